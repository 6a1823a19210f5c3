#!/bin/bash
# Round-2 bench run on one GPU: drop-in tests, the bench line, then the ncu launch list of the same command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in ${SD_R02_FILES:-test_gpu_dropin}; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; tail -n 3 gpurun_out/$f.log
done
SD_BENCH_VERBOSE=1 timeout 900 python bench.py --steps ${SD_STEPS:-20} --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
tail -n 5 gpurun_out/bench_r02.err; cat gpurun_out/bench_r02.json
