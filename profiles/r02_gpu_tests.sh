#!/bin/bash
# Round-2 GPU check: every test file in its own process (a trapped kernel poisons its CUDA context), logs to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in ${SD_R02_FILES:-test_gpu_zz_next test_gpu_ssc_head test_gpu_parity test_gpu_surface}; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; tail -n 3 gpurun_out/$f.log
done
if [ -n "$SD_R02_TIMINGS" ]; then
  timeout 600 python profiles/time_r02.py > gpurun_out/time_r02.json 2> gpurun_out/time_r02.err; echo "time rc=$?"; cat gpurun_out/time_r02.json
  SD_TC_HCOMP=1 timeout 600 python profiles/time_r02.py > gpurun_out/time_r02_hcomp.json 2> gpurun_out/time_r02_hcomp.err; echo "time hcomp rc=$?"; cat gpurun_out/time_r02_hcomp.json
fi
