timeout 300 python -m pytest tests/test_gpu_x3.py -x -q 2>&1 | grep -v "^$" | tail -15
SD_BENCH_VERBOSE=1 python bench.py --no-render --no-cpu-baseline > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err; tail -3 gpurun_out/bench_r02d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02d.json'))
print({k:d[k] for k in ('value','ms_per_step')}); print(d['fp32']); print(d['fp32_tc'])
PY
