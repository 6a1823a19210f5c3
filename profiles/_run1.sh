python -m pytest tests/test_gpu_binned.py -x -q -k "gen_voxel or btsnet" 2>&1 | tail -3
SD_BENCH_VERBOSE=1 python bench.py --no-render --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -3 gpurun_out/bench_r02c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'])
print('roofline', {k:d['roofline'][k] for k in ('frac','step_frac','kernel_ms')})
pf=d['per_frame']; print('per_frame', pf['ms'], pf['kernels_ms'])
PY
