python -m pytest tests/test_gpu_binned.py tests/test_gpu_ssc_head.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -4
SD_BENCH_VERBOSE=1 python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -5 gpurun_out/bench_r02b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02b.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', {k:d['roofline'][k] for k in ('frac','step_frac','kernel_ms','achieved')})
print('other', d['other_layout']); print('sorted_reuse', d['sorted_reuse'])
pf=d['per_frame']; print('per_frame', pf['ms'], pf['kernels_ms'], pf['ssc_head'])
for k,v in d['renders'].items(): print(k, v['ms'], v['msamples_per_s'], v['roofline']['frac'], v['e2e']['ms'])
print('fp32', d['fp32']); print('cpu', d['cpu_baseline'])
PY
