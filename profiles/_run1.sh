python -m pytest tests/test_gpu_parity.py tests/test_gpu_surface.py tests/test_gpu_zz_next.py -x -q -k "render or pass or surface or d768 or rays" 2>&1 | tail -3
python profiles/time_r02.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:round(v['ms'],3) for k,v in d.items() if isinstance(v,dict) and 'ms' in v})"
