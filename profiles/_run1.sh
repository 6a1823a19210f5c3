python -m pytest tests/test_gpu_binned.py tests/test_gpu_parity.py -x -q -k "binned or tile or ssc_grid or projected or sorted or graph" 2>&1 | tail -3
python profiles/time_bin.py 2>&1 | tail -1
