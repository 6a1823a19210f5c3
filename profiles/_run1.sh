timeout 300 python -m pytest tests/test_gpu_x3.py -x -q 2>&1 | grep -v "^$" | tail -25
