timeout 600 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | grep -v "^$" | tail -30
