python -m pytest tests/test_gpu_binned.py -q 2>&1 | grep -v "^$" | grep -B30 "Error\|^E " | head -80
