set -x
python -m pytest tests/test_gpu_ssc_head.py -x -q 2>&1 | tail -3
python profiles/time_r02.py > gpurun_out/time_r02b.json 2> gpurun_out/time_r02b.err
python profiles/dev_cfg.py cfg4 > gpurun_out/dev_cfg4.log 2>&1
python profiles/dev_cfg.py cfg1 > gpurun_out/dev_cfg1.log 2>&1
