python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python profiles/time_r02.py > gpurun_out/time_r02c.json 2> gpurun_out/time_r02c.err; cat gpurun_out/time_r02c.json; tail -2 gpurun_out/time_r02c.err
python profiles/time_bin.py 2>&1 | tail -1
