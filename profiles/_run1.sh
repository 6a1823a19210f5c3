python -m pytest tests/test_gpu_surface.py -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches2.log 2>&1; echo "launch list rc=$?"
for m in binned ssc x3 expand; do
  k=field_bin; [ $m = ssc ] && k=ssc_head; [ $m = expand ] && k=expand_tc
  python profiles/run_r02.py $m 3 > gpurun_out/plain2_$m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r02f_prof_$m -f python profiles/run_r02.py $m 3 > gpurun_out/ncu2_$m.log 2>&1; echo "$m rc=$?"
done
ls -la gpurun_out/r02f_prof_*.ncu-rep
