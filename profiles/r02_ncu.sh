#!/bin/bash
# Round-2 profile set: launch list of the bench command + one full capture per dominant kernel.  Every command runs
# plain first (exit 0 required) and only then under ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for m in ${SD_NCU_MODES:-query:field_bin ssc:ssc_head render64:field_tc render768:field_tc expand:expand_tc}; do
  mode=${m%%:*}; pat=${m##*:}
  python profiles/run_r02.py $mode 3 > gpurun_out/plain_$mode.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -o gpurun_out/r02_prof_$mode -f python profiles/run_r02.py $mode 3 > gpurun_out/ncu_$mode.log 2>&1
  echo "$mode rc=$?"; tail -n 1 gpurun_out/ncu_$mode.log
done
# the four sort kernels of the query: one full capture each
python profiles/run_r02.py query 2 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:bin_ -s 4 -c 4 -o gpurun_out/r02_prof_sort -f python profiles/run_r02.py query 2 > gpurun_out/ncu_sort.log 2>&1
echo "sort rc=$?"
ls -la gpurun_out/*.ncu-rep
