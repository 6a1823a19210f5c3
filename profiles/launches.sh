#!/bin/bash
# launch list of the projected-map SSC query (cold-cache serialised times): profiles/launches.sh <tag>
python profiles/run_bin.py > gpurun_out/plain_$1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$1.csv python profiles/run_bin.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$1.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[-8:]:
    print(r[ki][:60], r[vi])
PY
