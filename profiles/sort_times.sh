# parity of the query paths, then per-kernel times of the texel sort (ncu launch list) on the SSC grid
timeout 300 python -m pytest tests/test_gpu_binned.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python profiles/time_bin.py 2>&1 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bin_ -c 40 --csv --log-file gpurun_out/sort_launches.csv python profiles/time_bin.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/sort_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:40]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, v in d.items(): print(k, len(v), "median ns:", sorted(v)[len(v)//2])
PY
if [ "$1" = full ]; then
# one full capture of the count and the scatter pass (warp-state samples: what they wait on)
timeout 300 ncu --set full --clock-control none -k regex:'bin_count|bin_scatter' -s 6 -c 2 -o gpurun_out/sort_full -f python profiles/time_bin.py > /dev/null 2>&1
ncu -i gpurun_out/sort_full.ncu-rep --page raw --csv > gpurun_out/sort_full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/sort_full_raw.csv")))
hdr = rows[0]
keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled")]
keys = [h for h in hdr if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d["Kernel Name"][:30], "us", d.get("gpu__time_duration.sum"), "issue%", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
          "warps_active%", d.get("sm__warps_active.avg.pct_of_peak_sustained_active"), "inst", d.get("smsp__inst_executed.sum"),
          "dram%", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
          "lsu_wave%", d.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
          "bank_conf", d.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"))
    top = sorted(((float(d[k] or 0), k.split("issue_stalled_")[1].split("_per")[0]) for k in keys), reverse=True)[:6]
    print("   stalls per issue:", [(n, round(v, 2)) for v, n in top])
PY
fi
