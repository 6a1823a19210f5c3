#!/usr/bin/env python
"""Turns an ncu report (gpurun_out/*.ncu-rep) and a launch list (ncu --csv log) into the small text
summaries committed under profiles/.   python profiles/summarize.py <rep> <launches.csv> <out.md>"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]


def main():
    rep, launches, out = sys.argv[1:4]
    lines = ["# ncu summary", "", f"report: `{rep}` (ncu --set full --clock-control none), first captured launch", ""]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    lines += [f"kernel: `{r[idx['Kernel Name']]}`", "", "| metric | unit | value |", "|---|---|---|"]
    for k in KEYS:
        if k in idx:
            lines.append(f"| {k} | {units[idx[k]]} | {r[idx[k]]} |")
    rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
    lines += ["", f"traffic per launch (dram read + write): {rd + wr:.1f} {units[idx['dram__bytes_read.sum']]}", ""]
    if launches == "-":
        open(out, "w").write("\n".join(lines) + "\n")
        return
    lines += ["## launch list (ncu --metrics gpu__time_duration.sum, cold-cache serialised: compare shares)", ""]
    tot = collections.Counter(); cnt = collections.Counter()
    with open(launches) as f:
        rows = [x for x in csv.reader(l for l in f if not l.startswith("=="))]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    for x in rows[1:]:
        if len(x) <= vi:
            continue
        v = float(x[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(x[ui], 1.0)
        name = x[ki].split("(")[0][:90]
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    lines += ["| kernel | launches | total us | share |", "|---|---|---|---|"]
    for name, v in tot.most_common():
        lines.append(f"| {name} | {cnt[name]} | {v:.1f} | {100 * v / s:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
