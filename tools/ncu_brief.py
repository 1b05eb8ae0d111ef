#!/usr/bin/env python
"""Brief text summary of one .ncu-rep (first captured kernel): key metrics, warp-stall breakdown, hottest SASS lines.
usage: python tools/ncu_brief.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
print(f"# {rep}  kernel: {m.get('Kernel Name', ('?',''))[0]}")
for k in ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
    if k in m:
        print(f"{k:75s} {m[k][0]:>16s} {m[k][1]}")
st = [(h, float(v.replace(",", ""))) for h, (v, u) in m.items()
      if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v not in ("", "n/a")]
tot = sum(v for _, v in st) or 1
print("\n# warp stall samples")
for h, v in sorted(st, key=lambda x: -x[1])[:10]:
    print(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', ''):30s} {v:10.0f} {100 * v / tot:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
isrc, iss, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
body = rows[2:]
tot = sum(int(r[iss]) for r in body) or 1
print("\n# SASS lines with > 1.5 % of the stall samples (line, samples, executed)")
for k, r in enumerate(body):
    if int(r[iss]) > 0.015 * tot:
        print(f"{k:5d} {r[isrc].strip()[:80]:80s} {int(r[iss]):8d} {100 * int(r[iss]) / tot:5.1f}% {r[iex]:>10s}")
