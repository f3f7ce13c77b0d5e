#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics of each kernel + instructions / stall samples per source line.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rows[2:]:
    for k in KEYS:
        if k in hdr:
            print("%-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]]
    print("stalls/issue:", ", ".join("%s %.2f" % (h.split("issue_stalled_")[1].split("_per_")[0], v) for v, h in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
agg, samp = collections.Counter(), collections.Counter()
cur = hdr = None
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    key = (cur, int(r[0]))
    try: agg[key] += int(d.get("Instructions Executed") or 0)
    except ValueError: pass
    try: samp[key] += int(d.get("# Samples") or 0)
    except ValueError: pass
ti, ts = max(sum(agg.values()), 1), max(sum(samp.values()), 1)
print("source lines by stall samples (inst%% / samples%%), total inst %d samples %d" % (ti, ts))
for k in sorted(agg, key=lambda k: -samp[k])[:top]:
    print("  %5.2f%% inst %5.2f%% samp  %s:%d" % (100 * agg[k] / ti, 100 * samp[k] / ts, k[0], k[1]))
