#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics, stall reasons, and the hottest SASS lines.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex-for-source-page ...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")])
    for w in want:
        if w in hdr:
            print("   %-62s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print("   stalls: " + ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:7]))
for pat in sys.argv[2:]:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    i_src, i_s = hdr.index("Source"), hdr.index("# Samples")
    data = [r for r in rows[2:] if len(r) > i_s]
    # the page may repeat the kernel: keep the first instance
    first = []
    for r in data:
        if first and r[hdr.index("Address")] == first[0][hdr.index("Address")]:
            break
        first.append(r)
    def n(r):
        try:
            return int(r[i_s])
        except ValueError:
            return 0
    tot = sum(n(r) for r in first) or 1
    print("== hottest SASS lines of", pat, "(%d samples, %d instructions)" % (tot, len(first)))
    order = {id(r): k for k, r in enumerate(first)}
    for r in sorted(sorted(first, key=lambda r: -n(r))[:24], key=lambda r: order[id(r)]):
        print("   %5d %5.1f%%  %s" % (order[id(r)], 100.0 * n(r) / tot, r[i_src].strip()[:100]))
