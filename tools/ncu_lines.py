#!/usr/bin/env python
"""Executed warp instructions and stall samples of one kernel aggregated per CUDA source line.
usage: python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top_n]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
fname = "?"
hdr = None
seen_first_kernel = False
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if "Instructions Executed" in r:
        if hdr is not None and seen_first_kernel:
            pass
        hdr = r
        i_ln, i_src, i_addr, i_ie, i_s = 0, 1, hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= i_ie or r[i_addr] != "-" or not r[i_ln]:
        continue
    key = (fname, r[i_ln], r[i_src].strip()[:110])
    try:
        ie, ns = int(r[i_ie]), int(r[i_s])
    except ValueError:
        continue
    a = agg.setdefault(key, [0, 0, 0, {}])
    a[0] += ie; a[1] += ns; a[2] += 1
    for i, name in enumerate(hdr):
        if name.startswith("stall_") and "Not Issued" not in name:
            try:
                v = int(r[i])
            except ValueError:
                v = 0
            if v:
                a[3][name[6:]] = a[3].get(name[6:], 0) + v
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print("kernel %s: %d warp instructions, %d samples" % (pat, tot_i, tot_s))
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -(kv[1][1] if len(sys.argv) > 4 else kv[1][0]))[:top]:
    st = " ".join("%s=%d" % kv for kv in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s   [%s]" % (100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, f, ln, src[:70], st))
