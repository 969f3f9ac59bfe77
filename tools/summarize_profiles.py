#!/usr/bin/env python
"""Turn the raw captures of tools/final_profile.sh (gpurun_out/final_TAG/) into the tracked summaries under profiles/:
  TAG_launches.csv / TAG_launch_shares.txt   ncu launch list of the bench command and the kernels' shares of a step
  TAG_ncu_full_summary.txt                   selected ncu --set full metrics + stall breakdown per kernel
  TAG_traffic_c2.json                        measured DRAM bytes per launch (what bench.py reports as roofline.traffic)
usage: python tools/summarize_profiles.py TAG"""
import collections, csv, io, json, os, re, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = "gpurun_out/final_%s" % tag
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)

# ---- launch list ----
rows = [r for r in csv.reader(open(os.path.join(src, "launches.csv"))) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
per = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik]
    if "mcov::" not in name:
        continue
    per.setdefault(name, []).append(float(r[iv].replace(",", "")))
shutil.copy(os.path.join(src, "launches.csv"), "profiles/%s_launches.csv" % tag)
tot = sum(sum(v) / len(v) for v in per.values())
bench = json.loads(open(os.path.join(src, "bench_c2.json")).read().strip().splitlines()[-1])
kern = bench["roofline"]["kernels"]
ksum = sum(v["ms_per_launch"] for v in kern.values())
with open("profiles/%s_launch_shares.txt" % tag, "w") as fh:
    fh.write("ncu --metrics gpu__time_duration.sum --clock-control none (profiles/%s_launches.csv), bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong, final build of round 2\n" % tag)
    fh.write("mean ns per launch and share among the mcov kernels of one step (cold-cache, serialised: compare shares):\n")
    for k, v in per.items():
        m = sum(v) / len(v)
        fh.write("  %-62s n=%d mean=%8.0f ns share=%.3f\n" % (k[:62], len(v), m, m / tot))
    fh.write("CUDA-event shares of the same kernels among themselves (bench.py default run, profiles/%s_bench_c2.json): " % tag)
    fh.write(", ".join("%s %.3f" % (k, v["ms_per_launch"] / ksum) for k, v in kern.items()) + "\n")

# ---- ncu --set full ----
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
traffic = {}
out = []
for rep in ("full.ncu-rep", "full_blk.ncu-rep"):
    path = os.path.join(src, rep)
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, units = rr[0], rr[1]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")]
        out.append("== %s" % name)
        for w in WANT:
            if w in h:
                out.append("   %-62s %s %s" % (w, r[h.index(w)], units[h.index(w)]))
        stalls = {}
        for i, c in enumerate(h):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", c)
            if m and r[i]:
                try:
                    stalls[m.group(1) or m.group(2)] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        if stalls:
            t = sum(stalls.values()) or 1.0
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:7]
            out.append("   stalls: " + ", ".join("%s %d%%" % (k, round(100 * v / t)) for k, v in top))
        def num(metric):
            v = float(r[h.index(metric)].replace(",", ""))
            u = units[h.index(metric)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        short = name.split("(")[0].split("::")[-1]
        traffic[short] = int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
open("profiles/%s_ncu_full_summary.txt" % tag, "w").write("\n".join(out) + "\n")
json.dump({"source": "ncu --set full --clock-control none (tools/final_profile.sh %s -> profiles/%s_ncu_full_summary.txt), C2, one launch each; "
                     "dram__bytes_read.sum + dram__bytes_write.sum" % (tag, tag),
           "note": "k_fused_tile_tma: part of its 200 MB of depth is still dirty in the 126 MB L2 when the kernel ends",
           "dram_bytes_per_launch": traffic}, open("profiles/%s_traffic_c2.json" % tag, "w"), indent=1)
for f, t in (("bench_c2.json", "bench_c2.json"), ("reference_arm_c2.json", "reference_arm_c2.json"), ("bench_bam.json", "bench_bam_decode.json")):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), "profiles/%s_%s" % (tag, t))
print(open("profiles/%s_launch_shares.txt" % tag).read())
print(json.dumps(traffic))
