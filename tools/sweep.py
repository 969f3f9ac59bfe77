"""Run bench.py once per prebuilt library variant (tools/bin/lib_*.so) and print the per-kernel times.
Usage on the GPU box: python tools/sweep.py NAME[,NAME..] [bench args...]"""
import json, os, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "metacov_b200", "libmetacov_b200.so")
keep = lib + ".keep"
shutil.copy(lib, keep)
try:
    for name in sys.argv[1].split(","):
        shutil.copy(os.path.join(root, "tools", "bin", "lib_%s.so" % name), lib)
        os.utime(lib)
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu", "--no-e2e"] + sys.argv[2:],
                             capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            ks = {k: round(v["ms_per_launch"] * 1000, 1) for k, v in d["roofline"]["kernels"].items()}
            print("variant", name, "ms/step %.4f" % d["ms_per_step"], ks, flush=True)
        except Exception as e:
            print("variant", name, "FAILED", e, out.stderr[-800:], flush=True)
finally:
    shutil.move(keep, lib)
