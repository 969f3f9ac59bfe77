"""`metacov pileup` end to end on a C3-like file (many small contigs): wall time of the CLI call, per decode mode."""
import sys, time, os, tempfile, json
sys.path.insert(0, ".")
import numpy as np
from click.testing import CliRunner
from metacov_b200 import synth
from metacov_b200.cli import pileup as cli_pileup
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
w = synth.c3(scale)
hb, isz = synth.generate_host(w)
tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "c3.bam")
synth.write_bam(path, w, hb, isz)
out = {"reads": int(len(hb.tid)), "contigs": int(w.n_contigs), "bam_bytes": os.path.getsize(path)}
texts = {}
for mode in ("host", "auto"):
    best = None
    for _ in range(2):
        o = os.path.join(tmp, "cov_%s.csv" % mode)
        t0 = time.perf_counter()
        res = CliRunner().invoke(cli_pileup, ["-b", path, "-o", o, "--bam-decode", mode])
        dt = time.perf_counter() - t0
        assert res.exit_code == 0, res.output
        best = dt if best is None else min(best, dt)
    texts[mode] = open(o).read()
    out[mode] = {"cli_s": best, "csv_rows": texts[mode].count("\n") - 1}
out["same_csv"] = texts["host"] == texts["auto"]
print(json.dumps(out))
