"""BASELINE.json configs at FULL size on one B200 (C3 200 M reads / 500 k contigs, C4 1 B reads /
100 k contigs, C5 5 M long reads with 1.4e10 CIGAR ops -- 64-bit offsets): the oracle cannot finish
these in seconds, so parity is checked the way SURVEY.md 8(d) "Parity checks at scale" lays out:

  * size-independent properties of the whole result: mass conservation (sum of depth == sum of the
    per-contig `sum` records == aligned bases), the two independent GPU formulations (fused sorted
    path; atomics + decoupled-look-back scan of the push path) agree on every slot, idempotence;
  * 1 000 sampled contigs re-derived ON THE CPU from the counter-based generator and run through
    the C oracle: depth bit-exact, statistics records equal.

Part of the regular GPU suite (a case takes seconds on a B200); needs ~100 GB of HBM, so devices with
less than 120 GB skip it.  MCOV_FULLSIZE_OUT names a file the outcome is appended to (committed under
profiles/).
"""
import json
import os
import time

import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu

CASES = [("c3", 1.0), ("c4", 1.0), ("c5", 1.0)]
if os.environ.get("MCOV_FULLSIZE_CASES"):
    CASES = [(c.split(":")[0], float(c.split(":")[1])) for c in os.environ["MCOV_FULLSIZE_CASES"].split(",")]


def _host_sample(w, contigs):
    """Reads of the sampled contigs, regenerated on the host, as one batch with tids 0..S-1."""
    from metacov_b200 import ReadBatch, synth
    parts = []
    ops = 0
    for k, c in enumerate(contigs):
        i0, i1 = int(w.read_start[c]), int(w.read_start[c + 1])
        b, _ = synth.generate_host(w, i0=i0, n=i1 - i0, tid_base=int(c) - k)
        assert i1 == i0 or (b.tid[0] == k and b.tid[-1] == k)
        parts.append((b, ops))
        ops += len(b.cig)
    cat = lambda f: np.concatenate([getattr(b, f) for b, _ in parts])
    cig_off = np.concatenate([b.cig_off[:-1].astype(np.int64) + o for b, o in parts] + [np.array([ops], np.int64)])
    return ReadBatch(cat("tid"), cat("pos"), cat("flag"), cat("mapq"), cig_off.astype(np.uint32), cat("cig"))


@pytest.mark.parametrize("wl,scale", CASES)
def test_full_size_properties_and_sampled_contigs(wl, scale):
    import torch
    from metacov_b200 import CoverageEngine, synth
    if torch.cuda.get_device_properties(0).total_memory < 120 * 2 ** 30:
        pytest.skip("full-size configs need a device with at least 120 GB (B200: 180 GB)")
    torch.cuda.empty_cache()
    t_start = time.time()
    w = synth.WORKLOADS[wl](scale)
    db, _ = synth.generate_device(w, 0)
    lengths = w.contig_len
    out = {"workload": w.describe(), "scale": scale}
    with CoverageEngine(lengths) as eng:
        buf = torch.empty(eng.n_slots, dtype=torch.int32, device="cuda")
        eng.bind_depth(buf)
        eng.depth_sorted(db)
        pi = eng.pass_info()
        assert pi["cap_contigs"] == 0 and pi["max_depth_seen"] <= 8000
        fused = buf.clone()
        # idempotence
        eng.depth_sorted(db)
        torch.cuda.synchronize()
        assert torch.equal(buf, fused)
        # per-contig records of every contig
        tid = np.arange(w.n_contigs, dtype=np.int32)
        st = eng.region_stats(tid, np.zeros_like(tid), lengths)
        # mass conservation (no read of these workloads is clipped at a contig end)
        total = int(fused.sum(dtype=torch.int64).item())
        assert total == pi["aligned_bases"] == int(st["sum"].astype(np.int64).sum())
        sq = int((fused.to(torch.int64) ** 2).sum().item()) if eng.n_slots < 2 ** 31 else None
        if sq is not None:
            assert sq == int(st["sumsq"].astype(np.uint64).sum())
        assert int(fused.max().item()) == int(st["max"].max()) == pi["max_depth_seen"]
        # path equivalence: clear + atomics + look-back scan (push path)
        eng.begin()
        eng.push(db)
        eng.finalize()
        torch.cuda.synchronize()
        assert torch.equal(buf, fused), "push path differs from the fused path"
        pi2 = eng.pass_info()
        assert pi2["n_pass"] == pi["n_pass"] and pi2["aligned_bases"] == pi["aligned_bases"]
        out.update(n_pass=pi["n_pass"], aligned_bases=pi["aligned_bases"], max_depth_seen=pi["max_depth_seen"],
                   slots=int(eng.n_slots), depth_sum=total)
        # sampled contigs against the C oracle, reads re-derived on the CPU from the generator
        rng = np.random.Generator(np.random.PCG64(7))
        S = min(1000, w.n_contigs)
        contigs = np.sort(rng.choice(w.n_contigs, S, replace=False))
        hb = _host_sample(w, contigs)
        sl = lengths[contigs]
        want, off, info = cport.depth(hb, sl, mode="par", threads=os.cpu_count() or 4)
        ref = cport.region_stats(want, off, sl, np.arange(S, dtype=np.int32), np.zeros(S, np.int32), sl,
                                 threads=os.cpu_count() or 4)
        bad = 0
        for k, c in enumerate(contigs):
            o = int(eng.contig_offset(int(c)))
            got = fused[o:o + int(sl[k])].cpu().numpy()
            bad += 0 if np.array_equal(got, want[off[k]:off[k] + sl[k]]) else 1
        assert bad == 0, "%d of %d sampled contigs differ from the oracle" % (bad, S)
        for key in ("sum", "sumsq", "iq_sum", "min", "max", "med_lo", "med_hi", "n_ge1"):
            assert np.array_equal(st[key][contigs], ref[key]), key
        out.update(sampled_contigs=int(S), sampled_reads=int(len(hb.tid)), sampled_bit_exact=True,
                   checks=["idempotence", "sum(depth)==aligned_bases==sum(records.sum)", "sumsq", "max",
                           "push path == fused path on every slot", "sampled contigs depth+records == C oracle"],
                   seconds=round(time.time() - t_start, 1))
    dst = os.environ.get("MCOV_FULLSIZE_OUT")
    if dst:
        with open(dst, "a") as fh:
            fh.write(json.dumps(out) + "\n")
