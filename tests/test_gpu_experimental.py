"""``pileup.experimental`` on the GPU (mcov_experimental_run / mcov_exp_revsum through the drop-in
API) against (a) vectors made by the reference's own function and (b) the oracle's sums."""
import os

import numpy as np
import pytest

from helpers import GOLD, assert_experimental_equal, fake_bam, load_json
from oracle import bamio
from oracle import experimental as oexp

pytestmark = pytest.mark.gpu


class Fasta:
    """pysam.FastaFile protocol subset (reference cli.py:59, pileup.py:63)."""

    def __init__(self, seqs):
        self.seqs = seqs

    def fetch(self, ref, start, end):
        return self.seqs[ref][start:end]


def encode_bam(tmp, soa):
    z = np.load(os.path.join(GOLD, soa), allow_pickle=False)
    path = str(tmp / (soa + ".bam"))
    so = z["seq_off"]
    n = len(z["tid"])
    seqs = [z["seq"][so[i]:so[i + 1]] for i in range(n)]
    bamio.write_bam(path, [str(x) for x in z["references"]], z["lengths"].tolist(), z["tid"], z["pos"], z["flag"], z["mapq"],
                    z["cig_off"].astype(np.uint32), z["cig"], l_seq=z["l_seq"], isize=z["isize"],
                    names=[str(x) for x in z["names"]], seqs=seqs)
    return z, path


CASES = [("fixture_soa.npz", "fixture_experimental.json"), ("synth_pairs_soa.npz", "synth_pairs_experimental.json")]


@pytest.mark.parametrize("soa, js", CASES)
def test_experimental_matches_reference_vectors(soa, js, tmp_path):
    from metacov_b200 import AlignmentFile, pileup
    z, path = encode_bam(tmp_path, soa)
    gold = load_json(js)
    k_cor = oexp.synthetic_kcor(gold["k_len"])
    fasta = Fasta(gold["fasta"])
    with AlignmentFile(path) as bam:
        for row in gold["rows"]:
            where = (js, row["ref"], row["start"], row["end"])
            try:
                with np.errstate(all="ignore"):
                    res = pileup.experimental(bam, k_cor, gold["k_len"], fasta if row["fasta"] else None,
                                              row["ref"], row["start"], row["end"])
            except Exception as ex:
                assert row.get("raises") == type(ex).__name__, (where, ex)
                continue
            assert "raises" not in row, where
            assert_experimental_equal(res, row["result"], where)
            # value types follow the reference (np.mean -> np.float64, round(np.float64) -> int)
            assert isinstance(res["cov"], np.float64) and isinstance(res["cov2"], int)


@pytest.mark.parametrize("soa, js", CASES)
def test_experimental_sums_match_oracle_and_batching(soa, js, tmp_path):
    """The per-region sums (mcov_exp_stats): integers identical to the oracle's, float sums to 1e-9;
    one batched call equals region-by-region calls bit for bit (fixed-order reductions)."""
    from metacov_b200 import AlignmentFile, pileup
    z, path = encode_bam(tmp_path, soa)
    gold = load_json(js)
    k_len = gold["k_len"]
    k_cor = oexp.synthetic_kcor(k_len)
    obam = fake_bam(z)
    rows = gold["rows"]
    refs = [r["ref"] for r in rows]
    starts = [r["start"] for r in rows]
    ends = [r["end"] for r in rows]
    with AlignmentFile(path) as bam:
        kc_val, kc_has = pileup._kcor_tables(k_cor, k_len)
        st = bam.experimental_stats(refs, starts, ends, k_len, kc_val, kc_has)
        for i, r in enumerate(rows):
            want = oexp.region_sums(obam.recs, list(obam.references).index(r["ref"]), r["start"], r["end"], k_cor, k_len)
            for k in ("cov_sum", "cov2_sum", "n_starts", "nreads", "secondary", "improper", "no_reflen", "n_pairs"):
                assert int(st[i][k]) == want[k], (js, i, k, st[i][k], want[k])
            for k in ("covw_sum", "cor_sum", "wnf_sum"):
                assert abs(float(st[i][k]) - want[k]) <= 1e-9 * max(1.0, abs(want[k])), (js, i, k)
            one = bam.experimental_stats([r["ref"]], [r["start"]], [r["end"]], k_len, kc_val, kc_has)[0]
            assert one.tobytes() == st[i].tobytes(), (js, i)
        # no k-mer table at all: every lookup is a KeyError -> weights 1
        st0 = bam.experimental_stats(refs, starts, ends, k_len, None, None)
        for i in range(len(rows)):
            assert float(st0[i]["covw_sum"]) == float(st0[i]["cov_sum"])
            assert float(st0[i]["wnf_sum"]) == float(st0[i]["n_pairs"])
            assert float(st0[i]["cor_sum"]) == float(st0[i]["n_starts"])


def test_experimental_argument_errors(tmp_path):
    from metacov_b200 import AlignmentFile, pileup
    _, path = encode_bam(tmp_path, "fixture_soa.npz")
    with AlignmentFile(path) as bam:
        with pytest.raises(Exception, match="Length must be > 0"):        # reference pileup.py:40-41
            pileup.experimental(bam, None, 7, None, "ref1", 5, 5)
        with pytest.raises(KeyError):
            pileup.experimental(bam, None, 7, None, "nope", 0, 5)
    with pytest.raises(TypeError):
        pileup.experimental(object(), None, 7, None, "ref1", 0, 5)


def test_revsum_kernel_matches_numpy():
    from metacov_b200 import CoverageEngine
    rng = np.random.default_rng(5)
    eng = CoverageEngine([10])
    for L, n_w in ((1, 900), (37, 900), (2000, 900), (1500, 3)):
        cor_rev = rng.random(L)
        w = rng.random(n_w)
        got = eng.exp_revsum(cor_rev, w)
        want = np.array([np.dot(w[:min(L - i, n_w)], cor_rev[i:i + min(L - i, n_w)]) for i in range(L)])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    eng.close()
