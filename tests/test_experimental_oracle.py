"""The oracle restatement of ``pileup.experimental`` (oracle/experimental.py) against vectors made
by the reference's own function (oracle/make_golden.py: experimental_fixture / experimental_synthetic)."""
import numpy as np
import pytest

from helpers import GOLD, assert_experimental_equal, fake_bam, load_json
from oracle import experimental as oexp

import os


@pytest.mark.parametrize("soa, js", [("fixture_soa.npz", "fixture_experimental.json"),
                                     ("synth_pairs_soa.npz", "synth_pairs_experimental.json")])
def test_oracle_experimental_matches_reference_vectors(soa, js):
    z = np.load(os.path.join(GOLD, soa), allow_pickle=False)
    gold = load_json(js)
    bam = fake_bam(z)
    k_cor = oexp.synthetic_kcor(gold["k_len"])
    n_raise = 0
    for row in gold["rows"]:
        fa = gold["fasta"][row["ref"]] if row["fasta"] else None
        where = (js, row["ref"], row["start"], row["end"])
        try:
            with np.errstate(all="ignore"):
                sums, res = oexp.experimental(bam.recs, bam.references, k_cor, gold["k_len"], fa, row["ref"], row["start"], row["end"])
        except Exception as ex:
            assert row.get("raises") == type(ex).__name__, where
            n_raise += 1
            continue
        assert "raises" not in row, where
        assert_experimental_equal(res, row["result"], where)
        assert sums["secondary"] + sums["nreads"] + sums["improper"] >= 0
    assert n_raise == sum(1 for r in gold["rows"] if "raises" in r)


def test_python_slice_semantics():
    for n in (1, 2, 7):
        a = list(range(n))
        for i in range(-3 * n, 3 * n):
            for j in range(-3 * n, 3 * n):
                assert oexp._py_slice_count(i, j, n) == len(a[i:j])
