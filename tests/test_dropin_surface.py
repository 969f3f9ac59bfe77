"""The reference-facing Python surface (no GPU needed): region sources, flag table, accumulator
classes and their row formats (reference metacov/util.py, blast.py, scan.pyx)."""
import io

import numpy as np
import pytest

BLAST7 = """# TBLASTN 2.5.0+
# Query: demo
# Database: reference
# Fields: subject acc., s. start, s. end
# 4 hits found
ref1\t1\t425
ref2\t1\t575
ref2\t1\t300
ref2\t575\t301
"""


def test_blast7_regions():
    from metacov_b200 import blast, util
    hits = list(util.get_regions_from_blast7(io.StringIO(BLAST7)))
    assert [(h.sacc, h.sstart, h.send) for h in hits] == [("ref1", 1, 425), ("ref2", 1, 575), ("ref2", 1, 300),
                                                          ("ref2", 575, 301)]
    assert hits[0]._fields == ("sacc", "sstart", "send")
    with pytest.raises(ValueError):
        blast.reader(io.StringIO("no header here\n"))
    p = blast.reader(io.StringIO("# BLASTN 2.9\n# Fields: query acc., subject acc., % identity, evalue, weirdcol\n"
                                 "q1\ts1\t97.5\t1e-5\tx\n"))
    h = list(p)[0]
    assert (h.qacc, h.sacc, h.pident, h.evalue) == ("q1", "s1", 97.5, 1e-5) and p.get_fields()[-1] == "weirdcol"
    assert p.isfirsthit()


def test_csv_and_bam_regions():
    import click
    from metacov_b200 import util
    rows = list(util.get_regions_from_csv(io.StringIO("x,sequence_id,start,stop\n0,ctgA,5,90\n1,ctgB,7,3\n")))
    assert rows == [util.Region("", "ctgA", "5", "90"), util.Region("", "ctgB", "7", "3")]
    with pytest.raises(ValueError):
        list(util.get_regions_from_csv(io.StringIO("a,b,c\n1,2,3\n")))

    class Bam:
        references = ("r1", "r2 extra")
        lengths = (10, 20)
    assert list(util.get_regions_from_bam(Bam)) == [util.Region(0, "r1", 0, 10), util.Region(1, "r2 extra", 0, 20)]
    assert list(util.make_region_iterator(None, None, Bam))[1].send == 20
    with pytest.raises(click.BadParameter):
        util.make_region_iterator(io.StringIO(BLAST7), io.StringIO("sacc,start,end\n"), Bam)


def test_flag_table():
    from metacov_b200 import scan
    assert list(scan.Flags) == ["Paired", "PairedProperly", "Mapped", "MateMapped", "Readdir", "MateReaddir",
                                "IsRead1", "IsRead2", "Alignment", "QC", "Duplicate"]
    assert scan.Flags["Mapped"].flag == 0x4 and scan.Flags["Mapped"].name_true == "Unmapped"
    assert scan.FLAG_READ1.flag == 0x40 and scan.FLAG_DUP.name_false == "Singleton"
    assert scan.kmer_base2_to_ascii(0b11_10_01_00, 4) == "ACGT"


def test_accumulator_rows_and_byflag_tags():
    from copy import copy
    from metacov_b200 import scan
    h = scan.IsizeHist()
    assert h.counts.dtype == np.uint32 and len(h.counts) == 128 and h.max_isize == 0
    row = np.zeros(1024, np.uint32)
    row[[0, 3, 300]] = [7, 2, 1]
    h._ingest(row)
    assert h.max_isize == 300 and len(h.counts) == 512            # doubled from 128 until > 300
    rows = list(h.get_rows())
    assert rows[0] == ["n", "count"] and len(rows) == 302 and rows[4] == [3, 2]
    assert copy(h).max_isize == 0
    k = scan.KmerHist(2, 3, 2, 0)
    kr = list(k.get_rows())
    assert kr[0] == ["kmer", "n0", "n1", "n2"] and kr[1][0] == "NN" and kr[2][0] == "AA" and len(kr) == 18
    # ByFlag: copies and reversed tag columns (scan.pyx:393-403)
    bf = scan.ByFlag([scan.IsizeHist()], [scan.Flags["Mapped"], scan.Flags["IsRead1"]])
    assert len(bf.processors) == 4
    for n, p in enumerate(bf.processors):
        r = np.zeros(128, np.uint32)
        r[n] = n + 1
        p.processors[0]._ingest(r)
    rows = list(bf.get_rows(0))
    assert rows[0] == ["n", "count", "IsRead1", "Mapped"]
    assert rows[1] == [0, 1, "R2", "Mapped"]                        # n=0
    assert rows[2][:2] == [0, 0] and rows[3] == [1, 2, "R1", "Mapped"]          # n=1: bit0 = IsRead1
    assert rows[-1] == [3, 4, "R1", "Unmapped"]
    with pytest.raises(Exception, match="mah"):
        from metacov_b200 import AlignmentFile
        scan.scan_reads(object.__new__(AlignmentFile), None, 42)
    with pytest.raises(Exception, match="meh"):
        scan.scan_reads("not a bam", None, [])
