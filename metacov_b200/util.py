"""Region sources of ``metacov pileup`` (same semantics as reference
metacov/util.py:27-83): regions come from a BLAST7 file, a CSV file, or -- when
neither is given -- one whole-contig region per @SQ line of the BAM."""
import csv
import datetime
from collections import namedtuple

import click

from . import blast

Region = namedtuple("Region", ["qacc", "sacc", "sstart", "send"])


class Timeit:
    """Context manager printing the wall time of its block (reference util.py:10-24)."""

    def __init__(self, message="timed: {elapsed}s"):
        self.message = message

    def __enter__(self):
        self.begin = datetime.datetime.now()
        return self

    def __exit__(self, exc_type, exc, tb):
        self.stamp()

    def stamp(self):
        print(self.message.format(elapsed=datetime.datetime.now() - self.begin))


def get_regions_from_blast7(regionfile):
    """BLAST hits as they come (they carry sacc / sstart / send; reference util.py:30-33)."""
    yield from blast.reader(regionfile)


_CSV_COLUMNS = (("sacc", "sequence_id"), ("sstart", "start"), ("send", "end", "stop"))


def get_regions_from_csv(regionfile):
    """Regions from a CSV with a contig column (sacc | sequence_id), a start column
    (sstart | start) and an end column (send | end | stop); reference util.py:36-61."""
    rows = csv.reader(regionfile)
    header = next(rows)
    picked = []
    for names in _CSV_COLUMNS:
        col = next((header.index(n) for n in names if n in header), None)
        if col is None:
            raise ValueError("Region file must have a column with a name in {}".format(names))
        picked.append(col)
    for row in rows:
        yield Region("", row[picked[0]], row[picked[1]], row[picked[2]])


def get_regions_from_bam(bamfile):
    """(n, name, 0, length) for every reference of the BAM (reference util.py:64-69)."""
    for n, (length, name) in enumerate(zip(bamfile.lengths, bamfile.references)):
        yield Region(n, name, 0, length)


def make_region_iterator(regionfile_blast7, regionfile_csv, bam):
    if regionfile_blast7 and regionfile_csv:
        raise click.BadParameter("Only one of regionfile-blast7 and regionfile-csv may be specified")
    if regionfile_blast7:
        return get_regions_from_blast7(regionfile_blast7)
    if regionfile_csv:
        return get_regions_from_csv(regionfile_csv)
    return get_regions_from_bam(bam)
