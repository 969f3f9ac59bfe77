"""BLAST tabular-with-comments (outfmt 7) reader -> region records.

Same behaviour as reference metacov/blast.py:48-94 (`reader`, `fmt7_parser`):
the first line must mention BLAST, the ``# Fields:`` comment names the columns
(long names mapped to the usual short ones), numeric columns are converted, and
every hit line becomes a namedtuple ``BlastHit``.
"""
from collections import namedtuple

_SHORT = {
    "query acc.": "qacc", "subject acc.": "sacc", "% identity": "pident", "alignment length": "length",
    "mismatches": "mismatch", "gap opens": "gapopen", "q. start": "qstart", "q. end": "qend",
    "s. start": "sstart", "s. end": "send", "evalue": "evalue", "bit score": "bitscore",
    "subject strand": "sstrand", "sbjct frame": "sframe", "score": "score",
}
_CAST = {
    "pident": float, "length": int, "mismatch": int, "gapopen": int, "qstart": int, "qend": int,
    "sstart": int, "send": int, "evalue": float, "bitscore": float, "score": float, "sframe": int,
}


class Fmt7Reader:
    """Iterates the hits of one BLAST7 stream."""

    def __init__(self, fileobj):
        self.fileobj = fileobj
        self.fields = None
        self.query = None
        self.hits = None
        self.hit = 0
        self._tuple = None
        if "BLAST" not in fileobj.readline():
            raise ValueError("not a BLAST7 formatted file")

    def get_fields(self):
        return self.fields

    def isfirsthit(self):
        return self.hit == 1

    def __iter__(self):
        for line in self.fileobj:
            if line.startswith("# Fields: "):
                names = line[len("# Fields: "):].strip().split(", ")
                self.fields = [_SHORT.get(n, n) for n in names]
                self._tuple = namedtuple("BlastHit", self.fields)
            elif line.startswith("# Query: "):
                self.query, self.hit = line[len("# Query: "):].strip(), 0
            elif line.startswith("# Database: "):
                self.query, self.hit = line[len("# Database: "):].strip(), 0
            elif line.strip().endswith(" hits found"):
                self.hits, self.hit = int(line.split()[1]), 0
            elif line.startswith("#"):
                continue
            else:
                self.hit += 1
                cells = line.strip().split("\t")
                yield self._tuple(*[_CAST[k](v) if k in _CAST else v for k, v in zip(self.fields, cells)])


fmt7_parser = Fmt7Reader


def reader(fileobj, t=7):
    if t == 7:
        return Fmt7Reader(fileobj)
    raise ValueError("other formats not implemented")
