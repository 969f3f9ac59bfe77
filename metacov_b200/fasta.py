"""``FastaFile``: the subset of ``pysam.FastaFile`` the coverage path consumes (reference
metacov/cli.py:59 opens it, metacov/pileup.py:63 calls ``fetch(ref, start, end)``).

Host-side text parsing, outside the hot path: the whole file is read into memory (plain, gzip or
bgzip -- a BGZF file is a series of gzip members, which ``gzip`` reads as one stream).
"""
import gzip


class FastaFile:
    def __init__(self, filename):
        self.filename = filename
        with open(filename, "rb") as fh:
            magic = fh.read(2)
        opener = gzip.open if magic == b"\x1f\x8b" else open
        self._seqs = {}
        name, parts = None, []
        with opener(filename, "rt") as fh:
            for line in fh:
                if line.startswith(">"):
                    if name is not None:
                        self._seqs[name] = "".join(parts)
                    words = line[1:].split()
                    name, parts = (words[0] if words else ""), []
                elif name is not None:
                    parts.append(line.strip())
        if name is not None:
            self._seqs[name] = "".join(parts)
        self.references = tuple(self._seqs)
        self.lengths = tuple(len(v) for v in self._seqs.values())
        self.nreferences = len(self.references)

    def fetch(self, reference=None, start=None, end=None):
        """0-based half-open slice of a sequence; KeyError for an unknown name (pysam raises KeyError too)."""
        seq = self._seqs[reference]
        start = 0 if start is None else max(int(start), 0)
        end = len(seq) if end is None else int(end)
        return seq[start:end]

    def get_reference_length(self, reference):
        return len(self._seqs[reference])

    def close(self):
        self._seqs = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False
