"""metacov_b200 -- B200-native coverage hot path of epruesse/metacov.

Drop-in surface (same names, arguments and error behaviour as the reference):
``metacov_b200.pileup.classic`` / ``experimental`` (reference metacov/pileup.py:9, 38),
``metacov_b200.scan.scan_reads`` (metacov/scan.pyx:623) and the ``metacov``
CLI (metacov/cli.py), over ``metacov_b200.AlignmentFile`` (the pysam protocol
subset the path consumes).  All compute runs in hand-written sm_100a CUDA
kernels behind the C-ABI of include/metacov_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"

from ._capi import McovError  # noqa: F401  (import fails loudly if the native library is missing)
from .alignmentfile import AlignmentFile  # noqa: F401
from .engine import CoverageEngine, ReadBatch  # noqa: F401
from .fasta import FastaFile  # noqa: F401
