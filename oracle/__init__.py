"""CPU oracle for the metacov coverage hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package ``metacov_b200`` may
import, call, link or execute anything under ``oracle/``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm do,
and there only as the checker or the timed CPU baseline.

Parity status: **unpinned at the pysam/htslib boundary** -- the reference's own
tests assert only ``exit_code == 0`` for this path (reference
tests/test_cli.py:13-18,37,44,50) and pysam/htslib is absent from
/root/reference and from this image.  What *is* pinned: ``oracle.classic``
is checked against the reference's own ``metacov/pileup.py:classic`` imported
from /root/reference (see ``oracle/make_golden.py``), with both driven by the
restated boundary in ``oracle/pysam_boundary.py``; the resulting vectors are
committed under ``tests/golden/``.
"""
