"""CPU restatement of the read-statistics accumulators of metacov/scan.pyx (oracle; test
infrastructure only).  Pure Python loops: for fixture-sized inputs.

Follows reference metacov/scan.pyx:240-259 (`get_seq`: nt16 -> nt4, reverse-strand records
reverse-complemented back), 406-420 (`ByFlag` group index), 503-522 (`KmerHist.process_read`).
PARITY UNPINNED by the reference's tests (tests/test_cli.py only checks exit codes).
"""
import numpy as np

NT16_TO_NT4 = np.array([4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4], dtype=np.int64)   # scan.pyx:27


def nt4_comp(n):
    return np.where(n <= 3, 3 - n, 4)                                                        # scan.pyx:61-62


def get_seq(seq_nt16, flag):
    """Read in sequenced orientation, nt4 coded (scan.pyx:240-259)."""
    s = NT16_TO_NT4[np.asarray(seq_nt16, dtype=np.int64)]
    if flag & 0x10:
        s = nt4_comp(s)[::-1]
    return s


def group_index(flag, group_flags):
    n = 0
    for f in group_flags:                       # first selected flag = most significant bit
        n <<= 1
        if flag & f:
            n += 1
    return n


def kmer_hist(flags, seqs, K, NK, STEP, OFFSET, group_flags=()):
    """counts[group][kmer][i] as KmerHist fills them (scan.pyx:503-522)."""
    table = 4 ** K + 1
    out = np.zeros((1 << len(group_flags), table, NK), dtype=np.uint32)
    for flag, seq in zip(flags, seqs):
        rlen = len(seq)
        if rlen < OFFSET + STEP * NK:
            continue
        read = get_seq(seq, int(flag))
        g = group_index(int(flag), group_flags)
        for i in range(NK):
            k = 0
            for j in range(K):
                x = OFFSET + i * STEP + j
                c = int(read[x]) if x < rlen else 4      # the reference would read past the end here
                if c > 3:
                    k = 4 ** K
                    break
                k |= c << (2 * j)
            out[g, k, i] += 1
    return out
