"""The BASELINE.json workloads as seeded synthetic inputs, and the parity gate `bench.py` applies to its own results.

C3 (BASELINE.json configs[2], the configuration `metric` is quoted on): 24 contigs with hg38-like lengths, 3.1 Gbp,
contig i = synth.contig_c2(HG38_MBP[i] Mbp, seed = 100 + i).

Parity gate: the kept candidates (the calls of addSeedToSeedPositions* that reach the consumer's length cutoff,
parse_substitute_shiftxor.cpp:44, parse_anchored_shiftxor.cpp:153) of a region of a contig, in call order, digested with
md5. tests/golden/c3_digests.json holds the digests of the UNMODIFIED reference (oracle/_ref/ribbit_ref_cp) for
  * one 1 Mbp window of every contig (the reference run on the window as a contig of its own; candidates that lie
    GATE_MARGIN bases inside the window do not depend on what is outside it), and
  * the whole chr21-size contig (index 20, 47 Mbp),
written by tests/golden/make_golden_c3.py. Nothing here reads /root/reference or runs the oracle.
"""
import hashlib
import json
import os

import numpy as np

from . import synth

HG38_MBP = synth.HG38_MBP
C3_FULL_CONTIG = 20          # the chr21-size contig (47 Mbp), gated at full length
GATE_WINDOW = 1_000_000
GATE_MARGIN = 4096
DIGESTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c3_digests.json")


def c3_lengths(scale=1.0):
    return [int(mb * 1e6 * scale) for mb in HG38_MBP]


def c3_contig(i, scale=1.0):
    return synth.contig_c2(c3_lengths(scale)[i], seed=100 + i)


def c3_window(i, L):
    """[lo, hi) of the gated window of contig i: over the end of the leading N run, over the start of the 3 Mb N run at
    0.4 L, or in plain sequence, by i mod 3."""
    if L <= GATE_WINDOW:
        return 0, L
    k = i % 3
    lo = 40_000 if k == 0 else (int(0.4 * L) - 600_000 if k == 1 else int(0.7 * L))
    lo = max(0, min(lo, L - GATE_WINDOW))
    return lo, lo + GATE_WINDOW


def cut_subst(m):
    """parse_substitute_shiftxor.cpp:423"""
    return m // 3 if m > 30 else 10


def cut_anch(m):
    """parse_anchored_shiftxor.cpp:572-573"""
    c = m if m > 6 else 10
    if m >= 10:
        c = int(0.9 * m)
    return c


_CUTS = None


def kept_mask(stream, start, end, mlen):
    """stream 1..3; arrays of the raw calls -> bool mask of the calls that reach the consumer's cutoff."""
    global _CUTS
    if _CUTS is None:
        _CUTS = {1: np.zeros(1100, np.int64), 2: np.array([cut_subst(m) for m in range(1100)]),
                 3: np.array([cut_anch(m) for m in range(1100)])}
    return (np.asarray(end, np.int64) - np.asarray(start, np.int64)) >= _CUTS[stream][np.asarray(mlen, np.int64)]


def digest_rows(start, end, mlen, lo=None, hi=None, margin=GATE_MARGIN):
    """md5 over the (start, end, mlen) int32 rows, in order; with a window only the rows margin bases inside it."""
    start = np.asarray(start, np.int64); end = np.asarray(end, np.int64); mlen = np.asarray(mlen, np.int64)
    if lo is not None:
        sel = (start >= lo + margin) & (end <= hi - margin)
        start, end, mlen = start[sel], end[sel], mlen[sel]
    rows = np.stack([start, end, mlen], axis=1).astype("<i4") if len(start) else np.zeros((0, 3), "<i4")
    return [hashlib.md5(rows.tobytes()).hexdigest(), int(len(rows))]


def load_digests(path=DIGESTS):
    with open(path) as f:
        return json.load(f)


def gate_contig(i, L, streams, golden):
    """streams: {0|1|2: int rows (start, end, mlen, flags)} of contig i. Returns a list of mismatch descriptions
    (empty = parity ok): the window digest and, for the contig gated at full length, the whole-contig digest."""
    bad = []
    lo, hi = c3_window(i, L)
    want = golden["windows"][str(i)]
    assert want["L"] == L and want["lo"] == lo and want["hi"] == hi, "golden made for another workload"
    for s in range(3):
        r = streams[s]
        k = (r[:, 3] & 3) == 0  # neither DROPPED nor PSEUDO
        got = digest_rows(r[k, 0], r[k, 1], r[k, 2], lo, hi)
        if got != want["kept"][str(s + 1)]:
            bad.append("contig %d window [%d,%d) stream %d: %s != reference %s" % (i, lo, hi, s + 1, got, want["kept"][str(s + 1)]))
        if i == golden["full"]["contig"]:
            gotf = digest_rows(r[k, 0], r[k, 1], r[k, 2])
            if gotf != golden["full"]["kept"][str(s + 1)]:
                bad.append("contig %d full length stream %d: %s != reference %s" % (i, s + 1, gotf, golden["full"]["kept"][str(s + 1)]))
    return bad
