"""The oracle (oracle/scan_oracle.c, plain-C restatement) against the reference: committed golden vectors produced by
the unmodified reference sources (tests/golden/make_golden.py) and, where oracle/_ref is present, fresh fuzz inputs."""
import os
import tempfile

import numpy as np
import pytest

import oracle_util as ou
from ribbit_b200 import synth


def _check_cp1(ev, cp1, crashed):
    n = len(cp1)
    if crashed:  # the reference died in its merge: what it logged is a prefix of the call sequence
        assert len(ev) >= n
        assert (ev[:n, :4] == cp1).all()
    else:
        assert len(ev) == n
        assert (ev[:, :4] == cp1).all()


def test_oracle_matches_golden_cp1(golden):
    assert len(golden) >= 18
    for name, g in golden.items():
        seq = g["seq"].tobytes()
        ev = ou.scan_events(seq, int(g["args"][0]), int(g["args"][1]))
        _check_cp1(ev, g["cp1"], int(g["rc"][0]) != 0)


def test_oracle_pack_matches_ascii(golden):
    for name, g in golden.items():
        seq = g["seq"].tobytes()
        hi, lo, nn = ou.pack(seq)
        a = np.frombuffer(seq.upper(), dtype=np.uint8)
        pos = np.arange(len(a))
        bit = lambda pl: (pl[pos >> 5] >> (pos & 31).astype(np.uint32)) & 1 if len(a) else np.zeros(0, np.uint32)
        assert ((bit(hi) == 1) == ((a == ord("G")) | (a == ord("T")))).all()
        assert ((bit(lo) == 1) == ((a == ord("C")) | (a == ord("T")))).all()
        assert ((bit(nn) == 1) == ~np.isin(a, np.frombuffer(b"ACGT", dtype=np.uint8))).all()


@pytest.mark.skipif(not ou.have_ref(), reason="oracle/_ref (the compiled reference) is only present in the build container")
def test_oracle_matches_reference_live_fuzz():
    rng = np.random.default_rng(77)
    for L, nd, mlo, mhi in [(700, 0.01, 2, 100), (2500, 0.0, 2, 40), (2500, 0.05, 1, 6), (6000, 0.003, 2, 100)]:
        seq = synth.fuzz_contig(rng, L, nd)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "x.fa")
            synth.write_fasta(fa, [seq])
            contigs, _, rc = ou.ref_cp(fa, ["-m", mlo, "-M", mhi], stop_after_cp2=True)
        _check_cp1(ou.scan_events(seq, mlo, mhi), contigs[0]["cp1"], rc != 0)


def test_integer_cutoff_formula_equals_the_reference_double_arithmetic():
    """parse_anchored_shiftxor.cpp:572-573 computes int(0.9 * m) in double; the kernels use 9 * m // 10 (scan_core.h cut_anch)."""
    for m in range(10, 2001):
        assert int(0.9 * m) == (9 * m) // 10, m
