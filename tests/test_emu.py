"""Kernel lane logic (ribbit_b200/csrc/{scan_core,merge_core,layout}.h) run on the CPU warp emulator, against the
oracle and the golden vectors — including chunking with warm-up restarts and the ordered compaction."""
import numpy as np
import pytest

import emu_util
import oracle_util as ou
import stream_model as sm
from ribbit_b200 import synth


def _same(got, exp):
    for s in (1, 2, 3):
        assert got[s].shape == exp[s].shape, "stream %d: %s vs %s" % (s, got[s].shape, exp[s].shape)
        assert (got[s] == exp[s]).all(), "stream %d differs" % s


def test_emulator_matches_golden(golden):
    for name, g in golden.items():
        seq = g["seq"].tobytes()
        mlo, mhi = int(g["args"][0]), int(g["args"][1])
        ev = ou.scan_events(seq, mlo, mhi)
        exp = sm.expected_streams(seq, ev)
        for cw in (5, 33, 1 << 30):
            got, _ = emu_util.emu_streams(seq, mlo, mhi, cw)
            _same(got, exp)


@pytest.mark.parametrize("L,nd,mlo,mhi", [(1500, 0.0, 2, 100), (1500, 0.02, 1, 6), (4000, 0.002, 2, 100),
                                          (4000, 0.1, 5, 30), (12000, 0.001, 2, 100), (3000, 0.0, 90, 100)])
def test_emulator_matches_oracle_fuzz(L, nd, mlo, mhi):
    rng = np.random.default_rng(L + mlo)
    seq = synth.fuzz_contig(rng, L, nd)
    exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
    restarts = 0
    for cw in (7, 64, 1 << 30):
        got, rs = emu_util.emu_streams(seq, mlo, mhi, cw)
        restarts += rs
        _same(got, exp)


def test_emulator_long_repeat_forces_restarts():
    # a 6 kb perfect repeat: chunks that start inside it cannot synchronise with a short warm-up
    rng = np.random.default_rng(3)
    seq = synth.random_bases(rng, 2000).tobytes() + b"ACGTTGCA" * 750 + synth.random_bases(rng, 2000).tobytes()
    exp = sm.expected_streams(seq, ou.scan_events(seq, 2, 40))
    got, rs = emu_util.emu_streams(seq, 2, 40, 16)
    assert rs > 0
    _same(got, exp)


def _nrun_contig(rng, L):
    seq = bytearray(synth.fuzz_contig(rng, L, 0.001))
    for n in (40, 700, 2500):
        a = int(rng.integers(300, L - 3000))
        seq[a:a + n] = b"N" * n
        m = int(rng.integers(1, 30))
        rep = (bytes(synth.random_bases(rng, m)) * 300)[:int(rng.integers(20, 250))]
        seq[a - len(rep):a] = rep                      # a repeat running into the N run
        rep2 = (bytes(synth.random_bases(rng, m + 1)) * 300)[:int(rng.integers(20, 250))]
        seq[a + n:a + n + len(rep2)] = rep2            # and one starting right after it
    return bytes(seq)


def test_emulator_chunks_inside_long_n_runs_skip_instead_of_rescanning():
    rng = np.random.default_rng(9)
    for trial in range(3):
        seq = _nrun_contig(rng, 9000)
        for mlo, mhi in [(2, 100), (1, 6)]:
            exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
            skips = 0
            for cw in (3, 11, 40):
                got, _ = emu_util.emu_streams(seq, mlo, mhi, cw)
                skips += emu_util.emu_streams.last_skips
                _same(got, exp)
            assert skips > 0


def _regress_cases():
    import glob
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regress")
    for f in sorted(glob.glob(os.path.join(d, "*.txt"))):
        hdr, seq = open(f, "rb").read().split(b"\n", 1)
        mlo, mhi, cw = map(int, hdr.split())
        yield os.path.basename(f), seq, mlo, mhi, cw


def test_emulator_regression_cases():
    """Inputs a GPU fuzz campaign once got wrong (tests/golden/regress: first line = min_mlen max_mlen chunk_words): a
    fast -> slow transition inside an anchors-only warm-up word must not take the machine state from the fast view."""
    n = 0
    for name, seq, mlo, mhi, cw in _regress_cases():
        exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
        for c in {cw if cw else 1 << 30, 1, 2, 3}:
            got, _ = emu_util.emu_streams(seq, mlo, mhi, c)
            _same(got, exp)
        n += 1
    assert n >= 3


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_emulator_tiny_chunks_fuzz(seed):
    """Chunks of 1-3 words: every chunk start is a warm-up, transitions fall into warm-ups."""
    rng = np.random.default_rng(1000 + seed)
    for L, nd, mlo, mhi in ((2000, 0.01, 5, 30), (2000, 0.05, 2, 100), (1500, 0.01, 40, 100), (2500, 0.003, 2, 24)):
        seq = synth.fuzz_contig(rng, L, nd, m_range=(mlo, min(mhi, 60)))
        exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
        for cw in (1, 2, 3):
            got, _ = emu_util.emu_streams(seq, mlo, mhi, cw)
            _same(got, exp)


@pytest.mark.parametrize("mlo,mhi", [(700, 900), (990, 1000)])
def test_emulator_large_motif_sizes(mlo, mhi):
    """Motif sizes up to the ABI's limit (max_mlen 1000): multi-word shifts, the keep filter beyond its exact range, contigs
    only a few multiples of the shift long, N runs (the same cases run on the GPU in tests/test_gpu_parity.py)."""
    rng = np.random.default_rng(mlo)
    for case in range(4):
        L = int(rng.choice([mhi + 5, 2 * mhi + 77, 5 * mhi, 9000]))
        seq = bytearray(synth.fuzz_contig(rng, L, float(rng.choice([0, 0.001])), m_range=(2, 40)))
        m = int(rng.integers(mlo, mhi + 1))
        unit = bytes(synth.random_bases(rng, m))
        k = min(L, int(rng.integers(2 * m, 4 * m)))
        a = int(rng.integers(0, L - k + 1))
        seq[a:a + k] = (unit * 5)[:k]
        if case % 2:
            q = int(rng.integers(0, L - 20)); seq[q:q + 20] = b"N" * 20
        seq = bytes(seq)
        exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
        for cw in (1 << 30, 7):
            got, _ = emu_util.emu_streams(seq, mlo, mhi, cw)
            _same(got, exp)


def test_keep_filter_of_the_large_shift_loop():
    """scan_core.h keep_by_last (anchored cutoff from the position of the latest S bit) equals the bit-by-bit check for every
    cutoff >= 32, and smear_from_last hands a lane back to the smear network with a state that gives the same answers."""
    import ctypes
    L = emu_util.lib()
    L.emu_keep_filter_check.restype = ctypes.c_int
    L.emu_keep_filter_check.argtypes = [ctypes.c_uint64, ctypes.c_int]
    for seed in (1, 2, 3):
        assert L.emu_keep_filter_check(seed, 4000) == 0


def _kept_digest(rows, off, lo=None, hi=None):
    from ribbit_b200 import workloads as wl
    k = (rows[:, 3] & 3) == 0  # neither DROPPED nor PSEUDO
    return wl.digest_rows(rows[k, 0].astype(np.int64) + off, rows[k, 1].astype(np.int64) + off, rows[k, 2], lo, hi)


def test_emulator_and_oracle_equal_the_reference_digests_of_the_bench_workload():
    """The digests bench.py gates on (tests/golden/c3_digests.json, kept calls of the UNMODIFIED reference on the C3
    workload): the kernels' logic on the CPU emulator reproduces them for the whole 47 Mbp contig and for the gate windows
    of two more contigs (one over the end of the leading N run, one over the start of the 3 Mb N run); the oracle port
    reproduces one window. Pins emulator, oracle and the gate's own digest code to the reference at bench scale."""
    from ribbit_b200 import workloads as wl
    g = wl.load_digests()
    full = g["full"]["contig"]
    seq = wl.c3_contig(full)
    got, _ = emu_util.emu_streams(seq, 2, 100, chunk_words=8192)
    for s in (1, 2, 3):
        assert _kept_digest(got[s], 0) == g["full"]["kept"][str(s)], "contig %d stream %d" % (full, s)
    for i in (19, 21):
        seq = wl.c3_contig(i)
        lo, hi = wl.c3_window(i, len(seq))
        want = g["windows"][str(i)]
        assert (want["L"], want["lo"], want["hi"]) == (len(seq), lo, hi)
        got, _ = emu_util.emu_streams(seq[lo:hi], 2, 100, chunk_words=2048)
        for s in (1, 2, 3):
            assert _kept_digest(got[s], lo, lo, hi) == want["kept"][str(s)], "contig %d window stream %d" % (i, s)
        if i == 21:
            ev = ou.scan_events(seq[lo:hi], 2, 100)
            for s in (1, 2, 3):
                r = ev[ev[:, 0] == s]
                k = wl.kept_mask(s, r[:, 1], r[:, 2], r[:, 3])
                d = wl.digest_rows(r[k, 1].astype(np.int64) + lo, r[k, 2].astype(np.int64) + lo, r[k, 3], lo, hi)
                assert d == want["kept"][str(s)], "oracle, contig %d window stream %d" % (i, s)
