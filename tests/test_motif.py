"""K7 — the row search of mostFrequentLongerMotif (parse_seed.cpp:153-256).

CPU part: the plain-C oracle (oracle/motif_oracle.c) against checkpoint CP4 of the instrumented reference
(tests/golden/golden_motif.npz: every call the reference made on five seeded contigs, with the row it chose), and the
kernel's bit-parallel row scoring (ribbit_b200/csrc/motif_core.h, run on the CPU by tests/emu) against the oracle on fuzzed
seeds. GPU part: rb_motif_rows through the C ABI against both. Bit-exact (integer work)."""
import os

import numpy as np
import pytest

import emu_util as eu
import oracle_util as ou
from ribbit_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_motif.npz")


def golden_cases():
    g = np.load(GOLD)
    for name in sorted({k.rsplit("_", 1)[0] for k in g.files}):
        yield name, g[name + "_seq"].tobytes(), g[name + "_cp4"]


def fuzz_seeds(rng, seq, n):
    """Random (start, end, mlen) on N-free stretches of seq: short and long seeds, motif sizes 3..140 (up to five 32-base
    chunks per unit; below 3 the reference's unit walk need not advance), seeds shorter than the motif, seeds at position 0 / at the contig end / right behind an N."""
    L = len(seq)
    isn = np.frombuffer(seq, np.uint8)
    isn = ~np.isin(isn, np.frombuffer(b"ACGTacgt", np.uint8))
    nxt = np.full(L + 1, L, np.int64)          # next N at or after p
    for p in range(L - 1, -1, -1):
        nxt[p] = p if isn[p] else nxt[p + 1]
    out = []
    while len(out) < n:
        m = int(rng.choice([3, 4, 11, 12, 31, 32, 33, 37, 64, 65, 100, 129, 140])) if rng.random() < 0.5 else int(rng.integers(3, 141))
        kind = rng.random()
        if kind < 0.1:
            start = 0
        elif kind < 0.2 and isn.any():
            start = int(rng.choice(np.flatnonzero(isn))) + 1
        else:
            start = int(rng.integers(0, L))
        if start >= L or isn[start]:
            continue
        room = int(nxt[start] - start)
        ln = int(min(room, rng.integers(1, 6 * m + 40))) if rng.random() < 0.8 else room
        if kind > 0.9:                          # ends at the contig end / at the N
            ln = room
        ln = min(ln, 1500)
        out.append((start, start + ln, m))
    return np.array(out, np.int32)


def fuzz_contigs():
    rng = np.random.default_rng(77)
    yield synth.fuzz_contig(rng, 3000, 0.0)
    yield synth.fuzz_contig(rng, 3000, 0.01)
    s = bytearray(synth.contig_c4(4000, seed=5, n_repeats=12))
    s[700] = ord("N"); s[2500:2510] = b"N" * 10
    yield bytes(s)
    yield b"ACGTTGCA" * 40 + b"N" + b"A" * 300 + b"CAGCAGCAT" * 50


def oracle_rows(seq, seeds):
    return np.array([ou.motif_row(seq, int(s), int(e - s), int(m)) for s, e, m in seeds], np.int32).reshape(-1, 2)


def test_oracle_matches_reference_cp4():
    n = 0
    for name, seq, cp4 in golden_cases():
        got = np.array([ou.motif_row(seq, int(s), int(l), int(m))[0] for s, l, m, _ in cp4], np.int32)
        assert np.array_equal(got, cp4[:, 3]), name
        n += len(cp4)
    assert n > 4000


def test_lane_logic_matches_reference_cp4():
    for name, seq, cp4 in golden_cases():
        seeds = np.stack([cp4[:, 0], cp4[:, 0] + cp4[:, 1], cp4[:, 2]], axis=1)
        assert np.array_equal(eu.emu_motif_rows(seq, seeds)[:, 0], cp4[:, 3]), name


def test_lane_logic_matches_oracle_fuzz():
    rng = np.random.default_rng(123)
    for seq in fuzz_contigs():
        seeds = fuzz_seeds(rng, seq, 400)
        assert np.array_equal(eu.emu_motif_rows(seq, seeds), oracle_rows(seq, seeds))


@pytest.mark.gpu
def test_gpu_motif_rows_golden_and_fuzz():
    from ribbit_b200 import scan
    rng = np.random.default_rng(321)
    sc = scan.Scanner(2, 100)
    for name, seq, cp4 in golden_cases():
        sc.load([seq]); sc.scan_device()
        seeds = np.stack([np.zeros(len(cp4), np.int32), cp4[:, 0], cp4[:, 0] + cp4[:, 1], cp4[:, 2]], axis=1)
        assert np.array_equal(sc.motif_rows(seeds)[:, 0], cp4[:, 3]), name
    contigs = list(fuzz_contigs())
    sc.load(contigs); sc.scan_device()                 # one batch of several contigs
    seeds, want = [], []
    for ci, seq in enumerate(contigs):
        sd = fuzz_seeds(rng, seq, 1500)
        seeds.append(np.concatenate([np.full((len(sd), 1), ci, np.int32), sd], axis=1))
        want.append(oracle_rows(seq, sd))
    assert np.array_equal(sc.motif_rows(np.concatenate(seeds)), np.concatenate(want))
    assert sc.motif_rows(np.zeros((0, 4), np.int32)).shape == (0, 2)


@pytest.mark.gpu
def test_gpu_motif_rows_long_seed_and_errors():
    """A seed of many slabs (rows spread over several warps, combined by atomicMax) and the ABI's argument checks."""
    from ribbit_b200 import scan
    rng = np.random.default_rng(9)
    unit = bytes(rng.choice(list(b"ACGT"), 23).astype(np.uint8))
    tract = bytearray(unit * 120)
    for p in rng.integers(0, len(tract), 150):
        tract[p] = ord("ACGT"[int(rng.integers(0, 4))])
    seq = synth.fuzz_contig(rng, 500, 0.0) + bytes(tract) + synth.fuzz_contig(rng, 500, 0.0)
    sc = scan.Scanner(2, 100)
    sc.load([seq]); sc.scan_device()
    seeds = np.array([[0, 480, 480 + 2800, 23], [0, 0, len(seq), 46], [0, 500, 3260, 12]], np.int32)
    want = oracle_rows(seq, seeds[:, 1:])
    assert np.array_equal(sc.motif_rows(seeds), want)
    for bad in ([1, 0, 10, 12], [0, -1, 10, 12], [0, 0, len(seq) + 1, 12], [0, 10, 5, 12], [0, 0, 10, 2]):
        with pytest.raises(scan.RibbitScanError):
            sc.motif_rows(np.array([bad], np.int32))
