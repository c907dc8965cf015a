"""Parity of the CUDA path (through the C ABI) with the oracle and the golden vectors. Bit-exact: integer work."""
import numpy as np
import pytest

import oracle_util as ou
import stream_model as sm
from ribbit_b200 import scan, synth

pytestmark = pytest.mark.gpu


def _same(got, exp, what=""):
    for s in (1, 2, 3):
        assert got[s].shape == exp[s].shape, "%s stream %d: %s vs %s" % (what, s, got[s].shape, exp[s].shape)
        assert (got[s] == exp[s]).all(), "%s stream %d differs" % (what, s)


def _scan_one(seq, mlo, mhi, chunk_words=0):
    sc = scan.Scanner(mlo, mhi, chunk_words=chunk_words)
    sc.load([seq])
    res = sc.scan()
    t = sc.timing()
    planes = sc.planes(0)
    sc.close()
    return scan.contig_streams(res, 0), t, planes


def test_golden_vectors(built, golden):
    for name, g in golden.items():
        seq = g["seq"].tobytes()
        mlo, mhi = int(g["args"][0]), int(g["args"][1])
        exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
        for cw in (0, 5, 33):
            got, _, planes = _scan_one(seq, mlo, mhi, cw)
            _same(got, exp, "%s cw=%d" % (name, cw))
        hi, lo, nn = ou.pack(seq)
        assert (planes[0] == hi).all() and (planes[1] == lo).all() and (planes[2] == nn).all()
        # the kept records are exactly the reference's calls that pass the consumer cutoff (golden cp1)
        if int(g["rc"][0]) == 0:
            cp1 = g["cp1"]
            for s in (1, 2, 3):
                ref = cp1[cp1[:, 0] == s][:, 1:4]
                cut = {1: lambda m: 0, 2: sm.cut_subst, 3: sm.cut_anch}[s]
                keep = np.array([e - st >= cut(m) for st, e, m in ref.tolist()], dtype=bool)
                mine = got[s][(got[s][:, 3] & (scan.REC_DROPPED | scan.REC_PSEUDO)) == 0][:, :3]
                assert (mine == ref[keep]).all(), name


@pytest.mark.parametrize("L,nd,mlo,mhi", [(20000, 0.001, 2, 100), (20000, 0.0, 1, 6), (50000, 0.0005, 2, 100),
                                          (30000, 0.02, 5, 30), (30000, 0.0, 60, 224)])
def test_fuzz_vs_oracle(built, L, nd, mlo, mhi):
    rng = np.random.default_rng(L + mlo)
    seq = synth.fuzz_contig(rng, L, nd, m_range=(mlo, min(mhi, 120)))
    exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
    for cw in (0, 16):
        got, _, _ = _scan_one(seq, mlo, mhi, cw)
        _same(got, exp, "cw=%d" % cw)


def test_long_repeat_restarts(built):
    rng = np.random.default_rng(3)
    seq = synth.random_bases(rng, 2000).tobytes() + b"ACGTTGCA" * 750 + synth.random_bases(rng, 2000).tobytes()
    exp = sm.expected_streams(seq, ou.scan_events(seq, 2, 40))
    got, t, _ = _scan_one(seq, 2, 40, 16)
    assert t["restarts"] > 0
    _same(got, exp)


def test_long_n_runs_with_small_chunks(built):
    from test_emu import _nrun_contig
    rng = np.random.default_rng(9)
    for trial in range(3):
        seq = _nrun_contig(rng, 20000)
        for mlo, mhi in [(2, 100), (1, 6)]:
            exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
            for cw in (3, 11, 0):
                got, _, _ = _scan_one(seq, mlo, mhi, cw)
                _same(got, exp, "trial %d cw=%d" % (trial, cw))


def test_many_contigs_batch(built):
    rng = np.random.default_rng(11)
    contigs = [synth.fuzz_contig(rng, int(L), nd) for L, nd in
               [(0, 0), (3, 0), (1000, 0), (1000, 0.01), (31, 0), (32, 0), (33, 0), (64, 0.1), (5000, 0.001), (7, 0), (8, 0)] * 3]
    for mlo, mhi in [(1, 6), (2, 100)]:
        sc = scan.Scanner(mlo, mhi)
        sc.load(contigs)
        res = sc.scan()
        for i, seq in enumerate(contigs):
            exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
            _same(scan.contig_streams(res, i), exp, "contig %d" % i)
        sc.close()


def test_record_buffer_overflow_is_retried(built):
    # dense short-motif repeats produce far more candidates per word than the reservation
    seq = (b"ACACACACACAGACACACACATACACACAC" * 400)
    exp = sm.expected_streams(seq, ou.scan_events(seq, 2, 100))
    sc = scan.Scanner(2, 100, chunk_words=64)
    sc.load([seq])
    got = scan.contig_streams(sc.scan(), 0)
    _same(got, exp)
    sc.close()


def test_mid_size_properties(built):
    # 4 Mbp: too slow for the bit-serial oracle in a unit test; check size-independent properties instead
    seq = synth.contig_c2(4_000_000, seed=5)
    sc = scan.Scanner(2, 100)
    sc.load([seq])
    res = sc.scan()
    t = sc.timing()
    sc2 = scan.Scanner(2, 100, chunk_words=257)  # a different chunking must give the identical streams
    sc2.load([seq])
    res2 = sc2.scan()
    for s in range(3):
        a, _ = res[s]
        b, _ = res2[s]
        assert len(a) == len(b) and (a == b).all()
        real = a[(a["flags"] & scan.REC_PSEUDO) == 0]
        assert (np.diff(real["time"].astype(np.int64)) >= 0).all()          # emission order is time-major
        assert (real["end"] > real["start"]).all()
        same_t = np.diff(real["time"].astype(np.int64)) == 0
        assert (np.diff(real["mlen"].astype(np.int64))[same_t] >= 0).all()  # motifs ascending inside one position
    # a prefix of the contig scanned alone agrees with the oracle
    pre = seq[:60000]
    sc3 = scan.Scanner(2, 100)
    sc3.load([pre])
    _same(scan.contig_streams(sc3.scan(), 0), sm.expected_streams(pre, ou.scan_events(pre, 2, 100)))
    for x in (sc, sc2, sc3):
        x.close()
    assert t["launches"] >= 4


def test_seed_filter_matches_reference_definition(built):
    """K5 (rb_filter_seeds) against the definition in parse_seed.cpp:344-367 evaluated with the oracle's anchored plane."""
    rng = np.random.default_rng(12)
    seq = synth.fuzz_contig(rng, 30000, 0.002, m_range=(2, 60))
    L = len(seq)
    a = np.frombuffer(seq, dtype=np.uint8)
    isn = ~np.isin(a, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    sc = scan.Scanner(2, 100)
    sc.load([seq])
    res = sc.scan()
    # seeds: kept candidates of the three streams (clamped as the merges do: end <= L - m)
    seeds = []
    for s in range(3):
        r = res[s][0]
        r = r[(r["flags"] & (scan.REC_DROPPED | scan.REC_PSEUDO)) == 0][::7][:150]
        for st, en, m in zip(r["start"].tolist(), r["end"].tolist(), r["mlen"].tolist()):
            en = min(en, L - m)
            if en > st:
                seeds.append((0, st, en, m))
    seeds = np.array(seeds, dtype=np.int32)
    got = sc.filter_seeds(seeds)
    for (c, st, en, m), (seq_len, longest) in zip(seeds.tolist(), got.tolist()):
        nn = np.flatnonzero(isn[st:en + m])
        exp_len = int(nn[0]) if len(nn) else (en - st) + m
        bm = ou.anchored_plane(seq, 2, 100, m, st, en)
        runs = np.diff(np.flatnonzero(np.concatenate(([0], bm, [0])) == 0)) - 1
        assert seq_len == exp_len and longest == int(runs.max(initial=0)), (st, en, m)
    sc.close()


def test_anchor_planes_match_oracle(built):
    rng = np.random.default_rng(13)
    seq = synth.fuzz_contig(rng, 5000, 0.002, m_range=(2, 40)) + b"ACG" * 40
    sc = scan.Scanner(2, 30)
    sc.load([seq])
    sc.scan()
    planes = sc.anchor_planes(0, 1, 32)
    L = len(seq)
    pos = np.arange(L)
    for m in (2, 3, 7, 16, 30):
        bm = ou.anchored_plane(seq, 2, 30, m, 0, L)
        hi, lo, _ = sc.planes(0)
        code = (((hi[pos >> 5] >> (pos & 31).astype(np.uint32)) & 1) * 2 + ((lo[pos >> 5] >> (pos & 31).astype(np.uint32)) & 1)).astype(np.int64)
        x = np.where(pos + m < L, code == np.concatenate([code[m:], np.zeros(m, np.int64)])[:L], code == 0)
        got = x.copy()
        for s in range(max(m - 2, 1) if m > 2 else 1, m + 3):
            if s != m:
                got |= ((planes[s - 1][pos >> 5] >> (pos & 31).astype(np.uint32)) & 1).astype(bool)
        assert (got.astype(np.uint8) == bm).all(), m
    sc.close()


def test_sharded_scan_single_rank(built):
    from ribbit_b200 import shard
    rng = np.random.default_rng(4)
    contigs = [synth.fuzz_contig(rng, int(L), 0.005) for L in (900, 0, 2500, 1200, 40)]
    out = shard.scan_sharded(contigs, shard.gpu_scan_fn(2, 30), 0, 1)
    for seq, got in zip(contigs, out):
        _same(got, sm.expected_streams(seq, ou.scan_events(seq, 2, 30)))


def test_pipeline_two_contexts_gives_same_streams(built):
    from ribbit_b200 import pipeline
    rng = np.random.default_rng(21)
    seqs = [synth.fuzz_contig(rng, 40000, 0.001) for _ in range(5)]
    pipe = pipeline.ScanPipeline(2, 100, depth=2, copy=True)
    bufs = [np.frombuffer(s + b"\0", dtype=np.uint8) for s in seqs]
    futs = [pipe.submit_flat(b, [len(s)]) for b, s in zip(bufs, seqs)]
    for s, f in zip(seqs, futs):
        _same(scan.contig_streams(f.result(), 0), sm.expected_streams(s, ou.scan_events(s, 2, 100)))
    pipe.close()


def test_pipeline_grouped_batches_with_offsets(built):
    """The end-to-end path of bench.py: contigs at offsets of one host buffer, grouped into batches (group_contigs),
    three contexts, compact fetch; every contig equals the oracle."""
    from ribbit_b200 import pipeline
    rng = np.random.default_rng(22)
    seqs = [synth.fuzz_contig(rng, int(L), 0.002) for L in (30000, 12000, 0, 9000, 25000, 700, 18000)]
    lengths = [len(s) for s in seqs]
    offs, total = [], 0
    for s in seqs:
        offs.append(total); total += len(s) + 1
    host = np.zeros(total + 64, dtype=np.uint8)
    for o, s in zip(offs, seqs):
        host[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
    groups = pipeline.group_contigs(range(len(seqs)), lengths, 40000)
    assert [c for g in groups for c in g] == list(range(len(seqs))) and len(groups) >= 3
    pipe = pipeline.ScanPipeline(2, 100, depth=3, copy=True, compact=True, trace=True)
    futs = [pipe.submit_flat(host[offs[g[0]]:offs[g[-1]] + lengths[g[-1]] + 1], [lengths[c] for c in g],
                             offsets=[offs[c] - offs[g[0]] for c in g]) for g in groups]
    for g, f in zip(groups, futs):
        comp = f.result()
        for k, c in enumerate(g):
            exp = sm.expected_streams(seqs[c], ou.scan_events(seqs[c], 2, 100))
            for st in range(3):
                r8, off8, lg = comp[st]
                rows = scan.expand_compact(r8, lg)[off8[k]:off8[k + 1]]
                e = exp[st + 1][:, :4]  # (start, end, mlen, flags); the compact records carry no emission time
                assert rows.shape == e.shape and (rows == e).all(), "contig %d stream %d" % (c, st)
    assert len(pipe.trace) == len(groups)
    pipe.close()


def test_c5_shape_many_short_contigs(built):
    """BASELINE.json configs[4] shape (1 kb contigs, -m 1 -M 6), scaled to 20 000 contigs in one batch: 2 500 of the
    contigs against the oracle (the reference itself segfaults on this shape after logging CP1)."""
    contigs = synth.contigs_c5(n=20000, length=1000, seed=5)
    sc = scan.Scanner(1, 6)
    sc.load(contigs)
    res = sc.scan()
    t = sc.timing()
    rng = np.random.default_rng(0)
    for i in rng.choice(len(contigs), 2500, replace=False).tolist() + [0, 1, 2, 3, len(contigs) - 1]:
        _same(scan.contig_streams(res, i), sm.expected_streams(contigs[i], ou.scan_events(contigs[i], 1, 6)), "contig %d" % i)
    for s in range(3):
        off = res[s][1]
        assert off[0] == 0 and off[-1] == len(res[s][0]) and (np.diff(off) >= 0).all()
    sc.close()
    assert t["launches"] >= 4


def test_c3_shape_multi_contig_batch(built):
    """Several contigs of different sizes in one batch (configs[2] shape, scaled): two of them against the oracle, all of
    them identical to scanning each alone with another chunking."""
    lengths = [600_000, 50_000, 1_200_000, 333_333]
    contigs = [synth.contig_c2(L, seed=100 + i) for i, L in enumerate(lengths)]
    sc = scan.Scanner(2, 100)
    sc.load(contigs)
    res = sc.scan()
    for i in (1, 3):
        _same(scan.contig_streams(res, i), sm.expected_streams(contigs[i], ou.scan_events(contigs[i], 2, 100)), "contig %d" % i)
    for i, seq in enumerate(contigs):
        one = scan.Scanner(2, 100, chunk_words=211)
        one.load([seq])
        r1 = one.scan()
        for s in range(3):
            a = res[s][0][res[s][1][i]:res[s][1][i + 1]]
            assert len(a) == len(r1[s][0]) and (a == r1[s][0]).all(), (i, s)
        one.close()
    sc.close()


FUZZ_RANGES = [(2, 100), (1, 6), (2, 24), (5, 30), (3, 10), (40, 100), (1, 100), (2, 8), (90, 100), (2, 150)]


def fuzz_case(rng):
    """One case of the fuzz campaigns (tools/fuzz_gpu.py, tools/fuzz_emu.py): contig lengths 1 .. 20 000, N densities up to
    30 %, injected N runs and long repeats, ten motif ranges, chunk sizes 1 .. 64 words and automatic."""
    L = int(rng.choice([1, 7, 8, 31, 32, 33, 63, 64, 65, 100, 500, 2000, 7000, 20000]))
    nd = float(rng.choice([0, 0, 0.001, 0.01, 0.05, 0.3]))
    mlo, mhi = FUZZ_RANGES[int(rng.integers(len(FUZZ_RANGES)))]
    seq = synth.fuzz_contig(rng, L, nd, m_range=(mlo, min(mhi, 60)))
    if rng.random() < 0.3 and L > 200:
        b = bytearray(seq); a = int(rng.integers(0, L - 100)); k = int(rng.integers(50, min(3000, L - a)))
        if rng.random() < 0.5:
            b[a:a + k] = b"N" * k
        else:
            m = int(rng.integers(1, 40)); b[a:a + k] = (bytes(synth.random_bases(rng, m)) * (k // m + 1))[:k]
        seq = bytes(b)
    cw = int(rng.choice([0, 0, 1, 2, 3, 5, 17, 64]))
    return seq, mlo, mhi, cw


@pytest.mark.parametrize("seed", [101, 102, 103, 104])
def test_fuzz_slice(built, seed):
    """A deterministic slice (4 x 500 cases) of the campaign that found the round-1 state-conversion bug; the open-ended
    campaigns stay in tools/."""
    rng = np.random.default_rng(seed)
    for n in range(500):
        seq, mlo, mhi, cw = fuzz_case(rng)
        got, _, _ = _scan_one(seq, mlo, mhi, cw)
        _same(got, sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi)), "seed %d case %d (m %d-%d, cw %d, L %d)" % (seed, n, mlo, mhi, cw, len(seq)))


@pytest.mark.parametrize("mlo,mhi", [(700, 900), (990, 1000), (300, 420)])
def test_large_motif_sizes(built, mlo, mhi):
    """Motif sizes up to the ABI's limit (rb_create: max_mlen <= 1000): multi-word shifts, guard = s_hi/32 + 5, keep filter
    beyond its exact range (cutoffs above 96), contigs only a few multiples of the shift long, with N runs."""
    rng = np.random.default_rng(mlo)
    for case in range(6):
        L = int(rng.choice([mhi + 5, 2 * mhi + 77, 5 * mhi, 9000]))
        seq = bytearray(synth.fuzz_contig(rng, L, float(rng.choice([0, 0.001])), m_range=(2, 40)))
        m = int(rng.integers(mlo, mhi + 1))          # a repeat of a motif size inside the range
        unit = bytes(synth.random_bases(rng, m))
        k = min(L, int(rng.integers(2 * m, 4 * m)))
        a = int(rng.integers(0, L - k + 1))
        seq[a:a + k] = (unit * 5)[:k]
        if case % 2:
            q = int(rng.integers(0, L - 20)); seq[q:q + 20] = b"N" * 20
        seq = bytes(seq)
        exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
        for cw in (0, 7):
            got, _, _ = _scan_one(seq, mlo, mhi, cw)
            _same(got, exp, "m %d-%d case %d cw %d" % (mlo, mhi, case, cw))


def test_full_size_c2_properties(built):
    """The bench workload itself (46.7 Mbp): stream invariants, and agreement of a window in the middle of the contig with
    the oracle run on that window plus context (candidates fully inside the window, away from its ends)."""
    L = 46_700_000
    seq = synth.contig_c2(L, seed=21)
    sc = scan.Scanner(2, 100)
    sc.load([seq])
    res = sc.scan()
    a0, a1 = 30_000_000, 30_060_000
    ctx = 4000
    sub = seq[a0 - ctx:a1 + ctx]
    exp = sm.expected_streams(sub, ou.scan_events(sub, 2, 100))
    for s in range(3):
        a = res[s][0]
        real = a[(a["flags"] & scan.REC_PSEUDO) == 0]
        assert (np.diff(real["time"].astype(np.int64)) >= 0).all()
        mine = real[(real["start"] >= a0) & (real["end"] < a1) & (real["flags"] == 0)]
        rows = np.stack([mine["start"] - (a0 - ctx), mine["end"] - (a0 - ctx), mine["mlen"]], axis=1).astype(np.int64)
        e = exp[s + 1]
        e = e[(e[:, 3] == 0) & (e[:, 0] >= ctx) & (e[:, 1] < ctx + (a1 - a0))][:, :3]
        assert len(rows) == len(e) and (rows == e).all(), s
    sc.close()


def test_abi_error_behaviour(built):
    import ctypes
    lib = scan.load_library()
    p = scan.RbParams(2, 100, 0, 0)
    ctx = lib.rb_create(0, ctypes.byref(p))
    assert ctx
    out = scan.RbStreams()
    assert lib.rb_scan_device(ctx) == -4 and b"no contigs loaded" in lib.rb_last_error(ctx)      # RB_E_STATE
    assert lib.rb_fetch(ctx, ctypes.byref(out)) == -4
    lengths = np.array([-5], dtype=np.int32); offs = np.zeros(1, dtype=np.int64); buf = np.zeros(8, dtype=np.uint8)
    assert lib.rb_load_contigs(ctx, buf.ctypes.data, offs.ctypes.data, lengths.ctypes.data, 1) == -5  # RB_E_RANGE
    assert lib.rb_load_contigs(ctx, None, None, None, 3) == -1                                         # RB_E_ARG
    assert lib.rb_load_contigs(ctx, None, None, None, 0) == 0                                          # empty batch is fine
    assert lib.rb_scan(ctx, ctypes.byref(out)) == 0 and out.n_contigs == 0 and list(out.n) == [0, 0, 0]
    lib.rb_destroy(ctx)
    assert not lib.rb_create(99, ctypes.byref(p)) and b"not available" in lib.rb_last_error(None)


def test_python_cli_writes_kept_candidates(built, tmp_path):
    from ribbit_b200 import __main__ as cli
    rng = np.random.default_rng(31)
    seqs = [synth.fuzz_contig(rng, 3000, 0.002), synth.fuzz_contig(rng, 1200, 0.0)]
    fa = tmp_path / "x.fa"
    synth.write_fasta(str(fa), seqs, names=["one", "two"])
    outp = tmp_path / "c.tsv"
    assert cli.main(["scan", "-i", str(fa), "-o", str(outp), "-m", "2", "-M", "40"]) == 0
    rows = [l.split("\t") for l in open(outp).read().splitlines()]
    for name, seq in zip(["one", "two"], seqs):
        exp = sm.expected_streams(seq, ou.scan_events(seq, 2, 40))
        for s, tag in ((1, "P"), (2, "S"), (3, "A")):
            mine = [(int(r[2]), int(r[3]), int(r[4])) for r in rows if r[0] == name and r[1] == tag]
            e = exp[s]
            e = e[(e[:, 3] & (scan.REC_DROPPED | scan.REC_PSEUDO)) == 0]
            assert mine == [tuple(x) for x in e[:, :3].tolist()]


def test_compact_fetch_equals_full_fetch(built):
    rng = np.random.default_rng(41)
    long_rep = b"ACGGTCA" * 10000                                   # a 70 kb perfect repeat: candidates longer than 65535
    contigs = [synth.fuzz_contig(rng, 30000, 0.002), synth.random_bases(rng, 3000).tobytes() + long_rep + synth.random_bases(rng, 3000).tobytes(),
               b"", synth.fuzz_contig(rng, 900, 0.05)]
    sc = scan.Scanner(2, 100)
    sc.load(contigs)
    full = sc.scan()
    comp = sc.fetch_compact()
    nlong = 0
    for s in range(3):
        a, off = full[s]
        r8, off8, lg = comp[s]
        assert (off == off8).all() and len(a) == len(r8)
        rows = scan.expand_compact(r8, lg)
        exp = np.stack([a["start"], a["end"], a["mlen"], a["flags"]], axis=1).astype(np.int64)
        assert (rows == exp).all(), s
        nlong += len(lg)
    assert nlong > 0
    sc.close()


@pytest.mark.gpu
def test_contig_split_over_word_ranges():
    """SURVEY.md §8e, unit = (contig, chunk): one contig scanned as 2..9 word ranges (rb_set_word_range), stitched by
    shard.stitch_parts, equals the unsplit scan — boundaries fall inside N runs and repeats."""
    from ribbit_b200 import shard
    rng = np.random.default_rng(31)
    seq = bytearray(synth.contig_c2(300_000, seed=8, density_per_mbp=1500))
    seq[100_000:100_900] = b"N" * 900
    seq[150_016:150_016 + 37 * 60] = bytes(rng.choice(list(b"ACGT"), 37).astype(np.uint8)) * 60   # a repeat across the 2-way cut
    seq = bytes(seq)
    for lo, hi in ((2, 100), (1, 6)):
        sc = scan.Scanner(lo, hi)
        sc.load([seq])
        whole = scan.contig_streams(sc.scan(), 0)
        part_fn = shard.gpu_part_fn(lo, hi)
        nw = (len(seq) + 31) // 32
        for world in (2, 3, 9):
            parts = [part_fn(seq, a, b) for a, b in shard.split_words(nw, world)]
            got = shard.stitch_parts([p[0] for p in parts], [p[1] for p in parts])
            for s in (1, 2, 3):
                assert np.array_equal(got[s], whole[s]), (lo, hi, world, s)
        assert sum(len(p[0][3]) for p in parts) == len(whole[3])
        # a range set and then cleared gives the whole contig again; argument checks
        sc.load([seq]); sc.set_word_range(10, 20); sc.set_word_range(0, -1)
        again = scan.contig_streams(sc.scan(), 0)
        for s in (1, 2, 3):
            assert np.array_equal(again[s], whole[s])
        for bad in ((5, 5), (-1, 4), (3, nw + 1), (9, 2)):
            with pytest.raises(scan.RibbitScanError):
                sc.set_word_range(*bad)
        sc.load([seq, seq[:100]])
        with pytest.raises(scan.RibbitScanError):
            sc.set_word_range(0, 1)


@pytest.mark.gpu
def test_regression_cases_tiny_chunks():
    """tests/golden/regress (inputs a fuzz campaign once got wrong) and tiny chunk sizes through the C ABI."""
    import glob
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regress")
    files = sorted(glob.glob(os.path.join(d, "*.txt")))
    assert len(files) >= 3
    for f in files:
        hdr, seq = open(f, "rb").read().split(b"\n", 1)
        mlo, mhi, cw = map(int, hdr.split())
        exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
        for c in sorted({cw, 1, 2, 3}):
            sc = scan.Scanner(mlo, mhi, chunk_words=c)
            sc.load([seq])
            got = scan.contig_streams(sc.scan(), 0)
            sc.close()
            for s in (1, 2, 3):
                assert got[s].shape == exp[s].shape and (got[s] == exp[s]).all(), (os.path.basename(f), c, s)
