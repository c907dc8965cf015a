"""Host-side multi-GPU logic (ribbit_b200/shard.py) on CPU: world_size-2 gloo processes; the per-rank scan is stood in
by the oracle, so what is tested is the partition, the gather and the reassembly order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_util as ou
import stream_model as sm
from ribbit_b200 import shard, synth


def test_assign_is_balanced_and_deterministic():
    lengths = [248, 242, 198, 190, 181, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
    for world in (1, 2, 4, 8):
        owned = shard.assign(lengths, world)
        assert sorted(i for o in owned for i in o) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in o) for o in owned]
        assert max(loads) - min(loads) <= max(lengths)
        assert owned == shard.assign(lengths, world)
    assert shard.assign([], 2) == [[], []]
    assert shard.assign([5], 2) == [[0], []]


def _oracle_scan_fn(batch):
    return [sm.expected_streams(s, ou.scan_events(s, 2, 30)) for s in batch]


def _contigs():
    rng = np.random.default_rng(4)
    return [synth.fuzz_contig(rng, int(L), 0.005) for L in (900, 0, 2500, 1200, 40, 1800, 700)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = shard.scan_sharded(_contigs(), _oracle_scan_fn, rank, world)
    if rank == 0:
        q.put([{k: v.tolist() for k, v in r.items()} for r in out])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo_equal_single_process():
    ou.port()  # build the oracle before forking
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = shard.scan_sharded(_contigs(), _oracle_scan_fn, 0, 1)
    assert len(got) == len(single)
    for a, b in zip(got, single):
        for s in (1, 2, 3):
            assert np.array_equal(np.array(a[s], dtype=np.int64).reshape(-1, 5), b[s])


def _model_part_fn(contig, first, last):
    """A word range of the oracle's event list, encoded like the library encodes a part (pseudo ends from the part only)."""
    ev = ou.scan_events(contig, 2, 30)
    nw = (len(contig) + 31) // 32
    word = np.where(ev[:, 4] == -1, nw, ev[:, 4] >> 5)          # the tail flush belongs to the part that ends at nw
    keep = (word >= first) & ((word < last) | ((word == nw) & (last == nw)))
    streams, elided = sm.expected_streams(contig, ev[keep], with_elided=True)
    return streams, elided


def test_split_words_partitions():
    for nw, parts in ((0, 3), (1, 4), (7, 3), (100, 8), (5, 5), (1000, 1)):
        r = shard.split_words(nw, parts)
        assert r[0][0] == 0 and r[-1][1] == nw and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(b > a for a, b in r) or nw == 0


def test_contig_split_stitches_to_the_whole():
    """One contig cut into word ranges: concatenation + PSEUDO carry == the unsplit streams."""
    rng = np.random.default_rng(11)
    for L, nd in ((3000, 0.004), (2500, 0.0), (700, 0.02), (40, 0.0)):
        seq = synth.fuzz_contig(rng, L, nd)
        whole = sm.expected_streams(seq, ou.scan_events(seq, 2, 30))
        for world in (1, 2, 3, 7):
            parts = [_model_part_fn(seq, a, b) for a, b in shard.split_words((L + 31) // 32, world)]
            got = shard.stitch_parts([p[0] for p in parts], [p[1] for p in parts])
            for s in (1, 2, 3):
                assert np.array_equal(got[s], whole[s]), (L, world, s)
        assert any((whole[s][:, 3] & sm.PSEUDO).any() for s in (2, 3)) or L < 100


def _split_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = synth.fuzz_contig(np.random.default_rng(12), 2800, 0.004)
    out = shard.scan_contig_split(seq, _model_part_fn, rank, world)
    if rank == 0:
        q.put({k: v.tolist() for k, v in out.items()})
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_contig_split_two_ranks_gloo():
    ou.port()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    seq = synth.fuzz_contig(np.random.default_rng(12), 2800, 0.004)
    whole = sm.expected_streams(seq, ou.scan_events(seq, 2, 30))
    for s in (1, 2, 3):
        assert np.array_equal(np.array(got[s], dtype=np.int64).reshape(-1, 5), whole[s])


def test_group_contigs_batches_adjacent_contigs():
    from ribbit_b200 import pipeline
    lengths = [50, 40, 30, 100, 5, 5, 5, 70]
    g = pipeline.group_contigs(range(8), lengths, 80)
    assert g == [[0], [1, 2], [3], [4, 5, 6], [7]]
    assert sum(len(x) for x in g) == 8
    # contigs that are not neighbours in the host buffer never share a batch
    assert pipeline.group_contigs([0, 2, 3, 6], lengths, 1000) == [[0], [2, 3], [6]]
    assert pipeline.group_contigs([], lengths, 10) == []
