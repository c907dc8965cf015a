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
