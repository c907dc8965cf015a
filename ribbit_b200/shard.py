"""Multi-GPU sharding of the scan: contigs are independent units, so each rank (one process per GPU)
scans its own contigs and the per-contig streams are reassembled in input order. No collective is on the data path;
torch.distributed is used only to collect the results (gather_object) — or not at all when every rank writes its own
output.
"""
from typing import Callable, Dict, List, Sequence


def assign(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-first assignment of contigs to ranks by bases; deterministic (ties: lower contig index first,
    lower rank first). Returns, per rank, the contig indices in input order."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    owned: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owned[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(o) for o in owned]


def scan_local(contigs: Sequence[bytes], mine: Sequence[int], scan_fn: Callable[[List[bytes]], List[dict]]) -> Dict[int, dict]:
    """Scans the contigs this rank owns as one batch. scan_fn(list of contig bytes) -> list of per-contig results."""
    res = scan_fn([contigs[i] for i in mine]) if mine else []
    return {i: r for i, r in zip(mine, res)}


def scan_sharded(contigs: Sequence[bytes], scan_fn: Callable[[List[bytes]], List[dict]], rank: int = 0, world: int = 1,
                 gather: bool = True):
    """Every rank scans its share; with gather=True rank 0 returns the per-contig results in input order (other ranks
    return None), otherwise each rank returns its own {contig index: result}."""
    owned = assign([len(c) for c in contigs], world)
    local = scan_local(contigs, owned[rank], scan_fn)
    if not gather:
        return local
    if world == 1:
        return [local[i] for i in range(len(contigs))]
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0)
    if rank != 0:
        return None
    merged: Dict[int, dict] = {}
    for p in parts:
        merged.update(p)
    return [merged[i] for i in range(len(contigs))]


def gpu_scan_fn(min_mlen=2, max_mlen=100, device=0):
    """scan_fn backed by the CUDA library (one context on `device`)."""
    from . import scan
    sc = scan.Scanner(min_mlen, max_mlen, device=device)

    def fn(batch):
        sc.load(batch)
        res = sc.scan()
        return [scan.contig_streams(res, i) for i in range(len(batch))]

    return fn
