"""Multi-GPU sharding of the scan: contigs are independent units, so each rank (one process per GPU)
scans its own contigs and the per-contig streams are reassembled in input order. No collective is on the data path;
torch.distributed is used only to collect the results (gather_object) — or not at all when every rank writes its own
output.

One contig over several GPUs (scan_contig_split): every rank loads the contig, scans its own range of 32-base words
(rb_set_word_range) and the parts are concatenated in rank order; the only fix-up is the end carried by PSEUDO records
(stitch_parts).
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


def plan_parts(lengths: Sequence[int], world: int, tol: float = 0.02, align_words: int = 1024) -> List[List[tuple]]:
    """Strong-scaling partition by (contig, chunk): whole contigs by `assign`, then the most loaded rank
    hands the tail of its largest part to the least loaded one until the loads (in 32-base words) differ by at most
    tol x the mean. Returns, per rank, parts (contig, word_first, word_last) in (contig, word_first) order; a whole contig
    is (i, 0, ceil(L/32)). Deterministic. A rank that owns a range of a contig scans it with rb_set_word_range."""
    nw = [(int(L) + 31) // 32 for L in lengths]
    parts: List[List[tuple]] = [[(i, 0, nw[i]) for i in o] for o in assign(lengths, world)]
    if world > 1 and sum(nw) > 0:
        mean = sum(nw) / world
        for _ in range(8 * world):
            load = [sum(b - a for _, a, b in p) for p in parts]
            hi = max(range(world), key=lambda r: (load[r], -r))
            lo = min(range(world), key=lambda r: (load[r], r))
            diff = load[hi] - load[lo]
            if diff <= tol * mean:
                break
            take = (diff // 2) // align_words * align_words
            k = max(range(len(parts[hi])), key=lambda j: (parts[hi][j][2] - parts[hi][j][1], -j))
            c, a, b = parts[hi][k]
            if take <= 0 or b - a <= take:
                break
            parts[hi][k] = (c, a, b - take)
            parts[lo].append((c, b - take, b))
    return [sorted(p) for p in parts]


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


PSEUDO = 2  # rb_rec flag (include/ribbit_scan.h)


def split_words(n_words: int, parts: int) -> List[tuple]:
    """[first, last) word ranges of a contig of n_words words for `parts` ranks: equal shares, empty ranges dropped
    (a contig shorter than `parts` words keeps fewer parts)."""
    parts = max(1, min(parts, max(n_words, 1)))
    cuts = [n_words * k // parts for k in range(parts + 1)]
    return [(cuts[k], cuts[k + 1]) for k in range(parts) if cuts[k + 1] > cuts[k] or n_words == 0]


def stitch_parts(parts: Sequence[dict], elided: Sequence[Sequence[int]]) -> dict:
    """parts[r] = {1|2|3: rows (start, end, mlen, flags, time)} of word range r (in contig order), elided[r] = the largest
    end among the candidates elided in part r for the substitution and the anchored stream (-1: none). Returns the streams
    of the whole contig: the parts back to back, the end of every PSEUDO record raised to the largest elided end of the
    earlier parts (a PSEUDO record carries the largest end elided since the contig start, and a part only saw its own)."""
    import numpy as np
    out = {}
    for s in (1, 2, 3):
        rows = [np.array(p[s], copy=True) for p in parts]
        if s != 1:
            carry = -1
            for r, a in enumerate(rows):
                if r > 0 and carry >= 0 and len(a):
                    ps = (a[:, 3] & PSEUDO) != 0
                    a[ps, 1] = np.maximum(a[ps, 1], carry)
                carry = max(carry, int(elided[r][s - 2]))
        out[s] = np.concatenate(rows) if rows else np.zeros((0, 5), dtype=np.int64)
    return out


def scan_contig_split(contig: bytes, part_fn: Callable[[bytes, int, int], tuple], rank: int = 0, world: int = 1, gather: bool = True):
    """One contig over `world` ranks. part_fn(contig, word_first, word_last) -> (streams, [elided_S, elided_A]) scans one
    word range. With gather=True rank 0 returns the stitched streams of the whole contig (other ranks None)."""
    ranges = split_words((len(contig) + 31) // 32, world)
    local = part_fn(contig, *ranges[rank]) if rank < len(ranges) else None
    if not gather:
        return local
    if world == 1:
        gathered = [local]
    else:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(local, gathered, dst=0)
        if rank != 0:
            return None
    got = [g for g in gathered if g is not None]
    return stitch_parts([g[0] for g in got], [g[1] for g in got])


def gpu_part_fn(min_mlen=2, max_mlen=100, device=0):
    """part_fn backed by the CUDA library: the whole contig is loaded (warm-up and N-run handling read the words in
    front of the range), only the words of the range are scanned."""
    from . import scan
    sc = scan.Scanner(min_mlen, max_mlen, device=device)

    def fn(contig, word_first, word_last):
        sc.load([contig])
        sc.set_word_range(word_first, word_last)
        res = sc.scan()
        return scan.contig_streams(res, 0), sc.elided_max()

    return fn
