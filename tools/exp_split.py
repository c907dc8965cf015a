"""One contig over several GPUs (rb_set_word_range). Single process: the parts run one after the other on one GPU and the
device time of every part is printed (what a rank would spend). Under torchrun (one rank per GPU): every rank scans its
part, time = max over ranks of load-resident scan time; rank 0 checks the stitched candidate counts against an unsplit scan.
   python tools/exp_split.py [Mbp] [parts]      |      torchrun --nproc-per-node N tools/exp_split.py [Mbp]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ribbit_b200 import scan, shard, synth  # noqa: E402

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 248.0
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
seq = synth.contig_c2(int(mbp * 1e6), seed=100, density_per_mbp=300)
nw = (len(seq) + 31) // 32
sc = scan.Scanner(2, 100, device=local)
sc.load([seq])
if world == 1:
    sc.scan_device(); sc.scan_device()
    t = sc.timing(); whole = sc.counts()
    print("whole contig %.0f Mbp: %.3f ms (pack %.3f scan %.3f merge %.3f) counts %s" % (mbp, t["total_ms"], t["pack_ms"], t["scan_ms"], t["merge_ms"], whole), flush=True)
    for parts in ([int(sys.argv[2])] if len(sys.argv) > 2 else [2, 4, 8]):
        tot = [0, 0, 0]; times = []
        for a, b in shard.split_words(nw, parts):
            sc.set_word_range(a, b)
            sc.scan_device(); sc.scan_device()
            t = sc.timing(); c = sc.counts()
            times.append(t["total_ms"])
            tot = [x + y for x, y in zip(tot, c)]
        # PSEUDO records are per part: counts of P agree exactly, S/A differ only by pseudo records of slow buckets (none added or lost)
        print("%d parts: per-part device ms %s -> slowest part %.3f ms; summed counts %s %s" % (
            parts, ["%.2f" % x for x in times], max(times), tot, "== whole" if tot == whole else "!= whole"), flush=True)
else:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ranges = shard.split_words(nw, world)
    sc.set_word_range(*ranges[rank])
    for _ in range(2):
        sc.scan_device()
    dist.barrier(); torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        sc.scan_device()
        ms.append(sc.timing()["total_ms"])
    tm = torch.tensor([float(np.median(ms))], device="cuda", dtype=torch.float64)
    cnt = torch.tensor(sc.counts(), device="cuda", dtype=torch.int64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        sc.set_word_range(0, -1)
        sc.scan_device(); sc.scan_device()
        t = sc.timing()
        print("contig %.0f Mbp over %d GPUs: %.3f ms per scan (max over ranks) vs %.3f ms on one GPU -> %.2fx; summed counts %s, unsplit %s" % (
            mbp, world, float(tm.item()), t["total_ms"], t["total_ms"] / float(tm.item()), cnt.tolist(), sc.counts()), flush=True)
    dist.barrier()
    dist.destroy_process_group()
