"""Timeline of the end-to-end pipeline on C3 (diagnostic): per contig and context, when the load, the kernels and the
fetch started and ended.   python tools/exp_e2e.py [depth] [scale]"""
import sys, os, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concurrent.futures import ProcessPoolExecutor
import numpy as np
from ribbit_b200 import workloads

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0


def gen(i):
    return workloads.c3_contig(i, scale)


if __name__ == "__main__":
    lengths = workloads.c3_lengths(scale)
    with ProcessPoolExecutor(min(24, os.cpu_count() or 8)) as ex:
        seqs = list(ex.map(gen, range(len(lengths))))
    import torch
    from ribbit_b200 import pipeline, scan
    total = sum(lengths) + len(lengths)
    host = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
    hn = host.numpy()
    offs = []
    o = 0
    for s in seqs:
        offs.append(o); hn[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8); o += len(s) + 1
    cfgs = ((4, 1, 0), (4, 1, 150), (4, 1, 260), (4, 1, 400), (3, 1, 260), (4, 2, 260), (4, 1, 600)) if len(sys.argv) < 2 else ((depth, int(sys.argv[3]), int(sys.argv[4])),)
    for depth, slots, mbp in cfgs:
        pipe = pipeline.ScanPipeline(2, 100, device=0, depth=depth, compact=True, compute_slots=slots, trace=True)
        groups = pipeline.group_contigs(range(len(lengths)), lengths, mbp * 1_000_000) if mbp else [[c] for c in range(len(lengths))]
        for rep in range(3):
            pipe.trace.clear()
            T0 = time.perf_counter()
            futs = [pipe.submit_flat(hn[offs[g[0]]:offs[g[-1]] + lengths[g[-1]] + 1], [lengths[c] for c in g], offsets=[offs[c] - offs[g[0]] for c in g]) for g in groups]
            n = 0
            for f in futs:
                r = f.result(); n += sum(len(r[s][0]) for s in range(3))
            torch.cuda.synchronize()
            el = time.perf_counter() - T0
            print("rep %d depth %d slots %d batch %d Mbp (%d batches): %.1f ms, %.2f Gbp/s, records %d" % (rep, depth, slots, mbp, len(groups), el * 1e3, sum(lengths) / el / 1e9, n), flush=True)
        tr = sorted(pipe.trace, key=lambda x: x[2])
        for k, L, t0, t1, t2, t3, t4 in tr:
            print("ctx %d L %9d  load %7.1f..%7.1f  wait ..%7.1f scan ..%7.1f  fetch ..%7.1f ms | load %5.1f scan %5.1f fetch %5.1f" % (
                k, L, (t0 - T0) * 1e3, (t1 - T0) * 1e3, (t2 - T0) * 1e3, (t3 - T0) * 1e3, (t4 - T0) * 1e3, (t1 - t0) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
        pipe.close()
