"""Where an end-to-end step spends its time when `depth` contexts are pipelined: per-phase wall times of every step
(load = H2D + geometry, scan = kernels incl. the totals sync, fetch = compaction to 8-byte records + D2H)."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from ribbit_b200 import scan, synth
L = 46_700_000
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
seq = synth.contig_c2(L, seed=21)
host = torch.empty(L + 64, dtype=torch.uint8).pin_memory()
host[:L] = torch.frombuffer(bytearray(seq), dtype=torch.uint8)
hn = host.numpy()
scs = [scan.Scanner(2, 100) for _ in range(depth)]
pools = [ThreadPoolExecutor(max_workers=1) for _ in range(depth)]
log = []
def run(k, i):
    sc = scs[k]
    t0 = time.perf_counter(); sc.load_flat(hn[:L + 1], [L])
    t1 = time.perf_counter(); sc.scan_device()
    t2 = time.perf_counter(); sc.fetch_compact(copy=False)
    t3 = time.perf_counter()
    log.append((i, k, t0, t1, t2, t3, sc.timing()["total_ms"]))
for r in range(2):
    log.clear()
    T0 = time.perf_counter()
    futs = [pools[i % depth].submit(run, i % depth, i) for i in range(15)]
    for f in futs: f.result()
    T1 = time.perf_counter()
print("depth %d: %.2f ms/step" % (depth, (T1 - T0) / 15 * 1e3))
for i, k, t0, t1, t2, t3, dev in sorted(log)[3:12]:
    print("step %2d ctx %d: start %+7.2f  load %.2f  scan %.2f (device %.2f)  fetch %.2f  total %.2f" % (
        i, k, (t0 - T0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, dev, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
