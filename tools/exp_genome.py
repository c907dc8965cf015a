"""BASELINE.json configs[2] shape: 24 contigs with hg38-like lengths (3.1 Gbp), one batch on one GPU (device timings)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from ribbit_b200 import scan, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
t0 = time.time()
block = np.frombuffer(synth.contig_c2(60_000_000, seed=100), dtype=np.uint8)   # repeats + N runs
lengths = [int(mb * 1e6 * scale) for mb in synth.HG38_MBP]
total = sum(lengths)
buf = np.empty(total + 64, dtype=np.uint8)
pos = 0
for i, L in enumerate(lengths):
    shift = (i * 7_654_321) % len(block)
    idx = 0
    while idx < L:
        n = min(L - idx, len(block) - shift)
        buf[pos + idx:pos + idx + n] = block[shift:shift + n]
        idx += n; shift = 0
    pos += L
print("generated %d contigs, %.2f Gbp in %.0f s" % (len(lengths), total / 1e9, time.time() - t0), flush=True)
sc = scan.Scanner(2, 100)
t0 = time.perf_counter(); sc.load_flat(buf, lengths); t1 = time.perf_counter()
sc.scan_device(); t = sc.timing()
print("load (pageable H2D + geometry) %.2f s; device: %s; counts %s -> %.2f Gbp/s" % (t1 - t0, t, sc.counts(), total / t["total_ms"] / 1e6), flush=True)
sc.scan_device(); t = sc.timing()
print("second scan: %s -> %.2f Gbp/s" % (t, total / t["total_ms"] / 1e6), flush=True)
