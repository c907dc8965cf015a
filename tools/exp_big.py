"""Genome-shape runs: one chr1-size contig, and the 24-contig C3 shape (scaled) through the pipeline."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ribbit_b200 import scan, synth, pipeline
t0 = time.time()
big = synth.contig_c2(248_000_000, seed=100, density_per_mbp=100)
print("generated 248 Mbp in %.1f s" % (time.time() - t0), flush=True)
sc = scan.Scanner(2, 100); sc.load([big])
for _ in range(2): sc.scan_device()
t = sc.timing(); print("chr1-size contig:", t, sc.counts(), "-> %.2f Gbp/s device" % (len(big) / t["total_ms"] / 1e6), flush=True)
res = sc.fetch(copy=False)
for s in range(3):
    a = res[s][0]; real = a[(a["flags"] & 2) == 0]
    assert (np.diff(real["time"].astype(np.int64)) >= 0).all()
sc.close()
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
contigs = [big[:int(mb * 1e6 * scale)] for mb in synth.HG38_MBP]   # C3 shape: 24 contigs with hg38-like lengths
bufs = [np.frombuffer(c + b"\0", dtype=np.uint8) for c in contigs]
pipe = pipeline.ScanPipeline(2, 100, depth=2)
for f in [pipe.submit_flat(b, [len(c)]) for b, c in zip(bufs[:4], contigs[:4])]: f.result()
t0 = time.perf_counter()
futs = [pipe.submit_flat(b, [len(c)]) for b, c in zip(bufs, contigs)]
n = 0
for f in futs:
    r = f.result(); n += sum(len(r[s][0]) for s in range(3))
dt = time.perf_counter() - t0
tot = sum(len(c) for c in contigs)
print("C3 shape x%.2f: %d contigs, %.1f Mbp, %d records, %.1f ms end to end (pageable host buffers) -> %.2f Gbp/s" % (scale, len(contigs), tot / 1e6, n, dt * 1e3, tot / dt / 1e9), flush=True)
pipe.close()
