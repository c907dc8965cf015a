import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ribbit_b200 import pipeline, synth
L = 46_700_000
seq = synth.contig_c2(L, seed=21)
host = torch.empty(L + 64, dtype=torch.uint8).pin_memory()
host[:L] = torch.frombuffer(bytearray(seq), dtype=torch.uint8)
hn = host.numpy()
for depth in (1, 2, 3, 4):
    pipe = pipeline.ScanPipeline(2, 100, depth=depth)
    for f in [pipe.submit_flat(hn[:L + 1], [L]) for _ in range(2 * depth)]: f.result()
    t0 = time.perf_counter()
    futs = [pipe.submit_flat(hn[:L + 1], [L]) for _ in range(12)]
    for f in futs: f.result()
    dt = time.perf_counter() - t0
    print("depth %d: %.2f ms/step  %.2f Gbp/s" % (depth, dt / 12 * 1e3, 12 * L / dt / 1e9), flush=True)
    pipe.close()
