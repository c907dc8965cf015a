"""e2e breakdown: rb_load_contigs / rb_scan_device / rb_fetch wall times with a pinned host input."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ribbit_b200 import scan, synth
L = 46_700_000
seq = synth.contig_c2(L, seed=21)
host = torch.empty(L + 64, dtype=torch.uint8).pin_memory()
host[:L] = torch.frombuffer(bytearray(seq), dtype=torch.uint8)
hn = host.numpy()
sc = scan.Scanner(2, 100)
for it in range(4):
    t0 = time.perf_counter(); sc.load_flat(hn[:L + 1], [L]); t1 = time.perf_counter()
    sc.scan_device(); t2 = time.perf_counter()
    res = sc.fetch(copy=False); t3 = time.perf_counter()
    print("load %.2f ms  scan_device %.2f ms  fetch %.2f ms  total %.2f ms  (device %.2f)" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t3-t0)*1e3, sc.timing()["total_ms"]))
