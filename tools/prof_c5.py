"""Profiling driver: C5 shape (n contigs x 1 kb, -m 1 -M 6).  python tools/prof_c5.py [n_contigs] [steps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ribbit_b200 import scan, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
flat = b"".join(synth.contigs_c5(n=n, length=1000, seed=5))
h = torch.empty(len(flat) + 64, dtype=torch.uint8, pin_memory=True)
h.numpy()[:len(flat)] = np.frombuffer(flat, dtype=np.uint8)
d = h.cuda()
sc = scan.Scanner(1, 6, debug=int(os.environ.get("RB_DEBUG", "0")))
sc.load_device(d.data_ptr(), [1000] * n, keepalive=d)
for _ in range(steps):
    sc.scan_device()
    t = sc.timing()
    print(t, sc.counts(), "Gbp/s %.1f" % (len(flat) / t["total_ms"] / 1e6), flush=True)
sc.close()
