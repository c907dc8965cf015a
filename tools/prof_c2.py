"""Profiling driver: scans the C2 contig (46.7 Mbp, or argv[1] Mbp) a few times; run under ncu.
   python tools/prof_c2.py [mbp] [steps] [m_lo m_hi]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ribbit_b200 import scan, synth
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 46.7
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
seq = synth.contig_c2(int(mbp * 1e6), seed=21)
L = len(seq)
h = torch.empty(L + 64, dtype=torch.uint8, pin_memory=True)
h.numpy()[:L] = np.frombuffer(seq, dtype=np.uint8)
d = h.cuda()
sc = scan.Scanner(2, 100)
sc.load_device(d.data_ptr(), [L], keepalive=d)
for _ in range(steps):
    sc.scan_device()
    print(sc.timing(), sc.counts(), flush=True)
sc.close()
