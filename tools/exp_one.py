"""One scan of one synthetic input (for ncu captures): python tools/exp_one.py random|c2|c2n [L]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ribbit_b200 import scan, synth
mode = sys.argv[1]
L = int(sys.argv[2]) if len(sys.argv) > 2 else 46_700_000
if mode == "random":
    seq = synth.random_bases(np.random.default_rng(1), L).tobytes()
elif mode == "c2":
    seq = synth.contig_c2(L, seed=21)
else:
    seq = synth.contig_c2(L, seed=21, n_runs=False)
sc = scan.Scanner(2, 100)
sc.load([seq])
for _ in range(2):
    sc.scan_device()
print(mode, sc.timing(), sc.counts())
