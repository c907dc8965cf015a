import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ribbit_b200 import scan, synth
L = 46_700_000
seq = synth.contig_c2(L, seed=21)
for (a, b) in [(2, 100), (2, 26), (27, 51), (52, 76), (77, 100), (2, 6), (7, 11), (12, 26)]:
    sc = scan.Scanner(a, b); sc.load([seq])
    for _ in range(3): sc.scan_device()
    t = sc.timing()
    print("m %3d..%3d  scan %.3f ms merge %.3f  counts %s" % (a, b, t["scan_ms"], t["merge_ms"], sc.counts()), flush=True)
    sc.close()
