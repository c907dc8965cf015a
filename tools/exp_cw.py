import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ribbit_b200 import scan, synth
L = 46_700_000
seq = synth.contig_c2(L, seed=21)
for cw in (0, 200, 308, 400, 512, 800, 1232):
    sc = scan.Scanner(2, 100, chunk_words=cw); sc.load([seq])
    best = 1e9
    for _ in range(4):
        sc.scan_device(); best = min(best, sc.timing()["scan_ms"])
    print("cw %5d scan %.3f ms restarts %d" % (cw, best, sc.timing()["restarts"]), flush=True)
    sc.close()
