"""Ad-hoc timing experiments on the GPU box (not part of the product): scan time vs input shape / chunking."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ribbit_b200 import scan, synth

def run(name, seq, cw=0, reps=3):
    sc = scan.Scanner(2, 100, chunk_words=cw)
    sc.load([seq])
    best = None
    for _ in range(reps):
        sc.scan_device()
        t = sc.timing()
        if best is None or t["scan_ms"] < best["scan_ms"]:
            best = t
    print("%-28s L=%9d cw=%5d scan %.3f ms merge %.3f restarts %d  counts %s" % (name, len(seq), cw, best["scan_ms"], best["merge_ms"], best["restarts"], sc.counts()), flush=True)
    sc.close()

L = int(sys.argv[1]) if len(sys.argv) > 1 else 46_700_000
a = synth.contig_c2(L, seed=21)
b = synth.contig_c2(L, seed=21, n_runs=False)
rng = np.random.default_rng(1)
c = synth.random_bases(rng, L).tobytes()
run("c2 with N runs", a)
run("c2 without N runs", b)
run("pure random", c)
for cw in (64, 128, 256, 512, 1024):
    run("c2 without N runs", b, cw)
    run("c2 with N runs", a, cw)
