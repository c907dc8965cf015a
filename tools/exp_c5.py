import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ribbit_b200 import scan, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
t0 = time.time(); contigs = synth.contigs_c5(n=n, length=1000, seed=5); print("gen %.1f s" % (time.time() - t0), flush=True)
for (a, b) in [(1, 6), (2, 100)]:
    sc = scan.Scanner(a, b)
    t0 = time.perf_counter(); sc.load(contigs); t1 = time.perf_counter()
    for _ in range(2): sc.scan_device()
    t = sc.timing()
    print("C5 shape %d x 1 kb, m %d..%d: load %.1f ms, device %s -> %.2f Gbp/s, counts %s" % (n, a, b, (t1 - t0) * 1e3, {k: round(v, 3) if isinstance(v, float) else v for k, v in t.items()}, n * 1000 / t["total_ms"] / 1e6, sc.counts()), flush=True)
    sc.close()
