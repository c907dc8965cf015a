"""BASELINE.json configs[4] shape on the GPU box: n contigs x 1 kb (default 10^6 = 1 Gbp), motif sizes 1..6 (and 2..100 for
comparison): host-memory load vs rb_load_fasta of the same records as FASTA text (one 1000-column line per record), device
time of the scan, candidate counts; streams of both loads must agree.   usage: python tools/exp_c5.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ribbit_b200 import scan, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
t0 = time.time(); contigs = synth.contigs_c5(n=n, length=1000, seed=5); print("gen %.1f s" % (time.time() - t0), flush=True)
text = np.frombuffer(b"".join(b">ctg%d scaffold\n%s\n" % (i, s) for i, s in enumerate(contigs)), dtype=np.uint8)
for (a, b) in [(1, 6), (2, 100)]:
    sc = scan.Scanner(a, b)
    sc.load(contigs)
    t0 = time.perf_counter(); sc.load(contigs); t1 = time.perf_counter()
    for _ in range(2): sc.scan_device()
    t = sc.timing(); c1 = sc.counts()
    print("C5 shape %d x 1 kb, m %d..%d: host load %.1f ms, device %s -> %.2f Gbp/s, counts %s" % (
        n, a, b, (t1 - t0) * 1e3, {k: round(v, 3) if isinstance(v, float) else v for k, v in t.items()}, n * 1000 / t["total_ms"] / 1e6, c1), flush=True)
    sc.load_fasta(text)
    t0 = time.perf_counter(); names, lens = sc.load_fasta(text); t1 = time.perf_counter()
    sc.scan_device()
    assert sc.counts() == c1 and len(names) == n and names[-1] == "ctg%d" % (n - 1) and int(lens.min()) == 1000
    print("   rb_load_fasta of %.0f MB / %d records: %.1f ms (incl. building %d Python names)" % (text.size / 1e6, n, (t1 - t0) * 1e3, n), flush=True)
    sc.close()
