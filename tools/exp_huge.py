"""Sanity at the upper end of contig sizes: one contig of ~1.07 Gbp (positions beyond 2^30), invariants + a window near the
end against the oracle."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_util as ou, stream_model as sm
from ribbit_b200 import scan, synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_100_000_000
t0 = time.time()
block = np.frombuffer(synth.contig_c2(50_000_000, seed=9, n_runs=False), dtype=np.uint8)
seq = np.tile(block, L // len(block) + 1)[:L].copy()
seq[L - 7_000_000:L - 6_990_000] = ord("N")          # an N run near the end
seq = seq.tobytes()
print("generated %.2f Gbp in %.0f s" % (L / 1e9, time.time() - t0), flush=True)
sc = scan.Scanner(2, 100); sc.load([seq]); sc.scan_device(); t = sc.timing()
print("scan:", t, sc.counts(), "%.2f Gbp/s" % (L / t["total_ms"] / 1e6), flush=True)
res = sc.fetch(copy=False)
a0, a1, ctx = L - 7_050_000, L - 6_950_000, 5000
sub = seq[a0 - ctx:a1 + ctx]
exp = sm.expected_streams(sub, ou.scan_events(sub, 2, 100))
for s in range(3):
    a = res[s][0]
    real = a[(a["flags"] & 2) == 0]
    assert (np.diff(real["time"].astype(np.int64)) >= 0).all(), "order"
    mine = real[(real["start"] >= a0) & (real["end"] < a1) & (real["flags"] == 0)]
    rows = np.stack([mine["start"].astype(np.int64) - (a0 - ctx), mine["end"].astype(np.int64) - (a0 - ctx), mine["mlen"]], axis=1)
    e = exp[s + 1]; e = e[(e[:, 3] == 0) & (e[:, 0] >= ctx) & (e[:, 1] < ctx + (a1 - a0))][:, :3]
    print("stream", s, "window rows", len(rows), "oracle", len(e), "equal", len(rows) == len(e) and bool((rows == e).all()), flush=True)
