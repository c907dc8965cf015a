"""One-off: the CPU emulator (tests/emu) on the gate windows of all 24 C3 contigs against the reference digests
(tests/golden/c3_digests.json). Runs in the build container.   python tools/emu_c3_gate.py [jobs]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from concurrent.futures import ProcessPoolExecutor
import numpy as np


def one(i):
    import emu_util
    from ribbit_b200 import workloads as wl
    g = wl.load_digests()
    seq = wl.c3_contig(i)
    lo, hi = wl.c3_window(i, len(seq))
    got, _ = emu_util.emu_streams(seq[lo:hi], 2, 100, chunk_words=2048)
    bad = 0
    for s in (1, 2, 3):
        r = got[s]
        k = (r[:, 3] & 3) == 0
        d = wl.digest_rows(r[k, 0].astype(np.int64) + lo, r[k, 1].astype(np.int64) + lo, r[k, 2], lo, hi)
        bad += d != g["windows"][str(i)]["kept"][str(s)]
    return i, bad


if __name__ == "__main__":
    jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    with ProcessPoolExecutor(jobs) as ex:
        res = list(ex.map(one, range(24)))
    print("windows with a mismatch:", [i for i, b in res if b], "of", len(res))
