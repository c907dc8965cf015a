"""CPU fuzz campaign: the kernels' lane logic and control flow (tests/emu) against the oracle, same case distribution as
tools/fuzz_gpu.py (tiny chunk sizes included).   python tools/fuzz_emu.py [seed] [seconds]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import emu_util as eu, oracle_util as ou, stream_model as sm
from ribbit_b200 import synth
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
rng = np.random.default_rng(seed)
t0 = time.time(); n = 0; bad = 0
ranges = [(2, 100), (1, 6), (2, 24), (5, 30), (3, 10), (40, 100), (1, 100), (2, 8), (90, 100), (2, 150)]
while time.time() - t0 < budget:
    L = int(rng.choice([1, 7, 8, 31, 32, 33, 63, 64, 65, 100, 500, 2000, 7000]))
    nd = float(rng.choice([0, 0, 0.001, 0.01, 0.05, 0.3]))
    mlo, mhi = ranges[int(rng.integers(len(ranges)))]
    seq = synth.fuzz_contig(rng, L, nd, m_range=(mlo, min(mhi, 60)))
    if rng.random() < 0.3 and L > 200:
        b = bytearray(seq); a = int(rng.integers(0, L - 100)); k = int(rng.integers(50, min(3000, L - a)))
        if rng.random() < 0.5: b[a:a + k] = b"N" * k
        else:
            m = int(rng.integers(1, 40)); b[a:a + k] = (bytes(synth.random_bases(rng, m)) * (k // m + 1))[:k]
        seq = bytes(b)
    cw = int(rng.choice([1 << 30, 1, 2, 3, 5, 17, 64]))
    exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
    got, _ = eu.emu_streams(seq, mlo, mhi, chunk_words=cw)
    n += 1
    for s in (1, 2, 3):
        if got[s].shape != exp[s].shape or not (got[s] == exp[s]).all():
            bad += 1
            open("/tmp/fuzz_emu_fail_%d_%d.txt" % (seed, n), "wb").write(b"%d %d %d\n" % (mlo, mhi, cw) + seq)
            print("MISMATCH case", n, "L", L, "nd", nd, "m", mlo, mhi, "cw", cw, "stream", s, flush=True)
            break
print("emu fuzz seed %d: %d cases, %d mismatches, %.0f s" % (seed, n, bad, time.time() - t0))
