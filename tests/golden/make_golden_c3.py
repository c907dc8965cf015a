"""Golden digests for the bench workload C3 (24 contigs, 3.1 Gbp; ribbit_b200/workloads.py): runs the UNMODIFIED
reference (oracle/_ref/ribbit_ref_cp, stopped after the scan) on one 1 Mbp window of every contig and on the whole
chr21-size contig, and stores md5 digests of its kept CP1 calls in tests/golden/c3_digests.json. bench.py compares its own
streams with these inside every run ("parity": "ok"). Only runs in the build container (needs oracle/_ref).
Usage: python tests/golden/make_golden_c3.py [--jobs 8]"""
import argparse
import json
import os
import sys
import tempfile
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou  # noqa: E402
from ribbit_b200 import synth, workloads as wl  # noqa: E402


def kept_digests(cp1, off, lo=None, hi=None):
    out = {}
    for st in (1, 2, 3):
        r = cp1[cp1[:, 0] == st]
        k = wl.kept_mask(st, r[:, 1], r[:, 2], r[:, 3])
        out[str(st)] = wl.digest_rows(r[k, 1].astype(np.int64) + off, r[k, 2].astype(np.int64) + off, r[k, 3], lo, hi)
    return out


def ref_cp1(seq):
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "x.fa")
        synth.write_fasta(fa, [seq])
        contigs, _, rc = ou.ref_cp(fa, ["-m", "2", "-M", "100"], stop_after_cp2=True, timeout=7200)
    assert rc == 0 and len(contigs) == 1 and contigs[0]["L"] == len(seq), rc
    return contigs[0]["cp1"]


def one(i):
    seq = wl.c3_contig(i)
    L = len(seq)
    lo, hi = wl.c3_window(i, L)
    res = {"L": L, "lo": lo, "hi": hi, "kept": kept_digests(ref_cp1(seq[lo:hi]), lo, lo, hi)}
    full = None
    if i == wl.C3_FULL_CONTIG:
        full = {"contig": i, "L": L, "kept": kept_digests(ref_cp1(seq), 0)}
    return i, res, full


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=8)
    args = ap.parse_args()
    assert ou.have_ref()
    out = {"workload": "C3: contig i = synth.contig_c2(HG38_MBP[i] Mbp, seed=100+i), -m 2 -M 100", "margin": wl.GATE_MARGIN,
           "windows": {}, "full": None}
    order = sorted(range(len(wl.HG38_MBP)), key=lambda i: (i != wl.C3_FULL_CONTIG, -wl.HG38_MBP[i]))
    with ProcessPoolExecutor(args.jobs) as ex:
        for i, res, full in ex.map(one, order):
            out["windows"][str(i)] = res
            if full:
                out["full"] = full
            print(i, res, full, flush=True)
    json.dump(out, open(wl.DIGESTS, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
