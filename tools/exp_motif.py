"""K7 measurement (GPU box): rb_motif_rows on the calls the reference itself makes.

Runs the instrumented reference (oracle/_ref/ribbit_ref_cp, CP4 logging) on a C1-shape contig to collect every
mostFrequentLongerMotif call (top-level and recursive), then times
  * the reference's per-seed stage with those calls inside (wall of the reference run minus its scan stage), as context,
  * the plain-C oracle port on the same calls (1 core),
  * rb_motif_rows on the whole batch (CUDA events around the kernel are not exposed; wall time of the call incl. H2D of
    the seed list and D2H of the results, best of 5),
and checks the rows against the reference's.   usage: python tools/exp_motif.py [Mbp]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou  # noqa: E402
from ribbit_b200 import scan, synth  # noqa: E402


def main():
    mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    shape = sys.argv[2] if len(sys.argv) > 2 else "c1"
    L = int(mbp * 1e6)
    seq = synth.contig_c1(L, seed=7) if shape == "c1" else synth.contig_c2(L, seed=7)
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "x.fa")
        synth.write_fasta(fa, [seq])
        t = time.time()
        contigs, bed, rc = ou.ref_cp(fa, ["-m", 2, "-M", 100])
        t_ref = time.time() - t
        t = time.time()
        ou.ref_cp(fa, ["-m", 2, "-M", 100], stop_after_cp2=True)
        t_scan = time.time() - t
    cp4 = contigs[0]["cp4"].astype(np.int32)
    seeds = np.stack([np.zeros(len(cp4), np.int32), cp4[:, 0], cp4[:, 0] + cp4[:, 1], cp4[:, 2]], axis=1)
    ln = cp4[:, 1].astype(np.int64)
    work = int((5 * ln * ln).sum())
    print(f"{shape} {mbp} Mbp: reference total {t_ref:.1f} s, scan stage {t_scan:.1f} s, per-seed stage {t_ref - t_scan:.1f} s; "
          f"{len(cp4)} mostFrequentLongerMotif calls, seed length median {int(np.median(ln))} max {int(ln.max())}, "
          f"~{work / 1e9:.2f} G dot-matrix probes")
    n_cpu = min(len(cp4), 5000)
    t = time.time()
    rows_cpu = np.array([ou.motif_row(seq, int(s), int(l), int(m))[0] for s, l, m, _ in cp4[:n_cpu]], np.int32)
    t_cpu = (time.time() - t) * len(cp4) / n_cpu
    assert np.array_equal(rows_cpu, cp4[:n_cpu, 3])
    sc = scan.Scanner(2, 100)
    sc.load([seq]); sc.scan_device()
    best = 1e9
    for _ in range(5):
        t = time.time()
        out = sc.motif_rows(seeds)
        best = min(best, time.time() - t)
    ok = bool(np.array_equal(out[:, 0], cp4[:, 3]))
    res = {"shape": shape, "mbp": mbp, "calls": int(len(cp4)), "rows_equal_reference": ok, "gpu_batch_ms": best * 1e3,
           "gpu_calls_per_s": len(cp4) / best, "oracle_port_1core_s": t_cpu, "oracle_calls_per_s": len(cp4) / t_cpu,
           "reference_per_seed_stage_s": t_ref - t_scan, "reference_scan_stage_s": t_scan,
           "dot_probes": work, "gpu_probes_per_s": work / best}
    print(json.dumps(res))
    assert ok


if __name__ == "__main__":
    main()
