"""Golden digests for Mbp-scale inputs: runs the UNMODIFIED reference (oracle/_ref/ribbit_ref_cp) on seeded synthetic
contigs of the BASELINE.json shapes and stores md5 digests of its BED output and merged seed lists in
tests/golden/golden_large.json. The inputs are regenerated from their seeds at test time (ribbit_b200/synth.py).
Only runs in the build container. Usage: python tests/golden/make_golden_large.py"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou  # noqa: E402
from ribbit_b200 import synth  # noqa: E402

CASES = {
    # name: (generator, kwargs, extra command-line flags)
    "c1_1mbp_default": ("contig_c1", {"L": 1_000_000, "seed": 20261018}, []),
    "c1_300k_l12": ("contig_c1", {"L": 300_000, "seed": 7}, ["-l", "12"]),
    "c2_1mbp_nruns": ("contig_c2", {"L": 1_000_000, "seed": 21}, []),
    "c4_500k_p070": ("contig_c4", {"L": 500_000, "seed": 44, "n_repeats": 650}, ["-p", "0.70", "-M", "100"]),
}


def make_input(gen, kwargs):
    return getattr(synth, gen)(**kwargs)


def main():
    assert ou.have_ref()
    out = {}
    for name, (gen, kwargs, flags) in CASES.items():
        seq = make_input(gen, kwargs)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "x.fa")
            synth.write_fasta(fa, [seq])
            contigs, bed, rc = ou.ref_cp(fa, flags)
        cp2 = contigs[0]["cp2"]
        import numpy as np
        import stream_model as sm
        cp1 = contigs[0]["cp1"]
        kept_md5 = {}
        for st, cut in ((1, lambda m: 0), (2, sm.cut_subst), (3, sm.cut_anch)):
            rows = cp1[cp1[:, 0] == st][:, 1:4]
            cuts = np.array([cut(int(m)) for m in range(0, 1100)])
            keep = (rows[:, 1] - rows[:, 0]) >= cuts[rows[:, 2]]
            kept_md5[str(st)] = [hashlib.md5(rows[keep].astype("<i4").tobytes()).hexdigest(), int(keep.sum())]
        out[name] = {"gen": gen, "kwargs": kwargs, "flags": flags, "rc": rc, "seq_md5": hashlib.md5(seq).hexdigest(),
                     "bed_md5": hashlib.md5(bed).hexdigest(), "bed_rows": bed.count(b"\n"),
                     "cp2_md5": hashlib.md5(cp2.astype("<i4").tobytes()).hexdigest(), "cp2_rows": int(len(cp2)),
                     "cp1_md5": hashlib.md5(contigs[0]["cp1"].astype("<i4").tobytes()).hexdigest(), "cp1_rows": int(len(contigs[0]["cp1"])), "cp1_kept": kept_md5}
        print(name, out[name])
    json.dump(out, open(os.path.join(HERE, "golden_large.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
