"""Generates tests/golden/golden_motif.npz: checkpoint CP4 of the instrumented reference (oracle/_ref/ribbit_ref_cp) —
the arguments of every mostFrequentLongerMotif call (parse_seed.cpp:153, top-level and recursive) and the row it chose —
on small seeded contigs with planted repeats of motif sizes > 10. Only runs in the build container.
Usage:  python tests/golden/make_golden_motif.py

    <name>_seq   uint8  the contig (ASCII)
    <name>_args  int32  (min_mlen, max_mlen)
    <name>_cp4   int32  (n,4) rows (seed_start, seed_seq_len, mlen, row), in call order
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou  # noqa: E402
from ribbit_b200 import synth  # noqa: E402


def cases():
    out = [("c1", synth.contig_c1(30000, seed=41), 2, 100),
           ("c2n", synth.contig_c2(40000, seed=42, density_per_mbp=600), 2, 100),
           ("c4", synth.contig_c4(30000, seed=43, n_repeats=40), 2, 100),
           ("c2mid", synth.contig_c2(30000, seed=44, density_per_mbp=600), 8, 40)]
    rng = np.random.default_rng(45)
    s = bytearray(synth.fuzz_contig(rng, 6000, 0.0))
    unit = bytes(rng.choice(list(b"ACGT"), 37).astype(np.uint8))
    s[10:10 + 37 * 40] = unit * 40           # a long repeat at the contig start (upstream walk reaches position 0)
    s[3000] = ord("N")
    unit2 = bytes(rng.choice(list(b"ACGT"), 15).astype(np.uint8))
    s[3001:3001 + 15 * 30] = unit2 * 30      # repeat right behind an N (column seed_start-1 is an N)
    s[-13 * 25:] = bytes(rng.choice(list(b"ACGT"), 13).astype(np.uint8)) * 25   # repeat running to the contig end
    out.append(("edges", bytes(s), 2, 100))
    return out


def main():
    assert ou.have_ref(), "oracle/_ref not built (make -C oracle ref)"
    store = {}
    for name, seq, mlo, mhi in cases():
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "x.fa")
            synth.write_fasta(fa, [seq])
            contigs, bed, rc = ou.ref_cp(fa, ["-m", mlo, "-M", mhi])
        cp4 = contigs[0]["cp4"].astype(np.int32)
        store[name + "_seq"] = np.frombuffer(seq, dtype=np.uint8)
        store[name + "_args"] = np.array([mlo, mhi], dtype=np.int32)
        store[name + "_cp4"] = cp4
        print(name, len(seq), "cp4", len(cp4), "rc", rc, "max seed len", cp4[:, 1].max() if len(cp4) else 0)
    np.savez_compressed(os.path.join(HERE, "golden_motif.npz"), **store)


if __name__ == "__main__":
    main()
