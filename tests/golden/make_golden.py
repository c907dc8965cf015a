"""Generates the golden fixtures of tests/golden/ by running the UNMODIFIED reference (oracle/_ref/ribbit_ref_cp:
/root/reference sources + Boost stand-in + CP1/CP2 logging, see oracle/instrument.sh) on small seeded inputs.
Only runs in the build container (needs oracle/_ref). Usage:  python tests/golden/make_golden.py

Each case is stored in golden.npz as
    <name>_seq   uint8   the contig (ASCII)
    <name>_args  int32   (min_mlen, max_mlen)
    <name>_cp1   int32   (n,4) rows (stream 1..3, start, end, mlen): argument sequence of addSeedToSeedPositions*
    <name>_cp2   int32   (n,5) rows (list 1..3, start, end, mlen, rank): the three seed lists after all passes
    <name>_rc    int32   exit status of the reference (139/-11 = the reference's own segfault, SURVEY.md F6;
                         cp1 is then the prefix logged before the crash)
and the BED bytes of the full (un-instrumented stop) run in <name>_bed.
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
    rng = np.random.default_rng(20261018)
    out = []
    for i, (L, nd, mlo, mhi) in enumerate([
            (0, 0, 2, 100), (5, 0, 2, 100), (9, 0.1, 1, 6), (30, 0, 2, 24), (200, 0.02, 2, 100), (200, 0, 1, 6),
            (1500, 0.002, 2, 100), (1500, 0.02, 5, 30), (1500, 0.1, 2, 14), (4000, 0, 2, 100), (4000, 0.002, 2, 100),
            (4000, 0.02, 1, 6), (4000, 0.02, 3, 10), (9000, 0.002, 2, 100), (9000, 0, 2, 8)]):
        out.append(("fuzz%02d" % i, synth.fuzz_contig(rng, L, nd), mlo, mhi))
    # hand-built N stress: N at 0..2, isolated Ns, a 400-base N run, IUPAC / lower case at the tail
    s = bytearray(synth.fuzz_contig(rng, 5000, 0.0))
    s[0:3] = b"NNN"; s[100] = ord("N"); s[333] = ord("n"); s[1200:1600] = b"N" * 400
    rep = (b"ACGGT" * 30)
    s[1600:1600 + len(rep)] = rep          # repeat starting right after the N run
    s[1050:1200] = (b"AC" * 75)            # repeat running into the N run
    s[-3:] = b"Ryn"
    out.append(("nstress", bytes(s), 2, 100))
    # tail quirk: contig ending in a poly-A / repeat so that the zero shift-in matches
    t = synth.fuzz_contig(rng, 3000, 0.0) + b"A" * 150
    out.append(("tailA", t, 2, 100))
    t2 = synth.fuzz_contig(rng, 2000, 0.0) + b"CAG" * 60
    out.append(("tailCAG", t2, 2, 100))
    return out


def main():
    assert ou.have_ref(), "oracle/_ref not built (make -C oracle ref)"
    store = {}
    for name, seq, mlo, mhi in cases():
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "x.fa")
            synth.write_fasta(fa, [seq])
            contigs, _, rc = ou.ref_cp(fa, ["-m", mlo, "-M", mhi], stop_after_cp2=True)
            _, bed, rc_full = ou.ref_cp(fa, ["-m", mlo, "-M", mhi])
        c = contigs[0] if contigs else {"cp1": np.zeros((0, 4), np.int32), "cp2": np.zeros((0, 5), np.int32)}
        store[name + "_seq"] = np.frombuffer(seq, dtype=np.uint8)
        store[name + "_args"] = np.array([mlo, mhi], dtype=np.int32)
        store[name + "_cp1"] = c["cp1"].astype(np.int32)
        store[name + "_cp2"] = c["cp2"].astype(np.int32)
        store[name + "_rc"] = np.array([rc, rc_full], dtype=np.int32)
        store[name + "_bed"] = np.frombuffer(bed, dtype=np.uint8)
        print(name, len(seq), "cp1", len(c["cp1"]), "cp2", len(c["cp2"]), "rc", rc, rc_full, "bed bytes", len(bed))
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **store)


if __name__ == "__main__":
    main()
