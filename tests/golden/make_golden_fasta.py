"""Golden records of the reference's FASTA reader (ribbit.cpp:269-280): runs the UNMODIFIED reference
(oracle/_ref/ribbit_ref_cp, stopped after the scan of every record) on the quirk cases of tests/fasta_cases.py and on 40
seeded random files, and stores what it read: the record names it prints ("Processing sequence <name>", every record but
the last) and the length of every sequence it hands to processSequence (checkpoint tag 0, oracle/cp_hooks.h). Where the
reference dies on an input the records up to that point are stored. Only runs in the build container.
Usage: python tests/golden/make_golden_fasta.py"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_util as ou  # noqa: E402
from fasta_cases import QUIRKS, golden_random_texts  # noqa: E402


def reference_records(text):
    with tempfile.TemporaryDirectory() as td:
        fa, cp = os.path.join(td, "x.fa"), os.path.join(td, "cp.bin")
        open(fa, "wb").write(text)
        env = dict(os.environ, RB_CP_OUT=cp, RB_CP_STOP_AFTER_CP2="1")
        r = subprocess.run([ou.REF_CP_BIN, "-i", fa, "-o", os.path.join(td, "o.bed"), "-m", "2", "-M", "6"], env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=600)
        raw = np.fromfile(cp, dtype=np.int32) if os.path.exists(cp) else np.zeros(0, np.int32)
        raw = raw[: (raw.size // 5) * 5].reshape(-1, 5)
        lengths = raw[raw[:, 0] == 0][:, 2].tolist()
        names = [l[len(b"Processing sequence "):].decode(errors="replace") for l in r.stderr.split(b"\n") if l.startswith(b"Processing sequence ")]
        return {"rc": r.returncode, "names_but_last": names, "lengths": lengths}


def main():
    assert ou.have_ref()
    out = {"quirks": [], "random": []}
    for t in QUIRKS:
        out["quirks"].append(dict(reference_records(t), md5=hashlib.md5(t).hexdigest()))
    for t in golden_random_texts():
        out["random"].append(dict(reference_records(t), md5=hashlib.md5(t).hexdigest()))
    json.dump(out, open(os.path.join(HERE, "golden_fasta.json"), "w"), indent=0)
    print("quirks", [(g["rc"], g["names_but_last"], g["lengths"]) for g in out["quirks"]])
    print("random rc", [g["rc"] for g in out["random"]], "records", sum(len(g["lengths"]) for g in out["random"]))


if __name__ == "__main__":
    main()
