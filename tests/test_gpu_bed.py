"""End to end through the drop-in program baseline/_ref/ribbit_gpu (the reference's own main, merges, per-seed stage,
SSW and CIGAR code, with processSequence replaced by ribbit_b200/host/process_sequence_gpu.cpp + libribbit_scan.so):
the merged seed lists (CP2) and the BED bytes must equal what the unmodified reference produced (golden vectors)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from ribbit_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "baseline", "_ref", "ribbit_gpu")


def _run(seq, mlo, mhi, extra=()):
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "x.fa"); bed = os.path.join(td, "o.bed"); cp2 = os.path.join(td, "cp2.bin")
        synth.write_fasta(fa, [seq])
        r = subprocess.run([EXE, "-i", fa, "-o", bed, "-m", str(mlo), "-M", str(mhi), *extra], env=dict(os.environ, RB_CP2_OUT=cp2),
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=600)
        raw = np.fromfile(cp2, dtype=np.int32).reshape(-1, 5) if os.path.exists(cp2) else np.zeros((0, 5), np.int32)
        lists = raw[raw[:, 0] >= 11].copy()
        lists[:, 0] -= 10
        return r.returncode, (open(bed, "rb").read() if os.path.exists(bed) else b""), lists, r.stderr


@pytest.mark.skipif(not os.access(EXE, os.X_OK), reason="baseline/_ref/ribbit_gpu is built in the build container (make -C ribbit_b200/host)")
def test_bed_and_seed_lists_match_the_reference(golden):
    checked = 0
    for name, g in golden.items():
        if int(g["rc"][1]) != 0:
            continue  # the reference itself crashed on this input
        seq = g["seq"].tobytes()
        rc, bed, lists, err = _run(seq, int(g["args"][0]), int(g["args"][1]))
        assert rc == 0, (name, err[-500:])
        assert lists.shape == g["cp2"].shape and (lists == g["cp2"]).all(), "%s: merged seed lists differ" % name
        assert bed == g["bed"].tobytes(), "%s: BED differs" % name
        checked += 1
    assert checked >= 14


@pytest.mark.skipif(not os.access(EXE, os.X_OK), reason="baseline/_ref/ribbit_gpu is built in the build container")
def test_cli_flags_are_the_references(golden):
    # -p is accepted and ignored (the reference never reads it); missing -i is reported the reference's way
    g = golden["fuzz06"]
    seq = g["seq"].tobytes()
    rc, bed, _, _ = _run(seq, 2, 100, ("-p", "0.70"))
    assert rc == 0 and bed == g["bed"].tobytes()
    r = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert b"Please specify an input fasta file" in r.stderr
