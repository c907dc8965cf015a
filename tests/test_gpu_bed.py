"""End to end through the drop-in program ribbit_b200/bin/ribbit_gpu (the reference's own main, merges, per-seed stage,
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
EXE = os.path.join(ROOT, "ribbit_b200", "bin", "ribbit_gpu")


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


@pytest.mark.skipif(not os.access(EXE, os.X_OK), reason="ribbit_b200/bin/ribbit_gpu is built in the build container (make -C ribbit_b200/host)")
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


@pytest.mark.skipif(not os.access(EXE, os.X_OK), reason="ribbit_b200/bin/ribbit_gpu is built in the build container")
def test_cli_flags_are_the_references(golden):
    # -p is accepted and ignored (the reference never reads it); missing -i is reported the reference's way
    g = golden["fuzz06"]
    seq = g["seq"].tobytes()
    rc, bed, _, _ = _run(seq, 2, 100, ("-p", "0.70"))
    assert rc == 0 and bed == g["bed"].tobytes()
    r = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert b"Please specify an input fasta file" in r.stderr


def _large_cases():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_large.json")))


@pytest.mark.skipif(not os.access(EXE, os.X_OK), reason="ribbit_b200/bin/ribbit_gpu is built in the build container")
@pytest.mark.parametrize("name", ["c1_1mbp_default", "c1_300k_l12", "c2_1mbp_nruns", "c4_500k_p070"])
def test_mbp_scale_bed_digest_matches_the_reference(name):
    """BASELINE.json config shapes at 0.3-1 Mbp: md5 of the BED and of the merged seed lists vs the unmodified reference
    (tests/golden/make_golden_large.py); the kept candidates of the C-ABI streams vs the reference's call log."""
    import hashlib
    from ribbit_b200 import scan
    c = _large_cases()[name]
    seq = getattr(synth, c["gen"])(**c["kwargs"])
    assert hashlib.md5(seq).hexdigest() == c["seq_md5"], "synthetic input not reproduced"
    mlo, mhi = 2, 100
    # CP1: kept records of the three streams
    sc = scan.Scanner(mlo, mhi)
    sc.load([seq])
    res = sc.scan()
    for s in range(3):
        a, _ = res[s]
        real = a[(a["flags"] & (scan.REC_DROPPED | scan.REC_PSEUDO)) == 0]
        rows = np.stack([real["start"], real["end"], real["mlen"].astype(np.int32)], axis=1).astype("<i4")
        md5, n = c["cp1_kept"][str(s + 1)]
        assert len(rows) == n and hashlib.md5(rows.tobytes()).hexdigest() == md5, "stream %d" % s
    sc.close()
    # CP2 + CP3 through the drop-in program
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "x.fa"); bed = os.path.join(td, "o.bed"); cp2 = os.path.join(td, "cp2.bin")
        synth.write_fasta(fa, [seq])
        r = subprocess.run([EXE, "-i", fa, "-o", bed, *c["flags"]], env=dict(os.environ, RB_CP2_OUT=cp2),
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=1200)
        assert r.returncode == 0, r.stderr[-500:]
        raw = np.fromfile(cp2, dtype=np.int32).reshape(-1, 5)
        lists = raw[raw[:, 0] >= 11].copy()
        lists[:, 0] -= 10
        assert len(lists) == c["cp2_rows"] and hashlib.md5(lists.astype("<i4").tobytes()).hexdigest() == c["cp2_md5"]
        b = open(bed, "rb").read()
        assert b.count(b"\n") == c["bed_rows"] and hashlib.md5(b).hexdigest() == c["bed_md5"]
        if name == "c2_1mbp_nruns":
            # the per-seed stage runs in forked worker processes by default: one process and an odd number of workers give
            # the same bytes
            for procs in ("1", "3"):
                r = subprocess.run([EXE, "-i", fa, "-o", bed, *c["flags"]], env=dict(os.environ, RIBBIT_HOST_PROCS=procs, RIBBIT_VERBOSE="1"),
                                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=1200)
                assert r.returncode == 0, r.stderr[-500:]
                assert ("per-seed stage on %s process" % procs).encode() in r.stderr
                assert open(bed, "rb").read() == b
