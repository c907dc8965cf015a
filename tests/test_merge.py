"""The restated host merges (ribbit_b200/host/seed_merge.cpp: addSeedToSeedPositions{Perfect,Substitutions,Anchored} +
mergeAllLists) against checkpoint CP2 of the unmodified reference: the three seed lists after all
passes, entry for entry and rank for rank. The merges consume the candidate streams in the library's encoding (DROPPED /
PSEUDO / NOCOMMIT records), produced here from the oracle's CP1 events by tests/stream_model.py."""
import hashlib
import json
import os
import tempfile

import numpy as np
import pytest

import merge_util as mu
import oracle_util as ou
from ribbit_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_merges_equal_reference_cp2_on_golden_vectors(golden):
    n = 0
    for name, g in golden.items():
        if int(g["rc"][0]) != 0:
            continue  # the reference died on this input before it dumped its lists
        seq = g["seq"].tobytes()
        got = mu.merged_from_oracle(seq, int(g["args"][0]), int(g["args"][1]))
        assert got.shape == g["cp2"].shape and (got == g["cp2"]).all(), name
        n += len(got)
    assert n > 5000


def test_merges_equal_reference_cp2_digests_at_mbp_scale():
    """tests/golden/golden_large.json: md5 of the reference's CP2 dump on seeded synthetic contigs of the BASELINE shapes."""
    gl = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_large.json")))
    for name in ("c1_300k_l12", "c4_500k_p070"):
        g = gl[name]
        seq = getattr(synth, g["gen"])(**g["kwargs"])
        assert hashlib.md5(seq).hexdigest() == g["seq_md5"]
        m_hi = 100
        got = mu.merged_from_oracle(seq, 2, m_hi)
        assert len(got) == g["cp2_rows"], name
        assert hashlib.md5(got.astype("<i4").tobytes()).hexdigest() == g["cp2_md5"], name


def test_empty_substitution_list_counts_as_exhausted():
    """The one intentional divergence: the reference dereferences the empty substitution list (merge_types.cpp:47-50) and
    crashes; here the anchored candidate is merged against the other two lists."""
    seq = synth.contigs_c5(n=6, length=1000, seed=5)[1]
    got = mu.merged_from_oracle(seq, 1, 6)
    assert (got[:, 0] == 3).any()


@pytest.mark.skipif(not ou.have_ref(), reason="oracle/_ref is built in the build container")
def test_merges_equal_live_reference_on_fresh_inputs():
    rng = np.random.default_rng(77)
    checked = 0
    for it in range(10):
        L = int(rng.choice([2000, 20000, 60000]))
        mlo, mhi = [(2, 100), (2, 24), (5, 30), (1, 100)][it % 4]
        seq = synth.fuzz_contig(rng, L, float(rng.choice([0, 0.001])), m_range=(mlo, min(mhi, 60))) if it % 2 else \
            synth.contig_c2(L, seed=int(rng.integers(1 << 30)), density_per_mbp=3000)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "x.fa")
            synth.write_fasta(fa, [seq])
            contigs, _, rc = ou.ref_cp(fa, ["-m", mlo, "-M", mhi], stop_after_cp2=True)
        if rc != 0:
            continue
        got = mu.merged_from_oracle(seq, mlo, mhi)
        assert got.shape == contigs[0]["cp2"].shape and (got == contigs[0]["cp2"]).all(), it
        checked += 1
    assert checked >= 5
