"""K0 — FASTA text parsed on the device (rb_load_fasta) against the host reader ribbit_b200/fasta.py, which restates the
reference's getline loop (ribbit.cpp:269-280; its quirks are pinned in tests/test_fasta.py): record names and lengths, the
packed planes of every record against the oracle's pack of the expected sequence (bit-exact), and the candidate streams
against a load of the same sequences from host memory."""
import os
import tempfile
import time

import numpy as np
import pytest

import oracle_util as ou
from fasta_cases import QUIRKS, random_fasta
from ribbit_b200 import fasta, scan, synth

pytestmark = pytest.mark.gpu


def host_records(text: bytes):
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "x.fa")
        open(p, "wb").write(text)
        return fasta.read_fasta(p)


def check(sc, text: bytes, planes=True):
    names, seqs = host_records(text)
    got_names, got_len = sc.load_fasta(text)
    assert got_names == names
    assert got_len.tolist() == [len(s) for s in seqs]
    sc.scan_device()
    if planes:
        for i, s in enumerate(seqs):
            hi, lo, nn = sc.planes(i)
            ehi, elo, enn = ou.pack(s)
            assert np.array_equal(hi, ehi) and np.array_equal(lo, elo) and np.array_equal(nn, enn), (i, names[i])
    return names, seqs


def test_reader_quirks_on_device():
    sc = scan.Scanner(2, 30)
    for text in QUIRKS:
        check(sc, text)


def test_device_reader_equals_the_reference_reader():
    """rb_load_fasta against the records the unmodified reference read from the same bytes (golden_fasta.json)."""
    from test_fasta import check_against_reference
    sc = scan.Scanner(2, 6)

    def fn(text):
        names, lens = sc.load_fasta(text)
        return names, lens.tolist()
    assert check_against_reference(fn) > 100


def test_random_fasta_files():
    rng = np.random.default_rng(2027)
    sc = scan.Scanner(2, 24)
    for it in range(40):
        text = random_fasta(rng, int(rng.integers(1, 30)), int(rng.choice([50, 3000, 40000])), crlf=(it % 7 == 3))
        check(sc, text)


def test_many_short_records_and_streams():
    """BASELINE.json configs[4] shape (many 1 kb records): streams after load_fasta == streams after a host-memory load."""
    rng = np.random.default_rng(5)
    seqs = synth.contigs_c5(3000, 1000, seed=5)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "x.fa")
        synth.write_fasta(p, seqs, names=["ctg%d desc" % i for i in range(len(seqs))])
        text = open(p, "rb").read()
    sc = scan.Scanner(1, 6)
    names, got = check(sc, text, planes=False)
    assert names == ["ctg%d" % i for i in range(len(seqs))] and got == list(seqs)
    a = sc.fetch()
    sc.load(seqs)
    b = sc.scan()
    for s in range(3):
        assert np.array_equal(a[s][0], b[s][0]) and np.array_equal(a[s][1], b[s][1])


def test_large_contig_and_timing():
    seq = synth.contig_c2(6_000_000, seed=3)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "x.fa")
        synth.write_fasta(p, [seq, seq[:1_000_001]], names=["chrA", "chrB"])
        text = open(p, "rb").read()
    sc = scan.Scanner(2, 100)
    buf = np.frombuffer(text, np.uint8)
    sc.load_fasta(buf)
    t = time.time()
    names, lens = sc.load_fasta(buf)
    dt = time.time() - t
    assert names == ["chrA", "chrB"] and lens.tolist() == [6_000_000, 1_000_001]
    sc.scan_device()
    a = sc.fetch()
    for i, s in enumerate((seq, seq[:1_000_001])):
        hi, lo, nn = sc.planes(i)
        ehi, elo, enn = ou.pack(s)
        assert np.array_equal(hi, ehi) and np.array_equal(lo, elo) and np.array_equal(nn, enn)
    sc.load([seq, seq[:1_000_001]])
    b = sc.scan()
    for s in range(3):
        assert np.array_equal(a[s][0], b[s][0]) and np.array_equal(a[s][1], b[s][1])
    print("rb_load_fasta: %.1f MB in %.2f ms (pageable host memory, incl. H2D)" % (len(text) / 1e6, dt * 1e3))
