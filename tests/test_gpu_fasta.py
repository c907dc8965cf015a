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
    for text in (b">a desc more\nACGT\nAC\n>b\tx\nGG\r\n>c\n>d e\nTT\n", b"ACGT\nAC", b"", b"\n", b"\n\n>x\n", b">only\n\nAC\n\nGT\n",
                 b">h1\n>h2\n>h3", b"AC>GT\n>n\nA>C\n>", b">\nACGT\n> spaced name\nGG\n", b"ACGT\n>late header\nTTTT\nGG",
                 b">a\nACGT\n>b\n\n>c\nGG\n>d\n"):
        check(sc, text)


def random_fasta(rng, n_records, max_len, crlf=False):
    parts = []
    if rng.random() < 0.3:
        parts.append(synth.fuzz_contig(rng, int(rng.integers(1, 200)), 0.01) + b"\n")     # sequence in front of the first header
    for r in range(n_records):
        hdr = b">rec%d" % r
        k = rng.random()
        if k < 0.3:
            hdr += b" some description > with a bracket"
        elif k < 0.4:
            hdr += b"_" + b"x" * int(rng.integers(3000, 13000))                              # header spanning several 4 KiB tiles
        parts.append(hdr + (b"\r\n" if crlf else b"\n"))
        if rng.random() < 0.1:
            continue                                                                          # record without sequence
        seq = synth.fuzz_contig(rng, int(rng.integers(1, max_len)), float(rng.choice([0, 0.001, 0.05])))
        width = int(rng.choice([1, 7, 60, 61, 80, 4095, 4096, 4097, 100000]))
        for i in range(0, len(seq), width):
            line = seq[i:i + width]
            if rng.random() < 0.02:
                line = line[:len(line) // 2] + b">" + line[len(line) // 2:]                   # '>' inside a line is a base (N)
            parts.append(line + (b"\r\n" if crlf else b"\n"))
            if rng.random() < 0.01:
                parts.append(b"\n")
    text = b"".join(parts)
    if rng.random() < 0.5 and text.endswith(b"\n"):
        text = text[:-1]
    return text


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
