"""The FASTA reader reproduces the reference's reading quirks (ribbit.cpp:269-280)."""
import os
import tempfile

from ribbit_b200 import fasta


def _read(data: bytes):
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "x.fa")
        open(p, "wb").write(data)
        return fasta.read_fasta(p)


def test_reader_quirks():
    names, seqs = _read(b">a desc more\nACGT\nAC\n>b\tx\nGG\r\n>c\n>d e\nTT\n")
    # '>c' has no sequence: it does not close into a record of its own, the name is simply replaced by 'd'
    assert names == ["a", "b\tx", "d"]
    assert seqs == [b"ACGTAC", b"GG\r", b"TT"]
    # no header at all: one unnamed record; empty file: one empty record (processSequence is called unconditionally)
    assert _read(b"ACGT\nAC") == ([""], [b"ACGTAC"])
    assert _read(b"") == ([""], [b""])
    # blank lines are appended verbatim (nothing to append); a header without a space keeps the whole line
    assert _read(b">only\n\nAC\n\nGT\n") == (["only"], [b"ACGT"])


def _golden():
    import json
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_fasta.json")))


def check_against_reference(records_fn):
    """records_fn(text) -> (names, lengths). Compared with what the UNMODIFIED reference read from the same bytes
    (tests/golden/golden_fasta.json, made by tests/golden/make_golden_fasta.py with oracle/_ref/ribbit_ref_cp): the length of
    every record it handed to processSequence and the names it printed (every record but the last). Where the reference
    died on the input its records are a prefix."""
    import hashlib
    from fasta_cases import QUIRKS, golden_random_texts
    g = _golden()
    n = 0
    for texts, gold in ((QUIRKS, g["quirks"]), (golden_random_texts(), g["random"])):
        assert len(texts) == len(gold)
        for text, want in zip(texts, gold):
            assert hashlib.md5(text).hexdigest() == want["md5"], "golden made for another input"
            names, lengths = records_fn(text)
            k = len(want["lengths"])
            if want["rc"] == 0:
                assert list(lengths) == want["lengths"], text[:80]
            else:
                assert list(lengths)[:k] == want["lengths"], text[:80]
            kn = len(want["names_but_last"])
            assert list(names)[:kn] == want["names_but_last"], text[:80]
            if want["rc"] == 0:
                assert len(names) == kn + 1
            n += k
    return n


def test_reader_equals_the_reference_reader():
    def fn(text):
        names, seqs = _read(text)
        return names, [len(s) for s in seqs]
    assert check_against_reference(fn) > 100
