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
