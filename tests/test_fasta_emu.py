"""K0 on the CPU: the slice arithmetic and the three-pass tile logic of the FASTA ingest kernels (fasta_core.h, mirrored by
tests/emu) against a byte-serial statement of the reference reader (ribbit.cpp:269-280): a line is a header iff its first
byte is '>', every other byte except '\\n' is sequence."""
import numpy as np

import emu_util
from fasta_cases import QUIRKS, random_fasta


def serial_model(text: bytes):
    bases, hpos, hseq = bytearray(), [], []
    line_start, in_header = True, False
    for i, ch in enumerate(text):
        if ch == 0x0A:
            line_start, in_header = True, False
            continue
        if line_start and ch == 0x3E:
            in_header = True
            hpos.append(i); hseq.append(len(bases))
        elif not in_header:
            bases.append(ch)
        line_start = False
    return bytes(bases), np.array(hpos, np.int64), np.array(hseq, np.int64)


def _check(text):
    got = emu_util.emu_fasta(text)
    exp = serial_model(text)
    assert got[0] == exp[0]
    assert np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])


def test_quirk_cases():
    for text in QUIRKS:
        _check(text)


def test_random_files_and_tile_boundaries():
    rng = np.random.default_rng(2028)
    for it in range(25):
        _check(random_fasta(rng, int(rng.integers(1, 20)), int(rng.choice([50, 3000, 20000])), crlf=(it % 7 == 3)))
    # control bytes exactly at slice (16 B) and tile (4096 B) boundaries
    for pad in (4094, 4095, 4096, 4097, 8191, 8192, 15, 16, 17):
        _check(b"A" * pad + b"\n>h\nCC\n" + b"G" * 5000 + b"\n>" + b"x" * 9000 + b"\nTT")
        _check(b">" + b"n" * pad + b"\n" + b"ACGT" * 3000)
