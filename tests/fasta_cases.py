"""FASTA texts for the K0 tests (CPU emulator and GPU): the reader quirks of ribbit.cpp:269-280 and random files."""
import numpy as np

from ribbit_b200 import synth

QUIRKS = (b">a desc more\nACGT\nAC\n>b\tx\nGG\r\n>c\n>d e\nTT\n", b"ACGT\nAC", b"", b"\n", b"\n\n>x\n", b">only\n\nAC\n\nGT\n",
          b">h1\n>h2\n>h3", b"AC>GT\n>n\nA>C\n>", b">\nACGT\n> spaced name\nGG\n", b"ACGT\n>late header\nTTTT\nGG",
          b">a\nACGT\n>b\n\n>c\nGG\n>d\n")


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


def golden_random_texts():
    """The 40 seeded random files behind tests/golden/golden_fasta.json (tests/golden/make_golden_fasta.py)."""
    rng = np.random.default_rng(2031)
    return [random_fasta(rng, int(rng.integers(1, 12)), int(rng.choice([50, 600, 5000])), crlf=(it % 7 == 3)) for it in range(40)]
