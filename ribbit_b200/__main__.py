"""Command line access to the GPU scan (the candidate streams, checkpoint CP1):

    python -m ribbit_b200 scan -i in.fa [-o cands.tsv] [-m 2] [-M 100] [--device 0] [--all]

Writes one row per candidate: contig, stream (P|S|A), start, end, mlen, flags, time — by default only the candidates
that reach the consumer's length cutoff (flags 0 / NOCOMMIT); --all adds the DROPPED and PSEUDO bookkeeping records.
For BED output use the drop-in program ribbit_b200/bin/ribbit_gpu."""
import argparse
import os
import sys

import numpy as np

from . import fasta, scan


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m ribbit_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    sp = sub.add_parser("scan", help="scan a FASTA file on the GPU and write the candidate streams")
    sp.add_argument("-i", "--input-file", required=True)
    sp.add_argument("-o", "--output-file", default="-")
    sp.add_argument("-m", "--min-motif-length", type=int, default=2)
    sp.add_argument("-M", "--max-motif-length", type=int, default=100)
    sp.add_argument("--device", type=int, default=0)
    sp.add_argument("--all", action="store_true")
    sp.add_argument("--batch-bases", type=int, default=400_000_000, help="bases per GPU batch")
    args = ap.parse_args(argv)

    out = sys.stdout if args.output_file == "-" else open(args.output_file, "w")
    sc = scan.Scanner(args.min_motif_length, args.max_motif_length, device=args.device)
    whole = os.path.getsize(args.input_file) <= args.batch_bases
    if whole:   # one batch: the file goes to the device as it is and is parsed there (rb_load_fasta)
        names, lengths = sc.load_fasta(np.fromfile(args.input_file, dtype=np.uint8))
        seqs = [None] * len(names)
    else:       # several batches: records are cut on the host (same reader semantics, ribbit.cpp:269-280)
        names, seqs = fasta.read_fasta(args.input_file)
        lengths = [len(s) for s in seqs]
    i = 0
    while i < len(seqs):
        j, tot = i, 0
        while j < len(seqs) and (j == i or whole or tot + lengths[j] <= args.batch_bases):
            tot += lengths[j]; j += 1
        if not whole:
            sc.load(seqs[i:j])
        res = sc.scan(copy=False)
        for c in range(i, j):
            for s, tag in enumerate("PSA"):
                a, off = res[s]
                r = a[off[c - i]:off[c - i + 1]]
                if not args.all:
                    r = r[(r["flags"] & (scan.REC_DROPPED | scan.REC_PSEUDO)) == 0]
                for st, en, m, fl, t in zip(r["start"].tolist(), r["end"].tolist(), r["mlen"].tolist(), r["flags"].tolist(), r["time"].tolist()):
                    out.write("%s\t%s\t%d\t%d\t%d\t%d\t%d\n" % (names[c], tag, st, en, m, fl, t))
        i = j
    sc.close()
    if out is not sys.stdout:
        out.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
