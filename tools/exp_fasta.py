"""K0 measurement (GPU box): FASTA file -> packed contigs. Host reader (ribbit_b200/fasta.py, the reference's getline loop
restated) + rb_load_contigs versus rb_load_fasta (text to the device as it is, parsed there; from a file in the page cache
and from pinned memory).   usage: python tools/exp_fasta.py [Mbp]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribbit_b200 import fasta, scan, synth  # noqa: E402


def best(f, n=5):
    b = 1e9
    for _ in range(n):
        t = time.perf_counter(); f(); b = min(b, time.perf_counter() - t)
    return b


def main():
    mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 46.7
    seq = synth.contig_c2(int(mbp * 1e6), seed=21)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "x.fa")
        synth.write_fasta(path, [seq], names=["chr21"])
        sc = scan.Scanner(2, 100)
        text = np.fromfile(path, dtype=np.uint8)
        import torch
        pinned = torch.empty(len(text), dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[:] = text
        ptext = pinned.numpy()

        def host_arm():
            names, seqs = fasta.read_fasta(path)
            sc.load(seqs)

        def device_arm_file():
            sc.load_fasta(np.fromfile(path, dtype=np.uint8))

        t_host = best(host_arm, 3)
        t_dev_file = best(device_arm_file)
        t_dev_pinned = best(lambda: sc.load_fasta(ptext))
        names, lens = sc.load_fasta(ptext)
        assert names == ["chr21"] and lens.tolist() == [len(seq)]
        sc.scan_device()
        n1 = sc.counts()
        sc.load([seq]); sc.scan_device()
        assert sc.counts() == n1
        print(json.dumps({"mbp": mbp, "file_bytes": int(len(text)), "host_reader_plus_load_ms": t_host * 1e3,
                          "load_fasta_from_file_ms": t_dev_file * 1e3, "load_fasta_pinned_ms": t_dev_pinned * 1e3,
                          "pinned_gbps": len(seq) / t_dev_pinned / 1e9}))


if __name__ == "__main__":
    main()
