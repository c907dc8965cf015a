"""ribbit-b200: the tandem-repeat seed scan of SowpatiLab/ribbit on NVIDIA B200 (sm_100a).

    ribbit_b200.scan       ctypes binding of the C ABI (include/ribbit_scan.h, lib/libribbit_scan.so)
    ribbit_b200.pipeline   several scan contexts on one GPU: overlapped H2D / kernels / D2H
    ribbit_b200.shard      contigs over ranks (one process per GPU)
    ribbit_b200.fasta      FASTA reader with the reference's quirks
    ribbit_b200.synth      seeded synthetic inputs of the BASELINE.json shapes
    ribbit_b200.build      in-tree build of the CUDA library and of the test infrastructure
    python -m ribbit_b200 scan -i in.fa -o candidates.tsv

There is no CPU implementation of the scan in this package: without the CUDA library and a GPU it fails loudly.
"""
__version__ = "0.1.0"
