"""Throughput pipeline over the C ABI: several scan contexts (rb_ctx) on one GPU, one host thread each, so that the
H2D copy, the kernels and the D2H copy of consecutive batches (contigs) overlap. Each context has its own CUDA stream
and pinned result buffers; ctypes releases the GIL during the library calls. Results of a batch stay valid until the
same context is used again (every `depth`-th submission), so consume or copy them before that.

    pipe = ScanPipeline(2, 100, device=0, depth=3)
    futures = [pipe.submit_flat(buf, [L]) for buf in batches]     # buf: pinned uint8 numpy array
    for f in futures: streams = f.result()
"""
from concurrent.futures import ThreadPoolExecutor

from . import scan


class ScanPipeline:
    def __init__(self, min_mlen=2, max_mlen=100, device=0, depth=3, copy=False, compact=False):
        self.depth = depth
        self.copy = copy
        self.compact = compact  # fetch 8-byte records (rb_fetch_compact): half the D2H traffic
        self.scanners = [scan.Scanner(min_mlen, max_mlen, device=device) for _ in range(depth)]
        self.pools = [ThreadPoolExecutor(max_workers=1) for _ in range(depth)]  # one host thread per context
        self.n = 0

    def _run(self, k, buf, lengths, word_range):
        sc = self.scanners[k]
        sc.load_flat(buf, lengths)
        if word_range is not None:
            sc.set_word_range(*word_range)
        return sc.scan_compact(copy=self.copy) if self.compact else sc.scan(copy=self.copy)

    def submit_flat(self, buf, lengths, word_range=None):
        """buf: uint8 numpy array holding the contigs back to back (pinned for full-speed copies). word_range: scan only
        the words [first, last) of a single-contig batch (rb_set_word_range: one contig over several GPUs)."""
        k = self.n % self.depth
        self.n += 1
        return self.pools[k].submit(self._run, k, buf, lengths, word_range)

    def close(self):
        for p in self.pools:
            p.shutdown(wait=True)
        for s in self.scanners:
            s.close()
