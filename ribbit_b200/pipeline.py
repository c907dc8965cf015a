"""Throughput pipeline over the C ABI: several scan contexts (rb_ctx) on one GPU, one host thread each, so that the
H2D copy, the kernels and the D2H copy of consecutive batches (contigs) overlap. Each context has its own CUDA stream
and pinned result buffers; ctypes releases the GIL during the library calls. Results of a batch stay valid until the
same context is used again (every `depth`-th submission), so consume or copy them before that.

    pipe = ScanPipeline(2, 100, device=0, depth=3)
    futures = [pipe.submit_flat(buf, [L]) for buf in batches]     # buf: pinned uint8 numpy array
    for f in futures: streams = f.result()
"""
import threading
import time
from concurrent.futures import ThreadPoolExecutor

from . import scan


def group_contigs(indices, lengths, target_bases):
    """Consecutive runs of `indices` (contigs adjacent in the host buffer) with about `target_bases` each: the batches to
    submit. A contig longer than the target is a batch of its own."""
    out, cur, size = [], [], 0
    for c in indices:
        if cur and (c != cur[-1] + 1 or size + lengths[c] > target_bases):
            out.append(cur); cur, size = [], 0
        cur.append(c); size += lengths[c]
    if cur:
        out.append(cur)
    return out


class ScanPipeline:
    def __init__(self, min_mlen=2, max_mlen=100, device=0, depth=3, copy=False, compact=False, compute_slots=1, trace=False):
        self.depth = depth
        self.copy = copy
        self.compact = compact  # fetch 8-byte records (rb_fetch_compact): half the D2H traffic
        self.scanners = [scan.Scanner(min_mlen, max_mlen, device=device) for _ in range(depth)]
        self.pools = [ThreadPoolExecutor(max_workers=1) for _ in range(depth)]  # one host thread per context
        self.n = 0
        # The kernels of `compute_slots` batches run at a time; the other contexts copy meanwhile. Without the gate the
        # contexts drift into lockstep (all copy in, all scan, all copy out) and the copies stop hiding behind the kernels.
        self.compute = threading.Semaphore(compute_slots)
        self.h2d = threading.Lock()  # one copy in at a time: the first batch's kernels start after ITS copy, not after all
        self.trace = [] if trace else None  # (context, bases, t_load0, t_load1, t_scan0, t_scan1, t_fetch1) per batch
        self._lock = threading.Lock()

    def _run(self, k, buf, lengths, word_range, offsets):
        sc = self.scanners[k]
        t0 = time.perf_counter()
        with self.h2d:
            sc.load_flat(buf, lengths, offsets)
        if word_range is not None:
            sc.set_word_range(*word_range)
        t1 = time.perf_counter()
        with self.compute:
            t2 = time.perf_counter()
            sc.scan_device()
            t3 = time.perf_counter()
        res = sc.fetch_compact(self.copy) if self.compact else sc.fetch(self.copy)
        if self.trace is not None:
            with self._lock:
                self.trace.append((k, sum(lengths), t0, t1, t2, t3, time.perf_counter()))
        return res

    def submit_flat(self, buf, lengths, word_range=None, offsets=None):
        """buf: uint8 numpy array holding the contigs back to back, or at `offsets` (pinned for full-speed copies).
        word_range: scan only the words [first, last) of a single-contig batch (rb_set_word_range: one contig over
        several GPUs). Batches of a few hundred Mbp keep the kernels at their large-batch rate; see group_contigs."""
        k = self.n % self.depth
        self.n += 1
        return self.pools[k].submit(self._run, k, buf, lengths, word_range, offsets)

    def close(self):
        for p in self.pools:
            p.shutdown(wait=True)
        for s in self.scanners:
            s.close()
