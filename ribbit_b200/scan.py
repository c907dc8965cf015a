"""ctypes binding of the C ABI in include/ribbit_scan.h (libribbit_scan.so) — the reference-facing call a Python
user makes. Thin: no computation happens here, and there is no CPU fallback; loading fails loudly when the CUDA
library is missing or no GPU is usable.

The reference's own interface for this path is the group of functions processSequence calls
(/root/reference/fasta_utils.cpp:117-170); `Scanner.scan` returns what those functions hand to the host merges:
the three ordered candidate streams (perfect / substitution / anchored), see include/ribbit_scan.h.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RIBBIT_SCAN_LIB", os.path.join(_HERE, "lib", "libribbit_scan.so"))  # override: experiments only

REC_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("mlen", "<u2"), ("flags", "<u2"), ("time", "<i4")])
REC8_DTYPE = np.dtype([("start", "<i4"), ("len", "<u2"), ("mf", "<u2")])
LONG_DTYPE = np.dtype([("index", "<i8"), ("end", "<i8")])
REC_DROPPED, REC_PSEUDO, REC_NOCOMMIT = 1, 2, 4
STREAM_NAMES = ("perfect", "subst", "anchored")

EXPORTS = ("rb_abi_version", "rb_create", "rb_destroy", "rb_last_error", "rb_load_contigs", "rb_load_contigs_device",
           "rb_scan_device", "rb_fetch", "rb_scan", "rb_counts", "rb_get_timing", "rb_filter_seeds", "rb_get_planes",
           "rb_measure_int_peak", "rb_get_anchor_planes", "rb_fetch_compact", "rb_motif_rows", "rb_load_fasta", "rb_fasta_records", "rb_set_word_range", "rb_get_elided_max", "rb_measure_int_peak_modes", "rb_debug_item_clocks")


FASTA_REC_DTYPE = np.dtype([("name_off", "<i8"), ("name_len", "<i4"), ("length", "<i4")])


class RbParams(ctypes.Structure):
    _fields_ = [("min_mlen", ctypes.c_int32), ("max_mlen", ctypes.c_int32), ("chunk_words", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class RbStreams(ctypes.Structure):
    _fields_ = [("n_contigs", ctypes.c_int32), ("reserved", ctypes.c_int32), ("rec", ctypes.c_void_p * 3),
                ("contig_off", ctypes.c_void_p * 3), ("n", ctypes.c_int64 * 3)]


class RbStreams8(ctypes.Structure):
    _fields_ = [("n_contigs", ctypes.c_int32), ("reserved", ctypes.c_int32), ("rec", ctypes.c_void_p * 3),
                ("contig_off", ctypes.c_void_p * 3), ("n", ctypes.c_int64 * 3), ("long_end", ctypes.c_void_p * 3),
                ("n_long", ctypes.c_int64 * 3)]


class RbTiming(ctypes.Structure):
    _fields_ = [("pack_ms", ctypes.c_float), ("scan_ms", ctypes.c_float), ("merge_ms", ctypes.c_float),
                ("total_ms", ctypes.c_float), ("launches", ctypes.c_int32), ("restarts", ctypes.c_int32),
                ("retries", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class RbSeed(ctypes.Structure):
    _fields_ = [("contig", ctypes.c_int32), ("start", ctypes.c_int32), ("end", ctypes.c_int32), ("mlen", ctypes.c_int32)]


class RbSeedInfo(ctypes.Structure):
    _fields_ = [("seq_len", ctypes.c_int32), ("longest_run", ctypes.c_int32)]


class RibbitScanError(RuntimeError):
    pass


_lib = None


def load_library(path=LIB_PATH):
    """Loads libribbit_scan.so. Raises if it was not built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RibbitScanError("%s not found: the CUDA library is not built and there is no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    lib.rb_abi_version.restype = ctypes.c_int
    lib.rb_create.restype = ctypes.c_void_p
    lib.rb_create.argtypes = [ctypes.c_int, ctypes.POINTER(RbParams)]
    lib.rb_destroy.restype = None
    lib.rb_destroy.argtypes = [ctypes.c_void_p]
    lib.rb_last_error.restype = ctypes.c_char_p
    lib.rb_last_error.argtypes = [ctypes.c_void_p]
    lib.rb_load_contigs.restype = ctypes.c_int
    lib.rb_load_contigs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    lib.rb_load_contigs_device.restype = ctypes.c_int
    lib.rb_load_contigs_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    lib.rb_scan_device.restype = ctypes.c_int
    lib.rb_scan_device.argtypes = [ctypes.c_void_p]
    lib.rb_fetch.restype = ctypes.c_int
    lib.rb_fetch.argtypes = [ctypes.c_void_p, ctypes.POINTER(RbStreams)]
    lib.rb_fetch_compact.restype = ctypes.c_int
    lib.rb_fetch_compact.argtypes = [ctypes.c_void_p, ctypes.POINTER(RbStreams8)]
    lib.rb_scan.restype = ctypes.c_int
    lib.rb_scan.argtypes = [ctypes.c_void_p, ctypes.POINTER(RbStreams)]
    lib.rb_counts.restype = ctypes.c_int
    lib.rb_counts.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
    lib.rb_get_timing.restype = ctypes.c_int
    lib.rb_get_timing.argtypes = [ctypes.c_void_p, ctypes.POINTER(RbTiming)]
    lib.rb_filter_seeds.restype = ctypes.c_int
    lib.rb_filter_seeds.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    lib.rb_load_fasta.restype = ctypes.c_int
    lib.rb_load_fasta.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32)]
    lib.rb_fasta_records.restype = ctypes.c_int
    lib.rb_fasta_records.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    lib.rb_set_word_range.restype = ctypes.c_int
    lib.rb_set_word_range.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]
    lib.rb_get_elided_max.restype = ctypes.c_int
    lib.rb_get_elided_max.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
    lib.rb_motif_rows.restype = ctypes.c_int
    lib.rb_motif_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    lib.rb_measure_int_peak.restype = ctypes.c_int
    lib.rb_measure_int_peak.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    lib.rb_measure_int_peak_modes.restype = ctypes.c_int
    lib.rb_measure_int_peak_modes.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    lib.rb_debug_item_clocks.restype = ctypes.c_int
    lib.rb_debug_item_clocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    lib.rb_get_anchor_planes.restype = ctypes.c_int
    lib.rb_get_anchor_planes.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.rb_get_planes.restype = ctypes.c_int
    lib.rb_get_planes.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    _lib = lib
    return lib


class Scanner:
    """One scan context on one GPU (rb_ctx). Motif range = ribbit's -m / -M."""

    def __init__(self, min_mlen=2, max_mlen=100, device=0, chunk_words=0, debug=0):
        self.lib = load_library()
        p = RbParams(min_mlen, max_mlen, chunk_words, debug)
        self.ctx = self.lib.rb_create(device, ctypes.byref(p))
        if not self.ctx:
            raise RibbitScanError(self.lib.rb_last_error(None).decode())
        self.n_contigs = 0
        self.lengths = None

    def close(self):
        if self.ctx:
            self.lib.rb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RibbitScanError("rb error %d: %s" % (rc, self.lib.rb_last_error(self.ctx).decode()))

    @staticmethod
    def _tables(lengths):
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        offsets = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths) > 1:
            offsets[1:] = np.cumsum(lengths[:-1].astype(np.int64))
        return offsets, lengths

    def load(self, contigs):
        """contigs: list of bytes (one per contig), host memory. H2D copy + geometry."""
        lengths = [len(s) for s in contigs]
        buf = np.frombuffer(b"".join(contigs) + b"\0", dtype=np.uint8)
        return self.load_flat(buf, lengths)

    def load_flat(self, buf, lengths, offsets=None):
        """buf: uint8 numpy array (host) holding the contigs back to back (or at `offsets`)."""
        off, lengths = self._tables(lengths)
        if offsets is not None:
            off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._keep = (buf, off, lengths)
        self._check(self.lib.rb_load_contigs(self.ctx, buf.ctypes.data, off.ctypes.data, lengths.ctypes.data, len(lengths)))
        self.n_contigs = len(lengths)
        self.lengths = lengths

    def load_device(self, dev_ptr, lengths, offsets=None, keepalive=None):
        """dev_ptr: integer device address of the ASCII bytes already resident in HBM (e.g. tensor.data_ptr())."""
        off, lengths = self._tables(lengths)
        if offsets is not None:
            off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._keep = (keepalive, off, lengths)
        self._check(self.lib.rb_load_contigs_device(self.ctx, ctypes.c_void_p(dev_ptr), off.ctypes.data, lengths.ctypes.data,
                                                    len(lengths)))
        self.n_contigs = len(lengths)
        self.lengths = lengths

    def load_fasta(self, text):
        """text: FASTA file content (bytes or uint8 array). Parsed on the device (rb_load_fasta, the reference's reader
        quirks of ribbit.cpp:269-280); the records become the loaded contigs. Returns (names, lengths)."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, dtype=np.uint8)
        n = ctypes.c_int32(0)
        self._check(self.lib.rb_load_fasta(self.ctx, buf.ctypes.data if buf.size else None, buf.size, ctypes.byref(n)))
        rec = np.zeros(n.value, dtype=FASTA_REC_DTYPE)
        self._check(self.lib.rb_fasta_records(self.ctx, rec.ctypes.data, n.value))
        names = ["" if o < 0 else bytes(buf[o:o + l]).decode(errors="replace") for o, l in zip(rec["name_off"], rec["name_len"])]
        self.n_contigs = n.value
        self.lengths = rec["length"].astype(np.int32)
        self._keep = None
        return names, self.lengths.copy()

    def set_word_range(self, word_first, word_last=-1):
        """Single-contig batch: scan only the 32-base words [word_first, word_last) (rb_set_word_range)."""
        self._check(self.lib.rb_set_word_range(self.ctx, word_first, word_last))

    def elided_max(self):
        """Largest end of the candidates elided in the scanned part, per stream (subst, anchored); -1 = none."""
        out = (ctypes.c_int64 * 2)()
        self._check(self.lib.rb_get_elided_max(self.ctx, out))
        return [int(out[0]), int(out[1])]

    def scan_device(self):
        self._check(self.lib.rb_scan_device(self.ctx))

    def counts(self):
        n = (ctypes.c_int64 * 3)()
        self._check(self.lib.rb_counts(self.ctx, n))
        return list(n)

    def fetch(self, copy=True):
        out = RbStreams()
        self._check(self.lib.rb_fetch(self.ctx, ctypes.byref(out)))
        return self._wrap(out, copy)

    def fetch_compact(self, copy=True):
        """The streams as 8-byte records (rb_rec8): {stream: (records, contig offsets, long-candidate list)}."""
        out = RbStreams8()
        self._check(self.lib.rb_fetch_compact(self.ctx, ctypes.byref(out)))
        res = {}
        n = out.n_contigs
        for s in range(3):
            cnt = out.n[s]
            a = (np.ctypeslib.as_array(ctypes.cast(out.rec[s], ctypes.POINTER(ctypes.c_uint8)), shape=(cnt * 8,)).view(REC8_DTYPE)
                 if cnt else np.zeros(0, dtype=REC8_DTYPE))
            off = np.ctypeslib.as_array(ctypes.cast(out.contig_off[s], ctypes.POINTER(ctypes.c_int64)), shape=(n + 1,))
            nl = out.n_long[s]
            lg = (np.ctypeslib.as_array(ctypes.cast(out.long_end[s], ctypes.POINTER(ctypes.c_uint8)), shape=(nl * 16,)).view(LONG_DTYPE)
                  if nl else np.zeros(0, dtype=LONG_DTYPE))
            res[s] = (a.copy() if copy else a, off.copy() if copy else off, lg.copy())
        return res

    def scan_compact(self, copy=True):
        self.scan_device()
        return self.fetch_compact(copy)

    def scan(self, copy=True):
        out = RbStreams()
        self._check(self.lib.rb_scan(self.ctx, ctypes.byref(out)))
        return self._wrap(out, copy)

    def _wrap(self, out, copy):
        res = {}
        n = out.n_contigs
        for s in range(3):
            cnt = out.n[s]
            if cnt:
                a = np.ctypeslib.as_array(ctypes.cast(out.rec[s], ctypes.POINTER(ctypes.c_uint8)), shape=(cnt * 16,)).view(REC_DTYPE)
            else:
                a = np.zeros(0, dtype=REC_DTYPE)
            off = np.ctypeslib.as_array(ctypes.cast(out.contig_off[s], ctypes.POINTER(ctypes.c_int64)), shape=(n + 1,))
            res[s] = (a.copy() if copy else a, off.copy() if copy else off)
        return res

    def timing(self):
        t = RbTiming()
        self._check(self.lib.rb_get_timing(self.ctx, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in RbTiming._fields_ if k != "reserved"}

    def int_peak(self):
        """Measured LOP3+SHF lane-operations per second of this GPU (roofline denominator)."""
        v = ctypes.c_double()
        self._check(self.lib.rb_measure_int_peak(self.ctx, ctypes.byref(v)))
        return v.value

    def int_peak_modes(self):
        """Measured lane-operations per second for {SHF + LOP3, LOP3 only, SHF only, LOP3 + IMAD}."""
        v = (ctypes.c_double * 4)()
        self._check(self.lib.rb_measure_int_peak_modes(self.ctx, v))
        return {"shf_lop3": v[0], "lop3": v[1], "shf": v[2], "lop3_imad": v[3]}

    def planes(self, contig):
        nw = (int(self.lengths[contig]) + 31) // 32
        hi = np.zeros(max(nw, 1), np.uint32); lo = np.zeros(max(nw, 1), np.uint32); nn = np.zeros(max(nw, 1), np.uint32)
        self._check(self.lib.rb_get_planes(self.ctx, contig, hi.ctypes.data, lo.ctypes.data, nn.ctypes.data))
        return hi[:nw], lo[:nw], nn[:nw]

    def anchor_planes(self, contig, shift_lo, shift_hi):
        """(shift_hi-shift_lo+1, nw) uint32: anchor planes A_s of one contig, computed on the device."""
        nw = (int(self.lengths[contig]) + 31) // 32
        out = np.zeros((shift_hi - shift_lo + 1, max(nw, 1)), dtype=np.uint32)
        if nw:
            out = np.zeros((shift_hi - shift_lo + 1, nw), dtype=np.uint32)
            self._check(self.lib.rb_get_anchor_planes(self.ctx, contig, shift_lo, shift_hi, out.ctypes.data))
        return out[:, :nw]

    def filter_seeds(self, seeds):
        """seeds: (n,4) int32 rows (contig, start, end, mlen) -> (n,2) int32 rows (seq_len, longest_run)."""
        seeds = np.ascontiguousarray(seeds, dtype=np.int32).reshape(-1, 4)
        out = np.zeros((len(seeds), 2), dtype=np.int32)
        self._check(self.lib.rb_filter_seeds(self.ctx, seeds.ctypes.data, len(seeds), out.ctypes.data))
        return out

    def motif_rows(self, seeds):
        """K7, the row search of mostFrequentLongerMotif (parse_seed.cpp:153-256). seeds: (n,4) int32 rows
        (contig, seed_start, seed_start + seed_sequence_length, mlen) -> (n,2) int32 rows (row, count)."""
        seeds = np.ascontiguousarray(seeds, dtype=np.int32).reshape(-1, 4)
        out = np.zeros((len(seeds), 2), dtype=np.int32)
        self._check(self.lib.rb_motif_rows(self.ctx, seeds.ctypes.data, len(seeds), out.ctypes.data))
        return out


def expand_compact(rec8, long_end):
    """rb_rec8 array (+ its long-candidate list) -> rows (start, end, mlen, flags), the PSEUDO convention of rb_rec."""
    start = rec8["start"].astype(np.int64)
    end = start + rec8["len"].astype(np.int64)
    if len(long_end):
        end[long_end["index"]] = long_end["end"]
    mlen = (rec8["mf"] & 0xFFF).astype(np.int64)
    flags = (rec8["mf"] >> 12).astype(np.int64)
    ps = (flags & REC_PSEUDO) != 0
    end[ps] = start[ps]
    start[ps] = -1
    return np.stack([start, end, mlen, flags], axis=1)


def contig_streams(res, contig):
    """Rows (start, end, mlen, flags, time) of one contig per stream, from Scanner.scan() output."""
    out = {}
    for s in range(3):
        a, off = res[s]
        r = a[off[contig]:off[contig + 1]]
        out[s + 1] = np.stack([r["start"], r["end"], r["mlen"], r["flags"], r["time"]], axis=1).astype(np.int64) if len(r) else np.zeros((0, 5), np.int64)
    return out
