"""ctypes access to the CPU warp emulator (tests/emu/emu_scan.cpp): the kernels' lane logic, run lane by lane on the
CPU. TEST INFRASTRUCTURE — lets the CPU suite check the kernel logic against the oracle where no GPU exists."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        from ribbit_b200 import build
        _lib = ctypes.CDLL(build.build_emulator())
        _lib.emu_scan.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
                                  ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        _lib.emu_free.argtypes = [ctypes.c_void_p]
        _lib.emu_fasta.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.emu_motif_rows.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    return _lib


def emu_streams(seq: bytes, m_lo: int, m_hi: int, chunk_words: int = 1 << 30, warm0: int = 4):
    """Returns ({stream: (n,5) rows (start, end, mlen, flags, time)}, warm-up restarts)."""
    L = lib()
    out = (ctypes.c_void_p * 3)()
    n = (ctypes.c_int64 * 3)()
    rs = ctypes.c_int64()
    sk = ctypes.c_int64()
    rp = ctypes.c_int64()
    L.emu_scan(seq, len(seq), m_lo, m_hi, chunk_words, warm0, out, n, ctypes.byref(rs), ctypes.byref(sk), ctypes.byref(rp))
    emu_streams.last_skips = sk.value
    emu_streams.last_replays = rp.value
    res = {}
    for s in range(3):
        if n[s]:
            a = np.ctypeslib.as_array(ctypes.cast(out[s], ctypes.POINTER(ctypes.c_int32)), shape=(n[s], 4)).copy()
            rows = np.stack([a[:, 0], a[:, 1], a[:, 2] & 0xFFFF, (a[:, 2] >> 16) & 0xFFF, a[:, 3]], axis=1).astype(np.int64)
        else:
            rows = np.zeros((0, 5), np.int64)
        L.emu_free(out[s])
        res[s + 1] = rows
    return res, rs.value


def emu_motif_rows(seq: bytes, seeds):
    """K7 lane logic on the CPU. seeds: (n,3) rows (seed_start, seed_end, mlen) -> (n,2) rows (row, count)."""
    sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.int32).reshape(-1, 3))
    out = np.zeros((len(sd), 2), np.int32)
    lib().emu_motif_rows(seq, len(seq), sd.ctypes.data, len(sd), out.ctypes.data)
    return out


def emu_fasta(text: bytes):
    """K0 on the CPU: (sequence bytes, header offsets in the text, sequence bytes in front of each header)."""
    buf = np.frombuffer(text, dtype=np.uint8) if len(text) else np.zeros(1, np.uint8)
    n = len(text)
    bases = np.zeros(max(n, 1), np.uint8)
    hpos = np.zeros(max(n, 1), np.int64)
    hseq = np.zeros(max(n, 1), np.int64)
    tot = np.zeros(2, np.int64)
    lib().emu_fasta(buf.ctypes.data, n, bases.ctypes.data, hpos.ctypes.data, hseq.ctypes.data, tot.ctypes.data)
    return bases[:tot[0]].tobytes(), hpos[:tot[1]].copy(), hseq[:tot[1]].copy()
