"""ctypes access to the host seed-list merges (ribbit_b200/host/seed_merge.cpp, built by ribbit_b200.build.build_merge)."""
import ctypes

import numpy as np

import oracle_util as ou
import stream_model as sm

_lib = None


def lib():
    global _lib
    if _lib is None:
        from ribbit_b200 import build
        _lib = ctypes.CDLL(build.build_merge())
        _lib.rbm_run.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                 ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]
        _lib.rbm_free.argtypes = [ctypes.c_void_p]
    return _lib


def merged_lists(seq: bytes, streams, m_lo: int, m_hi: int):
    """streams: {1|2|3: rows (start, end, mlen, flags[, time])} as the scan library reports them. Returns the three seed
    lists as (n,5) rows (list 1..3, start, end, mlen, rank) = checkpoint CP2."""
    L = lib()
    hi, lo, nn = ou.pack(seq)
    hi = np.ascontiguousarray(hi); lo = np.ascontiguousarray(lo); nn = np.ascontiguousarray(nn)
    if len(hi) == 0:
        hi = lo = nn = np.zeros(1, np.uint32)
    cands = [np.ascontiguousarray(np.asarray(streams[s])[:, :4], dtype=np.int32).reshape(-1, 4) for s in (1, 2, 3)]
    out = (ctypes.c_void_p * 3)()
    n = (ctypes.c_int64 * 3)()
    rc = L.rbm_run(cands[0].ctypes.data, len(cands[0]), cands[1].ctypes.data, len(cands[1]), cands[2].ctypes.data, len(cands[2]),
                   hi.ctypes.data, lo.ctypes.data, nn.ctypes.data, len(seq), m_lo, m_hi, out, n)
    assert rc == 0
    rows = []
    for k in range(3):
        a = np.ctypeslib.as_array(ctypes.cast(out[k], ctypes.POINTER(ctypes.c_int32)), shape=(max(n[k], 1), 4))[:n[k]].copy()
        L.rbm_free(out[k])
        rows.append(np.concatenate([np.full((len(a), 1), k + 1, np.int32), a], axis=1))
    return np.concatenate(rows) if rows else np.zeros((0, 5), np.int32)


def merged_from_oracle(seq: bytes, m_lo: int, m_hi: int):
    return merged_lists(seq, sm.expected_streams(seq, ou.scan_events(seq, m_lo, m_hi)), m_lo, m_hi)
