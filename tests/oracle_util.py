"""Test-side access to the parity oracle (oracle/). TEST INFRASTRUCTURE only.

* ``scan_events(seq, m_lo, m_hi)``: CP1 streams from the plain-C restatement
  (oracle/scan_oracle.c), as an int32 array of rows (stream, start, end, mlen, time); time = position at which the reference makes the call, -1 for the tail flush.
* ``ref_cp(fasta_path, args)``: runs oracle/_ref/ribbit_ref_cp (the unmodified
  reference sources with checkpoint logging) and returns per-contig CP1/CP2
  arrays plus the BED bytes. Only available where oracle/_ref was built.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "_build", "librb_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "ribbit_ref")
REF_CP_BIN = os.path.join(ORACLE_DIR, "_ref", "ribbit_ref_cp")


class _Ev(ctypes.Structure):
    _fields_ = [("stream", ctypes.c_int32), ("start", ctypes.c_int32), ("end", ctypes.c_int32),
                ("mlen", ctypes.c_int32), ("time", ctypes.c_int32)]


_lib = None


def build_port():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def port():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < max(
                os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("scan_oracle.c", "motif_oracle.c")):
            build_port()
        lib = ctypes.CDLL(PORT_SO)
        lib.rbo_scan.restype = ctypes.c_int64
        lib.rbo_scan.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                 ctypes.POINTER(ctypes.POINTER(_Ev))]
        lib.rbo_scan_count.restype = ctypes.c_int64
        lib.rbo_scan_count.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int64)]
        lib.rbo_pack.restype = None
        lib.rbo_pack.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.rbo_anchored_plane.restype = None
        lib.rbo_anchored_plane.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        lib.rbo_free.argtypes = [ctypes.c_void_p]
        lib.rbo_motif_row.restype = ctypes.c_int32
        lib.rbo_motif_row.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.POINTER(ctypes.c_int32)]
        _lib = lib
    return _lib


def scan_events(seq: bytes, m_lo: int = 2, m_hi: int = 100) -> np.ndarray:
    lib = port()
    p = ctypes.POINTER(_Ev)()
    n = lib.rbo_scan(seq, len(seq), m_lo, m_hi, ctypes.byref(p))
    if n < 0:
        raise MemoryError("rbo_scan failed")
    if n == 0:
        out = np.zeros((0, 5), dtype=np.int32)
    else:
        out = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_int32)), shape=(n, 5)).copy()
    lib.rbo_free(p)
    return out


def scan_count(seq: bytes, m_lo: int = 2, m_hi: int = 100):
    lib = port()
    c = (ctypes.c_int64 * 4)()
    lib.rbo_scan_count(seq, len(seq), m_lo, m_hi, c)
    return list(c)


def pack(seq: bytes):
    lib = port()
    nw = (len(seq) + 31) // 32
    hi = np.zeros(max(nw, 1), np.uint32); lo = np.zeros(max(nw, 1), np.uint32); nn = np.zeros(max(nw, 1), np.uint32)
    lib.rbo_pack(seq, len(seq), hi.ctypes.data, lo.ctypes.data, nn.ctypes.data)
    return hi[:nw], lo[:nw], nn[:nw]


def anchored_plane(seq: bytes, m_lo: int, m_hi: int, m: int, p0: int, p1: int) -> np.ndarray:
    lib = port()
    out = np.zeros(max(p1 - p0, 1), np.uint8)
    lib.rbo_anchored_plane(seq, len(seq), m_lo, m_hi, m, p0, p1, out.ctypes.data)
    return out[:p1 - p0]


def motif_row(seq: bytes, seed_start: int, seed_len: int, m: int):
    """(row, count) of the oracle's mostFrequentLongerMotif row search."""
    lib = port()
    best = ctypes.c_int32(0)
    row = lib.rbo_motif_row(seq, len(seq), seed_start, seed_len, m, ctypes.byref(best))
    return int(row), int(best.value)


def have_ref() -> bool:
    return os.access(REF_CP_BIN, os.X_OK) and os.access(REF_BIN, os.X_OK)


def ref_cp(fasta_path: str, args=(), stop_after_cp2=False, timeout=3600):
    """Run the instrumented reference. Returns (contigs, bed_bytes, returncode) where contigs is a list of
    dicts {L, cp1: (n,4) int32 rows (stream,start,end,mlen), cp2: (n,5) rows (list,start,end,mlen,rank),
    cp4: (n,4) rows (seed_start, seed_seq_len, mlen, row) of every mostFrequentLongerMotif call}."""
    with tempfile.TemporaryDirectory() as td:
        cp = os.path.join(td, "cp.bin")
        bed = os.path.join(td, "out.bed")
        env = dict(os.environ, RB_CP_OUT=cp)
        if stop_after_cp2:
            env["RB_CP_STOP_AFTER_CP2"] = "1"
        r = subprocess.run([REF_CP_BIN, "-i", fasta_path, "-o", bed, *map(str, args)], env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=timeout)
        raw = np.fromfile(cp, dtype=np.int32) if os.path.exists(cp) else np.zeros(0, np.int32)
        raw = raw[: (raw.size // 5) * 5].reshape(-1, 5)
        contigs = []
        starts = np.flatnonzero(raw[:, 0] == 0)
        for i, s in enumerate(starts):
            e = starts[i + 1] if i + 1 < len(starts) else len(raw)
            blk = raw[s + 1:e]
            cp1 = blk[(blk[:, 0] >= 1) & (blk[:, 0] <= 3)][:, :4]
            cp2 = blk[(blk[:, 0] >= 11) & (blk[:, 0] <= 13)].copy()
            cp2[:, 0] -= 10
            cp4 = blk[blk[:, 0] == 21][:, 1:]
            contigs.append({"L": int(raw[s, 2]), "cp1": cp1, "cp2": cp2, "cp4": cp4})
        bed_bytes = open(bed, "rb").read() if os.path.exists(bed) else b""
        return contigs, bed_bytes, r.returncode


def read_fasta(path):
    """FASTA reader with the reference's quirks (ribbit.cpp:269-280)."""
    names, seqs = [], []
    name, parts = "", []
    with open(path, "rb") as f:
        for line in f.read().split(b"\n")[:-1] if True else []:
            if line[:1] == b">":
                if parts and b"".join(parts) != b"":
                    names.append(name); seqs.append(b"".join(parts))
                sp = line.find(b" ")
                name = (line[1:sp] if sp >= 0 else line[1:]).decode()
                parts = []
            else:
                parts.append(line)
    names.append(name); seqs.append(b"".join(parts))
    return names, seqs
