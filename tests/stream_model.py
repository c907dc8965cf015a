"""Expected contents of the library's candidate streams, derived from the oracle's raw CP1 events.

The C ABI (include/ribbit_scan.h) documents the stream encoding; this module restates it for the tests:

* every raw candidate (start, end, mlen) the reference passes to addSeedToSeedPositions* either
  - is kept (end-start >= the consumer's cutoff, or any perfect candidate)  -> record, flags 0
  - is below the cutoff and was emitted from a *slow* word (scan_core.h word_is_fast: a word where, or right
    after a word where, not all 8-base windows are valid - near an N / the contig start - and the last word of
    the contig) or from the tail flush                                       -> record, flags DROPPED
  - is below the cutoff and was emitted from a fast word                     -> elided
* anchored tail-flush calls whose returned cursors the reference discards (parse_anchored_shiftxor.cpp:688-719)
  carry NOCOMMIT when kept and are elided when below the cutoff
* in front of every non-empty slow bucket of the substitution / anchored stream sits a PSEUDO record whose `end`
  is the largest end among the candidates elided before it (-1 if none).
"""
import numpy as np

DROPPED, PSEUDO, NOCOMMIT = 1, 2, 4


def cut_subst(m):
    return m // 3 if m > 30 else 10


def cut_anch(m):
    c = m if m > 6 else 10
    if m >= 10:
        c = int(0.9 * m)
    return c


def fast_words(seq: bytes):
    """bool per word, as scan_core.h word_is_fast: every window ending in words w-1 and w is evaluated
    (v[p] = p>=7 and no N in [p-7,p]; positions 0..6 of the contig count as evaluated) and w is not the last word of the
    contig."""
    L = len(seq)
    a = np.frombuffer(seq, dtype=np.uint8)
    isn = ~np.isin(a, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    idx = np.arange(L)
    lastn = np.maximum.accumulate(np.where(isn, idx, -1)) if L else np.zeros(0, dtype=np.int64)
    v = (idx - lastn) >= 8
    nw = (L + 31) // 32
    vv = np.zeros(nw * 32, dtype=bool)
    vv[:L] = v
    allv = vv.reshape(nw, 32).all(axis=1) if nw else np.zeros(0, dtype=bool)
    # the contig start counts as evaluated (scan_core.h v_eff): the windows that would end at positions 0..6 do not exist
    if nw:
        allv[0] = bool(vv[7:32].all()) if L >= 32 else False
    fast = allv.copy()
    if nw:
        fast[1:] &= allv[:-1]
        fast[nw - 1] = False
    return fast


def expected_streams(seq: bytes, events: np.ndarray, with_elided: bool = False):
    """events: oracle rows (stream 1..3, start, end, mlen, time). Returns dict stream->(n,5) rows
    (start, end, mlen, flags, time); with_elided: also [largest elided end of stream 2, of stream 3] (-1: none)."""
    L = len(seq)
    nw = (L + 31) // 32
    fast = fast_words(seq)
    out = {}
    elided_max = []
    for stream in (1, 2, 3):
        ev = events[events[:, 0] == stream]
        rows = []
        last_elided_time = -1  # latest emission time among elided candidates so far (regular: end = time-8)
        cur_bucket = None
        pending_pseudo = None
        # anchored tail: per motif, which calls commit
        tail = ev[ev[:, 4] == -1]
        tail_counts = {}
        for r in tail:
            tail_counts[int(r[3])] = tail_counts.get(int(r[3]), 0) + 1
        tail_seen = {}
        for st, s, e, m, t in ev[:, :5].tolist():
            is_tail = t == -1
            bucket = nw if is_tail else t >> 5
            slow = True if is_tail else not bool(fast[bucket])
            if stream == 1:
                rows.append((s, e, m, 0, 32 * nw if is_tail else t))
                continue
            cut = cut_subst(m) if stream == 2 else cut_anch(m)
            kept = (e - s) >= cut
            flags = 0
            if stream == 3 and is_tail:
                k = tail_seen.get(m, 0)
                tail_seen[m] = k + 1
                commit = tail_counts[m] == 2 and k == 0
                if not commit:
                    flags |= NOCOMMIT
            if not kept:
                if flags & NOCOMMIT:
                    continue
                if not slow:
                    last_elided_time = max(last_elided_time, t)
                    continue
                flags |= DROPPED
            if slow and bucket != cur_bucket:
                # pseudo record in front of the first record of a slow bucket
                pe = -1
                if last_elided_time >= 0:
                    # only elided candidates of earlier buckets count
                    pe = last_elided_time - 8
                rows.append((-1, pe, 0, PSEUDO, 32 * bucket))
            cur_bucket = bucket
            rows.append((s, e, m, flags, 32 * nw if is_tail else t))
        out[stream] = np.array(rows, dtype=np.int64).reshape(-1, 5)
        if stream != 1:
            elided_max.append(last_elided_time - 8 if last_elided_time >= 0 else -1)
    return (out, elided_max) if with_elided else out
