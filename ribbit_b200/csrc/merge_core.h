// ribbit-b200: ordered compaction of the per-band candidate buckets into the three candidate streams
// (SURVEY.md §8 row a9 / kernel K6). Shared by kernels.cu and the CPU emulator used by tests/.
//
// The scan writes, per (band, word w), the records whose emission time falls into word w ("bucket" w) plus a
// Meta word. The reference emits candidates position-major, motifs ascending inside one position
// (parse_substitute_shiftxor.cpp:430-532), so the stream order is (time, mlen, seq): buckets in word order,
// and inside a bucket the records of all bands ordered by Rec::key. The tail flush (bucket = word after the last)
// is ordered by (mlen, seq) (parse_substitute_shiftxor.cpp:534-574).
#ifndef RB_MERGE_CORE_H
#define RB_MERGE_CORE_H

#include "scan_core.h"

namespace rb {

// m0  = nP | nS << 10 | nA << 20 (records of this band in this bucket, per stream)
// m1  = dmaxS | dmaxA << 8 | slow << 16; dmax = 1 + (latest emission time - 32*w) over the fast-word candidates
//       that were elided because they are below the consumer's length cutoff, 0 = none
// off = index of the bucket's first record in the raw record pool (records of one bucket are contiguous, the
//       three streams interleaved in arrival order; Rec::mflags bits 28-29 hold the stream)
struct alignas(16) Meta {
    uint32_t m0, m1, off, pad;
};
static const int REC_STREAM_SHIFT = 28;
RB_HD int meta_cnt(const Meta& m, int stream) { return (int)((m.m0 >> (10 * stream)) & 0x3FF); }
RB_HD int meta_total(const Meta& m) { return meta_cnt(m, 0) + meta_cnt(m, 1) + meta_cnt(m, 2); }
RB_HD int meta_dmax(const Meta& m, int stream) { return stream == STREAM_S ? (int)(m.m1 & 0xFF) : (int)((m.m1 >> 8) & 0xFF); }
RB_HD int meta_slow(const Meta& m) { return (int)((m.m1 >> 16) & 1); }
RB_HD Meta make_meta(uint32_t counts, int dS, int dA, int slow, uint32_t off) {
    Meta m;
    m.m0 = counts;
    m.m1 = (uint32_t)dS | ((uint32_t)dA << 8) | ((uint32_t)slow << 16);
    m.off = off;
    m.pad = 0u;
    return m;
}

// Elided candidates only move the consumer's from_index cursors (parse_substitute_shiftxor.cpp:34-44,
// parse_anchored_shiftxor.cpp:133-153), and only the largest `end` seen since the previous kept candidate matters.
// Every elided candidate comes from a fast word and has end = time - 8, so the largest end before bucket w is the
// latest elided emission time in an earlier bucket, minus 8. It can only change a cursor ahead of a candidate whose
// own end is smaller than that, i.e. one emitted late (at an N or in the tail flush) - those live in slow buckets,
// so a pseudo record is placed in front of every non-empty slow bucket.
// elided_end_code: 0 = no elided candidate in this bucket, else 1 + (largest end of an elided candidate)
RB_HD uint32_t elided_end_code(int w, int dmax) { return dmax ? (uint32_t)(32 * w + (dmax - 1) - 8 + 1) : 0u; }

RB_HD int bucket_has_pseudo(int stream, int slow, int nrec) { return stream != STREAM_P && slow && nrec > 0; }

// Rank of record `r` among the records of the same stream in its bucket (keys are unique: time, mlen, seq).
// src[b] = first raw record of band b in this bucket, n[b] = number of raw records (all streams).
RB_HD int rank_in_bucket(const Rec& r, const Rec* const* src, const int* n, int nbands) {
    const int stream = (r.mflags >> REC_STREAM_SHIFT) & 3;
    int rank = 0;
    for (int b = 0; b < nbands; ++b)
        for (int k = 0; k < n[b]; ++k) {
            const Rec o = src[b][k];
            rank += (((o.mflags >> REC_STREAM_SHIFT) & 3) == stream && o.key < r.key) ? 1 : 0;
        }
    return rank;
}
// Final form of a raw record: absolute emission time (tail bucket: 32*nw), stream bits cleared.
RB_HD Rec finalize_rec(const Rec& r, int w) {
    Rec o = r;
    o.mflags = r.mflags & ((1 << REC_STREAM_SHIFT) - 1);
    o.key = 32 * w + (r.key >> 18);
    return o;
}
RB_HD Rec pseudo_rec(int w, long long elided_end) {
    Rec r;
    r.start = -1;
    r.end = (int)elided_end;
    r.mflags = REC_PSEUDO << 16;
    r.key = 32 * w;
    return r;
}

}  // namespace rb
#endif
