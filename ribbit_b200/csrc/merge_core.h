// ribbit-b200: ordered compaction of the per-band candidate buckets into the three candidate streams
// (SURVEY.md §8 row a9 / kernel K6). Shared by scan_kernels.cu and the CPU emulator used by tests/.
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

// m0 = nP | nS << 8 | nA << 16 (records of this band in this bucket, per stream)
// m1 = dmaxS | dmaxA << 8 | slow << 16; dmax = 1 + (latest emission time - 32*w) over the fast-word candidates
//      that were elided because they are below the consumer's length cutoff, 0 = none
struct Meta {
    uint32_t m0, m1;
};
RB_HD int meta_cnt(const Meta& m, int stream) {
    return stream == STREAM_P ? (int)(m.m0 & 0xFF) : stream == STREAM_S ? (int)((m.m0 >> 8) & 0xFF) : (int)(m.m0 >> 16);
}
RB_HD int meta_dmax(const Meta& m, int stream) { return stream == STREAM_S ? (int)(m.m1 & 0xFF) : (int)((m.m1 >> 8) & 0xFF); }
RB_HD int meta_slow(const Meta& m) { return (int)((m.m1 >> 16) & 1); }
RB_HD Meta make_meta(int nP, int nS, int nA, int dS, int dA, int slow) {
    Meta m;
    m.m0 = (uint32_t)nP | ((uint32_t)nS << 8) | ((uint32_t)nA << 16);
    m.m1 = (uint32_t)dS | ((uint32_t)dA << 8) | ((uint32_t)slow << 16);
    return m;
}

// Elided candidates only move the consumer's from_index cursors (parse_substitute_shiftxor.cpp:34-44,
// parse_anchored_shiftxor.cpp:133-153), and only the largest `end` seen since the previous kept candidate matters.
// Every elided candidate comes from a fast word and has end = time - 8, so the largest end before bucket w is the
// latest elided emission time in an earlier bucket, minus 8. It can only change a cursor ahead of a candidate whose
// own end is smaller than that, i.e. one emitted late (at an N or in the tail flush) - those live in slow buckets,
// so a pseudo record is placed in front of every non-empty slow bucket. meta[b] points at word 0 of the contig.
RB_HD int pseudo_end_before(const Meta* const* meta, int nbands, int w, int stream) {
    for (int q = w - 1; q >= 0; --q) {
        int best = 0;
        for (int b = 0; b < nbands; ++b) {
            const int d = meta_dmax(meta[b][q], stream);
            if (d > best) best = d;
        }
        if (best) return 32 * q + (best - 1) - 8;
    }
    return -1;
}

RB_HD int bucket_has_pseudo(int stream, int slow, int nrec) { return stream != STREAM_P && slow && nrec > 0; }

// Writes bucket w of one stream in reference order. src[b] = first record of band b's part, n[b] its size.
// Returns the number of records written (including the pseudo record).
RB_HD int merge_bucket(Rec* dst, const Rec* const* src, const int* n, int nbands, int stream, int slow, int w,
                       const Meta* const* meta) {
    int total = 0;
    for (int b = 0; b < nbands; ++b) total += n[b];
    int o = 0;
    if (bucket_has_pseudo(stream, slow, total)) {
        Rec r;
        r.start = -1;
        r.end = pseudo_end_before(meta, nbands, w, stream);
        r.mflags = REC_PSEUDO << 16;
        r.key = 32 * w;
        dst[o++] = r;
    }
    for (int b = 0; b < nbands; ++b) {
        for (int i = 0; i < n[b]; ++i) {
            const Rec r = src[b][i];
            int rank = 0;  // keys are unique inside a bucket: (time, mlen, seq)
            for (int b2 = 0; b2 < nbands; ++b2)
                for (int k = 0; k < n[b2]; ++k) rank += (src[b2][k].key < r.key) ? 1 : 0;
            Rec out = r;
            out.key = 32 * w + (r.key >> 18);  // absolute emission time (tail bucket: 32*nw)
            dst[o + rank] = out;
        }
    }
    return o + total;
}

}  // namespace rb
#endif
