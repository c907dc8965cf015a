// ribbit-b200: ordered compaction of the per-band candidate buckets into the three candidate streams
// (SURVEY.md §8 row a9 / kernel K6). Shared by kernels.cu and the CPU emulator used by tests/.
//
// The scan writes, per (band, word w), the SLOTS whose emission time falls into word w ("bucket" w) plus a Meta
// word. A slot (16 bytes, struct Rec) is either
//   * a record: one candidate, written by the bit-serial path of slow words and by the tail flush, or
//   * a mask entry (REC_ENTRY), written by fast words: up to 32 candidates of one (motif size, stream) whose emission
//     times lie in this word, as bit masks. Layout: start = E mask (perfect stream: run ends), end = S mask (perfect
//     stream: run starts), key = position of the latest S bit in front of the word, mflags as for records.
// The reference emits candidates position-major, motifs ascending inside one position
// (parse_substitute_shiftxor.cpp:430-532), so the stream order is (time, mlen, seq): buckets in word order,
// and inside a bucket the candidates of all bands ordered by (time, mlen, seq). The tail flush (bucket = word after the
// last) is ordered by (mlen, seq) (parse_substitute_shiftxor.cpp:534-574).
#ifndef RB_MERGE_CORE_H
#define RB_MERGE_CORE_H

#include "scan_core.h"

namespace rb {

// off  = index of the bucket's first slot in the raw pool (slots of one bucket are contiguous, the three streams
//        interleaved in arrival order; Rec::mflags bits 28-29 hold the stream)
// info = nslots | dmaxS << 12 | dmaxA << 18 | slow << 24; dmax = 1 + (latest emission time - 32*w) over the fast-word
//        candidates that were elided (below the consumer's length cutoff), 0 = none
// Beside it, per (band, bucket), the candidate counts per stream: nP | nS << 10 | nA << 20 (pack_counts).
struct alignas(8) Meta {
    uint32_t off, info;
};
static const int REC_STREAM_SHIFT = 28;
static const int META_MAX_SLOTS = 0xFFF;
RB_HD int meta_slots(const Meta& m) { return (int)(m.info & 0xFFFu); }
RB_HD int meta_dmax(const Meta& m, int stream) { return stream == STREAM_S ? (int)((m.info >> 12) & 0x3F) : (int)((m.info >> 18) & 0x3F); }
RB_HD int meta_slow(const Meta& m) { return (int)((m.info >> 24) & 1); }
RB_HD Meta make_meta(int nslots, int dS, int dA, int slow, uint32_t off) {
    Meta m;
    m.off = off;
    m.info = (uint32_t)nslots | ((uint32_t)dS << 12) | ((uint32_t)dA << 18) | ((uint32_t)slow << 24);
    return m;
}

RB_HD uint32_t pack_counts(int stream, int n) { return (uint32_t)n << (10 * stream); }
RB_HD int counts_of(uint32_t c, int stream) { return stream == STREAM_A ? (int)(c >> 20) : (int)((c >> (10 * stream)) & 0x3FFu); }
// a field of the summed counts overflowed (more than 1023 perfect / substitution or 4095 anchored candidates in one
// (band, word)): cannot happen in fast words (29 motif lanes x 32 positions = 928), checked for slow words
RB_HD bool counts_overflow(int nP, int nS, int nA) { return nP > 0x3FF || nS > 0x3FF || nA > 0xFFF; }

RB_HD int rec_stream(const Rec& r) { return (r.mflags >> REC_STREAM_SHIFT) & 3; }
RB_HD int rec_mlen(const Rec& r) { return r.mflags & 0xFFFF; }
RB_HD bool rec_is_entry(const Rec& r) { return ((r.mflags >> 16) & REC_ENTRY) != 0; }
RB_HD Rec make_entry(int stream, int mlen, uint32_t mask, uint32_t smask, int last) {
    Rec r;
    r.start = (int32_t)mask; r.end = (int32_t)smask;
    r.mflags = mlen | (REC_ENTRY << 16) | (stream << REC_STREAM_SHIFT);
    r.key = last;
    return r;
}

// The candidate behind bit i of a mask entry of bucket w (scan_core.h, "fast word"): window streams: the component
// emitted at E-bit i is (ls, le) = (ts - 7, 32w + i - 8), ts = the latest S bit before i; perfect stream: the run that
// ends at 32w + i, started at the latest run start before i (looked up in the planes when it lies in an earlier word).
// Returns true when it reaches the consumer's cutoff (parse_perfect_shiftxor.cpp:193, parse_substitute_shiftxor.cpp:44,
// parse_anchored_shiftxor.cpp:153).
RB_HD bool entry_interval(const Rec& e, int w, const PlaneWord* cw, int i, int& start, int& end) {
    const int stream = rec_stream(e), m = rec_mlen(e);
    const uint32_t sb = (uint32_t)e.end & lowmask(i);
    const int p0 = 32 * w;
    if (stream == STREAM_P) {
        int a;
        if (sb) a = p0 + 31 - clz32(sb);
        else {
            a = 0;
            for (int k = w - 1; k >= 0; --k) {  // bits 0..i-1 of X_m[w] are ones: the run started in an earlier word
                const uint32_t z = ~(x_word(cw, k, m) & ~cw[k].n);
                if (z) { a = 32 * k + 32 - clz32(z); break; }
            }
        }
        start = a; end = p0 + i;
        return end - start >= cut_perfect(m);
    }
    const int ts = sb ? p0 + 31 - clz32(sb) : e.key;
    start = ts - 7; end = p0 + i - 8;
    return end - start >= (stream == STREAM_S ? cut_subst(m) : cut_anch(m));
}

// Mask of the bits of an entry that reach the cutoff; `elided` = 1 + the highest bit that does not (window streams), else 0.
RB_HD uint32_t entry_kept_mask(const Rec& e, int w, const PlaneWord* cw, int& elided) {
    uint32_t kept = 0u, x = (uint32_t)e.start;
    elided = 0;
    while (x) {
        const int i = ctz32(x);
        x &= x - 1u;
        int s0, e0;
        if (entry_interval(e, w, cw, i, s0, e0)) kept |= 1u << i;
        else elided = i + 1;
    }
    return kept;
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

// Order key of a candidate inside its bucket: (time - 32w) << 18 | mlen << 2 | seq  (Rec::key of a record)
RB_HD uint32_t entry_key(int i, int mlen) { return ((uint32_t)i << 18) | ((uint32_t)mlen << 2); }
// Candidates of slot `o` (same stream assumed; entries hold their kept mask in `start`) whose key is below `key`.
RB_HD int slot_count_below(const Rec& o, uint32_t key) {
    if (!rec_is_entry(o)) return (uint32_t)o.key < key ? 1 : 0;
    const int i = (int)(key >> 18);
    const uint32_t lowkey = key & 0x3FFFFu, mk = (uint32_t)rec_mlen(o) << 2;
    return popc32((uint32_t)o.start & lowmask(i + (mk < lowkey ? 1 : 0)));
}

RB_HD Rec pseudo_rec(int w, long long elided_end) {
    Rec r;
    r.start = -1;
    r.end = (int)elided_end;
    r.mflags = REC_PSEUDO << 16;
    r.key = 32 * w;
    return r;
}
// Final form of a record slot: absolute emission time (tail bucket: 32*nw), stream bits cleared.
RB_HD Rec finalize_rec(const Rec& r, int w) {
    Rec o = r;
    o.mflags = r.mflags & ((1 << REC_STREAM_SHIFT) - 1);
    o.key = 32 * w + (r.key >> 18);
    return o;
}

}  // namespace rb
#endif
