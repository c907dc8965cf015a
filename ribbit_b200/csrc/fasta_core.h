// ribbit-b200: K0 — slice arithmetic of the FASTA ingest (fasta_kernels.cu), shared with the CPU emulator of the tests.
// A slice = the 16 bytes one thread handles; everything here is pure mask arithmetic on those bytes.
#ifndef RB_FASTA_CORE_H
#define RB_FASTA_CORE_H
#include "scan_core.h"

namespace rb {

constexpr int FT_THREADS = 256;
constexpr int FT_PER = 16;
constexpr int FT_TILE = FT_THREADS * FT_PER;  // 4096 bytes

enum : int { LT_SEQ = 0, LT_HDR = 1, LT_NONE = 2 };

// 16 bytes of text as bit masks (bit i = byte i of the slice): newlines, header starts ('>' behind a '\n'), header bytes
// that follow a header start inside the slice, and the bytes in front of the slice's first control byte (their line type
// comes from earlier slices).
struct Slice {
    uint32_t w[4];     // the bytes
    uint32_t valid;    // bytes inside the text
    uint32_t nl, hs;   // control bytes
    uint32_t hdr;      // bytes of header lines that start inside the slice (the '>' included, the '\n' not)
    uint32_t before;   // valid bytes in front of the first control byte
};

// bit k = byte k of `word` equals the byte replicated in `pattern` (device: __vcmpeq4 + one multiply gathers the four
// compare results into a nibble)
RB_HD uint32_t eq_mask4(uint32_t word, uint32_t pattern) {
#ifdef __CUDA_ARCH__
    return ((__vcmpeq4(word, pattern) & 0x08040201u) * 0x01010101u) >> 24;
#else
    uint32_t m = 0;
    for (int k = 0; k < 4; ++k) m |= (uint32_t)(((word >> (8 * k)) & 0xFFu) == ((pattern >> (8 * k)) & 0xFFu)) << k;
    return m;
#endif
}

// r.w holds the slice's bytes (zero beyond n); n = bytes inside the text; prev_nl = the byte in front of the slice is a
// newline (or the slice starts the text)
RB_HD void slice_masks(Slice& r, int n, uint32_t prev_nl) {
    r.valid = (1u << n) - 1u;
    uint32_t nl = 0, gt = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        nl |= eq_mask4(r.w[k], 0x0A0A0A0Au) << (4 * k);
        gt |= eq_mask4(r.w[k], 0x3E3E3E3Eu) << (4 * k);
    }
    r.nl = nl & r.valid;
    r.hs = gt & ((r.nl << 1) | prev_nl) & r.valid;
    // a header start at bit a and its newline at bit b > a: (X - hs) ^ X covers a..b (no other control byte lies between)
    const uint32_t X = r.nl | 0x10000u;
    r.hdr = ((X - r.hs) ^ X) & ~r.nl & 0xFFFFu;
    const uint32_t ctrl = r.nl | r.hs;
    r.before = (ctrl ? ((ctrl & (0u - ctrl)) - 1u) : 0xFFFFu) & r.valid;
}

RB_HD uint32_t slice_byte(const Slice& d, int i) { return (d.w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }

// The slice's last control byte as (index in tile) * 2 + line type it starts, -1 if none.
RB_HD int last_control(const Slice& d, int first_index) {
    const uint32_t ctrl = d.nl | d.hs;
    if (!ctrl) return -1;
    const int pos = 31 - clz32(ctrl);
    return (first_index + pos) * 2 + (int)((d.hs >> pos) & 1u);
}

// Sequence bytes of a slice whose line type at its first byte is `type` (LT_NONE in the counting pass: the bytes in front
// of the first control byte are then reported apart, in `before`).
RB_HD uint32_t slice_seq_after(const Slice& d) { return d.valid & ~d.before & ~d.nl & ~d.hdr; }
RB_HD uint32_t slice_seq(const Slice& d, int type0) { return slice_seq_after(d) | (type0 == LT_SEQ ? d.before : 0u); }

}  // namespace rb
#endif
