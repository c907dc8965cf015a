// ribbit-b200: band / chunk layout shared by the host API, the kernels and the CPU emulator.
#ifndef RB_LAYOUT_H
#define RB_LAYOUT_H
#include <stdint.h>
#include "scan_core.h"

namespace rb {

// Shifts s_lo..s_hi (ribbit.cpp:240-243) are laid over warp lanes: a band owns up to 28 consecutive motif sizes and
// its lanes cover shifts [m0-2, m0-2+bw), i.e. two halo shifts on either side whose anchor words feed B_m
// (fasta_utils.cpp:150-156). When all shifts fit in fewer than 32 lanes several chunks share one warp (groups).
struct BandLayout {
    int m_lo, m_hi, s_lo, s_hi;
    int nbands;  // bands per chunk
    int mpb;     // motif sizes per band
    int bw;      // lanes per band (8, 16 or 32)
    int groups;  // 32 / bw chunks processed side by side in one warp
    int guard;   // zero words after each contig (and one before)
};

RB_HD BandLayout make_layout(int m_lo, int m_hi) {
    BandLayout b;
    b.m_lo = m_lo; b.m_hi = m_hi;
    b.s_lo = (m_lo > 2) ? m_lo - 2 : 1;
    b.s_hi = m_hi + 2;
    const int nm = m_hi - m_lo + 1;
    if (nm + 4 <= 32) {
        b.nbands = 1; b.mpb = nm;
        b.bw = 8;
        while (b.bw < nm + 4) b.bw *= 2;
    } else {
        b.nbands = (nm + 27) / 28;
        b.mpb = (nm + b.nbands - 1) / b.nbands;
        b.bw = 32;
    }
    b.groups = 32 / b.bw;
    b.guard = (b.s_hi >> 5) + 5;
    return b;
}
RB_HD int band_m0(const BandLayout& b, int band) { return b.m_lo + band * b.mpb; }
RB_HD int band_m1(const BandLayout& b, int band) {
    const int m1 = band_m0(b, band) + b.mpb - 1;
    return m1 < b.m_hi ? m1 : b.m_hi;
}
RB_HD LaneCfg band_lane_cfg(const BandLayout& b, int band, int j) {
    const int m0 = band_m0(b, band);
    return make_lane_cfg(m0 - 2 + j, b.m_lo, b.m_hi, b.s_lo, b.s_hi, m0, band_m1(b, band));
}

// first warm-up length in words (grown x4 on every failed synchronisation)
static const int WARMUP_WORDS = 4;

struct Contig {
    int64_t word_base;   // index of word 0 in the plane array (word_base-1 and nw..nw+guard-1 are guard words)
    int64_t ascii_off;   // offset of the first base in the ASCII buffer
    int32_t L;           // bases
    int32_t nw;          // ceil(L/32)
};
struct Chunk {
    int32_t contig;
    int32_t w0, w1;      // owned words [w0, w1) of the contig
    int32_t last;        // 1 if w1 == nw: this chunk also owns the tail bucket nw
};

}  // namespace rb
#endif
