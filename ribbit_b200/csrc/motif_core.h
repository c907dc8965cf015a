// ribbit-b200: row scoring of the consensus-motif search for motif sizes > 10 (K7), shared by the kernel
// (motif_kernels.cu) and the CPU emulator of the tests. Reference: mostFrequentLongerMotif, parse_seed.cpp:153-256.
//
// The reference walks, for every row (candidate motif start) of a seed, the repeat unit by unit downstream and upstream;
// at every unit it counts the matches of the row's motif against the five diagonals -2..+2 of the dot matrix
// (MATRIX[p][q] = base p is A/C/G/T and equals base q, fasta_utils.cpp:88-114), keeps the best diagonal (first wins)
// and drifts by its offset. Here the five counts of a unit come from the packed planes 32 bases at a time: the row's
// bases and a 64-base window of the columns are cut out with funnel shifts, compared with XOR/OR, masked to the
// reference's loop bounds and popcounted.
#ifndef RB_MOTIF_CORE_H
#define RB_MOTIF_CORE_H
#include "scan_core.h"

namespace rb {

// plane word w of a contig; words past the contig read as the guard word (all N). w >= -1 always (the guard in front).
RB_HD PlaneWord motif_ldw(const PlaneWord* __restrict__ cw, int w, int nw) {
    const int i = w < nw ? w : nw;
#ifdef __CUDA_ARCH__
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(cw + i));
    PlaneWord p;
    p.h = v.x; p.l = v.y; p.n = v.z; p.v = v.w;
    return p;
#else
    return cw[i];
#endif
}

RB_HD int clamp32(int v) { return v < 0 ? 0 : (v > 32 ? 32 : v); }

// cnt[x + 2] = #{ j in [jlo(x), jhi[x + 2]) : MATRIX[r0 + j][c0 + x + j] }, x = -2..2; jlo(x) = max(0, -x) if PREFIX else 0.
// r0 >= 0; c0 - 2 >= -2.
template <bool PREFIX>
RB_HD void diag5(const PlaneWord* __restrict__ cw, int nw, int r0, int c0, const int (&jhi)[5], int (&cnt)[5]) {
    int jmax = jhi[0];
#pragma unroll
    for (int x = 1; x < 5; ++x) jmax = jhi[x] > jmax ? jhi[x] : jmax;
#pragma unroll
    for (int x = 0; x < 5; ++x) cnt[x] = 0;
    for (int k = 0; k < jmax; k += 32) {
        const int rp = r0 + k, rw = rp >> 5, rb = rp & 31;
        const PlaneWord ra = motif_ldw(cw, rw, nw), rb1 = motif_ldw(cw, rw + 1, nw);
        const uint32_t rh = fsr(ra.h, rb1.h, rb), rl = fsr(ra.l, rb1.l, rb), rn = fsr(ra.n, rb1.n, rb);
        const int cp = c0 - 2 + k, cwd = cp >> 5, cb = cp & 31;  // arithmetic shift: cp may be -2 / -1 (front guard word)
        const PlaneWord c0w = motif_ldw(cw, cwd, nw), c1w = motif_ldw(cw, cwd + 1, nw), c2w = motif_ldw(cw, cwd + 2, nw);
        const uint32_t h0 = fsr(c0w.h, c1w.h, cb), h1 = fsr(c1w.h, c2w.h, cb);
        const uint32_t l0 = fsr(c0w.l, c1w.l, cb), l1 = fsr(c1w.l, c2w.l, cb);
        const uint32_t n0 = fsr(c0w.n, c1w.n, cb), n1 = fsr(c1w.n, c2w.n, cb);
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            const uint32_t ch = fsr(h0, h1, x), cl = fsr(l0, l1, x), cn = fsr(n0, n1, x);
            uint32_t eq = ~((rh ^ ch) | (rl ^ cl) | rn | cn);
            eq &= lowmask(clamp32(jhi[x] - k));
            if (PREFIX && x < 2) eq &= ~lowmask(clamp32(2 - x - k));
            cnt[x] += popc32(eq);
        }
    }
}

// best diagonal: first of x = -2..2 with the largest count; all zero -> x = -2 (parse_seed.cpp:185,196)
RB_HD void best5(const int (&cnt)[5], int& bc, int& bx) {
    bc = 0; bx = -2;
#pragma unroll
    for (int x = 0; x < 5; ++x)
        if (cnt[x] > bc) { bc = cnt[x]; bx = x - 2; }
}

// row_count of one row (parse_seed.cpp:179-238). seed_end = seed_start + seed_sequence_length; seed_start <= row <= seed_end - m.
RB_HD int motif_row_score(const PlaneWord* __restrict__ cw, int nw, int seed_start, int seed_end, int m, int row) {
    int total = 0, cnt[5], jhi[5], bc, bx;
    int d = row + m;  // downstream, parse_seed.cpp:183-199: columns stop at seed_end
    while (d < seed_end) {
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            const int room = seed_end - (d + x - 2);
            jhi[x] = room < m ? room : m;
        }
        diag5<false>(cw, nw, row, d, jhi, cnt);
        best5(cnt, bc, bx);
        total += bc;
        d += bx + m;
    }
    int u = row - m;  // upstream, parse_seed.cpp:201-218: a diagonal starting left of position 0 counts nothing
    while (u > seed_start) {
#pragma unroll
        for (int x = 0; x < 5; ++x) jhi[x] = (u + x - 2 < 0) ? 0 : m;
        diag5<false>(cw, nw, row, u, jhi, cnt);
        best5(cnt, bc, bx);
        total += bc;
        u += bx - m;
    }
    if (u < seed_start && seed_start - u < m) {  // partial unit at the seed start, parse_seed.cpp:220-238
        const int rows = m + (u - seed_start);   // 1 .. m-1
        const int pc = seed_start + rows - 1;
        // the reference walks i = 0..rows-1 from (row+m-1, pc+x) towards the top left and stops at the first column outside
        // [seed_start, seed_end); in ascending order j = rows-1-i: rows row+m-rows+j against columns seed_start+x+j
#pragma unroll
        for (int x = 0; x < 5; ++x) jhi[x] = (pc + x - 2 >= seed_end) ? 0 : rows;
        diag5<true>(cw, nw, row + m - rows, seed_start, jhi, cnt);
        best5(cnt, bc, bx);
        total += bc;
    }
    return total;
}

}  // namespace rb
#endif
