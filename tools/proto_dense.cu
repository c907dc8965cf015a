// Round-2 planning prototype (NOT part of the product, not parity-checked): how fast is the bit-sliced dense work of the
// scan when a warp's lanes are 32 adjacent strips of the SAME motif (uniform control flow) instead of 32 shifts?
// Block = one warp = 32 strips of K words; for a band of 32 shifts it first writes the anchor words of every shift
// to shared memory, then walks every motif of the band over its strip with all carries in registers. Events are
// only counted (popcount of the keep masks), nothing is emitted. Planes are synthetic random words.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ribbit_b200/csrc -o /tmp/proto_dense tools/proto_dense.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "scan_core.h"

using namespace rb;

constexpr int K = 8;          // words per strip
constexpr int WARM = 3;       // warm-up words in front of a strip (carries, smear reach)
constexpr int TILE = 32 * K;  // words per warp

__global__ void __launch_bounds__(32, 20) dense_kernel(const PlaneWord* __restrict__ planes, int nw, int m_lo, int m_hi,
                                                       unsigned long long* __restrict__ out) {
    __shared__ uint32_t s_anchor[32][TILE + 32 * 0 + 8];  // [shift in band][word of the tile] (no warm-up words: see below)
    const int lane = threadIdx.x;
    const int nbands = (m_hi - m_lo + 1 + 27) / 28;
    const int tile = blockIdx.x / nbands, band = blockIdx.x % nbands;
    const int w0 = tile * TILE + lane * K;  // first word of this lane's strip
    const PlaneWord* cw = planes + 8;       // guard words in front
    const int mpb = (m_hi - m_lo + 1 + nbands - 1) / nbands;
    const int bm0 = m_lo + band * mpb, bm1 = min(m_hi, bm0 + mpb - 1);
    if (w0 + K > nw) return;
    unsigned long long acc = 0;
    // phase 1: anchors of shifts bm0-2 .. bm1+2 for the strip's own words (stateless per word here, as the export kernel)
    for (int s = max(1, bm0 - 2); s <= bm1 + 2; ++s) {
        uint32_t xp = x_word(cw, w0 - 1, s), x = x_word(cw, w0, s);
        int lenL = clz32(~xp);
        for (int i = 0; i < K; ++i) {
            const uint32_t xn = x_word(cw, w0 + i + 1, s);
            const int K2 = 2 * s;
            const int lead = ctz32(~x), trail = clz32(~x), leadn = ctz32(~xn);
            const uint32_t l1 = fsl(xp, x, 1), l2 = fsl(xp, x, 2), r1 = fsr(x, xn, 1), r2 = fsr(x, xn, 2);
            uint32_t a = x & ((l1 & (l2 | r1)) | (r1 & r2));
            a &= (lenL + lead >= K2) ? ~lowmask(lead) : 0xFFFFFFFFu;
            a &= (trail + leadn >= K2) ? lowmask(32 - trail) : 0xFFFFFFFFu;
            lenL = (x == 0xFFFFFFFFu) ? lenL + 32 : trail;
            s_anchor[s - (bm0 - 2)][lane * K + i] = a;
            xp = x; x = xn;
        }
    }
    __syncwarp();
    // phase 2: every motif of the band over the strip (+ warm-up words), uniform control flow
    for (int m = bm0; m <= bm1; ++m) {
        LaneCfg cfg = make_lane_cfg(m, m_lo, m_hi, 1, m_hi + 2, bm0, bm1);
        WinCarry cs{}, ca{};
        EvCarry es{}, ea{};
        uint32_t sm[7] = {0, 0, 0, 0, 0, 0, 0};
        uint32_t pa2 = 0, pa6 = 0;
        uint32_t xp = x_word(cw, w0 - WARM - 1, m);
        XCache xc; xc.idx = -1 << 30; xc.h = xc.l = 0;
        uint32_t x = x_word_cached(cw, w0 - WARM, m, xc);
        for (int i = -WARM; i < K; ++i) {
            const uint32_t xn = x_word_next(cw, w0 + i + 1, m, xc);
            uint32_t b = x;
            if (i >= 0) {
                const int row = m - (bm0 - 2), col = lane * K + i;
                b |= s_anchor[row - 2 < 0 ? 0 : row - 2][col] | s_anchor[row - 1][col] | s_anchor[row + 1][col] | s_anchor[row + 2][col];
            }
            const uint32_t passS = ~fail_ge2(x, cs), passA = ~fail_ge3(b, ca);
            uint32_t sS, eS, sSp, sA, eA, sAp;
            ev_step(passS, es, sS, eS, sSp);
            ev_step(passA, ea, sA, eA, sAp);
            // smear with uniform shifts
            uint32_t v = fsl(sAp, sA, 1);
#pragma unroll
            for (int l = 0; l < 7; ++l) {
                const int d = (l < 5) ? (cfg.dA >> (6 * l)) & 63 : (cfg.dA2 >> (6 * (l - 5))) & 63;
                if (d) { const uint32_t nv = v | fslc(sm[l], v, d); sm[l] = v; v = nv; }
            }
            const uint32_t xs1 = fsl(xp, x, 1);
            const uint32_t cand = ~x & xs1 & six_ones(x, xs1, pa2, pa6);
            const uint32_t keepA = eA & ~v, keepS = eS & ~(fsl(sSp, sS, 9) | fsl(sSp, sS, 10));
            if (i >= 0) acc += __popc(keepA) + __popc(keepS) * 1000ull + __popc(cand) * 1000000ull;
            xp = x; x = xn;
        }
    }
    atomicAdd(out, acc);
}

int main(int argc, char** argv) {
    const long long L = argc > 1 ? atoll(argv[1]) : 46700000ll;
    const int nw = (int)((L + 31) / 32);
    std::vector<PlaneWord> h((size_t)nw + 64);
    srand(1);
    for (auto& p : h) { p.h = ((uint32_t)rand() << 16) ^ rand(); p.l = ((uint32_t)rand() << 16) ^ rand(); p.n = 0; p.v = ~0u; }
    PlaneWord* d; unsigned long long* out;
    cudaMalloc(&d, h.size() * sizeof(PlaneWord)); cudaMalloc(&out, 8);
    cudaMemcpy(d, h.data(), h.size() * sizeof(PlaneWord), cudaMemcpyHostToDevice);
    const int nbands = 4, tiles = (nw - 16) / TILE;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(out, 0, 8);
        cudaEventRecord(e0);
        dense_kernel<<<tiles * nbands, 32>>>(d, nw - 16, 2, 100, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long r; cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost);
        printf("dense prototype: %lld bases, %d tiles x %d bands, %.3f ms -> %.1f Gbp/s (acc %llu) %s\n", L, tiles, nbands, ms, L / ms / 1e6, r,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
