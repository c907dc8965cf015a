// ribbit-b200: K7 — consensus-motif row search for motif sizes > 10 (mostFrequentLongerMotif, parse_seed.cpp:153-256),
// batched over seeds. One warp per item = (seed, slab of MOTIF_SLAB rows); lane = row, so the 32 lanes read
// neighbouring plane words (L1-resident: a seed spans a few hundred bytes of planes). Every lane scores its rows with
// motif_row_score (motif_core.h); the warp's best (largest count, then smallest row = the reference's first-wins scan
// over rows, parse_seed.cpp:240-242) goes into the seed's 64-bit key with one atomicMax.
#include "kernels.h"
#include "motif_core.h"

namespace rb {

__global__ void __launch_bounds__(32) motif_row_kernel(DevBatch b, const int4* __restrict__ seeds, const int2* __restrict__ items,
                                                       unsigned long long* __restrict__ keys) {
    const int2 it = items[blockIdx.x];  // seed index, first row of the slab relative to the seed start
    const int4 sd = seeds[it.x];        // contig, seed_start, seed_end (= start + seed_sequence_length), mlen
    const Contig cg = b.contigs[sd.x];
    const PlaneWord* __restrict__ cw = b.planes + cg.word_base;
    const int seed_start = sd.y, seed_end = sd.z, m = sd.w;
    const int row0 = seed_start + it.y;
    const int row_last = min(seed_end - m, row0 + MOTIF_SLAB - 1);
    unsigned long long best = 0;
    for (int row = row0 + (int)threadIdx.x; row <= row_last; row += 32) {
        const int sc = motif_row_score(cw, cg.nw, seed_start, seed_end, m, row);
        const unsigned long long key = ((unsigned long long)(uint32_t)sc << 32) | (uint32_t)(0x7FFFFFFF - row);
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        best = other > best ? other : best;
    }
    if (threadIdx.x == 0 && (best >> 32)) atomicMax(keys + it.x, best);
}

void launch_motif_rows(const DevBatch& b, const void* seeds, const void* items, long long n_items, void* keys, cudaStream_t st) {
    if (n_items == 0) return;
    motif_row_kernel<<<(unsigned)n_items, 32, 0, st>>>(b, (const int4*)seeds, (const int2*)items, (unsigned long long*)keys);
}

}  // namespace rb
