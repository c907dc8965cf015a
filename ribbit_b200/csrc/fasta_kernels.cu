// ribbit-b200: K0 — FASTA text -> contig bases on the device. Reference reader: the getline loop
// ribbit.cpp:269-280 — a line whose FIRST byte is '>' is a header, every other line is appended verbatim (so a '\r'
// or a '>' inside a line stay in the sequence and become N bases later), '\n' is the only separator.
//
// HBM-bound byte work in three passes over 4 KiB tiles (one block of 256 threads, 16 bytes per thread):
//   fasta_tile_kernel   per tile: sequence bytes before the first control byte (their line type is decided by earlier
//                       tiles), sequence bytes after it, header count, line type at the tile's end
//   fasta_scan_kernel   one block: line type at every tile's start, exclusive prefixes of sequence bytes and headers
//   fasta_strip_kernel  per tile: sequence bytes staged in shared memory at their rank, then written out coalesced;
//                       per header: its offset in the text and the number of sequence bytes in front of it
// A control byte is '\n' (the next line is sequence until proven otherwise) or a '>' that follows a '\n' / starts the text.
#include "fasta_core.h"
#include "kernels.h"

namespace rb {

namespace {

__device__ __forceinline__ Slice load_slice(const uint8_t* __restrict__ text, long long nbytes, long long at) {
    Slice r;
    const int n = (int)max(0ll, min((long long)FT_PER, nbytes - at));
    if (n == FT_PER) {  // text comes from cudaMalloc and `at` is a multiple of 16
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + at));
        r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * k + j < n) x |= (uint32_t)__ldg(text + at + 4 * k + j) << (8 * j);
            r.w[k] = x;
        }
    }
    const uint32_t prev_nl = (at > 0 && at <= nbytes) ? (__ldg(text + at - 1) == '\n') : 1u;  // the text starts a line
    slice_masks(r, n, prev_nl);
    return r;
}

// inclusive max-scan over the block (values >= -1); returns the EXCLUSIVE result for this thread, *total = block max
__device__ __forceinline__ int block_excl_max(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x = max(x, y);
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    int before = -1, all = -1;
#pragma unroll
    for (int w = 0; w < FT_THREADS / 32; ++w) {
        const int t = s_warp[w];
        if (w < warp) before = max(before, t);
        all = max(all, t);
    }
    __syncthreads();
    int ex = __shfl_up_sync(0xFFFFFFFFu, x, 1);
    if (lane == 0) ex = -1;
    if (total) *total = all;
    return max(before, ex);
}

// exclusive sum-scan over the block; *total = block sum
__device__ __forceinline__ int block_excl_sum(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < FT_THREADS / 32; ++w) {
        const int t = s_warp[w];
        if (w < warp) before += t;
        all += t;
    }
    __syncthreads();
    if (total) *total = all;
    return before + x - v;
}

}  // namespace

// tile summary: x = sequence-candidate bytes before the tile's first control byte, y = sequence bytes after it,
// z = headers, w = line type at the tile's end (LT_NONE: no control byte in the tile)
__global__ void __launch_bounds__(FT_THREADS) fasta_tile_kernel(const uint8_t* __restrict__ text, long long nbytes, long long tile_first,
                                                                int4* __restrict__ tiles) {
    __shared__ int s_warp[FT_THREADS / 32];
    __shared__ int s_sum[3];
    const long long tile = tile_first + blockIdx.x;
    const long long at = tile * FT_TILE + (long long)threadIdx.x * FT_PER;
    const Slice d = load_slice(text, nbytes, at);
    if (threadIdx.x < 3) s_sum[threadIdx.x] = 0;
    int tile_last;
    const int before = block_excl_max(last_control(d, threadIdx.x * FT_PER), s_warp, &tile_last);
    const int type = before < 0 ? LT_NONE : (before & 1);
    const int n_before = __popc(d.before), n_after = __popc(slice_seq_after(d));
    int pre = type == LT_NONE ? n_before : 0;
    int known = n_after + (type == LT_SEQ ? n_before : 0);
    int nh = __popc(d.hs);
    pre = __reduce_add_sync(0xFFFFFFFFu, pre);
    known = __reduce_add_sync(0xFFFFFFFFu, known);
    nh = __reduce_add_sync(0xFFFFFFFFu, nh);
    if ((threadIdx.x & 31) == 0) {
        if (pre) atomicAdd(&s_sum[0], pre);
        if (known) atomicAdd(&s_sum[1], known);
        if (nh) atomicAdd(&s_sum[2], nh);
    }
    __syncthreads();
    if (threadIdx.x == 0) tiles[tile] = make_int4(s_sum[0], s_sum[1], s_sum[2], tile_last < 0 ? LT_NONE : (tile_last & 1));
}

// One block. info[t] = {sequence bytes before tile t, headers before tile t | line type at its start << 40};
// totals = {sequence bytes, headers}.
__global__ void __launch_bounds__(1024) fasta_scan_kernel(const int4* __restrict__ tiles, long long n_tiles, longlong2* __restrict__ info,
                                                          long long* __restrict__ totals) {
    __shared__ long long s_a[1024], s_b[1024];
    __shared__ int s_t[1024];
    const int t = threadIdx.x;
    const long long per = (n_tiles + 1023) / 1024, lo = min(n_tiles, t * per), hi = min(n_tiles, lo + per);
    // line type at the end of this thread's range
    int type = LT_NONE;
    for (long long i = lo; i < hi; ++i) { const int w = tiles[i].w; if (w != LT_NONE) type = w; }
    s_t[t] = type;
    __syncthreads();
    int entry = LT_SEQ;  // the text starts in a sequence line (a headerless file is one unnamed record)
    for (int k = t - 1; k >= 0; --k) if (s_t[k] != LT_NONE) { entry = s_t[k]; break; }
    long long seq = 0, hdr = 0;
    type = entry;
    for (long long i = lo; i < hi; ++i) {
        const int4 v = tiles[i];
        seq += (type == LT_SEQ ? v.x : 0) + v.y;
        hdr += v.z;
        if (v.w != LT_NONE) type = v.w;
    }
    s_a[t] = seq; s_b[t] = hdr;
    __syncthreads();
    if (t == 0) {
        long long a = 0, b = 0;
        for (int k = 0; k < 1024; ++k) { const long long x = s_a[k], y = s_b[k]; s_a[k] = a; s_b[k] = b; a += x; b += y; }
        totals[0] = a; totals[1] = b;
    }
    __syncthreads();
    seq = s_a[t]; hdr = s_b[t]; type = entry;
    for (long long i = lo; i < hi; ++i) {
        const int4 v = tiles[i];
        info[i] = make_longlong2(seq, hdr | ((long long)type << 40));
        seq += (type == LT_SEQ ? v.x : 0) + v.y;
        hdr += v.z;
        if (v.w != LT_NONE) type = v.w;
    }
}

// bases[] receives the sequence bytes in order; for header k: hdr_pos[k] = offset of its '>' in the text,
// hdr_seq[k] = number of sequence bytes in front of it (= where the record it names starts in bases[]).
__global__ void __launch_bounds__(FT_THREADS) fasta_strip_kernel(const uint8_t* __restrict__ text, long long nbytes,
                                                                 const longlong2* __restrict__ info, uint8_t* __restrict__ bases,
                                                                 long long* __restrict__ hdr_pos, long long* __restrict__ hdr_seq) {
    __shared__ int s_warp[FT_THREADS / 32];
    __shared__ uint8_t s_out[FT_TILE];
    const long long at = (long long)blockIdx.x * FT_TILE + (long long)threadIdx.x * FT_PER;
    const Slice d = load_slice(text, nbytes, at);
    const longlong2 ti = info[blockIdx.x];
    const long long seq_base = ti.x, hdr_base = ti.y & ((1ll << 40) - 1);
    const int entry = (int)(ti.y >> 40);
    const int before = block_excl_max(last_control(d, threadIdx.x * FT_PER), s_warp, nullptr);
    const int type0 = before < 0 ? entry : (before & 1);
    const uint32_t seq = slice_seq(d, type0);
    int tile_seq;
    const int rank = block_excl_sum(__popc(seq), s_warp, &tile_seq);
    const int hrank = block_excl_sum(__popc(d.hs), s_warp, nullptr);
    if (seq == 0xFFFFu) {
#pragma unroll
        for (int i = 0; i < FT_PER; ++i) s_out[rank + i] = (uint8_t)slice_byte(d, i);
    } else {
        int r = rank;
#pragma unroll
        for (int i = 0; i < FT_PER; ++i)
            if ((seq >> i) & 1u) s_out[r++] = (uint8_t)slice_byte(d, i);
    }
    for (uint32_t h = d.hs; h; h &= h - 1u) {
        const int i = __ffs((int)h) - 1;
        const long long k = hdr_base + hrank + __popc(d.hs & ((1u << i) - 1u));
        hdr_pos[k] = at + i;
        hdr_seq[k] = seq_base + rank + __popc(seq & ((1u << i) - 1u));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tile_seq; i += FT_THREADS) bases[seq_base + i] = s_out[i];
}

long long fasta_tiles(long long nbytes) { return (nbytes + FT_TILE - 1) / FT_TILE; }

// pass 1 over the tiles [tile_first, tile_first + n_tiles): needs the text up to the end of these tiles on the device (and
// the byte in front of them), so it can run while the rest of the file is still being copied
void launch_fasta_tiles(const void* text, long long nbytes, long long tile_first, long long n_tiles, void* tiles, cudaStream_t st) {
    if (n_tiles > 0) fasta_tile_kernel<<<(unsigned)n_tiles, FT_THREADS, 0, st>>>((const uint8_t*)text, nbytes, tile_first, (int4*)tiles);
}
void launch_fasta_scan(const void* tiles, long long nbytes, void* info, long long* totals, cudaStream_t st) {
    fasta_scan_kernel<<<1, 1024, 0, st>>>((const int4*)tiles, fasta_tiles(nbytes), (longlong2*)info, totals);
}
long long fasta_tile_bytes() { return FT_TILE; }

void launch_fasta_strip(const void* text, long long nbytes, const void* info, void* bases, long long* hdr_pos, long long* hdr_seq,
                        cudaStream_t st) {
    const long long nt = fasta_tiles(nbytes);
    if (nt > 0)
        fasta_strip_kernel<<<(unsigned)nt, FT_THREADS, 0, st>>>((const uint8_t*)text, nbytes, (const longlong2*)info, (uint8_t*)bases,
                                                                hdr_pos, hdr_seq);
}

}  // namespace rb
