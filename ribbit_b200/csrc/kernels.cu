// ribbit-b200: sm_100a kernels of the seed-scanning path.
//   K1 pack_kernel     ASCII -> (h, l, n, v) plane words                 fasta_utils.cpp:78-115
//   K2 scan_kernel     match words + perfect / substitution / anchored    fasta_utils.cpp:117-161,
//                      seed machines, one warp lane per shift             parse_{perfect,substitute,anchored}_shiftxor.cpp
//   K6 merge kernels   ordered compaction of the candidate buckets        call order of addSeedToSeedPositions*
// The per-lane logic lives in scan_core.h / merge_core.h, which the CPU warp emulator of tests/ compiles too.
#include "kernels.h"
#include "scan_tight.h"

namespace rb {

// ---------------------------------------------------------------------------------------------------------------
// K1: one thread per plane word (guard words included).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_segment(const long long* __restrict__ starts, int n, long long x) {
    // largest c in [0, n) with starts[c] <= x
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(starts + mid) <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) pack_kernel(DevBatch b) {
    const long long pw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = pw < b.n_plane_words;
    if (!valid) {  // the warp still votes below
        const unsigned m = __ballot_sync(0xFFFFFFFFu, false);
        if ((threadIdx.x & 31) == 0 && (pw >> 5) < ((b.n_plane_words + 31) >> 5)) b.nmask[pw >> 5] = m;
        return;
    }
    const int c = find_segment(b.plane_start, b.n_contigs, pw);
    const Contig cg = b.contigs[c];
    const long long w = pw - b.plane_start[c] - 1;  // word of the contig, -1 and >= nw are guard words
    PlaneWord o;
    o.h = 0u; o.l = 0u; o.n = 0xFFFFFFFFu; o.v = 0u;
    if (w >= 0 && w < cg.nw) {
        const long long p0 = 32 * w;
        const uint8_t* __restrict__ src = b.ascii + cg.ascii_off;
        unsigned long long nx = 0ull;  // bit i <-> position p0 - 7 + i
        uint32_t h = 0u, l = 0u;
        if (w >= 1 && p0 + 32 <= cg.L && (reinterpret_cast<unsigned long long>(src) & 15ull) == 0ull) {
            // interior word of a 16-byte aligned contig: two uint4 loads (32 bases) + one uint2 (the 8 bases before),
            // four bases at a time with byte-parallel arithmetic
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(src + p0));
            const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(src + p0 + 16));
            const uint2 qh = __ldg(reinterpret_cast<const uint2*>(src + p0 - 8));
            const uint32_t wd[10] = {qh.x, qh.y, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t nn = 0u, nhist = 0u;
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                const uint32_t y = wd[k] | 0x20202020u;  // fasta_utils.cpp:95-113: either case
                // byte == 'a' / 'c' / 'g' / 't' -> 0xFF in that byte
                const uint32_t acgt = (__vcmpeq4(y, 0x61616161u) | __vcmpeq4(y, 0x63636363u) | __vcmpeq4(y, 0x67676767u) |
                                       __vcmpeq4(y, 0x74747474u)) & 0x01010101u;
                // a=0x61 c=0x63 g=0x67 t=0x74: high code bit = bit 2, low code bit = bit 1 ^ bit 2
                const uint32_t hb = (y >> 2) & acgt, lb = ((y >> 1) ^ (y >> 2)) & acgt, nb = acgt ^ 0x01010101u;
                // one bit per byte -> 4 adjacent bits (byte 0 -> bit 0)
                const uint32_t h4 = (hb * 0x10204080u) >> 28, l4 = (lb * 0x10204080u) >> 28, n4 = (nb * 0x10204080u) >> 28;
                if (k < 2) nhist |= n4 << (4 * k);
                else { h |= h4 << (4 * (k - 2)); l |= l4 << (4 * (k - 2)); nn |= n4 << (4 * (k - 2)); }
            }
            nx = ((unsigned long long)nn << 7) | (unsigned long long)(nhist >> 1);  // positions p0-7 .. p0+31
        } else
#pragma unroll
        for (int i = 0; i < 39; ++i) {
            const long long p = p0 - 7 + i;
            uint32_t isn = 1u, hb = 0u, lb = 0u;
            if (p >= 0 && p < cg.L) {
                const uint32_t x = (uint32_t)__ldg(src + p) | 0x20u;  // fasta_utils.cpp:95-113: either case
                const uint32_t a = x == 'a', cc = x == 'c', g = x == 'g', t = x == 't';
                isn = !(a | cc | g | t);
                hb = g | t;
                lb = cc | t;
            }
            nx |= (unsigned long long)isn << i;
            if (i >= 7) { h |= hb << (i - 7); l |= lb << (i - 7); }
        }
        unsigned long long t = nx;
        t |= t >> 1; t |= t >> 2; t |= t >> 4;  // bit i: an N among positions p0-7+i .. p0+i
        o.h = h; o.l = l;
        o.n = (uint32_t)(nx >> 7);
        o.v = ~(uint32_t)t;
    }
    b.planes[pw] = o;
    // one bit per plane word: all 32 positions are N (or padding); the scan's N-run probe reads 32 words per load
    const unsigned m = __ballot_sync(0xFFFFFFFFu, o.n == 0xFFFFFFFFu);
    if ((threadIdx.x & 31) == 0) b.nmask[pw >> 5] = m;
}

void launch_pack(const DevBatch& b, cudaStream_t st) {
    if (b.n_plane_words == 0) return;
    const long long blocks = (b.n_plane_words + 255) / 256;
    pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(b);
}

// ---------------------------------------------------------------------------------------------------------------
// K2: scan. A warp holds 32 / BW items (chunk x band); lane j of an item handles shift band_m0 - 2 + j.
// ---------------------------------------------------------------------------------------------------------------
struct GpuSink {
    Rec* base;
    int cap;
    int* cnt;
    int nslots;        // slots this lane wrote into this bucket
    int n[3];          // candidates per stream this lane put into this bucket
    int dS, dA;
    __device__ __forceinline__ void put(const Rec& r) {
        const int idx = atomicAdd(cnt, 1);
        if (idx < cap) *reinterpret_cast<int4*>(base + idx) = make_int4(r.start, r.end, r.mflags, r.key);
        ++nslots;
    }
    __device__ __forceinline__ void rec(int stream, int start, int end, int mlen, int flags, int key) {
        Rec r;
        r.start = start; r.end = end; r.mflags = mlen | (flags << 16) | (stream << REC_STREAM_SHIFT); r.key = key;
        put(r);
        ++n[stream];
    }
    __device__ __forceinline__ void entry(int stream, int mlen, uint32_t mask, uint32_t smask, int last) {
        put(make_entry(stream, mlen, mask, smask, last));
        n[stream] += __popc(mask);
    }
    __device__ __forceinline__ void dropped_mask(int stream, uint32_t el) {
        const int d = 32 - __clz((int)el);  // 0 when nothing was elided
        if (stream == STREAM_S) dS = max(dS, d); else dA = max(dA, d);
    }
    __device__ __forceinline__ void reset() { nslots = 0; n[0] = n[1] = n[2] = 0; dS = 0; dA = 0; }
};

static const int SCAN_WARPS = 1;

// fast -> slow transition (scan_core.h, lane_to_slow) out of line and with scalar arguments, so that the rare call does
// not weigh on the register allocation of the scan loop
__device__ __noinline__ bool slow_entry(uint32_t Ps, uint32_t r8s, uint32_t Pa, uint32_t r8a, uint32_t x_prev, const PlaneWord* cw, int w,
                                        int s, SlowEntry& e) {
    EvCarry cs, ca;
    cs.P = Ps; cs.r8 = r8s; ca.P = Pa; ca.r8 = r8a;
    SlowEntry t;
    if (!win_from_fast(cs, 32 * (w - 1), t.S) || !win_from_fast(ca, 32 * (w - 1), t.A)) return false;
    t.pst = (x_prev >> 31) ? perfect_run_start(cw, w - 1, 32, s, x_prev) : -1;
    e = t;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// The tight loop (scan_tight.h): consecutive fast, emitting words of the item(s) of a warp. Out of line: it has its own
// register allocation, the general path's state waits in local memory meanwhile.
// ---------------------------------------------------------------------------------------------------------------
struct TightIO {
    TightState t;
    TightCfg c;
    const PlaneWord* cw;
    Meta* meta;        // bucket metadata of the item's (band, contig)
    uint32_t* bcnt;    // candidate counts of the item's (band, contig)
    Rec* raw;          // raw pool
    uint32_t off;      // raw-pool index of the next slot of the item
    uint32_t cap_end;  // first index past the item's reservation
    int w, wend, L;
    int on;            // the group takes part (0: finished or warming up; its lanes only keep the warp converged)
    unsigned lt;       // lanes below this one (read back from memory so that it is not rematerialised from S2R per use)
};

// rare path of the tight loop, out of line: keeps the loop body small (instruction cache) and its registers free
template <int TIER>
__device__ __noinline__ uint32_t tight_rare(const TightCfg& c, TightState& t, const PlaneWord* cw, int w, int L, uint32_t xn, uint32_t a,
                                            int lenL0, bool prev_rare) {
    bool full;
    a = tight_anchor_rare<TIER>(c, t, cw, w, L, xn, a, lenL0, prev_rare, full);
    if (__any_sync(0xFFFFFFFFu, full)) {
        if (full) a = tight_anchor_full(c, t, cw, w, L, xn);
    }
    return a;
}

// exact check of the survivors of the bit-parallel prefilters (scan_core.h kept_exact), out of line like the rare anchors
struct ExactIO {
    uint32_t xA, sA, xS, sS, cand, xc, sx;
    int lastA, lastS;
};
__device__ __noinline__ void tight_exact(ExactIO* e, const PlaneWord* cw, int w, int s, uint32_t needA, uint32_t needS) {
    const int p0 = 32 * w;
    if (needA) e->xA = kept_exact(cut_anch(s), p0, e->xA, e->sA, e->lastA);
    if (needS) e->xS = kept_exact(cut_subst(s), p0, e->xS, e->sS, e->lastS);
    if (e->cand) e->cand = kept_exact_perfect(cw, w, s, cut_perfect(s), e->xc, e->sx, e->cand);
}

template <typename T>
__device__ __forceinline__ T* uniform_ptr(T* p) {
    const unsigned long long v = (unsigned long long)p;
    const unsigned lo = __shfl_sync(0xFFFFFFFFu, (unsigned)v, 0), hi = __shfl_sync(0xFFFFFFFFu, (unsigned)(v >> 32), 0);
    return (T*)(((unsigned long long)hi << 32) | lo);
}

// everything the loop carries besides the per-word lane state
struct TightLoop {
    TightCfg c;
    const PlaneWord* cw;
    Rec* raw;
    const uint4* po;   // plane word w + 1 (all lanes)
    const uint2* pb;   // h / l of plane word w + 1 + (s >> 5) + 1 (per lane)
    Meta* mp;          // bucket metadata of word w
    uint32_t* cp;      // candidate counts of word w
    uint32_t off, cap_end;
    int w, wend, L;
    uint32_t vprev, vcur;
    uint4 o;           // plane word w + 1 and
    uint2 bb;          //   the lane's b operand, loaded one step ahead
    unsigned lt;
    int flags;         // bit 0 / 1: a lane flagged word w / w - 1 (TF_SUSC, TF_SUSP); bit 2: the previous word took the rare path
    bool on;
    int zc;
};
enum : int { TF_SUSC = 1, TF_SUSP = 2, TF_PREV_RARE = 4 };

template <int BW, int TIER>
__device__ __forceinline__ bool tight_leave_now(const TightLoop& q) {
    const bool go = !q.on || (q.w < q.wend && (q.vprev & q.vcur) == 0xFFFFFFFFu);
    if (BW == 32) return !go;
    // every group that takes part must have a fast word in front of it, else the warp leaves the loop
    return !__all_sync(0xFFFFFFFFu, go) || !__any_sync(0xFFFFFFFFu, q.on);
}

template <int TIER>
__device__ __forceinline__ void tight_finish(TightIO* io, const TightLoop& q, TightState t) {
    if (TIER != TIER_SMALL && q.c.s && q.on && !(q.flags & TF_PREV_RARE)) t.lenL = tight_lenL_lookup(q.c, t, q.cw, q.w);  // the general path carries the run length
    io->t = t;
    io->off = q.off;
    io->w = q.w;
}

// one word: state `in` (word w-1 rotated in) -> state `out`
template <int BW, int TIER>
__device__ __forceinline__ void tight_step(TightLoop& q, const TightState& in, TightState& t) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr bool WHOLE = BW == 32;  // one item per warp: everything about the item is warp-uniform
    const unsigned gmask = WHOLE ? FULL : (((1u << BW) - 1u) << (((threadIdx.x & 31) / BW) * BW));
    const unsigned lt = q.lt;
    const TightCfg& c = q.c;
    t = in;
    const uint4 o = q.o;
    const uint2 bb = q.bb;
    const int adv = (WHOLE || q.on) ? 1 : 0;  // lanes of a group that sits out keep reading the same (valid) words
    q.po += adv; q.pb += 2 * adv;
    q.o = __ldg(q.po);   // for the next step (always readable: guard words follow every contig)
    q.bb = __ldg(q.pb);
    uint32_t xn, l1;
    bool sus;
    const int lenL0 = t.lenL;
    uint32_t a = tight_phaseA<TIER>(c, t, o.x, o.y, bb.x, bb.y, xn, l1, sus);
    const bool any_sus = __any_sync(FULL, sus);
    bool rare = any_sus;
    if (TIER != TIER_SMALL) rare = any_sus || (q.flags & (TF_SUSC | TF_SUSP)) != 0;
    if (rare) {
        // only xp, xc and lenL of the state are touched: pass a small copy, not the whole TightState
        TightState tr;
        tr.xp = t.xp; tr.xc = t.xc; tr.lenL = t.lenL;
        a = tight_rare<TIER>(c, tr, q.cw, q.w, q.L, xn, a, lenL0, (q.flags & TF_PREV_RARE) != 0);
        t.lenL = tr.lenL;
    }
    q.flags = ((q.flags & TF_SUSC) << 1) | (any_sus ? TF_SUSC : 0) | (rare ? TF_PREV_RARE : 0);
    const uint32_t pair = a | __shfl_down_sync(FULL, a, 1);
    const uint32_t an = __shfl_up_sync(FULL, pair, 2) | __shfl_down_sync(FULL, pair, 1);
    uint32_t passS, passA, cand;
    tight_windows<TIER>(c, t, an, l1, passS, passA, cand);
    const int p0 = 32 * q.w;
    TightOut oa, os;
    tight_events_A<TIER>(c, t, p0, passA, oa);
    cand &= c.mmask;
    os.x = 0u; os.s = 0u; os.e = 0u; os.last = 0;
    const bool on = WHOLE ? true : q.on;
    // perfect and substitution candidates are rare: one vote decides whether their event logic runs at all
    const bool anyPS = __any_sync(FULL, (cand | passS) != 0u);
    const bool runS = anyPS || q.zc < 2;
    if (runS) {
        tight_events_S(t, p0, passS, os);
        q.zc = anyPS ? 0 : q.zc + 1;  // (a perfect-run candidate counts as activity too: one vote instead of two)
    }
    // survivors of the prefilters that are not decided yet: exact check, bit by bit (rare)
    // (substitution stream: the prefilter is the cutoff itself while that is 10, i.e. for motif sizes up to 30)
    // (the anchored filter is exact for every cutoff; LARGE items: every motif size is above 30)
    const uint32_t needA = 0u, needS = (TIER == TIER_LARGE || c.s > 30) ? os.x : 0u;
    if (__any_sync(FULL, (needA | needS | cand) != 0u)) {
        ExactIO ex;
        ex.xA = oa.x; ex.sA = oa.s; ex.lastA = oa.last; ex.xS = os.x; ex.sS = os.s; ex.lastS = os.last;
        ex.cand = cand; ex.xc = t.xc; ex.sx = t.xc & ~l1;
        tight_exact(&ex, q.cw, q.w, c.s, needA, needS);
        oa.x = ex.xA; os.x = ex.xS; cand = ex.cand;
    }
    const uint32_t elA = __reduce_or_sync(gmask, oa.e & ~oa.x);
    uint32_t counts = pack_counts(STREAM_A, __popc(oa.x));
    // mask entries: one slot per (lane, stream) with candidates, packed by ballot
    const uint32_t off_in = q.off;
    uint32_t off = q.off;
    const bool hA = on && oa.x != 0u;
    const unsigned balA = __ballot_sync(FULL, hA);
    if (balA) {
        const unsigned mine = balA & gmask;
        const uint32_t pos = off + __popc(mine & lt);
        if (hA && pos < q.cap_end) *reinterpret_cast<int4*>(q.raw + pos) = make_int4((int)oa.x, (int)oa.s, c.s | (REC_ENTRY << 16) | (STREAM_A << REC_STREAM_SHIFT), oa.last);
        off += __popc(mine);
    }
    uint32_t elS = 0u;
    if (runS) {
        const bool hP = on && cand != 0u, hS = on && os.x != 0u;
        const unsigned balP = __ballot_sync(FULL, hP), balS = __ballot_sync(FULL, hS);
        if (balP | balS) {
            const unsigned mineP = balP & gmask, mineS = balS & gmask;
            const uint32_t posP = off + __popc(mineP & lt);
            if (hP && posP < q.cap_end) *reinterpret_cast<int4*>(q.raw + posP) = make_int4((int)cand, (int)(t.xc & ~l1), c.s | (REC_ENTRY << 16) | (STREAM_P << REC_STREAM_SHIFT), 0);
            off += __popc(mineP);
            const uint32_t posS = off + __popc(mineS & lt);
            if (hS && posS < q.cap_end) *reinterpret_cast<int4*>(q.raw + posS) = make_int4((int)os.x, (int)os.s, c.s | (REC_ENTRY << 16) | (STREAM_S << REC_STREAM_SHIFT), os.last);
            off += __popc(mineS);
            counts += pack_counts(STREAM_P, __popc(cand)) + pack_counts(STREAM_S, __popc(os.x));
        }
        elS = __reduce_or_sync(gmask, os.e & ~os.x);
    }
    counts = __reduce_add_sync(gmask, counts);
    // every lane of the item stores the same words to the same addresses
    if (on) {
        *q.mp = make_meta((int)(off - off_in), 32 - __clz((int)elS), 32 - __clz((int)elA), 0, off_in);
        *q.cp = counts;
    }
    q.off = off;
    ++q.mp; ++q.cp;
    tight_rotate(t, xn);
    // (whole-warp items: the word is the same in every lane, which the compiler cannot see through the load; the broadcast
    // makes the loop condition provably uniform, so the warp-level operations of the next step need no convergence guards)
    q.vprev = q.vcur; q.vcur = WHOLE ? __shfl_sync(FULL, o.w, 0) : o.w;
    ++q.w;
}

template <int BW, int TIER>
__device__ __noinline__ void tight_run(TightIO* io) {
    TightLoop q;
    q.c = io->c;
    q.cw = io->cw; q.raw = io->raw;
    q.off = io->off; q.cap_end = io->cap_end;
    q.w = io->w; q.wend = io->wend; q.L = io->L;
    Meta* meta = io->meta;
    uint32_t* bcnt = io->bcnt;
    if (BW == 32) {  // uniform by construction; say so (see tight_step): the values can then live in uniform registers
        q.w = __shfl_sync(0xFFFFFFFFu, q.w, 0); q.wend = __shfl_sync(0xFFFFFFFFu, q.wend, 0); q.L = __shfl_sync(0xFFFFFFFFu, q.L, 0);
        q.off = __shfl_sync(0xFFFFFFFFu, q.off, 0); q.cap_end = __shfl_sync(0xFFFFFFFFu, q.cap_end, 0);
        q.cw = uniform_ptr(q.cw); q.raw = uniform_ptr(q.raw); meta = uniform_ptr(meta); bcnt = uniform_ptr(bcnt);
    }
    q.on = BW == 32 ? true : io->on != 0;
    q.po = reinterpret_cast<const uint4*>(q.cw + (q.w + 1));
    q.pb = reinterpret_cast<const uint2*>(q.cw + (q.w + 1 + (q.c.s >> 5) + 1));
    q.mp = meta + q.w;
    q.cp = bcnt + q.w;
    q.vprev = q.on ? v_eff(q.cw, q.w - 1) : 0xFFFFFFFFu; q.vcur = q.on ? q.cw[q.w].v : 0xFFFFFFFFu;
    if (BW == 32) { q.vprev = __shfl_sync(0xFFFFFFFFu, q.vprev, 0); q.vcur = __shfl_sync(0xFFFFFFFFu, q.vcur, 0); }
    // flags of the two words in front: unknown, so the first steps take the exact anchors; t.lenL is the general path's
    // carried run length
    q.flags = TF_SUSC | TF_SUSP | TF_PREV_RARE;
    q.lt = io->lt;
    q.o = __ldg(q.po); q.bb = __ldg(q.pb);
    q.zc = 0;                      // consecutive words without a passing substitution window or perfect-run candidate in any lane (saturating)
    TightState ta = io->t, tb;
    // two words per trip with the roles of the two state objects swapped: the carried words of a step are born in other
    // registers than the ones they replace, so no register moves are needed at the loop end
#pragma unroll 1
    for (;;) {
        if (tight_leave_now<BW, TIER>(q)) { tight_finish<TIER>(io, q, ta); return; }
        tight_step<BW, TIER>(q, ta, tb);
        if (tight_leave_now<BW, TIER>(q)) { tight_finish<TIER>(io, q, tb); return; }
        tight_step<BW, TIER>(q, tb, ta);
    }
}

template <int BW>
#ifndef RB_SCAN_BLOCKS_PER_SM
#define RB_SCAN_BLOCKS_PER_SM 16
#endif
__global__ void __launch_bounds__(SCAN_WARPS * 32, RB_SCAN_BLOCKS_PER_SM) scan_kernel(DevBatch b) {
    constexpr int GROUPS = 32 / BW;
    __shared__ int s_cnt[SCAN_WARPS][GROUPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / BW, j = lane % BW;
    const unsigned gmask = (BW == 32) ? 0xFFFFFFFFu : (((1u << BW) - 1u) << (g * BW));
    // launch order: band-major (all chunks of band 0, then band 1, ...), so that the warps resident on an SM run the same
    // variant of the tight loop most of the time (instruction cache) and the costliest band (small shifts) starts first;
    // the item tables are chunk-major (item = chunk * nbands + band)
    const int nbands = b.lay.nbands;
    const long long slot = ((long long)blockIdx.x * SCAN_WARPS + warp) * GROUPS + g;
    const long long item = nbands == 1 ? slot : (slot % b.n_chunks) * nbands + slot / b.n_chunks;
    bool active = slot < b.n_items;
    if (b.item_clk && active && j == 0) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        b.item_clk[4 * item] = (long long)t0;
    }
    const int band = active ? (int)(item % nbands) : 0;
    Chunk ch;
    ch.contig = 0; ch.w0 = 0; ch.w1 = 0; ch.last = 0;
    if (active) ch = b.chunks[item / nbands];
    const Contig cg = b.contigs[ch.contig];
    const PlaneWord* __restrict__ cw = b.planes + cg.word_base;
    const int L = cg.L;
    LaneCfg cfg = band_lane_cfg(b.lay, band, j);
    lane_cfg_set_contig(cfg, L);
    Meta* __restrict__ meta = b.meta + (long long)band * b.n_buckets + b.bucket_base[ch.contig];
    uint32_t* __restrict__ bcnt = b.bcnt + (long long)band * b.n_buckets + b.bucket_base[ch.contig];
    int* cnt = &s_cnt[warp][g];

    GpuSink sk;
    sk.base = b.raw + (active ? b.item_base[item] : 0);
    sk.cap = active ? b.item_cap[item] : 0;
    sk.cnt = cnt;
    const uint32_t off0 = active ? (uint32_t)b.item_base[item] : 0u;

    // a chunk that lies strictly inside an N run (the word before it and all its words are N) emits nothing:
    // no window is evaluated and every run was closed by the first N of the run
    {
        bool alln = active && ch.w0 > 0 && full_n(cw, ch.w0 - 1) && !ch.last && !(b.debug & 2);
        for (int base = 0; __any_sync(0xFFFFFFFFu, alln && ch.w0 + base < ch.w1); base += BW) {  // warp-uniform trip count
            const int wq = ch.w0 + base + j;
            const bool f = !alln || wq >= ch.w1 || full_n(cw, wq);
            const unsigned bits = __ballot_sync(0xFFFFFFFFu, f) & gmask;
            if (bits != gmask) alln = false;
        }
        if (alln) {
            for (int wq = ch.w0 + j; wq < ch.w1; wq += BW) { meta[wq] = make_meta(0, 0, 0, 1, off0); bcnt[wq] = 0u; }
            if (j == 0) b.item_count[item] = 0;
            active = false;
        }
    }
    // first word of the run of full-N words that ends right before the chunk: backward probe over the full-N bit mask of
    // the plane words (pack_kernel), 32 words per load, one load per lane and trip
    int nb = ch.w0;
    {
        bool go = active && ch.w0 > 0 && full_n(cw, ch.w0 - 1);
        long long p = cg.word_base + ch.w0 - 1;  // plane index of the word in front of the chunk
        if (go) {
            // the mask word that holds p: the ones at and below bit (p & 31)
            const int bit = (int)(p & 31);
            const uint32_t z = ~b.nmask[p >> 5] & lowmask(bit + 1);
            const int k = z ? bit - (31 - __clz((int)z)) : bit + 1;
            nb -= k; p -= k;
            if (z) go = false;
        }
        while (__any_sync(0xFFFFFFFFu, go)) {
            // p is the last word of a mask word now; lane j looks at the j-th mask word further back
            const long long gq = (p >> 5) - j;
            const uint32_t m = (go && gq >= 0) ? b.nmask[gq] : 0u;
            const unsigned bits = (__ballot_sync(0xFFFFFFFFu, m == 0xFFFFFFFFu) & gmask) >> (g * BW);
            const unsigned inv = ~bits & (BW == 32 ? 0xFFFFFFFFu : ((1u << BW) - 1u));
            const int c = inv ? __ffs((int)inv) - 1 : BW;  // mask words that are all ones
            const uint32_t mc = __shfl_sync(0xFFFFFFFFu, m, (g * BW + min(c, BW - 1)) & 31);  // the first one that is not
            if (go) {
                nb -= 32 * c; p -= 32ll * c;
                if (c < BW) {
                    nb -= __clz((int)~mc);  // its ones from the top
                    go = false;
                }
            }
        }
        if (nb < 0) nb = 0;  // the guard word in front of the contig is N as well
    }
    const int nb0 = nb;
    LaneState st;
    int H = b.warm0;
    const int e0 = ch.w0;  // first emitting word
    int we = e0;           // word at which the lane state must be complete (end of the current warm-up)
    int force = 0;         // the warm-up rebuilds the reference machines: process its words bit-serially
    int q = warmup_start(we, nb, H);
    int Ha = warmup_anchor_words(q, H);
    lane_init(cfg, st, cw, q);
    int w = q;
    // (the keep filter's look-back is trusted after four fast words in a row; at the contig start there is nothing behind)
    int prev_slow = 1, fastrun = q == 0 ? 3 : 0;
    uint32_t off = off0;  // raw-pool index of the next bucket's first slot
    if (j == 0) *cnt = 0;
    __syncwarp();
    int restarts = 0;
    int dbg_gen = 0, dbg_slow = 0;  // diagnostics: words through the general path / of them bit-serial
    if (active && ch.w0 >= ch.w1) {  // empty contig: only the tail bucket
        active = false;
        if (ch.last) {
            if (j == 0) { meta[ch.w1] = make_meta(0, 0, 0, 1, off0); bcnt[ch.w1] = 0u; }
        }
        if (j == 0) b.item_count[item] = 0;
    }

    const int guard = b.lay.guard;
    const int tier = tight_tier(band_m0(b.lay, band) - 2);  // by the item's smallest shift; uniform over the warp (see launch_scan)
    // The neighbour anchors of lanes at the edge of an item need no masking: lanes 0 and 1 of an item are halo
    // shifts (never motif lanes) and a motif lane has j + 2 <= mpb + 3 < BW (layout.h), so every value a motif lane
    // reads comes from its own item.
    for (;;) {
        // ---- end of the chunk: tail flush (last chunk of a contig), slot count ---------------------------------
        if (active && w >= ch.w1) {
            if (ch.last) {
                sk.reset();
                lane_tail(sk, cfg, st, L);
                const int ns = __reduce_add_sync(gmask, sk.nslots);
                const int nP = __reduce_add_sync(gmask, sk.n[0]), nS = __reduce_add_sync(gmask, sk.n[1]), nA = __reduce_add_sync(gmask, sk.n[2]);
                if (j == 0) {
                    meta[ch.w1] = make_meta(min(ns, META_MAX_SLOTS), 0, 0, 1, off);
                    bcnt[ch.w1] = pack_counts(STREAM_P, nP) + pack_counts(STREAM_S, nS) + pack_counts(STREAM_A, nA);
                    if (ns > META_MAX_SLOTS || counts_overflow(nP, nS, nA)) atomicAdd(b.counters + 2, 1);
                }
            }
            __syncwarp(gmask);
            if (j == 0) {
                const int n = *cnt;
                if (b.item_clk) {
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    b.item_clk[4 * item + 1] = (long long)t1;
                    b.item_clk[4 * item + 2] = dbg_gen;
                    b.item_clk[4 * item + 3] = ((long long)restarts << 32) | (unsigned)dbg_slow;
                }
                b.item_count[item] = n;
                if (n > sk.cap) atomicAdd(b.counters + 0, 1);
                if (restarts) atomicAdd(b.counters + 1, restarts);
            }
            active = false;
        }
        if (!__any_sync(0xFFFFFFFFu, active)) break;

        // ---- general path: one word (warm-up, slow words, restarts, N-run jumps) --------------------------------
        const int lim = max(we, e0) - 2;  // no word is emitted before max(we, e0)
        if (active && w >= q + Ha && w - 1 >= nb0) {
            // inside the N run that ends right before the chunk (words nb0 .. e0-1 are all N): one jump
            const int k = min(e0 - guard, lim) - w;
            if (k > 0) {
                w += k;
                lane_skip(cfg, st, cw, w, k);
            }
        }
        const bool can = active && w >= q + Ha && w < lim;
        if (__any_sync(0xFFFFFFFFu, can)) {
            // warming up inside an N run: jump over words whose neighbourhood is all N (scan_core.h, lane_skip)
            const int wq = w - 1 + j;
            const bool f = can && wq < cg.nw + guard && full_n(cw, wq);
            const unsigned bits = (__ballot_sync(0xFFFFFFFFu, f) & gmask) >> (g * BW);
            const unsigned inv = ~bits & (BW == 32 ? 0xFFFFFFFFu : ((1u << BW) - 1u));
            const int r = inv ? __ffs((int)inv) - 1 : BW;
            const int k = can ? min(r - guard - 1, lim - w) : 0;
            if (k > 0) {
                w += k;
                lane_skip(cfg, st, cw, w, k);
            }
        }
        int slow = active ? ((force && w < we) || !word_is_fast(cw, w, cg.nw)) : 1;
        {
            // fast -> slow: the reference machines' state comes from the bit-parallel view of the previous word when every
            // lane's word holds nine failing windows in a row (lane_to_slow), else it is rebuilt by a warm-up that ends here
            const bool trans = active && slow && !prev_slow;
            SlowEntry se;
            se.S = st.S; se.A = st.A; se.pst = st.pst;
            bool conv_ok = !trans || w >= q + Ha;  // the machines run in this word (not an anchors-only warm-up word)
            if (trans && conv_ok && cfg.motif) conv_ok = slow_entry(st.es.P, st.es.r8, st.ea.P, st.ea.r8, st.x_prev, cw, w, cfg.s, se);
            const unsigned conv_fail = __ballot_sync(0xFFFFFFFFu, !conv_ok) & gmask;
            if (trans) {
                if (!conv_fail) {
                    lane_enter_slow(st, se);
                    prev_slow = 1;
                } else {
                    we = w; nb = w; H = b.warm0; force = 1;
                    q = warmup_start(we, nb, H);
                    Ha = warmup_anchor_words(q, H);
                    lane_init(cfg, st, cw, q);
                    w = q;
                    prev_slow = 1; fastrun = q == 0 ? 3 : 0;
                    slow = 1;
                }
            }
        }
        uint32_t a = 0u;
        if (active) a = slow ? lane_phase1(cfg, st, cw, w, L) : lane_phase1_fast(cfg, st, cw, w, L);
        fastrun = slow ? 0 : min(fastrun + 1, 4);
        bool bad = false;
        if (active && q > 0) {
            if (w == q + Ha - 2 && cfg.s && !(st.sync & SYNC_X)) bad = true;
            if (w == we && cfg.motif && (st.sync & SYNC_ALL) != SYNC_ALL) bad = true;
        }
        const unsigned badmask = __ballot_sync(0xFFFFFFFFu, bad) & gmask;
        // anchor words of the neighbouring shifts (same item): lanes j-2, j-1, j+1, j+2
        uint32_t a_m2 = __shfl_up_sync(0xFFFFFFFFu, a, 2), a_m1 = __shfl_up_sync(0xFFFFFFFFu, a, 1);
        uint32_t a_p1 = __shfl_down_sync(0xFFFFFFFFu, a, 1), a_p2 = __shfl_down_sync(0xFFFFFFFFu, a, 2);
        if (badmask) {
            // the warm-up did not reach a history-free state: start earlier
            H = min(H * 4, we);
            if (H < 1) H = 1;
            q = warmup_start(we, nb, H);
            Ha = warmup_anchor_words(q, H);
            lane_init(cfg, st, cw, q);
            w = q;
            prev_slow = 1; fastrun = q == 0 ? 3 : 0;
            ++restarts;
        } else if (active) {
            IterCtx it;
            it.w = w; it.L = L; it.emit_on = w >= we && w >= e0; it.slow = slow; it.prev_slow = prev_slow; it.fastrun = fastrun;
            sk.reset();
            lane_phase2(sk, cfg, st, cw, it, a_m2, a_m1, a_p1, a_p2, w >= q + Ha);
            if (it.emit_on) {
                const int ns = __reduce_add_sync(gmask, sk.nslots);
                const int nP = __reduce_add_sync(gmask, sk.n[0]), nS = __reduce_add_sync(gmask, sk.n[1]), nA = __reduce_add_sync(gmask, sk.n[2]);
                const int dS = __reduce_max_sync(gmask, sk.dS), dA = __reduce_max_sync(gmask, sk.dA);
                if (j == 0) {
                    meta[w] = make_meta(min(ns, META_MAX_SLOTS), dS, dA, it.slow, off);
                    bcnt[w] = pack_counts(STREAM_P, nP) + pack_counts(STREAM_S, nS) + pack_counts(STREAM_A, nA);
                    if (ns > META_MAX_SLOTS || counts_overflow(nP, nS, nA)) atomicAdd(b.counters + 2, 1);
                }
                off += (uint32_t)ns;
            }
            prev_slow = slow;
            ++w;
            ++dbg_gen; dbg_slow += slow;
        }

        // ---- tight path: consecutive fast, emitting words ------------------------------------------------------------
        // word `we` (state check) and the first fast words of a stretch (keep filter not yet trusted) go through the
        // general path; the tight loop stops before the contig's tail zone (the anchor view differs from X_s there)
        const bool tight_ok = active && !badmask && w > we && w >= e0 && fastrun >= 4 && !prev_slow && !(b.debug & 1);
        const int wend = min(min(ch.w1, cg.nw - 1), ((L - b.lay.s_hi) >> 5) - 1);
        const bool want = tight_ok && w < wend && (v_eff(cw, w - 1) & cw[w].v) == 0xFFFFFFFFu;
        // every group with work must be able to enter, else the warp keeps to the general path for this word
        if (__any_sync(0xFFFFFFFFu, want) && __all_sync(0xFFFFFFFFu, want || !active)) {
            TightIO io;
            tight_enter(cfg, st, io.t);
            io.c = make_tight_cfg(cfg, tier == TIER_LARGE);
            io.cw = cw; io.meta = meta; io.bcnt = bcnt; io.raw = b.raw; io.off = off; io.cap_end = off0 + (uint32_t)sk.cap;
            io.w = w; io.wend = wend; io.L = L; io.on = want ? 1 : 0;
            io.lt = (1u << lane) - 1u;
            if (tier == TIER_SMALL) tight_run<BW, TIER_SMALL>(&io);
            else if (tier == TIER_MID) tight_run<BW, TIER_MID>(&io);
            else tight_run<BW, TIER_LARGE>(&io);
            if (want) {
                tight_leave(io.t, st);
                if (io.w != w && cfg.motif) smear_from_last(cfg, st, io.w);  // (scan_tight.h tight_events_A)
                if (io.w != w && cfg.s) st.xc.idx = io.w + (cfg.s >> 5) + 1;  // the cached plane word: the last b operand
                w = io.w;
                off = io.off;
                if (j == 0) *cnt = (int)(off - off0);
            }
            __syncwarp();
        }
    }
}

void launch_scan(const DevBatch& b, cudaStream_t st) {
    if (b.n_items == 0) return;
    const int groups = b.lay.groups;
    const long long warps = (b.n_items + groups - 1) / groups;
    const unsigned blocks = (unsigned)((warps + SCAN_WARPS - 1) / SCAN_WARPS);
    if (b.lay.bw == 32) scan_kernel<32><<<blocks, SCAN_WARPS * 32, 0, st>>>(b);
    else if (b.lay.bw == 16) scan_kernel<16><<<blocks, SCAN_WARPS * 32, 0, st>>>(b);
    else scan_kernel<8><<<blocks, SCAN_WARPS * 32, 0, st>>>(b);
}

// ---------------------------------------------------------------------------------------------------------------
// K6: ordered compaction. One thread per bucket (contig word, plus one tail bucket per contig).
//   M1 merge_count_kernel   per-bucket stream sizes -> per-block sums and elided-end maxima
//   M2 merge_scan_kernel    exclusive prefix over the blocks (one block)
//   M3 merge_write_kernel   per-bucket offsets, key ranking, final records
// ---------------------------------------------------------------------------------------------------------------
struct BucketInfo {
    int c, w;
    unsigned n[3];                  // final records per stream (pseudo included)
    unsigned pseudo[3];
    unsigned long long emax[2];     // contig << 32 | elided_end_code
};

__device__ __forceinline__ BucketInfo bucket_info(const DevBatch& b, long long gb) {
    BucketInfo bi;
    bi.c = find_segment(b.bucket_base, b.n_contigs, gb);
    bi.w = (int)(gb - b.bucket_base[bi.c]);
    unsigned tot[3] = {0u, 0u, 0u};
    int slow = 0, dS = 0, dA = 0;
    for (int k = 0; k < b.lay.nbands; ++k) {
        const Meta m = b.meta[(long long)k * b.n_buckets + gb];
        const uint32_t cn = b.bcnt[(long long)k * b.n_buckets + gb];
        tot[0] += counts_of(cn, 0); tot[1] += counts_of(cn, 1); tot[2] += counts_of(cn, 2);
        slow |= meta_slow(m);
        dS = max(dS, meta_dmax(m, STREAM_S));
        dA = max(dA, meta_dmax(m, STREAM_A));
    }
    for (int s = 0; s < 3; ++s) {
        bi.pseudo[s] = bucket_has_pseudo(s, slow, (int)tot[s]) ? 1u : 0u;
        bi.n[s] = tot[s] + bi.pseudo[s];
    }
    const uint32_t eS = elided_end_code(bi.w, dS), eA = elided_end_code(bi.w, dA);
    bi.emax[0] = eS ? (((unsigned long long)bi.c << 32) | eS) : 0ull;
    bi.emax[1] = eA ? (((unsigned long long)bi.c << 32) | eA) : 0ull;
    return bi;
}

__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// exclusive block scan of (3 sums, 2 maxima); returns the block totals in tot
__device__ __forceinline__ void block_scan(const unsigned (&n)[3], const unsigned long long (&e)[2], unsigned (&xs)[3],
                                           unsigned long long (&xe)[2], unsigned (&tot)[3], unsigned long long (&tote)[2]) {
    __shared__ unsigned s_sum[3][MERGE_BLOCK / 32];
    __shared__ unsigned long long s_max[2][MERGE_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc[3];
    unsigned long long ince[2];
    for (int s = 0; s < 3; ++s) {
        unsigned v = n[s];
        for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= d) v += t; }
        inc[s] = v;
        if (lane == 31) s_sum[s][warp] = v;
    }
    for (int s = 0; s < 2; ++s) {
        unsigned long long v = e[s];
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= d) v = umax64(v, t); }
        ince[s] = v;
        if (lane == 31) s_max[s][warp] = v;
    }
    __syncthreads();
    for (int s = 0; s < 3; ++s) {
        unsigned pre = 0u, all = 0u;
        for (int k = 0; k < MERGE_BLOCK / 32; ++k) { const unsigned v = s_sum[s][k]; if (k < warp) pre += v; all += v; }
        xs[s] = pre + inc[s] - n[s];
        tot[s] = all;
    }
    for (int s = 0; s < 2; ++s) {
        unsigned long long pre = 0ull, all = 0ull;
        for (int k = 0; k < MERGE_BLOCK / 32; ++k) { const unsigned long long v = s_max[s][k]; if (k < warp) pre = umax64(pre, v); all = umax64(all, v); }
        const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, ince[s], 1);
        xe[s] = umax64(pre, lane ? up : 0ull);
        tote[s] = all;
    }
}

__global__ void __launch_bounds__(MERGE_BLOCK) merge_count_kernel(DevBatch b) {
    const long long gb = b.gb_first + (long long)blockIdx.x * MERGE_BLOCK + threadIdx.x;
    unsigned n[3] = {0u, 0u, 0u};
    unsigned long long e[2] = {0ull, 0ull};
    if (gb < b.gb_first + b.n_active) {
        const BucketInfo bi = bucket_info(b, gb);
        for (int s = 0; s < 3; ++s) n[s] = bi.n[s];
        e[0] = bi.emax[0]; e[1] = bi.emax[1];
    }
    unsigned xs[3], tot[3];
    unsigned long long xe[2], tote[2];
    block_scan(n, e, xs, xe, tot, tote);
    if (threadIdx.x == 0) {
        BlockPartial p;
        for (int s = 0; s < 3; ++s) p.sum[s] = tot[s];
        p.emax[0] = tote[0]; p.emax[1] = tote[1];
        b.partial[blockIdx.x] = p;
    }
}

// Exclusive prefix (three sums, two maxima) over segments of `seg` partials, one block per segment: every thread owns a
// contiguous range (local prefix), one block-wide scan of the per-thread totals, then the exclusive prefixes are written
// back in place; seg_total[segment] receives the segment's total (and `totals` the three sums when given).
// Two levels: the per-block partials in segments of MERGE_SEG, then the segment totals in one block.
__global__ void __launch_bounds__(1024) merge_scan_kernel(BlockPartial* __restrict__ arr, int n, int seg, BlockPartial* __restrict__ seg_total,
                                                          long long* __restrict__ totals) {
    __shared__ unsigned long long s_w[5][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = blockIdx.x * seg, s1 = min(n, s0 + seg);
    const int per = (seg + 1023) / 1024;
    const int i0 = min(s1, s0 + (int)threadIdx.x * per), i1 = min(s1, i0 + per);
    unsigned long long v[5] = {0ull, 0ull, 0ull, 0ull, 0ull};
    for (int i = i0; i < i1; ++i) {
        const BlockPartial p = arr[i];
        v[0] += p.sum[0]; v[1] += p.sum[1]; v[2] += p.sum[2];
        v[3] = umax64(v[3], p.emax[0]); v[4] = umax64(v[4], p.emax[1]);
    }
    unsigned long long ex[5], all[5];
    for (int k = 0; k < 5; ++k) {
        unsigned long long x = v[k];
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x = (k < 3) ? x + t : umax64(x, t);
        }
        if (lane == 31) s_w[k][warp] = x;
        const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, x, 1);
        ex[k] = (k < 3) ? x - v[k] : (lane ? up : 0ull);  // exclusive inside the warp
    }
    __syncthreads();
    for (int k = 0; k < 5; ++k) {
        unsigned long long pre = 0ull, tot = 0ull;
        for (int u = 0; u < 32; ++u) {
            const unsigned long long x = s_w[k][u];
            if (k < 3) { if (u < warp) pre += x; tot += x; }
            else { if (u < warp) pre = umax64(pre, x); tot = umax64(tot, x); }
        }
        ex[k] = (k < 3) ? ex[k] + pre : umax64(ex[k], pre);
        all[k] = tot;
    }
    for (int i = i0; i < i1; ++i) {  // exclusive prefix of every partial of this thread's range
        const BlockPartial p = arr[i];
        BlockPartial o;
        o.sum[0] = ex[0]; o.sum[1] = ex[1]; o.sum[2] = ex[2]; o.emax[0] = ex[3]; o.emax[1] = ex[4];
        arr[i] = o;
        ex[0] += p.sum[0]; ex[1] += p.sum[1]; ex[2] += p.sum[2];
        ex[3] = umax64(ex[3], p.emax[0]); ex[4] = umax64(ex[4], p.emax[1]);
    }
    if (threadIdx.x == 0) {
        BlockPartial p;
        p.sum[0] = all[0]; p.sum[1] = all[1]; p.sum[2] = all[2]; p.emax[0] = all[3]; p.emax[1] = all[4];
        seg_total[blockIdx.x] = p;
        if (totals) { totals[0] = (long long)all[0]; totals[1] = (long long)all[1]; totals[2] = (long long)all[2]; }
    }
}
// the partial of merge block `blk` after both levels: its prefix inside the segment combined with the segment's prefix
__device__ __forceinline__ BlockPartial merge_block_prefix(const DevBatch& b, int blk) {
    BlockPartial p = b.partial[blk];
    const BlockPartial g = b.partial[b.n_merge_blocks + 1 + blk / MERGE_SEG];
    for (int s = 0; s < 3; ++s) p.sum[s] += g.sum[s];
    p.emax[0] = umax64(p.emax[0], g.emax[0]); p.emax[1] = umax64(p.emax[1], g.emax[1]);
    return p;
}

// Phase 1 (thread = bucket): stream offsets of every bucket of the block, pseudo records, per-contig offsets.
// Phase 2 (thread = slot): the block's raw slots are spread evenly over the threads; each finds its bucket in the shared
// prefix table; every candidate of the slot (one for a record, the set bits of a mask entry) ranks itself among the
// bucket's candidates of its stream by (time, mlen, seq) and is written in its final form.
__global__ void __launch_bounds__(MERGE_BLOCK) merge_write_kernel(DevBatch b) {
    __shared__ uint32_t s_off[8][MERGE_BLOCK];          // first raw slot of (band, bucket)
    __shared__ unsigned short s_n[8][MERGE_BLOCK];      // raw slots of (band, bucket)
    __shared__ uint32_t s_pre[MERGE_BLOCK + 1];         // exclusive prefix of the buckets' raw slot counts
    __shared__ uint32_t s_dst[3][MERGE_BLOCK];          // first final record of (stream, bucket), relative to the block,
                                                        // pseudo record already discounted
    __shared__ int s_w[MERGE_BLOCK];
    __shared__ int s_c[MERGE_BLOCK];
    __shared__ uint32_t s_wsum[MERGE_BLOCK / 32];
    __shared__ uint32_t s_x[MERGE_STAGE], s_z[MERGE_STAGE];
    __shared__ unsigned char s_bk_of[MERGE_STAGE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long gb = b.gb_first + (long long)blockIdx.x * MERGE_BLOCK + tid;
    const bool live = gb < b.gb_first + b.n_active;
    const int nbands = b.lay.nbands;
    BucketInfo bi;
    bi.c = 0; bi.w = 0;
    for (int s = 0; s < 3; ++s) { bi.n[s] = 0u; bi.pseudo[s] = 0u; }
    bi.emax[0] = bi.emax[1] = 0ull;
    uint32_t nrec = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s_off[k][tid] = 0u; s_n[k][tid] = 0; }
    if (live) {
        bi = bucket_info(b, gb);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < nbands) {
                const Meta m = b.meta[(long long)k * b.n_buckets + gb];
                const int n = meta_slots(m);
                s_off[k][tid] = m.off;
                s_n[k][tid] = (unsigned short)n;
                nrec += (uint32_t)n;
            }
    }
    unsigned xs[3], tot[3];
    unsigned long long xe[2], tote[2];
    block_scan(bi.n, bi.emax, xs, xe, tot, tote);
    const BlockPartial bp = merge_block_prefix(b, (int)blockIdx.x);
    const BlockPartial total = b.partial[b.n_merge_blocks];
    const long long sbase[3] = {0ll, (long long)total.sum[0], (long long)(total.sum[0] + total.sum[1])};
    // block prefix of the raw slot counts
    {
        uint32_t v = nrec;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= d) v += t; }
        if (lane == 31) s_wsum[warp] = v;
        __syncthreads();
        uint32_t pre = 0u;
        for (int k = 0; k < warp; ++k) pre += s_wsum[k];
        s_pre[tid] = pre + v - nrec;
        if (tid == MERGE_BLOCK - 1) s_pre[MERGE_BLOCK] = pre + v;
    }
    s_w[tid] = bi.w;
    s_c[tid] = bi.c;
    for (int s = 0; s < 3; ++s) s_dst[s][tid] = xs[s] + bi.pseudo[s];
    if (live) {
        long long o[3];
        for (int s = 0; s < 3; ++s) o[s] = (long long)bp.sum[s] + xs[s];
        if (bi.w == 0 || gb == b.gb_first)  // gb_first > 0: a word range of a single contig (rb_set_word_range)
            for (int s = 0; s < 3; ++s) b.contig_off[(long long)s * (b.n_contigs + 1) + bi.c] = o[s];
        if (gb == b.gb_first + b.n_active - 1)
            for (int s = 0; s < 3; ++s) b.contig_off[(long long)s * (b.n_contigs + 1) + b.n_contigs] = o[s] + bi.n[s];
        for (int s = 1; s < 3; ++s)
            if (bi.pseudo[s]) {
                const unsigned long long e = umax64(bp.emax[s - 1], xe[s - 1]);
                long long ee = -1;
                if (e != 0ull && (int)(e >> 32) == bi.c) ee = (long long)(e & 0xFFFFFFFFull) - 1;
                if (sbase[s] + o[s] < b.dst_cap) b.dst[sbase[s] + o[s]] = pseudo_rec(bi.w, ee);
            }
    }
    __syncthreads();

    // The slots of the block are staged in shared memory once (mask / key word, mflags, bucket): the rank loops then run
    // on shared memory. A block with more slots than fit ranks from global memory.
    const uint32_t T = s_pre[MERGE_BLOCK];
    const bool staged = T <= MERGE_STAGE;
    for (uint32_t r = tid; r < T; r += MERGE_BLOCK) {
        int lo = 0, hi = MERGE_BLOCK - 1;  // largest bucket with s_pre[bucket] <= r
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_pre[mid] <= r) lo = mid; else hi = mid - 1;
        }
        const int bk = lo;
        uint32_t idx = r - s_pre[bk];
        int band = 0;
        while (idx >= s_n[band][bk]) { idx -= s_n[band][bk]; ++band; }
        const int4 v = *reinterpret_cast<const int4*>(b.raw + s_off[band][bk] + idx);
        if (staged) {
            // what ranking needs of a slot: the kept mask (entry) or the key (record), and mflags
            s_x[r] = ((v.z >> 16) & REC_ENTRY) ? (uint32_t)v.x : (uint32_t)v.w;
            s_z[r] = (uint32_t)v.z;
        }
        s_bk_of[r < MERGE_STAGE ? r : 0] = (unsigned char)bk;
        if (!staged) {
            // (rare) rank from global memory
            const int st = (v.z >> REC_STREAM_SHIFT) & 3;
            const bool is_entry = ((v.z >> 16) & REC_ENTRY) != 0;
            const int w = s_w[bk], mlen = v.z & 0xFFFF;
            const long long dbase = sbase[st] + (long long)bp.sum[st] + s_dst[st][bk];
            const uint32_t mk = (uint32_t)mlen << 2;
            for (uint32_t bits = is_entry ? (uint32_t)v.x : 1u; bits; bits &= bits - 1u) {
                const int i = __ffs((int)bits) - 1;
                const uint32_t key = is_entry ? (((uint32_t)i << 18) | mk) : (uint32_t)v.w;
                const int ki = (int)(key >> 18);
                const uint32_t lowkey = key & 0x3FFFFu;
                int rank = 0;
                for (int k = 0; k < nbands; ++k) {
                    const int n = s_n[k][bk];
                    const int4* q = reinterpret_cast<const int4*>(b.raw + s_off[k][bk]);
                    for (int u = 0; u < n; ++u) {
                        const int4 ov = q[u];
                        if (((ov.z >> REC_STREAM_SHIFT) & 3) != st) continue;
                        if ((ov.z >> 16) & REC_ENTRY) rank += __popc((uint32_t)ov.x & lowmask(ki + ((((uint32_t)(ov.z & 0xFFFF)) << 2) < lowkey ? 1 : 0)));
                        else rank += (uint32_t)ov.w < key ? 1 : 0;
                    }
                }
                int4 out;
                if (is_entry) {
                    Rec e;
                    e.start = v.x; e.end = v.y; e.mflags = v.z; e.key = v.w;
                    int s0, e0;
                    entry_interval(e, w, b.planes + b.contigs[s_c[bk]].word_base, i, s0, e0);
                    out = make_int4(s0, e0, mlen, 32 * w + i);
                } else {
                    out = make_int4(v.x, v.y, v.z & ((1 << REC_STREAM_SHIFT) - 1), 32 * w + (v.w >> 18));
                }
                const long long at = dbase + rank;
                if (at < b.dst_cap) *reinterpret_cast<int4*>(b.dst + at) = out;
            }
        }
    }
    if (!staged) return;
    __syncthreads();
    for (uint32_t r = tid; r < T; r += MERGE_BLOCK) {
        const int bk = s_bk_of[r];
        uint32_t idx = r - s_pre[bk];
        int band = 0;
        while (idx >= s_n[band][bk]) { idx -= s_n[band][bk]; ++band; }
        const uint32_t z = s_z[r];
        const int st = (int)(z >> REC_STREAM_SHIFT) & 3;
        const bool is_entry = ((z >> 16) & REC_ENTRY) != 0;
        const int w = s_w[bk], mlen = (int)(z & 0xFFFF);
        const long long dbase = sbase[st] + (long long)bp.sum[st] + s_dst[st][bk];
        const uint32_t mk = (uint32_t)mlen << 2;
        const uint32_t r0 = s_pre[bk], r1 = s_pre[bk + 1];
        int4 v = make_int4(0, 0, 0, 0);
        bool have = false;
        for (uint32_t bits = is_entry ? s_x[r] : 1u; bits; bits &= bits - 1u) {
            const int i = __ffs((int)bits) - 1;
            const uint32_t key = is_entry ? (((uint32_t)i << 18) | mk) : s_x[r];
            const int ki = (int)(key >> 18);
            const uint32_t lowkey = key & 0x3FFFFu;
            int rank = 0;
            for (uint32_t u = r0; u < r1; ++u) {
                const uint32_t oz = s_z[u];
                if (((oz >> REC_STREAM_SHIFT) & 3) != (uint32_t)st) continue;
                if ((oz >> 16) & REC_ENTRY) rank += __popc(s_x[u] & lowmask(ki + (((oz & 0xFFFFu) << 2) < lowkey ? 1 : 0)));
                else rank += s_x[u] < key ? 1 : 0;
            }
            if (!have) { v = *reinterpret_cast<const int4*>(b.raw + s_off[band][bk] + idx); have = true; }  // the other fields
            int4 out;
            if (is_entry) {
                Rec e;
                e.start = v.x; e.end = v.y; e.mflags = v.z; e.key = v.w;
                int s0, e0;
                entry_interval(e, w, b.planes + b.contigs[s_c[bk]].word_base, i, s0, e0);
                out = make_int4(s0, e0, mlen, 32 * w + i);
            } else {
                out = make_int4(v.x, v.y, v.z & ((1 << REC_STREAM_SHIFT) - 1), 32 * w + (v.w >> 18));
            }
            const long long at = dbase + rank;
            if (at < b.dst_cap) *reinterpret_cast<int4*>(b.dst + at) = out;
        }
    }
}

// 16-byte records -> 8-byte records (rb_rec8); candidates of 65535 positions or more go to a side list.
// stream s occupies [sbase[s], sbase[s+1]) of both pools; long_cnt[s] counts the side-list entries of stream s.
__global__ void __launch_bounds__(256) compact_kernel(const Rec* __restrict__ src, long long n, long long b1, long long b2,
                                                      uint2* __restrict__ dst, long long* __restrict__ long_list, int long_cap,
                                                      int* __restrict__ long_cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 r = *reinterpret_cast<const int4*>(src + i);
    const int mlen = r.z & 0xFFFF, flags = (r.z >> 16) & 0xF;
    uint2 o;
    if (flags & REC_PSEUDO) {
        o.x = (uint32_t)r.y;
        o.y = (uint32_t)(mlen | (flags << 12)) << 16;
    } else {
        const int len = r.y - r.x;
        o.x = (uint32_t)r.x;
        o.y = (uint32_t)(len >= 0xFFFF ? 0xFFFF : len) | ((uint32_t)(mlen | (flags << 12)) << 16);
        if (len >= 0xFFFF) {
            const int s = i >= b2 ? 2 : (i >= b1 ? 1 : 0);
            const int k = atomicAdd(long_cnt + s, 1);
            if (k < long_cap) {
                long_list[((long long)s * long_cap + k) * 2] = i - (s == 2 ? b2 : (s == 1 ? b1 : 0));
                long_list[((long long)s * long_cap + k) * 2 + 1] = r.y;
            }
        }
    }
    dst[i] = o;
}

void launch_compact(const Rec* src, long long n, long long b1, long long b2, void* dst, long long* long_list, int long_cap,
                    int* long_cnt, cudaStream_t st) {
    if (n == 0) return;
    compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, n, b1, b2, (uint2*)dst, long_list, long_cap, long_cnt);
}

void launch_merge_count(const DevBatch& b, cudaStream_t st) {
    if (b.n_buckets == 0) return;
    merge_count_kernel<<<(unsigned)b.n_merge_blocks, MERGE_BLOCK, 0, st>>>(b);
    // partial[0 .. n) = per-block sums, partial[n] = grand total, partial[n + 1 ..] = segment prefixes
    const int nseg = (b.n_merge_blocks + MERGE_SEG - 1) / MERGE_SEG;
    BlockPartial* segs = b.partial + b.n_merge_blocks + 1;
    merge_scan_kernel<<<(unsigned)nseg, 1024, 0, st>>>(b.partial, b.n_merge_blocks, MERGE_SEG, segs, nullptr);
    merge_scan_kernel<<<1, 1024, 0, st>>>(segs, nseg, nseg, b.partial + b.n_merge_blocks, b.totals);
}
void launch_merge_write(const DevBatch& b, cudaStream_t st) {
    if (b.n_buckets == 0) return;
    merge_write_kernel<<<(unsigned)b.n_merge_blocks, MERGE_BLOCK, 0, st>>>(b);
}

// ---------------------------------------------------------------------------------------------------------------
// Anchor planes on request (generateAnchoredShiftXORs, parse_anchored_shiftxor.cpp:20-56): the host-side merges and
// the per-seed stage of the reference read the anchored planes B_m (fasta_utils.cpp:143-161). One thread per
// (word, shift), stateless: the run length in front of the word is found by looking back.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t anchor_word_stateless(const PlaneWord* __restrict__ cw, int L, int w, int s);
__global__ void __launch_bounds__(256) anchor_plane_kernel(const PlaneWord* __restrict__ cw, int L, int nw, int s_lo, int ns,
                                                           uint32_t* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nw * ns) return;
    const int w = (int)(t % nw), s = s_lo + (int)(t / nw);
    uint32_t a = anchor_word_stateless(cw, L, w, s);
    const long long rem = (long long)L - 32ll * w;  // bits past the contig end are not part of the plane
    if (rem < 32) a &= rem <= 0 ? 0u : ((1u << rem) - 1u);
    out[t] = a;
}

// stateless anchor word A_s[w] (shared by the plane export and the seed filter)
__device__ __forceinline__ uint32_t anchor_word_stateless(const PlaneWord* __restrict__ cw, int L, int w, int s) {
    const int K2 = 2 * s;
    int lenL = 0;
    for (int k = w - 1; k >= 0; --k) {
        const uint32_t xa = x_word(cw, k, s) | anchor_endmask(k, L, s);
        const int tr = clz32(~xa);
        lenL += tr;
        if (tr < 32 || lenL >= K2 + 32) break;
    }
    const uint32_t xa = x_word(cw, w, s) | anchor_endmask(w, L, s);
    const uint32_t xan = x_word(cw, w + 1, s) | anchor_endmask(w + 1, L, s);
    return anchor_word(cw, w, L, s, xa, xan, lenL);
}

// K5: the per-seed gate of processSeed / processSeedMotifWise (parse_seed.cpp:344-367, parse_smallmotif_seed.cpp:216-235):
// length of the seed sequence after truncation at the first N in [start, end + mlen), and the longest run of 1s of the
// anchored plane B_mlen over [start, end) (longestContinuousMatches, parse_seed.cpp:26-44). One thread per seed.
__global__ void __launch_bounds__(128) seed_filter_kernel(DevBatch b, const int4* __restrict__ seeds, long long n, int2* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int4 sd = seeds[t];  // contig, start, end, mlen
    const Contig cg = b.contigs[sd.x];
    const PlaneWord* __restrict__ cw = b.planes + cg.word_base;
    const int L = cg.L, start = sd.y, end = sd.z, m = sd.w;
    // first N in [start, end + m)
    int seq_len = (end - start) + m;
    {
        const int lim = min(end + m, L);
        for (int w = start >> 5; 32 * w < lim; ++w) {
            uint32_t nn = cw[w].n;
            if (w == (start >> 5)) nn &= ~lowmask(start & 31);
            if (nn) {
                const int p = 32 * w + ctz32(nn);
                if (p < lim) seq_len = p - start;
                break;
            }
        }
    }
    // longest run of B_m over [start, end)
    int best = 0, run = 0;
    const int lo = (m > 2) ? m - 2 : 1;
    for (int w = start >> 5; 32 * w < end; ++w) {
        uint32_t bm = x_word(cw, w, m);
        for (int s = lo; s <= m + 2; ++s)
            if (s != m && s >= b.lay.s_lo && s <= b.lay.s_hi) bm |= anchor_word_stateless(cw, L, w, s);
        if (w == (start >> 5)) bm &= ~lowmask(start & 31);
        const int rem = end - 32 * w;
        if (rem < 32) bm &= lowmask(rem);
        const int first = max(start - 32 * w, 0), last = min(rem, 32);
        if (bm == (lowmask(last) & ~lowmask(first))) { run += last - first; best = max(best, run); if (last < 32) run = 0; continue; }
        for (int i = first; i < last; ++i) {
            if ((bm >> i) & 1u) { ++run; best = max(best, run); } else run = 0;
        }
    }
    out[t] = make_int2(seq_len, best);
}

void launch_seed_filter(const DevBatch& b, const void* seeds, long long n, void* out, cudaStream_t st) {
    if (n == 0) return;
    seed_filter_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b, (const int4*)seeds, n, (int2*)out);
}

void launch_anchor_planes(const PlaneWord* cw, int L, int nw, int s_lo, int ns, uint32_t* out, cudaStream_t st) {
    const long long n = (long long)nw * ns;
    if (n == 0) return;
    anchor_plane_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cw, L, nw, s_lo, ns, out);
}

// ---------------------------------------------------------------------------------------------------------------
// Integer-pipe microbenchmark (roofline denominator of the scan): independent chains of funnel shifts and LOP3s,
// the instruction mix of the bit-sliced scan. 16 ops per thread per inner step.
// ---------------------------------------------------------------------------------------------------------------
// MODE 0: funnel shifts and LOP3s alternating (the mix of the bit-sliced scan), 1: LOP3 only, 2: funnel shifts only,
// 3: LOP3 alternating with IMAD (a multiply-add on the FMA pipe: shows what the second integer pipe adds)
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, int iters, uint32_t k) {
    uint32_t a0 = threadIdx.x, a1 = blockIdx.x, a2 = a0 * 3u + 1u, a3 = a1 * 5u + 2u, a4 = a0 ^ 0x55u, a5 = a1 ^ 0xAAu,
             a6 = a0 + 77u, a7 = a1 + 99u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) {
                a0 = __funnelshift_r(a0, a1, 7); a1 = (a1 & k) ^ a2;
                a2 = __funnelshift_r(a2, a3, 9); a3 = (a3 | k) ^ a4;
                a4 = __funnelshift_r(a4, a5, 11); a5 = (a5 & k) ^ a6;
                a6 = __funnelshift_r(a6, a7, 13); a7 = (a7 | k) ^ a0;
            } else if (MODE == 1) {
                a0 = (a0 | k) ^ a1; a1 = (a1 & k) ^ a2; a2 = (a2 | k) ^ a3; a3 = (a3 & k) ^ a4;
                a4 = (a4 | k) ^ a5; a5 = (a5 & k) ^ a6; a6 = (a6 | k) ^ a7; a7 = (a7 & k) ^ a0;
            } else if (MODE == 2) {
                a0 = __funnelshift_r(a0, a1, 7); a1 = __funnelshift_r(a1, a2, 5);
                a2 = __funnelshift_r(a2, a3, 9); a3 = __funnelshift_r(a3, a4, 3);
                a4 = __funnelshift_r(a4, a5, 11); a5 = __funnelshift_r(a5, a6, 2);
                a6 = __funnelshift_r(a6, a7, 13); a7 = __funnelshift_r(a7, a0, 6);
            } else {
                a0 = a0 * k + a1; a1 = (a1 & k) ^ a2;
                a2 = a2 * k + a3; a3 = (a3 | k) ^ a4;
                a4 = a4 * k + a5; a5 = (a5 & k) ^ a6;
                a6 = a6 * k + a7; a7 = (a7 | k) ^ a0;
            }
        }
    }
    const uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x12345678u) out[0] = r;
}

template <int MODE>
static double measure_mode(cudaStream_t st, uint32_t* scratch, int sms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, iters = 4096;
    int_peak_kernel<MODE><<<blocks, 256, 0, st>>>(scratch, 64, 0x0F0F0F0Fu);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        int_peak_kernel<MODE><<<blocks, 256, 0, st>>>(scratch, iters, 0x0F0F0F0Fu);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double ops = (double)blocks * 256.0 * iters * 32.0;
    return ops / (best * 1e-3);
}

double measure_int_peak(cudaStream_t st, uint32_t* scratch, int sms) { return measure_mode<0>(st, scratch, sms); }
void measure_int_peak_modes(cudaStream_t st, uint32_t* scratch, int sms, double out[4]) {
    out[0] = measure_mode<0>(st, scratch, sms);
    out[1] = measure_mode<1>(st, scratch, sms);
    out[2] = measure_mode<2>(st, scratch, sms);
    out[3] = measure_mode<3>(st, scratch, sms);
}

}  // namespace rb
