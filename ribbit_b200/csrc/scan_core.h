// ribbit-b200: per-lane scan logic shared by the sm_100a kernels (kernels.cu) and by the CPU warp emulator that
// tests/ uses to check the kernel logic without a GPU (tests/emu/emu_scan.cpp).
//
// Mapping: one warp lane owns one shift s of one band; a lane whose shift is a motif size m
// of the band additionally runs the three seed machines of the reference for that m:
//   perfect   parse_perfect_shiftxor.cpp:146-226      (runs of X_m & ~N)
//   subst     parse_substitute_shiftxor.cpp:391-577   (8-window, >=7 matches, Y = X_m)
//   anchored  parse_anchored_shiftxor.cpp:538-726     (8-window, >=6 matches, Y = B_m)
// with  X_s  fasta_utils.cpp:117-122,  anchors A_s  parse_anchored_shiftxor.cpp:20-56,
//       B_m  fasta_utils.cpp:143-161.
// A lane walks the words (32 positions each) of its chunk in order, all backward-looking state in registers.
//   FAST words (word_is_fast: every window ending in this and the previous word is evaluated, not the last word of
//   the contig) are handled bit-parallel: window pass masks (fail_ge2 / fail_ge3), component start / end masks
//   (ev_step), exact length filter (smear_step), run ends after six ones (perfect_fast); only candidates that reach
//   the consumer's cutoff are touched individually.
//   SLOW words (near N, contig start and end) go bit by bit through the reference's state machines verbatim
//   (lane_phase2, win_slow_bit, win_tail), which reproduces every quirk of SURVEY.md A.6.
// A fast -> slow transition rebuilds the machines' state by a warm-up that ends at the slow word; slow -> fast
// converts the machine state (win_to_fast).
#ifndef RB_SCAN_CORE_H
#define RB_SCAN_CORE_H

#include <stdint.h>

#ifdef __CUDACC__
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline
#endif

namespace rb {

// One word of the packed contig: 32 positions, bit i = position 32*w + i.
//  h, l : high / low bit of the 2-bit base code (A=00 C=01 G=10 T=11; N and padding 00)   fasta_utils.cpp:95-113
//  n    : N plane (anything that is not ACGTacgt); padding past the contig end is 1
//  v    : window-valid plane: v[p] = 1 iff p >= 7, p < L and no N in [p-7, p]  (the reference's
//         `valid_position >= window_length` test, parse_substitute_shiftxor.cpp:469)
struct alignas(16) PlaneWord {
    uint32_t h, l, n, v;
};

enum : int {
    STREAM_P = 0,
    STREAM_S = 1,
    STREAM_A = 2,
};
enum : int {
    REC_DROPPED = 1,   // below the consumer's length cutoff: only advances the consumer's cursors
    REC_PSEUDO = 2,    // synthetic cursor-advance record (end = max end of elided dropped candidates)
    REC_NOCOMMIT = 4,  // anchored tail flush whose returned cursors the reference discards
    REC_ENTRY = 8,     // raw pool only: a mask entry of a fast word (merge_core.h), expanded by the ordered compaction
};
static const int LEN_SAT = 1 << 24;

struct Rec {
    int32_t start, end;
    int32_t mflags;  // mlen | flags << 16
    int32_t key;     // (time - 32*bucket) << 18 | mlen << 2 | seq   (order inside a bucket)
};

RB_HD int ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return x ? __ffs((int)x) - 1 : 32;
#else
    return x ? __builtin_ctz(x) : 32;
#endif
}
RB_HD int clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
RB_HD int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// low 32 bits of (hi:lo) >> k, 0 <= k < 32
RB_HD uint32_t fsr(uint32_t lo, uint32_t hi, int k) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, k);
#else
    return k ? (lo >> k) | (hi << (32 - k)) : lo;
#endif
}
// high 32 bits of (hi:lo) << k, 0 <= k < 32: value[p] = in[p-k] with `lo` the previous word
RB_HD uint32_t fsl(uint32_t lo, uint32_t hi, int k) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, k);
#else
    return k ? (hi << k) | (lo >> (32 - k)) : hi;
#endif
}
// as fsl but 0 <= k <= 32 (k = 32 returns lo, the previous word)
RB_HD uint32_t fslc(uint32_t lo, uint32_t hi, int k) {
#ifdef __CUDA_ARCH__
    return __funnelshift_lc(lo, hi, k);
#else
    return k >= 32 ? lo : fsl(lo, hi, k);
#endif
}
// the low `nbits` bits set, 0 <= nbits <= 32
RB_HD uint32_t lowmask(int nbits) {
#ifdef __CUDA_ARCH__
    return __funnelshift_rc(0xFFFFFFFFu, 0u, 32 - nbits);  // (0 : ~0) >> (32 - nbits), shift clamped to 32
#else
    return nbits >= 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u);
#endif
}

struct LaneCfg {
    int s;       // shift handled by this lane (>= 1), 0 = idle lane
    int motif;   // 1 if the lane runs the seed machines for m = s
    int cutP;    // perfect cutoff, run closed by a mismatch or the contig end   parse_perfect_shiftxor.cpp:193,216
    int cutPN;   // perfect cutoff, run closed by an N                            parse_perfect_shiftxor.cpp:179
    int cutS;    // substitution keep cutoff  (m>30 ? m/3 : 10)                   parse_substitute_shiftxor.cpp:423
    int cutA;    // anchored keep cutoff                                          parse_anchored_shiftxor.cpp:572-573
    int wm;      // first word whose anchor view differs from X_s (positions >= L-s are forced to 1); per contig
    uint32_t dA; // smear shifts of the anchored keep filter, 6 bits each (smear_shifts): levels 0-4
    uint32_t dA2; //   levels 5-6
};

// consumer cutoffs
RB_HD int cut_perfect(int m) { return m <= 6 ? 12 - m : m; }        // parse_perfect_shiftxor.cpp:193,216
RB_HD int cut_subst(int m) { return m > 30 ? m / 3 : 10; }          // parse_substitute_shiftxor.cpp:423
RB_HD int cut_anch(int m) {                                         // parse_anchored_shiftxor.cpp:572-573
    // the reference computes (int)(0.9 * m) in double; 9 m / 10 in integers is the same number for every m up to 2000
    // (tests/test_oracle.py checks it), and keeps double arithmetic out of the kernels
    int c = m > 6 ? m : 10;
    if (m >= 10) c = (9 * m) / 10;
    return c;
}
static const int SMEAR_MAX = 96;  // positions the anchored keep filter looks back (exact for cutoffs up to this)

// Shifts d_0..d_6 (each <= 32) that smear a bit over exactly n <= 96 positions by doubling: with c_0 = 1,
// d_i = min(c_i, n - c_i, 32), c_{i+1} = c_i + d_i. Packed 6 bits each: d_0..d_4 in lo, d_5..d_6 in hi. For n >= 16 the
// first four are 1, 2, 4, 8.
RB_HD void smear_shifts(int n, uint32_t& lo, uint32_t& hi) {
    lo = 0u; hi = 0u;
    if (n < 1) n = 1;
    if (n > SMEAR_MAX) n = SMEAR_MAX;
    int c = 1;
    for (int i = 0; i < 7; ++i) {
        int d = (n - c < c) ? n - c : c;
        if (d > 32) d = 32;
        if (i < 5) lo |= (uint32_t)d << (6 * i); else hi |= (uint32_t)d << (6 * (i - 5));
        c += d;
    }
}

RB_HD LaneCfg make_lane_cfg(int s, int m_lo, int m_hi, int s_lo, int s_hi, int band_m0, int band_m1) {
    LaneCfg c;
    c.s = (s >= s_lo && s <= s_hi && s >= 1) ? s : 0;
    c.motif = (c.s != 0 && s >= m_lo && s <= m_hi && s >= band_m0 && s <= band_m1) ? 1 : 0;
    c.cutP = cut_perfect(s);
    c.cutPN = (s <= 6) ? 12 - s : s + (s - s_lo);
    c.cutS = cut_subst(s);
    c.cutA = cut_anch(s);
    c.wm = 0;
    smear_shifts(c.cutA, c.dA, c.dA2);
    return c;
}
// per-contig part of the lane configuration
RB_HD void lane_cfg_set_contig(LaneCfg& c, int L) { c.wm = (L - c.s) >> 5; }

// X_s word: bit p set iff code[p] == code[p+s] (zeros are shifted in past the contig end because the padding
// words are zero)  fasta_utils.cpp:121.  `cw` points at word 0 of the contig; indices -1 .. nw+guard-1 are readable.
RB_HD uint32_t x_word(const PlaneWord* cw, int w, int s) {
    const int off = s >> 5, sh = s & 31;
    const PlaneWord o = cw[w], a = cw[w + off], b = cw[w + off + 1];
    const uint32_t th = o.h ^ fsr(a.h, b.h, sh);
    const uint32_t tl = o.l ^ fsr(a.l, b.l, sh);
    return ~(th | tl);
}

// h / l halves of plane word `idx`, kept from the previous match-word computation (the "b" operand of word w is
// the "a" operand of word w+1)
struct XCache {
    uint32_t h, l;
    int idx;
};
RB_HD uint32_t x_word_cached(const PlaneWord* cw, int w, int s, XCache& xc) {
    const int off = s >> 5, sh = s & 31;
    const PlaneWord o = cw[w];
    uint32_t ah = xc.h, al = xc.l;
    if (xc.idx != w + off) {
        const PlaneWord a = cw[w + off];
        ah = a.h; al = a.l;
    }
    const PlaneWord b = cw[w + off + 1];
    xc.h = b.h; xc.l = b.l; xc.idx = w + off + 1;
    const uint32_t th = o.h ^ fsr(ah, b.h, sh);
    const uint32_t tl = o.l ^ fsr(al, b.l, sh);
    return ~(th | tl);
}

// same, when the previous call was for word w-1 (tight loop): the cache is known to be valid
RB_HD uint32_t x_word_next(const PlaneWord* cw, int w, int s, XCache& xc) {
    const int off = s >> 5, sh = s & 31;
    const PlaneWord o = cw[w];
    const PlaneWord b = cw[w + off + 1];
    const uint32_t th = o.h ^ fsr(xc.h, b.h, sh);
    const uint32_t tl = o.l ^ fsr(xc.l, b.l, sh);
    xc.h = b.h; xc.l = b.l; xc.idx = w + off + 1;
    return ~(th | tl);
}

// same with both plane words already loaded by the caller: o = cw[w], b = cw[w + (s >> 5) + 1]; xc.idx is left to the caller
RB_HD uint32_t x_word_next_ob(const PlaneWord& o, const PlaneWord& b, int s, XCache& xc) {
    const int sh = s & 31;
    const uint32_t th = o.h ^ fsr(xc.h, b.h, sh);
    const uint32_t tl = o.l ^ fsr(xc.l, b.l, sh);
    xc.h = b.h; xc.l = b.l;
    return ~(th | tl);
}
// running pointers of the tight loop: o = &cw[w + 1] (the same for all lanes), b = &cw[w + 1 + (s >> 5) + 1] (per lane)
struct SeqPtrs {
    const PlaneWord* o;
    const PlaneWord* b;
};

// positions p >= L - s never close an anchor run (parse_anchored_shiftxor.cpp:37): force them to 1 so that the
// run that reaches L-1-s looks unbounded and is dropped by the "< 2*s" test.
RB_HD uint32_t anchor_endmask(int w, int L, int s) {
    const long long lim = (long long)L - s - 32ll * w;  // first masked bit index in this word
    if (lim >= 32) return 0u;
    if (lim <= 0) return 0xFFFFFFFFu;
    return 0xFFFFFFFFu << lim;
}

// number of consecutive 1s of the anchor view of X_s starting at bit 0 of word w, capped at `cap`
RB_HD int anchor_ones_from(const PlaneWord* cw, int w, int L, int s, int cap) {
    int tot = 0;
    while (tot < cap) {
        if (32ll * w >= (long long)L - s) return cap;  // everything from here on is forced to 1
        const uint32_t xa = x_word(cw, w, s) | anchor_endmask(w, L, s);
        const int k = ctz32(~xa);
        tot += k;
        if (k < 32) break;
        ++w;
    }
    return tot < cap ? tot : cap;
}

// Anchor word A_s[w] (parse_anchored_shiftxor.cpp:34-55): runs of 1s of X_s with 3 <= len < 2s that are closed
// inside the scanned range. lenL = length of the run of 1s ending at bit 31 of word w-1 (saturating).
RB_HD uint32_t anchor_word(const PlaneWord* cw, int w, int L, int s, uint32_t xa, uint32_t xa_nxt, int& lenL) {
    const int K2 = 2 * s;
    if (xa == 0u) { lenL = 0; return 0u; }
    const int lead = ctz32(~xa);
    if (lead == 32) {  // the whole word is inside one run
        uint32_t a = 0u;
        if (lenL + 32 < K2) {
            int rext = ctz32(~xa_nxt);
            if (rext == 32) rext = 32 + anchor_ones_from(cw, w + 2, L, s, K2);
            if (lenL + 32 + rext < K2) a = 0xFFFFFFFFu;
        }
        lenL = (lenL + 32 > LEN_SAT) ? LEN_SAT : lenL + 32;
        return a;
    }
    const int trail = clz32(~xa);
    uint32_t a = 0u, mid = xa;
    if (lead > 0) {
        const int tot = lenL + lead;
        if (tot >= 3 && tot < K2) a |= lowmask(lead);
        mid &= ~lowmask(lead);
    }
    if (trail > 0) {
        int rext = ctz32(~xa_nxt);
        if (rext == 32 && trail + 32 < K2) rext = 32 + anchor_ones_from(cw, w + 2, L, s, K2);
        const int tot = trail + rext;
        if (tot >= 3 && tot < K2) a |= ~lowmask(32 - trail);
        mid &= lowmask(32 - trail);
    }
    // runs strictly inside the word
    const uint32_t r3 = mid & (mid >> 1) & (mid >> 2);
    uint32_t keep = r3 | (r3 << 1) | (r3 << 2);
    if (K2 <= 30 && keep) {
        uint32_t e = mid;
        for (int k = 1; k < K2;) { const int st = (k < K2 - k) ? k : K2 - k; e &= e >> st; k += st; }
        if (e) {
            uint32_t d = e;
            for (int k = 1; k < K2;) { const int st = (k < K2 - k) ? k : K2 - k; d |= d << st; k += st; }
            keep &= ~d;
        }
    }
    lenL = trail;
    return a | keep;
}

// Bit-sliced "number of mismatches in the last 8 positions". z = ~Y is the mismatch mask; per position t
//   o1 / t1 : at least one / two zeros among z[t-1..t]        o2 / t2 / u2 : at least one / two / three among z[t-3..t]
// The previous word's partial counts are carried (only their top bits are read, through funnel shifts).
struct WinCarryS {
    uint32_t o1, t1, o2, t2;
};
struct WinCarryA {
    uint32_t b, o1, t1, o2, t2, u2;  // b = the previous word of Y itself
};
// carries as if the previous word had the mismatch mask zp (exact in the top bits, which is all that is read)
RB_HD void win_carry_init(WinCarryS& c, uint32_t zp) {
    const uint32_t zp1 = zp << 1;
    c.o1 = zp | zp1; c.t1 = zp & zp1;
    const uint32_t o1s = c.o1 << 2, t1s = c.t1 << 2;
    c.o2 = c.o1 | o1s; c.t2 = c.t1 | t1s | (c.o1 & o1s);
}
RB_HD void win_carry_init(WinCarryA& c, uint32_t zp) {
    const uint32_t zp1 = zp << 1;
    c.b = ~zp;
    c.o1 = zp | zp1; c.t1 = zp & zp1;
    const uint32_t o1s = c.o1 << 2, t1s = c.t1 << 2;
    c.o2 = c.o1 | o1s; c.t2 = c.t1 | t1s | (c.o1 & o1s);
    c.u2 = (c.t1 & o1s) | (c.o1 & t1s);
}

// fail mask for ">= 7 of 8" (substitution pass): bit p set iff X[p-7..p] holds >= 2 zeros. l1 = fsl(x_prev, x, 1).
// cand: the perfect-run ends that follow six ones (x[t] = 0, x[t-6..t-1] = 1; every perfect cutoff is >= 6).
RB_HD uint32_t fail_ge2(uint32_t x, uint32_t l1, WinCarryS& c, uint32_t& cand) {
    const uint32_t o1 = ~(x & l1), t1 = ~(x | l1);
    const uint32_t o1s = fsl(c.o1, o1, 2), t1s = fsl(c.t1, t1, 2);
    const uint32_t o2 = o1 | o1s;
    const uint32_t t2 = t1 | t1s | (o1 & o1s);
    const uint32_t o2s = fsl(c.o2, o2, 4), t2s = fsl(c.t2, t2, 4);
    const uint32_t t3 = t2 | t2s | (o2 & o2s);
    cand = ~(x | fsl(c.o2, o2, 1) | fsl(c.o1, o1, 5));
    c.o1 = o1; c.t1 = t1; c.o2 = o2; c.t2 = t2;
    return t3;
}
// fail mask for ">= 6 of 8" (anchored pass): bit p set iff B[p-7..p] holds >= 3 zeros
RB_HD uint32_t fail_ge3(uint32_t b, WinCarryA& c) {
    const uint32_t l1 = fsl(c.b, b, 1);
    const uint32_t o1 = ~(b & l1), t1 = ~(b | l1);
    const uint32_t o1s = fsl(c.o1, o1, 2), t1s = fsl(c.t1, t1, 2);
    const uint32_t o2 = o1 | o1s;
    const uint32_t t2 = t1 | t1s | (o1 & o1s);
    const uint32_t u2 = (t1 & o1s) | (o1 & t1s);
    const uint32_t o2s = fsl(c.o2, o2, 4), t2s = fsl(c.t2, t2, 4), u2s = fsl(c.u2, u2, 4);
    const uint32_t u3 = u2 | u2s | (t2 & o2s) | (o2 & t2s);
    c.b = b; c.o1 = o1; c.t1 = t1; c.o2 = o2; c.t2 = t2; c.u2 = u2;
    return u3;
}

// reference window machine state for one motif and one stream (parse_substitute_shiftxor.cpp:408-410); maintained
// in slow words only
struct WinState {
    int cur, ls, le;
};

// Bit-parallel view of the same machine, used in fast words (every window of this word and of the previous word is
// evaluated). P bit t = the window ending at position t passes. With F = ~P:
//   R8[t] = F[t-7..t] all set (eight failing windows in a row, i.e. the reference's "gap > 7 windows")
//   S[t]  = P[t] & R8[t-1]           a component starts: first passing window after >= 8 failing ones
//   E[t]  = P[t-9] & R8[t-1]         the reference emits the component here: its last run ended with the failing
//                                    window t-8 (= le), and t is the first window start beyond le
//                                    (parse_substitute_shiftxor.cpp:475-530)
// The component emitted at E-bit t is (ls, le) = (ts - 7, t - 8) with ts the latest S bit before t.
struct EvCarry {
    uint32_t r2, r4, r8, P, S;  // previous word
    int lastS;                  // position of the latest S bit so far
};
RB_HD void ev_step(uint32_t P, EvCarry& c, uint32_t& S, uint32_t& E, uint32_t& Sprev) {
    const uint32_t r2 = ~(P | fsl(c.P, P, 1));  // two failing windows in a row
    const uint32_t r4 = r2 & fsl(c.r2, r2, 2);
    const uint32_t r8 = r4 & fsl(c.r4, r4, 4);
    const uint32_t R81 = fsl(c.r8, r8, 1);
    S = P & R81;
    E = fsl(c.P, P, 9) & R81;
    Sprev = c.S;
    c.r2 = r2; c.r4 = r4; c.r8 = r8; c.P = P; c.S = S;
}

struct LaneState {
    uint32_t x_prev, x_cur, x_nxt;  // X_s of words w-1, w, w+1
    int lenL;                       // anchor-view run length ending at the end of word w-1
    WinCarryS cs;
    WinCarryA ca;
    int pst;                        // perfect machine: start of the open run or -1 (last_starts); slow words
    WinState S, A;                  // valid while the previous word was a slow word
    EvCarry es, ea;                 // carries always valid; lastS valid while the previous word was a fast word
    uint32_t sm[7];                 // anchored keep filter: previous word of each smear level
    XCache xc;
    // warm-up bookkeeping (chunks that do not start at the contig start)
    int sync;                       // bit0 anchors exact, bit1 perfect, bit2 subst, bit3 anchored
    int zS, zA;                     // consecutive evaluated failing windows (saturating)
};
enum : int { SYNC_X = 1, SYNC_P = 2, SYNC_S = 4, SYNC_A = 8, SYNC_ALL = 15 };

// Word w is a FAST word when every window ending in words w-1 and w is evaluated (no N within reach) and w is not the
// last word of the contig; all other words are SLOW words and go through the reference's state machines bit by bit.
// The contig start counts as evaluated: the windows that would end at positions 0..6 do not exist
// (parse_substitute_shiftxor.cpp:469), which for the bit-parallel view is the same as failing windows with nothing but
// failing windows in front of them (lane_init) - the fast path only has to mask them out of its pass words (v_eff).
RB_HD uint32_t v_eff(const PlaneWord* cw, int w) { return cw[w].v | (w == 0 ? 0x7Fu : 0u); }
RB_HD int word_is_fast(const PlaneWord* cw, int w, int nw) {
    return v_eff(cw, w) == 0xFFFFFFFFu && (w == 0 || v_eff(cw, w - 1) == 0xFFFFFFFFu) && w != nw - 1;
}

// Context of one lane iteration; the Sink receives the events.
//   Sink::rec(stream, start, end, mlen, flags, key)   a record that goes to bucket `w` of the stream
//   Sink::dropped(stream, tw)                          a fast-word candidate below the cutoff (only its time counts)
struct IterCtx {
    int w;          // word (bucket) being processed
    int L;          // contig length
    int emit_on;    // 0 while warming up (w < first emitting word)
    int slow;       // 1 if this word goes through the bit-serial path
    int prev_slow;  // 1 if the previous word did
    int fastrun;    // consecutive fast words up to and including this one (saturating)
};

template <class Sink>
RB_HD void emit_win(Sink& sk, const IterCtx& it, const LaneCfg& cfg, int stream, int ls, int le, int time) {
    if (!it.emit_on) return;
    const int cut = (stream == STREAM_S) ? cfg.cutS : cfg.cutA;
    const int tw = time - 32 * it.w;
    if (le - ls >= cut) sk.rec(stream, ls, le, cfg.s, 0, (tw << 18) | (cfg.s << 2));
    else sk.rec(stream, ls, le, cfg.s, REC_DROPPED, (tw << 18) | (cfg.s << 2));
}
template <class Sink>
RB_HD void emit_perfect(Sink& sk, const IterCtx& it, const LaneCfg& cfg, int start, int end, int time) {
    if (!it.emit_on) return;
    sk.rec(STREAM_P, start, end, cfg.s, 0, ((time - 32 * it.w) << 18) | (cfg.s << 2));
}

// Anchored keep filter. M1[t] = OR of S[t-k], k = 1..n, n = min(cutA, SMEAR_MAX): an E bit with M1 set belongs to a
// component whose S bit is at most n positions back, i.e. whose length t - ts - 1 is below the consumer's cutoff
// (parse_anchored_shiftxor.cpp:153). Exact for cutA <= SMEAR_MAX, else a prefilter. Valid when the last four words were
// fast words (it looks back up to three words).
RB_HD uint32_t smear_step(const LaneCfg& cfg, LaneState& st, uint32_t S, uint32_t Sprev) {
    uint32_t v = fsl(Sprev, S, 1);
#define RB_SMEAR_LEVEL(i, d)                                  \
    {                                                         \
        const uint32_t nv = v | fslc(st.sm[i], v, (int)(d)); \
        st.sm[i] = v;                                         \
        v = nv;                                               \
    }
    RB_SMEAR_LEVEL(0, cfg.dA & 63u)
    RB_SMEAR_LEVEL(1, (cfg.dA >> 6) & 63u)
    RB_SMEAR_LEVEL(2, (cfg.dA >> 12) & 63u)
    RB_SMEAR_LEVEL(3, (cfg.dA >> 18) & 63u)
    RB_SMEAR_LEVEL(4, (cfg.dA >> 24) & 63u)
    RB_SMEAR_LEVEL(5, cfg.dA2 & 63u)
    RB_SMEAR_LEVEL(6, (cfg.dA2 >> 6) & 63u)
#undef RB_SMEAR_LEVEL
    return v;
}

// The same filter without the smear network, for lanes whose cutoff is at least 32 (exact for every such cutoff): the E
// bit at position p0 + i belongs to the component that started at the latest S bit in front of it. An S bit below i in the
// same word is less than 32 positions back: below the cutoff. Otherwise the start is lastS and the component is kept iff
// p0 + i - lastS - 1 >= cut, i.e. i >= lastS + cut + 1 - p0 (merge_core.h entry_interval computes the same interval).
// cut1 = cut + 1.
// bits thr..31 (all for thr <= 0, none for thr >= 32)
RB_HD uint32_t bits_from(int thr) {
    thr = thr < 0 ? 0 : thr;
#ifdef __CUDA_ARCH__
    return __funnelshift_lc(0u, 0xFFFFFFFFu, thr);  // the shift clamps at 32
#else
    return thr >= 32 ? 0u : (0xFFFFFFFFu << thr);
#endif
}
// The general form, for any cutoff: `near` = the positions that have an S bit of THIS word at most cut positions below
// them (in-word smear of S, by the caller). With an S bit below i in the word the component started there and is kept iff
// i is not near; with none it started at lastS.
RB_HD uint32_t keep_by_last_near(uint32_t E, uint32_t S, uint32_t near, int lastS, int cut1, int p0) {
    const uint32_t upto = S ^ (S - 1u);  // bits up to and including the lowest S bit (all bits when the word has none)
    return E & ~near & (~upto | bits_from(lastS + cut1 - p0));
}
RB_HD uint32_t keep_by_last(uint32_t E, uint32_t S, int lastS, int cut1, int p0) {
    const uint32_t below = S ^ (S - 1u);  // bits up to and including the lowest S bit of the word (all bits when it has none):
                                          // an S bit AT i starts the next component, it is not in front of i
    int thr = lastS + cut1 - p0;
    thr = thr < 0 ? 0 : thr;
#ifdef __CUDA_ARCH__
    const uint32_t far = __funnelshift_lc(0u, 0xFFFFFFFFu, thr);  // bits thr..31; none for thr >= 32 (the shift clamps)
#else
    const uint32_t far = thr >= 32 ? 0u : (0xFFFFFFFFu << thr);
#endif
    return E & below & far;
}
// The smear state a lane needs when it goes back to smear_step after words filtered by keep_by_last (w = the next word):
// what any later E bit asks is whether the LATEST S bit lies within n positions, so a history that holds just that one
// bit gives the same answers as the true one. Replays the four words in front of w (the network looks back 96 positions).
RB_HD void smear_from_last(const LaneCfg& cfg, LaneState& st, int w) {
    for (int i = 0; i < 7; ++i) st.sm[i] = 0u;
    const int lastS = st.ea.lastS;
    uint32_t Sprev = 0u;
    for (int k = 4; k >= 1; --k) {
        const uint32_t S = (lastS >= 0 && (lastS >> 5) == w - k) ? (1u << (lastS & 31)) : 0u;
        smear_step(cfg, st, S, Sprev);
        Sprev = S;
    }
}

// ---- fast word -----------------------------------------------------------------------------------------------------
// A fast word hands its candidates over as MASK ENTRIES (Sink::entry): per stream the E bits (perfect stream: the run ends)
// that survived the bit-parallel prefilters, the S mask (run starts) of the word and the position of the latest S bit in
// front of the word. The ordered compaction expands them into records and applies the consumer's cutoff exactly
// (merge_core.h entry_interval); E bits that fail it only contribute their emission time, like the prefiltered ones.
// lastS of the bit-parallel view from the reference machine's state (slow word -> fast word)
RB_HD void win_to_fast(const WinState& st, int& lastS) {
    if (st.ls != -1) lastS = st.ls + 7;
    else if (st.cur != -1) lastS = st.cur + 7;
}

// fast word -> slow word: the reference machine's state at the end of fast word w-1 (p0 = 32*(w-1)) from the bit-parallel
// view of that word: c.P = its pass mask, c.r8 = "eight failing windows in a row end here". After NINE failing windows in
// a row the machine is idle (the pending component was emitted at the ninth, parse_substitute_shiftxor.cpp:500-530); the
// windows of the word behind the last such point are run through the machine without emitting (the fast path emitted
// them). Returns false when the word holds no nine failing windows in a row: the caller rebuilds the state by a warm-up.
RB_HD bool win_from_fast(const EvCarry& c, int p0, WinState& st) {
    const uint32_t n9 = c.r8 & (c.r8 << 1);
    if (!n9) return false;
    st.cur = st.ls = st.le = -1;
    for (int i = 32 - clz32(n9); i < 32; ++i) {
        const int wp = p0 + i - 7;
        if ((c.P >> i) & 1u) {
            if (st.cur == -1) {
                st.cur = wp;
                if (st.le != -1 && st.le < st.cur) { st.ls = -1; st.le = -1; }
            }
        } else if (st.cur != -1) {
            if (st.ls == -1) st.ls = st.cur;
            st.le = wp + 7;
            st.cur = -1;
        } else if (st.le != -1 && st.le < wp) {
            st.ls = -1; st.le = -1;
        }
    }
    return true;
}
// Start of the run of G = X_s & ~N that contains position 32*w + i - 1 (bit i-1 of x = X_s[w] is set, i in [1, 32], no N
// in word w). Stateless: looks back over the earlier words as far as the run goes; this is the reference's last_starts
// (parse_perfect_shiftxor.cpp:194-200) without carrying it.
RB_HD int perfect_run_start(const PlaneWord* cw, int w, int i, int s, uint32_t x) {
    const uint32_t zb = ~x & lowmask(i);  // zeros below bit i
    if (zb) return 32 * w + 32 - clz32(zb);
    for (int k = w - 1; k >= 0; --k) {
        const uint32_t z = ~(x_word(cw, k, s) & ~cw[k].n);
        if (z) return 32 * k + 32 - clz32(z);
    }
    return 0;
}

// Exact keep filters of a fast word (the consumer's length cutoffs): the bits of x (E bits that survived the bit-parallel
// prefilters; perfect stream: run ends that follow six ones) whose candidate reaches `cut`.
// window streams: the component emitted at E-bit i is (ls, le) = (ts - 7, p0 + i - 8), ts = the latest S bit before i
RB_HD uint32_t kept_exact(int cut, int p0, uint32_t x, uint32_t S, int lastS) {
    uint32_t kept = 0u;
    while (x) {
        const int i = ctz32(x);
        x &= x - 1u;
        const uint32_t sb = S & lowmask(i);
        const int ts = sb ? p0 + 31 - clz32(sb) : lastS;
        if (p0 + i - ts - 1 >= cut) kept |= 1u << i;
    }
    return kept;
}
// perfect stream: the run that ends at p0 + i started at the latest run start before i (sx = run starts of the word,
// xw = X_s[w]); parse_perfect_shiftxor.cpp:190-208
RB_HD uint32_t kept_exact_perfect(const PlaneWord* cw, int w, int s, int cut, uint32_t xw, uint32_t sx, uint32_t cand) {
    uint32_t kept = 0u;
    const int p0 = 32 * w;
    while (cand) {
        const int i = ctz32(cand);
        cand &= cand - 1u;
        const uint32_t sb = sx & lowmask(i);
        const int a = sb ? p0 + 31 - clz32(sb) : perfect_run_start(cw, w, i, s, xw);
        if (p0 + i - a >= cut) kept |= 1u << i;
    }
    return kept;
}

// All three machines of a motif lane at a fast -> slow transition in front of word w (x_prev = X_m[w-1] already rotated).
// The new state is returned apart from the lane state: the caller applies it only if every lane of the item succeeded.
struct SlowEntry {
    WinState S, A;
    int pst;
};
RB_HD bool lane_to_slow(const LaneCfg& cfg, const LaneState& st, const PlaneWord* cw, int w, SlowEntry& e) {
    e.S = st.S; e.A = st.A; e.pst = st.pst;
    if (!cfg.motif) return true;
    if (!win_from_fast(st.es, 32 * (w - 1), e.S) || !win_from_fast(st.ea, 32 * (w - 1), e.A)) return false;
    // the run that is open at the word boundary (parse_perfect_shiftxor.cpp:194)
    e.pst = (st.x_prev >> 31) ? perfect_run_start(cw, w - 1, 32, cfg.s, st.x_prev) : -1;
    return true;
}
RB_HD void lane_enter_slow(LaneState& st, const SlowEntry& e) { st.S = e.S; st.A = e.A; st.pst = e.pst; }

// ---- slow word: the reference state machines bit by bit ----------------------------------------------------------
template <class Sink>
RB_HD void win_slow_bit(Sink& sk, const IterCtx& it, const LaneCfg& cfg, int stream, int p, int nbit, int vbit, int pass,
                        WinState& st, int& zrun, int& sync, int syncbit, int prev_v) {
    const int wp = p - 7;
    if (nbit) {  // parse_substitute_shiftxor.cpp:433-458
        if (st.cur != -1) {
            st.cur = wp;
            if (st.le != -1 && st.le < st.cur) {
                emit_win(sk, it, cfg, stream, st.ls, st.le, p);
                st.ls = -1; st.le = -1;
            }
        }
        st.cur = -1;
        zrun = 0;
    } else if (vbit) {  // :469-530
        if (pass) {
            if (st.cur == -1) {
                st.cur = wp;
                if (st.le != -1 && st.le < st.cur) {
                    emit_win(sk, it, cfg, stream, st.ls, st.le, p);
                    st.ls = -1; st.le = -1;
                }
            }
            zrun = 0;
        } else {
            if (st.cur != -1) {
                if (st.ls == -1) st.ls = st.cur;
                st.le = wp + 7;
                st.cur = -1;
            } else if (st.le != -1 && st.le < wp) {
                emit_win(sk, it, cfg, stream, st.ls, st.le, p);
                st.ls = -1; st.le = -1;
            }
            if (zrun < 64) zrun++;
        }
        // warm-up: the state is history-free after the first window of a valid stretch or after 9 failing windows
        if (!prev_v || zrun >= 9) sync |= syncbit;
    } else {
        zrun = 0;
    }
}

// Tail flush of one stream after the last position (parse_substitute_shiftxor.cpp:534-574,
// parse_anchored_shiftxor.cpp:681-723). Records go to the bucket after the last word; order key = (mlen, seq).
template <class Sink>
RB_HD void win_tail(Sink& sk, const LaneCfg& cfg, int stream, int L, const WinState& st) {
    const int cut = (stream == STREAM_S) ? cfg.cutS : cfg.cutA;
    const int m = cfg.s;
    int s0 = -1, e0 = 0, s1 = -1, e1 = 0, commit0 = 1;
    if (st.le == -1) {
        if (st.cur != -1) { s0 = st.cur; e0 = L; commit0 = 0; }
    } else if (st.cur == -1) {
        s0 = st.ls; e0 = st.le; commit0 = 0;
    } else if (st.le >= st.cur - m) {
        s0 = st.ls; e0 = L; commit0 = 0;
    } else {
        s0 = st.ls; e0 = st.le; commit0 = 1;
        s1 = st.cur; e1 = L;
    }
    // the substitution pass always stores the returned cursor; the anchored pass only for the first of two calls
    const int nc0 = (stream == STREAM_A && !commit0) ? REC_NOCOMMIT : 0;
    const int nc1 = (stream == STREAM_A) ? REC_NOCOMMIT : 0;
    if (s0 != -1) {
        if (e0 - s0 >= cut) sk.rec(stream, s0, e0, m, nc0, (m << 2) | 0);
        else if (!nc0) sk.rec(stream, s0, e0, m, REC_DROPPED, (m << 2) | 0);
    }
    if (s1 != -1) {
        if (e1 - s1 >= cut) sk.rec(stream, s1, e1, m, nc1, (m << 2) | 1);
        else if (!nc1) sk.rec(stream, s1, e1, m, REC_DROPPED, (m << 2) | 1);
    }
}

// ---- one lane, one word ------------------------------------------------------------------------------------------
// Phase 1: bring X_s[w+1] in and compute the anchor word A_s[w]. Returns A_s[w] (0 for idle lanes).
RB_HD uint32_t lane_phase1(const LaneCfg& cfg, LaneState& st, const PlaneWord* cw, int w, int L) {
    if (cfg.s == 0) return 0u;
    st.x_nxt = x_word_cached(cw, w + 1, cfg.s, st.xc);
    const uint32_t xa = st.x_cur | anchor_endmask(w, L, cfg.s);
    const uint32_t xan = st.x_nxt | anchor_endmask(w + 1, L, cfg.s);
    // a zero was seen: run lengths are exact from here on; a run already >= 2s long can never be an anchor,
    // whatever its true length
    if (xa != 0xFFFFFFFFu || st.lenL >= 2 * cfg.s) st.sync |= SYNC_X;
    return anchor_word(cw, w, L, cfg.s, xa, xan, st.lenL);
}

// Phase 1 for emitting words (warm-up bookkeeping not needed). Common case in straight-line code: no run touching
// this word can reach 2s positions and the word is not near the contig end; then A_s[w] = the positions of X_s that
// lie in a run of at least 3 (anchor_size, parse_anchored_shiftxor.cpp:44). Everything else takes anchor_word.
template <bool SEQ>
RB_HD uint32_t lane_phase1_fast_t(const LaneCfg& cfg, LaneState& st, const PlaneWord* cw, int w, int L, uint32_t* v_next = nullptr,
                                  SeqPtrs* sp = nullptr, bool small = true) {
    if (SEQ) {
        // every lane (idle ones too) reads the same word w+1: its v field is the caller's fast-word test for the next step
        const PlaneWord o = *sp->o, b = *sp->b;
        ++sp->o; ++sp->b;
        *v_next = o.v;
        if (cfg.s == 0) return 0u;
        st.x_nxt = x_word_next_ob(o, b, cfg.s, st.xc);
    } else {
        if (cfg.s == 0) return 0u;
        st.x_nxt = x_word_cached(cw, w + 1, cfg.s, st.xc);
    }
    const uint32_t x = st.x_cur, xn = st.x_nxt, xp = st.x_prev;
    const int K2 = 2 * cfg.s;
    // SEQ: the caller keeps w + 1 below every lane's wm
    if ((!SEQ && (w + 1 >= cfg.wm)) | (x == 0xFFFFFFFFu) | (xn == 0xFFFFFFFFu)) {  // rare: whole words of ones, contig end
        const uint32_t xa = x | anchor_endmask(w, L, cfg.s);
        const uint32_t xan = xn | anchor_endmask(w + 1, L, cfg.s);
        if (xa != 0xFFFFFFFFu || st.lenL >= K2) st.sync |= SYNC_X;
        return anchor_word(cw, w, L, cfg.s, xa, xan, st.lenL);
    }
    st.sync |= SYNC_X;  // this word holds a mismatch: run lengths are exact from here on
    const int lead = ctz32(~x), trail = clz32(~x), leadn = ctz32(~xn);
    const uint32_t l1 = fsl(xp, x, 1), l2 = fsl(xp, x, 2), r1 = fsr(x, xn, 1), r2 = fsr(x, xn, 2);
    uint32_t a = x & ((l1 & (l2 | r1)) | (r1 & r2));
    // runs that touch a word edge: their full length is known from the neighbours
    a &= (st.lenL + lead >= K2) ? ~lowmask(lead) : 0xFFFFFFFFu;
    a &= (trail + leadn >= K2) ? lowmask(32 - trail) : 0xFFFFFFFFu;
    if (small) {  // shifts <= 15 in this item: a run inside the word can be too long as well
        // K2 ones in a row by doubling, branch-free for every lane: five steps reach min(K2, 32) ones, and 32 ones in a
        // row cannot occur here (x is not all ones), so lanes with K2 > 30 end with e == 0
        uint32_t e = x;
        int k = 1;
#pragma unroll
        for (int step = 0; step < 5; ++step) {
            int sh = K2 - k;
            sh = sh < k ? sh : k;
            sh = sh < 0 ? 0 : sh;
            e &= e >> sh;
            k += sh;
        }
        if (e) {
            uint32_t d = e;
            for (int k = 1; k < K2;) { const int sh = (k < K2 - k) ? k : K2 - k; d |= d << sh; k += sh; }
            a &= ~d;
        }
    }
    st.lenL = trail;
    return a;
}

// nz = failing windows of this word; zrun = failing windows at the end of the previous words (saturating).
// Returns 1 if nine consecutive failing windows were seen.
RB_HD int nine_fails(uint32_t nz, int& zrun) {
    const int lead = ctz32(~nz);
    uint32_t e = nz & (nz >> 1);
    e &= e >> 2;
    e &= e >> 4;
    e &= nz >> 8;
    const int hit = (e != 0u) || (zrun + lead >= 9);
    const int tr = clz32(~nz);
    zrun = (tr == 32) ? ((zrun + 32 > 64) ? 64 : zrun + 32) : tr;
    return hit;
}

RB_HD uint32_t lane_phase1_fast(const LaneCfg& cfg, LaneState& st, const PlaneWord* cw, int w, int L) {
    return lane_phase1_fast_t<false>(cfg, st, cw, w, L);
}
// the previous call (either phase-1 variant) was for word w-1
// (tight loop; the caller owns the running pointers and restores st.xc.idx when it leaves the loop)
// small: some lane of the item has a shift <= 15 (uniform over the item; true is always correct)
RB_HD uint32_t lane_phase1_fast_seq(const LaneCfg& cfg, LaneState& st, const PlaneWord* cw, int w, int L, uint32_t& v_next, SeqPtrs& sp,
                                    bool small) {
    return lane_phase1_fast_t<true>(cfg, st, cw, w, L, &v_next, &sp, small);
}

// Phase 2 of a fast word (it.slow == 0, machines on): every window is evaluated, so o.v is all ones and the N plane is
// not consulted.
template <class Sink>
RB_HD void lane_phase2_fast(Sink& sk, const LaneCfg& cfg, LaneState& st, const PlaneWord* cw, const IterCtx& it, uint32_t a_m2,
                            uint32_t a_m1, uint32_t a_p1, uint32_t a_p2) {
    if (cfg.motif) {
        const uint32_t x = st.x_cur;
        const uint32_t l1 = fsl(st.x_prev, x, 1);
        const uint32_t b = x | a_m2 | a_m1 | a_p1 | a_p2;
        uint32_t cand;
        const uint32_t vm = it.w == 0 ? cw[0].v : 0xFFFFFFFFu;  // the windows that would end at positions 0..6 do not exist
        const uint32_t passS = ~fail_ge2(x, l1, st.cs, cand) & vm;
        const uint32_t passA = ~fail_ge3(b, st.ca) & vm;
        uint32_t sS, eS, sSp, sA, eA, sAp;
        ev_step(passS, st.es, sS, eS, sSp);
        ev_step(passA, st.ea, sA, eA, sAp);
        const uint32_t killA = smear_step(cfg, st, sA, sAp);
        if (it.prev_slow) {
            win_to_fast(st.S, st.es.lastS);
            win_to_fast(st.A, st.ea.lastS);
        }
        if (!it.emit_on) {
            // warm-up in fast words: lastS is exact once a component start was seen or nine windows in a row failed (any
            // later component starts inside the scanned range); the perfect stream keeps no state in fast words
            st.sync |= SYNC_P;
            if (sS != 0u || nine_fails(~passS, st.zS)) st.sync |= SYNC_S;
            if (sA != 0u || nine_fails(~passA, st.zA)) st.sync |= SYNC_A;
        }
        const int p0 = 32 * it.w;
        // prefilters: a substitution component whose S bit is 9 or 10 back has length 8 or 9, below every cutoff
        // (parse_substitute_shiftxor.cpp:423: >= 10); the smear looks back into the previous word: trust it once a few
        // fast words in a row were seen
        uint32_t xS = it.emit_on ? (eS & ~(fsl(sSp, sS, 9) | fsl(sSp, sS, 10))) : 0u;
        uint32_t xA = it.emit_on ? (eA & ~(it.fastrun >= 4 ? killA : 0u)) : 0u;
        uint32_t xP = it.emit_on ? cand : 0u;
        // the survivors are checked exactly, so that an entry only holds candidates the consumer keeps
        if (xS) xS = kept_exact(cfg.cutS, p0, xS, sS, st.es.lastS);
        if (xA && (cfg.cutA > SMEAR_MAX || it.fastrun < 4)) xA = kept_exact(cfg.cutA, p0, xA, sA, st.ea.lastS);
        if (xP) xP = kept_exact_perfect(cw, it.w, cfg.s, cfg.cutP, x, x & ~l1, xP);
        if (xP) sk.entry(STREAM_P, cfg.s, xP, x & ~l1, 0);
        if (xS) sk.entry(STREAM_S, cfg.s, xS, sS, st.es.lastS);
        if (xA) sk.entry(STREAM_A, cfg.s, xA, sA, st.ea.lastS);
        sk.dropped_mask(STREAM_S, eS & ~xS);
        sk.dropped_mask(STREAM_A, eA & ~xA);
        st.es.lastS = sS ? p0 + 31 - clz32(sS) : st.es.lastS;
        st.ea.lastS = sA ? p0 + 31 - clz32(sA) : st.ea.lastS;
    }
    st.x_prev = st.x_cur;
    st.x_cur = st.x_nxt;
}

// Phase 2: B_m[w] from the neighbouring anchors, window tests, seed machines, state rotation.
template <class Sink>
RB_HD void lane_phase2(Sink& sk, const LaneCfg& cfg, LaneState& st, const PlaneWord* cw, const IterCtx& it,
                       uint32_t a_m2, uint32_t a_m1, uint32_t a_p1, uint32_t a_p2, int machines_on) {
    if (!it.slow && machines_on) {
        lane_phase2_fast(sk, cfg, st, cw, it, a_m2, a_m1, a_p1, a_p2);
        return;
    }
    if (cfg.motif) {
        const PlaneWord o = cw[it.w];
        const uint32_t x = st.x_cur;
        const uint32_t b = x | a_m2 | a_m1 | a_p1 | a_p2;
        uint32_t cand;
        const uint32_t passS = ~fail_ge2(x, fsl(st.x_prev, x, 1), st.cs, cand) & o.v;
        const uint32_t passA = ~fail_ge3(b, st.ca) & o.v;
        uint32_t sS, eS, sSp, sA, eA, sAp;
        ev_step(passS, st.es, sS, eS, sSp);
        ev_step(passA, st.ea, sA, eA, sAp);
        (void)smear_step(cfg, st, sA, sAp);
        if (machines_on) {
            {
                const uint32_t prev_v31 = cw[it.w - 1].v >> 31;
                const int p0 = 32 * it.w;
                int pv = (int)prev_v31;
                for (int i = 0; i < 32; ++i) {
                    const int p = p0 + i;
                    if (p >= it.L) break;
                    const int nbit = (o.n >> i) & 1, vbit = (o.v >> i) & 1, xbit = (x >> i) & 1;
                    // perfect machine, parse_perfect_shiftxor.cpp:175-208
                    if (nbit) {
                        if (st.pst != -1) {
                            if (p - st.pst >= cfg.cutPN) emit_perfect(sk, it, cfg, st.pst, p, p);
                            st.pst = -1;
                        }
                        st.sync |= SYNC_P;
                    } else if (xbit) {
                        if (st.pst == -1) st.pst = p;
                    } else {
                        if (st.pst != -1 && p - st.pst >= cfg.cutP) emit_perfect(sk, it, cfg, st.pst, p, p);
                        st.pst = -1;
                        st.sync |= SYNC_P;
                    }
                    win_slow_bit(sk, it, cfg, STREAM_S, p, nbit, vbit, (passS >> i) & 1, st.S, st.zS, st.sync, SYNC_S, pv);
                    win_slow_bit(sk, it, cfg, STREAM_A, p, nbit, vbit, (passA >> i) & 1, st.A, st.zA, st.sync, SYNC_A, pv);
                    pv = vbit;
                    // the rest of the word is N (or padding): after one N every machine is idle and further Ns are no-ops
                    if (nbit && (o.n >> i) == (0xFFFFFFFFu >> i)) break;
                }
            }
        }
    }
    st.x_prev = st.x_cur;
    st.x_cur = st.x_nxt;
}

// Tail flush of a motif lane after the last word of the contig.
template <class Sink>
RB_HD void lane_tail(Sink& sk, const LaneCfg& cfg, LaneState& st, int L) {
    if (!cfg.motif) return;
    // parse_perfect_shiftxor.cpp:213-223: the run that reaches the end is reported with end = L-1
    if (st.pst != -1 && (L - 1) - st.pst >= cfg.cutP) sk.rec(STREAM_P, st.pst, L - 1, cfg.s, 0, cfg.s << 2);
    win_tail(sk, cfg, STREAM_S, L, st.S);
    win_tail(sk, cfg, STREAM_A, L, st.A);
}

// State of a lane that starts at word q. q == 0 is the true start of the contig (exact); any other q is a cold
// start whose state becomes exact once the sync bits are set.
RB_HD void lane_init(const LaneCfg& cfg, LaneState& st, const PlaneWord* cw, int q) {
    st.x_prev = 0u; st.x_cur = 0u; st.x_nxt = 0u; st.lenL = 0;
    win_carry_init(st.cs, 0u);
    win_carry_init(st.ca, 0u);
    st.pst = -1;
    st.S.cur = st.S.ls = st.S.le = -1;
    st.A = st.S;
    // nothing before the contig start: "failing" windows; a cold start must not invent a run of failing windows
    st.es.r2 = st.es.r4 = st.es.r8 = (q == 0) ? 0xFFFFFFFFu : 0u;
    st.es.P = (q == 0) ? 0u : 0xFFFFFFFFu;
    st.es.S = 0u;
    st.es.lastS = -1;
    st.ea = st.es;
    for (int i = 0; i < 7; ++i) st.sm[i] = 0u;
    st.xc.h = st.xc.l = 0u; st.xc.idx = -0x40000000;
    st.sync = (q == 0) ? SYNC_ALL : 0;
    st.zS = st.zA = 0;
    if (cfg.s == 0) return;
    st.x_cur = x_word(cw, q, cfg.s);
    if (q > 0) {
        st.x_prev = x_word(cw, q - 1, cfg.s);
        win_carry_init(st.cs, ~st.x_prev);
        win_carry_init(st.ca, ~st.x_prev);
    }
}

// ---- warm-up across N runs ---------------------------------------------------------------------------------------
// A word is "full-N" when all its 32 positions are N (or padding). Inside a long N run every match word is all ones
// (N is code 00, fasta_utils.cpp:109-113), no window is evaluated and every machine is idle, so a warming-up lane
// may jump over such words: only the anchor-view run length grows. Word w may be skipped when words w-1 .. w+guard
// are all full-N (guard covers the largest shift, so X_s of the skipped word and of its neighbours is all ones).
RB_HD int full_n(const PlaneWord* cw, int w) { return cw[w].n == 0xFFFFFFFFu; }
RB_HD void lane_skip(const LaneCfg& cfg, LaneState& st, const PlaneWord* cw, int w_new, int k) {
    if (cfg.s == 0) return;
    const long long l = (long long)st.lenL + 32ll * k;
    st.lenL = l > LEN_SAT ? LEN_SAT : (int)l;
    st.x_prev = x_word(cw, w_new - 1, cfg.s);
    st.x_cur = x_word(cw, w_new, cfg.s);
}
// words a warm-up spends in front of the run of full-N words that ends right before the chunk (nb = first word of
// that run, nb == w0 when there is none): start H words before the run when it is long, else H words before w0.
static const int NRUN_MIN_WORDS = 16;
RB_HD int warmup_start(int w0, int nb, int H) {
    const int anchor = (w0 - nb >= NRUN_MIN_WORDS) ? nb : w0;
    const int q = anchor - H;
    return q > 0 ? q : 0;
}
RB_HD int warmup_anchor_words(int q, int H) { return (q == 0) ? 0 : (H / 2 > 2 ? H / 2 : 2); }

}  // namespace rb
#endif
