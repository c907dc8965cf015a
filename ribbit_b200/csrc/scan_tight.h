// ribbit-b200: per-lane logic of the TIGHT LOOP of the scan kernel — consecutive fast words of an item, the bulk of every
// contig. Same mapping and same results as the general path (scan_core.h: lane = shift, fast-word logic of
// lane_phase1_fast / lane_phase2_fast), written for a minimal instruction stream:
//   * all carries of the bit-sliced window counters, of the component start / end detection and of the keep filter stay
//     in registers in their funnel-shift-ready form (TightState); nothing is recomputed from the previous word;
//   * no per-lane emission loops: a lane's candidates leave as mask entries (merge_core.h) that the ordered compaction
//     expands; the consumer's cutoff is exact in the bit-parallel filter for small motif sizes, the few survivors of the
//     other lanes are checked bit by bit in a rarely taken region;
//   * anchors (parse_anchored_shiftxor.cpp:20-56) in straight-line code; the run-length bound "< 2s" is only evaluated
//     when a lane of the item sees a run that could reach it (bands whose smallest shift is >= 17: a half word of ones;
//     the band with the small shifts: exact edge logic), which is rare outside repeats;
//   * the substitution stream's event logic is skipped while no lane of the item has a passing window.
// Shared with the CPU warp emulator (tests/emu/emu_scan.cpp), which runs these functions lane by lane.
#ifndef RB_SCAN_TIGHT_H
#define RB_SCAN_TIGHT_H

#include "scan_core.h"

namespace rb {

struct TightCfg {
    int s, sh;           // shift, s & 31
    int K2;              // 2 s (anchor run-length bound)
    int exactA;          // the bit-parallel anchored keep filter is exact for this lane (cutA <= SMEAR_MAX; LARGE items: always)
    uint32_t amask;      // ~0 for lanes that own a shift, 0 for idle lanes (their anchors must read as 0)
    uint32_t mmask;      // ~0 for motif lanes
    int d0, d1, d2, d3;  // anchored keep filter: shifts that smear an S bit over the min(cutoff, 31) positions above it
    int d4;              //   inside its word, levels 0-3 (SMALL items; MID: 1, 2, 4, 8) and 4 (LARGE: unused)
    int cut1;            // anchored cutoff + 1
    int e0, e1, e2, e3, e4;  // doubling shifts of "K2 ones in a row" (SMALL items)
};

struct TightState {
    uint32_t xp, xc;       // X_s of words w-1, w
    uint32_t bh, bl;       // h / l of the plane word that was the "b" operand of X_s[w]
    int lenL;              // SMALL: anchor-view run length ending at the end of word w-1
    WinCarryS cs;
    WinCarryA ca;
    EvCarry es, ea;
    uint32_t sm[7];
    uint32_t p8, p16;      // MID: previous word of "8 / 16 ones of X_s in a row end here" (perfect-run prefilter)
};

RB_HD TightCfg make_tight_cfg(const LaneCfg& c, bool large = false) {  // large: the item runs tight_run<.., TIER_LARGE>
    TightCfg t;
    t.s = c.s; t.sh = c.s & 31; t.K2 = 2 * c.s;
    t.exactA = 1;  // the tight loop's anchored keep filter is exact for every cutoff (tight_events_A)
    t.amask = c.s ? 0xFFFFFFFFu : 0u;
    t.mmask = c.motif ? 0xFFFFFFFFu : 0u;
    {
        const int n = c.cutA < 31 ? c.cutA : 31;
        int reach = 1, d[5];
        for (int i = 0; i < 5; ++i) {
            int sh = n - reach < reach ? n - reach : reach;
            sh = sh < 0 ? 0 : sh;
            d[i] = sh;
            reach += sh;
        }
        t.d0 = d[0]; t.d1 = d[1]; t.d2 = d[2]; t.d3 = d[3]; t.d4 = d[4];
    }
    t.cut1 = c.cutA + 1;
    (void)large;
    // five doubling steps reach min(K2, 32) ones in a row
    int k = 1, e[5];
    for (int i = 0; i < 5; ++i) {
        int sh = t.K2 - k;
        sh = sh < k ? sh : k;
        sh = sh < 0 ? 0 : sh;
        e[i] = sh;
        k += sh;
    }
    t.e0 = e[0]; t.e1 = e[1]; t.e2 = e[2]; t.e3 = e[3]; t.e4 = e[4];
    return t;
}

RB_HD void tight_enter(const LaneCfg& cfg, const LaneState& st, TightState& t) {
    t.xp = st.x_prev; t.xc = st.x_cur; t.bh = st.xc.h; t.bl = st.xc.l; t.lenL = st.lenL;
    t.cs = st.cs; t.ca = st.ca; t.es = st.es; t.ea = st.ea;
    for (int i = 0; i < 7; ++i) t.sm[i] = st.sm[i];
    // 8 / 16 ones in a row ending at t, from the previous word alone: exact in the top bits, which is all that is read
    uint32_t a = t.xp & (t.xp << 1);
    a &= a << 2; a &= a << 4;
    t.p8 = a; t.p16 = a & (a << 8);
    if (!cfg.motif) {
        // the general path keeps no event state for shifts that are not motif sizes: theirs reads "nothing ever passed"
        t.es.P = 0u; t.es.S = 0u; t.es.r2 = t.es.r4 = t.es.r8 = 0xFFFFFFFFu; t.es.lastS = -1;
        t.ea = t.es;
        for (int i = 0; i < 7; ++i) t.sm[i] = 0u;
    }
}
RB_HD void tight_leave(const TightState& t, LaneState& st) {
    st.x_prev = t.xp; st.x_cur = t.xc; st.xc.h = t.bh; st.xc.l = t.bl; st.lenL = t.lenL;
    st.cs = t.cs; st.ca = t.ca; st.es = t.es; st.ea = t.ea;
    for (int i = 0; i < 7; ++i) st.sm[i] = t.sm[i];
}

// ---- phase A: X_s[w+1] and the anchor word A_s[w] ---------------------------------------------------------------------
// Items come in three kinds by their smallest shift s_min (uniform over a warp):
//   TIER_SMALL (s_min <= 15)  runs inside a word can reach 2s: exact edge logic and in-word doubling in every step
//   TIER_MID   (s_min <= 47)  a run of >= 2s >= 32 positions that touches word w covers an aligned half word of w-1, w
//                             or w+1: the straight-line anchors are exact unless some lane sees such a half word
//   TIER_LARGE (s_min >= 48)  2s >= 96: such a run covers all of word w-1, w or w+1
// When a lane of the item raises its flag the item takes tight_anchor_rare for that word.
enum : int { TIER_SMALL = 0, TIER_MID = 1, TIER_LARGE = 2 };
RB_HD int tight_tier(int s_min) { return s_min <= 15 ? TIER_SMALL : (s_min <= 47 ? TIER_MID : TIER_LARGE); }

// oh / ol = h / l of plane word w+1 (the same for every lane), bh / bl = h / l of plane word w + 1 + (s >> 5) + 1.
// Returns the straight-line anchor word; `sus` = this lane's flag for word w+1 (SMALL: for words w, w+1).
template <int TIER>
RB_HD uint32_t tight_phaseA(const TightCfg& c, TightState& t, uint32_t oh, uint32_t ol, uint32_t bh, uint32_t bl, uint32_t& xn,
                            uint32_t& l1, bool& sus) {
    xn = ~((oh ^ fsr(t.bh, bh, c.sh)) | (ol ^ fsr(t.bl, bl, c.sh)));
    t.bh = bh; t.bl = bl;
    const uint32_t x = t.xc;
    l1 = fsl(t.xp, x, 1);
    const uint32_t l2 = fsl(t.xp, x, 2), r1 = fsr(x, xn, 1), r2 = fsr(x, xn, 2);
    // positions of X_s that lie in a run of at least 3 (anchor_size, parse_anchored_shiftxor.cpp:44)
    uint32_t a = x & ((l1 & (l2 | r1)) | (r1 & r2)) & c.amask;
    if (TIER == TIER_LARGE) {
        sus = c.s != 0 && xn == 0xFFFFFFFFu;
        return a;
    }
    if (TIER == TIER_MID) {
        const uint32_t u = ~xn;
        sus = c.s != 0 && (((u - 0x00010001u) & xn & 0x80008000u) != 0u);  // a half word of xn is all ones
        return a;
    }
    sus = c.s != 0 && ((x == 0xFFFFFFFFu) | (xn == 0xFFFFFFFFu));  // idle lanes (shift 0: all ones) never raise it
    // runs that touch a word edge: their full length is known from the neighbours
    const int lead = ctz32(~x), trail = clz32(~x), leadn = ctz32(~xn);
    a &= (t.lenL + lead >= c.K2) ? ~lowmask(lead) : 0xFFFFFFFFu;
    a &= (trail + leadn >= c.K2) ? lowmask(32 - trail) : 0xFFFFFFFFu;
    // a run inside the word can be too long as well: K2 ones in a row by doubling (32 ones in a row cannot occur when
    // !sus, so lanes with K2 > 30 end with e == 0)
    uint32_t e = x;
    e &= e >> c.e0; e &= e >> c.e1; e &= e >> c.e2; e &= e >> c.e3; e &= e >> c.e4;
    if (e) {
        uint32_t d = e;
        for (int k = 1; k < c.K2;) { const int sh = (k < c.K2 - k) ? k : c.K2 - k; d |= d << sh; k += sh; }
        a &= ~d;
    }
    t.lenL = trail;  // meaningless when sus: tight_anchor_rare sets it
    return a;
}

// Length of the run of ones of X_s that ends at the end of word w-1 (t.xp = X_s[w-1]), looked up in the planes when that
// word is all ones; saturates above K2 + 32. Used when the loop is left (items that do not carry the run length).
RB_HD int tight_lenL_lookup(const TightCfg& c, const TightState& t, const PlaneWord* cw, int w) {
    int lenL = clz32(~t.xp);
    if (lenL == 32)
        for (int k = w - 2; k >= 0; --k) {
            const int tr = clz32(~x_word(cw, k, c.s));
            lenL += tr;
            if (tr < 32 || lenL >= c.K2 + 32) break;
        }
    return lenL;
}

// Exact anchor word for a word some lane of the item flagged. a = the straight-line result of tight_phaseA, lenL_in =
// t.lenL before that call, prev_rare = the previous word took this path too (then t.lenL is the carried run length; if
// not, the previous word raised no flag, so X_s[w-1] has no all-ones (half) word and its trailing ones are the run).
// MID / LARGE: while neither this word nor the next is all ones the runs that touch an edge of the word are judged with
// their known lengths; whole words of ones take the general anchor_word.
template <int TIER>
RB_HD uint32_t tight_anchor_rare(const TightCfg& c, TightState& t, const PlaneWord* cw, int w, int L, uint32_t xn, uint32_t a,
                                 int lenL_in, bool prev_rare, bool& full) {
    const uint32_t x = t.xc;
    full = c.s != 0 && ((x == 0xFFFFFFFFu) | (xn == 0xFFFFFFFFu));
    if (TIER == TIER_SMALL) {
        if (full) t.lenL = lenL_in;  // tight_phaseA left the trailing ones of the word there
        return a;
    }
    const int lenL = prev_rare ? lenL_in : clz32(~t.xp);
    const int lead = ctz32(~x), trail = clz32(~x), leadn = ctz32(~xn);
    a &= (lenL + lead >= c.K2) ? ~lowmask(lead) : 0xFFFFFFFFu;
    a &= (trail + leadn >= c.K2) ? lowmask(32 - trail) : 0xFFFFFFFFu;
    t.lenL = full ? lenL : trail;
    return a;
}
// second half of the rare path, for the lanes with `full`: t.lenL holds the run length in front of the word
RB_HD uint32_t tight_anchor_full(const TightCfg& c, TightState& t, const PlaneWord* cw, int w, int L, uint32_t xn) {
    return anchor_word(cw, w, L, c.s, t.xc, xn, t.lenL);
}

// ---- phase B: window tests, component events, keep filter ---------------------------------------------------------------
// an = A_{m-2} | A_{m-1} | A_{m+1} | A_{m+2}. Non-motif lanes produce all-zero masks.
template <int TIER>
RB_HD void tight_windows(const TightCfg& c, TightState& t, uint32_t an, uint32_t l1, uint32_t& passS, uint32_t& passA, uint32_t& cand) {
    const uint32_t x = t.xc;
    const uint32_t o2p = t.cs.o2;
    passS = ~fail_ge2(x, l1, t.cs, cand) & c.mmask;
    passA = ~fail_ge3(x | an, t.ca) & c.mmask;
    if (TIER == TIER_LARGE) {
        // every perfect cutoff of a LARGE item is at least 50 (parse_perfect_shiftxor.cpp:193): a run that long which ends
        // inside this word started in an earlier one, so its end is the lowest zero bit of the word, and it covers at
        // least the top 18 bits of the previous word (t.p8 / t.p16 are not maintained)
        cand = (t.xp >= 0xFFFF0000u) ? (~x & (x + 1u)) : 0u;
    } else if (TIER != TIER_SMALL) {
        // every perfect cutoff of these items is at least 18 (parse_perfect_shiftxor.cpp:193): only run ends that follow 16
        // ones can be candidates, which leaves almost nothing for the exact check
        const uint32_t o8 = ~(t.cs.o2 | fsl(o2p, t.cs.o2, 4));   // X_s[t-7..t] all ones
        const uint32_t o16 = o8 & fsl(t.p8, o8, 8);
        cand = ~x & fsl(t.p16, o16, 1);
        t.p8 = o8; t.p16 = o16;
    }
}

struct TightOut {
    uint32_t x, s, e;  // E bits that survived the prefilter, S mask of the word, all E bits
    int last;          // position of the latest S bit in front of the word
};
template <int TIER>
RB_HD void tight_events_A(const TightCfg& c, TightState& t, int p0, uint32_t passA, TightOut& o) {
    uint32_t sA, eA, sAp;
    ev_step(passA, t.ea, sA, eA, sAp);
    // The cutoff is decided from the latest S bit in front of each E bit (scan_core.h keep_by_last): exact for every
    // cutoff, and the loop carries no smear state (t.sm is not maintained: smear_from_last when the loop is left).
    if (TIER == TIER_LARGE) {
        // every motif size of a LARGE item is at least 50: anchored cutoffs of 45 and more, an S bit in the same word is
        // always too close
        o.x = keep_by_last(eA, sA, t.ea.lastS, c.cut1, p0);
    } else {
        uint32_t near = sA << 1;
        if (TIER == TIER_SMALL) { near |= near << c.d0; near |= near << c.d1; near |= near << c.d2; near |= near << c.d3; }
        else { near |= near << 1; near |= near << 2; near |= near << 4; near |= near << 8; }  // (MID: every cutoff is at least 16)
        near |= near << c.d4;
        o.x = keep_by_last_near(eA, sA, near, t.ea.lastS, c.cut1, p0);
    }
    o.e = eA; o.s = sA; o.last = t.ea.lastS;
    t.ea.lastS = sA ? p0 + 31 - clz32(sA) : t.ea.lastS;
    (void)sAp;
}
RB_HD void tight_events_S(TightState& t, int p0, uint32_t passS, TightOut& o) {
    uint32_t sS, eS, sSp;
    ev_step(passS, t.es, sS, eS, sSp);
    // a substitution component whose S bit is 9 or 10 back has length 8 or 9: below every cutoff
    o.x = eS & ~(fsl(sSp, sS, 9) | fsl(sSp, sS, 10));
    o.e = eS; o.s = sS; o.last = t.es.lastS;
    t.es.lastS = sS ? p0 + 31 - clz32(sS) : t.es.lastS;
}
// rotation at the end of a step
RB_HD void tight_rotate(TightState& t, uint32_t xn) { t.xp = t.xc; t.xc = xn; }

}  // namespace rb
#endif
