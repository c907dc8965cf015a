// TEST INFRASTRUCTURE — CPU emulation of the scan + merge kernels, lane by lane, built from the very headers the
// CUDA kernels include (ribbit_b200/csrc/{scan_core,merge_core,layout}.h). It lets the CPU test-suite check the
// kernel logic against the oracle where no GPU exists. It is not a product path: the product has no CPU fallback.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fasta_core.h"
#include "layout.h"
#include "motif_core.h"
#include "merge_core.h"
#include "scan_core.h"
#include "scan_tight.h"

using namespace rb;

namespace {

struct ItemOut {
    std::vector<Rec> raw;  // the three streams interleaved in arrival order, as in the kernel's raw record pool
};

struct EmuSink {
    ItemOut* out;
    int nslots;
    int dmax[3];
    void rec(int stream, int start, int end, int mlen, int flags, int key) {
        Rec r; r.start = start; r.end = end; r.mflags = mlen | (flags << 16) | (stream << REC_STREAM_SHIFT); r.key = key;
        out->raw.push_back(r);
        ++nslots;
    }
    void entry(int stream, int mlen, uint32_t mask, uint32_t smask, int last) {
        out->raw.push_back(make_entry(stream, mlen, mask, smask, last));
        ++nslots;
    }
    void dropped_mask(int stream, uint32_t el) { if (el) dmax[stream] = std::max(dmax[stream], 32 - clz32(el)); }
    void reset(ItemOut* o) { out = o; nslots = 0; dmax[0] = dmax[1] = dmax[2] = 0; }
};

void pack(const char* seq, int64_t L, int guard, std::vector<PlaneWord>& pw) {
    const int64_t nw = (L + 31) / 32;
    pw.assign((size_t)(1 + nw + guard), PlaneWord{0u, 0u, 0xFFFFFFFFu, 0u});
    for (int64_t w = 0; w < nw; ++w) {
        PlaneWord o{0u, 0u, 0u, 0u};
        for (int i = 0; i < 32; ++i) {
            const int64_t p = 32 * w + i;
            if (p >= L) { o.n |= 1u << i; continue; }
            switch (seq[p]) {
                case 'A': case 'a': break;
                case 'C': case 'c': o.l |= 1u << i; break;
                case 'G': case 'g': o.h |= 1u << i; break;
                case 'T': case 't': o.h |= 1u << i; o.l |= 1u << i; break;
                default: o.n |= 1u << i; break;
            }
        }
        pw[(size_t)(1 + w)] = o;
    }
    int run = 0;  // consecutive non-N bases ending at p
    for (int64_t p = 0; p < L; ++p) {
        const bool isn = (pw[(size_t)(1 + p / 32)].n >> (p & 31)) & 1;
        run = isn ? 0 : run + 1;
        if (run >= 8) pw[(size_t)(1 + p / 32)].v |= 1u << (p & 31);
    }
}


// The kernel's tight_run for one item: the lanes step through scan_tight.h in lockstep, the warp-level operations
// (votes, shuffles, ballots, reductions) are done across the lane arrays.
template <int TIER>
long run_tight(const BandLayout& lay, const LaneCfg* cfg, LaneState* st, const PlaneWord* cw, int& w, int wend, int L, ItemOut& io,
               std::vector<Meta>& meta) {
    const int bw = lay.bw;
    TightState ts[32];
    TightCfg tc[32];
    for (int j = 0; j < bw; ++j) { tight_enter(cfg[j], st[j], ts[j]); tc[j] = make_tight_cfg(cfg[j], TIER == TIER_LARGE); }
    uint32_t vprev = v_eff(cw, w - 1), vcur = cw[w].v;
    bool susp = true, susc = true, prev_rare = true;
    int zc = 0;
    long steps = 0;
    const int w_in = w;
    while (w < wend && (vprev & vcur) == 0xFFFFFFFFu) {
        uint32_t a[32], xn[32], l1[32], pair[32];
        int lenL0[32];
        bool rare = false;
        for (int j = 0; j < bw; ++j) {
            const PlaneWord o = cw[w + 1], b = cw[w + 1 + (tc[j].s >> 5) + 1];
            bool sus;
            lenL0[j] = ts[j].lenL;
            a[j] = tight_phaseA<TIER>(tc[j], ts[j], o.h, o.l, b.h, b.l, xn[j], l1[j], sus);
            rare |= sus;
        }
        if (TIER != TIER_SMALL) { const bool r3 = rare | susc | susp; susp = susc; susc = rare; rare = r3; }
        if (rare)
            for (int j = 0; j < bw; ++j) {
                bool full;
                a[j] = tight_anchor_rare<TIER>(tc[j], ts[j], cw, w, L, xn[j], a[j], lenL0[j], prev_rare, full);
                if (full) a[j] = tight_anchor_full(tc[j], ts[j], cw, w, L, xn[j]);
            }
        prev_rare = rare;
        for (int j = 0; j < bw; ++j) pair[j] = a[j] | a[j + 1 < bw ? j + 1 : j];
        uint32_t passS[32], passA[32], cand[32];
        TightOut oa[32], os[32];
        for (int j = 0; j < bw; ++j) {
            const uint32_t an = pair[j >= 2 ? j - 2 : j] | pair[j + 1 < bw ? j + 1 : j];
            tight_windows<TIER>(tc[j], ts[j], an, l1[j], passS[j], passA[j], cand[j]);
            tight_events_A<TIER>(tc[j], ts[j], 32 * w, passA[j], oa[j]);
        }
        bool anyPS = false;
        for (int j = 0; j < bw; ++j) { cand[j] &= tc[j].mmask; anyPS |= (cand[j] | passS[j]) != 0u; }
        const bool runS = anyPS || zc < 2;
        for (int j = 0; j < bw; ++j) {
            os[j].x = os[j].s = os[j].e = 0u; os[j].last = 0;
            if (runS) tight_events_S(ts[j], 32 * w, passS[j], os[j]);
            // exact check of the survivors the bit-parallel filters do not decide
            if (!tc[j].exactA && oa[j].x) oa[j].x = kept_exact(cut_anch(tc[j].s), 32 * w, oa[j].x, oa[j].s, oa[j].last);
            if (tc[j].s > 30 && os[j].x) os[j].x = kept_exact(cut_subst(tc[j].s), 32 * w, os[j].x, os[j].s, os[j].last);
            if (cand[j]) cand[j] = kept_exact_perfect(cw, w, tc[j].s, cut_perfect(tc[j].s), ts[j].xc, ts[j].xc & ~l1[j], cand[j]);
        }
        if (runS) zc = anyPS ? 0 : zc + 1;
        const uint32_t off = (uint32_t)io.raw.size();
        uint32_t elA = 0u, elS = 0u;
        for (int j = 0; j < bw; ++j) if (oa[j].x) io.raw.push_back(make_entry(STREAM_A, tc[j].s, oa[j].x, oa[j].s, oa[j].last));
        for (int j = 0; j < bw; ++j) if (cand[j]) io.raw.push_back(make_entry(STREAM_P, tc[j].s, cand[j], ts[j].xc & ~l1[j], 0));
        for (int j = 0; j < bw; ++j) if (os[j].x) io.raw.push_back(make_entry(STREAM_S, tc[j].s, os[j].x, os[j].s, os[j].last));
        for (int j = 0; j < bw; ++j) { elA |= oa[j].e & ~oa[j].x; elS |= os[j].e & ~os[j].x; }
        meta[w] = make_meta((int)(io.raw.size() - off), elS ? 32 - clz32(elS) : 0, elA ? 32 - clz32(elA) : 0, 0, off);
        for (int j = 0; j < bw; ++j) tight_rotate(ts[j], xn[j]);
        vprev = vcur; vcur = cw[w + 1].v;
        ++w; ++steps;
    }
    for (int j = 0; j < bw; ++j) {
        if (TIER != TIER_SMALL && tc[j].s && !prev_rare) ts[j].lenL = tight_lenL_lookup(tc[j], ts[j], cw, w);
        tight_leave(ts[j], st[j]);
        if (w != w_in && cfg[j].motif) smear_from_last(cfg[j], st[j], w);
        if (w != w_in && cfg[j].s) st[j].xc.idx = w + (cfg[j].s >> 5) + 1;
    }
    return steps;
}

}  // namespace

extern "C" {

// Property check of the tight loop's keep filter (scan_core.h keep_by_last / keep_by_last_near, smear_from_last) on random
// event words:
//  (0) keep_by_last_near with the in-word smear of make_tight_cfg == kept_exact for every motif size 2..1000;
//  (1) keep_by_last == kept_exact for every cutoff >= 32;
//  (2) a lane that goes back to the smear network with the state smear_from_last builds gets the answers the network
//      gives when it ran all along (cutoffs <= SMEAR_MAX).
// Returns the number of disagreements.
int emu_keep_filter_check(uint64_t seed, int rounds) {
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    auto sparse = [&](int k) { uint32_t v = 0xFFFFFFFFu; for (int i = 0; i < k; ++i) v &= (uint32_t)rnd(); return v; };
    int bad = 0;
    for (int r = 0; r < rounds; ++r) {
        const int m = 2 + (int)(rnd() % 999);
        const LaneCfg cfg = make_lane_cfg(m, 2, 1000, 1, 1002, 2, 1000);
        const TightCfg tc = make_tight_cfg(cfg);
        if (cfg.cutA >= 16 && (tc.d0 != 1 || tc.d1 != 2 || tc.d2 != 4 || tc.d3 != 8)) ++bad;  // what the MID loop hard-codes
        int lastS = -1;
        for (int w = 0; w < 24; ++w) {
            const uint32_t S = w ? sparse(2 + (int)(rnd() % 4)) : 1u, E = w ? sparse(1 + (int)(rnd() % 4)) : 0u;
            uint32_t near = S << 1;
            near |= near << tc.d0; near |= near << tc.d1; near |= near << tc.d2; near |= near << tc.d3; near |= near << tc.d4;
            if (keep_by_last_near(E, S, near, lastS, tc.cut1, 32 * w) != kept_exact(cfg.cutA, 32 * w, E, S, lastS)) ++bad;
            if (S) lastS = 32 * w + 31 - clz32(S);
        }
    }
    for (int r = 0; r < rounds; ++r) {
        const int m = 36 + (int)(rnd() % 965);  // cut_anch(m) >= 32
        const LaneCfg cfg = make_lane_cfg(m, 2, 1000, 1, 1002, 2, 1000);
        const int cut = cfg.cutA;
        LaneState full, rebuilt;
        for (int i = 0; i < 7; ++i) full.sm[i] = 0u;
        int lastS = -1;
        uint32_t Sprev = 0u;
        const int nwords = 40, sw = 8 + (int)(rnd() % 24);  // the second lane switches to the network at word sw
        for (int w = 0; w < nwords; ++w) {
            // (an E bit always has an S bit in front of it: the first word starts a component and ends none)
            const uint32_t S = w ? sparse(3 + (int)(rnd() % 3)) : 1u, E = w ? sparse(2 + (int)(rnd() % 3)) : 0u;
            const int p0 = 32 * w;
            const uint32_t exact = kept_exact(cut, p0, E, S, lastS);
            if (keep_by_last(E, S, lastS, cut + 1, p0) != exact) ++bad;
            const uint32_t vfull = E & ~smear_step(cfg, full, S, Sprev);
            if (cut <= SMEAR_MAX && vfull != exact) ++bad;
            if (w == sw) { rebuilt.ea.lastS = lastS; smear_from_last(cfg, rebuilt, w); }
            if (w >= sw && cut <= SMEAR_MAX && (E & ~smear_step(cfg, rebuilt, S, Sprev)) != exact) ++bad;
            if (S) lastS = p0 + 31 - clz32(S);
            Sprev = S;
        }
    }
    return bad;
}

// Returns the three ordered streams concatenated per stream: out[stream] malloc'ed arrays of Rec, counts in n[3].
// restarts (optional) receives the number of warm-up restarts.
int emu_scan(const char* seq, int64_t L, int m_lo, int m_hi, int chunk_words, int warm0, Rec** out, int64_t* n,
             int64_t* restarts, int64_t* skips, int64_t* replays) {
    const BandLayout lay = make_layout(m_lo, m_hi);
    std::vector<PlaneWord> planes;
    pack(seq, L, lay.guard, planes);
    const PlaneWord* cw = planes.data() + 1;
    const int nw = (int)((L + 31) / 32);
    if (restarts) *restarts = 0;
    if (skips) *skips = 0;
    if (replays) *replays = 0;

    std::vector<Chunk> chunks;
    if (nw == 0) chunks.push_back(Chunk{0, 0, 0, 1});
    for (int w = 0; w < nw; w += chunk_words) chunks.push_back(Chunk{0, w, std::min(nw, w + chunk_words), w + chunk_words >= nw});

    std::vector<std::vector<Meta>> meta(lay.nbands, std::vector<Meta>((size_t)nw + 1, Meta{0u, 0u}));
    std::vector<ItemOut> items(chunks.size() * lay.nbands);
    long tight_steps = 0;

    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk ch = chunks[ci];
        for (int band = 0; band < lay.nbands; ++band) {
            ItemOut& io = items[ci * lay.nbands + band];
            LaneCfg cfg[32];
            LaneState st[32];
            for (int j = 0; j < lay.bw; ++j) { cfg[j] = band_lane_cfg(lay, band, j); lane_cfg_set_contig(cfg[j], (int)L); }
            int H = warm0;
            const bool tight = getenv("RB_EMU_NO_TIGHT") == nullptr;
            const int e0 = ch.w0;  // first emitting word
            int we = e0;           // word at which the lane state must be complete (end of the current warm-up)
            int force = 0;         // the warm-up rebuilds the reference machines: process its words bit-serially
            int nb = we;           // first word of the run of full-N words that ends right before it
            while (nb > 0 && full_n(cw, nb - 1)) --nb;
            const int nb0 = nb;
            for (;;) {
                const int q = warmup_start(we, nb, H);
                const int Ha = warmup_anchor_words(q, H);
                for (int j = 0; j < lay.bw; ++j) lane_init(cfg[j], st[j], cw, q);
                bool restart = false, replay = false;
                int prev_slow = 1, fastrun = q == 0 ? 3 : 0;
                for (int w = q; w < ch.w1 && !restart;) {
                    const int lim = std::max(we, e0) - 2;  // no word is emitted before max(we, e0)
                    if (w >= q + Ha && w - 1 >= nb0) {
                        // inside the N run that ends right before the chunk (words nb0 .. e0-1 are all N): one jump
                        const int k = std::min(e0 - lay.guard, lim) - w;
                        if (k > 0) {
                            w += k;
                            for (int j = 0; j < lay.bw; ++j) lane_skip(cfg[j], st[j], cw, w, k);
                            if (skips) ++*skips;
                            continue;
                        }
                    }
                    if (w >= q + Ha && w < lim) {  // warming up inside an N run: jump
                        int r = 0;
                        while (r < lay.bw && w - 1 + r < nw + lay.guard && full_n(cw, w - 1 + r)) ++r;
                        const int k = std::min(r - lay.guard - 1, lim - w);
                        if (k > 0) {
                            w += k;
                            for (int j = 0; j < lay.bw; ++j) lane_skip(cfg[j], st[j], cw, w, k);
                            if (skips) ++*skips;
                            continue;
                        }
                    }
                    const int slow = (force && w < we) || !word_is_fast(cw, w, nw);
                    if (slow && !prev_slow) {
                        // fast -> slow: the reference machines' state comes from the bit-parallel view of the previous word
                        // when every lane's word holds nine failing windows in a row ...
                        bool ok = w >= q + Ha;  // the machines run in this word (not an anchors-only warm-up word)
                        SlowEntry se[32];
                        for (int j = 0; j < lay.bw; ++j) ok = lane_to_slow(cfg[j], st[j], cw, w, se[j]) && ok;
                        if (ok) {
                            for (int j = 0; j < lay.bw; ++j) lane_enter_slow(st[j], se[j]);
                            prev_slow = 1;
                        } else {
                            // ... else it is rebuilt by a warm-up that ends here
                            we = w; nb = w; H = warm0; force = 1; replay = true;
                            if (replays) ++*replays;
                            break;
                        }
                    }
                    uint32_t a[32 + 4] = {0};
                    for (int j = 0; j < lay.bw; ++j)
                        a[j + 2] = slow ? lane_phase1(cfg[j], st[j], cw, w, (int)L) : lane_phase1_fast(cfg[j], st[j], cw, w, (int)L);
                    fastrun = slow ? 0 : std::min(fastrun + 1, 4);
                    if (q > 0 && w == q + Ha - 2)
                        for (int j = 0; j < lay.bw; ++j) if (cfg[j].s && !(st[j].sync & SYNC_X)) restart = true;
                    if (q > 0 && w == we)
                        for (int j = 0; j < lay.bw; ++j) if (cfg[j].motif && (st[j].sync & SYNC_ALL) != SYNC_ALL) restart = true;
                    if (restart) break;
                    IterCtx it; it.w = w; it.L = (int)L; it.emit_on = w >= we && w >= e0; it.slow = slow; it.prev_slow = prev_slow; it.fastrun = fastrun;
                    EmuSink sk; sk.reset(&io);
                    const uint32_t off = (uint32_t)io.raw.size();
                    for (int j = 0; j < lay.bw; ++j)
                        lane_phase2(sk, cfg[j], st[j], cw, it, j >= 2 ? a[j] : 0u, j >= 1 ? a[j + 1] : 0u,
                                    j + 1 < lay.bw ? a[j + 3] : 0u, j + 2 < lay.bw ? a[j + 4] : 0u, w >= q + Ha);
                    if (it.emit_on) meta[band][w] = make_meta(sk.nslots, sk.dmax[1], sk.dmax[2], it.slow, off);
                    prev_slow = slow;
                    ++w;
                    // tight path of the kernel (scan_tight.h): consecutive fast, emitting words. One item at a time here; the
                    // kernel couples the items of a warp only in WHEN it enters / leaves the loop and takes the rare paths
                    if (tight && w > we && w >= e0 && fastrun >= 4 && !prev_slow) {
                        const int wend = std::min(std::min(ch.w1, nw - 1), (((int)L - lay.s_hi) >> 5) - 1);
                        if (w < wend && (v_eff(cw, w - 1) & cw[w].v) == 0xFFFFFFFFu) {
                            const int tier = tight_tier(band_m0(lay, band) - 2);
                            tight_steps += tier == TIER_SMALL ? run_tight<TIER_SMALL>(lay, cfg, st, cw, w, wend, (int)L, io, meta[band])
                                         : tier == TIER_MID ? run_tight<TIER_MID>(lay, cfg, st, cw, w, wend, (int)L, io, meta[band])
                                                            : run_tight<TIER_LARGE>(lay, cfg, st, cw, w, wend, (int)L, io, meta[band]);
                        }
                    }
                }
                if (replay) continue;
                if (!restart) {
                    if (ch.last) {
                        EmuSink sk; sk.reset(&io);
                        const uint32_t off = (uint32_t)io.raw.size();
                        for (int j = 0; j < lay.bw; ++j) lane_tail(sk, cfg[j], st[j], (int)L);
                        meta[band][nw] = make_meta(sk.nslots, 0, 0, 1, off);
                    }
                    break;
                }
                if (restarts) ++*restarts;
                H = std::min(H * 4, we);
                if (H < 1) H = 1;
            }
        }
    }

    if (getenv("RB_EMU_TRACE")) fprintf(stderr, "emu: %ld tight-loop steps\n", tight_steps);
    // merge (kernels M1-M3): buckets in word order; per bucket a pseudo record, then the records ranked by key
    std::vector<Rec> res[3];
    long long emax[2] = {0, 0};  // running maximum of elided_end_code (exclusive prefix), streams S and A
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk ch = chunks[ci];
        const int wend = ch.last ? ch.w1 + 1 : ch.w1;
        for (int w = ch.w0; w < wend; ++w) {
            // merge_count: entries -> kept masks (written back into the slot), stream sizes, elided maxima
            std::vector<Rec*> slots;
            int tot[3] = {0, 0, 0}, slow = 0, dS = 0, dA = 0;
            for (int b = 0; b < lay.nbands; ++b) {
                const Meta m = meta[b][w];
                Rec* src = items[ci * lay.nbands + b].raw.data() + m.off;
                for (int i = 0; i < meta_slots(m); ++i) {
                    Rec& r = src[i];
                    const int s = rec_stream(r);
                    if (rec_is_entry(r)) {
                        int el;
                        const uint32_t kept = entry_kept_mask(r, w, cw, el);
                        r.start = (int32_t)kept;
                        tot[s] += popc32(kept);
                        if (s == STREAM_S) dS = std::max(dS, el);
                        if (s == STREAM_A) dA = std::max(dA, el);
                    } else {
                        tot[s] += 1;
                    }
                    slots.push_back(&r);
                }
                slow |= meta_slow(m);
                dS = std::max(dS, meta_dmax(m, STREAM_S));
                dA = std::max(dA, meta_dmax(m, STREAM_A));
            }
            size_t base[3];
            for (int s = 0; s < 3; ++s) {
                base[s] = res[s].size();
                const int ps = bucket_has_pseudo(s, slow, tot[s]);
                if (ps) { res[s].push_back(pseudo_rec(w, emax[s - 1] - 1)); base[s] += 1; }
                res[s].resize(base[s] + tot[s]);
            }
            // merge_write: every candidate ranks itself among the candidates of its stream in the bucket
            for (Rec* rp : slots) {
                const Rec r = *rp;
                const int s = rec_stream(r);
                auto rank_of = [&](uint32_t key) {
                    int rank = 0;
                    for (Rec* op : slots) if (rec_stream(*op) == s) rank += slot_count_below(*op, key);
                    return rank;
                };
                if (!rec_is_entry(r)) {
                    res[s][base[s] + rank_of((uint32_t)r.key)] = finalize_rec(r, w);
                    continue;
                }
                for (uint32_t x = (uint32_t)r.start; x; x &= x - 1u) {
                    const int i = ctz32(x);
                    Rec o;
                    entry_interval(r, w, cw, i, o.start, o.end);
                    o.mflags = rec_mlen(r);
                    o.key = 32 * w + i;
                    res[s][base[s] + rank_of(entry_key(i, rec_mlen(r)))] = o;
                }
            }
            emax[0] = std::max<long long>(emax[0], elided_end_code(w, dS));
            emax[1] = std::max<long long>(emax[1], elided_end_code(w, dA));
        }
    }
    for (int stream = 0; stream < 3; ++stream) {
        n[stream] = (int64_t)res[stream].size();
        out[stream] = (Rec*)malloc(std::max<size_t>(1, res[stream].size()) * sizeof(Rec));
        memcpy(out[stream], res[stream].data(), res[stream].size() * sizeof(Rec));
    }
    return 0;
}

void emu_free(void* p) { free(p); }

// K0 on the CPU: the three passes of fasta_kernels.cu with the shared slice arithmetic (fasta_core.h), tiles and slices
// walked in order. bases: capacity nbytes; hdr_pos / hdr_seq: capacity nbytes (one header needs at least one byte).
// totals[0] = sequence bytes, totals[1] = headers.
int emu_fasta(const uint8_t* text, int64_t nbytes, uint8_t* bases, int64_t* hdr_pos, int64_t* hdr_seq, int64_t* totals) {
    const int64_t nt = (nbytes + FT_TILE - 1) / FT_TILE;
    auto load = [&](int64_t at) {
        Slice r;
        const int n = (int)std::max<int64_t>(0, std::min<int64_t>(FT_PER, nbytes - at));
        for (int k = 0; k < 4; ++k) {
            uint32_t x = 0;
            for (int j = 0; j < 4; ++j) if (4 * k + j < n) x |= (uint32_t)text[at + 4 * k + j] << (8 * j);
            r.w[k] = x;
        }
        const uint32_t prev_nl = (at > 0 && at <= nbytes) ? (text[at - 1] == '\n') : 1u;
        slice_masks(r, n, prev_nl);
        return r;
    };
    struct Tile { int pre, known, nh, last; };
    std::vector<Tile> tiles((size_t)nt);
    for (int64_t t = 0; t < nt; ++t) {  // fasta_tile_kernel
        Tile ti{0, 0, 0, LT_NONE};
        int before = -1;
        for (int th = 0; th < FT_THREADS; ++th) {
            const Slice d = load(t * FT_TILE + (int64_t)th * FT_PER);
            const int type = before < 0 ? LT_NONE : (before & 1);
            const int nb = popc32(d.before), na = popc32(slice_seq_after(d));
            ti.pre += type == LT_NONE ? nb : 0;
            ti.known += na + (type == LT_SEQ ? nb : 0);
            ti.nh += popc32(d.hs);
            before = std::max(before, last_control(d, th * FT_PER));
        }
        ti.last = before < 0 ? LT_NONE : (before & 1);
        tiles[(size_t)t] = ti;
    }
    std::vector<int64_t> seq_base((size_t)nt), hdr_base((size_t)nt);
    std::vector<int> entry((size_t)nt);
    int64_t seq = 0, hdr = 0;
    int type = LT_SEQ;
    for (int64_t t = 0; t < nt; ++t) {  // fasta_scan_kernel
        const Tile& v = tiles[(size_t)t];
        seq_base[(size_t)t] = seq; hdr_base[(size_t)t] = hdr; entry[(size_t)t] = type;
        seq += (type == LT_SEQ ? v.pre : 0) + v.known;
        hdr += v.nh;
        if (v.last != LT_NONE) type = v.last;
    }
    totals[0] = seq; totals[1] = hdr;
    for (int64_t t = 0; t < nt; ++t) {  // fasta_strip_kernel
        int before = -1, rank = 0, hrank = 0;
        for (int th = 0; th < FT_THREADS; ++th) {
            const int64_t at = t * FT_TILE + (int64_t)th * FT_PER;
            const Slice d = load(at);
            const int type0 = before < 0 ? entry[(size_t)t] : (before & 1);
            const uint32_t sq = slice_seq(d, type0);
            int r = rank;
            for (int i = 0; i < FT_PER; ++i)
                if ((sq >> i) & 1u) bases[seq_base[(size_t)t] + r++] = (uint8_t)slice_byte(d, i);
            for (uint32_t h = d.hs; h; h &= h - 1u) {
                const int i = ctz32(h);
                const int64_t k = hdr_base[(size_t)t] + hrank + popc32(d.hs & ((1u << i) - 1u));
                hdr_pos[k] = at + i;
                hdr_seq[k] = seq_base[(size_t)t] + rank + popc32(sq & ((1u << i) - 1u));
            }
            rank += popc32(sq);
            hrank += popc32(d.hs);
            before = std::max(before, last_control(d, th * FT_PER));
        }
    }
    return 0;
}

// K7 on the CPU: the kernel's per-row scoring (motif_core.h) and its reduction rule (largest count, then smallest row; no
// scoring row -> 0). seeds = n x {start, end, mlen}; out = n x {row, count}.
int emu_motif_rows(const char* seq, int64_t L, const int32_t* seeds, int64_t n, int32_t* out) {
    std::vector<PlaneWord> planes;
    pack(seq, L, 8, planes);
    const PlaneWord* cw = planes.data() + 1;
    const int nw = (int)((L + 31) / 32);
    for (int64_t i = 0; i < n; ++i) {
        const int start = seeds[3 * i], end = seeds[3 * i + 1], m = seeds[3 * i + 2];
        unsigned long long best = 0;
        for (int row = start; row <= end - m; ++row) {
            const int sc = motif_row_score(cw, nw, start, end, m, row);
            const unsigned long long key = ((unsigned long long)(uint32_t)sc << 32) | (uint32_t)(0x7FFFFFFF - row);
            if (key > best) best = key;
        }
        out[2 * i + 1] = (int32_t)(best >> 32);
        out[2 * i] = (best >> 32) ? (int32_t)(0x7FFFFFFF - (uint32_t)(best & 0xFFFFFFFFull)) : 0;
    }
    return 0;
}
}
