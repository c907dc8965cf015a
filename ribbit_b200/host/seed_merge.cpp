// ribbit-b200 host side: restatement of ribbit's order-dependent seed-list merges (see seed_merge.h).
// Every decision is the reference's, including the ones that look unintended — the seed lists must come
// out entry for entry; the line numbers say where each rule comes from.
#include "seed_merge.h"

#include <algorithm>
#include <cstdlib>
#include <deque>
#include <map>

#include "../csrc/scan_core.h"

namespace rbm {

namespace {

enum Relation { IDENTICAL, NESTED, PARENT, OVERLAP };

// how the new seed [s, e) lies to an existing one [ls, le): the order of the tests is the reference's
inline Relation relate(int s, int e, int ls, int le) {
    if (s == ls && e == le) return IDENTICAL;
    if (ls <= s && e <= le) return NESTED;   // the new seed lies inside the existing one
    if (s <= ls && le <= e) return PARENT;   // the existing seed lies inside the new one
    return OVERLAP;
}

inline void retire(Seed& x) { x.rank = RANK_N; }

// overlap length and merged extent of two overlapping seeds, parse_substitute_shiftxor.cpp:322-346 =
// parse_anchored_shiftxor.cpp:352-376: the repeat of the shorter motif is taken to reach one motif length further
struct Overlap {
    int length, start, end;
};
inline Overlap overlap_of(int s, int e, int m, int ls, int le, int lm) {
    Overlap o;
    const int rend = e + m, lrend = le + lm;
    if (ls < s) {
        if (lm <= m) o.length = (e <= lrend) ? e - s : lrend - s;
        else o.length = (e <= le) ? e - s : le - s;
        o.start = ls; o.end = e;
    } else {
        if (m <= lm) o.length = (le <= rend) ? le - ls : rend - ls;
        else o.length = (le <= e) ? le - ls : e - ls;
        o.start = s; o.end = le;
    }
    return o;
}

// parse_substitute_shiftxor.cpp:34-42, parse_anchored_shiftxor.cpp:133-151: the cursor moves to the first seed that
// starts behind the candidate's end, but never past the last seed
inline int advance_cursor(const SeedList& list, int from, int seed_end) {
    const int n = (int)list.size();
    for (int i = from; i < n; ++i) {
        if (list[(size_t)i].start > seed_end) break;
        if (from == n - 1) break;
        ++from;
    }
    return from;
}

}  // namespace

// ---- perfect ------------------------------------------------------------------------------------------------------
void add_perfect(int seed_start, int seed_end, int mlen, SeedList& perfect, int contig_len) {
    const int seed_rlen = seed_end - seed_start + mlen;
    std::vector<int> drop;  // existing seeds the new one replaces (descending indices)
    for (int i = (int)perfect.size() - 1; i >= 0; --i) {
        const Seed last = perfect[(size_t)i];
        const int last_rlen = last.end - last.start + last.mlen;
        if (last.end < seed_start) break;  // the list is ordered by end (:71)
        if (last.start == seed_start && last.end == seed_end) {  // :74-77
            if (last.mlen < mlen) return;
            drop.push_back(i);
        } else if (last.start <= seed_start && last.end >= seed_end) {  // nested, :80-83
            if (seed_rlen < last.mlen / 3) continue;
            return;
        } else if (seed_start <= last.start && seed_end >= last.end) {  // parent, :86-89
            if (last_rlen < mlen / 3) continue;
            drop.push_back(i);
        } else {  // overlap, :92-126
            int overlap, ms, me;
            if (last.start < seed_start) { overlap = last.end - seed_start + last.mlen; ms = last.start; me = seed_end; }
            else { overlap = seed_end - last.start + mlen; ms = seed_start; me = last.end; }
            if (last.mlen == mlen) {
                add_perfect(ms, me, last.mlen, perfect, contig_len);
                return;
            }
            if (last.mlen < mlen) {
                if (mlen - overlap <= 1 && seed_rlen / mlen < 3) {
                    add_perfect(ms, me, last.mlen, perfect, contig_len);
                    return;
                }
                if (seed_rlen - mlen - overlap <= last.mlen) return;
            } else {
                if (last.mlen - overlap <= 1 && last_rlen / last.mlen < 3) {
                    add_perfect(ms, me, last.mlen, perfect, contig_len);
                    return;
                }
                if (last_rlen - last.mlen - overlap <= mlen) drop.push_back(i);
            }
        }
    }
    for (int i : drop) perfect.erase(perfect.begin() + i);
    if (seed_end > contig_len - mlen) seed_end = contig_len - mlen;  // :137-139
    perfect.push_back(Seed{seed_start, seed_end, mlen, RANK_P});
}

// ---- substitution ---------------------------------------------------------------------------------------------------
namespace {

struct Ref {
    int list;   // RANK_P / RANK_S / RANK_A: which list
    int index;
};

// One scratch vector per recursion depth and thread, so that a call does not allocate (there are ~0.2 calls per base).
struct Scratch {
    std::deque<std::vector<Ref>> near, ps;  // a deque: growing it does not move the vectors the shallower calls use
    int depth = 0;
};
thread_local Scratch g_scratch;
struct ScratchUse {
    std::vector<Ref>& near;
    std::vector<Ref>& ps;
    static std::vector<Ref>& at(std::deque<std::vector<Ref>>& pool, int d) {
        if ((int)pool.size() <= d) pool.resize((size_t)d + 1);
        pool[(size_t)d].clear();
        return pool[(size_t)d];
    }
    ScratchUse() : near(at(g_scratch.near, g_scratch.depth)), ps(at(g_scratch.ps, g_scratch.depth)) { ++g_scratch.depth; }
    ~ScratchUse() { --g_scratch.depth; }
};

// the seeds of the perfect and the substitution list that can touch a candidate starting at seed_start, interleaved by
// descending end (ties: perfect first), retired ones left out; parse_substitute_shiftxor.cpp:47-113
void candidates_ps(const SeedList& perfect, const SeedList& subst, int from_perfect, int seed_start, std::vector<Ref>& out) {
    bool more_p = !perfect.empty(), more_s = !subst.empty();
    int ip = from_perfect, is = (int)subst.size() - 1;
    while (more_p || more_s) {
        if (!more_s) {
            while (more_p) {
                const Seed& p = perfect[(size_t)ip];
                if (p.end >= seed_start) {
                    if (p.rank != RANK_N) out.push_back(Ref{RANK_P, ip});
                    --ip;
                }
                if (ip < 0 || p.end < seed_start) more_p = false;
            }
        } else if (!more_p) {
            while (more_s) {
                const Seed& q = subst[(size_t)is];
                if (q.end >= seed_start) {
                    if (q.rank != RANK_N) out.push_back(Ref{RANK_S, is});
                    --is;
                }
                if (is < 0 || q.end < seed_start) more_s = false;
            }
        } else {
            const Seed& p = perfect[(size_t)ip];
            const Seed& q = subst[(size_t)is];
            const int pe = p.end, qe = q.end;
            if (qe > pe) {
                if (q.rank != RANK_N) out.push_back(Ref{RANK_S, is});
                --is;
            } else {
                if (p.rank != RANK_N) out.push_back(Ref{RANK_P, ip});
                --ip;
            }
            if (ip < 0 || pe < seed_start) more_p = false;
            if (is < 0 || qe < seed_start) more_s = false;
        }
    }
}

}  // namespace

int add_subst(const MergeConfig& cfg, const PlaneCounts& planes, int seed_start, int seed_end, int mlen, SeedList& perfect,
              SeedList& subst, int from_index, int seed_type) {
    from_index = advance_cursor(perfect, from_index, seed_end);
    if (seed_end - seed_start < cfg.cut[mlen - cfg.min_mlen]) return from_index;  // :44

    ScratchUse scratch;
    std::vector<Ref>& near = scratch.near;
    candidates_ps(perfect, subst, from_index, seed_start, near);

    const int seed_length = seed_end - seed_start, seed_rlen = seed_length + mlen;
    auto again = [&](int s, int e, int m, int type) { return add_subst(cfg, planes, s, e, m, perfect, subst, from_index, type); };

    for (const Ref& r : near) {
        Seed& slot = (r.list == RANK_P) ? perfect[(size_t)r.index] : subst[(size_t)r.index];
        const Seed last = slot;
        const int last_length = last.end - last.start, last_rlen = last_length + last.mlen;
        if (last.end < seed_start) break;       // :145
        if (last.rank == RANK_N) continue;      // retired while this call was running
        if (seed_end < last.start) continue;    // :150
        const bool s_vs_pq = seed_type == RANK_S && (last.rank == RANK_P || last.rank == RANK_Q);
        const bool q_vs_p = seed_type == RANK_Q && last.rank == RANK_P;
        const bool q_vs_s = seed_type == RANK_Q && last.rank == RANK_S;
        const bool same = (seed_type == RANK_Q && last.rank == RANK_Q) || (seed_type == RANK_S && last.rank == RANK_S);

        switch (relate(seed_start, seed_end, last.start, last.end)) {
        case IDENTICAL:  // :153-188
            if (s_vs_pq || q_vs_p) return from_index;
            if (q_vs_s) { retire(subst[(size_t)r.index]); break; }
            if (same) {
                if (mlen % last.mlen == 0) return from_index;
                if (last.mlen % mlen == 0) {
                    retire(subst[(size_t)r.index]);
                    return again(seed_start, seed_end, mlen, seed_type);
                }
                // neither divides the other: the plane with more matches over the seed wins, ties go to the smaller motif
                const int mine = planes.match(mlen, seed_start, seed_end), theirs = planes.match(last.mlen, seed_start, seed_end);
                const bool keep_new = mine > theirs || (mine == theirs && mlen < last.mlen);
                if (!keep_new) return from_index;
                retire(subst[(size_t)r.index]);
                goto insert;  // the reference leaves the loop here (:186)
            }
            break;

        case NESTED:  // :191-229
            if (s_vs_pq || q_vs_p) return from_index;
            if (q_vs_s || same) {
                const int merged_type = (seed_type == RANK_S && last.rank == RANK_S) ? RANK_S : RANK_Q;
                if (mlen == last.mlen) {
                    subst[(size_t)r.index] = Seed{last.start, last.end, mlen, merged_type};
                    return from_index;
                }
                if (mlen % last.mlen == 0) return from_index;
                if (last.mlen % mlen == 0 || last.mlen < mlen) {
                    if (seed_rlen >= last.mlen - 1 || seed_rlen >= last_length - 1) {
                        subst[(size_t)r.index] = Seed{last.start, last.end, mlen, merged_type};
                        return from_index;
                    }
                } else if (planes.match(mlen, seed_start, seed_end) < planes.match(last.mlen, seed_start, seed_end)) {
                    return from_index;  // retainNestedSeed: fewer matches than the parent's plane over the seed
                }
            }
            break;

        case PARENT:  // :232-318
            if (s_vs_pq || q_vs_p) {
                if (last.mlen % mlen == 0) {
                    retire(slot);
                    return again(seed_start, seed_end, mlen, RANK_Q);
                }
                if (mlen % last.mlen == 0 || last.mlen < mlen) {
                    const bool many = seed_length / mlen > 3;
                    if ((many && last_rlen >= 3 * mlen - 1) || (!many && (last_rlen >= mlen - 1 || last_rlen >= seed_length - 1))) {
                        if (last.rank != RANK_P) retire(subst[(size_t)r.index]);
                        return again(seed_start, seed_end, last.mlen, RANK_Q);
                    }
                }
                // else both stay
            } else if (q_vs_s) {
                retire(subst[(size_t)r.index]);
                goto insert;  // :276
            } else if (same) {
                if (last.mlen % mlen == 0) {
                    retire(subst[(size_t)r.index]);
                } else if (mlen % last.mlen == 0 || mlen > last.mlen) {
                    if (last_rlen >= mlen - 1 || last_rlen >= seed_length - 1) {
                        retire(subst[(size_t)r.index]);
                        return again(seed_start, seed_end, last.mlen, seed_type);
                    }
                    if (planes.match(last.mlen, last.start, last.end) < planes.match(mlen, last.start, last.end)) retire(subst[(size_t)r.index]);
                } else if (last.mlen > mlen) {
                    if (planes.match(last.mlen, last.start, last.end) < planes.match(mlen, last.start, last.end)) {
                        retire(subst[(size_t)r.index]);
                        return again(seed_start, seed_end, mlen, seed_type);
                    }
                }
            }
            break;

        case OVERLAP: {  // :321-383
            const Overlap o = overlap_of(seed_start, seed_end, mlen, last.start, last.end, last.mlen);
            if (last.mlen % mlen == 0 || last.mlen > mlen) {
                const bool many = last_length / last.mlen > 3;
                if ((many && o.length >= 3 * last.mlen - 1) || (!many && (o.length >= last.mlen - 1 || o.length >= last_length - 1))) {
                    retire(slot);
                    return again(o.start, o.end, mlen, RANK_Q);
                }
            } else if (mlen % last.mlen == 0 || mlen > last.mlen) {
                const bool many = seed_length / mlen > 3;
                if ((many && o.length >= 3 * mlen - 1) || (!many && (o.length >= mlen - 1 || o.length >= seed_length - 1))) {
                    if (last.rank != RANK_P) retire(subst[(size_t)r.index]);
                    return again(o.start, o.end, last.mlen, RANK_Q);
                }
            }
            break;
        }
        }
    }
insert:
    if (seed_end > cfg.contig_len - mlen) seed_end = cfg.contig_len - mlen;  // :382-384
    subst.push_back(Seed{seed_start, seed_end, mlen, seed_type});
    return from_index;
}

// ---- anchored -------------------------------------------------------------------------------------------------------
namespace {

// merge_types.cpp:11-189: the seeds of the three lists that can touch a candidate starting at seed_start. First the
// perfect and the substitution list interleaved by descending end (as above, but the substitution cursor starts at
// from_subst); then that interleaving, walked from its END, against the anchored list from its end.
// An empty substitution list counts as exhausted (the reference reads its first element and crashes).
void candidates_psa(const SeedList& perfect, const SeedList& subst, const SeedList& anchored, int from_perfect, int from_subst,
                    int seed_start, std::vector<Ref>& ps, std::vector<Ref>& out) {
    bool done_p = perfect.empty(), done_s = subst.empty();
    int ip = from_perfect, is = from_subst;
    while (!(done_p && done_s)) {
        if (done_s) {
            for (;;) {
                const Seed& p = perfect[(size_t)ip];
                if (p.end >= seed_start) {
                    if (p.rank != RANK_N) ps.push_back(Ref{RANK_P, ip});
                    --ip;
                }
                if (ip < 0 || p.end < seed_start) { done_p = true; break; }
            }
        } else if (done_p) {
            for (;;) {
                const Seed& q = subst[(size_t)is];
                if (q.end >= seed_start) {
                    if (q.rank != RANK_N) ps.push_back(Ref{RANK_S, is});
                    --is;
                }
                if (is < 0 || q.end < seed_start) { done_s = true; break; }
            }
        } else {
            const Seed& p = perfect[(size_t)ip];
            const Seed& q = subst[(size_t)is];
            const int pe = p.end, qe = q.end;
            if (qe > pe) {
                if (q.rank != RANK_N) ps.push_back(Ref{RANK_S, is});
                --is;
            } else {
                if (p.rank != RANK_N) ps.push_back(Ref{RANK_P, ip});
                --ip;
            }
            if (ip < 0 || pe < seed_start) done_p = true;
            if (is < 0 || qe < seed_start) done_s = true;
        }
    }
    auto end_of = [&](const Ref& r) { return r.list == RANK_P ? perfect[(size_t)r.index].end : subst[(size_t)r.index].end; };
    if (anchored.empty()) {
        out.insert(out.end(), ps.begin(), ps.end());
        return;
    }
    int ia = (int)anchored.size() - 1;
    if (ps.empty()) {
        for (;;) {
            const Seed& a = anchored[(size_t)ia];
            if (a.end >= seed_start) {
                if (a.rank != RANK_N) out.push_back(Ref{RANK_A, ia});
                --ia;
            }
            if (ia < 0 || a.end < seed_start) break;
        }
        return;
    }
    bool done_ps = false, done_a = false;
    int k = (int)ps.size() - 1;
    while (!(done_ps && done_a)) {
        if (done_a) {
            for (;;) {
                const int e = end_of(ps[(size_t)k]);
                if (e >= seed_start) { out.push_back(ps[(size_t)k]); --k; }
                if (k < 0 || e < seed_start) { done_ps = true; break; }
            }
        } else if (done_ps) {
            for (;;) {
                const Seed& a = anchored[(size_t)ia];
                if (a.end >= seed_start) {
                    if (a.rank != RANK_N) out.push_back(Ref{RANK_A, ia});
                    --ia;
                }
                if (ia < 0 || a.end < seed_start) { done_a = true; break; }
            }
        } else {
            const int pe = end_of(ps[(size_t)k]), ae = anchored[(size_t)ia].end;
            if (ae > pe) { out.push_back(Ref{RANK_A, ia}); --ia; }   // retired anchored seeds are not left out here (:167-170)
            else { out.push_back(ps[(size_t)k]); --k; }
            if (k < 0 || pe < seed_start) done_ps = true;
            if (ia < 0 || ae < seed_start) done_a = true;
        }
    }
}

}  // namespace

std::pair<int, int> add_anchored(const MergeConfig& cfg, const PlaneCounts& planes, int seed_start, int seed_end, int mlen,
                                 SeedList& perfect, SeedList& subst, SeedList& anchored, std::pair<int, int> from_indices,
                                 int seed_type) {
    const int from_p = advance_cursor(perfect, from_indices.first, seed_end);
    const int from_s = advance_cursor(subst, from_indices.second, seed_end);
    const std::pair<int, int> cursors(from_p, from_s);
    if (seed_end - seed_start < cfg.cut[mlen - cfg.min_mlen]) return cursors;  // :153

    ScratchUse scratch;
    std::vector<Ref>& near = scratch.near;
    candidates_psa(perfect, subst, anchored, from_p, from_s, seed_start, scratch.ps, near);

    const int seed_length = seed_end - seed_start, seed_rlen = seed_length + mlen;
    // the recursive calls restart from the cursors this call was given, not from the advanced ones (:240, :261, ...)
    auto again = [&](int s, int e, int m, int type) { return add_anchored(cfg, planes, s, e, m, perfect, subst, anchored, from_indices, type); };
    auto retire_higher = [&](const Ref& r, int rank) {  // an existing perfect / substitution seed swallowed by the candidate
        if (rank == RANK_P) retire(perfect[(size_t)r.index]);
        else if (rank == RANK_S || rank == RANK_Q) retire(subst[(size_t)r.index]);
    };

    // The reference keeps these across the loop and reads them again in the coverage rules below, where they hold
    // whatever the last examined seed left in them (:166-167, :449-458): function scope here as well.
    int last_start = 0, last_end = 0, last_rend = 0, last_mlen = 0;
    struct Inside { int index, mlen, rank; };
    std::vector<Inside> inside_factor, inside_nonfactor;  // higher-ranked seeds inside the candidate whose motif divides / does not divide

    for (const Ref& r : near) {
        const Seed last = r.list == RANK_P ? perfect[(size_t)r.index] : (r.list == RANK_S ? subst[(size_t)r.index] : anchored[(size_t)r.index]);
        last_start = last.start; last_end = last.end; last_mlen = last.mlen; last_rend = last.end + last.mlen;
        const int last_type = last.rank;
        if (last_end < seed_start) break;     // :202
        if (last_type == RANK_N) continue;
        if (seed_end < last_start) continue;  // :207
        const int last_length = last_end - last_start, last_rlen = last_rend - last_start;
        const bool same = (seed_type == RANK_A && last_type == RANK_A) || (seed_type == RANK_C && last_type == RANK_C);

        switch (relate(seed_start, seed_end, last_start, last_end)) {
        case IDENTICAL:  // :214-226
            if (seed_type == RANK_A && last_type > RANK_A) return cursors;
            if (seed_type == RANK_C && last_type == RANK_A) retire(anchored[(size_t)r.index]);
            break;

        case NESTED:  // :229-262
            if (last_type > seed_type) return cursors;
            if (same) {
                if (mlen % last_mlen == 0 && mlen != 4) return cursors;
                if (last_mlen % mlen == 0 && last_mlen != 4) {
                    if (seed_rlen >= last_mlen - 1 || seed_rlen >= last_length) {
                        retire(anchored[(size_t)r.index]);
                        return again(last_start, last_end, mlen, seed_type);
                    }
                } else if (planes.anchored(mlen, seed_start, seed_end) < planes.anchored(last_mlen, seed_start, seed_end)) {
                    return cursors;  // retainNestedSeedAnchored
                }
            }
            break;

        case PARENT:  // :265-347
            if (last_type > seed_type) {
                if (mlen % last_mlen == 0) {
                    if (last_rlen >= mlen - 2 || last_rlen >= seed_length - 2) {
                        retire_higher(r, last_type);
                        return again(seed_start, seed_end, last_mlen, RANK_C);
                    }
                    inside_factor.push_back(Inside{r.index, last_mlen, last_type});
                } else if (last_mlen % mlen == 0 || last_mlen > mlen) {
                    if (last_mlen >= 4 * mlen || last_length >= 4 * mlen) {
                        retire_higher(r, last_type);
                        return again(seed_start, seed_end, mlen, RANK_C);
                    }
                } else {
                    inside_nonfactor.push_back(Inside{r.index, last_mlen, last_type});
                }
            } else if (seed_type == RANK_C && last_type == RANK_A) {
                retire(anchored[(size_t)r.index]);
            } else if (same) {
                if (last_mlen == mlen) {
                    retire(anchored[(size_t)r.index]);
                } else if (planes.anchored(last_mlen, last_start, last_end) < planes.anchored(mlen, last_start, last_end)) {
                    retire(anchored[(size_t)r.index]);
                } else if (mlen % last_mlen == 0) {
                    if (last_rlen >= mlen - 2 || last_rlen >= seed_length - 2) {
                        retire(anchored[(size_t)r.index]);
                        return again(seed_start, seed_end, last_mlen, seed_type);
                    }
                }
            }
            break;

        case OVERLAP: {  // :350-432
            const Overlap o = overlap_of(seed_start, seed_end, mlen, last_start, last_end, last_mlen);
            if (seed_type == RANK_A && last_type > RANK_C) {
                if (mlen == last_mlen && o.length >= 4 * mlen) {
                    retire_higher(r, last_type);
                    return again(o.start, o.end, mlen, RANK_C);
                }
                if (!(mlen % last_mlen == 0 || last_mlen % mlen == 0) && (o.length >= mlen - 1 || o.length >= seed_length - 1)) return cursors;
            } else if ((seed_type == RANK_A || seed_type == RANK_C) && (last_type == RANK_A || last_type == RANK_C)) {
                if (mlen == last_mlen) {
                    bool merge;  // both alternatives of each pair lead to the same merge
                    if (last_length >= seed_length) {
                        merge = (seed_length >= 3 * mlen && (o.length >= 3 * mlen - 1 || o.length >= seed_length - 1)) ||
                                (seed_length < 3 * mlen && (o.length >= mlen - 1 || o.length >= seed_length - 1));
                    } else {
                        merge = (last_length >= 3 * last_mlen && (o.length >= 3 * last_mlen - 1 || o.length >= last_length - 1)) ||
                                (seed_length < 3 * last_mlen && (o.length >= last_mlen - 1 || o.length >= last_length - 1));
                    }
                    if (merge) {
                        retire(anchored[(size_t)r.index]);
                        return again(o.start, o.end, last_mlen, seed_type);  // the type stays the candidate's (:392 is a comparison, not an assignment)
                    }
                }
            }
            break;
        }
        }
    }

    // ---- coverage rules, :434-526. They read the lists at the POSITION in the bookkeeping vectors (j), not at the stored
    // index, and fall back on the values the loop above left behind for types they do not handle — as the reference does.
    {
        uint32_t prev_start = (uint32_t)-1;
        if (!inside_nonfactor.empty()) {
            int coverage = 0;
            for (size_t j = 0; j < inside_nonfactor.size(); ++j) {
                const int t = inside_nonfactor[j].rank;
                const SeedList* src = t == RANK_P ? &perfect : (t == RANK_S ? &subst : nullptr);
                if (src && j < src->size()) {
                    const Seed& x = (*src)[j];
                    last_start = x.start; last_mlen = x.mlen; last_end = x.end; last_rend = x.end + x.mlen;
                }
                if ((uint32_t)last_rend >= prev_start) coverage += (int)(prev_start - (uint32_t)last_start);
                else if (last_rend < seed_end) coverage += last_rend - last_start;
                else coverage += seed_end - last_start;
                prev_start = (uint32_t)last_start;
            }
            if (coverage > 0.5 * seed_length) return cursors;
        }
        if (!inside_factor.empty()) {
            std::map<int, int> prev_starts, coverages;
            for (const Inside& x : inside_factor) { prev_starts[x.mlen] = -1; coverages[x.mlen] = 0; }
            for (size_t j = 0; j < inside_factor.size(); ++j) {
                const int t = inside_factor[j].rank;
                const SeedList* src = t == RANK_P ? &perfect : (t == RANK_S ? &subst : nullptr);
                if (src && j < src->size()) {
                    const Seed& x = (*src)[j];
                    last_start = x.start; last_mlen = x.mlen; last_end = x.end; last_rend = x.end + x.mlen;
                }
                prev_start = (uint32_t)prev_starts[last_mlen];
                if ((uint32_t)last_rend >= prev_start) coverages[last_mlen] += (int)(prev_start - (uint32_t)last_start);
                else if (last_rend < seed_end) coverages[last_mlen] += last_rend - last_start;
                else coverages[last_mlen] += seed_end - last_start;
                prev_starts[last_mlen] = last_start;
            }
            for (const auto& kv : coverages) {  // ascending motif size
                const int factor = kv.first;
                if (kv.second >= 0.8 * seed_length) {
                    mlen = factor; seed_type = RANK_C;
                    for (size_t j = 0; j < inside_factor.size(); ++j) {
                        const int t = inside_factor[j].rank;
                        SeedList* src = t == RANK_P ? &perfect : (t == RANK_S ? &subst : nullptr);
                        if (src && j < src->size()) {
                            last_mlen = (*src)[j].mlen;
                            if (last_mlen == factor) (*src)[j] = Seed{last_start, last_end, last_mlen, RANK_N};
                        }
                    }
                    break;
                }
            }
        }
    }
    if (seed_end > cfg.contig_len - mlen) seed_end = cfg.contig_len - mlen;  // :529-531
    anchored.push_back(Seed{seed_start, seed_end, mlen, seed_type});
    return cursors;
}

// ---- the three passes over the candidate streams ------------------------------------------------------------------------
// What processSequence does with the scan's candidates (fasta_utils.cpp:132-170 with the loops of
// parse_*_shiftxor.cpp inlined): every record is one call; a PSEUDO record stands for the calls the scan elided (only
// their cursor effect matters, and only the largest end counts: a call with an empty interval at that end moves the
// cursors the same way and is dropped by the cutoff); a NOCOMMIT record is a tail-flush call whose returned cursors the
// reference discards (parse_anchored_shiftxor.cpp:688-719).
void run_merges(const Cand* const cand[3], const int64_t n[3], const PlaneCounts& planes, int min_mlen, int max_mlen, int contig_len,
                SeedList& perfect, SeedList& subst, SeedList& anchored) {
    const int nm = max_mlen - min_mlen + 1;
    std::vector<int> cutS((size_t)nm), cutA((size_t)nm);
    for (int k = 0; k < nm; ++k) { cutS[(size_t)k] = rb::cut_subst(min_mlen + k); cutA[(size_t)k] = rb::cut_anch(min_mlen + k); }
    for (int64_t k = 0; k < n[0]; ++k) add_perfect(cand[0][k].start, cand[0][k].end, cand[0][k].mlen, perfect, contig_len);
    {
        const MergeConfig cfg{min_mlen, contig_len, cutS.data()};
        int from = 0;
        for (int64_t k = 0; k < n[1]; ++k) {
            const Cand& c = cand[1][k];
            if (c.flags & CAND_PSEUDO) {
                if (c.end >= 0) from = add_subst(cfg, planes, c.end, c.end, min_mlen, perfect, subst, from, RANK_S);
                continue;
            }
            from = add_subst(cfg, planes, c.start, c.end, c.mlen, perfect, subst, from, RANK_S);
        }
    }
    {
        const MergeConfig cfg{min_mlen, contig_len, cutA.data()};
        std::pair<int, int> from(0, 0);
        for (int64_t k = 0; k < n[2]; ++k) {
            const Cand& c = cand[2][k];
            if (c.flags & CAND_PSEUDO) {
                if (c.end >= 0) from = add_anchored(cfg, planes, c.end, c.end, min_mlen, perfect, subst, anchored, from, RANK_A);
                continue;
            }
            const std::pair<int, int> ret = add_anchored(cfg, planes, c.start, c.end, c.mlen, perfect, subst, anchored, from, RANK_A);
            if (!(c.flags & CAND_NOCOMMIT)) from = ret;
        }
    }
}

// ---- plane counts from the packed planes -----------------------------------------------------------------------------
struct PackedPlaneCounts::Impl {
    std::vector<rb::PlaneWord> words;  // one guard word in front, guard words behind (as the device layout)
    const rb::PlaneWord* cw;
    int L, nw, s_lo, s_hi;
    uint32_t x(int w, int s) const { return rb::x_word(cw, w, s); }
    // anchor word A_s[w] (parse_anchored_shiftxor.cpp:20-56) without carried state: the run length in front of the word is
    // found by looking back
    uint32_t anchor(int w, int s) const {
        const int K2 = 2 * s;
        int lenL = 0;
        for (int k = w - 1; k >= 0; --k) {
            const uint32_t xa = x(k, s) | rb::anchor_endmask(k, L, s);
            const int tr = rb::clz32(~xa);
            lenL += tr;
            if (tr < 32 || lenL >= K2 + 32) break;
        }
        const uint32_t xa = x(w, s) | rb::anchor_endmask(w, L, s);
        const uint32_t xan = x(w + 1, s) | rb::anchor_endmask(w + 1, L, s);
        return rb::anchor_word(cw, w, L, s, xa, xan, lenL);
    }
    template <class F>
    int count(int a, int b, F word) const {
        if (a < 0) a = 0;
        if (b > L) b = L;
        int n = 0;
        for (int w = a >> 5; 32 * w < b; ++w) {
            uint32_t v = word(w);
            if (w == (a >> 5)) v &= ~rb::lowmask(a & 31);
            const int rem = b - 32 * w;
            if (rem < 32) v &= rb::lowmask(rem);
            n += rb::popc32(v);
        }
        return n;
    }
};

PackedPlaneCounts::PackedPlaneCounts(const uint32_t* hi, const uint32_t* lo, const uint32_t* nn, int contig_len, int min_shift, int max_shift)
    : p_(new Impl) {
    p_->L = contig_len; p_->nw = (contig_len + 31) / 32; p_->s_lo = min_shift; p_->s_hi = max_shift;
    const int guard = (max_shift >> 5) + 5;
    p_->words.assign((size_t)(1 + p_->nw + guard), rb::PlaneWord{0u, 0u, 0xFFFFFFFFu, 0u});
    for (int w = 0; w < p_->nw; ++w) {
        rb::PlaneWord o{hi[w], lo[w], nn[w], 0u};
        const int rem = contig_len - 32 * w;
        if (rem < 32) o.n |= ~rb::lowmask(rem);
        p_->words[(size_t)(1 + w)] = o;
    }
    p_->cw = p_->words.data() + 1;
}
PackedPlaneCounts::~PackedPlaneCounts() { delete p_; }

int PackedPlaneCounts::match(int mlen, int a, int b) const {
    return p_->count(a, b, [&](int w) { return p_->x(w, mlen); });
}
int PackedPlaneCounts::anchored(int mlen, int a, int b) const {
    return p_->count(a, b, [&](int w) {
        uint32_t v = p_->x(w, mlen);
        for (int i = (mlen > 2) ? mlen - 2 : 1; i <= mlen + 2; ++i)
            if (i != mlen && i >= p_->s_lo && i <= p_->s_hi) v |= p_->anchor(w, i);
        return v;
    });
}

// ---- plane counts with the anchor planes given (the binding has them anyway: the per-seed stage reads the anchored planes)
struct AnchorArrayCounts::Impl {
    PackedPlaneCounts::Impl* base;
    const uint32_t* anchors;
    long nw;
};
AnchorArrayCounts::AnchorArrayCounts(const uint32_t* hi, const uint32_t* lo, const uint32_t* nn, int contig_len, int min_shift, int max_shift,
                                     const uint32_t* anchors)
    : PackedPlaneCounts(hi, lo, nn, contig_len, min_shift, max_shift), q_(new Impl) {
    q_->base = impl(); q_->anchors = anchors; q_->nw = ((long)contig_len + 31) / 32;
}
AnchorArrayCounts::~AnchorArrayCounts() { delete q_; }
int AnchorArrayCounts::anchored(int mlen, int a, int b) const {
    const PackedPlaneCounts::Impl* p = q_->base;
    return p->count(a, b, [&](int w) {
        uint32_t v = p->x(w, mlen);
        for (int i = (mlen > 2) ? mlen - 2 : 1; i <= mlen + 2; ++i)
            if (i != mlen && i >= p->s_lo && i <= p->s_hi) v |= q_->anchors[(size_t)(i - p->s_lo) * (size_t)q_->nw + (size_t)w];
        return v;
    });
}

}  // namespace rbm

// C entry for the tests (ctypes): candidates as rows {start, end, mlen, flags}, planes as rb_get_planes returns them;
// out[k] = malloc'ed rows {start, end, mlen, rank} of list k (perfect, substitution, anchored), n[k] rows.
extern "C" int rbm_run(const int32_t* cand_p, int64_t n_p, const int32_t* cand_s, int64_t n_s, const int32_t* cand_a, int64_t n_a,
                       const uint32_t* hi, const uint32_t* lo, const uint32_t* nn, int32_t contig_len, int32_t min_mlen, int32_t max_mlen,
                       int32_t** out, int64_t* n) {
    static_assert(sizeof(rbm::Cand) == 16, "Cand is four int32");
    const int s_lo = min_mlen > 2 ? min_mlen - 2 : 1, s_hi = max_mlen + 2;  // ribbit.cpp:241-242
    rbm::PackedPlaneCounts planes(hi, lo, nn, contig_len, s_lo, s_hi);
    const rbm::Cand* cand[3] = {(const rbm::Cand*)cand_p, (const rbm::Cand*)cand_s, (const rbm::Cand*)cand_a};
    const int64_t cn[3] = {n_p, n_s, n_a};
    rbm::SeedList lists[3];
    rbm::run_merges(cand, cn, planes, min_mlen, max_mlen, contig_len, lists[0], lists[1], lists[2]);
    for (int k = 0; k < 3; ++k) {
        n[k] = (int64_t)lists[k].size();
        out[k] = (int32_t*)malloc(std::max<size_t>(1, lists[k].size()) * 16);
        if (!out[k]) return -1;
        for (size_t i = 0; i < lists[k].size(); ++i) {
            out[k][4 * i] = lists[k][i].start; out[k][4 * i + 1] = lists[k][i].end; out[k][4 * i + 2] = lists[k][i].mlen; out[k][4 * i + 3] = lists[k][i].rank;
        }
    }
    return 0;
}
extern "C" void rbm_free(void* p) { free(p); }
