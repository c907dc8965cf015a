// ribbit-b200 host side: the three order-dependent seed-list merges of ribbit, restated.
//
//   add_perfect    addSeedToSeedPositionsPerfect         parse_perfect_shiftxor.cpp:47-142
//   add_subst      addSeedToSeedPositionsSubstitutions   parse_substitute_shiftxor.cpp:18-388
//   add_anchored   addSeedToSeedPositionsAnchored        parse_anchored_shiftxor.cpp:113-534  (+ mergeAllLists, merge_types.cpp:11-189)
//
// They consume the candidate streams of the scan library one call at a time, in the reference's call order, and build
// the same three seed lists (checkpoint CP2), entry for entry, rank for rank. Differences in form: no globals (the
// configuration is a struct), seeds are structs, and the tie-breakers that count ones of a match plane over a seed
// (retainNestedSeed / retainIdenticalSeeds, parse_perfect_shiftxor.cpp:18-43, parse_anchored_shiftxor.cpp:59-84) ask a
// PlaneCounts object, which computes them from the packed code planes on demand — the reference indexes 102 full-length
// bitsets for this. One intentional divergence: where the reference dereferences an empty substitution list
// (merge_types.cpp:47-50, SURVEY.md F6: it crashes) the list counts as exhausted.
#ifndef RB_SEED_MERGE_H
#define RB_SEED_MERGE_H

#include <cstdint>
#include <utility>
#include <vector>

namespace rbm {

// global_variables.cpp:28-34
enum : int { RANK_P = 5, RANK_Q = 4, RANK_S = 3, RANK_F = 2, RANK_C = 1, RANK_A = 0, RANK_N = -1 };

struct Seed {
    int start, end, mlen, rank;
};
typedef std::vector<Seed> SeedList;

struct MergeConfig {
    int min_mlen;     // MINIMUM_MLEN
    int contig_len;   // bset_size
    const int* cut;   // consumer cutoff per motif size, index mlen - min_mlen (parse_substitute_shiftxor.cpp:423 / parse_anchored_shiftxor.cpp:572-573)
};

// ones of a plane over positions [a, b)
struct PlaneCounts {
    virtual ~PlaneCounts() {}
    virtual int match(int mlen, int a, int b) const = 0;     // X_mlen: what the planes hold during the substitution pass
    virtual int anchored(int mlen, int a, int b) const = 0;  // B_mlen = X_mlen | anchors of the neighbouring shifts (fasta_utils.cpp:143-161)
};

void add_perfect(int seed_start, int seed_end, int mlen, SeedList& perfect, int contig_len);
int add_subst(const MergeConfig& cfg, const PlaneCounts& planes, int seed_start, int seed_end, int mlen, SeedList& perfect,
              SeedList& subst, int from_index, int seed_type);
std::pair<int, int> add_anchored(const MergeConfig& cfg, const PlaneCounts& planes, int seed_start, int seed_end, int mlen,
                                 SeedList& perfect, SeedList& subst, SeedList& anchored, std::pair<int, int> from_indices,
                                 int seed_type);

// one candidate of a stream as the scan library reports it (rb_rec without the time)
struct Cand {
    int32_t start, end, mlen, flags;
};
enum : int { CAND_DROPPED = 1, CAND_PSEUDO = 2, CAND_NOCOMMIT = 4 };  // RB_REC_* of include/ribbit_scan.h

// the three passes of processSequence over the candidate streams (perfect, substitution, anchored), in the reference's order
void run_merges(const Cand* const cand[3], const int64_t n[3], const PlaneCounts& planes, int min_mlen, int max_mlen, int contig_len,
                SeedList& perfect, SeedList& subst, SeedList& anchored);

// PlaneCounts over the packed planes of one contig (32 bases per word, bit i of word w = position 32 w + i), as
// rb_get_planes returns them. The match words and the anchor words are computed where they are asked for, with the
// very functions the kernels use (csrc/scan_core.h compiled for the host).
class PackedPlaneCounts : public PlaneCounts {
public:
    PackedPlaneCounts(const uint32_t* hi, const uint32_t* lo, const uint32_t* nn, int contig_len, int min_shift, int max_shift);
    ~PackedPlaneCounts();
    int match(int mlen, int a, int b) const override;
    int anchored(int mlen, int a, int b) const override;
    struct Impl;

protected:
    Impl* impl() const { return p_; }

private:
    Impl* p_;
};

// The same with the anchor planes A_s given as rb_get_anchor_planes returns them (anchors[(s - min_shift) * ceil(L/32) + w]):
// B_m words are then plain ORs. The drop-in binding fetches these planes anyway for the per-seed stage.
class AnchorArrayCounts : public PackedPlaneCounts {
public:
    AnchorArrayCounts(const uint32_t* hi, const uint32_t* lo, const uint32_t* nn, int contig_len, int min_shift, int max_shift,
                      const uint32_t* anchors);
    ~AnchorArrayCounts();
    int anchored(int mlen, int a, int b) const override;

private:
    struct Impl;
    Impl* q_;
};

}  // namespace rbm
#endif
