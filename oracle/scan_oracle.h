/* TEST INFRASTRUCTURE — not part of the product path.
 *
 * Plain-C restatement of ribbit's seed-scanning hot path (the CP1 checkpoint
 * of SURVEY.md §8c): the exact top-level argument sequence (start, end, mlen)
 * that the reference passes to addSeedToSeedPositions{Perfect,Substitutions,
 * Anchored}, in call order. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: PINNED against the reference itself — the reference has no
 * tests or golden vectors, so tests/test_oracle.py compares
 * this restatement with oracle/_ref/ribbit_ref_cp (the unmodified reference
 * sources with CP1 logging) on committed fixtures (tests/golden/) and, when
 * the _ref binary is present, on fresh fuzz inputs.
 */
#ifndef RB_SCAN_ORACLE_H
#define RB_SCAN_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t stream; /* 1 perfect, 2 substitution, 3 anchored */
    int32_t start;
    int32_t end;
    int32_t mlen;
    int32_t time; /* position being processed when the reference makes the call; -1 = tail flush after the loop */
} rbo_event;

/* Scans one contig (ASCII, length L) for motif sizes [min_mlen, max_mlen].
 * Returns the number of events; *out is malloc'ed (perfect events first, then
 * substitution, then anchored, each in reference call order). Free with
 * rbo_free. Returns -1 on allocation failure. */
int64_t rbo_scan(const char *seq, int64_t L, int min_mlen, int max_mlen, rbo_event **out);

/* Same, but only counts events per stream (counts[1..3]); used for timing. */
int64_t rbo_scan_count(const char *seq, int64_t L, int min_mlen, int max_mlen, int64_t counts[4]);

/* ASCII -> three bit planes, 32 bases per word, bit i of word w = position
 * 32*w+i (fasta_utils.cpp:90-115 semantics: A=00 C=01 G=10 T=11, anything
 * else N with code 00). nwords = ceil(L/32); bits past L are zero. */
void rbo_pack(const char *seq, int64_t L, uint32_t *hi, uint32_t *lo, uint32_t *nn);

/* Anchored plane B_m (fasta_utils.cpp:143-161) as bytes 0/1 for positions
 * [p0, p1) of one contig; used to check the seed filter. */
void rbo_anchored_plane(const char *seq, int64_t L, int min_mlen, int max_mlen, int m, int64_t p0, int64_t p1,
                        uint8_t *out);

void rbo_free(void *p);

/* oracle/motif_oracle.c: row search of mostFrequentLongerMotif (parse_seed.cpp:153-256); returns mmotif_index. */
int32_t rbo_motif_row(const char *seq, int64_t L, int32_t seed_start, int32_t seed_len, int32_t m, int32_t *best);

#ifdef __cplusplus
}
#endif
#endif
