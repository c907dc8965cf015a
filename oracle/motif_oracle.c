/* TEST INFRASTRUCTURE — not part of the product path (see oracle/Makefile).
 *
 * Plain-C, base-at-a-time restatement of the row search of the reference's consensus-motif step for motif sizes > 10,
 * mostFrequentLongerMotif (parse_seed.cpp:153-256): for every row (candidate motif start) of the seed, walk the repeat
 * downstream and upstream unit by unit, at every unit take the best of the five diagonals -2..+2 of the dot matrix
 * (MATRIX[p][q] = "base p is A/C/G/T and equals base q", fasta_utils.cpp:88-114), add a partial prefix unit, and return the
 * first row with the largest total. Pinned against the instrumented reference (checkpoint CP4 of oracle/instrument.sh)
 * by tests/test_oracle.py via tests/golden/golden_motif.npz.
 */
#include <stdint.h>
#include <stdlib.h>

static int code_of(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

static int dot(const char *seq, int64_t p, int64_t q) { /* (*MATRIX[p])[L-1-q] */
    const int c = code_of(seq[p]);
    return c >= 0 && c == code_of(seq[q]);
}

/* parse_seed.cpp:171-243. seed_len is the reference's seed_sequence_length (N-truncated, motif length included).
 * Returns mmotif_index (0 when no row scores, as the reference's initial value), *best = max_count. */
int32_t rbo_motif_row(const char *seq, int64_t L, int32_t seed_start, int32_t seed_len, int32_t m, int32_t *best) {
    const int64_t seed_end = (int64_t)seed_start + seed_len;
    int32_t best_row = 0, best_count = 0;
    (void)L;
    for (int64_t row = seed_start; row < seed_end - m + 1; ++row) {
        int32_t total = 0;
        int64_t d = row + m; /* parse_seed.cpp:183-199 */
        while (d < seed_end) {
            int bx = -2, bc = 0;
            for (int x = -2; x < 3; ++x) {
                int c = 0;
                for (int i = 0; i < m; ++i) {
                    if (d + x + i >= seed_end) break;
                    c += dot(seq, row + i, d + x + i);
                }
                if (c > bc) { bc = c; bx = x; }
            }
            total += bc;
            d += bx + m;
        }
        int64_t u = row - m; /* parse_seed.cpp:201-218 */
        while (u > seed_start) {
            int bx = -2, bc = 0;
            for (int x = -2; x < 3; ++x) {
                int c = 0;
                for (int i = 0; i < m; ++i) {
                    if (u + x + i < 0) break;
                    c += dot(seq, row + i, u + x + i);
                }
                if (c > bc) { bc = c; bx = x; }
            }
            total += bc;
            u += bx - m;
        }
        if (u < seed_start && llabs(u - seed_start) < m) { /* parse_seed.cpp:220-238: the partial unit at the seed start */
            const int64_t last_row = row + m - 1;
            const int64_t pc = seed_start + ((m + (u - seed_start)) - 1);
            const int64_t rows = m + (u - seed_start);
            int bc = 0;
            for (int x = -2; x < 3; ++x) {
                int c = 0;
                for (int64_t i = 0; i < rows; ++i) {
                    if (pc + x - i >= seed_end || pc + x - i < seed_start) break;
                    c += dot(seq, last_row - i, pc + x - i);
                }
                if (c > bc) bc = c;
            }
            total += bc;
        }
        if (total > best_count) { best_count = total; best_row = (int32_t)row; }
    }
    if (best) *best = best_count;
    return best_row;
}
