/* TEST INFRASTRUCTURE — not part of the product path. See scan_oracle.h.
 *
 * Bit-at-a-time, position-major restatement of the reference scan. Every
 * function cites the reference lines it follows (paths under /root/reference).
 * Nothing here is optimised: it is the checker, and (as the "port" kind of
 * cpu_baseline in bench.py) a single-threaded stand-in for the reference's
 * scan stage when oracle/_ref/ is unavailable.
 */
#include "scan_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
    rbo_event *v;
    int64_t n, cap;
    int count_only;
    int failed;
    int64_t counts[4];
} sink_t;

static void emit(sink_t *s, int stream, int start, int end, int mlen, int time) {
    s->counts[stream]++;
    if (s->count_only) { s->n++; return; }
    if (s->n == s->cap) {
        int64_t nc = s->cap ? s->cap * 2 : 4096;
        rbo_event *nv = (rbo_event *)realloc(s->v, (size_t)nc * sizeof(rbo_event));
        if (!nv) { s->failed = 1; return; }
        s->v = nv; s->cap = nc;
    }
    s->v[s->n].stream = stream; s->v[s->n].start = start; s->v[s->n].end = end; s->v[s->n].mlen = mlen;
    s->v[s->n].time = time;
    s->n++;
}

/* fasta_utils.cpp:90-115: A=00, C=01, G=10, T=11 (either case); anything else sets the N plane, code 00 */
static void encode(const char *seq, int64_t L, uint8_t *code, uint8_t *nn) {
    for (int64_t i = 0; i < L; i++) {
        switch (seq[i]) {
            case 'A': case 'a': code[i] = 0; nn[i] = 0; break;
            case 'C': case 'c': code[i] = 1; nn[i] = 0; break;
            case 'G': case 'g': code[i] = 2; nn[i] = 0; break;
            case 'T': case 't': code[i] = 3; nn[i] = 0; break;
            default: code[i] = 0; nn[i] = 1; break;
        }
    }
}

/* fasta_utils.cpp:121: ~(left ^ (left<<s)) & ~(right ^ (right<<s)); bit index L-1-p, so the left shift brings in
 * position p+s; zeros are shifted in past the end, hence code[p]==0 "matches" there. N is not consulted. */
static inline int X(const uint8_t *code, int64_t L, int s, int64_t p) {
    return (p + s < L) ? (code[p] == code[p + s]) : (code[p] == 0);
}

/* parse_perfect_shiftxor.cpp:146-226 */
static void scan_perfect(const uint8_t *code, const uint8_t *nn, int64_t L, int m_lo, int m_hi, int s_lo, sink_t *out) {
    int nm = m_hi - m_lo + 1;
    int *last = (int *)malloc(sizeof(int) * (size_t)nm);
    /* :161 `int last_starts[NMOTIFS] = {-1};` -> first element -1, the rest 0 */
    for (int d = 0; d < nm; d++) last[d] = d == 0 ? -1 : 0;
    for (int64_t p = 0; p < L; p++) {
        if (nn[p]) { /* :175-186 */
            for (int d = 0; d < nm; d++) {
                int m = m_lo + d, midx = m - s_lo;
                int cutoff = (m <= 6) ? 12 - m : m + midx; /* :179 */
                if (last[d] != -1) {
                    if ((int)p - last[d] >= cutoff) emit(out, 1, last[d], (int)p, m, (int)p);
                    last[d] = -1;
                }
            }
        } else { /* :190-208 */
            for (int d = 0; d < nm; d++) {
                int m = m_lo + d;
                int cutoff = (m <= 6) ? 12 - m : m; /* :193 */
                if (X(code, L, m, p)) {
                    if (last[d] == -1) last[d] = (int)p;
                } else {
                    if (last[d] != -1 && (int)p - last[d] >= cutoff) emit(out, 1, last[d], (int)p, m, (int)p);
                    last[d] = -1;
                }
            }
        }
    }
    int wp = (int)L - 1; /* :213 */
    for (int d = 0; d < nm; d++) { /* :214-223 */
        int m = m_lo + d;
        int cutoff = (m <= 6) ? 12 - m : m;
        if (last[d] != -1) {
            if (wp - last[d] >= cutoff) emit(out, 1, last[d], wp, m, -1);
            last[d] = -1;
        }
    }
    free(last);
}

/* parse_anchored_shiftxor.cpp:20-56; one bitset (64 positions per word) per shift */
static void anchors_for_shift(const uint8_t *code, int64_t L, int s, uint64_t *A) {
    int64_t run = -1;
    for (int64_t p = 0; p + s <= L - 1; p++) { /* :37 xor_idx >= s  <=>  p <= L-1-s */
        if (X(code, L, s, p)) {
            if (run == -1) run = p;
        } else {
            if (run != -1) {
                int64_t len = p - run;
                if (len >= 3 && len < 2 * (int64_t)s) /* :44 anchor_size = 3 */
                    for (int64_t q = run; q < p; q++) A[q >> 6] |= 1ull << (q & 63);
            }
            run = -1;
        }
    }
}

typedef struct {
    const uint8_t *code;
    int64_t L;
    int s_lo;
    int64_t wpl;        /* words per anchor plane */
    const uint64_t *A;  /* NULL -> substitution pass (Y = X) */
} yctx_t;

/* fasta_utils.cpp:146-160: B_m = X_m | A_{lo..m+2, != m}, lo = m>2 ? m-2 : 1 */
static inline int Y(const yctx_t *c, int m, int64_t p) {
    if (X(c->code, c->L, m, p)) return 1;
    if (!c->A) return 0;
    int lo = (m > 2) ? m - 2 : 1;
    for (int i = lo; i <= m + 2; i++) {
        if (i == m) continue;
        const uint64_t *a = c->A + (int64_t)(i - c->s_lo) * c->wpl;
        if ((a[p >> 6] >> (p & 63)) & 1) return 1;
    }
    return 0;
}

/* parse_substitute_shiftxor.cpp:391-577 (stream 2, T=7, Y=X) and
 * parse_anchored_shiftxor.cpp:538-726 (stream 3, T=6, Y=B): the same window machine */
static void scan_windows(const yctx_t *c, const uint8_t *nn, int m_lo, int m_hi, int T, int stream, sink_t *out) {
    int nm = m_hi - m_lo + 1;
    int64_t L = c->L;
    int *ls = (int *)malloc(sizeof(int) * (size_t)nm * 3), *le = ls + nm, *cur = le + nm;
    uint8_t *win = (uint8_t *)calloc((size_t)nm, 1);
    for (int d = 0; d < nm; d++) ls[d] = le[d] = cur[d] = -1;
    int valid = 0;
    int wp = -8; /* :429 window_position = -1*window_length */
    for (int64_t p = 0; p < L; p++) {
        wp += 1;
        if (nn[p]) { /* :433-458 */
            for (int d = 0; d < nm; d++) {
                if (cur[d] != -1) {
                    cur[d] = wp;
                    if (le[d] != -1 && le[d] < cur[d]) {
                        emit(out, stream, ls[d], le[d], m_lo + d, (int)p);
                        ls[d] = -1; le[d] = -1;
                    }
                }
                win[d] = 0; /* <<= window_length */
                cur[d] = -1;
            }
            valid = 0;
        } else { /* :460-531 */
            valid += 1;
            for (int d = 0; d < nm; d++) win[d] = (uint8_t)((win[d] << 1) | Y(c, m_lo + d, p));
            if (valid >= 8) {
                for (int d = 0; d < nm; d++) {
                    int m = m_lo + d;
                    if (__builtin_popcount(win[d]) >= T) {
                        if (cur[d] == -1) {
                            cur[d] = wp;
                            if (le[d] != -1 && le[d] < cur[d]) {
                                emit(out, stream, ls[d], le[d], m, (int)p);
                                ls[d] = -1; le[d] = -1;
                            }
                        }
                    } else {
                        if (cur[d] != -1) {
                            if (ls[d] == -1) { ls[d] = cur[d]; le[d] = wp + 8 - 1; }
                            else le[d] = wp + 8 - 1;
                            cur[d] = -1;
                        } else if (le[d] != -1 && le[d] < wp) {
                            emit(out, stream, ls[d], le[d], m, (int)p);
                            ls[d] = -1; le[d] = -1;
                        }
                    }
                }
            }
        }
    }
    /* tail flush :534-574 / :681-723; (bset_size - (xor_idx + 1)) == L after the loop */
    for (int d = 0; d < nm; d++) {
        int m = m_lo + d;
        if (le[d] == -1) {
            if (cur[d] != -1) emit(out, stream, cur[d], (int)L, m, -1);
        } else if (cur[d] == -1) {
            emit(out, stream, ls[d], le[d], m, -1);
        } else if (le[d] >= cur[d] - m) {
            emit(out, stream, ls[d], (int)L, m, -1);
        } else {
            emit(out, stream, ls[d], le[d], m, -1);
            emit(out, stream, cur[d], (int)L, m, -1);
        }
    }
    free(ls);
    free(win);
}

static int64_t scan_impl(const char *seq, int64_t L, int m_lo, int m_hi, sink_t *out) {
    /* ribbit.cpp:240-243 */
    int s_lo = (m_lo > 2) ? m_lo - 2 : 1, s_hi = m_hi + 2, ns = s_hi - s_lo + 1;
    uint8_t *code = (uint8_t *)malloc((size_t)(L > 0 ? L : 1) * 2), *nn = code + (L > 0 ? L : 1);
    if (!code) return -1;
    encode(seq, L, code, nn);

    scan_perfect(code, nn, L, m_lo, m_hi, s_lo, out);

    yctx_t c = {code, L, s_lo, (L + 63) / 64 + 1, NULL};
    scan_windows(&c, nn, m_lo, m_hi, 7, 2, out);

    uint64_t *A = (uint64_t *)calloc((size_t)(c.wpl * ns), sizeof(uint64_t));
    if (!A) { free(code); return -1; }
    for (int s = s_lo; s <= s_hi; s++) anchors_for_shift(code, L, s, A + (int64_t)(s - s_lo) * c.wpl);
    c.A = A;
    scan_windows(&c, nn, m_lo, m_hi, 6, 3, out); /* fasta_utils.cpp:165 threshold 6 */

    free(A);
    free(code);
    return out->failed ? -1 : out->n;
}

int64_t rbo_scan(const char *seq, int64_t L, int min_mlen, int max_mlen, rbo_event **outp) {
    sink_t s;
    memset(&s, 0, sizeof s);
    int64_t n = scan_impl(seq, L, min_mlen, max_mlen, &s);
    if (n < 0) { free(s.v); *outp = NULL; return -1; }
    *outp = s.v;
    return n;
}

int64_t rbo_scan_count(const char *seq, int64_t L, int min_mlen, int max_mlen, int64_t counts[4]) {
    sink_t s;
    memset(&s, 0, sizeof s);
    s.count_only = 1;
    int64_t n = scan_impl(seq, L, min_mlen, max_mlen, &s);
    for (int i = 0; i < 4; i++) counts[i] = s.counts[i];
    return n;
}

void rbo_pack(const char *seq, int64_t L, uint32_t *hi, uint32_t *lo, uint32_t *nn) {
    int64_t nw = (L + 31) / 32;
    memset(hi, 0, (size_t)nw * 4); memset(lo, 0, (size_t)nw * 4); memset(nn, 0, (size_t)nw * 4);
    for (int64_t i = 0; i < L; i++) {
        uint32_t b = 1u << (i & 31);
        switch (seq[i]) {
            case 'A': case 'a': break;
            case 'C': case 'c': lo[i >> 5] |= b; break;
            case 'G': case 'g': hi[i >> 5] |= b; break;
            case 'T': case 't': hi[i >> 5] |= b; lo[i >> 5] |= b; break;
            default: nn[i >> 5] |= b; break;
        }
    }
}

void rbo_anchored_plane(const char *seq, int64_t L, int min_mlen, int max_mlen, int m, int64_t p0, int64_t p1,
                        uint8_t *outb) {
    int s_lo = (min_mlen > 2) ? min_mlen - 2 : 1, s_hi = max_mlen + 2, ns = s_hi - s_lo + 1;
    uint8_t *code = (uint8_t *)malloc((size_t)(L > 0 ? L : 1) * 2), *nn = code + (L > 0 ? L : 1);
    encode(seq, L, code, nn);
    yctx_t c = {code, L, s_lo, (L + 63) / 64 + 1, NULL};
    uint64_t *A = (uint64_t *)calloc((size_t)(c.wpl * ns), sizeof(uint64_t));
    int lo = (m > 2) ? m - 2 : 1;
    for (int s = lo; s <= m + 2; s++)
        if (s != m) anchors_for_shift(code, L, s, A + (int64_t)(s - s_lo) * c.wpl);
    c.A = A;
    for (int64_t p = p0; p < p1; p++) outb[p - p0] = (uint8_t)Y(&c, m, p);
    free(A);
    free(code);
}

void rbo_free(void *p) { free(p); }
