/* ribbit-b200 — C ABI of the B200 seed-scanning library (libribbit_scan.so).
 *
 * This is the drop-in boundary for ribbit's tandem-repeat seed-scanning hot path. The reference has no FFI; the path
 * sits inside processSequence (/root/reference/fasta_utils.cpp:78-170), which packs the contig, builds one match plane
 * per shift and calls
 *     processShiftXORsPerfect             parse_perfect_shiftxor.h:10      (parse_perfect_shiftxor.cpp:146)
 *     processShiftXORswithSubstitutions   parse_substitute_shiftxor.h:9    (parse_substitute_shiftxor.cpp:391)
 *     generateAnchoredShiftXORs           parse_anchored_shiftxor.h:10     (parse_anchored_shiftxor.cpp:20)
 *     processShiftXORsAnchored            parse_anchored_shiftxor.h:14     (parse_anchored_shiftxor.cpp:538)
 * Those scan loops hand every candidate interval, one at a time and in a fixed order, to the order-dependent host
 * merges addSeedToSeedPositions{Perfect,Substitutions,Anchored} (parse_perfect_shiftxor.cpp:47,
 * parse_substitute_shiftxor.cpp:18, parse_anchored_shiftxor.cpp:113). This library replaces everything up to that
 * hand-off: it returns, per contig, the three candidate streams in exactly the reference's call order.
 * Around that path it also offers: rb_load_fasta (the reader loop of main, ribbit.cpp:269-280, on the device),
 * rb_filter_seeds (the gate at the top of processSeed / processSeedMotifWise, parse_seed.cpp:344-367),
 * rb_motif_rows (the row search of mostFrequentLongerMotif, parse_seed.cpp:153-256), rb_get_planes /
 * rb_get_anchor_planes (the planes the host-side consumers read) and rb_set_word_range (one contig over several GPUs).
 *
 * There is no CPU fallback behind this ABI: rb_create fails when no CUDA device is usable.
 * Plain C types only; one context per GPU, one host thread per context.
 */
#ifndef RIBBIT_SCAN_H
#define RIBBIT_SCAN_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_ABI_VERSION 1

/* error codes (negative returns) */
#define RB_OK 0
#define RB_E_ARG (-1)      /* bad argument */
#define RB_E_CUDA (-2)     /* CUDA runtime error; rb_last_error() has the text */
#define RB_E_NOMEM (-3)    /* host or device allocation failed */
#define RB_E_STATE (-4)    /* call order violated (e.g. rb_scan before rb_load_contigs) */
#define RB_E_RANGE (-5)    /* contig too long (>= 2^31 - 2^12 bases) or motif range unsupported */

/* Scan parameters. min_mlen/max_mlen are ribbit's -m / -M (ribbit.cpp:85-86); shifts are derived as
 * ribbit.cpp:240-243. The remaining reference constants are fixed in the reference (window 8, thresholds 7 then 6,
 * anchor 3: ribbit.cpp:191, fasta_utils.cpp:165) and are therefore not parameters; `-p` is never read by the
 * reference (ribbit.cpp:92) and has no counterpart here. */
typedef struct rb_params {
    int32_t min_mlen;     /* >= 1 */
    int32_t max_mlen;     /* >= min_mlen, <= 1000 */
    int32_t chunk_words;  /* 0 = automatic; words (32 bases) of a contig scanned by one warp item */
    int32_t reserved;     /* diagnostic switches, 0 in production: bit 0 = never enter the tight loop, bit 1 = do not skip
                           * chunks that lie inside an N run. Both only change which code path computes the (identical)
                           * result; values above 3 are rejected (RB_E_ARG) */
} rb_params;

/* One candidate record of a stream.
 *   flags == 0            the reference calls addSeedToSeedPositions*(start, end, mlen, ...) here
 *   RB_REC_DROPPED        same call, but end-start is below the consumer's length cutoff
 *                         (parse_substitute_shiftxor.cpp:44, parse_anchored_shiftxor.cpp:153): only its cursor
 *                         side effect (…:34-42, …:133-151) matters. Such calls are reported only where they can change
 *                         a later cursor: next to N runs and contig ends; elsewhere they are elided.
 *   RB_REC_PSEUDO         no reference call; `end` is the largest `end` among the calls elided before this point
 *                         (-1 if none). The consumer advances its cursors with it. start = -1, mlen = 0.
 *   RB_REC_NOCOMMIT       anchored tail-flush call whose returned cursors the reference discards
 *                         (parse_anchored_shiftxor.cpp:688-719).
 * `time` is the position the reference's scan loop is at when it makes the call (32*ceil(L/32) for the tail flush). */
typedef struct rb_rec {
    int32_t start;
    int32_t end;
    uint16_t mlen;
    uint16_t flags;
    int32_t time;
} rb_rec;
#define RB_REC_DROPPED 1
#define RB_REC_PSEUDO 2
#define RB_REC_NOCOMMIT 4

#define RB_STREAM_PERFECT 0
#define RB_STREAM_SUBST 1
#define RB_STREAM_ANCHORED 2

/* Result of rb_scan/rb_fetch. Arrays are owned by the context (pinned host memory) and stay valid until the next
 * rb_load_contigs / rb_scan / rb_fetch / rb_destroy. contig_off[s] has n_contigs+1 entries: records of contig c
 * in stream s are rec[s][contig_off[s][c] .. contig_off[s][c+1]). */
typedef struct rb_streams {
    int32_t n_contigs;
    int32_t reserved;
    const rb_rec *rec[3];
    const int64_t *contig_off[3];
    int64_t n[3];
} rb_streams;

/* Device timings of the last rb_scan_device, from CUDA events on the context's stream (milliseconds). */
typedef struct rb_timing {
    float pack_ms;    /* ASCII -> planes */
    float scan_ms;    /* match planes + perfect / substitution / anchored machines */
    float merge_ms;   /* ordered compaction into the three streams */
    float total_ms;   /* first kernel start -> last kernel end */
    int32_t launches; /* kernels launched */
    int32_t restarts; /* warm-up restarts inside the scan (diagnostic) */
    int32_t retries;  /* scan re-launches because a record buffer was too small */
    int32_t reserved;
} rb_timing;

typedef struct rb_ctx rb_ctx;

int rb_abi_version(void);

/* Creates a context on CUDA device `device`. NULL on failure (rb_last_error(NULL)). */
rb_ctx *rb_create(int device, const rb_params *params);
void rb_destroy(rb_ctx *ctx);

/* Text of the last error of `ctx` (or of the last failed rb_create when ctx == NULL). Never NULL. */
const char *rb_last_error(const rb_ctx *ctx);

/* Copies a batch of contigs to the device (host pointer `ascii`; contig i is ascii[offsets[i] .. offsets[i]+lengths[i]))
 * and packs them (fasta_utils.cpp:90-115). Replaces the previous batch. The caller keeps `ascii`. */
int rb_load_contigs(rb_ctx *ctx, const char *ascii, const int64_t *offsets, const int32_t *lengths, int32_t n);

/* Same, but `ascii_dev` is a device pointer on this context's device (no copy). The buffer must stay valid until the
 * next load. Used when the sequence is already resident in HBM. */
int rb_load_contigs_device(rb_ctx *ctx, const void *ascii_dev, const int64_t *offsets, const int32_t *lengths, int32_t n);

/* FASTA text -> contigs, parsed on the device (K0). Replaces the reference's reader loop ribbit.cpp:269-280 + the pack
 * loop: `text` (host memory, nbytes) is copied to the device as it is; lines whose first byte is '>' are headers, every
 * other line is sequence and is appended verbatim ('\n' is the only separator: a '\r' stays and becomes an N base, as in
 * the reference). Records follow the reference's quirks: a header closes the record in front of it only if that record
 * has sequence (an empty record's name is simply replaced), the text in front of the first header is an unnamed record,
 * and the last record is produced even when empty (ribbit.cpp:280 calls processSequence unconditionally). The records
 * become the loaded contigs (as after rb_load_contigs); rb_fasta_records returns, per record, where its name sits in
 * `text` (the bytes between '>' and the first ' ' or the end of the line, ribbit.cpp:275; name_off = -1: unnamed). */
typedef struct rb_fasta_record {
    int64_t name_off;
    int32_t name_len;
    int32_t length;  /* bases */
} rb_fasta_record;
int rb_load_fasta(rb_ctx *ctx, const char *text, int64_t nbytes, int32_t *n_records);
int rb_fasta_records(rb_ctx *ctx, rb_fasta_record *out, int32_t capacity);

/* One contig over several GPUs (unit = (contig, chunk)). After loading a batch of exactly ONE contig,
 * rb_set_word_range restricts the following scans to the 32-base words [word_first, word_last) of it (word_last = -1:
 * to the end; the part that ends at the last word also owns the tail flush). The streams then hold exactly the
 * candidates whose emission time falls into those words, in order, so the parts of a partition of the contig
 * concatenate to the full streams — with one fix-up: a PSEUDO record carries the largest end of the candidates elided
 * since the contig start, and a part only knows its own. rb_get_elided_max returns, for the substitution and the
 * anchored stream, the largest elided end of the scanned part (-1: none); the host raises every PSEUDO end of a later
 * part to the maximum over the earlier parts (ribbit_b200/shard.py: stitch_parts). Every part needs the whole contig
 * loaded: warm-up and N-run handling read the words in front of the range. */
int rb_set_word_range(rb_ctx *ctx, int32_t word_first, int32_t word_last);
int rb_get_elided_max(rb_ctx *ctx, int64_t out[2]);

/* Runs pack + scan + ordered compaction on the device; results stay in device memory. */
int rb_scan_device(rb_ctx *ctx);

/* Copies the streams of the last rb_scan_device to pinned host memory. */
int rb_fetch(rb_ctx *ctx, rb_streams *out);

/* Compact form of the same streams for hosts that are bound by the PCIe / memory traffic of the result (8 instead of 16
 * bytes per candidate): `time` is dropped (the order of the array is the call order), `end` = start + len.
 *   len == 0xFFFF          the candidate is at least 65535 long: its `end` is in the `long_end` list of the stream
 *                          (entries {index into rec8, end}, ascending index)
 *   flags & RB_REC_PSEUDO  `start` holds the PSEUDO record's `end` value, len = 0 */
typedef struct rb_rec8 {
    int32_t start;
    uint16_t len;
    uint16_t mf;  /* mlen | flags << 12 */
} rb_rec8;
typedef struct rb_long_end {
    int64_t index;
    int64_t end;
} rb_long_end;
typedef struct rb_streams8 {
    int32_t n_contigs;
    int32_t reserved;
    const rb_rec8 *rec[3];
    const int64_t *contig_off[3];
    int64_t n[3];
    const rb_long_end *long_end[3];
    int64_t n_long[3];
} rb_streams8;
int rb_fetch_compact(rb_ctx *ctx, rb_streams8 *out);

/* rb_scan_device + rb_fetch. */
int rb_scan(rb_ctx *ctx, rb_streams *out);

/* Record counts of the last rb_scan_device without copying the records. */
int rb_counts(rb_ctx *ctx, int64_t n[3]);

int rb_get_timing(const rb_ctx *ctx, rb_timing *out);

/* Diagnostic for the roofline: LOP3 + funnel-shift operations per second (32 lanes counted per warp instruction)
 * this GPU sustains, measured with a register-only microbenchmark of the scan's instruction mix. */
int rb_measure_int_peak(rb_ctx *ctx, double *ops_per_s);
/* The same microbenchmark with four instruction mixes: [0] funnel shifts + LOP3 (= rb_measure_int_peak), [1] LOP3 only,
 * [2] funnel shifts only, [3] LOP3 + IMAD (logic pipe and multiply-add pipe together). */
int rb_measure_int_peak_modes(rb_ctx *ctx, double ops_per_s[4]);
/* Diagnostic: per (chunk, motif band) work item of the scan kernel {start ns, end ns, words through the general path,
 * warm-up restarts << 32 | bit-serial words}. The first call switches the recording on (returns RB_E_STATE); call again
 * after the next rb_scan_device. */
int rb_debug_item_clocks(rb_ctx *ctx, int64_t *out, int64_t capacity_items, int64_t *n_items);

/* K5 — the per-seed gate of processSeed / processSeedMotifWise (parse_seed.cpp:344-367,
 * parse_smallmotif_seed.cpp:216-235), batched: for each seed (contig, start, end, mlen), end <= L (the merges clamp seeds to
 * end + mlen <= L, parse_perfect_shiftxor.cpp:137) returns the length of the seed sequence after truncation
 * at the first N in [start, end + mlen) and the longest run of 1s of the anchored plane B_mlen
 * (fasta_utils.cpp:143-161) over [start, end). The reference drops a seed when longest_run < 3 (ribbit.cpp:191). */
typedef struct rb_seed {
    int32_t contig, start, end, mlen;
} rb_seed;
typedef struct rb_seedinfo {
    int32_t seq_len;      /* seed_sequence_length (parse_seed.cpp:347-353) */
    int32_t longest_run;  /* longestContinuousMatches (parse_seed.cpp:26-44) */
} rb_seedinfo;
int rb_filter_seeds(rb_ctx *ctx, const rb_seed *seeds, int64_t n, rb_seedinfo *out);

/* K7: the row search of mostFrequentLongerMotif (parse_seed.cpp:153-256; the reference calls it for motif sizes > 10,
 * parse_seed.cpp:388-390), batched on the device. Here a seed is given as the function's own arguments: start =
 * seed_start, end = seed_start + seed_sequence_length (the N-truncated length rb_filter_seeds returns; the region
 * [start, end) must not contain N, as in the reference), mlen = motif_length >= 3 (with
 * smaller sizes the reference's unit walk, parse_seed.cpp:197-198, need not advance). Returns per seed the row the reference
 * settles on (mmotif_index, parse_seed.cpp:241: the first row with the largest diagonal-match count, 0 when no row scores
 * or the seed is shorter than the motif) and that count. The consensus motif is the mlen bases starting at row
 * (parse_seed.cpp:246-253). */
typedef struct rb_motifrow {
    int32_t row;    /* mmotif_index */
    int32_t count;  /* max_count */
} rb_motifrow;
int rb_motif_rows(rb_ctx *ctx, const rb_seed *seeds, int64_t n, rb_motifrow *out);

/* Anchor planes A_s (generateAnchoredShiftXORs, parse_anchored_shiftxor.cpp:20-56: runs of 1s of the match plane X_s
 * with 3 <= length < 2s that are closed inside the scanned range) of contig c for shifts shift_lo..shift_hi, computed on
 * the device: out[(s - shift_lo) * ceil(L/32) + w], bit i of word w = position 32*w+i. The host side of the reference
 * (order-dependent merges, per-seed stage) reads B_m = X_m | A_{m-2} | A_{m-1} | A_{m+1} | A_{m+2} (fasta_utils.cpp:143-161). */
int rb_get_anchor_planes(rb_ctx *ctx, int32_t contig, int32_t shift_lo, int32_t shift_hi, uint32_t *out);

/* Copies the packed planes of contig c to host arrays of ceil(L/32) words (bit i of word w = position 32*w+i):
 * hi/lo = the two code bits (A=00 C=01 G=10 T=11), nn = N plane. Any pointer may be NULL. */
int rb_get_planes(rb_ctx *ctx, int32_t contig, uint32_t *hi, uint32_t *lo, uint32_t *nn);

#ifdef __cplusplus
}
#endif
#endif
