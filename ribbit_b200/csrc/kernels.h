// ribbit-b200: device-side argument blocks and launch wrappers of the scan library (kernels.cu).
#ifndef RB_KERNELS_H
#define RB_KERNELS_H
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"
#include "merge_core.h"
#include "scan_core.h"

namespace rb {

// per-stream sums + prefix-max codes of one merge block (kernel M1 -> M2 -> M3)
struct BlockPartial {
    unsigned long long sum[3];
    unsigned long long emax[2];  // contig << 32 | elided_end_code, streams S and A
};

struct DevBatch {
    BandLayout lay;
    int n_contigs;
    int n_chunks;
    long long n_items;        // n_chunks * nbands
    long long n_buckets;      // sum over contigs of nw + 1
    long long gb_first;       // buckets the ordered compaction covers: [gb_first, gb_first + n_active); everything unless
    long long n_active;       //   a word range of a single contig is scanned (rb_set_word_range)
    long long n_plane_words;  // plane words including guard words
    int warm0;
    int debug;                // bit 0: no tight loop, bit 1: no all-N chunk skip (diagnostics)

    const uint8_t* ascii;
    const Contig* contigs;          // [n_contigs]
    const long long* plane_start;   // [n_contigs + 1] first plane word (the guard word in front) of each contig
    const long long* bucket_base;   // [n_contigs + 1]
    PlaneWord* planes;
    uint32_t* nmask;                // [ceil(n_plane_words / 32)] bit i of word g: plane word 32 g + i is all N / padding
    const Chunk* chunks;            // [n_chunks]
    const long long* item_base;     // [n_items] first raw record of the item
    const int* item_cap;            // [n_items]
    int* item_count;                // [n_items] records the item produced (may exceed item_cap: overflow)
    Meta* meta;                     // [nbands][n_buckets]
    uint32_t* bcnt;                 // [nbands][n_buckets] candidates per stream of (band, bucket): merge_core.h pack_counts
    Rec* raw;                       // raw slot pool
    long long* item_clk;            // diagnostics (rb_debug_item_clocks): [n_items][2] start / end of the item in globaltimer ns, or null
    int* counters;                  // [0] overflowed items, [1] warm-up restarts, [2] a count field overflowed (error)
    // merge
    BlockPartial* partial;          // [n_merge_blocks + 1 + segments]; after M2: exclusive prefixes inside the segment, [n] = totals,
                                    //   then the segments' exclusive prefixes
    int n_merge_blocks;
    Rec* dst;                       // final pool: stream P, then S, then A
    long long dst_cap;              // records the final pool holds (writes beyond it are dropped: the host re-runs M3)
    long long* contig_off;          // [3][n_contigs + 1]
    long long* totals;              // [3]
};

static const int MERGE_BLOCK = 256;
static const int MERGE_STAGE = 2560;  // slots of a merge block staged in shared memory (M3)
static const int MERGE_SEG = 4096;   // merge blocks per segment of the two-level prefix (M2)

void launch_pack(const DevBatch& b, cudaStream_t st);
void launch_scan(const DevBatch& b, cudaStream_t st);
void launch_merge_count(const DevBatch& b, cudaStream_t st);   // M1 + M2
void launch_merge_write(const DevBatch& b, cudaStream_t st);   // M3
// 16-byte -> 8-byte records; long_list[s][k] = {index in stream s, end}, long_cnt[3]
void launch_compact(const Rec* src, long long n, long long b1, long long b2, void* dst, long long* long_list, int long_cap,
                    int* long_cnt, cudaStream_t st);
// K5: seeds = int4 {contig, start, end, mlen}[n] (device), out = int2 {seq_len, longest_run}[n] (device)
void launch_seed_filter(const DevBatch& b, const void* seeds, long long n, void* out, cudaStream_t st);
// K7 (motif_kernels.cu): items = int2 {seed index, first row of the slab - seed_start}[n_items], keys[n seeds] zeroed by the
// caller; afterwards keys[i] = count << 32 | (0x7FFFFFFF - row) of the seed's best row, 0 if no row scores
static const int MOTIF_SLAB = 256;
void launch_motif_rows(const DevBatch& b, const void* seeds, const void* items, long long n_items, void* keys, cudaStream_t st);
// K0 (fasta_kernels.cu): FASTA text on the device -> sequence bytes + header table. tiles: int4[fasta_tiles], info:
// longlong2[fasta_tiles], totals: {sequence bytes, headers} (device)
long long fasta_tiles(long long nbytes);
void launch_fasta_tiles(const void* text, long long nbytes, long long tile_first, long long n_tiles, void* tiles, cudaStream_t st);
void launch_fasta_scan(const void* tiles, long long nbytes, void* info, long long* totals, cudaStream_t st);
long long fasta_tile_bytes();
void launch_fasta_strip(const void* text, long long nbytes, const void* info, void* bases, long long* hdr_pos, long long* hdr_seq,
                        cudaStream_t st);
// anchor planes A_s, s = s_lo .. s_lo+ns-1, of one contig: out[(s - s_lo) * nw + w]
void launch_anchor_planes(const PlaneWord* cw, int L, int nw, int s_lo, int ns, uint32_t* out, cudaStream_t st);
// LOP3 + SHF warp-lane operations per second the device sustains (integer-pipe roofline denominator)
double measure_int_peak(cudaStream_t st, uint32_t* scratch, int sms);
// the same for four instruction mixes: {SHF + LOP3, LOP3 only, SHF only, LOP3 + IMAD}
void measure_int_peak_modes(cudaStream_t st, uint32_t* scratch, int sms, double out[4]);

}  // namespace rb
#endif
