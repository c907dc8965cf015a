// ribbit-b200: C ABI of the scan library (include/ribbit_scan.h). Host-side orchestration only: geometry tables,
// buffer management, kernel launches, CUDA-event timing. No CPU implementation of the scan exists behind this ABI.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ribbit_scan.h"
#include "kernels.h"

using namespace rb;

static_assert(sizeof(rb_rec) == sizeof(Rec), "rb_rec and rb::Rec must have the same layout");

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

std::string g_create_error;

}  // namespace

struct rb_ctx {
    int device = 0;
    rb_params params{};
    BandLayout lay{};
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // rb_load_fasta: chunked H2D of the text, overlapped with the first kernel pass
    cudaEvent_t copy_ev = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;

    // host geometry
    std::vector<Contig> contigs;
    std::vector<long long> plane_start, bucket_base;
    std::vector<Chunk> chunks;
    std::vector<long long> item_base;
    std::vector<int> item_cap;
    long long raw_cap = 0;
    long long dst_cap = 0;
    bool loaded = false, scanned = false;
    bool ascii_external = false;
    const void* ascii_dev_ext = nullptr;

    DevBuf d_ascii, d_contigs, d_plane_start, d_bucket_base, d_planes, d_nmask, d_chunks, d_item_base, d_item_cap, d_item_count,
        d_meta, d_bcnt, d_item_clk, d_raw, d_counters, d_partial, d_dst, d_contig_off, d_totals, d_seeds, d_seedinfo, d_dst8, d_long, d_mitems, d_mkeys, d_text, d_ftiles, d_finfo, d_ftot, d_hpos, d_hseq;
    DevBatch batch{};

    // pinned host results
    void* h_rec = nullptr;
    size_t h_rec_cap = 0;
    void* h_off = nullptr;
    size_t h_off_cap = 0;
    void* h_long = nullptr;  // pinned: side list of long candidates + its counters
    size_t h_long_cap = 0;
    std::vector<rb_long_end> long_sorted[3];
    long long* h_small = nullptr;  // pinned: totals[3], counters[2]
    long long totals[3] = {0, 0, 0};
    rb_timing timing{};
    std::vector<rb_fasta_record> fasta_records;
};

namespace {

int fail(rb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define RB_CUDA(c, call)                                                                                      \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail((c), e_ == cudaErrorMemoryAllocation ? RB_E_NOMEM : RB_E_CUDA, "%s: %s", #call,       \
                        cudaGetErrorString(e_));                                                              \
    } while (0)

int ensure(rb_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return RB_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    const size_t want = std::max<size_t>(bytes + bytes / 8, 256);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(&b.p, std::max<size_t>(bytes, 256));
        if (e != cudaSuccess) { b.p = nullptr; return fail(c, RB_E_NOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e)); }
        b.cap = std::max<size_t>(bytes, 256);
    } else {
        b.cap = want;
    }
    return RB_OK;
}

template <class T>
int upload(rb_ctx* c, DevBuf& b, const std::vector<T>& v) {
    const int rc = ensure(c, b, v.size() * sizeof(T));
    if (rc) return rc;
    if (!v.empty()) RB_CUDA(c, cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return RB_OK;
}

int ensure_pinned(rb_ctx* c, void*& p, size_t& cap, size_t bytes) {
    if (bytes <= cap && p) return RB_OK;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    const size_t want = std::max<size_t>(bytes + bytes / 8, 4096);
    RB_CUDA(c, cudaHostAlloc(&p, want, cudaHostAllocDefault));
    cap = want;
    return RB_OK;
}

// raw slots per word reserved for one (chunk, band) item (a fast word needs one slot per motif size and stream with
// candidates, a slow word one per candidate); an item that needs more triggers an exact re-run
int records_per_word(const BandLayout& lay, int band) {
    const int m0 = band_m0(lay, band);
    return m0 <= 12 ? 8 : 4;
}

// Cuts the contigs' words into chunks. range_last >= 0: single-contig batch, only words [range_first, range_last).
void make_chunks(rb_ctx* c, long long total_words, int range_first, int range_last) {
    const BandLayout& lay = c->lay;
    int cw = c->params.chunk_words;
    if (cw <= 0) {
        // one-warp items (chunk x band), 16 resident per SM: about 16 waves of them keep the tail short; chunks of at least
        // 384 words amortise the warm-up in front of every chunk (about 25 word steps), longer ones are not needed for that
        const long long target_items = 148ll * 16 * 16;
        cw = (int)std::min<long long>(8192, std::max<long long>(384, total_words * lay.nbands / target_items));
    }
    c->chunks.clear();
    const int n = (int)c->contigs.size();
    for (int i = 0; i < n; ++i) {
        const int nw = c->contigs[i].nw;
        if (nw == 0) { c->chunks.push_back(Chunk{i, 0, 0, 1}); continue; }
        const int lo = range_last >= 0 ? range_first : 0, hi = range_last >= 0 ? range_last : nw;
        for (int w = lo; w < hi; w += cw) c->chunks.push_back(Chunk{i, w, std::min(hi, w + cw), std::min(hi, w + cw) >= nw ? 1 : 0});
    }
}

int build_geometry(rb_ctx* c, const int64_t* offsets, const int32_t* lengths, int32_t n) {
    const BandLayout& lay = c->lay;
    c->contigs.resize(n);
    c->plane_start.resize((size_t)n + 1);
    c->bucket_base.resize((size_t)n + 1);
    long long pw = 0, nb = 0, total_words = 0;
    for (int i = 0; i < n; ++i) {
        if (lengths[i] < 0 || (long long)lengths[i] > 0x7FFFFFFFll - 4096) return fail(c, RB_E_RANGE, "contig %d: length %d out of range", i, lengths[i]);
        if (offsets[i] < 0) return fail(c, RB_E_ARG, "contig %d: negative offset", i);
        Contig cg;
        cg.L = lengths[i];
        cg.nw = (int32_t)(((long long)lengths[i] + 31) / 32);
        cg.ascii_off = offsets[i];
        c->plane_start[i] = pw;
        cg.word_base = pw + 1;
        pw += 1 + cg.nw + lay.guard;
        c->bucket_base[i] = nb;
        nb += cg.nw + 1;
        total_words += cg.nw;
        c->contigs[i] = cg;
    }
    c->plane_start[n] = pw;
    c->bucket_base[n] = nb;

    make_chunks(c, total_words, 0, -1);
    return RB_OK;
}

void size_items(rb_ctx* c, const std::vector<int>* exact_counts) {
    const BandLayout& lay = c->lay;
    const size_t n_items = c->chunks.size() * (size_t)lay.nbands;
    c->item_base.resize(n_items);
    c->item_cap.resize(n_items);
    long long base = 0;
    for (size_t ci = 0; ci < c->chunks.size(); ++ci) {
        const int words = c->chunks[ci].w1 - c->chunks[ci].w0;
        for (int b = 0; b < lay.nbands; ++b) {
            const size_t it = ci * lay.nbands + b;
            long long cap = 64 + (long long)words * records_per_word(lay, b);
            if (exact_counts) cap = std::max<long long>(cap, (*exact_counts)[it]);
            cap = std::min<long long>(cap, 0x7FFFFFF0ll);
            c->item_base[it] = base;
            c->item_cap[it] = (int)cap;
            base += cap;
        }
    }
    c->raw_cap = base;
}

int upload_geometry(rb_ctx* c) {
    int rc;
    if ((rc = upload(c, c->d_contigs, c->contigs))) return rc;
    if ((rc = upload(c, c->d_plane_start, c->plane_start))) return rc;
    if ((rc = upload(c, c->d_bucket_base, c->bucket_base))) return rc;
    if ((rc = upload(c, c->d_chunks, c->chunks))) return rc;
    return RB_OK;
}

int upload_items(rb_ctx* c) {
    int rc;
    if ((rc = upload(c, c->d_item_base, c->item_base))) return rc;
    if ((rc = upload(c, c->d_item_cap, c->item_cap))) return rc;
    if (c->raw_cap >= 0xFFFFFFF0ll) return fail(c, RB_E_RANGE, "batch too large: %lld candidate slots (load fewer bases per batch)", c->raw_cap);
    if ((rc = ensure(c, c->d_raw, (size_t)c->raw_cap * sizeof(Rec)))) return rc;
    if ((rc = ensure(c, c->d_item_count, c->item_base.size() * sizeof(int)))) return rc;
    c->batch.item_base = (const long long*)c->d_item_base.p;
    c->batch.item_cap = (const int*)c->d_item_cap.p;
    c->batch.item_count = (int*)c->d_item_count.p;
    c->batch.raw = (Rec*)c->d_raw.p;
    return RB_OK;
}

int finish_load(rb_ctx* c, const void* ascii_dev, int32_t n) {
    int rc;
    if ((rc = upload_geometry(c))) return rc;
    size_items(c, nullptr);
    DevBatch& b = c->batch;
    b = DevBatch{};
    b.lay = c->lay;
    b.n_contigs = n;
    b.n_chunks = (int)c->chunks.size();
    b.n_items = (long long)c->chunks.size() * c->lay.nbands;
    b.n_buckets = c->bucket_base[n];
    b.gb_first = 0;
    b.n_active = b.n_buckets;
    b.n_plane_words = c->plane_start[n];
    b.warm0 = WARMUP_WORDS;
    if (const char* e = getenv("RB_WARM0")) b.warm0 = std::max(1, atoi(e));  // diagnostics
    b.debug = c->params.reserved;
    b.n_merge_blocks = (int)((b.n_buckets + MERGE_BLOCK - 1) / MERGE_BLOCK);
    if ((rc = ensure(c, c->d_planes, (size_t)b.n_plane_words * sizeof(PlaneWord)))) return rc;
    if ((rc = ensure(c, c->d_meta, (size_t)b.n_buckets * c->lay.nbands * sizeof(Meta)))) return rc;
    if ((rc = ensure(c, c->d_bcnt, (size_t)std::max<long long>(b.n_buckets, 1) * c->lay.nbands * sizeof(uint32_t)))) return rc;
    if ((rc = ensure(c, c->d_counters, 4 * sizeof(int)))) return rc;
    if ((rc = ensure(c, c->d_partial, ((size_t)b.n_merge_blocks + 2 + (size_t)b.n_merge_blocks / MERGE_SEG + 1) * sizeof(BlockPartial)))) return rc;
    if ((rc = ensure(c, c->d_contig_off, 3 * ((size_t)n + 1) * sizeof(long long)))) return rc;
    if ((rc = ensure(c, c->d_totals, 3 * sizeof(long long)))) return rc;
    b.ascii = (const uint8_t*)ascii_dev;
    b.contigs = (const Contig*)c->d_contigs.p;
    b.plane_start = (const long long*)c->d_plane_start.p;
    b.bucket_base = (const long long*)c->d_bucket_base.p;
    b.planes = (PlaneWord*)c->d_planes.p;
    if ((rc = ensure(c, c->d_nmask, (size_t)((b.n_plane_words + 31) / 32 + 1) * sizeof(uint32_t)))) return rc;
    b.nmask = (uint32_t*)c->d_nmask.p;
    b.chunks = (const Chunk*)c->d_chunks.p;
    b.meta = (Meta*)c->d_meta.p;
    b.bcnt = (uint32_t*)c->d_bcnt.p;
    b.counters = (int*)c->d_counters.p;
    b.partial = (BlockPartial*)c->d_partial.p;
    b.contig_off = (long long*)c->d_contig_off.p;
    b.totals = (long long*)c->d_totals.p;
    if ((rc = upload_items(c))) return rc;
    c->loaded = true;
    c->scanned = false;
    return RB_OK;
}

}  // namespace

extern "C" {

int rb_abi_version(void) { return RB_ABI_VERSION; }

const char* rb_last_error(const rb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

rb_ctx* rb_create(int device, const rb_params* params) {
    if (!params || params->min_mlen < 1 || params->max_mlen < params->min_mlen || params->max_mlen > 1000 ||
        params->max_mlen - params->min_mlen + 1 > 224 || params->reserved < 0 || params->reserved > 3 || params->chunk_words < 0) {
        fail(nullptr, RB_E_ARG, "rb_create: bad parameters (need 1 <= min_mlen <= max_mlen <= 1000, at most 224 motif sizes)");
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        fail(nullptr, RB_E_CUDA, "rb_create: CUDA device %d not available (%s); this library has no CPU path", device,
             e != cudaSuccess ? cudaGetErrorString(e) : "no such device");
        return nullptr;
    }
    rb_ctx* c = new rb_ctx();
    c->device = device;
    c->params = *params;
    c->lay = make_layout(params->min_mlen, params->max_mlen);
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 5; ++i) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
    if (ok) ok = cudaHostAlloc((void**)&c->h_small, 8 * sizeof(long long), cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        fail(nullptr, RB_E_CUDA, "rb_create: %s", cudaGetErrorString(cudaGetLastError()));
        rb_destroy(c);
        return nullptr;
    }
    return c;
}

void rb_destroy(rb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->d_ascii, &c->d_contigs, &c->d_plane_start, &c->d_bucket_base, &c->d_planes, &c->d_nmask, &c->d_chunks,
                      &c->d_item_base, &c->d_item_cap, &c->d_item_count, &c->d_meta, &c->d_bcnt, &c->d_item_clk, &c->d_raw, &c->d_counters,
                      &c->d_partial, &c->d_dst, &c->d_contig_off, &c->d_totals, &c->d_seeds, &c->d_seedinfo, &c->d_dst8, &c->d_long, &c->d_mitems, &c->d_mkeys,
                      &c->d_text, &c->d_ftiles, &c->d_finfo, &c->d_ftot, &c->d_hpos, &c->d_hseq};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    if (c->h_rec) cudaFreeHost(c->h_rec);
    if (c->h_off) cudaFreeHost(c->h_off);
    if (c->h_long) cudaFreeHost(c->h_long);
    if (c->h_small) cudaFreeHost(c->h_small);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->copy_ev) cudaEventDestroy(c->copy_ev);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rb_load_contigs(rb_ctx* c, const char* ascii, const int64_t* offsets, const int32_t* lengths, int32_t n) {
    if (!c) return RB_E_ARG;
    if (n < 0 || (n > 0 && (!ascii || !offsets || !lengths))) return fail(c, RB_E_ARG, "rb_load_contigs: null argument");
    RB_CUDA(c, cudaSetDevice(c->device));
    int rc = build_geometry(c, offsets, lengths, n);
    if (rc) return rc;
    long long hi = 0;
    for (int i = 0; i < n; ++i) hi = std::max<long long>(hi, offsets[i] + lengths[i]);
    if ((rc = ensure(c, c->d_ascii, (size_t)hi + 64))) return rc;
    if (hi > 0) RB_CUDA(c, cudaMemcpyAsync(c->d_ascii.p, ascii, (size_t)hi, cudaMemcpyHostToDevice, c->stream));
    c->ascii_external = false;
    rc = finish_load(c, c->d_ascii.p, n);
    if (rc) return rc;
    RB_CUDA(c, cudaStreamSynchronize(c->stream));  // the caller may reuse `ascii` and the geometry vectors
    return RB_OK;
}

int rb_load_contigs_device(rb_ctx* c, const void* ascii_dev, const int64_t* offsets, const int32_t* lengths, int32_t n) {
    if (!c) return RB_E_ARG;
    if (n < 0 || (n > 0 && (!ascii_dev || !offsets || !lengths))) return fail(c, RB_E_ARG, "rb_load_contigs_device: null argument");
    RB_CUDA(c, cudaSetDevice(c->device));
    int rc = build_geometry(c, offsets, lengths, n);
    if (rc) return rc;
    c->ascii_external = true;
    rc = finish_load(c, ascii_dev, n);
    if (rc) return rc;
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RB_OK;
}

int rb_load_fasta(rb_ctx* c, const char* text, int64_t nbytes, int32_t* n_records) {
    if (!c) return RB_E_ARG;
    if (nbytes < 0 || (nbytes > 0 && !text)) return fail(c, RB_E_ARG, "rb_load_fasta: null argument");
    RB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const long long nt = fasta_tiles(nbytes);
    int rc;
    if ((rc = ensure(c, c->d_text, (size_t)nbytes + 64))) return rc;
    if ((rc = ensure(c, c->d_ftiles, (size_t)std::max<long long>(nt, 1) * sizeof(int4)))) return rc;
    if ((rc = ensure(c, c->d_finfo, (size_t)std::max<long long>(nt, 1) * sizeof(longlong2)))) return rc;
    if ((rc = ensure(c, c->d_ftot, 2 * sizeof(long long)))) return rc;
    // the text goes over in chunks on a second stream; pass 1 (per-tile summaries) of a chunk starts as soon as the chunk
    // has arrived, while the next chunks are still on their way
    if (!c->copy_stream) RB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->copy_ev) RB_CUDA(c, cudaEventCreateWithFlags(&c->copy_ev, cudaEventDisableTiming));
    {
        const long long chunk = 2048 * fasta_tile_bytes();  // 8 MiB
        RB_CUDA(c, cudaEventRecord(c->copy_ev, st));         // the copies must not overtake earlier work on the text buffer
        RB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->copy_ev, 0));
        for (long long at = 0; at < nbytes; at += chunk) {
            const long long n = std::min<long long>(chunk, nbytes - at);
            RB_CUDA(c, cudaMemcpyAsync((char*)c->d_text.p + at, text + at, (size_t)n, cudaMemcpyHostToDevice, c->copy_stream));
            RB_CUDA(c, cudaEventRecord(c->copy_ev, c->copy_stream));
            RB_CUDA(c, cudaStreamWaitEvent(st, c->copy_ev, 0));
            launch_fasta_tiles(c->d_text.p, nbytes, at / fasta_tile_bytes(), (n + fasta_tile_bytes() - 1) / fasta_tile_bytes(), c->d_ftiles.p, st);
        }
    }
    launch_fasta_scan(c->d_ftiles.p, nbytes, c->d_finfo.p, (long long*)c->d_ftot.p, st);
    RB_CUDA(c, cudaGetLastError());
    RB_CUDA(c, cudaMemcpyAsync(c->h_small, c->d_ftot.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    RB_CUDA(c, cudaStreamSynchronize(st));
    const long long n_seq = c->h_small[0], n_hdr = c->h_small[1];
    if (n_hdr + 1 > 0x7FFFFFFFll) return fail(c, RB_E_RANGE, "rb_load_fasta: %lld records", n_hdr + 1);
    if ((rc = ensure(c, c->d_ascii, (size_t)n_seq + 64))) return rc;
    if ((rc = ensure(c, c->d_hpos, (size_t)std::max<long long>(n_hdr, 1) * sizeof(long long)))) return rc;
    if ((rc = ensure(c, c->d_hseq, (size_t)std::max<long long>(n_hdr, 1) * sizeof(long long)))) return rc;
    launch_fasta_strip(c->d_text.p, nbytes, c->d_finfo.p, c->d_ascii.p, (long long*)c->d_hpos.p, (long long*)c->d_hseq.p, st);
    RB_CUDA(c, cudaGetLastError());
    std::vector<long long> hpos((size_t)n_hdr), hseq((size_t)n_hdr);
    if (n_hdr > 0) {
        RB_CUDA(c, cudaMemcpyAsync(hpos.data(), c->d_hpos.p, (size_t)n_hdr * sizeof(long long), cudaMemcpyDeviceToHost, st));
        RB_CUDA(c, cudaMemcpyAsync(hseq.data(), c->d_hseq.p, (size_t)n_hdr * sizeof(long long), cudaMemcpyDeviceToHost, st));
    }
    RB_CUDA(c, cudaStreamSynchronize(st));
    // segment j = the sequence between header j-1 (none for j = 0) and header j; ribbit.cpp:271-279: a segment is a record
    // if it has sequence, the last one always
    c->fasta_records.clear();
    std::vector<int64_t> offsets;
    std::vector<int32_t> lengths;
    for (long long j = 0; j <= n_hdr; ++j) {
        const long long from = j == 0 ? 0 : hseq[(size_t)j - 1], to = j == n_hdr ? n_seq : hseq[(size_t)j];
        if (to == from && j != n_hdr) continue;
        if (to - from > 0x7FFFFFFFll - 4096) return fail(c, RB_E_RANGE, "rb_load_fasta: record %lld has %lld bases", j, to - from);
        rb_fasta_record r;
        r.name_off = -1; r.name_len = 0; r.length = (int32_t)(to - from);
        if (j > 0) {
            const long long p0 = hpos[(size_t)j - 1] + 1;
            long long p = p0;
            while (p < nbytes && text[p] != ' ' && text[p] != '\n') ++p;
            r.name_off = p0; r.name_len = (int32_t)std::min<long long>(p - p0, 0x7FFFFFFF);
        }
        c->fasta_records.push_back(r);
        offsets.push_back(from);
        lengths.push_back(r.length);
    }
    const int32_t n = (int32_t)lengths.size();
    if ((rc = build_geometry(c, offsets.data(), lengths.data(), n))) return rc;
    c->ascii_external = false;
    if ((rc = finish_load(c, c->d_ascii.p, n))) return rc;
    RB_CUDA(c, cudaStreamSynchronize(st));
    if (n_records) *n_records = n;
    return RB_OK;
}

int rb_fasta_records(rb_ctx* c, rb_fasta_record* out, int32_t capacity) {
    if (!c || capacity < 0 || (capacity > 0 && !out)) return RB_E_ARG;
    if ((size_t)capacity < c->fasta_records.size()) return fail(c, RB_E_ARG, "rb_fasta_records: %zu records, capacity %d", c->fasta_records.size(), capacity);
    for (size_t i = 0; i < c->fasta_records.size(); ++i) out[i] = c->fasta_records[i];
    return RB_OK;
}

int rb_set_word_range(rb_ctx* c, int32_t word_first, int32_t word_last) {
    if (!c) return RB_E_ARG;
    if (!c->loaded) return fail(c, RB_E_STATE, "rb_set_word_range: no contigs loaded");
    if (c->contigs.size() != 1) return fail(c, RB_E_STATE, "rb_set_word_range: needs a batch of exactly one contig");
    RB_CUDA(c, cudaSetDevice(c->device));
    const int nw = c->contigs[0].nw;
    if (word_last < 0) word_last = nw;
    if (word_first < 0 || word_first > word_last || word_last > nw || (nw > 0 && word_first == word_last))
        return fail(c, RB_E_ARG, "rb_set_word_range: [%d, %d) is not a non-empty range of the contig's %d words", word_first, word_last, nw);
    make_chunks(c, std::max(1, word_last - word_first), word_first, word_last);
    size_items(c, nullptr);
    int rc;
    if ((rc = upload(c, c->d_chunks, c->chunks))) return rc;
    c->batch.chunks = (const Chunk*)c->d_chunks.p;
    c->batch.n_chunks = (int)c->chunks.size();
    c->batch.n_items = (long long)c->chunks.size() * c->lay.nbands;
    // the ordered compaction covers the buckets of the range; the tail bucket belongs to the part that ends at the last word
    c->batch.gb_first = word_first;
    c->batch.n_active = (word_last - word_first) + (word_last == nw ? 1 : 0);
    c->batch.n_merge_blocks = (int)((c->batch.n_active + MERGE_BLOCK - 1) / MERGE_BLOCK);
    if ((rc = upload_items(c))) return rc;
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->scanned = false;
    return RB_OK;
}

int rb_get_elided_max(rb_ctx* c, int64_t out[2]) {
    if (!c || !out) return RB_E_ARG;
    if (!c->scanned) return fail(c, RB_E_STATE, "rb_get_elided_max: no scan result");
    if (c->contigs.size() != 1) return fail(c, RB_E_STATE, "rb_get_elided_max: needs a batch of exactly one contig");
    RB_CUDA(c, cudaSetDevice(c->device));
    BlockPartial tot{};
    if (c->batch.n_buckets > 0) {
        RB_CUDA(c, cudaMemcpyAsync(&tot, c->batch.partial + c->batch.n_merge_blocks, sizeof tot, cudaMemcpyDeviceToHost, c->stream));
        RB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    for (int s = 0; s < 2; ++s) out[s] = tot.emax[s] ? (int64_t)(tot.emax[s] & 0xFFFFFFFFull) - 1 : -1;
    return RB_OK;
}

int rb_scan_device(rb_ctx* c) {
    if (!c) return RB_E_ARG;
    if (!c->loaded) return fail(c, RB_E_STATE, "rb_scan_device: no contigs loaded");
    RB_CUDA(c, cudaSetDevice(c->device));
    DevBatch& b = c->batch;
    rb_timing tm{};
    cudaStream_t st = c->stream;
    // the final pool is sized before the step from the input size (or the previous result), so that the whole step is
    // enqueued without a host round trip; if it turns out too small only the last kernel runs again
    long long bases = 0;
    for (const Contig& cg : c->contigs) bases += cg.L;
    if (b.gb_first > 0 || b.n_active < b.n_buckets) bases = 32ll * b.n_active;
    long long want = std::max<long long>(1024 + bases / 4, c->scanned ? (c->totals[0] + c->totals[1] + c->totals[2]) * 9 / 8 : 0);
    int rc = ensure(c, c->d_dst, (size_t)want * sizeof(Rec));
    if (rc) return rc;
    b.dst = (Rec*)c->d_dst.p;
    b.dst_cap = (long long)(c->d_dst.cap / sizeof(Rec));
    RB_CUDA(c, cudaEventRecord(c->ev[0], st));
    launch_pack(b, st);
    tm.launches += b.n_plane_words ? 1 : 0;
    RB_CUDA(c, cudaEventRecord(c->ev[1], st));
    for (int attempt = 0;; ++attempt) {
        RB_CUDA(c, cudaMemsetAsync(b.counters, 0, 4 * sizeof(int), st));
        launch_scan(b, st);
        tm.launches += b.n_items ? 1 : 0;
        if (attempt == 0) RB_CUDA(c, cudaEventRecord(c->ev[2], st));
        launch_merge_count(b, st);
        launch_merge_write(b, st);
        tm.launches += b.n_buckets ? 4 : 0;
        RB_CUDA(c, cudaEventRecord(c->ev[3], st));
        RB_CUDA(c, cudaMemcpyAsync(c->h_small, b.totals, 3 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        RB_CUDA(c, cudaMemcpyAsync(c->h_small + 4, b.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        RB_CUDA(c, cudaStreamSynchronize(st));
        RB_CUDA(c, cudaGetLastError());
        const int* counters = (const int*)(c->h_small + 4);
        if (b.n_buckets == 0) { c->h_small[0] = c->h_small[1] = c->h_small[2] = 0; }
        tm.restarts += counters[1];
        if (b.n_items && counters[2]) return fail(c, RB_E_RANGE, "rb_scan_device: more candidates in one 32-base word than the bucket counters hold");
        if (b.n_items == 0 || counters[0] == 0) break;
        if (attempt >= 2) return fail(c, RB_E_CUDA, "rb_scan_device: candidate buffers still too small after an exact re-run");
        // some item produced more slots than reserved: size every item from the counts and run again
        std::vector<int> counts(c->item_base.size());
        RB_CUDA(c, cudaMemcpy(counts.data(), b.item_count, counts.size() * sizeof(int), cudaMemcpyDeviceToHost));
        size_items(c, &counts);
        rc = upload_items(c);
        if (rc) return rc;
        ++tm.retries;
    }
    for (int s = 0; s < 3; ++s) c->totals[s] = c->h_small[s];
    const long long total = c->totals[0] + c->totals[1] + c->totals[2];
    if (total > b.dst_cap) {  // the estimate was too small: the last kernel once more into a pool of the exact size
        rc = ensure(c, c->d_dst, (size_t)total * sizeof(Rec));
        if (rc) return rc;
        b.dst = (Rec*)c->d_dst.p;
        b.dst_cap = (long long)(c->d_dst.cap / sizeof(Rec));
        launch_merge_write(b, st);
        tm.launches += 1;
        ++tm.retries;
        RB_CUDA(c, cudaEventRecord(c->ev[3], st));
        RB_CUDA(c, cudaStreamSynchronize(st));
        RB_CUDA(c, cudaGetLastError());
    }
    RB_CUDA(c, cudaEventElapsedTime(&tm.pack_ms, c->ev[0], c->ev[1]));
    RB_CUDA(c, cudaEventElapsedTime(&tm.scan_ms, c->ev[1], c->ev[2]));
    RB_CUDA(c, cudaEventElapsedTime(&tm.merge_ms, c->ev[2], c->ev[3]));
    RB_CUDA(c, cudaEventElapsedTime(&tm.total_ms, c->ev[0], c->ev[3]));
    c->timing = tm;
    c->scanned = true;
    return RB_OK;
}

/* diagnostics: per (chunk, band) item of the last scan {start, end} in GPU globaltimer nanoseconds; the first call only
 * switches the recording on (returns RB_E_STATE), later calls after a scan return the n_items pairs */
int rb_debug_item_clocks(rb_ctx* c, int64_t* out, int64_t capacity_items, int64_t* n_items) {
    if (!c || !n_items) return RB_E_ARG;
    RB_CUDA(c, cudaSetDevice(c->device));
    const long long n = c->batch.n_items;
    *n_items = n;
    if (!c->batch.item_clk) {
        int rc = ensure(c, c->d_item_clk, (size_t)std::max<long long>(n, 1) * 4 * sizeof(long long));
        if (rc) return rc;
        RB_CUDA(c, cudaMemset(c->d_item_clk.p, 0, (size_t)std::max<long long>(n, 1) * 4 * sizeof(long long)));
        c->batch.item_clk = (long long*)c->d_item_clk.p;
        return fail(c, RB_E_STATE, "rb_debug_item_clocks: recording switched on; scan again");
    }
    if (!out || capacity_items < n) return fail(c, RB_E_ARG, "rb_debug_item_clocks: capacity");
    RB_CUDA(c, cudaMemcpy(out, c->d_item_clk.p, (size_t)n * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    return RB_OK;
}

int rb_counts(rb_ctx* c, int64_t n[3]) {
    if (!c || !n) return RB_E_ARG;
    if (!c->scanned) return fail(c, RB_E_STATE, "rb_counts: no scan result");
    for (int s = 0; s < 3; ++s) n[s] = c->totals[s];
    return RB_OK;
}

int rb_fetch(rb_ctx* c, rb_streams* out) {
    if (!c || !out) return RB_E_ARG;
    if (!c->scanned) return fail(c, RB_E_STATE, "rb_fetch: no scan result");
    RB_CUDA(c, cudaSetDevice(c->device));
    const long long total = c->totals[0] + c->totals[1] + c->totals[2];
    const int n = c->batch.n_contigs;
    int rc = ensure_pinned(c, c->h_rec, c->h_rec_cap, (size_t)std::max<long long>(total, 1) * sizeof(Rec));
    if (rc) return rc;
    rc = ensure_pinned(c, c->h_off, c->h_off_cap, 3 * ((size_t)n + 1) * sizeof(long long));
    if (rc) return rc;
    if (total > 0) RB_CUDA(c, cudaMemcpyAsync(c->h_rec, c->d_dst.p, (size_t)total * sizeof(Rec), cudaMemcpyDeviceToHost, c->stream));
    if (c->batch.n_buckets > 0)
        RB_CUDA(c, cudaMemcpyAsync(c->h_off, c->d_contig_off.p, 3 * ((size_t)n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    else
        memset(c->h_off, 0, 3 * ((size_t)n + 1) * sizeof(long long));
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    out->n_contigs = n;
    out->reserved = 0;
    long long base = 0;
    for (int s = 0; s < 3; ++s) {
        out->rec[s] = (const rb_rec*)c->h_rec + base;
        out->contig_off[s] = (const int64_t*)c->h_off + (size_t)s * (n + 1);
        out->n[s] = c->totals[s];
        base += c->totals[s];
    }
    return RB_OK;
}

int rb_fetch_compact(rb_ctx* c, rb_streams8* out) {
    if (!c || !out) return RB_E_ARG;
    if (!c->scanned) return fail(c, RB_E_STATE, "rb_fetch_compact: no scan result");
    RB_CUDA(c, cudaSetDevice(c->device));
    const long long total = c->totals[0] + c->totals[1] + c->totals[2];
    const int n = c->batch.n_contigs;
    const int long_cap = 1 << 16;  // candidates of >= 65535 positions per stream and batch; more than this is an error
    int rc = ensure(c, c->d_dst8, (size_t)std::max<long long>(total, 1) * 8);
    if (rc) return rc;
    rc = ensure(c, c->d_long, (size_t)3 * long_cap * 16 + 16);
    if (rc) return rc;
    rc = ensure_pinned(c, c->h_rec, c->h_rec_cap, (size_t)std::max<long long>(total, 1) * 8);
    if (rc) return rc;
    rc = ensure_pinned(c, c->h_off, c->h_off_cap, 3 * ((size_t)n + 1) * sizeof(long long));
    if (rc) return rc;
    rc = ensure_pinned(c, c->h_long, c->h_long_cap, (size_t)3 * long_cap * 16 + 16);
    if (rc) return rc;
    int* d_cnt = (int*)((char*)c->d_long.p + (size_t)3 * long_cap * 16);
    RB_CUDA(c, cudaMemsetAsync(d_cnt, 0, 16, c->stream));
    launch_compact((const Rec*)c->d_dst.p, total, c->totals[0], c->totals[0] + c->totals[1], c->d_dst8.p, (long long*)c->d_long.p,
                   long_cap, d_cnt, c->stream);
    if (total > 0) RB_CUDA(c, cudaMemcpyAsync(c->h_rec, c->d_dst8.p, (size_t)total * 8, cudaMemcpyDeviceToHost, c->stream));
    if (c->batch.n_buckets > 0)
        RB_CUDA(c, cudaMemcpyAsync(c->h_off, c->d_contig_off.p, 3 * ((size_t)n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    else
        memset(c->h_off, 0, 3 * ((size_t)n + 1) * sizeof(long long));
    int* h_cnt = (int*)((char*)c->h_long + (size_t)3 * long_cap * 16);
    RB_CUDA(c, cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    RB_CUDA(c, cudaGetLastError());
    for (int s = 0; s < 3; ++s) {
        if (h_cnt[s] > long_cap) return fail(c, RB_E_RANGE, "rb_fetch_compact: more than %d candidates of >= 65535 positions in stream %d; use rb_fetch", long_cap, s);
        c->long_sorted[s].resize((size_t)h_cnt[s]);
        if (h_cnt[s]) {
            RB_CUDA(c, cudaMemcpy(c->long_sorted[s].data(), (char*)c->d_long.p + (size_t)s * long_cap * 16, (size_t)h_cnt[s] * 16, cudaMemcpyDeviceToHost));
            std::sort(c->long_sorted[s].begin(), c->long_sorted[s].end(), [](const rb_long_end& a, const rb_long_end& b) { return a.index < b.index; });
        }
    }
    out->n_contigs = n;
    out->reserved = 0;
    long long base = 0;
    for (int s = 0; s < 3; ++s) {
        out->rec[s] = (const rb_rec8*)c->h_rec + base;
        out->contig_off[s] = (const int64_t*)c->h_off + (size_t)s * (n + 1);
        out->n[s] = c->totals[s];
        out->long_end[s] = c->long_sorted[s].data();
        out->n_long[s] = (int64_t)c->long_sorted[s].size();
        base += c->totals[s];
    }
    return RB_OK;
}

int rb_scan(rb_ctx* c, rb_streams* out) {
    const int rc = rb_scan_device(c);
    if (rc) return rc;
    return rb_fetch(c, out);
}

int rb_get_timing(const rb_ctx* c, rb_timing* out) {
    if (!c || !out) return RB_E_ARG;
    *out = c->timing;
    return RB_OK;
}

int rb_get_planes(rb_ctx* c, int32_t contig, uint32_t* hi, uint32_t* lo, uint32_t* nn) {
    if (!c) return RB_E_ARG;
    if (!c->scanned) return fail(c, RB_E_STATE, "rb_get_planes: planes exist after rb_scan_device");
    if (contig < 0 || contig >= c->batch.n_contigs) return fail(c, RB_E_ARG, "rb_get_planes: contig out of range");
    RB_CUDA(c, cudaSetDevice(c->device));
    const Contig& cg = c->contigs[contig];
    std::vector<PlaneWord> tmp((size_t)cg.nw);
    if (cg.nw) RB_CUDA(c, cudaMemcpy(tmp.data(), (const PlaneWord*)c->d_planes.p + cg.word_base, tmp.size() * sizeof(PlaneWord), cudaMemcpyDeviceToHost));
    for (int w = 0; w < cg.nw; ++w) {
        if (hi) hi[w] = tmp[w].h;
        if (lo) lo[w] = tmp[w].l;
        if (nn) {
            // bits past the contig end are 1 on the device (padding counts as N); report the reference's plane
            uint32_t m = 0xFFFFFFFFu;
            const long long rem = (long long)cg.L - 32ll * w;
            if (rem < 32) m = rem <= 0 ? 0u : ((1u << rem) - 1u);
            nn[w] = tmp[w].n & m;
        }
    }
    return RB_OK;
}

int rb_get_anchor_planes(rb_ctx* c, int32_t contig, int32_t shift_lo, int32_t shift_hi, uint32_t* out) {
    if (!c || !out) return RB_E_ARG;
    if (!c->scanned) return fail(c, RB_E_STATE, "rb_get_anchor_planes: planes exist after rb_scan_device");
    if (contig < 0 || contig >= c->batch.n_contigs) return fail(c, RB_E_ARG, "rb_get_anchor_planes: contig out of range");
    if (shift_lo < 1 || shift_hi < shift_lo || shift_hi > c->lay.s_hi) return fail(c, RB_E_ARG, "rb_get_anchor_planes: shifts out of range");
    RB_CUDA(c, cudaSetDevice(c->device));
    const Contig& cg = c->contigs[contig];
    const int ns = shift_hi - shift_lo + 1;
    const size_t n = (size_t)cg.nw * ns;
    if (n == 0) return RB_OK;
    int rc = ensure(c, c->d_seedinfo, n * sizeof(uint32_t));
    if (rc) return rc;
    launch_anchor_planes((const PlaneWord*)c->d_planes.p + cg.word_base, cg.L, cg.nw, shift_lo, ns, (uint32_t*)c->d_seedinfo.p, c->stream);
    RB_CUDA(c, cudaMemcpyAsync(out, c->d_seedinfo.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    RB_CUDA(c, cudaGetLastError());
    return RB_OK;
}

int rb_measure_int_peak(rb_ctx* c, double* ops_per_s) {
    if (!c || !ops_per_s) return RB_E_ARG;
    RB_CUDA(c, cudaSetDevice(c->device));
    int sms = 0;
    RB_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    int rc = ensure(c, c->d_counters, 4 * sizeof(int));
    if (rc) return rc;
    *ops_per_s = measure_int_peak(c->stream, (uint32_t*)c->d_counters.p, sms);
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    RB_CUDA(c, cudaGetLastError());
    return RB_OK;
}

int rb_measure_int_peak_modes(rb_ctx* c, double ops_per_s[4]) {
    if (!c || !ops_per_s) return RB_E_ARG;
    RB_CUDA(c, cudaSetDevice(c->device));
    int sms = 0;
    RB_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    int rc = ensure(c, c->d_counters, 4 * sizeof(int));
    if (rc) return rc;
    measure_int_peak_modes(c->stream, (uint32_t*)c->d_counters.p, sms, ops_per_s);
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    RB_CUDA(c, cudaGetLastError());
    return RB_OK;
}

int rb_filter_seeds(rb_ctx* c, const rb_seed* seeds, int64_t n, rb_seedinfo* out) {
    if (!c || n < 0 || (n > 0 && (!seeds || !out))) return RB_E_ARG;
    if (!c->scanned) return fail(c, RB_E_STATE, "rb_filter_seeds: planes exist after rb_scan_device");
    RB_CUDA(c, cudaSetDevice(c->device));
    for (int64_t i = 0; i < n; ++i) {
        const rb_seed& s = seeds[i];
        if (s.contig < 0 || s.contig >= c->batch.n_contigs || s.mlen < c->lay.m_lo || s.mlen > c->lay.m_hi || s.start < 0 ||
            s.end < s.start || s.end > c->contigs[s.contig].L)
            return fail(c, RB_E_ARG, "rb_filter_seeds: seed %lld out of range", (long long)i);
    }
    if (n == 0) return RB_OK;
    int rc = ensure(c, c->d_seeds, (size_t)n * sizeof(rb_seed));
    if (rc) return rc;
    rc = ensure(c, c->d_seedinfo, (size_t)n * sizeof(rb_seedinfo));
    if (rc) return rc;
    RB_CUDA(c, cudaMemcpyAsync(c->d_seeds.p, seeds, (size_t)n * sizeof(rb_seed), cudaMemcpyHostToDevice, c->stream));
    launch_seed_filter(c->batch, c->d_seeds.p, n, c->d_seedinfo.p, c->stream);
    RB_CUDA(c, cudaMemcpyAsync(out, c->d_seedinfo.p, (size_t)n * sizeof(rb_seedinfo), cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    RB_CUDA(c, cudaGetLastError());
    return RB_OK;
}

int rb_motif_rows(rb_ctx* c, const rb_seed* seeds, int64_t n, rb_motifrow* out) {
    if (!c || n < 0 || (n > 0 && (!seeds || !out))) return RB_E_ARG;
    if (!c->scanned) return fail(c, RB_E_STATE, "rb_motif_rows: planes exist after rb_scan_device");
    RB_CUDA(c, cudaSetDevice(c->device));
    std::vector<int2> items;
    for (int64_t i = 0; i < n; ++i) {
        const rb_seed& s = seeds[i];
        if (s.contig < 0 || s.contig >= c->batch.n_contigs || s.mlen < 3 || s.start < 0 || s.end < s.start ||
            s.end > c->contigs[s.contig].L)
            return fail(c, RB_E_ARG, "rb_motif_rows: seed %lld out of range", (long long)i);
        const int rows = s.end - s.start - s.mlen + 1;  // parse_seed.cpp:179
        for (int r = 0; r < rows; r += MOTIF_SLAB) items.push_back(make_int2((int)i, r));
        if (items.size() > 0x7FFFFFFFull) return fail(c, RB_E_RANGE, "rb_motif_rows: more than 2^31 row slabs in one call");
    }
    if (n == 0) return RB_OK;
    int rc = ensure(c, c->d_seeds, (size_t)n * sizeof(rb_seed));
    if (rc) return rc;
    rc = ensure(c, c->d_mkeys, (size_t)n * sizeof(unsigned long long));
    if (rc) return rc;
    rc = ensure(c, c->d_mitems, items.size() * sizeof(int2));
    if (rc) return rc;
    std::vector<unsigned long long> keys((size_t)n);
    RB_CUDA(c, cudaMemcpyAsync(c->d_seeds.p, seeds, (size_t)n * sizeof(rb_seed), cudaMemcpyHostToDevice, c->stream));
    RB_CUDA(c, cudaMemcpyAsync(c->d_mitems.p, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    RB_CUDA(c, cudaMemsetAsync(c->d_mkeys.p, 0, (size_t)n * sizeof(unsigned long long), c->stream));
    launch_motif_rows(c->batch, c->d_seeds.p, c->d_mitems.p, (long long)items.size(), c->d_mkeys.p, c->stream);
    RB_CUDA(c, cudaGetLastError());
    RB_CUDA(c, cudaMemcpyAsync(keys.data(), c->d_mkeys.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    RB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < n; ++i) {
        const unsigned long long k = keys[(size_t)i];
        out[i].count = (int32_t)(k >> 32);
        out[i].row = k ? (int32_t)(0x7FFFFFFF - (uint32_t)(k & 0xFFFFFFFFull)) : 0;
    }
    return RB_OK;
}

}  // extern "C"
