// ribbit-b200 host side: a drop-in replacement of ribbit's processSequence (fasta_utils.h:12, fasta_utils.cpp:59-250)
// that takes the seed scan from the B200 library (include/ribbit_scan.h) instead of the four bit-at-a-time loops
//   processShiftXORsPerfect              parse_perfect_shiftxor.cpp:146
//   processShiftXORswithSubstitutions    parse_substitute_shiftxor.cpp:391
//   generateAnchoredShiftXORs            parse_anchored_shiftxor.cpp:20
//   processShiftXORsAnchored             parse_anchored_shiftxor.cpp:538
// The order-dependent merges addSeedToSeedPositions{Perfect,Substitutions,Anchored} (+ mergeAllLists) are restated in
// seed_merge.cpp; the per-seed motif calling (processSeed, processSeedMotifWise), SSW and the CIGAR
// code stay the reference's own host C++ and are linked from its unmodified objects. Build: compile this file in
// place of fasta_utils.cpp and link libribbit_scan.so (ribbit_b200/host/Makefile).
//
// No scan happens on the CPU here: without a usable CUDA device the program stops with an error.
#include <boost/dynamic_bitset.hpp>

#include <poll.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <atomic>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <algorithm>
#include <vector>

#include "bitseq_utils.h"
#include "fasta_utils.h"
#include "global_variables.h"
#include "parse_seed.h"
#include "parse_smallmotif_seed.h"
#include "ribbit_scan.h"
#include "seed_merge.h"
#include "ssw_cpp.h"

using namespace std;
typedef vector<tuple<int, int, int, int>> SeedList;
typedef boost::dynamic_bitset<> Bitset;

// the functions of fasta_utils.cpp that other reference files may reference
void parseFai(string, int &, unordered_map<string, int> &) {}

namespace {

struct Gpu {
    rb_ctx *ctx = nullptr;
    int m_lo = 0, m_hi = 0;
    ~Gpu() { if (ctx) rb_destroy(ctx); }
};
Gpu g_gpu;

[[noreturn]] void die(const char *what, const rb_ctx *ctx) {
    cerr << "ERROR: " << what << ": " << rb_last_error(ctx) << "\n";
    exit(1);
}

rb_ctx *gpu_context() {
    if (g_gpu.ctx && (g_gpu.m_lo != MINIMUM_MLEN || g_gpu.m_hi != MAXIMUM_MLEN)) { rb_destroy(g_gpu.ctx); g_gpu.ctx = nullptr; }
    if (!g_gpu.ctx) {
        rb_params p = {MINIMUM_MLEN, MAXIMUM_MLEN, 0, 0};
        const char *dev = getenv("RIBBIT_CUDA_DEVICE");
        g_gpu.ctx = rb_create(dev ? atoi(dev) : 0, &p);
        if (!g_gpu.ctx) die("cannot create the GPU scan context", nullptr);
        g_gpu.m_lo = MINIMUM_MLEN; g_gpu.m_hi = MAXIMUM_MLEN;
    }
    return g_gpu.ctx;
}

// The planes of different shifts are independent: build them on several host cores (the reference is single-threaded,
// but nothing here touches its globals). RIBBIT_HOST_THREADS overrides the thread count.
template <class F>
void parallel_for(int n, F body) {
    const char *env = getenv("RIBBIT_HOST_THREADS");
    int nt = env ? atoi(env) : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    nt = std::max(1, std::min(nt, n));
    if (nt == 1) { for (int i = 0; i < n; ++i) body(i); return; }
    std::atomic<int> next(0);
    vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&] { for (int i = next++; i < n; i = next++) body(i); });
    for (auto &th : pool) th.join();
}

// A plane given as 32-base words (bit i of word w = position 32w+i) -> the reference's bitset, whose bit index is
// L-1-position (fasta_utils.cpp:93). Block k of the bitset holds positions L-64(k+1) .. L-64k-1 in reversed order.
inline uint64_t bits_at(const uint32_t *w, long nw, long p) {  // 64 positions starting at p (p may be negative)
    uint64_t v = 0;
    const long w0 = p >> 5;  // floor
    const int sh = (int)(p & 31);
    uint64_t a = (w0 >= 0 && w0 < nw) ? w[w0] : 0, b = (w0 + 1 >= 0 && w0 + 1 < nw) ? w[w0 + 1] : 0,
             c = (w0 + 2 >= 0 && w0 + 2 < nw) ? w[w0 + 2] : 0;
    v = (a >> sh) | (b << (32 - sh));
    if (sh) v |= c << (64 - sh);
    return v;
}
inline uint64_t rev64(uint64_t x) {
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    return __builtin_bswap64(x);
}
Bitset to_bitset(const uint32_t *w, long L) {
    const long nw = (L + 31) / 32, nb = (L + 63) / 64;
    vector<unsigned long> blocks((size_t)nb);
    for (long k = 0; k < nb; ++k) blocks[(size_t)k] = rev64(bits_at(w, nw, L - 64 * (k + 1)));
    Bitset b(blocks.begin(), blocks.end());
    b.resize((size_t)L);
    return b;
}

// ---- per-seed stage on several host cores ------------------------------------------------------------------------------
// processSeed / processSeedMotifWise keep their scratch state in globals (global_variables.h: MOTIF_*, REPEAT_CLASSES), so
// they cannot run on threads; they can run in forked worker processes: every global scratch slot is re-initialised on first
// touch inside a call (parse_smallmotif_seed.cpp:103-116), REPEAT_CLASSES memoises a pure function (bitseq_utils.cpp:195-217),
// so a seed's rows do not depend on the seeds processed before it. Workers claim parts of the seed list from a shared
// counter, write the BED rows of a part into a buffer, and hand the buffers back through an unlinked temp file; the parent
// emits the parts in order. A CUDA context does not survive fork(), so the parent keeps the GPU and serves the workers'
// single-seed rb_motif_rows requests (recursive flank seeds, parse_seed.cpp:443-463) over a socket pair.
int g_worker_fd = -1;

bool write_all(int fd, const void *p, size_t n) {
    const char *c = (const char *)p;
    while (n) {
        const ssize_t k = write(fd, c, n);
        if (k < 0) { if (errno == EINTR) continue; return false; }
        c += k; n -= (size_t)k;
    }
    return true;
}
bool read_all(int fd, void *p, size_t n) {
    char *c = (char *)p;
    while (n) {
        const ssize_t k = read(fd, c, n);
        if (k == 0) return false;
        if (k < 0) { if (errno == EINTR) continue; return false; }
        c += k; n -= (size_t)k;
    }
    return true;
}

// RIBBIT_VERBOSE=1: wall time of every stage of processSequence on stderr
struct StageClock {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    const bool on = getenv("RIBBIT_VERBOSE") != nullptr;
    void mark(const char *what) {
        const auto now = std::chrono::steady_clock::now();
        if (on) cerr << "  [stage] " << what << ": " << std::chrono::duration<double>(now - t).count() << " s\n";
        t = now;
    }
};

int host_procs() {
    const char *env = getenv("RIBBIT_HOST_PROCS");
    const int n = env ? atoi(env) : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    return std::max(1, n);
}

// Runs body(part, stream) for part = 0..n_parts-1 in `procs` forked workers and writes the parts' output to `out` in part
// order. serve(fd): the parent's answer to one request arriving on a worker's socket (returns false on EOF).
template <class Body, class Serve>
bool forked_parts(int n_parts, int procs, ostream &out, Body body, Serve serve) {
    // everything a worker needs is set up before the first fork; if the system refuses (no temp dir, descriptor or process
    // limit) the caller runs the stage in-process
    const char *tmpdir = getenv("TMPDIR") ? getenv("TMPDIR") : "/tmp";
    vector<int> files((size_t)procs, -1), sock_parent((size_t)procs, -1), sock_child((size_t)procs, -1);
    bool setup = true;
    for (int w = 0; w < procs && setup; ++w) {
        string path = string(tmpdir) + "/ribbit_gpu.XXXXXX";
        files[(size_t)w] = mkstemp(&path[0]);
        if (files[(size_t)w] < 0) { setup = false; break; }
        unlink(path.c_str());
        int sv[2];
        if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) { setup = false; break; }
        sock_parent[(size_t)w] = sv[0]; sock_child[(size_t)w] = sv[1];
    }
    std::atomic<int> *next = setup ? (std::atomic<int> *)mmap(nullptr, sizeof(std::atomic<int>), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0)
                                   : (std::atomic<int> *)MAP_FAILED;
    if (next == MAP_FAILED) {
        for (int w = 0; w < procs; ++w)
            for (int fd : {files[(size_t)w], sock_parent[(size_t)w], sock_child[(size_t)w]}) if (fd >= 0) close(fd);
        cerr << "note: per-seed workers not available (" << strerror(errno) << "), running in-process\n";
        return false;
    }
    new (next) std::atomic<int>(0);
    out.flush();
    cerr.flush();
    vector<pid_t> pids((size_t)procs, -1);
    vector<int> socks = sock_parent;
    for (int w = 0; w < procs; ++w) {
        const pid_t pid = fork();
        if (pid < 0) {  // fewer workers than planned: the parts are claimed dynamically, so the rest still gets done
            cerr << "note: fork: " << strerror(errno) << " (" << w << " per-seed workers)\n";
            for (int v = w; v < procs; ++v) { close(files[(size_t)v]); close(sock_parent[(size_t)v]); close(sock_child[(size_t)v]); }
            procs = w;
            break;
        }
        if (pid == 0) {  // worker: never touches CUDA, never returns
            for (int v = 0; v < (int)socks.size(); ++v) { close(sock_parent[(size_t)v]); if (v != w) close(sock_child[(size_t)v]); }
            g_worker_fd = sock_child[(size_t)w];
            for (int part = (*next)++; part < n_parts; part = (*next)++) {
                std::ostringstream os;
                body(part, os);
                const string text = os.str();
                const int64_t hdr[2] = {part, (int64_t)text.size()};
                if (!write_all(files[(size_t)w], hdr, sizeof hdr) || !write_all(files[(size_t)w], text.data(), text.size())) _exit(4);
            }
            _exit(0);
        }
        close(sock_child[(size_t)w]);
        pids[(size_t)w] = pid;
    }
    if (procs == 0) { munmap((void *)next, sizeof(std::atomic<int>)); return false; }
    // parent: serve GPU requests until every worker has closed its socket
    vector<pollfd> pf((size_t)procs);
    int open_socks = procs;
    for (int w = 0; w < procs; ++w) { pf[(size_t)w].fd = socks[(size_t)w]; pf[(size_t)w].events = POLLIN; }
    while (open_socks > 0) {
        if (poll(pf.data(), (nfds_t)procs, -1) < 0) { if (errno == EINTR) continue; cerr << "ERROR: poll: " << strerror(errno) << "\n"; exit(1); }
        for (int w = 0; w < procs; ++w)
            if (pf[(size_t)w].fd >= 0 && (pf[(size_t)w].revents & (POLLIN | POLLHUP | POLLERR)))
                if (!serve(pf[(size_t)w].fd)) { close(pf[(size_t)w].fd); pf[(size_t)w].fd = -1; --open_socks; }
    }
    bool ok = true;
    for (int w = 0; w < procs; ++w) {
        int status = 0;
        while (waitpid(pids[(size_t)w], &status, 0) < 0 && errno == EINTR) {}
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) ok = false;
    }
    if (!ok) { cerr << "ERROR: a per-seed worker process failed\n"; exit(1); }
    vector<string> parts((size_t)n_parts);
    for (int w = 0; w < procs; ++w) {
        const int fd = files[(size_t)w];
        lseek(fd, 0, SEEK_SET);
        int64_t hdr[2];
        while (read_all(fd, hdr, sizeof hdr)) {
            string &t = parts[(size_t)hdr[0]];
            t.resize((size_t)hdr[1]);
            if (hdr[1] && !read_all(fd, &t[0], (size_t)hdr[1])) { cerr << "ERROR: truncated worker output\n"; exit(1); }
        }
        close(fd);
    }
    for (const string &t : parts) out << t;
    munmap((void *)next, sizeof(std::atomic<int>));
    return true;
}

// K7: rows chosen by the consensus-motif search (rb_motif_rows) for the top-level seeds of the current contig, computed in
// one batch before the per-seed walk. key = seed_start | mlen << 32; value = {seed_sequence_length, row}.
struct MotifRows {
    std::unordered_map<uint64_t, std::pair<int, int>> rows;
    long hits = 0, misses = 0;
    static uint64_t key(int start, int mlen) { return (uint64_t)(uint32_t)start | ((uint64_t)(uint32_t)mlen << 32); }
} g_motif;

}  // namespace

// Replaces the reference's mostFrequentLongerMotif (parse_seed.cpp:153-256; ribbit_b200/host/Makefile renames that
// definition, so processSeed's call at parse_seed.cpp:389 binds here): the row search runs on the GPU — batched ahead for
// the top-level seeds, one rb_motif_rows call for the flank seeds processSeed recurses into (parse_seed.cpp:443-463) —
// and the motif is read off the code planes at the chosen row (parse_seed.cpp:246-253).
uint256_t mostFrequentLongerMotif(Bitset &left_bset, Bitset &right_bset, int &seed_start, int &seed_sequence_length,
                                  int &motif_length, int &sequence_length, vector<Bitset *> &) {
    int row;
    auto hit = g_motif.rows.find(MotifRows::key(seed_start, motif_length));
    if (hit != g_motif.rows.end() && hit->second.first == seed_sequence_length) {
        row = hit->second.second; ++g_motif.hits;
    } else {
        const rb_seed sd = {0, seed_start, seed_start + seed_sequence_length, motif_length};
        rb_motifrow r;
        if (g_worker_fd >= 0) {  // worker process of the per-seed stage: the parent owns the GPU
            if (!write_all(g_worker_fd, &sd, sizeof sd) || !read_all(g_worker_fd, &r, sizeof r)) { cerr << "ERROR: worker lost its parent\n"; _exit(3); }
        } else if (rb_motif_rows(g_gpu.ctx, &sd, 1, &r) != RB_OK) die("rb_motif_rows", g_gpu.ctx);
        row = r.row; ++g_motif.misses;
    }
    uint256_t unit = 0;
    for (int j = row; j < row + motif_length; ++j) {
        const size_t b = (size_t)(sequence_length - 1 - j);
        unit <<= 2;
        unit |= uint256_t((left_bset[b] ? 2 : 0) | (right_bset[b] ? 1 : 0));
    }
    return unit;
}

void processSequence(string &sequence_id, string &sequence, int window_length, int window_bitcount_threshold, int anchor_size,
                     int continuous_ones_threshold, ostream &out) {
    (void)window_length; (void)window_bitcount_threshold; (void)anchor_size;  // fixed in the reference: 8, 7 then 6, 3
    START_TIME = time(0);
    StageClock clock;
    int sequence_length = (int)sequence.length();
    rb_ctx *ctx = gpu_context();
    clock.mark("GPU context");

    // ---- the scan, on the GPU: K1 pack, K2 scan, K6 ordered streams -------------------------------------------
    const int64_t off0 = 0;
    const int32_t len0 = sequence_length;
    if (rb_load_contigs(ctx, sequence.data(), &off0, &len0, 1) != RB_OK) die("rb_load_contigs", ctx);
    rb_streams st;
    if (rb_scan(ctx, &st) != RB_OK) die("rb_scan", ctx);
    clock.mark("load + scan + fetch");
    cerr << "Generated shift XORs!\t Time elapsed:" << difftime(time(0), START_TIME) << "secs\n";

    // ---- planes for the host-side consumers (merge tie-breakers read popcounts of the match planes) -----------
    const long nw = ((long)sequence_length + 31) / 32;
    vector<uint32_t> hi((size_t)nw + 1), lo((size_t)nw + 1), nn((size_t)nw + 1);
    if (rb_get_planes(ctx, 0, hi.data(), lo.data(), nn.data()) != RB_OK) die("rb_get_planes", ctx);
    Bitset left_bset = to_bitset(hi.data(), sequence_length), right_bset = to_bitset(lo.data(), sequence_length),
           N_bset = to_bitset(nn.data(), sequence_length);
    // one-hot planes + per-base pointers (fasta_utils.cpp:83-115): only mostFrequentLongerMotif reads them, which runs on
    // the GPU from the packed planes (K7): MATRIX stays empty
    vector<Bitset *> MATRIX;

    // ---- the three order-dependent merges (addSeedToSeedPositions{Perfect,Substitutions,Anchored} + mergeAllLists),
    // restated in seed_merge.cpp: they consume the streams as they are and take the few plane popcounts their tie-breakers
    // need from the packed planes. Meanwhile worker threads build what only the per-seed stage reads: the anchored planes
    // B_m as the reference's bitsets (fasta_utils.cpp:117-122, 143-161), from the anchor planes the GPU computed.
    vector<uint32_t> anchors((size_t)nw * NSHIFTS + 1);
    if (rb_get_anchor_planes(ctx, 0, MINIMUM_SHIFT, MAXIMUM_SHIFT, anchors.data()) != RB_OK) die("rb_get_anchor_planes", ctx);
    vector<Bitset> lshift_xor_bsets((size_t)NSHIFTS);
    std::thread plane_builder([&] {
        vector<Bitset> lsxor_anchor_bsets((size_t)NSHIFTS);
        parallel_for(NSHIFTS, [&](int k) {
            const int i = MINIMUM_SHIFT + k;
            lshift_xor_bsets[(size_t)k] = ~(left_bset ^ (left_bset << (i))) & ~(right_bset ^ (right_bset << (i)));
            lsxor_anchor_bsets[(size_t)k] = to_bitset(anchors.data() + (size_t)k * nw, sequence_length);
        });
        // B_m = X_m | A_i for the shifts i within two of m (from 1 when m <= 2); plane m reads its own match plane and
        // anchor planes only, so the planes can be overwritten in any order
        parallel_for(NMOTIFS, [&](int k) {
            const int m = MINIMUM_MLEN + k;
            Bitset bm = lshift_xor_bsets[m - MINIMUM_SHIFT];
            for (int i = (m > 2) ? m - 2 : 1; i <= m + 2; ++i)
                if (i != m) bm |= lsxor_anchor_bsets[i - MINIMUM_SHIFT];
            lshift_xor_bsets[m - MINIMUM_SHIFT] = bm;
        });
    });

    SeedList seed_positions_perfect, seed_positions_substut, seed_positions_anchored;
    {
        static_assert(sizeof(rb_rec) == sizeof(rbm::Cand), "rb_rec is read as {start, end, mlen | flags << 16, time}");
        vector<rbm::Cand> cands[3];
        const rbm::Cand *cp[3];
        int64_t cn[3];
        for (int s = 0; s < 3; ++s) {
            cands[s].resize((size_t)st.n[s]);
            for (int64_t k = 0; k < st.n[s]; ++k) {
                const rb_rec &r = st.rec[s][k];
                cands[s][(size_t)k] = rbm::Cand{r.start, r.end, r.mlen, r.flags};
            }
            cp[s] = cands[s].data(); cn[s] = st.n[s];
        }
        rbm::AnchorArrayCounts counts(hi.data(), lo.data(), nn.data(), sequence_length, MINIMUM_SHIFT, MAXIMUM_SHIFT, anchors.data());
        rbm::SeedList lists[3];
        rbm::run_merges(cp, cn, counts, MINIMUM_MLEN, MAXIMUM_MLEN, sequence_length, lists[0], lists[1], lists[2]);
        SeedList *dst[3] = {&seed_positions_perfect, &seed_positions_substut, &seed_positions_anchored};
        for (int s = 0; s < 3; ++s) {
            dst[s]->reserve(lists[s].size());
            for (const rbm::Seed &x : lists[s]) dst[s]->emplace_back(x.start, x.end, x.mlen, x.rank);
        }
    }
    clock.mark("merges (restated, seed_merge.cpp)");
    {
        // the reference's three progress lines (fasta_utils.cpp:134, 139, 170); the first two counts are taken after all
        // passes here, so entries the anchored pass retired are already subtracted
        int failed = 0;
        for (auto &s : seed_positions_perfect) failed += get<3>(s) == -1;
        for (auto &s : seed_positions_substut) failed += get<3>(s) == -1;
        cerr << "Total number of perfect seeds: " << seed_positions_perfect.size() << "\t Time elapsed: " << difftime(time(0), START_TIME) << "secs\n";
        cerr << "Total number of seeds considering substitutions: " << seed_positions_perfect.size() + seed_positions_substut.size() - failed << "\t Time elapsed: " << difftime(time(0), START_TIME) << "secs\n";
        cerr << "Generated anchored shift XORs!\t Time elapsed: " << difftime(time(0), START_TIME) << "secs\n";
        for (auto &s : seed_positions_anchored) failed += get<3>(s) == -1;
        cerr << "Total number of seeds considering indels: " << seed_positions_perfect.size() + seed_positions_substut.size() + seed_positions_anchored.size() - failed << "\t Time elapsed: " << difftime(time(0), START_TIME) << "secs\n";
    }
    plane_builder.join();
    clock.mark("anchored planes as bitsets (overlapped with the merges)");
    if (const char *cp2 = getenv("RB_CP2_OUT")) {  // checkpoint CP2 for the parity tests: the three lists, same format as oracle/cp_hooks.h
        if (FILE *f = fopen(cp2, "ab")) {
            int32_t h[5] = {0, -1, sequence_length, 0, 0};
            fwrite(h, sizeof h, 1, f);
            int tag = 11;
            for (SeedList *l : {&seed_positions_perfect, &seed_positions_substut, &seed_positions_anchored}) {
                for (auto &s : *l) { int32_t r[5] = {tag, get<0>(s), get<1>(s), get<2>(s), get<3>(s)}; fwrite(r, sizeof r, 1, f); }
                ++tag;
            }
            fclose(f);
        }
        if (getenv("RB_CP_STOP_AFTER_CP2")) return;
    }

    // ---- per-seed stage (what fasta_utils.cpp:174-246 does): the three lists are walked head to head, the head with
    // the smallest start goes first (ties: perfect, then substitution, then anchored); entries re-tagged -1 by the
    // merges are skipped; seeds shorter than 0.9 motif lengths are not processed -----------------------------------
    clock.mark("anchored merge");
    StripedSmithWaterman::Aligner aligner;
    StripedSmithWaterman::Filter filter;
    StripedSmithWaterman::Alignment alignment;
    const SeedList *lists[3] = {&seed_positions_perfect, &seed_positions_substut, &seed_positions_anchored};
    size_t head[3] = {0, 0, 0};
    vector<rb_seed> todo;     // in processing order
    vector<int> todo_rank;
    for (;;) {
        int pick = -1;
        for (int k = 0; k < 3; ++k)
            if (head[k] < lists[k]->size() && (pick < 0 || get<0>((*lists[k])[head[k]]) < get<0>((*lists[pick])[head[pick]]))) pick = k;
        if (pick < 0) break;
        const tuple<int, int, int, int> seed = (*lists[pick])[head[pick]++];
        const int rank = get<3>(seed), mlen = get<2>(seed);
        if (rank == -1) continue;
        const int from = get<0>(seed), to = get<1>(seed);
        if (!(to - from >= 0.9 * mlen)) continue;
        todo.push_back(rb_seed{0, from, to, mlen});
        todo_rank.push_back(rank);
    }
    const int processed_seeds = (int)todo.size();
    // K5: the gate both per-seed functions apply first (parse_seed.cpp:344-367, parse_smallmotif_seed.cpp:216-235) —
    // longest run of 1s of the anchored plane over the seed below `continuous_ones_threshold` -> the reference returns
    // before doing anything — evaluated for all seeds in one batch on the GPU; such seeds are not handed over at all.
    vector<rb_seedinfo> info(todo.size());
    const bool use_filter = !getenv("RIBBIT_NO_SEED_FILTER");
    if (use_filter && !todo.empty() && rb_filter_seeds(ctx, todo.data(), (int64_t)todo.size(), info.data()) != RB_OK) die("rb_filter_seeds", ctx);
    // K7: processSeed calls mostFrequentLongerMotif(seed_start, N-truncated seed length, mlen) for every seed with mlen > 10
    // that passes the gate (parse_seed.cpp:344-390) — all of them in one batch
    g_motif.rows.clear();
    if (use_filter) {
        vector<rb_seed> mseeds;
        for (size_t k = 0; k < todo.size(); ++k)
            if (todo[k].mlen > 10 && info[k].longest_run >= continuous_ones_threshold)
                mseeds.push_back(rb_seed{0, todo[k].start, todo[k].start + info[k].seq_len, todo[k].mlen});
        vector<rb_motifrow> mrows(mseeds.size());
        if (!mseeds.empty() && rb_motif_rows(ctx, mseeds.data(), (int64_t)mseeds.size(), mrows.data()) != RB_OK) die("rb_motif_rows", ctx);
        g_motif.rows.reserve(mseeds.size() * 2);
        for (size_t k = 0; k < mseeds.size(); ++k)
            g_motif.rows.emplace(MotifRows::key(mseeds[k].start, mseeds[k].mlen), std::make_pair(mseeds[k].end - mseeds[k].start, mrows[k].row));
    }
    vector<size_t> live;  // seeds that are handed to the per-seed functions, in processing order
    for (size_t k = 0; k < todo.size(); ++k)
        if (!(use_filter && info[k].longest_run < continuous_ones_threshold)) live.push_back(k);
    auto run_seeds = [&](size_t i0, size_t i1, ostream &o) {
        for (size_t i = i0; i < i1; ++i) {
            const size_t k = live[i];
            int mlen = todo[k].mlen, rank = todo_rank[k];
            const int from = todo[k].start, to = todo[k].end;
            Bitset &plane = lshift_xor_bsets[mlen - MINIMUM_SHIFT];
            if (mlen <= 10)
                processSeedMotifWise(tuple<int, int>{from, to}, mlen, rank, sequence_id, sequence, sequence_length, plane, left_bset,
                                     right_bset, N_bset, continuous_ones_threshold, o, aligner, filter, alignment);
            else
                processSeed(tuple<int, int>{from, to}, mlen, rank, sequence_id, sequence, sequence_length, plane, left_bset, right_bset,
                            N_bset, continuous_ones_threshold, o, MATRIX, aligner, filter, alignment);
        }
    };
    clock.mark("seed gate (K5) + motif rows (K7)");
    const int procs = host_procs();
    long served = 0;
    bool in_workers = false;
    if (procs > 1 && live.size() >= 4096) {
        // parts of roughly equal work (seed length + motif size as the weight), several per worker
        const int n_parts = (int)std::min<size_t>((size_t)procs * 8, live.size() / 256);
        vector<double> acc(live.size() + 1, 0.0);
        for (size_t i = 0; i < live.size(); ++i) acc[i + 1] = acc[i] + 64.0 + (todo[live[i]].end - todo[live[i]].start) + todo[live[i]].mlen;
        vector<size_t> cut((size_t)n_parts + 1, live.size());
        cut[0] = 0;
        for (int p = 1; p < n_parts; ++p)
            cut[(size_t)p] = (size_t)(std::lower_bound(acc.begin(), acc.end(), acc.back() * p / n_parts) - acc.begin());
        for (int p = 1; p <= n_parts; ++p) cut[(size_t)p] = std::max(cut[(size_t)p], cut[(size_t)p - 1]);
        const bool done = forked_parts(n_parts, procs, out,
                     [&](int part, ostream &o) { run_seeds(cut[(size_t)part], cut[(size_t)part + 1], o); },
                     [&](int fd) {
                         rb_seed sd;
                         if (!read_all(fd, &sd, sizeof sd)) return false;
                         rb_motifrow r;
                         if (rb_motif_rows(ctx, &sd, 1, &r) != RB_OK) die("rb_motif_rows", ctx);
                         ++served;
                         return write_all(fd, &r, sizeof r);
                     });
        in_workers = done;
        if (!done) run_seeds(0, live.size(), out);
    } else
        run_seeds(0, live.size(), out);
    clock.mark("per-seed stage");
    if (getenv("RIBBIT_VERBOSE"))
        cerr << "K7 motif rows: " << g_motif.rows.size() << " in the batch, " << g_motif.misses + served << " single calls; per-seed stage on "
             << (in_workers ? procs : 1) << " process(es)\n";
    cerr << "Total number of seeds that are processed for alignment: " << processed_seeds << "\t Time elapsed: " << difftime(time(0), START_TIME) << "secs\n";
}
