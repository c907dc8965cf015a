// TEST INFRASTRUCTURE — not part of the product path.
//
// Checkpoint logging force-included (-include) into the instrumented oracle
// build (oracle/instrument.sh). The reference call sites are redirected to the
// RB_CP* macros by sed in a temp copy; the macros log and then forward to the
// untouched reference function, so behaviour is unchanged.
//
// Output (binary, little-endian int32 x5 per record) goes to $RB_CP_OUT:
//   {0, contig_index, L, 0, 0}                 contig header
//   {1|2|3, start, end, mlen, 0}               CP1: top-level argument sequence of
//                                              addSeedToSeedPositions{Perfect,Substitutions,Anchored}
//   {11|12|13, start, end, mlen, rank}         CP2: the three lists after all passes
//   {21, seed_start, seed_seq_len, mlen, row}  CP4: mostFrequentLongerMotif arguments and the row it chose
#ifndef RB_CP_HOOKS_H
#define RB_CP_HOOKS_H
#ifdef __cplusplus
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <tuple>
#include <vector>
#include <unistd.h>

inline FILE *&rb_cp_file() { static FILE *f = nullptr; return f; }
inline int &rb_cp_contig() { static int c = -1; return c; }
inline void rb_cp_segv(int sig) {
    if (rb_cp_file()) fflush(rb_cp_file());
    _exit(128 + sig);
}
inline FILE *rb_cp_open() {
    FILE *&f = rb_cp_file();
    static bool off = false;
    if (off) return nullptr;
    if (!f) {
        const char *p = getenv("RB_CP_OUT");
        if (!p) { off = true; return nullptr; }  /* timing runs: no logging at all */
        f = fopen(p, "wb");
        if (!f) { perror("RB_CP_OUT"); exit(2); }
        static char buf[1 << 20];
        setvbuf(f, buf, _IOFBF, sizeof buf);
        signal(SIGSEGV, rb_cp_segv);
        signal(SIGABRT, rb_cp_segv);
        atexit([] { if (rb_cp_file()) fflush(rb_cp_file()); });
    }
    return f;
}
inline void rb_cp_rec(int32_t tag, int32_t a, int32_t b, int32_t c, int32_t d) {
    FILE *f = rb_cp_open();
    if (!f) return;
    int32_t r[5] = {tag, a, b, c, d};
    fwrite(r, sizeof r, 1, f);
}
inline void rb_cp_contig_start(long len) { rb_cp_rec(0, ++rb_cp_contig(), (int32_t)len, 0, 0); }
inline void rb_cp2_dump(const std::vector<std::tuple<int, int, int, int>> &p,
                        const std::vector<std::tuple<int, int, int, int>> &s,
                        const std::vector<std::tuple<int, int, int, int>> &a) {
    for (auto &t : p) rb_cp_rec(11, std::get<0>(t), std::get<1>(t), std::get<2>(t), std::get<3>(t));
    for (auto &t : s) rb_cp_rec(12, std::get<0>(t), std::get<1>(t), std::get<2>(t), std::get<3>(t));
    for (auto &t : a) rb_cp_rec(13, std::get<0>(t), std::get<1>(t), std::get<2>(t), std::get<3>(t));
    if (rb_cp_open()) fflush(rb_cp_open());
}

#define RB_CP1_P(s, e, m, ...) (rb_cp_rec(1, (s), (e), (m), 0), addSeedToSeedPositionsPerfect((s), (e), (m), __VA_ARGS__))
#define RB_CP1_S(s, e, m, ...) (rb_cp_rec(2, (s), (e), (m), 0), addSeedToSeedPositionsSubstitutions((s), (e), (m), __VA_ARGS__))
#define RB_CP1_A(s, e, m, ...) (rb_cp_rec(3, (s), (e), (m), 0), addSeedToSeedPositionsAnchored((s), (e), (m), __VA_ARGS__))
#endif
#endif
