// CPU self-test of the library's host pool (pgr_tk_b200/csrc/hostpack.cpp): every index of a parallel_for is visited exactly once,
// also when several threads call it at the same time (the shards of one process do), for many back-to-back jobs of odd sizes;
// and pack_bases gives the same planes when the work is cut into pieces by the pool as in one piece.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../pgr_tk_b200/csrc/hostpack.hpp"

int main() {
    using namespace pgr;
    const unsigned T = pool_threads();
    // 1. exactly-once, repeated, odd sizes
    for (int rep = 0; rep < 300; rep++) {
        const size_t n = (size_t)(rep * 37 % 1000) + (rep % 3 == 0 ? 0 : 1);
        std::vector<std::atomic<int>> seen(n + 1);
        for (auto &s : seen) s.store(0);
        parallel_for(n, [&](size_t i) { seen[i].fetch_add(1); });
        for (size_t i = 0; i < n; i++) if (seen[i].load() != 1) { printf("FAIL rep %d index %zu seen %d\n", rep, i, seen[i].load()); return 1; }
        if (seen[n].load() != 0) { printf("FAIL: index past the end visited\n"); return 1; }
    }
    // 2. concurrent callers
    std::atomic<long> total{0};
    std::vector<std::thread> callers;
    for (int c = 0; c < 4; c++)
        callers.emplace_back([&, c] {
            for (int rep = 0; rep < 100; rep++) {
                std::atomic<long> local{0};
                const size_t n = 100 + (size_t)c * 13 + (size_t)rep;
                parallel_for(n, [&](size_t i) { local.fetch_add((long)i + 1); });
                if (local.load() != (long)(n * (n + 1) / 2)) { printf("FAIL caller %d rep %d\n", c, rep); _Exit(1); }
                total.fetch_add(local.load());
            }
        });
    for (auto &t : callers) t.join();
    // 3. packing in pool pieces == packing in one piece
    const size_t L = (1u << 22) + 77;
    std::vector<uint8_t> seq(L);
    uint32_t x = 12345;
    for (size_t i = 0; i < L; i++) { x = x * 1664525u + 1013904223u; seq[i] = (x >> 24) < 250 ? "ACGTacgt"[(x >> 8) & 7] : (uint8_t)(x >> 16); }
    const size_t nb = (L + 31) / 32;
    std::vector<uint32_t> a(3 * nb), b(3 * nb, 0xDEADBEEFu);
    const uint32_t all1 = pack_bases(seq.data(), L, a.data(), a.data() + nb, a.data() + 2 * nb);
    const size_t piece = 4096;   // blocks
    std::atomic<uint32_t> all2{0xFFFFFFFFu};
    parallel_for((nb + piece - 1) / piece, [&](size_t pc) {
        const size_t b0 = pc * piece, b1 = std::min(nb, b0 + piece);
        const size_t bytes = std::min(L, b1 * 32) - b0 * 32;
        all2.fetch_and(pack_bases(seq.data() + b0 * 32, bytes, b.data() + b0, b.data() + nb + b0, b.data() + 2 * nb + b0));
    });
    if (memcmp(a.data(), b.data(), a.size() * 4) != 0 || all1 != all2.load()) { printf("FAIL: pieces differ from one piece\n"); return 1; }
    printf("ok threads=%u isa=%s total=%ld all_valid=%08x\n", T, pack_isa(), total.load(), all1);
    return 0;
}
