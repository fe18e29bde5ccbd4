// hostpack.cpp — see hostpack.hpp.  Compiled by g++ (not nvcc: the AVX-512 intrinsics headers) and linked into libpgr_b200.so.
#include "hostpack.hpp"

#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace pgr {
namespace {

// one block or less, byte by byte (the reference's LUT, shmmrutils.rs:426-436)
inline void pack_block_scalar(const uint8_t *s, size_t n, uint32_t &p0, uint32_t &p1, uint32_t &v) {
    uint32_t a = 0, b = 0, c = 0;
    for (size_t j = 0; j < n; j++) {
        const uint8_t ch = s[j];
        uint32_t code;
        if (ch < 4) code = ch;
        else switch (ch) {
            case 'A': case 'a': code = 0; break;
            case 'C': case 'c': code = 1; break;
            case 'G': case 'g': code = 2; break;
            case 'T': case 't': code = 3; break;
            default: code = 4;
        }
        if (code < 4) { a |= (code & 1u) << j; b |= (code >> 1) << j; c |= 1u << j; }
    }
    if (n < 32) c |= 0xFFFFFFFFu << n;   // padding behind the end of a sequence counts as valid (code 0): its bytes are never read as sequence
    p0 = a; p1 = b; v = c;
}

// every packer returns the AND of the validity words it wrote (all ones = every byte a base)
uint32_t pack_scalar(const uint8_t *src, size_t n_bytes, uint32_t *p0, uint32_t *p1, uint32_t *v) {
    size_t b = 0;
    uint32_t all = 0xFFFFFFFFu;
    for (size_t i = 0; i < n_bytes; i += 32, b++) { pack_block_scalar(src + i, std::min<size_t>(32, n_bytes - i), p0[b], p1[b], v[b]); all &= v[b]; }
    return all;
}

// code bit 0 = bit 1 of (c ^ (c >> 1)), code bit 1 = bit 2 of c for the eight letters; a letter is recognised by looking the
// expected upper-case byte up by its low nibble (non-letters map to 0x20, which (c & 0xDF) never equals)
__attribute__((target("avx2"))) uint32_t pack_avx2(const uint8_t *src, size_t n_bytes, uint32_t *p0, uint32_t *p1, uint32_t *v) {
    const __m256i lut = _mm256_setr_epi8(0x20, 0x41, 0x20, 0x43, 0x54, 0x20, 0x20, 0x47, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20,
                                         0x20, 0x41, 0x20, 0x43, 0x54, 0x20, 0x20, 0x47, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20);
    const __m256i m_case = _mm256_set1_epi8((char)0xDF), m_nib = _mm256_set1_epi8(0x0F), m_fc = _mm256_set1_epi8((char)0xFC), zero = _mm256_setzero_si256();
    const size_t nb = n_bytes / 32;
    uint32_t all = 0xFFFFFFFFu;
    for (size_t b = 0; b < nb; b++) {
        const __m256i c = _mm256_loadu_si256((const __m256i *)(src + 32 * b));
        const __m256i t = _mm256_xor_si256(c, _mm256_srli_epi16(c, 1));
        uint32_t a = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(t, 6));
        uint32_t d = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(c, 5));
        const __m256i u = _mm256_and_si256(c, m_case);
        uint32_t ok = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, _mm256_and_si256(u, m_nib)), u));
        const uint32_t raw = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(c, m_fc), zero));
        if (raw) {   // bytes 0..3 are bases for the reference: code = the byte
            a = (a & ~raw) | ((uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(c, 7)) & raw);
            d = (d & ~raw) | ((uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(c, 6)) & raw);
            ok |= raw;
        }
        p0[b] = a & ok; p1[b] = d & ok; v[b] = ok;
        all &= ok;
    }
    if (n_bytes % 32) { pack_block_scalar(src + 32 * nb, n_bytes % 32, p0[nb], p1[nb], v[nb]); all &= v[nb]; }
    return all;
}

__attribute__((target("avx512f,avx512bw"))) uint32_t pack_avx512(const uint8_t *src, size_t n_bytes, uint32_t *p0, uint32_t *p1, uint32_t *v) {
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(0x20, 0x41, 0x20, 0x43, 0x54, 0x20, 0x20, 0x47, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20));
    const __m512i m_case = _mm512_set1_epi8((char)0xDF), m_nib = _mm512_set1_epi8(0x0F), m_fc = _mm512_set1_epi8((char)0xFC);
    const __m512i b1 = _mm512_set1_epi8(1), b2 = _mm512_set1_epi8(2), b4 = _mm512_set1_epi8(4);
    const size_t n64 = n_bytes / 64;
    uint64_t all = ~0ull;
    for (size_t q = 0; q < n64; q++) {
        _mm_prefetch((const char *)(src + 64 * q + 2048), _MM_HINT_NTA);
        const __m512i c = _mm512_loadu_si512((const void *)(src + 64 * q));
        const __m512i t = _mm512_xor_si512(c, _mm512_srli_epi16(c, 1));
        uint64_t a = _mm512_test_epi8_mask(t, b2);
        uint64_t d = _mm512_test_epi8_mask(c, b4);
        const __m512i u = _mm512_and_si512(c, m_case);
        uint64_t ok = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lut, _mm512_and_si512(u, m_nib)), u);
        const uint64_t raw = _mm512_testn_epi8_mask(c, m_fc);
        if (raw) {
            a = (a & ~raw) | (_mm512_test_epi8_mask(c, b1) & raw);
            d = (d & ~raw) | (_mm512_test_epi8_mask(c, b2) & raw);
            ok |= raw;
        }
        a &= ok; d &= ok;
        p0[2 * q] = (uint32_t)a; p0[2 * q + 1] = (uint32_t)(a >> 32);
        p1[2 * q] = (uint32_t)d; p1[2 * q + 1] = (uint32_t)(d >> 32);
        v[2 * q] = (uint32_t)ok; v[2 * q + 1] = (uint32_t)(ok >> 32);
        all &= ok;
    }
    uint32_t all32 = (uint32_t)all & (uint32_t)(all >> 32);
    const size_t done = 64 * n64;
    if (done < n_bytes) all32 &= pack_avx2(src + done, n_bytes - done, p0 + 2 * n64, p1 + 2 * n64, v + 2 * n64);
    return all32;
}

using PackFn = uint32_t (*)(const uint8_t *, size_t, uint32_t *, uint32_t *, uint32_t *);
struct Isa { PackFn fn; const char *name; };
Isa pick_isa() {
    const char *force = getenv("PGR_B200_PACK_ISA");   // test aid: "scalar" / "avx2"
    __builtin_cpu_init();
    const bool a512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f"), a2 = __builtin_cpu_supports("avx2");
    if (force && std::string(force) == "scalar") return {pack_scalar, "scalar"};
    if (force && std::string(force) == "avx2" && a2) return {pack_avx2, "avx2"};
    if (a512) return {pack_avx512, "avx512bw"};
    if (a2) return {pack_avx2, "avx2"};
    return {pack_scalar, "scalar"};
}
const Isa &isa() { static const Isa i = pick_isa(); return i; }

// ---- worker pool ---------------------------------------------------------------------------------------------------
struct Pool {
    std::mutex job_mu;                 // one parallel_for at a time
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    const std::function<void(size_t)> *fn = nullptr;
    size_t n = 0;
    std::atomic<size_t> next{0};
    uint64_t generation = 0;
    unsigned pending = 0;
    bool stop = false;
    unsigned n_threads = 1;

    Pool() {
        unsigned t = 0;
        if (const char *e = getenv("PGR_B200_HOST_THREADS")) t = (unsigned)atoi(e);
        if (!t) {
            cpu_set_t set;
            if (sched_getaffinity(0, sizeof set, &set) == 0) t = (unsigned)CPU_COUNT(&set);
            if (!t) t = std::thread::hardware_concurrency();
            // one process per GPU (torchrun): the processes of a node share its CPUs, so each takes its share; an oversubscribed
            // pool turns the per-slot joins of the packer into waits for descheduled threads
            if (const char *lw = getenv("LOCAL_WORLD_SIZE")) { const unsigned w = (unsigned)atoi(lw); if (w > 1) t = std::max(1u, t / w); }
            t = std::min(t, 32u);
        }
        n_threads = std::max(1u, t);
        for (unsigned i = 1; i < n_threads; i++) workers.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_work.notify_all();
        for (auto &w : workers) w.join();
    }
    void drain() {
        for (;;) {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n) break;
            (*fn)(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            drain();
            bool last;
            {
                std::lock_guard<std::mutex> lk(mu);
                last = --pending == 0;
            }
            if (last) cv_done.notify_one();
        }
    }
    // every worker acknowledges every job, so fn and n are never read after run() has returned
    void run(size_t count, const std::function<void(size_t)> &f) {
        if (count == 0) return;
        if (count == 1 || n_threads == 1) { for (size_t i = 0; i < count; i++) f(i); return; }
        std::lock_guard<std::mutex> job(job_mu);
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = &f; n = count; next.store(0); pending = (unsigned)workers.size(); generation++;
        }
        cv_work.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};
Pool &pool() { static Pool p; return p; }

}  // namespace

uint32_t pack_bases(const uint8_t *src, size_t n_bytes, uint32_t *p0, uint32_t *p1, uint32_t *v) { return isa().fn(src, n_bytes, p0, p1, v); }
const char *pack_isa() { return isa().name; }
void parallel_for(size_t n, const std::function<void(size_t)> &fn) { pool().run(n, fn); }
unsigned pool_threads() { return pool().n_threads; }

}  // namespace pgr
