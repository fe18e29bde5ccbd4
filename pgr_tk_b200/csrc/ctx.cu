// ctx.cu — device context, sequence store and the sequence_to_shmmrs pipeline (host orchestration).
#include "ctx.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "pack_upload.cuh"
#include "patch_kernels.cuh"
#include "sketch_kernels.cuh"

namespace pgr {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

void trace_mark(const char *what) {
    static const bool on = getenv("PGR_B200_TRACE") != nullptr;
    if (!on) return;
    static thread_local double last = 0;
    cudaDeviceSynchronize();
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    fprintf(stderr, "[pgr_b200 trace] %-40s +%.3f ms\n", what, last == 0 ? 0.0 : now - last);
    last = now;
}

int StageTimer::begin(const char *name, cudaStream_t st) {
    if (used == ev0.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        ev0.push_back(a); ev1.push_back(b); names.push_back(name); ms.push_back(0.f);
    }
    names[used] = name;
    cudaEventRecord(ev0[used], st);
    return (int)used++;
}
void StageTimer::end(int slot, cudaStream_t st) { cudaEventRecord(ev1[slot], st); }
void StageTimer::collect() {
    for (size_t i = 0; i < used; i++) {
        cudaEventSynchronize(ev1[i]);
        cudaEventElapsedTime(&ms[i], ev0[i], ev1[i]);
    }
    names_z.assign(names.begin(), names.begin() + used);
    names_z.push_back(nullptr);
}
void StageTimer::destroy() {
    for (auto e : ev0) cudaEventDestroy(e);
    for (auto e : ev1) cudaEventDestroy(e);
    ev0.clear(); ev1.clear();
}

// Result buffers handed to the caller.  Large ones come from a small pool of pinned host buffers so that the D2H copy
// runs at PCIe rate and no page faults are taken on every call; pgr_b200_free() returns them to the pool.
namespace {
struct PoolEntry { void *p; size_t cap; bool in_use; };
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool;
constexpr size_t POOL_MIN = 1u << 20;     // smaller results use malloc
constexpr size_t POOL_MAX_FREE = 16;      // idle pinned buffers kept around
}  // namespace

void *result_alloc(size_t bytes) {
    if (bytes < POOL_MIN) return malloc(bytes ? bytes : 1);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (size_t i = 0; i < g_pool.size(); i++)
        if (!g_pool[i].in_use && g_pool[i].cap >= bytes && (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = (int)i;
    if (best >= 0) { g_pool[best].in_use = true; return g_pool[best].p; }
    void *p = nullptr;
    const size_t cap = bytes + bytes / 8;
    if (cudaMallocHost(&p, cap) != cudaSuccess) { cudaGetLastError(); return malloc(bytes); }
    g_pool.push_back({p, cap, true});
    return p;
}

void result_free(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); i++) {
            if (g_pool[i].p != p) continue;
            g_pool[i].in_use = false;
            size_t idle = 0;
            for (auto &e : g_pool) idle += e.in_use ? 0 : 1;
            if (idle > POOL_MAX_FREE) { cudaFreeHost(p); g_pool.erase(g_pool.begin() + i); }
            return;
        }
    }
    free(p);
}

int check_spec(const pgr_shmmr_spec *s) {
    if (!s) { set_error("spec is NULL"); return PGR_E_ARG; }
    if (s->k == 0 || s->k > 56) { set_error("assert!(k <= 56) violated (k=%u)", s->k); return PGR_E_SPEC; }
    if (!(s->r > 0 && s->r < 13)) { set_error("assert!(r > 0 && r < 13) violated (r=%u)", s->r); return PGR_E_SPEC; }
    if (!s->sketch && (s->w == 0 || s->w > 128)) { set_error("assert!(w <= 128) violated (w=%u)", s->w); return PGR_E_SPEC; }
    return PGR_OK;
}

}  // namespace pgr

using namespace pgr;

// Page-locked scratch of a context (control read-backs, staging of small sequences) comes from a process-wide free list:
// cudaMallocHost / cudaFreeHost synchronise the device and serialise in the driver, and callers that make a fresh index (and
// with it a fresh context) per build — bench_index.py, the CLIs' shards — would pay them on every build.
namespace {
struct PinnedItem { void *p; size_t cap; };
std::mutex g_pinned_mu;
std::vector<PinnedItem> g_pinned_free;
void *pinned_take(size_t bytes, size_t *cap) {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        int best = -1;
        for (size_t i = 0; i < g_pinned_free.size(); i++)
            if (g_pinned_free[i].cap >= bytes && (best < 0 || g_pinned_free[i].cap < g_pinned_free[best].cap)) best = (int)i;
        if (best >= 0) { PinnedItem it = g_pinned_free[best]; g_pinned_free.erase(g_pinned_free.begin() + best); *cap = it.cap; return it.p; }
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *cap = bytes;
    return p;
}
void pinned_give(void *p, size_t cap) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    if (g_pinned_free.size() >= 32) { cudaFreeHost(p); return; }
    g_pinned_free.push_back({p, cap});
}
}  // namespace

int pgr_b200_ctx::ensure_stage(size_t bytes) {
    if (bytes <= h_stage_cap) return PGR_OK;
    pinned_give(h_stage, h_stage_cap);
    h_stage = nullptr; h_stage_cap = 0;
    if (!(h_stage = pinned_take(bytes, &h_stage_cap))) { set_error("cudaMallocHost(%zu) failed", bytes); return PGR_E_CUDA; }
    return PGR_OK;
}
int pgr_b200_ctx::ensure_ctl(size_t bytes) {
    if (bytes <= h_ctl_cap) return PGR_OK;
    pinned_give(h_ctl, h_ctl_cap);
    h_ctl = nullptr; h_ctl_cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    if (!(h_ctl = pinned_take(want, &h_ctl_cap))) { set_error("cudaMallocHost(%zu) failed", want); return PGR_E_CUDA; }
    return PGR_OK;
}


extern "C" {

int pgr_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char *pgr_b200_last_error(void) { return get_error(); }
void pgr_b200_free(void *p) { pgr::result_free(p); }
void *pgr_b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { set_error("cudaMallocHost(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
    return p;
}
void pgr_b200_host_free(void *p) { if (p) cudaFreeHost(p); }
int pgr_b200_host_register(void *p, size_t bytes) {
    if (!p || !bytes) { set_error("NULL argument"); return PGR_E_ARG; }
    if (pgr_b200_device_count() <= 0) { set_error("no CUDA device available"); return PGR_E_NO_DEVICE; }
    PGR_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return PGR_OK;
}
uint32_t pgr_b200_pack_bases(const uint8_t *src, size_t n_bytes, uint32_t *p0, uint32_t *p1, uint32_t *v) { return pgr::pack_bases(src, n_bytes, p0, p1, v); }
const char *pgr_b200_pack_isa(void) { return pgr::pack_isa(); }
int pgr_b200_pool_threads(void) { return (int)pgr::pool_threads(); }
int pgr_b200_last_transport(void) { return pgr::last_transport().load(); }
uint64_t pgr_b200_transport_bytes(void) { return pgr::transport_bytes().load(); }
int pgr_b200_set_transport(int mode) { return pgr::transport_mode().exchange(mode == PGR_TRANSPORT_DIRECT ? 1 : 0); }
int pgr_b200_host_unregister(void *p) {
    if (!p) return PGR_OK;
    PGR_CUDA(cudaHostUnregister(p));
    return PGR_OK;
}

// Contexts are recycled: a freed context keeps its streams, its grow-only device buffers and its page-locked scratch and waits
// in a short per-process list for the next pgr_b200_ctx_new / pgr_b200_index_new on the same device.  Callers that make an index
// per build (bench_index.py, the shards of the CLIs) then neither create streams nor size the pipeline buffers again; making a
// context from scratch was measured at 12-13 ms on a 4-GPU run against 9 ms for the whole HBM-resident build of config 3.
namespace {
std::mutex g_ctx_mu;
std::vector<pgr_b200_ctx *> g_ctx_free;
constexpr size_t CTX_KEEP = 4;
}  // namespace

pgr_b200_ctx *pgr_b200_ctx_new(int device) {
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        for (size_t i = 0; i < g_ctx_free.size(); i++) {
            if (g_ctx_free[i]->device != device) continue;
            pgr_b200_ctx *c = g_ctx_free[i];
            g_ctx_free.erase(g_ctx_free.begin() + i);
            if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); g_ctx_free.push_back(c); return nullptr; }
            return c;
        }
    }
    int n = pgr_b200_device_count();
    if (n <= 0) { set_error("no CUDA device available: libpgr_b200 has no CPU fallback"); return nullptr; }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return nullptr; }
    int cc_major = 0, n_sm = 0;
    if (cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { set_error("cudaDeviceGetAttribute failed"); return nullptr; }
    if (cc_major < 10) { set_error("device %d is sm_%dx; this library is built for sm_100a only", device, cc_major); return nullptr; }
    {   // keep freed device memory in the pool (see DevBuf)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    pgr_b200_ctx *ctx = new pgr_b200_ctx();
    ctx->device = device;
    ctx->n_sm = n_sm;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete ctx; return nullptr; }
    ctx->stream = ctx->own_stream;
    return ctx;
}

void pgr_b200_ctx_free(pgr_b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (g_ctx_free.size() < CTX_KEEP) {
            // back to a blank state: no sequences, no result, the caller's stream forgotten; buffers, bitmaps (with their lazily
            // cleared ranges), streams, timers and the pack ring stay
            ctx->stream = ctx->own_stream;
            ctx->d_seq = nullptr; ctx->n_seq = 0; ctx->r0 = ctx->rn = 0; ctx->total_bases = 0;
            ctx->h_off.clear(); ctx->h_len.clear(); ctx->h_rid.clear();
            ctx->d_result = nullptr; ctx->d_result_off = nullptr; ctx->n_result = 0; ctx->result_valid = false;
            ctx->l0_chunked = false; ctx->l0_chunk_cap = 0; ctx->l0_chunks = 0;
            memset(ctx->counters, 0, sizeof ctx->counters);
            ctx->timer.reset();
            g_ctx_free.push_back(ctx);
            return;
        }
    }
    pgr::DevBuf *bufs[] = {&ctx->seq_store, &ctx->d_off, &ctx->d_len, &ctx->d_rid, &ctx->tile_prefix, &ctx->cta_tile, &ctx->arena,
                           &ctx->chunk_count, &ctx->seq_count, &ctx->seq_flag, &ctx->replay_list, &ctx->replay_count,
                           &ctx->chunk_prefix, &ctx->seq_fast, &ctx->seq_dst, &ctx->bufA, &ctx->bufB, &ctx->flags,
                           &ctx->block_sum, &ctx->block_prefix, &ctx->block_chunk, &ctx->off_a, &ctx->off_b, &ctx->fix_mm, &ctx->fix_off, &ctx->mark_bits, &ctx->allinv_bits, &ctx->n_skips};
    for (auto b : bufs) b->release();
    for (auto &b : ctx->patch_buf) b.release();
    pinned_give(ctx->h_stage, ctx->h_stage_cap);
    pinned_give(ctx->h_ctl, ctx->h_ctl_cap);
    if (ctx->pack) pgr::pack_ring_release(ctx->pack);
    ctx->timer.destroy();
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int pgr_b200_ctx_set_stream(pgr_b200_ctx *ctx, void *cuda_stream) {
    if (!ctx) { set_error("ctx is NULL"); return PGR_E_ARG; }
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PGR_OK;
}

static int set_seq_tables(pgr_b200_ctx *ctx, size_t n, const uint32_t *rids) {
    ctx->n_seq = n;
    ctx->h_rid.resize(n);
    for (size_t i = 0; i < n; i++) ctx->h_rid[i] = rids ? rids[i] : (uint32_t)i;
    PGR_TRY(ctx->d_off.ensure((n + 1) * sizeof(uint64_t)));
    PGR_TRY(ctx->d_len.ensure((n + 1) * sizeof(uint32_t)));
    PGR_TRY(ctx->d_rid.ensure((n + 1) * sizeof(uint32_t)));
    if (n) {
        PGR_CUDA(cudaMemcpyAsync(ctx->d_off.p, ctx->h_off.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        PGR_CUDA(cudaMemcpyAsync(ctx->d_len.p, ctx->h_len.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        PGR_CUDA(cudaMemcpyAsync(ctx->d_rid.p, ctx->h_rid.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    PGR_CUDA(cudaStreamSynchronize(ctx->stream));  // host vectors may be reused by the next call
    ctx->result_valid = false;
    return PGR_OK;
}

// lay the batch out in the device sequence store (32-byte aligned starts, SEQ_SLACK bytes before and after) and
// publish the per-sequence tables; no sequence bytes are copied yet
static int upload_layout(pgr_b200_ctx *ctx, size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens) {
    ctx->h_off.resize(n);
    ctx->h_len.resize(n);
    uint64_t off = SEQ_SLACK, total = 0;
    for (size_t i = 0; i < n; i++) {
        if (lens[i] > 0x7FFFFF00ull) { set_error("sequence %zu longer than 2^31 (MM128 pos is 31 bits)", i); return PGR_E_LIMIT; }
        if (lens[i] && !seqs[i]) { set_error("sequence %zu is NULL", i); return PGR_E_ARG; }
        ctx->h_off[i] = off;
        ctx->h_len[i] = (uint32_t)lens[i];
        off += (lens[i] + 31) & ~(uint64_t)31;
        total += lens[i];
    }
    const uint64_t store_bytes = off + SEQ_SLACK;
    const bool grew = store_bytes > ctx->seq_store.cap;
    PGR_TRY(ctx->seq_store.ensure(store_bytes));
    if (grew) {
        PGR_CUDA(cudaMemsetAsync(ctx->seq_store.p, 0, ctx->seq_store.cap, ctx->stream));
        PGR_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ctx->d_seq = ctx->seq_store.as<uint8_t>();
    ctx->total_bases = total;
    return set_seq_tables(ctx, n, rids);
}

// copy sequences [i0, i1) into the store on stream `st` (asynchronous when the caller's buffers are pinned).
// Large sequences go straight from the caller's buffer; small ones are packed through a pinned staging buffer so
// that they share one copy.
static int upload_copy(pgr_b200_ctx *ctx, const uint8_t *const *seqs, const size_t *lens, size_t i0, size_t i1, cudaStream_t st) {
    uint8_t *base = ctx->seq_store.as<uint8_t>();
    const size_t BIG = 128u << 10, STAGE = 16u << 20;   // pinning is ~0.3 ms/MB: keep the staging buffers small
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int cur = 0;
    bool used[2] = {false, false};
    size_t i = i0;
    int rc = PGR_OK;
    while (i < i1 && rc == PGR_OK) {
        if (lens[i] >= BIG) {
            if (cudaMemcpyAsync(base + ctx->h_off[i], seqs[i], lens[i], cudaMemcpyHostToDevice, st) != cudaSuccess) rc = PGR_E_CUDA;
            i++;
            continue;
        }
        if (!stage[0]) {
            if ((rc = ctx->ensure_stage(2 * STAGE)) != PGR_OK) break;
            stage[0] = (uint8_t *)ctx->h_stage; stage[1] = stage[0] + STAGE;
            cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
        }
        // run of small sequences [i, j) whose device span fits the staging buffer
        const uint64_t span0 = ctx->h_off[i];
        size_t j = i;
        while (j < i1 && lens[j] < BIG && (ctx->h_off[j] - span0) + ((lens[j] + 31) & ~(size_t)31) <= STAGE) j++;
        if (used[cur]) cudaEventSynchronize(ev[cur]);
        uint64_t span = 0;
        {
            const size_t last = j - 1;
            span = (ctx->h_off[last] - span0) + ((lens[last] + 31) & ~(uint64_t)31);
            uint8_t *dst = stage[cur];
            auto fill = [&, dst](size_t q0, size_t q1) {
                for (size_t q = q0; q < q1; q++) {
                    const uint64_t rel = ctx->h_off[q] - span0;
                    if (lens[q]) memcpy(dst + rel, seqs[q], lens[q]);
                    const uint64_t padded = (lens[q] + 31) & ~(uint64_t)31;
                    if (padded > lens[q]) memset(dst + rel + lens[q], 0, padded - lens[q]);
                }
            };
            // many small sequences (a batch of queries): one host thread copies at ~10 GB/s, far below the PCIe rate the
            // staging buffer is emptied at, so a few threads share the fill
            const size_t n_run = j - i;
            const unsigned T = (span >= (4u << 20) && n_run >= 8) ? 4u : 1u;
            if (T == 1) {
                fill(i, j);
            } else {
                std::vector<std::thread> th;
                for (unsigned t = 1; t < T; t++) th.emplace_back(fill, i + n_run * t / T, i + n_run * (t + 1) / T);
                fill(i, i + n_run / T);
                for (auto &x : th) x.join();
            }
        }
        if (span && cudaMemcpyAsync(base + span0, stage[cur], span, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = PGR_E_CUDA;
        cudaEventRecord(ev[cur], st);
        used[cur] = true;
        cur ^= 1;
        i = j;
    }
    if (stage[0]) {
        // the staging buffers are reused by the next call: wait until the copies that read them are done
        if (used[0]) cudaEventSynchronize(ev[0]);
        if (used[1]) cudaEventSynchronize(ev[1]);
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
    }
    if (rc == PGR_E_CUDA) set_error("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int pgr_b200_ctx_upload(pgr_b200_ctx *ctx, size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens) {
    if (!ctx || (n && (!seqs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(ctx->device));
    PGR_TRY(upload_layout(ctx, n, rids, seqs, lens));
    if (choose_packed(seqs, lens, n, ctx->total_bases)) {
        if (!ctx->pack && !(ctx->pack = pack_ring_acquire(ctx->device))) return PGR_E_CUDA;
        PGR_TRY(upload_packed(ctx->pack, ctx->seq_store.as<uint8_t>(), ctx->h_off, ctx->h_len, seqs, 0, n, ctx->stream));
    } else {
        PGR_TRY(upload_copy(ctx, seqs, lens, 0, n, ctx->stream));
    }
    PGR_CUDA(cudaStreamSynchronize(ctx->stream));
    return PGR_OK;
}

int pgr_b200_ctx_set_device_seqs(pgr_b200_ctx *ctx, const uint8_t *dev_base, size_t n, const uint32_t *rids, const uint64_t *offs,
                                 const uint64_t *lens) {
    if (!ctx || (n && (!dev_base || !offs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(ctx->device));
    if (((uintptr_t)dev_base) & 31) { set_error("device base pointer must be 32-byte aligned"); return PGR_E_ARG; }
    ctx->h_off.resize(n);
    ctx->h_len.resize(n);
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        if (offs[i] & 31) { set_error("sequence offset %zu is not 32-byte aligned", i); return PGR_E_ARG; }
        if (offs[i] < SEQ_SLACK) { set_error("need %zu bytes of slack before the first sequence", SEQ_SLACK); return PGR_E_ARG; }
        if (lens[i] > 0x7FFFFF00ull) { set_error("sequence %zu longer than 2^31", i); return PGR_E_LIMIT; }
        ctx->h_off[i] = offs[i];
        ctx->h_len[i] = (uint32_t)lens[i];
        total += lens[i];
    }
    ctx->d_seq = dev_base;
    ctx->total_bases = total;
    return set_seq_tables(ctx, n, rids);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
namespace {

struct Level {
    const pgr_mm128 *in; uint64_t n_in; const uint64_t *off_in;
    pgr_mm128 *out; uint64_t *off_out;
};

// one reduce_shmmr level (kind 0) or the min_span filter (kind 1) in a single pass over the list (decoupled look-back)
int run_level_3pass(pgr_b200_ctx *ctx, int kind, const pgr_mm128 *in, uint64_t n_in, const uint64_t *off_in, pgr_mm128 *out,
                    uint64_t *off_out, const pgr_shmmr_spec &spec, int padding, bool patch_rid, uint64_t *n_out, const ChunkView &cv);

int run_level(pgr_b200_ctx *ctx, int kind, const pgr_mm128 *in, uint64_t n_in, const uint64_t *off_in, pgr_mm128 *out,
              uint64_t *off_out, const pgr_shmmr_spec &spec, int padding, bool patch_rid, uint64_t *n_out, const ChunkView &cv) {
    // measured on config 2 (1.2e8 level-0 entries): three-pass 2.84 ms, single-pass look-back 3.48 ms (2048-entry tiles are
    // too small to amortise the look-back latency with ~600 CTAs in flight) -> the three-pass path is the default
    static const bool one_pass = getenv("PGR_B200_LEVELS_1PASS") != nullptr;   // A/B switch (tuning aid)
    if (!one_pass || cv.prefix) return run_level_3pass(ctx, kind, in, n_in, off_in, out, off_out, spec, padding, patch_rid, n_out, cv);
    cudaStream_t st = ctx->stream;
    const size_t n = ctx->rn;
    if (n_in == 0) {
        PGR_CUDA(cudaMemsetAsync(off_out, 0, (n + 1) * sizeof(uint64_t), st));
        *n_out = 0;
        return PGR_OK;
    }
    const uint64_t n_tiles64 = ceil_div<uint64_t>(n_in, LF_TILE);
    if (n_tiles64 >= 0x7FFFFFFFull) { set_error("list too long"); return PGR_E_LIMIT; }
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    PGR_TRY(ctx->block_prefix.ensure(((size_t)n_tiles + 4) * sizeof(uint64_t)));
    PGR_CUDA(cudaMemsetAsync(ctx->block_prefix.p, 0, ((size_t)n_tiles + 4) * sizeof(uint64_t), st));
    LevelFusedParams p;
    p.in = in; p.n_in = n_in; p.seq_off_in = off_in; p.n_seq = (uint32_t)n;
    p.r = spec.r; p.padding = padding ? 1u : 0u; p.min_span = spec.min_span;
    p.out = out; p.seq_off_out = off_out; p.rid = ctx->d_rid.as<uint32_t>() + ctx->r0; p.patch_rid = patch_rid ? 1u : 0u;
    p.tile_status = (unsigned long long *)ctx->block_prefix.p;
    p.total_out = ctx->block_prefix.as<uint64_t>() + n_tiles;
    p.ticket = (uint32_t *)(ctx->block_prefix.as<uint64_t>() + n_tiles + 1);
    static bool attr_set = false;
    if (!attr_set) {
        PGR_CUDA(cudaFuncSetAttribute(level_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LevelFusedSmem)));
        PGR_CUDA(cudaFuncSetAttribute(level_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LevelFusedSmem)));
        attr_set = true;
    }
    if (kind == 0) level_fused_kernel<0><<<n_tiles, LF_NT, sizeof(LevelFusedSmem), st>>>(p);
    else level_fused_kernel<1><<<n_tiles, LF_NT, sizeof(LevelFusedSmem), st>>>(p);
    ctx->counters[0] += 1;
    PGR_CUDA(cudaGetLastError());
    PGR_TRY(ctx->ensure_ctl(64));
    PGR_CUDA(cudaMemcpyAsync(ctx->h_ctl, p.total_out, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    *n_out = *(uint64_t *)ctx->h_ctl;
    return PGR_OK;
}

// the same level as three kernels: flags -> block scan -> ordered scatter
int run_level_3pass(pgr_b200_ctx *ctx, int kind, const pgr_mm128 *in, uint64_t n_in, const uint64_t *off_in, pgr_mm128 *out,
              uint64_t *off_out, const pgr_shmmr_spec &spec, int padding, bool patch_rid, uint64_t *n_out, const ChunkView &cv) {
    cudaStream_t st = ctx->stream;
    const size_t n = ctx->rn;
    if (n_in == 0) {
        PGR_CUDA(cudaMemsetAsync(off_out, 0, (n + 1) * sizeof(uint64_t), st));
        *n_out = 0;
        return PGR_OK;
    }
    if (n_in >= (1ull << 32) * LV_BLK / 2) { set_error("list too long"); return PGR_E_LIMIT; }
    const uint32_t n_blocks = (uint32_t)ceil_div<uint64_t>(n_in, LV_BLK);
    PGR_TRY(ctx->flags.ensure(n_in));
    PGR_TRY(ctx->block_sum.ensure((size_t)n_blocks * sizeof(uint32_t)));
    PGR_TRY(ctx->block_prefix.ensure(((size_t)n_blocks + 1) * sizeof(uint64_t)));
    LevelParams p;
    p.in = in; p.n_in = n_in; p.seq_off_in = off_in; p.n_seq = (uint32_t)n;
    p.r = spec.r; p.padding = padding ? 1u : 0u; p.min_span = spec.min_span;
    p.flags = ctx->flags.as<uint8_t>(); p.block_sum = ctx->block_sum.as<uint32_t>();
    p.block_prefix = ctx->block_prefix.as<uint64_t>();
    p.out = out; p.seq_off_out = off_out; p.rid = (ctx->d_rid.as<uint32_t>() + ctx->r0); p.patch_rid = patch_rid ? 1u : 0u;
    // PGR_B200_LEVELS_UNTILED=1 selects the first-generation kernels (per-entry offset look-ups, strided scatter): A/B aid
    static const bool untiled = getenv("PGR_B200_LEVELS_UNTILED") != nullptr;
    if (untiled && !cv.prefix) {
        if (kind == 0) level_flags_kernel<0><<<n_blocks, LV_NT, 0, st>>>(p);
        else level_flags_kernel<1><<<n_blocks, LV_NT, 0, st>>>(p);
        block_scan_kernel<<<1, 1024, 0, st>>>(p.block_sum, p.block_prefix, n_blocks);
        level_scatter_kernel<<<n_blocks, LV_NT, 0, st>>>(p);
    } else {
        ChunkView v = cv;
        if (v.prefix) {   // chunked input: one small kernel finds every block's first chunk
            PGR_TRY(ctx->block_chunk.ensure((size_t)n_blocks * sizeof(uint32_t)));
            v.block_chunk = ctx->block_chunk.as<uint32_t>();
            block_chunk_kernel<<<ceil_div<uint32_t>(n_blocks, 256), 256, 0, st>>>(v, n_blocks, n_in, ctx->block_chunk.as<uint32_t>());
            ctx->counters[0] += 1;
        }
        const ChunkView &cv = v;
        if (kind == 0) level_flags_tiled_kernel<0><<<n_blocks, LV_NT, 0, st>>>(p, cv);
        else level_flags_tiled_kernel<1><<<n_blocks, LV_NT, 0, st>>>(p, cv);
        block_scan_kernel<<<1, 1024, 0, st>>>(p.block_sum, p.block_prefix, n_blocks);
        level_scatter_tiled_kernel<<<n_blocks, LV_NT, 0, st>>>(p, cv);
    }
    ctx->counters[0] += 3;
    PGR_CUDA(cudaGetLastError());
    PGR_TRY(ctx->ensure_ctl(64));
    PGR_CUDA(cudaMemcpyAsync(ctx->h_ctl, p.block_prefix + n_blocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    *n_out = *(uint64_t *)ctx->h_ctl;
    return PGR_OK;
}

template <int W, int K, int U>
int launch_l0u(const L0Params &p, int grid, cudaStream_t st) {
    PGR_CUDA(cudaFuncSetAttribute(l0_kernel<W, K, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(L0Smem)));
    l0_kernel<W, K, U><<<grid, L0_NT, sizeof(L0Smem), st>>>(p);
    PGR_CUDA(cudaGetLastError());
    return PGR_OK;
}
template <int W, int K>
int launch_l0(const L0Params &p, int grid, cudaStream_t st) {
    // PGR_B200_L0_UNROLL selects the key-loop unroll factor of the w=80 kernel (tuning aid; 4 measured best, 8 within 1 %)
    static const int u = getenv("PGR_B200_L0_UNROLL") ? atoi(getenv("PGR_B200_L0_UNROLL")) : 4;
    if (W == 80 && u == 8) return launch_l0u<W, K, 8>(p, grid, st);
    if (W == 80 && u == 16) return launch_l0u<W, K, 16>(p, grid, st);
    return launch_l0u<W, K, 4>(p, grid, st);
}

template <int W, int K>
int occupancy_l0(int *occ) {
    PGR_CUDA(cudaFuncSetAttribute(l0_kernel<W, K, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(L0Smem)));
    PGR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, l0_kernel<W, K, 4>, L0_NT, sizeof(L0Smem)));
    return PGR_OK;
}

// Replace the level-0 entries around the marked blocks (palindromes, bytes outside ACGTacgt) by an exact replay of the
// reference machine (patch_kernels.cuh), then splice: flat list moves bufA -> bufB -> (swap) bufA, offsets updated in seq_dst.
int apply_patches(pgr_b200_ctx *ctx, const pgr_shmmr_spec &spec, uint32_t n_marks, const std::vector<uint32_t> &flagged,
                  const std::vector<uint64_t> &off0, uint64_t word_lo, uint64_t word_hi, uint64_t *n_l0) {
    cudaStream_t st = ctx->stream;
    const size_t n = ctx->rn;
    const uint32_t gap = (6 * spec.w + spec.k + 64 + 31) / 32 + 1;
    // scratch of the patch pass lives in the context (grow-only): allocating and releasing ten buffers per call cost more
    // than the kernels (every release synchronises the device)
    DevBuf &d_sorted = ctx->patch_buf[0], &d_clusters = ctx->patch_buf[1], &d_q = ctx->patch_buf[2], &d_nadd = ctx->patch_buf[3],
           &d_nfill = ctx->patch_buf[4], &d_eoff = ctx->patch_buf[5], &d_entries = ctx->patch_buf[6], &d_meta = ctx->patch_buf[7],
           &d_slab = ctx->patch_buf[8], &d_list = ctx->patch_buf[9], &d_hslab = ctx->patch_buf[10];
    auto cleanup = [&]() {};
    // device -> host copies of the per-cluster tables go through the page-locked control buffer (a pageable destination
    // makes the driver stage the copy, ~1 ms for the 2-3 MB of a 95 k cluster batch)
    auto d2h = [&](void *dst, const void *src, size_t bytes) -> bool {
        if (bytes == 0) return true;
        if (ctx->ensure_ctl(bytes) != PGR_OK) return false;
        if (cudaMemcpyAsync(ctx->h_ctl, src, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return false;
        memcpy(dst, ctx->h_ctl, bytes);
        return true;
    };
    int rc = PGR_OK;
#define PATCH_TRY(x) do { rc = (x); if (rc != PGR_OK) { cleanup(); return rc; } } while (0)
#define PATCH_CUDA(x) do { if ((x) != cudaSuccess) { set_error("%s failed: %s", #x, cudaGetErrorString(cudaGetLastError())); cleanup(); return PGR_E_CUDA; } } while (0)
    const int slot_t = ctx->timer.begin("l0_patches", st);
    // 1. clusters of marked blocks.  The finder looks sequences up by store offset: sorted copy of the chunk's tables
    std::vector<uint32_t> ord(n);
    for (size_t i = 0; i < n; i++) ord[i] = (uint32_t)i;
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return ctx->h_off[ctx->r0 + a] < ctx->h_off[ctx->r0 + b]; });
    std::vector<uint64_t> s_off(n);
    std::vector<uint32_t> s_len(n);
    for (size_t i = 0; i < n; i++) { s_off[i] = ctx->h_off[ctx->r0 + ord[i]]; s_len[i] = ctx->h_len[ctx->r0 + ord[i]]; }
    PATCH_TRY(d_sorted.ensure(n * 16 + 64));
    uint64_t *ds_off = d_sorted.as<uint64_t>();
    uint32_t *ds_len = (uint32_t *)(ds_off + n), *ds_sid = ds_len + n;
    PATCH_CUDA(cudaMemcpyAsync(ds_off, s_off.data(), n * 8, cudaMemcpyHostToDevice, st));
    PATCH_CUDA(cudaMemcpyAsync(ds_len, s_len.data(), n * 4, cudaMemcpyHostToDevice, st));
    PATCH_CUDA(cudaMemcpyAsync(ds_sid, ord.data(), n * 4, cudaMemcpyHostToDevice, st));
    bool in_order = true;                              // store order == sequence order: the finder's output needs no sort
    for (size_t i = 0; i < n; i++) in_order = in_order && ord[i] == i;
    const uint32_t cf_grid = (uint32_t)ceil_div<uint64_t>(word_hi - word_lo, CF_NT);
    PATCH_TRY(ctx->block_sum.ensure((size_t)cf_grid * sizeof(uint32_t)));
    PATCH_TRY(ctx->block_prefix.ensure(((size_t)cf_grid + 1) * sizeof(uint64_t)));
    ClusterFindParams cf;
    cf.bits = ctx->mark_bits.as<uint32_t>(); cf.word_lo = word_lo; cf.word_hi = word_hi;
    cf.s_off = ds_off; cf.s_len = ds_len; cf.s_sid = ds_sid; cf.n_seq = (uint32_t)n; cf.gap = gap;
    cf.out = nullptr; cf.cap = 0; cf.n_out = nullptr;
    const int slot_find = ctx->timer.begin("patch_kernel_find_count", st);
    cluster_find_kernel<0><<<cf_grid, CF_NT, 0, st>>>(cf, ctx->block_sum.as<uint32_t>(), nullptr);
    block_scan_kernel<<<1, 1024, 0, st>>>(ctx->block_sum.as<uint32_t>(), ctx->block_prefix.as<uint64_t>(), cf_grid);
    ctx->timer.end(slot_find, st);
    PATCH_CUDA(cudaGetLastError());
    uint64_t n_cl64 = 0;
    PATCH_CUDA(cudaMemcpyAsync(&n_cl64, ctx->block_prefix.as<uint64_t>() + cf_grid, 8, cudaMemcpyDeviceToHost, st));
    PATCH_CUDA(cudaStreamSynchronize(st));
    if (n_cl64 > 0xFFFFFFF0ull) { set_error("too many clusters"); cleanup(); return PGR_E_LIMIT; }
    const uint32_t n_cl = (uint32_t)n_cl64;
    PATCH_TRY(d_clusters.ensure((size_t)std::max<uint32_t>(1, n_cl) * sizeof(Cluster)));
    cf.out = d_clusters.as<Cluster>(); cf.cap = n_cl;
    { const int sl = ctx->timer.begin("patch_kernel_find_write", st);
      cluster_find_kernel<1><<<cf_grid, CF_NT, 0, st>>>(cf, nullptr, ctx->block_prefix.as<uint64_t>());
      ctx->timer.end(sl, st); }
    PATCH_CUDA(cudaGetLastError());
    trace_mark("patches: clusters found");
    std::vector<Cluster> cl(n_cl);
    if (!d2h(cl.data(), d_clusters.p, (size_t)n_cl * sizeof(Cluster))) { set_error("D2H of the cluster table failed"); return PGR_E_CUDA; }
    if (!in_order) std::sort(cl.begin(), cl.end(), [](const Cluster &a, const Cluster &b) { return a.sid != b.sid ? a.sid < b.sid : a.pos < b.pos; });
    // whole-sequence replays (flagged) need no patch; the rest in (sequence, position) order
    bool any_flagged = false;
    for (size_t i = 0; i < n; i++) any_flagged = any_flagged || flagged[i] != 0;
    if (any_flagged) cl.erase(std::remove_if(cl.begin(), cl.end(), [&](const Cluster &c) { return flagged[c.sid] != 0; }), cl.end());
    size_t P = cl.size();
    if (P == 0) { ctx->timer.end(slot_t, st); cleanup(); return PGR_OK; }
    if (any_flagged || !in_order) PATCH_CUDA(cudaMemcpyAsync(d_clusters.p, cl.data(), P * sizeof(Cluster), cudaMemcpyHostToDevice, st));
    // 2. replay.  Pass A: one thread per cluster, entries into per-cluster slabs.  The few clusters that run out of steps or
    //    of slab (stuck machines waiting for a small key, Mb-scale gaps with many entries) are redone by one warp each:
    //    count pass, then write pass at exact offsets.
    PATCH_TRY(d_q.ensure(P * 12)); PATCH_TRY(d_nadd.ensure(P * 8)); PATCH_TRY(d_nfill.ensure(P * 4)); PATCH_TRY(d_eoff.ensure(P * 8));
    if ((rc = d_slab.ensure(P * (size_t)RC_SLAB * sizeof(pgr_mm128))) != PGR_OK) { cleanup(); return rc; }
    ReplayClusterParams rp;
    rp.seq = ctx->d_seq; rp.off = ctx->d_off.as<uint64_t>() + ctx->r0; rp.len = ctx->d_len.as<uint32_t>() + ctx->r0;
    rp.mark_bits = ctx->mark_bits.as<uint32_t>(); rp.allinv_bits = ctx->allinv_bits.as<uint32_t>();
    rp.clusters = d_clusters.as<Cluster>(); rp.list = nullptr; rp.n_list = (uint32_t)P; rp.w = spec.w; rp.k = spec.k; rp.gap = gap;
    rp.q0 = d_q.as<uint32_t>(); rp.q1 = rp.q0 + P; rp.t_stop = rp.q0 + 2 * P; rp.n_add = d_nadd.as<uint64_t>(); rp.n_fill = d_nfill.as<uint32_t>();
    rp.entry_off = d_eoff.as<uint64_t>(); rp.entries = nullptr; rp.slab = d_slab.as<pgr_mm128>(); rp.hslab = nullptr;
    DevBuf d_cyc;
    const bool dbg_cycles = getenv("PGR_B200_DEBUG_PATCHES") != nullptr;
    if (dbg_cycles && (rc = d_cyc.ensure(P * 8)) != PGR_OK) { cleanup(); return rc; }
    rp.cycles = dbg_cycles ? d_cyc.as<unsigned long long>() : nullptr;
    { const int sl = ctx->timer.begin("patch_kernel_thread_per_cluster", st);
      cluster_replay_kernel<0, 0><<<ceil_div<uint32_t>((uint32_t)P, RC_NT), RC_NT, 0, st>>>(rp);
      ctx->timer.end(sl, st); }
    if (cudaGetLastError() != cudaSuccess) { set_error("cluster_replay_kernel launch failed"); cleanup(); d_cyc.release(); return PGR_E_CUDA; }
    std::vector<uint32_t> h_q(3 * P), h_nfill(P);
    std::vector<uint64_t> h_nadd(P);
    auto fetch = [&]() -> bool {
        return d2h(h_q.data(), d_q.p, P * 12) && d2h(h_nadd.data(), d_nadd.p, P * 8) && d2h(h_nfill.data(), d_nfill.p, P * 4);
    };
    if (!fetch()) { set_error("D2H of the patch table failed"); cleanup(); d_cyc.release(); return PGR_E_CUDA; }
    trace_mark("patches: pass A (thread per cluster)");
    std::vector<uint32_t> heavy;
    for (size_t j = 0; j < P; j++) if (h_q[2 * P + j] == RC_HEAVY) heavy.push_back((uint32_t)j);
    std::vector<uint8_t> is_heavy(P, 0);
    for (uint32_t j : heavy) is_heavy[j] = 1;
    if (!heavy.empty()) {
        if ((rc = d_list.ensure(heavy.size() * 4)) != PGR_OK) { cleanup(); d_cyc.release(); return rc; }
        cudaMemcpyAsync(d_list.p, heavy.data(), heavy.size() * 4, cudaMemcpyHostToDevice, st);
        rp.list = d_list.as<uint32_t>(); rp.n_list = (uint32_t)heavy.size();
        if ((rc = d_hslab.ensure(heavy.size() * (size_t)RC_HSLAB * sizeof(pgr_mm128))) != PGR_OK) return rc;
        rp.hslab = d_hslab.as<pgr_mm128>();
        { const int sl = ctx->timer.begin("patch_kernel_warp_per_heavy_cluster", st);
          cluster_replay_kernel<0, 1><<<ceil_div<uint32_t>(rp.n_list, RC_NT / 32), RC_NT, 0, st>>>(rp);
          ctx->timer.end(sl, st); }
        if (cudaGetLastError() != cudaSuccess || !fetch()) { set_error("heavy cluster replay failed"); cleanup(); d_cyc.release(); return PGR_E_CUDA; }
        trace_mark("patches: heavy clusters, count pass");
    }
    ctx->counters[3] = heavy.size();
    if (dbg_cycles) {
        std::vector<unsigned long long> cyc(P);
        cudaMemcpy(cyc.data(), d_cyc.p, P * 8, cudaMemcpyDeviceToHost);
        std::vector<size_t> ix(P);
        for (size_t j = 0; j < P; j++) ix[j] = j;
        std::sort(ix.begin(), ix.end(), [&](size_t a, size_t b) { return cyc[a] > cyc[b]; });
        for (size_t t = 0; t < std::min<size_t>(12, P); t++) {
            const size_t j = ix[t];
            fprintf(stderr, "[slow patch] %.3f ms sid %u len %u pos %u q0 %u q1 %u t_stop %u n_add %llu n_fill %u heavy %d\n", cyc[j] / 1.9e6, cl[j].sid,
                    ctx->h_len[ctx->r0 + cl[j].sid], cl[j].pos, h_q[j], h_q[P + j], h_q[2 * P + j], (unsigned long long)h_nadd[j], h_nfill[j], (int)is_heavy[j]);
        }
        if (atoi(getenv("PGR_B200_DEBUG_PATCHES")) > 1)
            for (size_t j = 0; j < P; j++) fprintf(stderr, "[patch] sid %u pos %u q0 %u q1 %u t_stop %u n_add %llu n_fill %u\n", cl[j].sid, cl[j].pos, h_q[j], h_q[P + j], h_q[2 * P + j], (unsigned long long)h_nadd[j], h_nfill[j]);
    }
    d_cyc.release();
    // a replay that ran past the start of later clusters (stuck machine, or to the end of the sequence) subsumes them: their own
    // replays assumed a machine in its normal regime and are dropped
    std::vector<uint32_t> keep;       // indices into the cluster table of pass A
    {
        uint32_t cur_sid = 0xFFFFFFFFu, reach = 0;   // stop position of the last kept patch of cur_sid
        for (size_t j = 0; j < P; j++) {
            if (cl[j].sid == cur_sid && (reach == 0xFFFFFFFFu || cl[j].pos <= reach)) continue;
            cur_sid = cl[j].sid; reach = h_q[2 * P + j];
            keep.push_back((uint32_t)j);
        }
        ctx->counters[7] += P - keep.size();
    }
    const size_t P0 = P;              // size of the pass-A table (slab index space)
    // entries of the kept heavy clusters: exact offsets, write pass
    uint64_t n_entries = 0, n_fills = 0, n_heavy_entries = 0;
    std::vector<uint64_t> eoff(P0, 0);
    std::vector<uint32_t> heavy_keep;
    // heavy clusters whose entries fit the slab of the count pass are final; only the others are replayed once more
    std::vector<uint32_t> heavy_slot(P0, 0);   // position of a heavy cluster in the count pass's list = its slab
    for (size_t t = 0; t < heavy.size(); t++) heavy_slot[heavy[t]] = (uint32_t)t;
    for (uint32_t j : keep) {
        n_entries += h_nadd[j]; n_fills += h_nfill[j];
        if (is_heavy[j] && h_nadd[j] > (uint64_t)RC_HSLAB) { eoff[j] = n_heavy_entries; n_heavy_entries += h_nadd[j]; heavy_keep.push_back(j); }
    }
    if ((rc = d_entries.ensure(std::max<uint64_t>(1, n_heavy_entries) * sizeof(pgr_mm128))) != PGR_OK) { cleanup(); return rc; }
    if (!heavy_keep.empty()) {
        cudaMemcpyAsync(d_eoff.p, eoff.data(), P0 * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_list.p, heavy_keep.data(), heavy_keep.size() * 4, cudaMemcpyHostToDevice, st);
        rp.list = d_list.as<uint32_t>(); rp.n_list = (uint32_t)heavy_keep.size(); rp.entries = d_entries.as<pgr_mm128>(); rp.cycles = nullptr;
        cluster_replay_kernel<1, 1><<<ceil_div<uint32_t>(rp.n_list, RC_NT / 32), RC_NT, 0, st>>>(rp);
        if (cudaGetLastError() != cudaSuccess) { set_error("heavy cluster write pass failed"); cleanup(); return PGR_E_CUDA; }
    }
    // compact the per-patch tables to the kept patches
    std::vector<const pgr_mm128 *> psrc;
    {
        std::vector<Cluster> kcl; std::vector<uint32_t> kq0, kq1, knf; std::vector<uint64_t> kna;
        for (uint32_t j : keep) {
            kcl.push_back(cl[j]); kq0.push_back(h_q[j]); kq1.push_back(h_q[P0 + j]); kna.push_back(h_nadd[j]); knf.push_back(h_nfill[j]);
            psrc.push_back(!is_heavy[j] ? d_slab.as<pgr_mm128>() + (size_t)j * RC_SLAB
                           : h_nadd[j] > (uint64_t)RC_HSLAB ? d_entries.as<pgr_mm128>() + eoff[j] : d_hslab.as<pgr_mm128>() + (size_t)heavy_slot[j] * RC_HSLAB);
        }
        cl.swap(kcl); h_nadd.swap(kna); h_nfill.swap(knf);
        h_q.assign(kq0.begin(), kq0.end());
        h_q.insert(h_q.end(), kq1.begin(), kq1.end());
        P = keep.size();
    }
    trace_mark("patches: heavy write pass + tables");
    // 4. splice.  Patch metadata on the device: [pseq | pq0 | pq1 | plb | pub] u32, then pn_add, pentry_off, pdelta, pdst (8 B each),
    //    off1[n+1] u64, seq_first_patch[n] i32, seq_n_patch[n] u32
    std::vector<uint32_t> pseq(P);
    std::vector<int32_t> seq_first(n, -1);
    std::vector<uint32_t> seq_np(n, 0);
    for (size_t j = 0; j < P; j++) { pseq[j] = cl[j].sid; if (seq_first[cl[j].sid] < 0) seq_first[cl[j].sid] = (int32_t)j; seq_np[cl[j].sid]++; }
    PATCH_TRY(d_meta.ensure(P * 4 * 6 + P * 8 * 4 + (n + 1) * 8 + n * 8 + 64));
    uint32_t *m_u32 = d_meta.as<uint32_t>();
    uint64_t *m_u64 = (uint64_t *)(((uintptr_t)(m_u32 + 5 * P) + 7) & ~(uintptr_t)7);
    int32_t *m_first = (int32_t *)(m_u64 + 4 * P + (n + 1));
    uint32_t *m_np = (uint32_t *)(m_first + n);
    PATCH_CUDA(cudaMemcpyAsync(m_u32, pseq.data(), P * 4, cudaMemcpyHostToDevice, st));
    PATCH_CUDA(cudaMemcpyAsync(m_u32 + P, h_q.data(), P * 8, cudaMemcpyHostToDevice, st));           // pq0 | pq1
    PATCH_CUDA(cudaMemcpyAsync(m_u64, h_nadd.data(), P * 8, cudaMemcpyHostToDevice, st));
    PATCH_CUDA(cudaMemcpyAsync(m_u64 + P, psrc.data(), P * 8, cudaMemcpyHostToDevice, st));
    PATCH_CUDA(cudaMemcpyAsync(m_first, seq_first.data(), n * 4, cudaMemcpyHostToDevice, st));
    PATCH_CUDA(cudaMemcpyAsync(m_np, seq_np.data(), n * 4, cudaMemcpyHostToDevice, st));
    SpliceParams sp;
    sp.flat0 = ctx->bufA.as<pgr_mm128>(); sp.off0 = ctx->seq_dst.as<uint64_t>(); sp.n_seq = (uint32_t)n;
    sp.seq_first_patch = m_first; sp.seq_n_patch = m_np;
    sp.pseq = m_u32; sp.pq0 = m_u32 + P; sp.pq1 = m_u32 + 2 * P; sp.plb = m_u32 + 3 * P; sp.pub = m_u32 + 4 * P;
    sp.pn_add = m_u64; sp.psrc = (const pgr_mm128 *const *)(m_u64 + P); sp.pdelta = (const int64_t *)(m_u64 + 2 * P); sp.pdst = m_u64 + 3 * P; sp.off1 = m_u64 + 4 * P;
    sp.n_patches = (uint32_t)P; sp.n0 = off0[n];
    splice_bounds_kernel<<<ceil_div<uint32_t>((uint32_t)P, 64), 64, 0, st>>>(sp);
    PATCH_CUDA(cudaGetLastError());
    std::vector<uint32_t> plb(P), pub(P);
    if (!d2h(plb.data(), sp.plb, P * 4) || !d2h(pub.data(), sp.pub, P * 4)) { set_error("D2H of the patch bounds failed"); return PGR_E_CUDA; }
    trace_mark("patches: write pass + bounds");
    // consecutive patches of a sequence must not overlap (the cluster gap guarantees it)
    for (size_t j = 0; j < P; j++) {
        if (pub[j] < plb[j] || (j > 0 && pseq[j] == pseq[j - 1] && plb[j] < pub[j - 1])) {
            set_error("overlapping level-0 patches (sequence %u len %u w %u k %u gap %u): patch %zu pos %u q0 %u q1 %u lb %u ub %u n_add %llu | prev pos %u q0 %u q1 %u lb %u ub %u",
                      pseq[j], ctx->h_len[ctx->r0 + pseq[j]], spec.w, spec.k, gap, j, cl[j].pos, h_q[j], h_q[P + j], plb[j], pub[j], (unsigned long long)h_nadd[j],
                      j ? cl[j - 1].pos : 0u, j ? h_q[j - 1] : 0u, j ? h_q[P + j - 1] : 0u, j ? plb[j - 1] : 0u, j ? pub[j - 1] : 0u);
            cleanup(); return PGR_E_CUDA;
        }
    }
    std::vector<int64_t> pdelta(P), seq_delta(n, 0);
    std::vector<int64_t> before(P);
    for (size_t j = 0; j < P; j++) {
        before[j] = seq_delta[pseq[j]];
        seq_delta[pseq[j]] += (int64_t)h_nadd[j] - (int64_t)(pub[j] - plb[j]);
        pdelta[j] = seq_delta[pseq[j]];          // shift of the entries that follow patch j
    }
    std::vector<uint64_t> off1(n + 1), pdst(P);
    uint64_t acc = 0;
    for (size_t i = 0; i < n; i++) { off1[i] = acc; acc += (uint64_t)((int64_t)(off0[i + 1] - off0[i]) + seq_delta[i]); }
    off1[n] = acc;
    for (size_t j = 0; j < P; j++) pdst[j] = off1[pseq[j]] + plb[j] + (uint64_t)before[j];
    PATCH_CUDA(cudaMemcpyAsync((void *)sp.pdelta, pdelta.data(), P * 8, cudaMemcpyHostToDevice, st));
    PATCH_CUDA(cudaMemcpyAsync((void *)sp.pdst, pdst.data(), P * 8, cudaMemcpyHostToDevice, st));
    PATCH_CUDA(cudaMemcpyAsync((void *)sp.off1, off1.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    PATCH_TRY(ctx->bufB.ensure(std::max<uint64_t>(1, acc) * sizeof(pgr_mm128)));
    sp.flat1 = ctx->bufB.as<pgr_mm128>();
    { const int sl = ctx->timer.begin("patch_kernel_splice", st);
      if (sp.n0) splice_copy_kernel<<<(uint32_t)ceil_div<uint64_t>(sp.n0, 256), 256, 0, st>>>(sp);
      splice_patch_kernel<<<(uint32_t)P, 256, 0, st>>>(sp);
      ctx->timer.end(sl, st); }
    PATCH_CUDA(cudaGetLastError());
    PATCH_CUDA(cudaMemcpyAsync(ctx->seq_dst.p, off1.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    ctx->timer.end(slot_t, st);
    PATCH_CUDA(cudaStreamSynchronize(st));
    trace_mark("patches: splice");
    ctx->counters[0] += 6;
    ctx->counters[4] = P;
    ctx->counters[5] = n_fills;
    ctx->counters[6] = n_entries;
    std::swap(ctx->bufA, ctx->bufB);
    PATCH_TRY(ctx->bufB.ensure(std::max<uint64_t>(1, acc) * sizeof(pgr_mm128)));
    *n_l0 = acc;
    cleanup();
#undef PATCH_TRY
#undef PATCH_CUDA
    return PGR_OK;
}

// level-0 minimizers for the whole store -> flat list in ctx->bufA, per-sequence offsets in ctx->seq_dst
int run_l0(pgr_b200_ctx *ctx, const pgr_shmmr_spec &spec, uint64_t *n_l0) {
    ctx->l0_chunked = false;
    cudaStream_t st = ctx->stream;
    const size_t n = ctx->rn;
    const uint32_t w = spec.w, k = spec.k;
    const uint32_t halo = ((w - 1 + 31) / 32) * 32;
    const uint32_t TI = L0_KPOS - 2 * halo;
    // tiles per sequence
    std::vector<uint32_t> tile_prefix(n + 1);
    uint64_t nt64 = 0;
    for (size_t i = 0; i < n; i++) {
        tile_prefix[i] = (uint32_t)nt64;
        const uint32_t L = ctx->h_len[ctx->r0 + i];
        if (L > k) nt64 += ceil_div<uint32_t>(L, TI);
        if (nt64 >= 0xFFFFFFF0ull) { set_error("too many tiles in one batch"); return PGR_E_LIMIT; }
    }
    tile_prefix[n] = (uint32_t)nt64;
    const uint32_t n_tiles = (uint32_t)nt64;
    PGR_TRY(ctx->seq_dst.ensure((n + 1) * sizeof(uint64_t)));
    if (n_tiles == 0) {
        PGR_CUDA(cudaMemsetAsync(ctx->seq_dst.p, 0, (n + 1) * sizeof(uint64_t), st));
        *n_l0 = 0;
        return PGR_OK;
    }
    int occ = 1;
    const int variant = (w == 80 && k == 56) ? 1 : (w == 48 && k == 56) ? 2 : 0;
    if (variant == 1) PGR_TRY((occupancy_l0<80, 56>(&occ)));
    else if (variant == 2) PGR_TRY((occupancy_l0<48, 56>(&occ)));
    else PGR_TRY((occupancy_l0<0, 0>(&occ)));
    if (occ < 1) { set_error("l0_kernel does not fit on an SM"); return PGR_E_CUDA; }
    const uint32_t G = std::min<uint32_t>(n_tiles, (uint32_t)(ctx->n_sm * occ));
    // static partition of the tile sequence over the CTAs: contiguous, equal tile counts (every tile costs about the
    // same; contiguous ranges keep each CTA's output chunk in sequence/position order)
    std::vector<uint32_t> cta_tile(G + 1);
    uint64_t max_cost = 0;
    for (uint32_t c = 0; c <= G; c++) cta_tile[c] = (uint32_t)(((uint64_t)n_tiles * c) / G);
    for (uint32_t c = 0; c < G; c++) max_cost = std::max<uint64_t>(max_cost, (uint64_t)(cta_tile[c + 1] - cta_tile[c]) * TI);
    PGR_TRY(ctx->tile_prefix.ensure((n + 1) * sizeof(uint32_t)));
    PGR_TRY(ctx->cta_tile.ensure((G + 1) * sizeof(uint32_t)));
    PGR_TRY(ctx->chunk_count.ensure(G * sizeof(uint64_t)));
    PGR_TRY(ctx->seq_count.ensure(n * sizeof(uint32_t)));
    PGR_TRY(ctx->seq_flag.ensure(n * sizeof(uint32_t)));
    PGR_TRY(ctx->n_skips.ensure(64));
    // disturbance bitmaps over the 32-base blocks of the store region this chunk occupies (one bit per block, twice)
    uint64_t blk_lo = ~0ull, blk_hi = 0;
    for (size_t i = 0; i < n; i++) {
        const uint64_t o = ctx->h_off[ctx->r0 + i], l = ctx->h_len[ctx->r0 + i];
        blk_lo = std::min(blk_lo, o >> 5); blk_hi = std::max(blk_hi, (o + l + 31) >> 5);
    }
    const uint64_t word_lo = blk_lo >> 5, word_hi = (blk_hi + 31) >> 5;
    if ((size_t)word_hi * 4 + 64 > ctx->mark_bits.cap) {
        PGR_TRY(ctx->mark_bits.ensure((size_t)word_hi * 4 + 64));
        PGR_TRY(ctx->allinv_bits.ensure((size_t)word_hi * 4 + 64));
        ctx->bits_dirty_lo = 0; ctx->bits_dirty_hi = ctx->mark_bits.cap / 4;   // fresh allocation: clear everything once
    }
    trace_mark("run_l0: partition + buffers");
    PGR_CUDA(cudaMemcpyAsync(ctx->tile_prefix.p, tile_prefix.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemcpyAsync(ctx->cta_tile.p, cta_tile.data(), (G + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    // arena: expected density 2/(w+1) with 2x headroom; exact retry on overflow
    uint64_t chunk_cap = (uint64_t)((double)max_cost * std::min(1.0, 4.0 / (w + 1.0))) + 1024;
    chunk_cap = std::max(chunk_cap, ctx->chunk_cap * 0 + chunk_cap);
    PGR_TRY(ctx->ensure_ctl(G * sizeof(uint64_t) + 2 * n * sizeof(uint32_t) + 128));
    uint64_t *h_chunk = (uint64_t *)ctx->h_ctl;
    uint32_t *h_count = (uint32_t *)(h_chunk + G);
    uint32_t *h_flag = h_count + n;
    uint32_t *h_nskips = h_flag + n;
    for (int attempt = 0;; attempt++) {
        PGR_TRY(ctx->arena.ensure((size_t)G * chunk_cap * sizeof(pgr_mm128)));
        ctx->chunk_cap = chunk_cap;
        PGR_CUDA(cudaMemsetAsync(ctx->seq_count.p, 0, n * sizeof(uint32_t), st));
        PGR_CUDA(cudaMemsetAsync(ctx->seq_flag.p, 0, n * sizeof(uint32_t), st));
        PGR_CUDA(cudaMemsetAsync(ctx->n_skips.p, 0, sizeof(uint32_t), st));
        if (ctx->bits_dirty_hi > ctx->bits_dirty_lo) {   // the bitmaps are cleared only where an earlier launch marked blocks
            PGR_CUDA(cudaMemsetAsync(ctx->mark_bits.as<uint32_t>() + ctx->bits_dirty_lo, 0, (ctx->bits_dirty_hi - ctx->bits_dirty_lo) * 4, st));
            PGR_CUDA(cudaMemsetAsync(ctx->allinv_bits.as<uint32_t>() + ctx->bits_dirty_lo, 0, (ctx->bits_dirty_hi - ctx->bits_dirty_lo) * 4, st));
            ctx->bits_dirty_lo = ctx->bits_dirty_hi = 0;
        }
        L0Params p;
        p.seq = ctx->d_seq; p.off = (ctx->d_off.as<uint64_t>() + ctx->r0); p.len = (ctx->d_len.as<uint32_t>() + ctx->r0);
        p.tile_prefix = ctx->tile_prefix.as<uint32_t>(); p.cta_tile = ctx->cta_tile.as<uint32_t>();
        p.n_seq = (uint32_t)n; p.w = w; p.k = k; p.tile_stride = TI; p.halo = halo;
        p.arena = ctx->arena.as<pgr_mm128>(); p.chunk_cap = chunk_cap;
        p.chunk_count = ctx->chunk_count.as<uint64_t>(); p.seq_count = ctx->seq_count.as<uint32_t>();
        p.seq_flag = ctx->seq_flag.as<uint32_t>();
        p.mark_bits = ctx->mark_bits.as<uint32_t>(); p.allinv_bits = ctx->allinv_bits.as<uint32_t>(); p.n_marks = ctx->n_skips.as<uint32_t>();
        p.m1 = ~0ull;
        p.hs.c24 = 1u << (32 - 24); p.hs.c14 = 1u << (32 - 14); p.hs.c28 = 1u << (32 - 28);
        p.hs.chs = (k > 32) ? 1u << (32 - (64 - k)) : 0u;   // used by the K > 32 kernels only (64 - K < 32)
        trace_mark("run_l0: arena + memsets");
        const int slot = ctx->timer.begin("l0_minimizers", st);
        if (variant == 1) PGR_TRY((launch_l0<80, 56>(p, (int)G, st)));
        else if (variant == 2) PGR_TRY((launch_l0<48, 56>(p, (int)G, st)));
        else PGR_TRY((launch_l0<0, 0>(p, (int)G, st)));
        ctx->timer.end(slot, st);
        ctx->counters[0] += 1;
        trace_mark("run_l0: kernel");
        PGR_CUDA(cudaMemcpyAsync(h_chunk, ctx->chunk_count.p, G * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaMemcpyAsync(h_count, ctx->seq_count.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaMemcpyAsync(h_flag, ctx->seq_flag.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaMemcpyAsync(h_nskips, ctx->n_skips.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
        if (*h_nskips) { ctx->bits_dirty_lo = word_lo; ctx->bits_dirty_hi = word_hi; }
        uint64_t mx = 0;
        for (uint32_t c = 0; c < G; c++) mx = std::max(mx, h_chunk[c]);
        if (mx <= chunk_cap) break;
        if (attempt >= 2) { set_error("level-0 arena overflow persists"); return PGR_E_CUDA; }
        chunk_cap = mx + mx / 16 + 64;
        ctx->counters[3] += 1;
    }
    trace_mark("run_l0: l0 kernel + control read-back");
    // sequences flagged for sequential replay
    std::vector<uint32_t> replay;
    for (size_t i = 0; i < n; i++) if (h_flag[i]) replay.push_back((uint32_t)i);
    std::vector<uint32_t> final_count(h_count, h_count + n);
    if (!replay.empty()) {
        PGR_TRY(ctx->replay_list.ensure(replay.size() * sizeof(uint32_t)));
        PGR_TRY(ctx->replay_count.ensure(n * sizeof(uint32_t)));
        PGR_CUDA(cudaMemcpyAsync(ctx->replay_list.p, replay.data(), replay.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ReplayParams rp;
        rp.seq = ctx->d_seq; rp.off = (ctx->d_off.as<uint64_t>() + ctx->r0); rp.len = (ctx->d_len.as<uint32_t>() + ctx->r0);
        rp.list = ctx->replay_list.as<uint32_t>(); rp.n_list = (uint32_t)replay.size(); rp.w = w; rp.k = k;
        rp.count = ctx->replay_count.as<uint32_t>(); rp.dst_off = nullptr; rp.dst = nullptr;
        const int slot = ctx->timer.begin("l0_replay_count", st);
        replay_l0_kernel<0><<<ceil_div<uint32_t>(rp.n_list, 32), 32, 0, st>>>(rp);
        ctx->timer.end(slot, st);
        ctx->counters[0] += 1;
        PGR_CUDA(cudaGetLastError());
        std::vector<uint32_t> rc(n);
        PGR_CUDA(cudaMemcpyAsync(h_count, ctx->replay_count.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
        for (uint32_t sid : replay) final_count[sid] = h_count[sid];
        // h_count now holds replay counts at flagged slots only; restore semantics below via final_count
    }
    ctx->counters[2] = replay.size();
    // host scans (small arrays)
    std::vector<uint64_t> seq_fast(n), seq_dst(n + 1), chunk_prefix(G + 1);
    {
        // fast-path counts are what the kernel accumulated (before replay substitution)
        uint64_t a = 0, b = 0;
        // re-read the fast counts: they were overwritten in h_count only when a replay happened
        std::vector<uint32_t> fast(n);
        if (replay.empty()) {
            for (size_t i = 0; i < n; i++) fast[i] = final_count[i];
        } else {
            PGR_CUDA(cudaMemcpyAsync(h_count, ctx->seq_count.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            PGR_CUDA(cudaStreamSynchronize(st));
            for (size_t i = 0; i < n; i++) fast[i] = h_count[i];
        }
        for (size_t i = 0; i < n; i++) { seq_fast[i] = a; a += fast[i]; seq_dst[i] = b; b += final_count[i]; }
        seq_dst[n] = b;
        uint64_t c0 = 0;
        for (uint32_t c = 0; c < G; c++) { chunk_prefix[c] = c0; c0 += h_chunk[c]; }
        chunk_prefix[G] = c0;
    }
    const uint64_t total = seq_dst[n];
    *n_l0 = total;
    PGR_TRY(ctx->seq_fast.ensure(n * sizeof(uint64_t)));
    PGR_TRY(ctx->chunk_prefix.ensure((G + 1) * sizeof(uint64_t)));
    PGR_CUDA(cudaMemcpyAsync(ctx->seq_fast.p, seq_fast.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemcpyAsync(ctx->seq_dst.p, seq_dst.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemcpyAsync(ctx->chunk_prefix.p, chunk_prefix.data(), (G + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_TRY(ctx->bufA.ensure(std::max<uint64_t>(total, 1) * sizeof(pgr_mm128)));
    PGR_TRY(ctx->bufB.ensure(std::max<uint64_t>(total, 1) * sizeof(pgr_mm128)));
    // No sequence to replay and no palindrome / invalid-byte record: the level-0 list is never patched, so the next
    // level reads it straight from the arena chunks (ChunkView) and the gather pass is skipped.
    const uint32_t n_skips_seen = *h_nskips;
    static const bool always_gather = getenv("PGR_B200_ALWAYS_GATHER") != nullptr;   // A/B aid
    ctx->l0_chunked = total && replay.empty() && n_skips_seen == 0 && !always_gather;
    ctx->l0_chunk_cap = chunk_cap; ctx->l0_chunks = G;
    if (total && !ctx->l0_chunked) {
        GatherParams gp;
        gp.arena = ctx->arena.as<pgr_mm128>(); gp.chunk_cap = chunk_cap; gp.chunk_prefix = ctx->chunk_prefix.as<uint64_t>();
        gp.n_chunks = G; gp.seq_fast = ctx->seq_fast.as<uint64_t>(); gp.seq_dst = ctx->seq_dst.as<uint64_t>();
        gp.seq_flag = ctx->seq_flag.as<uint32_t>(); gp.flat = ctx->bufA.as<pgr_mm128>();
        const int slot = ctx->timer.begin("l0_gather", st);
        gather_l0_kernel<<<dim3(16, G), 256, 0, st>>>(gp);
        ctx->timer.end(slot, st);
        ctx->counters[0] += 1;
        PGR_CUDA(cudaGetLastError());
    }
    if (!replay.empty()) {
        ReplayParams rp;
        rp.seq = ctx->d_seq; rp.off = (ctx->d_off.as<uint64_t>() + ctx->r0); rp.len = (ctx->d_len.as<uint32_t>() + ctx->r0);
        rp.list = ctx->replay_list.as<uint32_t>(); rp.n_list = (uint32_t)replay.size(); rp.w = w; rp.k = k;
        rp.count = nullptr; rp.dst_off = ctx->seq_dst.as<uint64_t>(); rp.dst = ctx->bufA.as<pgr_mm128>();
        const int slot = ctx->timer.begin("l0_replay_write", st);
        replay_l0_kernel<1><<<ceil_div<uint32_t>(rp.n_list, 32), 32, 0, st>>>(rp);
        ctx->timer.end(slot, st);
        ctx->counters[0] += 1;
        PGR_CUDA(cudaGetLastError());
    }
    // the host vectors above are pageable: make sure the async copies are done before they go out of scope
    PGR_CUDA(cudaStreamSynchronize(st));
    trace_mark("run_l0: replay + gather");
    const uint32_t n_marks = *h_nskips;
    if (n_marks) {
        std::vector<uint32_t> flags(h_flag, h_flag + n);
        PGR_TRY(apply_patches(ctx, spec, n_marks, flags, seq_dst, word_lo, word_hi, n_l0));
        trace_mark("run_l0: patches");
    }
    return PGR_OK;
}

// sketch mode (shmmrutils.rs:558-630): flat list of kept k-mers in ctx->bufA, offsets in ctx->seq_dst (sketch_kernels.cuh)
int run_sketch(pgr_b200_ctx *ctx, const pgr_shmmr_spec &spec, uint64_t *n_l0) {
    cudaStream_t st = ctx->stream;
    const size_t n = ctx->rn;
    std::vector<uint32_t> tile_prefix(n + 1);
    uint64_t n_tiles64 = 0, blk_lo = ~0ull, blk_hi = 0;
    for (size_t i = 0; i < n; i++) {
        const uint64_t o = ctx->h_off[ctx->r0 + i], l = ctx->h_len[ctx->r0 + i];
        tile_prefix[i] = (uint32_t)n_tiles64;
        n_tiles64 += ceil_div<uint64_t>(l, SKT_KPOS);
        if (l) { blk_lo = std::min(blk_lo, o >> 5); blk_hi = std::max(blk_hi, (o + l + 31) >> 5); }
    }
    if (n_tiles64 >= 0x7FFFFFFFull) { set_error("too many tiles in one chunk"); return PGR_E_LIMIT; }
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    tile_prefix[n] = n_tiles;
    PGR_TRY(ctx->seq_dst.ensure((n + 1) * sizeof(uint64_t)));
    if (n_tiles == 0) {
        PGR_CUDA(cudaMemsetAsync(ctx->seq_dst.p, 0, (n + 1) * sizeof(uint64_t), st));
        *n_l0 = 0;
        return PGR_OK;
    }
    const uint64_t n_blk = blk_hi - blk_lo;
    const uint64_t word_lo = blk_lo >> 5, word_hi = (blk_hi + 31) >> 5;
    // buffers (reused from the minimizer pipeline): tile_prefix, seq_count <- tile counts, chunk_prefix <- tile offsets,
    // flags <- keep masks, block_chunk <- strand masks, mark_bits <- marked (not clean) blocks, allinv_bits
    PGR_TRY(ctx->tile_prefix.ensure((n + 1) * sizeof(uint32_t)));
    PGR_TRY(ctx->seq_count.ensure((size_t)n_tiles * sizeof(uint32_t)));
    PGR_TRY(ctx->chunk_prefix.ensure(((size_t)n_tiles + 1) * sizeof(uint64_t)));
    PGR_TRY(ctx->flags.ensure(n_blk * sizeof(uint32_t)));
    PGR_TRY(ctx->block_chunk.ensure(n_blk * sizeof(uint32_t)));
    PGR_TRY(ctx->n_skips.ensure(64));
    if ((size_t)word_hi * 4 + 64 > ctx->mark_bits.cap) {
        PGR_TRY(ctx->mark_bits.ensure((size_t)word_hi * 4 + 64));
        PGR_TRY(ctx->allinv_bits.ensure((size_t)word_hi * 4 + 64));
        ctx->bits_dirty_lo = 0; ctx->bits_dirty_hi = ctx->mark_bits.cap / 4;
    }
    if (ctx->bits_dirty_hi > ctx->bits_dirty_lo) {   // cleared only where an earlier launch marked blocks
        PGR_CUDA(cudaMemsetAsync(ctx->mark_bits.as<uint32_t>() + ctx->bits_dirty_lo, 0, (ctx->bits_dirty_hi - ctx->bits_dirty_lo) * 4, st));
        PGR_CUDA(cudaMemsetAsync(ctx->allinv_bits.as<uint32_t>() + ctx->bits_dirty_lo, 0, (ctx->bits_dirty_hi - ctx->bits_dirty_lo) * 4, st));
        ctx->bits_dirty_lo = ctx->bits_dirty_hi = 0;
    }
    PGR_CUDA(cudaMemsetAsync(ctx->n_skips.p, 0, sizeof(uint32_t), st));
    PGR_CUDA(cudaMemcpyAsync(ctx->tile_prefix.p, tile_prefix.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SketchTParams sp;
    sp.seq = ctx->d_seq; sp.off = (ctx->d_off.as<uint64_t>() + ctx->r0); sp.len = (ctx->d_len.as<uint32_t>() + ctx->r0);
    sp.tile_prefix = ctx->tile_prefix.as<uint32_t>(); sp.n_seq = (uint32_t)n; sp.n_tiles = n_tiles; sp.k = spec.k; sp.r = spec.r;
    sp.blk_base = blk_lo; sp.keep = ctx->flags.as<uint32_t>(); sp.strand = ctx->block_chunk.as<uint32_t>();
    sp.dirty_bits = ctx->mark_bits.as<uint32_t>(); sp.allinv_bits = ctx->allinv_bits.as<uint32_t>(); sp.n_dirty = ctx->n_skips.as<uint32_t>();
    sp.tile_count = ctx->seq_count.as<uint32_t>(); sp.tile_off = nullptr; sp.out = nullptr;
    int slot = ctx->timer.begin("sketch_masks", st);
    if (spec.k == 56) sketch_mask_kernel<56><<<n_tiles, L0_NT, 0, st>>>(sp);
    else sketch_mask_kernel<0><<<n_tiles, L0_NT, 0, st>>>(sp);
    ctx->timer.end(slot, st);
    PGR_CUDA(cudaGetLastError());
    PGR_TRY(ctx->ensure_ctl((size_t)n_tiles * sizeof(uint32_t) + 64));
    uint32_t *h_nd = (uint32_t *)ctx->h_ctl;
    PGR_CUDA(cudaMemcpyAsync(h_nd, ctx->n_skips.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    const uint32_t n_dirty = *h_nd;
    ctx->counters[4] = n_dirty;
    if (n_dirty) {
        // the marked blocks are looked up by store offset: sorted copy of the chunk's sequence table
        ctx->bits_dirty_lo = word_lo; ctx->bits_dirty_hi = word_hi;
        std::vector<uint32_t> ord(n);
        for (size_t i = 0; i < n; i++) ord[i] = (uint32_t)i;
        std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return ctx->h_off[ctx->r0 + a] < ctx->h_off[ctx->r0 + b]; });
        std::vector<uint64_t> s_off(n);
        std::vector<uint32_t> s_len(n);
        for (size_t i = 0; i < n; i++) { s_off[i] = ctx->h_off[ctx->r0 + ord[i]]; s_len[i] = ctx->h_len[ctx->r0 + ord[i]]; }
        DevBuf &d_sorted = ctx->patch_buf[0];
        PGR_TRY(d_sorted.ensure(n * 16 + 64));
        SketchSeqTable tb;
        uint64_t *ds_off = d_sorted.as<uint64_t>();
        uint32_t *ds_len = (uint32_t *)(ds_off + n), *ds_sid = ds_len + n;
        PGR_CUDA(cudaMemcpyAsync(ds_off, s_off.data(), n * 8, cudaMemcpyHostToDevice, st));
        PGR_CUDA(cudaMemcpyAsync(ds_len, s_len.data(), n * 4, cudaMemcpyHostToDevice, st));
        PGR_CUDA(cudaMemcpyAsync(ds_sid, ord.data(), n * 4, cudaMemcpyHostToDevice, st));
        tb.s_off = ds_off; tb.s_len = ds_len; tb.s_sid = ds_sid; tb.n = (uint32_t)n;
        slot = ctx->timer.begin("sketch_marked_blocks", st);
        sketch_dirty_kernel<<<(uint32_t)ceil_div<uint64_t>(word_hi - word_lo, 128), 128, 0, st>>>(sp, tb, word_lo, word_hi);
        ctx->timer.end(slot, st);
        PGR_CUDA(cudaGetLastError());
        PGR_CUDA(cudaStreamSynchronize(st));   // the sorted host tables go out of scope
        ctx->counters[0] += 1;
    }
    sketch_count_kernel<<<(uint32_t)ceil_div<uint64_t>((uint64_t)n_tiles * 32, 256), 256, 0, st>>>(sp);
    PGR_CUDA(cudaGetLastError());
    uint32_t *h_cnt = (uint32_t *)ctx->h_ctl;
    PGR_CUDA(cudaMemcpyAsync(h_cnt, ctx->seq_count.p, (size_t)n_tiles * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    std::vector<uint64_t> tile_off((size_t)n_tiles + 1), seq_dst(n + 1);
    uint64_t a = 0;
    for (uint32_t t = 0; t < n_tiles; t++) { tile_off[t] = a; a += h_cnt[t]; }
    tile_off[n_tiles] = a;
    for (size_t i = 0; i <= n; i++) seq_dst[i] = tile_off[tile_prefix[i]];
    *n_l0 = a;
    PGR_CUDA(cudaMemcpyAsync(ctx->chunk_prefix.p, tile_off.data(), ((size_t)n_tiles + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_CUDA(cudaMemcpyAsync(ctx->seq_dst.p, seq_dst.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGR_TRY(ctx->bufA.ensure(std::max<uint64_t>(a, 1) * sizeof(pgr_mm128)));
    PGR_TRY(ctx->bufB.ensure(std::max<uint64_t>(a, 1) * sizeof(pgr_mm128)));
    sp.tile_off = ctx->chunk_prefix.as<uint64_t>(); sp.out = ctx->bufA.as<pgr_mm128>();
    slot = ctx->timer.begin("sketch_write", st);
    sketch_write_kernel<<<n_tiles, L0_NT, 0, st>>>(sp);
    ctx->timer.end(slot, st);
    ctx->counters[0] += 3;
    PGR_CUDA(cudaGetLastError());
    PGR_CUDA(cudaStreamSynchronize(st));
    return PGR_OK;
}

}  // namespace

extern "C" {

}  // extern "C"

// sequence_to_shmmrs over the sequence range [ctx->r0, ctx->r0 + ctx->rn) of the store
int pgr::shmmrs_range(pgr_b200_ctx *ctx, const pgr_shmmr_spec &spec, int padding, size_t *n_shmmrs) {
    cudaStream_t st = ctx->stream;
    ctx->result_valid = false;
    const size_t n = ctx->rn;
    PGR_TRY(ctx->off_a.ensure((n + 1) * sizeof(uint64_t)));
    PGR_TRY(ctx->off_b.ensure((n + 1) * sizeof(uint64_t)));
    uint64_t n_cur = 0;
    if (spec.sketch) PGR_TRY(run_sketch(ctx, spec, &n_cur));
    else PGR_TRY(run_l0(ctx, spec, &n_cur));
    ctx->counters[1] = n_cur;
    PGR_TRY(ctx->bufA.ensure(std::max<uint64_t>(n_cur, 1) * sizeof(pgr_mm128)));
    PGR_TRY(ctx->bufB.ensure(std::max<uint64_t>(n_cur, 1) * sizeof(pgr_mm128)));
    const pgr_mm128 *cur = ctx->bufA.as<pgr_mm128>();
    const uint64_t *cur_off = ctx->seq_dst.as<uint64_t>();
    ChunkView cv = {nullptr, 0, 0, nullptr};   // flat list
    const ChunkView flat = cv;
    if (!spec.sketch && ctx->l0_chunked) {   // the level-0 list still lives in the arena chunks (run_l0)
        cur = ctx->arena.as<pgr_mm128>();
        cv.prefix = ctx->chunk_prefix.as<uint64_t>(); cv.cap = ctx->l0_chunk_cap; cv.n_chunks = ctx->l0_chunks;
    }
    // reduce_shmmr twice, then the min_span filter: each is flags -> block scan -> ordered scatter over the list.
    // (A fused single-kernel version with the levels as index lists in shared memory was measured slower: 6.6 ms vs
    // 2.8 ms on config 2 -- its per-CTA compactions serialise on barriers.)
    const int slot = ctx->timer.begin("reduce_and_span", st);
    if (!spec.sketch && spec.r > 1) {
        uint64_t n1 = 0, n2 = 0;
        PGR_TRY(run_level(ctx, 0, cur, n_cur, cur_off, ctx->bufB.as<pgr_mm128>(), ctx->off_a.as<uint64_t>(), spec, padding, false, &n1, cv));
        PGR_TRY(run_level(ctx, 0, ctx->bufB.as<pgr_mm128>(), n1, ctx->off_a.as<uint64_t>(), ctx->bufA.as<pgr_mm128>(),
                          ctx->off_b.as<uint64_t>(), spec, padding, false, &n2, flat));
        cur = ctx->bufA.as<pgr_mm128>(); cur_off = ctx->off_b.as<uint64_t>(); n_cur = n2; cv = flat;
    }
    uint64_t n_fin = 0;
    PGR_TRY(run_level(ctx, 1, cur, n_cur, cur_off, ctx->bufB.as<pgr_mm128>(), ctx->off_a.as<uint64_t>(), spec, 0, true, &n_fin, cv));
    ctx->timer.end(slot, st);
    ctx->d_result = ctx->bufB.as<pgr_mm128>();
    ctx->d_result_off = ctx->off_a.as<uint64_t>();
    ctx->n_result = n_fin;

    // padding=true with an empty level-0 list: reduce_shmmr keeps its MAX sentinels and the span filter leaves two
    // of them (shmmrutils.rs:367-380 twice, then :541-553).  Rare API corner; rebuilt on the host.
    if (padding && !spec.sketch && spec.r > 1 && n) {
        std::vector<uint64_t> l0_off(n + 1), fin_off(n + 1);
        PGR_CUDA(cudaMemcpyAsync(l0_off.data(), ctx->seq_dst.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaMemcpyAsync(fin_off.data(), ctx->d_result_off, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
        bool any = false;
        for (size_t i = 0; i < n; i++) any = any || (l0_off[i + 1] == l0_off[i]);
        if (any) {
            std::vector<pgr_mm128> mm(n_fin), out;
            if (n_fin) PGR_CUDA(cudaMemcpy(mm.data(), ctx->d_result, n_fin * sizeof(pgr_mm128), cudaMemcpyDeviceToHost));
            std::vector<uint64_t> noff(n + 1);
            for (size_t i = 0; i < n; i++) {
                noff[i] = out.size();
                if (l0_off[i + 1] == l0_off[i]) {
                    const pgr_mm128 s = {~0ull, ~0ull};
                    out.push_back(s); out.push_back(s);
                } else {
                    out.insert(out.end(), mm.begin() + fin_off[i], mm.begin() + fin_off[i + 1]);
                }
            }
            noff[n] = out.size();
            PGR_TRY(ctx->fix_mm.ensure(std::max<size_t>(1, out.size()) * sizeof(pgr_mm128)));
            PGR_TRY(ctx->fix_off.ensure((n + 1) * sizeof(uint64_t)));
            PGR_CUDA(cudaMemcpy(ctx->fix_mm.p, out.data(), out.size() * sizeof(pgr_mm128), cudaMemcpyHostToDevice));
            PGR_CUDA(cudaMemcpy(ctx->fix_off.p, noff.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
            ctx->d_result = ctx->fix_mm.as<pgr_mm128>();
            ctx->d_result_off = ctx->fix_off.as<uint64_t>();
            ctx->n_result = out.size();
        }
    }
    ctx->result_valid = true;
    if (n_shmmrs) *n_shmmrs = ctx->n_result;
    return PGR_OK;
}

extern "C" {

int pgr_b200_ctx_shmmrs(pgr_b200_ctx *ctx, const pgr_shmmr_spec *spec_in, int padding, size_t *n_shmmrs) {
    if (!ctx) { set_error("ctx is NULL"); return PGR_E_ARG; }
    PGR_TRY(check_spec(spec_in));
    PGR_CUDA(cudaSetDevice(ctx->device));
    ctx->timer.reset();
    memset(ctx->counters, 0, sizeof ctx->counters);
    ctx->r0 = 0;
    ctx->rn = ctx->n_seq;
    return pgr::shmmrs_range(ctx, *spec_in, padding, n_shmmrs);
}

int pgr_b200_ctx_shmmrs_device(pgr_b200_ctx *ctx, const pgr_mm128 **d_mm, const uint64_t **d_offsets) {
    if (!ctx || !ctx->result_valid) { set_error("no shimmer result available"); return PGR_E_ARG; }
    if (d_mm) *d_mm = ctx->d_result;
    if (d_offsets) *d_offsets = ctx->d_result_off;
    return PGR_OK;
}

int pgr_b200_ctx_shmmrs_download(pgr_b200_ctx *ctx, pgr_mm128 **out, size_t *offsets) {
    if (!ctx || !ctx->result_valid || !out || !offsets) { set_error("no shimmer result available / NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(ctx->device));
    const size_t n = ctx->rn;
    pgr_mm128 *o = (pgr_mm128 *)result_alloc(std::max<size_t>(1, ctx->n_result) * sizeof(pgr_mm128));
    if (!o) { set_error("out of host memory"); return PGR_E_ARG; }
    std::vector<uint64_t> off(n + 1);
    cudaError_t e = cudaSuccess;
    if (ctx->n_result) e = cudaMemcpyAsync(o, ctx->d_result, ctx->n_result * sizeof(pgr_mm128), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(off.data(), ctx->d_result_off, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { result_free(o); set_error("D2H failed: %s", cudaGetErrorString(e)); return PGR_E_CUDA; }
    for (size_t i = 0; i <= n; i++) offsets[i] = (size_t)off[i];
    *out = o;
    return PGR_OK;
}

int pgr_b200_ctx_timings(pgr_b200_ctx *ctx, const char *const **names, const float **ms, size_t *n) {
    if (!ctx) { set_error("ctx is NULL"); return PGR_E_ARG; }
    ctx->timer.collect();
    if (names) *names = ctx->timer.names_z.data();
    if (ms) *ms = ctx->timer.ms.data();
    if (n) *n = ctx->timer.used;
    return PGR_OK;
}

int pgr_b200_ctx_counters(pgr_b200_ctx *ctx, uint64_t out[8]) {
    if (!ctx || !out) { set_error("NULL argument"); return PGR_E_ARG; }
    memcpy(out, ctx->counters, sizeof ctx->counters);
    return PGR_OK;
}

// ---- one-shot host API (thread-local default context on device 0) ---------------------------------------------
}  // extern "C"

static thread_local int g_default_device = 0;
static thread_local pgr_b200_ctx *g_tls_ctx = nullptr;  // leaked at thread exit on purpose: CUDA may already be torn down
int pgr::default_device() { return g_default_device; }

extern "C" int pgr_b200_set_default_device(int device) {
    if (device < 0 || device >= pgr_b200_device_count()) { set_error("device %d out of range", device); return PGR_E_ARG; }
    if (g_tls_ctx && g_tls_ctx->device != device) { pgr_b200_ctx_free(g_tls_ctx); g_tls_ctx = nullptr; }
    g_default_device = device;
    return PGR_OK;
}

pgr_b200_ctx *pgr::tls_ctx() {
    if (!g_tls_ctx) g_tls_ctx = pgr_b200_ctx_new(g_default_device);
    return g_tls_ctx;
}

extern "C" {

}  // extern "C"

// Host batch -> device, chunk by chunk.  The batch is cut into chunks of whole sequences; the H2D copies of all chunks
// are queued on a copy stream up front and the shimmer pipeline of chunk c runs on the compute stream as soon as its
// bytes have landed, so that PCIe transfer and kernels overlap.  on_chunk(first_seq, n_seqs, n_shmmrs) consumes the
// chunk's device-resident result (ctx->d_result / ctx->d_result_off) before the next chunk overwrites it.
int pgr::run_chunked(pgr_b200_ctx *ctx, size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens,
                     const pgr_shmmr_spec &spec, int padding, const std::function<int(size_t, size_t, size_t)> &on_chunk) {
    PGR_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) PGR_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    trace_mark("run_chunked: begin");
    PGR_TRY(upload_layout(ctx, n, rids, seqs, lens));
    trace_mark("run_chunked: layout");
    ctx->timer.reset();
    // chunk boundaries: about 1/8 of the batch each, at least 192 MB (every chunk costs ~1.5 ms of launches and control
    // read-backs), whole sequences
    const uint64_t chunk_bytes = std::max<uint64_t>(192ull << 20, ctx->total_bases / 8);
    std::vector<size_t> cut(1, 0);
    {
        uint64_t acc = 0;
        for (size_t i = 0; i < n; i++) {
            acc += lens[i];
            if (acc >= chunk_bytes && i + 1 < n) { cut.push_back(i + 1); acc = 0; }
        }
        cut.push_back(n);
    }
    const size_t n_chunks = cut.size() - 1;
    std::vector<cudaEvent_t> ev(n_chunks);
    int rc = PGR_OK;
    for (size_t c = 0; c < n_chunks; c++) cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming);
    // Packed transport (pack_upload.cuh): the host packs the bases into bit planes while earlier chunks are copied and
    // computed, so the packing runs on its own host thread and hands the chunks over one by one.  Direct transport: the
    // copies are queued up front (asynchronous when the caller's buffers are page-locked).
    const bool packed = choose_packed(seqs, lens, n, ctx->total_bases);
    last_transport().store(packed ? 0 : 1);
    if (packed && !ctx->pack && !(ctx->pack = pack_ring_acquire(ctx->device))) rc = PGR_E_CUDA;
    std::thread uploader;
    std::mutex up_mu;
    std::condition_variable up_cv;
    size_t up_ready = 0;          // chunks whose event has been recorded
    int up_rc = PGR_OK;
    std::string up_err;
    const bool src_locked = packed && source_page_locked(seqs, lens, n);
    if (rc == PGR_OK && packed) {
        uploader = std::thread([&] {
            int r = cudaSetDevice(ctx->device) == cudaSuccess ? PGR_OK : PGR_E_CUDA;
            for (size_t c = 0; c < n_chunks; c++) {
                if (r == PGR_OK) r = upload_packed(ctx->pack, ctx->seq_store.as<uint8_t>(), ctx->h_off, ctx->h_len, seqs, cut[c], cut[c + 1], ctx->copy_stream, src_locked);
                if (r == PGR_OK && cudaEventRecord(ev[c], ctx->copy_stream) != cudaSuccess) r = PGR_E_CUDA;
                std::lock_guard<std::mutex> lk(up_mu);
                if (r != PGR_OK && up_rc == PGR_OK) { up_rc = r; up_err = get_error(); }
                up_ready = c + 1;
                up_cv.notify_all();
            }
        });
    } else {
        for (size_t c = 0; c < n_chunks && rc == PGR_OK; c++) {
            rc = upload_copy(ctx, seqs, lens, cut[c], cut[c + 1], ctx->copy_stream);
            cudaEventRecord(ev[c], ctx->copy_stream);
        }
    }
    uint64_t tot[8] = {0};
    trace_mark("run_chunked: H2D queued/staged");
    for (size_t c = 0; c < n_chunks && rc == PGR_OK; c++) {
        if (uploader.joinable()) {
            std::unique_lock<std::mutex> lk(up_mu);
            up_cv.wait(lk, [&] { return up_ready > c; });
            if (up_rc != PGR_OK) { rc = up_rc; set_error("%s", up_err.c_str()); break; }
        }
        cudaStreamWaitEvent(ctx->stream, ev[c], 0);
        ctx->r0 = cut[c];
        ctx->rn = cut[c + 1] - cut[c];
        memset(ctx->counters, 0, sizeof ctx->counters);
        size_t ns = 0;
        if ((rc = shmmrs_range(ctx, spec, padding, &ns)) != PGR_OK) break;
        for (int i = 0; i < 8; i++) tot[i] += ctx->counters[i];
        trace_mark("run_chunked: chunk reduce+span");
        rc = on_chunk(cut[c], ctx->rn, ns);
        trace_mark("run_chunked: on_chunk");
    }
    if (uploader.joinable()) uploader.join();
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    for (size_t c = 0; c < n_chunks; c++) cudaEventDestroy(ev[c]);
    ctx->r0 = 0; ctx->rn = ctx->n_seq;
    ctx->result_valid = false;  // the device holds only the last chunk
    memcpy(ctx->counters, tot, sizeof tot);
    return rc;
}

extern "C" {

// replaces CompactSeqDB::get_shmmrs_from_seqs: one call per batch with HOST buffers in and out
int pgr_b200_shmmrs_batch(size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens,
                          const pgr_shmmr_spec *spec, int padding, pgr_mm128 **out, size_t *offsets) {
    if (!out || !offsets || (n && (!seqs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_TRY(check_spec(spec));
    pgr_b200_ctx *ctx = pgr::tls_ctx();
    if (!ctx) return PGR_E_NO_DEVICE;
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) total += lens[i];
    // result buffer: expected density + headroom, grown on demand
    size_t cap = std::max<size_t>(4096, (size_t)(total / 128));
    pgr_mm128 *res = (pgr_mm128 *)result_alloc(cap * sizeof(pgr_mm128));
    if (!res) { set_error("out of host memory"); return PGR_E_ARG; }
    size_t n_res = 0;
    std::vector<uint64_t> off;
    offsets[0] = 0;
    const int rc = pgr::run_chunked(ctx, n, rids, seqs, lens, *spec, padding, [&](size_t c0, size_t cn, size_t ns) -> int {
        if (n_res + ns > cap) {
            const size_t ncap = std::max(n_res + ns, cap * 2);
            pgr_mm128 *nres = (pgr_mm128 *)result_alloc(ncap * sizeof(pgr_mm128));
            if (!nres) { set_error("out of host memory"); return PGR_E_ARG; }
            memcpy(nres, res, n_res * sizeof(pgr_mm128));  // earlier chunks were synchronised below
            result_free(res);
            res = nres; cap = ncap;
        }
        off.resize(cn + 1);
        cudaError_t e = cudaSuccess;
        if (ns) e = cudaMemcpyAsync(res + n_res, ctx->d_result, ns * sizeof(pgr_mm128), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(off.data(), ctx->d_result_off, (cn + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
        // the next chunk reuses the device result buffers and `off`: wait for this chunk's copies
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("D2H failed: %s", cudaGetErrorString(e)); return PGR_E_CUDA; }
        for (size_t i = 0; i < cn; i++) offsets[c0 + i + 1] = n_res + (size_t)off[i + 1];
        n_res += ns;
        return PGR_OK;
    });
    if (rc != PGR_OK) { result_free(res); return rc; }
    *out = res;
    return PGR_OK;
}

int pgr_b200_sequence_to_shmmrs(uint32_t rid, const uint8_t *seq, size_t len, const pgr_shmmr_spec *spec, int padding,
                                pgr_mm128 **out, size_t *n_out) {
    if (!n_out) { set_error("NULL output argument"); return PGR_E_ARG; }
    size_t offs[2] = {0, 0};
    const uint8_t *ptrs[1] = {seq};
    const size_t lens[1] = {len};
    const uint32_t rids[1] = {rid};
    PGR_TRY(pgr_b200_shmmrs_batch(1, rids, ptrs, lens, spec, padding, out, offs));
    *n_out = offs[1];
    return PGR_OK;
}

}  // extern "C"
