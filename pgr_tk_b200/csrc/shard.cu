// shard.cu — the multi-GPU ShmmrFragMap build inside the library.
//
//   every GPU: shimmers + tuples of its block of sequences (embarrassingly parallel, ctx.cu / index.cu)
//   -> all-gather of the fragment totals (global FASTX frg_id bases, seq_db.rs:203-231)
//   -> key splitters from an all-gathered sample of h0, sorted and cut on the device (identical on every rank)
//   -> stable partition of the 40-byte tuples into one key range per GPU
//   -> all-gather of the count matrix, ONE all-to-all (grouped ncclSend/ncclRecv over NVLink)
//   -> per owner: stable sort by (h0, h1) [+ insertion ordinal of the sequence when blocks interleave] and CSR
// Owner r holds the r-th key range, so the slices concatenate to the canonical key-ascending map: the .mdb written
// from N GPUs is byte-identical to the single-GPU one.
//
// Two process models share this code: one process per GPU (torchrun; pgr_b200_comm_init_rank with a broadcast NCCL
// unique id) and one process driving N GPUs with one host thread each (pgr_b200_mindex, ncclCommInitAll) — the form
// SURVEY §8b's `pgr_b200_index_new(spec, mode, n_gpus)` asks for and what `pgr-b200-make-frgdb --gpus N` uses.
// Shards that share a device (a test set-up on a 1-GPU box; NCCL refuses duplicate devices) exchange through
// device-to-device copies behind the same transport interface.
#include <nccl.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include "index.cuh"

using namespace pgr;

namespace pgr {

#define PGR_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) {                                                                   \
            ::pgr::set_error("%s failed: %s (%s:%d)", #call, ncclGetErrorString(r__), __FILE__, __LINE__); \
            return PGR_E_CUDA;                                                                      \
        }                                                                                           \
    } while (0)

// ---- transport ---------------------------------------------------------------------------------------------------
struct Transport {
    int rank = 0, n_ranks = 1;
    virtual ~Transport() {}
    // every rank contributes `bytes` at recv + rank*bytes (in place) and ends up with all contributions
    virtual int all_gather_inplace(void *recv, size_t bytes, cudaStream_t st) = 0;
    // send + send_off[p] .. send_off[p+1] goes to rank p; recv + recv_off[p] .. recv_off[p+1] comes from rank p (bytes)
    virtual int all_to_all(const uint8_t *send, const uint64_t *send_off, uint8_t *recv, const uint64_t *recv_off, cudaStream_t st) = 0;
    virtual const char *name() const = 0;
    // a shard of this process failed outside the transport: peers blocked in (or arriving at) a rendez-vous must not wait for it
    virtual void abort() {}
};

struct NcclTransport : Transport {
    ncclComm_t comm = nullptr;
    ~NcclTransport() override { if (comm) ncclCommDestroy(comm); }
    int all_gather_inplace(void *recv, size_t bytes, cudaStream_t st) override {
        PGR_NCCL(ncclAllGather((const uint8_t *)recv + (size_t)rank * bytes, recv, bytes, ncclUint8, comm, st));
        return PGR_OK;
    }
    int all_to_all(const uint8_t *send, const uint64_t *send_off, uint8_t *recv, const uint64_t *recv_off, cudaStream_t st) override {
        PGR_NCCL(ncclGroupStart());
        for (int p = 0; p < n_ranks; p++) {
            const uint64_t sb = send_off[p + 1] - send_off[p], rb = recv_off[p + 1] - recv_off[p];
            if (p == rank) continue;   // own part: device-to-device copy below
            if (sb) PGR_NCCL(ncclSend(send + send_off[p], sb, ncclUint8, p, comm, st));
            if (rb) PGR_NCCL(ncclRecv(recv + recv_off[p], rb, ncclUint8, p, comm, st));
        }
        PGR_NCCL(ncclGroupEnd());
        const uint64_t own = send_off[rank + 1] - send_off[rank];
        if (own) PGR_CUDA(cudaMemcpyAsync(recv + recv_off[rank], send + send_off[rank], own, cudaMemcpyDeviceToDevice, st));
        return PGR_OK;
    }
    const char *name() const override { return "nccl"; }
};

// shards of one process that cannot use NCCL (several shards on one device): peers publish their buffers, meet at a
// barrier and pull their parts with device-to-device copies
struct LocalBus {
    int n = 0;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    uint64_t generation = 0;
    std::vector<const uint8_t *> send;
    std::vector<const uint64_t *> send_off;
    std::vector<int> device;
    bool failed = false;              // set by abort(): a shard gave up, every barrier (current and later) returns false
    bool barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (failed) return false;
        const uint64_t g = generation;
        if (++waiting == n) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != g || failed; });
        return !failed;
    }
    void abort() {
        { std::lock_guard<std::mutex> lk(mu); failed = true; }
        cv.notify_all();
    }
};
#define PGR_BUS_BARRIER(bus) do { if (!(bus)->barrier()) { ::pgr::set_error("another shard of this process failed"); return PGR_E_CUDA; } } while (0)

struct LocalTransport : Transport {
    std::shared_ptr<LocalBus> bus;
    int device = 0;
    int all_gather_inplace(void *recv, size_t bytes, cudaStream_t st) override {
        PGR_CUDA(cudaStreamSynchronize(st));           // own contribution is complete
        bus->send[rank] = (const uint8_t *)recv;
        PGR_BUS_BARRIER(bus);
        for (int p = 0; p < n_ranks; p++) {
            if (p == rank) continue;
            PGR_CUDA(cudaMemcpyPeerAsync((uint8_t *)recv + (size_t)p * bytes, device, bus->send[p] + (size_t)p * bytes, bus->device[p], bytes, st));
        }
        PGR_CUDA(cudaStreamSynchronize(st));
        PGR_BUS_BARRIER(bus);                          // nobody reuses its buffer before every peer has read it
        return PGR_OK;
    }
    int all_to_all(const uint8_t *send, const uint64_t *send_off, uint8_t *recv, const uint64_t *recv_off, cudaStream_t st) override {
        PGR_CUDA(cudaStreamSynchronize(st));
        bus->send[rank] = send;
        bus->send_off[rank] = send_off;
        PGR_BUS_BARRIER(bus);
        for (int p = 0; p < n_ranks; p++) {
            const uint64_t rb = recv_off[p + 1] - recv_off[p];
            if (rb) PGR_CUDA(cudaMemcpyPeerAsync(recv + recv_off[p], device, bus->send[p] + bus->send_off[p][rank], bus->device[p], rb, st));
        }
        PGR_CUDA(cudaStreamSynchronize(st));
        PGR_BUS_BARRIER(bus);
        return PGR_OK;
    }
    const char *name() const override { return "local-d2d"; }
    void abort() override { bus->abort(); }
};

// ---- kernels of the merge -----------------------------------------------------------------------------------------
// evenly spaced sample of h0 (~0 where the shard has no tuple: not a valid 56-bit hash, sorts last)
__global__ void sample_h0_kernel(const FragTuple *t, uint64_t n, uint64_t *out, uint32_t n_sample) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sample) return;
    out[i] = n ? t[(uint64_t)(((unsigned __int128)i * n) / n_sample)].h0 : ~0ull;
}

// one CTA: bitonic sort of the gathered sample (padded to a power of two with ~0) in shared memory, then n_parts-1
// quantiles of the valid entries.  Minimizer hashes are skewed low, hence quantiles and not top bits.
__global__ void __launch_bounds__(1024) splitters_kernel(const uint64_t *sample, uint32_t n_sample, uint32_t n_pow2, uint32_t n_parts,
                                                          uint64_t *splitters) {
    extern __shared__ uint64_t s_sm[];
    __shared__ uint32_t n_valid;
    if (threadIdx.x == 0) n_valid = 0;
    for (uint32_t i = threadIdx.x; i < n_pow2; i += blockDim.x) s_sm[i] = i < n_sample ? sample[i] : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= n_pow2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const uint64_t a = s_sm[i], b = s_sm[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_sm[i] = b; s_sm[l] = a; }
                }
            }
            __syncthreads();
        }
    }
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x; i < n_pow2; i += blockDim.x) c += s_sm[i] != ~0ull;
    atomicAdd(&n_valid, c);
    __syncthreads();
    const uint32_t m = n_valid;
    for (uint32_t p = threadIdx.x; p + 1 < n_parts; p += blockDim.x)
        splitters[p] = m ? s_sm[min(m - 1, (uint32_t)(((uint64_t)m * (p + 1)) / n_parts))] : 0ull;
}

constexpr int PT_NT = 256, PT_WARPS = PT_NT / 32, PT_SEG = 2048, PT_MAX_PARTS = 64;

// destination part of a tuple: number of splitters <= h0
__device__ __forceinline__ uint32_t part_of(uint64_t h0, const uint64_t *sp, uint32_t n_split) {
    uint32_t lo = 0, hi = n_split;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sp[mid] <= h0) lo = mid + 1; else hi = mid; }
    return lo;
}

// per-warp-segment part histogram: hist[part * n_seg + seg]
__global__ void __launch_bounds__(PT_NT) part_hist_kernel(const FragTuple *t, uint64_t n, const uint64_t *splitters, uint32_t n_parts,
                                                           uint32_t *hist, uint32_t n_seg) {
    __shared__ uint64_t sp[PT_MAX_PARTS];
    __shared__ uint32_t h[PT_WARPS][PT_MAX_PARTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i + 1 < n_parts; i += blockDim.x) sp[i] = splitters[i];
    for (uint32_t i = lane; i < n_parts; i += 32) h[warp][i] = 0;
    __syncthreads();
    const uint32_t seg = blockIdx.x * PT_WARPS + warp;
    if (seg >= n_seg) return;
    const uint64_t b = (uint64_t)seg * PT_SEG, e = min(n, b + PT_SEG);
    for (uint64_t i = b + lane; i < e; i += 32) atomicAdd(&h[warp][part_of(t[i].h0, sp, n_parts - 1)], 1u);
    __syncwarp();
    for (uint32_t i = lane; i < n_parts; i += 32) hist[(uint64_t)i * n_seg + seg] = h[warp][i];
}

// exclusive scan of hist in (part-major, segment-minor) order, in place, and the per-part totals
__global__ void __launch_bounds__(1024) part_scan_kernel(uint32_t *hist, uint32_t n_parts, uint32_t n_seg, uint64_t n, uint64_t *counts) {
    __shared__ uint64_t wtot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_items = (uint64_t)n_parts * n_seg;
    const uint64_t per = (n_items + 1023) / 1024;
    const uint64_t b = min(n_items, (uint64_t)threadIdx.x * per), e = min(n_items, b + per);
    uint64_t sum = 0;
    for (uint64_t i = b; i < e; i++) sum += hist[i];
    uint64_t incl = sum;
    for (int d = 1; d < 32; d <<= 1) { const uint64_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint64_t v = wtot[lane];
        uint64_t wi = v;
        for (int d = 1; d < 32; d <<= 1) { const uint64_t x = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += x; }
        wtot[lane] = wi - v;
    }
    __syncthreads();
    uint64_t run = wtot[warp] + incl - sum;
    for (uint64_t i = b; i < e; i++) { const uint32_t v = hist[i]; hist[i] = (uint32_t)run; run += v; }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < n_parts; p += 1024) {
        const uint64_t start = hist[(uint64_t)p * n_seg];
        const uint64_t next = (p + 1 < n_parts) ? hist[(uint64_t)(p + 1) * n_seg] : n;
        counts[p] = next - start;
    }
}

// stable scatter of whole tuples: a warp walks its segment in order, lanes of the same part are ranked by lane id
__global__ void __launch_bounds__(PT_NT) part_scatter_kernel(const FragTuple *t, uint64_t n, const uint64_t *splitters, uint32_t n_parts,
                                                              const uint32_t *hist, uint32_t n_seg, FragTuple *out) {
    __shared__ uint64_t sp[PT_MAX_PARTS];
    __shared__ uint32_t off[PT_WARPS][PT_MAX_PARTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t seg = blockIdx.x * PT_WARPS + warp;
    for (uint32_t i = threadIdx.x; i + 1 < n_parts; i += blockDim.x) sp[i] = splitters[i];
    if (seg < n_seg) for (uint32_t i = lane; i < n_parts; i += 32) off[warp][i] = hist[(uint64_t)i * n_seg + seg];
    __syncthreads();
    if (seg >= n_seg) return;
    const uint64_t b = (uint64_t)seg * PT_SEG, e = min(n, b + PT_SEG);
    for (uint64_t i0 = b; i0 < e; i0 += 32) {
        const uint64_t i = i0 + lane;
        const bool live = i < e;
        FragTuple x;
        uint32_t d = 0xFFFFFFFFu;
        if (live) { x = t[i]; d = part_of(x.h0, sp, n_parts - 1); }
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        if (live) out[off[warp][d] + __popc(peers & ((1u << lane) - 1))] = x;
        __syncwarp();
        if (live && (peers & ((1u << lane) - 1)) == 0) off[warp][d] += __popc(peers);
        __syncwarp();
    }
}

__global__ void add_u64_kernel(uint64_t *p, uint64_t n, uint64_t v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += v;
}

}  // namespace pgr

// ---- handles ------------------------------------------------------------------------------------------------------
struct pgr_b200_comm {
    std::unique_ptr<pgr::Transport> tp;
    int device = 0;
};

namespace {

struct Ev {
    cudaEvent_t e = nullptr;
    Ev() { cudaEventCreate(&e); }
    ~Ev() { if (e) cudaEventDestroy(e); }
};
float ev_ms(const Ev &a, const Ev &b) { float ms = 0; cudaEventElapsedTime(&ms, a.e, b.e); return ms; }

// the exchange step: local tuples (insertion order) -> this rank's key range, sorted, CSR built
int shard_merge(pgr_b200_index *idx, pgr_b200_comm *c, bool ord_sort, pgr_shard_stats *stats) {
    pgr_b200_ctx *ctx = idx->ctx;
    PGR_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Transport *tp = c->tp.get();
    const int W = tp->n_ranks, R = tp->rank;
    if (idx->staged) { set_error("a staged batch is pending"); return PGR_E_ARG; }
    const uint64_t n = idx->n_tuples;
    pgr_shard_stats s;
    memset(&s, 0, sizeof s);
    s.n_ranks = (uint32_t)W; s.rank = (uint32_t)R;
    s.n_tuples_local = n;
    Ev e0, e1, e2, e3;
    cudaEventRecord(e0.e, st);
    if (W == 1) {
        idx->ord_sort = ord_sort;
        cudaEventRecord(e1.e, st); cudaEventRecord(e2.e, st);
        PGR_TRY(pgr_b200_index_finalize(idx));
        cudaEventRecord(e3.e, st);
        PGR_CUDA(cudaStreamSynchronize(st));
        s.n_tuples_owned = n; s.sort_ms = ev_ms(e2, e3);
        if (stats) *stats = s;
        return PGR_OK;
    }
    if (W > PT_MAX_PARTS) { set_error("at most %d shards", PT_MAX_PARTS); return PGR_E_ARG; }
    // 1. splitters: sample -> all-gather -> sort + quantiles on the device (every rank computes the same ones)
    uint32_t w2 = 1;
    while (w2 < (uint32_t)W) w2 <<= 1;
    const uint32_t S = std::min<uint32_t>(2048, 16384 / w2);
    const uint32_t n_all = S * (uint32_t)W;
    uint32_t n_pow2 = 1;
    while (n_pow2 < n_all) n_pow2 <<= 1;
    PGR_TRY(idx->scratch1.ensure((size_t)n_all * 8 + PT_MAX_PARTS * 8 + (size_t)W * W * 8 + 64));
    uint64_t *d_sample = idx->scratch1.as<uint64_t>();
    uint64_t *d_split = d_sample + n_all;
    uint64_t *d_matrix = d_split + PT_MAX_PARTS;          // [W][W]: row r = counts rank r sends to each part
    sample_h0_kernel<<<ceil_div<uint32_t>(S, 256), 256, 0, st>>>(idx->tuples.as<FragTuple>(), n, d_sample + (size_t)R * S, S);
    PGR_CUDA(cudaGetLastError());
    PGR_TRY(tp->all_gather_inplace(d_sample, (size_t)S * 8, st));
    PGR_CUDA(cudaFuncSetAttribute(splitters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    splitters_kernel<<<1, 1024, (size_t)n_pow2 * 8, st>>>(d_sample, n_all, n_pow2, (uint32_t)W, d_split);
    PGR_CUDA(cudaGetLastError());
    // 2. stable partition into W key ranges
    const uint32_t n_seg = (uint32_t)std::max<uint64_t>(1, ceil_div<uint64_t>(n, PT_SEG));
    const uint32_t grid = ceil_div<uint32_t>(n_seg, PT_WARPS);
    PGR_TRY(idx->hist.ensure((size_t)n_seg * W * sizeof(uint32_t) + 64));
    PGR_TRY(idx->sendbuf.ensure(std::max<uint64_t>(1, n) * sizeof(FragTuple)));
    part_hist_kernel<<<grid, PT_NT, 0, st>>>(idx->tuples.as<FragTuple>(), n, d_split, (uint32_t)W, idx->hist.as<uint32_t>(), n_seg);
    part_scan_kernel<<<1, 1024, 0, st>>>(idx->hist.as<uint32_t>(), (uint32_t)W, n_seg, n, d_matrix + (size_t)R * W);
    part_scatter_kernel<<<grid, PT_NT, 0, st>>>(idx->tuples.as<FragTuple>(), n, d_split, (uint32_t)W, idx->hist.as<uint32_t>(), n_seg,
                                                idx->sendbuf.as<FragTuple>());
    PGR_CUDA(cudaGetLastError());
    idx->launches += 5;
    cudaEventRecord(e1.e, st);
    // 3. count matrix -> host (sizes of the receive buffer and of the sends are host-side arguments)
    PGR_TRY(tp->all_gather_inplace(d_matrix, (size_t)W * 8, st));
    PGR_TRY(ctx->ensure_ctl((size_t)W * W * 8 + 64));
    uint64_t *h_matrix = (uint64_t *)ctx->h_ctl;
    PGR_CUDA(cudaMemcpyAsync(h_matrix, d_matrix, (size_t)W * W * 8, cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    std::vector<uint64_t> send_off(W + 1, 0), recv_off(W + 1, 0);
    for (int p = 0; p < W; p++) {
        send_off[p + 1] = send_off[p] + h_matrix[(size_t)R * W + p] * sizeof(FragTuple);
        recv_off[p + 1] = recv_off[p] + h_matrix[(size_t)p * W + R] * sizeof(FragTuple);
    }
    if (send_off[W] != n * sizeof(FragTuple)) { set_error("partition counts do not add up"); return PGR_E_CUDA; }
    const uint64_t n_recv = recv_off[W] / sizeof(FragTuple);
    if (n_recv >= 0xFFFFFFF0ull) { set_error("more than 2^32 tuples on one device"); return PGR_E_LIMIT; }
    // 4. the all-to-all; the received tuples become this shard's tuples (sources in rank order, each in its own order)
    PGR_TRY(idx->tuples.ensure(std::max<uint64_t>(1, n_recv) * sizeof(FragTuple)));   // old contents live on in sendbuf
    PGR_TRY(tp->all_to_all(idx->sendbuf.as<uint8_t>(), send_off.data(), idx->tuples.as<uint8_t>(), recv_off.data(), st));
    cudaEventRecord(e2.e, st);
    idx->n_tuples = n_recv;
    idx->finalized = false;
    idx->ord_sort = ord_sort;
    // 5. owner: stable sort + CSR
    PGR_TRY(pgr_b200_index_finalize(idx));
    cudaEventRecord(e3.e, st);
    PGR_CUDA(cudaStreamSynchronize(st));
    s.n_tuples_owned = n_recv;
    s.n_tuples_sent = n - h_matrix[(size_t)R * W + R];
    s.bytes_sent = s.n_tuples_sent * sizeof(FragTuple);
    s.bytes_recv = (n_recv - h_matrix[(size_t)R * W + R]) * sizeof(FragTuple);
    s.partition_ms = ev_ms(e0, e1); s.exchange_ms = ev_ms(e1, e2); s.sort_ms = ev_ms(e2, e3);
    if (stats) *stats = s;
    return PGR_OK;
}

// fragment-id base of this rank = fragments consumed by the lower ranks (FASTX numbering is one global running counter)
int exchange_frag_totals(pgr_b200_index *idx, pgr_b200_comm *c, uint64_t n_frags_local, uint64_t *base, uint64_t *total) {
    Transport *tp = c->tp.get();
    const int W = tp->n_ranks, R = tp->rank;
    pgr_b200_ctx *ctx = idx->ctx;
    cudaStream_t st = ctx->stream;
    PGR_TRY(idx->scratch1.ensure((size_t)W * 8 + 64));
    PGR_TRY(ctx->ensure_ctl((size_t)W * 8 + 64));
    uint64_t *h = (uint64_t *)ctx->h_ctl;
    h[R] = n_frags_local;
    PGR_CUDA(cudaMemcpyAsync(idx->scratch1.as<uint64_t>() + R, h + R, 8, cudaMemcpyHostToDevice, st));
    PGR_TRY(tp->all_gather_inplace(idx->scratch1.p, 8, st));
    PGR_CUDA(cudaMemcpyAsync(h, idx->scratch1.p, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    PGR_CUDA(cudaStreamSynchronize(st));
    uint64_t b = 0, t = 0;
    for (int p = 0; p < W; p++) { if (p < R) b += h[p]; t += h[p]; }
    if (t >= 0xFFFFFFFFull) { set_error("more than 2^32 fragments"); return PGR_E_LIMIT; }
    *base = b; *total = t;
    return PGR_OK;
}

}  // namespace

extern "C" {

int pgr_b200_comm_unique_id(uint8_t id[PGR_B200_COMM_ID_BYTES]) {
    if (!id) { set_error("id is NULL"); return PGR_E_ARG; }
    static_assert(sizeof(ncclUniqueId) <= PGR_B200_COMM_ID_BYTES, "ncclUniqueId larger than the ABI's id buffer");
    ncclUniqueId u;
    PGR_NCCL(ncclGetUniqueId(&u));
    memset(id, 0, PGR_B200_COMM_ID_BYTES);
    memcpy(id, &u, sizeof u);
    return PGR_OK;
}

pgr_b200_comm *pgr_b200_comm_init_rank(const uint8_t id[PGR_B200_COMM_ID_BYTES], int rank, int n_ranks, int device) {
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("bad argument"); return nullptr; }
    if (device < 0) device = default_device();
    if (device >= pgr_b200_device_count()) { set_error("no CUDA device %d: libpgr_b200 has no CPU fallback", device); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return nullptr; }
    auto tp = std::make_unique<NcclTransport>();
    tp->rank = rank; tp->n_ranks = n_ranks;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    const ncclResult_t r = ncclCommInitRank(&tp->comm, n_ranks, u, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", ncclGetErrorString(r)); tp->comm = nullptr; return nullptr; }
    pgr_b200_comm *c = new pgr_b200_comm();
    c->device = device;
    c->tp = std::move(tp);
    return c;
}

void pgr_b200_comm_free(pgr_b200_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    delete c;
}

int pgr_b200_index_merge(pgr_b200_index *idx, pgr_b200_comm *comm, pgr_shard_stats *stats) {
    if (!idx || !comm) { set_error("NULL argument"); return PGR_E_ARG; }
    return shard_merge(idx, comm, false, stats);
}

int pgr_b200_index_build_sharded(pgr_b200_index *idx, pgr_b200_comm *comm, size_t n, const uint32_t *sids, const uint8_t *const *seqs,
                                 const size_t *lens, pgr_shard_stats *stats) {
    if (!idx || !comm) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    Ev e0, e1;
    cudaEventRecord(e0.e, idx->ctx->stream);
    uint64_t nf = 0, base = 0, total = 0;
    PGR_TRY(pgr_b200_index_stage_batch(idx, n, sids, seqs, lens, &nf));
    PGR_TRY(exchange_frag_totals(idx, comm, nf, &base, &total));
    PGR_TRY(pgr_b200_index_commit_batch(idx, (uint32_t)base));
    cudaEventRecord(e1.e, idx->ctx->stream);
    pgr_shard_stats s;
    PGR_TRY(shard_merge(idx, comm, false, &s));
    s.stage_ms = ev_ms(e0, e1);
    s.total_frags = total;
    if (stats) *stats = s;
    return PGR_OK;
}

int pgr_b200_index_build_sharded_device(pgr_b200_index *idx, pgr_b200_comm *comm, const uint8_t *dev_base, size_t n, const uint32_t *sids,
                                        const uint64_t *offs, const uint64_t *lens, pgr_shard_stats *stats) {
    if (!idx || !comm) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    Ev e0, e1;
    cudaEventRecord(e0.e, idx->ctx->stream);
    uint64_t nf = 0, base = 0, total = 0;
    PGR_TRY(pgr_b200_index_stage_device(idx, dev_base, n, sids, offs, lens, &nf));
    PGR_TRY(exchange_frag_totals(idx, comm, nf, &base, &total));
    PGR_TRY(pgr_b200_index_commit_batch(idx, (uint32_t)base));
    cudaEventRecord(e1.e, idx->ctx->stream);
    pgr_shard_stats s;
    PGR_TRY(shard_merge(idx, comm, false, &s));
    s.stage_ms = ev_ms(e0, e1);
    s.total_frags = total;
    if (stats) *stats = s;
    return PGR_OK;
}

}  // extern "C"

// ---- one process, N GPUs ------------------------------------------------------------------------------------------
struct pgr_b200_mindex {
    pgr_shmmr_spec spec;
    int mode = 0;
    int n = 0;
    std::vector<pgr_b200_index *> shard;
    std::vector<pgr_b200_comm *> comm;
    uint64_t n_frags = 0;      // global running fragment counter (FASTX numbering)
    uint32_t n_seqs = 0;       // sequences added so far = insertion ordinal of the next one
    int n_batches = 0;
    bool finalized = false;
    std::vector<pgr_shard_stats> stats;
};

namespace {

// run f(g) on one host thread per shard; first failure wins, its message becomes the caller's last error
template <class F>
int for_each_shard(pgr_b200_mindex *m, F f) {
    std::vector<int> rc(m->n, PGR_OK);
    std::vector<std::string> err(m->n);
    std::vector<std::thread> th;
    for (int g = 0; g < m->n; g++)
        th.emplace_back([&, g] {
            rc[g] = f(g);
            if (rc[g] != PGR_OK) {
                err[g] = get_error();
                if (g < (int)m->comm.size() && m->comm[g]) m->comm[g]->tp->abort();   // peers waiting at a rendez-vous give up too
            }
        });
    for (auto &t : th) t.join();
    int first = -1;
    for (int g = 0; g < m->n; g++) {   // report the shard that failed on its own, not one that gave up at a rendez-vous because of it
        if (rc[g] == PGR_OK) continue;
        if (first < 0) first = g;
        if (err[g].rfind("another shard", 0) != 0) { first = g; break; }
    }
    if (first >= 0) { set_error("shard %d: %s", first, err[first].c_str()); return rc[first]; }
    return PGR_OK;
}

}  // namespace

extern "C" {

pgr_b200_mindex *pgr_b200_mindex_new_devices(const pgr_shmmr_spec *spec, int frg_id_mode, int n_shards, const int *devices) {
    if (check_spec(spec) != PGR_OK) return nullptr;
    if (n_shards < 1 || n_shards > PT_MAX_PARTS || !devices) { set_error("n_shards must be in 1..%d", PT_MAX_PARTS); return nullptr; }
    const int n_dev = pgr_b200_device_count();
    if (n_dev <= 0) { set_error("no CUDA device available: libpgr_b200 has no CPU fallback"); return nullptr; }
    bool distinct = true;
    for (int i = 0; i < n_shards; i++) {
        if (devices[i] < 0 || devices[i] >= n_dev) { set_error("device %d out of range (have %d)", devices[i], n_dev); return nullptr; }
        for (int j = 0; j < i; j++) distinct = distinct && devices[i] != devices[j];
    }
    auto m = std::make_unique<pgr_b200_mindex>();
    m->spec = *spec; m->mode = frg_id_mode; m->n = n_shards;
    auto fail = [&]() -> pgr_b200_mindex * {
        for (auto c : m->comm) pgr_b200_comm_free(c);
        for (auto s : m->shard) pgr_b200_index_free(s);
        return nullptr;
    };
    for (int g = 0; g < n_shards; g++) {
        pgr_b200_index *s = pgr_b200_index_new(spec, frg_id_mode, devices[g]);
        if (!s) return fail();
        m->shard.push_back(s);
    }
    if (distinct && n_shards > 1) {
        std::vector<ncclComm_t> comms(n_shards);
        const ncclResult_t r = ncclCommInitAll(comms.data(), n_shards, devices);
        if (r != ncclSuccess) { set_error("ncclCommInitAll failed: %s", ncclGetErrorString(r)); return fail(); }
        for (int g = 0; g < n_shards; g++) {
            auto tp = std::make_unique<NcclTransport>();
            tp->rank = g; tp->n_ranks = n_shards; tp->comm = comms[g];
            pgr_b200_comm *c = new pgr_b200_comm();
            c->device = devices[g]; c->tp = std::move(tp);
            m->comm.push_back(c);
        }
    } else {
        auto bus = std::make_shared<LocalBus>();
        bus->n = n_shards;
        bus->send.assign(n_shards, nullptr); bus->send_off.assign(n_shards, nullptr);
        bus->device.assign(devices, devices + n_shards);
        for (int g = 0; g < n_shards; g++) {
            auto tp = std::make_unique<LocalTransport>();
            tp->rank = g; tp->n_ranks = n_shards; tp->bus = bus; tp->device = devices[g];
            pgr_b200_comm *c = new pgr_b200_comm();
            c->device = devices[g]; c->tp = std::move(tp);
            m->comm.push_back(c);
        }
    }
    return m.release();
}

pgr_b200_mindex *pgr_b200_mindex_new(const pgr_shmmr_spec *spec, int frg_id_mode, int n_gpus) {
    if (n_gpus < 1 || n_gpus > pgr_b200_device_count()) {
        set_error("n_gpus = %d but %d CUDA device(s) are visible (no CPU fallback)", n_gpus, pgr_b200_device_count());
        return nullptr;
    }
    std::vector<int> dev(n_gpus);
    for (int g = 0; g < n_gpus; g++) dev[g] = g;
    return pgr_b200_mindex_new_devices(spec, frg_id_mode, n_gpus, dev.data());
}

void pgr_b200_mindex_free(pgr_b200_mindex *m) {
    if (!m) return;
    for (auto c : m->comm) pgr_b200_comm_free(c);
    for (auto s : m->shard) pgr_b200_index_free(s);
    delete m;
}

int pgr_b200_mindex_n_shards(const pgr_b200_mindex *m) { return m ? m->n : 0; }

pgr_b200_index *pgr_b200_mindex_shard(pgr_b200_mindex *m, int shard) {
    if (!m || shard < 0 || shard >= m->n) { set_error("shard out of range"); return nullptr; }
    return m->shard[shard];
}

// the batch is cut into n contiguous blocks of about equal bases (block g on shard g); every shard computes the
// shimmers and tuples of its block, then the fragment-id bases follow from the per-block totals in block order
int pgr_b200_mindex_add_batch(pgr_b200_mindex *m, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens) {
    if (!m || (n && (!sids || !seqs || !lens))) { set_error("NULL argument"); return PGR_E_ARG; }
    if (m->finalized) { set_error("the sharded index is finalized"); return PGR_E_ARG; }
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) total += lens[i];
    std::vector<size_t> cut(m->n + 1, n);
    cut[0] = 0;
    {
        uint64_t acc = 0;
        int g = 1;
        for (size_t i = 0; i < n && g < m->n; i++) {
            acc += lens[i];
            while (g < m->n && acc * (uint64_t)m->n >= total * (uint64_t)g) cut[g++] = i + 1;
        }
    }
    std::vector<uint64_t> nf(m->n, 0);
    const uint32_t ord0 = m->n_seqs;
    PGR_TRY(for_each_shard(m, [&](int g) -> int {
        const size_t b = cut[g], cn = cut[g + 1] - cut[g];
        m->shard[g]->ord_base = ord0 + (uint32_t)b;
        return pgr_b200_index_stage_batch(m->shard[g], cn, sids + b, seqs + b, lens + b, &nf[g]);
    }));
    uint64_t base = m->n_frags;
    for (int g = 0; g < m->n; g++) {
        if (base + nf[g] >= 0xFFFFFFFFull) { set_error("more than 2^32 fragments"); return PGR_E_LIMIT; }
        PGR_TRY(pgr_b200_index_commit_batch(m->shard[g], (uint32_t)base));
        base += nf[g];
    }
    m->n_frags = base;
    m->n_seqs += (uint32_t)n;
    m->n_batches += 1;
    return PGR_OK;
}

int pgr_b200_mindex_finalize(pgr_b200_mindex *m) {
    if (!m) { set_error("NULL argument"); return PGR_E_ARG; }
    if (m->finalized) return PGR_OK;
    m->stats.assign(m->n, pgr_shard_stats());
    const bool ord_sort = m->n_batches > 1;   // one batch: the blocks arrive at their owner in rank = insertion order
    PGR_TRY(for_each_shard(m, [&](int g) -> int { return shard_merge(m->shard[g], m->comm[g], ord_sort, &m->stats[g]); }));
    for (int g = 0; g < m->n; g++) m->stats[g].total_frags = m->n_frags;
    m->finalized = true;
    return PGR_OK;
}

int pgr_b200_mindex_stats(pgr_b200_mindex *m, int shard, pgr_shard_stats *out) {
    if (!m || !out || shard < 0 || shard >= m->n || !m->finalized) { set_error("bad argument / not finalized"); return PGR_E_ARG; }
    *out = m->stats[shard];
    return PGR_OK;
}

int pgr_b200_mindex_counts(pgr_b200_mindex *m, size_t *n_keys, size_t *n_sigs, uint32_t *n_frags) {
    if (!m) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_TRY(pgr_b200_mindex_finalize(m));
    size_t nk = 0, ns = 0;
    for (int g = 0; g < m->n; g++) {
        size_t k = 0, s = 0;
        PGR_TRY(pgr_b200_index_counts(m->shard[g], &k, &s, nullptr));
        nk += k; ns += s;
    }
    if (n_keys) *n_keys = nk;
    if (n_sigs) *n_sigs = ns;
    if (n_frags) *n_frags = (uint32_t)m->n_frags;
    return PGR_OK;
}

// slices in shard order = keys ascending
int pgr_b200_mindex_export_csr(pgr_b200_mindex *m, uint64_t *keys, uint64_t *offsets, pgr_frag_sig *sigs) {
    if (!m || !keys || !offsets || !sigs) { set_error("NULL argument"); return PGR_E_ARG; }
    PGR_TRY(pgr_b200_mindex_finalize(m));
    size_t k0 = 0, s0 = 0;
    offsets[0] = 0;
    for (int g = 0; g < m->n; g++) {
        size_t nk = 0, ns = 0;
        PGR_TRY(pgr_b200_index_counts(m->shard[g], &nk, &ns, nullptr));
        std::vector<uint64_t> off(nk + 1);
        uint64_t dk[2]; pgr_frag_sig ds;
        PGR_TRY(pgr_b200_index_export_csr(m->shard[g], nk ? keys + 2 * k0 : dk, off.data(), ns ? sigs + s0 : &ds));
        for (size_t i = 0; i <= nk; i++) offsets[k0 + i] = s0 + off[i];
        k0 += nk; s0 += ns;
    }
    return PGR_OK;
}

pgr_b200_index *pgr_b200_mindex_gather(pgr_b200_mindex *m, int device) {
    if (!m) { set_error("NULL argument"); return nullptr; }
    size_t nk = 0, ns = 0;
    if (pgr_b200_mindex_counts(m, &nk, &ns, nullptr) != PGR_OK) return nullptr;
    if (ns >= 0xFFFFFFF0ull) { set_error("more than 2^32 signatures on one device"); return nullptr; }
    pgr_b200_index *g = pgr_b200_index_new(&m->spec, m->mode, device);
    if (!g) return nullptr;
    auto fail = [&]() -> pgr_b200_index * { pgr_b200_index_free(g); return nullptr; };
    pgr_b200_ctx *ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    if (g->ukeys.ensure(std::max<size_t>(1, nk) * sizeof(SortKey)) != PGR_OK || g->offsets.ensure((nk + 2) * sizeof(uint64_t)) != PGR_OK ||
        g->sigs.ensure(std::max<size_t>(1, ns) * sizeof(pgr_frag_sig)) != PGR_OK) return fail();
    size_t k0 = 0, s0 = 0;
    for (int r = 0; r < m->n; r++) {
        pgr_b200_index *sh = m->shard[r];
        const size_t k = sh->n_keys, s = sh->n_tuples;
        cudaError_t e = cudaSuccess;
        if (k) e = cudaMemcpyPeerAsync(g->ukeys.as<SortKey>() + k0, ctx->device, sh->ukeys.p, sh->ctx->device, k * sizeof(SortKey), st);
        if (e == cudaSuccess && k) e = cudaMemcpyPeerAsync(g->offsets.as<uint64_t>() + k0, ctx->device, sh->offsets.p, sh->ctx->device, k * sizeof(uint64_t), st);
        if (e == cudaSuccess && s) e = cudaMemcpyPeerAsync(g->sigs.as<pgr_frag_sig>() + s0, ctx->device, sh->sigs.p, sh->ctx->device, s * sizeof(pgr_frag_sig), st);
        if (e != cudaSuccess) { set_error("peer copy of shard %d failed: %s", r, cudaGetErrorString(e)); return fail(); }
        if (k && s0) add_u64_kernel<<<(uint32_t)ceil_div<uint64_t>(k, 256), 256, 0, st>>>(g->offsets.as<uint64_t>() + k0, k, s0);
        k0 += k; s0 += s;
    }
    set_u64_kernel<<<1, 1, 0, st>>>(g->offsets.as<uint64_t>() + nk, ns);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { set_error("gather failed"); return fail(); }
    g->n_keys = nk; g->n_tuples = ns; g->n_frags = (uint32_t)m->n_frags; g->finalized = true; g->gathered = true;
    return g;
}

int pgr_b200_mindex_write_mdb(pgr_b200_mindex *m, const char *path) {
    if (!m || !path) { set_error("NULL argument"); return PGR_E_ARG; }
    size_t nk = 0, ns = 0;
    PGR_TRY(pgr_b200_mindex_counts(m, &nk, &ns, nullptr));
    std::vector<uint64_t> keys(2 * std::max<size_t>(1, nk)), offs(nk + 1);
    std::vector<pgr_frag_sig> sigs(std::max<size_t>(1, ns));
    PGR_TRY(pgr_b200_mindex_export_csr(m, keys.data(), offs.data(), sigs.data()));
    return pgr::write_mdb_file(m->spec, nk, keys.data(), offs.data(), sigs.data(), path);
}

}  // extern "C"
