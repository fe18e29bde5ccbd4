// pack_upload.cuh — packed host-to-device transport of sequence bytes (included by ctx.cu).
//
// A batch of bases normally crosses PCIe as ASCII, one byte per base, and PCIe is what bounds every host-buffer entry
// point (pgr_b200_shmmrs_batch: 5 GB at ~50 GB/s = 100 ms around 26 ms of kernels).  sequence_to_shmmrs only reads the
// class of a byte (shmmrutils.rs:426-436: A/a/0, C/c/1, G/g/2, T/t/3, anything else), i.e. 2 bits + 1 validity bit per base.
// The host packs those three bit planes with SIMD on all its cores (hostpack.cpp) straight out of the caller's buffers
// — pageable or page-locked alike — into a small ring of page-locked slots; a slot crosses PCIe as 8 bytes per 32 bases when
// every byte of it is a base (the validity plane stays at home) and as 12 otherwise, and unpack_kernel rewrites it as
// canonical ASCII ("ACGT", 'N') at its place in the device sequence store,
// so every kernel downstream is unchanged and sees a store that is equivalent for the reference's LUT.
//   PGR_B200_H2D=direct or pgr_b200_set_transport(PGR_TRANSPORT_DIRECT) disables the packed transport (A/B:
//   profiles/r2_e2e_packed_ab.txt)
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "hostpack.hpp"

namespace pgr {

constexpr uint32_t PACK_SLOT_BLOCKS = 1u << 20;   // 32-byte blocks per slot: 32 MiB of bases, 12 MiB packed
constexpr uint32_t PACK_PIECE_BLOCKS = 1u << 12;  // blocks per host work item (128 KiB of bases: 256 items per slot balance 16-32 threads)
constexpr int PACK_SLOTS = 4;
constexpr uint64_t PACK_MIN_BYTES = 4ull << 20;   // smaller uploads go the direct way

// one thread per block: three plane words -> 32 bytes (two 16-byte stores)
// has_v = 0: the slot holds bases only and its validity plane stayed on the host (2 bits per base crossed PCIe)
__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t *__restrict__ planes, uint32_t nb, uint8_t *__restrict__ dst, uint32_t has_v) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t p0 = planes[b], p1 = planes[nb + b], v = has_v ? planes[2 * nb + b] : 0xFFFFFFFFu;
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t a = (p0 >> (4 * j)) & 15u, c = (p1 >> (4 * j)) & 15u;
        // selector nibble of byte i = code of base 4j+i; pool bytes 0..3 = 'A','C','G','T'
        uint32_t sel = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) sel |= ((((a >> i) & 1u) | (((c >> i) & 1u) << 1))) << (4 * i);
        w[j] = __byte_perm(0x54474341u, 0u, sel);
    }
    if (v != 0xFFFFFFFFu) {   // rare: bytes that are no bases -> 'N'
#pragma unroll
        for (int j = 0; j < 8; j++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int bit = 4 * j + i;
                if (!((v >> bit) & 1u)) w[j] = (w[j] & ~(0xFFu << (8 * i))) | ((uint32_t)'N' << (8 * i));
            }
        }
    }
    uint4 *o = reinterpret_cast<uint4 *>(dst + (size_t)b * 32);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// page-locked + device slot ring of one device; kept in a process-wide free list because page-locking costs ~0.8 ms/MB
struct PackRing {
    int device = 0;
    uint32_t *h = nullptr;      // PACK_SLOTS x 3 x PACK_SLOT_BLOCKS words, page-locked
    uint32_t *d = nullptr;      // the same on the device
    cudaEvent_t h2d_done[PACK_SLOTS] = {}, unpack_done[PACK_SLOTS] = {};
    bool used[PACK_SLOTS] = {};
    cudaStream_t unpack_stream = nullptr;
    uint64_t seq = 0;
    // Hybrid feeding (page-locked sources): whenever the copies already queued will be over before the host has packed its next
    // slot, that slot is copied as it is instead — PCIe carries plain bytes in the gaps the packer leaves, the packer skips
    // the slot.  The backlog is estimated from the queued copies that have not completed yet.
    static constexpr int N_TRACK = 64;
    cudaEvent_t track_ev[N_TRACK] = {};
    struct Queued { int ev; float est_ms; };
    std::deque<Queued> queued;
    int track_next = 0;
    float pack_ms = 0.45f;          // running estimate of the host time to pack one full slot
    uint64_t n_direct_slots = 0, n_packed_slots = 0, bytes_sent = 0;   // since the ring was made (diagnostics)
    static constexpr size_t slot_words() { return 3ull * PACK_SLOT_BLOCKS; }
};
constexpr float PCIE_GB_PER_MS = 0.050f;   // ~50 GB/s: only used to size the backlog estimate
constexpr size_t HYBRID_MAX_SEGMENTS = 64;  // a slot made of more pieces than this (many short sequences) is always packed

inline std::mutex &pack_ring_mu() { static std::mutex m; return m; }
inline std::vector<PackRing *> &pack_ring_free() { static std::vector<PackRing *> v; return v; }

inline PackRing *pack_ring_acquire(int device) {
    {
        std::lock_guard<std::mutex> lk(pack_ring_mu());
        auto &fl = pack_ring_free();
        for (size_t i = 0; i < fl.size(); i++)
            if (fl[i]->device == device) { PackRing *r = fl[i]; fl.erase(fl.begin() + i); return r; }
    }
    PackRing *r = new PackRing();
    r->device = device;
    const size_t bytes = PACK_SLOTS * PackRing::slot_words() * 4;
    bool ok = cudaMallocHost((void **)&r->h, bytes) == cudaSuccess && cudaMalloc((void **)&r->d, bytes) == cudaSuccess;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    ok = ok && cudaStreamCreateWithPriority(&r->unpack_stream, cudaStreamNonBlocking, hi) == cudaSuccess;
    for (int s = 0; ok && s < PACK_SLOTS; s++)
        ok = cudaEventCreateWithFlags(&r->h2d_done[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&r->unpack_done[s], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < PackRing::N_TRACK; i++) ok = cudaEventCreateWithFlags(&r->track_ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        if (r->h) cudaFreeHost(r->h);
        if (r->d) cudaFree(r->d);
        delete r;
        set_error("cannot allocate the packed-upload ring");
        return nullptr;
    }
    return r;
}

inline void pack_ring_release(PackRing *r) {
    if (!r) return;
    cudaStreamSynchronize(r->unpack_stream);
    for (int s = 0; s < PACK_SLOTS; s++) r->used[s] = false;
    r->queued.clear();
    std::lock_guard<std::mutex> lk(pack_ring_mu());
    pack_ring_free().push_back(r);
}

// 0 = packed (default), 1 = direct; initialised from PGR_B200_H2D, changed by pgr_b200_set_transport
inline std::atomic<int> &transport_mode() {
    static std::atomic<int> m{[] { const char *e = getenv("PGR_B200_H2D"); return (e && std::string(e) == "direct") ? 1 : 0; }()};
    return m;
}
inline bool packed_upload_enabled() { return transport_mode().load(std::memory_order_relaxed) == 0; }

// is the batch's memory page-locked (judged by its first sequence of a megabyte or more)?
inline bool source_page_locked(const uint8_t *const *seqs, const size_t *lens, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (lens[i] < (1u << 20)) continue;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, seqs[i]) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    }
    return false;
}

// Which transport a batch takes when the mode is the default.  Packing pays when the host packs faster than PCIe copies
// (~5.6 GB/s per pool thread against ~50 GB/s for a page-locked source) or when the source is pageable, which the direct copy
// moves at ~10 GB/s through the driver's staging.  A page-locked source is fed in hybrid form (see PackRing): slots are copied
// as they are whenever the packer cannot keep the copy engine busy.  With few host threads per process (one process per GPU on a
// node with 24-32 CPUs) the node's aggregate host-to-device rate is the limit and packing only competes for memory bandwidth.
// Whole-job Gbases/s of the e2e leg, hybrid against direct, by ranks x pool threads: 1 x 16: 95-99 / 48-49; 2 x 12: 112 / 96;
// 4 x 8: 148 / 175; 8 x 4: 153 / 169 (packed only: 93) — so below PACK_MIN_THREADS a page-locked source is copied directly.
constexpr unsigned PACK_MIN_THREADS = 10;
inline std::atomic<uint64_t> &transport_bytes() { static std::atomic<uint64_t> b{0}; return b; }   // bytes the packed / hybrid transport put on PCIe (process-wide)
inline std::atomic<int> &last_transport() { static std::atomic<int> t{-1}; return t; }   // what the newest batch call took (0 packed, 1 direct)
inline bool choose_packed(const uint8_t *const *seqs, const size_t *lens, size_t n, uint64_t total_bases) {
    if (!packed_upload_enabled() || total_bases < PACK_MIN_BYTES) return false;
    static const unsigned min_threads = getenv("PGR_B200_PACK_MIN_THREADS") ? (unsigned)atoi(getenv("PGR_B200_PACK_MIN_THREADS")) : PACK_MIN_THREADS;   // tuning aid
    if (pool_threads() >= min_threads) return true;
    return !source_page_locked(seqs, lens, n);
}

// sequences [i0, i1) of the layout (h_off: 32-byte aligned consecutive store offsets, h_len) -> store, through the ring.
// Host work happens on the calling thread + the pool; copies on st_copy, expansion on the ring's own stream; on return
// everything is queued and st_copy waits for the last expansion (so an event recorded on st_copy covers the chunk).
inline int upload_packed(PackRing *ring, uint8_t *store, const std::vector<uint64_t> &h_off, const std::vector<uint32_t> &h_len,
                         const uint8_t *const *seqs, size_t i0, size_t i1, cudaStream_t st_copy, bool src_page_locked = false) {
    while (i0 < i1 && h_len[i0] == 0) i0++;
    while (i1 > i0 && h_len[i1 - 1] == 0) i1--;
    if (i0 >= i1) return PGR_OK;
    const uint64_t B0 = h_off[i0] >> 5;
    const uint64_t B1 = (h_off[i1 - 1] + (((uint64_t)h_len[i1 - 1] + 31) & ~31ull)) >> 5;
    int last_slot = -1;
    // copies queued on st_copy that are not over yet, as milliseconds of PCIe time
    auto backlog_ms = [&]() -> float {
        while (!ring->queued.empty() && cudaEventQuery(ring->track_ev[ring->queued.front().ev]) == cudaSuccess) ring->queued.pop_front();
        float t = 0;
        for (const auto &q : ring->queued) t += q.est_ms;
        return t;
    };
    auto track = [&](float est_ms) -> int {
        if ((int)ring->queued.size() >= PackRing::N_TRACK - 1) {   // keep the event ring ahead of the queue
            PGR_CUDA(cudaEventSynchronize(ring->track_ev[ring->queued.front().ev]));
            ring->queued.pop_front();
        }
        const int e = ring->track_next;
        ring->track_next = (e + 1) % PackRing::N_TRACK;
        PGR_CUDA(cudaEventRecord(ring->track_ev[e], st_copy));
        ring->queued.push_back({e, est_ms});
        return PGR_OK;
    };
    static const bool hybrid_off = getenv("PGR_B200_NO_HYBRID") != nullptr;   // A/B aid
    for (uint64_t sb = B0; sb < B1; sb += PACK_SLOT_BLOCKS) {
        const uint32_t nb = (uint32_t)std::min<uint64_t>(PACK_SLOT_BLOCKS, B1 - sb);
        if (src_page_locked && !hybrid_off && backlog_ms() < ring->pack_ms * ((float)nb / PACK_SLOT_BLOCKS)) {
            // the copy engine would run dry while this slot is packed: send the slot's bytes as they are
            size_t i = (size_t)(std::upper_bound(h_off.begin() + i0, h_off.begin() + i1, sb << 5) - h_off.begin()) - 1;
            // (count the pieces first: a slot of many short sequences is cheaper to pack than to copy piecewise)
            size_t n_seg = 0;
            { uint64_t b = sb; size_t ii = i; const uint64_t be = sb + nb;
              while (b < be && n_seg <= HYBRID_MAX_SEGMENTS) {
                  const uint64_t within = (b << 5) - h_off[ii], len = h_len[ii];
                  if (within >= len) { ii++; continue; }
                  const uint64_t take = std::min<uint64_t>(len - within, (be - b) << 5);
                  n_seg++; b += (take + 31) >> 5; if (within + take >= len) ii++;
              } }
            if (n_seg <= HYBRID_MAX_SEGMENTS) {
                uint64_t b = sb, bytes = 0;
                const uint64_t be = sb + nb;
                while (b < be) {
                    const uint64_t within = (b << 5) - h_off[i], len = h_len[i];
                    if (within >= len) { i++; continue; }
                    const uint64_t take = std::min<uint64_t>(len - within, (be - b) << 5);
                    PGR_CUDA(cudaMemcpyAsync(store + h_off[i] + within, seqs[i] + within, take, cudaMemcpyHostToDevice, st_copy));
                    bytes += take;
                    b += (take + 31) >> 5;
                    if (within + take >= len) i++;
                }
                PGR_TRY(track((float)bytes * 1e-9f / PCIE_GB_PER_MS));
                ring->n_direct_slots++;
                ring->bytes_sent += bytes;
                transport_bytes().fetch_add(bytes, std::memory_order_relaxed);
                continue;
            }
        }
        const auto t_pack0 = std::chrono::steady_clock::now();
        const int s = (int)(ring->seq++ % PACK_SLOTS);
        if (ring->used[s]) PGR_CUDA(cudaEventSynchronize(ring->h2d_done[s]));   // the copy that read this slot is over
        uint32_t *hp = ring->h + (size_t)s * PackRing::slot_words();
        uint32_t *p0 = hp, *p1 = hp + nb, *pv = hp + 2 * (size_t)nb;
        const size_t n_pieces = (nb + PACK_PIECE_BLOCKS - 1) / PACK_PIECE_BLOCKS;
        std::atomic<uint32_t> all_valid{0xFFFFFFFFu};
        parallel_for(n_pieces, [&](size_t pc) {
            uint32_t allv = 0xFFFFFFFFu;
            uint64_t b = sb + pc * PACK_PIECE_BLOCKS;
            const uint64_t be = std::min<uint64_t>(sb + nb, b + PACK_PIECE_BLOCKS);
            // sequence that holds block b: the last one of [i0, i1) whose offset is <= 32 b (empty sequences share the offset
            // of their successor, upper_bound steps over them)
            size_t i = (size_t)(std::upper_bound(h_off.begin() + i0, h_off.begin() + i1, b << 5) - h_off.begin()) - 1;
            while (b < be) {
                const uint64_t within = (b << 5) - h_off[i];
                const uint64_t len = h_len[i];
                if (within >= len) { i++; continue; }   // (only for empty sequences)
                const uint64_t take = std::min<uint64_t>(len - within, (be - b) << 5);
                allv &= pack_bases(seqs[i] + within, take, p0 + (b - sb), p1 + (b - sb), pv + (b - sb));
                b += (take + 31) >> 5;
                if (within + take >= len) i++;
            }
            if (allv != 0xFFFFFFFFu) all_valid.fetch_and(allv, std::memory_order_relaxed);
        });
        // a slot of bases only (the common case: gaps are few) leaves its validity plane at home: planes are laid out p0 | p1 | v
        const bool has_v = all_valid.load() != 0xFFFFFFFFu;
        const size_t slot_bytes = (size_t)nb * (has_v ? 12 : 8);
        if (nb == PACK_SLOT_BLOCKS) {
            const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_pack0).count();
            ring->pack_ms = 0.75f * ring->pack_ms + 0.25f * std::min(ms, 8.0f);
        }
        ring->n_packed_slots++;
        uint32_t *dp = ring->d + (size_t)s * PackRing::slot_words();
        if (ring->used[s]) PGR_CUDA(cudaStreamWaitEvent(st_copy, ring->unpack_done[s], 0));   // the device slot has been expanded
        PGR_CUDA(cudaMemcpyAsync(dp, hp, slot_bytes, cudaMemcpyHostToDevice, st_copy));
        PGR_CUDA(cudaEventRecord(ring->h2d_done[s], st_copy));
        PGR_TRY(track((float)slot_bytes * 1e-9f / PCIE_GB_PER_MS));
        ring->bytes_sent += slot_bytes;
        transport_bytes().fetch_add(slot_bytes, std::memory_order_relaxed);
        PGR_CUDA(cudaStreamWaitEvent(ring->unpack_stream, ring->h2d_done[s], 0));
        unpack_kernel<<<(nb + 255) / 256, 256, 0, ring->unpack_stream>>>(dp, nb, store + (sb << 5), has_v ? 1u : 0u);
        PGR_CUDA(cudaGetLastError());
        PGR_CUDA(cudaEventRecord(ring->unpack_done[s], ring->unpack_stream));
        ring->used[s] = true;
        last_slot = s;
    }
    if (last_slot >= 0) PGR_CUDA(cudaStreamWaitEvent(st_copy, ring->unpack_done[last_slot], 0));
    return PGR_OK;
}

// Raw bytes of a pageable source (fragment compression compares the caller's bytes as they are, so it cannot take the packed
// transport): the host pool copies them into the ring's page-locked slots, which then cross PCIe at its rate — a direct
// cudaMemcpyAsync from pageable memory moves ~10 GB/s through the driver's own staging.
inline int upload_raw_staged(PackRing *ring, uint8_t *dst_dev, const uint8_t *src, size_t bytes, cudaStream_t st) {
    const size_t slot_bytes = PackRing::slot_words() * 4, piece = 256u << 10;
    for (size_t o = 0; o < bytes; o += slot_bytes) {
        const size_t nbytes = std::min(slot_bytes, bytes - o);
        const int s = (int)(ring->seq++ % PACK_SLOTS);
        if (ring->used[s]) PGR_CUDA(cudaEventSynchronize(ring->h2d_done[s]));
        uint8_t *hp = reinterpret_cast<uint8_t *>(ring->h + (size_t)s * PackRing::slot_words());
        parallel_for((nbytes + piece - 1) / piece, [&](size_t i) { memcpy(hp + i * piece, src + o + i * piece, std::min(piece, nbytes - i * piece)); });
        PGR_CUDA(cudaMemcpyAsync(dst_dev + o, hp, nbytes, cudaMemcpyHostToDevice, st));
        PGR_CUDA(cudaEventRecord(ring->h2d_done[s], st));
        PGR_CUDA(cudaEventRecord(ring->unpack_done[s], st));   // the device slot of this index is not in use: keeps a later packed upload's wait trivial
        ring->used[s] = true;
    }
    return PGR_OK;
}

// The reverse for large one-off results: device -> ring slots at PCIe rate -> the host pool copies them into plain memory.
// Page-locking a fresh result buffer costs ~0.5-0.8 ms/MB, five times what the copy itself takes, and a copy into pageable memory
// runs at ~10 GB/s through the driver's staging; this keeps PCIe busy with neither.  Synchronous: the data is in dst on return.
inline int download_staged(PackRing *ring, uint8_t *dst, const uint8_t *src_dev, size_t bytes, cudaStream_t st) {
    const size_t slot_bytes = PackRing::slot_words() * 4, piece = 256u << 10;
    const size_t n = (bytes + slot_bytes - 1) / slot_bytes;
    PGR_CUDA(cudaStreamSynchronize(st));   // the ring's slots may still feed an upload queued on another stream: drain ours, then own the ring
    for (int s = 0; s < PACK_SLOTS; s++) if (ring->used[s]) { PGR_CUDA(cudaEventSynchronize(ring->h2d_done[s])); PGR_CUDA(cudaEventSynchronize(ring->unpack_done[s])); }
    auto issue = [&](size_t i) -> int {
        const int s = (int)(i % PACK_SLOTS);
        uint8_t *hp = reinterpret_cast<uint8_t *>(ring->h + (size_t)s * PackRing::slot_words());
        PGR_CUDA(cudaMemcpyAsync(hp, src_dev + i * slot_bytes, std::min(slot_bytes, bytes - i * slot_bytes), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaEventRecord(ring->h2d_done[s], st));
        return PGR_OK;
    };
    for (size_t i = 0; i < std::min<size_t>(n, PACK_SLOTS); i++) PGR_TRY(issue(i));
    for (size_t i = 0; i < n; i++) {
        const int s = (int)(i % PACK_SLOTS);
        PGR_CUDA(cudaEventSynchronize(ring->h2d_done[s]));
        const uint8_t *hp = reinterpret_cast<const uint8_t *>(ring->h + (size_t)s * PackRing::slot_words());
        const size_t nbytes = std::min(slot_bytes, bytes - i * slot_bytes);
        uint8_t *d = dst + i * slot_bytes;
        parallel_for((nbytes + piece - 1) / piece, [&](size_t q) { memcpy(d + q * piece, hp + q * piece, std::min(piece, nbytes - q * piece)); });
        if (i + PACK_SLOTS < n) PGR_TRY(issue(i + PACK_SLOTS));
    }
    for (int s = 0; s < PACK_SLOTS; s++) ring->used[s] = false;   // every slot is idle again
    return PGR_OK;
}

}  // namespace pgr
