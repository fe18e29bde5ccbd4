// ctx.cuh — device context: streams, grow-only device buffers, the device sequence store.
#pragma once
#include <functional>
#include <vector>

#include "common.cuh"

namespace pgr {

struct PackRing;

// grow-only device buffer, backed by the device's stream-ordered memory pool (cudaMallocAsync): the pool keeps freed
// blocks (release threshold = unlimited, set in pgr_b200_ctx_new), so contexts and indexes that come and go in one
// process do not pay cudaMalloc/cudaFree again
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return PGR_OK;
        release();
        size_t want = bytes + bytes / 8 + 256;
        PGR_CUDA(cudaMallocAsync(&p, want, (cudaStream_t)0));
        PGR_CUDA(cudaStreamSynchronize((cudaStream_t)0));
        cap = want;
        return PGR_OK;
    }
    void release() {
        if (!p) return;
        cudaDeviceSynchronize();   // growing/releasing is rare; nothing in flight may still use the old block
        cudaFreeAsync(p, (cudaStream_t)0);
        p = nullptr; cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct StageTimer {
    std::vector<const char *> names;
    std::vector<cudaEvent_t> ev0, ev1;
    std::vector<float> ms;
    std::vector<const char *> names_z;  // NULL-terminated copy handed to the C ABI
    size_t used = 0;
    void reset() { used = 0; }
    int begin(const char *name, cudaStream_t st);
    void end(int slot, cudaStream_t st);
    void collect();
    void destroy();
};

constexpr size_t SEQ_SLACK = 16384;  // readable bytes kept before the first and after the last sequence

}  // namespace pgr

struct pgr_b200_ctx {
    int device = 0;
    int n_sm = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    // sequence store
    pgr::DevBuf seq_store;              // owned copy (upload) — unused when sequences were adopted
    const uint8_t *d_seq = nullptr;     // base pointer used by kernels
    pgr::DevBuf d_off, d_len, d_rid;
    std::vector<uint64_t> h_off;
    std::vector<uint32_t> h_len, h_rid;
    size_t n_seq = 0;
    size_t r0 = 0, rn = 0;              // sequence range the pipeline currently works on (chunked batches)
    uint64_t total_bases = 0;
    // pinned staging for small sequences and control read-backs
    void *h_stage = nullptr; size_t h_stage_cap = 0;
    void *h_ctl = nullptr; size_t h_ctl_cap = 0;
    pgr::PackRing *pack = nullptr;      // packed host-to-device transport (pack_upload.cuh), acquired on first use
    // shimmer pipeline buffers
    pgr::DevBuf tile_prefix, cta_tile, arena, chunk_count, seq_count, seq_flag, replay_list, replay_count;
    pgr::DevBuf chunk_prefix, seq_fast, seq_dst, bufA, bufB, flags, block_sum, block_prefix, block_chunk, off_a, off_b, mark_bits, allinv_bits, n_skips;
    pgr::DevBuf patch_buf[11];          // scratch of apply_patches (cluster tables, slabs, splice metadata)
    uint64_t bits_dirty_lo = 0, bits_dirty_hi = 0;   // bitmap words an earlier launch may have set (cleared lazily)
    uint64_t chunk_cap = 0;
    // result of the last shmmrs call
    const pgr_mm128 *d_result = nullptr;
    const uint64_t *d_result_off = nullptr;
    size_t n_result = 0;
    bool result_valid = false;
    // level-0 list left in the arena chunks by run_l0 (no gather): read through a ChunkView by the next level
    bool l0_chunked = false;
    uint64_t l0_chunk_cap = 0;
    uint32_t l0_chunks = 0;
    // padding fix-up (rare): host-side rebuilt result
    pgr::DevBuf fix_mm, fix_off;
    pgr::StageTimer timer;
    uint64_t counters[8] = {0};

    int ensure_stage(size_t bytes);
    int ensure_ctl(size_t bytes);
};

namespace pgr {
int check_spec(const pgr_shmmr_spec *s);
int default_device();
pgr_b200_ctx *tls_ctx();
int shmmrs_range(pgr_b200_ctx *ctx, const pgr_shmmr_spec &spec, int padding, size_t *n_shmmrs);
int run_chunked(pgr_b200_ctx *ctx, size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens,
                const pgr_shmmr_spec &spec, int padding, const std::function<int(size_t, size_t, size_t)> &on_chunk);
}  // namespace pgr

#define PGR_TRY(x) do { int rc__ = (x); if (rc__ != PGR_OK) return rc__; } while (0)
