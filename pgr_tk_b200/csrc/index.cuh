// index.cuh — pgr_b200_index: device-resident ShmmrFragMap.
#pragma once
#include <vector>

#include "ctx.cuh"
#include "index_kernels.cuh"
#include "shmmr_kernels.cuh"

struct pgr_b200_index {
    pgr_shmmr_spec spec;
    int mode = 0;                     // 0 = FASTX global fragment counter, 1 = AGC per-sequence pair ordinal
    pgr_b200_ctx *ctx = nullptr;      // owned: shimmer pipeline + streams
    pgr::DevBuf tuples;               // FragTuple[n_tuples], insertion order
    uint64_t n_tuples = 0;
    uint32_t n_frags = 0;             // == frags.len() of the reference's FASTX path (seq_db.rs:203)
    // CSR, valid when finalized
    pgr::DevBuf ukeys, offsets, sigs;
    uint64_t n_keys = 0;
    bool finalized = false;
    // staged batch (multi-GPU build: shimmers first, fragment-id base later)
    bool staged = false;
    uint64_t staged_t0 = 0;           // first tuple of the staged batch
    uint32_t staged_frags = 0, staged_prev_frags = 0;
    // multi-GPU build (shard.cu): tuples carry the insertion ordinal of their sequence (FragTuple::ord = ord_base + position in
    // the batch); the owner's sort uses it as the minor key when blocks of several batches interleave (ord_sort)
    uint32_t ord_base = 0, ord_in_batch = 0;
    bool gathered = false;            // CSR copied from the shards of a multi-GPU build: no tuples behind it, read-only
    bool from_mdb = false;            // read from an .mdb: no sequences behind it, appending is refused by the host mirror
    bool ord_sort = false;
    pgr::DevBuf sendbuf;              // tuples partitioned by destination shard
    // scratch
    pgr::DevBuf keysA, keysB, idxA, idxB, hist, head, block_sum, block_prefix, d_sid, d_pair_off, d_frg_base;
    pgr::DevBuf qtuples, q_hit_begin, q_hit_count, scratch0, scratch1, scratch2, scratch3;
    pgr::DevBuf sid_count, hitsA, hitsB, seg_keys, seg_off, chain_f, chain_u, chain_b, chain_seg, asm_prefix, asm_has, asm_out, asm_out2;
    cudaStream_t d2h_stream = nullptr;   // result download of one query group while the next is computed
    bool sid_count_valid = false;
    uint64_t launches = 0;
};

namespace pgr {
int index_reserve_tuples(pgr_b200_index *idx, uint64_t need);
int index_batch_tuples(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                       bool query_mode, DevBuf *qbuf, uint64_t *n_pairs_total, std::vector<uint64_t> *pairs_per_seq);
int index_sort(pgr_b200_index *idx, uint64_t n, int first_pass, int last_pass);
__global__ void iota_kernel(uint32_t *p, uint64_t n);
__global__ void set_u64_kernel(uint64_t *p, uint64_t v);
int write_mdb_file(const pgr_shmmr_spec &spec, uint64_t nk, const uint64_t *keys, const uint64_t *offs, const pgr_frag_sig *sigs, const char *path);
}  // namespace pgr
