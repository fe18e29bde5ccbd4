/*
 * pgr_b200.h — C ABI of libpgr_b200.so: the B200 (sm_100a) implementation of pgr-tk's SHIMMER indexing hot path.
 *
 * Each entry point names the reference interface (GeneDx/pgr-tk, paths relative to the reference root) it replaces.
 * Conventions (modelled on the reference's only FFI precedent, libagc: pgr-db/build.rs:17-55, pgr-db/src/agc_io.rs):
 *   - opaque handles, int return codes (0 = ok, <0 = error; pgr_b200_last_error() gives the thread-local message),
 *   - every output buffer is allocated by the library and released with pgr_b200_free(),
 *   - inputs are borrowed for the duration of the call only,
 *   - PODs are #[repr(C)]-compatible.
 * There is NO CPU fallback: every compute entry point fails with PGR_E_NO_DEVICE when no CUDA device is usable.
 */
#ifndef PGR_B200_H
#define PGR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGR_OK 0
#define PGR_E_ARG (-1)        /* NULL / inconsistent argument                                              */
#define PGR_E_SPEC (-2)       /* violates assert!(k <= 56), assert!(w <= 128), assert!(r > 0 && r < 13)    */
                              /* (shmmrutils.rs:443-445, :575-576)                                         */
#define PGR_E_NO_DEVICE (-3)  /* no CUDA device / driver: the library never computes on the CPU            */
#define PGR_E_CUDA (-4)       /* a CUDA runtime call failed                                                */
#define PGR_E_IO (-5)         /* file could not be read / written / parsed                                 */
#define PGR_E_LIMIT (-6)      /* sequence longer than 2^31-1 (MM128 pos is 31 bits, shmmrutils.rs:261-263) */
#define PGR_E_ASSERT (-7)     /* a reference assert would have fired (e.g. aln.rs:24)                      */

/* shmmrutils.rs:225-229 — x = hash<<8 | span(k), y = rid<<32 | pos<<1 | strand */
typedef struct { uint64_t x, y; } pgr_mm128;
/* shmmrutils.rs:20-27 */
typedef struct { uint32_t w, k, r, min_span, sketch; } pgr_shmmr_spec;
/* seq_db.rs:75 FragmentSignature (frg_id, seq_id, bgn, end, orientation); the 17-byte packed form exists only on disk */
typedef struct { uint32_t frg_id, sid, bgn, end; uint8_t ori; uint8_t pad_[3]; } pgr_frag_sig;
/* seq_db.rs:1198 FragmentHit minus its Vec: ((h0,h1),(pos0,pos1,orientation)) */
typedef struct { uint64_t h0, h1; uint32_t bgn, end; uint8_t ori; uint8_t pad_[7]; } pgr_query_pair;
/* aln.rs:10 HitPair ((q_bgn,q_end,q_ori),(t_bgn,t_end,t_ori)) */
typedef struct { uint32_t qb, qe, tb, te; uint8_t qo, to; uint8_t pad_[2]; } pgr_hit_pair;
/* graph_utils.rs:47-52 AdjPair (sid, ShmmrGraphNode a, ShmmrGraphNode b) */
typedef struct { uint32_t sid; uint8_t ori0, ori1, pad_[2]; uint64_t a0, a1, b0, b1; } pgr_adj_pair;

/* graph_utils.rs:47 ShmmrGraphNode (hash0, hash1, orientation) */
typedef struct { uint64_t h0, h1; uint8_t ori; uint8_t pad_[7]; } pgr_graph_node;
/* seq_db.rs:1001-1010 PBundleNode (node, Option<previous_node>, node_weight, is_leaf, global_rank, branch, branch_rank) */
typedef struct {
    pgr_graph_node node, prev;
    uint8_t has_prev, is_leaf, pad_[2];
    uint32_t weight, rank, branch, branch_rank, pad2_;
} pgr_dfs_node;

/* seq_db.rs:34-41 AlnSegment: type 0 = FullMatch, 1 = Match(a, b), 2 = Insertion(a as u8) */
typedef struct { uint32_t type, a, b; } pgr_aln_seg;
/* seq_db.rs:48-55 Fragment, by reference to the caller's sequences: kind 0 = AlnSegments((ref_frag, reversed, len, segs)),
 * 1 = Prefix, 2 = Internal, 3 = Suffix; the bases of kinds 1-3 (and the aligned bases of kind 0) are [bgn, end) of
 * sequence sid (for kinds 0 and 2 including the leading k-mer); len = end - bgn */
typedef struct {
    uint8_t kind, reversed, pad_[2];
    uint32_t sid, bgn, end, len, ref_frag, n_segs, pad2_;
    uint64_t seg_off;
} pgr_fragment;

/* query_fragment_to_hps arguments (aln.rs:147-158); Option<u32> is encoded as a negative value = None */
typedef struct {
    float penalty;
    int64_t max_count, max_count_query, max_count_target, max_aln_span, max_gap;
    int32_t oriented;
} pgr_query_params;

/* flat result of a batch of query_fragment_to_hps calls (aln.rs:145 TargetHitPairLists per query).
 * query q owns targets [q_target_off[q], q_target_off[q+1]); target t (ascending sid within a query) owns chains
 * [target_chain_off[t], target_chain_off[t+1]); chain c owns hits [chain_hit_off[c], chain_hit_off[c+1]). */
typedef struct {
    size_t n_queries, n_targets, n_chains, n_hits;
    uint64_t *q_target_off;
    uint32_t *target_sid;
    uint64_t *target_chain_off;
    float *chain_score;
    uint64_t *chain_hit_off;
    pgr_hit_pair *hits;
} pgr_query_result;

typedef struct pgr_b200_ctx pgr_b200_ctx;      /* one per (thread, device): streams, device arenas            */
typedef struct pgr_b200_index pgr_b200_index;  /* CompactSeqDB.frag_map as a device-resident CSR (seq_db.rs:95-100) */

/* ---- library / device ------------------------------------------------------------------------------------ */
int pgr_b200_device_count(void);
/* device used by the one-shot host calls (pgr_b200_sequence_to_shmmrs, pgr_b200_shmmrs_batch) of the calling thread;
 * default 0.  One process per GPU sets it to its local rank. */
int pgr_b200_set_default_device(int device);
const char *pgr_b200_last_error(void);
void pgr_b200_free(void *p);
/* pinned host staging for callers that want the H2D copy at PCIe rate (bench.py e2e leg) */
void *pgr_b200_host_alloc(size_t bytes);
void pgr_b200_host_free(void *p);
/* page-lock memory the caller already owns (a file buffer a reader thread parsed in place), so that the batch calls copy
 * from it at PCIe rate without a staging copy; 0 on success.  Unregister before freeing it. */
int pgr_b200_host_register(void *p, size_t bytes);
int pgr_b200_host_unregister(void *p);
#define PGR_TRANSPORT_PACKED 0 /* default: large host inputs cross PCIe as bit planes (2-3 bits per base) */
#define PGR_TRANSPORT_DIRECT 1 /* the caller's bytes are copied as they are (asynchronous only from page-locked memory) */
int pgr_b200_set_transport(int mode); /* process-wide; returns the previous mode.  PGR_B200_H2D=direct sets the initial one */
/* In PGR_TRANSPORT_PACKED mode a batch is still copied directly when it is small (< 4 MB) or when the host has few threads for
 * this process (< 10, e.g. 4 or 8 ranks on a 32-CPU node) and the source is page-locked: the direct copy is then the faster
 * one.  Transport the newest batch call of this process took: PGR_TRANSPORT_PACKED / PGR_TRANSPORT_DIRECT, -1 before the first. */
int pgr_b200_last_transport(void);
/* bytes the packed / hybrid transport has put on PCIe in this process so far (bit planes of packed slots + plain bytes of the
 * slots copied as they are); the difference around a call is what that call really moved host to device */
uint64_t pgr_b200_transport_bytes(void);
/* Host half of the packed transport the batch calls use for large inputs (2-3 bits per base over PCIe instead of 8; the
 * device expands them again): bases -> three bit planes per 32-byte block, bit j = byte j; v = 1: a base with code p1:p0
 * in the reference's LUT order (shmmrutils.rs:426-436: A/a/0 -> 0, C/c/1 -> 1, G/g/2 -> 2, T/t/3 -> 3; the padding of the last
 * block counts as code 0); v = 0: any other byte.  Returns the AND of the validity words (all ones: every byte a base — the
 * transport then leaves the validity plane at home).  Runs on the host (SIMD picked at run time, see pgr_b200_pack_isa);
 * exported so that it can be checked without a device.  Each plane holds (n_bytes + 31) / 32 words. */
uint32_t pgr_b200_pack_bases(const uint8_t *src, size_t n_bytes, uint32_t *p0, uint32_t *p1, uint32_t *v);
const char *pgr_b200_pack_isa(void);
int pgr_b200_pool_threads(void); /* host threads the library's host-side loops use (PGR_B200_HOST_THREADS overrides) */

/* ---- sequence_to_shmmrs ---------------------------------------------------------------------------------- */
/* replaces shmmrutils::sequence_to_shmmrs(rid, &seq, &spec, padding) -> Vec<MM128>   (shmmrutils.rs:657-669);
 * re-entrant (the reference calls it from rayon workers, seq_db.rs:461, pgr-query.rs:135). */
int pgr_b200_sequence_to_shmmrs(uint32_t rid, const uint8_t *seq, size_t len, const pgr_shmmr_spec *spec, int padding,
                                pgr_mm128 **out, size_t *n_out);
/* replaces CompactSeqDB::get_shmmrs_from_seqs (seq_db.rs:456-469): one call per batch, HOST buffers in and out.
 * offsets has n+1 entries (caller-allocated); *out is library-allocated. */
int pgr_b200_shmmrs_batch(size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens,
                          const pgr_shmmr_spec *spec, int padding, pgr_mm128 **out, size_t *offsets);

/* ---- explicit device context (device-resident pipeline; what bench.py times as `value`) ------------------ */
pgr_b200_ctx *pgr_b200_ctx_new(int device);
void pgr_b200_ctx_free(pgr_b200_ctx *ctx);
/* use an existing CUDA stream (cudaStream_t as void*) for all work of this ctx; NULL = the ctx's own stream */
int pgr_b200_ctx_set_stream(pgr_b200_ctx *ctx, void *cuda_stream);
/* copy a batch of host sequences into the ctx's device sequence store (replaces what is there) */
int pgr_b200_ctx_upload(pgr_b200_ctx *ctx, size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens);
/* adopt sequences that already live in device memory: base + offs[i] (each offs[i] 32-byte aligned, with at least
 * 16 KiB of readable slack before offs[0] and after the last sequence) */
int pgr_b200_ctx_set_device_seqs(pgr_b200_ctx *ctx, const uint8_t *dev_base, size_t n, const uint32_t *rids,
                                 const uint64_t *offs, const uint64_t *lens);
/* run sequence_to_shmmrs over the store; results stay on the device. n_shmmrs = total count. Asynchronous w.r.t. the
 * host except for small control read-backs. */
int pgr_b200_ctx_shmmrs(pgr_b200_ctx *ctx, const pgr_shmmr_spec *spec, int padding, size_t *n_shmmrs);
/* device pointers of the last pgr_b200_ctx_shmmrs result: MM128[n_shmmrs] and uint64 offsets[n+1] */
int pgr_b200_ctx_shmmrs_device(pgr_b200_ctx *ctx, const pgr_mm128 **d_mm, const uint64_t **d_offsets);
/* copy the last result to the host (library-allocated *out; offsets caller-allocated, n+1) */
int pgr_b200_ctx_shmmrs_download(pgr_b200_ctx *ctx, pgr_mm128 **out, size_t *offsets);
/* per-stage device times (ms, CUDA events) of the last ctx call; names is a static NULL-terminated array */
int pgr_b200_ctx_timings(pgr_b200_ctx *ctx, const char *const **names, const float **ms, size_t *n);
/* counters of the last shmmrs call: [0] kernel launches, [1] level-0 minimizers, [2] sequences replayed sequentially,
 * [3] arena retries */
int pgr_b200_ctx_counters(pgr_b200_ctx *ctx, uint64_t out[8]);

/* ---- ShmmrFragMap index (CompactSeqDB.frag_map, seq_db.rs:72-100) ------------------------------------------- */
/* frg_id_mode 0: FASTX numbering, a global running fragment counter (prefix, one per pair, suffix; an empty sequence
 *                consumes two) — CompactSeqDB::seq_to_compressed, seq_db.rs:203-231,326-347 (pgr-make-frgdb, load_from_fastx)
 * frg_id_mode 1: AGC numbering, the per-sequence pair ordinal — seq_to_index + load_index_from_seq_vec,
 *                seq_db.rs:360-418,573-615 (pgr-mdb).   device < 0 = the calling thread's default device. */
pgr_b200_index *pgr_b200_index_new(const pgr_shmmr_spec *spec, int frg_id_mode, int device);
void pgr_b200_index_free(pgr_b200_index *idx);
int pgr_b200_index_get_spec(const pgr_b200_index *idx, pgr_shmmr_spec *spec);
/* replaces CompactSeqDB::load_seqs_from_seq_vec / load_index_from_seq_vec (seq_db.rs:507-525, :573-615) for the index
 * part: shimmers of the batch, pairs, tuples appended in (sequence, position) order.  HOST buffers. */
int pgr_b200_index_add_batch(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens);
/* two-step variant for the multi-GPU build: stage computes the batch's shimmers and reports how many fragment ids the
 * batch consumes; commit emits the tuples with `frag_base` = fragments that precede this batch globally. */
int pgr_b200_index_stage_batch(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                               uint64_t *n_frags_in_batch);
int pgr_b200_index_commit_batch(pgr_b200_index *idx, uint32_t frag_base);
/* stage_batch for sequences that already live in device memory (layout rules of pgr_b200_ctx_set_device_seqs) */
int pgr_b200_index_stage_device(pgr_b200_index *idx, const uint8_t *dev_base, size_t n, const uint32_t *sids, const uint64_t *offs,
                                const uint64_t *lens, uint64_t *n_frags_in_batch);
/* stable sort by (h0,h1) and CSR build; implied by every read access */
int pgr_b200_index_finalize(pgr_b200_index *idx);
int pgr_b200_index_counts(pgr_b200_index *idx, size_t *n_keys, size_t *n_sigs, uint32_t *n_frags);
/* canonical form of the map: keys ascending ((h0,h1) as 2 x u64 each), offsets[n_keys+1], per-key signatures in
 * insertion order.  Caller-allocated (sizes from pgr_b200_index_counts). */
int pgr_b200_index_export_csr(pgr_b200_index *idx, uint64_t *keys, uint64_t *offsets, pgr_frag_sig *sigs);
/* the raw tuples in device memory (40-byte records {u64 h0,h1; u32 frg_id,sid,bgn,end,ori,pad}) for the inter-GPU
 * exchange, and the way back in */
int pgr_b200_index_tuples_device(pgr_b200_index *idx, void **dev_tuples, size_t *n_tuples);
int pgr_b200_index_set_tuples_device(pgr_b200_index *idx, const void *dev_tuples, size_t n_tuples);
/* stable partition of the tuples into n_parts key ranges (part p holds h0 in [splitters[p-1], splitters[p])), the
 * layout an all-to-all needs; counts[n_parts] out */
int pgr_b200_index_partition(pgr_b200_index *idx, size_t n_parts, const uint64_t *splitters, uint64_t *counts);
/* write_shmmr_map_file (seq_db.rs:1291-1326; keys written ascending) / read_mdb_file (seq_db.rs:1328-1407) */
int pgr_b200_index_write_mdb(pgr_b200_index *idx, const char *path);
pgr_b200_index *pgr_b200_index_read_mdb(const char *path, int device);

/* ---- multi-GPU ShmmrFragMap build -------------------------------------------------------------------- */
/* Sequences shard over the GPUs in blocks; shimmers and tuples are computed where the sequences are; ONE all-to-all of
 * the 40-byte tuples (grouped ncclSend/ncclRecv over NVLink, key ranges cut at sampled splitters of h0) leaves every GPU
 * with one key range, which it sorts (stable) into its CSR slice.  Slices in rank order = the canonical key-ascending map
 * of load_seqs_from_seq_vec / load_index_from_seq_vec (seq_db.rs:507-525, :573-615): per-key vectors in insertion order,
 * FASTX fragment ids from one global running counter.  Two process models:
 *   (1) one process per GPU (torchrun): pgr_b200_comm_unique_id on rank 0, broadcast the bytes, pgr_b200_comm_init_rank
 *       everywhere, then the collective pgr_b200_index_build_sharded with each rank's contiguous block of the sequence list;
 *   (2) one process, n_gpus devices (what §8b's `pgr_b200_index_new(spec, mode, n_gpus)` names): pgr_b200_mindex_*.    */
#define PGR_B200_COMM_ID_BYTES 128
typedef struct pgr_b200_comm pgr_b200_comm;      /* one rank of an NCCL communicator */
typedef struct pgr_b200_mindex pgr_b200_mindex;  /* a ShmmrFragMap sharded by key range over the GPUs of one process */
typedef struct {
    uint32_t rank, n_ranks;
    uint64_t n_tuples_local;   /* tuples produced from this rank's sequences                    */
    uint64_t n_tuples_sent;    /* of those, tuples that left this GPU in the all-to-all          */
    uint64_t n_tuples_owned;   /* tuples of this rank's key range after the exchange             */
    uint64_t bytes_sent, bytes_recv;   /* all-to-all payload that crossed NVLink (40 B per tuple) */
    uint64_t total_frags;      /* global fragment count (frags.len() of the reference)           */
    float stage_ms, partition_ms, exchange_ms, sort_ms;   /* CUDA events on this rank's stream    */
} pgr_shard_stats;
int pgr_b200_comm_unique_id(uint8_t id[PGR_B200_COMM_ID_BYTES]);
pgr_b200_comm *pgr_b200_comm_init_rank(const uint8_t id[PGR_B200_COMM_ID_BYTES], int rank, int n_ranks, int device);
void pgr_b200_comm_free(pgr_b200_comm *comm);
/* collective over `comm`: this rank's block (HOST buffers; ranks hold consecutive blocks of the global sequence list in
 * rank order) -> idx holds this rank's key range, finalized.  stats may be NULL. */
int pgr_b200_index_build_sharded(pgr_b200_index *idx, pgr_b200_comm *comm, size_t n, const uint32_t *sids, const uint8_t *const *seqs,
                                 const size_t *lens, pgr_shard_stats *stats);
/* the same with the block already resident in this rank's HBM (layout rules of pgr_b200_ctx_set_device_seqs) */
int pgr_b200_index_build_sharded_device(pgr_b200_index *idx, pgr_b200_comm *comm, const uint8_t *dev_base, size_t n, const uint32_t *sids,
                                        const uint64_t *offs, const uint64_t *lens, pgr_shard_stats *stats);
/* the exchange step alone (collective): idx holds committed tuples of this rank's block */
int pgr_b200_index_merge(pgr_b200_index *idx, pgr_b200_comm *comm, pgr_shard_stats *stats);

/* one process, n_gpus devices (0 .. n_gpus-1), one host thread and one NCCL rank per device (ncclCommInitAll) */
pgr_b200_mindex *pgr_b200_mindex_new(const pgr_shmmr_spec *spec, int frg_id_mode, int n_gpus);
/* explicit device list; shards that share a device (testing the sharded build on one GPU) exchange through
 * device-to-device copies instead of NCCL, which refuses duplicate devices */
pgr_b200_mindex *pgr_b200_mindex_new_devices(const pgr_shmmr_spec *spec, int frg_id_mode, int n_shards, const int *devices);
void pgr_b200_mindex_free(pgr_b200_mindex *m);
int pgr_b200_mindex_n_shards(const pgr_b200_mindex *m);
/* replaces load_seqs_from_seq_vec / load_index_from_seq_vec for the index part: the batch is cut into n_gpus consecutive
 * blocks of about equal bases; may be called once per input file like the reference does */
int pgr_b200_mindex_add_batch(pgr_b200_mindex *m, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens);
/* the all-to-all merge + per-owner sort; implied by every read access */
int pgr_b200_mindex_finalize(pgr_b200_mindex *m);
int pgr_b200_mindex_counts(pgr_b200_mindex *m, size_t *n_keys, size_t *n_sigs, uint32_t *n_frags);
int pgr_b200_mindex_stats(pgr_b200_mindex *m, int shard, pgr_shard_stats *out);
/* shard `s` as an ordinary index (borrowed; owned by m): query / adjacency / export of one key range */
pgr_b200_index *pgr_b200_mindex_shard(pgr_b200_mindex *m, int shard);
/* canonical CSR / .mdb of the whole map = the slices in shard order (same layout as the single-GPU calls) */
int pgr_b200_mindex_export_csr(pgr_b200_mindex *m, uint64_t *keys, uint64_t *offsets, pgr_frag_sig *sigs);
int pgr_b200_mindex_write_mdb(pgr_b200_mindex *m, const char *path);
/* the whole map as ONE ordinary finalized index on `device` (slices copied GPU to GPU, offsets rebased): what queries,
 * frag_map_to_adj_list and fragment compression run against after a multi-GPU build.  New handle, freed by the caller
 * with pgr_b200_index_free. */
pgr_b200_index *pgr_b200_mindex_gather(pgr_b200_mindex *m, int device);

/* ---- query, chaining, adjacency ------------------------------------------------------------------------------ */
/* replaces seq_db::raw_query_fragment(&frag_map, &query, &spec) -> Vec<FragmentHit> (seq_db.rs:1200-1228): per query
 * shimmer pair (strict '<' canonicalisation, :1213-1217) its key/position/orientation and the signatures of the key:
 * hits[hit_off[i] .. hit_off[i+1]).  Library-allocated outputs (pgr_b200_free). */
int pgr_b200_raw_query(pgr_b200_index *idx, const uint8_t *seq, size_t len, pgr_query_pair **pairs, size_t *n_pairs, uint64_t **hit_off,
                       pgr_frag_sig **hits);
/* The .mdb-resident variant: read_mdb_file_to_frag_locations (seq_db.rs:1409-1471) keeps only key -> (file offset, count) in
 * memory and leaves the signatures in the memory-mapped file; raw_query_fragment_from_mmap_midx (seq_db.rs:1230-1269, caller
 * ext.rs:285-342) computes the query's shimmers (on the device here), looks every pair up in that table and reads its signatures
 * out of the map (get_fragment_signatures_from_mmap_file).  Same outputs as pgr_b200_raw_query; nothing of the index is loaded
 * into HBM, so an .mdb larger than the device (or a one-off look-up) costs the header pass only. */
typedef struct pgr_b200_mdb_map pgr_b200_mdb_map;
pgr_b200_mdb_map *pgr_b200_mdb_map_open(const char *mdb_path);
void pgr_b200_mdb_map_close(pgr_b200_mdb_map *m);
int pgr_b200_mdb_map_info(const pgr_b200_mdb_map *m, pgr_shmmr_spec *spec, size_t *n_keys, size_t *n_sigs);
int pgr_b200_raw_query_mmap(pgr_b200_mdb_map *m, const uint8_t *seq, size_t len, pgr_query_pair **pairs, size_t *n_pairs, uint64_t **hit_off,
                            pgr_frag_sig **hits);
/* query_fragment_to_hps_from_mmap_file (ext.rs:285-342 -> seq_db.rs:1230-1269 + aln.rs) for a batch of queries: the signature
 * vectors of the keys the queries hit are read out of the map into a temporary device index on `device`, which then answers
 * the batch exactly as pgr_b200_query_batch does on the whole index (the filters and the chaining only ever see those keys).
 * Device memory and transfer are proportional to the hits, not to the .mdb. */
int pgr_b200_query_batch_mmap(pgr_b200_mdb_map *m, int device, size_t n_queries, const uint8_t *const *seqs, const size_t *lens,
                              const pgr_query_params *params, pgr_query_result **out);
/* replaces SeqIndexDB::query_fragment_to_hps (ext.rs:252-282 -> aln.rs:147-242) for a batch of queries (the reference
 * runs one rayon task per query, pgr-query.rs:135).  Canonical forms where the reference leaks hash-map order: targets
 * ascending by sid; equal-score chain heads by position in the q_bgn-sorted hit list. */
int pgr_b200_query_batch(pgr_b200_index *idx, size_t n_queries, const uint8_t *const *seqs, const size_t *lens,
                         const pgr_query_params *params, pgr_query_result **out);
void pgr_b200_query_result_free(pgr_query_result *r);
/* replaces aln::sparse_aln(&mut sp_hits, max_span, penalty, max_gap, oriented) (aln.rs:12-142) on one hit list; hits is
 * sorted in place (stable by q_bgn, aln.rs:21); max_gap < 0 = None.  PGR_E_ASSERT when n < 2 (aln.rs:24). */
int pgr_b200_sparse_aln(pgr_hit_pair *hits, size_t n, uint32_t max_span, float penalty, int64_t max_gap, int oriented, size_t *n_chains,
                        uint64_t **chain_off, float **scores, pgr_hit_pair **chain_hits);
/* replaces seq_db::frag_map_to_adj_list(&frag_map, min_count, keeps) -> AdjList (seq_db.rs:876-944); has_keeps = 0
 * encodes None */
int pgr_b200_adj_list(pgr_b200_index *idx, size_t min_count, const uint32_t *keeps, size_t n_keeps, int has_keeps, pgr_adj_pair **out,
                      size_t *n_out);

/* replaces seq_db::generate_smp_adj_list_for_seq(&seq, sid, &frag_map, &spec, min_count) -> AdjList (seq_db.rs:946-1000) for a
 * batch of sequences — SeqIndexDB::generate_mapg_gfa runs it over every sequence of the database when method != "from_fragmap"
 * (ext.rs:696-722), with min_count 0 for the sequences in `keeps`: pass that per sequence in min_counts[].  The result is the
 * per-sequence lists concatenated in the order given (the reference concatenates in hash-map order of seq_info: unpinned). */
int pgr_b200_smp_adj_list_for_seqs(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                                   const uint64_t *min_counts, pgr_adj_pair **out, size_t *n_out);

/* ---- fragment compression (the .frg/.sdx content of pgr-make-frgdb) --------------------------------------------- */
/* replaces the alignment branch of CompactSeqDB::seq_to_compressed (seq_db.rs:189-358; shmmrutils::match_reads
 * shmmrutils.rs:57-223, deltas_to_aln_segs seq_db.rs:113-156) for every sequence of a FASTX-mode index: the caller passes the
 * sequences again (exactly the indexed ones); the result is `frags` of the reference in frg_id order (one record per
 * fragment) plus the alignment segments they point into.  Library-allocated (pgr_b200_free). */
int pgr_b200_index_compress_fragments(pgr_b200_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens,
                                      pgr_fragment **frags, size_t *n_frags, pgr_aln_seg **segs, size_t *n_segs);

/* ---- MAP-graph traversal (host-side walks in the reference too; vertex weights come from the device index) ------ */
/* replaces seq_db::sort_adj_list_by_weighted_dfs(&frag_map, &adj_list, start) -> Vec<PBundleNode> (seq_db.rs:1013-1061,
 * walking graph_utils.rs:63-290 BiDiGraphWeightedDfs).  PGR_E_ASSERT when start is not a vertex ("Node not found",
 * graph_utils.rs:107) or a vertex is not a key of the index (seq_db.rs:1031). */
int pgr_b200_sort_adj_list_by_weighted_dfs(pgr_b200_index *idx, const pgr_adj_pair *adj, size_t n_adj, const pgr_graph_node *start,
                                           pgr_dfs_node **out, size_t *n_out);
/* replaces seq_db::get_principal_bundles_from_adj_list(&frag_map, &adj_list, path_len_cutoff) -> (Vec<Vec<ShmmrGraphNode>>,
 * AdjList) (seq_db.rs:1063-1186): bundle b owns vertices[bundle_off[b] .. bundle_off[b+1]) (longest first, stable), plus the
 * adjacency pairs between vertices of the long paths.  PGR_E_ASSERT on an empty adjacency list (seq_db.rs:1068). */
int pgr_b200_principal_bundles(pgr_b200_index *idx, const pgr_adj_pair *adj, size_t n_adj, size_t path_len_cutoff, pgr_graph_node **vertices,
                               uint64_t **bundle_off, size_t *n_bundles, pgr_adj_pair **filtered, size_t *n_filtered);

#ifdef __cplusplus
}
#endif
#endif
