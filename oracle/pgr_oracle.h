/*
 * pgr_oracle.h — CPU ORACLE for the SHIMMER-index hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a C++ restatement (exposed through a C ABI for ctypes) of the reference's
 * algorithms on the path SURVEY.md §8 names.  The reference (GeneDx/pgr-tk) is Rust and
 * cannot be compiled in this image (no rustc/cargo), so there is no oracle/_ref build;
 * the restatement is pinned instead against the reference's own committed fixtures and
 * known-answer tests (tests/golden/, see tests/test_oracle_golden.py):
 *   - pgr-db/test/test_data/test_seqs_frag.{mdb,midx}  (index of test_seqs.fa, 80/56/4/64)
 *   - pgr-db/src/lib.rs:342-363   (boundary test, two literal sequences, 24/24/12/24, padding)
 *   - pgr-db/src/lib.rs:166-180   (rc_match on test_rev.fa, sketch spec)
 * Parity that no reference test pins (hash-map iteration order leaking into outputs:
 * .mdb key order, target order and equal-score chain-head choice of sparse_aln) is stated
 * as "parity unpinned" and compared on canonical forms.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libpgr_b200.so) never links or calls it.
 */
#ifndef PGR_ORACLE_H
#define PGR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x, y; } orc_mm128;                       /* shmmrutils.rs:225-229 */
typedef struct { uint32_t w, k, r, min_span, sketch; } orc_spec;   /* shmmrutils.rs:20-27   */
typedef struct { uint32_t frg_id, sid, bgn, end; uint8_t ori; uint8_t pad_[3]; } orc_sig; /* seq_db.rs:75 */
typedef struct { uint64_t h0, h1; uint32_t bgn, end; uint8_t ori; uint8_t pad_[7]; } orc_qpair; /* seq_db.rs:1198 (key, query pos) */
typedef struct { uint32_t qb, qe, tb, te; uint8_t qo, to; uint8_t pad_[2]; } orc_hitpair;  /* aln.rs:10 */
typedef struct { uint32_t sid; uint8_t ori0, ori1, pad_[2]; uint64_t a0, a1, b0, b1; } orc_adj; /* graph_utils.rs:47-52 */

typedef struct orc_index orc_index;

uint64_t orc_u64hash(uint64_t key);                                /* shmmrutils.rs:271-280 */

/* shmmrutils.rs:657-669.  *out is malloc'd (orc_free). Returns 0, or <0 on a violated assert. */
int orc_sequence_to_shmmrs(uint32_t rid, const uint8_t *seq, size_t len, const orc_spec *spec,
                           int padding, orc_mm128 **out, size_t *n_out);

/* Same, over a batch, `nthreads` worker threads each taking whole sequences (the reference's
 * rayon granularity, seq_db.rs:456-469).  offsets has n+1 entries. */
int orc_shmmrs_batch(size_t n, const uint32_t *rids, const uint8_t *const *seqs, const size_t *lens,
                     const orc_spec *spec, int padding, int nthreads, orc_mm128 **out, size_t *offsets);

void orc_free(void *p);

/* frg_id_mode: 0 = FASTX/global running fragment counter (seq_db.rs:203-231,326-347),
 *              1 = AGC/per-sequence pair ordinal (seq_db.rs:360-418,573-615). */
orc_index *orc_index_new(const orc_spec *spec, int frg_id_mode);
void orc_index_free(orc_index *idx);
int orc_index_add_batch(orc_index *idx, size_t n, const uint32_t *sids, const uint8_t *const *seqs,
                        const size_t *lens, int nthreads);
/* FASTA ingest following fasta_io.rs:46-118 + seq_db.rs:471-525 (sid = running record index). */
int orc_index_load_fasta(orc_index *idx, const char *path, int nthreads);
size_t orc_index_n_keys(const orc_index *idx);
size_t orc_index_n_sigs(const orc_index *idx);
size_t orc_index_n_seqs(const orc_index *idx);
/* canonical export: keys ascending by (h0,h1); keys[2*i], keys[2*i+1]; offsets n_keys+1; per-key
 * signatures in insertion order. Caller allocates. */
void orc_index_export(const orc_index *idx, uint64_t *keys, uint64_t *offsets, orc_sig *sigs);
/* .mdb writer (seq_db.rs:1291-1326) with keys ascending, .midx writer (seq_db.rs:795-807) */
int orc_index_write_mdb(const orc_index *idx, const char *path);
int orc_index_write_midx(const orc_index *idx, const char *path);
orc_index *orc_index_read_mdb(const char *path);                   /* seq_db.rs:1328-1407 */
void orc_index_get_spec(const orc_index *idx, orc_spec *spec);
/* sequence table, for .midx checks */
int orc_index_seq_info(const orc_index *idx, size_t i, uint32_t *sid, uint64_t *len, const char **name,
                       const char **source);
/* FASTA parser alone: returns number of records; arrays malloc'd. */
int orc_parse_fasta(const char *path, size_t *n, char ***names, uint8_t ***seqs, size_t **lens);

/* raw_query_fragment (seq_db.rs:1200-1228): per query pair the key/pos/ori and the hit range
 * [hit_off[i], hit_off[i+1]) into *hits.  Arrays malloc'd. */
int orc_raw_query(const orc_index *idx, const uint8_t *seq, size_t len, orc_qpair **pairs, size_t *n_pairs,
                  uint64_t **hit_off, orc_sig **hits);

/* sparse_aln (aln.rs:12-142) with the canonical chain-head tie rule (score desc, index in the
 * q_bgn-sorted list asc).  max_gap < 0 encodes None.  hits is sorted in place (stable, by qb).
 * Output: n_chains, chain_off[n_chains+1], scores[n_chains], chain_hits[...] malloc'd. */
int orc_sparse_aln(orc_hitpair *hits, size_t n, uint32_t max_span, float penalty, int64_t max_gap, int oriented,
                   size_t *n_chains, uint64_t **chain_off, float **scores, orc_hitpair **chain_hits);

/* query_fragment_to_hps (aln.rs:147-242 via ext.rs:252-282); targets ascending by sid (canonical).
 * Option<u32> arguments: negative encodes None. */
int orc_query_fragment_to_hps(const orc_index *idx, const uint8_t *seq, size_t len, float penalty,
                              int64_t max_count, int64_t max_count_query, int64_t max_count_target,
                              int64_t max_aln_span, int64_t max_gap, int oriented,
                              size_t *n_targets, uint32_t **target_sids, uint64_t **target_chain_off,
                              float **chain_scores, uint64_t **chain_hit_off, orc_hitpair **chain_hits);

/* frag_map_to_adj_list (seq_db.rs:876-944). keeps may be NULL (None). */
int orc_adj_list(const orc_index *idx, size_t min_count, const uint32_t *keeps, size_t n_keeps, int has_keeps,
                 orc_adj **out, size_t *n_out);

#ifdef __cplusplus
}
#endif
#endif
