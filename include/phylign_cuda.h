/*
 * phylign_cuda.h -- C ABI of libphylign_cuda.so (sm_100a), the drop-in boundary
 * for Phylign's match stage.
 *
 * What it replaces in the reference (file:line under /root/reference):
 *   scripts/run_cobs_streaming.sh:24-29, Snakefile:419-424, Snakefile:476-481
 *       `cobs query --load-complete -t THR -T N -i INDEX -f QUERIES`   -> phy_index_* + phy_match
 *   scripts/postprocess_cobs.py:21-38   per-batch top-N + ties          -> phy_match (top_n)
 *   scripts/filter_queries.py:105-156   global top-N + ties over batches -> phy_match_run(merge_top_n) + phy_merged_fetch, phy_merge_host
 *
 * Conventions: plain pointers and sizes only; every function returns PHY_OK (0)
 * or a negative status and leaves a message retrievable with phy_last_error();
 * buffers passed in are caller-owned and may be reused as soon as the call
 * returns; buffers handed out (phy_results*, phy_merged*) are owned by the
 * library and released only with phy_results_free / phy_merged_free.
 * One phy_ctx drives ONE GPU from one host thread (one process per GPU; the
 * multi-GPU exchange goes through phy_nccl_*).  There is no CPU fallback: every
 * compute entry point fails with PHY_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PHYLIGN_CUDA_H
#define PHYLIGN_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHY_OK 0
#define PHY_ERR_CUDA (-1)   /* CUDA runtime / no device */
#define PHY_ERR_ARG (-2)    /* invalid argument */
#define PHY_ERR_NOMEM (-3)  /* HBM budget or host memory exhausted */
#define PHY_ERR_STATE (-4)  /* call out of order (e.g. match before queries) */
#define PHY_ERR_NCCL (-5)   /* NCCL failure / libnccl not loadable */
#define PHY_ERR_QUERY (-6)  /* query holds a letter outside ACGT (cobs aborts too) */
#define PHY_ERR_IO (-7)     /* file could not be created / written / renamed */

#define PHY_ABI_VERSION 1

typedef struct phy_ctx phy_ctx;

int phy_abi_version(void);
/* number of visible CUDA devices (0 and PHY_ERR_CUDA when none) */
int phy_device_count(int* n);
/* hbm_budget = 0 -> use what cudaMemGetInfo reports free, minus a safety margin */
int phy_ctx_create(phy_ctx** out, int device, uint64_t hbm_budget);
void phy_ctx_destroy(phy_ctx* ctx);
/* ctx may be NULL: returns the message of the last failed call without a ctx */
const char* phy_last_error(const phy_ctx* ctx);

/* ------------------------------------------------------------------ index store
 * A COBS classic index is streamed in as the packed body
 * bytes (signature_size rows of ceil(n_docs/8) bytes, file order); the library
 * stages through its own pinned ring and re-strides rows on the device.  Header
 * parsing and document names stay with the caller. */
int phy_index_begin(phy_ctx* ctx, const char* batch_name, uint32_t term_size,
                    uint8_t canonicalize, uint64_t signature_size, uint64_t num_hashes,
                    uint32_t n_docs, int* idx_id);
int phy_index_push(phy_ctx* ctx, int idx_id, const void* host_chunk, uint64_t nbytes);
/* the body of a decompressed `.cobs_classic` file ({decompression_dir}/{batch}.cobs_classic of
 * Snakefile:364-387) straight from the file to HBM: n_threads readers fill a ring of page-locked
 * slots, DMA + re-stride on a dedicated upload stream (may run beside phy_match_run of the same
 * context from another host thread).  body_offset = header length; the file must hold exactly
 * body_offset + signature_size*row_size bytes.  Follow with phy_index_commit. */
int phy_index_load_file(phy_ctx* ctx, int idx_id, const char* path, uint64_t body_offset, int n_threads);
/* fails with PHY_ERR_STATE unless exactly signature_size*row_size bytes arrived */
int phy_index_commit(phy_ctx* ctx, int idx_id);
int phy_index_evict(phy_ctx* ctx, int idx_id);
/* sort ranks used by the merge key (-score, batch, ref) of filter_queries.py:135:
 * batch_rank = rank of the batch name among ALL batches of the job (global across
 * GPUs, < 4096), ref_rank[d] = rank of document d's accession inside the batch.
 * Defaults: batch_rank = idx_id, ref_rank[d] = d. */
int phy_index_set_ranks(phy_ctx* ctx, int idx_id, uint32_t batch_rank,
                        const uint32_t* ref_rank);

/* a resident index can be left out of phy_match_run without evicting it (default: active) */
int phy_index_set_active(phy_ctx* ctx, int idx_id, int active);

typedef struct phy_index_info {
    uint64_t signature_size, num_hashes, hbm_bytes;
    uint32_t term_size, n_docs, row_size, row_stride, batch_rank;
    uint8_t canonicalize, committed;
} phy_index_info;
int phy_index_info_get(phy_ctx* ctx, int idx_id, phy_index_info* out);
int phy_index_count(phy_ctx* ctx, int* n_resident);
/* packed body bytes back to the host (tests, cache files) */
int phy_index_download(phy_ctx* ctx, int idx_id, void* host_out, uint64_t nbytes);

/* ------------------------------------------------------------- pinned host memory
 * Optional: buffers from phy_host_alloc are page-locked, so phy_queries_set /
 * phy_index_push DMA straight from them (no staging copy).  Any other host pointer
 * works too (it is staged through the library's pinned ring). */
int phy_host_alloc(size_t bytes, void** out);
void phy_host_free(void* p);

/* --------------------------------------------------------------------- queries
 * seq_concat: ASCII bases of all queries back to back (upper-case ACGT only,
 * the contract of intermediate/01_queries_merged, Snakefile:326-332);
 * offs[nq+1]: start of each query in seq_concat.  Copies host -> HBM. */
int phy_queries_set(phy_ctx* ctx, const char* seq_concat, const uint64_t* offs, uint32_t nq);

/* Query preparation of rule fix_query (Snakefile:314-332, `seqtk seq -A -U -C | awk gsub(/[^ACGT]/,"A")`)
 * for the bases, on the device: upper-case, every letter outside ACGT becomes 'A'.  In place over n bytes
 * of a host buffer (upload, one kernel, download).  With the option "sanitize_queries" 1, phy_queries_set
 * applies the same transform to the uploaded queries instead of rejecting such letters. */
int phy_fix_bases(phy_ctx* ctx, char* bases, uint64_t n);

/* ----------------------------------------------------------------------- match */
typedef struct phy_match_params {
    double threshold;     /* cobs -t : doc reported iff score >= threshold*K */
    uint32_t top_n;       /* postprocess_cobs.py -n : keep N best + ties per (query,index); 0 = keep all */
    uint32_t floor_mode;  /* 0: T=ceil(t*K) (default)  1: T=floor(t*K)  (SURVEY A.6 switch) */
} phy_match_params;

typedef struct phy_hit { uint32_t doc; uint32_t score; } phy_hit;
/* one non-empty (query, index) block of the cobs output */
typedef struct phy_unit {
    uint32_t query;    /* position in phy_queries_set order */
    uint32_t index;    /* idx_id */
    uint32_t n_pass;   /* docs with score >= T (the header count, before top-N) */
    uint32_t n_kept;   /* hits stored: N best + ties */
    uint64_t offset;   /* first hit in phy_results.hits */
} phy_unit;

typedef struct phy_results {
    uint32_t n_queries, n_indexes;
    uint64_t n_units;      /* sorted by (index, query); absent pairs have 0 hits */
    phy_unit* units;
    uint64_t n_hits;
    phy_hit* hits;         /* per unit sorted by (score desc, doc asc) */
    const uint32_t* n_kmers;   /* [n_queries] K = max(L-k+1, 0) */
    uint64_t h2d_bytes, d2h_bytes;
} phy_results;

/* device-only part: hash + gather/count + threshold/top-N (+ local merge when
 * merge_top_n > 0) for the queries set last, against all committed indexes.
 * Results stay in HBM until phy_results_fetch / phy_merged_fetch. */
int phy_match_run(phy_ctx* ctx, const phy_match_params* p, uint32_t merge_top_n);
int phy_results_fetch(phy_ctx* ctx, phy_results** out);
void phy_results_free(phy_results* r);
/* convenience: phy_match_run + phy_results_fetch */
int phy_match(phy_ctx* ctx, const phy_match_params* p, phy_results** out);

/* dense score matrix of ONE index for the queries set last: scores[nq*n_docs]
 * (uint32, row = query).  Verification / `bit-exact scores` entry point. */
int phy_scores(phy_ctx* ctx, int idx_id, uint32_t* host_scores);

/* ----------------------------------------------------------------------- merge
 * filter_queries.py semantics over all indexes of all GPUs: per query every
 * candidate whose score >= the N-th largest score, ordered by
 * (score desc, batch_rank asc, ref_rank asc). */
typedef struct phy_cand { uint32_t score; uint32_t batch_rank; uint32_t doc; uint32_t ref_rank; } phy_cand;
typedef struct phy_merged {
    uint32_t n_queries;
    uint64_t* offs;       /* [n_queries+1] */
    phy_cand* cands;
    uint64_t d2h_bytes;
} phy_merged;
/* needs phy_match_run(..., merge_top_n > 0) before.  With NCCL initialised the
 * per-GPU lists are gathered over NVLink and merged on rank 0 (other ranks get
 * an empty list with n_queries set). */
int phy_merged_fetch(phy_ctx* ctx, phy_merged** out);
void phy_merged_free(phy_merged* m);
/* Which queries' merged lists this context holds after phy_match_run: [0, nq) on one GPU; with NCCL
 * either everything on rank 0 ("merge_mode" 0, the default) or one slice of the queries per rank
 * ("merge_mode" 1: the candidate lists are exchanged all-to-all over NVLink and every rank finalises
 * and downloads only its slice; offs of phy_merged stays nq+1 long, empty outside the slice). */
int phy_merged_range(phy_ctx* ctx, uint32_t* q_lo, uint32_t* q_hi);
/* Merge candidates that come from the host (the `filter_queries.py -n N -q fa match files...`
 * entry: match files parsed by the driver): cands[offs[q] .. offs[q+1]) belong to query q;
 * score, batch_rank (<4096) and ref_rank (<2^20) form the sort key, doc is carried along.
 * Result via phy_merged_fetch.  Local to this context (no collective). */
int phy_merge_host(phy_ctx* ctx, uint32_t n_queries, uint32_t top_n, const uint64_t* offs,
                   const phy_cand* cands);

/* ------------------------------------------------------------------ text output
 * Host-side formatters (no GPU work) so drivers need no per-line loops.
 * phy_format_cobs_text: what `cobs query` prints for index idx_id: per query
 *   "*<header>\t<n_pass>\n" then "<doc name>\t<score>\n" lines; skip[q] != 0 drops the block
 *   (records without sequence); strip_prefix writes names as postprocess_cobs.py:16-18 does.
 *   headers/names are concatenated strings with offsets [n+1].
 * phy_format_filter_fasta: filter_queries.py:152-156 ">{qname} {ref1,ref2,...}\n{seq}\n";
 *   ref_names[b]/ref_offs[b]/ref_counts[b] describe the accessions of batch_rank b by doc id.
 * Buffers are released with phy_text_free. */
int phy_format_cobs_text(const phy_results* r, uint32_t idx_id, const char* headers, const uint64_t* hoffs,
                         const uint8_t* skip, const char* names, const uint64_t* noffs, uint32_t n_docs,
                         int strip_prefix, char** out, uint64_t* out_len);
int phy_format_filter_fasta(const phy_merged* m, const char* qnames, const uint64_t* qnoffs,
                            const char* seqs, const uint64_t* soffs, uint32_t n_batches,
                            const char* const* ref_names, const uint64_t* const* ref_offs,
                            const uint32_t* ref_counts, char** out, uint64_t* out_len);
void phy_text_free(char* p);
/* Parser of a (decompressed) match file with the rules of filter_queries.py:27-66 (`cobs_iterator`):
 * per block the query name (offset/length into `text`) and its hits [first_hit[b], first_hit[b+1]);
 * per hit the accession (text after the single '_' of the name) and the k-mer count.  Host only.
 * Malformed input (hit before a header, a name without exactly one '_', non-integer counts, empty
 * file) returns PHY_ERR_ARG -- the reference raises on the same inputs. */
typedef struct phy_match_text {
    uint64_t n_blocks, n_hits;
    uint64_t* q_off;      /* [n_blocks] */
    uint32_t* q_len;      /* [n_blocks] */
    uint64_t* first_hit;  /* [n_blocks+1] */
    uint64_t* ref_off;    /* [n_hits] */
    uint32_t* ref_len;    /* [n_hits] */
    uint32_t* kmers;      /* [n_hits] */
} phy_match_text;
int phy_parse_match_text(const char* text, uint64_t len, phy_match_text** out);
void phy_match_text_free(phy_match_text* r);

/* ------------------------------------------------------------------ query files
 * `cobs query -f` record rules (the reader behind run_cobs_streaming.sh:24-29):
 * '>' or ';' opens a record, sequence lines are concatenated, empty lines skipped, records without
 * sequence dropped.  Flat arrays, no per-record objects: seqs/soffs go to phy_queries_set as they are,
 * headers/hoffs to the text writers.  name_len[q] = length
 * of the query name (header up to the first blank, filter_queries.py:59,80).  simple != 0: plain '>'
 * FASTA for which readfq (filter_queries.py:69-102) yields the same records.  Host only. */
typedef struct phy_fasta {
    uint32_t n;
    uint8_t simple, seqs_pinned;
    char* seqs;        uint64_t* soffs;    /* [n+1] */
    char* headers;     uint64_t* hoffs;    /* [n+1] header lines without their first character */
    uint32_t* name_len;                    /* [n] */
} phy_fasta;
int phy_fasta_read(const char* path, phy_fasta** out);
void phy_fasta_free(phy_fasta* f);
/* filter_queries.py:152-156,195-199 output written straight to `final_path` (tmp + rename):
 * ">{qname} {ref,ref,...}\n{seq}\n" for queries [q_begin, min(q_end, n_queries)) of `m` (a rank that holds
 * one slice of the merged lists writes one part); qname = headers[hoffs[q] .. +name_len[q]).
 * ref_names/ref_offs/ref_counts as in phy_format_filter_fasta.  append != 0: `final_path` is a file the
 * caller owns (its own temporary); the text is appended, nothing is renamed (query blocks written in order). */
int phy_write_filter_fasta(const char* final_path, const phy_merged* m, const char* headers,
                           const uint64_t* hoffs, const uint32_t* name_len, const char* seqs,
                           const uint64_t* soffs, uint32_t n_batches, const char* const* ref_names,
                           const uint64_t* const* ref_offs, const uint32_t* ref_counts,
                           uint32_t q_begin, uint32_t q_end, int append, uint64_t* file_bytes /* may be NULL */);

/* -------------------------------------------------------------- match-file writer
 * The tail of the reference's per-batch pipeline, `... | postprocess_cobs.py -n N | gzip --fast >
 * intermediate/03_match/{batch}____{qfile}.gz` (Snakefile:425-427,467-469,482-484), for the results of
 * one query block and any number of indexes at once: formatted and deflated (zlib, level 1 = --fast)
 * on n_threads host threads, one gzip member per (file, query range), appended in query order.
 * Host only (no GPU work), callable from a thread other than the one driving the ctx.
 * A phy_mfile is written as <final_path>.tmp.<pid> and renamed by phy_mfile_commit: nothing partial
 * is ever visible at the output path.  gzip_level 0 writes plain text. */
typedef struct phy_mfile phy_mfile;
int phy_mfile_open(const char* final_path, int gzip_level, phy_mfile** out);
int phy_mfile_commit(phy_mfile* f, uint64_t* file_bytes /* may be NULL */);
void phy_mfile_abort(phy_mfile* f);
typedef struct phy_mfile_job {
    phy_mfile* file;
    uint32_t idx_id;          /* units of this index go to `file` */
    uint32_t n_docs;
    const char* names;        /* document names of the index, concatenated */
    const uint64_t* noffs;    /* [n_docs+1] */
} phy_mfile_job;
typedef struct phy_write_stats {   /* accumulated (+=) over calls; *_s are thread-seconds except wall_s */
    double format_s, deflate_s, write_s, wall_s;
    uint64_t text_bytes, file_bytes, n_header_lines, n_hit_lines;
} phy_write_stats;
/* headers/hoffs/skip/strip_prefix as in phy_format_cobs_text */
int phy_write_match_blocks(const phy_results* r, const phy_mfile_job* jobs, uint32_t n_jobs,
                           const char* headers, const uint64_t* hoffs, const uint8_t* skip,
                           int strip_prefix, int n_threads, phy_write_stats* stats /* may be NULL */);

/* ------------------------------------------------------------------- multi-GPU */
#define PHY_NCCL_ID_BYTES 128
int phy_nccl_unique_id(void* id_out /* PHY_NCCL_ID_BYTES */);
int phy_nccl_init(phy_ctx* ctx, const void* id, int rank, int n_ranks);
/* collective end of the communicator (ncclCommFinalize + ncclCommDestroy): all ranks call it at the same
 * point.  A context destroyed without it aborts its communicator locally (never blocks on peers). */
int phy_nccl_finalize(phy_ctx* ctx);

/* -------------------------------------------------------------------- timing
 * CUDA-event timer on the stream every kernel of this ctx is launched on. */
int phy_timer_start(phy_ctx* ctx);
int phy_timer_stop(phy_ctx* ctx, float* ms);
int phy_sync(phy_ctx* ctx);
/* per-phase device times (ms) of the last phy_match_run:
 * [0] hash  [1] gather+count(+select)  [2] sort/merge  [3] launches of own kernels */
int phy_last_phase_ms(phy_ctx* ctx, float out[4]);
/* index-row bytes (file row size, not stride) the fused ring kernel really gathered in the last
 * phy_match_run: < sum K*row_size when threshold pruning ended units early */
int phy_last_gather_bytes(phy_ctx* ctx, uint64_t* bytes);
/* the same per index (per-batch log lines: rows read / all rows) */
int phy_last_gather_bytes_of(phy_ctx* ctx, int idx_id, uint64_t* bytes);
/* HBM accounting of this context: bytes it may use in total / bytes in use now */
int phy_ctx_budget(phy_ctx* ctx, uint64_t* budget, uint64_t* used);
/* switches: "prune" 0|1 (exact threshold pruning, default 1; 0 for A/B measurements),
 * "pinned_results" 0|1 (phy_results / phy_merged in page-locked memory from a reuse pool, default 1;
 * 0 = plain host memory, cheaper for a single fetch), "merge_mode" 0|1 (multi-GPU: merged lists on rank 0 |
 * one slice of the queries per rank, see phy_merged_range), "shard_query_upload" 0|1 (multi-GPU, all ranks
 * pass identical queries: each uploads 1/R of the bases, an NCCL all-gather completes them),
 * "sanitize_queries" 0|1 (see phy_fix_bases) */
int phy_ctx_set_option(phy_ctx* ctx, const char* name, int64_t value);
/* write a buffer larger than L2 (bench hygiene between timed iterations) */
int phy_flush_l2(phy_ctx* ctx);

/* ---------------------------------------------- synthetic workload (bench/tests)
 * Spec v1, restated independently in oracle/cobs_oracle.c (tests check byte
 * equality).  Builds an index on the device from procedural genomes
 * (`cobs classic-construct` shape: every k-mer of every document sets its bit). */
typedef struct phy_synth_spec {
    uint64_t seed;
    uint32_t n_docs, genome_len, clade_size, clade_sub_q16, doc_sub_q16;
} phy_synth_spec;
/* index must have been phy_index_begin()'d; fills and commits it */
int phy_index_synth(phy_ctx* ctx, int idx_id, const phy_synth_spec* spec);
/* `cobs classic-construct` step for real sequences: OR every k-mer of query q (queries set
 * last) into document doc_of_query[q] of a committed index (0xFFFFFFFF = skip the query). */
int phy_index_insert(phy_ctx* ctx, int idx_id, const uint32_t* doc_of_query);
/* n_reads reads of read_len bases into host_out (n_reads*read_len ASCII bytes) */
int phy_synth_reads(phy_ctx* ctx, const phy_synth_spec* specs, uint32_t n_specs,
                    uint64_t reads_seed, uint64_t first_read, uint32_t n_reads,
                    uint32_t read_len, uint32_t random_q8, uint32_t err_q16, char* host_out);

#ifdef __cplusplus
}
#endif
#endif
