/*
 * cobs_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the match stage of Phylign:
 *   `cobs query` (COBS 0.2.1, bioconda pin at /root/reference/envs/cobs.yaml:5),
 *   invoked by /root/reference/scripts/run_cobs_streaming.sh:24-29 and
 *   /root/reference/Snakefile:419-424,476-481.
 *
 * PARITY STATUS: **parity unpinned against the real `cobs` binary** -- the COBS
 * source is an un-vendored dependency (absent from /root/reference, no network),
 * so this file restates the published COBS classic-index algorithm (SURVEY.md
 * Appendix A).  What IS pinned: XXH64 against two independent implementations
 * (python-xxhash, libxxhash; SURVEY Appendix B KATs), the text protocol against
 * the reference's own consumers (postprocess_cobs.py / filter_queries.py, which
 * are executed unmodified to generate tests/golden/), and a pure-Python
 * brute-force re-derivation of scores on tiny inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (phylign_b200/) never does.
 */
#ifndef COBS_ORACLE_H
#define COBS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- XXH64 (public xxHash spec; COBS hashes terms with XXH64(term, k, seed=j)) */
uint64_t orc_xxh64(const void* data, size_t len, uint64_t seed);

/* ---- canonical k-mer: lexicographic min(kmer, revcomp) on ASCII (Appendix A.3).
 * Returns 0 on success, -1 if a non-ACGT letter is met (cobs aborts there;
 * Phylign sanitises queries upstream, /root/reference/Snakefile:326-332). */
int orc_canonical(const char* kmer, uint32_t k, char* out);

/* ---- classic index in memory (Appendix A.1): packed rows, row_size=(D+7)/8 */
typedef struct orc_index {
    uint32_t term_size;
    uint8_t canonicalize;
    uint32_t n_docs;
    uint64_t signature_size;
    uint64_t num_hashes;
    uint64_t row_size;
    char** doc_names;   /* n_docs NUL-terminated strings */
    uint8_t* body;      /* signature_size * row_size bytes, row-major */
} orc_index;

/* signature size for classic-construct (Appendix A.10) */
uint64_t orc_signature_size(uint64_t max_doc_kmers, uint64_t num_hashes, double fpr);

orc_index* orc_index_new(uint32_t term_size, uint8_t canonicalize, uint32_t n_docs,
                         uint64_t signature_size, uint64_t num_hashes,
                         const char* const* doc_names);
void orc_index_free(orc_index* idx);
/* add every k-mer of `seq` (k-mers with non-ACGT letters skipped) to document d */
int orc_index_add_doc(orc_index* idx, uint32_t d, const char* seq, uint64_t len);
/* serialise / parse the `.cobs_classic` byte layout (Appendix A.1) */
uint64_t orc_index_header_size(const orc_index* idx);
int orc_index_write(const orc_index* idx, const char* path);
orc_index* orc_index_read(const char* path);
/* parse from memory; returns NULL on malformed input */
orc_index* orc_index_parse(const uint8_t* buf, uint64_t len);

/* ---- query ------------------------------------------------------------------
 * threshold rounding switch (Appendix A.6): 0 = ceil(t*K) (default), 1 = floor */
uint32_t orc_threshold_terms(double threshold, uint32_t num_kmers, int floor_mode);

/* scores[d] = number of the K=len-k+1 query k-mers present in document d.
 * Returns K (0 if len<k), or -1 on non-ACGT letter. */
int64_t orc_query_scores(const orc_index* idx, const char* seq, uint64_t len,
                         uint32_t* scores);

/* Same arithmetic, in the shape cobs uses on CPU (Appendix A.9): hashes once,
 * then 128-document column slices spread over n_threads, each slice gathering
 * 16 B per row and adding a byte-expansion table with 16-bit SIMD adds
 * (flushed to 32-bit before overflow).  Used for the CPU baseline timing. */
int64_t orc_query_scores_sliced(const orc_index* idx, const char* seq, uint64_t len,
                                uint32_t* scores, int n_threads);

typedef struct orc_hit { uint32_t doc; uint32_t score; } orc_hit;
/* docs with score >= T sorted by (score desc, doc index asc); returns count.
 * `hits` must hold n_docs entries. */
uint32_t orc_select(const uint32_t* scores, uint32_t n_docs, uint32_t min_score,
                    orc_hit* hits);

/* Batch driver used for timing: queries concatenated in `seqs`, offs[nq+1].
 * mode 0: cobs shape (queries serial, slices over n_threads);
 * mode 1: queries spread over n_threads (scalar per-query path).
 * Writes per-query n_pass into n_pass[nq] (may be NULL) and returns total
 * number of passing (query,doc) pairs, or -1 on error. */
/* counting kernel: 0 = SSE2 byte-expansion adds (the cobs 0.2.1 shape, default), 1 = the same with
 * 256-bit adds (needs AVX2; refused otherwise).  ORC_SIMD=avx2 in the environment selects 1. */
int orc_simd_mode(void);
int orc_set_simd_mode(int avx2);
int64_t orc_query_batch(const orc_index* idx, const char* seqs, const uint64_t* offs,
                        uint32_t nq, double threshold, int floor_mode, int n_threads,
                        int mode, uint32_t* n_pass);

/* ---- synthetic workload spec v1 (restated independently in
 * phylign_b200/csrc/synth.cu; tests check byte equality) ------------------ */
typedef struct orc_synth {
    uint64_t seed;          /* per-index seed */
    uint32_t n_docs;
    uint32_t genome_len;
    uint32_t clade_size;    /* docs per clade (shared substitutions) */
    uint32_t clade_sub_q16; /* substitution prob * 65536, clade level */
    uint32_t doc_sub_q16;   /* substitution prob * 65536, private */
} orc_synth;
uint64_t orc_mix64(uint64_t x);
/* 2-bit base (0=A,1=C,2=G,3=T) of document d at position pos */
uint32_t orc_synth_base(const orc_synth* s, uint32_t d, uint32_t pos);
void orc_synth_genome(const orc_synth* s, uint32_t d, char* out /* genome_len */);
/* read r: from index (u>>8)%n_idx unless random; writes ASCII into out[read_len] */
void orc_synth_read(const orc_synth* idx_specs, uint32_t n_idx, uint64_t reads_seed,
                    uint64_t r, uint32_t read_len, uint32_t random_q8,
                    uint32_t err_q16, char* out);

#ifdef __cplusplus
}
#endif
#endif
