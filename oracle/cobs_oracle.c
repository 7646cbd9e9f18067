/*
 * cobs_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See cobs_oracle.h for scope and parity status ("parity unpinned" against the
 * real cobs binary; pinned against XXH64 KATs and the reference's own filters).
 *
 * Each function cites what it restates:
 *   [A.n]  = SURVEY.md Appendix A item n (COBS 0.2.1 classic index, restated
 *            from the published algorithm; source absent from /root/reference)
 *   ref:   = file:line under /root/reference that calls / consumes it
 */
#include "cobs_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <immintrin.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ XXH64 */
/* [A.4] public xxHash64 specification; cobs: XXH64(canon_kmer, k, seed=j). */
#define XP1 0x9E3779B185EBCA87ULL
#define XP2 0xC2B2AE3D27D4EB4FULL
#define XP3 0x165667B19E3779F9ULL
#define XP4 0x85EBCA77C2B2AE63ULL
#define XP5 0x27D4EB2F165667C5ULL

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xround(uint64_t acc, uint64_t in) {
    acc += in * XP2; acc = rotl64(acc, 31); return acc * XP1;
}
static inline uint64_t xmerge(uint64_t acc, uint64_t v) {
    acc ^= xround(0, v); return acc * XP1 + XP4;
}

uint64_t orc_xxh64(const void* data, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        const uint8_t* lim = end - 32;
        do {
            v1 = xround(v1, rd64(p)); p += 8;
            v2 = xround(v2, rd64(p)); p += 8;
            v3 = xround(v3, rd64(p)); p += 8;
            v4 = xround(v4, rd64(p)); p += 8;
        } while (p <= lim);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else {
        h = seed + XP5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) {
        h ^= xround(0, rd64(p));
        h = rotl64(h, 27) * XP1 + XP4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)rd32(p) * XP1;
        h = rotl64(h, 23) * XP2 + XP3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * XP5;
        h = rotl64(h, 11) * XP1;
        p++;
    }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}

/* ------------------------------------------------------------ canonical */
/* [A.3] canonical = lexicographic min(kmer, reverse complement), ASCII order. */
static inline int comp_base(char c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        default: return -1;
    }
}

int orc_canonical(const char* kmer, uint32_t k, char* out) {
    /* compare kmer with its reverse complement left to right */
    int use_rc = 0, decided = 0;
    for (uint32_t i = 0; i < k; i++) {
        int rc = comp_base(kmer[k - 1 - i]);
        if (rc < 0 || comp_base(kmer[i]) < 0) return -1;
        if (!decided && kmer[i] != (char)rc) {
            use_rc = ((char)rc < kmer[i]);
            decided = 1;
        }
    }
    if (use_rc) {
        for (uint32_t i = 0; i < k; i++) out[i] = (char)comp_base(kmer[k - 1 - i]);
    } else {
        memcpy(out, kmer, k);
    }
    return 0;
}

/* ------------------------------------------------------------------ index */
/* [A.10] signature_size = ceil(n * (-h / ln(1 - fpr^(1/h)))) */
uint64_t orc_signature_size(uint64_t max_doc_kmers, uint64_t num_hashes, double fpr) {
    double h = (double)num_hashes;
    double ratio = -h / log(1.0 - pow(fpr, 1.0 / h));
    return (uint64_t)ceil((double)max_doc_kmers * ratio);
}

orc_index* orc_index_new(uint32_t term_size, uint8_t canonicalize, uint32_t n_docs,
                         uint64_t signature_size, uint64_t num_hashes,
                         const char* const* doc_names) {
    orc_index* idx = (orc_index*)calloc(1, sizeof(orc_index));
    if (!idx) return NULL;
    idx->term_size = term_size;
    idx->canonicalize = canonicalize;
    idx->n_docs = n_docs;
    idx->signature_size = signature_size;
    idx->num_hashes = num_hashes;
    idx->row_size = ((uint64_t)n_docs + 7) / 8;
    idx->doc_names = (char**)calloc(n_docs ? n_docs : 1, sizeof(char*));
    for (uint32_t d = 0; d < n_docs; d++) {
        if (doc_names && doc_names[d]) {
            idx->doc_names[d] = strdup(doc_names[d]);
        } else {
            char tmp[48];
            snprintf(tmp, sizeof tmp, "r%06u_DOC%06u", d, d);
            idx->doc_names[d] = strdup(tmp);
        }
    }
    uint64_t nbytes = signature_size * idx->row_size;
    idx->body = (uint8_t*)calloc(nbytes ? nbytes : 1, 1);
    if (!idx->body) { orc_index_free(idx); return NULL; }
    return idx;
}

void orc_index_free(orc_index* idx) {
    if (!idx) return;
    if (idx->doc_names) {
        for (uint32_t d = 0; d < idx->n_docs; d++) free(idx->doc_names[d]);
        free(idx->doc_names);
    }
    free(idx->body);
    free(idx);
}

/* term -> row for hash function j.  [A.4] row = XXH64(term,k,seed=j) % signature_size */
static inline uint64_t term_row(const orc_index* idx, const char* term, uint64_t j) {
    return orc_xxh64(term, idx->term_size, j) % idx->signature_size;
}

/* [A.10] classic-construct: every k-mer position of the document sets bit d in
 * each of its num_hashes rows; k-mers containing non-ACGT letters are skipped. */
int orc_index_add_doc(orc_index* idx, uint32_t d, const char* seq, uint64_t len) {
    uint32_t k = idx->term_size;
    if (d >= idx->n_docs) return -1;
    if (len < k) return 0;
    char* buf = (char*)malloc(k);
    for (uint64_t i = 0; i + k <= len; i++) {
        const char* term = seq + i;
        if (idx->canonicalize) {
            if (orc_canonical(term, k, buf) != 0) continue;
            term = buf;
        }
        for (uint64_t j = 0; j < idx->num_hashes; j++) {
            uint64_t row = term_row(idx, term, j);
            idx->body[row * idx->row_size + d / 8] |= (uint8_t)(1u << (d % 8));
        }
    }
    free(buf);
    return 0;
}

/* [A.1] header: "COBS:" "CLASSIC_INDEX" u32 version=1, u32 k, u8 canon, u32 D,
 * u64 signature_size, u64 num_hashes, D x (name '\n'), "CLASSIC_INDEX", body. */
static const char MAGIC0[] = "COBS:";
static const char MAGIC1[] = "CLASSIC_INDEX";

uint64_t orc_index_header_size(const orc_index* idx) {
    uint64_t n = 5 + 13 + 4 + 4 + 1 + 4 + 8 + 8;
    for (uint32_t d = 0; d < idx->n_docs; d++) n += strlen(idx->doc_names[d]) + 1;
    return n + 13;
}

int orc_index_write(const orc_index* idx, const char* path) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    uint32_t version = 1, nd = idx->n_docs;
    fwrite(MAGIC0, 1, 5, f);
    fwrite(MAGIC1, 1, 13, f);
    fwrite(&version, 4, 1, f);
    fwrite(&idx->term_size, 4, 1, f);
    fwrite(&idx->canonicalize, 1, 1, f);
    fwrite(&nd, 4, 1, f);
    fwrite(&idx->signature_size, 8, 1, f);
    fwrite(&idx->num_hashes, 8, 1, f);
    for (uint32_t d = 0; d < nd; d++) {
        fwrite(idx->doc_names[d], 1, strlen(idx->doc_names[d]), f);
        fputc('\n', f);
    }
    fwrite(MAGIC1, 1, 13, f);
    uint64_t nbytes = idx->signature_size * idx->row_size;
    if (nbytes && fwrite(idx->body, 1, nbytes, f) != nbytes) { fclose(f); return -1; }
    return fclose(f) == 0 ? 0 : -1;
}

orc_index* orc_index_parse(const uint8_t* buf, uint64_t len) {
    uint64_t p = 0;
    if (len < 5 + 13 + 29 || memcmp(buf, MAGIC0, 5) || memcmp(buf + 5, MAGIC1, 13)) return NULL;
    p = 18;
    uint32_t version, k, nd; uint8_t canon; uint64_t sig, nh;
    memcpy(&version, buf + p, 4); p += 4;
    if (version != 1) return NULL;
    memcpy(&k, buf + p, 4); p += 4;
    canon = buf[p]; p += 1;
    memcpy(&nd, buf + p, 4); p += 4;
    memcpy(&sig, buf + p, 8); p += 8;
    memcpy(&nh, buf + p, 8); p += 8;
    char** names = (char**)calloc(nd ? nd : 1, sizeof(char*));
    for (uint32_t d = 0; d < nd; d++) {
        uint64_t s = p;
        while (p < len && buf[p] != '\n') p++;
        if (p >= len) { for (uint32_t e = 0; e < d; e++) free(names[e]); free(names); return NULL; }
        names[d] = strndup((const char*)buf + s, p - s);
        p++;
    }
    orc_index* idx = NULL;
    if (p + 13 <= len && !memcmp(buf + p, MAGIC1, 13)) {
        p += 13;
        uint64_t row_size = ((uint64_t)nd + 7) / 8;
        if (len - p == sig * row_size) {
            idx = orc_index_new(k, canon, nd, sig, nh, (const char* const*)names);
            if (idx) memcpy(idx->body, buf + p, sig * row_size);
        }
    }
    for (uint32_t d = 0; d < nd; d++) free(names[d]);
    free(names);
    return idx;
}

/* `cobs query --load-complete`: the whole index is read into RAM once.  Header from the first
 * bytes, then the body straight into its final place (no intermediate copy). */
orc_index* orc_index_read(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    orc_index* idx = NULL;
    uint8_t head[47];
    if (n < 47 + 13 || fread(head, 1, 47, f) != 47 || memcmp(head, MAGIC0, 5) || memcmp(head + 5, MAGIC1, 13)) {
        fclose(f);
        return NULL;
    }
    uint32_t version, k, nd; uint8_t canon; uint64_t sig, nh;
    memcpy(&version, head + 18, 4); memcpy(&k, head + 22, 4); canon = head[26];
    memcpy(&nd, head + 27, 4); memcpy(&sig, head + 31, 8); memcpy(&nh, head + 39, 8);
    if (version != 1) { fclose(f); return NULL; }
    char** names = (char**)calloc(nd ? nd : 1, sizeof(char*));
    char* line = NULL; size_t cap = 0; int ok = 1;
    uint64_t hdr = 47;
    for (uint32_t d = 0; d < nd && ok; d++) {
        ssize_t got = getline(&line, &cap, f);
        if (got <= 0 || line[got - 1] != '\n') { ok = 0; break; }
        hdr += (uint64_t)got;
        names[d] = strndup(line, (size_t)got - 1);
    }
    free(line);
    char endm[13];
    if (ok && (fread(endm, 1, 13, f) != 13 || memcmp(endm, MAGIC1, 13))) ok = 0;
    hdr += 13;
    uint64_t row_size = ((uint64_t)nd + 7) / 8;
    if (ok && (uint64_t)n - hdr == sig * row_size) {
        idx = orc_index_new(k, canon, nd, sig, nh, (const char* const*)names);
        if (idx && sig * row_size > 0 && fread(idx->body, 1, sig * row_size, f) != sig * row_size) {
            orc_index_free(idx);
            idx = NULL;
        }
    }
    for (uint32_t d = 0; d < nd; d++) free(names[d]);
    free(names);
    fclose(f);
    return idx;
}

/* ------------------------------------------------------------------ query */
/* [A.6] a document is reported iff score >= threshold * K in IEEE double;
 * default materialisation ceil(t*K) ("at least 70% of the query kmers",
 * ref: config.yaml:14), floor kept behind a switch. */
uint32_t orc_threshold_terms(double threshold, uint32_t num_kmers, int floor_mode) {
    double x = threshold * (double)num_kmers;
    double r = floor_mode ? floor(x) : ceil(x);
    if (r < 0) r = 0;
    return (uint32_t)r;
}

/* [A.4] rows of every query k-mer: rows[i*h + j] */
static int64_t query_rows(const orc_index* idx, const char* seq, uint64_t len, uint64_t** out) {
    uint32_t k = idx->term_size;
    *out = NULL;
    if (len < k) return 0;                      /* [A.5] (L): L<k -> no terms */
    uint64_t K = len - k + 1, h = idx->num_hashes;
    uint64_t* rows = (uint64_t*)malloc(sizeof(uint64_t) * K * h);
    char* buf = (char*)malloc(k);
    for (uint64_t i = 0; i < K; i++) {
        const char* term = seq + i;
        if (idx->canonicalize) {
            if (orc_canonical(term, k, buf) != 0) { free(rows); free(buf); return -1; }
            term = buf;
        }
        for (uint64_t j = 0; j < h; j++) rows[i * h + j] = term_row(idx, term, j);
    }
    free(buf);
    *out = rows;
    return (int64_t)K;
}

/* [A.5] score[d] = #{i : bit d set in ALL h rows of k-mer i}; every position
 * counts (no de-duplication).  Plain bit loop -- the definition. */
int64_t orc_query_scores(const orc_index* idx, const char* seq, uint64_t len,
                         uint32_t* scores) {
    memset(scores, 0, sizeof(uint32_t) * idx->n_docs);
    uint64_t* rows;
    int64_t K = query_rows(idx, seq, len, &rows);
    if (K <= 0) return K;
    uint64_t h = idx->num_hashes, rs = idx->row_size;
    uint8_t* acc = (uint8_t*)malloc(rs ? rs : 1);
    for (int64_t i = 0; i < K; i++) {
        memcpy(acc, idx->body + rows[i * h] * rs, rs);
        for (uint64_t j = 1; j < h; j++) {
            const uint8_t* r = idx->body + rows[i * h + j] * rs;
            for (uint64_t b = 0; b < rs; b++) acc[b] &= r[b];
        }
        for (uint32_t d = 0; d < idx->n_docs; d++) scores[d] += (acc[d / 8] >> (d % 8)) & 1u;
    }
    free(acc);
    free(rows);
    return K;
}

/* [A.9] byte -> 8 x uint16 expansion table (cobs: 256 x 128-bit, _mm_add_epi16) */
static __m128i g_expand[256];
static int g_expand_ready = 0;
static void init_expand(void) {
    if (g_expand_ready) return;
    for (int b = 0; b < 256; b++) {
        uint16_t v[8];
        for (int t = 0; t < 8; t++) v[t] = (uint16_t)((b >> t) & 1);
        memcpy(&g_expand[b], v, 16);
    }
    g_expand_ready = 1;
}

static void slice_count(const orc_index* idx, const uint64_t* rows, int64_t K, uint64_t s,
                        uint32_t* scores) {
    uint64_t h = idx->num_hashes, rs = idx->row_size;
    uint64_t off = s * 16, w = rs - off < 16 ? rs - off : 16;
    __m128i acc[16];
    uint32_t wide[128];
    memset(wide, 0, sizeof wide);
    for (int b = 0; b < 16; b++) acc[b] = _mm_setzero_si128();
    int64_t since = 0;
    for (int64_t i = 0; i < K; i++) {
        uint8_t g[16] = {0};
        /* the row addresses are known up front: pull the line 16 rows ahead into cache, as any
         * tuned CPU implementation would (keeps the CPU baseline honest) */
        if (i + 16 < K) __builtin_prefetch(idx->body + rows[(i + 16) * h] * rs + off, 0, 1);
        memcpy(g, idx->body + rows[i * h] * rs + off, w);
        for (uint64_t j = 1; j < h; j++) {
            const uint8_t* r = idx->body + rows[i * h + j] * rs + off;
            for (uint64_t b = 0; b < w; b++) g[b] &= r[b];
        }
        for (uint64_t b = 0; b < w; b++) acc[b] = _mm_add_epi16(acc[b], g_expand[g[b]]);
        if (++since == 65535 || i == K - 1) {
            uint16_t tmp[128];
            memcpy(tmp, acc, sizeof tmp);
            for (int t = 0; t < 128; t++) wide[t] += tmp[t];
            for (int b = 0; b < 16; b++) acc[b] = _mm_setzero_si128();
            since = 0;
        }
    }
    uint32_t d0 = (uint32_t)(s * 128);
    for (uint32_t t = 0; t < 128 && d0 + t < idx->n_docs; t++) scores[d0 + t] = wide[t];
}

/* The same counting with 256-bit adds: two row bytes (16 documents) per _mm256_add_epi16, the
 * 16 x uint16 addend assembled from two entries of the 128-bit expansion table.  Not what cobs
 * 0.2.1 ships (its kernel is the SSE2 one above); offered so the CPU baseline is not held back
 * by the older instruction set.  A slice is 32 row bytes = 256 documents here. */
__attribute__((target("avx2"))) static void slice_count_avx2(const orc_index* idx, const uint64_t* rows,
                                                             int64_t K, uint64_t s, uint32_t* scores) {
    uint64_t h = idx->num_hashes, rs = idx->row_size;
    uint64_t off = s * 32, w = rs - off < 32 ? rs - off : 32;
    __m256i acc[16];
    uint32_t wide[256];
    memset(wide, 0, sizeof wide);
    for (int b = 0; b < 16; b++) acc[b] = _mm256_setzero_si256();
    int64_t since = 0;
    for (int64_t i = 0; i < K; i++) {
        uint8_t g[32] = {0};
        if (i + 16 < K) __builtin_prefetch(idx->body + rows[(i + 16) * h] * rs + off, 0, 1);
        memcpy(g, idx->body + rows[i * h] * rs + off, w);
        for (uint64_t j = 1; j < h; j++) {
            const uint8_t* r = idx->body + rows[i * h + j] * rs + off;
            for (uint64_t b = 0; b < w; b++) g[b] &= r[b];
        }
        for (uint64_t b = 0; b < (w + 1) / 2; b++)
            acc[b] = _mm256_add_epi16(acc[b], _mm256_set_m128i(g_expand[g[2 * b + 1]], g_expand[g[2 * b]]));
        if (++since == 65535 || i == K - 1) {
            uint16_t tmp[256];
            memcpy(tmp, acc, sizeof tmp);
            for (int t = 0; t < 256; t++) wide[t] += tmp[t];
            for (int b = 0; b < 16; b++) acc[b] = _mm256_setzero_si256();
            since = 0;
        }
    }
    uint32_t d0 = (uint32_t)(s * 256);
    for (uint32_t t = 0; t < 256 && d0 + t < idx->n_docs; t++) scores[d0 + t] = wide[t];
}

static int g_simd = -1; /* -1: ask the environment (ORC_SIMD=sse2|avx2), 0 sse2, 1 avx2 */
int orc_simd_mode(void) {
    if (g_simd < 0) {
        const char* e = getenv("ORC_SIMD");
        g_simd = (e && !strcmp(e, "avx2") && __builtin_cpu_supports("avx2")) ? 1 : 0;
    }
    return g_simd;
}
/* returns the mode in effect (avx2 is refused on a CPU without it) */
int orc_set_simd_mode(int avx2) {
    g_simd = (avx2 && __builtin_cpu_supports("avx2")) ? 1 : 0;
    return g_simd;
}

int64_t orc_query_scores_sliced(const orc_index* idx, const char* seq, uint64_t len,
                                uint32_t* scores, int n_threads) {
    init_expand();
    memset(scores, 0, sizeof(uint32_t) * idx->n_docs);
    uint64_t* rows;
    int64_t K = query_rows(idx, seq, len, &rows);
    if (K <= 0) return K;
    const int avx2 = orc_simd_mode();
    int64_t n_slices = (int64_t)((idx->row_size + (avx2 ? 31 : 15)) / (avx2 ? 32 : 16));
    if (n_threads <= 1) {
        for (int64_t s = 0; s < n_slices; s++) {
            if (avx2) slice_count_avx2(idx, rows, K, (uint64_t)s, scores);
            else slice_count(idx, rows, K, (uint64_t)s, scores);
        }
    } else {
#pragma omp parallel for num_threads(n_threads) schedule(static)
        for (int64_t s = 0; s < n_slices; s++) {
            if (avx2) slice_count_avx2(idx, rows, K, (uint64_t)s, scores);
            else slice_count(idx, rows, K, (uint64_t)s, scores);
        }
    }
    free(rows);
    return K;
}

/* [A.7] results sorted by score descending; equal scores by document index
 * ascending (the build's canonical tie order, SURVEY 8(a) tie-order note). */
static int hit_cmp(const void* a, const void* b) {
    const orc_hit* x = (const orc_hit*)a;
    const orc_hit* y = (const orc_hit*)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    return x->doc < y->doc ? -1 : (x->doc > y->doc ? 1 : 0);
}

uint32_t orc_select(const uint32_t* scores, uint32_t n_docs, uint32_t min_score,
                    orc_hit* hits) {
    uint32_t n = 0;
    for (uint32_t d = 0; d < n_docs; d++)
        if (scores[d] >= min_score) { hits[n].doc = d; hits[n].score = scores[d]; n++; }
    qsort(hits, n, sizeof(orc_hit), hit_cmp);
    return n;
}

int64_t orc_query_batch(const orc_index* idx, const char* seqs, const uint64_t* offs,
                        uint32_t nq, double threshold, int floor_mode, int n_threads,
                        int mode, uint32_t* n_pass) {
    init_expand();
    int64_t total = 0;
    int err = 0;
    if (n_threads < 1) n_threads = 1;
    if (mode == 0) {
        uint32_t* scores = (uint32_t*)malloc(sizeof(uint32_t) * (idx->n_docs + 1));
        for (uint32_t q = 0; q < nq; q++) {
            uint64_t len = offs[q + 1] - offs[q];
            int64_t K = orc_query_scores_sliced(idx, seqs + offs[q], len, scores, n_threads);
            if (K < 0) { err = 1; break; }
            uint32_t T = orc_threshold_terms(threshold, (uint32_t)K, floor_mode), c = 0;
            for (uint32_t d = 0; d < idx->n_docs; d++) c += scores[d] >= T;
            if (n_pass) n_pass[q] = c;
            total += c;
        }
        free(scores);
    } else {
#pragma omp parallel num_threads(n_threads) reduction(+ : total)
        {
            uint32_t* scores = (uint32_t*)malloc(sizeof(uint32_t) * (idx->n_docs + 1));
#pragma omp for schedule(dynamic, 16)
            for (int64_t q = 0; q < (int64_t)nq; q++) {
                uint64_t len = offs[q + 1] - offs[q];
                int64_t K = orc_query_scores_sliced(idx, seqs + offs[q], len, scores, 1);
                if (K < 0) { err = 1; continue; }
                uint32_t T = orc_threshold_terms(threshold, (uint32_t)K, floor_mode), c = 0;
                for (uint32_t d = 0; d < idx->n_docs; d++) c += scores[d] >= T;
                if (n_pass) n_pass[q] = c;
                total += c;
            }
            free(scores);
        }
    }
    return err ? -1 : total;
}

/* ---------------------------------------------------- synthetic workload v1 */
uint64_t orc_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

static inline uint32_t sub_base(uint32_t b, uint64_t m) {
    return (b + 1u + (uint32_t)(((m >> 16) & 0xFFFFu) % 3u)) & 3u;
}

uint32_t orc_synth_base(const orc_synth* s, uint32_t d, uint32_t pos) {
    uint32_t b = (uint32_t)(orc_mix64(s->seed ^ ((uint64_t)pos * 0xD6E8FEB86659FD93ULL)) >> 62);
    uint64_t clade = d / (s->clade_size ? s->clade_size : 1);
    uint64_t m1 = orc_mix64((s->seed + (clade + 1) * 0x9E3779B97F4A7C15ULL) ^
                            ((uint64_t)pos * 0xC2B2AE3D27D4EB4FULL));
    if ((uint32_t)(m1 & 0xFFFFu) < s->clade_sub_q16) b = sub_base(b, m1);
    uint64_t m2 = orc_mix64((s->seed + ((uint64_t)d + 0x100000001ULL) * 0xBF58476D1CE4E5B9ULL) ^
                            ((uint64_t)pos * 0x94D049BB133111EBULL));
    if ((uint32_t)(m2 & 0xFFFFu) < s->doc_sub_q16) b = sub_base(b, m2);
    return b;
}

void orc_synth_genome(const orc_synth* s, uint32_t d, char* out) {
    for (uint32_t p = 0; p < s->genome_len; p++) out[p] = "ACGT"[orc_synth_base(s, d, p)];
}

void orc_synth_read(const orc_synth* specs, uint32_t n_idx, uint64_t reads_seed,
                    uint64_t r, uint32_t read_len, uint32_t random_q8,
                    uint32_t err_q16, char* out) {
    uint64_t u = orc_mix64(reads_seed + r * 0x9E3779B97F4A7C15ULL);
    if ((uint32_t)(u & 0xFFu) < random_q8 || n_idx == 0) {
        for (uint32_t j = 0; j < read_len; j++)
            out[j] = "ACGT"[orc_mix64(u + (uint64_t)j * 0xD6E8FEB86659FD93ULL) >> 62];
        return;
    }
    const orc_synth* s = &specs[(uint32_t)((u >> 8) & 0xFFFFFFu) % n_idx];
    uint32_t d = (uint32_t)(u >> 32) % s->n_docs;
    uint64_t u2 = orc_mix64(u);
    uint32_t span = s->genome_len >= read_len ? s->genome_len - read_len + 1 : 1;
    uint32_t pos = (uint32_t)((u2 >> 1) % span);
    uint32_t strand = (uint32_t)(u2 & 1u);
    uint64_t u3 = orc_mix64(u2);
    for (uint32_t j = 0; j < read_len; j++) {
        uint32_t b = orc_synth_base(s, d, pos + j);
        uint64_t e = orc_mix64(u3 + (uint64_t)j * 0xC2B2AE3D27D4EB4FULL);
        if ((uint32_t)(e & 0xFFFFu) < err_q16) b = sub_base(b, e);
        if (strand) out[read_len - 1 - j] = "ACGT"[3u - b];
        else out[j] = "ACGT"[b];
    }
}
