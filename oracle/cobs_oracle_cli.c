/*
 * cobs_oracle_cli.c -- CPU ORACLE command line (test infrastructure, NOT product).
 *
 *   cobs_oracle query [--load-complete] -t THR -T THREADS -i INDEX
 *                     [--index-sizes N] -f QUERY.fa [--floor] [--threads-over-slices] [--avx2]
 *       restates the `cobs query` call of
 *       /root/reference/scripts/run_cobs_streaming.sh:24-29: prints, per FASTA
 *       record, "*<header minus first char>\t<n>\n" followed by n lines
 *       "<doc_name>\t<score>\n" sorted by score descending ([A.8]).
 *   cobs_oracle construct -o OUT.cobs_classic [-k 31] [--num-hashes 1]
 *                     [--false-positive-rate 0.3] [--no-canonicalize] DOCS.fa
 *       restates `cobs classic-construct` for one multi-FASTA whose records are
 *       the documents (record name = document name), [A.10].
 *
 * parity unpinned against the real cobs binary (see cobs_oracle.h).
 */
#include "cobs_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { char* name; char* seq; size_t len, cap; } rec_t;

/* [A.8] records start at '>' or ';'; sequence lines concatenated; empty lines
 * skipped.  Returns malloc'd array. */
static rec_t* read_fasta(const char* path, size_t* n_out) {
    FILE* f = strcmp(path, "-") ? fopen(path, "r") : stdin;
    if (!f) return NULL;
    rec_t* recs = NULL;
    size_t n = 0, cap = 0;
    char* line = NULL;
    size_t lcap = 0;
    ssize_t got;
    while ((got = getline(&line, &lcap, f)) >= 0) {
        while (got > 0 && (line[got - 1] == '\n' || line[got - 1] == '\r')) line[--got] = 0;
        if (got == 0) continue;
        if (line[0] == '>' || line[0] == ';') {
            if (n == cap) { cap = cap ? cap * 2 : 64; recs = (rec_t*)realloc(recs, cap * sizeof(rec_t)); }
            recs[n].name = strdup(line + 1);
            recs[n].seq = NULL; recs[n].len = 0; recs[n].cap = 0;
            n++;
        } else if (n > 0) {
            rec_t* r = &recs[n - 1];
            if (r->len + (size_t)got + 1 > r->cap) {
                r->cap = (r->len + (size_t)got + 1) * 2;
                r->seq = (char*)realloc(r->seq, r->cap);
            }
            memcpy(r->seq + r->len, line, (size_t)got);
            r->len += (size_t)got;
            r->seq[r->len] = 0;
        }
    }
    free(line);
    if (f != stdin) fclose(f);
    *n_out = n;
    return recs;
}

static int cmd_query(int argc, char** argv) {
    const char* index = NULL; const char* qfile = NULL;
    double thr = 0.8; int threads = 1, floor_mode = 0, over_slices = 0;
    for (int i = 0; i < argc; i++) {
        if (!strcmp(argv[i], "-t") && i + 1 < argc) thr = strtod(argv[++i], NULL);
        else if (!strcmp(argv[i], "-T") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-i") && i + 1 < argc) index = argv[++i];
        else if (!strcmp(argv[i], "-f") && i + 1 < argc) qfile = argv[++i];
        else if (!strcmp(argv[i], "--index-sizes") && i + 1 < argc) ++i;
        else if (!strcmp(argv[i], "--load-complete")) {}
        else if (!strcmp(argv[i], "--floor")) floor_mode = 1;
        else if (!strcmp(argv[i], "--threads-over-slices")) over_slices = 1;   /* the cobs thread layout */
        else if (!strcmp(argv[i], "--avx2")) orc_set_simd_mode(1);
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 1; }
    }
    if (!index || !qfile) { fprintf(stderr, "query: -i and -f required\n"); return 1; }
    orc_index* idx = orc_index_read(index);
    if (!idx) { fprintf(stderr, "cannot read index %s\n", index); return 1; }
    size_t nq; rec_t* recs = read_fasta(qfile, &nq);
    if (!recs && nq) { fprintf(stderr, "cannot read %s\n", qfile); return 1; }
    if (threads > 1 && !over_slices) {
        /* queries over threads (the faster layout for many short queries), output kept in file
         * order: blocks of queries are scored in parallel into per-query text buffers */
        const size_t BLOCK = 2048;
        char** txt = (char**)calloc(BLOCK, sizeof(char*));
        size_t* tlen = (size_t*)calloc(BLOCK, sizeof(size_t));
        int bad = 0;
        for (size_t q0 = 0; q0 < nq && !bad; q0 += BLOCK) {
            size_t n = nq - q0 < BLOCK ? nq - q0 : BLOCK;
#pragma omp parallel num_threads(threads)
            {
                uint32_t* sc = (uint32_t*)malloc(sizeof(uint32_t) * (idx->n_docs + 1));
                orc_hit* hh = (orc_hit*)malloc(sizeof(orc_hit) * (idx->n_docs + 1));
#pragma omp for schedule(dynamic, 8)
                for (long i = 0; i < (long)n; i++) {
                    rec_t* r = &recs[q0 + (size_t)i];
                    txt[i] = NULL; tlen[i] = 0;
                    if (r->len == 0) continue;
                    int64_t K = orc_query_scores_sliced(idx, r->seq, r->len, sc, 1);
                    if (K < 0) { bad = 1; continue; }
                    uint32_t cnt = 0;
                    if (K > 0) cnt = orc_select(sc, idx->n_docs, orc_threshold_terms(thr, (uint32_t)K, floor_mode), hh);
                    size_t cap = strlen(r->name) + 32;
                    for (uint32_t j = 0; j < cnt; j++) cap += strlen(idx->doc_names[hh[j].doc]) + 16;
                    char* b = (char*)malloc(cap);
                    size_t w = (size_t)sprintf(b, "*%s\t%u\n", r->name, cnt);
                    for (uint32_t j = 0; j < cnt; j++)
                        w += (size_t)sprintf(b + w, "%s\t%u\n", idx->doc_names[hh[j].doc], hh[j].score);
                    txt[i] = b; tlen[i] = w;
                }
                free(sc); free(hh);
            }
            for (size_t i = 0; i < n; i++)
                if (txt[i]) { fwrite(txt[i], 1, tlen[i], stdout); free(txt[i]); }
        }
        if (bad) { fprintf(stderr, "invalid letter in a query\n"); return 1; }
        return 0;
    }
    uint32_t* scores = (uint32_t*)malloc(sizeof(uint32_t) * (idx->n_docs + 1));
    orc_hit* hits = (orc_hit*)malloc(sizeof(orc_hit) * (idx->n_docs + 1));
    for (size_t q = 0; q < nq; q++) {
        /* cobs's process_query only runs records with a non-empty sequence */
        if (recs[q].len == 0) continue;
        int64_t K = orc_query_scores_sliced(idx, recs[q].seq, recs[q].len, scores, threads);
        if (K < 0) { fprintf(stderr, "invalid letter in query %s\n", recs[q].name); return 1; }
        uint32_t n = 0;
        if (K > 0) {
            uint32_t T = orc_threshold_terms(thr, (uint32_t)K, floor_mode);
            n = orc_select(scores, idx->n_docs, T, hits);
        }
        printf("*%s\t%u\n", recs[q].name, n);
        for (uint32_t i = 0; i < n; i++) printf("%s\t%u\n", idx->doc_names[hits[i].doc], hits[i].score);
    }
    return 0;
}

static int cmd_construct(int argc, char** argv) {
    const char* out = NULL; const char* docs = NULL;
    uint32_t k = 31; uint64_t nh = 1; double fpr = 0.3; uint8_t canon = 1;
    for (int i = 0; i < argc; i++) {
        if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "-k") && i + 1 < argc) k = (uint32_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--num-hashes") && i + 1 < argc) nh = (uint64_t)atoll(argv[++i]);
        else if (!strcmp(argv[i], "--false-positive-rate") && i + 1 < argc) fpr = strtod(argv[++i], NULL);
        else if (!strcmp(argv[i], "--no-canonicalize")) canon = 0;
        else docs = argv[i];
    }
    if (!out || !docs) { fprintf(stderr, "construct: -o OUT DOCS.fa required\n"); return 1; }
    size_t nd; rec_t* recs = read_fasta(docs, &nd);
    if (!recs) { fprintf(stderr, "cannot read %s\n", docs); return 1; }
    uint64_t max_kmers = 1;
    const char** names = (const char**)malloc(sizeof(char*) * nd);
    for (size_t d = 0; d < nd; d++) {
        names[d] = recs[d].name;
        if (recs[d].len >= k && recs[d].len - k + 1 > max_kmers) max_kmers = recs[d].len - k + 1;
    }
    orc_index* idx = orc_index_new(k, canon, (uint32_t)nd, orc_signature_size(max_kmers, nh, fpr), nh, names);
    for (size_t d = 0; d < nd; d++) orc_index_add_doc(idx, (uint32_t)d, recs[d].seq ? recs[d].seq : "", recs[d].len);
    return orc_index_write(idx, out);
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "query")) return cmd_query(argc - 2, argv + 2);
    if (argc >= 2 && !strcmp(argv[1], "construct")) return cmd_construct(argc - 2, argv + 2);
    fprintf(stderr, "usage: cobs_oracle {query|construct} ...\n");
    return 1;
}
