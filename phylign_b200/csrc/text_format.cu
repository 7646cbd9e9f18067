// text_format.cu -- host-side formatters of the two text protocols of the match stage
// (plain C++; lives in the library so the Python driver has no per-line loops):
//   * `cobs query` stdout / intermediate/03_match content (SURVEY.md 3.2;
//     /root/reference/scripts/postprocess_cobs.py:16-18 for the "_<accession>" name form)
//   * the 04_filter FASTA of /root/reference/scripts/filter_queries.py:152-156,195-199
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "phy_internal.cuh"

namespace {
inline void put_u32(std::string& s, uint32_t v) {
    char buf[12];
    int n = 0;
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) s.push_back(buf[--n]);
}
char* to_c(const std::string& s, uint64_t* len) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    *len = s.size();
    return p;
}
}  // namespace

extern "C" int phy_format_cobs_text(const phy_results* r, uint32_t idx_id, const char* headers, const uint64_t* hoffs,
                                    const uint8_t* skip, const char* names, const uint64_t* noffs, uint32_t n_docs,
                                    int strip_prefix, char** out, uint64_t* out_len) {
    if (!r || !headers || !hoffs || !names || !noffs || !out || !out_len) return PHY_ERR_ARG;
    const phy_unit* lo = std::lower_bound(r->units, r->units + r->n_units, idx_id,
                                          [](const phy_unit& u, uint32_t i) { return u.index < i; });
    const phy_unit* hi = lo;
    while (hi < r->units + r->n_units && hi->index == idx_id) hi++;
    std::string s;
    uint64_t est = 0;
    for (const phy_unit* u = lo; u < hi; u++) est += (uint64_t)u->n_kept * 24;
    s.reserve(est + (hoffs[r->n_queries] - hoffs[0]) + (uint64_t)r->n_queries * 8 + 64);
    const phy_unit* u = lo;
    for (uint32_t q = 0; q < r->n_queries; q++) {
        while (u < hi && u->query < q) u++;
        if (skip && skip[q]) continue;   // record without sequence: cobs never runs it
        s.push_back('*');
        s.append(headers + hoffs[q], hoffs[q + 1] - hoffs[q]);
        s.push_back('\t');
        if (u < hi && u->query == q) {
            put_u32(s, u->n_pass);
            s.push_back('\n');
            const phy_hit* h = r->hits + u->offset;
            for (uint32_t i = 0; i < u->n_kept; i++) {
                if (h[i].doc >= n_docs) return PHY_ERR_ARG;
                const char* nm = names + noffs[h[i].doc];
                size_t len = noffs[h[i].doc + 1] - noffs[h[i].doc];
                if (strip_prefix) {  // "_" + text after the first underscore
                    const char* us = (const char*)memchr(nm, '_', len);
                    s.push_back('_');
                    if (us) s.append(us + 1, len - (size_t)(us + 1 - nm));
                } else {
                    s.append(nm, len);
                }
                s.push_back('\t');
                put_u32(s, h[i].score);
                s.push_back('\n');
            }
        } else {
            s.append("0\n");
        }
    }
    *out = to_c(s, out_len);
    return *out ? PHY_OK : PHY_ERR_NOMEM;
}

extern "C" int phy_format_filter_fasta(const phy_merged* m, const char* qnames, const uint64_t* qnoffs,
                                       const char* seqs, const uint64_t* soffs, uint32_t n_batches,
                                       const char* const* ref_names, const uint64_t* const* ref_offs,
                                       const uint32_t* ref_counts, char** out, uint64_t* out_len) {
    if (!m || !qnames || !qnoffs || !seqs || !soffs || !out || !out_len) return PHY_ERR_ARG;
    std::string s;
    s.reserve((soffs[m->n_queries] - soffs[0]) + (qnoffs[m->n_queries] - qnoffs[0]) + m->offs[m->n_queries] * 16 +
              (uint64_t)m->n_queries * 4 + 64);
    for (uint32_t q = 0; q < m->n_queries; q++) {
        s.push_back('>');
        s.append(qnames + qnoffs[q], qnoffs[q + 1] - qnoffs[q]);
        s.push_back(' ');
        for (uint64_t i = m->offs[q]; i < m->offs[q + 1]; i++) {
            const phy_cand& c = m->cands[i];
            if (c.batch_rank >= n_batches || c.doc >= ref_counts[c.batch_rank]) return PHY_ERR_ARG;
            const uint64_t* ro = ref_offs[c.batch_rank];
            if (i > m->offs[q]) s.push_back(',');
            s.append(ref_names[c.batch_rank] + ro[c.doc], ro[c.doc + 1] - ro[c.doc]);
        }
        s.push_back('\n');
        s.append(seqs + soffs[q], soffs[q + 1] - soffs[q]);
        s.push_back('\n');
    }
    *out = to_c(s, out_len);
    return *out ? PHY_OK : PHY_ERR_NOMEM;
}

extern "C" void phy_text_free(char* p) { free(p); }
