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
char* to_c(const std::string& s, uint64_t* len) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    *len = s.size();
    return p;
}
}  // namespace

// ---- formatting core shared by phy_format_cobs_text and the match-file writer --------------------
// Appends the cobs text of queries [q0,q1) for one index to `out`.  Capacity is computed up
// front from the header bytes and the kept-hit counts, so the inner loops are plain memcpy.
int phy_format_cobs_range(const phy_results* r, uint32_t idx_id, uint32_t q0, uint32_t q1, const char* headers,
                          const uint64_t* hoffs, const uint8_t* skip, const char* names, const uint64_t* noffs,
                          uint32_t n_docs, int strip_prefix, std::vector<char>& out, uint64_t* n_header_lines,
                          uint64_t* n_hit_lines) {
    const phy_unit* ulo = std::lower_bound(r->units, r->units + r->n_units, idx_id,
                                           [](const phy_unit& u, uint32_t i) { return u.index < i; });
    const phy_unit* uhi = std::upper_bound(ulo, (const phy_unit*)(r->units + r->n_units), idx_id,
                                           [](uint32_t i, const phy_unit& u) { return i < u.index; });
    const phy_unit* u = std::lower_bound(ulo, uhi, q0, [](const phy_unit& x, uint32_t q) { return x.query < q; });
    const phy_unit* uend = std::lower_bound(u, uhi, q1, [](const phy_unit& x, uint32_t q) { return x.query < q; });
    uint64_t max_name = 0;
    if (u < uend)
        for (uint32_t d = 0; d < n_docs; d++) max_name = std::max<uint64_t>(max_name, noffs[d + 1] - noffs[d]);
    uint64_t kept = 0;
    for (const phy_unit* x = u; x < uend; x++) kept += x->n_kept;
    const size_t base = out.size();
    out.resize(base + (hoffs[q1] - hoffs[q0]) + (uint64_t)(q1 - q0) * 14 + kept * (max_name + 13) + 16);
    char* w = out.data() + base;
    auto put = [&](uint32_t v) {
        char buf[12];
        int n = 0;
        do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (n) *w++ = buf[--n];
    };
    uint64_t nh = 0, nl = 0;
    for (uint32_t q = q0; q < q1; q++) {
        const bool mine = u < uend && u->query == q;
        const phy_unit* cur = u;
        if (mine) u++;
        if (skip && skip[q]) continue;   // record without sequence: cobs never runs it
        *w++ = '*';
        const size_t hl = hoffs[q + 1] - hoffs[q];
        memcpy(w, headers + hoffs[q], hl);
        w += hl;
        *w++ = '\t';
        nh++;
        if (!mine) {
            *w++ = '0';
            *w++ = '\n';
            continue;
        }
        put(cur->n_pass);
        *w++ = '\n';
        const phy_hit* h = r->hits + cur->offset;
        for (uint32_t i = 0; i < cur->n_kept; i++) {
            if (h[i].doc >= n_docs) return PHY_ERR_ARG;
            const char* nm = names + noffs[h[i].doc];
            size_t len = noffs[h[i].doc + 1] - noffs[h[i].doc];
            if (strip_prefix) {  // "_" + text after the first underscore (postprocess_cobs.py:16-18)
                const char* us = (const char*)memchr(nm, '_', len);
                *w++ = '_';
                if (us) {
                    const size_t rest = len - (size_t)(us + 1 - nm);
                    memcpy(w, us + 1, rest);
                    w += rest;
                }
            } else {
                memcpy(w, nm, len);
                w += len;
            }
            *w++ = '\t';
            put(h[i].score);
            *w++ = '\n';
        }
        nl += cur->n_kept;
    }
    out.resize((size_t)(w - out.data()));
    if (n_header_lines) *n_header_lines += nh;
    if (n_hit_lines) *n_hit_lines += nl;
    return PHY_OK;
}

extern "C" int phy_format_cobs_text(const phy_results* r, uint32_t idx_id, const char* headers, const uint64_t* hoffs,
                                    const uint8_t* skip, const char* names, const uint64_t* noffs, uint32_t n_docs,
                                    int strip_prefix, char** out, uint64_t* out_len) {
    if (!r || !headers || !hoffs || !names || !noffs || !out || !out_len) return PHY_ERR_ARG;
    std::vector<char> buf;
    PHY_TRY(phy_format_cobs_range(r, idx_id, 0, r->n_queries, headers, hoffs, skip, names, noffs, n_docs, strip_prefix,
                                  buf, nullptr, nullptr));
    char* p = (char*)malloc(buf.size() + 1);
    if (!p) return PHY_ERR_NOMEM;
    memcpy(p, buf.data(), buf.size());
    p[buf.size()] = 0;
    *out = p;
    *out_len = buf.size();
    return PHY_OK;
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

// ---- parser of match files (the reading side of the same protocol) -----------------------------
// Restates the rules of /root/reference/scripts/filter_queries.py:27-66 (`cobs_iterator`) natively:
// lines are stripped, empty lines skipped; "*<qname>[ comment]\t<int>" opens a block (qname = text up
// to the first space or tab); a hit line is "<name><ws><kmers>" with exactly two whitespace-separated
// fields, and <name> holds exactly ONE '_' (`rid, ref = tmp_name.split("_")`): ref = text after it.
namespace {
struct MatchText {
    std::vector<uint64_t> q_off, first_hit, ref_off;
    std::vector<uint32_t> q_len, ref_len, kmers;
};
inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
}  // namespace

extern "C" int phy_parse_match_text(const char* text, uint64_t len, phy_match_text** out) {
    if (!text || !out) return PHY_ERR_ARG;
    *out = nullptr;
    MatchText m;
    uint64_t line_no = 0, p = 0;
    while (p < len) {
        uint64_t e = p;
        while (e < len && text[e] != '\n') e++;
        uint64_t a = p, b = e;  // strip
        while (a < b && is_ws(text[a])) a++;
        while (b > a && is_ws(text[b - 1])) b--;
        p = e + 1;
        line_no++;
        if (a == b) continue;
        if (text[a] == '*') {
            uint64_t tab = a + 1;
            while (tab < b && text[tab] != '\t') tab++;
            if (tab == b) {
                phy_set_error(nullptr, "match file line %llu: header without a tab-separated count", (unsigned long long)line_no);
                return PHY_ERR_ARG;
            }
            uint64_t c = tab + 1;
            if (c == b) { phy_set_error(nullptr, "match file line %llu: empty count", (unsigned long long)line_no); return PHY_ERR_ARG; }
            uint64_t c2 = c;
            while (c2 < b && text[c2] != '\t') c2++;      // parts[1] of x[1:].split("\t")
            for (uint64_t i = c; i < c2; i++)
                if (text[i] < '0' || text[i] > '9') {
                    phy_set_error(nullptr, "match file line %llu: count is not an integer", (unsigned long long)line_no);
                    return PHY_ERR_ARG;
                }
            uint64_t qe = a + 1;
            while (qe < tab && text[qe] != ' ') qe++;   // qname = parts[0].split(" ")[0]
            m.q_off.push_back(a + 1);
            m.q_len.push_back((uint32_t)(qe - (a + 1)));
            m.first_hit.push_back(m.kmers.size());
        } else {
            if (m.q_off.empty()) {
                phy_set_error(nullptr, "match file line %llu: hit line before any query header", (unsigned long long)line_no);
                return PHY_ERR_ARG;
            }
            // exactly two whitespace-separated fields
            uint64_t n1 = a;
            while (n1 < b && !is_ws(text[n1])) n1++;
            uint64_t k0 = n1;
            while (k0 < b && is_ws(text[k0])) k0++;
            uint64_t k1 = k0;
            while (k1 < b && !is_ws(text[k1])) k1++;
            if (n1 == b || k0 == b || k1 != b) {
                phy_set_error(nullptr, "match file line %llu: expected '<name> <kmers>'", (unsigned long long)line_no);
                return PHY_ERR_ARG;
            }
            uint64_t us = a, n_us = 0, first_us = 0;
            for (; us < n1; us++)
                if (text[us] == '_') { if (!n_us) first_us = us; n_us++; }
            if (n_us != 1) {
                phy_set_error(nullptr, "match file line %llu: reference name must hold exactly one '_'", (unsigned long long)line_no);
                return PHY_ERR_ARG;
            }
            uint64_t v = 0;
            for (uint64_t i = k0; i < k1; i++) {
                if (text[i] < '0' || text[i] > '9' || v > 0xFFFFFFFFull) {
                    phy_set_error(nullptr, "match file line %llu: k-mer count is not an integer", (unsigned long long)line_no);
                    return PHY_ERR_ARG;
                }
                v = v * 10 + (uint64_t)(text[i] - '0');
            }
            m.ref_off.push_back(first_us + 1);
            m.ref_len.push_back((uint32_t)(n1 - first_us - 1));
            m.kmers.push_back((uint32_t)std::min<uint64_t>(v, 0xFFFFFFFFull));
        }
    }
    if (m.q_off.empty()) {
        phy_set_error(nullptr, "empty match file");
        return PHY_ERR_ARG;
    }
    m.first_hit.push_back(m.kmers.size());
    phy_match_text* r = (phy_match_text*)calloc(1, sizeof(phy_match_text));
    if (!r) return PHY_ERR_NOMEM;
    r->n_blocks = m.q_off.size();
    r->n_hits = m.kmers.size();
    auto dup = [](const void* src, size_t bytes) -> void* {
        void* p = malloc(bytes ? bytes : 1);
        if (p && bytes) memcpy(p, src, bytes);
        return p;
    };
    r->q_off = (uint64_t*)dup(m.q_off.data(), m.q_off.size() * 8);
    r->q_len = (uint32_t*)dup(m.q_len.data(), m.q_len.size() * 4);
    r->first_hit = (uint64_t*)dup(m.first_hit.data(), m.first_hit.size() * 8);
    r->ref_off = (uint64_t*)dup(m.ref_off.data(), m.ref_off.size() * 8);
    r->ref_len = (uint32_t*)dup(m.ref_len.data(), m.ref_len.size() * 4);
    r->kmers = (uint32_t*)dup(m.kmers.data(), m.kmers.size() * 4);
    if (!r->q_off || !r->q_len || !r->first_hit || !r->ref_off || !r->ref_len || !r->kmers) {
        phy_match_text_free(r);
        return PHY_ERR_NOMEM;
    }
    *out = r;
    return PHY_OK;
}

extern "C" void phy_match_text_free(phy_match_text* r) {
    if (!r) return;
    free(r->q_off); free(r->q_len); free(r->first_hit); free(r->ref_off); free(r->ref_len); free(r->kmers);
    free(r);
}
