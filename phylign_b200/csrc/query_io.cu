// query_io.cu -- host-side query file reader and 04_filter writer (plain C++, no GPU work).
//
//  * phy_fasta_read: the record rules of `cobs query -f`: a line starting with
//    '>' or ';' opens a record, the following lines are concatenated, empty lines are skipped, records
//    without sequence are dropped.  Everything lands in four flat arrays (sequences + offsets, header
//    lines + offsets) so the driver has no per-record Python objects: the sequence block goes to
//    phy_queries_set as it is (page-locked when a GPU is present), the header block to the match-file
//    writer.  `simple` tells the caller that the file is plain '>' FASTA without ';', '@' or '+' record
//    lines and without empty records -- then readfq (/root/reference/scripts/filter_queries.py:69-102)
//    sees exactly the same records, with name = header up to the first blank.
//  * phy_write_filter_fasta: ">{qname} {ref1,ref2,...}\n{seq}\n" of filter_queries.py:152-156,195-199
//    straight into intermediate/04_filter/{qfile}.fa (tmp + rename).
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "phy_internal.cuh"

void* phy_pinned_alloc(size_t bytes);
void phy_pinned_free(void* p);

namespace {
bool read_file(const char* path, std::vector<char>& buf) {
    int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return false;
    struct stat st;
    size_t hint = (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) ? (size_t)st.st_size : 0;
    buf.resize(hint ? hint + 1 : (1u << 20));  // +1: the read that reports EOF needs no regrowth
    size_t n = 0;
    for (;;) {
        if (n == buf.size()) buf.resize(buf.size() * 2);
        ssize_t r = ::read(fd, buf.data() + n, buf.size() - n);
        if (r < 0) {
            if (errno == EINTR) continue;
            ::close(fd);
            return false;
        }
        if (r == 0) break;
        n += (size_t)r;
    }
    ::close(fd);
    buf.resize(n);
    return true;
}
}  // namespace

extern "C" void phy_fasta_free(phy_fasta* f) {
    if (!f) return;
    if (f->seqs_pinned) phy_pinned_free(f->seqs); else free(f->seqs);
    free(f->soffs); free(f->headers); free(f->hoffs); free(f->name_len);
    free(f);
}

extern "C" int phy_fasta_read(const char* path, phy_fasta** out) {
    if (!path || !out) return PHY_ERR_ARG;
    *out = nullptr;
    std::vector<char> buf;
    if (!read_file(path, buf)) {
        phy_set_error(nullptr, "cannot read %s: %s", path, strerror(errno));
        return PHY_ERR_IO;
    }
    const size_t len = buf.size();
    const char* t = buf.data();
    // pass 1: count records and bytes (upper bounds)
    size_t n_rec = 0;
    for (size_t p = 0; p < len;) {
        if (t[p] == '>' || t[p] == ';') n_rec++;
        const char* nl = (const char*)memchr(t + p, '\n', len - p);
        p = nl ? (size_t)(nl - t) + 1 : len;
    }
    phy_fasta* f = (phy_fasta*)calloc(1, sizeof(phy_fasta));
    if (!f) return PHY_ERR_NOMEM;
    // plain host memory: page-locking ~100 MB costs more than the one staged upload it would save
    f->seqs = (char*)malloc(len + 64);
    f->seqs_pinned = 0;
    f->soffs = (uint64_t*)malloc((n_rec + 1) * sizeof(uint64_t));
    f->headers = (char*)malloc(len + 1);
    f->hoffs = (uint64_t*)malloc((n_rec + 1) * sizeof(uint64_t));
    f->name_len = (uint32_t*)malloc((n_rec + 1) * sizeof(uint32_t));
    if (!f->seqs || !f->soffs || !f->headers || !f->hoffs || !f->name_len) {
        phy_fasta_free(f);
        return PHY_ERR_NOMEM;
    }
    // pass 2
    uint64_t ns = 0, nh = 0;
    uint32_t n = 0;
    bool open_rec = false, simple = true;
    uint64_t rec_seq0 = 0, rec_h0 = 0;
    auto close_rec = [&]() {
        if (!open_rec) return;
        if (ns == rec_seq0) {  // no sequence: cobs never runs the record -> dropped
            nh = rec_h0;
            simple = false;
        } else {
            n++;
        }
        open_rec = false;
    };
    for (size_t p = 0; p < len;) {
        const char* nl = (const char*)memchr(t + p, '\n', len - p);
        size_t e = nl ? (size_t)(nl - t) : len, next = nl ? e + 1 : len;
        if (!nl) simple = false;                    // no trailing newline: readfq chops the last character
        while (e > p && (t[e - 1] == '\r')) { e--; simple = false; }
        if (e > p) {
            if (t[p] == '>' || t[p] == ';') {
                close_rec();
                if (t[p] == ';') simple = false;
                open_rec = true;
                rec_seq0 = ns;
                rec_h0 = nh;
                f->soffs[n] = ns;
                f->hoffs[n] = nh;
                const size_t hl = e - (p + 1);
                memcpy(f->headers + nh, t + p + 1, hl);
                size_t k = 0;
                while (k < hl && t[p + 1 + k] != ' ') k++;   // qname = header.split(" ")[0] (filter_queries.py:59,80)
                for (size_t i = 0; i < k; i++)
                    if (t[p + 1 + i] == '\t') simple = false;  // readfq also cuts at tabs: let the caller decide
                f->name_len[n] = (uint32_t)k;
                nh += hl;
            } else if (open_rec) {
                if (t[p] == '@' || t[p] == '+') simple = false;  // readfq would read these as FASTQ structure
                memcpy(f->seqs + ns, t + p, e - p);
                ns += e - p;
            } else {
                simple = false;  // text before the first header
            }
        }
        p = next;
    }
    close_rec();
    f->soffs[n] = ns;
    f->hoffs[n] = nh;
    f->n = n;
    f->simple = simple ? 1 : 0;
    *out = f;
    return PHY_OK;
}

extern "C" int phy_write_filter_fasta(const char* final_path, const phy_merged* m, const char* headers,
                                      const uint64_t* hoffs, const uint32_t* name_len, const char* seqs,
                                      const uint64_t* soffs, uint32_t n_batches, const char* const* ref_names,
                                      const uint64_t* const* ref_offs, const uint32_t* ref_counts,
                                      uint32_t q_begin, uint32_t q_end, int append, uint64_t* file_bytes) {
    if (!final_path || !m || !headers || !hoffs || !name_len || !seqs || !soffs) return PHY_ERR_ARG;
    q_end = std::min(q_end, m->n_queries);
    // append != 0: the caller owns the (temporary) file and its final rename; blocks are appended in order
    std::string tmp = append ? std::string(final_path)
                             : std::string(final_path) + ".tmp." + std::to_string((long)getpid());
    int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_CLOEXEC | (append ? O_APPEND : O_TRUNC), 0644);
    if (fd < 0) {
        phy_set_error(nullptr, "cannot create %s: %s", tmp.c_str(), strerror(errno));
        return PHY_ERR_IO;
    }
    std::vector<char> buf;
    buf.reserve(8u << 20);
    uint64_t total = 0;
    bool ok = true;
    auto flush = [&]() {
        size_t off = 0;
        while (ok && off < buf.size()) {
            ssize_t w = ::write(fd, buf.data() + off, buf.size() - off);
            if (w < 0) { if (errno == EINTR) continue; ok = false; break; }
            off += (size_t)w;
        }
        total += buf.size();
        buf.clear();
    };
    auto put = [&](const char* p, size_t n) { buf.insert(buf.end(), p, p + n); };
    int rc = PHY_OK;
    for (uint32_t q = q_begin; q < q_end && ok; q++) {
        buf.push_back('>');
        put(headers + hoffs[q], name_len[q]);
        buf.push_back(' ');
        for (uint64_t i = m->offs[q]; i < m->offs[q + 1]; i++) {
            const phy_cand& c = m->cands[i];
            if (c.batch_rank >= n_batches || c.doc >= ref_counts[c.batch_rank]) { rc = PHY_ERR_ARG; ok = false; break; }
            const uint64_t* ro = ref_offs[c.batch_rank];
            if (i > m->offs[q]) buf.push_back(',');
            put(ref_names[c.batch_rank] + ro[c.doc], ro[c.doc + 1] - ro[c.doc]);
        }
        buf.push_back('\n');
        put(seqs + soffs[q], soffs[q + 1] - soffs[q]);
        buf.push_back('\n');
        if (buf.size() >= (4u << 20)) flush();
    }
    if (ok) flush();
    ok = (::close(fd) == 0) && ok;
    if (ok && !append && ::rename(tmp.c_str(), final_path) != 0) ok = false;
    if (!ok) {
        if (rc == PHY_OK) {
            phy_set_error(nullptr, "cannot write %s: %s", final_path, strerror(errno));
            rc = PHY_ERR_IO;
        } else {
            phy_set_error(nullptr, "merged candidate refers to an unknown batch/document");
        }
        if (!append) ::unlink(tmp.c_str());
        return rc;
    }
    if (file_bytes) *file_bytes = total;
    return PHY_OK;
}
