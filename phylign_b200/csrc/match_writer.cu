// match_writer.cu -- host-side writer of the per-batch match files (plain C++, no GPU work).
//
// Replaces the tail of the reference's per-batch pipeline
//     cobs query ... | postprocess_cobs.py -n N | gzip --fast > intermediate/03_match/{batch}____{qfile}.gz
// (/root/reference/Snakefile:425-427, 467-469, 482-484): the device results of one query block are
// formatted (text_format.cu) and deflated (zlib level 1 = `gzip --fast`) by a pool of host threads, one
// gzip member per (file, query range) task, and appended to the files in query order.  A file of
// concatenated gzip members is a valid gzip stream (RFC 1952 2.2): `gzip -d`, Python's gzip module and
// xopen (filter_queries.py:13,40) read it as one text.  Files are written as <path>.tmp.<pid> and renamed
// on commit, so a failed job leaves nothing partial at the output path (SURVEY.md 8(b) error convention).
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "phy_internal.cuh"

int phy_format_cobs_range(const phy_results* r, uint32_t idx_id, uint32_t q0, uint32_t q1, const char* headers,
                          const uint64_t* hoffs, const uint8_t* skip, const char* names, const uint64_t* noffs,
                          uint32_t n_docs, int strip_prefix, std::vector<char>& out, uint64_t* n_header_lines,
                          uint64_t* n_hit_lines);

struct phy_mfile {
    int fd = -1;
    int gzip_level = 1;
    std::string tmp_path, final_path;
    uint64_t bytes = 0;
};

namespace {
using clk = std::chrono::steady_clock;
inline double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

bool write_all(int fd, const void* p, size_t n) {
    const char* s = (const char*)p;
    while (n) {
        ssize_t w = ::write(fd, s, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        s += w;
        n -= (size_t)w;
    }
    return true;
}

// one gzip member holding `n` bytes of text; `out` is scratch that keeps its capacity between calls
// (the caller copies the few bytes produced: fresh worst-case buffers per task would page-fault
// ~4x the text size on the first block of a run)
bool gzip_member(const char* text, size_t n, int level, std::vector<unsigned char>& out, size_t* out_len) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    const size_t bound = deflateBound(&zs, (uLong)n) + 32;
    if (out.size() < bound) out.resize(bound);
    size_t done = 0, produced = 0;
    int rc = Z_OK;
    do {  // avail_in is 32 bits wide: feed at most 1 GiB per call
        const size_t chunk = std::min<size_t>(n - done, (size_t)1 << 30);
        zs.next_in = (Bytef*)(text + done);
        zs.avail_in = (uInt)chunk;
        done += chunk;
        do {
            if (produced == out.size()) out.resize(out.size() * 2);
            zs.next_out = out.data() + produced;
            const size_t room = std::min<size_t>(out.size() - produced, (size_t)1 << 30);
            zs.avail_out = (uInt)room;
            rc = deflate(&zs, done == n ? Z_FINISH : Z_NO_FLUSH);
            produced += room - zs.avail_out;
        } while (rc == Z_OK && zs.avail_out == 0);
    } while (rc == Z_OK && done < n);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    *out_len = produced;
    return true;
}
}  // namespace

extern "C" int phy_mfile_open(const char* final_path, int gzip_level, phy_mfile** out) {
    if (!final_path || !out || gzip_level < 0 || gzip_level > 9) return PHY_ERR_ARG;
    *out = nullptr;
    phy_mfile* f = new phy_mfile();
    f->final_path = final_path;
    f->tmp_path = f->final_path + ".tmp." + std::to_string((long)getpid());
    f->gzip_level = gzip_level;
    f->fd = ::open(f->tmp_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    if (f->fd < 0) {
        phy_set_error(nullptr, "cannot create %s: %s", f->tmp_path.c_str(), strerror(errno));
        delete f;
        return PHY_ERR_IO;
    }
    *out = f;
    return PHY_OK;
}

extern "C" void phy_mfile_abort(phy_mfile* f) {
    if (!f) return;
    if (f->fd >= 0) ::close(f->fd);
    ::unlink(f->tmp_path.c_str());
    delete f;
}

extern "C" int phy_mfile_commit(phy_mfile* f, uint64_t* file_bytes) {
    if (!f) return PHY_ERR_ARG;
    bool ok = true;
    if (f->bytes == 0 && f->gzip_level > 0) {  // no block was written: still a valid (empty) gzip file
        std::vector<unsigned char> z;
        size_t zn = 0;
        ok = gzip_member("", 0, f->gzip_level, z, &zn) && write_all(f->fd, z.data(), zn);
        f->bytes += zn;
    }
    ok = ok && ::close(f->fd) == 0;
    f->fd = -1;
    if (ok && ::rename(f->tmp_path.c_str(), f->final_path.c_str()) != 0) ok = false;
    if (!ok) {
        phy_set_error(nullptr, "cannot finish %s: %s", f->final_path.c_str(), strerror(errno));
        phy_mfile_abort(f);
        return PHY_ERR_IO;
    }
    if (file_bytes) *file_bytes = f->bytes;
    delete f;
    return PHY_OK;
}

extern "C" int phy_write_match_blocks(const phy_results* r, const phy_mfile_job* jobs, uint32_t n_jobs,
                                      const char* headers, const uint64_t* hoffs, const uint8_t* skip,
                                      int strip_prefix, int n_threads, phy_write_stats* stats) {
    if (!r || (!jobs && n_jobs) || !headers || !hoffs) return PHY_ERR_ARG;
    for (uint32_t j = 0; j < n_jobs; j++)
        if (!jobs[j].file || jobs[j].file->fd < 0 || !jobs[j].names || !jobs[j].noffs) return PHY_ERR_ARG;
    const auto t_wall0 = clk::now();
    const uint32_t nq = r->n_queries;
    n_threads = std::max(1, std::min(n_threads, 256));
    // tasks: (job, query range); enough of them to keep every thread busy, none smaller than 4096 queries
    uint32_t per_job = n_jobs ? (uint32_t)((4u * (uint32_t)n_threads + n_jobs - 1) / n_jobs) : 1;
    per_job = std::max(1u, std::min(per_job, std::max(1u, nq / 4096u)));
    struct Task {
        uint32_t job, q0, q1;
        std::vector<unsigned char> z;   // gzip member (or the plain text when gzip_level == 0)
        uint64_t text_bytes = 0, n_head = 0, n_hit = 0;
        double fmt_s = 0, def_s = 0;
        int rc = PHY_OK;
    };
    std::vector<Task> tasks;
    std::vector<size_t> job_first(n_jobs + 1, 0);
    for (uint32_t j = 0; j < n_jobs; j++) {
        job_first[j] = tasks.size();
        for (uint32_t t = 0; t < per_job; t++) {
            Task k;
            k.job = j;
            k.q0 = (uint32_t)((uint64_t)nq * t / per_job);
            k.q1 = (uint32_t)((uint64_t)nq * (t + 1) / per_job);
            if (k.q1 > k.q0) tasks.push_back(std::move(k));
        }
    }
    job_first[n_jobs] = tasks.size();
    std::atomic<size_t> next{0};
    auto work = [&]() {
        std::vector<char> text;
        std::vector<unsigned char> zbuf;
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= tasks.size()) break;
            Task& k = tasks[i];
            const phy_mfile_job& jb = jobs[k.job];
            text.clear();
            const auto t0 = clk::now();
            k.rc = phy_format_cobs_range(r, jb.idx_id, k.q0, k.q1, headers, hoffs, skip, jb.names, jb.noffs, jb.n_docs,
                                         strip_prefix, text, &k.n_head, &k.n_hit);
            const auto t1 = clk::now();
            k.fmt_s = secs(t0, t1);
            k.text_bytes = text.size();
            if (k.rc != PHY_OK) continue;
            if (jb.file->gzip_level > 0) {
                if (!text.empty()) {
                    size_t zn = 0;
                    if (!gzip_member(text.data(), text.size(), jb.file->gzip_level, zbuf, &zn)) k.rc = PHY_ERR_NOMEM;
                    else k.z.assign(zbuf.begin(), zbuf.begin() + (ptrdiff_t)zn);
                }
            } else {
                k.z.assign(text.begin(), text.end());
            }
            k.def_s = secs(t1, clk::now());
        }
    };
    {
        std::vector<std::thread> pool;
        const int nt = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, tasks.size()));
        for (int t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    }
    int rc = PHY_OK;
    for (auto& k : tasks)
        if (k.rc != PHY_OK) rc = k.rc;
    if (rc != PHY_OK) {
        phy_set_error(nullptr, "formatting a match block failed (%d)", rc);
        return rc;
    }
    // append: tasks are in (job, query) order; files are independent, so the appends run in parallel too
    std::atomic<uint32_t> next_job{0};
    std::atomic<int> io_err{0};
    std::vector<double> write_s((size_t)n_threads, 0.0);
    auto append = [&](int tid) {
        for (;;) {
            const uint32_t j = next_job.fetch_add(1);
            if (j >= n_jobs) break;
            const auto t0 = clk::now();
            for (size_t i = job_first[j]; i < job_first[j + 1]; i++) {
                if (!tasks[i].z.empty()) {
                    if (!write_all(jobs[j].file->fd, tasks[i].z.data(), tasks[i].z.size())) io_err = errno ? errno : EIO;
                    jobs[j].file->bytes += tasks[i].z.size();
                }
            }
            write_s[(size_t)tid] += secs(t0, clk::now());
        }
    };
    {
        std::vector<std::thread> pool;
        const int nt = (int)std::min<uint32_t>((uint32_t)n_threads, std::max(1u, n_jobs));
        for (int t = 1; t < nt; t++) pool.emplace_back(append, t);
        append(0);
        for (auto& th : pool) th.join();
    }
    if (io_err) {
        phy_set_error(nullptr, "writing a match file failed: %s", strerror(io_err));
        return PHY_ERR_IO;
    }
    if (stats) {
        for (auto& k : tasks) {
            stats->format_s += k.fmt_s;
            stats->deflate_s += k.def_s;
            stats->text_bytes += k.text_bytes;
            stats->file_bytes += k.z.size();
            stats->n_header_lines += k.n_head;
            stats->n_hit_lines += k.n_hit;
        }
        for (double w : write_s) stats->write_s += w;
        stats->wall_s += secs(t_wall0, clk::now());
    }
    return PHY_OK;
}
