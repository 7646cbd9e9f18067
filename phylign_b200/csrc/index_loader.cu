// index_loader.cu -- decompressed index file -> HBM at PCIe speed.
//
// The reference keeps decompressed indexes on disk when `keep_cobs_indexes` is set
// ({decompression_dir}/{batch}.cobs_classic, /root/reference/Snakefile:364-387, config.yaml:134) and
// `cobs query --load-complete` then read()s the whole body into RAM.  Here the body goes from the
// file straight to the GPU: reader threads pread() consecutive chunks into a ring of page-locked
// slots, the calling thread DMAs each filled slot to a device staging buffer and launches the
// re-stride kernel behind it, all on a dedicated upload stream (so a match running on the context's
// main stream is not blocked).  No per-chunk host synchronisation: a slot is reused once the event
// recorded behind its DMA has completed.
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "phy_internal.cuh"

int phy_restride_chunk_on(phy_ctx* ctx, HostIndex& ix, const uint8_t* d_src, uint64_t body_off, uint64_t nbytes,
                          cudaStream_t st);

namespace {
constexpr size_t LD_CHUNK = 4u << 20;

int ensure_loader(phy_ctx* ctx) {
    if (ctx->ld_ready) return PHY_OK;
    PHY_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
    for (int i = 0; i < PHY_LD_SLOTS; i++) {
        PHY_CUDA(ctx, cudaHostAlloc((void**)&ctx->ld_pin[i], LD_CHUNK, cudaHostAllocDefault));
        PHY_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ld_ev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; i++) PHY_CUDA(ctx, cudaMalloc((void**)&ctx->ld_stage[i], LD_CHUNK));
    ctx->ld_ready = true;
    return PHY_OK;
}
}  // namespace

void phy_loader_destroy(phy_ctx* ctx) {
    if (!ctx->ld_ready) return;
    cudaStreamSynchronize(ctx->up_stream);
    for (int i = 0; i < PHY_LD_SLOTS; i++) {
        if (ctx->ld_pin[i]) cudaFreeHost(ctx->ld_pin[i]);
        if (ctx->ld_ev[i]) cudaEventDestroy(ctx->ld_ev[i]);
    }
    for (int i = 0; i < 2; i++)
        if (ctx->ld_stage[i]) cudaFree(ctx->ld_stage[i]);
    cudaStreamDestroy(ctx->up_stream);
    ctx->ld_ready = false;
}

extern "C" int phy_index_load_file(phy_ctx* ctx, int idx_id, const char* path, uint64_t body_offset, int n_threads) {
    if (!ctx || !path || idx_id < 0 || (size_t)idx_id >= ctx->idx.size() || !ctx->idx[idx_id].alive) {
        phy_set_error(ctx, "phy_index_load_file: unknown index id %d", idx_id);
        return PHY_ERR_ARG;
    }
    HostIndex& ix = ctx->idx[idx_id];
    if (ix.committed || ix.pushed) {
        phy_set_error(ctx, "index %d already holds data", idx_id);
        return PHY_ERR_STATE;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    PHY_TRY(ensure_loader(ctx));
    const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) {
        phy_set_error(ctx, "cannot open %s: %s", path, strerror(errno));
        return PHY_ERR_IO;
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || (uint64_t)st.st_size != body_offset + ix.body_bytes) {
        phy_set_error(ctx, "%s: file holds %llu bytes, expected header %llu + signature_size*row_size %llu", path,
                      (unsigned long long)st.st_size, (unsigned long long)body_offset,
                      (unsigned long long)ix.body_bytes);
        ::close(fd);
        return PHY_ERR_ARG;
    }
    const uint64_t total = ix.body_bytes;
    const uint64_t n_chunks = (total + LD_CHUNK - 1) / LD_CHUNK;
    n_threads = std::max(1, std::min(n_threads, PHY_LD_SLOTS));
    // chunk c lives in slot c % PHY_LD_SLOTS.  filled[c]: bytes are in the slot; issued[c]: its DMA is queued
    // and ld_ev[slot] recorded behind it (the reader of chunk c + PHY_LD_SLOTS waits for that event).
    std::mutex mu;
    std::condition_variable cv;
    std::vector<uint8_t> filled(n_chunks, 0), issued(n_chunks, 0);
    std::atomic<uint64_t> next{0};
    std::atomic<int> io_err{0};
    const int dev = ctx->device;
    auto reader = [&]() {
        cudaSetDevice(dev);
        for (;;) {
            const uint64_t c = next.fetch_add(1);
            if (c >= n_chunks || io_err.load()) break;
            const int slot = (int)(c % PHY_LD_SLOTS);
            if (c >= (uint64_t)PHY_LD_SLOTS) {  // the slot's previous chunk must have left for the device
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return issued[c - PHY_LD_SLOTS] || io_err.load(); });
                lk.unlock();
                if (io_err.load()) break;
                cudaEventSynchronize(ctx->ld_ev[slot]);
            }
            const uint64_t off = c * LD_CHUNK, n = std::min<uint64_t>(LD_CHUNK, total - off);
            uint64_t got = 0;
            while (got < n) {
                ssize_t r = ::pread(fd, ctx->ld_pin[slot] + got, n - got, (off_t)(body_offset + off + got));
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) { io_err = errno ? errno : EIO; break; }
                got += (uint64_t)r;
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                filled[c] = 1;
            }
            cv.notify_all();
        }
        cv.notify_all();
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(reader);
    int rc = PHY_OK;
    cudaError_t ce = cudaSuccess;
    for (uint64_t c = 0; c < n_chunks && rc == PHY_OK; c++) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return filled[c] || io_err.load(); });
        }
        if (io_err.load()) break;
        const int slot = (int)(c % PHY_LD_SLOTS);
        const uint64_t off = c * LD_CHUNK, n = std::min<uint64_t>(LD_CHUNK, total - off);
        uint8_t* stage = ctx->ld_stage[c & 1];
        ce = cudaMemcpyAsync(stage, ctx->ld_pin[slot], n, cudaMemcpyHostToDevice, ctx->up_stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(ctx->ld_ev[slot], ctx->up_stream);
        if (ce != cudaSuccess) { rc = PHY_ERR_CUDA; io_err = EIO; }
        else rc = phy_restride_chunk_on(ctx, ix, stage, off, n, ctx->up_stream);
        if (rc != PHY_OK) io_err = EIO;
        {
            std::lock_guard<std::mutex> lk(mu);
            issued[c] = 1;
        }
        cv.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        if (io_err.load()) std::fill(issued.begin(), issued.end(), 1);   // release every waiting reader
    }
    cv.notify_all();
    for (auto& th : pool) th.join();
    ::close(fd);
    if (rc == PHY_OK && ce != cudaSuccess) rc = PHY_ERR_CUDA;
    if (rc == PHY_ERR_CUDA && ce != cudaSuccess) phy_set_error(ctx, "index upload failed: %s", cudaGetErrorString(ce));
    if (rc == PHY_OK && io_err.load()) {
        phy_set_error(ctx, "reading %s failed: %s", path, strerror(io_err.load()));
        rc = PHY_ERR_IO;
    }
    if (rc != PHY_OK) {
        cudaStreamSynchronize(ctx->up_stream);
        return rc;
    }
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->up_stream));
    ix.pushed = total;
    ctx->h2d_bytes += total;
    return PHY_OK;
}
