// kmer_hash.cu -- K1: canonical k-mers of the queries and their XXH64 values.
//
// Replaces cobs `canonicalize_kmer` + `create_hashes` (SURVEY.md 8(a) a4, Appendix
// A.3/A.4; called from /root/reference/scripts/run_cobs_streaming.sh:24-29):
//   hash[j][g] = XXH64(canonical ASCII k-mer g, k, seed=j)
// The hash is index independent; `% signature_size` happens in the gather kernel.
// One thread per query k-mer; integer only; ~0.3% of the step time.
#include <algorithm>

#include "phy_internal.cuh"

namespace {

constexpr uint64_t XP1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t XP2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t XP3 = 0x165667B19E3779F9ULL;
constexpr uint64_t XP4 = 0x85EBCA77C2B2AE63ULL;
constexpr uint64_t XP5 = 0x27D4EB2F165667C5ULL;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xround(uint64_t acc, uint64_t in) {
    return rotl64(acc + in * XP2, 31) * XP1;
}

// XXH64 of the k (<32) bytes packed little-endian in w[0..3] (short-input path of the spec).
__device__ __forceinline__ uint64_t xxh64_packed(const uint64_t (&w)[4], const int k, uint64_t seed) {
    uint64_t h = seed + XP5 + (uint64_t)k;
    const int n8 = k >> 3;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (i < n8) {
            h ^= xround(0, w[i]);
            h = rotl64(h, 27) * XP1 + XP4;
        }
    }
    // remaining k & 7 bytes live in w[n8]
    uint64_t tail = n8 == 0 ? w[0] : (n8 == 1 ? w[1] : (n8 == 2 ? w[2] : w[3]));
    int rem = k & 7;
    if (rem >= 4) {
        h ^= (tail & 0xFFFFFFFFULL) * XP1;
        h = rotl64(h, 23) * XP2 + XP3;
        tail >>= 32;
        rem -= 4;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (i < rem) {
            h ^= (tail & 0xFFULL) * XP5;
            h = rotl64(h, 11) * XP1;
            tail >>= 8;
        }
    }
    h ^= h >> 33;
    h *= XP2;
    h ^= h >> 29;
    h *= XP3;
    h ^= h >> 32;
    return h;
}

// error word layout: bit 63 = an invalid letter was seen, low 32 bits = smallest query id
__device__ __forceinline__ void report_bad(unsigned long long* err, uint32_t q) {
    atomicMin(err, 0x8000000000000000ULL | (unsigned long long)q);
}

template <int KT>  // KT = 31 (fast path) or 0 (runtime k <= 31)
__global__ void __launch_bounds__(256) kmer_hash_kernel(
    const char* __restrict__ seq, const uint64_t* __restrict__ qoffs,
    const uint64_t* __restrict__ koffs, uint32_t nq, uint64_t total_kmers, int k_rt,
    int canonicalize, uint32_t num_hashes, uint64_t* __restrict__ hashes,
    unsigned long long* __restrict__ err) {
    const int k = KT ? KT : k_rt;
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_kmers) return;
    // query of k-mer g: last q with koffs[q] <= g
    uint32_t lo = 0, hi = nq;  // invariant: koffs[lo] <= g < koffs[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&koffs[mid]) <= g) lo = mid; else hi = mid;
    }
    const uint32_t q = lo;
    const char* s = seq + __ldg(&qoffs[q]) + (g - __ldg(&koffs[q]));

    uint64_t fwd = 0, rc = 0;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 31; i++) {
        if (i < k) {
            uint32_t c = (uint8_t)__ldg(&s[i]);
            uint32_t x = (c >> 1) & 3u;
            uint32_t code = x ^ (x >> 1);  // A0 C1 G2 T3
            bad |= ((0x54474341u >> (8 * code)) & 0xFFu) != c;
            fwd = (fwd << 2) | code;
            rc |= (uint64_t)(3u - code) << (2 * i);
        }
    }
    if (bad) {
        report_bad(err, q);
        return;
    }
    // ASCII order A<C<G<T equals the order of the MSB-first 2-bit packing
    uint64_t v = (canonicalize && rc < fwd) ? rc : fwd;
    uint64_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 31; i++) {
        if (i < k) {
            uint32_t code = (uint32_t)(v >> (2 * (k - 1 - i))) & 3u;
            uint64_t ch = (0x54474341u >> (8 * code)) & 0xFFu;
            w[i >> 3] |= ch << (8 * (i & 7));
        }
    }
    for (uint32_t j = 0; j < num_hashes; j++) hashes[(uint64_t)j * total_kmers + g] = xxh64_packed(w, k, j);
}

// k = 31 fast path: one thread takes 4 consecutive k-mers of one query.  Their 34 bases are read as ten
// aligned 32-bit words and packed to 2 bits four letters at a time (round 1 rolled byte by byte: ~170 of
// its ~390 instructions per k-mer were the per-base pack / validate / reverse-complement updates); the four
// k-mers are 62-bit windows of the packed stream, the reverse complement is a bit reversal; the canonical
// value becomes its 31 ASCII bytes through a 256-entry shared-memory table (4 bases -> 4 letters).
constexpr int KPT = 4;  // k-mers per thread
__global__ void __launch_bounds__(256) kmer_hash31_roll_kernel(
    const char* __restrict__ seq, const uint64_t* __restrict__ qoffs, const uint64_t* __restrict__ koffs,
    const uint64_t* __restrict__ ioffs, uint32_t nq, uint64_t total_items, uint64_t total_kmers,
    int canonicalize, uint32_t num_hashes, uint64_t* __restrict__ hashes, unsigned long long* __restrict__ err) {
    __shared__ uint32_t lut[256];
    {
        const uint32_t b = threadIdx.x;  // 4 bases MSB-first -> 4 ASCII bytes, first base in the low byte
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) v |= ((0x54474341u >> (8 * ((b >> (6 - 2 * i)) & 3u))) & 0xFFu) << (8 * i);
        lut[b] = v;
    }
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_items) return;
    // query of item t: ioffs[q] <= t < ioffs[q+1].  Reads mostly have similar lengths, so start from
    // the proportional guess and gallop outwards before bisecting (2-3 loads instead of log2(nq)).
    uint32_t lo, hi;
    {
        uint32_t g = (uint32_t)min((uint64_t)nq - 1, (uint64_t)((double)t * (double)nq / (double)total_items));
        uint32_t step = 1;
        if (__ldg(&ioffs[g]) <= t) {
            lo = g;
            hi = g + 1;
            while (hi < nq && __ldg(&ioffs[hi]) <= t) { lo = hi; step <<= 1; hi = min(nq, hi + step); }
        } else {
            hi = g;
            lo = g >= 1 ? g - 1 : 0;
            while (lo > 0 && __ldg(&ioffs[lo]) > t) { hi = lo; step <<= 1; lo = lo > step ? lo - step : 0; }
        }
    }
    while (hi - lo > 1) {  // ioffs[lo] <= t < ioffs[hi]
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&ioffs[mid]) <= t) lo = mid; else hi = mid;
    }
    const uint32_t q = lo;
    const uint64_t k0 = __ldg(&koffs[q]);
    const uint32_t K = (uint32_t)(__ldg(&koffs[q + 1]) - k0);
    const uint32_t i0 = (uint32_t)(t - __ldg(&ioffs[q])) * KPT;
    const uint32_t n = min((uint32_t)KPT, K - i0);
    // The 30 + n bases of this item, four at a time: aligned 32-bit loads re-aligned with a funnel shift,
    // the four 2-bit codes of a word extracted in parallel (A0 C1 G2 T3 from bits 2..1 of the letter) and
    // packed MSB-first with one multiply; a word is valid iff the table's letters for its codes equal it.
    constexpr int NG = (30 + KPT + 3) / 4;  // groups of 4 bases
    const uint64_t sbyte = __ldg(&qoffs[q]) + i0;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(seq + (sbyte & ~3ULL));  // (the buffer is padded by 64 B)
    const uint32_t sh = (uint32_t)(sbyte & 3ULL) * 8u;
    uint32_t raw[NG + 1];
#pragma unroll
    for (int i = 0; i <= NG; i++) raw[i] = __ldg(wp + i);
    const uint32_t valid_len = 30u + n;
    uint64_t p_hi = 0;
    uint32_t g_last = 0;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < NG; i++) {
        const uint32_t w = __funnelshift_r(raw[i], raw[i + 1], sh);  // letters 4i .. 4i+3, first in the low byte
        const uint32_t x = (w >> 1) & 0x03030303u;
        const uint32_t code = x ^ ((x >> 1) & 0x01010101u);
        const uint32_t g = (code * 0x40100401u) >> 24;             // c0<<6 | c1<<4 | c2<<2 | c3
        const uint32_t nv = valid_len > 4u * i ? min(4u, valid_len - 4u * i) : 0u;
        const uint32_t vm = nv >= 4u ? 0xFFFFFFFFu : ((1u << (8u * nv)) - 1u);
        bad |= ((lut[g] ^ w) & vm) != 0u;
        if (i < 8) p_hi |= (uint64_t)g << (56 - 8 * i);
        else g_last = g;
    }
    if (bad) {
        report_bad(err, q);
        return;
    }
    const uint64_t mask = (1ULL << 62) - 1;
#pragma unroll
    for (int m = 0; m < KPT; m++) {
        if (m < (int)n) {
            const uint64_t win = m == 0 ? p_hi : ((p_hi << (2 * m)) | ((uint64_t)g_last >> (8 - 2 * m)));  // bases m .. m+31
            const uint64_t fwd = win >> 2;                                                                  // bases m .. m+30
            // reverse complement: complement every 2-bit code, reverse the order of the 31 codes
            uint64_t r = __brevll(~fwd & mask);
            r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
            const uint64_t rc = r >> 2;
            const uint64_t v = ((canonicalize && rc < fwd) ? rc : fwd) << 2;  // 31 bases + 1 pad = 8 groups of 4
            uint64_t w[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                w[j] = (uint64_t)lut[(v >> (56 - 16 * j)) & 0xFF] | ((uint64_t)lut[(v >> (48 - 16 * j)) & 0xFF] << 32);
            w[3] &= 0x00FFFFFFFFFFFFFFULL;  // drop the pad letter
            const uint64_t g = k0 + i0 + (uint32_t)m;
            for (uint32_t j = 0; j < num_hashes; j++) hashes[(uint64_t)j * total_kmers + g] = xxh64_packed(w, 31, j);
        }
    }
}

// rule fix_query on the bases (Snakefile:326-332: `seqtk seq -U` + awk gsub(/[^ACGT]/,"A")): 16 bytes
// per thread, upper-case by clearing bit 5, everything that is not then A/C/G/T becomes 'A'
__device__ __forceinline__ uint32_t fix4(uint32_t w) {
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t c = ((w >> (8 * i)) & 0xFFu) & 0xDFu;
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') c = 'A';
        out |= c << (8 * i);
    }
    return out;
}
__global__ void __launch_bounds__(256) fix_bases_kernel(uint8_t* __restrict__ p, uint64_t n) {
    const uint64_t n16 = n / 16;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v = reinterpret_cast<uint4*>(p)[i];
        v.x = fix4(v.x); v.y = fix4(v.y); v.z = fix4(v.z); v.w = fix4(v.w);
        reinterpret_cast<uint4*>(p)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 15)) {
        uint32_t c = p[n16 * 16 + threadIdx.x] & 0xDFu;
        p[n16 * 16 + threadIdx.x] = (c != 'A' && c != 'C' && c != 'G' && c != 'T') ? 'A' : (uint8_t)c;
    }
}

}  // namespace

int phy_launch_fix_bases(phy_ctx* ctx, uint8_t* d_bases, uint64_t n) {
    if (n == 0) return PHY_OK;
    unsigned blocks = (unsigned)std::min<uint64_t>((n / 16 + 255) / 256 + 1, 148ull * 16);
    fix_bases_kernel<<<blocks, 256, 0, ctx->stream>>>(d_bases, n);
    ctx->launches++;
    PHY_CUDA(ctx, cudaGetLastError());
    return PHY_OK;
}

int phy_launch_hash(phy_ctx* ctx) {
    if (ctx->total_kmers == 0) {
        ctx->hashes_valid = true;
        return PHY_OK;
    }
    const uint32_t nh = ctx->q_num_hashes;
    PHY_TRY(phy_ensure(ctx, ctx->d_hashes, ctx->total_kmers * nh));
    PHY_TRY(phy_ensure(ctx, ctx->d_counters, 8));
    PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p + 2, 0xFF, sizeof(unsigned long long), ctx->stream));
    const uint64_t nblk = (ctx->total_kmers + 255) / 256;
    if (nblk > 0x7FFFFFFFull) {
        phy_set_error(ctx, "too many k-mers in one query block");
        return PHY_ERR_ARG;
    }
    if (ctx->q_term_size == 31) {
        // work items = runs of KPT consecutive k-mers of one query
        if (!ctx->ioffs_valid) {
            std::vector<uint64_t> ioffs(ctx->nq + 1);
            uint64_t run = 0;
            for (uint32_t q = 0; q < ctx->nq; q++) {
                ioffs[q] = run;
                run += (ctx->h_nk[q] + KPT - 1) / KPT;
            }
            ioffs[ctx->nq] = run;
            PHY_TRY(phy_ensure(ctx, ctx->d_ioffs, ctx->nq + 2));
            PHY_TRY(phy_h2d(ctx, ctx->d_ioffs.p, ioffs.data(), ioffs.size() * sizeof(uint64_t)));
            ctx->n_hash_items = run;
            ctx->ioffs_valid = true;
        }
        const uint64_t run = ctx->n_hash_items;
        const uint64_t nb = (run + 255) / 256;
        kmer_hash31_roll_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(
            ctx->d_seq.p, ctx->d_qoffs.p, ctx->d_koffs.p, ctx->d_ioffs.p, ctx->nq, run, ctx->total_kmers,
            (int)ctx->q_canon, nh, ctx->d_hashes.p, ctx->d_counters.p + 2);
    } else {
        kmer_hash_kernel<0><<<(unsigned)nblk, 256, 0, ctx->stream>>>(
            ctx->d_seq.p, ctx->d_qoffs.p, ctx->d_koffs.p, ctx->nq, ctx->total_kmers,
            (int)ctx->q_term_size, (int)ctx->q_canon, nh, ctx->d_hashes.p, ctx->d_counters.p + 2);
    }
    ctx->launches++;
    PHY_CUDA(ctx, cudaGetLastError());
    // the error word (a query with a letter outside ACGT) is read at the next point where the host
    // waits for the stream anyway: phy_check_hash_error, or with the counters after the gather
    ctx->hash_check_pending = true;
    ctx->hashes_valid = true;
    return PHY_OK;
}

int phy_hash_error_of(phy_ctx* ctx, unsigned long long word) {
    ctx->hash_check_pending = false;
    if (word != ~0ULL) {
        ctx->hashes_valid = false;
        phy_set_error(ctx, "query #%u holds a letter outside ACGT", (unsigned)(word & 0xFFFFFFFFu));
        return PHY_ERR_QUERY;
    }
    return PHY_OK;
}

int phy_check_hash_error(phy_ctx* ctx) {
    if (!ctx->hash_check_pending) return PHY_OK;
    unsigned long long e = 0;
    PHY_CUDA(ctx, cudaMemcpyAsync(&e, ctx->d_counters.p + 2, sizeof e, cudaMemcpyDeviceToHost, ctx->stream));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return phy_hash_error_of(ctx, e);
}
