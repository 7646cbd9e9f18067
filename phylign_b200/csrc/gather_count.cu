// gather_count.cu -- K2/K3: the hot kernel of the match stage.
//
// Replaces cobs `read_from_disk` (row gather) + AND across hash functions +
// `compute_counts` (SSE2 byte-expansion adds) + `counts_to_result` (threshold), and the
// per-batch top-N + ties of /root/reference/scripts/postprocess_cobs.py:21-38
// (Appendix A.4-A.7).
//
// Shape of the work: for every (query, index) and every query k-mer, one row of the
// bit-sliced index (ceil(D/8) bytes, random address) is read from HBM and added into D
// per-document counters.  HBM-bound: one 16-B piece per lane per row, and the counters are
// VERTICAL (bit-sliced): plane p of a lane holds bit p of the counters of the 128 documents
// that lane owns, so adding a row costs 2-3 LOP3 per 32 documents (Harley-Seal carry-save
// tree over blocks of 8/16/32 rows, then a ripple into the upper planes) instead of one add
// per document.  Threshold, top-N cut and the exact threshold pruning are evaluated
// bit-sliced on the planes; only kept documents ever get their score extracted.
//
// Geometry: a row is covered by LPR lanes x 16 B (LPR = 1..32, a power of two chosen from the
// row stride); a warp holds 32/LPR independent (query,index) units.
//
// Kernels (profiles/):
//   gather_count_ring_kernel   per-warp shared-memory ring fed by lane-private cp.async.cg, persistent
//                              warps pulling (query, index) units from a global counter, 8/10/14
//                              counter planes, exact threshold pruning; MULTI = indexes with several
//                              hash functions (the h rows of a k-mer are AND-ed as they leave the ring)
//   accum_scores_ring_kernel + select_scores_kernel   the same ring for K > 16383, D > 4096, phy_scores
// Earlier generations (rows staged in registers; cp.async.bulk + mbarrier ring) were measured in
// round 1 (profiles/r01_v1_*, r01_v2_*) and removed once the ring kernel covered every index shape.
#include "phy_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;

// full adder on 32 documents at once: 2 LOP3
__device__ __forceinline__ void csa(uint32_t& hi, uint32_t& lo, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t u = a ^ b;
    hi = (a & b) | (u & c);
    lo = u ^ c;
}

template <int LPR>
__device__ __forceinline__ unsigned group_mask(int lane) {
    if constexpr (LPR == 32) return FULL;
    else return ((1u << LPR) - 1u) << (lane & ~(LPR - 1));
}
template <int LPR>
__device__ __forceinline__ uint32_t group_sum(uint32_t v, unsigned gm) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gm, v, o);
    return v;
}
template <int LPR>
__device__ __forceinline__ uint32_t group_exscan(uint32_t v, unsigned gm, int col) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) {
        uint32_t t = __shfl_up_sync(gm, incl, o, LPR);
        if (col >= o) incl += t;
    }
    return incl - v;
}

// documents (bits) whose vertical counter is >= T
template <int P>
__device__ __forceinline__ uint32_t ge_mask(const uint32_t (&pl)[P][4], int w, uint32_t T) {
    uint32_t gt = 0, eq = FULL;
#pragma unroll
    for (int p = P - 1; p >= 0; p--) {
        uint32_t c = pl[p][w];
        if ((T >> p) & 1u) {
            eq &= c;
        } else {
            gt |= eq & c;
            eq &= ~c;
        }
    }
    return gt | eq;
}

template <int P>
__device__ __forceinline__ uint32_t extract_score(const uint32_t (&pl)[P][4], int w, int b) {
    uint32_t s = 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
        uint32_t x = w == 0 ? pl[p][0] : (w == 1 ? pl[p][1] : (w == 2 ? pl[p][2] : pl[p][3]));
        s |= ((x >> b) & 1u) << p;
    }
    return s;
}

struct GatherArgs {
    const DevIndex* indexes;
    const uint32_t* class_idx;  // index ids handled by this launch (same LPR class)
    uint32_t n_class_idx;
    const uint32_t* qlist;      // query ids with 1 <= K <= PHY_FUSED_KMAX
    uint32_t n_q;
    const uint64_t* koffs;
    const uint32_t* nk;
    const uint32_t* T;
    const uint64_t* hashes;
    uint64_t total_kmers;
    uint32_t top_n;
    phy_unit* units;
    uint64_t units_cap;
    phy_hit* hits;
    uint64_t hits_cap;
    unsigned long long* counters;  // [0] hits cursor, [1] units cursor
    uint32_t* qcount;
    uint32_t* unit_flag;  // [index slot][query] 0/1 (null: host orders the units)
    uint32_t* unit_id;    // [index slot][query] position in units[]
    uint32_t nq;
    uint32_t prune;       // threshold pruning on (ring kernel)
    uint32_t class_hashes;  // MULTI launches: num_hashes shared by the indexes of the launch
    unsigned long long* idx_bytes;  // [index slot] index-row bytes really gathered (ring kernel)
};

// Threshold (cobs counts_to_result), top-N + ties (postprocess_cobs.py:21-38) and emission of
// one (query, index) unit held as vertical counters by a group of LPR lanes.  Bit-sliced:
// no per-document score is materialised unless the document is kept.
template <int LPR, int P>
__device__ __forceinline__ void select_and_emit(const uint32_t (&pl)[P][4], const GatherArgs& a, uint32_t q,
                                                uint32_t idx_id, uint32_t n_docs, uint32_t nrows, int lane,
                                                unsigned gm) {
    const int col = lane & (LPR - 1);
    const uint32_t T = a.T[q];
    const uint32_t doc0 = (uint32_t)col * 128u;
    uint32_t pass[4], cnt = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint32_t d = doc0 + w * 32u;
        uint32_t valid = d >= n_docs ? 0u : (n_docs - d >= 32u ? FULL : ((1u << (n_docs - d)) - 1u));
        pass[w] = T > nrows ? 0u : (ge_mask<P>(pl, w, T) & valid);
        cnt += __popc(pass[w]);
    }
    const uint32_t n_pass = group_sum<LPR>(cnt, gm);
    if (n_pass == 0) return;

    // cut = N-th largest score, found bit by bit from the MSB
    uint32_t n_kept = n_pass;
    if (a.top_n != 0 && n_pass > a.top_n) {
        uint32_t cut = 0;
#pragma unroll
        for (int p = P - 1; p >= 0; p--) {
            uint32_t cand = cut | (1u << p), c = 0;
#pragma unroll
            for (int w = 0; w < 4; w++) c += __popc(ge_mask<P>(pl, w, cand) & pass[w]);
            c = group_sum<LPR>(c, gm);
            if (c >= a.top_n) cut = cand;
        }
        cnt = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            pass[w] &= ge_mask<P>(pl, w, cut);
            cnt += __popc(pass[w]);
        }
        n_kept = group_sum<LPR>(cnt, gm);
    }

    // emit (doc ascending inside the unit; sorted by score in sort_units_kernel)
    const uint32_t ex = group_exscan<LPR>(cnt, gm, col);
    unsigned long long off = 0;
    if (col == 0) {
        off = atomicAdd(&a.counters[0], (unsigned long long)n_kept);
        unsigned long long u = atomicAdd(&a.counters[1], 1ULL);
        if (u < a.units_cap) {
            phy_unit pu;
            pu.query = q; pu.index = idx_id; pu.n_pass = n_pass; pu.n_kept = n_kept; pu.offset = off;
            a.units[u] = pu;
            if (a.unit_flag) {
                const uint64_t cell = (uint64_t)idx_id * a.nq + q;
                a.unit_flag[cell] = 1u;
                a.unit_id[cell] = (uint32_t)u;
            }
        }
        atomicAdd(&a.qcount[q], n_kept);
    }
    off = __shfl_sync(gm, off, lane - col);
    if (off + n_kept > a.hits_cap) return;  // host regrows and reruns
    phy_hit* out = a.hits + off + ex;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint32_t m = pass[w];
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            phy_hit h;
            h.doc = doc0 + w * 32 + b;
            h.score = extract_score<P>(pl, w, b);
            *out++ = h;
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
constexpr int BULK_BATCH_BYTES = 4096;  // one ring batch: 8 rows x (32 lanes x 16 B)

// ---- the ring: lane-private cp.async staging (all strides) ------------------------------------
// Every warp owns a ring of NB batches (8 rows each, 4 KB) in shared memory.  Every lane copies
// exactly the 16 B it will read back (cp.async.cg, SASS LDGSTS.BYPASS) NB batches ahead of the
// carry-save adds and tracks completion with commit/wait groups: no mbarrier, no cross-lane
// visibility to arrange, 3 instructions per row (the cp.async.bulk + mbarrier variant of round 1
// needed a 9-instruction issue loop per row and was ALU-bound below 256-B rows,
// profiles/r01_gather_d1000_*).  The carry-save tree is three levels deep: per 32 rows and 32
// documents 4x7 + 2 + 1 full adders and ONE ripple into planes 5.. (2.25 LOP3 per row-word).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// add 8 rows (v) into the vertical counters; batch number j selects the tree level to close
template <int P>
__device__ __forceinline__ void csa_batch3(uint32_t (&pl)[P][4], const uint4 (&v)[8], uint32_t j,
                                           uint32_t (&p8)[4], uint32_t (&p16)[4]) {
    static_assert(P >= 6, "three-level tree needs at least 6 planes");
#pragma unroll
    for (int w = 0; w < 4; w++) {
        auto W = [&](const uint4& x) -> uint32_t { return w == 0 ? x.x : (w == 1 ? x.y : (w == 2 ? x.z : x.w)); };
        uint32_t twoA, twoB, fourA, fourB, eight;
        csa(twoA, pl[0][w], pl[0][w], W(v[0]), W(v[1]));
        csa(twoB, pl[0][w], pl[0][w], W(v[2]), W(v[3]));
        csa(fourA, pl[1][w], pl[1][w], twoA, twoB);
        csa(twoA, pl[0][w], pl[0][w], W(v[4]), W(v[5]));
        csa(twoB, pl[0][w], pl[0][w], W(v[6]), W(v[7]));
        csa(fourB, pl[1][w], pl[1][w], twoA, twoB);
        csa(eight, pl[2][w], pl[2][w], fourA, fourB);
        if ((j & 1u) == 0) {
            p8[w] = eight;
        } else {
            uint32_t sixteen;
            csa(sixteen, pl[3][w], pl[3][w], p8[w], eight);
            if ((j & 2u) == 0) {
                p16[w] = sixteen;
            } else {
                uint32_t carry;
                csa(carry, pl[4][w], pl[4][w], p16[w], sixteen);
#pragma unroll
                for (int p = 5; p < P; p++) {
                    uint32_t t = pl[p][w] & carry;
                    pl[p][w] ^= carry;
                    carry = t;
                }
            }
        }
    }
}
// fold the pending eights / sixteens of an unfinished group of 4 batches into the planes
template <int P>
__device__ __forceinline__ void csa_flush3(uint32_t (&pl)[P][4], uint32_t nb, const uint32_t (&p8)[4],
                                           const uint32_t (&p16)[4]) {
#pragma unroll
    for (int w = 0; w < 4; w++) {
        if (nb & 1u) {
            uint32_t carry = p8[w];
#pragma unroll
            for (int p = 3; p < P; p++) { uint32_t t = pl[p][w] & carry; pl[p][w] ^= carry; carry = t; }
        }
        if (nb & 2u) {
            uint32_t carry = p16[w];
#pragma unroll
            for (int p = 4; p < P; p++) { uint32_t t = pl[p][w] & carry; pl[p][w] ^= carry; carry = t; }
        }
    }
}

// Two-level variant for short queries (K <= 255, 8 planes): eights are paired into sixteens and
// rippled every 2 batches, so the planes are exact counts after every even batch and the
// pruning checkpoint can run every 16 rows instead of 32 (150-bp reads die at ~80 of 120 rows).
template <int P>
__device__ __forceinline__ void csa_batch2(uint32_t (&pl)[P][4], const uint4 (&v)[8], uint32_t j, uint32_t (&p8)[4]) {
    static_assert(P >= 5, "two-level tree needs at least 5 planes");
#pragma unroll
    for (int w = 0; w < 4; w++) {
        auto W = [&](const uint4& x) -> uint32_t { return w == 0 ? x.x : (w == 1 ? x.y : (w == 2 ? x.z : x.w)); };
        uint32_t twoA, twoB, fourA, fourB, eight;
        csa(twoA, pl[0][w], pl[0][w], W(v[0]), W(v[1]));
        csa(twoB, pl[0][w], pl[0][w], W(v[2]), W(v[3]));
        csa(fourA, pl[1][w], pl[1][w], twoA, twoB);
        csa(twoA, pl[0][w], pl[0][w], W(v[4]), W(v[5]));
        csa(twoB, pl[0][w], pl[0][w], W(v[6]), W(v[7]));
        csa(fourB, pl[1][w], pl[1][w], twoA, twoB);
        csa(eight, pl[2][w], pl[2][w], fourA, fourB);
        if ((j & 1u) == 0) {
            p8[w] = eight;
        } else {
            uint32_t carry;
            csa(carry, pl[3][w], pl[3][w], p8[w], eight);
#pragma unroll
            for (int p = 4; p < P; p++) {
                uint32_t t = pl[p][w] & carry;
                pl[p][w] ^= carry;
                carry = t;
            }
        }
    }
}
template <int P>
__device__ __forceinline__ void csa_flush2(uint32_t (&pl)[P][4], uint32_t nb, const uint32_t (&p8)[4]) {
    if (nb & 1u) {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t carry = p8[w];
#pragma unroll
            for (int p = 3; p < P; p++) { uint32_t t = pl[p][w] & carry; pl[p][w] ^= carry; carry = t; }
        }
    }
}

#ifndef PHY_RING_NB
#define PHY_RING_NB 3
#endif
#ifndef PHY_RING_WARPS
#define PHY_RING_WARPS 4
#endif
#ifndef PHY_RING_NB_SHORT
#define PHY_RING_NB_SHORT 2
#endif
constexpr int RING_NB = PHY_RING_NB, RING_WARPS = PHY_RING_WARPS;

// One unit per group of LPR lanes: add the rows of its `nrows` k-mers (hashes at hq) into pl.
// `nmax` = largest nrows among the groups of the warp (all lanes run the same trip count).
// THRESHOLD PRUNING (exact): after x of the K rows a document with count c can still reach
// the threshold T only if c + (K - x) >= T.  At every 4th batch (where the carry-save tree
// has no pending partial sums, so the planes ARE the counts) a lane whose 128 documents are
// all out of reach stops fetching its 16 B, and a unit whose documents are all out of reach
// ends.  Documents that can still pass keep being counted to the end, so n_pass, the kept
// set and every reported score are unchanged; only rows that cannot change the output are
// skipped.  With cobs_kmer_thres 0.7 and a 0.3 false-positive rate an unrelated batch dies
// after ~half of a read's k-mers.  prune_T = 0 switches it off (phy_scores, general path).
template <int LPR, int P, int NB>
__device__ __forceinline__ uint32_t ring_accumulate(uint32_t (&pl)[P][4], uint32_t ring, const uint8_t* colbase,
                                                bool lane_on, uint32_t stride, uint64_t sig, uint64_t magic,
                                                const uint64_t* __restrict__ hq, uint32_t nrows, uint32_t nmax,
                                                int lane, uint32_t prune_T, unsigned gm) {
    constexpr bool FINE = P <= 8;           // short queries: two-level tree, checkpoint every 2 batches
    constexpr int HB = LPR >= 8 ? LPR : 8;  // rows whose hashes are fetched per block
    constexpr int NH = HB / LPR;            // hashes per lane per block
    constexpr int BPH = HB / 8;             // batches per hash block
    const int col = lane & (LPR - 1), gbase = lane - col;
    uint32_t nb = (nmax + 7) >> 3;          // warp-uniform number of batches (shrinks when units die)
    const uint32_t K = nrows;               // this unit's k-mer count (nrows itself shrinks on early exit)
    uint32_t p8[4] = {0, 0, 0, 0}, p16[4] = {0, 0, 0, 0};
    uint32_t myrow[NH];
    uint64_t hnext[NH];
#pragma unroll
    for (int t = 0; t < NH; t++) {
        myrow[t] = PHY_ROW_INVALID;
        const uint32_t hidx = col + t * LPR;
        hnext[t] = hidx < nrows ? __ldg(hq + hidx) : 0;
    }
    auto issue = [&](uint32_t j) {  // queue the 8 rows of batch j (all groups of the warp)
        if (j < nb) {
            const uint32_t hb = j / BPH, s = j % BPH;
            if (s == 0) {
#pragma unroll
                for (int t = 0; t < NH; t++) {
                    const uint32_t hidx = hb * HB + col + t * LPR;
                    myrow[t] = hidx < nrows ? phy_fastmod(hnext[t], sig, magic) : PHY_ROW_INVALID;
                    const uint32_t hn = hidx + HB;
                    hnext[t] = hn < nrows ? __ldg(hq + hn) : 0;
                }
            }
            const uint32_t dst = ring + (j % NB) * BULK_BATCH_BYTES;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int src = LPR >= 8 ? gbase + (int)s * 8 + r : gbase + (r % LPR);
                const int slot = LPR >= 8 ? 0 : r / LPR;
                const uint32_t rr = __shfl_sync(FULL, myrow[slot], src);
                if (rr != PHY_ROW_INVALID && j * 8 + r < nrows && lane_on)
                    cp_async16(dst + r * 512, colbase + (uint64_t)rr * stride);
            }
        }
        cp_async_commit();  // (possibly empty) group: keeps the wait_group distance constant
    };
#pragma unroll 1
    for (uint32_t j = 0; j < (uint32_t)NB; j++) issue(j);
    // warp-uniform switch: the checkpoint below holds warp-wide votes, so every lane must take it
    const bool prune_any = __any_sync(FULL, prune_T != 0);
    uint32_t j = 0;
#pragma unroll 1
    for (; j < nb; j++) {
        cp_async_wait<NB - 1>();  // batch j has landed (this lane's own 16-B pieces)
        const uint32_t src = ring + (j % NB) * BULK_BATCH_BYTES;
        uint4 v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (j * 8 + r < nrows && lane_on) v[r] = lds128(src + r * 512);
            else v[r] = make_uint4(0, 0, 0, 0);
        }
        if constexpr (FINE) csa_batch2<P>(pl, v, j, p8);
        else csa_batch3<P>(pl, v, j, p8, p16);
        if (prune_any && ((j + 1) & (FINE ? 1u : 3u)) == 0) {  // planes are exact counts here
            const uint32_t x = min((j + 1) * 8u, nrows);  // rows of this unit counted so far
            bool alive = lane_on;
            if (lane_on && prune_T != 0 && prune_T + x > K) {  // need = T - (K - x) more than zero
                const uint32_t need = prune_T + x - K;
                uint32_t any = 0;
#pragma unroll
                for (int w = 0; w < 4; w++) any |= ge_mask<P>(pl, w, need);
                alive = any != 0;
                lane_on = alive;                      // my 128 documents cannot pass any more
            }
            const unsigned live = __ballot_sync(FULL, alive && x < nrows);
            if ((live & gm) == 0 && x < nrows) nrows = x;  // the whole unit is decided: stop here
            if (live != FULL) {                       // some unit ended: the warp may finish earlier
                uint32_t m = (nrows + 7) >> 3;
#pragma unroll
                for (int o = 16; o >= LPR; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
                nb = max(m, j + 1);
            }
        }
        issue(j + NB);  // refill the slot: its values were consumed above by this very lane
    }
    cp_async_wait<0>();
    if constexpr (FINE) csa_flush2<P>(pl, j, p8);
    else csa_flush3<P>(pl, j, p8, p16);
    return min(j * 8u, K);  // rows of this unit that were fetched and counted
}

// The same for indexes built with several hash functions (num_hashes = h > 1): a k-mer matches a
// document iff the bit is set in ALL h rows, so the rows of hash function 0..h-1 are
// AND-ed before they are counted.  The ring streams "virtual batches" jj = j*h + hj (8 k-mers of
// batch j under hash function hj): same slots, same commit/wait distance; the 8 row pieces of a
// lane stay in registers across the h virtual batches of one k-mer batch and enter the carry-save
// tree once, after the last AND.  Hashes are fetched 8 per virtual batch (hq[hj*hstride + i]).
// Pruning works on the same exact counts as with one hash function.
template <int LPR, int P, int NB>
__device__ __forceinline__ uint32_t ring_accumulate_multi(uint32_t (&pl)[P][4], uint32_t ring, const uint8_t* colbase,
                                                      bool lane_on, uint32_t stride, uint64_t sig, uint64_t magic,
                                                      const uint64_t* __restrict__ hq, uint64_t hstride, uint32_t h,
                                                      uint32_t nrows, uint32_t nmax, int lane, uint32_t prune_T,
                                                      unsigned gm) {
    constexpr bool FINE = P <= 8;
    constexpr int NH = LPR >= 8 ? 1 : 8 / LPR;  // hashes per lane per virtual batch (8 rows)
    const int col = lane & (LPR - 1), gbase = lane - col;
    uint32_t nb = (nmax + 7) >> 3;              // k-mer batches (warp uniform; shrinks when units die)
    const uint32_t K = nrows;
    uint32_t p8[4] = {0, 0, 0, 0}, p16[4] = {0, 0, 0, 0};
    uint32_t myrow[NH];
    uint64_t hnext[NH];
    auto load_hashes = [&](uint32_t jj) {       // hashes of virtual batch jj into hnext
        const uint32_t j = jj / h, hj = jj - j * h;
#pragma unroll
        for (int t = 0; t < NH; t++) {
            const uint32_t hidx = j * 8 + (uint32_t)col + (uint32_t)t * LPR;
            const bool mine = (LPR < 8 || col < 8) && hidx < nrows && j < nb;
            hnext[t] = mine ? __ldg(hq + (uint64_t)hj * hstride + hidx) : 0;
        }
    };
    load_hashes(0);
    auto issue = [&](uint32_t jj) {
        const uint32_t j = jj / h;
        if (j < nb) {
#pragma unroll
            for (int t = 0; t < NH; t++) {
                const uint32_t hidx = j * 8 + (uint32_t)col + (uint32_t)t * LPR;
                const bool mine = (LPR < 8 || col < 8) && hidx < nrows;
                myrow[t] = mine ? phy_fastmod(hnext[t], sig, magic) : PHY_ROW_INVALID;
            }
            load_hashes(jj + 1);
            const uint32_t dst = ring + (jj % NB) * BULK_BATCH_BYTES;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int src = LPR >= 8 ? gbase + r : gbase + (r % LPR);
                const int slot = LPR >= 8 ? 0 : r / LPR;
                const uint32_t rr = __shfl_sync(FULL, myrow[slot], src);
                if (rr != PHY_ROW_INVALID && j * 8 + r < nrows && lane_on)
                    cp_async16(dst + r * 512, colbase + (uint64_t)rr * stride);
            }
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (uint32_t jj = 0; jj < (uint32_t)NB; jj++) issue(jj);
    const bool prune_any = __any_sync(FULL, prune_T != 0);
    uint4 v[8];
    uint32_t j = 0, hj = 0, jj = 0;
#pragma unroll 1
    for (; j < nb; jj++) {
        cp_async_wait<NB - 1>();
        const uint32_t src = ring + (jj % NB) * BULK_BATCH_BYTES;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint4 x = make_uint4(0, 0, 0, 0);
            if (j * 8 + r < nrows && lane_on) x = lds128(src + r * 512);
            if (hj == 0) v[r] = x;
            else { v[r].x &= x.x; v[r].y &= x.y; v[r].z &= x.z; v[r].w &= x.w; }
        }
        if (++hj == h) {                        // all hash functions of k-mer batch j are in: count it
            hj = 0;
            if constexpr (FINE) csa_batch2<P>(pl, v, j, p8);
            else csa_batch3<P>(pl, v, j, p8, p16);
            if (prune_any && ((j + 1) & (FINE ? 1u : 3u)) == 0) {
                const uint32_t x = min((j + 1) * 8u, nrows);
                bool alive = lane_on;
                if (lane_on && prune_T != 0 && prune_T + x > K) {
                    const uint32_t need = prune_T + x - K;
                    uint32_t any = 0;
#pragma unroll
                    for (int w = 0; w < 4; w++) any |= ge_mask<P>(pl, w, need);
                    alive = any != 0;
                    lane_on = alive;
                }
                const unsigned live = __ballot_sync(FULL, alive && x < nrows);
                if ((live & gm) == 0 && x < nrows) nrows = x;
                if (live != FULL) {
                    uint32_t m = (nrows + 7) >> 3;
#pragma unroll
                    for (int o = 16; o >= LPR; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
                    nb = max(m, j + 1);
                }
            }
            j++;
        }
        issue(jj + NB);
    }
    cp_async_wait<0>();
    if constexpr (FINE) csa_flush2<P>(pl, j, p8);
    else csa_flush3<P>(pl, j, p8, p16);
    return min(j * 8u, K);
}

#ifndef PHY_RING_MINBLOCKS
#define PHY_RING_MINBLOCKS 1
#endif
#ifndef PHY_RING_MINBLOCKS_NARROW
#define PHY_RING_MINBLOCKS_NARROW 4   // 128..256-B rows: 4 CTAs/SM (<= 128 registers) instead of 3: +8 % (measured; no gain below)
#endif
#ifndef PHY_RING_MINBLOCKS_SHORT
#define PHY_RING_MINBLOCKS_SHORT 1
#endif
template <int LPR, int P, int NB, int WARPS, bool MULTI>
__global__ void __launch_bounds__(WARPS * 32, (P <= 8 ? PHY_RING_MINBLOCKS_SHORT : (P <= 10 && LPR >= 8 && LPR < 32 && !MULTI) ? PHY_RING_MINBLOCKS_NARROW : PHY_RING_MINBLOCKS))
gather_count_ring_kernel(const GatherArgs a) {
    constexpr int G = 32 / LPR;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int col = lane & (LPR - 1), g = lane / LPR;
    const unsigned gm = group_mask<LPR>(lane);
    // lane-private slots: batch b, row r -> ring + b*4096 + r*512 + lane*16 (conflict-free LDS.128)
    const uint32_t ring = smem_u32(smem) + wid * (NB * BULK_BATCH_BYTES) + lane * 16;
    const uint64_t total = (uint64_t)a.n_class_idx * a.n_q;
    const uint64_t n_wunits = (total + G - 1) / G;

    for (;;) {
        unsigned long long wu = 0;
        if (lane == 0) wu = atomicAdd(&a.counters[3], 1ULL);
        wu = __shfl_sync(FULL, wu, 0);
        if (wu >= n_wunits) break;
        const uint64_t gid = wu * G + g;
        const bool live = gid < total;
        uint32_t q = 0, nrows = 0, stride = 0, n_docs = 0, idx_id = 0;
        uint64_t sig = 1, magic = 0;
        const uint8_t* colbase = nullptr;
        const uint64_t* hq = nullptr;
        if (live) {
            const uint32_t ipos = (uint32_t)(gid / a.n_q);
            q = a.qlist[gid - (uint64_t)ipos * a.n_q];
            const DevIndex& ix = a.indexes[a.class_idx[ipos]];
            stride = ix.stride; n_docs = ix.n_docs; idx_id = ix.idx_id;
            sig = ix.sig; magic = ix.magic;
            colbase = ix.rows + col * 16;
            nrows = a.nk[q];
            hq = a.hashes + a.koffs[q];
        }
        uint32_t nmax = nrows;
#pragma unroll
        for (int o = 16; o >= LPR; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));

        uint32_t pl[P][4];
#pragma unroll
        for (int p = 0; p < P; p++) pl[p][0] = pl[p][1] = pl[p][2] = pl[p][3] = 0;
        const uint32_t prune_T = (a.prune && live) ? a.T[q] : 0u;
        uint32_t done;
        if constexpr (MULTI) {
            const uint32_t h = a.class_hashes;  // a launch holds indexes with the same number of hash functions
            done = ring_accumulate_multi<LPR, P, NB>(pl, ring, colbase, (uint32_t)col * 16u < stride, stride, sig, magic,
                                                     hq, a.total_kmers, h, nrows, nmax, lane, prune_T, gm) * h;
        } else {
            done = ring_accumulate<LPR, P, NB>(pl, ring, colbase, (uint32_t)col * 16u < stride, stride, sig, magic, hq,
                                               nrows, nmax, lane, prune_T, gm);
        }
        if (live && col == 0)  // bytes of index rows this unit really gathered (pruning makes it < K rows)
            atomicAdd(&a.idx_bytes[idx_id], (unsigned long long)done * ((n_docs + 7u) >> 3));
        if (live) select_and_emit<LPR, P>(pl, a, q, idx_id, n_docs, nrows, lane, gm);
        __syncwarp();
    }
}

// General path on the ring (one hash function): k-mers [k0,k1) of a query against ONE index and
// one 512-B column chunk, flushed into a dense uint32 score row with atomics.  Chunks hold up
// to 2^14-1 k-mers (14 planes).  For K > PHY_LONG_KMAX, rows wider than 512 B (D > 4096), phy_scores().
constexpr int PHY_GENERAL_PLANES = 14;
template <int LPR, int NB, int WARPS, bool MULTI>
__global__ void __launch_bounds__(WARPS * 32) accum_scores_ring_kernel(
    const DevIndex* __restrict__ ixp, const SlowItem* __restrict__ items, uint32_t n_items, uint32_t n_chunks,
    const uint64_t* __restrict__ koffs, const uint64_t* __restrict__ hashes, uint64_t hstride,
    uint32_t* __restrict__ scores, unsigned long long* __restrict__ counter) {
    constexpr int P = PHY_GENERAL_PLANES;
    constexpr int G = 32 / LPR;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int col = lane & (LPR - 1), g = lane / LPR;
    const uint32_t ring = smem_u32(smem) + wid * (NB * BULK_BATCH_BYTES) + lane * 16;
    const uint64_t total = (uint64_t)n_items * n_chunks;
    const uint64_t n_wunits = (total + G - 1) / G;
    const DevIndex& ix = *ixp;
    const uint32_t stride = ix.stride, n_docs = ix.n_docs;
    for (;;) {
        unsigned long long wu = 0;
        if (lane == 0) wu = atomicAdd(counter, 1ULL);
        wu = __shfl_sync(FULL, wu, 0);
        if (wu >= n_wunits) break;
        const uint64_t gid = wu * G + g;
        const bool live = gid < total;
        SlowItem it = {0, 0, 0, 0};
        uint32_t chunk = 0;
        if (live) {
            it = items[gid / n_chunks];
            chunk = (uint32_t)(gid % n_chunks);
        }
        const uint32_t nrows = it.k1 - it.k0;
        uint32_t nmax = nrows;
#pragma unroll
        for (int o = 16; o >= LPR; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));
        const uint32_t byte0 = chunk * PHY_CHUNK_BYTES + (uint32_t)col * 16u;
        uint32_t pl[P][4];
#pragma unroll
        for (int p = 0; p < P; p++) pl[p][0] = pl[p][1] = pl[p][2] = pl[p][3] = 0;
        const uint64_t* hq = hashes + (live ? koffs[it.query] + it.k0 : 0);
        if constexpr (MULTI)
            ring_accumulate_multi<LPR, P, NB>(pl, ring, ix.rows + byte0, live && byte0 < stride, stride, ix.sig, ix.magic,
                                              hq, hstride, ix.num_hashes, nrows, nmax, lane, 0u, group_mask<LPR>(lane));
        else
            ring_accumulate<LPR, P, NB>(pl, ring, ix.rows + byte0, live && byte0 < stride, stride, ix.sig, ix.magic, hq,
                                        nrows, nmax, lane, 0u, group_mask<LPR>(lane));
        if (live) {
            uint32_t* row = scores + (uint64_t)it.slot * n_docs;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t any = 0;
#pragma unroll
                for (int p = 0; p < P; p++) any |= pl[p][w];
                const uint32_t d0 = byte0 * 8u + w * 32u;
                while (any) {
                    int b = __ffs(any) - 1;
                    any &= any - 1;
                    if (d0 + b < n_docs) atomicAdd(&row[d0 + b], extract_score<P>(pl, w, b));
                }
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* sm) {
    // all 256 threads; returns the total to every thread
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t += sm[i];
    return t;
}

// threshold + top-N + ties on a dense score row (general path).  One block per slot.
__global__ void __launch_bounds__(256) select_scores_kernel(
    const uint32_t* __restrict__ scores, uint32_t n_docs, const uint32_t* __restrict__ slotq,
    const uint32_t* __restrict__ Tq, const uint32_t* __restrict__ nk, uint32_t top_n,
    uint32_t idx_id, phy_unit* units, uint64_t units_cap, phy_hit* hits, uint64_t hits_cap,
    unsigned long long* counters, uint32_t* qcount, uint32_t* unit_flag, uint32_t* unit_id, uint32_t nq) {
    __shared__ uint32_t sm[8];
    __shared__ uint32_t sm_scan[256];
    __shared__ unsigned long long sm_off;
    const uint32_t slot = blockIdx.x, q = slotq[slot];
    const uint32_t* sc = scores + (uint64_t)slot * n_docs;
    const uint32_t T = Tq[q], K = nk[q];
    const uint32_t per = (n_docs + 255) / 256;
    const uint32_t d_lo = min(threadIdx.x * per, n_docs), d_hi = min(d_lo + per, n_docs);
    auto count_ge = [&](uint32_t c) {
        uint32_t n = 0;
        for (uint32_t d = threadIdx.x; d < n_docs; d += 256) n += sc[d] >= c;
        return block_sum(n, sm);
    };
    const uint32_t n_pass = T > K ? 0 : count_ge(T);
    if (n_pass == 0) return;
    uint32_t cut = T, n_kept = n_pass;
    if (top_n != 0 && n_pass > top_n) {
        uint32_t lo = T, hi = K + 1;  // count_ge(lo) >= top_n, count_ge(hi) = 0
        while (hi - lo > 1) {
            uint32_t mid = lo + (hi - lo) / 2;
            if (count_ge(mid) >= top_n) lo = mid; else hi = mid;
        }
        cut = lo;
        n_kept = count_ge(cut);
    }
    uint32_t mine = 0;
    for (uint32_t d = d_lo; d < d_hi; d++) mine += sc[d] >= cut;
    sm_scan[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 256; i++) { uint32_t t = sm_scan[i]; sm_scan[i] = run; run += t; }
        unsigned long long off = atomicAdd(&counters[0], (unsigned long long)n_kept);
        unsigned long long u = atomicAdd(&counters[1], 1ULL);
        if (u < units_cap) {
            phy_unit pu;
            pu.query = q; pu.index = idx_id; pu.n_pass = n_pass; pu.n_kept = n_kept; pu.offset = off;
            units[u] = pu;
            if (unit_flag) {
                const uint64_t cell = (uint64_t)idx_id * nq + q;
                unit_flag[cell] = 1u;
                unit_id[cell] = (uint32_t)u;
            }
        }
        atomicAdd(&qcount[q], n_kept);
        sm_off = off;
    }
    __syncthreads();
    if (sm_off + n_kept > hits_cap) return;
    phy_hit* out = hits + sm_off + sm_scan[threadIdx.x];
    for (uint32_t d = d_lo; d < d_hi; d++) {
        uint32_t s = sc[d];
        if (s >= cut) { phy_hit h; h.doc = d; h.score = s; *out++ = h; }
    }
}

int lpr_for_stride(uint32_t stride) {
    uint32_t segs = (stride + 15) / 16;
    if (segs > 32) segs = 32;
    int lpr = 1;
    while ((uint32_t)lpr < segs) lpr <<= 1;
    return lpr;
}

template <int LPR, int P, int NB, int WARPS, bool MULTI>
int launch_ring(const GatherArgs& a, cudaStream_t st, int n_sm) {
    constexpr int SMEM = WARPS * NB * BULK_BATCH_BYTES;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gather_count_ring_kernel<LPR, P, NB, WARPS, MULTI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -1;
        configured = true;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_count_ring_kernel<LPR, P, NB, WARPS, MULTI>,
                                                      WARPS * 32, SMEM) != cudaSuccess || per_sm < 1) return -1;
    constexpr int G = 32 / LPR;
    static const int cap_per_sm = getenv("PHY_RING_BLOCKS_PER_SM") ? atoi(getenv("PHY_RING_BLOCKS_PER_SM")) : 0;
    if (cap_per_sm > 0) per_sm = std::min(per_sm, cap_per_sm);   // experiments: leave SM room for a co-running kernel
    uint64_t wunits = ((uint64_t)a.n_class_idx * a.n_q + G - 1) / G;
    uint64_t blocks = std::min<uint64_t>((wunits + WARPS - 1) / WARPS, (uint64_t)n_sm * per_sm);
    if (blocks) gather_count_ring_kernel<LPR, P, NB, WARPS, MULTI><<<(unsigned)blocks, WARPS * 32, SMEM, st>>>(a);
    return 0;
}
template <int P, bool MULTI>
int dispatch_ring(int c, const GatherArgs& a, cudaStream_t st, int n_sm) {
    // short-read class (8 planes): units of <= 255 rows die early under pruning, a 2-deep ring wastes
    // one batch less when they do (measured 7 % faster on 150-bp reads than the 3-deep ring)
    constexpr int NB = P <= 8 ? PHY_RING_NB_SHORT : RING_NB;
    switch (c) {
        case 0: return launch_ring<1, P, NB, RING_WARPS, MULTI>(a, st, n_sm);
        case 1: return launch_ring<2, P, NB, RING_WARPS, MULTI>(a, st, n_sm);
        case 2: return launch_ring<4, P, NB, RING_WARPS, MULTI>(a, st, n_sm);
        case 3: return launch_ring<8, P, NB, RING_WARPS, MULTI>(a, st, n_sm);
        case 4: return launch_ring<16, P, NB, RING_WARPS, MULTI>(a, st, n_sm);
        default: return launch_ring<32, P, NB, RING_WARPS, MULTI>(a, st, n_sm);
    }
}
template <bool MULTI>
int dispatch_ring_planes(int pass, int c, const GatherArgs& a, cudaStream_t st, int n_sm) {
    return pass == 0 ? dispatch_ring<10, MULTI>(c, a, st, n_sm)
         : pass == 1 ? dispatch_ring<8, MULTI>(c, a, st, n_sm)
                     : dispatch_ring<14, MULTI>(c, a, st, n_sm);
}

template <int LPR, bool MULTI>
int launch_accum_ring(const DevIndex* ixp, const SlowItem* items, uint32_t n_items, uint32_t n_chunks,
                      const uint64_t* koffs, const uint64_t* hashes, uint64_t hstride, uint32_t* scores,
                      unsigned long long* counter, cudaStream_t st, int n_sm) {
    constexpr int NB = RING_NB, WARPS = RING_WARPS;
    constexpr int SMEM = WARPS * NB * BULK_BATCH_BYTES;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(accum_scores_ring_kernel<LPR, NB, WARPS, MULTI>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -1;
        configured = true;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, accum_scores_ring_kernel<LPR, NB, WARPS, MULTI>,
                                                      WARPS * 32, SMEM) != cudaSuccess || per_sm < 1) return -1;
    constexpr int G = 32 / LPR;
    uint64_t wunits = ((uint64_t)n_items * n_chunks + G - 1) / G;
    uint64_t blocks = std::min<uint64_t>((wunits + WARPS - 1) / WARPS, (uint64_t)n_sm * per_sm);
    if (blocks)
        accum_scores_ring_kernel<LPR, NB, WARPS, MULTI><<<(unsigned)blocks, WARPS * 32, SMEM, st>>>(
            ixp, items, n_items, n_chunks, koffs, hashes, hstride, scores, counter);
    return 0;
}
template <bool MULTI>
int dispatch_accum_ring(int lpr, const DevIndex* ixp, const SlowItem* items, uint32_t n_items, uint32_t n_chunks,
                        const uint64_t* koffs, const uint64_t* hashes, uint64_t hstride, uint32_t* scores,
                        unsigned long long* counter, cudaStream_t st, int n_sm) {
    switch (lpr) {
        case 1: return launch_accum_ring<1, MULTI>(ixp, items, n_items, n_chunks, koffs, hashes, hstride, scores, counter, st, n_sm);
        case 2: return launch_accum_ring<2, MULTI>(ixp, items, n_items, n_chunks, koffs, hashes, hstride, scores, counter, st, n_sm);
        case 4: return launch_accum_ring<4, MULTI>(ixp, items, n_items, n_chunks, koffs, hashes, hstride, scores, counter, st, n_sm);
        case 8: return launch_accum_ring<8, MULTI>(ixp, items, n_items, n_chunks, koffs, hashes, hstride, scores, counter, st, n_sm);
        case 16: return launch_accum_ring<16, MULTI>(ixp, items, n_items, n_chunks, koffs, hashes, hstride, scores, counter, st, n_sm);
        default: return launch_accum_ring<32, MULTI>(ixp, items, n_items, n_chunks, koffs, hashes, hstride, scores, counter, st, n_sm);
    }
}

// chunk a query's k-mers into items of at most `cmax` rows (what the counter planes can hold)
void push_items(std::vector<SlowItem>& items, uint32_t slot, uint32_t q, uint32_t K, uint32_t cmax) {
    for (uint32_t k0 = 0; k0 < K; k0 += cmax) {
        SlowItem it;
        it.slot = slot; it.query = q; it.k0 = k0;
        it.k1 = K - k0 > cmax ? k0 + cmax : K;
        items.push_back(it);
    }
}

}  // namespace

int phy_lpr_for_stride(uint32_t stride) { return lpr_for_stride(stride); }

// Run the general path for `slots` (query ids) against index ix; scores into d_scores.
static int run_general(phy_ctx* ctx, const HostIndex& ix, int ipos, const std::vector<uint32_t>& slot_queries,
                       uint32_t* d_scores_out) {
    std::vector<SlowItem> items;
    for (uint32_t s = 0; s < slot_queries.size(); s++)
        push_items(items, s, slot_queries[s], ctx->h_nk[slot_queries[s]], PHY_LONG_KMAX);
    const size_t n_sc = slot_queries.size() * (size_t)ix.d.n_docs;
    PHY_CUDA(ctx, cudaMemsetAsync(d_scores_out, 0, n_sc * sizeof(uint32_t), ctx->stream));
    if (items.empty()) return PHY_OK;
    PHY_TRY(phy_ensure(ctx, ctx->d_items, items.size()));
    PHY_TRY(phy_h2d(ctx, ctx->d_items.p, items.data(), items.size() * sizeof(SlowItem)));
    const uint32_t n_chunks = (ix.d.stride + PHY_CHUNK_BYTES - 1) / PHY_CHUNK_BYTES;
    PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p + 4, 0, sizeof(unsigned long long), ctx->stream));
    const int rc = ix.d.num_hashes > 1
        ? dispatch_accum_ring<true>(ix.lpr, ctx->d_indexes.p + ipos, ctx->d_items.p, (uint32_t)items.size(), n_chunks,
                                    ctx->d_koffs.p, ctx->d_hashes.p, ctx->total_kmers, d_scores_out,
                                    ctx->d_counters.p + 4, ctx->stream, ctx->n_sm)
        : dispatch_accum_ring<false>(ix.lpr, ctx->d_indexes.p + ipos, ctx->d_items.p, (uint32_t)items.size(), n_chunks,
                                     ctx->d_koffs.p, ctx->d_hashes.p, ctx->total_kmers, d_scores_out,
                                     ctx->d_counters.p + 4, ctx->stream, ctx->n_sm);
    if (rc != 0) {
        phy_set_error(ctx, "cannot configure the ring score kernel: %s", cudaGetErrorString(cudaGetLastError()));
        return PHY_ERR_CUDA;
    }
    ctx->launches++;
    PHY_CUDA(ctx, cudaGetLastError());
    return PHY_OK;
}

int phy_launch_scores(phy_ctx* ctx, int idx_id, uint32_t* d_out_scores) {
    PHY_TRY(phy_check_hash_error(ctx));
    const HostIndex& ix = ctx->idx[idx_id];
    std::vector<uint32_t> qs;
    for (uint32_t q = 0; q < ctx->nq; q++) qs.push_back(q);
    PHY_TRY(phy_ensure(ctx, ctx->d_counters, 8));
    return run_general(ctx, ix, idx_id, qs, d_out_scores);
}

int phy_launch_gather(phy_ctx* ctx, const phy_match_params* p) {
    // per-query minimum score T (host double arithmetic, identical to the oracle / cobs) and the
    // query classes by k-mer count: <= 255 (8 counter planes), <= 1023 (10), <= 16383 (14) are fused
    // in the ring kernel; longer ones go through the chunked dense-score path.
    // These tables depend on the query set, the threshold and the set of active indexes only: they are
    // built and uploaded when one of those changes, not once per pass.
    GatherTables& gt = ctx->gt;
    if (!ctx->gather_tables_valid || gt.threshold != p->threshold || gt.floor_mode != p->floor_mode ||
        gt.index_version != ctx->index_version) {
        std::vector<uint32_t> T(ctx->nq), shortq, fastq, midq;
        gt.slowq.clear();
        for (uint32_t q = 0; q < ctx->nq; q++) {
            double x = p->threshold * (double)ctx->h_nk[q];
            double r = p->floor_mode ? floor(x) : ceil(x);
            T[q] = r < 0 ? 0u : (r > 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)r);
            const uint32_t K = ctx->h_nk[q];
            if (K == 0) continue;
            if (K <= PHY_SHORT_KMAX) shortq.push_back(q);
            else if (K <= PHY_FUSED_KMAX) fastq.push_back(q);
            else if (K <= PHY_LONG_KMAX) midq.push_back(q);
            else gt.slowq.push_back(q);
        }
        // longest first: balances the tail and keeps co-resident groups of a warp similar
        auto sort_by_len_desc = [&](std::vector<uint32_t>& v) {  // stable counting sort: K <= 16383
            if (v.size() < 2) return;
            uint32_t kmin = 0xFFFFFFFFu, kmax = 0;
            for (uint32_t q : v) { kmin = std::min(kmin, ctx->h_nk[q]); kmax = std::max(kmax, ctx->h_nk[q]); }
            if (kmin == kmax) return;  // all reads equally long (the common case): nothing to do
            std::vector<uint32_t> start(kmax - kmin + 2, 0), out(v.size());
            for (uint32_t q : v) start[kmax - ctx->h_nk[q] + 1]++;
            for (size_t i = 1; i < start.size(); i++) start[i] += start[i - 1];
            for (uint32_t q : v) out[start[kmax - ctx->h_nk[q]]++] = q;
            v.swap(out);
        };
        sort_by_len_desc(shortq);
        sort_by_len_desc(fastq);
        sort_by_len_desc(midq);
        // device list layout: [10-plane | 8-plane | 14-plane]; the first two together are "K <= 1023"
        gt.n_fast8 = (uint32_t)shortq.size();
        gt.n_fast10 = (uint32_t)fastq.size();
        gt.n_fast14 = (uint32_t)midq.size();
        gt.fastq = fastq;
        gt.fastq.insert(gt.fastq.end(), shortq.begin(), shortq.end());
        gt.fastq.insert(gt.fastq.end(), midq.begin(), midq.end());   // every fused-capable query
        PHY_TRY(phy_ensure(ctx, ctx->d_T, ctx->nq + 1));
        PHY_TRY(phy_h2d(ctx, ctx->d_T.p, T.data(), T.size() * sizeof(uint32_t)));
        PHY_TRY(phy_ensure(ctx, ctx->d_qlist, ctx->nq + 1));
        if (!gt.fastq.empty()) PHY_TRY(phy_h2d(ctx, ctx->d_qlist.p, gt.fastq.data(), gt.fastq.size() * sizeof(uint32_t)));

        // index classes: one launch per (lanes-per-row class, number of hash functions, query class);
        // rows wider than 512 B take the general path for every query
        gt.classes.clear();
        for (size_t i = 0; i < ctx->idx.size(); i++) {
            const HostIndex& ix = ctx->idx[i];
            if (!ix.alive || !ix.committed || !ix.active) continue;
            if (ix.d.stride > PHY_CHUNK_BYTES) continue;
            int c = 0;
            while ((1 << c) < ix.lpr) c++;
            GatherTables::IdxClass* k = nullptr;
            for (auto& x : gt.classes)
                if (x.c == c && x.h == ix.d.num_hashes) k = &x;
            if (!k) {
                gt.classes.push_back(GatherTables::IdxClass{c, ix.d.num_hashes, {}, 0});
                k = &gt.classes.back();
            }
            k->ids.push_back((uint32_t)i);
        }
        std::sort(gt.classes.begin(), gt.classes.end(), [](const GatherTables::IdxClass& x, const GatherTables::IdxClass& y) {
            return x.h != y.h ? x.h < y.h : x.c < y.c;
        });
        std::vector<uint32_t> class_flat;
        for (auto& k : gt.classes) { k.off = class_flat.size(); class_flat.insert(class_flat.end(), k.ids.begin(), k.ids.end()); }
        if (!class_flat.empty()) {
            PHY_TRY(phy_ensure(ctx, ctx->d_class, class_flat.size()));
            PHY_TRY(phy_h2d(ctx, ctx->d_class.p, class_flat.data(), class_flat.size() * sizeof(uint32_t)));
        }
        gt.threshold = p->threshold;
        gt.floor_mode = p->floor_mode;
        gt.index_version = ctx->index_version;
        ctx->gather_tables_valid = true;
    }
    const std::vector<uint32_t>& fastq = gt.fastq;
    const std::vector<uint32_t>& slowq = gt.slowq;
    const std::vector<GatherTables::IdxClass>& classes = gt.classes;
    using IdxClass = GatherTables::IdxClass;
    const uint32_t n_fast8 = gt.n_fast8, n_fast10 = gt.n_fast10, n_fast14 = gt.n_fast14;
    const uint32_t n_le1023 = n_fast10 + n_fast8;
    uint32_t* d_class = ctx->d_class.p;
    PHY_TRY(phy_ensure(ctx, ctx->d_qcount, ctx->nq + 1));
    PHY_TRY(phy_ensure(ctx, ctx->d_counters, 8));

    uint64_t units_cap = ctx->d_units.cap, hits_cap = ctx->d_hits.cap;
    // first sizing guess for a new workload (an overflow is still caught: the pass reports the needed
    // sizes and is rerun): one unit per 8 (query, index) cells
    uint64_t n_active_idx = 0;
    for (auto& ix : ctx->idx) n_active_idx += ix.alive && ix.committed && ix.active;
    if (units_cap < 4096)
        units_cap = std::min<uint64_t>(std::max<uint64_t>(1u << 16, (uint64_t)ctx->nq * n_active_idx / 8), 1ull << 24);
    if (hits_cap < 4096) {  // kept hits per query: ~2 x top_n when a top-N cut applies (ties, a few matching batches)
        const uint64_t per_q = p->top_n ? std::min<uint64_t>(std::max<uint64_t>(2ull * p->top_n, 64), 256) : 128;
        hits_cap = std::min<uint64_t>(std::max<uint64_t>(1u << 20, per_q * ctx->nq), 1ull << 28);
    }
    if (const char* e = getenv("PHY_TEST_TINY_CAPS")) {  // tests: force the overflow -> regrow -> rerun path
        if (atoi(e) && !ctx->d_units.p) { units_cap = 4; hits_cap = 4; }
    }
    for (int attempt = 0; attempt < 3; attempt++) {
        PHY_TRY(phy_ensure(ctx, ctx->d_units, units_cap));
        PHY_TRY(phy_ensure(ctx, ctx->d_hits, hits_cap));
        PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
        PHY_TRY(phy_ensure(ctx, ctx->d_idx_bytes, ctx->idx.size() + 1));
        PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_idx_bytes.p, 0, (ctx->idx.size() + 1) * sizeof(unsigned long long), ctx->stream));
        PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_qcount.p, 0, (ctx->nq + 1) * sizeof(uint32_t), ctx->stream));
        GatherArgs a;
        a.indexes = ctx->d_indexes.p;
        a.qlist = ctx->d_qlist.p; a.n_q = n_fast10;
        a.koffs = ctx->d_koffs.p; a.nk = ctx->d_nk.p; a.T = ctx->d_T.p;
        a.hashes = ctx->d_hashes.p; a.total_kmers = ctx->total_kmers;
        a.top_n = p->top_n;
        a.prune = ctx->prune ? 1u : 0u;
        a.units = ctx->d_units.p; a.units_cap = ctx->d_units.cap;
        a.hits = ctx->d_hits.p; a.hits_cap = ctx->d_hits.cap;
        a.counters = ctx->d_counters.p; a.qcount = ctx->d_qcount.p;
        a.idx_bytes = ctx->d_idx_bytes.p;
        // dense (index slot, query) table so the unit list can be put in (index, query) order on
        // the device (deterministic output, no host sort); skipped when it would be huge
        const uint64_t cells = (uint64_t)ctx->idx.size() * ctx->nq;
        a.unit_flag = a.unit_id = nullptr;
        a.nq = ctx->nq;
        if (cells && cells <= (1ull << 27)) {
            PHY_TRY(phy_ensure(ctx, ctx->d_unit_flag, cells + 1));
            PHY_TRY(phy_ensure(ctx, ctx->d_unit_id, cells + 1));
            PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_unit_flag.p, 0, cells * sizeof(uint32_t), ctx->stream));
            a.unit_flag = ctx->d_unit_flag.p;
            a.unit_id = ctx->d_unit_id.p;
        }
        if (!fastq.empty()) {
            for (const IdxClass& k : classes) {  // persistent warps pull (query, index) units of the class
                a.class_idx = d_class + k.off;
                a.n_class_idx = (uint32_t)k.ids.size();
                a.class_hashes = k.h;
                for (int pass = 0; pass < 3; pass++) {   // 10-plane, 8-plane, 14-plane query classes
                    a.qlist = ctx->d_qlist.p + (pass == 0 ? 0 : pass == 1 ? n_fast10 : n_le1023);
                    a.n_q = pass == 0 ? n_fast10 : pass == 1 ? n_fast8 : n_fast14;
                    if (a.n_q == 0) continue;
                    PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p + 3, 0, sizeof(unsigned long long), ctx->stream));
                    const int rc = k.h > 1 ? dispatch_ring_planes<true>(pass, k.c, a, ctx->stream, ctx->n_sm)
                                           : dispatch_ring_planes<false>(pass, k.c, a, ctx->stream, ctx->n_sm);
                    if (rc != 0) {
                        phy_set_error(ctx, "cannot configure the ring gather kernel: %s",
                                      cudaGetErrorString(cudaGetLastError()));
                        return PHY_ERR_CUDA;
                    }
                    ctx->launches++;
                    PHY_CUDA(ctx, cudaGetLastError());
                }
            }
        }
        // general path: long queries against every index; every query against wide indexes
        for (size_t i = 0; i < ctx->idx.size(); i++) {
            const HostIndex& ix = ctx->idx[i];
            if (!ix.alive || !ix.committed || !ix.active) continue;
            const bool is_wide = ix.d.stride > PHY_CHUNK_BYTES;
            std::vector<uint32_t> qs = slowq;
            if (is_wide) qs.insert(qs.end(), fastq.begin(), fastq.end());
            if (qs.empty()) continue;
            // bounded scratch: process slots in groups of <= 256 MB of scores
            size_t per = std::max<size_t>(1, (size_t)(64u << 20) / std::max<uint32_t>(1, ix.d.n_docs));
            for (size_t s0 = 0; s0 < qs.size(); s0 += per) {
                std::vector<uint32_t> part(qs.begin() + s0, qs.begin() + std::min(qs.size(), s0 + per));
                PHY_TRY(phy_ensure(ctx, ctx->d_scores, part.size() * (size_t)ix.d.n_docs));
                PHY_TRY(phy_ensure(ctx, ctx->d_slotq, part.size()));
                PHY_TRY(phy_h2d(ctx, ctx->d_slotq.p, part.data(), part.size() * sizeof(uint32_t)));
                PHY_TRY(run_general(ctx, ix, (int)i, part, ctx->d_scores.p));
                select_scores_kernel<<<(unsigned)part.size(), 256, 0, ctx->stream>>>(
                    ctx->d_scores.p, ix.d.n_docs, ctx->d_slotq.p, ctx->d_T.p, ctx->d_nk.p, p->top_n,
                    ix.d.idx_id, ctx->d_units.p, ctx->d_units.cap, ctx->d_hits.p, ctx->d_hits.cap,
                    ctx->d_counters.p, ctx->d_qcount.p, a.unit_flag, a.unit_id, ctx->nq);
                ctx->launches++;
                PHY_CUDA(ctx, cudaGetLastError());
            }
        }
        unsigned long long cnt[6];
        ctx->h_idx_bytes.assign(ctx->idx.size() + 1, 0);
        PHY_CUDA(ctx, cudaMemcpyAsync(cnt, ctx->d_counters.p, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
        PHY_CUDA(ctx, cudaMemcpyAsync(ctx->h_idx_bytes.data(), ctx->d_idx_bytes.p,
                                      ctx->idx.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->hash_check_pending) PHY_TRY(phy_hash_error_of(ctx, cnt[2]));  // K1's error word rode along
        ctx->gathered_bytes = 0;
        for (unsigned long long b : ctx->h_idx_bytes) ctx->gathered_bytes += b;
        if (cnt[0] <= ctx->d_hits.cap && cnt[1] <= ctx->d_units.cap) {
            ctx->n_hits = cnt[0];
            ctx->n_units = cnt[1];
            ctx->units_ordered = false;
            if (a.unit_flag && ctx->n_units) {
                PHY_TRY(phy_order_units(ctx, cells));
                ctx->units_ordered = true;
            }
            return PHY_OK;
        }
        hits_cap = std::max<uint64_t>(cnt[0], hits_cap);
        units_cap = std::max<uint64_t>(cnt[1], units_cap);
    }
    phy_set_error(ctx, "result buffers could not be sized");
    return PHY_ERR_NOMEM;
}
