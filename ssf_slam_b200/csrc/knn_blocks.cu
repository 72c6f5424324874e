// Exact k-nearest-neighbour search with spatial pruning: same results, bit for bit, as the brute-force `ssf_knn`
// (distance = ((dx*dx)+(dy*dy))+(dz*dz) without FMA, order = (distance, index) lexicographic),
// for `pointutils.knn / three_nn` call sites ASF/utils/utils.py:229,291 and ASF/utils/soflow.py:387-391,406,1243,1461.
//
// build : one CTA per reference cloud sorts the points along a 30-bit Hilbert curve over isotropic cells (ssf_hilbert30 keys,
//         ssf_cta_sort_u64 in shared memory), stores them as float4 (x, y, z, original index) and the bounding box of every
//         block of 32 consecutive points.  Any order is exact; this one makes the blocks compact (3.4 visits per query).
//         Above 64 blocks the search takes the bounds in two steps (knn_blocks_search_sb_kernel below).
// search: one warp per query.  Lane l owns the lower bounds lb(q, box) of blocks l, l+32, ...  The nearest block's 32
//         points (lane = point, one coalesced 512-byte load) are sorted across the lanes by a bitonic network and become
//         the candidate list (lane j = j-th best); a few more nearest-first visits tighten the k-th distance, then every
//         remaining block whose bound does not exceed it is visited in index order (re-tested as the bound shrinks).
//         Survivors of a visited block are merged into the lane-distributed list with ballot / shuffle-up inserts; when a
//         block has many of them (the first blocks of a query) the block is sorted across the lanes instead and merged with
//         the list by one bitonic merge (min of the list and the reversed block = the 32 smallest keys).
// Exactness: every box bound is computed with the SAME rounded operations as the point distance and each of them
// (fsub, fmul, fadd) is monotone, so bound <= distance of every point inside the box holds in floating point, not
// just in exact arithmetic; blocks are skipped only on bound > kth (strict), so index ties are never lost.
#include "ssf_common.cuh"
#include <math_constants.h>
#include <cstdlib>

#ifdef SSF_CV_TRACE
__device__ unsigned long long g_knn_stat[8];   // queries, visits, inserts (a merge = 8), sweep re-checks, merges, survivors
#define KSTAT(i, n) do { if (lane == 0) atomicAdd(&g_knn_stat[i], (unsigned long long)(n)); } while (0)
#else
#define KSTAT(i, n) do {} while (0)
#endif

namespace {

constexpr int KB_BUILD_T = 1024;
// Search CTAs are small (4 warps, <= 40 registers per thread, 8 KB of shared memory) ON PURPOSE: one such CTA still fits on an SM
// that a persistent tensor-core CTA of another stream occupies (cost volume: 576 threads x 96 registers, 207 KB), so the searches
// of one batch run in the issue slots the tensor kernels of another batch leave idle (they issue ~40 % of the time).
constexpr int KB_SEARCH_T = 128;                       // threads per search CTA
constexpr int KB_QPB = (KB_SEARCH_T / 32) * 8;         // queries per search CTA (8 per warp)

// ws layout per cloud (floats): pts4 [npad][4] | box_lo [nblk][4] | box_hi [nblk][4]
__global__ void __launch_bounds__(KB_BUILD_T) knn_blocks_build_kernel(const float* __restrict__ ref, int Nr, int npow2, int npad,
                                                                       int nblk, float* __restrict__ ws) {
    extern __shared__ __align__(16) unsigned long long skv[];   // [npow2] Hilbert key << 32 | point index
    __shared__ float sred[6][32];
    __shared__ float sbb[6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* rp = ref + (size_t)blockIdx.x * Nr * 3;
    float* wsb = ws + (size_t)blockIdx.x * ((size_t)npad * 4 + (size_t)nblk * 8);
    // cloud bounding box
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int i = tid; i < Nr; i += KB_BUILD_T)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = rp[3 * i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if (lane == 0) {
            sred[c][warp] = mn[c];
            sred[3 + c][warp] = mx[c];
        }
    }
    __syncthreads();
    if (tid < 6) {
        float v = sred[tid][0];
        for (int w = 1; w < KB_BUILD_T / 32; ++w) v = tid < 3 ? fminf(v, sred[tid][w]) : fmaxf(v, sred[tid][w]);
        sbb[tid] = v;
    }
    __syncthreads();
    // one scale for the three axes: the curve's cells are cubes, so the 32-point blocks are compact in the metric the search prunes
    // with (per-axis scales make 4 cm x 9 m slivers out of a 200 m x 23 m x 10 m LiDAR sweep and cost ~70 % more block visits)
    float sc[3];
    {
        const float ext = fmaxf(sbb[3] - sbb[0], fmaxf(sbb[4] - sbb[1], sbb[5] - sbb[2]));
        sc[0] = sc[1] = sc[2] = ext > 0.f ? 1023.0f / ext : 0.f;
    }
    for (int i = tid; i < npow2; i += KB_BUILD_T) {
        unsigned key = 0xFFFFFFFFu;
        if (i < Nr) {
            unsigned q[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float t = (rp[3 * i + c] - sbb[c]) * sc[c];
                q[c] = (unsigned)fminf(fmaxf(t, 0.f), 1023.f);
            }
            key = ssf_hilbert30(q[0], q[1], q[2]);
        }
        skv[i] = (unsigned long long)key << 32 | (unsigned)i;
    }
    __syncthreads();
    // sort on (key, index): the index breaks key ties, so the layout is a deterministic function of the cloud
    ssf_cta_sort_u64(skv, npow2, tid, KB_BUILD_T);
    // sorted points (padding: +inf coordinates, index INT_MAX -> never selected) and per-block boxes
    float4* pts4 = reinterpret_cast<float4*>(wsb);
    float4* blo = pts4 + npad;
    float4* bhi = blo + nblk;
    for (int i = tid; i < npad; i += KB_BUILD_T) {
        float4 p = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, __int_as_float(0x7fffffff));
        float lo3[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi3[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        if (i < Nr) {
            const int o = (int)(unsigned)skv[i];
            p = make_float4(rp[3 * o], rp[3 * o + 1], rp[3 * o + 2], __int_as_float(o));
            lo3[0] = hi3[0] = p.x; lo3[1] = hi3[1] = p.y; lo3[2] = hi3[2] = p.z;
        }
        pts4[i] = p;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo3[c] = fminf(lo3[c], __shfl_xor_sync(0xffffffffu, lo3[c], o));
                hi3[c] = fmaxf(hi3[c], __shfl_xor_sync(0xffffffffu, hi3[c], o));
            }
        if (lane == 0) {
            blo[i >> 5] = make_float4(lo3[0], lo3[1], lo3[2], 0.f);
            bhi[i >> 5] = make_float4(hi3[0], hi3[1], hi3[2], 0.f);
        }
    }
}

template <int NBL>
__global__ void __launch_bounds__(KB_SEARCH_T, 12) knn_blocks_search_kernel(int k, const float* __restrict__ query, const float* __restrict__ qadd,
                                                                const float* __restrict__ ws, int Nq, int npad, int nblk,
                                                                float* __restrict__ dist, int* __restrict__ idx) {
    // survivors of a block from which the sort-merge beats one-by-one insertion (measured: 6 for <= 64 blocks, 10-16 above)
    constexpr int merge_min = NBL == 2 ? 6 : 12;
    extern __shared__ __align__(16) float4 sbox[];   // [nblk] lo | [nblk] hi
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const float* wsb = ws + (size_t)b * ((size_t)npad * 4 + (size_t)nblk * 8);
    const float4* P = reinterpret_cast<const float4*>(wsb);
    {
        const float4* src = P + npad;
        for (int i = tid; i < 2 * nblk; i += KB_SEARCH_T) sbox[i] = __ldg(src + i);
    }
    __syncthreads();
    const int q_end = min(Nq, (int)(blockIdx.x + 1) * KB_QPB);
    for (int qi = blockIdx.x * KB_QPB + warp; qi < q_end; qi += KB_SEARCH_T / 32) {
        const float* qp = query + ((size_t)b * Nq + qi) * 3;
        float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        if (qadd != nullptr) {
            const float* ap = qadd + ((size_t)b * Nq + qi) * 3;
            qx = __fadd_rn(qx, __ldg(ap));
            qy = __fadd_rn(qy, __ldg(ap + 1));
            qz = __fadd_rn(qz, __ldg(ap + 2));
        }
        float lb[NBL];
#pragma unroll
        for (int s = 0; s < NBL; ++s) {
            const int blk = s * 32 + lane;
            lb[s] = CUDART_INF_F;
            if (blk < nblk) {
                const float4 lo = sbox[blk], hi = sbox[nblk + blk];
                const float gx = fmaxf(0.f, fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)));
                const float gy = fmaxf(0.f, fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)));
                const float gz = fmaxf(0.f, fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)));
                lb[s] = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
            }
        }
        // (distance, index) packed into one 64-bit key: distances are non-negative, so their bit patterns order like the
        // values and an unsigned 64-bit compare IS the lexicographic (distance, index) order of the specification
        unsigned long long list_k = 0x7f8000007fffffffull, kth = 0x7f8000007fffffffull;   // (+inf, INT_MAX)
        float kth_d = CUDART_INF_F;
        auto pack = [](float d, int i) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i; };
        // evaluates block `blk` (lane = point) and merges the survivors into the lane-distributed sorted list
        KSTAT(0, 1);
        // ascending bitonic sort of one key per lane
        auto sort32 = [&](unsigned long long key) -> unsigned long long {
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, j);
                    const bool take_min = ((lane & j) == 0) == ((lane & kk) == 0);
                    if (take_min == (ok < key)) key = ok;
                }
            }
            return key;
        };
        auto visit = [&](int blk) {
            KSTAT(1, 1);
            const float4 p = __ldg(P + (size_t)blk * 32 + lane);
            unsigned long long key = pack(ssf_sqdist(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
            unsigned mask = __ballot_sync(0xffffffffu, key < kth);
            KSTAT(5, __popc(mask));
            if (__popc(mask) >= merge_min) {
                KSTAT(4, 1);
                // many survivors (the first blocks of a query): sort the block and merge it with the list in one go -- the
                // element-wise min of the list and the reversed block is a bitonic sequence holding the 32 smallest keys
                key = sort32(key);
                const unsigned long long rev = __shfl_sync(0xffffffffu, key, 31 - lane);
                unsigned long long m = rev < list_k ? rev : list_k;
#pragma unroll
                for (int j = 16; j > 0; j >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, m, j);
                    if (((lane & j) == 0) == (ok < m)) m = ok;
                }
                list_k = m;
                kth = __shfl_sync(0xffffffffu, list_k, k - 1);
                KSTAT(2, 8);
            } else {
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const unsigned long long ck = __shfl_sync(0xffffffffu, key, src);
                    if (ck >= kth) continue;   // warp-uniform: the k-th key tightened meanwhile
                    KSTAT(2, 1);
                    const int pos = __popc(__ballot_sync(0xffffffffu, list_k < ck));
                    const unsigned long long uk = __shfl_up_sync(0xffffffffu, list_k, 1);
                    list_k = lane == pos ? ck : (lane > pos ? uk : list_k);
                    kth = __shfl_sync(0xffffffffu, list_k, k - 1);
                }
            }
            kth_d = __uint_as_float((unsigned)(kth >> 32));
        };
        // warp-wide argmin of the remaining bounds; marks the winner visited.  Returns -1 when nothing is left.
        auto pop_nearest = [&](float& best_out) -> int {
            float best = lb[0];
            int bs = 0;
#pragma unroll
            for (int s = 1; s < NBL; ++s)
                if (lb[s] < best) {
                    best = lb[s];
                    bs = s;
                }
            // warp-wide minimum by two integer reductions (bounds are non-negative floats: their bit patterns order like the
            // values): the smallest bound, then the lowest block index among the lanes that hold it
            const unsigned mn = __reduce_min_sync(0xffffffffu, __float_as_uint(best));
            const int bblk = (int)__reduce_min_sync(0xffffffffu, __float_as_uint(best) == mn ? (unsigned)(bs * 32 + lane) : 0x7fffffffu);
            best = __uint_as_float(mn);
            best_out = best;
            if (best == CUDART_INF_F) return -1;
#pragma unroll
            for (int s = 0; s < NBL; ++s)
                if (s == (bblk >> 5) && lane == (bblk & 31)) lb[s] = CUDART_INF_F;
            return bblk;
        };
        // 1) seed: the nearest block's 32 points, sorted across the lanes by a bitonic network, become the list
        {
            float best;
            const int blk = pop_nearest(best);   // >= 0: there is at least one block
            const float4 p = __ldg(P + (size_t)blk * 32 + lane);
            list_k = sort32(pack(ssf_sqdist(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w)));
            kth = __shfl_sync(0xffffffffu, list_k, k - 1);
            kth_d = __uint_as_float((unsigned)(kth >> 32));
        }
        // 2) a few nearest-first visits tighten the k-th distance quickly
#pragma unroll 1
        for (int rep = 0; rep < 3; ++rep) {
            float best;
            const int blk = pop_nearest(best);
            if (blk < 0 || best > kth_d) {   // nothing left, or no remaining box can hold a better point
                if (blk >= 0) {              // (put the popped block back: the sweep below re-tests it)
#pragma unroll
                    for (int s = 0; s < NBL; ++s)
                        if (s == (blk >> 5) && lane == (blk & 31)) lb[s] = best;
                }
                break;
            }
            visit(blk);
        }
        // 3) sweep: every remaining block whose bound does not exceed the (shrinking) k-th distance, in index order
#pragma unroll
        for (int s = 0; s < NBL; ++s) {
            unsigned m = __ballot_sync(0xffffffffu, lb[s] <= kth_d && lb[s] != CUDART_INF_F);
            while (m) {
                const int bl = __ffs(m) - 1;
                m &= m - 1;
                const float blb = __shfl_sync(0xffffffffu, lb[s], bl);
                KSTAT(3, 1);
                if (blb > kth_d) continue;   // warp-uniform
                visit(s * 32 + bl);
            }
        }
        if (lane < k) {
            const size_t o = ((size_t)b * Nq + qi) * k + lane;
            idx[o] = (int)(unsigned)(list_k & 0xffffffffull);
            if (dist != nullptr) dist[o] = __fsqrt_rn(__uint_as_float((unsigned)(list_k >> 32)));
        }
    }
}

// Search for 64 < nblk <= 512 blocks (2048 < Nr <= 16384).  Holding every block bound in registers (8-16 per lane) makes the
// bound computation and every arg-min over them the largest part of a query once the Hilbert order has cut the visits to 3-4
// blocks.  So the bounds are taken in two steps: a lane first bounds ONE super-box (the union of SUPB consecutive blocks: a
// compact run of the curve; the unions are made in the prologue from the block boxes already in shared memory), the warp
// picks the 32 / SUPB nearest super-boxes and spreads their 32 blocks over the lanes -- one bound, one register, and an
// arg-min is a reduction plus a ballot.  More super-boxes are opened only while one is at most the k-th distance away
// (rare).  Every block whose bound does not exceed the final k-th distance is still visited: a block of a super-box that was
// never opened is at least as far as the super-box, which was beyond the k-th distance when it was last tested, and that
// distance only shrinks.  Same keys and list operations as above, so the same result.
template <int SUPB>
__global__ void __launch_bounds__(KB_SEARCH_T, 12) knn_blocks_search_sb_kernel(int k, const float* __restrict__ query, const float* __restrict__ qadd,
                                                                   const float* __restrict__ ws, int Nq, int npad, int nblk,
                                                                   float* __restrict__ dist, int* __restrict__ idx) {
    constexpr int EXP = 32 / SUPB;      // super-boxes opened per round
    constexpr int merge_min = 12;
    extern __shared__ __align__(16) float4 sbox[];   // [nblk] lo | [nblk] hi | [32] super lo | [32] super hi
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const float* wsb = ws + (size_t)b * ((size_t)npad * 4 + (size_t)nblk * 8);
    const float4* P = reinterpret_cast<const float4*>(wsb);
    {
        const float4* src = P + npad;
        for (int i = tid; i < 2 * nblk; i += KB_SEARCH_T) sbox[i] = __ldg(src + i);
    }
    __syncthreads();
    float4* ssup = sbox + 2 * nblk;
    if (tid < 32) {
        float4 lo = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f), hi = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, 0.f);
        for (int u = 0; u < SUPB; ++u) {
            const int blk = tid * SUPB + u;
            if (blk < nblk) {
                const float4 l = sbox[blk], h = sbox[nblk + blk];
                lo.x = fminf(lo.x, l.x); lo.y = fminf(lo.y, l.y); lo.z = fminf(lo.z, l.z);
                hi.x = fmaxf(hi.x, h.x); hi.y = fmaxf(hi.y, h.y); hi.z = fmaxf(hi.z, h.z);
            }
        }
        ssup[tid] = lo;        // an empty super-box keeps (+inf, -inf): its bound is +inf
        ssup[32 + tid] = hi;
    }
    __syncthreads();
    auto box_bound = [](const float4 lo, const float4 hi, float qx, float qy, float qz) -> float {
        const float gx = fmaxf(0.f, fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)));
        const float gy = fmaxf(0.f, fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)));
        const float gz = fmaxf(0.f, fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)));
        return __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    };
    constexpr unsigned INF_BITS = 0x7f800000u;
    const int q_end = min(Nq, (int)(blockIdx.x + 1) * KB_QPB);
    for (int qi = blockIdx.x * KB_QPB + warp; qi < q_end; qi += KB_SEARCH_T / 32) {
        const float* qp = query + ((size_t)b * Nq + qi) * 3;
        float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        if (qadd != nullptr) {
            const float* ap = qadd + ((size_t)b * Nq + qi) * 3;
            qx = __fadd_rn(qx, __ldg(ap));
            qy = __fadd_rn(qy, __ldg(ap + 1));
            qz = __fadd_rn(qz, __ldg(ap + 2));
        }
        KSTAT(0, 1);
        unsigned slb = __float_as_uint(box_bound(ssup[lane], ssup[32 + lane], qx, qy, qz));   // bounds are >= 0: bits order like values
        unsigned long long list_k = 0x7f8000007fffffffull, kth = 0x7f8000007fffffffull;   // (+inf, INT_MAX)
        float kth_d = CUDART_INF_F;
        auto pack = [](float d, int i) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i; };
        auto sort32 = [&](unsigned long long key) -> unsigned long long {
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, j);
                    const bool take_min = ((lane & j) == 0) == ((lane & kk) == 0);
                    if (take_min == (ok < key)) key = ok;
                }
            }
            return key;
        };
        auto visit = [&](int blk) {
            KSTAT(1, 1);
            const float4 p = __ldg(P + (size_t)blk * 32 + lane);
            unsigned long long key = pack(ssf_sqdist(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
            unsigned mask = __ballot_sync(0xffffffffu, key < kth);
            KSTAT(5, __popc(mask));
            if (__popc(mask) >= merge_min) {
                KSTAT(4, 1);
                key = sort32(key);
                const unsigned long long rev = __shfl_sync(0xffffffffu, key, 31 - lane);
                unsigned long long m = rev < list_k ? rev : list_k;
#pragma unroll
                for (int j = 16; j > 0; j >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, m, j);
                    if (((lane & j) == 0) == (ok < m)) m = ok;
                }
                list_k = m;
                kth = __shfl_sync(0xffffffffu, list_k, k - 1);
                KSTAT(2, 8);
            } else {
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const unsigned long long ck = __shfl_sync(0xffffffffu, key, src);
                    if (ck >= kth) continue;   // warp-uniform: the k-th key tightened meanwhile
                    KSTAT(2, 1);
                    const int pos = __popc(__ballot_sync(0xffffffffu, list_k < ck));
                    const unsigned long long uk = __shfl_up_sync(0xffffffffu, list_k, 1);
                    list_k = lane == pos ? ck : (lane > pos ? uk : list_k);
                    kth = __shfl_sync(0xffffffffu, list_k, k - 1);
                }
            }
            kth_d = __uint_as_float((unsigned)(kth >> 32));
        };
        bool seeded = false;
#pragma unroll 1
        for (;;) {
            // open the (up to) EXP nearest super-boxes that can still hold a better point; lanes [e SUPB, (e + 1) SUPB) take the e-th
            int mysb = -1, opened = 0;
#pragma unroll
            for (int e = 0; e < EXP; ++e) {
                const unsigned mn = __reduce_min_sync(0xffffffffu, slb);
                if (mn == INF_BITS || __uint_as_float(mn) > kth_d) break;   // warp-uniform
                const int L = __ffs(__ballot_sync(0xffffffffu, slb == mn)) - 1;
                if (lane == L) slb = INF_BITS;
                if (lane / SUPB == e) mysb = L;
                ++opened;
            }
            if (opened == 0) break;
            int blk = -1;
            unsigned lbk = INF_BITS;
            if (mysb >= 0 && mysb * SUPB + lane % SUPB < nblk) {
                blk = mysb * SUPB + lane % SUPB;
                lbk = __float_as_uint(box_bound(sbox[blk], sbox[nblk + blk], qx, qy, qz));
            }
            // nearest-first: the seed block (its 32 points, sorted across the lanes, become the list) and a few visits that
            // tighten the k-th distance quickly
            const int nrep = seeded ? 1 : 4;
#pragma unroll 1
            for (int rep = 0; rep < nrep; ++rep) {
                const unsigned mn = __reduce_min_sync(0xffffffffu, lbk);
                if (mn == INF_BITS || __uint_as_float(mn) > kth_d) break;
                const int L = __ffs(__ballot_sync(0xffffffffu, lbk == mn)) - 1;
                const int vb = __shfl_sync(0xffffffffu, blk, L);
                if (lane == L) lbk = INF_BITS;
                if (!seeded) {
                    const float4 p = __ldg(P + (size_t)vb * 32 + lane);
                    list_k = sort32(pack(ssf_sqdist(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w)));
                    kth = __shfl_sync(0xffffffffu, list_k, k - 1);
                    kth_d = __uint_as_float((unsigned)(kth >> 32));
                    seeded = true;
                } else {
                    visit(vb);
                }
            }
            // the rest of the opened blocks in lane order under the (shrinking) k-th distance
            unsigned m = __ballot_sync(0xffffffffu, lbk != INF_BITS && __uint_as_float(lbk) <= kth_d);
            while (m) {
                const int L = __ffs(m) - 1;
                m &= m - 1;
                const float blb = __uint_as_float(__shfl_sync(0xffffffffu, lbk, L));
                KSTAT(3, 1);
                if (blb > kth_d) continue;   // warp-uniform
                visit(__shfl_sync(0xffffffffu, blk, L));
            }
        }
        if (lane < k) {
            const size_t o = ((size_t)b * Nq + qi) * k + lane;
            idx[o] = (int)(unsigned)(list_k & 0xffffffffull);
            if (dist != nullptr) dist[o] = __fsqrt_rn(__uint_as_float((unsigned)(list_k >> 32)));
        }
    }
}

// Brute-force scan, one warp per KS_Q queries (lane = reference point, 32 consecutive points per step), for reference clouds
// too large for the block index when there are too few queries to fill the machine with one thread each (BASELINE config 4:
// 2048 centres in 65536 points).  Same keys, same list operations and therefore the same result as every other kNN here.
constexpr int KS_Q = 2;

__global__ void __launch_bounds__(256) knn_warp_scan_kernel(int k, const float* __restrict__ query, const float* __restrict__ qadd,
                                                            const float* __restrict__ ref, int Nq, int Nr,
                                                            float* __restrict__ dist, int* __restrict__ idx) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
    const int q0 = ((int)blockIdx.x * 8 + warp) * KS_Q;
    if (q0 >= Nq) return;
    auto pack = [](float d, int i) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i; };
    float qx[KS_Q], qy[KS_Q], qz[KS_Q];
    unsigned long long list_k[KS_Q], kth[KS_Q];
#pragma unroll
    for (int u = 0; u < KS_Q; ++u) {
        const size_t qi = (size_t)b * Nq + min(q0 + u, Nq - 1);
        qx[u] = __ldg(query + qi * 3); qy[u] = __ldg(query + qi * 3 + 1); qz[u] = __ldg(query + qi * 3 + 2);
        if (qadd != nullptr) {
            qx[u] = __fadd_rn(qx[u], __ldg(qadd + qi * 3));
            qy[u] = __fadd_rn(qy[u], __ldg(qadd + qi * 3 + 1));
            qz[u] = __fadd_rn(qz[u], __ldg(qadd + qi * 3 + 2));
        }
        list_k[u] = kth[u] = 0x7f8000007fffffffull;   // (+inf, INT_MAX)
    }
    const float* P = ref + (size_t)b * Nr * 3;
    for (int base = 0; base < Nr; base += 32) {
        const int r = base + lane;
        const bool in = r < Nr;
        const int rc = in ? r : Nr - 1;
        const float x = __ldg(P + 3 * rc), y = __ldg(P + 3 * rc + 1), z = __ldg(P + 3 * rc + 2);
#pragma unroll
        for (int u = 0; u < KS_Q; ++u) {
            unsigned long long key = in ? pack(ssf_sqdist(qx[u], qy[u], qz[u], x, y, z), r) : 0x7f8000007fffffffull;
            unsigned mask = __ballot_sync(0xffffffffu, key < kth[u]);
            if (__popc(mask) >= 12) {   // sort the 32 keys across the lanes and merge them with the list (see the block search)
#pragma unroll
                for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                    for (int j = kk >> 1; j > 0; j >>= 1) {
                        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, j);
                        const bool take_min = ((lane & j) == 0) == ((lane & kk) == 0);
                        if (take_min == (ok < key)) key = ok;
                    }
                }
                const unsigned long long rev = __shfl_sync(0xffffffffu, key, 31 - lane);
                unsigned long long m = rev < list_k[u] ? rev : list_k[u];
#pragma unroll
                for (int j = 16; j > 0; j >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, m, j);
                    if (((lane & j) == 0) == (ok < m)) m = ok;
                }
                list_k[u] = m;
                kth[u] = __shfl_sync(0xffffffffu, m, k - 1);
            } else {
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const unsigned long long ck = __shfl_sync(0xffffffffu, key, src);
                    if (ck >= kth[u]) continue;   // warp-uniform
                    const int pos = __popc(__ballot_sync(0xffffffffu, list_k[u] < ck));
                    const unsigned long long uk = __shfl_up_sync(0xffffffffu, list_k[u], 1);
                    list_k[u] = lane == pos ? ck : (lane > pos ? uk : list_k[u]);
                    kth[u] = __shfl_sync(0xffffffffu, list_k[u], k - 1);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < KS_Q; ++u) {
        if (q0 + u < Nq && lane < k) {
            const size_t o = ((size_t)b * Nq + q0 + u) * k + lane;
            idx[o] = (int)(unsigned)(list_k[u] & 0xffffffffull);
            if (dist != nullptr) dist[o] = __fsqrt_rn(__uint_as_float((unsigned)(list_k[u] >> 32)));
        }
    }
}


// ------------------------------------------------------------------ clouds above 16384 points (BASELINE config 5: N = 65536)
// The curve sort no longer fits in shared memory, and one lane cannot hold the bounds of all blocks.  So:
//  build : one CTA per cloud, stable LSD radix sort of the (key, index) pairs in global scratch (eight 4-bit passes; thread t
//          owns a contiguous chunk, per-(digit, thread) counters in shared memory, one block scan per pass -- stable, so equal
//          keys stay in index order and the layout is the same deterministic function of the cloud as the bitonic build's),
//          then the same sorted float4 points + per-block boxes, plus the boxes of SUPER-BLOCKS of 32 consecutive blocks.
//  search: two levels.  A lane owns the bounds of <= 4 super-blocks; the warp pops super-blocks nearest-first while their bound
//          does not exceed the k-th distance, and inside a popped super-block lane l owns block l: the nearest few blocks are
//          visited first, the rest in index order while their bound does not exceed the (shrinking) k-th distance.  Same
//          keys, same list operations, same strict-skip rule as the one-level search: bit-identical to the brute-force scan.
// ws layout per cloud (floats): pts4 [npad][4] | box_lo [nblk][4] | box_hi [nblk][4] | sb_lo [nsb][4] | sb_hi [nsb][4] | scratch [4 npad]
constexpr int KB_RADIX_T = 1024;

__global__ void __launch_bounds__(KB_RADIX_T) knn_blocks_build_large_kernel(const float* __restrict__ ref, int Nr, int npad, int nblk,
                                                                            int nsb, long long ws_per_cloud, float* __restrict__ ws) {
    extern __shared__ __align__(16) int s_cnt[];          // [16][KB_RADIX_T]
    __shared__ float sred[6][32];
    __shared__ float sbb[6];
    __shared__ int s_scan[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* rp = ref + (size_t)blockIdx.x * Nr * 3;
    float* wsb = ws + (size_t)blockIdx.x * ws_per_cloud;
    float4* pts4 = reinterpret_cast<float4*>(wsb);
    float4* blo = pts4 + npad;
    float4* bhi = blo + nblk;
    float4* slo = bhi + nblk;
    float4* shi = slo + nsb;
    unsigned* keyA = reinterpret_cast<unsigned*>(shi + nsb);
    unsigned* keyB = keyA + npad;
    int* valA = reinterpret_cast<int*>(keyB + npad);
    int* valB = valA + npad;
    // cloud bounding box
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int i = tid; i < Nr; i += KB_RADIX_T)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = rp[3 * i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if (lane == 0) {
            sred[c][warp] = mn[c];
            sred[3 + c][warp] = mx[c];
        }
    }
    __syncthreads();
    if (tid < 6) {
        float v = sred[tid][0];
        for (int w = 1; w < KB_RADIX_T / 32; ++w) v = tid < 3 ? fminf(v, sred[tid][w]) : fmaxf(v, sred[tid][w]);
        sbb[tid] = v;
    }
    __syncthreads();
    // one scale for the three axes: the curve's cells are cubes, so the 32-point blocks are compact in the metric the search prunes
    // with (per-axis scales make 4 cm x 9 m slivers out of a 200 m x 23 m x 10 m LiDAR sweep and cost ~70 % more block visits)
    float sc[3];
    {
        const float ext = fmaxf(sbb[3] - sbb[0], fmaxf(sbb[4] - sbb[1], sbb[5] - sbb[2]));
        sc[0] = sc[1] = sc[2] = ext > 0.f ? 1023.0f / ext : 0.f;
    }
    for (int i = tid; i < npad; i += KB_RADIX_T) {
        unsigned key = 0xFFFFFFFFu;      // padding sorts last (real keys have 30 bits)
        if (i < Nr) {
            unsigned q[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float t = (rp[3 * i + c] - sbb[c]) * sc[c];
                q[c] = (unsigned)fminf(fmaxf(t, 0.f), 1023.f);
            }
            key = ssf_hilbert30(q[0], q[1], q[2]);
        }
        keyA[i] = key;
        valA[i] = i;
    }
    __syncthreads();
    // stable LSD radix sort, 4 bits per pass; thread t owns elements [t * chunk, t * chunk + chunk)
    const int chunk = (npad + KB_RADIX_T - 1) / KB_RADIX_T;
    const int e0 = min(npad, tid * chunk), e1 = min(npad, e0 + chunk);
    unsigned* kin = keyA; unsigned* kout = keyB;
    int* vin = valA; int* vout = valB;
    for (int shift = 0; shift < 32; shift += 4) {
#pragma unroll
        for (int d = 0; d < 16; ++d) s_cnt[d * KB_RADIX_T + tid] = 0;
        for (int e = e0; e < e1; ++e) s_cnt[((kin[e] >> shift) & 15u) * KB_RADIX_T + tid] += 1;
        __syncthreads();
        // exclusive scan of the 16 * KB_RADIX_T counters in (digit, thread) order: thread t scans entries [16 t, 16 t + 16)
        int loc[16], sum = 0;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            loc[u] = sum;
            sum += s_cnt[tid * 16 + u];
        }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_scan[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_scan[lane] = w;
        }
        __syncthreads();
        const int base = inc - sum + (warp > 0 ? s_scan[warp - 1] : 0);
#pragma unroll
        for (int u = 0; u < 16; ++u) s_cnt[tid * 16 + u] = base + loc[u];
        __syncthreads();
        for (int e = e0; e < e1; ++e) {
            const unsigned kk = kin[e];
            const int pos = s_cnt[((kk >> shift) & 15u) * KB_RADIX_T + tid]++;
            kout[pos] = kk;
            vout[pos] = vin[e];
        }
        __syncthreads();
        unsigned* tk = kin; kin = kout; kout = tk;
        int* tv = vin; vin = vout; vout = tv;
    }
    // sorted points (padding: +inf coordinates, index INT_MAX -> never selected) and per-block boxes (after 8 passes: in A)
    for (int i = tid; i < npad; i += KB_RADIX_T) {
        float4 p = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, __int_as_float(0x7fffffff));
        float lo3[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi3[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        const int o = vin[i];
        if (o < Nr && kin[i] != 0xFFFFFFFFu) {
            p = make_float4(rp[3 * o], rp[3 * o + 1], rp[3 * o + 2], __int_as_float(o));
            lo3[0] = hi3[0] = p.x; lo3[1] = hi3[1] = p.y; lo3[2] = hi3[2] = p.z;
        }
        pts4[i] = p;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) {
                lo3[c] = fminf(lo3[c], __shfl_xor_sync(0xffffffffu, lo3[c], o2));
                hi3[c] = fmaxf(hi3[c], __shfl_xor_sync(0xffffffffu, hi3[c], o2));
            }
        if (lane == 0) {
            blo[i >> 5] = make_float4(lo3[0], lo3[1], lo3[2], 0.f);
            bhi[i >> 5] = make_float4(hi3[0], hi3[1], hi3[2], 0.f);
        }
    }
    __syncthreads();   // block boxes (global, written by this CTA) are read back below
    for (int sb = warp; sb < nsb; sb += KB_RADIX_T / 32) {
        const int blk = sb * 32 + lane;
        float lo3[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi3[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        if (blk < nblk) {
            const float4 l = blo[blk], h = bhi[blk];
            lo3[0] = l.x; lo3[1] = l.y; lo3[2] = l.z; hi3[0] = h.x; hi3[1] = h.y; hi3[2] = h.z;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) {
                lo3[c] = fminf(lo3[c], __shfl_xor_sync(0xffffffffu, lo3[c], o2));
                hi3[c] = fmaxf(hi3[c], __shfl_xor_sync(0xffffffffu, hi3[c], o2));
            }
        if (lane == 0) {
            slo[sb] = make_float4(lo3[0], lo3[1], lo3[2], 0.f);
            shi[sb] = make_float4(hi3[0], hi3[1], hi3[2], 0.f);
        }
    }
}

template <int NSB>
__global__ void __launch_bounds__(KB_SEARCH_T, 12) knn_blocks_search2_kernel(int k, const float* __restrict__ query, const float* __restrict__ qadd,
                                                                            const float* __restrict__ ws, long long ws_per_cloud, int Nq, int npad,
                                                                            int nblk, int nsb, float* __restrict__ dist, int* __restrict__ idx) {
    extern __shared__ __align__(16) float4 sbox[];   // [nsb] lo | [nsb] hi of the super-blocks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const float* wsb = ws + (size_t)b * ws_per_cloud;
    const float4* P = reinterpret_cast<const float4*>(wsb);
    const float4* BLO = P + npad;
    const float4* BHI = BLO + nblk;
    {
        const float4* src = BHI + nblk;
        for (int i = tid; i < 2 * nsb; i += KB_SEARCH_T) sbox[i] = __ldg(src + i);
    }
    __syncthreads();
    auto box_bound = [](const float4 lo, const float4 hi, float qx, float qy, float qz) -> float {
        const float gx = fmaxf(0.f, fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)));
        const float gy = fmaxf(0.f, fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)));
        const float gz = fmaxf(0.f, fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)));
        return __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    };
    const int q_end = min(Nq, (int)(blockIdx.x + 1) * KB_QPB);
    for (int qi = blockIdx.x * KB_QPB + warp; qi < q_end; qi += KB_SEARCH_T / 32) {
        const float* qp = query + ((size_t)b * Nq + qi) * 3;
        float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        if (qadd != nullptr) {
            const float* ap = qadd + ((size_t)b * Nq + qi) * 3;
            qx = __fadd_rn(qx, __ldg(ap));
            qy = __fadd_rn(qy, __ldg(ap + 1));
            qz = __fadd_rn(qz, __ldg(ap + 2));
        }
        float slb[NSB];
#pragma unroll
        for (int s = 0; s < NSB; ++s) {
            const int sb = s * 32 + lane;
            slb[s] = sb < nsb ? box_bound(sbox[sb], sbox[nsb + sb], qx, qy, qz) : CUDART_INF_F;
        }
        unsigned long long list_k = 0x7f8000007fffffffull, kth = 0x7f8000007fffffffull;   // (+inf, INT_MAX)
        float kth_d = CUDART_INF_F;
        auto pack = [](float d, int i) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i; };
        auto sort32 = [&](unsigned long long key) -> unsigned long long {
#pragma unroll
            for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, j);
                    const bool take_min = ((lane & j) == 0) == ((lane & kk) == 0);
                    if (take_min == (ok < key)) key = ok;
                }
            }
            return key;
        };
        auto visit = [&](int blk) {
            const float4 p = __ldg(P + (size_t)blk * 32 + lane);
            unsigned long long key = pack(ssf_sqdist(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
            unsigned mask = __ballot_sync(0xffffffffu, key < kth);
            if (__popc(mask) >= 12) {
                key = sort32(key);
                const unsigned long long rev = __shfl_sync(0xffffffffu, key, 31 - lane);
                unsigned long long m = rev < list_k ? rev : list_k;
#pragma unroll
                for (int j = 16; j > 0; j >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, m, j);
                    if (((lane & j) == 0) == (ok < m)) m = ok;
                }
                list_k = m;
                kth = __shfl_sync(0xffffffffu, list_k, k - 1);
            } else {
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const unsigned long long ck = __shfl_sync(0xffffffffu, key, src);
                    if (ck >= kth) continue;   // warp-uniform: the k-th key tightened meanwhile
                    const int pos = __popc(__ballot_sync(0xffffffffu, list_k < ck));
                    const unsigned long long uk = __shfl_up_sync(0xffffffffu, list_k, 1);
                    list_k = lane == pos ? ck : (lane > pos ? uk : list_k);
                    kth = __shfl_sync(0xffffffffu, list_k, k - 1);
                }
            }
            kth_d = __uint_as_float((unsigned)(kth >> 32));
        };
        bool first = true;
#pragma unroll 1
        for (;;) {
            // nearest remaining super-block (two integer reductions: bounds are non-negative floats)
            float best = slb[0];
            int bs = 0;
#pragma unroll
            for (int s = 1; s < NSB; ++s)
                if (slb[s] < best) {
                    best = slb[s];
                    bs = s;
                }
            const unsigned mn = __reduce_min_sync(0xffffffffu, __float_as_uint(best));
            if (__uint_as_float(mn) == CUDART_INF_F || __uint_as_float(mn) > kth_d) break;   // no remaining box can hold a better point
            const int sb = (int)__reduce_min_sync(0xffffffffu, __float_as_uint(best) == mn ? (unsigned)(bs * 32 + lane) : 0x7fffffffu);
#pragma unroll
            for (int s = 0; s < NSB; ++s)
                if (s == (sb >> 5) && lane == (sb & 31)) slb[s] = CUDART_INF_F;
            // inside the super-block: lane = block
            const int blk = sb * 32 + lane;
            float lbk = CUDART_INF_F;
            if (blk < nblk) lbk = box_bound(__ldg(BLO + blk), __ldg(BHI + blk), qx, qy, qz);
            // nearest-first visits tighten the k-th distance quickly (more of them in the first super-block of a query)
#pragma unroll 1
            for (int rep = 0; rep < (first ? 4 : 1); ++rep) {
                const unsigned bm = __reduce_min_sync(0xffffffffu, __float_as_uint(lbk));
                if (__uint_as_float(bm) == CUDART_INF_F || __uint_as_float(bm) > kth_d) break;
                const int bl = (int)__reduce_min_sync(0xffffffffu, __float_as_uint(lbk) == bm ? (unsigned)lane : 0x7fffffffu);
                if (lane == bl) lbk = CUDART_INF_F;
                visit(sb * 32 + bl);
            }
            first = false;
            unsigned m = __ballot_sync(0xffffffffu, lbk <= kth_d && lbk != CUDART_INF_F);
            while (m) {
                const int bl = __ffs(m) - 1;
                m &= m - 1;
                const float blb = __shfl_sync(0xffffffffu, lbk, bl);
                if (blb > kth_d) continue;   // warp-uniform
                visit(sb * 32 + bl);
            }
        }
        if (lane < k) {
            const size_t o = ((size_t)b * Nq + qi) * k + lane;
            idx[o] = (int)(unsigned)(list_k & 0xffffffffull);
            if (dist != nullptr) dist[o] = __fsqrt_rn(__uint_as_float((unsigned)(list_k >> 32)));
        }
    }
}


// Ball query through the same spatial index (either level count): blocks whose box bound exceeds r^2 cannot hold a point in
// range (the bound is computed with the distance's own rounded operations and is monotone, see the header), every other block
// is evaluated lane = point.  The reference semantics want the FIRST nsample hits in ascending INDEX order (ASF/SetCover.py:
// 39-63), while the blocks come in curve order: the warp keeps the nsample smallest hit indices in a lane-distributed sorted
// list (the kNN list with the index as the key) and counts every hit.  Same output as ssf_ball_query, bit for bit.
template <bool TWO_LEVEL>
__global__ void __launch_bounds__(KB_SEARCH_T, 12) ball_query_blocks_kernel(float r2, int nsample, const float* __restrict__ query,
                                                                           const float* __restrict__ ws, long long ws_per_cloud, int S, int npad,
                                                                           int nblk, int nsb, int* __restrict__ idx, int* __restrict__ cnt) {
    extern __shared__ __align__(16) float4 sbox[];   // one level: [nblk] lo | [nblk] hi;  two levels: the super-block boxes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const float* wsb = ws + (size_t)b * ws_per_cloud;
    const float4* P = reinterpret_cast<const float4*>(wsb);
    const float4* BLO = P + npad;
    const float4* BHI = BLO + nblk;
    const int nbox = TWO_LEVEL ? nsb : nblk;
    {
        const float4* src = TWO_LEVEL ? BHI + nblk : BLO;
        for (int i = tid; i < 2 * nbox; i += KB_SEARCH_T) sbox[i] = __ldg(src + i);
    }
    __syncthreads();
    auto box_bound = [](const float4 lo, const float4 hi, float qx, float qy, float qz) -> float {
        const float gx = fmaxf(0.f, fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)));
        const float gy = fmaxf(0.f, fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)));
        const float gz = fmaxf(0.f, fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)));
        return __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    };
    const int q_end = min(S, (int)(blockIdx.x + 1) * KB_QPB);
    for (int qi = blockIdx.x * KB_QPB + warp; qi < q_end; qi += KB_SEARCH_T / 32) {
        const float* qp = query + ((size_t)b * S + qi) * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        unsigned list_i = 0xffffffffu, kth = 0xffffffffu;   // lane j = j-th smallest hit index so far; kth = the nsample-th
        int total = 0;
        auto visit = [&](int blk) {
            const float4 p = __ldg(P + (size_t)blk * 32 + lane);
            const bool hit = ssf_sqdist(qx, qy, qz, p.x, p.y, p.z) <= r2;   // padding points sit at +inf: never in range
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            total += __popc(mask);
            const unsigned mine = (unsigned)__float_as_int(p.w);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const unsigned ci = __shfl_sync(0xffffffffu, mine, src);
                if (ci >= kth) continue;   // warp-uniform
                const int pos = __popc(__ballot_sync(0xffffffffu, list_i < ci));
                const unsigned up = __shfl_up_sync(0xffffffffu, list_i, 1);
                list_i = lane == pos ? ci : (lane > pos ? up : list_i);
                kth = __shfl_sync(0xffffffffu, list_i, nsample - 1);
            }
        };
        if (TWO_LEVEL) {
            for (int g0 = 0; g0 < nsb; g0 += 32) {
                const int sb = g0 + lane;
                const bool near = sb < nsb && box_bound(sbox[sb], sbox[nsb + sb], qx, qy, qz) <= r2;
                unsigned sm = __ballot_sync(0xffffffffu, near);
                while (sm) {
                    const int sl = __ffs(sm) - 1;
                    sm &= sm - 1;
                    const int blk = (g0 + sl) * 32 + lane;
                    const bool bn = blk < nblk && box_bound(__ldg(BLO + blk), __ldg(BHI + blk), qx, qy, qz) <= r2;
                    unsigned bm = __ballot_sync(0xffffffffu, bn);
                    while (bm) {
                        const int bl = __ffs(bm) - 1;
                        bm &= bm - 1;
                        visit((g0 + sl) * 32 + bl);
                    }
                }
            }
        } else {
            for (int g0 = 0; g0 < nblk; g0 += 32) {
                const int blk = g0 + lane;
                const bool bn = blk < nblk && box_bound(sbox[blk], sbox[nblk + blk], qx, qy, qz) <= r2;
                unsigned bm = __ballot_sync(0xffffffffu, bn);
                while (bm) {
                    const int bl = __ffs(bm) - 1;
                    bm &= bm - 1;
                    visit(g0 + bl);
                }
            }
        }
        const unsigned first = __shfl_sync(0xffffffffu, list_i, 0);
        if (lane < nsample) {
            const int v = total == 0 ? 0 : (int)(lane < total ? list_i : first);
            idx[((size_t)b * S + qi) * nsample + lane] = v;
        }
        if (cnt != nullptr && lane == 0) cnt[(size_t)b * S + qi] = total;
    }
}

inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

#ifdef SSF_CV_TRACE
extern "C" int ssf_knn_stat_read(unsigned long long* out, int reset) {
    if (cudaMemcpyFromSymbol(out, g_knn_stat, sizeof(unsigned long long) * 8) != cudaSuccess) return 2;
    if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_knn_stat, z, sizeof(z)); }
    return 0;
}
#endif

// floats of workspace for B reference clouds of Nr points
constexpr int KB_ONE_LEVEL_MAX = 16384;    // one-level search / shared-memory bitonic build up to here
constexpr int KB_MAX_REF = 131072;         // two-level search: <= 128 super-blocks of 1024 points

static long long knn_ws_per_cloud(int Nr) {
    const long long npad = ((long long)Nr + 31) / 32 * 32, nblk = npad / 32;
    if (Nr <= KB_ONE_LEVEL_MAX) return npad * 4 + nblk * 8;
    const long long nsb = (nblk + 31) / 32;
    return npad * 4 + nblk * 8 + nsb * 8 + npad * 4;   // + super-block boxes + radix-sort scratch (2 x (key, index))
}

extern "C" long long ssf_knn_blocks_workspace_floats(int B, int Nr) { return (long long)B * knn_ws_per_cloud(Nr); }

extern "C" int ssf_knn_blocks_build(const float* ref, int B, int Nr, float* ws, void* stream) {
    if (B <= 0 || Nr <= 0) return ssf_arg_error("knn_blocks_build: empty input");
    if (Nr > KB_MAX_REF) return ssf_arg_error("knn_blocks_build: at most 131072 reference points");
    if (Nr > KB_ONE_LEVEL_MAX) {
        const int npad_l = (Nr + 31) / 32 * 32, nblk_l = npad_l / 32, nsb = (nblk_l + 31) / 32;
        static unsigned long long attr_l = 0;
        if (ssf_attr_needed(&attr_l)) {
            cudaError_t e = cudaFuncSetAttribute(knn_blocks_build_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * KB_RADIX_T * 4);
            if (e != cudaSuccess) return ssf_set_error(e);
            ssf_attr_done(&attr_l);
        }
        knn_blocks_build_large_kernel<<<B, KB_RADIX_T, 16 * KB_RADIX_T * 4, (cudaStream_t)stream>>>(ref, Nr, npad_l, nblk_l, nsb, knn_ws_per_cloud(Nr), ws);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    const int npow2 = next_pow2(Nr) < 256 ? 256 : next_pow2(Nr), npad = (Nr + 31) / 32 * 32, nblk = npad / 32;   // the CTA sort works on runs of 256
    const size_t smem = (size_t)npow2 * 8;
    static unsigned long long attr_set = 0;
    if (ssf_attr_needed(&attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(knn_blocks_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
        if (e != cudaSuccess) return ssf_set_error(e);
        ssf_attr_done(&attr_set);
    }
    knn_blocks_build_kernel<<<B, KB_BUILD_T, smem, (cudaStream_t)stream>>>(ref, Nr, npow2, npad, nblk, ws);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

extern "C" int ssf_knn_blocks_search(int k, const float* query, const float* query_add, const float* ws, int B, int Nq, int Nr,
                                     float* dist, int* idx, void* stream) {
    if (B <= 0 || Nq <= 0) return ssf_arg_error("knn_blocks_search: empty input");
    if (k <= 0 || k > 32) return ssf_arg_error("knn: k must be in [1,32]");
    if (k > Nr) return ssf_arg_error("knn: k exceeds the number of reference points");
    if (Nr > KB_MAX_REF) return ssf_arg_error("knn_blocks_search: at most 131072 reference points");
    const int npad = (Nr + 31) / 32 * 32, nblk = npad / 32;
    dim3 grid((Nq + KB_QPB - 1) / KB_QPB, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (Nr > KB_ONE_LEVEL_MAX) {
        const int nsb = (nblk + 31) / 32;
        const size_t smem2 = (size_t)nsb * 32;
        if (nsb <= 64)
            knn_blocks_search2_kernel<2><<<grid, KB_SEARCH_T, smem2, st>>>(k, query, query_add, ws, knn_ws_per_cloud(Nr), Nq, npad, nblk, nsb, dist, idx);
        else
            knn_blocks_search2_kernel<4><<<grid, KB_SEARCH_T, smem2, st>>>(k, query, query_add, ws, knn_ws_per_cloud(Nr), Nq, npad, nblk, nsb, dist, idx);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    const size_t smem = (size_t)nblk * 32;
    static int use_sb = -1;   // SSF_KNN_SB=0: the register-resident bounds for every size (measurement switch)
    if (use_sb < 0) {
        const char* e = getenv("SSF_KNN_SB");
        use_sb = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (use_sb && nblk > 64) {
        const size_t smem_sb = smem + 64 * 16;
        if (nblk <= 128)
            knn_blocks_search_sb_kernel<4><<<grid, KB_SEARCH_T, smem_sb, st>>>(k, query, query_add, ws, Nq, npad, nblk, dist, idx);
        else if (nblk <= 256)
            knn_blocks_search_sb_kernel<8><<<grid, KB_SEARCH_T, smem_sb, st>>>(k, query, query_add, ws, Nq, npad, nblk, dist, idx);
        else
            knn_blocks_search_sb_kernel<16><<<grid, KB_SEARCH_T, smem_sb, st>>>(k, query, query_add, ws, Nq, npad, nblk, dist, idx);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    if (nblk <= 64)
        knn_blocks_search_kernel<2><<<grid, KB_SEARCH_T, smem, st>>>(k, query, query_add, ws, Nq, npad, nblk, dist, idx);
    else if (nblk <= 256)
        knn_blocks_search_kernel<8><<<grid, KB_SEARCH_T, smem, st>>>(k, query, query_add, ws, Nq, npad, nblk, dist, idx);
    else
        knn_blocks_search_kernel<16><<<grid, KB_SEARCH_T, smem, st>>>(k, query, query_add, ws, Nq, npad, nblk, dist, idx);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// brute force with one warp per two queries: any Nr, meant for few queries (B * Nq <= 32768) against a large cloud
extern "C" int ssf_knn_warp_scan(int k, const float* query, const float* query_add, const float* ref, int B, int Nq, int Nr,
                                 float* dist, int* idx, void* stream) {
    if (B <= 0 || Nq <= 0) return ssf_arg_error("knn_warp_scan: empty input");
    if (k <= 0 || k > 32) return ssf_arg_error("knn: k must be in [1,32]");
    if (k > Nr) return ssf_arg_error("knn: k exceeds the number of reference points");
    dim3 grid((Nq + 8 * KS_Q - 1) / (8 * KS_Q), B);
    knn_warp_scan_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(k, query, query_add, ref, Nq, Nr, dist, idx);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// ball_query through the index built by ssf_knn_blocks_build for the same cloud (N points): same output as ssf_ball_query.
// nsample <= 32 (the hit list is lane-distributed).
extern "C" int ssf_ball_query_blocks(float radius, int nsample, const float* new_xyz, const float* ws, int B, int N, int S, int* idx,
                                     int* cnt, void* stream) {
    if (B <= 0 || S <= 0 || N <= 0) return ssf_arg_error("ball_query_blocks: empty input");
    if (nsample <= 0 || nsample > 32) return ssf_arg_error("ball_query_blocks: nsample must be in [1,32]");
    if (N > KB_MAX_REF) return ssf_arg_error("ball_query_blocks: at most 131072 points");
    const float r2 = radius * radius;  // fp32 product, as the spec
    const int npad = (N + 31) / 32 * 32, nblk = npad / 32, nsb = (nblk + 31) / 32;
    dim3 grid((S + KB_QPB - 1) / KB_QPB, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (N > KB_ONE_LEVEL_MAX)
        ball_query_blocks_kernel<true><<<grid, KB_SEARCH_T, (size_t)nsb * 32, st>>>(r2, nsample, new_xyz, ws, knn_ws_per_cloud(N), S, npad, nblk, nsb, idx, cnt);
    else
        ball_query_blocks_kernel<false><<<grid, KB_SEARCH_T, (size_t)nblk * 32, st>>>(r2, nsample, new_xyz, ws, knn_ws_per_cloud(N), S, npad, nblk, nsb, idx, cnt);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
