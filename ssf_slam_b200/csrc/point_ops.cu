// Point operators of the SSF-SLAM scene-flow front end, hand-written for sm_100a.
//
// Drop-in replacements for the reference's absent pointnet2 extension (`lib.pointnet2_utils`,
// call sites ASF/utils/utils.py:226-233,291-302 and ASF/utils/soflow.py:30,387-406,1241-1249,
// 1459-1470) and for torch_scatter (ASF/utils/soflow.py:474,481).  Arithmetic and tie-breaking follow
// the written spec: no-FMA squared distances, (distance, index) lexicographic
// kNN order, FPS from index 0 with lowest-index argmax.
//
//  * furthest_point_sample : one CTA per cloud, points and running minima in registers, argmax by two
//    redux.sync per warp + one shared-memory hop; clouds larger than one CTA's registers run on a
//    thread-block cluster and exchange candidates through distributed shared memory.
//  * knn / three_nn / ball_query : thread per query, reference cloud streamed through shared memory in
//    double-buffered tiles by the TMA engine (cp.async.bulk + mbarrier), float4 shared loads.
//  * grouping / gather / three_interpolate : vectorised gathers with streaming stores.
//  * scatter softmax / sum : deterministic (atomic-free accumulation) through a CSR of the key lists.
#include <cooperative_groups.h>
#include <cstdlib>

#include "ssf_common.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------ FPS

// One CTA per cloud.  Thread t owns points t, t+T, ...; xyz also sits in shared memory so that the
// coordinates of the winner can be broadcast without a global round trip.
template <int T, int PPT>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float* __restrict__ xyz, int N, int npoint, int* __restrict__ out) {
    extern __shared__ float s_xyz[];
    __shared__ unsigned s_val[2][32];
    __shared__ unsigned s_idx[2][32];
    constexpr int NW = T / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* p = xyz + (size_t)blockIdx.x * N * 3;
    int* o = out + (size_t)blockIdx.x * npoint;
    for (int i = tid; i < N * 3; i += T) s_xyz[i] = p[i];
    __syncthreads();

    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        int i = tid + j * T;
        bool ok = i < N;
        px[j] = ok ? s_xyz[3 * i] : 0.f;
        py[j] = ok ? s_xyz[3 * i + 1] : 0.f;
        pz[j] = ok ? s_xyz[3 * i + 2] : 0.f;
        md[j] = 1e10f;
    }
    int last = 0;
    for (int it = 0; it < npoint; ++it) {
        if (tid == 0) o[it] = last;
        if (it == npoint - 1) break;
        const float lx = s_xyz[3 * last], ly = s_xyz[3 * last + 1], lz = s_xyz[3 * last + 2];
        float best = -1.f;
        unsigned besti = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            int i = tid + j * T;
            if (i < N) {
                float d = ssf_sqdist(px[j], py[j], pz[j], lx, ly, lz);
                float m = fminf(md[j], d);
                md[j] = m;
                if (m > best) {
                    best = m;
                    besti = (unsigned)i;
                }
            }
        }
        // non-negative floats order like their bit patterns; ties resolve to the lowest index
        unsigned vb = besti == 0xffffffffu ? 0u : __float_as_uint(best);
        unsigned wm = __reduce_max_sync(0xffffffffu, vb);
        unsigned wi = __reduce_min_sync(0xffffffffu, vb == wm ? besti : 0xffffffffu);
        const int buf = it & 1;
        if (lane == 0) {
            s_val[buf][warp] = wm;
            s_idx[buf][warp] = wi;
        }
        __syncthreads();
        unsigned v2 = lane < NW ? s_val[buf][lane] : 0u;
        unsigned i2 = lane < NW ? s_idx[buf][lane] : 0xffffffffu;
        unsigned gm = __reduce_max_sync(0xffffffffu, v2);
        last = (int)__reduce_min_sync(0xffffffffu, v2 == gm ? i2 : 0xffffffffu);
    }
}

// Pruned variant (2048 < N <= 8192), same result bit for bit.  The distance update of an iteration can only lower the running
// minimum of points that are closer to the new sample than their current minimum, and after the first few dozen samples that is
// a small neighbourhood.  So the cloud is first sorted along a Hilbert curve (ssf_cta_sort_u64 in shared memory, as the kNN index
// build) and kept in shared memory as sorted x / y / z / running-minimum arrays.  32 consecutive sorted points -- a compact
// blob -- form a "row"; there is one thread per row: sorted row r belongs to warp r % NW (neighbouring rows, which are touched
// together, go to different warps) as its row j = r / NW, and lane j of that warp keeps the row's bounding box, its largest
// running minimum and the lowest original index attaining it.  Per iteration lane j bounds the distance from the new sample to
// box j with the distance's own rounded operations (monotone, so bound <= the computed distance of every point inside); a row
// is touched only if the bound is below its largest running minimum -- otherwise no minimum in it can change -- and only
// touched rows (about ten of 256 at N = 8192) update their points and recompute their maximum, up to four rows of a warp at a
// time so that their shared-memory and warp-reduction latencies overlap.  The iteration is a latency chain, not a throughput
// problem: few warps (one per 32 rows) keep the issue slots free for it.  The arg-max over rows, warps and the CTA is the same
// (value, lowest original index) reduction as in fps_kernel; the index travels as index << 13 | sorted slot so that the winner's
// coordinates are one shared-memory read away.
template <int U, int NW>
__device__ __forceinline__ void fps_rows(unsigned& touch, int warp, int lane, int N, const float* sx, const float* sy,
                                         const float* sz, float* smd, const int* sval, float lx, float ly, float lz,
                                         float& submax, unsigned& subidx) {
    int j[U], pos[U];
    unsigned id[U], rm[U], ri[U];
    float m[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        j[u] = touch ? __ffs(touch) - 1 : j[0];     // fewer than U rows left: row j[0] again (idempotent)
        touch &= touch - 1;
        pos[u] = (j[u] * NW + warp) * 32 + lane;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        id[u] = (unsigned)sval[pos[u]];
        const float d = ssf_sqdist(sx[pos[u]], sy[pos[u]], sz[pos[u]], lx, ly, lz);
        m[u] = fminf(smd[pos[u]], d);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) smd[pos[u]] = m[u];
    // non-negative floats order like their bit patterns; ties resolve to the lowest original index
#pragma unroll
    for (int u = 0; u < U; ++u) rm[u] = __reduce_max_sync(0xffffffffu, id[u] < (unsigned)N ? __float_as_uint(m[u]) : 0u);
#pragma unroll
    for (int u = 0; u < U; ++u)
        ri[u] = __reduce_min_sync(0xffffffffu, (id[u] < (unsigned)N && __float_as_uint(m[u]) == rm[u]) ? (id[u] << 13 | (unsigned)pos[u]) : 0xffffffffu);
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (lane == j[u]) {
            submax = __uint_as_float(rm[u]);
            subidx = ri[u];
        }
}

// T threads (one per row) run the sampling loop; the set-up (keys, sort, scatter) runs on TS = 4 T threads which then leave.
template <int T>
__global__ void __launch_bounds__(4 * T, 1)
fps_pruned_kernel(const float* __restrict__ xyz, int N, int npoint, int* __restrict__ out) {
    constexpr int NW = T / 32, NP = T * 32, TS = 4 * T, NWS = TS / 32;
    static_assert(NP <= 8192, "index << 13 | slot packing");
    extern __shared__ __align__(16) float s_dyn[];
    float* sx = s_dyn;                                   // [NP] sorted coordinates
    float* sy = sx + NP;
    float* sz = sy + NP;
    float* smd = sz + NP;                                // [NP] running minima (the curve keys during the sort)
    int* sval = reinterpret_cast<int*>(smd + NP);        // [NP] original index of sorted slot (>= N: padding)
    unsigned long long* skv = reinterpret_cast<unsigned long long*>(smd);   // [NP] key << 32 | index during the sort: smd + sval
    __shared__ unsigned long long s_cand[2][NW];        // per warp: value bits << 32 | index << 13 | slot
    __shared__ float sred[6][NWS];
    __shared__ float sbb[6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* p = xyz + (size_t)blockIdx.x * N * 3;
    int* o = out + (size_t)blockIdx.x * npoint;
    // cloud bounding box -> 10-bit cells, Hilbert keys
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < N; i += TS)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = p[3 * i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], of));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], of));
        }
        if (lane == 0) {
            sred[c][warp] = mn[c];
            sred[3 + c][warp] = mx[c];
        }
    }
    __syncthreads();
    if (tid < 6) {
        float v = sred[tid][0];
        for (int w = 1; w < NWS; ++w) v = tid < 3 ? fminf(v, sred[tid][w]) : fmaxf(v, sred[tid][w]);
        sbb[tid] = v;
    }
    __syncthreads();
    {
        // one scale for the three axes: cells are cubes, so rows are compact in the metric the bound uses
        const float ext = fmaxf(sbb[3] - sbb[0], fmaxf(sbb[4] - sbb[1], sbb[5] - sbb[2]));
        const float sc1 = ext > 0.f ? 1023.0f / ext : 0.f;
        const float sc[3] = {sc1, sc1, sc1};
        for (int i = tid; i < NP; i += TS) {
            unsigned key = 0xFFFFFFFFu;
            if (i < N) {
                unsigned q[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float t = (p[3 * i + c] - sbb[c]) * sc[c];
                    q[c] = (unsigned)fminf(fmaxf(t, 0.f), 1023.f);
                }
                key = ssf_hilbert30(q[0], q[1], q[2]);
            }
            skv[i] = (unsigned long long)key << 32 | (unsigned)i;
        }
    }
    __syncthreads();
    ssf_cta_sort_u64(skv, NP, tid, TS);
    __shared__ int s_first;                              // sorted slot of original point 0, the first sample
    // sorted coordinates; then the (key, index) pairs make room for the running minima and the indices
    for (int i = tid; i < NP; i += TS) {
        const int id = (int)(unsigned)skv[i];
        const bool ok = id < N;
        sx[i] = ok ? p[3 * id] : 0.f;
        sy[i] = ok ? p[3 * id + 1] : 0.f;
        sz[i] = ok ? p[3 * id + 2] : 0.f;
        if (id == 0) s_first = i;
    }
    {
        int ids[(NP + TS - 1) / TS];
#pragma unroll
        for (int u = 0; u < (NP + TS - 1) / TS; ++u) ids[u] = (int)(unsigned)skv[tid + u * TS];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < (NP + TS - 1) / TS; ++u) {
            smd[tid + u * TS] = 1e10f;
            sval[tid + u * TS] = ids[u];
        }
    }
    __syncthreads();
    if (tid >= T) return;
    auto loop_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory"); };
    // lane j: box of row j of this warp, its largest running minimum and the lowest index attaining it
    float blx = INFINITY, bly = INFINITY, blz = INFINITY, bhx = -INFINITY, bhy = -INFINITY, bhz = -INFINITY;
    float submax = 0.f;
    unsigned subidx = 0xffffffffu;
    for (int j = 0; j < 32; ++j) {
        const int pos = (j * NW + warp) * 32 + lane;
        const bool ok = sval[pos] < N;
        float l0 = ok ? sx[pos] : INFINITY, l1 = ok ? sy[pos] : INFINITY, l2 = ok ? sz[pos] : INFINITY;
        float h0 = ok ? sx[pos] : -INFINITY, h1 = ok ? sy[pos] : -INFINITY, h2 = ok ? sz[pos] : -INFINITY;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) {
            l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, of));
            l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, of));
            l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, of));
            h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, of));
            h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, of));
            h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, of));
        }
        const unsigned any = __ballot_sync(0xffffffffu, ok);
        if (lane == j) {
            blx = l0; bly = l1; blz = l2; bhx = h0; bhy = h1; bhz = h2;
            submax = any ? 1e10f : 0.f;      // an empty row never wins and is never touched
        }
    }
    unsigned last = (unsigned)s_first;       // original index << 13 | sorted slot
    unsigned wm = 0u, wi = 0xffffffffu;      // this warp's candidate; changes only when one of its rows is touched
#ifdef SSF_CV_TRACE
    long long tA = 0, tB = 0, tC = 0, tD = 0, tE = 0, nrows = 0;
#define FPS_TR(x) { long long c_ = clock64(); x += c_ - tr0; tr0 = c_; }
#else
#define FPS_TR(x)
#endif
    for (int it = 0; it < npoint; ++it) {
        if (tid == 0) o[it] = (int)(last >> 13);
        if (it == npoint - 1) break;
#ifdef SSF_CV_TRACE
        long long tr0 = clock64();
#endif
        const int lp = (int)(last & 8191u);
        const float lx = sx[lp], ly = sy[lp], lz = sz[lp];
        // lower bound of the distance from the new sample to every point of row `lane` (same rounded operations as ssf_sqdist)
        const float gx = fmaxf(0.f, fmaxf(__fsub_rn(blx, lx), __fsub_rn(lx, bhx)));
        const float gy = fmaxf(0.f, fmaxf(__fsub_rn(bly, ly), __fsub_rn(ly, bhy)));
        const float gz = fmaxf(0.f, fmaxf(__fsub_rn(blz, lz), __fsub_rn(lz, bhz)));
        const float lb = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
        unsigned touch = __ballot_sync(0xffffffffu, lb < submax);
        FPS_TR(tA)
        if (touch) {   // warp-uniform
#ifdef SSF_CV_TRACE
            nrows += __popc(touch);
#endif
            if ((touch & (touch - 1)) == 0) {
                fps_rows<1, NW>(touch, warp, lane, N, sx, sy, sz, smd, sval, lx, ly, lz, submax, subidx);
            } else if (__popc(touch) == 2) {
                fps_rows<2, NW>(touch, warp, lane, N, sx, sy, sz, smd, sval, lx, ly, lz, submax, subidx);
            } else {
                do {
                    fps_rows<4, NW>(touch, warp, lane, N, sx, sy, sz, smd, sval, lx, ly, lz, submax, subidx);
                } while (touch);
            }
            FPS_TR(tB)
            const unsigned vb = subidx != 0xffffffffu ? __float_as_uint(submax) : 0u;
            wm = __reduce_max_sync(0xffffffffu, vb);
            wi = __reduce_min_sync(0xffffffffu, vb == wm ? subidx : 0xffffffffu);
            FPS_TR(tC)
        }
        const int buf = it & 1;
        if (lane == 0) s_cand[buf][warp] = (unsigned long long)wm << 32 | wi;
        loop_sync();
        const unsigned long long c = lane < NW ? s_cand[buf][lane] : 0xffffffffull;
        FPS_TR(tD)
        const unsigned gm = __reduce_max_sync(0xffffffffu, (unsigned)(c >> 32));
        last = __reduce_min_sync(0xffffffffu, (unsigned)(c >> 32) == gm ? (unsigned)c : 0xffffffffu);
        FPS_TR(tE)
    }
#ifdef SSF_CV_TRACE
    if (blockIdx.x == 0 && lane == 0 && (warp < 4 || warp == NW - 1))
        printf("fps trace N=%d T=%d warp %d: per-iteration cycles  bound+ballot %.0f  rows %.0f (%.2f rows)  warp-redux %.0f  sts+bar+lds %.0f  block-redux %.0f\n",
               N, T, warp, (double)tA / npoint, (double)tB / npoint, (double)nrows / npoint, (double)tC / npoint, (double)tD / npoint, (double)tE / npoint);
#endif
#undef FPS_TR
}

template <int T>
static int launch_fps_pruned(const float* xyz, int B, int N, int npoint, int* out, cudaStream_t st) {
    const size_t smem = (size_t)T * 32 * 20;
    cudaError_t e = cudaFuncSetAttribute(fps_pruned_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ssf_set_error(e);
    fps_pruned_kernel<T><<<B, 4 * T, smem, st>>>(xyz, N, npoint, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// Cluster variant: CS CTAs of 1024 threads share one cloud (N <= CS * 1024 * PPT); each CTA keeps its slice
// in registers and publishes its candidate (value, index, xyz) into every peer's shared memory (DSMEM),
// one cluster barrier per iteration.
struct FpsCand {
    unsigned val;
    unsigned idx;
    float x, y, z;
    float pad[3];
};

template <int PPT>
__global__ void __launch_bounds__(1024, 1)
fps_cluster_kernel(const float* __restrict__ xyz, int N, int npoint, int* __restrict__ out) {
    constexpr int T = 1024;
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / cs;
    __shared__ unsigned s_val[32];
    __shared__ unsigned s_idx[32];
    __shared__ FpsCand s_cand[2][16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* p = xyz + (size_t)cloud * N * 3;
    int* o = out + (size_t)cloud * npoint;
    const int slice = (N + cs - 1) / cs;
    const int lo = rank * slice;
    const int hi = min(N, lo + slice);

    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        int i = lo + tid + j * T;
        bool ok = i < hi;
        px[j] = ok ? p[3 * i] : 0.f;
        py[j] = ok ? p[3 * i + 1] : 0.f;
        pz[j] = ok ? p[3 * i + 2] : 0.f;
        md[j] = 1e10f;
    }
    int last = 0;
    float lx = p[0], ly = p[1], lz = p[2];
    cluster.sync();
    for (int it = 0; it < npoint; ++it) {
        if (rank == 0 && tid == 0) o[it] = last;
        if (it == npoint - 1) break;
        float best = -1.f;
        unsigned besti = 0xffffffffu;
        int bestj = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            int i = lo + tid + j * T;
            if (i < hi) {
                float d = ssf_sqdist(px[j], py[j], pz[j], lx, ly, lz);
                float m = fminf(md[j], d);
                md[j] = m;
                if (m > best) {
                    best = m;
                    besti = (unsigned)i;
                    bestj = j;
                }
            }
        }
        unsigned vb = besti == 0xffffffffu ? 0u : __float_as_uint(best);
        unsigned wm = __reduce_max_sync(0xffffffffu, vb);
        unsigned wi = __reduce_min_sync(0xffffffffu, vb == wm ? besti : 0xffffffffu);
        if (lane == 0) {
            s_val[warp] = wm;
            s_idx[warp] = wi;
        }
        __syncthreads();
        unsigned v2 = s_val[lane];
        unsigned i2 = s_idx[lane];
        unsigned gm = __reduce_max_sync(0xffffffffu, v2);
        unsigned gi = __reduce_min_sync(0xffffffffu, v2 == gm ? i2 : 0xffffffffu);
        const int buf = it & 1;
        // the owner of the CTA-level winner publishes it to every CTA of the cluster
        if (gi != 0xffffffffu && besti == gi) {
            float wx = 0.f, wy = 0.f, wz = 0.f;
#pragma unroll
            for (int j = 0; j < PPT; ++j)
                if (j == bestj) {
                    wx = px[j];
                    wy = py[j];
                    wz = pz[j];
                }
            for (int r = 0; r < cs; ++r) {
                FpsCand* dst = cluster.map_shared_rank(&s_cand[buf][rank], r);
                dst->val = gm;
                dst->idx = gi;
                dst->x = wx;
                dst->y = wy;
                dst->z = wz;
            }
        } else if (gi == 0xffffffffu && tid == 0) {  // empty slice
            for (int r = 0; r < cs; ++r) {
                FpsCand* dst = cluster.map_shared_rank(&s_cand[buf][rank], r);
                dst->val = 0u;
                dst->idx = 0xffffffffu;
            }
        }
        cluster.sync();
        unsigned cv = lane < cs ? s_cand[buf][lane].val : 0u;
        unsigned ci = lane < cs ? s_cand[buf][lane].idx : 0xffffffffu;
        unsigned cm = __reduce_max_sync(0xffffffffu, cv);
        unsigned win = __reduce_min_sync(0xffffffffu, cv == cm ? ci : 0xffffffffu);
        int src = __ffs(__ballot_sync(0xffffffffu, lane < cs && cv == cm && ci == win)) - 1;
        last = (int)win;
        lx = s_cand[buf][src].x;
        ly = s_cand[buf][src].y;
        lz = s_cand[buf][src].z;
    }
    cluster.sync();  // no CTA may exit while peers can still write into its shared memory
}

template <int T, int PPT>
static int launch_fps(const float* xyz, int B, int N, int npoint, int* out, cudaStream_t st) {
    size_t smem = (size_t)N * 3 * sizeof(float);
    if (smem > 40 * 1024) {   // the 48 KB default covers static + dynamic shared memory: opt in with room for the static part
        cudaError_t e = cudaFuncSetAttribute(fps_kernel<T, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return ssf_set_error(e);
    }
    fps_kernel<T, PPT><<<B, T, smem, st>>>(xyz, N, npoint, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

template <int PPT>
static int launch_fps_cluster(const float* xyz, int B, int N, int npoint, int* out, int cs, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * cs);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (cs > 8) {
        cudaError_t e = cudaFuncSetAttribute(fps_cluster_kernel<PPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return ssf_set_error(e);
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, fps_cluster_kernel<PPT>, xyz, N, npoint, out);
    ssf_count_launch();
    if (e != cudaSuccess) return ssf_set_error(e);
    return SSF_OK;
}

extern "C" int ssf_furthest_point_sample(const float* xyz, int B, int N, int npoint, int* idx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || npoint <= 0) return ssf_arg_error("furthest_point_sample: empty input");
    static int prune = -1;   // SSF_FPS_PRUNE=0: the plain kernels everywhere (measurement / regression switch)
    if (prune < 0) {
        const char* e = getenv("SSF_FPS_PRUNE");
        prune = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (prune && N > 2048 && N <= 8192) {
        if (N <= 4096) return launch_fps_pruned<128>(xyz, B, N, npoint, idx, st);
        return launch_fps_pruned<256>(xyz, B, N, npoint, idx, st);
    }
    if (N <= 128) return launch_fps<128, 1>(xyz, B, N, npoint, idx, st);
    if (N <= 256) return launch_fps<128, 2>(xyz, B, N, npoint, idx, st);
    if (N <= 512) return launch_fps<256, 2>(xyz, B, N, npoint, idx, st);
    if (N <= 1024) return launch_fps<256, 4>(xyz, B, N, npoint, idx, st);
    if (N <= 2048) return launch_fps<512, 4>(xyz, B, N, npoint, idx, st);
    if (N <= 4096) return launch_fps<1024, 4>(xyz, B, N, npoint, idx, st);
    if (N <= 8192) return launch_fps<1024, 8>(xyz, B, N, npoint, idx, st);
    if (N <= 16384) return launch_fps_cluster<8>(xyz, B, N, npoint, idx, 2, st);
    if (N <= 32768) return launch_fps_cluster<8>(xyz, B, N, npoint, idx, 4, st);
    if (N <= 65536) return launch_fps_cluster<8>(xyz, B, N, npoint, idx, 8, st);
    if (N <= 131072) return launch_fps_cluster<8>(xyz, B, N, npoint, idx, 16, st);
    return ssf_arg_error("furthest_point_sample: N > 131072 not supported");
}

// ------------------------------------------------------------------------- tiled reference scan

constexpr int SCAN_T = 128;      // threads (= queries) per CTA
constexpr int SCAN_TILE = 1024;  // reference points per shared-memory tile (12 KB), double buffered

// Streams ref[b] through shared memory and calls op(x, y, z, global_index) for every reference point, in
// ascending index order.  The bulk-copy (TMA) path needs 16-byte aligned sources and sizes; ragged tiles
// fall back to cooperative loads.
template <class Op>
__device__ __forceinline__ void scan_reference(const float* __restrict__ refb, int Nr, float* tiles, uint64_t* bars, Op& op) {
    const int tid = threadIdx.x;
    const int ntiles = (Nr + SCAN_TILE - 1) / SCAN_TILE;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(refb) & 15) == 0;
    auto issue = [&](int t) {
        const int cnt = min(SCAN_TILE, Nr - t * SCAN_TILE);
        float* dst = tiles + (t & 1) * SCAN_TILE * 3;
        const float* src = refb + (size_t)t * SCAN_TILE * 3;
        const bool bulk = base_aligned && ((cnt * 12) % 16 == 0);
        if (bulk) {
            if (tid == 0) {
                ssf_mbar_expect_tx(&bars[t & 1], (uint32_t)cnt * 12u);
                ssf_bulk_g2s(dst, src, (uint32_t)cnt * 12u, &bars[t & 1]);
            }
        } else {
            for (int i = tid; i < cnt * 3; i += SCAN_T) dst[i] = src[i];
        }
        return bulk;
    };
    // `bulk` is a block-uniform function of (t): recompute instead of storing
    auto is_bulk = [&](int t) {
        const int cnt = min(SCAN_TILE, Nr - t * SCAN_TILE);
        return base_aligned && ((cnt * 12) % 16 == 0);
    };
    unsigned phases = 0u;  // bit i = parity to wait for on bars[i]
    issue(0);
    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue(t + 1);
        const int cur = t & 1;
        if (is_bulk(t)) {
            ssf_mbar_wait(&bars[cur], (phases >> cur) & 1u);
            phases ^= 1u << cur;
        } else {
            __syncthreads();
        }
        const float* tp = tiles + cur * SCAN_TILE * 3;
        const int cnt = min(SCAN_TILE, Nr - t * SCAN_TILE);
        const int base = t * SCAN_TILE;
        const float4* t4 = reinterpret_cast<const float4*>(tp);
        const int n4 = cnt >> 2;
        for (int g = 0; g < n4; ++g) {
            const float4 a = t4[3 * g], b = t4[3 * g + 1], c = t4[3 * g + 2];
            const int r = base + 4 * g;
            op(a.x, a.y, a.z, r);
            op(a.w, b.x, b.y, r + 1);
            op(b.z, b.w, c.x, r + 2);
            op(c.y, c.z, c.w, r + 3);
        }
        for (int r = n4 << 2; r < cnt; ++r) op(tp[3 * r], tp[3 * r + 1], tp[3 * r + 2], base + r);
        __syncthreads();  // everyone is done with this buffer before it is refilled
    }
}

template <int K>
struct KnnOp {
    float qx, qy, qz;
    float bd[K];
    int bi[K];
    __device__ __forceinline__ void operator()(float x, float y, float z, int r) {
        const float d = ssf_sqdist(qx, qy, qz, x, y, z);
        // candidates arrive in ascending index: strict '<' keeps the lowest index on equal distance
        if (d < bd[K - 1]) {
            bd[K - 1] = d;
            bi[K - 1] = r;
#pragma unroll
            for (int j = K - 1; j > 0; --j) {
                if (bd[j] < bd[j - 1]) {
                    float td = bd[j];
                    bd[j] = bd[j - 1];
                    bd[j - 1] = td;
                    int ti = bi[j];
                    bi[j] = bi[j - 1];
                    bi[j - 1] = ti;
                }
            }
        }
    }
};

// query [B,Nq,3] (+ optional per-query offset qadd [B,Nq,3], added with one rounded fp32 add exactly as
// `xyz1 + sf` at ASF/utils/soflow.py:389), ref [B,Nr,3] -> dist [B,Nq,k] (may be null), idx [B,Nq,k]
template <int K>
__global__ void __launch_bounds__(SCAN_T)
knn_kernel(const float* __restrict__ query, const float* __restrict__ qadd, const float* __restrict__ ref, int Nq, int Nr,
           int k, float* __restrict__ dist, int* __restrict__ idx) {
    __shared__ __align__(128) float tiles[2 * SCAN_TILE * 3];
    __shared__ __align__(8) uint64_t bars[2];
    const int b = blockIdx.y;
    const int q = blockIdx.x * SCAN_T + threadIdx.x;
    if (threadIdx.x == 0) {
        ssf_mbar_init(&bars[0], 1);
        ssf_mbar_init(&bars[1], 1);
        ssf_mbar_fence_init();
    }
    __syncthreads();
    KnnOp<K> op;
    const int qq = min(q, Nq - 1);
    const float* qp = query + ((size_t)b * Nq + qq) * 3;
    op.qx = qp[0];
    op.qy = qp[1];
    op.qz = qp[2];
    if (qadd != nullptr) {
        const float* ap = qadd + ((size_t)b * Nq + qq) * 3;
        op.qx = __fadd_rn(op.qx, ap[0]);
        op.qy = __fadd_rn(op.qy, ap[1]);
        op.qz = __fadd_rn(op.qz, ap[2]);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
        op.bd[j] = __int_as_float(0x7f800000);
        op.bi[j] = 0;
    }
    scan_reference(ref + (size_t)b * Nr * 3, Nr, tiles, bars, op);
    if (q < Nq) {
        const size_t o = ((size_t)b * Nq + q) * k;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (j < k) {
                idx[o + j] = op.bi[j];
                if (dist != nullptr) dist[o + j] = __fsqrt_rn(op.bd[j]);
            }
        }
    }
}

extern "C" int ssf_knn_offset(int k, const float* query, const float* query_add, const float* ref, int B, int Nq, int Nr,
                              float* dist, int* idx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || Nq <= 0) return ssf_arg_error("knn: empty input");
    if (k <= 0 || k > 32) return ssf_arg_error("knn: k must be in [1,32]");
    if (k > Nr) return ssf_arg_error("knn: k exceeds the number of reference points");
    dim3 grid((Nq + SCAN_T - 1) / SCAN_T, B);
    if (k <= 3)
        knn_kernel<3><<<grid, SCAN_T, 0, st>>>(query, query_add, ref, Nq, Nr, k, dist, idx);
    else if (k <= 5)
        knn_kernel<5><<<grid, SCAN_T, 0, st>>>(query, query_add, ref, Nq, Nr, k, dist, idx);
    else if (k <= 8)
        knn_kernel<8><<<grid, SCAN_T, 0, st>>>(query, query_add, ref, Nq, Nr, k, dist, idx);
    else if (k <= 16)
        knn_kernel<16><<<grid, SCAN_T, 0, st>>>(query, query_add, ref, Nq, Nr, k, dist, idx);
    else
        knn_kernel<32><<<grid, SCAN_T, 0, st>>>(query, query_add, ref, Nq, Nr, k, dist, idx);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

extern "C" int ssf_knn(int k, const float* query, const float* ref, int B, int Nq, int Nr, float* dist, int* idx,
                       void* stream) {
    return ssf_knn_offset(k, query, nullptr, ref, B, Nq, Nr, dist, idx, stream);
}

extern "C" int ssf_three_nn(const float* query, const float* ref, int B, int Nq, int Nr, float* dist, int* idx,
                            void* stream) {
    return ssf_knn_offset(3, query, nullptr, ref, B, Nq, Nr, dist, idx, stream);
}

struct BallOp {
    float qx, qy, qz, r2;
    int nsample, cnt, first;
    int* out;
    __device__ __forceinline__ void operator()(float x, float y, float z, int r) {
        const float d = ssf_sqdist(qx, qy, qz, x, y, z);
        if (d <= r2) {
            if (cnt == 0) first = r;
            if (cnt < nsample && out != nullptr) out[cnt] = r;
            ++cnt;
        }
    }
};

// xyz [B,N,3], new_xyz [B,S,3] -> idx [B,S,nsample], cnt [B,S] (may be null)
__global__ void __launch_bounds__(SCAN_T)
ball_query_kernel(float r2, int nsample, const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S,
                  int* __restrict__ idx, int* __restrict__ cnt) {
    __shared__ __align__(128) float tiles[2 * SCAN_TILE * 3];
    __shared__ __align__(8) uint64_t bars[2];
    const int b = blockIdx.y;
    const int s = blockIdx.x * SCAN_T + threadIdx.x;
    if (threadIdx.x == 0) {
        ssf_mbar_init(&bars[0], 1);
        ssf_mbar_init(&bars[1], 1);
        ssf_mbar_fence_init();
    }
    __syncthreads();
    BallOp op;
    const int ss = min(s, S - 1);
    const float* c = new_xyz + ((size_t)b * S + ss) * 3;
    op.qx = c[0];
    op.qy = c[1];
    op.qz = c[2];
    op.r2 = r2;
    op.nsample = nsample;
    op.cnt = 0;
    op.first = 0;
    op.out = s < S ? idx + ((size_t)b * S + s) * nsample : nullptr;
    scan_reference(xyz + (size_t)b * N * 3, N, tiles, bars, op);
    if (s < S) {
        for (int j = min(op.cnt, nsample); j < nsample; ++j) op.out[j] = op.first;
        if (cnt != nullptr) cnt[(size_t)b * S + s] = op.cnt;
    }
}

// Few query points (B * S below one thread per lane of the machine, e.g. BASELINE config 4: 2048 centres in 65536 points):
// one warp per BQ_Q queries, lane = reference point.  A step tests 32 consecutive points; the ballot of the hits is already in
// index order, so a hit's output slot is the running count plus the number of hits in lower lanes.  Same result as the
// thread-per-query kernel (first nsample hits in ascending index, padded with the first hit, total count).
constexpr int BQ_Q = 2;

__global__ void __launch_bounds__(256)
ball_query_warp_kernel(float r2, int nsample, const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S,
                       int* __restrict__ idx, int* __restrict__ cnt) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = ((int)blockIdx.x * 8 + warp) * BQ_Q;
    if (s0 >= S) return;
    float qx[BQ_Q], qy[BQ_Q], qz[BQ_Q];
    int c[BQ_Q], first[BQ_Q];
#pragma unroll
    for (int q = 0; q < BQ_Q; ++q) {
        const float* cp = new_xyz + ((size_t)b * S + min(s0 + q, S - 1)) * 3;
        qx[q] = __ldg(cp); qy[q] = __ldg(cp + 1); qz[q] = __ldg(cp + 2);
        c[q] = 0;
        first[q] = 0;
    }
    const float* P = xyz + (size_t)b * N * 3;
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < N; base += 32) {
        const int r = base + lane;
        const bool in = r < N;
        const int rc = in ? r : N - 1;
        const float x = __ldg(P + 3 * rc), y = __ldg(P + 3 * rc + 1), z = __ldg(P + 3 * rc + 2);
        bool done = cnt == nullptr;
#pragma unroll
        for (int q = 0; q < BQ_Q; ++q) {
            const bool hit = in && ssf_sqdist(qx[q], qy[q], qz[q], x, y, z) <= r2;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                if (c[q] == 0) first[q] = base + __ffs(m) - 1;
                const int pos = c[q] + __popc(m & lt);
                if (hit && pos < nsample && s0 + q < S) idx[((size_t)b * S + s0 + q) * nsample + pos] = r;
                c[q] += __popc(m);
            }
            done = done && c[q] >= nsample;
        }
        if (done) break;   // counts not asked for: stop once every query of the warp has its nsample hits
    }
#pragma unroll
    for (int q = 0; q < BQ_Q; ++q) {
        if (s0 + q >= S) continue;
        for (int j = min(c[q], nsample) + lane; j < nsample; j += 32) idx[((size_t)b * S + s0 + q) * nsample + j] = first[q];
        if (cnt != nullptr && lane == 0) cnt[(size_t)b * S + s0 + q] = c[q];
    }
}

extern "C" int ssf_ball_query(float radius, int nsample, const float* xyz, const float* new_xyz, int B, int N, int S,
                              int* idx, int* cnt, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || S <= 0 || N <= 0 || nsample <= 0) return ssf_arg_error("ball_query: empty input");
    const float r2 = radius * radius;  // fp32 product, as the spec
    if ((long long)B * S <= 32768) {   // too few queries to fill the machine with one thread each
        dim3 wgrid((S + 8 * BQ_Q - 1) / (8 * BQ_Q), B);
        ball_query_warp_kernel<<<wgrid, 256, 0, st>>>(r2, nsample, xyz, new_xyz, N, S, idx, cnt);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    dim3 grid((S + SCAN_T - 1) / SCAN_T, B);
    ball_query_kernel<<<grid, SCAN_T, 0, st>>>(r2, nsample, xyz, new_xyz, N, S, idx, cnt);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// --------------------------------------------------------------------------------- gathers

constexpr int GROUP_CH = 8;  // channels handled per thread (independent gathers in flight)

// feat [B,C,N], idx [B,MS] -> out [B,C,MS]; 4 consecutive outputs per thread, streaming stores
__global__ void __launch_bounds__(256)
group_vec4_kernel(const float* __restrict__ feat, const int* __restrict__ idx, int C, int N, int MS4, float* __restrict__ out) {
    const int j4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (j4 >= MS4) return;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GROUP_CH;
    const int4 id = __ldg(reinterpret_cast<const int4*>(idx) + (size_t)b * MS4 + j4);
    const float* f = feat + ((size_t)b * C + c0) * N;
    float4* o = reinterpret_cast<float4*>(out) + ((size_t)b * C + c0) * MS4 + j4;
    float4 v[GROUP_CH];
#pragma unroll
    for (int c = 0; c < GROUP_CH; ++c) {
        if (c0 + c < C) {
            const float* fc = f + (size_t)c * N;
            v[c] = make_float4(__ldg(fc + id.x), __ldg(fc + id.y), __ldg(fc + id.z), __ldg(fc + id.w));
        }
    }
#pragma unroll
    for (int c = 0; c < GROUP_CH; ++c)
        if (c0 + c < C) __stcs(o + (size_t)c * MS4, v[c]);
}

__global__ void __launch_bounds__(256)
group_scalar_kernel(const float* __restrict__ feat, const int* __restrict__ idx, int C, int N, int MS, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= MS) return;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GROUP_CH;
    const int id = __ldg(idx + (size_t)b * MS + j);
    const float* f = feat + ((size_t)b * C + c0) * N;
    float* o = out + ((size_t)b * C + c0) * MS + j;
#pragma unroll
    for (int c = 0; c < GROUP_CH; ++c)
        if (c0 + c < C) o[(size_t)c * MS] = __ldg(f + (size_t)c * N + id);
}

// Shared-memory staged variants: the GC source rows feat[b, c0 .. c0+GC, :] (N floats each) are brought in once per
// CTA by the TMA engine (cp.async.bulk), the random 4-byte gathers then hit shared memory instead of L1/L2, and the
// output leaves as coalesced streaming float4 stores -- the kernel runs at the speed of its (compulsory) write stream.
constexpr int STAGE_GC = 2;
constexpr int STAGE_T = 512;

__global__ void __launch_bounds__(STAGE_T)
group_staged_kernel(const float* __restrict__ feat, const int* __restrict__ idx, int C, int N, int MS4, int chunk4,
                    float* __restrict__ out) {
    extern __shared__ __align__(128) float srow[];   // [STAGE_GC][N]
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.z, c0 = blockIdx.y * STAGE_GC;
    const int nc = min(STAGE_GC, C - c0);
    if (threadIdx.x == 0) {
        ssf_mbar_init(&bar, 1);
        ssf_mbar_fence_init();
        ssf_mbar_expect_tx(&bar, (uint32_t)nc * (uint32_t)N * 4u);
        for (int c = 0; c < nc; ++c) ssf_bulk_g2s(srow + (size_t)c * N, feat + ((size_t)b * C + c0 + c) * N, (uint32_t)N * 4u, &bar);
    }
    __syncthreads();
    ssf_mbar_wait(&bar, 0);
    const int j_end = min(MS4, (int)(blockIdx.x + 1) * chunk4);
    const int4* ip = reinterpret_cast<const int4*>(idx) + (size_t)b * MS4;
    float4* o = reinterpret_cast<float4*>(out) + ((size_t)b * C + c0) * MS4;
    for (int j4 = blockIdx.x * chunk4 + threadIdx.x; j4 < j_end; j4 += STAGE_T) {
        const int4 id = __ldg(ip + j4);
#pragma unroll
        for (int c = 0; c < STAGE_GC; ++c)
            if (c < nc) {
                const float* r = srow + (size_t)c * N;
                __stcs(o + (size_t)c * MS4 + j4, make_float4(r[id.x], r[id.y], r[id.z], r[id.w]));
            }
    }
}

template <int GC, int T>
__global__ void __launch_bounds__(T)
three_interpolate_staged_kernel(const float* __restrict__ feat, const int* __restrict__ idx, const float* __restrict__ w, int C,
                                int M, int N, int chunk, int vec4, float* __restrict__ out) {
    extern __shared__ __align__(128) float srow[];   // [GC][M]
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.z, c0 = blockIdx.y * GC;
    const int nc = min(GC, C - c0);
    if (threadIdx.x == 0) {
        ssf_mbar_init(&bar, 1);
        ssf_mbar_fence_init();
        ssf_mbar_expect_tx(&bar, (uint32_t)nc * (uint32_t)M * 4u);
        for (int c = 0; c < nc; ++c) ssf_bulk_g2s(srow + (size_t)c * M, feat + ((size_t)b * C + c0 + c) * M, (uint32_t)M * 4u, &bar);
    }
    __syncthreads();
    ssf_mbar_wait(&bar, 0);
    const int n_end = min(N, (int)(blockIdx.x + 1) * chunk);
    if (vec4) {
        // four consecutive query points per thread: their 12 indices and 12 weights are three 16-byte loads each, the four
        // results of a channel leave as one streaming 16-byte store
        const int4* ip4 = reinterpret_cast<const int4*>(idx + (size_t)b * N * 3);
        const float4* wp4 = reinterpret_cast<const float4*>(w + (size_t)b * N * 3);
        for (int n4 = (blockIdx.x * chunk) / 4 + threadIdx.x; n4 * 4 < n_end; n4 += T) {
            const int4 ia = __ldg(ip4 + 3 * n4), ib = __ldg(ip4 + 3 * n4 + 1), ic = __ldg(ip4 + 3 * n4 + 2);
            const float4 wa = __ldg(wp4 + 3 * n4), wb = __ldg(wp4 + 3 * n4 + 1), wc = __ldg(wp4 + 3 * n4 + 2);
#pragma unroll
            for (int c = 0; c < GC; ++c)
                if (c < nc) {
                    const float* r = srow + (size_t)c * M;
                    float4 o;
                    o.x = __fadd_rn(__fadd_rn(__fmul_rn(r[ia.x], wa.x), __fmul_rn(r[ia.y], wa.y)), __fmul_rn(r[ia.z], wa.z));
                    o.y = __fadd_rn(__fadd_rn(__fmul_rn(r[ia.w], wa.w), __fmul_rn(r[ib.x], wb.x)), __fmul_rn(r[ib.y], wb.y));
                    o.z = __fadd_rn(__fadd_rn(__fmul_rn(r[ib.z], wb.z), __fmul_rn(r[ib.w], wb.w)), __fmul_rn(r[ic.x], wc.x));
                    o.w = __fadd_rn(__fadd_rn(__fmul_rn(r[ic.y], wc.y), __fmul_rn(r[ic.z], wc.z)), __fmul_rn(r[ic.w], wc.w));
                    __stcs(reinterpret_cast<float4*>(out + ((size_t)b * C + c0 + c) * N) + n4, o);
                }
        }
        return;
    }
    for (int n = blockIdx.x * chunk + threadIdx.x; n < n_end; n += T) {
        const int* ip = idx + ((size_t)b * N + n) * 3;
        const float* wp = w + ((size_t)b * N + n) * 3;
        const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
        const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
#pragma unroll
        for (int c = 0; c < GC; ++c)
            if (c < nc) {
                const float* r = srow + (size_t)c * M;
                __stcs(out + ((size_t)b * C + c0 + c) * N + n,
                       __fadd_rn(__fadd_rn(__fmul_rn(r[i0], w0), __fmul_rn(r[i1], w1)), __fmul_rn(r[i2], w2)));
            }
    }
}

// rows of `len` floats can be staged when they are 16-byte granular / aligned and STAGE_GC of them fit in shared memory
static bool can_stage(const float* feat, int len) {
    return (len % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0) && ((size_t)len * 4 * STAGE_GC <= 200 * 1024);
}

extern "C" int ssf_grouping_operation(const float* feat, const int* idx, int B, int C, int N, int M, int S, float* out,
                                      void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long MS = (long long)M * S;
    if (B <= 0 || C <= 0 || MS <= 0) return ssf_arg_error("grouping_operation: empty input");
    const bool vec = (MS % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec && can_stage(feat, N) && MS >= 4 * N) {
        // enough outputs per staged row to amortise the staging (each CTA produces >= 4 outputs per staged element)
        const int MS4 = (int)(MS / 4);
        const size_t smem = (size_t)N * 4 * STAGE_GC;
        int chunk4 = MS4;                                  // float4 outputs per CTA: at least N (4N outputs per channel)
        while (chunk4 / 2 >= N && chunk4 % 2 == 0) chunk4 /= 2;
        static unsigned long long attr = 0;
        if (ssf_attr_needed(&attr)) {
            cudaError_t e = cudaFuncSetAttribute(group_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return ssf_set_error(e);
            ssf_attr_done(&attr);
        }
        dim3 grid((MS4 + chunk4 - 1) / chunk4, (C + STAGE_GC - 1) / STAGE_GC, B);
        group_staged_kernel<<<grid, STAGE_T, smem, st>>>(feat, idx, C, N, MS4, chunk4, out);
    } else if (vec) {
        const int MS4 = (int)(MS / 4);
        dim3 grid((MS4 + 255) / 256, (C + GROUP_CH - 1) / GROUP_CH, B);
        group_vec4_kernel<<<grid, 256, 0, st>>>(feat, idx, C, N, MS4, out);
    } else {
        dim3 grid((unsigned)((MS + 255) / 256), (C + GROUP_CH - 1) / GROUP_CH, B);
        group_scalar_kernel<<<grid, 256, 0, st>>>(feat, idx, C, N, (int)MS, out);
    }
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

extern "C" int ssf_gather_operation(const float* feat, const int* idx, int B, int C, int N, int M, float* out, void* stream) {
    return ssf_grouping_operation(feat, idx, B, C, N, M, 1, out, stream);
}

// feat [B,C,M], idx [B,N,3], weight [B,N,3] -> out [B,C,N] = (g0*w0 + g1*w1) + g2*w2, every op rounded
__global__ void __launch_bounds__(256)
three_interpolate_kernel(const float* __restrict__ feat, const int* __restrict__ idx, const float* __restrict__ w, int C,
                         int M, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GROUP_CH;
    const int* ip = idx + ((size_t)b * N + n) * 3;
    const float* wp = w + ((size_t)b * N + n) * 3;
    const int i0 = ip[0], i1 = ip[1], i2 = ip[2];
    const float w0 = wp[0], w1 = wp[1], w2 = wp[2];
#pragma unroll
    for (int c = 0; c < GROUP_CH; ++c) {
        if (c0 + c < C) {
            const float* f = feat + ((size_t)b * C + c0 + c) * M;
            float v = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(f + i0), w0), __fmul_rn(__ldg(f + i1), w1)),
                                __fmul_rn(__ldg(f + i2), w2));
            out[((size_t)b * C + c0 + c) * N + n] = v;
        }
    }
}

extern "C" int ssf_three_interpolate(const float* feat, const int* idx, const float* weight, int B, int C, int M, int N,
                                     float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || C <= 0 || N <= 0) return ssf_arg_error("three_interpolate: empty input");
    if (can_stage(feat, M) && N >= M) {
        // every CTA re-reads the 24 N bytes of indices and weights of its points, so it stages three channels when two such
        // CTAs still share an SM (one CTA per SM with 4-6 rows measured slower: its load phase is not overlapped)
        const int gc = (size_t)M * 4 * 3 <= 100 * 1024 ? 3 : 2;
        const size_t smem = (size_t)M * 4 * gc;
        int chunk = N;
        while (chunk / 2 >= M && chunk % 2 == 0) chunk /= 2;
        static unsigned long long attr = 0;
        if (ssf_attr_needed(&attr)) {
            cudaError_t e = cudaFuncSetAttribute(three_interpolate_staged_kernel<2, STAGE_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(three_interpolate_staged_kernel<3, STAGE_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return ssf_set_error(e);
            ssf_attr_done(&attr);
        }
        dim3 grid((N + chunk - 1) / chunk, (C + gc - 1) / gc, B);
        const auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        const int vec4 = (N % 4 == 0 && chunk % 4 == 0 && al16(idx) && al16(weight) && al16(out)) ? 1 : 0;
        if (gc == 3) three_interpolate_staged_kernel<3, STAGE_T><<<grid, STAGE_T, smem, st>>>(feat, idx, weight, C, M, N, chunk, vec4, out);
        else three_interpolate_staged_kernel<2, STAGE_T><<<grid, STAGE_T, smem, st>>>(feat, idx, weight, C, M, N, chunk, vec4, out);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    dim3 grid((N + 255) / 256, (C + GROUP_CH - 1) / GROUP_CH, B);
    three_interpolate_kernel<<<grid, 256, 0, st>>>(feat, idx, weight, C, M, N, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// --------------------------------------------------------- deterministic segmented softmax / sum

// CSR of "which rows l carry key j", rows of a segment in ascending l.  Integer-only, so the result is
// reproducible although the fill uses atomics (each segment is sorted afterwards).
template <typename IdxT>
__global__ void csr_count_kernel(const IdxT* __restrict__ key, int L, int n_seg, int* __restrict__ count) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (l >= L) return;
    const long long k = (long long)key[(size_t)b * L + l];
    if (k >= 0 && k < n_seg) atomicAdd(&count[(size_t)b * (n_seg + 1) + k], 1);
}

// in-place exclusive scan of count[b][0..n_seg] (n_seg+1 entries; the last becomes the total)
__global__ void __launch_bounds__(1024) csr_scan_kernel(int* __restrict__ count, int n_seg) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    int* c = count + (size_t)blockIdx.x * (n_seg + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_seg + 1; base += 1024) {
        const int i = base + tid;
        const int v = i < n_seg + 1 ? c[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int incl = x + (warp > 0 ? s_warp[warp - 1] : 0) + carry;
        if (i < n_seg + 1) c[i] = incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = incl;
        __syncthreads();
    }
}

template <typename IdxT>
__global__ void csr_fill_kernel(const IdxT* __restrict__ key, int L, int n_seg, const int* __restrict__ offset,
                                int* __restrict__ cursor, int* __restrict__ rows) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (l >= L) return;
    const long long k = (long long)key[(size_t)b * L + l];
    if (k < 0 || k >= n_seg) return;
    const int pos = atomicAdd(&cursor[(size_t)b * n_seg + k], 1);
    rows[(size_t)b * L + offset[(size_t)b * (n_seg + 1) + k] + pos] = l;
}

// rows of every segment into ascending order (the fill above lands them in atomic order).  One warp per segment:
// segments of <= 32 rows (the common case: 16 on average) are sorted across the lanes by ranking, longer ones
// by a warp-cooperative bitonic network in global memory.
__global__ void __launch_bounds__(256) csr_sort_kernel(const int* __restrict__ offset, int n_seg, int L, int* __restrict__ rows) {
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (j >= n_seg) return;
    const int* off = offset + (size_t)b * (n_seg + 1);
    int* r = rows + (size_t)b * L;
    const int lo = off[j], hi = off[j + 1], cnt = hi - lo;
    if (cnt <= 1) return;
    if (cnt <= 32) {
        // rank sort: the rows of a segment are distinct, so a row's place is the number of smaller rows.  The cnt broadcasts
        // are independent of each other (a bitonic network is a chain of 15 dependent shuffles)
        const int v = lane < cnt ? r[lo + lane] : 0x7fffffff;
        int rank = 0;
        for (int u = 0; u < cnt; ++u) rank += __shfl_sync(0xffffffffu, v, u) < v ? 1 : 0;
        __syncwarp();
        if (lane < cnt) r[lo + rank] = v;
    } else {
        // long segment (many queries share a neighbour, e.g. zero-padded returns piled on one point): warp-cooperative bitonic
        // network in place, O(n log^2 n / 32) instead of a single lane's O(n^2) insertion sort.  All compare-exchanges are
        // ascending (first step of every merge mirrored), so the virtual +inf padding up to a power of two stays above `cnt`
        // and is never touched.
        int* seg = r + lo;
        int npow2 = 64;
        while (npow2 < cnt) npow2 <<= 1;
        const int half = npow2 >> 1;
        for (int k = 2; k <= npow2; k <<= 1) {
            const int hk = k >> 1;
            for (int i = lane; i < half; i += 32) {
                const int blk = i / hk, pos = i - blk * hk;
                const int a = blk * k + pos, c = blk * k + (k - 1 - pos);
                if (c < cnt) {
                    const int va = seg[a], vc = seg[c];
                    if (va > vc) { seg[a] = vc; seg[c] = va; }
                }
            }
            __syncwarp();
            for (int jj = k >> 2; jj >= 1; jj >>= 1) {
                for (int i = lane; i < half; i += 32) {
                    const int a = 2 * jj * (i / jj) + (i % jj), c = a + jj;
                    if (c < cnt) {
                        const int va = seg[a], vc = seg[c];
                        if (va > vc) { seg[a] = vc; seg[c] = va; }
                    }
                }
                __syncwarp();
            }
        }
    }
}

// workspace ints per batch item: (n_seg + 1) offsets + n_seg cursors + L rows
extern "C" long long ssf_csr_workspace_ints(int B, int L, int n_seg) {
    return (long long)B * ((long long)(n_seg + 1) + n_seg + L);
}

template <typename IdxT>
static int build_csr(const IdxT* key, int B, int L, int n_seg, int* ws, cudaStream_t st) {
    int* offset = ws;
    int* cursor = ws + (size_t)B * (n_seg + 1);
    int* rows = cursor + (size_t)B * n_seg;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(int) * ((size_t)B * (n_seg + 1) + (size_t)B * n_seg), st);
    if (e != cudaSuccess) return ssf_set_error(e);
    dim3 gl((L + 255) / 256, B);
    csr_count_kernel<IdxT><<<gl, 256, 0, st>>>(key, L, n_seg, offset);
    csr_scan_kernel<<<B, 1024, 0, st>>>(offset, n_seg);
    csr_fill_kernel<IdxT><<<gl, 256, 0, st>>>(key, L, n_seg, offset, cursor, rows);
    dim3 gs((n_seg + 7) / 8, B);
    csr_sort_kernel<<<gs, 256, 0, st>>>(offset, n_seg, L, rows);
    for (int i = 0; i < 4; ++i) ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

extern "C" int ssf_build_csr_i64(const long long* key, int B, int L, int n_seg, int* ws, void* stream) {
    return build_csr<long long>(key, B, L, n_seg, ws, (cudaStream_t)stream);
}
extern "C" int ssf_build_csr_i32(const int* key, int B, int L, int n_seg, int* ws, void* stream) {
    return build_csr<int>(key, B, L, n_seg, ws, (cudaStream_t)stream);
}

// src [B,L,C] -> out [B,L,C]: softmax over the rows of each segment, per channel (torch_scatter semantics)
__global__ void __launch_bounds__(128)
seg_softmax_kernel(const float* __restrict__ src, const int* __restrict__ ws, int B, int L, int C, int n_seg,
                   float* __restrict__ out) {
    const int b = blockIdx.y;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_seg * C) return;
    const int j = (int)(t / C), c = (int)(t % C);
    const int* off = ws + (size_t)b * (n_seg + 1);
    const int* rows = ws + (size_t)B * (n_seg + 1) + (size_t)B * n_seg + (size_t)b * L;
    const float* s = src + (size_t)b * L * C;
    float* o = out + (size_t)b * L * C;
    const int lo = off[j], hi = off[j + 1];
    float m = -INFINITY;
    for (int a = lo; a < hi; ++a) m = fmaxf(m, s[(size_t)rows[a] * C + c]);
    float den = 0.f;
    for (int a = lo; a < hi; ++a) den += expf(s[(size_t)rows[a] * C + c] - m);
    for (int a = lo; a < hi; ++a) {
        const size_t p = (size_t)rows[a] * C + c;
        o[p] = expf(s[p] - m) / den;
    }
}

// src [B,L,C] -> out [B,n_seg,C]: rows of a segment added in ascending row order
__global__ void __launch_bounds__(128)
seg_sum_kernel(const float* __restrict__ src, const int* __restrict__ ws, int B, int L, int C, int n_seg,
               float* __restrict__ out) {
    const int b = blockIdx.y;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_seg * C) return;
    const int j = (int)(t / C), c = (int)(t % C);
    const int* off = ws + (size_t)b * (n_seg + 1);
    const int* rows = ws + (size_t)B * (n_seg + 1) + (size_t)B * n_seg + (size_t)b * L;
    const float* s = src + (size_t)b * L * C;
    const int lo = off[j], hi = off[j + 1];
    float acc = 0.f;
    for (int a = lo; a < hi; ++a) acc += s[(size_t)rows[a] * C + c];
    out[((size_t)b * n_seg + j) * C + c] = acc;
}

extern "C" int ssf_segment_softmax(const float* src, const int* csr_ws, int B, int L, int C, int n_seg, float* out,
                                   void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || L <= 0 || C <= 0 || n_seg <= 0) return ssf_arg_error("segment_softmax: empty input");
    dim3 grid((unsigned)(((long long)n_seg * C + 127) / 128), B);
    seg_softmax_kernel<<<grid, 128, 0, st>>>(src, csr_ws, B, L, C, n_seg, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

extern "C" int ssf_segment_sum(const float* src, const int* csr_ws, int B, int L, int C, int n_seg, float* out,
                               void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || L <= 0 || C <= 0 || n_seg <= 0) return ssf_arg_error("segment_sum: empty input");
    dim3 grid((unsigned)(((long long)n_seg * C + 127) / 128), B);
    seg_sum_kernel<<<grid, 128, 0, st>>>(src, csr_ws, B, L, C, n_seg, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
