// Generic tensor-core dense layer for the 1x1-conv stacks of TFlow (ASF/utils/utils.py:236-247,304-313;
// ASF/utils/soflow.py:397-451,460-461,501-513): Y[rows, N] = epilogue(A[rows, K] . W[N, K]^T) on tcgen05
// `kind::tf32` with the fp32-faithful 3xTF32 split (tc_common.cuh), fp32 accumulators in TMEM.
//
// Persistent kernel: grid = (min(row tiles, SMs), column tiles of <= 256); every CTA walks 128-row tiles.  Roles of the base
// variant (the other role layouts -- wide, full, light -- are described at the kernel template below):
//   warps 0-3 / 4-7  two A-producer warpgroups (thread = row = TMEM lane) taking alternate tiles.  A is streamed in
//                    K chunks of 32: either plain rows of one or two concatenated inputs, or the *grouped first layer*
//                    evaluated on the fly,
//                       A[(b,n,s), c] = act1(G[b, idx[b,n,s], offG + c] + H[b, n, offH + c] + b1[c] + Wd1[:, c] . dir),
//                    so a gathered neighbourhood never exists in HBM.  Each chunk is split into (hi, lo) TF32 halves and
//                    written to a ring of TMEM stages with tcgen05.st.  The rows come in by cp.async, two chunks ahead,
//                    through a warp-private ring of staging tiles; neighbour indices and positions one tile ahead.
//   warps 8-11       epilogue (thread = row): tcgen05.ld of the accumulator, then
//                       STORE: y[row, :] = act(D + bias [+ Hq[row / S] + Wd2 . dir(row)])   (staged through shared memory so
//                              that full 128-byte row segments leave per store instruction)
//                       MAX  : y[row / S, :] = max over the S rows of a point (S in {8, 16}; butterfly over lanes)
//                       DOT  : y[row] = wvec . act(D + bias) + b0                      (weightnet1's last conv)
//   warp 12          one thread issues the MMAs (12 per chunk: 3 passes x 4 K-steps of 8)
//   warp 13          one thread brings the weight image in with cp.async.bulk (TMA): the host pre-arranges, per 32-wide
//                    K chunk, the [N x 32] hi and lo images in the no-swizzle K-major UMMA layout.  Layers whose image
//                    fits in shared memory keep it resident for the whole kernel; wider ones stream it through a ring.
// The accumulator is double buffered in TMEM when 2 N + 256 <= 512 columns, so the epilogue of one tile overlaps the
// MMAs of the next.
#include <cstdlib>
#include <cstring>
#include <cuda.h>   // CUtensorMap and the cuTensorMapEncodeTiled signature (types only: the entry point is fetched through cudart)
#include "tc_common.cuh"
#include "ssf_dense.h"

#ifdef SSF_CV_TRACE
// [role][event] accumulated cycles of CTA 0 (role 0: producer warp 0, 1: epilogue warp, 2: MMA warp); event 15 = count
__device__ long long g_dt_trace[3 * 16];
#define DTRACE_DECL long long tr_t = clock64(), tr_n
#define DTRACE(role, ev) do { if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0) { tr_n = clock64(); atomicAdd((unsigned long long*)&g_dt_trace[(role) * 16 + (ev)], (unsigned long long)(tr_n - tr_t)); tr_t = tr_n; } } while (0)
#else
#define DTRACE_DECL
#define DTRACE(role, ev) do {} while (0)
#endif

namespace {

constexpr int KC = 32;                    // K chunk
constexpr int AS = 4;                     // TMEM A stages (64 columns each: hi | lo)
constexpr int DT_THREADS = 448;           // heavy variant: 2 producer warpgroups, 1 CTA per SM
constexpr int DT_THREADS_FULL = 576;      // full variant: 2 producer + 2 epilogue warpgroups (96 registers per thread)
constexpr int DT_THREADS_LIGHT = 288;     // light variant: 1 producer warpgroup, 2 CTAs per SM (small layers)
constexpr int W_SMEM_MAX = 150 * 1024;    // bytes of shared memory for weight images
constexpr int LIGHT_SMEM_MAX = 112 * 1024; // per-CTA shared memory of the light variant (two CTAs per SM)
constexpr int STG_LD = KC + 4;            // row stride (floats) of a producer warp's transpose tile

// Tensor maps (TMA descriptors) of the plain-row inputs and of the stored output: 2-D fp32 tensors [rows, channels] with
// 32 x 32 boxes and the 128-byte swizzle, encoded on the host per launch and handed to the kernel as a __grid_constant__
// parameter.  A producer warp then fetches its 32 rows x 32 channels of a K chunk with ONE cp.async.bulk.tensor issued by one
// lane (rows beyond the tensor are zero-filled by the TMA engine), and a STORE-epilogue warp writes its 32 x 32 output box
// with one cp.async.bulk.tensor store (rows beyond the tensor are clipped) -- instead of 8 address computations + 8 cp.async,
// resp. 8 shared-memory loads + 8 predicated global stores, per lane.
struct alignas(64) DenseMaps {
    CUtensorMap x1, x2, y;
    CUtensorMap g;   // grouped mode: the source rows G [B * Nsrc, ldG] with 32 x 1 boxes, fetched four rows at a time (tile::gather4)
};

struct DenseCfg {
    int tma_a;       // plain-row A operand fetched by tensor-map TMA
    int tma_y;       // STORE epilogue written by tensor-map TMA
    int tma_g;       // grouped A operand: neighbour rows gathered by tensor-map TMA (tile::gather4)
    int Nt;          // columns of a CTA tile
    int nk;          // K chunks
    int nstage;      // weight stages in shared memory
    int resident;    // whole image resident (nk <= nstage)
    int nd;          // accumulator buffers
    int n_tiles;     // row tiles
    int pd;          // A chunks a producer warp keeps in flight ahead of the one it converts (ring of pd + 1 staging tiles)
    // division of a row / point number (< 2^31) by the invariant S resp. Nq as multiply-high + add + shift (Granlund-Montgomery):
    // the rows are 64-bit in the interface, and a 64-bit division is ~70 instructions -- there were six of them per thread per tile
    unsigned s_mul, s_sh, nq_mul, nq_sh;
};

__device__ __forceinline__ unsigned dt_fastdiv(unsigned n, unsigned mul, unsigned sh) { return ssf_fastdiv(n, mul, sh); }
static void dt_fastdiv_make(unsigned d, unsigned* mul, unsigned* sh) {
    const SsfFastDiv f = ssf_fastdiv_make(d);
    *mul = f.mul;
    *sh = f.sh;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ssf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ssf_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ssf_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// ---- tensor-map TMA (SASS: UTMALDG / UTMASTG).  Boxes are 32 rows x 128 bytes with CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte
// chunk j of box row r sits at r * 128 + ((j ^ (r & 7)) << 4) of the 1024-byte aligned tile, so thread = row accesses are
// bank-conflict free without padding.
constexpr int TMA_TILE_BYTES = 32 * 128;
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c_inner, int c_row, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     ssf_smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_row), "r"(ssf_smem_u32(bar))
                 : "memory");
}
// four rows r0..r3 of a 2-D tensor (box = 32 channels x 1 row), 128 bytes each, land consecutively at dst (SASS: UTMALDG.2D.GATHER4)
__device__ __forceinline__ void tma_gather4(void* dst_smem, const CUtensorMap* map, int c_inner, int r0, int r1, int r2, int r3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                     ssf_smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(ssf_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c_inner, int c_row, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(ssf_smem_u32(src_smem)), "r"(c_inner), "r"(c_row)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t tma_swz(int row, int chunk16) { return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4)); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return fmaxf(v, 0.1f * v);
    return v;
}
__device__ __forceinline__ void split8(const float (&v)[8], float (&hi)[8], float (&lo)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#ifdef SSF_SPLIT_TRUNC
        // the tensor core reads an fp32 container as TF32 by dropping the 13 low mantissa bits, so the value itself serves as
        // its truncated `hi` and only lo = v - trunc(v) (exact) has to be computed: 2 instead of 3 operations per element
        hi[j] = v[j];
        lo[j] = v[j] - __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);
#else
        hi[j] = __uint_as_float((__float_as_uint(v[j]) + 0x1000u) & 0xFFFFE000u);
        lo[j] = v[j] - hi[j];
#endif
    }
}
__device__ __forceinline__ void ldg8(const float* p, float (&o)[8]) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p)), y = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
}
__device__ __forceinline__ void lds8(const float* p, float (&o)[8]) {
    const float4 x = *reinterpret_cast<const float4*>(p), y = *reinterpret_cast<const float4*>(p + 4);
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
}

// per-row geometry of a tile
struct RowCtx {   // (rows < 2^31: 32-bit indices, widened only when they scale a pointer)
    unsigned row;     // global row (may be >= rows in the last tile)
    unsigned rowc;    // clamped
    unsigned pt;      // flat point index (row / S)
    unsigned srow;    // flat source row b * Nsrc + idx
    float px, py, pz; // pos_src[srow]
    float qx, qy, qz; // pos_q[pt]
};

// VARIANT 0 (heavy): two producer warpgroups, one epilogue warpgroup, 448 threads, all 512 TMEM columns, one CTA per SM.
// VARIANT 1 (light): for layers whose weight image is small: one producer warpgroup, 288 threads, 256 TMEM columns, <= 112 KB
//                    shared memory -> two CTAs per SM, i.e. two independent tile pipelines (the MMA warp also fetches the
//                    resident weights).
// VARIANT 3 (full):  two producer and two epilogue warpgroups (576 threads, 96 registers): grouped layers with a small weight
//                    image, where the gather needs both producer warpgroups and the pooling epilogue is the bottleneck.
// VARIANT 2 (wide):  for 256-column tiles, whose accumulator cannot be double buffered (the epilogue and the next tile's MMAs
//                    serialise): one producer warpgroup with all four A stages and TWO epilogue warpgroups, each draining
//                    half of the columns.
template <int VARIANT>
__global__ void __launch_bounds__(VARIANT == 1 ? DT_THREADS_LIGHT : (VARIANT == 3 ? DT_THREADS_FULL : DT_THREADS), VARIANT == 1 ? 2 : 1)
dense_tc_kernel(ssf_dense_args a, DenseCfg cfg, const __grid_constant__ DenseMaps maps) {
    constexpr int NPROD = (VARIANT == 0 || VARIANT == 3) ? 2 : 1;   // producer warpgroups
    constexpr int NEPI = VARIANT >= 2 ? 2 : 1;                      // epilogue warpgroups
    constexpr int SPW = VARIANT == 2 ? 4 : 2;                       // A stages per producer warpgroup
    constexpr int NTHR = VARIANT == 1 ? DT_THREADS_LIGHT : (VARIANT == 3 ? DT_THREADS_FULL : DT_THREADS);
    constexpr int W_EPI = 4 * NPROD;                                // first epilogue warp
    constexpr int W_MMA = W_EPI + 4 * NEPI, W_TMA = VARIANT == 1 ? W_MMA : W_MMA + 1;
    constexpr uint32_t TCOLS = VARIANT == 1 ? 256 : 512;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Nt = cfg.Nt, nk = cfg.nk;
    const int n0 = blockIdx.y * 256;                       // first output column of this CTA
    const uint32_t wchunk = (uint32_t)Nt * KC * 4 * 2;     // bytes of one weight chunk (hi + lo)
    uint8_t* sW = smem;
    float* sPar = reinterpret_cast<float*>(smem + (size_t)cfg.nstage * wchunk);
    float* sB1 = sPar;                 // [K]
    float* sWd1 = sB1 + a.K;           // [3][K]
    float* sBias = sWd1 + 3 * a.K;     // [Nt]
    float* sWd2 = sBias + Nt;          // [3][Nt]
    float* sWvec = sWd2 + 3 * Nt;      // [Nt]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sWvec + Nt);
    uint64_t* w_full = bars;                        // [nstage] (<= 16)
    uint64_t* w_empty = bars + 16;                  // [nstage]
    uint64_t* a_ready = bars + 32;                  // [AS]
    uint64_t* a_empty = bars + 36;                  // [AS]
    uint64_t* d_full = bars + 40;                   // [2]
    uint64_t* d_empty = bars + 42;                  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 44);
    uint64_t* a_tma = bars + 46;                    // [8 producer warps][4 slots]: completion of the tensor-map loads
    float* sStage = reinterpret_cast<float*>(bars + 78);   // [4 NPROD producer warps][pd + 1 slots][32 rows][STG_LD]
    float* sEpi = sStage + 4 * NPROD * (cfg.pd + 1) * (32 * STG_LD);   // [4 NEPI epilogue warps][32 rows][STG_LD]  (STORE epilogue)
    // tensor-map tiles: 4 KB each, 1024-byte aligned, carved out of the SAME regions as the padded tiles they replace (those
    // are 4.5 KB each, so the aligned 4 KB tiles of the same count (>= 4) always fit, and a kernel that uses the tensor-map
    // path for only one of the two roles never has the other role's padded tiles in the way)
    uint8_t* sTmaA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sStage) + 1023) & ~static_cast<uintptr_t>(1023));   // [producer warp][slot]
    uint8_t* sTmaE = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sEpi) + 1023) & ~static_cast<uintptr_t>(1023));     // [epilogue warp]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == W_MMA) tc_alloc(tmem_slot, TCOLS);
    if (tid == 0) {
        for (int i = 0; i < cfg.nstage; ++i) {
            ssf_mbar_init(&w_full[i], 1);
            ssf_mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < AS; ++i) {
            ssf_mbar_init(&a_ready[i], 128);
            ssf_mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ssf_mbar_init(&d_full[i], 1);
            ssf_mbar_init(&d_empty[i], 128 * NEPI);
        }
        for (int i = 0; i < 32; ++i) ssf_mbar_init(&a_tma[i], 1);
        ssf_mbar_fence_init();
    }
    for (int i = tid; i < a.K; i += NTHR) {
        sB1[i] = a.b1 ? __ldg(a.b1 + i) : 0.f;
        for (int c = 0; c < 3; ++c) sWd1[c * a.K + i] = a.Wd1 ? __ldg(a.Wd1 + c * a.K + i) : 0.f;
    }
    for (int i = tid; i < Nt; i += NTHR) {
        sBias[i] = a.bias ? __ldg(a.bias + n0 + i) : 0.f;
        for (int c = 0; c < 3; ++c) sWd2[c * Nt + i] = a.Wd2 ? __ldg(a.Wd2 + c * a.N + n0 + i) : 0.f;
        sWvec[i] = a.wvec ? __ldg(a.wvec + n0 + i) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t a_col0 = (uint32_t)(cfg.nd * Nt);
    const int n_my = ((int)blockIdx.x < cfg.n_tiles) ? (cfg.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    auto produce_weights = [&]() {   // ---- weight producer (one thread)
        const uint8_t* wsrc = static_cast<const uint8_t*>(a.wimg) + (size_t)blockIdx.y * nk * wchunk;
        if (cfg.resident) {
            if (n_my > 0)
                for (int kc = 0; kc < nk; ++kc) {
                    ssf_mbar_expect_tx(&w_full[kc], wchunk);
                    ssf_bulk_g2s(sW + (size_t)kc * wchunk, wsrc + (size_t)kc * wchunk, wchunk, &w_full[kc]);
                }
        } else {
            const int total = n_my * nk;
            for (int g = 0; g < total; ++g) {
                const int st = g % cfg.nstage;
                if (g >= cfg.nstage) ssf_mbar_wait(&w_empty[st], (uint32_t)((g / cfg.nstage - 1) & 1));
                ssf_mbar_expect_tx(&w_full[st], wchunk);
                ssf_bulk_g2s(sW + (size_t)st * wchunk, wsrc + (size_t)(g % nk) * wchunk, wchunk, &w_full[st]);
            }
        }
    };
    if (VARIANT != 1 && warp == W_TMA) {
        if (lane == 0) produce_weights();
        __syncwarp();
    } else if (warp == W_MMA) {
        if (VARIANT == 1) {
            if (lane == 0) produce_weights();
            __syncwarp();
        }
        {   // ---- MMA issuer: warp-uniform control flow, one elected lane executes the tcgen05 instructions
            const bool leader = tc_elect_one();
            const uint32_t idesc = tc_idesc_tf32(128, Nt);
            const uint32_t lbo = (uint32_t)(Nt / 8) * 128;
            const uint64_t lo_off = (uint64_t)(((uint32_t)Nt * KC * 4) >> 4), k_step = (uint64_t)((2 * lbo) >> 4);
            for (int it = 0; it < n_my; ++it) {
                const int db = it & (cfg.nd - 1);   // nd is 1 or 2
                if (it >= cfg.nd) {
                    ssf_mbar_wait(&d_empty[db], (uint32_t)(((it >> (cfg.nd >> 1)) - 1) & 1));
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem + (uint32_t)(db * Nt);
                for (int kc = 0; kc < nk; ++kc) {
                    const int g = it * nk + kc;
                    const int j = (it / NPROD) * nk + kc;             // chunk count of the producing warpgroup
                    const int ws = cfg.resident ? kc : g % cfg.nstage, as = (it % NPROD) * SPW + (j % SPW);
                    if (cfg.resident) {
                        if (it == 0) ssf_mbar_wait(&w_full[ws], 0);
                    } else {
                        ssf_mbar_wait(&w_full[ws], (uint32_t)((g / cfg.nstage) & 1));
                    }
                    ssf_mbar_wait(&a_ready[as], (uint32_t)((j / SPW) & 1));
                    tc_fence_after();
                    // the descriptors of the 4 K-steps / hi-lo images differ only in the start-address field
                    const uint64_t w_hi = tc_smem_desc(ssf_smem_u32(sW + (size_t)ws * wchunk), lbo, 128);
                    const uint32_t a_hi = tmem + a_col0 + as * 64;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t aa = a_hi + (pass == 0 ? 32u : 0u);
                        const uint64_t ww = w_hi + (pass == 1 ? lo_off : 0);
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks)
                            if (leader) tc_mma_ts(d_tmem, aa + ks * 8, ww + ks * k_step, idesc, (kc > 0 || pass > 0 || ks > 0) ? 1u : 0u);
                    }
                    if (leader) {
                        tc_commit(&a_empty[as]);
                        if (!cfg.resident) tc_commit(&w_empty[ws]);
                    }
                    __syncwarp();
                }
                if (leader) tc_commit(&d_full[db]);
                __syncwarp();
            }
        }
        __syncwarp();
    } else if (warp < W_EPI) {
        // ---- A producers: warpgroup wg takes tiles it = wg, wg + 2, ...
        const int wg = warp >> 2;
        DTRACE_DECL;
        const int r = tid & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const bool grouped = a.S > 0 && a.idx != nullptr;
        const bool need_dir = grouped && a.Wd1 != nullptr;

        const unsigned n_rows = (unsigned)a.rows;
        auto row_of = [&](int it) -> unsigned { return (blockIdx.x + (unsigned)it * gridDim.x) * 128u + (unsigned)r; };
        auto clampr = [&](unsigned row) -> unsigned { return row < n_rows ? row : n_rows - 1u; };
        auto load_idx = [&](int it) -> int {   // neighbour index of this thread's row in tile `it`
            if (!grouped || it >= n_my) return 0;
            return __ldg(a.idx + clampr(row_of(it)));
        };
        auto make_ctx = [&](int it, int id) -> RowCtx {   // issues the position loads (if needed)
            RowCtx c;
            c.row = row_of(it < n_my ? it : n_my - 1);
            c.rowc = clampr(c.row);
            c.pt = a.S > 0 ? dt_fastdiv(c.rowc, cfg.s_mul, cfg.s_sh) : 0u;
            c.srow = 0;
            c.px = c.py = c.pz = c.qx = c.qy = c.qz = 0.f;
            if (grouped) {
                c.srow = dt_fastdiv(c.pt, cfg.nq_mul, cfg.nq_sh) * (unsigned)a.Nsrc + (unsigned)id;
                if (need_dir) {
                    const float* ps = a.pos_src + (size_t)c.srow * 3u;
                    const float* pq = a.pos_q + (size_t)c.pt * 3u;
                    c.px = __ldg(ps); c.py = __ldg(ps + 1); c.pz = __ldg(ps + 2);
                    c.qx = __ldg(pq); c.qy = __ldg(pq + 1); c.qz = __ldg(pq + 2);
                }
            }
            return c;
        };
        // Coalesced asynchronous loads: 8 lanes x 16 bytes cover the 128-byte K chunk of one row, 4 rows per instruction (a
        // thread reading its own row would touch 32 different lines per instruction).  cp.async puts the pieces straight into
        // this warp's ring of staging tiles, `pd` chunks ahead of the chunk being converted (no registers held by loads in
        // flight); the chunk is then read back with thread = row.  The prefetch pointer runs through the same (tile, chunk)
        // sequence as the conversion loop below, with its own copy of the row indices.
        const int rg = lane >> 3, pc = lane & 7;
        const int PD = cfg.pd, D = PD + 1;
        float* stg = sStage + warp * D * (32 * STG_LD);
        const int total_n = wg < n_my ? ((n_my - wg + NPROD - 1) / NPROD) * nk : 0;
        auto my_of = [&](int t, int id) -> int {   // row of this thread in the source array for the warpgroup's t-th tile
            const int it = wg + t * NPROD;
            const unsigned rowc = clampr(row_of(it < n_my ? it : n_my - 1));
            if (a.a_mode == 0) return (int)rowc;
            const unsigned pt = dt_fastdiv(rowc, cfg.s_mul, cfg.s_sh);
            return (int)(dt_fastdiv(pt, cfg.nq_mul, cfg.nq_sh) * (unsigned)a.Nsrc + (unsigned)id);
        };
        int pf_n = 0, pf_t = 0, pf_kc = 0;
        int pf_slot = 0, tk_slot = 0;        // pf_n % D and n % D of the ring, kept as counters (D is a run-time 2 or 3)
        uint32_t tk_phase = 0;               // (n / D) & 1
        int my_pf = my_of(0, load_idx(wg));
        int idx_nx = load_idx(wg + NPROD);
        const bool tma_g = cfg.tma_g != 0;
        const bool tma_a = cfg.tma_a != 0 || tma_g;   // either way the chunk arrives in this warp's ring of swizzled tensor-map tiles
        uint8_t* ttile = sTmaA + (size_t)warp * D * TMA_TILE_BYTES;     // this warp's ring of tensor-map tiles
        uint64_t* tbar = a_tma + warp * 4;
        auto issue_next = [&]() {   // cp.async of the next chunk of the sequence (if any) + one commit group, always
            if (tma_g) {
                // gathered neighbour rows: lanes 0..7 each fetch four of the warp's 32 source rows (128 bytes of each) with one
                // tile::gather4 -- no per-lane address arithmetic, no load-store-unit queue slots held while the rows travel
                if (pf_n < total_n) {
                    const int k0 = pf_kc * KC;
                    const int slot = pf_slot;
                    const int g0 = (lane & 7) * 4;
                    const int r0 = __shfl_sync(0xffffffffu, my_pf, g0), r1 = __shfl_sync(0xffffffffu, my_pf, g0 + 1);
                    const int r2 = __shfl_sync(0xffffffffu, my_pf, g0 + 2), r3 = __shfl_sync(0xffffffffu, my_pf, g0 + 3);
                    if (lane == 0) ssf_mbar_expect_tx(&tbar[slot], TMA_TILE_BYTES);
                    __syncwarp();
                    if (lane < 8) tma_gather4(ttile + slot * TMA_TILE_BYTES + lane * 512, &maps.g, a.offG + k0, r0, r1, r2, r3, &tbar[slot]);
                    if (++pf_kc == nk) {
                        pf_kc = 0;
                        ++pf_t;
                        my_pf = my_of(pf_t, idx_nx);
                        idx_nx = load_idx(wg + (pf_t + 1) * NPROD);
                    }
                }
                ++pf_n;
                if (++pf_slot == D) pf_slot = 0;
                return;
            }
            if (tma_a) {
                // one elected lane fetches the warp's 32 rows x 32 channels box; rows past the end are zero-filled
                if (pf_n < total_n) {
                    const int k0 = pf_kc * KC;
                    const int it_pf = wg + pf_t * NPROD;
                    const int row0 = (int)(((long long)blockIdx.x + (long long)it_pf * gridDim.x) * 128 + (warp & 3) * 32);
                    DTRACE(0, 7);    // (trace build) address arithmetic of the prefetch
                    if (lane == 0) {
                        const int slot = pf_slot;
                        ssf_mbar_expect_tx(&tbar[slot], TMA_TILE_BYTES);
                        DTRACE(0, 8);    // expect_tx
                        if (k0 < a.c1) tma_load_2d(ttile + slot * TMA_TILE_BYTES, &maps.x1, k0, row0, &tbar[slot]);
                        else tma_load_2d(ttile + slot * TMA_TILE_BYTES, &maps.x2, k0 - a.c1, row0, &tbar[slot]);
                        DTRACE(0, 9);    // tensor-map load issue
                    }
                    if (++pf_kc == nk) {
                        pf_kc = 0;
                        ++pf_t;
                    }
                }
                ++pf_n;
                if (++pf_slot == D) pf_slot = 0;
                return;
            }
            if (pf_n < total_n) {
                const int k0 = pf_kc * KC;
                const float* base;
                unsigned ld;
                if (a.a_mode == 0) {
                    if (k0 < a.c1) { base = a.x1 + k0; ld = (unsigned)a.ld1; } else { base = a.x2 + (k0 - a.c1); ld = (unsigned)a.ld2; }
                } else {
                    base = a.G + a.offG + k0;
                    ld = (unsigned)a.ldG;
                }
                float* dst = stg + pf_slot * (32 * STG_LD) + pc * 4;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned rrow = (unsigned)__shfl_sync(0xffffffffu, my_pf, j * 4 + rg);
                    const float* src = base + ((size_t)rrow * ld + (unsigned)(pc * 4));   // one 32 x 32 -> 64-bit multiply-add
                    if (a.a_mode == 0) cp_async16_cg(dst + (j * 4 + rg) * STG_LD, src);   // streamed once
                    else cp_async16_ca(dst + (j * 4 + rg) * STG_LD, src);                 // gathered rows are re-used by neighbours
                }
                if (++pf_kc == nk) {
                    pf_kc = 0;
                    ++pf_t;
                    my_pf = my_of(pf_t, idx_nx);
                    idx_nx = load_idx(wg + (pf_t + 1) * NPROD);
                }
            }
            ++pf_n;
            if (++pf_slot == D) pf_slot = 0;
            cp_async_commit();
        };
        auto take_chunk = [&](int n, float (&v)[4][8]) {   // waits for chunk n of the sequence, reads this thread's row
            if (tma_a) {
                const int slot = tk_slot;
                ssf_mbar_wait(&tbar[slot], tk_phase);
                const uint8_t* src = ttile + slot * TMA_TILE_BYTES;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 x = *reinterpret_cast<const float4*>(src + tma_swz(lane, 2 * q));
                    const float4 y = *reinterpret_cast<const float4*>(src + tma_swz(lane, 2 * q + 1));
                    v[q][0] = x.x; v[q][1] = x.y; v[q][2] = x.z; v[q][3] = x.w; v[q][4] = y.x; v[q][5] = y.y; v[q][6] = y.z; v[q][7] = y.w;
                }
                __syncwarp();   // the slot is refilled by the tensor-map load issued for chunk n + D in the next step
                if (++tk_slot == D) { tk_slot = 0; tk_phase ^= 1u; }
                return;
            }
            if (PD == 1) cp_async_wait<1>();
            else cp_async_wait<2>();
            __syncwarp();
            const float* src = stg + tk_slot * (32 * STG_LD) + lane * STG_LD;
#pragma unroll
            for (int q = 0; q < 4; ++q) lds8(src + q * 8, v[q]);
            __syncwarp();   // the slot is overwritten by the cp.async issued for chunk n + D in the next step
            if (++tk_slot == D) { tk_slot = 0; tk_phase ^= 1u; }
        };

        if (wg < n_my) {
            RowCtx cur, nxt;
            int idx2;   // neighbour index two tiles (of this warpgroup) ahead
            float v[4][8];
            nxt = make_ctx(wg, load_idx(wg));
            idx2 = load_idx(wg + NPROD);
            for (int m = 0; m < PD; ++m) issue_next();
            int n_seq = 0;
            for (int it = wg; it < n_my; it += NPROD) {
                cur = nxt;
                nxt = make_ctx(it + NPROD, idx2);   // positions of the next tile: in flight while this tile is processed
                idx2 = load_idx(it + 2 * NPROD);
                const float dx = cur.px - cur.qx, dy = cur.py - cur.qy, dz = cur.pz - cur.qz;
                for (int kc = 0; kc < nk; ++kc) {
                    DTRACE(0, 0);   // rest of the loop body (previous arrive .. here)
                    issue_next();
                    DTRACE(0, 1);   // cp.async issue of the chunk `pd` ahead
                    take_chunk(n_seq++, v);
                    DTRACE(0, 2);   // wait for this chunk + read back with thread = row
                    const int k0 = kc * KC;
                    if (a.a_mode == 1) {
                        if (a.H != nullptr) {
                            const float* hs = a.H + ((size_t)cur.pt * (unsigned)a.ldH + (unsigned)(a.offH + k0));
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float h[8];
                                ldg8(hs + q * 8, h);
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[q][j] += h[j];
                            }
                        }
                        // bias + direction term on packed fp32x2 FMAs (two channels per instruction), then the activation
                        const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float bb[8];
                            lds8(sB1 + k0 + q * 8, bb);
                            float2 acc[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[j] = __fadd2_rn(make_float2(v[q][2 * j], v[q][2 * j + 1]), make_float2(bb[2 * j], bb[2 * j + 1]));
                            if (need_dir) {
                                float w0[8], w1[8], w2[8];
                                lds8(sWd1 + k0 + q * 8, w0);
                                lds8(sWd1 + a.K + k0 + q * 8, w1);
                                lds8(sWd1 + 2 * a.K + k0 + q * 8, w2);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    acc[j] = __ffma2_rn(dx2, make_float2(w0[2 * j], w0[2 * j + 1]), acc[j]);
                                    acc[j] = __ffma2_rn(dy2, make_float2(w1[2 * j], w1[2 * j + 1]), acc[j]);
                                    acc[j] = __ffma2_rn(dz2, make_float2(w2[2 * j], w2[2 * j + 1]), acc[j]);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                v[q][2 * j] = act_apply(acc[j].x, a.act1);
                                v[q][2 * j + 1] = act_apply(acc[j].y, a.act1);
                            }
                        }
                    }
                    // each warpgroup owns SPW of the A stages and waits on them strictly in order (an mbarrier
                    // parity wait must never run more than one phase ahead of the barrier)
                    const int j = (it / NPROD) * nk + kc, as = wg * SPW + (j % SPW);
                    DTRACE(0, 3);   // first-layer math
                    if (j >= SPW) {
                        ssf_mbar_wait(&a_empty[as], (uint32_t)(((j / SPW) - 1) & 1));
                        tc_fence_after();
                    }
                    DTRACE(0, 4);   // wait for the A stage
                    const uint32_t t_hi = tmem + lane_base + a_col0 + as * 64;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float hi[8], lo[8];
                        split8(v[q], hi, lo);
                        tc_st8(t_hi + q * 8, hi);
                        tc_st8(t_hi + 32 + q * 8, lo);
                    }
                    DTRACE(0, 5);   // split + tcgen05.st issue
                    tc_st_wait();
                    DTRACE(0, 6);   // tcgen05.wait::st
                    tc_fence_before();
                    mbar_arrive(&a_ready[as]);
                    DTRACE(0, 15);
                }
            }
        }
    } else {
        // ---- epilogue warps 8-11 (thread = row)
        const int r = tid & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const bool need_dir = a.Wd2 != nullptr;
        DTRACE_DECL;
        const int eg = (warp - W_EPI) >> 2;                            // epilogue warpgroup: columns [c_lo, c_hi) of the tile
        const int c_lo = eg * (Nt / NEPI), c_hi = c_lo + Nt / NEPI;
        for (int it = 0; it < n_my; ++it) {
            const int db = it & (cfg.nd - 1);   // nd is 1 or 2
            const long long row = ((long long)blockIdx.x + (long long)it * gridDim.x) * 128 + r;
            const bool valid = row < a.rows;
            const long long rowc = valid ? row : a.rows - 1;
            const long long pt = a.S > 0 ? (long long)dt_fastdiv((unsigned)rowc, cfg.s_mul, cfg.s_sh) : 0;
            float dx = 0.f, dy = 0.f, dz = 0.f;
            if (need_dir) {
                const long long b = (long long)dt_fastdiv((unsigned)pt, cfg.nq_mul, cfg.nq_sh);
                const float* ps = a.pos_src + (b * a.Nsrc + __ldg(a.idx + rowc)) * 3;
                const float* pq = a.pos_q + pt * 3;
                dx = __ldg(ps) - __ldg(pq);
                dy = __ldg(ps + 1) - __ldg(pq + 1);
                dz = __ldg(ps + 2) - __ldg(pq + 2);
            }
            DTRACE(1, 0);   // row setup
            ssf_mbar_wait(&d_full[db], (uint32_t)((it >> (cfg.nd >> 1)) & 1));
            tc_fence_after();
            DTRACE(1, 1);   // wait for the accumulator
            const uint32_t t_d = tmem + lane_base + (uint32_t)(db * Nt);
            float dot = 0.f;
            // bias, per-point rows, direction term and activation on 16 accumulator columns starting at tile column c
            auto finish16 = [&](int c, float (&v)[16]) {
                const int cg = n0 + c;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 bb = *reinterpret_cast<const float4*>(sBias + c + q * 4);
                    v[q * 4] += bb.x; v[q * 4 + 1] += bb.y; v[q * 4 + 2] += bb.z; v[q * 4 + 3] += bb.w;
                }
                if (a.Hq != nullptr) {
                    const float4* h4 = reinterpret_cast<const float4*>(a.Hq + pt * a.ldHq + cg);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 h = __ldg(h4 + q);
                        v[q * 4] += h.x; v[q * 4 + 1] += h.y; v[q * 4 + 2] += h.z; v[q * 4 + 3] += h.w;
                    }
                }
                if (need_dir) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w0 = *reinterpret_cast<const float4*>(sWd2 + c + q * 4);
                        const float4 w1 = *reinterpret_cast<const float4*>(sWd2 + Nt + c + q * 4);
                        const float4 w2 = *reinterpret_cast<const float4*>(sWd2 + 2 * Nt + c + q * 4);
                        v[q * 4] += dx * w0.x + dy * w1.x + dz * w2.x;
                        v[q * 4 + 1] += dx * w0.y + dy * w1.y + dz * w2.y;
                        v[q * 4 + 2] += dx * w0.z + dy * w1.z + dz * w2.z;
                        v[q * 4 + 3] += dx * w0.w + dy * w1.w + dz * w2.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = act_apply(v[j], a.act);
            };
            if (a.epi_mode == SSF_EPI_STORE) {
                // A thread storing its own row would touch 32 different lines per instruction (the LSU request rate, not
                // bandwidth, then bounds the epilogue).  32 columns per step: two TMEM loads in flight, the values are staged in
                // this warp's shared-memory tile and leave as full 128-byte row segments, four rows per store instruction.
                float* et = sEpi + (warp - W_EPI) * (32 * STG_LD);
                const int rg = lane >> 3, pc = lane & 7;
                const long long row0 = row - lane;     // first row of this warp's 32
                if (cfg.tma_y) {
                    // tensor-map store: the warp's 32 x 32 box is written swizzled into its 4 KB tile and leaves with one
                    // cp.async.bulk.tensor issued by one lane (rows past the end are clipped by the TMA engine)
                    uint8_t* tt = sTmaE + (size_t)(warp - W_EPI) * TMA_TILE_BYTES;
                    for (int c = c_lo; c < c_hi; c += 32) {
                        float va[16], vb[16];
                        tc_ld16(t_d + c, va);
                        tc_ld16(t_d + c + 16, vb);
                        tc_ld_wait();
                        finish16(c, va);
                        finish16(c + 16, vb);
                        if (lane == 0) tma_store_wait_read();   // the previous box has been read out of the tile
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            *reinterpret_cast<float4*>(tt + tma_swz(lane, q)) = make_float4(va[q * 4], va[q * 4 + 1], va[q * 4 + 2], va[q * 4 + 3]);
                            *reinterpret_cast<float4*>(tt + tma_swz(lane, 4 + q)) = make_float4(vb[q * 4], vb[q * 4 + 1], vb[q * 4 + 2], vb[q * 4 + 3]);
                        }
                        fence_proxy_async_smem();   // generic-proxy writes -> visible to the async proxy (TMA)
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&maps.y, n0 + c, (int)row0, tt);
                            tma_store_commit();
                        }
                    }
                } else
                for (int c = c_lo; c < c_hi; c += 32) {
                    float va[16], vb[16];
                    DTRACE(1, 3);   // previous iteration's tail
                    tc_ld16(t_d + c, va);
                    tc_ld16(t_d + c + 16, vb);
                    tc_ld_wait();
                    DTRACE(1, 4);   // tcgen05.ld + wait
                    finish16(c, va);
                    finish16(c + 16, vb);
                    DTRACE(1, 5);   // bias / per-point / direction terms, activation
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        *reinterpret_cast<float4*>(et + lane * STG_LD + q * 4) = make_float4(va[q * 4], va[q * 4 + 1], va[q * 4 + 2], va[q * 4 + 3]);
                        *reinterpret_cast<float4*>(et + lane * STG_LD + 16 + q * 4) = make_float4(vb[q * 4], vb[q * 4 + 1], vb[q * 4 + 2], vb[q * 4 + 3]);
                    }
                    DTRACE(1, 6);   // staging stores
                    __syncwarp();
                    float* dst = a.y + (n0 + c) + pc * 4;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int rl = j * 4 + rg;
                        if (row0 + rl < a.rows)
                            *reinterpret_cast<float4*>(dst + (row0 + rl) * a.ldy) = *reinterpret_cast<const float4*>(et + rl * STG_LD + pc * 4);
                    }
                    __syncwarp();
                    DTRACE(1, 7);   // coalesced row-segment stores
                }
            } else
            for (int c = c_lo; c < c_hi; c += 16) {
                float v[16];
                tc_ld16(t_d + c, v);
                tc_ld_wait();
                const int cg = n0 + c;   // global output column
                finish16(c, v);
                if (a.epi_mode == SSF_EPI_MAX) {
                    // butterfly max over the S rows (= lanes) of a point; rows beyond `rows` only exist in the last tile
                    // and belong to no stored point (rows % S == 0)
                    if (a.S == 16) {   // 16 values over 16 lanes: lane s ends up with column c + s
                        const bool u8 = (lane & 8) != 0, u4 = (lane & 4) != 0, u2 = (lane & 2) != 0, u1 = (lane & 1) != 0;
                        float w8[8], w4[4], w2[2];
#pragma unroll
                        for (int j = 0; j < 8; ++j) w8[j] = fmaxf(u8 ? v[j + 8] : v[j], __shfl_xor_sync(0xffffffffu, u8 ? v[j] : v[j + 8], 8));
#pragma unroll
                        for (int j = 0; j < 4; ++j) w4[j] = fmaxf(u4 ? w8[j + 4] : w8[j], __shfl_xor_sync(0xffffffffu, u4 ? w8[j] : w8[j + 4], 4));
#pragma unroll
                        for (int j = 0; j < 2; ++j) w2[j] = fmaxf(u2 ? w4[j + 2] : w4[j], __shfl_xor_sync(0xffffffffu, u2 ? w4[j] : w4[j + 2], 2));
                        const float w1 = fmaxf(u1 ? w2[1] : w2[0], __shfl_xor_sync(0xffffffffu, u1 ? w2[0] : w2[1], 1));
                        if (valid) a.y[pt * a.ldy + cg + (lane & 15)] = w1;
                    } else {           // S == 8: 16 values over 8 lanes: lane s ends up with columns c + 2 s, c + 2 s + 1
                        const bool u4 = (lane & 4) != 0, u2 = (lane & 2) != 0, u1 = (lane & 1) != 0;
                        float w8[8], w4[4], w2[2];
#pragma unroll
                        for (int j = 0; j < 8; ++j) w8[j] = fmaxf(u4 ? v[j + 8] : v[j], __shfl_xor_sync(0xffffffffu, u4 ? v[j] : v[j + 8], 4));
#pragma unroll
                        for (int j = 0; j < 4; ++j) w4[j] = fmaxf(u2 ? w8[j + 4] : w8[j], __shfl_xor_sync(0xffffffffu, u2 ? w8[j] : w8[j + 4], 2));
#pragma unroll
                        for (int j = 0; j < 2; ++j) w2[j] = fmaxf(u1 ? w4[j + 2] : w4[j], __shfl_xor_sync(0xffffffffu, u1 ? w4[j] : w4[j + 2], 1));
                        if (valid) *reinterpret_cast<float2*>(a.y + pt * a.ldy + cg + 2 * (lane & 7)) = make_float2(w2[0], w2[1]);
                    }
                } else {  // DOT
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(sWvec + c + q * 4);
                        dot = fmaf(v[q * 4], w.x, dot);
                        dot = fmaf(v[q * 4 + 1], w.y, dot);
                        dot = fmaf(v[q * 4 + 2], w.z, dot);
                        dot = fmaf(v[q * 4 + 3], w.w, dot);
                    }
                }
            }
            DTRACE(1, 2);   // column loop
            tc_fence_before();
            mbar_arrive(&d_empty[db]);
            if (a.epi_mode == SSF_EPI_DOT && valid) a.y[row] = dot + a.b0;
            DTRACE(1, 15);
        }
    }
    if (cfg.tma_y && lane == 0) tma_store_wait_all();   // (no-op for threads that issued no bulk store)
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tc_dealloc(tmem, TCOLS);
}

}  // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup (no link-time dependency on libcuda)
typedef CUresult (*ssf_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ssf_encode_tiled_fn encode_tiled_fn() {
    static ssf_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<ssf_encode_tiled_fn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}
// fp32 [rows, cols] row-major with leading dimension ld (floats): 32 x 32 boxes, 128-byte swizzle, zero fill / clipping out of bounds
static bool encode_rows_map(CUtensorMap* m, const float* base, long long rows, int cols, int ld, unsigned box_rows = 32) {
    ssf_encode_tiled_fn fn = encode_tiled_fn();
    if (fn == nullptr || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || ld % 4 != 0 || cols % 32 != 0 || rows <= 0 || rows > 0x7fffffffLL) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32, box_rows}, estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tensor-map TMA paths: 0 off, 1 (default) plain-row A operand + STORE epilogue, 2 also the grouped A operand by tile::gather4.
// Level 2 is correct but measured SLOWER than the per-lane cp.async gather on B200 (su0, 16.7 M rows: 2.11 vs 1.86 ms: the
// gathered rows are re-used by neighbouring queries and cp.async.ca keeps them in L1, the TMA engine goes to L2 every time), so it
// is not the default.  SSF_DENSE_TMA=0/1/2 or ssf_dense_set_tma() select the level (measurement / regression switch).
static int g_dense_tma = -1;
static int dense_tma() {
    if (g_dense_tma < 0) {
        const char* e = getenv("SSF_DENSE_TMA");
        g_dense_tma = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    return g_dense_tma;
}

extern "C" int ssf_dense_set_tma(int on) {
    const int prev = dense_tma();
    g_dense_tma = on < 0 ? 0 : (on > 2 ? 2 : on);
    return prev;
}

static int g_dense_variant = -1;   // 0: heavy variant everywhere, 1: light variant where it pays (default), 2: wherever eligible
static int dense_variant() {
    if (g_dense_variant < 0) {
        const char* e = getenv("SSF_DENSE_LIGHT");   // measurement switch for bench runs
        g_dense_variant = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    return g_dense_variant;
}
extern "C" int ssf_dense_set_variant(int light) {
    const int prev = dense_variant();
    g_dense_variant = light < 0 ? 0 : (light > 2 ? 2 : light);
    return prev;
}

#ifdef SSF_CV_TRACE
extern "C" int ssf_dense_trace_read(long long* out, int reset) {
    if (cudaMemcpyFromSymbol(out, g_dt_trace, sizeof(long long) * 48) != cudaSuccess) return 2;
    if (reset) { long long z[48] = {0}; cudaMemcpyToSymbol(g_dt_trace, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int ssf_dense_args_bytes(void) { return (int)sizeof(ssf_dense_args); }

extern "C" int ssf_dense_tc(const ssf_dense_args* args, void* stream) {
    ssf_dense_args a = *args;
    if (a.rows <= 0) return ssf_arg_error("dense_tc: empty input");
    if (a.K <= 0 || a.K % KC || a.K > 512) return ssf_arg_error("dense_tc: K must be a positive multiple of 32, <= 512");
    if (a.N < 32 || a.N % 32 || (a.N > 256 && a.N % 256)) return ssf_arg_error("dense_tc: N must be a multiple of 32 (of 256 above 256)");
    if (a.a_mode == 0) {
        if (a.x1 == nullptr || a.c1 % KC || (a.x2 != nullptr && a.c2 % KC) || a.c1 + (a.x2 ? a.c2 : 0) != a.K)
            return ssf_arg_error("dense_tc: input segments must be multiples of 32 wide and add up to K");
    } else {
        if (a.G == nullptr || a.idx == nullptr || a.S <= 0) return ssf_arg_error("dense_tc: gather mode needs G, idx, S");
    }
    if ((a.Hq != nullptr || a.Wd2 != nullptr || a.epi_mode == SSF_EPI_MAX) && a.S <= 0) return ssf_arg_error("dense_tc: S missing");
    if (a.epi_mode == SSF_EPI_MAX && a.S != 8 && a.S != 16) return ssf_arg_error("dense_tc: max epilogue needs S in {8,16}");
    if (a.epi_mode == SSF_EPI_MAX && a.rows % a.S) return ssf_arg_error("dense_tc: rows must be a multiple of S");
    if (a.epi_mode == SSF_EPI_DOT && (a.N > 256 || a.wvec == nullptr)) return ssf_arg_error("dense_tc: dot epilogue needs N <= 256 and wvec");
    if ((a.Wd1 != nullptr || a.Wd2 != nullptr) && (a.pos_src == nullptr || a.pos_q == nullptr || a.idx == nullptr))
        return ssf_arg_error("dense_tc: direction term needs pos_src, pos_q, idx");
    if (a.rows >= (1ll << 31)) return ssf_arg_error("dense_tc: at most 2^31 - 1 rows");
    DenseCfg cfg;
    dt_fastdiv_make(a.S > 0 ? (unsigned)a.S : 1u, &cfg.s_mul, &cfg.s_sh);
    dt_fastdiv_make(a.Nq > 0 ? (unsigned)a.Nq : 1u, &cfg.nq_mul, &cfg.nq_sh);
    cfg.Nt = a.N < 256 ? a.N : 256;
    cfg.nk = a.K / KC;
    const size_t wchunk = (size_t)cfg.Nt * KC * 4 * 2;
    cfg.resident = (size_t)cfg.nk * wchunk <= (size_t)W_SMEM_MAX;
    cfg.nstage = cfg.resident ? cfg.nk : (int)(W_SMEM_MAX / wchunk);
    if (cfg.nstage > 16) cfg.nstage = 16;
    cfg.n_tiles = (int)((a.rows + 127) / 128);
    const size_t smem_fixed = (size_t)cfg.nstage * wchunk + (size_t)(4 * a.K + 5 * cfg.Nt) * 4 + 80 * 8;
    const size_t tile_b = (size_t)32 * STG_LD * 4;   // one staging tile of a warp
    const int light_mode = dense_variant();
    // light_mode 1: only where it measured faster (pooled plain-row layers: the epilogue warps are the bottleneck and the
    // producers are cheap); 2: every eligible layer (tests)
    const size_t smem_light = smem_fixed + (4 * 2 + 4) * tile_b;   // prefetch depth 1 at least
    const bool light = light_mode && cfg.Nt <= 64 && cfg.resident && smem_light <= (size_t)LIGHT_SMEM_MAX &&
                       (light_mode == 2 || (a.epi_mode == SSF_EPI_MAX && a.a_mode == 0));
    // 256-column tiles (single accumulator): two epilogue warpgroups (the DOT epilogue needs whole rows in one thread)
    // ... and for plain-row layers of >= 64 columns, whose single producer warpgroup keeps up (measured: 8-25 % faster; the
    // grouped first layer needs both producer warpgroups)
    const size_t smem_full = smem_fixed + (size_t)(8 * 2 + 8) * tile_b;
    const bool full_ok = cfg.Nt >= 64 && cfg.Nt <= 128 && cfg.resident && smem_full <= (size_t)227 * 1024 && a.epi_mode != SSF_EPI_DOT;
    // plain-row layers of 64 columns with a small weight image: 2 + 2 warpgroups (measured 12-19 % faster than 1 + 2; no gain at 128)
    const bool full_r = light_mode && !light && a.a_mode == 0 && cfg.Nt == 64 && full_ok;
    const bool wide = light_mode && !light && !full_r && a.epi_mode != SSF_EPI_DOT && (cfg.Nt == 256 || (a.a_mode == 0 && cfg.Nt >= 64));
    cfg.nd = light ? 2 : ((2 * cfg.Nt + AS * 64 <= 512) ? 2 : 1);
    static unsigned long long attr_set = 0;
    if (ssf_attr_needed(&attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LIGHT_SMEM_MAX);
        if (e != cudaSuccess) return ssf_set_error(e);
        ssf_attr_done(&attr_set);
    }
    const int n_cta = light ? 2 * 148 : 148;
    dim3 grid((unsigned)(cfg.n_tiles < n_cta ? cfg.n_tiles : n_cta), (unsigned)((a.N + 255) / 256));
    // shared-memory plan: staging tiles (producer rings with 1 or 2 chunks in flight + the STORE epilogue tiles), then the
    // weight image: resident when it fits in what is left (at most W_SMEM_MAX), else a ring of whole chunks
    // (measured: the SA / SU pooling layers gain 4-8 %; layers that also add a per-point block H lose under the 96-register cap)
    const bool full = full_r || (light_mode && !light && !wide && a.a_mode == 1 && a.H == nullptr && full_ok);
    const int n_pw = (light || wide) ? 4 : 8, n_ew = (wide || full) ? 8 : 4;
    const size_t smem_cap = light ? (size_t)LIGHT_SMEM_MAX : (size_t)227 * 1024;
    const size_t smem_par = (size_t)(4 * a.K + 5 * cfg.Nt) * 4 + 80 * 8;
    size_t w_budget = smem_cap - smem_par - ((size_t)n_pw * 2 + n_ew) * tile_b;
    if (w_budget > (size_t)W_SMEM_MAX) w_budget = W_SMEM_MAX;
    cfg.resident = (size_t)cfg.nk * wchunk <= w_budget;
    cfg.nstage = cfg.resident ? cfg.nk : (int)(w_budget / wchunk);
    if (cfg.nstage > 16) cfg.nstage = 16;
    if (cfg.nstage < 1) return ssf_arg_error("dense_tc: shared-memory budget exceeded");
    const size_t smem_w = (size_t)cfg.nstage * wchunk + smem_par;
    cfg.pd = 2;
    while (cfg.pd > 1 && smem_w + ((size_t)n_pw * (cfg.pd + 1) + n_ew) * tile_b > smem_cap) --cfg.pd;
    const size_t smem = smem_w + ((size_t)n_pw * (cfg.pd + 1) + n_ew) * tile_b;
    DenseMaps maps;
    memset(&maps, 0, sizeof(maps));
    cfg.tma_a = cfg.tma_y = cfg.tma_g = 0;
    if (dense_tma()) {
        if (dense_tma() >= 2 && a.a_mode == 1 && a.Nq > 0 && a.Nsrc > 0 && a.ldG % 32 == 0 && a.offG % 4 == 0 &&
            encode_rows_map(&maps.g, a.G, (a.rows / ((long long)a.S * a.Nq)) * a.Nsrc, a.ldG, a.ldG, 1))
            cfg.tma_g = 1;
        if (a.a_mode == 0 && encode_rows_map(&maps.x1, a.x1, a.rows, a.c1, a.ld1) &&
            (a.x2 == nullptr || encode_rows_map(&maps.x2, a.x2, a.rows, a.c2, a.ld2)))
            cfg.tma_a = 1;
        if (a.epi_mode == SSF_EPI_STORE && encode_rows_map(&maps.y, a.y, a.rows, a.N, a.ldy)) cfg.tma_y = 1;
    }
    if (light) dense_tc_kernel<1><<<grid, DT_THREADS_LIGHT, smem, (cudaStream_t)stream>>>(a, cfg, maps);
    else if (full) dense_tc_kernel<3><<<grid, DT_THREADS_FULL, smem, (cudaStream_t)stream>>>(a, cfg, maps);
    else if (wide) dense_tc_kernel<2><<<grid, DT_THREADS, smem, (cudaStream_t)stream>>>(a, cfg, maps);
    else dense_tc_kernel<0><<<grid, DT_THREADS, smem, (cudaStream_t)stream>>>(a, cfg, maps);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
