// Tensor-core (tcgen05 / TMEM) realisation of the PointConvTransFlowV2 core for m = 64 (levels 0 and 1 of TFlow:
// ASF/utils/soflow.py:397-469,486; dataflow SURVEY.md Appendix F steps 3-8).  Same inputs/outputs as
// `ssf_cost_volume` (layers.cu), fp32-faithful through the 3xTF32 split (tc_common.cuh).
//
// One persistent CTA per SM walks 128-row tiles (8 query points x 16 neighbours).  576 threads:
//   warps 0-7   "forward" branch: warps 0-3 own columns 0..31 of the 128 rows (thread = row = TMEM lane), warps 4-7
//               columns 32..63
//   warps 8-15  "warped"  branch, same split (neighbours found on the warped cloud); emits Cw / gw
//   warp  16    one thread issues every tcgen05.mma (A operand = split activations in TMEM, B = weight images in smem)
//   warp  17    one thread streams the six weight images of the level through a 3-stage ring with cp.async.bulk (TMA)
// Per tile each branch runs five GEMMs (mlp_convs[1] | mlp_convs3[0] | mlp_convs3[1] | weightnet1[0] | weightnet1[3]);
// the two branches alternate on the tensor pipe so one branch's epilogue overlaps the other's MMAs, and the S x S
// attention (CUDA cores, packed fp32x2 FMAs, activations exchanged through shared memory) overlaps the mlp_convs3[0] MMAs.
// Gathers and the Cw store go through a shared-memory transpose so that every global access is a full 128-byte line; the
// gathered rows of the NEXT tile are fetched by cp.async as soon as the staging sub-tile is free.  Epilogue bias / direction /
// activation math runs on packed fp32x2 instructions.
// TMEM columns: IN_a hi|lo 0..127, IN_w hi|lo 128..255, D_a 256, D_w 320, C_a 384, C_w 448 (64 fp32 columns each).
#include "tc_common.cuh"

#ifdef SSF_CV_TRACE
__device__ long long g_cv_trace[3 * 8 * 32];
#define TRACE(who, ev) do { if (blockIdx.x == 0 && it < 8 && trace_me) g_cv_trace[((who) * 8 + it) * 32 + (ev)] = clock64(); } while (0)
#else
#define TRACE(who, ev) do {} while (0)
#endif

namespace {

constexpr int CM = 64;                      // channel width
constexpr int ROWS = 128;                   // rows per tile
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = 2 * CM * CM * 4;   // hi + lo image of one 64 x 64 layer
constexpr int LDA = CM + 4;
constexpr int LDQ = 20;
constexpr int NTHREADS = 576;
constexpr int NCHUNK = 6;                   // W2a, W2w, W3a, W3b, Wn1, Wn2 (Wn2 is 32 x 64)
// parameter block (floats): b2a[64] b2w[64] b3b[64] bn1[64] bn2[32] wn3[32] W3d[3][64] bn3
constexpr int P_B2A = 0, P_B2W = 64, P_B3B = 128, P_BN1 = 192, P_BN2 = 256, P_WN3 = 288, P_W3D = 320, P_BN3 = 512, P_TOTAL = 516;

constexpr uint32_t T_IN = 0, T_D = 256, T_C = 384;   // + 128 / 64 / 64 for the warped branch

struct CvTcArgs {
    const float* Gab; const float* Hab; const float* H3;
    const uint8_t* wblob; const float* params;
    const float* xyz1; const float* xyz2; const int* idx; const int* idxw;
    int B, N1, N2, tiles_per_cloud, n_tiles;
    unsigned tpc_mul, tpc_sh;   // tile / tiles_per_cloud == (umulhi(tile, tpc_mul) + tile) >> tpc_sh (tile < 2^31; a division is ~25 instructions, four per thread per tile)
    float* cost_fwd; float* cost_fwd_cm; float* gw; float* Cw;
};

__device__ __forceinline__ int cv_div(int tile, const CvTcArgs& a) { return (int)ssf_fastdiv((unsigned)tile, a.tpc_mul, a.tpc_sh); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ssf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ssf_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// integer round-to-nearest (ties away) to the TF32 grid == cvt.rna.tf32.f32 for finite values; lo is exact
__device__ __forceinline__ void split8(const float* v, float (&hi)[8], float (&lo)[8]) {
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
// 32 consecutive columns of this thread's row -> (hi, lo) TMEM images
// (one x32 store per image instead of four x8 measured slower: 7.55 -> 7.86 ms for both cost volumes, the 64 live split values spill)
__device__ __forceinline__ void split_store32(const float (&v)[32], uint32_t t_hi, uint32_t t_lo) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        float hi[8], lo[8];
        split8(v + c8 * 8, hi, lo);
        tc_st8(t_hi + c8 * 8, hi);
        tc_st8(t_lo + c8 * 8, lo);
    }
}
__device__ __forceinline__ float leaky(float v) { return fmaxf(v, 0.1f * v); }
// LeakyReLU(0.1) of (a + b) for two channels: packed fp32x2 add and multiply, scalar max
__device__ __forceinline__ void leaky_add2(float& a0, float& a1, float b0, float b1) {
    const float2 x = __fadd2_rn(make_float2(a0, a1), make_float2(b0, b1));
    const float2 y = __fmul2_rn(x, make_float2(0.1f, 0.1f));
    a0 = fmaxf(x.x, y.x);
    a1 = fmaxf(x.y, y.y);
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
    float a[16], b[16];
    tc_ld16(taddr, a);
    tc_ld16(taddr + 16, b);
    tc_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        v[j] = a[j];
        v[16 + j] = b[j];
    }
}

// Issue pacing measured on B200 (scripts/mma_rate.py): 42 cycles per M128 x N64 x K8 kind::tf32 MMA (A in TMEM) when the
// issuing warp runs warp-uniformly with one elected lane, ~55 in this kernel when it sat in an `if (lane == 0)` region;
// sharing or alternating accumulators makes no difference.
// D[128 x N] = A[128 x 64] . W[N x 64]^T, 3xTF32: Alo.Whi + Ahi.Wlo + Ahi.Whi  (one thread).  The shared-memory descriptors
// of the 8 K-steps differ only in the start-address field, so they are formed by one 64-bit add on a base descriptor.
template <int N>
__device__ __forceinline__ void issue_gemm(bool leader, uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi_smem) {
    constexpr uint32_t idesc = tc_idesc_tf32(128, N);
    constexpr uint32_t lbo = (uint32_t)(N / 8) * 128;
    const uint64_t d_hi = tc_smem_desc(w_hi_smem, lbo, 128);
    const uint64_t d_lo = d_hi + (uint64_t)(((uint32_t)N * CM * 4) >> 4);
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = pass == 0 ? a_lo : a_hi;
        const uint64_t w = pass == 1 ? d_lo : d_hi;
#pragma unroll
        for (int ks = 0; ks < CM / 8; ++ks) {
            const uint64_t bdesc = w + (uint64_t)((ks * 2 * lbo) >> 4);
            if (!leader) continue;   // whole warp runs the (uniform) address arithmetic, one elected lane issues
            if (pass == 0 && ks == 0)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                             "r"(a + ks * 8), "l"(bdesc), "r"(idesc)
                             : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                             "r"(a + ks * 8), "l"(bdesc), "r"(idesc)
                             : "memory");
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) cost_volume_tc64_kernel(CvTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sWst = smem;
    float* sA = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);   // [128][LDA]  forward-branch A rows
    float* sAw = sA + ROWS * LDA;                                         // [128][LDA]  warped-branch rows
    float* sQ = sAw + ROWS * LDA;                                         // [128][LDQ] attention logits / weights (2 x for alignment slack)
    float* sPar = sQ + 2 * ROWS * LDQ;                                    // [P_TOTAL]
    float* sNorm = sPar + 520;                                            // max | 1/sum per branch, [128] each
    float* sG = sNorm + 4 * ROWS;                                         // [2][128][2] half logits of weightnet1
    float* sOut = sG + 4 * ROWS;                                          // [64][8] forward cost, channel-major staging
    float* sH3 = sOut + CM * 8;                                           // [8][64] per-point block of mlp_convs3[0]
    float* sHab = sH3 + 8 * CM;                                           // [2][8][128] per-point blocks of mlp_convs/mlp_convs2[0]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sHab + 2 * 8 * 2 * CM);
    uint64_t* w_full = bars;            // [NSTAGE]
    uint64_t* w_empty = bars + NSTAGE;  // [NSTAGE]
    uint64_t* in_ready = bars + 2 * NSTAGE;      // [2]
    uint64_t* d_ready = bars + 2 * NSTAGE + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 16) tc_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            ssf_mbar_init(&w_full[i], 1);
            ssf_mbar_init(&w_empty[i], 1);
        }
        ssf_mbar_init(&in_ready[0], 256);
        ssf_mbar_init(&in_ready[1], 256);
        ssf_mbar_init(&d_ready[0], 1);
        ssf_mbar_init(&d_ready[1], 1);
        ssf_mbar_fence_init();
    }
    for (int i = tid; i < P_TOTAL; i += NTHREADS) sPar[i] = __ldg(a.params + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_my = ((int)blockIdx.x < a.n_tiles) ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 17) {
        // ------------------------------------------------------------------ weight producer (TMA bulk copies)
        if (lane == 0) {
            for (int it = 0; it < n_my; ++it) {
                for (int c = 0; c < NCHUNK; ++c) {
                    const int g = it * NCHUNK + c, st = g % NSTAGE;
                    if (g >= NSTAGE) ssf_mbar_wait(&w_empty[st], (uint32_t)((g / NSTAGE - 1) & 1));
                    const uint32_t bytes = c == NCHUNK - 1 ? STAGE_BYTES / 2 : STAGE_BYTES;
                    ssf_mbar_expect_tx(&w_full[st], bytes);
                    ssf_bulk_g2s(sWst + st * STAGE_BYTES, a.wblob + (size_t)c * STAGE_BYTES, bytes, &w_full[st]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 16) {
        // ------------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane issues)
        {
            const bool leader = tc_elect_one();
            uint32_t in_phase = 0;
            for (int it = 0; it < n_my; ++it) {
                for (int step = 0; step < 5; ++step) {
                    for (int br = 0; br < 2; ++br) {
                        const int c = step == 0 ? br : step + 1;
                        const int g = it * NCHUNK + c, st = g % NSTAGE;
                        if (step == 0 || br == 0) ssf_mbar_wait(&w_full[st], (uint32_t)((g / NSTAGE) & 1));
                        ssf_mbar_wait(&in_ready[br], in_phase);
                        tc_fence_after();
                        { const bool trace_me = lane == 0; TRACE(2, step * 4 + br * 2); }
                        const uint32_t in_hi = tmem + T_IN + br * 128;
                        const uint32_t d = tmem + (step == 2 ? T_C : T_D) + br * 64;
                        if (step == 4) issue_gemm<32>(leader, d, in_hi, in_hi + 64, ssf_smem_u32(sWst + st * STAGE_BYTES));
                        else issue_gemm<64>(leader, d, in_hi, in_hi + 64, ssf_smem_u32(sWst + st * STAGE_BYTES));
                        if (leader) {
                            tc_commit(&d_ready[br]);
                            if (step == 0 || br == 1) tc_commit(&w_empty[st]);
                        }
                        __syncwarp();
                        { const bool trace_me = lane == 0; TRACE(2, step * 4 + br * 2 + 1); }
                    }
                    in_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ branch warps (thread = row, half the columns)
        const int br = warp >> 3;             // 0 forward, 1 warped
        const int half = (warp >> 2) & 1;     // columns [32 half, 32 half + 32)
        const int quad = warp & 3;            // TMEM lane quadrant
        const int r = quad * 32 + lane, p = r >> 4, s = r & 15;
        const int c0 = half * 32;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const uint32_t t_hi = tmem + lane_base + T_IN + br * 128 + c0, t_lo = t_hi + 64;
        const uint32_t t_d = tmem + lane_base + T_D + br * 64 + c0;
        const uint32_t t_c = tmem + lane_base + T_C + br * 64 + c0;
        float* sOwn = br ? sAw : sA;
        const float* sOth = br ? sA : sAw;
        const int* idx_br = br ? a.idxw : a.idx;
        uint64_t* my_in = &in_ready[br];
        uint64_t* my_d = &d_ready[br];
        uint32_t dph = 0;
        // coalesced-transfer mapping inside a warp: 8 lanes x 16 bytes cover the 128-byte half row of one of 4 rows
        const int rg = lane >> 3, pc = lane & 7;
        float* stage_w = sOwn + (quad * 32) * LDA + c0;   // this warp's 32 rows x 32 columns of the staging tile

        const bool trace_me = (tid == 0 || tid == 256);
        // neighbour index of this thread's row, fetched one tile ahead
        auto tile_qrow = [&](int it_) -> size_t {
            const int tile_ = (int)blockIdx.x + it_ * (int)gridDim.x;
            const int b_ = cv_div(tile_, a), n_ = (tile_ - b_ * a.tiles_per_cloud) * 8 + p;
            return (size_t)b_ * a.N1 + (n_ < a.N1 ? n_ : a.N1 - 1);
        };
        int id_next = n_my > 0 ? __ldg(idx_br + tile_qrow(0) * 16 + s) : 0;
        // Gathered rows Gab[idx] of tile `it_` (this warp's 32 rows x 32 columns) -> the warp's own staging sub-tile, by
        // cp.async: issued as soon as the sub-tile is free in the previous tile, so the gather latency hides behind E2..E6
        auto gather_issue = [&](int it_, int id_) {
            const int tile_ = (int)blockIdx.x + it_ * (int)gridDim.x;
            const int b_ = cv_div(tile_, a);
            const float* gbase = a.Gab + (size_t)b_ * a.N2 * (2 * CM) + br * CM + c0 + pc * 4;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int rl = j * 4 + rg;
                const int sid = __shfl_sync(0xffffffffu, id_, rl);
                cp_async16(stage_w + rl * LDA + pc * 4, gbase + (size_t)sid * (2 * CM));
            }
            cp_async_commit();
        };
        if (n_my > 0) gather_issue(0, id_next);
        // Hab rows of the 8 points of a tile: 256 float4, one per thread of warps 0-7; tile `it` sits in buffer it & 1
        auto hab_load = [&](int it_) -> float4 {
            const int tile_ = (int)blockIdx.x + it_ * (int)gridDim.x;
            const int b_ = cv_div(tile_, a), n_ = (tile_ - b_ * a.tiles_per_cloud) * 8 + (tid >> 5);
            return __ldg(reinterpret_cast<const float4*>(a.Hab + ((size_t)b_ * a.N1 + (n_ < a.N1 ? n_ : a.N1 - 1)) * (2 * CM)) + (tid & 31));
        };
        if (tid < 256 && n_my > 0) *reinterpret_cast<float4*>(sHab + tid * 4) = hab_load(0);
        named_bar(1, 512);
        for (int it = 0; it < n_my; ++it) {
            TRACE(br, 0);
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int b = cv_div(tile, a);
            const int n0 = (tile - b * a.tiles_per_cloud) * 8;
            const bool valid = n0 + p < a.N1;
            const int n = valid ? n0 + p : a.N1 - 1;
            const size_t qrow = (size_t)b * a.N1 + n;
            const int id = id_next;
            if (it + 1 < n_my) id_next = __ldg(idx_br + tile_qrow(it + 1) * 16 + s);
            // per-point block of mlp_convs3[0] (needed in E2): 8 points x 64 floats, loaded now, parked in shared memory
            // just before the attention barrier
            float4 hab_pref = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < 256 && it + 1 < n_my) hab_pref = hab_load(it + 1);
            float4 h3_pref = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < 128) {
                const int np = n0 + (tid >> 4);
                h3_pref = __ldg(reinterpret_cast<const float4*>(a.H3 + ((size_t)b * a.N1 + (np < a.N1 ? np : a.N1 - 1)) * CM) + (tid & 15));
            }
            // direction = xyz2[idx] - xyz1[n] (un-warped xyz2 for both branches, soflow.py:407); only needed in E2, so the
            // loads are issued here and consumed there
            const float* pq_ = a.xyz1 + qrow * 3;
            const float* ps_ = a.xyz2 + ((size_t)b * a.N2 + id) * 3;
            const float psx = __ldg(ps_), psy = __ldg(ps_ + 1), psz = __ldg(ps_ + 2);
            const float pqx = __ldg(pq_), pqy = __ldg(pq_ + 1), pqz = __ldg(pq_ + 2);
            float v[32];
            // ---- prologue: x0 = leaky(Gab[idx] + Hab[n])  (first layers of mlp_convs / mlp_convs2, split algebraically)
            {
                cp_async_wait_all();   // rows gathered by gather_issue(it)
                __syncwarp();
                const float* hab = sHab + (it & 1) * (8 * 2 * CM) + p * (2 * CM) + br * CM + c0;
                const float* mine = sOwn + r * LDA + c0;
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 g4 = lds4(mine + q4 * 4);
                    const float4 h = lds4(hab + q4 * 4);
                    v[q4 * 4] = leaky(g4.x + h.x);
                    v[q4 * 4 + 1] = leaky(g4.y + h.y);
                    v[q4 * 4 + 2] = leaky(g4.z + h.z);
                    v[q4 * 4 + 3] = leaky(g4.w + h.w);
                }
                split_store32(v, t_hi, t_lo);
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(my_in);
                __syncwarp();   // staging rows are rewritten below by their owner threads
                TRACE(br, 1);
            }
            // ---- E1: A = leaky(D + b2): own half row to shared memory, split back to TMEM for mlp_convs3[0]
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            TRACE(br, 2);
            {
                const float* b2 = sPar + (br ? P_B2W : P_B2A) + c0;
                ld32(t_d, v);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 bb = lds4(b2 + q4 * 4);
                    leaky_add2(v[q4 * 4], v[q4 * 4 + 1], bb.x, bb.y);
                    leaky_add2(v[q4 * 4 + 2], v[q4 * 4 + 3], bb.z, bb.w);
                    *reinterpret_cast<float4*>(sOwn + r * LDA + c0 + q4 * 4) = make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
                }
                split_store32(v, t_hi, t_lo);
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(my_in);
            }
            if (tid < 128) *reinterpret_cast<float4*>(sH3 + tid * 4) = h3_pref;
            if (tid < 256) *reinterpret_cast<float4*>(sHab + ((it + 1) & 1) * (8 * 2 * CM) + tid * 4) = hab_pref;
            TRACE(br, 3);
            named_bar(1, 512);   // sA and sAw complete
            TRACE(br, 4);
            // ---- attention (soflow.py:420-422,453-458) on the CUDA cores, register-blocked so that every value read from
            // shared memory feeds >= 2 packed FMAs (the shared-memory -> register path, 128 B/clk, is the limiter here).
            // The mapping is per point, not per row: 64 threads per point.
            {
                const int ct = tid;                       // 0..511
                {   // Q[i][j] = <A_i, Aw_j>: thread = 4 x 4 block of one point over a 16-channel slice
                    const int pp = ct >> 6, tp = ct & 63, ks = tp & 3, bi = (tp >> 2) >> 2, bj = (tp >> 2) & 3;
                    // slice ks owns the channel quads ks, ks + 4, ks + 8, ks + 12: the 8 lanes of a quarter-warp then hit 8
                    // different 16-byte bank groups (or identical addresses), i.e. no shared-memory conflicts
                    const float* ap = sA + (pp * 16 + bi * 4) * LDA + ks * 4;
                    const float* wp = sAw + (pp * 16 + bj * 4) * LDA + ks * 4;
                    float2 acc[4][4];
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[x][y] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        float4 ar[4], wr[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            ar[x] = lds4(ap + x * LDA + c4 * 16);
                            wr[x] = lds4(wp + x * LDA + c4 * 16);
                        }
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y) {
                                acc[x][y] = __ffma2_rn(make_float2(ar[x].x, ar[x].y), make_float2(wr[y].x, wr[y].y), acc[x][y]);
                                acc[x][y] = __ffma2_rn(make_float2(ar[x].z, ar[x].w), make_float2(wr[y].z, wr[y].w), acc[x][y]);
                            }
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        float qv[4];
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            float t = acc[x][y].x + acc[x][y].y;
                            t += __shfl_xor_sync(0xffffffffu, t, 1);
                            t += __shfl_xor_sync(0xffffffffu, t, 2);
                            qv[y] = t;
                        }
                        if (ks == 0) *reinterpret_cast<float4*>(sQ + (pp * 16 + bi * 4 + x) * LDQ + bj * 4) = make_float4(qv[0], qv[1], qv[2], qv[3]);
                    }
                }
                TRACE(br, 14);
                named_bar(1, 512);
                TRACE(br, 15);
                if (ct < 256) {   // softmax statistics: threads 0..127 rows (over j, dim -1), 128..255 columns (over i, dim -2)
                    const int which = ct >> 7, e = ct & 127, pp = e >> 4, k = e & 15;
                    float t[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) t[u] = which == 0 ? sQ[e * LDQ + u] : sQ[(pp * 16 + u) * LDQ + k];
                    float mx = t[0];
#pragma unroll
                    for (int u = 1; u < 16; ++u) mx = fmaxf(mx, t[u]);
                    float sm = 0.f;
#pragma unroll
                    for (int u = 0; u < 16; ++u) sm += __expf(t[u] - mx);
                    sNorm[which * 256 + e] = mx;
                    sNorm[which * 256 + 128 + e] = 1.0f / sm;
                }
                TRACE(br, 16);
                named_bar(1, 512);
                {   // Q <- softmax_j(Q) * softmax_i(Q): four entries of one row per thread
                    const int row = ct >> 2, j0 = (ct & 3) * 4, pp = row >> 4;
                    float4 q4 = lds4(sQ + row * LDQ + j0);
                    const float rmx = sNorm[row], rinv = sNorm[128 + row];
                    const float4 cmx = lds4(sNorm + 256 + pp * 16 + j0), cinv = lds4(sNorm + 384 + pp * 16 + j0);
                    q4.x = (__expf(q4.x - rmx) * rinv) * (__expf(q4.x - cmx.x) * cinv.x);
                    q4.y = (__expf(q4.y - rmx) * rinv) * (__expf(q4.y - cmx.y) * cinv.y);
                    q4.z = (__expf(q4.z - rmx) * rinv) * (__expf(q4.z - cmx.z) * cinv.z);
                    q4.w = (__expf(q4.w - rmx) * rinv) * (__expf(q4.w - cmx.w) * cinv.w);
                    *reinterpret_cast<float4*>(sQ + row * LDQ + j0) = q4;
                }
                named_bar(1, 512);
                TRACE(br, 17);
                // mixes: threads 0..255 A' = A + Q.Aw (8 rows x 4 channels each), threads 256..511 Aw' = Aw + Q^T.A
                const int mt = ct & 255, pp = mt >> 5, blk = (mt >> 4) & 1, cq = mt & 15;
                float4 macc[8];
                {
                    const float* self = (ct < 256 ? sA : sAw) + (pp * 16 + blk * 8) * LDA + cq * 4;
                    const float* other = (ct < 256 ? sAw : sA) + (pp * 16) * LDA + cq * 4;
#pragma unroll
                    for (int k = 0; k < 8; ++k) macc[k] = lds4(self + k * LDA);
                    if (ct < 256) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            float4 o[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) o[u] = lds4(other + (j4 * 4 + u) * LDA);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const float4 q4 = lds4(sQ + (pp * 16 + blk * 8 + k) * LDQ + j4 * 4);   // Q[i = 8 blk + k][4 j4 ..]
                                const float qq[4] = {q4.x, q4.y, q4.z, q4.w};
                                float2 lo2 = make_float2(macc[k].x, macc[k].y), hi2 = make_float2(macc[k].z, macc[k].w);
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    lo2 = __ffma2_rn(make_float2(qq[u], qq[u]), make_float2(o[u].x, o[u].y), lo2);
                                    hi2 = __ffma2_rn(make_float2(qq[u], qq[u]), make_float2(o[u].z, o[u].w), hi2);
                                }
                                macc[k] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
                            }
                        }
                    } else {
#pragma unroll 4
                        for (int i = 0; i < 16; ++i) {
                            const float4 o = lds4(other + i * LDA);
                            const float4 qa = lds4(sQ + (pp * 16 + i) * LDQ + blk * 8), qb = lds4(sQ + (pp * 16 + i) * LDQ + blk * 8 + 4);
                            const float qq[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};   // Q[i][j = 8 blk + k]
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const float2 lo2 = __ffma2_rn(make_float2(qq[k], qq[k]), make_float2(o.x, o.y), make_float2(macc[k].x, macc[k].y));
                                const float2 hi2 = __ffma2_rn(make_float2(qq[k], qq[k]), make_float2(o.z, o.w), make_float2(macc[k].z, macc[k].w));
                                macc[k] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
                            }
                        }
                    }
                }
                TRACE(br, 18);
                named_bar(1, 512);   // every read of the un-mixed rows is done: overwrite them in place
                TRACE(br, 19);
                {
                    float* self = (ct < 256 ? sA : sAw) + (pp * 16 + blk * 8) * LDA + cq * 4;
#pragma unroll
                    for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(self + k * LDA) = macc[k];
                }
                named_bar(1, 512);
                // back to thread = row: this thread's 32 columns of the mixed row (A' or Aw')
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 t = lds4(sOwn + r * LDA + c0 + q4 * 4);
                    v[q4 * 4] = t.x; v[q4 * 4 + 1] = t.y; v[q4 * 4 + 2] = t.z; v[q4 * 4 + 3] = t.w;
                }
            }
            named_bar(1, 512);   // mixed rows consumed (sOwn is reused as a staging tile below and by the next tile)
            TRACE(br, 5);
            if (br == 0 && it + 1 < n_my) gather_issue(it + 1, id_next);   // the forward branch does not touch sA again in this tile
            // v[] now holds this thread's 32 columns of the mixed row (A' or Aw'); it stays in registers through E2,
            // which therefore works on 16 columns at a time
            // ---- E2: C1 = leaky(D + H3[n] + W3d . dir)   (first layer of mlp_convs3; A block came from the MMA)
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            TRACE(br, 6);
            {
                const float* h3 = sH3 + p * CM + c0;
                const float* wd = sPar + P_W3D + c0;
                const float dx = psx - pqx, dy = psy - pqy, dz = psz - pqz;
                const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float t[16];
                    tc_ld16(t_d + hh * 16, t);
                    tc_ld_wait();
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int cq = hh * 4 + q4;
                        const float4 h = lds4(h3 + cq * 4);
                        const float4 w0 = lds4(wd + cq * 4), w1 = lds4(wd + CM + cq * 4), w2 = lds4(wd + 2 * CM + cq * 4);
                        // D + H3[n] + W3d . dir on packed fp32x2 FMAs (two channels per instruction), then LeakyReLU
                        float2 a01 = __fadd2_rn(make_float2(t[q4 * 4], t[q4 * 4 + 1]), make_float2(h.x, h.y));
                        float2 a23 = __fadd2_rn(make_float2(t[q4 * 4 + 2], t[q4 * 4 + 3]), make_float2(h.z, h.w));
                        a01 = __ffma2_rn(dx2, make_float2(w0.x, w0.y), a01);
                        a23 = __ffma2_rn(dx2, make_float2(w0.z, w0.w), a23);
                        a01 = __ffma2_rn(dy2, make_float2(w1.x, w1.y), a01);
                        a23 = __ffma2_rn(dy2, make_float2(w1.z, w1.w), a23);
                        a01 = __ffma2_rn(dz2, make_float2(w2.x, w2.y), a01);
                        a23 = __ffma2_rn(dz2, make_float2(w2.z, w2.w), a23);
                        const float2 s01 = __fmul2_rn(a01, make_float2(0.1f, 0.1f)), s23 = __fmul2_rn(a23, make_float2(0.1f, 0.1f));
                        t[q4 * 4] = fmaxf(a01.x, s01.x);
                        t[q4 * 4 + 1] = fmaxf(a01.y, s01.y);
                        t[q4 * 4 + 2] = fmaxf(a23.x, s23.x);
                        t[q4 * 4 + 3] = fmaxf(a23.y, s23.y);
                    }
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) {
                        float hi[8], lo[8];
                        split8(t + c8 * 8, hi, lo);
                        tc_st8(t_hi + hh * 16 + c8 * 8, hi);
                        tc_st8(t_lo + hh * 16 + c8 * 8, lo);
                    }
                }
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(my_in);
            }
            TRACE(br, 7);
            // ---- E3: mlp_convs3[1] result stays in its own TMEM buffer; feed the mixed rows to weightnet1[0]
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            TRACE(br, 8);
            split_store32(v, t_hi, t_lo);
            tc_st_wait();
            tc_fence_before();
            mbar_arrive(my_in);
            if (br == 1) {   // warped-branch cost rows go to HBM for the segmented softmax/sum (soflow.py:471-481)
                const float* b3 = sPar + P_B3B + c0;
                ld32(t_c, v);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 bb = lds4(b3 + q4 * 4);
                    *reinterpret_cast<float4*>(sOwn + r * LDA + c0 + q4 * 4) =
                        make_float4(leaky(v[q4 * 4] + bb.x), leaky(v[q4 * 4 + 1] + bb.y), leaky(v[q4 * 4 + 2] + bb.z), leaky(v[q4 * 4 + 3] + bb.w));
                }
                __syncwarp();
                // this warp's 32 rows x 128 bytes, four rows per store instruction
                const int row0 = quad * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rl = j * 4 + rg;
                    const int rr = row0 + rl;
                    const int nn = n0 + (rr >> 4);
                    if (nn < a.N1) {
                        const float4 t = lds4(stage_w + rl * LDA + pc * 4);
                        __stcs(reinterpret_cast<float4*>(a.Cw + (((size_t)b * a.N1 + nn) * 16 + (rr & 15)) * CM + c0 + pc * 4), t);
                    }
                }
                __syncwarp();
                if (it + 1 < n_my) gather_issue(it + 1, id_next);   // staging sub-tile free again
            }
            TRACE(br, 9);
            // ---- E4: T1 = relu(D + bn1)
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            TRACE(br, 10);
            {
                const float* bn = sPar + P_BN1 + c0;
                ld32(t_d, v);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 bb = lds4(bn + q4 * 4);
                    const float2 r01 = __fadd2_rn(make_float2(v[q4 * 4], v[q4 * 4 + 1]), make_float2(bb.x, bb.y));
                    const float2 r23 = __fadd2_rn(make_float2(v[q4 * 4 + 2], v[q4 * 4 + 3]), make_float2(bb.z, bb.w));
                    v[q4 * 4] = fmaxf(r01.x, 0.f);
                    v[q4 * 4 + 1] = fmaxf(r01.y, 0.f);
                    v[q4 * 4 + 2] = fmaxf(r23.x, 0.f);
                    v[q4 * 4 + 3] = fmaxf(r23.y, 0.f);
                }
                split_store32(v, t_hi, t_lo);
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(my_in);
            }
            TRACE(br, 11);
            // ---- E5: logit = wn3 . relu(D[0:32] + bn2) + bn3   (each half sums 16 of the 32 columns)
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            TRACE(br, 12);
            float g;
            {
                float t[16];
                tc_ld16(tmem + lane_base + T_D + br * 64 + half * 16, t);
                tc_ld_wait();
                float part = 0.f;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 bb = lds4(sPar + P_BN2 + half * 16 + q4 * 4), w3 = lds4(sPar + P_WN3 + half * 16 + q4 * 4);
                    part = fmaf(fmaxf(t[q4 * 4] + bb.x, 0.f), w3.x, part);
                    part = fmaf(fmaxf(t[q4 * 4 + 1] + bb.y, 0.f), w3.y, part);
                    part = fmaf(fmaxf(t[q4 * 4 + 2] + bb.z, 0.f), w3.z, part);
                    part = fmaf(fmaxf(t[q4 * 4 + 3] + bb.w, 0.f), w3.w, part);
                }
                sG[(br * ROWS + r) * 2 + half] = part;
                if (br) named_bar(3, 256); else named_bar(2, 256);   // constant ids: ptxas then reserves 4 hardware barriers, not all 16
                g = (sG[(br * ROWS + r) * 2] + sG[(br * ROWS + r) * 2 + 1]) + sPar[P_BN3];
            }
            TRACE(br, 13);
            if (br == 1) {
                if (valid && half == 0) a.gw[qrow * 16 + s] = g;
            } else {
                // ---- E6: forward cost = sum_s softmax_s(g) * C[s]   (soflow.py:469,486)
                float mx = g;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float e = __expf(g - mx);
                float sm = e;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
                const float wgt = e / sm;
                const float* b3 = sPar + P_B3B + c0;
                ld32(t_c, v);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 bb = lds4(b3 + q4 * 4);
                    v[q4 * 4] = wgt * leaky(v[q4 * 4] + bb.x);
                    v[q4 * 4 + 1] = wgt * leaky(v[q4 * 4 + 1] + bb.y);
                    v[q4 * 4 + 2] = wgt * leaky(v[q4 * 4 + 2] + bb.z);
                    v[q4 * 4 + 3] = wgt * leaky(v[q4 * 4 + 3] + bb.w);
                }
                // butterfly sum over the 16 rows of the point: lane s ends up with columns c0 + 2s, c0 + 2s + 1
#pragma unroll
                for (int w = 16; w >= 2; w >>= 1) {          // w = values kept after this step
                    const bool up = (lane & (w >> 1)) != 0;  // xor distance = w/2: 8, 4, 2, 1
#pragma unroll
                    for (int j = 0; j < w; ++j) {
                        const float send = up ? v[j] : v[j + w];
                        const float keep = up ? v[j + w] : v[j];
                        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, w >> 1);
                    }
                }
                const int cc = c0 + 2 * s;
                if (valid) *reinterpret_cast<float2*>(a.cost_fwd + qrow * CM + cc) = make_float2(v[0], v[1]);
                sOut[cc * 8 + p] = v[0];
                sOut[(cc + 1) * 8 + p] = v[1];
                if (br) named_bar(3, 256); else named_bar(2, 256);   // constant ids: ptxas then reserves 4 hardware barriers, not all 16
                {   // channel-major copy: 2 consecutive points of one channel per thread
                    const int t = half * 128 + r;
                    const int c = t >> 2, part = t & 3;
                    const int nn = n0 + part * 2;
                    float* dst = a.cost_fwd_cm + ((size_t)b * CM + c) * a.N1 + nn;
                    const float* src = sOut + c * 8 + part * 2;
                    if ((a.N1 & 1) == 0 && nn + 1 < a.N1) {
                        *reinterpret_cast<float2*>(dst) = make_float2(src[0], src[1]);
                    } else {
                        if (nn < a.N1) dst[0] = src[0];
                        if (nn + 1 < a.N1) dst[1] = src[1];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tc_dealloc(tmem, 512);
}

constexpr size_t CV_TC_SMEM =
    (size_t)NSTAGE * STAGE_BYTES + (size_t)(2 * ROWS * LDA + 2 * ROWS * LDQ + 520 + 4 * ROWS + 4 * ROWS + CM * 8 + 8 * CM + 2 * 8 * 2 * CM) * 4 + 16 * 8;

}  // namespace

#ifdef SSF_CV_TRACE
extern "C" int ssf_cv_trace_read(long long* out) {
    return cudaMemcpyFromSymbol(out, g_cv_trace, sizeof(long long) * 3 * 8 * 32) == cudaSuccess ? 0 : 2;
}
#endif

extern "C" long long ssf_cost_volume_tc_blob_bytes() { return (long long)(5 * STAGE_BYTES + STAGE_BYTES / 2); }
extern "C" int ssf_cost_volume_tc_param_floats() { return P_TOTAL; }

extern "C" int ssf_cost_volume_tc(const float* Gab, const float* Hab, const float* H3, const void* wblob, const float* params,
                                  const float* xyz1, const float* xyz2, const int* idx, const int* idxw, int B, int N1, int N2,
                                  int m, float* cost_fwd, float* cost_fwd_cm, float* gw, float* Cw, int n_sm, void* stream) {
    if (B <= 0 || N1 <= 0) return ssf_arg_error("cost_volume_tc: empty input");
    if (m != CM) return ssf_arg_error("cost_volume_tc: m must be 64");
    CvTcArgs a;
    a.Gab = Gab; a.Hab = Hab; a.H3 = H3; a.wblob = static_cast<const uint8_t*>(wblob); a.params = params;
    a.xyz1 = xyz1; a.xyz2 = xyz2; a.idx = idx; a.idxw = idxw;
    a.B = B; a.N1 = N1; a.N2 = N2;
    a.tiles_per_cloud = (N1 + 7) / 8;
    {
        const SsfFastDiv f = ssf_fastdiv_make((unsigned)a.tiles_per_cloud);
        a.tpc_mul = f.mul;
        a.tpc_sh = f.sh;
    }
    a.n_tiles = a.tiles_per_cloud * B;
    a.cost_fwd = cost_fwd; a.cost_fwd_cm = cost_fwd_cm; a.gw = gw; a.Cw = Cw;
    cudaError_t e = cudaFuncSetAttribute(cost_volume_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CV_TC_SMEM);
    if (e != cudaSuccess) return ssf_set_error(e);
    if (n_sm <= 0) n_sm = 148;
    const int grid = a.n_tiles < n_sm ? a.n_tiles : n_sm;
    cost_volume_tc64_kernel<<<grid, NTHREADS, CV_TC_SMEM, (cudaStream_t)stream>>>(a);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
