// Tensor-core (tcgen05 / TMEM) realisation of the PointConvTransFlowV2 core for m = 64 (levels 0 and 1 of TFlow:
// ASF/utils/soflow.py:397-469,486; dataflow SURVEY.md Appendix F steps 3-8).  Same inputs/outputs as
// `ssf_cost_volume` (layers.cu), fp32-faithful through the 3xTF32 split (tc_common.cuh).
//
// One persistent CTA per SM walks 128-row tiles (8 query points x 16 neighbours).  Roles:
//   warps 0-3  "forward" branch  (thread = row): gather-add prologue, layer epilogues, attention, forward cost
//   warps 4-7  "warped"  branch  (thread = row): same for the warped-cloud neighbours, emits Cw / gw
//   warp  8    one thread issues every tcgen05.mma (A operand = split activations in TMEM, B = weight images in smem)
//   warp  9    one thread streams the six weight images of the level through a 3-stage ring with cp.async.bulk (TMA)
// Per tile each branch runs five GEMMs (mlp_convs[1] | mlp_convs3[0] | mlp_convs3[1] | weightnet1[0] | weightnet1[3]);
// the two branches alternate on the tensor pipe so one branch's epilogue overlaps the other's MMAs, and the S x S
// attention (CUDA cores, activations exchanged through shared memory) overlaps the mlp_convs3[0] MMAs.
// TMEM columns: IN_a hi|lo 0..127, IN_w hi|lo 128..255, D_a 256, D_w 320, C_a 384, C_w 448 (64 fp32 columns each).
#include "tc_common.cuh"

namespace {

constexpr int CM = 64;                      // channel width
constexpr int ROWS = 128;                   // rows per tile
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = 2 * CM * CM * 4;   // hi + lo image of one 64 x 64 layer
constexpr int LDA = CM + 4;
constexpr int NTHREADS = 320;
constexpr int NCHUNK = 6;                   // W2a, W2w, W3a, W3b, Wn1, Wn2 (Wn2 is 32 x 64)
// parameter block (floats): b2a[64] b2w[64] b3b[64] bn1[64] bn2[32] wn3[32] W3d[3][64] bn3
constexpr int P_B2A = 0, P_B2W = 64, P_B3B = 128, P_BN1 = 192, P_BN2 = 256, P_WN3 = 288, P_W3D = 320, P_BN3 = 512, P_TOTAL = 516;

constexpr uint32_t T_IN = 0, T_D = 256, T_C = 384;   // + 128 / 64 / 64 for the warped branch

struct CvTcArgs {
    const float* Gab; const float* Hab; const float* H3;
    const uint8_t* wblob; const float* params;
    const float* xyz1; const float* xyz2; const int* idx; const int* idxw;
    int B, N1, N2, tiles_per_cloud, n_tiles;
    float* cost_fwd; float* cost_fwd_cm; float* gw; float* Cw;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ssf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// integer round-to-nearest (ties away) to the TF32 grid == cvt.rna.tf32.f32 for finite values; lo is exact
__device__ __forceinline__ void split8(const float (&v)[8], float (&hi)[8], float (&lo)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        hi[j] = __uint_as_float((__float_as_uint(v[j]) + 0x1000u) & 0xFFFFE000u);
        lo[j] = v[j] - hi[j];
    }
}
__device__ __forceinline__ float leaky(float v) { return fmaxf(v, 0.1f * v); }
// 8 consecutive floats from 16-byte aligned shared / global memory
__device__ __forceinline__ void ld8s(const float* p, float (&o)[8]) {
    const float4 x = *reinterpret_cast<const float4*>(p), y = *reinterpret_cast<const float4*>(p + 4);
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
}

// D[128 x N] = A[128 x 64] . W[N x 64]^T, 3xTF32: Alo.Whi + Ahi.Wlo + Ahi.Whi  (one thread)
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi_smem, int N) {
    const uint32_t idesc = tc_idesc_tf32(128, N);
    const uint32_t lbo = (uint32_t)(N / 8) * 128;
    const uint32_t w_lo_smem = w_hi_smem + (uint32_t)N * CM * 4;
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = pass == 0 ? a_lo : a_hi;
        const uint32_t w = pass == 1 ? w_lo_smem : w_hi_smem;
#pragma unroll
        for (int ks = 0; ks < CM / 8; ++ks) {
            tc_mma_ts(d_tmem, a + ks * 8, tc_smem_desc(w + ks * 2 * lbo, lbo, 128), idesc, acc);
            acc = 1;
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) cost_volume_tc64_kernel(CvTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sWst = smem;
    float* sA = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);   // [128][LDA]  forward-branch A rows
    float* sAw = sA + ROWS * LDA;                                         // [128][LDA]  warped-branch rows
    float* sPar = sAw + ROWS * LDA;                                       // [P_TOTAL]
    float* sNorm = sPar + 520;                                            // rmax | rsum | cmax | csum, [128] each
    float* sOut = sNorm + 4 * ROWS;                                       // [64][8] forward cost, channel-major staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + CM * 8);
    uint64_t* w_full = bars;            // [NSTAGE]
    uint64_t* w_empty = bars + NSTAGE;  // [NSTAGE]
    uint64_t* in_ready = bars + 2 * NSTAGE;      // [2]
    uint64_t* d_ready = bars + 2 * NSTAGE + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 8) tc_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            ssf_mbar_init(&w_full[i], 1);
            ssf_mbar_init(&w_empty[i], 1);
        }
        ssf_mbar_init(&in_ready[0], 128);
        ssf_mbar_init(&in_ready[1], 128);
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

    if (warp == 9) {
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
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t in_phase = 0;
            for (int it = 0; it < n_my; ++it) {
                for (int step = 0; step < 5; ++step) {
                    for (int br = 0; br < 2; ++br) {
                        const int c = step == 0 ? br : step + 1;
                        const int g = it * NCHUNK + c, st = g % NSTAGE;
                        if (step == 0 || br == 0) ssf_mbar_wait(&w_full[st], (uint32_t)((g / NSTAGE) & 1));
                        ssf_mbar_wait(&in_ready[br], in_phase);
                        tc_fence_after();
                        const uint32_t in_hi = tmem + T_IN + br * 128;
                        const uint32_t d = tmem + (step == 2 ? T_C : T_D) + br * 64;
                        issue_gemm(d, in_hi, in_hi + 64, ssf_smem_u32(sWst + st * STAGE_BYTES), step == 4 ? 32 : 64);
                        tc_commit(&d_ready[br]);
                        if (step == 0 || br == 1) tc_commit(&w_empty[st]);
                    }
                    in_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ branch warpgroups (thread = row)
        const int wg = warp >> 2;
        const int r = tid & 127, p = r >> 4, s = r & 15;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t t_hi = tmem + lane_base + T_IN + wg * 128, t_lo = t_hi + 64;
        const uint32_t t_d = tmem + lane_base + T_D + wg * 64;
        const uint32_t t_c = tmem + lane_base + T_C + wg * 64;
        float* sOwn = wg ? sAw : sA;
        const float* sOth = wg ? sA : sAw;
        const int* idx_br = wg ? a.idxw : a.idx;
        uint64_t* my_in = &in_ready[wg];
        uint64_t* my_d = &d_ready[wg];
        uint32_t dph = 0;
        float hi[8], lo[8];

        for (int it = 0; it < n_my; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int b = tile / a.tiles_per_cloud;
            const int n0 = (tile % a.tiles_per_cloud) * 8;
            const bool valid = n0 + p < a.N1;
            const int n = valid ? n0 + p : a.N1 - 1;
            const size_t qrow = (size_t)b * a.N1 + n;
            const int id = __ldg(idx_br + qrow * 16 + s);
            float dx, dy, dz;
            {
                const float* pq = a.xyz1 + qrow * 3;
                const float* ps = a.xyz2 + ((size_t)b * a.N2 + id) * 3;   // un-warped xyz2 for both branches (soflow.py:407)
                dx = __ldg(ps) - __ldg(pq);
                dy = __ldg(ps + 1) - __ldg(pq + 1);
                dz = __ldg(ps + 2) - __ldg(pq + 2);
            }
            // ---- prologue: x0 = leaky(Gab[idx] + Hab[n])  (first layers of mlp_convs / mlp_convs2, split algebraically)
            {
                const float4* g4 = reinterpret_cast<const float4*>(a.Gab + ((size_t)b * a.N2 + id) * (2 * CM) + wg * CM);
                const float4* h4 = reinterpret_cast<const float4*>(a.Hab + qrow * (2 * CM) + wg * CM);
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const float4 g0 = __ldg(g4 + 2 * c8), g1 = __ldg(g4 + 2 * c8 + 1);
                    const float4 h0 = __ldg(h4 + 2 * c8), h1 = __ldg(h4 + 2 * c8 + 1);
                    const float v[8] = {leaky(g0.x + h0.x), leaky(g0.y + h0.y), leaky(g0.z + h0.z), leaky(g0.w + h0.w),
                                        leaky(g1.x + h1.x), leaky(g1.y + h1.y), leaky(g1.z + h1.z), leaky(g1.w + h1.w)};
                    split8(v, hi, lo);
                    tc_st8(t_hi + c8 * 8, hi);
                    tc_st8(t_lo + c8 * 8, lo);
                }
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(my_in);
            }
            // ---- E1: A = leaky(D + b2): own row to registers + shared memory, split back to TMEM for mlp_convs3[0]
            float av[64];
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            {
                const float* b2 = sPar + (wg ? P_B2W : P_B2A);
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    float v[8], bb[8];
                    tc_ld8(t_d + c8 * 8, v);
                    ld8s(b2 + c8 * 8, bb);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[j] = leaky(v[j] + bb[j]);
                        av[c8 * 8 + j] = v[j];
                    }
                    *reinterpret_cast<float4*>(sOwn + r * LDA + c8 * 8) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(sOwn + r * LDA + c8 * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                    split8(v, hi, lo);
                    tc_st8(t_hi + c8 * 8, hi);
                    tc_st8(t_lo + c8 * 8, lo);
                }
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(my_in);
            }
            named_bar(1, 256);   // sA and sAw complete
            // ---- attention (soflow.py:420-422,453-458).  Forward thread (p,i) holds row i of Q = <A_i, Aw_j>;
            // warped thread (p,j) holds column j.  Both evaluate the same fma chain, so the values agree bit for bit.
            float q[16];
            {
                const float* oth = sOth + (p * 16) * LDA;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float acc = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < 16; ++c4) {
                        const float4 o = *reinterpret_cast<const float4*>(oth + j * LDA + c4 * 4);
                        acc = fmaf(av[c4 * 4], o.x, acc);
                        acc = fmaf(av[c4 * 4 + 1], o.y, acc);
                        acc = fmaf(av[c4 * 4 + 2], o.z, acc);
                        acc = fmaf(av[c4 * 4 + 3], o.w, acc);
                    }
                    q[j] = acc;
                }
                float mx = q[0];
#pragma unroll
                for (int j = 1; j < 16; ++j) mx = fmaxf(mx, q[j]);
                float sm = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) sm += expf(q[j] - mx);
                sNorm[(wg * 2) * ROWS + r] = mx;        // forward: softmax over j (dim -1); warped: softmax over i (dim -2)
                sNorm[(wg * 2 + 1) * ROWS + r] = sm;
                named_bar(1, 256);
                const float* omx = sNorm + ((1 - wg) * 2) * ROWS + p * 16;
                const float* osm = omx + ROWS;
#pragma unroll
                for (int j = 0; j < 16; ++j) q[j] = (expf(q[j] - mx) / sm) * (expf(q[j] - omx[j]) / osm[j]);
                // mixed = own + sum_j Q * other-branch row j
#pragma unroll
                for (int j = 0; j < 16; ++j) {
#pragma unroll
                    for (int c4 = 0; c4 < 16; ++c4) {
                        const float4 o = *reinterpret_cast<const float4*>(oth + j * LDA + c4 * 4);
                        av[c4 * 4] = fmaf(q[j], o.x, av[c4 * 4]);
                        av[c4 * 4 + 1] = fmaf(q[j], o.y, av[c4 * 4 + 1]);
                        av[c4 * 4 + 2] = fmaf(q[j], o.z, av[c4 * 4 + 2]);
                        av[c4 * 4 + 3] = fmaf(q[j], o.w, av[c4 * 4 + 3]);
                    }
                }
            }
            named_bar(1, 256);   // all reads of sA / sAw / sNorm done
            // ---- E2: C1 = leaky(D + H3[n] + W3d . dir)   (first layer of mlp_convs3; A block came from the MMA)
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            {
                const float4* h4 = reinterpret_cast<const float4*>(a.H3 + qrow * CM);
                const float* wd = sPar + P_W3D;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    float v[8];
                    tc_ld8(t_d + c8 * 8, v);
                    const float4 h0 = __ldg(h4 + 2 * c8), h1 = __ldg(h4 + 2 * c8 + 1);
                    const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                    float w0[8], w1[8], w2[8];
                    ld8s(wd + c8 * 8, w0);
                    ld8s(wd + CM + c8 * 8, w1);
                    ld8s(wd + 2 * CM + c8 * 8, w2);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = leaky(v[j] + hh[j] + (dx * w0[j] + dy * w1[j] + dz * w2[j]));
                    split8(v, hi, lo);
                    tc_st8(t_hi + c8 * 8, hi);
                    tc_st8(t_lo + c8 * 8, lo);
                }
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(my_in);
            }
            // ---- E3: mlp_convs3[1] result stays in its own TMEM buffer; feed the mixed rows to weightnet1[0]
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                const float v[8] = {av[c8 * 8], av[c8 * 8 + 1], av[c8 * 8 + 2], av[c8 * 8 + 3],
                                    av[c8 * 8 + 4], av[c8 * 8 + 5], av[c8 * 8 + 6], av[c8 * 8 + 7]};
                split8(v, hi, lo);
                tc_st8(t_hi + c8 * 8, hi);
                tc_st8(t_lo + c8 * 8, lo);
            }
            tc_st_wait();
            tc_fence_before();
            mbar_arrive(my_in);
            if (wg == 1) {   // warped-branch cost rows go to HBM for the segmented softmax/sum (soflow.py:471-481)
                float* dst = a.Cw + (qrow * 16 + s) * CM;
                const float* b3 = sPar + P_B3B;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    float v[8], bb[8];
                    tc_ld8(t_c + c8 * 8, v);
                    ld8s(b3 + c8 * 8, bb);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = leaky(v[j] + bb[j]);
                    if (valid) {
                        __stcs(reinterpret_cast<float4*>(dst + c8 * 8), make_float4(v[0], v[1], v[2], v[3]));
                        __stcs(reinterpret_cast<float4*>(dst + c8 * 8 + 4), make_float4(v[4], v[5], v[6], v[7]));
                    }
                }
            }
            // ---- E4: T1 = relu(D + bn1)
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            {
                const float* bn = sPar + P_BN1;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    float v[8], bb[8];
                    tc_ld8(t_d + c8 * 8, v);
                    ld8s(bn + c8 * 8, bb);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + bb[j], 0.f);
                    split8(v, hi, lo);
                    tc_st8(t_hi + c8 * 8, hi);
                    tc_st8(t_lo + c8 * 8, lo);
                }
                tc_st_wait();
                tc_fence_before();
                mbar_arrive(my_in);
            }
            // ---- E5: logit = wn3 . relu(D[0:32] + bn2) + bn3
            ssf_mbar_wait(my_d, dph); dph ^= 1;
            tc_fence_after();
            float g = 0.f;
            {
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    float v[8], bb[8], w3[8];
                    tc_ld8(t_d + c8 * 8, v);
                    ld8s(sPar + P_BN2 + c8 * 8, bb);
                    ld8s(sPar + P_WN3 + c8 * 8, w3);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) g = fmaf(fmaxf(v[j] + bb[j], 0.f), w3[j], g);
                }
                g += sPar[P_BN3];
            }
            if (wg == 1) {
                if (valid) a.gw[qrow * 16 + s] = g;
            } else {
                // ---- E6: forward cost = sum_s softmax_s(g) * C[s]   (soflow.py:469,486)
                float mx = g;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float e = expf(g - mx);
                float sm = e;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
                const float wgt = e / sm;
                const float* b3 = sPar + P_B3B;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    float v[8], bb[8];
                    tc_ld8(t_c + c8 * 8, v);
                    ld8s(b3 + c8 * 8, bb);
                    tc_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) av[c8 * 8 + j] = wgt * leaky(v[j] + bb[j]);
                }
                // butterfly sum over the 16 rows of the point: lane s ends up with channels 4s..4s+3
#pragma unroll
                for (int w = 32; w >= 4; w >>= 1) {          // w = values kept after this step
                    const bool up = (lane & (w >> 2)) != 0;  // xor distance = w/4: 8, 4, 2, 1
#pragma unroll
                    for (int j = 0; j < w; ++j) {
                        const float send = up ? av[j] : av[j + w];
                        const float keep = up ? av[j + w] : av[j];
                        av[j] = keep + __shfl_xor_sync(0xffffffffu, send, w >> 2);
                    }
                }
                if (valid) *reinterpret_cast<float4*>(a.cost_fwd + qrow * CM + 4 * s) = make_float4(av[0], av[1], av[2], av[3]);
                sOut[(4 * s) * 8 + p] = av[0];
                sOut[(4 * s + 1) * 8 + p] = av[1];
                sOut[(4 * s + 2) * 8 + p] = av[2];
                sOut[(4 * s + 3) * 8 + p] = av[3];
                named_bar(2, 128);
                {   // channel-major copy: 4 consecutive points of one channel per thread
                    const int c = r >> 1, half = r & 1;
                    const int nn = n0 + half * 4;
                    float* dst = a.cost_fwd_cm + ((size_t)b * CM + c) * a.N1 + nn;
                    const float* src = sOut + c * 8 + half * 4;
                    if ((a.N1 & 3) == 0 && nn + 3 < a.N1) {
                        *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[1], src[2], src[3]);
                    } else {
                        for (int k = 0; k < 4; ++k)
                            if (nn + k < a.N1) dst[k] = src[k];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tc_dealloc(tmem, 512);
}

constexpr size_t CV_TC_SMEM = (size_t)NSTAGE * STAGE_BYTES + (size_t)(2 * ROWS * LDA + 520 + 4 * ROWS + CM * 8) * 4 + 16 * 8;

}  // namespace

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
