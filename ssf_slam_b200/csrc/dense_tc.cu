// Generic tensor-core dense layer for the 1x1-conv stacks of TFlow (ASF/utils/utils.py:236-247,304-313;
// ASF/utils/soflow.py:397-451,460-461,501-513): Y[rows, N] = epilogue(A[rows, K] . W[N, K]^T) on tcgen05
// `kind::tf32` with the fp32-faithful 3xTF32 split (tc_common.cuh), fp32 accumulators in TMEM.
//
// One CTA computes a 128-row x (<= 256)-column tile.  The K dimension is streamed in chunks of 32:
//   * A chunk: produced by the "row" threads (thread = row = TMEM lane).  Either plain rows of one or two
//     concatenated inputs, or the *grouped first layer* evaluated on the fly,
//         A[(b,n,s), c] = act1(G[b, idx[b,n,s], offG + c] + H[b, n, offH + c] + b1[c] + Wd1[:, c] . (pos_src[idx] - pos_q[n])),
//     so a gathered neighbourhood never exists in HBM.  The chunk is split into (hi, lo) TF32 halves and written
//     to TMEM with tcgen05.st; two row warpgroups alternate chunks so their global-load latencies overlap.
//   * W chunk: the host pre-arranges, per 32-wide K chunk, the [N x 32] hi and lo images in the no-swizzle
//     K-major UMMA layout; one cp.async.bulk (TMA) per chunk through a 3-stage shared-memory ring.
//   * one thread issues 12 MMAs per chunk (3 passes x 4 K-steps of 8).
// Epilogues (thread = row, the two warpgroups take half of the columns each):
//   STORE: y[row, :] = act(D + bias [+ Hq[row / S] + Wd2 . dir(row)])
//   MAX  : y[row / S, :] = max over the S rows of a point of the same expression   (S in {8, 16})
//   DOT  : y[row] = wvec . act(D + bias) + b0                                        (weightnet1's last conv)
#include "tc_common.cuh"
#include "ssf_dense.h"

namespace {

constexpr int KC = 32;         // K chunk
constexpr int WSTAGES = 3;
constexpr int DT_THREADS = 320;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ssf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return fmaxf(v, 0.1f * v);
    return v;
}
__device__ __forceinline__ void split8(const float (&v)[8], float (&hi)[8], float (&lo)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        hi[j] = __uint_as_float((__float_as_uint(v[j]) + 0x1000u) & 0xFFFFE000u);
        lo[j] = v[j] - hi[j];
    }
}
__device__ __forceinline__ void ldg8(const float* p, float (&o)[8]) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p)), y = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
}

__global__ void __launch_bounds__(DT_THREADS) dense_tc_kernel(ssf_dense_args a, int n_astage, int tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Nt = a.N < 256 ? a.N : 256;                 // columns of this CTA's tile
    const int n0 = blockIdx.y * 256;                       // first output column
    const uint32_t wchunk = (uint32_t)Nt * KC * 4 * 2;     // bytes of one weight chunk (hi + lo)
    uint8_t* sW = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WSTAGES * wchunk);
    uint64_t* w_full = bars;              // [WSTAGES]
    uint64_t* w_empty = bars + WSTAGES;   // [WSTAGES]
    uint64_t* a_ready = bars + 2 * WSTAGES;      // [4]
    uint64_t* a_empty = bars + 2 * WSTAGES + 4;  // [4]
    uint64_t* d_ready = bars + 2 * WSTAGES + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WSTAGES + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = a.K / KC;
    if (warp == 8) tc_alloc(tmem_slot, (uint32_t)tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < WSTAGES; ++i) {
            ssf_mbar_init(&w_full[i], 1);
            ssf_mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            ssf_mbar_init(&a_ready[i], 128);
            ssf_mbar_init(&a_empty[i], 1);
        }
        ssf_mbar_init(d_ready, 1);
        ssf_mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t a_col0 = (uint32_t)(tmem_cols == 512 ? 256 : 128);   // A stages live after the D columns

    if (warp == 9) {
        if (lane == 0) {   // ---- weight producer
            const uint8_t* wsrc = static_cast<const uint8_t*>(a.wimg) + (size_t)blockIdx.y * nk * wchunk;
            for (int kc = 0; kc < nk; ++kc) {
                const int st = kc % WSTAGES;
                if (kc >= WSTAGES) ssf_mbar_wait(&w_empty[st], (uint32_t)((kc / WSTAGES - 1) & 1));
                ssf_mbar_expect_tx(&w_full[st], wchunk);
                ssf_bulk_g2s(sW + st * wchunk, wsrc + (size_t)kc * wchunk, wchunk, &w_full[st]);
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        if (lane == 0) {   // ---- MMA issuer
            const uint32_t idesc = tc_idesc_tf32(128, Nt);
            const uint32_t lbo = (uint32_t)(Nt / 8) * 128;
            uint32_t acc = 0;
            for (int kc = 0; kc < nk; ++kc) {
                const int ws = kc % WSTAGES, as = kc % n_astage;
                ssf_mbar_wait(&w_full[ws], (uint32_t)((kc / WSTAGES) & 1));
                ssf_mbar_wait(&a_ready[as], (uint32_t)((kc / n_astage) & 1));
                tc_fence_after();
                const uint32_t w_hi = ssf_smem_u32(sW + ws * wchunk), w_lo = w_hi + (uint32_t)Nt * KC * 4;
                const uint32_t a_hi = tmem + a_col0 + as * 64, a_lo = a_hi + 32;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t aa = pass == 0 ? a_lo : a_hi;
                    const uint32_t ww = pass == 1 ? w_lo : w_hi;
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        tc_mma_ts(tmem, aa + ks * 8, tc_smem_desc(ww + ks * 2 * lbo, lbo, 128), idesc, acc);
                        acc = 1;
                    }
                }
                tc_commit(&a_empty[as]);
                tc_commit(&w_empty[ws]);
            }
            tc_commit(d_ready);
        }
        __syncwarp();
    } else {
        // ---- row threads: A producer, then epilogue
        const int wg = warp >> 2;
        const int r = tid & 127;
        const long long row = (long long)blockIdx.x * 128 + r;
        const bool valid = row < a.rows;
        const long long rowc = valid ? row : (long long)a.rows - 1;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        // grouped-row geometry (only when S > 0)
        long long pt = 0;        // flat point index b*Nq + n
        long long srow = 0;      // flat source row b*Nsrc + idx
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (a.S > 0) {
            pt = rowc / a.S;
            if (a.idx != nullptr) {
                const long long b = pt / a.Nq;
                srow = b * a.Nsrc + __ldg(a.idx + rowc);
                if (a.pos_src != nullptr) {
                    const float* ps = a.pos_src + srow * 3;
                    const float* pq = a.pos_q + pt * 3;
                    dx = __ldg(ps) - __ldg(pq);
                    dy = __ldg(ps + 1) - __ldg(pq + 1);
                    dz = __ldg(ps + 2) - __ldg(pq + 2);
                }
            }
        }
        for (int kc = wg; kc < nk; kc += 2) {
            const int as = kc % n_astage;
            const int k0 = kc * KC;
            float v[4][8];
            if (a.a_mode == 0) {
                const float* src = k0 < a.c1 ? a.x1 + rowc * a.ld1 + k0 : a.x2 + rowc * a.ld2 + (k0 - a.c1);
#pragma unroll
                for (int q = 0; q < 4; ++q) ldg8(src + q * 8, v[q]);
            } else {
                const float* gs = a.G + srow * a.ldG + a.offG + k0;
#pragma unroll
                for (int q = 0; q < 4; ++q) ldg8(gs + q * 8, v[q]);
                if (a.H != nullptr) {
                    const float* hs = a.H + pt * a.ldH + a.offH + k0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float h[8];
                        ldg8(hs + q * 8, h);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[q][j] += h[j];
                    }
                }
                if (a.Wd1 != nullptr) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float w0[8], w1[8], w2[8];
                        ldg8(a.Wd1 + k0 + q * 8, w0);
                        ldg8(a.Wd1 + a.K + k0 + q * 8, w1);
                        ldg8(a.Wd1 + 2 * a.K + k0 + q * 8, w2);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[q][j] += dx * w0[j] + dy * w1[j] + dz * w2[j];
                    }
                }
                if (a.b1 != nullptr) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float bb[8];
                        ldg8(a.b1 + k0 + q * 8, bb);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[q][j] += bb[j];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[q][j] = act_apply(v[q][j], a.act1);
            }
            if (kc >= n_astage) {
                ssf_mbar_wait(&a_empty[as], (uint32_t)((kc / n_astage - 1) & 1));
                tc_fence_after();
            }
            const uint32_t t_hi = tmem + lane_base + a_col0 + as * 64;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float hi[8], lo[8];
                split8(v[q], hi, lo);
                tc_st8(t_hi + q * 8, hi);
                tc_st8(t_hi + 32 + q * 8, lo);
            }
            tc_st_wait();
            tc_fence_before();
            mbar_arrive(&a_ready[as]);
        }
        // ---- epilogue
        ssf_mbar_wait(d_ready, 0);
        tc_fence_after();
        const int half = Nt >> 1;                      // columns per warpgroup (multiple of 8)
        const int cbeg = wg * half, cend = cbeg + half;
        const uint32_t t_d = tmem + lane_base;
        float dot = 0.f;
        for (int c = cbeg; c < cend; c += 8) {
            float v[8];
            tc_ld8(t_d + c, v);
            tc_ld_wait();
            const int cg = n0 + c;   // global output column
            if (a.bias != nullptr) {
                float bb[8];
                ldg8(a.bias + cg, bb);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += bb[j];
            }
            if (a.Hq != nullptr) {
                float h[8];
                ldg8(a.Hq + pt * a.ldHq + cg, h);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += h[j];
            }
            if (a.Wd2 != nullptr) {
                float w0[8], w1[8], w2[8];
                ldg8(a.Wd2 + cg, w0);
                ldg8(a.Wd2 + a.N + cg, w1);
                ldg8(a.Wd2 + 2 * a.N + cg, w2);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += dx * w0[j] + dy * w1[j] + dz * w2[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = act_apply(v[j], a.act);
            if (a.epi_mode == SSF_EPI_STORE) {
                if (valid) {
                    float* dst = a.y + row * a.ldy + cg;
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            } else if (a.epi_mode == SSF_EPI_MAX) {
                // rows beyond `rows` only exist in the last tile and belong to no stored point (rows % S == 0)
                if (a.S == 16) {
                    const bool u8 = (lane & 8) != 0, u4 = (lane & 4) != 0, u2 = (lane & 2) != 0;
                    float w4[4], w2[2];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        w4[j] = fmaxf(u8 ? v[j + 4] : v[j], __shfl_xor_sync(0xffffffffu, u8 ? v[j] : v[j + 4], 8));
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        w2[j] = fmaxf(u4 ? w4[j + 2] : w4[j], __shfl_xor_sync(0xffffffffu, u4 ? w4[j] : w4[j + 2], 4));
                    float w1 = fmaxf(u2 ? w2[1] : w2[0], __shfl_xor_sync(0xffffffffu, u2 ? w2[0] : w2[1], 2));
                    w1 = fmaxf(w1, __shfl_xor_sync(0xffffffffu, w1, 1));
                    if (valid && (lane & 1) == 0) a.y[pt * a.ldy + cg + ((lane >> 1) & 7)] = w1;
                } else {  // S == 8
                    const bool u4 = (lane & 4) != 0, u2 = (lane & 2) != 0, u1 = (lane & 1) != 0;
                    float w4[4], w2[2];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        w4[j] = fmaxf(u4 ? v[j + 4] : v[j], __shfl_xor_sync(0xffffffffu, u4 ? v[j] : v[j + 4], 4));
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        w2[j] = fmaxf(u2 ? w4[j + 2] : w4[j], __shfl_xor_sync(0xffffffffu, u2 ? w4[j] : w4[j + 2], 2));
                    const float w1 = fmaxf(u1 ? w2[1] : w2[0], __shfl_xor_sync(0xffffffffu, u1 ? w2[0] : w2[1], 1));
                    if (valid) a.y[pt * a.ldy + cg + (lane & 7)] = w1;
                }
            } else {  // DOT
                float w[8];
                ldg8(a.wvec + cg, w);
#pragma unroll
                for (int j = 0; j < 8; ++j) dot = fmaf(v[j], w[j], dot);
            }
        }
        if (a.epi_mode == SSF_EPI_DOT) {
            float* sDot = reinterpret_cast<float*>(tmem_slot + 4);   // [128]
            if (wg == 1) sDot[r] = dot;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (wg == 0 && valid) a.y[row] = dot + sDot[r] + a.b0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tc_dealloc(tmem, (uint32_t)tmem_cols);
}

}  // namespace

extern "C" int ssf_dense_args_bytes(void) { return (int)sizeof(ssf_dense_args); }

extern "C" int ssf_dense_tc(const ssf_dense_args* args, void* stream) {
    ssf_dense_args a = *args;
    if (a.rows <= 0) return ssf_arg_error("dense_tc: empty input");
    if (a.K <= 0 || a.K % KC) return ssf_arg_error("dense_tc: K must be a positive multiple of 32");
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
    const int Nt = a.N < 256 ? a.N : 256;
    const int tmem_cols = Nt <= 128 ? 256 : 512;
    const int n_astage = Nt <= 128 ? 2 : 4;
    const size_t smem = (size_t)WSTAGES * Nt * KC * 4 * 2 + 32 * 8 + 128 * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return ssf_set_error(e);
        attr_set = true;
    }
    dim3 grid((unsigned)((a.rows + 127) / 128), (unsigned)((a.N + 255) / 256));
    dense_tc_kernel<<<grid, DT_THREADS, smem, (cudaStream_t)stream>>>(a, n_astage, tmem_cols);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
