// Fused layers of the ActiveSceneFlow network (TFlow), hand-written for sm_100a.
//
// Data layout is point-major: features [B, N, C] with C contiguous, coordinates [B, N, 3]; weights are
// pre-transposed at model-load time to K-major [Cin, Cout] with eval-mode BatchNorm folded in.
// Grouped neighbourhoods ([B,C,S,N] in the reference) never reach HBM:
//
//  * The first 1x1 conv of every grouped MLP is split algebraically.  Its input is
//    cat[per-query block | gathered block | delta-xyz] (ASF/utils/utils.py:234,303; ASF/utils/soflow.py:397,428,494),
//    so  W.x = Wq.q[n] + Wg.g[idx[n,s]] + Wd.(xyz[idx]-xyz[n]).  Wq.q and Wg.g are computed ONCE per point by
//    `ssf_linear`; the fused kernels gather the already-projected rows (Cout wide instead of Cin wide),
//    add the exact delta-xyz term per (n,s) row and continue with the per-row layers in shared memory.
//  * `ssf_group_mlp_max`  = PointNetSetAbstraction / PointNetSetUpConv.mlp1 / mlp_convs4 (+ max over S)
//  * `ssf_cost_volume`    = PointConvTransFlowV2 core: both branches, SxS attention, mlp_convs3,
//    weightnet1, forward cost (ASF/utils/soflow.py:397-469,486); emits the warped-branch rows for the
//    deterministic segmented softmax/sum that forms the backward cost (:471-481).
//
// This file holds the SIMT fp32 (FFMA) realisations.  The product path runs the dense layers on tcgen05 (dense_tc.cu,
// cost_volume_tc.cu); these kernels remain (a) for the three layers tensor cores cannot take (3 input or 3 output
// channels: point_conv[0], fc), (b) as the in-repo fp32 reference the tensor-core kernels are tested against
// (functional.USE_TC = False), and they carry the small point-major helpers (transpose, row gather, interpolation,
// segmented softmax-sum).
#include "ssf_common.cuh"

constexpr int L_T = 256;   // threads per CTA in the fused kernels
constexpr int TD_KC = 32;  // K chunk staged per step
constexpr int TD_NC = 64;  // output columns per pass

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

__device__ __forceinline__ float act_fn(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LEAKY) return v > 0.f ? v : 0.1f * v;
    return v;
}

// Row owned by (thread-row tr, register r): interleaved so the two thread-rows of a warp touch adjacent
// shared-memory rows (distinct banks with the +4 padding).
__device__ __forceinline__ int td_row(int tr, int r) { return r * 16 + tr; }

// acc[r][0..3] += sum_k sIn[td_row(tr, r)][k] * Wt[k][o0 + tc*4 + 0..3]   for k in [0, Cin)
// sIn: shared [rows][ldIn] (ldIn % 4 == 0); Wt: global K-major [Cin][ldw]; sW: shared staging [TD_KC][TD_NC].
// Requires Cin % 32 == 0, ldw % 4 == 0, o0 % 4 == 0.  All L_T threads must call it (barriers inside).
template <int RM>
__device__ __forceinline__ void tile_dense_acc(const float* sIn, int ldIn, int Cin, const float* __restrict__ Wt, int ldw,
                                               int o0, int Cout, float* sW, float (&acc)[RM][4]) {
    const int tid = threadIdx.x;
    const int tr = tid >> 4, tc = tid & 15;
    for (int k0 = 0; k0 < Cin; k0 += TD_KC) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < (TD_KC * TD_NC / 4) / L_T; ++j) {
            const int f4 = tid + j * L_T;
            const int kk = f4 >> 4, c4 = f4 & 15;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o0 + c4 * 4 < Cout) w = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + kk) * ldw + o0 + c4 * 4));
            reinterpret_cast<float4*>(sW)[f4] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TD_KC; kk += 4) {
            float4 a[RM];
#pragma unroll
            for (int r = 0; r < RM; ++r) a[r] = *reinterpret_cast<const float4*>(sIn + (size_t)td_row(tr, r) * ldIn + k0 + kk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(sW + (kk + q) * TD_NC + tc * 4);
#pragma unroll
                for (int r = 0; r < RM; ++r) {
                    const float av = q == 0 ? a[r].x : (q == 1 ? a[r].y : (q == 2 ? a[r].z : a[r].w));
                    acc[r][0] = fmaf(av, w.x, acc[r][0]);
                    acc[r][1] = fmaf(av, w.y, acc[r][1]);
                    acc[r][2] = fmaf(av, w.z, acc[r][2]);
                    acc[r][3] = fmaf(av, w.w, acc[r][3]);
                }
            }
        }
    }
}

// sOut[row][c] = act(sIn[row] . Wt[:, c] + bias[c]) for all rows of the tile and c in [0, Cout); Cout % 4 == 0
template <int RM>
__device__ __forceinline__ void tile_dense(const float* sIn, int ldIn, int Cin, const float* __restrict__ Wt, int ldw,
                                           const float* __restrict__ bias, int Cout, int act, float* sOut, int ldOut, float* sW) {
    const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
    for (int o0 = 0; o0 < Cout; o0 += TD_NC) {
        float acc[RM][4];
#pragma unroll
        for (int r = 0; r < RM; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        tile_dense_acc<RM>(sIn, ldIn, Cin, Wt, ldw, o0, Cout, sW, acc);
        const int c = o0 + tc * 4;
        if (c < Cout) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(bias + c));
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                float4 v = make_float4(act_fn(acc[r][0] + b.x, act), act_fn(acc[r][1] + b.y, act),
                                       act_fn(acc[r][2] + b.z, act), act_fn(acc[r][3] + b.w, act));
                *reinterpret_cast<float4*>(sOut + (size_t)td_row(tr, r) * ldOut + c) = v;
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------- linear

struct LinearArgs {
    const float* x1; int c1; int ld1;   // first input block  [rows, c1]
    const float* x2; int c2; int ld2;   // optional second block (concatenated input), may be null
    const float* Wt; int ldw;           // K-major weights [c1 + c2 (+...), cout]; rows w_off1.. / w_off2..
    int w_off1; int w_off2;
    const float* bias;                  // [cout] or null
    int rows; int cout; int act;
    float clamp1;                       // > 0: clamp(act(.), +-clamp1)
    const float* add; int ld_add;       // optional residual [rows, cout], added after clamp1
    float clamp2;                       // > 0: final clamp
    float* y; int ldy;
};

// y[r, :] = epilogue(x1[r,:] . Wt[off1:off1+c1, :] + x2[r,:] . Wt[off2:off2+c2, :] + bias)
// 64 rows x 64 cols per CTA, 4x4 outputs per thread, K staged in chunks of 16 through shared memory.
__global__ void __launch_bounds__(256) linear_kernel(LinearArgs a) {
    __shared__ __align__(16) float sX[64][20];
    __shared__ __align__(16) float sWt[16][64];
    const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
    const int r0 = blockIdx.x * 64, o0 = blockIdx.y * 64;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    for (int seg = 0; seg < 2; ++seg) {
        const float* x = seg == 0 ? a.x1 : a.x2;
        if (x == nullptr) continue;
        const int C = seg == 0 ? a.c1 : a.c2, ld = seg == 0 ? a.ld1 : a.ld2, woff = seg == 0 ? a.w_off1 : a.w_off2;
        for (int k0 = 0; k0 < C; k0 += 16) {
            __syncthreads();
            {   // X tile: 64 rows x 16 k, one float4 per thread when aligned, scalars otherwise
                const int r = tid >> 2, kq = (tid & 3) * 4;
                const int row = r0 + r;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (row < a.rows) {
                    const float* p = x + (size_t)row * ld + k0 + kq;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (k0 + kq + q < C) v[q] = __ldg(p + q);
                }
                *reinterpret_cast<float4*>(&sX[r][kq]) = make_float4(v[0], v[1], v[2], v[3]);
            }
            {   // W tile: 16 k x 64 cols
                const int kk = tid >> 4, c4 = (tid & 15) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (k0 + kk < C) {
                    const float* p = a.Wt + (size_t)(woff + k0 + kk) * a.ldw + o0 + c4;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (o0 + c4 + q < a.cout) v[q] = __ldg(p + q);
                }
                *reinterpret_cast<float4*>(&sWt[kk][c4]) = make_float4(v[0], v[1], v[2], v[3]);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; kk += 4) {
                float4 xa[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) xa[r] = *reinterpret_cast<const float4*>(&sX[tr * 4 + r][kk]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w = *reinterpret_cast<const float4*>(&sWt[kk + q][tc * 4]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float av = q == 0 ? xa[r].x : (q == 1 ? xa[r].y : (q == 2 ? xa[r].z : xa[r].w));
                        acc[r][0] = fmaf(av, w.x, acc[r][0]);
                        acc[r][1] = fmaf(av, w.y, acc[r][1]);
                        acc[r][2] = fmaf(av, w.z, acc[r][2]);
                        acc[r][3] = fmaf(av, w.w, acc[r][3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = r0 + tr * 4 + r;
        if (row >= a.rows) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = o0 + tc * 4 + q;
            if (c >= a.cout) continue;
            float v = acc[r][q] + (a.bias ? __ldg(a.bias + c) : 0.f);
            v = act_fn(v, a.act);
            if (a.clamp1 > 0.f) v = fminf(fmaxf(v, -a.clamp1), a.clamp1);
            if (a.add != nullptr) v += __ldg(a.add + (size_t)row * a.ld_add + c);
            if (a.clamp2 > 0.f) v = fminf(fmaxf(v, -a.clamp2), a.clamp2);
            a.y[(size_t)row * a.ldy + c] = v;
        }
    }
}

// Narrow layers, where the 64 x 64 tile above wastes most of its lanes: same sums in the same order (one fmaf chain over
// ascending k from 0, then bias, activation, clamp, residual, clamp), so bit-identical to linear_kernel.
//   * few outputs (cout <= 4: the flow heads, 64-256 -> 3): lane = row; a warp reads 32 rows x 32 k with coalesced 128-byte loads,
//     transposes them through a padded shared-memory tile and every lane runs its row's chain; the layer reads its input once
//     at memory speed instead of spending 15 of 16 column-threads on padding (level 0: 173 -> ~30 us for 524288 rows)
//   * few inputs (K <= 4: the first per-point layer, 3 -> 32): thread = (row, 4 outputs), coalesced 16-byte stores
__global__ void __launch_bounds__(256) linear_narrow_out_kernel(LinearArgs a) {
    __shared__ float sT[8][32][33];
    __shared__ float sW[512 * 4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.c1;
    for (int i = tid; i < K * 4; i += 256) {
        const int k = i >> 2, c = i & 3;
        sW[i] = c < a.cout ? __ldg(a.Wt + (size_t)(a.w_off1 + k) * a.ldw + c) : 0.f;
    }
    __syncthreads();
    const long long r0 = ((long long)blockIdx.x * 8 + warp) * 32;
    if (r0 >= a.rows) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int kn = K - k0 < 32 ? K - k0 : 32;
        float v[32];   // all 32 row pieces of the chunk in flight before the first one is needed
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const long long row = r0 + j;
            v[j] = (row < a.rows && lane < kn) ? __ldg(a.x1 + (size_t)row * a.ld1 + k0 + lane) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) sT[warp][j][lane] = v[j];
        __syncwarp();
        for (int kk = 0; kk < kn; ++kk) {
            const float xv = sT[warp][lane][kk];
            const float4 w = *reinterpret_cast<const float4*>(sW + (k0 + kk) * 4);
            acc[0] = fmaf(xv, w.x, acc[0]);
            acc[1] = fmaf(xv, w.y, acc[1]);
            acc[2] = fmaf(xv, w.z, acc[2]);
            acc[3] = fmaf(xv, w.w, acc[3]);
        }
        __syncwarp();
    }
    const long long row = r0 + lane;
    if (row >= a.rows) return;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (c >= a.cout) continue;
        float v = acc[c] + (a.bias ? __ldg(a.bias + c) : 0.f);
        v = act_fn(v, a.act);
        if (a.clamp1 > 0.f) v = fminf(fmaxf(v, -a.clamp1), a.clamp1);
        if (a.add != nullptr) v += __ldg(a.add + (size_t)row * a.ld_add + c);
        if (a.clamp2 > 0.f) v = fminf(fmaxf(v, -a.clamp2), a.clamp2);
        a.y[(size_t)row * a.ldy + c] = v;
    }
}

__global__ void __launch_bounds__(256) linear_narrow_in_kernel(LinearArgs a, int q4) {   // q4 = cout / 4
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long row = t / q4;
    if (row >= a.rows) return;
    const int c = (int)(t - row * q4) * 4;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < a.c1) x[k] = __ldg(a.x1 + (size_t)row * a.ld1 + k);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < a.c1) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(a.Wt + (size_t)(a.w_off1 + k) * a.ldw + c));
            acc[0] = fmaf(x[k], w.x, acc[0]);
            acc[1] = fmaf(x[k], w.y, acc[1]);
            acc[2] = fmaf(x[k], w.z, acc[2]);
            acc[3] = fmaf(x[k], w.w, acc[3]);
        }
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[q] = acc[q] + (a.bias ? __ldg(a.bias + c + q) : 0.f);
        v[q] = act_fn(v[q], a.act);
        if (a.clamp1 > 0.f) v[q] = fminf(fmaxf(v[q], -a.clamp1), a.clamp1);
        if (a.add != nullptr) v[q] += __ldg(a.add + (size_t)row * a.ld_add + c + q);
        if (a.clamp2 > 0.f) v[q] = fminf(fmaxf(v[q], -a.clamp2), a.clamp2);
    }
    *reinterpret_cast<float4*>(a.y + (size_t)row * a.ldy + c) = make_float4(v[0], v[1], v[2], v[3]);
}

extern "C" int ssf_linear(const float* x1, int c1, int ld1, const float* x2, int c2, int ld2, const float* Wt, int ldw,
                          int w_off1, int w_off2, const float* bias, int rows, int cout, int act, float clamp1,
                          const float* add, int ld_add, float clamp2, float* y, int ldy, void* stream) {
    if (rows <= 0 || cout <= 0 || c1 <= 0) return ssf_arg_error("linear: empty input");
    LinearArgs a = {x1, c1, ld1, x2, x2 ? c2 : 0, ld2, Wt, ldw, w_off1, w_off2, bias, rows, cout, act, clamp1, add, ld_add, clamp2, y, ldy};
    // (small layers stay on the tiled kernel: one lane per row is a long serial chain when there are too few rows to hide it --
    // 32768 rows x 256: 57 us here, 17-27 us tiled)
    if (x2 == nullptr && cout <= 4 && c1 <= 512 && c1 >= 16 && rows > 65536) {
        linear_narrow_out_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    if (x2 == nullptr && c1 <= 4 && cout % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0 &&
        ((reinterpret_cast<uintptr_t>(Wt) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
        const int q4 = cout / 4;
        linear_narrow_in_kernel<<<(unsigned)(((long long)rows * q4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, q4);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    dim3 grid((rows + 63) / 64, (cout + 63) / 64);
    linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// ------------------------------------------------------------------------------ small helpers

// src [B,N,C], idx [B,M] -> out [B,M,C]  (point-major row gather: new_xyz = xyz[fps_idx])
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, int N, int M, int C,
                                   float* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (t >= (long long)M * C) return;
    const int m = (int)(t / C), c = (int)(t % C);
    out[((size_t)b * M + m) * C + c] = __ldg(src + ((size_t)b * N + idx[(size_t)b * M + m]) * C + c);
}

extern "C" int ssf_gather_rows(const float* src, const int* idx, int B, int N, int M, int C, float* out, void* stream) {
    if (B <= 0 || M <= 0 || C <= 0) return ssf_arg_error("gather_rows: empty input");
    dim3 grid((unsigned)(((long long)M * C + 255) / 256), B);
    gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, idx, N, M, C, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// [B,R,C] <-> [B,C,R] transposes (boundary layout conversion between the reference's channel-major tensors
// and the internal point-major ones)
__global__ void transpose_kernel(const float* __restrict__ in, int R, int C, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const float* ip = in + (size_t)b * R * C;
    float* op = out + (size_t)b * R * C;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? ip[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) op[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

extern "C" int ssf_transpose(const float* in, int B, int R, int C, float* out, void* stream) {
    if (B <= 0 || R <= 0 || C <= 0) return ssf_arg_error("transpose: empty input");
    dim3 grid((C + 31) / 32, (R + 31) / 32, B), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, R, C, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// Normalised inverse-distance interpolation (UpsampleFlow / PointWarping, ASF/utils/soflow.py:1462-1475,1244-1257):
//   w_j = (1/max(|src_pos[idx_j] - q|, 1e-10)) / sum_j(...),  v = sum_j w_j * src_val[idx_j]
//   mode 0: out = clamp(v, +-clampv)          (UpsampleFlow, clampv = 100)
//   mode 1: out = clamp(q - v, +-clampv)      (PointWarping,  clampv = 10, C == 3)
// query [B,N,3], src_pos [B,M,3], src_val [B,M,C], idx [B,N,k] -> out [B,N,C]; k <= 16
template <int VEC>
__global__ void __launch_bounds__(128)
interpolate_kernel(const float* __restrict__ query, const float* __restrict__ src_pos, const float* __restrict__ src_val,
                   const int* __restrict__ idx, int ld_idx, int N, int M, int C, int k, int mode, float clampv,
                   float* __restrict__ out) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * 4 + (threadIdx.x >> 5);  // one warp per query point
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const float* q = query + ((size_t)b * N + n) * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    float inv = 0.f;
    int id = 0;
    if (lane < k) {
        id = idx[((size_t)b * N + n) * ld_idx + lane];
        const float* s = src_pos + ((size_t)b * M + id) * 3;
        const float dx = s[0] - qx, dy = s[1] - qy, dz = s[2] - qz;
        const float d = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-10f);
        inv = 1.0f / d;
    }
    float norm = 0.f;
    for (int j = 0; j < k; ++j) norm += __shfl_sync(0xffffffffu, inv, j);  // slot order, as torch.sum over the last dim
    const float w = inv / norm;
    if (VEC > 1) {
        // C % VEC == 0, mode 0: every lane owns VEC consecutive channels (8- or 16-byte loads; 64 channels = one pass)
        for (int c0 = 0; c0 < C; c0 += 32 * VEC) {
            const int c = c0 + lane * VEC;
            float v[VEC];
#pragma unroll
            for (int u = 0; u < VEC; ++u) v[u] = 0.f;
            for (int j = 0; j < k; ++j) {
                const float wj = __shfl_sync(0xffffffffu, w, j);
                const int ij = __shfl_sync(0xffffffffu, id, j);
                if (c < C) {
                    const float* sp = src_val + ((size_t)b * M + ij) * C + c;
                    if (VEC == 2) {
                        const float2 x = __ldg(reinterpret_cast<const float2*>(sp));
                        v[0] += wj * x.x; v[1] += wj * x.y;
                    } else {
                        const float4 x = __ldg(reinterpret_cast<const float4*>(sp));
                        v[0] += wj * x.x; v[1] += wj * x.y; v[VEC - 2] += wj * x.z; v[VEC - 1] += wj * x.w;
                    }
                }
            }
            if (c < C) {
                float* op = out + ((size_t)b * N + n) * C + c;
#pragma unroll
                for (int u = 0; u < VEC; ++u) v[u] = fminf(fmaxf(v[u], -clampv), clampv);
                if (VEC == 2) *reinterpret_cast<float2*>(op) = make_float2(v[0], v[1]);
                else *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[VEC - 2], v[VEC - 1]);
            }
        }
        return;
    }
    for (int c0 = 0; c0 < C; c0 += 32) {  // warp-uniform trip count: every lane takes part in the shuffles
        const int c = c0 + lane;
        float v = 0.f;
        for (int j = 0; j < k; ++j) {
            const float wj = __shfl_sync(0xffffffffu, w, j);
            const int ij = __shfl_sync(0xffffffffu, id, j);
            if (c < C) v += wj * __ldg(src_val + ((size_t)b * M + ij) * C + c);
        }
        if (c < C) {
            if (mode == 1) v = (c == 0 ? qx : (c == 1 ? qy : qz)) - v;
            out[((size_t)b * N + n) * C + c] = fminf(fmaxf(v, -clampv), clampv);
        }
    }
}

// Up to 128 channels in 16-byte pieces: LPQ = 8 / 16 / 32 lanes per query point, so a warp interpolates 4 / 2 / 1 points at once
// (C = 64: half the warps of the warp-per-point kernel, C = 96: one 16-byte pass instead of three scalar ones).  mode 0,
// C % 4 == 0, C <= 4 LPQ, k <= LPQ; same operations in the same order as interpolate_kernel.
template <int LPQ>
__global__ void __launch_bounds__(128)
interpolate_group_kernel(const float* __restrict__ query, const float* __restrict__ src_pos, const float* __restrict__ src_val,
                         const int* __restrict__ idx, int ld_idx, int N, int M, int C, int k, float clampv,
                         float* __restrict__ out) {
    constexpr int QPW = 32 / LPQ;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, g = lane / LPQ, l = lane % LPQ;
    const int n_raw = (blockIdx.x * 4 + (threadIdx.x >> 5)) * QPW + g;
    const bool live = n_raw < N;
    const int n = live ? n_raw : N - 1;      // idle groups of the last warp shadow the last point (full-mask shuffles)
    const float* q = query + ((size_t)b * N + n) * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    float inv = 0.f;
    int id = 0;
    if (l < k) {
        id = idx[((size_t)b * N + n) * ld_idx + l];
        const float* s = src_pos + ((size_t)b * M + id) * 3;
        const float dx = s[0] - qx, dy = s[1] - qy, dz = s[2] - qz;
        const float d = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-10f);
        inv = 1.0f / d;
    }
    float norm = 0.f;
    for (int j = 0; j < k; ++j) norm += __shfl_sync(0xffffffffu, inv, g * LPQ + j);  // slot order, as torch.sum over the last dim
    const float w = inv / norm;
    const int c = l * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < k; ++j) {
        const float wj = __shfl_sync(0xffffffffu, w, g * LPQ + j);
        const int ij = __shfl_sync(0xffffffffu, id, g * LPQ + j);
        if (c < C) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(src_val + ((size_t)b * M + ij) * C + c));
            v[0] += wj * x.x; v[1] += wj * x.y; v[2] += wj * x.z; v[3] += wj * x.w;
        }
    }
    if (live && c < C) {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = fminf(fmaxf(v[u], -clampv), clampv);
        *reinterpret_cast<float4*>(out + ((size_t)b * N + n) * C + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// Few channels (C <= 4: the coarse flow of UpsampleFlow and PointWarping): one THREAD per query point -- a warp per point leaves
// 29 of 32 lanes idle and makes every point its own dependent load chain.  Same operations in the same order as above.
__global__ void __launch_bounds__(128)
interpolate_thread_kernel(const float* __restrict__ query, const float* __restrict__ src_pos, const float* __restrict__ src_val,
                          const int* __restrict__ idx, int ld_idx, int N, int M, int C, int k, int mode, float clampv,
                          float* __restrict__ out) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const float* q = query + ((size_t)b * N + n) * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    const int* ip = idx + ((size_t)b * N + n) * ld_idx;
    float inv[16];
    int id[16];
    float norm = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j < k) {
            id[j] = ip[j];
            const float* s = src_pos + ((size_t)b * M + id[j]) * 3;
            const float dx = s[0] - qx, dy = s[1] - qy, dz = s[2] - qz;
            const float d = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-10f);
            inv[j] = 1.0f / d;
            norm += inv[j];   // slot order, as torch.sum over the last dim
        }
    }
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j < k) {
            const float wj = inv[j] / norm;
            const float* sp = src_val + ((size_t)b * M + id[j]) * C;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < C) v[c] += wj * __ldg(sp + c);
        }
    }
    float* op = out + ((size_t)b * N + n) * C;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < C) {
            float r = v[c];
            if (mode == 1) r = (c == 0 ? qx : (c == 1 ? qy : qz)) - r;
            op[c] = fminf(fmaxf(r, -clampv), clampv);
        }
}

extern "C" int ssf_interpolate(const float* query, const float* src_pos, const float* src_val, const int* idx, int ld_idx, int B,
                               int N, int M, int C, int k, int mode, float clampv, float* out, void* stream) {
    if (B <= 0 || N <= 0 || C <= 0) return ssf_arg_error("interpolate: empty input");
    if (k <= 0 || k > 16 || ld_idx < k) return ssf_arg_error("interpolate: k must be in [1,16] and ld_idx >= k");
    if (mode == 1 && C != 3) return ssf_arg_error("interpolate: warp mode needs C == 3");
    if (C <= 4) {
        interpolate_thread_kernel<<<dim3((N + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(query, src_pos, src_val, idx, ld_idx, N, M, C, k, mode, clampv, out);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    const bool al16g = ((reinterpret_cast<uintptr_t>(src_val) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (mode == 0 && C % 4 == 0 && C <= 128 && k <= 8 && al16g) {
        cudaStream_t st = (cudaStream_t)stream;
        if (C <= 32)
            interpolate_group_kernel<8><<<dim3((N + 15) / 16, B), 128, 0, st>>>(query, src_pos, src_val, idx, ld_idx, N, M, C, k, clampv, out);
        else if (C <= 64)
            interpolate_group_kernel<16><<<dim3((N + 7) / 8, B), 128, 0, st>>>(query, src_pos, src_val, idx, ld_idx, N, M, C, k, clampv, out);
        else
            interpolate_group_kernel<32><<<dim3((N + 3) / 4, B), 128, 0, st>>>(query, src_pos, src_val, idx, ld_idx, N, M, C, k, clampv, out);
        ssf_count_launch();
        SSF_LAUNCH_CHECK();
        return SSF_OK;
    }
    dim3 grid((N + 3) / 4, B);
    const bool al16 = ((reinterpret_cast<uintptr_t>(src_val) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (mode == 0 && C % 128 == 0 && al16)
        interpolate_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>(query, src_pos, src_val, idx, ld_idx, N, M, C, k, mode, clampv, out);
    else if (mode == 0 && C % 64 == 0 && al16)
        interpolate_kernel<2><<<grid, 128, 0, (cudaStream_t)stream>>>(query, src_pos, src_val, idx, ld_idx, N, M, C, k, mode, clampv, out);
    else
        interpolate_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(query, src_pos, src_val, idx, ld_idx, N, M, C, k, mode, clampv, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// ------------------------------------------------------------------ grouped MLP + max over S

struct GroupMlpArgs {
    const float* G;        // [B, Nsrc, C1]  gathered block, already projected by the first layer's weights
    const float* H;        // [B, Nq, C1]    per-query block (projected) or null
    const float* bias1;    // [C1] or null (may already be folded into H)
    const float* Wd;       // [3, C1] K-major delta-xyz weights
    const float* pos_src;  // [B, Nsrc, 3]
    const float* pos_q;    // [B, Nq, 3]
    const int* idx;        // [B, Nq, S]
    const float* W2t; const float* b2; int C2;   // optional second layer (C2 == 0: absent)
    const float* W3t; const float* b3; int C3;   // optional third layer
    int Nsrc, Nq, S, C1, act;
    float* out;            // [B, Nq, Clast]
};

// rows (n,s) of a tile: x1 = act(G[idx] + H[n] + Wd.(pos_src[idx]-pos_q[n]) + b1); x2 = act(W2 x1 + b2); ...
// out[n] = max_s x_last.  ROWS = 16*RM rows per CTA (ROWS/S query points).
template <int RM>
__global__ void __launch_bounds__(L_T) group_mlp_max_kernel(GroupMlpArgs a) {
    constexpr int ROWS = 16 * RM;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int P = ROWS / a.S;  // query points per CTA
    const int n0 = blockIdx.x * P;
    const int ld1 = a.C1 + 4;
    const int ld2 = a.C2 + 4;
    float* sW = smem;                              // [32][64]
    float* sStage = sW + TD_KC * TD_NC;            // [ROWS][68]
    int* sIdx = reinterpret_cast<int*>(sStage + ROWS * 68);  // [ROWS]
    float* sDel = reinterpret_cast<float*>(sIdx + ROWS);     // [ROWS][4]
    float* sX1 = sDel + ROWS * 4;                  // [ROWS][ld1]
    float* sX2 = sX1 + ROWS * ld1;                 // [ROWS][ld2]  (only when a third layer exists)

    // neighbour indices and delta-xyz per row
    for (int r = tid; r < ROWS; r += L_T) {
        const int n = n0 + r / a.S;
        int id = 0;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (n < a.Nq) {
            id = a.idx[((size_t)b * a.Nq + n) * a.S + (r % a.S)];
            const float* ps = a.pos_src + ((size_t)b * a.Nsrc + id) * 3;
            const float* pq = a.pos_q + ((size_t)b * a.Nq + n) * 3;
            dx = ps[0] - pq[0];
            dy = ps[1] - pq[1];
            dz = ps[2] - pq[2];
        }
        sIdx[r] = id;
        sDel[4 * r] = dx;
        sDel[4 * r + 1] = dy;
        sDel[4 * r + 2] = dz;
    }
    __syncthreads();
    // first layer: gather-add
    {
        const int q4 = a.C1 >> 2;  // float4 per row
        for (int e = tid; e < ROWS * q4; e += L_T) {
            const int r = e / q4, c = (e % q4) * 4;
            const int n = min(n0 + r / a.S, a.Nq - 1);
            float4 v = __ldg(reinterpret_cast<const float4*>(a.G + ((size_t)b * a.Nsrc + sIdx[r]) * a.C1 + c));
            if (a.H != nullptr) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(a.H + ((size_t)b * a.Nq + n) * a.C1 + c));
                v.x += h.x; v.y += h.y; v.z += h.z; v.w += h.w;
            }
            const float dx = sDel[4 * r], dy = sDel[4 * r + 1], dz = sDel[4 * r + 2];
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.Wd + c));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.Wd + a.C1 + c));
            const float4 w2 = __ldg(reinterpret_cast<const float4*>(a.Wd + 2 * a.C1 + c));
            v.x += dx * w0.x + dy * w1.x + dz * w2.x;
            v.y += dx * w0.y + dy * w1.y + dz * w2.y;
            v.z += dx * w0.z + dy * w1.z + dz * w2.z;
            v.w += dx * w0.w + dy * w1.w + dz * w2.w;
            if (a.bias1 != nullptr) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias1 + c));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            v.x = act_fn(v.x, a.act); v.y = act_fn(v.y, a.act); v.z = act_fn(v.z, a.act); v.w = act_fn(v.w, a.act);
            *reinterpret_cast<float4*>(sX1 + (size_t)r * ld1 + c) = v;
        }
    }
    __syncthreads();
    const float* sLast = sX1;
    int ldL = ld1, CinL = a.C1;
    const float* Wl = a.W2t;
    const float* bl = a.b2;
    int Cl = a.C2;
    if (a.C3 > 0) {  // middle layer stays in shared memory
        tile_dense<RM>(sX1, ld1, a.C1, a.W2t, a.C2, a.b2, a.C2, a.act, sX2, ld2, sW);
        sLast = sX2; ldL = ld2; CinL = a.C2;
        Wl = a.W3t; bl = a.b3; Cl = a.C3;
    }
    // last layer, 64 columns at a time through the staging tile, then max over the S rows of each point
    const int tr = tid >> 4, tc = tid & 15;
    for (int o0 = 0; o0 < Cl; o0 += TD_NC) {
        float acc[RM][4];
#pragma unroll
        for (int r = 0; r < RM; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        tile_dense_acc<RM>(sLast, ldL, CinL, Wl, Cl, o0, Cl, sW, acc);
        const int c = o0 + tc * 4;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cl && bl != nullptr) bb = __ldg(reinterpret_cast<const float4*>(bl + c));
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            float4 v = make_float4(act_fn(acc[r][0] + bb.x, a.act), act_fn(acc[r][1] + bb.y, a.act),
                                   act_fn(acc[r][2] + bb.z, a.act), act_fn(acc[r][3] + bb.w, a.act));
            *reinterpret_cast<float4*>(sStage + (size_t)td_row(tr, r) * 68 + tc * 4) = v;
        }
        __syncthreads();
        for (int e = tid; e < P * TD_NC; e += L_T) {
            const int p = e / TD_NC, cc = e % TD_NC;
            const int n = n0 + p;
            if (n < a.Nq && o0 + cc < Cl) {
                float m = sStage[(size_t)(p * a.S) * 68 + cc];
                for (int s = 1; s < a.S; ++s) m = fmaxf(m, sStage[(size_t)(p * a.S + s) * 68 + cc]);
                a.out[((size_t)b * a.Nq + n) * Cl + o0 + cc] = m;
            }
        }
        __syncthreads();
    }
}

static size_t group_mlp_smem(int rows, int C1, int C2, int C3) {
    size_t f = (size_t)TD_KC * TD_NC + (size_t)rows * 68 + rows + (size_t)rows * 4 + (size_t)rows * (C1 + 4);
    if (C3 > 0) f += (size_t)rows * (C2 + 4);
    return f * sizeof(float);
}

extern "C" int ssf_group_mlp_max(const float* G, const float* H, const float* bias1, const float* Wd, const float* pos_src,
                                 const float* pos_q, const int* idx, const float* W2t, const float* b2, int C2,
                                 const float* W3t, const float* b3, int C3, int B, int Nsrc, int Nq, int S, int C1, int act,
                                 float* out, void* stream) {
    if (B <= 0 || Nq <= 0) return ssf_arg_error("group_mlp_max: empty input");
    if (S != 8 && S != 16) return ssf_arg_error("group_mlp_max: S must be 8 or 16");
    if (C1 % 32 || C2 % 32 || (C3 % 32) || C2 <= 0) return ssf_arg_error("group_mlp_max: channel widths must be multiples of 32");
    GroupMlpArgs a = {G, H, bias1, Wd, pos_src, pos_q, idx, W2t, b2, C2, W3t, b3, C3, Nsrc, Nq, S, C1, act, out};
    const int cmax = C3 > 0 ? (C1 > C2 ? C1 : C2) : C1;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (cmax <= 128) {
        const size_t smem = group_mlp_smem(128, C1, C2, C3);
        e = cudaFuncSetAttribute(group_mlp_max_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return ssf_set_error(e);
        dim3 grid((Nq + 128 / S - 1) / (128 / S), B);
        group_mlp_max_kernel<8><<<grid, L_T, smem, st>>>(a);
    } else if (cmax <= 256) {
        const size_t smem = group_mlp_smem(64, C1, C2, C3);
        e = cudaFuncSetAttribute(group_mlp_max_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return ssf_set_error(e);
        dim3 grid((Nq + 64 / S - 1) / (64 / S), B);
        group_mlp_max_kernel<4><<<grid, L_T, smem, st>>>(a);
    } else {
        return ssf_arg_error("group_mlp_max: hidden width > 256 not supported");
    }
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// --------------------------------------------------------------------------- cost volume core

struct CostVolArgs {
    const float* Gab;    // [B, N2, 2m]  f2 projected by mlp_convs[0] | mlp_convs2[0] (second halves of their inputs)
    const float* Hab;    // [B, N1, 2m]  f1 projected by the same layers (+ biases folded)
    const float* W2a; const float* b2a;   // mlp_convs[1]   K-major [m, m]
    const float* W2w; const float* b2w;   // mlp_convs2[1]
    const float* W3a;    // mlp_convs3[0] rows for the A block, K-major [m, m]
    const float* H3;     // [B, N1, m]   sf_feat projected by mlp_convs3[0] (+ bias folded); never null
    const float* W3d;    // [3, m]
    const float* W3b; const float* b3b;   // mlp_convs3[1]
    const float* Wn1; const float* bn1;   // weightnet1[0] + BN folded, [m, m]
    const float* Wn2; const float* bn2;   // weightnet1[3] + BN folded, [m, m/2]
    const float* wn3; float bn3;          // weightnet1[6]: [m/2] and scalar bias
    const float* xyz1;   // [B, N1, 3]
    const float* xyz2;   // [B, N2, 3]
    const int* idx;      // [B, N1, 16]
    const int* idxw;     // [B, N1, 16]
    int N1, N2, m;
    float* cost_fwd;     // [B, N1, m]   point-major
    float* cost_fwd_cm;  // [B, m, N1]   channel-major copy (the reference reinterprets this memory, soflow.py:490)
    float* gw;           // [B, N1*16]   warped-branch logits
    float* Cw;           // [B, N1*16, m] warped-branch cost rows
};

// One CTA handles ROWS = 16*RM rows = RM query points (S = 16 neighbours each).
template <int RM>
__global__ void __launch_bounds__(L_T) cost_volume_kernel(CostVolArgs a) {
    constexpr int ROWS = 16 * RM;
    constexpr int P = RM;  // points per CTA
    constexpr int S = 16;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * P;
    const int m = a.m, ld = m + 4, mh = m >> 1;
    float* sW = smem;                                   // [32][64]
    float* sQ = sW + TD_KC * TD_NC;                     // [P][16][17]
    float* sStat = sQ + P * 16 * 17;                    // row max/sum, col max/sum: [4][P*16]
    float* sG = sStat + 4 * P * 16;                     // [ROWS] logits
    float* sDir = sG + ROWS;                            // [ROWS][4]
    float* sDirW = sDir + ROWS * 4;                     // [ROWS][4]
    int* sIdx = reinterpret_cast<int*>(sDirW + ROWS * 4);   // [ROWS]
    int* sIdxW = sIdx + ROWS;                           // [ROWS]
    float* buf1 = reinterpret_cast<float*>(sIdxW + ROWS);   // A
    float* buf2 = buf1 + ROWS * ld;                     // Aw
    float* buf3 = buf2 + ROWS * ld;
    float* buf4 = buf3 + ROWS * ld;

    for (int r = tid; r < ROWS; r += L_T) {
        const int n = min(n0 + r / S, a.N1 - 1);
        const int i1 = a.idx[((size_t)b * a.N1 + n) * S + (r % S)];
        const int i2 = a.idxw[((size_t)b * a.N1 + n) * S + (r % S)];
        const float* pq = a.xyz1 + ((size_t)b * a.N1 + n) * 3;
        const float* p1 = a.xyz2 + ((size_t)b * a.N2 + i1) * 3;
        const float* p2 = a.xyz2 + ((size_t)b * a.N2 + i2) * 3;  // un-warped xyz2 with warped-cloud indices (soflow.py:407)
        sIdx[r] = i1;
        sIdxW[r] = i2;
        sDir[4 * r] = p1[0] - pq[0]; sDir[4 * r + 1] = p1[1] - pq[1]; sDir[4 * r + 2] = p1[2] - pq[2];
        sDirW[4 * r] = p2[0] - pq[0]; sDirW[4 * r + 1] = p2[1] - pq[1]; sDirW[4 * r + 2] = p2[2] - pq[2];
    }
    __syncthreads();

    // ---- first layers of mlp_convs / mlp_convs2 (gather-add), then their second layers -> A (buf1), Aw (buf2)
    const int q4 = m >> 2;
    for (int br = 0; br < 2; ++br) {
        const int* sI = br == 0 ? sIdx : sIdxW;
        for (int e = tid; e < ROWS * q4; e += L_T) {
            const int r = e / q4, c = (e % q4) * 4;
            const int n = min(n0 + r / S, a.N1 - 1);
            const float4 g = __ldg(reinterpret_cast<const float4*>(a.Gab + ((size_t)b * a.N2 + sI[r]) * (2 * m) + br * m + c));
            const float4 h = __ldg(reinterpret_cast<const float4*>(a.Hab + ((size_t)b * a.N1 + n) * (2 * m) + br * m + c));
            float4 v = make_float4(act_fn(g.x + h.x, ACT_LEAKY), act_fn(g.y + h.y, ACT_LEAKY), act_fn(g.z + h.z, ACT_LEAKY),
                                   act_fn(g.w + h.w, ACT_LEAKY));
            *reinterpret_cast<float4*>(buf3 + (size_t)r * ld + c) = v;
        }
        __syncthreads();
        tile_dense<RM>(buf3, ld, m, br == 0 ? a.W2a : a.W2w, m, br == 0 ? a.b2a : a.b2w, m, ACT_LEAKY, br == 0 ? buf1 : buf2, ld, sW);
    }

    // ---- S x S attention: Q[p][i][j] = <A[p,i,:], Aw[p,j,:]>, Q <- softmax_i(Q) * softmax_j(Q)   (soflow.py:420-422)
    for (int e = tid; e < P * 256; e += L_T) {
        const int p = e >> 8, i = (e >> 4) & 15, j = e & 15;
        const float* ar = buf1 + (size_t)(p * 16 + i) * ld;
        const float* wr = buf2 + (size_t)(p * 16 + j) * ld;
        float s = 0.f;
        for (int c = 0; c < m; c += 4) {
            const float4 x = *reinterpret_cast<const float4*>(ar + c);
            const float4 y = *reinterpret_cast<const float4*>(wr + c);
            s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
        }
        sQ[(p * 16 + i) * 17 + j] = s;
    }
    __syncthreads();
    for (int e = tid; e < P * 32; e += L_T) {
        const int p = e >> 5, which = (e >> 4) & 1, t = e & 15;  // which 0: row i = t (over j), 1: column j = t (over i)
        float mx = -INFINITY;
        for (int u = 0; u < 16; ++u) mx = fmaxf(mx, which == 0 ? sQ[(p * 16 + t) * 17 + u] : sQ[(p * 16 + u) * 17 + t]);
        float sm = 0.f;
        for (int u = 0; u < 16; ++u) sm += expf((which == 0 ? sQ[(p * 16 + t) * 17 + u] : sQ[(p * 16 + u) * 17 + t]) - mx);
        sStat[(which * 2) * P * 16 + p * 16 + t] = mx;
        sStat[(which * 2 + 1) * P * 16 + p * 16 + t] = sm;
    }
    __syncthreads();
    for (int e = tid; e < P * 256; e += L_T) {
        const int p = e >> 8, i = (e >> 4) & 15, j = e & 15;
        const float q = sQ[(p * 16 + i) * 17 + j];
        const float over_j = expf(q - sStat[p * 16 + i]) / sStat[P * 16 + p * 16 + i];                  // softmax(-1)
        const float over_i = expf(q - sStat[2 * P * 16 + p * 16 + j]) / sStat[3 * P * 16 + p * 16 + j];  // softmax(-2)
        sQ[(p * 16 + i) * 17 + j] = over_i * over_j;
    }
    __syncthreads();

    // helper lambdas ---------------------------------------------------------------------------
    // mixed = self + attention-weighted other branch   (soflow.py:453-458) -> dst
    auto mix = [&](const float* self, const float* other, bool transposed, float* dst) {
        for (int e = tid; e < ROWS * q4; e += L_T) {
            const int r = e / q4, c = (e % q4) * 4;
            const int p = r >> 4, i = r & 15;
            float4 acc = *reinterpret_cast<const float4*>(self + (size_t)r * ld + c);
            for (int j = 0; j < 16; ++j) {
                const float q = transposed ? sQ[(p * 16 + j) * 17 + i] : sQ[(p * 16 + i) * 17 + j];
                const float4 o = *reinterpret_cast<const float4*>(other + (size_t)(p * 16 + j) * ld + c);
                acc.x = fmaf(q, o.x, acc.x); acc.y = fmaf(q, o.y, acc.y); acc.z = fmaf(q, o.z, acc.z); acc.w = fmaf(q, o.w, acc.w);
            }
            *reinterpret_cast<float4*>(dst + (size_t)r * ld + c) = acc;
        }
        __syncthreads();
    };
    // weightnet1 on `src` (uses tmpA, tmpB), logits -> sG
    auto weightnet = [&](const float* src, float* tmpA, float* tmpB) {
        tile_dense<RM>(src, ld, m, a.Wn1, m, a.bn1, m, ACT_RELU, tmpA, ld, sW);
        tile_dense<RM>(tmpA, ld, m, a.Wn2, mh, a.bn2, mh, ACT_RELU, tmpB, ld, sW);
        for (int r = tid >> 1; r < ROWS; r += L_T / 2) {  // two threads per row
            const int half = tid & 1;
            float s = 0.f;
            for (int c = half; c < mh; c += 2) s = fmaf(tmpB[(size_t)r * ld + c], __ldg(a.wn3 + c), s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (half == 0) sG[r] = s + a.bn3;
        }
        __syncthreads();
    };
    // first layer of mlp_convs3 on `src` with direction `dir` -> dst  (cat[A | sf_feat | dir], soflow.py:428-444)
    auto mlp3_first = [&](const float* src, const float* dir, float* dst) {
        for (int o0 = 0; o0 < m; o0 += TD_NC) {
            float acc[RM][4];
#pragma unroll
            for (int r = 0; r < RM; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
            tile_dense_acc<RM>(src, ld, m, a.W3a, m, o0, m, sW, acc);
            const int c = o0 + tc * 4;
            if (c < m) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.W3d + c));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.W3d + m + c));
                const float4 w2 = __ldg(reinterpret_cast<const float4*>(a.W3d + 2 * m + c));
#pragma unroll
                for (int r = 0; r < RM; ++r) {
                    const int row = td_row(tr, r);
                    const int n = min(n0 + row / S, a.N1 - 1);
                    const float4 h = __ldg(reinterpret_cast<const float4*>(a.H3 + ((size_t)b * a.N1 + n) * m + c));
                    const float dx = dir[4 * row], dy = dir[4 * row + 1], dz = dir[4 * row + 2];
                    float4 v;
                    v.x = act_fn(acc[r][0] + h.x + (dx * w0.x + dy * w1.x + dz * w2.x), ACT_LEAKY);
                    v.y = act_fn(acc[r][1] + h.y + (dx * w0.y + dy * w1.y + dz * w2.y), ACT_LEAKY);
                    v.z = act_fn(acc[r][2] + h.z + (dx * w0.z + dy * w1.z + dz * w2.z), ACT_LEAKY);
                    v.w = act_fn(acc[r][3] + h.w + (dx * w0.w + dy * w1.w + dz * w2.w), ACT_LEAKY);
                    *reinterpret_cast<float4*>(dst + (size_t)row * ld + c) = v;
                }
            }
        }
        __syncthreads();
    };

    // ---- forward branch: logits g, cost rows C, forward cost
    mix(buf1, buf2, false, buf3);                 // A' = A + Q.Aw
    weightnet(buf3, buf4, buf3);                  // g -> sG   (buf3 reused for the m/2-wide layer)
    mlp3_first(buf1, sDir, buf3);
    tile_dense<RM>(buf3, ld, m, a.W3b, m, a.b3b, m, ACT_LEAKY, buf4, ld, sW);   // C -> buf4
    if (tid < P) {  // softmax over the 16 slots of each point (nn.Softmax(dim=2), soflow.py:469)
        float mx = -INFINITY;
        for (int s = 0; s < 16; ++s) mx = fmaxf(mx, sG[tid * 16 + s]);
        float sm = 0.f;
        for (int s = 0; s < 16; ++s) sm += expf(sG[tid * 16 + s] - mx);
        for (int s = 0; s < 16; ++s) sG[tid * 16 + s] = expf(sG[tid * 16 + s] - mx) / sm;
    }
    __syncthreads();
    for (int e = tid; e < P * m; e += L_T) {
        const int p = e / m, c = e % m;
        const int n = n0 + p;
        if (n < a.N1) {
            float s = 0.f;
            for (int u = 0; u < 16; ++u) s = fmaf(sG[p * 16 + u], buf4[(size_t)(p * 16 + u) * ld + c], s);
            a.cost_fwd[((size_t)b * a.N1 + n) * m + c] = s;
            a.cost_fwd_cm[((size_t)b * m + c) * a.N1 + n] = s;
        }
    }
    __syncthreads();

    // ---- warped branch: cost rows Cw and logits gw go to HBM for the segmented softmax/sum
    mlp3_first(buf2, sDirW, buf3);
    for (int o0 = 0; o0 < m; o0 += TD_NC) {
        float acc[RM][4];
#pragma unroll
        for (int r = 0; r < RM; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        tile_dense_acc<RM>(buf3, ld, m, a.W3b, m, o0, m, sW, acc);
        const int c = o0 + tc * 4;
        if (c < m) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b3b + c));
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                const int row = td_row(tr, r);
                const int n = n0 + row / S;
                if (n < a.N1) {
                    float4 v = make_float4(act_fn(acc[r][0] + bb.x, ACT_LEAKY), act_fn(acc[r][1] + bb.y, ACT_LEAKY),
                                           act_fn(acc[r][2] + bb.z, ACT_LEAKY), act_fn(acc[r][3] + bb.w, ACT_LEAKY));
                    *reinterpret_cast<float4*>(a.Cw + (((size_t)b * a.N1 + n) * S + (row % S)) * m + c) = v;
                }
            }
        }
    }
    __syncthreads();
    mix(buf2, buf1, true, buf3);                  // Aw' = Aw + A^T.Q
    weightnet(buf3, buf4, buf3);
    for (int r = tid; r < ROWS; r += L_T) {
        const int n = n0 + r / S;
        if (n < a.N1) a.gw[((size_t)b * a.N1 + n) * S + (r % S)] = sG[r];
    }
}

static size_t cost_volume_smem(int rows, int m) {
    const int P = rows / 16;
    size_t f = (size_t)TD_KC * TD_NC + (size_t)P * 16 * 17 + 4 * (size_t)P * 16 + rows + 2 * (size_t)rows * 4 + 2 * (size_t)rows +
               4 * (size_t)rows * (m + 4);
    return f * sizeof(float);
}

extern "C" int ssf_cost_volume(const float* Gab, const float* Hab, const float* W2a, const float* b2a, const float* W2w,
                               const float* b2w, const float* W3a, const float* H3, const float* W3d, const float* W3b,
                               const float* b3b, const float* Wn1, const float* bn1, const float* Wn2, const float* bn2,
                               const float* wn3, float bn3, const float* xyz1, const float* xyz2, const int* idx,
                               const int* idxw, int B, int N1, int N2, int m, float* cost_fwd, float* cost_fwd_cm, float* gw,
                               float* Cw, void* stream) {
    if (B <= 0 || N1 <= 0) return ssf_arg_error("cost_volume: empty input");
    if (m != 64 && m != 128 && m != 256) return ssf_arg_error("cost_volume: m must be 64, 128 or 256");
    CostVolArgs a = {Gab, Hab, W2a, b2a, W2w, b2w, W3a, H3, W3d, W3b, b3b, Wn1, bn1, Wn2, bn2, wn3, bn3,
                     xyz1, xyz2, idx, idxw, N1, N2, m, cost_fwd, cost_fwd_cm, gw, Cw};
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (m == 64) {
        const size_t smem = cost_volume_smem(128, m);
        e = cudaFuncSetAttribute(cost_volume_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return ssf_set_error(e);
        cost_volume_kernel<8><<<dim3((N1 + 7) / 8, B), L_T, smem, st>>>(a);
    } else if (m == 128) {
        const size_t smem = cost_volume_smem(64, m);
        e = cudaFuncSetAttribute(cost_volume_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return ssf_set_error(e);
        cost_volume_kernel<4><<<dim3((N1 + 3) / 4, B), L_T, smem, st>>>(a);
    } else {
        const size_t smem = cost_volume_smem(32, m);
        e = cudaFuncSetAttribute(cost_volume_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return ssf_set_error(e);
        cost_volume_kernel<2><<<dim3((N1 + 1) / 2, B), L_T, smem, st>>>(a);
    }
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// ------------------------------------------------- backward cost: segmented softmax-weighted sum

// cost_bwd[j, :] = sum_{l in seg(j)} softmax_seg(logit)[l] * val[l, :]     (soflow.py:471-481, one fused pass)
// One warp per target point j.  The softmax weights of up to 32 rows are computed lane-parallel (one expf and one
// division per row), then broadcast row by row while every lane accumulates its channels with coalesced vector loads.
// Rows are visited in ascending l and every reduction has a fixed shape, so the result is deterministic.
template <int VEC>
__global__ void __launch_bounds__(256)
seg_softmax_sum_kernel(const float* __restrict__ logit, const float* __restrict__ val, const int* __restrict__ ws, int B,
                       int L, int C, int n_seg, float* __restrict__ out) {
    constexpr int MAXV = 8;   // channels per lane: C <= 256
    const int b = blockIdx.y;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= n_seg) return;
    const int* off = ws + (size_t)b * (n_seg + 1);
    const int* rows = ws + (size_t)B * (n_seg + 1) + (size_t)B * n_seg + (size_t)b * L;
    const float* lg = logit + (size_t)b * L;
    const float* v = val + (size_t)b * L * C;
    const int lo = off[j], hi = off[j + 1];
    float mx = -INFINITY;
    for (int a = lo + lane; a < hi; a += 32) mx = fmaxf(mx, lg[rows[a]]);
    mx = ssf_warp_max(mx);
    float den = 0.f;
    for (int base = lo; base < hi; base += 32) {   // batches in ascending order, fixed-shape tree inside a batch
        const int a = base + lane;
        den += ssf_warp_sum(a < hi ? expf(lg[rows[a]] - mx) : 0.f);
    }
    float acc[MAXV];
#pragma unroll
    for (int q = 0; q < MAXV; ++q) acc[q] = 0.f;
    for (int base = lo; base < hi; base += 32) {
        const int a = base + lane;
        int r_l = 0;
        float w_l = 0.f;
        if (a < hi) {
            r_l = rows[a];
            w_l = expf(lg[r_l] - mx) / den;
        }
        const int nb = min(32, hi - base);
        // four rows per step: their loads are independent and in flight together (a segment is a short list of random
        // rows, so a one-row-at-a-time loop is a chain of exposed L2 latencies); accumulation order stays ascending
        for (int t0 = 0; t0 < nb; t0 += 4) {
            float wv[4];
            const float* vr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = min(t0 + u, nb - 1);
                const int r = __shfl_sync(0xffffffffu, r_l, t);
                wv[u] = t0 + u < nb ? __shfl_sync(0xffffffffu, w_l, t) : 0.f;
                vr[u] = v + (size_t)r * C;
            }
#pragma unroll
            for (int q = 0; q < MAXV / VEC; ++q) {
                const int c = (q * 32 + lane) * VEC;
                if (c < C) {
                    if (VEC == 2) {
                        float2 x[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) x[u] = __ldg(reinterpret_cast<const float2*>(vr[u] + c));
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            acc[q * 2] += x[u].x * wv[u];
                            acc[q * 2 + 1] += x[u].y * wv[u];
                        }
                    } else {
                        float x[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) x[u] = __ldg(vr[u] + c);
#pragma unroll
                        for (int u = 0; u < 4; ++u) acc[q] += x[u] * wv[u];
                    }
                }
            }
        }
    }
    float* o = out + ((size_t)b * n_seg + j) * C;
#pragma unroll
    for (int q = 0; q < MAXV / VEC; ++q) {
        const int c = (q * 32 + lane) * VEC;
        if (c < C) {
            if (VEC == 2) *reinterpret_cast<float2*>(o + c) = make_float2(acc[q * 2], acc[q * 2 + 1]);
            else o[c] = acc[q];
        }
    }
}

// C == 64: a row is 256 bytes = 16 lanes x 16 bytes, so the two half-warps take alternate rows of a segment (eight rows in
// flight per step, half as many load instructions); the two partial sums are added at the end.  Fixed shape -> deterministic.
__global__ void __launch_bounds__(256)
seg_softmax_sum64_kernel(const float* __restrict__ logit, const float* __restrict__ val, const int* __restrict__ ws, int B,
                         int L, int n_seg, float* __restrict__ out) {
    constexpr int C = 64;
    const int b = blockIdx.y;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
    if (j >= n_seg) return;
    const int* off = ws + (size_t)b * (n_seg + 1);
    const int* rows = ws + (size_t)B * (n_seg + 1) + (size_t)B * n_seg + (size_t)b * L;
    const float* lg = logit + (size_t)b * L;
    const float* v = val + (size_t)b * L * C;
    const int lo = off[j], hi = off[j + 1];
    float mx = -INFINITY;
    for (int a = lo + lane; a < hi; a += 32) mx = fmaxf(mx, lg[rows[a]]);
    mx = ssf_warp_max(mx);
    float den = 0.f;
    for (int base = lo; base < hi; base += 32) {
        const int a = base + lane;
        den += ssf_warp_sum(a < hi ? expf(lg[rows[a]] - mx) : 0.f);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = lo; base < hi; base += 32) {
        const int a = base + lane;
        int r_l = 0;
        float w_l = 0.f;
        if (a < hi) {
            r_l = rows[a];
            w_l = expf(lg[r_l] - mx) / den;
        }
        const int nb = min(32, hi - base);
        for (int t0 = 0; t0 < nb; t0 += 8) {
            float wv[4];
            float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + 2 * u + half;                       // this half-warp's u-th row of the step
                const int tc = min(t, nb - 1);
                const int r = __shfl_sync(0xffffffffu, r_l, tc);
                const float w = __shfl_sync(0xffffffffu, w_l, tc);
                wv[u] = t < nb ? w : 0.f;
                x[u] = __ldg(reinterpret_cast<const float4*>(v + (size_t)r * C) + l16);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x += x[u].x * wv[u];
                acc.y += x[u].y * wv[u];
                acc.z += x[u].z * wv[u];
                acc.w += x[u].w * wv[u];
            }
        }
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    if (half == 0) reinterpret_cast<float4*>(out + ((size_t)b * n_seg + j) * C)[l16] = acc;
}

extern "C" int ssf_segment_softmax_sum(const float* logit, const float* val, const int* csr_ws, int B, int L, int C,
                                       int n_seg, float* out, void* stream) {
    if (B <= 0 || L <= 0 || C <= 0 || n_seg <= 0) return ssf_arg_error("segment_softmax_sum: empty input");
    if (C > 256) return ssf_arg_error("segment_softmax_sum: at most 256 channels");
    dim3 grid((n_seg + 7) / 8, B);
    if (C == 64 && (reinterpret_cast<uintptr_t>(val) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        seg_softmax_sum64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logit, val, csr_ws, B, L, n_seg, out);
    else if (C % 2 == 0 && (reinterpret_cast<uintptr_t>(val) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0)
        seg_softmax_sum_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(logit, val, csr_ws, B, L, C, n_seg, out);
    else
        seg_softmax_sum_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(logit, val, csr_ws, B, L, C, n_seg, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
