// CUDA-core pieces of the un-fused (wide, m >= 128) PointConvTransFlowV2 path: the S x S cross attention between the
// two branches (ASF/utils/soflow.py:420-422,453-458) and the softmax-weighted forward cost (:469,486).  The dense
// layers around them run on tcgen05 (dense_tc.cu); at levels 2 and 3 a cloud has only 512 / 256 query points, so the
// activations exchanged through HBM are a few MB per cloud.
#include "ssf_common.cuh"

// A, Aw [P, 16, m] -> Amix = A + Q.Aw, Awmix = Aw + Q^T.A with Q = softmax_i(<A_i,Aw_j>) * softmax_j(<A_i,Aw_j>)
// Persistent CTAs (256 threads) walk over the points; the 2 x 16 rows of the NEXT point are fetched by the TMA engine
// (cp.async.bulk, one row per copy into the padded shared-memory rows, completion on an mbarrier) while the current point is
// processed: the kernel moves 4 x 16 x m floats per point and was latency-bound with one point per CTA (2.2-2.8 TB/s).
__global__ void __launch_bounds__(256, 3) attention_mix_kernel(const float* __restrict__ A, const float* __restrict__ Aw, int m, long long n_points,
                                                            float* __restrict__ Amix, float* __restrict__ Awmix) {
    extern __shared__ __align__(16) float sm[];
    const int ld = m + 4;
    constexpr int LQ = 20;          // row stride of sQ: 16-byte aligned rows
    float* sQ = sm;                 // [16][LQ]
    float* sQT = sQ + 16 * LQ;      // [16][LQ] the normalised Q transposed: both mixes read their four weights as one float4
    float* sStat = sQT + 16 * LQ;   // row max | row sum | col max | col sum, [16] each
    float* sBuf = sStat + 64;       // 2 x ([16][ld] A | [16][ld] Aw)
    __shared__ uint64_t bars[2];
    const int tid = threadIdx.x;
    const int q4 = m >> 2;
    if (tid == 0) {
        ssf_mbar_init(&bars[0], 1);
        ssf_mbar_init(&bars[1], 1);
        ssf_mbar_fence_init();
    }
    __syncthreads();
    auto fetch = [&](long long p, int buf) {   // warp 0: lane l < 16 copies row l of A and of Aw
        if (tid < 32) {
            if (tid == 0) ssf_mbar_expect_tx(&bars[buf], (uint32_t)(2 * 16 * m) * 4u);
            __syncwarp();
            if (tid < 16) {
                float* dA = sBuf + (size_t)buf * 32 * ld + tid * ld;
                const size_t src = (size_t)p * 16 * m + (size_t)tid * m;
                ssf_bulk_g2s(dA, A + src, (uint32_t)m * 4u, &bars[buf]);
                ssf_bulk_g2s(dA + 16 * ld, Aw + src, (uint32_t)m * 4u, &bars[buf]);
            }
        }
    };
    if ((long long)blockIdx.x < n_points) fetch(blockIdx.x, 0);
    int it = 0;
    for (long long p = blockIdx.x; p < n_points; p += gridDim.x, ++it) {
        const int buf = it & 1;
        // the other buffer was last read two barriers ago (end of the previous iteration)
        if (p + gridDim.x < n_points) fetch(p + gridDim.x, buf ^ 1);
        ssf_mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        const float* sA = sBuf + (size_t)buf * 32 * ld;   // [16][ld]
        const float* sW = sA + 16 * ld;                   // [16][ld]
        const size_t base = (size_t)p * 16 * m;
        {   // Q[i][j] = <A_i, Aw_j>: thread = 2 x 2 block over a quarter of the channels (every value read from shared memory
            // feeds two FMAs), partial sums combined over the 4 slices by shuffles
            const int blk = tid >> 2, ks = tid & 3, bi = blk >> 3, bj = blk & 7;
            const float* a0 = sA + (2 * bi) * ld, *a1 = a0 + ld;
            const float* w0 = sW + (2 * bj) * ld, *w1 = w0 + ld;
            float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
            for (int c = ks * 4; c < m; c += 16) {   // slice ks owns the channel quads ks, ks + 4, ...
                const float4 x0 = *reinterpret_cast<const float4*>(a0 + c), x1 = *reinterpret_cast<const float4*>(a1 + c);
                const float4 y0 = *reinterpret_cast<const float4*>(w0 + c), y1 = *reinterpret_cast<const float4*>(w1 + c);
                s00 = fmaf(x0.x, y0.x, s00); s00 = fmaf(x0.y, y0.y, s00); s00 = fmaf(x0.z, y0.z, s00); s00 = fmaf(x0.w, y0.w, s00);
                s01 = fmaf(x0.x, y1.x, s01); s01 = fmaf(x0.y, y1.y, s01); s01 = fmaf(x0.z, y1.z, s01); s01 = fmaf(x0.w, y1.w, s01);
                s10 = fmaf(x1.x, y0.x, s10); s10 = fmaf(x1.y, y0.y, s10); s10 = fmaf(x1.z, y0.z, s10); s10 = fmaf(x1.w, y0.w, s10);
                s11 = fmaf(x1.x, y1.x, s11); s11 = fmaf(x1.y, y1.y, s11); s11 = fmaf(x1.z, y1.z, s11); s11 = fmaf(x1.w, y1.w, s11);
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                s00 += __shfl_xor_sync(0xffffffffu, s00, o);
                s01 += __shfl_xor_sync(0xffffffffu, s01, o);
                s10 += __shfl_xor_sync(0xffffffffu, s10, o);
                s11 += __shfl_xor_sync(0xffffffffu, s11, o);
            }
            if (ks == 0) {
                sQ[(2 * bi) * LQ + 2 * bj] = s00;
                sQ[(2 * bi) * LQ + 2 * bj + 1] = s01;
                sQ[(2 * bi + 1) * LQ + 2 * bj] = s10;
                sQ[(2 * bi + 1) * LQ + 2 * bj + 1] = s11;
            }
        }
        __syncthreads();
        if (tid < 32) {
            const int which = tid >> 4, t = tid & 15;   // 0: row t (over j), 1: column t (over i)
            float mx = -INFINITY;
            for (int u = 0; u < 16; ++u) mx = fmaxf(mx, which == 0 ? sQ[t * LQ + u] : sQ[u * LQ + t]);
            float sum = 0.f;
            for (int u = 0; u < 16; ++u) sum += expf((which == 0 ? sQ[t * LQ + u] : sQ[u * LQ + t]) - mx);
            sStat[which * 32 + t] = mx;
            sStat[which * 32 + 16 + t] = sum;
        }
        __syncthreads();
        {
            const int i = tid >> 4, j = tid & 15;
            const float q = sQ[i * LQ + j];
            const float over_j = expf(q - sStat[i]) / sStat[16 + i];
            const float over_i = expf(q - sStat[32 + j]) / sStat[48 + j];
            __syncthreads();
            sQ[i * LQ + j] = over_j * over_i;
            sQT[j * LQ + i] = over_j * over_i;
        }
        __syncthreads();
        // mixes: thread = 4 rows x one channel quad, so that every row of the other branch read from shared memory feeds the four
        // rows at once (A'_r += Q[r][u] Aw_u, Aw'_r += Q[u][r] A_u)
        for (int e = tid; e < 4 * q4; e += 256) {
            const int r0 = (e / q4) * 4, c = (e % q4) * 4;
            float4 a[4], w[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                a[x] = *reinterpret_cast<const float4*>(sA + (r0 + x) * ld + c);
                w[x] = *reinterpret_cast<const float4*>(sW + (r0 + x) * ld + c);
            }
#pragma unroll 4
            for (int u = 0; u < 16; ++u) {
                const float4 ow = *reinterpret_cast<const float4*>(sW + u * ld + c);
                const float4 oa = *reinterpret_cast<const float4*>(sA + u * ld + c);
                const float4 qw4 = *reinterpret_cast<const float4*>(sQ + u * LQ + r0);    // Q[u][r0 .. r0 + 3]
                const float4 qa4 = *reinterpret_cast<const float4*>(sQT + u * LQ + r0);   // Q[r0 .. r0 + 3][u]
                const float qw[4] = {qw4.x, qw4.y, qw4.z, qw4.w}, qa[4] = {qa4.x, qa4.y, qa4.z, qa4.w};
                // packed fp32x2 FMAs: two channels per instruction, each lane of the pair rounded like a scalar fmaf
                const float2 ow01 = make_float2(ow.x, ow.y), ow23 = make_float2(ow.z, ow.w);
                const float2 oa01 = make_float2(oa.x, oa.y), oa23 = make_float2(oa.z, oa.w);
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const float2 qa2 = make_float2(qa[x], qa[x]), qw2 = make_float2(qw[x], qw[x]);
                    float2 t = __ffma2_rn(qa2, ow01, make_float2(a[x].x, a[x].y));
                    a[x].x = t.x; a[x].y = t.y;
                    t = __ffma2_rn(qa2, ow23, make_float2(a[x].z, a[x].w));
                    a[x].z = t.x; a[x].w = t.y;
                    t = __ffma2_rn(qw2, oa01, make_float2(w[x].x, w[x].y));
                    w[x].x = t.x; w[x].y = t.y;
                    t = __ffma2_rn(qw2, oa23, make_float2(w[x].z, w[x].w));
                    w[x].z = t.x; w[x].w = t.y;
                }
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                *reinterpret_cast<float4*>(Amix + base + (size_t)(r0 + x) * m + c) = a[x];
                *reinterpret_cast<float4*>(Awmix + base + (size_t)(r0 + x) * m + c) = w[x];
            }
        }
        __syncthreads();   // every thread is done with this buffer and with sQ before the next point reuses them
    }
}

extern "C" int ssf_attention_mix(const float* A, const float* Aw, long long n_points, int m, float* Amix, float* Awmix,
                                 void* stream) {
    if (n_points <= 0) return ssf_arg_error("attention_mix: empty input");
    if (m % 4 || m > 512) return ssf_arg_error("attention_mix: m must be a multiple of 4, <= 512");
    const size_t smem = (size_t)(2 * 2 * 16 * (m + 4) + 2 * 16 * 20 + 64) * sizeof(float);
    static unsigned long long attr = 0;
    if (ssf_attr_needed(&attr)) {
        cudaError_t e = cudaFuncSetAttribute(attention_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * 2 * 16 * (512 + 4) + 2 * 16 * 20 + 64) * sizeof(float)));
        if (e != cudaSuccess) return ssf_set_error(e);
        ssf_attr_done(&attr);
    }
    const long long per_sm = smem <= 72 * 1024 ? 3 : (smem <= 110 * 1024 ? 2 : 1);   // resident CTAs per SM (80 registers x 256 threads: at most 3)
    const long long n_cta = n_points < 148 * per_sm ? n_points : 148 * per_sm;
    attention_mix_kernel<<<(unsigned)n_cta, 256, smem, (cudaStream_t)stream>>>(A, Aw, m, n_points, Amix, Awmix);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// cost_fwd[b,n,:] = sum_s softmax_s(g[b,n,:])[s] * C[b,n,s,:]; also written channel-major [B, m, N1]
// one CTA (256 threads) per 8 consecutive points of a cloud
__global__ void __launch_bounds__(256) softmax_pool_kernel(const float* __restrict__ g, const float* __restrict__ C, int N1, int m,
                                                           float* __restrict__ out_pm, float* __restrict__ out_cm) {
    extern __shared__ __align__(16) float sm[];
    float* sWt = sm;            // [8][16] softmax weights
    float* sOut = sm + 128;     // [m][9]
    const int tid = threadIdx.x;
    const int b = blockIdx.y, n0 = blockIdx.x * 8;
    if (tid < 8) {
        const int n = n0 + tid;
        if (n < N1) {
            const float* gp = g + ((size_t)b * N1 + n) * 16;
            float mx = -INFINITY;
            for (int s = 0; s < 16; ++s) mx = fmaxf(mx, gp[s]);
            float sum = 0.f;
            for (int s = 0; s < 16; ++s) sum += expf(gp[s] - mx);
            for (int s = 0; s < 16; ++s) sWt[tid * 16 + s] = expf(gp[s] - mx) / sum;
        }
    }
    __syncthreads();
    for (int e = tid; e < 8 * m; e += 256) {
        const int p = e / m, c = e % m;
        const int n = n0 + p;
        float acc = 0.f;
        if (n < N1) {
            const float* cp = C + (((size_t)b * N1 + n) * 16) * m + c;
            for (int s = 0; s < 16; ++s) acc = fmaf(sWt[p * 16 + s], __ldg(cp + (size_t)s * m), acc);
            out_pm[((size_t)b * N1 + n) * m + c] = acc;
        }
        sOut[c * 9 + p] = acc;
    }
    __syncthreads();
    for (int e = tid; e < 8 * m; e += 256) {
        const int c = e >> 3, p = e & 7;
        if (n0 + p < N1) out_cm[((size_t)b * m + c) * N1 + n0 + p] = sOut[c * 9 + p];
    }
}

extern "C" int ssf_softmax_pool(const float* g, const float* C, int B, int N1, int m, float* out_pm, float* out_cm, void* stream) {
    if (B <= 0 || N1 <= 0 || m <= 0) return ssf_arg_error("softmax_pool: empty input");
    const size_t smem = (size_t)(128 + m * 9) * sizeof(float);
    softmax_pool_kernel<<<dim3((N1 + 7) / 8, B), 256, smem, (cudaStream_t)stream>>>(g, C, N1, m, out_pm, out_cm);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
