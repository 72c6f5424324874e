// CUDA-core pieces of the un-fused (wide, m >= 128) PointConvTransFlowV2 path: the S x S cross attention between the
// two branches (ASF/utils/soflow.py:420-422,453-458) and the softmax-weighted forward cost (:469,486).  The dense
// layers around them run on tcgen05 (dense_tc.cu); at levels 2 and 3 a cloud has only 512 / 256 query points, so the
// activations exchanged through HBM are a few MB per cloud.
#include "ssf_common.cuh"

// A, Aw [P, 16, m] -> Amix = A + Q.Aw, Awmix = Aw + Q^T.A with Q = softmax_i(<A_i,Aw_j>) * softmax_j(<A_i,Aw_j>)
// one CTA (256 threads) per point
__global__ void __launch_bounds__(256) attention_mix_kernel(const float* __restrict__ A, const float* __restrict__ Aw, int m,
                                                            float* __restrict__ Amix, float* __restrict__ Awmix) {
    extern __shared__ __align__(16) float sm[];
    const int ld = m + 4;
    float* sA = sm;                 // [16][ld]
    float* sW = sA + 16 * ld;       // [16][ld]
    float* sQ = sW + 16 * ld;       // [16][17]
    float* sStat = sQ + 16 * 17;    // row max | row sum | col max | col sum, [16] each
    const int tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * 16 * m;
    const int q4 = m >> 2;
    for (int e = tid; e < 16 * q4; e += 256) {
        const int r = e / q4, c = (e % q4) * 4;
        *reinterpret_cast<float4*>(sA + r * ld + c) = __ldg(reinterpret_cast<const float4*>(A + base + (size_t)r * m + c));
        *reinterpret_cast<float4*>(sW + r * ld + c) = __ldg(reinterpret_cast<const float4*>(Aw + base + (size_t)r * m + c));
    }
    __syncthreads();
    {
        const int i = tid >> 4, j = tid & 15;
        const float* ar = sA + i * ld;
        const float* wr = sW + j * ld;
        float s = 0.f;
        for (int c = 0; c < m; c += 4) {
            const float4 x = *reinterpret_cast<const float4*>(ar + c);
            const float4 y = *reinterpret_cast<const float4*>(wr + c);
            s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
        }
        sQ[i * 17 + j] = s;
    }
    __syncthreads();
    if (tid < 32) {
        const int which = tid >> 4, t = tid & 15;   // 0: row t (over j), 1: column t (over i)
        float mx = -INFINITY;
        for (int u = 0; u < 16; ++u) mx = fmaxf(mx, which == 0 ? sQ[t * 17 + u] : sQ[u * 17 + t]);
        float sum = 0.f;
        for (int u = 0; u < 16; ++u) sum += expf((which == 0 ? sQ[t * 17 + u] : sQ[u * 17 + t]) - mx);
        sStat[which * 32 + t] = mx;
        sStat[which * 32 + 16 + t] = sum;
    }
    __syncthreads();
    {
        const int i = tid >> 4, j = tid & 15;
        const float q = sQ[i * 17 + j];
        const float over_j = expf(q - sStat[i]) / sStat[16 + i];
        const float over_i = expf(q - sStat[32 + j]) / sStat[48 + j];
        __syncthreads();
        sQ[i * 17 + j] = over_j * over_i;
    }
    __syncthreads();
    for (int e = tid; e < 16 * q4; e += 256) {
        const int r = e / q4, c = (e % q4) * 4;
        float4 a = *reinterpret_cast<const float4*>(sA + r * ld + c);
        float4 w = *reinterpret_cast<const float4*>(sW + r * ld + c);
#pragma unroll 4
        for (int u = 0; u < 16; ++u) {
            const float qa = sQ[r * 17 + u];   // Q[r][u]  : A'_r  += Q[r][u] * Aw_u
            const float qw = sQ[u * 17 + r];   // Q[u][r]  : Aw'_r += Q[u][r] * A_u
            const float4 ow = *reinterpret_cast<const float4*>(sW + u * ld + c);
            const float4 oa = *reinterpret_cast<const float4*>(sA + u * ld + c);
            a.x = fmaf(qa, ow.x, a.x); a.y = fmaf(qa, ow.y, a.y); a.z = fmaf(qa, ow.z, a.z); a.w = fmaf(qa, ow.w, a.w);
            w.x = fmaf(qw, oa.x, w.x); w.y = fmaf(qw, oa.y, w.y); w.z = fmaf(qw, oa.z, w.z); w.w = fmaf(qw, oa.w, w.w);
        }
        *reinterpret_cast<float4*>(Amix + base + (size_t)r * m + c) = a;
        *reinterpret_cast<float4*>(Awmix + base + (size_t)r * m + c) = w;
    }
}

extern "C" int ssf_attention_mix(const float* A, const float* Aw, long long n_points, int m, float* Amix, float* Awmix,
                                 void* stream) {
    if (n_points <= 0) return ssf_arg_error("attention_mix: empty input");
    if (m % 4 || m > 512) return ssf_arg_error("attention_mix: m must be a multiple of 4, <= 512");
    const size_t smem = (size_t)(2 * 16 * (m + 4) + 16 * 17 + 64) * sizeof(float);
    attention_mix_kernel<<<(unsigned)n_points, 256, smem, (cudaStream_t)stream>>>(A, Aw, m, Amix, Awmix);
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
