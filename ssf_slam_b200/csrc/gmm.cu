// The reference's noSeg dynamic-point masker on the device: a 2-component full-covariance Gaussian mixture fitted by EM on the
// 6-D features [flow | xyz], majority component = background (scripts/PointCloudOdometry_noSeg.py:97-103,
// ASF/main_sju_occ_ros.py:257-263: `GaussianMixture(n_components=2).fit_predict(...)`, `Counter(...).most_common(1)`).
// The arithmetic follows scikit-learn's published algorithm with its defaults (tol 1e-3 on the mean log-likelihood, reg_covar
// 1e-6, <= 100 EM iterations, precision-Cholesky parameterisation, final E step for the labels); the k-means seeding, which
// the reference leaves to the global NumPy RNG, is defined deterministically.  The step-by-step specification, pinned against
// scikit-learn itself, is oracle/gmm.py; this kernel reproduces its labels, iteration count and lower bound.
//
// One CTA of 256 threads per cloud, everything in fp64 (features are 24 bytes per point and stay L1/L2 resident across the
// iterations).  Per EM iteration one pass over the points: E step and the M-step sums together.  The second moments are
// accumulated about the *current* means (sum r (x - m)(x - m)^T, corrected by the mean shift afterwards), so there is no
// cancellation and no second pass.  The 56 sums go through a fixed-order shuffle / shared-memory reduction; thread 0 does the
// 6 x 6 Cholesky factorisations.
#include "ssf_common.cuh"

namespace {

constexpr int GM_T = 256;
constexpr int GM_D = 6;
constexpr int GM_TRI = 21;            // upper triangle of a 6 x 6
constexpr int GM_NS = 28;             // sums per component: r, r d (6), r d d^T (21)
constexpr int GM_LLOYD_MAX = 30;
constexpr double GM_EPS10 = 10.0 * 2.220446049250313e-16;
constexpr double GM_LOG_2PI = 1.8378770664093453;

struct GmmShared {
    double mu[2][GM_D];        // current means (centred coordinates)
    double pc[2][GM_TRI];      // precision Cholesky factor, upper triangular, row-major packed (i <= j)
    double cst[2];             // -0.5 D log 2pi + log det PC + log w
    double centre[GM_D];       // column means of the raw features
    double part[GM_T / 32][2 * GM_NS + 2];
    double tot[2 * GM_NS + 2];
    int flag;
    int seed_lo, seed_hi;
};

__device__ __forceinline__ int tri(int i, int j) { return i * GM_D - (i * (i - 1)) / 2 + (j - i); }   // i <= j

__device__ __forceinline__ void load_x(const float* __restrict__ P, const float* __restrict__ F, int i, const double* centre, double (&x)[GM_D]) {
    x[0] = (double)F[3 * i] - centre[0];
    x[1] = (double)F[3 * i + 1] - centre[1];
    x[2] = (double)F[3 * i + 2] - centre[2];
    x[3] = (double)P[3 * i] - centre[3];
    x[4] = (double)P[3 * i + 1] - centre[4];
    x[5] = (double)P[3 * i + 2] - centre[5];
}

// CTA-wide sums of acc[0..n) in a fixed order; totals land in S.tot (valid for every thread after the call)
template <int NV>
__device__ __forceinline__ void block_sum(double (&acc)[NV], GmmShared& S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) S.part[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = S.part[0][threadIdx.x];
        for (int w = 1; w < GM_T / 32; ++w) t += S.part[w][threadIdx.x];
        S.tot[threadIdx.x] = t;
    }
    __syncthreads();
}

// (weights, means, covariances) of the M step from the sums about the shift points m_k = S.mu[k]; then the precision
// Cholesky factors and the per-component constants.  Thread 0 only.
__device__ void update_parameters(GmmShared& S) {
    double nk[2];
    double cov[2][GM_D][GM_D];
    for (int k = 0; k < 2; ++k) {
        const double* s = S.tot + k * GM_NS;
        const double r = s[0];
        nk[k] = r + GM_EPS10;
        double delta[GM_D], u[GM_D];
        for (int i = 0; i < GM_D; ++i) {
            u[i] = s[1 + i];
            delta[i] = u[i] / nk[k] - (GM_EPS10 / nk[k]) * S.mu[k][i];   // new mean - shift point: (u + r m) / nk - m
        }
        for (int i = 0; i < GM_D; ++i)
            for (int j = i; j < GM_D; ++j) {
                const double c = (s[7 + tri(i, j)] - delta[i] * u[j] - u[i] * delta[j] + r * delta[i] * delta[j]) / nk[k];
                cov[k][i][j] = cov[k][j][i] = c + (i == j ? 1e-6 : 0.0);
            }
        for (int i = 0; i < GM_D; ++i) S.mu[k][i] += delta[i];
    }
    const double wsum = nk[0] + nk[1];
    for (int k = 0; k < 2; ++k) {
        // cov = L L^T (lower); PC = L^-T (upper)
        double L[GM_D][GM_D];
        for (int i = 0; i < GM_D; ++i)
            for (int j = 0; j <= i; ++j) {
                double v = cov[k][i][j];
                for (int q = 0; q < j; ++q) v -= L[i][q] * L[j][q];
                L[i][j] = (i == j) ? sqrt(v) : v / L[j][j];
            }
        double Li[GM_D][GM_D];   // L^-1 (lower), by forward substitution on the identity
        for (int c = 0; c < GM_D; ++c)
            for (int i = 0; i < GM_D; ++i) {
                if (i < c) { Li[i][c] = 0.0; continue; }
                double v = (i == c) ? 1.0 : 0.0;
                for (int q = c; q < i; ++q) v -= L[i][q] * Li[q][c];
                Li[i][c] = v / L[i][i];
            }
        double logdet = 0.0;
        for (int i = 0; i < GM_D; ++i) {
            logdet += log(Li[i][i]);
            for (int j = i; j < GM_D; ++j) S.pc[k][tri(i, j)] = Li[j][i];
        }
        S.cst[k] = -0.5 * GM_D * GM_LOG_2PI + logdet + log(nk[k] / wsum);
    }
}

// log p(x, k) for both components; d[k] = x - mu_k is returned for the M-step sums
__device__ __forceinline__ void log_prob(const GmmShared& S, const double (&x)[GM_D], double (&d)[2][GM_D], double (&lp)[2]) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int i = 0; i < GM_D; ++i) d[k][i] = x[i] - S.mu[k][i];
        double q = 0.0;
#pragma unroll
        for (int j = 0; j < GM_D; ++j) {
            double y = 0.0;
#pragma unroll
            for (int i = 0; i <= j; ++i) y += d[k][i] * S.pc[k][tri(i, j)];
            q += y * y;
        }
        lp[k] = S.cst[k] - 0.5 * q;
    }
}

__global__ void __launch_bounds__(GM_T) gmm_mask_kernel(const float* __restrict__ points, const float* __restrict__ flow, int N,
                                                        int max_iter, double tol, unsigned char* __restrict__ mask_out,
                                                        double* __restrict__ info_out) {
    __shared__ GmmShared S;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* P = points + (size_t)b * N * 3;
    const float* F = flow + (size_t)b * N * 3;
    unsigned char* M = mask_out + (size_t)b * N;   // k-means labels, then the final mask; a thread only re-reads its own writes

    // ---- centre the features
    if (tid < GM_D) S.centre[tid] = 0.0;
    __syncthreads();
    {
        double acc[GM_D], x[GM_D];
        for (int i = 0; i < GM_D; ++i) acc[i] = 0.0;
        for (int i = tid; i < N; i += GM_T) {
            load_x(P, F, i, S.centre, x);
            for (int c = 0; c < GM_D; ++c) acc[c] += x[c];
        }
        block_sum<GM_D>(acc, S);
        if (tid < GM_D) S.centre[tid] = S.tot[tid] / (double)N;
        __syncthreads();
    }
    // ---- seeds of the 2-means: extreme points along the highest-variance feature (lowest index on ties)
    {
        double acc[GM_D], x[GM_D];
        for (int i = 0; i < GM_D; ++i) acc[i] = 0.0;
        for (int i = tid; i < N; i += GM_T) {
            load_x(P, F, i, S.centre, x);
            for (int c = 0; c < GM_D; ++c) acc[c] += x[c] * x[c];
        }
        block_sum<GM_D>(acc, S);
        int dstar = 0;
        for (int c = 1; c < GM_D; ++c)
            if (S.tot[c] > S.tot[dstar]) dstar = c;
        // argmin / argmax over the raw fp32 column (centring is monotone), packed as (ordered value, index)
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (int i = tid; i < N; i += GM_T) {
            const float v = dstar < 3 ? F[3 * i + dstar] : P[3 * i + dstar - 3];
            unsigned u = __float_as_uint(v);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // order-preserving map float -> unsigned
            const unsigned long long lo = ((unsigned long long)u << 32) | (unsigned)i;
            const unsigned long long hi = ((unsigned long long)u << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
            kmin = lo < kmin ? lo : kmin;
            kmax = hi > kmax ? hi : kmax;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), c = __shfl_xor_sync(0xffffffffu, kmax, o);
            kmin = a < kmin ? a : kmin;
            kmax = c > kmax ? c : kmax;
        }
        unsigned long long* red = reinterpret_cast<unsigned long long*>(&S.part[0][0]);
        __syncthreads();
        if ((tid & 31) == 0) {
            red[tid >> 5] = kmin;
            red[8 + (tid >> 5)] = kmax;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < GM_T / 32; ++w) {
                kmin = red[w] < kmin ? red[w] : kmin;
                kmax = red[8 + w] > kmax ? red[8 + w] : kmax;
            }
            S.seed_lo = (int)(unsigned)(kmin & 0xFFFFFFFFull);
            S.seed_hi = (int)(0xFFFFFFFFu - (unsigned)(kmax & 0xFFFFFFFFull));
        }
        __syncthreads();
        if (tid < 2) {
            double x2[GM_D];
            load_x(P, F, tid == 0 ? S.seed_lo : S.seed_hi, S.centre, x2);
            for (int c = 0; c < GM_D; ++c) S.mu[tid][c] = x2[c];
        }
        __syncthreads();
    }
    // ---- Lloyd iterations (labels in M; ties -> cluster 0; stop when no label changes)
    for (int itl = 0; itl < GM_LLOYD_MAX; ++itl) {
        double acc[15], x[GM_D];   // per cluster: count, sum x (6); [14] = changed labels
        for (int i = 0; i < 15; ++i) acc[i] = 0.0;
        for (int i = tid; i < N; i += GM_T) {
            load_x(P, F, i, S.centre, x);
            double d0 = 0.0, d1 = 0.0;
            for (int c = 0; c < GM_D; ++c) {
                const double a = x[c] - S.mu[0][c], e = x[c] - S.mu[1][c];
                d0 += a * a;
                d1 += e * e;
            }
            const int lab = d1 < d0 ? 1 : 0;
            if (itl == 0 || M[i] != lab) acc[14] += 1.0;
            M[i] = (unsigned char)lab;
            acc[lab * 7] += 1.0;
            for (int c = 0; c < GM_D; ++c) acc[lab * 7 + 1 + c] += x[c];
        }
        block_sum<15>(acc, S);
        const bool stop = itl > 0 && S.tot[14] == 0.0;
        __syncthreads();
        if (stop) break;
        if (tid < 2 && S.tot[tid * 7] > 0.0)
            for (int c = 0; c < GM_D; ++c) S.mu[tid][c] = S.tot[tid * 7 + 1 + c] / S.tot[tid * 7];
        __syncthreads();
    }
    // ---- initial parameters from the one-hot responsibilities
    double acc[2 * GM_NS + 1];
    {
        for (int i = 0; i < 2 * GM_NS; ++i) acc[i] = 0.0;
        double x[GM_D];
        for (int i = tid; i < N; i += GM_T) {
            load_x(P, F, i, S.centre, x);
            const int k = M[i];
            double d[GM_D];
            for (int c = 0; c < GM_D; ++c) d[c] = x[c] - S.mu[k][c];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                if (kk != k) continue;
                acc[kk * GM_NS] += 1.0;
#pragma unroll
                for (int c = 0; c < GM_D; ++c) acc[kk * GM_NS + 1 + c] += d[c];
#pragma unroll
                for (int c = 0; c < GM_D; ++c)
#pragma unroll
                    for (int e = c; e < GM_D; ++e) acc[kk * GM_NS + 7 + tri(c, e)] += d[c] * d[e];
            }
        }
        block_sum<2 * GM_NS>(reinterpret_cast<double(&)[2 * GM_NS]>(acc), S);
        if (tid == 0) update_parameters(S);
        __syncthreads();
    }
    // ---- EM
    double lb = -INFINITY;
    int n_iter = 0, converged = 0;
    for (n_iter = 1; n_iter <= max_iter; ++n_iter) {
        for (int i = 0; i < 2 * GM_NS + 1; ++i) acc[i] = 0.0;
        double x[GM_D], d[2][GM_D], lp[2];
        for (int i = tid; i < N; i += GM_T) {
            load_x(P, F, i, S.centre, x);
            log_prob(S, x, d, lp);
            const double mx = fmax(lp[0], lp[1]);
            const double lse = mx + log(exp(lp[0] - mx) + exp(lp[1] - mx));
            acc[2 * GM_NS] += lse;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double r = exp(lp[k] - lse);
                acc[k * GM_NS] += r;
#pragma unroll
                for (int c = 0; c < GM_D; ++c) {
                    const double rd = r * d[k][c];
                    acc[k * GM_NS + 1 + c] += rd;
#pragma unroll
                    for (int e = c; e < GM_D; ++e) acc[k * GM_NS + 7 + tri(c, e)] += rd * d[k][e];
                }
            }
        }
        block_sum<2 * GM_NS + 1>(acc, S);
        const double prev = lb;
        lb = S.tot[2 * GM_NS] / (double)N;
        __syncthreads();
        if (tid == 0) update_parameters(S);
        __syncthreads();
        if (fabs(lb - prev) < tol) {
            converged = 1;
            break;
        }
    }
    if (n_iter > max_iter) n_iter = max_iter;
    // ---- final E step: labels, majority component = background (ties -> the label of point 0, as Counter.most_common)
    {
        double cnt[1] = {0.0};
        double x[GM_D], d[2][GM_D], lp[2];
        for (int i = tid; i < N; i += GM_T) {
            load_x(P, F, i, S.centre, x);
            log_prob(S, x, d, lp);
            const int lab = lp[1] > lp[0] ? 1 : 0;
            M[i] = (unsigned char)lab;
            cnt[0] += (double)lab;
        }
        block_sum<1>(cnt, S);
        const int c1 = (int)S.tot[0], c0 = N - c1;
        if (tid == 0) S.flag = M[0];
        __syncthreads();
        const int bg = c1 > c0 ? 1 : (c1 < c0 ? 0 : S.flag);
        for (int i = tid; i < N; i += GM_T) M[i] = (M[i] != bg) ? 1 : 0;
        if (tid == 0 && info_out != nullptr) {
            double* o = info_out + (size_t)b * 4;
            o[0] = (double)n_iter;
            o[1] = lb;
            o[2] = (double)converged;
            o[3] = (double)(bg ? c1 : c0);
        }
    }
}

}  // namespace

// points, flow [B,N,3] f32 -> mask u8 [B,N] (0 = background component, 1 = the other), info f64 [B,4] =
// (EM iterations, lower bound of the last E step inside the loop, converged, background count); info may be null
extern "C" int ssf_gmm_mask(const float* points, const float* flow, int B, int N, int max_iter, double tol,
                            unsigned char* mask_out, double* info_out, void* stream) {
    if (B <= 0 || N < 2) return ssf_arg_error("gmm_mask: needs B >= 1 clouds of N >= 2 points");
    if (max_iter <= 0) return ssf_arg_error("gmm_mask: max_iter must be positive");
    gmm_mask_kernel<<<B, GM_T, 0, (cudaStream_t)stream>>>(points, flow, N, max_iter, tol, mask_out, info_out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
