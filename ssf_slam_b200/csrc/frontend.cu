// Per-frame front end after the scene-flow network: dynamic-point masking and static-point ego-motion.
//
// Replaces, on the device and in one launch per batch of clouds, the host code at
// ASF/main_sju_occ_ros.py:257-284 / scripts/PointCloudOdometry.py:91-101 (mask -> bg_index ->
// slove_RT_by_SVD -> quaternion).  The masker is the deterministic residual-flow-versus-rigid-flow test
// with per-instance voting specified in DESIGN.md (anchors: ASF/calc_coarse_flow.py:587-590,
// scripts/PointCloudOdometry.py:93-96); the bit-level specification is oracle/frontend.py::masker_spec,
// and every reduction order / rounding below mirrors it so masks are bit-exact:
//   * 16 fp64 sums per fit (W, sum a, sum b, sum a b^T), thread t adds points t, t+256, ... in order,
//     xor-butterfly across lanes, warp sums added in warp order by thread 0
//   * Horn's closed form: max eigenvector of the symmetric 4x4 by cyclic Jacobi, 12 fixed sweeps, fp64,
//     all products/sums individually rounded (__dmul_rn/__dadd_rn) so no FMA contraction can differ
//   * residual in fp32 with __fmul_rn/__fadd_rn, threshold fl32(fl32(tau)*m)^2, m in {8,4,2,1}
// One CTA of 256 threads per cloud; clouds of a batch run on different SMs.
#include "ssf_common.cuh"

constexpr int FE_T = 256;
constexpr int FE_MAX_INST = 4096;
constexpr int FE_SWEEPS = 12;

struct Pose {
    double q[4];  // w, x, y, z (w >= 0)
    double R[9];
    double t[3];
};

__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }

__device__ void pose_from_sums(const double* s, Pose& P) {
    const double W = s[0];
    if (!(W >= 3.0)) {
        P.q[0] = 1.0; P.q[1] = P.q[2] = P.q[3] = 0.0;
        for (int i = 0; i < 9; ++i) P.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        P.t[0] = P.t[1] = P.t[2] = 0.0;
        return;
    }
    double ma[3], mb[3], S[3][3];
    for (int i = 0; i < 3; ++i) {
        ma[i] = __ddiv_rn(s[1 + i], W);
        mb[i] = __ddiv_rn(s[4 + i], W);
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S[i][j] = xsub(s[7 + 3 * i + j], xmul(W, xmul(ma[i], mb[j])));
    const double Sxx = S[0][0], Sxy = S[0][1], Sxz = S[0][2];
    const double Syx = S[1][0], Syy = S[1][1], Syz = S[1][2];
    const double Szx = S[2][0], Szy = S[2][1], Szz = S[2][2];
    double A[4][4], V[4][4];
    A[0][0] = xadd(xadd(Sxx, Syy), Szz);
    A[0][1] = xsub(Syz, Szy);
    A[0][2] = xsub(Szx, Sxz);
    A[0][3] = xsub(Sxy, Syx);
    A[1][0] = A[0][1];
    A[1][1] = xsub(xsub(Sxx, Syy), Szz);
    A[1][2] = xadd(Sxy, Syx);
    A[1][3] = xadd(Szx, Sxz);
    A[2][0] = A[0][2];
    A[2][1] = A[1][2];
    A[2][2] = xsub(xsub(Syy, Sxx), Szz);
    A[2][3] = xadd(Syz, Szy);
    A[3][0] = A[0][3];
    A[3][1] = A[1][3];
    A[3][2] = A[2][3];
    A[3][3] = xsub(xsub(Szz, Sxx), Syy);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < FE_SWEEPS; ++sweep) {
        for (int p = 0; p < 3; ++p) {
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = __ddiv_rn(xsub(A[q][q], A[p][p]), xmul(2.0, apq));
                double t = __ddiv_rn(1.0, xadd(fabs(theta), __dsqrt_rn(xadd(xmul(theta, theta), 1.0))));
                if (theta < 0.0) t = -t;
                const double c = __ddiv_rn(1.0, __dsqrt_rn(xadd(xmul(t, t), 1.0)));
                const double sn = xmul(t, c);
                for (int k = 0; k < 4; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = xsub(xmul(c, akp), xmul(sn, akq));
                    A[k][q] = xadd(xmul(sn, akp), xmul(c, akq));
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = xsub(xmul(c, apk), xmul(sn, aqk));
                    A[q][k] = xadd(xmul(sn, apk), xmul(c, aqk));
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = xsub(xmul(c, vkp), xmul(sn, vkq));
                    V[k][q] = xadd(xmul(sn, vkp), xmul(c, vkq));
                }
            }
        }
    }
    int kk = 0;
    for (int i = 1; i < 4; ++i)
        if (A[i][i] > A[kk][kk]) kk = i;
    double q0 = V[0][kk], q1 = V[1][kk], q2 = V[2][kk], q3 = V[3][kk];
    const double n = __dsqrt_rn(xadd(xadd(xadd(xmul(q0, q0), xmul(q1, q1)), xmul(q2, q2)), xmul(q3, q3)));
    q0 = __ddiv_rn(q0, n); q1 = __ddiv_rn(q1, n); q2 = __ddiv_rn(q2, n); q3 = __ddiv_rn(q3, n);
    if (q0 < 0.0) { q0 = -q0; q1 = -q1; q2 = -q2; q3 = -q3; }
    const double w = q0, x = q1, y = q2, z = q3;
    P.q[0] = w; P.q[1] = x; P.q[2] = y; P.q[3] = z;
    P.R[0] = xsub(1.0, xmul(2.0, xadd(xmul(y, y), xmul(z, z))));
    P.R[1] = xmul(2.0, xsub(xmul(x, y), xmul(w, z)));
    P.R[2] = xmul(2.0, xadd(xmul(x, z), xmul(w, y)));
    P.R[3] = xmul(2.0, xadd(xmul(x, y), xmul(w, z)));
    P.R[4] = xsub(1.0, xmul(2.0, xadd(xmul(x, x), xmul(z, z))));
    P.R[5] = xmul(2.0, xsub(xmul(y, z), xmul(w, x)));
    P.R[6] = xmul(2.0, xsub(xmul(x, z), xmul(w, y)));
    P.R[7] = xmul(2.0, xadd(xmul(y, z), xmul(w, x)));
    P.R[8] = xsub(1.0, xmul(2.0, xadd(xmul(x, x), xmul(y, y))));
    for (int i = 0; i < 3; ++i)
        P.t[i] = xsub(mb[i], xadd(xadd(xmul(P.R[3 * i], ma[0]), xmul(P.R[3 * i + 1], ma[1])), xmul(P.R[3 * i + 2], ma[2])));
}

// CTA-wide sum of 16 doubles per thread in the order the oracle mirrors; result valid in thread 0.
__device__ __forceinline__ void block_sum16(double (&acc)[16], double (*s_part)[16]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = xadd(v, __shfl_xor_sync(0xffffffffu, v, o));
        acc[k] = v;
    }
    if (lane == 0)
        for (int k = 0; k < 16; ++k) s_part[warp][k] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 16; ++k) {
            double tot = s_part[0][k];
            for (int w = 1; w < FE_T / 32; ++w) tot = xadd(tot, s_part[w][k]);
            acc[k] = tot;
        }
    }
}

__device__ __forceinline__ void accumulate_point(double (&acc)[16], float px, float py, float pz, float fx, float fy, float fz,
                                                 bool two_clouds) {
    // a = p + f (fp32 add, as the reference forms `points + flow`), b = p; two_clouds: a = f given explicitly
    const double a0 = (double)(two_clouds ? fx : __fadd_rn(px, fx)), a1 = (double)(two_clouds ? fy : __fadd_rn(py, fy)),
                 a2 = (double)(two_clouds ? fz : __fadd_rn(pz, fz));
    const double b0 = (double)px, b1 = (double)py, b2 = (double)pz;
    acc[0] = xadd(acc[0], 1.0);
    acc[1] = xadd(acc[1], a0); acc[2] = xadd(acc[2], a1); acc[3] = xadd(acc[3], a2);
    acc[4] = xadd(acc[4], b0); acc[5] = xadd(acc[5], b1); acc[6] = xadd(acc[6], b2);
    acc[7] = xadd(acc[7], xmul(a0, b0)); acc[8] = xadd(acc[8], xmul(a0, b1)); acc[9] = xadd(acc[9], xmul(a0, b2));
    acc[10] = xadd(acc[10], xmul(a1, b0)); acc[11] = xadd(acc[11], xmul(a1, b1)); acc[12] = xadd(acc[12], xmul(a1, b2));
    acc[13] = xadd(acc[13], xmul(a2, b0)); acc[14] = xadd(acc[14], xmul(a2, b1)); acc[15] = xadd(acc[15], xmul(a2, b2));
}

__device__ __forceinline__ bool residual_dynamic(const float* R, const float* t, float px, float py, float pz, float fx,
                                                 float fy, float fz, float tau2) {
    const float qx = __fadd_rn(px, fx), qy = __fadd_rn(py, fy), qz = __fadd_rn(pz, fz);
    const float x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], qx), __fmul_rn(R[1], qy)), __fmul_rn(R[2], qz)), t[0]);
    const float y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], qx), __fmul_rn(R[4], qy)), __fmul_rn(R[5], qz)), t[1]);
    const float z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], qx), __fmul_rn(R[7], qy)), __fmul_rn(R[8], qz)), t[2]);
    const float dx = __fsub_rn(x, px), dy = __fsub_rn(y, py), dz = __fsub_rn(z, pz);
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return r2 > tau2;
}

// mode 0: Kabsch only with the given static weight (w = [weight_mask == 0]); mask_out = copy of the input mask
// mode 2: as mode 0 but `flow` holds the source cloud itself (fit points ~= R*flow + t): plain slove_RT_by_SVD(src, dst)
// mode 1: residual masker (+ optional sem seed / instance voting), then final Kabsch on the static set
__global__ void __launch_bounds__(FE_T)
frontend_kernel(const float* __restrict__ points, const float* __restrict__ flow, int N, int mode,
                const unsigned char* __restrict__ in_mask, const int* __restrict__ sem, unsigned long long movable_bits,
                const int* __restrict__ inst, int n_inst, float tau, unsigned char* __restrict__ mask_out,
                double* __restrict__ odom_out, double* __restrict__ pose_out) {
    __shared__ double s_part[FE_T / 32][16];
    __shared__ float s_R[9];
    __shared__ float s_t[3];
    __shared__ int s_flag;
    __shared__ int s_cnt[FE_MAX_INST];
    __shared__ int s_dyn[FE_MAX_INST];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* P = points + (size_t)b * N * 3;
    const float* F = flow + (size_t)b * N * 3;
    unsigned char* M = mask_out + (size_t)b * N;
    const unsigned char* IM = in_mask ? in_mask + (size_t)b * N : nullptr;
    const int* SEM = sem ? sem + (size_t)b * N : nullptr;
    const int* INST = inst ? inst + (size_t)b * N : nullptr;
    Pose pose;
    double acc[16];

    // ---- round-0 static set
    if (mode == 1 && SEM != nullptr && movable_bits != 0ull) {
        int c = 0;
        for (int i = tid; i < N; i += FE_T) {
            const int s = SEM[i];
            c += !(s >= 0 && s < 64 && ((movable_bits >> s) & 1ull));
        }
        if (tid == 0) s_flag = 0;
        __syncthreads();
        atomicAdd(&s_flag, c);
        __syncthreads();
        const bool use_sem = s_flag >= 3;
        for (int i = tid; i < N; i += FE_T) {
            const int s = SEM[i];
            const bool mov = (s >= 0 && s < 64 && ((movable_bits >> s) & 1ull));
            M[i] = use_sem ? (mov ? 1 : 0) : 0;
        }
    } else if (mode == 1) {
        for (int i = tid; i < N; i += FE_T) M[i] = 0;
    } else {
        for (int i = tid; i < N; i += FE_T) M[i] = IM ? (IM[i] != 0 ? 1 : 0) : 0;
    }
    __syncthreads();

    const int rounds = mode == 1 ? 4 : 0;
    for (int r = 0; r <= rounds; ++r) {
        // fit on the current static set (mask == 0); each thread reads back only its own writes to M
        for (int k = 0; k < 16; ++k) acc[k] = 0.0;
        for (int i = tid; i < N; i += FE_T) {
            if (M[i] == 0) accumulate_point(acc, P[3 * i], P[3 * i + 1], P[3 * i + 2], F[3 * i], F[3 * i + 1], F[3 * i + 2], mode == 2);
        }
        block_sum16(acc, s_part);
        if (tid == 0) {
            pose_from_sums(acc, pose);
            for (int k = 0; k < 9; ++k) s_R[k] = (float)pose.R[k];
            for (int k = 0; k < 3; ++k) s_t[k] = (float)pose.t[k];
        }
        __syncthreads();
        if (r == rounds) break;
        const float mult = r == 0 ? 8.f : (r == 1 ? 4.f : (r == 2 ? 2.f : 1.f));
        const float tau_r = __fmul_rn(tau, mult);
        const float tau2 = __fmul_rn(tau_r, tau_r);
        float Rr[9], tr[3];
        for (int k = 0; k < 9; ++k) Rr[k] = s_R[k];
        for (int k = 0; k < 3; ++k) tr[k] = s_t[k];
        const bool vote = (r == rounds - 1) && INST != nullptr;
        if (vote) {
            for (int k = tid; k < n_inst; k += FE_T) {
                s_cnt[k] = 0;
                s_dyn[k] = 0;
            }
            __syncthreads();
        }
        for (int i = tid; i < N; i += FE_T) {
            const bool dyn = residual_dynamic(Rr, tr, P[3 * i], P[3 * i + 1], P[3 * i + 2], F[3 * i], F[3 * i + 1], F[3 * i + 2], tau2);
            M[i] = dyn ? 1 : 0;
            if (vote) {
                const int id = INST[i];
                if (id >= 1 && id < n_inst) {
                    atomicAdd(&s_cnt[id], 1);
                    if (dyn) atomicAdd(&s_dyn[id], 1);
                }
            }
        }
        if (vote) {
            __syncthreads();
            for (int i = tid; i < N; i += FE_T) {
                const int id = INST[i];
                if (id >= 1 && id < n_inst) M[i] = (2 * s_dyn[id] > s_cnt[id]) ? 1 : 0;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        double* o = odom_out + (size_t)b * 7;
        o[0] = pose.t[0]; o[1] = pose.t[1]; o[2] = pose.t[2];
        o[3] = pose.q[1]; o[4] = pose.q[2]; o[5] = pose.q[3]; o[6] = pose.q[0];
        if (pose_out != nullptr) {
            double* p = pose_out + (size_t)b * 12;
            for (int k = 0; k < 9; ++k) p[k] = pose.R[k];
            for (int k = 0; k < 3; ++k) p[9 + k] = pose.t[k];
        }
    }
}

// slove_RT_by_SVD(src, dst) on float64 clouds, in the input precision like the reference (scripts/PointCloudOdometry.py:15-33
// runs numpy in the dtype it is given): the same 16 fp64 sums and Horn closed form as above, fed with doubles.
__global__ void __launch_bounds__(FE_T)
solve_rt_f64_kernel(const double* __restrict__ src, const double* __restrict__ dst, int M, double* __restrict__ odom_out,
                    double* __restrict__ pose_out) {
    __shared__ double s_part[FE_T / 32][16];
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* A = src + (size_t)b * M * 3;
    const double* Bp = dst + (size_t)b * M * 3;
    double acc[16];
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (int i = tid; i < M; i += FE_T) {
        const double a0 = A[3 * i], a1 = A[3 * i + 1], a2 = A[3 * i + 2];
        const double b0 = Bp[3 * i], b1 = Bp[3 * i + 1], b2 = Bp[3 * i + 2];
        acc[0] = xadd(acc[0], 1.0);
        acc[1] = xadd(acc[1], a0); acc[2] = xadd(acc[2], a1); acc[3] = xadd(acc[3], a2);
        acc[4] = xadd(acc[4], b0); acc[5] = xadd(acc[5], b1); acc[6] = xadd(acc[6], b2);
        acc[7] = xadd(acc[7], xmul(a0, b0)); acc[8] = xadd(acc[8], xmul(a0, b1)); acc[9] = xadd(acc[9], xmul(a0, b2));
        acc[10] = xadd(acc[10], xmul(a1, b0)); acc[11] = xadd(acc[11], xmul(a1, b1)); acc[12] = xadd(acc[12], xmul(a1, b2));
        acc[13] = xadd(acc[13], xmul(a2, b0)); acc[14] = xadd(acc[14], xmul(a2, b1)); acc[15] = xadd(acc[15], xmul(a2, b2));
    }
    block_sum16(acc, s_part);
    if (tid == 0) {
        Pose pose;
        pose_from_sums(acc, pose);
        double* o = odom_out + (size_t)b * 7;
        o[0] = pose.t[0]; o[1] = pose.t[1]; o[2] = pose.t[2];
        o[3] = pose.q[1]; o[4] = pose.q[2]; o[5] = pose.q[3]; o[6] = pose.q[0];
        double* p = pose_out + (size_t)b * 12;
        for (int k = 0; k < 9; ++k) p[k] = pose.R[k];
        for (int k = 0; k < 3; ++k) p[9 + k] = pose.t[k];
    }
}

extern "C" int ssf_solve_rt_f64(const double* src, const double* dst, int B, int M, double* odom_out, double* pose_out,
                                void* stream) {
    if (B <= 0 || M <= 0) return ssf_arg_error("solve_rt_f64: empty input");
    if (odom_out == nullptr || pose_out == nullptr) return ssf_arg_error("solve_rt_f64: outputs must not be NULL");
    solve_rt_f64_kernel<<<B, FE_T, 0, (cudaStream_t)stream>>>(src, dst, M, odom_out, pose_out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// points, flow [B,N,3] f32; optional in_mask u8 [B,N] (mode 0), sem / inst i32 [B,N]
// -> mask u8 [B,N], odom f64 [B,7] = [tx,ty,tz,qx,qy,qz,qw], pose f64 [B,12] = [R row-major, t] (may be null)
extern "C" int ssf_frontend(const float* points, const float* flow, int B, int N, int mode, const unsigned char* in_mask,
                            const int* sem, unsigned long long movable_bits, const int* inst, int n_inst, float tau,
                            unsigned char* mask_out, double* odom_out, double* pose_out, void* stream) {
    if (B <= 0 || N <= 0) return ssf_arg_error("frontend: empty input");
    if (mode < 0 || mode > 2) return ssf_arg_error("frontend: mode must be 0 (kabsch), 1 (masker) or 2 (two-cloud kabsch)");
    if (inst != nullptr && (n_inst <= 0 || n_inst > FE_MAX_INST)) return ssf_arg_error("frontend: n_inst must be in [1,4096]");
    frontend_kernel<<<B, FE_T, 0, (cudaStream_t)stream>>>(points, flow, N, mode, in_mask, sem, movable_bits, inst, n_inst,
                                                           tau, mask_out, odom_out, pose_out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
