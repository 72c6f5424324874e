// Bring-up / regression kernel for the tcgen05 path: Y[128,N] = X[128,K] . W[N,K]^T with the 3xTF32 scheme,
// exercising exactly the mechanisms the fused layers use (TMEM alloc, tcgen05.st of split activations, A-from-TMEM
// and A-from-smem MMAs against pre-arranged weight images, commit -> mbarrier, tcgen05.ld epilogue).
#include "../tc_common.cuh"

// mode 0: A operand from TMEM; mode 1: A operand from shared memory (no-swizzle K-major image written by the threads)
// passes 1: hi.hi only (plain TF32); 3: full 3xTF32
__global__ void __launch_bounds__(128) tc_gemm_test_kernel(const float* __restrict__ X, const float* __restrict__ Whi,
                                                           const float* __restrict__ Wlo, int K, int N, int mode, int passes,
                                                           float* __restrict__ Y) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* sWhi = reinterpret_cast<float*>(smem);
    float* sWlo = sWhi + N * K;
    float* sAhi = sWlo + N * K;    // only mode 1
    float* sAlo = sAhi + 128 * K;

    if (warp == 0) tc_alloc(&tmem_slot, 512);
    if (tid == 0) {
        ssf_mbar_init(&bar_w, 1);
        ssf_mbar_init(&bar_mma, 1);
        ssf_mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(N * K * 4);
        ssf_mbar_expect_tx(&bar_w, 2 * bytes);
        ssf_bulk_g2s(sWhi, Whi, bytes, &bar_w);
        ssf_bulk_g2s(sWlo, Wlo, bytes, &bar_w);
    }
    // activations: thread t owns row t
    const float* xr = X + (size_t)tid * K;
    for (int k0 = 0; k0 < K; k0 += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tc_split(xr[k0 + j], hi[j], lo[j]);
        if (mode == 0) {
            tc_st8(tc_addr(tmem, warp, k0), hi);
            tc_st8(tc_addr(tmem, warp, K + k0), lo);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sAhi[tc_img_off(tid, k0 + j, 128) >> 2] = hi[j];
                sAlo[tc_img_off(tid, k0 + j, 128) >> 2] = lo[j];
            }
        }
    }
    if (mode == 0) tc_st_wait();
    else fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    const int dcol = 256;
    if (tid == 0) {
        ssf_mbar_wait(&bar_w, 0);
        tc_fence_after();
        const uint32_t idesc = tc_idesc_tf32(128, N);
        const uint32_t lboW = (uint32_t)(N / 8) * 128, lboA = 16 * 128;
        uint32_t acc = 0;
        for (int pass = (passes == 3 ? 0 : 2); pass < 3; ++pass) {  // 0: Alo.Whi, 1: Ahi.Wlo, 2: Ahi.Whi
            const bool a_lo = pass == 0, w_lo = pass == 1;
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t bdesc = tc_smem_desc(ssf_smem_u32(w_lo ? sWlo : sWhi) + ks * 2 * lboW, lboW, 128);
                if (mode == 0) {
                    tc_mma_ts(tmem + dcol, tmem + (a_lo ? K : 0) + ks * 8, bdesc, idesc, acc);
                } else {
                    const uint64_t adesc = tc_smem_desc(ssf_smem_u32(a_lo ? sAlo : sAhi) + ks * 2 * lboA, lboA, 128);
                    tc_mma_ss(tmem + dcol, adesc, bdesc, idesc, acc);
                }
                acc = 1;
            }
        }
        tc_commit(&bar_mma);
    }
    ssf_mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int n0 = 0; n0 < N; n0 += 8) {
        float v[8];
        tc_ld8(tc_addr(tmem, warp, dcol + n0), v);
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) Y[(size_t)tid * N + n0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, 512);
}

extern "C" int ssf_tc_gemm_test(const float* X, const float* Whi_img, const float* Wlo_img, int K, int N, int mode, int passes,
                                float* Y, void* stream) {
    if (K % 8 || N % 16 || N > 256 || K > 128 || N < 16) return ssf_arg_error("tc_gemm_test: need K%8==0, K<=128, N%16==0, N<=256");
    const size_t smem = (size_t)(2 * N * K + (mode == 1 ? 2 * 128 * K : 0)) * 4;
    if (smem > 200 * 1024) return ssf_arg_error("tc_gemm_test: operands do not fit in shared memory");
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ssf_set_error(e);
    tc_gemm_test_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(X, Whi_img, Wlo_img, K, N, mode, passes, Y);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// ---- tensor-pipe pacing probe: `reps` back-to-back kind::tf32 MMAs (M128 x N x K8) from one thread, operands zero.
// mode 0: A from TMEM, 1: A from shared memory; acc_bufs accumulators used round-robin.  out[0] = cycles from first issue
// to completion (commit -> mbarrier), out[1] = cycles spent issuing.
__global__ void __launch_bounds__(128) tc_mma_rate_kernel(int N, int mode, int reps, int acc_bufs, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (N * 8 + 128 * 8) * 4 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (warp == 0) tc_alloc(&tmem_slot, 512);
    if (tid == 0) {
        ssf_mbar_init(&bar, 1);
        ssf_mbar_fence_init();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (mode < 2) {
        if (tid == 0) {   // divergent single-thread issue (what `if (lane == 0)` role code compiles to)
            const uint32_t idesc = tc_idesc_tf32(128, N);
            const uint32_t lboW = (uint32_t)(N / 8) * 128, lboA = 16 * 128;
            const uint64_t bdesc = tc_smem_desc(ssf_smem_u32(smem), lboW, 128);
            const uint64_t adesc = tc_smem_desc(ssf_smem_u32(smem + N * 8 * 4), lboA, 128);
            const long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                const uint32_t d = tmem + (uint32_t)((i % acc_bufs) * N);
                if (mode == 0) tc_mma_ts(d, tmem + 496, bdesc, idesc, 1);
                else tc_mma_ss(d, adesc, bdesc, idesc, 1);
            }
            const long long t1 = clock64();
            tc_commit(&bar);
            ssf_mbar_wait(&bar, 0);
            const long long t2 = clock64();
            out[0] = t2 - t0;
            out[1] = t1 - t0;
        }
    } else if (warp == 0) {
        // warp-uniform issue: every lane runs the loop, one elected lane executes the MMA.  Address arithmetic stays in
        // uniform registers (no per-instruction R2UR), the loop is unrolled by 8 with the K-step baked into the descriptor.
        uint32_t is_leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(is_leader));
        const uint32_t idesc = tc_idesc_tf32(128, N);
        const uint32_t lboW = (uint32_t)(N / 8) * 128, lboA = 16 * 128;
        const uint64_t bdesc = tc_smem_desc(ssf_smem_u32(smem), lboW, 128);
        const uint64_t adesc = tc_smem_desc(ssf_smem_u32(smem + N * 8 * 4), lboA, 128);
        const long long t0 = clock64();
        for (int i = 0; i < reps; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t d = tmem + (uint32_t)((u % 2) * (acc_bufs - 1) * N);
                if (is_leader) {
                    if (mode == 2) tc_mma_ts(d, tmem + 496, bdesc, idesc, 1);
                    else tc_mma_ss(d, adesc, bdesc, idesc, 1);
                }
            }
        }
        const long long t1 = clock64();
        if (is_leader) tc_commit(&bar);
        ssf_mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (is_leader) {
            out[0] = t2 - t0;
            out[1] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, 512);
}

extern "C" int ssf_tc_mma_rate(int N, int mode, int reps, int acc_bufs, long long* out, void* stream) {
    if (N % 16 || N < 16 || N > 256 || acc_bufs < 1 || acc_bufs * N > 480) return ssf_arg_error("tc_mma_rate: bad shape");
    tc_mma_rate_kernel<<<1, 128, (N * 8 + 128 * 8) * 4 + 1024, (cudaStream_t)stream>>>(N, mode, reps, acc_bufs, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

// host evaluation of the Hilbert key (regression test of the curve: tests/test_host.py)
extern "C" int ssf_dev_hilbert30(const unsigned* xyz, int n, unsigned* key) {
    for (int i = 0; i < n; ++i) key[i] = ssf_hilbert30(xyz[3 * i] & 1023u, xyz[3 * i + 1] & 1023u, xyz[3 * i + 2] & 1023u);
    return 0;
}

// host evaluation of the invariant division used by the persistent kernels: q[i] = n[i] / d through ssf_fastdiv
extern "C" int ssf_dev_fastdiv(const unsigned* n, int count, unsigned d, unsigned* q) {
    if (d == 0) return 1;
    const SsfFastDiv f = ssf_fastdiv_make(d);
    for (int i = 0; i < count; ++i) q[i] = ssf_fastdiv(n[i], f.mul, f.sh);
    return 0;
}
