/*
 * ssf_b200_dev.h -- developer / regression entry points, built into libssf_b200_dev.so (NOT part of the product ABI in
 * include/ssf_b200.h): tensor-core bring-up GEMM and the tensor-pipe pacing probe.  Used by tests/test_gpu_tc.py and scripts/.
 */
#ifndef SSF_B200_DEV_H
#define SSF_B200_DEV_H
#ifdef __cplusplus
extern "C" {
#endif

const char* ssf_last_error(void);

/* ---- tensor-core bring-up / regression: Y[128,N] = X[128,K].W[N,K]^T on tcgen05 kind::tf32 (3xTF32 when passes == 3);
 * Whi_img / Wlo_img are the split weights in the no-swizzle K-major UMMA image (ssf_slam_b200.tc.weight_image);
 * mode 0: A operand from TMEM, mode 1: A operand from shared memory */
int ssf_tc_gemm_test(const float* X, const float* Whi_img, const float* Wlo_img, int K, int N, int mode, int passes,
                     float* Y, void* stream);

/* tensor-pipe pacing probe (developer tool): cycles for `reps` back-to-back M128 x N x K8 kind::tf32 MMAs, A operand from
 * TMEM (mode 0) or shared memory (mode 1), acc_bufs accumulators round-robin; out[0] = total cycles, out[1] = issue cycles */
int ssf_tc_mma_rate(int N, int mode, int reps, int acc_bufs, long long* out, void* stream);

/* host evaluation of the curve key the spatial indices sort by (ssf_common.cuh ssf_hilbert30; no GPU needed): n cells with
 * 10-bit coordinates -> 30-bit keys */
int ssf_dev_hilbert30(const unsigned* xyz, int n, unsigned* key);

/* host evaluation of the invariant division of the persistent kernels (ssf_common.cuh ssf_fastdiv): q[i] = n[i] / d, n[i] < 2^31 */
int ssf_dev_fastdiv(const unsigned* n, int count, unsigned d, unsigned* q);

#ifdef __cplusplus
}
#endif
#endif /* SSF_B200_DEV_H */
