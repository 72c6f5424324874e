/*
 * ssf_b200.h -- C ABI of the B200-native SSF-SLAM scene-flow front end (libssf_b200.so).
 *
 * Every entry point takes plain DEVICE pointers, sizes and a CUDA stream (`void* stream` is a cudaStream_t);
 * nothing allocates, nothing synchronises, outputs are caller-owned.  Return value: 0 = ok, 1 = bad
 * argument, 2 = CUDA error; `ssf_last_error()` holds the message.  There is no CPU fallback.
 *
 * Each declaration cites the reference interface it replaces.  `ASF/` = scripts/ActiveSceneFlow/ in
 * YQChen8/SSF-SLAM.  The reference reaches these through the Python module `lib.pointnet2_utils`
 * (absent from its tree: .gitignore:74, README.md:22-27) and `torch_scatter`; INTEGRATION.md shows the
 * ctypes binding that re-creates that module on top of this header.
 *
 * Layouts: "cm" = the reference's channel-major [B,C,N]; "pm" = point-major [B,N,C] (internal, fused path).
 * dtypes: float = fp32, int = int32.
 */
#ifndef SSF_B200_H
#define SSF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plumbing ---- */
int ssf_abi_version(void);
const char* ssf_last_error(void);
unsigned long long ssf_launch_count(void);   /* kernels launched by this library so far */
int ssf_require_device(void);                 /* non-zero unless the current device is sm_100 */

/* ---- B-op: lib.pointnet2_utils drop-ins (reference layouts) ---- */

/* furthest_point_sample(xyz[B,N,3], npoint) -> idx[B,npoint]; call site ASF/utils/utils.py:226.
 * First index 0, running min initialised 1e10, argmax ties -> lowest index. N <= 131072.
 * 2048 < N <= 8192 runs an exactly pruned variant (Hilbert-sorted rows with box bounds; same indices bit for bit; the
 * environment variable SSF_FPS_PRUNE=0 selects the plain scan for comparisons), N > 8192 a thread-block cluster per cloud. */
int ssf_furthest_point_sample(const float* xyz, int B, int N, int npoint, int* idx, void* stream);

/* gather_operation(features cm[B,C,N], idx[B,M]) -> cm[B,C,M]; ASF/utils/utils.py:228 */
int ssf_gather_operation(const float* feat, const int* idx, int B, int C, int N, int M, float* out, void* stream);

/* knn(k, unknown[B,Nq,3], known[B,Nr,3]) -> (dist[B,Nq,k] = sqrt(L2^2) ascending, idx[B,Nq,k]);
 * ASF/utils/utils.py:229,291; ASF/utils/soflow.py:387-391,406,1243,1461.  Order: (distance, index)
 * lexicographic; distances are ((dx*dx)+(dy*dy))+(dz*dz) without FMA.  dist may be NULL.  1 <= k <= 32, k <= Nr. */
int ssf_knn(int k, const float* query, const float* ref, int B, int Nq, int Nr, float* dist, int* idx, void* stream);

/* same, with the query formed on the fly as query + query_add (one rounded fp32 add per coordinate), i.e.
 * `pointutils.knn(nsample, xyz1 + sf, xyz2)` at ASF/utils/soflow.py:389 without materialising the sum */
int ssf_knn_offset(int k, const float* query, const float* query_add, const float* ref, int B, int Nq, int Nr,
                   float* dist, int* idx, void* stream);

/* Same result as ssf_knn_offset, bit for bit, through a spatial index: `build` sorts each reference cloud along a Hilbert curve (isotropic 10-bit cells) into
 * blocks of 32 points with bounding boxes (workspace of ssf_knn_blocks_workspace_floats(B,Nr) floats, Nr <= 131072; above
 * 16384 points the sort is a global-memory radix sort and super-blocks of 32 blocks add a second level),
 * `search` walks the blocks nearest-first and stops when no remaining box can hold a closer point. */
long long ssf_knn_blocks_workspace_floats(int B, int Nr);
int ssf_knn_blocks_build(const float* ref, int B, int Nr, float* ws, void* stream);
int ssf_knn_blocks_search(int k, const float* query, const float* query_add, const float* ws, int B, int Nq, int Nr,
                          float* dist, int* idx, void* stream);

/* ball_query through the index `ws` built by ssf_knn_blocks_build for the cloud xyz[B,N,3]: identical output to ssf_ball_query
 * (ASF/SetCover.py:39-63 semantics), visiting only the blocks whose box can reach the ball.  nsample <= 32. */
int ssf_ball_query_blocks(float radius, int nsample, const float* new_xyz, const float* ws, int B, int N, int S, int* idx, int* cnt,
                          void* stream);

/* Same result by a brute-force scan with one warp per two queries (lane = reference point): for the few-queries / large-cloud
 * corner (B * Nq <= 32768 and Nr > 16384, e.g. 2048 centres in 65536 points) where one thread per query leaves the GPU empty */
int ssf_knn_warp_scan(int k, const float* query, const float* query_add, const float* ref, int B, int Nq, int Nr,
                      float* dist, int* idx, void* stream);

/* three_nn(unknown[B,n,3], known[B,m,3]) -> (dist[B,n,3], idx[B,n,3]); ASF/utils/soflow.py:1241,1459 */
int ssf_three_nn(const float* query, const float* ref, int B, int Nq, int Nr, float* dist, int* idx, void* stream);

/* ball_query(radius, nsample, xyz[B,N,3], new_xyz[B,S,3]) -> idx[B,S,nsample] (+ cnt[B,S] hits in range, may be
 * NULL); semantics of ASF/SetCover.py:39-63 (d <= r*r, ascending index, pad with first hit; no hit -> zeros) */
int ssf_ball_query(float radius, int nsample, const float* xyz, const float* new_xyz, int B, int N, int S, int* idx,
                   int* cnt, void* stream);

/* grouping_operation(features cm[B,C,N], idx[B,M,S]) -> cm[B,C,M,S]; ASF/utils/utils.py:231,233,296,302;
 * ASF/utils/soflow.py:30,1244,1249,1462,1470 */
int ssf_grouping_operation(const float* feat, const int* idx, int B, int C, int N, int M, int S, float* out,
                           void* stream);

/* three_interpolate(features cm[B,C,M], idx[B,N,3], weight[B,N,3]) -> cm[B,C,N] (upstream pointnet2 API; the
 * reference spells it as grouping * weight -> sum at ASF/utils/soflow.py:1462-1471) */
int ssf_three_interpolate(const float* feat, const int* idx, const float* weight, int B, int C, int M, int N,
                          float* out, void* stream);

/* ---- B-scatter: torch_scatter drop-ins (ASF/utils/soflow.py:474,481), deterministic ---- */

/* CSR of the key lists: workspace = ssf_csr_workspace_ints(B,L,n_seg) int32s; keys outside [0,n_seg) are skipped */
long long ssf_csr_workspace_ints(int B, int L, int n_seg);
int ssf_build_csr_i64(const long long* key, int B, int L, int n_seg, int* ws, void* stream);
int ssf_build_csr_i32(const int* key, int B, int L, int n_seg, int* ws, void* stream);
/* scatter_softmax(src[B,L,C], index, dim=1) -> [B,L,C] */
int ssf_segment_softmax(const float* src, const int* csr_ws, int B, int L, int C, int n_seg, float* out, void* stream);
/* scatter_sum(src[B,L,C], index, dim=1) -> [B,n_seg,C]; rows added in ascending l */
int ssf_segment_sum(const float* src, const int* csr_ws, int B, int L, int C, int n_seg, float* out, void* stream);
/* fused: out[b,j,:] = sum_l softmax_seg(logit)[l] * val[l,:]  (the backward cost, soflow.py:471-481) */
int ssf_segment_softmax_sum(const float* logit, const float* val, const int* csr_ws, int B, int L, int C, int n_seg,
                            float* out, void* stream);

/* ---- B-layer: fused building blocks of TFlow (point-major, K-major BN-folded weights) ---- */

/* y = clamp2(clamp1(act(x1.Wt[off1:] + x2.Wt[off2:] + bias)) + add); act 0 none / 1 ReLU / 2 LeakyReLU(0.1).
 * All 1x1 Conv1d layers: ASF/TFlowV3_Occlussion.py:22-38,68,98-100; ASF/utils/utils.py:312-313;
 * ASF/utils/soflow.py:511-525 (flow_mlp, fc, clamp +-50, + sf, clamp) and the per-point halves of split convs */
int ssf_linear(const float* x1, int c1, int ld1, const float* x2, int c2, int ld2, const float* Wt, int ldw,
               int w_off1, int w_off2, const float* bias, int rows, int cout, int act, float clamp1, const float* add,
               int ld_add, float clamp2, float* y, int ldy, void* stream);

/* pm row gather: out[b,m,:] = src[b,idx[b,m],:]  (new_xyz = xyz[fps_idx], ASF/utils/utils.py:228) */
int ssf_gather_rows(const float* src, const int* idx, int B, int N, int M, int C, float* out, void* stream);

/* [B,R,C] -> [B,C,R] */
int ssf_transpose(const float* in, int B, int R, int C, float* out, void* stream);

/* UpsampleFlow.forward (mode 0, clamp 100; ASF/utils/soflow.py:1443-1475) and the interpolation + warp of
 * PointWarping.forward (mode 1, clamp 10; ASF/utils/soflow.py:1244-1257), given the neighbour indices.  idx rows are
 * ld_idx apart and their first k entries are used: the k nearest neighbours are a prefix of any longer (distance, index)
 * ordered list, so one kNN call serves every interpolation between the same two clouds. */
int ssf_interpolate(const float* query, const float* src_pos, const float* src_val, const int* idx, int ld_idx, int B,
                    int N, int M, int C, int k, int mode, float clampv, float* out, void* stream);

/* gather -> MLP -> max over S: PointNetSetAbstraction.forward (ASF/utils/utils.py:231-247), the mlp1 half of
 * PointNetSetUpConv.forward (:296-307) and mlp_convs4 + max of PointConvTransFlowV2 (ASF/utils/soflow.py:489-509).
 * G/H are the first layer's per-point partial products; S in {8,16}; widths multiples of 32, hidden <= 256. */
int ssf_group_mlp_max(const float* G, const float* H, const float* bias1, const float* Wd, const float* pos_src,
                      const float* pos_q, const int* idx, const float* W2t, const float* b2, int C2, const float* W3t,
                      const float* b3, int C3, int B, int Nsrc, int Nq, int S, int C1, int act, float* out, void* stream);

/* PointConvTransFlowV2.forward core (ASF/utils/soflow.py:397-469,486): both branches, SxS attention, mlp_convs3,
 * weightnet1, forward cost; emits warped-branch logits gw[B,N1*16] and rows Cw[B,N1*16,m]; m in {64,128,256} */
int ssf_cost_volume(const float* Gab, const float* Hab, const float* W2a, const float* b2a, const float* W2w,
                    const float* b2w, const float* W3a, const float* H3, const float* W3d, const float* W3b,
                    const float* b3b, const float* Wn1, const float* bn1, const float* Wn2, const float* bn2,
                    const float* wn3, float bn3, const float* xyz1, const float* xyz2, const int* idx, const int* idxw,
                    int B, int N1, int N2, int m, float* cost_fwd, float* cost_fwd_cm, float* gw, float* Cw, void* stream);

/* tcgen05 / TMEM realisation of the same core for m == 64 (levels 0 and 1): 3xTF32 split GEMMs with fp32 TMEM
 * accumulators, persistent CTA per SM.  wblob / params come from ssf_slam_b200.tc.cost_volume_tc_pack
 * (ssf_cost_volume_tc_blob_bytes() bytes, ssf_cost_volume_tc_param_floats() floats); n_sm <= 0 -> 148. */
long long ssf_cost_volume_tc_blob_bytes(void);
int ssf_cost_volume_tc_param_floats(void);
int ssf_cost_volume_tc(const float* Gab, const float* Hab, const float* H3, const void* wblob, const float* params,
                       const float* xyz1, const float* xyz2, const int* idx, const int* idxw, int B, int N1, int N2,
                       int m, float* cost_fwd, float* cost_fwd_cm, float* gw, float* Cw, int n_sm, void* stream);

/* CUDA-core pieces of the un-fused wide (m >= 128) cost volume; the dense layers between them are ssf_dense_tc calls.
 * attention_mix: A, Aw [P,16,m] -> A + Q.Aw, Aw + Q^T.A, Q = softmax(-2) * softmax(-1) of <A_i,Aw_j> (soflow.py:420-422,453-458)
 * softmax_pool : cost_fwd = sum_s softmax_s(g) * C  -> point-major [B,N1,m] and channel-major [B,m,N1] (soflow.py:469,486) */
int ssf_attention_mix(const float* A, const float* Aw, long long n_points, int m, float* Amix, float* Awmix, void* stream);
int ssf_softmax_pool(const float* g, const float* C, int B, int N1, int m, float* out_pm, float* out_cm, void* stream);

/* tensor-core dense layer (argument block and documentation in ssf_dense.h) */
#include "ssf_dense.h"

/* ---- B-frontend: dynamic mask + static-point ego-motion (ASF/main_sju_occ_ros.py:257-284,455-473;
 * scripts/PointCloudOdometry.py:15-33,91-101) ----
 * mode 0: weighted Kabsch on the points whose in_mask is 0 (GT / dataset mask variants);
 * mode 2: as mode 0 with `flow` holding the source cloud itself: slove_RT_by_SVD(src = flow, dst = points);
 * mode 1: residual-vs-rigid-flow masker with optional semantic seed (movable class bitmask) and per-instance
 *         voting, then Kabsch on the static set.  Bit-level spec: oracle/frontend.py::masker_spec.
 * points, flow [B,N,3]; mask_out u8 [B,N] (1 = dynamic); odom_out f64 [B,7] = [tx,ty,tz,qx,qy,qz,qw]
 * (the frame_odom1 payload); pose_out f64 [B,12] = [R row-major | t] or NULL. */
int ssf_frontend(const float* points, const float* flow, int B, int N, int mode, const unsigned char* in_mask,
                 const int* sem, unsigned long long movable_bits, const int* inst, int n_inst, float tau,
                 unsigned char* mask_out, double* odom_out, double* pose_out, void* stream);

/* slove_RT_by_SVD(src[M,3], dst[M,3]) -> dst ~= R src + t (scripts/PointCloudOdometry.py:15-33) for float64 clouds, computed in
 * float64 like the reference computes in its input dtype.  src, dst f64 [B,M,3] -> odom f64 [B,7], pose f64 [B,12] = [R | t].
 * Fewer than 3 points -> identity. */
int ssf_solve_rt_f64(const double* src, const double* dst, int B, int M, double* odom_out, double* pose_out, void* stream);

/* ---- the reference's noSeg masker on the device: 2-component full-covariance Gaussian mixture by EM on [flow | xyz],
 * majority component = background.  Replaces `GaussianMixture(n_components=2).fit_predict(...)` + `Counter.most_common(1)` at
 * scripts/PointCloudOdometry_noSeg.py:97-103 and ASF/main_sju_occ_ros.py:257-263 (scikit-learn defaults: tol 1e-3, reg_covar
 * 1e-6, max_iter 100, one initialisation); the k-means seeding the reference draws from the global RNG is deterministic here
 * (specification pinned against scikit-learn: oracle/gmm.py).  points, flow [B,N,3] f32 -> mask u8 [B,N] (0 = background),
 * info f64 [B,4] = (EM iterations, lower bound, converged, background count) or NULL.  Feed the mask to ssf_frontend mode 0
 * for the pose. */
int ssf_gmm_mask(const float* points, const float* flow, int B, int N, int max_iter, double tol, unsigned char* mask_out,
                 double* info_out, void* stream);

/* ---- before the path: the CARLA frame subsampler's data movement on the device, ASF/utils/datasets/carla.py:179-305.
 * Clouds are index lists into the raw frame.  select: entries of `pre` (NULL = 0..n-1) that pass the ground cut (keep unless
 * pts[src, ld-1] < ground_z; carla.py:237) and the mask test (mask_mode 0 none, 1 mask != 0, 2 mask == 0, 3 mask == 1; mask u8 indexed
 * by raw index), in order -> sel (raw indices, capacity n) and *count.  index_compose: out[j] = sel[ind[j]] (sel NULL = identity),
 * *err set to 1 if an index is outside [0, n_sel).  gather_u8: out[j] = src[idx[j]]. */
int ssf_dataset_select(const float* pts, int ld, const unsigned char* mask, const int* pre, int n, int ground_cut, float ground_z,
                       int mask_mode, int* sel, int* count, void* stream);
int ssf_index_compose(const int* sel, int n_sel, const int* ind, int m, int* out, int* err, void* stream);
int ssf_gather_u8(const unsigned char* src, int n, const int* idx, int m, unsigned char* out, void* stream);

/* ---- next after the path: plane-feature extraction of the back end's first node, src/frameFeature.cpp:45-127
 * (scan-line id from elevation, stable regrouping per line, 11-tap curvature, greedy selection curvature < plane_min with a
 * skip of plane_span).  points [B,N,3] -> out [B,N,4] = (x, y, z, intensity = indexInRow + line/100), out_count [B].
 * n_rows in {16, 64}; the node's parameters are (16: 0.05, 3, rows 0..15) and (64: 0.005, 25, rows 5..58). */
long long ssf_plane_features_workspace_bytes(int B, int N);
int ssf_plane_features(const float* points, int B, int N, int n_rows, int row_start, int row_end, float plane_min,
                       int plane_span, void* ws, float* out, int* out_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSF_B200_H */
