// Device side of the CARLA frame subsampler (the step right before the hot path,
// ASF/utils/datasets/carla.py:179-305).  The reference filters, compacts and gathers with NumPy on the host; here the raw frame
// lives in HBM, every intermediate cloud is an INDEX LIST into the raw arrays, and only the list lengths (which size the
// host's np.random draws) and the drawn indices cross PCIe:
//   ssf_dataset_select   walks the current list `pre` (NULL = all n raw points) and keeps the entries that pass the ground cut
//                        (carla.py:237-247: keep unless z < -3.3) and / or the mask test (:262-265 mask != 0; :184-190 mask == 0 /
//                        mask == 1), in order (== boolean-indexing / np.argwhere order) -> new list of RAW indices + its length.
//                        One CTA, stable block-wide compaction (ballot ranks + scan of warp totals).
//   ssf_index_compose    out[j] = sel[ind[j]]   (sequence[0][mask][ind1] as one index list)
//   ssf_gather_u8        out[j] = src[idx[j]]   (mask[0][ind1])
// Point / flow rows are then gathered by ssf_gather_rows.  Integer / byte work: bit-exact against the host implementation.
#include "ssf_common.cuh"

namespace {

constexpr int DS_T = 1024;

// mask_mode: 0 ignore mask, 1 keep mask != 0, 2 keep mask == 0, 3 keep mask == 1
__global__ void __launch_bounds__(DS_T)
dataset_select_kernel(const float* __restrict__ pts, int ld, const unsigned char* __restrict__ mask, const int* __restrict__ pre,
                      int n, int ground_cut, float ground_z, int mask_mode, int* __restrict__ sel, int* __restrict__ count) {
    __shared__ int s_warp[DS_T / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += DS_T) {
        const int i = start + tid;
        bool keep = false;
        int src = 0;
        if (i < n) {
            src = pre != nullptr ? pre[i] : i;               // raw index of the i-th point of the current cloud
            keep = true;
            if (ground_cut) keep = !(pts[(size_t)src * ld + (ld - 1)] < ground_z);   // np.logical_not(pc[:, -1] < -3.3): NaN is kept
            if (keep && mask_mode != 0) {
                const unsigned char m = mask[src];
                keep = mask_mode == 1 ? (m != 0) : (mask_mode == 2 ? (m == 0) : (m == 1));
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            int v = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += y;
            }
            s_warp[lane] = v;   // inclusive
        }
        __syncthreads();
        const int base = s_base + (warp > 0 ? s_warp[warp - 1] : 0);
        if (keep) sel[base + __popc(bal & ((1u << lane) - 1u))] = src;
        __syncthreads();
        if (tid == 0) s_base += s_warp[DS_T / 32 - 1];
        __syncthreads();
    }
    if (tid == 0) *count = s_base;
}

__global__ void index_compose_kernel(const int* __restrict__ sel, int n_sel, const int* __restrict__ ind, int m, int* __restrict__ out,
                                     int* __restrict__ err) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int i = ind[j];
    if (i < 0 || i >= n_sel) {
        atomicExch(err, 1);
        out[j] = 0;
        return;
    }
    out[j] = sel != nullptr ? sel[i] : i;
}

__global__ void gather_u8_kernel(const unsigned char* __restrict__ src, int n, const int* __restrict__ idx, int m,
                                 unsigned char* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int i = idx[j];
    out[j] = (i >= 0 && i < n) ? src[i] : 0;
}

}  // namespace

extern "C" int ssf_dataset_select(const float* pts, int ld, const unsigned char* mask, const int* pre, int n, int ground_cut,
                                  float ground_z, int mask_mode, int* sel, int* count, void* stream) {
    if (n < 0 || ld < 1) return ssf_arg_error("dataset_select: bad size");
    if (mask_mode < 0 || mask_mode > 3) return ssf_arg_error("dataset_select: mask_mode must be 0..3");
    if (mask_mode != 0 && mask == nullptr) return ssf_arg_error("dataset_select: mask_mode needs a mask");
    if (ground_cut && pts == nullptr) return ssf_arg_error("dataset_select: ground cut needs the points");
    dataset_select_kernel<<<1, DS_T, 0, (cudaStream_t)stream>>>(pts, ld, mask, pre, n, ground_cut, ground_z, mask_mode, sel, count);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

extern "C" int ssf_index_compose(const int* sel, int n_sel, const int* ind, int m, int* out, int* err, void* stream) {
    if (m <= 0) return ssf_arg_error("index_compose: empty input");
    index_compose_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sel, n_sel, ind, m, out, err);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}

extern "C" int ssf_gather_u8(const unsigned char* src, int n, const int* idx, int m, unsigned char* out, void* stream) {
    if (m <= 0) return ssf_arg_error("gather_u8: empty input");
    gather_u8_kernel<<<(m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, n, idx, m, out);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
