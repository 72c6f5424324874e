// Plane-feature extraction of the LOAM back end's first node, on the GPU (the step that consumes the
// `velodyne_points` cloud this front end publishes): src/frameFeature.cpp:35-127 of the reference.
//   1. scan-line id of every point from its elevation angle            (:53-73; 16-line and 64-line rules)
//   2. points regrouped per scan line, input order kept                (:74-82; indexInRow, intensity = index + id/100)
//   3. 11-tap curvature along each line, rounded op by op as in C++    (:85-108)
//   4. greedy selection per line: curvature < planeMin, then skip planeSpan positions  (:110-126)
// One CTA per cloud.  Steps 1-3 are data parallel (the stable regrouping is a counting sort whose ranks come from
// __match_any_sync ballots, so the order is deterministic); step 4 is a short sequential scan per line (<= 64 lines).
// The voxel-grid filter at :128-131 writes a temporary the node never publishes; it is not reproduced.
#include "ssf_common.cuh"

namespace {

constexpr int PF_T = 1024;
constexpr int PF_MAXROWS = 64;

__device__ __forceinline__ int scan_id(float x, float y, float z, int n_rows) {
    // float angle = atan(point.z / sqrt(x*x + y*y)) * 180 / M_PI;   (float overloads, the final division in double)
    const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
    const float a = (float)atan((double)__fdiv_rn(z, r));   // double atan rounded to float == correctly rounded atanf
    const float angle = (float)((double)__fmul_rn(a, 180.0f) / 3.14159265358979323846);
    if (isnan(angle)) return -1;
    int id = -1;
    if (n_rows == 16) {
        id = (int)((double)__fdiv_rn(__fadd_rn(angle, 15.0f), 2.0f) + 0.5);
    } else if (n_rows == 64) {
        if ((double)angle >= -8.83) id = (int)((double)__fsub_rn(2.0f, angle) * 3.0 + 0.5);   // (2 - angle) is a float op in C++
        else id = n_rows / 2 + (int)((-8.83 - (double)angle) * 2.0 + 0.5);
    }
    return (id > -1 && id < n_rows) ? id : -1;
}

// ws ints per cloud: sid [N] | sorted [N] | selected [N];  ws floats: value [N]
__global__ void __launch_bounds__(PF_T) plane_features_kernel(const float* __restrict__ pts, int N, int n_rows, int row_start,
                                                              int row_end, float plane_min, int plane_span, int* __restrict__ wsi,
                                                              float* __restrict__ wsf, float* __restrict__ out,
                                                              int* __restrict__ out_count) {
    __shared__ int s_cnt[PF_MAXROWS], s_off[PF_MAXROWS + 1], s_base[PF_MAXROWS], s_sel[PF_MAXROWS], s_ooff[PF_MAXROWS + 1];
    __shared__ int s_w[32][PF_MAXROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const float* P = pts + (size_t)b * N * 3;
    int* sid = wsi + (size_t)b * 3 * N;
    int* sorted = sid + N;
    int* selected = sorted + N;
    float* value = wsf + (size_t)b * N;
    float* O = out + (size_t)b * N * 4;

    if (tid < PF_MAXROWS) { s_cnt[tid] = 0; s_base[tid] = 0; }
    __syncthreads();
    // 1. scan-line ids + histogram
    for (int i = tid; i < N; i += PF_T) {
        const int id = scan_id(P[3 * i], P[3 * i + 1], P[3 * i + 2], n_rows);
        sid[i] = id;
        if (id >= 0) atomicAdd(&s_cnt[id], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int r = 0; r < n_rows; ++r) { s_off[r] = acc; acc += s_cnt[r]; }
        s_off[n_rows] = acc;
    }
    __syncthreads();
    // 2. stable regrouping: chunks of PF_T points in input order
    for (int c0 = 0; c0 < N; c0 += PF_T) {
        for (int e = tid; e < 32 * PF_MAXROWS; e += PF_T) (&s_w[0][0])[e] = 0;
        __syncthreads();
        const int i = c0 + tid;
        const int id = i < N ? sid[i] : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, id);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (id >= 0 && rank == 0) s_w[warp][id] = __popc(peers);
        __syncthreads();
        if (id >= 0) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += s_w[w][id];
            sorted[s_off[id] + s_base[id] + before + rank] = i;
        }
        __syncthreads();
        if (tid < n_rows) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += s_w[w][tid];
            s_base[tid] += tot;
        }
        __syncthreads();
    }
    const int M = s_off[n_rows];   // points with a valid scan line
    // 3. curvature (value stays 0 for the first / last five points of a line and for excluded lines)
    for (int p = tid; p < M; p += PF_T) {
        const int id = sid[sorted[p]];
        const int j = p - s_off[id], size = s_cnt[id];
        float v = 0.f;
        if (id >= row_start && id < n_rows - row_end && j >= 5 && j < size - 5) {
            float d[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                auto at = [&](int k) { return P[3 * sorted[p + k] + c]; };
                float s = __fadd_rn(at(-5), at(-4));
                s = __fadd_rn(s, at(-3));
                s = __fadd_rn(s, at(-2));
                s = __fadd_rn(s, at(-1));
                s = __fsub_rn(s, __fmul_rn(10.0f, at(0)));
                s = __fadd_rn(s, at(1));
                s = __fadd_rn(s, at(2));
                s = __fadd_rn(s, at(3));
                s = __fadd_rn(s, at(4));
                s = __fadd_rn(s, at(5));
                d[c] = s;
            }
            v = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
        }
        value[p] = v;
        selected[p] = 0;
    }
    __syncthreads();
    // 4. greedy selection, one thread per scan line
    if (tid < n_rows) {
        int cnt = 0;
        if (tid >= row_start && tid < n_rows - row_end) {
            const int o = s_off[tid], size = s_cnt[tid];
            int jstart = 0;
            for (int j = 0; j < size; ++j)
                if (j >= jstart && value[o + j] < plane_min) {
                    selected[o + j] = 1;
                    ++cnt;
                    jstart = j + plane_span;
                }
        }
        s_sel[tid] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int r = 0; r < n_rows; ++r) { s_ooff[r] = acc; acc += s_sel[r]; }
        s_ooff[n_rows] = acc;
        out_count[b] = acc;
    }
    __syncthreads();
    if (tid < n_rows && s_sel[tid] > 0) {
        const int o = s_off[tid], size = s_cnt[tid];
        int k = s_ooff[tid];
        for (int j = 0; j < size; ++j)
            if (selected[o + j]) {
                const int i = sorted[o + j];
                float* q = O + (size_t)k * 4;
                q[0] = P[3 * i]; q[1] = P[3 * i + 1]; q[2] = P[3 * i + 2];
                q[3] = (float)((double)j + (double)tid / 100.0);   // intensity = indexInRow + scanID / 100.0
                ++k;
            }
    }
}

}  // namespace

extern "C" long long ssf_plane_features_workspace_bytes(int B, int N) { return (long long)B * N * 16; }

// points [B,N,3] -> out [B,N,4] (x, y, z, intensity), out_count [B]; n_rows in {16, 64}.  ws: workspace of
// ssf_plane_features_workspace_bytes(B, N) bytes.
extern "C" int ssf_plane_features(const float* points, int B, int N, int n_rows, int row_start, int row_end, float plane_min,
                                  int plane_span, void* ws, float* out, int* out_count, void* stream) {
    if (B <= 0 || N <= 0) return ssf_arg_error("plane_features: empty input");
    if (n_rows != 16 && n_rows != 64) return ssf_arg_error("plane_features: n_rows must be 16 or 64");
    if (row_start < 0 || row_end < 0 || plane_span < 0) return ssf_arg_error("plane_features: negative parameter");
    int* wsi = static_cast<int*>(ws);
    float* wsf = reinterpret_cast<float*>(wsi + (size_t)B * 3 * N);
    plane_features_kernel<<<B, PF_T, 0, (cudaStream_t)stream>>>(points, N, n_rows, row_start, row_end, plane_min, plane_span, wsi, wsf,
                                                                out, out_count);
    ssf_count_launch();
    SSF_LAUNCH_CHECK();
    return SSF_OK;
}
