/*
 * oracle/c/ssf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the point operators on SSF-SLAM's scene-flow path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * The reference's own pointnet2 CUDA extension is absent from the reference tree
 * (.gitignore:74, README.md:22-27), so these follow the pure-PyTorch look-alikes the
 * reference ships and the arithmetic / tie-breaking spec of SURVEY.md Appendix C:
 *
 *   squared distance  d = ((dx*dx) + (dy*dy)) + (dz*dz), IEEE fp32 RN, no FMA contraction
 *   (build with -ffp-contract=off; follows scripts/ActiveSceneFlow/utils/utils.py:85,106)
 *   FPS   : utils.py:68-89, first index 0 (upstream extension), running min initialised 1e10,
 *           argmax ties -> lowest index
 *   kNN   : utils.py:92-108, k smallest by (d, index) lexicographic, ascending, dist = sqrtf(d)
 *   ball  : SetCover.py:39-63, d <= r*r (r*r in fp32), ascending index, first nsample,
 *           pad with first hit; no hit -> all zeros (upstream CUDA kernel behaviour), cnt = hits
 *   group : grouping_operation / gather_operation as called at utils.py:228-233
 *
 * PARITY UNPINNED: the reference holds no test or golden vector for these operators.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist3(const float *a, const float *b)
{
    float dx = a[0] - b[0];
    float dy = a[1] - b[1];
    float dz = a[2] - b[2];
    float xx = dx * dx;
    float yy = dy * dy;
    float zz = dz * dz;
    float s = xx + yy;
    return s + zz;
}

/* xyz [B,N,3] -> idx [B,npoint] */
void ssf_oracle_fps(const float *xyz, int B, int N, int npoint, int32_t *idx)
{
#pragma omp parallel for
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        int32_t *out = idx + (size_t)b * npoint;
        float *mind = (float *)malloc(sizeof(float) * (size_t)N);
        for (int i = 0; i < N; ++i) mind[i] = 1e10f;
        int last = 0;
        for (int j = 0; j < npoint; ++j) {
            out[j] = last;
            if (j == npoint - 1) break;
            float best = -1.0f;
            int besti = 0;
            for (int i = 0; i < N; ++i) {
                float d = sqdist3(p + 3 * i, p + 3 * last);
                float m = mind[i];
                if (d < m) m = d;
                mind[i] = m;
                if (m > best) { best = m; besti = i; }
            }
            last = besti;
        }
        free(mind);
    }
}

/* query [B,Nq,3], ref [B,Nr,3] -> dist [B,Nq,k] (sqrt), idx [B,Nq,k]; needs k <= Nr */
void ssf_oracle_knn(int k, const float *query, const float *ref, int B, int Nq, int Nr,
                    float *dist, int32_t *idx)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int q = 0; q < Nq; ++q) {
            const float *qp = query + ((size_t)b * Nq + q) * 3;
            const float *rp = ref + (size_t)b * Nr * 3;
            float bd[64];
            int32_t bi[64];
            int cnt = 0;
            for (int r = 0; r < Nr; ++r) {
                float d = sqdist3(qp, rp + 3 * r);
                /* candidates arrive in ascending index: strict < keeps the lowest index on ties */
                if (cnt == k && !(d < bd[k - 1])) continue;
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && d < bd[pos - 1]) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = r;
                if (cnt < k) ++cnt;
            }
            float *od = dist + ((size_t)b * Nq + q) * k;
            int32_t *oi = idx + ((size_t)b * Nq + q) * k;
            for (int j = 0; j < k; ++j) {
                od[j] = sqrtf(bd[j]);
                oi[j] = bi[j];
            }
        }
    }
}

/* xyz [B,N,3], new_xyz [B,S,3] -> idx [B,S,nsample], cnt [B,S] */
void ssf_oracle_ball_query(float radius, int nsample, const float *xyz, const float *new_xyz,
                           int B, int N, int S, int32_t *idx, int32_t *cnt)
{
    float r2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int s = 0; s < S; ++s) {
            const float *c = new_xyz + ((size_t)b * S + s) * 3;
            const float *p = xyz + (size_t)b * N * 3;
            int32_t *o = idx + ((size_t)b * S + s) * nsample;
            int n = 0;
            for (int i = 0; i < N; ++i) {
                float d = sqdist3(c, p + 3 * i);
                if (d <= r2) {
                    if (n < nsample) o[n] = i;
                    ++n;
                }
            }
            int filled = n < nsample ? n : nsample;
            int32_t pad = filled > 0 ? o[0] : 0;
            for (int j = filled; j < nsample; ++j) o[j] = pad;
            cnt[(size_t)b * S + s] = n;
        }
    }
}

/* feat [B,C,N], idx [B,M,S] -> out [B,C,M,S]  (gather_operation is the S == 1 case) */
void ssf_oracle_group(const float *feat, const int32_t *idx, int B, int C, int N, int M, int S,
                      float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int c = 0; c < C; ++c) {
            const float *f = feat + ((size_t)b * C + c) * N;
            const int32_t *ii = idx + (size_t)b * M * S;
            float *o = out + ((size_t)b * C + c) * M * S;
            for (size_t j = 0; j < (size_t)M * S; ++j) o[j] = f[ii[j]];
        }
    }
}

/* ---- plane-feature extraction: restatement of src/frameFeature.cpp:45-127 (test infrastructure).  Pinned: the node itself
 * compiles unmodified against stand-in ROS / PCL headers (oracle/ref_build/ -> oracle/_ref/libframe_feature_ref.so) and
 * oracle/gen_golden_plane_features.py asserts this restatement equals what it publishes, bit for bit.  Float/double types of
 * every sub-expression are those C++ evaluates with the <math.h> float overloads.
 * points [N,3]; out [N,4]; returns the number of plane points.  tmp: N ints + N ints + N floats + N ints. */
int ssf_oracle_plane_features(const float* P, int N, int n_rows, int row_start, int row_end, float plane_min, int plane_span,
                              float* out) {
    int* row_of = (int*)malloc(sizeof(int) * (size_t)N);
    int* cnt = (int*)calloc((size_t)n_rows, sizeof(int));
    int* off = (int*)calloc((size_t)n_rows + 1, sizeof(int));
    for (int i = 0; i < N; ++i) {
        const float x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
        float angle = (float)((double)(atanf(z / sqrtf(x * x + y * y)) * 180) / M_PI);   /* :56 */
        int id = -1;
        if (n_rows == 16) {
            if (angle >= -15 || angle <= 15) id = (int)((angle + 15) / 2 + 0.5);          /* :58-60 */
        }
        if (n_rows == 64) {
            if (angle >= -24.33 || angle <= 2) {                                            /* :63-69 */
                if (angle >= -8.83) id = (int)((2 - angle) * 3.0 + 0.5);
                else id = n_rows / 2 + (int)((-8.83 - angle) * 2.0 + 0.5);
            }
        }
        if (angle != angle) id = -1;
        row_of[i] = (id > -1 && id < n_rows) ? id : -1;
        if (row_of[i] >= 0) cnt[row_of[i]]++;
    }
    for (int r = 0; r < n_rows; ++r) off[r + 1] = off[r] + cnt[r];
    int* sorted = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    int* fill = (int*)calloc((size_t)n_rows, sizeof(int));
    for (int i = 0; i < N; ++i)
        if (row_of[i] >= 0) sorted[off[row_of[i]] + fill[row_of[i]]++] = i;               /* push_back order, :74-82 */
    float* value = (float*)calloc((size_t)(N > 0 ? N : 1), sizeof(float));
    for (int r = row_start; r < n_rows - row_end; ++r) {                                    /* :85-108 */
        const int* s = sorted + off[r];
        for (int j = 5; j < cnt[r] - 5; ++j) {
            float d[3];
            for (int c = 0; c < 3; ++c) {
#define AT(k) P[3 * s[j + (k)] + c]
                d[c] = AT(-5) + AT(-4) + AT(-3) + AT(-2) + AT(-1) - 10 * AT(0) + AT(1) + AT(2) + AT(3) + AT(4) + AT(5);
#undef AT
            }
            value[off[r] + j] = (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        }
    }
    int n_out = 0;
    for (int r = row_start; r < n_rows - row_end; ++r) {                                    /* :110-126 */
        size_t jstart = 0;
        for (size_t j = 0; j < (size_t)cnt[r]; ++j) {
            if (j >= jstart && value[off[r] + j] < plane_min) {
                const int i = sorted[off[r] + j];
                out[4 * n_out] = P[3 * i]; out[4 * n_out + 1] = P[3 * i + 1]; out[4 * n_out + 2] = P[3 * i + 2];
                out[4 * n_out + 3] = (float)((int)j + r / 100.0);                          /* :79 */
                ++n_out;
                jstart = j + plane_span;
            }
        }
    }
    free(row_of); free(cnt); free(off); free(sorted); free(fill); free(value);
    return n_out;
}
