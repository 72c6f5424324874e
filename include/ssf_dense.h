/* Argument block of ssf_dense_tc (tensor-core dense layer, csrc/dense_tc.cu).  Plain C, device pointers only.
 * Replaces the 1x1 Conv2d/Conv1d (+ folded BatchNorm + activation [+ max over nsample]) stacks of
 * ASF/utils/utils.py:236-247,304-313 and ASF/utils/soflow.py:397-451,460-461,501-513. */
#ifndef SSF_DENSE_H
#define SSF_DENSE_H

enum { SSF_EPI_STORE = 0, SSF_EPI_MAX = 1, SSF_EPI_DOT = 2 };

typedef struct ssf_dense_args {
    /* ---- A operand [rows, K] */
    int a_mode;                 /* 0: rows of x1 | x2 (concatenated along K); 1: grouped first layer (see below) */
    int K;                      /* multiple of 32 */
    long long rows;             /* grouped: rows = B * Nq * S, row = (b * Nq + n) * S + s */
    const float* x1; int c1; int ld1;
    const float* x2; int c2; int ld2;
    /* a_mode 1: A[row, c] = act1(G[b, idx[row], offG + c] + H[b*Nq+n, offH + c] + b1[c] + Wd1[:, c] . dir(row)) */
    const float* G; int ldG; int offG;
    const float* H; int ldH; int offH;        /* may be NULL */
    const float* b1;                          /* [K] or NULL */
    const float* Wd1;                         /* [3, K] or NULL */
    int act1;
    /* grouped-row geometry (needed by a_mode 1, Hq, Wd2 and the MAX epilogue) */
    const int* idx;                           /* [B, Nq, S] neighbour indices into the source cloud */
    int S; int Nq; int Nsrc;
    const float* pos_src;                     /* [B, Nsrc, 3]; dir(row) = pos_src[idx[row]] - pos_q[n] */
    const float* pos_q;                       /* [B, Nq, 3] */
    /* ---- weights: image from ssf_slam_b200.tc.dense_image (per 256-column tile, per 32-wide K chunk: hi | lo) */
    const void* wimg; int N;                  /* N multiple of 32; multiple of 256 when > 256 */
    /* ---- epilogue: v = act(D + bias + Hq[point] + Wd2 . dir(row)) */
    const float* bias;                        /* [N] or NULL */
    const float* Hq; int ldHq;                /* per-point rows [B*Nq, >= N] or NULL */
    const float* Wd2;                         /* [3, N] or NULL */
    int act;                                  /* 0 none, 1 ReLU, 2 LeakyReLU(0.1) */
    int epi_mode;                             /* SSF_EPI_* */
    const float* wvec; float b0;              /* DOT: y[row] = wvec . v + b0 */
    float* y; int ldy;                        /* STORE: [rows, ldy]; MAX: [rows / S, ldy]; DOT: [rows] */
} ssf_dense_args;

#ifdef __cplusplus
extern "C" {
#endif
int ssf_dense_tc(const ssf_dense_args* args, void* stream);
/* Layers with N <= 64 and a small weight image can run a light kernel variant (288 threads, 256 TMEM columns, two CTAs per
 * SM).  mode 0: never; 1 (default): where it measured faster (pooled plain-row layers); 2: every eligible layer.  Results are
 * bit-identical across modes.  Returns the previous mode. */
int ssf_dense_set_variant(int mode);
/* Plain-row inputs and stored outputs travel by tensor-map TMA (cp.async.bulk.tensor, 32 x 32 fp32 boxes, 128-byte swizzle) when the
 * arrays are 16-byte aligned with leading dimensions that are multiples of 4 floats (level 1, default); 0 switches those paths off
 * (cp.async / st.global instead); 2 also fetches the gathered rows of grouped layers by tile::gather4 (correct, measured slower than
 * the cp.async gather, hence not the default).  Results are bit-identical at every level.  Returns the previous level. */
int ssf_dense_set_tma(int on);
int ssf_dense_args_bytes(void);   /* sizeof(ssf_dense_args) as compiled, for binding self-checks */
#ifdef __cplusplus
}
#endif
#endif /* SSF_DENSE_H */
