// Shared device helpers for the SSF-SLAM scene-flow front end kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SSF_OK 0
#define SSF_ERR_ARG 1
#define SSF_ERR_CUDA 2

#define SSF_LAUNCH_CHECK()                                  \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return ssf_set_error(e__);  \
    } while (0)

// Kernel function attributes (opt-in shared memory above 48 KB) are PER DEVICE: a process that drives several GPUs must set
// them once on each.  `seen` is a per-call-site bitmask over device ordinals; ssf_attr_needed() is true until ssf_attr_done()
// has been called on the current device (setting an attribute twice from racing threads is harmless).
static inline unsigned long long ssf_device_bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
}
static inline bool ssf_attr_needed(const unsigned long long* seen) { return !(__atomic_load_n(seen, __ATOMIC_ACQUIRE) & ssf_device_bit()); }
static inline void ssf_attr_done(unsigned long long* seen) { __atomic_fetch_or(seen, ssf_device_bit(), __ATOMIC_RELEASE); }

int ssf_set_error(cudaError_t e);
int ssf_arg_error(const char* msg);
void ssf_count_launch();

// Squared distance exactly as the written spec (mirrors
// ASF/utils/utils.py:85,106): every multiply and add rounded to fp32, no FMA contraction.
__device__ __forceinline__ float ssf_sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx);
    float dy = __fsub_rn(ay, by);
    float dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Division of n < 2^31 by a run-time invariant d >= 1 as multiply-high + add + shift (Granlund-Montgomery):
// n / d == (umulhi(n, mul) + n) >> sh with sh = ceil(log2 d), mul = floor(2^32 (2^sh - d) / d) + 1.  A 32-bit division is ~25
// instructions on the device and a 64-bit one ~70; the persistent tensor-core kernels did several per thread per tile.
struct SsfFastDiv {
    unsigned mul, sh;
};
static inline SsfFastDiv ssf_fastdiv_make(unsigned d) {
    SsfFastDiv f;
    unsigned l = 0;
    while ((1ull << l) < d) ++l;
    f.mul = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.sh = l;
    return f;
}
__host__ __device__ __forceinline__ unsigned ssf_fastdiv(unsigned n, unsigned mul, unsigned sh) {
#ifdef __CUDA_ARCH__
    return (__umulhi(n, mul) + n) >> sh;
#else
    return (unsigned)((((unsigned long long)n * mul) >> 32) + n) >> sh;
#endif
}

// 30-bit Hilbert-curve key of a point quantised to 10 bits per axis (Skilling's transpose algorithm, then bit interleave).
// Sorting a cloud by it makes runs of consecutive points spatially compact; every index built on such runs (kNN / ball-query
// blocks, the pruned sampler's rows) is exact for ANY order, the curve only decides how many blocks a query has to open:
// on LiDAR sweeps 4.5 blocks of 32 intersect a 16-NN ball with Hilbert order, 6.2 with Morton (Z) order of the same cells.
__host__ __device__ __forceinline__ unsigned ssf_spread10(unsigned v) {   // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__host__ __device__ __forceinline__ unsigned ssf_hilbert30(unsigned x, unsigned y, unsigned z) {
    unsigned X[3] = {x, y, z};
#pragma unroll
    for (unsigned Q = 512u; Q > 1u; Q >>= 1) {
        const unsigned P = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const unsigned t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    unsigned t = 0u;
#pragma unroll
    for (unsigned Q = 512u; Q > 1u; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1u;
    return (ssf_spread10(X[0] ^ t) << 2) | (ssf_spread10(X[1] ^ t) << 1) | ssf_spread10(X[2] ^ t);
}

// Ascending in-place sort of n 64-bit keys in shared memory by a whole CTA (n a power of two >= 256, nthreads a multiple of 32;
// every thread of the CTA calls it).  Bitonic network, but only the exchanges at distance >= 256 go through shared memory with
// a CTA barrier per stage: a warp takes a run of 256 consecutive keys into registers (lane l holds keys l, l + 32, ...), where
// distances 32-128 are exchanges between a thread's own registers and distances 1-16 are warp shuffles -- all merges up to
// size 256 in one visit, and the last eight stages of every larger merge in one more.  8192 keys: 21 barrier rounds instead
// of 91.
template <int FIRST>   // stages at distances FIRST, FIRST / 2, ..., 1 of the merge of size `size` on a 256-key run starting at `base`
__device__ __forceinline__ void ssf_sort_run_stages(unsigned long long (&e)[8], int base, int lane, int size) {
#pragma unroll
    for (int stride = FIRST; stride >= 32; stride >>= 1) {
        const int dj = stride >> 5;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if ((j & dj) == 0) {
                const bool up = ((base + j * 32 + lane) & size) == 0;
                const unsigned long long a = e[j], b = e[j | dj];
                if ((a > b) == up) {
                    e[j] = b;
                    e[j | dj] = a;
                }
            }
        }
    }
#pragma unroll
    for (int stride = FIRST < 16 ? FIRST : 16; stride >= 1; stride >>= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, e[j], stride);
            const bool up = ((base + j * 32 + lane) & size) == 0;
            const bool keep_min = ((lane & stride) == 0) == up;
            e[j] = (o < e[j]) == keep_min ? o : e[j];
        }
    }
}

__device__ __forceinline__ void ssf_cta_sort_u64(unsigned long long* s, int n, int tid, int nthreads) {
    const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
    unsigned long long e[8];
    for (int base = warp * 256; base < n; base += nw * 256) {
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = s[base + j * 32 + lane];
        ssf_sort_run_stages<1>(e, base, lane, 2);
        ssf_sort_run_stages<2>(e, base, lane, 4);
        ssf_sort_run_stages<4>(e, base, lane, 8);
        ssf_sort_run_stages<8>(e, base, lane, 16);
        ssf_sort_run_stages<16>(e, base, lane, 32);
        ssf_sort_run_stages<32>(e, base, lane, 64);
        ssf_sort_run_stages<64>(e, base, lane, 128);
        ssf_sort_run_stages<128>(e, base, lane, 256);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[base + j * 32 + lane] = e[j];
    }
    __syncthreads();
    for (int size = 512; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride >= 256; stride >>= 1) {
            for (int t = tid; t < (n >> 1); t += nthreads) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = s[lo], b = s[hi];
                if ((a > b) == up) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
            __syncthreads();
        }
        for (int base = warp * 256; base < n; base += nw * 256) {
#pragma unroll
            for (int j = 0; j < 8; ++j) e[j] = s[base + j * 32 + lane];
            ssf_sort_run_stages<128>(e, base, lane, size);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[base + j * 32 + lane] = e[j];
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float ssf_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float ssf_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + bulk async copy (TMA engine, 1-D; SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ uint32_t ssf_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void ssf_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ssf_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ssf_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ssf_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ssf_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void ssf_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(ssf_smem_u32(bar)),
        "r"(parity)
        : "memory");
    // (no explicit suspend-time hint: a 10 ms hint measured +0.4 %, within noise, and a wake-up that arrived late would cost the
    // whole hint; the default time limit bounds a missed notification to microseconds)
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void ssf_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ssf_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(ssf_smem_u32(bar))
                 : "memory");
}
