// plb_common.cuh -- shared helpers for libparllel_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/parllel_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libparllel_b200 is written for sm_100a (B200) only"
#endif

namespace plb {

// ---------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define PLB_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            ::plb::set_error(__VA_ARGS__); \
            return (code);                \
        }                                 \
    } while (0)

#define PLB_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::plb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                             __FILE__, __LINE__);                                        \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

#define PLB_LAUNCH_CHECK() PLB_CUDA(cudaGetLastError())

int sm_count();  // cached per device

static inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------
// (count, mean, M2) moments with the Chan/Welford parallel merge, float64 throughout
// ---------------------------------------------------------------------------------
struct Moments {
    double n, mean, m2;
};

__host__ __device__ __forceinline__ Moments merge(const Moments &a, const Moments &b)
{
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    const double tot = a.n + b.n;
    const double d = b.mean - a.mean;
    Moments r;
    r.n = tot;
    r.mean = a.mean + d * (b.n / tot);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / tot);
    return r;
}

// Per-thread accumulator: sums of (x - K) and (x - K)^2 around a pivot K (the first
// element seen), so the later conversion M2 = s2 - s1^2/n does not cancel.
struct PivotAcc {
    double K, s1, s2;
    unsigned n;
    __device__ __forceinline__ void init()
    {
        K = 0.0; s1 = 0.0; s2 = 0.0; n = 0u;
    }
    __device__ __forceinline__ void add(float xf)
    {
        const double x = (double)xf;
        if (n == 0u) K = x;
        const double d = x - K;
        s1 += d;
        s2 = fma(d, d, s2);
        ++n;
    }
    __device__ __forceinline__ Moments finish() const
    {
        Moments m;
        m.n = (double)n;
        if (n == 0u) {
            m.mean = 0.0; m.m2 = 0.0;
        } else {
            m.mean = K + s1 / m.n;
            m.m2 = s2 - s1 * s1 / m.n;
            if (m.m2 < 0.0) m.m2 = 0.0;
        }
        return m;
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ Moments shfl_down(const Moments &m, int delta)
{
    Moments r;
    r.n = __shfl_down_sync(0xffffffffu, m.n, delta);
    r.mean = __shfl_down_sync(0xffffffffu, m.mean, delta);
    r.m2 = __shfl_down_sync(0xffffffffu, m.m2, delta);
    return r;
}

// Fixed-shape tree: lane 0 ends up with merge over lanes in a deterministic order.
__device__ __forceinline__ Moments warp_merge(Moments m)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = merge(m, shfl_down(m, d));
    return m;
}

// Block-wide merge; result valid in thread 0.  `scratch` needs >= 32 Moments.
__device__ __forceinline__ Moments block_merge(Moments m, Moments *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    m = warp_merge(m);
    if (lane == 0) scratch[warp] = m;
    __syncthreads();
    if (warp == 0) {
        Moments z;
        z.n = 0.0; z.mean = 0.0; z.m2 = 0.0;
        m = (lane < nwarps) ? scratch[lane] : z;
        m = warp_merge(m);
    }
    return m;
}

__device__ __forceinline__ bool is_nan_bound(double b) { return b != b; }
#endif

}  // namespace plb
