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

// every kernel launch site ends with PLB_LAUNCH_CHECK(): it also feeds plb_launch_counters()
void count_launch(int resident);
#define PLB_LAUNCH_CHECK()          \
    do {                            \
        ::plb::count_launch(0);     \
        PLB_CUDA(cudaGetLastError()); \
    } while (0)

int sm_count();  // cached per device

// Programmatic dependent launch (plb_set_tuning key 18, default on).  The streaming kernels of the batch chain are
// launched with cudaLaunchAttributeProgrammaticStreamSerialization: every CTA signals `launch_dependents` on entry, so
// the NEXT kernel of the stream is scheduled, runs its prologue (barrier init, descriptor prefetch) and parks in
// `griddepcontrol.wait` while this one is still streaming -- the ~2 us launch gap at each kernel boundary is hidden.
// Rule for every kernel launched this way: ALL threads execute griddep_wait() before their first global access and
// before any early return (the completion of a grid must imply the completion of its predecessor).
int pdl_enabled();
void set_pdl_enabled(int on);

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

// Division-free reduction currency: sums of (x - piv) and (x - piv)^2 over n elements.
// Two Sums with the same pivot add component-wise; `rebase` moves a Sums to another pivot
// exactly (up to float64 rounding) without a division, so whole block / grid reductions are
// plain additions and only the very last thread divides (to_moments).
struct Sums {
    double n, s1, s2, piv;
};

__host__ __device__ __forceinline__ Sums rebase(Sums a, double P)
{
    const double d = a.piv - P;
    a.s2 = a.s2 + 2.0 * d * a.s1 + a.n * d * d;
    a.s1 = a.s1 + a.n * d;
    a.piv = P;
    return a;
}

__host__ __device__ __forceinline__ Moments to_moments(const Sums &a)
{
    Moments m;
    m.n = a.n;
    if (a.n == 0.0) {
        m.mean = 0.0; m.m2 = 0.0;
    } else {
        m.mean = a.piv + a.s1 / a.n;
        m.m2 = a.s2 - a.s1 * a.s1 / a.n;
        if (m.m2 < 0.0) m.m2 = 0.0;
    }
    return m;
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
    __device__ __forceinline__ Sums sums() const
    {
        Sums r;
        r.n = (double)n; r.s1 = s1; r.s2 = s2; r.piv = K;
        return r;
    }
    __device__ __forceinline__ Moments finish() const { return to_moments(sums()); }
};

#ifdef __CUDACC__
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// <<<grid, block, smem, stream>>> with the programmatic-serialization attribute (when enabled)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Block-wide reduction of per-thread Sums without divisions: agree on a pivot (the pivot of the
// lowest-numbered thread that saw data), rebase, add.  Result valid in thread 0.
// `scratch` needs >= 32 Sums; all threads of the block must call.
__device__ __forceinline__ Sums block_reduce_sums(Sums a, Sums *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    // warp pivot: first lane with data
    const unsigned has = __ballot_sync(0xffffffffu, a.n > 0.0);
    const int src = has ? (__ffs(has) - 1) : 0;
    const double wp = __shfl_sync(0xffffffffu, a.piv, src);
    if (lane == 0) {
        scratch[warp].piv = wp;
        scratch[warp].n = has ? 1.0 : 0.0;
    }
    __syncthreads();
    double P = 0.0;
    for (int w = 0; w < nwarps; ++w) {
        if (scratch[w].n > 0.0) { P = scratch[w].piv; break; }
    }
    __syncthreads();
    a = rebase(a, P);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a.n += __shfl_down_sync(0xffffffffu, a.n, d);
        a.s1 += __shfl_down_sync(0xffffffffu, a.s1, d);
        a.s2 += __shfl_down_sync(0xffffffffu, a.s2, d);
    }
    if (lane == 0) scratch[warp] = a;
    __syncthreads();
    if (warp == 0) {
        Sums z;
        z.n = 0.0; z.s1 = 0.0; z.s2 = 0.0; z.piv = P;
        a = (lane < nwarps) ? scratch[lane] : z;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            a.n += __shfl_down_sync(0xffffffffu, a.n, d);
            a.s1 += __shfl_down_sync(0xffffffffu, a.s1, d);
            a.s2 += __shfl_down_sync(0xffffffffu, a.s2, d);
        }
        a.piv = P;
    }
    return a;
}

// Cheaper variant for callers whose pivot is already uniform within each warp: plain shuffle adds,
// one barrier, then thread 0 rebases the (<= 32) warp totals in warp order.  Result valid in thread 0.
__device__ __forceinline__ Sums block_reduce_sums_warp_pivot(Sums a, Sums *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a.n += __shfl_down_sync(0xffffffffu, a.n, d);
        a.s1 += __shfl_down_sync(0xffffffffu, a.s1, d);
        a.s2 += __shfl_down_sync(0xffffffffu, a.s2, d);
    }
    if (lane == 0) scratch[warp] = a;
    __syncthreads();
    Sums acc;
    acc.n = 0.0; acc.s1 = 0.0; acc.s2 = 0.0; acc.piv = 0.0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < nwarps; ++w) {
            Sums q = scratch[w];
            if (q.n > 0.0) {
                if (acc.n == 0.0) acc = q;
                else {
                    q = rebase(q, acc.piv);
                    acc.n += q.n; acc.s1 += q.s1; acc.s2 += q.s2;
                }
            }
        }
    }
    return acc;
}

__device__ __forceinline__ bool is_nan_bound(double b) { return b != b; }
#endif

}  // namespace plb
