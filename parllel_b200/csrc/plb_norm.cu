// plb_norm.cu -- elementwise normalisation / clipping passes and the NormalizeObservations step.
//   reward scale + clip      parllel/transforms/norm_rewards.py:114, clip_rewards.py:36
//   advantage normalisation  parllel/transforms/norm_advantage.py:51-54
//   observation normalise    parllel/transforms/norm_obs.py:57-76
#include <math.h>

#include "plb_common.cuh"
#include "plb_scan_shared.cuh"

namespace plb {

// Generic 2-D elementwise driver: F(value, column) over a [rows, cols] region in place,
// 128-bit accesses when the region allows it.
template <bool VEC4, class F>
__global__ void __launch_bounds__(256)
elementwise_kernel(float *__restrict__ x, int64_t ld_x, int64_t rows, int64_t cols, F f)
{
    f.prepare();  // per-launch constants -> registers (the stores to x could alias them otherwise)
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    if (VEC4) {
        const int64_t vpr = cols >> 2;
        const int64_t total = rows * vpr;
        constexpr int U = 4;
        for (int64_t base = tid; base < total; base += nthreads * U) {
            float4 v[U];
            int64_t off[U], col[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + (int64_t)u * nthreads;
                if (i < total) {
                    const int64_t r = (rows == 1) ? 0 : i / vpr;
                    col[u] = (i - r * vpr) << 2;
                    off[u] = r * ld_x + col[u];
                    v[u] = *reinterpret_cast<const float4 *>(x + off[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + (int64_t)u * nthreads;
                if (i < total) {
                    float4 o;
                    o.x = f(v[u].x, col[u]);
                    o.y = f(v[u].y, col[u] + 1);
                    o.z = f(v[u].z, col[u] + 2);
                    o.w = f(v[u].w, col[u] + 3);
                    *reinterpret_cast<float4 *>(x + off[u]) = o;
                }
            }
        }
    } else {
        const int64_t total = rows * cols;
        for (int64_t i = tid; i < total; i += nthreads) {
            const int64_t r = (rows == 1) ? 0 : i / cols;
            const int64_t c = i - r * cols;
            x[r * ld_x + c] = f(x[r * ld_x + c], c);
        }
    }
}

// grid cap of the streaming elementwise kernels, in CTAs of 256 threads per SM (C1 step: 21.0 us with 4,
// 22.0 with 8, 22.4 with 2); plb_set_tuning key 11
static int g_normadv_ctas_per_sm = 4;
void set_normadv_ctas_per_sm(int v) { g_normadv_ctas_per_sm = v > 0 ? v : 4; }
static int g_normadv_reduce_mode = 0;     // experiments: 1 = serial partial loads in the normalise prologue (key 21)
void set_normadv_reduce_mode(int v) { g_normadv_reduce_mode = v; }
// CTA size of the sharded normalise kernel (plb_set_tuning key 20).  256 is the configuration of the best 8-GPU
// measurement (26.8 us per step); 512 / 1024 (fewer polling warps per SM) measured the same at two GPUs and were
// not measured in isolation at eight.
static int g_normadv_xch_threads = 256;
void set_normadv_xch_threads(int v) { g_normadv_xch_threads = (v == 256 || v == 512 || v == 1024) ? v : 256; }

template <class F>
static int launch_elementwise(float *x, int64_t ld_x, int64_t rows, int64_t cols, bool col_dependent, F f,
                              cudaStream_t stream)
{
    if (rows * cols == 0) return PLB_OK;
    if (ld_x == cols && !col_dependent) {
        cols = rows * cols;
        rows = 1;
        ld_x = cols;
    }
    const bool vec = (cols % 4 == 0) && (ld_x % 4 == 0) && aligned(x, 16);
    const int64_t total = rows * cols;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    int64_t blocks = ceil_div(vec ? total / 4 : total, 256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)sms * g_normadv_ctas_per_sm) blocks = (int64_t)sms * g_normadv_ctas_per_sm;
    if (vec) elementwise_kernel<true, F><<<(unsigned)blocks, 256, 0, stream>>>(x, ld_x, rows, cols, f);
    else elementwise_kernel<false, F><<<(unsigned)blocks, 256, 0, stream>>>(x, ld_x, rows, cols, f);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

// reward = clip( float32( float64(reward) / sqrt(var + eps) ) )
struct ScaleClipOp {
    const double *var;
    double eps;
    float lo, hi;
    double denom;
    __device__ __forceinline__ void prepare() { denom = (var != nullptr) ? 1.0 / sqrt(__dadd_rn(*var, eps)) : 1.0; }
    __device__ __forceinline__ float operator()(float x, int64_t) const
    {
        if (var != nullptr) {
            // float64 product with the reciprocal of sqrt(var + eps): <= 1 ulp(float64) from the reference's
            // float64 division (norm_rewards.py:114) before the single rounding to float32
            x = __double2float_rn(__dmul_rn((double)x, denom));
        }
        return fminf(fmaxf(x, lo), hi);
    }
};

// advantage = (adv - mean32) / (std32 + eps32), float32 arithmetic (norm_advantage.py:51-54)
struct NormAdvOp {
    const double *moments;  // [inner][3]
    int64_t inner;
    float eps;
    float mean0, denom0, rdenom0;    // single-agent case, hoisted
    __device__ __forceinline__ void prepare()
    {
        mean0 = __double2float_rn(moments[1]);
        denom0 = __fadd_rn(__double2float_rn(sqrt(moments[2] / moments[0])), eps);
        rdenom0 = __frcp_rn(denom0);
    }
    __device__ __forceinline__ float operator()(float x, int64_t col) const
    {
        if (inner == 1) return div_by(__fsub_rn(x, mean0), denom0, rdenom0);
        const double *m = moments + 3 * (col % inner);
        const float mean = __double2float_rn(m[1]);
        const float stdv = __double2float_rn(sqrt(m[2] / m[0]));
        return __fdiv_rn(__fsub_rn(x, mean), __fadd_rn(stdv, eps));
    }
};

// obs = float32( (float64(obs) - mean[f]) / sqrt(var[f] + eps) )   (norm_obs.py:72-74)
struct NormColsOp {
    const double *mean, *var;
    double eps;
    __device__ __forceinline__ void prepare() {}
    __device__ __forceinline__ float operator()(float x, int64_t col) const
    {
        return __double2float_rn(__ddiv_rn(__dsub_rn((double)x, mean[col]), sqrt(__dadd_rn(var[col], eps))));
    }
};

// NormalizeAdvantage tail fed directly by the partial Sums the advantage scan published: each CTA
// reduces the (<= few hundred) partials itself while its first loads are in flight, so the scan needs
// no serial last-CTA epilogue.  CTA 0 also writes the final (count, mean, M2) for API consumers.
// CONTIG (dense [T, B] leaves, seen as one row): no per-element row arithmetic.  (Capping it at 32 registers so that
// four of these CTAs fit next to a resident scan CTA spilled the first batch of loads: not kept.)
template <bool CONTIG>
__global__ void __launch_bounds__(256, 4)
normalize_advantage_partials_kernel(float *__restrict__ x, int64_t ld_x, int64_t rows, int64_t cols,
                                    const Sums *__restrict__ partials, const unsigned *__restrict__ n_partials,
                                    float eps, double *__restrict__ moments_out, int reduce_mode)
{
    __shared__ float s_mean, s_denom;
    griddep_launch_dependents();
    griddep_wait();  // the scan that produced x and the partials has completed
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t vpr = cols >> 2;
    const int64_t total = rows * vpr;
    constexpr int U = 4;
    auto offset_of = [&](int64_t i) -> int64_t {
        if (CONTIG) return i << 2;
        const int64_t r = (rows == 1) ? 0 : i / vpr;
        return r * ld_x + ((i - r * vpr) << 2);
    };
    // first batch of loads before the reduction, so the two overlap
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = tid + (int64_t)u * nthreads;
        if (i < total) v[u] = *reinterpret_cast<const float4 *>(x + offset_of(i));
    }
    if (threadIdx.x < 32) {  // one warp reduces the partial sums; the others go straight to the barrier
        const Moments m = reduce_mode == 1 ? reduce_published_partials_warp_serial(partials, *n_partials)
                                            : reduce_published_partials_warp(partials, n_partials);
        if (threadIdx.x == 0) {
            s_mean = __double2float_rn(m.mean);
            s_denom = __fadd_rn(__double2float_rn(sqrt(m.m2 / m.n)), eps);
            if (blockIdx.x == 0) {
                moments_out[0] = m.n; moments_out[1] = m.mean; moments_out[2] = m.m2;
            }
        }
    }
    __syncthreads();
    const float mean = s_mean, denom = s_denom, rdenom = __frcp_rn(s_denom);
    for (int64_t base = tid; base < total; base += nthreads * U) {
        if (base != tid) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + (int64_t)u * nthreads;
                if (i < total) v[u] = *reinterpret_cast<const float4 *>(x + offset_of(i));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * nthreads;
            if (i < total) {
                float4 o;
                o.x = div_by(__fsub_rn(v[u].x, mean), denom, rdenom);
                o.y = div_by(__fsub_rn(v[u].y, mean), denom, rdenom);
                o.z = div_by(__fsub_rn(v[u].z, mean), denom, rdenom);
                o.w = div_by(__fsub_rn(v[u].w, mean), denom, rdenom);
                *reinterpret_cast<float4 *>(x + offset_of(i)) = o;
            }
        }
    }
}

// Multi-GPU variant: the producer's tail pushed this rank's moments to every peer; each CTA waits
// for all ranks' flags and merges in rank order in its prologue (warp 0), the first loads in flight.
// (CTAs of up to 1024 threads can be selected with plb_set_tuning key 20: a quarter of the polling warps.)
__global__ void __launch_bounds__(1024)
normalize_advantage_xch_kernel(float *__restrict__ x, int64_t ld_x, int64_t rows, int64_t cols, const XchRef xch,
                               float eps, double *__restrict__ moments_out, unsigned long long *dbg)
{
    __shared__ float s_mean, s_denom;
    griddep_launch_dependents();
    griddep_wait();  // this rank's scan (x, the epoch word) has completed; the peers' records are polled below
    if (dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0) dbg[4] = global_timer_ns();
    // The wait comes first, on an idle memory pipe: a hop costs ~0.25 us there, but microseconds when the polls queue
    // behind the ~10 MB of first-wave streaming loads of the grid (measured on the grid-exchange consumer that was
    // tried in round 2: 3.3 us to read one word, 4.6 us for the records, with the loads issued first).
    if (threadIdx.x < 32) {
        const Moments m = xch_wait_merge3(xch);
        if (threadIdx.x == 0) {
            if (dbg != nullptr && blockIdx.x == 0) dbg[5] = global_timer_ns();
            s_mean = __double2float_rn(m.mean);
            s_denom = __fadd_rn(__double2float_rn(sqrt(m.m2 / m.n)), eps);
            if (blockIdx.x == 0) {
                moments_out[0] = m.n; moments_out[1] = m.mean; moments_out[2] = m.m2;
            }
        }
    }
    __syncthreads();
    const float mean = s_mean, denom = s_denom, rdenom = __frcp_rn(s_denom);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t vpr = cols >> 2;
    const int64_t total = rows * vpr;
    constexpr int U = 4;
    for (int64_t base = tid; base < total; base += nthreads * U) {
        float4 v[U];
        int64_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * nthreads;
            if (i < total) {
                const int64_t r = (rows == 1) ? 0 : i / vpr;
                off[u] = r * ld_x + ((i - r * vpr) << 2);
                v[u] = *reinterpret_cast<const float4 *>(x + off[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * nthreads;
            if (i < total) {
                float4 o;
                o.x = div_by(__fsub_rn(v[u].x, mean), denom, rdenom);
                o.y = div_by(__fsub_rn(v[u].y, mean), denom, rdenom);
                o.z = div_by(__fsub_rn(v[u].z, mean), denom, rdenom);
                o.w = div_by(__fsub_rn(v[u].w, mean), denom, rdenom);
                *reinterpret_cast<float4 *>(x + off[u]) = o;
            }
        }
    }
    if (dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0) dbg[6] = global_timer_ns();
}

int normalize_advantage_after_exchange(float *x, int64_t ld_x, int64_t rows, int64_t cols, const XchRef &xch, double eps,
                                       double *moments_out, cudaStream_t stream)
{
    if (rows * cols == 0) return PLB_OK;
    if (ld_x == cols) {
        cols = rows * cols;
        rows = 1;
        ld_x = cols;
    }
    if (!((cols % 4 == 0) && (ld_x % 4 == 0) && aligned(x, 16))) return PLB_ERR_UNSUPPORTED;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    const int threads = g_normadv_xch_threads;
    const int ctas_per_sm = (256 * g_normadv_ctas_per_sm + threads - 1) / threads;  // same threads per SM as the one-GPU kernel
    int64_t blocks = ceil_div(rows * cols / 4, (int64_t)threads * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)sms * ctas_per_sm) blocks = (int64_t)sms * ctas_per_sm;
    PLB_CUDA(launch_pdl(normalize_advantage_xch_kernel, dim3((unsigned)blocks), dim3(threads), 0, stream, x, ld_x, rows, cols, xch,
                        (float)eps, moments_out, debug_buffer()));
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

// returns PLB_ERR_UNSUPPORTED when the region cannot take the 128-bit path
int normalize_advantage_from_partials(float *x, int64_t ld_x, int64_t rows, int64_t cols, void *tail_workspace,
                                      double eps, double *moments_out, cudaStream_t stream)
{
    if (rows * cols == 0) return PLB_OK;
    if (ld_x == cols) {
        cols = rows * cols;
        rows = 1;
        ld_x = cols;
    }
    if (!((cols % 4 == 0) && (ld_x % 4 == 0) && aligned(x, 16))) return PLB_ERR_UNSUPPORTED;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    int64_t blocks = ceil_div(rows * cols / 4, 256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)sms * g_normadv_ctas_per_sm) blocks = (int64_t)sms * g_normadv_ctas_per_sm;
    const MomentsTail t = make_tail(moments_out, tail_workspace);
    if (rows == 1)
        PLB_CUDA(launch_pdl(normalize_advantage_partials_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, stream, x, ld_x, rows,
                            cols, t.partials, t.ticket + 1, (float)eps, moments_out, g_normadv_reduce_mode));
    else
        PLB_CUDA(launch_pdl(normalize_advantage_partials_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, stream, x, ld_x, rows,
                            cols, t.partials, t.ticket + 1, (float)eps, moments_out, g_normadv_reduce_mode));
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

// column-tiled normalisation with the per-column (mean, 1/denominator) hoisted out of the row loop
template <int VEC>
__global__ void __launch_bounds__(128)
normalize_columns_kernel(float *__restrict__ x, int64_t ld_x, int64_t rows, int64_t cols,
                         const double *__restrict__ mean, const double *__restrict__ var, double eps,
                         int64_t rows_per_split, float clip_lo, float clip_hi)
{
    const int64_t c0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (c0 >= cols) return;
    double mu[VEC], den[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        mu[k] = mean[c0 + k];
        den[k] = 1.0 / sqrt(__dadd_rn(var[c0 + k], eps));  // multiplied below: <= 1 ulp(float64) from the division
    }
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_hi = (r_lo + rows_per_split < rows) ? r_lo + rows_per_split : rows;
    constexpr int U = 4;
    for (int64_t r = r_lo; r < r_hi; r += U) {
        float v[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r + u < r_hi) {
                if (VEC == 4) {
                    const float4 q = *reinterpret_cast<const float4 *>(x + (r + u) * ld_x + c0);
                    v[u][0] = q.x; v[u][1 % VEC] = q.y; v[u][2 % VEC] = q.z; v[u][3 % VEC] = q.w;
                } else {
                    v[u][0] = x[(r + u) * ld_x + c0];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r + u < r_hi) {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    v[u][k] = fminf(fmaxf(__double2float_rn(__dmul_rn(__dsub_rn((double)v[u][k], mu[k]), den[k])), clip_lo), clip_hi);
                if (VEC == 4) {
                    float4 q;
                    q.x = v[u][0]; q.y = v[u][1 % VEC]; q.z = v[u][2 % VEC]; q.w = v[u][3 % VEC];
                    *reinterpret_cast<float4 *>(x + (r + u) * ld_x + c0) = q;
                } else {
                    x[(r + u) * ld_x + c0] = v[u][0];
                }
            }
        }
    }
}

int normalize_columns(float *x, int64_t ld_x, int64_t rows, int64_t cols, const double *mean, const double *var,
                      double eps, float clip_lo, float clip_hi, cudaStream_t stream)
{
    PLB_REQUIRE(rows >= 0 && cols >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (rows * cols == 0) return PLB_OK;
    PLB_REQUIRE(x && mean && var, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(ld_x >= cols, PLB_ERR_INVALID_ARGUMENT, "row stride smaller than row length");
    const int splits = column_moments_splits(rows, cols);
    const int64_t rows_per_split = ceil_div(rows, splits);
    const bool vec = (cols % 4 == 0) && (ld_x % 4 == 0) && aligned(x, 16);
    if (vec) {
        dim3 grid((unsigned)ceil_div(cols / 4, 128), (unsigned)splits);
        normalize_columns_kernel<4><<<grid, 128, 0, stream>>>(x, ld_x, rows, cols, mean, var, eps, rows_per_split, clip_lo, clip_hi);
    } else {
        dim3 grid((unsigned)ceil_div(cols, 128), (unsigned)splits);
        normalize_columns_kernel<1><<<grid, 128, 0, stream>>>(x, ld_x, rows, cols, mean, var, eps, rows_per_split, clip_lo, clip_hi);
    }
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

// finalize of the observation step: merge row splits, fold into the running state
// (update_from_moments), leaving mean/var updated for the normalise pass.  `count` is
// read by every CTA and written by nobody here; a follow-up single-thread kernel bumps it.
__global__ void __launch_bounds__(256)
obs_finalize_update_kernel(const Moments *__restrict__ partials, int64_t cols, int n_splits,
                           double *mean, double *var, const double *count, double *batch_count_out, int flags)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    Moments acc = partials[c];
    for (int s = 1; s < n_splits; ++s) acc = merge(acc, partials[(int64_t)s * cols + c]);
    const double cnt = count[0];
    // running_mean_std.py:23-30 with batch_var = M2 / n
    double batch_mean = acc.mean;
    double batch_var = acc.m2 / acc.n;
    double m_b;
    if (flags & PLB_MOMENTS_F32_FLOW) {
        batch_mean = (double)__double2float_rn(batch_mean);
        const float bv32 = __double2float_rn(batch_var);
        m_b = (double)__fmul_rn(bv32, __double2float_rn(acc.n));
    } else {
        m_b = __dmul_rn(batch_var, acc.n);
    }
    const double m = mean[c], v = var[c];
    const double delta = __dsub_rn(batch_mean, m);
    const double total = __dadd_rn(acc.n, cnt);
    mean[c] = __dadd_rn(m, __ddiv_rn(__dmul_rn(delta, acc.n), total));
    const double cross = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), cnt), acc.n), total);
    var[c] = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(v, cnt), m_b), cross), total);
    if (c == 0) batch_count_out[0] = acc.n;
}

__global__ void bump_count_kernel(double *count, const double *batch_count) { count[0] = batch_count[0] + count[0]; }

}  // namespace plb

extern "C" {

int plb_scale_clip_f32(float *x, int64_t ld_x, int64_t rows, int64_t cols, const double *var, double eps,
                       double clip_lo, double clip_hi, plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(rows >= 0 && cols >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (rows * cols == 0) return PLB_OK;
    PLB_REQUIRE(x != nullptr && ld_x >= cols, PLB_ERR_INVALID_ARGUMENT, "bad array");
    ScaleClipOp op;
    op.var = var;
    op.eps = eps;
    op.lo = (clip_lo != clip_lo) ? -INFINITY : (float)clip_lo;
    op.hi = (clip_hi != clip_hi) ? INFINITY : (float)clip_hi;
    return launch_elementwise(x, ld_x, rows, cols, false, op, (cudaStream_t)stream);
}

int plb_normalize_advantage_f32(float *advantage, int64_t ld, int64_t rows, int64_t cols, int64_t inner,
                                const double *moments, double eps, plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(rows >= 0 && cols >= 0 && inner >= 1, PLB_ERR_INVALID_ARGUMENT, "bad extent");
    if (rows * cols == 0) return PLB_OK;
    PLB_REQUIRE(advantage && moments && ld >= cols && cols % inner == 0, PLB_ERR_INVALID_ARGUMENT, "bad array");
    NormAdvOp op;
    op.moments = moments;
    op.inner = inner;
    op.eps = (float)eps;
    return launch_elementwise(advantage, ld, rows, cols, inner != 1, op, (cudaStream_t)stream);
}

int plb_normalize_columns_f32(float *x, int64_t ld_x, int64_t rows, int64_t cols, const double *mean,
                              const double *var, double eps, plb_stream_t stream)
{
    return plb::normalize_columns(x, ld_x, rows, cols, mean, var, eps, -INFINITY, INFINITY, (cudaStream_t)stream);
}

int plb_normalize_columns_clip_f32(float *x, int64_t ld_x, int64_t rows, int64_t cols, const double *mean,
                                   const double *var, double eps, double clip_lo, double clip_hi, plb_stream_t stream)
{
    const float lo = (clip_lo != clip_lo) ? -INFINITY : (float)clip_lo;
    const float hi = (clip_hi != clip_hi) ? INFINITY : (float)clip_hi;
    return plb::normalize_columns(x, ld_x, rows, cols, mean, var, eps, lo, hi, (cudaStream_t)stream);
}

size_t plb_normalize_observations_workspace_bytes(int64_t rows, int64_t cols)
{
    if (rows < 0 || cols <= 0) return 64;
    return 64 + sizeof(plb::Moments) * (size_t)plb::column_moments_splits(rows, cols) * (size_t)cols;
}

int plb_normalize_observations_step_f32(float *obs, int64_t ld_obs, int64_t rows, int64_t cols,
                                        const uint8_t *row_valid, double *mean, double *var, double *count,
                                        double eps, int flags, double clip_lo_, double clip_hi_, void *workspace,
                                        size_t workspace_bytes, plb_stream_t stream_)
{
    using namespace plb;
    cudaStream_t stream = (cudaStream_t)stream_;
    const float clip_lo = (clip_lo_ != clip_lo_) ? -INFINITY : (float)clip_lo_;
    const float clip_hi = (clip_hi_ != clip_hi_) ? INFINITY : (float)clip_hi_;
    PLB_REQUIRE(rows >= 0 && cols >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (cols == 0) return PLB_OK;
    PLB_REQUIRE(obs && mean && var && count, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(ld_obs >= cols, PLB_ERR_INVALID_ARGUMENT, "row stride smaller than row length");
    const size_t need = plb_normalize_observations_workspace_bytes(rows, cols);
    PLB_REQUIRE(workspace != nullptr && workspace_bytes >= need, PLB_ERR_WORKSPACE,
                "observation step needs %zu workspace bytes, got %zu", need, workspace_bytes);
    double *batch_count = reinterpret_cast<double *>(workspace);
    Moments *partials = reinterpret_cast<Moments *>(reinterpret_cast<unsigned char *>(workspace) + 64);

    // 8 bytes of HBM traffic per element: a column slab is either held in shared memory (TMA in/out) or
    // streamed twice with the second pass served by L2
    {
        // measured at C2 (B = 1024, F = 21168): shared-memory slab filled by cp.async / drained by 128-bit
        // stores 39.5 us, the same slab through TMA 42.5 us, two-CTA cluster slab 53.9 us (removed), two-pass L2 slab
        // 55 us, three kernels 85 us -> prefer the load/store-unit slab, then TMA (no cols % 4 requirement),
        // then the L2 variant when the slab is too tall for shared memory
        int rc1 = PLB_ERR_UNSUPPORTED;
        const int variant = obs_variant();
        if (variant == 4 || variant == 0)
            rc1 = obs_step_lsu(obs, ld_obs, rows, cols, row_valid, mean, var, count, batch_count,
                               eps, flags, clip_lo, clip_hi, stream);
        if (rc1 == PLB_ERR_UNSUPPORTED && variant != 2)
            rc1 = obs_step_single_pass(obs, ld_obs, rows, cols, row_valid, mean, var, count, batch_count,
                                       eps, flags, clip_lo, clip_hi, stream);
        if (rc1 == PLB_ERR_UNSUPPORTED && variant != 1)
            rc1 = obs_step_two_pass_l2(obs, ld_obs, rows, cols, row_valid, mean, var, count, batch_count, eps, flags,
                                       clip_lo, clip_hi, stream);
        if (rc1 == PLB_OK) {
            bump_count_kernel<<<1, 1, 0, stream>>>(count, batch_count);
            PLB_LAUNCH_CHECK();
            return PLB_OK;
        }
        if (rc1 != PLB_ERR_UNSUPPORTED) return rc1;
    }

    const int splits = column_moments_splits(rows, cols);
    int rc = column_moments_partial(obs, ld_obs, rows, cols, row_valid, partials, splits, stream);
    if (rc) return rc;
    obs_finalize_update_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, stream>>>(partials, cols, splits, mean, var,
                                                                                   count, batch_count, flags);
    PLB_LAUNCH_CHECK();
    bump_count_kernel<<<1, 1, 0, stream>>>(count, batch_count);
    PLB_LAUNCH_CHECK();
    return normalize_columns(obs, ld_obs, rows, cols, mean, var, eps, clip_lo, clip_hi, stream);
}

}  // extern "C"
