// plb_obs.cu -- single-pass NormalizeObservations step (parllel/transforms/norm_obs.py:57-76):
//   stats.update(obs[valid]); obs[:] = (obs - stats.mean) / sqrt(stats.var + eps)
// for one [B, F] float32 step with per-feature running statistics (float64).
//
// Three designs live here; plb_norm.cu picks (default: variant D, then A, then B, then three kernels):
//   A  obs::obs_step_kernel            shared-memory slab, TMA in / TMA out
//   B  obs_l2::obs_step_l2_kernel      two passes over a 256-byte-wide slab, second one served by L2
//   D  obs_lsu::obs_step_lsu_kernel    shared-memory slab, cp.async in / 128-bit global stores out
//
// Variant A.  Each CTA owns a [B rows x 16 columns] slab: TMA brings the whole slab into shared memory once,
// the per-column batch moments are reduced from shared memory (pivot-shifted float64 sums,
// division-free tree), merged into the running state with the reference's Chan update
// (running_mean_std.py:23-30), the slab is normalised in place in shared memory and leaves by TMA
// store.  HBM traffic is the compulsory 8 bytes per element (the three-kernel fallback in
// plb_norm.cu moves 12).  Slabs are 64 KB at B = 1024, so three CTAs share an SM and one CTA's
// arithmetic overlaps the others' loads and stores; neighbouring CTAs take neighbouring column
// slabs, so together they stream whole rows.
#include <cuda.h>
#include <math.h>

#include "plb_common.cuh"
#include "plb_scan_shared.cuh"

namespace plb {
namespace obs {

constexpr int FC = 16;           // columns per slab (64-byte row segments)
constexpr int THREADS = 256;
constexpr int GROUPS = THREADS / FC;  // 16 row groups; thread (g, c) takes rows g, g+16, g+32, ...
constexpr int BOX_ROWS = 256;    // rows per TMA box (box dimensions are limited to 256)
constexpr int MAX_BUFS = 4;

struct Params {
    int64_t rows, cols;
    const uint8_t *row_valid;
    double *mean, *var;
    const double *count;
    double *batch_count_out;  // workspace slot read by the count-bump kernel
    double eps;
    float clip_lo, clip_hi;   // -inf / +inf when off
    int flags;
    int n_slabs;
    int n_boxes;              // ceil(rows / BOX_ROWS)
    int n_bufs;               // slab buffers in the shared-memory ring (>= 1)
    int slab_floats;          // floats per slab buffer = n_boxes * BOX_ROWS * FC
};

__global__ void __launch_bounds__(THREADS)
obs_step_kernel(const __grid_constant__ CUtensorMap tm, const Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);  // n_bufs x [n_boxes * BOX_ROWS][FC]
    __shared__ unsigned long long full[MAX_BUFS];
    __shared__ Sums part[GROUPS][FC];
    __shared__ double col_mean[FC], col_inv[FC];

    const int c = threadIdx.x % FC, g = threadIdx.x / FC;
    const int nrows = (int)p.rows;
    const int n_bufs = p.n_bufs;
    const uint32_t slab_tx = (uint32_t)p.slab_floats * sizeof(float);
    const int my_n = (p.n_slabs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    auto load_slab = [&](int j) {  // thread 0 only
        const int buf = j % n_bufs;
        const int c0 = ((int)blockIdx.x + j * (int)gridDim.x) * FC;
        float *dst = ring + (size_t)buf * p.slab_floats;
        mbar_arrive_expect_tx(&full[buf], slab_tx);
        for (int bx = 0; bx < p.n_boxes; ++bx) tma_load_2d(dst + (size_t)bx * BOX_ROWS * FC, &tm, c0, bx * BOX_ROWS, &full[buf]);
    };

    if (threadIdx.x == 0) {
        for (int k = 0; k < n_bufs; ++k) mbar_init(&full[k], 1);
        fence_proxy_async_smem();
        tma_prefetch_desc(&tm);
    }
    __syncthreads();
    const int lead = n_bufs > 1 ? n_bufs - 1 : 1;  // slabs requested ahead of the one being processed
    if (threadIdx.x == 0)
        for (int j = 0; j < lead && j < my_n; ++j) load_slab(j);

    const double cnt = p.count[0];
    for (int i = 0; i < my_n; ++i) {
        const int buf = i % n_bufs;
        const uint32_t parity = (uint32_t)((i / n_bufs) & 1);
        const int s = (int)blockIdx.x + i * (int)gridDim.x;
        const int c0 = s * FC;
        float (*slab)[FC] = reinterpret_cast<float (*)[FC]>(ring + (size_t)buf * p.slab_floats);
        const bool col_ok = (int64_t)c0 + c < p.cols;
        const double old_mean = (g == 0 && col_ok) ? p.mean[c0 + c] : 0.0;   // overlaps the load
        const double old_var = (g == 0 && col_ok) ? p.var[c0 + c] : 1.0;
        mbar_wait(&full[buf], parity);

        // ---- batch moments of this slab: rows g, g+GROUPS, ... of column c --------------
        // float32 pivot-shifted sums over runs of 16 rows, folded into float64 per run: one
        // conversion pair per 16 elements instead of one per element (F2F is a quarter-rate pipe)
        const float piv = (g < nrows) ? slab[g][c] : 0.f;  // pivot: any value of the column's magnitude
        double S1 = 0.0, S2 = 0.0;
        unsigned n_acc = 0;
        if (p.row_valid == nullptr) {
            int r0 = g;
            for (; r0 + 15 * GROUPS < nrows; r0 += GROUPS * 16) {  // full runs: no predicates
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float d = slab[r0 + k * GROUPS][c] - piv;
                    s1 += d;
                    s2 = fmaf(d, d, s2);
                }
                S1 += (double)s1;
                S2 += (double)s2;
                n_acc += 16;
            }
            float s1 = 0.f, s2 = 0.f;
            for (int r = r0; r < nrows; r += GROUPS) {
                const float d = slab[r][c] - piv;
                s1 += d;
                s2 = fmaf(d, d, s2);
                ++n_acc;
            }
            S1 += (double)s1;
            S2 += (double)s2;
        } else {
            for (int r0 = g; r0 < nrows; r0 += GROUPS * 16) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
                for (int k = 0; k < 16; ++k) {
                    const int r = r0 + k * GROUPS;
                    if (r < nrows && p.row_valid[r] != 0) {
                        const float d = slab[r][c] - piv;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                        ++n_acc;
                    }
                }
                S1 += (double)s1;
                S2 += (double)s2;
            }
        }
        Sums mine;
        mine.n = (double)n_acc; mine.s1 = S1; mine.s2 = S2; mine.piv = (double)piv;
        part[g][c] = mine;
        __syncthreads();
        if (g == 0) {
            // division-free merge of the row groups (fixed order), then the reference's update
            Sums tot = part[0][c];
            for (int k = 1; k < GROUPS; ++k) {
                Sums q = part[k][c];
                if (q.n > 0.0) {
                    if (tot.n == 0.0) tot = q;
                    else {
                        q = rebase(q, tot.piv);
                        tot.n += q.n; tot.s1 += q.s1; tot.s2 += q.s2;
                    }
                }
            }
            const Moments m = to_moments(tot);
            // running_mean_std.py:23-30 with batch_var = M2 / n (population variance)
            double batch_mean = m.mean;
            double batch_var = m.m2 / m.n;
            double m_b;
            if (p.flags & PLB_MOMENTS_F32_FLOW) {
                batch_mean = (double)__double2float_rn(batch_mean);
                const float bv32 = __double2float_rn(batch_var);
                m_b = (double)__fmul_rn(bv32, __double2float_rn(m.n));
            } else {
                m_b = __dmul_rn(batch_var, m.n);
            }
            const double delta = __dsub_rn(batch_mean, old_mean);
            const double total = __dadd_rn(m.n, cnt);
            const double new_mean = __dadd_rn(old_mean, __ddiv_rn(__dmul_rn(delta, m.n), total));
            const double cross = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), cnt), m.n), total);
            const double new_var = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(old_var, cnt), m_b), cross), total);
            if (col_ok) {
                p.mean[c0 + c] = new_mean;
                p.var[c0 + c] = new_var;
            }
            col_mean[c] = new_mean;
            col_inv[c] = 1.0 / sqrt(__dadd_rn(new_var, p.eps));
            if (s == 0 && c == 0) p.batch_count_out[0] = m.n;
        }
        __syncthreads();

        // ---- normalise the slab in place --------------------------------------------------
        // y = (x - mean) * inv with mean and inv = 1/sqrt(var+eps) carried as float32 hi+lo pairs:
        // (x - mean_hi) is exact or correctly rounded, the lo parts restore the float64 constants;
        // |y - y_float64| <= ~1.5 ulp(float32) of y, without any float64 conversion per element.
        const double mu = col_mean[c], inv = col_inv[c];
        const float mu_hi = __double2float_rn(mu), mu_lo = __double2float_rn(mu - (double)mu_hi);
        const float inv_hi = __double2float_rn(inv), inv_lo = __double2float_rn(inv - (double)inv_hi);
#pragma unroll 8
        for (int r = g; r < nrows; r += GROUPS) {
            const float d = (slab[r][c] - mu_hi) - mu_lo;
            slab[r][c] = fminf(fmaxf(fmaf(d, inv_hi, d * inv_lo), p.clip_lo), p.clip_hi);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int bx = 0; bx < p.n_boxes; ++bx) tma_store_2d(&tm, &slab[bx * BOX_ROWS][0], c0, bx * BOX_ROWS);
            tma_store_commit();
            // request the slab `lead` ahead: it lands in the buffer whose store was committed one
            // iteration ago (ring of >= 2 buffers) or in this very buffer (single buffer)
            const int j = i + lead;
            if (j < my_n) {
                if (n_bufs > 1) tma_store_wait_read<1>();
                else tma_store_wait_read<0>();
                load_slab(j);
            }
        }
    }
    if (threadIdx.x == 0) tma_store_wait<0>();
}

}  // namespace obs

// ---------------------------------------------------------------------------------------------
// Variant B: two passes over a [rows x 64 columns] slab, the second one served by L2.
// Each CTA owns 256-byte-wide column slabs (full DRAM bursts, plain 128-bit loads: the TMA unit
// tops out at ~4.4 TB/s for the 64-byte rows a shared-memory-resident slab allows at B = 1024,
// tools/microbench/tma_width.cu).  Pass 1 streams the slab from HBM and reduces the per-column
// moments; pass 2 re-reads it while it is still L2-resident (one CTA per SM keeps only ~40 MB of
// slabs in flight in the 126 MB L2), normalises and stores.  HBM traffic stays 8 B per element and
// there is no limit on the number of rows.
// ---------------------------------------------------------------------------------------------
namespace obs_l2 {

constexpr int WCOLS = 64;              // columns per slab: 16 lanes x float4 = 256 bytes per row
constexpr int THREADS = 256;
constexpr int LANES = WCOLS / 4;       // 16 threads cover one row
constexpr int GROUPS = THREADS / LANES;  // 16 row groups; thread (g, l) takes rows g, g+16, ...

struct Params {
    float *x; int64_t ld;
    int64_t rows, cols;
    const uint8_t *row_valid;
    double *mean, *var;
    const double *count;
    double *batch_count_out;
    double eps;
    float clip_lo, clip_hi;
    int flags;
    int n_slabs;
};

__global__ void __launch_bounds__(THREADS)
obs_step_l2_kernel(const Params p)
{
    __shared__ Sums part[GROUPS][WCOLS];
    __shared__ float4 s_mu_hi[LANES], s_mu_lo[LANES], s_inv_hi[LANES], s_inv_lo[LANES];
    const int l = threadIdx.x % LANES, g = threadIdx.x / LANES;
    const int nrows = (int)p.rows;
    const double cnt = p.count[0];
    for (int s = blockIdx.x; s < p.n_slabs; s += gridDim.x) {
        const int64_t c0 = (int64_t)s * WCOLS + 4 * l;
        const bool col_ok = c0 < p.cols;  // cols % 4 == 0: a float4 is entirely in or out
        float *base = p.x + c0;

        // ---- pass 1: per-column moments (float32 pivot-shifted sums per run of 8 rows, float64 across runs)
        float4 piv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok && g < nrows) piv = *reinterpret_cast<const float4 *>(base + (int64_t)g * p.ld);
        double S1[4] = {0.0, 0.0, 0.0, 0.0}, S2[4] = {0.0, 0.0, 0.0, 0.0};
        unsigned n_acc = 0;
        if (col_ok) {
            constexpr int RUN = 8;
            for (int r0 = g; r0 < nrows; r0 += GROUPS * RUN) {
                float4 v[RUN];
                bool ok[RUN];
#pragma unroll
                for (int k = 0; k < RUN; ++k) {
                    const int r = r0 + k * GROUPS;
                    ok[k] = r < nrows && (p.row_valid == nullptr || p.row_valid[r] != 0);
                    if (r < nrows) v[k] = *reinterpret_cast<const float4 *>(base + (int64_t)r * p.ld);
                }
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < RUN; ++k) {
                    if (ok[k]) {
                        const float d0 = v[k].x - piv.x, d1 = v[k].y - piv.y, d2 = v[k].z - piv.z, d3 = v[k].w - piv.w;
                        s1[0] += d0; s2[0] = fmaf(d0, d0, s2[0]);
                        s1[1] += d1; s2[1] = fmaf(d1, d1, s2[1]);
                        s1[2] += d2; s2[2] = fmaf(d2, d2, s2[2]);
                        s1[3] += d3; s2[3] = fmaf(d3, d3, s2[3]);
                        ++n_acc;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { S1[q] += (double)s1[q]; S2[q] += (double)s2[q]; }
            }
        }
        const float pv[4] = {piv.x, piv.y, piv.z, piv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            Sums m;
            m.n = (double)n_acc; m.s1 = S1[q]; m.s2 = S2[q]; m.piv = (double)pv[q];
            part[g][4 * l + q] = m;
        }
        __syncthreads();
        if (threadIdx.x < WCOLS) {
            const int c = threadIdx.x;
            const int64_t col = (int64_t)s * WCOLS + c;
            Sums tot = part[0][c];
            for (int k = 1; k < GROUPS; ++k) {
                Sums q = part[k][c];
                if (q.n > 0.0) {
                    if (tot.n == 0.0) tot = q;
                    else {
                        q = rebase(q, tot.piv);
                        tot.n += q.n; tot.s1 += q.s1; tot.s2 += q.s2;
                    }
                }
            }
            const Moments m = to_moments(tot);
            double mean = 0.0, var = 1.0;
            if (col < p.cols) {
                mean = p.mean[col];
                var = p.var[col];
                chan_update_scalar(m.n, m.mean, m.m2, cnt, p.flags, mean, var);
                p.mean[col] = mean;
                p.var[col] = var;
            }
            const double inv = 1.0 / sqrt(__dadd_rn(var, p.eps));
            const float mh = __double2float_rn(mean), ih = __double2float_rn(inv);
            reinterpret_cast<float *>(s_mu_hi)[c] = mh;
            reinterpret_cast<float *>(s_mu_lo)[c] = __double2float_rn(mean - (double)mh);
            reinterpret_cast<float *>(s_inv_hi)[c] = ih;
            reinterpret_cast<float *>(s_inv_lo)[c] = __double2float_rn(inv - (double)ih);
            if (s == 0 && c == 0) p.batch_count_out[0] = m.n;
        }
        __syncthreads();

        // ---- pass 2: re-read (L2), normalise with float32 hi+lo constants, store ----------------
        if (col_ok) {
            const float4 mh = s_mu_hi[l], ml = s_mu_lo[l], ih = s_inv_hi[l], il = s_inv_lo[l];
            constexpr int RUN = 8;
            for (int r0 = g; r0 < nrows; r0 += GROUPS * RUN) {
                float4 v[RUN];
#pragma unroll
                for (int k = 0; k < RUN; ++k) {
                    const int r = r0 + k * GROUPS;
                    if (r < nrows) v[k] = *reinterpret_cast<const float4 *>(base + (int64_t)r * p.ld);
                }
#pragma unroll
                for (int k = 0; k < RUN; ++k) {
                    const int r = r0 + k * GROUPS;
                    if (r < nrows) {
                        float4 o;
                        float d;
                        d = (v[k].x - mh.x) - ml.x; o.x = fminf(fmaxf(fmaf(d, ih.x, d * il.x), p.clip_lo), p.clip_hi);
                        d = (v[k].y - mh.y) - ml.y; o.y = fminf(fmaxf(fmaf(d, ih.y, d * il.y), p.clip_lo), p.clip_hi);
                        d = (v[k].z - mh.z) - ml.z; o.z = fminf(fmaxf(fmaf(d, ih.z, d * il.z), p.clip_lo), p.clip_hi);
                        d = (v[k].w - mh.w) - ml.w; o.w = fminf(fmaxf(fmaf(d, ih.w, d * il.w), p.clip_lo), p.clip_hi);
                        *reinterpret_cast<float4 *>(base + (int64_t)r * p.ld) = o;
                    }
                }
            }
        }
        __syncthreads();  // part[] / s_* are reused by the next slab
    }
}

}  // namespace obs_l2

// (Variant C of round 1 -- a [rows x 32 columns] slab split by rows over a two-CTA cluster with a DSMEM exchange of the
//  per-column sums -- measured 53.9 us at C2 against 39.5 us for variant D and was removed in round 2.)

// ---------------------------------------------------------------------------------------------
// Variant D: the shared-memory slab of variant A filled and drained by the load/store units instead
// of the TMA: cp.async (16 bytes per thread) in, 128-bit global stores straight from registers out.
// Same [rows x 16 columns] slab and three CTAs per SM; the pivot of every column is its row-0 value,
// so all partial sums of a column share it and merge by plain addition.
// ---------------------------------------------------------------------------------------------
namespace obs_lsu {

constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;

struct Params {
    float *x; int64_t ld;
    int64_t rows, cols;
    const uint8_t *row_valid;
    double *mean, *var;
    const double *count;
    double *batch_count_out;
    double eps;
    float clip_lo, clip_hi;
    int flags;
    int n_slabs;
};

// FC columns per slab; thread (g, l) takes the float4 lane l of rows g, g+GROUPS, ...
template <int FC>
__global__ void __launch_bounds__(THREADS)
obs_step_lsu_kernel(const Params p)
{
    constexpr int LANES = FC / 4;
    constexpr int GROUPS = THREADS / LANES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 (*slab)[LANES] = reinterpret_cast<float4 (*)[LANES]>(smem_raw);  // [rows][4 x float4]
    __shared__ double w_s1[NWARPS][FC], w_s2[NWARPS][FC];
    __shared__ unsigned w_n[NWARPS];
    __shared__ float4 s_mu_hi[LANES], s_mu_lo[LANES], s_inv_hi[LANES], s_inv_lo[LANES];

    const int l = threadIdx.x % LANES, g = threadIdx.x / LANES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nrows = (int)p.rows;
    const double cnt = p.count[0];
    if ((int)blockIdx.x < p.n_slabs && (int64_t)blockIdx.x * FC + 4 * l < p.cols) {
        const float *first = p.x + (int64_t)blockIdx.x * FC + 4 * l;
        for (int r = g; r < nrows; r += GROUPS) {  // all requests of the first slab in flight at once
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&slab[r][l]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(first + (int64_t)r * p.ld) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int s = blockIdx.x; s < p.n_slabs; s += gridDim.x) {
        const int64_t c0 = (int64_t)s * FC + 4 * l;
        const bool col_ok = c0 < p.cols;  // cols % 4 == 0: a float4 is entirely in or out
        float *base = p.x + c0;
        // ---- the slab was requested by this thread itself (before the loop, or while it drained the
        // previous slab): wait for the own requests, then for everybody's
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();

        // ---- per-column moments around the row-0 pivot -----------------------------------------
        const float4 piv = col_ok ? slab[0][l] : make_float4(0.f, 0.f, 0.f, 0.f);
        double S1[4] = {0.0, 0.0, 0.0, 0.0}, S2[4] = {0.0, 0.0, 0.0, 0.0};
        unsigned n_acc = 0;
        if (col_ok) {
            constexpr int RUN = 16;
            for (int r0 = g; r0 < nrows; r0 += GROUPS * RUN) {
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                for (int k = 0; k < RUN; ++k) {
                    const int r = r0 + k * GROUPS;
                    if (r < nrows && (p.row_valid == nullptr || p.row_valid[r] != 0)) {
                        const float4 v = slab[r][l];
                        const float d0 = v.x - piv.x, d1 = v.y - piv.y, d2 = v.z - piv.z, d3 = v.w - piv.w;
                        s1[0] += d0; s2[0] = fmaf(d0, d0, s2[0]);
                        s1[1] += d1; s2[1] = fmaf(d1, d1, s2[1]);
                        s1[2] += d2; s2[2] = fmaf(d2, d2, s2[2]);
                        s1[3] += d3; s2[3] = fmaf(d3, d3, s2[3]);
                        ++n_acc;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { S1[q] += (double)s1[q]; S2[q] += (double)s2[q]; }
            }
        }
        // the 8 row groups of a warp share their columns' pivots: plain shuffle adds
#pragma unroll
        for (int d = LANES; d < 32; d <<= 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                S1[q] += __shfl_xor_sync(0xffffffffu, S1[q], d);
                S2[q] += __shfl_xor_sync(0xffffffffu, S2[q], d);
            }
            n_acc += __shfl_xor_sync(0xffffffffu, n_acc, d);
        }
        if (lane < LANES) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { w_s1[warp][4 * lane + q] = S1[q]; w_s2[warp][4 * lane + q] = S2[q]; }
            if (lane == 0) w_n[warp] = n_acc;
        }
        __syncthreads();
        if (threadIdx.x < FC) {
            const int c = threadIdx.x;
            const int64_t col = (int64_t)s * FC + c;
            Sums tot;
            tot.n = 0.0; tot.s1 = 0.0; tot.s2 = 0.0;
            tot.piv = (double)reinterpret_cast<const float *>(&slab[0][0])[c];
            for (int w = 0; w < NWARPS; ++w) { tot.n += (double)w_n[w]; tot.s1 += w_s1[w][c]; tot.s2 += w_s2[w][c]; }
            double mean = 0.0, var = 1.0;
            if (col < p.cols) {
                const Moments m = to_moments(tot);
                mean = p.mean[col];
                var = p.var[col];
                chan_update_scalar(m.n, m.mean, m.m2, cnt, p.flags, mean, var);
                p.mean[col] = mean;
                p.var[col] = var;
                if (s == 0 && c == 0) p.batch_count_out[0] = m.n;
            }
            const double inv = 1.0 / sqrt(__dadd_rn(var, p.eps));
            const float mh = __double2float_rn(mean), ih = __double2float_rn(inv);
            reinterpret_cast<float *>(s_mu_hi)[c] = mh;
            reinterpret_cast<float *>(s_mu_lo)[c] = __double2float_rn(mean - (double)mh);
            reinterpret_cast<float *>(s_inv_hi)[c] = ih;
            reinterpret_cast<float *>(s_inv_lo)[c] = __double2float_rn(inv - (double)ih);
        }
        __syncthreads();

        // ---- normalise from shared memory, store straight to global; every slab element is owned by
        // one thread in all three phases, so as soon as a thread has drained a batch of its rows it
        // requests the same rows of the CTA's next slab into the same places: the next load overlaps
        // this slab's store phase
        const int s_next = s + (int)gridDim.x;
        const bool next_ok = s_next < p.n_slabs && (int64_t)s_next * FC + 4 * l < p.cols;
        const float *next_base = p.x + (int64_t)s_next * FC + 4 * l;
        {
            float4 mh, ml, ih, il;
            if (col_ok) { mh = s_mu_hi[l]; ml = s_mu_lo[l]; ih = s_inv_hi[l]; il = s_inv_lo[l]; }
            constexpr int U = 4;
            for (int r0 = g; r0 < nrows; r0 += GROUPS * U) {
                if (col_ok) {
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = r0 + u * GROUPS;
                        if (r < nrows) v[u] = slab[r][l];
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = r0 + u * GROUPS;
                        if (r < nrows) {
                            float4 o;
                            float d;
                            d = (v[u].x - mh.x) - ml.x; o.x = fminf(fmaxf(fmaf(d, ih.x, d * il.x), p.clip_lo), p.clip_hi);
                            d = (v[u].y - mh.y) - ml.y; o.y = fminf(fmaxf(fmaf(d, ih.y, d * il.y), p.clip_lo), p.clip_hi);
                            d = (v[u].z - mh.z) - ml.z; o.z = fminf(fmaxf(fmaf(d, ih.z, d * il.z), p.clip_lo), p.clip_hi);
                            d = (v[u].w - mh.w) - ml.w; o.w = fminf(fmaxf(fmaf(d, ih.w, d * il.w), p.clip_lo), p.clip_hi);
                            *reinterpret_cast<float4 *>(base + (int64_t)r * p.ld) = o;
                        }
                    }
                }
                if (next_ok) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = r0 + u * GROUPS;
                        if (r < nrows) {
                            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&slab[r][l]);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(next_base + (int64_t)r * p.ld) : "memory");
                        }
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // no barrier here: the next iteration's first barrier orders this slab's shared-memory reads
        // (column constants, warp sums) before their next writes; slab elements are thread-private
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace obs_lsu

static int g_obs_lsu_cols = 16;  // slab width of variant D (plb_set_tuning key 7: 8, 16 or 32)
void set_obs_lsu_cols(int v) { g_obs_lsu_cols = v; }

template <int FC>
static int obs_step_lsu_inst(const obs_lsu::Params &p0, cudaStream_t stream)
{
    using namespace obs_lsu;
    Params p = p0;
    const size_t slab_bytes = (size_t)p.rows * FC * sizeof(float);
    static thread_local size_t static_bytes = 0;
    if (static_bytes == 0) {
        cudaFuncAttributes attr;
        PLB_CUDA(cudaFuncGetAttributes(&attr, obs_step_lsu_kernel<FC>));
        static_bytes = attr.sharedSizeBytes;
    }
    const size_t avail = 227 * 1024 - static_bytes;
    if (slab_bytes > avail) return PLB_ERR_UNSUPPORTED;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    p.n_slabs = (int)ceil_div(p.cols, FC);
    int ctas_per_sm = (int)((228 * 1024) / (slab_bytes + static_bytes + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    const int grid = p.n_slabs < sms * ctas_per_sm ? p.n_slabs : sms * ctas_per_sm;
    static thread_local bool configured = false;
    if (!configured) {
        PLB_CUDA(cudaFuncSetAttribute(obs_step_lsu_kernel<FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
        configured = true;
    }
    obs_step_lsu_kernel<FC><<<grid, THREADS, slab_bytes, stream>>>(p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

int obs_step_lsu(float *x, int64_t ld, int64_t rows, int64_t cols, const uint8_t *row_valid,
                 double *mean, double *var, const double *count, double *batch_count_out,
                 double eps, int flags, float clip_lo, float clip_hi, cudaStream_t stream)
{
    using namespace obs_lsu;
    if (!(aligned(x, 16) && ld % 4 == 0 && cols % 4 == 0) || rows < 1 || rows > (int64_t)INT32_MAX / 2 ||
        cols > (int64_t)INT32_MAX / 2)
        return PLB_ERR_UNSUPPORTED;
    Params p = {};
    p.x = x; p.ld = ld; p.rows = rows; p.cols = cols;
    p.row_valid = row_valid;
    p.mean = mean; p.var = var; p.count = count;
    p.batch_count_out = batch_count_out;
    p.eps = eps; p.flags = flags;
    p.clip_lo = clip_lo; p.clip_hi = clip_hi;
    if (g_obs_lsu_cols == 8) return obs_step_lsu_inst<8>(p, stream);
    if (g_obs_lsu_cols == 32) return obs_step_lsu_inst<32>(p, stream);
    return obs_step_lsu_inst<16>(p, stream);
}

static int g_obs_variant = 0;  // 0 auto, 1 shared-memory slab (TMA), 2 two-pass L2 slab, 3 two-CTA cluster slab, 4 load/store-unit slab
int obs_variant() { return g_obs_variant; }
void set_obs_variant(int v) { g_obs_variant = v; }

int obs_step_two_pass_l2(float *x, int64_t ld, int64_t rows, int64_t cols, const uint8_t *row_valid,
                         double *mean, double *var, const double *count, double *batch_count_out,
                         double eps, int flags, float clip_lo, float clip_hi, cudaStream_t stream)
{
    using namespace obs_l2;
    if (!(aligned(x, 16) && ld % 4 == 0 && cols % 4 == 0) || rows < 1 || rows > (int64_t)INT32_MAX / 2)
        return PLB_ERR_UNSUPPORTED;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    Params p = {};
    p.x = x; p.ld = ld; p.rows = rows; p.cols = cols;
    p.row_valid = row_valid;
    p.mean = mean; p.var = var; p.count = count;
    p.batch_count_out = batch_count_out;
    p.eps = eps; p.flags = flags;
    p.clip_lo = clip_lo; p.clip_hi = clip_hi;
    p.n_slabs = (int)ceil_div(cols, WCOLS);
    // keep the slabs in flight (grid x rows x 256 B) well inside L2 so that pass 2 hits
    const int64_t slab_bytes = rows * WCOLS * 4;
    int64_t max_ctas = (int64_t)48 * 1024 * 1024 / (slab_bytes > 0 ? slab_bytes : 1);
    if (max_ctas < sms / 2) return PLB_ERR_UNSUPPORTED;  // slabs too tall for L2 reuse: three-kernel path
    int grid = (int)(max_ctas < 3 * sms ? max_ctas : 3 * sms);
    if (grid > p.n_slabs) grid = p.n_slabs;
    obs_step_l2_kernel<<<grid, THREADS, 0, stream>>>(p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

// Returns PLB_ERR_UNSUPPORTED when the shape does not fit the single-pass design.
int obs_step_single_pass(float *x, int64_t ld, int64_t rows, int64_t cols, const uint8_t *row_valid,
                         double *mean, double *var, const double *count, double *batch_count_out,
                         double eps, int flags, float clip_lo, float clip_hi, cudaStream_t stream)
{
    using namespace obs;
    if (!(aligned(x, 16) && ld % 4 == 0) || rows < 1 || rows > (int64_t)INT32_MAX / 2 || cols > (int64_t)INT32_MAX / 2)
        return PLB_ERR_UNSUPPORTED;
    const int n_boxes = (int)ceil_div(rows, BOX_ROWS);
    const size_t slab_bytes = (size_t)n_boxes * BOX_ROWS * FC * sizeof(float);
    static thread_local size_t static_bytes = 0;
    if (static_bytes == 0) {
        cudaFuncAttributes attr;
        PLB_CUDA(cudaFuncGetAttributes(&attr, obs_step_kernel));
        static_bytes = attr.sharedSizeBytes;
    }
    const size_t avail = 227 * 1024 - static_bytes;
    if (slab_bytes > avail) return PLB_ERR_UNSUPPORTED;
    CUtensorMap tm;
    int rc = make_tensor_map_2d(&tm, x, 4, rows, cols, ld, BOX_ROWS, FC);
    if (rc) return rc;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    Params p = {};
    p.rows = rows; p.cols = cols;
    p.row_valid = row_valid;
    p.mean = mean; p.var = var; p.count = count;
    p.batch_count_out = batch_count_out;
    p.eps = eps; p.flags = flags;
    p.clip_lo = clip_lo; p.clip_hi = clip_hi;
    p.n_slabs = (int)ceil_div(cols, FC);
    p.n_boxes = n_boxes;
    p.slab_floats = n_boxes * BOX_ROWS * FC;
    // Measured at C2 (B = 1024, 64 KB slabs): three independent single-buffer CTAs per SM (52 us) beat
    // one persistent CTA per SM with a 3-slab ring (66 us) -- with 64-byte rows the TMA unit is the
    // limiter (tools/microbench/tma_width.cu: 4.4 TB/s at 64 B, 5.3 at 128 B, 5.9 at 256 B per row
    // request) and independent CTAs keep more row requests in flight.  The ring is used only when
    // slabs are small enough that several fit next to several CTAs.
    int ctas_per_sm = (int)((228 * 1024) / (slab_bytes + static_bytes + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    int n_bufs = (int)(((228 * 1024) / ctas_per_sm - static_bytes - 1024) / slab_bytes);
    if (n_bufs > MAX_BUFS) n_bufs = MAX_BUFS;
    if (n_bufs < 1) n_bufs = 1;
    p.n_bufs = n_bufs;
    const int grid = p.n_slabs < sms * ctas_per_sm ? p.n_slabs : sms * ctas_per_sm;
    static thread_local bool configured = false;
    if (!configured) {
        PLB_CUDA(cudaFuncSetAttribute(obs_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
        configured = true;
    }
    obs_step_kernel<<<grid, THREADS, slab_bytes * n_bufs, stream>>>(tm, p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

}  // namespace plb
