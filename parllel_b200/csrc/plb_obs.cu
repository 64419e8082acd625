// plb_obs.cu -- single-pass NormalizeObservations step (parllel/transforms/norm_obs.py:57-76):
//   stats.update(obs[valid]); obs[:] = (obs - stats.mean) / sqrt(stats.var + eps)
// for one [B, F] float32 step with per-feature running statistics (float64).
//
// Each CTA owns a [B rows x 16 columns] slab: TMA brings the whole slab into shared memory once,
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

constexpr int FC = 16;          // columns per slab (64-byte row segments)
constexpr int THREADS = 256;
constexpr int GROUPS = THREADS / FC;  // 16 row groups; thread (g, c) takes rows g, g+16, g+32, ...
constexpr int BOX_ROWS = 256;   // rows per TMA box (box dimensions are limited to 256)

struct Params {
    int64_t rows, cols;
    const uint8_t *row_valid;
    double *mean, *var;
    const double *count;
    double *batch_count_out;  // workspace slot read by the count-bump kernel
    double eps;
    int flags;
    int n_slabs;
    int n_boxes;              // ceil(rows / BOX_ROWS)
};

__global__ void __launch_bounds__(THREADS)
obs_step_kernel(const __grid_constant__ CUtensorMap tm, const Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float (*slab)[FC] = reinterpret_cast<float (*)[FC]>(smem_raw);  // [n_boxes * BOX_ROWS][FC]
    __shared__ unsigned long long bar;
    __shared__ Sums part[GROUPS][FC];
    __shared__ double col_mean[FC], col_inv[FC];

    const int c = threadIdx.x % FC, g = threadIdx.x / FC;
    const int64_t rows = p.rows;
    const uint32_t box_bytes = BOX_ROWS * FC * sizeof(float);

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_proxy_async_smem();
        tma_prefetch_desc(&tm);
    }
    __syncthreads();

    uint32_t phase = 0;
    for (int s = blockIdx.x; s < p.n_slabs; s += gridDim.x) {
        const int c0 = s * FC;
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&bar, box_bytes * p.n_boxes);
            for (int b = 0; b < p.n_boxes; ++b) tma_load_2d(&slab[b * BOX_ROWS][0], &tm, c0, b * BOX_ROWS, &bar);
        }
        const bool col_ok = (int64_t)c0 + c < p.cols;
        const double old_mean = col_ok ? p.mean[c0 + c] : 0.0;   // overlaps the load
        const double old_var = col_ok ? p.var[c0 + c] : 1.0;
        const double cnt = p.count[0];
        mbar_wait(&bar, phase);
        phase ^= 1u;

        // ---- batch moments of this slab: rows g, g+16, ... of column c -----------------
        PivotAcc acc;
        acc.init();
        if (p.row_valid == nullptr) {
#pragma unroll 4
            for (int64_t r = g; r < rows; r += GROUPS) acc.add(slab[r][c]);
        } else {
            for (int64_t r = g; r < rows; r += GROUPS)
                if (p.row_valid[r] != 0) acc.add(slab[r][c]);
        }
        part[g][c] = acc.sums();
        __syncthreads();
        if (g == 0) {
            // division-free merge of the 16 row groups (fixed order), then the reference's update
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

        // ---- normalise the slab in place (float64 arithmetic, one rounding to float32) --
        const double mu = col_mean[c], inv = col_inv[c];
#pragma unroll 4
        for (int64_t r = g; r < rows; r += GROUPS)
            slab[r][c] = __double2float_rn(((double)slab[r][c] - mu) * inv);
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int b = 0; b < p.n_boxes; ++b) tma_store_2d(&tm, &slab[b * BOX_ROWS][0], c0, b * BOX_ROWS);
            tma_store_commit();
            tma_store_wait_read<0>();  // the slab buffer is reused by the next TMA load
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) tma_store_wait<0>();
}

}  // namespace obs

// Returns PLB_ERR_UNSUPPORTED when the shape does not fit the single-pass design.
int obs_step_single_pass(float *x, int64_t ld, int64_t rows, int64_t cols, const uint8_t *row_valid,
                         double *mean, double *var, const double *count, double *batch_count_out,
                         double eps, int flags, cudaStream_t stream)
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
    if (slab_bytes + static_bytes > 227 * 1024) return PLB_ERR_UNSUPPORTED;
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
    p.n_slabs = (int)ceil_div(cols, FC);
    p.n_boxes = n_boxes;
    int ctas_per_sm = (int)((228 * 1024) / (slab_bytes + static_bytes + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    const int grid = p.n_slabs < sms * ctas_per_sm ? p.n_slabs : sms * ctas_per_sm;
    static thread_local size_t configured = 0;
    if (slab_bytes > configured) {
        PLB_CUDA(cudaFuncSetAttribute(obs_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(227 * 1024 - static_bytes)));
        configured = 227 * 1024;
    }
    obs_step_kernel<<<grid, THREADS, slab_bytes, stream>>>(tm, p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

}  // namespace plb
