// plb_stats.cu -- RunningMeanStd on the device (parllel/transforms/running_mean_std.py):
//   batch moments  (np.mean / np.var, :52-53)  -> Welford/Chan parallel-merge reductions
//   update_from_moments (:7-32)               -> float64 Chan merge into the running state
#include "plb_common.cuh"
#include "plb_scan_shared.cuh"

namespace plb {

// ---------------------------------------------------------------------------------
// scalar moments over a [rows, cols] region (optionally masked per element)
// ---------------------------------------------------------------------------------
// HAS_SUB: moments of the float32 difference x - sub (explained_variance's residual)
template <bool VEC4, bool MASKED, bool HAS_SUB = false>
__global__ void __launch_bounds__(256)
batch_moments_kernel(const float *__restrict__ x, int64_t ld_x, int64_t rows, int64_t cols,
                     const uint8_t *__restrict__ valid, int64_t ld_v, MomentsTail tail,
                     const float *__restrict__ sub = nullptr, int64_t ld_s = 0)
{
    __shared__ Sums scratch[32];
    PivotAcc acc;
    acc.init();
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    if (VEC4) {
        const int64_t vpr = cols >> 2;  // float4 per row
        const int64_t total = rows * vpr;
        constexpr int U = 4;
        for (int64_t base = tid; base < total; base += nthreads * U) {
            float4 v[U];
            uchar4 m[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + (int64_t)u * nthreads;
                if (i < total) {
                    const int64_t r = (rows == 1) ? 0 : i / vpr;
                    const int64_t c = (i - r * vpr) << 2;
                    v[u] = *reinterpret_cast<const float4 *>(x + r * ld_x + c);
                    if (HAS_SUB) {
                        const float4 q = *reinterpret_cast<const float4 *>(sub + r * ld_s + c);
                        v[u].x = __fsub_rn(v[u].x, q.x); v[u].y = __fsub_rn(v[u].y, q.y);
                        v[u].z = __fsub_rn(v[u].z, q.z); v[u].w = __fsub_rn(v[u].w, q.w);
                    }
                    if (MASKED) m[u] = *reinterpret_cast<const uchar4 *>(valid + r * ld_v + c);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + (int64_t)u * nthreads;
                if (i < total) {
                    if (!MASKED || m[u].x) acc.add(v[u].x);
                    if (!MASKED || m[u].y) acc.add(v[u].y);
                    if (!MASKED || m[u].z) acc.add(v[u].z);
                    if (!MASKED || m[u].w) acc.add(v[u].w);
                }
            }
        }
    } else {
        const int64_t total = rows * cols;
        for (int64_t i = tid; i < total; i += nthreads) {
            const int64_t r = (rows == 1) ? 0 : i / cols;
            const int64_t c = i - r * cols;
            if (!MASKED || valid[r * ld_v + c])
                acc.add(HAS_SUB ? __fsub_rn(x[r * ld_x + c], sub[r * ld_s + c]) : x[r * ld_x + c]);
        }
    }
    moments_tail(acc.sums(), tail, scratch);
}

// ---------------------------------------------------------------------------------
// per-column moments: partial pass (row splits) + finalize
//   partial layout in workspace: [split][col] Moments
// ---------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(128)
column_moments_partial_kernel(const float *__restrict__ x, int64_t ld_x, int64_t rows, int64_t cols,
                              const uint8_t *__restrict__ row_valid, int64_t rows_per_split,
                              Moments *__restrict__ partials)
{
    const int64_t c0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (c0 >= cols) return;
    const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_hi = (r_lo + rows_per_split < rows) ? r_lo + rows_per_split : rows;
    PivotAcc acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k].init();
    constexpr int U = 4;
    for (int64_t r = r_lo; r < r_hi; r += U) {
        float v[U][VEC];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ok[u] = (r + u < r_hi) && (row_valid == nullptr || row_valid[r + u] != 0);
            if (ok[u]) {
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
            if (ok[u]) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k].add(v[u][k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) partials[(int64_t)blockIdx.y * cols + c0 + k] = acc[k].finish();
}

// merge the row splits of each column in split order; optionally fold the result into the
// running state (update_from_moments) in the same pass.
__global__ void __launch_bounds__(256)
column_moments_finalize_kernel(const Moments *__restrict__ partials, int64_t cols, int n_splits,
                               double *__restrict__ count_out, double *__restrict__ mean_out,
                               double *__restrict__ m2_out)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    Moments acc = partials[c];
    for (int s = 1; s < n_splits; ++s) acc = merge(acc, partials[(int64_t)s * cols + c]);
    mean_out[c] = acc.mean;
    m2_out[c] = acc.m2;
    if (c == 0) count_out[0] = acc.n;
}

// ---------------------------------------------------------------------------------
// valid mask of a recurrent batch (parllel/samplers/recurrent.py:102-103,141-142):
//   valid[0] = True; valid[t+1] = valid[t] & ~done[t]     (one thread per environment column)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
valid_prefix_kernel(const uint8_t *__restrict__ done, int64_t ld_d, int64_t T, int64_t B,
                    uint8_t *__restrict__ valid, int64_t ld_v)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    uint8_t ok = 1;
    constexpr int U = 8;
    for (int64_t t0 = 0; t0 < T; t0 += U) {
        uint8_t d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) d[u] = (t0 + u < T) ? done[(t0 + u) * ld_d + b] : (uint8_t)0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t0 + u < T) {
                valid[(t0 + u) * ld_v + b] = ok;
                ok = (ok && !d[u]) ? 1 : 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// update_from_moments: single CTA so that the shared `count` is read by everyone
// before it is overwritten.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
update_from_moments_kernel(const double *__restrict__ batch_count, const double *__restrict__ batch_mean,
                           const double *__restrict__ batch_m2, double *mean, double *var, double *count,
                           int64_t n, int flags)
{
    const double b_n = batch_count[0];
    const double cnt = count[0];
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double m = mean[i], v = var[i];
        chan_update_scalar(b_n, batch_mean[i], batch_m2[i], cnt, flags, m, v);
        mean[i] = m;
        var[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) count[0] = b_n + cnt;
}

// merge per-rank moments in rank order (identical result on every rank)
//   layout 0: parts[n_parts][n_cols][3] triples     -> out[n_cols][3]
//   layout 1: parts[n_parts][1 + 2*n_cols] (count, mean[n_cols], m2[n_cols]) -> out[1 + 2*n_cols]
__global__ void __launch_bounds__(256)
merge_moments_kernel(const double *__restrict__ parts, int n_parts, int64_t n_cols, int layout,
                     double *__restrict__ out)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int64_t part_stride = layout == 0 ? 3 * n_cols : 1 + 2 * n_cols;
    Moments acc;
    acc.n = 0.0; acc.mean = 0.0; acc.m2 = 0.0;
    for (int p = 0; p < n_parts; ++p) {
        const double *q = parts + (int64_t)p * part_stride;
        Moments m;
        if (layout == 0) {
            m.n = q[3 * c]; m.mean = q[3 * c + 1]; m.m2 = q[3 * c + 2];
        } else {
            m.n = q[0]; m.mean = q[1 + c]; m.m2 = q[1 + n_cols + c];
        }
        acc = merge(acc, m);
    }
    if (layout == 0) {
        out[3 * c] = acc.n; out[3 * c + 1] = acc.mean; out[3 * c + 2] = acc.m2;
    } else {
        if (c == 0) out[0] = acc.n;
        out[1 + c] = acc.mean;
        out[1 + n_cols + c] = acc.m2;
    }
}

__global__ void __launch_bounds__(256)
finalize_partials_kernel(const Sums *partials, const unsigned *n_partials, double *moments_out)
{
    __shared__ Sums scratch[32];
    const Moments m = reduce_published_partials(partials, *n_partials, scratch);
    if (threadIdx.x == 0) {
        moments_out[0] = m.n; moments_out[1] = m.mean; moments_out[2] = m.m2;
    }
}

int finalize_published_partials(void *tail_workspace, double *moments_out, cudaStream_t stream)
{
    const MomentsTail t = make_tail(moments_out, tail_workspace);
    finalize_partials_kernel<<<1, 256, 0, stream>>>(t.partials, t.ticket + 1, moments_out);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

int column_moments_splits(int64_t rows, int64_t cols)
{
    // enough CTAs to fill the machine: column blocks x row splits ~ 8 waves of 148
    const int64_t col_blocks = ceil_div(ceil_div(cols, 4), 128);
    int64_t want = ceil_div((int64_t)148 * 8, col_blocks);
    if (want < 1) want = 1;
    int64_t max_splits = ceil_div(rows, 8);  // at least 8 rows per split
    if (max_splits < 1) max_splits = 1;
    if (want > max_splits) want = max_splits;
    if (want > 256) want = 256;
    return (int)want;
}

int column_moments_partial(const float *x, int64_t ld_x, int64_t rows, int64_t cols, const uint8_t *row_valid,
                           Moments *partials, int splits, cudaStream_t stream)
{
    const int64_t rows_per_split = rows > 0 ? ceil_div(rows, splits) : 1;
    const bool vec = (cols % 4 == 0) && (ld_x % 4 == 0) && aligned(x, 16);
    if (vec) {
        dim3 grid((unsigned)ceil_div(cols / 4, 128), (unsigned)splits);
        column_moments_partial_kernel<4><<<grid, 128, 0, stream>>>(x, ld_x, rows, cols, row_valid, rows_per_split, partials);
    } else {
        dim3 grid((unsigned)ceil_div(cols, 128), (unsigned)splits);
        column_moments_partial_kernel<1><<<grid, 128, 0, stream>>>(x, ld_x, rows, cols, row_valid, rows_per_split, partials);
    }
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

int column_moments(const float *x, int64_t ld_x, int64_t rows, int64_t cols, const uint8_t *row_valid,
                   double *count_out, double *mean_out, double *m2_out,
                   void *workspace, size_t workspace_bytes, cudaStream_t stream)
{
    PLB_REQUIRE(rows >= 0 && cols >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (cols == 0) return PLB_OK;
    PLB_REQUIRE(x && count_out && mean_out && m2_out, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(ld_x >= cols, PLB_ERR_INVALID_ARGUMENT, "row stride smaller than row length");
    const int splits = column_moments_splits(rows, cols);
    const size_t need = sizeof(Moments) * (size_t)splits * (size_t)cols;
    PLB_REQUIRE(workspace != nullptr && workspace_bytes >= need, PLB_ERR_WORKSPACE,
                "column moments need %zu workspace bytes, got %zu", need, workspace_bytes);
    Moments *partials = reinterpret_cast<Moments *>(workspace);
    const int rc = column_moments_partial(x, ld_x, rows, cols, row_valid, partials, splits, stream);
    if (rc) return rc;
    column_moments_finalize_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, stream>>>(
        partials, cols, splits, count_out, mean_out, m2_out);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

}  // namespace plb

extern "C" {

size_t plb_batch_moments_workspace_bytes(void) { return plb::MOMENTS_WS_BYTES; }

static int batch_moments_impl(const float *x, int64_t ld_x, const float *sub, int64_t ld_sub, int64_t rows, int64_t cols,
                              const uint8_t *valid, int64_t ld_valid, double *moments_out,
                              void *workspace, size_t workspace_bytes, cudaStream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(rows >= 0 && cols >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    PLB_REQUIRE(moments_out != nullptr, PLB_ERR_INVALID_ARGUMENT, "null output");
    PLB_REQUIRE(rows * cols == 0 || x != nullptr, PLB_ERR_INVALID_ARGUMENT, "null input");
    PLB_REQUIRE(workspace != nullptr && workspace_bytes >= MOMENTS_WS_BYTES, PLB_ERR_WORKSPACE,
                "batch moments need %zu workspace bytes, got %zu", MOMENTS_WS_BYTES, workspace_bytes);
    PLB_REQUIRE(rows * cols == 0 || (ld_x >= cols && (valid == nullptr || ld_valid >= cols) &&
                                     (sub == nullptr || ld_sub >= cols)),
                PLB_ERR_INVALID_ARGUMENT, "row stride smaller than row length");
    // flatten dense regions so the kernel skips the row/column split
    if (ld_x == cols && (valid == nullptr || ld_valid == cols) && (sub == nullptr || ld_sub == cols)) {
        cols = rows * cols;
        rows = cols > 0 ? 1 : 0;
        ld_x = cols;
        ld_valid = cols;
        ld_sub = cols;
    }
    const int64_t total = rows * cols;
    const bool vec = (cols % 4 == 0) && (ld_x % 4 == 0) && aligned(x, 16) &&
                     (valid == nullptr || ((ld_valid % 4 == 0) && aligned(valid, 4))) &&
                     (sub == nullptr || ((ld_sub % 4 == 0) && aligned(sub, 16)));
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    int64_t blocks = ceil_div(vec ? ceil_div(total, 4) : total, 256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    if (blocks > MOMENTS_MAX_PARTIALS) blocks = MOMENTS_MAX_PARTIALS;
    const MomentsTail tail = make_tail(moments_out, workspace);
#define PLB_BM_LAUNCH(V, M, S) \
    batch_moments_kernel<V, M, S><<<(unsigned)blocks, 256, 0, stream>>>(x, ld_x, rows, cols, valid, ld_valid, tail, sub, ld_sub)
    if (sub != nullptr) {
        if (vec) { if (valid) PLB_BM_LAUNCH(true, true, true); else PLB_BM_LAUNCH(true, false, true); }
        else { if (valid) PLB_BM_LAUNCH(false, true, true); else PLB_BM_LAUNCH(false, false, true); }
    } else {
        if (vec) { if (valid) PLB_BM_LAUNCH(true, true, false); else PLB_BM_LAUNCH(true, false, false); }
        else { if (valid) PLB_BM_LAUNCH(false, true, false); else PLB_BM_LAUNCH(false, false, false); }
    }
#undef PLB_BM_LAUNCH
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

int plb_batch_moments_f32(const float *x, int64_t ld_x, int64_t rows, int64_t cols,
                          const uint8_t *valid, int64_t ld_valid, double *moments_out,
                          void *workspace, size_t workspace_bytes, plb_stream_t stream_)
{
    return batch_moments_impl(x, ld_x, nullptr, 0, rows, cols, valid, ld_valid, moments_out, workspace, workspace_bytes,
                              (cudaStream_t)stream_);
}

int plb_residual_moments_f32(const float *x, int64_t ld_x, const float *sub, int64_t ld_sub, int64_t rows, int64_t cols,
                             const uint8_t *valid, int64_t ld_valid, double *moments_out,
                             void *workspace, size_t workspace_bytes, plb_stream_t stream_)
{
    PLB_REQUIRE(rows * cols == 0 || sub != nullptr, PLB_ERR_INVALID_ARGUMENT, "null subtrahend");
    return batch_moments_impl(x, ld_x, sub, ld_sub, rows, cols, valid, ld_valid, moments_out, workspace, workspace_bytes,
                              (cudaStream_t)stream_);
}

int plb_valid_prefix_u8(const uint8_t *done, int64_t ld_done, int64_t T, int64_t B, uint8_t *valid_out, int64_t ld_valid,
                        plb_stream_t stream_)
{
    using namespace plb;
    PLB_REQUIRE(T >= 0 && B >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (T == 0 || B == 0) return PLB_OK;
    PLB_REQUIRE(done != nullptr && valid_out != nullptr, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(ld_done >= B && ld_valid >= B, PLB_ERR_INVALID_ARGUMENT, "row stride smaller than row length");
    valid_prefix_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, (cudaStream_t)stream_>>>(done, ld_done, T, B, valid_out, ld_valid);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

size_t plb_column_moments_workspace_bytes(int64_t rows, int64_t cols)
{
    if (rows < 0 || cols <= 0) return 0;
    return sizeof(plb::Moments) * (size_t)plb::column_moments_splits(rows, cols) * (size_t)cols;
}

int plb_column_moments_f32(const float *x, int64_t ld_x, int64_t rows, int64_t cols, const uint8_t *row_valid,
                           double *count_out, double *mean_out, double *m2_out,
                           void *workspace, size_t workspace_bytes, plb_stream_t stream)
{
    return plb::column_moments(x, ld_x, rows, cols, row_valid, count_out, mean_out, m2_out,
                               workspace, workspace_bytes, (cudaStream_t)stream);
}

int plb_merge_moments_f64(const double *parts, int32_t n_parts, int64_t n_cols, int layout, double *out,
                          plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(n_parts >= 1 && n_cols >= 1 && (layout == 0 || layout == 1), PLB_ERR_INVALID_ARGUMENT, "bad extent");
    PLB_REQUIRE(parts && out, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    merge_moments_kernel<<<(unsigned)ceil_div(n_cols, 256), 256, 0, (cudaStream_t)stream>>>(parts, n_parts, n_cols,
                                                                                             layout, out);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

int plb_update_from_moments_f64(const double *batch_count, const double *batch_mean, const double *batch_m2,
                                double *mean, double *var, double *count, int64_t n, int flags,
                                plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(n >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    PLB_REQUIRE(batch_count && batch_mean && batch_m2 && mean && var && count, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    const int threads = n >= 1024 ? 1024 : (n > 32 ? (int)ceil_div(n, 32) * 32 : 32);
    update_from_moments_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(batch_count, batch_mean, batch_m2,
                                                                         mean, var, count, n, flags);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

}  // extern "C"
