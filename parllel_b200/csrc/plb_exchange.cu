// plb_exchange.cu -- multi-GPU statistics exchange over NVLink peer memory.
//
// The only exchange step of the path is an all-gather of tiny per-rank (count, mean, M2) moments
// followed by a rank-ordered merge.  Through NCCL that costs ~25 us of
// collective latency per exchange even inside a CUDA graph -- more than the 22 us single-GPU step.
// This kernel does it in one launch over peer-mapped ("symmetric") buffers: every rank stores its
// moments directly into every peer's buffer (NVLink/NVSwitch P2P stores), publishes a flag with
// release semantics at system scope, spins on its own flags (acquire), and merges the gathered
// moments in rank order -- bit-identical on every rank.
//
// Buffer layout on every rank (same on all ranks, `exchange_bytes(world, n_vals)` long):
//     uint32  epoch                         this rank's exchange counter (device-resident so that a
//                                           CUDA graph can replay the kernel)
//     uint32  flags[2][world]   (+ padding) flags[p][r] == e  <=>  rank r's epoch-e data has landed
//     double  data[2][world][n_vals]        p = epoch parity (double buffer)
// One process per GPU; never run two ranks of one exchange on the same GPU (they would spin on each
// other without any guarantee of co-scheduling).
#include "plb_common.cuh"
#include "plb_scan_shared.cuh"

namespace plb {

constexpr int EX_MAX_WORLD = 64;
constexpr size_t EX_HEADER_BYTES = 1024;  // epoch + flags, keeps data 1 KB aligned

struct ExchangeParams {
    const double *local;      // [n_vals] this rank's moments
    unsigned char *const *peers;  // device array [world]: every rank's exchange buffer, peer-mapped
    int rank, world;
    int64_t n_vals;
    int layout;               // merge layout (see plb_merge_moments_f64)
    int64_t n_cols;
    double *merged_out;
};

__global__ void __launch_bounds__(256)
exchange_moments_kernel(const ExchangeParams p)
{
    __shared__ uint32_t s_epoch;
    __shared__ int s_timed_out;
    unsigned char *mine = p.peers[p.rank];
    if (threadIdx.x == 0) {
        s_timed_out = 0;
        uint32_t *ep = reinterpret_cast<uint32_t *>(mine);
        s_epoch = *ep + 1u;
        *ep = s_epoch;
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    const int par = (int)(epoch & 1u);
    const size_t data_off = EX_HEADER_BYTES + (size_t)par * p.world * p.n_vals * sizeof(double);

    // 1. push my moments into slot [par][rank] of every rank's buffer (including my own)
    for (int r = 0; r < p.world; ++r) {
        double *dst = reinterpret_cast<double *>(p.peers[r] + data_off) + (size_t)p.rank * p.n_vals;
        for (int64_t i = threadIdx.x; i < p.n_vals; i += blockDim.x) dst[i] = p.local[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: flags[par][rank] = epoch on every rank; 3. wait for every rank's flag on my buffer
    if (threadIdx.x < p.world) {
        const int r = threadIdx.x;
        uint32_t *remote_flag = reinterpret_cast<uint32_t *>(p.peers[r] + 64) + par * p.world + p.rank;
        st_release_sys(remote_flag, epoch);
        const uint32_t *my_flag = reinterpret_cast<const uint32_t *>(mine + 64) + par * p.world + r;
        const long long t0 = clock64();
        while (ld_acquire_sys(my_flag) != epoch) {
            // ~10 s: a peer died.  Do not hang the GPU, do not kill the context either: the merged moments
            // come out as NaN (as in xch_wait_merge3), which the caller sees in its statistics
            if (clock64() - t0 > 20000000000LL) { s_timed_out = 1; break; }
        }
    }
    __syncthreads();
    const bool timed_out = s_timed_out != 0;
    // 4. rank-ordered merge of the gathered moments (same order on every rank)
    const double *parts = reinterpret_cast<const double *>(mine + data_off);
    const int64_t part_stride = p.n_vals;
    for (int64_t c = threadIdx.x; c < p.n_cols; c += blockDim.x) {
        Moments acc;
        acc.n = 0.0; acc.mean = 0.0; acc.m2 = 0.0;
        for (int r = 0; r < p.world; ++r) {
            const double *q = parts + (size_t)r * part_stride;
            Moments m;
            if (p.layout == 0) {
                m.n = q[3 * c]; m.mean = q[3 * c + 1]; m.m2 = q[3 * c + 2];
            } else {
                m.n = q[0]; m.mean = q[1 + c]; m.m2 = q[1 + p.n_cols + c];
            }
            acc = merge(acc, m);
        }
        if (timed_out) acc.mean = acc.m2 = __longlong_as_double(0x7ff8000000000000LL);
        if (p.layout == 0) {
            p.merged_out[3 * c] = acc.n; p.merged_out[3 * c + 1] = acc.mean; p.merged_out[3 * c + 2] = acc.m2;
        } else {
            if (c == 0) p.merged_out[0] = acc.n;
            p.merged_out[1 + c] = acc.mean;
            p.merged_out[1 + p.n_cols + c] = acc.m2;
        }
    }
}

// wait + merge (+ optional running-state update) for moments pushed by a producer kernel's tail
__global__ void __launch_bounds__(32)
exchange_finish_kernel(const XchRef xch, double *moments_out, double *upd_mean, double *upd_var, double *upd_count,
                       int upd_flags)
{
    const Moments m = xch_wait_merge3(xch);
    if (threadIdx.x == 0) {
        moments_out[0] = m.n; moments_out[1] = m.mean; moments_out[2] = m.m2;
        if (upd_mean != nullptr) {
            double mean = *upd_mean, var = *upd_var;
            const double cnt = *upd_count;
            chan_update_scalar(m.n, m.mean, m.m2, cnt, upd_flags, mean, var);
            *upd_mean = mean; *upd_var = var; *upd_count = m.n + cnt;
        }
    }
}

int exchange_finish(const XchRef &xch, double *moments_out, double *upd_mean, double *upd_var, double *upd_count,
                    int upd_flags, cudaStream_t stream)
{
    exchange_finish_kernel<<<1, 32, 0, stream>>>(xch, moments_out, upd_mean, upd_var, upd_count, upd_flags);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

}  // namespace plb

extern "C" {

size_t plb_exchange_buffer_bytes(int32_t world, int64_t n_vals)
{
    if (world < 1 || n_vals < 1) return 0;
    const size_t generic = plb::EX_HEADER_BYTES + 2 * (size_t)world * (size_t)n_vals * sizeof(double);
    // buffers for scalar statistics also carry the low-latency area of the in-kernel exchange
    if (n_vals == 3) return plb::xch_total_bytes3(world);
    return generic;
}

int plb_exchange_moments_p2p(const double *local, int64_t n_vals, void *const *peer_buffers, int32_t rank,
                             int32_t world, int layout, int64_t n_cols, double *merged_out, plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(local && peer_buffers && merged_out, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(world >= 1 && world <= EX_MAX_WORLD && rank >= 0 && rank < world, PLB_ERR_INVALID_ARGUMENT, "bad rank/world");
    PLB_REQUIRE(64 + 2 * (size_t)world * 4 <= EX_HEADER_BYTES, PLB_ERR_INVALID_ARGUMENT, "world too large");
    PLB_REQUIRE(n_vals >= 1 && n_cols >= 1 && (layout == 0 ? n_vals == 3 * n_cols : n_vals == 1 + 2 * n_cols),
                PLB_ERR_INVALID_ARGUMENT, "n_vals does not match the layout");
    ExchangeParams p;
    p.local = local;
    p.peers = reinterpret_cast<unsigned char *const *>(peer_buffers);
    p.rank = rank; p.world = world;
    p.n_vals = n_vals; p.layout = layout; p.n_cols = n_cols;
    p.merged_out = merged_out;
    exchange_moments_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

}  // extern "C"
