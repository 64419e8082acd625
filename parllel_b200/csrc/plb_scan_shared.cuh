// plb_scan_shared.cuh -- PTX wrappers (mbarrier, TMA) and the cross-CTA moments tail
// shared by the scan, statistics and observation kernels.
#pragma once

#include <cuda.h>

#include "plb_common.cuh"

namespace plb {

// ---------------------------------------------------------------------------------
// Workspace layout for a grid-wide (count, mean, M2) reduction:
//   [0, 64)            ticket counter (uint32) -- zero before first use, reset by the kernel
//   [64, 64 + 24*MAX)  per-CTA partial moments
// ---------------------------------------------------------------------------------
constexpr int MOMENTS_MAX_PARTIALS = 4096;
constexpr size_t MOMENTS_WS_BYTES = 64 + sizeof(Sums) * MOMENTS_MAX_PARTIALS;

// Reference to the peer-mapped exchange buffers of plb_exchange.cu (n_vals = 3 layout):
// peers == nullptr switches the in-kernel exchange off.
struct XchRef {
    unsigned char *const *peers;  // device array [world] of every rank's buffer, peer-mapped
    int rank, world;
};
constexpr size_t XCH_HEADER_BYTES = 1024;  // uint32 epoch @0, uint32 flags[2][world] @64, data @1024

// Grid-wide + cross-GPU statistics exchange of the resident advantage kernel (one protocol for both):
// every CTA of every rank stores its partial (count, mean, M2) as an epoch-tagged record straight into
// EVERY rank's buffer (its own included; NVLink P2P stores for the peers), then polls its own rank's
// buffer until all world x slots records of the launch epoch have landed and merges them in a fixed
// order -- so the grid barrier of one GPU and the all-gather across GPUs are the same step, with no
// atomics, no last-CTA reduction and no second kernel.  Buffer layout (plb_grid_exchange_bytes):
// (every word that is polled or hit by atomics has a 128-byte line of its own: hundreds of CTAs on one line
//  serialise in its L2 bank)
//     line 0:  uint32 epoch @0 | uint32 status @8 (0 ok, 1 timed out)
//     line 1:  uint32 exit_ticket @128
//     line 2:  uint32 arrived[2] @256     (see below)
//     uint32 arrived[2] @256 (by epoch parity): slots filled so far, a HINT for when to read -- the
//            records validate themselves through their epoch tags, so no fence orders it after them
//     records @512: [2 (epoch parity)][world][slots] x 64 B = 4 x {lo, epoch, hi, epoch} 16-byte units
//            carrying the division-free Sums (n, s1, s2, pivot) of one CTA
struct GridXch {
    unsigned char *const *peers;  // device array [world]: every rank's buffer (world == 1: just ours)
    unsigned char *self;          // this rank's buffer (= peers[rank], passed by value: no dependent table load)
    int rank, world;
    int slots;                    // records per rank and epoch (>= the largest grid of any rank)
};
constexpr size_t GX_HEADER_BYTES = 512;
constexpr int GX_TICKET_WORD = 32, GX_ARRIVED_WORD = 64;  // header offsets in 32-bit words
constexpr size_t GX_RECORD_BYTES = 64;
constexpr int GX_RECORD_VALS = 4;
constexpr int GX_MAX_WORLD = 32;
constexpr int GX_MAX_SLOTS = 1024;
__host__ __device__ __forceinline__ size_t gx_total_bytes(int world, int slots)
{
    return GX_HEADER_BYTES + 2 * (size_t)world * (size_t)slots * GX_RECORD_BYTES;
}

struct MomentsTail {
    double *moments_out;  // [3] = (count, mean, M2); nullptr disables the tail
    Sums *partials;
    unsigned *ticket;     // workspace word 0: last-CTA ticket; word 1: number of partials written
    int partials_only;    // 1: every CTA only publishes its partial Sums (and the CTA count); the
                          //    consumer kernel reduces them itself (no serial last-CTA epilogue)
    // optional: the CTA that finishes the reduction also folds the batch moments into a scalar
    // running state (update_from_moments, running_mean_std.py:7-32) -- saves a kernel launch
    double *upd_mean, *upd_var, *upd_count;
    int upd_flags;
    // optional (multi-GPU): the finishing CTA pushes the final moments straight into every peer's
    // exchange buffer (NVLink P2P stores + release flags); the consumer kernel waits and merges
    XchRef xch;
    unsigned long long *dbg;  // experiments: %globaltimer stamps of the sharded step (plb_debug_set_buffer), or NULL
};

static inline MomentsTail make_tail(double *moments_out, void *workspace)
{
    MomentsTail t;
    t.moments_out = moments_out;
    t.ticket = reinterpret_cast<unsigned *>(workspace);
    t.partials = reinterpret_cast<Sums *>(reinterpret_cast<unsigned char *>(workspace) + 64);
    t.partials_only = 0;
    t.upd_mean = t.upd_var = t.upd_count = nullptr;
    t.upd_flags = 0;
    t.xch.peers = nullptr; t.xch.rank = 0; t.xch.world = 1;
    t.dbg = nullptr;
    return t;
}

// Optional work fused into the chunked scans (host-side description).
struct ScanFusion {
    int has_pre;               // apply reward scale/clip on the load path
    const double *reward_var;  // device running variance (float64) or nullptr
    double reward_eps;
    double clip_lo, clip_hi;   // NaN = no bound
    float *reward_out;         // write the transformed reward back (may alias the input)
    int64_t ld_reward_out;
    const uint8_t *valid;      // mask for the fused moments (nullptr = all)
    int64_t ld_valid;
    MomentsTail tail;          // fused moments of advantage / past_return
    float *carry_ret_out;      // forward scan: receives out[T-1] per column (may alias prev_ret), or nullptr
    uint8_t *carry_done_out;   // forward scan: receives done[T-1] per column (may alias prev_done), or nullptr
    // reverse scans, optional: NormalizeAdvantage fused into the same launch (resident kernel): the
    // advantage stays on chip (TMEM) until the statistics of the whole (multi-GPU) batch are known
    int resident;              // 1: requested; scan_reverse returns PLB_ERR_UNSUPPORTED if the shape cannot
    GridXch gx;
    double adv_eps;
    double *adv_moments_out;   // [3] merged (count, mean, M2)
};
bool scan_chunked_supported(const float *reward, int64_t ld_r, const float *value, int64_t ld_v,
                            const uint8_t *done, int64_t ld_d, int64_t T, int64_t B, int64_t inner);
bool resident_wanted(int world);  // policy (plb_set_tuning key 12)
int sharded_route();              // plb_set_tuning key 17
unsigned long long *debug_buffer();  // plb_debug_set_buffer
bool resident_supported(const float *reward, int64_t ld_r, const float *value, int64_t ld_v, const uint8_t *done,
                        int64_t ld_d, const float *adv, int64_t ld_a, const float *ret, int64_t ld_t,
                        int64_t T, int64_t B);

int scan_reverse(int mode, const float *reward, int64_t ld_r, const float *value, int64_t ld_v,
                 const uint8_t *done, int64_t ld_d, const float *boot,
                 float *adv, int64_t ld_a, float *ret, int64_t ld_t,
                 int64_t T, int64_t B, int64_t inner, double gamma, double lambda, int algo,
                 const ScanFusion *fusion, cudaStream_t stream);
int scan_forward(const float *reward, int64_t ld_r, const uint8_t *done, int64_t ld_d,
                 const float *prev_ret, const uint8_t *prev_done, float *out, int64_t ld_o,
                 int64_t T, int64_t B, double gamma, int algo, const ScanFusion *fusion, cudaStream_t stream);
int column_moments_splits(int64_t rows, int64_t cols);
int column_moments_partial(const float *x, int64_t ld_x, int64_t rows, int64_t cols, const uint8_t *row_valid,
                           Moments *partials, int splits, cudaStream_t stream);
int column_moments(const float *x, int64_t ld_x, int64_t rows, int64_t cols, const uint8_t *row_valid,
                   double *count_out, double *mean_out, double *m2_out,
                   void *workspace, size_t workspace_bytes, cudaStream_t stream);
int obs_variant();
void set_obs_variant(int v);
void set_normadv_ctas_per_sm(int v);
void set_normadv_xch_threads(int v);
void set_normadv_reduce_mode(int v);
void set_obs_lsu_cols(int v);
void set_rng_cluster(int v);
void set_gather_mode(int v);
int obs_step_two_pass_l2(float *x, int64_t ld, int64_t rows, int64_t cols, const uint8_t *row_valid,
                         double *mean, double *var, const double *count, double *batch_count_out,
                         double eps, int flags, float clip_lo, float clip_hi, cudaStream_t stream);
int obs_step_lsu(float *x, int64_t ld, int64_t rows, int64_t cols, const uint8_t *row_valid,
                 double *mean, double *var, const double *count, double *batch_count_out,
                 double eps, int flags, float clip_lo, float clip_hi, cudaStream_t stream);
int obs_step_single_pass(float *x, int64_t ld, int64_t rows, int64_t cols, const uint8_t *row_valid,
                         double *mean, double *var, const double *count, double *batch_count_out,
                         double eps, int flags, float clip_lo, float clip_hi, cudaStream_t stream);
int finalize_published_partials(void *tail_workspace, double *moments_out, cudaStream_t stream);
int exchange_finish(const XchRef &xch, double *moments_out, double *upd_mean, double *upd_var, double *upd_count,
                    int upd_flags, cudaStream_t stream);
int normalize_advantage_after_exchange(float *x, int64_t ld_x, int64_t rows, int64_t cols, const XchRef &xch, double eps,
                                       double *moments_out, cudaStream_t stream);
int normalize_advantage_from_partials(float *x, int64_t ld_x, int64_t rows, int64_t cols, void *tail_workspace,
                                      double eps, double *moments_out, cudaStream_t stream);
int make_tensor_map_2d(CUtensorMap *tm, const void *base, int elem_bytes, int64_t rows, int64_t cols,
                       int64_t ld_elems, int box_rows, int box_cols);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------
// mbarrier + TMA (inline PTX; SASS: SYNCS.*, UTMALDG / UBLKCP)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    unsigned spins = 0;
    const long long t_start = clock64();
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        // a lost TMA would otherwise hang the GPU: give up after ~10 s of polling (checked every 2^20 polls)
        if ((++spins & ((1u << 20) - 1u)) == 0u && clock64() - t_start > 20000000000LL) __trap();
    }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// 2-D tiled load: box at (c0 = column, c1 = row) -> shared, completes on `bar`.
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1,
                                            unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 2-D tiled store shared -> global (bulk async group)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, const void *src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// at most `n` of this thread's bulk groups may still be reading shared memory (n clamped to [0, 8])
__device__ __forceinline__ void tma_store_wait_read_dyn(int n)
{
    switch (n) {
        case 0: tma_store_wait_read<0>(); break;
        case 1: tma_store_wait_read<1>(); break;
        case 2: tma_store_wait_read<2>(); break;
        case 3: tma_store_wait_read<3>(); break;
        case 4: tma_store_wait_read<4>(); break;
        case 5: tma_store_wait_read<5>(); break;
        case 6: tma_store_wait_read<6>(); break;
        case 7: tma_store_wait_read<7>(); break;
        default: tma_store_wait_read<8>(); break;
    }
}
template <int N>
__device__ __forceinline__ void tma_store_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// 1-D bulk copies (no tensor map): global -> shared on an mbarrier, shared -> global in a bulk group
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void bulk_store_1d(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(reinterpret_cast<uint64_t>(dst)), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}

// ---------------------------------------------------------------------------------
// In-kernel statistics exchange over peer memory (see plb_exchange.cu for the protocol)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// In-kernel exchange of one record of partial sums per rank, low-latency ("LL") protocol: every 32-bit half of
// the payload travels in its own naturally aligned 8-byte store {half, epoch}, so the epoch tag doubles as the
// arrival flag -- no fence between data and flag, one NVLink store latency end to end.  The payload is the
// division-free Sums (n, s1, s2, pivot), so the consumer adds the ranks up with rebase + add and divides once.
// LL area of a rank's buffer (after the generic area of plb_exchange.cu):
//     uint2 ll[2 (parity)][world][8]     (8 = 4 doubles x 2 halves)
__host__ __device__ __forceinline__ size_t xch_ll_offset(int world)
{
    const size_t generic = XCH_HEADER_BYTES + 2 * (size_t)world * 3 * sizeof(double);
    return (generic + 127) / 128 * 128;
}
__host__ __device__ __forceinline__ size_t xch_total_bytes3(int world)
{
    return xch_ll_offset(world) + 2 * (size_t)world * 8 * 8;
}

// What xch_push3 needs from memory, fetched by the pushing warp long before the push (at kernel start: the epoch word
// only changes in the push itself): the pointer table entry of lane's peer, the own buffer, the current epoch.
struct XchPre {
    unsigned char *mine, *peer;
    uint32_t epoch;
    bool valid;
};
__device__ __forceinline__ XchPre xch_preload(const XchRef &x)
{
    XchPre pre;
    const int lane = threadIdx.x & 31;
    pre.mine = x.peers[x.rank];
    pre.peer = lane < x.world ? x.peers[lane] : nullptr;
    pre.epoch = *reinterpret_cast<const volatile uint32_t *>(pre.mine);
    pre.valid = true;
    return pre;
}

// Push this rank's sums to every rank.  Called by ONE full warp (world <= 32); lane 0 supplies them.
__device__ __forceinline__ XchPre xch_no_preload()
{
    XchPre pre;
    pre.mine = nullptr; pre.peer = nullptr; pre.epoch = 0u; pre.valid = false;
    return pre;
}
__device__ __forceinline__ void xch_push3(const XchRef &x, Sums mine_sums, const XchPre pre = xch_no_preload())
{
    const int lane = threadIdx.x & 31;
    const bool have = pre.valid;
    unsigned char *mine = have ? pre.mine : x.peers[x.rank];
    unsigned char *peer = have ? pre.peer : (lane < x.world ? x.peers[lane] : nullptr);  // in flight with the epoch round trip
    uint32_t epoch = 0;
    if (lane == 0) {
        uint32_t *ep = reinterpret_cast<uint32_t *>(mine);
        epoch = (have ? pre.epoch : *ep) + 1u;
        *ep = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    double vals[4] = {mine_sums.n, mine_sums.s1, mine_sums.s2, mine_sums.piv};
#pragma unroll
    for (int i = 0; i < 4; ++i) vals[i] = __shfl_sync(0xffffffffu, vals[i], 0);
    const int par = (int)(epoch & 1u);
    if (lane < x.world) {
        uint2 *dst = reinterpret_cast<uint2 *>(peer + xch_ll_offset(x.world)) +
                     ((size_t)par * x.world + x.rank) * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[i]);
            asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};"
                         ::"l"(dst + 2 * i), "r"((uint32_t)bits), "r"(epoch), "r"((uint32_t)(bits >> 32)), "r"(epoch) : "memory");
        }
    }
}

// Wait for every rank's record of the current epoch and add them up in rank order.  Called by ONE full warp;
// every lane returns the same moments.  The epoch word and the record are fetched together (the producer kernel
// of this rank advanced the epoch before this kernel started), so an arrived record costs one round trip.
__device__ __forceinline__ Moments xch_wait_merge3(const XchRef &x)
{
    const int lane = threadIdx.x & 31;
    const unsigned char *mine = x.peers[x.rank];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    bool timed_out = false;
    if (lane < x.world) {
        const long long t0 = clock64();
        unsigned ns = 32;
        while (true) {
            const uint32_t epoch = *reinterpret_cast<const volatile uint32_t *>(mine);
            const uint2 *src = reinterpret_cast<const uint2 *>(mine + xch_ll_offset(x.world)) +
                               ((size_t)(epoch & 1u) * x.world + lane) * 8;
            uint4 q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(q[i].x), "=r"(q[i].y), "=r"(q[i].z), "=r"(q[i].w) : "l"(src + 2 * i) : "memory");
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 4; ++i) ok = ok && q[i].y == epoch && q[i].w == epoch;
            if (ok) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    v[i] = __longlong_as_double((long long)(((unsigned long long)q[i].z << 32) | (unsigned long long)q[i].x));
                break;
            }
            if (clock64() - t0 > 20000000000LL) { timed_out = true; break; }  // a peer died: give up after ~10 s
            // back off: hundreds of CTAs spinning on the same few lines would slow down the very NVLink stores
            // (and the epoch update) they are waiting for
            __nanosleep(ns);
            if (ns < 256) ns <<= 1;
        }
    }
    Sums acc;
    acc.n = 0.0; acc.s1 = 0.0; acc.s2 = 0.0; acc.piv = 0.0;
    for (int r = 0; r < x.world; ++r) {  // rank order, identical on every rank and lane
        Sums q;
        q.n = __shfl_sync(0xffffffffu, v[0], r);
        q.s1 = __shfl_sync(0xffffffffu, v[1], r);
        q.s2 = __shfl_sync(0xffffffffu, v[2], r);
        q.piv = __shfl_sync(0xffffffffu, v[3], r);
        if (q.n > 0.0) {
            if (acc.n == 0.0) acc = q;
            else {
                q = rebase(q, acc.piv);
                acc.n += q.n; acc.s1 += q.s1; acc.s2 += q.s2;
            }
        }
    }
    Moments m = to_moments(acc);
    if (__any_sync(0xffffffffu, timed_out)) m.mean = __longlong_as_double(0x7ff8000000000000LL);  // poison, no trap
    return m;
}

// ---------------------------------------------------------------------------------
// Tensor memory (TMEM) as an on-chip parking area: 128 lanes x 512 columns x 32 bit per SM.
// Warp w may touch lanes [32*(w%4), 32*(w%4)+32); with the 32x32b shape thread i of the warp owns
// lane base+i and register k goes to column base+k.  SASS: STTM / LDTM.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t cols)  // one full warp; cols = 2^k in [32, 512]
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)  // the allocating warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&x)[N])
{
    static_assert(N == 8 || N == 16, "tmem_st: 8 or 16 columns");
    if constexpr (N == 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     ::"r"(taddr), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]), "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]),
                       "f"(x[8]), "f"(x[9]), "f"(x[10]), "f"(x[11]), "f"(x[12]), "f"(x[13]), "f"(x[14]), "f"(x[15])
                     : "memory");
    } else {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                     ::"r"(taddr), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]), "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7])
                     : "memory");
    }
}
// wait::ld that also "touches" the destination registers, so that no use of them can be scheduled above it
template <int N>
__device__ __forceinline__ void tmem_wait_ld_regs(float (&x)[N])
{
    static_assert(N == 8 || N == 16, "8 or 16 registers");
    if constexpr (N == 16) {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+f"(x[0]), "+f"(x[1]), "+f"(x[2]), "+f"(x[3]), "+f"(x[4]), "+f"(x[5]), "+f"(x[6]), "+f"(x[7]),
                       "+f"(x[8]), "+f"(x[9]), "+f"(x[10]), "+f"(x[11]), "+f"(x[12]), "+f"(x[13]), "+f"(x[14]), "+f"(x[15])
                     :
                     : "memory");
    } else {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+f"(x[0]), "+f"(x[1]), "+f"(x[2]), "+f"(x[3]), "+f"(x[4]), "+f"(x[5]), "+f"(x[6]), "+f"(x[7])
                     :
                     : "memory");
    }
}
// WAIT = false: the caller overlaps the load with other work and calls tmem_wait_ld_regs(x) before reading x
template <int N, bool WAIT = true>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&x)[N])
{
    static_assert(N == 8 || N == 16, "tmem_ld: 8 or 16 columns");
    if constexpr (N == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]), "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7]),
                       "=f"(x[8]), "=f"(x[9]), "=f"(x[10]), "=f"(x[11]), "=f"(x[12]), "=f"(x[13]), "=f"(x[14]), "=f"(x[15])
                     : "r"(taddr)
                     : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]), "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7])
                     : "r"(taddr)
                     : "memory");
    }
    if (WAIT) tmem_wait_ld_regs<N>(x);
}

// x / d through the correctly rounded reciprocal r = 1/d and one FMA residual step (Markstein): the IEEE
// quotient except in rare double-rounding cases (1 ulp), at 3 instructions instead of the ~30 of the
// division subroutine -- the normalise passes divide every element by the same denominator.
__device__ __forceinline__ float div_by(float x, float d, float r)
{
    const float q = x * r;
    const float e = fmaf(-q, d, x);
    return (fabsf(q) < INFINITY) ? fmaf(e, r, q) : q;
}

// ---------------------------------------------------------------------------------
// Grid exchange (see GridXch above).  gx_epoch / gx_publish / gx_collect / gx_retire are called by
// every CTA of the launch, in that order; all ranks must issue the same sequence of launches on one
// buffer set (a collective), because the epoch tag is each rank's own launch counter.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gx_epoch(const GridXch &g)  // any thread; the epoch of THIS launch
{
    return *reinterpret_cast<const volatile uint32_t *>(g.self) + 1u;
}

__device__ __forceinline__ unsigned char *gx_record(unsigned char *buf, const GridXch &g, uint32_t epoch, int src_rank, int slot)
{
    return buf + GX_HEADER_BYTES + (((size_t)(epoch & 1u) * g.world + src_rank) * g.slots + slot) * GX_RECORD_BYTES;
}

// Publish this CTA's Sums (n, s1, s2, pivot -- read from `mine`, e.g. shared memory, by every participating thread) into
// slot blockIdx.x of every rank, and empty records into the slots no CTA of this grid owns
// (blockIdx.x + k * gridDim.x < slots).  Call with all threads of the CTA.
// `peer_ptrs`: optional shared-memory copy of g.peers[0 .. world) made early in the kernel (gx_preload_peers),
// so that no dependent pointer load sits between the end of the scan and the record stores.
__device__ __forceinline__ void gx_preload_peers(const GridXch &g, unsigned char **peer_ptrs)
{
    if ((int)threadIdx.x < g.world) peer_ptrs[threadIdx.x] = g.peers[threadIdx.x];
}

__device__ __forceinline__ void gx_publish(const GridXch &g, uint32_t epoch, const double *mine,
                                           unsigned char *const *peer_ptrs = nullptr)
{
    if (peer_ptrs == nullptr) peer_ptrs = g.peers;
    const int extra = (g.slots - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;  // slots this CTA fills
    for (int i = threadIdx.x; i < extra * g.world; i += blockDim.x) {
        const int k = i / g.world, r = i - k * g.world;
        const int slot = (int)blockIdx.x + k * (int)gridDim.x;
        unsigned char *dst = gx_record(peer_ptrs[r], g, epoch, g.rank, slot);
#pragma unroll
        for (int v = 0; v < GX_RECORD_VALS; ++v) {
            const unsigned long long bits = (k == 0) ? (unsigned long long)__double_as_longlong(mine[v]) : 0ull;
            asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};"
                         ::"l"(dst + 16 * v), "r"((uint32_t)bits), "r"(epoch), "r"((uint32_t)(bits >> 32)), "r"(epoch)
                         : "memory");
        }
    }
    __syncthreads();  // the record stores of all threads are issued before the arrival hints
    if ((int)threadIdx.x < g.world) {
        uint32_t *cnt = reinterpret_cast<uint32_t *>(peer_ptrs[threadIdx.x]) + GX_ARRIVED_WORD + (epoch & 1u);
        asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(cnt), "r"((uint32_t)extra) : "memory");
    }
}

// Wait until the arrival hint says every slot of `epoch` has been filled (one thread per CTA spins on one
// word with back-off: hundreds of CTAs polling whole record arrays would congest the very L2 lines the
// late CTAs' records have to reach).  Returns false on timeout.
__device__ __forceinline__ bool gx_wait_arrivals(const GridXch &g, uint32_t epoch)
{
    const uint32_t *cnt = reinterpret_cast<const uint32_t *>(g.self) + GX_ARRIVED_WORD + (epoch & 1u);
    const uint32_t want = (uint32_t)(g.world * g.slots);
    const long long t_start = clock64();
    unsigned ns = 20;
    while (true) {
        uint32_t v;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        if (v >= want) return true;
        if (clock64() - t_start > 20000000000LL) return false;
        __nanosleep(ns);
        if (ns < 160) ns <<= 1;
    }
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Wait for all world x slots records of `epoch` in this rank's buffer and add them up.  Records are dealt
// to threads round-robin; each thread keeps up to four polls in flight (one L2 round trip per batch) and
// adds its records in index order around the first pivot it meets (rebase: no divisions anywhere), then
// the block reduction agrees on one pivot and adds in a fixed order: the same operations on the same data
// in every CTA of every rank, so all of them derive identical bits.  Result (as moments: the only
// divisions of the whole exchange) in thread 0.  `scratch` >= 32 Sums.  All threads of the CTA must call.
__device__ __forceinline__ Moments gx_collect(const GridXch &g, uint32_t epoch, Sums *scratch)
{
    unsigned char *mine = g.self;
    const int n_rec = g.world * g.slots;
    Sums acc;
    acc.n = 0.0; acc.s1 = 0.0; acc.s2 = 0.0; acc.piv = 0.0;
    const long long t_start = clock64();
    bool timed_out = false;
    for (int base = threadIdx.x; base < n_rec; base += 4 * blockDim.x) {
        double v[4][GX_RECORD_VALS];
        unsigned pending = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (base + j * (int)blockDim.x < n_rec) pending |= 1u << j;
#pragma unroll
            for (int c = 0; c < GX_RECORD_VALS; ++c) v[j][c] = 0.0;
        }
        while (pending != 0u) {
            uint4 q[4][GX_RECORD_VALS];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (pending & (1u << j)) {
                    const int idx = base + j * (int)blockDim.x;  // = src_rank * slots + slot
                    const unsigned char *src = mine + GX_HEADER_BYTES + ((size_t)(epoch & 1u) * n_rec + idx) * GX_RECORD_BYTES;
#pragma unroll
                    for (int c = 0; c < GX_RECORD_VALS; ++c)
                        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(q[j][c].x), "=r"(q[j][c].y), "=r"(q[j][c].z), "=r"(q[j][c].w)
                                     : "l"(src + 16 * c)
                                     : "memory");
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (pending & (1u << j)) {
                    bool ok = true;
#pragma unroll
                    for (int c = 0; c < GX_RECORD_VALS; ++c) ok = ok && q[j][c].y == epoch && q[j][c].w == epoch;
                    if (ok) {
#pragma unroll
                        for (int c = 0; c < GX_RECORD_VALS; ++c)
                            v[j][c] = __longlong_as_double((long long)(((unsigned long long)q[j][c].z << 32) | (unsigned long long)q[j][c].x));
                        pending &= ~(1u << j);
                    }
                }
            }
            if (pending != 0u && clock64() - t_start > 20000000000LL) {  // ~10 s: a peer never arrived
                timed_out = true;
                break;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            Sums q;
            q.n = v[j][0]; q.s1 = v[j][1]; q.s2 = v[j][2]; q.piv = v[j][3];
            if (q.n > 0.0) {
                if (acc.n == 0.0) acc = q;
                else {
                    q = rebase(q, acc.piv);
                    acc.n += q.n; acc.s1 += q.s1; acc.s2 += q.s2;
                }
            }
        }
        if (timed_out) break;
    }
    if (timed_out) reinterpret_cast<volatile uint32_t *>(mine)[2] = 1u;  // no trap: plb_grid_exchange_status() reports it
    const int any_timeout = __syncthreads_or(timed_out ? 1 : 0);
    const Sums all = block_reduce_sums(acc, scratch);
    Moments m = to_moments(all);
    if (any_timeout) m.mean = __longlong_as_double(0x7ff8000000000000LL);  // poison the result
    return m;
}

// Last CTA out advances the epoch for the next launch (off the critical path).  One thread per CTA,
// after that CTA's gx_collect.
__device__ __forceinline__ void gx_retire(const GridXch &g, uint32_t epoch)
{
    uint32_t *hdr = reinterpret_cast<uint32_t *>(g.self);
    __threadfence();
    const uint32_t t = atomicAdd(hdr + GX_TICKET_WORD, 1u);
    if (t == gridDim.x - 1) {
        hdr[GX_TICKET_WORD] = 0u;
        // every rank's hints of this epoch have arrived (all CTAs here waited for them), and nobody can send
        // hints of epoch + 2 before this rank has published epoch + 1, i.e. after this launch: safe to clear
        hdr[GX_ARRIVED_WORD + (epoch & 1u)] = 0u;
        __threadfence();
        *reinterpret_cast<volatile uint32_t *>(hdr) = epoch;
    }
}

// running_mean_std.py:23-30 for one feature: (mean, var) <- Chan merge with batch moments
// (b_n, b_mean, b_m2); `cnt` is the running count BEFORE the update.  Same operation order as
// the reference; PLB_MOMENTS_F32_FLOW reproduces numba's float32 `batch_var * batch_count`.
__device__ __forceinline__ void chan_update_scalar(double b_n, double b_mean, double b_m2, double cnt, int flags,
                                                   double &mean, double &var)
{
    double batch_mean = b_mean;
    double batch_var = b_m2 / b_n;
    double m_b;
    if (flags & PLB_MOMENTS_F32_FLOW) {
        batch_mean = (double)__double2float_rn(batch_mean);
        const float bv32 = __double2float_rn(batch_var);
        m_b = (double)__fmul_rn(bv32, __double2float_rn(b_n));  // float32 array * int64 scalar -> float32
    } else {
        m_b = __dmul_rn(batch_var, b_n);
    }
    const double delta = __dsub_rn(batch_mean, mean);
    const double total = __dadd_rn(b_n, cnt);
    const double new_mean = __dadd_rn(mean, __ddiv_rn(__dmul_rn(delta, b_n), total));
    const double m_a = __dmul_rn(var, cnt);
    const double cross = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), cnt), b_n), total);
    const double m_2 = __dadd_rn(__dadd_rn(m_a, m_b), cross);
    mean = new_mean;
    var = __ddiv_rn(m_2, total);
}

// ---------------------------------------------------------------------------------
// Grid-wide moments tail: every CTA contributes one partial (division-free Sums); the last CTA
// to arrive rebases all partials to one pivot, adds them in a fixed order (deterministic) and
// performs the only divisions of the whole reduction to write (count, mean, M2).
// Must be called by ALL threads of EVERY CTA of the grid.  `scratch` >= 32 Sums.
// ---------------------------------------------------------------------------------
// WARP_PIVOT: `mine.piv` is uniform within every warp (the scans broadcast lane 0's first value).
template <bool WARP_PIVOT = false>
__device__ __forceinline__ void moments_tail(const Sums &mine, const MomentsTail &tail, Sums *scratch,
                                             const XchPre xch_pre = xch_no_preload())
{
    __shared__ bool is_last;
    __syncthreads();  // scratch may have been used before
    const Sums blk = WARP_PIVOT ? block_reduce_sums_warp_pivot(mine, scratch) : block_reduce_sums(mine, scratch);
    if (tail.partials_only) {
        if (threadIdx.x == 0) {
            tail.partials[blockIdx.x] = blk;
            if (blockIdx.x == 0) tail.ticket[1] = gridDim.x;
        }
        return;
    }
    if (threadIdx.x == 0) {
        tail.partials[blockIdx.x] = blk;
        __threadfence();
        const unsigned t = atomicAdd(tail.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tail.dbg != nullptr && threadIdx.x == 0) tail.dbg[1] = global_timer_ns();
    // contiguous slices per thread keep the summation order = CTA index order
    const unsigned n = gridDim.x;
    const unsigned per = (n + blockDim.x - 1) / blockDim.x;
    const unsigned lo = threadIdx.x * per;
    Sums acc;
    acc.n = 0.0; acc.s1 = 0.0; acc.s2 = 0.0; acc.piv = 0.0;
    for (unsigned i = lo; i < lo + per && i < n; ++i) {
        const volatile Sums *vp = tail.partials + i;
        Sums q;
        q.n = vp->n; q.s1 = vp->s1; q.s2 = vp->s2; q.piv = vp->piv;
        if (q.n > 0.0) {
            if (acc.n == 0.0) acc = q;
            else {
                q = rebase(q, acc.piv);
                acc.n += q.n; acc.s1 += q.s1; acc.s2 += q.s2;
            }
        }
    }
    __syncthreads();
    acc = block_reduce_sums(acc, scratch);
    if (tail.xch.peers != nullptr) {
        // multi-GPU: hand the rank's moments to every peer; whoever consumes them waits and merges
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) *tail.ticket = 0u;
            if (tail.dbg != nullptr && threadIdx.x == 0) tail.dbg[2] = global_timer_ns();
            xch_push3(tail.xch, acc, xch_pre);  // lane 0's sums travel; the consumer writes the merged moments_out
            if (tail.dbg != nullptr && threadIdx.x == 0) tail.dbg[3] = global_timer_ns();
        }
        return;
    }
    if (threadIdx.x == 0) {
        const Moments m = to_moments(acc);
        tail.moments_out[0] = m.n;
        tail.moments_out[1] = m.mean;
        tail.moments_out[2] = m.m2;
        if (tail.upd_mean != nullptr) {
            double mean = *tail.upd_mean, var = *tail.upd_var;
            const double cnt = *tail.upd_count;
            chan_update_scalar(m.n, m.mean, m.m2, cnt, tail.upd_flags, mean, var);
            *tail.upd_mean = mean;
            *tail.upd_var = var;
            *tail.upd_count = m.n + cnt;
        }
        *tail.ticket = 0u;  // ready for the next launch
    }
}
// Block-cooperative reduction of the partial Sums a producer grid published with `partials_only`:
// every CTA of the consumer derives the same moments (same data, same order).  Result in thread 0.
__device__ __forceinline__ Moments reduce_published_partials(const Sums *partials, unsigned n, Sums *scratch)
{
    Sums acc;
    acc.n = 0.0; acc.s1 = 0.0; acc.s2 = 0.0; acc.piv = 0.0;
    const unsigned per = (n + blockDim.x - 1) / blockDim.x;
    const unsigned lo = threadIdx.x * per;
    for (unsigned i = lo; i < lo + per && i < n; ++i) {
        Sums q = partials[i];
        if (q.n > 0.0) {
            if (acc.n == 0.0) acc = q;
            else {
                q = rebase(q, acc.piv);
                acc.n += q.n; acc.s1 += q.s1; acc.s2 += q.s2;
            }
        }
    }
    acc = block_reduce_sums(acc, scratch);
    return to_moments(acc);
}
// Same reduction by ONE warp (no block barrier inside): lanes take partials lane, lane+32, ... in index
// order, rebase onto the first non-empty pivot of the warp, then plain shuffle adds.  Every lane returns
// the same moments; summation order depends only on n.
// The count and the first four partials of every lane are fetched in ONE round trip (slots beyond the count exist --
// the workspace holds MOMENTS_MAX_PARTIALS of them -- and are masked afterwards): the seven other warps of the CTA
// sit at a barrier until this returns, and a chain of dependent L2 loads (count, then partial after partial) was 40 %
// of the normalise kernel's stall samples.
__device__ __forceinline__ Moments reduce_published_partials_warp(const Sums *partials, const unsigned *n_partials)
{
    static_assert(MOMENTS_MAX_PARTIALS >= 128, "the first four partials per lane are fetched before their count is known");
    const unsigned lane = threadIdx.x & 31u;
    Sums q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 *src = reinterpret_cast<const double2 *>(partials + lane + 32u * k);
        const double2 lo = __ldcg(src), hi = __ldcg(src + 1);
        q[k].n = lo.x; q[k].s1 = lo.y; q[k].s2 = hi.x; q[k].piv = hi.y;
    }
    const unsigned n = __ldcg(n_partials);
    Sums acc;
    acc.n = 0.0; acc.s1 = 0.0; acc.s2 = 0.0; acc.piv = 0.0;
    for (unsigned base = lane;; base += 128u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (base + 32u * k < n && q[k].n > 0.0) {
                if (acc.n == 0.0) acc = q[k];
                else {
                    const Sums r = rebase(q[k], acc.piv);
                    acc.n += r.n; acc.s1 += r.s1; acc.s2 += r.s2;
                }
            }
        }
        if (base + 128u >= n) break;  // (uniform enough: lanes beyond n just carry empty sums)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned i = base + 128u + 32u * k;
            if (i < n) {
                const double2 *src = reinterpret_cast<const double2 *>(partials + i);
                const double2 lo = __ldcg(src), hi = __ldcg(src + 1);
                q[k].n = lo.x; q[k].s1 = lo.y; q[k].s2 = hi.x; q[k].piv = hi.y;
            } else {
                q[k].n = 0.0;
            }
        }
    }
    const unsigned has = __ballot_sync(0xffffffffu, acc.n > 0.0);
    const int src_lane = has ? (__ffs(has) - 1) : 0;
    const double P = __shfl_sync(0xffffffffu, acc.piv, src_lane);
    acc = rebase(acc, P);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc.n += __shfl_xor_sync(0xffffffffu, acc.n, d);
        acc.s1 += __shfl_xor_sync(0xffffffffu, acc.s1, d);
        acc.s2 += __shfl_xor_sync(0xffffffffu, acc.s2, d);
    }
    acc.piv = P;
    return to_moments(acc);
}
// (experiments, plb_set_tuning key 21 = 1: the round-2 loop, count first, then one partial after the other)
__device__ __forceinline__ Moments reduce_published_partials_warp_serial(const Sums *partials, unsigned n)
{
    const unsigned lane = threadIdx.x & 31u;
    Sums acc;
    acc.n = 0.0; acc.s1 = 0.0; acc.s2 = 0.0; acc.piv = 0.0;
    for (unsigned i = lane; i < n; i += 32u) {
        Sums q = partials[i];
        if (q.n > 0.0) {
            if (acc.n == 0.0) acc = q;
            else {
                q = rebase(q, acc.piv);
                acc.n += q.n; acc.s1 += q.s1; acc.s2 += q.s2;
            }
        }
    }
    const unsigned has = __ballot_sync(0xffffffffu, acc.n > 0.0);
    const int src = has ? (__ffs(has) - 1) : 0;
    const double P = __shfl_sync(0xffffffffu, acc.piv, src);
    acc = rebase(acc, P);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc.n += __shfl_xor_sync(0xffffffffu, acc.n, d);
        acc.s1 += __shfl_xor_sync(0xffffffffu, acc.s1, d);
        acc.s2 += __shfl_xor_sync(0xffffffffu, acc.s2, d);
    }
    acc.piv = P;
    return to_moments(acc);
}
#endif  // __CUDACC__

}  // namespace plb
