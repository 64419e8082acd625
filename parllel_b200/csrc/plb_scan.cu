// plb_scan.cu -- masked linear-recurrence scans over the time axis of [T, B] batches:
//   GAE advantage            (parllel/transforms/advantage.py:36-57)
//   discounted return        (parllel/transforms/advantage.py:11-33)
//   past discounted return   (parllel/transforms/norm_rewards.py:16-33)
//
// Two implementations behind one C ABI:
//
//   SERIAL   one thread per column walks T with the reference's exact numerics (float64
//            step, float32 store, no FMA contraction) -> bit-identical to oracle/plb_oracle.c.
//            Memory-bound once B*inner >= ~32k columns; also the path for shapes TMA cannot
//            describe (row strides not multiples of 16 bytes, trailing agent axis).
//
//   CHUNKED  persistent CTAs stream [TC x 32]-element tiles of each 32-column strip through a
//            multi-stage TMA (cp.async.bulk.tensor) + mbarrier ring.  Inside a tile each warp
//            scans its own rows as an affine map  x_t = A_t + C_t * carry  (float64), the
//            per-warp aggregates are composed through shared memory, and the carry of the strip
//            lives in registers across tiles.  HBM traffic is the compulsory bytes only.
//            Rounding differs from the per-step float32 rounding of the reference by <= ~3e-6
//            absolute at C1 sizes (tolerance stated in tests/test_gpu_scan.py).
#include <cuda.h>
#include <math.h>

#include <type_traits>

#include "plb_common.cuh"
#include "plb_scan_shared.cuh"

namespace plb {

// =====================================================================================
// SERIAL kernels (strict reference numerics)
// =====================================================================================
constexpr int SERIAL_U = 8;  // rows prefetched per register block

// MODE 0: GAE, MODE 1: discounted return.  Column j in [0, B*inner); reward/done column j/inner.
template <int MODE>
__global__ void __launch_bounds__(128, 4)
scan_reverse_serial_kernel(const float *__restrict__ reward, int64_t ld_r,
                           const float *__restrict__ value, int64_t ld_v,
                           const uint8_t *__restrict__ done, int64_t ld_d,
                           const float *__restrict__ boot,
                           float *__restrict__ adv, int64_t ld_a,
                           float *__restrict__ ret, int64_t ld_t,
                           int64_t T, int64_t ncols, int64_t inner, double gamma, double lambda)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const int64_t jr = (inner == 1) ? j : j / inner;
    const double gl = __dmul_rn(gamma, lambda);

    float rA[SERIAL_U], vA[SERIAL_U], rB[SERIAL_U], vB[SERIAL_U];
    uint8_t dA[SERIAL_U], dB[SERIAL_U];

    // running pointers (one 64-bit register pair per array instead of one per unrolled row)
    const float *rp = reward + (T - 1) * ld_r + jr;
    const float *vp = value + (T - 1) * ld_v + j;
    const uint8_t *dp = done + (T - 1) * ld_d + jr;
    float *ap = adv + (T - 1) * ld_a + j;
    float *tp = ret + (T - 1) * ld_t + j;
    int64_t t_load = T - 1;  // row the load pointers address

    auto load_block = [&](float(&rb)[SERIAL_U], float(&vb)[SERIAL_U], uint8_t(&db)[SERIAL_U]) {
#pragma unroll
        for (int u = 0; u < SERIAL_U; ++u) {
            if (t_load >= 0) {
                rb[u] = *rp;
                vb[u] = *vp;
                db[u] = *dp;
            }
            rp -= ld_r; vp -= ld_v; dp -= ld_d;
            --t_load;
        }
    };

    double vnext = (double)boot[j];  // value[t+1]; bootstrap for t = T-1
    float carry = 0.f;               // advantage[t+1] (GAE) or return_[t+1]
    bool last = true;
    int64_t t = T - 1;               // row being computed

    auto compute_block = [&](const float(&rb)[SERIAL_U], const float(&vb)[SERIAL_U],
                             const uint8_t(&db)[SERIAL_U]) {
#pragma unroll
        for (int u = 0; u < SERIAL_U; ++u) {
            if (t < 0) break;
            const double r = (double)rb[u];
            const float vf = vb[u];
            const double v = (double)vf;
            const double nd = db[u] ? 0.0 : 1.0;
            if (MODE == 0) {
                // delta = r + g*v[t+1]*nd - v ; adv = delta + g*l*nd*adv[t+1]
                const double delta = __dsub_rn(__dadd_rn(r, __dmul_rn(__dmul_rn(gamma, vnext), nd)), v);
                const double a = last ? delta
                                      : __dadd_rn(delta, __dmul_rn(__dmul_rn(gl, nd), (double)carry));
                carry = __double2float_rn(a);
                *ap = carry;
                *tp = __fadd_rn(carry, vf);
                vnext = v;
            } else {
                // ret[T-1] = r + g*boot*nd ; ret[t] = r + ret[t+1]*g*nd
                const double x = last ? __dadd_rn(r, __dmul_rn(__dmul_rn(gamma, vnext), nd))
                                      : __dadd_rn(r, __dmul_rn(__dmul_rn((double)carry, gamma), nd));
                carry = __double2float_rn(x);
                *tp = carry;
                *ap = __fsub_rn(carry, vf);
            }
            last = false;
            ap -= ld_a; tp -= ld_t;
            --t;
        }
    };

    load_block(rA, vA, dA);
    while (t >= 0) {
        load_block(rB, vB, dB);
        compute_block(rA, vA, dA);
        load_block(rA, vA, dA);
        compute_block(rB, vB, dB);
    }
}

__global__ void __launch_bounds__(128, 4)
scan_forward_serial_kernel(const float *__restrict__ reward, int64_t ld_r,
                           const uint8_t *__restrict__ done, int64_t ld_d,
                           const float *__restrict__ prev_ret, const uint8_t *__restrict__ prev_done,
                           float *__restrict__ out, int64_t ld_o,
                           int64_t T, int64_t ncols, double gamma)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    float rA[SERIAL_U], rB[SERIAL_U];
    uint8_t dA[SERIAL_U], dB[SERIAL_U];
    const float *rp = reward + j;
    const uint8_t *dp = done + j;
    float *op = out + j;
    int64_t t_load = 0, t = 0;
    auto load_block = [&](float(&rb)[SERIAL_U], uint8_t(&db)[SERIAL_U]) {
#pragma unroll
        for (int u = 0; u < SERIAL_U; ++u) {
            if (t_load < T) {
                rb[u] = *rp;
                db[u] = *dp;
            }
            rp += ld_r; dp += ld_d;
            ++t_load;
        }
    };
    float carry = prev_ret[j];
    double nd = prev_done[j] ? 0.0 : 1.0;  // not_done[t-1]
    auto compute_block = [&](const float(&rb)[SERIAL_U], const uint8_t(&db)[SERIAL_U]) {
#pragma unroll
        for (int u = 0; u < SERIAL_U; ++u) {
            if (t >= T) break;
            const double x = __dadd_rn((double)rb[u], __dmul_rn(__dmul_rn((double)carry, gamma), nd));
            carry = __double2float_rn(x);
            *op = carry;
            nd = db[u] ? 0.0 : 1.0;
            op += ld_o;
            ++t;
        }
    };
    load_block(rA, dA);
    while (t < T) {
        load_block(rB, dB);
        compute_block(rA, dA);
        load_block(rA, dA);
        compute_block(rB, dB);
    }
}

// =====================================================================================
// CHUNKED kernels (TMA + mbarrier ring, affine-map scan)
// =====================================================================================
namespace chunked {

constexpr int W = 32;        // columns per strip (one lane per column)
constexpr int MAX_STAGES = 12;

// one input pipeline stage: three TMA boxes of TC rows x 32 columns
// (TC = NW warps x R rows; warp w owns rows [w*R, (w+1)*R) of a tile)
template <int TC>
struct __align__(128) Stage {
    float r[TC][W];
    float v[TC][W];
    uint8_t d[TC][W];
};

// one output staging buffer (TMA-stored to global; the TMA clips rows/columns outside the tensor)
template <int TC, int N_OUT>
struct __align__(128) OutTile {
    float o[N_OUT][TC][W];
};

template <typename ACC, int NW>
struct Shared {
    ACC aggA[2][NW][W];
    ACC aggC[2][NW][W];
    ACC carry[2][W];
    float halo[2][W];
    unsigned long long full[MAX_STAGES];
};

struct Params {
    const float *boot;            // reverse: bootstrap value [B]
    const float *prev_ret;        // forward: previous return [B]
    const uint8_t *prev_done;     // forward: previous done [B]
    int64_t T, B;
    double gamma, lambda;
    int n_stages;
    int n_strips, n_chunks;
    // optional fused reward transform (NormalizeRewards scale + ClipRewards) on the load path
    const double *reward_var;     // device float64 running variance, or NULL
    double reward_eps;
    float clip_lo, clip_hi;       // -inf / +inf when absent
    // optional fused moments of out0
    const uint8_t *valid; int64_t ld_valid;
    MomentsTail tail;
    // forward scan, optional: out[T-1] and done[T-1] per column for the caller's carry rows
    float *carry_ret_out; uint8_t *carry_done_out; const uint8_t *done_raw; int64_t ld_done_raw;
    int debug;                    // experiments only (plb_set_tuning key 5): 1 = skip the in-loop sums, 2 = skip the tail
    // resident variant (NormalizeAdvantage fused into the reverse scan)
    GridXch gx;
    float adv_eps;
    double *adv_moments_out;
    float *out0_raw; int64_t ld_out0_raw;   // out0 as a plain pointer (ragged rows of the final store phase)
    int tmem_cols;                // TMEM columns to allocate: power of two >= tiles per CTA x 2 x R
    unsigned long long *dbg;      // experiments: per-CTA phase timestamps [grid][8] (plb_debug_set_buffer), or NULL
    int poll_mode;                // 0: arrival hint, then one batched read of the records (default); 1: no hint, poll the records
    int n_out_bufs;               // whole-tile staging buffers of the final store phase carved from the input ring
};

struct TensorMaps {
    CUtensorMap r, v, d;      // inputs
    CUtensorMap o0, o1, o2;   // outputs: (advantage | past_return), return_, transformed reward
    // Reverse scans align tiles to the END of the batch, so the earliest tile may start at a
    // negative row.  TMA loads zero-fill there, but a TMA STORE may not start out of bounds:
    // that ragged tile is stored through maps whose box is only T % TC rows tall, from row 0.
    CUtensorMap o0_rag, o1_rag, o2_rag;
};

__device__ __forceinline__ float to_f32(double x) { return __double2float_rn(x); }
__device__ __forceinline__ float to_f32(float x) { return x; }

// MODE 0: GAE, 1: discounted return (both reverse), 2: past return (forward).
// ACC: arithmetic type of the affine maps (double; a float variant measured only ~4 % faster and was removed).
// Output slots: MODE 0: o0 = advantage, o1 = return_; MODE 1: same; MODE 2: o0 = past_return;
//               HAS_PRE adds the transformed reward as the last slot.
//
// ---------------------------------------------------------------------------------
// Software-pipelined tile loop: the two register passes of consecutive tiles are fused -- between two
// CTA barriers every warp finalises tile j (independent FMAs + staging stores) while it scans tile j+1
// (the dependent chain), row by row in one unrolled loop, so the chain's latency hides behind the
// independent work instead of being exposed twice per tile.  The aggregates of the warps are
// fetched with independent shared-memory loads before the (predicated) compose chain.
// ---------------------------------------------------------------------------------
//
// RESIDENT (reverse scans only) fuses NormalizeAdvantage (norm_advantage.py:30-56) into the same launch:
// the un-normalised advantage of every tile is parked in TENSOR MEMORY (tcgen05.st, 16 columns per
// warp and tile -- the 256 KB of TMEM are otherwise idle on this path) instead of being written out;
// after the last tile every CTA publishes its partial moments to all CTAs of all ranks (gx_publish /
// gx_collect: grid barrier and multi-GPU all-gather in one step), then reads its tiles back
// (tcgen05.ld), normalises them in float32 and stores the advantage ONCE, each warp streaming its own
// R-row slabs through the (now idle) input ring by TMA.  HBM traffic = the compulsory 17 B/transition.
// The grid must be co-resident (cooperative launch, one CTA per SM).
template <int MODE, typename ACC, int NW, int R, bool HAS_PRE, bool HAS_STATS, bool RESIDENT = false>
__global__ void __launch_bounds__(NW * 32)
scan_pipelined_kernel(const __grid_constant__ TensorMaps tm, const Params p)
{
    constexpr int TC = NW * R;
    constexpr bool REVERSE = (MODE != 2);
    static_assert(!RESIDENT || (REVERSE && HAS_STATS), "the resident variant is a reverse scan with statistics");
    constexpr int N_OUT = (REVERSE ? 2 : 1) + (HAS_PRE ? 1 : 0);
    constexpr int N_STG = RESIDENT ? N_OUT - 1 : N_OUT;   // staging slots: out0 does not pass through smem when resident
    constexpr int RET_SLOT = RESIDENT ? 0 : 1;
    constexpr int PRE_SLOT = N_STG - 1;
    constexpr uint32_t TX_BYTES = REVERSE ? (sizeof(float) * TC * W * 2 + TC * W) : (sizeof(float) * TC * W + TC * W);
    using StageT = Stage<TC>;
    using OutT = OutTile<TC, N_STG>;
    using SharedT = Shared<ACC, NW>;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    StageT *stages = reinterpret_cast<StageT *>(smem_raw);
    OutT *outs = reinterpret_cast<OutT *>(smem_raw + sizeof(StageT) * p.n_stages);
    SharedT *sh = reinterpret_cast<SharedT *>(smem_raw + sizeof(StageT) * p.n_stages + 2 * sizeof(OutT));
    __shared__ Sums red_scratch[32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_stages = p.n_stages, n_chunks = p.n_chunks, n_strips = p.n_strips;
    const int64_t T = p.T;

    __shared__ uint32_t s_tmem_base, s_epoch;
    __shared__ unsigned char *s_peers[GX_MAX_WORLD];
    // programmatic dependent launch: the next kernel of the stream may be scheduled from here on (it parks in its own
    // griddepcontrol.wait until this grid has completed); this grid's shared-memory prologue overlaps its predecessor
    griddep_launch_dependents();
    if (RESIDENT) griddep_wait();
    if (RESIDENT) gx_preload_peers(p.gx, s_peers);
    if (RESIDENT && p.dbg != nullptr && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 0] = global_timer_ns();
    if (!RESIDENT && HAS_STATS && p.tail.dbg != nullptr && threadIdx.x == 0 && blockIdx.x == 0) p.tail.dbg[0] = global_timer_ns();
    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&sh->full[s], 1);
        fence_proxy_async_smem();
        if (RESIDENT) s_epoch = gx_epoch(p.gx);
    }
    if (RESIDENT) {
        if (warp == 0) {
            __syncwarp();  // .sync.aligned: the whole warp, converged
            tmem_alloc(&s_tmem_base, (uint32_t)p.tmem_cols);
        }
        tmem_fence_before_sync();
    }
    // the tensor maps are kernel parameters: their fetch can run ahead of the predecessor's completion
    if (!RESIDENT && threadIdx.x == 0) {
        tma_prefetch_desc(&tm.r);
        if (REVERSE) tma_prefetch_desc(&tm.v);
        tma_prefetch_desc(&tm.d);
    }
    if (!RESIDENT) griddep_wait();  // everything below reads or writes global memory the predecessor may own
    __syncthreads();
    // sharded: what the last CTA's push will need from memory, fetched now by the warp that pushes (two dependent L2
    // round trips off the tail, which every rank's normalise kernel waits for)
    // (parked in shared memory, not in registers: the tile loop has none to spare)
    __shared__ unsigned char *s_xch_peer[32];
    __shared__ unsigned char *s_xch_mine;
    __shared__ uint32_t s_xch_epoch;
    const bool xch_preloaded = !RESIDENT && HAS_STATS && p.tail.xch.peers != nullptr && !p.tail.partials_only;
    if (xch_preloaded && warp == 0) {
        const XchPre pre = xch_preload(p.tail.xch);
        s_xch_peer[lane] = pre.peer;
        if (lane == 0) { s_xch_mine = pre.mine; s_xch_epoch = pre.epoch; }
    }
    uint32_t tmem_warp = 0;  // this warp's lane quarter and column half of tile 0
    if (RESIDENT) {
        tmem_fence_after_sync();
        tmem_warp = s_tmem_base + ((uint32_t)(warp & 3) << 21) + (uint32_t)((warp >> 2) * R);
    }

    // ---- producer state (thread 0 only) -------------------------------------------------
    int iss_strip = blockIdx.x, iss_chunk = 0, iss_stage = 0;
    auto issue_one = [&]() {
        const int c0 = iss_strip * W;
        const int t0 = REVERSE ? (int)T - (iss_chunk + 1) * TC : iss_chunk * TC;
        unsigned long long *bar = &sh->full[iss_stage];
        StageT &S = stages[iss_stage];
        mbar_arrive_expect_tx(bar, TX_BYTES);
        tma_load_2d(&S.r[0][0], &tm.r, c0, t0, bar);
        if (REVERSE) {
            tma_load_2d(&S.v[0][0], &tm.v, c0, t0, bar);
            tma_load_2d(&S.d[0][0], &tm.d, c0, t0, bar);
        } else {
            tma_load_2d(&S.d[0][0], &tm.d, c0, t0 - 1, bar);  // done[t-1]
        }
        if (++iss_chunk == n_chunks) { iss_chunk = 0; iss_strip += gridDim.x; }
        if (++iss_stage == n_stages) iss_stage = 0;
    };
    // ---- store state (thread 0 only) ----------------------------------------------------
    bool st_pending = false;
    int st_c0 = 0, st_t0 = 0;
    auto store_tile = [&](int buf) {
        if (REVERSE && st_t0 < 0) {
            const int skip = -st_t0;
            if (!RESIDENT) tma_store_2d(&tm.o0_rag, &outs[buf].o[0][skip][0], st_c0, 0);
            tma_store_2d(&tm.o1_rag, &outs[buf].o[RET_SLOT][skip][0], st_c0, 0);
            if (HAS_PRE) tma_store_2d(&tm.o2_rag, &outs[buf].o[PRE_SLOT][skip][0], st_c0, 0);
        } else {
            if (!RESIDENT) tma_store_2d(&tm.o0, &outs[buf].o[0][0][0], st_c0, st_t0);
            if (REVERSE) tma_store_2d(&tm.o1, &outs[buf].o[RET_SLOT][0][0], st_c0, st_t0);
            if (HAS_PRE) tma_store_2d(&tm.o2, &outs[buf].o[PRE_SLOT][0][0], st_c0, st_t0);
        }
        tma_store_commit();
    };
    if (threadIdx.x == 0) {
        if (RESIDENT) {
            tma_prefetch_desc(&tm.r);
            if (REVERSE) tma_prefetch_desc(&tm.v);
            tma_prefetch_desc(&tm.d);
        }
        for (int s = 0; s < n_stages && iss_strip < n_strips; ++s) issue_one();
        if (!RESIDENT) tma_prefetch_desc(&tm.o0);
        if (REVERSE) tma_prefetch_desc(&tm.o1);
        if (HAS_PRE) tma_prefetch_desc(&tm.o2);
    }

    const ACC gamma = (ACC)p.gamma;
    const float gamma_f = (float)p.gamma;
    (void)gamma_f;
    const ACC gl = (ACC)(p.gamma * p.lambda);
    double inv_denom = 1.0;
    float clip_lo = 0.f, clip_hi = 0.f;
    if (HAS_PRE) {
        if (p.reward_var != nullptr) inv_denom = 1.0 / sqrt(*p.reward_var + p.reward_eps);
        clip_lo = p.clip_lo;
        clip_hi = p.clip_hi;
    }

    float st_piv = 0.f;
    bool st_has = false;
    double st_s1 = 0.0, st_s2 = 0.0;
    unsigned st_n = 0;

    const int i0 = warp * R;
    const int n_my_strips = blockIdx.x < n_strips ? (n_strips - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_tiles = n_my_strips * n_chunks;

    // register state of the tile between its two passes
    ACC A[R], C[R];
    float vreg[R], rpre[HAS_PRE ? R : 1];
    (void)vreg; (void)rpre;
#pragma unroll
    for (int u = 0; u < R; ++u) { A[u] = (ACC)0; C[u] = (ACC)0; vreg[u] = 0.f; }

    // cursor of the tile being scanned (pass 1) and of the tile being finalised (pass 2)
    int nx_strip = blockIdx.x, nx_k = 0;
    int cu_strip = 0, cu_k = 0;
    int cu_idx = -1;  // index of the tile being finalised among this CTA's tiles (resident: its TMEM slot)
    ACC nx_carry0 = (ACC)0, cu_carry0 = (ACC)0;
    int stage = 0, par = 0;  // input stage and staging/aggregate parity of the tile being scanned
    uint32_t phase = 0;

    auto body = [&](auto p2_tag, auto p1_tag) {
        constexpr bool P2 = decltype(p2_tag)::value, P1 = decltype(p1_tag)::value;
        const int pc = par ^ 1;  // parity of the tile being finalised
        ACC cin = (ACC)0;
        int64_t cu_t0 = 0;
        if (P2) {
            cu_t0 = REVERSE ? T - (int64_t)(cu_k + 1) * TC : (int64_t)cu_k * TC;
            // outs[pc] was handed to the TMA two tiles ago (store issued one iteration back): it must
            // have been read before this iteration overwrites it
            if (threadIdx.x == 0) tma_store_wait_read<0>();
            __syncthreads();  // aggregates of this tile visible; its input stage fully read; outs[par] complete
            if (threadIdx.x == 0) {
                if (iss_strip < n_strips) issue_one();
                if (st_pending) store_tile(par);
                st_pending = true;
                st_c0 = cu_strip * W;
                st_t0 = (int)cu_t0;
            }
            ACC aC[NW - 1], aA[NW - 1];
            if (REVERSE) {
#pragma unroll
                for (int w = 1; w < NW; ++w) { aC[w - 1] = sh->aggC[pc][w][lane]; aA[w - 1] = sh->aggA[pc][w][lane]; }
            } else {
#pragma unroll
                for (int w = 0; w < NW - 1; ++w) { aC[w] = sh->aggC[pc][w][lane]; aA[w] = sh->aggA[pc][w][lane]; }
            }
            cin = (cu_k == 0) ? cu_carry0 : sh->carry[pc ^ 1][lane];
            if (REVERSE) {
#pragma unroll
                for (int w = NW - 1; w >= 1; --w) if (w > warp) cin = fma(aC[w - 1], cin, aA[w - 1]);
                if (warp == 0) sh->carry[pc][lane] = fma(C[0], cin, A[0]);
            } else {
#pragma unroll
                for (int w = 0; w < NW - 1; ++w) if (w < warp) cin = fma(aC[w], cin, aA[w]);
                if (warp == NW - 1) sh->carry[pc][lane] = fma(C[R - 1], cin, A[R - 1]);
            }
        }

        // ---- set-up of the tile being scanned ------------------------------------------
        const StageT &S = stages[stage];
        float vnext_f = 0.f;
        bool first_prev_done = false;
        if (P1) {
            const int64_t ncol = (int64_t)nx_strip * W + lane;
            const bool ncol_ok = ncol < p.B;
            if (nx_k == 0) {
                if (MODE == 0) nx_carry0 = (ACC)0;
                else if (MODE == 1) nx_carry0 = ncol_ok ? (ACC)p.boot[ncol] : (ACC)0;
                else nx_carry0 = ncol_ok ? (ACC)p.prev_ret[ncol] : (ACC)0;
            }
            if (MODE == 0 && warp == NW - 1 && nx_k == 0) vnext_f = ncol_ok ? p.boot[ncol] : 0.f;
            if (MODE == 2 && warp == 0 && nx_k == 0) first_prev_done = ncol_ok ? (p.prev_done[ncol] != 0) : false;
            mbar_wait(&sh->full[stage], phase);
            if (MODE == 0) {
                if (warp == NW - 1) { if (nx_k != 0) vnext_f = sh->halo[par ^ 1][lane]; }
                else vnext_f = S.v[i0 + R][lane];
            }
        }
        OutT &O = outs[pc];
        // statistics of the tile being finalised: interior tiles without a mask accumulate inside the
        // row loop (fills the issue slots the scan chain leaves free); others go through ystat[]
        bool stats_fast = false;
        int64_t s_col = 0;
        bool s_col_ok = false;
        if (P2 && HAS_STATS && !(p.debug & 1)) {
            s_col = (int64_t)cu_strip * W + lane;
            s_col_ok = s_col < p.B;
            const bool strip_full = (int64_t)(cu_strip + 1) * W <= p.B;
            stats_fast = p.valid == nullptr && strip_full && cu_t0 >= 0 && cu_t0 + TC <= T;
            if (!st_has) {  // pivot: any value of the data's magnitude will do (uniform within the warp)
                constexpr int u0 = REVERSE ? R - 1 : 0;
                float y = to_f32(fma(C[u0], cin, A[u0]));
                if (MODE == 1) y = __fsub_rn(y, vreg[u0]);
                st_piv = __shfl_sync(0xffffffffu, y, 0);
                st_has = true;
            }
        }
        ACC a = (ACC)0, c = (ACC)1;
        float xa[R];  // resident: this warp's R advantages of the tile, bound for TMEM (dead otherwise)
        (void)xa;
        auto rows = [&](auto fast_tag) {
            constexpr bool FAST = decltype(fast_tag)::value;
            float ystat[(HAS_STATS && !FAST) ? R : 1];
            (void)ystat;
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
            for (int uu = 0; uu < R; ++uu) {
                const int u = REVERSE ? R - 1 - uu : uu;
                if (P2) {  // finalise row u of the previous tile from registers
                    const float x = to_f32(fma(C[u], cin, A[u]));
                    float y0;
                    if (MODE == 0) {
                        y0 = x;
                        if (!RESIDENT) O.o[0][i0 + u][lane] = x;
                        O.o[RET_SLOT][i0 + u][lane] = __fadd_rn(x, vreg[u]);
                    } else if (MODE == 1) {
                        y0 = __fsub_rn(x, vreg[u]);
                        O.o[RET_SLOT][i0 + u][lane] = x;
                        if (!RESIDENT) O.o[0][i0 + u][lane] = y0;
                    } else {
                        y0 = x;
                        O.o[0][i0 + u][lane] = x;
                    }
                    if (RESIDENT) xa[u] = y0;
                    if (HAS_PRE) O.o[PRE_SLOT][i0 + u][lane] = rpre[u];
                    if (HAS_STATS) {
                        if (FAST) {
                            const float d = y0 - st_piv;
                            if (uu & 1) { s1b += d; s2b = fmaf(d, d, s2b); }
                            else { s1a += d; s2a = fmaf(d, d, s2a); }
                        } else {
                            ystat[u] = y0;
                        }
                    }
                }
                if (P1) {  // scan row u of the next tile
                    float rf = S.r[i0 + u][lane];
                    bool dn = S.d[i0 + u][lane] != 0;
                    if (HAS_PRE) {
                        rf = __double2float_rn((double)rf * inv_denom);
                        rf = fminf(fmaxf(rf, clip_lo), clip_hi);
                        rpre[u] = rf;
                    }
                    if (REVERSE) {
                        const float vf = S.v[i0 + u][lane];
                        vreg[u] = vf;
                        const ACC cc = dn ? (ACC)0 : (MODE == 0 ? gl : gamma);
                        if (MODE == 0) {
#ifdef PLB_SCAN_F64_DELTA
                            const ACC delta = fma(gamma, (ACC)(dn ? 0.f : vnext_f), (ACC)rf) - (ACC)vf;
#else
                            // delta in float32 (one rounding of ~6e-8 relative per row, far inside the 1e-5 parity
                            // tolerance once discounted and summed), then ONE conversion for the float64 recurrence:
                            // saves two F2F and two FP64 instructions per row on the scan's busiest pipe
                            const ACC delta = (ACC)__fsub_rn(fmaf(gamma_f, dn ? 0.f : vnext_f, rf), vf);
#endif
                            a = fma(cc, a, delta);
                            vnext_f = vf;
                        } else {
                            a = fma(cc, a, (ACC)rf);
                        }
                        c = cc * c;
                    } else {
                        if (uu == 0 && warp == 0 && nx_k == 0) dn = first_prev_done;
                        const ACC cc = dn ? (ACC)0 : gamma;
                        a = fma(cc, a, (ACC)rf);
                        c = cc * c;
                    }
                    A[u] = a;
                    C[u] = c;
                }
            }
            if (P2 && HAS_STATS && !(p.debug & 1)) {
                if (FAST) {
                    st_n += R;
                } else {
                    const int u_lo = (int)((cu_t0 + i0 < 0) ? -(cu_t0 + i0) : 0);
                    const int u_hi = (int)((T - (cu_t0 + i0) < R) ? (T - (cu_t0 + i0)) : R);
                    const uint8_t *vp = (p.valid != nullptr) ? p.valid + (cu_t0 + i0) * p.ld_valid + s_col : nullptr;
#pragma unroll
                    for (int u = 0; u < R; ++u) {
                        const float d = ystat[u] - st_piv;
                        bool ok = s_col_ok && u >= u_lo && u < u_hi;
                        if (ok && vp != nullptr) ok = vp[(int64_t)u * p.ld_valid] != 0;
                        if (ok) {
                            s1a += d;
                            s2a = fmaf(d, d, s2a);
                            ++st_n;
                        }
                    }
                }
                st_s1 += (double)(s1a + s1b);
                st_s2 += (double)(s2a + s2b);
            }
        };
        if constexpr (P2 && HAS_STATS) {
            if (stats_fast) rows(std::true_type{});
            else rows(std::false_type{});
        } else {
            rows(std::false_type{});
        }
        if (P2 && RESIDENT) tmem_st<R>(tmem_warp + (uint32_t)(cu_idx * 2 * R), xa);
        if (P1) {
            sh->aggA[par][warp][lane] = a;
            sh->aggC[par][warp][lane] = c;
            if (MODE == 0 && warp == 0) sh->halo[par][lane] = vreg[0];
        }
        if (MODE == 2 && P2 && p.carry_ret_out != nullptr && cu_k == n_chunks - 1) {
            // the warp that holds t = T-1 hands the strip's last row to the caller's carry buffers
            const int64_t r_last = T - 1 - cu_t0;
            const int64_t ccol = (int64_t)cu_strip * W + lane;
            if (r_last >= i0 && r_last < i0 + R && ccol < p.B) {
                p.carry_ret_out[ccol] = O.o[0][r_last][lane];
                p.carry_done_out[ccol] = p.done_raw[(T - 1) * p.ld_done_raw + ccol];
            }
        }
        if (P2) fence_proxy_async_smem();  // staging-tile writes -> visible to the TMA store issued after the next barrier

        // ---- advance the cursors ---------------------------------------------------------
        if (P1) {
            cu_strip = nx_strip; cu_k = nx_k; cu_carry0 = nx_carry0;
            ++cu_idx;
            if (++nx_k == n_chunks) { nx_k = 0; nx_strip += gridDim.x; }
            par ^= 1;
            if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
    };

    if (n_tiles > 0) {
        body(std::false_type{}, std::true_type{});
        for (int j = 0; j + 1 < n_tiles; ++j) body(std::true_type{}, std::true_type{});
        body(std::true_type{}, std::false_type{});
    }

    __syncthreads();
    // after the last body: the final tile sits in outs[par ^ 1]
    if (threadIdx.x == 0 && st_pending) store_tile(par ^ 1);
    if constexpr (RESIDENT) {
        // ---- statistics of the whole batch: publish this CTA's sums, collect everybody's -------------
        __shared__ double s_pub[GX_RECORD_VALS];
        __shared__ float s_mean, s_denom;
        const uint32_t epoch = s_epoch;
        if (p.dbg != nullptr && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 1] = global_timer_ns();
        Sums acc;
        acc.n = (double)st_n; acc.s1 = st_s1; acc.s2 = st_s2; acc.piv = (double)st_piv;
        const Sums blk = block_reduce_sums_warp_pivot(acc, red_scratch);
        if (threadIdx.x == 0) { s_pub[0] = blk.n; s_pub[1] = blk.s1; s_pub[2] = blk.s2; s_pub[3] = blk.piv; }
        __syncthreads();
        gx_publish(p.gx, epoch, s_pub, s_peers);
        if (p.dbg != nullptr && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 2] = global_timer_ns();
        // one thread waits for the arrival hint (cheap to poll), then everybody reads its records in one batch
        if (p.poll_mode == 0 && threadIdx.x == 0) gx_wait_arrivals(p.gx, epoch);
        __syncthreads();
        const Moments all = gx_collect(p.gx, epoch, red_scratch);
        if (threadIdx.x == 0) {
            if (p.dbg != nullptr) p.dbg[blockIdx.x * 8 + 3] = global_timer_ns();
            s_mean = __double2float_rn(all.mean);
            s_denom = __fadd_rn(__double2float_rn(sqrt(all.m2 / all.n)), p.adv_eps);
            if (blockIdx.x == 0) {
                p.adv_moments_out[0] = all.n; p.adv_moments_out[1] = all.mean; p.adv_moments_out[2] = all.m2;
            }
            gx_retire(p.gx, epoch);
        }
        __syncthreads();
        const float mean = s_mean, denom = s_denom, rdenom = __frcp_rn(s_denom);
        // ---- advantage = (adv - mean) / (std + eps), float32 (norm_advantage.py:51-54), stored once --
        tmem_wait_st();
        // The input ring is idle now (every requested tile has been consumed): it becomes `n_out_bufs` whole-tile
        // staging buffers.  Per tile: all warps write their slabs (from tensor memory, normalised), one barrier,
        // one tile-sized TMA store (the TMA clips the columns beyond B; the tile that starts before t = 0 goes
        // through the ragged map, as in the scan) -- so the drain of tile j overlaps the arithmetic of tile j+1,
        // whose tensor-memory load is already in flight.
        float *bufs = reinterpret_cast<float *>(smem_raw);
        const int n_bufs = p.n_out_bufs < 9 ? p.n_out_bufs : 9;
        float x[R], xn[R];
        if (n_tiles > 0) tmem_ld<R>(tmem_warp, x);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) tmem_ld<R, false>(tmem_warp + (uint32_t)((j + 1) * 2 * R), xn);
            const int strip = (int)blockIdx.x + (j / n_chunks) * (int)gridDim.x;
            const int k = j - (j / n_chunks) * n_chunks;
            const int t0 = (int)T - (k + 1) * TC;
            float *tile = bufs + (size_t)(j % n_bufs) * TC * W;
            if (!(p.debug & 4)) {
                if (j >= n_bufs) {  // the store that last used this buffer (n_bufs tiles ago) must have read it
                    if (threadIdx.x == 0) tma_store_wait_read_dyn(n_bufs - 1);
                    __syncthreads();
                }
                float *slab = tile + (size_t)i0 * W;
#pragma unroll
                for (int u = 0; u < R; ++u) slab[u * W + lane] = div_by(__fsub_rn(x[u], mean), denom, rdenom);
                fence_proxy_async_smem();
                __syncthreads();
                if (threadIdx.x == 0) {
                    if (t0 < 0) tma_store_2d(&tm.o0_rag, tile + (size_t)(-t0) * W, strip * W, 0);
                    else tma_store_2d(&tm.o0, tile, strip * W, t0);
                    tma_store_commit();
                }
            } else {  // experiment switch: plain coalesced stores straight from registers
                const int64_t row0 = (int64_t)t0 + i0;
                const int64_t col = (int64_t)strip * W + lane;
                if (col < p.B) {
#pragma unroll
                    for (int u = 0; u < R; ++u)
                        if (row0 + u >= 0)
                            p.out0_raw[(row0 + u) * p.ld_out0_raw + col] = div_by(__fsub_rn(x[u], mean), denom, rdenom);
                }
            }
            if (j + 1 < n_tiles) {
                tmem_wait_ld_regs<R>(xn);
#pragma unroll
                for (int u = 0; u < R; ++u) x[u] = xn[u];
            }
        }
        if (p.dbg != nullptr && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 4] = global_timer_ns();
        if (threadIdx.x == 0) tma_store_wait<0>();  // shared memory must outlive the bulk reads
        if (p.dbg != nullptr && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 5] = global_timer_ns();
        tmem_fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            __syncwarp();
            tmem_dealloc(s_tmem_base, (uint32_t)p.tmem_cols);
        }
    } else if (HAS_STATS && !(p.debug & 2)) {
        Sums acc;
        acc.n = (double)st_n; acc.s1 = st_s1; acc.s2 = st_s2; acc.piv = (double)st_piv;
        XchPre xch_pre = xch_no_preload();
        if (xch_preloaded && warp == 0) {  // written by this very warp
            xch_pre.mine = s_xch_mine; xch_pre.peer = s_xch_peer[lane]; xch_pre.epoch = s_xch_epoch; xch_pre.valid = true;
        }
        moments_tail<true>(acc, p.tail, red_scratch, xch_pre);
    }
    if (threadIdx.x == 0) tma_store_wait<0>();
}

}  // namespace chunked

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D row-major [rows, cols] tensor map with a [box_rows x box_cols] box, zero OOB fill.
int make_tensor_map_2d(CUtensorMap *tm, const void *base, int elem_bytes, int64_t rows, int64_t cols,
                       int64_t ld_elems, int box_rows, int box_cols)
{
    // a descriptor depends only on these arguments: a small per-thread cache saves the driver call
    // (~1 us each, up to nine per scan launch) when the same leaves come back, as they do every step
    struct Key { const void *base; int64_t rows, cols, ld; int elem, box_rows, box_cols; };
    struct Entry { Key key; CUtensorMap map; bool used; };
    constexpr int CACHE = 64;
    static thread_local Entry cache[CACHE];
    static thread_local int next_slot = 0;
    const Key key = {base, rows, cols, ld_elems, elem_bytes, box_rows, box_cols};
    for (int i = 0; i < CACHE; ++i) {
        const Entry &e = cache[i];
        if (e.used && e.key.base == key.base && e.key.rows == key.rows && e.key.cols == key.cols && e.key.ld == key.ld &&
            e.key.elem == key.elem && e.key.box_rows == key.box_rows && e.key.box_cols == key.box_cols) {
            *tm = e.map;
            return PLB_OK;
        }
    }
    EncodeTiledFn fn = get_encode_fn();
    PLB_REQUIRE(fn != nullptr, PLB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld_elems * (cuuint64_t)elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    const CUresult r = fn(tm, dt, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PLB_REQUIRE(r == CUDA_SUCCESS, PLB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    Entry &slot = cache[next_slot];
    next_slot = (next_slot + 1) % CACHE;
    slot.key = key; slot.map = *tm; slot.used = true;
    return PLB_OK;
}

static bool tma_ok_f32(const float *p, int64_t ld) { return aligned(p, 16) && (ld % 4 == 0); }
static bool tma_ok_u8(const uint8_t *p, int64_t ld) { return aligned(p, 16) && (ld % 16 == 0); }

bool scan_chunked_supported(const float *reward, int64_t ld_r, const float *value, int64_t ld_v,
                            const uint8_t *done, int64_t ld_d, int64_t T, int64_t B, int64_t inner)
{
    if (inner != 1) return false;
    if (T > (int64_t)INT32_MAX - 1024 || B > (int64_t)INT32_MAX - 1024) return false;
    if (!tma_ok_f32(reward, ld_r)) return false;
    if (value != nullptr && !tma_ok_f32(value, ld_v)) return false;
    // the forward scan reads done one row early: base - ld must keep 16-byte alignment (ld % 16 == 0 does)
    return tma_ok_u8(done, ld_d);
}

// tile geometry: 0 = auto, 1 = 8 warps x 8 rows (64-row tiles), 2 = 8 warps x 16 rows (128-row tiles)
static int g_chunked_geom = 0;
static int g_scan_debug = 0;
static int g_scan_max_stages = 0;     // experiments: cap on the input ring depth (plb_set_tuning key 9)
// resident (tensor-memory) advantage kernel: 0 never, 1 auto (when the batch is sharded over several GPUs: there
// it replaces a kernel boundary plus a cross-GPU exchange; on one GPU the two-kernel route, whose second pass
// is served by L2, measured faster: 20.5 vs 22-24 us at C1), 2 wherever the batch fits (plb_set_tuning key 12)
static int g_resident = 1;
static int g_resident_plain_launch = 0;  // experiments: 1 = no cooperative-launch attribute (key 14)
static int g_resident_poll = 0;          // experiments: 1 = every thread polls its own records (key 15)
static int g_sharded_route = 0;          // sharded step: 0 / 3 one record per rank (default), 1 resident kernel (key 17)
int sharded_route() { return g_sharded_route; }
static unsigned long long *g_debug_buf = nullptr;  // plb_debug_set_buffer
unsigned long long *debug_buffer() { return g_debug_buf; }

template <int MODE, typename ACC, int NW, int R, bool HAS_PRE, bool HAS_STATS>
static int launch_chunked_inst(const chunked::TensorMaps &tm, chunked::Params p, int sms, cudaStream_t stream)
{
    using namespace chunked;
    constexpr int TC = NW * R;
    constexpr int N_OUT = (MODE != 2 ? 2 : 1) + (HAS_PRE ? 1 : 0);
    const size_t stage_bytes = sizeof(Stage<TC>);
    const size_t fixed_bytes = sizeof(Shared<ACC, NW>) + 2 * sizeof(OutTile<TC, N_OUT>);
    p.n_strips = (int)ceil_div(p.B, W);
    p.debug = g_scan_debug;
    p.n_chunks = (int)ceil_div(p.T, TC);
    // few strips: one CTA per SM with a deep ring; many strips: several CTAs per SM, shallow rings.
    // 228 KB of shared memory per SM, 1 KB reserved per CTA, ~1 KB of static scratch per CTA.
    int ctas_per_sm = p.n_strips <= sms ? 1 : (p.n_strips <= 2 * sms ? 2 : 3);
    if (NW > 8) ctas_per_sm = 1;
    auto kern = scan_pipelined_kernel<MODE, ACC, NW, R, HAS_PRE, HAS_STATS>;
    static thread_local size_t static_smem = ~(size_t)0;
    if (static_smem == ~(size_t)0) {
        cudaFuncAttributes attr;
        PLB_CUDA(cudaFuncGetAttributes(&attr, kern));
        static_smem = attr.sharedSizeBytes;
    }
    const size_t smem_cap = 227 * 1024 - static_smem;  // per-CTA limit: static + dynamic <= 227 KB
    int stages = 0;
    for (; ctas_per_sm >= 1; --ctas_per_sm) {
        size_t budget = (size_t)(228 * 1024) / ctas_per_sm - 1024 - static_smem;  // 1 KB reserved per CTA
        if (budget > smem_cap) budget = smem_cap;
        stages = budget > fixed_bytes ? (int)((budget - fixed_bytes) / stage_bytes) : 0;
        if (stages >= 2 || ctas_per_sm == 1) break;
    }
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (g_scan_max_stages > 0 && stages > g_scan_max_stages) stages = g_scan_max_stages;
    if (stages > p.n_chunks) stages = p.n_chunks;
    PLB_REQUIRE(stages >= 1, PLB_ERR_UNSUPPORTED, "chunked scan does not fit in shared memory");
    p.n_stages = stages;
    const size_t smem = stage_bytes * stages + fixed_bytes;
    const int grid = p.n_strips < sms * ctas_per_sm ? p.n_strips : sms * ctas_per_sm;
    static thread_local bool configured = false;
    if (!configured) {
        PLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        configured = true;
    }
    PLB_CUDA(launch_pdl(kern, dim3(grid), dim3(NW * 32), smem, stream, tm, p));
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

// Geometry of the resident variant: one persistent CTA per SM, every CTA's advantage tiles parked in
// its SM's tensor memory (512 columns = 16 tiles of 128 rows x 32 columns, or 32 tiles of 64 rows).
struct ResidentGeometry {
    int R;          // rows per warp: 16 (128-row tiles) or 8 (64-row tiles, T <= 64)
    int grid, n_chunks, tiles_per_cta, tmem_cols;
};
static bool resident_geometry(int64_t T, int64_t B, int sms, ResidentGeometry *g)
{
    using namespace chunked;
    if (sms <= 0 || sms > GX_MAX_SLOTS) return false;
    const int64_t strips = ceil_div(B, W);
    g->R = T <= 64 ? 8 : 16;
    const int TC = 8 * g->R;
    g->grid = (int)(strips < sms ? strips : sms);
    const int64_t chunks = ceil_div(T, TC);
    const int64_t tiles = ceil_div(strips, g->grid) * chunks;
    const int64_t cols = tiles * 2 * g->R;
    if (cols > 512) return false;
    g->n_chunks = (int)chunks;
    g->tiles_per_cta = (int)tiles;
    int c = 32;
    while (c < cols) c <<= 1;
    g->tmem_cols = c;
    return true;
}

template <int MODE, int R, bool HAS_PRE>
static int launch_resident_inst(const chunked::TensorMaps &tm, chunked::Params p, const ResidentGeometry &geo,
                                cudaStream_t stream)
{
    using namespace chunked;
    constexpr int NW = 8, TC = NW * R;
    constexpr int N_STG = 1 + (HAS_PRE ? 1 : 0);
    auto kern = scan_pipelined_kernel<MODE, double, NW, R, HAS_PRE, true, true>;
    const size_t stage_bytes = sizeof(Stage<TC>);
    const size_t fixed_bytes = sizeof(Shared<double, NW>) + 2 * sizeof(OutTile<TC, N_STG>);
    static thread_local size_t static_smem = ~(size_t)0;
    if (static_smem == ~(size_t)0) {
        cudaFuncAttributes attr;
        PLB_CUDA(cudaFuncGetAttributes(&attr, kern));
        static_smem = attr.sharedSizeBytes;
        PLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024 - static_smem)));
    }
    const size_t smem_cap = 227 * 1024 - static_smem;
    int stages = (int)((smem_cap - fixed_bytes) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (g_scan_max_stages > 0 && stages > g_scan_max_stages) stages = g_scan_max_stages;
    // (no cap at the number of chunks: the final store phase reuses the ring as staging slabs)
    PLB_REQUIRE(stages >= 1, PLB_ERR_UNSUPPORTED, "resident scan does not fit in shared memory");
    p.n_stages = stages;
    p.n_strips = (int)ceil_div(p.B, W);
    p.n_chunks = geo.n_chunks;
    p.debug = g_scan_debug;
    p.tmem_cols = geo.tmem_cols;
    const size_t tile_bytes = (size_t)TC * W * sizeof(float);
    p.n_out_bufs = (int)(stage_bytes * stages / tile_bytes);  // whole-tile staging buffers the idle ring holds
    PLB_REQUIRE(p.n_out_bufs >= 1, PLB_ERR_UNSUPPORTED, "resident scan: ring too small for the store phase");
    const size_t smem = stage_bytes * stages + fixed_bytes;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)geo.grid);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // every CTA waits for every other: the grid must be co-resident
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_resident_plain_launch ? 0 : 1;
    p.dbg = g_debug_buf;
    p.poll_mode = g_resident_poll;
    PLB_CUDA(cudaLaunchKernelEx(&cfg, kern, tm, p));
    count_launch(1);
    return PLB_OK;
}

struct ChunkedIO {
    const float *reward; int64_t ld_r;
    const float *value; int64_t ld_v;
    const uint8_t *done; int64_t ld_d;
    float *out0; int64_t ld0;
    float *out1; int64_t ld1;
    float *reward_out; int64_t ld_ro;
};

// tensor maps of one scan launch with TC-row tiles (warp_rows > 0: also out0 with a warp-slab box)
template <int MODE>
static int build_tensor_maps(const ChunkedIO &io, const chunked::Params &p, bool has_pre, int TC, int warp_rows,
                             chunked::TensorMaps *out)
{
    using namespace chunked;
    TensorMaps &tm = *out;
    int rc = make_tensor_map_2d(&tm.r, io.reward, 4, p.T, p.B, io.ld_r, TC, W);
    if (rc) return rc;
    if (MODE != 2) {
        rc = make_tensor_map_2d(&tm.v, io.value, 4, p.T, p.B, io.ld_v, TC, W);
        if (rc) return rc;
    } else {
        tm.v = tm.r;
    }
    rc = make_tensor_map_2d(&tm.d, io.done, 1, p.T, p.B, io.ld_d, TC, W);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tm.o0, io.out0, 4, p.T, p.B, io.ld0, TC, W);
    if (rc) return rc;
    if (MODE != 2) {
        rc = make_tensor_map_2d(&tm.o1, io.out1, 4, p.T, p.B, io.ld1, TC, W);
        if (rc) return rc;
    } else {
        tm.o1 = tm.o0;
    }
    if (has_pre) {
        rc = make_tensor_map_2d(&tm.o2, io.reward_out, 4, p.T, p.B, io.ld_ro, TC, W);
        if (rc) return rc;
    } else {
        tm.o2 = tm.o0;
    }
    tm.o0_rag = tm.o0; tm.o1_rag = tm.o1; tm.o2_rag = tm.o2;
    const int ragged = (int)(p.T % TC);
    if (MODE != 2 && ragged != 0) {
        rc = make_tensor_map_2d(&tm.o0_rag, io.out0, 4, p.T, p.B, io.ld0, ragged, W);
        if (rc) return rc;
        rc = make_tensor_map_2d(&tm.o1_rag, io.out1, 4, p.T, p.B, io.ld1, ragged, W);
        if (rc) return rc;
        if (has_pre) {
            rc = make_tensor_map_2d(&tm.o2_rag, io.reward_out, 4, p.T, p.B, io.ld_ro, ragged, W);
            if (rc) return rc;
        }
    }
    (void)warp_rows;
    return PLB_OK;
}

template <int MODE>
static int launch_chunked(const ChunkedIO &io, chunked::Params p, bool has_pre, cudaStream_t stream)
{
    using namespace chunked;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    // one CTA per SM (few strips): 8 warps x 16 rows = 128-row tiles (measured fastest at C1: 13.8 us
    // vs 15.2 us for 16 warps x 8 rows and 18.3 us for 64-row tiles); several CTAs per SM (many
    // strips): 8 warps x 8 rows = 64-row tiles.
    int geom = g_chunked_geom;
    if (geom == 0) {
        // few strips (<= one per SM) or many long strips (>= 4 per SM, T >= 256): one persistent CTA per SM
        // with 128-row tiles (C4: 49.8 us vs 55.0 us for 64-row tiles); in between: several CTAs per SM
        const int64_t strips = ceil_div(p.B, W);
        geom = (strips <= sms || (strips >= 4 * (int64_t)sms && p.T >= 256)) ? 2 : 1;
    }
    if (p.T <= 64) geom = 1;
    const int TC = geom == 1 ? 64 : 128;
    TensorMaps tm;
    const int rc = build_tensor_maps<MODE>(io, p, has_pre, TC, 0, &tm);
    if (rc) return rc;
    const bool stats = p.tail.moments_out != nullptr;
#define PLB_DISPATCH(ACC_, NW_, R_)                                                                          \
    do {                                                                                                     \
        if constexpr (MODE != 2) {                                                                           \
            if (has_pre && stats) return launch_chunked_inst<MODE, ACC_, NW_, R_, true, true>(tm, p, sms, stream);   \
            if (has_pre) return launch_chunked_inst<MODE, ACC_, NW_, R_, true, false>(tm, p, sms, stream);           \
        }                                                                                                    \
        if (stats) return launch_chunked_inst<MODE, ACC_, NW_, R_, false, true>(tm, p, sms, stream);         \
        return launch_chunked_inst<MODE, ACC_, NW_, R_, false, false>(tm, p, sms, stream);                   \
    } while (0)
    if (geom == 1) PLB_DISPATCH(double, 8, 8);
    PLB_DISPATCH(double, 8, 16);
#undef PLB_DISPATCH
}

// reverse scan + NormalizeAdvantage in one cooperative launch (MODE 0 / 1)
template <int MODE>
static int launch_resident(const ChunkedIO &io, chunked::Params p, bool has_pre, cudaStream_t stream)
{
    using namespace chunked;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    ResidentGeometry geo;
    PLB_REQUIRE(resident_geometry(p.T, p.B, sms, &geo), PLB_ERR_UNSUPPORTED,
                "resident scan: the advantage of this batch does not fit in tensor memory");
    PLB_REQUIRE(p.gx.peers != nullptr && p.gx.self != nullptr && p.gx.world >= 1 && p.gx.world <= GX_MAX_WORLD && p.gx.rank >= 0 &&
                    p.gx.rank < p.gx.world && p.gx.slots >= geo.grid && p.gx.slots <= GX_MAX_SLOTS,
                PLB_ERR_INVALID_ARGUMENT, "resident scan: bad grid-exchange description (world %d, slots %d, grid %d)",
                p.gx.world, p.gx.slots, geo.grid);
    TensorMaps tm;
    const int rc = build_tensor_maps<MODE>(io, p, has_pre, 8 * geo.R, geo.R, &tm);
    if (rc) return rc;
    p.out0_raw = io.out0; p.ld_out0_raw = io.ld0;
    if (geo.R == 16) {
        if (has_pre) return launch_resident_inst<MODE, 16, true>(tm, p, geo, stream);
        return launch_resident_inst<MODE, 16, false>(tm, p, geo, stream);
    }
    if (has_pre) return launch_resident_inst<MODE, 8, true>(tm, p, geo, stream);
    return launch_resident_inst<MODE, 8, false>(tm, p, geo, stream);
}

bool resident_wanted(int world) { return g_resident == 2 || (g_resident == 1 && world > 1); }  // one GPU: only when forced

bool resident_supported(const float *reward, int64_t ld_r, const float *value, int64_t ld_v, const uint8_t *done,
                        int64_t ld_d, const float *adv, int64_t ld_a, const float *ret, int64_t ld_t,
                        int64_t T, int64_t B)
{
    if (T <= 0 || B <= 0) return false;
    if (!scan_chunked_supported(reward, ld_r, value, ld_v, done, ld_d, T, B, 1)) return false;
    if (!tma_ok_f32(adv, ld_a) || !tma_ok_f32(ret, ld_t)) return false;
    ResidentGeometry geo;
    return resident_geometry(T, B, sm_count(), &geo);
}

int scan_reverse(int mode, const float *reward, int64_t ld_r, const float *value, int64_t ld_v,
                 const uint8_t *done, int64_t ld_d, const float *boot,
                 float *adv, int64_t ld_a, float *ret, int64_t ld_t,
                 int64_t T, int64_t B, int64_t inner, double gamma, double lambda, int algo,
                 const ScanFusion *fusion, cudaStream_t stream)
{
    PLB_REQUIRE(T >= 0 && B >= 0 && inner >= 1, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (T == 0 || B == 0) return PLB_OK;
    PLB_REQUIRE(reward && value && done && boot && adv && ret, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(ld_r >= B && ld_d >= B && ld_v >= B * inner && ld_a >= B * inner && ld_t >= B * inner,
                PLB_ERR_INVALID_ARGUMENT, "row stride smaller than row length");
    bool can_chunk = scan_chunked_supported(reward, ld_r, value, ld_v, done, ld_d, T, B, inner) &&
                     tma_ok_f32(adv, ld_a) && tma_ok_f32(ret, ld_t);
    if (fusion != nullptr && fusion->has_pre && fusion->reward_out != nullptr)
        can_chunk = can_chunk && tma_ok_f32(fusion->reward_out, fusion->ld_reward_out);
    if (algo == PLB_ALGO_AUTO) {
        // the chunked scan is the faster one at every measured shape (C1: 13.5 us; C4: 49.8 us vs 53.8 us
        // for the serial walk); PLB_ALGO_SERIAL stays available as the bit-exact (reference rounding) mode
        algo = can_chunk ? PLB_ALGO_CHUNKED : PLB_ALGO_SERIAL;
    }
    if (algo == PLB_ALGO_CHUNKED) {
        PLB_REQUIRE(can_chunk, PLB_ERR_UNSUPPORTED,
                    "chunked scan needs inner == 1, 16-byte aligned bases, ld %% 4 == 0 (float) and ld %% 16 == 0 (done)");
        chunked::Params p = {};
        p.boot = boot;
        ChunkedIO io = {reward, ld_r, value, ld_v, done, ld_d, adv, ld_a, ret, ld_t, nullptr, 0};
        p.T = T; p.B = B;
        p.gamma = gamma; p.lambda = lambda;
        p.clip_lo = -INFINITY; p.clip_hi = INFINITY;
        bool has_pre = false;
        if (fusion != nullptr) {
            has_pre = fusion->has_pre != 0;
            p.reward_var = fusion->reward_var;
            p.reward_eps = fusion->reward_eps;
            if (!(fusion->clip_lo != fusion->clip_lo)) p.clip_lo = (float)fusion->clip_lo;
            if (!(fusion->clip_hi != fusion->clip_hi)) p.clip_hi = (float)fusion->clip_hi;
            if (has_pre) {
                PLB_REQUIRE(fusion->reward_out != nullptr, PLB_ERR_INVALID_ARGUMENT,
                            "fused reward transform needs reward_out (it may alias reward)");
                io.reward_out = fusion->reward_out; io.ld_ro = fusion->ld_reward_out;
            }
            p.valid = fusion->valid; p.ld_valid = fusion->ld_valid;
            p.tail = fusion->tail;
            if (fusion->resident) {
                PLB_REQUIRE(resident_supported(reward, ld_r, value, ld_v, done, ld_d, adv, ld_a, ret, ld_t, T, B),
                            PLB_ERR_UNSUPPORTED, "resident scan not available for this shape");
                PLB_REQUIRE(fusion->adv_moments_out != nullptr, PLB_ERR_INVALID_ARGUMENT, "adv_moments_out is null");
                p.gx = fusion->gx;
                p.adv_eps = (float)fusion->adv_eps;
                p.adv_moments_out = fusion->adv_moments_out;
                return mode == 0 ? launch_resident<0>(io, p, has_pre, stream) : launch_resident<1>(io, p, has_pre, stream);
            }
        }
        return mode == 0 ? launch_chunked<0>(io, p, has_pre, stream) : launch_chunked<1>(io, p, has_pre, stream);
    }
    PLB_REQUIRE(algo == PLB_ALGO_SERIAL, PLB_ERR_INVALID_ARGUMENT, "unknown algo %d", algo);
    PLB_REQUIRE(fusion == nullptr, PLB_ERR_UNSUPPORTED, "fused transforms need the chunked scan");
    const int64_t ncols = B * inner;
    const int threads = 128;
    const int64_t blocks = ceil_div(ncols, threads);
    if (mode == 0)
        scan_reverse_serial_kernel<0><<<(unsigned)blocks, threads, 0, stream>>>(
            reward, ld_r, value, ld_v, done, ld_d, boot, adv, ld_a, ret, ld_t, T, ncols, inner, gamma, lambda);
    else
        scan_reverse_serial_kernel<1><<<(unsigned)blocks, threads, 0, stream>>>(
            reward, ld_r, value, ld_v, done, ld_d, boot, adv, ld_a, ret, ld_t, T, ncols, inner, gamma, lambda);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

// carry rows for callers without padding rows, when the scan kernel did not write them itself
static int copy_carry_rows(const ScanFusion *fusion, const float *out, int64_t ld_o, const uint8_t *done, int64_t ld_d,
                           int64_t T, int64_t B, cudaStream_t stream)
{
    if (fusion == nullptr || fusion->carry_ret_out == nullptr || fusion->carry_done_out == nullptr) return PLB_OK;
    PLB_CUDA(cudaMemcpyAsync(fusion->carry_ret_out, out + (T - 1) * ld_o, (size_t)B * sizeof(float),
                             cudaMemcpyDeviceToDevice, stream));
    PLB_CUDA(cudaMemcpyAsync(fusion->carry_done_out, done + (T - 1) * ld_d, (size_t)B, cudaMemcpyDeviceToDevice, stream));
    return PLB_OK;
}

int scan_forward(const float *reward, int64_t ld_r, const uint8_t *done, int64_t ld_d,
                 const float *prev_ret, const uint8_t *prev_done, float *out, int64_t ld_o,
                 int64_t T, int64_t B, double gamma, int algo, const ScanFusion *fusion, cudaStream_t stream)
{
    PLB_REQUIRE(T >= 0 && B >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (T == 0 || B == 0) return PLB_OK;
    PLB_REQUIRE(reward && done && prev_ret && prev_done && out, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(ld_r >= B && ld_d >= B && ld_o >= B, PLB_ERR_INVALID_ARGUMENT, "row stride smaller than row length");
    const bool can_chunk = scan_chunked_supported(reward, ld_r, nullptr, 0, done, ld_d, T, B, 1) &&
                           tma_ok_f32(out, ld_o);
    if (algo == PLB_ALGO_AUTO) {
        algo = can_chunk ? PLB_ALGO_CHUNKED : PLB_ALGO_SERIAL;
    }
    if (algo == PLB_ALGO_CHUNKED) {
        PLB_REQUIRE(can_chunk, PLB_ERR_UNSUPPORTED,
                    "chunked scan needs 16-byte aligned bases, ld %% 4 == 0 (float) and ld %% 16 == 0 (done)");
        chunked::Params p = {};
        p.prev_ret = prev_ret; p.prev_done = prev_done;
        ChunkedIO io = {reward, ld_r, nullptr, 0, done, ld_d, out, ld_o, nullptr, 0, nullptr, 0};
        p.T = T; p.B = B;
        p.gamma = gamma; p.lambda = 1.0;
        p.clip_lo = -INFINITY; p.clip_hi = INFINITY;
        bool carry_in_kernel = false;
        if (fusion != nullptr) {
            p.valid = fusion->valid; p.ld_valid = fusion->ld_valid;
            p.tail = fusion->tail;
            if (fusion->carry_ret_out != nullptr && fusion->carry_done_out != nullptr) {
                p.carry_ret_out = fusion->carry_ret_out; p.carry_done_out = fusion->carry_done_out;
                p.done_raw = done; p.ld_done_raw = ld_d;
                carry_in_kernel = true;
            }
        }
        const int rc = launch_chunked<2>(io, p, false, stream);
        if (rc != PLB_OK || carry_in_kernel) return rc;
        return copy_carry_rows(fusion, out, ld_o, done, ld_d, T, B, stream);
    }
    PLB_REQUIRE(algo == PLB_ALGO_SERIAL, PLB_ERR_INVALID_ARGUMENT, "unknown algo %d", algo);
    PLB_REQUIRE(fusion == nullptr || fusion->tail.moments_out == nullptr, PLB_ERR_UNSUPPORTED,
                "fused statistics need the chunked scan");
    const int threads = 128;
    scan_forward_serial_kernel<<<(unsigned)ceil_div(B, threads), threads, 0, stream>>>(
        reward, ld_r, done, ld_d, prev_ret, prev_done, out, ld_o, T, B, gamma);
    PLB_LAUNCH_CHECK();
    return copy_carry_rows(fusion, out, ld_o, done, ld_d, T, B, stream);
}

}  // namespace plb

extern "C" {

int plb_set_tuning(int key, int value)
{
    switch (key) {
        case 0: plb::g_chunked_geom = value; return PLB_OK;
        case 12: plb::g_resident = value; return PLB_OK;
        case 13: plb::set_rng_cluster(value); return PLB_OK;
        case 16: plb::set_gather_mode(value); return PLB_OK;
        case 17: plb::g_sharded_route = value; return PLB_OK;
        case 14: plb::g_resident_plain_launch = value; return PLB_OK;
        case 15: plb::g_resident_poll = value; return PLB_OK;
        case 18: plb::set_pdl_enabled(value); return PLB_OK;
        case 20: plb::set_normadv_xch_threads(value); return PLB_OK;
        case 21: plb::set_normadv_reduce_mode(value); return PLB_OK;
        case 5: plb::g_scan_debug = value; return PLB_OK;
        case 9: plb::g_scan_max_stages = value; return PLB_OK;
        case 11: plb::set_normadv_ctas_per_sm(value); return PLB_OK;
        case 7: plb::set_obs_lsu_cols(value); return PLB_OK;
        case 3: plb::set_obs_variant(value); return PLB_OK;
        default: plb::set_error("unknown tuning key %d", key); return PLB_ERR_INVALID_ARGUMENT;
    }
}

void plb_debug_set_buffer(void *device_buffer) { plb::g_debug_buf = reinterpret_cast<unsigned long long *>(device_buffer); }

int plb_compute_gae_advantage_f32(const float *reward, int64_t ld_reward, const float *value, int64_t ld_value,
                                  const uint8_t *done, int64_t ld_done, const float *bootstrap_value,
                                  float *out_advantage, int64_t ld_advantage, float *out_return, int64_t ld_return,
                                  int64_t T, int64_t B, int64_t inner, double discount, double gae_lambda,
                                  int algo, plb_stream_t stream)
{
    return plb::scan_reverse(0, reward, ld_reward, value, ld_value, done, ld_done, bootstrap_value,
                             out_advantage, ld_advantage, out_return, ld_return, T, B, inner,
                             discount, gae_lambda, algo, nullptr, (cudaStream_t)stream);
}

int plb_compute_discount_return_f32(const float *reward, int64_t ld_reward, const float *value, int64_t ld_value,
                                    const uint8_t *done, int64_t ld_done, const float *bootstrap_value,
                                    float *out_advantage, int64_t ld_advantage, float *out_return, int64_t ld_return,
                                    int64_t T, int64_t B, int64_t inner, double discount,
                                    int algo, plb_stream_t stream)
{
    return plb::scan_reverse(1, reward, ld_reward, value, ld_value, done, ld_done, bootstrap_value,
                             out_advantage, ld_advantage, out_return, ld_return, T, B, inner,
                             discount, 1.0, algo, nullptr, (cudaStream_t)stream);
}

int plb_compute_past_discount_return_f32(const float *reward, int64_t ld_reward, const uint8_t *done, int64_t ld_done,
                                         const float *previous_return, const uint8_t *previous_done,
                                         float *out_return, int64_t ld_out, int64_t T, int64_t B,
                                         double discount, int algo, plb_stream_t stream)
{
    return plb::scan_forward(reward, ld_reward, done, ld_done, previous_return, previous_done,
                             out_return, ld_out, T, B, discount, algo, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
