/*
 * parllel_b200.h -- C ABI of libparllel_b200.so: the B200 (sm_100a) implementation of
 * parllel's batch post-processing hot path (Transform pipeline + ReplayBuffer gather).
 *
 * This is the drop-in boundary.  Each entry point replaces one compiled/numpy
 * routine of the reference (file:line relative to the upstream repository root is
 * cited at every declaration); INTEGRATION.md shows the ctypes binding a parllel
 * maintainer would add on the reference side.
 *
 * Conventions (all entry points):
 *   - plain C types only; every data pointer is a DEVICE pointer unless it says HOST;
 *   - time-major [T, B(, inner)] arrays, the column axis contiguous, one explicit row
 *     stride per array in ELEMENTS (`ld_*`), exactly the views parllel's Array hands
 *     to numba (parllel/arrays/array.py:420-431);
 *   - float arrays are float32, masks are 1-byte bool (0 / non-zero), statistics and
 *     scalar hyper-parameters are float64 -- the reference's dtypes;
 *   - asynchronous and stream-ordered on `stream` (a cudaStream_t passed as void*);
 *     no allocation, no host synchronisation, no host<->device copies inside;
 *   - the caller owns all memory, including `workspace` (size from the matching
 *     `*_workspace_bytes` query; must be zero-filled ONCE before its first use and is
 *     left ready for reuse by every call; one workspace must not be shared by calls
 *     that may run concurrently on different streams);
 *   - returns 0 on success, a positive cudaError_t, or a negative PLB_ERR_* code;
 *     `plb_last_error()` returns a thread-local description of the last failure.
 */
#ifndef PARLLEL_B200_H
#define PARLLEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PLB_API __attribute__((visibility("default")))
#else
#define PLB_API
#endif

#define PLB_VERSION 200 /* 0.2.0 */

#define PLB_OK 0
#define PLB_ERR_INVALID_ARGUMENT (-1)
#define PLB_ERR_UNSUPPORTED (-2)
#define PLB_ERR_WORKSPACE (-3)

/* scan algorithm selection */
#define PLB_ALGO_AUTO 0    /* chunked wherever the layout allows it (TMA alignment, inner == 1), else serial */
#define PLB_ALGO_SERIAL 1  /* one thread per column, float64 step + float32 store: the
                              reference's per-step rounding, bit-identical to oracle/plb_oracle.c */
#define PLB_ALGO_CHUNKED 2 /* TMA-staged chunked affine scan (float64 carries); |err| <= ~3e-6 */

/* flags for plb_update_from_moments_f64 */
#define PLB_MOMENTS_F32_FLOW 1 /* batch mean/var rounded to float32 and batch_var*count evaluated
                                  in float32, the numba typing of running_mean_std.py:52-53,27 */

/* flags for the normalisation entry points */
#define PLB_EPS_ON_VAR 0 /* x / sqrt(var + eps)   (norm_rewards.py:114, norm_obs.py:72-74) */
#define PLB_EPS_ON_STD 1 /* x / (std + eps)       (norm_advantage.py:54)                   */

typedef void *plb_stream_t; /* cudaStream_t */

PLB_API int plb_version(void);
PLB_API const char *plb_last_error(void);
/* Kernels this process has ENQUEUED through the library so far (launches recorded into a CUDA graph
 * under capture count once, at capture): out2[0] = all kernels, out2[1] = resident advantage kernels. */
PLB_API void plb_launch_counters(int64_t *out2);
/* cudaGraphLaunch of an instantiated graph (a cudaGraphExec_t, e.g. torch.cuda.CUDAGraph.raw_cuda_graph_exec()) on
 * `stream`: the replay fast path of the Python classes, a few microseconds cheaper per call than going through the
 * framework's replay wrapper -- with B sharded over 8 GPUs the slowest host paces every rank. */
PLB_API int plb_graph_launch(void *graph_exec, plb_stream_t stream);
/* Number of SMs of the current device (grid sizing); <0 on error. */
PLB_API int plb_sm_count(void);
/* Tuning / diagnostics knobs (process-wide; used by tools/ and the tests, not by the transforms).
 *   key 0: tile geometry of the chunked scan: 0 auto, 1 = 8 warps x 8 rows, 2 = 8 warps x 16 rows
 *   key 3: NormalizeObservations step kernel: 0 auto (4, then 1, then 2, then three kernels),
 *          1 shared-memory slab through TMA, 2 two-pass L2 slab,
 *          4 shared-memory slab through cp.async / direct stores
 *   key 5: experiments only: 1 skip the fused statistics, 2 skip their tail (results then incomplete),
 *          4 resident kernel: final advantage through plain stores instead of TMA stores
 *   key 7: slab width of observation variant 4: 8, 16 (default) or 32 columns
 *   key 9: cap on the input ring depth of the chunked scan (0 = as many stages as fit)
 *   key 11: grid cap of the streaming normalise kernels, CTAs of 256 threads per SM (default 4)
 *   key 12: resident (tensor-memory) advantage kernel: 0 never, 1 (default) when B is sharded over several GPUs,
 *           2 wherever the batch fits (also on one GPU)
 *   key 14 / 15: experiments on the resident kernel (1 = plain instead of cooperative launch / every thread polls)
 *   key 17: route of EstimateAdvantage + NormalizeAdvantage on a batch sharded over several GPUs: 0/3 last-CTA reduction +
 *           one record of partial sums per rank (default), 1 resident kernel
 *   key 18: programmatic dependent launch of the streaming kernels (scans, advantage normalise, index generation):
 *           1 (default) on, 0 off -- the kernels then start only when their predecessor on the stream has retired
 *   key 20: CTA size of the sharded advantage-normalise kernel: 256 (default), 512 or 1024 threads (one CTA per SM at 1024)
 *   key 21: experiments: 1 = the normalise prologue reads the scan's partial sums one after the other (round-2 code)
 *   key 16: 1 = replay gather through the load/store-unit path only (default 0: TMA-staged rows of >= 128 B)
 *   key 13: 0 = index generation always by the single-CTA kernel (default 1: 8-CTA cluster kernel)
 * The environment variable PLB_TUNING="key=value,..." applies them when the Python package loads
 * the library. */
#define PLB_TUNE_CHUNK_GEOMETRY 0
#define PLB_TUNE_OBS_VARIANT 3
#define PLB_TUNE_SCAN_DEBUG 5
#define PLB_TUNE_OBS_SLAB_COLS 7
#define PLB_TUNE_SCAN_MAX_STAGES 9
#define PLB_TUNE_NORMADV_CTAS_PER_SM 11
#define PLB_TUNE_RESIDENT 12
#define PLB_TUNE_RNG_CLUSTER 13
PLB_API int plb_set_tuning(int key, int value);
/* Diagnostics: with a DEVICE buffer of >= 8 * 8 * plb_sm_count() bytes set, every CTA of the resident
 * advantage kernel leaves %globaltimer stamps there ([cta][8]: start, scan done, published, collected,
 * stores issued, stores drained); NULL (default) switches it off. */
PLB_API void plb_debug_set_buffer(void *device_buffer);

/* ------------------------------------------------------------------------------------
 * EstimateAdvantage            parllel/transforms/advantage.py:36-57 (compute_gae_advantage)
 *   adv[T-1] = r[T-1] + g*boot*nd[T-1] - v[T-1]
 *   adv[t]   = r[t] + g*v[t+1]*nd[t] - v[t] + g*l*nd[t]*adv[t+1];   ret = adv + v (float32)
 * `inner` > 1: value/bootstrap/advantage/return_ carry a trailing agent axis of that size
 * while reward/done do not (parllel/transforms/multi_advantage.py:118-146 broadcast).
 * ---------------------------------------------------------------------------------- */
PLB_API int plb_compute_gae_advantage_f32(
    const float *reward, int64_t ld_reward,
    const float *value, int64_t ld_value,
    const uint8_t *done, int64_t ld_done,
    const float *bootstrap_value,
    float *out_advantage, int64_t ld_advantage,
    float *out_return, int64_t ld_return,
    int64_t T, int64_t B, int64_t inner,
    double discount, double gae_lambda,
    int algo, plb_stream_t stream);

/* parllel/transforms/advantage.py:11-33 (compute_discount_return), the lambda == 1 branch
 *   ret[T-1] = r[T-1] + g*boot*nd[T-1];  ret[t] = r[t] + ret[t+1]*g*nd[t];  adv = ret - v */
PLB_API int plb_compute_discount_return_f32(
    const float *reward, int64_t ld_reward,
    const float *value, int64_t ld_value,
    const uint8_t *done, int64_t ld_done,
    const float *bootstrap_value,
    float *out_advantage, int64_t ld_advantage,
    float *out_return, int64_t ld_return,
    int64_t T, int64_t B, int64_t inner,
    double discount,
    int algo, plb_stream_t stream);

/* parllel/transforms/norm_rewards.py:16-33 (compute_past_discount_return)
 *   out[0] = r[0] + prev_return*g*~prev_done;  out[t] = r[t] + out[t-1]*g*~done[t-1]
 * previous_return / previous_done are the [B] padding rows past_return[-1] / done[-1]
 * (norm_rewards.py:92,95). */
PLB_API int plb_compute_past_discount_return_f32(
    const float *reward, int64_t ld_reward,
    const uint8_t *done, int64_t ld_done,
    const float *previous_return, const uint8_t *previous_done,
    float *out_return, int64_t ld_out,
    int64_t T, int64_t B,
    double discount,
    int algo, plb_stream_t stream);

/* ------------------------------------------------------------------------------------
 * RunningMeanStd               parllel/transforms/running_mean_std.py:49-57
 * Batch moments (np.mean / np.var, :52-53) as a Welford/Chan parallel reduction.
 * Moments are always the triple (count, mean, M2 = sum (x-mean)^2) in float64.
 * ---------------------------------------------------------------------------------- */
/* Moments of ALL elements of a [rows, cols] region (or of those with valid != 0;
 * `valid` may be NULL).  Writes moments_out[3] = (count, mean, M2). */
PLB_API size_t plb_batch_moments_workspace_bytes(void);
PLB_API int plb_batch_moments_f32(
    const float *x, int64_t ld_x, int64_t rows, int64_t cols,
    const uint8_t *valid, int64_t ld_valid,
    double *moments_out,
    void *workspace, size_t workspace_bytes, plb_stream_t stream);

/* Moments of the float32 difference x - sub over the same region: the residual term of
 * explained_variance (parllel/torch/utils.py:135-156: 1 - Var[y_true - y_pred] / Var[y_true]). */
PLB_API int plb_residual_moments_f32(const float *x, int64_t ld_x, const float *sub, int64_t ld_sub,
                                     int64_t rows, int64_t cols, const uint8_t *valid, int64_t ld_valid,
                                     double *moments_out, void *workspace, size_t workspace_bytes,
                                     plb_stream_t stream);

/* Valid mask of a recurrent batch (parllel/samplers/recurrent.py:102-103,141-142):
 * valid[0] = 1; valid[t+1] = valid[t] & ~done[t], for t + 1 < T.  1-byte masks, row strides in elements. */
PLB_API int plb_valid_prefix_u8(const uint8_t *done, int64_t ld_done, int64_t T, int64_t B,
                                uint8_t *valid_out, int64_t ld_valid, plb_stream_t stream);

/* Per-column moments of a [rows, cols] region over the rows with row_valid != 0
 * (row_valid may be NULL).  Writes count_out[1], mean_out[cols], m2_out[cols]. */
PLB_API size_t plb_column_moments_workspace_bytes(int64_t rows, int64_t cols);
PLB_API int plb_column_moments_f32(
    const float *x, int64_t ld_x, int64_t rows, int64_t cols,
    const uint8_t *row_valid,
    double *count_out, double *mean_out, double *m2_out,
    void *workspace, size_t workspace_bytes, plb_stream_t stream);

/* Multi-GPU exchange step: merge `n_parts` per-rank moments (all-gathered by the host through
 * NCCL) in rank order, so every rank derives bit-identical statistics.  This is the exact
 * parallel merge that replaces the reference's only collective, the mean-of-variances
 * all_reduce in parllel/torch/models.py:212-218.
 *   layout 0: parts[n_parts][n_cols][3] (count, mean, M2) triples -> out[n_cols][3]
 *   layout 1: parts[n_parts][1 + 2*n_cols] (count, mean[], M2[])  -> out[1 + 2*n_cols] */
PLB_API int plb_merge_moments_f64(
    const double *parts, int32_t n_parts, int64_t n_cols, int layout, double *out,
    plb_stream_t stream);

/* The same exchange as ONE kernel over NVLink peer memory (no NCCL on the path): all-gather of the
 * per-rank moments by direct P2P stores into every rank's symmetric buffer + release/acquire flags
 * at system scope + the rank-ordered merge.  `peer_buffers` is a DEVICE array of `world` pointers to
 * every rank's exchange buffer as mapped into this process (e.g. torch symmetric memory), each
 * plb_exchange_buffer_bytes() long and zero-filled once.  One process per GPU. */
PLB_API size_t plb_exchange_buffer_bytes(int32_t world, int64_t n_vals);
PLB_API int plb_exchange_moments_p2p(
    const double *local, int64_t n_vals, void *const *peer_buffers, int32_t rank, int32_t world,
    int layout, int64_t n_cols, double *merged_out, plb_stream_t stream);

/* parllel/transforms/running_mean_std.py:7-32 (update_from_moments): float64 Chan merge
 * of device-resident batch moments into the device-resident running state
 * (mean[n], var[n], count[1]).  batch_var = batch_m2 / batch_count. */
PLB_API int plb_update_from_moments_f64(
    const double *batch_count, const double *batch_mean, const double *batch_m2,
    double *mean, double *var, double *count, int64_t n,
    int flags, plb_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Elementwise passes
 * ---------------------------------------------------------------------------------- */
/* NormalizeRewards tail + ClipRewards fused:  x = clip( float32( x / sqrt(var + eps) ) )
 *   parllel/transforms/norm_rewards.py:114 and parllel/transforms/clip_rewards.py:36.
 * `var` is a DEVICE pointer to the running float64 variance, or NULL for "no scaling";
 * clip bounds are NaN for "no bound".  In place over a [rows, cols] region. */
PLB_API int plb_scale_clip_f32(
    float *x, int64_t ld_x, int64_t rows, int64_t cols,
    const double *var, double eps, double clip_lo, double clip_hi,
    plb_stream_t stream);

/* NormalizeAdvantage tail      parllel/transforms/norm_advantage.py:51-54
 *   adv = (adv - float32(mean)) / (float32(std) + float32(eps)),  all in float32,
 * with per-`inner` statistics moments[inner][3] = (count, mean, M2) on the device. */
PLB_API int plb_normalize_advantage_f32(
    float *advantage, int64_t ld, int64_t rows, int64_t cols, int64_t inner,
    const double *moments, double eps, plb_stream_t stream);

/* NormalizeObservations tail   parllel/transforms/norm_obs.py:72-74
 *   obs[r, f] = float32( (obs[r, f] - mean[f]) / sqrt(var[f] + eps) )   (float64 arithmetic) */
PLB_API int plb_normalize_columns_f32(
    float *x, int64_t ld_x, int64_t rows, int64_t cols,
    const double *mean, const double *var, double eps, plb_stream_t stream);

/* same pass with the clamp EXTENSION of plb_normalize_observations_step_f32 fused in (NaN = no bound) */
PLB_API int plb_normalize_columns_clip_f32(
    float *x, int64_t ld_x, int64_t rows, int64_t cols,
    const double *mean, const double *var, double eps, double clip_lo, double clip_hi, plb_stream_t stream);

/* NormalizeObservations.__call__ for one step   parllel/transforms/norm_obs.py:57-76
 * update running stats from obs[row_valid] FIRST, then normalise all rows with the
 * updated stats.  Single entry; chooses the on-chip single-pass kernel when a
 * [rows x column-tile] slab fits in shared memory, else a 3-kernel sequence. */
PLB_API size_t plb_normalize_observations_workspace_bytes(int64_t rows, int64_t cols);
PLB_API int plb_normalize_observations_step_f32(
    float *obs, int64_t ld_obs, int64_t rows, int64_t cols,
    const uint8_t *row_valid,
    double *mean, double *var, double *count, double eps, int flags,
    double clip_lo, double clip_hi, /* EXTENSION (not in the reference): clamp the normalised values; NaN = off */
    void *workspace, size_t workspace_bytes, plb_stream_t stream);

/* ------------------------------------------------------------------------------------
 * ReplayBuffer.sample_batch gather      parllel/replays/replay.py:72  (tree[T_idxs, B_idxs])
 * One launch gathers every leaf of the ring tree: dst[i, :] = src[t_idx[i], b_idx[i], :].
 * ---------------------------------------------------------------------------------- */
typedef struct {
    const void *src;        /* leaf base: element [0, 0, ...] of the visible ring window   */
    void *dst;              /* [n, row_bytes] dense output                                 */
    int64_t row_bytes;      /* bytes of one (t, b) entry (prod(feature_shape) * itemsize)  */
    int64_t stride_t_bytes; /* bytes between consecutive t                                 */
    int64_t stride_b_bytes; /* bytes between consecutive b                                 */
} plb_gather_leaf;

#define PLB_GATHER_MAX_LEAVES 16

/* `leaves` is a HOST array (copied into the kernel's parameter space); t_idx / b_idx are
 * DEVICE int64[n]. */
PLB_API int plb_gather_tree(
    const plb_gather_leaf *leaves, int32_t n_leaves,
    const int64_t *t_idx, const int64_t *b_idx, int64_t n,
    plb_stream_t stream);

/* ReplayBuffer.sample_batch index generation on the device   parllel/replays/replay.py:55-70
 *   T_idxs = rng.integers(0, valid_length, n) (+ (cursor + oldest_invalid)) % size_T when full),
 *   B_idxs = rng.integers(0, B, n), T draw before B draw, for `k` consecutive minibatches,
 * bit-exact with numpy.random.Generator(PCG64) (Lemire rejection on 32-bit half-words, the
 * half-word buffer persisting across calls).  `pcg64_state` is DEVICE memory, 5 x uint64:
 * {state_hi, state_lo, inc_hi, inc_lo, has_uint32 | (uinteger << 32)} exactly as numpy's
 * bit_generator.state reports them; it is advanced in place.  Outputs are int64 [k, n]. */
PLB_API int plb_replay_sample_indices(
    uint64_t *pcg64_state, int64_t size_T, int64_t B, int64_t cursor, int full,
    int64_t newest_invalid, int64_t oldest_invalid, int64_t n, int64_t k,
    int64_t *t_idx, int64_t *b_idx, plb_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Fused batch-transform chain        parllel/samplers/basic.py:124-125 (`for transform in
 * batch_transforms: sample_tree = transform(sample_tree)`) specialised to the chain the examples
 * build (examples/cartpole/train_ppo.py:101-121):
 *   NormalizeRewards -> ClipRewards -> EstimateAdvantage -> NormalizeAdvantage
 * Each stage can be switched off.  Results equal the four separate calls (same kernels' maths);
 * the reward rescale/clip is folded into the advantage scan's load path and the advantage
 * moments into its store path.  `phases` selects which part to enqueue so that a multi-GPU host
 * can merge `return_moments` (after A) and `advantage_moments` (after B) across ranks with
 * plb_merge_moments_f64 before the next phase; single-GPU callers pass PLB_PHASE_ALL.
 * ---------------------------------------------------------------------------------- */
#define PLB_PHASE_A 1 /* past-return scan + moments of past_return            */
#define PLB_PHASE_B 2 /* update running stats; advantage scan (+scale/clip, +advantage moments) */
#define PLB_PHASE_C 4 /* normalise advantage                                  */
#define PLB_PHASE_ALL 7
/* modifier for single-GPU callers that issue B and C as separate calls on the same descriptor and
 * workspace: the advantage statistics travel as the scan's per-CTA partial sums (no serial last-CTA
 * reduction in the scan; `advantage_moments` is written by the C call) -- pass it to both calls */
#define PLB_PHASE_PARTIALS 8

typedef struct {
    int64_t T, B;
    /* leaves ([T, B] views, row strides in elements) */
    float *reward; int64_t ld_reward;                /* in/out */
    const float *value; int64_t ld_value;
    const uint8_t *done; int64_t ld_done;
    const float *bootstrap_value;                    /* [B] */
    float *advantage; int64_t ld_advantage;          /* out */
    float *return_; int64_t ld_return;               /* out */
    float *past_return; int64_t ld_past_return;      /* out (NormalizeRewards) */
    const float *previous_return;                    /* [B]  past_return[-1] */
    const uint8_t *previous_done;                    /* [B]  done[-1]        */
    const uint8_t *valid; int64_t ld_valid;          /* optional mask for the statistics, or NULL */
    /* stages */
    int32_t normalize_rewards, clip_rewards, estimate_advantage, normalize_advantage;
    double discount, gae_lambda;
    double clip_lo, clip_hi;                         /* NaN = no bound */
    double eps_reward, eps_advantage;                /* 1e-6 in the reference */
    /* device-resident state / scratch */
    double *return_mean, *return_var, *return_count; /* RunningMeanStd(shape=()) of NormalizeRewards */
    double *return_moments;                          /* [3] batch moments of past_return   */
    double *advantage_moments;                       /* [3] batch moments of advantage     */
    int32_t moments_flags;                           /* PLB_MOMENTS_F32_FLOW or 0 */
    void *workspace; size_t workspace_bytes;
    /* multi-GPU (B sharded), optional: with `xch_peers` != NULL and phases == PLB_PHASE_ALL both
     * statistics are exchanged INSIDE the kernels over NVLink peer memory -- the scan that produces a
     * statistic pushes the rank's moments into every peer's buffer from its last CTA, the kernel that
     * consumes it waits for all ranks and merges in rank order in its prologue.  `xch_peers` is the
     * DEVICE pointer table of plb_exchange_moments_p2p (buffers sized for n_vals = 3). */
    void *const *xch_peers; int32_t xch_rank, xch_world;
    /* optional [B] buffers that receive past_return[T-1] / done[T-1] at the end of phase A -- the rows
     * Array.rotate() (arrays/array.py:394-418) would copy into the padding for the next batch; for leaves
     * without padding rows.  May alias previous_return / previous_done.  NULL = not wanted. */
    float *carry_return_out; uint8_t *carry_done_out;
    /* optional: grid-exchange buffers (plb_grid_exchange_bytes) of every rank as a DEVICE pointer table
     * [gx_world]; on one GPU a table with the single local buffer.  With them, EstimateAdvantage +
     * NormalizeAdvantage (norm_advantage.py:30-56 after advantage.py:91-106) run as ONE cooperative
     * launch wherever the batch allows it (PLB_PLAN_RESIDENT): the advantage waits in tensor memory while
     * every CTA of every rank exchanges its partial moments through these buffers, and is written once,
     * normalised.  `gx_slots` = records per rank (plb_grid_exchange_slots(), identical on all ranks). */
    void *const *gx_peers; int32_t gx_rank, gx_world, gx_slots;
    void *gx_self;                                   /* this rank's buffer, i.e. gx_peers[gx_rank] (saves the kernels a dependent load) */
    /* which fused routes the call may take: PLB_PLAN_AUTO on one GPU; on several GPUs the bitwise AND over
     * all ranks of plb_batch_pipeline_plan(), so that every rank takes the same exchange protocol */
    int32_t plan;
} plb_batch_pipeline;

#define PLB_PLAN_AUTO (-1)
#define PLB_PLAN_CHUNKED 1  /* both scans can run as TMA-staged chunked scans (in-kernel moment exchange possible) */
#define PLB_PLAN_RESIDENT 2 /* EstimateAdvantage + NormalizeAdvantage fit the single-launch resident kernel */

PLB_API size_t plb_batch_pipeline_workspace_bytes(void);
PLB_API int plb_batch_pipeline_f32(const plb_batch_pipeline *desc, int phases, plb_stream_t stream);
/* Which fused routes THIS rank's leaves allow (PLB_PLAN_* bits; >= 0).  Pure host-side inspection of
 * the descriptor (shapes, alignment, strides); nothing is enqueued. */
PLB_API int plb_batch_pipeline_plan(const plb_batch_pipeline *desc);

/* Grid-exchange buffers of the resident kernel: `slots` = plb_grid_exchange_slots() records per rank;
 * a buffer is plb_grid_exchange_bytes(world, slots) long, zero-filled ONCE, and (for world > 1) mapped
 * into every peer (e.g. torch symmetric memory).  All ranks must issue the same sequence of calls on one
 * buffer set.  plb_grid_exchange_status copies the status word back (synchronises `stream`):
 * 0 = ok, 1 = a launch timed out waiting for a peer (its results are NaN), <0 = library error. */
PLB_API int plb_grid_exchange_slots(void);
PLB_API size_t plb_grid_exchange_bytes(int32_t world, int32_t slots);
PLB_API int plb_grid_exchange_status(const void *local_buffer, plb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PARLLEL_B200_H */
