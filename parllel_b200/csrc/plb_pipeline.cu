// plb_pipeline.cu -- the fused batch-transform chain of the samplers (samplers/basic.py:124-125):
//
//     NormalizeRewards -> ClipRewards -> EstimateAdvantage -> NormalizeAdvantage
//     (norm_rewards.py:86-116, clip_rewards.py:31-38, advantage.py:91-106, norm_advantage.py:30-56)
//
// One C call enqueues the whole chain.  Compared with calling the four transforms one by one,
// the reward rescale + clip is folded into the load path of the advantage scan (the transformed
// reward is written back once), and the moments NormalizeAdvantage needs are accumulated in the
// store path of the same scan, so the batch is streamed 3 times instead of 6:
//
//   phase A  forward past-return scan  + fused moments of past_return      (read 5 B, write 4 B /tr)
//   phase B  update running statistics; advantage scan with fused reward
//            scale/clip on load and advantage moments on store             (read 9 B, write 12 B /tr)
//   phase C  normalise advantage                                           (read 4 B, write 4 B /tr)
//
// The phases are separately selectable so a multi-GPU host can all-gather + merge the moments
// (plb_merge_moments_f64) between them.  Any stage can be switched off; shapes the TMA scan cannot
// describe fall back to the unfused kernels with identical results.
#include <math.h>

#include "plb_common.cuh"
#include "plb_scan_shared.cuh"

namespace plb {

static bool is_nan(double x) { return x != x; }

static int run_pipeline(const plb_batch_pipeline *d, int phases, cudaStream_t stream)
{
    PLB_REQUIRE(d != nullptr, PLB_ERR_INVALID_ARGUMENT, "null descriptor");
    PLB_REQUIRE(d->T >= 0 && d->B >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (d->T == 0 || d->B == 0) return PLB_OK;
    const int64_t T = d->T, B = d->B;
    const bool norm_rew = d->normalize_rewards != 0;
    const bool clip = d->clip_rewards != 0;
    const bool est_adv = d->estimate_advantage != 0;
    const bool norm_adv = d->normalize_advantage != 0;
    PLB_REQUIRE(d->reward != nullptr || !(norm_rew || clip || est_adv), PLB_ERR_INVALID_ARGUMENT, "reward is null");
    const size_t ws_need = 2 * MOMENTS_WS_BYTES;
    PLB_REQUIRE(d->workspace != nullptr && d->workspace_bytes >= ws_need, PLB_ERR_WORKSPACE,
                "pipeline needs %zu workspace bytes, got %zu", ws_need, d->workspace_bytes);
    unsigned char *ws = reinterpret_cast<unsigned char *>(d->workspace);
    void *ws_a = ws, *ws_b = ws + MOMENTS_WS_BYTES;
    int rc;
    // PLB_PHASE_PARTIALS: phases B and C are issued by separate calls on one GPU and the advantage
    // statistics travel between them as the scan's per-CTA partial sums (as in a PLB_PHASE_ALL call)
    const bool split_partials = (phases & PLB_PHASE_PARTIALS) != 0;
    phases &= ~PLB_PHASE_PARTIALS;
    PLB_REQUIRE(!split_partials || !(d->xch_peers != nullptr && d->xch_world > 1), PLB_ERR_INVALID_ARGUMENT,
                "PLB_PHASE_PARTIALS is a single-GPU mode");
    // whole chain in one call on one GPU: the advantage scan only publishes per-CTA partial sums and
    // the normalise kernel reduces them in its prologue (no serial last-CTA epilogue in the scan)
    const bool direct_partials = (phases == PLB_PHASE_ALL || split_partials) && est_adv && norm_adv && (T * B) % 4 == 0 &&
                                 !(d->xch_peers != nullptr && d->xch_world > 1);
    bool adv_partials_published = direct_partials && split_partials && !(phases & PLB_PHASE_B);
    // same idea for the reward statistics: on one GPU the CTA that finishes the past-return moments
    // also applies update_from_moments to the running state (no separate update launch)
    // multi-GPU: exchange both statistics inside the kernels over peer memory
    XchRef xch;
    xch.peers = reinterpret_cast<unsigned char *const *>(d->xch_peers);
    xch.rank = d->xch_rank; xch.world = d->xch_world;
    const bool use_xch = (phases == PLB_PHASE_ALL) && xch.peers != nullptr && xch.world > 1;
    PLB_REQUIRE(!use_xch || (xch.world <= 32 && xch.rank >= 0 && xch.rank < xch.world), PLB_ERR_INVALID_ARGUMENT,
                "in-kernel exchange supports up to 32 ranks");
    // the in-kernel exchange only exists in the chunked scans: every rank must be able to take them, which
    // the caller establishes by AND-ing plb_batch_pipeline_plan() over the ranks (uneven shards may differ)
    PLB_REQUIRE(!use_xch || (d->plan != PLB_PLAN_AUTO && (d->plan & PLB_PLAN_CHUNKED)), PLB_ERR_INVALID_ARGUMENT,
                "in-kernel exchange needs an agreed plan with PLB_PLAN_CHUNKED (see plb_batch_pipeline_plan)");
    bool adv_pushed = false;
    // EstimateAdvantage + NormalizeAdvantage as one cooperative launch (advantage resident in tensor memory)
    const int plan = d->plan;
    GridXch gx;
    gx.peers = reinterpret_cast<unsigned char *const *>(d->gx_peers);
    gx.self = reinterpret_cast<unsigned char *>(d->gx_self);
    gx.rank = d->gx_rank; gx.world = d->gx_world; gx.slots = d->gx_slots;
    PLB_REQUIRE(gx.peers == nullptr || gx.self != nullptr, PLB_ERR_INVALID_ARGUMENT, "gx_peers without gx_self");
    // sharded batches (gx.world > 1) have two routes for EstimateAdvantage + NormalizeAdvantage (tuning key 17):
    //   3 (auto)  last-CTA reduction in the scan, ONE record of partial sums per rank pushed to every rank, every CTA
    //             of the normalise kernel adds the `world` records up itself (xch_push3 / xch_wait_merge3);
    //   1         the resident kernel (one launch; its write-only final phase runs at ~half of HBM speed).
    // (A third one -- per-CTA records + a collector CTA in the normalise kernel -- measured slower at every N and was
    //  removed: DESIGN.md section 4.)
    const bool multi = gx.peers != nullptr && gx.world > 1;
    const int route = sharded_route() == 1 ? 1 : 3;
    const bool both_phases = est_adv && norm_adv && (phases & PLB_PHASE_B) && (phases & PLB_PHASE_C) && !split_partials;
    const bool want_resident = both_phases && gx.peers != nullptr && (plan & PLB_PLAN_RESIDENT) &&
                               (multi ? (route == 1 && plan != PLB_PLAN_AUTO) : resident_wanted(1));
    // a sharded single-call chain must exchange the advantage statistics one way or another
    PLB_REQUIRE(!(both_phases && phases == PLB_PHASE_ALL && (multi || (d->xch_peers != nullptr && d->xch_world > 1))) ||
                    want_resident || (d->xch_peers != nullptr && d->xch_world > 1),
                PLB_ERR_INVALID_ARGUMENT, "sharded chain without an exchange route (plan %d, route %d)", plan, route);
    const bool update_in_tail = (phases == PLB_PHASE_ALL) && norm_rew && !use_xch;
    bool state_updated = false;

    // ---------------------------------------------------------------- phase A
    if ((phases & PLB_PHASE_A) && norm_rew) {
        PLB_REQUIRE(d->done && d->past_return && d->previous_return && d->previous_done && d->return_moments,
                    PLB_ERR_INVALID_ARGUMENT, "NormalizeRewards stage: missing leaf");
        ScanFusion f = {};
        f.valid = d->valid; f.ld_valid = d->ld_valid;
        f.tail = make_tail(d->return_moments, ws_a);
        if (update_in_tail) {
            PLB_REQUIRE(d->return_mean && d->return_var && d->return_count, PLB_ERR_INVALID_ARGUMENT,
                        "NormalizeRewards stage: missing running statistics");
            f.tail.upd_mean = d->return_mean; f.tail.upd_var = d->return_var; f.tail.upd_count = d->return_count;
            f.tail.upd_flags = d->moments_flags;
        }
        if (use_xch) f.tail.xch = xch;
        f.carry_ret_out = d->carry_return_out; f.carry_done_out = d->carry_done_out;
        rc = scan_forward(d->reward, d->ld_reward, d->done, d->ld_done, d->previous_return, d->previous_done,
                          d->past_return, d->ld_past_return, T, B, d->discount, PLB_ALGO_CHUNKED, &f, stream);
        if (rc == PLB_OK && update_in_tail) state_updated = true;
        if (rc == PLB_OK && use_xch) {
            // wait for every rank's push, merge in rank order, fold into the running state
            PLB_REQUIRE(d->return_mean && d->return_var && d->return_count, PLB_ERR_INVALID_ARGUMENT,
                        "NormalizeRewards stage: missing running statistics");
            rc = exchange_finish(xch, d->return_moments, d->return_mean, d->return_var, d->return_count,
                                 d->moments_flags, stream);
            if (rc) return rc;
            state_updated = true;
        }
        if (rc == PLB_ERR_UNSUPPORTED && use_xch) return rc;  // agreed plan broken: never switch protocol alone
        if (rc == PLB_ERR_UNSUPPORTED) {  // shape TMA cannot describe: serial scan + separate moments
            ScanFusion fc = {};  // only the carry rows travel with the serial scan
            fc.carry_ret_out = d->carry_return_out; fc.carry_done_out = d->carry_done_out;
            rc = scan_forward(d->reward, d->ld_reward, d->done, d->ld_done, d->previous_return, d->previous_done,
                              d->past_return, d->ld_past_return, T, B, d->discount, PLB_ALGO_SERIAL, &fc, stream);
            if (rc) return rc;
            rc = plb_batch_moments_f32(d->past_return, d->ld_past_return, T, B, d->valid, d->ld_valid,
                                       d->return_moments, ws_a, MOMENTS_WS_BYTES, stream);
            if (rc == PLB_OK && use_xch)
                rc = plb_exchange_moments_p2p(d->return_moments, 3, d->xch_peers, xch.rank, xch.world, 0, 1,
                                              d->return_moments, stream);
        }
        if (rc) return rc;
    }

    // ---------------------------------------------------------------- phase B
    if (phases & PLB_PHASE_B) {
        if (norm_rew && !state_updated) {
            PLB_REQUIRE(d->return_mean && d->return_var && d->return_count, PLB_ERR_INVALID_ARGUMENT,
                        "NormalizeRewards stage: missing running statistics");
            rc = plb_update_from_moments_f64(d->return_moments, d->return_moments + 1, d->return_moments + 2,
                                             d->return_mean, d->return_var, d->return_count, 1,
                                             d->moments_flags, stream);
            if (rc) return rc;
        }
        const double lo = clip ? d->clip_lo : NAN, hi = clip ? d->clip_hi : NAN;
        const bool has_pre = norm_rew || (clip && !(is_nan(lo) && is_nan(hi)));
        if (est_adv) {
            PLB_REQUIRE(d->value && d->done && d->bootstrap_value && d->advantage && d->return_,
                        PLB_ERR_INVALID_ARGUMENT, "EstimateAdvantage stage: missing leaf");
            PLB_REQUIRE(!norm_adv || d->advantage_moments, PLB_ERR_INVALID_ARGUMENT, "advantage_moments is null");
            ScanFusion f = {};
            f.has_pre = has_pre ? 1 : 0;
            f.reward_var = norm_rew ? d->return_var : nullptr;
            f.reward_eps = d->eps_reward;
            f.clip_lo = lo; f.clip_hi = hi;
            f.reward_out = d->reward; f.ld_reward_out = d->ld_reward;
            if (norm_adv) {
                f.valid = d->valid; f.ld_valid = d->ld_valid;
                f.tail = make_tail(d->advantage_moments, ws_b);
                f.tail.partials_only = direct_partials ? 1 : 0;
                if (use_xch) f.tail.xch = xch;
                f.tail.dbg = debug_buffer();
            }
            const int mode = (d->gae_lambda == 1.0) ? 1 : 0;
            const bool resident = want_resident &&
                                  resident_supported(d->reward, d->ld_reward, d->value, d->ld_value, d->done, d->ld_done,
                                                     d->advantage, d->ld_advantage, d->return_, d->ld_return, T, B) &&
                                  (!has_pre || (aligned(d->reward, 16) && d->ld_reward % 4 == 0));
            // several GPUs: the route was agreed through `plan`; a rank that cannot follow must not pick another
            PLB_REQUIRE(resident || !(want_resident && gx.world > 1), PLB_ERR_UNSUPPORTED,
                        "plan includes PLB_PLAN_RESIDENT but this rank's leaves do not allow it");
            if (resident) {
                ScanFusion fr = f;
                fr.tail = make_tail(d->advantage_moments, ws_b);
                fr.resident = 1;
                fr.gx = gx;
                fr.adv_eps = d->eps_advantage;
                fr.adv_moments_out = d->advantage_moments;
                rc = scan_reverse(mode, d->reward, d->ld_reward, d->value, d->ld_value, d->done, d->ld_done,
                                  d->bootstrap_value, d->advantage, d->ld_advantage, d->return_, d->ld_return,
                                  T, B, 1, d->discount, d->gae_lambda, PLB_ALGO_CHUNKED, &fr, stream);
                if (rc) return rc;
                return PLB_OK;  // phases B and C are complete
            }
            rc = scan_reverse(mode, d->reward, d->ld_reward, d->value, d->ld_value, d->done, d->ld_done,
                              d->bootstrap_value, d->advantage, d->ld_advantage, d->return_, d->ld_return,
                              T, B, 1, d->discount, d->gae_lambda, PLB_ALGO_CHUNKED,
                              (has_pre || norm_adv) ? &f : nullptr, stream);
            if (rc == PLB_OK && norm_adv && direct_partials) adv_partials_published = true;
            else if (rc == PLB_OK && norm_adv && use_xch) adv_pushed = true;
            if (rc == PLB_ERR_UNSUPPORTED && use_xch) return rc;
            if (rc == PLB_ERR_UNSUPPORTED) {  // unfused fallback, same results
                if (has_pre) {
                    rc = plb_scale_clip_f32(d->reward, d->ld_reward, T, B, f.reward_var, d->eps_reward, lo, hi, stream);
                    if (rc) return rc;
                }
                rc = scan_reverse(mode, d->reward, d->ld_reward, d->value, d->ld_value, d->done, d->ld_done,
                                  d->bootstrap_value, d->advantage, d->ld_advantage, d->return_, d->ld_return,
                                  T, B, 1, d->discount, d->gae_lambda, PLB_ALGO_SERIAL, nullptr, stream);
                if (rc) return rc;
                if (norm_adv) {
                    rc = plb_batch_moments_f32(d->advantage, d->ld_advantage, T, B, d->valid, d->ld_valid,
                                               d->advantage_moments, ws_b, MOMENTS_WS_BYTES, stream);
                    if (rc == PLB_OK && use_xch)
                        rc = plb_exchange_moments_p2p(d->advantage_moments, 3, d->xch_peers, xch.rank, xch.world, 0, 1,
                                                      d->advantage_moments, stream);
                }
            }
            if (rc) return rc;
        } else {
            if (has_pre) {
                rc = plb_scale_clip_f32(d->reward, d->ld_reward, T, B, norm_rew ? d->return_var : nullptr,
                                        d->eps_reward, lo, hi, stream);
                if (rc) return rc;
            }
            if (norm_adv) {
                PLB_REQUIRE(d->advantage && d->advantage_moments, PLB_ERR_INVALID_ARGUMENT, "advantage is null");
                rc = plb_batch_moments_f32(d->advantage, d->ld_advantage, T, B, d->valid, d->ld_valid,
                                           d->advantage_moments, ws_b, MOMENTS_WS_BYTES, stream);
                if (rc == PLB_OK && use_xch)
                    rc = plb_exchange_moments_p2p(d->advantage_moments, 3, d->xch_peers, xch.rank, xch.world, 0, 1,
                                                  d->advantage_moments, stream);
                if (rc) return rc;
            }
        }
    }

    // ---------------------------------------------------------------- phase C
    if ((phases & PLB_PHASE_C) && norm_adv && adv_pushed) {
        rc = normalize_advantage_after_exchange(d->advantage, d->ld_advantage, T, B, xch, d->eps_advantage,
                                                d->advantage_moments, stream);
        if (rc == PLB_OK) return PLB_OK;
        if (rc != PLB_ERR_UNSUPPORTED) return rc;
        rc = exchange_finish(xch, d->advantage_moments, nullptr, nullptr, nullptr, 0, stream);
        if (rc) return rc;
    }
    if ((phases & PLB_PHASE_C) && norm_adv && adv_partials_published) {
        rc = normalize_advantage_from_partials(d->advantage, d->ld_advantage, T, B, ws_b, d->eps_advantage,
                                               d->advantage_moments, stream);
        if (rc == PLB_OK) return PLB_OK;
        if (rc != PLB_ERR_UNSUPPORTED) return rc;
        // not 128-bit addressable: finish the reduction with a one-CTA pass, then the generic kernel
        rc = finalize_published_partials(ws_b, d->advantage_moments, stream);
        if (rc) return rc;
    }
    if ((phases & PLB_PHASE_C) && norm_adv) {
        rc = plb_normalize_advantage_f32(d->advantage, d->ld_advantage, T, B, 1, d->advantage_moments,
                                         d->eps_advantage, stream);
        if (rc) return rc;
    }
    return PLB_OK;
}

// host-side inspection: which fused routes do this rank's leaves allow?
static int pipeline_plan(const plb_batch_pipeline *d)
{
    PLB_REQUIRE(d != nullptr, PLB_ERR_INVALID_ARGUMENT, "null descriptor");
    if (d->T <= 0 || d->B <= 0) return PLB_PLAN_CHUNKED | PLB_PLAN_RESIDENT;  // nothing to do: follows any route
    int plan = 0;
    const bool fwd_ok = !d->normalize_rewards ||
                        (scan_chunked_supported(d->reward, d->ld_reward, nullptr, 0, d->done, d->ld_done, d->T, d->B, 1) &&
                         aligned(d->past_return, 16) && d->ld_past_return % 4 == 0);
    const bool rev_ok = !d->estimate_advantage ||
                        (scan_chunked_supported(d->reward, d->ld_reward, d->value, d->ld_value, d->done, d->ld_done, d->T, d->B, 1) &&
                         aligned(d->advantage, 16) && d->ld_advantage % 4 == 0 && aligned(d->return_, 16) && d->ld_return % 4 == 0);
    const bool adv_vec = !d->normalize_advantage ||
                         (aligned(d->advantage, 16) && d->ld_advantage % 4 == 0 && (d->ld_advantage == d->B ? (d->T * d->B) % 4 == 0 : d->B % 4 == 0));
    if (fwd_ok && rev_ok && adv_vec) plan |= PLB_PLAN_CHUNKED;
    if (d->estimate_advantage && d->normalize_advantage && rev_ok &&
        resident_supported(d->reward, d->ld_reward, d->value, d->ld_value, d->done, d->ld_done, d->advantage, d->ld_advantage,
                           d->return_, d->ld_return, d->T, d->B))
        plan |= PLB_PLAN_RESIDENT;
    return plan;
}

}  // namespace plb

extern "C" {

size_t plb_batch_pipeline_workspace_bytes(void) { return 2 * plb::MOMENTS_WS_BYTES; }

int plb_batch_pipeline_plan(const plb_batch_pipeline *desc) { return plb::pipeline_plan(desc); }

int plb_grid_exchange_slots(void)
{
    const int sms = plb::sm_count();
    if (sms <= 0) { plb::set_error("no CUDA device available"); return PLB_ERR_UNSUPPORTED; }
    return sms < plb::GX_MAX_SLOTS ? sms : plb::GX_MAX_SLOTS;
}

size_t plb_grid_exchange_bytes(int32_t world, int32_t slots)
{
    if (world < 1 || world > plb::GX_MAX_WORLD || slots < 1 || slots > plb::GX_MAX_SLOTS) return 0;
    return plb::gx_total_bytes(world, slots);
}

int plb_grid_exchange_status(const void *local_buffer, plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(local_buffer != nullptr, PLB_ERR_INVALID_ARGUMENT, "null buffer");
    uint32_t status = 0;
    PLB_CUDA(cudaMemcpyAsync(&status, reinterpret_cast<const unsigned char *>(local_buffer) + 8, sizeof(status),
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PLB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (status != 0) set_error("grid exchange timed out waiting for a peer rank (results of that launch are NaN)");
    return (int)status;
}

int plb_batch_pipeline_f32(const plb_batch_pipeline *desc, int phases, plb_stream_t stream)
{
    return plb::run_pipeline(desc, phases, (cudaStream_t)stream);
}

}  // extern "C"
