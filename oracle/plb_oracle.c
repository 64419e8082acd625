/*
 * plb_oracle.c -- CPU restatement of parllel's batch post-processing arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (parllel_b200/) may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and there only as the
 * checker / the timed CPU arm.
 *
 * Every function names the reference file:line (relative to /root/reference)
 * whose arithmetic it restates.  The reference is numba `njit(fastmath=True)`
 * code in which the python-float scalars (discount, gae_lambda) are float64, so
 * every right-hand side is evaluated in float64 per time step and ROUNDED TO
 * float32 WHEN STORED; the next step re-reads the float32 value.  This file
 * reproduces exactly that (no -ffast-math, no contraction: see oracle/Makefile),
 * so it is a strict IEEE emulation of the reference.  numba's fastmath may
 * contract/reassociate, so the live reference differs from this file by <=~1e-6
 * abs at C1 sizes (pinned by tests/golden, see tests/test_oracle_golden.py).
 *
 * Layout everywhere: time-major [T, n] with the n "column" axis contiguous and
 * an explicit row stride (in elements) per array, which is how the reference's
 * Array views into padded base buffers look (parllel/arrays/array.py:133,420-424).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(PLBO_NO_CLONES)
#define PLBO_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define PLBO_CLONES
#endif

#define PLBO_API __attribute__((visibility("default")))

/* parllel/transforms/advantage.py:36-57  compute_gae_advantage
 *   advantage[-1] = reward[-1] + discount * bootstrap_value * not_done[-1] - value[-1]
 *   for t = T-2 .. 0:
 *       delta        = reward[t] + discount * value[t+1] * not_done[t] - value[t]
 *       advantage[t] = delta + discount * gae_lambda * not_done[t] * advantage[t+1]
 *   return_[...] = advantage + value                      (float32 + float32)
 * `inner` > 1 restates the trailing-dim broadcast of multi_advantage.py:118-146:
 * value/bootstrap/advantage/return_ have n*inner columns, reward/done have n. */
PLBO_CLONES PLBO_API void plbo_gae_advantage(
    const float *reward, int64_t ld_reward,
    const float *value, int64_t ld_value,
    const uint8_t *done, int64_t ld_done,
    const float *bootstrap,
    float *advantage, int64_t ld_adv,
    float *return_, int64_t ld_ret,
    int64_t T, int64_t n, int64_t inner,
    double discount, double gae_lambda)
{
    if (T <= 0 || n <= 0) return;
    const int64_t nv = n * inner;
    {
        const int64_t t = T - 1;
        for (int64_t j = 0; j < nv; ++j) {
            const int64_t jr = j / inner;
            const double nd = done[t * ld_done + jr] ? 0.0 : 1.0;
            const double a = (double)reward[t * ld_reward + jr] +
                             discount * (double)bootstrap[j] * nd -
                             (double)value[t * ld_value + j];
            advantage[t * ld_adv + j] = (float)a;
        }
    }
    for (int64_t t = T - 2; t >= 0; --t) {
        const float *r = reward + t * ld_reward;
        const float *v = value + t * ld_value;
        const float *vn = value + (t + 1) * ld_value;
        const uint8_t *d = done + t * ld_done;
        const float *an = advantage + (t + 1) * ld_adv;
        float *a = advantage + t * ld_adv;
        if (inner == 1) {
            for (int64_t j = 0; j < n; ++j) {
                const double nd = d[j] ? 0.0 : 1.0;
                const double delta = (double)r[j] + discount * (double)vn[j] * nd - (double)v[j];
                a[j] = (float)(delta + discount * gae_lambda * nd * (double)an[j]);
            }
        } else {
            for (int64_t j = 0; j < nv; ++j) {
                const int64_t jr = j / inner;
                const double nd = d[jr] ? 0.0 : 1.0;
                const double delta = (double)r[jr] + discount * (double)vn[j] * nd - (double)v[j];
                a[j] = (float)(delta + discount * gae_lambda * nd * (double)an[j]);
            }
        }
    }
    for (int64_t t = 0; t < T; ++t)
        for (int64_t j = 0; j < nv; ++j)
            return_[t * ld_ret + j] = advantage[t * ld_adv + j] + value[t * ld_value + j];
}

/* parllel/transforms/advantage.py:11-33  compute_discount_return
 *   return_[-1] = reward[-1] + discount * bootstrap_value * not_done[-1]
 *   for t = T-2 .. 0: return_[t] = reward[t] + return_[t+1] * discount * not_done[t]
 *   advantage[...] = return_ - value                      (float32 - float32) */
PLBO_CLONES PLBO_API void plbo_discount_return(
    const float *reward, int64_t ld_reward,
    const float *value, int64_t ld_value,
    const uint8_t *done, int64_t ld_done,
    const float *bootstrap,
    float *advantage, int64_t ld_adv,
    float *return_, int64_t ld_ret,
    int64_t T, int64_t n, int64_t inner,
    double discount)
{
    if (T <= 0 || n <= 0) return;
    const int64_t nv = n * inner;
    {
        const int64_t t = T - 1;
        for (int64_t j = 0; j < nv; ++j) {
            const int64_t jr = j / inner;
            const double nd = done[t * ld_done + jr] ? 0.0 : 1.0;
            return_[t * ld_ret + j] =
                (float)((double)reward[t * ld_reward + jr] + discount * (double)bootstrap[j] * nd);
        }
    }
    for (int64_t t = T - 2; t >= 0; --t) {
        const float *r = reward + t * ld_reward;
        const uint8_t *d = done + t * ld_done;
        const float *rn = return_ + (t + 1) * ld_ret;
        float *o = return_ + t * ld_ret;
        if (inner == 1) {
            for (int64_t j = 0; j < n; ++j) {
                const double nd = d[j] ? 0.0 : 1.0;
                o[j] = (float)((double)r[j] + (double)rn[j] * discount * nd);
            }
        } else {
            for (int64_t j = 0; j < nv; ++j) {
                const int64_t jr = j / inner;
                const double nd = d[jr] ? 0.0 : 1.0;
                o[j] = (float)((double)r[jr] + (double)rn[j] * discount * nd);
            }
        }
    }
    for (int64_t t = 0; t < T; ++t)
        for (int64_t j = 0; j < nv; ++j)
            advantage[t * ld_adv + j] = return_[t * ld_ret + j] - value[t * ld_value + j];
}

/* parllel/transforms/norm_rewards.py:16-33  compute_past_discount_return
 *   return_[0] = reward[0] + previous_return * discount * previous_not_done
 *   for t = 1 .. T-1: return_[t] = reward[t] + return_[t-1] * discount * not_done[t-1] */
PLBO_CLONES PLBO_API void plbo_past_discount_return(
    const float *reward, int64_t ld_reward,
    const uint8_t *done, int64_t ld_done,
    const float *previous_return, const uint8_t *previous_done,
    float *out_return, int64_t ld_out,
    int64_t T, int64_t n, double discount)
{
    if (T <= 0 || n <= 0) return;
    for (int64_t j = 0; j < n; ++j) {
        const double pnd = previous_done[j] ? 0.0 : 1.0;
        out_return[j] = (float)((double)reward[j] + (double)previous_return[j] * discount * pnd);
    }
    for (int64_t t = 1; t < T; ++t) {
        const float *r = reward + t * ld_reward;
        const uint8_t *d = done + (t - 1) * ld_done;
        const float *p = out_return + (t - 1) * ld_out;
        float *o = out_return + t * ld_out;
        for (int64_t j = 0; j < n; ++j) {
            const double nd = d[j] ? 0.0 : 1.0;
            o[j] = (float)((double)r[j] + (double)p[j] * discount * nd);
        }
    }
}

/* parllel/transforms/running_mean_std.py:7-32  update_from_moments (float64 Chan merge)
 *   delta = batch_mean - mean; total = batch_count + count
 *   mean += delta * batch_count / total
 *   m_2   = var*count + batch_var*batch_count + delta^2 * count * batch_count / total
 *   var   = m_2 / total ; count = total
 * Measured numba typing (pinned by tests/golden/norm_obs_valid.npz): when the batch
 * is float32, `batch_var * batch_count` is float32-array x int64-scalar and is
 * evaluated in FLOAT32 (int scalars do not promote a float32 array); every other
 * term involves a float64 array and is float64.  `batch_f32` selects that flow. */
PLBO_API void plbo_update_from_moments(
    const double *batch_mean, const double *batch_var, double batch_count,
    double *mean, double *var, double *count, int64_t n, int batch_f32)
{
    const double cnt = *count;
    const double total = batch_count + cnt;
    for (int64_t i = 0; i < n; ++i) {
        const double delta = batch_mean[i] - mean[i];
        mean[i] = mean[i] + delta * batch_count / total;
        const double m_a = var[i] * cnt;
        const double m_b = batch_f32 ? (double)((float)batch_var[i] * (float)batch_count)
                                     : batch_var[i] * batch_count;
        const double m_2 = m_a + m_b + delta * delta * cnt * batch_count / total;
        var[i] = m_2 / total;
    }
    *count = total;
}

/* parllel/replays/replay.py:72  `self.tree[T_idxs, B_idxs]` for one leaf:
 * advanced-index row gather out of a [size_T, B, row_bytes] ring (byte copy). */
PLBO_API void plbo_gather_rows(
    const uint8_t *src, int64_t row_bytes, int64_t stride_t_bytes, int64_t stride_b_bytes,
    const int64_t *t_idx, const int64_t *b_idx, int64_t n, uint8_t *dst)
{
    for (int64_t i = 0; i < n; ++i)
        memcpy(dst + i * row_bytes, src + t_idx[i] * stride_t_bytes + b_idx[i] * stride_b_bytes,
               (size_t)row_bytes);
}

/* ---------------------------------------------------------------------------
 * numpy.random.Generator(PCG64).integers(0, high, size) restated.
 * Third-party dependency of parllel/replays/replay.py:44,64,68,70 (numpy is not
 * vendored under /root/reference; pinned numpy~=1.22.0 in requirements.txt:1,
 * installed 2.3.5).  Published algorithm:
 *   - PCG64 = 128-bit LCG (multiplier 0x2360ED051FC65DA44385DF649FCCF645) with
 *     XSL-RR 128/64 output; next_uint32 hands out the LOW then the HIGH half of
 *     one 64-bit draw, and the unused half is buffered in the bit generator
 *     ACROSS calls (has_uint32 / uinteger).
 *   - for int64 output with (high-1) < 2^32-1, Generator.integers uses Lemire's
 *     multiply-shift rejection on 32-bit draws (use_masked = False).
 * Pinned bit-for-bit against numpy itself in tests/test_oracle_rng.py.
 * ------------------------------------------------------------------------- */
typedef struct {
    uint64_t state_hi, state_lo; /* 128-bit LCG state            */
    uint64_t inc_hi, inc_lo;     /* 128-bit increment (odd)      */
    int32_t has_uint32;          /* buffered half-word available */
    uint32_t uinteger;           /* the buffered high half       */
} plbo_pcg64;

static const uint64_t PCG_MULT_HI = 0x2360ED051FC65DA4ULL;
static const uint64_t PCG_MULT_LO = 0x4385DF649FCCF645ULL;

static inline uint64_t plbo_pcg64_next64(plbo_pcg64 *g)
{
    unsigned __int128 s = ((unsigned __int128)g->state_hi << 64) | g->state_lo;
    const unsigned __int128 m = ((unsigned __int128)PCG_MULT_HI << 64) | PCG_MULT_LO;
    const unsigned __int128 inc = ((unsigned __int128)g->inc_hi << 64) | g->inc_lo;
    s = s * m + inc; /* step first, then output from the NEW state */
    g->state_hi = (uint64_t)(s >> 64);
    g->state_lo = (uint64_t)s;
    const uint64_t x = g->state_hi ^ g->state_lo;
    const unsigned rot = (unsigned)(g->state_hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}

static inline uint32_t plbo_pcg64_next32(plbo_pcg64 *g)
{
    if (g->has_uint32) {
        g->has_uint32 = 0;
        return g->uinteger;
    }
    const uint64_t v = plbo_pcg64_next64(g);
    g->has_uint32 = 1;
    g->uinteger = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

PLBO_API uint64_t plbo_pcg64_random_raw(plbo_pcg64 *g) { return plbo_pcg64_next64(g); }

/* Generator.integers(0, high, size=n, dtype=int64) for 1 <= high <= 2^32 - 1.
 * rng = high - 1 (inclusive range). */
PLBO_API int plbo_pcg64_integers(plbo_pcg64 *g, int64_t high, int64_t n, int64_t *out)
{
    if (high < 1 || high > 0xFFFFFFFFLL) return -1;
    const uint32_t rng = (uint32_t)(high - 1);
    if (rng == 0) {
        for (int64_t i = 0; i < n; ++i) out[i] = 0;
        return 0;
    }
    if (rng == 0xFFFFFFFFu) {
        for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)plbo_pcg64_next32(g);
        return 0;
    }
    const uint32_t rng_excl = rng + 1u;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t m = (uint64_t)plbo_pcg64_next32(g) * (uint64_t)rng_excl;
        uint32_t leftover = (uint32_t)m;
        if (leftover < rng_excl) {
            const uint32_t threshold = (uint32_t)((0xFFFFFFFFu - rng) % rng_excl);
            while (leftover < threshold) {
                m = (uint64_t)plbo_pcg64_next32(g) * (uint64_t)rng_excl;
                leftover = (uint32_t)m;
            }
        }
        out[i] = (int64_t)(m >> 32);
    }
    return 0;
}

/* parllel/replays/replay.py:54-70  index math of ReplayBuffer.sample_batch():
 * T draw, then B draw, wrap by (cursor + oldest_invalid) when the ring is full. */
PLBO_API int plbo_replay_sample_indices(
    plbo_pcg64 *g, int64_t size_T, int64_t B, int64_t cursor, int full,
    int64_t newest_invalid, int64_t oldest_invalid, int64_t n,
    int64_t *t_idx, int64_t *b_idx)
{
    int rc;
    if (full) {
        const int64_t offset = cursor + oldest_invalid;
        const int64_t valid_length = size_T - oldest_invalid - newest_invalid;
        rc = plbo_pcg64_integers(g, valid_length, n, t_idx);
        if (rc) return rc;
        for (int64_t i = 0; i < n; ++i) t_idx[i] = (t_idx[i] + offset) % size_T;
    } else {
        rc = plbo_pcg64_integers(g, cursor - newest_invalid, n, t_idx);
        if (rc) return rc;
    }
    return plbo_pcg64_integers(g, B, n, b_idx);
}
