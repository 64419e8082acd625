// plb_rng.cu -- device-side index generation for ReplayBuffer.sample_batch()
// (parllel/replays/replay.py:54-70): bit-exact restatement of
//     T_idxs = rng.integers(0, valid_length, size=(n,));  B_idxs = rng.integers(0, B, size=(n,))
// for numpy.random.Generator(PCG64), so that a minibatch needs no host->device index upload.
//
// numpy's algorithm (third-party dependency of the reference, pinned against numpy itself in
// tests/test_oracle_rng.py and tests/test_gpu_replay.py):
//   * PCG64: 128-bit LCG, state <- state * M + inc, output XSL-RR of the NEW state;
//   * next_uint32 hands out the low then the high half of one 64-bit output; the unused half stays
//     buffered in the bit generator across calls (has_uint32 / uinteger);
//   * integers() with range < 2^32 - 1 uses Lemire's multiply-shift with rejection on 32-bit words.
// Whether a raw 32-bit word is rejected depends only on the word itself, so the draw parallelises
// exactly: every thread jumps the LCG ahead to its own word (O(log) 128-bit multiplies, then a fixed
// stride), the accept flags are prefix-summed, and the first `n` accepted words are the result.
// One CTA; the generator state lives in device memory and is advanced in place.
#include "plb_common.cuh"

namespace plb {

typedef unsigned __int128 u128;

constexpr int RNG_THREADS = 1024;

__device__ __forceinline__ u128 pcg_mult()
{
    return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
}

// (mult, plus) of `delta` LCG steps: state_delta = mult * state + plus
__device__ __forceinline__ void lcg_skip(u128 inc, uint64_t delta, u128 &acc_mult, u128 &acc_plus)
{
    u128 cur_mult = pcg_mult(), cur_plus = inc;
    acc_mult = 1;
    acc_plus = 0;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
}

__device__ __forceinline__ uint64_t xsl_rr(u128 s)
{
    const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

struct RngShared {
    u128 state, inc;          // generator state at the start of the current draw
    int has;                  // buffered half-word available
    uint32_t buffered;        // the buffered half-word
    u128 stride_mult, stride_plus;  // RNG_THREADS/2 LCG steps
    int warp_sums[32];
    long long last_word;      // index of the word that completed the draw
};

// Exclusive prefix sum of a 0/1 flag over the CTA; returns the rank, writes the CTA total.
__device__ __forceinline__ int block_rank(int flag, RngShared &sh, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    const int in_warp = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) sh.warp_sums[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, all = 0;
    for (int w = 0; w < RNG_THREADS / 32; ++w) {
        const int c = sh.warp_sums[w];
        if (w < warp) before += c;
        all += c;
    }
    __syncthreads();
    total = all;
    return before + in_warp;
}

// One `integers(0, high, n)` call: out[i] = (draw_i + add) % mod  (mod == 0: no post-op).
__device__ void draw_integers(RngShared &sh, uint64_t high, long long n, int64_t *out, int64_t add, int64_t mod)
{
    if (n <= 0) return;
    const uint32_t rng = (uint32_t)(high - 1);
    if (rng == 0u) {  // nothing is drawn (numpy returns `low` without touching the generator)
        for (long long i = threadIdx.x; i < n; i += RNG_THREADS) out[i] = mod ? (add % mod) : add;
        return;
    }
    const bool full_range = (rng == 0xFFFFFFFFu);
    const uint32_t rng_excl = rng + 1u;
    const uint32_t threshold = full_range ? 0u : (uint32_t)((0xFFFFFFFFu - rng) % rng_excl);
    const u128 s0 = sh.state, inc = sh.inc;
    const int base = sh.has;  // word 0 is the buffered half-word when present
    __syncthreads();

    // this thread's first word j = tid: output index q = (j - base) >> 1, half = (j - base) & 1
    u128 st = 0;
    {
        const long long j = threadIdx.x;
        if (j >= base) {
            u128 m, c;
            lcg_skip(inc, (uint64_t)(((j - base) >> 1) + 1), m, c);
            st = m * s0 + c;
        }
    }
    long long filled = 0, pos = 0;
    while (true) {
        const long long j = pos + threadIdx.x;
        uint32_t w;
        if (j < base) w = sh.buffered;
        else {
            const uint64_t o = xsl_rr(st);
            w = ((j - base) & 1) ? (uint32_t)(o >> 32) : (uint32_t)o;
        }
        uint64_t m = (uint64_t)w * (uint64_t)rng_excl;
        int accept = 1;
        if (!full_range) {
            const uint32_t leftover = (uint32_t)m;
            accept = !(leftover < rng_excl && leftover < threshold);
        } else {
            m = (uint64_t)w << 32;
        }
        int total;
        const int rank = block_rank(accept, sh, total);
        const long long need = n - filled;
        if (accept && rank < need) {
            int64_t v = (int64_t)(m >> 32);
            if (mod) v = (v + add) % mod;
            out[filled + rank] = v;
            if (rank == need - 1) sh.last_word = j;
        }
        if (total >= need) break;
        filled += total;
        pos += RNG_THREADS;
        if (pos == RNG_THREADS && threadIdx.x < base) {
            // the thread that served the buffered word joins the stream: word j = pos + tid
            u128 mm, cc;
            lcg_skip(inc, (uint64_t)(((pos + threadIdx.x - base) >> 1) + 1), mm, cc);
            st = mm * s0 + cc;
        } else {
            st = sh.stride_mult * st + sh.stride_plus;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // words consumed from fresh outputs, outputs consumed, new buffered half-word
        const long long words = sh.last_word + 1 - base;  // >= 0
        if (words <= 0) {
            sh.has = 0;  // only the buffered word was used
        } else {
            const uint64_t outs = (uint64_t)((words + 1) >> 1);
            u128 m, c;
            lcg_skip(inc, outs, m, c);
            const u128 ns = m * s0 + c;
            sh.state = ns;
            if (words & 1) {
                sh.has = 1;
                sh.buffered = (uint32_t)(xsl_rr(ns) >> 32);
            } else {
                sh.has = 0;
            }
        }
    }
    __syncthreads();
}

struct RngParams {
    uint64_t *state;  // [0..3] state_hi, state_lo, inc_hi, inc_lo; [4] = has_uint32 | (uinteger << 32)
    int64_t size_T, B, cursor;
    int full;
    int64_t newest_invalid, oldest_invalid;
    long long n, k;
    int64_t *t_idx, *b_idx;
};

__global__ void __launch_bounds__(RNG_THREADS)
replay_indices_kernel(const RngParams p)
{
    __shared__ RngShared sh;
    if (threadIdx.x == 0) {
        sh.state = ((u128)p.state[0] << 64) | (u128)p.state[1];
        sh.inc = ((u128)p.state[2] << 64) | (u128)p.state[3];
        sh.has = (int)(p.state[4] & 1ull);
        sh.buffered = (uint32_t)(p.state[4] >> 32);
        u128 m, c;
        lcg_skip(sh.inc, RNG_THREADS / 2, m, c);
        sh.stride_mult = m;
        sh.stride_plus = c;
    }
    __syncthreads();
    // replay.py:55-70
    uint64_t high_t;
    int64_t add = 0, mod = 0;
    if (p.full) {
        high_t = (uint64_t)(p.size_T - p.oldest_invalid - p.newest_invalid);
        add = p.cursor + p.oldest_invalid;
        mod = p.size_T;
    } else {
        high_t = (uint64_t)(p.cursor - p.newest_invalid);
    }
    for (long long kk = 0; kk < p.k; ++kk) {
        draw_integers(sh, high_t, p.n, p.t_idx + kk * p.n, add, mod);
        draw_integers(sh, (uint64_t)p.B, p.n, p.b_idx + kk * p.n, 0, 0);
    }
    if (threadIdx.x == 0) {
        p.state[0] = (uint64_t)(sh.state >> 64);
        p.state[1] = (uint64_t)sh.state;
        p.state[4] = (uint64_t)(sh.has & 1) | ((uint64_t)sh.buffered << 32);
    }
}

}  // namespace plb

extern "C" int plb_replay_sample_indices(uint64_t *pcg64_state, int64_t size_T, int64_t B, int64_t cursor, int full,
                                         int64_t newest_invalid, int64_t oldest_invalid, int64_t n, int64_t k,
                                         int64_t *t_idx, int64_t *b_idx, plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(pcg64_state && t_idx && b_idx, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(n >= 0 && k >= 0, PLB_ERR_INVALID_ARGUMENT, "negative count");
    if (n == 0 || k == 0) return PLB_OK;
    const int64_t high_t = full ? size_T - oldest_invalid - newest_invalid : cursor - newest_invalid;
    PLB_REQUIRE(high_t >= 1 && B >= 1, PLB_ERR_INVALID_ARGUMENT,
                "empty sampling range (valid_length %lld, B %lld)", (long long)high_t, (long long)B);
    PLB_REQUIRE(high_t <= 0x100000000LL && B <= 0x100000000LL && size_T >= 1, PLB_ERR_UNSUPPORTED,
                "index ranges above 2^32 take numpy's 64-bit path, which is not implemented");
    RngParams p;
    p.state = pcg64_state;
    p.size_T = size_T; p.B = B; p.cursor = cursor; p.full = full;
    p.newest_invalid = newest_invalid; p.oldest_invalid = oldest_invalid;
    p.n = n; p.k = k;
    p.t_idx = t_idx; p.b_idx = b_idx;
    replay_indices_kernel<<<1, RNG_THREADS, 0, (cudaStream_t)stream>>>(p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}
