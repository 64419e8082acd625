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
// The generator state lives in device memory and is advanced in place.
//
// Two kernels: the usual ranges reject a word with probability < 2^-14, and there the whole draw is
// speculated as one stream by a CLUSTER of eight 1024-thread CTAs (32768 words per round; the first
// rejected word of a round is agreed through distributed shared memory, everything before it is final).
// Every LCG jump comes from compile-time tables (plb_rng_tables.inc: M^d and the geometric sums G(d) do
// not depend on the generator), so a launch costs a handful of 128-bit multiplies per thread instead of
// O(log d) of them on a serial critical path.  Ranges just above a power of two (up to half of all words
// rejected) take the single-CTA compaction kernel.
#include <cooperative_groups.h>

#include "plb_common.cuh"

namespace plb {

namespace cg = cooperative_groups;

typedef unsigned __int128 u128;

#include "plb_rng_tables.inc"

struct Jump {
    u128 m, g;  // s_d = m * s + inc * g
};
__device__ __forceinline__ Jump table_jump(const uint64_t (&e)[4])
{
    Jump j;
    j.m = ((u128)e[0] << 64) | (u128)e[1];
    j.g = ((u128)e[2] << 64) | (u128)e[3];
    return j;
}
// `a` steps, then `b` steps
__device__ __forceinline__ Jump compose(const Jump &a, const Jump &b)
{
    Jump r;
    r.m = a.m * b.m;
    r.g = b.m * a.g + b.g;
    return r;
}
__device__ __forceinline__ u128 apply_jump(const Jump &j, u128 s, u128 inc) { return j.m * s + inc * j.g; }
// arbitrary distance by binary decomposition (rare path: after a rejected word)
__device__ __forceinline__ u128 jump_from(u128 s, u128 inc, uint64_t delta)
{
    for (int i = 0; delta != 0; ++i, delta >>= 1)
        if (delta & 1) s = apply_jump(table_jump(RNG_POW2[i]), s, inc);
    return s;
}

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

constexpr int WORDS_PER_THREAD = 4;                       // two consecutive 64-bit outputs per thread
constexpr int WORDS_PER_ITER = RNG_THREADS * WORDS_PER_THREAD;
constexpr int OUTS_PER_ITER = WORDS_PER_ITER / 2;

struct RngShared {
    u128 state, inc;          // generator state at the start of the current draw
    int has;                  // buffered half-word available
    uint32_t buffered;        // the buffered half-word
    u128 stride_mult, stride_plus;  // OUTS_PER_ITER LCG steps
    int warp_prefix[32];
    int total;
    long long last_word;      // index (in fresh words) of the word that completed the draw
};

// Exclusive prefix sum of per-thread counts over the CTA; returns this thread's offset, CTA total in sh.total.
__device__ __forceinline__ int block_exclusive_scan(int count, RngShared &sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) sh.warp_prefix[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = sh.warp_prefix[lane];
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        sh.warp_prefix[lane] = wi - w;  // exclusive prefix of the warp totals
        if (lane == 31) sh.total = wi;
    }
    __syncthreads();
    const int off = sh.warp_prefix[warp] + incl - count;
    return off;
}

// One `integers(0, high, n)` call: out[i] = (draw_i + add) % mod  (mod == 0: no post-op).
// `skip_mult/skip_plus`: this thread's precomputed LCG jump to its first output (2*tid + 1 steps).
__device__ void draw_integers(RngShared &sh, uint64_t high, long long n, int64_t *out, int64_t add, int64_t mod,
                              u128 skip_mult, u128 skip_plus)
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
    auto lemire = [&](uint32_t w, int64_t &v) -> bool {
        uint64_t m = (uint64_t)w * (uint64_t)rng_excl;
        if (full_range) { v = (int64_t)w; return true; }
        v = (int64_t)(m >> 32);
        return (uint32_t)m >= threshold;
    };
    auto post = [&](int64_t v) -> int64_t { return mod ? (v + add) % mod : v; };

    long long filled = 0;
    // the buffered half-word (if any) is the first word of the stream
    if (sh.has) {
        int64_t v;
        const bool ok = lemire(sh.buffered, v);
        __syncthreads();
        if (threadIdx.x == 0) {
            sh.has = 0;
            if (ok) out[0] = post(v);
        }
        if (ok) filled = 1;
        __syncthreads();
        if (filled >= n) return;
    }
    const u128 s0 = sh.state, inc = sh.inc;
    // fresh words: word f (0-based) is the low (f even) / high (f odd) half of output f >> 1; this thread
    // serves words 4*tid .. 4*tid+3 of every 4096-word block = outputs 2*tid, 2*tid+1
    u128 st = skip_mult * s0 + skip_plus;  // state after 2*tid + 1 steps
    long long pos = 0;                     // fresh words consumed by earlier blocks
    while (true) {
        const u128 st2 = st * pcg_mult() + inc;
        const uint64_t o0 = xsl_rr(st), o1 = xsl_rr(st2);
        const uint32_t w[4] = {(uint32_t)o0, (uint32_t)(o0 >> 32), (uint32_t)o1, (uint32_t)(o1 >> 32)};
        int64_t v[4];
        bool ok[4];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ok[k] = lemire(w[k], v[k]);
            cnt += ok[k] ? 1 : 0;
        }
        int rank = block_exclusive_scan(cnt, sh);
        const int total = sh.total;
        const long long need = n - filled;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (ok[k]) {
                if (rank < need) {
                    out[filled + rank] = post(v[k]);
                    if (rank == need - 1) sh.last_word = pos + 4LL * threadIdx.x + k;
                }
                ++rank;
            }
        }
        if (total >= need) break;
        filled += total;
        pos += WORDS_PER_ITER;
        st = sh.stride_mult * st + sh.stride_plus;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long words = sh.last_word + 1;            // fresh words consumed
        const uint64_t outs = (uint64_t)((words + 1) >> 1);  // 64-bit outputs consumed
        u128 m, c;
        lcg_skip(inc, outs, m, c);
        const u128 ns = m * s0 + c;
        sh.state = ns;
        if (words & 1) {  // the high half of the last output stays buffered
            sh.has = 1;
            sh.buffered = (uint32_t)(xsl_rr(ns) >> 32);
        } else {
            sh.has = 0;
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

// ---------------------------------------------------------------------------------------------
// Fast path: all 2k draws of a launch as ONE stream.  With the usual ranges a word is rejected with
// probability < 2^-16, so the kernel speculates that every word of a 4096-word block is accepted:
// word f then lands at output position g0 + (f - pos), which fixes its segment (minibatch, T or B)
// and therefore its range.  The first rejected word of the block (if any) is found with a block-wide
// minimum; everything before it is final, and the next block restarts right behind it (threads
// re-jump the LCG only then).  Segments with a range of 1 draw nothing and are filled separately.
// ---------------------------------------------------------------------------------------------
struct StreamSpec {
    uint32_t excl[2], thr[2];   // range (exclusive) and Lemire threshold of the T / B draws
    int full_range[2];          // 1: range == 2^32, raw words
    int draws[2];               // 0: the draw consumes no words (range 1)
    int64_t add, mod;           // post-op of the T draw
};

__device__ void draw_stream(RngShared &sh, const RngParams &p, const StreamSpec &sp)
{
    __shared__ long long s_first_reject;
    const long long n = p.n;
    const int n_drawing = sp.draws[0] + sp.draws[1];           // word-consuming segments per minibatch
    const long long total_out = (long long)n_drawing * n * p.k;  // outputs that consume words
    // segments that consume nothing
    for (int which = 0; which < 2; ++which) {
        if (!sp.draws[which]) {
            int64_t *dst = which == 0 ? p.t_idx : p.b_idx;
            const int64_t v = (which == 0 && sp.mod) ? (sp.add % sp.mod) : (which == 0 ? sp.add : 0);
            for (long long i = threadIdx.x; i < n * p.k; i += RNG_THREADS) dst[i] = v;
        }
    }
    if (total_out == 0) return;
    // position bookkeeping without per-word divisions: a cursor (minibatch, segment kind, index) is
    // computed once per thread and block and stepped word by word
    struct Cursor {
        long long mb, idx;
        int which;
    };
    auto locate = [&](long long g) -> Cursor {
        Cursor c;
        const long long seg = g / n;
        c.idx = g - seg * n;
        c.mb = seg / n_drawing;
        c.which = (n_drawing == 2) ? (int)(seg & 1) : (sp.draws[0] ? 0 : 1);
        return c;
    };
    auto step = [&](Cursor &c) {
        if (++c.idx == n) {
            c.idx = 0;
            if (n_drawing == 2) {
                if (c.which == 1) ++c.mb;
                c.which ^= 1;
            } else {
                ++c.mb;
            }
        }
    };
    auto rejected = [&](const Cursor &c, uint32_t w) -> bool {
        if (sp.full_range[c.which]) return false;
        return (uint32_t)((uint64_t)w * (uint64_t)sp.excl[c.which]) < sp.thr[c.which];
    };
    auto emit = [&](const Cursor &c, uint32_t w) {
        int64_t v = sp.full_range[c.which] ? (int64_t)w : (int64_t)(((uint64_t)w * (uint64_t)sp.excl[c.which]) >> 32);
        if (c.which == 0) {
            if (sp.mod) {  // (v + add) % size_T with v < size_T and add < 2 * size_T
                v += sp.add;
                if (v >= sp.mod) v -= sp.mod;
                if (v >= sp.mod) v -= sp.mod;
            }
            p.t_idx[c.mb * n + c.idx] = v;
        } else {
            p.b_idx[c.mb * n + c.idx] = v;
        }
    };

    long long g0 = 0;  // outputs produced so far
    if (sh.has) {      // the buffered half-word is the first word of the stream
        const uint32_t w = sh.buffered;
        __syncthreads();
        if (threadIdx.x == 0) {
            sh.has = 0;
            const Cursor c0 = locate(0);
            const bool ok = !rejected(c0, w);
            if (ok) emit(c0, w);
            sh.total = ok ? 1 : 0;
        }
        __syncthreads();
        g0 = sh.total;
        if (g0 >= total_out) return;
    }
    const u128 s0 = sh.state, inc = sh.inc;
    long long pos = 0;  // fresh words consumed so far (word f = half f&1 of output f>>1)
    bool rejump = true;
    u128 st = 0;
    while (true) {
        const long long f0 = pos + 4LL * threadIdx.x;  // this thread's first word of the block
        if (rejump) {
            u128 m, c;
            lcg_skip(inc, (uint64_t)(f0 >> 1) + 1ull, m, c);
            st = m * s0 + c;
            rejump = false;
        }
        const u128 st2 = st * pcg_mult() + inc;
        const u128 st3 = st2 * pcg_mult() + inc;
        const uint64_t o0 = xsl_rr(st), o1 = xsl_rr(st2), o2 = xsl_rr(st3);
        uint32_t w[4];
        if ((f0 & 1) == 0) {
            w[0] = (uint32_t)o0; w[1] = (uint32_t)(o0 >> 32); w[2] = (uint32_t)o1; w[3] = (uint32_t)(o1 >> 32);
        } else {
            w[0] = (uint32_t)(o0 >> 32); w[1] = (uint32_t)o1; w[2] = (uint32_t)(o1 >> 32); w[3] = (uint32_t)o2;
        }
        if (threadIdx.x == 0) s_first_reject = 1LL << 62;
        __syncthreads();
        // pass 1: find the first word of the block that would be rejected at its speculative position
        long long my_reject = 1LL << 62;
        const long long gbase = g0 + (f0 - pos);
        const Cursor cur0 = locate(gbase < total_out ? gbase : 0);
        {
            Cursor c = cur0;
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {
                if (gbase + kq < total_out) {
                    if (my_reject == (1LL << 62) && rejected(c, w[kq])) my_reject = f0 + kq;
                    step(c);
                }
            }
        }
        // warp minimum, then one atomicMin per warp that saw a rejection (rare)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const long long o = __shfl_xor_sync(0xffffffffu, my_reject, d);
            my_reject = o < my_reject ? o : my_reject;
        }
        if ((threadIdx.x & 31) == 0 && my_reject < (1LL << 62))
            atomicMin(reinterpret_cast<unsigned long long *>(&s_first_reject), (unsigned long long)my_reject);
        __syncthreads();
        const long long first_reject = s_first_reject;
        // pass 2: everything before the first rejection (and inside the job) is final
        {
            Cursor c = cur0;
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {
                const long long f = f0 + kq;
                if (f < first_reject && gbase + kq < total_out) {
                    emit(c, w[kq]);
                    step(c);
                }
            }
        }
        const long long block_end = pos + WORDS_PER_ITER;
        const long long stop = first_reject < block_end ? first_reject : block_end;  // words [pos, stop) accepted
        const long long produced = stop - pos;
        if (g0 + produced >= total_out) {  // done inside this block: the last consumed word
            pos = pos + (total_out - g0);
            break;
        }
        g0 += produced;
        if (first_reject < block_end) {
            pos = first_reject + 1;  // the rejected word is consumed, nothing is produced for it
            rejump = true;
        } else {
            pos = block_end;
            st = sh.stride_mult * st + sh.stride_plus;
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long words = pos;                         // fresh words consumed
        const uint64_t outs = (uint64_t)((words + 1) >> 1);  // 64-bit outputs consumed
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
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Cluster version of draw_stream: eight CTAs speculate 32768 consecutive words per round.
// ---------------------------------------------------------------------------------------------
constexpr int CL_CTAS = 8;
constexpr int CL_WORDS_PER_ROUND = CL_CTAS * RNG_THREADS * WORDS_PER_THREAD;  // 32768
constexpr long long NO_REJECT = 1LL << 62;

__global__ void __cluster_dims__(CL_CTAS, 1, 1) __launch_bounds__(RNG_THREADS)
replay_indices_cluster_kernel(const RngParams p, const StreamSpec sp)
{
    __shared__ unsigned long long s_reject[2];  // first rejected word of the round, this CTA (by round parity)
    __shared__ long long s_first;               // ... of the whole cluster
    __shared__ u128 s_base;                     // generator state at output index pos >> 1 (after a restart)
    griddep_launch_dependents();
    griddep_wait();  // the generator state is whatever the previous draw left
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long gtid = (long long)cta * RNG_THREADS + tid;
    const long long nthreads = (long long)CL_CTAS * RNG_THREADS;

    const u128 s0 = ((u128)p.state[0] << 64) | (u128)p.state[1];
    const u128 inc = ((u128)p.state[2] << 64) | (u128)p.state[3];
    const int has = (int)(p.state[4] & 1ull);
    const uint32_t buffered = (uint32_t)(p.state[4] >> 32);
    if (tid == 0) { s_reject[0] = (unsigned long long)NO_REJECT; s_reject[1] = (unsigned long long)NO_REJECT; }

    const long long n = p.n;
    const int n_drawing = sp.draws[0] + sp.draws[1];
    const long long total_out = (long long)n_drawing * n * p.k;
    for (int which = 0; which < 2; ++which) {  // segments that consume no words (range 1)
        if (!sp.draws[which]) {
            int64_t *dst = which == 0 ? p.t_idx : p.b_idx;
            const int64_t v = (which == 0 && sp.mod) ? (sp.add % sp.mod) : (which == 0 ? sp.add : 0);
            for (long long i = gtid; i < n * p.k; i += nthreads) dst[i] = v;
        }
    }
    struct Cursor {
        long long mb, idx;
        int which;
    };
    auto locate = [&](long long g) -> Cursor {
        Cursor c;
        const long long seg = g / n;
        c.idx = g - seg * n;
        c.mb = seg / n_drawing;
        c.which = (n_drawing == 2) ? (int)(seg & 1) : (sp.draws[0] ? 0 : 1);
        return c;
    };
    auto step = [&](Cursor &c) {
        if (++c.idx == n) {
            c.idx = 0;
            if (n_drawing == 2) {
                if (c.which == 1) ++c.mb;
                c.which ^= 1;
            } else {
                ++c.mb;
            }
        }
    };
    auto rejected = [&](const Cursor &c, uint32_t w) -> bool {
        if (sp.full_range[c.which]) return false;
        return (uint32_t)((uint64_t)w * (uint64_t)sp.excl[c.which]) < sp.thr[c.which];
    };
    auto emit = [&](const Cursor &c, uint32_t w) {
        int64_t v = sp.full_range[c.which] ? (int64_t)w : (int64_t)(((uint64_t)w * (uint64_t)sp.excl[c.which]) >> 32);
        if (c.which == 0) {
            if (sp.mod) {  // (v + add) % size_T with v < size_T and add < 2 * size_T
                v += sp.add;
                if (v >= sp.mod) v -= sp.mod;
                if (v >= sp.mod) v -= sp.mod;
            }
            p.t_idx[c.mb * n + c.idx] = v;
        } else {
            p.b_idx[c.mb * n + c.idx] = v;
        }
    };
    auto finish = [&](u128 ns, long long words) {  // exactly one thread of the cluster calls this
        p.state[0] = (uint64_t)(ns >> 64);
        p.state[1] = (uint64_t)ns;
        const int h = (int)(words & 1);  // the high half of the last output stays buffered
        p.state[4] = (uint64_t)h | ((uint64_t)(h ? (uint32_t)(xsl_rr(ns) >> 32) : 0u) << 32);
    };

    long long g0 = 0;  // outputs produced so far
    bool done = total_out == 0;
    if (!done && has) {  // the buffered half-word is the first word of the stream (every thread evaluates it alike)
        const Cursor c0 = locate(0);
        const bool ok = !rejected(c0, buffered);
        if (ok && gtid == 0) emit(c0, buffered);
        g0 = ok ? 1 : 0;
        if (g0 >= total_out) {
            if (gtid == 0) finish(s0, 0);
            done = true;
        }
    }
    if (done) {  // nothing (more) to draw: uniform over the cluster, so leaving before any barrier is safe
        return;
    }
    // this thread's offset inside a round: outputs 2 * gtid and the two after it
    const Jump mine = compose(table_jump(RNG_WARP[cta * 32 + warp]), table_jump(RNG_LANE[lane]));
    const Jump stride = table_jump(RNG_POW2[14]);  // CL_WORDS_PER_ROUND / 2 = 16384 outputs per round
    static_assert(CL_WORDS_PER_ROUND / 2 == (1 << 14), "round stride must match RNG_POW2[14]");
    u128 st = apply_jump(mine, s0, inc);
    long long pos = 0;  // fresh words consumed so far (word f = half f&1 of output f>>1)
    int par = 0;
    __syncthreads();
    while (true) {
        const long long f0 = pos + 4LL * gtid;  // this thread's first word of the round
        const u128 st2 = st * pcg_mult() + inc;
        const u128 st3 = st2 * pcg_mult() + inc;
        const uint64_t o0 = xsl_rr(st), o1 = xsl_rr(st2), o2 = xsl_rr(st3);
        uint32_t w[4];
        if ((f0 & 1) == 0) {
            w[0] = (uint32_t)o0; w[1] = (uint32_t)(o0 >> 32); w[2] = (uint32_t)o1; w[3] = (uint32_t)(o1 >> 32);
        } else {
            w[0] = (uint32_t)(o0 >> 32); w[1] = (uint32_t)o1; w[2] = (uint32_t)(o1 >> 32); w[3] = (uint32_t)o2;
        }
        // pass 1: the first word of the round that would be rejected at its speculative position
        long long my_reject = NO_REJECT;
        const long long gbase = g0 + (f0 - pos);
        const Cursor cur0 = locate(gbase < total_out ? gbase : 0);
        {
            Cursor c = cur0;
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {
                if (gbase + kq < total_out) {
                    if (my_reject == NO_REJECT && rejected(c, w[kq])) my_reject = f0 + kq;
                    step(c);
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const long long o = __shfl_xor_sync(0xffffffffu, my_reject, d);
            my_reject = o < my_reject ? o : my_reject;
        }
        if (lane == 0 && my_reject < NO_REJECT) atomicMin(&s_reject[par], (unsigned long long)my_reject);
        cluster.sync();  // every CTA's minimum of this round is final
        if (warp == 0) {
            long long r = NO_REJECT;
            if (lane < CL_CTAS) r = (long long)*cluster.map_shared_rank(&s_reject[par], lane);
#pragma unroll
            for (int d = 4; d > 0; d >>= 1) {
                const long long o = __shfl_xor_sync(0xffffffffu, r, d);
                r = o < r ? o : r;
            }
            if (lane == 0) {
                s_first = r;
                // last round's slot: every CTA passed this round's barrier, so nobody reads it any more
                s_reject[par ^ 1] = (unsigned long long)NO_REJECT;
            }
        }
        __syncthreads();
        const long long first_reject = s_first;
        // pass 2: everything before the first rejection (and inside the job) is final
        {
            Cursor c = cur0;
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {
                const long long f = f0 + kq;
                if (f < first_reject && gbase + kq < total_out) {
                    emit(c, w[kq]);
                    step(c);
                }
            }
        }
        const long long round_end = pos + CL_WORDS_PER_ROUND;
        const long long stop = first_reject < round_end ? first_reject : round_end;  // words [pos, stop) accepted
        const long long produced = stop - pos;
        if (g0 + produced >= total_out) {  // the draw ends inside this round
            const long long words = pos + (total_out - g0);  // fresh words consumed by the whole call
            const long long last = words - 1;
            if (last >= f0 && last < f0 + 4) {
                // new state = the state that produced the last consumed output
                const long long k_out = (last >> 1) - (f0 >> 1);
                finish(k_out == 0 ? st : (k_out == 1 ? st2 : st3), words);
            }
            break;
        }
        g0 += produced;
        par ^= 1;
        if (first_reject < round_end) {
            pos = first_reject + 1;  // the rejected word is consumed, nothing is produced for it
            __syncthreads();         // s_base of an earlier restart has been read by everyone
            if (tid == 0) s_base = jump_from(s0, inc, (uint64_t)(pos >> 1));
            __syncthreads();
            st = apply_jump(mine, s_base, inc);
        } else {
            pos = round_end;
            st = apply_jump(stride, st, inc);
        }
    }
    cluster.sync();  // no CTA may retire while a sibling can still read its shared memory
}

__global__ void __launch_bounds__(RNG_THREADS)
replay_indices_kernel(const RngParams p)
{
    __shared__ RngShared sh;
    griddep_launch_dependents();
    griddep_wait();
    if (threadIdx.x == 0) {
        sh.state = ((u128)p.state[0] << 64) | (u128)p.state[1];
        sh.inc = ((u128)p.state[2] << 64) | (u128)p.state[3];
        sh.has = (int)(p.state[4] & 1ull);
        sh.buffered = (uint32_t)(p.state[4] >> 32);
        u128 m, c;
        lcg_skip(sh.inc, OUTS_PER_ITER, m, c);
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
    // rejection probabilities of the two draws decide between the speculative stream and the
    // compaction path (ranges just above a power of two reject up to half of all words)
    StreamSpec sp;
    const uint64_t highs[2] = {high_t, (uint64_t)p.B};
    bool rare = true;
    for (int i = 0; i < 2; ++i) {
        const uint32_t rng = (uint32_t)(highs[i] - 1);
        sp.draws[i] = rng != 0u;
        sp.full_range[i] = rng == 0xFFFFFFFFu;
        sp.excl[i] = rng + 1u;
        sp.thr[i] = (sp.full_range[i] || rng == 0u) ? 0u : (uint32_t)((0xFFFFFFFFu - rng) % sp.excl[i]);
        if (sp.thr[i] > (1u << 18)) rare = false;  // > 2^-14 per word: > 1/4 expected rejection per block
    }
    sp.add = add; sp.mod = mod;
    if (rare) {
        draw_stream(sh, p, sp);
    } else {
        u128 skip_mult, skip_plus;
        lcg_skip(sh.inc, 2ull * threadIdx.x + 1ull, skip_mult, skip_plus);
        for (long long kk = 0; kk < p.k; ++kk) {
            draw_integers(sh, high_t, p.n, p.t_idx + kk * p.n, add, mod, skip_mult, skip_plus);
            draw_integers(sh, (uint64_t)p.B, p.n, p.b_idx + kk * p.n, 0, 0, skip_mult, skip_plus);
        }
    }
    if (threadIdx.x == 0) {
        p.state[0] = (uint64_t)(sh.state >> 64);
        p.state[1] = (uint64_t)sh.state;
        p.state[4] = (uint64_t)(sh.has & 1) | ((uint64_t)sh.buffered << 32);
    }
}

static int g_rng_cluster = 1;  // 0: always the single-CTA kernel (plb_set_tuning key 13)
void set_rng_cluster(int v) { g_rng_cluster = v; }

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
    // rejection rates of the two draws (same arithmetic as numpy's Lemire thresholds)
    StreamSpec sp;
    const uint64_t highs[2] = {(uint64_t)high_t, (uint64_t)B};
    bool rare = true;
    for (int i = 0; i < 2; ++i) {
        const uint32_t rng = (uint32_t)(highs[i] - 1);
        sp.draws[i] = rng != 0u;
        sp.full_range[i] = rng == 0xFFFFFFFFu;
        sp.excl[i] = rng + 1u;
        sp.thr[i] = (sp.full_range[i] || rng == 0u) ? 0u : (uint32_t)((0xFFFFFFFFu - rng) % sp.excl[i]);
        if (sp.thr[i] > (1u << 18)) rare = false;  // > 2^-14 per word
    }
    sp.add = full ? cursor + oldest_invalid : 0;
    sp.mod = full ? size_T : 0;
    if (rare && g_rng_cluster) {
        PLB_CUDA(launch_pdl(replay_indices_cluster_kernel, dim3(CL_CTAS), dim3(RNG_THREADS), 0, (cudaStream_t)stream, p, sp));
    } else {
        PLB_CUDA(launch_pdl(replay_indices_kernel, dim3(1), dim3(RNG_THREADS), 0, (cudaStream_t)stream, p));
    }
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}
