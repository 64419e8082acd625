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
constexpr int TC = 64;       // rows per tile
constexpr int NWARPS = 8;    // warps per CTA
constexpr int R = TC / NWARPS;
constexpr int THREADS = NWARPS * 32;
constexpr int MAX_STAGES = 8;

// one pipeline stage: three TMA boxes
struct __align__(128) Stage {
    float r[TC][W];
    float v[TC][W];
    uint8_t d[TC][W];
};
constexpr uint32_t STAGE_TX_REVERSE = sizeof(float) * TC * W * 2 + TC * W;  // r + v + d
constexpr uint32_t STAGE_TX_FORWARD = sizeof(float) * TC * W + TC * W;      // r + d

struct Shared {
    double aggA[2][NWARPS][W];
    double aggC[2][NWARPS][W];
    float halo[2][W];
    unsigned long long full[MAX_STAGES];
};

struct Params {
    const float *boot;            // reverse: bootstrap value [B]
    const float *prev_ret;        // forward: previous return [B]
    const uint8_t *prev_done;     // forward: previous done [B]
    float *out0; int64_t ld0;     // reverse: advantage   | forward: past_return
    float *out1; int64_t ld1;     // reverse: return_
    int64_t T, B;
    double gamma, lambda;
    int n_stages;
    int n_strips, n_chunks;
    // optional fused reward transform (NormalizeRewards scale + ClipRewards) on the load path
    const double *reward_var;     // device float64 running variance, or NULL
    double reward_eps;
    float clip_lo, clip_hi;       // -inf / +inf when absent
    int has_pre;
    float *reward_out; int64_t ld_reward_out;  // transformed reward written back (may be NULL)
    // optional fused moments of out0
    const uint8_t *valid; int64_t ld_valid;
    MomentsTail tail;             // moments_out == NULL -> off
};

// MODE 0: GAE, 1: discounted return (both reverse), 2: past return (forward)
template <int MODE>
__global__ void __launch_bounds__(THREADS)
scan_chunked_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_v,
                    const __grid_constant__ CUtensorMap tm_d, const Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage *stages = reinterpret_cast<Stage *>(smem_raw);
    Shared *sh = reinterpret_cast<Shared *>(smem_raw + sizeof(Stage) * p.n_stages);
    __shared__ Moments red_scratch[32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool REVERSE = (MODE != 2);
    const uint32_t tx_bytes = REVERSE ? STAGE_TX_REVERSE : STAGE_TX_FORWARD;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_stages; ++s) mbar_init(&sh->full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // this CTA's item sequence: strips blockIdx.x, +gridDim.x, ...; chunks in scan order
    const int my_strips = (p.n_strips - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int64_t Q = (int64_t)my_strips * p.n_chunks;

    auto issue = [&](int64_t q) {
        const int s = (int)(q % p.n_stages);
        const int strip = (int)blockIdx.x + (int)(q / p.n_chunks) * (int)gridDim.x;
        const int k = (int)(q % p.n_chunks);
        const int c0 = strip * W;
        int t0;
        if (REVERSE) t0 = (int)p.T - (k + 1) * TC;   // tiles aligned to the END of the batch
        else t0 = k * TC;                            // tiles aligned to the START
        mbar_arrive_expect_tx(&sh->full[s], tx_bytes);
        tma_load_2d(&stages[s].r[0][0], &tm_r, c0, t0, &sh->full[s]);
        if (REVERSE) {
            tma_load_2d(&stages[s].v[0][0], &tm_v, c0, t0, &sh->full[s]);
            tma_load_2d(&stages[s].d[0][0], &tm_d, c0, t0, &sh->full[s]);
        } else {
            tma_load_2d(&stages[s].d[0][0], &tm_d, c0, t0 - 1, &sh->full[s]);  // done[t-1]
        }
    };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_r);
        if (REVERSE) tma_prefetch_desc(&tm_v);
        tma_prefetch_desc(&tm_d);
        for (int64_t q = 0; q < Q && q < p.n_stages; ++q) issue(q);
    }

    // fused reward transform constants
    double inv_denom = 1.0;
    if (p.has_pre && p.reward_var != nullptr) inv_denom = 1.0 / sqrt(*p.reward_var + p.reward_eps);
    const double gl = p.gamma * p.lambda;

    PivotAcc acc;
    acc.init();

    double carry = 0.0;      // scan value just outside the tile (strip-level carry), per lane
    for (int64_t q = 0; q < Q; ++q) {
        const int s = (int)(q % p.n_stages);
        const uint32_t parity = (uint32_t)((q / p.n_stages) & 1);
        const int strip = (int)blockIdx.x + (int)(q / p.n_chunks) * (int)gridDim.x;
        const int k = (int)(q % p.n_chunks);
        const int64_t col = (int64_t)strip * W + lane;
        const bool col_ok = col < p.B;
        const int par = (int)(q & 1);
        int64_t t0;
        if (REVERSE) t0 = p.T - (int64_t)(k + 1) * TC;
        else t0 = (int64_t)k * TC;
        const Stage &S = stages[s];

        mbar_wait(&sh->full[s], parity);

        double A[R], C[R];
        float vreg[R];
        (void)vreg;

        if (REVERSE) {
            if (k == 0) carry = (MODE == 1 && col_ok) ? (double)p.boot[col] : 0.0;
            const int i_hi = (warp + 1) * R - 1;
            // value[t+1] for the warp's last row
            double vnext;
            if (MODE == 0) {
                if (i_hi + 1 < TC) vnext = (double)S.v[i_hi + 1][lane];
                else if (k == 0) vnext = col_ok ? (double)p.boot[col] : 0.0;
                else vnext = (double)sh->halo[par ^ 1][lane];
            }
            double a = 0.0, c = 1.0;
#pragma unroll
            for (int u = R - 1; u >= 0; --u) {
                const int i = warp * R + u;
                float rf = S.r[i][lane];
                const float vf = S.v[i][lane];
                const bool dn = S.d[i][lane] != 0;
                if (p.has_pre) {
                    rf = __double2float_rn((double)rf * inv_denom);
                    rf = fminf(fmaxf(rf, p.clip_lo), p.clip_hi);
                    const int64_t t = t0 + i;
                    if (p.reward_out != nullptr && t >= 0 && col_ok) p.reward_out[t * p.ld_reward_out + col] = rf;
                }
                vreg[u] = vf;
                if (MODE == 0) {
                    const double v = (double)vf;
                    const double gnd = dn ? 0.0 : p.gamma;
                    const double delta = fma(gnd, vnext, (double)rf) - v;
                    const double cc = dn ? 0.0 : gl;
                    a = fma(cc, a, delta);
                    c = cc * c;
                    vnext = v;
                } else {
                    const double cc = dn ? 0.0 : p.gamma;
                    a = fma(cc, a, (double)rf);
                    c = cc * c;
                }
                A[u] = a;
                C[u] = c;
            }
            sh->aggA[par][warp][lane] = a;
            sh->aggC[par][warp][lane] = c;
            if (MODE == 0 && warp == 0) sh->halo[par][lane] = vreg[0];
        } else {
            // forward: x_t = r_t + g*~done[t-1]*x_{t-1}; S.d holds done rows t0-1 .. t0+TC-2
            if (k == 0) carry = col_ok ? (double)p.prev_ret[col] : 0.0;
            double a = 0.0, c = 1.0;
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int i = warp * R + u;
                const float rf = S.r[i][lane];
                bool dn = S.d[i][lane] != 0;
                if (k == 0 && i == 0) dn = col_ok ? (p.prev_done[col] != 0) : false;
                const double cc = dn ? 0.0 : p.gamma;
                a = fma(cc, a, (double)rf);
                c = cc * c;
                A[u] = a;
                C[u] = c;
            }
            sh->aggA[par][warp][lane] = a;
            sh->aggC[par][warp][lane] = c;
        }

        __syncthreads();  // all smem reads of stage s done; aggregates visible

        if (threadIdx.x == 0 && q + p.n_stages < Q) issue(q + p.n_stages);

        // compose the carry entering this warp's rows, and the strip carry leaving the tile
        double cin = carry;
        if (REVERSE) {
            double x = carry;
#pragma unroll
            for (int w = NWARPS - 1; w >= 0; --w) {
                if (w == warp) cin = x;
                x = fma(sh->aggC[par][w][lane], x, sh->aggA[par][w][lane]);
            }
            carry = x;
        } else {
            double x = carry;
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) {
                if (w == warp) cin = x;
                x = fma(sh->aggC[par][w][lane], x, sh->aggA[par][w][lane]);
            }
            carry = x;
        }

        // pass 2: finalise from registers and store
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int i = warp * R + u;
            const int64_t t = t0 + i;
            const bool ok = col_ok && t >= 0 && t < p.T;
            const float x = __double2float_rn(fma(C[u], cin, A[u]));
            if (!ok) continue;
            float o0;
            if (MODE == 0) {
                o0 = x;
                p.out0[t * p.ld0 + col] = x;
                p.out1[t * p.ld1 + col] = __fadd_rn(x, vreg[u]);
            } else if (MODE == 1) {
                o0 = __fsub_rn(x, vreg[u]);
                p.out1[t * p.ld1 + col] = x;
                p.out0[t * p.ld0 + col] = o0;
            } else {
                o0 = x;
                p.out0[t * p.ld0 + col] = x;
            }
            if (p.tail.moments_out != nullptr) {
                if (p.valid == nullptr || p.valid[t * p.ld_valid + col] != 0) acc.add(o0);
            }
        }
    }

    if (p.tail.moments_out != nullptr) moments_tail(acc.finish(), p.tail, red_scratch);
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

template <int MODE>
static int launch_chunked(const float *reward, int64_t ld_r, const float *value, int64_t ld_v,
                          const uint8_t *done, int64_t ld_d, chunked::Params p, cudaStream_t stream)
{
    using namespace chunked;
    CUtensorMap tm_r, tm_v, tm_d;
    int rc = make_tensor_map_2d(&tm_r, reward, 4, p.T, p.B, ld_r, TC, W);
    if (rc) return rc;
    if (MODE != 2) {
        rc = make_tensor_map_2d(&tm_v, value, 4, p.T, p.B, ld_v, TC, W);
        if (rc) return rc;
    } else {
        tm_v = tm_r;
    }
    rc = make_tensor_map_2d(&tm_d, done, 1, p.T, p.B, ld_d, TC, W);
    if (rc) return rc;

    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    p.n_strips = (int)ceil_div(p.B, W);
    p.n_chunks = (int)ceil_div(p.T, TC);
    int stages, ctas_per_sm;
    if (p.n_strips <= sms) { stages = 8; ctas_per_sm = 1; }
    else if (p.n_strips <= 2 * sms) { stages = 5; ctas_per_sm = 2; }
    else { stages = 3; ctas_per_sm = 3; }
    if (stages > p.n_chunks) stages = p.n_chunks < 1 ? 1 : p.n_chunks;
    p.n_stages = stages;
    const size_t smem = sizeof(Stage) * stages + sizeof(Shared);
    int grid = p.n_strips < sms * ctas_per_sm ? p.n_strips : sms * ctas_per_sm;
    auto kern = scan_chunked_kernel<MODE>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        PLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Stage) * MAX_STAGES + sizeof(Shared))));
        configured = sizeof(Stage) * MAX_STAGES + sizeof(Shared);
    }
    kern<<<grid, THREADS, smem, stream>>>(tm_r, tm_v, tm_d, p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
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
    const bool can_chunk = scan_chunked_supported(reward, ld_r, value, ld_v, done, ld_d, T, B, inner);
    if (algo == PLB_ALGO_AUTO) {
        // the strict serial walk is memory-bound (and bit-exact) once there are enough columns
        algo = (can_chunk && B * inner < 32768) ? PLB_ALGO_CHUNKED : PLB_ALGO_SERIAL;
        if (fusion != nullptr && can_chunk) algo = PLB_ALGO_CHUNKED;
    }
    if (algo == PLB_ALGO_CHUNKED) {
        PLB_REQUIRE(can_chunk, PLB_ERR_UNSUPPORTED,
                    "chunked scan needs inner == 1, 16-byte aligned bases, ld %% 4 == 0 (float) and ld %% 16 == 0 (done)");
        chunked::Params p = {};
        p.boot = boot;
        p.out0 = adv; p.ld0 = ld_a;
        p.out1 = ret; p.ld1 = ld_t;
        p.T = T; p.B = B;
        p.gamma = gamma; p.lambda = lambda;
        p.clip_lo = -INFINITY; p.clip_hi = INFINITY;
        if (fusion != nullptr) {
            p.has_pre = fusion->has_pre;
            p.reward_var = fusion->reward_var;
            p.reward_eps = fusion->reward_eps;
            if (!(fusion->clip_lo != fusion->clip_lo)) p.clip_lo = (float)fusion->clip_lo;
            if (!(fusion->clip_hi != fusion->clip_hi)) p.clip_hi = (float)fusion->clip_hi;
            p.reward_out = fusion->reward_out; p.ld_reward_out = fusion->ld_reward_out;
            p.valid = fusion->valid; p.ld_valid = fusion->ld_valid;
            p.tail = fusion->tail;
        }
        return mode == 0 ? launch_chunked<0>(reward, ld_r, value, ld_v, done, ld_d, p, stream)
                         : launch_chunked<1>(reward, ld_r, value, ld_v, done, ld_d, p, stream);
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

int scan_forward(const float *reward, int64_t ld_r, const uint8_t *done, int64_t ld_d,
                 const float *prev_ret, const uint8_t *prev_done, float *out, int64_t ld_o,
                 int64_t T, int64_t B, double gamma, int algo, const ScanFusion *fusion, cudaStream_t stream)
{
    PLB_REQUIRE(T >= 0 && B >= 0, PLB_ERR_INVALID_ARGUMENT, "negative extent");
    if (T == 0 || B == 0) return PLB_OK;
    PLB_REQUIRE(reward && done && prev_ret && prev_done && out, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    PLB_REQUIRE(ld_r >= B && ld_d >= B && ld_o >= B, PLB_ERR_INVALID_ARGUMENT, "row stride smaller than row length");
    const bool can_chunk = scan_chunked_supported(reward, ld_r, nullptr, 0, done, ld_d, T, B, 1);
    if (algo == PLB_ALGO_AUTO) {
        algo = (can_chunk && B < 32768) ? PLB_ALGO_CHUNKED : PLB_ALGO_SERIAL;
        if (fusion != nullptr && can_chunk) algo = PLB_ALGO_CHUNKED;
    }
    if (algo == PLB_ALGO_CHUNKED) {
        PLB_REQUIRE(can_chunk, PLB_ERR_UNSUPPORTED,
                    "chunked scan needs 16-byte aligned bases, ld %% 4 == 0 (float) and ld %% 16 == 0 (done)");
        chunked::Params p = {};
        p.prev_ret = prev_ret; p.prev_done = prev_done;
        p.out0 = out; p.ld0 = ld_o;
        p.T = T; p.B = B;
        p.gamma = gamma; p.lambda = 1.0;
        p.clip_lo = -INFINITY; p.clip_hi = INFINITY;
        if (fusion != nullptr) {
            p.valid = fusion->valid; p.ld_valid = fusion->ld_valid;
            p.tail = fusion->tail;
        }
        return launch_chunked<2>(reward, ld_r, nullptr, 0, done, ld_d, p, stream);
    }
    PLB_REQUIRE(algo == PLB_ALGO_SERIAL, PLB_ERR_INVALID_ARGUMENT, "unknown algo %d", algo);
    PLB_REQUIRE(fusion == nullptr, PLB_ERR_UNSUPPORTED, "fused statistics need the chunked scan");
    const int threads = 128;
    scan_forward_serial_kernel<<<(unsigned)ceil_div(B, threads), threads, 0, stream>>>(
        reward, ld_r, done, ld_d, prev_ret, prev_done, out, ld_o, T, B, gamma);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}

}  // namespace plb

extern "C" {

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
