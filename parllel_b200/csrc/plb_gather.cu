// plb_gather.cu -- ReplayBuffer.sample_batch() gather (parllel/replays/replay.py:72):
//   batch[leaf][i, :] = ring[leaf][T_idxs[i], B_idxs[i], :]   for every leaf, in ONE launch.
// Pure byte movement (bit-exact by construction).  Rows are copied in the widest power-of-two
// chunks (<= 16 bytes) that the leaf's row size, strides and base alignment allow; consecutive
// threads take consecutive chunks of consecutive samples, so both the row reads and the dense
// output writes are coalesced.
//
// Two roles in the one launch:
//   * rows of >= 128 bytes (16-byte aligned): TMA-STAGED.  A warp takes 32 samples of one leaf at a time: every
//     lane reads its sample's (t, b) and issues ONE bulk copy (cp.async.bulk, the 1-D TMA) of its row into the
//     warp's shared-memory tile, all 32 completing on the warp's mbarrier; the tile is then exactly the dense
//     [32, row] block of the output and leaves as ONE bulk store (8 KB for 256-byte rows).  No data passes
//     through registers; ~28 warps per SM keep > 200 KB of row reads in flight;
//   * smaller / unaligned rows (reward, terminated, 32-byte actions): the load/store-unit path below, 16-byte
//     chunks, consecutive threads on consecutive chunks.
#include "plb_common.cuh"
#include "plb_scan_shared.cuh"

namespace plb {

struct GatherLeafDev {
    const unsigned char *src;
    unsigned char *dst;
    int64_t stride_t, stride_b;  // bytes
    int32_t row_bytes;
    int32_t chunk_log2;          // log2(chunk bytes), 0..4
    int64_t first_block;         // first blockIdx.x that works on this leaf
    int64_t n_chunks;            // n * row_bytes >> chunk_log2
};

struct GatherParams {
    GatherLeafDev leaf[PLB_GATHER_MAX_LEAVES];
    int32_t n_leaves;
    const int64_t *t_idx, *b_idx;
    int64_t n;
    // TMA-staged role: leaves [0, n_tma_leaves) of `tma_leaf` (indices into leaf[]), work items = (leaf, group of
    // `group` samples), dealt round-robin to the warps of CTAs [0, tma_blocks)
    int32_t n_tma_leaves;
    int32_t tma_leaf[PLB_GATHER_MAX_LEAVES];
    int32_t tma_group[PLB_GATHER_MAX_LEAVES];       // samples per work item (<= 32, tile <= GATHER_TILE_BYTES)
    int64_t tma_first_item[PLB_GATHER_MAX_LEAVES + 1];
    int64_t tma_blocks;
};

constexpr int GATHER_TILE_BYTES = 8192;   // per warp
constexpr int GATHER_TMA_WARPS = 8;       // per CTA: every warp of the 256-thread CTA
constexpr int GATHER_TMA_MIN_ROW = 128;

constexpr int GATHER_THREADS = 256;
constexpr int GATHER_U = 4;  // chunks per thread, all loads issued before any store

template <typename V>
__device__ __forceinline__ void gather_chunks(const GatherLeafDev &L, const int64_t *__restrict__ t_idx,
                                              const int64_t *__restrict__ b_idx, int64_t block_in_leaf)
{
    const int32_t cpr = L.row_bytes / (int32_t)sizeof(V);  // chunks per row
    const int64_t base = block_in_leaf * (GATHER_THREADS * GATHER_U) + threadIdx.x;
    V v[GATHER_U];
    int64_t out_off[GATHER_U];
#pragma unroll
    for (int u = 0; u < GATHER_U; ++u) {
        const int64_t i = base + (int64_t)u * GATHER_THREADS;
        if (i < L.n_chunks) {
            const int64_t s = i / cpr;
            const int64_t c = i - s * cpr;
            const int64_t t = t_idx[s], b = b_idx[s];
            v[u] = *reinterpret_cast<const V *>(L.src + t * L.stride_t + b * L.stride_b + c * (int64_t)sizeof(V));
            out_off[u] = i * (int64_t)sizeof(V);
        }
    }
#pragma unroll
    for (int u = 0; u < GATHER_U; ++u) {
        const int64_t i = base + (int64_t)u * GATHER_THREADS;
        if (i < L.n_chunks) *reinterpret_cast<V *>(L.dst + out_off[u]) = v[u];
    }
}

// TMA-staged role (see the file header); called by CTAs [0, tma_blocks) with GATHER_TMA_WARPS warps each
__device__ __forceinline__ void gather_rows_tma(const GatherParams &p)
{
    extern __shared__ __align__(128) unsigned char tiles[];  // [GATHER_TMA_WARPS][GATHER_TILE_BYTES], dynamic
    __shared__ unsigned long long bars[GATHER_TMA_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long *bar = &bars[warp];
    unsigned char *tile = tiles + (size_t)warp * GATHER_TILE_BYTES;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_proxy_async_smem();
    }
    __syncwarp();
    const int64_t n_items = p.tma_first_item[p.n_tma_leaves];
    const int64_t n_warps = p.tma_blocks * GATHER_TMA_WARPS;
    uint32_t phase = 0;
    for (int64_t item = (int64_t)blockIdx.x * GATHER_TMA_WARPS + warp; item < n_items; item += n_warps) {
        int k = 0;
#pragma unroll 1
        for (int q = 1; q < p.n_tma_leaves; ++q)
            if (item >= p.tma_first_item[q]) k = q;
        const GatherLeafDev &L = p.leaf[p.tma_leaf[k]];
        const int group = p.tma_group[k];
        const int64_t s0 = (item - p.tma_first_item[k]) * group;
        const int cnt = (int)((p.n - s0 < group) ? (p.n - s0) : group);
        const uint32_t row = (uint32_t)L.row_bytes;
        const unsigned char *src = nullptr;
        if (lane < cnt) src = L.src + p.t_idx[s0 + lane] * L.stride_t + p.b_idx[s0 + lane] * L.stride_b;
        if (lane == 0) mbar_arrive_expect_tx(bar, row * (uint32_t)cnt);
        __syncwarp();
        if (lane < cnt) bulk_load_1d(tile + (size_t)lane * row, src, row, bar);
        mbar_wait(bar, phase);
        phase ^= 1u;
        if (lane == 0) {
            bulk_store_1d(L.dst + s0 * (int64_t)row, tile, row * (uint32_t)cnt);
            tma_store_commit();
            tma_store_wait_read<0>();  // the tile is free again once the store has read it
        }
        __syncwarp();
    }
    if (lane == 0) tma_store_wait<0>();
}

__global__ void __launch_bounds__(GATHER_THREADS)
gather_tree_kernel(const __grid_constant__ GatherParams p)
{
    if ((int64_t)blockIdx.x < p.tma_blocks) {
        gather_rows_tma(p);
        return;
    }
    // which leaf does this CTA serve?  (n_leaves <= 16: linear search is free)
    int l = 0;
#pragma unroll 1
    for (int k = 1; k < p.n_leaves; ++k)
        if ((int64_t)blockIdx.x >= p.leaf[k].first_block) l = k;
    const GatherLeafDev &L = p.leaf[l];
    const int64_t blk = (int64_t)blockIdx.x - L.first_block;
    switch (L.chunk_log2) {
        case 4: gather_chunks<uint4>(L, p.t_idx, p.b_idx, blk); break;
        case 3: gather_chunks<uint2>(L, p.t_idx, p.b_idx, blk); break;
        case 2: gather_chunks<uint32_t>(L, p.t_idx, p.b_idx, blk); break;
        case 1: gather_chunks<uint16_t>(L, p.t_idx, p.b_idx, blk); break;
        default: gather_chunks<uint8_t>(L, p.t_idx, p.b_idx, blk); break;
    }
}

static int g_gather_mode = 0;  // 0: TMA-staged rows where eligible (default); 1: load/store-unit path only (key 16)
void set_gather_mode(int v) { g_gather_mode = v; }

}  // namespace plb

extern "C" int plb_gather_tree(const plb_gather_leaf *leaves, int32_t n_leaves, const int64_t *t_idx,
                               const int64_t *b_idx, int64_t n, plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(n_leaves >= 0 && n_leaves <= PLB_GATHER_MAX_LEAVES, PLB_ERR_INVALID_ARGUMENT,
                "n_leaves must be in [0, %d]", PLB_GATHER_MAX_LEAVES);
    PLB_REQUIRE(n >= 0, PLB_ERR_INVALID_ARGUMENT, "negative sample count");
    if (n == 0 || n_leaves == 0) return PLB_OK;
    PLB_REQUIRE(leaves && t_idx && b_idx, PLB_ERR_INVALID_ARGUMENT, "null pointer");
    GatherParams p = {};
    p.n_leaves = n_leaves;
    p.t_idx = t_idx;
    p.b_idx = b_idx;
    p.n = n;
    // pass 1: which leaves take the TMA-staged role
    int64_t tma_items = 0;
    bool is_tma[PLB_GATHER_MAX_LEAVES] = {};
    for (int l = 0; l < n_leaves; ++l) {
        const plb_gather_leaf &h = leaves[l];
        PLB_REQUIRE(h.src && h.dst && h.row_bytes > 0 && h.row_bytes <= (int64_t)INT32_MAX, PLB_ERR_INVALID_ARGUMENT,
                    "leaf %d: bad descriptor", l);
        is_tma[l] = g_gather_mode != 1 && h.row_bytes >= GATHER_TMA_MIN_ROW && h.row_bytes <= GATHER_TILE_BYTES &&
                    h.row_bytes % 16 == 0 && h.stride_t_bytes % 16 == 0 && h.stride_b_bytes % 16 == 0 &&
                    aligned(h.src, 16) && aligned(h.dst, 16);
        if (is_tma[l]) {
            const int k = p.n_tma_leaves++;
            int group = (int)(GATHER_TILE_BYTES / h.row_bytes);
            if (group > 32) group = 32;
            p.tma_leaf[k] = l;
            p.tma_group[k] = group;
            p.tma_first_item[k] = tma_items;
            tma_items += ceil_div(n, group);
        }
    }
    p.tma_first_item[p.n_tma_leaves] = tma_items;
    const int sms = sm_count();
    PLB_REQUIRE(sms > 0, PLB_ERR_UNSUPPORTED, "no CUDA device");
    p.tma_blocks = ceil_div(tma_items, GATHER_TMA_WARPS);
    const int64_t tma_cap = (int64_t)sms * 3;  // persistent beyond 3 CTAs (24 warps, 192 KB of tiles) per SM
    if (p.tma_blocks > tma_cap) p.tma_blocks = tma_cap;
    int64_t blocks = p.tma_blocks;
    for (int l = 0; l < n_leaves; ++l) {
        const plb_gather_leaf &h = leaves[l];
        int lg = 4;
        while (lg > 0) {
            const int64_t a = (int64_t)1 << lg;
            if (h.row_bytes % a == 0 && h.stride_t_bytes % a == 0 && h.stride_b_bytes % a == 0 &&
                aligned(h.src, (size_t)a) && aligned(h.dst, (size_t)a))
                break;
            --lg;
        }
        GatherLeafDev &d = p.leaf[l];
        d.src = reinterpret_cast<const unsigned char *>(h.src);
        d.dst = reinterpret_cast<unsigned char *>(h.dst);
        d.stride_t = h.stride_t_bytes;
        d.stride_b = h.stride_b_bytes;
        d.row_bytes = (int32_t)h.row_bytes;
        d.chunk_log2 = lg;
        d.first_block = blocks;
        d.n_chunks = is_tma[l] ? 0 : n * (h.row_bytes >> lg);  // TMA-staged leaves need no load/store-unit CTAs
        blocks += ceil_div(d.n_chunks, GATHER_THREADS * GATHER_U);
    }
    PLB_REQUIRE(blocks <= (int64_t)INT32_MAX, PLB_ERR_UNSUPPORTED, "gather too large for one launch");
    const size_t smem = p.tma_blocks > 0 ? (size_t)GATHER_TMA_WARPS * GATHER_TILE_BYTES : 0;
    static thread_local bool configured = false;
    if (!configured) {
        PLB_CUDA(cudaFuncSetAttribute(gather_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GATHER_TMA_WARPS * GATHER_TILE_BYTES));
        configured = true;
    }
    // (programmatic dependent launch measured here: back-to-back gathers of n = 16 384 went from 5.1 to 6.4 us -- the
    //  early-scheduled CTAs of the next launch hold shared memory the TMA role wants -- so this stays a plain launch)
    gather_tree_kernel<<<(unsigned)blocks, GATHER_THREADS, smem, (cudaStream_t)stream>>>(p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}
