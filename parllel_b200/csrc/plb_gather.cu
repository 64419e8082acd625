// plb_gather.cu -- ReplayBuffer.sample_batch() gather (parllel/replays/replay.py:72):
//   batch[leaf][i, :] = ring[leaf][T_idxs[i], B_idxs[i], :]   for every leaf, in ONE launch.
// Pure byte movement (bit-exact by construction).  Rows are copied in the widest power-of-two
// chunks (<= 16 bytes) that the leaf's row size, strides and base alignment allow; consecutive
// threads take consecutive chunks of consecutive samples, so both the row reads and the dense
// output writes are coalesced.
#include "plb_common.cuh"

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
};

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

__global__ void __launch_bounds__(GATHER_THREADS)
gather_tree_kernel(const __grid_constant__ GatherParams p)
{
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
    int64_t blocks = 0;
    for (int l = 0; l < n_leaves; ++l) {
        const plb_gather_leaf &h = leaves[l];
        PLB_REQUIRE(h.src && h.dst && h.row_bytes > 0 && h.row_bytes <= (int64_t)INT32_MAX, PLB_ERR_INVALID_ARGUMENT,
                    "leaf %d: bad descriptor", l);
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
        d.n_chunks = n * (h.row_bytes >> lg);
        blocks += ceil_div(d.n_chunks, GATHER_THREADS * GATHER_U);
    }
    PLB_REQUIRE(blocks <= (int64_t)INT32_MAX, PLB_ERR_UNSUPPORTED, "gather too large for one launch");
    gather_tree_kernel<<<(unsigned)blocks, GATHER_THREADS, 0, (cudaStream_t)stream>>>(p);
    PLB_LAUNCH_CHECK();
    return PLB_OK;
}
