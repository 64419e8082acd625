// Microbenchmark: TMA 2-D tile load throughput vs inner box width (bytes per row request).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_width tma_width.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(unsigned long long *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load(void *dst, const CUtensorMap *tm, int c0, int c1, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store(const CUtensorMap *tm, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)tm), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

// each CTA walks its own column strip of `width` bytes top to bottom, `stages` boxes in flight
__global__ void k_load(const __grid_constant__ CUtensorMap tm, int width, int box_rows, int n_boxes, int stages, int do_store)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar[8];
    const uint32_t box_bytes = (uint32_t)width * box_rows;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int c0 = blockIdx.x * width;
        for (int s = 0; s < stages && s < n_boxes; ++s) { mbar_expect(&bar[s], box_bytes); tma_load(smem + s * box_bytes, &tm, c0, s * box_rows, &bar[s]); }
        for (int i = 0; i < n_boxes; ++i) {
            const int s = i % stages;
            mbar_wait(&bar[s], (i / stages) & 1);
            if (do_store) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tma_store(&tm, smem + s * box_bytes, c0, i * box_rows);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (i + stages < n_boxes) { mbar_expect(&bar[s], box_bytes); tma_load(smem + s * box_bytes, &tm, c0, (i + stages) * box_rows, &bar[s]); }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main()
{
    typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                            const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *sym; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    Enc enc = (Enc)sym;
    const int sms = 148;
    const size_t rows = 16384, row_bytes = 65536;  // 1 GiB tensor: far beyond L2
    unsigned char *buf; CK(cudaMalloc(&buf, rows * row_bytes)); CK(cudaMemset(buf, 1, rows * row_bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int do_store = 0; do_store < 2; ++do_store)
    for (int width : {32, 64, 128, 256, 512}) {
        for (int box_bytes : {16384, 32768}) {
            int box_rows = box_bytes / width; if (box_rows > 256) box_rows = 256;
            const int real_box = box_rows * width;
            const int stages = 4;
            const int n_boxes = (int)(rows / box_rows);
            CUtensorMap tm;
            cuuint64_t gdim[2] = {row_bytes, rows}; cuuint64_t gstr[1] = {row_bytes};
            cuuint32_t box[2] = {(cuuint32_t)width, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            const int grid = sms;
            k_load<<<grid, 32, stages * real_box>>>(tm, width, box_rows, n_boxes, stages, do_store);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            k_load<<<grid, 32, stages * real_box>>>(tm, width, box_rows, n_boxes, stages, do_store);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)grid * n_boxes * real_box * (do_store ? 2 : 1);
            const double reqs = (double)n_boxes * box_rows * (do_store ? 2 : 1);
            printf("%s width %4d B  box %3d rows (%5d B) x%d stages: %8.1f GB/s   %.1f cycles/row-request/SM (at 1.9 GHz)\n",
                   do_store ? "load+store" : "load      ", width, box_rows, real_box, stages, bytes / ms / 1e6, ms * 1e-3 * 1.9e9 / reqs);
        }
    }
    return 0;
}
