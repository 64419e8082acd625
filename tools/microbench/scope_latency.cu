// scope_latency.cu -- how long does one flag hop between two CTAs take, by memory-ordering scope?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scope_latency scope_latency.cu && ./scope_latency
// CTA 0 and CTA 1 (different SMs) bounce a counter ROUNDS times through two words in device memory:
//   CTA 0: store flagA = i          ; spin until flagB == i
//   CTA 1: spin until flagA == i    ; store flagB = i
// One round = two store->visible->load hops.  Variants: relaxed.sys / relaxed.gpu / volatile / .cg loads with
// plain or .sys stores.  Also a dependent-load chain (pointer chase in an L2-resident buffer) per load flavour.
#include <cstdio>
#include <cuda_runtime.h>

template <int LD>
__device__ __forceinline__ unsigned load_flag(const unsigned *p)
{
    unsigned v;
    if (LD == 0) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else if (LD == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else if (LD == 2) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else if (LD == 3) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int ST>
__device__ __forceinline__ void store_flag(unsigned *p, unsigned v)
{
    if (ST == 0) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else if (ST == 1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else if (ST == 2) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

template <int LD, int ST>
__global__ void pingpong(unsigned *flags, int rounds, long long *ns_out)
{
    if (threadIdx.x != 0) return;
    unsigned *a = flags, *b = flags + 64;  // separate 128-byte lines
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 1; i <= rounds; ++i) {
        if (blockIdx.x == 0) {
            store_flag<ST>(a, (unsigned)i);
            while (load_flag<LD>(b) != (unsigned)i) {}
        } else {
            while (load_flag<LD>(a) != (unsigned)i) {}
            store_flag<ST>(b, (unsigned)i);
        }
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (blockIdx.x == 0) *ns_out = (long long)(t1 - t0);
}

template <int LD>
__global__ void chase(const unsigned *buf, int hops, long long *ns_out, unsigned *sink)
{
    unsigned long long t0, t1;
    unsigned idx = 0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int i = 0; i < hops; ++i) idx = load_flag<LD>(buf + idx);
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    *ns_out = (long long)(t1 - t0);
    *sink = idx;
}

int main()
{
    unsigned *flags, *buf, *sink;
    long long *ns;
    cudaMalloc(&flags, 1024);
    cudaMalloc(&sink, 4);
    cudaMallocManaged(&ns, 8);
    const int N = 1 << 16, rounds = 2000, hops = 2000;
    cudaMalloc(&buf, N * 4);
    unsigned *h = new unsigned[N];
    for (int i = 0; i < N; ++i) h[i] = (unsigned)((i * 1237u + 4099u) % N);
    cudaMemcpy(buf, h, N * 4, cudaMemcpyHostToDevice);
    const char *ld_names[] = {"ld.relaxed.sys", "ld.relaxed.gpu", "ld.volatile", "ld.global.cg", "ld.acquire.sys"};
    const char *st_names[] = {"st.relaxed.sys", "st.relaxed.gpu", "st.volatile", "red.relaxed.sys.add(+1 per round: flag == i)"};
#define PP(LD, ST)                                                                                        \
    cudaMemset(flags, 0, 1024);                                                                           \
    pingpong<LD, ST><<<2, 32>>>(flags, rounds, ns);                                                       \
    cudaDeviceSynchronize();                                                                              \
    printf("ping-pong %-16s / %-16s : %7.1f ns per one-way hop\n", ld_names[LD], st_names[ST], (double)*ns / rounds / 2);
    PP(0, 0) PP(1, 1) PP(2, 2) PP(3, 1) PP(3, 0) PP(1, 0) PP(0, 1) PP(4, 0) PP(0, 3) PP(1, 3)
#define CH(LD)                                                                                            \
    chase<LD><<<1, 1>>>(buf, hops, ns, sink);                                                             \
    cudaDeviceSynchronize();                                                                              \
    printf("dependent loads  %-16s : %7.1f ns per load\n", ld_names[LD], (double)*ns / hops);
    CH(0) CH(1) CH(2) CH(3) CH(4)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
