// plb_api.cu -- library-level entry points: version, error string, device query.
#include <stdarg.h>

#include "plb_common.cuh"

namespace plb {

static thread_local char g_last_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0, g_resident_launches = 0;  // enqueued by this process (any thread)
void count_launch(int resident)
{
    __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    if (resident) __atomic_add_fetch(&g_resident_launches, 1, __ATOMIC_RELAXED);
}

static int g_pdl = 1;
int pdl_enabled() { return g_pdl; }
void set_pdl_enabled(int on) { g_pdl = on ? 1 : 0; }

int sm_count()
{
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

}  // namespace plb

extern "C" {

int plb_version(void) { return PLB_VERSION; }

const char *plb_last_error(void) { return plb::g_last_error; }

void plb_launch_counters(int64_t *out2)
{
    if (out2 == nullptr) return;
    out2[0] = __atomic_load_n(&plb::g_launches, __ATOMIC_RELAXED);
    out2[1] = __atomic_load_n(&plb::g_resident_launches, __ATOMIC_RELAXED);
}

int plb_graph_launch(void *graph_exec, plb_stream_t stream)
{
    using namespace plb;
    PLB_REQUIRE(graph_exec != nullptr, PLB_ERR_INVALID_ARGUMENT, "null graph");
    PLB_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), (cudaStream_t)stream));
    return PLB_OK;
}

int plb_sm_count(void)
{
    const int n = plb::sm_count();
    if (n < 0) plb::set_error("no CUDA device available");
    return n;
}

}  // extern "C"
