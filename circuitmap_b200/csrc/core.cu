// Library-wide C-ABI plumbing: version, thread-local error text, launch accounting.
#include "common.cuh"

namespace cm {
static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static thread_local bool g_timed = false;
void main_kernel_begin(cudaStream_t st) {
    if (!g_ev0) { cudaEventCreate(&g_ev0); cudaEventCreate(&g_ev1); }
    cudaEventRecord(g_ev0, st);
}
void main_kernel_end(cudaStream_t st) { cudaEventRecord(g_ev1, st); g_timed = true; }
void reset_launch_count() { g_launches = 0; }
}  // namespace cm

extern "C" int cm_version(void) { return CM_VERSION; }
extern "C" const char* cm_last_error(void) { return cm::g_err; }
extern "C" int cm_last_launch_count(void) { return cm::g_launches; }
extern "C" float cm_last_main_kernel_ms(void) {
    if (!cm::g_timed) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(cm::g_ev1) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, cm::g_ev0, cm::g_ev1) != cudaSuccess) return -1.f;
    return ms;
}
extern "C" int cm_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    CM_CUDA_CHECK(cudaGetDevice(&dev));
    if (sm_count) CM_CUDA_CHECK(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) CM_CUDA_CHECK(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) CM_CUDA_CHECK(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return CM_OK;
}
