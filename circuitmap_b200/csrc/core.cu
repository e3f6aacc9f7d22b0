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
// Timing of the dominant kernel is a diagnostic: it must never fail the compute call.  One event pair per device
// (events belong to the device that was current when they were created), created lazily, every CUDA return checked
// and swallowed; a failure only makes cm_last_main_kernel_ms() report "no timed kernel".
constexpr int MAX_DEV = 64;
static thread_local cudaEvent_t g_ev0[MAX_DEV] = {}, g_ev1[MAX_DEV] = {};
static thread_local int g_timed_dev = -1;          // device whose event pair holds the last timed kernel, -1 = none
static thread_local int g_begin_dev = -1;
void main_kernel_begin(cudaStream_t st) {
    g_begin_dev = -1;
    g_timed_dev = -1;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) { (void)cudaGetLastError(); return; }
    if (!g_ev0[dev]) {
        if (cudaEventCreate(&g_ev0[dev]) != cudaSuccess || cudaEventCreate(&g_ev1[dev]) != cudaSuccess) {
            (void)cudaGetLastError();
            g_ev0[dev] = g_ev1[dev] = nullptr;
            return;
        }
    }
    if (cudaEventRecord(g_ev0[dev], st) != cudaSuccess) { (void)cudaGetLastError(); return; }
    g_begin_dev = dev;
}
void main_kernel_end(cudaStream_t st) {
    const int dev = g_begin_dev;
    g_begin_dev = -1;
    if (dev < 0) return;
    if (cudaEventRecord(g_ev1[dev], st) != cudaSuccess) { (void)cudaGetLastError(); return; }
    g_timed_dev = dev;
}
void reset_launch_count() { g_launches = 0; }
}  // namespace cm

extern "C" int cm_version(void) { return CM_VERSION; }
extern "C" const char* cm_last_error(void) { return cm::g_err; }
extern "C" int cm_last_launch_count(void) { return cm::g_launches; }
extern "C" float cm_last_main_kernel_ms(void) {
    const int dev = cm::g_timed_dev;
    if (dev < 0) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(cm::g_ev1[dev]) != cudaSuccess) { (void)cudaGetLastError(); return -1.f; }
    if (cudaEventElapsedTime(&ms, cm::g_ev0[dev], cm::g_ev1[dev]) != cudaSuccess) { (void)cudaGetLastError(); return -1.f; }
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
