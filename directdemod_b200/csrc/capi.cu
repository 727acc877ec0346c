// Library-level entry points of libddemod.so: version, error string, launch counter.
#include <atomic>
#include <mutex>

#include "ddm_common.cuh"

namespace ddm {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count(int device) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0)
        sms = 148;
    return sms;
}

}  // namespace ddm

extern "C" {

int ddm_version(void) { return 100; }   // 0.1.0

const char *ddm_last_error(void) { return ddm::g_last_error.c_str(); }

int64_t ddm_launch_count(void) { return ddm::g_launches.load(std::memory_order_relaxed); }

int ddm_device_count(int *count) {
    DDM_REQUIRE(count != nullptr, "ddm_device_count: NULL argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        // no driver / no device is a valid answer on a CPU-only box
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return DDM_OK;
}

}  // extern "C"
