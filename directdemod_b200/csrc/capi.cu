// Library-level entry points of libddemod.so: version, error string, launch counter.
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

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

// Per-device scratch pool for the operators that need temporary device memory (correlation
// prefix sums, selection histograms, compaction offsets).  Slots grow on demand and are kept, so
// steady-state calls do no cudaMalloc/cudaFree.  Calls that use the pool must be serialised per
// device by the caller (the same rule as for handles).
namespace {
constexpr int kScratchSlots = 10;     // 0..6: ncc.cu, 8..9: filtfilt temporaries (filter.cu)
constexpr int kScratchDevices = 64;
struct ScratchSlot {
    void *p = nullptr;
    size_t cap = 0;
};
ScratchSlot g_scratch[kScratchDevices][kScratchSlots];
std::mutex g_scratch_mu;
}  // namespace

void *scratch_get(int device, int slot, size_t bytes) {
    if (device < 0 || device >= kScratchDevices || slot < 0 || slot >= kScratchSlots) return nullptr;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    ScratchSlot &s = g_scratch[device][slot];
    if (bytes <= s.cap && s.p) return s.p;
    if (s.p) cudaFree(s.p);           // implicit device synchronisation
    s.p = nullptr;
    s.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&s.p, want) != cudaSuccess) {
        cudaGetLastError();
        s.p = nullptr;
        set_error("scratch allocation of %zu bytes failed on device %d", want, device);
        return nullptr;
    }
    s.cap = want;
    return s.p;
}

void scratch_release(int device) {
    if (device < 0 || device >= kScratchDevices) return;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    for (ScratchSlot &s : g_scratch[device]) {
        if (s.p) cudaFree(s.p);
        s.p = nullptr;
        s.cap = 0;
    }
}

// Small-block pool: the fixed buffers of a handle (taps, halos, delay lines, tables) are a few KB;
// decoders create and drop handles per call, and cudaMalloc / cudaFree cost milliseconds each on
// a context that holds gigabytes.  Blocks are rounded up to a power of two (>= 4 KB), recycled
// through per-device free lists and only returned to the driver by ddm_release_scratch.
namespace {
struct PoolEntry {
    int device;
    int cls;
};
std::mutex g_pool_mu;
std::vector<void *> g_pool_free[kScratchDevices][24];
std::unordered_map<void *, PoolEntry> g_pool_live;
int pool_class(size_t bytes) {
    int c = 12;                                   // 4 KB
    while ((static_cast<size_t>(1) << c) < bytes) ++c;
    return c;
}
}  // namespace

cudaError_t pool_alloc(void **out, size_t bytes) {
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    const int cls = pool_class(bytes ? bytes : 1);
    if (device < 0 || device >= kScratchDevices || cls >= 36) return cudaMalloc(out, bytes);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (cls - 12 < 24 && !g_pool_free[device][cls - 12].empty()) {
        *out = g_pool_free[device][cls - 12].back();
        g_pool_free[device][cls - 12].pop_back();
    } else {
        e = cudaMalloc(out, static_cast<size_t>(1) << cls);
        if (e != cudaSuccess) return e;
    }
    g_pool_live[*out] = PoolEntry{device, cls};
    return cudaSuccess;
}

void pool_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool_live.find(p);
    if (it == g_pool_live.end()) {
        cudaFree(p);                              // not ours (allocated with cudaMalloc directly)
        return;
    }
    const PoolEntry e = it->second;
    g_pool_live.erase(it);
    if (e.cls - 12 < 24 && g_pool_free[e.device][e.cls - 12].size() < 4096) g_pool_free[e.device][e.cls - 12].push_back(p);
    else cudaFree(p);
}

void pool_release(int device) {
    if (device < 0 || device >= kScratchDevices) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto &lst : g_pool_free[device]) {
        for (void *p : lst) cudaFree(p);
        lst.clear();
    }
}

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

int ddm_release_scratch(int device) {
    int ndev = 0;
    DDM_CUDA(cudaGetDeviceCount(&ndev));
    DDM_REQUIRE(device >= 0 && device < ndev, "ddm_release_scratch: no such device %d", device);
    ddm::DeviceGuard guard(device);
    ddm::scratch_release(device);
    ddm::pool_release(device);
    return DDM_OK;
}

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
