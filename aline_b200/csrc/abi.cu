// Library-wide state: error string, launch counter, device info.
#include "common.cuh"
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace aline {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return 1;
}

thread_local bool g_pdl_chain = false;

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("ALINE_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

const DeviceInfo& device_info() {
    static thread_local DeviceInfo info;
    static thread_local int cached_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&info.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cached_dev = dev;
    }
    return info;
}

int ensure_dyn_smem(const void* kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return 0;
    const DeviceInfo& di = device_info();
    ALINE_REQUIRE(bytes <= (size_t)di.max_smem_optin, "kernel needs %zu bytes of shared memory (max %d)", bytes,
                  di.max_smem_optin);
    int dev = 0;
    cudaGetDevice(&dev);
    static std::mutex mu;
    static std::unordered_map<uint64_t, int> done;
    const uint64_t key = (uint64_t)(uintptr_t)kernel * 64u + (uint64_t)(dev & 63);
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(key);
    if (it != done.end()) return 0;
    // the opt-in limit covers static + dynamic shared memory of the kernel
    cudaFuncAttributes fa;
    ALINE_CHECK_CUDA(cudaFuncGetAttributes(&fa, kernel));
    const int most = di.max_smem_optin - (int)fa.sharedSizeBytes;
    ALINE_REQUIRE((size_t)most >= bytes, "kernel needs %zu dynamic + %zu static bytes of shared memory (max %d)", bytes,
                  fa.sharedSizeBytes, di.max_smem_optin);
    ALINE_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
    done[key] = most;
    return 0;
}

}  // namespace aline

extern "C" {
int aline_abi_version(void) { return ALINE_ABI_VERSION; }
const char* aline_last_error(void) { return aline::g_last_error.c_str(); }
uint64_t aline_kernel_launches(void) { return aline::g_launches.load(); }
}
