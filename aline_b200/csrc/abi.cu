// Library-wide state: error string, launch counter, device info.
#include "common.cuh"
#include <cstdarg>
#include <cstdlib>

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

}  // namespace aline

extern "C" {
int aline_abi_version(void) { return ALINE_ABI_VERSION; }
const char* aline_last_error(void) { return aline::g_last_error.c_str(); }
uint64_t aline_kernel_launches(void) { return aline::g_launches.load(); }
}
