// Shared helpers for the aline_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cfloat>
#include <string>
#include <atomic>
#include <type_traits>

#include "../../include/aline_b200.h"

namespace aline {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

int set_error(const char* fmt, ...);

#define ALINE_CHECK_CUDA(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::aline::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                                      cudaGetErrorString(_e));                              \
    } while (0)

#define ALINE_REQUIRE(cond, ...)                                                            \
    do {                                                                                    \
        if (!(cond)) return ::aline::set_error(__VA_ARGS__);                                \
    } while (0)

// Check the launch that was just enqueued and count it.
#define ALINE_LAUNCH_OK()                                                                   \
    do {                                                                                    \
        ::aline::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
        ALINE_CHECK_CUDA(cudaGetLastError());                                               \
    } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// A kernel launched with `pdl = true` may start (prologue: barrier init, TMEM allocation, TMA staging of the weights --
// nothing the preceding kernel of the stream writes) while the preceding kernel drains; pdl_wait() then blocks until
// that kernel has completed and its writes are visible.  pdl_trigger() at the top of a kernel lets ITS successor do
// the same.  Both are no-ops for ordinary launches.  Only the back-to-back launches of aline_rollout use pdl = true:
// there every kernel of the chain executes pdl_wait(), so completion is transitive (ALINE_PDL=0 disables it).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
extern thread_local bool g_pdl_chain;   // set by aline_rollout while the preceding launch of the stream is its own

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                            Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Opt a kernel in to more than 48 KB of dynamic shared memory ONCE per (kernel, device): the attribute is raised to the
// device's opt-in maximum the first time, later launches only pay a hash lookup (a cudaFuncSetAttribute per launch costs
// microseconds on the host-bound rollout chain).  Returns 0 / 1 like the launchers.
int ensure_dyn_smem(const void* kernel, size_t bytes);

struct DeviceInfo {
    int sm_count = 0;
    int max_smem_optin = 0;
};
const DeviceInfo& device_info();

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- streaming loads / stores ------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// coherent streaming load for buffers the same kernel later overwrites (seq accumulator)
__device__ __forceinline__ float ld_stream1(const float* p) {
    float r;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// exp(x) = MUFU.EX2(x * log2 e), flushing to zero below 2^-126 (x <= 0 on every call site)
__device__ __forceinline__ float exp_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.44269504088896340736f));
    return r;
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 -- two fp32 lanes per issue slot) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// ---- online log-sum-exp pair (m, s): value = m + log(s) ---------------------------
struct Lse {
    float m, s;
    __device__ __forceinline__ void init() { m = -FLT_MAX; s = 0.f; }
    __device__ __forceinline__ void push(float v) {
        if (v > m) { s *= exp_fast(m - v); m = v; }   // rescale only when the max moves (rare)
        s += exp_fast(v - m);                         // MUFU.EX2; terms that matter have |v - m| small
    }
    __device__ __forceinline__ void merge(float m2, float s2) {
        float M = fmaxf(m, m2);
        s = s * expf(m - M) + s2 * expf(m2 - M);
        m = M;
    }
};

}  // namespace aline
