// Hardware self-test of the tcgen05 building blocks (csrc/tc.cuh): D[128, N] = A[128, K] * B[N, K]^T with bf16
// operands staged in the core-matrix tiled shared-memory layout, fp32 accumulation in TMEM, read back with
// tcgen05.ld.  tests/test_tc_gpu.py compares it with a bf16-rounded torch matmul.
#include "tc.cuh"

namespace aline {

__global__ void __launch_bounds__(128)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K, float* __restrict__ D,
                   int use_bulk, const unsigned char* __restrict__ Bpacked, int a_in_tmem) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar, bar_tma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* As = smem;                         // 128 rows x K
    unsigned char* Bs = smem + (size_t)128 * K * 2;   // N rows x K
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init(&bar_tma, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    // A: thread = row
    for (int c = 0; c < K / 8; ++c) {
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * K + 8 * c + i];
        tc::store_row_bf16<8>(As, 128, tid, v, c);
    }
    if (use_bulk) {
        // B arrives pre-packed (bf16, tiled layout) through one TMA bulk copy
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&bar_tma, (uint32_t)(N * K * 2));
            tc::bulk_g2s(Bs, Bpacked, (uint32_t)(N * K * 2), &bar_tma);
        }
        tc::mbar_wait(&bar_tma, 0);
    } else {
        for (int r = tid; r < N; r += 128)
            for (int c = 0; c < K / 8; ++c) {
                float v[8];
                for (int i = 0; i < 8; ++i) v[i] = B[(size_t)r * K + 8 * c + i];
                tc::store_row_bf16<8>(Bs, N, r, v, c);
            }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (a_in_tmem) {
        // A operand staged in tensor memory (columns 256 ..): packed bf16 pairs, 8 columns per K = 16 step
        const uint32_t ta = tmem + 256 + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < K / 16; ++c) {
            uint32_t pk[8];
            for (int i = 0; i < 8; ++i)
                pk[i] = tc::pack_bf16(A[(size_t)tid * K + 16 * c + 2 * i], A[(size_t)tid * K + 16 * c + 2 * i + 1]);
            tc::tmem_st8(ta + 8 * c, pk);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
    }
    if (tid == 0) {
        tc::tc_fence_after();
        if (a_in_tmem) {
            for (int s = 0; s < K / 16; ++s) {
                uint64_t bd = tc::smem_desc(tc::smem_u32(Bs) + (uint32_t)(2 * s) * N * 16, N * 16, 128);
                tc::umma_bf16_ts(tmem, tmem + 256 + 8 * s, bd, tc::idesc_bf16(128, N), s > 0 ? 1u : 0u);
            }
        } else {
            tc::umma_gemm(tmem, tc::smem_u32(As), 128, tc::smem_u32(Bs), N, K, tc::idesc_bf16(128, N));
        }
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
        tc::tmem_ld_wait();
        for (int i = 0; i < 32; ++i) D[(size_t)tid * N + c0 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace aline

static int tc_selftest_launch(const float* A, const float* B, int32_t N, int32_t K, float* D, const void* B_packed,
                              int a_in_tmem, void* stream) {
    using namespace aline;
    ALINE_REQUIRE(A && B && D, "aline_tc_selftest: NULL tensor");
    ALINE_REQUIRE(N >= 32 && N <= 256 && N % 32 == 0 && K >= 16 && K % 16 == 0 && K <= 256,
                  "aline_tc_selftest: need N in 32..256 (multiple of 32) and K in 16..256 (multiple of 16)");
    size_t smem = (size_t)(128 + N) * K * 2;
    if (smem > 48 * 1024)
        ALINE_CHECK_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, N, K, D, B_packed != nullptr,
                                                               (const unsigned char*)B_packed, a_in_tmem);
    ALINE_LAUNCH_OK();
    return 0;
}

extern "C" int aline_tc_selftest(const float* A, const float* B, int32_t N, int32_t K, float* D, const void* B_packed,
                                 void* stream) {
    return tc_selftest_launch(A, B, N, K, D, B_packed, 0, stream);
}

extern "C" int aline_tc_selftest_tmem_a(const float* A, const float* B, int32_t N, int32_t K, float* D, void* stream) {
    return tc_selftest_launch(A, B, N, K, D, nullptr, 1, stream);
}
