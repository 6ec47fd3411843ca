// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
// SC'11; the generator behind curand's and torch's CUDA streams), restated from the published algorithm.
// Stream layout of this library: key = 64-bit seed; counter = (row_lo, row_hi, column b, call index), where `row` is the
// GLOBAL index of the contrastive draw -- so a draw does not depend on how the rows are sharded over ranks or passes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace aline {

struct Philox4 { uint32_t x[4]; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.x[0] = c0; o.x[1] = c1; o.x[2] = c2; o.x[3] = c3;
    return o;
}

// 24-bit uniform in [0, 1) (the resolution of torch.rand for float32)
__host__ __device__ inline float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

}  // namespace aline
