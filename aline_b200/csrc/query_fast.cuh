// Pieces shared by the fast tcgen05 candidate-query kernels (csrc/query_tc3.cu: one thread per candidate row;
// csrc/query_tc4.cu: two threads per row): the bf16 weight-blob shape, the key / value operand-block sizes and the
// bf16 packing helpers.  Operand layout: csrc/tc.cuh.
#pragma once
#include "model.cuh"
#include "tc.cuh"

namespace aline {
namespace tcq {

constexpr int kT2D = 32;
constexpr int kT2Tile = 128;
constexpr int kT2Chunk = kT2Tile * 16;      // bytes of one 8-column chunk of a 128-row operand tile

struct Tc2Shape {
    int FF, HH, NL;
    int off_wq, off_wo, off_w1, off_w2, layer_bytes, off_acq, total_bytes;      // bytes inside the bf16 blob
    int vec_layer, v_acq_w2, v_acq_b2, vec_total;                               // floats inside the fp32 block
};

__host__ __device__ inline Tc2Shape make_tc2_shape(const Dims& m) {
    Tc2Shape s;
    const int D = m.D, KA = D + 16;          // 32 (csrc/query_tc3.cu, query_tc4.cu) or 64 (csrc/query_tc5.cu)
    s.FF = m.FF; s.HH = m.HH; s.NL = m.NL;
    s.off_wq = 0;
    s.off_wo = s.off_wq + D * KA * 2;
    s.off_w1 = s.off_wo + D * KA * 2;
    s.off_w2 = s.off_w1 + m.FF * KA * 2;
    s.layer_bytes = s.off_w2 + D * (m.FF + 16) * 2;
    s.off_acq = s.layer_bytes * m.NL;
    s.total_bytes = s.off_acq + m.HH * KA * 2;
    s.vec_layer = 4 * D;                     // g1, be1, g2, be2
    s.v_acq_w2 = s.vec_layer * m.NL;
    s.v_acq_b2 = s.v_acq_w2 + m.HH;
    s.vec_total = s.v_acq_b2 + 4;
    return s;
}

// bytes of the bf16 key / value operand block of one (layer, rollout): K part 5 chunks x nkp rows x 16 B
// (4 heads + mask chunk), V part 4 heads x nkp/8 chunks x 16 rows x 16 B (8 features, ones row, 7 zero rows)
__host__ __device__ inline int tc2_kv_block_bytes(int nkp) { return 208 * nkp; }
__host__ __device__ inline int tc2_k_bytes(int nkp) { return 80 * nkp; }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float ex2f(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 8 fp32 -> one 16-byte chunk row of a 128-row operand tile
__device__ __forceinline__ void store_chunk(unsigned char* tile, int chunk, int r, const float* v) {
    uint4 q;
    q.x = pack2(v[0], v[1]); q.y = pack2(v[2], v[3]); q.z = pack2(v[4], v[5]); q.w = pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + (size_t)chunk * kT2Chunk + (size_t)r * 16) = q;
}
__device__ __forceinline__ void store_chunk_relu(unsigned char* tile, int chunk, int r, const float* v) {
    uint4 q;
    q.x = pack2_relu(v[0], v[1]); q.y = pack2_relu(v[2], v[3]); q.z = pack2_relu(v[4], v[5]); q.w = pack2_relu(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + (size_t)chunk * kT2Chunk + (size_t)r * 16) = q;
}

}  // namespace tcq
}  // namespace aline
