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

// ---- folded operands of csrc/query_tc3.cu (FOLD): K'_h = c Wq_h^T (K_h - K_0h), V'_h = V_h Wo_h^T ----
// Emitted by the context kernels (one block = one rollout) right after the plain bf16 operand blocks, from those blocks
// (already in L2) and the fp32 parameters: the fold region follows the NL x B plain blocks in the same buffer, one
// kFoldKeyBytes * nkf block per (layer, rollout) (nkf = the key count padded to 8 -- the MMA shapes only need 4 nkf to be
// a multiple of 16 -- or to 16 for the two-warpgroup form) = the shared-memory image the query stream loads with one bulk copy:
//   K' [4 nkp rows (head, key)] x [48 columns: 32 token features | bias_hi, bias_lo, 0 x 6 | 0 x 8]  (6 chunks)
//   V' [32 rows (output features)] x [4 nkp columns (head, key)]                                      (nkp / 2 chunks)
// Rows of slots that hold no key: K' = 0 with -200 in the bias column (probability 2^-200 = 0), V' = 0.
// Work items (layer, part, head, key), part 0..3 = K' chunk, 4..7 = 8 V' features: a warp shares the weights it reads.
constexpr int kFoldKeyBytes = 640;           // per key and layer: 4 heads x (6 chunks x 16 B of K' + 32 features x 2 B of V')
constexpr float kFoldScale = 0.51006973272324049f;      // log2(e) / sqrt(8)
__host__ __device__ inline size_t tc2_fold_offset(int NL, int B, int nkp) { return (size_t)NL * B * tc2_kv_block_bytes(nkp); }

// nkp: padding of the plain blocks (multiple of 16); nkf: padding of the folded blocks (multiple of 8, <= nkp)
__device__ __forceinline__ void fold_kv_emit(unsigned char* tckv, int b, int B, int nkp, int nkf, int NL,
                                             const float* __restrict__ P, const Layout& L, int tid, int nthreads) {
    const int kvblk = tc2_kv_block_bytes(nkp);
    unsigned char* fold = tckv + tc2_fold_offset(NL, B, nkp);
    const uint32_t chunk = (uint32_t)(4 * nkf) * 16u;
    for (int it = tid; it < NL * 32 * nkf; it += nthreads) {
        const int key = it % nkf, h = (it / nkf) & 3, part = (it / (4 * nkf)) & 7, l = it / (32 * nkf);
        const unsigned char* blk = tckv + ((size_t)l * B + b) * kvblk;
        const unsigned short* vb = reinterpret_cast<const unsigned short*>(blk + tc2_k_bytes(nkp)) +
                                   ((size_t)h * (nkp / 8) + (key >> 3)) * 128 + (key & 7);
        const bool used = __ldcg(vb + 64) != 0;                          // the "ones" row marks the slots that hold a key
        const float* Pl = P + L.layer0 + (size_t)l * L.layer_stride;
        unsigned char* kp = fold + ((size_t)l * B + b) * ((size_t)kFoldKeyBytes * nkf);
        const int n = h * nkf + key;
        if (part < 4) {
            const uint4 kq = __ldcg(reinterpret_cast<const uint4*>(blk + ((size_t)h * nkp + key) * 16));
            const uint32_t w[4] = {kq.x, kq.y, kq.z, kq.w};
            float kd[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) { kd[2 * e] = __uint_as_float(w[e] << 16); kd[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
            float o[8];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const float4* wr = reinterpret_cast<const float4*>(Pl + L.wq + (size_t)(8 * part + ii) * kT2D + 8 * h);
                const float4 w0 = __ldg(wr), w1 = __ldg(wr + 1);
                float a = w0.x * kd[0];
                a = fmaf(w0.y, kd[1], a); a = fmaf(w0.z, kd[2], a); a = fmaf(w0.w, kd[3], a);
                a = fmaf(w1.x, kd[4], a); a = fmaf(w1.y, kd[5], a); a = fmaf(w1.z, kd[6], a); a = fmaf(w1.w, kd[7], a);
                o[ii] = used ? a * kFoldScale : 0.f;
            }
            uint4 q;
            q.x = pack2(o[0], o[1]); q.y = pack2(o[2], o[3]); q.z = pack2(o[4], o[5]); q.w = pack2(o[6], o[7]);
            *reinterpret_cast<uint4*>(kp + part * chunk + (size_t)n * 16) = q;
            if (part == 0) {
                float bias = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) bias = fmaf(__ldg(Pl + L.bq + 8 * h + e), kd[e], bias);
                bias *= kFoldScale;
                const float hi = __bfloat162float(__float2bfloat16_rn(bias));
                *reinterpret_cast<uint4*>(kp + 4 * chunk + (size_t)n * 16) =
                    make_uint4(used ? pack2(hi, bias - hi) : 0xC348u, 0u, 0u, 0u);           // 0xC348 = bf16(-200)
                *reinterpret_cast<uint4*>(kp + 5 * chunk + (size_t)n * 16) = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
            float v[8];
#pragma unroll
            for (int f = 0; f < 8; ++f) v[f] = __uint_as_float((uint32_t)__ldcg(vb + f * 8) << 16);
            const int o0 = 8 * (part - 4);
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4* wr = reinterpret_cast<const float4*>(Pl + L.wo + (size_t)(8 * h + e) * kT2D + o0);
                const float4 w0 = __ldg(wr), w1 = __ldg(wr + 1);
                a[0] = fmaf(v[e], w0.x, a[0]); a[1] = fmaf(v[e], w0.y, a[1]); a[2] = fmaf(v[e], w0.z, a[2]); a[3] = fmaf(v[e], w0.w, a[3]);
                a[4] = fmaf(v[e], w1.x, a[4]); a[5] = fmaf(v[e], w1.y, a[5]); a[6] = fmaf(v[e], w1.z, a[6]); a[7] = fmaf(v[e], w1.w, a[7]);
            }
            // V' column n of the [32 x 4 nkp] operand (K-major: 8 keys of a feature row = one 16-byte unit)
            __nv_bfloat16* vo = reinterpret_cast<__nv_bfloat16*>(kp + 384 * nkf + (size_t)(n >> 3) * (kT2D * 16)) + (n & 7);
#pragma unroll
            for (int j = 0; j < 8; ++j) vo[(o0 + j) * 8] = __float2bfloat16_rn(used ? a[j] : 0.f);
        }
    }
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
