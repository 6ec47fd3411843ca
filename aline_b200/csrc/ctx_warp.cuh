// Pieces shared by the warp-per-token context kernels (csrc/ctx_warp.cu: d = 32 / 4 heads, csrc/ctx_warp64.cu: d = 64 / 8
// heads): attention of a warp's tokens over the context keys with head_dim 8.
#pragma once
#include "common.cuh"

namespace aline {

// softmax(q K^T) V over the n_c context keys for NP tokens of one warp, FOUR heads of 8 features (32 consecutive
// features of rows with stride KS; d = 64 calls it once per half of the features).  qrow[i]: the token's scaled query (shared row);
// lane = (head h = lane >> 3, jj = lane & 7) scores keys jj, jj + 8, ...; on return o[i] = attention output feature `lane`.
template <int NP, int KS>
__device__ __forceinline__ void warp_attention(float (&o)[NP], const float* const (&qrow)[NP], const float* Ks,
                                               const float* Vs, int n_c, int lane) {
    const int h = lane >> 3, jj = lane & 7;
    float q[NP][8], s[NP][8];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(qrow[i] + 8 * h), c = *reinterpret_cast<const float4*>(qrow[i] + 8 * h + 4);
        q[i][0] = a.x; q[i][1] = a.y; q[i][2] = a.z; q[i][3] = a.w; q[i][4] = c.x; q[i][5] = c.y; q[i][6] = c.z; q[i][7] = c.w;
    }
    float mx[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) mx[i] = -INFINITY;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
        if (8 * sl < n_c) {
            const int j = 8 * sl + jj, jc = j < n_c ? j : n_c - 1;
            const float4 a = *reinterpret_cast<const float4*>(Ks + jc * KS + 8 * h);
            const float4 c = *reinterpret_cast<const float4*>(Ks + jc * KS + 8 * h + 4);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                float d = q[i][0] * a.x;
                d = fmaf(q[i][1], a.y, d); d = fmaf(q[i][2], a.z, d); d = fmaf(q[i][3], a.w, d);
                d = fmaf(q[i][4], c.x, d); d = fmaf(q[i][5], c.y, d); d = fmaf(q[i][6], c.z, d); d = fmaf(q[i][7], c.w, d);
                s[i][sl] = j < n_c ? d : -INFINITY;
                mx[i] = fmaxf(mx[i], s[i][sl]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NP; ++i) s[i][sl] = -INFINITY;
        }
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
        for (int i = 0; i < NP; ++i) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], off));
    }
    float den[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) den[i] = 0.f;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
        if (8 * sl < n_c) {
#pragma unroll
            for (int i = 0; i < NP; ++i) { s[i][sl] = expf(s[i][sl] - mx[i]); den[i] += s[i][sl]; }
        }
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
        for (int i = 0; i < NP; ++i) den[i] += __shfl_xor_sync(0xffffffffu, den[i], off);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) o[i] = 0.f;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
        if (8 * sl < n_c) {
#pragma unroll
            for (int j2 = 0; j2 < 8; ++j2) {
                const int j = 8 * sl + j2;
                if (j < n_c) {
                    const float v = Vs[j * KS + lane];
#pragma unroll
                    for (int i = 0; i < NP; ++i)
                        o[i] = fmaf(__shfl_sync(0xffffffffu, s[i][sl], (lane & 24) | j2), v, o[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) o[i] *= 1.0f / den[i];
}

}  // namespace aline
