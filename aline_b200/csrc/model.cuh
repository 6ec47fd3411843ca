// ALINE model description, packed-parameter layout and the per-token fp32 building blocks
// shared by the context/target stack kernel and the candidate-query stream kernel.
//
// reference: model/embedder.py:47-57 (x / y MLPs), model/encoder.py:76-79 (post-norm
// nn.TransformerEncoderLayer, ReLU, eps 1e-5), model/head.py:27-33 (acquisition MLP),
// model/head.py:214-224 (GMM heads).  Arithmetic: SURVEY.md appendix A.1 / A.2.
//
// Execution model of the fp32 path: one thread owns one token.  Its activation vector
// lives in a shared-memory *column* X[k][tid] (conflict-free: consecutive threads,
// consecutive addresses); the weights of the current layer are staged once per block
// in shared memory, transposed to [in][out], and read as warp-wide broadcasts
// (LDS.128 -> 4 FMAs).  Accumulators are register arrays with static indices.
#pragma once
#include "common.cuh"

namespace aline {

struct Dims {
    int D, FF, H, NL, dx, dy, ntok, C, EH, HH, tt;
    float std_min;
};

// Offsets (in floats) into the packed parameter blob.  Python packs in exactly this order
// (aline_b200/model/packing.py); every matrix is stored transposed, [in][out].
struct Layout {
    // embedders
    size_t x_w1, x_b1, x_w2, x_b2, y_w1, y_b1, y_w2, y_b2, tok;
    // encoder layer l at layer0 + l * layer_stride, fields relative to the layer base
    size_t layer0, layer_stride;
    size_t wq, wk, wv, bq, bk, bv, wo, bo, g1, be1, w1, b1, w2, b2, g2, be2;
    // acquisition head
    size_t a_w1, a_b1, a_w2, a_b2;
    // GMM head c at gmm0 + c * gmm_stride: w1 [D][HH], b1 [HH], w2 [3][HH], b2 [3] (+1 pad)
    size_t gmm0, gmm_stride, g_w1, g_b1, g_w2, g_b2;
    size_t total;
};

__host__ __device__ inline size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }

__host__ __device__ inline Layout make_layout(const Dims& m) {
    Layout L;
    size_t o = 0;
    const size_t D = m.D, FF = m.FF, EH = m.EH, HH = m.HH;
    L.x_w1 = o; o += pad4((size_t)m.dx * EH);
    L.x_b1 = o; o += EH;
    L.x_w2 = o; o += EH * D;
    L.x_b2 = o; o += D;
    L.y_w1 = o; o += pad4((size_t)m.dy * EH);
    L.y_b1 = o; o += EH;
    L.y_w2 = o; o += EH * D;
    L.y_b2 = o; o += D;
    L.tok = o; o += pad4((size_t)m.ntok * D);
    L.layer0 = o;
    size_t r = 0;
    L.wq = r; r += D * D;
    L.wk = r; r += D * D;
    L.wv = r; r += D * D;
    L.bq = r; r += D;
    L.bk = r; r += D;
    L.bv = r; r += D;
    L.wo = r; r += D * D;
    L.bo = r; r += D;
    L.g1 = r; r += D;
    L.be1 = r; r += D;
    L.w1 = r; r += D * FF;
    L.b1 = r; r += FF;
    L.w2 = r; r += FF * D;
    L.b2 = r; r += D;
    L.g2 = r; r += D;
    L.be2 = r; r += D;
    L.layer_stride = r;
    o += r * m.NL;
    L.a_w1 = o; o += (D + m.tt) * HH;
    L.a_b1 = o; o += HH;
    L.a_w2 = o; o += HH;
    L.a_b2 = o; o += 4;
    L.gmm0 = o;
    r = 0;
    L.g_w1 = r; r += D * HH;
    L.g_b1 = r; r += HH;
    L.g_w2 = r; r += 3 * HH;
    L.g_b2 = r; r += 4;
    L.gmm_stride = r;
    o += r * m.C;
    L.total = o;
    return L;
}

// cooperative global -> shared copy of n floats (n % 4 == 0, both 16-byte aligned)
__device__ __forceinline__ void stage_floats(float* dst, const float* __restrict__ src, int n) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

// y[0..OUT) += sum_k xcol[k * stride] * WT[k * ldw + j]      (x from a shared-memory column, W broadcast)
template <int OUT>
__device__ __forceinline__ void matvec_col(float (&y)[OUT], const float* xcol, int stride, int IN,
                                           const float* WT, int ldw) {
#pragma unroll 4
    for (int k = 0; k < IN; ++k) {
        const float xk = xcol[k * stride];
        const float4* w = reinterpret_cast<const float4*>(WT + (size_t)k * ldw);
#pragma unroll
        for (int j = 0; j < OUT / 4; ++j) {
            float4 ww = w[j];
            y[4 * j + 0] = fmaf(xk, ww.x, y[4 * j + 0]);
            y[4 * j + 1] = fmaf(xk, ww.y, y[4 * j + 1]);
            y[4 * j + 2] = fmaf(xk, ww.z, y[4 * j + 2]);
            y[4 * j + 3] = fmaf(xk, ww.w, y[4 * j + 3]);
        }
    }
}

// y[0..OUT) += sum_{k<IN} x[k] * WT[k * ldw + j]   with x in registers (IN compile-time, fully unrolled)
template <int IN, int OUT>
__device__ __forceinline__ void matvec_reg(float (&y)[OUT], const float (&x)[IN], const float* WT, int ldw) {
#pragma unroll
    for (int k = 0; k < IN; ++k) {
        const float4* w = reinterpret_cast<const float4*>(WT + (size_t)k * ldw);
#pragma unroll
        for (int j = 0; j < OUT / 4; ++j) {
            float4 ww = w[j];
            y[4 * j + 0] = fmaf(x[k], ww.x, y[4 * j + 0]);
            y[4 * j + 1] = fmaf(x[k], ww.y, y[4 * j + 1]);
            y[4 * j + 2] = fmaf(x[k], ww.z, y[4 * j + 2]);
            y[4 * j + 3] = fmaf(x[k], ww.w, y[4 * j + 3]);
        }
    }
}

template <int N>
__device__ __forceinline__ void load_vec(float (&y)[N], const float* v) {
#pragma unroll
    for (int j = 0; j < N / 4; ++j) {
        float4 t = reinterpret_cast<const float4*>(v)[j];
        y[4 * j] = t.x; y[4 * j + 1] = t.y; y[4 * j + 2] = t.z; y[4 * j + 3] = t.w;
    }
}

// LayerNorm over D features, biased variance, eps 1e-5 (torch native_layer_norm)
template <int D>
__device__ __forceinline__ void layer_norm(float (&v)[D], const float* g, const float* b) {
    float mu = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) mu += v[i];
    mu *= (1.0f / D);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) { float d = v[i] - mu; var = fmaf(d, d, var); }
    const float rstd = 1.0f / sqrtf(var * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = (v[i] - mu) * rstd * g[i] + b[i];
}

// 2-layer MLP embedder  in -> EH (ReLU) -> D, input in registers (IN <= 8), result ADDED to out[D]
template <int D>
__device__ __forceinline__ void embed_mlp(float (&out)[D], const float* xin, int IN, const float* W1T, const float* b1,
                                          const float* W2T, const float* b2, int EH) {
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] += b2[i];
    for (int c = 0; c < EH; c += 32) {
        float h[32];
        load_vec<32>(h, b1 + c);
        for (int k = 0; k < IN; ++k) {
            const float xk = xin[k];
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = fmaf(xk, W1T[k * EH + c + j], h[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
        matvec_reg<32, D>(out, h, W2T + (size_t)c * D, D);
    }
}

// One post-norm encoder layer for the token whose activation column is xcol (stride = tokens per block).
//   q = (x Wq^T + bq) / sqrt(8);  per head softmax(q K^T) V over the n_keys rows of Ks / Vs (shared, [key][D]);
//   h = LN1(x + o Wo^T + bo);  x' = LN2(h + W2 relu(W1 h + b1) + b2)
// tcol is a second per-token scratch column.  On return xcol holds x'.
template <int D>
__device__ __forceinline__ void encoder_layer_token(float* xcol, float* tcol, int stride, const float* W, const Layout& L,
                                                    int FF, const float* Ks, const float* Vs, int n_keys) {
    constexpr int H = D / 8;
    float q[D];
    load_vec<D>(q, W + L.bq);
    matvec_col<D>(q, xcol, stride, D, W + L.wq, D);
#pragma unroll
    for (int i = 0; i < D; ++i) q[i] *= 0.35355339059327376220f;      // 1/sqrt(head_dim = 8)

    // pass 1: per-head maximum of the scores
    float mx[H];
#pragma unroll
    for (int h = 0; h < H; ++h) mx[h] = -INFINITY;
    for (int j = 0; j < n_keys; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + (size_t)j * D);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float4 a = kr[2 * h], b = kr[2 * h + 1];
            float s = q[8 * h] * a.x;
            s = fmaf(q[8 * h + 1], a.y, s); s = fmaf(q[8 * h + 2], a.z, s); s = fmaf(q[8 * h + 3], a.w, s);
            s = fmaf(q[8 * h + 4], b.x, s); s = fmaf(q[8 * h + 5], b.y, s); s = fmaf(q[8 * h + 6], b.z, s);
            s = fmaf(q[8 * h + 7], b.w, s);
            mx[h] = fmaxf(mx[h], s);
        }
    }
    // pass 2: exp, normaliser, weighted values
    float acc[D], den[H];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) den[h] = 0.f;
    for (int j = 0; j < n_keys; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + (size_t)j * D);
        const float4* vr = reinterpret_cast<const float4*>(Vs + (size_t)j * D);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float4 a = kr[2 * h], b = kr[2 * h + 1];
            float s = q[8 * h] * a.x;
            s = fmaf(q[8 * h + 1], a.y, s); s = fmaf(q[8 * h + 2], a.z, s); s = fmaf(q[8 * h + 3], a.w, s);
            s = fmaf(q[8 * h + 4], b.x, s); s = fmaf(q[8 * h + 5], b.y, s); s = fmaf(q[8 * h + 6], b.z, s);
            s = fmaf(q[8 * h + 7], b.w, s);
            const float p = expf(s - mx[h]);
            den[h] += p;
            float4 va = vr[2 * h], vb = vr[2 * h + 1];
            acc[8 * h + 0] = fmaf(p, va.x, acc[8 * h + 0]); acc[8 * h + 1] = fmaf(p, va.y, acc[8 * h + 1]);
            acc[8 * h + 2] = fmaf(p, va.z, acc[8 * h + 2]); acc[8 * h + 3] = fmaf(p, va.w, acc[8 * h + 3]);
            acc[8 * h + 4] = fmaf(p, vb.x, acc[8 * h + 4]); acc[8 * h + 5] = fmaf(p, vb.y, acc[8 * h + 5]);
            acc[8 * h + 6] = fmaf(p, vb.z, acc[8 * h + 6]); acc[8 * h + 7] = fmaf(p, vb.w, acc[8 * h + 7]);
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float inv = 1.0f / den[h];
#pragma unroll
        for (int i = 0; i < 8; ++i) tcol[(8 * h + i) * stride] = acc[8 * h + i] * inv;
    }
    // out-projection + residual + LN1
    float hv[D];
    load_vec<D>(hv, W + L.bo);
    matvec_col<D>(hv, tcol, stride, D, W + L.wo, D);
#pragma unroll
    for (int i = 0; i < D; ++i) hv[i] += xcol[i * stride];
    layer_norm<D>(hv, W + L.g1, W + L.be1);
#pragma unroll
    for (int i = 0; i < D; ++i) tcol[i * stride] = hv[i];
    // feed-forward in chunks of 32 hidden units, + residual + LN2
    float out[D];
    load_vec<D>(out, W + L.b2);
    for (int c = 0; c < FF; c += 32) {
        float hid[32];
        load_vec<32>(hid, W + L.b1 + c);
        matvec_col<32>(hid, tcol, stride, D, W + L.w1 + c, FF);
#pragma unroll
        for (int j = 0; j < 32; ++j) hid[j] = fmaxf(hid[j], 0.f);
        matvec_reg<32, D>(out, hid, W + L.w2 + (size_t)c * D, D);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] += tcol[i * stride];
    layer_norm<D>(out, W + L.g2, W + L.be2);
#pragma unroll
    for (int i = 0; i < D; ++i) xcol[i * stride] = out[i];
}

}  // namespace aline
