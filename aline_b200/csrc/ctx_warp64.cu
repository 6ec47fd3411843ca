// Context + target stack, warp-per-token fp32 kernel for d = 64 / 8 heads (the trained psychometric variant of the
// reference, notebooks/eval_psychometric.ipynb cell 5: dim_embedding 64, dim_feedforward 128, n_head 8).
//
// Same contract as ctx_stack_kernel (csrc/rollout.cu) and the same mapping idea as ctx_stack_warp_kernel
// (csrc/ctx_warp.cu; reference: model/embedder.py:128-214 + model/encoder.py:128-141 restricted to the context / target
// rows): a warp owns NTK tokens for the whole kernel, lane f owns features f and f + 32 of each, so a weight row is two
// conflict-free 128-byte wavefronts feeding 2 NTK FMAs, inputs are LDS.128 broadcasts, the 128 hidden units of the MLPs
// are 4 per lane, attention runs as two passes of four heads (lane = (head, key mod 8)), LayerNorm is a warp reduction.
//
// What differs from d = 32 is the weight traffic: a layer is 134 KB of fp32 weights (d = 32: 50 KB), and two blocks
// must share an SM (200 rollouts on 148 SMs would otherwise take two waves of a latency-bound kernel).  The weights
// therefore stream through a three-slot ring of ~17 KB segments -- ONE 64 x 64 matrix (or half of a 64 x 128 / 128 x 64
// one) per segment, 8 per layer -- by TMA bulk copies, each slot refilled as soon as every warp is past its segment;
// accumulators that span two segments (the two halves of W1 / W2) stay in registers across the block barrier, which is
// why a warp keeps its tokens for the whole kernel (n_rows <= 16 warps x 4 tokens).  The old kernel for this shape
// (thread = (token, head), every lane walking the full k loop) took 454 us per step at the cfg5 launch shape.
#include "model.cuh"
#include "tc.cuh"
#include "select.cuh"
#include "ctx_warp.cuh"
#include <cstdlib>

namespace aline {

constexpr int kC6D = 64;
constexpr int kC6KS = 68;          // padded K / V row stride in floats (16-byte reads of 8 rows hit 32 distinct banks)
constexpr int kC6MaxWarps = 16;
constexpr int kC6Mat = 64 * 64;    // floats of one segment's matrix part

// acc[i][j] += sum_{k < K} row_i[k] * W[k * 64 + lane + 32 j]     (row_i: shared, 16-byte aligned, read as broadcasts)
// two partial sums (even / odd k) per accumulator as the halves of packed fp32x2 operands
template <int NTK, int K>
__device__ __forceinline__ void warp_matvec64(float (&acc)[NTK][2], const float* const (&row)[NTK], const float* W, int lane) {
    f32x2 a2[NTK][2];
#pragma unroll
    for (int i = 0; i < NTK; ++i) { a2[i][0] = pk2(acc[i][0], 0.f); a2[i][1] = pk2(acc[i][1], 0.f); }
#pragma unroll 4
    for (int k4 = 0; k4 < K / 4; ++k4) {
        float4 xv[NTK];
#pragma unroll
        for (int i = 0; i < NTK; ++i) xv[i] = *reinterpret_cast<const float4*>(row[i] + 4 * k4);
        const float* w = W + (4 * k4) * kC6D + lane;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const f32x2 w01 = pk2(w[32 * j], w[kC6D + 32 * j]);
            const f32x2 w23 = pk2(w[2 * kC6D + 32 * j], w[3 * kC6D + 32 * j]);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                a2[i][j] = fma2(pk2(xv[i].x, xv[i].y), w01, a2[i][j]);
                a2[i][j] = fma2(pk2(xv[i].z, xv[i].w), w23, a2[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NTK; ++i) {
        float lo, hi;
        upk2(a2[i][0], lo, hi); acc[i][0] = lo + hi;
        upk2(a2[i][1], lo, hi); acc[i][1] = lo + hi;
    }
}

// hid[i][0..3] += sum_{k < K} row_i[k] * W1[k * 128 + 4 lane + (0..3)]     (hidden units 4 lane .. 4 lane + 3)
template <int NTK, int K>
__device__ __forceinline__ void warp_hidden64(float (&hid)[NTK][4], const float* const (&row)[NTK], const float* W1, int lane) {
#pragma unroll 2
    for (int k4 = 0; k4 < K / 4; ++k4) {
        float4 xv[NTK];
#pragma unroll
        for (int i = 0; i < NTK; ++i) xv[i] = *reinterpret_cast<const float4*>(row[i] + 4 * k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 w = *reinterpret_cast<const float4*>(W1 + (size_t)(4 * k4 + kk) * 128 + 4 * lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const float xk = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
                const f32x2 xx = pk2(xk, xk);
                f32x2 h01 = pk2(hid[i][0], hid[i][1]), h23 = pk2(hid[i][2], hid[i][3]);
                h01 = fma2(xx, pk2(w.x, w.y), h01); h23 = fma2(xx, pk2(w.z, w.w), h23);
                upk2(h01, hid[i][0], hid[i][1]); upk2(h23, hid[i][2], hid[i][3]);
            }
        }
    }
}

// LayerNorm over the 64 features of NTK tokens (two per lane; biased variance, eps 1e-5)
template <int NTK>
__device__ __forceinline__ void warp_layer_norm64(float (&v)[NTK][2], const float* g, const float* be, int lane) {
    float mu[NTK], q[NTK];
#pragma unroll
    for (int i = 0; i < NTK; ++i) mu[i] = v[i][0] + v[i][1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NTK; ++i) mu[i] += __shfl_xor_sync(0xffffffffu, mu[i], o);
    }
#pragma unroll
    for (int i = 0; i < NTK; ++i) {
        v[i][0] -= mu[i] * (1.0f / 64); v[i][1] -= mu[i] * (1.0f / 64);
        q[i] = fmaf(v[i][0], v[i][0], v[i][1] * v[i][1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NTK; ++i) q[i] += __shfl_xor_sync(0xffffffffu, q[i], o);
    }
    const float g0 = g[lane], g1 = g[lane + 32], b0 = be[lane], b1 = be[lane + 32];
#pragma unroll
    for (int i = 0; i < NTK; ++i) {
        const float rstd = 1.0f / sqrtf(q[i] * (1.0f / 64) + 1e-5f);
        v[i][0] = v[i][0] * rstd * g0 + b0;
        v[i][1] = v[i][1] * rstd * g1 + b1;
    }
}

// embedder MLP  in(IN <= 8, registers) -> 128 (ReLU) -> hs (warp-private shared rows of 128 floats)
template <int NTK>
__device__ __forceinline__ void warp_embed_hidden(const float (&xin)[NTK][8], int IN, const float* W1, const float* b1,
                                                  float* hs, int lane) {
    float hid[NTK][4];
    const float4 bb = *reinterpret_cast<const float4*>(b1 + 4 * lane);
#pragma unroll
    for (int i = 0; i < NTK; ++i) { hid[i][0] = bb.x; hid[i][1] = bb.y; hid[i][2] = bb.z; hid[i][3] = bb.w; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < IN) {
            const float4 w = *reinterpret_cast<const float4*>(W1 + (size_t)k * 128 + 4 * lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                hid[i][0] = fmaf(xin[i][k], w.x, hid[i][0]); hid[i][1] = fmaf(xin[i][k], w.y, hid[i][1]);
                hid[i][2] = fmaf(xin[i][k], w.z, hid[i][2]); hid[i][3] = fmaf(xin[i][k], w.w, hid[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NTK; ++i)
        *reinterpret_cast<float4*>(hs + i * 128 + 4 * lane) =
            make_float4(fmaxf(hid[i][0], 0.f), fmaxf(hid[i][1], 0.f), fmaxf(hid[i][2], 0.f), fmaxf(hid[i][3], 0.f));
    __syncwarp();
}

// Weight segments in stream order: 4 for the embedders, then 8 per layer.  Every segment starts with its matrix part.
//   0: x_w1 | x_b1 | x_w2 rows 0..63      1: x_w2 rows 64..127 | x_b2      2, 3: the same for the y embedder
//   layer: 0 Wq   1 Wk   2 Wv | bq bk bv   3 Wo | bo g1 be1   4 W1 rows 0..31   5 W1 rows 32..63 | b1
//          6 W2 rows 0..63   7 W2 rows 64..127 | b2 g2 be2
__host__ __device__ inline void c6_segment(const Layout& L, int s, size_t& off, size_t& n) {
    if (s < 4) {
        const bool y = s >= 2;
        const size_t beg = y ? L.y_w1 : L.x_w1, w2 = y ? L.y_w2 : L.x_w2, end = y ? L.tok : L.y_w1;
        if ((s & 1) == 0) { off = beg; n = w2 + kC6Mat - beg; }
        else { off = w2 + kC6Mat; n = end - off; }
        return;
    }
    const int l = (s - 4) >> 3, k = (s - 4) & 7;
    const size_t base = L.layer0 + (size_t)l * L.layer_stride;
    switch (k) {
        case 0: off = L.wq; n = kC6Mat; break;
        case 1: off = L.wk; n = kC6Mat; break;
        case 2: off = L.wv; n = L.wo - L.wv; break;
        case 3: off = L.wo; n = L.w1 - L.wo; break;
        case 4: off = L.w1; n = kC6Mat; break;
        case 5: off = L.w1 + kC6Mat; n = L.w2 - off; break;
        case 6: off = L.w2; n = kC6Mat; break;
        default: off = L.w2 + kC6Mat; n = L.layer_stride - off; break;
    }
    off += base;
}

template <int NTK>
__global__ void __launch_bounds__(32 * kC6MaxWarps, 2)
ctx_stack_warp64_kernel(const Dims m, const Layout L, const float* __restrict__ P, const float* cx,
                        const float* cy, int n_c, int ctx_cap, const float* __restrict__ target_x, int n_td,
                        const int* __restrict__ tgt_slot, float* __restrict__ kv, int kv_slots, int B,
                        float* __restrict__ z_tgt, float* __restrict__ z_ctx, int WB, int n_slots,
                        unsigned char* __restrict__ tckv, int n_keys_tc, const SelectArgs sel, int do_select) {
    constexpr int D = kC6D, KS = kC6KS, G = 8;
    constexpr int NP = NTK >= 2 ? 2 : 1;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int b = blockIdx.x;
    const int n_t = n_td + m.ntok, n_tok = n_c + n_t;
    float* Wbuf = smem;                                   // 3 ring slots of WB floats
    float* X = Wbuf + 3 * (size_t)WB;                     // [n_tok][64] layer input / output
    float* T = X + (size_t)n_tok * D;                     // [n_tok][64] scaled query, then attention output, then LN1 output
    float* Hs = T + (size_t)n_tok * D;                    // [NW][NTK][128] hidden units (warp-private)
    float* Ks = Hs + (size_t)NW * NTK * 128;              // [n_slots][68]
    float* Vs = Ks + (size_t)n_slots * KS;                // [n_slots][68]
    int* slot_s = reinterpret_cast<int*>(Vs + (size_t)n_slots * KS);      // [n_tok] key slot of a token (-1: not attended)
    int* orig_s = slot_s + n_tok;                                         // [n_tok] token id of a processed row
    __shared__ int n_eff_s;
    float* hs = Hs + (size_t)warp * NTK * 128;

    pdl_trigger();                                        // the next kernel of the stream may start its own prologue
    const bool rollout_mode = z_tgt == nullptr && z_ctx == nullptr;   // nothing downstream of the last layer's K, V
    const bool ctx_last = z_ctx != nullptr;               // the value head reads the context rows' final encodings
    const int n_seg = 4 + 8 * m.NL - (rollout_mode ? 5 : 0);
    auto issue = [&](int s) {                             // one thread
        if (s >= n_seg) return;
        size_t off, n;
        c6_segment(L, s, off, n);
        tc::mbar_arrive_expect_tx(&bar[s % 3], (uint32_t)n * 4u);
        tc::bulk_g2s(Wbuf + (size_t)(s % 3) * WB, P + off, (uint32_t)n * 4u, &bar[s % 3]);
    };
    auto wait_seg = [&](int s) -> const float* {
        tc::mbar_wait(&bar[s % 3], (uint32_t)((s / 3) & 1));
        return Wbuf + (size_t)(s % 3) * WB;
    };
    auto done_seg = [&](int s) {                          // every warp is past segment s: refill its slot
        __syncthreads();
        if (tid == 0) issue(s + 3);
    };

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) tc::mbar_init(&bar[i], 1);
        tc::fence_mbar_init();
    }
    for (int t = tid; t < n_tok; t += blockDim.x) {
        int sl = t;
        if (t >= n_c) { const int si = __ldg(tgt_slot + (t - n_c)); sl = si >= 0 ? n_c + si : -1; }
        slot_s[t] = sl;
    }
    __syncthreads();
    if (tid == 0) { issue(0); issue(1); issue(2); }
    // Everything above reads only the weights and the target map.  From here on the kernel touches what the preceding
    // kernel of the stream wrote (logits, alive) and buffers it may still be reading (K / V): wait for it.
    pdl_wait();
    if (do_select) {
        // fused design step: choose the previous step's design from its logits and append it as context point n_c - 1
        select_block(sel, b);
        __syncthreads();
    }
    // rows to process: in rollout mode a target the candidates do not attend to feeds nothing downstream and is dropped
    for (int t = tid; t < n_c; t += blockDim.x) orig_s[t] = t;
    if (tid == 0) {
        int n = n_c;
        for (int t = n_c; t < n_tok; ++t)
            if (!rollout_mode || slot_s[t] >= 0) orig_s[n++] = t;
        n_eff_s = n;
    }
    __syncthreads();
    const int n_eff = n_eff_s;

    // this warp's rows (kept for the whole kernel; the launcher guarantees NW * NTK >= n_eff)
    const int base = warp * NTK;
    const bool has = base < n_eff;
    const float* xrow[NTK];
    const float* trow[NTK];
    const float* hrow[NTK];
#pragma unroll
    for (int i = 0; i < NTK; ++i) {
        const int rc = base + i < n_eff ? base + i : n_eff - 1;         // padding rows recompute the last row, never store
        xrow[i] = X + (size_t)rc * D;
        trow[i] = T + (size_t)rc * D;
        hrow[i] = hs + i * 128;
    }

    // ---- embedding (model/embedder.py:128-214): X[row] = MLPx(x) (+ MLPy(y) for context points) | theta token ----
    {
        float e[NTK][2];
        const float* W = wait_seg(0);
        if (has) {
            float xin[NTK][8];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = orig_s[base + i < n_eff ? base + i : n_eff - 1];
                const int ti = tok - n_c;
                const bool is_data = tok < n_c || ti < n_td;
#pragma unroll
                for (int k = 0; k < 8; ++k) xin[i][k] = 0.f;
                if (is_data) {
                    const float* src = tok < n_c ? cx + ((size_t)b * ctx_cap + tok) * m.dx : target_x + ((size_t)b * n_td + ti) * m.dx;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < m.dx) xin[i][k] = src[k];
                }
                e[i][0] = 0.f; e[i][1] = 0.f;
            }
            warp_embed_hidden<NTK>(xin, m.dx, W, W + (L.x_b1 - L.x_w1), hs, lane);
            warp_matvec64<NTK, 64>(e, hrow, W + (L.x_w2 - L.x_w1), lane);
        }
        done_seg(0);
        W = wait_seg(1);
        if (has) {
            const float* h2[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) h2[i] = hrow[i] + 64;
            warp_matvec64<NTK, 64>(e, h2, W, lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int row = base + i;
                if (row < n_eff) {
                    const int tok = orig_s[row], ti = tok - n_c;
                    const bool is_data = tok < n_c || ti < n_td;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int f = lane + 32 * j;
                        X[(size_t)row * D + f] = is_data ? e[i][j] + W[kC6Mat + f]
                                                         : __ldg(P + L.tok + (size_t)(ti - n_td) * D + f);
                    }
                }
            }
        }
        done_seg(1);
        const bool has_ctx = base < n_c;                  // context tokens come first
        W = wait_seg(2);
        if (has_ctx) {
            float yin[NTK][8];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i < n_c ? base + i : n_c - 1;
#pragma unroll
                for (int k = 0; k < 8; ++k) yin[i][k] = 0.f;
                yin[i][0] = cy[(size_t)b * ctx_cap + tok];
                e[i][0] = 0.f; e[i][1] = 0.f;
            }
            warp_embed_hidden<NTK>(yin, 1, W, W + (L.y_b1 - L.y_w1), hs, lane);
            warp_matvec64<NTK, 64>(e, hrow, W + (L.y_w2 - L.y_w1), lane);
        }
        done_seg(2);
        W = wait_seg(3);
        if (has_ctx) {
            const float* h2[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) h2[i] = hrow[i] + 64;
            warp_matvec64<NTK, 64>(e, h2, W, lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i;
                if (tok < n_c) {
                    X[(size_t)tok * D + lane] += e[i][0] + W[kC6Mat + lane];
                    X[(size_t)tok * D + lane + 32] += e[i][1] + W[kC6Mat + lane + 32];
                }
            }
        }
        done_seg(3);
    }

    // ---- encoder layers ----
    for (int l = 0; l < m.NL; ++l) {
        const bool last = l + 1 == m.NL;
        const int s0 = 4 + 8 * l;
        const float* Pl = P + L.layer0 + (size_t)l * L.layer_stride;
        const int nkp = (n_keys_tc + 15) / 16 * 16, kbytes = (G + 1) * 16 * nkp, blk_bytes = kbytes + G * 32 * nkp;
        unsigned char* blk = tckv ? tckv + ((size_t)l * B + b) * blk_bytes : nullptr;
        if (blk) {                                        // clear this (layer, rollout) operand block, set the key mask
            for (int i = tid * 16; i < blk_bytes; i += blockDim.x * 16) {
                uint4 z = make_uint4(0, 0, 0, 0);
                const int mrow = (i - G * 16 * nkp) >> 4;  // row of the mask chunk (chunk G of the K part)
                if (i >= G * 16 * nkp && i < kbytes && mrow >= n_keys_tc) z.x = 0xC348u;       // bf16(-200) in element 0
                *reinterpret_cast<uint4*>(blk + i) = z;
            }
        }
        // phase A: q (scaled) -> T, k / v -> shared slots + global; one matrix per segment
        {
            float bq[2], bk[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) { bq[j] = __ldg(Pl + L.bq + lane + 32 * j); bk[j] = __ldg(Pl + L.bk + lane + 32 * j); }
            int sl[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) sl[i] = base + i < n_eff ? slot_s[orig_s[base + i]] : -1;
            const float* W = wait_seg(s0);
            if (has) {
                float a[NTK][2];
#pragma unroll
                for (int i = 0; i < NTK; ++i) { a[i][0] = 0.f; a[i][1] = 0.f; }
                warp_matvec64<NTK, 64>(a, xrow, W, lane);
#pragma unroll
                for (int i = 0; i < NTK; ++i) {
                    if (base + i < n_eff) {
                        float* t = T + (size_t)(base + i) * D;
                        t[lane] = (a[i][0] + bq[0]) * 0.35355339059327376220f;
                        t[lane + 32] = (a[i][1] + bq[1]) * 0.35355339059327376220f;
                    }
                }
            }
            done_seg(s0);
            W = wait_seg(s0 + 1);
            if (has) {
                float a[NTK][2];
#pragma unroll
                for (int i = 0; i < NTK; ++i) { a[i][0] = 0.f; a[i][1] = 0.f; }
                warp_matvec64<NTK, 64>(a, xrow, W, lane);
#pragma unroll
                for (int i = 0; i < NTK; ++i) {
                    if (sl[i] >= 0) {
                        float* gk = kv + (((size_t)l * B + b) * kv_slots + sl[i]) * (2 * D);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float v = a[i][j] + bk[j];
                            Ks[sl[i] * KS + lane + 32 * j] = v;
                            gk[lane + 32 * j] = v;
                        }
                    }
                }
            }
            done_seg(s0 + 1);
            W = wait_seg(s0 + 2);
            if (has) {
                float a[NTK][2];
#pragma unroll
                for (int i = 0; i < NTK; ++i) { a[i][0] = 0.f; a[i][1] = 0.f; }
                warp_matvec64<NTK, 64>(a, xrow, W, lane);
                const float bv0 = W[(L.bv - L.wv) + lane], bv1 = W[(L.bv - L.wv) + lane + 32];
#pragma unroll
                for (int i = 0; i < NTK; ++i) {
                    if (sl[i] >= 0) {
                        float* gk = kv + (((size_t)l * B + b) * kv_slots + sl[i]) * (2 * D) + D;
                        const float v0 = a[i][0] + bv0, v1 = a[i][1] + bv1;
                        Vs[sl[i] * KS + lane] = v0; Vs[sl[i] * KS + lane + 32] = v1;
                        gk[lane] = v0; gk[lane + 32] = v1;
                    }
                }
            }
            done_seg(s0 + 2);                              // also publishes Ks / Vs to the block
        }
        if (blk) {
            // bf16 operands of the tensor-core query stream (csrc/query_tc5.cu).  K part: chunk h (= head) row `slot` =
            // K[slot] - K[0] (the softmax is evaluated relative to key 0); V part: head h, 16-row chunks of 8 keys:
            // rows 0..7 = features, row 8 = 1 (returns the softmax denominator), rows 9..15 = 0
            for (int i = tid; i < n_slots * G; i += blockDim.x) {
                const int sl = i >> 3, h = i & 7;
                bool used = sl < n_c;
                if (!used) {
                    for (int t = n_c; t < n_tok; ++t) used |= slot_s[t] == sl;
                }
                if (!used) continue;
                const float* kr = Ks + sl * KS + 8 * h, *k0 = Ks + 8 * h, *vr = Vs + sl * KS + 8 * h;
                uint4 q4;
                q4.x = tc::pack_bf16(kr[0] - k0[0], kr[1] - k0[1]); q4.y = tc::pack_bf16(kr[2] - k0[2], kr[3] - k0[3]);
                q4.z = tc::pack_bf16(kr[4] - k0[4], kr[5] - k0[5]); q4.w = tc::pack_bf16(kr[6] - k0[6], kr[7] - k0[7]);
                *reinterpret_cast<uint4*>(blk + ((size_t)h * nkp + sl) * 16) = q4;
                __nv_bfloat16* vb = reinterpret_cast<__nv_bfloat16*>(blk + kbytes) + ((size_t)h * (nkp / 8) + (sl >> 3)) * 128 + (sl & 7);
#pragma unroll
                for (int f = 0; f < 8; ++f) vb[f * 8] = __float2bfloat16_rn(vr[f]);
                vb[64] = __float2bfloat16_rn(1.0f);
            }
        }
        if (last && rollout_mode) break;

        // last layer: only the targets continue (z_tgt), unless the value head wants the context rows too
        const bool run = has && !(last && !ctx_last && base + NTK <= n_c);
        // phase B: attention over the context keys (two passes of four heads), out-projection + residual, LayerNorm 1 -> T
        const float* W = wait_seg(s0 + 3);
        if (run) {
            float o[NTK][2];
#pragma unroll
            for (int p = 0; p < NTK; p += NP) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float* qr[NP];
                    float op[NP];
#pragma unroll
                    for (int i = 0; i < NP; ++i) qr[i] = trow[p + i] + 32 * j;
                    warp_attention<NP, KS>(op, qr, Ks + 32 * j, Vs + 32 * j, n_c, lane);
#pragma unroll
                    for (int i = 0; i < NP; ++i) o[p + i][j] = op[i];
                }
            }
            __syncwarp();                                  // every lane has read the query rows
            float hres[NTK][2];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                if (base + i < n_eff) {
                    float* t = T + (size_t)(base + i) * D;
                    t[lane] = o[i][0]; t[lane + 32] = o[i][1];
                }
                hres[i][0] = W[(L.bo - L.wo) + lane] + xrow[i][lane];
                hres[i][1] = W[(L.bo - L.wo) + lane + 32] + xrow[i][lane + 32];
            }
            __syncwarp();
            warp_matvec64<NTK, 64>(hres, trow, W, lane);
            warp_layer_norm64<NTK>(hres, W + (L.g1 - L.wo), W + (L.be1 - L.wo), lane);
            __syncwarp();                                  // every lane has read the attention-output rows
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                if (base + i < n_eff) {
                    float* t = T + (size_t)(base + i) * D;
                    t[lane] = hres[i][0]; t[lane + 32] = hres[i][1];
                }
            }
            __syncwarp();
        }
        done_seg(s0 + 3);

        // phase C: x' = LayerNorm2(h + W2 relu(W1 h + b1) + b2) -> X; W1 and W2 arrive as two halves each
        float hid[NTK][4];
#pragma unroll
        for (int i = 0; i < NTK; ++i) { hid[i][0] = 0.f; hid[i][1] = 0.f; hid[i][2] = 0.f; hid[i][3] = 0.f; }
        W = wait_seg(s0 + 4);
        if (run) warp_hidden64<NTK, 32>(hid, trow, W, lane);
        done_seg(s0 + 4);
        W = wait_seg(s0 + 5);
        if (run) {
            const float* t2[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) t2[i] = trow[i] + 32;
            warp_hidden64<NTK, 32>(hid, t2, W, lane);
            const float4 bb = *reinterpret_cast<const float4*>(W + kC6Mat + 4 * lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i)
                *reinterpret_cast<float4*>(hs + i * 128 + 4 * lane) =
                    make_float4(fmaxf(hid[i][0] + bb.x, 0.f), fmaxf(hid[i][1] + bb.y, 0.f), fmaxf(hid[i][2] + bb.z, 0.f),
                                fmaxf(hid[i][3] + bb.w, 0.f));
            __syncwarp();
        }
        done_seg(s0 + 5);
        float acc[NTK][2];
        W = wait_seg(s0 + 6);
        if (run) {
#pragma unroll
            for (int i = 0; i < NTK; ++i) { acc[i][0] = trow[i][lane]; acc[i][1] = trow[i][lane + 32]; }
            warp_matvec64<NTK, 64>(acc, hrow, W, lane);
        }
        done_seg(s0 + 6);
        W = wait_seg(s0 + 7);
        if (run) {
            const float* h2[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) h2[i] = hrow[i] + 64;
            warp_matvec64<NTK, 64>(acc, h2, W, lane);
            const float* vb = W + kC6Mat;                  // b2 | g2 | be2
#pragma unroll
            for (int i = 0; i < NTK; ++i) { acc[i][0] += vb[lane]; acc[i][1] += vb[lane + 32]; }
            warp_layer_norm64<NTK>(acc, vb + D, vb + 2 * D, lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                if (base + i < n_eff) {
                    float* xr = X + (size_t)(base + i) * D;
                    xr[lane] = acc[i][0]; xr[lane + 32] = acc[i][1];
                }
            }
            __syncwarp();
        }
        done_seg(s0 + 7);
    }
    if (z_tgt || z_ctx) __syncthreads();
    if (z_tgt)
        for (int i = tid; i < n_t * D; i += blockDim.x) z_tgt[(size_t)b * n_t * D + i] = X[(size_t)n_c * D + i];
    if (z_ctx)
        for (int i = tid; i < n_c * D; i += blockDim.x) z_ctx[(size_t)b * n_c * D + i] = X[i];
}

struct C6Plan { int ntk, warps, n_slots; size_t smem; int wb; };

static bool c6_plan(const Dims& d, const Layout& L, int n_c, int n_tok, int kv_slots, C6Plan& p, int min_warps = 1,
                    int n_rows = 0, int B = 1) {
    if (d.D != kC6D || d.H != 8 || d.FF != 128 || d.EH != 128 || d.dx > 8 || n_c > 64 || n_c < 1) return false;
    if (n_tok > kC6MaxWarps * 4) return false;            // a warp keeps its (<= 4) tokens for the whole kernel
    if (n_rows < 1 || n_rows > n_tok) n_rows = n_tok;      // rows actually processed (rollout mode drops dead targets)
    p.ntk = n_rows <= 6 ? 1 : n_rows <= 22 ? 2 : 4;       // tokens per warp: thresholds measured for d = 32 (ctx_warp.cu)
    static const int force_ntk = [] { const char* e = getenv("ALINE_CTX_NTK"); return e ? atoi(e) : 0; }();
    if (B >= 3 * device_info().sm_count && n_rows >= 4) p.ntk = n_rows <= 24 ? 2 : 4;     // throughput regime (ctx_warp.cu)
    if ((force_ntk == 1 || force_ntk == 2 || force_ntk == 4) && force_ntk * kC6MaxWarps >= n_rows) p.ntk = force_ntk;
    p.warps = (n_rows + p.ntk - 1) / p.ntk;
    if (p.warps < min_warps) p.warps = min_warps;        // the fused select wants a few warps over the candidates
    if (p.warps > kC6MaxWarps) return false;
    p.n_slots = kv_slots < n_tok ? kv_slots : n_tok;
    size_t wb = 0;
    for (int s = 0; s < 4 + 8; ++s) {                     // every layer has the same segment sizes
        size_t off, n;
        c6_segment(L, s, off, n);
        if (n % 4 != 0 || off % 4 != 0) return false;     // 16-byte bulk copies
        if (n > wb) wb = n;
    }
    p.wb = (int)((wb + 31) & ~(size_t)31);
    size_t fl = 3 * (size_t)p.wb + 2 * (size_t)n_tok * kC6D + (size_t)p.warps * p.ntk * 128 + 2 * (size_t)p.n_slots * kC6KS + 2 * n_tok;
    p.smem = fl * sizeof(float) + 16;
    return p.smem <= (size_t)device_info().max_smem_optin;
}

bool ctx_stack_warp64_supported(const Dims& d, const Layout& L, const float* P, int n_c, int n_tok, int kv_slots) {
    C6Plan p;
    return ((uintptr_t)P % 16 == 0) && c6_plan(d, L, n_c, n_tok, kv_slots, p, 8);
}

int ctx_stack_warp64(const Dims& d, const Layout& L, const float* P, const float* cx, const float* cy, int B, int n_c,
                     int ctx_cap, const float* target_x, int n_td, const int* tgt_slot, float* kv, int kv_slots,
                     float* z_tgt, float* z_ctx, void* tckv, int n_keys_tc, const SelectArgs* sel, int n_rows_hint,
                     cudaStream_t st) {
    const int n_tok = n_c + n_td + d.ntok;
    C6Plan p;
    ALINE_REQUIRE(c6_plan(d, L, n_c, n_tok, kv_slots, p, sel ? 8 : 1, (z_tgt || z_ctx) ? 0 : n_rows_hint, B),
                  "ctx_stack_warp64: unsupported shape");
    const SelectArgs sa = sel ? *sel : SelectArgs{};
#define ALINE_C6_LAUNCH(NTKV)                                                                                          \
    do {                                                                                                               \
        if (ensure_dyn_smem((const void*)ctx_stack_warp64_kernel<NTKV>, p.smem)) return 1;                            \
        ALINE_CHECK_CUDA(launch_k(ctx_stack_warp64_kernel<NTKV>, dim3(B), dim3(32 * p.warps), p.smem, st, g_pdl_chain, \
                                  d, L, P, cx, cy, n_c, ctx_cap, target_x, n_td, tgt_slot, kv, kv_slots, B, z_tgt,    \
                                  z_ctx, p.wb, p.n_slots, (unsigned char*)tckv, n_keys_tc, sa, (int)(sel != nullptr)));  \
    } while (0)
    if (p.ntk == 1) ALINE_C6_LAUNCH(1);
    else if (p.ntk == 2) ALINE_C6_LAUNCH(2);
    else ALINE_C6_LAUNCH(4);
#undef ALINE_C6_LAUNCH
    ALINE_LAUNCH_OK();
    return 0;
}

}  // namespace aline
