// Context + target stack, warp-per-token fp32 kernel for d = 32 (the shape of every shipped config).
//
// Same contract as ctx_stack_kernel (csrc/rollout.cu; reference: model/embedder.py:128-214 + model/encoder.py:128-141
// restricted to the context / target rows): one thread block per rollout runs the few context + target tokens through
// the embedders and all encoder layers and emits, per layer, the keys / values the candidate-query stream needs.
//
// Mapping: a warp owns NTK tokens, lane f owns feature f of each (d = 32 = one warp), so every weight matrix read
// W[k][f] is one conflict-free 128-byte shared-memory wavefront that feeds NTK FMAs; the per-token input vector is
// read as LDS.128 broadcasts.  The 128 hidden units of the MLPs are split 4 per lane (W1 as LDS.128), exchanged through
// a warp-private shared row and contracted with W2 lane-per-feature again.  Attention: lane = (head, key mod 8), the
// probabilities are handed to the feature lanes with shuffles.  LayerNorm statistics are warp reductions.
// Ten times more warps in flight than the lane-per-head kernel at the same instruction count: the step is latency
// bound (a few thousand dependent instructions), so parallelism across warps is what shortens it.
//
// Weights are streamed through a three-slot shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier) in
// segments <= 20 KB -- x-embedder, y-embedder, then per layer {Wq Wk Wv Wo + LN1}, {W1}, {W2 + LN2} -- each issued as
// soon as the segment that used its slot has been consumed, so the copy of segment s+2 overlaps the math of s, s+1.
#include "model.cuh"
#include "tc.cuh"
#include "select.cuh"
#include "ctx_warp.cuh"
#include "query_fast.cuh"
#include <cstdlib>

namespace aline {

constexpr int kCwD = 32;
constexpr int kCwKS = 36;          // padded K / V row stride in floats: 8 rows x 16 bytes hit 32 distinct banks
constexpr int kCwMaxWarps = 16;

__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LayerNorm of NTK feature-per-lane vectors (biased variance, eps 1e-5)
template <int NTK>
__device__ __forceinline__ void warp_layer_norm(float (&v)[NTK], float g, float be) {
    float mu[NTK], q[NTK];
#pragma unroll
    for (int i = 0; i < NTK; ++i) mu[i] = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NTK; ++i) mu[i] += __shfl_xor_sync(0xffffffffu, mu[i], o);
    }
#pragma unroll
    for (int i = 0; i < NTK; ++i) { v[i] -= mu[i] * (1.0f / 32); q[i] = v[i] * v[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NTK; ++i) q[i] += __shfl_xor_sync(0xffffffffu, q[i], o);
    }
#pragma unroll
    for (int i = 0; i < NTK; ++i) v[i] = v[i] * (1.0f / sqrtf(q[i] * (1.0f / 32) + 1e-5f)) * g + be;
}

// acc[i] += sum_{k < 32} row_i[k] * W[k * ldw + lane]       (row_i: shared, 16-byte aligned, read as broadcasts)
template <int NTK>
__device__ __forceinline__ void warp_matvec32(float (&acc)[NTK], const float* const (&row)[NTK], const float* W, int ldw,
                                              int lane) {
    // two partial sums (even / odd k) per token as the halves of packed fp32x2 operands: half the FMA issue slots and
    // half the length of the dependent chain
    f32x2 a2[NTK];
#pragma unroll
    for (int i = 0; i < NTK; ++i) a2[i] = pk2(acc[i], 0.f);
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        float4 xv[NTK];
#pragma unroll
        for (int i = 0; i < NTK; ++i) xv[i] = *reinterpret_cast<const float4*>(row[i] + 4 * k4);
        const f32x2 w01 = pk2(W[(4 * k4 + 0) * ldw + lane], W[(4 * k4 + 1) * ldw + lane]);
        const f32x2 w23 = pk2(W[(4 * k4 + 2) * ldw + lane], W[(4 * k4 + 3) * ldw + lane]);
#pragma unroll
        for (int i = 0; i < NTK; ++i) {
            a2[i] = fma2(pk2(xv[i].x, xv[i].y), w01, a2[i]);
            a2[i] = fma2(pk2(xv[i].z, xv[i].w), w23, a2[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NTK; ++i) { float lo, hi; upk2(a2[i], lo, hi); acc[i] = lo + hi; }
}

// 2-layer MLP  in(IN) -> HID (ReLU) -> 32 for NTK tokens of one warp; HID in chunks of 128 (4 hidden units per lane).
//   IN_SMEM: the inputs are 32-float shared rows (xin[i] = row pointer), else registers xr[i][0..IN)
//   W1 [IN][HID], b1 [HID] from w1s; W2 [HID][32] from w2s (may live in different ring slots)
// acc[i] += W2^T relu(W1^T x_i + b1)
template <int NTK, bool IN_SMEM>
__device__ __forceinline__ void warp_mlp(float (&acc)[NTK], const float* const (&xrow)[NTK], const float (&xr)[NTK][8],
                                         int IN, const float* W1, const float* b1, const float* W2, int HID, float* hs,
                                         int lane) {
    for (int c0 = 0; c0 < HID; c0 += 128) {
        float hid[NTK][4];
        {
            const float4 bb = *reinterpret_cast<const float4*>(b1 + c0 + 4 * lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i) { hid[i][0] = bb.x; hid[i][1] = bb.y; hid[i][2] = bb.z; hid[i][3] = bb.w; }
        }
        if constexpr (IN_SMEM) {
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                float4 xv[NTK];
#pragma unroll
                for (int i = 0; i < NTK; ++i) xv[i] = *reinterpret_cast<const float4*>(xrow[i] + 4 * k4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 w = *reinterpret_cast<const float4*>(W1 + (size_t)(4 * k4 + kk) * HID + c0 + 4 * lane);
#pragma unroll
                    for (int i = 0; i < NTK; ++i) {
                        const float xk = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
                        const f32x2 xx = pk2(xk, xk);                 // hidden units in pairs: packed fp32x2 FMAs
                        f32x2 h01 = pk2(hid[i][0], hid[i][1]), h23 = pk2(hid[i][2], hid[i][3]);
                        h01 = fma2(xx, pk2(w.x, w.y), h01); h23 = fma2(xx, pk2(w.z, w.w), h23);
                        upk2(h01, hid[i][0], hid[i][1]); upk2(h23, hid[i][2], hid[i][3]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k < IN) {
                    const float4 w = *reinterpret_cast<const float4*>(W1 + (size_t)k * HID + c0 + 4 * lane);
#pragma unroll
                    for (int i = 0; i < NTK; ++i) {
                        hid[i][0] = fmaf(xr[i][k], w.x, hid[i][0]); hid[i][1] = fmaf(xr[i][k], w.y, hid[i][1]);
                        hid[i][2] = fmaf(xr[i][k], w.z, hid[i][2]); hid[i][3] = fmaf(xr[i][k], w.w, hid[i][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NTK; ++i)
            *reinterpret_cast<float4*>(hs + i * 128 + 4 * lane) =
                make_float4(fmaxf(hid[i][0], 0.f), fmaxf(hid[i][1], 0.f), fmaxf(hid[i][2], 0.f), fmaxf(hid[i][3], 0.f));
        __syncwarp();
        const float* w2 = W2 + (size_t)c0 * kCwD + lane;
        f32x2 a2[NTK];                                              // even / odd hidden units: two packed partial sums
#pragma unroll
        for (int i = 0; i < NTK; ++i) a2[i] = pk2(acc[i], 0.f);
#pragma unroll 8
        for (int c4 = 0; c4 < 32; ++c4) {
            float4 hv[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) hv[i] = *reinterpret_cast<const float4*>(hs + i * 128 + 4 * c4);
            const f32x2 w01 = pk2(w2[(4 * c4 + 0) * kCwD], w2[(4 * c4 + 1) * kCwD]);
            const f32x2 w23 = pk2(w2[(4 * c4 + 2) * kCwD], w2[(4 * c4 + 3) * kCwD]);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                a2[i] = fma2(pk2(hv[i].x, hv[i].y), w01, a2[i]);
                a2[i] = fma2(pk2(hv[i].z, hv[i].w), w23, a2[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < NTK; ++i) { float lo, hi; upk2(a2[i], lo, hi); acc[i] = lo + hi; }
        __syncwarp();
    }
}

template <int NTK>
__global__ void __launch_bounds__(32 * kCwMaxWarps, 2)
ctx_stack_warp_kernel(const Dims m, const Layout L, const float* __restrict__ P, const float* cx,
                      const float* cy, int n_c, int ctx_cap, const float* __restrict__ target_x, int n_td,
                      const int* __restrict__ tgt_slot, float* __restrict__ kv, int kv_slots, int B,
                      float* __restrict__ z_tgt, float* __restrict__ z_ctx, int WB, int n_slots,
                      unsigned char* __restrict__ tckv, int n_keys_tc, const SelectArgs sel, int do_select, int emit_fold, int nkf) {
    constexpr int D = kCwD;
    constexpr int NP = NTK >= 2 ? 2 : 1;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int b = blockIdx.x;
    const int n_t = n_td + m.ntok, n_tok = n_c + n_t;
    float* Wbuf = smem;                                   // 3 ring slots of WB floats
    float* X = Wbuf + 3 * (size_t)WB;                     // [n_tok][32] layer input / output
    float* T = X + (size_t)n_tok * D;                     // [n_tok][32] scaled query, then attention output, then LN1 output
    float* Hs = T + (size_t)n_tok * D;                    // [NW][NTK][128] hidden units (warp-private)
    float* Ks = Hs + (size_t)NW * NTK * 128;              // [n_slots][36]
    float* Vs = Ks + (size_t)n_slots * kCwKS;             // [n_slots][36]
    int* slot_s = reinterpret_cast<int*>(Vs + (size_t)n_slots * kCwKS);   // [n_tok] key slot of a token (-1: not attended)
    int* orig_s = slot_s + n_tok;                                         // [n_tok] token id of a processed row
    __shared__ int n_eff_s;
    float* hs = Hs + (size_t)warp * NTK * 128;

    pdl_trigger();                                        // the next kernel of the stream may start its own prologue
    const bool rollout_mode = z_tgt == nullptr && z_ctx == nullptr;   // nothing downstream of the last layer's K, V
    const bool ctx_last = z_ctx != nullptr;               // the value head reads the context rows' final encodings
    const int n_seg = 2 + 3 * m.NL - (rollout_mode ? 2 : 0);
    auto issue = [&](int s) {                             // one thread
        if (s >= n_seg) return;
        const float* src;
        int n;
        if (s == 0) { src = P + L.x_w1; n = (int)(L.y_w1 - L.x_w1); }
        else if (s == 1) { src = P + L.y_w1; n = (int)(L.tok - L.y_w1); }
        else {
            const int l = (s - 2) / 3, k = (s - 2) % 3;
            const float* base = P + L.layer0 + (size_t)l * L.layer_stride;
            if (k == 0) { src = base + L.wq; n = (int)(L.w1 - L.wq); }
            else if (k == 1) { src = base + L.w1; n = (int)(L.w2 - L.w1); }
            else { src = base + L.w2; n = (int)(L.layer_stride - L.w2); }
        }
        tc::mbar_arrive_expect_tx(&bar[s % 3], (uint32_t)n * 4u);
        tc::bulk_g2s(Wbuf + (size_t)(s % 3) * WB, src, (uint32_t)n * 4u, &bar[s % 3]);
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
    // kernel of the stream wrote (logits, alive) and buffers it may still be reading (K / V): wait for it (a no-op
    // unless launched as a programmatic dependent, see common.cuh).
    pdl_wait();
    if (do_select) {
        // fused design step: choose the previous step's design from its logits and append it as context point n_c - 1
        // (written by thread 0, read below with ordinary loads after the barrier)
        select_block(sel, b);
        __syncthreads();
    }
    // Rows to process.  In rollout mode a target the candidates do not attend to (slot < 0) feeds nothing downstream --
    // targets only attend to the context -- so it is dropped: with a 'split' target mask attending to the 3 theta
    // tokens, 100 of 103 targets of the GP configuration disappear from the work list.
    for (int t = tid; t < n_c; t += blockDim.x) orig_s[t] = t;
    if (tid == 0) {
        int n = n_c;
        for (int t = n_c; t < n_tok; ++t)
            if (!rollout_mode || slot_s[t] >= 0) orig_s[n++] = t;
        n_eff_s = n;
    }
    __syncthreads();
    const int n_eff = n_eff_s;

    const int per_round = NW * NTK;
    const int rounds = (n_eff + per_round - 1) / per_round;

    // ---- embedding (model/embedder.py:128-214): X[tok] = MLPx(x) (+ MLPy(y) for context points) | theta token ----
    {
        const float* Wx = wait_seg(0);
        for (int rd = 0; rd < rounds; ++rd) {
            const int base = (rd * NW + warp) * NTK;
            if (base >= n_eff) break;
            float xin[NTK][8];
            const float* none[NTK];
            float e[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = orig_s[base + i < n_eff ? base + i : n_eff - 1];      // token id of this row
                const int ti = tok - n_c;
                none[i] = nullptr;
                const bool is_data = tok < n_c || ti < n_td;
#pragma unroll
                for (int k = 0; k < 8; ++k) xin[i][k] = 0.f;
                if (is_data) {
                    const float* src = tok < n_c ? cx + ((size_t)b * ctx_cap + tok) * m.dx : target_x + ((size_t)b * n_td + ti) * m.dx;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < m.dx) xin[i][k] = src[k];
                }
                e[i] = Wx[(L.x_b2 - L.x_w1) + lane];
            }
            warp_mlp<NTK, false>(e, none, xin, m.dx, Wx, Wx + (L.x_b1 - L.x_w1), Wx + (L.x_w2 - L.x_w1), m.EH, hs, lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int row = base + i;
                if (row < n_eff) {
                    const int tok = orig_s[row], ti = tok - n_c;
                    const bool is_data = tok < n_c || ti < n_td;
                    X[(size_t)row * D + lane] = is_data ? e[i] : __ldg(P + L.tok + (size_t)(ti - n_td) * D + lane);
                }
            }
        }
        done_seg(0);
        const float* Wy = wait_seg(1);
        for (int rd = 0; rd < rounds; ++rd) {
            const int base = (rd * NW + warp) * NTK;
            if (base >= n_c) break;                       // context tokens come first: no context point in this warp
            float yin[NTK][8];
            const float* none[NTK];
            float e[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i < n_c ? base + i : n_c - 1;
                none[i] = nullptr;
#pragma unroll
                for (int k = 0; k < 8; ++k) yin[i][k] = 0.f;
                yin[i][0] = cy[(size_t)b * ctx_cap + tok];
                e[i] = Wy[(L.y_b2 - L.y_w1) + lane];
            }
            warp_mlp<NTK, false>(e, none, yin, 1, Wy, Wy + (L.y_b1 - L.y_w1), Wy + (L.y_w2 - L.y_w1), m.EH, hs, lane);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i;
                if (tok < n_c) X[(size_t)tok * D + lane] += e[i];
            }
        }
        done_seg(1);
    }

    // ---- encoder layers ----
    unsigned long long used_mask = 0ull;                 // key slots that hold a key (folded operands)
    for (int l = 0; l < m.NL; ++l) {
        const bool last = l + 1 == m.NL;
        const int sA = 2 + 3 * l;
        const int nkp = (n_keys_tc + 15) / 16 * 16, kbytes = 80 * nkp, blk_bytes = 208 * nkp;
        unsigned char* blk = tckv ? tckv + ((size_t)l * B + b) * blk_bytes : nullptr;
        const bool plain = blk && emit_fold != 2;         // emit_fold == 2: only the folded operands will be read
        if (plain) {                                      // clear this (layer, rollout) operand block, set the key mask
            for (int i = tid * 16; i < blk_bytes; i += blockDim.x * 16) {
                uint4 z = make_uint4(0, 0, 0, 0);
                const int mrow = (i - 64 * nkp) >> 4;      // row of the mask chunk (chunk 4 of the K part)
                if (i >= 64 * nkp && i < kbytes && mrow >= n_keys_tc) z.x = 0xC348u;       // bf16(-200) in element 0
                *reinterpret_cast<uint4*>(blk + i) = z;
            }
        }
        const float* WA = wait_seg(sA);
        const float* Wq = WA, *Wk = WA + (L.wk - L.wq), *Wv = WA + (L.wv - L.wq);
        // phase A: q (scaled) -> T, k / v -> shared slots + global
        for (int rd = 0; rd < rounds; ++rd) {
            const int base = (rd * NW + warp) * NTK;
            if (base >= n_eff) break;
            const float* xrow[NTK];
            float aq[NTK], ak[NTK], av[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i < n_eff ? base + i : n_eff - 1;
                xrow[i] = X + (size_t)tok * D;
                aq[i] = WA[(L.bq - L.wq) + lane]; ak[i] = WA[(L.bk - L.wq) + lane]; av[i] = WA[(L.bv - L.wq) + lane];
            }
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                float4 xv[NTK];
#pragma unroll
                for (int i = 0; i < NTK; ++i) xv[i] = *reinterpret_cast<const float4*>(xrow[i] + 4 * k4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k = 4 * k4 + kk;
                    const float wq = Wq[k * D + lane], wk = Wk[k * D + lane], wv = Wv[k * D + lane];
#pragma unroll
                    for (int i = 0; i < NTK; ++i) {
                        const float xk = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
                        aq[i] = fmaf(xk, wq, aq[i]); ak[i] = fmaf(xk, wk, ak[i]); av[i] = fmaf(xk, wv, av[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i;
                if (tok < n_eff) {
                    T[(size_t)tok * D + lane] = aq[i] * 0.35355339059327376220f;
                    const int sl = slot_s[orig_s[tok]];
                    if (sl >= 0) {
                        Ks[sl * kCwKS + lane] = ak[i];
                        Vs[sl * kCwKS + lane] = av[i];
                        float* gk = kv + (((size_t)l * B + b) * kv_slots + sl) * (2 * D);
                        gk[lane] = ak[i];
                        gk[D + lane] = av[i];
                    }
                }
            }
        }
        __syncthreads();
        if (plain) {
            // bf16 operands of the fast tensor-core query stream (csrc/query_tc3.cu).  K part: chunk h (= head) row
            // `slot` = K[slot] - K[0] (the softmax is evaluated relative to key 0); V part: head h, 16-row chunks of 8
            // keys: rows 0..7 = features, row 8 = 1 (returns the softmax denominator), rows 9..15 = 0
            for (int i = tid; i < n_slots * 4; i += blockDim.x) {
                const int sl = i >> 2, h = i & 3;
                bool used = sl < n_c;
                if (!used) {
                    for (int t = n_c; t < n_tok; ++t) used |= slot_s[t] == sl;
                }
                if (!used) continue;
                const float* kr = Ks + sl * kCwKS + 8 * h, *k0 = Ks + 8 * h, *vr = Vs + sl * kCwKS + 8 * h;
                uint4 q4;
                q4.x = tc::pack_bf16(kr[0] - k0[0], kr[1] - k0[1]); q4.y = tc::pack_bf16(kr[2] - k0[2], kr[3] - k0[3]);
                q4.z = tc::pack_bf16(kr[4] - k0[4], kr[5] - k0[5]); q4.w = tc::pack_bf16(kr[6] - k0[6], kr[7] - k0[7]);
                *reinterpret_cast<uint4*>(blk + ((size_t)h * nkp + sl) * 16) = q4;
                __nv_bfloat16* vb = reinterpret_cast<__nv_bfloat16*>(blk + 80 * nkp) + ((size_t)h * (nkp / 8) + (sl >> 3)) * 128 + (sl & 7);
#pragma unroll
                for (int f = 0; f < 8; ++f) vb[f * 8] = __float2bfloat16_rn(vr[f]);
                vb[64] = __float2bfloat16_rn(1.0f);
            }
        }
        if (blk && emit_fold) {
            // folded operands of the candidate stream (query_fast.cuh: K'_h = c Wq_h^T (K_h - K_0h), V'_h = V_h Wo_h^T), from
            // the fp32 K / V rows and the layer's weights in shared memory.  Items: K' (chunk c, head h, slot) -> one
            // 16-byte row unit (+ the bias / mask and zero chunks with c = 0); V' (head h, 8 slots, feature o) -> one
            // 16-byte unit; slot / feature fastest, so a warp shares its weight (K') or value (V') reads.
            if (l == 0) {
                used_mask = n_c >= 64 ? ~0ull : ((1ull << n_c) - 1ull);
                for (int t = n_c; t < n_tok; ++t) {
                    const int sl = slot_s[t];
                    if (sl >= 0 && sl < 64) used_mask |= 1ull << sl;
                }
            }
            // nkf: the key count padded to 8 (16 for the two-warpgroup form) -- rows / columns of the folded blocks
            unsigned char* kp = tckv + tcq::tc2_fold_offset(m.NL, B, nkp) + ((size_t)l * B + b) * ((size_t)tcq::kFoldKeyBytes * nkf);
            const float* Wo_ = WA + (L.wo - L.wq), *bq_ = WA + (L.bq - L.wq);
            const uint32_t chunk = (uint32_t)(4 * nkf) * 16u;
            const int nK = 16 * nkf, nV = 16 * nkf;                    // 4 chunks x 4 heads x nkf ; 4 heads x nkf / 8 x 32
            // index arithmetic by shifts for the power-of-two paddings; the used slots as a bit mask limited to the slots
            // that exist
            const bool p2 = (nkf & (nkf - 1)) == 0;
            const int sh = 31 - __clz(nkf);
            const unsigned long long um = n_slots >= 64 ? used_mask : (used_mask & ((1ull << n_slots) - 1ull));
#ifndef ALINE_FOLD_SKIP_EMIT                               // development: timing without the operand fold (results invalid)
            for (int it = tid; it < nK + nV; it += blockDim.x) {
                if (it < nK) {
                    const int sl = p2 ? (it & (nkf - 1)) : it % nkf;
                    const int hc = p2 ? (it >> sh) : it / nkf, h = hc & 3, c = hc >> 2;
                    const bool used = (um >> sl) & 1ull;
                    float kd[8];
                    {
                        const float* kr = Ks + (used ? sl : 0) * kCwKS + 8 * h, *k0 = Ks + 8 * h;
                        const float4 a0 = *reinterpret_cast<const float4*>(kr), a1 = *reinterpret_cast<const float4*>(kr + 4);
                        const float4 z0 = *reinterpret_cast<const float4*>(k0), z1 = *reinterpret_cast<const float4*>(k0 + 4);
                        kd[0] = a0.x - z0.x; kd[1] = a0.y - z0.y; kd[2] = a0.z - z0.z; kd[3] = a0.w - z0.w;
                        kd[4] = a1.x - z1.x; kd[5] = a1.y - z1.y; kd[6] = a1.z - z1.z; kd[7] = a1.w - z1.w;
                    }
                    float o[8];
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) {
                        const float4* wr = reinterpret_cast<const float4*>(Wq + (8 * c + ii) * D + 8 * h);
                        const float4 w0 = wr[0], w1 = wr[1];
                        float a = w0.x * kd[0];
                        a = fmaf(w0.y, kd[1], a); a = fmaf(w0.z, kd[2], a); a = fmaf(w0.w, kd[3], a);
                        a = fmaf(w1.x, kd[4], a); a = fmaf(w1.y, kd[5], a); a = fmaf(w1.z, kd[6], a); a = fmaf(w1.w, kd[7], a);
                        o[ii] = used ? a * tcq::kFoldScale : 0.f;
                    }
                    uint4 q;
                    q.x = tcq::pack2(o[0], o[1]); q.y = tcq::pack2(o[2], o[3]); q.z = tcq::pack2(o[4], o[5]); q.w = tcq::pack2(o[6], o[7]);
                    const int n = h * nkf + sl;
                    *reinterpret_cast<uint4*>(kp + c * chunk + (size_t)n * 16) = q;
                    if (c == 0) {
                        float bias = 0.f;
#pragma unroll
                        for (int e = 0; e < 8; ++e) bias = fmaf(bq_[8 * h + e], kd[e], bias);
                        bias *= tcq::kFoldScale;
                        const float hi = __bfloat162float(__float2bfloat16_rn(bias));
                        *reinterpret_cast<uint4*>(kp + 4 * chunk + (size_t)n * 16) =
                            make_uint4(used ? tcq::pack2(hi, bias - hi) : 0xC348u, 0u, 0u, 0u);       // bf16(-200): p = 0
                        *reinterpret_cast<uint4*>(kp + 5 * chunk + (size_t)n * 16) = make_uint4(0u, 0u, 0u, 0u);
                    }
                } else {
                    const int j = it - nK, o = j & 31;
                    const int h = p2 ? (j >> (sh + 2)) : j / (4 * nkf), kg = (j >> 5) - h * (nkf >> 3);
                    const unsigned um8 = (unsigned)(um >> (8 * kg)) & 0xffu;
                    float w[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) w[e] = Wo_[(8 * h + e) * D + o];
                    float a[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int sl = 8 * kg + k;
                        const bool used = (um8 >> k) & 1u;
                        const float* vr = Vs + (used ? sl : 0) * kCwKS + 8 * h;
                        const float4 v0 = *reinterpret_cast<const float4*>(vr), v1 = *reinterpret_cast<const float4*>(vr + 4);
                        float acc = v0.x * w[0];
                        acc = fmaf(v0.y, w[1], acc); acc = fmaf(v0.z, w[2], acc); acc = fmaf(v0.w, w[3], acc);
                        acc = fmaf(v1.x, w[4], acc); acc = fmaf(v1.y, w[5], acc); acc = fmaf(v1.z, w[6], acc); acc = fmaf(v1.w, w[7], acc);
                        a[k] = used ? acc : 0.f;
                    }
                    uint4 q;
                    q.x = tcq::pack2(a[0], a[1]); q.y = tcq::pack2(a[2], a[3]); q.z = tcq::pack2(a[4], a[5]); q.w = tcq::pack2(a[6], a[7]);
                    *reinterpret_cast<uint4*>(kp + 384 * nkf + ((size_t)(h * (nkf / 8) + kg) * 32 + o) * 16) = q;
                }
            }
#else
            for (int it = tid; it < 40 * nkf; it += blockDim.x) reinterpret_cast<uint4*>(kp)[it] = make_uint4(0u, 0u, 0u, 0u);
#endif
        }
        if (last && rollout_mode) break;

        // phase B: attention over the context keys, out-projection + residual, LayerNorm 1 -> T (and registers)
        const float* Wo = WA + (L.wo - L.wq);
        for (int rd = 0; rd < rounds; ++rd) {
            const int base = (rd * NW + warp) * NTK;
            if (base >= n_eff) break;
            if (last && !ctx_last && base + NTK <= n_c) continue;     // last layer: only the targets continue (z_tgt)
            const float* trow[NTK];
            float hres[NTK];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i < n_eff ? base + i : n_eff - 1;
                trow[i] = T + (size_t)tok * D;
            }
            float o[NTK];
#pragma unroll
            for (int p = 0; p < NTK; p += NP) {
                const float* qr[NP];
                float op[NP];
#pragma unroll
                for (int i = 0; i < NP; ++i) qr[i] = trow[p + i];
                warp_attention<NP, kCwKS>(op, qr, Ks, Vs, n_c, lane);
#pragma unroll
                for (int i = 0; i < NP; ++i) o[p + i] = op[i];
            }
            __syncwarp();                                  // every lane has read the query rows
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i;
                if (tok < n_eff) T[(size_t)tok * D + lane] = o[i];
                const int tc_ = tok < n_eff ? tok : n_eff - 1;
                hres[i] = WA[(L.bo - L.wq) + lane] + X[(size_t)tc_ * D + lane];
            }
            __syncwarp();
            warp_matvec32<NTK>(hres, trow, Wo, D, lane);
            warp_layer_norm<NTK>(hres, WA[(L.g1 - L.wq) + lane], WA[(L.be1 - L.wq) + lane]);
            __syncwarp();                                  // every lane has read the attention-output rows
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i;
                if (tok < n_eff) T[(size_t)tok * D + lane] = hres[i];
            }
        }
        done_seg(sA);

        // phase C: x' = LayerNorm2(h + W2 relu(W1 h + b1) + b2) -> X
        const float* WB1 = wait_seg(sA + 1);
        const float* WB2 = wait_seg(sA + 2);
        for (int rd = 0; rd < rounds; ++rd) {
            const int base = (rd * NW + warp) * NTK;
            if (base >= n_eff) break;
            if (last && !ctx_last && base + NTK <= n_c) continue;
            const float* trow[NTK];
            float acc[NTK];
            float dummy[NTK][8];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i < n_eff ? base + i : n_eff - 1;
                trow[i] = T + (size_t)tok * D;
                acc[i] = WB2[(L.b2 - L.w2) + lane] + trow[i][lane];
            }
            warp_mlp<NTK, true>(acc, trow, dummy, D, WB1, WB1 + (L.b1 - L.w1), WB2, m.FF, hs, lane);
            warp_layer_norm<NTK>(acc, WB2[(L.g2 - L.w2) + lane], WB2[(L.be2 - L.w2) + lane]);
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                const int tok = base + i;
                if (tok < n_eff) X[(size_t)tok * D + lane] = acc[i];
            }
        }
        done_seg(sA + 1);
        if (tid == 0) issue(sA + 2 + 3);
    }
    if (z_tgt || z_ctx) __syncthreads();
    if (z_tgt)
        for (int i = tid; i < n_t * D; i += blockDim.x) z_tgt[(size_t)b * n_t * D + i] = X[(size_t)n_c * D + i];
    if (z_ctx)
        for (int i = tid; i < n_c * D; i += blockDim.x) z_ctx[(size_t)b * n_c * D + i] = X[i];
}

static size_t cw_ring_floats(const Dims& d, const Layout& L) {
    size_t w = L.y_w1 - L.x_w1;
    auto mx = [&](size_t v) { if (v > w) w = v; };
    mx(L.tok - L.y_w1); mx(L.w1 - L.wq); mx(L.w2 - L.w1); mx(L.layer_stride - L.w2);
    (void)d;
    return (w + 31) & ~(size_t)31;
}

struct CwPlan { int ntk, warps, n_slots; size_t smem; int wb; };

static bool cw_plan(const Dims& d, const Layout& L, int n_c, int n_tok, int kv_slots, CwPlan& p, int min_warps = 1,
                    int n_rows = 0, int B = 1) {
    if (d.D != kCwD || d.FF % 128 != 0 || d.EH % 128 != 0 || n_c > 64 || n_c < 1) return false;
    if (n_rows < 1 || n_rows > n_tok) n_rows = n_tok;      // rows actually processed (rollout mode drops dead targets)
    // Tokens per warp.  Measured at cfg2 (B = 200, us per launch; rows = context points + 2 theta tokens):
    //   rows      4     8     12    16    20    24    28    32    36
    //   1 token   17.9  23.4  31.7  39.7  52.2  58.1  67.8  75.1  88.8
    //   2 tokens  19.0  21.3  25.4  29.6  35.6  41.1  48.4  52.6  68.3
    //   4 tokens  29.6  28.5  30.0  32.0  37.6  38.2  43.6  46.1  55.9
    // Two tokens that share every weight read beat twice the warps from ~7 rows on (round 1 used one token per warp up
    // to 16 rows and two up to 32: 39.7 / 52.5 us at 16 / 32 rows).
    static const bool old_rule = [] { const char* e = getenv("ALINE_CTX_RULE"); return e && e[0] == '1'; }();   // A/B: round-1 rule
    p.ntk = n_rows <= 6 ? 1 : n_rows <= 22 ? 2 : 4;
    if (old_rule) { p.ntk = n_rows <= kCwMaxWarps ? 1 : n_rows <= 2 * kCwMaxWarps ? 2 : 4; if (min_warps > 8) min_warps = 8; }
    // Many rollouts (several blocks per SM, more than one wave): throughput bound, tokens share each weight read.
    static const int force_ntk = [] { const char* e = getenv("ALINE_CTX_NTK"); return e ? atoi(e) : 0; }();
    // Measured at B = 1000 (us, 16 / 20 / 32 rows): 1 token per warp 126 / 167 / 245, 2: 87 / 107 / 171, 4: 102 / 119 / 162.
    if (B >= 3 * device_info().sm_count && n_rows >= 4) p.ntk = n_rows <= 24 ? 2 : 4;
    if (force_ntk == 1 || force_ntk == 2 || force_ntk == 4) p.ntk = force_ntk;
    p.warps = (n_rows + p.ntk - 1) / p.ntk;
    if (p.warps > kCwMaxWarps) p.warps = kCwMaxWarps;
    // the fused select (four passes over the rollout's logits) wants all 16 warps whatever the context needs -- in the
    // latency regime: with 8 the new tokens-per-warp rule gained nothing in the cfg2 rollout (6.57 ms), with 16: 6.40 ms;
    // with many rollouts per SM (cfg1, B = 1000) the idle warps cost more than they give (6.22 -> 6.64 ms), and a few
    // hundred candidates (cfg4 / cfg5) do not need them: 8 there
    // -- and only while two blocks still share an SM (cfg4: 154 token rows of activations; 16 warps of hidden-unit
    // scratch pushed the block past half of the shared memory: 5.33 -> 5.78 ms)
    p.n_slots = kv_slots < n_tok ? kv_slots : n_tok;
    p.wb = (int)cw_ring_floats(d, L);
    auto smem_for = [&](int warps) {
        size_t fl = 3 * (size_t)p.wb + 2 * (size_t)n_tok * kCwD + (size_t)warps * p.ntk * 128 + 2 * (size_t)p.n_slots * kCwKS + 2 * n_tok;
        return fl * sizeof(float) + 16;
    };
    if (p.warps < min_warps) {
        const size_t half = (size_t)device_info().max_smem_optin / 2 - 1024;
        int w = min_warps;
        if (w > 8 && smem_for(w) > half) w = p.warps > 8 ? p.warps : 8;
        p.warps = w;
    }
    p.smem = smem_for(p.warps);
    return p.smem <= (size_t)device_info().max_smem_optin;
}

// csrc/ctx_warp64.cu: the same mapping for d = 64 / 8 heads
bool ctx_stack_warp64_supported(const Dims& d, const Layout& L, const float* P, int n_c, int n_tok, int kv_slots);
int ctx_stack_warp64(const Dims& d, const Layout& L, const float* P, const float* cx, const float* cy, int B, int n_c,
                     int ctx_cap, const float* target_x, int n_td, const int* tgt_slot, float* kv, int kv_slots,
                     float* z_tgt, float* z_ctx, void* tckv, int n_keys_tc, const SelectArgs* sel, int n_rows_hint,
                     cudaStream_t st);

bool query_tc3_fold_emitted(const Dims& d, int n_keys);     // csrc/query_tc3.cu
int query_tc3_fold_keys(const Dims& d, int n_keys);          // ... with which padded key count (0: not emitted)
bool query_tc3_fold_only();                                  // ... and nothing will read the plain blocks

bool ctx_stack_warp_supported(const Dims& d, const Layout& L, const float* P, int n_c, int n_tok, int kv_slots) {
    if (d.D == 64) return ctx_stack_warp64_supported(d, L, P, n_c, n_tok, kv_slots);
    CwPlan p;
    return ((uintptr_t)P % 16 == 0) && cw_plan(d, L, n_c, n_tok, kv_slots, p, 8);       // as the launch with a fused select would size it
}

int ctx_stack_warp(const Dims& d, const Layout& L, const float* P, const float* cx, const float* cy, int B, int n_c,
                   int ctx_cap, const float* target_x, int n_td, const int* tgt_slot, float* kv, int kv_slots,
                   float* z_tgt, float* z_ctx, void* tckv, int n_keys_tc, const SelectArgs* sel, int n_rows_hint,
                   cudaStream_t st) {
    if (d.D == 64)
        return ctx_stack_warp64(d, L, P, cx, cy, B, n_c, ctx_cap, target_x, n_td, tgt_slot, kv, kv_slots, z_tgt, z_ctx, tckv,
                                n_keys_tc, sel, n_rows_hint, st);
    const int n_tok = n_c + n_td + d.ntok;
    CwPlan p;
    ALINE_REQUIRE(cw_plan(d, L, n_c, n_tok, kv_slots, p, sel ? ((B >= 3 * device_info().sm_count || sel->nq < 1024) ? 8 : kCwMaxWarps) : 1, (z_tgt || z_ctx) ? 0 : n_rows_hint, B),
                  "ctx_stack_warp: unsupported shape");
    const SelectArgs sa = sel ? *sel : SelectArgs{};
    // 0: plain operand blocks only; 1: + the folded operands; 2: the folded operands only (inside aline_rollout, when the
    // candidate stream that follows reads nothing else)
    const int nkf = tckv != nullptr ? query_tc3_fold_keys(d, n_keys_tc) : 0;            // padded key count of the folded operands
    const int emit_fold = nkf > 0 ? (query_tc3_fold_only() ? 2 : 1) : 0;
#define ALINE_CW_LAUNCH(NTKV)                                                                                          \
    do {                                                                                                               \
        if (ensure_dyn_smem((const void*)ctx_stack_warp_kernel<NTKV>, p.smem)) return 1;                              \
        ALINE_CHECK_CUDA(launch_k(ctx_stack_warp_kernel<NTKV>, dim3(B), dim3(32 * p.warps), p.smem, st, g_pdl_chain,   \
                                  d, L, P, cx, cy, n_c, ctx_cap, target_x, n_td, tgt_slot, kv, kv_slots, B, z_tgt,    \
                                  z_ctx, p.wb, p.n_slots, (unsigned char*)tckv, n_keys_tc, sa, (int)(sel != nullptr),     \
                                  emit_fold, nkf));                                                                   \
    } while (0)
    if (p.ntk == 1) ALINE_CW_LAUNCH(1);
    else if (p.ntk == 2) ALINE_CW_LAUNCH(2);
    else ALINE_CW_LAUNCH(4);
#undef ALINE_CW_LAUNCH
    ALINE_LAUNCH_OK();
    return 0;
}

}  // namespace aline
