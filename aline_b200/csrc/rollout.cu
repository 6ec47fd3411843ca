// ALINE forward pass and resident T-step design rollout, fp32 path.
//
// Replaces, in the reference (eval mode, torch.no_grad):
//   model/embedder.py:67-214  Embedder.forward            -> embed_query_kernel / ctx_stack_kernel (embedding phase)
//   model/encoder.py:83-141   create_mask + TransformerEncoder -> ctx_stack_kernel + query_stream_kernel
//   model/head.py:27-33,319-358 AcquisitionHead + argmax  -> query_stream_kernel (logit) + select_kernel
//   model/head.py:152-186     GMMTargetHead.forward       -> gmm_head_kernel
//   tasks/base_task.py:103-154 Task.update_batch          -> select_kernel (in-place append) / move_selected_kernel
//   utils/eval.py:9-39        get_traces T-loop           -> aline_rollout
//
// The reference's [N, N] additive mask is never built: its structure (model/encoder.py:83-126) is
// compiled into the kernels.  Context rows attend to context; target rows attend to context; candidate-query rows
// attend to context + the selected targets; nothing attends to a query.  So per layer only the keys / values of the
// context and selected-target tokens are needed by the (many) query tokens: ctx_stack_kernel runs the few
// context + target tokens of one rollout per thread block through all layers and emits those K, V per layer;
// query_stream_kernel then runs every candidate independently through all layers + the acquisition MLP.
#include <mutex>
#include "model.cuh"
#include "tc.cuh"
#include "select.cuh"
#include "query_fast.cuh"

namespace aline {

constexpr int kQueryTile = 256;       // candidate-query tokens (= threads) per block, embedding kernel
// largest block (tokens = threads) of query_stream_kernel: at d = 64 the layer weights the candidate rows need (Wq, Wo,
// the MLP and the vectors -- NOT Wk / Wv, which only the context stack uses) take 101 KB of shared memory, leaving room
// for 128 tokens (round 1 staged the whole layer, 134 KB, and ran 64 tokens = 2 warps per SM)
constexpr int query_tile(int D) { return D == 32 ? 256 : 128; }

static int dims_from(const aline_model* m, Dims& d) {
    ALINE_REQUIRE(m != nullptr && m->params != nullptr, "aline_model / params is NULL");
    d.D = m->d; d.FF = m->ff; d.H = m->n_head; d.NL = m->n_layer; d.dx = m->dim_x; d.dy = m->dim_y;
    d.ntok = m->n_theta_tok; d.C = m->n_comp; d.EH = m->emb_hidden; d.HH = m->head_hidden; d.tt = m->time_token ? 1 : 0;
    d.std_min = m->std_min;
    ALINE_REQUIRE(d.D == 32 || d.D == 64, "dim_embedding %d unsupported (32 or 64)", d.D);
    ALINE_REQUIRE(d.H * 8 == d.D, "n_head %d: head_dim must be 8 (dim_embedding %d)", d.H, d.D);
    ALINE_REQUIRE(d.FF >= 32 && d.FF % 32 == 0 && d.EH % 32 == 0 && d.HH % 32 == 0 && d.EH >= 32 && d.HH >= 32,
                  "feed-forward widths must be multiples of 32 (ff=%d emb=%d head=%d)", d.FF, d.EH, d.HH);
    ALINE_REQUIRE(d.NL >= 1 && d.NL <= 16, "num_layers %d unsupported", d.NL);
    ALINE_REQUIRE(d.dx >= 1 && d.dx <= 8 && d.dy == 1, "dim_x must be 1..8 and dim_y 1 (got %d, %d)", d.dx, d.dy);
    ALINE_REQUIRE(d.C >= 1 && d.C <= 32, "num_components %d unsupported", d.C);
    Layout L = make_layout(d);
    ALINE_REQUIRE(m->n_params == L.total, "packed parameter blob has %llu floats, layout needs %llu",
                  (unsigned long long)m->n_params, (unsigned long long)L.total);
    return 0;
}

// ------------------------------------------------------ query embedding ----
// E_q = MLPx(query_x), written k-major: eq[b][k][j]  (model/embedder.py:143-147)
template <int D>
__global__ void __launch_bounds__(kQueryTile)
embed_query_kernel(const Dims m, const Layout L, const float* __restrict__ P, const float* __restrict__ qx, int nq,
                   float* __restrict__ eq, float* __restrict__ eq_rm) {
    extern __shared__ __align__(16) float smem[];
    const int n_w = (int)(L.y_w1 - L.x_w1);
    stage_floats(smem, P + L.x_w1, n_w);
    __syncthreads();
    const int b = blockIdx.y, j = blockIdx.x * kQueryTile + threadIdx.x;
    if (j >= nq) return;
    float xin[8];
    for (int k = 0; k < m.dx; ++k) xin[k] = __ldg(qx + ((size_t)b * nq + j) * m.dx + k);
    float e[D];
#pragma unroll
    for (int i = 0; i < D; ++i) e[i] = 0.f;
    embed_mlp<D>(e, xin, m.dx, smem, smem + (L.x_b1 - L.x_w1), smem + (L.x_w2 - L.x_w1), smem + (L.x_b2 - L.x_w1), m.EH);
    if (eq) {
#pragma unroll
        for (int i = 0; i < D; ++i) eq[((size_t)b * D + i) * nq + j] = e[i];
    }
    if (eq_rm) {                                   // row-major copy for the two-threads-per-row tensor-core stream
        float4* o = reinterpret_cast<float4*>(eq_rm + ((size_t)b * nq + j) * D);
#pragma unroll
        for (int i = 0; i < D / 4; ++i) o[i] = make_float4(e[4 * i], e[4 * i + 1], e[4 * i + 2], e[4 * i + 3]);
    }
}

// ------------------------------------------------ context + target stack ----
// One block per rollout b.  Token t < n_c is context point t, then n_td data-target tokens, then the theta tokens.
// Each token is owned by G = D/8 adjacent lanes (one per attention head): lane g computes outputs [8g, 8g+8) of
// every projection, runs head g of the attention, and a 1/G slice of the MLP hidden units; LayerNorm statistics
// and the second MLP projection are combined with warp shuffles.  Emits per layer the K, V rows of the context
// tokens (slots 0..n_c-1) and of the selected targets (slot n_c + tgt_slot[i]) -- fp32, and optionally as the bf16
// operand blocks of the fast tensor-core query stream (csrc/query_tc3.cu: keys relative to key 0 + mask chunk,
// values + ones row) -- and optionally the final target encodings z_tgt.
// Sum over the G lanes of a token.  Called from warp-uniform control flow only (every lane of the warp executes the
// row code; lanes without a row just do not store): per-group member masks would split the warp into G-lane
// fragments that then run one after the other.
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LayerNorm over D = 8 G features held 8 per lane
template <int G>
__device__ __forceinline__ void group_layer_norm(float (&v)[8], const float* g, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mu = group_sum<G>(s) * (1.0f / (8 * G));
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float d = v[i] - mu; q = fmaf(d, d, q); }
    const float rstd = 1.0f / sqrtf(group_sum<G>(q) * (1.0f / (8 * G)) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mu) * rstd * g[i] + b[i];
}

// out8 += this lane's 8 outputs of a 2-layer MLP (in -> HID relu -> D): the lane evaluates HID/G hidden units,
// multiplies them into all D outputs, and the partial outputs are summed over the G lanes
template <int D>
__device__ __forceinline__ void group_mlp(float (&out8)[8], int g, const float* xin, int IN, bool x_in_smem, int xstride,
                                          const float* W1T, const float* b1, const float* W2T, int HID) {
    constexpr int G = D / 8;
    float acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.f;
    const int per = HID / G;
    for (int c = g * per; c < (g + 1) * per; c += 16) {
        float h[16];
        load_vec<16>(h, b1 + c);
        for (int k = 0; k < IN; ++k) {
            const float xk = x_in_smem ? xin[k * xstride] : xin[k];
            const float4* w = reinterpret_cast<const float4*>(W1T + (size_t)k * HID + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 ww = w[j];
                h[4 * j] = fmaf(xk, ww.x, h[4 * j]); h[4 * j + 1] = fmaf(xk, ww.y, h[4 * j + 1]);
                h[4 * j + 2] = fmaf(xk, ww.z, h[4 * j + 2]); h[4 * j + 3] = fmaf(xk, ww.w, h[4 * j + 3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) h[j] = fmaxf(h[j], 0.f);
        matvec_reg<16, D>(acc, h, W2T + (size_t)c * D, D);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = group_sum<G>(acc[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = 0.f;
#pragma unroll
        for (int gg = 0; gg < G; ++gg) v = (g == gg) ? acc[8 * gg + i] : v;
        out8[i] += v;
    }
}

constexpr int ctx_max_threads(int D) { return D == 32 ? 640 : 512; }

template <int D>
__global__ void __launch_bounds__(ctx_max_threads(D))
ctx_stack_kernel(const Dims m, const Layout L, const float* __restrict__ P, const float* __restrict__ cx,
                 const float* __restrict__ cy, int n_c, int ctx_cap, const float* __restrict__ target_x, int n_td,
                 const int* __restrict__ tgt_slot, float* __restrict__ kv, int kv_slots, int B,
                 float* __restrict__ z_tgt, float* __restrict__ z_ctx, int w_floats, int NT,
                 unsigned char* __restrict__ tckv, int n_keys_tc, int emit_fold) {
    constexpr int G = D / 8;
    extern __shared__ __align__(16) float smem[];
    float* Wsm = smem;                             // [w_floats]
    float* Ks = Wsm + w_floats;                    // [n_c][D]
    float* Vs = Ks + (size_t)n_c * D;              // [n_c][D]
    float* X = Vs + (size_t)n_c * D;               // [D][NT]   activations, one column per token
    float* T = X + (size_t)D * NT;                 // [D][NT]   scratch column
    const int b = blockIdx.x, tid = threadIdx.x;
    const int tok = tid / G, g = tid % G;
    const int n_t = n_td + m.ntok, n_tok = n_c + n_t;
    const bool live = tok < n_tok;
    const int tokc = tok < NT ? tok : NT - 1;      // lanes without a token compute on a valid column, never store
    float* xcol = X + tokc;
    float* tcol = T + tokc;

    // ---- embedding (model/embedder.py:128-214) ----
    const int n_emb = (int)(L.tok - L.x_w1);
    stage_floats(Wsm, P + L.x_w1, n_emb);
    __syncthreads();
    {
        // warp-uniform: every lane evaluates both embedders (on zeros where they do not apply), then selects
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = 0.f;
        const int ti = tok - n_c;
        const bool is_ctx = live && tok < n_c, is_data = live && (tok < n_c || ti < n_td);
        float xin[8], yin[1] = {0.f};
#pragma unroll
        for (int k = 0; k < 8; ++k) xin[k] = 0.f;
        if (is_data) {
            const float* src = tok < n_c ? cx + ((size_t)b * ctx_cap + tok) * m.dx : target_x + ((size_t)b * n_td + ti) * m.dx;
            for (int k = 0; k < m.dx; ++k) xin[k] = __ldg(src + k);
        }
        if (is_ctx) yin[0] = __ldg(cy + (size_t)b * ctx_cap + tok);
        group_mlp<D>(e, g, xin, m.dx, false, 0, Wsm, Wsm + (L.x_b1 - L.x_w1), Wsm + (L.x_w2 - L.x_w1), m.EH);
        float ey[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ey[i] = 0.f;
        group_mlp<D>(ey, g, yin, 1, false, 0, Wsm + (L.y_w1 - L.x_w1), Wsm + (L.y_b1 - L.x_w1), Wsm + (L.y_w2 - L.x_w1), m.EH);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            e[i] += Wsm[(L.x_b2 - L.x_w1) + 8 * g + i];
            if (is_ctx) e[i] += ey[i] + Wsm[(L.y_b2 - L.x_w1) + 8 * g + i];
        }
        if (live && !is_data) {
            const float* tk = P + L.tok + (size_t)(ti - n_td) * D + 8 * g;
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = __ldg(tk + i);
        }
        if (live) {
#pragma unroll
            for (int i = 0; i < 8; ++i) xcol[(8 * g + i) * NT] = e[i];
        }
    }
    __syncthreads();

    // ---- encoder layers ----
    for (int l = 0; l < m.NL; ++l) {
        stage_floats(Wsm, P + L.layer0 + (size_t)l * L.layer_stride, (int)L.layer_stride);
        __syncthreads();
        // the last layer's outputs are only needed for the targets (no value head), and only if z_tgt is wanted
        const bool last = l + 1 == m.NL;
        const bool run_row = live && (!last || (tok >= n_c && z_tgt != nullptr) || (tok < n_c && z_ctx != nullptr));
        float q8[8], k8[8], v8[8];
        int slot = -1;
        if (tckv) {                                    // clear this (layer, rollout) operand block, set the key mask
            const int nkp = (n_keys_tc + 15) / 16 * 16, kbytes = (G + 1) * 16 * nkp, blk_bytes = kbytes + G * 32 * nkp;
            unsigned char* blk = tckv + ((size_t)l * B + b) * blk_bytes;
            for (int i = tid * 16; i < blk_bytes; i += blockDim.x * 16) {
                uint4 z = make_uint4(0, 0, 0, 0);
                const int mrow = (i - G * 16 * nkp) >> 4;          // row of the mask chunk (chunk G of the K part)
                if (i >= G * 16 * nkp && i < kbytes && mrow >= n_keys_tc) z.x = 0xC348u;   // bf16(-200) in element 0
                *reinterpret_cast<uint4*>(blk + i) = z;
            }
        }
        if (live) {
            slot = tok < n_c ? tok : -1;
            if (tok >= n_c) { int sidx = __ldg(tgt_slot + (tok - n_c)); slot = sidx >= 0 ? n_c + sidx : -1; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                q8[i] = Wsm[L.bq + 8 * g + i]; k8[i] = Wsm[L.bk + 8 * g + i]; v8[i] = Wsm[L.bv + 8 * g + i];
            }
            if (slot >= 0 || run_row) {
#pragma unroll 4
                for (int k = 0; k < D; ++k) {
                    const float xk = xcol[k * NT];
                    const float4* wq = reinterpret_cast<const float4*>(Wsm + L.wq + (size_t)k * D + 8 * g);
                    const float4* wk = reinterpret_cast<const float4*>(Wsm + L.wk + (size_t)k * D + 8 * g);
                    const float4* wv = reinterpret_cast<const float4*>(Wsm + L.wv + (size_t)k * D + 8 * g);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float4 a = wq[j], c = wk[j], e = wv[j];
                        q8[4 * j] = fmaf(xk, a.x, q8[4 * j]); q8[4 * j + 1] = fmaf(xk, a.y, q8[4 * j + 1]);
                        q8[4 * j + 2] = fmaf(xk, a.z, q8[4 * j + 2]); q8[4 * j + 3] = fmaf(xk, a.w, q8[4 * j + 3]);
                        k8[4 * j] = fmaf(xk, c.x, k8[4 * j]); k8[4 * j + 1] = fmaf(xk, c.y, k8[4 * j + 1]);
                        k8[4 * j + 2] = fmaf(xk, c.z, k8[4 * j + 2]); k8[4 * j + 3] = fmaf(xk, c.w, k8[4 * j + 3]);
                        v8[4 * j] = fmaf(xk, e.x, v8[4 * j]); v8[4 * j + 1] = fmaf(xk, e.y, v8[4 * j + 1]);
                        v8[4 * j + 2] = fmaf(xk, e.z, v8[4 * j + 2]); v8[4 * j + 3] = fmaf(xk, e.w, v8[4 * j + 3]);
                    }
                }
            }
            if (slot >= 0) {
                float* gk = kv + (((size_t)l * B + b) * kv_slots + slot) * (2 * D) + 8 * g;
                reinterpret_cast<float4*>(gk)[0] = make_float4(k8[0], k8[1], k8[2], k8[3]);
                reinterpret_cast<float4*>(gk)[1] = make_float4(k8[4], k8[5], k8[6], k8[7]);
                reinterpret_cast<float4*>(gk + D)[0] = make_float4(v8[0], v8[1], v8[2], v8[3]);
                reinterpret_cast<float4*>(gk + D)[1] = make_float4(v8[4], v8[5], v8[6], v8[7]);
                if (tok < n_c) {
                    float* sk = Ks + (size_t)tok * D + 8 * g;
                    float* sv = Vs + (size_t)tok * D + 8 * g;
                    reinterpret_cast<float4*>(sk)[0] = make_float4(k8[0], k8[1], k8[2], k8[3]);
                    reinterpret_cast<float4*>(sk)[1] = make_float4(k8[4], k8[5], k8[6], k8[7]);
                    reinterpret_cast<float4*>(sv)[0] = make_float4(v8[0], v8[1], v8[2], v8[3]);
                    reinterpret_cast<float4*>(sv)[1] = make_float4(v8[4], v8[5], v8[6], v8[7]);
                }
            }
        }
        __syncthreads();
        if (tckv && slot >= 0) {
            // bf16 operands of the fast tensor-core query stream (csrc/query_tc3.cu, query_tc5.cu).  K part: chunk g (= head) row
            // `slot` = K[slot] - K[0] (the softmax is evaluated relative to key 0); V part: head g, 16-row chunks of 8
            // keys: rows 0..7 = features, row 8 = 1 (returns the softmax denominator), rows 9..15 = 0
            const int nkp = (n_keys_tc + 15) / 16 * 16;
            const int kbytes = (G + 1) * 16 * nkp;
            unsigned char* blk = tckv + ((size_t)l * B + b) * (kbytes + G * 32 * nkp);
            const float* k0 = Ks + 8 * g;
            uint4 q4;
            q4.x = tc::pack_bf16(k8[0] - k0[0], k8[1] - k0[1]); q4.y = tc::pack_bf16(k8[2] - k0[2], k8[3] - k0[3]);
            q4.z = tc::pack_bf16(k8[4] - k0[4], k8[5] - k0[5]); q4.w = tc::pack_bf16(k8[6] - k0[6], k8[7] - k0[7]);
            *reinterpret_cast<uint4*>(blk + ((size_t)g * nkp + slot) * 16) = q4;
            __nv_bfloat16* vb = reinterpret_cast<__nv_bfloat16*>(blk + kbytes) + ((size_t)g * (nkp / 8) + (slot >> 3)) * 128 + (slot & 7);
#pragma unroll
            for (int f = 0; f < 8; ++f) vb[f * 8] = __float2bfloat16_rn(v8[f]);
            vb[64] = __float2bfloat16_rn(1.0f);
        }
        if (last && z_tgt == nullptr && z_ctx == nullptr) break;   // rollout mode: nothing downstream of the last K, V
        // head g of the attention over the context keys, then the rest of the layer.  Executed by EVERY lane
        // (warp-uniform: the shuffles below need all 32 lanes); only rows that continue store anything.
        {
#pragma unroll
            for (int i = 0; i < 8; ++i) q8[i] *= 0.35355339059327376220f;
            float mx = -INFINITY;
            for (int j = 0; j < n_c; ++j) {
                const float4* kr = reinterpret_cast<const float4*>(Ks + (size_t)j * D + 8 * g);
                float4 a = kr[0], c = kr[1];
                float sdot = q8[0] * a.x;
                sdot = fmaf(q8[1], a.y, sdot); sdot = fmaf(q8[2], a.z, sdot); sdot = fmaf(q8[3], a.w, sdot);
                sdot = fmaf(q8[4], c.x, sdot); sdot = fmaf(q8[5], c.y, sdot); sdot = fmaf(q8[6], c.z, sdot);
                sdot = fmaf(q8[7], c.w, sdot);
                mx = fmaxf(mx, sdot);
            }
            float o8[8], den = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) o8[i] = 0.f;
            for (int j = 0; j < n_c; ++j) {
                const float4* kr = reinterpret_cast<const float4*>(Ks + (size_t)j * D + 8 * g);
                const float4* vr = reinterpret_cast<const float4*>(Vs + (size_t)j * D + 8 * g);
                float4 a = kr[0], c = kr[1];
                float sdot = q8[0] * a.x;
                sdot = fmaf(q8[1], a.y, sdot); sdot = fmaf(q8[2], a.z, sdot); sdot = fmaf(q8[3], a.w, sdot);
                sdot = fmaf(q8[4], c.x, sdot); sdot = fmaf(q8[5], c.y, sdot); sdot = fmaf(q8[6], c.z, sdot);
                sdot = fmaf(q8[7], c.w, sdot);
                const float p = expf(sdot - mx);
                den += p;
                float4 va = vr[0], vb2 = vr[1];
                o8[0] = fmaf(p, va.x, o8[0]); o8[1] = fmaf(p, va.y, o8[1]); o8[2] = fmaf(p, va.z, o8[2]);
                o8[3] = fmaf(p, va.w, o8[3]); o8[4] = fmaf(p, vb2.x, o8[4]); o8[5] = fmaf(p, vb2.y, o8[5]);
                o8[6] = fmaf(p, vb2.z, o8[6]); o8[7] = fmaf(p, vb2.w, o8[7]);
            }
            const float inv = 1.0f / den;
            if (run_row) {
#pragma unroll
                for (int i = 0; i < 8; ++i) tcol[(8 * g + i) * NT] = o8[i] * inv;
            }
            __syncwarp();
            // out-projection slice + residual, LayerNorm 1
            float h8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h8[i] = Wsm[L.bo + 8 * g + i] + xcol[(8 * g + i) * NT];
#pragma unroll 4
            for (int k = 0; k < D; ++k) {
                const float ok = tcol[k * NT];
                const float4* w = reinterpret_cast<const float4*>(Wsm + L.wo + (size_t)k * D + 8 * g);
                float4 a = w[0], c = w[1];
                h8[0] = fmaf(ok, a.x, h8[0]); h8[1] = fmaf(ok, a.y, h8[1]); h8[2] = fmaf(ok, a.z, h8[2]);
                h8[3] = fmaf(ok, a.w, h8[3]); h8[4] = fmaf(ok, c.x, h8[4]); h8[5] = fmaf(ok, c.y, h8[5]);
                h8[6] = fmaf(ok, c.z, h8[6]); h8[7] = fmaf(ok, c.w, h8[7]);
            }
            group_layer_norm<G>(h8, Wsm + L.g1 + 8 * g, Wsm + L.be1 + 8 * g);
            __syncwarp();                              // every lane has read the attention output column
            if (run_row) {
#pragma unroll
                for (int i = 0; i < 8; ++i) tcol[(8 * g + i) * NT] = h8[i];
            }
            __syncwarp();
            // MLP (hidden units split over the lanes) + residual, LayerNorm 2
#pragma unroll
            for (int i = 0; i < 8; ++i) h8[i] += Wsm[L.b2 + 8 * g + i];
            group_mlp<D>(h8, g, tcol, D, true, NT, Wsm + L.w1, Wsm + L.b1, Wsm + L.w2, m.FF);
            group_layer_norm<G>(h8, Wsm + L.g2 + 8 * g, Wsm + L.be2 + 8 * g);
            if (run_row) {
#pragma unroll
                for (int i = 0; i < 8; ++i) xcol[(8 * g + i) * NT] = h8[i];
            }
        }
        __syncthreads();
    }
    if (z_ctx && live && tok < n_c) {                  // value head input (model/head.py:368-370)
        float* z = z_ctx + ((size_t)b * n_c + tok) * D + 8 * g;
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = xcol[(8 * g + i) * NT];
    }
    if (z_tgt && live && tok >= n_c) {
        float* z = z_tgt + ((size_t)b * n_t + (tok - n_c)) * D + 8 * g;
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = xcol[(8 * g + i) * NT];
    }
    if constexpr (D == 32) {
        if (emit_fold) {                               // folded operands of the candidate stream (query_fast.cuh)
            __syncthreads();                           // this block's plain operand blocks are complete
            tcq::fold_kv_emit(tckv, b, B, (n_keys_tc + 15) / 16 * 16, emit_fold, m.NL, P, L, (int)threadIdx.x, (int)blockDim.x);
        }
    }
}

// ------------------------------------------------- candidate-query stream ----
// grid (tiles, B); thread = one candidate of rollout b.  All layers + acquisition logit, candidates never interact.
template <int D>
__global__ void __launch_bounds__(query_tile(D), D == 32 ? 2 : 1)
query_stream_kernel(const Dims m, const Layout L, const float* __restrict__ P, const float* __restrict__ eq,
                    const unsigned char* __restrict__ alive, int nq, const float* __restrict__ kv, int n_keys,
                    int kv_slots, int B, float t_value, float* __restrict__ logits, float* __restrict__ zq,
                    int w_floats, const int* __restrict__ flag = nullptr, int epoch = 0) {
    extern __shared__ __align__(16) float smem[];
    if (flag != nullptr) {             // conditional fallback of the fast tensor-core kernel: runs only if it flagged this launch
        pdl_wait();
        if (*flag != epoch) return;
    }
    const int NT = blockDim.x;
    float* Wsm = smem;
    float* Ks = Wsm + w_floats;
    float* Vs = Ks + (size_t)n_keys * D;
    float* X = Vs + (size_t)n_keys * D;
    float* T = X + (size_t)D * NT;
    const int b = blockIdx.y, tid = threadIdx.x, j = blockIdx.x * NT + tid;
    const bool live = j < nq && (alive == nullptr || alive[(size_t)b * nq + j] != 0);
    float* xcol = X + tid;
    float* tcol = T + tid;
    if (live) {
#pragma unroll
        for (int i = 0; i < D; ++i) xcol[i * NT] = __ldg(eq + ((size_t)b * D + i) * nq + j);
    }
    // the candidate rows never use Wk / Wv / bk / bv: stage [Wq] and [bq .. be2] back to back and address them through a
    // layout whose offsets behind Wq are shifted down by the two skipped matrices
    Layout Lq = L;
    const size_t skip = 2 * (size_t)D * D;
    Lq.bq -= skip; Lq.bk -= skip; Lq.bv -= skip; Lq.wo -= skip; Lq.bo -= skip; Lq.g1 -= skip; Lq.be1 -= skip;
    Lq.w1 -= skip; Lq.b1 -= skip; Lq.w2 -= skip; Lq.b2 -= skip; Lq.g2 -= skip; Lq.be2 -= skip;
    for (int l = 0; l < m.NL; ++l) {
        __syncthreads();                      // previous layer done with Wsm / Ks / Vs
        const float* Pl = P + L.layer0 + (size_t)l * L.layer_stride;
        stage_floats(Wsm, Pl + L.wq, D * D);
        stage_floats(Wsm + D * D, Pl + L.bq, (int)(L.layer_stride - L.bq));
        const float* g = kv + ((size_t)l * B + b) * kv_slots * (2 * D);
        for (int i = tid; i < n_keys * (D / 4); i += NT) {
            int key = i / (D / 4), c4 = i - key * (D / 4);
            const float4* src = reinterpret_cast<const float4*>(g + (size_t)key * 2 * D);
            reinterpret_cast<float4*>(Ks + (size_t)key * D)[c4] = __ldg(src + c4);
            reinterpret_cast<float4*>(Vs + (size_t)key * D)[c4] = __ldg(src + D / 4 + c4);
        }
        __syncthreads();
        if (live) encoder_layer_token<D>(xcol, tcol, NT, Wsm, Lq, m.FF, Ks, Vs, n_keys);
    }
    __syncthreads();
    // ---- acquisition MLP (model/head.py:27-31): logit = w2 . relu(W1 [z ; t] + b1) + b2 ----
    const int n_acq = (int)(L.gmm0 - L.a_w1);
    stage_floats(Wsm, P + L.a_w1, n_acq);
    __syncthreads();
    if (j < nq) {
        float logit = -INFINITY;
        if (live) {
            const float* W1 = Wsm, *b1 = Wsm + (L.a_b1 - L.a_w1), *w2 = Wsm + (L.a_w2 - L.a_w1);
            logit = Wsm[L.a_b2 - L.a_w1];
            for (int c = 0; c < m.HH; c += 32) {
                float hid[32];
                load_vec<32>(hid, b1 + c);
                matvec_col<32>(hid, xcol, NT, D, W1 + c, m.HH);
                if (m.tt) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) hid[jj] = fmaf(t_value, W1[(size_t)D * m.HH + c + jj], hid[jj]);
                }
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) logit = fmaf(fmaxf(hid[jj], 0.f), w2[c + jj], logit);
            }
            if (zq) {
                float* z = zq + ((size_t)b * nq + j) * D;
#pragma unroll
                for (int i = 0; i < D; ++i) z[i] = xcol[i * NT];
            }
        }
        logits[(size_t)b * nq + j] = logit;
    }
}

// --------------------------------------------------- softmax / argmax / append ----
// One block per rollout.  zt = softmax over the live candidates; idx = first argmax of zt; log_prob = log(zt[idx])
// (model/head.py:355-358); then the chosen (x, y) is appended to the context in place and the candidate retired
// (tasks/base_task.py:133-154 without the compaction).  idx_out is the index in the *compacted* live set, which is
// what the reference's design_out.idx means.  Body: select_block (csrc/select.cuh).
__global__ void __launch_bounds__(256)
select_kernel(const float* __restrict__ logits, unsigned char* __restrict__ alive, int nq, const float* __restrict__ qx,
              const float* __restrict__ qy, int dx, int dy, float* __restrict__ cx, float* __restrict__ cy, int n_c,
              int ctx_cap, long long* __restrict__ idx_out, int idx_stride, float* __restrict__ logp_out,
              int logp_stride, long long* __restrict__ idx_orig_out, float* __restrict__ zt, int sample = 0,
              unsigned long long seed = 0, int step = 0) {
    SelectArgs a{logits, alive, nq, qx, qy, dx, dy, cx, cy, n_c, ctx_cap, idx_out, idx_stride, logp_out, logp_stride,
                 idx_orig_out, zt};
    a.sample = sample; a.seed = seed; a.step = step;
    select_block(a, blockIdx.x);
}

// ------------------------------------------------------------- GMM head ----
// z [n_tok][D] -> means / stds / weights [n_tok][C]   (model/head.py:152-186, 252-266) and / or the predictive
// variance of the mixture  var = sum_c w_c (sigma_c^2 + (mu_c - sum_c w_c mu_c)^2)  (utils/misc.py:244-279), fused so
// that the uncertainty-sampling baseline never materialises posterior_out_query.  Any output may be NULL.
template <int D>
__global__ void __launch_bounds__(128)
gmm_head_kernel(const Dims m, const Layout L, const float* __restrict__ P, const float* __restrict__ z, long long n_tok,
                float* __restrict__ means, float* __restrict__ stds, float* __restrict__ weights,
                float* __restrict__ var_out) {
    extern __shared__ __align__(16) float smem[];
    const int NT = blockDim.x;
    float* Wsm = smem;                               // one component's weights
    float* X = Wsm + L.gmm_stride;                   // [D][NT]
    float* R = X + (size_t)D * NT;                   // raw weights [C][NT]
    float* Ms = R + (size_t)m.C * NT;                // means [C][NT]
    float* Ss = Ms + (size_t)m.C * NT;               // stds [C][NT]
    const int tid = threadIdx.x;
    const long long tok = (long long)blockIdx.x * NT + tid;
    const bool live = tok < n_tok;
    if (live) {
#pragma unroll
        for (int i = 0; i < D; ++i) X[i * NT + tid] = __ldg(z + tok * D + i);
    }
    for (int c = 0; c < m.C; ++c) {
        __syncthreads();
        stage_floats(Wsm, P + L.gmm0 + (size_t)c * L.gmm_stride, (int)L.gmm_stride);
        __syncthreads();
        if (live) {
            const float* W1 = Wsm + L.g_w1, *b1 = Wsm + L.g_b1, *W2 = Wsm + L.g_w2, *b2 = Wsm + L.g_b2;
            float o0 = b2[0], o1 = b2[1], o2 = b2[2];
            for (int cc = 0; cc < m.HH; cc += 32) {
                float hid[32];
                load_vec<32>(hid, b1 + cc);
                matvec_col<32>(hid, X + tid, NT, D, W1 + cc, m.HH);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float h = fmaxf(hid[j], 0.f);
                    o0 = fmaf(h, W2[cc + j], o0);
                    o1 = fmaf(h, W2[m.HH + cc + j], o1);
                    o2 = fmaf(h, W2[2 * m.HH + cc + j], o2);
                }
            }
            const float sd = (o1 > 20.f ? o1 : log1pf(expf(o1))) + m.std_min;             // softplus + std_min
            if (means) means[tok * m.C + c] = o0;
            if (stds) stds[tok * m.C + c] = sd;
            Ms[c * NT + tid] = o0;
            Ss[c * NT + tid] = sd;
            R[c * NT + tid] = o2;
        }
    }
    if (live) {                                                                          // softmax over components
        float mx = -INFINITY;
        for (int c = 0; c < m.C; ++c) mx = fmaxf(mx, R[c * NT + tid]);
        float s = 0.f;
        for (int c = 0; c < m.C; ++c) { float e = expf(R[c * NT + tid] - mx); R[c * NT + tid] = e; s += e; }
        for (int c = 0; c < m.C; ++c) {
            const float w = R[c * NT + tid] / s;
            R[c * NT + tid] = w;
            if (weights) weights[tok * m.C + c] = w;
        }
        if (var_out) {
            float wm = 0.f;
            for (int c = 0; c < m.C; ++c) wm += R[c * NT + tid] * Ms[c * NT + tid];
            float v = 0.f;
            for (int c = 0; c < m.C; ++c) {
                const float dm = Ms[c * NT + tid] - wm, sd = Ss[c * NT + tid];
                v += R[c * NT + tid] * (sd * sd + dm * dm);
            }
            var_out[tok] = v;
        }
    }
}

// calculate_gmm_variance (utils/misc.py:244-279) on given mixture parameters [n][C]; the weights row of token i is
// i / tok_per_w (tok_per_w = 1: per-token weights [n][C]; n_query: weights [B][C] shared by a rollout's queries)
__global__ void gmm_variance_kernel(const float* __restrict__ means, const float* __restrict__ stds,
                                    const float* __restrict__ weights, long long n, int C, long long tok_per_w,
                                    float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* w = weights + (i / tok_per_w) * C;
    float wm = 0.f;
    for (int c = 0; c < C; ++c) wm += w[c] * means[i * C + c];
    float v = 0.f;
    for (int c = 0; c < C; ++c) {
        const float dm = means[i * C + c] - wm, sd = stds[i * C + c];
        v += w[c] * (sd * sd + dm * dm);
    }
    out[i] = v;
}

// ValueHead.forward (model/head.py:84-111): mean over the context tokens of  w2 . relu(W1 z + b1) + b2.
// One block per rollout, one thread per hidden unit; W1 / w2 in torch's own [out][in] layout (the value head is not part
// of the packed blob: no shipped config enables it).
__global__ void value_head_kernel(const float* __restrict__ z_ctx, int n_c, int D, int FF, const float* __restrict__ w1,
                                  const float* __restrict__ b1, const float* __restrict__ w2,
                                  const float* __restrict__ b2, float* __restrict__ value) {
    extern __shared__ float zs[];                        // [n_c][D]
    __shared__ float red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < n_c * D; i += blockDim.x) zs[i] = z_ctx[(size_t)b * n_c * D + i];
    __syncthreads();
    float acc = 0.f;
    for (int j = tid; j < FF; j += blockDim.x) {
        const float* w = w1 + (size_t)j * D;
        const float bj = b1[j], wj = w2[j];
        for (int t = 0; t < n_c; ++t) {
            float h = bj;
            for (int k = 0; k < D; ++k) h = fmaf(w[k], zs[t * D + k], h);
            acc = fmaf(fmaxf(h, 0.f), wj, acc);
        }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s += red[w];
        value[b] = s / (float)n_c + b2[0];
    }
}

// logsumexp_c( Normal(mu_c, sigma_c).log_prob(v) + log w_c )   (utils/eval.py:200-207)
__global__ void gmm_loglik_kernel(const float* __restrict__ value, const float* __restrict__ means,
                                  const float* __restrict__ stds, const float* __restrict__ weights, long long n, int C,
                                  float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = value[i];
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
        float mu = means[i * C + c], sd = stds[i * C + c];
        float d = v - mu;
        float lp = -(d * d) / (2.0f * (sd * sd)) - logf(sd) - 0.91893853320467274178f + logf(weights[i * C + c]);
        mx = fmaxf(mx, lp);
    }
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
        float mu = means[i * C + c], sd = stds[i * C + c];
        float d = v - mu;
        float lp = -(d * d) / (2.0f * (sd * sd)) - logf(sd) - 0.91893853320467274178f + logf(weights[i * C + c]);
        s += expf(lp - mx);
    }
    out[i] = mx + logf(s);
}

// Task.update_batch, out of place: drop row idx[b] of query [B,N,D] (order preserving), append it to ctx [B,M,D]
__global__ void move_selected_kernel(const float* __restrict__ query, const float* __restrict__ ctx,
                                     const long long* __restrict__ idx, int N, int M, int D,
                                     float* __restrict__ new_query, float* __restrict__ new_ctx) {
    const int b = blockIdx.x;
    const long long sel = idx[b];
    const float* q = query + (size_t)b * N * D;
    float* nq = new_query + (size_t)b * (N - 1) * D;
    for (int i = threadIdx.x; i < (N - 1) * D; i += blockDim.x) {
        int row = i / D, col = i - row * D;
        nq[i] = q[(size_t)(row + (row >= sel ? 1 : 0)) * D + col];
    }
    const float* c = ctx + (size_t)b * M * D;
    float* nc = new_ctx + (size_t)b * (M + 1) * D;
    for (int i = threadIdx.x; i < M * D; i += blockDim.x) nc[i] = c[i];
    for (int i = threadIdx.x; i < D; i += blockDim.x) nc[(size_t)M * D + i] = q[(size_t)sel * D + i];
}

// ------------------------------------------------------------ host side ----
static size_t layer_w_floats(const Dims& d, const Layout& L) {
    size_t a = L.layer_stride, e = L.tok - L.x_w1, q = L.gmm0 - L.a_w1;
    size_t w = a > e ? a : e;
    return pad4(w > q ? w : q);
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
    ALINE_REQUIRE(bytes <= (size_t)device_info().max_smem_optin, "kernel needs %zu bytes of shared memory (max %d)",
                  bytes, device_info().max_smem_optin);
    return ensure_dyn_smem((const void*)kernel, bytes);
}

static int embed_queries(const Dims& d, const Layout& L, const float* P, const float* qx, int B, int nq, float* eq,
                         float* eq_rm, cudaStream_t st) {
    size_t smem = (L.y_w1 - L.x_w1) * sizeof(float);
    dim3 grid(ceil_div(nq, kQueryTile), B);
    if (d.D == 32) {
        if (set_smem(embed_query_kernel<32>, smem)) return 1;
        embed_query_kernel<32><<<grid, kQueryTile, smem, st>>>(d, L, P, qx, nq, eq, eq_rm);
    } else {
        if (set_smem(embed_query_kernel<64>, smem)) return 1;
        embed_query_kernel<64><<<grid, kQueryTile, smem, st>>>(d, L, P, qx, nq, eq, eq_rm);
    }
    ALINE_LAUNCH_OK();
    return 0;
}

// csrc/ctx_warp.cu: warp-per-token kernel (d = 32), the default when the shape has one
bool ctx_stack_warp_supported(const Dims& d, const Layout& L, const float* P, int n_c, int n_tok, int kv_slots);
bool query_tc3_fold_emitted(const Dims& d, int n_keys);     // csrc/query_tc3.cu: do the context kernels emit K' / V'?
int query_tc3_fold_keys(const Dims& d, int n_keys);         // ... and with which padded key count (0: no)
int ctx_stack_warp(const Dims& d, const Layout& L, const float* P, const float* cx, const float* cy, int B, int n_c,
                   int ctx_cap, const float* target_x, int n_td, const int* tgt_slot, float* kv, int kv_slots,
                   float* z_tgt, float* z_ctx, void* tckv, int n_keys_tc, const SelectArgs* sel, int n_rows_hint,
                   cudaStream_t st);

static bool ctx_warp_enabled() {                       // ALINE_CTX_KERNEL=head: A/B switch to the lane-per-head kernel
    static const bool on = [] {
        const char* e = getenv("ALINE_CTX_KERNEL");
        return !(e && e[0] == 'h');
    }();
    return on;
}
static bool ctx_fuse_select_enabled() {                // ALINE_FUSE_SELECT=0: A/B switch, separate select kernel
    static const bool on = [] {
        const char* e = getenv("ALINE_FUSE_SELECT");
        return !(e && e[0] == '0');
    }();
    return on;
}
// will ctx_stack run the previous step's select in its prologue for this shape?
static bool ctx_fuses_select(const Dims& d, const Layout& L, const float* P, int n_c, int n_tok, int kv_slots) {
    return ctx_warp_enabled() && ctx_fuse_select_enabled() && ctx_stack_warp_supported(d, L, P, n_c, n_tok, kv_slots);
}

static int ctx_stack(const Dims& d, const Layout& L, const float* P, const float* cx, const float* cy, int B, int n_c,
                     int ctx_cap, const float* target_x, int n_td, const int* tgt_slot, float* kv, int kv_slots,
                     float* z_tgt, float* z_ctx, void* tckv, int n_keys_tc, cudaStream_t st,
                     const SelectArgs* sel = nullptr, bool* sel_fused = nullptr, int n_rows_hint = 0) {
    if (sel_fused) *sel_fused = false;
    ALINE_REQUIRE(!tckv || ((d.D == 32 || (d.D == 64 && d.H == 8)) && n_keys_tc >= n_c && n_keys_tc <= 160 && n_keys_tc <= kv_slots),
                  "ctx_stack: bf16 key / value operand blocks need d = 32 (or d = 64 with 8 heads) and n_c <= n_keys (%d) <= "
                  "min(160, kv_slots %d)", n_keys_tc, kv_slots);
    const int n_tok = n_c + n_td + d.ntok;
    if (ctx_warp_enabled() && ctx_stack_warp_supported(d, L, P, n_c, n_tok, kv_slots)) {
        const SelectArgs* s2 = (sel && ctx_fuse_select_enabled()) ? sel : nullptr;
        if (sel_fused) *sel_fused = s2 != nullptr;
        return ctx_stack_warp(d, L, P, cx, cy, B, n_c, ctx_cap, target_x, n_td, tgt_slot, kv, kv_slots, z_tgt, z_ctx,
                              tckv, n_keys_tc, s2, n_rows_hint, st);
    }
    const int G = d.D / 8;
    ALINE_REQUIRE(n_tok * G <= ctx_max_threads(d.D), "context + target tokens per rollout (%d) exceed %d", n_tok,
                  ctx_max_threads(d.D) / G);
    const int NT = (n_tok + 7) / 8 * 8;
    int threads = (n_tok * G + 31) / 32 * 32;
    if (threads < 128) threads = 128;                 // more threads for the cooperative weight staging
    const int wf = (int)layer_w_floats(d, L);
    size_t smem = ((size_t)wf + 2 * (size_t)n_c * d.D + 2 * (size_t)d.D * NT) * sizeof(float);
    const int emit_fold = tckv != nullptr ? query_tc3_fold_keys(d, n_keys_tc) : 0;      // padded key count of the folded operands, 0: none
    if (d.D == 32) {
        if (set_smem(ctx_stack_kernel<32>, smem)) return 1;
        ctx_stack_kernel<32><<<B, threads, smem, st>>>(d, L, P, cx, cy, n_c, ctx_cap, target_x, n_td, tgt_slot, kv,
                                                        kv_slots, B, z_tgt, z_ctx, wf, NT, (unsigned char*)tckv, n_keys_tc, emit_fold);
    } else {
        if (set_smem(ctx_stack_kernel<64>, smem)) return 1;
        ctx_stack_kernel<64><<<B, threads, smem, st>>>(d, L, P, cx, cy, n_c, ctx_cap, target_x, n_td, tgt_slot, kv,
                                                        kv_slots, B, z_tgt, z_ctx, wf, NT, (unsigned char*)tckv, n_keys_tc, emit_fold);
    }
    ALINE_LAUNCH_OK();
    return 0;
}

static int query_stream(const Dims& d, const Layout& L, const float* P, const float* eq, const unsigned char* alive,
                        int B, int nq, const float* kv, int n_keys, int kv_slots, float t_value, float* logits,
                        float* zq, cudaStream_t st, const int* flag = nullptr, int epoch = 0) {
    // shared-memory words for the staged weights: the layer minus Wk / Wv, or the acquisition block
    const size_t w_layer = L.layer_stride - 2 * (size_t)d.D * d.D, w_acq = L.gmm0 - L.a_w1;
    const int wf = (int)pad4(w_layer > w_acq ? w_layer : w_acq);
    int NT = query_tile(d.D);
    auto need = [&](int nt) { return ((size_t)wf + 2 * (size_t)n_keys * d.D + 2 * (size_t)d.D * nt) * sizeof(float); };
    while (NT > 32 && need(NT) > (size_t)device_info().max_smem_optin) NT /= 2;
    const size_t smem = need(NT);
    dim3 grid(ceil_div(nq, NT), B);
    if (d.D == 32) {
        if (set_smem(query_stream_kernel<32>, smem)) return 1;
        query_stream_kernel<32><<<grid, NT, smem, st>>>(d, L, P, eq, alive, nq, kv, n_keys, kv_slots, B, t_value,
                                                                logits, zq, wf, flag, epoch);
    } else {
        if (set_smem(query_stream_kernel<64>, smem)) return 1;
        query_stream_kernel<64><<<grid, NT, smem, st>>>(d, L, P, eq, alive, nq, kv, n_keys, kv_slots, B, t_value,
                                                                logits, zq, wf, flag, epoch);
    }
    ALINE_LAUNCH_OK();
    return 0;
}

// csrc/query_tc.cu (general / robust tcgen05 kernel)
int query_stream_tc(const Dims& d, const Layout& L, const float* P, const void* wb, const float* eq,
                    const unsigned char* alive, int B, int nq, const float* kv, int n_keys, int kv_slots, float t_value,
                    float* logits, float* zq, const int* flag, int epoch, cudaStream_t st);
int query_stream_tc_max_keys(const Dims& d);
uint64_t query_tc_weight_bytes(const Dims& d);
// csrc/query_tc3.cu: the fast kernel (<= 48 keys) with P / relu(F) as tensor-memory A operands
bool query_tc3_supported(const Dims& d, int n_keys);
uint64_t query_tc3_weight_bytes(const Dims& d);
int query_stream_tc3(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq,
                     const unsigned char* alive, int B, int nq, int n_keys, float t_value, float* logits, float* zq,
                     const void* tckv, int* flag, int epoch, cudaStream_t st);

// csrc/query_tc4.cu: the fast kernel with two threads per candidate row (row-major embeddings)
bool query_tc4_supported(const Dims& d, int n_keys);
int query_stream_tc4(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq_rm,
                     const unsigned char* alive, int B, int nq, int n_keys, float t_value, float* logits, float* zq,
                     const void* tckv, int* flag, int epoch, cudaStream_t st);
// Which fast kernel: measured on B200 at the cfg2 launch shape (tools/bench_query.py, us per launch at 16 / 32 / 48 padded
// keys): one thread per row (query_tc3, 4 / 4 / 2 tiles in flight) 118 / 137 / 190; two threads per row (query_tc4, 3 / 3 /
// 2 tiles in flight) 129 / 143 / 183.  Default: query_tc4 above 32 keys.  ALINE_QUERY_TC4=1 / 0 forces it on / off.
void set_ces_fast_pow(int v);                           // csrc/spce.cu
// csrc/query_tc5.cu: d = 64 / 8 heads, weights streamed layer by layer (<= 48 keys)
bool query_tc5_supported(const Dims& d, int n_keys);
int query_tc5_kv_block_bytes(int nkp);
int query_stream_tc5(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq,
                     const unsigned char* alive, int B, int nq, int n_keys, float t_value, float* logits, float* zq,
                     const void* tckv, int* flag, int epoch, cudaStream_t st);
// does the model shape / key count have a fast kernel (bf16 operand blocks from ctx_stack)?
static bool query_fast_supported(const Dims& d, int n_keys) {
    return d.D == 64 ? query_tc5_supported(d, n_keys) : query_tc3_supported(d, n_keys);
}
void query_tc3_set_fold(int v);
void query_tc3_set_nq_hint(int nq, bool plain_needed); // csrc/query_tc3.cu: candidates per rollout of the launches that follow
static std::atomic<int> g_tc4_mode{-2};                // -2: not read yet; -1 auto; 0 off; 1 on (aline_set_option "query_tc4")
static bool tc4_wanted(const Dims& d, int n_keys) {
    int mode = g_tc4_mode.load(std::memory_order_relaxed);
    if (mode == -2) {
        const char* e = getenv("ALINE_QUERY_TC4");
        mode = e ? (e[0] == '0' ? 0 : 1) : -1;
        g_tc4_mode.store(mode, std::memory_order_relaxed);
    }
    // automatic: two threads per row above 32 keys unless the folded one-thread-per-row form covers the shape
    return mode < 0 ? (n_keys > 32 && !query_tc3_fold_emitted(d, n_keys)) : mode == 1;
}

// Overflow flags of the fast kernel: a ring of per-launch slots in device memory (one ring per device, allocated on
// first use, never freed); a launch owns slot (epoch % kFlagSlots) and stores its epoch there when a softmax row
// overflowed -- no reset needed, and launches in flight on different streams do not disturb each other.
constexpr int kFlagSlots = 256;
static int tc_flag_slot(int** flag, int* epoch) {
    static std::atomic<int> counter{1};
    static thread_local int* ring[64] = {};
    int dev = 0;
    ALINE_CHECK_CUDA(cudaGetDevice(&dev));
    ALINE_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    if (!ring[dev]) {
        static std::mutex mu;
        static int* shared_ring[64] = {};
        std::lock_guard<std::mutex> lk(mu);
        if (!shared_ring[dev]) {
            ALINE_CHECK_CUDA(cudaMalloc(&shared_ring[dev], kFlagSlots * sizeof(int)));
            ALINE_CHECK_CUDA(cudaMemset(shared_ring[dev], 0, kFlagSlots * sizeof(int)));
        }
        ring[dev] = shared_ring[dev];
    }
    const int e = counter.fetch_add(1, std::memory_order_relaxed) & 0x3fffffff;
    *epoch = e ? e : 1;
    *flag = ring[dev] + (e % kFlagSlots);
    return 0;
}

// tensor-core query stream: the fast kernel when the shape has one and its operand blocks are given (followed by
// the conditional robust launch), else the general kernel
static int query_stream_tc_any(const Dims& d, const Layout& L, const float* P, const void* wb, const float* eq,
                               const float* eq_rm, const unsigned char* alive, int B, int nq, const float* kv, int n_keys,
                               int kv_slots, float t_value, float* logits, float* zq, const void* tckv, cudaStream_t st) {
    const bool use4 = tckv && eq_rm && tc4_wanted(d, n_keys) && query_tc4_supported(d, n_keys);
    ALINE_REQUIRE(eq || use4, "tensor-core query stream: the k-major embeddings eq are required for this shape");
    // the general kernel (d = 32) holds fp32 K / V in shared memory
    const bool general_ok = d.D == 32 && n_keys <= query_stream_tc_max_keys(d);
    if (tckv && query_fast_supported(d, n_keys)) {
        int* flag = nullptr;
        int epoch = 0;
        if (tc_flag_slot(&flag, &epoch)) return 1;
        if (d.D == 64) {                               // the bf16 blob of a d = 64 model has only the fast section
            if (query_stream_tc5(d, L, P, wb, eq, alive, B, nq, n_keys, t_value, logits, zq, tckv, flag, epoch, st)) return 1;
            return query_stream(d, L, P, eq, alive, B, nq, kv, n_keys, kv_slots, t_value, logits, zq, st, flag, epoch);
        }
        const unsigned char* wb2 = (const unsigned char*)wb + query_tc_weight_bytes(d);
        if (use4) {
            if (query_stream_tc4(d, L, P, wb2, eq_rm, alive, B, nq, n_keys, t_value, logits, zq, tckv, flag, epoch, st)) return 1;
            if (!eq) return 0;                         // no k-major embeddings: the conditional robust launch is not possible
        } else if (query_stream_tc3(d, L, P, wb2, eq, alive, B, nq, n_keys, t_value, logits, zq, tckv, flag, epoch, st)) {
            return 1;
        }
        // conditional robust recomputation (runs only if the fast kernel flagged an overflowing softmax row): the general
        // tcgen05 kernel while its fp32 K / V fit in shared memory, the FFMA kernel beyond
        if (general_ok)
            return query_stream_tc(d, L, P, wb, eq, alive, B, nq, kv, n_keys, kv_slots, t_value, logits, zq, flag, epoch, st);
        return query_stream(d, L, P, eq, alive, B, nq, kv, n_keys, kv_slots, t_value, logits, zq, st, flag, epoch);
    }
    if (!general_ok)        // no operand blocks for the fast kernel and too many keys for the general one
        return query_stream(d, L, P, eq, alive, B, nq, kv, n_keys, kv_slots, t_value, logits, zq, st);
    return query_stream_tc(d, L, P, wb, eq, alive, B, nq, kv, n_keys, kv_slots, t_value, logits, zq, nullptr, 0, st);
}

}  // namespace aline

using namespace aline;

extern "C" {

int aline_set_option(const char* name, int32_t value) {
    ALINE_REQUIRE(name != nullptr, "aline_set_option: NULL name");
    if (std::string(name) == "query_tc4") {
        ALINE_REQUIRE(value >= -1 && value <= 1, "aline_set_option(query_tc4): value must be -1 (auto), 0 or 1");
        g_tc4_mode.store(value, std::memory_order_relaxed);
        return 0;
    }
    if (std::string(name) == "query_fold") {
        ALINE_REQUIRE(value >= -1 && value <= 2, "aline_set_option(query_fold): value must be -1 (auto), 0, 1 or 2 (<= 32 keys only)");
        query_tc3_set_fold(value);
        return 0;
    }
    if (std::string(name) == "ces_fast_pow") {
        set_ces_fast_pow(value);
        return 0;
    }
    return set_error("aline_set_option: unknown option '%s'", name);
}

uint64_t aline_model_param_count(const aline_model* m) {
    if (!m) return 0;
    Dims d;
    d.D = m->d; d.FF = m->ff; d.H = m->n_head; d.NL = m->n_layer; d.dx = m->dim_x; d.dy = m->dim_y;
    d.ntok = m->n_theta_tok; d.C = m->n_comp; d.EH = m->emb_hidden; d.HH = m->head_hidden; d.tt = m->time_token ? 1 : 0;
    d.std_min = m->std_min;
    return make_layout(d).total;
}

int aline_embed_queries(const aline_model* m, const float* query_x, int32_t B, int32_t nq, float* eq, void* stream) {
    Dims d;
    if (dims_from(m, d)) return 1;
    ALINE_REQUIRE(query_x && eq && B >= 1 && nq >= 1, "aline_embed_queries: bad arguments");
    return embed_queries(d, make_layout(d), m->params, query_x, B, nq, eq, nullptr, (cudaStream_t)stream);
}

int aline_embed_queries_ex(const aline_model* m, const float* query_x, int32_t B, int32_t nq, float* eq, float* eq_rm,
                           void* stream) {
    Dims d;
    if (dims_from(m, d)) return 1;
    ALINE_REQUIRE(query_x && (eq || eq_rm) && B >= 1 && nq >= 1, "aline_embed_queries_ex: bad arguments");
    return embed_queries(d, make_layout(d), m->params, query_x, B, nq, eq, eq_rm, (cudaStream_t)stream);
}

int aline_ctx_stack_ex(const aline_model* m, const float* cx, const float* cy, int32_t B, int32_t n_c, int32_t ctx_cap,
                    const float* target_x, int32_t n_td, const int32_t* tgt_slot, float* kv, int32_t kv_slots,
                    float* z_tgt, float* z_ctx, void* tckv, int32_t n_keys_tc, void* stream) {
    Dims d;
    if (dims_from(m, d)) return 1;
    ALINE_REQUIRE(cx && cy && kv && tgt_slot && B >= 1, "aline_ctx_stack: NULL tensor");
    ALINE_REQUIRE(n_c >= 1 && n_c <= ctx_cap, "aline_ctx_stack: n_context %d must be in 1..%d (an empty context leaves "
                  "every attention row fully masked)", n_c, ctx_cap);
    ALINE_REQUIRE(n_td == 0 || target_x, "aline_ctx_stack: target_x required for %d data targets", n_td);
    ALINE_REQUIRE(kv_slots >= n_c, "aline_ctx_stack: kv_slots %d < n_context %d", kv_slots, n_c);
    return ctx_stack(d, make_layout(d), m->params, cx, cy, B, n_c, ctx_cap, target_x, n_td, tgt_slot, kv, kv_slots, z_tgt,
                     z_ctx, tckv, n_keys_tc, (cudaStream_t)stream);
}

int aline_ctx_stack(const aline_model* m, const float* cx, const float* cy, int32_t B, int32_t n_c, int32_t ctx_cap,
                    const float* target_x, int32_t n_td, const int32_t* tgt_slot, float* kv, int32_t kv_slots,
                    float* z_tgt, void* tckv, int32_t n_keys_tc, void* stream) {
    return aline_ctx_stack_ex(m, cx, cy, B, n_c, ctx_cap, target_x, n_td, tgt_slot, kv, kv_slots, z_tgt, nullptr, tckv,
                              n_keys_tc, stream);
}

int aline_query_stream(const aline_model* m, const float* eq, const uint8_t* alive, int32_t B, int32_t nq,
                       const float* kv, int32_t n_keys, int32_t kv_slots, float t_value, float* logits, float* zq,
                       void* stream) {
    Dims d;
    if (dims_from(m, d)) return 1;
    ALINE_REQUIRE(eq && kv && logits && B >= 1 && nq >= 1 && n_keys >= 1 && n_keys <= kv_slots,
                  "aline_query_stream: bad arguments");
    return query_stream(d, make_layout(d), m->params, eq, alive, B, nq, kv, n_keys, kv_slots, t_value, logits, zq,
                        (cudaStream_t)stream);
}

static void dims_unchecked(const aline_model* m, Dims& d) {
    d.D = m->d; d.FF = m->ff; d.H = m->n_head; d.NL = m->n_layer; d.dx = m->dim_x; d.dy = m->dim_y;
    d.ntok = m->n_theta_tok; d.C = m->n_comp; d.EH = m->emb_hidden; d.HH = m->head_hidden; d.tt = m->time_token ? 1 : 0;
    d.std_min = m->std_min;
}

uint64_t aline_tc_weight_bytes(const aline_model* m) {
    if (!m) return 0;
    Dims d;
    dims_unchecked(m, d);
    if (d.D == 64) return query_tc5_supported(d, 1) ? query_tc3_weight_bytes(d) : 0;       // fast section only
    return query_tc_weight_bytes(d) + query_tc3_weight_bytes(d);
}

uint64_t aline_tc_kv_bytes(const aline_model* m, int32_t B, int32_t n_keys) {
    if (!m || B < 1 || n_keys < 1) return 0;
    // per key: (heads + 1) x 16 B of K operand rows + heads x 32 B of V operand columns, heads = d / 8; d = 32: + the
    // folded operands of csrc/query_tc3.cu (640 B per key, query_fast.cuh)
    const uint64_t per_key = (uint64_t)(6 * m->d + 16) + (m->d == 32 ? (uint64_t)tcq::kFoldKeyBytes : 0);
    return (uint64_t)m->n_layer * (uint64_t)B * per_key * (uint64_t)((n_keys + 15) / 16 * 16);
}

int32_t aline_tc_fast_max_keys(const aline_model* m) {
    if (!m) return 0;
    Dims d;
    dims_unchecked(m, d);
    int best = 0;
    for (int k = 16; k <= 160; k += 16)
        if (query_fast_supported(d, k)) best = k;
    return best;
}

int32_t aline_tc_max_keys(const aline_model* m) {
    Dims d;
    if (!m || m->d != 32) return 0;
    d.D = m->d; d.FF = m->ff; d.H = m->n_head; d.NL = m->n_layer; d.dx = m->dim_x; d.dy = m->dim_y;
    d.ntok = m->n_theta_tok; d.C = m->n_comp; d.EH = m->emb_hidden; d.HH = m->head_hidden; d.tt = m->time_token ? 1 : 0;
    d.std_min = m->std_min;
    if (d.FF > 128 || d.HH > 128) return 0;
    return query_stream_tc_max_keys(d);
}

int aline_query_stream_tc_ex(const aline_model* m, const void* tc_weights, const float* eq, const float* eq_rm,
                             const uint8_t* alive, int32_t B, int32_t nq, const float* kv, int32_t n_keys,
                             int32_t kv_slots, float t_value, float* logits, float* zq, const void* tckv, void* stream) {
    Dims d;
    if (dims_from(m, d)) return 1;
    ALINE_REQUIRE(tc_weights && (eq || eq_rm) && kv && logits && B >= 1 && nq >= 1 && n_keys >= 1 && n_keys <= kv_slots,
                  "aline_query_stream_tc: bad arguments");
    return query_stream_tc_any(d, make_layout(d), m->params, tc_weights, eq, eq_rm, alive, B, nq, kv, n_keys, kv_slots,
                               t_value, logits, zq, tckv, (cudaStream_t)stream);
}

int aline_query_stream_tc(const aline_model* m, const void* tc_weights, const float* eq, const uint8_t* alive, int32_t B,
                          int32_t nq, const float* kv, int32_t n_keys, int32_t kv_slots, float t_value, float* logits,
                          float* zq, const void* tckv, void* stream) {
    ALINE_REQUIRE(eq, "aline_query_stream_tc: eq is NULL");
    return aline_query_stream_tc_ex(m, tc_weights, eq, nullptr, alive, B, nq, kv, n_keys, kv_slots, t_value, logits, zq,
                                    tckv, stream);
}

int aline_select(const float* logits, uint8_t* alive, int32_t B, int32_t nq, const float* qx, const float* qy,
                 int32_t dx, int32_t dy, float* cx, float* cy, int32_t n_c, int32_t ctx_cap, int64_t* idx_out,
                 int32_t idx_stride, float* logp_out, int32_t logp_stride, int64_t* idx_orig_out, float* zt,
                 void* stream) {
    ALINE_REQUIRE(logits && idx_out && logp_out && B >= 1 && nq >= 1, "aline_select: bad arguments");
    ALINE_REQUIRE(!cx || (qx && qy && cy && n_c < ctx_cap), "aline_select: append needs qx, qy, cy and n_c < ctx_cap");
    ALINE_REQUIRE(!(zt && alive), "aline_select: zt output requires a fully live candidate set (alive = NULL)");
    select_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(logits, alive, nq, qx, qy, dx, dy, cx, cy, n_c, ctx_cap,
                                                        (long long*)idx_out, idx_stride, logp_out, logp_stride,
                                                        (long long*)idx_orig_out, zt);
    ALINE_LAUNCH_OK();
    return 0;
}

int aline_select_sample(const float* logits, uint8_t* alive, int32_t B, int32_t nq, const float* qx, const float* qy,
                        int32_t dx, int32_t dy, float* cx, float* cy, int32_t n_c, int32_t ctx_cap, int64_t* idx_out,
                        int32_t idx_stride, float* logp_out, int32_t logp_stride, int64_t* idx_orig_out, float* zt,
                        uint64_t seed, int32_t step, void* stream) {
    ALINE_REQUIRE(logits && idx_out && logp_out && B >= 1 && nq >= 1, "aline_select_sample: bad arguments");
    ALINE_REQUIRE(!cx || (qx && qy && cy && n_c < ctx_cap), "aline_select_sample: append needs qx, qy, cy and n_c < ctx_cap");
    ALINE_REQUIRE(!(zt && alive), "aline_select_sample: zt output requires a fully live candidate set (alive = NULL)");
    select_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(logits, alive, nq, qx, qy, dx, dy, cx, cy, n_c, ctx_cap,
                                                        (long long*)idx_out, idx_stride, logp_out, logp_stride,
                                                        (long long*)idx_orig_out, zt, 1, (unsigned long long)seed, step);
    ALINE_LAUNCH_OK();
    return 0;
}

static int gmm_head_launch(const aline_model* m, const float* z, int64_t n_tok, float* means, float* stds,
                           float* weights, float* var_out, void* stream) {
    Dims d;
    if (dims_from(m, d)) return 1;
    ALINE_REQUIRE(z && n_tok >= 1 && (var_out || (means && stds && weights)), "aline_gmm_head: bad arguments");
    Layout L = make_layout(d);
    const int NT = 128;
    size_t smem = (L.gmm_stride + (size_t)d.D * NT + 3 * (size_t)d.C * NT) * sizeof(float);
    int blocks = (int)ceil_div64(n_tok, NT);
    if (d.D == 32) {
        if (set_smem(gmm_head_kernel<32>, smem)) return 1;
        gmm_head_kernel<32><<<blocks, NT, smem, (cudaStream_t)stream>>>(d, L, m->params, z, n_tok, means, stds, weights,
                                                                        var_out);
    } else {
        if (set_smem(gmm_head_kernel<64>, smem)) return 1;
        gmm_head_kernel<64><<<blocks, NT, smem, (cudaStream_t)stream>>>(d, L, m->params, z, n_tok, means, stds, weights,
                                                                        var_out);
    }
    ALINE_LAUNCH_OK();
    return 0;
}

int aline_gmm_head(const aline_model* m, const float* z, int64_t n_tok, float* means, float* stds, float* weights,
                   void* stream) {
    ALINE_REQUIRE(means && stds && weights, "aline_gmm_head: NULL output");
    return gmm_head_launch(m, z, n_tok, means, stds, weights, nullptr, stream);
}

int aline_gmm_head_variance(const aline_model* m, const float* z, int64_t n_tok, float* variance, void* stream) {
    ALINE_REQUIRE(variance, "aline_gmm_head_variance: NULL output");
    return gmm_head_launch(m, z, n_tok, nullptr, nullptr, nullptr, variance, stream);
}

int aline_gmm_variance(const float* means, const float* stds, const float* weights, int64_t n, int32_t C,
                       int64_t tok_per_w, float* out, void* stream) {
    ALINE_REQUIRE(means && stds && weights && out && n >= 1 && C >= 1 && tok_per_w >= 1, "aline_gmm_variance: bad arguments");
    gmm_variance_kernel<<<(int)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(means, stds, weights, n, C, tok_per_w, out);
    ALINE_LAUNCH_OK();
    return 0;
}

int aline_value_head(const float* z_ctx, int32_t B, int32_t n_c, int32_t d, int32_t ff, const float* w1, const float* b1,
                     const float* w2, const float* b2, float* value, void* stream) {
    ALINE_REQUIRE(z_ctx && w1 && b1 && w2 && b2 && value && B >= 1 && d >= 1 && ff >= 1, "aline_value_head: bad arguments");
    ALINE_REQUIRE(n_c >= 1, "aline_value_head: empty context (the reference returns head.value_head.empty_value there; "
                  "the encoder itself needs n_context >= 1)");
    const size_t smem = (size_t)n_c * d * sizeof(float);
    if (set_smem(value_head_kernel, smem)) return 1;
    value_head_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(z_ctx, n_c, d, ff, w1, b1, w2, b2, value);
    ALINE_LAUNCH_OK();
    return 0;
}

int aline_gmm_log_likelihood(const float* value, const float* means, const float* stds, const float* weights, int64_t n,
                             int32_t C, float* out, void* stream) {
    ALINE_REQUIRE(value && means && stds && weights && out && n >= 1 && C >= 1, "aline_gmm_log_likelihood: bad arguments");
    gmm_loglik_kernel<<<(int)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(value, means, stds, weights, n, C, out);
    ALINE_LAUNCH_OK();
    return 0;
}

int aline_move_selected(const float* query, const float* ctx, const int64_t* idx, int32_t B, int32_t N, int32_t M,
                        int32_t D, float* new_query, float* new_ctx, void* stream) {
    ALINE_REQUIRE(query && ctx && idx && new_query && new_ctx && B >= 1 && N >= 1 && M >= 0 && D >= 1,
                  "aline_move_selected: bad arguments");
    move_selected_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(query, ctx, (const long long*)idx, N, M, D, new_query,
                                                               new_ctx);
    ALINE_LAUNCH_OK();
    return 0;
}

int aline_rollout(const aline_model* m, const float* qx, const float* qy, uint8_t* alive, const float* eq, float* cx,
                  float* cy, int32_t B, int32_t nq, int32_t n_c0, int32_t ctx_cap, const float* target_x, int32_t n_td,
                  const int32_t* tgt_slot, int32_t n_sel, float* kv, int32_t kv_slots, float* logits, int32_t T,
                  const float* t_values_host, int64_t* idx_hist, float* logp_hist, const void* tc_weights, void* tckv,
                  void* stream) {
    ALINE_REQUIRE(eq, "aline_rollout: eq is NULL");
    return aline_rollout_ex(m, qx, qy, alive, eq, nullptr, cx, cy, B, nq, n_c0, ctx_cap, target_x, n_td, tgt_slot, n_sel, kv,
                            kv_slots, logits, T, t_values_host, idx_hist, logp_hist, tc_weights, tckv, stream);
}

int aline_rollout_ex(const aline_model* m, const float* qx, const float* qy, uint8_t* alive, const float* eq,
                     const float* eq_rm, float* cx, float* cy, int32_t B, int32_t nq, int32_t n_c0, int32_t ctx_cap,
                     const float* target_x, int32_t n_td, const int32_t* tgt_slot, int32_t n_sel, float* kv,
                     int32_t kv_slots, float* logits, int32_t T, const float* t_values_host, int64_t* idx_hist,
                     float* logp_hist, const void* tc_weights, void* tckv, void* stream) {
    Dims d;
    if (dims_from(m, d)) return 1;
    ALINE_REQUIRE(qx && qy && alive && (eq || eq_rm) && cx && cy && tgt_slot && kv && logits && idx_hist && logp_hist,
                  "aline_rollout: NULL tensor");
    ALINE_REQUIRE(eq || (tc_weights && tckv && eq_rm), "aline_rollout: eq is required without the fast tensor-core path");
    ALINE_REQUIRE(T >= 1 && n_c0 >= 1 && n_c0 + T <= ctx_cap && T <= nq, "aline_rollout: need 1 <= T <= n_query and "
                  "n_context_init + T <= ctx_cap (T=%d n_c0=%d cap=%d nq=%d)", T, n_c0, ctx_cap, nq);
    ALINE_REQUIRE(kv_slots >= n_c0 + T - 1 + n_sel, "aline_rollout: kv_slots %d too small", kv_slots);
    Layout L = make_layout(d);
    cudaStream_t st = (cudaStream_t)stream;
    // Step t: [select of step t-1 fused into] ctx_stack -> query stream -> (last step, or no fused kernel) select.
    bool pending = false;                              // step t-1's logits are written, its design not chosen yet
    struct ChainGuard { ~ChainGuard() { g_pdl_chain = false; query_tc3_set_nq_hint(0, true); } } chain_guard;   // any return path leaves the chain mode
    // the context kernel emits the folded operands only if this nq uses them, and then only those unless the
    // two-threads-per-row kernel (which reads the plain blocks) is forced
    query_tc3_set_nq_hint(nq, tc4_wanted(d, 1));
    for (int t = 0; t < T; ++t) {
        // from the second launch on, the preceding launch of the stream is this rollout's own: the kernels of the chain
        // may start as programmatic dependents (common.cuh) and overlap their prologues with the predecessor's tail
        g_pdl_chain = t > 0;
        const int n_c = n_c0 + t;
        const int n_keys = n_c + n_sel;
        const bool fast = tc_weights && tckv && query_fast_supported(d, n_keys);
        SelectArgs sel{logits, alive, nq, qx, qy, d.dx, d.dy, cx, cy, n_c - 1, ctx_cap, (long long*)idx_hist + (t - 1), T,
                       logp_hist + (t - 1), T, nullptr, nullptr};
        bool fused = false;
        if (pending) {
            // try the fused kernel first; if the shape has none, run the stand-alone select, then the context stack
            if (!ctx_fuses_select(d, L, m->params, n_c, n_c + n_td + d.ntok, kv_slots)) {
                select_kernel<<<B, 256, 0, st>>>(logits, alive, nq, qx, qy, d.dx, d.dy, cx, cy, n_c - 1, ctx_cap,
                                                 (long long*)idx_hist + (t - 1), T, logp_hist + (t - 1), T, nullptr, nullptr);
                ALINE_LAUNCH_OK();
                pending = false;
            }
        }
        if (ctx_stack(d, L, m->params, cx, cy, B, n_c, ctx_cap, target_x, n_td, tgt_slot, kv, kv_slots, nullptr, nullptr,
                      fast ? tckv : nullptr, fast ? n_keys : 0, st, pending ? &sel : nullptr, &fused, n_c + n_sel))
            return 1;
        if (pending && !fused) return set_error("aline_rollout: internal error (design step %d was not selected)", t - 1);
        g_pdl_chain = true;
        float tv = t_values_host ? t_values_host[t] : 0.f;
        if (tc_weights) {
            if (query_stream_tc_any(d, L, m->params, tc_weights, eq, eq_rm, alive, B, nq, kv, n_keys, kv_slots, tv, logits,
                                    nullptr, fast ? tckv : nullptr, st)) return 1;
        } else if (query_stream(d, L, m->params, eq, alive, B, nq, kv, n_keys, kv_slots, tv, logits, nullptr, st)) {
            return 1;
        }
        pending = true;
    }
    // the last step's design
    g_pdl_chain = false;
    select_kernel<<<B, 256, 0, st>>>(logits, alive, nq, qx, qy, d.dx, d.dy, cx, cy, n_c0 + T - 1, ctx_cap,
                                     (long long*)idx_hist + (T - 1), T, logp_hist + (T - 1), T, nullptr, nullptr);
    ALINE_LAUNCH_OK();
    return 0;
}

}  // extern "C"
