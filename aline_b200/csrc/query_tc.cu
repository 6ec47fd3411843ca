// Candidate-query stream on the 5th-generation tensor cores (bf16 operands, fp32 accumulation in TMEM): the
// general / robust variant (any key count that fits in shared memory, max-subtracted softmax).  The fast variant
// for <= 48 keys is csrc/query_tc3.cu; this kernel is also its fallback: launched right after it with the same
// arguments plus (flag, epoch), it returns immediately unless the fast kernel flagged an overflowing softmax row.
//
// Same contract as query_stream_kernel (csrc/rollout.cu; reference: model/encoder.py:128-141 restricted to the
// query rows + model/head.py:27-31), d = 32.  The dense contractions of every layer -- Q projection, attention
// out-projection, both MLP projections -- and the first layer of the acquisition MLP are tcgen05.mma instructions
// on 128-token tiles; bias / scale / softmax / LayerNorm / ReLU / residual epilogues run in fp32 on the accumulator
// rows read back from TMEM (one thread = one token = one TMEM lane).  The attention itself (<= ~80 keys of 8-wide
// heads) stays on the FFMA pipe with fp32 K, V.
//
// One persistent CTA per SM, two warpgroups.  A work unit is (rollout b, pair of 128-token tiles): each warpgroup
// owns one tile, its own operand staging buffers, TMEM columns and mbarrier, so while one warpgroup runs an
// epilogue the other one's MMAs occupy the tensor pipe.  Shared memory:
//   bf16 weights of all layers + acquisition W1, core-matrix tiled (csrc/tc.cuh), one TMA bulk copy per CTA  68 KB
//   fp32 bias / LayerNorm / head vectors                                                                       6 KB
//   fp32 K, V of the unit's rollout for all layers, TMA bulk copies                                   n_keys*768 B
//   per warpgroup: A operand buffers  X / O / H [128x32] bf16 (8 KB) and F1 [128x128] bf16 (32 KB)
#include "model.cuh"
#include "tc.cuh"

namespace aline {

constexpr int kTcD = 32;
constexpr int kTcTile = 128;

struct TcShape {
    int FF, HH, NL, n_keys;
    // byte offsets inside the bf16 weight blob
    int layer_bytes, off_wq, off_wo, off_w1, off_w2, off_acq, total_bytes;
    // float offsets inside the fp32 vector block
    int vec_layer, v_bq, v_bo, v_g1, v_be1, v_b1, v_b2, v_g2, v_be2, v_acq_b1, v_acq_w2, v_acq_wt, v_acq_b2, vec_total;
};

__host__ __device__ inline TcShape make_tc_shape(const Dims& m, int n_keys) {
    TcShape s;
    const int D = kTcD;
    s.FF = m.FF; s.HH = m.HH; s.NL = m.NL; s.n_keys = n_keys;
    s.off_wq = 0;
    s.off_wo = s.off_wq + D * D * 2;
    s.off_w1 = s.off_wo + D * D * 2;
    s.off_w2 = s.off_w1 + m.FF * D * 2;
    s.layer_bytes = s.off_w2 + D * m.FF * 2;
    s.off_acq = s.layer_bytes * m.NL;
    s.total_bytes = s.off_acq + m.HH * D * 2;
    int o = 0;
    s.v_bq = o; o += D;
    s.v_bo = o; o += D;
    s.v_g1 = o; o += D;
    s.v_be1 = o; o += D;
    s.v_b1 = o; o += m.FF;
    s.v_b2 = o; o += D;
    s.v_g2 = o; o += D;
    s.v_be2 = o; o += D;
    s.vec_layer = o;
    o = s.vec_layer * m.NL;
    s.v_acq_b1 = o; o += m.HH;
    s.v_acq_w2 = o; o += m.HH;
    s.v_acq_wt = o; o += m.HH;
    s.v_acq_b2 = o; o += 4;
    s.vec_total = o;
    return s;
}

// softmax(q K^T) V for one token; K row of key j at kv[j*2D], V row at kv[j*2D + D]; q already scaled
__device__ __forceinline__ void attention_row(const float (&q)[kTcD], const float* kv, int n_keys, float (&o)[kTcD]) {
    constexpr int D = kTcD, H = D / 8;
    float mx[H];
#pragma unroll
    for (int h = 0; h < H; ++h) mx[h] = -INFINITY;
    for (int j = 0; j < n_keys; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(kv + (size_t)j * 2 * D);
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
    float den[H];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) den[h] = 0.f;
    for (int j = 0; j < n_keys; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(kv + (size_t)j * 2 * D);
        const float4* vr = kr + D / 4;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float4 a = kr[2 * h], b = kr[2 * h + 1];
            float s = q[8 * h] * a.x;
            s = fmaf(q[8 * h + 1], a.y, s); s = fmaf(q[8 * h + 2], a.z, s); s = fmaf(q[8 * h + 3], a.w, s);
            s = fmaf(q[8 * h + 4], b.x, s); s = fmaf(q[8 * h + 5], b.y, s); s = fmaf(q[8 * h + 6], b.z, s);
            s = fmaf(q[8 * h + 7], b.w, s);
            const float p = exp_fast(s - mx[h]);
            den[h] += p;
            float4 va = vr[2 * h], vb = vr[2 * h + 1];
            o[8 * h + 0] = fmaf(p, va.x, o[8 * h + 0]); o[8 * h + 1] = fmaf(p, va.y, o[8 * h + 1]);
            o[8 * h + 2] = fmaf(p, va.z, o[8 * h + 2]); o[8 * h + 3] = fmaf(p, va.w, o[8 * h + 3]);
            o[8 * h + 4] = fmaf(p, vb.x, o[8 * h + 4]); o[8 * h + 5] = fmaf(p, vb.y, o[8 * h + 5]);
            o[8 * h + 6] = fmaf(p, vb.z, o[8 * h + 6]); o[8 * h + 7] = fmaf(p, vb.w, o[8 * h + 7]);
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const float inv = 1.0f / den[h];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[8 * h + i] *= inv;
    }
}

__global__ void __launch_bounds__(256, 1)
query_stream_tc_kernel(const Dims m, const Layout L, const TcShape S, const float* __restrict__ P,
                       const unsigned char* __restrict__ Wb_g, const float* __restrict__ eq,
                       const unsigned char* __restrict__ alive, int nq, const float* __restrict__ kv, int kv_slots, int B,
                       float t_value, float* __restrict__ logits, float* __restrict__ zq, int n_units, int pairs_per_b,
                       const int* __restrict__ flag, int epoch) {
    extern __shared__ __align__(1024) unsigned char smem[];
    pdl_trigger();
    pdl_wait();                                              // (programmatic dependent of the fast kernel: its flag, its logits)
    if (flag != nullptr && *flag != epoch) return;           // fallback launch and the fast kernel was fine
    constexpr int D = kTcD;
    __shared__ __align__(8) uint64_t bar_w, bar_kv, bar_mma[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, r = tid & 127;
    // ---- carve shared memory ----
    unsigned char* Wb = smem;                                                 // bf16 weights
    float* Vec = reinterpret_cast<float*>(Wb + ((S.total_bytes + 127) & ~127));
    float* KV = Vec + ((S.vec_total + 31) & ~31);                             // [NL][n_keys][2][D] fp32 (FFMA attention)
    unsigned char* Abase = reinterpret_cast<unsigned char*>(KV + (size_t)S.NL * S.n_keys * 2 * D);
    Abase = reinterpret_cast<unsigned char*>(((uintptr_t)Abase + 127) & ~(uintptr_t)127);
    const int a_cols = S.FF > S.HH ? S.FF : S.HH;
    const int a_x_bytes = kTcTile * D * 2, a_f_bytes = kTcTile * a_cols * 2;
    unsigned char* Ax = Abase + (size_t)wg * (a_x_bytes + a_f_bytes);         // X / H operand  [128 x 32]
    unsigned char* Af = Ax + a_x_bytes;                                       // O (first 8 KB) / F1 operand [128 x FF]

    if (tid == 0) {
        tc::mbar_init(&bar_w, 1);
        tc::mbar_init(&bar_kv, 1);
        tc::mbar_init(&bar_mma[0], 1);
        tc::mbar_init(&bar_mma[1], 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s + (uint32_t)wg * 256;
    const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
    const uint32_t tQ = tmem + lane_off, tF = tmem + 32 + lane_off;           // 32 + 128 columns per warpgroup

    // weights: one TMA bulk copy for the whole CTA; fp32 vectors gathered from the parameter blob
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(&bar_w, (uint32_t)S.total_bytes);
        tc::bulk_g2s(Wb, Wb_g, (uint32_t)S.total_bytes, &bar_w);
    }
    for (int l = 0; l < S.NL; ++l) {
        const float* Pl = P + L.layer0 + (size_t)l * L.layer_stride;
        float* V = Vec + l * S.vec_layer;
        for (int i = tid; i < D; i += 256) {
            V[S.v_bq + i] = Pl[L.bq + i]; V[S.v_bo + i] = Pl[L.bo + i]; V[S.v_g1 + i] = Pl[L.g1 + i];
            V[S.v_be1 + i] = Pl[L.be1 + i]; V[S.v_b2 + i] = Pl[L.b2 + i]; V[S.v_g2 + i] = Pl[L.g2 + i];
            V[S.v_be2 + i] = Pl[L.be2 + i];
        }
        for (int i = tid; i < S.FF; i += 256) V[S.v_b1 + i] = Pl[L.b1 + i];
    }
    for (int i = tid; i < S.HH; i += 256) {
        Vec[S.v_acq_b1 + i] = P[L.a_b1 + i];
        Vec[S.v_acq_w2 + i] = P[L.a_w2 + i];
        Vec[S.v_acq_wt + i] = m.tt ? P[L.a_w1 + (size_t)D * S.HH + i] : 0.f;
    }
    if (tid == 0) Vec[S.v_acq_b2] = P[L.a_b2];
    tc::mbar_wait(&bar_w, 0);
    __syncthreads();

    const uint32_t wb_s = tc::smem_u32(Wb), ax_s = tc::smem_u32(Ax), af_s = tc::smem_u32(Af);
    uint32_t ph_mma = 0, ph_kv = 0;
    const uint32_t kv_layer_bytes = (uint32_t)S.n_keys * 2 * D * sizeof(float);

    // one MMA phase of this warpgroup: publish the operand stores, let one thread issue, wait for completion
    auto mma_phase = [&](auto&& issue) {
        tc::fence_async_smem();                      // this thread's operand stores -> async proxy
        tc::tc_fence_before();
        tc::named_sync(1 + wg, 128);
        if (r == 0) {
            tc::tc_fence_after();
            issue();
            tc::umma_commit(&bar_mma[wg]);
        }
        tc::mbar_wait(&bar_mma[wg], ph_mma);
        ph_mma ^= 1;
        tc::tc_fence_after();
    };
    // D[128 x N] = A[128 x Kd] * W[N x Kd]^T
    auto gemm = [&](uint32_t d_tmem_cols, uint32_t a_s, uint32_t w_s, int N, int Kd) {
        mma_phase([&] { tc::umma_gemm(d_tmem_cols, a_s, kTcTile, w_s, N, Kd, tc::idesc_bf16(128, N)); });
    };

    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int b = unit / pairs_per_b, pair = unit - b * pairs_per_b;
        // ---- K, V of rollout b for all layers: TMA bulk copies ----
        __syncthreads();                                                       // everyone is done with the previous K, V
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&bar_kv, kv_layer_bytes * S.NL);
            for (int l = 0; l < S.NL; ++l)
                tc::bulk_g2s(KV + (size_t)l * S.n_keys * 2 * D, kv + ((size_t)l * B + b) * kv_slots * (2 * D),
                             kv_layer_bytes, &bar_kv);
        }
        const int j = (2 * pair + wg) * kTcTile + r;
        const bool in_range = j < nq;
        const bool live = in_range && (alive == nullptr || alive[(size_t)b * nq + j] != 0);
        float x[D];
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = live ? __ldg(eq + ((size_t)b * D + i) * nq + j) : 0.f;
        tc::mbar_wait(&bar_kv, ph_kv);
        ph_kv ^= 1;

        for (int l = 0; l < S.NL; ++l) {
            const float* V = Vec + l * S.vec_layer;
            const uint32_t wl = wb_s + (uint32_t)l * S.layer_bytes;
            const float* kvl = KV + (size_t)l * S.n_keys * 2 * D;
            // Q = x Wq^T
            tc::store_row_bf16<D>(Ax, kTcTile, r, x);
            gemm(tmem, ax_s, wl + S.off_wq, D, D);
            float q[D], o[D];
            tc::tmem_ld32(tQ, q);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < D; ++i) q[i] = (q[i] + V[S.v_bq + i]) * 0.35355339059327376220f;
            attention_row(q, kvl, S.n_keys, o);
            tc::store_row_bf16<D>(Af, kTcTile, r, o);
            // y = o Wo^T ; h = LN1(x + y + bo)
            gemm(tmem, af_s, wl + S.off_wo, D, D);
            tc::tmem_ld32(tQ, q);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < D; ++i) x[i] += q[i] + V[S.v_bo + i];
            layer_norm<D>(x, V + S.v_g1, V + S.v_be1);
            // f1 = relu(h W1^T + b1)
            tc::store_row_bf16<D>(Ax, kTcTile, r, x);
            gemm(tmem + 32, ax_s, wl + S.off_w1, S.FF, D);
            for (int c0 = 0; c0 < S.FF; c0 += 32) {
                tc::tmem_ld32(tF + c0, q);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) q[i] = fmaxf(q[i] + V[S.v_b1 + c0 + i], 0.f);
                tc::store_row_bf16<32>(Af, kTcTile, r, q, c0 / 8);
            }
            // z = f1 W2^T ; x' = LN2(h + z + b2)
            gemm(tmem, af_s, wl + S.off_w2, D, S.FF);
            tc::tmem_ld32(tQ, q);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < D; ++i) x[i] += q[i] + V[S.v_b2 + i];
            layer_norm<D>(x, V + S.v_g2, V + S.v_be2);
        }
        // ---- acquisition MLP: logit = w2 . relu(W1 [z ; t] + b1) + b2 ----
        tc::store_row_bf16<D>(Ax, kTcTile, r, x);
        gemm(tmem + 32, ax_s, wb_s + S.off_acq, S.HH, D);
        float logit = Vec[S.v_acq_b2];
        for (int c0 = 0; c0 < S.HH; c0 += 32) {
            float hv[32];
            tc::tmem_ld32(tF + c0, hv);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float hh = hv[i] + Vec[S.v_acq_b1 + c0 + i] + t_value * Vec[S.v_acq_wt + c0 + i];
                logit = fmaf(fmaxf(hh, 0.f), Vec[S.v_acq_w2 + c0 + i], logit);
            }
        }
        if (in_range) {
            logits[(size_t)b * nq + j] = live ? logit : -INFINITY;
            if (zq && live) {
                float4* z = reinterpret_cast<float4*>(zq + ((size_t)b * nq + j) * D);
#pragma unroll
                for (int i = 0; i < D / 4; ++i) z[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

static size_t tc_smem_bytes(const TcShape& S) {
    size_t w = (S.total_bytes + 127) & ~127;
    size_t v = (size_t)((S.vec_total + 31) & ~31) * 4;
    size_t k = (size_t)S.NL * S.n_keys * 2 * kTcD * 4;
    int a_cols = S.FF > S.HH ? S.FF : S.HH;
    size_t a = 2 * ((size_t)kTcTile * kTcD * 2 + (size_t)kTcTile * a_cols * 2);
    return w + v + k + 128 + a;
}

uint64_t query_tc_weight_bytes(const Dims& d) { return (uint64_t)make_tc_shape(d, 0).total_bytes; }

// flag != NULL: fallback launch of the fast kernel (csrc/query_tc3.cu), runs only if *flag == epoch
int query_stream_tc(const Dims& d, const Layout& L, const float* P, const void* wb, const float* eq,
                    const unsigned char* alive, int B, int nq, const float* kv, int n_keys, int kv_slots, float t_value,
                    float* logits, float* zq, const int* flag, int epoch, cudaStream_t st) {
    ALINE_REQUIRE(d.D == kTcD, "tensor-core query stream supports dim_embedding 32 (got %d)", d.D);
    ALINE_REQUIRE(d.FF % 32 == 0 && d.FF <= 128 && d.HH % 32 == 0 && d.HH <= 128,
                  "tensor-core query stream supports feed-forward widths <= 128 (ff=%d head=%d)", d.FF, d.HH);
    TcShape S = make_tc_shape(d, n_keys);
    size_t smem = tc_smem_bytes(S);
    ALINE_REQUIRE(smem <= (size_t)device_info().max_smem_optin,
                  "tensor-core query stream: %d keys need %zu bytes of shared memory (max %d)", n_keys, smem,
                  device_info().max_smem_optin);
    if (ensure_dyn_smem((const void*)query_stream_tc_kernel, smem)) return 1;
    const int tiles = ceil_div(nq, kTcTile), pairs = ceil_div(tiles, 2);
    const int n_units = B * pairs;
    int grid = device_info().sm_count;
    if (grid > n_units) grid = n_units;
    // as the conditional fallback (flag given) the preceding launch of the stream is the fast kernel: dependent launch
    ALINE_CHECK_CUDA(launch_k(query_stream_tc_kernel, dim3(grid), dim3(256), smem, st, flag != nullptr || g_pdl_chain, d, L, S,
                              P, (const unsigned char*)wb, eq, alive, nq, kv, kv_slots, B, t_value, logits, zq, n_units,
                              pairs, (const int*)flag, epoch));
    ALINE_LAUNCH_OK();
    return 0;
}

// Largest key count the tensor-core kernel can hold in shared memory for this model
int query_stream_tc_max_keys(const Dims& d) {
    TcShape S = make_tc_shape(d, 0);
    size_t fixed = tc_smem_bytes(S);
    size_t lim = 227 * 1024;
    if (fixed >= lim) return 0;
    return (int)((lim - fixed) / ((size_t)d.NL * 2 * kTcD * 4));
}

}  // namespace aline
