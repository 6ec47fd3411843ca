// Candidate-query stream, fast tcgen05 path for dim_embedding = 64 / 8 heads (the psychometric model of
// config/model/aline_psychometric.yaml in the reference).  Same contract as the d = 32 kernels (csrc/query_tc3.cu;
// reference: model/encoder.py:128-141 restricted to the candidate rows + model/head.py:27-31) and the same numerics:
// bf16 operands, fp32 accumulation in tensor memory, every bias / the softmax shift / the softmax normaliser folded into
// the contractions, P = 2^S and relu(F) kept in TENSOR MEMORY as A operands.
//
// What differs from d = 32 is where things live.  The folded bf16 weights of one layer are 58 KB (Wq', Wo' 64 x 80,
// W1' 128 x 80, W2' 64 x 144), so three layers + the acquisition layer do not fit next to the operand tiles: the weights
// are STREAMED, one layer per stage, through two shared-memory slots (cp.async.bulk from L2, one stage ahead; the two
// warpgroups of a CTA walk the stages in lock step, one __syncthreads per stage).  A tile has 8 heads, so the scores of
// a 16-key block take 128 of the warpgroup's 256 tensor-memory columns and the 8 PV accumulators (16 columns each, the
// extra "ones" value column returns the softmax denominator) the other 128; the keys are processed in blocks of 16 and
// P_j V_j accumulates over the blocks (scores are relative to key 0, so no running maximum is needed).
//
// Per 128-candidate tile and layer: Q, (S, PV) per key block, O, F, Z -- 4 + 2 * ceil(n_keys / 16) MMA phases; one
// thread = one candidate row = one tensor-memory lane, two tiles (warpgroups) in flight per SM.
#include <type_traits>
#include "query_fast.cuh"

namespace aline {
namespace tc5 {

using namespace tcq;

constexpr int D = 64, H = 8, KA = D + 16;
constexpr int kChunks = D / 8 + 2;               // operand tile: 8 data chunks, the "ones" chunk, a zero chunk
constexpr int kOnes = D / 8, kZero = D / 8 + 1;
constexpr int kKeyBlk = 16;
constexpr int kMaxKeys = 48;                     // K / V operand blocks of 3 layers x 48 keys = 56 KB next to 116 KB of weight slots
constexpr uint32_t kTM = 256;                    // tensor-memory columns per warpgroup
constexpr uint32_t kPvCol = 128, kZCol = 128;

// bytes of the bf16 key / value operand block of one (layer, rollout): K part (H + 1) chunks x nkp rows x 16 B
// (heads + mask chunk), V part H heads x nkp/8 chunks x 16 rows x 16 B (8 features, ones row, 7 zero rows)
__host__ __device__ inline int k_bytes(int nkp) { return (H + 1) * 16 * nkp; }
__host__ __device__ inline int kv_block_bytes(int nkp) { return k_bytes(nkp) + H * 32 * nkp; }

// x <- LayerNorm(x + y) over 64 features (biased variance, eps 1e-5), packed fp32x2 arithmetic
__device__ __forceinline__ void add_ln64(float (&x)[D], const float (&y)[D], const float* g, const float* b) {
    f32x2 v[D / 2];
#pragma unroll
    for (int i = 0; i < D / 2; ++i) v[i] = add2(pk2(x[2 * i], x[2 * i + 1]), pk2(y[2 * i], y[2 * i + 1]));
    f32x2 s0 = v[0], s1 = v[1], s2 = v[2], s3 = v[3];
#pragma unroll
    for (int i = 4; i < D / 2; i += 4) { s0 = add2(s0, v[i]); s1 = add2(s1, v[i + 1]); s2 = add2(s2, v[i + 2]); s3 = add2(s3, v[i + 3]); }
    float lo, hi;
    upk2(add2(add2(s0, s1), add2(s2, s3)), lo, hi);
    const float nmu = (lo + hi) * (-1.0f / D);
    const f32x2 nmu2 = pk2(nmu, nmu);
    f32x2 q0 = pk2(0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
#pragma unroll
    for (int i = 0; i < D / 2; i += 4) {
        v[i] = add2(v[i], nmu2); v[i + 1] = add2(v[i + 1], nmu2); v[i + 2] = add2(v[i + 2], nmu2); v[i + 3] = add2(v[i + 3], nmu2);
        q0 = fma2(v[i], v[i], q0); q1 = fma2(v[i + 1], v[i + 1], q1);
        q2 = fma2(v[i + 2], v[i + 2], q2); q3 = fma2(v[i + 3], v[i + 3], q3);
    }
    upk2(add2(add2(q0, q1), add2(q2, q3)), lo, hi);
    const float rstd = rsqrtf((lo + hi) * (1.0f / D) + 1e-5f);
    const f32x2 r2 = pk2(rstd, rstd);
#pragma unroll
    for (int i = 0; i < D / 2; i += 2) {
        const float4 gg = *reinterpret_cast<const float4*>(g + 2 * i), bb = *reinterpret_cast<const float4*>(b + 2 * i);
        upk2(fma2(mul2(v[i], r2), pk2(gg.x, gg.y), pk2(bb.x, bb.y)), x[2 * i], x[2 * i + 1]);
        upk2(fma2(mul2(v[i + 1], r2), pk2(gg.z, gg.w), pk2(bb.z, bb.w)), x[2 * i + 2], x[2 * i + 3]);
    }
}

__device__ __forceinline__ float (&half32(float* p))[32] { return *reinterpret_cast<float(*)[32]>(p); }

__host__ __device__ inline uint32_t slot_bytes(const Tc2Shape& S) {
    const uint32_t acq = (uint32_t)(S.total_bytes - S.off_acq);
    const uint32_t m = (uint32_t)S.layer_bytes > acq ? (uint32_t)S.layer_bytes : acq;
    return (m + 127u) & ~127u;
}

__global__ void __launch_bounds__(256, 1)
query_tc5_kernel(const Dims m, const Layout L, const Tc2Shape S, const float* __restrict__ P,
                 const unsigned char* __restrict__ Wb_g, const float* __restrict__ eq,
                 const unsigned char* __restrict__ alive, int nq, int B, float t_hi, float t_lo,
                 float* __restrict__ logits, float* __restrict__ zq, int n_units, int tiles_per_b,
                 const unsigned char* __restrict__ tckv, int nkp, int* __restrict__ flag, int epoch) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int NWG = 2;
    __shared__ __align__(8) uint64_t bar_w[2], bar_kv, bar_mma[NWG];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, r = tid & 127;
    const int kvblk = kv_block_bytes(nkp), kbytes = k_bytes(nkp);
    // ---- carve shared memory: two weight slots | fp32 vectors | K / V operand blocks of the rollout | operand tiles ----
    const uint32_t slot = slot_bytes(S);
    const uint32_t vec_bytes = (uint32_t)((S.vec_total + 31) & ~31) * 4u;
    const uint32_t kv_all = (uint32_t)(((size_t)S.NL * kvblk + 127) & ~(size_t)127);
    constexpr uint32_t xt_bytes = kChunks * kT2Chunk;
    float* Vec = reinterpret_cast<float*>(smem + 2 * slot);
    unsigned char* KVb = smem + 2 * slot + vec_bytes;
    unsigned char* Xt = KVb + kv_all + (size_t)wg * xt_bytes;           // [128 x 80]: x / Q / o / h, ones chunk, zero chunk

    // Schedule (as csrc/query_tc3.cu): `full` whole units (two tiles of one rollout, one per warpgroup) per CTA; the
    // left-over units are split into single-tile sub-units dealt over all CTAs when they fit.
    const int grid = (int)gridDim.x, full = n_units / grid, n_tail = n_units - full * grid;
    const bool split_tail = n_tail * 2 <= grid;
    const int n_sub = split_tail ? n_tail * 2 : n_tail;
    const int n_iter = full + ((int)blockIdx.x < n_sub ? 1 : 0);       // iterations of THIS CTA
    const int n_stage_cta = n_iter * (S.NL + 1);                        // weight stages this CTA walks through

    // stage n of this CTA = layer (n mod (NL + 1)) of its tile round (the last one: the acquisition layer)
    auto load_stage = [&](int n) {
        const int s = n % (S.NL + 1);
        const uint32_t bytes = s < S.NL ? (uint32_t)S.layer_bytes : (uint32_t)(S.total_bytes - S.off_acq);
        const unsigned char* src = Wb_g + (s < S.NL ? (size_t)s * S.layer_bytes : (size_t)S.off_acq);
        tc::mbar_arrive_expect_tx(&bar_w[n & 1], bytes);
        tc::bulk_g2s(smem + (size_t)(n & 1) * slot, src, bytes, &bar_w[n & 1]);
    };

    pdl_trigger();
    if (tid == 0) {
        tc::mbar_init(&bar_w[0], 1);
        tc::mbar_init(&bar_w[1], 1);
        tc::mbar_init(&bar_kv, 1);
        for (int i = 0; i < NWG; ++i) tc::mbar_init(&bar_mma[i], 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s + (uint32_t)wg * kTM;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);    // this warp's lanes

    if (tid == 0 && n_stage_cta > 0) load_stage(0);
    for (int l = 0; l < S.NL; ++l) {
        const float* Pl = P + L.layer0 + (size_t)l * L.layer_stride;
        float* V = Vec + l * S.vec_layer;
        for (int i = tid; i < D; i += 128 * NWG) {
            V[i] = Pl[L.g1 + i]; V[D + i] = Pl[L.be1 + i]; V[2 * D + i] = Pl[L.g2 + i]; V[3 * D + i] = Pl[L.be2 + i];
        }
    }
    for (int i = tid; i < S.HH; i += 128 * NWG) Vec[S.v_acq_w2 + i] = P[L.a_w2 + i];
    if (tid == 0) Vec[S.v_acq_b2] = P[L.a_b2];
    {   // constant operand chunks of this thread's row
        const float ones[8] = {1.f, 1.f, t_hi, t_lo, 0.f, 0.f, 0.f, 0.f};
        const float zeros[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_chunk(Xt, kOnes, r, ones);
        store_chunk(Xt, kZero, r, zeros);
    }
    __syncthreads();
    // barriers, vectors and tensor memory are in place; what follows reads what the preceding kernel of the stream wrote
    // (K / V operand blocks, alive flags) and writes the logits it may still read
    pdl_wait();

    uint32_t ones_pk[8];
    ones_pk[0] = pack2(1.f, 1.f); ones_pk[1] = pack2(t_hi, t_lo);
#pragma unroll
    for (int i = 2; i < 8; ++i) ones_pk[i] = 0u;
    uint32_t ph_mma = 0, ph_kv = 0;

    // ---- MMA issue: the warpgroup's first warp, convergently, from warp-uniform operands (see csrc/query_tc3.cu) ----
    enum { kQ = 0, kS, kPV, kO, kF, kZ, kAcq };
    const uint32_t smem_s = tc::smem_u32(smem);
    const uint32_t tmem_cta = tmem_base_s;
    auto issue_phase = [&](auto g_c, auto kind_c, int l, int k0, int wsl) {       // keys [k0, k0 + 16) of the layer's nkp
        constexpr int G = decltype(g_c)::value, KIND = decltype(kind_c)::value;
        const uint32_t w_s = smem_s + (uint32_t)wsl * slot;                                 // this stage's weights
        const uint32_t kvb_u = smem_s + 2 * slot + vec_bytes;
        const uint32_t xt_u = kvb_u + kv_all + (uint32_t)G * xt_bytes;
        const uint32_t tm = tmem_cta + (uint32_t)G * kTM;
        const uint32_t kb_u = kvb_u + (uint32_t)l * kvblk;
        const uint32_t vb_u = kb_u + (uint32_t)kbytes;
        if (tc::elect_one_sync()) {
            if constexpr (KIND == kQ) tc::umma_gemm(tm, xt_u, kT2Tile, w_s + S.off_wq, D, KA, tc::idesc_bf16(128, D));
            if constexpr (KIND == kO) tc::umma_gemm(tm, xt_u, kT2Tile, w_s + S.off_wo, D, KA, tc::idesc_bf16(128, D));
            if constexpr (KIND == kF) tc::umma_gemm(tm, xt_u, kT2Tile, w_s + S.off_w1, S.FF, KA, tc::idesc_bf16(128, S.FF));
            if constexpr (KIND == kAcq) tc::umma_gemm(tm, xt_u, kT2Tile, w_s, S.HH, KA, tc::idesc_bf16(128, S.HH));
            if constexpr (KIND == kS) {
                const uint32_t idesc = tc::idesc_bf16(128, kKeyBlk);
#pragma unroll
                for (int h = 0; h < H; ++h)
                    tc::umma_bf16(tm + h * kKeyBlk, tc::smem_desc(xt_u + h * kT2Chunk, (kOnes - h) * kT2Chunk, 128),
                                  tc::smem_desc(kb_u + (h * nkp + k0) * 16, (H - h) * nkp * 16, 128), idesc, 0u);
            }
            if constexpr (KIND == kPV) {
                const uint32_t idesc = tc::idesc_bf16(128, 16);
#pragma unroll
                for (int h = 0; h < H; ++h)
                    tc::umma_bf16_ts(tm + kPvCol + 16 * h, tm + (uint32_t)(8 * h),
                                     tc::smem_desc(vb_u + (uint32_t)(h * (nkp / 8) + k0 / 8) * 256u, 256, 128), idesc,
                                     k0 ? 1u : 0u);
            }
            if constexpr (KIND == kZ) {                                                    // A = [f | ones] from tensor memory
                const uint32_t idesc = tc::idesc_bf16(128, D);
                const uint32_t w2_s = w_s + S.off_w2;                   // W2' chunks: [ones (2 chunks) | f (FF / 8 chunks)]
                tc::umma_bf16_ts(tm + kZCol, tm + (uint32_t)(S.FF / 2), tc::smem_desc(w2_s, D * 16, 128), idesc, 0u);
                for (int s2 = 0; s2 < S.FF / 16; ++s2)
                    tc::umma_bf16_ts(tm + kZCol, tm + (uint32_t)(8 * s2),
                                     tc::smem_desc(w2_s + (uint32_t)(2 * (s2 + 1)) * D * 16, D * 16, 128), idesc, 1u);
            }
            tc::umma_commit(&bar_mma[G]);
        }
        __syncwarp();
    };
    auto mma_phase = [&](auto kind_c, int l, int wsl, int k0 = 0) {
        constexpr int KIND_ = decltype(kind_c)::value;
        if constexpr (KIND_ == kPV || KIND_ == kZ) tc::tmem_st_wait();
        else tc::fence_async_smem();
        tc::tc_fence_before();
        tc::named_sync(1 + wg, 128);
        if ((warp & 3) == 0) {                       // the warpgroup's first warp issues, convergently
            tc::tc_fence_after();
            if (warp == 0) issue_phase(std::integral_constant<int, 0>{}, kind_c, l, k0, wsl);
            else issue_phase(std::integral_constant<int, 1>{}, kind_c, l, k0, wsl);
        }
        tc::mbar_wait(&bar_mma[wg], ph_mma);
        ph_mma ^= 1;
        tc::tc_fence_after();
    };
    // start of weight stage `sc` (CTA-uniform counter): every warpgroup is done with stage sc - 1, so its slot can take
    // stage sc + 1; then wait for this stage's weights (issued one stage ago, normally long landed)
    auto stage_begin = [&](int sc) {
        __syncthreads();
        if (tid == 0 && sc + 1 < n_stage_cta) load_stage(sc + 1);
        tc::mbar_wait(&bar_w[sc & 1], (uint32_t)((sc >> 1) & 1));
    };

    int b_loaded = -1;
    bool bad = false;
    int sc = 0;
    float x[D];

    for (int it = 0; it < n_iter; ++it) {
        // tile of this warpgroup: rollout b, tile index within the rollout
        int unit, wg_off = 0, n_act = NWG;
        if (it < full) {
            unit = (int)blockIdx.x * full + it;
        } else if (split_tail) {
            unit = grid * full + (int)blockIdx.x / 2;
            n_act = 1;
            wg_off = (int)blockIdx.x % 2;
        } else {
            unit = grid * full + (int)blockIdx.x;
        }
        const int b = unit / tiles_per_b, tg = unit - b * tiles_per_b;
        const int tile = NWG * tg + wg_off + wg;
        const bool active = wg < n_act && tile * kT2Tile < nq;
        const int j = tile * kT2Tile + r;
        const bool in_range = active && j < nq;
        const bool live = in_range && (alive == nullptr || alive[(size_t)b * nq + j] != 0);
        if (b != b_loaded) {
            __syncthreads();                                            // everyone is done with the previous K, V
            if (tid == 0) {
                tc::mbar_arrive_expect_tx(&bar_kv, (uint32_t)(S.NL * kvblk));
                for (int l = 0; l < S.NL; ++l)
                    tc::bulk_g2s(KVb + (size_t)l * kvblk, tckv + ((size_t)l * B + b) * kvblk, (uint32_t)kvblk, &bar_kv);
            }
        }
        {   // embeddings [B][d][nq], candidate-minor: one 64-bit base per row, 32-bit feature offsets
            const float* pe = eq + (size_t)b * D * nq + (live ? j : 0);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const float v = __ldg(pe + (unsigned)(i * nq));
                x[i] = live ? v : 0.f;
            }
        }
        if (b != b_loaded) {
            tc::mbar_wait(&bar_kv, ph_kv);
            ph_kv ^= 1;
            b_loaded = b;
        }

        float q[D];
        for (int l = 0; l < S.NL; ++l, ++sc) {
            stage_begin(sc);
            if (!active) continue;
            const int wsl = sc & 1;
            const float* V = Vec + l * S.vec_layer;
            // ---- Q ----
#pragma unroll
            for (int c = 0; c < D / 8; ++c) store_chunk(Xt, c, r, x + 8 * c);
            mma_phase(std::integral_constant<int, kQ>{}, l, wsl);
            tc::tmem_ld32(tl, half32(q));
            tc::tmem_ld32(tl + 32, half32(q + 32));
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < D / 8; ++c) store_chunk(Xt, c, r, q + 8 * c);
            // ---- per block of 16 keys: S = Q_h (K_h - K_0h)^T + mask (all heads: 8 x 16 columns); P = 2^S packed to
            //      bf16 IN PLACE (head h: columns [16 h, 16 h + 16) -> [8 h, 8 h + 8)); PV_h (+)= P_h [V_h | 1] ----
            for (int k0 = 0; k0 < nkp; k0 += kKeyBlk) {
                mma_phase(std::integral_constant<int, kS>{}, l, wsl, k0);
                float sa[16], sb[16];                                    // two heads in flight
                tc::tmem_ld16(tl, sa);
#pragma unroll 1
                for (int h = 0; h < H; h += 2) {
                    tc::tmem_ld_wait16(sa);
                    tc::tmem_ld16(tl + 16 * (h + 1), sb);
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack2(ex2f(sa[2 * i]), ex2f(sa[2 * i + 1]));
                    tc::tmem_st8(tl + 8 * h, pk);
                    tc::tmem_ld_wait16(sb);
                    if (h + 2 < H) tc::tmem_ld16(tl + 16 * (h + 2), sa);
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack2(ex2f(sb[2 * i]), ex2f(sb[2 * i + 1]));
                    tc::tmem_st8(tl + 8 * (h + 1), pk);
                }
                mma_phase(std::integral_constant<int, kPV>{}, l, wsl, k0);
            }
            // ---- o = PV / denominator, two heads at a time, straight to the operand tile ----
#pragma unroll
            for (int hp = 0; hp < H / 2; ++hp) {
                tc::tmem_ld32(tl + kPvCol + 32 * hp, half32(q));
                tc::tmem_ld_wait();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const float den = q[16 * hh + 8];
                    bad |= !(den < 1e30f);
                    const float inv = __fdividef(1.0f, den);
                    const f32x2 inv2 = pk2(inv, inv);
                    float o8[8];
#pragma unroll
                    for (int i = 0; i < 8; i += 2)
                        upk2(mul2(pk2(q[16 * hh + i], q[16 * hh + i + 1]), inv2), o8[i], o8[i + 1]);
                    store_chunk(Xt, 2 * hp + hh, r, o8);
                }
            }
            // ---- y = [o | 1] Wo'^T ; h = LN1(x + y) ----
            mma_phase(std::integral_constant<int, kO>{}, l, wsl);
            tc::tmem_ld32(tl, half32(q));
            tc::tmem_ld32(tl + 32, half32(q + 32));
            tc::tmem_ld_wait();
            add_ln64(x, q, V, V + D);
            // ---- f = relu([h | 1] W1'^T) ----
#pragma unroll
            for (int c = 0; c < D / 8; ++c) store_chunk(Xt, c, r, x + 8 * c);
            mma_phase(std::integral_constant<int, kF>{}, l, wsl);
            {   // relu + bf16 pack IN PLACE: accumulator columns [32 j, 32 j + 32) -> packed columns [16 j, 16 j + 16)
                float fa[32], fb[32];
                auto put = [&](const float* v, int blk) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = pack2_relu(v[2 * i], v[2 * i + 1]);
                    tc::tmem_st16(tl + 16 * blk, pk);
                };
                tc::tmem_ld32(tl, fa);
                for (int c0 = 0; c0 < S.FF / 32; c0 += 2) {
                    tc::tmem_ld_wait32(fa);
                    const bool has_b = c0 + 1 < S.FF / 32;
                    if (has_b) tc::tmem_ld32(tl + 32 * (c0 + 1), fb);
                    put(fa, c0);
                    if (has_b) {
                        tc::tmem_ld_wait32(fb);
                        if (c0 + 2 < S.FF / 32) tc::tmem_ld32(tl + 32 * (c0 + 2), fa);
                        put(fb, c0 + 1);
                    }
                }
                tc::tmem_st8(tl + S.FF / 2, ones_pk);                    // bias / time-token operand chunk
            }
            // ---- z = [1 | f] W2'^T ; x' = LN2(h + z) ----
            mma_phase(std::integral_constant<int, kZ>{}, l, wsl);
            tc::tmem_ld32(tl + kZCol, half32(q));
            tc::tmem_ld32(tl + kZCol + 32, half32(q + 32));
            tc::tmem_ld_wait();
            add_ln64(x, q, V + 2 * D, V + 3 * D);
        }
        // ---- acquisition MLP: logit = w2 . relu([z | 1, t] Wa'^T) + b2 ----
        stage_begin(sc);
        const int wsl_a = sc & 1;
        ++sc;
        if (!active) continue;
#pragma unroll
        for (int c = 0; c < D / 8; ++c) store_chunk(Xt, c, r, x + 8 * c);
        mma_phase(std::integral_constant<int, kAcq>{}, 0, wsl_a);
        float lg0 = Vec[S.v_acq_b2], lg1 = 0.f;
        {
            float ha[32], hb[32];
            const int nb = S.HH / 32;
            auto acc32 = [&](const float* cur, int c0) {
                const float* w2 = Vec + S.v_acq_w2 + 32 * c0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(w2 + i);
                    lg0 = fmaf(fmaxf(cur[i], 0.f), w.x, lg0); lg1 = fmaf(fmaxf(cur[i + 1], 0.f), w.y, lg1);
                    lg0 = fmaf(fmaxf(cur[i + 2], 0.f), w.z, lg0); lg1 = fmaf(fmaxf(cur[i + 3], 0.f), w.w, lg1);
                }
            };
            tc::tmem_ld32(tl, ha);
            for (int c0 = 0; c0 < nb; c0 += 2) {
                tc::tmem_ld_wait32(ha);
                const bool has_b = c0 + 1 < nb;
                if (has_b) tc::tmem_ld32(tl + 32 * (c0 + 1), hb);
                acc32(ha, c0);
                if (has_b) {
                    tc::tmem_ld_wait32(hb);
                    if (c0 + 2 < nb) tc::tmem_ld32(tl + 32 * (c0 + 2), ha);
                    acc32(hb, c0 + 1);
                }
            }
        }
        if (in_range) {
            logits[(size_t)b * nq + j] = live ? lg0 + lg1 : -INFINITY;
            if (zq && live) {
                float4* z = reinterpret_cast<float4*>(zq + ((size_t)b * nq + j) * D);
#pragma unroll
                for (int i = 0; i < D / 4; ++i) z[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
            }
            if (bad && live) *flag = epoch;
        }
        bad = false;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

static size_t smem_bytes(const Tc2Shape& S, int nkp) {
    return 2 * (size_t)slot_bytes(S) + (size_t)((S.vec_total + 31) & ~31) * 4 +
           (((size_t)S.NL * kv_block_bytes(nkp) + 127) & ~(size_t)127) + 2 * (size_t)kChunks * kT2Chunk;
}

bool supported(const Dims& d, int n_keys) {
    if (d.D != D || d.H != H || d.FF % 32 != 0 || d.FF > 128 || d.FF < 32 || d.HH % 32 != 0 || d.HH > 128 || d.HH < 32) return false;
    if (n_keys < 1 || n_keys > kMaxKeys) return false;
    Tc2Shape S = make_tc2_shape(d);
    return smem_bytes(S, (n_keys + 15) / 16 * 16) <= (size_t)device_info().max_smem_optin;
}

int launch(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq, const unsigned char* alive,
           int B, int nq, int n_keys, float t_value, float* logits, float* zq, const void* tckv, int* flag, int epoch,
           cudaStream_t st) {
    ALINE_REQUIRE(supported(d, n_keys), "fast tensor-core query stream (d = 64): unsupported shape (d=%d heads=%d ff=%d "
                  "head=%d keys=%d)", d.D, d.H, d.FF, d.HH, n_keys);
    Tc2Shape S = make_tc2_shape(d);
    const int nkp = (n_keys + 15) / 16 * 16;
    const size_t smem = smem_bytes(S, nkp);
    const int tiles = ceil_div(nq, kT2Tile), groups = ceil_div(tiles, 2);
    const int n_units = B * groups;
    int grid = device_info().sm_count;
    if (grid > n_units) grid = n_units;
    const __nv_bfloat16 th = __float2bfloat16_rn(t_value);
    const float t_hi = __bfloat162float(th), t_lo = t_value - t_hi;
    if (ensure_dyn_smem((const void*)query_tc5_kernel, smem)) return 1;
    ALINE_CHECK_CUDA(launch_k(query_tc5_kernel, dim3(grid), dim3(256), smem, st, g_pdl_chain, d, L, S, P,
                              (const unsigned char*)wb2, eq, alive, nq, B, t_hi, t_lo, logits, zq, n_units, groups,
                              (const unsigned char*)tckv, nkp, flag, epoch));
    ALINE_LAUNCH_OK();
    return 0;
}

}  // namespace tc5

bool query_tc5_supported(const Dims& d, int n_keys) { return tc5::supported(d, n_keys); }
int query_tc5_kv_block_bytes(int nkp) { return tc5::kv_block_bytes(nkp); }

int query_stream_tc5(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq,
                     const unsigned char* alive, int B, int nq, int n_keys, float t_value, float* logits, float* zq,
                     const void* tckv, int* flag, int epoch, cudaStream_t st) {
    return tc5::launch(d, L, P, wb2, eq, alive, B, nq, n_keys, t_value, logits, zq, tckv, flag, epoch, st);
}

}  // namespace aline
