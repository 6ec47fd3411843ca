// Candidate-query stream, fast tcgen05 path with the softmax probabilities and the MLP hidden activations kept in TENSOR
// MEMORY as MMA A-operands (tcgen05.st + TS-mode tcgen05.mma): P = 2^S is packed to bf16 in place over the score columns,
// relu(F) in place over the MLP1 accumulator columns, so neither tile is written to shared memory (no 16-byte stores, no
// generic->async proxy fence for them, 36 KB less shared memory per warpgroup) and the softmax needs one pass and one PV
// phase for any key count.  Every GEMM bias, the softmax shift and the softmax normaliser are folded into the tensor-core contractions, so the
// CUDA-core epilogues only convert, exponentiate and LayerNorm.
//
// Same contract as query_stream_kernel / query_stream_tc_kernel (reference: model/encoder.py:128-141 restricted to
// the candidate rows + model/head.py:27-31).  Per 128-token tile and layer, six MMA phases (fp32 accumulators in
// TMEM, one thread = one token = one TMEM lane):
//   Q   [x | 1] Wq'^T            Wq' = c Wq with the bias c bq in the "ones" columns, c = log2(e) / sqrt(8)
//   S   per head  [Q_h | 1] [K_h - K_0,h | mask]^T    scores relative to key 0 (softmax is shift-invariant), padded
//                                keys pushed to -200 by the mask column: p = 2^S needs no max pass and no select
//   PV  per head  P_h [V_h | 1]  the extra "ones" value column returns the softmax denominator
//   O   [o | 1] Wo'^T ;  F  [h | 1] W1'^T (ReLU fused into the bf16 conversion) ;  Z  [1 | f] W2'^T
// The operand "ones" chunk is [1, 1, t_hi, t_lo, 0...]: a bias is stored as bf16 hi + lo parts in the matching
// weight columns (fp32-accurate), the acquisition head's time-token weight multiplies t.
// FOLD (the shipped form up to 32 keys with >= 3 tiles per rollout): the Q and O phases and the o epilogue
// disappear -- the context kernel folds Wq into the key operand and Wo into the value operand (query_fast.cuh), the scores
// of all heads come from ONE MMA [x | 1] K'^T, the probabilities are normalised on the CUDA cores before they are packed
// and y = Pn V' accumulates over all heads: four phases per layer, 13 per tile.
// 2^S can overflow only if a score exceeds the score of key 0 by > 127 (88 nats); such a row makes its denominator
// non-finite, which raises a per-launch flag and the robust kernel (query_tc.cu, max-subtracted softmax, always
// enqueued right after, exits immediately when the flag is clear) recomputes the launch.
#include <atomic>
#include <cstdlib>
#include <type_traits>
#include "query_fast.cuh"

namespace aline {
namespace tc3 {

using namespace tcq;

// MLP2 accumulator columns: MLP1's columns [0, FF) hold relu(F) packed in [0, FF / 2 + 8) by then, so with only 128
// columns per warpgroup (four warpgroups) the accumulator goes over MLP1's consumed upper columns [96, 128)
__host__ __device__ constexpr uint32_t z_col(int nwg) { return nwg == 4 ? 96u : 128u; }

// x <- LayerNorm(x + y) over 32 features (biased variance, eps 1e-5) on packed fp32x2 operands (FADD2 / FFMA2 / FMUL2:
// half the issue slots of the scalar form -- the kernel is bound by the issue / latency of its epilogues); four
// independent accumulation chains
__device__ __forceinline__ void add_ln32(float (&x)[32], const float (&y)[32], const float* g, const float* b) {
    f32x2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = add2(pk2(x[2 * i], x[2 * i + 1]), pk2(y[2 * i], y[2 * i + 1]));
    f32x2 s0 = v[0], s1 = v[1], s2 = v[2], s3 = v[3];
#pragma unroll
    for (int i = 4; i < 16; i += 4) { s0 = add2(s0, v[i]); s1 = add2(s1, v[i + 1]); s2 = add2(s2, v[i + 2]); s3 = add2(s3, v[i + 3]); }
    float lo, hi;
    upk2(add2(add2(s0, s1), add2(s2, s3)), lo, hi);
    const float nmu = (lo + hi) * (-1.0f / 32);
    const f32x2 nmu2 = pk2(nmu, nmu);
    f32x2 q0 = pk2(0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        v[i] = add2(v[i], nmu2); v[i + 1] = add2(v[i + 1], nmu2); v[i + 2] = add2(v[i + 2], nmu2); v[i + 3] = add2(v[i + 3], nmu2);
        q0 = fma2(v[i], v[i], q0); q1 = fma2(v[i + 1], v[i + 1], q1);
        q2 = fma2(v[i + 2], v[i + 2], q2); q3 = fma2(v[i + 3], v[i + 3], q3);
    }
    upk2(add2(add2(q0, q1), add2(q2, q3)), lo, hi);
    const float rstd = rsqrtf((lo + hi) * (1.0f / 32) + 1e-5f);
    const f32x2 r2 = pk2(rstd, rstd);
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const float4 gg = *reinterpret_cast<const float4*>(g + 2 * i), bb = *reinterpret_cast<const float4*>(b + 2 * i);
        upk2(fma2(mul2(v[i], r2), pk2(gg.x, gg.y), pk2(bb.x, bb.y)), x[2 * i], x[2 * i + 1]);
        upk2(fma2(mul2(v[i + 1], r2), pk2(gg.z, gg.w), pk2(bb.z, bb.w)), x[2 * i + 2], x[2 * i + 3]);
    }
}

constexpr int kMaxFastKeys = 160;            // K / V operand blocks of 3 layers x 160 keys = 100 KB next to 93 KB of weights

// ---- folded operands (FOLD): two of the six MMA phases of a layer disappear ----
// S_h = c (x Wq_h^T + bq_h) (K_h - K_0h)^T = x K'_h^T + bias_h with K'_h[key] = c Wq_h^T (K_h - K_0h)[key]: the scores come
// straight from the token, no Q phase.  y = sum_h softmax_h V_h Wo_h^T = sum_h Pn_h V'_h with V'_h = V_h Wo_h^T and the
// probabilities NORMALISED before they are packed (the row sum is taken on the CUDA cores while exponentiating): all heads
// accumulate into ONE 32-column accumulator, no o epilogue and no O phase.  K' / V' come from the context kernel
// (fold_kv_emit, query_fast.cuh), once per rollout.  (First version: every CTA folded the plain blocks itself at each
// rollout change -- 7 / 12 us of a 111 / 139 us launch at 16 / 32 keys.)

template <int NWG, bool FOLD>
__global__ void __launch_bounds__(128 * NWG, 1)
query_tc3_kernel(const Dims m, const Layout L, const Tc2Shape S, const float* __restrict__ P,
                 const unsigned char* __restrict__ Wb_g, const float* __restrict__ eq,
                 const unsigned char* __restrict__ alive, int nq, int B, float t_hi, float t_lo,
                 float* __restrict__ logits, float* __restrict__ zq, int n_units, int tiles_per_b,
                 const unsigned char* __restrict__ tckv, int nkp, int rpu, int* __restrict__ flag, int epoch) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int D = kT2D;
    constexpr int TM = (512 / NWG) & ~7;                              // TMEM columns per warpgroup (256, 168 or 128)
    // PV accumulators (4 heads x 16 columns): above the scores when they fit, else over the upper half of the score
    // columns (every score has been read and the packed probabilities occupy only the lower half by then)
    // More than 48 keys (NWG == 2 only, 256 columns): the keys are processed in blocks of kKeyBlk = 32 -- scores relative
    // to key 0 need no running maximum, so P_j V_j simply ACCUMULATES over the blocks into PV accumulators that live above
    // the block's score columns.
    constexpr int kKeyBlk = 32;
    const bool key_blocks = nkp > 48;
    const uint32_t pv_col = FOLD ? z_col(NWG)
                            : key_blocks ? (uint32_t)(4 * kKeyBlk)
                                         : ((4 * nkp + 64 <= TM) ? (uint32_t)(4 * nkp) : (uint32_t)(2 * nkp));
    __shared__ __align__(8) uint64_t bar_w, bar_kv, bar_mma[NWG];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bo_s[FOLD ? 4 * kT2D : 4];          // FOLD: the out-projection bias joins LayerNorm 1

    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, r = tid & 127;
    const int kvblk = FOLD ? kFoldKeyBytes * nkp : tc2_kv_block_bytes(nkp), kbytes = FOLD ? 384 * nkp : tc2_k_bytes(nkp);
    // ---- carve shared memory ----
    unsigned char* Wb = smem;
    float* Vec = reinterpret_cast<float*>(Wb + ((S.total_bytes + 127) & ~127));
    unsigned char* KVb = reinterpret_cast<unsigned char*>(Vec + ((S.vec_total + 31) & ~31));
    const uint32_t kv_stride = (uint32_t)(((size_t)S.NL * kvblk + 127) & ~(size_t)127);     // one rollout's K / V blocks
    unsigned char* Abase = KVb + (size_t)rpu * kv_stride;
    const int xt_bytes = 6 * kT2Chunk;
    unsigned char* Xt = Abase + (size_t)wg * xt_bytes;                  // [128 x 48]: x / Q / o / h, ones chunk, zero chunk

    pdl_trigger();
    if (tid == 0) {
        tc::mbar_init(&bar_w, 1);
        tc::mbar_init(&bar_kv, 1);
        for (int i = 0; i < NWG; ++i) tc::mbar_init(&bar_mma[i], 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s + (uint32_t)wg * TM;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);    // this warp's lanes

    if (tid == 0) {
        tc::mbar_arrive_expect_tx(&bar_w, (uint32_t)S.total_bytes);
        tc::bulk_g2s(Wb, Wb_g, (uint32_t)S.total_bytes, &bar_w);
    }
    for (int l = 0; l < S.NL; ++l) {
        const float* Pl = P + L.layer0 + (size_t)l * L.layer_stride;
        float* V = Vec + l * S.vec_layer;
        for (int i = tid; i < D; i += 128 * NWG) {
            V[i] = Pl[L.g1 + i]; V[D + i] = Pl[L.be1 + i]; V[2 * D + i] = Pl[L.g2 + i]; V[3 * D + i] = Pl[L.be2 + i];
        }
    }
    for (int i = tid; i < S.HH; i += 128 * NWG) Vec[S.v_acq_w2 + i] = P[L.a_w2 + i];
    if (tid == 0) Vec[S.v_acq_b2] = P[L.a_b2];
    if constexpr (FOLD)
        for (int i = tid; i < S.NL * D; i += 128 * NWG) bo_s[i] = P[L.layer0 + (size_t)(i / D) * L.layer_stride + L.bo + (i % D)];
    {   // constant operand chunks of this thread's row
        const float ones[8] = {1.f, 1.f, t_hi, t_lo, 0.f, 0.f, 0.f, 0.f};
        const float zeros[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_chunk(Xt, 4, r, ones);
        store_chunk(Xt, 5, r, zeros);
    }
    tc::mbar_wait(&bar_w, 0);
    __syncthreads();
    // weights, barriers and tensor memory are in place; what follows reads what the preceding kernel of the stream wrote
    // (K / V operand blocks, alive flags) and writes the logits it may still read
    pdl_wait();

    const unsigned char* kv_src = tckv;               // FOLD: the host passes the fold region and nkp = its key padding
    const uint32_t wb_s = tc::smem_u32(Wb), xt_s = tc::smem_u32(Xt), kvb_s = tc::smem_u32(KVb);
    // the "ones" operand chunk [1, 1, t_hi, t_lo, 0 x 12] as 8 packed TMEM columns (A operand of the MLP2 bias step)
    uint32_t ones_pk[8];
    ones_pk[0] = pack2(1.f, 1.f); ones_pk[1] = pack2(t_hi, t_lo);
#pragma unroll
    for (int i = 2; i < 8; ++i) ones_pk[i] = 0u;
    uint32_t ph_mma = 0, ph_kv = 0;

    // ---- MMA issue ----
    // Clock stamps (profiles/r2_q4_phase_trace.txt) showed that a thread issuing a phase's 3-9 tcgen05.mma from
    // per-thread operands spends ~150 cycles per MMA: the descriptors live in ordinary registers and every UTCHMMA is
    // preceded by six R2UR.BROADCAST moves -- with the whole warpgroup waiting.  Here the FIRST WARP of the warpgroup
    // issues, convergently, from operands that are warp-uniform by construction: the warpgroup index is a compile-time
    // constant of the (unrolled) dispatch, everything else a kernel parameter or loop counter, so ptxas keeps the
    // descriptors in uniform registers; one elected lane executes the MMAs and the commit.
    enum { kQ = 0, kS, kPV, kO, kF, kZ, kAcq };
    const uint32_t smem_s = tc::smem_u32(smem);
    const uint32_t tmem_cta = tmem_base_s;
    auto issue_phase = [&](auto g_c, auto kind_c, int l, int k0, int nkb) {      // keys [k0, k0 + nkb) of the layer's nkp
        constexpr int G = decltype(g_c)::value, KIND = decltype(kind_c)::value;
        const uint32_t w_s = smem_s;                                                        // weights at the base
        const uint32_t vec_b = (uint32_t)((S.total_bytes + 127) & ~127) + (uint32_t)((S.vec_total + 31) & ~31) * 4u;
        const uint32_t kvb_u = smem_s + vec_b;
        const uint32_t xt_u = kvb_u + (uint32_t)rpu * kv_stride + (uint32_t)G * (uint32_t)xt_bytes;
        const uint32_t tm = tmem_cta + (uint32_t)G * TM;
        const uint32_t wl_u = w_s + (uint32_t)l * S.layer_bytes;
        const int bsel_u = G / (NWG / rpu);
        const uint32_t kb_u = kvb_u + (rpu == 1 ? 0u : (uint32_t)bsel_u * kv_stride) + (uint32_t)l * kvblk;
        const uint32_t vb_u = kb_u + (uint32_t)kbytes;
        if (tc::elect_one_sync()) {
            if constexpr (KIND == kQ) tc::umma_gemm(tm, xt_u, kT2Tile, wl_u + S.off_wq, D, D + 16, tc::idesc_bf16(128, D));
            if constexpr (KIND == kO) tc::umma_gemm(tm, xt_u, kT2Tile, wl_u + S.off_wo, D, D + 16, tc::idesc_bf16(128, D));
            if constexpr (KIND == kF) tc::umma_gemm(tm, xt_u, kT2Tile, wl_u + S.off_w1, S.FF, D + 16, tc::idesc_bf16(128, S.FF));
            if constexpr (KIND == kAcq) tc::umma_gemm(tm, xt_u, kT2Tile, w_s + S.off_acq, S.HH, D + 16, tc::idesc_bf16(128, S.HH));
            if constexpr (KIND == kS && FOLD)                 // all heads at once: [x | 1] K'^T
                tc::umma_gemm(tm, xt_u, kT2Tile, kb_u, 4 * nkp, D + 16, tc::idesc_bf16(128, 4 * nkp));
            if constexpr (KIND == kPV && FOLD) {              // y = Pn V' over the 4 nkp (head, key) columns
                const uint32_t idesc = tc::idesc_bf16(128, D);
                for (int s2 = 0; s2 < nkp / 4; ++s2)
                    tc::umma_bf16_ts(tm + pv_col, tm + (uint32_t)(8 * s2),
                                     tc::smem_desc(vb_u + (uint32_t)(2 * s2) * D * 16, D * 16, 128), idesc, s2 ? 1u : 0u);
            }
            if constexpr (KIND == kS && !FOLD) {
                const uint32_t idesc = tc::idesc_bf16(128, nkb);
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    tc::umma_bf16(tm + h * nkb, tc::smem_desc(xt_u + h * kT2Chunk, (4 - h) * kT2Chunk, 128),
                                  tc::smem_desc(kb_u + (h * nkp + k0) * 16, (4 - h) * nkp * 16, 128), idesc, 0u);
            }
            if constexpr (KIND == kPV && !FOLD) {
                const uint32_t idesc = tc::idesc_bf16(128, 16);
                const uint64_t vd0 = tc::smem_desc(vb_u, 256, 128);                        // + 16 per 256-byte V chunk
#pragma unroll
                for (int sblk = 0; sblk < 3; ++sblk) {                                     // key blocks outermost: the heads'
                    if (16 * sblk < nkb) {                                                 // accumulation chains interleave
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            tc::umma_bf16_ts(tm + pv_col + 16 * h, tm + (uint32_t)(h * (nkb / 2) + 8 * sblk),
                                             vd0 + (uint64_t)((h * (nkp / 8) + k0 / 8 + 2 * sblk) * 16), idesc,
                                             (sblk || k0) ? 1u : 0u);
                    }
                }
            }
            if constexpr (KIND == kZ) {                                                    // A = [f | ones] from tensor memory
                const uint32_t idesc = tc::idesc_bf16(128, D);
                const uint32_t w2_s = wl_u + S.off_w2;                  // W2' chunks: [ones (2 chunks) | f (FF / 8 chunks)]
                tc::umma_bf16_ts(tm + z_col(NWG), tm + (uint32_t)(S.FF / 2), tc::smem_desc(w2_s, D * 16, 128), idesc, 0u);
                for (int s2 = 0; s2 < S.FF / 16; ++s2)
                    tc::umma_bf16_ts(tm + z_col(NWG), tm + (uint32_t)(8 * s2),
                                     tc::smem_desc(w2_s + (uint32_t)(2 * (s2 + 1)) * D * 16, D * 16, 128), idesc, 1u);
            }
            tc::umma_commit(&bar_mma[G]);
        }
        __syncwarp();
    };
    auto mma_phase = [&](auto kind_c, int l, int k0 = 0, int nkb = 0) {
        // operands written with tcgen05.st (P before PV, relu(F) before Z) -> wait for those stores; operands written to
        // the shared-memory tile -> make them visible to the async proxy.  Each costs ~130 cycles even with nothing
        // outstanding (profiles/r2_q4_phase_trace_*.txt), and no phase needs both.
        constexpr int KIND_ = decltype(kind_c)::value;
        if constexpr (KIND_ == kPV || KIND_ == kZ) tc::tmem_st_wait();
        else tc::fence_async_smem();
        tc::tc_fence_before();
        tc::named_sync(1 + wg, 128);
        if ((warp & 3) == 0) {                       // the warpgroup's first warp issues, convergently
            tc::tc_fence_after();
#pragma unroll
            for (int g = 0; g < NWG; ++g) {
                if (warp == 4 * g) {
                    if (g == 0) issue_phase(std::integral_constant<int, 0>{}, kind_c, l, k0, nkb);
                    if (g == 1) issue_phase(std::integral_constant<int, 1>{}, kind_c, l, k0, nkb);
                    if (g == 2) issue_phase(std::integral_constant<int, (NWG > 2 ? 2 : 0)>{}, kind_c, l, k0, nkb);
                    if (g == 3) issue_phase(std::integral_constant<int, (NWG > 3 ? 3 : 0)>{}, kind_c, l, k0, nkb);
                }
            }
        }
        tc::mbar_wait(&bar_mma[wg], ph_mma);
        ph_mma ^= 1;
        tc::tc_fence_after();
    };

    // Schedule.  Every CTA first processes `full` = n_units / grid whole units (NWG tiles of one rollout, one per
    // warpgroup).  The n_units % grid left-over units would cost a whole extra round on some SMs while the others idle;
    // when they fit, they are split into SUB sub-units (NWG / SUB tiles of one rollout) dealt over ALL CTAs, so the last
    // round runs with half the warpgroups per SM -- and correspondingly faster -- instead of on a fraction of the SMs.
    constexpr int SUB = (NWG % 2 == 0) ? 2 : 1;
    const int grid = (int)gridDim.x, full = n_units / grid, n_tail = n_units - full * grid;
    const bool split_tail = SUB > 1 && rpu == 1 && n_tail * SUB <= grid;
    // Few candidates (the whole rollout fits NWG / rpu tiles): a unit = rpu consecutive rollouts, each with its own K / V
    // buffer and NWG / rpu warpgroups, so that all warpgroups have a tile.
    const int wpr = NWG / rpu, bsel = wg / wpr;
    const int n_sub = split_tail ? n_tail * SUB : n_tail;
    const int n_iter = full + (n_sub > 0 ? 1 : 0);
    int b_loaded = -1;
    bool bad = false;

    // tile of this warpgroup in iteration `it`: rollout b, first-rollout-of-the-unit b0, candidate j of this thread
    struct Coords { int b0, b, j; bool valid, active, in_range, live; };
    auto coords = [&](int it) {
        Coords c{};
        int unit, wg_off = 0, n_act = NWG;
        if (it < full) {
            unit = (int)blockIdx.x * full + it;
        } else {
            if (it >= n_iter || (int)blockIdx.x >= n_sub) return c;     // no (left-over) work for this CTA (uniform)
            if (split_tail) {
                unit = grid * full + (int)blockIdx.x / SUB;
                n_act = NWG / SUB;
                wg_off = ((int)blockIdx.x % SUB) * n_act;
            } else {
                unit = grid * full + (int)blockIdx.x;
            }
        }
        c.valid = true;
        // rpu == 1: unit = (rollout, tile group tg of NWG consecutive tiles); else unit = rpu consecutive rollouts
        c.b0 = rpu == 1 ? unit / tiles_per_b : unit * rpu;
        const int tg = rpu == 1 ? unit - c.b0 * tiles_per_b : 0;
        c.b = c.b0 + (rpu == 1 ? 0 : bsel);
        // a warpgroup without a tile (sub-unit round, or a tile group reaching past the last candidate) only takes part
        // in the K / V hand-over
        const int tile = rpu == 1 ? NWG * tg + wg_off + wg : wg - bsel * wpr;
        c.active = wg < n_act && c.b < B && tile * kT2Tile < nq;
        c.j = tile * kT2Tile + r;
        c.in_range = c.active && c.j < nq;
        c.live = c.in_range && (alive == nullptr || alive[(size_t)c.b * nq + c.j] != 0);
        return c;
    };
    float x[D];
    auto load_x = [&](const Coords& c) {
        // one 64-bit base per tile row, 32-bit feature offsets (the embeddings are [B][d][nq], candidate-minor)
        const float* pe = eq + (size_t)c.b * D * nq + (c.live ? c.j : 0);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float v = __ldg(pe + (unsigned)(i * nq));
            x[i] = c.live ? v : 0.f;
        }
    };
    bool have_x = false;                              // x already holds this iteration's embeddings (prefetched)

    for (int it = 0; it < n_iter; ++it) {
        const Coords cc = coords(it);
        if (!cc.valid) break;
        const int b0 = cc.b0, b = cc.b, j = cc.j;
        const bool active = cc.active, in_range = cc.in_range, live = cc.live;
        if (b0 != b_loaded || rpu > 1) {
            __syncthreads();                                            // everyone is done with the previous K, V
            if (tid == 0) {
                const int nb = (B - b0 < rpu) ? B - b0 : rpu;
                tc::mbar_arrive_expect_tx(&bar_kv, (uint32_t)(nb * S.NL * kvblk));
                for (int k = 0; k < nb; ++k)
                    for (int l = 0; l < S.NL; ++l)
                        tc::bulk_g2s(KVb + (size_t)k * kv_stride + (size_t)l * kvblk,
                                     kv_src + ((size_t)l * B + b0 + k) * kvblk, (uint32_t)kvblk, &bar_kv);
            }
        }
        if (!have_x) load_x(cc);
        have_x = false;
        if (b0 != b_loaded || rpu > 1) {
            tc::mbar_wait(&bar_kv, ph_kv);
            ph_kv ^= 1;
            b_loaded = b0;
        }
        if (!active) continue;

        float q[D];
        for (int l = 0; l < S.NL; ++l) {
            const float* V = Vec + l * S.vec_layer;
            if constexpr (FOLD) {
                // ---- S (all heads) straight from the token; Pn = 2^S / row sum packed IN PLACE; y = Pn V' + bo; LN1 ----
#pragma unroll
                for (int c = 0; c < 4; ++c) store_chunk(Xt, c, r, x + 8 * c);
                mma_phase(std::integral_constant<int, kS>{}, l);
                // one tcgen05.ld in flight at a time (wait::ld waits for every outstanding load), issued right after the
                // wait so that it overlaps the exponentials of the block that has just arrived; 16-column blocks in three
                // rotating buffers
                auto exp_sum = [&](float* p, f32x2& s0, f32x2& s1) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        p[i] = ex2f(p[i]); p[i + 1] = ex2f(p[i + 1]); p[i + 2] = ex2f(p[i + 2]); p[i + 3] = ex2f(p[i + 3]);
                        s0 = add2(s0, pk2(p[i], p[i + 1])); s1 = add2(s1, pk2(p[i + 2], p[i + 3]));
                    }
                };
                auto inverse = [&](f32x2 s0, f32x2 s1) {
                    float lo, hi;
                    upk2(add2(s0, s1), lo, hi);
                    const float den = lo + hi;
                    bad |= !(den < 1e30f);
                    const float inv = __fdividef(1.0f, den);
                    return pk2(inv, inv);
                };
                auto scale_pack = [&](const float* p, f32x2 inv2, uint32_t* pk) {
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        float a, c2;
                        upk2(mul2(pk2(p[i], p[i + 1]), inv2), a, c2);
                        pk[i / 2] = pack2(a, c2);
                    }
                };
                // NB blocks of 16 keys per head in NB + 1 rotating buffers (block k -> buffer k % (NB + 1))
                auto softmax_heads = [&](auto nb_c) {
                    constexpr int NB = decltype(nb_c)::value;
                    float pb[NB + 1][16];
                    tc::tmem_ld16(tl, pb[0]);
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        f32x2 s0 = pk2(0.f, 0.f), s1 = s0;
#pragma unroll
                        for (int j = 0; j < NB; ++j) {
                            const int k = NB * h + j;
                            tc::tmem_ld_wait16(pb[k % (NB + 1)]);
                            if (k + 1 < 4 * NB) tc::tmem_ld16(tl + 16 * (k + 1), pb[(k + 1) % (NB + 1)]);
                            exp_sum(pb[k % (NB + 1)], s0, s1);
                        }
                        const f32x2 inv2 = inverse(s0, s1);
                        uint32_t pk[8 * NB];
#pragma unroll
                        for (int j = 0; j < NB; ++j) scale_pack(pb[(NB * h + j) % (NB + 1)], inv2, pk + 8 * j);
                        if constexpr (NB == 1) tc::tmem_st8(tl + 8 * h, pk);
                        if constexpr (NB >= 2) tc::tmem_st16(tl + 8 * NB * h, pk);
                        if constexpr (NB == 3) tc::tmem_st8(tl + 8 * NB * h + 16, pk + 16);
                    }
                };
                auto exp_sum8 = [&](float* p, f32x2& s0, f32x2& s1) {
                    p[0] = ex2f(p[0]); p[1] = ex2f(p[1]); p[2] = ex2f(p[2]); p[3] = ex2f(p[3]);
                    p[4] = ex2f(p[4]); p[5] = ex2f(p[5]); p[6] = ex2f(p[6]); p[7] = ex2f(p[7]);
                    s0 = add2(s0, add2(pk2(p[0], p[1]), pk2(p[4], p[5])));
                    s1 = add2(s1, add2(pk2(p[2], p[3]), pk2(p[6], p[7])));
                };
                auto scale_pack8 = [&](const float* p, f32x2 inv2, uint32_t* pk) {
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        float a, c2;
                        upk2(mul2(pk2(p[i], p[i + 1]), inv2), a, c2);
                        pk[i / 2] = pack2(a, c2);
                    }
                };
                if (nkp == 32) softmax_heads(std::integral_constant<int, 2>{});
                else if (nkp == 16) softmax_heads(std::integral_constant<int, 1>{});
                else if (nkp == 8) {                                     // 8 keys: the four heads in one load and one store
                    float p[32];
                    uint32_t pk[16];
                    tc::tmem_ld32(tl, p);
                    tc::tmem_ld_wait32(p);
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        f32x2 s0 = pk2(0.f, 0.f), s1 = s0;
                        exp_sum8(p + 8 * h, s0, s1);
                        scale_pack8(p + 8 * h, inverse(s0, s1), pk + 4 * h);
                    }
                    tc::tmem_st16(tl, pk);
                } else if (nkp == 24) {                                  // 24 keys: 16 + 8 columns per head, two buffer sets
                    float pa[2][16], pc[2][8];
                    tc::tmem_ld16(tl, pa[0]);
                    tc::tmem_ld8(tl + 16, pc[0]);
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        float* a = pa[h & 1];
                        float* c = pc[h & 1];
                        tc::tmem_ld_wait16(a);
                        tc::tmem_ld_wait8(c);
                        if (h < 3) {
                            tc::tmem_ld16(tl + 24 * (h + 1), pa[(h + 1) & 1]);
                            tc::tmem_ld8(tl + 24 * (h + 1) + 16, pc[(h + 1) & 1]);
                        }
                        f32x2 s0 = pk2(0.f, 0.f), s1 = s0;
                        exp_sum(a, s0, s1);
                        exp_sum8(c, s0, s1);
                        const f32x2 inv2 = inverse(s0, s1);
                        uint32_t pk[12];
                        scale_pack(a, inv2, pk);
                        scale_pack8(c, inv2, pk + 8);
                        tc::tmem_st8(tl + 12 * h, pk);
                        tc::tmem_st4(tl + 12 * h + 8, pk + 8);
                    }
                } else if constexpr (NWG == 2) softmax_heads(std::integral_constant<int, 3>{});
                mma_phase(std::integral_constant<int, kPV>{}, l);
                tc::tmem_ld32(tl + pv_col, q);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < D; i += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(bo_s + l * D + i);
                    q[i] += bb.x; q[i + 1] += bb.y; q[i + 2] += bb.z; q[i + 3] += bb.w;
                }
                add_ln32(x, q, V, V + D);
            } else {
            // ---- Q ----
#pragma unroll
            for (int c = 0; c < 4; ++c) store_chunk(Xt, c, r, x + 8 * c);
            mma_phase(std::integral_constant<int, kQ>{}, l);
            tc::tmem_ld32(tl, q);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) store_chunk(Xt, c, r, q + 8 * c);
            // ---- per key block: S = Q_h (K_h - K_0h)^T + mask (all heads); P = 2^S packed to bf16 IN PLACE over the score
            //      columns (16 fp32 columns -> 8 packed columns); PV_h (+)= P_h [V_h | 1] with P read from tensor memory ----
            for (int k0 = 0; k0 < nkp; k0 += (key_blocks ? kKeyBlk : nkp)) {
                const int nkb = key_blocks ? (nkp - k0 < kKeyBlk ? nkp - k0 : kKeyBlk) : nkp;
                mma_phase(std::integral_constant<int, kS>{}, l, k0, nkb);
                const int nblk = 4 * nkb / 16;                           // 16-column blocks, all heads (even)
                float sa[16], sb[16];                                    // two blocks in flight
                tc::tmem_ld16(tl, sa);
                for (int blk = 0; blk < nblk; blk += 2) {
                    tc::tmem_ld_wait16(sa);
                    tc::tmem_ld16(tl + 16 * (blk + 1), sb);
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack2(ex2f(sa[2 * i]), ex2f(sa[2 * i + 1]));
                    tc::tmem_st8(tl + 8 * blk, pk);
                    tc::tmem_ld_wait16(sb);
                    if (blk + 2 < nblk) tc::tmem_ld16(tl + 16 * (blk + 2), sa);
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack2(ex2f(sb[2 * i]), ex2f(sb[2 * i + 1]));
                    tc::tmem_st8(tl + 8 * (blk + 1), pk);
                }
                mma_phase(std::integral_constant<int, kPV>{}, l, k0, nkb);
            }
            // ---- o = PV / denominator ----
            if constexpr (NWG >= 3) {
                // register-lean form (128 / 168 registers per thread): two heads at a time, straight to the operand tile
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    tc::tmem_ld32(tl + pv_col + 32 * half, q);
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
                        store_chunk(Xt, 2 * half + hh, r, o8);
                    }
                }
            } else {
                float o[D];
                float pv[2][32];
                tc::tmem_ld32(tl + pv_col, pv[0]);
                tc::tmem_ld32(tl + pv_col + 32, pv[1]);
                tc::tmem_ld_wait();
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float* dh = &pv[h >> 1][16 * (h & 1)];
                    const float den = dh[8];
                    bad |= !(den < 1e30f);
                    const float inv = __fdividef(1.0f, den);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[8 * h + i] = dh[i] * inv;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) store_chunk(Xt, c, r, o + 8 * c);
            }
            // ---- y = [o | 1] Wo'^T ; h = LN1(x + y) ----
            mma_phase(std::integral_constant<int, kO>{}, l);
            tc::tmem_ld32(tl, q);
            tc::tmem_ld_wait();
            add_ln32(x, q, V, V + D);
            }
            // ---- f = relu([h | 1] W1'^T) ----
#pragma unroll
            for (int c = 0; c < 4; ++c) store_chunk(Xt, c, r, x + 8 * c);
            mma_phase(std::integral_constant<int, kF>{}, l);
            {   // relu + bf16 pack IN PLACE: accumulator columns [32 j, 32 j + 32) -> packed columns [16 j, 16 j + 16)
                float fa[32], fb[32];
                auto put = [&](const float* v, int blk) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = pack2_relu(v[2 * i], v[2 * i + 1]);
                    tc::tmem_st16(tl + 16 * blk, pk);
                };
                if constexpr (NWG >= 3) {
                    for (int c0 = 0; c0 < S.FF / 32; ++c0) {              // one block in flight: other warps hide the latency
                        tc::tmem_ld32(tl + 32 * c0, fa);
                        tc::tmem_ld_wait32(fa);
                        put(fa, c0);
                    }
                } else {
                    tc::tmem_ld32(tl, fa);
                    for (int c0 = 0; c0 < S.FF / 32; c0 += 2) {            // FF / 32 blocks, two per trip
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
                }
                tc::tmem_st8(tl + S.FF / 2, ones_pk);                    // bias / time-token operand chunk
            }
            // ---- z = [1 | f] W2'^T ; x' = LN2(h + z) ----
            mma_phase(std::integral_constant<int, kZ>{}, l);
            tc::tmem_ld32(tl + z_col(NWG), q);
            tc::tmem_ld_wait();
            add_ln32(x, q, V + 2 * D, V + 3 * D);
        }
        // ---- acquisition MLP: logit = w2 . relu([z | 1, t] Wa'^T) + b2 ----
#pragma unroll
        for (int c = 0; c < 4; ++c) store_chunk(Xt, c, r, x + 8 * c);
        mma_phase(std::integral_constant<int, kAcq>{}, 0);
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
            if constexpr (NWG >= 3) {
                for (int c0 = 0; c0 < nb; ++c0) {
                    tc::tmem_ld32(tl + 32 * c0, ha);
                    tc::tmem_ld_wait32(ha);
                    acc32(ha, c0);
                }
            } else {
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

static size_t tc2_smem_bytes(const Tc2Shape& S, int nkp, int nwg, int* f_chunks_out, int rpu = 1, bool fold = false) {
    if (f_chunks_out) *f_chunks_out = 0;
    size_t w = (S.total_bytes + 127) & ~127;
    size_t v = (size_t)((S.vec_total + 31) & ~31) * 4;
    size_t k = ((size_t)S.NL * (fold ? kFoldKeyBytes * nkp : tc2_kv_block_bytes(nkp)) + 127) & ~(size_t)127;
    size_t a = (size_t)nwg * (size_t)6 * kT2Chunk;
    return w + v + (size_t)rpu * k + a;
}

bool supported(const Dims& d, int n_keys) {
    if (d.D != kT2D || d.FF % 32 != 0 || d.FF > 128 || d.FF < 32 || d.HH % 32 != 0 || d.HH > 128 || d.HH < 32) return false;
    if (n_keys < 1 || n_keys > kMaxFastKeys) return false;
    Tc2Shape S = make_tc2_shape(d);
    return tc2_smem_bytes(S, (n_keys + 15) / 16 * 16, 2, nullptr) <= (size_t)device_info().max_smem_optin;
}


// folded operands (query_tc3_kernel<4, true> up to 32 keys, <2, true> at 33-48): -1 / 2 up to 32 keys (the shipped
// rule: at 33-48 keys the kernel gains 13 us on query_tc4 but the context kernel pays 18 us for the operands), 0 off,
// 1 up to 48 keys
static std::atomic<int> g_fold_mode{-2};
void set_fold(int v) { g_fold_mode.store(v, std::memory_order_relaxed); }
static int fold_mode() {
    int mode = g_fold_mode.load(std::memory_order_relaxed);
    if (mode == -2) {
        const char* e = getenv("ALINE_QUERY_FOLD");
        mode = e ? atoi(e) : -1;
        g_fold_mode.store(mode, std::memory_order_relaxed);
    }
    return mode;
}

// Launch plan of the one-thread-per-row kernel for (shape, key count, candidates): warpgroups per CTA, rollouts per
// unit, and whether the folded form runs.  nq = 0: candidates unknown (the context kernel of a stand-alone call) -- the
// fold is then assumed for the shapes that have one for SOME candidate count.
struct Plan { int NWG, rpu; bool fold; int nkf; };      // nkf: key padding of the folded operands (8 with four warpgroups)
static Plan make_plan(const Dims& d, const Tc2Shape& S, int n_keys, int nq) {
    const int nkp = (n_keys + 15) / 16 * 16;
    // warpgroups (= tiles in flight) per CTA.  Measured at cfg2 (us per launch at 16 / 32 padded keys): 2 -> 187 / 208,
    // 3 -> 199 / 217 (18-for-16 tile padding and 8.1 -> 9 unit quantisation eat its +15 % tiles/s), 4 -> 170 / 194
    // (128 registers per thread, 300 B of spills, +20 % tiles/s per SM).  ALINE_QUERY_WG overrides.
    static const int want_wg = [] {
        const char* e = getenv("ALINE_QUERY_WG");
        return e ? atoi(e) : 4;
    }();
    Plan p;
    // three / four warpgroups (tiles in flight per SM) need the scores + PV accumulators in 168 / 128 TMEM columns:
    // <= 32 keys
    p.NWG = ((want_wg == 3 || want_wg == 4) && 4 * nkp <= 128 && S.FF / 2 + 8 <= 96) ? want_wg : 2;
    // rollouts per unit: when a rollout has only 1 or 2 tiles, NWG / tiles rollouts share a unit (one K / V buffer each)
    p.rpu = 1;
    const int tiles = nq > 0 ? ceil_div(nq, kT2Tile) : 3;
    if ((tiles == 1 || tiles == 2) && p.NWG % tiles == 0 && p.NWG / tiles > 1) {
        p.rpu = p.NWG / tiles;
        while (p.rpu > 1 && tc2_smem_bytes(S, nkp, p.NWG, nullptr, p.rpu) > (size_t)device_info().max_smem_optin) p.rpu /= 2;
    }
    // folded operands: four warpgroups up to 32 keys (two at 33-48 with query_fold = 1), <= 4 layers (bias staging), one
    // rollout per unit (with 1-2 tiles per rollout the fold measured neutral to +1 %: cfg1 6.29 -> 6.34 ms, cfg4 theta
    // 5.53 -> 5.61, cfg5 2.26 -> 2.18 / 2.15 -> 2.16 -- the context kernel pays as much as the few tiles gain)
    p.nkf = p.NWG == 4 ? (n_keys + 7) / 8 * 8 : nkp;
    p.fold = p.rpu == 1 && d.D == kT2D && d.H == 4 && d.NL <= 4 && n_keys >= 1 && fold_mode() != 0 &&
             n_keys <= (fold_mode() == 1 ? 48 : 32) && (p.NWG == 4 || (p.NWG == 2 && nkp == 48)) &&
             tc2_smem_bytes(S, p.nkf, p.NWG, nullptr, p.rpu, true) <= (size_t)device_info().max_smem_optin;
    return p;
}

// Candidates per rollout of the launches that follow on this thread (aline_rollout sets it around its chain, 0 =
// unknown): lets the context kernel skip the folded operands when the candidate stream will not use them.
static thread_local int g_nq_hint = 0;
static thread_local bool g_plain_needed = true;       // a kernel that reads the plain operand blocks may follow (query_tc4 forced)
void set_nq_hint(int nq, bool plain_needed) { g_nq_hint = nq; g_plain_needed = nq <= 0 || plain_needed; }
bool fold_only() { return g_nq_hint > 0 && !g_plain_needed; }

// do the context kernels emit the folded operands for this shape?
int fold_keys(const Dims& d, int n_keys) {             // padded key count of the emitted folded operands, 0: none
    if (d.D != kT2D || !supported(d, n_keys)) return 0;
    const Plan p = make_plan(d, make_tc2_shape(d), n_keys, g_nq_hint);
    return p.fold ? p.nkf : 0;
}
bool fold_emitted(const Dims& d, int n_keys) { return fold_keys(d, n_keys) > 0; }

// launch; flag / epoch: see the header comment
int launch(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq,
                     const unsigned char* alive, int B, int nq, int n_keys, float t_value, float* logits, float* zq,
                     const void* tckv, int* flag, int epoch, cudaStream_t st) {
    ALINE_REQUIRE(supported(d, n_keys), "fast tensor-core query stream: unsupported shape (d=%d ff=%d head=%d "
                  "keys=%d)", d.D, d.FF, d.HH, n_keys);
    Tc2Shape S = make_tc2_shape(d);
    int nkp = (n_keys + 15) / 16 * 16;
    const Plan plan = make_plan(d, S, n_keys, nq);
    // the folded form needs the operands the context kernel emitted: same rule, evaluated with the hint it saw
    const bool fold = plan.fold && fold_keys(d, n_keys) == plan.nkf;
    const int NWG = plan.NWG, rpu = plan.rpu;
    const int tiles = ceil_div(nq, kT2Tile);
    // the folded kernel sees only its own region of the buffer and its own key padding
    const int nkp_plain = nkp;
    if (fold) {
        tckv = (const unsigned char*)tckv + tc2_fold_offset(S.NL, B, nkp_plain);
        nkp = plan.nkf;
    }
    const size_t smem = tc2_smem_bytes(S, nkp, NWG, nullptr, rpu, fold);
    const int groups = ceil_div(tiles, NWG);
    const int n_units = rpu > 1 ? ceil_div(B, rpu) : B * groups;
    int grid = device_info().sm_count;
    if (grid > n_units) grid = n_units;
    const __nv_bfloat16 th = __float2bfloat16_rn(t_value);
    const float t_hi = __bfloat162float(th), t_lo = t_value - t_hi;
    if (fold && NWG == 2) {
        if (ensure_dyn_smem((const void*)query_tc3_kernel<2, true>, smem)) return 1;
        ALINE_CHECK_CUDA(launch_k(query_tc3_kernel<2, true>, dim3(grid), dim3(256), smem, st, g_pdl_chain, d, L, S, P,
                                  (const unsigned char*)wb2, eq, alive, nq, B, t_hi, t_lo, logits, zq, n_units, groups,
                                  (const unsigned char*)tckv, nkp, rpu, flag, epoch));
    } else if (fold) {
        if (ensure_dyn_smem((const void*)query_tc3_kernel<4, true>, smem)) return 1;
        ALINE_CHECK_CUDA(launch_k(query_tc3_kernel<4, true>, dim3(grid), dim3(512), smem, st, g_pdl_chain, d, L, S, P,
                                  (const unsigned char*)wb2, eq, alive, nq, B, t_hi, t_lo, logits, zq, n_units, groups,
                                  (const unsigned char*)tckv, nkp, rpu, flag, epoch));
    } else if (NWG == 4) {
        if (ensure_dyn_smem((const void*)query_tc3_kernel<4, false>, smem)) return 1;
        ALINE_CHECK_CUDA(launch_k(query_tc3_kernel<4, false>, dim3(grid), dim3(512), smem, st, g_pdl_chain, d, L, S, P,
                                  (const unsigned char*)wb2, eq, alive, nq, B, t_hi, t_lo, logits, zq, n_units, groups,
                                  (const unsigned char*)tckv, nkp, rpu, flag, epoch));
    } else if (NWG == 3) {
        if (ensure_dyn_smem((const void*)query_tc3_kernel<3, false>, smem)) return 1;
        ALINE_CHECK_CUDA(launch_k(query_tc3_kernel<3, false>, dim3(grid), dim3(384), smem, st, g_pdl_chain, d, L, S, P,
                                  (const unsigned char*)wb2, eq, alive, nq, B, t_hi, t_lo, logits, zq, n_units, groups,
                                  (const unsigned char*)tckv, nkp, rpu, flag, epoch));
    } else {
        if (ensure_dyn_smem((const void*)query_tc3_kernel<2, false>, smem)) return 1;
        ALINE_CHECK_CUDA(launch_k(query_tc3_kernel<2, false>, dim3(grid), dim3(256), smem, st, g_pdl_chain, d, L, S, P,
                                  (const unsigned char*)wb2, eq, alive, nq, B, t_hi, t_lo, logits, zq, n_units, groups,
                                  (const unsigned char*)tckv, nkp, rpu, flag, epoch));
    }
    ALINE_LAUNCH_OK();
    return 0;
}

}  // namespace tc3

bool query_tc3_supported(const Dims& d, int n_keys) { return tc3::supported(d, n_keys); }
void query_tc3_set_fold(int v) { tc3::set_fold(v); }
bool query_tc3_fold_emitted(const Dims& d, int n_keys) { return tc3::fold_emitted(d, n_keys); }
int query_tc3_fold_keys(const Dims& d, int n_keys) { return tc3::fold_keys(d, n_keys); }
void query_tc3_set_nq_hint(int nq, bool plain_needed) { tc3::set_nq_hint(nq, plain_needed); }
bool query_tc3_fold_only() { return tc3::fold_only(); }
uint64_t query_tc3_weight_bytes(const Dims& d) { return (uint64_t)tc3::make_tc2_shape(d).total_bytes; }

int query_stream_tc3(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq,
                     const unsigned char* alive, int B, int nq, int n_keys, float t_value, float* logits, float* zq,
                     const void* tckv, int* flag, int epoch, cudaStream_t st) {
    return tc3::launch(d, L, P, wb2, eq, alive, B, nq, n_keys, t_value, logits, zq, tckv, flag, epoch, st);
}

}  // namespace aline
