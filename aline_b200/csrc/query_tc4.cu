// Candidate-query stream, fast tcgen05 path with TWO THREADS PER CANDIDATE ROW (d = 32, ff = head = 128, <= 48 keys).
//
// Same arithmetic, weight blob, key / value operand blocks and MMA phases as csrc/query_tc3.cu (see its header: biases,
// softmax shift and softmax denominator folded into the contractions; P = 2^S and relu(F) packed to bf16 in place in
// tensor memory and consumed as TMEM A operands).  What changes is who runs the CUDA-core epilogues.  query_tc3 gives a
// 128-row tile to one warpgroup, one thread per row: ncu showed it latency-bound with 4 warps per scheduler (issue
// slots 39 % busy; a third of the warp time waiting for an MMA phase or the phase barrier; the 128-register budget
// and the 512 TMEM columns forbid a fifth tile in flight).  A warp may only touch its own 32-lane quadrant of tensor
// memory, but TWO warps (w and w + 4) may share a quadrant: here a tile belongs to a group of 8 warps, and the two
// threads of a row split every epilogue by columns --
//     16 of the 32 model features (operand chunks 2c, 2c+1), attention heads {2c, 2c+1}, 64 of the 128 hidden units --
// so a thread keeps 16 residual values instead of 32, the per-phase dependency chain halves, and NT = 3 (80 registers)
// or 4 (64 registers, <= 16 keys) tiles = 6-8 warps per scheduler are in flight.  LayerNorm needs the statistics of
// the whole row: each thread reduces its half to (mean, M2), the pair exchanges them through shared memory (one
// 64-thread named barrier) and merges them with the parallel-variance formula (as accurate as the two-pass form).
// The acquisition head's second layer (128 -> 1) also runs on the tensor pipe: relu(F) is packed like the MLP
// activations and multiplied with [w2_hi | w2_lo] (bf16 split of the fp32 weights, built in shared memory by the
// prologue), so the epilogue reads two columns instead of 128.
//
// In-place packing with two threads per row: a thread may only overwrite columns it has itself consumed.  Thread c owns
// score columns [2c nkp, (2c+2) nkp) and packs P for its two heads into the lower half of that range; relu(F) of
// columns [64c, 64c+64) goes to [64c, 64c+32); the PV / MLP2 / logit accumulators are placed over consumed inputs.
//
// Candidate embeddings are read row-major ([B][nq][32], aline_embed_queries_ex): a thread's 16 features are four
// 16-byte loads, the pair covers one 128-byte line.
#include <type_traits>
#include "query_fast.cuh"

namespace aline {
namespace tc4 {

using namespace tcq;

constexpr int kD = kT2D;
constexpr int kFF = 128, kHH = 128;                          // the only feed-forward / head widths this kernel is built for
// byte offsets inside one layer of the bf16 weight blob (make_tc2_shape with D = 32, FF = 128): Wq', Wo', W1', W2'
constexpr int kKA = kD + 16;
constexpr uint32_t kOffWq = 0, kOffWo = kD * kKA * 2, kOffW1 = 2 * kD * kKA * 2, kOffW2 = kOffW1 + kFF * kKA * 2;
constexpr uint32_t kLayerBytes = kOffW2 + kD * (kFF + 16) * 2;
constexpr int kVecLayer = 4 * kD;
constexpr int kXtBytes = 6 * kT2Chunk;                       // [128 x 48] bf16 operand tile: x / Q / o / h, ones, zeros
constexpr int kLnBytes = 2 * 2 * kT2Tile * (int)sizeof(float2);   // LayerNorm exchange slots: 2 sets x 2 halves x 128 rows
constexpr int kWa2Rows = 16;                                 // acquisition layer 2 as an N = 16 operand (rows 0, 1 used)

__host__ __device__ constexpr int tm_cols(int nt) { return (512 / nt) & ~31; }     // 256, 160, 128

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
    a = __uint_as_float(r0); b = __uint_as_float(r1);
}

// x <- LayerNorm(x + y) over the 32 features of a row held as two 16-feature halves by two threads (this thread: half c).
// Each thread reduces its half to (mean_c, M2_c); the partner's pair arrives through `mine` / `other` (shared memory)
// after the pair barrier; merged with Chan's parallel formula: mean = (m0 + m1) / 2, M2 = M2_0 + M2_1 + 8 (m0 - m1)^2.
__device__ __forceinline__ void add_ln_half(float (&x)[16], const float (&y)[16], const float* g, const float* b,
                                            float2* mine, const float2* other, int bar_id, int bar_threads) {
    f32x2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = add2(pk2(x[2 * i], x[2 * i + 1]), pk2(y[2 * i], y[2 * i + 1]));
    float lo, hi;
    upk2(add2(add2(add2(v[0], v[1]), add2(v[2], v[3])), add2(add2(v[4], v[5]), add2(v[6], v[7]))), lo, hi);
    const float mc = (lo + hi) * (1.0f / 16);
    const f32x2 nmc = pk2(-mc, -mc);
    f32x2 q0 = pk2(0.f, 0.f), q1 = q0;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        v[i] = add2(v[i], nmc); v[i + 1] = add2(v[i + 1], nmc);
        q0 = fma2(v[i], v[i], q0); q1 = fma2(v[i + 1], v[i + 1], q1);
    }
    upk2(add2(q0, q1), lo, hi);
    const float m2c = lo + hi;
    *mine = make_float2(mc, m2c);
    tc::named_sync(bar_id, bar_threads);
    const float2 o = *other;
    const float dm = mc - o.x;
    const float var = (m2c + o.y + 8.0f * dm * dm) * (1.0f / 32);
    const float rstd = rsqrtf(var + 1e-5f);
    const float shift = 0.5f * dm;                                  // mean_c - mean
    const f32x2 sh2 = pk2(shift, shift), r2 = pk2(rstd, rstd);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const float4 gg = *reinterpret_cast<const float4*>(g + 2 * i), bb = *reinterpret_cast<const float4*>(b + 2 * i);
        upk2(fma2(mul2(add2(v[i], sh2), r2), pk2(gg.x, gg.y), pk2(bb.x, bb.y)), x[2 * i], x[2 * i + 1]);
        upk2(fma2(mul2(add2(v[i + 1], sh2), r2), pk2(gg.z, gg.w), pk2(bb.z, bb.w)), x[2 * i + 2], x[2 * i + 3]);
    }
}

// ---- who issues the MMAs ------------------------------------------------------------------------------------------
// Clock stamps inside the first version of this kernel (profiles/r2_q4_phase_trace.txt) showed where a phase went: of
// ~2300 cycles, ~800 were the ELECTED EPILOGUE THREAD issuing the phase's 3-9 tcgen05.mma with all eight warps of the
// tile waiting on it -- ~150 cycles per MMA, because its descriptors lived in ordinary registers and every UTCHMMA was
// preceded by six R2UR.BROADCAST moves into uniform registers (a second version that replayed precomputed descriptor
// records from shared memory had the same six moves per MMA and, with one issuing thread for three tiles, was 2x
// slower).  So each tile group gets its own ISSUER WARP whose control flow and operands are warp-uniform by
// construction (kernel parameters, loop counters, compile-time group index): ptxas keeps the descriptors in uniform
// registers and an MMA costs a handful of uniform-datapath instructions.  The issuer walks the fixed phase sequence of
// a tile; before each phase it waits on the group's "operands ready" mbarrier (8 arrivals: lane 0 of every epilogue
// warp, after the warp's proxy / tcgen05 fences), then one elected lane issues and commits to the group's "MMA done"
// mbarrier.  Epilogue warps never issue and never meet at a block-level barrier on the phase path.
__host__ __device__ inline uint32_t pv_col_hd(int h, int nkp, int TM) {
    if (4 * nkp + 64 <= TM) return (uint32_t)(4 * nkp + 16 * h);
    return (uint32_t)(h < 2 ? nkp + 16 * h : 3 * nkp + 16 * (h - 2));       // over the consumed scores of heads 1 / 3
}
__host__ __device__ inline uint32_t p_col_hd(int h, int nkp) { return (uint32_t)((h >> 1) * 2 * nkp + (h & 1) * (nkp / 2)); }

// shared-memory carve (bytes from the dynamic shared-memory base), the same on host and device
struct Carve {
    uint32_t wb, vec, wa2, per, kv_bytes, per_tg, total;
};
__host__ __device__ inline Carve make_carve(int total_w_bytes, int vec_total, int NL, int nkp, int nt) {
    Carve c;
    c.wb = 0;
    c.vec = (uint32_t)((total_w_bytes + 127) & ~127);
    c.wa2 = c.vec + (uint32_t)((vec_total + 31) & ~31) * 4u;
    c.per = c.wa2 + (uint32_t)((((kHH / 8 + 2) * kWa2Rows * 16) + 127) & ~127);
    c.kv_bytes = (uint32_t)((NL * tc2_kv_block_bytes(nkp) + 127) & ~127);
    c.per_tg = c.kv_bytes + kXtBytes + kLnBytes;
    c.total = c.per + (uint32_t)nt * c.per_tg;
    return c;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

#ifdef ALINE_Q4_TRACE
__device__ long long g_q4_trace[3 * 4096];
#define Q4_T(slot) do { if (trace_on) { long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)); if ((trace_n & 4095) < 4095) g_q4_trace[trace_n++] = c_ * 8 + (slot); } } while (0)
#else
#define Q4_T(slot) do { } while (0)
#endif

template <int NT>
__global__ void __launch_bounds__(256 * NT + 32 * NT, 1)
query_tc4_kernel(const Dims m, const Layout L, const Tc2Shape S, const float* __restrict__ P,
                 const unsigned char* __restrict__ Wb_g, const float* __restrict__ eq_rm,
                 const unsigned char* __restrict__ alive, int nq, int B, float t_hi, float t_lo,
                 float* __restrict__ logits, float* __restrict__ zq, int tiles_per_b,
                 const unsigned char* __restrict__ tckv, int nkp, int* __restrict__ flag, int epoch) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int TM = tm_cols(NT);
    constexpr int kThreads = 256 * NT + 32 * NT;
    constexpr int FF = kFF, HH = kHH;
    constexpr bool kLean = NT >= 3;                                  // <= 80 registers: one TMEM block in flight, 16-column blocks
    __shared__ __align__(8) uint64_t bar_w, bar_kv[NT], bar_mma[NT], bar_rdy[NT];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, tg = tid >> 8, tt = tid & 255, wl = tt >> 5, lane = tid & 31;
    const bool is_issuer = tid >= 256 * NT;                          // the last NT warps: one MMA issuer per tile group
    const int qd = wl & 3, c = wl >> 2, r = 32 * qd + lane;          // lane quadrant, column half, tile row
    const int kvblk = tc2_kv_block_bytes(nkp), kbytes = tc2_k_bytes(nkp);
    (void)kbytes;
    // ---- carve shared memory (host twin: make_carve) ----
    const Carve cv = make_carve(S.total_bytes, S.vec_total, S.NL, nkp, NT);
    unsigned char* Wb = smem + cv.wb;
    float* Vec = reinterpret_cast<float*>(smem + cv.vec);
    unsigned char* Wa2 = smem + cv.wa2;
    unsigned char* KVb = smem + cv.per + (size_t)(is_issuer ? 0 : tg) * cv.per_tg;
    unsigned char* Xt = KVb + cv.kv_bytes;
    float2* Ln = reinterpret_cast<float2*>(Xt + kXtBytes);

    pdl_trigger();
    if (tid == 0) {
        tc::mbar_init(&bar_w, 1);
        for (int i = 0; i < NT; ++i) {
            tc::mbar_init(&bar_kv[i], 1);
            tc::mbar_init(&bar_mma[i], 1);
            tc::mbar_init(&bar_rdy[i], 8);                            // lane 0 of each of the group's 8 epilogue warps
        }
        tc::fence_mbar_init();
    }
    if (tid < 32) tc::tmem_alloc(&tmem_base_s, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t tmem = tmem_base + (uint32_t)(is_issuer ? 0 : tg) * TM;
    const uint32_t tl = tmem + ((uint32_t)(32 * qd) << 16);           // this warp's lanes

    if (tid == 0) {
        tc::mbar_arrive_expect_tx(&bar_w, (uint32_t)S.total_bytes);
        tc::bulk_g2s(Wb, Wb_g, (uint32_t)S.total_bytes, &bar_w);
    }
    for (int l = 0; l < S.NL; ++l) {
        const float* Pl = P + L.layer0 + (size_t)l * L.layer_stride;
        float* V = Vec + l * kVecLayer;
        for (int i = tid; i < kD; i += kThreads) {
            V[i] = Pl[L.g1 + i]; V[kD + i] = Pl[L.be1 + i]; V[2 * kD + i] = Pl[L.g2 + i]; V[3 * kD + i] = Pl[L.be2 + i];
        }
    }
    // acquisition layer 2 as a [16 x (16 + HH)] bf16 operand: row 0 = [b2_hi, b2_lo, 0 .. | w2_hi], row 1 = [0 .. | w2_lo]
    const int wa2_bytes = (HH / 8 + 2) * kWa2Rows * 16;
    for (int i = tid; i < wa2_bytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(Wa2)[i] = 0u;
    __syncthreads();
    {
        __nv_bfloat16* w = reinterpret_cast<__nv_bfloat16*>(Wa2);
        auto at = [&](int row, int col) -> __nv_bfloat16& { return w[((col >> 3) * kWa2Rows + row) * 8 + (col & 7)]; };
        for (int k = tid; k < HH; k += kThreads) {
            const float v = P[L.a_w2 + k];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            at(0, 16 + k) = h;
            at(1, 16 + k) = __float2bfloat16_rn(v - __bfloat162float(h));
        }
        if (tid == 0) {
            const float v = P[L.a_b2];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            at(0, 0) = h;
            at(0, 1) = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
    if (!is_issuer) {   // constant operand chunks of this row: [1, 1, t_hi, t_lo, 0 ..] (chunk 4) and zeros (chunk 5)
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (c == 0) { q.x = pack2(1.f, 1.f); q.y = pack2(t_hi, t_lo); }
        *reinterpret_cast<uint4*>(Xt + (size_t)(4 + c) * kT2Chunk + (size_t)r * 16) = q;
    }
    tc::fence_async_smem();
    tc::mbar_wait(&bar_w, 0);
    __syncthreads();
    // weights, barriers and tensor memory are in place; what follows reads what the preceding kernel of the stream wrote
    // (K / V operand blocks, alive flags) and writes the logits it may still read
    pdl_wait();

    // tiles of this CTA: a contiguous, balanced range of the B x tiles_per_b tiles, dealt round-robin to its NT groups
    const long long n_tiles = (long long)B * tiles_per_b;
    const int t0 = (int)(n_tiles * blockIdx.x / gridDim.x), t1 = (int)(n_tiles * (blockIdx.x + 1) / gridDim.x);

    if (is_issuer) {
        // ================= MMA issuers: warp 8 NT + t serves tile group t (everything below is warp-uniform) =================
        const uint32_t smem_s = tc::smem_u32(smem);
        const int kbytes_u = tc2_k_bytes(nkp);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if ((tid >> 5) != 8 * NT + t) continue;
#ifdef ALINE_Q4_TRACE
            const bool trace_on = blockIdx.x == 1 && t == 0 && lane == 0;
            int trace_n = 2 * 4096;
#endif
            const int n_t = t0 + t < t1 ? (t1 - t0 - t + NT - 1) / NT : 0;
            const uint32_t kvb_s = smem_s + cv.per + (uint32_t)t * cv.per_tg, xt_s = kvb_s + cv.kv_bytes;
            const uint32_t wb_s = smem_s + cv.wb, wa2_s = smem_s + cv.wa2;
            const uint32_t tm = tmem_base + (uint32_t)(t * TM);
            uint32_t par = 0;
            auto wait_ready = [&] {
                tc::mbar_wait(&bar_rdy[t], par);
                par ^= 1;
                tc::tc_fence_after();
                Q4_T(6);
            };
            auto done = [&] {
                tc::umma_commit(&bar_mma[t]);
                Q4_T(7);
            };
            // [operand tile | 1 | 0] (K = 48) times W[N x 48]^T -> columns 0..N
            auto gemm_xt = [&](uint32_t w_s, int N) {
                wait_ready();
                if (elect_one()) {
                    tc::umma_gemm(tm, xt_s, kT2Tile, w_s, N, kD + 16, tc::idesc_bf16(128, N));
                    done();
                }
                __syncwarp();
            };
            // [ones | relu(f)] (tensor memory) times W'[N x (16 + n_hidden)]^T -> columns 32..32+N
            auto gemm_f = [&](uint32_t w_s, int N, int n_hidden) {
                wait_ready();
                if (elect_one()) {
                    const uint32_t idesc = tc::idesc_bf16(128, N);
                    tc::umma_bf16_ts(tm + 32, tm + (uint32_t)(n_hidden / 2 + n_hidden / 4), tc::smem_desc(w_s, N * 16, 128), idesc, 0u);
#pragma unroll
                    for (int s2 = 0; s2 < n_hidden / 16; ++s2)
                        tc::umma_bf16_ts(tm + 32, tm + (uint32_t)(8 * s2 + (s2 >= n_hidden / 32 ? n_hidden / 4 : 0)),
                                         tc::smem_desc(w_s + (uint32_t)(2 * (s2 + 1)) * N * 16, N * 16, 128), idesc, 1u);
                    done();
                }
                __syncwarp();
            };
            uint32_t pv_d[4], pv_a[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) { pv_d[h] = tm + pv_col_hd(h, nkp, TM); pv_a[h] = tm + p_col_hd(h, nkp); }
            for (int it = 0; it < n_t; ++it) {
                for (int l = 0; l < S.NL; ++l) {
                    const uint32_t wl_s = wb_s + (uint32_t)l * kLayerBytes;
                    const uint32_t kb_s = kvb_s + (uint32_t)(l * kvblk), vb_s = kb_s + (uint32_t)kbytes_u;
                    gemm_xt(wl_s + kOffWq, kD);                              // Q
                    wait_ready();                                            // S: per head [Q_h | 1] [K_h - K_0h | mask]^T
                    if (elect_one()) {
                        const uint32_t idesc = tc::idesc_bf16(128, nkp);
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            tc::umma_bf16(tm + (uint32_t)(h * nkp), tc::smem_desc(xt_s + h * kT2Chunk, (4 - h) * kT2Chunk, 128),
                                          tc::smem_desc(kb_s + h * nkp * 16, (4 - h) * nkp * 16, 128), idesc, 0u);
                        done();
                    }
                    __syncwarp();
                    wait_ready();                                            // PV: per head P_h [V_h | 1]
                    if (elect_one()) {
                        const uint32_t idesc = tc::idesc_bf16(128, 16);
                        const uint64_t vd0 = tc::smem_desc(vb_s, 256, 128);           // + 16 per 256-byte V chunk
#pragma unroll
                        for (int sblk = 0; sblk < 3; ++sblk) {                        // key blocks outermost: the four heads'
                            if (16 * sblk < nkp) {                                    // accumulation chains interleave
#pragma unroll
                                for (int h = 0; h < 4; ++h)
                                    tc::umma_bf16_ts(pv_d[h], pv_a[h] + (uint32_t)(8 * sblk),
                                                     vd0 + (uint64_t)((h * (nkp / 8) + 2 * sblk) * 16), idesc, sblk ? 1u : 0u);
                            }
                        }
                        done();
                    }
                    __syncwarp();
                    gemm_xt(wl_s + kOffWo, kD);                              // O
                    gemm_xt(wl_s + kOffW1, FF);                              // F
                    gemm_f(wl_s + kOffW2, kD, FF);                           // Z
                }
                gemm_xt(wb_s + (uint32_t)S.NL * kLayerBytes, HH);            // acquisition layer 1
                gemm_f(wa2_s, kWa2Rows, HH);                                 // acquisition layer 2
            }
        }
    } else {
    // ================= epilogue warps =================
    const uint32_t f_in = (uint32_t)(c * (FF / 2));                   // this thread's MLP1 / head-1 accumulator columns
    constexpr uint32_t ones_col = FF / 2 + FF / 4;                    // bias / time-token operand chunk (8 packed columns)
    constexpr uint32_t kZCol = 32;                                    // MLP2 / logit accumulator: over consumed MLP1 columns
    uint32_t ph_mma = 0, ph_kv = 0;
    const int bar_tg = 1 + tg;
    const int bar_ln = 1 + NT + 4 * tg + qd;                          // named barrier of this row-pair of warps (64 threads)
    static_assert(1 + NT + 4 * NT <= 16, "named barrier ids");

#ifdef ALINE_Q4_TRACE
    const bool trace_on = blockIdx.x == 1 && (tid == 0 || tid == 32 * 5);
    int trace_n = tid == 0 ? 0 : 4096;
#endif
    // hand the operands of the next MMA phase to the issuer warp and wait for the phase's completion
    // kTmemOps: the operands were written with tcgen05.st (P, relu(F)) -> wait for those stores; otherwise they were
    // written to the shared-memory operand tile -> make them visible to the async proxy.  (Each of the two costs
    // ~140 cycles even with nothing outstanding; no phase needs both.)
    auto mma_phase = [&](auto tmem_ops) {
        Q4_T(0);
        if constexpr (decltype(tmem_ops)::value) tc::tmem_st_wait();
        Q4_T(1);
        if constexpr (!decltype(tmem_ops)::value) tc::fence_async_smem();
        tc::tc_fence_before();
        Q4_T(2);
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_rdy[tg]);
        Q4_T(3);
        tc::mbar_wait(&bar_mma[tg], ph_mma);
        ph_mma ^= 1;
        tc::tc_fence_after();
        Q4_T(5);
    };
    // relu + bf16 pack IN PLACE of this thread's 64 accumulator columns: [f_in + 32 j, +32) -> [f_in + 16 j, +16)
    auto relu_pack_half = [&](int n_half) {
        if constexpr (kLean) {
            float fa[16];
            for (int blk = 0; blk < n_half / 16; ++blk) {
                tc::tmem_ld16(tl + f_in + 16 * blk, fa);
                tc::tmem_ld_wait16(fa);
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pk[i] = pack2_relu(fa[2 * i], fa[2 * i + 1]);
                tc::tmem_st8(tl + f_in + 8 * blk, pk);
            }
        } else {
            float fa[32];
            for (int blk = 0; blk < n_half / 32; ++blk) {
                tc::tmem_ld32(tl + f_in + 32 * blk, fa);
                tc::tmem_ld_wait32(fa);
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack2_relu(fa[2 * i], fa[2 * i + 1]);
                tc::tmem_st16(tl + f_in + 16 * blk, pk);
            }
        }
        if (c == 1) {
            uint32_t ones_pk[8];
            ones_pk[0] = pack2(1.f, 1.f); ones_pk[1] = pack2(t_hi, t_lo);
#pragma unroll
            for (int i = 2; i < 8; ++i) ones_pk[i] = 0u;
            tc::tmem_st8(tl + ones_col, ones_pk);
        }
    };

    constexpr std::integral_constant<bool, true> kTmemOps{};
    constexpr std::integral_constant<bool, false> kSmemOps{};
    int b_loaded = -1;
    bool bad = false;

    for (int tile = t0 + tg; tile < t1; tile += NT) {
        const int b = tile / tiles_per_b;
        const int j = (tile - b * tiles_per_b) * kT2Tile + r;
        const bool in_range = j < nq;
        const bool live = in_range && (alive == nullptr || alive[(size_t)b * nq + j] != 0);
        const bool new_kv = b != b_loaded;
        if (new_kv) {
            tc::named_sync(bar_tg, 256);                                // every thread of the group has left the previous tile
            if (tt == 0) {
                tc::mbar_arrive_expect_tx(&bar_kv[tg], (uint32_t)(S.NL * kvblk));
                for (int l = 0; l < S.NL; ++l)
                    tc::bulk_g2s(KVb + (size_t)l * kvblk, tckv + ((size_t)l * B + b) * kvblk, (uint32_t)kvblk, &bar_kv[tg]);
            }
        }
        float x[16];
        {
            const float4* pe = reinterpret_cast<const float4*>(eq_rm + ((size_t)b * nq + (live ? j : 0)) * kD + 16 * c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = __ldg(pe + i);
                x[4 * i] = live ? v.x : 0.f; x[4 * i + 1] = live ? v.y : 0.f;
                x[4 * i + 2] = live ? v.z : 0.f; x[4 * i + 3] = live ? v.w : 0.f;
            }
        }
        if (new_kv) {
            tc::mbar_wait(&bar_kv[tg], ph_kv);
            ph_kv ^= 1;
            b_loaded = b;
        }

        float y[16];
        for (int l = 0; l < S.NL; ++l) {
            const float* V = Vec + l * kVecLayer + 16 * c;
            float2* ln_a = Ln + (size_t)(0 * 2) * kT2Tile;              // exchange slots of LN1 / LN2
            float2* ln_b = Ln + (size_t)(1 * 2) * kT2Tile;
            // ---- Q ----
            store_chunk(Xt, 2 * c, r, x);
            store_chunk(Xt, 2 * c + 1, r, x + 8);
            mma_phase(kSmemOps);
            tc::tmem_ld16(tl + 16 * c, y);
            tc::tmem_ld_wait16(y);
            store_chunk(Xt, 2 * c, r, y);
            store_chunk(Xt, 2 * c + 1, r, y + 8);
            // ---- S = Q_h (K_h - K_0h)^T + mask, all heads ----
            mma_phase(kSmemOps);
            // ---- P = 2^S for this thread's two heads, packed to bf16 in place over its own score columns ----
            {
                const uint32_t s_in = tl + (uint32_t)(2 * c * nkp);
                const int nblk = 2 * nkp / 16;                           // 16-column blocks of the two heads (even)
                if constexpr (kLean) {
                    float sa[16];
                    for (int blk = 0; blk < nblk; ++blk) {
                        tc::tmem_ld16(s_in + 16 * blk, sa);
                        tc::tmem_ld_wait16(sa);
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = pack2(ex2f(sa[2 * i]), ex2f(sa[2 * i + 1]));
                        tc::tmem_st8(s_in + 8 * blk, pk);
                    }
                } else {
                    float sa[16], sb[16];
                    tc::tmem_ld16(s_in, sa);
                    for (int blk = 0; blk < nblk; blk += 2) {
                        tc::tmem_ld_wait16(sa);
                        tc::tmem_ld16(s_in + 16 * (blk + 1), sb);
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = pack2(ex2f(sa[2 * i]), ex2f(sa[2 * i + 1]));
                        tc::tmem_st8(s_in + 8 * blk, pk);
                        tc::tmem_ld_wait16(sb);
                        if (blk + 2 < nblk) tc::tmem_ld16(s_in + 16 * (blk + 2), sa);
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = pack2(ex2f(sb[2 * i]), ex2f(sb[2 * i + 1]));
                        tc::tmem_st8(s_in + 8 * (blk + 1), pk);
                    }
                }
            }
            mma_phase(kTmemOps);                                       // PV
            // ---- o = PV / denominator for heads 2c, 2c+1 ----
            {
                float pv[32];
                tc::tmem_ld32(tl + pv_col_hd(2 * c, nkp, TM), pv);
                tc::tmem_ld_wait32(pv);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const float den = pv[16 * hh + 8];
                    bad |= !(den < 1e30f);
                    const float inv = __fdividef(1.0f, den);
                    const f32x2 inv2 = pk2(inv, inv);
                    float o8[8];
#pragma unroll
                    for (int i = 0; i < 8; i += 2)
                        upk2(mul2(pk2(pv[16 * hh + i], pv[16 * hh + i + 1]), inv2), o8[i], o8[i + 1]);
                    store_chunk(Xt, 2 * c + hh, r, o8);
                }
            }
            // ---- y = [o | 1] Wo'^T ; h = LN1(x + y) ----
            mma_phase(kSmemOps);
            tc::tmem_ld16(tl + 16 * c, y);
            tc::tmem_ld_wait16(y);
            add_ln_half(x, y, V, V + kD, ln_a + c * kT2Tile + r, ln_a + (1 - c) * kT2Tile + r, bar_ln, 64);
            // ---- f = relu([h | 1] W1'^T) ----
            store_chunk(Xt, 2 * c, r, x);
            store_chunk(Xt, 2 * c + 1, r, x + 8);
            mma_phase(kSmemOps);
            relu_pack_half(FF / 2);
            // ---- z = [1 | f] W2'^T ; x' = LN2(h + z) ----
            mma_phase(kTmemOps);
            tc::tmem_ld16(tl + kZCol + 16 * c, y);
            tc::tmem_ld_wait16(y);
            add_ln_half(x, y, V + 2 * kD, V + 3 * kD, ln_b + c * kT2Tile + r, ln_b + (1 - c) * kT2Tile + r, bar_ln, 64);
        }
        // ---- acquisition MLP: logit = [b2 | w2] . [1 | relu([z | 1, t] Wa'^T)], both layers on the tensor pipe ----
        store_chunk(Xt, 2 * c, r, x);
        store_chunk(Xt, 2 * c + 1, r, x + 8);
        mma_phase(kSmemOps);
        relu_pack_half(HH / 2);
        mma_phase(kTmemOps);
        if (c == 0) {
            float l0, l1;
            tmem_ld2(tl + kZCol, l0, l1);
            tc::tmem_ld_wait();
            if (in_range) logits[(size_t)b * nq + j] = live ? l0 + l1 : -INFINITY;
        }
        if (zq && live) {
            float4* z = reinterpret_cast<float4*>(zq + ((size_t)b * nq + j) * kD + 16 * c);
#pragma unroll
            for (int i = 0; i < 4; ++i) z[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        }
        if (bad && live) *flag = epoch;
        bad = false;
    }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid < 32) tc::tmem_dealloc(tmem_base_s, 512);
}

// tiles in flight per SM for a padded key count: the scores of 4 heads (4 nkp columns), the MLP accumulator (128) and
// the PV accumulators must fit in 512 / NT tensor-memory columns, the per-group K / V blocks in shared memory
static int pick_nt(const Tc2Shape& S, int nkp) {
    static const int want = [] { const char* e = getenv("ALINE_Q4_NT"); return e ? atoi(e) : 3; }();
    int nt = nkp <= 32 ? want : 2;
    if (nt < 2 || nt > 3) nt = 3;
    if (nkp > 32) nt = 2;
    while (nt > 2 && make_carve(S.total_bytes, S.vec_total, S.NL, nkp, nt).total + 256 > (size_t)device_info().max_smem_optin) --nt;
    return nt;
}

bool supported(const Dims& d, int n_keys) {
    if (d.D != kD || d.FF != kFF || d.HH != kHH) return false;
    if (n_keys < 1 || n_keys > 48) return false;
    Tc2Shape S = make_tc2_shape(d);
    const int nkp = (n_keys + 15) / 16 * 16;
    return make_carve(S.total_bytes, S.vec_total, S.NL, nkp, 2).total + 256 <= (size_t)device_info().max_smem_optin;
}

int launch(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq_rm, const unsigned char* alive,
           int B, int nq, int n_keys, float t_value, float* logits, float* zq, const void* tckv, int* flag, int epoch,
           cudaStream_t st) {
    ALINE_REQUIRE(supported(d, n_keys), "two-thread tensor-core query stream: unsupported shape (d=%d ff=%d head=%d keys=%d)",
                  d.D, d.FF, d.HH, n_keys);
    Tc2Shape S = make_tc2_shape(d);
    const int nkp = (n_keys + 15) / 16 * 16;
    const int nt = pick_nt(S, nkp);
    const size_t smem = make_carve(S.total_bytes, S.vec_total, S.NL, nkp, nt).total;
    const int tiles = ceil_div(nq, kT2Tile);
    const long long n_tiles = (long long)B * tiles;
    int grid = device_info().sm_count;
    if ((long long)grid * nt > n_tiles) grid = (int)((n_tiles + nt - 1) / nt);
    const __nv_bfloat16 th = __float2bfloat16_rn(t_value);
    const float t_hi = __bfloat162float(th), t_lo = t_value - t_hi;
#define ALINE_Q4_LAUNCH(NTV)                                                                                           \
    do {                                                                                                               \
        if (ensure_dyn_smem((const void*)query_tc4_kernel<NTV>, smem)) return 1;                                       \
        ALINE_CHECK_CUDA(launch_k(query_tc4_kernel<NTV>, dim3(grid), dim3(288 * NTV), smem, st, g_pdl_chain, d, L, S, P, \
                                  (const unsigned char*)wb2, eq_rm, alive, nq, B, t_hi, t_lo, logits, zq, tiles,       \
                                  (const unsigned char*)tckv, nkp, flag, epoch));                                      \
    } while (0)
    if (nt == 3) ALINE_Q4_LAUNCH(3);
    else ALINE_Q4_LAUNCH(2);
#undef ALINE_Q4_LAUNCH
    ALINE_LAUNCH_OK();
    return 0;
}

}  // namespace tc4

#ifdef ALINE_Q4_TRACE
extern "C" int aline_debug_q4_trace(long long* host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, tc4::g_q4_trace, sizeof(long long) * (size_t)n) == cudaSuccess ? 0 : 1;
}
#endif

bool query_tc4_supported(const Dims& d, int n_keys) { return tc4::supported(d, n_keys); }

int query_stream_tc4(const Dims& d, const Layout& L, const float* P, const void* wb2, const float* eq_rm,
                     const unsigned char* alive, int B, int nq, int n_keys, float t_value, float* logits, float* zq,
                     const void* tckv, int* flag, int epoch, cudaStream_t st) {
    return tc4::launch(d, L, P, wb2, eq_rm, alive, B, nq, n_keys, t_value, logits, zq, tckv, flag, epoch, st);
}

}  // namespace aline
