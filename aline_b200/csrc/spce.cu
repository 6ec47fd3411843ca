// sPCE / sNMC evaluation: streaming likelihood + online log-sum-exp over the
// contrastive prior draws.
//
// Replaces, in the reference:
//   loss/eig.py:154-209   EIGStepLoss.step / .forward   -> aline_spce_step
//   loss/eig.py:22-48     EIGBounds.compute_seq_logprobs -> aline_spce_history
//   utils/eval.py:43-80   compute_EIG_from_history       -> aline_spce_history + aline_lse_combine
//
// Layout.  thetas [n_rows, B, dth] and seq [n_rows, B] are the reference's own
// tensors (row = contrastive draw l, column = trajectory b).  One thread owns
// one column b and walks rows l with a grid stride, so its running
// (max, sum-exp) pairs -- one per history point in the current pass -- live in
// registers for the whole kernel; consecutive threads read consecutive b, i.e.
// consecutive addresses of both tensors.  The per-(b,t) history record H (y,
// design, and whatever part of the likelihood does not depend on theta) is
// prepared once by a tiny kernel in [t][f][b] layout and, for the cheap
// likelihoods, held in registers across all rows.
//
// A pass covers up to TC history points: theta is read once per pass and
// seq (the accumulated log-likelihood) is read / written once per pass, so the
// traffic per (l, b) is  4*dth + 8  bytes per pass instead of per history
// point.  TC = 1 is the drop-in EIGStepLoss.step (HBM-bound); TC = 16 is the
// fused-history evaluation (issue/MUFU-bound).
#include "lik.cuh"
#include "tc.cuh"
#include "philox.cuh"
#include <cstdlib>

namespace aline {

constexpr int kMaxGridX = 640;        // upper bound used for scratch sizing
constexpr int kMaxColsPerBlock = 512;
constexpr int kMaxNH = 24;
constexpr int kMaxPass = 36;        // history points per pass (largest compiled TC)

template <class LK> struct is_ces : std::false_type {};
template <bool F> struct is_ces<CesLikT<F>> : std::true_type {};

// ---------------------------------------------------------------- H prep ----
template <class LK>
__global__ void prep_hist_generic(const float* __restrict__ y, const float* __restrict__ xi, int B, int T,
                                  int dim_x, float* __restrict__ H) {
    // location / psychometric records: H[t][0][b] and H[t][1+d][b]
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    int b = i / T, t = i % T;
    float yy = y[i];
    const float* x = xi + (size_t)i * dim_x;
    float* h = H + (size_t)t * LK::NH * B + b;
    if constexpr (std::is_same<LK, PsychometricLik>::value) {
        h[0] = x[0];
        h[(size_t)B] = yy;
    } else {
        h[0] = yy;
        for (int d = 0; d < dim_x; ++d) h[(size_t)(1 + d) * B] = x[d];
    }
}

__global__ void prep_hist_ces(const float* __restrict__ y, const float* __restrict__ xi, int B, int T,
                              float epsilon, float* __restrict__ H) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    int b = i / T, t = i % T;
    float rec[CesLik::NH];
    ces_prepare(xi + (size_t)i * 6, y[i], epsilon, rec);
    float* h = H + (size_t)t * CesLik::NH * B + b;
#pragma unroll
    for (int f = 0; f < CesLik::NH; ++f) h[(size_t)f * B] = rec[f];
}

// ------------------------------------------------------- streaming kernel ----
// grid.x: row groups (persistent, grid-stride);  grid.y: column chunks of CB columns.
// block = RS x CB threads: thread (r, c) owns column b = blockIdx.y*CB + c and rows
// l = row_begin + blockIdx.x*RS + r + k * gridDim.x*RS  (l < row_end).
// HOT = true: every row is a contrastive draw (no theta_0 handling in the loop);
// HOT = false: the rows are the leading non-contrastive rows (row 0 = theta_0): only out_lp0 / seq are produced.
// Shared memory: the pass' history records Hs[t][f][c] (read conflict-free: consecutive threads, consecutive c),
// then reused for the block-level merge of the per-thread (max, sum-exp) pairs.
// register budget per thread ~ 2*TC (m,s pairs) + ~40: cap the block size accordingly
constexpr int max_threads_for(int TC) { return TC > 18 ? 448 : (TC > 12 ? 640 : 1024); }

template <class LK, int TC, int U, bool HOT>
__global__ void __launch_bounds__(max_threads_for(TC))
spce_stream_kernel(const LK lk, const float* __restrict__ H, int t0, int nT, int Ttot,
                   const float* __restrict__ thetas, int dth, float* __restrict__ seq,
                   long long row_begin, long long row_end, int B, int CB, int RS, int read_seq, int write_seq,
                   float2* __restrict__ part, float* __restrict__ out_lp0, int* __restrict__ bad_flag,
                   const int* __restrict__ run_flag) {
    extern __shared__ float smem[];
    if (run_flag && *run_flag == 0) return;       // conditional re-run after the fast pass: nothing to redo
    const int tid = threadIdx.x;
    const int r = tid / CB, c = tid - r * CB;
    const int b = blockIdx.y * CB + c;
    const bool active = (r < RS) && (b < B);

    // stage this pass' history records for the block's columns as [t][c][NHP]: the fields of a thread's record are
    // NHP = NH | 1 floats apart from the next column's (odd: consecutive columns hit distinct banks) and at IMMEDIATE
    // offsets from one per-point base -- the former [t][f][c] layout cost an address add per field (CES: 15 IADD3 of 208
    // instructions per evaluation)
    constexpr int NHP = LK::NH | 1;
    for (int i = tid; i < nT * LK::NH * CB; i += blockDim.x) {
        int tf = i / CB, cc = i - tf * CB;
        int bb = blockIdx.y * CB + cc;
        const int t = tf / LK::NH, f = tf - t * LK::NH;
        smem[(t * CB + cc) * NHP + f] = (bb < B) ? __ldg(H + ((size_t)t0 * LK::NH + tf) * B + bb) : 0.f;
    }
    __syncthreads();

    Lse acc[TC];
#pragma unroll
    for (int t = 0; t < TC; ++t) acc[t].init();

    if (active) {
        // history records: registers when the pass is short (no shared-memory traffic in the hot loop)
        constexpr bool kHRegs = TC * LK::NH <= 12;
        const float* hs = smem + c * NHP;
        const int t_stride = CB * NHP;
        float hreg[kHRegs ? TC : 1][LK::NH];
        if constexpr (kHRegs) {
#pragma unroll
            for (int t = 0; t < TC; ++t)
#pragma unroll
                for (int f = 0; f < LK::NH; ++f) hreg[t][f] = (t < nT) ? hs[t * t_stride + f] : 0.f;
        }
        bool bad = false;
        // this thread's rows: l = first + k*stride, k = 0 .. n_mine-1; pointers advance by a constant
        const long long stride = (long long)gridDim.x * RS;
        const long long first = row_begin + (long long)blockIdx.x * RS + r;
        const long long n_mine = first < row_end ? (row_end - first + stride - 1) / stride : 0;
        const float* pth = thetas + ((size_t)first * B + b) * dth;
        float* pseq = seq + ((size_t)first * B + b);
        const size_t th_step = (size_t)stride * B * dth, seq_step = (size_t)stride * B;

        auto eval_row = [&](const typename LK::Theta& th, float s_run, float* seq_out, bool is_row0) {
#pragma unroll
            for (int t = 0; t < TC; ++t) {
                if (t < nT) {
                    float v;
                    if constexpr (kHRegs) {
                        v = lk.ll(th, hreg[t]);
                    } else {
                        float hl[LK::NH];
                        const float* ht = hs + t * t_stride;
#pragma unroll
                        for (int f = 0; f < LK::NH; ++f) hl[f] = ht[f];
                        v = lk.ll(th, hl);
                    }
                    if constexpr (LK::CHECK_BAD) bad |= !isfinite(v);
                    s_run += v;
                    if constexpr (HOT) acc[t].push(s_run);
                    else if (is_row0 && out_lp0) out_lp0[(size_t)b * Ttot + t0 + t] = s_run;
                }
            }
            if (write_seq) *seq_out = s_run;
        };

        // software pipeline over groups of U rows: group k+1 is loaded before group k is evaluated, so the HBM
        // latency of theta / seq overlaps the likelihood arithmetic
        typename LK::Theta th[U], th_n[U];
        float S[U], S_n[U];
        const long long n_full = n_mine / U;
        if (n_full > 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                lk.load_theta(th[u], pth + u * th_step);
                S[u] = read_seq ? ld_stream1(pseq + u * seq_step) : 0.f;
            }
        }
        for (long long g = 0; g < n_full; ++g) {
            const float* pth_n = pth + U * th_step;
            float* pseq_n = pseq + U * seq_step;
            if (g + 1 < n_full) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    lk.load_theta(th_n[u], pth_n + u * th_step);
                    S_n[u] = read_seq ? ld_stream1(pseq_n + u * seq_step) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                eval_row(th[u], S[u], pseq + u * seq_step, !HOT && first + (g * U + u) * stride == 0);
#pragma unroll
            for (int u = 0; u < U; ++u) { th[u] = th_n[u]; S[u] = S_n[u]; }
            pth = pth_n; pseq = pseq_n;
        }
        for (long long k = n_full * U; k < n_mine; ++k) {       // tail rows
            typename LK::Theta t1;
            lk.load_theta(t1, pth);
            float s1 = read_seq ? ld_stream1(pseq) : 0.f;
            eval_row(t1, s1, pseq, !HOT && first + k * stride == 0);
            pth += th_step; pseq += seq_step;
        }
        if constexpr (LK::CHECK_BAD) {
            if (bad && bad_flag) atomicOr(bad_flag, 1);
        }
    }
    if constexpr (!HOT) return;

    // merge the RS row-threads of each column, one history point at a time
    __syncthreads();
    float2* sh = reinterpret_cast<float2*>(smem);
#pragma unroll 1
    for (int t = 0; t < nT; ++t) {
        float2 mine;
#pragma unroll
        for (int tt = 0; tt < TC; ++tt)
            if (tt == t) mine = make_float2(acc[tt].m, acc[tt].s);
        sh[tid] = mine;
        __syncthreads();
        if (r == 0 && b < B) {
            Lse a; a.m = mine.x; a.s = mine.y;
            for (int rr = 1; rr < RS; ++rr) {
                float2 o = sh[rr * CB + c];
                a.merge(o.x, o.y);
            }
            part[((size_t)blockIdx.x * Ttot + t0 + t) * B + b] = make_float2(a.m, a.s);
        }
        __syncthreads();
    }
}

// ---- EIGStepLoss.step, location K = 1, D = 2: the HBM-bound case gets its own lean kernel ----
// Thread (r, c2) owns the column pair b = 2*c2, 2*c2+1 (theta of both = one 16-byte load, seq = one 8-byte
// load / store) and walks rows with U rows in flight, so ~24*U bytes per thread are outstanding: enough to cover
// the HBM latency at ~40 % occupancy without software pipelining.
template <int U>
__global__ void __launch_bounds__(512, 2)
spce_step_loc12_kernel(const LocationLik<1, 2> lk, const float* __restrict__ H, const float* __restrict__ thetas,
                       float* __restrict__ seq, long long row_begin, long long row_end, int B, int CB2, int RS,
                       float2* __restrict__ part) {
    extern __shared__ float smem[];
    const int tid = threadIdx.x;
    const int r = tid / CB2, c2 = tid - r * CB2;
    const int b = 2 * (blockIdx.y * CB2 + c2);
    const bool active = (r < RS) && (b < B);
    Lse a0, a1;
    a0.init(); a1.init();
    if (active) {
        float h0[3], h1[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) { h0[f] = __ldg(H + (size_t)f * B + b); h1[f] = __ldg(H + (size_t)f * B + b + 1); }
        // each block streams one contiguous range of rows (DRAM page locality: RS*U consecutive rows in flight)
        const long long per_block = (row_end - row_begin + gridDim.x - 1) / gridDim.x;
        const long long blk_begin = row_begin + (long long)blockIdx.x * per_block;
        const long long blk_end = blk_begin + per_block < row_end ? blk_begin + per_block : row_end;
        const long long stride = RS;
        const long long first = blk_begin + r;
        const long long n_mine = first < blk_end ? (blk_end - first + stride - 1) / stride : 0;
        const float4* pth = reinterpret_cast<const float4*>(thetas + ((size_t)first * B + b) * 2);
        float2* pseq = reinterpret_cast<float2*>(seq + (size_t)first * B + b);
        const size_t th_step = (size_t)stride * B / 2, seq_step = (size_t)stride * B / 2;     // in float4 / float2 units
        auto load_group = [&](const float4* pt, const float2* ps, float4* th, float2* sv) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                th[u] = ldg_stream4(reinterpret_cast<const float*>(pt + u * th_step));
                asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(sv[u].x), "=f"(sv[u].y) : "l"(ps + u * seq_step));
            }
        };
        auto eval_pair = [&](const float4& t4, const float2& s2, float2* dst) {
            LocationLik<1, 2>::Theta t0, t1;
            t0.v[0] = t4.x; t0.v[1] = t4.y; t1.v[0] = t4.z; t1.v[1] = t4.w;
            float s0 = s2.x + lk.ll(t0, h0), s1 = s2.y + lk.ll(t1, h1);
            a0.push(s0); a1.push(s1);
            *dst = make_float2(s0, s1);
        };
        // groups of U rows, the next group's loads issued before the current group is evaluated
        const long long n_groups = n_mine / U;
        float4 th[U], th_n[U];
        float2 sv[U], sv_n[U];
        if (n_groups > 0) load_group(pth, pseq, th, sv);
        for (long long g = 0; g < n_groups; ++g) {
            const float4* pth_n = pth + U * th_step;
            float2* pseq_n = pseq + U * seq_step;
            if (g + 1 < n_groups) load_group(pth_n, pseq_n, th_n, sv_n);
#pragma unroll
            for (int u = 0; u < U; ++u) eval_pair(th[u], sv[u], pseq + u * seq_step);
#pragma unroll
            for (int u = 0; u < U; ++u) { th[u] = th_n[u]; sv[u] = sv_n[u]; }
            pth = pth_n; pseq = pseq_n;
        }
        for (long long k = n_groups * U; k < n_mine; ++k) {
            float4 t4 = ldg_stream4(reinterpret_cast<const float*>(pth));
            float2 s2 = *pseq;
            eval_pair(t4, s2, pseq);
            pth += th_step; pseq += seq_step;
        }
    }
    float4* sh = reinterpret_cast<float4*>(smem);
    sh[tid] = make_float4(a0.m, a0.s, a1.m, a1.s);
    __syncthreads();
    if (r == 0 && b < B) {
        for (int rr = 1; rr < RS; ++rr) {
            float4 o = sh[rr * CB2 + c2];
            a0.merge(o.x, o.y); a1.merge(o.z, o.w);
        }
        part[(size_t)blockIdx.x * B + b] = make_float2(a0.m, a0.s);
        part[(size_t)blockIdx.x * B + b + 1] = make_float2(a1.m, a1.s);
    }
}

// ---- EIGStepLoss.step, location K = 1, D = 2, TMA-staged (the default for B % 4 == 0, B <= 480) ----
// One persistent block per SM streams ONE contiguous range of contrastive rows.  A producer thread moves chunks of R
// rows (theta: R*B*8 contiguous bytes, seq: R*B*4) global -> shared with cp.async.bulk into an NS-stage ring; the
// consumer warps (thread = fixed column b, RS rows in parallel) add the log-likelihood to seq IN PLACE in shared memory
// and keep the online (max, sum-exp) pair in registers; the producer writes the updated seq chunk back with a bulk
// store and refills the stage.  No register-staged global loads, no per-thread address arithmetic, every DRAM request
// is a full contiguous burst.  Block 0 also advances the theta_0 row (out_lp0), and the last block to finish merges
// the per-block partials (ticket counter), so the whole step is one launch.
constexpr int kStepMaxStages = 8;

__global__ void __launch_bounds__(512, 1)
spce_step_tma_kernel(const LocationLik<1, 2> lk, const float* __restrict__ y, const float* __restrict__ xi,
                     const float* __restrict__ thetas, float* __restrict__ seq, long long skip_rows, long long n_rows,
                     int B, int R, int RS, int NS, float2* __restrict__ part, float* __restrict__ out_lp0,
                     unsigned int* __restrict__ ticket, float* __restrict__ out_m, float* __restrict__ out_s) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t full[kStepMaxStages], done[kStepMaxStages];
    __shared__ int is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_cons = RS * B, n_cons_warps = (n_cons + 31) >> 5;
    const uint32_t th_bytes = (uint32_t)R * B * 8u, stage_bytes = (uint32_t)R * B * 12u;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&done[i], n_cons_warps); }
        tc::fence_mbar_init();
    }
    __syncthreads();
    const long long total = n_rows - skip_rows;
    long long per_block = (total + gridDim.x - 1) / gridDim.x;
    per_block = (per_block + R - 1) / R * R;                       // whole chunks: only the last block has a short chunk
    long long blk_begin = skip_rows + (long long)blockIdx.x * per_block;
    if (blk_begin > n_rows) blk_begin = n_rows;
    const long long blk_end = blk_begin + per_block < n_rows ? blk_begin + per_block : n_rows;
    const int n_chunks = (int)((blk_end - blk_begin + R - 1) / R);

    Lse a0, a1;
    a0.init(); a1.init();
    if (warp == n_cons_warps) {
        // ---- producer ----
        if (lane == 0) {
            auto load = [&](int c) {
                const int st = c % NS;
                const long long row0 = blk_begin + (long long)c * R;
                const uint32_t rows = (uint32_t)(blk_end - row0 < R ? blk_end - row0 : R);
                unsigned char* sb = ring + (size_t)st * stage_bytes;
                tc::mbar_arrive_expect_tx(&full[st], rows * (uint32_t)B * 12u);
                tc::bulk_g2s(sb, thetas + (size_t)row0 * B * 2, rows * (uint32_t)B * 8u, &full[st]);
                tc::bulk_g2s(sb + th_bytes, seq + (size_t)row0 * B, rows * (uint32_t)B * 4u, &full[st]);
            };
            for (int c = 0; c < NS && c < n_chunks; ++c) load(c);
            for (int c = 0; c < n_chunks; ++c) {
                const int st = c % NS;
                tc::mbar_wait(&done[st], (uint32_t)((c / NS) & 1));
                const long long row0 = blk_begin + (long long)c * R;
                const uint32_t rows = (uint32_t)(blk_end - row0 < R ? blk_end - row0 : R);
                tc::bulk_s2g(seq + (size_t)row0 * B, ring + (size_t)st * stage_bytes + th_bytes, rows * (uint32_t)B * 4u);
                tc::bulk_commit();
                if (c + NS < n_chunks) {
                    tc::bulk_wait_read0();
                    load(c + NS);
                }
            }
            tc::bulk_wait0();
        }
    } else {
        // ---- consumers (the unused lanes of the last consumer warp only take part in its arrivals) ----
        const bool active = tid < n_cons;
        const int r = active ? tid / B : 0, b = active ? tid - r * B : 0;
        float h[3];
        h[0] = __ldg(y + b); h[1] = __ldg(xi + 2 * b); h[2] = __ldg(xi + 2 * b + 1);
        if (blockIdx.x == 0 && active && r == 0) {                     // theta_0 row(s): advance seq, report lp0
            for (long long row = 0; row < skip_rows; ++row) {
                LocationLik<1, 2>::Theta t;
                const float2 tv = *reinterpret_cast<const float2*>(thetas + ((size_t)row * B + b) * 2);
                t.v[0] = tv.x; t.v[1] = tv.y;
                const float s = seq[(size_t)row * B + b] + lk.ll(t, h);
                seq[(size_t)row * B + b] = s;
                if (row == 0 && out_lp0) out_lp0[b] = s;
            }
        }
        for (int c = 0; c < n_chunks; ++c) {
            const int st = c % NS;
            const long long row0 = blk_begin + (long long)c * R;
            const int rows = (int)(blk_end - row0 < R ? blk_end - row0 : R);
            const float2* sth = reinterpret_cast<const float2*>(ring + (size_t)st * stage_bytes);
            float* ssq = reinterpret_cast<float*>(ring + (size_t)st * stage_bytes + th_bytes);
            tc::mbar_wait(&full[st], (uint32_t)((c / NS) & 1));
            auto one = [&](int rr, Lse& a) {
                const int i = rr * B + b;
                const float2 tv = sth[i];
                LocationLik<1, 2>::Theta t;
                t.v[0] = tv.x; t.v[1] = tv.y;
                const float s = ssq[i] + lk.ll(t, h);
                a.push(s);
                ssq[i] = s;
            };
            if (active) {
                if (rows == R) {
#pragma unroll 4
                    for (int rr = r; rr < R; rr += 2 * RS) { one(rr, a0); one(rr + RS, a1); }     // R % (2 RS) == 0
                } else {
                    for (int rr = r; rr < rows; rr += RS) one(rr, a0);
                }
                tc::fence_async_smem();                                 // in-place results -> visible to the bulk store
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&done[st]);
        }
    }
    __syncthreads();
    float4* sh = reinterpret_cast<float4*>(ring);
    a0.merge(a1.m, a1.s);
    if (tid < n_cons) sh[tid] = make_float4(a0.m, a0.s, 0.f, 0.f);
    __syncthreads();
    if (tid < B) {
        for (int rr = 1; rr < RS; ++rr) { const float4 o = sh[rr * B + tid]; a0.merge(o.x, o.y); }
        part[(size_t)blockIdx.x * B + tid] = make_float2(a0.m, a0.s);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int b = tid; b < B; b += blockDim.x) {
            Lse a; a.init();
            for (int g = 0; g < (int)gridDim.x; ++g) {
                const float2 o = __ldcg(part + (size_t)g * B + b);
                a.merge(o.x, o.y);
            }
            out_m[b] = a.m;
            out_s[b] = a.s;
        }
    }
}

// ---- fast history pass: shifted accumulation against a fixed per-(b,t) reference ----
// With M_t = seq-log-likelihood of theta_0 after history point t (known from the cold pass) the contrastive sum is
// accumulated as  s_t = sum_l 2^(S2_t[l]),  S2_t = (S_t - M_t) * log2 e,  i.e. one register, one FADD and one EX2 per
// evaluation instead of the online (max, sum) pair.  S2 is advanced by  ll_t * log2 e - (M_t - M_{t-1}) * log2 e; the
// second term (and the constant of the Normal density) is the per-(b,t) record field c2.  Terms more than ~87 nats
// below theta_0's flush to zero (irrelevant for sPCE, which contains theta_0's own term); the finalize kernel
// raises a flag if a sum is 0 or not finite, and the robust kernels then recompute the bound.
template <class LK>
__global__ void prep_fast_kernel(const LK lk, const float* __restrict__ H, const float* __restrict__ lp0, int B, int T,
                                 float* __restrict__ HF) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    int t = i / B, b = i - t * B;
    constexpr int NF = LK::NH + 1;
#pragma unroll
    for (int f = 0; f < LK::NH; ++f) HF[((size_t)t * NF + f) * B + b] = H[((size_t)t * LK::NH + f) * B + b];
    const float m_t = lp0[(size_t)b * T + t], m_prev = t > 0 ? lp0[(size_t)b * T + t - 1] : 0.f;
    HF[((size_t)t * NF + LK::NH) * B + b] = lk.const_log2() - (m_t - m_prev) * 1.44269504088896340736f;
}

template <class LK, int TC, bool FULL>      // FULL: the pass covers exactly TC history points (no per-point guard)
__global__ void __launch_bounds__(1024)
spce_fast_kernel(const LK lk, const float* __restrict__ HF, int t0, int nT, int Ttot, const float* __restrict__ thetas,
                 int dth, float* __restrict__ seq, long long row_begin, long long row_end, int B, int CB, int RS,
                 int read_seq, int write_seq, float* __restrict__ part) {
    extern __shared__ float smem[];
    constexpr int NF = LK::NH + 1;
    const int tid = threadIdx.x;
    const int r = tid / CB, c = tid - r * CB;
    const int b = blockIdx.y * CB + c;
    const bool active = (r < RS) && (b < B);
    float acc[TC];
#pragma unroll
    for (int t = 0; t < TC; ++t) acc[t] = 0.f;
    if (active) {
        float h[TC][NF];
#pragma unroll
        for (int t = 0; t < TC; ++t)
#pragma unroll
            for (int f = 0; f < NF; ++f) h[t][f] = (t < nT) ? __ldg(HF + ((size_t)(t0 + t) * NF + f) * B + b) : 0.f;
        const long long stride = (long long)gridDim.x * RS;
        const long long first = row_begin + (long long)blockIdx.x * RS + r;
        const long long n_mine = first < row_end ? (row_end - first + stride - 1) / stride : 0;
        const float* pth = thetas + ((size_t)first * B + b) * dth;
        float* pseq = seq + ((size_t)first * B + b);
        const size_t th_step = (size_t)stride * B * dth, seq_step = (size_t)stride * B;
        typename LK::Theta th, th_n;
        float S2 = 0.f, S2_n = 0.f;
        if (n_mine > 0) {
            lk.load_theta(th, pth);
            S2 = read_seq ? ld_stream1(pseq) : 0.f;
        }
        for (long long k = 0; k < n_mine; ++k) {
            if (k + 1 < n_mine) {                                  // prefetch the next row
                lk.load_theta(th_n, pth + th_step);
                S2_n = read_seq ? ld_stream1(pseq + seq_step) : 0.f;
            }
#pragma unroll
            for (int t = 0; t < TC; ++t) {
                if (FULL || t < nT) {
                    S2 += lk.ll_log2(th, h[t], h[t][LK::NH]);
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(S2));
                    acc[t] += e;
                }
            }
            if (write_seq) *pseq = S2;
            th = th_n; S2 = S2_n;
            pth += th_step; pseq += seq_step;
        }
    }
    // block-level sums over the RS row-threads of each column
    for (int t = 0; t < nT; ++t) {
        float mine = 0.f;
#pragma unroll
        for (int tt = 0; tt < TC; ++tt)
            if (tt == t) mine = acc[tt];
        __syncthreads();
        smem[tid] = mine;
        __syncthreads();
        if (r == 0 && b < B) {
            float a = mine;
            for (int rr = 1; rr < RS; ++rr) a += smem[rr * CB + c];
            part[((size_t)blockIdx.x * Ttot + t0 + t) * B + b] = a;
        }
    }
}

// ---- fast history pass, location K = 1, D = 2, packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2) ----
// Same contract and same shifted accumulation as spce_fast_kernel; two consecutive history points (t, t+1) of one
// contrastive row are evaluated as the two halves of 64-bit packed operands, which halves the issue slots of the
// FMA-pipe work (the scalar kernel is issue-bound: ~19 instructions per evaluation for 2 MUFU).  Per pair:
// 13 packed FMA-pipe instructions + 2 scalar FADD (the running sum S2 is sequential in t) + 2 integer seeds + 4 MUFU
// (2 lg2, 2 ex2).  The reciprocal runs on the FMA pipe on the NEGATED value (seed 0xFEF311C7 - bits(x), three Newton
// steps nr <- nr + nr (1 + x nr)), so that base + 1/x comes out negated and the sign is absorbed by MUFU.LG2's
// operand modifier.
struct GenArgs {                        // in-kernel prior draws (GEN): Philox key, global row offset, box prior of (theta_0, theta_1)
    uint32_t k0, k1;
    long long row_offset;
    float lo0, sc0, lo1, sc1;
};

// TP pairs = 2 TP history points per pass, NM of them with the MUFU reciprocal (NM == kJointRcp: one MUFU reciprocal
// shared by two pairs); GEN: rows generated, not loaded
constexpr int kJointRcp = -4;
template <int TP, int NM, bool FULL, bool GEN = false>
__global__ void __launch_bounds__(608, 1)
spce_fast_loc12x2_kernel(const LocationLik<1, 2> lk, const float* __restrict__ HF, int t0, int nT, int Ttot,
                         const float* __restrict__ thetas, float* __restrict__ seq, long long row_begin,
                         long long row_end, int B, int CB, int RS, int read_seq, int write_seq, float* __restrict__ part,
                         const GenArgs gen = GenArgs{}) {
    extern __shared__ float smem[];
    constexpr int NF = 4;
    const int tid = threadIdx.x;
    const int r = tid / CB, c = tid - r * CB;
    const int b = blockIdx.y * CB + c;
    const bool active = (r < RS) && (b < B);
    f32x2 acc[TP];
#pragma unroll
    for (int p = 0; p < TP; ++p) acc[p] = pk2(0.f, 0.f);
    if (active) {
        f32x2 hy[TP], hx0[TP], hx1[TP], hc[TP];
#pragma unroll
        for (int p = 0; p < TP; ++p) {
            float v[2][NF];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int f = 0; f < NF; ++f)
                    v[q][f] = (2 * p + q < nT) ? __ldg(HF + ((size_t)(t0 + 2 * p + q) * NF + f) * B + b) : 0.f;
            hy[p] = pk2(v[0][0], v[1][0]); hx0[p] = pk2(v[0][1], v[1][1]);
            hx1[p] = pk2(v[0][2], v[1][2]); hc[p] = pk2(v[0][3], v[1][3]);
        }
        const f32x2 c_max = pk2(lk.max_signal, lk.max_signal), c_one = pk2(1.f, 1.f);
        const f32x2 c_nbase = pk2(-lk.base_signal, -lk.base_signal), c_base = pk2(lk.base_signal, lk.base_signal);
        const f32x2 c_nln2 = pk2(-0.69314718055994530942f, -0.69314718055994530942f), c_k2 = pk2(lk.k2, lk.k2);
        const long long stride = (long long)gridDim.x * RS;
        const long long first = row_begin + (long long)blockIdx.x * RS + r;
        const int n_mine = first < row_end ? (int)((row_end - first + stride - 1) / stride) : 0;   // < 2^31 rows per thread
        const float2* pth = reinterpret_cast<const float2*>(thetas + ((size_t)first * B + b) * 2);
        float* pseq = seq + ((size_t)first * B + b);
        const size_t th_step = (size_t)stride * B, seq_step = (size_t)stride * B;
        float2 th = make_float2(0.f, 0.f), th_n = th, th_nn = th;   // rows k, k+1, k+2: loads run two rows ahead
        float S2 = 0.f, S2_n = 0.f, S2_nn = 0.f;
        auto draw = [&](long long row) {                           // same function of (seed, row, b) as prior_box_kernel
            const unsigned long long g = (unsigned long long)(gen.row_offset + row);
            const Philox4 rr = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)b, 0u, gen.k0, gen.k1);
            return make_float2(fmaf(u01(rr.x[0]), gen.sc0, gen.lo0), fmaf(u01(rr.x[1]), gen.sc1, gen.lo1));
        };
        if (n_mine > 0) {
            th = GEN ? draw(first) : __ldg(pth);
            S2 = read_seq ? ld_stream1(pseq) : 0.f;
        }
        if (n_mine > 1) {
            th_n = GEN ? draw(first + stride) : __ldg(pth + th_step);
            S2_n = read_seq ? ld_stream1(pseq + seq_step) : 0.f;
        }
        for (int k = 0; k < n_mine; ++k) {
            if (k + 2 < n_mine) {                                  // prefetch (or draw) the row after the next
                th_nn = GEN ? draw(first + (long long)(k + 2) * stride) : __ldg(pth + 2 * th_step);
                S2_nn = read_seq ? ld_stream1(pseq + 2 * seq_step) : 0.f;
            }
            const f32x2 nt0 = pk2(-th.x, -th.x), nt1 = pk2(-th.y, -th.y);
            // log-density of pair p from lg2(base + 1/sq), running sum and shifted exponentials
            auto tail = [&](const int p, const float gl, const float gh) {
                const f32x2 d = fma2(c_nln2, pk2(gl, gh), hy[p]);
                float ll, lh, el, eh;
                upk2(fma2(mul2(d, d), c_k2, hc[p]), ll, lh);
                S2 += ll;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(el) : "f"(S2));
                if (FULL || 2 * p + 1 < nT) S2 += lh;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eh) : "f"(S2));
                if (FULL || 2 * p < nT) acc[p] = add2(acc[p], pk2(el, eh));
            };
            if constexpr (NM == kJointRcp) {
                // ONE MUFU reciprocal for the four evaluations of two pairs: with sqA = (a, b), sqB = (c, d) and
                // r = 1 / (ac * bd):  (1/a, 1/b) = (r bd, r ac) * sqB,  (1/c, 1/d) = (r bd, r ac) * sqA  (<= ~2.5 ulp;
                // the product stays normal for max_signal >= 1e-7, checked by the launcher).  2.25 MUFU per evaluation.
                static_assert(TP % 2 == 0, "joint reciprocal works on two pairs");
#pragma unroll
                for (int p = 0; p < TP; p += 2) {
                    const f32x2 d0a = add2(hx0[p], nt0), d1a = add2(hx1[p], nt1);
                    const f32x2 d0b = add2(hx0[p + 1], nt0), d1b = add2(hx1[p + 1], nt1);
                    const f32x2 sqa = fma2(d1a, d1a, fma2(d0a, d0a, c_max));
                    const f32x2 sqb = fma2(d1b, d1b, fma2(d0b, d0b, c_max));
                    float pl, ph, r, ta, tb, tc, td, ga, gb, gc, gd;
                    upk2(mul2(sqa, sqb), pl, ph);                  // (ac, bd)
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(pl * ph));
                    const f32x2 w = pk2(r * ph, r * pl);
                    upk2(fma2(w, sqb, c_base), ta, tb);
                    upk2(fma2(w, sqa, c_base), tc, td);
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(ga) : "f"(ta));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gb) : "f"(tb));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gc) : "f"(tc));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gd) : "f"(td));
                    tail(p, ga, gb);
                    tail(p + 1, gc, gd);
                }
            } else {
#pragma unroll
            for (int p = 0; p < TP; ++p) {
                const f32x2 d0 = add2(hx0[p], nt0), d1 = add2(hx1[p], nt1);
                const f32x2 sq = fma2(d1, d1, fma2(d0, d0, c_max));
                float sl, sh, gl, gh;
                upk2(sq, sl, sh);
                if (p < NM) {
                    // reciprocal on the MUFU pipe (rcp.approx: <= 1 ulp), for NM of the TP pairs
                    float rl, rh, tl, th2;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(sl));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rh) : "f"(sh));
                    upk2(add2(pk2(rl, rh), c_base), tl, th2);
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gl) : "f"(tl));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gh) : "f"(th2));
                } else {
                    f32x2 nr = pk2(__int_as_float((int)(0xFEF311C7u - (unsigned)__float_as_int(sl))),
                                   __int_as_float((int)(0xFEF311C7u - (unsigned)__float_as_int(sh))));
                    nr = fma2(nr, fma2(sq, nr, c_one), nr);
                    nr = fma2(nr, fma2(sq, nr, c_one), nr);
                    nr = fma2(nr, fma2(sq, nr, c_one), nr);
                    float nl, nh;
                    upk2(add2(nr, c_nbase), nl, nh);               // -(base + 1/sq)
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gl) : "f"(-nl));
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gh) : "f"(-nh));
                }
                tail(p, gl, gh);
            }
            }
            if (write_seq) *pseq = S2;
            th = th_n; S2 = S2_n;
            th_n = th_nn; S2_n = S2_nn;
            pth += th_step; pseq += seq_step;
        }
    }
    // block-level sums over the RS row-threads of each column
    for (int t = 0; t < nT; ++t) {
        float mine = 0.f;
#pragma unroll
        for (int p = 0; p < TP; ++p) {
            float lo, hi;
            upk2(acc[p], lo, hi);
            if (t == 2 * p) mine = lo;
            if (t == 2 * p + 1) mine = hi;
        }
        __syncthreads();
        smem[tid] = mine;
        __syncthreads();
        if (r == 0 && b < B) {
            float a = mine;
            for (int rr = 1; rr < RS; ++rr) a += smem[rr * CB + c];
            part[((size_t)blockIdx.x * Ttot + t0 + t) * B + b] = a;
        }
    }
}

// ---- the whole history in ONE pass (location K = 1, D = 2, T <= 2 * kOnePairs) ----
// The multi-pass kernel above keeps the 12 history points of a pass in registers and therefore reads the thetas three
// times and carries the accumulated log-likelihood through a [rows, B] scratch array between the passes (cfg2: 7.4 GB
// of DRAM traffic for 1.6 GB of thetas), with one row in flight per thread.  Here a thread (fixed column b) holds R
// ROWS in registers -- (theta_0, theta_1) negated and duplicated for the packed operands, and the running S2 -- and
// walks all history points once: the records of a pair of points (t, t + 1) are four LDS.64 from shared memory
// ([pair][field][b] as float2), shared by the R rows; the per-point sums are 2 * kOnePairs register accumulators.
// One MUFU reciprocal serves two rows x two points (same identity as above, the partner is the next ROW instead of
// the next pair), or -- J8, max_signal >= 1e-4 -- all four rows x two points: 2.125 MUFU per evaluation.  R independent
// rows per thread supply the instruction-level parallelism that the 19 warps of the multi-pass kernel lacked; thetas
// are read once, nothing is written but the per-block sums.
// LAST (ALINE_SPCE_LAST_ONLY, the reference's stepwise = False): only the bound after the final history point is
// wanted -- the loop only accumulates log-likelihoods, one exponential per ROW at the end (1.125 MUFU per evaluation);
// the sums of the other history points are written as 1 (never read by the caller).
constexpr int kOnePairs = 18;

template <int R, bool LAST, int MAXT, bool J8, bool GEN = false>
__global__ void __launch_bounds__(MAXT, 1)
spce_fast_loc_onepass_kernel(const LocationLik<1, 2> lk, const float* __restrict__ HF, int T,
                             const float* __restrict__ thetas, long long row_begin, long long row_end, int B, int CB,
                             int RS, float* __restrict__ part, const GenArgs gen = GenArgs{}) {
    extern __shared__ __align__(16) float4 hsm[];           // [n_pairs][2][CB]: (y, x0) and (x1, c2) of points (2p, 2p + 1)
    static_assert(R % 2 == 0, "rows are evaluated two at a time (joint reciprocal)");
    const int tid = threadIdx.x;
    const int r = tid / CB, c = tid - r * CB;
    const int b = blockIdx.y * CB + c;
    const bool active = (r < RS) && (b < B);
    const int n_pairs = (T + 1) >> 1;
    for (int i = tid; i < n_pairs * 2 * CB; i += blockDim.x) {
        const int pg = i / CB, cc = i - pg * CB, p = pg >> 1, g = pg & 1, bb = blockIdx.y * CB + cc;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);        // (field 2g of t, of t+1, field 2g+1 of t, of t+1)
        if (bb < B) {
            v.x = __ldg(HF + ((size_t)(2 * p) * 4 + 2 * g) * B + bb);
            v.z = __ldg(HF + ((size_t)(2 * p) * 4 + 2 * g + 1) * B + bb);
            if (2 * p + 1 < T) {
                v.y = __ldg(HF + ((size_t)(2 * p + 1) * 4 + 2 * g) * B + bb);
                v.w = __ldg(HF + ((size_t)(2 * p + 1) * 4 + 2 * g + 1) * B + bb);
            }
        }
        hsm[i] = v;
    }
    __syncthreads();
    f32x2 acc[kOnePairs];
#pragma unroll
    for (int p = 0; p < kOnePairs; ++p) acc[p] = pk2(0.f, 0.f);
    if (active) {
        const f32x2 c_max = pk2(lk.max_signal, lk.max_signal), c_base = pk2(lk.base_signal, lk.base_signal);
        const f32x2 c_nln2 = pk2(-0.69314718055994530942f, -0.69314718055994530942f), c_k2 = pk2(lk.k2, lk.k2);
        const long long stride = (long long)gridDim.x * RS;
        const long long first = row_begin + (long long)blockIdx.x * RS + r;
        const int n_mine = first < row_end ? (int)((row_end - first + stride - 1) / stride) : 0;
        const float2* pth = reinterpret_cast<const float2*>(thetas + ((size_t)first * B + b) * 2);
        const size_t th_step = (size_t)stride * B;
        const bool odd_tail = (T & 1) != 0;
        float2 th[R], th_n[R];
        auto draw = [&](long long row) {                           // same function of (seed, row, b) as prior_box_kernel
            const unsigned long long g = (unsigned long long)(gen.row_offset + row);
            const Philox4 rr = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)b, 0u, gen.k0, gen.k1);
            return make_float2(fmaf(u01(rr.x[0]), gen.sc0, gen.lo0), fmaf(u01(rr.x[1]), gen.sc1, gen.lo1));
        };
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if constexpr (GEN) th[i] = draw(first + (long long)i * stride);
            else th[i] = i < n_mine ? __ldg(pth + i * th_step) : make_float2(0.f, 0.f);
        }
        for (int k = 0; k < n_mine; k += R) {
#pragma unroll
            for (int i = 0; i < R; ++i) {                              // next group of rows (loaded, or drawn: never in HBM)
                if constexpr (GEN) th_n[i] = draw(first + (long long)(k + R + i) * stride);
                else th_n[i] = k + R + i < n_mine ? __ldg(pth + (R + i) * th_step) : make_float2(0.f, 0.f);
            }
            f32x2 nt0[R], nt1[R];
            float S2[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                nt0[i] = pk2(-th[i].x, -th[i].x); nt1[i] = pk2(-th[i].y, -th[i].y);
                S2[i] = k + i < n_mine ? 0.f : -INFINITY;              // a row past the end contributes 2^-inf = 0
            }
            const float4* hp = hsm + c;
#pragma unroll
            for (int p = 0; p < kOnePairs; ++p) {
                if (p < n_pairs) {
                    const float4 h0 = hp[0], h1 = hp[CB];
                    hp += 2 * CB;
                    const f32x2 hy = pk2(h0.x, h0.y), hx0 = pk2(h0.z, h0.w), hx1 = pk2(h1.x, h1.y), hc = pk2(h1.z, h1.w);
                    auto tail = [&](const int row, const float gl, const float gh) {
                        const f32x2 d = fma2(c_nln2, pk2(gl, gh), hy);
                        float ll, lh;
                        upk2(fma2(mul2(d, d), c_k2, hc), ll, lh);
                        if constexpr (LAST) {
                            // the padded half of an odd history's last pair must not reach the final exponential
                            S2[row] += ll + ((odd_tail && p == n_pairs - 1) ? 0.f : lh);
                        } else {
                            float el, eh;
                            S2[row] += ll;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(el) : "f"(S2[row]));
                            S2[row] += lh;                             // (garbage after the last valid point: never read again)
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eh) : "f"(S2[row]));
                            acc[p] = add2(acc[p], pk2(el, eh));
                        }
                    };
                    auto lg2_tail = [&](const int row, const f32x2 t) {
                        float tl, th2, gl, gh;
                        upk2(t, tl, th2);
                        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gl) : "f"(tl));
                        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(gh) : "f"(th2));
                        tail(row, gl, gh);
                    };
                    if constexpr (J8) {
                        // ONE MUFU reciprocal for four rows x two points: with the packed products pA = sq0 sq1,
                        // pB = sq2 sq3, pAB = pA pB and r = 1 / (pAB.lo pAB.hi):  u = (r pAB.hi, r pAB.lo) = 1 / pAB,
                        // 1 / pA = u pB, 1 / pB = u pA, 1 / sq0 = sq1 / pA, ...  (eight factors >= max_signal >= 1e-4:
                        // the product stays normal; <= ~4 ulp).  2.125 MUFU per evaluation.
                        static_assert(!J8 || R == 4, "the eight-way reciprocal takes four rows");
                        f32x2 sq[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const f32x2 d0 = add2(hx0, nt0[i]), d1 = add2(hx1, nt1[i]);
                            sq[i] = fma2(d1, d1, fma2(d0, d0, c_max));
                        }
                        const f32x2 pA = mul2(sq[0], sq[1]), pB = mul2(sq[2], sq[3]);
                        float ql, qh, rr;
                        upk2(mul2(pA, pB), ql, qh);
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(ql * qh));
                        const f32x2 u = pk2(rr * qh, rr * ql);
                        const f32x2 iA = mul2(u, pB), iB = mul2(u, pA);
                        lg2_tail(0, fma2(iA, sq[1], c_base));
                        lg2_tail(1, fma2(iA, sq[0], c_base));
                        lg2_tail(2, fma2(iB, sq[3], c_base));
                        lg2_tail(3, fma2(iB, sq[2], c_base));
                    } else {
#pragma unroll
                    for (int i = 0; i < R; i += 2) {
                        const f32x2 d0a = add2(hx0, nt0[i]), d1a = add2(hx1, nt1[i]);
                        const f32x2 d0b = add2(hx0, nt0[i + 1]), d1b = add2(hx1, nt1[i + 1]);
                        const f32x2 sqa = fma2(d1a, d1a, fma2(d0a, d0a, c_max));
                        const f32x2 sqb = fma2(d1b, d1b, fma2(d0b, d0b, c_max));
                        float pl, ph, rr;
                        upk2(mul2(sqa, sqb), pl, ph);
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(pl * ph));
                        const f32x2 w = pk2(rr * ph, rr * pl);
                        lg2_tail(i, fma2(w, sqb, c_base));             // base + 1 / sqa
                        lg2_tail(i + 1, fma2(w, sqa, c_base));         // base + 1 / sqb
                    }
                    }
                }
            }
            if constexpr (LAST) {
                float e = 0.f;
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    float ei;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ei) : "f"(S2[i]));
                    e += ei;
                }
                acc[0] = add2(acc[0], pk2(e, 0.f));
            }
#pragma unroll
            for (int i = 0; i < R; ++i) th[i] = th_n[i];
            pth += R * th_step;
        }
    }
    // block-level sums over the RS row-threads of each column
    float* red = reinterpret_cast<float*>(hsm);
    for (int t = 0; t < T; ++t) {
        float mine = 0.f;
        if constexpr (LAST) {
            float lo, hi;
            upk2(acc[0], lo, hi);
            mine = t == T - 1 ? lo : 1.f;                              // other points: a finite, positive placeholder
        } else {
#pragma unroll
            for (int p = 0; p < kOnePairs; ++p) {
                float lo, hi;
                upk2(acc[p], lo, hi);
                if (t == 2 * p) mine = lo;
                if (t == 2 * p + 1) mine = hi;
            }
        }
        __syncthreads();
        red[tid] = mine;
        __syncthreads();
        if (r == 0 && b < B) {
            float a = mine;
            if (!(LAST && t != T - 1))
                for (int q = 1; q < RS; ++q) a += red[q * CB + c];
            part[((size_t)blockIdx.x * T + t) * B + b] = a;
        }
    }
}

// out_m = M_t (theta_0's accumulated log-likelihood), out_s = sum over blocks; flags invalid sums
// redo: the shifted sum a = sum_l 2^(S_l - M) (M = theta_0's accumulated log-likelihood) is only trusted while the terms
// the ex2.approx.ftz flushed to zero (each < 2^-126) cannot matter: n_rows * 2^-126 <= 1e-5 a.  Below that -- every
// contrastive draw ~60 nats less likely than theta_0 -- sPCE would still be fine (theta_0's own term dominates it), but
// sNMC excludes theta_0 and would lose up to log(L) nats: the robust online-(max, sum) kernels recompute the launch.
__global__ void spce_fast_finalize_kernel(const float* __restrict__ part, int G, int B, int T, const float* __restrict__ lp0,
                                          float* __restrict__ out_m, float* __restrict__ out_s, int* __restrict__ redo,
                                          float a_min) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // i = t*B + b
    if (i >= B * T) return;
    const int t = i / B, b = i % B;
    float a = 0.f;
    for (int g = lane; g < G; g += 32) a += part[((size_t)g * T + t) * B + b];
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
        out_m[(size_t)b * T + t] = lp0[(size_t)b * T + t];
        out_s[(size_t)b * T + t] = a;
        if (!(a > a_min) || !isfinite(a)) atomicOr(redo, 1);
    }
}

// Merge the per-block partials of history points [t0, t0+nT): out_m/out_s [B, T].  One warp per (t, b): lanes
// stride over the G blocks, then a shuffle tree merges the 32 (max, sum-exp) pairs.
__global__ void spce_finalize_kernel(const float2* __restrict__ part, int G, int B, int T, int t0, int nT,
                                     float* __restrict__ out_m, float* __restrict__ out_s,
                                     const int* __restrict__ run_flag) {
    if (run_flag && *run_flag == 0) return;
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // i = tt*B + b
    if (i >= B * nT) return;
    const int t = t0 + i / B, b = i % B;
    Lse a; a.init();
    for (int g = lane; g < G; g += 32) {
        float2 o = part[((size_t)g * T + t) * B + b];
        a.merge(o.x, o.y);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, a.m, o), s2 = __shfl_xor_sync(0xffffffffu, a.s, o);
        a.merge(m2, s2);
    }
    if (lane == 0) {
        out_m[(size_t)b * T + t] = a.m;
        out_s[(size_t)b * T + t] = a.s;
    }
}

// pce_loss = logsumexp_{l=0..L} seq - seq[0];  nmc_loss = logsumexp_{l=1..L} seq - seq[0]   (loss/eig.py:200-202)
__global__ void lse_combine_kernel(const float* __restrict__ m, const float* __restrict__ s,
                                   const float* __restrict__ lp0, int R, long long n,
                                   float* __restrict__ pce, float* __restrict__ nmc) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Lse a; a.init();
    for (int r = 0; r < R; ++r) a.merge(m[(size_t)r * n + i], s[(size_t)r * n + i]);
    float l0 = lp0[i];
    if (nmc) nmc[i] = (a.m + logf(a.s)) - l0;
    if (pce) {
        float M2 = fmaxf(a.m, l0);                     // theta_0 usually dominates: fold it in under a safe max
        pce[i] = (M2 + logf(a.s * expf(a.m - M2) + expf(l0 - M2))) - l0;
    }
}

template <class LK>
__global__ void loglik_kernel(const LK lk, const float* __restrict__ H, const float* __restrict__ thetas, int dth,
                              float* __restrict__ out, long long n_rows, int B, int* __restrict__ bad_flag) {
    long long total = n_rows * B;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int b = (int)(e % B);
        typename LK::Theta th;
        lk.load_theta(th, thetas + (size_t)e * dth);
        float hl[LK::NH];
#pragma unroll
        for (int f = 0; f < LK::NH; ++f) hl[f] = __ldg(H + (size_t)f * B + b);
        float v = lk.ll(th, hl);
        if constexpr (LK::CHECK_BAD) {
            if (!isfinite(v) && bad_flag) atomicOr(bad_flag, 1);
        }
        out[e] = v;
    }
}

// CensoredSigmoidNormal.log_prob, element-wise (distributions/censored_sigmoid_normal.py:47-86)
__global__ void csn_logprob_kernel(const float* __restrict__ loc, const float* __restrict__ scale,
                                   const float* __restrict__ value, float lo, float hi, long long n,
                                   float* __restrict__ out, int* __restrict__ bad_flag) {
    const float crit = 2.0f * kFltTiny;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float mu = loc[i], sigma = scale[i], y = value[i];
        float log_sigma = logf(sigma), two_var = 2.0f * (sigma * sigma);
        float t = sigmoid_inv(y);
        float lp = normal_logpdf(t, mu, two_var, log_sigma) - (-softplus_t(-t) - softplus_t(t));
        if (y == hi || y == lo) {
            float tl = sigmoid_inv(y == hi ? hi : lo);
            float cdf = 0.5f * (1.0f + erff((tl - mu) * (1.0f / sigma) / 1.41421356237309504880f));
            float c = (y == hi) ? 1.0f - cdf : cdf;
            lp = (c < crit) ? lp - logf(crit + fabsf((tl - mu) / sigma)) : logf(c);
        }
        if (y > hi || y < lo) lp = -INFINITY;
        if (!isfinite(lp) && bad_flag) atomicOr(bad_flag, 1);
        out[i] = lp;
    }
}

// ------------------------------------------------------------ host side ----
static int g_block_threads = 1024;   // measured on B200: one large block per SM beats several small ones here
static int g_fast_history = 1;       // shifted-accumulation fast path of aline_spce_history_ex
static int g_step_threads = 400;     // block size of the lean step kernel
static int g_fast_onepass = 1;       // whole history in one pass over the thetas (ALINE_SPCE_ONEPASS=0: the multi-pass kernel)
static int g_fast_packed = 1;        // packed fp32x2 fast history pass for location K=1, D=2 (ALINE_SPCE_PACKED=0: scalar)
static int g_fast_mufu_pairs = kJointRcp;  // kJointRcp (-4, default): one MUFU reciprocal per two pairs; 6: one per evaluation;
                                           // 0: reciprocal on the FMA pipe (ALINE_SPCE_MUFU_PAIRS=-4|6|0, development A/B switch)
static int g_step_tma = 1;           // TMA-staged single-launch step kernel (ALINE_SPCE_STEP_TMA=0: register-staged one)
static int g_step_rows = 0;          // rows per chunk (0 = auto: ~36 KB stages)
static int g_step_stages = 5;        // ring depth
static int g_pass_len_ces = 18;     // CES: see max_pass_len
static int g_pass_len = 9;          // default history points per pass (tuned on B200)

// tuning knobs (development only): ALINE_SPCE_PASS, ALINE_SPCE_THREADS, ALINE_SPCE_STEP_THREADS
static void read_env_once() {
    static bool done = false;
    if (done) return;
    done = true;
    if (const char* e = getenv("ALINE_SPCE_FAST")) g_fast_history = atoi(e) != 0;
    if (const char* e = getenv("ALINE_SPCE_PASS")) { int v = atoi(e); if (v >= 1 && v <= kMaxPass) g_pass_len = g_pass_len_ces = v; }
    if (const char* e = getenv("ALINE_SPCE_STEP_THREADS")) { int v = atoi(e); if (v >= 32 && v <= 512) g_step_threads = v; }
    if (const char* e = getenv("ALINE_SPCE_PACKED")) g_fast_packed = atoi(e) != 0;
    if (const char* e = getenv("ALINE_SPCE_ONEPASS")) g_fast_onepass = atoi(e) != 0;
    if (const char* e = getenv("ALINE_SPCE_MUFU_PAIRS")) { int v = atoi(e); if (v == 0 || v == 6 || v == kJointRcp) g_fast_mufu_pairs = v; }
    if (const char* e = getenv("ALINE_SPCE_STEP_TMA")) g_step_tma = atoi(e) != 0;
    if (const char* e = getenv("ALINE_SPCE_STEP_ROWS")) { int v = atoi(e); if (v >= 1 && v <= 4096) g_step_rows = v; }
    if (const char* e = getenv("ALINE_SPCE_STEP_STAGES")) { int v = atoi(e); if (v >= 2 && v <= kStepMaxStages) g_step_stages = v; }
    if (const char* e = getenv("ALINE_SPCE_THREADS")) { int v = atoi(e); if (v >= 32 && v <= 1024) g_block_threads = v; }
}

// does the fused location pass (K = 1, D = 2, seq as scratch) run as ONE pass over the contrastive rows?
static bool onepass_applies(float max_signal, int T) {
    read_env_once();
    return g_fast_history && g_fast_onepass && g_fast_packed && T > 1 && T <= 2 * kOnePairs && max_signal >= 1e-7f;
}

struct Plan {
    int CB, RS, threads, gx, gy;
    size_t smem;
};

static size_t pass_smem(int NH, int nT, int CB, int threads) {
    size_t a = (size_t)nT * (NH | 1) * CB * sizeof(float), b2 = (size_t)threads * sizeof(float2);
    return a > b2 ? a : b2;
}

static void plan_cols(int B, Plan& p, int max_threads = 512) {
    int cols = max_threads < kMaxColsPerBlock ? max_threads : kMaxColsPerBlock;
    int nc = ceil_div(B, cols);
    p.CB = ceil_div(B, nc);
    p.gy = nc;
    int tgt = max_threads < g_block_threads ? max_threads : g_block_threads;
    p.RS = tgt / p.CB;
    if (p.RS < 1) p.RS = 1;
    p.threads = p.CB * p.RS;
}

// history points one pass may cover: bounded by the compiled TC and by ~96 KB of shared memory for the records
// (CES: up to 18 points in one pass of 640 threads -- the per-draw work of load_theta (an exponential, two divisions)
// and the theta / seq traffic are paid once: cfg3, T = 15: 22.7 -> 21.9 ms against two passes of 8 + 7)
static int max_pass_len(int NH, int B, bool ces = false) {
    Plan p; plan_cols(B, p, max_threads_for(36));
    int by_smem = (int)((96 * 1024) / ((size_t)(NH | 1) * p.CB * sizeof(float)));
    if (by_smem < 1) by_smem = 1;
    read_env_once();
    const int cap = ces ? g_pass_len_ces : g_pass_len;
    return by_smem < cap ? by_smem : cap;
}

template <class K>
static int make_plan(K kernel, int TC, int NH, int nT, long long n_rows, int B, Plan& p) {
    plan_cols(B, p, max_threads_for(TC));
    p.smem = pass_smem(NH, nT, p.CB, p.threads);
    if (ensure_dyn_smem((const void*)kernel, p.smem)) return 1;
    int occ = 0;
    ALINE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, p.threads, p.smem));
    if (occ < 1) occ = 1;
    int sms = device_info().sm_count;
    long long want = ceil_div64(n_rows, p.RS);
    long long cap = (long long)sms * occ / p.gy;
    if (cap < 1) cap = 1;
    if (cap > kMaxGridX) cap = kMaxGridX;
    p.gx = (int)(want < cap ? want : cap);
    if (p.gx < 1) p.gx = 1;
    return 0;
}

template <class LK>
static int prep_hist(const LK&, const aline_lik* lik, const float* y, const float* xi, int B, int T, float* H,
                     cudaStream_t st) {
    int n = B * T, th = 128;
    if constexpr (is_ces<LK>::value) {
        prep_hist_ces<<<ceil_div(n, th), th, 0, st>>>(y, xi, B, T, lik->c1, H);
    } else {
        prep_hist_generic<LK><<<ceil_div(n, th), th, 0, st>>>(y, xi, B, T, lik->dim_x, H);
    }
    ALINE_LAUNCH_OK();
    return 0;
}

// One pass over history points [t0, t0+nT): the leading `skip_rows` rows (theta_0) by the cold variant,
// the contrastive rows by the hot one.
template <class LK, int TC, int U>
static int launch_pass(const LK& lk, const float* H, int t0, int nT, int T, const float* thetas, int dth, float* seq,
                       long long n_rows, int B, int skip_rows, int read_seq, int write_seq, float2* part,
                       float* out_lp0, int* bad_flag, int& G, cudaStream_t st, const int* run_flag = nullptr,
                       bool cold_only = false, bool hot_only = false) {
    Plan p;
    if (skip_rows > 0 && !hot_only) {
        if (make_plan(spce_stream_kernel<LK, TC, U, false>, TC, LK::NH, nT, skip_rows, B, p)) return 1;
        spce_stream_kernel<LK, TC, U, false><<<dim3(p.gx, p.gy), p.threads, p.smem, st>>>(
            lk, H, t0, nT, T, thetas, dth, seq, 0, skip_rows, B, p.CB, p.RS, read_seq, write_seq, part, out_lp0,
            bad_flag, run_flag);
        ALINE_LAUNCH_OK();
    }
    G = 0;
    if (n_rows > skip_rows && !cold_only) {
        if (make_plan(spce_stream_kernel<LK, TC, U, true>, TC, LK::NH, nT, n_rows - skip_rows, B, p)) return 1;
        spce_stream_kernel<LK, TC, U, true><<<dim3(p.gx, p.gy), p.threads, p.smem, st>>>(
            lk, H, t0, nT, T, thetas, dth, seq, skip_rows, n_rows, B, p.CB, p.RS, read_seq, write_seq, part, nullptr,
            bad_flag, run_flag);
        ALINE_LAUNCH_OK();
        G = p.gx;
    }
    return 0;
}

// CES power arithmetic (csrc/lik.cuh): 1 = exp2 / log2 form (default), 0 = eight powf.  ALINE_CES_POW=powf or
// aline_set_option("ces_fast_pow", 0) selects the latter (tests, A/B runs).
static std::atomic<int> g_ces_fast_pow{-1};
int ces_fast_pow_mode() {
    int m = g_ces_fast_pow.load(std::memory_order_relaxed);
    if (m < 0) {
        const char* e = getenv("ALINE_CES_POW");
        m = (e && e[0] == 'p') ? 0 : 1;
        g_ces_fast_pow.store(m, std::memory_order_relaxed);
    }
    return m;
}
void set_ces_fast_pow(int v) { g_ces_fast_pow.store(v ? 1 : 0, std::memory_order_relaxed); }

static size_t hist_bytes(int NH, int B, int T) { return ((size_t)T * NH * B * 4 + 255) / 256 * 256; }

static int finalize(const float2* part, int G, int B, int T, int t0, int nT, float* out_m, float* out_s,
                    cudaStream_t st, const int* run_flag = nullptr) {
    spce_finalize_kernel<<<ceil_div(B * nT * 32, 256), 256, 0, st>>>(part, G, B, T, t0, nT, out_m, out_s, run_flag);
    ALINE_LAUNCH_OK();
    return 0;
}

template <class LK>
static int run_history(const LK& lk, const aline_lik* lik, const float* y, const float* xi, const float* thetas,
                       float* seq, long long n_rows, int B, int T, int skip_rows, float* out_m, float* out_s,
                       float* out_lp0, int* bad_flag, void* scratch, size_t scratch_bytes, cudaStream_t st,
                       int flags = 0, const GenArgs* gen = nullptr, int* redo_out = nullptr) {
    const int dth = lik->dim_theta;
    size_t h_bytes = hist_bytes(kMaxNH + 1, B, T) + 256;      // robust records + fast records (one more field) + flag
    h_bytes += hist_bytes(kMaxNH + 1, B, T);
    size_t need = h_bytes + (size_t)kMaxGridX * T * B * sizeof(float2);
    ALINE_REQUIRE(scratch && scratch_bytes >= need, "aline_spce: scratch too small (%zu < %zu)", scratch_bytes, need);
    float* H = (float*)scratch;
    float2* part = (float2*)((char*)scratch + h_bytes);
    const int has_seq = seq != nullptr;
    int G = 0;
    read_env_once();
    if constexpr (!std::is_same<LK, LocationLik<1, 2>>::value) {
        if (prep_hist(lk, lik, y, xi, B, T, H, st)) return 1;
    }
    if constexpr (std::is_same<LK, LocationLik<1, 2>>::value) {
        if (T == 1 && has_seq && g_step_tma && B % 4 == 0 && B <= 480 && (((uintptr_t)thetas | (uintptr_t)seq) & 15) == 0) {
            // TMA-staged single-launch step (reads y / xi directly: no record prep, no cold pass, no finalize launch)
            int RS = 448 / B;
            if (RS < 1) RS = 1;
            if (RS > 8) RS = 8;
            int R = g_step_rows > 0 ? g_step_rows : (int)(36 * 1024 / ((size_t)B * 12));
            R = R / (2 * RS) * (2 * RS);
            if (R < 2 * RS) R = 2 * RS;
            int NS = g_step_stages;
            const size_t stage = (size_t)R * B * 12;
            while (NS > 2 && NS * stage > (size_t)device_info().max_smem_optin - 1024) --NS;
            size_t smem = NS * stage;
            const size_t merge_bytes = (size_t)RS * B * sizeof(float4);
            if (smem < merge_bytes) smem = merge_bytes;
            if (smem <= (size_t)device_info().max_smem_optin - 1024) {
                unsigned int* ticket = (unsigned int*)((char*)scratch + 2 * hist_bytes(kMaxNH + 1, B, T) + 64);
                ALINE_CHECK_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));
                if (ensure_dyn_smem((const void*)spce_step_tma_kernel, smem)) return 1;
                const int threads = 32 * ((RS * B + 31) / 32 + 1);
                long long want = ceil_div64(n_rows - skip_rows, R);
                int gx = device_info().sm_count;
                if (gx > kMaxGridX) gx = kMaxGridX;
                if (want < gx) gx = (int)(want < 1 ? 1 : want);
                spce_step_tma_kernel<<<gx, threads, smem, st>>>(lk, y, xi, thetas, seq, skip_rows, n_rows, B, R, RS, NS, part,
                                                                out_lp0, ticket, out_m, out_s);
                ALINE_LAUNCH_OK();
                return 0;
            }
        }
        if (prep_hist(lk, lik, y, xi, B, T, H, st)) return 1;
        if (T == 1 && has_seq && B % 2 == 0 && B / 2 <= 512) {
            // cold row(s) first (theta_0: out_lp0 + seq), then the lean HBM-bound kernel over the contrastive rows
            Plan p;
            if (skip_rows > 0) {
                if (make_plan(spce_stream_kernel<LK, 1, 4, false>, 1, LK::NH, 1, skip_rows, B, p)) return 1;
                spce_stream_kernel<LK, 1, 4, false><<<dim3(p.gx, p.gy), p.threads, p.smem, st>>>(
                    lk, H, 0, 1, T, thetas, dth, seq, 0, skip_rows, B, p.CB, p.RS, 1, 1, part, out_lp0, bad_flag, nullptr);
                ALINE_LAUNCH_OK();
            }
            const int CB2 = B / 2;
            int RS = g_step_threads / CB2;
            if (RS < 1) RS = 1;
            const int threads = CB2 * RS;
            const size_t smem = (size_t)threads * sizeof(float4);
            int occ = 1;
            ALINE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spce_step_loc12_kernel<4>, threads, smem));
            if (occ < 1) occ = 1;
            long long want = ceil_div64(n_rows - skip_rows, (long long)RS * 4);
            long long cap = (long long)device_info().sm_count * occ;
            if (cap > kMaxGridX) cap = kMaxGridX;
            int gx = (int)(want < cap ? want : cap);
            if (gx < 1) gx = 1;
            spce_step_loc12_kernel<4><<<gx, threads, smem, st>>>(lk, H, thetas, seq, skip_rows, n_rows, B, CB2, RS, part);
            ALINE_LAUNCH_OK();
            return finalize(part, gx, B, T, 0, 1, out_m, out_s, st);
        }
    }
    if (T == 1) {
        // EIGStepLoss.step: one history point, HBM-bound -> 4 rows in flight per thread
        if (launch_pass<LK, 1, 4>(lk, H, 0, 1, T, thetas, dth, seq, n_rows, B, skip_rows, has_seq, has_seq, part,
                                  out_lp0, bad_flag, G, st)) return 1;
        return finalize(part, G, B, T, 0, 1, out_m, out_s, st);
    }
    int cap = max_pass_len(LK::NH, B, is_ces<LK>::value);
    int npass = ceil_div(T, cap);
    ALINE_REQUIRE(npass == 1 || has_seq, "aline_spce_history: T=%d needs %d passes, seq must not be NULL", T, npass);
    int per = ceil_div(T, npass);
    // robust multi-pass evaluation (online (max, sum) pairs); seq_zero: treat seq as zeros on entry
    auto robust = [&](const int* run_flag, bool seq_zero, bool hot_only) -> int {
        for (int t0 = 0; t0 < T; t0 += per) {
            int nT = (T - t0 < per) ? T - t0 : per;
            int rd = has_seq && !(seq_zero && t0 == 0);
            int rc;
#define ALINE_PASS(TCV)                                                                                       \
            rc = launch_pass<LK, TCV, 1>(lk, H, t0, nT, T, thetas, dth, seq, n_rows, B, skip_rows, rd, has_seq, \
                                         part, out_lp0, bad_flag, G, st, run_flag, false, hot_only)
            if (nT <= 6) ALINE_PASS(6);
            else if (nT <= 9) ALINE_PASS(9);
            else if (nT <= 12) ALINE_PASS(12);
            else if (nT <= 18) ALINE_PASS(18);
            else ALINE_PASS(36);
#undef ALINE_PASS
            if (rc) return 1;
            if (finalize(part, G, B, T, t0, nT, out_m, out_s, st, run_flag)) return 1;
        }
        return 0;
    };

    if constexpr (LK::HAS_LL_LOG2) {
        // fast path: needs theta_0 as row 0 and a seq buffer that is pure scratch (zeros on entry, content not needed
        // afterwards); B small enough for the 1024-thread column mapping
        if ((flags & ALINE_SPCE_SEQ_SCRATCH) && skip_rows == 1 && has_seq && g_fast_history) {
            constexpr int NF = LK::NH + 1;
            float* HF = (float*)((char*)scratch + hist_bytes(kMaxNH + 1, B, T));
            int* redo = (int*)((char*)scratch + 2 * hist_bytes(kMaxNH + 1, B, T));
            float* partf = (float*)part;
            ALINE_CHECK_CUDA(cudaMemsetAsync(redo, 0, sizeof(int), st));
            // (1) theta_0's accumulated log-likelihood for every history point (cold rows only; writes seq row 0)
            for (int t0 = 0; t0 < T; t0 += per) {
                int nT = (T - t0 < per) ? T - t0 : per;
                int rc, rd = t0 > 0;
#define ALINE_COLD(TCV)                                                                                        \
                rc = launch_pass<LK, TCV, 1>(lk, H, t0, nT, T, thetas, dth, seq, n_rows, B, skip_rows, rd, 1, part, \
                                             out_lp0, bad_flag, G, st, nullptr, true, false)
                if (nT <= 6) ALINE_COLD(6);
                else if (nT <= 9) ALINE_COLD(9);
                else if (nT <= 12) ALINE_COLD(12);
                else if (nT <= 18) ALINE_COLD(18);
                else ALINE_COLD(36);
#undef ALINE_COLD
                if (rc) return 1;
            }
            prep_fast_kernel<LK><<<ceil_div(B * T, 128), 128, 0, st>>>(lk, H, out_lp0, B, T, HF);
            ALINE_LAUNCH_OK();
            // (2) contrastive rows, shifted accumulation
            if constexpr (std::is_same<LK, LocationLik<1, 2>>::value) {
                if (onepass_applies(lk.max_signal, T)) {
                    // the whole history in one pass over the thetas (R rows per thread, records from shared memory)
                    const int n_pairs = (T + 1) / 2;
                    int cols_cap = (int)((size_t)(device_info().max_smem_optin - 2048) / ((size_t)n_pairs * 4 * sizeof(float2)));
                    // measured at cfg2 (B = 200 -> 3 x 200 threads, ms): R = 2 / 4 / 6 rows per thread 4.33 / 4.14 / 4.27;
                    // 2 x 200 threads 4.89, 4 x 200 threads (72 registers, spills) 4.37; eight-way reciprocal 3.89
                    constexpr int R = 4, maxt = 608;
                    if (cols_cap > maxt) cols_cap = maxt;
                    Plan p;
                    p.gy = ceil_div(B, cols_cap);
                    p.CB = ceil_div(B, p.gy);
                    p.RS = maxt / p.CB;
                    p.threads = p.CB * p.RS;
                    size_t smem = (size_t)n_pairs * 2 * p.CB * sizeof(float4);
                    if (smem < (size_t)p.threads * sizeof(float)) smem = (size_t)p.threads * sizeof(float);
                    long long want = ceil_div64(n_rows - skip_rows, (long long)p.RS * R);
                    long long capg = (long long)device_info().sm_count / p.gy;
                    if (capg < 1) capg = 1;
                    if (capg > kMaxGridX) capg = kMaxGridX;
                    const int gx = (int)(want < capg ? want : capg);
#define ALINE_1P(LASTV, J8V, GENV)                                                                                     \
                    do {                                                                                               \
                        if (ensure_dyn_smem((const void*)spce_fast_loc_onepass_kernel<R, LASTV, maxt, J8V, GENV>, smem)) return 1; \
                        spce_fast_loc_onepass_kernel<R, LASTV, maxt, J8V, GENV><<<dim3(gx, p.gy), p.threads, smem, st>>>( \
                            lk, HF, T, thetas, skip_rows, n_rows, B, p.CB, p.RS, partf, gen ? *gen : GenArgs{});       \
                    } while (0)
                    const bool last = (flags & ALINE_SPCE_LAST_ONLY) != 0;
                    // the eight-way reciprocal multiplies eight (max_signal + distance^2) terms: keep the product normal
                    const bool j8 = g_fast_mufu_pairs == kJointRcp && lk.max_signal >= 1e-4f;
                    if (gen) { if (j8) ALINE_1P(false, true, true); else ALINE_1P(false, false, true); }
                    else if (j8) { if (last) ALINE_1P(true, true, false); else ALINE_1P(false, true, false); }
                    else { if (last) ALINE_1P(true, false, false); else ALINE_1P(false, false, false); }
#undef ALINE_1P
                    ALINE_LAUNCH_OK();
                    spce_fast_finalize_kernel<<<ceil_div(B * T * 32, 256), 256, 0, st>>>(partf, gx, B, T, out_lp0, out_m, out_s, redo,
                                                                                          (float)n_rows * 1.5e-33f);
                    ALINE_LAUNCH_OK();
                    if (gen) {
                        // contrastive rows drawn inside the pass (thetas / seq hold row 0 only): an invalid sum is reported
                        // to the caller instead of being recomputed here (the robust kernels read thetas)
                        if (redo_out) ALINE_CHECK_CUDA(cudaMemcpyAsync(redo_out, redo, sizeof(int), cudaMemcpyDeviceToDevice, st));
                        return 0;
                    }
                    return robust(redo, true, true);
                }
                if (g_fast_packed) {
                    constexpr int TP = 6, PTC = 2 * TP;               // 12 history points per pass, evaluated in pairs
                    Plan p;
                    plan_cols(B, p, 608);
                    const size_t smem = (size_t)p.threads * sizeof(float);
                    long long want = ceil_div64(n_rows - skip_rows, p.RS);
                    long long capg = (long long)device_info().sm_count / p.gy;
                    if (capg < 1) capg = 1;
                    if (capg > kMaxGridX) capg = kMaxGridX;
                    const int gx = (int)(want < capg ? want : capg);
                    // the joint reciprocal multiplies four (max_signal + distance^2) terms: keep the product normal
                    const bool joint = g_fast_mufu_pairs == kJointRcp && lk.max_signal >= 1e-7f;
                    if (gen) {
                        // contrastive rows drawn inside the pass (thetas holds row 0 only); an invalid sum is reported
                        // to the caller instead of being recomputed here (the robust kernels read thetas)
                        for (int t0 = 0; t0 < T; t0 += PTC) {
                            const int nT = (T - t0 < PTC) ? T - t0 : PTC;
#define ALINE_X2G(NMV, FULLV)                                                                                          \
                            spce_fast_loc12x2_kernel<TP, NMV, FULLV, true><<<dim3(gx, p.gy), p.threads, smem, st>>>(   \
                                lk, HF, t0, nT, T, thetas, seq, skip_rows, n_rows, B, p.CB, p.RS, t0 > 0, T > PTC, partf, *gen)
                            if (joint) { if (nT == PTC) ALINE_X2G(kJointRcp, true); else ALINE_X2G(kJointRcp, false); }
                            else       { if (nT == PTC) ALINE_X2G(6, true); else ALINE_X2G(6, false); }
#undef ALINE_X2G
                            ALINE_LAUNCH_OK();
                        }
                        spce_fast_finalize_kernel<<<ceil_div(B * T * 32, 256), 256, 0, st>>>(partf, gx, B, T, out_lp0, out_m, out_s, redo,
                                                                                          (float)n_rows * 1.5e-33f);
                        ALINE_LAUNCH_OK();
                        if (redo_out) ALINE_CHECK_CUDA(cudaMemcpyAsync(redo_out, redo, sizeof(int), cudaMemcpyDeviceToDevice, st));
                        return 0;
                    }
                    for (int t0 = 0; t0 < T; t0 += PTC) {
                        const int nT = (T - t0 < PTC) ? T - t0 : PTC;
#define ALINE_X2(NMV)                                                                                                  \
                        do {                                                                                           \
                            if (nT == PTC)                                                                             \
                                spce_fast_loc12x2_kernel<TP, NMV, true><<<dim3(gx, p.gy), p.threads, smem, st>>>(      \
                                    lk, HF, t0, nT, T, thetas, seq, skip_rows, n_rows, B, p.CB, p.RS, t0 > 0, T > PTC, partf); \
                            else                                                                                       \
                                spce_fast_loc12x2_kernel<TP, NMV, false><<<dim3(gx, p.gy), p.threads, smem, st>>>(     \
                                    lk, HF, t0, nT, T, thetas, seq, skip_rows, n_rows, B, p.CB, p.RS, t0 > 0, T > PTC, partf); \
                        } while (0)
                        if (joint) ALINE_X2(kJointRcp);                // one MUFU reciprocal per two pairs (default)
                        else if (g_fast_mufu_pairs == 0) ALINE_X2(0);  // reciprocal on the FMA pipe (development A/B switch)
                        else ALINE_X2(6);                              // one MUFU reciprocal per evaluation
#undef ALINE_X2
                        ALINE_LAUNCH_OK();
                    }
                    spce_fast_finalize_kernel<<<ceil_div(B * T * 32, 256), 256, 0, st>>>(partf, gx, B, T, out_lp0, out_m, out_s, redo,
                                                                                          (float)n_rows * 1.5e-33f);
                    ALINE_LAUNCH_OK();
                    return robust(redo, true, true);
                }
            }
            constexpr int FTC = 9;
            Plan p;
            plan_cols(B, p, 1024);
            size_t smem = (size_t)p.threads * sizeof(float);
            int occ = 1;
            ALINE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spce_fast_kernel<LK, FTC, false>, p.threads, smem));
            if (occ < 1) occ = 1;
            long long want = ceil_div64(n_rows - skip_rows, p.RS);
            long long capg = (long long)device_info().sm_count * occ / p.gy;
            if (capg < 1) capg = 1;
            if (capg > kMaxGridX) capg = kMaxGridX;
            const int gx = (int)(want < capg ? want : capg);
            const int fper = FTC;                                  // full passes of FTC points, one shorter tail pass
            for (int t0 = 0; t0 < T; t0 += fper) {
                int nT = (T - t0 < fper) ? T - t0 : fper;
                if (nT == FTC)
                    spce_fast_kernel<LK, FTC, true><<<dim3(gx, p.gy), p.threads, smem, st>>>(
                        lk, HF, t0, nT, T, thetas, dth, seq, skip_rows, n_rows, B, p.CB, p.RS, t0 > 0, T > fper, partf);
                else
                    spce_fast_kernel<LK, FTC, false><<<dim3(gx, p.gy), p.threads, smem, st>>>(
                        lk, HF, t0, nT, T, thetas, dth, seq, skip_rows, n_rows, B, p.CB, p.RS, t0 > 0, T > fper, partf);
                ALINE_LAUNCH_OK();
            }
            spce_fast_finalize_kernel<<<ceil_div(B * T * 32, 256), 256, 0, st>>>(partf, gx, B, T, out_lp0, out_m, out_s, redo,
                                                                                          (float)n_rows * 1.5e-33f);
            ALINE_LAUNCH_OK();
            // (3) conditional robust recomputation of the contrastive rows (early-exits unless a sum was invalid)
            return robust(redo, true, true);
        }
    }
    return robust(nullptr, (flags & ALINE_SPCE_SEQ_SCRATCH) != 0, false);
}

template <class LK>
static int run_loglik(const LK& lk, const aline_lik* lik, const float* y, const float* xi, const float* thetas,
                      float* out, long long n_rows, int B, int* bad_flag, void* scratch, size_t scratch_bytes,
                      cudaStream_t st) {
    size_t need = hist_bytes(LK::NH, B, 1);
    ALINE_REQUIRE(scratch && scratch_bytes >= need, "aline_log_likelihood: scratch too small (%zu < %zu)",
                  scratch_bytes, need);
    float* H = (float*)scratch;
    if (prep_hist(lk, lik, y, xi, B, 1, H, st)) return 1;
    long long total = n_rows * B;
    int th = 256;
    long long blocks = ceil_div64(total, th);
    long long cap = (long long)device_info().sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    loglik_kernel<LK><<<(int)blocks, th, 0, st>>>(lk, H, thetas, lik->dim_theta, out, n_rows, B, bad_flag);
    ALINE_LAUNCH_OK();
    return 0;
}

static int check_lik(const aline_lik* lik) {
    ALINE_REQUIRE(lik != nullptr, "aline_lik is NULL");
    switch (lik->task) {
        case ALINE_TASK_LOCATION:
            ALINE_REQUIRE(lik->K >= 1 && lik->dim_x >= 1 && lik->dim_theta == lik->K * lik->dim_x,
                          "location: dim_theta (%d) must equal K*dim_x (%d*%d)", lik->dim_theta, lik->K, lik->dim_x);
            ALINE_REQUIRE(lik->dim_theta <= 16 && lik->dim_x <= 7, "location: K*D <= 16 and D <= 7 supported");
            ALINE_REQUIRE(lik->c0 > 0.f, "location: noise_scale must be > 0");
            return 0;
        case ALINE_TASK_CES:
            ALINE_REQUIRE(lik->dim_x == 6 && lik->dim_theta == 5, "ces: dim_x must be 6 and dim_theta 5");
            return 0;
        case ALINE_TASK_PSYCHOMETRIC:
            ALINE_REQUIRE(lik->dim_x == 1 && lik->dim_theta == 4, "psychometric: dim_x must be 1 and dim_theta 4");
            return 0;
        default:
            return set_error("unknown task id %d (no CPU or generic fallback exists)", lik->task);
    }
}

// Dispatch on the task (and, for location, on the compiled (K, D) specialisations).
template <class F>
static int dispatch_lik(const aline_lik* lik, F&& f) {
    if (check_lik(lik)) return 1;
    if (lik->task == ALINE_TASK_CES) {
        if (ces_fast_pow_mode()) {
            CesLikT<true> lk; lk.noise_scale = lik->c0; lk.log_noise = logf(lik->c0);
            return f(lk);
        }
        CesLikT<false> lk; lk.noise_scale = lik->c0; lk.log_noise = logf(lik->c0);
        return f(lk);
    }
    if (lik->task == ALINE_TASK_PSYCHOMETRIC) {
        PsychometricLik lk;
        return f(lk);
    }
    float two_var = 2.0f * (lik->c0 * lik->c0), log_scale = logf(lik->c0);
#define ALINE_LOC_CASE(KK, DD)                                                          \
    if (lik->K == KK && lik->dim_x == DD) {                                              \
        LocationLik<KK, DD> lk;                                                          \
        lk.neg_inv_two_var = -1.0f / two_var; lk.lp_const = -log_scale - kLogSqrt2Pi;    \
        lk.k2 = lk.neg_inv_two_var * 1.44269504088896340736f;                            \
        lk.base_signal = lik->c1; lk.max_signal = lik->c2;                               \
        return f(lk);                                                                    \
    }
    ALINE_LOC_CASE(1, 2)
    ALINE_LOC_CASE(2, 2)
    ALINE_LOC_CASE(1, 1)
#undef ALINE_LOC_CASE
    LocationLikDyn lk;
    lk.two_var = two_var; lk.log_scale = log_scale; lk.base_signal = lik->c1; lk.max_signal = lik->c2;
    lk.K = lik->K; lk.D = lik->dim_x;
    return f(lk);
}

}  // namespace aline

using namespace aline;

extern "C" {

size_t aline_spce_scratch_bytes(int32_t B, int32_t T) {
    if (B < 1 || T < 1) return 0;
    return 2 * hist_bytes(kMaxNH + 1, B, T) + 256 + (size_t)kMaxGridX * T * B * sizeof(float2);
}

int32_t aline_spce_pass_len(const aline_lik* lik, int32_t B) {
    if (!lik || B < 1) return 0;
    int nh = lik->task == ALINE_TASK_CES ? CesLik::NH : (lik->task == ALINE_TASK_PSYCHOMETRIC ? 2 : 1 + lik->dim_x);
    if (lik->task == ALINE_TASK_LOCATION && !((lik->K == 1 && lik->dim_x == 2) || (lik->K == 2 && lik->dim_x == 2) ||
                                              (lik->K == 1 && lik->dim_x == 1)))
        nh = LocationLikDyn::NH;
    return max_pass_len(nh, B, lik->task == ALINE_TASK_CES);
}

int aline_spce_history_ex(const aline_lik* lik, const float* y, const float* xi, const float* thetas, float* seq,
                          int64_t n_rows, int32_t B, int32_t T, int32_t skip_rows, float* out_m, float* out_s,
                          float* out_lp0, int32_t* bad_flag, void* scratch, size_t scratch_bytes, int32_t flags,
                          void* stream);

int aline_spce_history(const aline_lik* lik, const float* y, const float* xi, const float* thetas, float* seq,
                       int64_t n_rows, int32_t B, int32_t T, int32_t skip_rows, float* out_m, float* out_s,
                       float* out_lp0, int32_t* bad_flag, void* scratch, size_t scratch_bytes, void* stream) {
    return aline_spce_history_ex(lik, y, xi, thetas, seq, n_rows, B, T, skip_rows, out_m, out_s, out_lp0, bad_flag,
                                 scratch, scratch_bytes, 0, stream);
}

int aline_spce_history_ex(const aline_lik* lik, const float* y, const float* xi, const float* thetas, float* seq,
                          int64_t n_rows, int32_t B, int32_t T, int32_t skip_rows, float* out_m, float* out_s,
                          float* out_lp0, int32_t* bad_flag, void* scratch, size_t scratch_bytes, int32_t flags,
                          void* stream) {
    ALINE_REQUIRE(y && xi && thetas && out_m && out_s, "aline_spce_history: NULL tensor");
    ALINE_REQUIRE(n_rows >= 1 && B >= 1 && T >= 1, "aline_spce_history: empty problem (n_rows=%lld B=%d T=%d)",
                  (long long)n_rows, B, T);
    ALINE_REQUIRE(skip_rows >= 0 && skip_rows <= 1, "aline_spce_history: skip_rows must be 0 or 1");
    ALINE_REQUIRE(skip_rows == 0 || out_lp0, "aline_spce_history: out_lp0 required when skip_rows = 1");
    return dispatch_lik(lik, [&](auto lk) {
        return run_history(lk, lik, y, xi, thetas, seq, n_rows, B, T, skip_rows, out_m, out_s,
                           skip_rows ? out_lp0 : nullptr, bad_flag, scratch, scratch_bytes, (cudaStream_t)stream, flags);
    });
}

int64_t aline_spce_device_prior_seq_rows(const aline_lik* lik, int64_t n_rows, int32_t T) {
    if (!lik || n_rows < 1) return n_rows;
    return onepass_applies(lik->c2, T) ? 1 : n_rows;
}

int aline_spce_history_device_prior(const aline_lik* lik, const aline_prior* prior, uint64_t seed, int64_t row_offset,
                                    const float* y, const float* xi, const float* theta0, float* seq, int64_t n_rows,
                                    int32_t B, int32_t T, float* out_m, float* out_s, float* out_lp0, int32_t* redo_flag,
                                    void* scratch, size_t scratch_bytes, void* stream) {
    ALINE_REQUIRE(lik && prior && y && xi && theta0 && seq && out_m && out_s && out_lp0 && redo_flag,
                  "aline_spce_history_device_prior: NULL argument");
    ALINE_REQUIRE(n_rows >= 2 && B >= 1 && T >= 1 && row_offset >= 0, "aline_spce_history_device_prior: empty problem");
    ALINE_REQUIRE(lik->task == ALINE_TASK_LOCATION && lik->K == 1 && lik->dim_x == 2 && prior->kind == ALINE_PRIOR_BOX &&
                  prior->dim_theta == 2, "aline_spce_history_device_prior: in-kernel draws exist for location K=1, D=2 with "
                  "a box prior; materialise other priors with aline_prior_sample and call aline_spce_history");
    read_env_once();
    ALINE_REQUIRE(g_fast_history && g_fast_packed, "aline_spce_history_device_prior: the packed fast pass is disabled");
    GenArgs g;
    g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    g.row_offset = row_offset;
    g.lo0 = prior->lo[0]; g.sc0 = prior->hi[0] - prior->lo[0]; g.lo1 = prior->lo[1]; g.sc1 = prior->hi[1] - prior->lo[1];
    return dispatch_lik(lik, [&](auto lk) {
        if constexpr (std::is_same<decltype(lk), LocationLik<1, 2>>::value) {
            return run_history(lk, lik, y, xi, theta0, seq, n_rows, B, T, 1, out_m, out_s, out_lp0, nullptr, scratch,
                               scratch_bytes, (cudaStream_t)stream, ALINE_SPCE_SEQ_SCRATCH, &g, redo_flag);
        } else {
            return set_error("aline_spce_history_device_prior: unsupported likelihood");
        }
    });
}

int aline_spce_step(const aline_lik* lik, const float* y, const float* xi, const float* thetas, float* seq,
                    int64_t n_rows, int32_t B, int32_t skip_rows, float* out_m, float* out_s, float* out_lp0,
                    int32_t* bad_flag, void* scratch, size_t scratch_bytes, void* stream) {
    ALINE_REQUIRE(seq != nullptr, "aline_spce_step: seq (EIGStepLoss.seq_logprobs) must not be NULL");
    return aline_spce_history(lik, y, xi, thetas, seq, n_rows, B, 1, skip_rows, out_m, out_s, out_lp0, bad_flag,
                              scratch, scratch_bytes, stream);
}

int aline_lse_combine(const float* m, const float* s, const float* lp0, int32_t R, int64_t n, float* pce_loss,
                      float* nmc_loss, void* stream) {
    ALINE_REQUIRE(m && s && lp0 && R >= 1 && n >= 1, "aline_lse_combine: bad arguments");
    int th = 128;
    lse_combine_kernel<<<(int)ceil_div64(n, th), th, 0, (cudaStream_t)stream>>>(m, s, lp0, R, n, pce_loss, nmc_loss);
    ALINE_LAUNCH_OK();
    return 0;
}

int aline_log_likelihood(const aline_lik* lik, const float* y, const float* xi, const float* thetas, float* out,
                         int64_t n_rows, int32_t B, int32_t* bad_flag, void* scratch, size_t scratch_bytes,
                         void* stream) {
    ALINE_REQUIRE(y && xi && thetas && out && n_rows >= 1 && B >= 1, "aline_log_likelihood: bad arguments");
    return dispatch_lik(lik, [&](auto lk) {
        return run_loglik(lk, lik, y, xi, thetas, out, n_rows, B, bad_flag, scratch, scratch_bytes,
                          (cudaStream_t)stream);
    });
}

int aline_censored_sigmoid_normal_log_prob(const float* loc, const float* scale, const float* value, float lower_lim,
                                           float upper_lim, int64_t n, float* out, int32_t* bad_flag, void* stream) {
    ALINE_REQUIRE(loc && scale && value && out && n >= 1, "aline_censored_sigmoid_normal_log_prob: bad arguments");
    int th = 256;
    long long blocks = ceil_div64(n, th);
    long long cap = (long long)device_info().sm_count * 8;
    if (blocks > cap) blocks = cap;
    csn_logprob_kernel<<<(int)blocks, th, 0, (cudaStream_t)stream>>>(loc, scale, value, lower_lim, upper_lim, n, out,
                                                                     bad_flag);
    ALINE_LAUNCH_OK();
    return 0;
}

}  // extern "C"
