// Softmax over the live candidates of one rollout, first-argmax, log-prob and the in-place Task.update_batch, as a
// block-level device function: the body of select_kernel (csrc/rollout.cu) and the prologue of the fused context kernel
// (csrc/ctx_warp.cu), which selects step t-1's design before it encodes the context of step t.
// reference: model/head.py:355-358 (eval-mode max), tasks/base_task.py:133-154 (append, without the compaction).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace aline {

struct ArgMax {
    float v; int i;
};
// larger value, then lower index (torch.max); a NaN counts as the maximum, like torch.max: one NaN / +inf logit makes
// every softmax probability NaN, and the reference then returns the FIRST live candidate with a NaN log-prob
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
    const bool an = a.v != a.v, bn = b.v != b.v;
    if (an || bn) return (bn && (!an || b.i < a.i)) ? b : a;
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

struct SelectArgs {
    const float* logits;            // [B, nq]
    unsigned char* alive;           // [B, nq] or NULL
    int nq;
    const float* qx; const float* qy;
    int dx, dy;
    float* cx; float* cy;           // append target (NULL: no append)
    int n_c, ctx_cap;               // the chosen point goes to context position n_c
    long long* idx_out; int idx_stride;
    float* logp_out; int logp_stride;
    long long* idx_orig_out;
    float* zt;
    // train-mode design choice (model/head.py:350-354): sample != 0 draws idx ~ Categorical(zt) by inverse CDF from one
    // Philox uniform keyed by (seed, rollout b, step) instead of taking the argmax
    int sample = 0;
    unsigned long long seed = 0;
    int step = 0;
};

// all threads of the block call this with the same arguments; b = rollout.  Ends with every result written by thread 0
// (callers that read them in the same kernel need a __syncthreads()).
__device__ __forceinline__ void select_block(const SelectArgs& a, int b) {
    const float* logits = a.logits;
    unsigned char* alive = a.alive;
    const int nq = a.nq, dx = a.dx, dy = a.dy, n_c = a.n_c, ctx_cap = a.ctx_cap, idx_stride = a.idx_stride,
              logp_stride = a.logp_stride;
    const float* qx = a.qx; const float* qy = a.qy;
    float* cx = a.cx; float* cy = a.cy;
    long long* idx_out = a.idx_out; float* logp_out = a.logp_out; long long* idx_orig_out = a.idx_orig_out;
    float* zt = a.zt;
    __shared__ float red_f[32];
    __shared__ ArgMax red_a[32];
    __shared__ int red_i[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const float* lg = logits + (size_t)b * nq;
    const unsigned char* al = alive ? alive + (size_t)b * nq : nullptr;

    float mx = -INFINITY;
    for (int j = tid; j < nq; j += blockDim.x)
        if (!al || al[j]) mx = fmaxf(mx, lg[j]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red_f[warp] = mx;
    __syncthreads();
    mx = red_f[0];
    for (int w = 1; w < nw; ++w) mx = fmaxf(mx, red_f[w]);
    __syncthreads();

    float sum = 0.f;
    for (int j = tid; j < nq; j += blockDim.x)
        if (!al || al[j]) sum += expf(lg[j] - mx);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red_f[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < nw; ++w) sum += red_f[w];

    ArgMax best{-1.f, 0x7fffffff};
    for (int j = tid; j < nq; j += blockDim.x) {
        if (!al || al[j]) {
            float p = expf(lg[j] - mx) / sum;
            if (zt) zt[(size_t)b * nq + j] = p;
            best = better(best, ArgMax{p, j});
        }
    }
    for (int o = 16; o; o >>= 1) {
        ArgMax other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
        best = better(best, other);
    }
    if (lane == 0) red_a[warp] = best;
    __syncthreads();
    best = red_a[0];
    for (int w = 1; w < nw; ++w) best = better(best, red_a[w]);
    if (a.sample && best.i >= 0 && best.i < nq && best.v == best.v) {
        // ---- Categorical(zt).sample(): inverse CDF over the live candidates in candidate order ----
        __shared__ float scan_w[32];
        __shared__ int win_tid, last_live_tid;
        const Philox4 rnd = philox4x32_10((uint32_t)b, (uint32_t)a.step, 0u, 0x5e1ec7u, (uint32_t)a.seed,
                                          (uint32_t)(a.seed >> 32));
        const float u = u01(rnd.x[0]);
        const int chunk = (nq + (int)blockDim.x - 1) / (int)blockDim.x;
        const int j0 = tid * chunk, j1 = min(nq, j0 + chunk);
        float part = 0.f;
        for (int j = j0; j < j1; ++j)
            if (!al || al[j]) part += expf(lg[j] - mx);
        float incl = part;                                        // inclusive scan over the threads' chunk sums
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (tid == 0) { win_tid = 0x7fffffff; last_live_tid = -1; }
        if (lane == 31) scan_w[warp] = incl;
        __syncthreads();
        float base = 0.f, total = 0.f;
        for (int w = 0; w < nw; ++w) { if (w < warp) base += scan_w[w]; total += scan_w[w]; }
        incl += base;
        const float target = u * total;
        if (part > 0.f) {
            atomicMax(&last_live_tid, tid);
            if (incl > target) atomicMin(&win_tid, tid);
        }
        __syncthreads();
        const int wt = win_tid != 0x7fffffff ? win_tid : last_live_tid;     // rounding: u * total == total
        if (tid == wt) {
            float cum = incl - part;
            int pick = -1;
            for (int j = j0; j < j1; ++j) {
                if (!al || al[j]) {
                    cum += expf(lg[j] - mx);
                    pick = j;                                     // last live candidate of the chunk if rounding runs out
                    if (cum > target) break;
                }
            }
            red_a[0] = ArgMax{expf(lg[pick] - mx) / sum, pick};
        }
        __syncthreads();
        best = red_a[0];
        __syncthreads();
    }
    if (best.i < 0 || best.i >= nq) {       // no live candidate left (uniform over the block): report it, touch nothing
        if (tid == 0) {
            idx_out[(size_t)b * idx_stride] = -1;
            logp_out[(size_t)b * logp_stride] = __int_as_float(0x7fc00000);
            if (idx_orig_out) idx_orig_out[b] = -1;
        }
        return;
    }

    // compacted index = number of live candidates before the winner
    int before = 0;
    for (int j = tid; j < best.i; j += blockDim.x)
        if (!al || al[j]) ++before;
    for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) red_i[warp] = before;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < nw; ++w) tot += red_i[w];
        idx_out[(size_t)b * idx_stride] = tot;
        // argmax: log of the fp32 probability (model/head.py:357); sample: Categorical.log_prob, whose logits are
        // log(clamp(probs, eps, 1 - eps)) (torch.distributions.utils.probs_to_logits)
        logp_out[(size_t)b * logp_stride] = a.sample ? logf(fminf(fmaxf(best.v, 1.1920929e-07f), 1.0f - 1.1920929e-07f))
                                                     : logf(best.v);
        if (idx_orig_out) idx_orig_out[b] = best.i;
        if (cx) {
            for (int k = 0; k < dx; ++k) cx[((size_t)b * ctx_cap + n_c) * dx + k] = qx[((size_t)b * nq + best.i) * dx + k];
            for (int k = 0; k < dy; ++k) cy[((size_t)b * ctx_cap + n_c) * dy + k] = qy[((size_t)b * nq + best.i) * dy + k];
        }
        if (alive) alive[(size_t)b * nq + best.i] = 0;
    }
}

}  // namespace aline
