// Simulator log-likelihoods of the three BED tasks, as device functors.
//
// Each functor describes, for one history point (b, t), a small record H of
// theta-independent values ("history record", NH floats, laid out [t][f][b] so
// that consecutive threads = consecutive b read consecutive addresses), and a
// per-theta evaluation ll(theta, H).  fp32 throughout, precise libm (no fast
// math): the arithmetic follows the reference op for op so the sPCE bound
// agrees to ~1e-6 (tolerance 1e-4).
//
//   location     tasks/location_finding.py:110-130 (total_density), 149-164 (log_likelihood)
//   ces          tasks/ces.py:96-115,169-210 + distributions/censored_sigmoid_normal.py:47-86
//   psychometric tasks/psychometric.py:107-134,178-195
#pragma once
#include "common.cuh"

namespace aline {

constexpr float kLogSqrt2Pi = 0.91893853320467274178f;   // log(sqrt(2*pi))
constexpr float kFltTiny = 1.17549435e-38f;               // torch.finfo(float32).tiny
constexpr float kFltEps = 1.1920928955078125e-07f;        // torch.finfo(float32).eps

// torch.distributions.Normal.log_prob with precomputed 2*var and log(scale)
__device__ __forceinline__ float normal_logpdf(float v, float loc, float two_var, float log_scale) {
    float d = v - loc;
    return -(d * d) / two_var - log_scale - kLogSqrt2Pi;
}

// 1/x for normal positive x: MUFU.RCP + one Newton step (<= 1 ulp; the reference's pow(-1) is IEEE division)
__device__ __forceinline__ float rcp_nr(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

// 1/x on the FMA pipe only (no MUFU): integer-trick seed (5 % error) + 3 Newton steps -> ~1 ulp.  Used where the
// MUFU pipe is the bottleneck (fused sPCE history pass: rcp + lg2 + ex2 per evaluation).
__device__ __forceinline__ float rcp_fma(float x) {
    float r = __int_as_float(0x7EF311C7 - __float_as_int(x));
    r = fmaf(r, fmaf(-x, r, 1.0f), r);
    r = fmaf(r, fmaf(-x, r, 1.0f), r);
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

#ifndef ALINE_RCP_FMA
#define ALINE_RCP_FMA 1
#endif

// log(x) for normal positive x.  ALINE_FAST_LOG: MUFU.LG2 * ln2 (abs err 2^-21.4 on [0.5,2], 3 ulp elsewhere),
// measured effect on the location sPCE bound < 1e-5 relative (tolerance 1e-4); otherwise libm logf (1 ulp).
#ifndef ALINE_FAST_LOG
#define ALINE_FAST_LOG 1
#endif
__device__ __forceinline__ float log_pos(float x) {
#if ALINE_FAST_LOG
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.69314718055994530942f;
#else
    return logf(x);
#endif
}

// torch.nn.functional.softplus (beta=1, threshold=20)
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// torch SigmoidTransform._inverse
__device__ __forceinline__ float sigmoid_inv(float y) {
    y = fminf(fmaxf(y, kFltTiny), 1.0f - kFltEps);
    return logf(y) - log1pf(-y);
}

// ------------------------------------------------------------- location ----
// H = [y, xi_0 .. xi_{D-1}];  theta = [K][D]
template <int K_, int D_>
struct LocationLik {
    static constexpr int NH = 1 + D_;
    static constexpr int DTH = K_ * D_;
        static constexpr bool CHECK_BAD = false;
    float neg_inv_two_var, lp_const, base_signal, max_signal;     // lp_const = -log(scale) - log(sqrt(2 pi))
    float k2;                                                      // neg_inv_two_var * log2(e)
    struct Theta { float v[DTH]; };

    __device__ __forceinline__ void load_theta(Theta& th, const float* __restrict__ p) const {
        if constexpr (DTH % 4 == 0) {
#pragma unroll
            for (int i = 0; i < DTH / 4; ++i) {
                float4 q = __ldg(reinterpret_cast<const float4*>(p) + i);
                th.v[4 * i] = q.x; th.v[4 * i + 1] = q.y; th.v[4 * i + 2] = q.z; th.v[4 * i + 3] = q.w;
            }
        } else if constexpr (DTH % 2 == 0) {
#pragma unroll
            for (int i = 0; i < DTH / 2; ++i) {
                float2 q = __ldg(reinterpret_cast<const float2*>(p) + i);
                th.v[2 * i] = q.x; th.v[2 * i + 1] = q.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < DTH; ++i) th.v[i] = __ldg(p + i);
        }
    }
    // h[0] = y, h[1+d] = xi_d
    __device__ __forceinline__ float ll(const Theta& th, const float* h) const {
        float tot = base_signal;
#pragma unroll
        for (int k = 0; k < K_; ++k) {
            float sq = max_signal;
#pragma unroll
            for (int d = 0; d < D_; ++d) {
                float df = h[1 + d] - th.v[k * D_ + d];
                sq = fmaf(df, df, sq);
            }
            tot += rcp_nr(sq);                         // pow(-1)
        }
        float d = h[0] - log_pos(tot);
        return fmaf(d * d, neg_inv_two_var, lp_const);  // Normal(signal, scale).log_prob(y)
    }
    // log2-domain variant for the shifted accumulation of the fast history pass:
    // returns ll * log2(e) + c2 with the constant term of the Normal log-density already folded into c2
    static constexpr bool HAS_LL_LOG2 = true;
    __device__ __forceinline__ float ll_log2(const Theta& th, const float* h, float c2) const {
        float tot = base_signal;
#pragma unroll
        for (int k = 0; k < K_; ++k) {
            float sq = max_signal;
#pragma unroll
            for (int d = 0; d < D_; ++d) {
                float df = h[1 + d] - th.v[k * D_ + d];
                sq = fmaf(df, df, sq);
            }
#if ALINE_RCP_FMA
            tot += rcp_fma(sq);
#else
            tot += rcp_nr(sq);
#endif
        }
        float lg;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(tot));
        float d = fmaf(-0.69314718055994530942f, lg, h[0]);
        return fmaf(d * d, k2, c2);
    }
    __device__ __forceinline__ float const_log2() const { return lp_const * 1.44269504088896340736f; }
};

// Run-time (K, D) fallback for shapes without a compiled specialisation: K*D <= 16, D <= 7.
struct LocationLikDyn {
    static constexpr bool HAS_LL_LOG2 = false;
    static constexpr int NH = 8;
    static constexpr int DTH = 16;
        static constexpr bool CHECK_BAD = false;
    float two_var, log_scale, base_signal, max_signal;
    int K, D;
    struct Theta { float v[DTH]; };
    __device__ __forceinline__ void load_theta(Theta& th, const float* __restrict__ p) const {
#pragma unroll
        for (int i = 0; i < DTH; ++i) th.v[i] = (i < K * D) ? __ldg(p + i) : 0.f;
    }
    __device__ __forceinline__ float ll(const Theta& th, const float* h) const {
        float tot = 0.f;
        for (int k = 0; k < K; ++k) {
            float sq = 0.f;
            for (int d = 0; d < D; ++d) {
                float df = h[1 + d] - th.v[k * D + d];
                sq += df * df;
            }
            tot += 1.0f / (max_signal + sq);
        }
        float signal = logf(base_signal + tot);
        return normal_logpdf(h[0], signal, two_var, log_scale);
    }
};

// ------------------------------------------------------------------ CES ----
// H = [x0..x5 (clamped), 1+||b1-b2||, t=g(y), J(t), case, log2(x0..x5) hi, log2(x0..x5) lo];  theta = [rho, a1, a2, a3, log u]
// case: 0 interior, 1 y==hi (upper censor), 2 y==lo (lower censor), 3 outside -> -inf
//
// Powers.  The reference evaluates U(x) = (sum_i a_i x_i^rho)^(1/rho) with eight fp32 pow() per (draw, history point);
// with powf that is ~85 % of this kernel's instructions (cfg3: 86 ms for L = 1e7).  fast_pow = 1 (default):
//   x_i^rho = 2^(rho log2 x_i): log2 x_i depends on the HISTORY only, so it is computed once per history point in double
//     and stored as an fp32 hi + lo pair; the product rho * log2 x_i is carried as hi + lo (one FMA recovers the rounding
//     error of the product) and 2^hi (ex2.approx, <= 2^-22 relative) is corrected by (1 + ln2 * lo);
//   g^(1/rho) = 2^(log2(g) / rho): log2(g) from an exponent split and the atanh series of the mantissa in fp32 FMAs
//     (~1e-7 relative -- lg2.approx's 2^-22 ABSOLUTE error would be amplified by 1 / rho <= 100 where g ~ 1), the quotient
//     again carried as hi + lo.
// Neither form reproduces the reference bit for bit; the reference's own round-off (pow to 1 ulp, amplified by 1 / rho)
// is of the same size, and the gate is the bound itself: tests/golden/spce_ces_large.npz (L = 1e5, 1e-4 relative).
// FAST: the exp2 / log2 power arithmetic below (default); false: eight powf and the reference's op order
// (ALINE_CES_POW=powf) -- a compile-time choice so that the hot kernel carries only one of the two
template <bool FAST>
struct CesLikT {
    static constexpr bool HAS_LL_LOG2 = false;
    static constexpr int NH = 24;
    static constexpr bool fast_pow = FAST;
    static constexpr int DTH = 5;
        static constexpr bool CHECK_BAD = true;
    float noise_scale, log_noise = 0.f;         // log_noise = log(noise_scale), set by the host
    struct Theta { float rho, inv_rho, a1, a2, a3, u, inv_nu, log_nu; };

    __device__ __forceinline__ void load_theta(Theta& th, const float* __restrict__ p) const {
        th.rho = __ldg(p);
        th.a1 = __ldg(p + 1); th.a2 = __ldg(p + 2); th.a3 = __ldg(p + 3);
        const float lu = __ldg(p + 4);
        th.u = expf(lu);
        th.inv_rho = 1.0f / th.rho;
        th.inv_nu = 1.0f / (noise_scale * th.u);        // once per draw: sigma = h6 noise u never needs a division per point
        th.log_nu = log_noise + lu;
    }
    // 2^(a * b) with the product carried as hi + lo
    static __device__ __forceinline__ float exp2_prod(float a, float b_hi, float b_lo) {
        const float p = a * b_hi;
        const float e = fmaf(a, b_hi, -p) + a * b_lo;
        float v;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(p));
        return fmaf(v, e * 0.69314718055994530942f, v);
    }
    // log2(g), g > 0 normal: g = m 2^k with m in [sqrt(1/2), sqrt(2)), log2 m = (2 / ln 2) atanh((m - 1) / (m + 1)).
    // The quotient: MUFU reciprocal of m + 1 in [1.7, 2.5) + one Newton step + one residual correction of the quotient
    // (<= 1 ulp, branch-free; an IEEE division here is ~14 instructions and a slow-path branch)
    static __device__ __forceinline__ float log2_acc(float g) {
        const int ib = __float_as_int(g);
        const int k = (ib - 0x3f3504f3) >> 23;
        const float m = __int_as_float(ib - (k << 23));
        const float n = m - 1.0f, d = m + 1.0f;
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
        r = fmaf(r, fmaf(-d, r, 1.0f), r);
        float s = n * r;
        s = fmaf(r, fmaf(-s, d, n), s);
        const float s2 = s * s;
        float q = fmaf(s2, 1.0f / 11, 1.0f / 9);
        q = fmaf(q, s2, 1.0f / 7);
        q = fmaf(q, s2, 1.0f / 5);
        q = fmaf(q, s2, 1.0f / 3);
        const float lm = 2.8853900817779268f * fmaf(s * s2, q, s);           // 2 / ln 2
        return (float)k + lm;
    }
    __device__ __forceinline__ float ll(const Theta& th, const float* h) const {
        float mu, z, ez, base_lp;                     // z = (t - mu) / sigma, ez = z / sqrt(2) (the erf argument)
        const float t = h[7];
        if constexpr (FAST) {
            const float g1 = th.a1 * exp2_prod(th.rho, h[10], h[16]) + th.a2 * exp2_prod(th.rho, h[11], h[17]) +
                             th.a3 * exp2_prod(th.rho, h[12], h[18]);
            const float g2 = th.a1 * exp2_prod(th.rho, h[13], h[19]) + th.a2 * exp2_prod(th.rho, h[14], h[20]) +
                             th.a3 * exp2_prod(th.rho, h[15], h[21]);
            const float U1 = exp2_prod(th.inv_rho, log2_acc(g1), 0.f);
            const float U2 = exp2_prod(th.inv_rho, log2_acc(g2), 0.f);
            mu = (U1 - U2) * th.u;
            // sigma = h6 noise u:  1 / sigma = (1 / h6) (1 / (noise u)),  log sigma = log h6 + log noise + log u
            z = (t - mu) * (h[22] * th.inv_nu);
            base_lp = fmaf(-0.5f * z, z, -(h[23] + th.log_nu)) - kLogSqrt2Pi - h[8];
            ez = z * 0.70710678118654752440f;
        } else {
            const float g1 = th.a1 * powf(h[0], th.rho) + th.a2 * powf(h[1], th.rho) + th.a3 * powf(h[2], th.rho);
            const float g2 = th.a1 * powf(h[3], th.rho) + th.a2 * powf(h[4], th.rho) + th.a3 * powf(h[5], th.rho);
            const float U1 = powf(g1, th.inv_rho);
            const float U2 = powf(g2, th.inv_rho);
            mu = (U1 - U2) * th.u;
            const float sigma = h[6] * noise_scale * th.u;
            base_lp = normal_logpdf(t, mu, 2.0f * (sigma * sigma), logf(sigma)) - h[8];
            z = (t - mu) / sigma;
            ez = (t - mu) * (1.0f / sigma) / 1.41421356237309504880f;
        }
        int cs = __float_as_int(h[9]);
        if (cs == 0) return base_lp;
        if (cs == 3) return -INFINITY;
        const float crit = 2.0f * kFltTiny;
        float cdf = 0.5f * (1.0f + erff(ez));
        float c = (cs == 1) ? 1.0f - cdf : cdf;
        if constexpr (FAST) return (c < crit) ? base_lp - log_pos(crit + fabsf(z)) : log_pos(c);
        if (c < crit) return base_lp - logf(crit + fabsf(z));
        return logf(c);
    }
};

using CesLik = CesLikT<true>;

// H record builder for CES (theta independent part of the likelihood).
__device__ __forceinline__ void ces_prepare(const float* xi6, float y, float epsilon, float* out /*NH*/) {
    float x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { x[i] = fminf(fmaxf(xi6[i], 0.01f), 100.0f); out[i] = x[i]; }
    float d0 = x[0] - x[3], d1 = x[1] - x[4], d2 = x[2] - x[5];
    out[6] = 1.0f + sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    float lo = epsilon, hi = 1.0f - epsilon;
    float t = sigmoid_inv(y);
    out[7] = t;
    out[8] = -softplus_t(-t) - softplus_t(t);
    int cs = 0;
    if (y == hi) cs = 1;
    if (y == lo) cs = 2;
    if (y > hi || y < lo) cs = 3;
    out[9] = __int_as_float(cs);
#pragma unroll
    for (int i = 0; i < 6; ++i) {                      // log2 of the clamped design as fp32 hi + lo (once per history point)
        const double l = log2((double)x[i]);
        const float hi_ = (float)l;
        out[10 + i] = hi_;
        out[16 + i] = (float)(l - (double)hi_);
    }
    out[22] = 1.0f / out[6];                           // sigma = out[6] * noise * u (tasks/ces.py:190-197)
    out[23] = logf(out[6]);
}

// --------------------------------------------------------- psychometric ----
// H = [x, y];  theta = [alpha, beta, gamma, lambda]
struct PsychometricLik {
    static constexpr bool HAS_LL_LOG2 = false;
    static constexpr int NH = 2;
    static constexpr int DTH = 4;
        static constexpr bool CHECK_BAD = false;
    struct Theta { float a, b, g, lam; };
    __device__ __forceinline__ void load_theta(Theta& th, const float* __restrict__ p) const {
        float4 q = __ldg(reinterpret_cast<const float4*>(p));
        th.a = q.x; th.b = q.y; th.g = q.z; th.lam = q.w;
    }
    __device__ __forceinline__ float ll(const Theta& th, const float* h) const {
        float z = (h[0] - th.a) / th.b;
        float F = 1.0f - expf(-powf(10.0f, z));
        float p = th.lam * th.g + (1.0f - th.lam) * F;
        float y = h[1];
        return y * logf(p + 1e-10f) + (1.0f - y) * logf(1.0f - p + 1e-10f);
    }
};

}  // namespace aline
