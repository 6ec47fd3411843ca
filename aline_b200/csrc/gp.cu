// Gaussian-process prior draws: kernel matrix + jitter, Cholesky, f = L z, y = f + noise * eps.
//
// Replaces, in the reference, the per-batch-element Python loop of GPTask.generate_gp_data
// (tasks/gaussian_process.py:366-417: compute_kernel_matrix -> + jitter I -> torch.linalg.cholesky -> L @ randn
// -> + noise * randn) and the four kernel functions (194-317).  One thread block per matrix; the lower triangle
// lives packed in shared memory (N (N+1) / 2 floats: N = 301 -> 182 KB of the 227 KB), is factorised in place
// (right-looking, blocked by panels of 8 columns with a register-tiled rank-8 trailing update), then multiplied
// into the normal variates.  Matrices too large for shared memory use a caller-provided global scratch.
#include "common.cuh"

namespace aline {

__device__ __forceinline__ float gp_kernel_value(float sq, float scale, int type) {
    // sq = sum_d (x_d - x'_d)^2 / l_d^2
    if (type == 0) return scale * expf(-0.5f * sq);                        // rbf
    const float dist = sqrtf(sq);
    if (type == 1) return scale * expf(-dist);                              // matern 1/2
    if (type == 2) {                                                        // matern 3/2
        const float s3 = 1.7320508075688772f;
        return scale * (1.0f + s3 * dist) * expf(-s3 * dist);
    }
    const float s5 = 2.2360679774997898f;                                   // matern 5/2
    return scale * (1.0f + s5 * dist + (5.0f / 3.0f) * (dist * dist)) * expf(-s5 * dist);
}

__device__ __forceinline__ float gp_sqdist(const float* a, const float* b, const float* ls2, int dx) {
    float sq = 0.f;
    for (int d = 0; d < dx; ++d) {
        float df = a[d] - b[d];
        sq += (df * df) / ls2[d];
    }
    return sq;
}

// K [N, M] for one (x1, x2) pair   (GPTask.compute_kernel_matrix)
__global__ void gp_kernel_matrix_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int N, int M, int dx,
                                        const float* __restrict__ ls, const float* __restrict__ scale, int type,
                                        float* __restrict__ K) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * M) return;
    int i = e / M, k = e - i * M;
    float ls2[8];
    for (int d = 0; d < dx; ++d) ls2[d] = ls[d] * ls[d];
    K[e] = gp_kernel_value(gp_sqdist(x1 + (size_t)i * dx, x2 + (size_t)k * dx, ls2, dx), scale[0], type);
}

__device__ __forceinline__ size_t tri(int i, int k) { return (size_t)i * (i + 1) / 2 + k; }

__global__ void __launch_bounds__(512)
gp_sample_kernel(const float* __restrict__ x, int N, int dx, const float* __restrict__ ls, const float* __restrict__ scale,
                 const int* __restrict__ ktype, const float* __restrict__ z, const float* __restrict__ eps, float jitter,
                 float noise, float* __restrict__ y, float* __restrict__ L_out, float* __restrict__ K_out,
                 int* __restrict__ info, float* __restrict__ gscratch) {
    extern __shared__ __align__(16) float smem[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const size_t ntri = (size_t)N * (N + 1) / 2;
    float* zs = smem + N;                     // [N]
    float* A = gscratch ? gscratch + (size_t)b * ntri : smem + 2 * (size_t)N;
    const float* xb = x + (size_t)b * N * dx;
    float ls2[8];
    for (int d = 0; d < dx; ++d) { float l = ls[(size_t)b * dx + d]; ls2[d] = l * l; }
    const float sc = scale[b];
    const int type = ktype[b];

    for (int i = tid; i < N; i += blockDim.x) zs[i] = z[(size_t)b * N + i];
    // kernel matrix, lower triangle: warp per row, lanes over columns
    for (int i = warp; i < N; i += nw) {
        for (int k = lane; k <= i; k += 32) {
            float v = gp_kernel_value(gp_sqdist(xb + (size_t)i * dx, xb + (size_t)k * dx, ls2, dx), sc, type);
            if (k == i) v += jitter;
            A[tri(i, k)] = v;
            if (K_out) {
                K_out[((size_t)b * N + i) * N + k] = v;
                K_out[((size_t)b * N + k) * N + i] = v;
            }
        }
    }
    __syncthreads();

    // In-place Cholesky, right-looking, BLOCKED by panels of NB = 8 columns.  Per panel: (1) one thread factors the
    // 8 x 8 diagonal block (85 FMAs); (2) a thread per row solves its 8 panel entries against it (forward substitution,
    // the diagonal block broadcast from shared memory); (3) rank-8 update of the trailing triangle in 4 x 4 register
    // tiles: a thread loads the 8 panel values of its 4 rows and 4 columns once (64 reads) for 128 FMAs and 16
    // read-modify-writes -- 0.75 shared-memory accesses per FMA against 3 for the column-by-column rank-1 form, and 4
    // block barriers per 8 columns instead of 16.  (The unblocked form took 0.9 ms per 301 x 301 matrix = 1.80 ms for
    // 200 draws on 148 SMs; batched cuSOLVER potrf through torch.linalg.cholesky does the same draws in 1.36 ms.)
    constexpr int NB = 8;
    __shared__ float Ld[NB][NB + 1];          // the factored diagonal block (lower) ...
    __shared__ float inv_d[NB];               // ... and the reciprocals of its diagonal
    __shared__ int fail_s;
    if (tid == 0) fail_s = 0;
    __syncthreads();
    bool failed = false;
    for (int j0 = 0; j0 < N; j0 += NB) {
        const int nb = min(NB, N - j0), j1 = j0 + nb;
        // (1) diagonal block
        if (tid == 0) {
            for (int p = 0; p < nb; ++p) {
                for (int q = 0; q <= p; ++q) {
                    float v = A[tri(j0 + p, j0 + q)];
                    for (int t = 0; t < q; ++t) v = fmaf(-Ld[p][t], Ld[q][t], v);
                    if (q == p) {
                        if (!(v > 0.f)) { fail_s = 1; v = 1.f; }
                        const float d = sqrtf(v);
                        Ld[p][p] = d;
                        inv_d[p] = 1.0f / d;
                    } else {
                        Ld[p][q] = v * inv_d[q];
                    }
                }
            }
            for (int p = 0; p < nb; ++p)
                for (int q = 0; q <= p; ++q) A[tri(j0 + p, j0 + q)] = Ld[p][q];
        }
        __syncthreads();
        if (fail_s) { failed = true; break; }                  // uniform
        // (2) panel rows below the diagonal block: a[p] = (a[p] - sum_{q<p} a[q] Ld[p][q]) / Ld[p][p]
        for (int i = j1 + tid; i < N; i += blockDim.x) {
            float* row = A + tri(i, j0);
            float a[NB];
#pragma unroll
            for (int p = 0; p < NB; ++p) a[p] = p < nb ? row[p] : 0.f;
#pragma unroll
            for (int p = 0; p < NB; ++p) {
                if (p < nb) {
                    float v = a[p];
#pragma unroll
                    for (int q = 0; q < p; ++q) v = fmaf(-a[q], Ld[p][q], v);
                    a[p] = v * inv_d[p];
                }
            }
#pragma unroll
            for (int p = 0; p < NB; ++p)
                if (p < nb) row[p] = a[p];
        }
        __syncthreads();
        // (3) trailing update A[i][k] -= sum_p L[i][j0+p] L[k][j0+p], j1 <= k <= i < N, in 4 x 4 register tiles
        const int nt = (N - j1 + 3) >> 2;
        for (int ti = warp; ti < nt; ti += nw) {
            const int i0 = j1 + 4 * ti;
            float pi[4][NB];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int i = i0 + rr;
                const float* src = A + tri(min(i, N - 1), j0);
#pragma unroll
                for (int p = 0; p < NB; ++p) pi[rr][p] = (i < N && p < nb) ? src[p] : 0.f;
            }
            for (int tk = lane; tk <= ti; tk += 32) {
                const int k0 = j1 + 4 * tk;
                float pk[4][NB];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = k0 + c;
                    const float* src = A + tri(min(k, N - 1), j0);
#pragma unroll
                    for (int p = 0; p < NB; ++p) pk[c][p] = (k < N && p < nb) ? src[p] : 0.f;
                }
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int i = i0 + rr;
                    if (i >= N) continue;
                    float* row = A + tri(i, 0);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int k = k0 + c;
                        if (k <= i) {
                            float acc = 0.f;
#pragma unroll
                            for (int p = 0; p < NB; ++p) acc = fmaf(pi[rr][p], pk[c][p], acc);
                            row[k] -= acc;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (failed) {
        if (tid == 0 && info) info[b] = 1;
        for (int i = tid; i < N; i += blockDim.x) y[(size_t)b * N + i] = __int_as_float(0x7fc00000);
        return;
    }
    if (tid == 0 && info) info[b] = 0;

    // f = L z, y = f + noise * eps
    for (int i = warp; i < N; i += nw) {
        const float* row = A + tri(i, 0);
        float acc = 0.f;
        for (int k = lane; k <= i; k += 32) acc = fmaf(row[k], zs[k], acc);
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[(size_t)b * N + i] = acc + noise * eps[(size_t)b * N + i];
        if (L_out) {
            for (int k = lane; k < N; k += 32) L_out[((size_t)b * N + i) * N + k] = k <= i ? row[k] : 0.f;
        }
    }
}

}  // namespace aline

using namespace aline;

extern "C" {

size_t aline_gp_scratch_bytes(int32_t B, int32_t N) {
    // matrices whose packed triangle (+ 2 N floats) fits the 227 KB of shared memory need no global scratch
    size_t need = ((size_t)N * (N + 1) / 2 + 2 * (size_t)N) * sizeof(float);
    if (need <= 227 * 1024) return 0;
    return (size_t)B * ((size_t)N * (N + 1) / 2) * sizeof(float);
}

int aline_gp_sample(const float* x, int32_t B, int32_t N, int32_t dim_x, const float* lengthscales, const float* scale,
                    const int32_t* kernel_type, const float* z, const float* eps, float jitter, float noise_scale,
                    float* y, float* L_out, float* K_out, int32_t* info, void* scratch, size_t scratch_bytes,
                    void* stream) {
    ALINE_REQUIRE(x && lengthscales && scale && kernel_type && z && eps && y, "aline_gp_sample: NULL tensor");
    ALINE_REQUIRE(B >= 1 && N >= 1 && dim_x >= 1 && dim_x <= 8, "aline_gp_sample: bad sizes (B=%d N=%d dx=%d)", B, N, dim_x);
    size_t need = aline_gp_scratch_bytes(B, N);
    ALINE_REQUIRE(need == 0 || (scratch && scratch_bytes >= need), "aline_gp_sample: N=%d needs %zu bytes of scratch", N, need);
    size_t smem = need == 0 ? ((size_t)N * (N + 1) / 2 + 2 * (size_t)N) * sizeof(float) : 2 * (size_t)N * sizeof(float);
    if (smem > 48 * 1024)
        ALINE_CHECK_CUDA(cudaFuncSetAttribute(gp_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gp_sample_kernel<<<B, 512, smem, (cudaStream_t)stream>>>(x, N, dim_x, lengthscales, scale, kernel_type, z, eps,
                                                              jitter, noise_scale, y, L_out, K_out, info,
                                                              need ? (float*)scratch : nullptr);
    ALINE_LAUNCH_OK();
    return 0;
}

int aline_gp_kernel_matrix(const float* x1, const float* x2, int32_t N, int32_t M, int32_t dim_x,
                           const float* lengthscales, const float* scale, int32_t kernel_type, float* K, void* stream) {
    ALINE_REQUIRE(x1 && x2 && lengthscales && scale && K && N >= 1 && M >= 1 && dim_x >= 1 && dim_x <= 8,
                  "aline_gp_kernel_matrix: bad arguments");
    ALINE_REQUIRE(kernel_type >= 0 && kernel_type <= 3, "Unknown kernel type: %d", kernel_type);
    gp_kernel_matrix_kernel<<<ceil_div(N * M, 256), 256, 0, (cudaStream_t)stream>>>(x1, x2, N, M, dim_x, lengthscales,
                                                                                     scale, kernel_type, K);
    ALINE_LAUNCH_OK();
    return 0;
}

}  // extern "C"
