// kernels_project.cuh — fp64 projection / evaluation kernels behind predict(), nnls() and evaluate()
// (SURVEY.md §8f-2, §8f-3). The reference runs these in double on the CPU and has no GPU entry for them:
//   Rcpp_predict  src/RcppFunctions_utils.cpp:23-53   (Gram + tiny twice + L2, RHS, cold CD with L1 inside)
//   c_nnls        src/RcppFunctions_utils.cpp:314-366 (same; warm start: B -= G·h, CD without tolerance)
//   compute_mse / compute_loss_general  src/RcppFunctions_utils.cpp:60-149 (dense reconstruction in the
//   reference — O(m·n) memory; here per non-zero dot products and the Gram trick)
// Same fused gather + coordinate-descent structure as the fp32 ALS kernel, one warp per column, a lane owns
// coordinates lane, lane+32, …; all arithmetic in IEEE double with separately rounded mul/add.
#pragma once

#include "common.cuh"

namespace b200 {

// G[j*KP+i] = Σ_c F[c][i]·F[c][j] over this CTA's rows, fp64. partials[cta][KP*KP].
template <int KP>
static __global__ void __launch_bounds__(256) gram_f64_kernel(const double* __restrict__ F, long long ncols,
                                                              double* __restrict__ partials) {
    constexpr int PER = (KP * KP + 255) / 256;
    __shared__ double srow[8][KP];
    double acc[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) acc[u] = 0.0;
    for (long long c0 = static_cast<long long>(blockIdx.x) * 8; c0 < ncols; c0 += static_cast<long long>(gridDim.x) * 8) {
        for (int t = threadIdx.x; t < 8 * KP; t += 256) {
            const long long c = c0 + t / KP;
            srow[t / KP][t % KP] = (c < ncols) ? F[c * KP + (t % KP)] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int e = threadIdx.x + u * 256;
            if (e < KP * KP) {
                const int i = e % KP, j = e / KP;
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[u] = fma(srow[r][i], srow[r][j], acc[u]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int e = threadIdx.x + u * 256;
        if (e < KP * KP) partials[static_cast<size_t>(blockIdx.x) * KP * KP + e] = acc[u];
    }
}

// G = Σ partials (fixed order) + diag_add on the first k diagonal entries; padded entries 0.
static __global__ void gram_f64_finalize_kernel(const double* __restrict__ partials, int nparts, int KP, int k,
                                                double diag_add1, double diag_add2, double diag_add3,
                                                double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= KP * KP) return;
    const int i = e % KP, j = e / KP;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += partials[static_cast<size_t>(c) * KP * KP + e];
    if (i == j) { s += diag_add1; s += diag_add2; s += diag_add3; }    // tiny (gram.hpp:51), tiny again, L2
    if (i >= k || j >= k) s = 0.0;
    G[e] = s;
}

struct ProjectParams {
    const int* __restrict__ colptr;
    const int* __restrict__ rowidx;
    const double* __restrict__ vals;
    const double* __restrict__ F;      // [rows][KP] fixed factor (w)
    double* __restrict__ X;            // [ncols][KP] solution (in: warm start)
    const double* __restrict__ G;      // KP×KP col-major
    int ncols, k;
    double L1, ub, cd_tol;
    int cd_maxit, nonneg, warm;
    int* work_counter;
};

template <int KP>
__global__ void __launch_bounds__(256) project_f64_kernel(const ProjectParams p) {
    constexpr int NC = (KP + 31) / 32;
    extern __shared__ __align__(16) double sGd[];          // KP*KP
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) sGd[e] = p.G[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int k = p.k;
    const bool nonneg = p.nonneg != 0, has_ub = p.ub > 0.0;
    // c_nnls' warm-start branch calls CD without a tolerance (src/RcppFunctions_utils.cpp:352-355)
    const bool check = (p.cd_tol > 0.0) && !p.warm;
    const double inv_k = 1.0 / static_cast<double>(k);
    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicAdd(p.work_counter, 1);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= p.ncols) break;
        double b[NC], x[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) b[t] = 0.0;
        for (int e = p.colptr[j]; e < p.colptr[j + 1]; ++e) {        // rhs.hpp:64-68, CSC order
            const double v = p.vals[e];
            const double* f = p.F + static_cast<size_t>(p.rowidx[e]) * KP;
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                const int c = lane + 32 * t;
                if (c < KP) b[t] = __dadd_rn(b[t], __dmul_rn(v, f[c]));
            }
        }
        double* xcol = p.X + static_cast<size_t>(j) * KP;
        if (p.warm) {                                                 // B -= G·h (gemv: tmp over columns, then subtract)
            double tmp[NC];
#pragma unroll
            for (int t = 0; t < NC; ++t) { const int c = lane + 32 * t; x[t] = (c < KP) ? xcol[c] : 0.0; tmp[t] = 0.0; }
            for (int i = 0; i < k; ++i) {
                double xi = 0.0;
#pragma unroll
                for (int t = 0; t < NC; ++t) if (t == (i >> 5)) xi = __shfl_sync(0xffffffffu, x[t], i & 31);
#pragma unroll
                for (int t = 0; t < NC; ++t) {
                    const int r = lane + 32 * t;
                    if (r < k) tmp[t] = __dadd_rn(tmp[t], __dmul_rn(sGd[i * KP + r], xi));
                }
            }
#pragma unroll
            for (int t = 0; t < NC; ++t) b[t] = __dsub_rn(b[t], tmp[t]);
        } else {
#pragma unroll
            for (int t = 0; t < NC; ++t) x[t] = 0.0;                  // nnls_batch.hpp:172-174
        }
        for (int it = 0; it < p.cd_maxit; ++it) {                     // nnls_batch.hpp:86-131
            double tol_sum = 0.0;
            for (int i = 0; i < k; ++i) {
                const int owner = i & 31, slot = i >> 5;
                double bi = 0.0, xi = 0.0;
#pragma unroll
                for (int t = 0; t < NC; ++t)
                    if (t == slot) {
                        bi = __shfl_sync(0xffffffffu, b[t], owner);
                        xi = __shfl_sync(0xffffffffu, x[t], owner);
                    }
                const double gd = sGd[i * KP + i];
                double ad = 0.0, xn = xi;
                if (gd > 0.0) {
                    double diff = __ddiv_rn(bi, gd);
                    if (p.L1 != 0.0) diff = __dsub_rn(diff, p.L1);    // :94
                    const double nval = __dadd_rn(xi, diff);
                    if (nonneg && nval < 0.0) { ad = -xi; xn = 0.0; }
                    else if (has_ub && nval > p.ub) { ad = __dsub_rn(p.ub, xi); xn = p.ub; }
                    else { ad = diff; xn = (diff == 0.0) ? xi : nval; }
                }
                if (ad != 0.0) {
                    if (check) tol_sum = __dadd_rn(tol_sum, __ddiv_rn(fabs(ad), __dadd_rn(fabs(xn), 1e-15)));
#pragma unroll
                    for (int t = 0; t < NC; ++t) {
                        if (t == slot && lane == owner) x[t] = xn;
                        const int r = lane + 32 * t;
                        if (r < k) b[t] = __dsub_rn(b[t], __dmul_rn(sGd[i * KP + r], ad));
                    }
                }
            }
            if (check && __dmul_rn(tol_sum, inv_k) < p.cd_tol) break;
        }
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int c = lane + 32 * t;
            if (c < KP) xcol[c] = (c < k) ? x[t] : 0.0;
        }
    }
}

// Per non-zero: pred = Σ_f (W(i,f)·d_f)·H(f,j) in fp64; accumulates Σ(a−pred)², Σ a·pred and Σ a².
static __global__ void __launch_bounds__(256) eval_nnz_kernel(const int* __restrict__ colptr,
                                                              const int* __restrict__ rowidx,
                                                              const double* __restrict__ vals, int ncols, int KP, int k,
                                                              const double* __restrict__ W_T,
                                                              const double* __restrict__ H,
                                                              const double* __restrict__ d,
                                                              double* __restrict__ partials /*[grid][3]*/) {
    __shared__ double s[3][256];
    double sq = 0.0, cross = 0.0, aa = 0.0;
    for (int j = blockIdx.x; j < ncols; j += gridDim.x) {
        const double* h = H + static_cast<size_t>(j) * KP;
        for (int e = colptr[j] + threadIdx.x; e < colptr[j + 1]; e += blockDim.x) {
            const double* w = W_T + static_cast<size_t>(rowidx[e]) * KP;
            double pred = 0.0;
            for (int f = 0; f < k; ++f) pred = fma(w[f] * d[f], h[f], pred);
            const double a = vals[e], r = a - pred;
            sq += r * r; cross += a * pred; aa += a * a;
        }
    }
    s[0][threadIdx.x] = sq; s[1][threadIdx.x] = cross; s[2][threadIdx.x] = aa;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int q = 0; q < 3; ++q) s[q][threadIdx.x] += s[q][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 3) partials[blockIdx.x * 3 + threadIdx.x] = s[threadIdx.x][0];
}

template <class T>
static __global__ void pad_f64_kernel(const T* __restrict__ src, double* __restrict__ dst, long long ncols, int k, int KP) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= ncols * KP) return;
    const long long c = e / KP;
    const int i = static_cast<int>(e % KP);
    dst[e] = (i < k) ? static_cast<double>(src[c * k + i]) : 0.0;
}
static __global__ void unpad_f64_kernel(const double* __restrict__ src, double* __restrict__ dst, long long ncols, int k, int KP) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= ncols * k) return;
    dst[e] = src[(e / k) * KP + (e % k)];
}

}  // namespace b200
