// kernels_cv.cuh — speckled-mask cross-validation NMF.
//
// Replaces (reference): nmf/fit_cv.hpp:409-476 (H update), :560-735 (W update), :1349-1548 (test / train
// loss), :1565-1621 (early stopping, convergence), nmf/speckled_cv.hpp:58-160 (lazy mask),
// nmf/cv_detail.hpp:67-85 (apply_gram_correction), :305-405 (compute_train_rhs[_W]).
//
// The hold-out mask is never stored: is_holdout(i, j) = SplitMix64::hash(seed, i, j) < UINT64_MAX / inv_prob
// is evaluated in-kernel (pure 64-bit integer arithmetic -> identical on CPU and GPU). Per column the kernel
// walks the entries in ascending order; a train entry is accumulated into b, a held-out entry triggers a
// rank-1 downdate of the warp's private copy of the Gram (G_local = G − Σ_test f fᵀ), then the column is
// solved with its own matrix — CD with L1 inside the sweep and NO tolerance (fit_cv.hpp:469-472), or a
// per-column LLT + clip (:463-467). One warp per column, G_local in shared memory (k×(k+1) floats).
#pragma once

#include "common.cuh"
#include "kernels_dense.cuh"
#include "kernels_masked.cuh"

namespace b200 {

__device__ __forceinline__ bool cv_is_holdout(unsigned long long seed, unsigned i, unsigned j,
                                              unsigned long long threshold, int enabled) {
    return enabled && splitmix_hash(seed, i, j) < threshold;                // rng.hpp:164-170
}

struct CvState {
    float prev_conv_loss;      // fit_cv.hpp:348
    float best_test_loss;      // :354
    int best_iter;
    int patience_count;
    float train_loss, test_loss;
    long long n_test;
};

struct CvParams {
    const int* __restrict__ colptr;    // sparse operand: A (H update) or A[I,:]ᵀ (W update), CSC
    const int* __restrict__ rowidx;
    const float* __restrict__ vals;
    const float* __restrict__ F;       // gathered factor [rows][KP]
    float* __restrict__ X;             // solved factor [ncols][KP] (in: warm x)
    const float* __restrict__ G;       // Gram + tiny + L2, KP×KP col-major
    int ncols;                         // columns of the operand
    int nrows;                         // rows of the operand (inner dimension)
    int k;
    int transposed;                    // 0: mask(i = inner, j = column)   1: mask(i = column, j = inner)
    int mask_zeros;
    unsigned long long seed, threshold;
    int holdout_enabled;
    float L1, ub;
    int cd_maxit, nonneg, solver, norm_type;
    int want_cross;                    // W update: also Σ <x, b_full> for the Gram-trick train loss
    int* work_counter;
    double* partials;                  // [gridDim.x][KP+1]
    DevState* state;
    // sharded fits: local column j is column / row j + col_offset of the whole matrix (the hold-out hash takes GLOBAL
    // indices) and row j + col_offset of the replicated factor X; solved columns also go into the peers' replicas
    int col_offset;
    float* peerX[7];
    int npeers;
    float* mcX;                        // multicast alias of X (see HalfStepParams::mcX); nullptr: local store
};

template <int KP>
__global__ void __launch_bounds__(384, 1) cv_half_step_kernel(const CvParams p) {   // blockDim = WARPS*32
    constexpr int NC = (KP + 31) / 32;
    constexpr int LD = KP + 1;
    constexpr int WARPS = (KP <= 64) ? 12 : 3;      // 12 x 16.9 KB of per-warp Gram copies at k = 64 (one CTA per SM)
    extern __shared__ __align__(16) float smem[];
    __shared__ double sred[12][KP + 1];
    if (p.state->stop) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Gl = smem + static_cast<size_t>(warp) * (KP * LD + KP);
    float* sf = Gl + KP * LD;
    const int k = p.k;

    double rs[NC];
#pragma unroll
    for (int t = 0; t < NC; ++t) rs[t] = 0.0;
    double cross = 0.0;
    int chol_fail = 0;

    auto held = [&](int inner, int col) {
        return p.transposed ? cv_is_holdout(p.seed, static_cast<unsigned>(col), static_cast<unsigned>(inner), p.threshold, p.holdout_enabled)
                            : cv_is_holdout(p.seed, static_cast<unsigned>(inner), static_cast<unsigned>(col), p.threshold, p.holdout_enabled);
    };
    auto accumulate = [&](float (&acc)[NC], float v, int r) {
        const float* f = p.F + static_cast<size_t>(r) * KP;
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int c = lane + 32 * t;
            if (c < KP) acc[t] = __fadd_rn(acc[t], __fmul_rn(v, __ldg(f + c)));
        }
    };

    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicAdd(p.work_counter, 1);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= p.ncols) break;
        const int p0 = p.colptr[j], p1 = p.colptr[j + 1];
        const int jg = j + p.col_offset;                            // global column (H update) / row (W update) index

        for (int e = lane; e < KP * KP; e += 32) Gl[(e / KP) * LD + (e % KP)] = p.G[e];      // cv_detail.hpp:74
        float b[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) b[t] = 0.f;

        // ---- pass 1: ascending walk; train entries -> b, held-out entries -> G_local downdate
        if (p.mask_zeros) {                                         // only the non-zeros can be held out
            for (int e0 = p0; e0 < p1; e0 += 32) {
                const int e = e0 + lane;
                int r = 0; float v = 0.f; bool h = false;
                if (e < p1) { r = __ldg(p.rowidx + e); v = __ldg(p.vals + e); h = held(r, jg); }
                const unsigned hb = __ballot_sync(0xffffffffu, h);
                const int cnt = min(32, p1 - e0);
                // The factor rows of 8 entries are requested together before any of them is consumed: a train
                // entry and a held-out entry both need their row, and one L2 round trip per entry, taken one
                // after the other, was what a column spent most of its time on.
                constexpr int PF = 8;
                for (int t0 = 0; t0 < cnt; t0 += PF) {
                    float fr[PF][NC];
#pragma unroll
                    for (int u = 0; u < PF; ++u) {
                        const int rt = __shfl_sync(0xffffffffu, r, (t0 + u) & 31);
                        const float* f = p.F + static_cast<size_t>(rt) * KP;
#pragma unroll
                        for (int t = 0; t < NC; ++t) {
                            const int c = lane + 32 * t;
                            fr[u][t] = (t0 + u < cnt && c < KP) ? __ldg(f + c) : 0.f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PF; ++u) {
                        if (t0 + u >= cnt) break;
                        const float vt = __shfl_sync(0xffffffffu, v, t0 + u);
                        if ((hb >> (t0 + u)) & 1u) {
                            warp_stage_row<KP>(sf, fr[u], lane);
                            if (p.solver == 1) warp_rank1_downdate_staged<KP, true>(Gl, sf, fr[u], k, lane);
                            else warp_rank1_downdate_staged<KP, false>(Gl, sf, fr[u], k, lane);
                        } else {
#pragma unroll
                            for (int t = 0; t < NC; ++t)
                                if (lane + 32 * t < KP) b[t] = __fadd_rn(b[t], __fmul_rn(vt, fr[u][t]));
                        }
                    }
                }
            }
        } else {                                                    // every (row, column) cell is hashed
            int e = p0;
            for (int i0 = 0; i0 < p.nrows; i0 += 32) {
                const int i = i0 + lane;
                const unsigned hb = __ballot_sync(0xffffffffu, i < p.nrows && held(i, jg));
                unsigned rest = hb;
                for (;;) {                                          // merge held-out rows and non-zero rows of this chunk
                    const int nh = rest ? i0 + __ffs(rest) - 1 : 0x7fffffff;
                    int nz = 0x7fffffff;
                    if (e < p1) { const int r = __ldg(p.rowidx + e); if (r < i0 + 32) nz = r; }
                    const int row = min(nh, nz);
                    if (row == 0x7fffffff) break;
                    const bool is_h = (row == nh);
                    if (is_h) {
                        if (p.solver == 1) warp_rank1_downdate_lower<KP>(Gl, sf, p.F + static_cast<size_t>(row) * KP, k, lane);
                        else warp_rank1_downdate<KP>(Gl, sf, p.F + static_cast<size_t>(row) * KP, k, lane);
                        rest &= rest - 1;
                        if (row == nz) ++e;                         // held-out non-zero: skipped in b
                    } else {
                        const float v = __ldg(p.vals + e);
                        if (v != 0.f) accumulate(b, v, row);
                        ++e;
                    }
                }
            }
        }
        __syncwarp();

        float btrain[NC];
        if (p.want_cross) {
#pragma unroll
            for (int t = 0; t < NC; ++t) btrain[t] = b[t];
        }
        float* xcol = p.X + static_cast<size_t>(jg) * KP;
        float x[NC];
        if (p.solver == 1) {                                        // cholesky_clip_col(G_local, b, x, k, L1, 0, nonneg, ...)
            if (p.L1 > 0.f) {
#pragma unroll
                for (int t = 0; t < NC; ++t) if (lane + 32 * t < k) b[t] = __fsub_rn(b[t], p.L1);
            }
            const int f = warp_chol_solve<KP>(Gl, b, k, lane);
            if (f && !chol_fail) chol_fail = f;
#pragma unroll
            for (int t = 0; t < NC; ++t) { float v = b[t]; if (p.nonneg && v < 0.f) v = 0.f; x[t] = v; }
        } else {                                                    // cd_nnls_col_fixed(G_local, b, x, k, L1, 0, nonneg, cd_maxit)
#pragma unroll
            for (int t = 0; t < NC; ++t) { const int c = lane + 32 * t; x[t] = (c < KP) ? xcol[c] : 0.f; }   // x_local = X.col(j)
            warp_cd_solve<KP>(Gl, b, x, k, p.L1, p.nonneg != 0, p.cd_maxit, 0.f, 0.f, lane);
        }
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int c = lane + 32 * t;
            if (c < KP) {
                float v = (c < k) ? x[t] : 0.f;
                if (p.ub > 0.f) v = fminf(v, p.ub);                 // fit_cv.hpp:528 / :843 (post-hoc)
                x[t] = v;
                if (p.mcX) multimem_store1(p.mcX + static_cast<size_t>(jg) * KP + c, v);
                else xcol[c] = v;
                for (int q2 = 0; q2 < p.npeers; ++q2) p.peerX[q2][static_cast<size_t>(jg) * KP + c] = v;
                if (p.norm_type == 0) rs[t] += static_cast<double>(fabsf(v));
                else if (p.norm_type == 1) rs[t] += static_cast<double>(v) * static_cast<double>(v);
            }
        }
        // ---- pass 2 (W update): b_full = b_train + held-out contributions in test order (fit_cv.hpp:609-653)
        if (p.want_cross) {
            if (p.mask_zeros) {
                for (int e0 = p0; e0 < p1; e0 += 32) {
                    const int e = e0 + lane;
                    int r = 0; float v = 0.f; bool h = false;
                    if (e < p1) { r = __ldg(p.rowidx + e); v = __ldg(p.vals + e); h = held(r, jg); }
                    unsigned hb = __ballot_sync(0xffffffffu, h);
                    while (hb) {
                        const int t = __ffs(hb) - 1;
                        hb &= hb - 1;
                        accumulate(btrain, __shfl_sync(0xffffffffu, v, t), __shfl_sync(0xffffffffu, r, t));
                    }
                }
            } else {
                for (int e = p0; e < p1; ++e) {
                    const int r = __ldg(p.rowidx + e);
                    const float v = __ldg(p.vals + e);
                    if (held(r, jg) && v != 0.f) accumulate(btrain, v, r);
                }
            }
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < NC; ++t) s += static_cast<double>(x[t]) * static_cast<double>(btrain[t]);
            cross += s;
        }
    }
#pragma unroll
    for (int t = 0; t < NC; ++t) {
        const int c = lane + 32 * t;
        if (c < KP) sred[warp][c] = rs[t];
    }
    for (int o = 16; o > 0; o >>= 1) cross += __shfl_xor_sync(0xffffffffu, cross, o);
    if (lane == 0) sred[warp][KP] = cross;
    if (chol_fail && lane == 0) atomicCAS(&p.state->chol_fail, 0, chol_fail);
    __syncthreads();
    for (int c = threadIdx.x; c < KP + 1; c += WARPS * 32) {
        double s = 0.0;
        for (int w = 0; w < WARPS; ++w) s += sred[w][c];
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + c] = s;
    }
}

// Σ over held-out entries of (a − <W_T[:,i]·d, H[:,j]>)² and their count (fit_cv.hpp:1453-1496).
// One CTA per column (grid-stride). partials[cta] = {sum, count}.
static __global__ void __launch_bounds__(256) cv_test_loss_kernel(const int* __restrict__ colptr,
                                                                  const int* __restrict__ rowidx,
                                                                  const float* __restrict__ vals, int ncols, int col_offset,
                                                                  int nrows, int KP, int k, int mask_zeros,
                                                                  unsigned long long seed, unsigned long long threshold,
                                                                  int holdout_enabled, const float* __restrict__ W_T,
                                                                  const float* __restrict__ H,
                                                                  const float* __restrict__ d,
                                                                  double* __restrict__ partials,
                                                                  const int* __restrict__ stop_flag) {
    __shared__ double ssum[256];
    __shared__ long long scnt[256];
    if (*stop_flag) return;
    double acc = 0.0;
    long long cnt = 0;
    auto sq_err = [&](int i, int j, float a) {
        const float* w = W_T + static_cast<size_t>(i) * KP;
        const float* h = H + static_cast<size_t>(j) * KP;
        double s = 0.0;
        for (int f = 0; f < k; ++f) s += static_cast<double>(__fmul_rn(w[f], d[f])) * static_cast<double>(h[f]);
        const float df = __fsub_rn(a, static_cast<float>(s));
        return static_cast<double>(__fmul_rn(df, df));
    };
    for (int jl = blockIdx.x; jl < ncols; jl += gridDim.x) {
        const int p0 = colptr[jl], p1 = colptr[jl + 1];
        const int j = jl + col_offset;                              // global column: hash argument and row of H
        if (mask_zeros) {
            for (int e = p0 + threadIdx.x; e < p1; e += blockDim.x) {
                const int i = rowidx[e];
                if (cv_is_holdout(seed, i, j, threshold, holdout_enabled)) { acc += sq_err(i, j, vals[e]); ++cnt; }
            }
        } else {
            for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
                if (!cv_is_holdout(seed, i, j, threshold, holdout_enabled)) continue;
                int lo = p0, hi = p1;                               // value of A(i, j), 0 when structurally zero
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (rowidx[mid] < i) lo = mid + 1; else hi = mid; }
                const float a = (lo < p1 && rowidx[lo] == i) ? vals[lo] : 0.f;
                acc += sq_err(i, j, a);
                ++cnt;
            }
        }
    }
    ssum[threadIdx.x] = acc; scnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { ssum[threadIdx.x] += ssum[threadIdx.x + w]; scnt[threadIdx.x] += scnt[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = ssum[0];
        partials[2 * blockIdx.x + 1] = static_cast<double>(scnt[0]);
    }
}

// Train loss by the Gram trick (fit_cv.hpp:1498-1542), test loss, early stopping and convergence (:1565-1621).
static __global__ void __launch_bounds__(256) cv_loss_finalize_kernel(const float* __restrict__ G_wnew,
                                                                      const float* __restrict__ G_hsaved,
                                                                      const float* __restrict__ d, int KP, int k,
                                                                      const double* __restrict__ cross_ptr,
                                                                      const double* __restrict__ test_partials, int nparts,
                                                                      float trAtA, long long total_entries, float tol,
                                                                      int cv_patience, float* __restrict__ train_hist,
                                                                      float* __restrict__ test_hist, int hist_cap,
                                                                      DevState* __restrict__ st, CvState* __restrict__ cv) {
    __shared__ double sred[256];
    if (st->stop) return;
    double r = 0.0;
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
        const int i = e % KP, j = e / KP;
        if (i < k && j < k) r += static_cast<double>(__fmul_rn(__fmul_rn(__fmul_rn(d[i], d[j]), G_wnew[e]), G_hsaved[e]));
    }
    sred[threadIdx.x] = r;
    __syncthreads();
    if (threadIdx.x != 0) return;
    double recon = 0.0;
    for (int t = 0; t < 256; ++t) recon += sred[t];
    double tsq = 0.0, tcnt = 0.0;
    for (int c = 0; c < nparts; ++c) { tsq += test_partials[2 * c]; tcnt += test_partials[2 * c + 1]; }
    const float test_sq_error = static_cast<float>(tsq);
    const long long n_test = static_cast<long long>(tcnt);
    const float total = fmaxf(__fadd_rn(__fsub_rn(trAtA, __fmul_rn(2.f, static_cast<float>(*cross_ptr))), static_cast<float>(recon)), 0.f);
    const float train_sq_error = fmaxf(__fsub_rn(total, test_sq_error), 0.f);
    const long long n_train = total_entries - n_test;
    const float train_loss = n_train > 0 ? __fdiv_rn(train_sq_error, static_cast<float>(n_train)) : 0.f;
    const float test_loss = n_test > 0 ? __fdiv_rn(test_sq_error, static_cast<float>(n_test)) : 0.f;
    const int iter = st->iter;
    if (iter < hist_cap) { train_hist[iter] = train_loss; test_hist[iter] = test_loss; }
    cv->train_loss = train_loss; cv->test_loss = test_loss; cv->n_test = n_test;
    st->train_loss = train_loss;
    float rel = 0.f;
    if (iter > 0) rel = __fdiv_rn(fabsf(__fsub_rn(cv->prev_conv_loss, test_loss)), __fadd_rn(fabsf(cv->prev_conv_loss), 1e-15f));
    if (test_loss < cv->best_test_loss) { cv->best_test_loss = test_loss; cv->best_iter = iter; cv->patience_count = 0; }
    else cv->patience_count++;
    st->iter = iter + 1;
    if (cv_patience > 0 && cv->patience_count >= cv_patience) { st->converged = 0; st->stop = 1; return; }
    if (iter > 0) {
        st->final_tol = rel;
        if (rel < tol) { st->converged = 1; st->stop = 1; return; }
    }
    cv->prev_conv_loss = test_loss;
}

// H[:, j] *= d (fit_cv.hpp:1639-1641, "absorb d into H for final output").
static __global__ void cv_absorb_d_kernel(float* __restrict__ H, long long ncols, int KP, const float* __restrict__ d) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < ncols * KP) H[e] = __fmul_rn(H[e], d[e % KP]);
}

}  // namespace b200
