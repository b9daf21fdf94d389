// kernels_masked.cuh — the explicit-user-mask path.
//
// Replaces (reference): nmf/masked_nnls.hpp:97-154 (masked_nnls_h), :178-242 (masked_nnls_w),
// :46-67 (masked_solve_col), primitives/cpu/cholesky_clip.hpp:65-106 (cholesky_clip_col) and
// nmf/masked_nnls.hpp:251-282 (masked_loss).
//
// Per column j: b sums only the non-masked non-zeros; the Gram is corrected per column,
// G_local = G − Σ_{r∈mask(j)} f_r f_rᵀ, then b −= L1, diag += L2, and the column is solved with ITS OWN
// matrix (CD, or a per-column LLT). One warp per column; G_local lives in the warp's shared memory
// (k×(k+1) floats), a lane owns coordinates lane, lane+32, …  Every operation keeps the CPU path's
// order and roundings (separate mul / sub, IEEE div / sqrt), including its quirks: the warm start
// copies x but does NOT correct b (masked_nnls.hpp:146-151), and iteration 0 starts from x = 0.
#pragma once

#include "common.cuh"
#include "kernels_dense.cuh"

namespace b200 {

// ---- warp-cooperative dense solvers on a per-warp k×k matrix in shared memory (col-major, ld = KP+1) ----
// A lane owns coordinates lane, lane+32, … (NC of them). Used by the explicit-mask and the CV kernels.

// cd_nnls_col_fixed (nnls_batch.hpp:71-132) with upper_bound = 0, L2 = 0. L1 != 0 acts inside the sweep (:94).
template <int KP>
__device__ __forceinline__ int warp_cd_solve(const float* Gl, float (&b)[(KP + 31) / 32], float (&x)[(KP + 31) / 32],
                                             int k, float L1, bool nonneg, int maxit, float cd_tol, float inv_k,
                                             int lane) {
    constexpr int NC = (KP + 31) / 32;
    constexpr int LD = KP + 1;
    const bool check = cd_tol > 0.f;
    for (int it = 0; it < maxit; ++it) {
        float tol_sum = 0.f;
        for (int i = 0; i < k; ++i) {
            const int owner = i & 31, slot = i >> 5;
            float bi = 0.f, xi = 0.f;
#pragma unroll
            for (int t = 0; t < NC; ++t)
                if (t == slot) {
                    bi = __shfl_sync(0xffffffffu, b[t], owner);
                    xi = __shfl_sync(0xffffffffu, x[t], owner);
                }
            const float gd = Gl[i * LD + i];
            float ad = 0.f, xn = xi;
            if (gd > 0.f) {
                float diff = __fdiv_rn(bi, gd);
                if (L1 != 0.f) diff = __fsub_rn(diff, L1);
                const float nval = __fadd_rn(xi, diff);
                if (nonneg && nval < 0.f) { ad = -xi; xn = 0.f; }
                else { ad = diff; xn = (diff == 0.f) ? xi : nval; }
            }
            if (ad != 0.f) {
                if (check) tol_sum = __fadd_rn(tol_sum, __fdiv_rn(fabsf(ad), __fadd_rn(fabsf(xn), 1e-15f)));
#pragma unroll
                for (int t = 0; t < NC; ++t) {
                    if (t == slot && lane == owner) x[t] = xn;
                    const int row = lane + 32 * t;
                    if (row < k) b[t] = __fsub_rn(b[t], __fmul_rn(Gl[i * LD + row], ad));
                }
            }
        }
        if (check && __fmul_rn(tol_sum, inv_k) < cd_tol) return it + 1;
    }
    return maxit;
}

// Eigen::LLT restated (oracle order): L(i,j) = (G(i,j) − t_ij) / L(j,j), t_ij = Σ_{p<j} L(i,p)·L(j,p) accumulated
// sequentially in p (separately rounded mul and add), L(j,j) = sqrt(G(j,j) − t_jj); then x := L⁻ᵀ L⁻¹ b by
// column-oriented substitution with IEEE division. b holds the rhs on entry and x on exit. Returns the first
// non-positive pivot index + 1 (0 = ok).
// The dots are kept as RUNNING sums (like prepare_solver_kernel): finished column j adds L(i,j)·L(c,j) to T(i,c) for
// every pair j < c <= i at once — the same additions in the same order as the left-looking loop, but the k³/6
// updates are independent of each other instead of forming k²/2 dependent chains of length j (the left-looking
// form made a warp wait ~130 K cycles per column on those chains). Only the lower triangle of Gl is read by the
// factorisation and the substitutions, so T lives in the unused upper triangle: T(i,c), i > c, at Gl[i*LD + c];
// T(c,c) in the padding row, Gl[c*LD + KP].
template <int KP>
__device__ __forceinline__ int warp_chol_solve(float* Gl, float (&b)[(KP + 31) / 32], int k, int lane) {
    constexpr int NC = (KP + 31) / 32;
    constexpr int LD = KP + 1;
    int fail = 0;
    for (int c = 0; c < k; ++c) {                                   // T = 0
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int i = lane + 32 * t;
            if (i > c && i < k) Gl[i * LD + c] = 0.f;
            else if (i == c) Gl[c * LD + KP] = 0.f;
        }
    }
    __syncwarp();
    for (int jj = 0; jj < k; ++jj) {
        const float xx = __fsub_rn(Gl[jj * LD + jj], Gl[jj * LD + KP]);
        float ljj = 0.f;
        if (!(xx > 0.f)) { if (!fail) fail = jj + 1; } else ljj = __fsqrt_rn(xx);
        float lij[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int i = lane + 32 * t;
            lij[t] = 0.f;
            if (i > jj && i < k) lij[t] = __fdiv_rn(__fsub_rn(Gl[jj * LD + i], Gl[i * LD + jj]), ljj);
        }
        __syncwarp();                                               // every lane has read G(jj,jj)
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int i = lane + 32 * t;
            if (i > jj && i < k) Gl[jj * LD + i] = lij[t];
            if (i == jj) { Gl[jj * LD + jj] = ljj; lij[t] = ljj; }
        }
        __syncwarp();                                               // L(:, jj) visible
        for (int c = jj + 1; c < k; ++c) {                          // T(i,c) += L(i,jj)·L(c,jj), c <= i
            const float lc = Gl[jj * LD + c];
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                if (32 * t + 31 < c) continue;                      // warp-uniform: no row of this pass reaches c
                const int i = lane + 32 * t;
                if (i >= c && i < k) {
                    float* tp = (i == c) ? (Gl + c * LD + KP) : (Gl + i * LD + c);
                    *tp = __fadd_rn(*tp, __fmul_rn(lij[t], lc));
                }
            }
        }
        __syncwarp();                                               // T(jj+1, jj+1) visible to every lane
    }
    for (int pp = 0; pp < k; ++pp) {
        const int owner = pp & 31, slot = pp >> 5;
        float bp = 0.f;
#pragma unroll
        for (int t = 0; t < NC; ++t) if (t == slot) bp = __shfl_sync(0xffffffffu, b[t], owner);
        const float y = __fdiv_rn(bp, Gl[pp * LD + pp]);
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int i = lane + 32 * t;
            if (t == slot && lane == owner) b[t] = y;
            else if (i > pp && i < k) b[t] = __fsub_rn(b[t], __fmul_rn(Gl[pp * LD + i], y));
        }
    }
    for (int pp = k - 1; pp >= 0; --pp) {
        const int owner = pp & 31, slot = pp >> 5;
        float yp = 0.f;
#pragma unroll
        for (int t = 0; t < NC; ++t) if (t == slot) yp = __shfl_sync(0xffffffffu, b[t], owner);
        const float xp = __fdiv_rn(yp, Gl[pp * LD + pp]);
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int i = lane + 32 * t;
            if (t == slot && lane == owner) b[t] = xp;
            else if (i < pp) b[t] = __fsub_rn(b[t], __fmul_rn(Gl[i * LD + pp], xp));   // L(pp, i)
        }
    }
    return fail;
}

// Gl -= f fᵀ for one factor row f (staged through sf), separately rounded, column by column.
template <int KP>
__device__ __forceinline__ void warp_rank1_downdate(float* Gl, float* sf, const float* __restrict__ f, int k, int lane) {
    constexpr int NC = (KP + 31) / 32;
    constexpr int LD = KP + 1;
    __syncwarp();
    for (int c = lane; c < KP; c += 32) sf[c] = __ldg(f + c);
    __syncwarp();
    for (int col = 0; col < k; ++col) {
        const float fc = sf[col];
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int row = lane + 32 * t;
            if (row < k) Gl[col * LD + row] = __fsub_rn(Gl[col * LD + row], __fmul_rn(sf[row], fc));
        }
    }
    __syncwarp();
}

// Stage a factor row that is already in registers (a lane holds coordinates lane, lane+32, ...) into sf.
template <int KP>
__device__ __forceinline__ void warp_stage_row(float* sf, const float (&fr)[(KP + 31) / 32], int lane) {
    constexpr int NC = (KP + 31) / 32;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NC; ++t) if (lane + 32 * t < KP) sf[lane + 32 * t] = fr[t];
    __syncwarp();
}
// Gl -= f fᵀ with f already staged in sf (full square / lower triangle only).
template <int KP, bool LOWER>
__device__ __forceinline__ void warp_rank1_downdate_staged(float* Gl, const float* sf, const float (&fr)[(KP + 31) / 32], int k, int lane) {
    constexpr int NC = (KP + 31) / 32;
    constexpr int LD = KP + 1;
    for (int col = 0; col < k; ++col) {
        const float fc = sf[col];
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            if (LOWER && 32 * t + 31 < col) continue;
            const int row = lane + 32 * t;
            if (row < k && (!LOWER || row >= col)) Gl[col * LD + row] = __fsub_rn(Gl[col * LD + row], __fmul_rn(fr[t], fc));
        }
    }
    __syncwarp();
}

// The same downdate restricted to the lower triangle (rows >= column) — all a per-column LLT reads
// (cv_detail.hpp:80-84 updates the Lower view and mirrors it). Same products, same roundings.
template <int KP>
__device__ __forceinline__ void warp_rank1_downdate_lower(float* Gl, float* sf, const float* __restrict__ f, int k, int lane) {
    constexpr int NC = (KP + 31) / 32;
    constexpr int LD = KP + 1;
    __syncwarp();
    for (int c = lane; c < KP; c += 32) sf[c] = __ldg(f + c);
    __syncwarp();
    float fr[NC];
#pragma unroll
    for (int t = 0; t < NC; ++t) fr[t] = (lane + 32 * t < KP) ? sf[lane + 32 * t] : 0.f;
    for (int col = 0; col < k; ++col) {
        const float fc = sf[col];
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            if (32 * t + 31 < col) continue;
            const int row = lane + 32 * t;
            if (row >= col && row < k) Gl[col * LD + row] = __fsub_rn(Gl[col * LD + row], __fmul_rn(fr[t], fc));
        }
    }
    __syncwarp();
}

struct MaskedParams {
    const int* __restrict__ colptr;    // sparse operand (A or Aᵀ), CSC
    const int* __restrict__ rowidx;
    const float* __restrict__ vals;
    const int* __restrict__ mptr;      // mask pattern in the same orientation, CSC (sorted rows)
    const int* __restrict__ midx;
    const float* __restrict__ F;       // gathered factor [rows][KP]
    float* __restrict__ X;             // solved factor   [ncols][KP]
    const float* __restrict__ G;       // unmodified Gram, KP×KP col-major (fit_cpu.hpp:562 / :801)
    int ncols;
    int k;
    float L1, L2, ub, cd_tol, inv_k;
    int cd_maxit;
    int nonneg;
    int warm;
    int solver;                        // 0 CD, else per-column Cholesky + clip
    int norm_type;
    int* work_counter;
    double* partials;                  // [gridDim.x][KP+1] (norm sums; slot KP unused here)
    DevState* state;
    unsigned long long* sweep_counter;
    // sharded fits: local column j is row j + col_offset of the replicated factor X; solved columns are also stored
    // into the peers' replicas (NVLink P2P), like half_step_kernel does
    int col_offset;
    float* peerX[7];
    int npeers;
    float* mcX;                        // multicast alias of X (see HalfStepParams::mcX); nullptr: local store
};

template <int KP>
__global__ void __launch_bounds__(384, 1) masked_half_step_kernel(const MaskedParams p) {   // blockDim = WARPS*32
    constexpr int NC = (KP + 31) / 32;            // coordinates per lane
    constexpr int LD = KP + 1;                     // padded leading dimension: row reads are conflict-free
    constexpr int WARPS = (KP <= 64) ? 12 : 3;      // 12 x 16.9 KB of per-warp Gram copies at k = 64 (one CTA per SM)
    extern __shared__ __align__(16) float smem[];
    if (p.state->stop) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Gl = smem + static_cast<size_t>(warp) * (KP * LD + KP);
    float* sf = Gl + KP * LD;                      // one factor row (broadcast source)
    const int k = p.k;

    double rs[NC];
#pragma unroll
    for (int t = 0; t < NC; ++t) rs[t] = 0.0;
    unsigned long long my_sweeps = 0;
    int chol_fail = 0;

    for (;;) {
        int j = 0;
        if (lane == 0) j = atomicAdd(p.work_counter, 1);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= p.ncols) break;
        const int p0 = p.colptr[j], p1 = p.colptr[j + 1];
        const int mb = p.mptr[j], me = p.mptr[j + 1];

        // ---- b over the non-masked non-zeros, CSC order (masked_nnls.hpp:121-133); two-pointer merge
        float b[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) b[t] = 0.f;
        int q = mb;
        for (int e = p0; e < p1; ++e) {
            const int r = __ldg(p.rowidx + e);
            while (q < me && __ldg(p.midx + q) < r) ++q;
            if (q < me && __ldg(p.midx + q) == r) continue;            // masked entry
            const float v = __ldg(p.vals + e);
            const float* f = p.F + static_cast<size_t>(r) * KP;
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                const int c = lane + 32 * t;
                if (c < KP) b[t] = __fadd_rn(b[t], __fmul_rn(v, __ldg(f + c)));
            }
        }
        // ---- G_local = G_full − Σ_{r masked} f_r f_rᵀ (masked_nnls.hpp:136-138)
        for (int e = lane; e < KP * KP; e += 32) Gl[(e / KP) * LD + (e % KP)] = p.G[e];
        __syncwarp();
        for (int e = mb; e < me; ++e)
            if (p.solver == 0) warp_rank1_downdate<KP>(Gl, sf, p.F + static_cast<size_t>(__ldg(p.midx + e)) * KP, k, lane);
            else warp_rank1_downdate_lower<KP>(Gl, sf, p.F + static_cast<size_t>(__ldg(p.midx + e)) * KP, k, lane);
        // ---- L1 / L2 (masked_nnls.hpp:141-144): unconditional, like the reference
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int c = lane + 32 * t;
            if (c < k) {
                b[t] = __fsub_rn(b[t], p.L1);
                Gl[c * LD + c] = __fadd_rn(Gl[c * LD + c], p.L2);
            }
        }
        __syncwarp();

        float* xcol = p.X + static_cast<size_t>(j + p.col_offset) * KP;
        float x[NC];
        if (p.solver == 0) {
            // ---- cd_nnls_col_fixed(G_local, b, x, k, 0, 0, nonneg, maxit, 0, cd_tol)  (masked_nnls.hpp:62-65)
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                const int c = lane + 32 * t;
                x[t] = (p.warm && c < KP) ? xcol[c] : 0.f;          // :146-148 (b is NOT corrected)
            }
            my_sweeps += warp_cd_solve<KP>(Gl, b, x, k, 0.f, p.nonneg != 0, p.cd_maxit, p.cd_tol, p.inv_k, lane);
        } else {
            // ---- cholesky_clip_col (cholesky_clip.hpp:65-106): per-column LLT of G_local
            const int f = warp_chol_solve<KP>(Gl, b, k, lane);
            if (f && !chol_fail) chol_fail = f;
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                float v = b[t];
                if (p.nonneg && v < 0.f) v = 0.f;
                x[t] = v;
            }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            const int c = lane + 32 * t;
            if (c < KP) {
                float v = (c < k) ? x[t] : 0.f;
                if (p.ub > 0.f) v = fminf(v, p.ub);                         // fit_cpu.hpp:636 / :884 (post-hoc)
                if (p.mcX) multimem_store1(p.mcX + static_cast<size_t>(j + p.col_offset) * KP + c, v);
                else xcol[c] = v;
                for (int q2 = 0; q2 < p.npeers; ++q2) p.peerX[q2][static_cast<size_t>(j + p.col_offset) * KP + c] = v;
                if (p.norm_type == 0) rs[t] += static_cast<double>(fabsf(v));
                else if (p.norm_type == 1) rs[t] += static_cast<double>(v) * static_cast<double>(v);
            }
        }
    }
    // per-CTA partial row sums in a fixed order
    __shared__ double sred[12][KP];
#pragma unroll
    for (int t = 0; t < NC; ++t) {
        const int c = lane + 32 * t;
        if (c < KP) sred[warp][c] = rs[t];
    }
    if (chol_fail && lane == 0) atomicCAS(&p.state->chol_fail, 0, chol_fail);
    if (p.sweep_counter && lane == 0 && my_sweeps) atomicAdd(p.sweep_counter, my_sweeps);
    __syncthreads();
    for (int c = threadIdx.x; c < KP + 1; c += WARPS * 32) {
        double s = 0.0;
        if (c < KP)
            for (int w = 0; w < WARPS; ++w) s += sred[w][c];
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + c] = s;
    }
}

template <int KP>
inline int masked_warps() { return (KP <= 64) ? 12 : 3; }
template <int KP>
inline size_t masked_smem_bytes() {
    return static_cast<size_t>(masked_warps<KP>()) * (KP * (KP + 1) + KP) * sizeof(float);
}

// masked_loss (masked_nnls.hpp:251-282): Σ over NON-MASKED NON-ZEROS of (a − <W_T[:,i]·d, H[:,j]>)², the
// prediction accumulated sequentially in fp32 like the reference's inner loop (:274-276).
static __global__ void __launch_bounds__(256) masked_loss_kernel(const int* __restrict__ colptr,
                                                                 const int* __restrict__ rowidx,
                                                                 const float* __restrict__ vals,
                                                                 const int* __restrict__ mptr,
                                                                 const int* __restrict__ midx, int ncols, int col_offset,
                                                                 int KP, int k, const float* __restrict__ W_T,
                                                                 const float* __restrict__ H,
                                                                 const float* __restrict__ d,
                                                                 double* __restrict__ partials,
                                                                 const int* __restrict__ stop_flag) {
    __shared__ double sred[256];
    if (*stop_flag) return;
    double acc = 0.0;
    for (int j = blockIdx.x; j < ncols; j += gridDim.x) {
        const int p0 = colptr[j], p1 = colptr[j + 1], mb = mptr[j], me = mptr[j + 1];
        const float* h = H + static_cast<size_t>(j + col_offset) * KP;
        for (int e = p0 + threadIdx.x; e < p1; e += blockDim.x) {
            const int r = rowidx[e];
            int lo = mb, hi = me;                               // binary search of r in the mask column
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (midx[mid] < r) lo = mid + 1; else hi = mid;
            }
            if (lo < me && midx[lo] == r) continue;
            const float* w = W_T + static_cast<size_t>(r) * KP;
            float pred = 0.f;
            for (int f = 0; f < k; ++f) pred = __fadd_rn(pred, __fmul_rn(__fmul_rn(w[f], d[f]), h[f]));
            const float res = __fsub_rn(vals[e], pred);
            acc += static_cast<double>(__fmul_rn(res, res));
        }
    }
    sred[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sred[threadIdx.x] += sred[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sred[0];
}

// Convergence bookkeeping for a loss value that is already a sum (masked path): fit_cpu.hpp:1769-1811.
static __global__ void masked_loss_finalize_kernel(const double* __restrict__ partials, int nparts, float tol,
                                                   int patience, float* __restrict__ loss_hist, int hist_cap,
                                                   DevState* __restrict__ st) {
    if (st->stop || threadIdx.x != 0) return;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += partials[c];
    const float loss = static_cast<float>(s);
    const int iter = st->iter;
    if (iter < hist_cap) loss_hist[iter] = loss;
    bool loss_conv = false;
    if (iter > 0) {
        const float rel = __fdiv_rn(fabsf(__fsub_rn(st->prev_loss, loss)), __fadd_rn(fabsf(st->prev_loss), 1e-15f));
        st->final_tol = rel;
        if (rel < tol) loss_conv = true;
    }
    st->prev_loss = loss;
    st->train_loss = loss;
    st->iter = iter + 1;
    if (iter > 0) {
        if (loss_conv) {
            if (++st->patience_counter >= patience) { st->converged = 1; st->stop = 1; }
        } else {
            st->patience_counter = 0;
        }
    }
}

}  // namespace b200
