// oracle/nmf_oracle.cpp — TEST INFRASTRUCTURE ONLY. NOT PRODUCT CODE.
//
// Eigen-free CPU restatement of the reference's sparse-MSE NMF ALS path
// (zdebruine/RcppML @ df69ddd, CPU backend). Every function cites the
// reference file:line it follows (paths relative to /root/reference/).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference leg may build, load or call this file. The product path
// (rcppml_b200/csrc) never links or calls it.
//
// PARITY STATUS: pinned against the reference's OWN nmf_fit and nmf_fit_cv —
// fit_cpu.hpp / fit_cv.hpp compiled unmodified against an Eigen stand-in (oracle/ref_hotpath/ref_fit.cpp ->
// oracle/_ref/libref_fit.so); tests/test_reference_fit.py: W, d, H of this file are
// bit-identical to it (CD / Cholesky, L1/L2, bounds, norms, masks, sorting,
// patience). **parity unpinned** only for the rounding inside Eigen (the stand-in
// shares this file's definitions) and for the fp32-vs-fp64 loss accumulation.
//  * The reference's hot-path headers (rng.hpp, nnls_batch.hpp, fused_nnls.hpp,
//    cholesky_clip.hpp, gram.hpp, primitives.hpp, constants.hpp) compile
//    unmodified from /root/reference against a minimal stand-in for the Eigen
//    types they use (oracle/ref_hotpath, `make -C oracle ref_hotpath` ->
//    oracle/_ref/libref_hotpath.so); tests/test_reference_sources.py holds this
//    file against that library BIT FOR BIT: SplitMix64 / hash / is_holdout /
//    fill_uniform / factor initialisation, cd_nnls_col_fixed with every switch
//    (sweep counts included), nnls_batch, the fused CD and Cholesky column loops
//    (L1, warm start, clip, bound), the Gram wrapper, the explicit-mask half-steps,
//    extract_scaling, apply_upper_bound, LazySpeckledMask, and the CV building
//    blocks compute_train_rhs / compute_train_rhs_W / apply_gram_correction.
//  * Eigen itself (and R/Rcpp) is not in the image, so the whole nmf_fit<> cannot
//    be built and the reference's test-suite holds no golden W/d/H vectors
//    (SURVEY.md §8c): the order of the reductions INSIDE Eigen (rankUpdate, gemv,
//    LLT, dot) is this file's definition (below), which the stand-in shares.
//  * Also pinned (tests/test_oracle_kats.py): the known-answer tests of
//    tests/cpp/test_nnls.cpp, test_gram.cpp, test_rng.cpp, the published
//    SplitMix64 vectors, and frozen outputs on the reference's movielens matrix.
//
// Arithmetic conventions of this restatement (documented in DESIGN.md §3):
//  * fp32 everywhere the reference is fp32; compiled with -ffp-contract=off
//    (the package builds without -march, i.e. SSE2, no FMA: src/Makevars:6).
//  * Element-wise Eigen expressions whose evaluation order is defined by the
//    source (b += v*col in CSC order, the CD residual axpy) are restated in
//    exactly that order and are bit-reproducible.
//  * Reductions whose order is hidden inside Eigen (rankUpdate Gram,
//    rowwise().sum(), dot, OpenMP reduction(+)) are restated as
//    fp64-accumulated, then rounded once to fp32 — the order-independent
//    definition any accurate implementation agrees with to ~1 ulp.
//  * Eigen gemv `b -= G*x` is restated as tmp = sum_i G(:,i)*x_i (sequential
//    in i, fp32), b -= tmp — the order of Eigen's col-major gemv kernel.
//  * Eigen::LLT is restated as the unblocked left-looking factorisation
//    (dot-then-subtract, as Eigen's llt_inplace::unblocked) and column
//    oriented (axpy) forward/backward substitution with IEEE division.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ---------------------------------------------------------------------------
// rng/rng.hpp:60-221 — SplitMix64
// ---------------------------------------------------------------------------
struct SplitMix64 {
    uint64_t state;
    explicit SplitMix64(uint64_t seed) : state(seed == 0 ? 12345ULL : seed) {}  // rng.hpp:73
    uint64_t next() {                                                           // rng.hpp:89-95
        state += 0x9e3779b97f4a7c15ULL;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    template <class T> T uniform() {                                            // rng.hpp:102-104
        return static_cast<T>(next()) / static_cast<T>(UINT64_MAX);
    }
    static uint64_t hash(uint64_t seed, uint32_t i, uint32_t j) {               // rng.hpp:129-138
        uint64_t h = seed + static_cast<uint64_t>(i) * 0x9e3779b97f4a7c15ULL +
                     static_cast<uint64_t>(j) * 0x6c62272e07bb0142ULL;
        h = (h ^ (h >> 30)) * 0xbf58476d1ce4e5b9ULL;
        h = (h ^ (h >> 27)) * 0x94d049bb133111ebULL;
        return h ^ (h >> 31);
    }
    static bool is_holdout(uint64_t seed, uint32_t i, uint32_t j, uint64_t inv_prob) {  // rng.hpp:164-170
        if (inv_prob == 0) return false;
        return hash(seed, i, j) < (UINT64_MAX / inv_prob);
    }
    template <class T> void fill_uniform(T* data, long rows, long cols) {       // rng.hpp:195-201
        for (long j = 0; j < cols; ++j)
            for (long i = 0; i < rows; ++i) data[j * rows + i] = uniform<T>();
    }
};

constexpr double CD_ABS_TOL = 1e-15;   // core/constants.hpp:76

// ---------------------------------------------------------------------------
// primitives/cpu/gram.hpp:58-67 — G = F·Fᵀ (+ tiny_num on the diagonal)
// Order-opaque in Eigen (selfadjointView::rankUpdate) → fp64 accumulate.
// ---------------------------------------------------------------------------
template <class T>
void gram(const T* F, int k, long n, T* G, int threads) {
    const long kk = static_cast<long>(k) * k;
    int nt = std::max(1, threads);
    std::vector<double> acc(static_cast<size_t>(nt) * kk, 0.0);
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double* a = acc.data() + static_cast<size_t>(tid) * kk;
#pragma omp for schedule(static)
        for (long c = 0; c < n; ++c) {
            const T* f = F + c * k;
            for (int j = 0; j < k; ++j) {
                const double fj = static_cast<double>(f[j]);
                double* aj = a + static_cast<long>(j) * k;
                for (int i = j; i < k; ++i) aj[i] += static_cast<double>(f[i]) * fj;  // lower triangle
            }
        }
    }
    for (int j = 0; j < k; ++j)
        for (int i = j; i < k; ++i) {
            double s = 0.0;
            for (int t = 0; t < nt; ++t) s += acc[static_cast<size_t>(t) * kk + static_cast<long>(j) * k + i];
            T v = static_cast<T>(s);
            G[static_cast<long>(j) * k + i] = v;
            G[static_cast<long>(i) * k + j] = v;   // gram.hpp:65 mirror
        }
    for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += static_cast<T>(1e-15);  // gram.hpp:66
}

// ---------------------------------------------------------------------------
// primitives/cpu/nnls_batch.hpp:71-132 — cd_nnls_col_fixed (verbatim semantics)
// G is k×k column-major with leading dimension k.
// ---------------------------------------------------------------------------
template <class T>
int cd_nnls_col_fixed(const T* G, T* b, T* x, int k, T L1, T L2, bool nonneg, int maxit,
                      T upper_bound, T cd_tol) {
    const bool has_upper = (upper_bound > 0);
    const bool check_convergence = (cd_tol > 0);
    const T inv_k = T(1) / static_cast<T>(k);
    for (int iter = 0; iter < maxit; ++iter) {
        T tol_sum = 0;
        for (int i = 0; i < k; ++i) {
            const T g_diag = G[static_cast<long>(i) * k + i];
            if (g_diag <= T(0)) continue;
            T diff = b[i] / g_diag;
            if (L1 != 0) diff -= L1;
            if (L2 != 0) diff += L2 * x[i];
            T new_val = x[i] + diff;
            T actual_diff;
            if (nonneg && new_val < T(0)) {
                actual_diff = -x[i];
                if (actual_diff == T(0)) continue;
                x[i] = T(0);
            } else if (has_upper && new_val > upper_bound) {
                actual_diff = upper_bound - x[i];
                if (actual_diff == T(0)) continue;
                x[i] = upper_bound;
            } else {
                if (diff == T(0)) continue;
                actual_diff = diff;
                x[i] = new_val;
            }
            if (check_convergence) {
                const T abs_diff = (actual_diff >= 0) ? actual_diff : -actual_diff;
                tol_sum += abs_diff / (std::abs(x[i]) + static_cast<T>(CD_ABS_TOL));
            }
            const T* g_col = G + static_cast<long>(i) * k;
            for (int r = 0; r < k; ++r) b[r] -= g_col[r] * actual_diff;
        }
        if (check_convergence && tol_sum * inv_k < cd_tol) return iter + 1;
    }
    return maxit;
}

// ---------------------------------------------------------------------------
// Eigen::LLT restated (see header). L is k×k column-major, lower triangle.
// Returns 0 on success, j+1 if the pivot at column j is not positive
// (Eigen: info()=NumericalIssue; the reference never checks it and goes on —
// we stop and leave the remaining columns untouched, flagged to the caller).
// ---------------------------------------------------------------------------
template <class T>
int cholesky_factor(const T* G, int k, T* L) {
    for (long e = 0; e < static_cast<long>(k) * k; ++e) L[e] = T(0);
    for (int j = 0; j < k; ++j) {
        T s = T(0);
        for (int p = 0; p < j; ++p) {
            const T l = L[static_cast<long>(p) * k + j];
            s += l * l;
        }
        T x = G[static_cast<long>(j) * k + j] - s;
        if (!(x > T(0))) return j + 1;
        x = std::sqrt(x);
        L[static_cast<long>(j) * k + j] = x;
        for (int i = j + 1; i < k; ++i) {
            T t = T(0);
            for (int p = 0; p < j; ++p) t += L[static_cast<long>(p) * k + i] * L[static_cast<long>(p) * k + j];
            L[static_cast<long>(j) * k + i] = (G[static_cast<long>(j) * k + i] - t) / x;
        }
    }
    return 0;
}

// x := L⁻ᵀ L⁻¹ x, column-oriented substitution, IEEE division.
template <class T>
void cholesky_solve_inplace(const T* L, int k, T* x) {
    for (int p = 0; p < k; ++p) {               // forward: L y = b
        const T* lc = L + static_cast<long>(p) * k;
        const T y = x[p] / lc[p];
        x[p] = y;
        for (int i = p + 1; i < k; ++i) x[i] -= lc[i] * y;
    }
    for (int p = k - 1; p >= 0; --p) {          // backward: Lᵀ x = y
        const T xp = x[p] / L[static_cast<long>(p) * k + p];
        x[p] = xp;
        for (int i = 0; i < p; ++i) x[i] -= L[static_cast<long>(i) * k + p] * xp;
    }
}

// b = Σ_{p∈col j} x[p]·F[:, i[p]] in CSC order — fused_nnls.hpp:111-114
template <class T>
inline void gather_rhs(const int* Ap, const int* Ai, const T* Ax, long j, const T* F, int k, T* b) {
    for (int t = 0; t < k; ++t) b[t] = T(0);
    for (long p = Ap[j]; p < Ap[j + 1]; ++p) {
        const T v = Ax[p];
        const T* f = F + static_cast<long>(Ai[p]) * k;
        for (int t = 0; t < k; ++t) b[t] += v * f[t];
    }
}

// b -= G·x restated as Eigen's col-major gemv (tmp accumulated over columns).
template <class T>
inline void warm_start_correct(const T* G, const T* x, int k, T* b, T* tmp) {
    for (int r = 0; r < k; ++r) tmp[r] = T(0);
    for (int i = 0; i < k; ++i) {
        const T xi = x[i];
        const T* gc = G + static_cast<long>(i) * k;
        for (int r = 0; r < k; ++r) tmp[r] += gc[r] * xi;
    }
    for (int r = 0; r < k; ++r) b[r] -= tmp[r];
}

// ---------------------------------------------------------------------------
// primitives/cpu/fused_nnls.hpp:71-134 — fused_rhs_nnls_sparse (solver_mode 0)
// ---------------------------------------------------------------------------
template <class T>
void fused_rhs_nnls_sparse(const int* Ap, const int* Ai, const T* Ax, long n, const T* F, int k,
                           const T* G, T* X, int cd_maxit, T cd_tol, T L1, bool nonneg, int threads,
                           bool warm_start, T upper_bound, long* sweeps_out) {
    const bool has_L1 = (L1 > T(0));
    long sweeps_total = 0;
#pragma omp parallel num_threads(std::max(1, threads)) reduction(+ : sweeps_total)
    {
        std::vector<T> b(k), tmp(k);
#pragma omp for schedule(dynamic)
        for (long j = 0; j < n; ++j) {
            gather_rhs(Ap, Ai, Ax, j, F, k, b.data());                     // :111-114
            if (has_L1) for (int t = 0; t < k; ++t) b[t] -= L1;            // :117
            T* xj = X + j * k;
            if (warm_start) warm_start_correct(G, xj, k, b.data(), tmp.data());  // :121-123
            sweeps_total += cd_nnls_col_fixed(G, b.data(), xj, k, T(0), T(0), nonneg, cd_maxit,
                                              upper_bound, cd_tol);        // :126-131
        }
    }
    if (sweeps_out) *sweeps_out = sweeps_total;
}

// ---------------------------------------------------------------------------
// primitives/cpu/fused_nnls.hpp:156-221 — fused_rhs_cholesky_sparse (modes 1,2)
// ---------------------------------------------------------------------------
template <class T>
int fused_rhs_cholesky_sparse(const int* Ap, const int* Ai, const T* Ax, long n, const T* F, int k,
                              const T* G, T* X, T L1, bool nonneg, int threads, T upper_bound) {
    std::vector<T> L(static_cast<size_t>(k) * k);
    const int info = cholesky_factor(G, k, L.data());                       // :185
    const bool has_L1 = (L1 > T(0));
    const bool do_upper = (upper_bound > T(0));
#pragma omp parallel num_threads(std::max(1, threads))
    {
        std::vector<T> b(k);
#pragma omp for schedule(dynamic)
        for (long j = 0; j < n; ++j) {
            gather_rhs(Ap, Ai, Ax, j, F, k, b.data());                     // :196-199
            if (has_L1) for (int t = 0; t < k; ++t) b[t] -= L1;            // :202
            cholesky_solve_inplace(L.data(), k, b.data());                 // :210
            T* xj = X + j * k;
            for (int t = 0; t < k; ++t) {
                T v = b[t];
                if (nonneg) v = std::max(v, T(0));                          // :212-214
                if (do_upper) v = std::min(v, upper_bound);                 // :216
                xj[t] = v;                                                  // :218
            }
        }
    }
    return info;
}

// ---------------------------------------------------------------------------
// nmf/variant_helpers.hpp:287-305 — extract_scaling. norm_type: 0=L1 1=L2 2=None
// ---------------------------------------------------------------------------
template <class T>
void extract_scaling(T* X, int k, long n, T* d, int norm_type, int threads) {
    if (norm_type == 2) { for (int i = 0; i < k; ++i) d[i] = T(1); return; }
    int nt = std::max(1, threads);
    std::vector<double> acc(static_cast<size_t>(nt) * k, 0.0);
#pragma omp parallel num_threads(nt)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double* a = acc.data() + static_cast<size_t>(tid) * k;
#pragma omp for schedule(static)
        for (long j = 0; j < n; ++j) {
            const T* x = X + j * k;
            if (norm_type == 0) for (int i = 0; i < k; ++i) a[i] += std::abs(static_cast<double>(x[i]));
            else                for (int i = 0; i < k; ++i) a[i] += static_cast<double>(x[i]) * static_cast<double>(x[i]);
        }
    }
    for (int i = 0; i < k; ++i) {
        double s = 0.0;
        for (int t = 0; t < nt; ++t) s += acc[static_cast<size_t>(t) * k + i];
        T di = static_cast<T>(s);
        if (norm_type == 1) di = std::sqrt(di);
        d[i] = di + static_cast<T>(1e-15);                                   // :301
    }
#pragma omp parallel for num_threads(nt) schedule(static)
    for (long j = 0; j < n; ++j) {
        T* x = X + j * k;
        for (int i = 0; i < k; ++i) x[i] /= d[i];                            // :302-304
    }
}

// features/bounds.hpp:38 — apply_upper_bound
template <class T>
void apply_upper_bound(T* X, long count, T ub) {
    for (long e = 0; e < count; ++e) X[e] = std::min(X[e], ub);
}

// primitives/primitives.hpp:101-115 — trace_AtA. The reference accumulates in
// Scalar (fp32) sequentially, which saturates beyond 2^24 terms of O(1); the
// restatement accumulates in fp64 (order-independent definition).
template <class T>
T trace_AtA(const T* Ax, long nnz) {
    double s = 0.0;
    for (long p = 0; p < nnz; ++p) s += static_cast<double>(Ax[p]) * static_cast<double>(Ax[p]);
    return static_cast<T>(s);
}

// ---------------------------------------------------------------------------
// primitives/cpu/fused_nnls.hpp:306-362 — loss_cross_term_sparse_via_At
// ---------------------------------------------------------------------------
template <class T>
T loss_cross_term_via_At(const int* Atp, const int* Ati, const T* Atx, long m, const T* W_T, const T* H,
                         const T* d, int k, int threads) {
    double cross = 0.0;
#pragma omp parallel num_threads(std::max(1, threads)) reduction(+ : cross)
    {
        std::vector<T> h_at(k);
#pragma omp for schedule(static)
        for (long ell = 0; ell < m; ++ell) {
            gather_rhs(Atp, Ati, Atx, ell, H, k, h_at.data());              // :342-345
            const T* w = W_T + ell * k;
            double s = 0.0;
            for (int i = 0; i < k; ++i) {
                const T u = w[i] * d[i];                                     // :327-329 U = diag(d)·W_T
                s += static_cast<double>(u) * static_cast<double>(h_at[i]);  // :347
            }
            cross += s;
        }
    }
    return static_cast<T>(cross);
}

// Sparse transpose with ascending inner indices (Eigen SparseMatrix::transpose,
// nmf/fit_cpu.hpp:251-253). Stable counting sort.
template <class T>
void transpose_csc(const int* Ap, const int* Ai, const T* Ax, long m, long n, int* Tp, int* Ti, T* Tx) {
    const long nnz = Ap[n];
    std::vector<long> cnt(m + 1, 0);
    for (long p = 0; p < nnz; ++p) cnt[Ai[p] + 1]++;
    for (long r = 0; r < m; ++r) cnt[r + 1] += cnt[r];
    for (long r = 0; r <= m; ++r) Tp[r] = static_cast<int>(cnt[r]);
    for (long j = 0; j < n; ++j)
        for (long p = Ap[j]; p < Ap[j + 1]; ++p) {
            const long q = cnt[Ai[p]]++;
            Ti[q] = static_cast<int>(j);
            Tx[q] = Ax[p];
        }
}

// ---------------------------------------------------------------------------
// nmf/masked_nnls.hpp:97-154 / 178-242 — masked_nnls_h / masked_nnls_w.
// One routine: "data" is A (H-update) or Aᵀ (W-update); "mask" is the mask in
// the same orientation (pattern only; stored entries are masked when their
// value is non-zero — the caller passes the pattern of non-zero mask values).
// ---------------------------------------------------------------------------
template <class T>
void masked_nnls(const int* Ap, const int* Ai, const T* Ax, long n, long m_rows, const T* F, int k,
                 const T* G_full, T* X, const int* Mp, const int* Mi, T L1, T L2, bool nonneg,
                 int cd_maxit, T cd_tol, int solver_mode, int threads, bool warm_start) {
#pragma omp parallel num_threads(std::max(1, threads))
    {
        std::vector<T> b(k), x(k), Gl(static_cast<size_t>(k) * k), Lc(static_cast<size_t>(k) * k);
        std::vector<char> is_masked;
#pragma omp for schedule(dynamic, 64)
        for (long j = 0; j < n; ++j) {
            const long mb = Mp[j], me = Mp[j + 1];
            for (int t = 0; t < k; ++t) b[t] = T(0);
            if (mb == me) {
                gather_rhs(Ap, Ai, Ax, j, F, k, b.data());                  // :123-125
            } else {
                is_masked.assign(m_rows, 0);                                // :127-128
                for (long q = mb; q < me; ++q) is_masked[Mi[q]] = 1;
                for (long p = Ap[j]; p < Ap[j + 1]; ++p) {                  // :129-132
                    if (is_masked[Ai[p]]) continue;
                    const T v = Ax[p];
                    const T* f = F + static_cast<long>(Ai[p]) * k;
                    for (int t = 0; t < k; ++t) b[t] += v * f[t];
                }
            }
            std::copy(G_full, G_full + static_cast<long>(k) * k, Gl.begin());   // :136
            for (long q = mb; q < me; ++q) {                                // :137-138
                const T* f = F + static_cast<long>(Mi[q]) * k;
                for (int c = 0; c < k; ++c)
                    for (int r = 0; r < k; ++r) Gl[static_cast<long>(c) * k + r] -= f[r] * f[c];
            }
            for (int i = 0; i < k; ++i) {                                   // :141-144
                b[i] -= L1;
                Gl[static_cast<long>(i) * k + i] += L2;
            }
            T* xj = X + j * k;
            if (warm_start) std::copy(xj, xj + k, x.begin());               // :146-148
            else std::fill(x.begin(), x.end(), T(0));
            if (solver_mode == 1) {                                         // :56-60 → cholesky_clip.hpp:65-106
                cholesky_factor(Gl.data(), k, Lc.data());
                std::copy(b.begin(), b.end(), x.begin());
                cholesky_solve_inplace(Lc.data(), k, x.data());
                if (nonneg) for (int t = 0; t < k; ++t) x[t] = std::max(x[t], T(0));
            } else {                                                        // :62-65
                cd_nnls_col_fixed(Gl.data(), b.data(), x.data(), k, T(0), T(0), nonneg, cd_maxit, T(0), cd_tol);
            }
            std::copy(x.begin(), x.end(), xj);                              // :152
        }
    }
}

// nmf/masked_nnls.hpp:251-282 — masked_loss (MSE): Σ over unmasked NON-ZEROS of A.
template <class T>
T masked_loss(const int* Ap, const int* Ai, const T* Ax, long m, long n, const T* W_Td, const T* H, int k,
              const int* Mp, const int* Mi, int threads) {
    double total = 0.0;
#pragma omp parallel num_threads(std::max(1, threads)) reduction(+ : total)
    {
        std::vector<char> is_masked;
#pragma omp for schedule(static)
        for (long j = 0; j < n; ++j) {
            is_masked.assign(m, 0);
            for (long q = Mp[j]; q < Mp[j + 1]; ++q) is_masked[Mi[q]] = 1;
            for (long p = Ap[j]; p < Ap[j + 1]; ++p) {
                if (is_masked[Ai[p]]) continue;
                const T* w = W_Td + static_cast<long>(Ai[p]) * k;
                const T* h = H + j * k;
                T pred = 0;
                for (int f = 0; f < k; ++f) pred += w[f] * h[f];            // :274-276 (fp32 sequential)
                const T r = Ax[p] - pred;
                total += static_cast<double>(r * r);
            }
        }
    }
    return static_cast<T>(total);
}

}  // namespace orc

// ===========================================================================
// C interface (ctypes)
// ===========================================================================
static int orc_max_threads_impl() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

extern "C" {

typedef struct {
    int k;
    int max_iter;
    float tol;
    float L1_W, L1_H, L2_W, L2_H, ub_W, ub_H;
    int nonneg_W, nonneg_H;
    int cd_maxit;
    float cd_tol;
    int norm_type;     // 0=L1 1=L2 2=None (core/config.hpp NormType)
    int solver_mode;   // 0=CD, !=0 Cholesky+clip in the fused path
    int patience;
    int threads;
    int sort_model;
    int has_mask;
    double time_budget_s;   // > 0: stop after the first iteration that ends past the budget (bench only)
} orc_config;

typedef struct {
    int iterations;
    int converged;
    float train_loss;
    float final_tol;
    int chol_info;          // first non-positive Cholesky pivot seen (0 = none)
    double loop_seconds;    // wall time of the iteration loop only
    long cd_sweeps;         // total CD sweeps over all columns (solver_mode 0)
    double* iter_seconds;   // optional [max_iter]: wall time at the end of each iteration (since loop start)
} orc_result;

void orc_splitmix_next(uint64_t seed, int count, uint64_t* out) {
    orc::SplitMix64 r(seed);
    for (int i = 0; i < count; ++i) out[i] = r.next();
}
uint64_t orc_splitmix_hash(uint64_t seed, uint32_t i, uint32_t j) { return orc::SplitMix64::hash(seed, i, j); }
int orc_is_holdout(uint64_t seed, uint32_t i, uint32_t j, uint64_t inv_prob) {
    return orc::SplitMix64::is_holdout(seed, i, j, inv_prob) ? 1 : 0;
}
// Fills `count` uniforms continuing from *state (pass the seed-remapped state).
void orc_fill_uniform_f32(uint64_t* state, float* out, long count) {
    orc::SplitMix64 r(1); r.state = *state;
    for (long e = 0; e < count; ++e) out[e] = r.uniform<float>();
    *state = r.state;
}
void orc_fill_uniform_f64(uint64_t* state, double* out, long count) {
    orc::SplitMix64 r(1); r.state = *state;
    for (long e = 0; e < count; ++e) out[e] = r.uniform<double>();
    *state = r.state;
}
void orc_gram_f32(const float* F, int k, long n, float* G, int threads) { orc::gram(F, k, n, G, threads); }
void orc_gram_f64(const double* F, int k, long n, double* G, int threads) { orc::gram(F, k, n, G, threads); }

int orc_cd_nnls_col_f32(const float* G, float* b, float* x, int k, float L1, float L2, int nonneg, int maxit,
                        float ub, float cd_tol) {
    return orc::cd_nnls_col_fixed(G, b, x, k, L1, L2, nonneg != 0, maxit, ub, cd_tol);
}
int orc_cd_nnls_col_f64(const double* G, double* b, double* x, int k, double L1, double L2, int nonneg, int maxit,
                        double ub, double cd_tol) {
    return orc::cd_nnls_col_fixed(G, b, x, k, L1, L2, nonneg != 0, maxit, ub, cd_tol);
}
// nnls_batch<CPU,double> (nnls_batch.hpp:150-185): warm start B -= G·X else X=0, then CD per column.
void orc_nnls_batch_f64(const double* G, double* B, double* X, int k, long n, int cd_maxit, double cd_tol, double L1,
                        double L2, int nonneg, double ub, int warm_start) {
    std::vector<double> tmp(k);
    for (long j = 0; j < n; ++j) {
        if (warm_start) orc::warm_start_correct(G, X + j * k, k, B + j * k, tmp.data());
        else for (int i = 0; i < k; ++i) X[j * k + i] = 0.0;
        orc::cd_nnls_col_fixed(G, B + j * k, X + j * k, k, L1, L2, nonneg != 0, cd_maxit, ub, cd_tol);
    }
}
int orc_cholesky_factor_f32(const float* G, int k, float* L) { return orc::cholesky_factor(G, k, L); }
void orc_cholesky_solve_f32(const float* L, int k, float* x) { orc::cholesky_solve_inplace(L, k, x); }

void orc_transpose_csc_f32(const int* Ap, const int* Ai, const float* Ax, long m, long n, int* Tp, int* Ti, float* Tx) {
    orc::transpose_csc(Ap, Ai, Ax, m, n, Tp, Ti, Tx);
}
void orc_extract_scaling_f32(float* X, int k, long n, float* d, int norm_type, int threads) {
    orc::extract_scaling(X, k, n, d, norm_type, threads);
}
float orc_trace_AtA_f32(const float* Ax, long nnz) { return orc::trace_AtA(Ax, nnz); }

// One fused half-step (H-update when given A and W_T; W-update when given Aᵀ and H).
// Returns Cholesky info (0 ok) for solver_mode != 0, else 0.
int orc_half_step_f32(const int* Ap, const int* Ai, const float* Ax, long n_cols, const float* F, int k,
                      const float* G, float* X, int solver_mode, int cd_maxit, float cd_tol, float L1, int nonneg,
                      int warm_start, float ub_in_solver, int threads, long* sweeps_out) {
    if (solver_mode == 0) {
        orc::fused_rhs_nnls_sparse(Ap, Ai, Ax, n_cols, F, k, G, X, cd_maxit, cd_tol, L1, nonneg != 0, threads,
                                   warm_start != 0, ub_in_solver, sweeps_out);
        return 0;
    }
    if (sweeps_out) *sweeps_out = 0;
    return orc::fused_rhs_cholesky_sparse(Ap, Ai, Ax, n_cols, F, k, G, X, L1, nonneg != 0, threads, ub_in_solver);
}

// Raw RHS B[:, j] = F·A[:, j] for all columns (primitives/cpu/rhs.hpp:52-70); used by the
// sharded (multi-rank) restatement where partial right-hand sides are summed across ranks.
void orc_rhs_f32(const int* Ap, const int* Ai, const float* Ax, long n_cols, const float* F, int k, float* B,
                 int threads) {
#pragma omp parallel for num_threads(std::max(1, threads)) schedule(dynamic, 64)
    for (long j = 0; j < n_cols; ++j) orc::gather_rhs(Ap, Ai, Ax, j, F, k, B + j * k);
}

// Solve step of a half-step with the right-hand side already formed (B is consumed).
int orc_solve_given_rhs_f32(float* B, long n_cols, int k, const float* G, float* X, int solver_mode, int cd_maxit,
                            float cd_tol, float L1, int nonneg, int warm_start, int threads) {
    std::vector<float> L;
    int info = 0;
    if (solver_mode != 0) { L.resize(static_cast<size_t>(k) * k); info = orc::cholesky_factor(G, k, L.data()); }
#pragma omp parallel num_threads(std::max(1, threads))
    {
        std::vector<float> tmp(k);
#pragma omp for schedule(dynamic, 64)
        for (long j = 0; j < n_cols; ++j) {
            float* b = B + j * k;
            float* x = X + j * k;
            if (L1 > 0.f) for (int t = 0; t < k; ++t) b[t] -= L1;
            if (solver_mode == 0) {
                if (warm_start) orc::warm_start_correct(G, x, k, b, tmp.data());
                orc::cd_nnls_col_fixed(G, b, x, k, 0.f, 0.f, nonneg != 0, cd_maxit, 0.f, cd_tol);
            } else {
                orc::cholesky_solve_inplace(L.data(), k, b);
                for (int t = 0; t < k; ++t) x[t] = nonneg ? std::max(b[t], 0.f) : b[t];
            }
        }
    }
    return info;
}

void orc_masked_nnls_f32(const int* Ap, const int* Ai, const float* Ax, long n, long m_rows, const float* F, int k,
                         const float* G_full, float* X, const int* Mp, const int* Mi, float L1, float L2, int nonneg,
                         int cd_maxit, float cd_tol, int solver_mode, int threads, int warm_start) {
    orc::masked_nnls(Ap, Ai, Ax, n, m_rows, F, k, G_full, X, Mp, Mi, L1, L2, nonneg != 0, cd_maxit, cd_tol,
                     solver_mode, threads, warm_start != 0);
}

float orc_loss_cross_term_f32(const int* Atp, const int* Ati, const float* Atx, long m, const float* W_T,
                              const float* H, const float* d, int k, int threads) {
    return orc::loss_cross_term_via_At(Atp, Ati, Atx, m, W_T, H, d, k, threads);
}

// ---------------------------------------------------------------------------
// nmf/fit_cpu.hpp:172-1855 — nmf_fit<CPU,float,SparseMatrix<float>>, sparse MSE
// standard variant (no projective/symmetric/IRLS/graph/L21/angular/target).
// W_T (k×m col-major) and H (k×n) hold the initial factors on entry (the
// caller reproduces fit_cpu.hpp:195-218) and the result on exit. d is output.
// mask (optional, CSC pattern m×n) selects the masked path (:560-564, :799-810).
// loss_history (optional) must hold max_iter floats.
// ---------------------------------------------------------------------------
int orc_nmf_fit_f32(const int* Ap, const int* Ai, const float* Ax, long m, long n, const orc_config* cfg,
                    float* W_T, float* H, float* d, const int* Mp, const int* Mi, float* loss_history,
                    orc_result* res) {
    using namespace orc;
    const int k = cfg->k;
    if (k <= 0 || cfg->max_iter <= 0 || cfg->tol < 0 || cfg->cd_maxit <= 0) return -1;   // core/config.hpp:421-432
    const long nnz = Ap[n];
    const int threads = cfg->threads > 0 ? cfg->threads :
#ifdef _OPENMP
        omp_get_max_threads();
#else
        1;
#endif
    const bool use_mask = cfg->has_mask && Mp && Mi;
    for (int i = 0; i < k; ++i) d[i] = 1.f;                                  // :198 / nmf_init.hpp:181

    const float trAtA = trace_AtA(Ax, nnz);                                  // :224
    std::vector<int> Atp(m + 1), Ati(nnz);
    std::vector<float> Atx(nnz);
    transpose_csc(Ap, Ai, Ax, m, n, Atp.data(), Ati.data(), Atx.data());     // :251-253
    std::vector<int> MTp, MTi;
    if (use_mask) {                                                          // :273-278
        const long mnnz = Mp[n];
        MTp.resize(m + 1); MTi.resize(mnnz);
        std::vector<float> ones(mnnz, 1.f), onesT(mnnz);
        transpose_csc(Mp, Mi, ones.data(), m, n, MTp.data(), MTi.data(), onesT.data());
    }
    std::vector<float> G(static_cast<size_t>(k) * k), G_w_saved(static_cast<size_t>(k) * k),
        G_wt(static_cast<size_t>(k) * k);
    float prev_loss = std::numeric_limits<float>::max();                     // :281
    int patience_counter = 0;
    res->iterations = 0; res->converged = 0; res->train_loss = 0.f; res->final_tol = 0.f;
    res->chol_info = 0; res->cd_sweeps = 0;
    const auto t0 = std::chrono::high_resolution_clock::now();

    for (int iter = 0; iter < cfg->max_iter; ++iter) {                       // :444
        long sw = 0;
        // ---- H update ----
        gram(W_T, k, m, G.data(), threads);                                  // :491
        if (!use_mask) {
            if (cfg->L2_H > 0) for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += cfg->L2_H;   // :506
            if (cfg->solver_mode == 0) {                                     // :516-524
                fused_rhs_nnls_sparse(Ap, Ai, Ax, n, W_T, k, G.data(), H, cfg->cd_maxit, cfg->cd_tol, cfg->L1_H,
                                      cfg->nonneg_H != 0, threads, iter > 0, 0.f, &sw);
                res->cd_sweeps += sw;
            } else {                                                         // :527-534
                int info = fused_rhs_cholesky_sparse(Ap, Ai, Ax, n, W_T, k, G.data(), H, cfg->L1_H,
                                                     cfg->nonneg_H != 0, threads, 0.f);
                if (info && !res->chol_info) res->chol_info = info;
            }
        } else {                                                             // :560-564 (G rebuilt unmodified :562)
            masked_nnls(Ap, Ai, Ax, n, m, W_T, k, G.data(), H, Mp, Mi, cfg->L1_H, cfg->L2_H, cfg->nonneg_H != 0,
                        cfg->cd_maxit, cfg->cd_tol, cfg->solver_mode, threads, iter > 0);
        }
        if (cfg->ub_H > 0) apply_upper_bound(H, static_cast<long>(k) * n, cfg->ub_H);    // :636-637
        extract_scaling(H, k, n, d, cfg->norm_type, threads);                // :644

        // ---- W update ----
        gram(H, k, n, G.data(), threads);                                    // :715
        if (!use_mask) G_w_saved = G;                                        // :719-722
        if (!use_mask) {
            if (cfg->L2_W > 0) for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += cfg->L2_W;   // :738
            if (cfg->solver_mode == 0) {                                     // :748-756
                fused_rhs_nnls_sparse(Atp.data(), Ati.data(), Atx.data(), m, H, k, G.data(), W_T, cfg->cd_maxit,
                                      cfg->cd_tol, cfg->L1_W, cfg->nonneg_W != 0, threads, iter > 0, 0.f, &sw);
                res->cd_sweeps += sw;
            } else {                                                         // :759-766
                int info = fused_rhs_cholesky_sparse(Atp.data(), Ati.data(), Atx.data(), m, H, k, G.data(), W_T,
                                                     cfg->L1_W, cfg->nonneg_W != 0, threads, 0.f);
                if (info && !res->chol_info) res->chol_info = info;
            }
        } else {                                                             // :799-810 (G rebuilt :801)
            masked_nnls(Atp.data(), Ati.data(), Atx.data(), m, n, H, k, G.data(), W_T, MTp.data(), MTi.data(),
                        cfg->L1_W, cfg->L2_W, cfg->nonneg_W != 0, cfg->cd_maxit, cfg->cd_tol, cfg->solver_mode,
                        threads, iter > 0);
        }
        if (cfg->ub_W > 0) apply_upper_bound(W_T, static_cast<long>(k) * m, cfg->ub_W);  // :884-885
        extract_scaling(W_T, k, m, d, cfg->norm_type, threads);              // :892

        // ---- loss ----
        float loss_val;
        if (use_mask) {                                                      // :1686-1691
            std::vector<float> W_Td(static_cast<size_t>(k) * m);
            for (long c = 0; c < m; ++c)
                for (int i = 0; i < k; ++i) W_Td[c * k + i] = W_T[c * k + i] * d[i];
            loss_val = masked_loss(Ap, Ai, Ax, m, n, W_Td.data(), H, k, Mp, Mi, threads);
        } else {                                                             // :1729-1753
            gram(W_T, k, m, G_wt.data(), threads);                           // :1735
            const float cross = loss_cross_term_via_At(Atp.data(), Ati.data(), Atx.data(), m, W_T, H, d, k,
                                                       threads);             // :1740
            double recon = 0.0;                                              // :1748-1751
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j)
                    recon += static_cast<double>(d[i] * d[j] * G_wt[static_cast<long>(j) * k + i] *
                                                 G_w_saved[static_cast<long>(j) * k + i]);
            loss_val = trAtA - 2.f * cross + static_cast<float>(recon);      // :1753
        }
        if (loss_history) loss_history[iter] = loss_val;

        bool loss_converged = false;                                         // :1769-1776
        if (iter > 0) {
            const float rel = std::abs(prev_loss - loss_val) / (std::abs(prev_loss) + 1e-15f);
            res->final_tol = rel;
            if (rel < cfg->tol) loss_converged = true;
        }
        prev_loss = loss_val;
        if (iter > 0) {                                                      // :1797-1809
            if (loss_converged) {
                if (++patience_counter >= cfg->patience) {
                    res->converged = 1;
                    res->train_loss = prev_loss;
                    res->iterations = iter + 1;
                    break;
                }
            } else patience_counter = 0;
        }
        res->iterations = iter + 1;                                          // :1811
        if (res->iter_seconds || cfg->time_budget_s > 0) {
            const double el = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
            if (res->iter_seconds) res->iter_seconds[iter] = el;
            if (cfg->time_budget_s > 0 && el > cfg->time_budget_s) break;
        }
    }
    res->loop_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    if (res->train_loss == 0.f) res->train_loss = prev_loss;                 // :1842-1845

    if (cfg->sort_model) {                                                   // :1847-1848, core/result.hpp:169-189
        std::vector<int> idx(k);
        std::iota(idx.begin(), idx.end(), 0);
        std::sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] > d[b]; });
        std::vector<float> tmp(static_cast<size_t>(k) * std::max(m, n)), dd(k);
        for (long c = 0; c < m; ++c) for (int i = 0; i < k; ++i) tmp[c * k + i] = W_T[c * k + idx[i]];
        std::copy(tmp.begin(), tmp.begin() + static_cast<long>(k) * m, W_T);
        for (long c = 0; c < n; ++c) for (int i = 0; i < k; ++i) tmp[c * k + i] = H[c * k + idx[i]];
        std::copy(tmp.begin(), tmp.begin() + static_cast<long>(k) * n, H);
        for (int i = 0; i < k; ++i) dd[i] = d[idx[i]];
        std::copy(dd.begin(), dd.end(), d);
    }
    return 0;
}

// SURVEY.md §8d synthetic generator (host twin of csrc/kernels_sparse.cuh synth_column_kernel;
// rcppml_b200/synth.py is the numpy twin). Two passes: counts, then fill. Rows >= m_keep are
// dropped (m_keep = m keeps everything): the leading m_keep × n_local block is the CPU sample.
long orc_synth_csc(long m, long n_local, long col_begin, double density, uint64_t seed, long m_keep, int pass,
                   int* Ap, int* Ai, float* Ax, int threads) {
    const long cnt = std::llround(static_cast<double>(m) * density);
    if (pass == 0) Ap[0] = 0;
#pragma omp parallel num_threads(std::max(1, threads))
    {
        std::vector<uint32_t> rows(cnt);
#pragma omp for schedule(static)
        for (long jl = 0; jl < n_local; ++jl) {
            const uint32_t j = static_cast<uint32_t>(col_begin + jl);
            for (long t = 0; t < cnt; ++t)
                rows[t] = static_cast<uint32_t>(orc::SplitMix64::hash(seed, static_cast<uint32_t>(t), j) %
                                                static_cast<uint64_t>(m));
            std::sort(rows.begin(), rows.end());
            const long u = std::unique(rows.begin(), rows.end()) - rows.begin();
            long kept = 0;
            while (kept < u && static_cast<long>(rows[kept]) < m_keep) ++kept;
            if (pass == 0) {
                Ap[jl + 1] = static_cast<int>(kept);
            } else {
                long w = Ap[jl];
                for (long t = 0; t < kept; ++t, ++w) {
                    Ai[w] = static_cast<int>(rows[t]);
                    const uint64_t h = orc::SplitMix64::hash(seed + 1, rows[t], j);
                    Ax[w] = 0.5f + static_cast<float>(h) / static_cast<float>(UINT64_MAX);
                }
            }
        }
    }
    if (pass == 0) {
        long total = 0;
        for (long jl = 0; jl < n_local; ++jl) { const long c = Ap[jl + 1]; Ap[jl + 1] = static_cast<int>(total += c); }
        return total;
    }
    return Ap[n_local];
}

// ---------------------------------------------------------------------------
// src/RcppFunctions_utils.cpp:23-53 (Rcpp_predict) and :314-366 (c_nnls), fp64.
// w_T: k×m col-major; h: k×n (in: warm start when warm_start, out: solution).
// ---------------------------------------------------------------------------
void orc_project_f64(const int* Ap, const int* Ai, const double* Ax, long m, long n, int k, const double* w_T,
                     double* h, double L1, double L2, double upper_bound, int nonneg, int cd_maxit, double cd_tol,
                     int warm_start, int threads) {
    std::vector<double> G(static_cast<size_t>(k) * k);
    orc::gram(w_T, k, m, G.data(), threads);                                   // :32 (adds tiny_num once)
    for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += 1e-15;      // :33 / :327 tiny_num again
    if (L2 > 0) for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += L2;
#pragma omp parallel num_threads(std::max(1, threads))
    {
        std::vector<double> b(k), tmp(k);
#pragma omp for schedule(dynamic, 16)
        for (long j = 0; j < n; ++j) {
            orc::gather_rhs(Ap, Ai, Ax, j, w_T, k, b.data());                  // rhs.hpp:64-68
            double* x = h + j * k;
            if (warm_start) {                                                  // :346-356 — no cd_tol passed
                orc::warm_start_correct(G.data(), x, k, b.data(), tmp.data());
                orc::cd_nnls_col_fixed(G.data(), b.data(), x, k, L1, 0.0, nonneg != 0, cd_maxit, upper_bound, 0.0);
            } else {                                                           // nnls_batch.hpp:167-184
                for (int i = 0; i < k; ++i) x[i] = 0.0;
                orc::cd_nnls_col_fixed(G.data(), b.data(), x, k, L1, 0.0, nonneg != 0, cd_maxit, upper_bound, cd_tol);
            }
        }
    }
}

// src/RcppFunctions_utils.cpp:60-90 (compute_mse): explicit over the non-zeros (mask_zeros) or over all m·n
// entries through the dense reconstruction W·diag(d)·H, fp64. Dense form only for small shapes.
double orc_evaluate_mse_f64(const int* Ap, const int* Ai, const double* Ax, long m, long n, int k, const double* w_T,
                            const double* d, const double* h, int mask_zeros) {
    double total = 0.0;
    if (mask_zeros) {
        long cnt = 0;
        for (long j = 0; j < n; ++j)
            for (long p = Ap[j]; p < Ap[j + 1]; ++p) {
                double pred = 0.0;
                for (int f = 0; f < k; ++f) pred += w_T[static_cast<long>(Ai[p]) * k + f] * d[f] * h[j * k + f];
                const double r = Ax[p] - pred;
                total += r * r;
                ++cnt;
            }
        return cnt > 0 ? total / static_cast<double>(cnt) : 0.0;
    }
    std::vector<double> col(m);
    for (long j = 0; j < n; ++j) {
        for (long i = 0; i < m; ++i) {
            double pred = 0.0;
            for (int f = 0; f < k; ++f) pred += w_T[i * k + f] * d[f] * h[j * k + f];
            col[i] = -pred;
        }
        for (long p = Ap[j]; p < Ap[j + 1]; ++p) col[Ai[p]] += Ax[p];
        for (long i = 0; i < m; ++i) total += col[i] * col[i];
    }
    return total / (static_cast<double>(m) * static_cast<double>(n));
}

// ---------------------------------------------------------------------------
// nmf/fit_cv.hpp:124-1667 — nmf_fit_cv<CPU,float,Sparse>: speckled-mask cross-validation NMF, restricted to
// the sparse / MSE / standard-variant / no-user-mask path. The lazy mask is
// nmf/speckled_cv.hpp:58-160 (SplitMix64 position hash, integer — bit-exact everywhere).
// Order-opaque pieces: the per-column Gram correction (Eigen rankUpdate(-1), cv_detail.hpp:67-85) is restated
// as sequential fp32 rank-1 downdates in test-row order; dots / norms / loss sums are fp64-accumulated.
// ---------------------------------------------------------------------------
typedef struct {
    int k, max_iter;
    float tol;
    float L1_W, L1_H, L2_W, L2_H, ub_W, ub_H;
    int nonneg_W, nonneg_H;
    int cd_maxit;
    int norm_type, solver_mode;
    int cv_patience;             // core/config.hpp:257 (default 5; not on the bridge wire)
    int threads;
    float holdout_fraction;      // core/config.hpp:236
    uint32_t cv_seed, seed;      // effective_cv_seed(): cv_seed != 0 ? cv_seed : seed (core/config.hpp:416-418)
    int mask_zeros;
} orc_cv_config;

typedef struct {
    int iterations, converged, best_iter;
    float train_loss, test_loss, best_test_loss, final_tol;
    double loop_seconds;
    long n_test;                 // held-out entries (last evaluation)
} orc_cv_result;

// nmf/cv_detail.hpp:305-340 (compute_train_rhs: operand A, column j, mask(i = inner, j)) and :357-405
// (compute_train_rhs_W: operand Aᵀ, column i, mask(i, col = inner)). Train entries accumulate into b in ascending
// inner order; held-out entries are listed (index and stored value, 0 for a structural zero) in the same order.
// mask_zeros: only the stored entries can be held out; otherwise every cell of the column is hashed.
extern "C++" {
template <class Held>
static void cv_train_rhs(const int* Cp, const int* Ci, const float* Cx, long n_inner, long col, const float* F, int k,
                         bool transposed, bool mz, const Held& held, float* b, std::vector<int>& test,
                         std::vector<float>& tval) {
    std::fill(b, b + k, 0.f);
    test.clear();
    tval.clear();
    auto is_held = [&](long inner) { return transposed ? held(col, inner) : held(inner, col); };
    if (mz) {
        for (long p = Cp[col]; p < Cp[col + 1]; ++p) {
            if (is_held(Ci[p])) { test.push_back(Ci[p]); tval.push_back(Cx[p]); }
            else { const float v = Cx[p]; const float* f = F + static_cast<long>(Ci[p]) * k;
                   for (int t = 0; t < k; ++t) b[t] += v * f[t]; }
        }
    } else {
        long p = Cp[col];
        for (long c = 0; c < n_inner; ++c) {
            float v = 0.f;
            if (p < Cp[col + 1] && Ci[p] == c) { v = Cx[p]; ++p; }
            if (is_held(c)) { test.push_back(static_cast<int>(c)); tval.push_back(v); }
            else if (v != 0.f) { const float* f = F + c * k; for (int t = 0; t < k; ++t) b[t] += v * f[t]; }
        }
    }
}
}  // extern "C++"

static void cv_solve_col(const float* G, const float* F, const std::vector<int>& test, int k, float* Gl, float* Lc,
                         float* b, float* x, float L1, bool nonneg, int cd_maxit, int solver_mode) {
    std::copy(G, G + static_cast<long>(k) * k, Gl);                              // cv_detail.hpp:74
    for (int r : test) {                                                         // :75-84 (restated, see header)
        const float* f = F + static_cast<long>(r) * k;
        for (int c = 0; c < k; ++c)
            for (int a = 0; a < k; ++a) Gl[static_cast<long>(c) * k + a] -= f[a] * f[c];
    }
    if (solver_mode == 1) {                                                      // cholesky_clip.hpp:65-106
        if (L1 > 0) for (int i = 0; i < k; ++i) b[i] -= L1;
        orc::cholesky_factor(Gl, k, Lc);
        std::copy(b, b + k, x);
        orc::cholesky_solve_inplace(Lc, k, x);
        if (nonneg) for (int i = 0; i < k; ++i) x[i] = std::max(x[i], 0.f);
    } else {                                                                     // fit_cv.hpp:469-472: no cd_tol
        orc::cd_nnls_col_fixed(Gl, b, x, k, L1, 0.f, nonneg, cd_maxit, 0.f, 0.f);
    }
}

// The two per-column building blocks of the CV path on their own (tests/test_reference_sources.py holds them against
// the reference's compute_train_rhs / compute_train_rhs_W / apply_gram_correction). Returns the number of held-out
// entries; test_idx must have room for every candidate (n_inner).
long orc_cv_train_rhs_f32(const int* Cp, const int* Ci, const float* Cx, long n_inner, long col, const float* F, int k,
                          int transposed, int mask_zeros, uint64_t seed, uint64_t inv_prob, float* b, int* test_idx) {
    auto held = [&](long i, long j) {
        return orc::SplitMix64::is_holdout(seed, static_cast<uint32_t>(i), static_cast<uint32_t>(j), inv_prob);
    };
    std::vector<int> test;
    std::vector<float> tval;
    cv_train_rhs(Cp, Ci, Cx, n_inner, col, F, k, transposed != 0, mask_zeros != 0, held, b, test, tval);
    for (size_t q = 0; q < test.size(); ++q) test_idx[q] = test[q];
    return static_cast<long>(test.size());
}
// G_local = G − Σ_{r in test} f_r f_rᵀ (cv_detail.hpp:67-85, restated as sequential fp32 downdates in test order)
void orc_cv_gram_correction_f32(const float* G, const float* F, const int* test_idx, long n_test, int k, float* Gl) {
    std::copy(G, G + static_cast<long>(k) * k, Gl);
    for (long q = 0; q < n_test; ++q) {
        const float* f = F + static_cast<long>(test_idx[q]) * k;
        for (int c = 0; c < k; ++c)
            for (int a = 0; a < k; ++a) Gl[static_cast<long>(c) * k + a] -= f[a] * f[c];
    }
}

int orc_nmf_fit_cv_f32(const int* Ap, const int* Ai, const float* Ax, long m, long n, const orc_cv_config* cfg,
                       float* W_T, float* H, float* d, float* train_hist, float* test_hist, orc_cv_result* res) {
    using namespace orc;
    const int k = cfg->k;
    if (k <= 0 || cfg->max_iter <= 0 || cfg->cd_maxit <= 0) return -1;
    const long nnz = Ap[n];
    const int threads = cfg->threads > 0 ? cfg->threads : orc_max_threads_impl();
    // speckled_cv.hpp:117-127
    const uint32_t eff = cfg->cv_seed != 0 ? cfg->cv_seed : cfg->seed;
    const uint64_t mseed = (eff == 0) ? 12345ULL : static_cast<uint64_t>(eff);
    const double hf = static_cast<double>(cfg->holdout_fraction);               // fit_cv.hpp:183
    const uint64_t inv_prob = hf > 0 ? static_cast<uint64_t>(1.0 / hf) : 0;
    auto held = [&](long i, long j) {
        return SplitMix64::is_holdout(mseed, static_cast<uint32_t>(i), static_cast<uint32_t>(j), inv_prob);
    };
    const bool mz = cfg->mask_zeros != 0;
    for (int i = 0; i < k; ++i) d[i] = 1.f;
    const float trAtA = trace_AtA(Ax, nnz);                                      // :351
    std::vector<int> Atp(m + 1), Ati(nnz);
    std::vector<float> Atx(nnz);
    transpose_csc(Ap, Ai, Ax, m, n, Atp.data(), Ati.data(), Atx.data());
    std::vector<float> G(static_cast<size_t>(k) * k), G_H_saved(G.size()), G_W_new(G.size());
    std::vector<float> B_W_full(static_cast<size_t>(k) * m);
    float prev_conv_loss = std::numeric_limits<float>::max();                   // :348
    float best_test_loss = std::numeric_limits<float>::max();                   // :354
    int best_iter = 0, patience_count = 0;
    *res = orc_cv_result{};
    const auto t0 = std::chrono::high_resolution_clock::now();

    auto normalise = [&](float* X, long cols) {                                  // :536-548 / :849-858
        if (cfg->norm_type == 2) { for (int i = 0; i < k; ++i) d[i] = 1.f; return; }
        extract_scaling(X, k, cols, d, cfg->norm_type, threads);
    };

    for (int iter = 0; iter < cfg->max_iter; ++iter) {
        // ---------------- H update (:409-476)
        gram(W_T, k, m, G.data(), threads);
        for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += 1e-15f;   // :414
        if (cfg->L2_H > 0) for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += cfg->L2_H;
#pragma omp parallel num_threads(threads)
        {
            std::vector<float> b(k), x(k), Gl(static_cast<size_t>(k) * k), Lc(Gl.size());
            std::vector<int> test;
            std::vector<float> tval;
#pragma omp for schedule(dynamic, 64)
            for (long j = 0; j < n; ++j) {
                cv_train_rhs(Ap, Ai, Ax, m, j, W_T, k, false, mz, held, b.data(), test, tval);   // cv_detail.hpp:305-340
                std::copy(H + j * k, H + (j + 1) * k, x.begin());                // :460 x_local = H.col(j)
                cv_solve_col(G.data(), W_T, test, k, Gl.data(), Lc.data(), b.data(), x.data(), cfg->L1_H,
                             cfg->nonneg_H != 0, cfg->cd_maxit, cfg->solver_mode);
                std::copy(x.begin(), x.end(), H + j * k);
            }
        }
        if (cfg->ub_H > 0) apply_upper_bound(H, static_cast<long>(k) * n, cfg->ub_H);     // :528-529
        normalise(H, n);

        // ---------------- W update (:560-735)
        gram(H, k, n, G.data(), threads);
        G_H_saved = G;                                                           // :573-576
        for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += 1e-15f;   // :578
        if (cfg->L2_W > 0) for (int i = 0; i < k; ++i) G[static_cast<long>(i) * k + i] += cfg->L2_W;
#pragma omp parallel num_threads(threads)
        {
            std::vector<float> b(k), bf(k), x(k), Gl(static_cast<size_t>(k) * k), Lc(Gl.size());
            std::vector<int> test;
            std::vector<float> tval;
#pragma omp for schedule(dynamic, 64)
            for (long i = 0; i < m; ++i) {
                cv_train_rhs(Atp.data(), Ati.data(), Atx.data(), n, i, H, k, true, mz, held, b.data(), test, tval);   // cv_detail.hpp:357-405
                bf = b;                                                          // :609-653 full RHS (train, then test entries)
                for (size_t q = 0; q < test.size(); ++q)
                    if (tval[q] != 0.f) { const float* f = H + static_cast<long>(test[q]) * k;
                                          for (int t = 0; t < k; ++t) bf[t] += tval[q] * f[t]; }
                std::copy(bf.begin(), bf.end(), B_W_full.begin() + i * k);
                std::copy(W_T + i * k, W_T + (i + 1) * k, x.begin());            // :700
                cv_solve_col(G.data(), H, test, k, Gl.data(), Lc.data(), b.data(), x.data(), cfg->L1_W,
                             cfg->nonneg_W != 0, cfg->cd_maxit, cfg->solver_mode);
                std::copy(x.begin(), x.end(), W_T + i * k);
            }
        }
        if (cfg->ub_W > 0) apply_upper_bound(W_T, static_cast<long>(k) * m, cfg->ub_W);   // :843-844
        normalise(W_T, m);

        // ---------------- loss (:1349-1548), every iteration (cv_patience > 0, track_train_loss)
        double test_sq = 0.0;
        long n_test = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 64) reduction(+ : test_sq, n_test)
        for (long j = 0; j < n; ++j) {
            auto pred_at = [&](long i) {
                double s = 0.0;
                for (int f = 0; f < k; ++f) s += static_cast<double>(W_T[i * k + f] * d[f]) * static_cast<double>(H[j * k + f]);
                return static_cast<float>(s);
            };
            if (mz) {
                for (long p = Ap[j]; p < Ap[j + 1]; ++p)
                    if (held(Ai[p], j)) { const float df = Ax[p] - pred_at(Ai[p]); test_sq += static_cast<double>(df * df); ++n_test; }
            } else {
                long p = Ap[j];
                for (long i = 0; i < m; ++i) {
                    if (held(i, j)) {
                        float a = 0.f;
                        if (p < Ap[j + 1] && Ai[p] == i) { a = Ax[p]; ++p; }
                        const float df = a - pred_at(i);
                        test_sq += static_cast<double>(df * df);
                        ++n_test;
                    } else if (p < Ap[j + 1] && Ai[p] == i) ++p;
                }
            }
        }
        const float test_sq_error = static_cast<float>(test_sq);
        double cross = 0.0;                                                      // :1511-1515
        for (long i = 0; i < m; ++i)
            for (int r = 0; r < k; ++r)
                cross += static_cast<double>(d[r] * W_T[i * k + r]) * static_cast<double>(B_W_full[i * k + r]);
        gram(W_T, k, m, G_W_new.data(), threads);                                // :1518-1519
        double recon = 0.0;
        for (int r = 0; r < k; ++r)
            for (int s2 = 0; s2 < k; ++s2)
                recon += static_cast<double>(d[r] * d[s2] * G_W_new[static_cast<long>(s2) * k + r] *
                                             G_H_saved[static_cast<long>(s2) * k + r]);
        const float total_sq = std::max(trAtA - 2.f * static_cast<float>(cross) + static_cast<float>(recon), 0.f);
        const float train_sq_error = std::max(total_sq - test_sq_error, 0.f);     // :1533
        const long total_entries = mz ? nnz : m * n;
        const long n_train = total_entries - n_test;
        const float train_loss = n_train > 0 ? train_sq_error / static_cast<float>(n_train) : 0.f;
        const float test_loss = n_test > 0 ? test_sq_error / static_cast<float>(n_test) : 0.f;
        if (train_hist) train_hist[iter] = train_loss;
        if (test_hist) test_hist[iter] = test_loss;
        res->train_loss = train_loss; res->test_loss = test_loss; res->n_test = n_test;

        float rel = 0.f;                                                         // :1565-1569
        if (iter > 0) rel = std::abs(prev_conv_loss - test_loss) / (std::abs(prev_conv_loss) + 1e-15f);
        if (test_loss < best_test_loss) { best_test_loss = test_loss; best_iter = iter; patience_count = 0; }   // :1584-1590
        else ++patience_count;
        if (cfg->cv_patience > 0 && patience_count >= cfg->cv_patience) {        // :1600-1606
            res->iterations = iter + 1; res->converged = 0; break;
        }
        if (iter > 0) {                                                          // :1609-1619
            res->final_tol = rel;
            if (rel < cfg->tol) { res->iterations = iter + 1; res->converged = 1; break; }
        }
        prev_conv_loss = test_loss;
        res->iterations = iter + 1;
    }
    res->loop_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    res->best_test_loss = best_test_loss;
    res->best_iter = best_iter;
    for (long j = 0; j < n; ++j)                                                 // :1639-1641 absorb d into H
        for (int i = 0; i < k; ++i) H[j * k + i] *= d[i];
    return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
