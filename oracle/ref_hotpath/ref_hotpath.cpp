// C entry points over the REFERENCE'S OWN hot-path headers, compiled from /root/reference where they lie against the
// minimal Eigen stand-in in shim/ (Eigen itself is not in this image). TEST INFRASTRUCTURE ONLY: the library is
// built into oracle/_ref/ (git-ignored) and loaded by tests/test_reference_sources.py, which holds the oracle's
// restatement (oracle/nmf_oracle.cpp) against it bit for bit.
//
//   make -C oracle ref_hotpath
//
// Reference sources compiled here (nothing is copied into this repository):
//   inst/include/FactorNet/rng/rng.hpp                      SplitMix64, uniform, hash, is_holdout, fill_uniform
//   inst/include/FactorNet/primitives/cpu/nnls_batch.hpp     cd_nnls_col_fixed, nnls_batch<CPU,float|double>
//   inst/include/FactorNet/primitives/cpu/fused_nnls.hpp     fused_rhs_nnls_sparse, fused_rhs_cholesky_sparse,
//                                                            loss_cross_term_sparse_via_At
//   inst/include/FactorNet/primitives/cpu/cholesky_clip.hpp  cholesky_clip_col
//   inst/include/FactorNet/primitives/cpu/gram.hpp           gram<CPU,float|double>
//   inst/include/FactorNet/primitives/cpu/rhs.hpp            rhs<CPU,double> (sparse) — with gram + nnls_batch: the body of c_nnls / Rcpp_predict
//   inst/include/FactorNet/primitives/primitives.hpp         trace_AtA
//   inst/include/FactorNet/core/constants.hpp                tiny_num, CD_TOL, CD_MAXIT, CD_ABS_TOL, NMF_PATIENCE
//   inst/include/FactorNet/nmf/masked_nnls.hpp               masked_nnls_h / masked_nnls_w / masked_loss (+ core/config.hpp)
//   inst/include/FactorNet/nmf/variant_helpers.hpp           extract_scaling
//   inst/include/FactorNet/features/bounds.hpp               apply_upper_bound
//   inst/include/FactorNet/nmf/speckled_cv.hpp               LazySpeckledMask (seed / inv_prob conventions)
//   inst/include/FactorNet/nmf/cv_detail.hpp                 compute_train_rhs, compute_train_rhs_W, apply_gram_correction
#ifndef FACTORNET_HOST_DEVICE
#define FACTORNET_HOST_DEVICE
#endif
#include <Eigen/Dense>
#include <Eigen/Sparse>
namespace Eigen { template <class D> struct DenseBase; }   // named by a fill_uniform overload that is never instantiated

namespace Eigen { template <class M> class SelfAdjointEigenSolver; }   // named by a projective-NMF helper, never instantiated

#include <FactorNet/rng/rng.hpp>
#include <FactorNet/primitives/cpu/gram.hpp>
#include <FactorNet/primitives/cpu/fused_nnls.hpp>
#include <FactorNet/primitives/cpu/rhs.hpp>
#include <FactorNet/features/bounds.hpp>
#include <FactorNet/nmf/masked_nnls.hpp>
#include <FactorNet/nmf/speckled_cv.hpp>
#include <FactorNet/nmf/variant_helpers.hpp>
#include <FactorNet/nmf/cv_detail.hpp>

#include <cstdint>
#include <cstring>

using namespace FactorNet;
using FactorNet::primitives::CPU;
using SpF = Eigen::SparseMatrix<float, Eigen::ColMajor, int>;

template <class S>
static DenseMatrix<S> dense_from(const S* p, long rows, long cols) {
    DenseMatrix<S> M(rows, cols);
    std::memcpy(M.data(), p, sizeof(S) * static_cast<size_t>(rows * cols));
    return M;
}

extern "C" {

// ---- rng/rng.hpp
void ref_splitmix_next(uint64_t seed, int count, uint64_t* out) {
    rng::SplitMix64 g(seed);
    for (int i = 0; i < count; ++i) out[i] = g.next();
}
uint64_t ref_splitmix_hash(uint64_t seed, uint32_t i, uint32_t j) { return rng::SplitMix64::hash(seed, i, j); }
int ref_is_holdout(uint64_t seed, uint32_t i, uint32_t j, uint64_t inv_prob) { return rng::SplitMix64::is_holdout(seed, i, j, inv_prob) ? 1 : 0; }
void ref_fill_uniform_f32(uint64_t seed, float* data, int rows, int cols) { rng::SplitMix64 g(seed); g.fill_uniform(data, rows, cols); }
void ref_fill_uniform_f64(uint64_t seed, double* data, int rows, int cols) { rng::SplitMix64 g(seed); g.fill_uniform(data, rows, cols); }
// nmf/nmf_init.hpp:173-181 draws W_T (k x m) then H (k x n) from ONE stream — two fill_uniform calls on one generator
void ref_init_factors_f32(uint64_t seed, int k, int m, int n, float* W_T, float* H) {
    rng::SplitMix64 g(seed);
    g.fill_uniform(W_T, k, m);
    g.fill_uniform(H, k, n);
}

// ---- core/constants.hpp
void ref_constants(double* out) {
    out[0] = static_cast<double>(tiny_num<float>());
    out[1] = tiny_num<double>();
    out[2] = CD_TOL;
    out[3] = static_cast<double>(CD_MAXIT);
    out[4] = CD_ABS_TOL;
    out[5] = static_cast<double>(NMF_PATIENCE);
}

// ---- primitives/cpu/nnls_batch.hpp
int ref_cd_nnls_col_fixed_f32(const float* G, float* b, float* x, int k, float L1, float L2, int nonneg, int maxit,
                              float ub, float cd_tol) {
    const DenseMatrix<float> Gm = dense_from(G, k, k);
    return primitives::detail::cd_nnls_col_fixed<float>(Gm, b, x, k, L1, L2, nonneg != 0, maxit, ub, cd_tol);
}
int ref_cd_nnls_col_fixed_f64(const double* G, double* b, double* x, int k, double L1, double L2, int nonneg, int maxit,
                              double ub, double cd_tol) {
    const DenseMatrix<double> Gm = dense_from(G, k, k);
    return primitives::detail::cd_nnls_col_fixed<double>(Gm, b, x, k, L1, L2, nonneg != 0, maxit, ub, cd_tol);
}
void ref_nnls_batch_f64(const double* G, double* B, double* X, int k, long n, int cd_maxit, double cd_tol, double L1,
                        double L2, int nonneg, double ub, int warm_start) {
    const DenseMatrix<double> Gm = dense_from(G, k, k);
    DenseMatrix<double> Bm = dense_from(B, k, n), Xm = dense_from(X, k, n);
    primitives::nnls_batch<CPU, double>(Gm, Bm, Xm, cd_maxit, cd_tol, L1, L2, nonneg != 0, 1, ub, warm_start != 0);
    std::memcpy(B, Bm.data(), sizeof(double) * static_cast<size_t>(k) * n);
    std::memcpy(X, Xm.data(), sizeof(double) * static_cast<size_t>(k) * n);
}

// ---- c_nnls / Rcpp_predict (src/RcppFunctions_utils.cpp:314-366, 23-53): those two functions need Rcpp, but their
// bodies are these calls into the reference's own primitives, in this order, in double. w_T: k x m, h: k x n in/out.
void ref_c_nnls_sparse_f64(const int* Ap, const int* Ai, const double* Ax, long m, long n, const double* w_T, int k, double* h,
                           int cd_maxit, double cd_tol, double L1, double L2, double ub, int nonneg, int warm_start) {
    using SpD = Eigen::SparseMatrix<double, Eigen::ColMajor, int>;
    const SpD A(m, n, Ap, Ai, Ax);
    const DenseMatrix<double> Wm = dense_from(w_T, k, m);
    DenseMatrix<double> G(k, k);
    primitives::gram<CPU, double>(Wm, G);                                             // :326 / :33
    for (int i = 0; i < k; ++i) G(i, i) += tiny_num<double>();                          // :327
    if (L2 > 0) for (int i = 0; i < k; ++i) G(i, i) += L2;                              // :328
    DenseMatrix<double> B;
    B.resize(k, n);
    primitives::rhs<CPU, double>(A, Wm, B);                                           // :338
    DenseMatrix<double> hm = dense_from(h, k, n);
    if (warm_start) {                                                                   // :347-356
        B.noalias() -= G * hm;
        for (long j = 0; j < n; ++j)
            primitives::detail::cd_nnls_col_fixed(G, B.col(j).data(), hm.col(j).data(), k, L1, 0.0, nonneg != 0, cd_maxit, ub);
    } else {                                                                            // :358-360
        primitives::nnls_batch<CPU, double>(G, B, hm, cd_maxit, cd_tol, L1, 0.0, nonneg != 0, 1, ub);
    }
    std::memcpy(h, hm.data(), sizeof(double) * static_cast<size_t>(k) * n);
}

// ---- primitives/cpu/gram.hpp (F is k x ncols col-major)
void ref_gram_f32(const float* F, int k, long ncols, float* G) {
    const DenseMatrix<float> Fm = dense_from(F, k, ncols);
    DenseMatrix<float> Gm;
    primitives::gram<CPU, float>(Fm, Gm);
    std::memcpy(G, Gm.data(), sizeof(float) * static_cast<size_t>(k) * k);
}

// ---- primitives/cpu/fused_nnls.hpp (A: CSC m x n; Factor k x m; X k x n, in/out)
void ref_fused_rhs_nnls_sparse_f32(const int* Ap, const int* Ai, const float* Ax, long m, long n, const float* Factor,
                                   const float* G, float* X, int k, int cd_maxit, float cd_tol, float L1, int nonneg,
                                   int warm_start, float ub) {
    const SpF A(m, n, Ap, Ai, Ax);
    const DenseMatrix<float> Fm = dense_from(Factor, k, m), Gm = dense_from(G, k, k);
    DenseMatrix<float> Xm = dense_from(X, k, n);
    primitives::fused_rhs_nnls_sparse<SpF, float>(A, Fm, Gm, Xm, cd_maxit, cd_tol, L1, nonneg != 0, 1, warm_start != 0, ub);
    std::memcpy(X, Xm.data(), sizeof(float) * static_cast<size_t>(k) * n);
}
void ref_fused_rhs_cholesky_sparse_f32(const int* Ap, const int* Ai, const float* Ax, long m, long n, const float* Factor,
                                       const float* G, float* X, int k, int solver_mode, float L1, int nonneg, float ub) {
    const SpF A(m, n, Ap, Ai, Ax);
    const DenseMatrix<float> Fm = dense_from(Factor, k, m), Gm = dense_from(G, k, k);
    DenseMatrix<float> Xm = dense_from(X, k, n);
    primitives::fused_rhs_cholesky_sparse<SpF, float>(A, Fm, Gm, Xm, solver_mode, L1, nonneg != 0, 1, false, ub);
    std::memcpy(X, Xm.data(), sizeof(float) * static_cast<size_t>(k) * n);
}
// At: CSC n x m (the transpose of A); W_T k x m; H k x n
float ref_loss_cross_term_via_At_f32(const int* Atp, const int* Ati, const float* Atx, long n, long m, const float* W_T,
                                     const float* H, const float* d, int k) {
    const SpF At(n, m, Atp, Ati, Atx);
    const DenseMatrix<float> Wm = dense_from(W_T, k, m), Hm = dense_from(H, k, n);
    DenseVector<float> dv(k);
    for (int i = 0; i < k; ++i) dv(i) = d[i];
    return primitives::loss_cross_term_sparse_via_At<SpF, float>(At, Wm, Hm, dv, 1);
}
float ref_trace_AtA_f32(const int* Ap, const int* Ai, const float* Ax, long m, long n) {
    const SpF A(m, n, Ap, Ai, Ax);
    return primitives::trace_AtA<CPU, float>(A);
}

// ---- primitives/cpu/cholesky_clip.hpp
void ref_cholesky_clip_col_f32(const float* G, float* b, float* x, int k, float L1, float L2, int nonneg, float ub) {
    const DenseMatrix<float> Gm = dense_from(G, k, k);
    primitives::detail::cholesky_clip_col<float>(Gm, b, x, k, L1, L2, nonneg != 0, 0, 0.f, ub);
}

// ---- nmf/masked_nnls.hpp. A: CSC m x n; mask: CSC pattern m x n of the masked entries (values 1); F = W_T (k x m);
//      X = H (k x n). The W half-step is the same call on the transposes (masked_nnls_w over At, mask_T).
static NMFConfig<float> masked_cfg(float L1, float L2, int nonneg, int cd_maxit, float cd_tol, int solver_mode, bool for_w) {
    NMFConfig<float> c;
    c.cd_max_iter = cd_maxit; c.cd_tol = cd_tol; c.solver_mode = solver_mode;
    auto& f = for_w ? c.W : c.H;
    f.L1 = L1; f.L2 = L2; f.nonneg = nonneg != 0;
    return c;
}
void ref_masked_nnls_h_f32(const int* Ap, const int* Ai, const float* Ax, long m, long n, const float* W_T, const float* G,
                           float* H, int k, const int* Mp, const int* Mi, float L1, float L2, int nonneg, int cd_maxit,
                           float cd_tol, int solver_mode, int warm_start) {
    const SpF A(m, n, Ap, Ai, Ax);
    std::vector<float> ones(static_cast<size_t>(Mp[n]) + 1, 1.f);
    const SpF M(m, n, Mp, Mi, ones.data());
    const DenseMatrix<float> Wm = dense_from(W_T, k, m), Gm = dense_from(G, k, k);
    DenseMatrix<float> Hm = dense_from(H, k, n);
    nmf::mask_detail::masked_nnls_h<float, SpF>(A, Wm, Gm, Hm, M, masked_cfg(L1, L2, nonneg, cd_maxit, cd_tol, solver_mode, false), 1, warm_start != 0);
    std::memcpy(H, Hm.data(), sizeof(float) * static_cast<size_t>(k) * n);
}
void ref_masked_nnls_w_f32(const int* Atp, const int* Ati, const float* Atx, long n, long m, const float* H, const float* G,
                           float* W_T, int k, const int* MTp, const int* MTi, float L1, float L2, int nonneg, int cd_maxit,
                           float cd_tol, int solver_mode, int warm_start) {
    const SpF At(n, m, Atp, Ati, Atx);
    std::vector<float> ones(static_cast<size_t>(MTp[m]) + 1, 1.f);
    const SpF MT(n, m, MTp, MTi, ones.data());
    const DenseMatrix<float> Hm = dense_from(H, k, n), Gm = dense_from(G, k, k);
    DenseMatrix<float> Wm = dense_from(W_T, k, m);
    nmf::mask_detail::masked_nnls_w<float, SpF>(At, Hm, Gm, Wm, MT, masked_cfg(L1, L2, nonneg, cd_maxit, cd_tol, solver_mode, true), 1, warm_start != 0);
    std::memcpy(W_T, Wm.data(), sizeof(float) * static_cast<size_t>(k) * m);
}
float ref_masked_loss_f32(const int* Ap, const int* Ai, const float* Ax, long m, long n, const float* W_Td, const float* H,
                          int k, const int* Mp, const int* Mi) {
    const SpF A(m, n, Ap, Ai, Ax);
    std::vector<float> ones(static_cast<size_t>(Mp[n]) + 1, 1.f);
    const SpF M(m, n, Mp, Mi, ones.data());
    const DenseMatrix<float> Wm = dense_from(W_Td, k, m), Hm = dense_from(H, k, n);
    LossConfig<float> lc{};
    return nmf::mask_detail::masked_loss<float, SpF>(A, Wm, Hm, M, lc, 1);
}

// ---- nmf/variant_helpers.hpp:287-305 and features/bounds.hpp:38 (X is k x n col-major, in place)
void ref_extract_scaling_f32(float* X, int k, long n, float* d, int norm_type) {
    DenseMatrix<float> Xm = dense_from(X, k, n);
    DenseVector<float> dv(k);
    nmf::variant::extract_scaling<float>(Xm, dv, norm_type == 0 ? NormType::L1 : norm_type == 1 ? NormType::L2 : NormType::None);
    std::memcpy(X, Xm.data(), sizeof(float) * static_cast<size_t>(k) * n);
    for (int i = 0; i < k; ++i) d[i] = dv(i);
}
void ref_apply_upper_bound_f32(float* X, int k, long n, float ub) {
    DenseMatrix<float> Xm = dense_from(X, k, n);
    features::apply_upper_bound<float>(Xm, ub);
    std::memcpy(X, Xm.data(), sizeof(float) * static_cast<size_t>(k) * n);
}

// ---- nmf/speckled_cv.hpp: LazySpeckledMask(n_rows, n_cols, nnz, holdout_fraction, seed, mask_zeros).is_holdout(i, j)
void ref_speckled_mask_f32(int n_rows, int n_cols, double holdout_fraction, uint64_t seed, const int* ii, const int* jj,
                           int count, int* out) {
    nmf::LazySpeckledMask<float> mask(n_rows, n_cols, 0, holdout_fraction, seed, true);
    for (int t = 0; t < count; ++t) out[t] = mask.is_holdout(ii[t], jj[t]) ? 1 : 0;
}

// ---- nmf/cv_detail.hpp. C = A (CSC m x n) with F = W_T for the H side (transposed = 0), C = Aᵀ (CSC n x m) with F = H
//      for the W side (transposed = 1). Returns the number of held-out entries of column `col`, listed in test_idx.
long ref_cv_train_rhs_f32(const int* Cp, const int* Ci, const float* Cx, long n_inner, long n_cols, long col, const float* F,
                          int k, int transposed, int mask_zeros, double holdout_fraction, uint64_t seed, float* b, int* test_idx) {
    const SpF Cm(n_inner, n_cols, Cp, Ci, Cx);
    const DenseMatrix<float> Fm = dense_from(F, k, n_inner);
    // LazySpeckledMask(n_rows, n_cols, ...) is indexed (row of A, column of A) on both sides
    const nmf::LazySpeckledMask<float> mask(transposed ? static_cast<int>(n_cols) : static_cast<int>(n_inner),
                                            transposed ? static_cast<int>(n_inner) : static_cast<int>(n_cols), 0,
                                            holdout_fraction, seed, mask_zeros != 0);
    DenseVector<float> bv(k);
    std::vector<int> test;
    if (transposed) nmf::detail::compute_train_rhs_W<float, SpF>(Cm, Fm, static_cast<int>(col), mask, bv, test, mask_zeros != 0, static_cast<int>(n_inner));
    else nmf::detail::compute_train_rhs<float, SpF>(Cm, Fm, static_cast<int>(col), mask, bv, test, mask_zeros != 0, static_cast<int>(n_inner));
    for (int i = 0; i < k; ++i) b[i] = bv(i);
    for (size_t q = 0; q < test.size(); ++q) test_idx[q] = test[q];
    return static_cast<long>(test.size());
}
void ref_cv_gram_correction_f32(const float* G, const float* F, long n_rows, const int* test_idx, long n_test, int k, float* Gl) {
    const DenseMatrix<float> Gm = dense_from(G, k, k), Fm = dense_from(F, k, n_rows);
    DenseMatrix<float> Gl_m(k, k), buf(k, n_test > 0 ? n_test : 1);
    const std::vector<int> test(test_idx, test_idx + n_test);
    nmf::detail::apply_gram_correction<float>(Gm, Fm, test, Gl_m, buf);
    std::memcpy(Gl, Gl_m.data(), sizeof(float) * static_cast<size_t>(k) * k);
}

}  // extern "C"
