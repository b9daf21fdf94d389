// The REFERENCE'S OWN ALS LOOPS — inst/include/FactorNet/nmf/fit_cpu.hpp, nmf_fit<CPU, float, SparseMatrix<float>>,
// and nmf/fit_cv.hpp, nmf_fit_cv<CPU, float, SparseMatrix<float>> (speckled-mask cross-validation) —
// compiled unmodified from /root/reference against the Eigen stand-in (shim/) and run as an oracle of the oracle.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_fit.so; tests/test_reference_fit.py).
//
// What is the reference's own source here: the whole orchestration of a fit (initialisation from W_init/H_init, the
// transpose, which Gram feeds which half-step, where L2 / L1 / the bounds / the scaling sit, the fused-path
// predicate and its first-iteration quirk, the explicit-mask branch, the loss by the Gram trick, the patience
// logic, the packing of the result) together with every header already pinned piecewise (tests/
// test_reference_sources.py). What is NOT: (1) the arithmetic inside Eigen, implemented in shim/ with the oracle's
// definitions; (2) the features outside the hot path — SVD initialisation, IRLS losses, graph / L21 /
// angular regularisers — whose headers are shadowed by out_of_path/ declarations that throw if ever called
// (fit_cpu.hpp instantiates them behind run-time switches; an MSE fit with given factors never takes them).
//
//   make -C oracle ref_hotpath
#ifndef FACTORNET_HOST_DEVICE
#define FACTORNET_HOST_DEVICE
#endif
#include <cstdio>
#define Rprintf(...) std::printf(__VA_ARGS__)          // R's printf (only reached with config.verbose)
#include <Eigen/Dense>
#include <Eigen/Sparse>
namespace Eigen { template <class D> struct DenseBase; }

#include <FactorNet/nmf/fit_cpu.hpp>
#include <FactorNet/nmf/fit_cv.hpp>

#include <cstring>

using namespace FactorNet;
using SpF = Eigen::SparseMatrix<float, Eigen::ColMajor, int>;

struct reffit_params {
    int k, max_iter;
    float tol;
    float L1_W, L1_H, L2_W, L2_H, ub_W, ub_H;
    int nonneg_W, nonneg_H, cd_maxit;
    float cd_tol;
    int norm_type, solver_mode, patience, threads, sort_model;
    unsigned seed;
};
struct reffit_result {
    int iterations, converged;
    float train_loss, final_tol;
    int n_loss;
};

extern "C" {

// A: CSC m x n. W_init: m x k column-major; H_init: k x n column-major (nullptr: the reference draws H itself,
// fit_cpu.hpp:203-206). mask: CSC pattern m x n of the masked entries or nullptr. Outputs: W (m x k column-major),
// H (k x n column-major), d (k), loss history (up to max_iter values). Returns 0, or -1 when nmf_fit threw.
int reffit_nmf_sparse_f32(const int* Ap, const int* Ai, const float* Ax, int m, int n, const reffit_params* q,
                          const float* W_init, const float* H_init, const int* Mp, const int* Mi, float* W_out,
                          float* H_out, float* d_out, float* loss_hist, reffit_result* res, char* err, int err_len) {
    try {
        const SpF A(m, n, Ap, Ai, Ax);
        NMFConfig<float> c;
        c.rank = q->k; c.max_iter = q->max_iter; c.tol = q->tol; c.patience = q->patience; c.seed = q->seed;
        c.threads = q->threads; c.verbose = false;
        c.W.L1 = q->L1_W; c.H.L1 = q->L1_H; c.W.L2 = q->L2_W; c.H.L2 = q->L2_H;
        c.W.upper_bound = q->ub_W; c.H.upper_bound = q->ub_H;
        c.W.nonneg = q->nonneg_W != 0; c.H.nonneg = q->nonneg_H != 0;
        c.cd_max_iter = q->cd_maxit; c.cd_tol = q->cd_tol;
        c.norm_type = q->norm_type == 0 ? NormType::L1 : q->norm_type == 1 ? NormType::L2 : NormType::None;
        c.solver_mode = q->solver_mode;
        c.sort_model = q->sort_model != 0;
        c.track_loss_history = true;
        c.loss_every = 1;
        std::vector<float> ones;
        SpF M;
        if (Mp && Mi) {
            ones.assign(static_cast<size_t>(Mp[n]) + 1, 1.f);
            M = SpF(m, n, Mp, Mi, ones.data());
            c.mask = &M;
        }
        DenseMatrix<float> W0(m, q->k), H0(q->k, n);
        std::memcpy(W0.data(), W_init, sizeof(float) * static_cast<size_t>(m) * q->k);
        if (H_init) std::memcpy(H0.data(), H_init, sizeof(float) * static_cast<size_t>(n) * q->k);
        const NMFResult<float> r = nmf::nmf_fit<primitives::CPU, float, SpF>(A, c, &W0, H_init ? &H0 : nullptr);
        std::memcpy(W_out, r.W.data(), sizeof(float) * static_cast<size_t>(m) * q->k);
        std::memcpy(H_out, r.H.data(), sizeof(float) * static_cast<size_t>(n) * q->k);
        for (int i = 0; i < q->k; ++i) d_out[i] = r.d(i);
        res->iterations = r.iterations; res->converged = r.converged ? 1 : 0;
        res->train_loss = r.train_loss; res->final_tol = r.final_tol;
        res->n_loss = static_cast<int>(r.loss_history.size());
        for (int i = 0; i < res->n_loss && i < q->max_iter; ++i) loss_hist[i] = r.loss_history[static_cast<size_t>(i)];
        return 0;
    } catch (const std::exception& e) {
        if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "%s", e.what());
        return -1;
    }
}

struct reffit_cv_params {
    float holdout_fraction;
    unsigned cv_seed;
    int mask_zeros, cv_patience;
};
struct reffit_cv_result {
    float test_loss, best_test_loss;
    int best_iter, n_test_hist;
};

// nmf_fit_cv (nmf/fit_cv.hpp:124). Same layouts as above; train / test loss histories up to max_iter values each.
int reffit_nmf_cv_sparse_f32(const int* Ap, const int* Ai, const float* Ax, int m, int n, const reffit_params* q,
                             const reffit_cv_params* cvq, const float* W_init, const float* H_init, float* W_out,
                             float* H_out, float* d_out, float* train_hist, float* test_hist, reffit_result* res,
                             reffit_cv_result* cvres, char* err, int err_len) {
    try {
        const SpF A(m, n, Ap, Ai, Ax);
        NMFConfig<float> c;
        c.rank = q->k; c.max_iter = q->max_iter; c.tol = q->tol; c.patience = q->patience; c.seed = q->seed;
        c.threads = q->threads; c.verbose = false;
        c.W.L1 = q->L1_W; c.H.L1 = q->L1_H; c.W.L2 = q->L2_W; c.H.L2 = q->L2_H;
        c.W.upper_bound = q->ub_W; c.H.upper_bound = q->ub_H;
        c.W.nonneg = q->nonneg_W != 0; c.H.nonneg = q->nonneg_H != 0;
        c.cd_max_iter = q->cd_maxit; c.cd_tol = q->cd_tol;
        c.norm_type = q->norm_type == 0 ? NormType::L1 : q->norm_type == 1 ? NormType::L2 : NormType::None;
        c.solver_mode = q->solver_mode;
        c.sort_model = q->sort_model != 0;
        c.track_loss_history = true;
        c.loss_every = 1;
        c.holdout_fraction = cvq->holdout_fraction; c.cv_seed = cvq->cv_seed; c.mask_zeros = cvq->mask_zeros != 0;
        c.cv_patience = cvq->cv_patience;
        DenseMatrix<float> W0(m, q->k), H0(q->k, n);
        std::memcpy(W0.data(), W_init, sizeof(float) * static_cast<size_t>(m) * q->k);
        std::memcpy(H0.data(), H_init, sizeof(float) * static_cast<size_t>(n) * q->k);
        const NMFResult<float> r = nmf::nmf_fit_cv<primitives::CPU, float, SpF>(A, c, &W0, &H0);
        std::memcpy(W_out, r.W.data(), sizeof(float) * static_cast<size_t>(m) * q->k);
        std::memcpy(H_out, r.H.data(), sizeof(float) * static_cast<size_t>(n) * q->k);
        for (int i = 0; i < q->k; ++i) d_out[i] = r.d(i);
        res->iterations = r.iterations; res->converged = r.converged ? 1 : 0;
        res->train_loss = r.train_loss; res->final_tol = r.final_tol;
        res->n_loss = static_cast<int>(r.loss_history.size());
        for (int i = 0; i < res->n_loss && i < q->max_iter; ++i) train_hist[i] = r.loss_history[static_cast<size_t>(i)];
        cvres->test_loss = r.test_loss; cvres->best_test_loss = r.best_test_loss; cvres->best_iter = r.best_iter;
        cvres->n_test_hist = static_cast<int>(r.test_loss_history.size());
        for (int i = 0; i < cvres->n_test_hist && i < q->max_iter; ++i) test_hist[i] = r.test_loss_history[static_cast<size_t>(i)];
        return 0;
    } catch (const std::exception& e) {
        if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "%s", e.what());
        return -1;
    }
}

}  // extern "C"
