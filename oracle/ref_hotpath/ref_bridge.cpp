// The REFERENCE-SIDE CALLER of the GPU boundary — inst/include/FactorNet/gpu/bridge_nmf.hpp (bridge_nmf_sparse,
// bridge_nmf_cv_sparse: the code that packs the 73 / 51 pointer arguments and dlsym's the entry points) and
// gpu/loader.hpp (detect_gpus_via_bridge) — compiled unmodified from /root/reference against the Eigen stand-in.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_bridge.so): tests load RcppML_gpu.so with RTLD_GLOBAL, as R's
// dyn.load(local = FALSE) does (R/gpu_backend.R:85-88), then call through THIS library, so the drop-in claim is
// checked with the reference's own packing code rather than with this repository's ctypes twin of it.
//
//   make -C oracle ref_hotpath
#ifndef FACTORNET_HOST_DEVICE
#define FACTORNET_HOST_DEVICE
#endif
#include <Eigen/Dense>
#include <Eigen/Sparse>
namespace Eigen { template <class D> struct DenseBase; }

#include <FactorNet/gpu/bridge_nmf.hpp>

#include <cstdio>
#include <cstring>

using namespace FactorNet;
using SpF = Eigen::SparseMatrix<float, Eigen::ColMajor, int>;

struct refbridge_params {
    int k, max_iter;
    float tol;
    float L1_W, L1_H, L2_W, L2_H, ub_W, ub_H;
    int nonneg_W, nonneg_H, cd_maxit, norm_type, solver_mode;
    unsigned seed;
    // CV only
    float holdout_fraction;
    unsigned cv_seed;
    int mask_zeros;
};
struct refbridge_result {
    int iterations, converged;
    float train_loss, final_tol, test_loss, best_test_loss;
    int best_iter;
};

static NMFConfig<float> make_config(const refbridge_params& q) {
    NMFConfig<float> c;
    c.rank = q.k; c.max_iter = q.max_iter; c.tol = q.tol;
    c.W.L1 = q.L1_W; c.H.L1 = q.L1_H; c.W.L2 = q.L2_W; c.H.L2 = q.L2_H;
    c.W.upper_bound = q.ub_W; c.H.upper_bound = q.ub_H;
    c.W.nonneg = q.nonneg_W != 0; c.H.nonneg = q.nonneg_H != 0;
    c.cd_max_iter = q.cd_maxit;
    c.norm_type = q.norm_type == 0 ? NormType::L1 : q.norm_type == 1 ? NormType::L2 : NormType::None;
    c.solver_mode = q.solver_mode;
    c.seed = q.seed;
    c.verbose = false;
    return c;
}

extern "C" {

// gpu/loader.hpp:73-92 — what nmf() calls first on every fit
int refbridge_detect(int* gpu_count, double* total_mem_bytes) {
    int count = 0;
    size_t mem = 0;
    const bool ok = gpu::detect_gpus_via_bridge(count, mem);
    *gpu_count = count;
    *total_mem_bytes = static_cast<double>(mem);
    return ok ? 1 : 0;
}

// A: CSC m x n; W_init: m x k column-major (R's w); H_init: k x n column-major (R's h). Outputs in the same layouts.
// Returns 0, or -1 when the reference bridge threw (message in err).
int refbridge_nmf_sparse_f32(const int* Ap, const int* Ai, const float* Ax, int m, int n, const refbridge_params* q,
                             const float* W_init, const float* H_init, float* W_out, float* H_out, float* d_out,
                             refbridge_result* res, char* err, int err_len) {
    try {
        const SpF A(m, n, Ap, Ai, Ax);
        DenseMatrix<float> W0(m, q->k), H0(q->k, n);
        std::memcpy(W0.data(), W_init, sizeof(float) * static_cast<size_t>(m) * q->k);
        std::memcpy(H0.data(), H_init, sizeof(float) * static_cast<size_t>(n) * q->k);
        const NMFConfig<float> cfg = make_config(*q);
        const NMFResult<float> r = gpu::bridge_nmf_sparse<float, SpF>(A, cfg, &W0, &H0);
        std::memcpy(W_out, r.W.data(), sizeof(float) * static_cast<size_t>(m) * q->k);
        std::memcpy(H_out, r.H.data(), sizeof(float) * static_cast<size_t>(n) * q->k);
        for (int i = 0; i < q->k; ++i) d_out[i] = r.d(i);
        res->iterations = r.iterations; res->converged = r.converged ? 1 : 0;
        res->train_loss = r.train_loss; res->final_tol = r.final_tol;
        res->test_loss = 0; res->best_test_loss = 0; res->best_iter = 0;
        return 0;
    } catch (const std::exception& e) {
        if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "%s", e.what());
        return -1;
    }
}

// Cross-validation entry (bridge_nmf.hpp:401-...): H is always drawn by the bridge from seed + 0x9E3779B9.
int refbridge_nmf_cv_sparse_f32(const int* Ap, const int* Ai, const float* Ax, int m, int n, const refbridge_params* q,
                                const float* W_init, float* W_out, float* H_out, float* d_out, refbridge_result* res,
                                char* err, int err_len) {
    try {
        const SpF A(m, n, Ap, Ai, Ax);
        DenseMatrix<float> W0(m, q->k);
        std::memcpy(W0.data(), W_init, sizeof(float) * static_cast<size_t>(m) * q->k);
        NMFConfig<float> cfg = make_config(*q);
        cfg.holdout_fraction = q->holdout_fraction;
        cfg.cv_seed = q->cv_seed;
        cfg.mask_zeros = q->mask_zeros != 0;
        const NMFResult<float> r = gpu::bridge_nmf_cv_sparse<float, SpF>(A, cfg, &W0, nullptr);
        std::memcpy(W_out, r.W.data(), sizeof(float) * static_cast<size_t>(m) * q->k);
        std::memcpy(H_out, r.H.data(), sizeof(float) * static_cast<size_t>(n) * q->k);
        for (int i = 0; i < q->k; ++i) d_out[i] = r.d(i);
        res->iterations = r.iterations; res->converged = r.converged ? 1 : 0;
        res->train_loss = r.train_loss; res->final_tol = r.final_tol;
        res->test_loss = r.test_loss; res->best_test_loss = r.best_test_loss; res->best_iter = r.best_iter;
        return 0;
    } catch (const std::exception& e) {
        if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "%s", e.what());
        return -1;
    }
}

}  // extern "C"
