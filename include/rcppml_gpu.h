/*
 * rcppml_gpu.h — C ABI of the B200-native sparse-NMF ALS engine (RcppML_gpu.so).
 *
 * Part 1 are the symbols the reference's CPU-side bridge dlsym()s / .C()s out of
 * RcppML_gpu.so (citations are into the reference tree). A maintainer switches
 * backends by dropping this library where R/gpu_backend.R:149-171 looks for it.
 * Part 2 are NEW symbols (prefix rcppml_b200_) for callers that keep data resident
 * on the device (bench, multi-GPU, masked fits); they are ABI extensions and are
 * listed as such in INTEGRATION.md.
 * Part 3 is the on-disk ingest: the reference's rcppml_sp_read_gpu / rcppml_sp_free_gpu (a StreamPress v2 .spz file
 * decoded and left on the device) and the reader behind them as rcppml_b200_spz_* (host) / rcppml_b200_set_matrix_spz.
 *
 * All symbols: plain C, pointers + sizes only, never throw, never call into R.
 */
#ifndef RCPPML_GPU_H
#define RCPPML_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ========================================================================
 * Part 1 — reference ABI
 * ===================================================================== */

/* Replaces src/gpu_bridge_cluster.cu:24-47 (called from R/gpu_backend.R:101-106 via .C and
 * from inst/include/FactorNet/gpu/loader.hpp:73-92 via dlsym on every nmf()).
 * out_status = 0 and num_gpus > 0 mean "GPU plan available". */
void rcppml_gpu_detect(int* num_gpus, double* total_mem_mb, double* free_mem_mb, int* max_gpus,
                       int* out_status);

/* Replaces src/gpu_bridge_nmf.cu:460-624; function-pointer type at
 * inst/include/FactorNet/gpu/bridge_nmf.hpp:39-75; called at bridge_nmf.hpp:310-342.
 * CSC of A (m x n): col_ptr[n+1], row_idx[nnz], values[nnz] (double on the wire, fp32 inside).
 * W is k x m column-major (= W_T), H is k x n column-major, d is k; all in/out.
 * Supported: loss_type 0 (MSE), solver_mode 0 (CD) / !=0 (Cholesky+clip), L1/L2, upper
 * bounds, nonneg flags, norm_type 0/1/2. Anything else (L21, ortho, graph, guides,
 * projective, symmetric, non-MSE loss, k > 128) sets *out_status = -1 so that the
 * reference gateway (nmf/fit.hpp:125-133) takes its CPU path; there is no CPU fallback here.
 * Environment: RCPPML_NUM_GPUS = G | "all" (default 1) runs the fit sharded over G devices inside this
 * one call (one engine + host thread per device, NVLink peer memory); the signature carries no device
 * count (core/config.hpp:86 `max_gpus` is dead). Same results, bit for bit. */
void rcppml_gpu_nmf_unified_float(
    const int* col_ptr, const int* row_idx, const double* values,
    int* m, int* n, int* nnz, int* k,
    double* W, double* H, double* d,
    int* max_iter, double* tol,
    double* L1_H, double* L1_W, double* L2_H, double* L2_W,
    double* L21_H, double* L21_W,
    double* ortho_H, double* ortho_W,
    double* ub_H, double* ub_W,
    int* cd_maxit, int* verbose, int* seed,
    int* loss_every, int* patience,
    int* nonneg_W, int* nonneg_H,
    int* loss_type, double* huber_delta,
    int* irls_max_iter, double* irls_tol,
    int* norm_type,
    int* projective, int* symmetric,
    int* solver_mode,
    const int* graph_W_p, const int* graph_W_i, const double* graph_W_x,
    int* graph_W_dim, int* graph_W_nnz, double* graph_W_lambda,
    const int* graph_H_p, const int* graph_H_i, const double* graph_H_x,
    int* graph_H_dim, int* graph_H_nnz, double* graph_H_lambda,
    int* gp_dispersion_mode,
    double* gp_theta_init, double* gp_theta_max, double* gp_theta_min,
    double* nb_size_init, double* nb_size_max, double* nb_size_min,
    double* gamma_phi_init, double* gamma_phi_max, double* gamma_phi_min,
    double* robust_delta, double* tweedie_power,
    double* out_theta, int* out_theta_len,
    const int* guide_H_labels_flat, const int* guide_H_ns,
    const double* guide_H_lambdas, const int* guide_H_ncs, int* guide_H_count,
    int* out_iter, int* out_converged, double* out_loss,
    int* out_status,
    double* out_tol);

/* Replaces src/gpu_bridge_nmf.cu:879-967 (39 pointers), called through R's .C by .gpu_nmf_zerocopy
 * (R/gpu_backend.R:183-265). col_ptr / row_idx (int32) and values (double) are DEVICE arrays (sp_read_gpu); their
 * addresses are passed encoded as doubles. W (k x m), H (k x n), d are host doubles, in/out. No solver_mode on
 * this wire: Cholesky+clip (the reference's struct default, core/config.hpp:133). The reference computes this
 * entry in fp64; this engine runs its fp32 ALS loop (values converted on the device). */
void rcppml_gpu_nmf_zerocopy_double(
    double* d_col_ptr_addr, double* d_row_idx_addr, double* d_values_addr,
    int* m, int* n, double* nnz_d, int* k,
    double* W, double* H, double* d,
    int* max_iter, double* tol,
    double* L1_H, double* L1_W, double* L2_H, double* L2_W,
    double* L21_H, double* L21_W,
    double* ortho_H, double* ortho_W,
    double* ub_H, double* ub_W,
    int* cd_maxit, int* verbose, int* seed,
    int* loss_every, int* patience,
    int* nonneg_W, int* nonneg_H,
    int* loss_type, double* huber_delta,
    int* irls_max_iter, double* irls_tol,
    int* norm_type,
    int* out_iter, int* out_converged, double* out_loss,
    int* out_status,
    double* out_tol);

/* Replaces src/gpu_bridge_nmf.cu:340-455; function-pointer type gpu/bridge_nmf.hpp:78-99 (51 pointers), called by
 * bridge_nmf_cv_sparse (gpu/bridge_nmf.hpp:399+). Speckled-mask cross-validation NMF; cv_patience is not on the
 * wire (default 5). Unsupported (status -1): non-MSE loss, graphs, projective, symmetric, k > 128. */
void rcppml_gpu_nmf_cv_unified_float(
    const int* col_ptr, const int* row_idx, const double* values,
    int* m, int* n, int* nnz, int* k,
    double* W, double* H, double* d,
    int* max_iter, double* tol,
    double* L1_H, double* L1_W, double* L2_H, double* L2_W,
    int* cd_maxit, int* verbose, int* seed,
    double* holdout_frac, int* cv_seed, int* mask_zeros,
    int* nonneg_W, int* nonneg_H,
    int* norm_type,
    int* loss_type, double* huber_delta,
    int* irls_max_iter, double* irls_tol,
    const int* graph_W_p, const int* graph_W_i, const double* graph_W_x,
    int* graph_W_dim, int* graph_W_nnz, double* graph_W_lambda,
    const int* graph_H_p, const int* graph_H_i, const double* graph_H_x,
    int* graph_H_dim, int* graph_H_nnz, double* graph_H_lambda,
    int* projective, int* symmetric, int* solver_mode,
    int* out_iter, int* out_converged,
    double* out_train_loss, double* out_test_loss,
    double* out_best_test, int* out_best_iter,
    int* out_status);

/* ABI EXTENSION of Part 1: the same call with the explicit user mask, which the reference bridge does not carry
 * (gpu/bridge_nmf.hpp:39-75). mask_p[n+1] / mask_i[*mask_nnz]: CSC pattern of the masked entries of A
 * (nmf/masked_nnls.hpp:97-282; the fit then follows fit_cpu.hpp:560-564, :799-810, :1686-1691).
 * The 73 arguments of rcppml_gpu_nmf_unified_float follow the three mask arguments. */
void rcppml_gpu_nmf_masked_unified_float(
    const int* mask_p, const int* mask_i, int* mask_nnz,
    const int* col_ptr, const int* row_idx, const double* values,
    int* m, int* n, int* nnz, int* k,
    double* W, double* H, double* d,
    int* max_iter, double* tol,
    double* L1_H, double* L1_W, double* L2_H, double* L2_W,
    double* L21_H, double* L21_W,
    double* ortho_H, double* ortho_W,
    double* ub_H, double* ub_W,
    int* cd_maxit, int* verbose, int* seed,
    int* loss_every, int* patience,
    int* nonneg_W, int* nonneg_H,
    int* loss_type, double* huber_delta,
    int* irls_max_iter, double* irls_tol,
    int* norm_type,
    int* projective, int* symmetric,
    int* solver_mode,
    const int* graph_W_p, const int* graph_W_i, const double* graph_W_x,
    int* graph_W_dim, int* graph_W_nnz, double* graph_W_lambda,
    const int* graph_H_p, const int* graph_H_i, const double* graph_H_x,
    int* graph_H_dim, int* graph_H_nnz, double* graph_H_lambda,
    int* gp_dispersion_mode,
    double* gp_theta_init, double* gp_theta_max, double* gp_theta_min,
    double* nb_size_init, double* nb_size_max, double* nb_size_min,
    double* gamma_phi_init, double* gamma_phi_max, double* gamma_phi_min,
    double* robust_delta, double* tweedie_power,
    double* out_theta, int* out_theta_len,
    const int* guide_H_labels_flat, const int* guide_H_ns,
    const double* guide_H_lambdas, const int* guide_H_ncs, int* guide_H_count,
    int* out_iter, int* out_converged, double* out_loss,
    int* out_status,
    double* out_tol);

/* ABI EXTENSIONS for predict() / nnls() / evaluate() — the reference runs these on the CPU in fp64 and has no GPU
 * entry for them. Same .C conventions: every scalar behind a pointer, status out.
 *
 * rcppml_gpu_nnls_double: Rcpp_predict (src/RcppFunctions_utils.cpp:23-53) and c_nnls (:314-366) in fp64 on the GPU.
 *   w_T: k x m column-major fixed factor; h: k x n column-major output (and warm start when *warm_start != 0).
 *   predict(): nonneg=1, cd_maxit=100, cd_tol=1e-8, warm_start=0.  nnls(): as passed by R/solve.R:311,339.
 * rcppml_gpu_evaluate_double: MSE of A ~ W diag(d) H (compute_mse, :60-90); mask_zeros != 0 -> over non-zeros only. */
void rcppml_gpu_nnls_double(const int* col_ptr, const int* row_idx, const double* values, int* m, int* n, int* nnz,
                            int* k, const double* w_T, double* h, double* L1, double* L2, double* upper_bound,
                            int* nonneg, int* cd_maxit, double* cd_tol, int* warm_start, int* out_status);
void rcppml_gpu_evaluate_double(const int* col_ptr, const int* row_idx, const double* values, int* m, int* n, int* nnz,
                                int* k, const double* w_T, const double* d, const double* h, int* mask_zeros,
                                double* out_loss, int* out_status);

/* ========================================================================
 * Part 2 — device-resident engine (ABI extension)
 * ===================================================================== */

typedef struct rcppml_b200_engine rcppml_b200_engine;

/* Mirrors the fields of NMFConfig<float> (core/config.hpp:54-454) that exist on this path.
 * The W/H pairs follow src/RcppFunctions_nmf.cpp:59-72. */
typedef struct {
    int   k;
    int   max_iter;      /* used by rcppml_b200_fit only */
    float tol;
    float L1_W, L1_H, L2_W, L2_H, ub_W, ub_H;
    int   nonneg_W, nonneg_H;
    int   cd_maxit;      /* <=0 -> 10  (src/RcppFunctions_nmf.cpp:75) */
    float cd_tol;        /* <=0 -> 1e-8 (src/RcppFunctions_nmf.cpp:76; core/constants.hpp:64) */
    int   norm_type;     /* 0=L1 1=L2 2=None */
    int   solver_mode;   /* 0=CD, else Cholesky+clip (fused_nnls.hpp:156) */
    int   patience;      /* core/constants.hpp:89 default 5 */
    int   verbose;
} rcppml_b200_config;

typedef struct {
    int    iterations;
    int    converged;
    float  train_loss;
    float  final_tol;
    int    status;          /* 0 ok; 1 = non-positive Cholesky pivot seen */
    int    gpu_launches;    /* kernels launched by the engine since begin_fit */
    double loop_ms;         /* CUDA-event time of iterations run by the last fit/iterate */
} rcppml_b200_result;

/* Cross-validation (nmf/fit_cv.hpp): speckled hold-out mask evaluated in-kernel from a position hash. */
typedef struct {
    float    holdout_fraction;   /* core/config.hpp:236; inv_prob = (uint64)(1/(double)fraction): 0.1f -> 9 */
    uint32_t cv_seed;            /* 0 -> use seed (core/config.hpp:416-418); effective 0 -> 12345 */
    uint32_t seed;
    int      mask_zeros;         /* 1: only non-zeros can be held out; 0: every cell of A is hashed */
    int      cv_patience;        /* core/config.hpp:257 (default 5; <0 -> 5); 0 disables early stopping */
} rcppml_b200_cv_config;

typedef struct {
    float   train_loss, test_loss, best_test_loss;   /* mean squared errors (fit_cv.hpp:1544-1547) */
    int     best_iter;
    int64_t n_test;
} rcppml_b200_cv_result;

/* Named sections follow the reference profiler (profiling/cpu_timer.hpp; fit_cpu.hpp:490-536). */
enum {
    RCPPML_B200_SEC_GRAM_H = 0,        /* "gram_H"            */
    RCPPML_B200_SEC_SOLVE_H = 1,       /* "fused_rhs_nnls_H"  */
    RCPPML_B200_SEC_SCALE_H = 2,       /* "scaling"           */
    RCPPML_B200_SEC_GRAM_W = 3,        /* "gram_W"            */
    RCPPML_B200_SEC_SOLVE_W = 4,       /* "fused_rhs_nnls_W"  */
    RCPPML_B200_SEC_SCALE_W = 5,       /* "scaling"           */
    RCPPML_B200_SEC_LOSS = 6,          /* "loss"              */
    RCPPML_B200_SEC_COMM = 7,          /* multi-GPU exchange  */
    RCPPML_B200_NUM_SECTIONS = 8
};

const char* rcppml_b200_last_error(void);
int  rcppml_b200_engine_create(rcppml_b200_engine** out, int device);
void rcppml_b200_engine_destroy(rcppml_b200_engine* e);

/* A is m x n CSC with ascending row indices per column (dgCMatrix). The engine copies it,
 * builds CSC(A^T) on the device (stable, ascending — Eigen transpose(), nmf/fit_cpu.hpp:251-253)
 * and tr(A^T A) (primitives/primitives.hpp:101-115). Host pointers. */
int rcppml_b200_set_matrix_f32(rcppml_b200_engine* e, int m, int n, int64_t nnz, const int* col_ptr,
                               const int* row_idx, const float* values);
int rcppml_b200_set_matrix_f64(rcppml_b200_engine* e, int m, int n, int64_t nnz, const int* col_ptr,
                               const int* row_idx, const double* values);
/* The caller already holds CSC(A^T) with ascending column ids per row — e.g. a StreamPress .spz file written with
 * include_transpose and decoded by the package's own reader (streampress/sparsepress_v2.hpp:58, :652, :1318; SURVEY.md
 * 8f-4): both operands are uploaded as they are, the device transpose is skipped. Single GPU. */
int rcppml_b200_set_matrix_with_transpose_f32(rcppml_b200_engine* e, int m, int n, int64_t nnz, const int* col_ptr,
                                              const int* row_idx, const float* values, const int* t_col_ptr,
                                              const int* t_row_idx, const float* t_values);
int rcppml_b200_set_matrix_with_transpose_f64(rcppml_b200_engine* e, int m, int n, int64_t nnz, const int* col_ptr,
                                              const int* row_idx, const double* values, const int* t_col_ptr,
                                              const int* t_row_idx, const double* t_values);
/* Synthetic generator of SURVEY.md §8d, on the device. Columns [col_begin, col_begin+n_local)
 * of the m x n_global matrix; per column round(m*density) candidate rows
 * SplitMix64::hash(seed, t, j) mod m, sorted + deduplicated; value 0.5 + uniform<float>(seed+1, r, j). */
int rcppml_b200_set_matrix_synthetic(rcppml_b200_engine* e, int m, int n_local, int col_begin, double density,
                                     uint64_t seed);
/* Sharded variants (after rcppml_b200_comm_init). Rank g of N owns the column block J_g and the row block
 * I_g of equal-size partitions (block = ceil(n/N) resp. ceil(m/N), see rcppml_b200_get_shard) and needs
 * A[:, J_g] (CSC, global row ids) for the H half-step and A[I_g, :] (CSC over all n columns, row ids relative
 * to the block) for the W half-step. */
int rcppml_b200_set_matrix_synthetic_sharded(rcppml_b200_engine* e, int m, int n, double density, uint64_t seed);
int rcppml_b200_set_matrix_sharded_f32(rcppml_b200_engine* e, int m, int n, const int* colblk_ptr,
                                       const int* colblk_idx, const float* colblk_val, const int* rowblk_ptr,
                                       const int* rowblk_idx, const float* rowblk_val);
int rcppml_b200_get_shard(rcppml_b200_engine* e, int* col_begin, int* n_loc, int* row_begin, int* m_loc,
                          int64_t* nnz_global);
/* Explicit partition for the following set_matrix_* calls (after comm_init, same cuts on every rank): rank r owns
 * columns [col_cuts[r], col_cuts[r+1]) of H and rows [row_cuts[r], row_cuts[r+1]) of W_T; world+1 ascending cuts
 * each, 0 .. n and 0 .. m. NULL restores equal blocks. Contiguous ranges balanced by work:
 * rcppml_b200/shard.py balanced_cuts. */
int rcppml_b200_set_partition(rcppml_b200_engine* e, const int* col_cuts, const int* row_cuts);
/* 64-bit checksums of the logical factors {W_T, H, d} (padding excluded), computed on the device: equal
 * checksums <=> bit-identical factors. Used to compare sharded fits with the one-GPU fit of the same seed. */
int rcppml_b200_factor_checksum(rcppml_b200_engine* e, uint64_t* out3);
/* Copies the device CSC operands back (for checking the generator / transpose). Pass NULL to skip an array.
 * get_matrix: A[:, J] (n_loc columns); get_matrix_t: A[I, :]^T (m_loc columns, global column ids). */
int rcppml_b200_get_matrix(rcppml_b200_engine* e, int64_t* nnz, int* col_ptr, int* row_idx, float* values);
int rcppml_b200_get_matrix_t(rcppml_b200_engine* e, int* col_ptr, int* row_idx, float* values);

/* Explicit user mask: CSC pattern (m x n, the WHOLE pattern on every rank of a sharded fit) of the masked entries; mask_nnz = 0 clears it.
 * With a mask set, fits follow the reference's masked path (nmf/masked_nnls.hpp). */
int rcppml_b200_set_mask(rcppml_b200_engine* e, int64_t mask_nnz, const int* mask_col_ptr, const int* mask_row_idx);

/* Factors: W_T is k x m column-major, H is k x n column-major (host, leading dimension k). Always the FULL
 * factors, also when sharded (they are replicated on every rank). */
int rcppml_b200_set_factors_f32(rcppml_b200_engine* e, int k, const float* W_T, const float* H);
int rcppml_b200_set_factors_f64(rcppml_b200_engine* e, int k, const double* W_T, const double* H);
/* nmf/nmf_init.hpp:167-182 on the device: one SplitMix64(seed) stream, W_T first, then H.
 * h_col_begin/h_cols_global place a column shard of H inside the global stream. */
int rcppml_b200_init_factors(rcppml_b200_engine* e, int k, uint32_t seed, int h_col_begin);
int rcppml_b200_get_factors_f32(rcppml_b200_engine* e, float* W_T, float* H, float* d);
int rcppml_b200_get_factors_f64(rcppml_b200_engine* e, double* W_T, double* H, double* d);
/* Sharded fits: only this rank's blocks cross PCIe — W_blk is k x m_loc (rows [row_begin, row_begin + m_loc) of W_T),
 * H_blk is k x n_loc (columns [col_begin, col_begin + n_loc) of H), see rcppml_b200_get_shard. set_: the replicas on
 * every rank are completed by one all-gather per factor over NVLink (every rank must call it). With one rank the
 * blocks are the whole factors. */
int rcppml_b200_set_factor_blocks_f32(rcppml_b200_engine* e, int k, const float* W_blk, const float* H_blk);
int rcppml_b200_get_factor_blocks_f32(rcppml_b200_engine* e, float* W_blk, float* H_blk, float* d);

/* nmf_fit loop (nmf/fit_cpu.hpp:444-1825). begin_fit resets iteration state (iter = 0);
 * iterate enqueues up to n_iters further ALS iterations and returns after they finished. */
int rcppml_b200_begin_fit(rcppml_b200_engine* e, const rcppml_b200_config* cfg);
int rcppml_b200_iterate(rcppml_b200_engine* e, int n_iters);
int rcppml_b200_fit(rcppml_b200_engine* e, const rcppml_b200_config* cfg);   /* begin_fit + iterate(max_iter) */
int rcppml_b200_get_result(rcppml_b200_engine* e, rcppml_b200_result* out);
/* nmf_fit_cv (one GPU or sharded): MSE, standard variant, no user mask. On return H has d absorbed (fit_cv.hpp:1639-1641),
 * get_result gives iterations / converged / final_tol, get_cv_result the losses. */
int rcppml_b200_fit_cv(rcppml_b200_engine* e, const rcppml_b200_config* cfg, const rcppml_b200_cv_config* cv);
int rcppml_b200_get_cv_result(rcppml_b200_engine* e, rcppml_b200_cv_result* out);
int rcppml_b200_get_cv_history(rcppml_b200_engine* e, float* train, float* test, int capacity);
int rcppml_b200_get_loss_history(rcppml_b200_engine* e, float* out, int capacity);
/* Per-section CUDA-event times (ms) and launch counts accumulated since begin_fit. */
int rcppml_b200_set_profiling(rcppml_b200_engine* e, int enabled);
int rcppml_b200_get_profile(rcppml_b200_engine* e, double* ms /*[NUM_SECTIONS]*/, int* launches /*[NUM_SECTIONS]*/);

/* Single half-steps on the resident data, for unit parity tests (fused_nnls.hpp:71,156).
 * which: 0 = H-update (gram(W_T) + solve over columns of A), 1 = W-update. warm_start as fit_cpu.hpp:523. */
int rcppml_b200_half_step(rcppml_b200_engine* e, const rcppml_b200_config* cfg, int which, int warm_start,
                          int normalize_after);

/* Diagnostics: total CD sweeps of the last fit; host<->device bytes moved by the set/get calls; bitwise
 * self-test of the FMA-corrected division the solvers use (must report 0 mismatches vs IEEE). */
int64_t rcppml_b200_cd_sweeps(rcppml_b200_engine* e);
int rcppml_b200_get_counters(rcppml_b200_engine* e, int64_t* h2d_bytes, int64_t* d2h_bytes);
int rcppml_b200_selftest_division(int64_t n, uint64_t seed, int64_t* mismatches);

/* Multi-GPU (one process per GPU). id is an ncclUniqueId (128 bytes) created on rank 0 and distributed by
 * the host launcher. comm_init must be the first call after engine_create; then use the *_sharded setters. */
int rcppml_b200_nccl_unique_id(char* id128);
int rcppml_b200_comm_init(rcppml_b200_engine* e, int rank, int world, const char* id128);
/* Peer-memory fast path (NVLink / NVSwitch P2P, world <= 8). After the factors exist on every rank:
 * export this rank's 192-byte CUDA IPC handle triple {W_T, H, exchange buffer}, all-gather the triples with
 * the host launcher (rank-major, world x 192 bytes), import. From then on the sharded ALS loop makes no
 * NCCL call: half_step_kernel stores every solved column into all replicas while it runs (the factor
 * all-gather overlaps the solve), and the k x k / k-vector fp64 all-reduces are one-shot peer-memory kernels
 * that sum in rank order (bit-identical on every rank). Without these two calls the loop uses NCCL. */
int rcppml_b200_comm_ipc_export(rcppml_b200_engine* e, char* handles192);
int rcppml_b200_comm_ipc_import(rcppml_b200_engine* e, const char* all_handles);
/* NVSwitch multicast replication (NVLS; preferred over the unicast peer stores where the devices support it and
 * RCPPML_B200_MC != 0 — comm_mc_wanted says so after comm_init): the factors are VMM allocations bound to two
 * multicast objects, and the kernel that normalises a freshly solved block writes it into EVERY replica with one
 * multimem.st per word. Same call pattern as the IPC pair: export a 128-byte blob per rank after the factors exist,
 * all-gather the blobs (rank-major, world x 128 bytes), import on every rank, check that every rank succeeded, bind, barrier, finish. The blobs carry POSIX
 * file descriptors that a peer duplicates with pidfd_getfd (same user). On failure on any rank: last_error, then
 * comm_mc_disable on every rank and the IPC pair (unicast peer stores) instead. */
int rcppml_b200_comm_mc_wanted(rcppml_b200_engine* e);
int rcppml_b200_comm_mc_ready(rcppml_b200_engine* e);
int rcppml_b200_comm_mc_export(rcppml_b200_engine* e, char* blob128);
int rcppml_b200_comm_mc_import(rcppml_b200_engine* e, const char* all_blobs);
int rcppml_b200_comm_mc_bind(rcppml_b200_engine* e);      /* after import succeeded on EVERY rank (binding blocks until all joined) */
int rcppml_b200_comm_mc_finish(rcppml_b200_engine* e);
/* The multicast set-up failed on some rank: give multicast up for this engine, keeping the factors (they move into plain
 * allocations that the IPC pair can export). Call on every rank, then use comm_ipc_export / import. */
int rcppml_b200_comm_mc_disable(rcppml_b200_engine* e);
/* Drops every peer mapping (IPC or multicast): the loop falls back to NCCL until the next export / import. */
int rcppml_b200_comm_p2p_close(rcppml_b200_engine* e);

/* The engine behind the reference entry points (part 1) is cached per process: device buffers and staging areas
 * are grow-only, so repeated calls make no cudaMalloc / cudaFree (the reference builds its GPUContext per call,
 * nmf/fit_gpu.cuh:552; SURVEY.md 8b allows process-global caching). RCPPML_B200_CACHE=0 disables the cache;
 * release_cache frees the device memory now. last_call_phases: host wall-clock (ms) of the last part-1 call:
 * [0] matrix upload (+fp64->fp32) [1] device transpose + tr(AtA) [2] factor upload [3] ALS loop [4] factor download. */
int rcppml_b200_release_cache(void);
int rcppml_b200_last_call_phases(double* ms5);
/* Wall clock (ms) of the last rcppml_gpu_nmf_unified_float call, entry to return, measured inside the library. */
double rcppml_b200_last_call_wall_ms(void);
/* The column partition of the in-process multi-GPU path (RCPPML_NUM_GPUS), host only: world + 1 ascending cuts of the n
 * columns, balanced by work = non-zeros + per_item per column (rcppml_b200/shard.py balanced_cuts). */
int rcppml_b200_balanced_col_cuts(const int* col_ptr, int n, int world, int per_item, int* cuts);

/* ---- Part 3: on-disk ingest — StreamPress v2 `.spz` files -------------------------------------------
 * rcppml_sp_read_gpu / rcppml_sp_free_gpu replace src/sp_gpu_bridge.cu:42-123 and :133-155, called through R's .C by
 * st_read_gpu / st_free_gpu (R/sp_gpu.R:53-141): the file is decoded (v2 only: status 4 otherwise; 1 cannot open,
 * 2 read failed, 3 too small, 5 decode error) and left on `device_id` as CSC — int32 col_ptr[n+1], int32 row_idx[nnz],
 * double values[nnz] — the three device addresses returned encoded as doubles, which is what
 * rcppml_gpu_nmf_zerocopy_double takes. The reference's test helper looks the first one up as rcppml_st_read_gpu
 * (tests/testthat/helper-test-utils.R:278); both spellings are exported. */
void rcppml_sp_read_gpu(const char** path_ptr, int* device_id, double* out_col_ptr_addr, double* out_row_idx_addr,
                        double* out_values_addr, int* out_m, int* out_n, double* out_nnz, int* out_status);
void rcppml_sp_free_gpu(double* col_ptr_addr, double* row_idx_addr, double* values_addr, int* out_status);
void rcppml_st_read_gpu(const char** path_ptr, int* device_id, double* out_col_ptr_addr, double* out_row_idx_addr,
                        double* out_values_addr, int* out_m, int* out_n, double* out_nnz, int* out_status);
void rcppml_st_free_gpu(double* col_ptr_addr, double* row_idx_addr, double* values_addr, int* out_status);

/* The reader by itself (host only, no device): what Rcpp_sp_read / Rcpp_sp_read_transpose / Rcpp_sp_metadata
 * (src/sparsepress_bridge.cpp:226-266, :273-286, :293-395) do with streampress::v2::decompress_v2 /
 * decompress_v2_transpose (sparsepress_v2.hpp:897, :1318). Functions return 0 or the status codes above
 * (6: the file has no pre-stored transpose, 7: bad argument, -1: other; text in rcppml_b200_last_error). */
typedef struct rcppml_b200_spz rcppml_b200_spz;
typedef struct {
    int      m, n;
    int64_t  nnz;
    int      chunk_cols, num_chunks;
    int      value_type;          /* header_v2.hpp:45-53: 0 uint8 1 uint16 2 uint32 3 float32 4 float16 5 quant8 6 float64 */
    int      row_sorted;
    int      has_transpose, transpose_chunks, transp_chunk_cols;
    int      has_obs, has_var, has_metadata;
    int      row_permutation_len; /* entries of the stored row permutation (0: none) */
    float    density;
    int64_t  file_bytes, transpose_offset, metadata_offset, metadata_bytes;
    uint32_t stored_crc32;        /* footer (header_v2.hpp:233-266); compare with rcppml_b200_spz_crc32 */
} rcppml_b200_spz_info;
int  rcppml_b200_spz_open(const char* path, rcppml_b200_spz** out);
void rcppml_b200_spz_close(rcppml_b200_spz* h);
int  rcppml_b200_spz_get_info(const rcppml_b200_spz* h, rcppml_b200_spz_info* out);
int  rcppml_b200_spz_crc32(const rcppml_b200_spz* h, uint32_t* computed);
/* section 0: A (n columns); section 1: the pre-stored CSC(A^T) (m columns). Any column range [c0, c1). */
int  rcppml_b200_spz_range_nnz(const rcppml_b200_spz* h, int section, int c0, int c1, int64_t* nnz);
int  rcppml_b200_spz_col_counts(const rcppml_b200_spz* h, int section, int threads, int* counts);
/* col_ptr: c1 - c0 + 1 entries rebased to 0; row_idx / values: range_nnz entries. reorder: apply the stored row
 * permutation as decompress_v2 does (section 0 only). threads <= 0: every core. */
int  rcppml_b200_spz_read_f32(const rcppml_b200_spz* h, int section, int c0, int c1, int reorder, int threads,
                              int* col_ptr, int* row_idx, float* values);
int  rcppml_b200_spz_read_f64(const rcppml_b200_spz* h, int section, int c0, int c1, int reorder, int threads,
                              int* col_ptr, int* row_idx, double* values);
/* A[row_begin : row_begin + m_loc, :] as CSC over all n columns with block-relative row ids (the row-block operand of
 * rcppml_b200_set_matrix_sharded_f32) from a file WITHOUT a transpose section: full decode + host filter. capacity:
 * entries row_idx / values can take (the file's nnz always suffices). */
int  rcppml_b200_spz_row_block_f32(const rcppml_b200_spz* h, int row_begin, int m_loc, int threads, int* col_ptr,
                                   int* row_idx, float* values, int64_t capacity, int64_t* nnz);
/* Raw bytes of a metadata record (header_v2.hpp:108-113: 0 rownames, 1 colnames — NUL-separated —, 2 row permutation). */
int  rcppml_b200_spz_metadata(const rcppml_b200_spz* h, int key, unsigned char* buf, int64_t capacity, int64_t* bytes);
/* File -> engine. After comm_init (+ optional set_partition): this rank decodes only its column block of A and its row
 * block (columns of the stored transpose section); collective like every sharded set_matrix_*. stored_transpose: 1 use
 * the file's transpose section when it has a usable one, 0 never (device transpose / host filter instead), < 0 automatic:
 * sharded yes, one GPU no (the device transposes 1e8 entries in 6 ms; entropy-decoding them takes the host ~0.5 s).
 * *used_stored_transpose reports what happened. */
int  rcppml_b200_set_matrix_spz(rcppml_b200_engine* e, const rcppml_b200_spz* h, int threads, int stored_transpose,
                                int* used_stored_transpose);
/* Sharded operands with the row block already transposed (what the two sections of a .spz file deliver):
 * A[:, J] as CSC (n_loc columns, global row ids) and (A[I, :])^T as CSC (m_loc columns, global column ids). */
int  rcppml_b200_set_matrix_sharded_with_transpose_f32(rcppml_b200_engine* e, int m, int n, const int* colblk_ptr,
                                                       const int* colblk_idx, const float* colblk_val,
                                                       const int* tblk_ptr, const int* tblk_idx, const float* tblk_val);

#ifdef __cplusplus
}
#endif
#endif /* RCPPML_GPU_H */
