// engine.hpp — the device-resident ALS engine (one instance per GPU / per process rank).
#pragma once

#include "../../include/rcppml_gpu.h"
#include "common.cuh"
#include "kernels_dense.cuh"
#include "kernels_solve.cuh"
#include "kernels_sparse.cuh"

#include <array>
#include <string>
#include <utility>
#include <vector>

struct ncclComm;

namespace b200 {

extern thread_local std::string g_last_error;

// grid_out != nullptr: only report the persistent grid this instantiation uses (buffer sizing).
void launch_half_step(int lanes, int solver, int bsrc, int out, const HalfStepParams& p, int num_sms,
                      cudaStream_t s, int* grid_out = nullptr);

class Engine {
public:
    explicit Engine(int device);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    void use_device() const;

    // matrix (column shard [col_begin, col_begin + n) of an m × n_global matrix when world > 1)
    template <class ValT>
    void set_matrix_host(int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx, const ValT* values);
    void set_matrix_synthetic(int m, int n_local, int col_begin, double density, uint64_t seed);

    // factors
    template <class T> void set_factors_host(int k, const T* W_T, const T* H);
    template <class T> void get_factors_host(T* W_T, T* H, T* d);
    void init_factors(int k, uint32_t seed, int h_col_begin);

    // fit
    void begin_fit(const rcppml_b200_config& cfg);
    void iterate(int n_iters);
    void half_step_only(const rcppml_b200_config& cfg, int which, bool warm, bool normalize_after);
    void get_result(rcppml_b200_result* out);

    // multi-GPU (comm.cu)
    void comm_init(int rank, int world, const char* id128);
    void comm_destroy();
    bool comm_ready() const { return comm != nullptr; }

    // ---- state (public: the C ABI shims read it) ------------------------------------------
    int device = 0;
    int num_sms = kNumSMs;
    cudaStream_t stream = nullptr;

    int m = 0, n = 0, col_begin = 0;
    int64_t nnz = 0;
    bool matrix_ready = false, factors_ready = false, fit_active = false;
    DeviceBuffer<int> Ap, Ai, Atp, Ati;
    DeviceBuffer<float> Ax, Atx;
    float trAtA = 0.f;          // global tr(AᵀA) (summed over ranks after comm_init)
    double trAtA_local = 0.0;

    int k = 0, KP = 0, LANES = 0, nv_override = 0;
    int geometry_for(long long nnz, long long ncols) const;
    DeviceBuffer<float> W_T, H, d;
    DeviceBuffer<float> G_w, G_h, M1, M2, dblk, rcp;
    DeviceBuffer<double> gram_partials, solve_partials;   // per-CTA partials (fixed-order reductions)
    DeviceBuffer<double> red_gram, red_small;             // reduced sums: k×k Gram | k row sums + cross term
    DeviceBuffer<int> counters;
    DeviceBuffer<unsigned long long> sweep_counter;
    DeviceBuffer<DevState> state;
    DeviceBuffer<float> loss_hist;
    DevState* h_state = nullptr;
    int gram_grid = 0, solve_grid_max = 0, last_solve_grid = 0;

    rcppml_b200_config cfg{};
    int iters_enqueued = 0;
    double loop_ms = 0.0;
    unsigned long long cd_sweeps = 0;
    size_t h2d_bytes = 0, d2h_bytes = 0;

    bool profiling = false;
    std::array<std::vector<std::pair<cudaEvent_t, cudaEvent_t>>, RCPPML_B200_NUM_SECTIONS> prof_events;
    std::array<int, RCPPML_B200_NUM_SECTIONS> prof_used{};
    std::array<double, RCPPML_B200_NUM_SECTIONS> prof_ms{};
    std::array<int, RCPPML_B200_NUM_SECTIONS> launches{};

    // multi-GPU
    ncclComm* comm = nullptr;
    int rank = 0, world = 1;
    DeviceBuffer<float> B_part;       // m_pad × KP partial right-hand side of the W-update (this rank's columns)
    DeviceBuffer<float> B_blk;        // reduced row block owned by this rank
    int m_pad = 0;                    // m rounded up to a multiple of world (equal row blocks)
    int row_begin = 0, row_count = 0; // rows of A (columns of Aᵀ) this rank solves in the W-update

private:
    cudaEvent_t ev_loop_begin = nullptr, ev_loop_end = nullptr;

    void finish_matrix();
    void build_transpose();
    void alloc_factors(int k);
    void normalize_cfg(const rcppml_b200_config& c);
    void sec_begin(int sec);
    void sec_end(int sec);
    void collect_profile();
    void gram(float* X, long long ncols, bool normalize, float* G_out, int sec, bool reduce_over_ranks = false);
    void loss(int sec);
    void allreduce_f64(double* buf, size_t count);
    void prepare_solver(const float* G, float L2, int sec);
    HalfStepParams solve_params(int which, bool warm) const;
    void solve(int which, bool warm, int sec);
    void scale_finalize(int sec, bool reduce_over_ranks = false);
    void enqueue_iteration();
    void enqueue_iteration_sharded();
};

}  // namespace b200

// The opaque handle of the C ABI.
struct rcppml_b200_engine {
    b200::Engine impl;
    explicit rcppml_b200_engine(int dev) : impl(dev) {}
};
