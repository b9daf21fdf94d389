// engine.hpp — the device-resident ALS engine (one instance per GPU / per process rank).
#pragma once

#include "../../include/rcppml_gpu.h"
#include "common.cuh"
#include "vmm.hpp"
#include "kernels_dense.cuh"
#include "kernels_masked.cuh"
#include "kernels_cv.cuh"
#include "kernels_solve.cuh"
#include "kernels_cd.cuh"
#include "kernels_tiled.cuh"
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
void launch_cd_half_step(int geom, const HalfStepParams& p, int num_sms, cudaStream_t s, int* grid_out = nullptr);
void launch_tiled_half_step(int gather_geom, int solver, const HalfStepParams& p, int num_sms, cudaStream_t s,
                            int* grid_out = nullptr);

class Engine {
public:
    explicit Engine(int device);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    void use_device() const;

    // matrix: single GPU (whole A) or this rank's column block + row block (see engine.cu)
    template <class ValT>
    void set_matrix_host(int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx, const ValT* values);
    template <class ValT>
    void set_matrix_host_with_transpose(int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx, const ValT* values,
                                        const int* t_col_ptr, const int* t_row_idx, const ValT* t_values);
    template <class ValT>
    void set_matrix_sharded(int m, int n, const int* cb_ptr, const int* cb_idx, const ValT* cb_val, const int* rb_ptr,
                            const int* rb_idx, const ValT* rb_val);
    // Sharded, the row block already transposed: A[:, J] (CSC, n_loc columns, global row ids) and (A[I, :])^T (CSC,
    // m_loc columns, global column ids ascending) — e.g. two column ranges of a .spz file with a pre-stored transpose.
    template <class ValT>
    void set_matrix_sharded_with_transpose(int m, int n, const int* cb_ptr, const int* cb_idx, const ValT* cb_val,
                                           const int* tb_ptr, const int* tb_idx, const ValT* tb_val);
    // The blocks set_dims would give this rank for an m x n matrix (pending explicit cuts, else equal blocks).
    void planned_blocks(int m, int n, int* col_begin_out, int* n_loc_out, int* row_begin_out, int* m_loc_out) const;
    void set_matrix_synthetic(int m, int n_local, int col_begin, double density, uint64_t seed);
    void set_matrix_synthetic_sharded(int m, int n, double density, uint64_t seed);

    void set_mask(int64_t mask_nnz, const int* mask_ptr, const int* mask_idx);

    // factors
    template <class T> void set_factors_host(int k, const T* W_T, const T* H);
    // Host matrix + host factors in one call (the reference entry points): the factor H2D copies run on the copy
    // engine while the device transposes A, instead of after it.
    template <class ValT, class T>
    void set_matrix_and_factors_host(int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx, const ValT* values,
                                     int k, const T* W_T, const T* H);
    template <class T> void get_factors_host(T* W_T, T* H, T* d);
    // Sharded fits: move only this rank's blocks over PCIe — rows [row_begin, row_begin + m_loc) of W_T and columns
    // [col_begin, col_begin + n_loc) of H; the replicas are completed by one all-gather per factor over NVLink.
    template <class T> void set_factor_blocks_host(int k, const T* W_blk, const T* H_blk);
    template <class T> void get_factor_blocks_host(T* W_blk, T* H_blk, T* d);
    void init_factors(int k, uint32_t seed, int h_col_begin);

    // fit
    void begin_fit(const rcppml_b200_config& cfg);
    void iterate(int n_iters);
    void half_step_only(const rcppml_b200_config& cfg, int which, bool warm, bool normalize_after);
    void get_result(rcppml_b200_result* out);
    void fit_cv(const rcppml_b200_config& cfg, const rcppml_b200_cv_config& cv);
    void get_cv_result(rcppml_b200_cv_result* out);

    // multi-GPU (comm.cu)
    void comm_init(int rank, int world, const char* id128);
    void comm_destroy();
    bool comm_ready() const { return comm != nullptr; }
    // Peer-memory fast path (NVLink P2P): exchange CUDA IPC handles of {W_T, H, exchange buffer} between ranks.
    void comm_ipc_export(char* handles192);
    void comm_ipc_import(const char* all_handles /* world x 192 bytes, rank-major */);
    void comm_ipc_close();
    // In-process multi-GPU (RCPPML_NUM_GPUS behind the reference entry points, abi_reference.cu): one Engine per
    // device inside ONE process, each driven by its own host thread. No NCCL and no IPC: the peers' buffers are
    // plain device pointers once peer access is enabled, and the peer-memory loop is the only loop.
    void comm_init_local(int rank, int world);
    void comm_prepare_local(const int* devices);          // exchange buffer + cudaDeviceEnablePeerAccess to every peer
    void comm_attach_local(Engine* const* all);           // after EVERY engine ran comm_prepare_local
    void enable_peer_access(const int* devices);          // cudaDeviceEnablePeerAccess to every peer (idempotent)
    // Sharded ingest of a HOST matrix (engine.cu, "in-process multi-GPU ingest"): every device uploads only its column
    // block; row blocks are assembled from the peers' column blocks over NVLink; factors travel block-wise.
    template <class ValT>
    void upload_col_block_host(int m, int n, const int* col_ptr, const int* row_idx, const ValT* values);
    void balanced_row_cuts(Engine* const* all, int per_item, int* cuts_out);
    void install_row_cuts(const int* row_cuts_in);
    double assemble_row_block(Engine* const* all);
    void finish_matrix_local(double sumsq_total, int64_t nnz_total);
    template <class T> void start_factor_block_upload(int k, const T* W_full, const T* H_full);
    template <class T> void finish_factor_block_upload(int k);
    void pull_factor_blocks_from_peers();
    DeviceBuffer<int> row_hist;                           // [m] entries per row of the own column block

    // ---- state (public: the C ABI shims read it) ------------------------------------------
    int device = 0;
    int num_sms = kNumSMs;
    cudaStream_t stream = nullptr;

    int m = 0, n = 0;                       // global shape of A
    int col_begin = 0, n_loc = 0;           // this rank's column block J (H half-step)
    int row_begin = 0, m_loc = 0;           // this rank's row block I (W half-step)
    int m_pad = 0, n_pad = 0;               // m, n rounded up to a multiple of world (equal all-gather blocks)
    // Partition over the ranks: rank r owns columns [col_cuts[r], col_cuts[r+1]) of H and rows [row_cuts[r],
    // row_cuts[r+1]) of W_T. Default: equal blocks of ceil(total/world). set_partition installs explicit cuts
    // (contiguous ranges balanced by work) for the following set_matrix_* calls; nullptr restores
    // the default. Every rank must install the same cuts.
    std::vector<int> col_cuts, row_cuts, pending_col_cuts, pending_row_cuts;
    bool equal_partition = true;
    void set_partition(const int* col_cuts_in, const int* row_cuts_in);
    void factor_checksum(unsigned long long* out3);
    int64_t nnz = 0, nnz_w = 0, nnz_global = 0;   // nnz of A[:,J], of A[I,:], of A
    bool matrix_ready = false, factors_ready = false, fit_active = false;
    DeviceBuffer<int> Ap, Ai, Atp, Ati;
    DeviceBuffer<float> Ax, Atx;
    bool has_mask = false;                  // explicit user mask (nmf/masked_nnls.hpp); pattern + transposed pattern
    int64_t mask_nnz = 0;
    DeviceBuffer<int> Mp, Mi, MTp, MTi;
    DeviceBuffer<double> loss_partials;
    // speckled-mask cross-validation state (nmf/fit_cv.hpp)
    bool cv_active = false;
    rcppml_b200_cv_config cv{};
    unsigned long long cv_seed_state = 0, cv_inv_prob = 0, cv_threshold = 0;
    DeviceBuffer<CvState> cv_state;
    DeviceBuffer<float> test_hist;
    float trAtA = 0.f;          // tr(AᵀA) of the whole matrix

    int k = 0, KP = 0, LANES = 0, nv_override = 0, nv_short_override = 0;
    int geometry_for(long long nnz, long long ncols) const;
    int cd_geom = 0, cd_geom_long = 0;      // lane-group geometry of cd_half_step_kernel for short / long columns
                                            // (0: use half_step_kernel<CD>)
    void launch_solver(int kind, int geom, int solver, const HalfStepParams& p, int* grid_out = nullptr);
    // kernels_tiled.cuh (wide gather -> shared-memory tile -> narrow solve). 0: never, 1: CD always and Cholesky
    // for short columns (default), 2: always. RCPPML_B200_TILED overrides.
    int tiled_mode = 1;
    double tiled_min_batches = 0.5, narrow_min_cols = 0.0;
    int side_ctas_per_sm = 8;               // grid of the peers' re-normalisation kernel (side stream), CTAs per SM
    int tiled_cta_mode = 0;                 // k = 64 Cholesky tiled kernel: 0 = 3 x 256-thread CTAs per SM, 1 = one 768-thread CTA,
                                            // 2 = one 768-thread CTA + hybrid register / cp.async-ring gather (RCPPML_B200_TILED_CTA)
    int tiled_sl_override = 0;              // k = 64 tiled kernel: solve lanes per column (0: rule, 2: 16-column batches, 4: 8)
    bool use_narrow_cd(long long ncols) const;
    bool use_tiled(int solver, long long cnt, long long ncols) const;
    int tiled_gather_geom(long long cnt, long long ncols, int solver) const;
    FactorBuffer W_T, H;                    // replicated factors (VMM-backed + multicast-bound in sharded fits, vmm.hpp)
    DeviceBuffer<float> d;
    DeviceBuffer<float> G_w, G_h, M1, M2, dblk;     // dblk: SolverConsts image (diagonal blocks, then reciprocals)
    int const_slot = 0;
    DeviceBuffer<double> gram_partials, solve_partials;   // per-CTA partials (fixed-order reductions)
    DeviceBuffer<double> red_gram, red_small;             // reduced sums: k×k Gram | k row sums + cross term
    // row-panel passes for half-steps whose gathered factor exceeds L2 (build_panels, kernels_solve.cuh)
    DeviceBuffer<int> panel_bounds[2];                    // [which][(P+1) x ncols]
    int npanels[2] = {1, 1};
    DeviceBuffer<float> carry;                            // [max(n_loc, m_loc)][KP] running right-hand sides
    void build_panels();
    DeviceBuffer<int> counters;
    DeviceBuffer<unsigned long long> sweep_counter;
    DeviceBuffer<DevState> state;
    DeviceBuffer<float> loss_hist;
    DevState* h_state = nullptr;
    int gram_grid = 0, solve_grid_max = 0, last_solve_grid = 0;

    rcppml_b200_config cfg{};
    // Steady-state iteration (iteration >= 1, single GPU, plain path) captured once per fit as a CUDA graph and
    // replayed: on small matrices (C2/C3) an iteration is ~20 launches of a few microseconds each and the loop is
    // launch bound. RCPPML_B200_GRAPH=0 disables. Not used while per-section profiling records events.
    cudaGraphExec_t iter_graph = nullptr;
    bool graphs_enabled = true;
    bool capturing = false;
    std::array<int, RCPPML_B200_NUM_SECTIONS> graph_launches{};
    void capture_iteration_graph();
    void drop_iteration_graph();
    int iters_enqueued = 0;
    double loop_ms = 0.0;
    unsigned long long cd_sweeps = 0;
    size_t h2d_bytes = 0, d2h_bytes = 0;

    // Grow-only staging buffers (fp64 wire copies, transpose temporaries): a cached engine (abi_reference.cu)
    // makes its second and later calls without a single cudaMalloc / cudaFree.
    std::array<DeviceBuffer<unsigned char>, 13> scratch_;
    template <class T> T* scratch(int slot, size_t count) {
        scratch_[slot].ensure(count * sizeof(T));
        return reinterpret_cast<T*>(scratch_[slot].ptr);
    }
    void release_scratch() { for (auto& b : scratch_) b.release(); }
    // Host wall-clock of the phases of the last set_matrix / set_factors / iterate / get_factors sequence (ms):
    // [0] matrix upload (+fp64->fp32), [1] device transpose + tr(AtA), [2] factor upload, [3] ALS loop, [4] factor download
    double phase_ms[5] = {0, 0, 0, 0, 0};

    bool profiling = false;
    std::array<std::vector<std::pair<cudaEvent_t, cudaEvent_t>>, RCPPML_B200_NUM_SECTIONS> prof_events;
    std::array<int, RCPPML_B200_NUM_SECTIONS> prof_used{};
    std::array<double, RCPPML_B200_NUM_SECTIONS> prof_ms{};
    std::array<int, RCPPML_B200_NUM_SECTIONS> launches{};

    // multi-GPU
    ncclComm* comm = nullptr;
    int rank = 0, world = 1;
    bool peers_ready = false;
    bool peers_local = false;         // peers live in this process (comm_attach_local): nothing to IPC-close
    bool peer_access_enabled = false; // in-process peers: cudaDeviceEnablePeerAccess done for this engine's device
    bool peers_vmm = false;           // cross-process peers mapped from imported VMM handles (multicast path), not cudaIpc
    float* peer_W[8] = {};            // every rank's W_T / H / exchange buffer (own pointers at [rank])
    float* peer_H[8] = {};
    double* peer_x[8] = {};
    DeviceBuffer<double> xbuf;        // data[2][world][ne_max] + flags[2][8] + sequence counter (kXchgTailWords)
    int xchg_ne_max = 0;
    // NVSwitch multicast (NVLS) replication of the factors. mc_wanted is decided with the communicator (device support,
    // RCPPML_B200_MC != 0): the factors are then VMM allocations. Once every rank has bound its W_T / H to the two
    // multicast objects (mc_ready), the kernel that NORMALISES a freshly solved block (normalize_gram_*) writes it with
    // multimem.st — one store per word lands in all N replicas through the switch — and the solve kernels stop pushing
    // unicast copies; the "every rank re-normalises the peers' blocks" pass disappears with them.
    bool mc_wanted = false, mc_ready = false;
    int mc_mode = 2;                  // 1: the normalising Gram kernel writes the block into every replica (no peer-side
                                      // re-normalisation); 2: the solve kernel's stores go through the multicast alias
                                      // (overlapped with the solve; peers re-normalise as with unicast). RCPPML_B200_MC.
    MappedHandle mcW, mcH;            // the multicast objects mapped on this device
    bool mc_owner = false;            // this engine created the multicast objects
    bool mc_local = false;            // in-process group: the handles are shared and owned by the orchestrator
    MappedHandle peer_map_W[8], peer_map_H[8];   // cross-process: the peers' physical allocations mapped here
    int mc_export_fds[4] = {-1, -1, -1, -1};
    bool mc_ptracer_window = false;   // PR_SET_PTRACER_ANY is in force (export .. finish)
    void mc_decide();                                        // after rank / world are known
    void mc_grant_local_access(const int* devices);          // in-process: every device of the group may access W_T / H
    void mc_create(CUmemGenericAllocationHandle* hW, CUmemGenericAllocationHandle* hH, bool shareable);
    void mc_add_device(CUmemGenericAllocationHandle hW, CUmemGenericAllocationHandle hH);
    void mc_bind_and_map(CUmemGenericAllocationHandle hW, CUmemGenericAllocationHandle hH, bool owner);
    void mc_close();
    void comm_mc_export(char* blob128);                      // cross-process (one process per GPU)
    void comm_mc_import(const char* all_blobs);
    void comm_mc_bind();
    void comm_mc_finish();
    void mc_disable_keep_factors();                          // multicast set-up failed somewhere: plain buffers, same contents
    float* mc_alias(const float* replica_ptr) const;         // multicast address of a word of W_T / H (nullptr: not ready)

private:
    cudaEvent_t ev_loop_begin = nullptr, ev_loop_end = nullptr;
    void init_device_objects();

    void set_dims(int m, int n);
    void finish_matrix();
    void finish_matrix_from(const float* vals, int64_t cnt, bool reduce_over_ranks);
    void transpose_csc(const int* sp, const int* si, const float* sx, int ncols, int nrows, int64_t cnt,
                       DeviceBuffer<int>& dp, DeviceBuffer<int>& di, DeviceBuffer<float>& dx, int col_id_offset);
    template <class ValT>
    void upload_csc(int ncols, int64_t cnt, const int* col_ptr, const int* row_idx, const ValT* values,
                    DeviceBuffer<int>& dp, DeviceBuffer<int>& di, DeviceBuffer<float>& dx);
    void synth_block(int m, int c0, int nc, int r0, int r1, double density, uint64_t seed, DeviceBuffer<int>& dp,
                     DeviceBuffer<int>& di, DeviceBuffer<float>& dx, int64_t* cnt_out);
    void allgather_rows(float* buf, const std::vector<int>& cuts, int rows_padded, int sec);
    void alloc_factors(int k);
    void normalize_cfg(const rcppml_b200_config& c);
    void sec_begin(int sec, cudaStream_t on = nullptr);
    void sec_end(int sec, cudaStream_t on = nullptr);
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_pending = false, side_forked = false;
    void fork_side_stream();
    void join_side_stream();
    void collect_profile();
    void gram(float* X, long long ncols, bool normalize, float* G_out, int sec, bool reduce_over_ranks = false,
              bool replicate = false);
    void normalize_peer_blocks(float* X, long long ncols, long long lo, long long hi, bool normalize);
    void loss(int sec);
    void allreduce_f64(double* buf, size_t count);
    void prepare_solver(const float* G, float L2, int sec);
    HalfStepParams solve_params(int which, bool warm) const;
    void solve(int which, bool warm, int sec);
    void scale_finalize(int sec, bool reduce_over_ranks = false);
    void enqueue_iteration();
    void enqueue_iteration_masked();
    void enqueue_iteration_cv();
    void cv_solve(int which, int sec);
    void masked_solve(int which, bool warm, const float* G, int sec);
};

}  // namespace b200

// The opaque handle of the C ABI.
struct rcppml_b200_engine {
    b200::Engine impl;
    explicit rcppml_b200_engine(int dev) : impl(dev) {}
};
