// comm.cu — column-sharded ALS over N GPUs (one process per GPU, NCCL over NVLink/NVSwitch).
//
// Rank g holds A[:, J_g] (CSC), its transpose, H[:, J_g]; W_T, d and the k×k Grams are replicated.
//   H half-step : local fused gather+solve; Σ|H| row sums all-reduced (k doubles); the Gram of the
//                 normalised H is a sum of per-rank Grams (k×k doubles, all-reduce).
//   W half-step : B = Σ_g H_g·A_gᵀ — every rank forms its k×m partial with the gather kernel
//                 (OUT_RHS), ONE reduce-scatter hands rank g the fully reduced rows of its row block,
//                 rank g solves those m/N rows (BSRC_LOAD), normalises them and ONE all-gather
//                 replicates the new W_T. Row sums, Gram and the loss cross term are all-reduced.
// The reference has no multi-GPU path; this follows SURVEY.md §8e.
#include "engine.hpp"

#include <nccl.h>

#include <cstring>

namespace b200 {

#define B200_NCCL_CHECK(expr)                                                                       \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess)                                                                      \
            throw std::runtime_error(std::string(#expr) + " failed: " + ncclGetErrorString(_r));    \
    } while (0)

static inline ncclComm_t as_comm(ncclComm* c) { return reinterpret_cast<ncclComm_t>(c); }

void Engine::comm_init(int rank_, int world_, const char* id128) {
    use_device();
    B200_REQUIRE(world_ >= 1 && rank_ >= 0 && rank_ < world_, "comm_init: bad rank/world");
    B200_REQUIRE(!factors_ready, "comm_init must precede set_factors/init_factors (W_T is padded to equal row blocks)");
    B200_REQUIRE(matrix_ready, "comm_init: set the local column shard first");
    rank = rank_;
    world = world_;
    if (world == 1) return;
    ncclUniqueId id;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    B200_NCCL_CHECK(ncclCommInitRank(&c, world, id, rank));
    comm = reinterpret_cast<ncclComm*>(c);
    // tr(AᵀA) over all shards (fp64)
    DeviceBuffer<double> t;
    t.ensure(1);
    B200_CUDA_CHECK(cudaMemcpyAsync(t.ptr, &trAtA_local, sizeof(double), cudaMemcpyHostToDevice, stream));
    B200_NCCL_CHECK(ncclAllReduce(t.ptr, t.ptr, 1, ncclDouble, ncclSum, as_comm(comm), stream));
    double total = 0.0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&total, t.ptr, sizeof(double), cudaMemcpyDeviceToHost, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    trAtA = static_cast<float>(total);
}

void Engine::comm_destroy() {
    if (comm) {
        ncclCommDestroy(as_comm(comm));
        comm = nullptr;
    }
}

void Engine::allreduce_f64(double* buf, size_t count) {
    B200_NCCL_CHECK(ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, as_comm(comm), stream));
}

void Engine::enqueue_iteration_sharded() {
    const bool warm = iters_enqueued > 0;
    const bool normalize = cfg.norm_type != 2;
    const int solver = cfg.solver_mode == 0 ? SOLVER_CD : SOLVER_CHOL;
    const int mb = m_pad / world;                         // rows per rank
    row_begin = rank * mb;
    row_count = std::max(0, std::min(m, row_begin + mb) - row_begin);
    if (B_part.count < static_cast<size_t>(m_pad) * KP) {
        B_part.ensure(static_cast<size_t>(m_pad) * KP);
        B_blk.ensure(static_cast<size_t>(mb) * KP);
        B200_CUDA_CHECK(cudaMemsetAsync(B_part.ptr, 0, static_cast<size_t>(m_pad) * KP * sizeof(float), stream));
    }

    // ---- H update: local columns, replicated W_T / G_w
    if (iters_enqueued == 0) gram(W_T.ptr, m, false, G_w.ptr, RCPPML_B200_SEC_GRAM_H);     // identical on every rank
    prepare_solver(G_w.ptr, cfg.L2_H, RCPPML_B200_SEC_GRAM_H);
    solve(0, warm, RCPPML_B200_SEC_SOLVE_H);
    scale_finalize(RCPPML_B200_SEC_SCALE_H, /*reduce_over_ranks=*/true);
    gram(H.ptr, n, normalize, G_h.ptr, RCPPML_B200_SEC_GRAM_W, /*reduce_over_ranks=*/true);
    prepare_solver(G_h.ptr, cfg.L2_W, RCPPML_B200_SEC_GRAM_W);

    // ---- W update, step 1: partial right-hand side of every row from the local columns
    {
        HalfStepParams p = solve_params(1, warm);
        p.B = B_part.ptr;
        p.X = nullptr;
        p.work_counter = counters.ptr + 2;
        sec_begin(RCPPML_B200_SEC_SOLVE_W);
        B200_CUDA_CHECK(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));
        launch_half_step(LANES, SOLVER_CD, BSRC_GATHER, OUT_RHS, p, num_sms, stream);
        launches[RCPPML_B200_SEC_SOLVE_W] += 1;
        sec_end(RCPPML_B200_SEC_SOLVE_W);
    }
    // ---- step 2: reduce-scatter by row blocks (NVLink; NCCL sums in a fixed order for a fixed communicator)
    sec_begin(RCPPML_B200_SEC_COMM);
    B200_NCCL_CHECK(ncclReduceScatter(B_part.ptr, B_blk.ptr, static_cast<size_t>(mb) * KP, ncclFloat, ncclSum,
                                      as_comm(comm), stream));
    sec_end(RCPPML_B200_SEC_COMM);
    // ---- step 3: solve my row block from the reduced right-hand sides
    {
        HalfStepParams p = solve_params(1, warm);
        p.B = B_blk.ptr;
        p.nslots = 1;
        p.slot_stride = 0;
        p.b_local_index = 1;
        p.ncols = row_count;
        p.col_offset = row_begin;
        p.cols_per_fetch = 8;
        p.work_counter = counters.ptr + 3;
        const int geom = geometry_for(0, row_count);          // pure solve: two words per lane when available
        int grid = 0;
        launch_half_step(geom, solver, BSRC_LOAD, OUT_SOLVE, p, num_sms, stream, &grid);
        last_solve_grid = grid;
        sec_begin(RCPPML_B200_SEC_SOLVE_W);
        B200_CUDA_CHECK(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));
        launch_half_step(geom, solver, BSRC_LOAD, OUT_SOLVE, p, num_sms, stream);
        launches[RCPPML_B200_SEC_SOLVE_W] += 1;
        sec_end(RCPPML_B200_SEC_SOLVE_W);
    }
    scale_finalize(RCPPML_B200_SEC_SCALE_W, /*reduce_over_ranks=*/true);       // d and the loss cross term
    // ---- step 4: normalise my rows + my share of gram(W_T); all-reduce the Gram; all-gather W_T
    sec_begin(RCPPML_B200_SEC_LOSS);
    const bool was = profiling; profiling = false;
    gram(W_T.ptr + static_cast<size_t>(row_begin) * KP, row_count, normalize, G_w.ptr, RCPPML_B200_SEC_LOSS,
         /*reduce_over_ranks=*/true);
    profiling = was;
    loss(RCPPML_B200_SEC_LOSS);
    sec_end(RCPPML_B200_SEC_LOSS);
    sec_begin(RCPPML_B200_SEC_COMM);
    B200_NCCL_CHECK(ncclAllGather(W_T.ptr + static_cast<size_t>(row_begin) * KP, W_T.ptr, static_cast<size_t>(mb) * KP,
                                  ncclFloat, as_comm(comm), stream));
    sec_end(RCPPML_B200_SEC_COMM);
    ++iters_enqueued;
}

}  // namespace b200

extern "C" {

int rcppml_b200_nccl_unique_id(char* id128) {
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) {
        b200::g_last_error = std::string("ncclGetUniqueId failed: ") + ncclGetErrorString(r);
        return -1;
    }
    std::memcpy(id128, &id, sizeof(id));
    return 0;
}

int rcppml_b200_comm_init(rcppml_b200_engine* e, int rank, int world, const char* id128) {
    try {
        e->impl.comm_init(rank, world, id128);
        return 0;
    } catch (const std::exception& ex) {
        b200::g_last_error = ex.what();
        return -1;
    }
}

}  // extern "C"
