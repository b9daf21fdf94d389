// comm.cu — NCCL plumbing of the sharded ALS (one process per GPU, NVLink/NVSwitch).
//
// Rank g owns column block J_g of H and row block I_g of W_T and holds the sparse operands A[:,J_g] and
// A[I_g,:]ᵀ (engine.cu). Per iteration the only traffic is: all-gather of the H blocks (k×n floats in total),
// all-gather of the W_T blocks (k×m floats), and fp64 all-reduces of the k×k Grams, the k row sums and the
// loss cross term. The k×m right-hand side of the W-update is never exchanged: each rank gathers the
// right-hand sides of ITS rows from the replicated H, so sharded and single-GPU fits perform the same
// per-column arithmetic. The reference has no multi-GPU path.
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
    B200_REQUIRE(!matrix_ready && !factors_ready, "comm_init must come first: blocks and padding depend on (rank, world)");
    rank = rank_;
    world = world_;
    if (world == 1) return;
    ncclUniqueId id;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    B200_NCCL_CHECK(ncclCommInitRank(&c, world, id, rank));
    comm = reinterpret_cast<ncclComm*>(c);
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

// In-place all-gather of equal row blocks of a replicated factor: rank g contributes rows
// [g*rows_per_rank, (g+1)*rows_per_rank).
void Engine::allgather_rows(float* buf, int rows_per_rank, int sec) {
    sec_begin(sec);
    const size_t cnt = static_cast<size_t>(rows_per_rank) * KP;
    B200_NCCL_CHECK(ncclAllGather(buf + static_cast<size_t>(rank) * cnt, buf, cnt, ncclFloat, as_comm(comm), stream));
    sec_end(sec);
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
