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

void Engine::comm_ipc_close() {
    if (!peers_ready) return;
    cudaStreamSynchronize(stream);
    for (int r = 0; r < world && !peers_local; ++r) {
        if (r == rank) continue;
        cudaIpcCloseMemHandle(peer_W[r]);
        cudaIpcCloseMemHandle(peer_H[r]);
        cudaIpcCloseMemHandle(peer_x[r]);
    }
    peers_ready = false;
    peers_local = false;
}

// ---- in-process multi-GPU (abi_reference.cu: RCPPML_NUM_GPUS) ------------------------------------------------
void Engine::comm_init_local(int rank_, int world_) {
    use_device();
    B200_REQUIRE(world_ >= 1 && world_ <= 8 && rank_ >= 0 && rank_ < world_, "comm_init_local: bad rank/world");
    B200_REQUIRE(!matrix_ready && !factors_ready, "comm_init_local must come first: blocks and padding depend on (rank, world)");
    rank = rank_;
    world = world_;
}

void Engine::enable_peer_access(const int* devices) {
    use_device();
    if (peer_access_enabled) return;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        int can = 0;
        B200_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, device, devices[r]));
        B200_REQUIRE(can, "no peer access between the selected devices (NVLink / NVSwitch required)");
        const cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else B200_CUDA_CHECK(e);
    }
    peer_access_enabled = true;
}

void Engine::comm_prepare_local(const int* devices) {
    use_device();
    B200_REQUIRE(world > 1 && factors_ready, "comm_prepare_local: needs comm_init_local and allocated factors");
    comm_ipc_close();
    enable_peer_access(devices);
    xchg_ne_max = KP * KP;
    xbuf.ensure(static_cast<size_t>(2) * world * xchg_ne_max + kXchgTailWords);
    B200_CUDA_CHECK(cudaMemsetAsync(xbuf.ptr, 0, xbuf.bytes(), stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

void Engine::comm_attach_local(Engine* const* all) {
    B200_REQUIRE(world > 1 && xbuf.ptr != nullptr, "comm_attach_local: call comm_prepare_local on every engine first");
    for (int r = 0; r < world; ++r) {
        B200_REQUIRE(all[r] && all[r]->world == world && all[r]->rank == r && all[r]->KP == KP && all[r]->xbuf.ptr,
                     "comm_attach_local: engines do not form one group");
        peer_W[r] = all[r]->W_T.ptr;
        peer_H[r] = all[r]->H.ptr;
        peer_x[r] = all[r]->xbuf.ptr;
    }
    peers_local = true;
    peers_ready = true;
}

void Engine::comm_destroy() {
    comm_ipc_close();
    if (comm) {
        ncclCommDestroy(as_comm(comm));
        comm = nullptr;
    }
}

void Engine::allreduce_f64(double* buf, size_t count) {
    if (peers_ready && static_cast<int>(count) <= xchg_ne_max) {
        // one-shot all-reduce over peer memory (kernels_dense.cuh xchg_allreduce_kernel)
        XchgParams x{};
        for (int r = 0; r < world; ++r) x.peer[r] = peer_x[r];
        x.rank = rank; x.world = world; x.ne_max = xchg_ne_max;
        xchg_allreduce_kernel<<<1, 1024, 0, stream>>>(buf, static_cast<int>(count), buf, x, state.ptr);
        launches[RCPPML_B200_SEC_COMM] += 1;
        return;
    }
    B200_NCCL_CHECK(ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, as_comm(comm), stream));
}

// CUDA IPC handles of this rank's W_T, H and exchange buffer (3 x 64 bytes). Call after the factors exist.
void Engine::comm_ipc_export(char* handles192) {
    use_device();
    B200_REQUIRE(world > 1 && factors_ready, "comm_ipc_export: needs a communicator and allocated factors");
    comm_ipc_close();
    xchg_ne_max = KP * KP;
    const size_t doubles = static_cast<size_t>(2) * world * xchg_ne_max + kXchgTailWords;
    xbuf.ensure(doubles);
    B200_CUDA_CHECK(cudaMemsetAsync(xbuf.ptr, 0, xbuf.bytes(), stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    cudaIpcMemHandle_t hs[3];
    B200_CUDA_CHECK(cudaIpcGetMemHandle(&hs[0], W_T.ptr));
    B200_CUDA_CHECK(cudaIpcGetMemHandle(&hs[1], H.ptr));
    B200_CUDA_CHECK(cudaIpcGetMemHandle(&hs[2], xbuf.ptr));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handles192, hs, sizeof(hs));
}

void Engine::comm_ipc_import(const char* all) {
    use_device();
    B200_REQUIRE(world > 1 && world <= 8 && xbuf.ptr != nullptr, "comm_ipc_import: call comm_ipc_export first (world <= 8)");
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            peer_W[r] = W_T.ptr; peer_H[r] = H.ptr; peer_x[r] = xbuf.ptr;
            continue;
        }
        cudaIpcMemHandle_t hs[3];
        std::memcpy(hs, all + static_cast<size_t>(r) * 192, sizeof(hs));
        void* p[3] = {nullptr, nullptr, nullptr};
        for (int q = 0; q < 3; ++q)
            B200_CUDA_CHECK(cudaIpcOpenMemHandle(&p[q], hs[q], cudaIpcMemLazyEnablePeerAccess));
        peer_W[r] = static_cast<float*>(p[0]);
        peer_H[r] = static_cast<float*>(p[1]);
        peer_x[r] = static_cast<double*>(p[2]);
    }
    peers_ready = true;
}

// In-place all-gather of the row blocks of a replicated factor: rank r contributes rows [cuts[r], cuts[r+1]).
// Equal (padded) blocks go through ncclAllGather; an explicit partition through grouped per-rank broadcasts.
void Engine::allgather_rows(float* buf, const std::vector<int>& cuts, int rows_padded, int sec) {
    sec_begin(sec);
    if (equal_partition) {
        const size_t cnt = static_cast<size_t>(rows_padded / world) * KP;
        B200_NCCL_CHECK(ncclAllGather(buf + static_cast<size_t>(rank) * cnt, buf, cnt, ncclFloat, as_comm(comm), stream));
    } else {
        B200_NCCL_CHECK(ncclGroupStart());
        for (int r = 0; r < world; ++r) {
            const size_t cnt = static_cast<size_t>(cuts[r + 1] - cuts[r]) * KP;
            float* blk = buf + static_cast<size_t>(cuts[r]) * KP;
            if (cnt > 0) B200_NCCL_CHECK(ncclBroadcast(blk, blk, cnt, ncclFloat, r, as_comm(comm), stream));
        }
        B200_NCCL_CHECK(ncclGroupEnd());
    }
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
