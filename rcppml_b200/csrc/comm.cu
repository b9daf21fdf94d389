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

#include <sys/prctl.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <vector>

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
    mc_decide();
    if (world == 1) return;
    ncclUniqueId id;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    B200_NCCL_CHECK(ncclCommInitRank(&c, world, id, rank));
    comm = reinterpret_cast<ncclComm*>(c);
}

void Engine::comm_ipc_close() {
    cudaSetDevice(device);
    if (peers_ready || mc_ready) {
        cudaStreamSynchronize(stream);
        cudaStreamSynchronize(side_stream);
    }
    mc_close();
    for (int r = 0; r < 8; ++r) {                            // imported VMM mappings (also of an import that failed half-way)
        if (peer_map_W[r].va || peer_map_W[r].handle) peer_map_W[r].release();
        if (peer_map_H[r].va || peer_map_H[r].handle) peer_map_H[r].release();
    }
    for (int r = 0; r < world && peers_ready && !peers_local; ++r) {
        if (r == rank) continue;
        if (!peers_vmm) {
            cudaIpcCloseMemHandle(peer_W[r]);
            cudaIpcCloseMemHandle(peer_H[r]);
        }
        cudaIpcCloseMemHandle(peer_x[r]);
    }
    comm_mc_finish();
    peers_ready = false;
    peers_local = false;
    peers_vmm = false;
}

// ---- NVSwitch multicast replication of the factors (vmm.hpp) ---------------------------------------------------
void Engine::mc_decide() {
    mc_wanted = false;
    if (world <= 1) return;
    const char* env = std::getenv("RCPPML_B200_MC");
    if (env && env[0] == '0') return;
    mc_mode = (env && env[0] == '1') ? 1 : 2;
    if (!multicast_supported(device)) return;
    try {
        const size_t g = multicast_granularity(world);
        W_T.granularity = H.granularity = g;
        W_T.device = H.device = device;
        mc_wanted = g > 0;
    } catch (...) {
        mc_wanted = false;
    }
}

float* Engine::mc_alias(const float* replica_ptr) const {
    if (!mc_ready) return nullptr;
    if (replica_ptr >= W_T.ptr && replica_ptr < W_T.ptr + W_T.count) return mcW.fptr() + (replica_ptr - W_T.ptr);
    if (replica_ptr >= H.ptr && replica_ptr < H.ptr + H.count) return mcH.fptr() + (replica_ptr - H.ptr);
    return nullptr;
}

// In-process peers read / write each other's replicas through plain pointers: a VMM allocation is accessible from a
// device only if access was granted on it (cudaDeviceEnablePeerAccess covers cudaMalloc memory only).
void Engine::mc_grant_local_access(const int* devices) {
    if (!W_T.vmm || !H.vmm) return;
    const DriverApi& drv = DriverApi::get();
    std::vector<CUmemAccessDesc> acc(world);
    for (int r = 0; r < world; ++r) {
        acc[r].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc[r].location.id = devices[r];
        acc[r].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    }
    B200_CU_CHECK(drv.MemSetAccess(reinterpret_cast<CUdeviceptr>(W_T.ptr), W_T.phys_size, acc.data(), acc.size()));
    B200_CU_CHECK(drv.MemSetAccess(reinterpret_cast<CUdeviceptr>(H.ptr), H.phys_size, acc.data(), acc.size()));
}

void Engine::mc_create(CUmemGenericAllocationHandle* hW, CUmemGenericAllocationHandle* hH, bool shareable) {
    use_device();
    B200_REQUIRE(W_T.vmm && H.vmm, "mc_create: the factors are not VMM allocations");
    const DriverApi& drv = DriverApi::get();
    CUmulticastObjectProp mp{};
    mp.numDevices = static_cast<unsigned>(world);
    mp.handleTypes = shareable ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : 0;
    mp.size = W_T.phys_size;
    B200_CU_CHECK(drv.MulticastCreate(hW, &mp));
    mp.size = H.phys_size;
    const CUresult r = drv.MulticastCreate(hH, &mp);
    if (r != CUDA_SUCCESS) { drv.MemRelease(*hW); B200_CU_CHECK(r); }
}

void Engine::mc_add_device(CUmemGenericAllocationHandle hW, CUmemGenericAllocationHandle hH) {
    use_device();
    const DriverApi& drv = DriverApi::get();
    CUdevice dev;
    B200_CU_CHECK(drv.DeviceGet(&dev, device));
    B200_CU_CHECK(drv.MulticastAddDevice(hW, dev));
    B200_CU_CHECK(drv.MulticastAddDevice(hH, dev));
}

// After EVERY rank added its device: bind this rank's replicas at offset 0 and map the multicast objects here.
void Engine::mc_bind_and_map(CUmemGenericAllocationHandle hW, CUmemGenericAllocationHandle hH, bool owner) {
    use_device();
    const DriverApi& drv = DriverApi::get();
    mc_owner = owner;
    mc_local = peers_local;
    mcW.handle = hW; mcW.size = W_T.phys_size; mcW.device = device;
    mcH.handle = hH; mcH.size = H.phys_size; mcH.device = device;
    B200_CU_CHECK(drv.MulticastBindMem(hW, 0, W_T.phys, 0, W_T.phys_size, 0));
    mcW.bound = true;
    B200_CU_CHECK(drv.MulticastBindMem(hH, 0, H.phys, 0, H.phys_size, 0));
    mcH.bound = true;
    mcW.map(hW, W_T.phys_size, W_T.granularity, device);
    mcH.map(hH, H.phys_size, H.granularity, device);
    mcW.bound = mcH.bound = true;                            // (map() keeps handle / size / device)
    mc_ready = true;
}

void Engine::mc_close() {
    if (!mcW.handle && !mcH.handle) { mc_ready = false; return; }
    cudaSetDevice(device);
    // in-process groups share one handle per factor, owned by the orchestrator (abi_reference.cu): an engine only
    // unmaps and unbinds. Cross-process: every rank holds its own (created or imported) handle and releases it.
    mcW.release(!mc_local);
    mcH.release(!mc_local);
    mc_ready = false;
    mc_owner = false;
}

// ---- cross-process multicast set-up (one process per GPU) ---------------------------------------------------------
// Blob (128 bytes per rank): {int pid, fd_W, fd_H, fd_mcW, fd_mcH, pad; size_t size_W, size_H; cudaIpcMemHandle_t xbuf}.
// The file descriptors are POSIX-fd exports of the physical allocations (every rank) and of the two multicast objects
// (rank 0); a peer duplicates them into its own process with pidfd_open + pidfd_getfd (same user; no socket plumbing).
namespace {
struct McBlob {
    int pid, fd_W, fd_H, fd_mcW, fd_mcH, pad;
    unsigned long long size_W, size_H;
    cudaIpcMemHandle_t xbuf;
};
static_assert(sizeof(McBlob) <= 128, "McBlob must fit the 128-byte wire slot");

int dup_fd_from(int pid, int fd) {
    const int pfd = static_cast<int>(syscall(SYS_pidfd_open, pid, 0));
    B200_REQUIRE(pfd >= 0, "pidfd_open failed (multicast set-up needs Linux >= 5.6 and same-user processes)");
    const int got = static_cast<int>(syscall(SYS_pidfd_getfd, pfd, fd, 0));
    close(pfd);
    B200_REQUIRE(got >= 0, "pidfd_getfd failed (ptrace permission between the ranks' processes)");
    return got;
}
}  // namespace

void Engine::comm_mc_export(char* blob128) {
    use_device();
    B200_REQUIRE(world > 1 && factors_ready, "comm_mc_export: needs a communicator and allocated factors");
    B200_REQUIRE(mc_wanted && W_T.vmm && H.vmm, "comm_mc_export: multicast is not available for this engine");
    comm_ipc_close();
    const DriverApi& drv = DriverApi::get();
    xchg_ne_max = KP * KP;
    xbuf.ensure(static_cast<size_t>(2) * world * xchg_ne_max + kXchgTailWords);
    B200_CUDA_CHECK(cudaMemsetAsync(xbuf.ptr, 0, xbuf.bytes(), stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    McBlob b{};
    b.pid = static_cast<int>(getpid());
    b.fd_W = b.fd_H = b.fd_mcW = b.fd_mcH = -1;
    // Yama (ptrace_scope 1) would refuse pidfd_getfd between sibling processes: open the window for the duration of the
    // hand-over only — comm_mc_finish closes it again together with the exported descriptors
    prctl(PR_SET_PTRACER, PR_SET_PTRACER_ANY, 0, 0, 0);
    mc_ptracer_window = true;
    B200_CU_CHECK(drv.MemExportToShareableHandle(&b.fd_W, W_T.phys, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    B200_CU_CHECK(drv.MemExportToShareableHandle(&b.fd_H, H.phys, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    mc_export_fds[0] = b.fd_W; mc_export_fds[1] = b.fd_H;
    if (rank == 0) {
        CUmemGenericAllocationHandle hW = 0, hH = 0;
        mc_create(&hW, &hH, true);
        mcW.handle = hW; mcH.handle = hH;                        // kept here until import maps them
        B200_CU_CHECK(drv.MemExportToShareableHandle(&b.fd_mcW, hW, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        B200_CU_CHECK(drv.MemExportToShareableHandle(&b.fd_mcH, hH, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        mc_export_fds[2] = b.fd_mcW; mc_export_fds[3] = b.fd_mcH;
    }
    b.size_W = W_T.phys_size;
    b.size_H = H.phys_size;
    B200_CUDA_CHECK(cudaIpcGetMemHandle(&b.xbuf, xbuf.ptr));
    std::memset(blob128, 0, 128);
    std::memcpy(blob128, &b, sizeof(b));
}

void Engine::comm_mc_import(const char* all) {
    use_device();
    B200_REQUIRE(world > 1 && world <= 8 && xbuf.ptr != nullptr, "comm_mc_import: call comm_mc_export first (world <= 8)");
    // (test hook: exercise the fall-back of the host launcher — VMM-backed factors on the NCCL loop)
    B200_REQUIRE(!std::getenv("RCPPML_B200_MC_TEST_FAIL"), "comm_mc_import: failure requested by RCPPML_B200_MC_TEST_FAIL");
    const DriverApi& drv = DriverApi::get();
    std::vector<McBlob> blobs(world);
    for (int r = 0; r < world; ++r) std::memcpy(&blobs[r], all + static_cast<size_t>(r) * 128, sizeof(McBlob));
    for (int r = 0; r < world; ++r)
        B200_REQUIRE(blobs[r].size_W == W_T.phys_size && blobs[r].size_H == H.phys_size, "comm_mc_import: ranks disagree on the factor sizes");
    auto import_fd = [&](int pid, int fd) {
        const int mine = dup_fd_from(pid, fd);
        CUmemGenericAllocationHandle h = 0;
        const CUresult rc = drv.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(mine)),
                                                             CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(mine);
        B200_CU_CHECK(rc);
        return h;
    };
    // the peers' replicas, mapped here (unicast reads / writes: initial all-gather of blocks, diagnostics) + exchange buffers
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            peer_W[r] = W_T.ptr; peer_H[r] = H.ptr; peer_x[r] = xbuf.ptr;
            continue;
        }
        peer_map_W[r].map(import_fd(blobs[r].pid, blobs[r].fd_W), W_T.phys_size, W_T.granularity, device);
        peer_map_H[r].map(import_fd(blobs[r].pid, blobs[r].fd_H), H.phys_size, H.granularity, device);
        peer_W[r] = peer_map_W[r].fptr();
        peer_H[r] = peer_map_H[r].fptr();
        void* px = nullptr;
        B200_CUDA_CHECK(cudaIpcOpenMemHandle(&px, blobs[r].xbuf, cudaIpcMemLazyEnablePeerAccess));
        peer_x[r] = static_cast<double*>(px);
    }
    peers_ready = true;
    peers_local = false;
    peers_vmm = true;
    // the multicast objects: rank 0 created them; everybody adds its device, binds its replicas (blocks until all
    // devices were added) and maps the objects
    CUmemGenericAllocationHandle hW = mcW.handle, hH = mcH.handle;
    if (rank != 0) {
        hW = import_fd(blobs[0].pid, blobs[0].fd_mcW);
        hH = import_fd(blobs[0].pid, blobs[0].fd_mcH);
    }
    mcW.handle = hW; mcH.handle = hH;
    mc_add_device(hW, hH);
}

// Second half of the import, after EVERY rank's comm_mc_import succeeded (the host launcher checks): binding blocks
// until all devices joined the multicast objects, so it must not start while a rank may still fail before joining.
void Engine::comm_mc_bind() {
    use_device();
    B200_REQUIRE(peers_ready && peers_vmm && mcW.handle && mcH.handle, "comm_mc_bind: call comm_mc_import first");
    mc_bind_and_map(mcW.handle, mcH.handle, rank == 0);
}

// Fall-back of the host launcher when the multicast set-up failed on some rank: give up multicast for this engine but
// KEEP the factors — they move (device to device) from the VMM allocations into plain cudaMalloc buffers, which the
// CUDA-IPC path (unicast peer stores) can export. Every rank must call it (same decision on all ranks).
void Engine::mc_disable_keep_factors() {
    use_device();
    comm_ipc_close();
    mc_wanted = false;
    if (!factors_ready) return;
    auto move = [&](FactorBuffer& buf) {
        if (!buf.vmm || !buf.ptr) return;
        const size_t n_floats = buf.count;
        float* plain = nullptr;
        B200_CUDA_CHECK(cudaMalloc(&plain, n_floats * sizeof(float)));
        B200_CUDA_CHECK(cudaMemcpyAsync(plain, buf.ptr, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        buf.want_vmm = false;
        buf.release();                                       // unmaps and frees the VMM allocation
        buf.ptr = plain;
        buf.count = n_floats;
    };
    drop_iteration_graph();
    move(W_T);
    move(H);
}

// After every rank imported (a barrier of the host launcher): the exported descriptors are no longer needed.
void Engine::comm_mc_finish() {
    for (int& fd : mc_export_fds) {
        if (fd >= 0) close(fd);
        fd = -1;
    }
    if (mc_ptracer_window) {
        prctl(PR_SET_PTRACER, 0, 0, 0, 0);
        mc_ptracer_window = false;
    }
}

// ---- in-process multi-GPU (abi_reference.cu: RCPPML_NUM_GPUS) ------------------------------------------------
void Engine::comm_init_local(int rank_, int world_) {
    use_device();
    B200_REQUIRE(world_ >= 1 && world_ <= 8 && rank_ >= 0 && rank_ < world_, "comm_init_local: bad rank/world");
    B200_REQUIRE(!matrix_ready && !factors_ready, "comm_init_local must come first: blocks and padding depend on (rank, world)");
    rank = rank_;
    world = world_;
    mc_decide();
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
    // (a cached in-process group keeps its peer pointers and multicast mappings from call to call: alloc_factors
    // closes them when a buffer moved, comm_attach_local refreshes the pointers)
    if (!peers_local) comm_ipc_close();
    enable_peer_access(devices);
    mc_grant_local_access(devices);
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

int rcppml_b200_comm_mc_wanted(rcppml_b200_engine* e) { return e->impl.mc_wanted ? 1 : 0; }
int rcppml_b200_comm_mc_ready(rcppml_b200_engine* e) { return e->impl.mc_ready ? 1 : 0; }
int rcppml_b200_comm_mc_export(rcppml_b200_engine* e, char* blob128) {
    try { e->impl.comm_mc_export(blob128); return 0; }
    catch (const std::exception& ex) { b200::g_last_error = ex.what(); return -1; }
}
int rcppml_b200_comm_mc_import(rcppml_b200_engine* e, const char* all_blobs) {
    try { e->impl.comm_mc_import(all_blobs); return 0; }
    catch (const std::exception& ex) { b200::g_last_error = ex.what(); return -1; }
}
int rcppml_b200_comm_mc_bind(rcppml_b200_engine* e) {
    try { e->impl.comm_mc_bind(); return 0; }
    catch (const std::exception& ex) { b200::g_last_error = ex.what(); return -1; }
}
int rcppml_b200_comm_mc_disable(rcppml_b200_engine* e) {
    try { e->impl.mc_disable_keep_factors(); return 0; }
    catch (const std::exception& ex) { b200::g_last_error = ex.what(); return -1; }
}
int rcppml_b200_comm_p2p_close(rcppml_b200_engine* e) {
    try { e->impl.comm_ipc_close(); return 0; }
    catch (const std::exception& ex) { b200::g_last_error = ex.what(); return -1; }
}
int rcppml_b200_comm_mc_finish(rcppml_b200_engine* e) {
    try { e->impl.comm_mc_finish(); return 0; }
    catch (const std::exception& ex) { b200::g_last_error = ex.what(); return -1; }
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
