// mc_probe.cu — can this box do NVSwitch multicast stores (multimem.st) into factor replicas?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/mc_probe tools/mc_probe.cu -lcuda
//   ./tools/mc_probe            (needs >= 2 GPUs)
// 1. CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED per device, multicast granularities;
// 2. one process, G devices: cuMulticastCreate / AddDevice / cuMemCreate / BindMem / map; a kernel on device 0 writes a
//    pattern with multimem.st through the multicast mapping; every device's local mapping must hold it; timing of a
//    32 MB multicast store vs G-1 unicast peer stores;
// 3. two processes (fork before CUDA is touched): the multicast object and the physical allocations exported as POSIX
//    file descriptors and duplicated into the other process with pidfd_getfd (no Unix-socket plumbing).
#include <cuda.h>
#include <cuda_runtime.h>

#include <sys/syscall.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CU(x)                                                                                   \
    do {                                                                                        \
        CUresult r_ = (x);                                                                      \
        if (r_ != CUDA_SUCCESS) {                                                               \
            const char* s_ = nullptr;                                                           \
            cuGetErrorString(r_, &s_);                                                          \
            std::printf("FAIL %s -> %d (%s) at line %d\n", #x, (int)r_, s_ ? s_ : "?", __LINE__); \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)
#define RT(x)                                                                                   \
    do {                                                                                        \
        cudaError_t r_ = (x);                                                                   \
        if (r_ != cudaSuccess) {                                                                \
            std::printf("FAIL %s -> %s at line %d\n", #x, cudaGetErrorString(r_), __LINE__);      \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

__global__ void mc_store_kernel(float4* mc, long long n4, float base) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float v = base + (float)(i & 1023);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + i), "f"(v), "f"(v + 1.f),
                     "f"(v + 2.f), "f"(v + 3.f)
                     : "memory");
    }
}
struct Peers { float4* p[8]; };
__global__ void uc_store_kernel(Peers peers, int np, long long n4, float base) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float v = base + (float)(i & 1023);
        const float4 x = make_float4(v, v + 1.f, v + 2.f, v + 3.f);
        for (int q = 0; q < np; ++q) peers.p[q][i] = x;
    }
}
__global__ void check_kernel(const float4* loc, long long n4, float base, unsigned long long* bad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float v = base + (float)(i & 1023);
        const float4 x = loc[i];
        if (x.x != v || x.y != v + 1.f || x.z != v + 2.f || x.w != v + 3.f) atomicAdd(bad, 1ULL);
    }
}

static size_t round_up(size_t a, size_t g) { return (a + g - 1) / g * g; }

static int in_process(int G) {
    std::printf("== in-process multicast over %d devices\n", G);
    const size_t want = 32u << 20;
    CUmulticastObjectProp mp{};
    mp.numDevices = G;
    mp.handleTypes = 0;
    mp.flags = 0;
    mp.size = want;
    size_t gmin = 0, grec = 0;
    CU(cuMulticastGetGranularity(&gmin, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
    CU(cuMulticastGetGranularity(&grec, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    std::printf("multicast granularity: min %zu, recommended %zu\n", gmin, grec);
    const size_t size = round_up(want, grec);
    mp.size = size;
    CUmemGenericAllocationHandle mc;
    CU(cuMulticastCreate(&mc, &mp));
    std::vector<CUmemGenericAllocationHandle> phys(G);
    std::vector<CUdeviceptr> loc(G), mcva(G);
    for (int g = 0; g < G; ++g) {
        RT(cudaSetDevice(g));
        RT(cudaFree(0));
        CUdevice dev;
        CU(cuDeviceGet(&dev, g));
        CU(cuMulticastAddDevice(mc, dev));
    }
    for (int g = 0; g < G; ++g) {
        RT(cudaSetDevice(g));
        CUmemAllocationProp ap{};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = g;
        size_t ag = 0;
        CU(cuMemGetAllocationGranularity(&ag, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
        if (g == 0) std::printf("allocation granularity (recommended): %zu\n", ag);
        CU(cuMemCreate(&phys[g], size, &ap, 0));
        CU(cuMemAddressReserve(&loc[g], size, grec, 0, 0));
        CU(cuMemMap(loc[g], size, 0, phys[g], 0));
        std::vector<CUmemAccessDesc> acc(G);
        for (int q = 0; q < G; ++q) {
            acc[q].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            acc[q].location.id = q;
            acc[q].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        }
        CU(cuMemSetAccess(loc[g], size, acc.data(), G));          // every device may read / write this replica (unicast)
        CU(cuMulticastBindMem(mc, 0, phys[g], 0, size, 0));
    }
    for (int g = 0; g < G; ++g) {
        RT(cudaSetDevice(g));
        CU(cuMemAddressReserve(&mcva[g], size, grec, 0, 0));
        CU(cuMemMap(mcva[g], size, 0, mc, 0));
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = g;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CU(cuMemSetAccess(mcva[g], size, &acc, 1));
        RT(cudaMemset((void*)loc[g], 0, size));
        RT(cudaDeviceSynchronize());
    }
    const long long n4 = (long long)(want / 16);
    RT(cudaSetDevice(0));
    cudaEvent_t e0, e1;
    RT(cudaEventCreate(&e0));
    RT(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep) {
        RT(cudaEventRecord(e0));
        mc_store_kernel<<<148 * 8, 256>>>((float4*)mcva[0], n4, 7.f + rep);
        RT(cudaEventRecord(e1));
        RT(cudaDeviceSynchronize());
        float ms = 0;
        RT(cudaEventElapsedTime(&ms, e0, e1));
        std::printf("multimem.st of %zu MB from device 0: %.3f ms (%.1f GB/s of payload)\n", want >> 20, ms, want / ms / 1e6);
    }
    unsigned long long* bad;
    RT(cudaMallocManaged(&bad, sizeof(*bad)));
    for (int g = 0; g < G; ++g) {
        RT(cudaSetDevice(g));
        *bad = 0;
        check_kernel<<<148 * 4, 256>>>((const float4*)loc[g], n4, 9.f, bad);
        RT(cudaDeviceSynchronize());
        std::printf("device %d local replica after the multicast store: %llu mismatching words %s\n", g, *bad, *bad ? "FAIL" : "ok");
    }
    RT(cudaSetDevice(0));
    Peers peers{};
    for (int g = 1; g < G; ++g) peers.p[g - 1] = (float4*)loc[g];
    for (int rep = 0; rep < 3; ++rep) {
        RT(cudaEventRecord(e0));
        uc_store_kernel<<<148 * 8, 256>>>(peers, G - 1, n4, 20.f + rep);
        RT(cudaEventRecord(e1));
        RT(cudaDeviceSynchronize());
        float ms = 0;
        RT(cudaEventElapsedTime(&ms, e0, e1));
        std::printf("unicast stores of %zu MB to %d peers from device 0: %.3f ms (%.1f GB/s egress)\n", want >> 20, G - 1, ms,
                    (double)want * (G - 1) / ms / 1e6);
    }
    std::printf("in-process multicast: OK\n");
    return 0;
}

// Two processes: parent = rank 0 (device 0), child = rank 1 (device 1). pipes carry pid / fd numbers and go-ahead bytes.
static int cross_process() {
    std::printf("== cross-process multicast (POSIX fd handles duplicated with pidfd_getfd)\n");
    std::fflush(stdout);
    int p2c[2], c2p[2];
    if (pipe(p2c) || pipe(c2p)) return 1;
    const pid_t child = fork();
    if (child < 0) return 1;
    const int rank = child == 0 ? 1 : 0;
    const int rd = rank == 0 ? c2p[0] : p2c[0], wr = rank == 0 ? p2c[1] : c2p[1];
    auto send = [&](const void* p, size_t n) { return write(wr, p, n) == (ssize_t)n; };
    auto recv = [&](void* p, size_t n) { return read(rd, p, n) == (ssize_t)n; };
    auto dup_from = [&](pid_t pid, int fd) -> int {
        const int pfd = (int)syscall(SYS_pidfd_open, pid, 0);
        if (pfd < 0) { std::perror("pidfd_open"); return -1; }
        const int got = (int)syscall(SYS_pidfd_getfd, pfd, fd, 0);
        if (got < 0) std::perror("pidfd_getfd");
        close(pfd);
        return got;
    };
    auto body = [&]() -> int {
        CU(cuInit(0));
        RT(cudaSetDevice(rank));
        RT(cudaFree(0));
        CUdevice dev;
        CU(cuDeviceGet(&dev, rank));
        const size_t want = 32u << 20;
        CUmulticastObjectProp mp{};
        mp.numDevices = 2;
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        mp.size = want;
        size_t grec = 0;
        CU(cuMulticastGetGranularity(&grec, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
        const size_t size = round_up(want, grec);
        mp.size = size;
        CUmemGenericAllocationHandle mc;
        const pid_t me = getpid();
        pid_t other = 0;
        if (rank == 0) {
            CU(cuMulticastCreate(&mc, &mp));
            int fd = -1;
            CU(cuMemExportToShareableHandle(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
            if (!send(&me, sizeof(me)) || !send(&fd, sizeof(fd))) return 1;
            if (!recv(&other, sizeof(other))) return 1;
        } else {
            int fd = -1;
            if (!recv(&other, sizeof(other)) || !recv(&fd, sizeof(fd))) return 1;
            if (!send(&me, sizeof(me))) return 1;
            const int mine = dup_from(other, fd);
            if (mine < 0) { std::printf("FAIL pidfd_getfd\n"); return 1; }
            CU(cuMemImportFromShareableHandle(&mc, (void*)(uintptr_t)mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
            close(mine);
        }
        CU(cuMulticastAddDevice(mc, dev));
        // local physical memory, shareable; mapped locally; bound to the multicast object
        CUmemAllocationProp ap{};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = rank;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CUmemGenericAllocationHandle phys;
        CU(cuMemCreate(&phys, size, &ap, 0));
        CUdeviceptr loc = 0, mcva = 0, peer = 0;
        CU(cuMemAddressReserve(&loc, size, grec, 0, 0));
        CU(cuMemMap(loc, size, 0, phys, 0));
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = rank;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CU(cuMemSetAccess(loc, size, &acc, 1));
        RT(cudaMemset((void*)loc, 0, size));
        RT(cudaDeviceSynchronize());
        // exchange the physical handles too (the peer's replica mapped here: unicast reads / writes as with cudaIpc)
        int myfd = -1, theirfd = -1;
        CU(cuMemExportToShareableHandle(&myfd, phys, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        if (!send(&myfd, sizeof(myfd)) || !recv(&theirfd, sizeof(theirfd))) return 1;
        const int dupfd = dup_from(other, theirfd);
        if (dupfd < 0) return 1;
        CUmemGenericAllocationHandle pphys;
        CU(cuMemImportFromShareableHandle(&pphys, (void*)(uintptr_t)dupfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
        close(dupfd);
        CU(cuMemAddressReserve(&peer, size, grec, 0, 0));
        CU(cuMemMap(peer, size, 0, pphys, 0));
        CU(cuMemSetAccess(peer, size, &acc, 1));
        CU(cuMulticastBindMem(mc, 0, phys, 0, size, 0));          // blocks until both devices were added
        CU(cuMemAddressReserve(&mcva, size, grec, 0, 0));
        CU(cuMemMap(mcva, size, 0, mc, 0));
        CU(cuMemSetAccess(mcva, size, &acc, 1));
        char go = 1;
        if (!send(&go, 1) || !recv(&go, 1)) return 1;             // both sides bound and mapped
        const long long n4 = (long long)(want / 16);
        if (rank == 0) {
            mc_store_kernel<<<148 * 8, 256>>>((float4*)mcva, n4, 3.f);
            RT(cudaDeviceSynchronize());
        }
        if (!send(&go, 1) || !recv(&go, 1)) return 1;             // store done
        unsigned long long* bad;
        RT(cudaMallocManaged(&bad, sizeof(*bad)));
        *bad = 0;
        check_kernel<<<148 * 4, 256>>>((const float4*)loc, n4, 3.f, bad);
        RT(cudaDeviceSynchronize());
        std::printf("rank %d: local replica after rank 0's multicast store: %llu mismatching words %s\n", rank, *bad, *bad ? "FAIL" : "ok");
        *bad = 0;
        check_kernel<<<148 * 4, 256>>>((const float4*)peer, n4, 3.f, bad);
        RT(cudaDeviceSynchronize());
        std::printf("rank %d: the peer's replica read through the imported mapping: %llu mismatching words %s\n", rank, *bad, *bad ? "FAIL" : "ok");
        std::fflush(stdout);
        if (!send(&go, 1) || !recv(&go, 1)) return 1;
        return 0;
    };
    const int rc = body();
    if (rank == 1) _exit(rc);
    int st = 0;
    waitpid(child, &st, 0);
    std::printf("cross-process multicast: parent rc=%d child rc=%d\n", rc, WIFEXITED(st) ? WEXITSTATUS(st) : -1);
    return rc || !WIFEXITED(st) || WEXITSTATUS(st);
}

int main(int argc, char** argv) {
    const bool only_cross = argc > 1 && !std::strcmp(argv[1], "cross");
    if (!only_cross) {
        // the cross-process test forks BEFORE this process touches CUDA: run it in a child of its own first
        const pid_t c = fork();
        if (c == 0) {
            execl(argv[0], argv[0], "cross", (char*)nullptr);
            _exit(127);
        }
        int st = 0;
        waitpid(c, &st, 0);
        std::printf("(cross-process test exit status %d)\n", WIFEXITED(st) ? WEXITSTATUS(st) : -1);
    } else {
        return cross_process();
    }
    CU(cuInit(0));
    int n = 0;
    RT(cudaGetDeviceCount(&n));
    for (int g = 0; g < n; ++g) {
        CUdevice dev;
        CU(cuDeviceGet(&dev, g));
        int mc = 0, vmm = 0, fab = 0;
        CU(cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
        CU(cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev));
        CU(cuDeviceGetAttribute(&fab, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev));
        std::printf("device %d: multicast %d, vmm %d, fabric handles %d\n", g, mc, vmm, fab);
    }
    if (n < 2) { std::printf("needs >= 2 GPUs for the multicast tests\n"); return 0; }
    return in_process(n > 8 ? 8 : n);
}
