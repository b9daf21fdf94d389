// engine.cu — host side of the sm_100a ALS engine: device memory, the iteration schedule of
// nmf_fit (nmf/fit_cpu.hpp:444-1825) as a stream of kernels with NO per-iteration host
// synchronisation (convergence bookkeeping lives in DevState on the device), and the
// rcppml_b200_* C ABI declared in include/rcppml_gpu.h.
#include "engine.hpp"

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges are no-ops unless a profiler injects itself

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace b200 {

thread_local std::string g_last_error;

// Opt-in shared memory (cudaFuncSetAttribute) and the occupancy of a kernel are properties of the (kernel, DEVICE)
// pair, so they are cached per device — a host thread may drive engines on several devices (zero-copy entry,
// Engine(1) after Engine(0)), and the in-process multi-GPU path runs one thread per device.
constexpr int kMaxDevices = 64;
struct OccCache {
    std::atomic<int> v[kMaxDevices];
    OccCache() { for (auto& x : v) x.store(-1); }
};
template <class Kern>
static int cached_occupancy(OccCache& cache, Kern kern, int threads, size_t smem, const char* what) {
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    B200_REQUIRE(dev >= 0 && dev < kMaxDevices, "device ordinal out of range");
    int occ = cache.v[dev].load(std::memory_order_acquire);
    if (occ < 0) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
        B200_REQUIRE(occ > 0, what);
        cache.v[dev].store(occ, std::memory_order_release);
    }
    return occ;
}

// ---------------------------------------------------------------------------------------------
// kernel dispatch by padded rank
// ---------------------------------------------------------------------------------------------
template <int LANES, int NV, int SOLVER, int BSRC, int OUT>
static void launch_half_step_t(const HalfStepParams& p, int num_sms, cudaStream_t stream, int* grid_out) {
    auto kern = half_step_kernel<LANES, NV, SOLVER, BSRC, OUT>;
    const size_t smem = half_step_smem_bytes<LANES, NV, SOLVER, OUT>();
    static OccCache cache;
    const int cached_occ = cached_occupancy(cache, kern, 256, smem, "half_step_kernel does not fit on an SM");
    // Persistent grid: a multiple of the SM count (148 on B200) x resident CTAs per SM.
    const int grid = num_sms * cached_occ;
    if (grid_out) { *grid_out = grid; return; }
    kern<<<grid, 256, smem, stream>>>(p);
}

template <int SOLVER, int BSRC, int OUT>
static void launch_half_step_l(int lanes, const HalfStepParams& p, int num_sms, cudaStream_t s, int* grid_out) {
    // `lanes` encodes (LANES, NV) as LANES + 100*(NV-1): a lane owns NV 128-bit words of a factor row
    switch (lanes) {
        case 4: launch_half_step_t<4, 1, SOLVER, BSRC, OUT>(p, num_sms, s, grid_out); break;
        case 8: launch_half_step_t<8, 1, SOLVER, BSRC, OUT>(p, num_sms, s, grid_out); break;
        case 16: launch_half_step_t<16, 1, SOLVER, BSRC, OUT>(p, num_sms, s, grid_out); break;
        case 32: launch_half_step_t<32, 1, SOLVER, BSRC, OUT>(p, num_sms, s, grid_out); break;
        case 108: launch_half_step_t<8, 2, SOLVER, BSRC, OUT>(p, num_sms, s, grid_out); break;
        case 116: launch_half_step_t<16, 2, SOLVER, BSRC, OUT>(p, num_sms, s, grid_out); break;
        case 304: launch_half_step_t<4, 4, SOLVER, BSRC, OUT>(p, num_sms, s, grid_out); break;
        case 308: launch_half_step_t<8, 4, SOLVER, BSRC, OUT>(p, num_sms, s, grid_out); break;
        default: throw std::runtime_error("unsupported lane-group geometry");
    }
}

// grid_out != nullptr: only report the grid this configuration would use (for buffer sizing).
void launch_half_step(int lanes, int solver, int bsrc, int out, const HalfStepParams& p, int num_sms,
                      cudaStream_t s, int* grid_out) {
    B200_REQUIRE(bsrc == BSRC_GATHER, "right-hand sides are always gathered in-kernel");
    if (out == OUT_RHS) {
        launch_half_step_l<SOLVER_CD, BSRC_GATHER, OUT_RHS>(lanes, p, num_sms, s, grid_out);
    } else {
        if (solver == SOLVER_CD) launch_half_step_l<SOLVER_CD, BSRC_GATHER, OUT_SOLVE>(lanes, p, num_sms, s, grid_out);
        else launch_half_step_l<SOLVER_CHOL, BSRC_GATHER, OUT_SOLVE>(lanes, p, num_sms, s, grid_out);
    }
}

// The coordinate-descent kernel (kernels_cd.cuh): narrow lane groups, blocked pivots.
template <int LANES, int NV>
static void launch_cd_t(const HalfStepParams& p, int num_sms, cudaStream_t stream, int* grid_out) {
    auto kern = cd_half_step_kernel<LANES, NV>;
    const size_t smem = cd_half_step_smem_bytes<LANES, NV>();
    static OccCache cache;
    const int cached_occ = cached_occupancy(cache, kern, 256, smem, "cd_half_step_kernel does not fit on an SM");
    const int grid = num_sms * cached_occ;
    if (grid_out) { *grid_out = grid; return; }
    kern<<<grid, 256, smem, stream>>>(p);
}

// `geom` encodes (LANES, NV) as LANES + 100*(NV-1), LANES*4*NV == KP.
void launch_cd_half_step(int geom, const HalfStepParams& p, int num_sms, cudaStream_t s, int* grid_out) {
    switch (geom) {
        case 301: launch_cd_t<1, 4>(p, num_sms, s, grid_out); break;     // KP = 16
        case 102: launch_cd_t<2, 2>(p, num_sms, s, grid_out); break;
        case 4: launch_cd_t<4, 1>(p, num_sms, s, grid_out); break;
        case 701: launch_cd_t<1, 8>(p, num_sms, s, grid_out); break;     // KP = 32
        case 302: launch_cd_t<2, 4>(p, num_sms, s, grid_out); break;
        case 104: launch_cd_t<4, 2>(p, num_sms, s, grid_out); break;
        case 702: launch_cd_t<2, 8>(p, num_sms, s, grid_out); break;     // KP = 64
        case 304: launch_cd_t<4, 4>(p, num_sms, s, grid_out); break;
        case 108: launch_cd_t<8, 2>(p, num_sms, s, grid_out); break;
        case 704: launch_cd_t<4, 8>(p, num_sms, s, grid_out); break;     // KP = 128
        case 308: launch_cd_t<8, 4>(p, num_sms, s, grid_out); break;
        case 116: launch_cd_t<16, 2>(p, num_sms, s, grid_out); break;
        default: throw std::runtime_error("unsupported coordinate-descent lane-group geometry");
    }
}

template <int GL, int GNV, int SL, int SNV, int SOLVER, int WARPS = 8, bool HYB = false>
static void launch_tiled_t(const HalfStepParams& p, int num_sms, cudaStream_t stream, int* grid_out) {
    auto kern = tiled_half_step_kernel<GL, GNV, SL, SNV, SOLVER, WARPS, HYB>;
    const size_t smem = tiled_smem_bytes<GL, GNV, SL, SNV, SOLVER, WARPS, HYB>();
    static OccCache cache;
    const int cached_occ = cached_occupancy(cache, kern, WARPS * 32, smem, "tiled_half_step_kernel does not fit on an SM");
    const int grid = num_sms * cached_occ;
    if (grid_out) { *grid_out = grid; return; }
    kern<<<grid, WARPS * 32, smem, stream>>>(p);
}

template <int SOLVER>
static void launch_tiled_s(int gather_geom, const HalfStepParams& p, int num_sms, cudaStream_t s, int* grid_out) {
    // gather_geom = LANES + 100*(NV-1) of the gather phase (+ 1000 * solve lanes when the solve geometry is not the
    // default narrowest one: k = 64 with 4 lanes x 4 words = batches of 8 columns instead of 16 — fewer columns per
    // warp batch quantise better when a rank holds few columns, Engine::tiled_solve_lanes)
    switch (gather_geom) {
        case 8108: launch_tiled_t<8, 2, 8, 2, SOLVER>(p, num_sms, s, grid_out); return;   // KP = 64, 4-column batches (experiment)
        case 4016: launch_tiled_t<16, 1, 4, 4, SOLVER>(p, num_sms, s, grid_out); return;  // KP = 64, 8-column batches
        case 4108: launch_tiled_t<8, 2, 4, 4, SOLVER>(p, num_sms, s, grid_out); return;
        // + 10000: one 768-thread CTA per SM (L / Lᵀ once per SM); + 20000: the same with the hybrid register +
        // cp.async-ring gather (kernels_tiled.cuh) — Cholesky, short columns; Engine::tiled_gather_geom
        case 14108: if (SOLVER == SOLVER_CHOL) { launch_tiled_t<8, 2, 4, 4, SOLVER_CHOL, 24, false>(p, num_sms, s, grid_out); return; } break;
        case 24108: if (SOLVER == SOLVER_CHOL) { launch_tiled_t<8, 2, 4, 4, SOLVER_CHOL, 24, true>(p, num_sms, s, grid_out); return; } break;
        default: break;
    }
    switch (gather_geom) {                          // gather (LANES + 100*(NV-1)); the solve geometry follows from KP
        case 4: launch_tiled_t<4, 1, 1, 4, SOLVER>(p, num_sms, s, grid_out); break;       // KP = 16
        case 8: launch_tiled_t<8, 1, 1, 8, SOLVER>(p, num_sms, s, grid_out); break;       // KP = 32
        case 16: launch_tiled_t<16, 1, 2, 8, SOLVER>(p, num_sms, s, grid_out); break;     // KP = 64
        case 108: launch_tiled_t<8, 2, 2, 8, SOLVER>(p, num_sms, s, grid_out); break;
        case 32: launch_tiled_t<32, 1, 4, 8, SOLVER>(p, num_sms, s, grid_out); break;     // KP = 128
        case 116: launch_tiled_t<16, 2, 4, 8, SOLVER>(p, num_sms, s, grid_out); break;
        default: throw std::runtime_error("unsupported gather geometry for the tiled kernel");
    }
}

void launch_tiled_half_step(int gather_geom, int solver, const HalfStepParams& p, int num_sms, cudaStream_t s, int* grid_out) {
    if (solver == SOLVER_CD) launch_tiled_s<SOLVER_CD>(gather_geom, p, num_sms, s, grid_out);
    else launch_tiled_s<SOLVER_CHOL>(gather_geom, p, num_sms, s, grid_out);
}

static bool cd_geometry_matches(int geom, int KP) {
    const int lanes = geom % 100, nv = geom / 100 + 1;
    switch (geom) {
        case 301: case 102: case 4: case 701: case 302: case 104: case 702: case 304: case 108: case 704: case 308: case 116:
            return lanes * 4 * nv == KP;
        default: return false;
    }
}

static void launch_normalize_gram(int KP, float* X, long long ncols, const float* d, int normalize, double* partials,
                                  const int* stop, int grid, cudaStream_t s, float* mcX) {
#ifdef B200_GRAM_DFMA      // register-tiled DFMA version (kept for comparison; bound by shared-memory delivery)
    switch (KP) {
        case 16: normalize_gram_kernel<16, 64><<<grid, kGramThreads, 0, s>>>(X, ncols, d, normalize, partials, stop, mcX); break;
        case 32: normalize_gram_kernel<32, 64><<<grid, kGramThreads, 0, s>>>(X, ncols, d, normalize, partials, stop, mcX); break;
        case 64: normalize_gram_kernel<64, 32><<<grid, kGramThreads, 0, s>>>(X, ncols, d, normalize, partials, stop, mcX); break;
        case 128: normalize_gram_kernel<128, 32><<<grid, kGramThreads, 0, s>>>(X, ncols, d, normalize, partials, stop, mcX); break;
        default: throw std::runtime_error("unsupported padded rank");
    }
    return;
#endif
    if (mcX) {                                        // multicast replication by the Gram kernel (Engine::mc_mode 1)
        switch (KP) {
            case 16: normalize_gram_mma_kernel<16, 64, true><<<grid, GramMmaGeom<16>::WARPS * 32, 0, s>>>(X, ncols, d, normalize, partials, stop, mcX); break;
            case 32: normalize_gram_mma_kernel<32, 64, true><<<grid, GramMmaGeom<32>::WARPS * 32, 0, s>>>(X, ncols, d, normalize, partials, stop, mcX); break;
            case 64: normalize_gram_mma_kernel<64, 32, true><<<grid, GramMmaGeom<64>::WARPS * 32, 0, s>>>(X, ncols, d, normalize, partials, stop, mcX); break;
            case 128: normalize_gram_mma_kernel<128, 32, true><<<grid, GramMmaGeom<128>::WARPS * 32, 0, s>>>(X, ncols, d, normalize, partials, stop, mcX); break;
            default: throw std::runtime_error("unsupported padded rank");
        }
        return;
    }
    switch (KP) {
        case 16: normalize_gram_mma_kernel<16, 64><<<grid, GramMmaGeom<16>::WARPS * 32, 0, s>>>(X, ncols, d, normalize, partials, stop, nullptr); break;
        case 32: normalize_gram_mma_kernel<32, 64><<<grid, GramMmaGeom<32>::WARPS * 32, 0, s>>>(X, ncols, d, normalize, partials, stop, nullptr); break;
        case 64: normalize_gram_mma_kernel<64, 32><<<grid, GramMmaGeom<64>::WARPS * 32, 0, s>>>(X, ncols, d, normalize, partials, stop, nullptr); break;
        case 128: normalize_gram_mma_kernel<128, 32><<<grid, GramMmaGeom<128>::WARPS * 32, 0, s>>>(X, ncols, d, normalize, partials, stop, nullptr); break;
        default: throw std::runtime_error("unsupported padded rank");
    }
}

template <int KP>
static void launch_masked_t(const MaskedParams& p, int num_sms, cudaStream_t s, int* grid_out) {
    auto kern = masked_half_step_kernel<KP>;
    const size_t smem = masked_smem_bytes<KP>();
    const int threads = masked_warps<KP>() * 32;
    static OccCache cache;
    const int cached_occ = cached_occupancy(cache, kern, threads, smem, "masked_half_step_kernel does not fit on an SM");
    const int grid = num_sms * cached_occ;
    if (grid_out) { *grid_out = grid; return; }
    kern<<<grid, threads, smem, s>>>(p);
}
static void launch_masked(int KP, const MaskedParams& p, int num_sms, cudaStream_t s, int* grid_out = nullptr) {
    switch (KP) {
        case 16: launch_masked_t<16>(p, num_sms, s, grid_out); break;
        case 32: launch_masked_t<32>(p, num_sms, s, grid_out); break;
        case 64: launch_masked_t<64>(p, num_sms, s, grid_out); break;
        case 128: launch_masked_t<128>(p, num_sms, s, grid_out); break;
        default: throw std::runtime_error("unsupported padded rank");
    }
}

template <int KP>
static void launch_cv_t(const CvParams& p, int num_sms, cudaStream_t s, int* grid_out) {
    auto kern = cv_half_step_kernel<KP>;
    const size_t smem = masked_smem_bytes<KP>();
    const int threads = masked_warps<KP>() * 32;
    static OccCache cache;
    const int cached_occ = cached_occupancy(cache, kern, threads, smem, "cv_half_step_kernel does not fit on an SM");
    const int grid = num_sms * cached_occ;
    if (grid_out) { *grid_out = grid; return; }
    kern<<<grid, threads, smem, s>>>(p);
}
static void launch_cv(int KP, const CvParams& p, int num_sms, cudaStream_t s, int* grid_out = nullptr) {
    switch (KP) {
        case 16: launch_cv_t<16>(p, num_sms, s, grid_out); break;
        case 32: launch_cv_t<32>(p, num_sms, s, grid_out); break;
        case 64: launch_cv_t<64>(p, num_sms, s, grid_out); break;
        case 128: launch_cv_t<128>(p, num_sms, s, grid_out); break;
        default: throw std::runtime_error("unsupported padded rank");
    }
}

// out = G with (G_ii + 1e-15) + L2 on the diagonal (fit_cv.hpp:414-417 / :578-581)
static __global__ void cv_prepare_gram_kernel(const float* __restrict__ G, int KP, int k, float L2, float* __restrict__ out,
                                              const int* __restrict__ stop_flag) {
    if (*stop_flag) return;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= KP * KP) return;
    const int i = e % KP, j = e / KP;
    float v = G[e];
    if (i == j && i < k) {
        v = __fadd_rn(v, 1e-15f);
        if (L2 > 0.f) v = __fadd_rn(v, L2);
    }
    out[e] = v;
}

// ---------------------------------------------------------------------------------------------
// Engine
// ---------------------------------------------------------------------------------------------
// __constant__ SolverConsts slots (kernels_solve.cuh) are per device: a live engine owns one slot of ITS device, so
// two engines running on their own streams can never overwrite each other's diagonal blocks / reciprocals.
static std::mutex g_slot_mu;
static bool g_slot_used[kMaxDevices][kConstSlots] = {};

Engine::Engine(int dev) : device(dev) {
    B200_REQUIRE(dev >= 0 && dev < kMaxDevices, "device ordinal out of range");
    {
        std::lock_guard<std::mutex> g(g_slot_mu);
        const_slot = -1;
        for (int s = 0; s < kConstSlots && const_slot < 0; ++s)
            if (!g_slot_used[dev][s]) { g_slot_used[dev][s] = true; const_slot = s; }
    }
    B200_REQUIRE(const_slot >= 0, "too many live engines on this device (8 constant-memory solver slots); destroy one first");
    try {
        init_device_objects();
    } catch (...) {
        std::lock_guard<std::mutex> g(g_slot_mu);
        g_slot_used[dev][const_slot] = false;
        throw;
    }
}

void Engine::init_device_objects() {
    B200_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    B200_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    num_sms = prop.multiProcessorCount;
    // The main stream carries the critical path (solve -> exchange -> Gram -> exchange -> LLT); the side stream only
    // re-normalises the peers' blocks of a replicated factor (pure HBM streaming with a machine-filling grid). With
    // priorities the Gram kernel's CTAs are placed first and the streaming kernel fills what is left, instead of the
    // Gram waiting for a full wave of it (N = 8: the "loss" section was 0.19 ms of a 1.14 ms iteration).
    int prio_lo = 0, prio_hi = 0;
    B200_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));     // lo = least (numerically greatest)
    B200_CUDA_CHECK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_hi));
    B200_CUDA_CHECK(cudaStreamCreateWithPriority(&side_stream, cudaStreamNonBlocking, prio_lo));
    B200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    B200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    B200_CUDA_CHECK(cudaEventCreate(&ev_loop_begin));
    B200_CUDA_CHECK(cudaEventCreate(&ev_loop_end));
    state.ensure(1);
    counters.ensure(16);
    sweep_counter.ensure(1);
    B200_CUDA_CHECK(cudaMallocHost(&h_state, sizeof(DevState) * 2));
    B200_CUDA_CHECK(cudaMemsetAsync(state.ptr, 0, sizeof(DevState), stream));
    B200_CUDA_CHECK(cudaMemsetAsync(sweep_counter.ptr, 0, sizeof(unsigned long long), stream));
}

Engine::~Engine() {
    cudaSetDevice(device);
    cudaStreamSynchronize(stream);
    cudaStreamSynchronize(side_stream);
    comm_destroy();                                       // peer / multicast mappings go first (needs both streams alive)
    if (iter_graph) cudaGraphExecDestroy(iter_graph);
    for (auto& sec : prof_events)
        for (auto& pr : sec) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    cudaStreamSynchronize(side_stream);
    cudaEventDestroy(ev_fork);
    cudaEventDestroy(ev_join);
    cudaStreamDestroy(side_stream);
    cudaEventDestroy(ev_loop_begin);
    cudaEventDestroy(ev_loop_end);
    if (h_state) cudaFreeHost(h_state);
    cudaStreamDestroy(stream);
    std::lock_guard<std::mutex> g(g_slot_mu);
    if (const_slot >= 0) g_slot_used[device][const_slot] = false;
}

void Engine::use_device() const { B200_CUDA_CHECK(cudaSetDevice(device)); }

// ---- matrix ---------------------------------------------------------------------------------
// Two sparse operands live on the device:
//   opH = A[:, J]  as CSC (n_loc columns, inner = global row id)    — gathered by the H half-step
//   opW = A[I, :]ᵀ as CSC (m_loc columns, inner = global column id) — gathered by the W half-step
// Single GPU: J = all columns, I = all rows, opW = opHᵀ. Sharded (world > 1): J / I are this rank's
// column / row block, so BOTH half-steps stay fully fused and local; only factors are exchanged.
static void block_of(int total, int world, int rank, int* begin, int* count) {
    const int nb = (total + world - 1) / world;
    const int lo = std::min(total, rank * nb);
    *begin = lo;
    *count = std::min(total, lo + nb) - lo;
}

void Engine::set_dims(int m_, int n_) {
    B200_REQUIRE(m_ > 0 && n_ > 0, "set_matrix: bad dimensions");
    drop_iteration_graph();
    fit_active = false;                                   // a fit in flight does not survive a new matrix
    m = m_; n = n_;
    equal_partition = true;
    auto fill = [&](std::vector<int>& cuts, const std::vector<int>& pending, int total, const char* what) {
        if (!pending.empty()) {
            B200_REQUIRE(static_cast<int>(pending.size()) == world + 1 && pending.front() == 0 && pending.back() == total, what);
            for (int r = 0; r < world; ++r) B200_REQUIRE(pending[r] <= pending[r + 1], what);
            cuts = pending;
            equal_partition = false;
            return;
        }
        cuts.assign(world + 1, total);
        for (int r = 0; r < world; ++r) { int lo = 0, cnt = 0; block_of(total, world, r, &lo, &cnt); cuts[r] = lo; }
    };
    fill(col_cuts, pending_col_cuts, n, "set_partition: the column cuts do not cover this matrix (need world+1 ascending cuts, 0 .. n)");
    fill(row_cuts, pending_row_cuts, m, "set_partition: the row cuts do not cover this matrix (need world+1 ascending cuts, 0 .. m)");
    col_begin = col_cuts[rank]; n_loc = col_cuts[rank + 1] - col_cuts[rank];
    row_begin = row_cuts[rank]; m_loc = row_cuts[rank + 1] - row_cuts[rank];
    m_pad = ((m + world - 1) / world) * world;
    n_pad = ((n + world - 1) / world) * world;
    matrix_ready = false;
    factors_ready = false;
    npanels[0] = npanels[1] = 1;
}

void Engine::set_partition(const int* cc, const int* rc) {
    pending_col_cuts.clear();
    pending_row_cuts.clear();
    if (cc) pending_col_cuts.assign(cc, cc + world + 1);
    if (rc) pending_row_cuts.assign(rc, rc + world + 1);
}

void Engine::factor_checksum(unsigned long long* out3) {
    use_device();
    B200_REQUIRE(factors_ready, "no factors");
    unsigned long long* acc = scratch<unsigned long long>(6, 4);
    B200_CUDA_CHECK(cudaMemsetAsync(acc, 0, 4 * sizeof(unsigned long long), stream));
    checksum_kernel<<<num_sms * 8, 256, 0, stream>>>(W_T.ptr, m, k, KP, acc);
    checksum_kernel<<<num_sms * 8, 256, 0, stream>>>(H.ptr, n, k, KP, acc + 1);
    checksum_kernel<<<1, 256, 0, stream>>>(d.ptr, 1, k, KP, acc + 2);
    B200_CUDA_CHECK(cudaMemcpyAsync(out3, acc, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

void Engine::finish_matrix() { finish_matrix_from(Ax.ptr, nnz, world > 1); }

void Engine::finish_matrix_from(const float* vals, int64_t vcnt, bool reduce_over_ranks) {
    // tr(AᵀA) in fp64 (primitives/primitives.hpp:101-115) over the given values (this rank's column block, then
    // summed over ranks — or the whole matrix when every device holds it, set_matrix_host_shard)
    const int nb = 1024;
    struct { double* ptr; } part{scratch<double>(6, nb)};
    sumsq_kernel<<<nb, 256, 0, stream>>>(vals, vcnt, part.ptr);
    std::vector<double> hp(nb);
    B200_CUDA_CHECK(cudaMemcpyAsync(hp.data(), part.ptr, nb * sizeof(double), cudaMemcpyDeviceToHost, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    double s = 0.0;
    for (double v : hp) s += v;
    double cnt = static_cast<double>(vcnt);
    if (reduce_over_ranks) {
        DeviceBuffer<double> t;
        t.ensure(2);
        const double h[2] = {s, cnt};
        B200_CUDA_CHECK(cudaMemcpyAsync(t.ptr, h, sizeof(h), cudaMemcpyHostToDevice, stream));
        allreduce_f64(t.ptr, 2);
        double r[2];
        B200_CUDA_CHECK(cudaMemcpyAsync(r, t.ptr, sizeof(r), cudaMemcpyDeviceToHost, stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        s = r[0]; cnt = r[1];
    }
    trAtA = static_cast<float>(s);
    nnz_global = static_cast<int64_t>(cnt);
    matrix_ready = true;
}

// dst = srcᵀ with ascending inner indices. src: CSC with `ncols` columns and `nrows` rows.
// Stable LSD radix sort of (row, position): positions — hence column indices — stay ascending
// within each row, exactly the order Eigen's transpose() produces (nmf/fit_cpu.hpp:251-253).
void Engine::transpose_csc(const int* sp, const int* si, const float* sx, int ncols, int nrows, int64_t cnt,
                           DeviceBuffer<int>& dp, DeviceBuffer<int>& di, DeviceBuffer<float>& dx, int col_id_offset) {
    dp.ensure(static_cast<size_t>(nrows) + 1);
    di.ensure(std::max<int64_t>(cnt, 1) + 4);
    dx.ensure(std::max<int64_t>(cnt, 1) + 4);
    if (cnt == 0) {
        B200_CUDA_CHECK(cudaMemsetAsync(dp.ptr, 0, (static_cast<size_t>(nrows) + 1) * sizeof(int), stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        return;
    }
    struct P { int* ptr; };
    struct U { unsigned* ptr; };
    const P col_of{scratch<int>(1, cnt)}, keys_out{scratch<int>(2, cnt)};
    const U perm_in{scratch<unsigned>(3, cnt)}, perm_out{scratch<unsigned>(4, cnt)};
    const int T = 256;
    const unsigned gb = static_cast<unsigned>((cnt + T - 1) / T);
    expand_columns_kernel<<<gb, T, 0, stream>>>(sp, ncols, cnt, col_of.ptr, col_id_offset);
    iota_kernel<<<gb, T, 0, stream>>>(perm_in.ptr, cnt);
    int end_bit = 1;
    while ((1LL << end_bit) < static_cast<long long>(nrows) && end_bit < 31) ++end_bit;
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, si, keys_out.ptr, perm_in.ptr, perm_out.ptr, cnt, 0, end_bit, stream);
    struct { unsigned char* ptr; } temp{scratch<unsigned char>(5, temp_bytes)};
    B200_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(temp.ptr, temp_bytes, si, keys_out.ptr, perm_in.ptr, perm_out.ptr,
                                                    cnt, 0, end_bit, stream));
    permute_gather_kernel<<<gb, T, 0, stream>>>(perm_out.ptr, col_of.ptr, sx, cnt, di.ptr, dx.ptr);
    row_pointers_kernel<<<(nrows + 1 + T - 1) / T, T, 0, stream>>>(keys_out.ptr, cnt, nrows, dp.ptr);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

template <class ValT>
void Engine::upload_csc(int ncols, int64_t cnt, const int* col_ptr, const int* row_idx, const ValT* values,
                        DeviceBuffer<int>& dp, DeviceBuffer<int>& di, DeviceBuffer<float>& dx) {
    // cudaMemcpyDefault: the sources may be host OR device pointers (zero-copy entry, UVA decides)
    B200_REQUIRE(cnt >= 0 && cnt < (1LL << 31), "set_matrix: nnz must fit int32 (reference boundary, bridge_nmf.hpp:196)");
    dp.ensure(static_cast<size_t>(ncols) + 1);
    di.ensure(std::max<int64_t>(cnt, 1) + 4);
    dx.ensure(std::max<int64_t>(cnt, 1) + 4);
    B200_CUDA_CHECK(cudaMemcpyAsync(dp.ptr, col_ptr, (static_cast<size_t>(ncols) + 1) * sizeof(int), cudaMemcpyDefault, stream));
    if (cnt > 0) {
        B200_CUDA_CHECK(cudaMemcpyAsync(di.ptr, row_idx, cnt * sizeof(int), cudaMemcpyDefault, stream));
        if (std::is_same<ValT, float>::value) {
            B200_CUDA_CHECK(cudaMemcpyAsync(dx.ptr, values, cnt * sizeof(float), cudaMemcpyDefault, stream));
        } else {
            cudaPointerAttributes at{};
            const bool on_device = cudaPointerGetAttributes(&at, values) == cudaSuccess && at.type == cudaMemoryTypeDevice;
            cudaGetLastError();
            const double* src = reinterpret_cast<const double*>(values);
            if (!on_device) {                                  // host doubles: stage once, convert on the device
                double* tmp = scratch<double>(0, cnt);
                B200_CUDA_CHECK(cudaMemcpyAsync(tmp, values, cnt * sizeof(double), cudaMemcpyDefault, stream));
                src = tmp;
            }
            f64_to_f32_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, stream>>>(src, dx.ptr, cnt);
            B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    h2d_bytes += (static_cast<size_t>(ncols) + 1) * sizeof(int) + cnt * (sizeof(int) + sizeof(ValT));
}

template <class ValT>
void Engine::set_matrix_host(int m_, int n_, int64_t nnz_, const int* col_ptr, const int* row_idx, const ValT* values) {
    use_device();
    B200_REQUIRE(world == 1, "set_matrix: with a communicator use set_matrix_sharded");
    set_dims(m_, n_);
    nnz = nnz_;
    const auto t0 = std::chrono::steady_clock::now();
    upload_csc<ValT>(n, nnz, col_ptr, row_idx, values, Ap, Ai, Ax);
    const auto t1 = std::chrono::steady_clock::now();
    transpose_csc(Ap.ptr, Ai.ptr, Ax.ptr, n, m, nnz, Atp, Ati, Atx, 0);
    nnz_w = nnz;
    finish_matrix();
    has_mask = false;                                      // a new matrix invalidates the previous mask
    phase_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    phase_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
}
template void Engine::set_matrix_host<float>(int, int, int64_t, const int*, const int*, const float*);
template void Engine::set_matrix_host<double>(int, int, int64_t, const int*, const int*, const double*);

// The caller already holds CSC(Aᵀ) — e.g. a StreamPress .spz file written with include_transpose, decoded by the package's
// own reader (streampress/sparsepress_v2.hpp:58, :652, :1318): both operands are uploaded as they are and
// the device transpose (radix sort, 6 ms at C4) is skipped. The transpose must list, for every row of A, its entries
// with ascending column indices (what Eigen's transpose() and the .spz writer produce); it is trusted, like A itself.
template <class ValT>
void Engine::set_matrix_host_with_transpose(int m_, int n_, int64_t nnz_, const int* col_ptr, const int* row_idx,
                                            const ValT* values, const int* t_col_ptr, const int* t_row_idx, const ValT* t_values) {
    use_device();
    B200_REQUIRE(world == 1, "set_matrix_with_transpose: single-GPU entry (sharded fits slice their operands themselves)");
    B200_REQUIRE(col_ptr[n_] == nnz_ && t_col_ptr[m_] == nnz_, "set_matrix_with_transpose: A and its transpose disagree on nnz");
    set_dims(m_, n_);
    nnz = nnz_w = nnz_;
    const auto t0 = std::chrono::steady_clock::now();
    upload_csc<ValT>(n, nnz, col_ptr, row_idx, values, Ap, Ai, Ax);
    upload_csc<ValT>(m, nnz, t_col_ptr, t_row_idx, t_values, Atp, Ati, Atx);
    const auto t1 = std::chrono::steady_clock::now();
    finish_matrix();
    has_mask = false;
    phase_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    phase_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
}
template void Engine::set_matrix_host_with_transpose<float>(int, int, int64_t, const int*, const int*, const float*, const int*, const int*, const float*);
template void Engine::set_matrix_host_with_transpose<double>(int, int, int64_t, const int*, const int*, const double*, const int*, const int*, const double*);

// Sharded: the caller hands this rank its column block A[:, J] (CSC, n_loc columns, global row ids) and
// its row block A[I, :] (CSC, n columns, row ids relative to the block). See rcppml_b200/shard.py.
template <class ValT>
void Engine::set_matrix_sharded(int m_, int n_, const int* cb_ptr, const int* cb_idx, const ValT* cb_val,
                                const int* rb_ptr, const int* rb_idx, const ValT* rb_val) {
    use_device();
    set_dims(m_, n_);
    nnz = cb_ptr[n_loc];
    upload_csc<ValT>(n_loc, nnz, cb_ptr, cb_idx, cb_val, Ap, Ai, Ax);
    DeviceBuffer<int> rp, ri;
    DeviceBuffer<float> rx;
    nnz_w = rb_ptr[n];
    upload_csc<ValT>(n, nnz_w, rb_ptr, rb_idx, rb_val, rp, ri, rx);
    transpose_csc(rp.ptr, ri.ptr, rx.ptr, n, std::max(m_loc, 1), nnz_w, Atp, Ati, Atx, 0);
    finish_matrix();
}
template void Engine::set_matrix_sharded<float>(int, int, const int*, const int*, const float*, const int*, const int*, const float*);
template void Engine::set_matrix_sharded<double>(int, int, const int*, const int*, const double*, const int*, const int*, const double*);

template <class ValT>
void Engine::set_matrix_sharded_with_transpose(int m_, int n_, const int* cb_ptr, const int* cb_idx, const ValT* cb_val,
                                               const int* tb_ptr, const int* tb_idx, const ValT* tb_val) {
    use_device();
    set_dims(m_, n_);
    nnz = cb_ptr[n_loc];
    upload_csc<ValT>(n_loc, nnz, cb_ptr, cb_idx, cb_val, Ap, Ai, Ax);
    nnz_w = tb_ptr[m_loc];
    upload_csc<ValT>(std::max(m_loc, 0), nnz_w, tb_ptr, tb_idx, tb_val, Atp, Ati, Atx);
    finish_matrix();
    has_mask = false;
}
template void Engine::set_matrix_sharded_with_transpose<float>(int, int, const int*, const int*, const float*, const int*, const int*, const float*);
template void Engine::set_matrix_sharded_with_transpose<double>(int, int, const int*, const int*, const double*, const int*, const int*, const double*);

void Engine::planned_blocks(int m_, int n_, int* cb, int* nl, int* rb, int* ml) const {
    auto one = [&](const std::vector<int>& pending, int total, int* begin, int* count) {
        if (static_cast<int>(pending.size()) == world + 1) { *begin = pending[rank]; *count = pending[rank + 1] - pending[rank]; }
        else block_of(total, world, rank, begin, count);
    };
    one(pending_col_cuts, n_, cb, nl);
    one(pending_row_cuts, m_, rb, ml);
}

// ---- in-process multi-GPU ingest (abi_reference.cu: RCPPML_NUM_GPUS) ------------------------------------------------
// Each device receives ONLY its column block of the host matrix over its own PCIe link (nnz/G entries instead of nnz);
// the row block it needs for the W half-step is assembled from all devices' column blocks over NVLink
// (assemble_row_block). The steps are separate methods because the host threads meet at barriers between them.

// Step 1: dimensions + column partition (pending cuts, set_partition), own column block up. Host-synchronised.
template <class ValT>
void Engine::upload_col_block_host(int m_, int n_, const int* col_ptr, const int* row_idx, const ValT* values) {
    use_device();
    B200_REQUIRE(world > 1, "upload_col_block_host: needs comm_init_local first");
    const auto t0 = std::chrono::steady_clock::now();
    set_dims(m_, n_);
    const int64_t p0 = col_ptr[col_begin], p1 = col_ptr[col_begin + n_loc];
    nnz = p1 - p0;
    upload_csc<ValT>(n_loc, nnz, col_ptr + col_begin, row_idx + p0, values + p0, Ap, Ai, Ax);
    if (p0 != 0) rebase_int_kernel<<<(n_loc + 1 + 255) / 256, 256, 0, stream>>>(Ap.ptr, n_loc + 1, static_cast<int>(p0));
    // row work of this block (balanced row partition): hist[r] = entries of row r in A[:, J]
    row_hist.ensure(static_cast<size_t>(m));
    B200_CUDA_CHECK(cudaMemsetAsync(row_hist.ptr, 0, static_cast<size_t>(m) * sizeof(int), stream));
    if (nnz > 0) row_histogram_kernel<<<num_sms * 8, 256, 0, stream>>>(Ai.ptr, nnz, row_hist.ptr);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    phase_ms[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
template void Engine::upload_col_block_host<float>(int, int, const int*, const int*, const float*);
template void Engine::upload_col_block_host<double>(int, int, const int*, const int*, const double*);

// Step 2 (one device, after every device finished step 1): row cuts balanced by work = non-zeros of the row +
// per_item (the k x k solve is a constant per row), from the sum of all devices' histograms. cuts_out: world + 1 ints.
void Engine::balanced_row_cuts(Engine* const* all, int per_item, int* cuts_out) {
    use_device();
    HistPtrs hp{};
    for (int g = 0; g < world; ++g) hp.p[g] = all[g]->row_hist.ptr;
    long long* work = scratch<long long>(8, static_cast<size_t>(m));
    long long* inc = scratch<long long>(9, static_cast<size_t>(m));
    sum_histograms_kernel<<<(m + 255) / 256, 256, 0, stream>>>(hp, world, m, per_item, work);
    size_t temp_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, temp_bytes, work, inc, m, stream);
    unsigned char* temp = scratch<unsigned char>(5, temp_bytes);
    B200_CUDA_CHECK(cub::DeviceScan::InclusiveSum(temp, temp_bytes, work, inc, m, stream));
    int* dcuts = scratch<int>(10, 16);
    balanced_cuts_kernel<<<1, 32, 0, stream>>>(inc, m, world, dcuts);
    B200_CUDA_CHECK(cudaMemcpyAsync(cuts_out, dcuts, (world + 1) * sizeof(int), cudaMemcpyDeviceToHost, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

// Step 3: install the row cuts, pull the runs of every column that fall into this device's rows out of all column
// blocks (peer reads), transpose the row block locally. Returns this device's share of tr(AtA) (over its column block).
void Engine::install_row_cuts(const int* row_cuts_in) {
    row_cuts.assign(row_cuts_in, row_cuts_in + world + 1);
    B200_REQUIRE(row_cuts.front() == 0 && row_cuts.back() == m, "install_row_cuts: the row cuts do not cover the matrix");
    for (int r = 0; r < world; ++r) B200_REQUIRE(row_cuts[r] <= row_cuts[r + 1], "install_row_cuts: cuts must ascend");
    row_begin = row_cuts[rank]; m_loc = row_cuts[rank + 1] - row_cuts[rank];
    equal_partition = false;
}

double Engine::assemble_row_block(Engine* const* all) {
    use_device();
    const auto t0 = std::chrono::steady_clock::now();
    int* rstart = scratch<int>(8, static_cast<size_t>(n) + 1);
    int* rcnt = scratch<int>(9, static_cast<size_t>(n) + 1);
    int* rp = scratch<int>(10, static_cast<size_t>(n) + 1);
    for (int g = 0; g < world; ++g) {
        const Engine& P = *all[g];
        B200_REQUIRE(P.col_begin == col_cuts[g] && P.n_loc == col_cuts[g + 1] - col_cuts[g], "assemble_row_block: peers disagree on the column partition");
        if (P.n_loc > 0)
            rowblock_count_kernel<<<(P.n_loc + 255) / 256, 256, 0, stream>>>(P.Ap.ptr, P.Ai.ptr, P.n_loc, row_begin, row_begin + m_loc,
                                                                             rstart + P.col_begin, rcnt + P.col_begin);
    }
    B200_CUDA_CHECK(cudaMemsetAsync(rcnt + n, 0, sizeof(int), stream));
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, rcnt, rp, n + 1, stream);
    unsigned char* temp = scratch<unsigned char>(5, temp_bytes);
    B200_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, rcnt, rp, n + 1, stream));
    int total = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&total, rp + n, sizeof(int), cudaMemcpyDeviceToHost, stream));
    // tr(AtA) share of the own column block meanwhile (fixed-order fp64 partials, summed on the host below)
    const int nb = 1024;
    double* part = scratch<double>(6, nb);
    sumsq_kernel<<<nb, 256, 0, stream>>>(Ax.ptr, nnz, part);
    std::vector<double> hp(nb);
    B200_CUDA_CHECK(cudaMemcpyAsync(hp.data(), part, nb * sizeof(double), cudaMemcpyDeviceToHost, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    B200_REQUIRE(total >= 0, "assemble_row_block: nnz overflowed int32");
    nnz_w = total;
    int* ri = scratch<int>(11, static_cast<size_t>(std::max(total, 1)) + 4);
    float* rx = scratch<float>(12, static_cast<size_t>(std::max(total, 1)) + 4);
    for (int g = 0; g < world; ++g) {
        const Engine& P = *all[g];
        if (P.n_loc > 0)
            rowblock_copy_kernel<<<num_sms * 8, 256, 0, stream>>>(P.Ai.ptr, P.Ax.ptr, P.n_loc, rstart + P.col_begin, rp + P.col_begin,
                                                                  row_begin, ri, rx);
    }
    B200_CUDA_CHECK(cudaGetLastError());
    transpose_csc(rp, ri, rx, n, std::max(m_loc, 1), nnz_w, Atp, Ati, Atx, 0);      // host-synchronised at its end
    double s = 0.0;
    for (double v : hp) s += v;
    phase_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return s;
}

// Step 4 (after every device finished step 3 — nobody reads this device's column block any more): tr(AtA) and nnz
// of the whole matrix as summed by the caller in rank order.
void Engine::finish_matrix_local(double sumsq_total, int64_t nnz_total) {
    trAtA = static_cast<float>(sumsq_total);
    nnz_global = nnz_total;
    matrix_ready = true;
    has_mask = false;
}

// Step 5: own factor blocks up (rows [row_begin, +m_loc) of the full host W_T, columns [col_begin, +n_loc) of H), in two
// halves: the H2D copies into a staging buffer start on the side stream as soon as the row cuts are installed and run
// on the copy engine WHILE the main stream assembles and transposes the row block (NVLink + SMs); padding / conversion
// into the replicas follows once the matrix is complete.
template <class T>
void Engine::start_factor_block_upload(int k_, const T* W_full, const T* H_full) {
    use_device();
    B200_REQUIRE(k_ >= 1 && k_ <= kMaxKP, "rank must be in [1, 128]");
    const size_t wcount = static_cast<size_t>(m_loc) * k_, hcount = static_cast<size_t>(n_loc) * k_;
    T* stage = scratch<T>(7, wcount + hcount + 1);
    if (wcount) B200_CUDA_CHECK(cudaMemcpyAsync(stage, W_full + static_cast<size_t>(row_begin) * k_, wcount * sizeof(T), cudaMemcpyHostToDevice, side_stream));
    if (hcount) B200_CUDA_CHECK(cudaMemcpyAsync(stage + wcount, H_full + static_cast<size_t>(col_begin) * k_, hcount * sizeof(T), cudaMemcpyHostToDevice, side_stream));
    B200_CUDA_CHECK(cudaEventRecord(ev_join, side_stream));
    h2d_bytes += (wcount + hcount) * sizeof(T);
}
template void Engine::start_factor_block_upload<float>(int, const float*, const float*);
template void Engine::start_factor_block_upload<double>(int, const double*, const double*);

template <class T>
void Engine::finish_factor_block_upload(int k_) {
    use_device();
    alloc_factors(k_);
    const auto t0 = std::chrono::steady_clock::now();
    const size_t wcount = static_cast<size_t>(m_loc) * k, hcount = static_cast<size_t>(n_loc) * k;
    T* stage = scratch<T>(7, wcount + hcount + 1);                           // (same size as in start_: no reallocation)
    B200_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_join, 0));
    if (wcount) pad_convert_kernel<T><<<static_cast<unsigned>((static_cast<long long>(m_loc) * KP + 255) / 256), 256, 0, stream>>>(
        stage, W_T.ptr + static_cast<size_t>(row_begin) * KP, m_loc, k, KP);
    if (hcount) pad_convert_kernel<T><<<static_cast<unsigned>((static_cast<long long>(n_loc) * KP + 255) / 256), 256, 0, stream>>>(
        stage + wcount, H.ptr + static_cast<size_t>(col_begin) * KP, n_loc, k, KP);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    phase_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
template void Engine::finish_factor_block_upload<float>(int);
template void Engine::finish_factor_block_upload<double>(int);

// Step 6 (after comm_attach_local on every engine): complete the replicas with the peers' blocks over NVLink.
void Engine::pull_factor_blocks_from_peers() {
    use_device();
    B200_REQUIRE(peers_ready && peers_local, "pull_factor_blocks_from_peers: peers not attached");
    for (int g = 0; g < world; ++g) {
        if (g == rank) continue;
        const size_t wo = static_cast<size_t>(row_cuts[g]) * KP, wc = static_cast<size_t>(row_cuts[g + 1] - row_cuts[g]) * KP;
        const size_t ho = static_cast<size_t>(col_cuts[g]) * KP, hc = static_cast<size_t>(col_cuts[g + 1] - col_cuts[g]) * KP;
        if (wc) B200_CUDA_CHECK(cudaMemcpyAsync(W_T.ptr + wo, peer_W[g] + wo, wc * sizeof(float), cudaMemcpyDefault, stream));
        if (hc) B200_CUDA_CHECK(cudaMemcpyAsync(H.ptr + ho, peer_H[g] + ho, hc * sizeof(float), cudaMemcpyDefault, stream));
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

// SURVEY.md §8d generator. Columns [c0, c0+nc), rows kept in [r0, r1) and stored relative to r0.
void Engine::synth_block(int m_, int c0, int nc, int r0, int r1, double density, uint64_t seed, DeviceBuffer<int>& dp,
                         DeviceBuffer<int>& di, DeviceBuffer<float>& dx, int64_t* cnt_out) {
    const long long cnt_ll = std::llround(static_cast<double>(m_) * density);
    B200_REQUIRE(cnt_ll >= 1 && cnt_ll <= 8192, "synthetic: round(m*density) must be in [1, 8192]");
    const int cnt = static_cast<int>(cnt_ll);
    DeviceBuffer<int> counts;
    counts.ensure(static_cast<size_t>(nc) + 1);
    dp.ensure(static_cast<size_t>(nc) + 1);
    auto run = [&](int pass, int* rows, float* vals) {
        if (cnt <= 1024) synth_column_kernel<1024><<<nc, 256, 0, stream>>>(m_, nc, c0, cnt, r0, r1, seed, pass, counts.ptr, dp.ptr, rows, vals);
        else if (cnt <= 4096) synth_column_kernel<4096><<<nc, 256, 0, stream>>>(m_, nc, c0, cnt, r0, r1, seed, pass, counts.ptr, dp.ptr, rows, vals);
        else synth_column_kernel<8192><<<nc, 256, 0, stream>>>(m_, nc, c0, cnt, r0, r1, seed, pass, counts.ptr, dp.ptr, rows, vals);
        B200_CUDA_CHECK(cudaGetLastError());
    };
    run(0, nullptr, nullptr);
    B200_CUDA_CHECK(cudaMemsetAsync(counts.ptr + nc, 0, sizeof(int), stream));
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, counts.ptr, dp.ptr, nc + 1, stream);
    DeviceBuffer<unsigned char> temp;
    temp.ensure(temp_bytes);
    B200_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp.ptr, temp_bytes, counts.ptr, dp.ptr, nc + 1, stream));
    int total = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&total, dp.ptr + nc, sizeof(int), cudaMemcpyDeviceToHost, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    B200_REQUIRE(total >= 0, "synthetic: nnz overflowed int32");
    *cnt_out = total;
    di.ensure(std::max<int64_t>(total, 1) + 4);
    dx.ensure(std::max<int64_t>(total, 1) + 4);
    run(1, di.ptr, dx.ptr);
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

// Stand-alone matrix made of columns [col_begin, col_begin + n_local) of the generator's m × ∞ family.
void Engine::set_matrix_synthetic(int m_, int n_local, int col_begin_, double density, uint64_t seed) {
    use_device();
    B200_REQUIRE(world == 1, "set_matrix_synthetic: with a communicator use set_matrix_synthetic_sharded");
    set_dims(m_, n_local);
    synth_block(m, col_begin_, n, 0, m, density, seed, Ap, Ai, Ax, &nnz);
    transpose_csc(Ap.ptr, Ai.ptr, Ax.ptr, n, m, nnz, Atp, Ati, Atx, 0);
    nnz_w = nnz;
    finish_matrix();
}

// The m × n generator matrix, this rank's column block and row block (any world size).
void Engine::set_matrix_synthetic_sharded(int m_, int n_, double density, uint64_t seed) {
    use_device();
    set_dims(m_, n_);
    synth_block(m, col_begin, std::max(n_loc, 1), 0, m, density, seed, Ap, Ai, Ax, &nnz);
    if (n_loc == 0) nnz = 0;
    if (world == 1) {
        transpose_csc(Ap.ptr, Ai.ptr, Ax.ptr, n, m, nnz, Atp, Ati, Atx, 0);
        nnz_w = nnz;
    } else {
        DeviceBuffer<int> rp, ri;
        DeviceBuffer<float> rx;
        synth_block(m, 0, n, row_begin, row_begin + m_loc, density, seed, rp, ri, rx, &nnz_w);
        transpose_csc(rp.ptr, ri.ptr, rx.ptr, n, std::max(m_loc, 1), nnz_w, Atp, Ati, Atx, 0);
    }
    finish_matrix();
}

// ---- factors --------------------------------------------------------------------------------
void Engine::alloc_factors(int k_) {
    B200_REQUIRE(matrix_ready, "set a matrix before the factors");
    B200_REQUIRE(k_ >= 1 && k_ <= kMaxKP, "rank must be in [1, 128]");
    drop_iteration_graph();
    fit_active = false;                                   // a fit in flight does not survive new factors
    k = k_;
    LANES = lanes_for_rank(k);
    KP = padded_rank(k);
    nv_override = 0;
    if (const char* env = std::getenv("RCPPML_B200_NV")) nv_override = std::atoi(env);   // tuning knobs (1, 2 or 4)
    nv_short_override = 0;
    if (const char* env = std::getenv("RCPPML_B200_NV_SHORT")) nv_short_override = std::atoi(env);
    // Coordinate descent runs in its own kernel with narrow lane groups (kernels_cd.cuh). RCPPML_B200_CD_GEOM
    // selects another geometry (LANES + 100*(NV-1)); RCPPML_B200_CD_KERNEL=1 falls back to half_step_kernel<CD>.
    // Measured at C4 (profiles/r01p_cd_geometries.json): the narrowest group wins wherever the solve dominates
    // (short columns: the W half-step); long columns (the H half-step, 1000 non-zeros) gather better one step wider.
    cd_geom = (KP == 16) ? 301 : (KP == 32) ? 701 : (KP == 64) ? 702 : 704;
    cd_geom_long = (KP == 16) ? 102 : (KP == 32) ? 302 : (KP == 64) ? 304 : 704;
    if (const char* env = std::getenv("RCPPML_B200_CD_GEOM")) {
        const int g = std::atoi(env);
        B200_REQUIRE(cd_geometry_matches(g, KP), "RCPPML_B200_CD_GEOM does not match the padded rank");
        cd_geom = cd_geom_long = g;
    }
    if (const char* env = std::getenv("RCPPML_B200_CD_KERNEL")) { if (std::atoi(env) == 1) cd_geom = cd_geom_long = 0; }
    tiled_mode = 1;
    if (const char* env = std::getenv("RCPPML_B200_TILED")) tiled_mode = std::atoi(env);
    tiled_min_batches = 0.5;
    if (const char* env = std::getenv("RCPPML_B200_TILED_MIN_BATCHES")) tiled_min_batches = std::atof(env);
    // measured on 8 GPUs (profiles/r02k_bench_c4_n8*.json): 8 CTAs/SM 1.067 ms per iteration, 2 CTAs/SM 1.120 ms — the
    // throttled kernel no longer fits under the Gram chain and the next solve waits for it
    side_ctas_per_sm = 8;
    if (const char* env = std::getenv("RCPPML_B200_SIDE_CTAS")) side_ctas_per_sm = std::max(1, std::atoi(env));
    tiled_sl_override = 0;
    if (const char* env = std::getenv("RCPPML_B200_TILED_SL")) tiled_sl_override = std::atoi(env);
    tiled_cta_mode = 0;
    if (const char* env = std::getenv("RCPPML_B200_TILED_CTA")) tiled_cta_mode = std::max(0, std::min(2, std::atoi(env)));
    narrow_min_cols = 8.0 * num_sms * 24;
    if (std::getenv("RCPPML_B200_CD_GEOM")) narrow_min_cols = 0.0;           // an explicit geometry applies to every size
    if (const char* env = std::getenv("RCPPML_B200_NARROW_MIN_COLS")) narrow_min_cols = std::atof(env);
    W_T.want_vmm = H.want_vmm = mc_wanted;
    if ((peers_ready || mc_ready) &&
        (static_cast<size_t>(m_pad) * KP > W_T.count || static_cast<size_t>(n_pad) * KP > H.count || KP * KP > xchg_ne_max ||
         W_T.vmm != W_T.want_vmm || H.vmm != H.want_vmm))
        comm_ipc_close();                                 // the mapped buffers are about to move: back to NCCL
    W_T.ensure(static_cast<size_t>(m_pad) * KP);          // padded to equal row / column blocks (all-gather)
    H.ensure(static_cast<size_t>(n_pad) * KP);
    B200_CUDA_CHECK(cudaMemsetAsync(W_T.ptr, 0, static_cast<size_t>(m_pad) * KP * sizeof(float), stream));
    B200_CUDA_CHECK(cudaMemsetAsync(H.ptr, 0, static_cast<size_t>(n_pad) * KP * sizeof(float), stream));
    d.ensure(KP);
    G_w.ensure(static_cast<size_t>(KP) * KP);
    G_h.ensure(static_cast<size_t>(KP) * KP);
    M1.ensure(static_cast<size_t>(KP) * KP);
    M2.ensure(static_cast<size_t>(KP) * KP);
    dblk.ensure(sizeof(SolverConsts) / sizeof(float));       // [kMaxKP*4] diagonal blocks, then [kMaxKP] reciprocals
    gram_grid = num_sms * 4;
    gram_partials.ensure(static_cast<size_t>(gram_grid) * KP * KP);
    B200_CUDA_CHECK(cudaMemsetAsync(gram_partials.ptr, 0, gram_partials.bytes(), stream));   // upper tiles are never written
    // the solve grid depends on the instantiation; size the partial buffers for the largest
    int gmax = 0;
    HalfStepParams dummy{};
    for (int solver = 0; solver < 2; ++solver) {
        for (int geom : {LANES, (KP == 64 || KP == 128) ? 100 + KP / 8 : LANES, (KP == 64 || KP == 128) ? 300 + KP / 16 : LANES}) {
            int g = 0;
            launch_half_step(geom, solver, BSRC_GATHER, OUT_SOLVE, dummy, num_sms, stream, &g);
            gmax = std::max(gmax, g);
        }
    }
    if (cd_geom) {
        for (int geom : {cd_geom, cd_geom_long}) {
            int g = 0;
            launch_cd_half_step(geom, dummy, num_sms, stream, &g);
            gmax = std::max(gmax, g);
        }
    }
    if (tiled_mode) {
        for (int solver = 0; solver < 2; ++solver)
            for (int geom : {LANES, (KP == 64 || KP == 128) ? 100 + KP / 8 : LANES, KP == 64 ? 4016 : LANES, KP == 64 ? 4108 : LANES,
                             (KP == 64 && solver == SOLVER_CHOL) ? 14108 : LANES, (KP == 64 && solver == SOLVER_CHOL) ? 24108 : LANES}) {
                int g = 0;
                launch_tiled_half_step(geom, solver, dummy, num_sms, stream, &g);
                gmax = std::max(gmax, g);
            }
    }
    { int g = 0; MaskedParams md{}; launch_masked(KP, md, num_sms, stream, &g); gmax = std::max(gmax, g); }
    { int g = 0; CvParams cd{}; launch_cv(KP, cd, num_sms, stream, &g); gmax = std::max(gmax, g); }
    solve_grid_max = gmax;
    solve_partials.ensure(static_cast<size_t>(gmax) * (KP + 1));
    red_gram.ensure(static_cast<size_t>(KP) * KP);
    red_small.ensure(KP + 1);
    std::vector<float> ones(KP, 1.f);
    B200_CUDA_CHECK(cudaMemcpyAsync(d.ptr, ones.data(), KP * sizeof(float), cudaMemcpyHostToDevice, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    factors_ready = true;
}

// W_T (k × m) and H (k × n) are FULL (replicated on every rank when sharded).
template <class T>
void Engine::set_factors_host(int k_, const T* W_T_host, const T* H_host) {
    use_device();
    alloc_factors(k_);
    const auto t0 = std::chrono::steady_clock::now();
    auto upload = [&](const T* src, float* dst, long long ncols) {
        struct { T* ptr; } tmp{scratch<T>(0, static_cast<size_t>(ncols) * k)};
        B200_CUDA_CHECK(cudaMemcpyAsync(tmp.ptr, src, static_cast<size_t>(ncols) * k * sizeof(T), cudaMemcpyHostToDevice, stream));
        const long long total = ncols * KP;
        pad_convert_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(tmp.ptr, dst, ncols, k, KP);
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        h2d_bytes += static_cast<size_t>(ncols) * k * sizeof(T);
    };
    upload(W_T_host, W_T.ptr, m);
    upload(H_host, H.ptr, n);
    phase_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
template void Engine::set_factors_host<float>(int, const float*, const float*);
template void Engine::set_factors_host<double>(int, const double*, const double*);

// set_matrix_host + set_factors_host with the factor upload overlapped: the CSC goes up first (the transpose needs
// all of it), then the factors stream into a staging buffer on the side stream (copy engine) WHILE the main stream
// sorts and transposes A and reduces tr(AtA); the padding/conversion kernels wait for the copies by event.
// Same bytes over PCIe, the ~6 ms device transpose of C4 disappears from the wall time of the call.
template <class ValT, class T>
void Engine::set_matrix_and_factors_host(int m_, int n_, int64_t nnz_, const int* col_ptr, const int* row_idx,
                                         const ValT* values, int k_, const T* W_T_host, const T* H_host) {
    use_device();
    B200_REQUIRE(world == 1, "set_matrix: with a communicator use set_matrix_sharded");
    B200_REQUIRE(k_ >= 1 && k_ <= kMaxKP, "rank must be in [1, 128]");
    set_dims(m_, n_);
    nnz = nnz_;
    const auto t0 = std::chrono::steady_clock::now();
    upload_csc<ValT>(n, nnz, col_ptr, row_idx, values, Ap, Ai, Ax);            // host-synchronised at its end
    const auto t1 = std::chrono::steady_clock::now();
    const size_t wcount = static_cast<size_t>(m) * k_, hcount = static_cast<size_t>(n) * k_;
    T* stage = scratch<T>(7, wcount + hcount);
    B200_CUDA_CHECK(cudaMemcpyAsync(stage, W_T_host, wcount * sizeof(T), cudaMemcpyHostToDevice, side_stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(stage + wcount, H_host, hcount * sizeof(T), cudaMemcpyHostToDevice, side_stream));
    B200_CUDA_CHECK(cudaEventRecord(ev_join, side_stream));
    transpose_csc(Ap.ptr, Ai.ptr, Ax.ptr, n, m, nnz, Atp, Ati, Atx, 0);
    nnz_w = nnz;
    finish_matrix();
    has_mask = false;
    const auto t2 = std::chrono::steady_clock::now();
    alloc_factors(k_);
    B200_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_join, 0));
    pad_convert_kernel<T><<<static_cast<unsigned>((static_cast<long long>(m) * KP + 255) / 256), 256, 0, stream>>>(stage, W_T.ptr, m, k, KP);
    pad_convert_kernel<T><<<static_cast<unsigned>((static_cast<long long>(n) * KP + 255) / 256), 256, 0, stream>>>(stage + wcount, H.ptr, n, k, KP);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    h2d_bytes += (wcount + hcount) * sizeof(T);
    phase_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    phase_ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    phase_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t2).count();
}
template void Engine::set_matrix_and_factors_host<double, double>(int, int, int64_t, const int*, const int*, const double*, int, const double*, const double*);
template void Engine::set_matrix_and_factors_host<float, float>(int, int, int64_t, const int*, const int*, const float*, int, const float*, const float*);

// nmf/nmf_init.hpp:167-182: one SplitMix64(seed) stream, W_T first, then H. h_col_begin shifts H inside
// a wider stream (the matrix is columns [h_col_begin, h_col_begin + n) of a larger problem).
void Engine::init_factors(int k_, uint32_t seed, int h_col_begin) {
    use_device();
    alloc_factors(k_);
    const unsigned long long state0 = (seed == 0) ? 12345ULL : static_cast<unsigned long long>(seed);   // rng.hpp:73
    const long long tw = static_cast<long long>(m) * KP, th = static_cast<long long>(n) * KP;
    init_uniform_kernel<<<static_cast<unsigned>((tw + 255) / 256), 256, 0, stream>>>(W_T.ptr, m, k, KP, state0, 0ULL);
    const unsigned long long first_h = static_cast<unsigned long long>(m) * k + static_cast<unsigned long long>(h_col_begin) * k;
    init_uniform_kernel<<<static_cast<unsigned>((th + 255) / 256), 256, 0, stream>>>(H.ptr, n, k, KP, state0, first_h);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

template <class T>
void Engine::get_factors_host(T* W_T_host, T* H_host, T* d_host) {
    use_device();
    B200_REQUIRE(factors_ready, "no factors");
    const auto t0 = std::chrono::steady_clock::now();
    auto download = [&](const float* src, T* dst, long long ncols) {
        if (!dst) return;
        struct { T* ptr; } tmp{scratch<T>(0, static_cast<size_t>(ncols) * k)};
        const long long total = ncols * k;
        unpad_convert_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src, tmp.ptr, ncols, k, KP);
        B200_CUDA_CHECK(cudaMemcpyAsync(dst, tmp.ptr, static_cast<size_t>(total) * sizeof(T), cudaMemcpyDeviceToHost, stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        d2h_bytes += static_cast<size_t>(total) * sizeof(T);
    };
    download(W_T.ptr, W_T_host, m);
    download(H.ptr, H_host, n);
    download(d.ptr, d_host, 1);
    phase_ms[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
template void Engine::get_factors_host<float>(float*, float*, float*);
template void Engine::get_factors_host<double>(double*, double*, double*);

// Block-wise factor I/O for sharded fits (world > 1; with world == 1 the blocks are the whole factors). Upload: own
// blocks into place in the zero-filled replicas, then one in-place all-gather per factor (equal padded blocks, NCCL).
template <class T>
void Engine::set_factor_blocks_host(int k_, const T* W_blk, const T* H_blk) {
    use_device();
    B200_REQUIRE(world == 1 || comm_ready(), "set_factor_blocks: needs the communicator (comm_init)");
    alloc_factors(k_);
    const auto t0 = std::chrono::steady_clock::now();
    auto upload = [&](const T* src, float* dst, long long ncols) {
        if (ncols <= 0) return;
        struct { T* ptr; } tmp{scratch<T>(0, static_cast<size_t>(ncols) * k)};
        B200_CUDA_CHECK(cudaMemcpyAsync(tmp.ptr, src, static_cast<size_t>(ncols) * k * sizeof(T), cudaMemcpyHostToDevice, stream));
        const long long total = ncols * KP;
        pad_convert_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(tmp.ptr, dst, ncols, k, KP);
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        h2d_bytes += static_cast<size_t>(ncols) * k * sizeof(T);
    };
    upload(W_blk, W_T.ptr + static_cast<size_t>(row_begin) * KP, m_loc);
    upload(H_blk, H.ptr + static_cast<size_t>(col_begin) * KP, n_loc);
    if (world > 1) {
        allgather_rows(W_T.ptr, row_cuts, m_pad, RCPPML_B200_SEC_COMM);
        allgather_rows(H.ptr, col_cuts, n_pad, RCPPML_B200_SEC_COMM);
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    phase_ms[2] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
template void Engine::set_factor_blocks_host<float>(int, const float*, const float*);
template void Engine::set_factor_blocks_host<double>(int, const double*, const double*);

template <class T>
void Engine::get_factor_blocks_host(T* W_blk, T* H_blk, T* d_host) {
    use_device();
    B200_REQUIRE(factors_ready, "no factors");
    const auto t0 = std::chrono::steady_clock::now();
    auto download = [&](const float* src, T* dst, long long ncols) {
        if (!dst || ncols <= 0) return;
        struct { T* ptr; } tmp{scratch<T>(0, static_cast<size_t>(ncols) * k)};
        const long long total = ncols * k;
        unpad_convert_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src, tmp.ptr, ncols, k, KP);
        B200_CUDA_CHECK(cudaMemcpyAsync(dst, tmp.ptr, static_cast<size_t>(total) * sizeof(T), cudaMemcpyDeviceToHost, stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
        d2h_bytes += static_cast<size_t>(total) * sizeof(T);
    };
    download(W_T.ptr + static_cast<size_t>(row_begin) * KP, W_blk, m_loc);
    download(H.ptr + static_cast<size_t>(col_begin) * KP, H_blk, n_loc);
    download(d.ptr, d_host, 1);
    phase_ms[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
template void Engine::get_factor_blocks_host<float>(float*, float*, float*);
template void Engine::get_factor_blocks_host<double>(double*, double*, double*);

// ---- profiling sections -----------------------------------------------------------------------
// Sections of an iteration, named as the reference's own profiling sections — also the NVTX ranges a
// profiler sees around the enqueue of each section's kernels (host side; nested for the Gram inside the loss section).
static const char* const kSectionNames[RCPPML_B200_NUM_SECTIONS] = {"gram_H", "fused_rhs_nnls_H", "scaling_H", "gram_W",
                                                                     "fused_rhs_nnls_W", "scaling_W", "loss", "comm"};

void Engine::sec_begin(int sec, cudaStream_t on) {
    nvtxRangePushA(kSectionNames[sec]);
    if (!profiling) return;
    if (!on) on = stream;
    auto& pool = prof_events[sec];
    if (prof_used[sec] == static_cast<int>(pool.size())) {
        cudaEvent_t a, b;
        B200_CUDA_CHECK(cudaEventCreate(&a));
        B200_CUDA_CHECK(cudaEventCreate(&b));
        pool.emplace_back(a, b);
    }
    B200_CUDA_CHECK(cudaEventRecord(pool[prof_used[sec]].first, on));
}
void Engine::sec_end(int sec, cudaStream_t on) {
    nvtxRangePop();
    if (!profiling) return;
    if (!on) on = stream;
    B200_CUDA_CHECK(cudaEventRecord(prof_events[sec][prof_used[sec]].second, on));
    ++prof_used[sec];
}
void Engine::collect_profile() {
    if (!profiling) return;
    for (int s = 0; s < RCPPML_B200_NUM_SECTIONS; ++s) {
        for (int i = 0; i < prof_used[s]; ++i) {
            float ms = 0.f;
            B200_CUDA_CHECK(cudaEventElapsedTime(&ms, prof_events[s][i].first, prof_events[s][i].second));
            prof_ms[s] += ms;
        }
        prof_used[s] = 0;
    }
}

// ---- one iteration ----------------------------------------------------------------------------
void Engine::normalize_cfg(const rcppml_b200_config& c) {
    cfg = c;
    B200_REQUIRE(cfg.k == k, "config rank differs from the factors' rank");
    if (cfg.cd_maxit <= 0) cfg.cd_maxit = 10;          // src/RcppFunctions_nmf.cpp:75
    if (!(cfg.cd_tol > 0.f)) cfg.cd_tol = 1e-8f;       // src/RcppFunctions_nmf.cpp:76
    if (cfg.patience <= 0) cfg.patience = 5;           // core/constants.hpp:89
    B200_REQUIRE(cfg.tol >= 0.f, "tol must be non-negative");                          // core/config.hpp:426
    B200_REQUIRE(cfg.norm_type >= 0 && cfg.norm_type <= 2, "norm_type must be 0, 1 or 2");
}

// Peer-memory sharding: the other ranks' half_step_kernel pushed their solved (un-normalised) columns into this
// replica; divide them by d here (same IEEE division as the owner applies to its block).
// Runs on the side stream, forked after `d` is final and joined before the next solve kernel gathers from X:
// it overlaps with the Gram of the own block, the small all-reduces and the solver set-up on the main stream
// (those touch only this rank's block of X and k x k data).
void Engine::fork_side_stream() {
    B200_CUDA_CHECK(cudaEventRecord(ev_fork, stream));
    side_forked = true;
}

// Enqueued AFTER the main-stream work it overlaps with (fork_side_stream marks the point in the main stream it
// depends on), so that the main stream's kernels reach the hardware queues first.
void Engine::normalize_peer_blocks(float* X, long long ncols, long long lo, long long hi, bool normalize) {
    const bool forked = side_forked;
    side_forked = false;
    if (!normalize || ncols == hi - lo) return;
    if (!forked) B200_CUDA_CHECK(cudaEventRecord(ev_fork, stream));
    B200_CUDA_CHECK(cudaStreamWaitEvent(side_stream, ev_fork, 0));
    sec_begin(RCPPML_B200_SEC_COMM, side_stream);
    scale_columns_kernel<<<num_sms * side_ctas_per_sm, 256, 0, side_stream>>>(X, ncols, KP, d.ptr, lo, hi, &state.ptr->stop);
    launches[RCPPML_B200_SEC_COMM] += 1;
    sec_end(RCPPML_B200_SEC_COMM, side_stream);
    B200_CUDA_CHECK(cudaEventRecord(ev_join, side_stream));
    side_pending = true;
}

void Engine::join_side_stream() {
    if (!side_pending) return;
    B200_CUDA_CHECK(cudaStreamWaitEvent(stream, ev_join, 0));
    side_pending = false;
}

void Engine::gram(float* X, long long ncols, bool normalize, float* G_out, int sec, bool reduce_over_ranks, bool replicate) {
    sec_begin(sec);
    // one CTA per column tile at most: a rank's block of a sharded factor can be a few hundred tiles, and every CTA's
    // k x k fp64 partial is summed afterwards (592 partials of 32 KB at k = 64)
    const long long tc = (KP >= 64) ? 32 : 64;
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(gram_grid, (ncols + tc - 1) / tc)));
    // replicate: X is this rank's freshly solved block of a replicated factor and the factors are multicast-bound —
    // the kernel writes the normalised block into every replica (multimem.st) instead of only this one
    float* mcX = (replicate && mc_ready) ? mc_alias(X) : nullptr;
    launch_normalize_gram(KP, X, ncols, d.ptr, normalize ? 1 : 0, gram_partials.ptr, &state.ptr->stop, grid, stream, mcX);
    const int nelem = KP * KP;
    sum_partials_kernel<<<(nelem + 31) / 32, dim3(32, 8), 0, stream>>>(gram_partials.ptr, grid, nelem, red_gram.ptr, &state.ptr->stop);
    launches[sec] += 2;
    if (reduce_over_ranks && world > 1) allreduce_f64(red_gram.ptr, nelem);
    gram_from_sums_kernel<<<(nelem + 255) / 256, 256, 0, stream>>>(red_gram.ptr, KP, k, G_out, &state.ptr->stop);
    launches[sec] += 1;
    sec_end(sec);
}

template <int KP>
static void launch_prepare_solver(const float* G, int k, float L2, int solver, float* M1, float* M2, float* dblk, float* rcp,
                                  DevState* st, cudaStream_t stream) {
    auto kern = prepare_solver_kernel<KP>;
    const size_t smem = static_cast<size_t>(KP) * KP * sizeof(float);
    static OccCache cache;                                   // per device (see cached_occupancy)
    cached_occupancy(cache, kern, kPrepThreads, smem, "prepare_solver_kernel does not fit on an SM");
    kern<<<1, kPrepThreads, smem, stream>>>(G, k, L2, solver, M1, M2, dblk, rcp, st);
}

void Engine::prepare_solver(const float* G, float L2, int sec) {
    sec_begin(sec);
    const int solver = cfg.solver_mode == 0 ? SOLVER_CD : SOLVER_CHOL;
    float* rcp = dblk.ptr + kMaxKP * 4;
    switch (KP) {
        case 16: launch_prepare_solver<16>(G, k, L2, solver, M1.ptr, M2.ptr, dblk.ptr, rcp, state.ptr, stream); break;
        case 32: launch_prepare_solver<32>(G, k, L2, solver, M1.ptr, M2.ptr, dblk.ptr, rcp, state.ptr, stream); break;
        case 64: launch_prepare_solver<64>(G, k, L2, solver, M1.ptr, M2.ptr, dblk.ptr, rcp, state.ptr, stream); break;
        case 128: launch_prepare_solver<128>(G, k, L2, solver, M1.ptr, M2.ptr, dblk.ptr, rcp, state.ptr, stream); break;
        default: throw std::runtime_error("unsupported padded rank");
    }
    // warp-uniform operands -> this engine's constant-memory slot (kernels_solve.cuh SolverConsts); rcp follows
    // dblk in one device buffer, laid out like the struct
    B200_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_solver, dblk.ptr, sizeof(SolverConsts), sizeof(SolverConsts) * const_slot,
                                            cudaMemcpyDeviceToDevice, stream));
    launches[sec] += 1;
    sec_end(sec);
}

// Lane geometry per launch. A lane owns NV 128-bit words of a factor row. Short columns (few
// non-zeros, the solve dominates) run faster with half as many lanes per column owning two words
// each (twice the columns per warp share every pivot broadcast); long columns (gather dominates)
// prefer one word per lane. Measured on B200 at k = 64: W-update (100 nnz/col) 3.78 -> 2.94 ms with
// NV = 2, H-update (1000 nnz/col) 1.87 -> 1.97 ms.
int Engine::geometry_for(long long nnz_, long long ncols) const {
    int nv = 1;
    if (KP == 64 || KP == 128) {
        const double avg = ncols > 0 ? static_cast<double>(nnz_) / static_cast<double>(ncols) : 0.0;
        nv = (avg < 400.0) ? 2 : 1;
        if (avg < 400.0 && (nv_short_override == 1 || nv_short_override == 2 || nv_short_override == 4)) nv = nv_short_override;
        if (nv_override == 1 || nv_override == 2 || nv_override == 4) nv = nv_override;
    }
    return nv == 1 ? LANES : 100 * (nv - 1) + KP / (4 * nv);
}

static int pick_cols_per_fetch(long long nnz, long long ncols, int num_sms) {
    // ~2K non-zeros per group fetch bounds the tail imbalance; 1..16 columns. With few columns per rank
    // (sharded runs) also keep >= ~8 fetches per resident warp, or half the warps would sit idle while the
    // others work through an oversized batch (measured: 125K rows on 8 GPUs took 2x their share).
    const double avg = ncols > 0 ? static_cast<double>(nnz) / static_cast<double>(ncols) : 1.0;
    int c = static_cast<int>(2048.0 / std::max(1.0, avg));
    const long long groups = static_cast<long long>(num_sms) * 3 * 8 * 4;      // resident lane groups (upper bound)
    c = static_cast<int>(std::min<long long>(c, std::max<long long>(1, ncols / (groups * 8))));
    return std::max(1, std::min(16, c));
}

HalfStepParams Engine::solve_params(int which, bool warm) const {
    HalfStepParams p{};
    const bool h = (which == 0);
    p.colptr = h ? Ap.ptr : Atp.ptr;
    p.rowidx = h ? Ai.ptr : Ati.ptr;
    p.vals = h ? Ax.ptr : Atx.ptr;
    p.F = h ? W_T.ptr : H.ptr;                          // full (replicated) factor being gathered
    p.X = h ? H.ptr : W_T.ptr;                          // full factor being solved; this rank owns a block of it
    p.M1 = M1.ptr; p.M2 = M2.ptr; p.dblk = dblk.ptr; p.rcp = dblk.ptr + kMaxKP * 4; p.cslot = const_slot;
    p.B = nullptr;
    p.ncols = h ? n_loc : m_loc;
    p.col_offset = h ? col_begin : row_begin;
    p.k = k;
    p.L1 = h ? cfg.L1_H : cfg.L1_W;
    p.ub = h ? cfg.ub_H : cfg.ub_W;
    p.cd_tol = cfg.cd_tol;
    p.inv_k = 1.0f / static_cast<float>(k);            // nnls_batch.hpp:84
    p.cd_maxit = cfg.cd_maxit;
    p.nonneg = h ? cfg.nonneg_H : cfg.nonneg_W;
    p.warm = warm ? 1 : 0;
    p.norm_type = cfg.norm_type;
    p.want_cross = h ? 0 : 1;
    p.cols_per_fetch = pick_cols_per_fetch(h ? nnz : nnz_w, p.ncols, num_sms);
    p.work_counter = counters.ptr + which;
    p.partials = solve_partials.ptr;
    p.stop_flag = &state.ptr->stop;
    p.sweep_counter = sweep_counter.ptr;
    p.npeers = 0;
    p.mcX = nullptr;
    if (peers_ready && mc_ready) {                      // multicast: mode 1 — the normalising Gram kernel replicates the
        if (mc_mode == 2) p.mcX = mc_alias(p.X);        // block; mode 2 — this kernel's stores go through the multicast alias
    } else if (peers_ready) {
        for (int r = 0; r < world; ++r)
            if (r != rank) p.peerX[p.npeers++] = h ? peer_H[r] : peer_W[r];
    }
    return p;
}

// Row panels: when the factor a half-step gathers from is larger than what stays resident in the 126 MB L2
// (C4: W_T is 256 MB; C5: both factors), nearly every round of row loads pays one DRAM miss. The rows of a
// column are sorted, so the entries that fall in a panel of rows are a contiguous run: the half-step runs as P
// launches, pass q touching only panel q of the factor (L2-resident after first touch) and carrying the running
// right-hand sides in `carry`. RCPPML_B200_PANEL_MB sets the panel size (0 disables).
void Engine::build_panels() {
    // Measured on B200 (profiles/r01l_*, r01m_*): C5 (k = 128, factors 2.56 GB / 256 MB) 289.8 -> 147.8 ms per
    // iteration with 40 MB panels; C4's H half-step (k = 64, W_T 256 MB, 72 % L2 hits) already runs at the L2
    // throughput cap (14 TB/s) and gains nothing (1.718 vs 1.732 ms). Default rule: panels when the factor is
    // > 4 x L2, or > 60 MB with 512-byte rows; an explicit RCPPML_B200_PANEL_MB overrides the rule.
    double panel_mb = 40.0;
    bool explicit_mb = false;
    if (const char* env = std::getenv("RCPPML_B200_PANEL_MB")) { panel_mb = std::atof(env); explicit_mb = true; }
    for (int which = 0; which < 2; ++which) {
        const bool h = (which == 0);
        const long long frows = h ? m : n;                                  // rows of the gathered factor
        const long long ncols = h ? n_loc : m_loc;
        const double fbytes = static_cast<double>(frows) * KP * sizeof(float);
        int P = 1;
        const bool wanted = explicit_mb || fbytes > 4.0 * 126.0 * 1048576.0 || KP >= 128;
        if (wanted && panel_mb > 0.0 && fbytes > 1.5 * panel_mb * 1048576.0) P = static_cast<int>(std::ceil(fbytes / (panel_mb * 1048576.0)));
        P = std::min(P, 64);
        if (ncols <= 0 || (h ? nnz : nnz_w) == 0) P = 1;
        npanels[which] = P;
        if (P == 1) continue;
        const int rows_per_panel = static_cast<int>((frows + P - 1) / P);
        panel_bounds[which].ensure(static_cast<size_t>(P + 1) * ncols);
        const long long total = static_cast<long long>(P + 1) * ncols;
        panel_bounds_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
            h ? Ap.ptr : Atp.ptr, h ? Ai.ptr : Ati.ptr, static_cast<int>(ncols), P, rows_per_panel, panel_bounds[which].ptr);
        B200_CUDA_CHECK(cudaGetLastError());
        carry.ensure(static_cast<size_t>(std::max(n_loc, m_loc)) * KP);
    }
}

// Kernel selection for a half-step (all choices give bit-identical results; measured in profiles/r01p_*):
//   * few columns (< narrow_min_cols, default 8 per resident warp): the one-geometry kernel with WIDE lane groups.
//     Narrow groups put 16-32 columns in a warp; with few columns most of the machine idles and each column is
//     slower (pbmc3k k=32 CD: 1.7 ms/iteration wide, 4.2 ms narrow; Cholesky 0.49 vs 1.2 ms);
//   * coordinate descent otherwise: narrow groups — the tiled kernel (robust to skewed column lengths; long
//     columns gather wide), or cd_half_step_kernel with RCPPML_B200_TILED=0;
//   * Cholesky with >= tiled_min_batches batches per resident warp: the tiled kernel for short columns (C4 W
//     half-step 2.64 -> 2.35 ms; every pass of a row-panelled half-step is "short": C5 147 -> 114 ms) and for long
//     columns at k = 32 / 128; long columns at k = 64 / 16 (C4 H half-step) stay one-geometry.
bool Engine::use_narrow_cd(long long ncols) const {
    return cd_geom != 0 && static_cast<double>(ncols) >= narrow_min_cols;
}

bool Engine::use_tiled(int solver, long long cnt, long long ncols) const {
    if (tiled_mode == 0) return false;
    if (tiled_mode == 2) return true;
    if (solver == SOLVER_CD) return use_narrow_cd(ncols);
    // Cholesky. `cnt` is the non-zero count of the pass that ends in the solve (one row panel's share when the
    // half-step is split into panel passes).
    const int cb = (KP == 64) ? 16 : (KP == 128) ? 8 : 32;                       // columns per warp batch
    const long long resident_warps = static_cast<long long>(num_sms) * ((KP == 128) ? 8 : 24);
    // fewer than ~half a batch per resident warp: the narrow solve geometry starves the machine (pbmc3k, 0.12
    // batches per warp: 0.49 -> 1.2 ms); from ~0.9 batches per warp on the tiled kernel is ahead (k = 32, 100 K columns)
    if (static_cast<double>(ncols) < tiled_min_batches * static_cast<double>(cb) * static_cast<double>(resident_warps)) return false;
    const double avg = ncols > 0 ? static_cast<double>(cnt) / static_cast<double>(ncols) : 0.0;
    if (avg < 400.0) return true;                 // short columns: the solve dominates at every rank
    // Long columns (measured on the C4 shape, 1000 non-zeros per column, profiles/r01r_chol_kernel_selection.json):
    // k = 128: 4.48 -> 3.61 ms and k = 32: 1.11 -> 0.84 ms with the tiled kernel, but k = 64: 1.73 -> 3.87 ms and
    // k = 16: 0.70 -> 0.76 ms — those stay on the one-geometry kernel.
    return KP == 128 || KP == 32;
}

int Engine::tiled_gather_geom(long long cnt, long long ncols, int solver) const {
    int g = geometry_for(cnt, ncols);                           // LANES + 100*(NV-1), NV in {1, 2, 4}
    if (g >= 300) g = 100 + KP / 8;                             // NV = 4 is not instantiated for the tiled kernel
    // k = 64, Cholesky: batches of 8 columns (4 lanes x 4 words per column) instead of 16. Measured on the C4 W half-step
    // (profiles/r02a_rank_shapes.jsonl, r02b_rank_shapes_n1.jsonl; ms, 16- vs 8-column batches): one GPU 2.355 -> 2.215;
    // per-rank shapes of N = 2 / 4 / 8: 1.227 -> 1.138, 0.657 -> 0.603, 0.378 -> 0.341 (a rank with 125 K rows holds only
    // 2.2 sixteen-column batches per resident warp: one warp in five works a third batch while the others idle).
    // Coordinate descent keeps the narrowest groups (its cost is the group-uniform part of every coordinate step).
    // RCPPML_B200_TILED_SL = 2 / 4 forces a geometry.
    if (KP == 64) {
        int sl = tiled_sl_override;
        if (sl == 0) sl = (solver == SOLVER_CHOL) ? 4 : 2;
        if (sl == 4) g += 4000;
        if (sl == 8 && g == 108) g += 8000;
        // RCPPML_B200_TILED_CTA: 1 = one 768-thread CTA per SM, 2 = that + hybrid gather (Cholesky, 8 lanes x 2 words)
        if (g == 4108 && solver == SOLVER_CHOL && tiled_cta_mode > 0) g += 10000 * tiled_cta_mode;
    }
    return g;
}

// kind 0: half_step_kernel (one geometry, wide groups); 1: cd_half_step_kernel (one geometry, narrow groups);
// 2: tiled_half_step_kernel (geom = gather geometry).
void Engine::launch_solver(int kind, int geom, int solver, const HalfStepParams& p, int* grid_out) {
    if (kind == 2) launch_tiled_half_step(geom, solver, p, num_sms, stream, grid_out);
    else if (kind == 1) launch_cd_half_step(geom, p, num_sms, stream, grid_out);
    else launch_half_step(geom, solver, BSRC_GATHER, OUT_SOLVE, p, num_sms, stream, grid_out);
}

void Engine::solve(int which, bool warm, int sec) {
    HalfStepParams p = solve_params(which, warm);
    const int solver = cfg.solver_mode == 0 ? SOLVER_CD : SOLVER_CHOL;
    const long long cnt = which == 0 ? nnz : nnz_w;
    const int P = npanels[which];
    int geom = geometry_for(cnt, p.ncols);
    const bool tiled = use_tiled(solver, cnt / std::max(1, P), p.ncols);
    const bool narrow_cd = !tiled && solver == SOLVER_CD && use_narrow_cd(p.ncols);
    const int kind = tiled ? 2 : (narrow_cd ? 1 : 0);
    if (tiled) {
        geom = tiled_gather_geom(cnt, p.ncols, solver);
    } else if (narrow_cd) {
        const double avg = p.ncols > 0 ? static_cast<double>(cnt) / static_cast<double>(p.ncols) : 0.0;
        geom = avg >= 400.0 ? cd_geom_long : cd_geom;
        p.cols_per_fetch = 1;                                      // a CD column is thousands of instructions
    }
    if ((tiled || narrow_cd) && p.want_cross) {                    // parking space for the pre-L1 right-hand sides
        carry.ensure(static_cast<size_t>(std::max(n_loc, m_loc)) * KP);
        p.braw = carry.ptr;                                        // (a panel pass reads carry[j] before it parks there)
    }
    int grid = 0;
    launch_solver(kind, geom, solver, p, &grid);
    last_solve_grid = grid;
    sec_begin(sec);
    if (P > 1) {
        // passes 0..P-2 only gather (their segments are short: use the short-column geometry); the last pass
        // gathers its segment on top of the carried sums and solves
        const int geom_pass = geometry_for(cnt / P, p.ncols);
        HalfStepParams q = p;
        q.carry = carry.ptr;
        for (int pass = 0; pass < P; ++pass) {
            q.seg_begin = panel_bounds[which].ptr + static_cast<size_t>(pass) * p.ncols;
            q.seg_end = panel_bounds[which].ptr + static_cast<size_t>(pass + 1) * p.ncols;
            q.carry_load = pass > 0 ? 1 : 0;
            B200_CUDA_CHECK(cudaMemsetAsync(q.work_counter, 0, sizeof(int), stream));
            if (pass + 1 < P) {
                q.cols_per_fetch = pick_cols_per_fetch(cnt / P, p.ncols, num_sms);
                launch_half_step(geom_pass, SOLVER_CD, BSRC_GATHER, OUT_RHS, q, num_sms, stream);
            } else {
                q.cols_per_fetch = narrow_cd ? 1 : pick_cols_per_fetch(cnt / P, p.ncols, num_sms);
                launch_solver(kind, geom, solver, q);
            }
            launches[sec] += 1;
        }
    } else {
        B200_CUDA_CHECK(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));
        launch_solver(kind, geom, solver, p);
        launches[sec] += 1;
    }
    sec_end(sec);
}

void Engine::scale_finalize(int sec, bool reduce_over_ranks) {
    sec_begin(sec);
    const int nelem = KP + 1;                                   // k row sums + the loss cross term
    sum_partials_kernel<<<(nelem + 31) / 32, dim3(32, 8), 0, stream>>>(solve_partials.ptr, last_solve_grid, nelem, red_small.ptr, &state.ptr->stop);
    if (reduce_over_ranks && world > 1) allreduce_f64(red_small.ptr, nelem);
    scale_from_sums_kernel<<<1, 128, 0, stream>>>(red_small.ptr, KP, k, cfg.norm_type, d.ptr, &state.ptr->stop);
    launches[sec] += 2;
    sec_end(sec);
}

void Engine::loss(int sec) {
    loss_kernel<<<1, 256, 0, stream>>>(G_w.ptr, G_h.ptr, d.ptr, KP, k, red_small.ptr + KP, 1, trAtA, cfg.tol,
                                       cfg.patience, loss_hist.ptr, static_cast<int>(loss_hist.count), state.ptr);
    launches[sec] += 1;
}

// One ALS iteration (nmf/fit_cpu.hpp:444-1825). With world > 1 every rank solves its own column block of
// H and row block of W_T with the SAME fused kernels (its sparse operands are A[:,J_g] and A[I_g,:]ᵀ), the
// k×k Grams / row sums / loss cross term are all-reduced in fp64, and the two factor blocks are all-gathered.
// No right-hand side ever crosses NVLink, and every column's arithmetic is identical to the single-GPU run.
void Engine::enqueue_iteration() {
    const bool warm = iters_enqueued > 0;                                   // fit_cpu.hpp:523 / :755
    const bool normalize = cfg.norm_type != 2;
    const bool sharded = world > 1;
    // With peer-mapped factors (comm_ipc_import) the solve kernels replicate every solved column into the
    // other GPUs' copies while they run, the small all-reduces are one-shot peer-memory kernels, and each
    // rank normalises the whole replicated factor locally: no NCCL call in the loop, no exposed all-gather.
    const bool p2p = sharded && peers_ready;
    const bool repl = p2p && mc_ready && mc_mode == 1;                      // multicast replication by the Gram kernel
    const bool ucast = p2p && !repl;                                        // the solve kernel replicates (unicast peer stores or
                                                                            // multicast), every rank re-normalises the peers' blocks
    float* Hblk = H.ptr + static_cast<size_t>(col_begin) * KP;
    float* Wblk = W_T.ptr + static_cast<size_t>(row_begin) * KP;
    // ---- H update (fit_cpu.hpp:488-645). The Gram of W_T and the solver operands built from it are produced at the END
    // of the previous iteration (the Gram doubles as the loss's; the LLT then runs while the side stream still
    // normalises the peers' blocks of W_T), so a steady-state iteration starts directly with the solve.
    if (iters_enqueued == 0) {
        gram(Wblk, m_loc, false, G_w.ptr, RCPPML_B200_SEC_GRAM_H, sharded);   // :491
        prepare_solver(G_w.ptr, cfg.L2_H, RCPPML_B200_SEC_GRAM_H);            // :506
    }
    join_side_stream();                                                     // W_T fully normalised (peer blocks)
    solve(0, warm, RCPPML_B200_SEC_SOLVE_H);                                // :516-535 (+ :636 upper bound)
    scale_finalize(RCPPML_B200_SEC_SCALE_H, sharded);                       // :644 (p2p: also the barrier)
    // ---- W update (fit_cpu.hpp:713-893)
    if (ucast) fork_side_stream();                                          // d is final from here on
    gram(Hblk, n_loc, normalize, G_h.ptr, RCPPML_B200_SEC_GRAM_W, sharded, repl); // :644 (normalise) + :715
    if (ucast) normalize_peer_blocks(H.ptr, n, col_begin, col_begin + n_loc, normalize);   // side stream, behind the Gram in the queues
    if (sharded && !p2p) allgather_rows(H.ptr, col_cuts, n_pad, RCPPML_B200_SEC_COMM);
    prepare_solver(G_h.ptr, cfg.L2_W, RCPPML_B200_SEC_GRAM_W);              // :738
    join_side_stream();
    solve(1, warm, RCPPML_B200_SEC_SOLVE_W);                                // :748-767 (+ :884)
    scale_finalize(RCPPML_B200_SEC_SCALE_W, sharded);                       // :892
    // ---- loss (fit_cpu.hpp:1729-1809): Gram of the new W_T doubles as next iteration's gram_H
    if (ucast) fork_side_stream();                                          // d is final from here on
    sec_begin(RCPPML_B200_SEC_LOSS);
    const bool was = profiling; profiling = false;                         // nested section: account under LOSS
    gram(Wblk, m_loc, normalize, G_w.ptr, RCPPML_B200_SEC_LOSS, sharded, repl);   // :892 (normalise) + :1735
    profiling = was;
    if (ucast) normalize_peer_blocks(W_T.ptr, m, row_begin, row_begin + m_loc, normalize);   // side stream, behind the Gram in the queues
    loss(RCPPML_B200_SEC_LOSS);
    sec_end(RCPPML_B200_SEC_LOSS);
    if (sharded && !p2p) allgather_rows(W_T.ptr, row_cuts, m_pad, RCPPML_B200_SEC_COMM);
    prepare_solver(G_w.ptr, cfg.L2_H, RCPPML_B200_SEC_GRAM_H);              // next iteration's :506 (no-op once stopped)
    if (capturing) join_side_stream();                                      // a captured graph has no dangling branch
    ++iters_enqueued;
}

// ---- explicit user mask (nmf/masked_nnls.hpp) ---------------------------------------------------
// mask: CSC pattern m × n of the masked entries (the reference masks entries whose stored value is
// non-zero, masked_nnls.hpp:116-118 — callers pass that pattern). The transposed pattern is built on
// the device (fit_cpu.hpp:273-278). Pass nnz = 0 to clear.
void Engine::set_mask(int64_t mnnz, const int* mask_ptr, const int* mask_idx) {
    use_device();
    B200_REQUIRE(matrix_ready, "set_mask: set the matrix first");
    has_mask = false;
    if (mnnz <= 0) return;
    // Every rank receives the WHOLE pattern (index arrays only; masks are small next to A), transposes it on the
    // device and — sharded — keeps two contiguous slices: columns J of the pattern (H half-step) and columns I of its
    // transpose (= rows I, W half-step), pointers rebased; exactly the two operands it holds of A itself.
    DeviceBuffer<int> fp, fi, tp, ti;
    DeviceBuffer<float> ones, onesT;
    DeviceBuffer<int>& Fp = (world == 1) ? Mp : fp;
    DeviceBuffer<int>& Fi = (world == 1) ? Mi : fi;
    DeviceBuffer<int>& Tp = (world == 1) ? MTp : tp;
    DeviceBuffer<int>& Ti = (world == 1) ? MTi : ti;
    Fp.ensure(static_cast<size_t>(n) + 1);
    Fi.ensure(mnnz + 4);
    B200_CUDA_CHECK(cudaMemcpyAsync(Fp.ptr, mask_ptr, (static_cast<size_t>(n) + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(Fi.ptr, mask_idx, mnnz * sizeof(int), cudaMemcpyHostToDevice, stream));
    ones.ensure(mnnz + 4);
    B200_CUDA_CHECK(cudaMemsetAsync(ones.ptr, 0, (mnnz + 4) * sizeof(float), stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    transpose_csc(Fp.ptr, Fi.ptr, ones.ptr, n, m, mnnz, Tp, Ti, onesT, 0);
    if (world > 1) {
        auto slice = [&](const DeviceBuffer<int>& sp, const DeviceBuffer<int>& si, int first, int count, DeviceBuffer<int>& dp,
                         DeviceBuffer<int>& di) {
            int ends[2] = {0, 0};
            B200_CUDA_CHECK(cudaMemcpyAsync(&ends[0], sp.ptr + first, sizeof(int), cudaMemcpyDeviceToHost, stream));
            B200_CUDA_CHECK(cudaMemcpyAsync(&ends[1], sp.ptr + first + count, sizeof(int), cudaMemcpyDeviceToHost, stream));
            B200_CUDA_CHECK(cudaStreamSynchronize(stream));
            const int64_t cnt = static_cast<int64_t>(ends[1]) - ends[0];
            dp.ensure(static_cast<size_t>(std::max(count, 1)) + 1);
            di.ensure(std::max<int64_t>(cnt, 1) + 4);
            B200_CUDA_CHECK(cudaMemcpyAsync(dp.ptr, sp.ptr + first, (static_cast<size_t>(count) + 1) * sizeof(int), cudaMemcpyDeviceToDevice, stream));
            rebase_int_kernel<<<(count + 1 + 255) / 256, 256, 0, stream>>>(dp.ptr, count + 1, ends[0]);
            if (cnt > 0) B200_CUDA_CHECK(cudaMemcpyAsync(di.ptr, si.ptr + ends[0], cnt * sizeof(int), cudaMemcpyDeviceToDevice, stream));
            B200_CUDA_CHECK(cudaGetLastError());
        };
        slice(fp, fi, col_begin, n_loc, Mp, Mi);
        slice(tp, ti, row_begin, m_loc, MTp, MTi);
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    mask_nnz = mnnz;
    has_mask = true;
}

void Engine::masked_solve(int which, bool warm, const float* G, int sec) {
    MaskedParams p{};
    const bool h = (which == 0);
    p.colptr = h ? Ap.ptr : Atp.ptr;
    p.rowidx = h ? Ai.ptr : Ati.ptr;
    p.vals = h ? Ax.ptr : Atx.ptr;
    p.mptr = h ? Mp.ptr : MTp.ptr;
    p.midx = h ? Mi.ptr : MTi.ptr;
    p.F = h ? W_T.ptr : H.ptr;
    p.X = h ? H.ptr : W_T.ptr;
    p.G = G;
    p.ncols = h ? n_loc : m_loc;
    p.col_offset = h ? col_begin : row_begin;
    p.npeers = 0;
    p.mcX = nullptr;
    if (peers_ready && mc_ready) {
        if (mc_mode == 2) p.mcX = mc_alias(p.X);
    } else if (peers_ready) {
        for (int r = 0; r < world; ++r)
            if (r != rank) p.peerX[p.npeers++] = h ? peer_H[r] : peer_W[r];
    }
    p.k = k;
    p.L1 = h ? cfg.L1_H : cfg.L1_W;
    p.L2 = h ? cfg.L2_H : cfg.L2_W;
    p.ub = h ? cfg.ub_H : cfg.ub_W;
    p.cd_tol = cfg.cd_tol;
    p.inv_k = 1.0f / static_cast<float>(k);
    p.cd_maxit = cfg.cd_maxit;
    p.nonneg = h ? cfg.nonneg_H : cfg.nonneg_W;
    p.warm = warm ? 1 : 0;
    p.solver = (cfg.solver_mode == 1) ? 1 : 0;       // masked_solve_col tests `== 1` (masked_nnls.hpp:56)
    p.norm_type = cfg.norm_type;
    p.work_counter = counters.ptr + 4 + which;
    p.partials = solve_partials.ptr;
    p.state = state.ptr;
    p.sweep_counter = sweep_counter.ptr;
    int grid = 0;
    launch_masked(KP, p, num_sms, stream, &grid);
    B200_REQUIRE(static_cast<size_t>(grid) * (KP + 1) <= solve_partials.count, "partials buffer too small");
    last_solve_grid = grid;
    sec_begin(sec);
    B200_CUDA_CHECK(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));
    launch_masked(KP, p, num_sms, stream);
    launches[sec] += 1;
    sec_end(sec);
}

// fit_cpu.hpp with use_mask: :560-564 (H), :799-810 (W), :1686-1691 (loss). Sharded like enqueue_iteration: every rank
// solves its column block of H and its row block of W_T with its slices of A and of the mask pattern; Grams, row
// sums and the explicit loss are all-reduced in fp64.
void Engine::enqueue_iteration_masked() {
    const bool warm = iters_enqueued > 0;
    const bool normalize = cfg.norm_type != 2;
    const bool sharded = world > 1;
    const bool p2p = sharded && peers_ready;
    const bool repl = p2p && mc_ready && mc_mode == 1;                      // multicast replication by the Gram kernel
    const bool ucast = p2p && !repl;                                        // the solve kernel replicates (unicast peer stores or
                                                                            // multicast), every rank re-normalises the peers' blocks
    float* Hblk = H.ptr + static_cast<size_t>(col_begin) * KP;
    float* Wblk = W_T.ptr + static_cast<size_t>(row_begin) * KP;
    if (iters_enqueued == 0) gram(Wblk, m_loc, false, G_w.ptr, RCPPML_B200_SEC_GRAM_H, sharded);   // :562 rebuilt unmodified
    join_side_stream();
    masked_solve(0, warm, G_w.ptr, RCPPML_B200_SEC_SOLVE_H);
    scale_finalize(RCPPML_B200_SEC_SCALE_H, sharded);
    if (ucast) fork_side_stream();                                          // d is final from here on
    gram(Hblk, n_loc, normalize, G_h.ptr, RCPPML_B200_SEC_GRAM_W, sharded, repl);                       // :644 + :801
    if (ucast) normalize_peer_blocks(H.ptr, n, col_begin, col_begin + n_loc, normalize);   // side stream, behind the Gram in the queues
    if (sharded && !p2p) allgather_rows(H.ptr, col_cuts, n_pad, RCPPML_B200_SEC_COMM);
    join_side_stream();
    masked_solve(1, warm, G_h.ptr, RCPPML_B200_SEC_SOLVE_W);
    scale_finalize(RCPPML_B200_SEC_SCALE_W, sharded);
    if (ucast) fork_side_stream();                                          // d is final from here on
    sec_begin(RCPPML_B200_SEC_LOSS);
    const bool was = profiling; profiling = false;
    gram(Wblk, m_loc, normalize, G_w.ptr, RCPPML_B200_SEC_LOSS, sharded, repl);                         // normalise + next gram_H
    profiling = was;
    if (ucast) normalize_peer_blocks(W_T.ptr, m, row_begin, row_begin + m_loc, normalize);   // side stream, behind the Gram in the queues
    if (sharded && !p2p) allgather_rows(W_T.ptr, row_cuts, m_pad, RCPPML_B200_SEC_COMM);
    join_side_stream();                                                     // the loss reads every row of W_T
    const int lgrid = num_sms * 4;
    masked_loss_kernel<<<lgrid, 256, 0, stream>>>(Ap.ptr, Ai.ptr, Ax.ptr, Mp.ptr, Mi.ptr, n_loc, col_begin, KP, k, W_T.ptr, H.ptr,
                                                  d.ptr, loss_partials.ptr, &state.ptr->stop);
    // fixed-order sum of the per-CTA partials, then (sharded) over the ranks' column blocks
    sum_partials_kernel<<<1, dim3(32, 8), 0, stream>>>(loss_partials.ptr, lgrid, 1, red_small.ptr, &state.ptr->stop);
    if (sharded) allreduce_f64(red_small.ptr, 1);
    masked_loss_finalize_kernel<<<1, 32, 0, stream>>>(red_small.ptr, 1, cfg.tol, cfg.patience, loss_hist.ptr,
                                                      static_cast<int>(loss_hist.count), state.ptr);
    launches[RCPPML_B200_SEC_LOSS] += 3;
    sec_end(RCPPML_B200_SEC_LOSS);
    ++iters_enqueued;
}

// ---- speckled-mask cross-validation (nmf/fit_cv.hpp) -----------------------------------------------
void Engine::cv_solve(int which, int sec) {
    CvParams p{};
    const bool h = (which == 0);
    p.colptr = h ? Ap.ptr : Atp.ptr;
    p.rowidx = h ? Ai.ptr : Ati.ptr;
    p.vals = h ? Ax.ptr : Atx.ptr;
    p.F = h ? W_T.ptr : H.ptr;
    p.X = h ? H.ptr : W_T.ptr;
    p.G = M1.ptr;
    p.ncols = h ? n_loc : m_loc;
    p.nrows = h ? m : n;
    p.col_offset = h ? col_begin : row_begin;
    p.npeers = 0;
    p.mcX = nullptr;
    if (peers_ready && mc_ready) {
        if (mc_mode == 2) p.mcX = mc_alias(p.X);
    } else if (peers_ready) {
        for (int r = 0; r < world; ++r)
            if (r != rank) p.peerX[p.npeers++] = h ? peer_H[r] : peer_W[r];
    }
    p.k = k;
    p.transposed = h ? 0 : 1;
    p.mask_zeros = cv.mask_zeros;
    p.seed = cv_seed_state;
    p.threshold = cv_threshold;
    p.holdout_enabled = cv_inv_prob != 0;
    p.L1 = h ? cfg.L1_H : cfg.L1_W;
    p.ub = h ? cfg.ub_H : cfg.ub_W;
    p.cd_maxit = cfg.cd_maxit;
    p.nonneg = h ? cfg.nonneg_H : cfg.nonneg_W;
    p.solver = (cfg.solver_mode == 1) ? 1 : 0;              // fit_cv.hpp:461 tests `== 1`
    p.norm_type = cfg.norm_type;
    p.want_cross = h ? 0 : 1;
    p.work_counter = counters.ptr + 6 + which;
    p.partials = solve_partials.ptr;
    p.state = state.ptr;
    int grid = 0;
    launch_cv(KP, p, num_sms, stream, &grid);
    B200_REQUIRE(static_cast<size_t>(grid) * (KP + 1) <= solve_partials.count, "partials buffer too small");
    last_solve_grid = grid;
    sec_begin(sec);
    B200_CUDA_CHECK(cudaMemsetAsync(p.work_counter, 0, sizeof(int), stream));
    launch_cv(KP, p, num_sms, stream);
    launches[sec] += 1;
    sec_end(sec);
}

void Engine::enqueue_iteration_cv() {
    const bool normalize = cfg.norm_type != 2;
    const bool sharded = world > 1;
    const bool p2p = sharded && peers_ready;
    const bool repl = p2p && mc_ready && mc_mode == 1;                      // multicast replication by the Gram kernel
    const bool ucast = p2p && !repl;                                        // the solve kernel replicates (unicast peer stores or
                                                                            // multicast), every rank re-normalises the peers' blocks
    float* Hblk = H.ptr + static_cast<size_t>(col_begin) * KP;
    float* Wblk = W_T.ptr + static_cast<size_t>(row_begin) * KP;
    const int ge = (KP * KP + 255) / 256;
    if (iters_enqueued == 0) gram(Wblk, m_loc, false, G_w.ptr, RCPPML_B200_SEC_GRAM_H, sharded);   // fit_cv.hpp:413
    cv_prepare_gram_kernel<<<ge, 256, 0, stream>>>(G_w.ptr, KP, k, cfg.L2_H, M1.ptr, &state.ptr->stop);   // :414-417
    join_side_stream();
    cv_solve(0, RCPPML_B200_SEC_SOLVE_H);                                                    // :431-476, :528
    scale_finalize(RCPPML_B200_SEC_SCALE_H, sharded);                                        // :536-548
    if (ucast) fork_side_stream();                                          // d is final from here on
    gram(Hblk, n_loc, normalize, G_h.ptr, RCPPML_B200_SEC_GRAM_W, sharded, repl);                  // :570 (= G_H_saved)
    if (ucast) normalize_peer_blocks(H.ptr, n, col_begin, col_begin + n_loc, normalize);   // side stream, behind the Gram in the queues
    if (sharded && !p2p) allgather_rows(H.ptr, col_cuts, n_pad, RCPPML_B200_SEC_COMM);
    cv_prepare_gram_kernel<<<ge, 256, 0, stream>>>(G_h.ptr, KP, k, cfg.L2_W, M1.ptr, &state.ptr->stop);   // :578-581
    join_side_stream();
    cv_solve(1, RCPPML_B200_SEC_SOLVE_W);                                                    // :598-735, :843
    scale_finalize(RCPPML_B200_SEC_SCALE_W, sharded);                                        // :849-858 (+ cross term)
    if (ucast) fork_side_stream();                                          // d is final from here on
    sec_begin(RCPPML_B200_SEC_LOSS);
    const bool was = profiling; profiling = false;
    gram(Wblk, m_loc, normalize, G_w.ptr, RCPPML_B200_SEC_LOSS, sharded, repl);                    // normalise + :1518
    profiling = was;
    if (ucast) normalize_peer_blocks(W_T.ptr, m, row_begin, row_begin + m_loc, normalize);   // side stream, behind the Gram in the queues
    if (sharded && !p2p) allgather_rows(W_T.ptr, row_cuts, m_pad, RCPPML_B200_SEC_COMM);
    join_side_stream();                                                     // the test loss reads every row of W_T
    const int lgrid = num_sms * 4;
    cv_test_loss_kernel<<<lgrid, 256, 0, stream>>>(Ap.ptr, Ai.ptr, Ax.ptr, n_loc, col_begin, m, KP, k, cv.mask_zeros, cv_seed_state,
                                                   cv_threshold, cv_inv_prob != 0, W_T.ptr, H.ptr, d.ptr,
                                                   loss_partials.ptr, &state.ptr->stop);
    // {sum, count} of the held-out entries: fixed-order over the CTAs, then (sharded) over the ranks' column blocks
    sum_partials_kernel<<<1, dim3(32, 8), 0, stream>>>(loss_partials.ptr, lgrid, 2, red_gram.ptr, &state.ptr->stop);
    if (sharded) allreduce_f64(red_gram.ptr, 2);
    const long long total_entries = cv.mask_zeros ? nnz_global : static_cast<long long>(m) * n;
    cv_loss_finalize_kernel<<<1, 256, 0, stream>>>(G_w.ptr, G_h.ptr, d.ptr, KP, k, red_small.ptr + KP, red_gram.ptr,
                                                   1, trAtA, total_entries, cfg.tol, cv.cv_patience, loss_hist.ptr,
                                                   test_hist.ptr, static_cast<int>(loss_hist.count), state.ptr,
                                                   cv_state.ptr);
    launches[RCPPML_B200_SEC_GRAM_H] += 1; launches[RCPPML_B200_SEC_GRAM_W] += 1; launches[RCPPML_B200_SEC_LOSS] += 3;
    sec_end(RCPPML_B200_SEC_LOSS);
    ++iters_enqueued;
}

void Engine::fit_cv(const rcppml_b200_config& c, const rcppml_b200_cv_config& cvc) {
    use_device();
    B200_REQUIRE(!has_mask, "fit_cv: a user mask together with the speckled mask is not supported");
    B200_REQUIRE(cvc.holdout_fraction >= 0.f && cvc.holdout_fraction < 1.f, "holdout_fraction must be in [0, 1)");   // core/config.hpp:428
    B200_REQUIRE(c.max_iter > 0, "max_iter must be positive");
    cv = cvc;
    if (cv.cv_patience < 0) cv.cv_patience = 5;
    // nmf/speckled_cv.hpp:117-127: seed remap, inv_prob from the FLOAT fraction (0.1f -> 9, i.e. 11.1 % held out)
    const uint32_t eff = cv.cv_seed != 0 ? cv.cv_seed : cv.seed;                 // core/config.hpp:416-418
    cv_seed_state = (eff == 0) ? 12345ULL : static_cast<unsigned long long>(eff);
    const double hf = static_cast<double>(cv.holdout_fraction);
    cv_inv_prob = hf > 0 ? static_cast<unsigned long long>(1.0 / hf) : 0ULL;
    cv_threshold = cv_inv_prob ? (0xFFFFFFFFFFFFFFFFULL / cv_inv_prob) : 0ULL;
    begin_fit(c);
    test_hist.ensure(loss_hist.count);
    cv_state.ensure(1);
    CvState s0{};
    s0.prev_conv_loss = 3.402823466e+38f;
    s0.best_test_loss = 3.402823466e+38f;
    B200_CUDA_CHECK(cudaMemcpyAsync(cv_state.ptr, &s0, sizeof(CvState), cudaMemcpyHostToDevice, stream));
    cv_active = true;
    iterate(cfg.max_iter);
    cv_active = false;
    const long long th = static_cast<long long>(n) * KP;
    cv_absorb_d_kernel<<<static_cast<unsigned>((th + 255) / 256), 256, 0, stream>>>(H.ptr, n, KP, d.ptr);   // :1639-1641
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

void Engine::get_cv_result(rcppml_b200_cv_result* out) {
    use_device();
    B200_REQUIRE(cv_state.ptr != nullptr, "no cross-validation fit has run");
    CvState s{};
    B200_CUDA_CHECK(cudaMemcpy(&s, cv_state.ptr, sizeof(CvState), cudaMemcpyDeviceToHost));
    out->train_loss = s.train_loss; out->test_loss = s.test_loss; out->best_test_loss = s.best_test_loss;
    out->best_iter = s.best_iter; out->n_test = s.n_test;
}

void Engine::begin_fit(const rcppml_b200_config& c) {
    use_device();
    B200_REQUIRE(matrix_ready && factors_ready, "begin_fit: matrix and factors must be set first");
    normalize_cfg(c);
    B200_REQUIRE(world == 1 || comm_ready() || peers_ready, "begin_fit: communicator not initialised");
    iters_enqueued = 0;
    drop_iteration_graph();                                                 // new configuration / buffers
    graphs_enabled = true;
    if (const char* env = std::getenv("RCPPML_B200_GRAPH")) graphs_enabled = std::atoi(env) != 0;
    build_panels();
    loss_hist.ensure(static_cast<size_t>(std::max(cfg.max_iter, 1024)));
    // Every buffer the loop touches exists before the first iteration is enqueued: a cudaMalloc inside the loop would
    // synchronise the device against peers that are already spinning in an exchange kernel (sharded fits).
    carry.ensure(static_cast<size_t>(std::max(std::max(n_loc, m_loc), 1)) * KP);
    loss_partials.ensure(static_cast<size_t>(num_sms) * 4 * 2);
    DevState s0{};
    s0.prev_loss = 3.402823466e+38f;                                        // fit_cpu.hpp:281
    B200_CUDA_CHECK(cudaMemcpyAsync(state.ptr, &s0, sizeof(DevState), cudaMemcpyHostToDevice, stream));
    B200_CUDA_CHECK(cudaMemsetAsync(sweep_counter.ptr, 0, sizeof(unsigned long long), stream));
    std::vector<float> ones(KP, 1.f);                                       // d = 1 (fit_cpu.hpp:198)
    B200_CUDA_CHECK(cudaMemcpyAsync(d.ptr, ones.data(), KP * sizeof(float), cudaMemcpyHostToDevice, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    for (int s = 0; s < RCPPML_B200_NUM_SECTIONS; ++s) { prof_ms[s] = 0.0; launches[s] = 0; prof_used[s] = 0; }
    // Peer-memory loop: no rank may push solved columns into a replica whose owner is still initialising it.
    // One exchange (a k-independent barrier on the stream) orders every rank's set-up before any first solve.
    if (world > 1 && peers_ready) {
        red_small.ensure(KP + 1);
        allreduce_f64(red_small.ptr, 1);
    }
    fit_active = true;
    loop_ms = 0.0;
}

void Engine::drop_iteration_graph() {
    if (iter_graph) cudaGraphExecDestroy(iter_graph);
    iter_graph = nullptr;
}

// Capture one steady-state iteration (warm start, Gram of W_T carried over) into a graph. Every launcher has
// run once by now (iteration 0), so no attribute / occupancy query or allocation happens inside the capture.
void Engine::capture_iteration_graph() {
    const auto before = launches;
    const int it0 = iters_enqueued;
    cudaGraph_t graph = nullptr;
    join_side_stream();                                                     // nothing of a plain iteration dangles into the capture
    if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        graphs_enabled = false;
        return;
    }
    bool ok = true;
    capturing = true;
    try {
        enqueue_iteration();
    } catch (...) {
        ok = false;
    }
    capturing = false;
    if (cudaStreamEndCapture(stream, &graph) != cudaSuccess || !graph) ok = false;
    iters_enqueued = it0;                                                   // nothing has executed
    for (int s = 0; s < RCPPML_B200_NUM_SECTIONS; ++s) graph_launches[s] = launches[s] - before[s];
    launches = before;
    if (ok && cudaGraphInstantiate(&iter_graph, graph, 0) != cudaSuccess) { iter_graph = nullptr; ok = false; }
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {                                                              // fall back to plain launches
        cudaGetLastError();
        drop_iteration_graph();
        graphs_enabled = false;
    }
}

void Engine::iterate(int n_iters) {
    use_device();
    B200_REQUIRE(fit_active, "iterate: call begin_fit first");
    B200_CUDA_CHECK(cudaEventRecord(ev_loop_begin, stream));
    // The host never waits inside the loop: convergence is decided on the device (DevState) and
    // kernels enqueued after `stop` return immediately. Every 8 iterations the host peeks at the
    // flag only to avoid enqueuing a long tail of no-op launches.
    for (int it = 0; it < n_iters; ++it) {
        // (sharded fits: only the peer-memory loop — its exchange kernels carry no per-call argument; the NCCL loop
        // is launched plainly)
        const bool graphable = graphs_enabled && !cv_active && !has_mask && (world == 1 || peers_ready) && !profiling &&
                               iters_enqueued >= 1;
        if (graphable && !iter_graph) capture_iteration_graph();
        if (graphable && iter_graph) {
            B200_CUDA_CHECK(cudaGraphLaunch(iter_graph, stream));
            for (int s = 0; s < RCPPML_B200_NUM_SECTIONS; ++s) launches[s] += graph_launches[s];
            ++iters_enqueued;
        } else if (cv_active) enqueue_iteration_cv(); else if (has_mask) enqueue_iteration_masked(); else enqueue_iteration();
        // (sharded fits poll as well, less often: a peer-memory exchange that timed out sets comm_error + stop)
        if (((it & 7) == 7 && (cfg.tol > 0.f || cv_active)) || ((it & 31) == 31 && world > 1)) {
            B200_CUDA_CHECK(cudaMemcpyAsync(h_state, state.ptr, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
            B200_CUDA_CHECK(cudaStreamSynchronize(stream));
            if (h_state->stop || h_state->comm_error) break;
        }
    }
    join_side_stream();
    B200_CUDA_CHECK(cudaEventRecord(ev_loop_end, stream));
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    float ms = 0.f;
    B200_CUDA_CHECK(cudaEventElapsedTime(&ms, ev_loop_begin, ev_loop_end));
    loop_ms = ms;
    phase_ms[3] = ms;
    collect_profile();
}

void Engine::half_step_only(const rcppml_b200_config& c, int which, bool warm, bool normalize_after) {
    use_device();
    B200_REQUIRE(matrix_ready && factors_ready, "half_step: matrix and factors must be set first");
    B200_REQUIRE(world == 1, "half_step: single-GPU diagnostic entry");
    normalize_cfg(c);
    drop_iteration_graph();
    fit_active = false;
    build_panels();
    DevState s0{};
    s0.prev_loss = 3.402823466e+38f;
    B200_CUDA_CHECK(cudaMemcpyAsync(state.ptr, &s0, sizeof(DevState), cudaMemcpyHostToDevice, stream));
    const bool h = (which == 0);
    gram(h ? W_T.ptr : H.ptr, h ? m : n, false, h ? G_w.ptr : G_h.ptr, RCPPML_B200_SEC_GRAM_H);
    prepare_solver(h ? G_w.ptr : G_h.ptr, h ? cfg.L2_H : cfg.L2_W, RCPPML_B200_SEC_GRAM_H);
    solve(which, warm, h ? RCPPML_B200_SEC_SOLVE_H : RCPPML_B200_SEC_SOLVE_W);
    scale_finalize(RCPPML_B200_SEC_SCALE_H);
    if (normalize_after && cfg.norm_type != 2)
        gram(h ? H.ptr : W_T.ptr, h ? n : m, true, h ? G_h.ptr : G_w.ptr, RCPPML_B200_SEC_GRAM_W);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

void Engine::get_result(rcppml_b200_result* out) {
    use_device();
    DevState s{};
    B200_CUDA_CHECK(cudaMemcpyAsync(&s, state.ptr, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
    unsigned long long sw = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&sw, sweep_counter.ptr, sizeof(sw), cudaMemcpyDeviceToHost, stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    out->iterations = s.iter;
    out->converged = s.converged;
    out->train_loss = s.train_loss;
    out->final_tol = s.final_tol;
    out->status = s.chol_fail ? 1 : (s.comm_error ? 2 : 0);
    int total = 0;
    for (int i = 0; i < RCPPML_B200_NUM_SECTIONS; ++i) total += launches[i];
    out->gpu_launches = total;
    out->loop_ms = loop_ms;
    cd_sweeps = sw;
}

}  // namespace b200

// =============================================================================================
// C ABI (include/rcppml_gpu.h, part 2)
// =============================================================================================
using b200::Engine;

#define B200_API_BEGIN try {
#define B200_API_END                                                      \
    return 0;                                                             \
    }                                                                     \
    catch (const std::exception& ex) {                                    \
        b200::g_last_error = ex.what();                                   \
        return -1;                                                        \
    }                                                                     \
    catch (...) {                                                         \
        b200::g_last_error = "unknown error";                             \
        return -1;                                                        \
    }

extern "C" {

const char* rcppml_b200_last_error(void) { return b200::g_last_error.c_str(); }

int rcppml_b200_engine_create(rcppml_b200_engine** out, int device) {
    B200_API_BEGIN
    B200_REQUIRE(out != nullptr, "engine_create: null output");
    int count = 0;
    B200_CUDA_CHECK(cudaGetDeviceCount(&count));
    B200_REQUIRE(device >= 0 && device < count, "engine_create: no such CUDA device");
    *out = new rcppml_b200_engine(device);
    B200_API_END
}
void rcppml_b200_engine_destroy(rcppml_b200_engine* e) { delete e; }

int rcppml_b200_set_matrix_f32(rcppml_b200_engine* e, int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx, const float* values) {
    B200_API_BEGIN e->impl.set_matrix_host<float>(m, n, nnz, col_ptr, row_idx, values); B200_API_END
}
int rcppml_b200_set_matrix_f64(rcppml_b200_engine* e, int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx, const double* values) {
    B200_API_BEGIN e->impl.set_matrix_host<double>(m, n, nnz, col_ptr, row_idx, values); B200_API_END
}
int rcppml_b200_set_matrix_with_transpose_f32(rcppml_b200_engine* e, int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx,
                                              const float* values, const int* t_col_ptr, const int* t_row_idx, const float* t_values) {
    B200_API_BEGIN e->impl.set_matrix_host_with_transpose<float>(m, n, nnz, col_ptr, row_idx, values, t_col_ptr, t_row_idx, t_values); B200_API_END
}
int rcppml_b200_set_matrix_with_transpose_f64(rcppml_b200_engine* e, int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx,
                                              const double* values, const int* t_col_ptr, const int* t_row_idx, const double* t_values) {
    B200_API_BEGIN e->impl.set_matrix_host_with_transpose<double>(m, n, nnz, col_ptr, row_idx, values, t_col_ptr, t_row_idx, t_values); B200_API_END
}
int rcppml_b200_set_matrix_synthetic(rcppml_b200_engine* e, int m, int n_local, int col_begin, double density, uint64_t seed) {
    B200_API_BEGIN e->impl.set_matrix_synthetic(m, n_local, col_begin, density, seed); B200_API_END
}
int rcppml_b200_set_matrix_synthetic_sharded(rcppml_b200_engine* e, int m, int n, double density, uint64_t seed) {
    B200_API_BEGIN e->impl.set_matrix_synthetic_sharded(m, n, density, seed); B200_API_END
}
int rcppml_b200_set_matrix_sharded_f32(rcppml_b200_engine* e, int m, int n, const int* cb_ptr, const int* cb_idx,
                                       const float* cb_val, const int* rb_ptr, const int* rb_idx, const float* rb_val) {
    B200_API_BEGIN e->impl.set_matrix_sharded<float>(m, n, cb_ptr, cb_idx, cb_val, rb_ptr, rb_idx, rb_val); B200_API_END
}
int rcppml_b200_set_matrix_sharded_with_transpose_f32(rcppml_b200_engine* e, int m, int n, const int* cb_ptr, const int* cb_idx,
                                                      const float* cb_val, const int* tb_ptr, const int* tb_idx, const float* tb_val) {
    B200_API_BEGIN e->impl.set_matrix_sharded_with_transpose<float>(m, n, cb_ptr, cb_idx, cb_val, tb_ptr, tb_idx, tb_val); B200_API_END
}
int rcppml_b200_set_partition(rcppml_b200_engine* e, const int* col_cuts, const int* row_cuts) {
    B200_API_BEGIN e->impl.set_partition(col_cuts, row_cuts); B200_API_END
}
int rcppml_b200_factor_checksum(rcppml_b200_engine* e, uint64_t* out3) {
    B200_API_BEGIN
    B200_REQUIRE(out3 != nullptr, "factor_checksum: null output");
    unsigned long long h[3] = {0, 0, 0};
    e->impl.factor_checksum(h);
    for (int i = 0; i < 3; ++i) out3[i] = static_cast<uint64_t>(h[i]);
    B200_API_END
}
int rcppml_b200_get_shard(rcppml_b200_engine* e, int* col_begin, int* n_loc, int* row_begin, int* m_loc, int64_t* nnz_global) {
    B200_API_BEGIN
    if (col_begin) *col_begin = e->impl.col_begin;
    if (n_loc) *n_loc = e->impl.n_loc;
    if (row_begin) *row_begin = e->impl.row_begin;
    if (m_loc) *m_loc = e->impl.m_loc;
    if (nnz_global) *nnz_global = e->impl.nnz_global;
    B200_API_END
}
static void copy_csc(Engine& E, const int* dp, const int* di, const float* dx, int ncols, int64_t cnt, int* p, int* i, float* x) {
    E.use_device();
    if (p) B200_CUDA_CHECK(cudaMemcpy(p, dp, (static_cast<size_t>(ncols) + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    if (i && cnt) B200_CUDA_CHECK(cudaMemcpy(i, di, cnt * sizeof(int), cudaMemcpyDeviceToHost));
    if (x && cnt) B200_CUDA_CHECK(cudaMemcpy(x, dx, cnt * sizeof(float), cudaMemcpyDeviceToHost));
}
int rcppml_b200_get_matrix(rcppml_b200_engine* e, int64_t* nnz, int* col_ptr, int* row_idx, float* values) {
    B200_API_BEGIN
    B200_REQUIRE(e->impl.matrix_ready, "no matrix");
    if (nnz) *nnz = e->impl.nnz;
    copy_csc(e->impl, e->impl.Ap.ptr, e->impl.Ai.ptr, e->impl.Ax.ptr, e->impl.n_loc, e->impl.nnz, col_ptr, row_idx, values);
    B200_API_END
}
int rcppml_b200_get_matrix_t(rcppml_b200_engine* e, int* col_ptr, int* row_idx, float* values) {
    B200_API_BEGIN
    B200_REQUIRE(e->impl.matrix_ready, "no matrix");
    copy_csc(e->impl, e->impl.Atp.ptr, e->impl.Ati.ptr, e->impl.Atx.ptr, e->impl.m_loc, e->impl.nnz_w, col_ptr, row_idx, values);
    B200_API_END
}
int rcppml_b200_comm_ipc_export(rcppml_b200_engine* e, char* handles192) {
    B200_API_BEGIN e->impl.comm_ipc_export(handles192); B200_API_END
}
int rcppml_b200_comm_ipc_import(rcppml_b200_engine* e, const char* all_handles) {
    B200_API_BEGIN e->impl.comm_ipc_import(all_handles); B200_API_END
}
int rcppml_b200_set_mask(rcppml_b200_engine* e, int64_t mask_nnz, const int* mask_col_ptr, const int* mask_row_idx) {
    B200_API_BEGIN e->impl.set_mask(mask_nnz, mask_col_ptr, mask_row_idx); B200_API_END
}
int rcppml_b200_set_factors_f32(rcppml_b200_engine* e, int k, const float* W_T, const float* H) {
    B200_API_BEGIN e->impl.set_factors_host<float>(k, W_T, H); B200_API_END
}
int rcppml_b200_set_factors_f64(rcppml_b200_engine* e, int k, const double* W_T, const double* H) {
    B200_API_BEGIN e->impl.set_factors_host<double>(k, W_T, H); B200_API_END
}
int rcppml_b200_init_factors(rcppml_b200_engine* e, int k, uint32_t seed, int h_col_begin) {
    B200_API_BEGIN e->impl.init_factors(k, seed, h_col_begin); B200_API_END
}
int rcppml_b200_get_factors_f32(rcppml_b200_engine* e, float* W_T, float* H, float* d) {
    B200_API_BEGIN e->impl.get_factors_host<float>(W_T, H, d); B200_API_END
}
int rcppml_b200_get_factors_f64(rcppml_b200_engine* e, double* W_T, double* H, double* d) {
    B200_API_BEGIN e->impl.get_factors_host<double>(W_T, H, d); B200_API_END
}
int rcppml_b200_set_factor_blocks_f32(rcppml_b200_engine* e, int k, const float* W_blk, const float* H_blk) {
    B200_API_BEGIN e->impl.set_factor_blocks_host<float>(k, W_blk, H_blk); B200_API_END
}
int rcppml_b200_get_factor_blocks_f32(rcppml_b200_engine* e, float* W_blk, float* H_blk, float* d) {
    B200_API_BEGIN e->impl.get_factor_blocks_host<float>(W_blk, H_blk, d); B200_API_END
}
int rcppml_b200_begin_fit(rcppml_b200_engine* e, const rcppml_b200_config* cfg) {
    B200_API_BEGIN e->impl.begin_fit(*cfg); B200_API_END
}
int rcppml_b200_iterate(rcppml_b200_engine* e, int n_iters) {
    B200_API_BEGIN e->impl.iterate(n_iters); B200_API_END
}
int rcppml_b200_fit(rcppml_b200_engine* e, const rcppml_b200_config* cfg) {
    B200_API_BEGIN
    B200_REQUIRE(cfg->max_iter > 0, "max_iter must be positive");        // core/config.hpp:424
    e->impl.begin_fit(*cfg);
    e->impl.iterate(cfg->max_iter);
    B200_API_END
}
int rcppml_b200_fit_cv(rcppml_b200_engine* e, const rcppml_b200_config* cfg, const rcppml_b200_cv_config* cv) {
    B200_API_BEGIN e->impl.fit_cv(*cfg, *cv); B200_API_END
}
int rcppml_b200_get_cv_result(rcppml_b200_engine* e, rcppml_b200_cv_result* out) {
    B200_API_BEGIN e->impl.get_cv_result(out); B200_API_END
}
int rcppml_b200_get_cv_history(rcppml_b200_engine* e, float* train, float* test, int capacity) {
    B200_API_BEGIN
    Engine& E = e->impl;
    E.use_device();
    const int cnt = std::min<int>(capacity, static_cast<int>(E.test_hist.count));
    if (cnt > 0 && train) B200_CUDA_CHECK(cudaMemcpy(train, E.loss_hist.ptr, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    if (cnt > 0 && test) B200_CUDA_CHECK(cudaMemcpy(test, E.test_hist.ptr, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    B200_API_END
}
int rcppml_b200_get_result(rcppml_b200_engine* e, rcppml_b200_result* out) {
    B200_API_BEGIN e->impl.get_result(out); B200_API_END
}
int rcppml_b200_get_loss_history(rcppml_b200_engine* e, float* out, int capacity) {
    B200_API_BEGIN
    Engine& E = e->impl;
    E.use_device();
    const int cnt = std::min<int>(capacity, static_cast<int>(E.loss_hist.count));
    if (cnt > 0) B200_CUDA_CHECK(cudaMemcpy(out, E.loss_hist.ptr, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    B200_API_END
}
int rcppml_b200_set_profiling(rcppml_b200_engine* e, int enabled) {
    B200_API_BEGIN e->impl.profiling = enabled != 0; B200_API_END
}
int rcppml_b200_get_profile(rcppml_b200_engine* e, double* ms, int* launches) {
    B200_API_BEGIN
    for (int s = 0; s < RCPPML_B200_NUM_SECTIONS; ++s) {
        if (ms) ms[s] = e->impl.prof_ms[s];
        if (launches) launches[s] = e->impl.launches[s];
    }
    B200_API_END
}
int rcppml_b200_half_step(rcppml_b200_engine* e, const rcppml_b200_config* cfg, int which, int warm_start, int normalize_after) {
    B200_API_BEGIN e->impl.half_step_only(*cfg, which, warm_start != 0, normalize_after != 0); B200_API_END
}
// Bitwise check of the FMA-corrected division used by the solvers against __fdiv_rn.
int rcppml_b200_selftest_division(int64_t n, uint64_t seed, int64_t* mismatches) {
    B200_API_BEGIN
    unsigned long long* d_bad = nullptr;
    B200_CUDA_CHECK(cudaMalloc(&d_bad, sizeof(unsigned long long)));
    B200_CUDA_CHECK(cudaMemset(d_bad, 0, sizeof(unsigned long long)));
    b200::selftest_division_kernel<<<148 * 8, 256>>>(n, seed, d_bad);
    unsigned long long bad = 0;
    B200_CUDA_CHECK(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    *mismatches = static_cast<int64_t>(bad);
    B200_API_END
}
int64_t rcppml_b200_cd_sweeps(rcppml_b200_engine* e) { return static_cast<int64_t>(e->impl.cd_sweeps); }
int rcppml_b200_get_counters(rcppml_b200_engine* e, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    B200_API_BEGIN
    if (h2d_bytes) *h2d_bytes = static_cast<int64_t>(e->impl.h2d_bytes);
    if (d2h_bytes) *d2h_bytes = static_cast<int64_t>(e->impl.d2h_bytes);
    B200_API_END
}

}  // extern "C"
