// abi_reference.cu — the symbols the reference's bridge resolves out of RcppML_gpu.so
// (include/rcppml_gpu.h, part 1). Thin shims over the device-resident engine.
#include "engine.hpp"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// FACTORNET_GPU_WARN (core/logging.hpp:132) prints to stderr; the .so never calls into R.
void warn(const char* what) { std::fprintf(stderr, "[RcppML_gpu/b200] %s\n", what); }

// The reference creates its GPUContext per call (nmf/fit_gpu.cuh:552). Here the engine behind the reference entry
// points is cached per process (SURVEY.md §8b "may cache contexts in process-global state"): its device buffers
// and staging areas are grow-only, so repeated nmf() calls make no cudaMalloc / cudaFree at all. The calls are
// synchronous on the R main thread; the mutex only guards against misuse. RCPPML_B200_CACHE=0 restores one
// engine per call; rcppml_b200_release_cache() frees the device memory (also runs at dyn.unload / exit).
std::mutex g_mu;
std::unique_ptr<b200::Engine> g_engine;
double g_phases[5] = {0, 0, 0, 0, 0};
double g_call_wall_ms = 0.0;                     // wall clock of the last part-1 call, entry to return, measured in here

struct EngineLease {
    std::unique_lock<std::mutex> lock;
    std::unique_ptr<b200::Engine> own;
    b200::Engine* e = nullptr;
    explicit EngineLease(int dev) : lock(g_mu) {
        const char* env = std::getenv("RCPPML_B200_CACHE");
        if (env && env[0] == '0') {
            own.reset(new b200::Engine(dev));
            e = own.get();
            return;
        }
        if (!g_engine || g_engine->device != dev) {
            g_engine.reset();
            g_engine.reset(new b200::Engine(dev));
        }
        e = g_engine.get();
    }
    bool ok = false;
    void commit() { ok = true; }                     // the call succeeded: keep the cached engine
    ~EngineLease() {
        if (e) for (int i = 0; i < 5; ++i) g_phases[i] = e->phase_ms[i];
        if (!ok && !own) g_engine.reset();           // any failure: never reuse an engine in an unknown state
    }
    b200::Engine& get() { return *e; }
};

// ---- RCPPML_NUM_GPUS: multi-GPU behind the single-process reference caller (SURVEY.md §8b "Multi-GPU knob", §8e) ----
// The bridge signature has no device-count argument and R calls it from one thread of one process, so the knob is an
// environment variable read here (default 1; "all" or 0 = every usable device, capped at 8). With G > 1 the call runs
// the same sharded ALS as the one-process-per-GPU path (engine.cu: column block of H + row block of W_T per device,
// solved columns stored straight into every replica over NVLink, one-shot peer-memory all-reduces), with one Engine
// per device and one host thread per Engine inside this process; peers are plain device pointers after
// cudaDeviceEnablePeerAccess — no NCCL, no IPC.
// Ingest is sharded too: device g pulls only A[:, J_g] (nnz/G entries) and its blocks of W / H over ITS PCIe link; the
// row blocks A[I_g, :] are assembled from the peers' column blocks over NVLink (Engine::assemble_row_block), the factor
// replicas are completed by peer copies, and every device writes its own blocks of the result straight into the
// caller's W / H. Partition: contiguous ranges balanced by work (non-zeros + k per column).
// Results are bit-identical to one GPU (same per-column arithmetic; fp64 reductions only re-associate).
// The sm_100+ devices of this process, found once (cudaGetDeviceProperties costs milliseconds per call; the entry
// points are called once per nmf()).
int usable_devices(int* ids, int cap) {
    static std::once_flag once;
    static std::vector<int> found;
    std::call_once(once, [] {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return; }
        for (int dev = 0; dev < count; ++dev) {
            int major = 0;
            if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess && major >= 10) found.push_back(dev);
        }
    });
    int usable = 0;
    for (size_t i = 0; i < found.size() && usable < cap; ++i) ids[usable++] = found[i];
    return usable;
}

int requested_gpus() {
    const char* env = std::getenv("RCPPML_NUM_GPUS");
    if (!env || !env[0]) return 1;
    if (std::string(env) == "all") return 8;
    const int v = std::atoi(env);
    return v <= 0 ? 8 : std::min(v, 8);
}

// Host-side barrier of the per-device threads. fail() releases everybody: wait() then returns false on every thread,
// so one device's exception can never leave the others blocked.
struct PhaseBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n = 0, waiting = 0;
    unsigned gen = 0;
    bool failed = false;
    explicit PhaseBarrier(int n_) : n(n_) {}
    bool wait() {
        std::unique_lock<std::mutex> l(mu);
        if (failed) return false;
        const unsigned g = gen;
        if (++waiting == n) {
            waiting = 0;
            ++gen;
            cv.notify_all();
            return true;
        }
        cv.wait(l, [&] { return gen != g || failed; });
        return !failed;
    }
    void fail() {
        std::lock_guard<std::mutex> l(mu);
        failed = true;
        cv.notify_all();
    }
};

// Contiguous column ranges balanced by work = non-zeros + per_item per column, straight from the host col_ptr
// (rcppml_b200/shard.py balanced_cuts: cut r = smallest j with prefix(j) >= total * r / G).
void balanced_col_cuts(const int* col_ptr, int n, int G, int per_item, int* cuts) {
    auto prefix = [&](int j) { return static_cast<double>(col_ptr[j]) + static_cast<double>(per_item) * j; };
    const double total = prefix(n);
    cuts[0] = 0;
    for (int r = 1; r < G; ++r) {
        const double target = total * r / G;
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if (prefix(mid) < target) lo = mid + 1; else hi = mid;
        }
        cuts[r] = std::max(lo, cuts[r - 1]);
    }
    cuts[G] = n;
}

// The engines of the last multi-GPU call are kept (grow-only device buffers, like the single-GPU cache) and reused
// when the next call asks for the same devices; any failure drops them. Guarded by g_mu.
std::vector<std::unique_ptr<b200::Engine>> g_multi;
std::vector<int> g_multi_devices;
// The two multicast objects (W_T, H) of the cached in-process group: created by device 0's thread, shared by all
// engines, released here once every engine has unbound (engine destruction or mc_close).
CUmemGenericAllocationHandle g_mc_handles[2] = {0, 0};
bool g_mc_valid = false;
void release_group_multicast() {
    if (!g_mc_valid) return;
    const b200::DriverApi& drv = b200::DriverApi::get();
    if (drv.ok) { drv.MemRelease(g_mc_handles[0]); drv.MemRelease(g_mc_handles[1]); }
    g_mc_handles[0] = g_mc_handles[1] = 0;
    g_mc_valid = false;
}
void drop_group() {                                            // engines first (they unmap / unbind), then the objects
    g_multi.clear();
    g_multi_devices.clear();
    release_group_multicast();
}

// Returns false (after warn) on any failure; fills res / W / H / d on success. Caller holds g_mu.
// Optional: cv / cv_res — the speckled-mask cross-validation fit (nmf_fit_cv) instead of the plain one; mask_* — the
// explicit user mask (CSC pattern of the whole matrix; every device keeps its two slices).
bool fit_in_process_multi_gpu(int G, const int* devices, int m, int n, int64_t nnz, const int* col_ptr, const int* row_idx,
                              const double* values, int k, double* W, double* H, double* d, const rcppml_b200_config& cfg,
                              rcppml_b200_result* res, const rcppml_b200_cv_config* cv = nullptr,
                              rcppml_b200_cv_result* cv_res = nullptr, int64_t mask_nnz = 0, const int* mask_p = nullptr,
                              const int* mask_i = nullptr) {
    const auto t_call = std::chrono::steady_clock::now();
    const char* env = std::getenv("RCPPML_B200_CACHE");
    const bool cache = !(env && env[0] == '0');
    if (!(cache && static_cast<int>(g_multi.size()) == G && g_multi_devices == std::vector<int>(devices, devices + G))) {
        drop_group();
        g_multi.resize(G);
        g_multi_devices.assign(devices, devices + G);
    }
    std::vector<std::unique_ptr<b200::Engine>>& eng = g_multi;
    // RCPPML_B200_BALANCE=0: equal blocks (the partition of the one-process-per-GPU path's default)
    const char* benv = std::getenv("RCPPML_B200_BALANCE");
    const bool balance = !(benv && benv[0] == '0');
    std::vector<int> col_cuts(G + 1, 0), row_cuts(G + 1, 0);
    if (balance) balanced_col_cuts(col_ptr, n, G, k, col_cuts.data());
    std::vector<b200::Engine*> all(G, nullptr);
    std::vector<double> sumsq(G, 0.0);
    std::vector<rcppml_b200_result> rr(G);
    std::vector<std::string> err(G);
    PhaseBarrier bar(G);
    double marks[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};        // rank 0's wall clock at the phase boundaries (ms)
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };

    auto body = [&](int g) {
        if (!eng[g]) {
            eng[g].reset(new b200::Engine(devices[g]));
            eng[g]->comm_init_local(g, G);
        }
        b200::Engine& E = *eng[g];
        all[g] = &E;
        E.enable_peer_access(devices);
        E.set_partition(balance ? col_cuts.data() : nullptr, nullptr);
        E.upload_col_block_host<double>(m, n, col_ptr, row_idx, values);     // own column block + row histogram
        if (!bar.wait()) return;                                             // A: every column block is resident
        if (g == 0) {
            marks[0] = since();
            if (balance) E.balanced_row_cuts(all.data(), k, row_cuts.data());
            else row_cuts = E.row_cuts;                                      // equal blocks from set_dims
        }
        if (!bar.wait()) return;                                             // B: row cuts known
        E.install_row_cuts(row_cuts.data());
        E.start_factor_block_upload<double>(k, W, H);                        // copy engine, behind the row-block assembly
        sumsq[g] = E.assemble_row_block(all.data());                         // NVLink pulls + local transpose
        if (!bar.wait()) return;                                             // C: nobody reads a peer's column block any more
        if (g == 0) marks[1] = since();
        double s = 0.0;
        for (int r = 0; r < G; ++r) s += sumsq[r];                           // rank order: identical on every device
        E.finish_matrix_local(s, nnz);
        if (mask_nnz > 0) E.set_mask(mask_nnz, mask_p, mask_i);
        E.finish_factor_block_upload<double>(k);                             // own rows of W_T, own columns of H into place
        E.comm_prepare_local(devices);
        if (!bar.wait()) return;                                             // D: every replica buffer exists, own blocks in place
        E.comm_attach_local(all.data());
        // NVSwitch multicast (vmm.hpp): device 0's thread creates the two multicast objects, every device joins, binds
        // its replicas and maps them — from then on a normalised block is written once and lands in all replicas
        bool want_mc = true, have_mc = true;
        for (int r = 0; r < G; ++r) {
            want_mc = want_mc && all[r]->mc_wanted && all[r]->W_T.vmm && all[r]->H.vmm;
            have_mc = have_mc && all[r]->mc_ready;
        }
        if (!bar.wait()) return;                                             // F: every thread sampled the same state
        if (want_mc && !have_mc) {                                           // (a cached group keeps its mappings)
            E.mc_close();
            if (!bar.wait()) return;                                         // E0: nobody is bound to the old objects
            if (g == 0) {
                release_group_multicast();
                try { E.mc_create(&g_mc_handles[0], &g_mc_handles[1], false); g_mc_valid = true; }
                catch (const std::exception& ex) { warn((std::string("multicast unavailable, unicast peer stores instead: ") + ex.what()).c_str()); }
            }
            if (!bar.wait()) return;                                         // E1: the objects exist (or not)
            if (g_mc_valid) {
                E.mc_add_device(g_mc_handles[0], g_mc_handles[1]);
                if (!bar.wait()) return;                                     // E2: every device joined
                E.mc_bind_and_map(g_mc_handles[0], g_mc_handles[1], g == 0);
            }
        } else if (!want_mc && E.mc_ready) {
            E.mc_close();
        }
        E.pull_factor_blocks_from_peers();
        if (g == 0) marks[2] = since();
        if (cv) {
            E.fit_cv(cfg, *cv);                                              // begin_fit + iterations + d absorbed into H
            if (g == 0) { marks[5] = since(); E.get_cv_result(cv_res); }
        } else {
            E.begin_fit(cfg);                                                // (its first exchange orders the pulls against the peers' first stores)
            if (g == 0) marks[5] = since();
            E.iterate(cfg.max_iter);
        }
        E.get_result(&rr[g]);
        if (g == 0) marks[3] = since();
        if (rr[g].status == 0)
            E.get_factor_blocks_host<double>(W + static_cast<size_t>(E.row_begin) * k, H + static_cast<size_t>(E.col_begin) * k,
                                             g == 0 ? d : nullptr);
        if (g == 0) marks[4] = since();
    };
    {
        std::vector<std::thread> th;
        th.reserve(G);
        auto guarded = [&](int g) {
            try { body(g); }
            catch (const std::exception& ex) { err[g] = ex.what()[0] ? ex.what() : "error"; bar.fail(); }
            catch (...) { err[g] = "unknown error"; bar.fail(); }
        };
        try {
            for (int g = 1; g < G; ++g) th.emplace_back(guarded, g);
        } catch (...) {                                        // thread creation failed: release the ranks that did start
            err[G - 1] = "could not start a host thread";
            bar.fail();
        }
        if (err[G - 1].empty()) guarded(0);                    // the calling thread drives device 0
        for (auto& t : th) t.join();
    }
    bool ok = true;
    for (int g = 0; g < G && ok; ++g)
        if (!err[g].empty()) { warn(("multi-GPU rank " + std::to_string(g) + ": " + err[g]).c_str()); ok = false; }
    if (ok && bar.failed) { warn("multi-GPU: a rank stopped early"); ok = false; }
    if (ok)
        for (int g = 0; g < G; ++g)
            if (rr[g].status != 0) {
                warn(rr[g].status == 2 ? "multi-GPU exchange timed out (a peer stopped)" : "Gram matrix not positive definite (Cholesky pivot <= 0)");
                ok = false;
                break;
            }
    if (ok) {
        *res = rr[0];
        // phases of the call as rank 0 saw them: column-block upload | row-block assembly + transpose | factor blocks
        // up + replicas | ALS loop (CUDA events) | factor blocks down
        g_phases[0] = marks[0];
        g_phases[1] = marks[1] - marks[0];
        g_phases[2] = marks[2] - marks[1];
        g_phases[3] = rr[0].loop_ms;
        g_phases[4] = marks[4] - marks[3];
    }
    marks[6] = since();
    for (int g = 0; g < G; ++g) {                              // every device idle before anything is reused or freed
        if (!eng[g]) continue;
        cudaSetDevice(devices[g]);
        cudaDeviceSynchronize();
    }
    if (!ok || !cache) drop_group();                           // never reuse engines in an unknown state
    cudaSetDevice(devices[0]);
    if (std::getenv("RCPPML_B200_TRACE"))
        std::fprintf(stderr, "[RcppML_gpu/b200] multi-GPU call (G=%d, %s), rank-0 wall clock ms: col blocks up %.2f | row blocks %.2f | "
                             "factor blocks + replicas %.2f | begin_fit %.2f | iterate+result %.2f (loop events %.2f) | blocks down %.2f | "
                             "threads joined %.2f | end %.2f\n",
                     G, (eng.size() && eng[0] && eng[0]->mc_ready) ? "multicast replication" : "unicast peer stores", marks[0], marks[1] - marks[0], marks[2] - marks[1], marks[5] - marks[2], marks[3] - marks[5], rr[0].loop_ms,
                     marks[4] - marks[3], marks[6], since());
    return ok;
}

}  // namespace

extern "C" {

// src/gpu_bridge_cluster.cu:24-47
void rcppml_gpu_detect(int* num_gpus, double* total_mem_mb, double* free_mem_mb, int* max_gpus, int* out_status) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) count = 0;
    int usable = 0;
    const int cap = max_gpus ? *max_gpus : 0;
    for (int dev = 0; dev < count; ++dev) {
        cudaDeviceProp prop{};
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) continue;
        if (prop.major < 10) continue;                       // this library carries sm_100a code only
        size_t free_b = 0, total_b = 0;
        if (cudaSetDevice(dev) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) continue;
        if (usable < cap) {
            if (total_mem_mb) total_mem_mb[usable] = static_cast<double>(total_b) / (1024.0 * 1024.0);
            if (free_mem_mb) free_mem_mb[usable] = static_cast<double>(free_b) / (1024.0 * 1024.0);
        }
        ++usable;
    }
    if (count > 0) cudaSetDevice(0);                         // src/gpu_bridge_common.cuh:60
    if (num_gpus) *num_gpus = usable;
    if (out_status) *out_status = usable > 0 ? 0 : -1;
}

// src/gpu_bridge_nmf.cu:340-455 (51 pointers, gpu/bridge_nmf.hpp:78-99)
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
    int* loss_type, double* /*huber_delta*/,
    int* /*irls_max_iter*/, double* /*irls_tol*/,
    const int*, const int*, const double*, int*, int* graph_W_nnz, double*,
    const int*, const int*, const double*, int*, int* graph_H_nnz, double*,
    int* projective, int* symmetric, int* solver_mode,
    int* out_iter, int* out_converged,
    double* out_train_loss, double* out_test_loss,
    double* out_best_test, int* out_best_iter,
    int* out_status)
{
    if (!out_status) return;
    *out_status = -1;
    try {
        const char* refuse = nullptr;
        if (*loss_type != 0) refuse = "non-MSE loss is outside the B200 CV path";
        else if (*projective || *symmetric) refuse = "projective/symmetric NMF is outside the B200 CV path";
        else if ((graph_W_nnz && *graph_W_nnz > 0) || (graph_H_nnz && *graph_H_nnz > 0)) refuse = "graph regularisation is outside the B200 CV path";
        else if (*k < 1 || *k > b200::kMaxKP) refuse = "rank must be in [1, 128] on the B200 CV path";
        else if (*max_iter <= 0) refuse = "max_iter must be positive";
        if (refuse) { warn(refuse); return; }
        int dev0 = 0;
        if (usable_devices(&dev0, 1) < 1) { warn("no sm_100+ CUDA device"); return; }
        rcppml_b200_config cfg{};
        cfg.k = *k; cfg.max_iter = *max_iter; cfg.tol = static_cast<float>(*tol);
        cfg.L1_H = static_cast<float>(*L1_H); cfg.L1_W = static_cast<float>(*L1_W);
        cfg.L2_H = static_cast<float>(*L2_H); cfg.L2_W = static_cast<float>(*L2_W);
        cfg.nonneg_W = *nonneg_W != 0; cfg.nonneg_H = *nonneg_H != 0;
        cfg.cd_maxit = *cd_maxit; cfg.cd_tol = 1e-8f;
        cfg.norm_type = *norm_type; cfg.solver_mode = *solver_mode; cfg.patience = 5; cfg.verbose = *verbose;
        rcppml_b200_cv_config cv{};
        cv.holdout_fraction = static_cast<float>(*holdout_frac);         // src/gpu_bridge_nmf.cu:379
        cv.cv_seed = static_cast<uint32_t>(*cv_seed);
        cv.seed = static_cast<uint32_t>(*seed);
        cv.mask_zeros = *mask_zeros != 0;
        cv.cv_patience = 5;                                              // core/config.hpp:257, not on the wire
        rcppml_b200_result res{};
        rcppml_b200_cv_result cr{};
        int devices[8];
        int G = 1;                                                       // RCPPML_NUM_GPUS, as for the standard entry
        if (requested_gpus() > 1) G = std::min(requested_gpus(), usable_devices(devices, 8));
        while (G > 1 && (*n < 64 * G || *m < 64 * G)) --G;
        std::unique_ptr<EngineLease> lease_holder;
        if (G > 1) {
            std::lock_guard<std::mutex> guard(g_mu);
            if (!fit_in_process_multi_gpu(G, devices, *m, *n, static_cast<int64_t>(*nnz), col_ptr, row_idx, values, *k, W, H, d,
                                          cfg, &res, &cv, &cr))
                return;
        } else {
            lease_holder.reset(new EngineLease(dev0));
            b200::Engine& E = lease_holder->get();
            E.set_matrix_and_factors_host<double, double>(*m, *n, static_cast<int64_t>(*nnz), col_ptr, row_idx, values, *k, W, H);
            E.fit_cv(cfg, cv);
            E.get_result(&res);
            if (res.status != 0) { warn("Gram matrix not positive definite (Cholesky pivot <= 0)"); return; }
            E.get_cv_result(&cr);
            E.get_factors_host<double>(W, H, d);
        }
        if (out_iter) *out_iter = res.iterations;
        if (out_converged) *out_converged = res.converged;
        if (out_train_loss) *out_train_loss = cr.train_loss;
        if (out_test_loss) *out_test_loss = cr.test_loss;
        if (out_best_test) *out_best_test = cr.best_test_loss;
        if (out_best_iter) *out_best_iter = cr.best_iter;
        if (lease_holder) lease_holder->commit();
        *out_status = 0;
    } catch (const std::exception& ex) {
        warn(ex.what());
        *out_status = -1;
    } catch (...) {
        warn("unknown error");
        *out_status = -1;
    }
}

// src/gpu_bridge_nmf.cu:879-967 (39 pointers; called through R's .C from R/gpu_backend.R:225-265). The CSC arrays
// already live in DEVICE memory (sp_read_gpu): their addresses arrive encoded as doubles because R has no int64.
// The reference runs this entry in fp64 with its config defaults (no solver_mode on this wire: the struct default
// is Cholesky+clip, core/config.hpp:133); this engine converts the values to fp32 on the device and runs the same
// fp32 ALS loop as the standard entry — factors agree with the reference's fp64 run to fp32 accuracy.
void rcppml_gpu_nmf_zerocopy_double(
    double* d_col_ptr_addr, double* d_row_idx_addr, double* d_values_addr,
    int* m, int* n, double* nnz_d, int* k,
    double* W, double* H, double* d,
    int* max_iter, double* tol,
    double* L1_H, double* L1_W, double* L2_H, double* L2_W,
    double* L21_H, double* L21_W,
    double* ortho_H, double* ortho_W,
    double* ub_H, double* ub_W,
    int* cd_maxit, int* verbose, int* /*seed*/,
    int* /*loss_every*/, int* patience,
    int* nonneg_W, int* nonneg_H,
    int* loss_type, double* /*huber_delta*/,
    int* /*irls_max_iter*/, double* /*irls_tol*/,
    int* norm_type,
    int* out_iter, int* out_converged, double* out_loss,
    int* out_status,
    double* out_tol)
{
    if (!out_status) return;
    *out_status = -1;
    try {
        const char* refuse = nullptr;
        if (*loss_type != 0) refuse = "non-MSE loss is outside the B200 ALS path";
        else if (*L21_H > 0 || *L21_W > 0) refuse = "L21 is outside the B200 ALS path";
        else if (*ortho_H > 0 || *ortho_W > 0) refuse = "angular/ortho penalty is outside the B200 ALS path";
        else if (*k < 1 || *k > b200::kMaxKP) refuse = "rank must be in [1, 128] on the B200 ALS path";
        else if (*max_iter <= 0) refuse = "max_iter must be positive";
        else if (!(*nnz_d >= 0 && *nnz_d < 2147483648.0)) refuse = "nnz must fit int32";
        if (refuse) { warn(refuse); return; }
        auto to_ptr = [](double addr) { return reinterpret_cast<void*>(static_cast<uintptr_t>(addr)); };
        const int* dev_col_ptr = static_cast<const int*>(to_ptr(*d_col_ptr_addr));
        const int* dev_row_idx = static_cast<const int*>(to_ptr(*d_row_idx_addr));
        const double* dev_values = static_cast<const double*>(to_ptr(*d_values_addr));
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, dev_col_ptr) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
            cudaGetLastError();
            warn("zero-copy entry: col_ptr is not a device pointer");
            return;
        }
        EngineLease lease(at.device);                        // run where the matrix already is
        b200::Engine& E = lease.get();
        E.set_matrix_host<double>(*m, *n, static_cast<int64_t>(*nnz_d), dev_col_ptr, dev_row_idx, dev_values);
        E.h2d_bytes = 0;                                     // the matrix never crossed PCIe
        E.set_factors_host<double>(*k, W, H);
        rcppml_b200_config cfg{};
        cfg.k = *k; cfg.max_iter = *max_iter; cfg.tol = static_cast<float>(*tol);
        cfg.L1_H = static_cast<float>(*L1_H); cfg.L1_W = static_cast<float>(*L1_W);
        cfg.L2_H = static_cast<float>(*L2_H); cfg.L2_W = static_cast<float>(*L2_W);
        cfg.ub_H = static_cast<float>(*ub_H); cfg.ub_W = static_cast<float>(*ub_W);
        cfg.nonneg_W = *nonneg_W != 0; cfg.nonneg_H = *nonneg_H != 0;
        cfg.cd_maxit = *cd_maxit; cfg.cd_tol = 1e-8f;
        cfg.norm_type = *norm_type;
        cfg.solver_mode = 1;                                 // core/config.hpp:133 (not on this wire)
        cfg.patience = *patience; cfg.verbose = *verbose;
        E.begin_fit(cfg);
        E.iterate(cfg.max_iter);
        rcppml_b200_result res{};
        E.get_result(&res);
        if (res.status != 0) { warn("Gram matrix not positive definite (Cholesky pivot <= 0)"); return; }
        E.get_factors_host<double>(W, H, d);
        if (out_iter) *out_iter = res.iterations;
        if (out_converged) *out_converged = res.converged;
        if (out_loss) *out_loss = static_cast<double>(res.train_loss);
        if (out_tol) *out_tol = static_cast<double>(res.final_tol);
        lease.commit();
        *out_status = 0;
    } catch (const std::exception& ex) {
        warn(ex.what());
        *out_status = -1;
    } catch (...) {
        warn("unknown error");
        *out_status = -1;
    }
}

// Shared body of the standard entry point and its masked extension.
static void nmf_unified_impl(
    const int* mask_p, const int* mask_i, const int* mask_nnz,
    const int* col_ptr, const int* row_idx, const double* values,
    int* m, int* n, int* nnz, int* k,
    double* W, double* H, double* d,
    int* max_iter, double* tol,
    double* L1_H, double* L1_W, double* L2_H, double* L2_W,
    double* L21_H, double* L21_W,
    double* ortho_H, double* ortho_W,
    double* ub_H, double* ub_W,
    int* cd_maxit, int* verbose, int* /*seed*/,
    int* /*loss_every*/, int* patience,
    int* nonneg_W, int* nonneg_H,
    int* loss_type, double* /*huber_delta*/,
    int* /*irls_max_iter*/, double* /*irls_tol*/,
    int* norm_type,
    int* projective, int* symmetric,
    int* solver_mode,
    const int*, const int*, const double*, int*, int* graph_W_nnz, double*,
    const int*, const int*, const double*, int*, int* graph_H_nnz, double*,
    int* /*gp_dispersion_mode*/,
    double*, double*, double*, double*, double*, double*, double*, double*, double*,
    double* robust_delta, double* /*tweedie_power*/,
    double* /*out_theta*/, int* out_theta_len,
    const int*, const int*, const double*, const int*, int* guide_H_count,
    int* out_iter, int* out_converged, double* out_loss,
    int* out_status,
    double* out_tol);

// src/gpu_bridge_nmf.cu:460-624
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
    const int* gWp, const int* gWi, const double* gWx, int* gWd, int* graph_W_nnz, double* gWl,
    const int* gHp, const int* gHi, const double* gHx, int* gHd, int* graph_H_nnz, double* gHl,
    int* gp_dispersion_mode,
    double* t1, double* t2, double* t3, double* t4, double* t5, double* t6, double* t7, double* t8, double* t9,
    double* robust_delta, double* tweedie_power,
    double* out_theta, int* out_theta_len,
    const int* g1, const int* g2, const double* g3, const int* g4, int* guide_H_count,
    int* out_iter, int* out_converged, double* out_loss,
    int* out_status,
    double* out_tol)
{
    nmf_unified_impl(nullptr, nullptr, nullptr, col_ptr, row_idx, values, m, n, nnz, k, W, H, d, max_iter, tol, L1_H, L1_W,
                     L2_H, L2_W, L21_H, L21_W, ortho_H, ortho_W, ub_H, ub_W, cd_maxit, verbose, seed, loss_every, patience,
                     nonneg_W, nonneg_H, loss_type, huber_delta, irls_max_iter, irls_tol, norm_type, projective, symmetric,
                     solver_mode, gWp, gWi, gWx, gWd, graph_W_nnz, gWl, gHp, gHi, gHx, gHd, graph_H_nnz, gHl,
                     gp_dispersion_mode, t1, t2, t3, t4, t5, t6, t7, t8, t9, robust_delta, tweedie_power, out_theta,
                     out_theta_len, g1, g2, g3, g4, guide_H_count, out_iter, out_converged, out_loss, out_status, out_tol);
}

// ABI EXTENSION (not in the reference): the standard entry point plus the user mask, which the reference
// bridge does not carry (gpu/bridge_nmf.hpp:39-75). mask_p[n+1] / mask_i[mask_nnz] are the CSC
// pattern of the masked entries (nmf/masked_nnls.hpp). Same 73 arguments follow.
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
    const int* gWp, const int* gWi, const double* gWx, int* gWd, int* graph_W_nnz, double* gWl,
    const int* gHp, const int* gHi, const double* gHx, int* gHd, int* graph_H_nnz, double* gHl,
    int* gp_dispersion_mode,
    double* t1, double* t2, double* t3, double* t4, double* t5, double* t6, double* t7, double* t8, double* t9,
    double* robust_delta, double* tweedie_power,
    double* out_theta, int* out_theta_len,
    const int* g1, const int* g2, const double* g3, const int* g4, int* guide_H_count,
    int* out_iter, int* out_converged, double* out_loss,
    int* out_status,
    double* out_tol)
{
    nmf_unified_impl(mask_p, mask_i, mask_nnz, col_ptr, row_idx, values, m, n, nnz, k, W, H, d, max_iter, tol, L1_H, L1_W,
                     L2_H, L2_W, L21_H, L21_W, ortho_H, ortho_W, ub_H, ub_W, cd_maxit, verbose, seed, loss_every, patience,
                     nonneg_W, nonneg_H, loss_type, huber_delta, irls_max_iter, irls_tol, norm_type, projective, symmetric,
                     solver_mode, gWp, gWi, gWx, gWd, graph_W_nnz, gWl, gHp, gHi, gHx, gHd, graph_H_nnz, gHl,
                     gp_dispersion_mode, t1, t2, t3, t4, t5, t6, t7, t8, t9, robust_delta, tweedie_power, out_theta,
                     out_theta_len, g1, g2, g3, g4, guide_H_count, out_iter, out_converged, out_loss, out_status, out_tol);
}

static void nmf_unified_impl(
    const int* mask_p, const int* mask_i, const int* mask_nnz,
    const int* col_ptr, const int* row_idx, const double* values,
    int* m, int* n, int* nnz, int* k,
    double* W, double* H, double* d,
    int* max_iter, double* tol,
    double* L1_H, double* L1_W, double* L2_H, double* L2_W,
    double* L21_H, double* L21_W,
    double* ortho_H, double* ortho_W,
    double* ub_H, double* ub_W,
    int* cd_maxit, int* verbose, int* /*seed*/,
    int* /*loss_every*/, int* patience,
    int* nonneg_W, int* nonneg_H,
    int* loss_type, double* /*huber_delta*/,
    int* /*irls_max_iter*/, double* /*irls_tol*/,
    int* norm_type,
    int* projective, int* symmetric,
    int* solver_mode,
    const int*, const int*, const double*, int*, int* graph_W_nnz, double*,
    const int*, const int*, const double*, int*, int* graph_H_nnz, double*,
    int* /*gp_dispersion_mode*/,
    double*, double*, double*, double*, double*, double*, double*, double*, double*,
    double* robust_delta, double* /*tweedie_power*/,
    double* /*out_theta*/, int* out_theta_len,
    const int*, const int*, const double*, const int*, int* guide_H_count,
    int* out_iter, int* out_converged, double* out_loss,
    int* out_status,
    double* out_tol)
{
    if (!out_status) return;
    *out_status = -1;
    const auto t_entry = std::chrono::steady_clock::now();
    struct WallClock {
        std::chrono::steady_clock::time_point t0;
        ~WallClock() { g_call_wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
    } wall_clock{t_entry};
    try {
        // Features outside the sparse-MSE ALS path are refused, not emulated: the caller
        // (nmf/fit.hpp:125-133) owns the CPU fallback.
        const char* refuse = nullptr;
        if (*loss_type != 0) refuse = "non-MSE loss is outside the B200 ALS path";
        else if (robust_delta && *robust_delta > 0) refuse = "robust IRLS is outside the B200 ALS path";
        else if (*projective || *symmetric) refuse = "projective/symmetric NMF is outside the B200 ALS path";
        else if (*L21_H > 0 || *L21_W > 0) refuse = "L21 is outside the B200 ALS path";
        else if (*ortho_H > 0 || *ortho_W > 0) refuse = "angular/ortho penalty is outside the B200 ALS path";
        else if ((graph_W_nnz && *graph_W_nnz > 0) || (graph_H_nnz && *graph_H_nnz > 0)) refuse = "graph regularisation is outside the B200 ALS path";
        else if (guide_H_count && *guide_H_count > 0) refuse = "classifier guides are outside the B200 ALS path";
        else if (*k < 1 || *k > b200::kMaxKP) refuse = "rank must be in [1, 128] on the B200 ALS path";
        else if (*max_iter <= 0) refuse = "max_iter must be positive";
        if (refuse) { warn(refuse); return; }

        int dev0 = 0;
        if (usable_devices(&dev0, 1) < 1) { warn("no sm_100+ CUDA device"); return; }

        rcppml_b200_config cfg{};
        cfg.k = *k;
        cfg.max_iter = *max_iter;
        cfg.tol = static_cast<float>(*tol);
        cfg.L1_H = static_cast<float>(*L1_H); cfg.L1_W = static_cast<float>(*L1_W);
        cfg.L2_H = static_cast<float>(*L2_H); cfg.L2_W = static_cast<float>(*L2_W);
        cfg.ub_H = static_cast<float>(*ub_H); cfg.ub_W = static_cast<float>(*ub_W);
        cfg.nonneg_W = *nonneg_W != 0; cfg.nonneg_H = *nonneg_H != 0;
        cfg.cd_maxit = *cd_maxit;
        cfg.cd_tol = 1e-8f;                       // not on the wire: core/constants.hpp:64
        cfg.norm_type = *norm_type;
        cfg.solver_mode = *solver_mode;
        cfg.patience = *patience;
        cfg.verbose = *verbose;

        const bool masked = mask_p && mask_i && mask_nnz && *mask_nnz > 0;
        int devices[8];
        int G = 1;
        if (requested_gpus() > 1) G = std::min(requested_gpus(), usable_devices(devices, 8));
        while (G > 1 && (*n < 64 * G || *m < 64 * G)) --G;             // every device needs a real block of each factor
        rcppml_b200_result res{};
        std::unique_ptr<EngineLease> lease_holder;
        if (G > 1) {
            std::lock_guard<std::mutex> guard(g_mu);
            if (!fit_in_process_multi_gpu(G, devices, *m, *n, static_cast<int64_t>(*nnz), col_ptr, row_idx, values, *k, W, H, d,
                                          cfg, &res, nullptr, nullptr, masked ? *mask_nnz : 0, mask_p, mask_i))
                return;
        } else {
            lease_holder.reset(new EngineLease(dev0));
            b200::Engine& E = lease_holder->get();
            if (masked) {
                E.set_matrix_host<double>(*m, *n, static_cast<int64_t>(*nnz), col_ptr, row_idx, values);
                E.set_mask(*mask_nnz, mask_p, mask_i);
                E.set_factors_host<double>(*k, W, H);
            } else {
                E.set_matrix_and_factors_host<double, double>(*m, *n, static_cast<int64_t>(*nnz), col_ptr, row_idx, values, *k, W, H);
            }
            E.begin_fit(cfg);
            E.iterate(cfg.max_iter);
            E.get_result(&res);
            if (res.status != 0) { warn("Gram matrix not positive definite (Cholesky pivot <= 0)"); return; }
            E.get_factors_host<double>(W, H, d);
        }

        if (out_theta_len) *out_theta_len = 0;
        if (out_iter) *out_iter = res.iterations;
        if (out_converged) *out_converged = res.converged;
        if (out_loss) *out_loss = static_cast<double>(res.train_loss);
        if (out_tol) *out_tol = static_cast<double>(res.final_tol);
        if (*verbose)
            std::fprintf(stderr, "[RcppML_gpu/b200] %d iterations, loss %.6g, loop %.3f ms, %d launches\n",
                         res.iterations, res.train_loss, res.loop_ms, res.gpu_launches);
        if (lease_holder) lease_holder->commit();
        *out_status = 0;
    } catch (const std::exception& ex) {
        warn(ex.what());
        *out_status = -1;
    } catch (...) {
        warn("unknown error");
        *out_status = -1;
    }
}

// ABI EXTENSIONS around the cached engine (see EngineLease above).
int rcppml_b200_release_cache(void) {
    std::lock_guard<std::mutex> g(g_mu);
    g_engine.reset();
    drop_group();
    return 0;
}
// Host wall-clock (ms) of the phases of the last reference-ABI call: matrix upload (+fp64->fp32), device
// transpose + tr(AtA), factor upload, ALS loop (CUDA events), factor download.
int rcppml_b200_last_call_phases(double* ms5) {
    std::lock_guard<std::mutex> g(g_mu);
    for (int i = 0; i < 5; ++i) ms5[i] = g_phases[i];
    return 0;
}
double rcppml_b200_last_call_wall_ms(void) { return g_call_wall_ms; }
// The column partition the in-process multi-GPU path uses (host only, no device work): world + 1 ascending cuts of the
// n columns of a CSC matrix, balanced by work = non-zeros + per_item per column.
int rcppml_b200_balanced_col_cuts(const int* col_ptr, int n, int world, int per_item, int* cuts) {
    if (!col_ptr || !cuts || n < 0 || world < 1) return -1;
    balanced_col_cuts(col_ptr, n, world, per_item, cuts);
    return 0;
}

}  // extern "C"
