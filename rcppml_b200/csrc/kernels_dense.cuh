// kernels_dense.cuh — the dense k×k / k×n pieces of one ALS iteration:
//   * normalize_gram_kernel : X[i,:] /= d_i fused with the Gram G = X·Xᵀ of the normalised factor
//     (nmf/variant_helpers.hpp:302-304 + primitives/cpu/gram.hpp:58-67), one pass over X
//   * gram_reduce_kernel    : fixed-order fp64 reduction of the per-CTA Gram partials
//   * prepare_solver_kernel : G += L2·I (fit_cpu.hpp:506,738) and, for solver_mode != 0, the LLT
//     factorisation (fused_nnls.hpp:185) in the oracle's operation order
//   * scale_finalize_kernel : d_i = Σ partials (+sqrt) + 1e-15 (variant_helpers.hpp:297-301)
//   * loss_kernel           : Gram-trick loss + convergence/patience (fit_cpu.hpp:1729-1809)
#pragma once

#include "common.cuh"
#include <type_traits>

namespace b200 {

// Device-resident iteration state: lets the host enqueue iterations without synchronising.
struct DevState {
    int stop;               // set when converged (patience reached) -> later kernels exit at once
    int iter;               // iterations completed (result.iterations)
    int converged;
    int patience_counter;
    int chol_fail;          // first non-positive pivot index + 1 (0 = none)
    int comm_error;         // a peer-memory exchange timed out
    float prev_loss;
    float final_tol;
    float train_loss;
};

// One 128-bit store into EVERY replica bound to the multicast object behind `mc` (NVSwitch multicast, NVLS): the
// switch replicates the packet, the sender's NVLink egress carries it once.
__device__ __forceinline__ void multimem_store4(float4* mc, const float4& v) {
    // (no "memory" clobber: it would be a compiler barrier in every kernel that merely CONTAINS the store — measured:
    // the Gram kernel lost its load/MMA overlap, 0.32 -> 0.52 ms, with the multicast path not even taken)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

__device__ __forceinline__ void multimem_store1(float* mc, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v));
}

// ---------------------------------------------------------------------------------------------
// normalize + Gram. X is [ncols][KP] (row c = column c of the reference's k×ncols matrix).
// G is symmetric: of the 16×16 grid of R×R tiles (R = KP/16) only the lower triangle is needed. A warp owns
// a 4×8 block of tiles (lane -> tile row 4·bi + lane/8, tile column 8·bj + lane%8) so that a shared-memory
// read of the warp touches only 4 (rows) or 8 (columns) distinct 32-byte chunks — 1 wavefront per LDS.128;
// 6 of the 8 blocks intersect the lower triangle -> 6 warps. Column tiles of TC columns are staged in shared
// memory ALREADY CONVERTED to fp64 (rows padded by 16 B per 128 B against bank conflicts); the next tile is
// prefetched into registers while the current one is multiplied; the inner loop is LDS.128 + DFMA only.
// ---------------------------------------------------------------------------------------------
constexpr int kGramThreads = 192;

template <int KP, int TC>
static __global__ void __launch_bounds__(kGramThreads) normalize_gram_kernel(float* __restrict__ X, long long ncols,
                                                             const float* __restrict__ d, int normalize,
                                                             double* __restrict__ partials,
                                                             const int* __restrict__ stop_flag, float* __restrict__ mcX) {
    constexpr int R = KP / 16;
    constexpr int V4 = KP / 4;                       // float4 per column
    constexpr int ROWD = KP + (KP / 16) * 2;         // doubles per staged row incl. padding (2 doubles per 16)
    __shared__ __align__(16) double sX[TC][ROWD];
    __shared__ float sD[KP];
    if (*stop_flag) return;
    for (int t = threadIdx.x; t < KP; t += blockDim.x) sD[t] = normalize ? d[t] : 1.f;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bi = (warp < 4) ? warp : warp - 2;     // blocks (0,0) (1,0) (2,0) (3,0) (2,1) (3,1)
    const int bj = (warp < 4) ? 0 : 1;
    const int ti = 4 * bi + (lane >> 3), tj = 8 * bj + (lane & 7);
    auto pos = [](int i) { return i + (i / 16) * 2; };          // padded position of coordinate i in a staged row
    const int pa = pos(ti * R), pb = pos(tj * R);               // R <= 8 consecutive coordinates never straddle a pad
    // fp64 accumulation of exact fp32×fp32 products: the result, rounded once to fp32 by
    // gram_from_sums_kernel, is the order-independent "correctly rounded" Gram the oracle defines
    // (oracle/nmf_oracle.cpp gram()). A G that matches bit for bit keeps every column solve
    // bit-identical to the CPU path; B200 runs DFMA at half the FFMA rate and the Gram is a few
    // percent of an iteration.
    double acc[R][R];
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int b = 0; b < R; ++b) acc[a][b] = 0.0;

    constexpr int PER = (TC * V4 + kGramThreads - 1) / kGramThreads;     // float4 per thread per tile
    const long long ntiles = (ncols + TC - 1) / TC;
    float4 pre[PER];
    auto fetch = [&](long long tile) {
        const long long c0 = tile * TC;
        const long long left = ncols - c0;
        const int nc = left < TC ? static_cast<int>(left) : TC;
        float4* X4 = reinterpret_cast<float4*>(X + c0 * KP);
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int t = threadIdx.x + u * kGramThreads;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < TC * V4 && t / V4 < nc) {
                const int q = t % V4;
                v = X4[t];
                if (normalize) {                     // padded coordinates hold 0 and d = 1 there
                    v.x = __fdiv_rn(v.x, sD[q * 4 + 0]);
                    v.y = __fdiv_rn(v.y, sD[q * 4 + 1]);
                    v.z = __fdiv_rn(v.z, sD[q * 4 + 2]);
                    v.w = __fdiv_rn(v.w, sD[q * 4 + 3]);
                    if (!mcX) X4[t] = v;
                }
                // sharded fits with multicast-bound factors: the (normalised) block goes into ALL replicas, this rank's
                // included, with one store per word
                if (mcX) multimem_store4(reinterpret_cast<float4*>(mcX + c0 * KP) + t, v);
            }
            pre[u] = v;
        }
    };
    long long tile = blockIdx.x;
    if (tile < ntiles) fetch(tile);
    for (; tile < ntiles; tile += gridDim.x) {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int t = threadIdx.x + u * kGramThreads;
            if (t < TC * V4) {
                const int c = t / V4, q = t % V4;
                double2* dst = reinterpret_cast<double2*>(&sX[c][pos(q * 4)]);
                dst[0] = make_double2(static_cast<double>(pre[u].x), static_cast<double>(pre[u].y));
                dst[1] = make_double2(static_cast<double>(pre[u].z), static_cast<double>(pre[u].w));
            }
        }
        __syncthreads();
        if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);      // overlaps with the DFMA loop below
#pragma unroll 4
        for (int c = 0; c < TC; ++c) {
            double av[R], bv[R];
#pragma unroll
            for (int a = 0; a < R; ++a) av[a] = sX[c][pa + a];
#pragma unroll
            for (int b = 0; b < R; ++b) bv[b] = sX[c][pb + b];
#pragma unroll
            for (int a = 0; a < R; ++a)
#pragma unroll
                for (int b = 0; b < R; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
    if (ti >= tj) {                                  // tiles above the diagonal are redundant
        double* out = partials + static_cast<size_t>(blockIdx.x) * KP * KP;
#pragma unroll
        for (int a = 0; a < R; ++a)
#pragma unroll
            for (int b = 0; b < R; ++b)
                out[(tj * R + b) * KP + (ti * R + a)] = acc[a][b];   // G(i,j), i >= j tile-wise, at [j*KP+i]
    }
}

// ---------------------------------------------------------------------------------------------
// normalize + Gram on the FP64 tensor cores (DMMA m8n8k4). Same contract as normalize_gram_kernel — exact
// fp32xfp32 products accumulated in fp64, rounded once by gram_from_sums_kernel — but the operands reach the
// math pipe as MMA fragments: one LDS.64 per lane feeds 256 FMAs, where the register-tiled DFMA version needs
// 8 loads per 16 FMAs and is bound by shared-memory delivery (ncu r01c/r01g: l1tex 74 %, FP64 pipe 38 %).
// G is cut into 8x8 tiles; only the NT(NT+1)/2 tiles of the lower triangle are computed, TPW per warp.
// Fragment of tile-row t for the 4 staged columns c..c+3: lane l holds X[t*8 + l/4][c + l%4] — this is both
// the A fragment (rows of G) and the col-major B fragment (columns of G). Staged rows are padded to
// KP + 4 doubles so that the 16 lanes of a half-warp hit 16 distinct 8-byte banks.
// ---------------------------------------------------------------------------------------------
template <int KP> struct GramMmaGeom;
template <> struct GramMmaGeom<16> { static constexpr int WARPS = 3, TPW = 1; };
template <> struct GramMmaGeom<32> { static constexpr int WARPS = 5, TPW = 2; };
template <> struct GramMmaGeom<64> { static constexpr int WARPS = 6, TPW = 6; };
template <> struct GramMmaGeom<128> { static constexpr int WARPS = 8, TPW = 17; };

// REPL: also write the (normalised) block into every replica through the multicast alias mcX (Engine::mc_mode 1) — a
// separate instantiation, so that the common kernel keeps its 80 registers (4 CTAs/SM; with the extra pointer
// arithmetic it needed 94 and lost a resident CTA: 0.32 -> 0.52 ms).
template <int KP, int TC, bool REPL = false>
static __global__ void __launch_bounds__(GramMmaGeom<KP>::WARPS * 32) normalize_gram_mma_kernel(
    float* __restrict__ X, long long ncols, const float* __restrict__ d, int normalize, double* __restrict__ partials,
    const int* __restrict__ stop_flag, float* __restrict__ mcX) {
    constexpr int WARPS = GramMmaGeom<KP>::WARPS, TPW = GramMmaGeom<KP>::TPW, THREADS = WARPS * 32;
    constexpr int NT = KP / 8;
    static_assert(WARPS * TPW == NT * (NT + 1) / 2, "tiles of the lower triangle must split evenly over the warps");
    constexpr int V4 = KP / 4;
    constexpr int ROWD = KP + 4;
    __shared__ __align__(16) double sX[TC][ROWD];
    __shared__ float sD[KP];
    if (*stop_flag) return;
    for (int t = threadIdx.x; t < KP; t += blockDim.x) sD[t] = normalize ? d[t] : 1.f;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile u of this warp: index e = warp*TPW + u in the row-major enumeration of the lower triangle
    int fa[TPW], fb[TPW];                              // fragment offsets (doubles) inside a 4-column slab
    int t_i[TPW], t_j[TPW];
#pragma unroll
    for (int u = 0; u < TPW; ++u) {
        const int e = warp * TPW + u;
        int ti = 0;
        while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;
        const int tj = e - ti * (ti + 1) / 2;
        t_i[u] = ti; t_j[u] = tj;
        fa[u] = (lane & 3) * ROWD + ti * 8 + (lane >> 2);
        fb[u] = (lane & 3) * ROWD + tj * 8 + (lane >> 2);
    }
    double acc[TPW][2];
#pragma unroll
    for (int u = 0; u < TPW; ++u) acc[u][0] = acc[u][1] = 0.0;

    constexpr int PER = (TC * V4 + THREADS - 1) / THREADS;
    const long long ntiles = (ncols + TC - 1) / TC;
    float4 pre[PER];
    auto fetch = [&](long long tile) {
        const long long c0 = tile * TC;
        const long long left = ncols - c0;
        const int nc = left < TC ? static_cast<int>(left) : TC;
        float4* X4 = reinterpret_cast<float4*>(X + c0 * KP);
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int t = threadIdx.x + u * THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < TC * V4 && t / V4 < nc) {
                const int q = t % V4;
                v = X4[t];
                if (normalize) {                     // padded coordinates hold 0 and d = 1 there
                    v.x = __fdiv_rn(v.x, sD[q * 4 + 0]);
                    v.y = __fdiv_rn(v.y, sD[q * 4 + 1]);
                    v.z = __fdiv_rn(v.z, sD[q * 4 + 2]);
                    v.w = __fdiv_rn(v.w, sD[q * 4 + 3]);
                    if (!REPL) X4[t] = v;
                }
                // sharded fits with multicast-bound factors (REPL): the (normalised) block goes into ALL replicas, this
                // rank's included, with one store per word
                if (REPL) multimem_store4(reinterpret_cast<float4*>(mcX + c0 * KP) + t, v);
            }
            pre[u] = v;
        }
    };
    long long tile = blockIdx.x;
    if (tile < ntiles) fetch(tile);
    for (; tile < ntiles; tile += gridDim.x) {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int t = threadIdx.x + u * THREADS;
            if (t < TC * V4) {
                const int c = t / V4, q = t % V4;
                double2* dst = reinterpret_cast<double2*>(&sX[c][q * 4]);
                dst[0] = make_double2(static_cast<double>(pre[u].x), static_cast<double>(pre[u].y));
                dst[1] = make_double2(static_cast<double>(pre[u].z), static_cast<double>(pre[u].w));
            }
        }
        __syncthreads();
        if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);      // overlaps with the MMA loop below
        const double* slab = &sX[0][0];
#pragma unroll 2
        for (int c = 0; c < TC; c += 4, slab += 4 * ROWD) {
#pragma unroll
            for (int u = 0; u < TPW; ++u) {
                const double a = slab[fa[u]];
                const double b = slab[fb[u]];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(acc[u][0]), "+d"(acc[u][1]) : "d"(a), "d"(b));
            }
        }
        __syncthreads();
    }
    // C fragment: lane l holds C[l/4][2*(l%4) + {0,1}] of tile (ti, tj) -> G(i, j) stored at [j*KP + i]
    double* out = partials + static_cast<size_t>(blockIdx.x) * KP * KP;
#pragma unroll
    for (int u = 0; u < TPW; ++u) {
        const int i = t_i[u] * 8 + (lane >> 2);
        const int j = t_j[u] * 8 + 2 * (lane & 3);
        out[j * KP + i] = acc[u][0];
        out[(j + 1) * KP + i] = acc[u][1];
    }
}

// X[c][:] /= d for the columns c in [0, lo) and [hi, ncols): peer-memory sharded runs normalise the blocks the
// OTHER ranks solved (pushed un-normalised into this rank's replica by their half_step_kernel) with the same
// IEEE division normalize_gram_kernel applies to the own block. Pure HBM streaming.
static __global__ void __launch_bounds__(256) scale_columns_kernel(float* __restrict__ X, long long ncols, int KP,
                                                                   const float* __restrict__ d, long long lo,
                                                                   long long hi, const int* __restrict__ stop_flag) {
    __shared__ float sD[kMaxKP];
    if (*stop_flag) return;
    for (int t = threadIdx.x; t < KP; t += blockDim.x) sD[t] = d[t];
    __syncthreads();
    const int V4 = KP / 4;
    const long long head = lo * V4, skip = (hi - lo) * V4, total = (ncols - (hi - lo)) * V4;
    float4* X4 = reinterpret_cast<float4*>(X);
    // Four independent 128-bit loads in flight per thread. The kernel runs BESIDE the Gram chain of the main stream, on
    // the low-priority side stream (grid: Engine::side_ctas_per_sm CTAs per SM).
    constexpr int UN = 4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long t0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t0 < total; t0 += stride * UN) {
        float4 v[UN];
        long long w[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long t = t0 + u * stride;
            w[u] = t < head ? t : t + skip;
            if (t < total) v[u] = X4[w[u]];
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (t0 + u * stride >= total) break;
            const int q = static_cast<int>(w[u] % V4);
            v[u].x = __fdiv_rn(v[u].x, sD[q * 4 + 0]);
            v[u].y = __fdiv_rn(v[u].y, sD[q * 4 + 1]);
            v[u].z = __fdiv_rn(v[u].z, sD[q * 4 + 2]);
            v[u].w = __fdiv_rn(v[u].w, sD[q * 4 + 3]);
            X4[w[u]] = v[u];
        }
    }
}

// out[e] = Σ_c partials[c][e] in a FIXED order: 8 interleaved slices per element (slice s takes
// c ≡ s mod 8, ascending), combined in slice order. blockDim = (32, 8).
static __global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ partials, int nparts,
                                                                  int nelem, double* __restrict__ out,
                                                                  const int* __restrict__ stop_flag) {
    __shared__ double s[8][33];
    if (*stop_flag) return;
    const int e = blockIdx.x * 32 + threadIdx.x;
    double a = 0.0;
    if (e < nelem)
        for (int c = threadIdx.y; c < nparts; c += 8) a += partials[static_cast<size_t>(c) * nelem + e];
    s[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.y == 0 && e < nelem) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += s[q][threadIdx.x];
        out[e] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// One-shot all-reduce over NVLink peer memory, fused with nothing but itself: every rank stores its
// vector into slot[rank] of EVERY rank's exchange buffer (P2P stores), publishes a sequence number, waits
// for the world's sequence numbers, and sums the slots in rank order — the same order on every rank, so
// the result is bit-identical everywhere. Replaces a ~40 us ncclAllReduce of a few KB with a ~10 us kernel.
// Because a rank only signals after its previous stream work completed, the exchange is also the barrier
// that orders the P2P factor replication of half_step_kernel against the peers' next reads.
// xbuf layout (per rank): double data[2][world][ne_max]; unsigned long long flags[2][8].
// ---------------------------------------------------------------------------------------------
struct XchgParams {
    double* peer[8];                 // exchange buffer base of every rank (peer[rank] = own)
    int rank, world;
    int ne_max;
};
// The sequence number of an exchange lives in the rank's own exchange buffer (u64 slot 16 behind the flags) and is
// advanced by the kernel itself: every rank runs the same sequence of exchanges, so the counters stay in lock-step,
// and a launch carries no per-call argument — the sharded iteration can be captured once as a CUDA graph and replayed.
constexpr int kXchgTailWords = 24;   // flags[2][8] + sequence counter (+ padding), in 8-byte words

__device__ __forceinline__ unsigned long long* xchg_flags(double* base, int world, int ne_max, int phase) {
    return reinterpret_cast<unsigned long long*>(base + static_cast<size_t>(2) * world * ne_max) + phase * 8;
}
__device__ __forceinline__ unsigned long long* xchg_seq_counter(double* base, int world, int ne_max) {
    return reinterpret_cast<unsigned long long*>(base + static_cast<size_t>(2) * world * ne_max) + 16;
}

static __global__ void __launch_bounds__(1024) xchg_allreduce_kernel(const double* __restrict__ local, int nelem,
                                                                     double* __restrict__ out, const XchgParams x,
                                                                     DevState* __restrict__ st) {
    if (st->stop) return;            // identical on every rank (convergence is decided from all-reduced data)
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) {
        unsigned long long* ctr = xchg_seq_counter(x.peer[x.rank], x.world, x.ne_max);
        s_seq = *ctr + 1ULL;
        *ctr = s_seq;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int phase = static_cast<int>(seq & 1ULL);
    const size_t slot = (static_cast<size_t>(phase) * x.world + x.rank) * x.ne_max;
    for (int r = 0; r < x.world; ++r) {
        double* dst = x.peer[r] + slot;
        for (int e = threadIdx.x; e < nelem; e += blockDim.x) dst[e] = local[e];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < x.world) {
        unsigned long long* f = xchg_flags(x.peer[threadIdx.x], x.world, x.ne_max, phase) + x.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
        const unsigned long long* mine = xchg_flags(x.peer[x.rank], x.world, x.ne_max, phase) + threadIdx.x;
        unsigned long long v = 0;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if (v >= seq) break;
            // ~2 s: never hang the GPU. `stop` turns every later kernel of this fit into a no-op, so a dead peer costs
            // one time-out, not one per exchange; the host sees comm_error at its next poll (Engine::iterate).
            if (clock64() - t0 > 4000000000LL) { st->comm_error = 1; st->stop = 1; break; }
        }
    }
    __syncthreads();
    const double* base = x.peer[x.rank] + static_cast<size_t>(phase) * x.world * x.ne_max;
    for (int e = threadIdx.x; e < nelem; e += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < x.world; ++r) s += __ldcg(base + static_cast<size_t>(r) * x.ne_max + e);
        out[e] = s;
    }
}

// G = float(sums) symmetrised from the lower triangle (gram.hpp:64-65) with tiny_num on the
// diagonal (gram.hpp:66). Padded rows/cols stay 0. `sums` may have been all-reduced over ranks.
static __global__ void gram_from_sums_kernel(const double* __restrict__ sums, int KP, int k, float* __restrict__ G,
                                             const int* __restrict__ stop_flag) {
    if (*stop_flag) return;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= KP * KP) return;
    const int i = e % KP, j = e / KP;
    const int lo = (i >= j) ? (j * KP + i) : (i * KP + j);     // lower-triangle source element
    float v = static_cast<float>(sums[lo]);
    if (i == j) v = __fadd_rn(v, 1e-15f);
    if (i >= k || j >= k) v = 0.f;
    G[e] = v;
}

// One CTA. Builds the k×k operands of the solve kernel from G (+ L2·I, fit_cpu.hpp:506,738):
//   CD  : M1 = G (+L2·I), rcp = RN(1/G_ii) (0 if G_ii <= 0)
//   CHOL: unblocked left-looking LLT in the oracle's order (dot accumulated sequentially, then
//         subtracted, IEEE sqrt/div — fused_nnls.hpp:185 restated); M1 = strictly-lower L (col-major),
//         M2[p*KP+i] = L(p,i) (i<p), both with the diagonal blocks zeroed; dblk = diagonal blocks of L
//         (lower triangle incl. diagonal, [q][row][col]); rcp = RN(1/L_pp).
// Padded pivots get dblk diagonal 1 (CHOL) / 0 (CD) so that they solve to 0 / are skipped.
constexpr int kPrepThreads = 1024;
template <int KP>
static __global__ void __launch_bounds__(kPrepThreads) prepare_solver_kernel(const float* __restrict__ G, int k, float L2,
                                                            int solver, float* __restrict__ M1,
                                                            float* __restrict__ M2, float* __restrict__ dblk,
                                                            float* __restrict__ rcp, DevState* __restrict__ st) {
    extern __shared__ float sL[];          // KP*KP col-major: G (+L2·I), overwritten column by column with L (below the diagonal)
    __shared__ float sDiag[KP];            // L(j,j) (the diagonal of the working copy keeps G(j,j): every row reads it)
    if (st->stop) return;
    const int tid = threadIdx.x;
    for (int e = tid; e < KP * KP; e += kPrepThreads) {
        float v = G[e];
        const int i = e % KP, j = e / KP;
        if (i == j && i < k && L2 > 0.f) v = __fadd_rn(v, L2);
        sL[e] = v;
    }
    if (tid < KP) sDiag[tid] = 0.f;
    __syncthreads();
    if (solver != 0) {
        // LLT in the oracle's operation order: L(i,j) = (G(i,j) - t_ij) / L(j,j) with the dot product
        // t_ij = sum_{p<j} L(i,p)·L(j,p) accumulated sequentially in p (separately rounded mul and add), and
        // L(j,j) = sqrt(G(j,j) - t_jj). The dots are RUNNING sums that every finished column p updates for all pairs
        // (i >= c > p) at once — the same additions in the same order as the left-looking loop, k steps of O(1) depth.
        // Thread (ti, tq) OWNS the running dots T(ti, c) of the columns c = tq, tq + TJS, ... in registers, and — so that
        // no pivot has to be broadcast — its own copy of the diagonal dots T(c, c) of those columns (the multiplier
        // L(c,p) it reads for T(ti,c) += L(ti,p)·L(c,p) is all that needs: T(c,c) += L(c,p)·L(c,p), the very
        // operation the owner of row c performs). Column j is then finished by the threads with tq == j mod TJS from
        // registers alone, and ONE CTA barrier per column publishes it (round 1: three barriers and the dots in shared
        // memory, ~45 us at k = 64). IEEE division through div_exact with RN(1/L(j,j)) — bit-identical to __fdiv_rn
        // (rcppml_b200_selftest_division).
        constexpr int TJS = kPrepThreads / KP;             // column stride between a thread's columns
        constexpr int NU = (KP + TJS - 1) / TJS;           // columns per thread
        const int ti = tid % KP, tq = tid / KP;
        float T[NU], Td[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) T[u] = Td[u] = 0.f;
        for (int j = 0; j < k; ++j) {
            if (tq == j % TJS) {
                const int uj = j / TJS;
                float t = 0.f, td = 0.f;
#pragma unroll
                for (int u = 0; u < NU; ++u)
                    if (u == uj) { t = T[u]; td = Td[u]; }
                const float x = __fsub_rn(sL[j * KP + j], td);
                float ljj = 0.f;
                if (!(x > 0.f)) {
                    if (ti == 0 && st->chol_fail == 0) st->chol_fail = j + 1;
                } else {
                    ljj = __fsqrt_rn(x);
                }
                if (ti > j && ti < k) sL[j * KP + ti] = div_exact(__fsub_rn(sL[j * KP + ti], t), ljj, __frcp_rn(ljj));
                if (ti == j) sDiag[j] = ljj;
            }
            __syncthreads();                                // L(:, j) published
            const float li = sL[j * KP + ti];               // (rows <= j read G or garbage: their dots are never used)
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int c = tq + u * TJS;
                if (c > j && c < k) {
                    const float lc = sL[j * KP + c];
                    T[u] = __fadd_rn(T[u], __fmul_rn(li, lc));
                    Td[u] = __fadd_rn(Td[u], __fmul_rn(lc, lc));
                }
            }
        }
        __syncthreads();
        for (int j = tid; j < k; j += kPrepThreads) sL[j * KP + j] = sDiag[j];
        __syncthreads();
    }
    for (int e = tid; e < KP * KP; e += kPrepThreads) {
        const int i = e % KP, j = e / KP;                  // e = j*KP + i  -> element (row i, col j)
        const bool in_block = (i / 4) == (j / 4);
        if (solver == 0) {
            M1[e] = sL[e];
        } else {
            M1[e] = (i > j && i < k && !in_block) ? sL[e] : 0.f;            // strictly lower, col-major
            // M2[p*KP + c] = L(p, c) for c < p: slot row p = j, c = i
            M2[e] = (i < j && j < k && !in_block) ? sL[i * KP + j] : 0.f;
        }
    }
    for (int e = tid; e < KP * 4; e += kPrepThreads) {     // dblk[q][a][b] = M(4q+a, 4q+b)
        const int q = e / 16, a = (e / 4) % 4, b = e % 4;
        const int i = q * 4 + a, j = q * 4 + b;
        float v;
        if (solver == 0) v = (i < k && j < k) ? sL[j * KP + i] : 0.f;
        else v = (i < k && j < k) ? ((j <= i) ? sL[j * KP + i] : 0.f) : ((i == j) ? 1.f : 0.f);
        dblk[e] = v;
    }
    for (int i = tid; i < KP; i += kPrepThreads) {
        const float dg = (i < k) ? sL[i * KP + i] : ((solver == 0) ? 0.f : 1.f);
        rcp[i] = (dg > 0.f) ? __frcp_rn(dg) : 0.f;
    }
}

// d_i = float(sums_i) (+sqrt for L2) + 1e-15 (variant_helpers.hpp:297-301); padded d_i = 1.
// `sums` are the fp64 row sums over all columns (all-reduced over ranks when sharded).
static __global__ void scale_from_sums_kernel(const double* __restrict__ sums, int KP, int k, int norm_type,
                                              float* __restrict__ d, const int* __restrict__ stop_flag) {
    if (*stop_flag) return;
    const int i = threadIdx.x;
    if (i >= KP) return;
    if (i >= k || norm_type == 2) { d[i] = 1.f; return; }
    float v = static_cast<float>(sums[i]);
    if (norm_type == 1) v = __fsqrt_rn(v);
    d[i] = __fadd_rn(v, 1e-15f);
}

// loss = trAtA − 2·cross + Σ_ij d_i d_j G_wt(i,j) G_h(i,j)   (fit_cpu.hpp:1747-1753), then the
// convergence / patience bookkeeping of fit_cpu.hpp:1769-1811. One CTA of 256 threads.
static __global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ G_wt, const float* __restrict__ G_h,
                                                   const float* __restrict__ d, int KP, int k,
                                                   const double* __restrict__ cross_partials, int nparts,
                                                   float trAtA, float tol, int patience,
                                                   float* __restrict__ loss_hist, int hist_cap,
                                                   DevState* __restrict__ st) {
    __shared__ double sred[256];
    if (st->stop) return;
    double r = 0.0;
    for (int e = threadIdx.x; e < KP * KP; e += blockDim.x) {
        const int i = e % KP, j = e / KP;
        if (i < k && j < k) {
            const float t = __fmul_rn(__fmul_rn(__fmul_rn(d[i], d[j]), G_wt[e]), G_h[e]);
            r += static_cast<double>(t);
        }
    }
    sred[threadIdx.x] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        double recon = 0.0;
        for (int t = 0; t < 256; ++t) recon += sred[t];
        double cross = 0.0;
        for (int c = 0; c < nparts; ++c) cross += cross_partials[c];
        const float loss = __fadd_rn(__fsub_rn(trAtA, __fmul_rn(2.f, static_cast<float>(cross))),
                                     static_cast<float>(recon));
        const int iter = st->iter;
        if (iter < hist_cap) loss_hist[iter] = loss;
        bool loss_conv = false;
        if (iter > 0) {
            const float rel = __fdiv_rn(fabsf(__fsub_rn(st->prev_loss, loss)),
                                        __fadd_rn(fabsf(st->prev_loss), 1e-15f));
            st->final_tol = rel;
            if (rel < tol) loss_conv = true;
        }
        st->prev_loss = loss;
        st->train_loss = loss;
        st->iter = iter + 1;
        if (iter > 0) {
            if (loss_conv) {
                if (++st->patience_counter >= patience) {
                    st->converged = 1;
                    st->stop = 1;
                }
            } else {
                st->patience_counter = 0;
            }
        }
    }
}

// ---- small utilities -----------------------------------------------------------------------

// dst[c][0..KP) = (float)src[c][0..k), zero padded.   (double→float at the R boundary:
// src/RcppFunctions_nmf.cpp:4-5; src/gpu_bridge_nmf.cu:578-592)
template <class SrcT>
static __global__ void pad_convert_kernel(const SrcT* __restrict__ src, float* __restrict__ dst, long long ncols, int k,
                                   int KP) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= ncols * KP) return;
    const long long c = e / KP;
    const int i = static_cast<int>(e % KP);
    dst[e] = (i < k) ? static_cast<float>(src[c * k + i]) : 0.f;
}
template <class DstT>
static __global__ void unpad_convert_kernel(const float* __restrict__ src, DstT* __restrict__ dst, long long ncols, int k,
                                     int KP) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= ncols * k) return;
    const long long c = e / k;
    const int i = static_cast<int>(e % k);
    dst[e] = static_cast<DstT>(src[c * KP + i]);
}
static __global__ void f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, long long n) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < n) dst[e] = static_cast<float>(src[e]);
}

__device__ __forceinline__ unsigned long long splitmix_mix(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long splitmix_hash(unsigned long long seed, unsigned i, unsigned j) {
    return splitmix_mix(seed + static_cast<unsigned long long>(i) * 0x9e3779b97f4a7c15ULL +
                        static_cast<unsigned long long>(j) * 0x6c62272e07bb0142ULL);          // rng.hpp:129-138
}
// uniform<float>() = float(u64) / float(UINT64_MAX) ; float(UINT64_MAX) == 2^64 (rng.hpp:102-104)
__device__ __forceinline__ float u64_to_unit_float(unsigned long long z) {
    return __fdiv_rn(__ull2float_rn(z), 18446744073709551616.0f);
}

// Element e (0-based) of the sequential stream SplitMix64(seed): state after e+1 increments is
// seed + (e+1)·γ, so the stream is random-access (rng.hpp:89-95). dst is [ncols][KP] padded;
// stream element of (column c, coordinate i) is first_elem + c·k + i (fill_uniform, rng.hpp:195-201).
static __global__ void init_uniform_kernel(float* __restrict__ dst, long long ncols, int k, int KP,
                                    unsigned long long state0, unsigned long long first_elem) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= ncols * KP) return;
    const long long c = e / KP;
    const int i = static_cast<int>(e % KP);
    if (i >= k) { dst[e] = 0.f; return; }
    const unsigned long long idx = first_elem + static_cast<unsigned long long>(c) * k + i;
    const unsigned long long state = state0 + (idx + 1ULL) * 0x9e3779b97f4a7c15ULL;
    dst[e] = u64_to_unit_float(splitmix_mix(state));
}

// 64-bit checksum of the logical k x ncols factor (padding excluded), independent of the launch geometry: the sum
// (mod 2^64) over elements of mix(position) ^ mix(bits). Equal checksums <=> bit-identical factors (up to 2^-64);
// bench.py and the multi-GPU checks compare sharded fits with the one-GPU fit of the same seed this way.
static __global__ void __launch_bounds__(256) checksum_kernel(const float* __restrict__ X, long long ncols, int k, int KP,
                                                              unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    const long long total = ncols * k;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long c = e / k;
        const int i = static_cast<int>(e % k);
        const unsigned bits = __float_as_uint(X[c * KP + i]);
        acc += splitmix_mix(static_cast<unsigned long long>(e + 1) * 0x9e3779b97f4a7c15ULL) ^
               splitmix_mix(0x5851f42d4c957f2dULL + bits);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// tr(AᵀA) partials (primitives/primitives.hpp:101-115) in fp64; reduced on the host.
static __global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n,
                                                    double* __restrict__ partials) {
    __shared__ double s[256];
    double a = 0.0;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
        a += static_cast<double>(x[e]) * static_cast<double>(x[e]);
    s[threadIdx.x] = a;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = s[0];
}

}  // namespace b200
