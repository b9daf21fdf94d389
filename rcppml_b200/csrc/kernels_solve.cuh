// kernels_solve.cuh — the fused half-step kernel: per sparse column, gather b = F·a_j in CSC
// order, apply L1, warm-start, solve the k×k NNLS (coordinate descent or Cholesky+clip), clip /
// bound, write x_j, and emit the partial sums the rest of the iteration needs (row norms for
// the diagonal scaling, <x, b_raw> for the Gram-trick loss). B is never materialised.
//
// Replaces (reference): primitives/cpu/fused_nnls.hpp:71-134 (CD) and :156-221 (Cholesky+clip),
// primitives/cpu/nnls_batch.hpp:71-132 (cd_nnls_col_fixed), features/bounds.hpp:38, the
// accumulation half of nmf/variant_helpers.hpp:287-305, and fused_nnls.hpp:306-362 (loss cross term).
//
// Arithmetic contract: every per-column operation is performed in the SAME order
// and with the SAME roundings as the CPU path — separate IEEE mul and add (no FMA contraction:
// the reference package is built without FMA), IEEE division, sequential CSC order per
// coordinate. A lane group owns the column; a lane owns 4 consecutive coordinates, so no
// cross-lane reduction ever touches b or x: shuffles only BROADCAST the pivot coordinate.
#pragma once

#include "common.cuh"

namespace b200 {

enum : int { SOLVER_CD = 0, SOLVER_CHOL = 1 };
enum : int { BSRC_GATHER = 0 };   // right-hand sides are always gathered in-kernel (an exchanged-RHS variant existed once)
enum : int { OUT_SOLVE = 0, OUT_RHS = 1 };

struct HalfStepParams {
    // sparse operand (CSC of A for the H-update, CSC of Aᵀ for the W-update)
    const int* __restrict__ colptr;
    const int* __restrict__ rowidx;
    const float* __restrict__ vals;
    // dense operands, leading dimension KP (zero padded)
    const float* __restrict__ F;       // gathered factor [rows][KP]
    float* __restrict__ X;             // solution        [ncols][KP] (in: warm start, out: x)
    // k×k operands prepared by prepare_solver_kernel; "z" = 4×4 diagonal blocks zeroed (CHOL only)
    const float* __restrict__ M1;      // CD: G+L2·I (col-major, full) | CHOL: Lz (strictly-lower L, col-major)
    const float* __restrict__ M2;      // CHOL: LTz[p*KP+i] = L(p,i) for i<p
    const float* __restrict__ dblk;    // [KP/4][4][4] diagonal blocks (CD: of G; CHOL: of L incl. diagonal)
    const float* __restrict__ rcp;     // [KP] RN(1/diag) (0 where the diagonal is <= 0)
    // OUT_RHS without a carry buffer: raw right-hand sides [ncols][KP] (diagnostics)
    float* __restrict__ B;
    int ncols;
    int col_offset;                    // X row of local column 0 (this rank's block of the replicated factor)
    int k;
    float L1;
    float ub;
    float cd_tol;
    float inv_k;
    int cd_maxit;
    int nonneg;
    int warm;
    int norm_type;                     // 0: sum|x|, 1: sum x², 2: none
    int want_cross;
    int cols_per_fetch;
    int* work_counter;
    double* partials;                  // [gridDim.x][KP+1]: Σ|x| (or Σx²) per coordinate, then <x, b_raw>
    const int* stop_flag;
    unsigned long long* sweep_counter; // optional: total CD sweeps (diagnostics)
    // Sharded runs with peer-mapped factors: every solved column is ALSO stored straight into the replicas of
    // X on the other GPUs (NVLink P2P stores), so the "all-gather" of the factor overlaps with the solve.
    float* peerX[7];
    int npeers;
    // ... or, with multicast-bound factors (vmm.hpp, Engine::mc_mode 2): ONE multimem.st per word through mcX, the
    // multicast alias of X, lands in every replica — the sender's NVLink egress carries the block once instead of
    // N-1 times. (The local store stays unconditional: the copy that loops back into this rank's replica carries the
    // same bits, and the hot path of a one-GPU fit keeps its instruction schedule.) nullptr: no multicast.
    float* mcX;
    int cslot;                         // c_solver slot holding this launch's diagonal blocks / reciprocals
    // Row-panel passes (engine.cu build_panels): when the gathered factor is larger than L2, a half-step runs as
    // P launches; pass q gathers only the entries of every column whose rows fall in panel q (a contiguous run,
    // rows are sorted), so the rows it touches stay L2-resident. The running right-hand side is carried between
    // passes in `carry` ([ncols][KP], local column index) — the additions happen in exactly the CSC order of the
    // single-pass kernel, so the result is bit-identical.
    const int* __restrict__ seg_begin; // [ncols] first entry of this pass per column (nullptr: colptr[j])
    const int* __restrict__ seg_end;   // [ncols] one past the last entry of this pass (nullptr: colptr[j+1])
    float* __restrict__ carry;
    int carry_load;                    // start from carry[j] instead of 0 (every pass but the first)
    float* __restrict__ braw;          // cd_half_step_kernel, want_cross: [ncols][KP] parking space for the pre-L1 RHS
};

// Warp-uniform solver operands (the 4x4 diagonal blocks and the pivot reciprocals) live in CONSTANT memory:
// every lane reads the same address, and an LDS.128 of a broadcast address still costs four wavefronts of the
// L1/shared data pipe — the pipe that bounds this kernel (ncu: l1tex__data_pipe_lsu_wavefronts 69-79 %). The
// constant cache serves them off that pipe. One slot per engine instance (prepare_solver copies device->symbol).
constexpr int kConstSlots = 8;
struct SolverConsts {
    float dblk[kMaxKP * 4];
    float rcp[kMaxKP];
};
static __constant__ SolverConsts c_solver[kConstSlots];

template <int LANES>
__device__ __forceinline__ float gshfl(unsigned mask, float v, int src) {
    return __shfl_sync(mask, v, src, LANES);
}
template <int LANES>
__device__ __forceinline__ int gshfl(unsigned mask, int v, int src) {
    return __shfl_sync(mask, v, src, LANES);
}

#ifndef B200_SCALAR_FP32   // default: packed pairs (-DB200_SCALAR_FP32 selects the scalar FMUL + FADD form)
// Blackwell packed fp32 pairs (FFMA2 / FADD2): two IEEE-rounded operations per issue slot. ptxas contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under --fmad=false, which would break the separately-rounded
// contract, so the product is written as fma(a, b, +0): RN(a·b + 0) == RN(a·b) except that an exact-zero product
// comes out as +0 instead of -0 — invisible here, because the accumulators it is added to / subtracted from
// are never -0 (they start at +0, and RN(x + y) = -0 only when x = y = -0).
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long p, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ unsigned long long mul2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(0ULL));
    return r;
}
__device__ __forceinline__ unsigned long long add2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long sub2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void axpy4(float (&acc)[4], float v, const float4& f) {
    const unsigned long long vv = pk2(v, v);
    unsigned long long a0 = pk2(acc[0], acc[1]), a1 = pk2(acc[2], acc[3]);
    a0 = add2_rn(a0, mul2_rn(vv, pk2(f.x, f.y)));
    a1 = add2_rn(a1, mul2_rn(vv, pk2(f.z, f.w)));
    upk2(a0, acc[0], acc[1]);
    upk2(a1, acc[2], acc[3]);
}
__device__ __forceinline__ void sub_scaled4(float (&b)[4], const float4& g, float s) {
    const unsigned long long ss = pk2(s, s);
    unsigned long long b0 = pk2(b[0], b[1]), b1 = pk2(b[2], b[3]);
    b0 = sub2_rn(b0, mul2_rn(pk2(g.x, g.y), ss));
    b1 = sub2_rn(b1, mul2_rn(pk2(g.z, g.w), ss));
    upk2(b0, b[0], b[1]);
    upk2(b1, b[2], b[3]);
}
#else
// acc[e] = acc[e] + v*f[e], separately rounded (matches SSE2 Eigen `b += v * col`).
__device__ __forceinline__ void axpy4(float (&acc)[4], float v, const float4& f) {
    acc[0] = __fadd_rn(acc[0], __fmul_rn(v, f.x));
    acc[1] = __fadd_rn(acc[1], __fmul_rn(v, f.y));
    acc[2] = __fadd_rn(acc[2], __fmul_rn(v, f.z));
    acc[3] = __fadd_rn(acc[3], __fmul_rn(v, f.w));
}
// b[e] = b[e] - g[e]*s
__device__ __forceinline__ void sub_scaled4(float (&b)[4], const float4& g, float s) {
    b[0] = __fsub_rn(b[0], __fmul_rn(g.x, s));
    b[1] = __fsub_rn(b[1], __fmul_rn(g.y, s));
    b[2] = __fsub_rn(b[2], __fmul_rn(g.z, s));
    b[3] = __fsub_rn(b[3], __fmul_rn(g.w, s));
}
#endif

// 128-bit load of a factor-row word on the read-only path. -DB200_GATHER_NOALLOC selects
// ld.global.nc.L1::no_allocate (rows are used once per SM; don't let them evict the CSC segments).
__device__ __forceinline__ float4 ldg_row(const float4* ptr) {
#ifdef B200_GATHER_NOALLOC
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr));
    return r;
#else
    return __ldg(ptr);
#endif
}

// ---------------------------------------------------------------------------------------------
// Gather: b = Σ_p vals[p] · F[:, rowidx[p]] for p in [p0, p1), CSC order, per-lane coordinates.
// idx/val are fetched LANES at a time (one coalesced segment per group) and broadcast by shuffle;
// up to UN independent 128-bit row loads are in flight per lane.
// ---------------------------------------------------------------------------------------------
template <int LANES, int NV>
__device__ __forceinline__ void gather_column(const HalfStepParams& p, int p0, int p1, int gl, unsigned gmask,
                                              float (&b)[NV][4]) {
    constexpr int KP = LANES * 4 * NV;
#ifndef B200_GATHER_UN
#define B200_GATHER_UN 8
#endif
    constexpr int UNW = B200_GATHER_UN;                                          // 128-bit loads in flight per lane
    constexpr int UN = (LANES * NV <= UNW) ? LANES : (UNW / NV > 0 ? UNW / NV : 1);
    // b is initialised by the caller (zeros, or the running sums carried over from the previous row panel)

    int nidx = 0;
    float nval = 0.f;
    if (p0 + gl < p1) {
        nidx = __ldg(p.rowidx + p0 + gl);
        nval = __ldg(p.vals + p0 + gl);
    }
#ifdef B200_PREFETCH
    // Two-deep index pipeline: segment s+2 of (idx, val) is requested while segment s is consumed, so that the
    // row addresses of segment s+1 are known one whole segment early and can be prefetched
    // (B200_PREFETCH=1: first half of the next segment into L1; =2: the whole next segment into L2).
    int nnidx = 0;
    float nnval = 0.f;
    if (p0 + LANES + gl < p1) {
        nnidx = __ldg(p.rowidx + p0 + LANES + gl);
        nnval = __ldg(p.vals + p0 + LANES + gl);
    }
#endif
    const float4* Fl = reinterpret_cast<const float4*>(p.F) + gl;
    for (int base = p0; base < p1; base += LANES) {
        const int ridx = nidx;
        const float rval = nval;
#ifdef B200_PREFETCH
        nidx = nnidx;
        nval = nnval;
        const int nb = base + 2 * LANES + gl;        // request segment s+2
        nnidx = 0;
        nnval = 0.f;
        if (nb < p1) {
            nnidx = __ldg(p.rowidx + nb);
            nnval = __ldg(p.vals + nb);
        }
        if (base + LANES + gl < p1 && (B200_PREFETCH == 2 || gl < LANES / 2)) {     // rows of segment s+1
            const char* rowp = reinterpret_cast<const char*>(p.F) + static_cast<size_t>(nidx) * (KP * 4);
#pragma unroll
            for (int l = 0; l < KP * 4; l += 128) {
                if (B200_PREFETCH == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + l));
                else asm volatile("prefetch.global.L1 [%0];" ::"l"(rowp + l));
            }
        }
#else
        const int nb = base + LANES + gl;            // prefetch the next idx/val segment
        nidx = 0;
        nval = 0.f;
        if (nb < p1) {
            nidx = __ldg(p.rowidx + nb);
            nval = __ldg(p.vals + nb);
        }
#endif
        const int cnt = p1 - base;
        if (cnt >= LANES) {                          // full batch: no predication at all
#pragma unroll
            for (int s = 0; s < LANES; s += UN) {
                float4 f[UN][NV];
                float v[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int r = gshfl<LANES>(gmask, ridx, s + u);
                    v[u] = gshfl<LANES>(gmask, rval, s + u);
                    const float4* row = Fl + static_cast<size_t>(r) * (KP / 4);
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv) f[u][nv] = ldg_row(row + nv * LANES);
                }
#pragma unroll
                for (int u = 0; u < UN; ++u)
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv) axpy4(b[nv], v[u], f[u][nv]);
            }
        } else {                                     // ragged tail (< LANES entries), element by element
            for (int s = 0; s < cnt; ++s) {
                const int r = gshfl<LANES>(gmask, ridx, s);
                const float v = gshfl<LANES>(gmask, rval, s);
                const float4* row = Fl + static_cast<size_t>(r) * (KP / 4);
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) axpy4(b[nv], v, __ldg(row + nv * LANES));
            }
        }
    }
}

// b -= G·x restated as tmp = Σ_i G(:,i)·x_i (sequential in i), b -= tmp  (fused_nnls.hpp:121-123).
template <int LANES, int NV>
__device__ __forceinline__ void warm_start_correct(const float* sG, int k, int gl, unsigned gmask,
                                                   const float (&x)[NV][4], float (&b)[NV][4]) {
    constexpr int KP = LANES * 4 * NV;
    const float4* sG4 = reinterpret_cast<const float4*>(sG);
    float tmp[NV][4];
#pragma unroll
    for (int nv = 0; nv < NV; ++nv)
#pragma unroll
        for (int e = 0; e < 4; ++e) tmp[nv][e] = 0.f;
#pragma unroll
    for (int nv = 0; nv < NV; ++nv) {
#pragma unroll 1
        for (int owner = 0; owner < LANES; ++owner) {
            const int i0 = (nv * LANES + owner) * 4;
            if (i0 >= k) break;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xi = gshfl<LANES>(gmask, x[nv][e], owner);
                if (i0 + e < k) {
#pragma unroll
                    for (int nv2 = 0; nv2 < NV; ++nv2) {
                        const float4 g = sG4[(i0 + e) * (KP / 4) + nv2 * LANES + gl];
                        axpy4(tmp[nv2], xi, g);   // tmp += g*xi (commutative product, same rounding)
                    }
                }
            }
        }
    }
#pragma unroll
    for (int nv = 0; nv < NV; ++nv)
#pragma unroll
        for (int e = 0; e < 4; ++e) b[nv][e] = __fsub_rn(b[nv][e], tmp[nv][e]);
}

// cd_nnls_col_fixed (nnls_batch.hpp:71-132) with L1 = L2 = upper_bound = 0 as the fused path calls it.
template <int LANES, int NV>
__device__ __forceinline__ int cd_solve(const HalfStepParams& p, const float* sG, const float* sDblk, const float* sRcp, int gl,
                                        unsigned gmask, float (&x)[NV][4], float (&b)[NV][4]) {
    constexpr int KP = LANES * 4 * NV;
    const float4* sG4 = reinterpret_cast<const float4*>(sG);
    const int k = p.k;
    const bool nonneg = p.nonneg != 0;
    const bool check = p.cd_tol > 0.f;
    int sweeps = p.cd_maxit;
    for (int it = 0; it < p.cd_maxit; ++it) {
        float tol_sum = 0.f;
#pragma unroll
        for (int nv = 0; nv < NV; ++nv) {
#pragma unroll 1
            for (int owner = 0; owner < LANES; ++owner) {
                const int i0 = (nv * LANES + owner) * 4;
                if (i0 >= k) break;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = i0 + e;
                    const float bi = gshfl<LANES>(gmask, b[nv][e], owner);
                    const float xi = gshfl<LANES>(gmask, x[nv][e], owner);
                    const float gd = sDblk[(i >> 2) * 16 + (i & 3) * 5];   // G_ii from the diagonal blocks (0 when padded)
                    float ad = 0.f, xn = xi;
                    if (gd > 0.f) {                                // :90
                        const float diff = div_exact(bi, gd, sRcp[i]);   // :92
                        const float nval = __fadd_rn(xi, diff);    // :97
                        if (nonneg && nval < 0.f) {                // :100-103
                            ad = -xi;
                            xn = 0.f;
                        } else {                                   // :108-112
                            ad = diff;
                            xn = (diff == 0.f) ? xi : nval;
                        }
                    }
                    if (ad != 0.f) {                               // `continue` when nothing changes
                        if (check)                                 // :115-118
                            tol_sum = __fadd_rn(tol_sum, __fdiv_rn(fabsf(ad), __fadd_rn(fabsf(xn), 1e-15f)));
                        if (gl == owner) x[nv][e] = xn;
#pragma unroll
                        for (int nv2 = 0; nv2 < NV; ++nv2) {       // :121-124 residual update
                            const float4 g = sG4[i * (KP / 4) + nv2 * LANES + gl];
                            sub_scaled4(b[nv2], g, ad);
                        }
                    }
                }
            }
        }
        if (check && __fmul_rn(tol_sum, p.inv_k) < p.cd_tol) {     // :127-129
            sweeps = it + 1;
            break;
        }
    }
    return sweeps;
}

// x = L⁻ᵀ L⁻¹ b, column-oriented substitution with IEEE-exact division (Eigen LLT::solve restated,
// fused_nnls.hpp:210), blocked by 4 pivots like cd_solve. sLz / sLTz: strictly-lower L (col-major)
// and its rows (LT[p*KP+i] = L(p,i)) with the 4×4 diagonal blocks zeroed; sDblk: the diagonal
// blocks [q][row][col] (lower triangle incl. the diagonal); sRcp: RN(1/L_pp). On exit b holds x.
template <int LANES, int NV>
__device__ __forceinline__ void chol_solve(const float* sLz, const float* sLTz, const float* sDblk,
                                           const float* sRcp, int k, int gl, unsigned gmask, float (&b)[NV][4]) {
    constexpr int KP = LANES * 4 * NV;
    const float4* sL4 = reinterpret_cast<const float4*>(sLz);
    const float4* sLT4 = reinterpret_cast<const float4*>(sLTz);
    const float4* sD4 = reinterpret_cast<const float4*>(sDblk);
    const float4* sR4 = reinterpret_cast<const float4*>(sRcp);
    // forward: L y = b
#pragma unroll
    for (int nv = 0; nv < NV; ++nv) {
#pragma unroll 1
        for (int owner = 0; owner < LANES; ++owner) {
            const int q = nv * LANES + owner;
            if (q * 4 >= k) break;
            float t[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) t[e] = gshfl<LANES>(gmask, b[nv][e], owner);
            const float4 d1 = sD4[q * 4 + 1], d2 = sD4[q * 4 + 2], d3 = sD4[q * 4 + 3];
            const float d00 = sDblk[q * 16];
            const float4 r = sR4[q];
            float y[4];
            y[0] = div_exact(t[0], d00, r.x);
            t[1] = __fsub_rn(t[1], __fmul_rn(d1.x, y[0]));
            y[1] = div_exact(t[1], d1.y, r.y);
            t[2] = __fsub_rn(t[2], __fmul_rn(d2.x, y[0]));
            t[2] = __fsub_rn(t[2], __fmul_rn(d2.y, y[1]));
            y[2] = div_exact(t[2], d2.z, r.z);
            t[3] = __fsub_rn(t[3], __fmul_rn(d3.x, y[0]));
            t[3] = __fsub_rn(t[3], __fmul_rn(d3.y, y[1]));
            t[3] = __fsub_rn(t[3], __fmul_rn(d3.z, y[2]));
            y[3] = div_exact(t[3], d3.w, r.w);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
#pragma unroll
                for (int nv2 = 0; nv2 < NV; ++nv2) {
                    if (nv2 >= nv) {                                // rows above the pivot block are zero
                        const float4 l = sL4[(q * 4 + e) * (KP / 4) + nv2 * LANES + gl];
                        sub_scaled4(b[nv2], l, y[e]);               // zero on rows <= block: exact no-op
                    }
                }
            }
            if (gl == owner) {
#pragma unroll
                for (int e = 0; e < 4; ++e) b[nv][e] = y[e];
            }
        }
    }
    // backward: Lᵀ x = y   (pivots descending; within a block rows 3,2,1,0)
#pragma unroll
    for (int nv = NV - 1; nv >= 0; --nv) {
#pragma unroll 1
        for (int owner = LANES - 1; owner >= 0; --owner) {
            const int q = nv * LANES + owner;
            if (q * 4 >= k) continue;
            float t[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) t[e] = gshfl<LANES>(gmask, b[nv][e], owner);
            const float4 d1 = sD4[q * 4 + 1], d2 = sD4[q * 4 + 2], d3 = sD4[q * 4 + 3];
            const float d00 = sDblk[q * 16];
            const float4 r = sR4[q];
            float x[4];
            x[3] = div_exact(t[3], d3.w, r.w);
            t[2] = __fsub_rn(t[2], __fmul_rn(d3.z, x[3]));      // y_i -= L(p,i)·x_p for i < p
            t[1] = __fsub_rn(t[1], __fmul_rn(d3.y, x[3]));
            t[0] = __fsub_rn(t[0], __fmul_rn(d3.x, x[3]));
            x[2] = div_exact(t[2], d2.z, r.z);
            t[1] = __fsub_rn(t[1], __fmul_rn(d2.y, x[2]));
            t[0] = __fsub_rn(t[0], __fmul_rn(d2.x, x[2]));
            x[1] = div_exact(t[1], d1.y, r.y);
            t[0] = __fsub_rn(t[0], __fmul_rn(d1.x, x[1]));
            x[0] = div_exact(t[0], d00, r.x);
#pragma unroll
            for (int e = 3; e >= 0; --e) {
#pragma unroll
                for (int nv2 = 0; nv2 < NV; ++nv2) {
                    if (nv2 <= nv) {
                        const float4 l = sLT4[(q * 4 + e) * (KP / 4) + nv2 * LANES + gl];
                        sub_scaled4(b[nv2], l, x[e]);
                    }
                }
            }
            if (gl == owner) {
#pragma unroll
                for (int e = 0; e < 4; ++e) b[nv][e] = x[e];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The kernel. Persistent CTAs; a warp pulls batches of columns from a global counter; each
// LANES-wide lane group handles one column at a time.
//   BSRC_GATHER: b gathered from the CSC operand
//   OUT_SOLVE  : solve and write X;  OUT_RHS: write the raw gathered b to B[0] (partial RHS)
// ---------------------------------------------------------------------------------------------
#ifndef B200_SOLVE_MIN_CTAS
#define B200_SOLVE_MIN_CTAS 3   // <= 85 registers: 3 CTAs (24 warps) per SM; 4 CTAs (64 regs) spills and is no faster
#endif
template <int LANES, int NV, int SOLVER, int BSRC, int OUT>
__global__ void __launch_bounds__(256, (NV >= 4) ? 2 : B200_SOLVE_MIN_CTAS) half_step_kernel(const HalfStepParams p) {
    constexpr int KP = LANES * 4 * NV;
    constexpr int GPW = 32 / LANES;   // groups per warp
    static_assert(BSRC == BSRC_GATHER, "right-hand sides are gathered in-kernel");
    extern __shared__ __align__(16) float smem[];
    if (*p.stop_flag) return;

    float* sM1 = smem;
    float* sM2 = smem + ((OUT == OUT_SOLVE) ? KP * KP : 0);
    float* sEnd = sM2 + ((OUT == OUT_SOLVE && SOLVER == SOLVER_CHOL) ? KP * KP : 0);
    const float* sDblk = c_solver[p.cslot].dblk;              // constant memory (see SolverConsts)
    const float* sRcp = c_solver[p.cslot].rcp;
    double* sRed = reinterpret_cast<double*>(sEnd);     // [256/LANES][KP] norms + [256] cross; 8-byte aligned (KP % 16 == 0)

    if (OUT == OUT_SOLVE) {
        const float4* g4 = reinterpret_cast<const float4*>(p.M1);
        float4* s4 = reinterpret_cast<float4*>(sM1);
        for (int t = threadIdx.x; t < KP * KP / 4; t += blockDim.x) s4[t] = g4[t];
        if (SOLVER == SOLVER_CHOL) {
            const float4* l4 = reinterpret_cast<const float4*>(p.M2);
            float4* t4 = reinterpret_cast<float4*>(sM2);
            for (int t = threadIdx.x; t < KP * KP / 4; t += blockDim.x) t4[t] = l4[t];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int gl = lane % LANES;
    const int gw = lane / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (gw * LANES));

    // Running Σ|x| (or Σx²) over the columns this group solved. fp64: the column→group assignment
    // is dynamic, so only an (effectively) order-independent accumulator keeps d run-to-run stable.
    double rs[NV][4];
#pragma unroll
    for (int nv = 0; nv < NV; ++nv)
#pragma unroll
        for (int e = 0; e < 4; ++e) rs[nv][e] = 0.0;
    double cross = 0.0;
    unsigned long long my_sweeps = 0;

    const int fetch = p.cols_per_fetch * GPW;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(p.work_counter, fetch);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= p.ncols) break;
        for (int c = 0; c < p.cols_per_fetch; ++c) {
            const int jl = base + c * GPW + gw;     // neighbouring groups take neighbouring columns
            if (jl >= p.ncols) continue;            // whole group skips together
            const int j = jl + p.col_offset;

            float b[NV][4];
            if (BSRC == BSRC_GATHER) {
                const int p0 = p.seg_begin ? __ldg(p.seg_begin + jl) : __ldg(p.colptr + jl);    // local to this rank
                const int p1 = p.seg_end ? __ldg(p.seg_end + jl) : __ldg(p.colptr + jl + 1);
                if (p.carry_load) {
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv) {
                        const float4 c = __ldcg(reinterpret_cast<const float4*>(p.carry + static_cast<size_t>(jl) * KP +
                                                                                 (nv * LANES + gl) * 4));
                        b[nv][0] = c.x; b[nv][1] = c.y; b[nv][2] = c.z; b[nv][3] = c.w;
                    }
                } else {
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                        for (int e = 0; e < 4; ++e) b[nv][e] = 0.f;
                }
                gather_column<LANES, NV>(p, p0, p1, gl, gmask, b);
            }

            if (OUT == OUT_RHS) {                   // row-panel pass (carry) or raw right-hand side (B)
                float* dst = p.carry ? p.carry + static_cast<size_t>(jl) * KP : p.B + static_cast<size_t>(j) * KP;
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
                    __stcg(reinterpret_cast<float4*>(dst + (nv * LANES + gl) * 4),
                           make_float4(b[nv][0], b[nv][1], b[nv][2], b[nv][3]));
                continue;
            }

            float braw[NV][4];
            if (p.want_cross) {
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e) braw[nv][e] = b[nv][e];
            }
            if (p.L1 > 0.f) {                                       // fused_nnls.hpp:117 / :202
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((nv * LANES + gl) * 4 + e < p.k) b[nv][e] = __fsub_rn(b[nv][e], p.L1);
            }

            float* xcol = p.X + static_cast<size_t>(j) * KP;
            float x[NV][4];
            if (SOLVER == SOLVER_CD) {
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) {
                    const float4 xv = *reinterpret_cast<const float4*>(xcol + (nv * LANES + gl) * 4);
                    x[nv][0] = xv.x; x[nv][1] = xv.y; x[nv][2] = xv.z; x[nv][3] = xv.w;
                }
                if (p.warm) warm_start_correct<LANES, NV>(sM1, p.k, gl, gmask, x, b);
                my_sweeps += cd_solve<LANES, NV>(p, sM1, sDblk, sRcp, gl, gmask, x, b);
            } else {
                chol_solve<LANES, NV>(sM1, sM2, sDblk, sRcp, p.k, gl, gmask, b);
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float v = b[nv][e];
                        if (p.nonneg && v < 0.f) v = 0.f;           // fused_nnls.hpp:212-214
                        x[nv][e] = v;
                    }
            }
            if (p.ub > 0.f) {                                       // features/bounds.hpp:38 (post-hoc)
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[nv][e] = fminf(x[nv][e], p.ub);
            }
#pragma unroll
            for (int nv = 0; nv < NV; ++nv)
                *reinterpret_cast<float4*>(xcol + (nv * LANES + gl) * 4) =
                    make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
            for (int q = 0; q < p.npeers; ++q) {                    // replicate to the peers' copies of X
                float* pc = p.peerX[q] + static_cast<size_t>(j) * KP;
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
                    *reinterpret_cast<float4*>(pc + (nv * LANES + gl) * 4) =
                        make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
            }
            if (p.mcX) {                                            // ... or to every replica at once (NVSwitch multicast; the
                float* mc = p.mcX + static_cast<size_t>(j) * KP;    // copy that loops back into this replica carries the same bits)
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
                    multimem_store4(reinterpret_cast<float4*>(mc + (nv * LANES + gl) * 4), make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]));
            }

            if (p.norm_type == 0) {
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e) rs[nv][e] += static_cast<double>(fabsf(x[nv][e]));
            } else if (p.norm_type == 1) {
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        rs[nv][e] += static_cast<double>(x[nv][e]) * static_cast<double>(x[nv][e]);
            }
            if (p.want_cross) {                                     // Σ_i x_i · b_raw,i  (x = d∘w_normalised)
                double s = 0.0;
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e) s += static_cast<double>(x[nv][e]) * static_cast<double>(braw[nv][e]);
                cross += s;
            }
        }
    }

    if (OUT == OUT_RHS) return;

    // CTA reduction in fp64 through shared memory in a FIXED order, then one partial per CTA
    // (the finalize kernels sum the per-CTA partials in CTA order).
    constexpr int NGROUPS = 256 / LANES;
    const int grp = threadIdx.x / LANES;
    if (p.norm_type != 2) {
#pragma unroll
        for (int nv = 0; nv < NV; ++nv)
#pragma unroll
            for (int e = 0; e < 4; ++e) sRed[grp * KP + (nv * LANES + gl) * 4 + e] = rs[nv][e];
    }
    double* sCross = sRed + NGROUPS * KP;
    sCross[threadIdx.x] = cross;
    if (p.sweep_counter && gl == 0 && my_sweeps) atomicAdd(p.sweep_counter, my_sweeps);
    __syncthreads();
    if (p.norm_type != 2) {
        for (int t = threadIdx.x; t < KP; t += blockDim.x) {
            double s = 0.0;
            for (int g = 0; g < NGROUPS; ++g) s += sRed[g * KP + t];
            p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + t] = s;
        }
    } else {
        for (int t = threadIdx.x; t < KP; t += blockDim.x) p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + t] = 0.0;
    }
    if (threadIdx.x == 0) {
        double s = 0.0;
        if (p.want_cross)
            for (int t = 0; t < 256; ++t) s += sCross[t];
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + KP] = s;
    }
}

template <int LANES, int NV, int SOLVER, int OUT>
inline size_t half_step_smem_bytes() {
    constexpr int KP = LANES * 4 * NV;
    size_t f = 0;
    if (OUT == OUT_SOLVE) f += static_cast<size_t>(KP) * KP * (SOLVER == SOLVER_CHOL ? 2 : 1);
    return f * sizeof(float) + (static_cast<size_t>(256 / LANES) * KP + 256) * sizeof(double);
}

// Self-test: div_exact vs __fdiv_rn on pseudo-random operands (normal range), bitwise.
static __global__ void selftest_division_kernel(long long n, unsigned long long seed,
                                                unsigned long long* __restrict__ mismatches) {
    unsigned long long bad = 0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        unsigned long long z = seed + static_cast<unsigned long long>(i + 1) * 0x9e3779b97f4a7c15ULL;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        z ^= z >> 31;
        // random mantissas/signs, exponents within 2^±20 of 1.0; every 4th sample uses near-1 mantissas
        unsigned ma = static_cast<unsigned>(z) & 0x007fffffu, md = static_cast<unsigned>(z >> 23) & 0x007fffffu;
        if ((i & 3) == 3) { ma |= 0x007ff000u; md &= 0x00000fffu; }
        const unsigned ea = 107u + (static_cast<unsigned>(z >> 46) % 41u), ed = 107u + (static_cast<unsigned>(z >> 52) % 41u);
        const float a = __uint_as_float(((static_cast<unsigned>(z >> 63) & 1u) << 31) | (ea << 23) | ma);
        const float d = __uint_as_float((ed << 23) | md);
        const float want = __fdiv_rn(a, d);
        const float got = div_exact(a, d, __frcp_rn(d));
        if (__float_as_uint(want) != __float_as_uint(got)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace b200
