// kernels_cd.cuh — the fused half-step kernel specialised for the coordinate-descent solver
// (solver_mode 0): same gather, same arithmetic contract and same results (bit for bit, sweep counts
// included) as half_step_kernel<.., SOLVER_CD, ..> in kernels_solve.cuh, re-cut for what bounds CD.
// cd_solve_blocked() below is also the CD solver of tiled_half_step_kernel (kernels_tiled.cuh), which is what a fit
// runs by default when columns are plentiful; cd_half_step_kernel (one narrow geometry for gather and solve) remains
// selectable with RCPPML_B200_TILED=0.
//
// Replaces (reference): primitives/cpu/fused_nnls.hpp:71-134 and primitives/cpu/nnls_batch.hpp:71-132.
//
// Why a second kernel. A CD column costs sweeps x k coordinate steps (C4: ~80 x 64), each a short
// group-uniform decision (pivot / divide / clamp) followed by a k-long residual update b -= G(:,i)·Δ. In the
// Cholesky-tuned geometry (16 or 8 lanes per column) the uniform part is re-executed by every lane of the
// group and dominates the issue slots: ≈21 (16x1) / ≈12 (8x2) warp instructions per column-coordinate, the
// kernel is issue bound, and the gather is < 10 % of the time. So:
//   * narrow lane groups: 2 lanes x 8 words (k = 64) put 16 columns in a warp — the uniform part is shared
//     by 16 columns instead of 2, ≈4.5 instructions per column-coordinate;
//   * pivots in blocks of 4 (like chol_solve): the 4 right-hand sides and 4 warm-start values of a block are
//     broadcast with 8 independent shuffles, the 4 coordinate steps run on the block's private copy `t`
//     against the 4x4 diagonal block of G in constant memory (G is bitwise symmetric), and the 4 residual
//     updates of the lane-owned words follow as straight-line packed arithmetic. Everything inside a block
//     is branch-free: a step that changes nothing has Δ = ±0, and b − G·(±0) = b exactly (the product is
//     formed as fma(g, Δ, +0) = +0, and x − (+0) = x for every x including −0), so executing it is the
//     reference's `continue`;
//   * the sweep's convergence sum Σ|Δ|/(|x|+1e-15) keeps its order (coordinate order, fp32) but the four
//     IEEE divisions of a block are spread over the lanes of the group and brought back by shuffle; and since
//     the sum of non-negative terms only grows under monotone rounding, once tol_sum/k >= cd_tol the sweep-end
//     test is decided and the quotients are skipped for the rest of the sweep (exact, not a heuristic);
//   * registers hold only the lane's words of b (32 at k = 64): x lives in shared memory (it is touched once per
//     block), the fp64 row sums are per-warp shared-memory arrays, and the pre-L1 right-hand side kept for the
//     loss cross term is parked in global memory (L2) — 3 CTAs per SM instead of 2;
//   * the block loop is rolled (the body is ~6 KB of SASS): unrolled it is 96 KB, 3x the instruction cache.
#pragma once

#include "kernels_solve.cuh"

#include <algorithm>

namespace b200 {

// div_exact with the slow path suppressed for non-positive pivots (their result is discarded by the caller).
__device__ __forceinline__ float div_exact_pos(float a, float d, float r) {
    float q = __fmul_rn(a, r);
    float e = __fmaf_rn(-d, q, a);
    q = __fmaf_rn(e, r, q);
    e = __fmaf_rn(-d, q, a);
    q = __fmaf_rn(e, r, q);
    const float aq = fabsf(q);
    if (!(aq > 1e-30f && aq < 1e30f)) {
        if (a != 0.f && r != 0.f) q = __fdiv_rn(a, d);
    }
    return q;
}

// cd_nnls_col_fixed (nnls_batch.hpp:71-132) with L1 = L2 = upper_bound = 0, blocked by 4 pivots. Same
// operations on every element of b and x, in the same order, as cd_solve() in kernels_solve.cuh.
// b: lane-owned registers. x: the group's row of shared memory `sxg` (KP floats, every lane of the group reads
// the pivot word and every lane writes the identical new value back, so a lane only ever reads what it wrote
// itself or what was published before the __syncwarp that precedes the call).
// The block loop is ROLLED (one copy of the block body, ~400 instructions): the fully unrolled version is
// 96 KB of SASS at k = 64, three times the 32 KB instruction cache, and ncu showed the `no_instruction` stall
// at 17 % of the issue slots. Only the pivot broadcast needs a static register index: a switch on the word.
template <int LANES, int NV>
__device__ __forceinline__ int cd_solve_blocked(const HalfStepParams& p, const float* sG, float* sxg, const float* cD,
                                                const float* cR, int gl, unsigned gmask, float (&b)[NV][4]) {
    constexpr int KP = LANES * 4 * NV;
    static_assert(NV <= 8, "lane-group geometry outside what the pivot switch covers");
    const float4* sG4 = reinterpret_cast<const float4*>(sG);
    float4* sx4 = reinterpret_cast<float4*>(sxg);
    const float4* cD4 = reinterpret_cast<const float4*>(cD);
    const float4* cR4 = reinterpret_cast<const float4*>(cR);
    const int nblocks = (p.k + 3) >> 2;
    const bool nonneg = p.nonneg != 0;
    const bool check = p.cd_tol > 0.f;
    int sweeps = p.cd_maxit;
    for (int it = 0; it < p.cd_maxit; ++it) {
        float tol_sum = 0.f;
        bool decided = false;                                           // tol_sum/k >= cd_tol already: not converged
#pragma unroll 1
        for (int q = 0; q < nblocks; ++q) {
            const int nv = q / LANES, owner = q % LANES;
            const float4 x4 = sx4[q];
            float t[4];
#define B200_CD_PIVOT(v)                                                                     \
    case v:                                                                                  \
        if (v < NV) {                                                                        \
            _Pragma("unroll") for (int e = 0; e < 4; ++e)                                    \
                t[e] = (LANES == 1) ? b[v < NV ? v : 0][e] : gshfl<LANES>(gmask, b[v < NV ? v : 0][e], owner);         \
        }                                                                                    \
        break;
            switch (nv) {
                B200_CD_PIVOT(0) B200_CD_PIVOT(1) B200_CD_PIVOT(2) B200_CD_PIVOT(3)
                B200_CD_PIVOT(4) B200_CD_PIVOT(5) B200_CD_PIVOT(6) B200_CD_PIVOT(7)
                default: t[0] = t[1] = t[2] = t[3] = 0.f; break;
            }
#undef B200_CD_PIVOT
            const float xo[4] = {x4.x, x4.y, x4.z, x4.w};
            const float4 r4 = cR4[q];
            const float rc[4] = {r4.x, r4.y, r4.z, r4.w};
            float ad[4], xn[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 drow = cD4[q * 4 + e];                     // row e of the diagonal block == column e
                const float gd = (e == 0) ? drow.x : (e == 1) ? drow.y : (e == 2) ? drow.z : drow.w;
                const float diff = div_exact_pos(t[e], gd, rc[e]);      // :92 (discarded when G_ii <= 0, :90)
                const float nval = __fadd_rn(xo[e], diff);              // :97
                const bool neg = nonneg && (nval < 0.f);                // :100
                float a = neg ? -xo[e] : diff;                          // :101 / :108
                if (!(gd > 0.f)) a = 0.f;
                float xv = neg ? 0.f : nval;
                xv = (a != 0.f) ? xv : xo[e];                           // nothing changes: x keeps its bits
                ad[e] = a;
                xn[e] = xv;
                sub_scaled4(t, drow, a);                                // the block's own rows of :121-124
            }
            __syncwarp(gmask);                                          // every lane of the group has read x4
            sx4[q] = make_float4(xn[0], xn[1], xn[2], xn[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
#pragma unroll
                for (int nv2 = 0; nv2 < NV; ++nv2) {                    // :121-124 residual update
                    const float4 g = sG4[(q * 4 + e) * (KP / 4) + nv2 * LANES + gl];
                    sub_scaled4(b[nv2], g, ad[e]);
                }
            }
            // :115-118. The sum only grows (non-negative terms, monotone rounding), so once tol_sum/k has reached
            // cd_tol the sweep-end test `tol_sum/k < cd_tol` is decided and the exact quotients are not needed
            // any more: in all but a column's last sweeps that happens in the first block or two.
            if (check && !decided) {
                if (LANES >= 4) {                                       // quotients split over the lanes of the group
                    const int me = gl & 3;
                    const float num = fabsf(me == 0 ? ad[0] : me == 1 ? ad[1] : me == 2 ? ad[2] : ad[3]);
                    const float den = fabsf(me == 0 ? xn[0] : me == 1 ? xn[1] : me == 2 ? xn[2] : xn[3]);
                    const float qv = __fdiv_rn(num, __fadd_rn(den, 1e-15f));
#pragma unroll
                    for (int e = 0; e < 4; ++e) tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, qv, e));
                } else if (LANES == 2) {
                    const float n0 = fabsf(gl == 0 ? ad[0] : ad[1]), n1 = fabsf(gl == 0 ? ad[2] : ad[3]);
                    const float d0 = fabsf(gl == 0 ? xn[0] : xn[1]), d1 = fabsf(gl == 0 ? xn[2] : xn[3]);
                    const float q0 = __fdiv_rn(n0, __fadd_rn(d0, 1e-15f));
                    const float q1 = __fdiv_rn(n1, __fadd_rn(d1, 1e-15f));
                    tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, q0, 0));
                    tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, q0, 1));
                    tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, q1, 0));
                    tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, q1, 1));
                } else {                                                // one thread per column
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        tol_sum = __fadd_rn(tol_sum, __fdiv_rn(fabsf(ad[e]), __fadd_rn(fabsf(xn[e]), 1e-15f)));
                }
                decided = !(__fmul_rn(tol_sum, p.inv_k) < p.cd_tol);
            }
        }
        __syncwarp(gmask);
        if (check && !decided) {                                        // :127-129: tol_sum/k < cd_tol
            sweeps = it + 1;
            break;
        }
    }
    return sweeps;
}

// Shared memory: G (KP x KP floats) | x of the columns in flight, one padded row per lane group
// ([256/LANES][KP+4]: the +4 staggers the groups' pivot words over the banks) | fp64 row sums, one [KP] array
// per warp (the groups of a warp add their column one after the other) | 256 fp64 cross partials.
template <int LANES, int NV>
inline size_t cd_half_step_smem_bytes() {
    constexpr int KP = LANES * 4 * NV;
    return static_cast<size_t>(KP) * KP * sizeof(float) + static_cast<size_t>(256 / LANES) * (KP + 4) * sizeof(float) +
           static_cast<size_t>(8) * KP * sizeof(double) + 256 * sizeof(double);
}

#ifndef B200_CD_MIN_CTAS
#define B200_CD_MIN_CTAS 3
#endif
template <int LANES, int NV>
__global__ void __launch_bounds__(256, B200_CD_MIN_CTAS) cd_half_step_kernel(const HalfStepParams p) {
    constexpr int KP = LANES * 4 * NV;
    constexpr int GPW = 32 / LANES;
    constexpr int NGROUPS = 256 / LANES;
    extern __shared__ __align__(16) float smem[];
    if (*p.stop_flag) return;

    float* sG = smem;
    float* sX = sG + KP * KP;                                               // [NGROUPS][KP+4]
    double* sRS = reinterpret_cast<double*>(sX + NGROUPS * (KP + 4));       // [8][KP]
    double* sCross = sRS + 8 * KP;                                          // [256]
    const float* cD = c_solver[p.cslot].dblk;
    const float* cR = c_solver[p.cslot].rcp;
    const int tid = threadIdx.x;

    {
        const float4* g4 = reinterpret_cast<const float4*>(p.M1);
        float4* s4 = reinterpret_cast<float4*>(sG);
        for (int t = tid; t < KP * KP / 4; t += 256) s4[t] = g4[t];
        for (int t = tid; t < 8 * KP; t += 256) sRS[t] = 0.0;
    }
    __syncthreads();

    const int lane = tid & 31;
    const int gl = lane % LANES;
    const int gw = lane / LANES;
    const unsigned gmask = ((1u << LANES) - 1u) << (gw * LANES);
    float* sxg = sX + (tid / LANES) * (KP + 4);
    double* wrs = sRS + (tid >> 5) * KP;
    double cross = 0.0;
    unsigned long long my_sweeps = 0;

    const int fetch = p.cols_per_fetch * GPW;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(p.work_counter, fetch);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= p.ncols) break;
        for (int c = 0; c < p.cols_per_fetch; ++c) {
            const int jl = base + c * GPW + gw;
            const bool active = jl < p.ncols;                               // group-uniform
            float x[NV][4];
#pragma unroll
            for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                for (int e = 0; e < 4; ++e) x[nv][e] = 0.f;
            if (active) {
                const int j = jl + p.col_offset;
                float b[NV][4];
                const int p0 = p.seg_begin ? __ldg(p.seg_begin + jl) : __ldg(p.colptr + jl);
                const int p1 = p.seg_end ? __ldg(p.seg_end + jl) : __ldg(p.colptr + jl + 1);
                if (p.carry_load) {
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv) {
                        const float4 cv = __ldcg(reinterpret_cast<const float4*>(p.carry + static_cast<size_t>(jl) * KP +
                                                                                  (nv * LANES + gl) * 4));
                        b[nv][0] = cv.x; b[nv][1] = cv.y; b[nv][2] = cv.z; b[nv][3] = cv.w;
                    }
                } else {
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                        for (int e = 0; e < 4; ++e) b[nv][e] = 0.f;
                }
                gather_column<LANES, NV>(p, p0, p1, gl, gmask, b);

                if (p.want_cross) {                                         // park b_raw (fused_nnls.hpp:340-347)
                    float* br = p.braw + static_cast<size_t>(jl) * KP;
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv)
                        __stcg(reinterpret_cast<float4*>(br + (nv * LANES + gl) * 4),
                               make_float4(b[nv][0], b[nv][1], b[nv][2], b[nv][3]));
                }
                if (p.L1 > 0.f) {                                           // fused_nnls.hpp:117
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if ((nv * LANES + gl) * 4 + e < p.k) b[nv][e] = __fsub_rn(b[nv][e], p.L1);
                }

                float* xcol = p.X + static_cast<size_t>(j) * KP;
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) {
                    const float4 xv = *reinterpret_cast<const float4*>(xcol + (nv * LANES + gl) * 4);
                    x[nv][0] = xv.x; x[nv][1] = xv.y; x[nv][2] = xv.z; x[nv][3] = xv.w;
                }
                if (p.warm) warm_start_correct<LANES, NV>(sG, p.k, gl, gmask, x, b);      // fused_nnls.hpp:121-123
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
                    *reinterpret_cast<float4*>(sxg + (nv * LANES + gl) * 4) = make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
                __syncwarp(gmask);
                my_sweeps += cd_solve_blocked<LANES, NV>(p, sG, sxg, cD, cR, gl, gmask, b);   // :126-131
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) {
                    const float4 xv = *reinterpret_cast<const float4*>(sxg + (nv * LANES + gl) * 4);
                    x[nv][0] = xv.x; x[nv][1] = xv.y; x[nv][2] = xv.z; x[nv][3] = xv.w;
                }
                __syncwarp(gmask);                                          // the row is reused by the next column

                if (p.ub > 0.f) {                                           // features/bounds.hpp:38 (post-hoc)
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[nv][e] = fminf(x[nv][e], p.ub);
                }
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
                    *reinterpret_cast<float4*>(xcol + (nv * LANES + gl) * 4) =
                        make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
                for (int q = 0; q < p.npeers; ++q) {                        // replicate to the peers' copies of X
                    float* pc = p.peerX[q] + static_cast<size_t>(j) * KP;
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv)
                        *reinterpret_cast<float4*>(pc + (nv * LANES + gl) * 4) =
                            make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
                }
                if (p.mcX) {                                                // ... or to every replica at once (multicast)
                    float* mc = p.mcX + static_cast<size_t>(j) * KP;
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv)
                        multimem_store4(reinterpret_cast<float4*>(mc + (nv * LANES + gl) * 4), make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]));
                }
                if (p.want_cross) {                                         // Σ_i x_i · b_raw,i
                    const float* br = p.braw + static_cast<size_t>(jl) * KP;
                    double s = 0.0;
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv) {
                        const float4 r = __ldcg(reinterpret_cast<const float4*>(br + (nv * LANES + gl) * 4));
                        s += static_cast<double>(x[nv][0]) * static_cast<double>(r.x);
                        s += static_cast<double>(x[nv][1]) * static_cast<double>(r.y);
                        s += static_cast<double>(x[nv][2]) * static_cast<double>(r.z);
                        s += static_cast<double>(x[nv][3]) * static_cast<double>(r.w);
                    }
                    cross += s;
                }
            }
            // Row sums (fp64): the groups of the warp add their column to the warp's array one after the other
            // (x is zero for a group without a column). ~100 instructions per column against ~25 000 for the solve.
            if (p.norm_type != 2) {
#pragma unroll 1
                for (int g = 0; g < GPW; ++g) {
                    if (gw == g) {
#pragma unroll
                        for (int nv = 0; nv < NV; ++nv) {
                            double2* w2 = reinterpret_cast<double2*>(wrs + (nv * LANES + gl) * 4);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                double2 s = w2[h];
                                const float v0 = x[nv][2 * h], v1 = x[nv][2 * h + 1];
                                if (p.norm_type == 0) {
                                    s.x += static_cast<double>(fabsf(v0));
                                    s.y += static_cast<double>(fabsf(v1));
                                } else {
                                    s.x += static_cast<double>(v0) * static_cast<double>(v0);
                                    s.y += static_cast<double>(v1) * static_cast<double>(v1);
                                }
                                w2[h] = s;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }

    // CTA reduction in fp64, fixed order (warps ascending), one partial per CTA — same layout as half_step_kernel.
    if (p.sweep_counter && gl == 0 && my_sweeps) atomicAdd(p.sweep_counter, my_sweeps);
    sCross[tid] = cross;
    __syncthreads();
    for (int t = tid; t < KP; t += 256) {
        double s = 0.0;
        if (p.norm_type != 2)
            for (int w = 0; w < 8; ++w) s += sRS[w * KP + t];
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + t] = s;
    }
    if (tid == 0) {
        double s = 0.0;
        if (p.want_cross)
            for (int t = 0; t < 256; ++t) s += sCross[t];
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + KP] = s;
    }
}

}  // namespace b200
