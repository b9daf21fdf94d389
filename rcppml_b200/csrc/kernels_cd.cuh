// kernels_cd.cuh — the fused half-step kernel specialised for the coordinate-descent solver
// (solver_mode 0): same gather, same arithmetic contract and same results (bit for bit, sweep counts
// included) as half_step_kernel<.., SOLVER_CD, ..> in kernels_solve.cuh, re-cut for what bounds CD.
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
//     IEEE divisions of a block are spread over the lanes of the group and brought back by shuffle;
//   * the per-coordinate row sums (fp64) and the pre-L1 right-hand side kept for the loss cross term live in
//     per-thread shared-memory slots instead of registers, which is what lets a lane own 8 words.
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
template <int LANES, int NV>
__device__ __forceinline__ int cd_solve_blocked(const HalfStepParams& p, const float* sG, const float* cD,
                                                const float* cR, int gl, unsigned gmask, float (&x)[NV][4],
                                                float (&b)[NV][4]) {
    constexpr int KP = LANES * 4 * NV;
    static_assert(LANES >= 2, "a lane group needs at least two lanes (tolerance quotients are split over lanes)");
    const float4* sG4 = reinterpret_cast<const float4*>(sG);
    const float4* cD4 = reinterpret_cast<const float4*>(cD);
    const float4* cR4 = reinterpret_cast<const float4*>(cR);
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
                const int q = nv * LANES + owner;
                if (q * 4 >= k) break;
                float t[4], xo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    t[e] = gshfl<LANES>(gmask, b[nv][e], owner);
                    xo[e] = gshfl<LANES>(gmask, x[nv][e], owner);
                }
                const float4 r4 = cR4[q];
                const float rc[4] = {r4.x, r4.y, r4.z, r4.w};
                float ad[4], xn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 drow = cD4[q * 4 + e];             // row e of the diagonal block == column e
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
                if (gl == owner) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[nv][e] = xn[e];
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
#pragma unroll
                    for (int nv2 = 0; nv2 < NV; ++nv2) {                    // :121-124 residual update
                        const float4 g = sG4[(q * 4 + e) * (KP / 4) + nv2 * LANES + gl];
                        sub_scaled4(b[nv2], g, ad[e]);
                    }
                }
                if (check) {                                                // :115-118, quotients split over lanes
                    if (LANES >= 4) {
                        const int me = gl & 3;
                        const float num = fabsf(me == 0 ? ad[0] : me == 1 ? ad[1] : me == 2 ? ad[2] : ad[3]);
                        const float den = fabsf(me == 0 ? xn[0] : me == 1 ? xn[1] : me == 2 ? xn[2] : xn[3]);
                        const float qv = __fdiv_rn(num, __fadd_rn(den, 1e-15f));
#pragma unroll
                        for (int e = 0; e < 4; ++e) tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, qv, e));
                    } else {
                        const float n0 = fabsf(gl == 0 ? ad[0] : ad[1]), n1 = fabsf(gl == 0 ? ad[2] : ad[3]);
                        const float d0 = fabsf(gl == 0 ? xn[0] : xn[1]), d1 = fabsf(gl == 0 ? xn[2] : xn[3]);
                        const float q0 = __fdiv_rn(n0, __fadd_rn(d0, 1e-15f));
                        const float q1 = __fdiv_rn(n1, __fadd_rn(d1, 1e-15f));
                        tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, q0, 0));
                        tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, q0, 1));
                        tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, q1, 0));
                        tol_sum = __fadd_rn(tol_sum, gshfl<LANES>(gmask, q1, 1));
                    }
                }
            }
        }
        if (check && __fmul_rn(tol_sum, p.inv_k) < p.cd_tol) {              // :127-129
            sweeps = it + 1;
            break;
        }
    }
    return sweeps;
}

// Shared memory: G (KP x KP floats) | per-thread fp64 slots for the running Σ|x| / Σx² (NV*4 per thread,
// stored as double2 [NV*2][256]) | when the launch wants the loss cross term: per-thread float4 slots for
// the pre-L1 right-hand side ([NV][256]); that region doubles as the 256 fp64 cross partials at the end.
template <int LANES, int NV>
inline size_t cd_half_step_smem_bytes(bool want_cross) {
    constexpr int KP = LANES * 4 * NV;
    size_t bytes = static_cast<size_t>(KP) * KP * sizeof(float) + static_cast<size_t>(NV) * 2 * 256 * sizeof(double2);
    if (want_cross) bytes += std::max<size_t>(static_cast<size_t>(NV) * 256 * sizeof(float4), 256 * sizeof(double));
    return bytes;
}

template <int LANES, int NV>
__global__ void __launch_bounds__(256, 2) cd_half_step_kernel(const HalfStepParams p) {
    constexpr int KP = LANES * 4 * NV;
    constexpr int GPW = 32 / LANES;
    constexpr int NGROUPS = 256 / LANES;
    extern __shared__ __align__(16) float smem[];
    if (*p.stop_flag) return;

    float* sG = smem;
    double2* sRS = reinterpret_cast<double2*>(smem + KP * KP);              // [NV*2][256]
    float4* sBR = reinterpret_cast<float4*>(sRS + NV * 2 * 256);            // [NV][256]   (want_cross launches only)
    const float* cD = c_solver[p.cslot].dblk;
    const float* cR = c_solver[p.cslot].rcp;
    const int tid = threadIdx.x;

    {
        const float4* g4 = reinterpret_cast<const float4*>(p.M1);
        float4* s4 = reinterpret_cast<float4*>(sG);
        for (int t = tid; t < KP * KP / 4; t += 256) s4[t] = g4[t];
#pragma unroll
        for (int s = 0; s < NV * 2; ++s) sRS[s * 256 + tid] = make_double2(0.0, 0.0);
    }
    __syncthreads();

    const int lane = tid & 31;
    const int gl = lane % LANES;
    const int gw = lane / LANES;
    const unsigned gmask = ((1u << LANES) - 1u) << (gw * LANES);
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
            if (jl >= p.ncols) continue;                                    // whole group skips together
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

            if (p.want_cross) {                                             // park b_raw (fused_nnls.hpp:340-347)
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) sBR[nv * 256 + tid] = make_float4(b[nv][0], b[nv][1], b[nv][2], b[nv][3]);
            }
            if (p.L1 > 0.f) {                                               // fused_nnls.hpp:117
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((nv * LANES + gl) * 4 + e < p.k) b[nv][e] = __fsub_rn(b[nv][e], p.L1);
            }

            float* xcol = p.X + static_cast<size_t>(j) * KP;
            float x[NV][4];
#pragma unroll
            for (int nv = 0; nv < NV; ++nv) {
                const float4 xv = *reinterpret_cast<const float4*>(xcol + (nv * LANES + gl) * 4);
                x[nv][0] = xv.x; x[nv][1] = xv.y; x[nv][2] = xv.z; x[nv][3] = xv.w;
            }
            if (p.warm) warm_start_correct<LANES, NV>(sG, p.k, gl, gmask, x, b);          // fused_nnls.hpp:121-123
            my_sweeps += cd_solve_blocked<LANES, NV>(p, sG, cD, cR, gl, gmask, x, b);     // :126-131

            if (p.ub > 0.f) {                                               // features/bounds.hpp:38 (post-hoc)
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[nv][e] = fminf(x[nv][e], p.ub);
            }
#pragma unroll
            for (int nv = 0; nv < NV; ++nv)
                *reinterpret_cast<float4*>(xcol + (nv * LANES + gl) * 4) =
                    make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
            for (int q = 0; q < p.npeers; ++q) {                            // replicate to the peers' copies of X
                float* pc = p.peerX[q] + static_cast<size_t>(j) * KP;
#pragma unroll
                for (int nv = 0; nv < NV; ++nv)
                    *reinterpret_cast<float4*>(pc + (nv * LANES + gl) * 4) =
                        make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
            }

            if (p.norm_type != 2) {
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double2 s = sRS[(nv * 2 + h) * 256 + tid];
                        const float v0 = x[nv][2 * h], v1 = x[nv][2 * h + 1];
                        if (p.norm_type == 0) {
                            s.x += static_cast<double>(fabsf(v0));
                            s.y += static_cast<double>(fabsf(v1));
                        } else {
                            s.x += static_cast<double>(v0) * static_cast<double>(v0);
                            s.y += static_cast<double>(v1) * static_cast<double>(v1);
                        }
                        sRS[(nv * 2 + h) * 256 + tid] = s;
                    }
                }
            }
            if (p.want_cross) {                                             // Σ_i x_i · b_raw,i
                double s = 0.0;
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) {
                    const float4 br = sBR[nv * 256 + tid];
                    s += static_cast<double>(x[nv][0]) * static_cast<double>(br.x);
                    s += static_cast<double>(x[nv][1]) * static_cast<double>(br.y);
                    s += static_cast<double>(x[nv][2]) * static_cast<double>(br.z);
                    s += static_cast<double>(x[nv][3]) * static_cast<double>(br.w);
                }
                cross += s;
            }
        }
    }

    // CTA reduction in fp64, fixed order (groups ascending), one partial per CTA — same layout as half_step_kernel.
    if (p.sweep_counter && gl == 0 && my_sweeps) atomicAdd(p.sweep_counter, my_sweeps);
    __syncthreads();
    const double* sRSd = reinterpret_cast<const double*>(sRS);
    for (int t = tid; t < KP; t += 256) {
        double s = 0.0;
        if (p.norm_type != 2) {
            const int w = t >> 2, e = t & 3;
            const int nv = w / LANES, l = w % LANES;
            const int slot = nv * 2 + (e >> 1);
            for (int g = 0; g < NGROUPS; ++g) s += sRSd[(static_cast<size_t>(slot) * 256 + g * LANES + l) * 2 + (e & 1)];
        }
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + t] = s;
    }
    if (p.want_cross) {
        double* sCross = reinterpret_cast<double*>(sBR);                    // every thread is past its last b_raw read
        sCross[tid] = cross;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int t = 0; t < 256; ++t) s += sCross[t];
            p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + KP] = s;
        }
    } else if (tid == 0) {
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + KP] = 0.0;
    }
}

}  // namespace b200
