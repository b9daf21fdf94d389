// kernels_tiled.cuh — the fused half-step kernel with SEPARATE lane-group geometries for its two phases.
//
// Replaces (reference): primitives/cpu/fused_nnls.hpp:71-134 (CD) and :156-221 (Cholesky+clip), with
// primitives/cpu/nnls_batch.hpp:71-132, features/bounds.hpp:38, the accumulation half of
// nmf/variant_helpers.hpp:287-305 and the cross term of fused_nnls.hpp:306-362 — same arithmetic contract and the
// same results, bit for bit, as half_step_kernel / cd_half_step_kernel.
//
// Why. The gather wants WIDE lane groups: a group of 8 or 16 lanes reads a factor row as whole 128-byte lines
// (L1 wavefronts, which bound the gather, scale with the number of lines a request touches), and a long column is
// walked by many lanes at once. The k x k solve wants NARROW groups: every pivot step starts with a warp-uniform
// part (broadcast, IEEE divisions, the 4x4 diagonal block) that each lane re-executes, so the fewer lanes a column
// occupies the more columns share it — ncu on the one-geometry kernels: the W half-step of C4 (1M columns of 100
// non-zeros, Cholesky) issues 1600 instructions per column, ~1100 of them in the substitutions; thread-per-column
// CD on skewed data (pbmc3k genes: 1 .. 2700 non-zeros) serialises the long columns behind one lane.
// So a warp works on a batch of CB = 32/SL consecutive columns in three steps:
//   1. gather, ROUNDS x (32/GL columns at a time) in the (GL lanes x GNV words) geometry, right-hand sides into a
//      per-warp shared-memory tile [CB][KP+4];
//   2. solve all CB columns at once in the (SL lanes x SNV words) geometry (k = 64: 2 x 8, 16 columns per warp;
//      k <= 32: one thread per column) — blocked CD (kernels_cd.cuh) or the rolled Cholesky substitution below —
//      and write x back into the tile (for CD the tile row IS x during the sweeps);
//   3. store the CB solved columns with whole-row 128-bit stores (also into the peer replicas when sharded) and
//      accumulate the fp64 row norms in the same pass.
// The pre-L1 right-hand side needed by the loss cross term is parked in global memory (L2) between 2a and 2c.
#pragma once

#include "kernels_cd.cuh"

namespace b200 {

// x = L⁻ᵀ L⁻¹ b — chol_solve() of kernels_solve.cuh with the block loop ROLLED (one copy of the block body; the
// unrolled form is 8 copies per direction at NV = 8, beyond the 32 KB instruction cache). The pivot word and the
// range of words below / above the pivot need static register indices: the pivot broadcast is a switch on the
// word, the trailing update a fall-through switch (Duff) that enters at the pivot's word.
// Same operations on every element, in the same order, as chol_solve().
template <int LANES, int NV>
__device__ __forceinline__ void chol_solve_rolled(const float* sLz, const float* sLTz, const float* cD, const float* cR,
                                                  int k, int gl, unsigned gmask, float (&b)[NV][4]) {
    constexpr int KP = LANES * 4 * NV;
    static_assert(NV <= 8, "word switch covers 8 words");
    const float4* sL4 = reinterpret_cast<const float4*>(sLz);
    const float4* sLT4 = reinterpret_cast<const float4*>(sLTz);
    const float4* cD4 = reinterpret_cast<const float4*>(cD);
    const float4* cR4 = reinterpret_cast<const float4*>(cR);
    const int nblocks = (k + 3) >> 2;

#define B200_PIVOT(v)                                                                                     \
    case v:                                                                                               \
        if (v < NV) {                                                                                     \
            _Pragma("unroll") for (int e = 0; e < 4; ++e)                                                 \
                t[e] = (LANES == 1) ? b[v < NV ? v : 0][e] : gshfl<LANES>(gmask, b[v < NV ? v : 0][e], owner); \
        }                                                                                                 \
        break;
#define B200_STORE(v, src)                                                                                \
    case v:                                                                                               \
        if (v < NV && gl == owner) {                                                                      \
            _Pragma("unroll") for (int e = 0; e < 4; ++e) b[v < NV ? v : 0][e] = src[e];                  \
        }                                                                                                 \
        break;
#define B200_UPDATE(v, M4, src)                                                                           \
    case v:                                                                                               \
        if (v < NV) {                                                                                     \
            _Pragma("unroll") for (int e = 0; e < 4; ++e)                                                 \
                sub_scaled4(b[v < NV ? v : 0], M4[(q * 4 + e) * (KP / 4) + (v < NV ? v : 0) * LANES + gl], src[e]); \
        }

    // forward: L y = b
#pragma unroll 1
    for (int q = 0; q < nblocks; ++q) {
        const int nv = q / LANES, owner = q % LANES;
        float t[4];
        switch (nv) {
            B200_PIVOT(0) B200_PIVOT(1) B200_PIVOT(2) B200_PIVOT(3) B200_PIVOT(4) B200_PIVOT(5) B200_PIVOT(6) B200_PIVOT(7)
            default: t[0] = t[1] = t[2] = t[3] = 0.f; break;
        }
        const float4 d0 = cD4[q * 4], d1 = cD4[q * 4 + 1], d2 = cD4[q * 4 + 2], d3 = cD4[q * 4 + 3];
        const float4 r = cR4[q];
        float y[4];
        y[0] = div_exact(t[0], d0.x, r.x);
        t[1] = __fsub_rn(t[1], __fmul_rn(d1.x, y[0]));
        y[1] = div_exact(t[1], d1.y, r.y);
        t[2] = __fsub_rn(t[2], __fmul_rn(d2.x, y[0]));
        t[2] = __fsub_rn(t[2], __fmul_rn(d2.y, y[1]));
        y[2] = div_exact(t[2], d2.z, r.z);
        t[3] = __fsub_rn(t[3], __fmul_rn(d3.x, y[0]));
        t[3] = __fsub_rn(t[3], __fmul_rn(d3.y, y[1]));
        t[3] = __fsub_rn(t[3], __fmul_rn(d3.z, y[2]));
        y[3] = div_exact(t[3], d3.w, r.w);
        switch (nv) {                                   // rows at and below the pivot block: words nv .. NV-1
            B200_UPDATE(0, sL4, y) B200_UPDATE(1, sL4, y) B200_UPDATE(2, sL4, y) B200_UPDATE(3, sL4, y)
            B200_UPDATE(4, sL4, y) B200_UPDATE(5, sL4, y) B200_UPDATE(6, sL4, y) B200_UPDATE(7, sL4, y)
            default: break;
        }
        switch (nv) {
            B200_STORE(0, y) B200_STORE(1, y) B200_STORE(2, y) B200_STORE(3, y) B200_STORE(4, y) B200_STORE(5, y)
            B200_STORE(6, y) B200_STORE(7, y)
            default: break;
        }
    }
    // backward: Lᵀ x = y   (pivot blocks descending; within a block rows 3,2,1,0)
#pragma unroll 1
    for (int q = nblocks - 1; q >= 0; --q) {
        const int nv = q / LANES, owner = q % LANES;
        float t[4];
        switch (nv) {
            B200_PIVOT(0) B200_PIVOT(1) B200_PIVOT(2) B200_PIVOT(3) B200_PIVOT(4) B200_PIVOT(5) B200_PIVOT(6) B200_PIVOT(7)
            default: t[0] = t[1] = t[2] = t[3] = 0.f; break;
        }
        const float4 d0 = cD4[q * 4], d1 = cD4[q * 4 + 1], d2 = cD4[q * 4 + 2], d3 = cD4[q * 4 + 3];
        const float4 r = cR4[q];
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
        x[0] = div_exact(t[0], d0.x, r.x);
        switch (nv) {                                   // rows at and above the pivot block: words nv .. 0

#define B200_UPDATE_T(v)                                                                                  \
    case v:                                                                                               \
        if (v < NV) {                                                                                     \
            _Pragma("unroll") for (int e = 3; e >= 0; --e)                                                \
                sub_scaled4(b[v < NV ? v : 0], sLT4[(q * 4 + e) * (KP / 4) + (v < NV ? v : 0) * LANES + gl], x[e]); \
        }
            B200_UPDATE_T(7) B200_UPDATE_T(6) B200_UPDATE_T(5) B200_UPDATE_T(4)
            B200_UPDATE_T(3) B200_UPDATE_T(2) B200_UPDATE_T(1) B200_UPDATE_T(0)
#undef B200_UPDATE_T
            default: break;
        }
        switch (nv) {
            B200_STORE(0, x) B200_STORE(1, x) B200_STORE(2, x) B200_STORE(3, x) B200_STORE(4, x) B200_STORE(5, x)
            B200_STORE(6, x) B200_STORE(7, x)
            default: break;
        }
    }
#undef B200_PIVOT
#undef B200_STORE
#undef B200_UPDATE
}

// ---------------------------------------------------------------------------------------------
// Hybrid gather (tiled kernel, WARPS = 24 variant): twice the factor rows in flight per lane group without a register
// more. Of every segment of GL (index, value) pairs the first half of the rows is requested into registers (LDG.128)
// as in gather_column, the second half is requested AT THE SAME TIME into a per-group shared-memory ring with cp.async
// (LDGSTS, 16 bytes per lane, L1 bypassed) — landing space the one-CTA-per-SM layout frees by keeping L / Lᵀ once per
// SM instead of three times. A lane copies exactly the words it later consumes, so no cross-lane synchronisation is
// needed: cp.async.wait_group makes its own copies visible to it. The accumulation order is the CSC order of
// gather_column (entries 0 .. GL-1 of a segment in turn): bit-identical.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int GL, int GNV>
__device__ __forceinline__ void gather_column_hybrid(const HalfStepParams& p, int p0, int p1, int gl, unsigned gmask,
                                                     float* ring /* this group's [GL/2][KP] floats */, float (&b)[GNV][4]) {
    constexpr int KP = GL * 4 * GNV;
    constexpr int HALF = GL / 2;
    static_assert(GL >= 2 && HALF * GNV <= 8, "register half must stay within 8 loads per lane");
    int nidx = 0;
    float nval = 0.f;
    if (p0 + gl < p1) {
        nidx = __ldg(p.rowidx + p0 + gl);
        nval = __ldg(p.vals + p0 + gl);
    }
    const float4* Fl = reinterpret_cast<const float4*>(p.F) + gl;
    float4* ringl = reinterpret_cast<float4*>(ring) + gl;
    for (int base = p0; base < p1; base += GL) {
        const int ridx = nidx;
        const float rval = nval;
        const int nb = base + GL + gl;               // prefetch the next idx/val segment
        nidx = 0;
        nval = 0.f;
        if (nb < p1) {
            nidx = __ldg(p.rowidx + nb);
            nval = __ldg(p.vals + nb);
        }
        const int cnt = p1 - base;
        if (cnt >= GL) {
            float4 f[HALF][GNV];
            float v[GL];
#pragma unroll
            for (int u = 0; u < HALF; ++u) {         // first half: registers
                const int r = gshfl<GL>(gmask, ridx, u);
                v[u] = gshfl<GL>(gmask, rval, u);
                const float4* row = Fl + static_cast<size_t>(r) * (KP / 4);
#pragma unroll
                for (int nv = 0; nv < GNV; ++nv) f[u][nv] = ldg_row(row + nv * GL);
            }
#pragma unroll
            for (int u = HALF; u < GL; ++u) {        // second half: shared-memory ring, in flight together with the first
                const int r = gshfl<GL>(gmask, ridx, u);
                v[u] = gshfl<GL>(gmask, rval, u);
                const float4* row = Fl + static_cast<size_t>(r) * (KP / 4);
#pragma unroll
                for (int nv = 0; nv < GNV; ++nv) cp_async16(ringl + (u - HALF) * (KP / 4) + nv * GL, row + nv * GL);
            }
#pragma unroll
            for (int u = 0; u < HALF; ++u)
#pragma unroll
                for (int nv = 0; nv < GNV; ++nv) axpy4(b[nv], v[u], f[u][nv]);
            cp_async_commit_wait_all();
#pragma unroll
            for (int u = HALF; u < GL; ++u)
#pragma unroll
                for (int nv = 0; nv < GNV; ++nv) axpy4(b[nv], v[u], ringl[(u - HALF) * (KP / 4) + nv * GL]);
        } else {                                     // ragged tail (< GL entries), element by element
            for (int s2 = 0; s2 < cnt; ++s2) {
                const int r = gshfl<GL>(gmask, ridx, s2);
                const float v = gshfl<GL>(gmask, rval, s2);
                const float4* row = Fl + static_cast<size_t>(r) * (KP / 4);
#pragma unroll
                for (int nv = 0; nv < GNV; ++nv) axpy4(b[nv], v, __ldg(row + nv * GL));
            }
        }
    }
}

template <int GL, int GNV, int SL, int SNV, int SOLVER, int WARPS = 8, bool HYB = false>
inline size_t tiled_smem_bytes() {
    constexpr int KP = GL * 4 * GNV;
    constexpr int CB = 32 / SL;
    return static_cast<size_t>(KP) * KP * sizeof(float) * (SOLVER == SOLVER_CHOL ? 2 : 1) +
           static_cast<size_t>(WARPS) * CB * (KP + 4) * sizeof(float) + static_cast<size_t>(WARPS) * 128 * sizeof(double) +
           (HYB ? static_cast<size_t>(WARPS) * (32 / GL) * (GL / 2) * KP * sizeof(float) : 0);
    // (k = 64, Cholesky: 32 KB + 34 KB + 8 KB = 74 KB -> three CTAs per SM; the 256 fp64 cross partials of the
    // final reduction reuse the tile)
}

#ifndef B200_TILED_MIN_CTAS
#define B200_TILED_MIN_CTAS 3
#endif
// WARPS = 8 (default): 256-thread CTAs, three per SM at k = 64. WARPS = 24: ONE 768-thread CTA per SM — L / Lᵀ exist once
// per SM (32 KB instead of 96), which is what pays for the shared-memory ring of the hybrid gather (HYB).
template <int GL, int GNV, int SL, int SNV, int SOLVER, int WARPS = 8, bool HYB = false>
__global__ void __launch_bounds__(WARPS * 32, (WARPS > 8) ? 1 : ((GL * 4 * GNV >= 128) ? (SOLVER == SOLVER_CHOL ? 1 : 2) : B200_TILED_MIN_CTAS)) tiled_half_step_kernel(const HalfStepParams p) {
    constexpr int KP = GL * 4 * GNV;
    static_assert(KP == SL * 4 * SNV, "gather and solve geometries must cover the same padded rank");
    constexpr int GGPW = 32 / GL;                 // gather groups (columns per gather round) per warp
    constexpr int CB = 32 / SL;                   // columns per warp batch = solve groups per warp
    static_assert(CB % GGPW == 0, "a batch must be a whole number of gather rounds");
    constexpr int ROUNDS = CB / GGPW;
    constexpr int PITCH = KP + 4;                 // floats per tile row: staggers the rows over the banks
    constexpr int W4 = KP / 4;                    // 128-bit words per row
    constexpr int RPI = (32 / W4 > 0) ? 32 / W4 : 1;   // rows covered by one warp-wide 128-bit access (KP=128: 1)
    static_assert(W4 <= 32, "a row must fit one warp-wide access");
    extern __shared__ __align__(16) float smem[];
    if (*p.stop_flag) return;

    float* sM1 = smem;
    float* sM2 = sM1 + KP * KP;
    constexpr int THREADS = WARPS * 32;
    float* sTile = sM2 + (SOLVER == SOLVER_CHOL ? KP * KP : 0);                     // [WARPS][CB][PITCH]
    double* sRS = reinterpret_cast<double*>(sTile + WARPS * CB * PITCH);            // [WARPS][RPI][KP]  (RPI*KP = 128)
    float* sRing = reinterpret_cast<float*>(sRS + WARPS * 128);                     // HYB: [WARPS][GGPW][GL/2][KP]
    double* sCross = reinterpret_cast<double*>(sTile);                              // [THREADS], after the main loop only
    const float* cD = c_solver[p.cslot].dblk;
    const float* cR = c_solver[p.cslot].rcp;
    const int tid = threadIdx.x;
    {
        const float4* g4 = reinterpret_cast<const float4*>(p.M1);
        float4* s4 = reinterpret_cast<float4*>(sM1);
        for (int t = tid; t < KP * KP / 4; t += THREADS) s4[t] = g4[t];
        if (SOLVER == SOLVER_CHOL) {
            const float4* l4 = reinterpret_cast<const float4*>(p.M2);
            float4* t4 = reinterpret_cast<float4*>(sM2);
            for (int t = tid; t < KP * KP / 4; t += THREADS) t4[t] = l4[t];
        }
        for (int t = tid; t < WARPS * 128; t += THREADS) sRS[t] = 0.0;
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int ggl = lane % GL, ggw = lane / GL;                                    // gather geometry
    const unsigned ggmask = (GL == 32) ? 0xffffffffu : (((1u << GL) - 1u) << (ggw * GL));
    const int sgl = lane % SL, sgw = lane / SL;                                    // solve geometry
    const unsigned sgmask = ((1u << SL) - 1u) << (sgw * SL);
    const int srow = lane / W4, sw = lane % W4;                                    // store geometry (KP=128: srow 0)
    float* tile = sTile + warp * CB * PITCH;
    float* trow = tile + sgw * PITCH;                                              // this solve group's column
    // fp64 row sums of coordinates sw*4 .. sw*4+3 over the rows this lane stores: a private shared-memory slot
    // (kept out of the register file, which the solve phase needs)
    double2* myrs = reinterpret_cast<double2*>(sRS + (warp * RPI + srow) * KP + sw * 4);
    double cross = 0.0;
    unsigned long long my_sweeps = 0;

    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(p.work_counter, CB);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= p.ncols) break;

        // ---- 1. gather: ROUNDS x GGPW columns in the wide geometry -> tile rows
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            const int cl = r * GGPW + ggw;
            const int jl = base + cl;
            float b[GNV][4];
#pragma unroll
            for (int nv = 0; nv < GNV; ++nv)
#pragma unroll
                for (int e = 0; e < 4; ++e) b[nv][e] = 0.f;
            if (jl < p.ncols) {                                                     // group-uniform
                const int p0 = p.seg_begin ? __ldg(p.seg_begin + jl) : __ldg(p.colptr + jl);
                const int p1 = p.seg_end ? __ldg(p.seg_end + jl) : __ldg(p.colptr + jl + 1);
                if (p.carry_load) {
#pragma unroll
                    for (int nv = 0; nv < GNV; ++nv) {
                        const float4 cv = __ldcg(reinterpret_cast<const float4*>(p.carry + static_cast<size_t>(jl) * KP +
                                                                                  (nv * GL + ggl) * 4));
                        b[nv][0] = cv.x; b[nv][1] = cv.y; b[nv][2] = cv.z; b[nv][3] = cv.w;
                    }
                }
                if constexpr (HYB) gather_column_hybrid<GL, GNV>(p, p0, p1, ggl, ggmask, sRing + ((warp * GGPW + ggw) * (GL / 2)) * KP, b);
                else gather_column<GL, GNV>(p, p0, p1, ggl, ggmask, b);
            }
#pragma unroll
            for (int nv = 0; nv < GNV; ++nv)
                *reinterpret_cast<float4*>(tile + cl * PITCH + (nv * GL + ggl) * 4) =
                    make_float4(b[nv][0], b[nv][1], b[nv][2], b[nv][3]);
        }
        __syncwarp();

        // ---- 2. solve: CB columns in the narrow geometry
        const int jl = base + sgw;
        const bool active = jl < p.ncols;                                           // group-uniform
        float b[SNV][4];
#pragma unroll
        for (int nv = 0; nv < SNV; ++nv) {
            const float4 v = *reinterpret_cast<const float4*>(trow + (nv * SL + sgl) * 4);
            b[nv][0] = v.x; b[nv][1] = v.y; b[nv][2] = v.z; b[nv][3] = v.w;
        }
        __syncwarp();                                                               // rows are reused for x below
        if (active) {
            const int j = jl + p.col_offset;
            if (p.want_cross) {                                                     // park b_raw (fused_nnls.hpp:340-347)
                float* br = p.braw + static_cast<size_t>(jl) * KP;
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv)
                    __stcg(reinterpret_cast<float4*>(br + (nv * SL + sgl) * 4), make_float4(b[nv][0], b[nv][1], b[nv][2], b[nv][3]));
            }
            if (p.L1 > 0.f) {                                                       // fused_nnls.hpp:117 / :202
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((nv * SL + sgl) * 4 + e < p.k) b[nv][e] = __fsub_rn(b[nv][e], p.L1);
            }
            float x[SNV][4];
            if (SOLVER == SOLVER_CD) {
                const float* xcol = p.X + static_cast<size_t>(j) * KP;
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv) {
                    const float4 xv = *reinterpret_cast<const float4*>(xcol + (nv * SL + sgl) * 4);
                    x[nv][0] = xv.x; x[nv][1] = xv.y; x[nv][2] = xv.z; x[nv][3] = xv.w;
                }
                if (p.warm) warm_start_correct<SL, SNV>(sM1, p.k, sgl, sgmask, x, b);         // fused_nnls.hpp:121-123
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv)
                    *reinterpret_cast<float4*>(trow + (nv * SL + sgl) * 4) = make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
                __syncwarp(sgmask);
                my_sweeps += cd_solve_blocked<SL, SNV>(p, sM1, trow, cD, cR, sgl, sgmask, b);   // :126-131
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv) {
                    const float4 xv = *reinterpret_cast<const float4*>(trow + (nv * SL + sgl) * 4);
                    x[nv][0] = xv.x; x[nv][1] = xv.y; x[nv][2] = xv.z; x[nv][3] = xv.w;
                }
            } else {
                chol_solve_rolled<SL, SNV>(sM1, sM2, cD, cR, p.k, sgl, sgmask, b);               // fused_nnls.hpp:210
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float v = b[nv][e];
                        if (p.nonneg && v < 0.f) v = 0.f;                           // fused_nnls.hpp:212-214
                        x[nv][e] = v;
                    }
            }
            if (p.ub > 0.f) {                                                       // features/bounds.hpp:38 (post-hoc)
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv)
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[nv][e] = fminf(x[nv][e], p.ub);
            }
            if (SOLVER != SOLVER_CD || p.ub > 0.f) {
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv)
                    *reinterpret_cast<float4*>(trow + (nv * SL + sgl) * 4) = make_float4(x[nv][0], x[nv][1], x[nv][2], x[nv][3]);
            }
            if (p.want_cross) {                                                     // Σ_i x_i · b_raw,i
                const float* br = p.braw + static_cast<size_t>(jl) * KP;
                double s = 0.0;
#pragma unroll
                for (int nv = 0; nv < SNV; ++nv) {
                    const float4 rr = __ldcg(reinterpret_cast<const float4*>(br + (nv * SL + sgl) * 4));
                    s += static_cast<double>(x[nv][0]) * static_cast<double>(rr.x);
                    s += static_cast<double>(x[nv][1]) * static_cast<double>(rr.y);
                    s += static_cast<double>(x[nv][2]) * static_cast<double>(rr.z);
                    s += static_cast<double>(x[nv][3]) * static_cast<double>(rr.w);
                }
                cross += s;
            }
        }
        __syncwarp();

        // ---- 3. store the batch with whole-row accesses (+ peer replicas) and accumulate the row norms
#pragma unroll 1
        for (int c = 0; c < CB; c += RPI) {
            const int row = c + srow;
            const int jr = base + row;
            if (row < CB && jr < p.ncols) {
                const float4 v = *reinterpret_cast<const float4*>(tile + row * PITCH + sw * 4);
                const size_t off = static_cast<size_t>(jr + p.col_offset) * KP + sw * 4;
                *reinterpret_cast<float4*>(p.X + off) = v;
                for (int q = 0; q < p.npeers; ++q) *reinterpret_cast<float4*>(p.peerX[q] + off) = v;
                if (p.mcX) multimem_store4(reinterpret_cast<float4*>(p.mcX + off), v);      // ... or every replica at once
                if (p.norm_type != 2) {
                    double2 a0 = myrs[0], a1 = myrs[1];
                    if (p.norm_type == 0) {
                        a0.x += static_cast<double>(fabsf(v.x)); a0.y += static_cast<double>(fabsf(v.y));
                        a1.x += static_cast<double>(fabsf(v.z)); a1.y += static_cast<double>(fabsf(v.w));
                    } else {
                        a0.x += static_cast<double>(v.x) * static_cast<double>(v.x);
                        a0.y += static_cast<double>(v.y) * static_cast<double>(v.y);
                        a1.x += static_cast<double>(v.z) * static_cast<double>(v.z);
                        a1.y += static_cast<double>(v.w) * static_cast<double>(v.w);
                    }
                    myrs[0] = a0; myrs[1] = a1;
                }
            }
        }
        __syncwarp();                                                               // tile is rewritten by the next batch
    }

    // CTA reduction in fp64, fixed order, one partial per CTA — same layout as half_step_kernel.
    if (p.sweep_counter && sgl == 0 && my_sweeps) atomicAdd(p.sweep_counter, my_sweeps);
    __syncthreads();                                                                // every warp is done with its tile
    sCross[tid] = cross;
    __syncthreads();
    for (int t = tid; t < KP; t += THREADS) {
        double s = 0.0;
        if (p.norm_type != 2)
            for (int g = 0; g < WARPS * RPI; ++g) s += sRS[g * KP + t];
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + t] = s;
    }
    if (tid == 0) {
        double s = 0.0;
        if (p.want_cross)
            for (int t = 0; t < THREADS; ++t) s += sCross[t];
        p.partials[static_cast<size_t>(blockIdx.x) * (KP + 1) + KP] = s;
    }
}

}  // namespace b200
