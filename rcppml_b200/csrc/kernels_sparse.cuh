// kernels_sparse.cuh — one-time sparse set-up on the device:
//   * CSC(Aᵀ) with ascending inner indices (Eigen's A.transpose(), nmf/fit_cpu.hpp:251-253) by a
//     STABLE radix sort of (row key, position) pairs — stability keeps columns ascending per row
//   * the O(nnz) synthetic matrix generator of SURVEY.md §8d
// CUB (CUDA toolkit header library) provides the radix sort / scans; this is set-up, not the hot loop.
#pragma once

#include "common.cuh"
#include "kernels_dense.cuh"   // splitmix helpers

#include <cub/cub.cuh>

namespace b200 {

// col_of[p] = j such that colptr[j] <= p < colptr[j+1]
static __global__ void expand_columns_kernel(const int* __restrict__ colptr, int n, long long nnz, int* __restrict__ col_of,
                                             int col_id_offset) {
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int lo = 0, hi = n;                 // invariant: colptr[lo] <= p < colptr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(colptr + mid) <= p) lo = mid; else hi = mid;
    }
    col_of[p] = lo + col_id_offset;
}

static __global__ void iota_kernel(unsigned* __restrict__ x, long long n) {
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p < n) x[p] = static_cast<unsigned>(p);
}

// Ti[q] = col_of[perm[q]], Tx[q] = vals[perm[q]]
static __global__ void permute_gather_kernel(const unsigned* __restrict__ perm, const int* __restrict__ col_of,
                                      const float* __restrict__ vals, long long nnz, int* __restrict__ Ti,
                                      float* __restrict__ Tx) {
    const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    const unsigned p = perm[q];
    Ti[q] = col_of[p];
    Tx[q] = vals[p];
}

// Tp[r] = first position q with sorted_rows[q] >= r  (r = 0..m)
static __global__ void row_pointers_kernel(const int* __restrict__ sorted_rows, long long nnz, int m, int* __restrict__ Tp) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > m) return;
    long long lo = 0, hi = nnz;         // first index with key >= r
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (sorted_rows[mid] < r) lo = mid + 1; else hi = mid;
    }
    Tp[r] = static_cast<int>(lo);
}

// --------------------------------------------------------------------------------------------
// Synthetic generator. One CTA per column; candidates are sorted in shared memory (bitonic),
// de-duplicated, counted (pass 0) or written (pass 1).
// --------------------------------------------------------------------------------------------
template <int NMAX>   // power of two >= candidates per column
static __global__ void __launch_bounds__(256) synth_column_kernel(int m, int n_local, int col_begin, int cnt,
                                                           int row_lo, int row_hi,   // keep rows in [row_lo,row_hi), stored relative
                                                           unsigned long long seed, int pass,
                                                           int* __restrict__ counts,
                                                           const int* __restrict__ colptr,
                                                           int* __restrict__ rowidx, float* __restrict__ vals) {
    __shared__ unsigned sk[NMAX];
    __shared__ int sbase[257];
    const int jl = blockIdx.x;
    if (jl >= n_local) return;
    const unsigned j = static_cast<unsigned>(col_begin + jl);
    for (int t = threadIdx.x; t < NMAX; t += 256)
        sk[t] = (t < cnt) ? static_cast<unsigned>(splitmix_hash(seed, static_cast<unsigned>(t), j) %
                                                  static_cast<unsigned long long>(m))
                          : 0xFFFFFFFFu;
    __syncthreads();
    for (int size = 2; size <= NMAX; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < NMAX / 2; t += 256) {
                const int lo = 2 * t - (t & (stride - 1));      // index with the `stride` bit clear
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned a = sk[lo], b = sk[hi];
                if ((a > b) == up) { sk[lo] = b; sk[hi] = a; }
            }
            __syncthreads();
        }
    }
    // unique count per contiguous chunk
    constexpr int CH = NMAX / 256 > 0 ? NMAX / 256 : 1;
    const int c0 = threadIdx.x * CH;
    int uniq = 0;
    for (int t = c0; t < c0 + CH && t < NMAX; ++t) {
        const unsigned v = sk[t];
        if (v != 0xFFFFFFFFu && (t == 0 || sk[t - 1] != v) && v >= static_cast<unsigned>(row_lo) &&
            v < static_cast<unsigned>(row_hi))
            ++uniq;
    }
    sbase[threadIdx.x + 1] = uniq;
    if (threadIdx.x == 0) sbase[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int t = 1; t <= 256; ++t) sbase[t] += sbase[t - 1];
    __syncthreads();
    if (pass == 0) {
        if (threadIdx.x == 0) counts[jl] = sbase[256];
        return;
    }
    const int out0 = colptr[jl];
    int w = out0 + sbase[threadIdx.x];
    for (int t = c0; t < c0 + CH && t < NMAX; ++t) {
        const unsigned v = sk[t];
        if (v != 0xFFFFFFFFu && (t == 0 || sk[t - 1] != v) && v >= static_cast<unsigned>(row_lo) &&
            v < static_cast<unsigned>(row_hi)) {
            rowidx[w] = static_cast<int>(v) - row_lo;
            vals[w] = __fadd_rn(0.5f, u64_to_unit_float(splitmix_hash(seed + 1ULL, v, j)));
            ++w;
        }
    }
}

// Row-panel boundaries inside every column: out[q*ncols + j] = first entry of column j whose row is
// >= q*rows_per_panel (q = 0..P), by binary search over the sorted row indices. out[0][j] = colptr[j],
// out[P][j] = colptr[j+1]. One thread per (q, j).
static __global__ void panel_bounds_kernel(const int* __restrict__ colptr, const int* __restrict__ rowidx, int ncols,
                                           int npanels, int rows_per_panel, int* __restrict__ out) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(npanels + 1) * ncols) return;
    const int q = static_cast<int>(t / ncols), j = static_cast<int>(t % ncols);
    int lo = colptr[j], hi = colptr[j + 1];
    if (q == npanels) { out[t] = hi; return; }
    const long long bound = static_cast<long long>(q) * rows_per_panel;
    while (lo < hi) {                                  // first position with rowidx >= bound
        const int mid = lo + ((hi - lo) >> 1);
        if (rowidx[mid] < bound) lo = mid + 1; else hi = mid;
    }
    out[t] = lo;
}

// --------------------------------------------------------------------------------------------
// In-process multi-GPU ingest (abi_reference.cu, Engine::assemble_row_block): every device uploads only its own
// column block A[:, J_g]; the row block A[I_g, :] each device needs for the W half-step is assembled from the
// column blocks of ALL devices over NVLink peer memory. Rows are ascending inside a CSC column, so the entries of
// column j that fall in rows [r0, r1) are ONE contiguous run, found by two binary searches.
// --------------------------------------------------------------------------------------------
// Per column j of a (peer's) column block: start[j] = first entry with row >= r0, cnt[j] = entries with row in [r0, r1).
static __global__ void rowblock_count_kernel(const int* __restrict__ sp, const int* __restrict__ si, int ncols, int r0,
                                             int r1, int* __restrict__ start, int* __restrict__ cnt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const int p0 = sp[j], p1 = sp[j + 1];
    int lo = p0, hi = p1;
    while (lo < hi) {                                  // first position with row >= r0
        const int mid = lo + ((hi - lo) >> 1);
        if (si[mid] < r0) lo = mid + 1; else hi = mid;
    }
    const int first = lo;
    hi = p1;
    while (lo < hi) {                                  // first position with row >= r1
        const int mid = lo + ((hi - lo) >> 1);
        if (si[mid] < r1) lo = mid + 1; else hi = mid;
    }
    start[j] = first;
    cnt[j] = lo - first;
}

// One warp per column: copies the run found above into the row block's CSC (row ids relative to r0).
static __global__ void __launch_bounds__(256) rowblock_copy_kernel(const int* __restrict__ si, const float* __restrict__ sx,
                                                                   int ncols, const int* __restrict__ start,
                                                                   const int* __restrict__ dst_ptr, int r0,
                                                                   int* __restrict__ di, float* __restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < ncols; j += warps) {
        const int d0 = dst_ptr[j], cnt = dst_ptr[j + 1] - d0, s0 = start[j];
        for (int e = lane; e < cnt; e += 32) {
            di[d0 + e] = si[s0 + e] - r0;
            dx[d0 + e] = sx[s0 + e];
        }
    }
}

// hist[row] += 1 per stored entry (row work of the W half-step, for the balanced row partition)
static __global__ void row_histogram_kernel(const int* __restrict__ si, long long nnz, int* __restrict__ hist) {
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
        atomicAdd(hist + si[e], 1);
}

// work[i] = per_item + sum over devices of hist_g[i]   (hist pointers are peer memory)
struct HistPtrs { const int* p[8]; };
static __global__ void sum_histograms_kernel(HistPtrs h, int ndev, int m, int per_item, long long* __restrict__ work) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    long long s = per_item;
    for (int g = 0; g < ndev; ++g) s += h.p[g][i];
    work[i] = s;
}

// cuts[r] (r = 1..world-1) = smallest j in [0, m] with prefix(j) >= total*r/world, prefix(j) = work of items < j
// (inc is the INCLUSIVE scan: prefix(j) = inc[j-1]); the rule of rcppml_b200/shard.py balanced_cuts. cuts[0] = 0,
// cuts[world] = m. One thread per cut; monotone by construction.
static __global__ void balanced_cuts_kernel(const long long* __restrict__ inc, int m, int world, int* __restrict__ cuts) {
    const int r = threadIdx.x;
    if (r > world) return;
    if (r == 0) { cuts[0] = 0; return; }
    if (r == world) { cuts[world] = m; return; }
    const double total = static_cast<double>(inc[m - 1]);
    const double target = total * static_cast<double>(r) / static_cast<double>(world);
    int lo = 0, hi = m;                                // smallest j with prefix(j) >= target
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        const double pm = (mid == 0) ? 0.0 : static_cast<double>(inc[mid - 1]);
        if (pm < target) lo = mid + 1; else hi = mid;
    }
    cuts[r] = lo;
}

static __global__ void rebase_int_kernel(int* __restrict__ x, int count, int base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) x[i] -= base;
}

}  // namespace b200
