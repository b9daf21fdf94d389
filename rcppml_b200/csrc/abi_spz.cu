// abi_spz.cu — on-disk ingest: StreamPress v2 `.spz` files into device memory.
//   * rcppml_sp_read_gpu / rcppml_sp_free_gpu: the reference's own entry points (src/sp_gpu_bridge.cu:42-152, called by
//     st_read_gpu / st_free_gpu, R/sp_gpu.R:53-141) — device-resident CSC (int32 pointers and indices, double values)
//     whose addresses travel as doubles into rcppml_gpu_nmf_zerocopy_double. The reference decodes v2 on the CPU and
//     uploads ("v2 adapter: CPU decode + GPU upload", sp_gpu_bridge.cu:20); so does this: the decode is the threaded
//     host reader of spz_reader.cpp, the upload three copies.
//   * rcppml_b200_spz_*: the reader by itself (host buffers; any column range of A or of the stored transpose).
//   * rcppml_b200_set_matrix_spz: file -> engine. One GPU: A and, when the file carries it, the pre-stored transpose
//     (no device transpose). Sharded: this rank decodes ONLY its column block of A and its row block from the
//     transpose section — no rank ever holds the whole matrix.
#include "engine.hpp"
#include "spz_reader.hpp"

#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

struct rcppml_b200_spz {
    b200::spz::File file;
    explicit rcppml_b200_spz(const char* path) : file(path) {}
};

namespace {

using b200::spz::Error;
using b200::spz::File;

// Reader errors keep their status code (the reference's 1..5, spz_reader.hpp); anything else is -1.
template <class F>
int guarded(F&& body) {
    try { body(); return 0; }
    catch (const Error& e) { b200::g_last_error = e.what(); return e.status; }
    catch (const std::exception& e) { b200::g_last_error = e.what(); return -1; }
    catch (...) { b200::g_last_error = "unknown error"; return -1; }
}

// Decode targets are written exactly once, by the decode threads: leave them uninitialised. (A value-initialising
// resize() memsets them first, single-threaded — 80 ms of a 210 ms decode + allocate at 2e7 non-zeros.)
template <class T>
struct NoInit : std::allocator<T> {
    template <class U> struct rebind { using other = NoInit<U>; };
    template <class U, class... A>
    void construct(U* ptr, A&&... args) {
        if constexpr (sizeof...(A) == 0) ::new (static_cast<void*>(ptr)) U;
        else ::new (static_cast<void*>(ptr)) U(std::forward<A>(args)...);
    }
};

struct HostCsc {
    std::vector<int, NoInit<int>> p, i;
    std::vector<float, NoInit<float>> x;
    int64_t nnz = 0;
};

void decode_range(const File& f, int section, uint32_t c0, uint32_t c1, bool reorder, int threads, HostCsc& out) {
    out.nnz = static_cast<int64_t>(f.range_nnz(section, c0, c1));
    out.p.resize(static_cast<size_t>(c1 - c0) + 1);
    out.i.resize(static_cast<size_t>(std::max<int64_t>(out.nnz, 1)));
    out.x.resize(static_cast<size_t>(std::max<int64_t>(out.nnz, 1)));
    f.decode<float>(section, c0, c1, out.p.data(), out.i.data(), out.x.data(), reorder, threads);
}

// A[I, :] for a file WITHOUT a usable transpose section: CSC over all n columns, row ids relative to the block (the
// set_matrix_sharded contract, rcppml_b200/shard.py extract_row_block) — a full decode of A filtered on the host.
void row_block_by_filter(const File& f, int row_begin, int m_loc, int threads, HostCsc& out) {
    const auto& in = f.info();
    const int n = static_cast<int>(in.n);
    HostCsc full;
    decode_range(f, 0, 0, in.n, true, threads, full);
    out.p.assign(static_cast<size_t>(n) + 1, 0);
    out.i.clear(); out.x.clear();
    for (int j = 0; j < n; ++j) {
        for (int q = full.p[j]; q < full.p[j + 1]; ++q) {
            const int r = full.i[q];
            if (r >= row_begin && r < row_begin + m_loc) { out.i.push_back(r - row_begin); out.x.push_back(full.x[q]); }
        }
        out.p[j + 1] = static_cast<int>(out.i.size());
    }
    out.nnz = static_cast<int64_t>(out.i.size());
    if (out.i.empty()) { out.i.push_back(0); out.x.push_back(0.f); }
}

// The stored transpose describes the matrix as written: with a row permutation in play the reference's reader maps
// the row indices of A (sparsepress_v2.hpp:1093-1104) but not the columns of the transpose section (:1318-1467), so
// the two sections no longer describe the same matrix and the transpose is rebuilt on the device instead.
bool transpose_usable(const File& f) {
    const auto& in = f.info();
    return in.transpose_chunks > 0 && in.transpose_nnz == in.nnz && !(in.row_sorted && in.row_permutation_len > 0);
}

void read_gpu(const char* path, int dev, double* out_col_ptr_addr, double* out_row_idx_addr, double* out_values_addr,
              int* out_m, int* out_n, double* out_nnz) {
    File f(path);
    const auto& in = f.info();
    const int64_t nnz = static_cast<int64_t>(in.nnz);
    std::vector<int, NoInit<int>> p(static_cast<size_t>(in.n) + 1), i(static_cast<size_t>(std::max<int64_t>(nnz, 1)));
    std::vector<double, NoInit<double>> x(static_cast<size_t>(std::max<int64_t>(nnz, 1)));
    f.decode<double>(0, 0, in.n, p.data(), i.data(), x.data(), /*reorder=*/true, /*threads=*/0);

    int prev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&prev));
    B200_CUDA_CHECK(cudaSetDevice(dev));
    int* dp = nullptr; int* di = nullptr; double* dx = nullptr;
    auto release = [&] { if (dp) cudaFree(dp); if (di) cudaFree(di); if (dx) cudaFree(dx); cudaSetDevice(prev); };
    try {
        B200_CUDA_CHECK(cudaMalloc(&dp, p.size() * sizeof(int)));
        B200_CUDA_CHECK(cudaMalloc(&di, i.size() * sizeof(int)));
        B200_CUDA_CHECK(cudaMalloc(&dx, x.size() * sizeof(double)));
        B200_CUDA_CHECK(cudaMemcpy(dp, p.data(), p.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (nnz > 0) {
            B200_CUDA_CHECK(cudaMemcpy(di, i.data(), static_cast<size_t>(nnz) * sizeof(int), cudaMemcpyHostToDevice));
            B200_CUDA_CHECK(cudaMemcpy(dx, x.data(), static_cast<size_t>(nnz) * sizeof(double), cudaMemcpyHostToDevice));
        }
    } catch (...) { release(); throw; }
    cudaSetDevice(prev);
    *out_m = static_cast<int>(in.m);
    *out_n = static_cast<int>(in.n);
    *out_nnz = static_cast<double>(nnz);
    *out_col_ptr_addr = static_cast<double>(reinterpret_cast<uintptr_t>(dp));
    *out_row_idx_addr = static_cast<double>(reinterpret_cast<uintptr_t>(di));
    *out_values_addr = static_cast<double>(reinterpret_cast<uintptr_t>(dx));
}

}  // namespace

extern "C" {

// ---- the reference's entry points ---------------------------------------------------------------------------------

void rcppml_sp_read_gpu(const char** path_ptr, int* device_id, double* out_col_ptr_addr, double* out_row_idx_addr,
                        double* out_values_addr, int* out_m, int* out_n, double* out_nnz, int* out_status) {
    *out_status = -1;
    const int st = guarded([&] {
        read_gpu(*path_ptr, *device_id, out_col_ptr_addr, out_row_idx_addr, out_values_addr, out_m, out_n, out_nnz);
    });
    if (st != 0) std::fprintf(stderr, "[sp_read_gpu] %s\n", b200::g_last_error.c_str());
    // 1 cannot open, 2 read failed, 3 too small, 4 not v2, 5 everything the decode can throw (sp_gpu_bridge.cu:57-120)
    *out_status = (st >= 0 && st <= 4) ? st : 5;
}

void rcppml_sp_free_gpu(double* col_ptr_addr, double* row_idx_addr, double* values_addr, int* out_status) {
    *out_status = 0;
    for (double* a : {col_ptr_addr, row_idx_addr, values_addr}) {
        if (*a != 0.0) cudaFree(reinterpret_cast<void*>(static_cast<uintptr_t>(*a)));
        *a = 0.0;
    }
    cudaGetLastError();
}

// The reference's test helper probes the GPU library for `rcppml_st_read_gpu` (tests/testthat/helper-test-utils.R:278)
// while its bridge defines `rcppml_sp_read_gpu`; both names are exported.
void rcppml_st_read_gpu(const char** path_ptr, int* device_id, double* a, double* b, double* c, int* out_m, int* out_n,
                        double* out_nnz, int* out_status) {
    rcppml_sp_read_gpu(path_ptr, device_id, a, b, c, out_m, out_n, out_nnz, out_status);
}
void rcppml_st_free_gpu(double* a, double* b, double* c, int* out_status) { rcppml_sp_free_gpu(a, b, c, out_status); }

// ---- the reader by itself (no device involved) ----------------------------------------------------------------------

int rcppml_b200_spz_open(const char* path, rcppml_b200_spz** out) {
    *out = nullptr;
    return guarded([&] { *out = new rcppml_b200_spz(path); });
}

void rcppml_b200_spz_close(rcppml_b200_spz* h) { delete h; }

int rcppml_b200_spz_get_info(const rcppml_b200_spz* h, rcppml_b200_spz_info* out) {
    return guarded([&] {
        const auto& in = h->file.info();
        std::memset(out, 0, sizeof(*out));
        out->m = static_cast<int>(in.m); out->n = static_cast<int>(in.n); out->nnz = static_cast<int64_t>(in.nnz);
        out->chunk_cols = static_cast<int>(in.chunk_cols); out->num_chunks = static_cast<int>(in.num_chunks);
        out->value_type = in.value_type; out->row_sorted = in.row_sorted;
        out->has_transpose = in.transpose_chunks > 0 || in.transpose_offset != 0;
        out->transpose_chunks = static_cast<int>(in.transpose_chunks);
        out->transp_chunk_cols = static_cast<int>(in.transp_chunk_cols ? in.transp_chunk_cols : in.chunk_cols);
        out->has_obs = in.obs_table_offset != 0; out->has_var = in.var_table_offset != 0;
        out->has_metadata = in.metadata_offset != 0;
        out->row_permutation_len = static_cast<int>(in.row_permutation_len);
        out->density = in.density;
        out->file_bytes = static_cast<int64_t>(in.file_bytes);
        out->transpose_offset = static_cast<int64_t>(in.transpose_offset);
        out->metadata_offset = static_cast<int64_t>(in.metadata_offset);
        out->metadata_bytes = static_cast<int64_t>(in.metadata_bytes);
        out->stored_crc32 = in.footer_crc32;
    });
}

int rcppml_b200_spz_crc32(const rcppml_b200_spz* h, uint32_t* computed) {
    return guarded([&] { *computed = h->file.compute_crc32(); });
}

int rcppml_b200_spz_range_nnz(const rcppml_b200_spz* h, int section, int c0, int c1, int64_t* nnz) {
    return guarded([&] {
        B200_REQUIRE(c0 >= 0 && c1 >= c0, "spz: bad column range");
        *nnz = static_cast<int64_t>(h->file.range_nnz(section, static_cast<uint32_t>(c0), static_cast<uint32_t>(c1)));
    });
}

int rcppml_b200_spz_col_counts(const rcppml_b200_spz* h, int section, int threads, int* counts) {
    return guarded([&] { h->file.col_counts(section, counts, threads); });
}

int rcppml_b200_spz_read_f32(const rcppml_b200_spz* h, int section, int c0, int c1, int reorder, int threads, int* col_ptr,
                             int* row_idx, float* values) {
    return guarded([&] {
        B200_REQUIRE(c0 >= 0 && c1 >= c0, "spz: bad column range");
        h->file.decode<float>(section, static_cast<uint32_t>(c0), static_cast<uint32_t>(c1), col_ptr, row_idx, values,
                              reorder != 0, threads);
    });
}

int rcppml_b200_spz_read_f64(const rcppml_b200_spz* h, int section, int c0, int c1, int reorder, int threads, int* col_ptr,
                             int* row_idx, double* values) {
    return guarded([&] {
        B200_REQUIRE(c0 >= 0 && c1 >= c0, "spz: bad column range");
        h->file.decode<double>(section, static_cast<uint32_t>(c0), static_cast<uint32_t>(c1), col_ptr, row_idx, values,
                               reorder != 0, threads);
    });
}

int rcppml_b200_spz_metadata(const rcppml_b200_spz* h, int key, unsigned char* buf, int64_t capacity, int64_t* bytes) {
    return guarded([&] {
        const auto rec = h->file.metadata_record(static_cast<uint8_t>(key));
        *bytes = static_cast<int64_t>(rec.size());
        if (buf && capacity >= static_cast<int64_t>(rec.size()) && !rec.empty()) std::memcpy(buf, rec.data(), rec.size());
    });
}

// Host half of the sharded ingest of a file without a transpose section (see row_block_by_filter); capacity: entries the
// caller's row_idx / values can take (the file's nnz always suffices). Exposed so that the host logic is testable
// without a device.
int rcppml_b200_spz_row_block_f32(const rcppml_b200_spz* h, int row_begin, int m_loc, int threads, int* col_ptr, int* row_idx,
                                  float* values, int64_t capacity, int64_t* nnz) {
    return guarded([&] {
        const auto& in = h->file.info();
        B200_REQUIRE(row_begin >= 0 && m_loc >= 0 && static_cast<int64_t>(row_begin) + m_loc <= static_cast<int64_t>(in.m),
                     "spz: row block outside the matrix");
        HostCsc t;
        row_block_by_filter(h->file, row_begin, m_loc, threads, t);
        *nnz = t.nnz;
        std::memcpy(col_ptr, t.p.data(), t.p.size() * sizeof(int));
        B200_REQUIRE(capacity >= t.nnz, "spz: row block larger than the caller's buffers");
        if (t.nnz > 0) {
            std::memcpy(row_idx, t.i.data(), static_cast<size_t>(t.nnz) * sizeof(int));
            std::memcpy(values, t.x.data(), static_cast<size_t>(t.nnz) * sizeof(float));
        }
    });
}

// ---- file -> engine -------------------------------------------------------------------------------------------------

int rcppml_b200_set_matrix_spz(rcppml_b200_engine* e, const rcppml_b200_spz* h, int threads, int stored_transpose,
                               int* used_stored_transpose) {
    return guarded([&] {
        const File& f = h->file;
        const auto& in = f.info();
        const int m = static_cast<int>(in.m), n = static_cast<int>(in.n);
        b200::Engine& eng = e->impl;
        // stored_transpose: 1 use the file's transpose section when it is usable, 0 never, < 0 decide here. On ONE GPU
        // the section is not worth decoding: entropy-decoding 1e8 entries costs ~0.5 s of host time, the
        // device transpose of the same matrix 6 ms. Sharded, it is what lets a rank get its row block without decoding
        // (or receiving) the rest of the matrix.
        if (threads <= 0 && eng.world > 1)   // one process per GPU on one host: share the cores between the ranks
            threads = static_cast<int>(std::max(1u, std::thread::hardware_concurrency() / static_cast<unsigned>(eng.world)));
        const bool want = stored_transpose > 0 || (stored_transpose < 0 && eng.world > 1);
        const bool stored = want && transpose_usable(f);
        if (used_stored_transpose) *used_stored_transpose = stored ? 1 : 0;
        HostCsc a, t;
        if (eng.world == 1) {
            decode_range(f, 0, 0, in.n, true, threads, a);
            if (stored) {
                decode_range(f, 1, 0, in.m, false, threads, t);
                eng.set_matrix_host_with_transpose<float>(m, n, a.nnz, a.p.data(), a.i.data(), a.x.data(), t.p.data(),
                                                          t.i.data(), t.x.data());
            } else {
                eng.set_matrix_host<float>(m, n, a.nnz, a.p.data(), a.i.data(), a.x.data());
            }
            return;
        }
        // sharded: this rank's column block of A, and its row block — as columns of the stored transpose when the file
        // has one, else filtered out of a full decode (rows relative to the block, the set_matrix_sharded contract)
        int cb = 0, nl = 0, rb = 0, ml = 0;
        eng.planned_blocks(m, n, &cb, &nl, &rb, &ml);
        decode_range(f, 0, static_cast<uint32_t>(cb), static_cast<uint32_t>(cb + nl), true, threads, a);
        if (stored) {
            decode_range(f, 1, static_cast<uint32_t>(rb), static_cast<uint32_t>(rb + ml), false, threads, t);
            eng.set_matrix_sharded_with_transpose<float>(m, n, a.p.data(), a.i.data(), a.x.data(), t.p.data(), t.i.data(),
                                                         t.x.data());
            return;
        }
        row_block_by_filter(f, rb, ml, threads, t);
        eng.set_matrix_sharded<float>(m, n, a.p.data(), a.i.data(), a.x.data(), t.p.data(), t.i.data(), t.x.data());
    });
}

}  // extern "C"
