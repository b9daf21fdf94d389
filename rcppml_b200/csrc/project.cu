// project.cu — GPU entry points for predict() / nnls() / evaluate() (ABI extensions).
// fp64 throughout, like the reference's CPU implementations (src/RcppFunctions_utils.cpp:23-53, 97-149, 314-366).
#include "engine.hpp"
#include "kernels_project.cuh"

#include <cstdio>
#include <vector>

namespace {

using b200::DeviceBuffer;

// First sm_100+ device (this library carries sm_100a code only; rcppml_gpu_detect counts the same devices), made
// current; its SM count sizes the grids. Throws when there is none.
int select_device(int* num_sms) {
    int count = 0;
    B200_REQUIRE(cudaGetDeviceCount(&count) == cudaSuccess && count > 0, "no CUDA device");
    for (int dev = 0; dev < count; ++dev) {
        cudaDeviceProp prop{};
        if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.major >= 10) {
            B200_CUDA_CHECK(cudaSetDevice(dev));
            if (num_sms) *num_sms = prop.multiProcessorCount;
            return dev;
        }
    }
    throw std::runtime_error("no sm_100+ device");
}

// The stream of one entry-point call: destroyed on every exit path (a thrown B200_REQUIRE / CUDA error included).
struct ScopedStream {
    cudaStream_t s = nullptr;
    ScopedStream() { B200_CUDA_CHECK(cudaStreamCreate(&s)); }
    ~ScopedStream() { if (s) cudaStreamDestroy(s); }
    ScopedStream(const ScopedStream&) = delete;
    ScopedStream& operator=(const ScopedStream&) = delete;
};

int g_num_sms = b200::kNumSMs;     // of the device select_device() made current (entry points are synchronous)

template <int KP>
void launch_gram_f64(const double* F, long long ncols, double* partials, int grid, cudaStream_t s) {
    b200::gram_f64_kernel<KP><<<grid, 256, 0, s>>>(F, ncols, partials);
}
void gram_f64(int KP, int k, const double* F, long long ncols, double add1, double add2, double add3, double* G,
              cudaStream_t s) {
    const int grid = g_num_sms * 2;
    DeviceBuffer<double> partials;
    partials.ensure(static_cast<size_t>(grid) * KP * KP);
    switch (KP) {
        case 16: launch_gram_f64<16>(F, ncols, partials.ptr, grid, s); break;
        case 32: launch_gram_f64<32>(F, ncols, partials.ptr, grid, s); break;
        case 64: launch_gram_f64<64>(F, ncols, partials.ptr, grid, s); break;
        case 128: launch_gram_f64<128>(F, ncols, partials.ptr, grid, s); break;
        default: throw std::runtime_error("unsupported padded rank");
    }
    b200::gram_f64_finalize_kernel<<<(KP * KP + 255) / 256, 256, 0, s>>>(partials.ptr, grid, KP, k, add1, add2, add3, G);
    B200_CUDA_CHECK(cudaStreamSynchronize(s));
}

template <int KP>
void launch_project(const b200::ProjectParams& p, cudaStream_t s) {
    const size_t smem = static_cast<size_t>(KP) * KP * sizeof(double);
    B200_CUDA_CHECK(cudaFuncSetAttribute(b200::project_f64_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
    int occ = 0;
    B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, b200::project_f64_kernel<KP>, 256, smem));
    B200_REQUIRE(occ > 0, "project_f64_kernel does not fit on an SM");
    b200::project_f64_kernel<KP><<<g_num_sms * occ, 256, smem, s>>>(p);
}

struct DeviceCsc64 {
    DeviceBuffer<int> p, i;
    DeviceBuffer<double> x;
    void upload(int ncols, int64_t nnz, const int* cp, const int* ri, const double* v, cudaStream_t s) {
        p.ensure(static_cast<size_t>(ncols) + 1);
        i.ensure(std::max<int64_t>(nnz, 1));
        x.ensure(std::max<int64_t>(nnz, 1));
        B200_CUDA_CHECK(cudaMemcpyAsync(p.ptr, cp, (static_cast<size_t>(ncols) + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
        if (nnz > 0) {
            B200_CUDA_CHECK(cudaMemcpyAsync(i.ptr, ri, nnz * sizeof(int), cudaMemcpyHostToDevice, s));
            B200_CUDA_CHECK(cudaMemcpyAsync(x.ptr, v, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
        }
    }
};

void upload_padded(const double* host, long long ncols, int k, int KP, DeviceBuffer<double>& dst, cudaStream_t s) {
    DeviceBuffer<double> tmp;
    tmp.ensure(static_cast<size_t>(ncols) * k);
    dst.ensure(static_cast<size_t>(ncols) * KP);
    B200_CUDA_CHECK(cudaMemcpyAsync(tmp.ptr, host, static_cast<size_t>(ncols) * k * sizeof(double), cudaMemcpyHostToDevice, s));
    const long long total = ncols * KP;
    b200::pad_f64_kernel<double><<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(tmp.ptr, dst.ptr, ncols, k, KP);
    B200_CUDA_CHECK(cudaStreamSynchronize(s));
}

void warn(const char* what) { std::fprintf(stderr, "[RcppML_gpu/b200] %s\n", what); }

}  // namespace

extern "C" {

// h = argmin_{h>=0} ||A - w^T h|| column by column: Rcpp_predict / c_nnls on the GPU.
//   w_T : k x m column-major (the fixed factor, transposed as R's predict passes it)
//   h   : k x n column-major, output; also the warm start when *warm_start != 0 (c_nnls semantics:
//         B -= G·h, then CD WITHOUT a tolerance — src/RcppFunctions_utils.cpp:346-356)
// G = gram(w) + tiny (gram.hpp:51) + tiny (:33 / :327) + L2; L1 and the upper bound act inside the CD
// sweep (nnls_batch.hpp:94, :104-107). predict(): nonneg=1, cd_maxit=100, cd_tol=1e-8, warm_start=0.
void rcppml_gpu_nnls_double(const int* col_ptr, const int* row_idx, const double* values, int* m, int* n, int* nnz,
                            int* k, const double* w_T, double* h, double* L1, double* L2, double* upper_bound,
                            int* nonneg, int* cd_maxit, double* cd_tol, int* warm_start, int* out_status) {
    if (!out_status) return;
    *out_status = -1;
    try {
        B200_REQUIRE(*k >= 1 && *k <= b200::kMaxKP, "rank must be in [1, 128]");
        B200_REQUIRE(*m > 0 && *n > 0 && *nnz >= 0 && *cd_maxit > 0, "bad dimensions");
        select_device(&g_num_sms);
        ScopedStream stream_guard;
        cudaStream_t s = stream_guard.s;
        const int KP = b200::padded_rank(*k);
        DeviceCsc64 A;
        A.upload(*n, *nnz, col_ptr, row_idx, values, s);
        DeviceBuffer<double> W, H, G;
        upload_padded(w_T, *m, *k, KP, W, s);
        G.ensure(static_cast<size_t>(KP) * KP);
        gram_f64(KP, *k, W.ptr, *m, 1e-15, 1e-15, (*L2 > 0) ? *L2 : 0.0, G.ptr, s);
        if (*warm_start) upload_padded(h, *n, *k, KP, H, s);
        else H.ensure(static_cast<size_t>(*n) * KP);
        DeviceBuffer<int> counter;
        counter.ensure(1);
        B200_CUDA_CHECK(cudaMemsetAsync(counter.ptr, 0, sizeof(int), s));
        b200::ProjectParams p{};
        p.colptr = A.p.ptr; p.rowidx = A.i.ptr; p.vals = A.x.ptr;
        p.F = W.ptr; p.X = H.ptr; p.G = G.ptr;
        p.ncols = *n; p.k = *k;
        p.L1 = *L1; p.ub = *upper_bound; p.cd_tol = *cd_tol;
        p.cd_maxit = *cd_maxit; p.nonneg = *nonneg; p.warm = *warm_start;
        p.work_counter = counter.ptr;
        switch (KP) {
            case 16: launch_project<16>(p, s); break;
            case 32: launch_project<32>(p, s); break;
            case 64: launch_project<64>(p, s); break;
            default: launch_project<128>(p, s); break;
        }
        B200_CUDA_CHECK(cudaGetLastError());
        DeviceBuffer<double> out;
        out.ensure(static_cast<size_t>(*n) * *k);
        const long long total = static_cast<long long>(*n) * *k;
        b200::unpad_f64_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(H.ptr, out.ptr, *n, *k, KP);
        B200_CUDA_CHECK(cudaMemcpyAsync(h, out.ptr, total * sizeof(double), cudaMemcpyDeviceToHost, s));
        B200_CUDA_CHECK(cudaStreamSynchronize(s));
        *out_status = 0;
    } catch (const std::exception& ex) {
        warn(ex.what());
    } catch (...) {
        warn("unknown error");
    }
}

// evaluate(): mean squared error of A ~ W·diag(d)·H (loss = "mse").
//   mask_zeros != 0 : mean over the non-zeros of A (src/RcppFunctions_utils.cpp:68-79, :115-125)
//   mask_zeros == 0 : mean over all m·n entries (:80-85); computed as (||A||² − 2<A, WdH> + Σ d_i d_j G_W G_H)/(m·n)
//                     in fp64 instead of materialising the dense m×n reconstruction.
void rcppml_gpu_evaluate_double(const int* col_ptr, const int* row_idx, const double* values, int* m, int* n, int* nnz,
                                int* k, const double* w_T, const double* d, const double* h, int* mask_zeros,
                                double* out_loss, int* out_status) {
    if (!out_status) return;
    *out_status = -1;
    try {
        B200_REQUIRE(*k >= 1 && *k <= b200::kMaxKP, "rank must be in [1, 128]");
        select_device(&g_num_sms);
        ScopedStream stream_guard;
        cudaStream_t s = stream_guard.s;
        const int KP = b200::padded_rank(*k);
        DeviceCsc64 A;
        A.upload(*n, *nnz, col_ptr, row_idx, values, s);
        DeviceBuffer<double> W, H, D, part;
        upload_padded(w_T, *m, *k, KP, W, s);
        upload_padded(h, *n, *k, KP, H, s);
        upload_padded(d, 1, *k, KP, D, s);
        const int grid = g_num_sms * 4;
        part.ensure(static_cast<size_t>(grid) * 3);
        b200::eval_nnz_kernel<<<grid, 256, 0, s>>>(A.p.ptr, A.i.ptr, A.x.ptr, *n, KP, *k, W.ptr, H.ptr, D.ptr, part.ptr);
        std::vector<double> hp(static_cast<size_t>(grid) * 3);
        B200_CUDA_CHECK(cudaMemcpyAsync(hp.data(), part.ptr, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        B200_CUDA_CHECK(cudaStreamSynchronize(s));
        double sq = 0, cross = 0, aa = 0;
        for (int c = 0; c < grid; ++c) { sq += hp[c * 3]; cross += hp[c * 3 + 1]; aa += hp[c * 3 + 2]; }
        if (*mask_zeros) {
            *out_loss = (*nnz > 0) ? sq / static_cast<double>(*nnz) : 0.0;
        } else {
            DeviceBuffer<double> Gw, Gh;
            Gw.ensure(static_cast<size_t>(KP) * KP);
            Gh.ensure(static_cast<size_t>(KP) * KP);
            gram_f64(KP, *k, W.ptr, *m, 0.0, 0.0, 0.0, Gw.ptr, s);
            gram_f64(KP, *k, H.ptr, *n, 0.0, 0.0, 0.0, Gh.ptr, s);
            std::vector<double> gw(static_cast<size_t>(KP) * KP), gh(gw.size());
            B200_CUDA_CHECK(cudaMemcpy(gw.data(), Gw.ptr, gw.size() * sizeof(double), cudaMemcpyDeviceToHost));
            B200_CUDA_CHECK(cudaMemcpy(gh.data(), Gh.ptr, gh.size() * sizeof(double), cudaMemcpyDeviceToHost));
            double recon = 0.0;
            for (int i = 0; i < *k; ++i)
                for (int j = 0; j < *k; ++j) recon += d[i] * d[j] * gw[static_cast<size_t>(j) * KP + i] * gh[static_cast<size_t>(j) * KP + i];
            *out_loss = (aa - 2.0 * cross + recon) / (static_cast<double>(*m) * static_cast<double>(*n));
        }
        *out_status = 0;
    } catch (const std::exception& ex) {
        warn(ex.what());
    } catch (...) {
        warn("unknown error");
    }
}

}  // extern "C"
