// common.cuh — shared helpers for the sm_100a ALS engine.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace b200 {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define B200_CUDA_CHECK(expr)                                                                      \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            throw ::b200::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +    \
                                    " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");       \
    } while (0)

#define B200_REQUIRE(cond, msg)                                                                    \
    do {                                                                                           \
        if (!(cond)) throw std::runtime_error(std::string(msg));                                   \
    } while (0)

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr int kMaxKP = 128;           // largest padded rank with shared-memory resident solver matrices

// Padded rank and lane-group geometry. A sparse column is solved by a group of LANES lanes,
// each lane owning 4 consecutive coordinates (one 128-bit word of a factor row).
inline int lanes_for_rank(int k) {
    if (k <= 16) return 4;
    if (k <= 32) return 8;
    if (k <= 64) return 16;
    return 32;
}
inline int padded_rank(int k) { return lanes_for_rank(k) * 4; }

// IEEE-exact a/d from a correctly rounded reciprocal r = RN(1/d): two FMA correction steps
// (Markstein): after the first q is faithful, after the second it is the correctly rounded
// quotient. 5 instructions instead of the ~15 (MUFU + slow path) of a generic __fdiv_rn; the
// guard hands results outside the comfortable exponent range to __fdiv_rn. Verified bit for bit
// against __fdiv_rn by rcppml_b200_selftest_division (tests/test_gpu_parity.py).
__device__ __forceinline__ float div_exact(float a, float d, float r) {
    float q = __fmul_rn(a, r);
    float e = __fmaf_rn(-d, q, a);
    q = __fmaf_rn(e, r, q);
    e = __fmaf_rn(-d, q, a);
    q = __fmaf_rn(e, r, q);
    const float aq = fabsf(q);
    if (!(aq > 1e-30f && aq < 1e30f)) {
        if (a != 0.f) q = __fdiv_rn(a, d);
    }
    return q;
}

template <class T>
struct DeviceBuffer {
    T* ptr = nullptr;
    size_t count = 0;
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { release(); }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
    }
    // Grow-only allocation; contents are NOT preserved.
    void ensure(size_t n) {
        if (n <= count && ptr) return;
        release();
        if (n == 0) n = 1;
        B200_CUDA_CHECK(cudaMalloc(&ptr, n * sizeof(T)));
        count = n;
    }
    size_t bytes() const { return count * sizeof(T); }
};

}  // namespace b200
