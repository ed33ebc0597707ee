// Shared helpers for the kinisi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kinisi_b200.h"

namespace kb2 {

void set_error(const char* fmt, ...);
int device_sm_count();  // SM count of the current device (api.cu)

#define KB2_CHECK(cond, ...)          \
    do {                              \
        if (!(cond)) {                \
            kb2::set_error(__VA_ARGS__); \
            return 1;                 \
        }                             \
    } while (0)

#define KB2_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            kb2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

#define KB2_LAUNCH_CHECK()                                                                       \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            kb2::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 3;                                                                            \
        }                                                                                        \
    } while (0)

constexpr int kWarp = 32;

__device__ __forceinline__ double shfl_xor_f64(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
__device__ __forceinline__ double shfl_up_f64(double v, int delta) { return __shfl_up_sync(0xffffffffu, v, delta); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}

// Inclusive scan across the 32 lanes of a warp.
__device__ __forceinline__ double warp_inclusive_scan(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double up = shfl_up_f64(v, d);
        if (lane >= d) v += up;
    }
    return v;
}

// Philox4x32-10 counter-based generator (Salmon et al., SC'11).
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = (uint64_t)a * (uint64_t)b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    __host__ __device__ static inline uint4 run(uint4 ctr, uint2 key) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(M0, ctr.x, hi0, lo0);
            mulhilo(M1, ctr.z, hi1, lo1);
            ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
            key.x += W0;
            key.y += W1;
        }
        return ctr;
    }
};

// 53-bit uniform in the open interval (0, 1).
__host__ __device__ inline double u01_open(uint32_t a, uint32_t b) {
    uint64_t bits = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace kb2
