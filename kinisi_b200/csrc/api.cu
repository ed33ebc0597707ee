// Error reporting, device queries and the two roofline micro-benchmarks.
#include <stdarg.h>

#include "common.cuh"

namespace kb2 {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// SM count of the CURRENT device (the launchers size persistent grids from it), cached per device ordinal.
int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    int sm = 148;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm <= 0) sm = 148;
    if (dev >= 0 && dev < 64) cached[dev] = sm;
    return sm;
}
}  // namespace kb2

extern "C" int kb2_abi_version(void) { return KB2_ABI_VERSION; }
extern "C" const char* kb2_last_error(void) { return kb2::g_err; }

extern "C" int kb2_device_info(int* sm_count, int* max_smem_optin, int* cc_major, int* cc_minor) {
    int dev = 0;
    KB2_CUDA(cudaGetDevice(&dev));
    if (sm_count) KB2_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (max_smem_optin) KB2_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (cc_major) KB2_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) KB2_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return 0;
}

// 16 independent DFMA chains per thread: enough ILP to saturate the FP64 pipe with 8 warps/SM.
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double* sink) {
    double a[16];
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * (blockIdx.x + 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (double)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123456.789) sink[0] = s;  // never true; keeps the loop alive
}

extern "C" int kb2_fp64_peak(int grid, int iters, double* sink, void* stream) {
    KB2_CHECK(grid > 0 && iters > 0, "kb2_fp64_peak: bad arguments");
    fp64_peak_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    KB2_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256) stream_copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                                          size_t n2) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) dst[i] = src[i];
}

extern "C" int kb2_stream_copy(const double* src, double* dst, size_t n, void* stream) {
    KB2_CHECK(n % 2 == 0, "kb2_stream_copy: n must be even");
    const int sm = kb2::device_sm_count();
    stream_copy_kernel<<<sm * 8, 256, 0, (cudaStream_t)stream>>>((const double2*)src, (double2*)dst, n / 2);
    KB2_LAUNCH_CHECK();
    return 0;
}
