// Dependent-chain latency probe for sm_100a (cycles per operation, one warp unless stated).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd(double* out, long long* cyc, double a) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma(float* out, long long* cyc, float a, float b) {
    float x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fmaf(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsqrtf(float* out, long long* cyc) {
    float x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = rsqrtf(x) + 1.0f;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_cvt(double* out, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = (double)((float)x);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_drcp(double* out, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = 1.0 / x;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dsqrt(double* out, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) x = sqrt(x);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(int* out, long long* cyc) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + 33) & 1023;
    __syncthreads();
    int x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = s[x];
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sts_lds(double* out, long long* cyc) {  // store then load back (same thread), dependent
    __shared__ double s[1024];
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        s[threadIdx.x] = x;
        __syncwarp();
        x = s[threadIdx.x ^ 1] + 1.0;
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 5) & 31);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* cyc) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar_sts_lds(double* out, long long* cyc) {  // producer/consumer through smem with a block barrier
    __shared__ double s[1024];
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        s[threadIdx.x] = x;
        __syncthreads();
        x = s[(threadIdx.x + 37) % blockDim.x] + 1.0;
        __syncthreads();
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput of LDS.64/STS.64 from 16 warps (cycles per warp-instruction at the SM level)
__global__ void k_lds_tp(double* out, long long* cyc) {
    __shared__ double s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
    __syncthreads();
    double acc = 0;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) acc += s[(threadIdx.x + i * 32) & 4095];
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 8192 * 8); cudaMalloc(&c, 64); cudaMemset(d, 0x3f, 8192 * 8);
    long long h;
#define RUN(name, call, threads) call; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-28s threads %4d : %.1f cycles/op\n", name, threads, (double)h / N);
    for (int rep = 0; rep < 2; ++rep) {
    RUN("dfma chain", (k_dfma<<<1, 32>>>(d, c, 1.0000001, 1e-9)), 32)
    RUN("dadd chain", (k_dadd<<<1, 32>>>(d, c, 1e-9)), 32)
    RUN("ffma chain", (k_ffma<<<1, 32>>>((float*)d, c, 1.0000001f, 1e-9f)), 32)
    RUN("rsqrtf+fadd chain", (k_rsqrtf<<<1, 32>>>((float*)d, c)), 32)
    RUN("cvt d->f->d chain", (k_cvt<<<1, 32>>>(d, c)), 32)
    RUN("1.0/x double chain", (k_drcp<<<1, 32>>>(d, c)), 32)
    RUN("sqrt double chain", (k_dsqrt<<<1, 32>>>(d, c)), 32)
    RUN("lds pointer chase", (k_lds<<<1, 32>>>((int*)d, c)), 32)
    RUN("sts+syncwarp+lds+dadd", (k_sts_lds<<<1, 32>>>(d, c)), 32)
    RUN("shfl f64 chain", (k_shfl<<<1, 32>>>(d, c)), 32)
    RUN("syncthreads", (k_bar<<<1, 64>>>(d, c)), 64)
    RUN("syncthreads", (k_bar<<<1, 256>>>(d, c)), 256)
    RUN("syncthreads", (k_bar<<<1, 512>>>(d, c)), 512)
    RUN("syncthreads", (k_bar<<<1, 1024>>>(d, c)), 1024)
    RUN("sts+bar+lds+dadd+bar", (k_bar_sts_lds<<<1, 64>>>(d, c)), 64)
    RUN("sts+bar+lds+dadd+bar", (k_bar_sts_lds<<<1, 512>>>(d, c)), 512)
    RUN("lds.64 throughput 16 warps", (k_lds_tp<<<1, 512>>>(d, c)), 512)
    RUN("lds.64 throughput 32 warps", (k_lds_tp<<<1, 1024>>>(d, c)), 1024)
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
