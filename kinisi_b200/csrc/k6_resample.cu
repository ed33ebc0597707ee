// `bootstrap=True` (kinisi/diffusion.py:258-289, 578-583, 663-669, 755-761, 783-801):
// the distribution of the mean of n_o observations drawn WITH replacement from the squared displacements of one lag.
//
// The reference draws int(n_o) indices per resample from numpy's Mersenne twister and gathers on the host, once per
// resample, 1000+ times per lag.  Here one launch makes all resamples of a lag: the population (the squared
// displacements of the lag, written once by kb2_lag_squares / kb2_disp_squares: at most N * T doubles, L2-resident for
// the sizes this option is used at) is gathered by one CTA per resample with indices from a counter-based generator,
// so draw d of resample r is a pure function of (seed, stream, r, d) -- the result does not depend on the launch shape
// and a later batch of resamples (sample_until_normal's "+100 until normal") continues the same sequence.
//   index(d) = floor(u64(d) * M / 2^64),  u64 = one half of Philox4x32-10(counter = {d/2 lo, d/2 hi, r, stream}, key = seed)
// (multiply-high range reduction: the bias is M / 2^64 < 2^-30 for any population that fits the device).
// The per-thread partial sums are reduced in a fixed order: the same seed gives the same bits.
#include <algorithm>

#include "common.cuh"

namespace kb2 {

constexpr int kRsThreads = 256;

// out[a * n_obs + m] = |dc[a][:, k - 1 + m] - (m > 0 ? dc[a][:, m - 1] : 0)|^2 over the dim_mask components, n_obs = T - k + 1:
// entry 0 is the reference's first-observation term dc[:, k - 1] (kinisi/parser.py:205-207).
__global__ void __launch_bounds__(256) lag_squares_kernel(const double* __restrict__ dc, int n_frames, int64_t pitch,
                                                          int lag, int dim_mask, double* __restrict__ out) {
    const int n_obs = n_frames - lag + 1;
    const int m = blockIdx.x * 256 + threadIdx.x;
    if (m >= n_obs) return;
    const double* row = dc + (size_t)blockIdx.y * 3 * pitch;
    double d2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (!((dim_mask >> c) & 1)) continue;
        const double hi = row[c * pitch + lag - 1 + m];
        const double lo = m > 0 ? row[c * pitch + m - 1] : 0.0;
        const double d = hi - lo;
        d2 = fma(d, d, d2);
    }
    out[(size_t)blockIdx.y * n_obs + m] = d2;
}

// disp [n_atoms][n_obs][3] (an entry of the reference's disp_3d list).  collective == 0: out[a * n_obs + j] = |disp[a][j]|^2
// over dim_mask; collective != 0: out[j] = |sum_a w_a disp[a][j]|^2 over dim_mask (w == NULL: unit weights).
__global__ void __launch_bounds__(256) disp_squares_kernel(const double* __restrict__ disp, int n_atoms, int n_obs,
                                                           int dim_mask, const double* __restrict__ w, int collective,
                                                           double* __restrict__ out) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_obs) return;
    if (!collective) {
        const double* p = disp + ((size_t)blockIdx.y * n_obs + j) * 3;
        double d2 = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if ((dim_mask >> c) & 1) d2 = fma(p[c], p[c], d2);
        out[(size_t)blockIdx.y * n_obs + j] = d2;
        return;
    }
    double s[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < n_atoms; ++a) {
        const double* p = disp + ((size_t)a * n_obs + j) * 3;
        const double wa = w ? w[a] : 1.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) s[c] = fma(wa, p[c], s[c]);
    }
    double c2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        if ((dim_mask >> c) & 1) c2 = fma(s[c], s[c], c2);
    out[j] = c2;
}

__device__ __forceinline__ uint64_t pick(uint32_t lo, uint32_t hi, uint64_t M) {
    return __umul64hi(((uint64_t)hi << 32) | lo, M);
}

// one CTA per resample: thread t takes the draw pairs t, t + 256, ...; four pairs (eight gathers) in flight per thread
__global__ void __launch_bounds__(kRsThreads) resample_means_kernel(const double* __restrict__ values, uint64_t M,
                                                                    int64_t n_draws, uint32_t first_resample,
                                                                    uint32_t stream_id, uint2 key,
                                                                    double* __restrict__ means) {
    __shared__ double s_part[kRsThreads / 32];
    const uint32_t r = first_resample + blockIdx.x;
    const int64_t n_pairs = n_draws >> 1;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int64_t p = threadIdx.x;
    for (; p + 3 * kRsThreads < n_pairs; p += 4 * kRsThreads) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t q = (uint64_t)(p + u * kRsThreads);
            const uint4 x = Philox::run(make_uint4((uint32_t)q, (uint32_t)(q >> 32), r, stream_id), key);
            a[u] = __ldg(values + pick(x.x, x.y, M));
            b[u] = __ldg(values + pick(x.z, x.w, M));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += a[u] + b[u];
    }
    for (; p < n_pairs; p += kRsThreads) {
        const uint4 x = Philox::run(make_uint4((uint32_t)p, (uint32_t)((uint64_t)p >> 32), r, stream_id), key);
        acc[0] += __ldg(values + pick(x.x, x.y, M)) + __ldg(values + pick(x.z, x.w, M));
    }
    if ((n_draws & 1) && threadIdx.x == 0) {  // the odd last draw: first half of pair n_pairs
        const uint4 x = Philox::run(make_uint4((uint32_t)n_pairs, (uint32_t)((uint64_t)n_pairs >> 32), r, stream_id), key);
        acc[1] += __ldg(values + pick(x.x, x.y, M));
    }
    double v = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kRsThreads / 32; ++i) t += s_part[i];
        means[blockIdx.x] = t / (double)n_draws;
    }
}

// ---- block=True: Flyvbjerg-Petersen reblocking (kinisi/diffusion.py:584-595 via pyblock.blocking.reblock) ----------------
// One launch per level: with the level's sum known (acc[0], finished by the previous launch), accumulate the sum of
// squared deviations from the level's mean into acc[1] (the two-pass variance numpy.cov computes), write the next level
// y[i] = (x[2i] + x[2i+1]) / 2 -- the last point of an odd level takes part in its statistics and is then dropped -- and
// accumulate the next level's sum into acc[2].
__global__ void __launch_bounds__(256) reblock_level_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ acc,
                                                            double* __restrict__ y) {
    __shared__ double s_part[2][8];
    const double mean = acc[0] / (double)n;
    const int64_t n_pairs = n >> 1;
    double ssd = 0.0, sy = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * 256) {
        const double2 v = *reinterpret_cast<const double2*>(x + 2 * i);
        const double da = v.x - mean, db = v.y - mean;
        ssd = fma(da, da, ssd);
        ssd = fma(db, db, ssd);
        const double m = 0.5 * (v.x + v.y);
        y[i] = m;
        sy += m;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double d = x[n - 1] - mean;
        ssd = fma(d, d, ssd);
    }
    ssd = warp_sum(ssd);
    sy = warp_sum(sy);
    if ((threadIdx.x & 31) == 0) {
        s_part[0][threadIdx.x >> 5] = ssd;
        s_part[1][threadIdx.x >> 5] = sy;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_part[threadIdx.x][i];
        atomicAdd(acc + 1 + threadIdx.x, t);
    }
}

__global__ void __launch_bounds__(256) reblock_sum_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ acc) {
    __shared__ double s_part[8];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s += x[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_part[i];
        atomicAdd(acc, t);
    }
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_reblock_levels(int64_t n_values) {
    int levels = 0;
    for (int64_t n = n_values; n >= 2; n >>= 1) ++levels;
    return levels;
}

extern "C" size_t kb2_reblock_workspace_bytes(int64_t n_values) {
    size_t n = 0;
    for (int64_t m = n_values; m >= 2; m >>= 1) n += (size_t)(((m >> 1) + 1) & ~(int64_t)1);
    return sizeof(double) * (n + 2);
}

extern "C" int kb2_reblock(const double* values, int64_t n_values, double* workspace, double* stats, void* stream) {
    KB2_CHECK(n_values >= 2, "kb2_reblock: at least two values are needed");
    KB2_CHECK(values && workspace && stats, "kb2_reblock: null pointer");
    KB2_CHECK(((uintptr_t)values & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "kb2_reblock: values and workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int levels = kb2_reblock_levels(n_values);
    KB2_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (levels + 1), st));
    auto grid_for = [](int64_t n) { return (int)std::min<int64_t>(std::max<int64_t>((n + 255) / 256, 1), 148 * 8); };
    reblock_sum_kernel<<<grid_for(n_values), 256, 0, st>>>(values, n_values, stats);
    KB2_LAUNCH_CHECK();
    const double* x = values;
    double* y = workspace;
    int64_t n = n_values;
    for (int l = 0; l < levels; ++l) {
        // stats[2l] = sum of level l, stats[2l + 1] = its sum of squared deviations, stats[2l + 2] = sum of level l + 1
        reblock_level_kernel<<<grid_for(n >> 1), 256, 0, st>>>(x, n, stats + 2 * l, y);
        KB2_LAUNCH_CHECK();
        x = y;
        y += ((n >> 1) + 1) & ~(int64_t)1;  // keep every level 16-byte aligned
        n >>= 1;
    }
    return 0;
}

extern "C" int kb2_lag_squares(const double* dc, int n_atoms, int n_frames, int64_t pitch, int lag, int dim_mask,
                               double* out, void* stream) {
    KB2_CHECK(n_atoms >= 0 && n_frames >= 1 && pitch >= n_frames, "kb2_lag_squares: bad extents");
    KB2_CHECK(lag >= 1 && lag <= n_frames, "kb2_lag_squares: lag %d out of range [1, %d]", lag, n_frames);
    KB2_CHECK(dim_mask >= 1 && dim_mask <= 7, "kb2_lag_squares: dim_mask must be in [1, 7]");
    if (n_atoms == 0) return 0;
    KB2_CHECK(dc && out, "kb2_lag_squares: null pointer");
    KB2_CHECK(n_atoms <= 65535 * 1024, "kb2_lag_squares: too many atoms for one launch");
    const int n_obs = n_frames - lag + 1;
    for (int a0 = 0; a0 < n_atoms; a0 += 65535) {
        const int na = min(65535, n_atoms - a0);
        dim3 grid(ceil_div(n_obs, 256), na);
        lag_squares_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dc + (size_t)a0 * 3 * pitch, n_frames, pitch, lag, dim_mask,
                                                                   out + (size_t)a0 * n_obs);
        KB2_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int kb2_disp_squares(const double* disp, int n_atoms, int n_obs, int dim_mask, const double* w, int collective,
                                double* out, void* stream) {
    KB2_CHECK(n_atoms >= 0 && n_obs >= 0, "kb2_disp_squares: bad extents");
    KB2_CHECK(dim_mask >= 1 && dim_mask <= 7, "kb2_disp_squares: dim_mask must be in [1, 7]");
    if (n_obs == 0 || (n_atoms == 0 && !collective)) return 0;
    KB2_CHECK(out && (disp || n_atoms == 0), "kb2_disp_squares: null pointer");
    if (collective) {
        disp_squares_kernel<<<dim3(ceil_div(n_obs, 256), 1), 256, 0, (cudaStream_t)stream>>>(disp, n_atoms, n_obs, dim_mask, w, 1,
                                                                                            out);
        KB2_LAUNCH_CHECK();
        return 0;
    }
    for (int a0 = 0; a0 < n_atoms; a0 += 65535) {
        const int na = min(65535, n_atoms - a0);
        disp_squares_kernel<<<dim3(ceil_div(n_obs, 256), na), 256, 0, (cudaStream_t)stream>>>(
            disp + (size_t)a0 * n_obs * 3, na, n_obs, dim_mask, nullptr, 0, out + (size_t)a0 * n_obs);
        KB2_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int kb2_resample_means(const double* values, int64_t n_values, int64_t n_draws, int n_resamples,
                                  int first_resample, uint64_t seed, int stream_id, double* means, void* stream) {
    KB2_CHECK(n_values >= 1, "kb2_resample_means: empty population");
    KB2_CHECK(n_draws >= 1, "kb2_resample_means: n_draws must be at least 1");
    KB2_CHECK(n_resamples >= 0 && first_resample >= 0, "kb2_resample_means: bad resample range");
    if (n_resamples == 0) return 0;
    KB2_CHECK(values && means, "kb2_resample_means: null pointer");
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    resample_means_kernel<<<n_resamples, kRsThreads, 0, (cudaStream_t)stream>>>(
        values, (uint64_t)n_values, n_draws, (uint32_t)first_resample, (uint32_t)stream_id, key, means);
    KB2_LAUNCH_CHECK();
    return 0;
}
