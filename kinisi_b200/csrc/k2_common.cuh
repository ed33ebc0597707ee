// Pieces shared by the two self-moment kernels: shared-memory accumulators, the per-warp stash
// used to fold per-lane partial sums, and the first-observation term.
#pragma once
#include "common.cuh"

namespace kb2 {

constexpr int kLagGroup = 4;    // lags folded per transposed reduction
constexpr int kStashRow = 36;   // double2 per stash row: 32 lanes + 4 pad -> conflict-free transposed reads
constexpr int kConsumerWarps = 8;
constexpr int kMaxLags = 768;   // lags per launch (longer grids are split by the host entry point)
constexpr int kMaxGrid = 512;   // upper bound on persistent CTAs (partials buffer sizing)

__host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

struct MomentSmem {
    double* acc;     // [K][2] running sums of d^2 and d^4 for this CTA
    double2* stash;  // [warps][kLagGroup][kStashRow]
    int* lags;       // [K]
    __device__ MomentSmem(unsigned char* raw, int K) {
        acc = reinterpret_cast<double*>(raw);
        stash = reinterpret_cast<double2*>(raw + align16(16 * (size_t)K));
        lags = reinterpret_cast<int*>(raw + align16(16 * (size_t)K) +
                                      sizeof(double2) * kConsumerWarps * kLagGroup * kStashRow);
    }
    __host__ __device__ static size_t bytes(int K) {
        return align16(16 * (size_t)K) + sizeof(double2) * kConsumerWarps * kLagGroup * kStashRow +
               align16(4 * (size_t)K);
    }
    __device__ __forceinline__ void stash_put(int warp, int g, int lane, double s1, double s2) {
        stash[(warp * kLagGroup + g) * kStashRow + lane] = make_double2(s1, s2);
    }
};

// Folds the kLagGroup x 32 per-lane partials a warp has stashed: 32/kLagGroup lanes per lag each sum
// kLagGroup entries (interleaved so a quarter-warp touches 8 consecutive 16-byte slots), then
// log2(32/kLagGroup) shuffle steps, then one shared-memory atomic per lag and moment.
__device__ __forceinline__ void fold_lag_group(MomentSmem& sm, int warp, int lane, int i0, int K) {
    constexpr int LPG = 32 / kLagGroup;  // lanes per lag
    __syncwarp();
    const int g = lane / LPG, q = lane % LPG;
    const double2* rowp = sm.stash + (warp * kLagGroup + g) * kStashRow;
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int m = 0; m < kLagGroup; ++m) {
        const double2 v = rowp[q + LPG * m];
        t1 += v.x;
        t2 += v.y;
    }
#pragma unroll
    for (int mask = 1; mask < LPG; mask <<= 1) {
        t1 += shfl_xor_f64(t1, mask);
        t2 += shfl_xor_f64(t2, mask);
    }
    if (q == 0 && i0 + g < K) {
        atomicAdd(&sm.acc[2 * (i0 + g)], t1);
        atomicAdd(&sm.acc[2 * (i0 + g) + 1], t2);
    }
    __syncwarp();
}

// The extra first observation of every lag, dc[a][k-1] (kinisi/parser.py:209): one term per atom
// and lag, added by the CTA that owns tile 0 of the atom.
template <int NDIM>
__device__ __forceinline__ void first_observation_terms(const double* const (&row)[3], MomentSmem& sm, int K, int tid) {
    for (int i = tid; i < K; i += 256) {
        const int p = sm.lags[i] - 1;
        double d2 = 0.0;
#pragma unroll
        for (int c = 0; c < NDIM; ++c) {
            const double d = __ldg(row[c] + p);
            d2 = fma(d, d, d2);
        }
        atomicAdd(&sm.acc[2 * i], d2);
        atomicAdd(&sm.acc[2 * i + 1], d2 * d2);
    }
}

}  // namespace kb2
