// Pieces shared by the two self-moment kernels: shared-memory accumulators, the per-warp stash
// used to fold per-lane partial sums, the first-observation term and the term loop body.
#pragma once
#include "common.cuh"

namespace kb2 {

constexpr int kStashRow = 36;   // double2 per stash row: 32 lanes + 4 pad -> conflict-free transposed reads
constexpr int kMaxLags = 768;   // lags per launch (longer grids are split by the host entry point)
constexpr int kMaxGrid = 512;   // upper bound on persistent CTAs (partials buffer sizing)

__host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Shared-memory block of a moment kernel with `warps` consumer warps that fold `lg` lags at a time.
struct MomentSmem {
    double* acc;     // [K][2] running sums of d^2 and d^4 for this CTA
    double2* stash;  // [warps][lg][kStashRow]
    int* lags;       // [K]
    int lg;
    __device__ MomentSmem(unsigned char* raw, int K, int warps, int lg_) : lg(lg_) {
        acc = reinterpret_cast<double*>(raw);
        stash = reinterpret_cast<double2*>(raw + align16(16 * (size_t)K));
        lags = reinterpret_cast<int*>(raw + align16(16 * (size_t)K) + sizeof(double2) * warps * lg_ * kStashRow);
    }
    __host__ __device__ static size_t bytes(int K, int warps, int lg_) {
        return align16(16 * (size_t)K) + sizeof(double2) * warps * lg_ * kStashRow + align16(4 * (size_t)K);
    }
    __device__ __forceinline__ void stash_put(int warp, int g, int lane, double s1, double s2) {
        stash[(warp * lg + g) * kStashRow + lane] = make_double2(s1, s2);
    }
};

// Folds the LG x 32 per-lane partials a warp has stashed: 32/LG lanes per lag each sum LG entries
// (interleaved so a quarter-warp touches 8 consecutive 16-byte slots), then log2(32/LG) shuffle
// steps, then one shared-memory atomic per lag and moment.
template <int LG>
__device__ __forceinline__ void fold_lag_group(MomentSmem& sm, int warp, int lane, int i0, int K) {
    constexpr int LPG = 32 / LG;  // lanes per lag
    __syncwarp();
    const int g = lane / LPG, q = lane % LPG;
    const double2* rowp = sm.stash + (warp * LG + g) * kStashRow;
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int m = 0; m < LG; ++m) {
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
__device__ __forceinline__ void first_observation_terms(const double* const (&row)[3], MomentSmem& sm, int K, int tid,
                                                        int nthreads) {
    for (int i = tid; i < K; i += nthreads) {
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

// One lag of one work item for one thread: R terms, origins in registers, partners fetched by
// `load(c, li)` (li = index of the origin inside the tile).  Loads are issued RB terms at a time before
// any arithmetic so that RB*NDIM loads are in flight per thread; an origin without a partner
// (li >= nv) substitutes itself for the partner and so contributes exactly zero, which keeps the
// loop free of predicated arithmetic.  Two accumulator pairs halve the dependent chains.
template <int NDIM, int R, int RB, int NT, bool FULL, typename Load>
__device__ __forceinline__ void lag_terms(const double (&o)[3][R], int nv, int tid, Load load, double& s1, double& s2) {
    double a1 = 0.0, a2 = 0.0, b1 = 0.0, b2 = 0.0;
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += RB) {
        double p[NDIM][RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int li = (r0 + j) * NT + tid;
            const bool ok = FULL || li < nv;
#pragma unroll
            for (int c = 0; c < NDIM; ++c) p[c][j] = ok ? load(c, li) : o[c][r0 + j];
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            double d2 = 0.0;
#pragma unroll
            for (int c = 0; c < NDIM; ++c) {
                const double d = p[c][j] - o[c][r0 + j];
                d2 = fma(d, d, d2);
            }
            if (j & 1) {
                b1 += d2;
                b2 = fma(d2, d2, b2);
            } else {
                a1 += d2;
                a2 = fma(d2, d2, a2);
            }
        }
    }
    s1 = a1 + b1;
    s2 = a2 + b2;
}

}  // namespace kb2
