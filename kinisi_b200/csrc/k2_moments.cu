// Stage 2: second and fourth moments of the displacement over atoms x time origins x dt, fused with
// the per-dt displacement construction (kinisi/parser.py:208-213 + kinisi/diffusion.py:571-600).
//
// Work decomposition (both variants)
//   work item = (atom, origin tile of B = 256*R consecutive time origins)
//   each thread keeps R origins (x, y, z) in registers for the whole item and walks the lag grid in
//   ascending order; for lag k the partner of origin j is sample j + k.  Per term: 3 DADD + 3 DFMA
//   (d^2) + DADD (sum d^2) + DFMA (sum d^4) = 8 FP64 instructions and 24 B of partner data.
//   Per-lag partial sums are staged per warp in shared memory and folded, kLagGroup lags at a time,
//   with a transposed read + a few shuffles, so the cross-lane reduction costs ~5 DADD per lag per lane.
//   Persistent CTAs pull work items from an atomic counter and keep [n_lags][2] accumulators in
//   shared memory; one partial per CTA goes to HBM and a second kernel folds the partials, so the
//   result layout is deterministic.
//
// Variant 0 (this file): partners are read with read-only global loads; overlapping partner windows
//   of consecutive lags are served by L1.
// Variant 1 (k2_moments_tma.cu): partner windows are staged through a shared-memory ring filled by
//   the TMA engine (cp.async.bulk + mbarrier) by a producer warp.
//
// Algorithmic bytes: one read of the trajectory, n_atoms * n_frames * 24 B.  Algorithmic FP64 work:
// 8 instructions per term, terms = n_atoms * sum_k (n_frames - k + 1).  The kernel is bound by the
// FP64 pipe / shared-memory bandwidth, not by HBM.
#include <algorithm>

#include "common.cuh"
#include "peer.cuh"
#include "k2_common.cuh"

namespace kb2 {

constexpr int kDirectWarps = 8;
constexpr int kDirectLG = 4;

template <int NDIM, int R>
__global__ void __launch_bounds__(256, (R <= 8) ? 2 : 1)
    self_moments_direct_kernel(const double* __restrict__ dc, int n_atoms, int Tn, int64_t pitch, int c0, int c1, int c2,
                               const int32_t* __restrict__ lags, int K, int ntiles, unsigned int* work_counter,
                               double* __restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MomentSmem sm(smem_raw, K, kDirectWarps, kDirectLG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int B = 256 * R;
    constexpr int RB = 4;
    __shared__ int s_work;

    for (int i = tid; i < 2 * K; i += 256) sm.acc[i] = 0.0;
    for (int i = tid; i < K; i += 256) sm.lags[i] = lags[i];
    __syncthreads();

    const int n_work = n_atoms * ntiles;
    const int comp[3] = {c0, c1, c2};
    while (true) {
        if (tid == 0) s_work = (int)atomicAdd(work_counter, 1u);
        __syncthreads();
        const int wk = s_work;
        __syncthreads();
        if (wk >= n_work) break;
        const int atom = wk / ntiles, tile = wk - atom * ntiles;
        const int64_t j0 = (int64_t)tile * B;
        const double* row[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) row[c] = dc + ((size_t)atom * 3 + comp[c < NDIM ? c : 0]) * pitch;

        double o[3][R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t j = j0 + r * 256 + tid;
#pragma unroll
            for (int c = 0; c < NDIM; ++c) o[c][r] = (j < Tn) ? __ldg(row[c] + j) : 0.0;
        }

        if (tile == 0) first_observation_terms<NDIM>(row, sm, K, tid, 256);

        for (int i0 = 0; i0 < K; i0 += kDirectLG) {
            if ((int64_t)Tn - sm.lags[i0] - j0 <= 0) break;  // ascending lags: nothing left for this tile
#pragma unroll 1
            for (int g = 0; g < kDirectLG; ++g) {
                const int i = i0 + g;
                double s1 = 0.0, s2 = 0.0;
                if (i < K) {
                    const int k = sm.lags[i];
                    const int64_t nv64 = (int64_t)Tn - k - j0;
                    const double* p0 = row[0] + j0 + k;
                    const double* p1 = row[1] + j0 + k;
                    const double* p2 = row[2] + j0 + k;
                    auto load = [&](int c, int li) { return __ldg((c == 0 ? p0 : (c == 1 ? p1 : p2)) + li); };
                    if (nv64 >= B) {
                        lag_terms<NDIM, R, RB, 256, true>(o, B, tid, load, s1, s2);
                    } else if (nv64 > 0) {
                        lag_terms<NDIM, R, RB, 256, false>(o, (int)nv64, tid, load, s1, s2);
                    }
                }
                sm.stash_put(warp, g, lane, s1, s2);
            }
            fold_lag_group<kDirectLG>(sm, warp, lane, i0, K);
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * K; i += 256) partials[(size_t)blockIdx.x * 2 * K + i] = sm.acc[i];
}

__global__ void moments_finalize_kernel(const double* __restrict__ partials, int G, int K2, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K2) return;
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += partials[(size_t)g * K2 + i];
    out[i] = s;
}

__global__ void moment_stats_kernel(const double* __restrict__ ms, const double* __restrict__ cnt,
                                    const double* __restrict__ divisor, const double* __restrict__ gs,
                                    const double* __restrict__ gcnt, int n, double* __restrict__ out_n,
                                    double* __restrict__ out_v, double* __restrict__ out_s,
                                    double* __restrict__ out_ngp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m1 = ms[2 * i], m2 = ms[2 * i + 1], c = cnt[i];
    const double mean = m1 / c;
    const double var = (m2 - m1 * mean) / (c - 1.0);
    const double v = var / divisor[i];
    out_n[i] = mean;
    out_v[i] = v;
    out_s[i] = sqrt(v);
    const double g = gcnt[i];
    const double gm = gs[2 * i] / g;
    out_ngp[i] = (gs[2 * i + 1] / g * 3.0) / (gm * gm * 5.0) - 1.0;
}

// The same with the per-lag sums of the atoms' own displacements still local to this rank (atoms sharded over GPUs): ONE CTA
// first joins gs [n][2] over the ranks through peer memory (peer_join_slice: the protocol of kb2_peer_allreduce_f64 for an
// array of one slice, so 2 n <= 1024), in place, then finishes the statistics from the joined sums -- the join and its
// consumer are one kernel.  ms may be gs (MSD) or an array that is already the same on every rank (the collective sums).
__global__ void __launch_bounds__(256) moment_stats_joined_kernel(const double* __restrict__ ms, const double* __restrict__ cnt,
                                                                  const double* __restrict__ divisor, double* gs,
                                                                  const double* __restrict__ gcnt, int n, double* __restrict__ out_n,
                                                                  double* __restrict__ out_v, double* __restrict__ out_s,
                                                                  double* __restrict__ out_ngp, PeerJoin J) {
    peer_join_slice(J, gs, 1, 2ll * n, 0, 0, 2ll * n);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double m1 = __ldcg(ms + 2 * i), m2 = __ldcg(ms + 2 * i + 1), c = cnt[i];
        const double mean = m1 / c;
        const double var = (m2 - m1 * mean) / (c - 1.0);
        const double v = var / divisor[i];
        out_n[i] = mean;
        out_v[i] = v;
        out_s[i] = sqrt(v);
        const double g = gcnt[i];
        const double gm = __ldcg(gs + 2 * i) / g;
        out_ngp[i] = (__ldcg(gs + 2 * i + 1) / g * 3.0) / (gm * gm * 5.0) - 1.0;
    }
}

// ---- moments of one materialised displacement array (the reference seam) ----------------------
__global__ void __launch_bounds__(256) disp_self_kernel(const double* __restrict__ disp, int64_t n_elem, int m0, int m1,
                                                        int m2, double* __restrict__ out) {
    double s1 = 0.0, s2 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n_elem; e += stride) {
        const double* p = disp + e * 3;
        double d2 = 0.0;
        if (m0) d2 = fma(p[0], p[0], d2);
        if (m1) d2 = fma(p[1], p[1], d2);
        if (m2) d2 = fma(p[2], p[2], d2);
        s1 += d2;
        s2 = fma(d2, d2, s2);
    }
    __shared__ double sh[2][8];
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = s1;
        sh[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) {
            a += sh[0][w];
            b += sh[1][w];
        }
        atomicAdd(out + 0, a);
        atomicAdd(out + 1, b);
    }
}

__global__ void __launch_bounds__(256) disp_collective_kernel(const double* __restrict__ disp, int n_atoms, int n_obs,
                                                              const double* __restrict__ w, int m0, int m1, int m2,
                                                              double* __restrict__ out) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    double c = 0.0;
    if (j < n_obs) {
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int a = 0; a < n_atoms; ++a) {
            const double* p = disp + ((size_t)a * n_obs + j) * 3;
            const double wa = w ? w[a] : 1.0;
            sx += wa * p[0];
            sy += wa * p[1];
            sz += wa * p[2];
        }
        if (m0) c = fma(sx, sx, c);
        if (m1) c = fma(sy, sy, c);
        if (m2) c = fma(sz, sz, c);
    }
    double c1 = warp_sum(c), c2 = warp_sum(c * c);
    __shared__ double sh[2][8];
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = c1;
        sh[1][threadIdx.x >> 5] = c2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int wd = 0; wd < 8; ++wd) {
            a += sh[0][wd];
            b += sh[1][wd];
        }
        atomicAdd(out + 2, a);
        atomicAdd(out + 3, b);
    }
}

template <int NDIM>
static int launch_direct(int R, int grid, size_t smem, cudaStream_t st, const double* dc, int n_atoms, int Tn,
                         int64_t pitch, const int comp[3], const int32_t* lags, int K, int ntiles, unsigned int* counter,
                         double* partials) {
    if (R == 8) {
        KB2_CUDA(cudaFuncSetAttribute(self_moments_direct_kernel<NDIM, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        self_moments_direct_kernel<NDIM, 8><<<grid, 256, smem, st>>>(dc, n_atoms, Tn, pitch, comp[0], comp[1], comp[2],
                                                                      lags, K, ntiles, counter, partials);
    } else {
        KB2_CUDA(cudaFuncSetAttribute(self_moments_direct_kernel<NDIM, 16>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        self_moments_direct_kernel<NDIM, 16><<<grid, 256, smem, st>>>(dc, n_atoms, Tn, pitch, comp[0], comp[1],
                                                                       comp[2], lags, K, ntiles, counter, partials);
    }
    KB2_LAUNCH_CHECK();
    return 0;
}

int launch_self_moments_tma(int config, int ndim, int sm_count, cudaStream_t st, const double* dc, int n_atoms, int Tn,
                            int64_t pitch, const int comp[3], const int32_t* lags, int K, unsigned int* counter,
                            double* partials, int* grid_out);

// Launches the direct kernel on `K` <= kMaxLags lags; partials [grid][K][2], counter zeroed by the caller.
int launch_self_moments_direct(int R, int ndim, int sm_count, cudaStream_t st, const double* dc, int n_atoms, int Tn,
                               int64_t pitch, const int comp[3], const int32_t* lags, int K, unsigned int* counter,
                               double* partials, int* grid_out) {
    if (R != 8 && R != 16) R = 8;
    const int B = 256 * R;
    const int ntiles = ceil_div(Tn, B);
    const int64_t n_work = (int64_t)n_atoms * ntiles;
    KB2_CHECK(n_work < (1ll << 31), "kb2_self_moments: too many work items");
    int grid = (int)std::min<int64_t>(n_work, (int64_t)sm_count * (R <= 8 ? 2 : 1));
    grid = min(grid, kMaxGrid);
    *grid_out = grid;
    const size_t smem = MomentSmem::bytes(K, kDirectWarps, kDirectLG);
    if (ndim == 1) return launch_direct<1>(R, grid, smem, st, dc, n_atoms, Tn, pitch, comp, lags, K, ntiles, counter, partials);
    if (ndim == 2) return launch_direct<2>(R, grid, smem, st, dc, n_atoms, Tn, pitch, comp, lags, K, ntiles, counter, partials);
    return launch_direct<3>(R, grid, smem, st, dc, n_atoms, Tn, pitch, comp, lags, K, ntiles, counter, partials);
}


}  // namespace kb2

using namespace kb2;

extern "C" size_t kb2_self_moments_workspace_bytes(int n_lags) {
    // [kMaxGrid][min(n_lags, kMaxLags)][2] partial sums + the work counter (kept 256-byte aligned)
    const int k = n_lags < 1 ? 1 : (n_lags > kMaxLags ? kMaxLags : n_lags);
    return 256 + (size_t)kMaxGrid * 2 * (size_t)k * sizeof(double);
}

extern "C" int kb2_self_moments(const double* dc, int n_atoms, int n_frames, int64_t pitch, const int32_t* lags,
                                int n_lags, int dim_mask, int variant, void* workspace, double* sums, void* stream) {
    KB2_CHECK(dc && lags && workspace && sums, "kb2_self_moments: null pointer");
    KB2_CHECK(n_atoms > 0 && n_frames > 0 && n_lags > 0, "kb2_self_moments: empty problem");
    KB2_CHECK(pitch >= n_frames && pitch % 2 == 0, "kb2_self_moments: pitch must be even and >= n_frames");
    KB2_CHECK(dim_mask >= 1 && dim_mask <= 7, "kb2_self_moments: dim_mask must be in 1..7");
    KB2_CHECK((variant & 0xff) == 0 || (variant & 0xff) == 1, "kb2_self_moments: variant must be 0 (direct) or 1 (tma)");
    cudaStream_t st = (cudaStream_t)stream;
    int comp[3] = {0, 0, 0}, ndim = 0;
    for (int c = 0; c < 3; ++c)
        if (dim_mask & (1 << c)) comp[ndim++] = c;
    const int sm_count = device_sm_count();

    unsigned int* counter = (unsigned int*)workspace;
    double* partials = (double*)((char*)workspace + 256);

    // Long lag grids are processed kMaxLags at a time (the accumulators live in shared memory).
    for (int l0 = 0; l0 < n_lags; l0 += kMaxLags) {
        const int K = min(kMaxLags, n_lags - l0);
        const int32_t* lg = lags + l0;
        KB2_CUDA(cudaMemsetAsync(counter, 0, 256, st));
        int grid = 0;
        if ((variant & 0xff) == 0) {
            // bits 8.. of `variant` optionally force R (8 or 16); default 8 (two CTAs per SM)
            if (int rc = launch_self_moments_direct((variant >> 8) & 0xff, ndim, sm_count, st, dc, n_atoms, n_frames, pitch,
                                                    comp, lg, K, counter, partials, &grid))
                return rc;
        } else {
            // bits 8.. of `variant` select the TMA configuration (0/1: one 16-warp CTA per SM with a
            // 4096-origin tile; 2: two 8-warp CTAs per SM with 2048-origin tiles)
            int rc = launch_self_moments_tma((variant >> 8) & 0xff, ndim, sm_count, st, dc, n_atoms, n_frames, pitch,
                                             comp, lg, K, counter, partials, &grid);
            if (rc) return rc;
        }
        moments_finalize_kernel<<<ceil_div(2 * K, 128), 128, 0, st>>>(partials, grid, 2 * K, sums + 2 * l0);
        KB2_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int kb2_moment_stats(const double* main_sums, const double* cnt, const double* divisor,
                                const double* ngp_sums, const double* gcnt, int n, double* out_n, double* out_v,
                                double* out_s, double* out_ngp, void* stream) {
    KB2_CHECK(main_sums && cnt && divisor && ngp_sums && gcnt && out_n && out_v && out_s && out_ngp,
              "kb2_moment_stats: null pointer");
    KB2_CHECK(n > 0, "kb2_moment_stats: n must be positive");
    moment_stats_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(main_sums, cnt, divisor, ngp_sums, gcnt, n,
                                                                            out_n, out_v, out_s, out_ngp);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_moment_stats_joined(const double* main_sums, const double* cnt, const double* divisor, double* ngp_sums,
                                       const double* gcnt, int n, double* out_n, double* out_v, double* out_s, double* out_ngp,
                                       void* const* windows_host, int world, int rank, int64_t capacity, uint64_t seq,
                                       void* stream) {
    KB2_CHECK(main_sums && cnt && divisor && ngp_sums && gcnt && out_n && out_v && out_s && out_ngp,
              "kb2_moment_stats_joined: null pointer");
    KB2_CHECK(n > 0 && 2 * n <= 1024, "kb2_moment_stats_joined: 1 <= n <= 512 (longer grids: kb2_peer_allreduce_f64 + kb2_moment_stats)");
    KB2_CHECK(2ll * n <= capacity, "kb2_moment_stats_joined: 2 n exceeds the capacity of the windows");
    KB2_CHECK((reinterpret_cast<uintptr_t>(ngp_sums) & 15) == 0, "kb2_moment_stats_joined: ngp_sums must be 16-byte aligned");
    PeerJoin J;
    if (int rc = peer_join_setup(&J, windows_host, world, rank, capacity, seq, 1, "kb2_moment_stats_joined")) return rc;
    moment_stats_joined_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(main_sums, cnt, divisor, ngp_sums, gcnt, n, out_n, out_v, out_s,
                                                                    out_ngp, J);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_disp_moments(const double* disp, int n_atoms, int n_obs, int dim_mask, const double* w,
                                int coll_mask, double* out, void* stream) {
    KB2_CHECK(disp && out, "kb2_disp_moments: null pointer");
    KB2_CHECK(n_atoms > 0 && n_obs > 0, "kb2_disp_moments: empty array");
    KB2_CHECK(dim_mask >= 1 && dim_mask <= 7 && coll_mask >= 0 && coll_mask <= 7, "kb2_disp_moments: bad mask");
    cudaStream_t st = (cudaStream_t)stream;
    KB2_CUDA(cudaMemsetAsync(out, 0, 4 * sizeof(double), st));
    const int sm_count = device_sm_count();
    const int64_t n_elem = (int64_t)n_atoms * n_obs;
    const int grid = (int)std::min<int64_t>(ceil_div(n_elem, 256), (int64_t)sm_count * 8);
    disp_self_kernel<<<grid, 256, 0, st>>>(disp, n_elem, dim_mask & 1, dim_mask & 2, dim_mask & 4, out);
    KB2_LAUNCH_CHECK();
    if (coll_mask) {
        disp_collective_kernel<<<ceil_div(n_obs, 256), 256, 0, st>>>(disp, n_atoms, n_obs, w, coll_mask & 1,
                                                                     coll_mask & 2, coll_mask & 4, out);
        KB2_LAUNCH_CHECK();
    }
    return 0;
}
