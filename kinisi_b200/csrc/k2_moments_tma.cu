// Stage 2, variant 1: the self-moment kernel with partner samples staged through a shared-memory
// ring that the TMA engine fills (cp.async.bulk + mbarrier complete_tx), one producer warp feeding
// the consumer warps.  See k2_moments.cu for the work decomposition and the arithmetic.
//
// Shared memory per CTA:
//   ring   [NDIM][CAP] f64        CAP = 8192 (192 KB at NDIM=3, one 16-warp CTA per SM) or 4096 (two 8-warp
//                                 CTAs per SM); frame p of the current atom lives at (p + off) & (CAP-1)
//   full / empty mbarriers        CAP/512 each
//   accumulators [K][2], lag table [K], per-warp stash
// The trajectory row of one component is cut into absolute 512-frame chunks (4 KB); chunk q of a
// work item is the (seq)-th chunk this CTA ever loads and goes to slot seq % 16.  A consumer warp
// waits on `full` for the chunks a lag's partner window [j0+k, j0+k+B) touches, and arrives on
// `empty` for every chunk that lies wholly below the next lag's window.  The producer re-arms a
// slot once all consumer warps have released its previous occupant.  Consecutive lags' windows
// overlap by B - dk frames, so each frame is copied from L2 once per origin tile, not once per lag:
// L2->SM traffic is 24 B * min(dk, B) / B per term instead of 24 B per term.
#include <algorithm>

#include "common.cuh"
#include "k2_common.cuh"

namespace kb2 {

namespace tma {
constexpr int CH = 512;  // frames per bulk copy (4 KB per component)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "KB2_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KB2_WAIT_DONE;\n\t"
        "bra KB2_WAIT_LOOP;\n\t"
        "KB2_WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr size_t kBarrierBytes = 512;
__host__ __device__ constexpr size_t ring_bytes(int ndim, int cap) { return (size_t)ndim * cap * sizeof(double); }
}  // namespace tma

// WARPS consumer warps + one producer warp; R origins per consumer thread (tile B = 32*WARPS*R);
// ring of CAP frames per component; LG lags folded at a time.
template <int NDIM, int WARPS, int R, int CAP, int LG, int MINB>
__global__ void __launch_bounds__(32 * WARPS + 32, MINB)
    self_moments_tma_kernel(const double* __restrict__ dc, int n_atoms, int Tn, int64_t pitch, int c0, int c1, int c2,
                            const int32_t* __restrict__ lags, int K, int ntiles, unsigned int* work_counter,
                            double* __restrict__ partials) {
    using namespace tma;
    constexpr int NT = 32 * WARPS;       // consumer threads
    constexpr int THREADS = NT + 32;
    constexpr int B = NT * R;            // origins per work item
    constexpr int NSLOT = CAP / CH;
    constexpr int RB = 4;
    static_assert(CAP >= B + 4 * CH, "ring too small for the partner window plus prefetch slack");
    static_assert((CAP & (CAP - 1)) == 0 && 2 * NSLOT * 8 <= (int)kBarrierBytes, "bad ring geometry");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + ring_bytes(NDIM, CAP));
    uint64_t* empty = full + NSLOT;
    MomentSmem sm(smem_raw + ring_bytes(NDIM, CAP) + kBarrierBytes, K, WARPS, LG);
    __shared__ int s_work;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == WARPS;

    for (int i = tid; i < 2 * K; i += THREADS) sm.acc[i] = 0.0;
    for (int i = tid; i < K; i += THREADS) sm.lags[i] = lags[i];
    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int n_work = n_atoms * ntiles;
    const int comp[3] = {c0, c1, c2};
    uint32_t seq_base = 0;  // chunks loaded by this CTA so far (identical in every thread)

    while (true) {
        if (tid == 0) s_work = (int)atomicAdd(work_counter, 1u);
        __syncthreads();
        const int wk = s_work;
        __syncthreads();
        if (wk >= n_work) break;
        const int atom = wk / ntiles, tile = wk - atom * ntiles;
        const int j0 = tile * B;
        const double* row[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) row[c] = dc + ((size_t)atom * 3 + comp[c < NDIM ? c : 0]) * pitch;

        // number of lags that still have origins in this tile (ascending grid -> a prefix)
        int nlv = 0;
        {
            int lo = 0, hi = K;
            const int limit = Tn - j0;  // lag k contributes iff k < limit
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sm.lags[mid] < limit) lo = mid + 1; else hi = mid;
            }
            nlv = lo;
        }
        int q_first = 0, nq = 0;
        if (nlv > 0) {
            q_first = (j0 + sm.lags[0]) / CH;
            const int e_last = min(Tn, j0 + sm.lags[nlv - 1] + B);
            nq = (e_last - 1) / CH - q_first + 1;
        }

        if (producer) {
            if (lane == 0) {
                for (int qi = 0; qi < nq; ++qi) {
                    const uint32_t s = seq_base + qi;
                    const uint32_t slot = s % NSLOT, use = s / NSLOT;
                    if (use >= 1) mbar_wait(&empty[slot], (use - 1) & 1);
                    const int64_t f0 = (int64_t)(q_first + qi) * CH;
                    const int64_t left = pitch - f0;
                    const uint32_t bytes = (uint32_t)(left < CH ? left : CH) * (uint32_t)sizeof(double);
                    mbar_arrive_expect_tx(&full[slot], NDIM * bytes);
#pragma unroll
                    for (int c = 0; c < NDIM; ++c) bulk_g2s(ring + c * CAP + slot * CH, row[c] + f0, bytes, &full[slot]);
                }
            }
            __syncwarp();
        } else {
            double o[3][R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = j0 + r * NT + tid;
#pragma unroll
                for (int c = 0; c < NDIM; ++c) o[c][r] = (j < Tn) ? __ldg(row[c] + j) : 0.0;
            }
            if (tile == 0) first_observation_terms<NDIM>(row, sm, K, tid, NT);

            // ring position of frame p is (p + off) & (CAP-1)
            const int off = (int)(((seq_base % NSLOT) + NSLOT - (uint32_t)(q_first % NSLOT)) % NSLOT) * CH;
            int waited = q_first - 1;    // last chunk this warp has seen complete
            int released = q_first - 1;  // last chunk this warp has handed back

            for (int i0 = 0; i0 < nlv; i0 += LG) {
#pragma unroll 1
                for (int g = 0; g < LG; ++g) {
                    const int i = i0 + g;
                    double s1 = 0.0, s2 = 0.0;
                    if (i < nlv) {
                        const int start = j0 + sm.lags[i];
                        const int nv = min(B, Tn - start);
                        const int q_hi = (start + nv - 1) / CH;
                        while (waited < q_hi) {
                            ++waited;
                            const uint32_t s = seq_base + (uint32_t)(waited - q_first);
                            mbar_wait(&full[s % NSLOT], (s / NSLOT) & 1);
                        }
                        const int pbase = start + off;
                        auto load = [&](int c, int li) { return ring[c * CAP + ((pbase + li) & (CAP - 1))]; };
                        if (nv == B)
                            lag_terms<NDIM, R, RB, NT, true>(o, B, tid, load, s1, s2);
                        else
                            lag_terms<NDIM, R, RB, NT, false>(o, nv, tid, load, s1, s2);
                        // hand back the chunks that lie wholly below the next lag's window
                        // (a chunk in a gap between two windows is waited for before it is released, so an
                        //  `empty` arrival can never land in the slot's previous phase)
                        const int next_lo = (i + 1 < nlv) ? (j0 + sm.lags[i + 1]) / CH : q_first + nq;
                        __syncwarp();
                        while (released < next_lo - 1) {
                            ++released;
                            if (released > waited) {
                                waited = released;
                                const uint32_t sw = seq_base + (uint32_t)(waited - q_first);
                                mbar_wait(&full[sw % NSLOT], (sw / NSLOT) & 1);
                            }
                            if (lane == 0) {
                                const uint32_t s = seq_base + (uint32_t)(released - q_first);
                                mbar_arrive(&empty[s % NSLOT]);
                            }
                        }
                    }
                    sm.stash_put(warp, g, lane, s1, s2);
                }
                fold_lag_group<LG>(sm, warp, lane, i0, K);
            }
        }
        seq_base += (uint32_t)nq;
    }
    __syncthreads();
    for (int i = tid; i < 2 * K; i += THREADS) partials[(size_t)blockIdx.x * 2 * K + i] = sm.acc[i];
}

template <int NDIM, int WARPS, int R, int CAP, int LG, int MINB>
static int launch_tma_cfg(int sm_count, cudaStream_t st, const double* dc, int n_atoms, int Tn, int64_t pitch,
                          const int comp[3], const int32_t* lags, int K, unsigned int* counter, double* partials,
                          int* grid_out) {
    constexpr int B = 32 * WARPS * R;
    const int ntiles = ceil_div(Tn, B);
    const int64_t n_work = (int64_t)n_atoms * ntiles;
    KB2_CHECK(n_work < (1ll << 31), "kb2_self_moments: too many work items");
    const size_t smem = tma::ring_bytes(NDIM, CAP) + tma::kBarrierBytes + MomentSmem::bytes(K, WARPS, LG);
    auto kern = self_moments_tma_kernel<NDIM, WARPS, R, CAP, LG, MINB>;
    KB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    KB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * WARPS + 32, smem));
    KB2_CHECK(per_sm >= 1, "kb2_self_moments: the TMA kernel does not fit on this device (%zu bytes of shared memory)",
              smem);
    const int grid = (int)std::min<int64_t>(std::min<int64_t>((int64_t)sm_count * per_sm, kMaxGrid), n_work);
    kern<<<grid, 32 * WARPS + 32, smem, st>>>(dc, n_atoms, Tn, pitch, comp[0], comp[1], comp[2], lags, K, ntiles, counter,
                                              partials);
    KB2_LAUNCH_CHECK();
    *grid_out = grid;
    return 0;
}

template <int NDIM>
static int launch_tma_ndim(int config, int sm_count, cudaStream_t st, const double* dc, int n_atoms, int Tn,
                           int64_t pitch, const int comp[3], const int32_t* lags, int K, unsigned int* counter,
                           double* partials, int* grid_out) {
    if (config == 2)  // two 8-warp CTAs per SM, 2048-origin tiles, 4096-frame ring
        return launch_tma_cfg<NDIM, 8, 8, 4096, 2, 2>(sm_count, st, dc, n_atoms, Tn, pitch, comp, lags, K, counter,
                                                      partials, grid_out);
    // one 16-warp CTA per SM, 4096-origin tiles, 8192-frame ring
    return launch_tma_cfg<NDIM, 16, 8, 8192, 2, 1>(sm_count, st, dc, n_atoms, Tn, pitch, comp, lags, K, counter, partials,
                                                   grid_out);
}

int launch_self_moments_tma(int config, int ndim, int sm_count, cudaStream_t st, const double* dc, int n_atoms, int Tn,
                            int64_t pitch, const int comp[3], const int32_t* lags, int K, unsigned int* counter,
                            double* partials, int* grid_out) {
    KB2_CHECK(((uintptr_t)dc & 15) == 0, "kb2_self_moments: dc must be 16-byte aligned for the TMA variant");
    KB2_CHECK((int64_t)Tn + 8192 < (1ll << 30), "kb2_self_moments: trajectory too long for the TMA variant");
    if (ndim == 1)
        return launch_tma_ndim<1>(config, sm_count, st, dc, n_atoms, Tn, pitch, comp, lags, K, counter, partials, grid_out);
    if (ndim == 2)
        return launch_tma_ndim<2>(config, sm_count, st, dc, n_atoms, Tn, pitch, comp, lags, K, counter, partials, grid_out);
    return launch_tma_ndim<3>(config, sm_count, st, dc, n_atoms, Tn, pitch, comp, lags, K, counter, partials, grid_out);
}

}  // namespace kb2
