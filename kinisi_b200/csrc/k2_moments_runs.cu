// Stage 2, variant 2: the self-moment kernel for lag grids with arithmetic structure.
//
// The direct kernel (k2_moments.cu) fetches one partner sample (24 B) per term, which pins it to the L1 / shared
// memory bandwidth of the SM (128 B/clk against 8 FP64 instructions per term at 64 lanes/clk: at most 2/3 of the
// FP64 pipe, ~0.34 measured).  kinisi's default lag grid is np.linspace(min_dt, max_dt, n_steps, dtype=int)
// (kinisi/parser.py:150-168): an arithmetic progression up to integer truncation.  For a run of lags
// k_i = k0 + i*D (i < L) and origins spaced by the same D,
//      partner(origin j + m*D, lag k0 + i*D) = x[j + k0 + (m + i)*D]
// depends on m + i only, so a thread that keeps M origins {j + m*D} in registers and walks s = m + i upwards needs
// ONE new partner sample per step for up to M terms: L1 traffic drops from 24 B to 24/M' B per term
// (M' = M*L/(M+L-1) = 7.5 for M = 8, L = 100) and the kernel becomes FP64-pipe bound.
//
// Per thread (one "column" = M origins at stride D) and step s:
//      for m in 0..M-1:  lag i = s - m;  acc[i mod M] += |P_s - o_m|^2, acc2[i mod M] += (...)^2
// Lag i receives its last contribution at step i + M - 1 and is then flushed to the per-warp stash and folded
// across the warp four lags at a time (k2_common.cuh).  A term is valid iff its partner index is inside the
// trajectory, which depends on s only, so interior blocks of M steps run fully unrolled with static register
// indices and no predication.  The triangular start, the end of the run and the end of the trajectory go through
// one generic step that keeps the same register layout by rotating the accumulators.
//
// Partner samples are staged through a per-warp shared-memory ring with cp.async (LDGSTS), each lane copying
// the 8 bytes it will read itself kRunPrefetch steps later: deep prefetch (L2 latency is ~300 cycles, HBM more)
// without spending registers, and no cross-lane synchronisation.
//
// The host splits the lag grid into arithmetic runs (greedy cover; a lag that fits no run of >= 3 becomes a run of
// one).  Work items are (run, 32-column strip, atom); every warp of the persistent CTAs pulls its own items from
// an atomic counter (no block-level barrier inside the loop).  Atoms are taken in batches whose trajectories fit
// in L2 together (each sample is a partner in ~(M+L)/M strips, so a batch is read from HBM once and served from
// L2 afterwards); inside a batch the strips are issued in order of decreasing cost.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "k2_common.cuh"

namespace kb2 {

constexpr int kRunM = 8;                   // origins per thread = ring slots
constexpr int kRunLG = 4;                  // lags folded at a time
constexpr int kRunCols = 32;               // columns per work item: one warp-wide strip
constexpr int kRunMinLen = 3;              // shortest progression taken as a run
constexpr int kRunSingleStride = 32;       // origin stride of a run of one lag: a strip covers 256 contiguous origins
constexpr int kRunCandidates = 24;         // successors tried as the common difference of a run
constexpr int kRunPrefetch = 6;            // steps between the cp.async of a partner sample and its use (< kRunM)
constexpr long long kRunBatchBytes = 24ll << 20;  // trajectory bytes of one atom batch (L2 is 126 MB)
constexpr int kRunRingBytes = kRunM * 3 * 32 * 8;  // per warp
constexpr int kRunMaxWarps = 8;

struct RunDesc {
    int k0, delta, len, off;  // lags k0 + i*delta, i < len; accumulator index of lag i is off + i
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// As fold_lag_group, for a group whose first `cnt` stash rows are live; lag g of the group goes to acc[i0 + g].
template <int LG>
__device__ __forceinline__ void fold_lag_group_n(MomentSmem& sm, int warp, int lane, int i0, int cnt) {
    constexpr int LPG = 32 / LG;
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
    if (q == 0 && g < cnt) {
        atomicAdd(&sm.acc[2 * (i0 + g)], t1);
        atomicAdd(&sm.acc[2 * (i0 + g) + 1], t2);
    }
    __syncwarp();
}

// State of one column while it walks its steps.
template <int NDIM, int M>
struct Column {
    double o[3][M];        // origins
    double A1[M], A2[M];   // running sums; at a step s = 0 (mod M) lag i lives in slot i mod M
    const double* pn[3];   // address of the partner sample of the next step to prefetch
    double* ring;          // this lane's column of the warp's ring: sample (slot, c) at ring[(slot*3 + c)*32]
    int nvalid;            // steps whose partner lies inside the trajectory (this lane)
    int fetched;           // steps prefetched so far

    __device__ __forceinline__ void prefetch(int D) {
        if (fetched < nvalid) {
#pragma unroll
            for (int c = 0; c < NDIM; ++c) cp_async8(ring + ((fetched % M) * 3 + c) * 32, pn[c]);
        }
#pragma unroll
        for (int c = 0; c < NDIM; ++c) pn[c] += D;
        ++fetched;
        cp_async_commit();
    }
};

// M consecutive steps s0 .. s0+M-1 (s0 a multiple of M) whose partners lie inside the trajectory for all lanes
// (s0 + M <= min over lanes of nvalid) and whose (step, origin) pairs are lags of the run in a static pattern:
//   FIRST = false: all of them (s0 >= M, s0 + M <= L);
//   FIRST = true:  s0 = 0 and L >= M, the triangular start: origin m joins at step m, only lag 0 completes.
template <int NDIM, int M, int LG, bool FIRST>
__device__ __forceinline__ void run_block_interior(Column<NDIM, M>& col, int D, int s0, MomentSmem& sm, int warp,
                                                   int lane, int acc_base) {
    static_assert(M % LG == 0, "the fold group must divide the register window");
#pragma unroll
    for (int u = 0; u < M; ++u) {
        cp_async_wait<kRunPrefetch - 1>();
        double P[3];
#pragma unroll
        for (int c = 0; c < NDIM; ++c) P[c] = col.ring[(u * 3 + c) * 32];
        col.prefetch(D);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            if (FIRST && m > u) continue;
            double d2 = 0.0;
#pragma unroll
            for (int c = 0; c < NDIM; ++c) {
                const double d = P[c] - col.o[c][m];
                d2 = fma(d, d, d2);
            }
            col.A1[(u - m + M) % M] += d2;
            col.A2[(u - m + M) % M] = fma(d2, d2, col.A2[(u - m + M) % M]);
        }
        // lag s - (M-1) is complete: register slot (u+1) % M, stash row (lag % LG) = (u+1) % LG
        if (!FIRST || u == M - 1) {
            const int sl = (u + 1) % M, gs = (u + 1) % LG;
            sm.stash_put(warp, gs, lane, col.A1[sl], col.A2[sl]);
            col.A1[sl] = 0.0;
            col.A2[sl] = 0.0;
            if (gs == LG - 1) fold_lag_group_n<LG>(sm, warp, lane, acc_base + s0 + u - (M - 1) - (LG - 1), LG);
        }
    }
}

// One step at any position: origin m pairs with slot (M - m) % M (the layout of a step s = 0 mod M), and the
// accumulators are rotated by one slot afterwards, so M generic steps from a block boundary restore the layout the
// interior blocks assume.
template <int NDIM, int M, int LG>
__device__ __forceinline__ void run_step_generic(Column<NDIM, M>& col, int D, int s, int L, MomentSmem& sm, int warp,
                                                 int lane, int acc_base) {
    cp_async_wait<kRunPrefetch - 1>();
    double P[3];
#pragma unroll
    for (int c = 0; c < NDIM; ++c) P[c] = col.ring[((s % M) * 3 + c) * 32];
    col.prefetch(D);
    const bool ok = s < col.nvalid;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        if ((unsigned)(s - m) < (unsigned)L) {  // lag s - m belongs to the run: warp-uniform
            double d2 = 0.0;
#pragma unroll
            for (int c = 0; c < NDIM; ++c) {
                const double d = P[c] - col.o[c][m];
                d2 = fma(d, d, d2);
            }
            d2 = ok ? d2 : 0.0;
            col.A1[(M - m) % M] += d2;
            col.A2[(M - m) % M] = fma(d2, d2, col.A2[(M - m) % M]);
        }
    }
    const int i_f = s - (M - 1);  // complete after this step; it is the lag of origin M-1, slot 1
    if (i_f >= 0 && i_f < L) {
        const int gs = i_f % LG;
        sm.stash_put(warp, gs, lane, col.A1[1], col.A2[1]);
        if (gs == LG - 1 || i_f == L - 1) fold_lag_group_n<LG>(sm, warp, lane, acc_base + i_f - gs, gs + 1);
    }
    // rotate by one slot: the lag of origin m becomes the lag of origin m+1 (slot j <- slot j+1); the flushed
    // slot 1 re-enters as the empty slot 0 and slot 0 moves to slot M-1
    const double r1 = col.A1[0], r2 = col.A2[0];
    col.A1[0] = 0.0;
    col.A2[0] = 0.0;
#pragma unroll
    for (int j = 1; j < M - 1; ++j) {
        col.A1[j] = col.A1[j + 1];
        col.A2[j] = col.A2[j + 1];
    }
    col.A1[M - 1] = r1;
    col.A2[M - 1] = r2;
}

// WARPS warps per CTA, two CTAs per SM: 6 warps leave 168 registers per thread (no spills), 8 warps 128.
template <int NDIM, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 2)
    self_moments_runs_kernel(const double* __restrict__ dc, int n_atoms, int Tn, int64_t pitch, int c0, int c1, int c2,
                             const RunDesc* __restrict__ runs, const int2* __restrict__ pairs, int n_pairs, int batch,
                             int K, unsigned int* work_counter, const int32_t* __restrict__ slot, double* __restrict__ out) {
    constexpr int M = kRunM, LG = kRunLG, NT = 32 * WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MomentSmem sm(smem_raw, K, WARPS, LG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Column<NDIM, M> col;
    col.ring = reinterpret_cast<double*>(smem_raw + MomentSmem::bytes(K, WARPS, LG)) + warp * (kRunRingBytes / 8) + lane;

#pragma unroll 1
    for (int i = tid; i < 2 * K; i += NT) sm.acc[i] = 0.0;
    __syncthreads();

    const long long n_work = (long long)n_pairs * n_atoms;
    const long long per_batch = (long long)n_pairs * batch;
    const int comp[3] = {c0, c1, c2};
    while (true) {
        int wk = 0;
        if (lane == 0) wk = (int)atomicAdd(work_counter, 1u);
        wk = __shfl_sync(0xffffffffu, wk, 0);
        if (wk >= n_work) break;
        // work index -> (atom batch, strip in cost order, atom of the batch)
        const int bi = (int)(wk / per_batch);
        const int rem = (int)(wk - bi * per_batch);
        const int bsize = min(batch, n_atoms - bi * batch);
        const int pi = rem / bsize, atom = bi * batch + (rem - pi * bsize);
        const int2 pr = __ldg(pairs + pi);
        const int4 rd4 = __ldg(reinterpret_cast<const int4*>(runs) + pr.x);
        const int k0 = rd4.x, D = rd4.y, L = rd4.z, acc_base = rd4.w;
        const double* row[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) row[c] = dc + ((size_t)atom * 3 + comp[c < NDIM ? c : 0]) * pitch;

        // the extra first observation dc[a][k-1] of every lag of the run (kinisi/parser.py:209)
        if (pr.y == 0) {
#pragma unroll 1
            for (int i = lane; i < L; i += 32) {
                const int p = k0 + i * D - 1;
                double d2 = 0.0;
#pragma unroll
                for (int c = 0; c < NDIM; ++c) {
                    const double d = __ldg(row[c] + p);
                    d2 = fma(d, d, d2);
                }
                atomicAdd(&sm.acc[2 * (acc_base + i)], d2);
                atomicAdd(&sm.acc[2 * (acc_base + i) + 1], d2 * d2);
            }
        }

        const int g = pr.y * kRunCols + lane;  // column: tile q = g / D, offset t = g % D
        const int q = g / D, t = g - q * D;
        const long long j0 = (long long)q * M * D + t;
        const long long p0 = j0 + k0;
        col.nvalid = 0;
        if (p0 < Tn) {
            const long long room = (Tn - 1 - p0) / D + 1;
            col.nvalid = (int)(room < (long long)(M + L - 1) ? room : (long long)(M + L - 1));
        }
        const int wmax = __reduce_max_sync(0xffffffffu, col.nvalid);
        if (wmax == 0) continue;  // warp-uniform
        const int wmin = __reduce_min_sync(0xffffffffu, col.nvalid);

        col.fetched = 0;
#pragma unroll
        for (int c = 0; c < NDIM; ++c) col.pn[c] = row[c] + p0;
#pragma unroll
        for (int f = 0; f < kRunPrefetch; ++f) col.prefetch(D);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const long long jm = j0 + (long long)m * D;
#pragma unroll
            for (int c = 0; c < NDIM; ++c) col.o[c][m] = (jm < Tn) ? __ldg(row[c] + jm) : 0.0;
            col.A1[m] = col.A2[m] = 0.0;
        }

        // steps in which at least one lane has a valid term; blocks of M steps with a static pattern run unrolled
        const int work_steps = min(M + L - 1, wmax);
        const int full = min(L, wmin);  // steps below `full` have every lane valid and lag s - m < L for all m <= s
        int s = 0;
        if (full >= M) {
            run_block_interior<NDIM, M, LG, true>(col, D, 0, sm, warp, lane, acc_base);
            s = M;
#pragma unroll 1
            for (; s + M <= full; s += M) run_block_interior<NDIM, M, LG, false>(col, D, s, sm, warp, lane, acc_base);
        }
#pragma unroll 1
        for (; s < work_steps; ++s) run_step_generic<NDIM, M, LG>(col, D, s, L, sm, warp, lane, acc_base);
        cp_async_wait<0>();  // the ring is reused by the next item
        // drain: the lags of origins M-1 .. 1 at the step after the last one are still in their slots; step on
        // without terms (flush slot 1, rotate) until they are out
#pragma unroll 1
        for (int sd = work_steps; sd < work_steps + M - 1; ++sd) {
            const int i = sd - (M - 1);
            if ((unsigned)i < (unsigned)L) {
                const int gs = i % LG;
                sm.stash_put(warp, gs, lane, col.A1[1], col.A2[1]);
                if (gs == LG - 1 || i == L - 1) fold_lag_group_n<LG>(sm, warp, lane, acc_base + i - gs, gs + 1);
            }
#pragma unroll
            for (int j = 1; j < M - 1; ++j) {
                col.A1[j] = col.A1[j + 1];
                col.A2[j] = col.A2[j + 1];
            }
        }
        // a fold group left open (the steps stopped before the run's last lag)
        const int i_last = work_steps - 1;
        if (i_last < L - 1 && (i_last % LG) != LG - 1)
            fold_lag_group_n<LG>(sm, warp, lane, acc_base + i_last - (i_last % LG), (i_last % LG) + 1);
    }
    __syncthreads();
    // one FP64 reduction per lag, moment and CTA straight into the caller's array (zeroed by the host call)
#pragma unroll 1
    for (int i = tid; i < 2 * K; i += NT) atomicAdd(out + 2 * slot[i >> 1] + (i & 1), sm.acc[i]);
}

// ---- host-side planning -------------------------------------------------------------------------
struct RunPlan {
    std::vector<RunDesc> runs;
    std::vector<int2> pairs;        // (run, strip), most expensive first
    std::vector<int32_t> slot;      // accumulator index -> position in the caller's lag array
};

// Greedy cover of the (ascending) lag grid by arithmetic progressions: each lag, smallest first, starts the longest
// progression of still-uncovered lags among the differences to its next kRunCandidates successors; progressions
// shorter than kRunMinLen are dropped and their first lag becomes a run of one.
static void plan_runs(const int32_t* lags, int K, int Tn, RunPlan& plan) {
    std::vector<char> used(K, 0);
    auto find = [&](int value, int from) -> int {  // index of an unused lag == value at or after `from`, or -1
        const int32_t* lo = std::lower_bound(lags + from, lags + K, value);
        for (; lo < lags + K && *lo == value; ++lo)
            if (!used[lo - lags]) return (int)(lo - lags);
        return -1;
    };
    int n_acc = 0;
    for (int a = 0; a < K; ++a) {
        if (used[a]) continue;
        int best_len = 1, best_d = 0;
        int tried = 0;
        for (int b = a + 1; b < K && tried < kRunCandidates; ++b) {
            if (used[b] || lags[b] == lags[a]) continue;
            ++tried;
            const int d = lags[b] - lags[a];
            int len = 1, at = a;
            while (true) {
                const long long nxt = (long long)lags[a] + (long long)len * d;
                if (nxt > Tn) break;
                const int idx = find((int)nxt, at + 1);
                if (idx < 0) break;
                at = idx;
                ++len;
            }
            if (len > best_len) {
                best_len = len;
                best_d = d;
            }
        }
        used[a] = 1;
        plan.slot.push_back(a);
        if (best_len >= kRunMinLen) {
            int at = a;
            for (int i = 1; i < best_len; ++i) {
                at = find(lags[a] + i * best_d, at + 1);
                used[at] = 1;
                plan.slot.push_back(at);
            }
            plan.runs.push_back(RunDesc{lags[a], best_d, best_len, n_acc});
            n_acc += best_len;
        } else {
            plan.runs.push_back(RunDesc{lags[a], kRunSingleStride, 1, n_acc});
            n_acc += 1;
        }
    }
    // (run, strip) pairs with the number of steps of their first column as the cost
    struct Item {
        int run, strip, cost;
    };
    std::vector<Item> items;
    for (size_t r = 0; r < plan.runs.size(); ++r) {
        const RunDesc& rd = plan.runs[r];
        const long long span = (long long)Tn - rd.k0;  // origins with at least the first lag valid
        const long long tiles = span <= 0 ? 0 : (span + (long long)kRunM * rd.delta - 1) / ((long long)kRunM * rd.delta);
        const long long cols = tiles * rd.delta;
        // strip 0 always exists: it also carries the first-observation terms of the run
        const int nstrip = std::max<int>(1, (int)((cols + kRunCols - 1) / kRunCols));
        for (int sp = 0; sp < nstrip; ++sp) {
            const long long g = (long long)sp * kRunCols;
            const long long q = g / rd.delta, t = g - q * rd.delta;
            const long long p0 = q * kRunM * rd.delta + t + rd.k0;
            long long cost = 0;
            if (p0 < Tn) cost = std::min<long long>(kRunM + rd.len - 1, (Tn - 1 - p0) / rd.delta + 1);
            items.push_back({(int)r, sp, (int)cost});
        }
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.cost > y.cost; });
    plan.pairs.reserve(items.size());
    for (const Item& it : items) plan.pairs.push_back(make_int2(it.run, it.strip));
}

static size_t runs_pair_capacity(int K, int Tn) {
    // per run: columns <= (Tn / (M*D) + 1) * D <= Tn / M + D; a run of >= 3 lags has D <= Tn / 2, a run of one has
    // D = kRunSingleStride
    const size_t long_runs = (size_t)K / kRunMinLen + 1;
    const size_t per_long = (size_t)(5ll * Tn / (8 * kRunCols)) + 3;
    const size_t per_single = (size_t)(Tn / (kRunM * kRunCols)) + 3;
    return long_runs * per_long + (size_t)K * per_single;
}

template <int NDIM, int WARPS>
static int launch_runs_w(int sm_count, long long n_work, cudaStream_t st, const double* dc, int n_atoms, int Tn,
                         int64_t pitch, const int comp[3], const RunDesc* runs, const int2* pairs, int n_pairs, int batch,
                         int K, unsigned int* counter, const int32_t* slot, double* out) {
    const int grid = (int)std::min<long long>((n_work + WARPS - 1) / WARPS, 2 * sm_count);
    const size_t smem = MomentSmem::bytes(K, WARPS, kRunLG) + (size_t)WARPS * kRunRingBytes;
    KB2_CUDA(cudaFuncSetAttribute(self_moments_runs_kernel<NDIM, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    self_moments_runs_kernel<NDIM, WARPS><<<grid, 32 * WARPS, smem, st>>>(dc, n_atoms, Tn, pitch, comp[0], comp[1], comp[2],
                                                                          runs, pairs, n_pairs, batch, K, counter, slot, out);
    KB2_LAUNCH_CHECK();
    return 0;
}

template <int NDIM>
static int launch_runs(int warps, int sm_count, long long n_work, cudaStream_t st, const double* dc, int n_atoms, int Tn,
                       int64_t pitch, const int comp[3], const RunDesc* runs, const int2* pairs, int n_pairs, int batch,
                       int K, unsigned int* counter, const int32_t* slot, double* out) {
    if (warps == 8)
        return launch_runs_w<NDIM, 8>(sm_count, n_work, st, dc, n_atoms, Tn, pitch, comp, runs, pairs, n_pairs, batch, K,
                                      counter, slot, out);
    return launch_runs_w<NDIM, 6>(sm_count, n_work, st, dc, n_atoms, Tn, pitch, comp, runs, pairs, n_pairs, batch, K, counter,
                                  slot, out);
}

// Tuning knob for experiments (tools/k2probe.py): KB2_RUNS_WARPS=6|8 picks the CTA size of the run kernel.
static int runs_warps() {
    static int w = [] {
        const char* e = getenv("KB2_RUNS_WARPS");
        return (e && atoi(e) == 8) ? 8 : 6;
    }();
    return w;
}

// workspace layout: the plan: 64 ints (work counter) | runs | pairs | slots
static size_t runs_plan_bytes(int k, int t) {
    return 256 + align256((size_t)k * sizeof(RunDesc)) + align256(runs_pair_capacity(k, t) * sizeof(int2)) +
           align256((size_t)k * sizeof(int32_t));
}

}  // namespace kb2

using namespace kb2;

extern "C" size_t kb2_self_moments_runs_workspace_bytes(int n_lags, int n_frames) {
    const int k = n_lags < 1 ? 1 : (n_lags > kMaxLags ? kMaxLags : n_lags);
    const int t = n_frames < 1 ? 1 : n_frames;
    return runs_plan_bytes(k, t);
}

extern "C" int kb2_self_moments_runs(const double* dc, int n_atoms, int n_frames, int64_t pitch,
                                     const int32_t* lags_host, int n_lags, int dim_mask, void* workspace, double* sums,
                                     void* stream) {
    KB2_CHECK(dc && lags_host && workspace && sums, "kb2_self_moments_runs: null pointer");
    KB2_CHECK(n_atoms > 0 && n_frames > 0 && n_lags > 0, "kb2_self_moments_runs: empty problem");
    KB2_CHECK(pitch >= n_frames && pitch % 2 == 0, "kb2_self_moments_runs: pitch must be even and >= n_frames");
    KB2_CHECK(dim_mask >= 1 && dim_mask <= 7, "kb2_self_moments_runs: dim_mask must be in 1..7");
    for (int i = 0; i < n_lags; ++i) {
        KB2_CHECK(lags_host[i] >= 1 && lags_host[i] <= n_frames, "kb2_self_moments_runs: lag %d out of range", i);
        KB2_CHECK(i == 0 || lags_host[i] >= lags_host[i - 1], "kb2_self_moments_runs: lags must ascend");
    }
    cudaStream_t st = (cudaStream_t)stream;
    int comp[3] = {0, 0, 0}, ndim = 0;
    for (int c = 0; c < 3; ++c)
        if (dim_mask & (1 << c)) comp[ndim++] = c;
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, 0);

    const int kcap = n_lags > kMaxLags ? kMaxLags : n_lags;
    char* d_plan = (char*)workspace;
    KB2_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_lags * 2 * sizeof(double), st));
    const size_t pair_cap = runs_pair_capacity(kcap, n_frames);

    std::vector<unsigned char> packed;
    for (int l0 = 0; l0 < n_lags; l0 += kMaxLags) {
        const int K = min(kMaxLags, n_lags - l0);
        RunPlan plan;
        plan_runs(lags_host + l0, K, n_frames, plan);
        KB2_CHECK(plan.pairs.size() <= pair_cap && (int)plan.slot.size() == K,
                  "kb2_self_moments_runs: internal planning error");
        // one packed upload: zeroed work counter, runs, pairs, slots
        const size_t off_runs = 256;
        const size_t off_pairs = off_runs + align256(plan.runs.size() * sizeof(RunDesc));
        const size_t off_slot = off_pairs + align256(plan.pairs.size() * sizeof(int2));
        const size_t total = off_slot + align256((size_t)K * sizeof(int32_t));
        KB2_CHECK(total <= runs_plan_bytes(kcap, n_frames), "kb2_self_moments_runs: plan exceeds the workspace");
        packed.assign(total, 0);
        memcpy(packed.data() + off_runs, plan.runs.data(), plan.runs.size() * sizeof(RunDesc));
        memcpy(packed.data() + off_pairs, plan.pairs.data(), plan.pairs.size() * sizeof(int2));
        memcpy(packed.data() + off_slot, plan.slot.data(), (size_t)K * sizeof(int32_t));
        KB2_CUDA(cudaMemcpyAsync(d_plan, packed.data(), total, cudaMemcpyHostToDevice, st));

        const long long n_work = (long long)plan.pairs.size() * n_atoms;
        KB2_CHECK(n_work < (1ll << 31) - (1 << 20), "kb2_self_moments_runs: too many work items");
        // atoms per batch: their trajectories (3 * pitch * 8 B each) fit in a fraction of the L2 together
        const int batch = (int)std::max<long long>(1, std::min<long long>(n_atoms, kRunBatchBytes / (3 * pitch * 8)));
        const RunDesc* d_runs = (const RunDesc*)(d_plan + off_runs);
        const int2* d_pairs = (const int2*)(d_plan + off_pairs);
        const int32_t* d_slot = (const int32_t*)(d_plan + off_slot);
        int rc;
        if (ndim == 1)
            rc = launch_runs<1>(runs_warps(), sm_count, n_work, st, dc, n_atoms, n_frames, pitch, comp, d_runs, d_pairs,
                                (int)plan.pairs.size(), batch, K, (unsigned int*)d_plan, d_slot, sums + 2 * l0);
        else if (ndim == 2)
            rc = launch_runs<2>(runs_warps(), sm_count, n_work, st, dc, n_atoms, n_frames, pitch, comp, d_runs, d_pairs,
                                (int)plan.pairs.size(), batch, K, (unsigned int*)d_plan, d_slot, sums + 2 * l0);
        else
            rc = launch_runs<3>(runs_warps(), sm_count, n_work, st, dc, n_atoms, n_frames, pitch, comp, d_runs, d_pairs,
                                (int)plan.pairs.size(), batch, K, (unsigned int*)d_plan, d_slot, sums + 2 * l0);
        if (rc) return rc;
    }
    return 0;
}

// The plan for a lag grid, for tests and diagnostics: fills run_k0/run_delta/run_len (capacity max_runs) with the
// runs of >= 3 lags and returns their number; *n_rest receives the number of lags that became runs of one.
extern "C" int kb2_self_moments_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* run_k0,
                                     int32_t* run_delta, int32_t* run_len, int max_runs, int* n_rest) {
    if (!lags_host || n_lags <= 0) return 0;
    RunPlan plan;
    plan_runs(lags_host, std::min(n_lags, kMaxLags), n_frames, plan);
    int n = 0, rest = 0;
    for (const RunDesc& rd : plan.runs) {
        if (rd.len < kRunMinLen) {
            ++rest;
            continue;
        }
        if (n < max_runs) {
            if (run_k0) run_k0[n] = rd.k0;
            if (run_delta) run_delta[n] = rd.delta;
            if (run_len) run_len[n] = rd.len;
        }
        ++n;
    }
    if (n_rest) *n_rest = rest;
    return n;
}
