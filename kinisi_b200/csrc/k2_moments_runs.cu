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
// Drift.  np.linspace(..., dtype=int) truncates: the spacing is D most of the time and D+1 (or D-1) every so often
// (every 11th lag at 100,000 frames / 100 lags).  Such a grid is ONE run with drift, k_i = k0 + i*D + e_i, e_i
// changing by delta = +-1 at a few lag indices at least M apart.  The partner stream of a column stays where it is,
// x[j + k0 + s*D + e_fix]; what moves is the ORIGIN of a slot, by one frame, when the lag of that slot passes a change
// (see DriftCursor): one origin reload per step for the M steps after a change, nothing per term.
//
// The host splits the lag grid into arithmetic runs (greedy cover; a lag that fits no run of >= 3 becomes a run of
// one) and then chains runs that continue each other's progression with a one-sample shift into drift runs.  Work items are (run, 32-column strip, atom); every warp of the persistent CTAs pulls its own items from
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
constexpr int kRunChainMaxSeg = 16;        // progressions longer than this are not chained into drift runs
constexpr int kRunPrefetch = 6;            // steps between the cp.async of a partner sample and its use (< kRunM)
constexpr long long kRunBatchBytes = 24ll << 20;  // trajectory bytes of one atom batch (L2 is 126 MB)
constexpr int kRunRingBytes = kRunM * 3 * 2 * 32 * 8;  // per warp: slot x component x {sample, neighbour} x lane

constexpr int kRunMaxLags = 192;           // lags per launch: every warp keeps private accumulators [K][2] in shared memory

struct RunDesc {
    int k0, delta, len, off;  // lags k0 + i*delta + e_i, i < len; accumulator index of lag i is off + i
    int chg_off, n_chg;       // e_i = shift * #{changes <= i}; the change indices are chg[chg_off .. chg_off + n_chg)
    int shift, pad;           // +1 or -1
};
static_assert(sizeof(RunDesc) == 32, "RunDesc is read as two int4");

// Warp-uniform position of a step in the drift pattern of its run.
//
// Lags k_i = k0 + i*D + shift * E_i, E_i = #{changes <= i}.  The PARTNER stream of a column is fixed for the whole run,
// x[j + k0 + s*D + e_fix] with e_fix = max_i (shift * E_i); the ORIGIN of slot m for lag i = s - m is then
// x[j + m*D + sigma_i], sigma_i = e_fix - shift * E_i >= 0: when the lag of a slot passes a change (at step c + m for the
// change at lag index c) the slot takes the neighbouring origin, one frame down (shift = +1) or up (shift = -1).  So a
// change costs one origin reload per step for M steps and no per-term work.  The origins [0, sigma_i) of lag i that no
// column reaches are added by the strip that also adds the first-observation terms.
struct DriftCursor {
    int cnext;  // lag index of the change whose slots are being switched, or of the next one
    int ci;     // its position in the change list
    __device__ __forceinline__ void init(const int32_t* chg, int n) {
        ci = 0;
        cnext = n > 0 ? __ldg(chg) : 0x7fffffff;
    }
    // at the start of step s: a change is done once all M slots have switched (steps cnext .. cnext + M - 1)
    __device__ __forceinline__ void advance(int s, const int32_t* chg, int n) {
        if (s - cnext >= kRunM) {
            ++ci;
            cnext = ci < n ? __ldg(chg + ci) : 0x7fffffff;
        }
    }
    // the slot that switches at step s, or a value outside [0, M)
    __device__ __forceinline__ int slot(int s) const { return s - cnext; }
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

// As fold_lag_group, for a group whose first `cnt` stash rows are live; lag g of the group goes to the WARP's own
// accumulators wacc[i0 + g] with a plain read-modify-write (a shared-memory FP64 atomicAdd is a compare-and-swap loop:
// 19 % of the kernel's stall samples when all the warps of an SM fold into one array).
template <int LG>
__device__ __forceinline__ void fold_lag_group_n(MomentSmem& sm, double* wacc, int warp, int lane, int i0, int cnt) {
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
        double2* a = reinterpret_cast<double2*>(wacc) + (i0 + g);
        double2 v = *a;
        v.x += t1;
        v.y += t2;
        *a = v;
    }
    __syncwarp();
}

// State of one column while it walks its steps.
template <int NDIM, int M, bool DR>
struct Column {
    double o[3][M];        // origins
    double A1[M], A2[M];   // running sums; at a step s = 0 (mod M) lag i lives in slot i mod M
    const double* pn[3];   // address of the undrifted partner sample of the next step to prefetch
    double* ring;          // this lane's column of the warp's ring: sample (slot, c, h) at ring[((slot*3 + c)*2 + h)*32]
                           // (h = 0 only and no factor 2 in kernels without drift, DR = false)
    int rem;               // Tn - 1 - (undrifted partner index) of the current step: offset e is valid iff e <= rem
    int rem_pf;            // the same for the next step to prefetch
    int fetched;           // steps prefetched so far
    DriftCursor pf;        // drift position of the prefetch stream

    // Prefetch the partner sample of step `fetched` and, if a slot switches its origin in that step, the new origin
    // (h = 1 of the ring slot): its address is the partner's plus a warp-uniform offset,
    //   j + m*D + sigma - (j + k0 + e_fix + fetched*D) = sigma_after - k0 - e_fix - c*D   with m = fetched - c.
    static constexpr int H = DR ? 2 : 1;  // samples per (slot, component)
    __device__ __forceinline__ void prefetch(int D, const int32_t* chg, int n_chg, int shift, int k0, int e_fix) {
        double* dst = ring + (fetched % M) * (3 * H * 32);
        if (0 <= rem_pf) {
#pragma unroll
            for (int c = 0; c < NDIM; ++c) cp_async8(dst + (c * H) * 32, pn[c]);
        }
        if (DR) {
            pf.advance(fetched, chg, n_chg);
            const int ms = pf.slot(fetched);
            if (ms >= 0 && ms < M) {
                const int sigma_after = e_fix - shift * (pf.ci + 1);
                const int ooff = sigma_after - k0 - e_fix - pf.cnext * D;  // < 0: the origin precedes the partner
                if (ooff <= rem_pf) {
#pragma unroll
                    for (int c = 0; c < NDIM; ++c) cp_async8(dst + (c * 2 + 1) * 32, pn[c] + ooff);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NDIM; ++c) pn[c] += D;
        rem_pf -= D;
        ++fetched;
        cp_async_commit();
    }

    // slot ms takes the origin staged at h = 1 of ring slot `rs`: M predicated loads straight into the slot's registers
    // (a register cannot be indexed at run time; a switch is if-converted into twice as many predicated moves)
    __device__ __forceinline__ void switch_origin(int ms, int rs) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            if (m == ms) {
#pragma unroll
                for (int c = 0; c < NDIM; ++c) o[c][m] = ring[((rs * 3 + c) * 2 + 1) * 32];
            }
        }
    }
};

struct RunCtx {  // warp-uniform description of the current item
    int D, L, acc_base, n_chg, shift, k0, e_fix;
    const int32_t* chg;
};

// M consecutive steps s0 .. s0+M-1 (s0 a multiple of M) whose partners lie inside the trajectory for all lanes and
// whose (step, origin) pairs are lags of the run in a static pattern:
//   FIRST = false: all of them (s0 >= M, s0 + M <= L);
//   FIRST = true:  s0 = 0 and L >= M, the triangular start: origin m joins at step m, only lag 0 completes.
// DRIFT = false requires that no slot switches its origin in these steps (the previous change is done and
// cur.cnext > s0 + M - 1).
template <int NDIM, int M, int LG, bool FIRST, bool DRIFT, bool DR>
__device__ __forceinline__ void run_block_interior(Column<NDIM, M, DR>& col, DriftCursor& cur, const RunCtx& rc, int s0,
                                                   MomentSmem& sm, double* wacc, int warp, int lane) {
    static_assert(M % LG == 0, "the fold group must divide the register window");
#pragma unroll
    for (int u = 0; u < M; ++u) {
        cp_async_wait<kRunPrefetch - 1>();
        double P[3];
#pragma unroll
        for (int c = 0; c < NDIM; ++c) P[c] = col.ring[((u * 3 + c) * (DR ? 2 : 1)) * 32];
        if (DRIFT) {
            cur.advance(s0 + u, rc.chg, rc.n_chg);
            const int ms = cur.slot(s0 + u);
            if (ms >= 0 && ms < M) col.switch_origin(ms, u);  // warp-uniform
        }
        col.prefetch(rc.D, rc.chg, rc.n_chg, rc.shift, rc.k0, rc.e_fix);
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
            if (gs == LG - 1) fold_lag_group_n<LG>(sm, wacc, warp, lane, rc.acc_base + s0 + u - (M - 1) - (LG - 1), LG);
        }
    }
    col.rem -= M * rc.D;
}

// One step at any position: origin m pairs with slot (M - m) % M (the layout of a step s = 0 mod M), and the
// accumulators are rotated by one slot afterwards, so M generic steps from a block boundary restore the layout the
// interior blocks assume.  Validity is checked per term (lag inside the run: warp-uniform; partner inside the
// trajectory: per lane).
template <int NDIM, int M, int LG, bool DR>
__device__ __forceinline__ void run_step_generic(Column<NDIM, M, DR>& col, DriftCursor& cur, const RunCtx& rc, int s,
                                                 MomentSmem& sm, double* wacc, int warp, int lane) {
    cp_async_wait<kRunPrefetch - 1>();
    double P[3];
#pragma unroll
    for (int c = 0; c < NDIM; ++c) P[c] = col.ring[(((s % M) * 3 + c) * (DR ? 2 : 1)) * 32];
    if (DR) {
        cur.advance(s, rc.chg, rc.n_chg);
        const int ms = cur.slot(s);
        if (ms >= 0 && ms < M) col.switch_origin(ms, s % M);  // warp-uniform
    }
    col.prefetch(rc.D, rc.chg, rc.n_chg, rc.shift, rc.k0, rc.e_fix);
    const bool ok_p = 0 <= col.rem;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        if ((unsigned)(s - m) < (unsigned)rc.L) {  // lag s - m belongs to the run: warp-uniform
            double d2 = 0.0;
#pragma unroll
            for (int c = 0; c < NDIM; ++c) {
                const double d = P[c] - col.o[c][m];
                d2 = fma(d, d, d2);
            }
            d2 = ok_p ? d2 : 0.0;
            col.A1[(M - m) % M] += d2;
            col.A2[(M - m) % M] = fma(d2, d2, col.A2[(M - m) % M]);
        }
    }
    col.rem -= rc.D;
    const int i_f = s - (M - 1);  // complete after this step; it is the lag of origin M-1, slot 1
    if (i_f >= 0 && i_f < rc.L) {
        const int gs = i_f % LG;
        sm.stash_put(warp, gs, lane, col.A1[1], col.A2[1]);
        if (gs == LG - 1 || i_f == rc.L - 1) fold_lag_group_n<LG>(sm, wacc, warp, lane, rc.acc_base + i_f - gs, gs + 1);
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
template <int NDIM, int WARPS, bool DR>
__global__ void __launch_bounds__(32 * WARPS, 2)
    self_moments_runs_kernel(const double* __restrict__ dc, int n_atoms, int Tn, int64_t pitch, int c0, int c1, int c2,
                             const RunDesc* __restrict__ runs, const int2* __restrict__ pairs, int n_pairs, int batch,
                             const int32_t* __restrict__ chg, int K, unsigned int* work_counter,
                             const int32_t* __restrict__ slot, double* __restrict__ out) {
    constexpr int M = kRunM, LG = kRunLG, NT = 32 * WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MomentSmem sm(smem_raw, K, WARPS, LG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Column<NDIM, M, DR> col;
    col.ring = reinterpret_cast<double*>(smem_raw + MomentSmem::bytes(K, WARPS, LG)) +
               warp * (kRunRingBytes / 8 / (DR ? 1 : 2)) + lane;

    // per-warp accumulators [WARPS][K][2] behind the rings
    double* const wacc_all = reinterpret_cast<double*>(smem_raw + MomentSmem::bytes(K, WARPS, LG) +
                                                       (size_t)WARPS * kRunRingBytes / (DR ? 1 : 2));
    double* const wacc = wacc_all + (size_t)warp * 2 * K;
#pragma unroll 1
    for (int i = tid; i < 2 * K * WARPS; i += NT) wacc_all[i] = 0.0;
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
        const int4 rd4 = __ldg(reinterpret_cast<const int4*>(runs) + 2 * pr.x);
        const int4 rd5 = __ldg(reinterpret_cast<const int4*>(runs) + 2 * pr.x + 1);
        const int k0 = rd4.x;
        RunCtx rc;
        rc.D = rd4.y;
        rc.L = rd4.z;
        rc.acc_base = rd4.w;
        rc.chg = chg + rd5.x;
        rc.n_chg = rd5.y;
        rc.shift = rd5.z;
        rc.k0 = k0;
        rc.e_fix = rc.shift > 0 ? rc.n_chg : 0;  // max_i shift * E_i
        const int D = rc.D, L = rc.L;
        const double* row[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) row[c] = dc + ((size_t)atom * 3 + comp[c < NDIM ? c : 0]) * pitch;

        // the extra first observation dc[a][k-1] of every lag of the run (kinisi/parser.py:209), and the origins
        // [0, sigma_i) of a drifting run that no column reaches (see DriftCursor)
        if (pr.y == 0) {
#pragma unroll 1
            for (int i = lane; i < L; i += 32) {
                int e = 0;
                for (int j = 0; j < rc.n_chg; ++j) e += (__ldg(rc.chg + j) <= i) ? rc.shift : 0;
                const int k = k0 + i * D + e;
                double s1 = 0.0, s2 = 0.0;
                {
                    double d2 = 0.0;
#pragma unroll
                    for (int c = 0; c < NDIM; ++c) {
                        const double d = __ldg(row[c] + k - 1);
                        d2 = fma(d, d, d2);
                    }
                    s1 = d2;
                    s2 = d2 * d2;
                }
                for (int jj = 0; jj < rc.e_fix - e && jj + k < Tn; ++jj) {
                    double d2 = 0.0;
#pragma unroll
                    for (int c = 0; c < NDIM; ++c) {
                        const double d = __ldg(row[c] + jj + k) - __ldg(row[c] + jj);
                        d2 = fma(d, d, d2);
                    }
                    s1 += d2;
                    s2 = fma(d2, d2, s2);
                }
                wacc[2 * (rc.acc_base + i)] += s1;  // one lane per lag
                wacc[2 * (rc.acc_base + i) + 1] += s2;
            }
            __syncwarp();
        }

        const int g = pr.y * kRunCols + lane;  // column: tile q = g / D, offset t = g % D
        const int q = g / D, t = g - q * D;
        const long long j0 = (long long)q * M * D + t;
        const long long p0 = j0 + k0 + rc.e_fix;  // partner index of step 0 (the partner stream never drifts)
        // steps in which this lane has a valid partner
        int n_ok = 0;
        if (p0 < Tn) n_ok = (int)min((long long)(M + L - 1), (Tn - 1 - p0) / D + 1);
        const int wmax = __reduce_max_sync(0xffffffffu, n_ok);
        if (wmax == 0) continue;  // warp-uniform
        const int wmin = __reduce_min_sync(0xffffffffu, n_ok);

        col.fetched = 0;
        col.pf.init(rc.chg, rc.n_chg);
        // Tn - 1 - p0 can be far below zero for lanes past the end; clamp so that the per-step decrements cannot wrap
        col.rem = col.rem_pf = (int)max((long long)(Tn - 1) - p0, -(1ll << 30));
#pragma unroll
        for (int c = 0; c < NDIM; ++c) col.pn[c] = row[c] + p0;
#pragma unroll
        for (int f = 0; f < kRunPrefetch; ++f) col.prefetch(D, rc.chg, rc.n_chg, rc.shift, rc.k0, rc.e_fix);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const long long jm = j0 + (long long)m * D + rc.e_fix;  // sigma of lag 0 is e_fix
#pragma unroll
            for (int c = 0; c < NDIM; ++c) col.o[c][m] = (jm < Tn) ? __ldg(row[c] + jm) : 0.0;
            col.A1[m] = col.A2[m] = 0.0;
        }
        DriftCursor cur;
        cur.init(rc.chg, rc.n_chg);

        // steps in which at least one lane has a valid term; blocks of M steps with a static pattern run unrolled
        const int work_steps = min(M + L - 1, wmax);
        const int full = min(L, wmin);  // steps below `full` have every lane valid and lag s - m < L for all m <= s
        int s = 0;
#pragma unroll 1
        while (s < work_steps) {
            if (DR) cur.advance(s, rc.chg, rc.n_chg);  // (idempotent) so that the dispatch below sees the current change
            if ((s % M) == 0 && s + M <= full && (s > 0 || cur.cnext > M - 1)) {
                if (s == 0)
                    run_block_interior<NDIM, M, LG, true, false, DR>(col, cur, rc, 0, sm, wacc, warp, lane);
                else if (!DR || cur.cnext > s + M - 1)
                    run_block_interior<NDIM, M, LG, false, false, DR>(col, cur, rc, s, sm, wacc, warp, lane);
                else
                    run_block_interior<NDIM, M, LG, false, DR, DR>(col, cur, rc, s, sm, wacc, warp, lane);
                s += M;
            } else {
                run_step_generic<NDIM, M, LG, DR>(col, cur, rc, s, sm, wacc, warp, lane);
                ++s;
            }
        }
        const int acc_base = rc.acc_base;
        cp_async_wait<0>();  // the ring is reused by the next item
        // drain: the lags of origins M-1 .. 1 at the step after the last one are still in their slots; step on
        // without terms (flush slot 1, rotate) until they are out
#pragma unroll 1
        for (int sd = work_steps; sd < work_steps + M - 1; ++sd) {
            const int i = sd - (M - 1);
            if ((unsigned)i < (unsigned)L) {
                const int gs = i % LG;
                sm.stash_put(warp, gs, lane, col.A1[1], col.A2[1]);
                if (gs == LG - 1 || i == L - 1) fold_lag_group_n<LG>(sm, wacc, warp, lane, acc_base + i - gs, gs + 1);
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
            fold_lag_group_n<LG>(sm, wacc, warp, lane, acc_base + i_last - (i_last % LG), (i_last % LG) + 1);
    }
    __syncthreads();
    // one FP64 reduction per lag, moment and CTA straight into the caller's array (zeroed by the host call)
#pragma unroll 1
    for (int i = tid; i < 2 * K; i += NT) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) t += wacc_all[(size_t)w * 2 * K + i];
        atomicAdd(out + 2 * slot[i >> 1] + (i & 1), t);
    }
}

// ---- host-side planning -------------------------------------------------------------------------
struct RunPlan {
    std::vector<RunDesc> runs;
    std::vector<int2> pairs;        // (run, strip), most expensive first
    std::vector<int32_t> slot;      // accumulator index -> position in the caller's lag array
    std::vector<int32_t> chg;       // drift change indices of all runs
};

// Step 1, greedy cover of the (ascending) lag grid by arithmetic progressions: each lag, smallest first, starts the
// longest progression of still-uncovered lags among the differences to its next kRunCandidates successors;
// progressions shorter than kRunMinLen are dropped and their first lag becomes a segment of one.
// Step 2, chaining: a segment whose first lag is (last lag of a chain) + D + shift, shift = +-1 fixed per chain, with
// the same D (or of one lag), continues that chain as a drift change, provided the changes stay >= M lags apart.
static void plan_runs(const int32_t* lags, int K, int Tn, RunPlan& plan) {
    struct Segment {
        int k0, d, len;
        std::vector<int32_t> slots;
        bool taken;
    };
    std::vector<Segment> segs;
    std::vector<char> used(K, 0);
    auto find = [&](int value, int from) -> int {  // index of an unused lag == value at or after `from`, or -1
        const int32_t* lo = std::lower_bound(lags + from, lags + K, value);
        for (; lo < lags + K && *lo == value; ++lo)
            if (!used[lo - lags]) return (int)(lo - lags);
        return -1;
    };
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
        Segment sg{lags[a], kRunSingleStride, 1, {a}, false};
        used[a] = 1;
        if (best_len >= kRunMinLen) {
            sg.d = best_d;
            sg.len = best_len;
            int at = a;
            for (int i = 1; i < best_len; ++i) {
                at = find(lags[a] + i * best_d, at + 1);
                used[at] = 1;
                sg.slots.push_back(at);
            }
        }
        segs.push_back(std::move(sg));
    }
    // segments are in ascending order of their first lag (the greedy walks the grid upwards)
    int n_acc = 0;
    for (size_t h = 0; h < segs.size(); ++h) {
        if (segs[h].taken) continue;
        segs[h].taken = true;
        RunDesc rd{segs[h].k0, segs[h].d, segs[h].len, n_acc, (int)plan.chg.size(), 0, 1, 0};
        plan.slot.insert(plan.slot.end(), segs[h].slots.begin(), segs[h].slots.end());
        static const int chain_max = [] { const char* e = getenv("KB2_RUNS_CHAIN_MAX"); return e ? atoi(e) : kRunChainMaxSeg; }();
        if (segs[h].len >= kRunMinLen && segs[h].len <= chain_max) {
            int last_len = segs[h].len;  // length of the chain's last segment
            bool first = true;           // ... which is still the chain's first segment
            int shift = 0;
            while (true) {
                if (!first && last_len < kRunM) break;  // the last segment would become interior: too short
                const long long last_lag = (long long)rd.k0 + (long long)(rd.len - 1) * rd.delta + (long long)shift * rd.n_chg;
                size_t found = segs.size();
                int found_shift = 0;
                for (size_t n = h + 1; n < segs.size() && found == segs.size(); ++n) {
                    if (segs[n].taken) continue;
                    if (segs[n].len > 1 && (segs[n].d != rd.delta || segs[n].len > chain_max)) continue;
                    for (int sh = -1; sh <= 1; sh += 2) {
                        if (shift != 0 && sh != shift) continue;
                        if ((long long)segs[n].k0 == last_lag + rd.delta + sh) {
                            found = n;
                            found_shift = sh;
                        }
                    }
                }
                if (found == segs.size()) break;
                shift = found_shift;
                segs[found].taken = true;
                plan.chg.push_back(rd.len);
                rd.n_chg += 1;
                rd.len += segs[found].len;
                plan.slot.insert(plan.slot.end(), segs[found].slots.begin(), segs[found].slots.end());
                last_len = segs[found].len;
                first = false;
            }
            rd.shift = shift == 0 ? 1 : shift;
        }
        n_acc += rd.len;
        plan.runs.push_back(rd);
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

template <int NDIM, int WARPS, bool DR>
static int launch_runs_w(int sm_count, long long n_work, cudaStream_t st, const double* dc, int n_atoms, int Tn,
                         int64_t pitch, const int comp[3], const RunDesc* runs, const int2* pairs, int n_pairs, int batch,
                         const int32_t* chg, int K, unsigned int* counter, const int32_t* slot, double* out) {
    const int grid = (int)std::min<long long>((n_work + WARPS - 1) / WARPS, 2 * sm_count);
    const size_t smem = MomentSmem::bytes(K, WARPS, kRunLG) + (size_t)WARPS * kRunRingBytes / (DR ? 1 : 2) +
                        (size_t)WARPS * 16 * K;
    KB2_CUDA(cudaFuncSetAttribute(self_moments_runs_kernel<NDIM, WARPS, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    self_moments_runs_kernel<NDIM, WARPS, DR><<<grid, 32 * WARPS, smem, st>>>(dc, n_atoms, Tn, pitch, comp[0], comp[1], comp[2],
                                                                          runs, pairs, n_pairs, batch, chg, K, counter, slot, out);
    KB2_LAUNCH_CHECK();
    return 0;
}

template <int NDIM>
static int launch_runs(int warps, bool drift, int sm_count, long long n_work, cudaStream_t st, const double* dc,
                       int n_atoms, int Tn, int64_t pitch, const int comp[3], const RunDesc* runs, const int2* pairs,
                       int n_pairs, int batch, const int32_t* chg, int K, unsigned int* counter, const int32_t* slot,
                       double* out) {
#define KB2_RUNS(W, DRF)                                                                                                \
    return launch_runs_w<NDIM, W, DRF>(sm_count, n_work, st, dc, n_atoms, Tn, pitch, comp, runs, pairs, n_pairs, batch, chg, \
                                       K, counter, slot, out)
    if (warps == 8) {
        if (drift) KB2_RUNS(8, true);
        KB2_RUNS(8, false);
    }
    if (drift) KB2_RUNS(6, true);
    KB2_RUNS(6, false);
#undef KB2_RUNS
}

// Tuning knob for experiments (tools/k2probe.py): KB2_RUNS_WARPS=6|8 picks the CTA size of the run kernel.
static int runs_warps() {
    static int w = [] {
        const char* e = getenv("KB2_RUNS_WARPS");
        return (e && atoi(e) == 8) ? 8 : 6;
    }();
    return w;
}

// Packed plan of one launch (<= kRunMaxLags lags): header (256 B) | runs | pairs | slots | drift change indices.
// A grid longer than kRunMaxLags is a sequence of such blocks at a fixed stride.
struct PlanHeader {
    int K, n_runs, n_pairs, drift;
    int off_runs, off_pairs, off_slot, off_chg;
    int l0, pad[7];
};
static_assert(sizeof(PlanHeader) <= 256, "the header has 256 bytes");

static size_t runs_plan_bytes(int k, int t) {
    return 256 + align256((size_t)k * sizeof(RunDesc)) + align256(runs_pair_capacity(k, t) * sizeof(int2)) +
           2 * align256((size_t)k * sizeof(int32_t));
}

static int check_lags(const char* fn, const int32_t* lags_host, int n_lags, int n_frames) {
    KB2_CHECK(lags_host && n_lags > 0 && n_frames > 0, "%s: empty problem", fn);
    for (int i = 0; i < n_lags; ++i) {
        KB2_CHECK(lags_host[i] >= 1 && lags_host[i] <= n_frames, "%s: lag %d out of range", fn, i);
        KB2_CHECK(i == 0 || lags_host[i] >= lags_host[i - 1], "%s: lags must ascend", fn);
    }
    return 0;
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_self_moments_runs_blocks(int n_lags) { return n_lags < 1 ? 1 : (n_lags + kRunMaxLags - 1) / kRunMaxLags; }

extern "C" size_t kb2_self_moments_runs_workspace_bytes(int n_lags, int n_frames) {
    const int t = n_frames < 1 ? 1 : n_frames;
    const int n = n_lags < 1 ? 1 : n_lags;
    const int blocks = (n + kRunMaxLags - 1) / kRunMaxLags;
    return (size_t)blocks * runs_plan_bytes(n > kRunMaxLags ? kRunMaxLags : n, t);
}

// Plans the lag grid on the host into `plan_host` (kb2_self_moments_runs_workspace_bytes bytes of HOST memory; pinned
// if the caller wants its upload to be asynchronous).  The plan depends on (lags, n_frames) only: callers cache it.
extern "C" int kb2_self_moments_runs_plan_pack(const int32_t* lags_host, int n_lags, int n_frames, void* plan_host) {
    if (int rc = check_lags("kb2_self_moments_runs_plan_pack", lags_host, n_lags, n_frames)) return rc;
    KB2_CHECK(plan_host, "kb2_self_moments_runs_plan_pack: null pointer");
    const int kcap = n_lags > kRunMaxLags ? kRunMaxLags : n_lags;
    const size_t stride = runs_plan_bytes(kcap, n_frames);
    const size_t pair_cap = runs_pair_capacity(kcap, n_frames);
    unsigned char* base = (unsigned char*)plan_host;
    int b = 0;
    for (int l0 = 0; l0 < n_lags; l0 += kRunMaxLags, ++b) {
        const int K = std::min(kRunMaxLags, n_lags - l0);
        RunPlan plan;
        plan_runs(lags_host + l0, K, n_frames, plan);
        KB2_CHECK(plan.pairs.size() <= pair_cap && (int)plan.slot.size() == K,
                  "kb2_self_moments_runs_plan_pack: internal planning error");
        PlanHeader h{};
        h.K = K;
        h.n_runs = (int)plan.runs.size();
        h.n_pairs = (int)plan.pairs.size();
        h.drift = plan.chg.empty() ? 0 : 1;
        h.off_runs = 256;
        h.off_pairs = h.off_runs + (int)align256(plan.runs.size() * sizeof(RunDesc));
        h.off_slot = h.off_pairs + (int)align256(plan.pairs.size() * sizeof(int2));
        h.off_chg = h.off_slot + (int)align256((size_t)K * sizeof(int32_t));
        h.l0 = l0;
        const size_t total = h.off_chg + align256((size_t)K * sizeof(int32_t));
        KB2_CHECK(total <= stride, "kb2_self_moments_runs_plan_pack: plan exceeds the workspace");
        unsigned char* p = base + (size_t)b * stride;
        memset(p, 0, stride);
        memcpy(p, &h, sizeof(h));
        memcpy(p + h.off_runs, plan.runs.data(), plan.runs.size() * sizeof(RunDesc));
        memcpy(p + h.off_pairs, plan.pairs.data(), plan.pairs.size() * sizeof(int2));
        memcpy(p + h.off_slot, plan.slot.data(), (size_t)K * sizeof(int32_t));
        if (!plan.chg.empty()) memcpy(p + h.off_chg, plan.chg.data(), plan.chg.size() * sizeof(int32_t));
    }
    return 0;
}

// Launches the run kernels of a packed plan: `plan_host` is the host copy (its headers are read here), `plan_dev` the
// same bytes in device memory (uploaded by the caller, before this call in stream order), `counters` one zero-initialised
// 256-byte slot of device memory per block of kRunMaxLags lags (zeroed by this call).  sums[n_lags][2] is zeroed first.
extern "C" int kb2_self_moments_runs_planned(const double* dc, int n_atoms, int n_frames, int64_t pitch, int n_lags,
                                             int dim_mask, const void* plan_host, const void* plan_dev, void* counters,
                                             double* sums, void* stream) {
    KB2_CHECK(dc && plan_host && plan_dev && counters && sums, "kb2_self_moments_runs_planned: null pointer");
    KB2_CHECK(n_atoms > 0 && n_frames > 0 && n_lags > 0, "kb2_self_moments_runs_planned: empty problem");
    KB2_CHECK(pitch >= n_frames && pitch % 2 == 0, "kb2_self_moments_runs_planned: pitch must be even and >= n_frames");
    KB2_CHECK(dim_mask >= 1 && dim_mask <= 7, "kb2_self_moments_runs_planned: dim_mask must be in 1..7");
    cudaStream_t st = (cudaStream_t)stream;
    int comp[3] = {0, 0, 0}, ndim = 0;
    for (int c = 0; c < 3; ++c)
        if (dim_mask & (1 << c)) comp[ndim++] = c;
    const int sm_count = device_sm_count();
    const int kcap = n_lags > kRunMaxLags ? kRunMaxLags : n_lags;
    const size_t stride = runs_plan_bytes(kcap, n_frames);
    const int blocks = (n_lags + kRunMaxLags - 1) / kRunMaxLags;
    KB2_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_lags * 2 * sizeof(double), st));
    KB2_CUDA(cudaMemsetAsync(counters, 0, (size_t)blocks * 256, st));
    for (int b = 0; b < blocks; ++b) {
        const unsigned char* ph = (const unsigned char*)plan_host + (size_t)b * stride;
        const char* pd = (const char*)plan_dev + (size_t)b * stride;
        PlanHeader h;
        memcpy(&h, ph, sizeof(h));
        KB2_CHECK(h.K > 0 && h.K <= kRunMaxLags && h.l0 == b * kRunMaxLags && h.l0 + h.K <= n_lags,
                  "kb2_self_moments_runs_planned: the plan does not match the lag count");
        const long long n_work = (long long)h.n_pairs * n_atoms;
        KB2_CHECK(n_work < (1ll << 31) - (1 << 20), "kb2_self_moments_runs_planned: too many work items");
        // atoms per batch: their trajectories (3 * pitch * 8 B each) fit in a fraction of the L2 together
        const int batch = (int)std::max<long long>(1, std::min<long long>(n_atoms, kRunBatchBytes / (3 * pitch * 8)));
        const RunDesc* d_runs = (const RunDesc*)(pd + h.off_runs);
        const int2* d_pairs = (const int2*)(pd + h.off_pairs);
        const int32_t* d_slot = (const int32_t*)(pd + h.off_slot);
        const int32_t* d_chg = (const int32_t*)(pd + h.off_chg);
        unsigned int* counter = (unsigned int*)((char*)counters + (size_t)b * 256);
        int rc;
        if (ndim == 1)
            rc = launch_runs<1>(runs_warps(), h.drift != 0, sm_count, n_work, st, dc, n_atoms, n_frames, pitch, comp, d_runs, d_pairs,
                                h.n_pairs, batch, d_chg, h.K, counter, d_slot, sums + 2 * h.l0);
        else if (ndim == 2)
            rc = launch_runs<2>(runs_warps(), h.drift != 0, sm_count, n_work, st, dc, n_atoms, n_frames, pitch, comp, d_runs, d_pairs,
                                h.n_pairs, batch, d_chg, h.K, counter, d_slot, sums + 2 * h.l0);
        else
            rc = launch_runs<3>(runs_warps(), h.drift != 0, sm_count, n_work, st, dc, n_atoms, n_frames, pitch, comp, d_runs, d_pairs,
                                h.n_pairs, batch, d_chg, h.K, counter, d_slot, sums + 2 * h.l0);
        if (rc) return rc;
    }
    return 0;
}

// Convenience form: plan, upload (a pageable copy: it synchronises the stream) and launch.  `workspace` holds the
// device copy of the plan; its first 256 bytes per block double as the work counters.
extern "C" int kb2_self_moments_runs(const double* dc, int n_atoms, int n_frames, int64_t pitch,
                                     const int32_t* lags_host, int n_lags, int dim_mask, void* workspace, double* sums,
                                     void* stream) {
    KB2_CHECK(dc && lags_host && workspace && sums, "kb2_self_moments_runs: null pointer");
    const size_t bytes = kb2_self_moments_runs_workspace_bytes(n_lags, n_frames);
    std::vector<unsigned char> packed(bytes);
    if (int rc = kb2_self_moments_runs_plan_pack(lags_host, n_lags, n_frames, packed.data())) return rc;
    KB2_CUDA(cudaMemcpyAsync(workspace, packed.data(), bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    // the header of each block is host-only information: its 256 bytes serve as the block's work counter slot
    const int kcap = n_lags > kRunMaxLags ? kRunMaxLags : n_lags;
    const size_t stride = runs_plan_bytes(kcap, n_frames);
    const int blocks = (n_lags + kRunMaxLags - 1) / kRunMaxLags;
    KB2_CHECK(stride % 256 == 0, "kb2_self_moments_runs: internal layout error");
    // counters must be contiguous 256-byte slots: use a small array behind the packed plan when there are several blocks
    if (blocks == 1)
        return kb2_self_moments_runs_planned(dc, n_atoms, n_frames, pitch, n_lags, dim_mask, packed.data(), workspace, workspace,
                                             sums, stream);
    for (int b = 0; b < blocks; ++b) {
        const int l0 = b * kRunMaxLags, K = std::min(kRunMaxLags, n_lags - l0);
        // one block at a time: shift the views so that the block looks like a single-block plan starting at lag 0
        std::vector<unsigned char> one(packed.begin() + (size_t)b * stride, packed.begin() + (size_t)(b + 1) * stride);
        PlanHeader h;
        memcpy(&h, one.data(), sizeof(h));
        h.l0 = 0;
        memcpy(one.data(), &h, sizeof(h));
        char* wd = (char*)workspace + (size_t)b * stride;
        if (int rc = kb2_self_moments_runs_planned(dc, n_atoms, n_frames, pitch, K, dim_mask, one.data(), wd, wd, sums + 2 * l0,
                                                   stream))
            return rc;
    }
    return 0;
}

// The plan for a lag grid, for tests and diagnostics: fills run_k0/run_delta/run_len/run_n_chg (capacity max_runs)
// with the runs of >= 3 lags (n_chg = number of one-sample drift changes inside the run) and returns their number;
// *n_rest receives the number of lags that became runs of one.
extern "C" int kb2_self_moments_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* run_k0,
                                     int32_t* run_delta, int32_t* run_len, int32_t* run_n_chg, int max_runs, int* n_rest) {
    if (!lags_host || n_lags <= 0) return 0;
    RunPlan plan;
    plan_runs(lags_host, std::min(n_lags, kRunMaxLags), n_frames, plan);
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
            if (run_n_chg) run_n_chg[n] = rd.n_chg * rd.shift;
        }
        ++n;
    }
    if (n_rest) *n_rest = rest;
    return n;
}
