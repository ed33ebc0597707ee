// Stage 2, variant 3 (default): the lag-window kernel for lag grids with arithmetic structure.
//
// Replaces the per-dt loop of MSDBootstrap.__init__ (kinisi/diffusion.py:571-602) over the list Parser.get_disps
// builds (kinisi/parser.py:208-216): for every lag k, sum and sum of squares of |x[j+k] - x[j]|^2 over atoms and origins.
//
// kinisi's default grid np.linspace(min_dt, max_dt, n_steps, dtype=int) (kinisi/parser.py:150-168) is a sequence of
// arithmetic progressions (the spacing changes by one sample wherever the truncation carries).  A WINDOW is M <= 8
// consecutive lags of one progression, kappa + w*D (w < M).  For the origins of one residue class, j = t + a*D,
//      partner(a, w) = x[t + kappa + (a + w)*D]
// depends on a + w only.  A thread owns one residue t (a "column"), keeps the accumulators of the M lags and the M
// most recent partner samples in registers, and walks the origins DOWNWARDS, a = B, B-1, ..., 0 (B = the last origin
// whose first lag still has a partner): every step loads ONE origin x[t + a*D] and ONE partner x[t + kappa + a*D],
// and pairs the origin with the whole partner window: M terms (8 FP64 instructions each) for two 24-byte samples.
//
// Compared with the arithmetic-run kernel (k2_moments_runs.cu: M origins in registers, partners streamed, the
// accumulator of a lag complete after M terms and then folded across the warp):
//   * the accumulators stay in registers for the whole column; the cross-lane fold happens once per work item,
//   * walking downwards, the partner window fills by itself: the first M-1 steps have 1, 2, ... M-1 terms in a static
//     pattern, there is no tail, and no step is predicated per lane,
//   * every window is a pure progression (the planner cuts at the truncation carries), so there is no drift handling;
//     a lag that fits no progression is a window of one.
//
// All 32 lanes of a work item run the same number of steps.  Columns are therefore grouped by their number of
// origins before they are packed into strips of 32: residues t < t* have one origin more than residues t >= t*
// (two GROUPS per window), and long columns (small D: few residues, many origins each) are cut into SEGMENTS of the
// origin range, top / middle / bottom, where a lower segment first loads the M-1 partners above its range without
// forming terms.  Within a group the columns of a batch of atoms are numbered consecutively, (atom, segment, t), so
// strips are full whatever D is; a lane past the end repeats the last column with stride 0 and partner = origin:
// every term it forms is exactly zero.
//
// Samples travel global -> shared with 8-byte cp.async (each lane copies what it reads itself P steps later: no
// cross-lane synchronisation) and shared -> registers as one 16-byte load per component ({origin, partner}).
// Persistent CTAs, one per SM; every warp draws its own items from an atomic counter, atoms in L2-sized batches, the
// longest items of a batch first.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "k2_plan.h"

namespace kb2 {

constexpr int kWinM = 8;                          // lags per window (register budget: 10 registers per lag)
constexpr int kWinStashRow = 36;                  // double2 per row of the fold stash (32 lanes + pad)
constexpr int kWinStashBytes = kWinM * kWinStashRow * 16;  // per warp
constexpr int kWinMaxLags = 256;                  // lags per launch (CTA accumulators in shared memory)
constexpr int kWinMinLen = 3;                     // shortest progression taken as such
constexpr int kWinCandidates = 24;                // successors tried as the common difference
constexpr int kWinSingleStride = 64;              // D of a window of one lag (any value works: nothing is reused)
constexpr int kWinSegSteps = 128;                 // target number of origins per column segment
constexpr long long kWinBatchBytes = 24ll << 20;  // trajectory bytes of one atom batch (L2 is 126 MB)
constexpr int kWinMaxPairs = 3 * 2 * kWinMaxLags; // (window, group, segment class) triples per launch

struct WinDesc {
    int kappa, D, m, acc_off;  // lags kappa + w*D, w < m; accumulator index of lag w is acc_off + w
};
static_assert(sizeof(WinDesc) == 16, "WinDesc is read as one int4");

// Columns of one (window, group, segment class): `segc` segments per residue, segment g covers the origins
// a_top0 - g*seg_stride downwards, nsteps of them.
struct WinPair {
    int win, tbase, ncol, a_top0;
    int segc, seg_stride, nsteps, top;
};
static_assert(sizeof(WinPair) == 32, "WinPair is read as two int4");

// Register rings: origins of DU + 1 steps (RU slots), partners of the window of M plus DV steps in flight (RV slots).
// RV is a multiple of RU and also the number of steps per unrolled block, so that every register index is static.
template <int M, int RU_>
struct WinCfg {
    static constexpr int RU = RU_;                                 // 2 or 4
    static constexpr int RV = ((M + 2 + RU - 1) / RU) * RU;         // multiple of RU, >= M + 2
    static constexpr int DV = RV - M;                              // steps between the load of a partner and its first use
    static constexpr int DU = RU - 1;                              // the same for an origin
    static constexpr int U = RV;
};

// Per-lane state of a column walk.  Samples go global -> registers directly (LDG.64, ~2.6 L1 wavefronts per 256-byte
// row of a warp; the cp.async.ca 8-byte copies of the first version of this kernel cost ~10 each and saturated the
// L1 data pipe at a third of the FP64 peak): the origin of step s+1 and the partner of step s+DV are requested at
// step s.  The walk stops at origin index 0; the loads run ahead of it and are clamped there instead of being
// predicated (a clamped load re-reads a sample of the column; its value is never used).
template <int NDIM>
struct WinLane {
    const double* row[3];  // start of the trajectory of this lane's atom, per component
    int eu;                // frame index of the next origin sample to load
    int ew;                // frame index of the origin whose partner is loaded next
    int kap;               // partner = origin + kap frames (0 for a lane without a column)
    int stride;            // D, or 0 for a lane without a column

    __device__ __forceinline__ void load_origin(double (&dst)[3]) {
#pragma unroll
        for (int c = 0; c < NDIM; ++c) dst[c] = __ldg(row[c] + (unsigned)eu);
        eu = max(eu - stride, 0);
    }
    __device__ __forceinline__ void load_partner(double (&dst)[3]) {
        const int ev = ew + kap;
#pragma unroll
        for (int c = 0; c < NDIM; ++c) dst[c] = __ldg(row[c] + (unsigned)ev);
        ew = max(ew - stride, 0);
    }
};

template <int NDIM, int M, int RU>
struct WinRegs {
    double vb[WinCfg<M, RU>::RV][3];  // partner ring: the sample of step s lives in slot s mod RV
    double ub[RU][3];                 // origin of step s in slot s mod RU
    double A1[M], A2[M];          // sums of d^2 and d^4 of the window's lags
};

// One step at phase PH = s mod U: requests for the steps ahead, then NT terms (lags 0 .. NT-1; lag w pairs the origin
// with the partner loaded w steps earlier).
template <int NDIM, int M, int RU, int PH, int NT>
__device__ __forceinline__ void win_step(WinLane<NDIM>& ln, WinRegs<NDIM, M, RU>& r) {
    constexpr int RV = WinCfg<M, RU>::RV, DV = WinCfg<M, RU>::DV, DU = WinCfg<M, RU>::DU;
    ln.load_partner(r.vb[(PH + DV) % RV]);
    ln.load_origin(r.ub[(PH + DU) % RU]);
#pragma unroll
    for (int w = 0; w < NT; ++w) {
        const int sl = (PH - w + RV) % RV;
        double d2 = 0.0;
#pragma unroll
        for (int c = 0; c < NDIM; ++c) {
            const double d = r.vb[sl][c] - r.ub[PH % RU][c];
            d2 = fma(d, d, d2);
        }
        r.A1[w] += d2;
        r.A2[w] = fma(d2, d2, r.A2[w]);
    }
}

template <int NDIM, int M, int RU, int F>
struct WinPrologue {  // partners of steps 0 .. DV-1, origins of steps 0 .. DU-1
    __device__ __forceinline__ static void run(WinLane<NDIM>& ln, WinRegs<NDIM, M, RU>& r) {
        constexpr int DV = WinCfg<M, RU>::DV, DU = WinCfg<M, RU>::DU;
        if constexpr (F < (DV > DU ? DV : DU)) {  // (DV < DU for windows of 2 and 6 lags)
            if constexpr (F < DV) ln.load_partner(r.vb[F]);
            if constexpr (F < DU) ln.load_origin(r.ub[F]);
            WinPrologue<NDIM, M, RU, F + 1>::run(ln, r);
        }
    }
};

// The fill: steps 0 .. M-2.  TOP columns form 1, 2, ... M-1 terms; lower segments only load (the partners above their
// origin range).  Returns false when the column ended inside the fill.
template <int NDIM, int M, int RU, int R>
struct WinFill {
    __device__ __forceinline__ static bool run(WinLane<NDIM>& ln, WinRegs<NDIM, M, RU>& r, int total, bool top) {
        if constexpr (R < M - 1) {
            if (R >= total) return false;
            if (top)
                win_step<NDIM, M, RU, R % WinCfg<M, RU>::U, R + 1>(ln, r);
            else
                win_step<NDIM, M, RU, R % WinCfg<M, RU>::U, 0>(ln, r);
            return WinFill<NDIM, M, RU, R + 1>::run(ln, r, total, top);
        } else {
            return true;
        }
    }
};

template <int NDIM, int M, int RU, int UU>
struct WinMain {
    // steps of one unrolled block, the first at phase (M - 1) mod U; returns true when the column is done
    __device__ __forceinline__ static bool run(WinLane<NDIM>& ln, WinRegs<NDIM, M, RU>& r, int& left) {
        constexpr int U = WinCfg<M, RU>::U;
        if constexpr (UU < U) {
            win_step<NDIM, M, RU, (M - 1 + UU) % U, M>(ln, r);
            if (--left == 0) return true;
            return WinMain<NDIM, M, RU, UU + 1>::run(ln, r, left);
        } else {
            return false;
        }
    }
};

// Walks one column and folds the accumulators of the warp into the CTA's sums.
template <int NDIM, int M, int RU>
__device__ __forceinline__ void win_column(WinLane<NDIM>& ln, int total, bool top, int acc_off, double* acc,
                                           double2* stash, int lane) {
    WinRegs<NDIM, M, RU> r;
#pragma unroll
    for (int w = 0; w < M; ++w) {
        r.A1[w] = 0.0;
        r.A2[w] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < WinCfg<M, RU>::RV; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) r.vb[i][c] = 0.0;
    WinPrologue<NDIM, M, RU, 0>::run(ln, r);
    if (WinFill<NDIM, M, RU, 0>::run(ln, r, total, top)) {
        int left = total - (M - 1);
        if (left > 0) {
#pragma unroll 1
            while (!WinMain<NDIM, M, RU, 0>::run(ln, r, left)) {
            }
        }
    }
    // fold: rows of kWinStashRow double2 in the warp's stash
    __syncwarp();
#pragma unroll
    for (int w = 0; w < M; ++w) stash[w * kWinStashRow + lane] = make_double2(r.A1[w], r.A2[w]);
    __syncwarp();
    {
        const int g = lane >> 2, q = lane & 3;  // lag g of the window, quarter q of its row
        double t1 = 0.0, t2 = 0.0;
        if (g < M) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double2 v = stash[g * kWinStashRow + q + 4 * j];
                t1 += v.x;
                t2 += v.y;
            }
        }
        t1 += shfl_xor_f64(t1, 1);
        t2 += shfl_xor_f64(t2, 1);
        t1 += shfl_xor_f64(t1, 2);
        t2 += shfl_xor_f64(t2, 2);
        if (q == 0 && g < M) {
            atomicAdd(acc + 2 * (acc_off + g), t1);
            atomicAdd(acc + 2 * (acc_off + g) + 1, t2);
        }
    }
    __syncwarp();
}

template <int NDIM, int WARPS, int RU>
__global__ void __launch_bounds__(32 * WARPS, 1)
    self_moments_win_kernel(const double* __restrict__ dc, int n_atoms, int Tn, int64_t pitch, int c0, int c1, int c2,
                            const WinDesc* __restrict__ wins, const WinPair* __restrict__ pairs, int n_pairs, int batch,
                            const int32_t* __restrict__ acc_lag, int K, unsigned int* work_counter,
                            const int32_t* __restrict__ slot, double* __restrict__ out) {
    constexpr int NT = 32 * WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [WARPS] fold stashes | acc [K][2] | prefix_full [n_pairs + 1] | prefix_last [n_pairs + 1]
    double* const acc = reinterpret_cast<double*>(smem_raw + (size_t)WARPS * kWinStashBytes);
    int* const prefix_full = reinterpret_cast<int*>(acc + 2 * K);
    int* const prefix_last = prefix_full + (n_pairs + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int n_batches = (n_atoms + batch - 1) / batch;
    const int nb_last = n_atoms - (n_batches - 1) * batch;
#pragma unroll 1
    for (int i = tid; i < 2 * K; i += NT) acc[i] = 0.0;
    // strips per pair for a full batch and for the last batch; serial prefix sums (a few hundred entries at most)
#pragma unroll 1
    for (int p = tid; p < n_pairs; p += NT) {
        const int4 b = __ldg(reinterpret_cast<const int4*>(pairs) + 2 * p);
        const int4 e = __ldg(reinterpret_cast<const int4*>(pairs) + 2 * p + 1);
        const long long per_atom = (long long)b.z * e.x;  // ncol * segc
        prefix_full[p + 1] = (int)((per_atom * batch + 31) / 32);
        prefix_last[p + 1] = (int)((per_atom * nb_last + 31) / 32);
    }
    __syncthreads();
    if (warp == 0) {
        // inclusive scan of both arrays by one warp, 32 entries at a time
        int carry_f = 0, carry_l = 0;
#pragma unroll 1
        for (int base = 0; base < n_pairs; base += 32) {
            const int idx = base + lane;
            int vf = idx < n_pairs ? prefix_full[idx + 1] : 0;
            int vl = idx < n_pairs ? prefix_last[idx + 1] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int uf = __shfl_up_sync(0xffffffffu, vf, d);
                const int ul = __shfl_up_sync(0xffffffffu, vl, d);
                if (lane >= d) {
                    vf += uf;
                    vl += ul;
                }
            }
            vf += carry_f;
            vl += carry_l;
            if (idx < n_pairs) {
                prefix_full[idx + 1] = vf;
                prefix_last[idx + 1] = vl;
            }
            carry_f = __shfl_sync(0xffffffffu, vf, 31);
            carry_l = __shfl_sync(0xffffffffu, vl, 31);
        }
        if (lane == 0) {
            prefix_full[0] = 0;
            prefix_last[0] = 0;
        }
    }
    __syncthreads();
    const int items_full = prefix_full[n_pairs], items_last = prefix_last[n_pairs];
    const long long n_front = (long long)(n_batches - 1) * items_full;
    const long long n_work = n_front + items_last;
    const int comp[3] = {c0, c1, c2};

    // the extra first observation dc[a][k-1] of every lag (kinisi/parser.py:209): units of (lag, 32 atoms)
    {
        const int gw = blockIdx.x * WARPS + warp, n_gw = gridDim.x * WARPS;
        const int ablocks = (n_atoms + 31) / 32;
#pragma unroll 1
        for (long long unit = gw; unit < (long long)K * ablocks; unit += n_gw) {
            const int i = (int)(unit % K), ab = (int)(unit / K);
            const int atom = ab * 32 + lane;
            const int k = __ldg(acc_lag + i);
            double d2 = 0.0;
            if (atom < n_atoms) {
#pragma unroll
                for (int c = 0; c < NDIM; ++c) {
                    const double d = __ldg(dc + ((size_t)atom * 3 + comp[c]) * pitch + (k - 1));
                    d2 = fma(d, d, d2);
                }
            }
            const double s1 = warp_sum(d2), s2 = warp_sum(d2 * d2);
            if (lane == 0) {
                atomicAdd(acc + 2 * i, s1);
                atomicAdd(acc + 2 * i + 1, s2);
            }
        }
    }

    WinLane<NDIM> ln;
    double2* const stash = reinterpret_cast<double2*>(smem_raw + (size_t)warp * kWinStashBytes);

    while (true) {
        long long wk = 0;
        if (lane == 0) wk = (long long)atomicAdd(work_counter, 1u);
        wk = __shfl_sync(0xffffffffu, wk, 0);
        if (wk >= n_work) break;
        int bi, r;
        const int* prefix;
        if (wk < n_front) {
            bi = (int)(wk / items_full);
            r = (int)(wk - (long long)bi * items_full);
            prefix = prefix_full;
        } else {
            bi = n_batches - 1;
            r = (int)(wk - n_front);
            prefix = prefix_last;
        }
        const int nb = bi == n_batches - 1 ? nb_last : batch;
        // pair p with prefix[p] <= r < prefix[p + 1]
        int p = 0;
#pragma unroll 1
        for (int base = 0; base < n_pairs; base += 32) {
            const int idx = base + lane;
            const bool below = idx < n_pairs && prefix[idx + 1] <= r;
            const int cnt = __popc(__ballot_sync(0xffffffffu, below));
            p += cnt;
            if (cnt < 32) break;
        }
        const int strip = r - prefix[p];
        const int4 pb = __ldg(reinterpret_cast<const int4*>(pairs) + 2 * p);
        const int4 pe = __ldg(reinterpret_cast<const int4*>(pairs) + 2 * p + 1);
        const int4 wd = __ldg(reinterpret_cast<const int4*>(wins) + pb.x);
        const int kappa = wd.x, D = wd.y, m = wd.z, acc_off = wd.w;
        const int tbase = pb.y, ncol = pb.z, a_top0 = pb.w, segc = pe.x, seg_stride = pe.y, nsteps = pe.z;
        const bool top = pe.w != 0;

        // column of this lane: (atom of the batch, segment, residue), residue fastest
        const long long per_atom = (long long)ncol * segc;
        const long long ncols = per_atom * nb;
        long long g = (long long)strip * 32 + lane;
        const bool active = g < ncols;
        if (!active) g = ncols - 1;
        const int al = (int)(g / per_atom);
        const int rem = (int)(g - (long long)al * per_atom);
        const int sg = rem / ncol;
        const int t = tbase + (rem - sg * ncol);
        const int atom = bi * batch + al;
        const int a_top = a_top0 - sg * seg_stride;
        const int a_start = top ? a_top : a_top + (m - 1);
        const int total = nsteps + (top ? 0 : m - 1);
        ln.stride = active ? D : 0;
        ln.kap = active ? kappa : 0;
        ln.eu = ln.ew = t + a_start * D;  // <= n_frames - 1 - kappa
#pragma unroll
        for (int c = 0; c < NDIM; ++c) {
            ln.row[c] = dc + ((size_t)atom * 3 + comp[c]) * pitch;
            asm volatile("" : "+l"(ln.row[c]));  // keep the pointer in registers: the compiler otherwise recomputes it per load
        }
        switch (m) {
            case 8: win_column<NDIM, 8, RU>(ln, total, top, acc_off, acc, stash, lane); break;
            case 7: win_column<NDIM, 7, RU>(ln, total, top, acc_off, acc, stash, lane); break;
            case 6: win_column<NDIM, 6, RU>(ln, total, top, acc_off, acc, stash, lane); break;
            case 5: win_column<NDIM, 5, RU>(ln, total, top, acc_off, acc, stash, lane); break;
            case 4: win_column<NDIM, 4, RU>(ln, total, top, acc_off, acc, stash, lane); break;
            case 3: win_column<NDIM, 3, RU>(ln, total, top, acc_off, acc, stash, lane); break;
            case 2: win_column<NDIM, 2, RU>(ln, total, top, acc_off, acc, stash, lane); break;
            default: win_column<NDIM, 1, RU>(ln, total, top, acc_off, acc, stash, lane); break;
        }
    }
    __syncthreads();
    // one FP64 reduction per lag, moment and CTA straight into the caller's array (zeroed by the host call)
#pragma unroll 1
    for (int i = tid; i < 2 * K; i += NT) atomicAdd(out + 2 * slot[i >> 1] + (i & 1), acc[i]);
}

// ---- host-side planning -------------------------------------------------------------------------
struct WinPlan {
    std::vector<WinDesc> wins;
    std::vector<WinPair> pairs;      // most expensive items first
    std::vector<int32_t> slot;       // accumulator index -> position in the caller's lag array
    std::vector<int32_t> acc_lag;    // accumulator index -> lag
};

static void plan_windows(const int32_t* lags, int K, int Tn, WinPlan& plan) {
    const std::vector<LagSegment> segs = cover_by_progressions(lags, K, Tn, kWinMinLen, kWinCandidates);
    for (const LagSegment& sg : segs) {
        // cut the progression into windows of nearly equal size <= kWinM
        const int nw = (sg.len + kWinM - 1) / kWinM;
        const int base = sg.len / nw, extra = sg.len % nw;
        int at = 0;
        for (int w = 0; w < nw; ++w) {
            const int m = base + (w < extra ? 1 : 0);
            WinDesc wd{sg.k0 + at * sg.d, sg.len == 1 ? kWinSingleStride : sg.d, m, (int)plan.slot.size()};
            for (int i = 0; i < m; ++i) {
                plan.slot.push_back(sg.slots[at + i]);
                plan.acc_lag.push_back(sg.k0 + (at + i) * sg.d);
            }
            plan.wins.push_back(wd);
            at += m;
        }
    }
    struct Item {
        WinPair pr;
        long long cost;
    };
    std::vector<Item> items;
    for (size_t wi = 0; wi < plan.wins.size(); ++wi) {
        const WinDesc& wd = plan.wins[wi];
        const long long span = (long long)Tn - 1 - wd.kappa;  // largest origin index with a partner at the first lag
        if (span < 0) continue;                               // lag == n_frames: only the first-observation term
        const int b_hi = (int)(span / wd.D);
        const int tstar = (int)(span - (long long)b_hi * wd.D) + 1;  // residues t < tstar have origins a <= b_hi
        for (int grp = 0; grp < 2; ++grp) {
            const int tbase = grp == 0 ? 0 : tstar;
            const int ncol = grp == 0 ? tstar : wd.D - tstar;
            const int B = grp == 0 ? b_hi : b_hi - 1;
            if (ncol <= 0 || B < 0) continue;
            const int n = B + 1;  // origins per column
            const int nseg0 = std::max(1, (n + kWinSegSteps / 2) / kWinSegSteps);
            const int ls = (n + nseg0 - 1) / nseg0;
            const int nseg = (n + ls - 1) / ls;  // (nseg - 1) * ls < n: the bottom segment is never empty
            auto add = [&](int a_top0, int segc, int nsteps, int top) {
                if (segc <= 0 || nsteps <= 0) return;
                WinPair pr{(int)wi, tbase, ncol, a_top0, segc, ls, nsteps, top};
                const long long steps = nsteps + (top ? 0 : wd.m - 1);
                items.push_back({pr, steps * (16ll * wd.m + 24)});
            };
            // segments from the top: segment g covers origins B - g*ls downwards
            if (nseg == 1) {
                add(B, 1, n, 1);
            } else {
                add(B, 1, ls, 1);
                add(B - ls, nseg - 2, ls, 0);
                add(B - (nseg - 1) * ls, 1, n - (nseg - 1) * ls, 0);
            }
        }
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.cost > y.cost; });
    for (const Item& it : items) plan.pairs.push_back(it.pr);
}

// Packed plan of one launch (<= kWinMaxLags lags): header (256 B) | windows | pairs | slots | lags by accumulator.
struct WinHeader {
    int K, n_wins, n_pairs, pad0;
    int off_wins, off_pairs, off_slot, off_lag;
    int l0, pad[7];
};
static_assert(sizeof(WinHeader) <= 256, "the header has 256 bytes");

static size_t win_plan_bytes(int k) {
    return 256 + align256((size_t)k * sizeof(WinDesc)) + align256((size_t)6 * k * sizeof(WinPair)) +
           2 * align256((size_t)k * sizeof(int32_t));
}

template <int NDIM, int WARPS, int RU>
static int launch_win(int sm_count, cudaStream_t st, const double* dc, int n_atoms, int Tn, int64_t pitch,
                      const int comp[3], const WinDesc* wins, const WinPair* pairs, int n_pairs, int batch,
                      const int32_t* acc_lag, int K, unsigned int* counter, const int32_t* slot, double* out) {
    const size_t smem = (size_t)WARPS * kWinStashBytes + 16 * (size_t)K + 2 * sizeof(int) * (size_t)(n_pairs + 1);
    KB2_CUDA(cudaFuncSetAttribute(self_moments_win_kernel<NDIM, WARPS, RU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    self_moments_win_kernel<NDIM, WARPS, RU><<<sm_count, 32 * WARPS, smem, st>>>(dc, n_atoms, Tn, pitch, comp[0], comp[1], comp[2],
                                                                           wins, pairs, n_pairs, batch, acc_lag, K, counter,
                                                                           slot, out);
    KB2_LAUNCH_CHECK();
    return 0;
}

static int win_warps() {
    static int w = [] {
        const char* e = getenv("KB2_WIN_WARPS");
        const int v = e ? atoi(e) : 0;
        return (v == 122 || v == 12 || v == 16) ? v : 12;
    }();
    return w;
}

static int win_check_lags(const char* fn, const int32_t* lags_host, int n_lags, int n_frames) {
    KB2_CHECK(lags_host && n_lags > 0 && n_frames > 0, "%s: empty problem", fn);
    for (int i = 0; i < n_lags; ++i) {
        KB2_CHECK(lags_host[i] >= 1 && lags_host[i] <= n_frames, "%s: lag %d out of range", fn, i);
        KB2_CHECK(i == 0 || lags_host[i] >= lags_host[i - 1], "%s: lags must ascend", fn);
    }
    return 0;
}

}  // namespace kb2

using namespace kb2;

extern "C" size_t kb2_self_moments_win_workspace_bytes(int n_lags) {
    const int n = n_lags < 1 ? 1 : n_lags;
    const int blocks = (n + kWinMaxLags - 1) / kWinMaxLags;
    return (size_t)blocks * win_plan_bytes(n > kWinMaxLags ? kWinMaxLags : n);
}

extern "C" int kb2_self_moments_win_blocks(int n_lags) { return n_lags < 1 ? 1 : (n_lags + kWinMaxLags - 1) / kWinMaxLags; }

extern "C" int kb2_self_moments_win_plan_pack(const int32_t* lags_host, int n_lags, int n_frames, void* plan_host) {
    if (int rc = win_check_lags("kb2_self_moments_win_plan_pack", lags_host, n_lags, n_frames)) return rc;
    KB2_CHECK(plan_host, "kb2_self_moments_win_plan_pack: null pointer");
    const int kcap = n_lags > kWinMaxLags ? kWinMaxLags : n_lags;
    const size_t stride = win_plan_bytes(kcap);
    unsigned char* base = (unsigned char*)plan_host;
    int b = 0;
    for (int l0 = 0; l0 < n_lags; l0 += kWinMaxLags, ++b) {
        const int K = std::min(kWinMaxLags, n_lags - l0);
        WinPlan plan;
        plan_windows(lags_host + l0, K, n_frames, plan);
        KB2_CHECK((int)plan.slot.size() == K && (int)plan.wins.size() <= K && plan.pairs.size() <= (size_t)6 * K,
                  "kb2_self_moments_win_plan_pack: internal planning error");
        WinHeader h{};
        h.K = K;
        h.n_wins = (int)plan.wins.size();
        h.n_pairs = (int)plan.pairs.size();
        h.off_wins = 256;
        h.off_pairs = h.off_wins + (int)align256((size_t)K * sizeof(WinDesc));
        h.off_slot = h.off_pairs + (int)align256((size_t)6 * K * sizeof(WinPair));
        h.off_lag = h.off_slot + (int)align256((size_t)K * sizeof(int32_t));
        h.l0 = l0;
        unsigned char* p = base + (size_t)b * stride;
        memset(p, 0, stride);
        memcpy(p, &h, sizeof(h));
        memcpy(p + h.off_wins, plan.wins.data(), plan.wins.size() * sizeof(WinDesc));
        if (!plan.pairs.empty()) memcpy(p + h.off_pairs, plan.pairs.data(), plan.pairs.size() * sizeof(WinPair));
        memcpy(p + h.off_slot, plan.slot.data(), (size_t)K * sizeof(int32_t));
        memcpy(p + h.off_lag, plan.acc_lag.data(), (size_t)K * sizeof(int32_t));
    }
    return 0;
}

extern "C" int kb2_self_moments_win_planned(const double* dc, int n_atoms, int n_frames, int64_t pitch, int n_lags,
                                            int dim_mask, const void* plan_host, const void* plan_dev, void* counters,
                                            double* sums, void* stream) {
    KB2_CHECK(dc && plan_host && plan_dev && counters && sums, "kb2_self_moments_win_planned: null pointer");
    KB2_CHECK(n_atoms > 0 && n_frames > 0 && n_lags > 0, "kb2_self_moments_win_planned: empty problem");
    KB2_CHECK(pitch >= n_frames, "kb2_self_moments_win_planned: pitch must be >= n_frames");
    KB2_CHECK(dim_mask >= 1 && dim_mask <= 7, "kb2_self_moments_win_planned: dim_mask must be in 1..7");
    cudaStream_t st = (cudaStream_t)stream;
    int comp[3] = {0, 0, 0}, ndim = 0;
    for (int c = 0; c < 3; ++c)
        if (dim_mask & (1 << c)) comp[ndim++] = c;
    const int sm_count = device_sm_count();
    KB2_CHECK(sm_count > 0, "kb2_self_moments_win_planned: no CUDA device");
    const int kcap = n_lags > kWinMaxLags ? kWinMaxLags : n_lags;
    const size_t stride = win_plan_bytes(kcap);
    const int blocks = (n_lags + kWinMaxLags - 1) / kWinMaxLags;
    KB2_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_lags * 2 * sizeof(double), st));
    KB2_CUDA(cudaMemsetAsync(counters, 0, (size_t)blocks * 256, st));
    const int batch = (int)std::max<long long>(1, std::min<long long>(n_atoms, kWinBatchBytes / (3 * pitch * 8)));
    for (int b = 0; b < blocks; ++b) {
        const unsigned char* ph = (const unsigned char*)plan_host + (size_t)b * stride;
        const char* pd = (const char*)plan_dev + (size_t)b * stride;
        WinHeader h;
        memcpy(&h, ph, sizeof(h));
        KB2_CHECK(h.K > 0 && h.K <= kWinMaxLags && h.l0 == b * kWinMaxLags && h.l0 + h.K <= n_lags,
                  "kb2_self_moments_win_planned: the plan does not match the lag count");
        const WinDesc* d_wins = (const WinDesc*)(pd + h.off_wins);
        const WinPair* d_pairs = (const WinPair*)(pd + h.off_pairs);
        const int32_t* d_slot = (const int32_t*)(pd + h.off_slot);
        const int32_t* d_lag = (const int32_t*)(pd + h.off_lag);
        unsigned int* counter = (unsigned int*)((char*)counters + (size_t)b * 256);
        const int warps = win_warps();
        int rc = 0;
#define KB2_WIN(ND, W, RU)                                                                                                  \
    rc = launch_win<ND, W, RU>(sm_count, st, dc, n_atoms, n_frames, pitch, comp, d_wins, d_pairs, h.n_pairs, batch, d_lag, \
                               h.K, counter, d_slot, sums + 2 * h.l0)
        // experiments (KB2_WIN_WARPS): 16 warps x 2 origin slots (128 registers) or 12 warps x 4 origin slots (168)
        if (ndim == 3) {
            if (warps == 16)
                KB2_WIN(3, 16, 2);
            else if (warps == 122)
                KB2_WIN(3, 12, 2);
            else
                KB2_WIN(3, 12, 4);
        } else if (ndim == 2) {
            KB2_WIN(2, 12, 4);
        } else {
            KB2_WIN(1, 12, 4);
        }
#undef KB2_WIN
        if (rc) return rc;
    }
    return 0;
}

// The windows of the plan for (the first kWinMaxLags lags of) a grid, for tests and diagnostics: fills up to max_wins
// entries of win_kappa / win_delta / win_len (host arrays, may be NULL) and returns the number of windows.
extern "C" int kb2_self_moments_win_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* win_kappa,
                                         int32_t* win_delta, int32_t* win_len, int max_wins) {
    if (!lags_host || n_lags <= 0) return 0;
    WinPlan plan;
    plan_windows(lags_host, std::min(n_lags, kWinMaxLags), n_frames, plan);
    int n = 0;
    for (const WinDesc& wd : plan.wins) {
        if (n < max_wins) {
            if (win_kappa) win_kappa[n] = wd.kappa;
            if (win_delta) win_delta[n] = wd.D;
            if (win_len) win_len[n] = wd.m;
        }
        ++n;
    }
    return n;
}
