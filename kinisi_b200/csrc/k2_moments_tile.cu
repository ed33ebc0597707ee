// Stage 2, variant 4 (default): the tiled lag-window kernel.  Trajectory tiles are staged in shared memory by the TMA
// engine (cp.async.bulk + mbarrier), the windows of a tile are walked by the consumer warps out of shared memory.
//
// Replaces the per-dt loop of MSDBootstrap.__init__ (kinisi/diffusion.py:571-602) over the list Parser.get_disps
// builds (kinisi/parser.py:208-216): for every lag k, sum and sum of squares of |x[j+k] - x[j]|^2 over atoms and origins.
//
// Geometry.  Pick a row length D (>= 32) and read one component of one atom's trajectory as a matrix with rows of
// D frames: frame f = a*D + t is (row a, column t).  Any lag is k = I*D + r (0 <= r < D): the partner of the origin
// (a, t) is the frame (a + I)*D + (t + r), i.e. row a + I, column t + r, counted linearly (a column index >= D simply
// runs into the next row).  A WINDOW is m <= 8 lags with consecutive I and one r -- kinisi's default grid
// np.linspace(min_dt, max_dt, n_steps, dtype=int) (kinisi/parser.py:150-168) is a sequence of arithmetic progressions,
// and D is a multiple of their common difference (the multiple splits one progression into interleaved ones; it is
// chosen so that D fills whole strips of 32 columns).  A thread owns one column t, keeps the accumulators of the m lags
// and the m most recent partner samples in registers and walks the origin rows DOWNWARDS: every step loads one origin
// sample and one partner sample and forms m terms (8 FP64 instructions each).  Walking downwards the partner window
// fills by itself at the top of a column (1, 2, ... m-1 terms, a static pattern): no tail, no per-lane predicate except
// on the one partner sample of the top row that lies past the end of the trajectory for some columns.
//
// Tiles.  The lag-window kernel of the previous version (k2_moments_win.cu) loaded both samples of a step from global
// memory into registers: 48 bytes per 8 terms from the L2, and the six scoreboards of a warp alias on the LDGs in flight,
// so a step waited for the loads issued the step before (ncu: 4.3 long-scoreboard stall cycles per issue, FP64 pipe 49 %,
// profiles/r02_k2_win_ncu_summary.txt).  Here a TILE is what a group of windows needs for a block of origin rows of one
// strip of 32 columns of one atom: the origin rows (32 columns wide) and the partner rows (40 wide: the windows of a tile
// may differ in r by up to 8).  One producer warp copies the rows of a tile into one of kTileBufs shared-memory buffers
// with cp.async.bulk (global -> shared, completion on the buffer's `full` mbarrier); the consumer warps take the
// (window, row block) UNITS of the tile from a counter in shared memory, walk them with LDS only, and arrive on the
// buffer's `empty` mbarrier when they leave the tile.  Tiles are drawn from an atomic counter, the most expensive first,
// atoms in L2-sized batches.  L2 -> SM traffic falls to ~2-3 bytes per term and is asynchronous, a whole tile ahead.
//
// Lags that fit no progression are windows of one (two samples per term: shared-memory bound, about the rate of the
// direct kernel).  The extra first observation dc[a][k-1] of every lag (kinisi/parser.py:209) is added by a pre-pass.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "common.cuh"
#include "k2_plan.h"

namespace kb2 {

constexpr int kTileM = 8;                    // lags per window
constexpr int kTileWp = 40;                  // columns of a partner row = component stride within any row
constexpr int kTileRowP = 3 * kTileWp;       // doubles per row (three components)
constexpr int kTileRowO = kTileRowP;         // origin rows have the layout of partner rows (a SHARED tile has one row set)
constexpr int kTileMaxBufs = 3;              // tile buffers in flight: 3 x 64 KB or 2 x 96 KB (chosen by the planner)
constexpr int kTileSmemBufBytes = 192 * 1024;  // shared memory of all buffers
constexpr int kTileGuardBytes = 4096;        // shared memory in front of the first buffer (look-ahead reads run below a tile)
constexpr int kTileCW = 14;                  // consumer warps
constexpr int kTileMaxLags = 256;            // lags per launch (CTA accumulators in shared memory)
constexpr int kTileRA = 64;                  // most origin rows per unit (sweep: tools/k2tile_sweep.sh, profiles/r02_k2_tile_sweeps.txt)
constexpr int kTileMaxUnits = 96;            // units per tile
constexpr int kTileMinLen = 3;               // shortest progression taken as such
constexpr int kTileCandidates = 24;          // successors tried as the common difference
constexpr int kTileSingleD = 64;             // D of a window of one lag (nothing is reused across its lags)
constexpr long long kTileBatchBytes = 24ll << 20;  // trajectory bytes of one atom batch (L2 is 126 MB)

// One (window, row block): lags (I0 + w)*D + r, w < m; origins a_top, a_top - 1, ... (nsteps of them).
struct TileUnit {
    int I0, r, m, acc_off;
    int a_top, nsteps, top, pad;
};
static_assert(sizeof(TileUnit) == 32, "TileUnit is read as two int4");

// One tile template (per strip and atom): origin rows [oLo, oLo + nO), partner rows [pLo, pLo + nP) starting at the
// column offset cP0 (even), units [unit0, unit0 + nunits).
struct TileTpl {
    int D, oLo, nO, pLo;
    int nP, cP0, unit0, nunits;
    int nstrips, strip_prefix, spt, pack_w;  // strips of 32 columns; spt of them side by side in one tile (nstrips here =
                                             // number of strip GROUPS); strip_prefix = groups of the templates before this one;
                                             // pack_w > 0: the last strip has pack_w columns and its tile takes that strip of
                                             // SEVERAL atoms side by side (see tile_pack)
};
static_assert(sizeof(TileTpl) == 48, "TileTpl is read as three int4");

struct TileDesc {  // what the producer publishes with a filled buffer
    int stop, t0, D, oLo;
    int pLo, cP0, unit0, nunits;
    int nO, atom, ns, sub_doubles;  // ns strips in this tile, one every sub_doubles doubles of the buffer
    int pack_w, pack_n, pack_wp, pad;  // pack_w > 0: pack_n atoms' pieces of pack_w columns, one every pack_wp columns of a row
};

// A ragged last strip (D mod 32 = w columns) leaves 32 - w lanes of every warp that walks it idle: 22 of 32 at kinisi's
// default grid on 20 000 frames (D = 202 = 6 * 32 + 10, a tenth of all lane-steps).  A piece of such a strip needs
// w + 8 columns of a 40-column row (8: the room for the residues r of its windows), so the tile of the last strip holds
// that strip of P = 40 / wp consecutive ATOMS side by side, wp = w + 8 rounded up to even (16-byte copies); lane l walks
// column l mod w of atom l / w.  All atoms share the template, so the units and their control flow are those of one atom.
__host__ __device__ inline int tile_pack_wp(int w) { return (w + 8 + 1) & ~1; }
__host__ __device__ inline int tile_pack_atoms(int w) { return w > 0 && w < 32 ? kTileWp / tile_pack_wp(w) : 1; }

namespace tile {

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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "KB2_TILE_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KB2_TILE_DONE;\n\t"
        "bra KB2_TILE_WAIT;\n\t"
        "KB2_TILE_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completion (bytes) is signalled on `bar`.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only), and the arrival on an mbarrier that fires when all earlier
// cp.async of the executing thread have landed (the pending count of the barrier is not incremented).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Register rings: the partner of step s lives in slot s mod RV, the origin of step s in slot s mod 2.  RV is even (so
// that one unrolled block of RV steps keeps every register index static) and >= M + 1: a partner is loaded DV = RV - M
// (1 or 2) steps before its first use, an origin one step before its use.  Shared-memory latency is ~30 cycles against
// >= 16 m cycles of FP64 issue per step.
template <int M>
struct Cfg {
    static constexpr int RV = ((M + 2) / 2) * 2;
    static constexpr int DV = RV - M;
    static constexpr int U = RV;
};

template <int NDIM, int M>
struct Regs {
    double vb[Cfg<M>::RV][3];
    double ub[2][3];
    double A1[M], A2[M];
};

// Sample pointers of a column walk.  With an odd row length D the rows of a tile start alternately at an even and an
// odd frame, and a row is copied from the even frame at or below its start (16-byte copies): the samples of the rows of
// one parity sit one column further right.  o[0] / p[0] serve the even steps counted from the step the pointers refer
// to, o[1] / p[1] the odd ones (all equal for an even D).
struct Ptrs {
    const double* o[2];
    const double* p[2];
    __device__ __forceinline__ void advance(int steps) {  // `steps` rows down; an odd count swaps the parities
        const double* const o0 = o[0] - steps * kTileRowO;
        const double* const o1 = o[1] - steps * kTileRowO;
        const double* const p0 = p[0] - steps * kTileRowP;
        const double* const p1 = p[1] - steps * kTileRowP;
        const bool swap = steps & 1;
        o[0] = swap ? o1 : o0;
        o[1] = swap ? o0 : o1;
        p[0] = swap ? p1 : p0;
        p[1] = swap ? p0 : p1;
    }
};

// One step.  The pointers refer to step 0 of the current block; S = the step's index relative to them (compile time),
// PH = absolute step index mod U.  Loads the samples of the steps ahead, then forms NT terms (lags 0 .. NT-1: lag w
// pairs the origin with the partner loaded w steps earlier).  DIAG: the last term uses the partner of the top row, which
// lies past the end of the trajectory for the lanes with `kill` set.
template <int NDIM, int M, int S, int PH, int NT, bool DIAG>
__device__ __forceinline__ void step(const Ptrs& q, Regs<NDIM, M>& r, bool kill) {
    constexpr int RV = Cfg<M>::RV, DV = Cfg<M>::DV;
#pragma unroll
    for (int c = 0; c < NDIM; ++c) r.vb[(PH + DV) % RV][c] = q.p[(S + DV) & 1][-(S + DV) * kTileRowP + c * kTileWp];
#pragma unroll
    for (int c = 0; c < NDIM; ++c) r.ub[(PH + 1) % 2][c] = q.o[(S + 1) & 1][-(S + 1) * kTileRowO + c * kTileWp];
#pragma unroll
    for (int w = 0; w < NT; ++w) {
        const int sl = (PH - w + RV) % RV;
        double d2 = 0.0;
#pragma unroll
        for (int c = 0; c < NDIM; ++c) {
            const double d = r.vb[sl][c] - r.ub[PH % 2][c];
            d2 = fma(d, d, d2);
        }
        if (DIAG && w == NT - 1) d2 = kill ? 0.0 : d2;
        r.A1[w] += d2;
        r.A2[w] = fma(d2, d2, r.A2[w]);
    }
}

template <int NDIM, int M, int F>
struct Prologue {  // partners of steps 0 .. DV-1, origin of step 0
    __device__ __forceinline__ static void run(const Ptrs& q, Regs<NDIM, M>& r) {
        if constexpr (F < Cfg<M>::DV) {
#pragma unroll
            for (int c = 0; c < NDIM; ++c) r.vb[F][c] = q.p[F & 1][-F * kTileRowP + c * kTileWp];
            Prologue<NDIM, M, F + 1>::run(q, r);
        } else {
#pragma unroll
            for (int c = 0; c < NDIM; ++c) r.ub[0][c] = q.o[0][c * kTileWp];
        }
    }
};

// The fill: steps 0 .. M-1.  TOP columns form 1, 2, ... M terms, the last of each masked for the short lanes; lower
// blocks only load during the first M-1 steps (the partners above their origin range) and form all M terms at step M-1.
// Returns false when the column ended inside the fill.
template <int NDIM, int M, int R>
struct Fill {
    __device__ __forceinline__ static bool run(const Ptrs& q, Regs<NDIM, M>& r, int total, bool top, bool kill) {
        if constexpr (R < M) {
            if (R >= total) return false;
            if (top)
                step<NDIM, M, R, R % Cfg<M>::U, R + 1, true>(q, r, kill);
            else
                step<NDIM, M, R, R % Cfg<M>::U, (R == M - 1 ? M : 0), false>(q, r, false);
            return Fill<NDIM, M, R + 1>::run(q, r, total, top, kill);
        } else {
            return true;
        }
    }
};

template <int NDIM, int M, int UU>
struct Main {
    // steps of one unrolled block, the first at phase M mod U; returns true when the column is done
    __device__ __forceinline__ static bool run(const Ptrs& q, Regs<NDIM, M>& r, int& left) {
        constexpr int U = Cfg<M>::U;
        if constexpr (UU < U) {
            step<NDIM, M, UU, (M + UU) % U, M, false>(q, r, false);
            if (--left == 0) return true;
            return Main<NDIM, M, UU + 1>::run(q, r, left);
        } else {
            return false;
        }
    }
};

// Sums 2M values over the 32 lanes with a halving butterfly: after the exchange with lane ^ 16 a lane keeps half of its
// values, after lane ^ 8 a quarter, ...; the lanes whose index has bit 0 clear end up with one complete sum each.
template <int N>
__device__ __forceinline__ void butterfly(double (&v)[16], int lane, int bit) {
    if constexpr (N >= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double send = up ? v[i] : v[i + N];
            const double keep = up ? v[i + N] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
        butterfly<N / 2>(v, lane, bit >> 1);
    }
}

// Walks one column and adds the sums of the warp to the CTA's accumulators.
template <int NDIM, int M>
__device__ __forceinline__ void column(Ptrs q, int total, bool top, bool kill, bool active, int acc_off, double* acc,
                                       int lane) {
    static_assert(Cfg<M>::U % 2 == 0, "a block of the main loop keeps the row parities");
    Regs<NDIM, M> r;
#pragma unroll
    for (int w = 0; w < M; ++w) {
        r.A1[w] = 0.0;
        r.A2[w] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < Cfg<M>::RV; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) r.vb[i][c] = 0.0;
    Prologue<NDIM, M, 0>::run(q, r);
    if (Fill<NDIM, M, 0>::run(q, r, total, top, kill)) {
        int left = total - M;
        q.advance(M);
        if (left > 0) {
#pragma unroll 1
            while (!Main<NDIM, M, 0>::run(q, r, left)) q.advance(Cfg<M>::U);
        }
    }
    // a lane without a column (ragged last strip) has read its neighbour's data: its sums are dropped
    double v[16];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        v[w] = (w < M && active) ? r.A1[w] : 0.0;
        v[8 + w] = (w < M && active) ? r.A2[w] : 0.0;
    }
    butterfly<8>(v, lane, 16);
    const double s = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
    // value index held by this lane: bit 4 of the lane selects the upper half of 16, bit 3 of 8, bit 2 of 4, bit 1 of 2
    const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    const int w = idx & 7, which = idx >> 3;
    if ((lane & 1) == 0 && w < M) atomicAdd(acc + 2 * (acc_off + w) + which, s);
}

}  // namespace tile

// NPW producer warps share the rows of a tile and copy them with 16-byte cp.async (LDGSTS, two warp instructions per row).
// (cp.async.bulk, one TMA request per row and component, was measured first: a 1-D bulk copy of 256-320 bytes costs the
// SM's TMA unit ~33 cycles, 260 of them per tile made the kernel 2.1-2.6 times slower, profiles/r02_k2_tile_copy_modes.txt.)
template <int NDIM, int NPW>
__global__ void __launch_bounds__(32 * (kTileCW + NPW), 1)
    self_moments_tile_kernel(const double* __restrict__ dc, int n_atoms, int Tn, int64_t pitch, int c0, int c1, int c2,
                             const TileTpl* __restrict__ tpls, int n_tpls, int strips_total,
                             const TileUnit* __restrict__ units, int batch, const int32_t* __restrict__ acc_lag, int K,
                             int nbuf, int buf_bytes, unsigned int* work_counter, const int32_t* __restrict__ slot,
                             double* __restrict__ out) {
    using namespace tile;
    constexpr int NT = 32 * (kTileCW + NPW);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // acc [2 * kTileMaxLags] | guard area: full[], empty[], next_unit[], desc[] | the tile buffers.  The look-ahead loads of
    // a column run up to two rows below its tile: below the first buffer they land in the guard area (values unused).
    constexpr int kAccBytes = 2 * kTileMaxLags * 8;
    double* const acc = reinterpret_cast<double*>(smem_raw);
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + kAccBytes);
    uint64_t* const empty = full + kTileMaxBufs;
    int* const next_unit = reinterpret_cast<int*>(empty + kTileMaxBufs);
    TileDesc* const desc = reinterpret_cast<TileDesc*>(reinterpret_cast<unsigned char*>(next_unit) + 16);
    unsigned char* const bufs = smem_raw + kAccBytes + kTileGuardBytes;
    static_assert(kTileMaxBufs <= 4 && 2 * kTileMaxBufs * 8 + 16 + kTileMaxBufs * sizeof(TileDesc) <= kTileGuardBytes, "guard area too small");
    static_assert(kTileGuardBytes + kAccBytes >= 3 * kTileRowP * 8, "look-ahead reads below the first buffer must stay in shared memory");

    __shared__ long long s_item[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp >= kTileCW;
    const int comp[3] = {c0, c1, c2};

#pragma unroll 1
    for (int i = tid; i < 2 * K; i += NT) acc[i] = 0.0;
    if (tid == 0) {
        for (int b = 0; b < kTileMaxBufs; ++b) {
            mbar_init(&full[b], 32 * NPW + 1);
            mbar_init(&empty[b], kTileCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int n_batches = (n_atoms + batch - 1) / batch;
    const int nb_last = n_atoms - (n_batches - 1) * batch;
    const long long items_full = (long long)strips_total * batch, items_last = (long long)strips_total * nb_last;
    const long long n_items = (long long)(n_batches - 1) * items_full + items_last;

    if (producer) {
        // Tiles are drawn from an atomic counter (a CTA that starts late, e.g. behind another stream's kernel on its SM,
        // simply takes fewer): producer warp 0 draws, one tile ahead so that the round trip hides behind the copies, and
        // hands the number to the other producer warps through shared memory and a named barrier.
        const int pw = warp - kTileCW;
        long long drawn = 0;
        if (pw == 0 && lane == 0) s_item[0] = (long long)atomicAdd(work_counter, 1u);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");
#pragma unroll 1
        for (unsigned tau = 0;; ++tau) {
            const int b = (int)(tau % (unsigned)nbuf);
            long long item = s_item[tau & 1];
            if (pw == 0 && lane == 0 && item < n_items) drawn = (long long)atomicAdd(work_counter, 1u);
            int bi = 0, nb = 1, y = 0, al = 0, pack_w = 0, pack_n = 1;
            int4 ta = make_int4(0, 0, 0, 0), tb = ta, tc = ta;
#pragma unroll 1
            while (item < n_items) {
                // item -> (atom batch, strip slot, atom of the batch); strip slot -> (template, strip) by the prefix sums
                long long x;
                if (item < (long long)(n_batches - 1) * items_full) {
                    bi = (int)(item / items_full);
                    x = item - (long long)bi * items_full;
                    nb = batch;
                } else {
                    bi = n_batches - 1;
                    x = item - (long long)(n_batches - 1) * items_full;
                    nb = nb_last;
                }
                y = (int)(x / nb);
                al = (int)(x - (long long)y * nb);
                int ti = 0;
#pragma unroll 1
                for (int base = 0; base < n_tpls; base += 32) {
                    const int idx = base + lane;
                    // strip_prefix of template idx + 1 <= y  <=>  the slot lies beyond template idx
                    const bool below = idx + 1 < n_tpls && __ldg(&tpls[idx + 1].strip_prefix) <= y;
                    const int cnt = __popc(__ballot_sync(0xffffffffu, below));
                    ti += cnt;
                    if (cnt < 32) break;
                }
                ta = __ldg(reinterpret_cast<const int4*>(tpls + ti));
                tb = __ldg(reinterpret_cast<const int4*>(tpls + ti) + 1);
                tc = __ldg(reinterpret_cast<const int4*>(tpls + ti) + 2);
                // the packed last strip of a template: the tile of atom al (a multiple of P) serves atoms al .. al + P - 1,
                // the items of the others are empty -- the item drawn ahead is taken at once and another one drawn
                pack_w = (tc.w > 0 && y - tc.y == tc.x - 1) ? tc.w : 0;
                const int P = tile_pack_atoms(pack_w);
                pack_n = min(P, nb - al);
                if (pack_w == 0 || al % P == 0) break;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");  // (every producer thread has read the slot)
                if (pw == 0 && lane == 0) s_item[tau & 1] = drawn;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");
                item = s_item[tau & 1];
                if (pw == 0 && lane == 0 && item < n_items) drawn = (long long)atomicAdd(work_counter, 1u);
            }
            if (item >= n_items) {
                if (tau >= (unsigned)nbuf) mbar_wait(&empty[b], ((tau / (unsigned)nbuf) - 1) & 1);
                if (pw == 0 && lane == 0) {
                    desc[b].stop = 1;
                    next_unit[b] = 0;
                    mbar_arrive(&full[b]);
                }
                cp_async_arrive(&full[b]);
                break;
            }
            const int D = ta.x, oLo = ta.y, nO = ta.z, pLo = ta.w, nP = tb.x, cP0 = tb.y, spt = tc.z;
            const int t0 = (y - tc.y) * 32 * spt;
            const int ns = min(spt, (D - t0 + 31) / 32);  // strips of this tile (the last group of a row may be short)
            const int sub_bytes = (nO + nP) * (kTileRowP * 8);
            const int atom = bi * batch + al;
            const int pack_wp = tile_pack_wp(pack_w);
            const uint32_t buf_s = smem_u32(bufs + (size_t)b * buf_bytes);
            const double* const src_atom = dc + (size_t)atom * 3 * pitch;
            const int ipitch = (int)pitch;
            // (the tile was decoded while the consumers were still busy with the buffer's previous occupant)
            if (tau >= (unsigned)nbuf) mbar_wait(&empty[b], ((tau / (unsigned)nbuf) - 1) & 1);
            if (pw == 0 && lane == 0) {
                TileDesc d{0, t0, D, oLo, pLo, cP0, tb.z, tb.w, nO, atom, ns, sub_bytes / 8, pack_w, pack_n, pack_wp, 0};
                desc[b] = d;
                next_unit[b] = 0;
                mbar_arrive(&full[b]);  // (release: the descriptor is visible with the phase)
            }
            // Frames are 32-bit here (pitch < 2^30 is checked by the host call).  A chunk that starts at or beyond the pitch
            // is not copied: what is missing is only ever read by masked lanes.  A row is NDIM * 16 (origin) or NDIM * 20
            // (partner) chunks of two frames: chunk i of a row goes to lane i mod 32; rows alternate between the producer
            // warps.
            static_assert(NPW % 2 == 0, "a producer warp copies rows of one parity (odd D)");
            const int c_lo = lane / 20, p_lo = lane - c_lo * 20;                // chunks 0..31 of a row of NDIM * 20
            const int c_hi = (lane + 32) / 20, p_hi = (lane + 32) - c_hi * 20;  // chunks 32..59
            // a packed tile: chunk p of a row belongs to the piece of atom + p / (wp / 2), columns 2 (p mod wp / 2) of it
            const int half_wp = pack_w ? pack_wp / 2 : 20;
            const int j_lo = p_lo / half_wp, j_hi = p_hi / half_wp;
            const int q_lo = 2 * (p_lo - j_lo * half_wp), q_hi = 2 * (p_hi - j_hi * half_wp);  // first column within the piece
            const bool has_lo = c_lo < NDIM && j_lo < pack_n, has_hi = c_hi < NDIM && lane + 32 < NDIM * 20 && j_hi < pack_n;
            const double* const g_lo = src_atom + ((size_t)(has_lo ? j_lo : 0) * 3 + comp[has_lo ? c_lo : 0]) * pitch + q_lo;
            const double* const g_hi = src_atom + ((size_t)(has_hi ? j_hi : 0) * 3 + comp[has_hi ? c_hi : 0]) * pitch + q_hi;
            const uint32_t o_lo = (uint32_t)(c_lo * kTileWp + 2 * p_lo) * 8u + (uint32_t)pw * (kTileRowP * 8);
            const uint32_t o_hi = (uint32_t)(c_hi * kTileWp + 2 * p_hi) * 8u + (uint32_t)pw * (kTileRowP * 8);
            const size_t sstep = (size_t)NPW * D;
#pragma unroll 1
            for (int sidx = 0; sidx < ns; ++sidx) {
                const int ts = t0 + 32 * sidx;
                const uint32_t sub_s = buf_s + (uint32_t)(sidx * sub_bytes);
                // the origin rows, then the partner rows (from column cP0); every row 40 columns from an even frame
#pragma unroll 1
                for (int kind = 0; kind < 2; ++kind) {
                    const int n_rows = kind ? nP : nO, rho0 = (kind ? pLo : oLo) + pw;
                    const int f_first = rho0 * D + ts + (kind ? cP0 : 0) - (D & rho0 & 1);
                    const double* s_lo = g_lo + f_first;
                    const double* s_hi = g_hi + f_first;
                    uint32_t d_lo = sub_s + o_lo + (kind ? (uint32_t)nO * (kTileRowP * 8) : 0u);
                    uint32_t d_hi = sub_s + o_hi + (kind ? (uint32_t)nO * (kTileRowP * 8) : 0u);
                    // rows whose 40 frames lie wholly below the pitch need no test
                    int f = f_first, row = pw;
                    const int whole = ipitch - kTileWp;
#pragma unroll 2
                    for (; row < n_rows && f <= whole; row += NPW, f += NPW * D) {
                        if (has_lo) cp_async16(d_lo, s_lo);
                        if (has_hi) cp_async16(d_hi, s_hi);
                        s_lo += sstep, s_hi += sstep, d_lo += NPW * (kTileRowP * 8), d_hi += NPW * (kTileRowP * 8);
                    }
#pragma unroll 1
                    for (; row < n_rows; row += NPW, f += NPW * D) {
                        if (has_lo && f + q_lo < ipitch) cp_async16(d_lo, s_lo);
                        if (has_hi && f + q_hi < ipitch) cp_async16(d_hi, s_hi);
                        s_lo += sstep, s_hi += sstep, d_lo += NPW * (kTileRowP * 8), d_hi += NPW * (kTileRowP * 8);
                    }
                }
            }
            cp_async_arrive(&full[b]);
            if (pw == 0 && lane == 0) s_item[(tau + 1) & 1] = drawn;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");
        }
    } else {
        // the extra first observation dc[a][k-1] of every lag (kinisi/parser.py:209): units of (lag, 32 atoms)
        {
            const int gw = blockIdx.x * kTileCW + warp, n_gw = gridDim.x * kTileCW;
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
#pragma unroll 1
        for (unsigned tau = 0;; ++tau) {
            const int b = (int)(tau % (unsigned)nbuf);
            mbar_wait(&full[b], (tau / (unsigned)nbuf) & 1);
            const TileDesc d = desc[b];
            if (d.stop) break;
            const double* const buf0 = reinterpret_cast<const double*>(bufs + (size_t)b * buf_bytes);
            const int n_work = d.nunits * d.ns;
#pragma unroll 1
            while (true) {
                int u = 0;
                if (lane == 0) u = atomicAdd(&next_unit[b], 1);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u >= n_work) break;
                // unit-major: the longest units of every strip of the tile go first
                const int ui = u / d.ns, sidx = u - ui * d.ns;
                const double* const bufO = buf0 + (size_t)sidx * d.sub_doubles;
                const double* const bufP = bufO + (size_t)d.nO * kTileRowO;
                // (a packed tile: lane -> column lane mod w of the piece of atom lane / w; the lanes beyond the pieces read
                // piece 0 and their sums are dropped)
                const int pj = d.pack_w ? lane / d.pack_w : 0;
                const int pcl = d.pack_w ? lane - pj * d.pack_w : lane;
                const int lane_off = d.pack_w ? (pj < d.pack_n ? pj * d.pack_wp : 0) + pcl : lane;
                const int t = d.t0 + 32 * sidx + pcl;
                const bool active = t < d.D && pj < d.pack_n;
                const int4 ua = __ldg(reinterpret_cast<const int4*>(units + d.unit0 + ui));
                const int4 ub = __ldg(reinterpret_cast<const int4*>(units + d.unit0 + ui) + 1);
                const int I0 = ua.x, r = ua.y, m = ua.z, acc_off = ua.w, a_top = ub.x, nsteps = ub.y;
                const bool top = ub.z != 0;
                const int a_start = top ? a_top : a_top + (m - 1);
                const int total = nsteps + (top ? 0 : m - 1);
                // the partner of (a_top, first lag) lies past the end of the trajectory for the short lanes of a top unit
                const bool kill = top && ((long long)(a_top + I0) * d.D + t + r > (long long)Tn - 1);
                // (odd D: the row with global index rho is stored from the even frame at or below its start, i.e. shifted
                // right by rho & 1 columns)
                const int odd = d.D & 1, p_start = a_start + I0;
                const double* const po = bufO + (a_start - d.oLo) * kTileRowO + lane_off;
                const double* const pp = bufP + (p_start - d.pLo) * kTileRowP + (r - d.cP0) + lane_off;
                tile::Ptrs q;
                q.o[0] = po + (odd & a_start);
                q.o[1] = po + (odd & ~a_start);
                q.p[0] = pp + (odd & p_start);
                q.p[1] = pp + (odd & ~p_start);
                switch (m) {
                    case 8: column<NDIM, 8>(q, total, top, kill, active, acc_off, acc, lane); break;
                    case 7: column<NDIM, 7>(q, total, top, kill, active, acc_off, acc, lane); break;
                    case 6: column<NDIM, 6>(q, total, top, kill, active, acc_off, acc, lane); break;
                    case 5: column<NDIM, 5>(q, total, top, kill, active, acc_off, acc, lane); break;
                    case 4: column<NDIM, 4>(q, total, top, kill, active, acc_off, acc, lane); break;
                    case 3: column<NDIM, 3>(q, total, top, kill, active, acc_off, acc, lane); break;
                    case 2: column<NDIM, 2>(q, total, top, kill, active, acc_off, acc, lane); break;
                    default: column<NDIM, 1>(q, total, top, kill, active, acc_off, acc, lane); break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[b]);
        }
    }
    __syncthreads();
    // one FP64 reduction per lag, moment and CTA straight into the caller's array (zeroed by the host call)
#pragma unroll 1
    for (int i = tid; i < 2 * K; i += NT) atomicAdd(out + 2 * slot[i >> 1] + (i & 1), acc[i]);
}

// ---- host-side planning -------------------------------------------------------------------------
struct TileWinHost {
    int D, I0, r, m, acc_off;
};

struct TilePlan {
    std::vector<TileTpl> tpls;      // most expensive first
    std::vector<TileUnit> units;
    std::vector<int32_t> slot;      // accumulator index -> position in the caller's lag array
    std::vector<int32_t> acc_lag;   // accumulator index -> lag
    std::vector<TileWinHost> wins;  // for diagnostics
    int strips_total = 0;
    int nbuf = 2;                   // buffers in flight, kTileSmemBufBytes / nbuf bytes each
    long long bytes_loaded = 0;     // per atom
};

static bool tile_packing() {  // KB2_TILE_PACK=0: ragged strips one atom at a time (experiments)
    const char* e = getenv("KB2_TILE_PACK");
    return !(e && !strcmp(e, "0"));
}

static int tile_ra() {  // most origin rows per unit (KB2_TILE_RA: experiments)
    const char* e = getenv("KB2_TILE_RA");
    const int v = e ? atoi(e) : 0;
    return v >= kTileM ? v : kTileRA;
}

// Row length for a progression of `len` lags with common difference d: a multiple of d, >= 32, that fills whole
// strips of 32 columns and still leaves windows of several lags.
static int choose_mult(int d, int len, int Tn) {
    if (const char* e = getenv("KB2_TILE_MULT")) {  // experiments: force the multiple when it is admissible
        const int mult = atoi(e);
        if (mult >= 1 && (long long)mult * d >= 32) return mult;
    }
    double best = 1e300;
    int best_mult = 0;
    for (int mult = 1; mult <= 64; ++mult) {
        const long long D = (long long)mult * d;
        if (D < 32) continue;
        if (D > 2ll * Tn && best_mult) break;
        const int ls = (len + mult - 1) / mult;            // lags of the longest interleaved progression
        if (ls < 3 && best_mult) continue;
        const int nw = (ls + kTileM - 1) / kTileM;
        const double m_avg = (double)ls / nw;
        const double waste = (double)(((D + 31) / 32) * 32) / (double)D;
        const double rows = std::max(1.0, (double)Tn / (double)D);
        // (the interleaved progressions beyond the first have a large r: their tiles need separate partner rows)
        const double score = waste * (1.0 + 0.5 / m_avg) * (1.0 + 2.0 / rows) * (1.0 + 0.15 * (1.0 - 1.0 / mult));
        if (score < best - 1e-12) {
            best = score;
            best_mult = mult;
        }
    }
    if (!best_mult) {
        best_mult = 1;
        while ((long long)best_mult * d < 32) ++best_mult;
    }
    return best_mult;
}

static void plan_tiles_for(const int32_t* lags, int K, int Tn, int buf_bytes, int cap, bool eights, TilePlan& plan) {
    const std::vector<LagSegment> segs = cover_by_progressions(lags, K, Tn, kTileMinLen, kTileCandidates);
    for (const LagSegment& sg : segs) {
        if (sg.len == 1) {
            const int k = sg.k0;
            plan.wins.push_back({kTileSingleD, k / kTileSingleD, k % kTileSingleD, 1, (int)plan.slot.size()});
            plan.slot.push_back(sg.slots[0]);
            plan.acc_lag.push_back(k);
            continue;
        }
        const int mult = choose_mult(sg.d, sg.len, Tn);
        const int D = mult * sg.d;
        for (int u = 0; u < mult && u < sg.len; ++u) {
            const int ls = (sg.len - u + mult - 1) / mult;  // lags k0 + (u + i*mult)*d, i < ls
            // windows of 8 and one remainder ("eights": one instantiation of the walk does nearly all the work, which
            // matters for the instruction cache) or windows of nearly equal size ("even")
            const int nw = (ls + cap - 1) / cap;
            const int base = ls / nw, extra = ls % nw;
            int at = 0;
            for (int w = 0; w < nw; ++w) {
                const int m = eights ? std::min(cap, ls - at) : base + (w < extra ? 1 : 0);
                const long long kfirst = (long long)sg.k0 + (long long)(u + (long long)at * mult) * sg.d;
                plan.wins.push_back({D, (int)(kfirst / D), (int)(kfirst % D), m, (int)plan.slot.size()});
                for (int i = 0; i < m; ++i) {
                    plan.slot.push_back(sg.slots[u + (at + i) * mult]);
                    plan.acc_lag.push_back(sg.k0 + (u + (at + i) * mult) * sg.d);
                }
                at += m;
            }
        }
    }
    // windows by family (D), then by r and I0
    std::vector<int> order(plan.wins.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const TileWinHost &x = plan.wins[a], &y = plan.wins[b];
        if (x.D != y.D) return x.D < y.D;
        if (x.r != y.r) return x.r < y.r;
        return x.I0 < y.I0;
    });
    struct Tile {
        TileTpl tp;
        std::vector<TileUnit> units;
        long long cost;
    };
    std::vector<Tile> tiles;
    size_t g0 = 0;
    while (g0 < order.size()) {
        // an r-group: windows of one D whose r fits one partner strip (r - cP0 <= kTileWp - 32; one column less for an
        // odd D, whose rows may sit one column to the right)
        const TileWinHost& first = plan.wins[order[g0]];
        const int cP0 = first.r & ~1;
        const int r_room = kTileWp - 32 - (first.D & 1);
        size_t g1 = g0;
        while (g1 < order.size() && plan.wins[order[g1]].D == first.D && plan.wins[order[g1]].r - cP0 <= r_room) ++g1;
        std::vector<int> grp(order.begin() + g0, order.begin() + g1);
        g0 = g1;
        std::stable_sort(grp.begin(), grp.end(), [&](int a, int b) { return plan.wins[a].I0 < plan.wins[b].I0; });
        const int D = first.D;
        // SHARED: the partner strip starts at the origin strip's first column, so a row of 40 columns can serve both roles
        // (used for a tile when its origin and partner rows overlap enough that one row set is the smaller one)
        const bool shared = cP0 == 0;
        std::vector<int> rhat(grp.size());  // last origin row of every window (-1: only the first-observation term)
        int rows = 0;
        for (size_t i = 0; i < grp.size(); ++i) {
            const TileWinHost& w = plan.wins[grp[i]];
            const long long span = (long long)Tn - 1 - ((long long)w.I0 * D + w.r);
            rhat[i] = span < 0 ? -1 : (int)(span / D);
            rows = std::max(rows, rhat[i] + 1);
        }
        if (rows == 0) continue;
        // A tile = (block of origin rows [aLo, aHi], run of consecutive windows).  Builds the tiles for blocks of H rows
        // and returns the bytes they load; `out` == nullptr only counts.
        const int RA = tile_ra();
        auto build = [&](int H, std::vector<Tile>* out) -> long long {
            long long total = 0;
            for (int aLo = 0; aLo < rows; aLo += H) {
                const int aHi = std::min(rows, aLo + H) - 1;
                Tile cur;
                int oHi = -1, pLo = 0, pHi = -1;
                auto bytes_of = [&](int o_hi, int p_lo, int p_hi) -> long long {
                    const long long sep = (long long)(o_hi - aLo + 1 + p_hi - p_lo + 1) * kTileRowP * 8;
                    const long long one = (long long)(std::max(o_hi, p_hi) - std::min(aLo, p_lo) + 1) * kTileRowP * 8;
                    return shared && one < sep ? one : sep;
                };
                auto flush = [&]() {
                    if (cur.units.empty()) return;
                    total += bytes_of(oHi, pLo, pHi) + 4096;  // (+ a charge per tile)
                    if (out) {
                        const long long sep = (long long)(oHi - aLo + 1 + pHi - pLo + 1), one = std::max(oHi, pHi) - std::min(aLo, pLo) + 1;
                        if (shared && one < sep) {
                            const int lo = std::min(aLo, pLo), hi = std::max(oHi, pHi);
                            cur.tp = TileTpl{D, lo, 0, lo, hi - lo + 1, 0, 0, (int)cur.units.size(), 0, 0, 1, 0};
                        } else {
                            cur.tp = TileTpl{D, aLo, oHi - aLo + 1, pLo, pHi - pLo + 1, cP0, 0, (int)cur.units.size(), 0, 0, 1, 0};
                        }
                        // small tiles take several strips side by side: enough units for the consumer warps, one fill
                        const int strips = (D + 31) / 32;
                        const long long tb = (long long)(cur.tp.nO + cur.tp.nP) * kTileRowP * 8;
                        int spt = (int)std::min<long long>(std::min(strips, 8), std::max<long long>(1, buf_bytes / tb));
                        while (spt > 1 && (long long)spt * (long long)cur.units.size() > 4 * kTileMaxUnits) --spt;
                        cur.tp.spt = spt;
                        cur.tp.nstrips = (strips + spt - 1) / spt;
                        // a ragged last strip that is a tile of its own takes several atoms side by side (tile_pack_atoms)
                        const int w_last = D - 32 * (strips - 1);
                        cur.tp.pack_w = (spt == 1 && tile_packing() && tile_pack_atoms(w_last) >= 2) ? w_last : 0;
                        std::stable_sort(cur.units.begin(), cur.units.end(), [](const TileUnit& x, const TileUnit& y) {
                            return (long long)x.nsteps * x.m > (long long)y.nsteps * y.m;
                        });
                        cur.cost = 0;
                        for (const TileUnit& u : cur.units) cur.cost += (long long)(u.nsteps + (u.top ? 0 : u.m - 1)) * (16 * u.m + 24);
                        out->push_back(cur);
                    }
                    cur = Tile();
                    oHi = pHi = -1;
                };
                for (size_t i = 0; i < grp.size(); ++i) {
                    const TileWinHost& w = plan.wins[grp[i]];
                    if (rhat[i] < aLo) continue;
                    // The rows of a window are cut at the block boundaries, except that a top piece of fewer than m rows
                    // goes to the block below it: a lower piece loads the m - 1 partner rows above its first origin, and
                    // they must exist.
                    const int top_lo = (rhat[i] / H) * H;
                    const bool merged = rhat[i] - top_lo + 1 < w.m && top_lo >= H;
                    if (merged && aLo == top_lo) continue;
                    const bool top_w = rhat[i] <= aHi || (merged && aLo + H == top_lo);
                    const int hi = top_w ? rhat[i] : aHi;
                    const int plo = aLo + w.I0, phi = hi + w.I0 + (top_w ? 0 : w.m - 1);
                    int n_ohi = hi, n_plo = plo, n_phi = phi;
                    if (!cur.units.empty()) {
                        n_ohi = std::max(oHi, hi);
                        n_plo = std::min(pLo, plo);
                        n_phi = std::max(pHi, phi);
                        const int n_units = (int)cur.units.size() + (hi - aLo + RA) / RA;
                        if (bytes_of(n_ohi, n_plo, n_phi) > buf_bytes || n_units > kTileMaxUnits) {
                            flush();
                            n_ohi = hi, n_plo = plo, n_phi = phi;
                        }
                    }
                    if (bytes_of(n_ohi, n_plo, n_phi) > buf_bytes) return -1;  // one window does not fit: smaller blocks
                    oHi = n_ohi, pLo = n_plo, pHi = n_phi;
                    // units of this window: its range cut from the top into blocks of nearly equal height <= kTileRA
                    const int n_rows = hi - aLo + 1, n_blk = (n_rows + RA - 1) / RA;
                    int top_row = hi;
                    for (int q = 0; q < n_blk; ++q) {
                        const int h = n_rows / n_blk + (q < n_rows % n_blk ? 1 : 0);
                        cur.units.push_back(TileUnit{w.I0, w.r, w.m, w.acc_off, top_row, h, (top_w && q == 0) ? 1 : 0, 0});
                        top_row -= h;
                    }
                }
                flush();
            }
            return total;
        };
        // block height: the one that loads the fewest bytes (whole columns when they fit)
        int best_H = 0;
        long long best_bytes = -1;
        for (int nb = 1; nb <= rows; ++nb) {
            const int H = (rows + nb - 1) / nb;
            if (H < kTileM && nb > 1 && best_H) break;
            const long long bytes = build(H, nullptr);
            if (bytes >= 0 && (best_bytes < 0 || bytes < best_bytes)) {
                best_bytes = bytes;
                best_H = H;
            }
            if (H <= kTileM) break;
            if (best_H && nb >= 8 && H < 4 * RA) break;
        }
        if (!best_H) best_H = kTileM;
        build(best_H, &tiles);
    }
    std::stable_sort(tiles.begin(), tiles.end(), [](const Tile& x, const Tile& y) { return x.cost > y.cost; });
    int prefix = 0;
    for (Tile& tl : tiles) {
        tl.tp.unit0 = (int)plan.units.size();
        tl.tp.strip_prefix = prefix;
        prefix += tl.tp.nstrips;
        plan.units.insert(plan.units.end(), tl.units.begin(), tl.units.end());
        plan.tpls.push_back(tl.tp);
    }
    plan.strips_total = prefix;
    plan.bytes_loaded = 0;
    for (const TileTpl& tp : plan.tpls) plan.bytes_loaded += (long long)(tp.nO + tp.nP) * kTileRowP * 8 * ((tp.D + 31) / 32);
}

// Three buffers of 64 KB keep the producer a tile further ahead; two of 96 KB hold whole columns of ~100 rows.  The smaller
// buffers are taken when they do not cost more than a tenth more traffic.
static void plan_tiles(const int32_t* lags, int K, int Tn, TilePlan& plan) {
    // (measured, tools/k2tile_sweep.sh: two buffers of 96 KB were never slower than three of 64 KB -- tiles of several strips
    // and whole columns matter more than a third tile in flight -- so two is the default; KB2_TILE_BUFS=3 / 0 = by traffic)
    const char* e = getenv("KB2_TILE_BUFS");
    const int force = e ? atoi(e) : 2;
    // How the progressions are cut into windows.  The column walk is instantiated per window length M, its main loop
    // unrolled over the register ring (~0.29 KB x RV(M) x M of code: 23 KB for M = 8, 16 for 7, 14 for 6, 8.6 for 5), and the
    // walks that run side by side on an SM have to share its 32 KB instruction cache: windows of 8 + 3 + 6 (the first 50
    // points of the default grid on 100 000 frames, groups of 11 lags between carries) ran at 0.61 of the FP64 peak with
    // 2.5 no-instruction stall cycles per issue, the same lags cut 6 + 5 at 0.79; windows of 7 + 6 on four progressions of 25
    // at 0.58 against 0.77 for 8 + 8 + 8 + 1 (profiles/r02_k2_tile_split.txt).  So: the cheapest cut by the unit cost model
    // among those whose walks (each with >= 5 % of the work) fit 28 KB; the smallest footprint when none does.
    struct Cut { int cap; bool eights; };
    std::vector<Cut> cuts = {{8, true}, {8, false}, {7, true}, {6, true}, {6, false}, {5, true}};
    if (const char* sp = getenv("KB2_TILE_SPLIT")) {  // experiments: KB2_TILE_SPLIT=even|eights, KB2_TILE_CAP=5..8
        const char* cp = getenv("KB2_TILE_CAP");
        const int cap = cp ? std::min(kTileM, std::max(2, atoi(cp))) : kTileM;
        cuts = {{cap, strcmp(sp, "even") != 0}};
    }
    TilePlan best_plan;
    double best_cost = 0.0, best_foot = 0.0;
    bool have = false, best_fits = false;
    int best_nbuf = 2;
    for (const Cut& cut : cuts) {
        TilePlan three, two;
        if (force != 2) plan_tiles_for(lags, K, Tn, kTileSmemBufBytes / 3, cut.cap, cut.eights, three);
        if (force != 3) plan_tiles_for(lags, K, Tn, kTileSmemBufBytes / 2, cut.cap, cut.eights, two);
        const bool take3 = force == 3 || (force != 2 && three.bytes_loaded * 10 <= two.bytes_loaded * 11);
        TilePlan& cand = take3 ? three : two;
        double by_m[kTileM + 1] = {0}, total = 0.0;
        for (const TileTpl& tp : cand.tpls) {
            const int strips = (tp.D + 31) / 32;
            const double w = tp.pack_w ? strips - 1 + 1.0 / tile_pack_atoms(tp.pack_w) : (double)strips;
            for (int u = 0; u < tp.nunits; ++u) {
                const TileUnit& un = cand.units[tp.unit0 + u];
                const double c = w * (double)(un.nsteps + (un.top ? 0 : un.m - 1)) * (16 * un.m + 24);
                by_m[un.m] += c;
                total += c;
            }
        }
        double foot = 0.0;
        for (int m = 1; m <= kTileM; ++m)
            if (by_m[m] >= 0.05 * total) foot += 0.29 * (((m + 2) / 2) * 2) * m;
        const bool fits = foot <= 28.0;
        const bool better = !have || (fits && !best_fits) || (fits == best_fits && (fits ? total < 0.97 * best_cost : foot < best_foot - 0.01));
        if (better) {
            best_plan = cand;
            best_cost = total, best_foot = foot, best_fits = fits, best_nbuf = take3 ? 3 : 2;
            have = true;
        }
    }
    plan = best_plan;
    plan.nbuf = best_nbuf;
    if (getenv("KB2_TILE_DEBUG")) {
        fprintf(stderr, "tile plan: K %d T %d nbuf %d windows %zu templates %zu units %zu bytes/atom %lld\n", K, Tn, plan.nbuf,
                plan.wins.size(), plan.tpls.size(), plan.units.size(), plan.bytes_loaded);
        for (const TileTpl& tp : plan.tpls) {
            long long steps = 0, terms = 0;
            for (int u = 0; u < tp.nunits; ++u) {
                const TileUnit& un = plan.units[tp.unit0 + u];
                steps += un.nsteps + (un.top ? 0 : un.m - 1);
                terms += (long long)un.nsteps * un.m - (un.top ? (long long)un.m * (un.m - 1) / 2 : 0);
            }
            fprintf(stderr, "  D %5d rows o[%d,+%d) p[%d,+%d) cP0 %d units %d spt %d groups %d steps %lld terms/col %lld KB %.1f\n", tp.D,
                    tp.oLo, tp.nO, tp.pLo, tp.nP, tp.cP0, tp.nunits, tp.spt, tp.nstrips, steps, terms,
                    (tp.nO + tp.nP) * kTileRowP * 8 * tp.spt / 1024.0);
        }
    }
}

// Packed plan of one launch (<= kTileMaxLags lags): header (256 B) | templates | units | slots | lags by accumulator.
struct TileHeader {
    int K, n_tpls, n_units, strips_total;
    int off_tpls, off_units, off_slot, off_lag;
    int l0, bytes, nbuf, pad[5];
};
static_assert(sizeof(TileHeader) <= 256, "the header has 256 bytes");

static size_t tile_plan_bytes(const TilePlan& p, int K) {
    return 256 + align256(p.tpls.size() * sizeof(TileTpl)) + align256(p.units.size() * sizeof(TileUnit)) +
           2 * align256((size_t)K * sizeof(int32_t));
}

static int tile_check_lags(const char* fn, const int32_t* lags_host, int n_lags, int n_frames) {
    KB2_CHECK(lags_host && n_lags > 0 && n_frames > 0, "%s: empty problem", fn);
    for (int i = 0; i < n_lags; ++i) {
        KB2_CHECK(lags_host[i] >= 1 && lags_host[i] <= n_frames, "%s: lag %d out of range", fn, i);
        KB2_CHECK(i == 0 || lags_host[i] >= lags_host[i - 1], "%s: lags must ascend", fn);
    }
    return 0;
}

static size_t tile_smem_bytes() { return (size_t)kTileGuardBytes + 2 * kTileMaxLags * 8 + (size_t)kTileSmemBufBytes; }

}  // namespace kb2

using namespace kb2;

// Size of the packed plan for a lag grid (0 on invalid input): the planner runs to find out.
extern "C" size_t kb2_self_moments_tile_workspace_bytes(const int32_t* lags_host, int n_lags, int n_frames) {
    if (!lags_host || n_lags <= 0 || n_frames <= 0) return 0;
    size_t total = 0;
    for (int l0 = 0; l0 < n_lags; l0 += kTileMaxLags) {
        const int K = std::min(kTileMaxLags, n_lags - l0);
        TilePlan plan;
        plan_tiles(lags_host + l0, K, n_frames, plan);
        total += tile_plan_bytes(plan, K);
    }
    return total;
}

extern "C" int kb2_self_moments_tile_blocks(int n_lags) { return n_lags < 1 ? 1 : (n_lags + kTileMaxLags - 1) / kTileMaxLags; }

// Plans the lag grid on the host into plan_host (kb2_self_moments_tile_workspace_bytes bytes).  Depends on (lags, n_frames) only.
extern "C" int kb2_self_moments_tile_plan_pack(const int32_t* lags_host, int n_lags, int n_frames, void* plan_host) {
    if (int rc = tile_check_lags("kb2_self_moments_tile_plan_pack", lags_host, n_lags, n_frames)) return rc;
    KB2_CHECK(plan_host, "kb2_self_moments_tile_plan_pack: null pointer");
    unsigned char* p = (unsigned char*)plan_host;
    for (int l0 = 0; l0 < n_lags; l0 += kTileMaxLags) {
        const int K = std::min(kTileMaxLags, n_lags - l0);
        TilePlan plan;
        plan_tiles(lags_host + l0, K, n_frames, plan);
        KB2_CHECK((int)plan.slot.size() == K, "kb2_self_moments_tile_plan_pack: internal planning error");
        for (const TileTpl& tp : plan.tpls)
            KB2_CHECK((long long)tp.nO * kTileRowO * 8 + (long long)tp.nP * kTileRowP * 8 <= kTileSmemBufBytes / plan.nbuf && tp.nunits > 0,
                      "kb2_self_moments_tile_plan_pack: a tile exceeds its buffer");
        TileHeader h{};
        h.K = K;
        h.n_tpls = (int)plan.tpls.size();
        h.n_units = (int)plan.units.size();
        h.strips_total = plan.strips_total;
        h.off_tpls = 256;
        h.off_units = h.off_tpls + (int)align256(plan.tpls.size() * sizeof(TileTpl));
        h.off_slot = h.off_units + (int)align256(plan.units.size() * sizeof(TileUnit));
        h.off_lag = h.off_slot + (int)align256((size_t)K * sizeof(int32_t));
        h.l0 = l0;
        h.nbuf = plan.nbuf;
        h.bytes = (int)tile_plan_bytes(plan, K);
        memset(p, 0, (size_t)h.bytes);
        memcpy(p, &h, sizeof(h));
        if (!plan.tpls.empty()) memcpy(p + h.off_tpls, plan.tpls.data(), plan.tpls.size() * sizeof(TileTpl));
        if (!plan.units.empty()) memcpy(p + h.off_units, plan.units.data(), plan.units.size() * sizeof(TileUnit));
        memcpy(p + h.off_slot, plan.slot.data(), (size_t)K * sizeof(int32_t));
        memcpy(p + h.off_lag, plan.acc_lag.data(), (size_t)K * sizeof(int32_t));
        p += h.bytes;
    }
    return 0;
}

template <int NDIM>
static int launch_tile(int sm_count, cudaStream_t st, const double* dc, int n_atoms, int Tn, int64_t pitch, const int comp[3],
                       const TileTpl* tpls, int n_tpls, int strips_total, const TileUnit* units, int batch,
                       const int32_t* acc_lag, int K, int nbuf, unsigned int* counter, const int32_t* slot, double* out) {
    const size_t smem = tile_smem_bytes();
    constexpr int NPW = 2;
    auto kern = self_moments_tile_kernel<NDIM, NPW>;
    const int threads = 32 * (kTileCW + NPW);
    KB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // The CTAs take every register of their SM: a few SMs are left to the one-CTA kernels of the regression (Jacobi, the
    // one-warp sampler) of the analyses in flight on other streams, which would otherwise wait for this kernel to end.
    static const int spare = [] {
        const char* e = getenv("KB2_TILE_SPARE_SMS");
        return e ? std::max(0, atoi(e)) : 2;
    }();
    const int grid = std::max(1, sm_count - spare);
    kern<<<grid, threads, smem, st>>>(dc, n_atoms, Tn, pitch, comp[0], comp[1], comp[2], tpls, n_tpls, strips_total, units,
                                         batch, acc_lag, K, nbuf, kTileSmemBufBytes / nbuf, counter, slot, out);
    KB2_LAUNCH_CHECK();
    return 0;
}

// Launches the tile kernels of a packed plan: plan_host = the host copy (headers are read here), plan_dev = the same bytes
// in device memory (uploaded before this call in stream order), counters = kb2_self_moments_tile_blocks(n_lags) * 256
// bytes of device memory (zeroed by this call).  sums[n_lags][2] is zeroed first.
extern "C" int kb2_self_moments_tile_planned(const double* dc, int n_atoms, int n_frames, int64_t pitch, int n_lags,
                                             int dim_mask, const void* plan_host, const void* plan_dev, void* counters,
                                             double* sums, void* stream) {
    KB2_CHECK(dc && plan_host && plan_dev && counters && sums, "kb2_self_moments_tile_planned: null pointer");
    KB2_CHECK(n_atoms > 0 && n_frames > 0 && n_lags > 0, "kb2_self_moments_tile_planned: empty problem");
    KB2_CHECK(pitch >= n_frames && pitch % 16 == 0 && pitch < (1ll << 30),
              "kb2_self_moments_tile_planned: pitch must be a multiple of 16, >= n_frames and < 2^30");
    KB2_CHECK(((uintptr_t)dc & 15) == 0, "kb2_self_moments_tile_planned: dc must be 16-byte aligned");
    KB2_CHECK(dim_mask >= 1 && dim_mask <= 7, "kb2_self_moments_tile_planned: dim_mask must be in 1..7");
    cudaStream_t st = (cudaStream_t)stream;
    int comp[3] = {0, 0, 0}, ndim = 0;
    for (int c = 0; c < 3; ++c)
        if (dim_mask & (1 << c)) comp[ndim++] = c;
    const int sm_count = device_sm_count();
    KB2_CHECK(sm_count > 0, "kb2_self_moments_tile_planned: no CUDA device");
    const int blocks = (n_lags + kTileMaxLags - 1) / kTileMaxLags;
    KB2_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_lags * 2 * sizeof(double), st));
    KB2_CUDA(cudaMemsetAsync(counters, 0, (size_t)blocks * 256, st));
    const int batch = (int)std::max<long long>(1, std::min<long long>(n_atoms, kTileBatchBytes / (3 * pitch * 8)));
    const unsigned char* ph = (const unsigned char*)plan_host;
    const char* pd = (const char*)plan_dev;
    for (int b = 0; b < blocks; ++b) {
        TileHeader h;
        memcpy(&h, ph, sizeof(h));
        KB2_CHECK(h.K > 0 && h.K <= kTileMaxLags && h.l0 == b * kTileMaxLags && h.l0 + h.K <= n_lags && h.bytes >= 256,
                  "kb2_self_moments_tile_planned: the plan does not match the lag count");
        KB2_CHECK((long long)h.strips_total * n_atoms < (1ll << 31) - (1 << 20), "kb2_self_moments_tile_planned: too many tiles");
        const TileTpl* d_tpls = (const TileTpl*)(pd + h.off_tpls);
        const TileUnit* d_units = (const TileUnit*)(pd + h.off_units);
        const int32_t* d_slot = (const int32_t*)(pd + h.off_slot);
        const int32_t* d_lag = (const int32_t*)(pd + h.off_lag);
        KB2_CHECK(h.nbuf == 2 || h.nbuf == 3, "kb2_self_moments_tile_planned: corrupt plan");
        unsigned int* counter = (unsigned int*)((char*)counters + (size_t)b * 256);
        int rc;
        if (ndim == 3)
            rc = launch_tile<3>(sm_count, st, dc, n_atoms, n_frames, pitch, comp, d_tpls, h.n_tpls, h.strips_total, d_units, batch,
                                d_lag, h.K, h.nbuf, counter, d_slot, sums + 2 * h.l0);
        else if (ndim == 2)
            rc = launch_tile<2>(sm_count, st, dc, n_atoms, n_frames, pitch, comp, d_tpls, h.n_tpls, h.strips_total, d_units, batch,
                                d_lag, h.K, h.nbuf, counter, d_slot, sums + 2 * h.l0);
        else
            rc = launch_tile<1>(sm_count, st, dc, n_atoms, n_frames, pitch, comp, d_tpls, h.n_tpls, h.strips_total, d_units, batch,
                                d_lag, h.K, h.nbuf, counter, d_slot, sums + 2 * h.l0);
        if (rc) return rc;
        ph += h.bytes;
        pd += h.bytes;
    }
    return 0;
}

// What the plan of a grid costs per atom (host only; the variant choice of the host side reads it): out[0] = bytes the
// producers copy into shared memory (all launches, all strips), out[1] = consumer issue cycles (the planner's own unit
// cost, steps x (16 m + 24), over all strips), out[2] = terms formed in the column walks.
extern "C" int kb2_self_moments_tile_plan_cost(const int32_t* lags_host, int n_lags, int n_frames, double* out) {
    if (int rc = tile_check_lags("kb2_self_moments_tile_plan_cost", lags_host, n_lags, n_frames)) return rc;
    KB2_CHECK(out, "kb2_self_moments_tile_plan_cost: null pointer");
    out[0] = out[1] = out[2] = 0.0;
    for (int l0 = 0; l0 < n_lags; l0 += kTileMaxLags) {
        const int K = std::min(kTileMaxLags, n_lags - l0);
        TilePlan plan;
        plan_tiles(lags_host + l0, K, n_frames, plan);
        out[0] += (double)plan.bytes_loaded;
        for (const TileTpl& tp : plan.tpls) {
            double cyc = 0.0, terms = 0.0;
            for (int u = 0; u < tp.nunits; ++u) {
                const TileUnit& un = plan.units[tp.unit0 + u];
                cyc += (double)(un.nsteps + (un.top ? 0 : un.m - 1)) * (16 * un.m + 24);
                terms += (double)un.nsteps * un.m - (un.top ? (double)un.m * (un.m - 1) / 2 : 0.0);
            }
            const int strips = (tp.D + 31) / 32;
            out[1] += cyc * (tp.pack_w ? strips - 1 + 1.0 / tile_pack_atoms(tp.pack_w) : (double)strips);
            out[2] += terms * tp.D;
        }
    }
    return 0;
}

extern "C" int kb2_self_moments_win_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* win_kappa,
                                         int32_t* win_delta, int32_t* win_len, int max_wins);
extern "C" int kb2_self_moments_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* run_k0,
                                     int32_t* run_delta, int32_t* run_len, int32_t* run_n_chg, int max_runs, int* n_rest);

// Which moment kernel for a lag grid (host only).  The tiled kernel reaches 0.6-0.75 of the FP64 peak while its producers
// copy less than ~2 bytes per term; beyond that it is bound by the L2 -> shared-memory traffic (fraction ~ 1 / (1.2 + 0.5 b),
// b = bytes per term), which is what happens when the lags of a grid spread over many residues r of the row length (dense
// grids, short trajectories, grids whose spacing drifts every few lags).  The other kernels do not depend on r:
//   * arithmetic runs: long runs of one spacing; model 0.67 * terms / sum_runs origins * (len + 20 + 10 changes), 32 per
//     term of a lag outside the runs;
//   * register windows: as long as the windows stay long (terms in windows of <= 3 lags <= 10 %, mean length >= 5);
//   * direct: everything else (no structure needed).
// Thresholds from a survey of 68 grids np.linspace(1, T, n, dtype=int), T = 1000 ... 100 000, n = 50 ... 2000
// (tools/k2grid_survey.py, profiles/r02_k2_grid_survey.txt): the choice is within 10 % of the best kernel on 65 of them,
// the geometric mean of chosen / best time is 1.014 (always tiled: 1.75).  features (may be NULL) receives b, the run
// model, the share of short windows and the mean window length.  Returns 0 (direct), 2 (runs), 3 (windows), 4 (tiles),
// or -1 on invalid input.
extern "C" int kb2_self_moments_choose(const int32_t* lags_host, int n_lags, int n_frames, double* features) {
    if (tile_check_lags("kb2_self_moments_choose", lags_host, n_lags, n_frames)) return -1;
    const double T = (double)n_frames;
    double cost[3];
    if (kb2_self_moments_tile_plan_cost(lags_host, n_lags, n_frames, cost)) return -1;
    double terms = 0.0;
    for (int i = 0; i < n_lags; ++i) terms += T - lags_host[i] + 1;
    const double b = cost[0] / std::max(terms, 1.0);
    // runs: the plan of the first launch (192 lags)
    double runs_eff = 0.0;
    {
        const int K = std::min(n_lags, 192);
        std::vector<int32_t> k0(K), dl(K), ln(K), ch(K);
        int rest = 0;
        const int n = kb2_self_moments_plan(lags_host, K, n_frames, k0.data(), dl.data(), ln.data(), ch.data(), K, &rest);
        double t2 = 0.0, run_terms = 0.0, c = 0.0;
        for (int i = 0; i < K; ++i) t2 += T - lags_host[i];
        for (int i = 0; i < n && i < K; ++i) {
            const double origins = T - k0[i] - 0.5 * (ln[i] - 1) * dl[i];
            run_terms += origins * ln[i];
            c += origins * (ln[i] + 20.0 + 10.0 * abs(ch[i]));
        }
        c += 32.0 * std::max(t2 - run_terms, 0.0);
        runs_eff = c > 0.0 ? 0.67 * t2 / c : 0.0;
    }
    // windows: the plan of the first launch (256 lags)
    double small = 1.0, mbar = 0.0;
    {
        const int K = std::min(n_lags, 256);
        std::vector<int32_t> kp(K), dl(K), ln(K);
        const int n = std::min(K, kb2_self_moments_win_plan(lags_host, K, n_frames, kp.data(), dl.data(), ln.data(), K));
        double w_all = 0.0, w_small = 0.0, o_all = 0.0;
        for (int i = 0; i < n; ++i) {
            const double origins = T - kp[i] - 0.5 * (ln[i] - 1) * dl[i];
            w_all += origins * ln[i];
            o_all += origins;
            if (ln[i] <= 3) w_small += origins * ln[i];
        }
        if (w_all > 0.0 && o_all > 0.0) {
            small = w_small / w_all;
            mbar = w_all / o_all;
        }
    }
    if (features) {
        features[0] = b;
        features[1] = runs_eff;
        features[2] = small;
        features[3] = mbar;
    }
    if (b <= 2.0) return 4;
    if (runs_eff >= 0.55) return 2;
    if (small <= 0.1 && mbar >= 5.0) return 3;
    return 0;
}

// The windows of the plan for (the first 256 lags of) a grid, for tests and diagnostics: fills up to max_wins entries of
// win_first (first lag) / win_delta (row length D = lag spacing inside the window) / win_len and returns the number of
// windows; *n_tiles receives the number of tile templates.
extern "C" int kb2_self_moments_tile_plan(const int32_t* lags_host, int n_lags, int n_frames, int32_t* win_first,
                                          int32_t* win_delta, int32_t* win_len, int max_wins, int* n_tiles) {
    if (!lags_host || n_lags <= 0) return 0;
    TilePlan plan;
    plan_tiles(lags_host, std::min(n_lags, kTileMaxLags), n_frames, plan);
    int n = 0;
    for (const TileWinHost& w : plan.wins) {
        if (n < max_wins) {
            if (win_first) win_first[n] = w.I0 * w.D + w.r;
            if (win_delta) win_delta[n] = w.D;
            if (win_len) win_len[n] = w.m;
        }
        ++n;
    }
    if (n_tiles) *n_tiles = (int)plan.tpls.size();
    return n;
}
