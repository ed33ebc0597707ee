// Stage 1: periodic unwrapping (TOR scheme), cumulative displacement, drift correction.
// Replaces Parser.get_disp / Parser.correct_drift (kinisi/parser.py:101-148).
//
// Layout in HBM
//   coords  [N_all][F][3]  f32 or f64, atom-major (what np.concatenate(coords, axis=1) builds, parser.py:120)
//   latt    [F][3][3] f64 (row-vector convention cart = frac @ latt), latt_inv likewise
//   dc out  [n_sel][3][pitch] f64, component-major so that the moment kernels stream one component
//           of one atom as one contiguous run (bulk-copy friendly, conflict-free in shared memory)
//
// Algorithmic HBM bytes per (atom, frame): sizeof(coord)*3 read + 24 written for the scan kernel,
// sizeof(coord)*3 read for the frame-sum kernel.  Both are HBM-bound.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "peer.cuh"

namespace kb2 {

__device__ __forceinline__ void load9(const double* __restrict__ p, double (&m)[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = __ldg(p + i);
}

// Unwrapped displacement between input frames f-1 and f of one atom (parser.py:122-127):
//   diff = frac_f @ L_f - frac_{f-1} @ L_{f-1};  u = diff - floor(diff @ L_f^-1 + 1/2) @ L_f
template <typename T, typename P>
__device__ __forceinline__ void unwrapped_step(const P& cur, const P& prev,
                                               const double (&Lc)[9], const double (&Lp)[9], const double (&Ic)[9],
                                               double (&u)[3]) {
    const double c0 = (double)cur[0], c1 = (double)cur[1], c2 = (double)cur[2];
    const double p0 = (double)prev[0], p1 = (double)prev[1], p2 = (double)prev[2];
    double d[3], img[3];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const double cc = c0 * Lc[l] + c1 * Lc[3 + l] + c2 * Lc[6 + l];
        const double pp = p0 * Lp[l] + p1 * Lp[3 + l] + p2 * Lp[6 + l];
        d[l] = cc - pp;
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) img[l] = floor(d[0] * Ic[l] + d[1] * Ic[3 + l] + d[2] * Ic[6 + l] + 0.5);
#pragma unroll
    for (int l = 0; l < 3; ++l) u[l] = d[l] - (img[0] * Lc[l] + img[1] * Lc[3 + l] + img[2] * Lc[6 + l]);
}

// The same step when every frame shares one lattice L: diff @ L^-1 is the fractional difference itself,
// so u = (w - floor(w + 1/2)) @ L with w = frac_f - frac_{f-1} (21 FP64 operations instead of 48, and
// no lattice traffic).  Differs from the general formula by rounding only.
template <typename T>
__device__ __forceinline__ void unwrapped_step_const(const T (&cur)[3], const T (&prev)[3], const double (&L)[9],
                                                     double (&u)[3]) {
    double w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double d = (double)cur[k] - (double)prev[k];
        w[k] = d - floor(d + 0.5);
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) u[l] = w[0] * L[l] + w[1] * L[3 + l] + w[2] * L[6 + l];
}

__global__ void invert_lattice_kernel(const double* __restrict__ latt, double* __restrict__ inv, int n) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const double* m = latt + (size_t)f * 9;
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], g = m[5], h = m[6], i = m[7], j = m[8];
    const double A = e * j - g * i, B = -(d * j - g * h), C = d * i - e * h;
    const double det = a * A + b * B + c * C;
    const double r = 1.0 / det;
    double* o = inv + (size_t)f * 9;
    o[0] = A * r;
    o[1] = -(b * j - c * i) * r;
    o[2] = (b * g - c * e) * r;
    o[3] = B * r;
    o[4] = (a * j - c * h) * r;
    o[5] = -(a * g - c * d) * r;
    o[6] = C * r;
    o[7] = -(a * i - b * h) * r;
    o[8] = (a * e - b * d) * r;
}

constexpr int kFsBatch = 4;

// a read-only global load the compiler keeps where it is written (volatile asm statements are not reordered among
// themselves): used to issue a batch of independent loads back to back
__device__ __forceinline__ float ld_nc_issue(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_nc_issue(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// ---- frame sums that end with the sum over ranks (atoms sharded over GPUs) -------------------------------------------
// With `enabled`, the frame-sum kernels below finish the join themselves: the CTAs that add into a tile of kPeerFrameTile
// frames count themselves on the tile's arrival counter, and the one that arrives last -- the tile's local sums are then
// complete -- pushes the tile's three rows into the peers' windows, raises its flags, waits for the peers' copies of the
// same tile, adds them in rank order and writes the all-rank sums over the tile (peer_join_slice, csrc/peer.cuh).  Tiles
// are exchanged while the CTAs of later tiles still compute: compute and collective are one kernel, tile by tile, and what
// the next kernel on the stream reads is the joined array.  The protocol is the one of kb2_peer_allreduce_rows_f64, so a
// rank without local framework atoms (or a caller that summed several blocks first) joins the same call with that kernel.
struct FrameJoin {
    PeerJoin J;
    unsigned int* counters;  // [kPeerMaxSlices] arrival counters of this call's tiles, zero on entry and on exit
    int enabled;
    int pad;
};

// called by all threads of a CTA after its additions into tile `tile` (columns [tile * kPeerFrameTile, ...) of the three
// rows of sums): `expected` CTAs add into the tile
__device__ __forceinline__ void frame_join_epilogue(const FrameJoin& fj, double* __restrict__ sums, int64_t pitch, int tile,
                                                    unsigned int expected) {
    __shared__ int s_join;
    __threadfence();  // this thread's additions are performed before the arrival below
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int before = atomicAdd(fj.counters + tile, 1u);
        s_join = (before + 1 == expected);
        if (s_join) {
            atomicExch(fj.counters + tile, 0u);
            __threadfence();
        }
    }
    __syncthreads();
    if (s_join) {
        const long long c0 = (long long)tile * kPeerFrameTile;
        peer_join_slice(fj.J, sums, 3, (long long)pitch, tile, c0, min((long long)pitch, c0 + kPeerFrameTile));
    }
}

// One thread per output frame, looping over a chunk of atoms: per-frame weighted sums of the
// unwrapped displacement.  The sum over atoms commutes with the cumulative sum over time, so the
// framework drift and the collective (MSTD/MSCD) sums need no per-atom scan and no per-atom atomics.
template <typename T, int MODE, bool CONSTL>
__global__ void __launch_bounds__(256) frame_sums_kernel(const T* __restrict__ coords, int F,
                                                         const double* __restrict__ latt,
                                                         const double* __restrict__ inv, int lstride,
                                                         const int32_t* __restrict__ idx, const double* __restrict__ w,
                                                         int n_idx, int per_chunk, double* __restrict__ sums,
                                                         int64_t pitch, int Tn, FrameJoin fj) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t < Tn) {
    const int i0 = blockIdx.y * per_chunk;
    const int i1 = min(n_idx, i0 + per_chunk);
    double acc[3] = {0.0, 0.0, 0.0};
    if (MODE == KB2_MODE_FRACTIONAL && CONSTL) {
        // constant cell: sum_i w_i (frac diff - image) @ L = (sum_i w_i (frac diff - image)) @ L -- the lattice product
        // leaves the atom loop (12 FP64 operations per atom and frame instead of 24)
        double L[9];
        load9(latt, L);
        double aw[3] = {0.0, 0.0, 0.0};
        int i = i0;
        if (!w) {
            // kFsBatch atoms at a time with all their loads issued before the first use: the kernel is bound by the
            // bytes it keeps in flight (one frame pair of one atom is 24 bytes per thread), not by issue or the pipes.
            // (Taking the second frame from the neighbour lane instead of loading it -- half of these loads hit L1 --
            // was slower: 98 us against 80 us; so were batches of 2, 6 and 8.)
            for (; i + kFsBatch <= i1; i += kFsBatch) {
                T v[kFsBatch][6];
                int a[kFsBatch];
#pragma unroll
                for (int u = 0; u < kFsBatch; ++u) a[u] = idx[i + u];
#pragma unroll
                for (int u = 0; u < kFsBatch; ++u) {
                    const T* base = coords + ((size_t)a[u] * F + t) * 3;
#pragma unroll
                    for (int k = 0; k < 6; ++k) v[u][k] = ld_nc_issue(base + k);
                }
#pragma unroll
                for (int u = 0; u < kFsBatch; ++u) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double d = (double)v[u][3 + k] - (double)v[u][k];
                        aw[k] += d - floor(d + 0.5);
                    }
                }
            }
        }
#pragma unroll 4
        for (; i < i1; ++i) {
            const T* base = coords + ((size_t)idx[i] * F + t) * 3;
            const double wi = w ? w[i] : 1.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double d = (double)base[3 + k] - (double)base[k];
                const double wk = d - floor(d + 0.5);
                aw[k] = w ? fma(wi, wk, aw[k]) : aw[k] + wk;
            }
        }
#pragma unroll
        for (int l = 0; l < 3; ++l) acc[l] = aw[0] * L[l] + aw[1] * L[3 + l] + aw[2] * L[6 + l];
    } else if (MODE == KB2_MODE_FRACTIONAL) {
        double Lc[9], Lp[9], Ic[9];
        load9(latt + (size_t)(t + 1) * lstride, Lc);
        load9(latt + (size_t)t * lstride, Lp);
        load9(inv + (size_t)(t + 1) * lstride, Ic);
#pragma unroll 4
        for (int i = i0; i < i1; ++i) {
            const T* base = coords + ((size_t)idx[i] * F + t) * 3;
            double u[3];
            unwrapped_step<T>(base + 3, base, Lc, Lp, Ic, u);
            const double wi = w ? w[i] : 1.0;
            acc[0] += wi * u[0];
            acc[1] += wi * u[1];
            acc[2] += wi * u[2];
        }
    } else {
#pragma unroll 4
        for (int i = i0; i < i1; ++i) {
            const T* base = coords + ((size_t)idx[i] * F + t) * 3;
            const double wi = w ? w[i] : 1.0;
            acc[0] += wi * (double)base[0];
            acc[1] += wi * (double)base[1];
            acc[2] += wi * (double)base[2];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(sums + c * pitch + t, acc[c]);
    }
    if (fj.enabled) {
        // four 256-frame blocks (fewer at the end of the trajectory) times gridDim.y chunks add into one tile
        constexpr int PER = kPeerFrameTile / 256;
        const int tile = blockIdx.x / PER;
        const unsigned int blocks = (unsigned int)min(PER, (int)gridDim.x - tile * PER);
        frame_join_epilogue(fj, sums, pitch, tile, blocks * gridDim.y);
    }
}

// ---- frame sums through a bulk-copy ring -------------------------------------------------------------------------
// The kernel above keeps as many bytes in flight as its registers hold (24 bytes per thread and atom of a batch of four:
// ncu showed 20 long-scoreboard stall cycles per issue at 3.2 TB/s).  Here a CTA owns a tile of kFbFrames frames and a chunk
// of atoms; a producer thread copies the tile of one atom at a time -- one contiguous run of (kFbFrames + 1) * 3 elements
// -- into a ring of kFbStages shared-memory stages with cp.async.bulk (the TMA engine; completion on the stage's `full`
// mbarrier), so kFbStages * 12 KB per CTA are in flight whatever the register count; the consumer warps read the stage
// (conflict-free: a lane's frame is 3 elements), accumulate w - floor(w + 1/2) per frame in registers over the atoms of
// the chunk, and release the stage on its `empty` mbarrier.  Rows of the coordinate array are only element-aligned: a
// copy starts at the 16-byte boundary at or below the tile (`mis` elements early) and is a multiple of 16 bytes; what
// would run past the end of the array is fetched with plain loads.  Constant cell, unit weights (the framework drift of
// every front-end; kinisi/parser.py:131-148); everything else takes the kernel above.
// ncu (profiles/r02_k1_frame_sums_bulk_ncu_summary.txt): 60 us, DRAM 54 %, first stall reason the shared-memory instruction
// queue (4.8 mio_throttle cycles per issue: six 4-byte LDS per frame).  Three CONSECUTIVE frames per thread (12 LDS per 3
// frames, odd lane stride) was measured on the same box: 4.46 -> 4.90 TB/s at 4000 atoms x 50 000 frames but 62.9 -> 66.8 us
// at configs[1] (27 tiles of 768 frames: more, shorter CTAs) -- configs[1] is the only BASELINE shape with a framework.
constexpr int kFbFrames = 1024;
constexpr int kFbStages = 4;
constexpr int kFbConsumers = 256;
constexpr int kFbPer = kFbFrames / kFbConsumers;

template <typename T>
__host__ __device__ constexpr int fb_stage_bytes() {
    return (int)((((kFbFrames + 1) * 3 * sizeof(T) + 16 + 15) / 16) * 16 + 16);
}

namespace fb {
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
        "KB2_FB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KB2_FB_DONE;\n\t"
        "bra KB2_FB_WAIT;\n\t"
        "KB2_FB_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
}  // namespace fb

template <typename T, bool JOIN>
__global__ void __launch_bounds__(kFbConsumers + 32, sizeof(T) == 4 ? 4 : 2) frame_sums_bulk_kernel(const T* __restrict__ coords, int64_t n_coords, int F,
                                                                           const double* __restrict__ latt,
                                                                           const int32_t* __restrict__ idx, int n_idx,
                                                                           int per_chunk, double* __restrict__ sums,
                                                                           int64_t pitch, int Tn, FrameJoin fj) {
    using namespace fb;
    static_assert(kFbFrames == kPeerFrameTile, "a CTA's tile is a slice of the join");
    constexpr int SB = fb_stage_bytes<T>();
    extern __shared__ __align__(128) unsigned char fb_smem[];
    __shared__ uint64_t full[kFbStages], empty[kFbStages];
    __shared__ int s_mis[kFbStages];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = blockIdx.x * kFbFrames;
    const int nf = min(kFbFrames, Tn - t0);
    const int i0 = blockIdx.y * per_chunk, i1 = min(n_idx, i0 + per_chunk);
    if (tid == 0) {
        for (int s = 0; s < kFbStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kFbConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == kFbConsumers / 32) {
        if (lane == 0) {
            const char* const arr_end = reinterpret_cast<const char*>(coords + n_coords);
#pragma unroll 1
            for (int i = i0; i < i1; ++i) {
                const int q = i - i0, s = q % kFbStages;
                if (q >= kFbStages) mbar_wait(&empty[s], ((q / kFbStages) - 1) & 1);
                const T* first = coords + ((size_t)__ldg(idx + i) * F + t0) * 3;
                const int mis = (int)((reinterpret_cast<uintptr_t>(first) & 15) / sizeof(T));
                const char* src = reinterpret_cast<const char*>(first - mis);
                unsigned char* dst = fb_smem + (size_t)s * SB;
                const int n_el = mis + 3 * (nf + 1);
                int bytes = (int)(((size_t)n_el * sizeof(T) + 15) / 16) * 16;
                const long long room = ((arr_end - src) / 16) * 16;  // whole 16-byte chunks before the end of the array
                int bulk = bytes;
                if ((long long)bulk > room) {
                    bulk = (int)room;
                    // the last elements of the whole array: plain loads (ordered before the arrival below)
                    const T* g = reinterpret_cast<const T*>(src + bulk);
                    T* d = reinterpret_cast<T*>(dst + bulk);
                    for (int k = 0; g + k < coords + n_coords && (long long)bulk + (k + 1) * (long long)sizeof(T) <= bytes; ++k) d[k] = g[k];
                }
                s_mis[s] = mis;
                if (bulk > 0) {
                    mbar_arrive_expect_tx(&full[s], (uint32_t)bulk);
                    bulk_g2s(dst, src, (uint32_t)bulk, &full[s]);
                } else {
                    mbar_arrive(&full[s]);
                }
            }
        }
    } else {
    double aw[kFbPer][3];
#pragma unroll
    for (int j = 0; j < kFbPer; ++j) aw[j][0] = aw[j][1] = aw[j][2] = 0.0;
#pragma unroll 1
    for (int i = i0; i < i1; ++i) {
        const int q = i - i0, s = q % kFbStages;
        mbar_wait(&full[s], (q / kFbStages) & 1);
        const T* raw = reinterpret_cast<const T*>(fb_smem + (size_t)s * SB) + s_mis[s];
#pragma unroll
        for (int j = 0; j < kFbPer; ++j) {
            const int f = tid + j * kFbConsumers;
            if (f < nf) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double d = (double)raw[(f + 1) * 3 + k] - (double)raw[f * 3 + k];
                    aw[j][k] += d - floor(d + 0.5);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    double L[9];
    load9(latt, L);
#pragma unroll
    for (int j = 0; j < kFbPer; ++j) {
        const int f = tid + j * kFbConsumers;
        if (f < nf) {
#pragma unroll
            for (int l = 0; l < 3; ++l)
                atomicAdd(sums + l * pitch + t0 + f, aw[j][0] * L[l] + aw[j][1] * L[3 + l] + aw[j][2] * L[6 + l]);
        }
    }
    }
    // (the producer warp comes here too: the join of a tile is the work of the whole CTA; an instantiation of its own, so that
    // the single-GPU kernel does not carry the join's registers: 66 against 56)
    if (JOIN) frame_join_epilogue(fj, sums, pitch, blockIdx.x, gridDim.y);
}

template <typename T>
static int launch_frame_sums_bulk(const void* coords, int n_atoms_total, int F, const double* latt, const int32_t* idx, int n_idx,
                                  double* sums, int64_t pitch, int Tn, const FrameJoin& fj, cudaStream_t st) {
    const int sm = device_sm_count();
    const size_t smem = (size_t)kFbStages * fb_stage_bytes<T>();
    const int tiles = ceil_div(Tn, kFbFrames);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(6, (200 * 1024) / smem));
    int chunks = ceil_div((int64_t)sm * per_sm, tiles);
    chunks = std::max(1, std::min(chunks, ceil_div(n_idx, 2 * kFbStages)));
    chunks = std::min(chunks, 65535);
    const int per = ceil_div(n_idx, chunks);
    chunks = ceil_div(n_idx, per);
    if (fj.enabled) {
        KB2_CUDA(cudaFuncSetAttribute(frame_sums_bulk_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        frame_sums_bulk_kernel<T, true><<<dim3(tiles, chunks), kFbConsumers + 32, smem, st>>>((const T*)coords, (int64_t)n_atoms_total * F * 3,
                                                                                            F, latt, idx, n_idx, per, sums, pitch, Tn, fj);
    } else {
        KB2_CUDA(cudaFuncSetAttribute(frame_sums_bulk_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        frame_sums_bulk_kernel<T, false><<<dim3(tiles, chunks), kFbConsumers + 32, smem, st>>>((const T*)coords, (int64_t)n_atoms_total * F * 3,
                                                                                             F, latt, idx, n_idx, per, sums, pitch, Tn, fj);
    }
    KB2_LAUNCH_CHECK();
    return 0;
}

static bool frame_sums_bulk_enabled() {
    static bool on = [] {
        const char* e = getenv("KB2_FS_BULK");
        return !(e && e[0] == '0');
    }();
    return on;
}

// One CTA per component: optional cumulative sum over time, then an affine map.
__global__ void __launch_bounds__(1024) scan_frames_kernel(double* __restrict__ buf, int n, int64_t pitch,
                                                           int cumulative, double divide,
                                                           const double* __restrict__ sub, double sub_scale) {
    __shared__ double s_tot[32];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* row = buf + (size_t)c * pitch;
    double carry = 0.0;
    for (int t0 = 0; t0 < n; t0 += 1024) {
        const int t = t0 + tid;
        double v = (t < n) ? row[t] : 0.0;
        if (cumulative) {
            v = warp_inclusive_scan(v, lane);
            if (lane == 31) s_tot[warp] = v;
            __syncthreads();
            double pre = carry, tot = 0.0;
#pragma unroll
            for (int wdx = 0; wdx < 32; ++wdx) {
                const double x = s_tot[wdx];
                if (wdx < warp) pre += x;
                tot += x;
            }
            v += pre;
            carry += tot;
            __syncthreads();
        }
        if (t < n) {
            double r = v / divide;
            if (sub) r -= sub_scale * sub[(size_t)c * pitch + t];
            row[t] = r;
        }
    }
}

// Persistent CTAs walk (atom, tile of kUwTile frames) work items taken from a ticket counter; the tiles of an atom are
// processed concurrently by different CTAs and exchange their totals:
//   raw tile   -> shared memory with 16-byte cp.async from the 16-byte-aligned address at or below the tile's first
//                 element (the atom-major [F][3] rows are only element-aligned; the misalignment m shifts the tile by
//                 m elements inside the buffer), double buffered: the loads of the NEXT item are issued before the
//                 current one is scanned, and its ticket is drawn one item earlier still
//   thread     -> kUwE consecutive frames: unwrapped steps, minus step_scale * step_sums (the per-frame framework
//                 mean, subtracted BEFORE the cumulative sum: the sum over atoms commutes with the sum over time),
//                 sequential prefix, then one warp scan + one block scan of the thread totals.  kUwE is odd, so the
//                 per-thread walks through the raw buffer (stride 3 kUwE elements) and the staging buffer (stride kUwE
//                 doubles) are bank-conflict free without padding
//   carry      -> the tile publishes its total in the workspace and, after staging its own prefix, adds the totals of
//                 ALL the earlier tiles of its atom, one thread per predecessor.  A total is its own ready flag (the
//                 workspace is preset to an all-ones bit pattern no sum can produce), so publishing is three plain
//                 stores and collecting is three L2 loads: no fences.  Tickets are drawn in order and an item only waits
//                 for smaller tickets, so the smallest unfinished ticket always makes progress; items are numbered
//                 tile-major (all atoms' tile 0, then tile 1, ...) so that a tile's predecessors are n_idx tickets
//                 back: finished or running, not queued behind another item of the CTA that drew them
//   result     -> staged through shared memory and written as full rows of the component-major dc
// History: one CTA per atom walking its trajectory was 1.5 waves of 296 CTAs that alternated between loading and
// scanning (1.8 TB/s); one CTA per tile exposed four dependent global latencies per tile (ticket, index, tile, flags);
// element-wise cp.async with padded strides spent 340 instructions per frame.
constexpr int kUwE = 5;
constexpr int kUwThreads = 256;
constexpr int kUwCtasPerSm = 3;  // (128 threads x 6 CTAs, the same bytes in flight with twice the tiles per atom: 154 us against 119 us at configs[1])
constexpr int kUwTile = kUwThreads * kUwE;
constexpr unsigned long long kUwNotReady = 0xFFFFFFFFFFFFFFFFull;

template <typename T>
__host__ __device__ constexpr int uw_raw_chunks() {  // 16-byte chunks that cover (kUwTile + 1) frames at any misalignment
    return (int)(((kUwTile + 1) * 3 * sizeof(T) + 15 + 15) / 16);
}
template <typename T>
__host__ __device__ constexpr size_t uw_raw_bytes() { return (size_t)uw_raw_chunks<T>() * 16; }
template <typename T>
__host__ __device__ constexpr size_t uw_smem_bytes() { return 2 * uw_raw_bytes<T>() + 3 * (size_t)kUwTile * sizeof(double); }
// workspace: ticket counter (256 B) | totals [n_idx][n_tiles][3] double
__host__ __device__ inline size_t uw_workspace_bytes(int64_t n_idx, int64_t n_tiles) { return 256 + align256(24 * n_idx * n_tiles); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "n"(sizeof(T))
                 : "memory");
}

__device__ __forceinline__ unsigned long long ld_l2_u64(const double* p) {  // L2 load the compiler may not cache or hoist
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <typename T, int MODE, bool CONSTL>
__global__ void __launch_bounds__(kUwThreads, kUwCtasPerSm) unwrap_cumsum_kernel(const T* __restrict__ coords, int64_t n_coords, int F,
                                                                     const double* __restrict__ latt,
                                                                     const double* __restrict__ inv, int lstride,
                                                                     const int32_t* __restrict__ idx, int n_idx,
                                                                     const double* __restrict__ step_sums, double step_scale,
                                                                     double* __restrict__ out, int64_t pitch, int Tn,
                                                                     int n_tiles, unsigned char* __restrict__ workspace) {
    constexpr int E = kUwE, TF = kUwTile, NT = kUwThreads;
    constexpr bool FRAC = (MODE == KB2_MODE_FRACTIONAL);
    constexpr int PER16 = 16 / (int)sizeof(T);  // elements per chunk
    extern __shared__ __align__(16) unsigned char uw_smem[];
    double* outs = reinterpret_cast<double*>(uw_smem + 2 * uw_raw_bytes<T>());  // [3][TF]
    __shared__ double s_tot[3][NT / 32];
    __shared__ double s_carry[3][NT / 32];
    __shared__ unsigned int s_ticket[4];  // ring of drawn tickets
    __shared__ int s_mis[2];              // misalignment (elements) of the tile in each raw buffer
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int n_work = (unsigned int)n_idx * (unsigned int)n_tiles;
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
    double* totals = reinterpret_cast<double*>(workspace + 256);
    const T* const coords_end = coords + n_coords;

    // loads of work item w into raw buffer b: the tile's first element (frame t0, component 0) lands at element s_mis[b]
    auto issue = [&](unsigned int w, int b) {
        if (w < n_work) {
            const int atom = (int)(w % (unsigned)n_idx), tile = (int)(w / (unsigned)n_idx);
            const int64_t t0 = (int64_t)tile * TF;
            if (t0 < Tn) {
                unsigned char* dst = uw_smem + (size_t)b * uw_raw_bytes<T>();
                const T* first = coords + ((size_t)__ldg(idx + atom) * F + t0) * 3;
                const int mis = (int)((reinterpret_cast<uintptr_t>(first) & 15) / sizeof(T));
                const T* src = first - mis;  // 16-byte aligned
                // elements needed: [0, 3 * frames) past `first`, frames = min(TF, Tn - t0) (+ 1 in fractional mode)
                const int n_el = mis + 3 * ((int)min((int64_t)TF, (int64_t)Tn - t0) + (FRAC ? 1 : 0));
                const int n_chunks = (n_el + PER16 - 1) / PER16;
                if (tid == 0) s_mis[b] = mis;
                for (int ch = tid; ch < n_chunks; ch += NT) {
                    const T* g = src + (size_t)ch * PER16;
                    if (g >= coords && g + PER16 <= coords_end) {
                        cp_async16(dst + (size_t)ch * 16, g);
                    } else {  // the first / last chunk of the whole array: element by element
#pragma unroll
                        for (int k = 0; k < PER16; ++k)
                            if (g + k >= coords && g + k < coords_end)
                                cp_async_elem<T>(reinterpret_cast<T*>(dst + (size_t)ch * 16) + k, g + k);
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double Lconst[9];
    if (FRAC && CONSTL) load9(latt, Lconst);
    if (tid == 0) {
        s_ticket[0] = atomicAdd(counter, 1u);
        s_ticket[1] = atomicAdd(counter, 1u);
    }
    __syncthreads();
    issue(s_ticket[0], 0);
    for (int it = 0;; ++it) {
        const unsigned int w = s_ticket[it & 3];
        if (w >= n_work) break;
        const unsigned int w_next = s_ticket[(it + 1) & 3];
        issue(w_next, (it + 1) & 1);                                      // in flight while this item is scanned
        if (tid == 0) s_ticket[(it + 2) & 3] = atomicAdd(counter, 1u);  // read two barriers from here at the earliest
        const int atom = (int)(w % (unsigned)n_idx), tile = (int)(w / (unsigned)n_idx);  // tile-major order (see above)
        const int64_t t0 = (int64_t)tile * TF;
        double* tot_atom = totals + (size_t)atom * n_tiles * 3;
        // the body of an item, instantiated without bounds tests for tiles that lie fully inside the trajectory
        auto body = [&](auto full_tag) __attribute__((always_inline)) {
            constexpr bool FULL = decltype(full_tag)::value;
        // the drift of this thread's frames: L2 loads that land while the tile is unwrapped
        const int f0 = tid * E;
        // look back early: thread j asks for the total of tile j of this atom now (an L2 round trip that overlaps the
        // unwrapping); whatever is not ready yet is asked for again after the tile's own prefix has been staged
        unsigned long long early[3] = {kUwNotReady, kUwNotReady, kUwNotReady};
        if (FRAC && tid < tile) {
#pragma unroll
            for (int c = 0; c < 3; ++c) early[c] = ld_l2_u64(tot_atom + tid * 3 + c);
        }
        double drift[E][3];
        {
            const double* ds[3] = {step_sums + t0 + f0, step_sums + pitch + t0 + f0, step_sums + 2 * pitch + t0 + f0};
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int64_t t = t0 + f0 + e;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    drift[e][c] = (step_sums && (FULL || t < Tn)) ? step_scale * __ldg(ds[c] + e) : 0.0;
            }
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const T* raw = reinterpret_cast<const T*>(uw_smem + (size_t)(it & 1) * uw_raw_bytes<T>()) + s_mis[it & 1];

        double val[E][3];
        if (FULL || t0 + f0 < Tn) {
            // the thread's frames f0 .. f0 + E (one more than it owns in fractional mode), read once
            T r[E + 1][3];
#pragma unroll
            for (int e = 0; e < E + (FRAC ? 1 : 0); ++e)
#pragma unroll
                for (int k = 0; k < 3; ++k) r[e][k] = (FULL || t0 + f0 + e < Tn + (FRAC ? 1 : 0)) ? raw[(f0 + e) * 3 + k] : (T)0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int64_t t = t0 + f0 + e;
                double u[3] = {0.0, 0.0, 0.0};
                if (FULL || t < Tn) {
                    if (FRAC) {
                        if (CONSTL) {
                            unwrapped_step_const<T>(r[e + 1], r[e], Lconst, u);
                        } else {
                            double Lc[9], Lp[9], Ic[9];
                            load9(latt + (size_t)(t + 1) * lstride, Lc);
                            load9(latt + (size_t)t * lstride, Lp);
                            load9(inv + (size_t)(t + 1) * lstride, Ic);
                            unwrapped_step<T>(r[e + 1], r[e], Lc, Lp, Ic, u);
                        }
                    } else {
                        u[0] = (double)r[e][0];
                        u[1] = (double)r[e][1];
                        u[2] = (double)r[e][2];
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) u[c] -= drift[e][c];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) val[e][c] = u[c];
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e)
#pragma unroll
                for (int c = 0; c < 3; ++c) val[e][c] = 0.0;
        }
        if (FRAC) {
            // thread-sequential prefix, then the exclusive offset of this thread from a warp scan and a block scan
#pragma unroll
            for (int e = 1; e < E; ++e)
#pragma unroll
                for (int c = 0; c < 3; ++c) val[e][c] += val[e - 1][c];
            double incl[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                incl[c] = warp_inclusive_scan(val[E - 1][c], lane);
                if (lane == 31) s_tot[c][warp] = incl[c];
            }
            __syncthreads();
            // publish the tile total before collecting: no tile ever waits for another tile's wait
            if (tid >= NT - 3 && n_tiles > 1) {
                const int c = tid - (NT - 3);
                double tot = 0.0;
#pragma unroll
                for (int wdx = 0; wdx < NT / 32; ++wdx) tot += s_tot[c][wdx];
                __stcg(tot_atom + tile * 3 + c, tot);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double pre = 0.0;
#pragma unroll
                for (int wdx = 0; wdx < NT / 32; ++wdx) pre += (wdx < warp) ? s_tot[c][wdx] : 0.0;
                const double off = pre + (incl[c] - val[E - 1][c]);
#pragma unroll
                for (int e = 0; e < E; ++e) val[e][c] += off;
            }
        }
        // the prefix inside the tile goes to the staging buffer ...
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int c = 0; c < 3; ++c) outs[c * TF + f0 + e] = val[e][c];
        // ... and only now the totals of the earlier tiles of this atom are collected, one thread per predecessor
        if (FRAC && tile > 0) {
            double pc[3] = {0.0, 0.0, 0.0};
            for (int j = tid; j < tile; j += NT) {
                unsigned long long bits[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) bits[c] = (j == tid) ? early[c] : ld_l2_u64(tot_atom + j * 3 + c);
                while (bits[0] == kUwNotReady || bits[1] == kUwNotReady || bits[2] == kUwNotReady) {
                    __nanosleep(32);
#pragma unroll
                    for (int c = 0; c < 3; ++c) bits[c] = ld_l2_u64(tot_atom + j * 3 + c);  // three loads in flight
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) pc[c] += __longlong_as_double((long long)bits[c]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                pc[c] = warp_sum(pc[c]);
                if (lane == 0) s_carry[c][warp] = pc[c];
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double carry = 0.0;
            if (FRAC && tile > 0) {
#pragma unroll
                for (int wdx = 0; wdx < NT / 32; ++wdx) carry += s_carry[c][wdx];
            }
            double* orow = out + ((size_t)atom * 3 + c) * pitch + t0;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int f = tid + i * NT;
                const int64_t t = t0 + f;
                if (FULL)
                    orow[f] = outs[c * TF + f] + carry;
                else if (t < pitch)
                    orow[f] = (t < Tn) ? outs[c * TF + f] + carry : 0.0;
            }
        }
        };
        if (t0 + TF <= Tn)
            body(std::true_type{});
        else
            body(std::false_type{});
        // the next item overwrites outs / s_tot / s_carry only after its own first barrier; its raw buffer is the one
        // this item read, refilled by the issue() at the top of the NEXT iteration, after every thread passed the
        // barrier above
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- unwrap + cumulative sum, one warp per atom ------------------------------------------------------------------
// The kernel above cuts an atom into tiles handled by different CTAs and pays for it: three block barriers per tile (ncu:
// 3.4 barrier-stall cycles per issue), a look-back through L2 flags, two memsets per call, 200 instructions per frame.
// With at least a couple of atoms per SM there is enough parallelism ACROSS atoms, and the scan along time can stay inside
// one warp: a warp owns an atom and walks its trajectory in sub-tiles of kWuE * 32 frames, carrying the running sum in
// registers -- no barrier, no flag, no workspace.  What a warp cannot do with its own loads it gets from the TMA engine:
//   stage s of the warp's ring (kWuStages stages)
//     drift / out  [3][TW] f64    the per-frame framework sums of the sub-tile, bulk-copied from step_sums (L2-resident);
//                                 the results overwrite them in place and leave with three bulk stores
//     raw          (TW + 1) * 3   the sub-tile of the atom-major coordinate row, one contiguous bulk copy from the 16-byte
//                                 boundary at or below its first element (`mis` elements early; the last bytes of the
//                                 whole array, which a 16-byte copy would overrun, come with plain loads)
//   lane 0 issues the copies of sub-tile i + kWuStages - 1 after the stores of sub-tile i - 1 have read their stage
//   (cp.async.bulk.wait_group.read), every lane waits on the stage's mbarrier; a lane owns kWuE consecutive frames
//   (kWuE odd: its walks through `raw`, stride 3 kWuE words, and through `out`, stride kWuE doubles, are conflict-free),
//   unwraps them, scans them, adds the exclusive prefix of the lower lanes (five shuffle steps per component) and the
//   running sum, and lane 31's last value is the new running sum.
// So the bytes in flight no longer depend on registers or occupancy (6 warps per SM keep 18 stages = 146 KB in flight), a
// frame costs ~60 instructions, and the sub-tiles of consecutive atoms follow each other without draining the ring.
// Constant cell only (every front-end but an NPT run); the kernel above keeps the other modes and the calls with too few
// atoms to fill the device this way.
constexpr int kWuE = 7;
constexpr int kWuTW = 32 * kWuE;  // frames per sub-tile (a multiple of 16: a sub-tile of the output row is 16-byte aligned)
constexpr int kWuStages = 4;
constexpr int kWuMaxWarps = 4;

template <typename T>
__host__ __device__ constexpr int wu_raw_bytes() {
    return (int)((((kWuTW + 1) * 3 * sizeof(T) + 15 + 15) / 16) * 16);
}
template <typename T>
__host__ __device__ constexpr int wu_stage_bytes() {
    return 3 * kWuTW * 8 + wu_raw_bytes<T>();
}
template <typename T>
__host__ __device__ constexpr int wu_warps_per_cta() {
    return sizeof(T) == 4 ? 3 : 2;  // two CTAs per SM either way (97 KB / 105 KB of shared memory each)
}

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(fb::smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}

template <typename T, bool DRIFT>
__global__ void __launch_bounds__(kWuMaxWarps * 32) unwrap_cumsum_warp_kernel(const T* __restrict__ coords, int64_t n_coords, int F,
                                                                         const double* __restrict__ latt,
                                                                         const int32_t* __restrict__ idx, int n_idx,
                                                                         const double* __restrict__ step_sums, double step_scale,
                                                                         double* __restrict__ out, int64_t pitch, int Tn) {
    using namespace fb;
    constexpr int E = kWuE, TW = kWuTW, S = kWuStages, SB = wu_stage_bytes<T>();
    extern __shared__ __align__(128) unsigned char wu_smem[];
    __shared__ uint64_t full[kWuMaxWarps][S];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    unsigned char* ring = wu_smem + (size_t)warp * S * SB;
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    // atoms of this warp: i = blockIdx.x + gridDim.x * (warp + wpc * m), m = 0, 1, ... (consecutive atoms go to different SMs)
    const int first_atom = blockIdx.x + gridDim.x * warp;
    const int atom_step = gridDim.x * wpc;
    if (first_atom >= n_idx) return;
    const int n_my = (n_idx - first_atom + atom_step - 1) / atom_step;
    const int n_sub = (int)((pitch + TW - 1) / TW);
    const int64_t n_jobs = (int64_t)n_my * n_sub;
    const char* const arr_end = reinterpret_cast<const char*>(coords + n_coords);

    // lane 0: the copies of job j (atom m = j / n_sub, sub-tile j % n_sub) into stage j % S
    auto issue = [&](int64_t j) {
        const int m = (int)(j / n_sub), sub = (int)(j - (int64_t)m * n_sub), s = (int)(j % S);
        const int a = first_atom + m * atom_step;
        const int t0 = sub * TW;
        const int nf = max(0, min(TW, Tn - t0));
        unsigned char* st = ring + (size_t)s * SB;
        if (nf == 0) {  // a sub-tile of the zero pad: nothing to fetch
            mbar_arrive(&full[warp][s]);
            return;
        }
        const T* first = coords + ((size_t)__ldg(idx + a) * F + t0) * 3;
        const int mis = (int)((reinterpret_cast<uintptr_t>(first) & 15) / sizeof(T));
        const char* src = reinterpret_cast<const char*>(first - mis);
        unsigned char* dst = st + 3 * TW * 8;
        const int bytes = (int)((((size_t)(mis + 3 * (nf + 1)) * sizeof(T)) + 15) / 16) * 16;
        const long long room = ((arr_end - src) / 16) * 16;
        int bulk = bytes;
        if ((long long)bulk > room) {
            bulk = (int)room;
            const T* g = reinterpret_cast<const T*>(src + bulk);
            T* d = reinterpret_cast<T*>(dst + bulk);
            for (int k = 0; g + k < coords + n_coords && (long long)bulk + (k + 1) * (long long)sizeof(T) <= bytes; ++k) d[k] = g[k];
        }
        const int nd = (int)min((int64_t)TW, pitch - t0) * 8;  // bytes of one component of the drift (pitch is even: 16 | nd)
        mbar_arrive_expect_tx(&full[warp][s], (uint32_t)(bulk + (DRIFT ? 3 * nd : 0)));
        if (bulk > 0) bulk_g2s(dst, src, (uint32_t)bulk, &full[warp][s]);
        if (DRIFT) {
#pragma unroll
            for (int c = 0; c < 3; ++c) bulk_g2s(st + c * TW * 8, step_sums + c * pitch + t0, (uint32_t)nd, &full[warp][s]);
        }
    };

    double L[9];
    load9(latt, L);
    if (lane == 0) {
        for (int64_t j = 0; j < min((int64_t)(S - 1), n_jobs); ++j) issue(j);
    }
    double carry[3] = {0.0, 0.0, 0.0};
    int sub = 0, m = 0, mis = 0;
    int a = first_atom;
#pragma unroll 1
    for (int64_t j = 0; j < n_jobs; ++j) {
        const int s = (int)(j % S);
        if (sub == 0) {
            const T* row = coords + (size_t)__ldg(idx + a) * F * 3;
            mis = (int)((reinterpret_cast<uintptr_t>(row) & 15) / sizeof(T));  // TW * 3 elements are a multiple of 16 bytes:
            carry[0] = carry[1] = carry[2] = 0.0;                             // the same for every sub-tile of the atom
        }
        const int t0 = sub * TW;
        unsigned char* st = ring + (size_t)s * SB;
        double* outs = reinterpret_cast<double*>(st);
        mbar_wait(&full[warp][s], (uint32_t)((j / S) & 1));
        const int f0 = lane * E;
        double val[E][3];
        if (t0 + f0 < Tn) {
            const T* raw = reinterpret_cast<const T*>(st + 3 * TW * 8) + mis + f0 * 3;
            T r[E + 1][3];
#pragma unroll
            for (int e = 0; e <= E; ++e)
#pragma unroll
                for (int k = 0; k < 3; ++k) r[e][k] = raw[e * 3 + k];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                double u[3];
                unwrapped_step_const<T>(r[e + 1], r[e], L, u);
                const bool live = (t0 + f0 + e < Tn);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double x = u[c];
                    if (DRIFT) x -= step_scale * outs[c * TW + f0 + e];
                    val[e][c] = live ? x : 0.0;
                }
            }
#pragma unroll
            for (int e = 1; e < E; ++e)
#pragma unroll
                for (int c = 0; c < 3; ++c) val[e][c] += val[e - 1][c];
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e)
#pragma unroll
                for (int c = 0; c < 3; ++c) val[e][c] = 0.0;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double incl = warp_inclusive_scan(val[E - 1][c], lane);
            const double off = carry[c] + (incl - val[E - 1][c]);
#pragma unroll
            for (int e = 0; e < E; ++e) val[e][c] += off;
            carry[c] = __shfl_sync(0xffffffffu, val[E - 1][c], 31);
        }
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int c = 0; c < 3; ++c) outs[c * TW + f0 + e] = (t0 + f0 + e < Tn) ? val[e][c] : 0.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const uint32_t nd = (uint32_t)min((int64_t)TW, pitch - t0) * 8u;
#pragma unroll
            for (int c = 0; c < 3; ++c) bulk_s2g(out + ((size_t)a * 3 + c) * pitch + t0, outs + c * TW, nd);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (j + S - 1 < n_jobs) {
                // stage (j - 1) % S is refilled once the stores of job j - 1 have read it (the group before the newest)
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                issue(j + S - 1);
            }
        }
        __syncwarp();
        if (++sub == n_sub) {
            sub = 0;
            ++m;
            a += atom_step;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <typename T>
static int launch_unwrap_warp(const void* coords, int n_atoms_total, int F, const double* latt, const int32_t* idx, int n_idx,
                              const double* step_sums, double step_scale, double* out, int64_t pitch, int Tn, cudaStream_t st) {
    const int sm = device_sm_count();
    const int wpc = wu_warps_per_cta<T>();
    const size_t smem = (size_t)wpc * kWuStages * wu_stage_bytes<T>();
    const int grid = std::min(2 * sm, ceil_div(n_idx, wpc));
#define KB2_WU(D)                                                                                                            \
    do {                                                                                                                     \
        KB2_CUDA(cudaFuncSetAttribute(unwrap_cumsum_warp_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        unwrap_cumsum_warp_kernel<T, D><<<grid, wpc * 32, smem, st>>>((const T*)coords, (int64_t)n_atoms_total * F * 3, F, latt, idx, \
                                                                      n_idx, step_sums, step_scale, out, pitch, Tn);        \
    } while (0)
    if (step_sums)
        KB2_WU(true);
    else
        KB2_WU(false);
#undef KB2_WU
    KB2_LAUNCH_CHECK();
    return 0;
}

// ---- unwrap + cumulative sum, tiles through bulk copies -------------------------------------------------------------
// The (atom, 1280-frame tile) items, the ticket counter and the look-back over tile totals of unwrap_cumsum_kernel, with
// the three things ncu charged it for taken out:
//   * the 256-thread 16-byte cp.async loop with its bounds tests, and the 15 strided 8-byte drift loads per thread, are
//     four bulk copies issued by ONE thread of a producer warp (TMA engine, completion on the stage's `full` mbarrier);
//     the producer draws the tickets and runs up to two items ahead of the consumers;
//   * the drift lands in the [3][TF] staging area, the results overwrite it in place and leave through bulk stores issued
//     per warp (no output pass, no barrier in front of it); a stage goes back to the producer (`empty` mbarrier) once the
//     stores have read it;
//   * one named barrier per tile (the warp totals) instead of three block barriers: the look-back is done by every warp
//     for itself (a tile has at most a few dozen predecessors), before the results are staged.
// Constant cell; NPT cells and displacement input take the kernels above.
constexpr int kUbE = 5;
// consumer threads NT (+ one producer warp); a tile is NT * kUbE frames.  float32 coordinates: 256 threads, 1280-frame tiles,
// two 45 KB stages, two CTAs per SM.  float64 coordinates (what ASE and pymatgen hand over): the same stage would be 61 KB --
// one CTA per SM -- so 128 threads, 640-frame tiles, two 31 KB stages, three CTAs per SM.
template <typename T>
__host__ __device__ constexpr int ub_threads() { return sizeof(T) == 4 ? 256 : 128; }
constexpr int kUbMinTile = 128 * kUbE;  // the workspace is sized for the shorter tile (kb2_unwrap_cumsum_workspace_bytes)
static_assert(256 * kUbE == kUwTile, "the float32 tile is the tile of unwrap_cumsum_kernel");

template <typename T>
__host__ __device__ constexpr int ub_raw_bytes() {
    return (int)((((ub_threads<T>() * kUbE + 1) * 3 * sizeof(T) + 15 + 15) / 16) * 16);
}
template <typename T>
__host__ __device__ constexpr int ub_stage_bytes() {
    return 3 * ub_threads<T>() * kUbE * 8 + ub_raw_bytes<T>();
}

template <typename T, bool DRIFT>
__global__ void __launch_bounds__(ub_threads<T>() + 32, sizeof(T) == 4 ? 2 : 3) unwrap_cumsum_bulk_kernel(const T* __restrict__ coords, int64_t n_coords, int F,
                                                                              const double* __restrict__ latt,
                                                                              const int32_t* __restrict__ idx, int n_idx,
                                                                              const double* __restrict__ step_sums, double step_scale,
                                                                              double* __restrict__ out, int64_t pitch, int Tn,
                                                                              int n_tiles, unsigned char* __restrict__ workspace, int opts) {
    using namespace fb;
    constexpr int E = kUbE, NT = ub_threads<T>(), TF = NT * E, NW = NT / 32, SB = ub_stage_bytes<T>();
    extern __shared__ __align__(128) unsigned char ub_smem[];
    __shared__ uint64_t full[2], empty[2];
    __shared__ unsigned int s_item[2];
    __shared__ int s_mis[2], s_atom[2], s_tile[2];
    __shared__ double s_tot[2][3][NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int n_work = (unsigned int)n_idx * (unsigned int)n_tiles;
    unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);  // preset to 0xFFFFFFFF: the first ticket wraps to 0
    double* totals = reinterpret_cast<double*>(workspace + 256);
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&empty[b], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == NW) {
        // ---- producer: tickets and bulk copies, up to two items ahead
        if (lane == 0) {
            const char* const arr_end = reinterpret_cast<const char*>(coords + n_coords);
            // The ticket of an item is drawn one item early and its coordinate run is prefetched into L2 then
            // (cp.async.bulk.prefetch.L2): the copy into the stage, issued one item later, finds it there -- the consumers
            // spent 16 % of their stall samples waiting for `full` with the HBM latency inside the one item of lead that
            // two stages give.
            // (Only with many atoms, opts bit 1: the extra ticket in flight per CTA brings the tiles of one atom closer together
            // in time, and at configs[1] -- 448 atoms, 296 CTAs -- the look-back waits cost more than the prefetch gains:
            // 76 -> 81-93 us, against 4.39 -> 4.59 TB/s at 4000 atoms.  Without it the ticket is drawn when the stage is free.)
            const bool ahead = (opts & 2) != 0;
            unsigned int w_ahead = ahead ? atomicAdd(counter, 1u) + 1u : 0u;
#pragma unroll 1
            for (int it = 0;; ++it) {
                const int b = it & 1;
                if (!ahead) {
                    if (it >= 2) mbar_wait(&empty[b], (uint32_t)(((it >> 1) - 1) & 1));
                    w_ahead = atomicAdd(counter, 1u) + 1u;
                }
                const unsigned int w = w_ahead;
                if (ahead && w < n_work) {
                    w_ahead = atomicAdd(counter, 1u) + 1u;
                    if (w_ahead < n_work) {
                        const int atom_a = (int)(w_ahead % (unsigned)n_idx), tile_a = (int)(w_ahead / (unsigned)n_idx);
                        const int t0_a = tile_a * TF;
                        const int nf_a = min(TF, Tn - t0_a);
                        if (nf_a > 0) {
                            const T* first_a = coords + ((size_t)__ldg(idx + atom_a) * F + t0_a) * 3;
                            const char* src_a = reinterpret_cast<const char*>(first_a) - (reinterpret_cast<uintptr_t>(first_a) & 15);
                            long long bytes_a = (((reinterpret_cast<const char*>(first_a + 3 * (nf_a + 1)) - src_a) + 15) / 16) * 16;
                            bytes_a = min(bytes_a, (long long)(((arr_end - src_a) / 16) * 16));
                            if (bytes_a > 0)
                                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_a), "r"((uint32_t)bytes_a) : "memory");
                        }
                    }
                }
                if (ahead && it >= 2) mbar_wait(&empty[b], (uint32_t)(((it >> 1) - 1) & 1));
                s_item[b] = w;
                if (w >= n_work) {
                    mbar_arrive(&full[b]);
                    break;
                }
                const int atom = (int)(w % (unsigned)n_idx), tile = (int)(w / (unsigned)n_idx);  // tile-major order
                const int t0 = tile * TF;
                const int nf = max(0, min(TF, Tn - t0));
                unsigned char* st = ub_smem + (size_t)b * SB;
                s_atom[b] = atom;
                s_tile[b] = tile;
                if (nf == 0) {  // a tile of the zero pad
                    s_mis[b] = 0;
                    mbar_arrive(&full[b]);
                    continue;
                }
                const T* first = coords + ((size_t)__ldg(idx + atom) * F + t0) * 3;
                const int mis = (int)((reinterpret_cast<uintptr_t>(first) & 15) / sizeof(T));
                const char* src = reinterpret_cast<const char*>(first - mis);
                unsigned char* dst = st + 3 * TF * 8;
                const int bytes = (int)((((size_t)(mis + 3 * (nf + 1)) * sizeof(T)) + 15) / 16) * 16;
                const long long room = ((arr_end - src) / 16) * 16;
                int bulk = bytes;
                if ((long long)bulk > room) {  // the last bytes of the whole array: plain loads
                    bulk = (int)room;
                    const T* g = reinterpret_cast<const T*>(src + bulk);
                    T* d = reinterpret_cast<T*>(dst + bulk);
                    for (int k = 0; g + k < coords + n_coords && (long long)bulk + (k + 1) * (long long)sizeof(T) <= bytes; ++k) d[k] = g[k];
                }
                s_mis[b] = mis;
                const int nd = (int)min((int64_t)TF, pitch - t0) * 8;
                mbar_arrive_expect_tx(&full[b], (uint32_t)(bulk + (DRIFT ? 3 * nd : 0)));
                if (bulk > 0) bulk_g2s(dst, src, (uint32_t)bulk, &full[b]);
                if (DRIFT) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) bulk_g2s(st + c * TF * 8, step_sums + c * pitch + t0, (uint32_t)nd, &full[b]);
                }
            }
        }
        return;
    }
    // ---- consumers
    double L[9];
    load9(latt, L);
    const int f0 = tid * E;
#pragma unroll 1
    for (int it = 0;; ++it) {
        const int b = it & 1;
        mbar_wait(&full[b], (uint32_t)((it >> 1) & 1));
        const unsigned int w = s_item[b];
        if (w >= n_work) break;
        if (it >= 1 && lane == 0) {  // the stores of the previous item have read the other stage: back to the producer at once
            // (handing it back after the unwrapping instead, when the wait is free, costs the producer half an item of lead:
            // 84 -> 91 us at configs[1])
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&empty[b ^ 1]);
        }
        const int atom = s_atom[b], tile = s_tile[b];
        const int t0 = tile * TF;
        double* tot_atom = totals + (size_t)atom * n_tiles * 3;
        unsigned char* st = ub_smem + (size_t)b * SB;
        double* outs = reinterpret_cast<double*>(st);
        // the body of an item, instantiated without the end-of-trajectory tests for tiles that lie inside it
        auto body = [&](auto full_tag) __attribute__((always_inline)) {
            constexpr bool FULL = decltype(full_tag)::value;
        // look back early: lane j asks for the total of tile j of this atom now (an L2 round trip under the unwrapping)
        unsigned long long early[3] = {kUwNotReady, kUwNotReady, kUwNotReady};
        if (lane < tile) {
#pragma unroll
            for (int c = 0; c < 3; ++c) early[c] = ld_l2_u64(tot_atom + lane * 3 + c);
        }
        double val[E][3];
        if (FULL || t0 + f0 < Tn) {
            const T* raw = reinterpret_cast<const T*>(st + 3 * TF * 8) + s_mis[b] + f0 * 3;
            T r[E + 1][3];
#pragma unroll
            for (int e = 0; e <= E; ++e)
#pragma unroll
                for (int k = 0; k < 3; ++k) r[e][k] = raw[e * 3 + k];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                double u[3];
                unwrapped_step_const<T>(r[e + 1], r[e], L, u);
                const bool live = FULL || (t0 + f0 + e < Tn);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double x = u[c];
                    if (DRIFT) x -= step_scale * outs[c * TF + f0 + e];
                    val[e][c] = live ? x : 0.0;
                }
            }
#pragma unroll
            for (int e = 1; e < E; ++e)
#pragma unroll
                for (int c = 0; c < 3; ++c) val[e][c] += val[e - 1][c];
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e)
#pragma unroll
                for (int c = 0; c < 3; ++c) val[e][c] = 0.0;
        }
        double incl[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            incl[c] = warp_inclusive_scan(val[E - 1][c], lane);
            if (lane == 31) s_tot[b][c][warp] = incl[c];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // the consumer warps only
        // publish the tile total before collecting: no tile ever waits for another tile's wait
        if (tid >= NT - 3 && n_tiles > 1) {
            const int c = tid - (NT - 3);
            double tot = 0.0;
#pragma unroll
            for (int wdx = 0; wdx < NW; ++wdx) tot += s_tot[b][c][wdx];
            __stcg(tot_atom + tile * 3 + c, tot);
        }
        // the totals of the earlier tiles of this atom, collected by every warp for itself
        double pc[3] = {0.0, 0.0, 0.0};
        for (int j = lane; j < tile; j += 32) {
            unsigned long long bits[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) bits[c] = (j == lane) ? early[c] : ld_l2_u64(tot_atom + j * 3 + c);
            while (bits[0] == kUwNotReady || bits[1] == kUwNotReady || bits[2] == kUwNotReady) {
                __nanosleep(32);
#pragma unroll
                for (int c = 0; c < 3; ++c) bits[c] = ld_l2_u64(tot_atom + j * 3 + c);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) pc[c] += __longlong_as_double((long long)bits[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // the totals of the earlier tiles (one per lane) and of the lower warps of this tile (lanes 0 .. warp - 1) in ONE
            // butterfly: a loop over the lower warps cost the last warp 80 instructions behind the barrier
            double off = warp_sum(pc[c] + ((lane < warp) ? s_tot[b][c][lane] : 0.0));
            off += incl[c] - val[E - 1][c];
#pragma unroll
            for (int e = 0; e < E; ++e) outs[c * TF + f0 + e] = (FULL || t0 + f0 + e < Tn) ? val[e][c] + off : 0.0;
        }
        };
        if (!(opts & 1) && t0 + TF <= Tn)
            body(std::true_type{});
        else
            body(std::false_type{});
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const int64_t w0 = (int64_t)t0 + warp * 32 * E;  // this warp's 160 frames of the three rows
            const int64_t n_out = min((int64_t)(32 * E), pitch - w0);
            if (n_out > 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    bulk_s2g(out + ((size_t)atom * 3 + c) * pitch + w0, outs + c * TF + warp * 32 * E, (uint32_t)(n_out * 8));
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <typename T>
static int launch_unwrap_bulk(const void* coords, int n_atoms_total, int F, const double* latt, const int32_t* idx, int n_idx,
                              const double* step_sums, double step_scale, double* out, int64_t pitch, int Tn, void* workspace,
                              cudaStream_t st) {
    const int sm = device_sm_count();
    const size_t smem = 2 * (size_t)ub_stage_bytes<T>();
    const int n_tiles = ceil_div(pitch, ub_threads<T>() * kUbE);
    KB2_CHECK((int64_t)n_idx * n_tiles < (1ll << 31), "kb2_unwrap_cumsum: too many (atom, tile) pairs for one call");
    // ticket counter = 0xFFFFFFFF (the first atomicAdd + 1 wraps to ticket 0); totals = the "not ready" pattern: one memset
    KB2_CUDA(cudaMemsetAsync(workspace, 0xFF, uw_workspace_bytes(n_idx, n_tiles), st));
    const int grid = (int)std::min<int64_t>((int64_t)n_idx * n_tiles, (int64_t)sm * (sizeof(T) == 4 ? 2 : 3));
    const char* eo = getenv("KB2_UB_OPTS");  // bit 0: no separate instantiation for full tiles (diagnostics); bit 1: L2 prefetch
    const int opts = eo ? atoi(eo) : ((int64_t)n_idx >= 4ll * grid ? 2 : 0);
#define KB2_UB(D)                                                                                                            \
    do {                                                                                                                     \
        KB2_CUDA(cudaFuncSetAttribute(unwrap_cumsum_bulk_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        unwrap_cumsum_bulk_kernel<T, D><<<grid, ub_threads<T>() + 32, smem, st>>>((const T*)coords, (int64_t)n_atoms_total * F * 3, F,  \
                                                                             latt, idx, n_idx, step_sums, step_scale, out, pitch, \
                                                                             Tn, n_tiles, (unsigned char*)workspace, opts); \
    } while (0)
    if (step_sums)
        KB2_UB(true);
    else
        KB2_UB(false);
#undef KB2_UB
    KB2_LAUNCH_CHECK();
    return 0;
}

// which kernel kb2_unwrap_cumsum takes for float32 / float64 fractional coordinates in a constant cell.  KB2_UW_KERNEL:
// 0 the cp.async tile kernel, 1 the warp-per-atom kernel, 2 the bulk-copy tile kernel, unset: by atom count (see the caller)
static int unwrap_kernel_choice() {
    const char* e = getenv("KB2_UW_KERNEL");  // (read per call: the tests switch it)
    return e && e[0] ? atoi(e) : -1;
}


// Weighted sum over atoms of a component-major displacement array: out[c][t] += sum_a w[a] dc[a][c][t].
// The collective displacement of MSTD (diffusion.py:659) and the charge-weighted one of MSCD (:751).
__global__ void __launch_bounds__(256) atom_sum_kernel(const double* __restrict__ dc, int n_atoms, int64_t pitch,
                                                       const double* __restrict__ w, int per_chunk,
                                                       double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= pitch) return;
    const int c = blockIdx.y;
    const int a0 = blockIdx.z * per_chunk, a1 = min(n_atoms, a0 + per_chunk);
    double acc = 0.0;
#pragma unroll 4
    for (int a = a0; a < a1; ++a) {
        const double x = dc[((size_t)a * 3 + c) * pitch + t];
        acc = w ? fma(w[a], x, acc) : acc + x;
    }
    atomicAdd(out + (size_t)c * pitch + t, acc);
}

template <typename T, int MODE, bool CONSTL>
static int launch_unwrap(const void* coords, int n_atoms_total, int F, const double* latt, const double* inv, int lstride,
                         const int32_t* idx, int n_idx, const double* step_sums, double step_scale, double* out,
                         int64_t pitch, int Tn, void* workspace, cudaStream_t st) {
    const T* c = (const T*)coords;
    const size_t smem = uw_smem_bytes<T>();
    const int n_tiles = ceil_div(pitch, kUwTile);
    KB2_CHECK((int64_t)n_idx * n_tiles < (1ll << 31), "kb2_unwrap_cumsum: too many (atom, tile) pairs for one call");
    KB2_CUDA(cudaFuncSetAttribute(unwrap_cumsum_kernel<T, MODE, CONSTL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    // ticket counter = 0; totals = the "not ready" pattern
    KB2_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
    KB2_CUDA(cudaMemsetAsync((unsigned char*)workspace + 256, 0xFF, uw_workspace_bytes(n_idx, n_tiles) - 256, st));
    const int sm = device_sm_count();
    const int grid = (int)std::min<int64_t>((int64_t)n_idx * n_tiles, (int64_t)sm * kUwCtasPerSm);
    unwrap_cumsum_kernel<T, MODE, CONSTL><<<grid, kUwThreads, smem, st>>>(
        c, (int64_t)n_atoms_total * F * 3, F, latt, inv, lstride, idx, n_idx, step_sums, step_scale, out, pitch, Tn, n_tiles,
        (unsigned char*)workspace);
    KB2_LAUNCH_CHECK();
    return 0;
}

static int check_stage1(const char* fn, const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in,
                        const double* latt, const double* latt_inv, int latt_stride, const int32_t* idx, int n_idx,
                        int64_t pitch, int* Tn) {
    KB2_CHECK(coords && idx, "%s: null pointer", fn);
    KB2_CHECK(dtype == KB2_F32 || dtype == KB2_F64, "%s: dtype must be KB2_F32 or KB2_F64", fn);
    KB2_CHECK(mode == KB2_MODE_FRACTIONAL || mode == KB2_MODE_DISPLACEMENT, "%s: bad mode", fn);
    KB2_CHECK(n_atoms_total > 0 && n_idx > 0, "%s: no atoms", fn);
    if (mode == KB2_MODE_FRACTIONAL) {
        KB2_CHECK(latt && latt_inv, "%s: lattices required in fractional mode", fn);
        KB2_CHECK(latt_stride == 9 || latt_stride == 0, "%s: latt_stride must be 9 or 0", fn);
        KB2_CHECK(n_frames_in >= 2, "%s: need at least two frames", fn);
        *Tn = n_frames_in - 1;
    } else {
        KB2_CHECK(n_frames_in >= 1, "%s: need at least one frame", fn);
        *Tn = n_frames_in;
    }
    KB2_CHECK(pitch >= *Tn && pitch % 2 == 0, "%s: pitch must be even and >= the number of output frames", fn);
    return 0;
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_invert_lattice(const double* latt, double* latt_inv, int n, void* stream) {
    KB2_CHECK(latt && latt_inv && n > 0, "kb2_invert_lattice: bad arguments");
    invert_lattice_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(latt, latt_inv, n);
    KB2_LAUNCH_CHECK();
    return 0;
}

static int frame_sums_impl(const char* fn, const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in,
                           const double* latt, const double* latt_inv, int latt_stride, const int32_t* idx,
                           const double* w, int n_idx, double* frame_sums, int64_t pitch, const FrameJoin& fj, void* stream) {
    int Tn = 0;
    if (int rc = check_stage1(fn, coords, dtype, mode, n_atoms_total, n_frames_in, latt, latt_inv,
                              latt_stride, idx, n_idx, pitch, &Tn))
        return rc;
    KB2_CHECK(frame_sums, "%s: null output", fn);
    if (fj.enabled)
        KB2_CHECK(ceil_div(pitch, kPeerFrameTile) == ceil_div(Tn, kPeerFrameTile) && fj.J.n_slices == ceil_div(pitch, kPeerFrameTile),
                  "%s: the row pad must not reach into a tile of %d frames of its own", fn, kPeerFrameTile);
    cudaStream_t st = (cudaStream_t)stream;
    KB2_CUDA(cudaMemsetAsync(frame_sums, 0, sizeof(double) * 3 * pitch, st));
    if (mode == KB2_MODE_FRACTIONAL && latt_stride == 0 && !w && frame_sums_bulk_enabled() && n_idx >= 2 * kFbStages) {
        if (dtype == KB2_F32)
            return launch_frame_sums_bulk<float>(coords, n_atoms_total, n_frames_in, latt, idx, n_idx, frame_sums, pitch, Tn, fj, st);
        return launch_frame_sums_bulk<double>(coords, n_atoms_total, n_frames_in, latt, idx, n_idx, frame_sums, pitch, Tn, fj, st);
    }
    const int sm = device_sm_count();
    const int tiles = ceil_div(Tn, 256);
    int chunks = ceil_div((int64_t)sm * 8, tiles);
    chunks = max(1, min(chunks, ceil_div(n_idx, 8)));
    chunks = min(chunks, 65535);
    const int per = ceil_div(n_idx, chunks);
    chunks = ceil_div(n_idx, per);
    dim3 grid(tiles, chunks);
#define KB2_FS(T, M, C)                                                                                             \
    frame_sums_kernel<T, M, C><<<grid, 256, 0, st>>>((const T*)coords, n_frames_in, latt, latt_inv, latt_stride, idx, w, \
                                                     n_idx, per, frame_sums, pitch, Tn, fj)
    const bool constl = (mode == KB2_MODE_FRACTIONAL && latt_stride == 0);
    if (dtype == KB2_F32 && mode == KB2_MODE_FRACTIONAL) { if (constl) KB2_FS(float, KB2_MODE_FRACTIONAL, true); else KB2_FS(float, KB2_MODE_FRACTIONAL, false); }
    else if (dtype == KB2_F32) KB2_FS(float, KB2_MODE_DISPLACEMENT, false);
    else if (mode == KB2_MODE_FRACTIONAL) { if (constl) KB2_FS(double, KB2_MODE_FRACTIONAL, true); else KB2_FS(double, KB2_MODE_FRACTIONAL, false); }
    else KB2_FS(double, KB2_MODE_DISPLACEMENT, false);
#undef KB2_FS
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_frame_sums(const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in,
                              const double* latt, const double* latt_inv, int latt_stride, const int32_t* idx,
                              const double* w, int n_idx, double* frame_sums, int64_t pitch, void* stream) {
    FrameJoin fj;
    memset(&fj, 0, sizeof(fj));
    return frame_sums_impl("kb2_frame_sums", coords, dtype, mode, n_atoms_total, n_frames_in, latt, latt_inv, latt_stride, idx, w,
                           n_idx, frame_sums, pitch, fj, stream);
}

extern "C" int kb2_frame_sums_joined(const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in,
                                     const double* latt, const double* latt_inv, int latt_stride, const int32_t* idx,
                                     const double* w, int n_idx, double* frame_sums, int64_t pitch, void* const* windows_host,
                                     int world, int rank, int64_t capacity, uint64_t seq, unsigned int* tile_counters,
                                     void* stream) {
    KB2_CHECK(tile_counters, "kb2_frame_sums_joined: null counters");
    KB2_CHECK(3 * pitch <= capacity, "kb2_frame_sums_joined: 3 * pitch exceeds the capacity of the windows");
    KB2_CHECK((reinterpret_cast<uintptr_t>(frame_sums) & 15) == 0 && pitch % 2 == 0, "kb2_frame_sums_joined: frame_sums must be 16-byte aligned, pitch even");
    const int slices = kb2_peer_rows_slices(pitch);
    KB2_CHECK(slices > 0, "kb2_frame_sums_joined: too many frames for one join (use kb2_frame_sums + kb2_peer_allreduce_f64)");
    FrameJoin fj;
    memset(&fj, 0, sizeof(fj));
    if (int rc = peer_join_setup(&fj.J, windows_host, world, rank, capacity, seq, slices, "kb2_frame_sums_joined")) return rc;
    fj.counters = tile_counters;
    fj.enabled = 1;
    return frame_sums_impl("kb2_frame_sums_joined", coords, dtype, mode, n_atoms_total, n_frames_in, latt, latt_inv, latt_stride,
                           idx, w, n_idx, frame_sums, pitch, fj, stream);
}

extern "C" int kb2_scan_frames(double* buf, int n_frames, int64_t pitch, int cumulative, double divide,
                               const double* sub, double sub_scale, void* stream) {
    KB2_CHECK(buf && n_frames > 0 && pitch >= n_frames, "kb2_scan_frames: bad arguments");
    KB2_CHECK(divide != 0.0, "kb2_scan_frames: divide must be non-zero");
    scan_frames_kernel<<<3, 1024, 0, (cudaStream_t)stream>>>(buf, n_frames, pitch, cumulative, divide, sub, sub_scale);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t kb2_unwrap_cumsum_workspace_bytes(int n_idx, int64_t pitch) {
    if (n_idx < 1 || pitch < 1) return 256;
    return uw_workspace_bytes(n_idx, ceil_div(pitch, kUbMinTile));  // (the kernel with the shortest tile: float64 bulk copies)
}

extern "C" int kb2_unwrap_cumsum(const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in,
                                 const double* latt, const double* latt_inv, int latt_stride, const int32_t* idx,
                                 int n_idx, const double* step_sums, double step_scale, double* dc_out, int64_t pitch,
                                 void* workspace, void* stream) {
    int Tn = 0;
    if (int rc = check_stage1("kb2_unwrap_cumsum", coords, dtype, mode, n_atoms_total, n_frames_in, latt, latt_inv,
                              latt_stride, idx, n_idx, pitch, &Tn))
        return rc;
    KB2_CHECK(dc_out && workspace, "kb2_unwrap_cumsum: null output or workspace");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == KB2_MODE_FRACTIONAL && latt_stride == 0 && (reinterpret_cast<uintptr_t>(dc_out) & 15) == 0 &&
        (!step_sums || (reinterpret_cast<uintptr_t>(step_sums) & 15) == 0)) {
        int choice = unwrap_kernel_choice();
        if (choice < 0) {
            // The tickets of an atom's consecutive tiles are n_idx apart: with fewer atoms than about 4/3 of the persistent grid
            // a tile runs at the same time as its predecessor and waits for its total in the middle of the item, before its
            // results can be staged -- the cp.async kernel looks back AFTER staging and loses less.  Measured at 20 000 frames
            // (tools/k1probe.py, profiles/r02_k1_probe.txt): float32 133 atoms 95 against 83 us, 300 atoms equal, 448 atoms 91
            // against 118 us, 2000 atoms 4.4 against 3.5 TB/s; float64 448 atoms 215 against 135 us, 700 atoms 5.1 against
            // 3.4 TB/s.
            const int64_t grid = (int64_t)device_sm_count() * (dtype == KB2_F32 ? 2 : 3);
            choice = ((int64_t)n_idx * 3 >= grid * 4) ? 2 : 0;
        }
        if (choice == 2) {
            if (dtype == KB2_F32)
                return launch_unwrap_bulk<float>(coords, n_atoms_total, n_frames_in, latt, idx, n_idx, step_sums, step_scale, dc_out, pitch, Tn, workspace, st);
            return launch_unwrap_bulk<double>(coords, n_atoms_total, n_frames_in, latt, idx, n_idx, step_sums, step_scale, dc_out, pitch, Tn, workspace, st);
        }
        if (choice == 1) {
            if (dtype == KB2_F32)
                return launch_unwrap_warp<float>(coords, n_atoms_total, n_frames_in, latt, idx, n_idx, step_sums, step_scale, dc_out, pitch, Tn, st);
            return launch_unwrap_warp<double>(coords, n_atoms_total, n_frames_in, latt, idx, n_idx, step_sums, step_scale, dc_out, pitch, Tn, st);
        }
    }
#define KB2_UC(T, M, C) \
    return launch_unwrap<T, M, C>(coords, n_atoms_total, n_frames_in, latt, latt_inv, latt_stride, idx, n_idx, step_sums, step_scale, dc_out, pitch, Tn, workspace, st)
    const bool constl = (mode == KB2_MODE_FRACTIONAL && latt_stride == 0);
    if (dtype == KB2_F32 && mode == KB2_MODE_FRACTIONAL) { if (constl) KB2_UC(float, KB2_MODE_FRACTIONAL, true); KB2_UC(float, KB2_MODE_FRACTIONAL, false); }
    if (dtype == KB2_F32) KB2_UC(float, KB2_MODE_DISPLACEMENT, false);
    if (mode == KB2_MODE_FRACTIONAL) { if (constl) KB2_UC(double, KB2_MODE_FRACTIONAL, true); KB2_UC(double, KB2_MODE_FRACTIONAL, false); }
    KB2_UC(double, KB2_MODE_DISPLACEMENT, false);
#undef KB2_UC
}

extern "C" int kb2_atom_sum(const double* dc, int n_atoms, int64_t pitch, const double* w, double* out, void* stream) {
    KB2_CHECK(dc && out && n_atoms > 0 && pitch > 0, "kb2_atom_sum: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    KB2_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 3 * pitch, st));
    const int sm = device_sm_count();
    const int tiles = ceil_div(pitch, 256);
    int chunks = ceil_div((int64_t)sm * 8, (int64_t)tiles * 3);
    chunks = max(1, min(chunks, ceil_div(n_atoms, 8)));
    chunks = min(chunks, 65535);
    const int per = ceil_div(n_atoms, chunks);
    chunks = ceil_div(n_atoms, per);
    atom_sum_kernel<<<dim3(tiles, 3, chunks), 256, 0, st>>>(dc, n_atoms, pitch, w, per, out);
    KB2_LAUNCH_CHECK();
    return 0;
}
