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
#include "common.cuh"

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

// One thread per output frame, looping over a chunk of atoms: per-frame weighted sums of the
// unwrapped displacement.  The sum over atoms commutes with the cumulative sum over time, so the
// framework drift and the collective (MSTD/MSCD) sums need no per-atom scan and no per-atom atomics.
template <typename T, int MODE, bool CONSTL>
__global__ void __launch_bounds__(256) frame_sums_kernel(const T* __restrict__ coords, int F,
                                                         const double* __restrict__ latt,
                                                         const double* __restrict__ inv, int lstride,
                                                         const int32_t* __restrict__ idx, const double* __restrict__ w,
                                                         int n_idx, int per_chunk, double* __restrict__ sums,
                                                         int64_t pitch, int Tn) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= Tn) return;
    const int i0 = blockIdx.y * per_chunk;
    const int i1 = min(n_idx, i0 + per_chunk);
    double acc[3] = {0.0, 0.0, 0.0};
    if (MODE == KB2_MODE_FRACTIONAL) {
        double Lc[9], Lp[9], Ic[9];
        load9(latt + (size_t)(t + 1) * lstride, Lc);
        if (!CONSTL) {
            load9(latt + (size_t)t * lstride, Lp);
            load9(inv + (size_t)(t + 1) * lstride, Ic);
        }
#pragma unroll 4
        for (int i = i0; i < i1; ++i) {
            const T* base = coords + ((size_t)idx[i] * F + t) * 3;
            double u[3];
            if (CONSTL) {
                const T cur[3] = {base[3], base[4], base[5]}, prev[3] = {base[0], base[1], base[2]};
                unwrapped_step_const<T>(cur, prev, Lc, u);
            } else {
                unwrapped_step<T>(base + 3, base, Lc, Lp, Ic, u);
            }
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

// One CTA per selected atom, walking the trajectory in tiles of kUwTile = 256 * kUwE frames with a running carry.
//   raw tile   -> shared memory with cp.async (coalesced 4- or 8-byte copies, the next tile in flight while the
//                 current one is scanned; the atom-major [F][3] rows are only element-aligned, so no bulk copy)
//   thread     -> kUwE consecutive frames: unwrapped steps, minus step_scale * step_sums (the per-frame framework
//                 mean, subtracted BEFORE the cumulative sum: the sum over atoms commutes with the sum over time),
//                 sequential prefix, then one warp scan + one block scan of the thread totals per tile
//   result     -> staged through shared memory and written as full rows of the component-major dc
// Shared-memory strides are padded by one element per kUwE frames so that the per-thread walks are conflict free.
constexpr int kUwE = 8;
constexpr int kUwTile = 256 * kUwE;

__host__ __device__ constexpr int uw_raw_elems() { return (kUwTile + 1) * 3 + (kUwTile + 1) / kUwE + 4; }
__host__ __device__ constexpr int uw_out_elems() { return kUwTile + kUwTile / kUwE; }
template <typename T>
__host__ __device__ constexpr size_t uw_smem_bytes() {
    return ((2 * uw_raw_elems() * sizeof(T) + 15) & ~(size_t)15) + 3 * uw_out_elems() * sizeof(double);
}

template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "n"(sizeof(T))
                 : "memory");
}

template <typename T, int MODE, bool CONSTL>
__global__ void __launch_bounds__(256, 2) unwrap_cumsum_kernel(const T* __restrict__ coords, int F,
                                                               const double* __restrict__ latt,
                                                               const double* __restrict__ inv, int lstride,
                                                               const int32_t* __restrict__ idx, int n_idx,
                                                               const double* __restrict__ step_sums, double step_scale,
                                                               double* __restrict__ out, int64_t pitch, int Tn) {
    constexpr int E = kUwE, TF = kUwTile;
    constexpr bool FRAC = (MODE == KB2_MODE_FRACTIONAL);
    extern __shared__ __align__(16) unsigned char uw_smem[];
    T* raw = reinterpret_cast<T*>(uw_smem);  // [2][uw_raw_elems()]
    double* outs = reinterpret_cast<double*>(uw_smem + ((2 * uw_raw_elems() * sizeof(T) + 15) & ~(size_t)15));
    __shared__ double s_tot[3][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int atom = blockIdx.x;
    const T* base = coords + (size_t)idx[atom] * F * 3;
    const int64_t n_in = (int64_t)F * 3;  // elements of this atom's row

    // raw element e of the tile starting at frame t0 is global element t0*3 + e; frame f, component k sits at
    // f*3 + k + f/E
    auto issue = [&](int64_t t0, int buf) {
        T* dst = raw + (size_t)buf * uw_raw_elems();
        const int64_t g0 = t0 * 3;
        const int n = (TF + (FRAC ? 1 : 0)) * 3;
        for (int e = tid; e < n; e += 256) {
            if (g0 + e < n_in) {
                const int f = e / 3;
                cp_async_elem<T>(dst + e + f / E, base + g0 + e);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    double Lconst[9];
    if (FRAC && CONSTL) load9(latt, Lconst);
    double carry[3] = {0.0, 0.0, 0.0};
    const int n_tiles = (int)((pitch + TF - 1) / TF);
    issue(0, 0);
    for (int tile = 0; tile < n_tiles; ++tile) {
        const int64_t t0 = (int64_t)tile * TF;
        if (tile + 1 < n_tiles && t0 + TF < Tn) issue(t0 + TF, (tile + 1) & 1); else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const T* r = raw + (size_t)(tile & 1) * uw_raw_elems();
        double val[E][3];
        const int f0 = tid * E;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int f = f0 + e;
            const int64_t t = t0 + f;
            const bool valid = t < Tn;
            double u[3] = {0.0, 0.0, 0.0};
            if (valid) {
                const int o0 = f * 3 + f / E;
                if (FRAC) {
                    const int o1 = (f + 1) * 3 + (f + 1) / E;
                    const T prev[3] = {r[o0], r[o0 + 1], r[o0 + 2]};
                    const T cur[3] = {r[o1], r[o1 + 1], r[o1 + 2]};
                    if (CONSTL) {
                        unwrapped_step_const<T>(cur, prev, Lconst, u);
                    } else {
                        double Lc[9], Lp[9], Ic[9];
                        load9(latt + (size_t)(t + 1) * lstride, Lc);
                        load9(latt + (size_t)t * lstride, Lp);
                        load9(inv + (size_t)(t + 1) * lstride, Ic);
                        unwrapped_step<T>(cur, prev, Lc, Lp, Ic, u);
                    }
                } else {
                    u[0] = (double)r[o0];
                    u[1] = (double)r[o0 + 1];
                    u[2] = (double)r[o0 + 2];
                }
                if (step_sums) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) u[c] -= step_scale * __ldg(step_sums + (size_t)c * pitch + t);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) val[e][c] = u[c];
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
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double pre = carry[c], tot = 0.0;
#pragma unroll
                for (int wdx = 0; wdx < 8; ++wdx) {
                    const double x = s_tot[c][wdx];
                    if (wdx < warp) pre += x;
                    tot += x;
                }
                const double off = pre + (incl[c] - val[E - 1][c]);
                carry[c] += tot;
#pragma unroll
                for (int e = 0; e < E; ++e) val[e][c] += off;
            }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int f = f0 + e;
            const bool valid = (t0 + f) < Tn;
#pragma unroll
            for (int c = 0; c < 3; ++c) outs[c * uw_out_elems() + f + f / E] = valid ? val[e][c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double* orow = out + ((size_t)atom * 3 + c) * pitch;
            for (int f = tid; f < TF; f += 256) {
                if (t0 + f < pitch) orow[t0 + f] = outs[c * uw_out_elems() + f + f / E];
            }
        }
        // the next tile overwrites s_tot / outs only after its own barriers
    }
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
static int launch_unwrap(const void* coords, int F, const double* latt, const double* inv, int lstride,
                         const int32_t* idx, int n_idx, const double* step_sums, double step_scale, double* out,
                         int64_t pitch, int Tn, cudaStream_t st) {
    const T* c = (const T*)coords;
    const size_t smem = uw_smem_bytes<T>();
    KB2_CUDA(cudaFuncSetAttribute(unwrap_cumsum_kernel<T, MODE, CONSTL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    unwrap_cumsum_kernel<T, MODE, CONSTL><<<n_idx, 256, smem, st>>>(c, F, latt, inv, lstride, idx, n_idx, step_sums,
                                                                    step_scale, out, pitch, Tn);
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

extern "C" int kb2_frame_sums(const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in,
                              const double* latt, const double* latt_inv, int latt_stride, const int32_t* idx,
                              const double* w, int n_idx, double* frame_sums, int64_t pitch, void* stream) {
    int Tn = 0;
    if (int rc = check_stage1("kb2_frame_sums", coords, dtype, mode, n_atoms_total, n_frames_in, latt, latt_inv,
                              latt_stride, idx, n_idx, pitch, &Tn))
        return rc;
    KB2_CHECK(frame_sums, "kb2_frame_sums: null output");
    cudaStream_t st = (cudaStream_t)stream;
    KB2_CUDA(cudaMemsetAsync(frame_sums, 0, sizeof(double) * 3 * pitch, st));
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const int tiles = ceil_div(Tn, 256);
    int chunks = ceil_div((int64_t)sm * 8, tiles);
    chunks = max(1, min(chunks, ceil_div(n_idx, 8)));
    chunks = min(chunks, 65535);
    const int per = ceil_div(n_idx, chunks);
    chunks = ceil_div(n_idx, per);
    dim3 grid(tiles, chunks);
#define KB2_FS(T, M, C)                                                                                             \
    frame_sums_kernel<T, M, C><<<grid, 256, 0, st>>>((const T*)coords, n_frames_in, latt, latt_inv, latt_stride, idx, w, \
                                                     n_idx, per, frame_sums, pitch, Tn)
    const bool constl = (mode == KB2_MODE_FRACTIONAL && latt_stride == 0);
    if (dtype == KB2_F32 && mode == KB2_MODE_FRACTIONAL) { if (constl) KB2_FS(float, KB2_MODE_FRACTIONAL, true); else KB2_FS(float, KB2_MODE_FRACTIONAL, false); }
    else if (dtype == KB2_F32) KB2_FS(float, KB2_MODE_DISPLACEMENT, false);
    else if (mode == KB2_MODE_FRACTIONAL) { if (constl) KB2_FS(double, KB2_MODE_FRACTIONAL, true); else KB2_FS(double, KB2_MODE_FRACTIONAL, false); }
    else KB2_FS(double, KB2_MODE_DISPLACEMENT, false);
#undef KB2_FS
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_scan_frames(double* buf, int n_frames, int64_t pitch, int cumulative, double divide,
                               const double* sub, double sub_scale, void* stream) {
    KB2_CHECK(buf && n_frames > 0 && pitch >= n_frames, "kb2_scan_frames: bad arguments");
    KB2_CHECK(divide != 0.0, "kb2_scan_frames: divide must be non-zero");
    scan_frames_kernel<<<3, 1024, 0, (cudaStream_t)stream>>>(buf, n_frames, pitch, cumulative, divide, sub, sub_scale);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_unwrap_cumsum(const void* coords, int dtype, int mode, int n_atoms_total, int n_frames_in,
                                 const double* latt, const double* latt_inv, int latt_stride, const int32_t* idx,
                                 int n_idx, const double* step_sums, double step_scale, double* dc_out, int64_t pitch,
                                 void* stream) {
    int Tn = 0;
    if (int rc = check_stage1("kb2_unwrap_cumsum", coords, dtype, mode, n_atoms_total, n_frames_in, latt, latt_inv,
                              latt_stride, idx, n_idx, pitch, &Tn))
        return rc;
    KB2_CHECK(dc_out, "kb2_unwrap_cumsum: null output");
    cudaStream_t st = (cudaStream_t)stream;
#define KB2_UC(T, M, C) \
    return launch_unwrap<T, M, C>(coords, n_frames_in, latt, latt_inv, latt_stride, idx, n_idx, step_sums, step_scale, dc_out, pitch, Tn, st)
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
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
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
