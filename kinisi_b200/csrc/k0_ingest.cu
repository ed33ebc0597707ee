// Ingest helpers that sit in front of stage 1.
//
// kb2_frames_to_atoms: MD codes write trajectories frame by frame, [frame][atom][3]; the kernels of this path stream
//   one atom over time, [atom][frame][3] (what np.concatenate(coords, axis=1) builds on the host in the reference,
//   kinisi/parser.py:120, 309-325).  The transposition is a tiled shared-memory transpose of 3-vectors on the device,
//   so the front-ends only ever append contiguous frames on the host.  HBM-bound: every element read and written once.
// kb2_molecule_centres: the pseudo-centre-of-mass of groups of atoms in fractional coordinates under periodic boundary
//   conditions (kinisi/parser.py:674-738, arXiv:2501.14578): the circular mean of the members fixes a reference point,
//   the members are re-centred on it, averaged (mass weighted), and the result is wrapped back into the cell.
#include <math.h>

#include "common.cuh"

namespace kb2 {

constexpr int kTrTile = 32;  // atoms x frames per tile

// in [n_frames][n_atoms][3] -> out[(a * F_out + f_off + f) * 3 + k]; dup_first also writes frame 0 to f_off - 1
template <typename T>
__global__ void __launch_bounds__(256) frames_to_atoms_kernel(const T* __restrict__ in, int n_atoms, int n_frames,
                                                              T* __restrict__ out, int64_t F_out, int64_t f_off,
                                                              int dup_first) {
    __shared__ T tile[kTrTile][kTrTile * 3 + 1];  // [frame][atom * 3 + k], padded: conflict-free transposed reads
    const int a0 = blockIdx.x * kTrTile, f0 = blockIdx.y * kTrTile;
    const int na = min(kTrTile, n_atoms - a0), nf = min(kTrTile, n_frames - f0);
    // read: a frame's atoms a0 .. a0 + na are na * 3 contiguous elements
    for (int e = threadIdx.x; e < nf * na * 3; e += 256) {
        const int f = e / (na * 3), r = e - f * (na * 3);
        tile[f][r] = in[((size_t)(f0 + f) * n_atoms + a0) * 3 + r];
    }
    __syncthreads();
    // write: an atom's frames f0 .. f0 + nf are nf * 3 contiguous elements
    for (int e = threadIdx.x; e < na * nf * 3; e += 256) {
        const int a = e / (nf * 3), r = e - a * (nf * 3);
        const int f = r / 3, k = r - f * 3;
        const T v = tile[f][a * 3 + k];
        T* row = out + (size_t)(a0 + a) * F_out * 3;
        row[(f_off + f0 + f) * 3 + k] = v;
        if (dup_first && f0 + f == 0) row[(f_off - 1) * 3 + k] = v;
    }
}

// The MDAnalysis front-end (kinisi/parser.py:620-638) holds CARTESIAN positions and the cell of every frame: the reference
// computes np.dot(positions, inv(cell)) per frame on the host.  Same tiled transpose, with the product by the frame's
// inverse cell (row vector times matrix, float64 whatever the input type) applied on the way out.
template <typename T>
__global__ void __launch_bounds__(256) frames_to_atoms_fractional_kernel(const T* __restrict__ in, int n_atoms, int n_frames,
                                                                         const double* __restrict__ cell_inv, int cell_stride,
                                                                         double* __restrict__ out, int64_t F_out, int64_t f_off,
                                                                         int dup_first) {
    __shared__ T tile[kTrTile][kTrTile * 3 + 1];
    __shared__ double inv[kTrTile][9];
    const int a0 = blockIdx.x * kTrTile, f0 = blockIdx.y * kTrTile;
    const int na = min(kTrTile, n_atoms - a0), nf = min(kTrTile, n_frames - f0);
    for (int e = threadIdx.x; e < nf * na * 3; e += 256) {
        const int f = e / (na * 3), r = e - f * (na * 3);
        tile[f][r] = in[((size_t)(f0 + f) * n_atoms + a0) * 3 + r];
    }
    for (int e = threadIdx.x; e < nf * 9; e += 256) inv[e / 9][e % 9] = cell_inv[(size_t)(f0 + e / 9) * cell_stride + e % 9];
    __syncthreads();
    for (int e = threadIdx.x; e < na * nf * 3; e += 256) {
        const int a = e / (nf * 3), r = e - a * (nf * 3);
        const int f = r / 3, k = r - f * 3;
        const double x = (double)tile[f][a * 3], y = (double)tile[f][a * 3 + 1], z = (double)tile[f][a * 3 + 2];
        const double v = fma(z, inv[f][6 + k], fma(y, inv[f][3 + k], x * inv[f][k]));
        double* row = out + (size_t)(a0 + a) * F_out * 3;
        row[(f_off + f0 + f) * 3 + k] = v;
        if (dup_first && f0 + f == 0) row[(f_off - 1) * 3 + k] = v;
    }
}

// one thread per (molecule, frame): members[mol * m + j] are atom rows of coords [n_atoms][F][3]
template <typename T>
__global__ void __launch_bounds__(128) molecule_centres_kernel(const T* __restrict__ coords, int64_t F,
                                                               const int32_t* __restrict__ members, int n_mol, int m,
                                                               const double* __restrict__ weights,
                                                               double* __restrict__ out) {
    const int64_t f = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int mol = blockIdx.y;
    if (f >= F) return;
    const int32_t* mem = members + (size_t)mol * m;
    const double two_pi = 6.283185307179586476925286766559;
    double xi[3] = {0, 0, 0}, zeta[3] = {0, 0, 0}, wsum = 0.0;
    for (int j = 0; j < m; ++j) {
        const T* p = coords + ((size_t)mem[j] * F + f) * 3;
        const double w = weights ? weights[j] : 1.0;
        wsum += w;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double s, c;
            sincos((double)p[k] * two_pi, &s, &c);  // parser.py:713-715
            xi[k] = fma(w, c, xi[k]);
            zeta[k] = fma(w, s, zeta[k]);
        }
    }
    double ref[3], acc[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // np.average(..., weights): the common 1 / sum(w) does not change the angle but is kept for exact zeros
        const double theta_bar = atan2(-(zeta[k] / wsum), -(xi[k] / wsum)) + 3.14159265358979323846;  // parser.py:718
        ref[k] = theta_bar / two_pi + 0.5;                                                           // new_s_coords + 0.5
    }
    for (int j = 0; j < m; ++j) {
        const T* p = coords + ((size_t)mem[j] * F + f) * 3;
        const double w = weights ? weights[j] : 1.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double d = (double)p[k] - ref[k];
            acc[k] = fma(w, d - floor(d), acc[k]);  // (s - ref) % 1, parser.py:722
        }
    }
    double* o = out + ((size_t)mol * F + f) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double v = acc[k] / wsum + ref[k];
        o[k] = v - floor(v);  // % 1, parser.py:724
    }
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_frames_to_atoms(const void* frames, int dtype, int n_atoms, int n_frames, void* coords_out,
                                   int64_t n_frames_out, int64_t frame_offset, int duplicate_first, void* stream) {
    KB2_CHECK(frames && coords_out, "kb2_frames_to_atoms: null pointer");
    KB2_CHECK(dtype == KB2_F32 || dtype == KB2_F64, "kb2_frames_to_atoms: dtype must be KB2_F32 or KB2_F64");
    KB2_CHECK(n_atoms > 0 && n_frames > 0, "kb2_frames_to_atoms: empty input");
    KB2_CHECK(frame_offset >= (duplicate_first ? 1 : 0) && frame_offset + n_frames <= n_frames_out,
              "kb2_frames_to_atoms: the frames do not fit the output rows");
    dim3 grid(ceil_div(n_atoms, kTrTile), ceil_div(n_frames, kTrTile));
    KB2_CHECK(grid.y <= 65535, "kb2_frames_to_atoms: at most %d frames per call", 65535 * kTrTile);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KB2_F32)
        frames_to_atoms_kernel<float><<<grid, 256, 0, st>>>((const float*)frames, n_atoms, n_frames, (float*)coords_out,
                                                            n_frames_out, frame_offset, duplicate_first);
    else
        frames_to_atoms_kernel<double><<<grid, 256, 0, st>>>((const double*)frames, n_atoms, n_frames, (double*)coords_out,
                                                             n_frames_out, frame_offset, duplicate_first);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_frames_to_atoms_fractional(const void* frames, int dtype, int n_atoms, int n_frames, const double* cell_inv,
                                              int cell_stride, double* coords_out, int64_t n_frames_out, int64_t frame_offset,
                                              int duplicate_first, void* stream) {
    KB2_CHECK(frames && coords_out && cell_inv, "kb2_frames_to_atoms_fractional: null pointer");
    KB2_CHECK(dtype == KB2_F32 || dtype == KB2_F64, "kb2_frames_to_atoms_fractional: dtype must be KB2_F32 or KB2_F64");
    KB2_CHECK(cell_stride == 0 || cell_stride == 9, "kb2_frames_to_atoms_fractional: cell_stride must be 0 or 9");
    KB2_CHECK(n_atoms > 0 && n_frames > 0, "kb2_frames_to_atoms_fractional: empty input");
    KB2_CHECK(frame_offset >= (duplicate_first ? 1 : 0) && frame_offset + n_frames <= n_frames_out,
              "kb2_frames_to_atoms_fractional: the frames do not fit the output rows");
    dim3 grid(ceil_div(n_atoms, kTrTile), ceil_div(n_frames, kTrTile));
    KB2_CHECK(grid.y <= 65535, "kb2_frames_to_atoms_fractional: at most %d frames per call", 65535 * kTrTile);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KB2_F32)
        frames_to_atoms_fractional_kernel<float><<<grid, 256, 0, st>>>((const float*)frames, n_atoms, n_frames, cell_inv,
                                                                       cell_stride, coords_out, n_frames_out, frame_offset,
                                                                       duplicate_first);
    else
        frames_to_atoms_fractional_kernel<double><<<grid, 256, 0, st>>>((const double*)frames, n_atoms, n_frames, cell_inv,
                                                                        cell_stride, coords_out, n_frames_out, frame_offset,
                                                                        duplicate_first);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_molecule_centres(const void* coords, int dtype, int n_atoms_total, int64_t n_frames_in,
                                    const int32_t* members, int n_molecules, int molecule_size, const double* weights,
                                    double* centres, void* stream) {
    KB2_CHECK(coords && members && centres, "kb2_molecule_centres: null pointer");
    KB2_CHECK(dtype == KB2_F32 || dtype == KB2_F64, "kb2_molecule_centres: dtype must be KB2_F32 or KB2_F64");
    KB2_CHECK(n_atoms_total > 0 && n_frames_in > 0 && n_molecules > 0 && molecule_size > 0,
              "kb2_molecule_centres: empty input");
    KB2_CHECK(n_molecules <= 65535, "kb2_molecule_centres: at most 65535 molecules per call");
    dim3 grid(ceil_div(n_frames_in, 128), n_molecules);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KB2_F32)
        molecule_centres_kernel<float><<<grid, 128, 0, st>>>((const float*)coords, n_frames_in, members, n_molecules,
                                                             molecule_size, weights, centres);
    else
        molecule_centres_kernel<double><<<grid, 128, 0, st>>>((const double*)coords, n_frames_in, members, n_molecules,
                                                              molecule_size, weights, centres);
    KB2_LAUNCH_CHECK();
    return 0;
}
