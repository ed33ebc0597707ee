// Stage 3, spectral step: everything the regression needs from the model covariance, in one CTA.
//
// The reference reconditions the covariance by clipping its spectrum (minimum_eigenvalue_method,
// kinisi/diffusion.py:870-888), repairs it (cov_nearest, :425), takes slogdet (:350) and inverts it
// with scipy.linalg.pinvh (:352), i.e. through its spectrum again with a relative cut-off.  All of
// that is a function of ONE symmetric eigendecomposition C = V L V^T:
//     clipped spectrum   l'_r = max(l_r, l_max / cond_max)
//     kept directions    |l'_r| > n * eps * max|l'|               (pinvh)
//     whitened model     a = S V^T dt, b = S V^T 1, w = S V^T msd,  S = diag(|l'_r|^-1/2)
//     log-determinant    sum_r log l'_r
// and only V^T applied to three vectors is needed, never V itself.  A cyclic two-sided Jacobi
// iteration in shared memory delivers exactly that: each plane rotation J is applied to the matrix
// (J^T A J, symmetric storage, fused 2x2 block updates) and to the three vectors (J^T u), so the cost
// per round is one pass over the upper triangle instead of the n^2/2 rotations an eigenvector matrix
// would add.  n <= 128 (the default 100-point dt grid gives n <= 100); larger problems take the
// library eigensolver on the host side.  Replaces ~450 cuSOLVER launches (1.45 ms at n = 90).
#include <math.h>

#include "common.cuh"

namespace kb2 {

constexpr int kJacobiMaxN = 128;
constexpr int kJacobiThreads = 512;
constexpr int kJacobiMaxSweeps = 40;

__device__ __forceinline__ double block_sum_512(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kJacobiThreads / 32; ++w) t += sh[w];
    return t;
}

// Round-robin tournament: pair k of round r among m (even) indices, returned with p < q.
__device__ __forceinline__ void tournament_pair(int k, int r, int m, int& p, int& q) {
    int a, b;
    if (k == 0) {
        a = m - 1;
        b = r;
    } else {
        a = (r + k) % (m - 1);
        b = (r + (m - 1) - k) % (m - 1);
    }
    p = min(a, b);
    q = max(a, b);
}

// Rotation angle for the pivot block [[app, apq], [apq, aqq]].  The tangent is evaluated in single
// precision on operands rescaled by an exact power of two (FP64 division and square root are long
// software sequences and sit on the critical path of every round); c and s are then formed in double
// precision from that tangent, so the rotation is orthogonal to FP64 rounding and merely leaves a
// residual of ~1e-7 |apq| in the pivot instead of zero.  Jacobi converges with inexact angles: the
// off-diagonal norm still shrinks by >= 1e-7 per sweep on top of the quadratic term.
__device__ __forceinline__ void jacobi_angle(double app, double aqq, double apq, double& c, double& s) {
    const double d = aqq - app, h = 2.0 * apq;
    const double mx = fmax(fabs(d), fabs(h));
    // 2^-e with e the unbiased exponent of mx (exact rescale into [1, 2))
    const long long bits = __double_as_longlong(mx);
    const long long expo = (bits >> 52) & 0x7ff;
    const double scale = __longlong_as_double((long long)(2046 - expo) << 52);
    const float df = (float)(d * scale), hf = (float)(h * scale);
    // t = h / (d + sign(d) sqrt(d^2 + h^2)): the smaller root, |t| <= 1
    const float rf = sqrtf(fmaf(df, df, hf * hf));
    const float tf = hf / (df + copysignf(rf, df));
    const double t = (double)tf;
    c = rsqrt(fma(t, t, 1.0));
    s = t * c;
}

__global__ void __launch_bounds__(kJacobiThreads)
    jacobi_project_kernel(const double* __restrict__ cov, int n, const double* __restrict__ x,
                          const double* __restrict__ y, double cond_max, double* __restrict__ evals,
                          double* __restrict__ proj, double* __restrict__ logdet_term, int* __restrict__ info) {
    extern __shared__ __align__(16) double smj[];
    __shared__ double red[kJacobiThreads / 32];
    const int m = (n + 1) & ~1;  // even number of indices (a zero dummy is appended when n is odd)
    const int ld = m | 1;        // odd leading dimension: column walks are bank-conflict free
    double* A = smj;             // [m][ld]; element (i, j) lives at A[min(i,j)][max(i,j)]
    double* u = A + (size_t)m * ld;  // [3][m]: dt, 1, msd -> rotated into the eigenbasis
    double* cs = u + 3 * m;          // [m/2][2] rotation (c, s) of each pair of the current round
    int* prs = reinterpret_cast<int*>(cs + m);  // [m/2][2] the pairs (p, q) of the current round
    const int tid = threadIdx.x;
    const int half = m / 2;

    for (int e = tid; e < m * m; e += kJacobiThreads) {
        const int i = e / m, j = e - i * m;
        A[i * ld + j] = (i < n && j < n) ? cov[(size_t)i * n + j] : 0.0;
    }
    for (int i = tid; i < m; i += kJacobiThreads) {
        u[i] = i < n ? x[i] : 0.0;
        u[m + i] = i < n ? 1.0 : 0.0;
        u[2 * m + i] = i < n ? y[i] : 0.0;
    }
    // The off-diagonal 2x2 blocks (ka < kb) are the same index set every round: each thread owns up to
    // kMaxOwn of them for the whole run.
    constexpr int kMaxOwn = (kJacobiMaxN / 2) * (kJacobiMaxN / 2 - 1) / 2 / kJacobiThreads + 1;
    const int n_blocks = half * (half - 1) / 2;
    int own_ka[kMaxOwn], own_kb[kMaxOwn];
#pragma unroll
    for (int o = 0; o < kMaxOwn; ++o) {
        int e = tid + o * kJacobiThreads;
        own_ka[o] = -1;
        own_kb[o] = 0;
        if (e < n_blocks) {  // decode the e-th pair (ka < kb) of the upper triangle, row by row
            int ka = 0, row = half - 1;
            while (e >= row) {
                e -= row;
                --row;
                ++ka;
            }
            own_ka[o] = ka;
            own_kb[o] = ka + 1 + e;
        }
    }
    __syncthreads();

    int sweeps = 0;
    bool polishing = false;
    for (; sweeps < kJacobiMaxSweeps; ++sweeps) {
        // Converged when the off-diagonal norm is <= 1e-13 of the matrix norm; ONE more sweep after that
        // takes it to rounding level (the floor of the off-diagonal norm is ~sqrt(n) eps, so a much
        // tighter test could never fire).
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < m * m; e += kJacobiThreads) {
            const int i = e / m, j = e - i * m;
            if (i < j) {
                const double v = A[i * ld + j];
                off = fma(v, v, off);
            } else if (i == j) {
                const double v = A[i * ld + i];
                dg = fma(v, v, dg);
            }
        }
        off = block_sum_512(off, red);
        dg = block_sum_512(dg, red);
        if (polishing || off == 0.0) break;
        if (off <= 1e-26 * (dg + 2.0 * off)) polishing = true;

        for (int r = 0; r < m - 1; ++r) {
            // phase 1: the rotation of every pair (the only serial arithmetic of the round)
            if (tid < half) {
                int p, q;
                tournament_pair(tid, r, m, p, q);
                prs[2 * tid] = p;
                prs[2 * tid + 1] = q;
                const double apq = A[p * ld + q];
                double c = 1.0, s = 0.0;
                if (apq != 0.0) jacobi_angle(A[p * ld + p], A[q * ld + q], apq, c, s);
                cs[2 * tid] = c;
                cs[2 * tid + 1] = s;
            }
            __syncthreads();
            // phase 2, all independent given (c, s): pivot blocks and the three vectors (threads < 4*half) ...
            if (tid < half) {
                const int p = prs[2 * tid], q = prs[2 * tid + 1];
                const double c = cs[2 * tid], s = cs[2 * tid + 1];
                const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
                // J^T [[app, apq], [apq, aqq]] J with J = [[c, s], [-s, c]]
                const double cc = c * c, ss = s * s, sc = s * c;
                A[p * ld + p] = fma(cc, app, fma(ss, aqq, -2.0 * sc * apq));
                A[q * ld + q] = fma(ss, app, fma(cc, aqq, 2.0 * sc * apq));
                A[p * ld + q] = fma(sc, app - aqq, (cc - ss) * apq);
            } else if (tid < 4 * half) {
                const int v = tid / half - 1, k = tid - (v + 1) * half;
                const int p = prs[2 * k], q = prs[2 * k + 1];
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double up = u[v * m + p], uq = u[v * m + q];
                u[v * m + p] = c * up - s * uq;
                u[v * m + q] = s * up + c * uq;
            }
            // ... and the off-diagonal 2x2 blocks (pair ka rows) x (pair kb columns), ka < kb:  M <- J_ka^T M J_kb
#pragma unroll
            for (int o = 0; o < kMaxOwn; ++o) {
                const int ka = own_ka[o], kb = own_kb[o];
                if (ka < 0) continue;
                const int p1 = prs[2 * ka], q1 = prs[2 * ka + 1], p2 = prs[2 * kb], q2 = prs[2 * kb + 1];
                const double c1 = cs[2 * ka], s1 = cs[2 * ka + 1], c2 = cs[2 * kb], s2 = cs[2 * kb + 1];
                double* e00 = A + min(p1, p2) * ld + max(p1, p2);
                double* e01 = A + min(p1, q2) * ld + max(p1, q2);
                double* e10 = A + min(q1, p2) * ld + max(q1, p2);
                double* e11 = A + min(q1, q2) * ld + max(q1, q2);
                const double m00 = *e00, m01 = *e01, m10 = *e10, m11 = *e11;
                // columns: (x_p, x_q) <- (c x_p - s x_q, s x_p + c x_q) with the column pair's rotation
                const double t00 = c2 * m00 - s2 * m01, t01 = s2 * m00 + c2 * m01;
                const double t10 = c2 * m10 - s2 * m11, t11 = s2 * m10 + c2 * m11;
                // rows: the same map with the row pair's rotation
                *e00 = c1 * t00 - s1 * t10;
                *e10 = s1 * t00 + c1 * t10;
                *e01 = c1 * t01 - s1 * t11;
                *e11 = s1 * t01 + c1 * t11;
            }
            __syncthreads();
        }
    }

    // spectrum -> clipped spectrum -> pinvh cut-off -> whitened projections and log-determinant
    double lmax = -INFINITY;
    for (int i = tid; i < n; i += kJacobiThreads) lmax = fmax(lmax, A[i * ld + i]);
#pragma unroll
    for (int msk = 16; msk > 0; msk >>= 1) lmax = fmax(lmax, shfl_xor_f64(lmax, msk));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = lmax;
    __syncthreads();
    lmax = red[0];
#pragma unroll
    for (int w = 1; w < kJacobiThreads / 32; ++w) lmax = fmax(lmax, red[w]);
    const double floor_l = lmax / cond_max;                                // diffusion.py:886
    const double cutoff = (double)n * 2.220446049250313e-16 * fabs(lmax);  // scipy.linalg.pinvh
    double ld_sum = 0.0;
    for (int r = tid; r < n; r += kJacobiThreads) {
        const double raw = A[r * ld + r];
        const double lam = fmax(raw, floor_l);  // diffusion.py:887
        evals[r] = raw;
        ld_sum += log(lam);
        const bool keep = fabs(lam) > cutoff;
        const double s = keep ? 1.0 / sqrt(fabs(lam)) : 0.0;
        proj[r] = s * u[r];
        proj[n + r] = s * u[m + r];
        proj[2 * n + r] = s * u[2 * m + r];
        proj[3 * n + r] = keep ? (lam > 0.0 ? 1.0 : -1.0) : 0.0;
    }
    ld_sum = block_sum_512(ld_sum, red);
    if (tid == 0) {
        logdet_term[0] = ld_sum + (double)n * 1.8378770664093453;  // + n ln(2 pi), diffusion.py:351
        if (info) info[0] = sweeps;
    }
}

// defined in k3_regress.cu
int launch_gls_solve(const double* proj, int n, int fit_intercept, double* gls, cudaStream_t st);

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_spectral_prepare_max_n(void) { return kJacobiMaxN; }

extern "C" int kb2_spectral_prepare(const double* cov, const double* x, const double* y, int n, double cond_max,
                                    int fit_intercept, double* evals, double* proj, double* gls, double* logdet_term,
                                    int* info, void* stream) {
    KB2_CHECK(cov && x && y && evals && proj && gls && logdet_term, "kb2_spectral_prepare: null pointer");
    KB2_CHECK(n >= 1 && n <= kJacobiMaxN, "kb2_spectral_prepare: n must be in [1, %d] (got %d)", kJacobiMaxN, n);
    KB2_CHECK(cond_max > 0.0, "kb2_spectral_prepare: cond_max must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const int m = (n + 1) & ~1, ld = m | 1;
    const size_t smem = sizeof(double) * ((size_t)m * ld + 3 * (size_t)m + (size_t)m) + sizeof(int) * (size_t)m;
    KB2_CUDA(cudaFuncSetAttribute(jacobi_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jacobi_project_kernel<<<1, kJacobiThreads, smem, st>>>(cov, n, x, y, cond_max, evals, proj, logdet_term, info);
    KB2_LAUNCH_CHECK();
    return launch_gls_solve(proj, n, fit_intercept, gls, st);
}
