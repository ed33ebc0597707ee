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
// and only V^T applied to three vectors is needed, never V itself.  A cyclic two-sided Jacobi iteration in
// shared memory delivers exactly that: each plane rotation J is applied to the matrix (J^T A J) and to the
// three vectors (J^T u).  n <= 128 (the default 100-point dt grid gives n <= 100); larger problems take the
// library eigensolver on the host side.
//
// SEAT space.  The m = 2h indices sit on the 2h seats of a round-robin tournament: a top row T_0..T_{h-1}
// (seat ids 0..h-1) and a bottom row B_0..B_{h-1} (seat ids h..2h-1).  Slot k rotates the pair (T_k, B_k); after
// every round all seats but T_0 move one step round the circle T_1 -> ... -> T_{h-1} -> B_{h-1} -> ... -> B_0
// -> T_1, so after m - 1 rounds (one sweep) every pair has met once and every index is back on its own seat.
// The matrix is stored by SEAT, so every round has the same data flow: the 2x2 block between slots ka < kb is
// four elements at fixed shared-memory addresses (four planes indexed by the block number), it is rotated by
// (c, s) of the two slots and scattered to four fixed addresses of the other buffer.  Every thread computes its
// byte offsets once; the layout strides are compile-time constants and the rounds are instantiated for both
// buffer parities, so every access of the round loop is [register + immediate].
//
// One barrier per round, three kinds of warps:
//   * warps 0-1, thread k: the "special" block of slot k -- the one block that holds the element which becomes
//     the pivot of slot k in the NEXT round.  The thread rotates it, evaluates the two diagonal entries that
//     arrive on slot k (from the current rotations of the slots they leave) and the slot's next rotation.  This
//     serial chain (~400 cycles, two MUFU) is the critical path of a round and runs beside
//   * warps 2-17: the regular blocks, one task = one column slot kb x two row slots (2j, 2j+1), consecutive lanes
//     = consecutive kb: contiguous 8-byte lanes on every load, (c, s) of kb fetched once for two blocks.  These
//     warps are bound by shared-memory bandwidth (one read and one write of the triangle per round);
//   * the first 4h regular threads also carry the residual pivots (the angle is inexact: the rotated pivot is
//     kept, not dropped) and the three vectors.
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace kb2 {

constexpr int kJacobiMaxN = 128;
constexpr int kJacobiMaxH = kJacobiMaxN / 2;
constexpr int kJacobiSpecial = 64;   // threads of the special warps (>= kJacobiMaxH)
constexpr int kJacobiRegular = 512;  // threads of the regular warps
constexpr int kJacobiThreads = kJacobiSpecial + kJacobiRegular;
constexpr int kJacobiMaxSweeps = 40;
constexpr int kJacobiMinH = 3;  // smaller problems are padded with zero rows (their rotations are the identity)

// Shared-memory image of the matrix in one buffer, in doubles.
namespace jl {
constexpr int kPlane = 1024;                     // >= number of tasks; plane (q, r) holds element q of block r of every task
constexpr int kPiv = 8 * kPlane;                 // [kJacobiMaxH] pivots (T_k, B_k)
constexpr int kDiagT = kPiv + kJacobiMaxH;       // [kJacobiMaxH] diagonal entries of the top seats
constexpr int kDiagB = kDiagT + kJacobiMaxH;     // [kJacobiMaxH] ... of the bottom seats
constexpr int kCS = kDiagB + kJacobiMaxH;        // [kJacobiMaxH][2] rotation (c, s) of every slot
constexpr int kU = kCS + 2 * kJacobiMaxH;        // [3][2 kJacobiMaxH] vectors: top seats, then bottom seats
constexpr int kBuf = kU + 6 * kJacobiMaxH;       // 8896 doubles = 71168 bytes
}  // namespace jl

__device__ __forceinline__ double block_sum_all(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kJacobiThreads / 32; ++w) t += sh[w];
    return t;
}

// Rotation for the pivot block [[app, apq], [apq, aqq]]: with d = aqq - app, h = 2 apq, r = sqrt(d^2 + h^2) the
// small-angle solution has cos 2t = |d| / r, sin 2t = sign(d) h / r, c = sqrt((1 + cos 2t) / 2), s = sin 2t / (2 c).
// The angle only has to be good enough for the iteration to converge (an error of 1e-7 leaves a residual pivot of
// 1e-7 |apq| instead of zero, and the residual is kept); what must hold to FP64 rounding is c^2 + s^2 = 1.  The
// chain sits on the critical path of every round, so (c0, s0) are formed in single precision (two MUFU.RSQ) on
// operands rescaled by an exact power of two and then renormalised in double precision with the series
// (1 + e)^-1/2 = 1 - e/2 + 3 e^2 / 8, e = c0^2 + s0^2 - 1 ~ 1e-7 (error ~e^3).
__device__ __forceinline__ float rsqrt_approx(float v) {  // MUFU.RSQ without the denormal rescue (operands are O(1))
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void jacobi_angle(double app, double aqq, double apq, double& c, double& s) {
    const double d = aqq - app, h = apq + apq;
    // 2^-E with E the larger exponent of d and h: an exact rescale of (d, h) into (-2, 2)
    const int ed = __double2hiint(d) & 0x7ff00000, eh = __double2hiint(h) & 0x7ff00000;
    const double scale = __hiloint2double(0x7fe00000 - max(ed, eh), 0);
    const float df = (float)(d * scale), hf = (float)(h * scale);
    const float ir = rsqrt_approx(fmaf(df, df, hf * hf));
    const float x = fabsf(df) * ir;                                                                // cos 2t
    const float y = __int_as_float(__float_as_int(hf * ir) ^ (__float_as_int(df) & 0x80000000));  // sin 2t
    const float w = fmaf(0.5f, x, 0.5f);                                                           // cos^2 t
    const float ic = rsqrt_approx(w);
    const double c0 = (double)(w * ic), s0 = (double)(0.5f * y * ic);
    const double e = fma(c0, c0, fma(s0, s0, -1.0));
    const double g = fma(e, fma(e, 0.375, -0.5), 1.0);
    const bool rot = apq != 0.0;
    c = rot ? c0 * g : 1.0;
    s = rot ? s0 * g : 0.0;
}

struct JacobiSeats {
    int h;
    __device__ __forceinline__ int next(int p) const {  // where the index on seat p sits in the next round
        if (p == 0) return 0;                 // T_0 is fixed
        if (p < h - 1) return p + 1;          // T_k -> T_{k+1}
        if (p == h - 1) return 2 * h - 1;     // T_{h-1} -> B_{h-1}
        if (p == h) return 1;                 // B_0 -> T_1
        return p - 1;                         // B_k -> B_{k-1}
    }
    __device__ __forceinline__ int prev(int p) const {  // inverse of next
        if (p == 0) return 0;
        if (p == 1) return h;                 // T_1 <- B_0
        if (p < h) return p - 1;              // T_k <- T_{k-1}
        if (p == 2 * h - 1) return h - 1;     // B_{h-1} <- T_{h-1}
        return p + 1;                         // B_k <- B_{k+1}
    }
    __device__ __forceinline__ int slot(int p) const { return p < h ? p : p - h; }
    __device__ __forceinline__ int bottom(int p) const { return p >= h; }
    // task = (row pair j, column slot kb > 2j): blocks (2j, kb) and (2j + 1, kb); tasks are numbered row pair by row pair
    __device__ __forceinline__ int task(int j, int kb) const { return j * (h - j) + (kb - 2 * j - 1); }
    __device__ __forceinline__ int n_tasks() const { const int J = h / 2; return J * (h - J); }
    // offset (doubles) of element TT of block (ka, kb); element q = (row bottom, column bottom) is 2 q planes further
    __device__ __forceinline__ int block(int ka, int kb) const { return (ka & 1) * jl::kPlane + task(ka >> 1, kb); }
    // Offset (doubles) of the element between seats x != y: for seats of different slots the planes are indexed by the
    // task number (consecutive lanes = consecutive tasks = consecutive addresses); kPiv + k for the pivot of slot k.
    __device__ __forceinline__ int addr(int x, int y) const {
        int sx = slot(x), sy = slot(y), tx = bottom(x), ty = bottom(y);
        if (sx == sy) return jl::kPiv + sx;
        if (sx > sy) {
            int q = sx; sx = sy; sy = q;
            q = tx; tx = ty; ty = q;
        }
        return (tx * 2 + ty) * 2 * jl::kPlane + block(sx, sy);
    }
    __device__ __forceinline__ int diag(int p) const { return (bottom(p) ? jl::kDiagB : jl::kDiagT) + slot(p); }
    __device__ __forceinline__ int vec(int v, int p) const { return jl::kU + v * 2 * kJacobiMaxH + bottom(p) * kJacobiMaxH + slot(p); }
    // destination of element q = (row bottom, column bottom) of block (ka, kb)
    __device__ __forceinline__ int block_dst(int ka, int kb, int q) const {
        return addr(next(ka + (q >> 1) * h), next(kb + (q & 1) * h));
    }
    // a block is special when one of its elements lands on a pivot
    __device__ __forceinline__ bool special(int ka, int kb) const {
        bool sp = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) sp = sp || block_dst(ka, kb, q) >= jl::kPiv;
        return sp;
    }
};

__device__ __forceinline__ double lds_f64(const unsigned char* sm, uint32_t off) {
    return *reinterpret_cast<const double*>(sm + off);
}
__device__ __forceinline__ double2 lds_f64x2(const unsigned char* sm, uint32_t off) {
    return *reinterpret_cast<const double2*>(sm + off);
}
__device__ __forceinline__ void sts_f64(unsigned char* sm, uint32_t off, double v) { *reinterpret_cast<double*>(sm + off) = v; }

// M <- J_a^T M J_b for the 2x2 block (TT, TB, BT, BB) between row slot a and column slot b
__device__ __forceinline__ void rotate_block(const double (&mm)[4], double2 ra, double2 rb, double (&ev)[4]) {
    const double t00 = rb.x * mm[0] - rb.y * mm[1], t01 = rb.y * mm[0] + rb.x * mm[1];
    const double t10 = rb.x * mm[2] - rb.y * mm[3], t11 = rb.y * mm[2] + rb.x * mm[3];
    ev[0] = ra.x * t00 - ra.y * t10;
    ev[1] = ra.x * t01 - ra.y * t11;
    ev[2] = ra.y * t00 + ra.x * t10;
    ev[3] = ra.y * t01 + ra.x * t11;
}

template <int NTASK>
__global__ void __launch_bounds__(kJacobiThreads, 1)
    jacobi_project_kernel(const double* __restrict__ cov, int n, int h, const double* __restrict__ x,
                          const double* __restrict__ y, double cond_max, double* __restrict__ evals,
                          double* __restrict__ proj, double* __restrict__ logdet_term, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smj[];
    __shared__ double red[kJacobiThreads / 32];
    double* const buf0 = reinterpret_cast<double*>(smj);
    const int m = 2 * h;             // seats; indices >= n are zero padding
    const JacobiSeats seats{h};
    const int tid = threadIdx.x;
    constexpr uint32_t kBufBytes = jl::kBuf * 8, kPlaneBytes = jl::kPlane * 8;

    // ---- load: seat p holds index p -------------------------------------------------------------------------------
    for (int e = tid; e < 2 * jl::kBuf; e += kJacobiThreads) buf0[e] = 0.0;
    __syncthreads();
    for (int e = tid; e < m * m; e += kJacobiThreads) {
        const int i = e / m, j = e - i * m;
        if (i < j)
            buf0[seats.addr(i, j)] = (j < n) ? cov[(size_t)i * n + j] : 0.0;
        else if (i == j)
            buf0[seats.diag(i)] = (i < n) ? cov[(size_t)i * n + i] : 0.0;
    }
    for (int i = tid; i < m; i += kJacobiThreads) {
        buf0[seats.vec(0, i)] = i < n ? x[i] : 0.0;
        buf0[seats.vec(1, i)] = i < n ? 1.0 : 0.0;
        buf0[seats.vec(2, i)] = i < n ? y[i] : 0.0;
    }

    // ---- static roles (byte offsets into buffer 0) --------------------------------------------------------------------
    // A thread is either special or regular, so both roles share the offset registers O[0][.]:
    //   special thread k: block (ka, kb) holding the element {X, Y}, X = prev(T_k) -> T_k, Y = prev(B_k) -> B_k
    //     O = {block, cs ka, cs kb, diag own a, diag other a, pivot a, diag own b, diag other b, pivot b, dst[4]}
    //   regular thread: task o = (row pair j, column slot kb) -> blocks (2j, kb) and (2j + 1, kb)
    //     O = {task, cs ka0, cs kb, -, cs ka1, dst 0 [4], dst 1 [4]}
    uint32_t O[NTASK][13];
    bool ok[NTASK][2];
#pragma unroll
    for (int o = 0; o < NTASK; ++o) {
        ok[o][0] = ok[o][1] = false;
#pragma unroll
        for (int i = 0; i < 13; ++i) O[o][i] = 0;
    }
    int sp_q = 0;
    bool sp_x_in_ka = true;
    double sp_sga = 0.0, sp_sgb = 0.0;
    const bool is_special = tid < h;
    const int rt = tid - kJacobiSpecial;
    if (is_special) {
        const int X = seats.prev(tid), Y = seats.prev(h + tid);
        const int sX = seats.slot(X), sY = seats.slot(Y);
        const int ka = min(sX, sY), kb = max(sX, sY);
        sp_x_in_ka = sX < sY;
        const int fa = sp_x_in_ka ? seats.bottom(X) : seats.bottom(Y);  // is the element's row seat (slot ka) a bottom seat
        const int fb = sp_x_in_ka ? seats.bottom(Y) : seats.bottom(X);  // ... its column seat (slot kb)
        sp_q = fa * 2 + fb;
        O[0][0] = 8u * seats.block(ka, kb);
        O[0][1] = 8u * (jl::kCS + 2 * ka);
        O[0][2] = 8u * (jl::kCS + 2 * kb);
        // new diagonal of a seat after its slot's rotation: cc * own + ss * other -/+ 2 s c pivot (top / bottom)
        O[0][3] = 8u * ((fa ? jl::kDiagB : jl::kDiagT) + ka);
        O[0][4] = 8u * ((fa ? jl::kDiagT : jl::kDiagB) + ka);
        O[0][5] = 8u * (jl::kPiv + ka);
        sp_sga = fa ? 2.0 : -2.0;
        O[0][6] = 8u * ((fb ? jl::kDiagB : jl::kDiagT) + kb);
        O[0][7] = 8u * ((fb ? jl::kDiagT : jl::kDiagB) + kb);
        O[0][8] = 8u * (jl::kPiv + kb);
        sp_sgb = fb ? 2.0 : -2.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) O[0][9 + q] = 8u * seats.block_dst(ka, kb, q);
    } else if (rt >= 0) {
#pragma unroll
        for (int o = 0; o < NTASK; ++o) {
            int li = rt + o * kJacobiRegular, j = 0;
            while (2 * j < h - 1 && li >= h - 1 - 2 * j) {
                li -= h - 1 - 2 * j;
                ++j;
            }
            if (2 * j < h - 1) {
                const int kb = 2 * j + 1 + li;
                O[o][0] = 8u * seats.task(j, kb);
                O[o][2] = 8u * (jl::kCS + 2 * kb);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int ka = 2 * j + r;
                    if (ka < kb && !seats.special(ka, kb)) {
                        ok[o][r] = true;
                        O[o][3 * r + 1] = 8u * (jl::kCS + 2 * ka);
#pragma unroll
                        for (int q = 0; q < 4; ++q) O[o][5 + 4 * r + q] = 8u * seats.block_dst(ka, kb, q);
                    }
                }
            }
        }
    }
    // auxiliary task of regular thread rt < 4h: the residual pivot of slot rt (rt < h), or the pair (T_k, B_k) of vector v
    const int aux_kind = (rt >= 0 && rt < h) ? 1 : ((rt >= h && rt < 4 * h) ? 2 : 0);
    uint32_t aux_cs = 0, aux_l0 = 0, aux_l1 = 0, aux_l2 = 0, aux_d0 = 0, aux_d1 = 0;
    if (aux_kind == 1) {
        aux_cs = 8u * (jl::kCS + 2 * rt);
        aux_l0 = 8u * (jl::kDiagT + rt);
        aux_l1 = 8u * (jl::kDiagB + rt);
        aux_l2 = 8u * (jl::kPiv + rt);
        aux_d0 = 8u * seats.addr(seats.next(rt), seats.next(h + rt));
    } else if (aux_kind == 2) {
        const int v = (rt - h) / h, k = (rt - h) - v * h;
        aux_cs = 8u * (jl::kCS + 2 * k);
        aux_l0 = 8u * seats.vec(v, k);
        aux_l1 = 8u * seats.vec(v, h + k);
        aux_l2 = aux_l0;
        aux_d0 = 8u * seats.vec(v, seats.next(k));
        aux_d1 = 8u * seats.vec(v, seats.next(h + k));
    }
    __syncthreads();
    // rotations of the first round
    if (tid < h) {
        double c, s;
        jacobi_angle(buf0[jl::kDiagT + tid], buf0[jl::kDiagB + tid], buf0[jl::kPiv + tid], c, s);
        buf0[jl::kCS + 2 * tid] = c;
        buf0[jl::kCS + 2 * tid + 1] = s;
    }
    __syncthreads();

    // ---- one round: reads buffer P, writes buffer 1 - P ------------------------------------------------------------------
    // Loads are unconditional (offset 0 where a thread has no block) and come first: the compiler cannot move them
    // above the scattered stores.
    auto round = [&](auto parity) __attribute__((always_inline)) {
        constexpr int P = decltype(parity)::value;
        constexpr uint32_t R = P * kBufBytes, W = (1 - P) * kBufBytes;
        if (tid < kJacobiSpecial) {
            const double mm[4] = {lds_f64(smj, O[0][0] + R), lds_f64(smj, O[0][0] + R + 2 * kPlaneBytes),
                                  lds_f64(smj, O[0][0] + R + 4 * kPlaneBytes), lds_f64(smj, O[0][0] + R + 6 * kPlaneBytes)};
            const double2 ra = lds_f64x2(smj, O[0][1] + R), rb = lds_f64x2(smj, O[0][2] + R);
            const double a1 = lds_f64(smj, O[0][3] + R), a2 = lds_f64(smj, O[0][4] + R), ap = lds_f64(smj, O[0][5] + R);
            const double b1 = lds_f64(smj, O[0][6] + R), b2 = lds_f64(smj, O[0][7] + R), bp = lds_f64(smj, O[0][8] + R);
            double ev[4];
            rotate_block(mm, ra, rb, ev);
            const double apq = (sp_q & 2) ? ((sp_q & 1) ? ev[3] : ev[2]) : ((sp_q & 1) ? ev[1] : ev[0]);
            const double da = fma(ra.x * ra.x, a1, fma(ra.y * ra.y, a2, sp_sga * (ra.x * ra.y) * ap));
            const double db = fma(rb.x * rb.x, b1, fma(rb.y * rb.y, b2, sp_sgb * (rb.x * rb.y) * bp));
            const double app = sp_x_in_ka ? da : db, aqq = sp_x_in_ka ? db : da;
            double c, s;
            jacobi_angle(app, aqq, apq, c, s);
            if (is_special) {
                *reinterpret_cast<double2*>(smj + W + 8u * (jl::kCS + 2 * tid)) = make_double2(c, s);
                sts_f64(smj, W + 8u * (jl::kDiagT + tid), app);
                sts_f64(smj, W + 8u * (jl::kDiagB + tid), aqq);
#pragma unroll
                for (int q = 0; q < 4; ++q) sts_f64(smj, O[0][9 + q] + W, ev[q]);
            }
        } else {
            double mm[NTASK][2][4];
            double2 ra[NTASK][2], rb[NTASK];
#pragma unroll
            for (int o = 0; o < NTASK; ++o) {
                rb[o] = lds_f64x2(smj, O[o][2] + R);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    ra[o][r] = lds_f64x2(smj, O[o][3 * r + 1] + R);
#pragma unroll
                    for (int q = 0; q < 4; ++q) mm[o][r][q] = lds_f64(smj, O[o][0] + R + (2 * q + r) * kPlaneBytes);
                }
            }
            const double2 ac = lds_f64x2(smj, aux_cs + R);
            const double l0 = lds_f64(smj, aux_l0 + R), l1 = lds_f64(smj, aux_l1 + R), l2 = lds_f64(smj, aux_l2 + R);
#pragma unroll
            for (int o = 0; o < NTASK; ++o) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    double ev[4];
                    rotate_block(mm[o][r], ra[o][r], rb[o], ev);
                    if (ok[o][r]) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) sts_f64(smj, O[o][5 + 4 * r + q] + W, ev[q]);
                    }
                }
            }
            if (aux_kind == 1) {
                // J^T [[app, apq], [apq, aqq]] J, off-diagonal element
                sts_f64(smj, aux_d0 + W, fma(ac.y * ac.x, l0 - l1, (ac.x * ac.x - ac.y * ac.y) * l2));
            } else if (aux_kind == 2) {
                sts_f64(smj, aux_d0 + W, ac.x * l0 - ac.y * l1);
                sts_f64(smj, aux_d1 + W, ac.y * l0 + ac.x * l1);
            }
        }
        __syncthreads();
    };

    int par = 0;  // buffer holding the current matrix
    int sweeps = 0;
    bool polishing = false;
    for (; sweeps < kJacobiMaxSweeps; ++sweeps) {
        // Once the off-diagonal norm is <= 1e-9 of the matrix norm the iteration converges quadratically: ONE more
        // sweep takes it to rounding level.
        const double* cur = buf0 + par * jl::kBuf;
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < seats.n_tasks(); e += kJacobiThreads) {
#pragma unroll
            for (int q = 0; q < 8; ++q) off = fma(cur[q * jl::kPlane + e], cur[q * jl::kPlane + e], off);
        }
        if (tid < h) {
            off = fma(cur[jl::kPiv + tid], cur[jl::kPiv + tid], off);
            dg = fma(cur[jl::kDiagT + tid], cur[jl::kDiagT + tid], cur[jl::kDiagB + tid] * cur[jl::kDiagB + tid]);
        }
        off = block_sum_all(off, red);
        dg = block_sum_all(dg, red);
        if (polishing || off == 0.0) break;
        if (off <= 1e-18 * (dg + 2.0 * off)) polishing = true;
        // m - 1 rounds (an odd number): the buffers swap roles from sweep to sweep
        if (par == 0) {
            for (int r = 0; r < h - 1; ++r) {
                round(std::integral_constant<int, 0>{});
                round(std::integral_constant<int, 1>{});
            }
            round(std::integral_constant<int, 0>{});
        } else {
            for (int r = 0; r < h - 1; ++r) {
                round(std::integral_constant<int, 1>{});
                round(std::integral_constant<int, 0>{});
            }
            round(std::integral_constant<int, 1>{});
        }
        par ^= 1;
    }

    // ---- spectrum -> clipped spectrum -> pinvh cut-off -> whitened projections and log-determinant ------------
    // (whole sweeps only: every index is back on its own seat, seat p = index p; seats >= n are padding)
    const double* cur = buf0 + par * jl::kBuf;
    double lmax = -INFINITY;
    for (int i = tid; i < n; i += kJacobiThreads) lmax = fmax(lmax, cur[seats.diag(i)]);
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
        const double raw = cur[seats.diag(r)];
        const double lam = fmax(raw, floor_l);  // diffusion.py:887
        evals[r] = raw;
        ld_sum += log(lam);
        const bool keep = fabs(lam) > cutoff;
        const double s = keep ? 1.0 / sqrt(fabs(lam)) : 0.0;
        proj[r] = s * cur[seats.vec(0, r)];
        proj[n + r] = s * cur[seats.vec(1, r)];
        proj[2 * n + r] = s * cur[seats.vec(2, r)];
        proj[3 * n + r] = keep ? (lam > 0.0 ? 1.0 : -1.0) : 0.0;
    }
    ld_sum = block_sum_all(ld_sum, red);
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
    const int h = (n + 1) / 2 < kJacobiMinH ? kJacobiMinH : (n + 1) / 2;
    const size_t smem = sizeof(double) * 2 * (size_t)jl::kBuf;
    int n_tasks = 0;  // (row pair, column) tasks
    for (int j = 0; 2 * j < h - 1; ++j) n_tasks += h - 1 - 2 * j;
    if (n_tasks <= kJacobiRegular) {
        KB2_CUDA(cudaFuncSetAttribute(jacobi_project_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jacobi_project_kernel<1><<<1, kJacobiThreads, smem, st>>>(cov, n, h, x, y, cond_max, evals, proj, logdet_term, info);
    } else {
        KB2_CUDA(cudaFuncSetAttribute(jacobi_project_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jacobi_project_kernel<2><<<1, kJacobiThreads, smem, st>>>(cov, n, h, x, y, cond_max, evals, proj, logdet_term, info);
    }
    KB2_LAUNCH_CHECK();
    return launch_gls_solve(proj, n, fit_intercept, gls, st);
}
