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
// four elements at fixed shared-memory addresses (four planes indexed by the block number: conflict-free for
// consecutive lanes), it is rotated by (c, s) of the two slots and scattered to four fixed addresses of the
// other buffer.  Each thread computes its source and destination offsets once, before the first sweep.
//
// One barrier per round.  The element that becomes the pivot of slot k in the NEXT round sits in exactly one
// block ("special" block of slot k).  Thread k owns that block: besides rotating it, it evaluates the two
// diagonal entries that arrive on slot k (from the current rotations of the slots they leave) and the slot's
// next rotation, so that serial chain (a few hundred cycles) overlaps with the other warps' block updates,
// which are bound by shared-memory bandwidth (one read and one write of the triangle per round).
#include <math.h>

#include "common.cuh"

namespace kb2 {

constexpr int kJacobiMaxN = 128;
constexpr int kJacobiThreads = 512;
constexpr int kJacobiMaxSweeps = 40;
constexpr int kJacobiMinH = 3;  // smaller problems are padded with zero rows (their rotations are the identity)

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

// Rotation for the pivot block [[app, apq], [apq, aqq]]: t = tan(theta) is the smaller root of
// h t^2 + 2 d t - h = 0 with d = aqq - app, h = 2 apq, i.e. t = h / (d + sign(d) sqrt(d^2 + h^2)).
// The angle only has to be good enough for the iteration to converge (an error of 1e-7 leaves a residual pivot
// of 1e-7 |apq| instead of zero, and the residual is kept, not dropped); what must hold to FP64 rounding is
// c^2 + s^2 = 1.  This chain sits on the critical path of every round, so (c0, s0) are formed in single precision
// (MUFU) on operands rescaled by an exact power of two and then renormalised in double precision with the series
// (1 + e)^-1/2 = 1 - e/2 + 3 e^2 / 8, e = c0^2 + s0^2 - 1 ~ 1e-7 (error ~e^3).
__device__ __forceinline__ void jacobi_angle(double app, double aqq, double apq, double& c, double& s) {
    const double d = aqq - app, h = 2.0 * apq;
    const double mx = fmax(fabs(d), fabs(h));
    // 2^-e with e the unbiased exponent of mx: an exact rescale of (d, h) into [1, 2)
    const int hi = __double2hiint(mx);
    const double scale = __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
    const float df = (float)(d * scale), hf = (float)(h * scale);
    const float q = fmaf(df, df, hf * hf);
    const float rf = q * rsqrtf(q);
    const float tf = __fdividef(hf, df + copysignf(rf, df));
    const float cf = rsqrtf(fmaf(tf, tf, 1.0f));
    const double c0 = (double)cf, s0 = (double)(tf * cf);
    const double e = fma(c0, c0, fma(s0, s0, -1.0));
    const double g = fma(e, fma(e, 0.375, -0.5), 1.0);
    c = c0 * g;
    s = s0 * g;
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
    __device__ __forceinline__ int tri(int ka, int kb) const { return ka * (2 * h - ka - 1) / 2 + (kb - ka - 1); }
    // Offset of the element between seats x != y: plane * NT + block for seats of different slots, 4 NT + k for
    // the pivot of slot k.
    __device__ __forceinline__ int addr(int x, int y, int NT) const {
        int sx = slot(x), sy = slot(y), tx = x >= h, ty = y >= h;
        if (sx == sy) return 4 * NT + sx;
        if (sx > sy) {
            int q = sx; sx = sy; sy = q;
            q = tx; tx = ty; ty = q;
        }
        return (tx * 2 + ty) * NT + tri(sx, sy);
    }
    __device__ __forceinline__ void block_of(int e, int& ka, int& kb) const {  // inverse of tri
        ka = 0;
        int row = h - 1;
        while (e >= row) {
            e -= row;
            --row;
            ++ka;
        }
        kb = ka + 1 + e;
    }
};

// New diagonal entry of a seat of a slot after the slot's rotation J^T [[app, apq], [apq, aqq]] J,
// J = [[c, s], [-s, c]] (rows: T' = c T - s B, B' = s T + c B).
__device__ __forceinline__ double rotated_diag(bool bottom, double app, double aqq, double apq, double c, double s) {
    const double cc = c * c, ss = s * s, sc2 = 2.0 * s * c * apq;
    return bottom ? fma(ss, app, fma(cc, aqq, sc2)) : fma(cc, app, fma(ss, aqq, -sc2));
}

// Shared-memory image of the matrix in one buffer (doubles): elements [4 NT + h] (four planes, then the
// pivots), diagonal [m], rotations [h][2], vectors [3][m].
struct JacobiLayout {
    int NT, h, m;
    __host__ __device__ int elems() const { return 4 * NT + h; }
    __host__ __device__ int off_diag() const { return (elems() + 1) & ~1; }
    __host__ __device__ int off_cs() const { return off_diag() + m; }
    __host__ __device__ int off_u() const { return off_cs() + 2 * h; }
    __host__ __device__ int size() const { return (off_u() + 3 * m + 1) & ~1; }
};

template <int OWN>
__global__ void __launch_bounds__(kJacobiThreads)
    jacobi_project_kernel(const double* __restrict__ cov, int n, int h, const double* __restrict__ x,
                          const double* __restrict__ y, double cond_max, double* __restrict__ evals,
                          double* __restrict__ proj, double* __restrict__ logdet_term, int* __restrict__ info) {
    extern __shared__ __align__(16) double smj[];
    __shared__ double red[kJacobiThreads / 32];
    const int m = 2 * h;             // seats; indices >= n are zero padding
    const int NT = h * (h - 1) / 2;  // blocks ka < kb
    const JacobiSeats seats{h};
    const JacobiLayout L{NT, h, m};
    double* const buf0 = smj;
    double* const buf1 = smj + L.size();
    const int tid = threadIdx.x;

    // ---- load: seat p holds index p ----------------------------------------------------------------------
    for (int e = tid; e < m * m; e += kJacobiThreads) {
        const int i = e / m, j = e - i * m;
        if (i < j)
            buf0[seats.addr(i, j, NT)] = (j < n) ? cov[(size_t)i * n + j] : 0.0;
        else if (i == j)
            buf0[L.off_diag() + i] = (i < n) ? cov[(size_t)i * n + i] : 0.0;
    }
    for (int i = tid; i < m; i += kJacobiThreads) {
        double* u = buf0 + L.off_u();
        u[i] = i < n ? x[i] : 0.0;
        u[m + i] = i < n ? 1.0 : 0.0;
        u[2 * m + i] = i < n ? y[i] : 0.0;
    }

    // ---- static ownership ----------------------------------------------------------------------------------
    // special block of slot k: holds the element {prev(T_k), prev(B_k)}, the pivot of slot k in the next round
    int sp_e = -1, sp_ka = 0, sp_kb = 0, sp_q = 0, sp_dst[4] = {0, 0, 0, 0};
    bool sp_x_bottom = false, sp_y_bottom = false, sp_x_in_ka = true;
    int res_dst = 0;
    if (tid < h) {
        const int X = seats.prev(tid), Y = seats.prev(h + tid);  // X lands on T_tid, Y on B_tid
        const int sX = seats.slot(X), sY = seats.slot(Y);
        sp_ka = min(sX, sY);
        sp_kb = max(sX, sY);
        sp_e = seats.tri(sp_ka, sp_kb);
        sp_x_in_ka = sX < sY;
        sp_x_bottom = X >= h;
        sp_y_bottom = Y >= h;
        const bool row_bottom = sp_x_in_ka ? sp_x_bottom : sp_y_bottom, col_bottom = sp_x_in_ka ? sp_y_bottom : sp_x_bottom;
        sp_q = (row_bottom ? 2 : 0) | (col_bottom ? 1 : 0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            sp_dst[q] = seats.addr(seats.next(sp_ka + (q >> 1) * h), seats.next(sp_kb + (q & 1) * h), NT);
        res_dst = seats.addr(seats.next(tid), seats.next(h + tid), NT);  // where the residual pivot of slot tid goes
    }
    // regular blocks: all the others, dealt to the threads from the last one down so that the threads with a
    // special block get the fewest
    int* const s_special = reinterpret_cast<int*>(buf1);  // [NT] flags, scratch before the first round
    for (int e = tid; e < NT; e += kJacobiThreads) s_special[e] = 0;
    __syncthreads();
    if (tid < h) s_special[sp_e] = 1;
    __syncthreads();
    int own_e[OWN], own_k[OWN], own_dst[OWN][4];
    {
        // list of the regular blocks (warp 0 compacts the flags); thread t takes entries (last - t) + o * threads
        int* const s_list = s_special + NT;  // [NT]
        if (tid < 32) {
            int base = 0;
            for (int e0 = 0; e0 < NT; e0 += 32) {
                const int e = e0 + tid;
                const bool regular = e < NT && s_special[e] == 0;
                const unsigned mask = __ballot_sync(0xffffffffu, regular);
                if (regular) s_list[base + __popc(mask & ((1u << tid) - 1))] = e;
                base += __popc(mask);
            }
        }
        __syncthreads();
        const int n_regular = NT - h;  // the h special blocks are distinct for h >= 3
#pragma unroll
        for (int o = 0; o < OWN; ++o) {
            own_e[o] = -1;
            own_k[o] = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) own_dst[o][q] = 0;
            const int li = (kJacobiThreads - 1 - tid) + o * kJacobiThreads;
            if (li < n_regular) {
                const int e = s_list[li];
                int ka, kb;
                seats.block_of(e, ka, kb);
                own_e[o] = e;
                own_k[o] = ka | (kb << 8);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    own_dst[o][q] = seats.addr(seats.next(ka + (q >> 1) * h), seats.next(kb + (q & 1) * h), NT);
            }
        }
    }
    // the three vectors: task (v, k) rotates (u[v][T_k], u[v][B_k]); given to the threads after the special ones
    const int vt = tid - h;
    const bool has_vec = vt >= 0 && vt < 3 * h;
    const int vec_v = has_vec ? vt / h : 0, vec_k = has_vec ? vt - vec_v * h : 0;
    const int vec_dt = seats.next(vec_k), vec_db = seats.next(h + vec_k);
    __syncthreads();
    // buf1 was scratch: clear what the rounds never write (nothing is read before being written, but keep it clean)
    for (int e = tid; e < L.size(); e += kJacobiThreads) buf1[e] = 0.0;
    __syncthreads();
    // rotations of the first round
    if (tid < h) {
        const double apq = buf0[4 * NT + tid];
        double c = 1.0, s = 0.0;
        if (apq != 0.0) jacobi_angle(buf0[L.off_diag() + tid], buf0[L.off_diag() + h + tid], apq, c, s);
        buf0[L.off_cs() + 2 * tid] = c;
        buf0[L.off_cs() + 2 * tid + 1] = s;
    }
    __syncthreads();

    double* cur = buf0;
    double* nxt = buf1;
    int sweeps = 0;
    bool polishing = false;
    for (; sweeps < kJacobiMaxSweeps; ++sweeps) {
        // Once the off-diagonal norm is <= 1e-9 of the matrix norm the iteration converges quadratically: ONE more
        // sweep takes it to rounding level.
        double off = 0.0, dg = 0.0;
        for (int e = tid; e < L.elems(); e += kJacobiThreads) off = fma(cur[e], cur[e], off);
        if (tid < m) dg = cur[L.off_diag() + tid] * cur[L.off_diag() + tid];
        off = block_sum_512(off, red);
        dg = block_sum_512(dg, red);
        if (polishing || off == 0.0) break;
        if (off <= 1e-18 * (dg + 2.0 * off)) polishing = true;

        for (int r = 0; r < m - 1; ++r) {
            const double* __restrict__ E = cur;
            const double* __restrict__ Dg = cur + L.off_diag();
            const double2* __restrict__ CS = reinterpret_cast<const double2*>(cur + L.off_cs());
            double* __restrict__ En = nxt;
            // ---- loads first (the compiler cannot hoist them above the scattered stores): two regular blocks, the
            // special block with the diagonal / pivot entries its chain needs, the vector pair --------------------
            double bm[2][4];
            double2 b1[2], b2[2];
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                if (own_e[o] >= 0) {
                    const int e = own_e[o];
                    bm[o][0] = E[e];
                    bm[o][1] = E[NT + e];
                    bm[o][2] = E[2 * NT + e];
                    bm[o][3] = E[3 * NT + e];
                    b1[o] = CS[own_k[o] & 0xff];
                    b2[o] = CS[own_k[o] >> 8];
                }
            }
            double sm4[4] = {0, 0, 0, 0}, sd[4] = {0, 0, 0, 0}, spv[2] = {0, 0}, rd[2] = {0, 0}, rp = 0;
            double2 s1 = make_double2(1, 0), s2 = make_double2(1, 0), rcs = make_double2(1, 0);
            if (tid < h) {
                sm4[0] = E[sp_e];
                sm4[1] = E[NT + sp_e];
                sm4[2] = E[2 * NT + sp_e];
                sm4[3] = E[3 * NT + sp_e];
                s1 = CS[sp_ka];
                s2 = CS[sp_kb];
                sd[0] = Dg[sp_ka];
                sd[1] = Dg[h + sp_ka];
                sd[2] = Dg[sp_kb];
                sd[3] = Dg[h + sp_kb];
                spv[0] = E[4 * NT + sp_ka];
                spv[1] = E[4 * NT + sp_kb];
                rcs = CS[tid];
                rd[0] = Dg[tid];
                rd[1] = Dg[h + tid];
                rp = E[4 * NT + tid];
            }
            double2 vcs = make_double2(1, 0);
            double vu0 = 0, vu1 = 0;
            if (has_vec) {
                const double* uc = cur + L.off_u() + vec_v * m;
                vcs = CS[vec_k];
                vu0 = uc[vec_k];
                vu1 = uc[h + vec_k];
            }
            // ---- the special block, the next pivot of slot tid and its rotation -------------------------------
            if (tid < h) {
                const double c1 = s1.x, sn1 = s1.y, c2 = s2.x, sn2 = s2.y;
                const double t00 = c2 * sm4[0] - sn2 * sm4[1], t01 = sn2 * sm4[0] + c2 * sm4[1];
                const double t10 = c2 * sm4[2] - sn2 * sm4[3], t11 = sn2 * sm4[2] + c2 * sm4[3];
                double ev[4];
                ev[0] = c1 * t00 - sn1 * t10;
                ev[2] = sn1 * t00 + c1 * t10;
                ev[1] = c1 * t01 - sn1 * t11;
                ev[3] = sn1 * t01 + c1 * t11;
                const double apq = sp_q == 0 ? ev[0] : (sp_q == 1 ? ev[1] : (sp_q == 2 ? ev[2] : ev[3]));
                // new diagonal entries of the seats X (in slot ka or kb) and Y that arrive on slot tid
                const double da = rotated_diag(sp_x_in_ka ? sp_x_bottom : sp_y_bottom, sd[0], sd[1], spv[0], c1, sn1);
                const double db = rotated_diag(sp_x_in_ka ? sp_y_bottom : sp_x_bottom, sd[2], sd[3], spv[1], c2, sn2);
                const double app = sp_x_in_ka ? da : db, aqq = sp_x_in_ka ? db : da;
                double c = 1.0, s = 0.0;
                if (apq != 0.0) jacobi_angle(app, aqq, apq, c, s);
#pragma unroll
                for (int q = 0; q < 4; ++q) En[sp_dst[q]] = ev[q];
                nxt[L.off_diag() + tid] = app;
                nxt[L.off_diag() + h + tid] = aqq;
                reinterpret_cast<double2*>(nxt + L.off_cs())[tid] = make_double2(c, s);
                // residual of the current pivot of slot tid (the angle is inexact: the element is kept)
                En[res_dst] = fma(rcs.y * rcs.x, rd[0] - rd[1], (rcs.x * rcs.x - rcs.y * rcs.y) * rp);
            }
            // ---- regular blocks  M <- J_ka^T M J_kb ----------------------------------------------------------------
            auto rotate_store = [&](const double (&mm)[4], double2 r1, double2 r2, const int (&dst)[4]) {
                const double c1 = r1.x, sn1 = r1.y, c2 = r2.x, sn2 = r2.y;
                const double t00 = c2 * mm[0] - sn2 * mm[1], t01 = sn2 * mm[0] + c2 * mm[1];
                const double t10 = c2 * mm[2] - sn2 * mm[3], t11 = sn2 * mm[2] + c2 * mm[3];
                En[dst[0]] = c1 * t00 - sn1 * t10;
                En[dst[2]] = sn1 * t00 + c1 * t10;
                En[dst[1]] = c1 * t01 - sn1 * t11;
                En[dst[3]] = sn1 * t01 + c1 * t11;
            };
#pragma unroll
            for (int o = 0; o < 2; ++o)
                if (own_e[o] >= 0) rotate_store(bm[o], b1[o], b2[o], own_dst[o]);
            if (OWN > 2) {  // n > 92: two more blocks per thread, loaded late to stay inside the register budget
#pragma unroll
                for (int o = 2; o < OWN; ++o) {
                    if (own_e[o] >= 0) {
                        const int e = own_e[o];
                        const double mm[4] = {E[e], E[NT + e], E[2 * NT + e], E[3 * NT + e]};
                        rotate_store(mm, CS[own_k[o] & 0xff], CS[own_k[o] >> 8], own_dst[o]);
                    }
                }
            }
            if (has_vec) {
                double* un = nxt + L.off_u() + vec_v * m;
                un[vec_dt] = vcs.x * vu0 - vcs.y * vu1;
                un[vec_db] = vcs.y * vu0 + vcs.x * vu1;
            }
            __syncthreads();
            double* t = cur;
            cur = nxt;
            nxt = t;
        }
    }

    // ---- spectrum -> clipped spectrum -> pinvh cut-off -> whitened projections and log-determinant ------------
    // (whole sweeps only: every index is back on its own seat, seat p = index p; seats >= n are padding)
    const double* D = cur + L.off_diag();
    const double* U = cur + L.off_u();
    double lmax = -INFINITY;
    for (int i = tid; i < n; i += kJacobiThreads) lmax = fmax(lmax, D[i]);
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
        const double raw = D[r];
        const double lam = fmax(raw, floor_l);  // diffusion.py:887
        evals[r] = raw;
        ld_sum += log(lam);
        const bool keep = fabs(lam) > cutoff;
        const double s = keep ? 1.0 / sqrt(fabs(lam)) : 0.0;
        proj[r] = s * U[r];
        proj[n + r] = s * U[m + r];
        proj[2 * n + r] = s * U[2 * m + r];
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
    const int h = (n + 1) / 2 < kJacobiMinH ? kJacobiMinH : (n + 1) / 2;
    const JacobiLayout L{h * (h - 1) / 2, h, 2 * h};
    const size_t smem = sizeof(double) * 2 * (size_t)L.size();
    const int n_regular = L.NT - h;
    if (n_regular <= 2 * kJacobiThreads) {
        KB2_CUDA(cudaFuncSetAttribute(jacobi_project_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jacobi_project_kernel<2><<<1, kJacobiThreads, smem, st>>>(cov, n, h, x, y, cond_max, evals, proj, logdet_term, info);
    } else {
        KB2_CUDA(cudaFuncSetAttribute(jacobi_project_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jacobi_project_kernel<4><<<1, kJacobiThreads, smem, st>>>(cov, n, h, x, y, cond_max, evals, proj, logdet_term, info);
    }
    KB2_LAUNCH_CHECK();
    return launch_gls_solve(proj, n, fit_intercept, gls, st);
}
