// The slice-level join over peer memory, shared by the stand-alone all-reduce kernel (k7_peer.cu) and by kernels that end
// with a sum over ranks (the fused frame sums of k1_unwrap.cu).  Protocol and window layout: see k7_peer.cu.
#pragma once
#include "common.cuh"

namespace kb2 {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerMaxSlices = 128;
constexpr int kPeerSlots = 8;
constexpr int kPeerFrameTile = 1024;  // columns per slice of a [rows][row_len] array joined by rows (the frame sums)

struct PeerWindows {
    unsigned char* base[kPeerMaxWorld];
};

// everything a CTA needs to join its slice of one call
struct PeerJoin {
    PeerWindows win;
    int world, rank;
    int n_slices;  // slices (CTAs) of this call on every rank: the one that finishes last acknowledges the slot
    int pad;
    long long capacity;       // doubles per (slot, rank) of the window
    unsigned long long seq;   // number of earlier calls on these windows (the same on every rank)
    long long limit;          // clock cycles a wait may take
    int* status;              // mapped host word: raised when a wait ran out
};

// window layout: flags [slots][W][kPeerMaxSlices] u64 | acks [slots][W] u64, done [slots] u64 (in a region of the size of
// the flags) | data [slots][W][capacity] f64
__host__ __device__ inline size_t peer_flag_words(int world) { return (size_t)kPeerSlots * world * kPeerMaxSlices; }
__host__ __device__ inline size_t peer_data_offset(int world) { return align256(2 * peer_flag_words(world) * 8); }
__host__ __device__ inline size_t peer_window_bytes(int world, int64_t capacity) {
    return peer_data_offset(world) + (size_t)kPeerSlots * world * (size_t)capacity * 8;
}

__device__ __forceinline__ unsigned long long peer_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// true when *p reached `token` within `limit` clock cycles
__device__ __forceinline__ bool peer_wait_token(const unsigned long long* p, unsigned long long token, long long limit) {
    const long long t0 = clock64();
    unsigned int ns = 20;
    while (peer_ld_acquire_sys(p) < token) {
        __nanosleep(ns);
        if (ns < 1000) ns += ns;
        if (clock64() - t0 > limit) return false;
    }
    return true;
}

// Joins slice c of a [rows][row_len] float64 array (row stride row_len; columns [c0, c1) of every row; c0 and row_len even)
// over the ranks, in place: called by ALL threads of a CTA (blockDim.x >= world, a multiple of 32), which must not be
// divergent.  `buf` is read through L2 (its producers may be other CTAs of the calling kernel, which have fenced).
__device__ __forceinline__ void peer_join_slice(const PeerJoin& J, double* __restrict__ buf, int rows, long long row_len, int c,
                                                long long c0, long long c1) {
    const int tid = threadIdx.x, nthr = blockDim.x, world = J.world, rank = J.rank;
    const int slot = (int)(J.seq % (unsigned long long)kPeerSlots);
    const unsigned long long token = J.seq + 1;
    const size_t words = peer_flag_words(world);
    const size_t data_off = peer_data_offset(world);
    unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(J.win.base[rank]);
    unsigned long long* my_acks = my_flags + words;
    const size_t cell = ((size_t)slot * world) * kPeerMaxSlices;  // + q * kPeerMaxSlices + c
    __shared__ int s_bad, s_last;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // (1) credit: the previous user of this slot (sequence seq - slots, however it was sliced) has been consumed by every rank
    if (J.seq >= (unsigned long long)kPeerSlots && tid < world) {
        if (!peer_wait_token(my_acks + (size_t)slot * world + tid, token - kPeerSlots, J.limit)) s_bad = 1;
    }
    __syncthreads();
    // (2) the slice goes to every window, the nearest "next" rank first so that the W ranks do not all write to rank 0 at once
    const long long half = (c1 - c0 + 1) / 2;  // pairs of columns per row
    for (long long p = tid; p < half * rows; p += nthr) {
        const int r = (int)(p / half);
        const long long i = (long long)r * row_len + c0 + 2 * (p - (long long)r * half);
        const bool pair = (c0 + 2 * (p - (long long)r * half) + 1 < c1);
        double2 v;
        if (pair) {
            v = __ldcg(reinterpret_cast<const double2*>(buf + i));
        } else {
            v.x = __ldcg(buf + i);
            v.y = 0.0;
        }
#pragma unroll 1
        for (int h = 0; h < world; ++h) {
            int q = rank + 1 + h;
            if (q >= world) q -= world;
            double* dst = reinterpret_cast<double*>(J.win.base[q] + data_off) + ((size_t)slot * world + rank) * (size_t)J.capacity + i;
            if (pair)
                *reinterpret_cast<double2*>(dst) = v;
            else
                dst[0] = v.x;
        }
    }
    // (3) every thread's stores are fenced at system scope, then one thread per destination raises the flag
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        unsigned long long* flags_q = reinterpret_cast<unsigned long long*>(J.win.base[tid]);
        peer_st_release_sys(flags_q + cell + (size_t)rank * kPeerMaxSlices + c, token);
        // (4) ... and waits for the copy of rank `tid` in this rank's window
        if (!peer_wait_token(my_flags + cell + (size_t)tid * kPeerMaxSlices + c, token, J.limit)) s_bad = 2;
    }
    __syncthreads();
    // (5) the W copies added in rank order (L2 loads: the copies were written by other devices)
    const double* mine = reinterpret_cast<const double*>(J.win.base[rank] + data_off) + (size_t)slot * world * (size_t)J.capacity;
    for (long long p = tid; p < half * rows; p += nthr) {
        const int r = (int)(p / half);
        const long long i = (long long)r * row_len + c0 + 2 * (p - (long long)r * half);
        if (c0 + 2 * (p - (long long)r * half) + 1 < c1) {
            double2 acc = __ldcg(reinterpret_cast<const double2*>(mine + i));
            for (int q = 1; q < world; ++q) {
                const double2 x = __ldcg(reinterpret_cast<const double2*>(mine + (size_t)q * J.capacity + i));
                acc.x += x.x;
                acc.y += x.y;
            }
            *reinterpret_cast<double2*>(buf + i) = acc;
        } else {
            double acc = __ldcg(mine + i);
            for (int q = 1; q < world; ++q) acc += __ldcg(mine + (size_t)q * J.capacity + i);
            buf[i] = acc;
        }
    }
    // (6) the CTA that finishes last on this rank releases the slot to its writers
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned long long* done = my_acks + (size_t)kPeerSlots * world + slot;
        const unsigned long long before = atomicAdd(done, 1ull);
        s_last = (before + 1 == (unsigned long long)J.n_slices);
        if (s_last) {
            atomicExch(done, 0ull);  // the next user of the slot cannot get here before the acknowledgements below
            __threadfence();
        }
    }
    __syncthreads();
    if (s_last && tid < world) {
        unsigned long long* acks_q = reinterpret_cast<unsigned long long*>(J.win.base[tid]) + words;
        peer_st_release_sys(acks_q + (size_t)slot * world + rank, token);
    }
    if (tid == 0 && s_bad) atomicExch(J.status, s_bad);  // mapped host memory: visible without a synchronisation
}

// host side (k7_peer.cu): fills J from the caller's arguments; non-zero on error
int peer_join_setup(PeerJoin* J, void* const* windows_host, int world, int rank, int64_t capacity, uint64_t seq, int n_slices,
                    const char* fn);

}  // namespace kb2
