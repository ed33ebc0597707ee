// Multi-GPU joins of the path through peer memory (NVLink / NVSwitch), one kernel per join.
//
// When the atoms of a trajectory are sharded over ranks, two small float64 arrays are summed over ranks per
// analysis: the per-frame sums of the framework atoms, [3][pitch] (the drift of kinisi/parser.py:131-148; the collective
// displacement of diffusion.py:659, 751 has the same shape), and the per-lag sums of the moment kernels, [K][2]
// (diffusion.py:571-602).  Both are latency-sized (1.6 KB and 0.5 MB): a ring all-reduce kernel plus its host-side call
// cost more than the data movement.  Here every rank owns a window of device memory that its peers map (CUDA IPC), and
// ONE launch per rank does the whole join:
//
//   CTA c of rank r owns slice c of the array
//     (1) waits until every peer has released the slot (credit from the previous use, normally long there),
//     (2) reads its slice once and stores it into window[q].data[slot][r][slice] of EVERY rank q (its own included),
//     (3) fences at system scope and raises window[q].flag[slot][r][c] to the call's token,
//     (4) waits until the W flags of its slice in its OWN window carry the token,
//     (5) adds the W copies in rank order -- the same order on every rank, so all ranks hold bit-identical sums,
//         which NCCL's ring does not promise -- and writes the result over the input slice,
//     (6) counts itself done; the last CTA of the launch tells every peer that this rank has consumed the slot
//         (window[q].ack[slot][r] = token).
//
// Slices are independent: there is no barrier across the CTAs of a launch, and the exchange of one slice overlaps the
// copies of the others.  Tokens are the call's sequence number + 1, the same on every rank because every rank issues
// the joins of a channel in the same program order; calls `slots` apart share a slot, and step (1) makes that safe
// whatever the streams they were issued on.  Every wait is bounded (KB2_PEER_TIMEOUT_S, 120 s unless set): a peer that
// never arrives raises a status word in mapped host memory instead of hanging the device.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace kb2 {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerMaxSlices = 64;
constexpr int kPeerThreads = 256;
constexpr int kPeerSliceMin = 1024;  // doubles per slice before a second CTA is worth its launch

struct PeerWindows {
    unsigned char* base[kPeerMaxWorld];
};

// window layout: flags [slots][W][kPeerMaxSlices] u64 | acks [slots][W] u64, done [slots] u64 (in a region of the size of
// the flags) | data [slots][W][capacity] f64
__host__ __device__ inline size_t peer_flag_words(int world, int slots) { return (size_t)slots * world * kPeerMaxSlices; }
__host__ __device__ inline size_t peer_data_offset(int world, int slots) { return align256(2 * peer_flag_words(world, slots) * 8); }
__host__ __device__ inline size_t peer_window_bytes(int world, int slots, int64_t capacity) {
    return peer_data_offset(world, slots) + (size_t)slots * world * (size_t)capacity * 8;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// true when *p reached `token` within `limit` clock cycles
__device__ __forceinline__ bool wait_token(const unsigned long long* p, unsigned long long token, long long limit) {
    const long long t0 = clock64();
    unsigned int ns = 20;
    while (ld_acquire_sys(p) < token) {
        __nanosleep(ns);
        if (ns < 1000) ns += ns;
        if (clock64() - t0 > limit) return false;
    }
    return true;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(double* __restrict__ buf, int64_t n, int64_t per_slice,
                                                                       PeerWindows win, int world, int rank, int slots,
                                                                       int64_t capacity, unsigned long long seq,
                                                                       long long limit, int* __restrict__ status) {
    const int c = blockIdx.x, tid = threadIdx.x;
    const int slot = (int)(seq % (unsigned long long)slots);
    const unsigned long long token = seq + 1;
    const size_t words = peer_flag_words(world, slots);
    const size_t data_off = peer_data_offset(world, slots);
    const int64_t i0 = (int64_t)c * per_slice, i1 = min(n, i0 + per_slice);
    unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(win.base[rank]);
    unsigned long long* my_acks = my_flags + words;
    const size_t cell = ((size_t)slot * world) * kPeerMaxSlices;  // + q * kPeerMaxSlices + c
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // (1) credit: the previous user of this slot (sequence seq - slots, however it was sliced) has been consumed by every rank
    if (seq >= (unsigned long long)slots && tid < world) {
        if (!wait_token(my_acks + (size_t)slot * world + tid, token - slots, limit)) s_bad = 1;
    }
    __syncthreads();
    // (2) the slice goes to every window, the nearest "next" rank first so that the W ranks do not all write to rank 0 at once
    for (int64_t i = i0 + 2 * tid; i < i1; i += 2 * kPeerThreads) {
        const bool pair = (i + 1 < i1);
        double2 v;
        if (pair) {
            v = *reinterpret_cast<const double2*>(buf + i);
        } else {
            v.x = buf[i];
            v.y = 0.0;
        }
#pragma unroll 1
        for (int h = 0; h < world; ++h) {
            int q = rank + 1 + h;
            if (q >= world) q -= world;
            double* dst = reinterpret_cast<double*>(win.base[q] + data_off) + ((size_t)slot * world + rank) * (size_t)capacity + i;
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
        unsigned long long* flags_q = reinterpret_cast<unsigned long long*>(win.base[tid]);
        st_release_sys(flags_q + cell + (size_t)rank * kPeerMaxSlices + c, token);
        // (4) ... and waits for the copy of rank `tid` in this rank's window
        if (!wait_token(my_flags + cell + (size_t)tid * kPeerMaxSlices + c, token, limit)) s_bad = 2;
    }
    __syncthreads();
    // (5) the W copies added in rank order (L2 loads: the copies were written by other devices)
    const double* mine = reinterpret_cast<const double*>(win.base[rank] + data_off) + (size_t)slot * world * (size_t)capacity;
    for (int64_t i = i0 + 2 * tid; i < i1; i += 2 * kPeerThreads) {
        if (i + 1 < i1) {
            double2 acc = __ldcg(reinterpret_cast<const double2*>(mine + i));
            for (int q = 1; q < world; ++q) {
                const double2 x = __ldcg(reinterpret_cast<const double2*>(mine + (size_t)q * capacity + i));
                acc.x += x.x;
                acc.y += x.y;
            }
            *reinterpret_cast<double2*>(buf + i) = acc;
        } else {
            double acc = __ldcg(mine + i);
            for (int q = 1; q < world; ++q) acc += __ldcg(mine + (size_t)q * capacity + i);
            buf[i] = acc;
        }
    }
    // (6) the CTA that finishes last on this rank releases the slot to its writers
    __shared__ int s_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned long long* done = my_acks + (size_t)slots * world + slot;
        const unsigned long long before = atomicAdd(done, 1ull);
        s_last = (before + 1 == (unsigned long long)gridDim.x);
        if (s_last) {
            atomicExch(done, 0ull);  // the next user of the slot cannot get here before the acknowledgements below
            __threadfence();
        }
    }
    __syncthreads();
    if (s_last && tid < world) {
        unsigned long long* acks_q = reinterpret_cast<unsigned long long*>(win.base[tid]) + words;
        st_release_sys(acks_q + (size_t)slot * world + rank, token);
    }
    if (tid == 0 && s_bad) atomicExch(status, s_bad);  // mapped host memory: visible without a synchronisation
}

static int* g_status_host = nullptr;  // one status word per process, mapped into the device's address space
static int* g_status_dev = nullptr;

static long long peer_wait_limit() {  // clock cycles
    static long long limit = [] {
        const char* e = getenv("KB2_PEER_TIMEOUT_S");
        double s = e ? atof(e) : 120.0;
        if (!(s > 0.0)) s = 120.0;
        return (long long)(s * 2.0e9);
    }();
    return limit;
}

static int peer_status_word() {
    if (g_status_host) return 0;
    KB2_CUDA(cudaHostAlloc((void**)&g_status_host, 256, cudaHostAllocMapped));
    *g_status_host = 0;
    KB2_CUDA(cudaHostGetDevicePointer((void**)&g_status_dev, g_status_host, 0));
    return 0;
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_peer_slots(void) { return 8; }

extern "C" size_t kb2_peer_window_bytes(int world, int64_t capacity) {
    if (world < 1 || world > kPeerMaxWorld || capacity < 2) return 0;
    return peer_window_bytes(world, kb2_peer_slots(), capacity);
}

extern "C" int kb2_peer_window_create(int world, int64_t capacity, void** window, unsigned char* handle64) {
    KB2_CHECK(window && handle64, "kb2_peer_window_create: null pointer");
    KB2_CHECK(world >= 1 && world <= kPeerMaxWorld, "kb2_peer_window_create: 1 <= world <= %d", kPeerMaxWorld);
    KB2_CHECK(capacity >= 2 && capacity % 2 == 0, "kb2_peer_window_create: capacity must be even and >= 2");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const size_t bytes = peer_window_bytes(world, kb2_peer_slots(), capacity);
    void* p = nullptr;
    KB2_CUDA(cudaMalloc(&p, bytes));
    KB2_CUDA(cudaMemset(p, 0, bytes));  // flags and acks start below every token
    KB2_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("kb2_peer_window_create: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return 2;
    }
    memcpy(handle64, &h, 64);
    *window = p;
    return 0;
}

extern "C" int kb2_peer_window_open(const unsigned char* handle64, void** window) {
    KB2_CHECK(window && handle64, "kb2_peer_window_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    KB2_CUDA(cudaIpcOpenMemHandle(window, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int kb2_peer_window_close(void* window) {
    KB2_CHECK(window, "kb2_peer_window_close: null pointer");
    KB2_CUDA(cudaIpcCloseMemHandle(window));
    return 0;
}

extern "C" int kb2_peer_window_destroy(void* window) {
    KB2_CHECK(window, "kb2_peer_window_destroy: null pointer");
    KB2_CUDA(cudaFree(window));
    return 0;
}

extern "C" int kb2_peer_status(void) { return g_status_host ? *(volatile int*)g_status_host : 0; }

extern "C" int kb2_peer_allreduce_f64(double* buf, int64_t n, void* const* windows_host, int world, int rank, int64_t capacity,
                                      uint64_t seq, void* stream) {
    KB2_CHECK(buf && windows_host, "kb2_peer_allreduce_f64: null pointer");
    KB2_CHECK(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "kb2_peer_allreduce_f64: bad world / rank");
    KB2_CHECK(n >= 1 && n <= capacity, "kb2_peer_allreduce_f64: n must be in [1, capacity]");
    KB2_CHECK((reinterpret_cast<uintptr_t>(buf) & 15) == 0, "kb2_peer_allreduce_f64: buf must be 16-byte aligned");
    if (int rc = peer_status_word()) return rc;
    PeerWindows win;
    for (int q = 0; q < kPeerMaxWorld; ++q) win.base[q] = q < world ? (unsigned char*)windows_host[q] : nullptr;
    for (int q = 0; q < world; ++q) KB2_CHECK(win.base[q], "kb2_peer_allreduce_f64: window %d is null", q);
    // the slicing depends on n alone: every rank cuts the array the same way
    int slices = (int)std::min<int64_t>(kPeerMaxSlices, (n + kPeerSliceMin - 1) / kPeerSliceMin);
    int64_t per = (n + slices - 1) / slices;
    per += per & 1;
    slices = (int)((n + per - 1) / per);
    peer_allreduce_kernel<<<slices, kPeerThreads, 0, (cudaStream_t)stream>>>(buf, n, per, win, world, rank, kb2_peer_slots(),
                                                                           capacity, (unsigned long long)seq, peer_wait_limit(), g_status_dev);
    KB2_LAUNCH_CHECK();
    return 0;
}
