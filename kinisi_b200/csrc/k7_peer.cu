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

#include "peer.cuh"

namespace kb2 {

constexpr int kPeerThreads = 256;
constexpr int kPeerSliceMin = 1024;  // doubles per slice before a second CTA is worth its launch

// slice c = columns [c * per, (c + 1) * per) of every row
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(double* __restrict__ buf, int rows, long long row_len,
                                                                       long long per, PeerJoin J) {
    const int c = blockIdx.x;
    const long long c0 = (long long)c * per, c1 = min(row_len, c0 + per);
    peer_join_slice(J, buf, rows, row_len, c, c0, c1);
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

int peer_join_setup(PeerJoin* J, void* const* windows_host, int world, int rank, int64_t capacity, uint64_t seq, int n_slices,
                    const char* fn) {
    KB2_CHECK(windows_host, "%s: null window list", fn);
    KB2_CHECK(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "%s: bad world / rank", fn);
    KB2_CHECK(n_slices >= 1 && n_slices <= kPeerMaxSlices, "%s: bad slice count", fn);
    if (int rc = peer_status_word()) return rc;
    for (int q = 0; q < kPeerMaxWorld; ++q) J->win.base[q] = q < world ? (unsigned char*)windows_host[q] : nullptr;
    for (int q = 0; q < world; ++q) KB2_CHECK(J->win.base[q], "%s: window %d is null", fn, q);
    J->world = world, J->rank = rank, J->n_slices = n_slices, J->pad = 0;
    J->capacity = capacity, J->seq = seq, J->limit = peer_wait_limit(), J->status = g_status_dev;
    return 0;
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_peer_slots(void) { return kPeerSlots; }

extern "C" size_t kb2_peer_window_bytes(int world, int64_t capacity) {
    if (world < 1 || world > kPeerMaxWorld || capacity < 2) return 0;
    return peer_window_bytes(world, capacity);
}

extern "C" int kb2_peer_window_create(int world, int64_t capacity, void** window, unsigned char* handle64) {
    KB2_CHECK(window && handle64, "kb2_peer_window_create: null pointer");
    KB2_CHECK(world >= 1 && world <= kPeerMaxWorld, "kb2_peer_window_create: 1 <= world <= %d", kPeerMaxWorld);
    KB2_CHECK(capacity >= 2 && capacity % 2 == 0, "kb2_peer_window_create: capacity must be even and >= 2");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const size_t bytes = peer_window_bytes(world, capacity);
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
    KB2_CHECK(buf, "kb2_peer_allreduce_f64: null pointer");
    KB2_CHECK(n >= 1 && n <= capacity, "kb2_peer_allreduce_f64: n must be in [1, capacity]");
    KB2_CHECK((reinterpret_cast<uintptr_t>(buf) & 15) == 0, "kb2_peer_allreduce_f64: buf must be 16-byte aligned");
    // the slicing depends on n alone: every rank cuts the array the same way
    int slices = (int)std::min<int64_t>(kPeerMaxSlices, (n + kPeerSliceMin - 1) / kPeerSliceMin);
    int64_t per = (n + slices - 1) / slices;
    per += per & 1;
    slices = (int)((n + per - 1) / per);
    PeerJoin J;
    if (int rc = peer_join_setup(&J, windows_host, world, rank, capacity, seq, slices, "kb2_peer_allreduce_f64")) return rc;
    peer_allreduce_kernel<<<slices, kPeerThreads, 0, (cudaStream_t)stream>>>(buf, 1, (long long)n, (long long)per, J);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_peer_rows_slices(int64_t row_len) {
    const int64_t s = (row_len + kPeerFrameTile - 1) / kPeerFrameTile;
    return (row_len < 2 || (row_len & 1) || s > kPeerMaxSlices) ? 0 : (int)s;
}

extern "C" int kb2_peer_allreduce_rows_f64(double* buf, int rows, int64_t row_len, void* const* windows_host, int world, int rank,
                                           int64_t capacity, uint64_t seq, void* stream) {
    KB2_CHECK(buf, "kb2_peer_allreduce_rows_f64: null pointer");
    KB2_CHECK(rows >= 1 && row_len >= 2 && (int64_t)rows * row_len <= capacity, "kb2_peer_allreduce_rows_f64: rows * row_len must be in [2, capacity]");
    KB2_CHECK((reinterpret_cast<uintptr_t>(buf) & 15) == 0, "kb2_peer_allreduce_rows_f64: buf must be 16-byte aligned");
    const int slices = kb2_peer_rows_slices(row_len);
    KB2_CHECK(slices > 0, "kb2_peer_allreduce_rows_f64: row_len must be even and at most %d columns", kPeerMaxSlices * kPeerFrameTile);
    PeerJoin J;
    if (int rc = peer_join_setup(&J, windows_host, world, rank, capacity, seq, slices, "kb2_peer_allreduce_rows_f64")) return rc;
    peer_allreduce_kernel<<<slices, kPeerThreads, 0, (cudaStream_t)stream>>>(buf, rows, (long long)row_len, (long long)kPeerFrameTile, J);
    KB2_LAUNCH_CHECK();
    return 0;
}
