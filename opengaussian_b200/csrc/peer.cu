// peer.cu -- small-message all-reduce over NVLink peer memory, inside one kernel.
//
// The k-means Lloyd iteration on R GPUs needs the sum of a few KB per pass (coarse: [64, 10] floats = 2.6 KB;
// fine, all clusters: [640, 7] int64 = 35 KB) -- SURVEY.md section 8e.  A library collective costs ~45 us of
// launch + protocol per call against ~30 us of compute per pass at 625 k points per rank, so the collective is
// done by ONE tiny kernel over memory the ranks map from each other (cudaIpc handles over NVLink/NVSwitch):
//
//   every rank owns  inbox[2][R][slot]  and  flag[2][R]  (2 = parity of the call's sequence number)
//   1. push   : rank r stores its vector into inbox[par][r] of EVERY rank (remote stores, fire and forget)
//   2. signal : fence.sys, then flag[par][r] = seq on every rank (st.release.sys)
//   3. wait   : spin on the own flags until all R show seq (ld.acquire.sys)
//   4. reduce : sum the R slots of the own inbox in RANK ORDER -- every rank adds the same numbers in the same
//               order, so the replicated centres stay bit-identical across ranks.
//
// Two parities suffice: a rank can start call s+1 only after it finished s, i.e. after every peer pushed s, i.e.
// after every peer finished reading s-1 -- so a push into parity (s+1) % 2 never overwrites a slot still in use.
// A rank that waits longer than ~20 s gives up and raises the comm's error word (the host reports it) instead of
// hanging the GPU.  The sequence number lives in device memory and is advanced by the kernel itself, so a captured
// CUDA graph can replay the call.  The device side (peer.cuh) is also called by the LAST CTA of the k-means kernels:
// assign + centroid partials + all-reduce + centre update of a Lloyd pass are then ONE launch.
#include <string.h>

#include "peer.cuh"

namespace ogs {

#define PEER_THREADS 512

// one CTA.  T = float or long long.
template <typename T>
__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(PeerDev pd, T* __restrict__ buf, int n) {
    peer_allreduce_cta<T>(pd, buf, n);
}

}  // namespace ogs

using namespace ogs;

struct ogs_peer_comm {
    int rank, world, dev;
    size_t slot_bytes, flag_off, total;
    char* local;
    PeerDev dev_desc;     // passed by value to the kernels
    bool opened[PEER_MAX_RANKS];
    int* err_dev;
    int* err_host;        // pinned mirror read by ogs_peer_comm_error
    cudaIpcMemHandle_t handle;
};

namespace ogs {
const PeerDev* peer_comm_dev(const ogs_peer_comm* c) { return c ? &c->dev_desc : nullptr; }
size_t peer_comm_slot_bytes(const ogs_peer_comm* c) { return c ? c->slot_bytes : 0; }
}

extern "C" {

int ogs_peer_comm_create(int32_t rank, int32_t world, int64_t max_bytes, ogs_peer_comm** out, void* handle_out64) {
    if (!out || !handle_out64 || world < 1 || world > PEER_MAX_RANKS || rank < 0 || rank >= world || max_bytes < 8) {
        set_error("peer_comm_create: bad arguments (world <= %d)", PEER_MAX_RANKS);
        return -1;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    ogs_peer_comm* c = new ogs_peer_comm();
    memset(c, 0, sizeof *c);
    c->rank = rank; c->world = world;
    OGS_CUDA(cudaGetDevice(&c->dev));
    c->slot_bytes = align_up((size_t)max_bytes, 256);
    c->flag_off = 2 * (size_t)world * c->slot_bytes;
    c->total = c->flag_off + align_up(2 * (size_t)world * 8, 256) + 256;
    cudaError_t e = cudaMalloc((void**)&c->local, c->total);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->total);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&c->handle, c->local);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->err_host, 64);
    if (e != cudaSuccess) {
        if (c->local) cudaFree(c->local);
        delete c;
        return cuda_fail(e, "peer_comm_create");
    }
    *c->err_host = 0;
    c->err_dev = reinterpret_cast<int*>(c->local + c->total - 256);
    c->dev_desc.base[rank] = c->local;
    c->dev_desc.rank = rank; c->dev_desc.world = world;
    c->dev_desc.slot_bytes = c->slot_bytes; c->dev_desc.flag_off = c->flag_off;
    c->dev_desc.seq = reinterpret_cast<unsigned long long*>(c->local + c->total - 128);   // zeroed with the buffer
    c->dev_desc.err = c->err_dev;
    memcpy(handle_out64, &c->handle, 64);
    OGS_CUDA(cudaDeviceSynchronize());
    *out = c;
    return 0;
}

int ogs_peer_comm_connect(ogs_peer_comm* c, const void* all_handles) {
    if (!c || !all_handles) { set_error("peer_comm_connect: NULL"); return -1; }
    const char* h = (const char*)all_handles;
    for (int p = 0; p < c->world; p++) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, h + (size_t)p * 64, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle (peer memory over NVLink unavailable?)");
        c->dev_desc.base[p] = (char*)ptr;
        c->opened[p] = true;
    }
    return 0;
}

/* dtype: 0 = float32, 1 = int64.  In place on buf [n]; n * sizeof(T) <= max_bytes of create. */
int ogs_peer_allreduce(ogs_peer_comm* c, void* buf, int64_t n, int32_t dtype, void* stream_) {
    if (!c || !buf || n < 0 || (dtype != 0 && dtype != 1)) { set_error("peer_allreduce: bad arguments"); return -1; }
    const size_t bytes = (size_t)n * (dtype == 0 ? 4 : 8);
    if (bytes > c->slot_bytes) { set_error("peer_allreduce: %zu bytes exceed the slot size %zu", bytes, c->slot_bytes); return -1; }
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream_;
    if (dtype == 0) peer_allreduce_kernel<float><<<1, PEER_THREADS, 0, s>>>(c->dev_desc, (float*)buf, (int)n);
    else peer_allreduce_kernel<long long><<<1, PEER_THREADS, 0, s>>>(c->dev_desc, (long long*)buf, (int)n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "peer_allreduce");
    return 0;
}

/* 1 if a call timed out waiting for a peer since the last check (synchronises the stream). */
int ogs_peer_comm_error(ogs_peer_comm* c, void* stream_) {
    if (!c) return -1;
    cudaStream_t s = (cudaStream_t)stream_;
    OGS_CUDA(cudaMemcpyAsync(c->err_host, c->err_dev, 4, cudaMemcpyDeviceToHost, s));
    OGS_CUDA(cudaStreamSynchronize(s));
    return *c->err_host ? 1 : 0;
}

int ogs_peer_comm_destroy(ogs_peer_comm* c) {
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world; p++)
        if (c->opened[p]) cudaIpcCloseMemHandle(c->dev_desc.base[p]);
    if (c->local) cudaFree(c->local);
    if (c->err_host) cudaFreeHost(c->err_host);
    delete c;
    return 0;
}

}  // extern "C"
