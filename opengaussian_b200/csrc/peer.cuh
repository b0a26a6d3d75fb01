// peer.cuh -- device side of the small all-reduce over NVLink peer memory (see peer.cu for the protocol).
// Shared by the stand-alone kernel of ogs_peer_allreduce and by the k-means kernels, whose LAST CTA pushes the
// reduced centroid partials to the peers, waits for theirs and finishes the Lloyd update in the same launch.
#pragma once
#include "common.cuh"

namespace ogs {

#define PEER_MAX_RANKS 16

struct PeerDev {
    char* base[PEER_MAX_RANKS];     // every rank's inbox, mapped into this process (base[rank] = own)
    int rank, world;
    size_t slot_bytes, flag_off;
    unsigned long long* seq;        // call counter of this rank (device memory, owned by the comm)
    int* err;                       // raised when a wait times out
};

#ifdef __CUDACC__
__device__ __forceinline__ void peer_st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of ONE CTA (blockDim.x >= world).  In-place sum of buf[0..n) over the ranks, contributions
// added in rank order.  buf may live in shared or global memory.  Returns false if a peer did not arrive in time.
template <typename T>
__device__ __forceinline__ bool peer_allreduce_cta(const PeerDev& pd, T* buf, int n) {
    __shared__ unsigned long long s_seq;
    __shared__ int s_fail;
    if (threadIdx.x == 0) {
        s_seq = *pd.seq + 1ull;
        *pd.seq = s_seq;
        s_fail = 0;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = (int)(seq & 1ull);
    const int world = pd.world, rank = pd.rank;
    for (int p = 0; p < world; p++) {                                     // 1. push
        T* dst = reinterpret_cast<T*>(pd.base[p] + (size_t)(par * world + rank) * pd.slot_bytes);
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {                                       // 2. signal, 3. wait
        unsigned long long* remote = reinterpret_cast<unsigned long long*>(pd.base[threadIdx.x] + pd.flag_off) + (par * world + rank);
        peer_st_release_sys(remote, seq);
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pd.base[rank] + pd.flag_off) + (par * world + threadIdx.x);
        const long long t0 = clock64();
        while (peer_ld_acquire_sys(mine) < seq) {
            if (clock64() - t0 > 40000000000ll) { s_fail = 1; break; }    // ~20 s at 2 GHz
        }
    }
    __syncthreads();
    if (s_fail) {
        if (threadIdx.x == 0) atomicExch(pd.err, 1);
        return false;
    }
    const char* inbox = pd.base[rank] + (size_t)(par * world) * pd.slot_bytes;   // 4. reduce in rank order
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T s = 0;
        for (int p = 0; p < world; p++) s += __ldcg(reinterpret_cast<const T*>(inbox + (size_t)p * pd.slot_bytes) + i);
        buf[i] = s;
    }
    __syncthreads();
    return true;
}
#endif

}  // namespace ogs

// host-side view of a communicator (defined in peer.cu); the k-means launchers read its device descriptor
struct ogs_peer_comm;
namespace ogs {
const PeerDev* peer_comm_dev(const ogs_peer_comm* c);
size_t peer_comm_slot_bytes(const ogs_peer_comm* c);
}
