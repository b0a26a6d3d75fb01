// nvls.cu -- all-reduce of the step's gradient buffer through the NVSwitch multicast mapping, in one kernel.
//
// View-parallel training sums the Gaussian-parameter gradients of all ranks once per step:
// 236 MB at 1 M Gaussians, 708 MB at 3 M.  The buffer lives in symmetric memory that every rank maps at the same
// offset of one MULTICAST address range (torch.distributed._symmetric_memory owns the allocation and the rendezvous;
// the kernel is ours).  Rank r owns slice r of the buffer:
//     multimem.ld_reduce.add.v4.f32  reads the 16 bytes from ALL ranks' copies and returns their sum -- the NVSwitch
//                                    adds in flight, each GPU link carries the slice once --
//     multimem.st.v4.f32             writes the sum back to ALL copies.
// Per GPU the links move the buffer once in and once out (the two-shot schedule of a library all-reduce, without its
// protocol, staging copies and channel bookkeeping).  The caller brackets the launch with a cross-rank barrier on
// the stream (every rank's gradients written before / every slice stored after).
#include <stdlib.h>

#include "common.cuh"

namespace ogs {

#define NVLS_THREADS 512

__device__ __forceinline__ void mm_ld_reduce(const float* mc, uint32_t (&v)[4]) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "l"(mc)
                 : "memory");
}
__device__ __forceinline__ void mm_st(float* mc, const uint32_t (&v)[4]) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
                 : "memory");
}

template <int NVLS_UNROLL>
__global__ void __launch_bounds__(NVLS_THREADS) multimem_allreduce_f32_kernel(float* __restrict__ mc, size_t n4, int rank, int world) {
    const size_t per = (n4 + (size_t)world - 1) / (size_t)world;
    const size_t lo = (size_t)rank * per;
    const size_t hi = lo + per < n4 ? lo + per : n4;
    const size_t stride = (size_t)gridDim.x * NVLS_THREADS;
    for (size_t i = lo + (size_t)blockIdx.x * NVLS_THREADS + threadIdx.x; i < hi; i += stride * NVLS_UNROLL) {
        uint32_t v[NVLS_UNROLL][4];
#pragma unroll
        for (int u = 0; u < NVLS_UNROLL; u++)
            if (i + u * stride < hi) mm_ld_reduce(mc + 4 * (i + u * stride), v[u]);
#pragma unroll
        for (int u = 0; u < NVLS_UNROLL; u++)
            if (i + u * stride < hi) mm_st(mc + 4 * (i + u * stride), v[u]);
    }
}

}  // namespace ogs

using namespace ogs;

extern "C" int ogs_multimem_allreduce_f32(void* multicast_ptr, int64_t n_floats, int32_t rank, int32_t world, void* stream_) {
    if (!multicast_ptr || n_floats < 0 || world < 1 || rank < 0 || rank >= world || (n_floats & 3) || ((uintptr_t)multicast_ptr & 15)) {
        set_error("multimem_allreduce_f32: needs a 16-byte aligned multicast pointer and a multiple of 4 floats");
        return -1;
    }
    if (n_floats == 0) return 0;
    static const int env_grid = getenv("OGS_NVLS_GRID") ? atoi(getenv("OGS_NVLS_GRID")) : 0;        // tuning knobs
    static const int env_unroll = getenv("OGS_NVLS_UNROLL") ? atoi(getenv("OGS_NVLS_UNROLL")) : 4;
    const size_t n4 = (size_t)n_floats / 4;
    const size_t per = (n4 + world - 1) / world;
    const size_t want = (per + (size_t)NVLS_THREADS * env_unroll - 1) / ((size_t)NVLS_THREADS * env_unroll);
    const size_t cap = env_grid > 0 ? (size_t)env_grid : (size_t)OGS_NUM_SMS * 2;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    cudaStream_t s = (cudaStream_t)stream_;
    if (env_unroll >= 8) multimem_allreduce_f32_kernel<8><<<grid, NVLS_THREADS, 0, s>>>((float*)multicast_ptr, n4, rank, world);
    else if (env_unroll <= 2) multimem_allreduce_f32_kernel<2><<<grid, NVLS_THREADS, 0, s>>>((float*)multicast_ptr, n4, rank, world);
    else multimem_allreduce_f32_kernel<4><<<grid, NVLS_THREADS, 0, s>>>((float*)multicast_ptr, n4, rank, world);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "multimem_allreduce_f32");
    return 0;
}
