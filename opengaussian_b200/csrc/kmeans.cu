// kmeans.cu -- codebook assign / centroid-sum / finalize / straight-through gather kernels.
//
// Replaces the PyTorch ops of scene/kmeans_quantize.py: torch.cdist + argmin (:38-55,:181-182,
// :200-201,:223-224,:237-238), one-hot @ feat centroid sums and counts (:82-87,:183-187,:202-205),
// centres = sums / counts (:208-214), gather + straight-through (:273-275).
//
// Arithmetic contract (shared with oracle/kmeans_oracle.c so that ids are bit-exact):
//   dist2(x, c) = fold_d fmaf(x_d - c_d, x_d - c_d, acc), acc0 = 0, d ascending;
//   id = lowest index of the minimum (strict '<' scan) -- torch.argmin's tie rule.
// The contraction is D <= 16 deep: FMA cores, not tensor cores (BASELINE.json north_star).
//
// Fusion: the centroid sums are accumulated in the SAME pass that assigns.  Each warp owns a
// private [k][D+1] accumulator in shared memory; lanes with equal ids are serialised by their rank
// inside the __match_any_sync peer group, so every update is a plain LDS/FADD/STS (no shared or
// global atomics) and the result is deterministic.  Block partials go to a [grid][k][D+1] scratch
// that a second tiny kernel reduces in fixed order.
// HBM-bound at the fine level (k <= 10), FP32-bound at the coarse level (k = 64, D = 9).
#include "common.cuh"

namespace ogs {

#define KM_THREADS 256
#define KM_MAX_D 16

template <int D>
__global__ void __launch_bounds__(KM_THREADS) kmeans_assign_kernel(
    int64_t N, const float* __restrict__ a, int Da, const float* __restrict__ b, int Db, float scale_b,
    const float* __restrict__ centers, int k, const int64_t* __restrict__ select_ids, int64_t selected,
    int64_t id_offset, int64_t* __restrict__ ids_out, float* __restrict__ partials /* [grid][k][D+1] or NULL */) {
    extern __shared__ float smem[];
    float* s_c = smem;                       // [k][D]
    float* s_acc = smem + (size_t)k * D;     // [8][k][D+1] (only when partials)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int ROW = D + 1;

    for (int e = threadIdx.x; e < k * D; e += KM_THREADS) s_c[e] = centers[e];
    if (partials)
        for (int e = threadIdx.x; e < 8 * k * ROW; e += KM_THREADS) s_acc[e] = 0.f;
    __syncthreads();
    float* my_acc = s_acc + (size_t)warp * k * ROW;

    const int64_t stride = (int64_t)gridDim.x * KM_THREADS;
    const int64_t n_round = (N + stride - 1) / stride;
    for (int64_t it = 0; it < n_round; it++) {
        const int64_t i = it * stride + (int64_t)blockIdx.x * KM_THREADS + threadIdx.x;
        bool active = i < N;
        if (active && select_ids) active = (select_ids[i] == selected);
        float x[D];
        int best_j = -1;
        if (active) {
#pragma unroll
            for (int d = 0; d < D; d++) {
                x[d] = (d < Da) ? __ldg(a + i * Da + d) : __fmul_rn(__ldg(b + i * Db + (d - Da)), scale_b);
            }
            float best = INFINITY;
            best_j = 0;
            for (int j = 0; j < k; j++) {
                const float* c = s_c + j * D;
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < D; d++) {
                    const float df = __fsub_rn(x[d], c[d]);
                    acc = __fmaf_rn(df, df, acc);
                }
                if (acc < best) { best = acc; best_j = j; }
            }
            ids_out[i] = id_offset + best_j;
        }
        if (partials) {
            // conflict-free, atomic-free accumulate: lanes sharing an id go in rank order
            const unsigned peers = __match_any_sync(0xffffffffu, best_j);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            int max_rank = active ? rank : 0;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) max_rank = max(max_rank, __shfl_xor_sync(0xffffffffu, max_rank, m));
            for (int r = 0; r <= max_rank; r++) {
                if (active && rank == r) {
                    float* row = my_acc + best_j * ROW;
#pragma unroll
                    for (int d = 0; d < D; d++) row[d] += x[d];
                    row[D] += 1.0f;
                }
                __syncwarp();
            }
        }
    }
    if (partials) {
        __syncthreads();
        float* out = partials + (size_t)blockIdx.x * k * ROW;
        for (int e = threadIdx.x; e < k * ROW; e += KM_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) s += s_acc[(size_t)w * k * ROW + e];
            out[e] = s;
        }
    }
}

__global__ void kmeans_reduce_partials_kernel(int nblocks, int k, int D, const float* __restrict__ partials,
                                              float* __restrict__ sums, float* __restrict__ counts) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int ROW = D + 1;
    if (e >= k * ROW) return;
    float s = 0.f;
    for (int bk = 0; bk < nblocks; bk++) s += partials[(size_t)bk * k * ROW + e];
    const int j = e / ROW, d = e - j * ROW;
    if (d < D) { if (sums) sums[j * D + d] += s; }
    else if (counts) counts[j] += s;
}

template <int D>
static int launch_assign_d(int64_t N, const float* a, int Da, const float* b, int Db, float scale_b,
                           const float* centers, int k, const int64_t* select_ids, int64_t selected,
                           int64_t id_offset, int64_t* ids_out, float* sums, float* counts, cudaStream_t s) {
    const bool fuse = (sums != nullptr) || (counts != nullptr);
    size_t smem = (size_t)k * D * sizeof(float);
    if (fuse) smem += (size_t)8 * k * (D + 1) * sizeof(float);
    if (smem > 200 * 1024) { set_error("kmeans_assign: k=%d D=%d needs %zu B shared memory", k, D, smem); return -5; }
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        OGS_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    int64_t want = (N + KM_THREADS - 1) / KM_THREADS;
    int grid = (int)(want < (int64_t)OGS_NUM_SMS * 4 ? want : (int64_t)OGS_NUM_SMS * 4);
    if (grid < 1) grid = 1;
    float* partials = nullptr;
    if (fuse) OGS_CUDA(cudaMallocAsync((void**)&partials, (size_t)grid * k * (D + 1) * sizeof(float), s));
    kmeans_assign_kernel<D><<<grid, KM_THREADS, smem, s>>>(N, a, Da, b, Db, scale_b, centers, k, select_ids, selected,
                                                           id_offset, ids_out, partials);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && fuse) {
        const int tot = k * (D + 1);
        kmeans_reduce_partials_kernel<<<(tot + 127) / 128, 128, 0, s>>>(grid, k, D, partials, sums, counts);
        e = cudaGetLastError();
    }
    if (partials) cudaFreeAsync(partials, s);
    if (e != cudaSuccess) return cuda_fail(e, "kmeans_assign");
    return 0;
}

int launch_kmeans_assign(int64_t N, const float* a, int Da, const float* b, int Db, float scale_b,
                         const float* centers, int k, const int64_t* select_ids, int64_t selected, int64_t id_offset,
                         int64_t* ids_out, float* sums, float* counts, cudaStream_t s) {
    const int D = Da + Db;
#define OGS_KM_CASE(DD) case DD: return launch_assign_d<DD>(N, a, Da, b, Db, scale_b, centers, k, select_ids, selected, id_offset, ids_out, sums, counts, s);
    switch (D) {
        OGS_KM_CASE(1) OGS_KM_CASE(2) OGS_KM_CASE(3) OGS_KM_CASE(4) OGS_KM_CASE(5) OGS_KM_CASE(6) OGS_KM_CASE(7)
        OGS_KM_CASE(8) OGS_KM_CASE(9) OGS_KM_CASE(10) OGS_KM_CASE(11) OGS_KM_CASE(12) OGS_KM_CASE(13)
        OGS_KM_CASE(14) OGS_KM_CASE(15) OGS_KM_CASE(16)
    }
#undef OGS_KM_CASE
    set_error("kmeans_assign: unsupported point dimension %d (1..16)", D);
    return -5;
}

__global__ void kmeans_finalize_kernel(int k, int D, const float* __restrict__ sums, const float* __restrict__ counts,
                                       float eps, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= k * D) return;
    const int j = e / D;
    out[e] = __fdiv_rn(sums[e], __fadd_rn(counts[j], eps));
}

int launch_kmeans_finalize(int k, int D, const float* sums, const float* counts, float eps, float* out, cudaStream_t s) {
    if (k * D <= 0) return 0;
    kmeans_finalize_kernel<<<(k * D + 127) / 128, 128, 0, s>>>(k, D, sums, counts, eps, out);
    return 0;
}

__global__ void kmeans_gather_st_kernel(int64_t N, const float* __restrict__ feat, int Dout,
                                        const float* __restrict__ centers, int Dc, const int64_t* __restrict__ ids,
                                        float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * Dout) return;
    const int64_t i = e / Dout;
    const int d = (int)(e - i * Dout);
    const float f = feat[e];
    out[e] = __fadd_rn(__fsub_rn(f, f), __ldg(centers + ids[i] * Dc + d));
}

int launch_kmeans_gather_st(int64_t N, const float* feat, int Dout, const float* centers, int Dc, const int64_t* ids,
                            float* out, cudaStream_t s) {
    const int64_t tot = N * Dout;
    if (tot <= 0) return 0;
    kmeans_gather_st_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(N, feat, Dout, centers, Dc, ids, out);
    return 0;
}

__global__ void kmeans_count_kernel(int64_t N, const int64_t* __restrict__ ids, int k, unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int s_cnt[];
    for (int e = threadIdx.x; e < k; e += blockDim.x) s_cnt[e] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = ids[i];
        if (id >= 0 && id < k) atomicAdd(&s_cnt[id], 1u);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < k; e += blockDim.x)
        if (s_cnt[e]) atomicAdd(&counts[e], (unsigned long long)s_cnt[e]);
}

int launch_kmeans_count(int64_t N, const int64_t* ids, int k, int64_t* counts, cudaStream_t s) {
    if (k <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)k * 8, s);
    if (e != cudaSuccess) return cuda_fail(e, "memset counts");
    if (N <= 0) return 0;
    if ((size_t)k * 4 > 48 * 1024) { set_error("kmeans_count: k=%d too large", k); return -5; }
    int64_t want = (N + 255) / 256;
    int grid = (int)(want < OGS_NUM_SMS * 8 ? want : OGS_NUM_SMS * 8);
    kmeans_count_kernel<<<grid, 256, (size_t)k * 4, s>>>(N, ids, k, (unsigned long long*)counts);
    return 0;
}

}  // namespace ogs
