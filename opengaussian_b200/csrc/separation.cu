// separation.cu -- the inter-mask contrastive loss of Stage 1 (train.py:123-155 `separation_loss`), value and gradient.
//
// The reference evaluates it with ~20 tiny torch kernels on an [N, N] matrix (expand, pow, sum, masked_fill, two
// argsorts, ...) and as many again in the backward; at N = 120 masks that is launch latency, not work.  Here:
//   kernel 1 (one CTA per row i): inv_ij = 1 / (|m_i - m_j|^2 + 1) (0 on the diagonal), the RANK of every inv_ij inside
//            row i by counting (ascending, ties by column: what argsort().argsort() yields), the weight
//            w_ij = rank / (N - 1) * 0.9 + 0.1 (0.1 wherever w < 0.9 once iteration > 35000), the row's loss sum;
//   kernel 2 (one CTA per row i): dL/dm_i = sum_j (w_ij + w_ji) * (-2 inv_ij^2) (m_i - m_j) / (N (N - 1)) -- the weights
//            are rank-derived constants for autograd, as in the reference -- and, in CTA 0, the loss = sum of the row
//            sums in row order / (N (N - 1)).
#include "common.cuh"

namespace ogs {

#define SEP_THREADS 128
#define SEP_MAX_C 16

__global__ void __launch_bounds__(SEP_THREADS) separation_rows_kernel(int N, int C, const float* __restrict__ mean,
                                                                       int small_weights, float* __restrict__ w /*[N][N]*/,
                                                                       float* __restrict__ row_sum /*[N]*/) {
    extern __shared__ float s_inv[];          // [N]
    __shared__ float s_red[SEP_THREADS / 32];
    const int i = blockIdx.x;
    float mi[SEP_MAX_C];
    for (int c = 0; c < C; c++) mi[c] = mean[(size_t)i * C + c];
    for (int j = threadIdx.x; j < N; j += SEP_THREADS) {
        float d2 = 0.f;
        for (int c = 0; c < C; c++) {
            const float d = __fsub_rn(mi[c], mean[(size_t)j * C + c]);
            d2 = __fadd_rn(d2, __fmul_rn(d, d));
        }
        s_inv[j] = (j == i) ? 0.f : __fdiv_rn(1.0f, __fadd_rn(d2, 1.0f));
    }
    __syncthreads();
    const float inv_nm1 = (float)(N - 1);
    float part = 0.f;
    for (int j = threadIdx.x; j < N; j += SEP_THREADS) {
        const float v = s_inv[j];
        int rank = 0;
        for (int k = 0; k < N; k++) {
            const float u = s_inv[k];
            rank += (u < v || (u == v && k < j)) ? 1 : 0;
        }
        float wt = __fadd_rn(__fmul_rn(__fdiv_rn((float)rank, inv_nm1), 0.9f), 0.1f);
        if (small_weights && wt < 0.9f) wt = 0.1f;
        w[(size_t)i * N + j] = wt;
        part = __fadd_rn(part, __fmul_rn(v, wt));
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int q = 0; q < SEP_THREADS / 32; q++) t += s_red[q];
        row_sum[i] = t;
    }
}

__global__ void __launch_bounds__(SEP_THREADS) separation_grad_kernel(int N, int C, const float* __restrict__ mean,
                                                                       const float* __restrict__ w, const float* __restrict__ row_sum,
                                                                       float* __restrict__ loss_out, float* __restrict__ dmean) {
    __shared__ float s_red[SEP_THREADS / 32][SEP_MAX_C];
    const int i = blockIdx.x;
    const float norm = (float)N * (float)(N - 1);
    float mi[SEP_MAX_C], acc[SEP_MAX_C];
    for (int c = 0; c < C; c++) { mi[c] = mean[(size_t)i * C + c]; acc[c] = 0.f; }
    for (int j = threadIdx.x; j < N; j += SEP_THREADS) {
        if (j == i) continue;
        float d[SEP_MAX_C], d2 = 0.f;
        for (int c = 0; c < C; c++) {
            d[c] = __fsub_rn(mi[c], mean[(size_t)j * C + c]);
            d2 = __fadd_rn(d2, __fmul_rn(d[c], d[c]));
        }
        const float inv = __fdiv_rn(1.0f, __fadd_rn(d2, 1.0f));
        const float coef = -2.0f * inv * inv * (w[(size_t)i * N + j] + w[(size_t)j * N + i]) / norm;
        for (int c = 0; c < C; c++) acc[c] = fmaf(coef, d[c], acc[c]);
    }
    for (int c = 0; c < C; c++) {
        float v = acc[c];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
        float t = 0.f;
        for (int q = 0; q < SEP_THREADS / 32; q++) t += s_red[q][threadIdx.x];
        dmean[(size_t)i * C + threadIdx.x] = t;
    }
    if (i == 0 && threadIdx.x == 0) {
        float t = 0.f;
        for (int r = 0; r < N; r++) t += row_sum[r];
        *loss_out = t / norm;
    }
}

int launch_separation_loss(int N, int C, const float* mean, int small_weights, float* scratch, float* loss_out, float* dmean,
                           cudaStream_t s) {
    if (C < 1 || C > SEP_MAX_C) { set_error("separation_loss: C=%d out of range (1..%d)", C, SEP_MAX_C); return -4; }
    if ((size_t)N * 4 > 48 * 1024) { set_error("separation_loss: N=%d too large (max 12288 masks)", N); return -4; }
    float* w = scratch;
    float* row_sum = scratch + (size_t)N * N;
    separation_rows_kernel<<<N, SEP_THREADS, (size_t)N * 4, s>>>(N, C, mean, small_weights, w, row_sum);
    separation_grad_kernel<<<N, SEP_THREADS, 0, s>>>(N, C, mean, w, row_sum, loss_out, dmean);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "separation_loss");
    return 0;
}

}  // namespace ogs
