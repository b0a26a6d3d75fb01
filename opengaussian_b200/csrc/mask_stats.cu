// mask_stats.cu -- per-mask feature statistics and the Stage-1 cohesion loss as streaming
// segmented reductions.
//
// Replaces utils/opengs_utlis.py::mask_feature_mean (:240-283, with its chunked helpers :203-238) and
// train.py::cohesion_loss (:102-121).  The reference expands feat_map [C,H,W] and gt_masks [M,H,W] to
// [M,C,H,W] float tensors (GBs per step, Python loops over 5x5 chunks); here every pass streams the
// M*H*W mask BYTES once (plus 4(C+1) B per pixel) -- HBM-bound:
//   mask_mean_fwd    : sums[m][c] = sum_p feat[c][p] w[m][p],  counts[m] = sum_p w[m][p],  w = mask * image_mask
//   mask_mean_bwd    : dfeat[c][p] = img[p] sum_m mask[m][p] G[m][c];  dimg[p] = sum_m mask[m][p] (<G[m], feat[:,p]> - K[m])
//   mask_var_fwd     : sq[m][c] = sum_p mask[m][p] (feat[c][p] w - mean[m][c])^2          (return_var=True, :270-283)
//   cohesion_fwd     : dsum[m] = sum_p mask[m][p] ||feat[:,p] - mean[m]||_2,  n[m] = sum_p mask[m][p]
//   cohesion_bwd     : dfeat[c][p] = sum_m mask coef[m] (feat - mean[m][c]) / dist;  dmean[m][c] = -sum_p (same)
// Layout: a warp owns 128 consecutive pixels (4 per lane: one 32-bit load brings a lane's 4 mask
// bytes); the features of its pixels stay in registers while it walks the M masks.  SAM masks are
// spatially compact, so for most masks the warp's 128 bytes are all zero: one coalesced load + one
// vote, nothing else.  Present masks are warp-reduced into a per-CTA shared table that is flushed
// with one red.global per value.
#include "common.cuh"

namespace ogs {

#define MS_THREADS 256
#define MS_PPL 4                                  // pixels per lane
#define MS_CTA_PIX (MS_THREADS * MS_PPL)          // 1024

struct PixelBlock {
    int64_t p0;       // first pixel of the lane
    int valid;        // number of valid pixels (0..4)
};

__device__ __forceinline__ PixelBlock lane_pixels(int64_t HW) {
    PixelBlock b;
    b.p0 = ((int64_t)blockIdx.x * MS_THREADS + threadIdx.x) * MS_PPL;
    const int64_t rem = HW - b.p0;
    b.valid = rem >= MS_PPL ? MS_PPL : (rem > 0 ? (int)rem : 0);
    return b;
}

// 4 mask bytes of the lane for mask m (bit 8j set <=> pixel j in the mask); rows need not be 4-byte aligned
__device__ __forceinline__ uint32_t load_mask4(const uint8_t* __restrict__ row, const PixelBlock& b, bool aligned) {
    if (b.valid == MS_PPL && aligned) return __ldg(reinterpret_cast<const uint32_t*>(row + b.p0));
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < MS_PPL; j++)
        if (j < b.valid) v |= (uint32_t)(row[b.p0 + j] != 0) << (8 * j);
    return v;
}
__device__ __forceinline__ bool in_mask(uint32_t bits, int j) { return ((bits >> (8 * j)) & 0xFFu) != 0; }

// Walks the masks MS_UNROLL at a time: the (independent, coalesced) mask-word loads of a group are all
// issued before the first vote, so the walk is bound by bandwidth and not by one load latency per mask.
// The body runs for masks that have at least one pixel among the warp's 128; close with MS_END_FOR.
#define MS_UNROLL 8
#define MS_FOR_EACH_PRESENT_MASK(m, bits)                                                     \
    for (int m##_0 = 0; m##_0 < M; m##_0 += MS_UNROLL) {                                      \
        uint32_t bits##_g[MS_UNROLL];                                                         \
        _Pragma("unroll") for (int u = 0; u < MS_UNROLL; u++)                                 \
            bits##_g[u] = (m##_0 + u < M) ? load_mask4(masks + (size_t)(m##_0 + u) * HW, b, aligned) : 0u; \
        _Pragma("unroll") for (int u = 0; u < MS_UNROLL; u++) {                               \
            const uint32_t bits = bits##_g[u];                                                \
            const int m = m##_0 + u;                                                          \
            if (!__any_sync(0xffffffffu, bits != 0)) continue;
#define MS_END_FOR }

template <int C>
__device__ __forceinline__ void load_feat(const float* __restrict__ feat, const float* __restrict__ img, int64_t HW,
                                          const PixelBlock& b, float (&f)[MS_PPL][C], float (&w)[MS_PPL]) {
#pragma unroll
    for (int j = 0; j < MS_PPL; j++) {
        const bool ok = j < b.valid;
        w[j] = ok ? (img ? __ldg(img + b.p0 + j) : 1.0f) : 0.0f;
#pragma unroll
        for (int c = 0; c < C; c++) f[j][c] = ok ? __ldg(feat + (size_t)c * HW + b.p0 + j) : 0.0f;
    }
}

template <int NV>
__device__ __forceinline__ void warp_sum(float (&v)[NV]) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] += __shfl_xor_sync(0xffffffffu, v[k], m);
}

// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(MS_THREADS) mask_mean_fwd_kernel(int M, int64_t HW, const float* __restrict__ feat,
                                                                   const uint8_t* __restrict__ masks, const float* __restrict__ img,
                                                                   float* __restrict__ sums, float* __restrict__ counts) {
    extern __shared__ float s_acc[];   // [M][C+1]
    for (int e = threadIdx.x; e < M * (C + 1); e += MS_THREADS) s_acc[e] = 0.f;
    __syncthreads();
    const PixelBlock b = lane_pixels(HW);
    const bool aligned = (HW & 3) == 0;
    float f[MS_PPL][C], w[MS_PPL];
    load_feat<C>(feat, img, HW, b, f, w);
    const int lane = threadIdx.x & 31;
    MS_FOR_EACH_PRESENT_MASK(m, bits)
        float v[C + 1];
#pragma unroll
        for (int k = 0; k <= C; k++) v[k] = 0.f;
#pragma unroll
        for (int j = 0; j < MS_PPL; j++) {
            const float wj = in_mask(bits, j) ? w[j] : 0.f;
#pragma unroll
            for (int c = 0; c < C; c++) v[c] = fmaf(f[j][c], wj, v[c]);
            v[C] += wj;
        }
        warp_sum<C + 1>(v);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k <= C; k++) atomicAdd(&s_acc[m * (C + 1) + k], v[k]);
        }
    }
    MS_END_FOR
    __syncthreads();
    for (int e = threadIdx.x; e < M * (C + 1); e += MS_THREADS) {
        const float s = s_acc[e];
        if (s != 0.f) {
            const int m = e / (C + 1), k = e - m * (C + 1);
            if (k < C) atomicAdd(sums + (size_t)m * C + k, s);
            else atomicAdd(counts + m, s);
        }
    }
}

// G[m][c] = dL/dmean[m][c] / max(count[m], 1);  K[m] = (count[m] > 1) ? sum_c G[m][c] mean[m][c] : 0
template <int C>
__global__ void __launch_bounds__(MS_THREADS) mask_mean_bwd_kernel(int M, int64_t HW, const float* __restrict__ feat,
                                                                   const uint8_t* __restrict__ masks, const float* __restrict__ img,
                                                                   const float* __restrict__ G, const float* __restrict__ K,
                                                                   float* __restrict__ dfeat, float* __restrict__ dimg) {
    extern __shared__ float s_g[];   // [M][C+1]: G then K
    for (int e = threadIdx.x; e < M * (C + 1); e += MS_THREADS) {
        const int m = e / (C + 1), k = e - m * (C + 1);
        s_g[e] = k < C ? G[(size_t)m * C + k] : K[m];
    }
    __syncthreads();
    const PixelBlock b = lane_pixels(HW);
    const bool aligned = (HW & 3) == 0;
    float f[MS_PPL][C], w[MS_PPL];
    load_feat<C>(feat, img, HW, b, f, w);
    float acc[MS_PPL][C], ai[MS_PPL];
#pragma unroll
    for (int j = 0; j < MS_PPL; j++) {
        ai[j] = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) acc[j][c] = 0.f;
    }
    MS_FOR_EACH_PRESENT_MASK(m, bits)
        const float* g = s_g + m * (C + 1);
#pragma unroll
        for (int j = 0; j < MS_PPL; j++) {
            if (in_mask(bits, j)) {
                float dot = -g[C];
#pragma unroll
                for (int c = 0; c < C; c++) {
                    acc[j][c] += g[c];
                    dot = fmaf(g[c], f[j][c], dot);
                }
                ai[j] += dot;
            }
        }
    }
    MS_END_FOR
#pragma unroll
    for (int j = 0; j < MS_PPL; j++) {
        if (j < b.valid) {
#pragma unroll
            for (int c = 0; c < C; c++) dfeat[(size_t)c * HW + b.p0 + j] = acc[j][c] * w[j];
            if (dimg) dimg[b.p0 + j] = ai[j];
        }
    }
}

// sq[m][c] = sum_p mask (feat w - mean[m][c])^2   (the reference squares masked_feats - mean inside the mask)
template <int C>
__global__ void __launch_bounds__(MS_THREADS) mask_var_fwd_kernel(int M, int64_t HW, const float* __restrict__ feat,
                                                                  const uint8_t* __restrict__ masks, const float* __restrict__ img,
                                                                  const float* __restrict__ mean, float* __restrict__ sq) {
    extern __shared__ float s_mem[];   // [M][C] mean, [M][C] acc
    float* s_mean = s_mem;
    float* s_acc = s_mem + (size_t)M * C;
    for (int e = threadIdx.x; e < M * C; e += MS_THREADS) { s_mean[e] = mean[e]; s_acc[e] = 0.f; }
    __syncthreads();
    const PixelBlock b = lane_pixels(HW);
    const bool aligned = (HW & 3) == 0;
    float f[MS_PPL][C], w[MS_PPL];
    load_feat<C>(feat, img, HW, b, f, w);
    const int lane = threadIdx.x & 31;
    MS_FOR_EACH_PRESENT_MASK(m, bits)
        float v[C];
#pragma unroll
        for (int c = 0; c < C; c++) v[c] = 0.f;
#pragma unroll
        for (int j = 0; j < MS_PPL; j++) {
            if (in_mask(bits, j)) {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const float d = f[j][c] * w[j] - s_mean[m * C + c];
                    v[c] = fmaf(d, d, v[c]);
                }
            }
        }
        warp_sum<C>(v);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; c++) atomicAdd(&s_acc[m * C + c], v[c]);
        }
    }
    MS_END_FOR
    __syncthreads();
    for (int e = threadIdx.x; e < M * C; e += MS_THREADS)
        if (s_acc[e] != 0.f) atomicAdd(sq + e, s_acc[e]);
}

template <int C>
__global__ void __launch_bounds__(MS_THREADS) cohesion_fwd_kernel(int M, int64_t HW, const float* __restrict__ feat,
                                                                  const uint8_t* __restrict__ masks, const float* __restrict__ mean,
                                                                  float* __restrict__ dsum, float* __restrict__ npix) {
    extern __shared__ float s_mem[];   // [M][C] mean, [M][2] acc
    float* s_mean = s_mem;
    float* s_acc = s_mem + (size_t)M * C;
    for (int e = threadIdx.x; e < M * C; e += MS_THREADS) s_mean[e] = mean[e];
    for (int e = threadIdx.x; e < M * 2; e += MS_THREADS) s_acc[e] = 0.f;
    __syncthreads();
    const PixelBlock b = lane_pixels(HW);
    const bool aligned = (HW & 3) == 0;
    float f[MS_PPL][C], w[MS_PPL];
    load_feat<C>(feat, nullptr, HW, b, f, w);
    const int lane = threadIdx.x & 31;
    MS_FOR_EACH_PRESENT_MASK(m, bits)
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < MS_PPL; j++) {
            if (in_mask(bits, j)) {
                float q = 0.f;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const float d = f[j][c] - s_mean[m * C + c];
                    q = fmaf(d, d, q);
                }
                v[0] += sqrtf(q);
                v[1] += 1.0f;
            }
        }
        warp_sum<2>(v);
        if (lane == 0) {
            atomicAdd(&s_acc[2 * m], v[0]);
            atomicAdd(&s_acc[2 * m + 1], v[1]);
        }
    }
    MS_END_FOR
    __syncthreads();
    for (int e = threadIdx.x; e < M; e += MS_THREADS) {
        if (s_acc[2 * e + 1] != 0.f) {
            atomicAdd(dsum + e, s_acc[2 * e]);
            atomicAdd(npix + e, s_acc[2 * e + 1]);
        }
    }
}

// coef[m] = dL/dloss / (M max(n[m], 1)).  dfeat is WRITTEN (every pixel), dmean is accumulated (zeroed by the launcher).
template <int C>
__global__ void __launch_bounds__(MS_THREADS) cohesion_bwd_kernel(int M, int64_t HW, const float* __restrict__ feat,
                                                                  const uint8_t* __restrict__ masks, const float* __restrict__ mean,
                                                                  const float* __restrict__ coef, float* __restrict__ dfeat,
                                                                  float* __restrict__ dmean) {
    extern __shared__ float s_mem[];   // [M][C+1] mean | coef, [M][C] acc
    float* s_mean = s_mem;
    float* s_acc = s_mem + (size_t)M * (C + 1);
    for (int e = threadIdx.x; e < M * (C + 1); e += MS_THREADS) {
        const int m = e / (C + 1), k = e - m * (C + 1);
        s_mean[e] = k < C ? mean[(size_t)m * C + k] : coef[m];
    }
    for (int e = threadIdx.x; e < M * C; e += MS_THREADS) s_acc[e] = 0.f;
    __syncthreads();
    const PixelBlock b = lane_pixels(HW);
    const bool aligned = (HW & 3) == 0;
    float f[MS_PPL][C], w[MS_PPL];
    load_feat<C>(feat, nullptr, HW, b, f, w);
    float acc[MS_PPL][C];
#pragma unroll
    for (int j = 0; j < MS_PPL; j++)
#pragma unroll
        for (int c = 0; c < C; c++) acc[j][c] = 0.f;
    const int lane = threadIdx.x & 31;
    MS_FOR_EACH_PRESENT_MASK(m, bits)
        const float* mu = s_mean + m * (C + 1);
        float v[C];
#pragma unroll
        for (int c = 0; c < C; c++) v[c] = 0.f;
#pragma unroll
        for (int j = 0; j < MS_PPL; j++) {
            if (in_mask(bits, j)) {
                float d[C], q = 0.f;
#pragma unroll
                for (int c = 0; c < C; c++) { d[c] = f[j][c] - mu[c]; q = fmaf(d[c], d[c], q); }
                const float s = q > 0.f ? mu[C] * rsqrtf(q) : 0.f;   // d ||x|| = x / ||x||, 0 at the origin (torch's convention)
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const float r = d[c] * s;
                    acc[j][c] += r;
                    v[c] -= r;
                }
            }
        }
        warp_sum<C>(v);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; c++) atomicAdd(&s_acc[m * C + c], v[c]);
        }
    }
    MS_END_FOR
#pragma unroll
    for (int j = 0; j < MS_PPL; j++)
        if (j < b.valid) {
#pragma unroll
            for (int c = 0; c < C; c++) dfeat[(size_t)c * HW + b.p0 + j] = acc[j][c];
        }
    __syncthreads();
    for (int e = threadIdx.x; e < M * C; e += MS_THREADS)
        if (s_acc[e] != 0.f) atomicAdd(dmean + e, s_acc[e]);
}

// ---------------------------------------------------------------------------------------------
static int check_shapes(int M, int C, int64_t HW, size_t smem_floats, const char* what) {
    if (M < 0 || HW < 0) { set_error("%s: bad sizes M=%d HW=%lld", what, M, (long long)HW); return -1; }
    if (C != 3 && C != 6) { set_error("%s: unsupported channel count %d (3 or 6)", what, C); return -4; }
    if (smem_floats * 4 > 200 * 1024) { set_error("%s: %d masks need %zu B of shared memory (max 200 KB)", what, M, smem_floats * 4); return -5; }
    return 0;
}

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
    if (bytes > 48 * 1024) OGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

#define MS_GRID(HW) (unsigned)(((HW) + MS_CTA_PIX - 1) / MS_CTA_PIX)
#define MS_DISPATCH(CALL3, CALL6) \
    do {                          \
        if (C == 3) { CALL3; }    \
        else { CALL6; }           \
    } while (0)

int launch_mask_mean_forward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const float* img, float* sums,
                             float* counts, cudaStream_t s) {
    const size_t sm = (size_t)M * (C + 1);
    int rc = check_shapes(M, C, HW, sm, "mask_mean_forward");
    if (rc) return rc;
    OGS_CUDA(cudaMemsetAsync(sums, 0, (size_t)M * C * 4, s));
    OGS_CUDA(cudaMemsetAsync(counts, 0, (size_t)M * 4, s));
    if (M == 0 || HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(mask_mean_fwd_kernel<3>, sm * 4), mask_mean_fwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, img, sums, counts)),
                (rc = set_smem(mask_mean_fwd_kernel<6>, sm * 4), mask_mean_fwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, img, sums, counts)));
    return rc;
}

int launch_mask_mean_backward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const float* img, const float* G,
                              const float* K, float* dfeat, float* dimg, cudaStream_t s) {
    const size_t sm = (size_t)M * (C + 1);
    int rc = check_shapes(M, C, HW, sm, "mask_mean_backward");
    if (rc) return rc;
    if (HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(mask_mean_bwd_kernel<3>, sm * 4), mask_mean_bwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, img, G, K, dfeat, dimg)),
                (rc = set_smem(mask_mean_bwd_kernel<6>, sm * 4), mask_mean_bwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, img, G, K, dfeat, dimg)));
    return rc;
}

int launch_mask_var_forward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const float* img, const float* mean,
                            float* sq, cudaStream_t s) {
    const size_t sm = (size_t)M * C * 2;
    int rc = check_shapes(M, C, HW, sm, "mask_var_forward");
    if (rc) return rc;
    OGS_CUDA(cudaMemsetAsync(sq, 0, (size_t)M * C * 4, s));
    if (M == 0 || HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(mask_var_fwd_kernel<3>, sm * 4), mask_var_fwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, img, mean, sq)),
                (rc = set_smem(mask_var_fwd_kernel<6>, sm * 4), mask_var_fwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, img, mean, sq)));
    return rc;
}

int launch_cohesion_forward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const float* mean, float* dsum,
                            float* npix, cudaStream_t s) {
    const size_t sm = (size_t)M * (C + 2);
    int rc = check_shapes(M, C, HW, sm, "cohesion_forward");
    if (rc) return rc;
    OGS_CUDA(cudaMemsetAsync(dsum, 0, (size_t)M * 4, s));
    OGS_CUDA(cudaMemsetAsync(npix, 0, (size_t)M * 4, s));
    if (M == 0 || HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(cohesion_fwd_kernel<3>, sm * 4), cohesion_fwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, mean, dsum, npix)),
                (rc = set_smem(cohesion_fwd_kernel<6>, sm * 4), cohesion_fwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, mean, dsum, npix)));
    return rc;
}

int launch_cohesion_backward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const float* mean, const float* coef,
                             float* dfeat, float* dmean, cudaStream_t s) {
    const size_t sm = (size_t)M * (2 * C + 1);
    int rc = check_shapes(M, C, HW, sm, "cohesion_backward");
    if (rc) return rc;
    OGS_CUDA(cudaMemsetAsync(dmean, 0, (size_t)M * C * 4, s));
    if (HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(cohesion_bwd_kernel<3>, sm * 4), cohesion_bwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, mean, coef, dfeat, dmean)),
                (rc = set_smem(cohesion_bwd_kernel<6>, sm * 4), cohesion_bwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, mean, coef, dfeat, dmean)));
    return rc;
}

}  // namespace ogs
