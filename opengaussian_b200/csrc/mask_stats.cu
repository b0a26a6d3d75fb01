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
// body(m, bits) runs for masks that have at least one pixel among the warp's 128.
//
// When the masks are a PARTITION of the image (every pixel in at most one mask -- what get_SAM_mask_and_feat builds
// from a SAM id map, utils/opengs_utlis.py:125-182), a per-pixel id map (int16, -1 = no mask) replaces the M mask
// rows: the warp reads 2 B per pixel instead of M, and visits only the ids that occur among its 128 pixels.
#define MS_UNROLL 8
template <typename Body>
__device__ __forceinline__ void for_each_present_mask(int M, int64_t HW, const uint8_t* __restrict__ masks,
                                                      const int16_t* __restrict__ ids, const PixelBlock& b, bool aligned,
                                                      Body body) {
    if (ids) {
        int id0 = -1, id1 = -1, id2 = -1, id3 = -1;
        if (b.valid == MS_PPL) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(ids + b.p0));      // p0 is a multiple of 4: 8-byte aligned
            id0 = (int16_t)(v.x & 0xFFFFu); id1 = (int16_t)(v.x >> 16);
            id2 = (int16_t)(v.y & 0xFFFFu); id3 = (int16_t)(v.y >> 16);
        } else {
            if (b.valid > 0) id0 = ids[b.p0];
            if (b.valid > 1) id1 = ids[b.p0 + 1];
            if (b.valid > 2) id2 = ids[b.p0 + 2];
        }
        uint32_t todo = (id0 >= 0 ? 1u : 0u) | (id1 >= 0 ? 2u : 0u) | (id2 >= 0 ? 4u : 0u) | (id3 >= 0 ? 8u : 0u);
        for (;;) {
            const unsigned ball = __ballot_sync(0xffffffffu, todo != 0);
            if (!ball) break;
            const int mine = (todo & 1u) ? id0 : (todo & 2u) ? id1 : (todo & 4u) ? id2 : id3;
            const int m = __shfl_sync(0xffffffffu, mine, __ffs(ball) - 1);
            uint32_t bits = 0;
            if ((todo & 1u) && id0 == m) { bits |= 1u; todo &= ~1u; }
            if ((todo & 2u) && id1 == m) { bits |= 1u << 8; todo &= ~2u; }
            if ((todo & 4u) && id2 == m) { bits |= 1u << 16; todo &= ~4u; }
            if ((todo & 8u) && id3 == m) { bits |= 1u << 24; todo &= ~8u; }
            if (m < M) body(m, bits);
        }
        return;
    }
    for (int m0 = 0; m0 < M; m0 += MS_UNROLL) {
        uint32_t g[MS_UNROLL];
#pragma unroll
        for (int u = 0; u < MS_UNROLL; u++) g[u] = (m0 + u < M) ? load_mask4(masks + (size_t)(m0 + u) * HW, b, aligned) : 0u;
#pragma unroll
        for (int u = 0; u < MS_UNROLL; u++)
            if (__any_sync(0xffffffffu, g[u] != 0)) body(m0 + u, g[u]);
    }
}
// the id map is used only when the device-side flag says the masks it was built from did not overlap
__device__ __forceinline__ const int16_t* usable_ids(const int16_t* ids, const int32_t* overlap) {
    return (ids && overlap && *overlap == 0) ? ids : nullptr;
}

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
                                                                   const uint8_t* __restrict__ masks, const int16_t* __restrict__ ids,
        const int32_t* __restrict__ overlap, const float* __restrict__ img,
                                                                   float* __restrict__ sums, float* __restrict__ counts) {
    extern __shared__ float s_acc[];   // [M][C+1]
    for (int e = threadIdx.x; e < M * (C + 1); e += MS_THREADS) s_acc[e] = 0.f;
    __syncthreads();
    const PixelBlock b = lane_pixels(HW);
    const bool aligned = (HW & 3) == 0;
    float f[MS_PPL][C], w[MS_PPL];
    load_feat<C>(feat, img, HW, b, f, w);
    const int lane = threadIdx.x & 31;
    for_each_present_mask(M, HW, masks, usable_ids(ids, overlap), b, aligned, [&](int m, uint32_t bits) {
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
    });
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
                                                                   const uint8_t* __restrict__ masks, const int16_t* __restrict__ ids,
        const int32_t* __restrict__ overlap, const float* __restrict__ img,
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
    for_each_present_mask(M, HW, masks, usable_ids(ids, overlap), b, aligned, [&](int m, uint32_t bits) {
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
    });
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
                                                                  const uint8_t* __restrict__ masks, const int16_t* __restrict__ ids,
        const int32_t* __restrict__ overlap, const float* __restrict__ img,
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
    for_each_present_mask(M, HW, masks, usable_ids(ids, overlap), b, aligned, [&](int m, uint32_t bits) {
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
    });
    __syncthreads();
    for (int e = threadIdx.x; e < M * C; e += MS_THREADS)
        if (s_acc[e] != 0.f) atomicAdd(sq + e, s_acc[e]);
}

template <int C>
__global__ void __launch_bounds__(MS_THREADS) cohesion_fwd_kernel(int M, int64_t HW, const float* __restrict__ feat,
                                                                  const uint8_t* __restrict__ masks, const int16_t* __restrict__ ids,
        const int32_t* __restrict__ overlap, const float* __restrict__ mean,
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
    for_each_present_mask(M, HW, masks, usable_ids(ids, overlap), b, aligned, [&](int m, uint32_t bits) {
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
    });
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
                                                                  const uint8_t* __restrict__ masks, const int16_t* __restrict__ ids,
        const int32_t* __restrict__ overlap, const float* __restrict__ mean,
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
    for_each_present_mask(M, HW, masks, usable_ids(ids, overlap), b, aligned, [&](int m, uint32_t bits) {
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
    });
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

// ids[p] = the mask that holds pixel p (-1: none); *overlap is set when some pixel sits in two or more masks, in which
// case the statistics kernels ignore the id map and walk the mask rows.
__global__ void __launch_bounds__(MS_THREADS) mask_id_map_kernel(int M, int64_t HW, const uint8_t* __restrict__ masks,
                                                                 int16_t* __restrict__ ids, int32_t* __restrict__ overlap) {
    const PixelBlock b = lane_pixels(HW);
    const bool aligned = (HW & 3) == 0;
    int id[MS_PPL], n[MS_PPL];
#pragma unroll
    for (int j = 0; j < MS_PPL; j++) { id[j] = -1; n[j] = 0; }
    for_each_present_mask(M, HW, masks, nullptr, b, aligned, [&](int m, uint32_t bits) {
#pragma unroll
        for (int j = 0; j < MS_PPL; j++)
            if (in_mask(bits, j)) { id[j] = m; n[j]++; }
    });
    bool clash = false;
#pragma unroll
    for (int j = 0; j < MS_PPL; j++) {
        if (j < b.valid) ids[b.p0 + j] = (int16_t)id[j];
        clash |= n[j] > 1;
    }
    if (__any_sync(0xffffffffu, clash) && (threadIdx.x & 31) == 0) atomicOr(overlap, 1);
}

// get_SAM_mask_and_feat (utils/opengs_utlis.py:134-148,168-169) per pixel: v = max(level_id - offset, -1);
// mask_id = v + 1 (0 = invalid), invalid_pix = (v < 0), ids = v (or -1 when v >= M: in none of the M masks)
__global__ void __launch_bounds__(256) sam_ids_kernel(int M, int64_t HW, const int32_t* __restrict__ level_ids, int offset,
                                                      int64_t* __restrict__ mask_id, uint8_t* __restrict__ invalid_pix,
                                                      int16_t* __restrict__ ids) {
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= HW) return;
    int v = __ldg(level_ids + p) - offset;
    v = v < -1 ? -1 : v;
    mask_id[p] = (int64_t)v + 1;
    invalid_pix[p] = v < 0;
    ids[p] = (int16_t)(v < M ? v : -1);
}

// masks[m][p] = (ids[p] == m): the one-hot expansion, 16 pixels per thread (one 16-byte store when rows are aligned)
__global__ void __launch_bounds__(256) mask_expand_kernel(int64_t HW, const int16_t* __restrict__ ids, uint8_t* __restrict__ masks) {
    const int64_t p = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
    if (p >= HW) return;
    const int m = blockIdx.y;
    uint8_t* row = masks + (size_t)m * HW;
    if (p + 16 <= HW && (HW & 15) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(ids + p));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(ids + p) + 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t lo = w[2 * k], hi = w[2 * k + 1];
            o[k] = ((int)(int16_t)(lo & 0xFFFFu) == m ? 1u : 0u) | ((int)(int16_t)(lo >> 16) == m ? 1u << 8 : 0u) |
                   ((int)(int16_t)(hi & 0xFFFFu) == m ? 1u << 16 : 0u) | ((int)(int16_t)(hi >> 16) == m ? 1u << 24 : 0u);
        }
        *reinterpret_cast<uint4*>(row + p) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        for (int64_t q = p; q < p + 16 && q < HW; q++) row[q] = (int)ids[q] == m;
    }
}

// The same expansion with the comparisons done four pixels at a time (M <= 254: an id fits a byte, "no mask" = 0xFF):
// a thread packs its 16 ids into four words once and then, per mask of its slice, finds the equal bytes with the
// exact zero-byte test  ~(((x & 0x7F7F7F7F) + 0x7F7F7F7F) | x | 0x7F7F7F7F) >> 7  of  x = word ^ (m * 0x01010101)  (the
// shorter (x - 0x01010101) & ~x form lets a borrow flag a 0x01 byte above a zero byte) -- one ALU operation per
// output byte instead of two (ncu: the per-mask kernel above is ALU-bound at 85 % of the pipe, 42 us for 120 masks
// at 1296x968), and the ids are read once per MX_SLICES masks instead of once per mask.
#define MX_SLICES 4
__global__ void __launch_bounds__(256) mask_expand_bytes_kernel(int M, int64_t HW, const int16_t* __restrict__ ids,
                                                                uint8_t* __restrict__ masks) {
    const int64_t p = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
    if (p >= HW) return;                       // launcher guarantees HW % 16 == 0
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(ids + p));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(ids + p) + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t q[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {              // int16 pairs -> bytes (ids are -1 or < 255: the low byte identifies them)
        const uint32_t lo = w[2 * k], hi = w[2 * k + 1];
        q[k] = __byte_perm(lo, hi, 0x6420);
    }
    const int per = (M + MX_SLICES - 1) / MX_SLICES;
    const int m0 = blockIdx.y * per, m1 = min(M, m0 + per);
    uint8_t* row = masks + (size_t)m0 * HW + p;
    for (int m = m0; m < m1; m++, row += HW) {
        const uint32_t rep = (uint32_t)m * 0x01010101u;
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t x = q[k] ^ rep;
            o[k] = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu) >> 7;
        }
        *reinterpret_cast<uint4*>(row) = make_uint4(o[0], o[1], o[2], o[3]);
    }
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

int launch_mask_mean_forward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const int16_t* ids, const int32_t* overlap, const float* img, float* sums,
                             float* counts, cudaStream_t s) {
    const size_t sm = (size_t)M * (C + 1);
    int rc = check_shapes(M, C, HW, sm, "mask_mean_forward");
    if (rc) return rc;
    OGS_CUDA(cudaMemsetAsync(sums, 0, (size_t)M * C * 4, s));
    OGS_CUDA(cudaMemsetAsync(counts, 0, (size_t)M * 4, s));
    if (M == 0 || HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(mask_mean_fwd_kernel<3>, sm * 4), mask_mean_fwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, img, sums, counts)),
                (rc = set_smem(mask_mean_fwd_kernel<6>, sm * 4), mask_mean_fwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, img, sums, counts)));
    return rc;
}

int launch_mask_mean_backward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const int16_t* ids, const int32_t* overlap, const float* img, const float* G,
                              const float* K, float* dfeat, float* dimg, cudaStream_t s) {
    const size_t sm = (size_t)M * (C + 1);
    int rc = check_shapes(M, C, HW, sm, "mask_mean_backward");
    if (rc) return rc;
    if (HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(mask_mean_bwd_kernel<3>, sm * 4), mask_mean_bwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, img, G, K, dfeat, dimg)),
                (rc = set_smem(mask_mean_bwd_kernel<6>, sm * 4), mask_mean_bwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, img, G, K, dfeat, dimg)));
    return rc;
}

int launch_mask_var_forward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const int16_t* ids, const int32_t* overlap, const float* img, const float* mean,
                            float* sq, cudaStream_t s) {
    const size_t sm = (size_t)M * C * 2;
    int rc = check_shapes(M, C, HW, sm, "mask_var_forward");
    if (rc) return rc;
    OGS_CUDA(cudaMemsetAsync(sq, 0, (size_t)M * C * 4, s));
    if (M == 0 || HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(mask_var_fwd_kernel<3>, sm * 4), mask_var_fwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, img, mean, sq)),
                (rc = set_smem(mask_var_fwd_kernel<6>, sm * 4), mask_var_fwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, img, mean, sq)));
    return rc;
}

int launch_cohesion_forward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const int16_t* ids, const int32_t* overlap, const float* mean, float* dsum,
                            float* npix, cudaStream_t s) {
    const size_t sm = (size_t)M * (C + 2);
    int rc = check_shapes(M, C, HW, sm, "cohesion_forward");
    if (rc) return rc;
    OGS_CUDA(cudaMemsetAsync(dsum, 0, (size_t)M * 4, s));
    OGS_CUDA(cudaMemsetAsync(npix, 0, (size_t)M * 4, s));
    if (M == 0 || HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(cohesion_fwd_kernel<3>, sm * 4), cohesion_fwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, mean, dsum, npix)),
                (rc = set_smem(cohesion_fwd_kernel<6>, sm * 4), cohesion_fwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, mean, dsum, npix)));
    return rc;
}

int launch_cohesion_backward(int M, int C, int64_t HW, const float* feat, const uint8_t* masks, const int16_t* ids, const int32_t* overlap, const float* mean, const float* coef,
                             float* dfeat, float* dmean, cudaStream_t s) {
    const size_t sm = (size_t)M * (2 * C + 1);
    int rc = check_shapes(M, C, HW, sm, "cohesion_backward");
    if (rc) return rc;
    OGS_CUDA(cudaMemsetAsync(dmean, 0, (size_t)M * C * 4, s));
    if (HW == 0) return 0;
    MS_DISPATCH((rc = set_smem(cohesion_bwd_kernel<3>, sm * 4), cohesion_bwd_kernel<3><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, mean, coef, dfeat, dmean)),
                (rc = set_smem(cohesion_bwd_kernel<6>, sm * 4), cohesion_bwd_kernel<6><<<MS_GRID(HW), MS_THREADS, sm * 4, s>>>(M, HW, feat, masks, ids, overlap, mean, coef, dfeat, dmean)));
    return rc;
}

int launch_mask_id_map(int M, int64_t HW, const uint8_t* masks, int16_t* ids, int32_t* overlap, cudaStream_t s) {
    if (M < 0 || HW < 0 || M > 32767) { set_error("mask_id_map: bad sizes M=%d HW=%lld (at most 32767 masks)", M, (long long)HW); return -1; }
    OGS_CUDA(cudaMemsetAsync(overlap, 0, 4, s));
    if (HW == 0) return 0;
    mask_id_map_kernel<<<MS_GRID(HW), MS_THREADS, 0, s>>>(M, HW, masks, ids, overlap);
    return 0;
}

int launch_sam_masks(int M, int64_t HW, const int32_t* level_ids, int offset, int64_t* mask_id, uint8_t* invalid_pix,
                     int16_t* ids, uint8_t* masks, cudaStream_t s) {
    if (M < 0 || HW < 0 || M > 32767) { set_error("sam_masks: bad sizes M=%d HW=%lld (at most 32767 masks)", M, (long long)HW); return -1; }
    if (HW == 0) return 0;
    sam_ids_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, s>>>(M, HW, level_ids, offset, mask_id, invalid_pix, ids);
    if (M > 0) {
        if (M <= 254 && (HW & 15) == 0) {
            const dim3 grid((unsigned)((HW + 16 * 256 - 1) / (16 * 256)), MX_SLICES);
            mask_expand_bytes_kernel<<<grid, 256, 0, s>>>(M, HW, ids, masks);
        } else {
            const dim3 grid((unsigned)((HW + 16 * 256 - 1) / (16 * 256)), (unsigned)M);
            mask_expand_kernel<<<grid, 256, 0, s>>>(HW, ids, masks);
        }
    }
    return 0;
}

}  // namespace ogs
