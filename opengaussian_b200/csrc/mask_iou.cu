// mask_iou.cu -- pairwise intersection counts of two sets of boolean masks.
//
// Replaces utils/opengs_utlis.py::calculate_iou (:90-123).  The reference broadcasts masks1 [n,H,W] and
// masks2 [m,H,W] to [m,n,H,W] twice (AND, OR), converts to float and sums: ~10 m n H W bytes of traffic.
// Here:
//   mask_pack_kernel : every mask row is read ONCE (n+m rows of H*W bytes, HBM-bound) and written as a
//                      bit string, 1 bit per pixel (H*W/8 bytes per row; stays in L2), with its pixel count.
//                      A warp takes 512 pixels per step (one 16-byte load per lane) and turns them into 16
//                      words with 16 ballots; word k holds byte k of every lane, i.e. the pixel order inside a
//                      512-pixel group is permuted -- identically for every row, which is all AND/popc needs.
//   mask_pair_kernel : inter[j][i] = sum_w popc(A[i][w] & B[j][w]) over the packed words; a warp owns one row
//                      of masks1 and a 1024-word slice, keeps 8 rows of masks2 in flight, skips zero words
//                      (SAM masks are spatially compact), one integer red.global per (i, j, slice).
// The union is |a| + |b| - inter, so the counts are exact integers and the IoU computed from them equals the
// reference's float32 sums bit for bit while H*W < 2^24.
#include "common.cuh"

namespace ogs {

#define MI_THREADS 256
#define MI_WARPS (MI_THREADS / 32)
#define MI_GROUP 512        // pixels per warp step
#define MI_CHUNK 1024       // packed words of a row per CTA of the pair kernel
#define MI_JT 8             // rows of masks2 per register tile

static inline int64_t packed_words(int64_t HW) { return ((HW + MI_GROUP - 1) / MI_GROUP) * 16; }

template <bool VEC>
__global__ void __launch_bounds__(MI_THREADS) mask_pack_kernel(int n1, int64_t HW, int64_t W, const uint8_t* __restrict__ m1,
                                                               const uint8_t* __restrict__ m2, uint32_t* __restrict__ packed,
                                                               int32_t* __restrict__ counts) {
    const int r = blockIdx.y;
    const uint8_t* __restrict__ row = r < n1 ? m1 + (int64_t)r * HW : m2 + (int64_t)(r - n1) * HW;
    uint32_t* __restrict__ out = packed + (int64_t)r * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ngroups = W / 16;
    int cnt = 0;
    for (int64_t g = (int64_t)blockIdx.x * MI_WARPS + warp; g < ngroups; g += (int64_t)gridDim.x * MI_WARPS) {
        const int64_t p = g * MI_GROUP + lane * 16;
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        if (VEC) {                                        // H*W % 16 == 0 and 16-byte aligned bases: p < HW => p + 16 <= HW
            if (p < HW) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + p));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (p + k < HW && row[p + k]) v[k >> 2] |= 0xFFu << (8 * (k & 3));
        }
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, ((v[k >> 2] >> (8 * (k & 3))) & 0xFFu) != 0u);
            if (lane == k) mine = b;
        }
        if (lane < 16) {
            out[g * 16 + lane] = mine;
            cnt += __popc(mine);
        }
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0 && cnt) atomicAdd(counts + r, cnt);
}

__global__ void __launch_bounds__(MI_THREADS) mask_pair_kernel(int n1, int n2, int64_t W, const uint32_t* __restrict__ packed,
                                                               int32_t* __restrict__ inter) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.y * MI_WARPS + warp;
    if (i >= n1) return;                                   // warp-uniform; no block-level barrier below
    const int64_t w0 = (int64_t)blockIdx.x * MI_CHUNK;
    const int64_t w1 = (w0 + MI_CHUNK < W) ? w0 + MI_CHUNK : W;
    const uint32_t* __restrict__ A = packed + (int64_t)i * W;
    const uint32_t* __restrict__ B = packed + (int64_t)n1 * W;
    for (int j0 = 0; j0 < n2; j0 += MI_JT) {
        int acc[MI_JT];
#pragma unroll
        for (int jj = 0; jj < MI_JT; ++jj) acc[jj] = 0;
        for (int64_t w = w0 + lane; w < w1; w += 32) {
            const uint32_t a = __ldg(A + w);
            if (a != 0u) {
#pragma unroll
                for (int jj = 0; jj < MI_JT; ++jj) {
                    const int j = (j0 + jj < n2) ? j0 + jj : n2 - 1;      // clamped rows are never written back
                    acc[jj] += __popc(a & __ldg(B + (int64_t)j * W + w));
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < MI_JT; ++jj) {
            const int s = __reduce_add_sync(0xFFFFFFFFu, acc[jj]);
            if (lane == 0 && s != 0 && j0 + jj < n2) atomicAdd(inter + (int64_t)(j0 + jj) * n1 + i, s);
        }
    }
}

int64_t mask_iou_scratch_bytes(int n1, int n2, int64_t HW) {
    return (int64_t)(n1 + n2) * packed_words(HW) * 4;
}

int launch_mask_pair_counts(int n1, int n2, int64_t HW, const uint8_t* masks1, const uint8_t* masks2, uint32_t* scratch,
                            int32_t* inter, int32_t* counts, cudaStream_t s) {
    if (n1 + n2 > 65535) { set_error("mask_pair_counts: at most 65535 masks in total"); return -1; }
    if (HW >= ((int64_t)1 << 31)) { set_error("mask_pair_counts: H*W must be < 2^31"); return -1; }
    if (n1 + n2 > 0) OGS_CUDA(cudaMemsetAsync(counts, 0, (size_t)(n1 + n2) * 4, s));
    if (n1 > 0 && n2 > 0) OGS_CUDA(cudaMemsetAsync(inter, 0, (size_t)n1 * n2 * 4, s));
    if (n1 + n2 == 0 || HW == 0) return 0;
    const int64_t W = packed_words(HW);
    const int64_t ngroups = W / 16;
    unsigned gx = (unsigned)((ngroups + MI_WARPS * 4 - 1) / (MI_WARPS * 4));     // ~4 groups per warp
    if (gx < 1) gx = 1;
    const dim3 pg(gx, (unsigned)(n1 + n2));
    const bool vec = (HW % 16 == 0) && (((uintptr_t)masks1 | (uintptr_t)masks2) % 16 == 0);
    if (vec) mask_pack_kernel<true><<<pg, MI_THREADS, 0, s>>>(n1, HW, W, masks1, masks2, scratch, counts);
    else mask_pack_kernel<false><<<pg, MI_THREADS, 0, s>>>(n1, HW, W, masks1, masks2, scratch, counts);
    if (n1 > 0 && n2 > 0) {
        const dim3 qg((unsigned)((W + MI_CHUNK - 1) / MI_CHUNK), (unsigned)((n1 + MI_WARPS - 1) / MI_WARPS));
        mask_pair_kernel<<<qg, MI_THREADS, 0, s>>>(n1, n2, W, scratch, inter);
    }
    return 0;
}

}  // namespace ogs
