// footprint.cu -- batched single-splat footprints and their weighted SAM-id votes.
//
// Replaces the per-Gaussian loop of utils/sam_refinement_utils.py::MultiViewSAMMaskRefiner:
//   get_splat_id_and_weights (:902-913) = render_single_gaussian (:330-403, a P = 1 rasterizer call with white,
//   view-independent SH) -> fix_image (:143-176, uint8 quantisation) -> rgb_to_weight_map (:103-141) ->
//   get_most_common_id_in_mask_weighted (:645-702, torch.bincount with weights + argmax).
// For ONE camera and B selected Gaussians this is B rasterizer launches sequences, B full-image quantisations and
// B full-image bincounts in the reference.  Here: one preprocess launch over the B Gaussians (the SAME kernel the
// rasterizer uses, so radii / tile rectangles / conics are identical), then one warp per splat walks the pixels
// of the splat's tile rectangle, evaluates alpha with the blend kernels' pinned arithmetic (ogs_stage /
// ogs_pair_power / ogs_ex2: bit-identical contribute decisions), quantises colour * alpha * 255 to the uint8 the
// reference sees, and votes the quantised weights into a small per-warp hash table keyed by SAM id.
//   single Gaussian, black background:  pixel = colour * alpha  (T = 1, so the transmittance floor never triggers)
//   weight = uint8(pixel * 255) / 255 / max  -- the vote compares integer sums of q = uint8(...), which orders
//   ids exactly as the normalised float weights do (ties -> lowest id, like argmax over the shifted ids).
// Outputs per splat: dominant id (empty_id when the footprint is empty, like argmax over all-zero counts), the
// integer weight of that id (-1 if the footprint touched more than 128 distinct ids), the number of non-black pixels, the largest q (the normaliser), the radius.
#include "common.cuh"

namespace ogs {

static_assert(sizeof(ogs_footprint_inputs) == 96, "ogs_footprint_inputs layout is part of the C ABI (ctypes mirror in _lib.py)");

#define FP_WARPS 4
#define FP_SLOTS 128
#define FP_EMPTY ((int)0x80000000)

// same arithmetic as tile_rect in preprocess.cu / binning.cu (no multiply-add, so contraction cannot differ)
__device__ __forceinline__ void fp_tile_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1, int& y1) {
    const float r = (float)radius;
    x0 = min(gx, max(0, (int)((px - r) / 16.0f)));
    y0 = min(gy, max(0, (int)((py - r) / 16.0f)));
    x1 = min(gx, max(0, (int)((px + r + 15.0f) / 16.0f)));
    y1 = min(gy, max(0, (int)((py + r + 15.0f) / 16.0f)));
}

__global__ void __launch_bounds__(FP_WARPS * 32) footprint_vote_kernel(int P, int W, int H, const float4* __restrict__ rec0,
                                                                       const float4* __restrict__ rec1,
                                                                       const uint32_t* __restrict__ tiles,
                                                                       const int32_t* __restrict__ sam_ids, int empty_id, float color,
                                                                       int32_t* __restrict__ dominant_id,
                                                                       int32_t* __restrict__ dominant_weight,
                                                                       int32_t* __restrict__ footprint_pixels,
                                                                       int32_t* __restrict__ q_max_out, int32_t* __restrict__ overflow) {
    __shared__ int s_key[FP_WARPS][FP_SLOTS];
    __shared__ int s_val[FP_WARPS][FP_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * FP_WARPS + warp;
    if (i >= P) return;                                     // warp-uniform, no block barrier below
    int* key = s_key[warp];
    int* val = s_val[warp];
    for (int s = lane; s < FP_SLOTS; s += 32) { key[s] = FP_EMPTY; val[s] = 0; }
    __syncwarp();

    int npix = 0, qmax = 0, over = 0;
    if (tiles[i] != 0u) {
        const float4 r0 = rec0[i], r1 = rec1[i];
        float4 sa;
        float2 sb;
        ogs_stage(r0, r1, sa, sb);
        const int gx = (W + 15) / 16, gy = (H + 15) / 16;
        int x0, y0, x1, y1;
        fp_tile_rect(r0.x, r0.y, __float_as_int(r1.w), gx, gy, x0, y0, x1, y1);
        const int px0 = x0 * 16, px1 = min(W, x1 * 16), py0 = y0 * 16, py1 = min(H, y1 * 16);
        for (int py = py0; py < py1; ++py) {
            const float2 npy = make_float2(-(float)py, -(float)py);
            for (int xb = px0; xb < px1; xb += 32) {
                const int px = xb + lane;
                int q = 0;
                if (px < px1) {
                    const float dx = sa.x - (float)px;
                    const float adx = __fmul_rn(__fmul_rn(sa.z, dx), dx), bdx = __fmul_rn(sa.w, dx);
                    float2 dy;
                    const float2 pw = ogs_pair_power(adx, bdx, sb.x, sa.y, npy, dy);
                    float al = fminf(0.99f, __fmul_rn(sb.y, ogs_ex2(pw.x)));
                    if (pw.x <= 0.0f && al >= (1.0f / 255.0f)) {
                        const float pix = __fmul_rn(color, al);                        // acc = fma(colour, alpha * T, 0), T = 1
                        q = (int)fminf(fmaxf(__fmul_rn(pix, 255.0f), 0.0f), 255.0f);   // clamp(x * 255, 0, 255).to(uint8)
                    }
                }
                const uint32_t live = __ballot_sync(0xFFFFFFFFu, q > 0);
                if (live == 0u) continue;
                npix += __popc(live);
                qmax = max(qmax, (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)q));
                if (q > 0) {
                    const int id = __ldg(sam_ids + (size_t)py * W + px);
                    const uint32_t peers = __match_any_sync(live, id);
                    const int sum = __reduce_add_sync(peers, q);
                    if ((int)(__ffs(peers) - 1) == lane) {
                        uint32_t h = ((uint32_t)id * 2654435761u) >> 25;                // 7 bits
                        int probes = 0;
                        for (; probes < FP_SLOTS; ++probes) {
                            const int prev = atomicCAS(key + h, FP_EMPTY, id);
                            if (prev == FP_EMPTY || prev == id) { atomicAdd(val + h, sum); break; }
                            h = (h + 1) & (FP_SLOTS - 1);
                        }
                        if (probes == FP_SLOTS) over = 1;
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    // argmax over the table: largest weight, ties -> lowest id
    int best_w = 0, best_id = empty_id;
    bool have = false;
    for (int s = lane; s < FP_SLOTS; s += 32) {
        const int k = key[s];
        if (k != FP_EMPTY) {
            const int w = val[s];
            if (!have || w > best_w || (w == best_w && k < best_id)) { best_w = w; best_id = k; have = true; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ow = __shfl_xor_sync(0xFFFFFFFFu, best_w, o);
        const int oi = __shfl_xor_sync(0xFFFFFFFFu, best_id, o);
        const int oh = __shfl_xor_sync(0xFFFFFFFFu, (int)have, o);
        if (oh && (!have || ow > best_w || (ow == best_w && oi < best_id))) { best_w = ow; best_id = oi; have = true; }
    }
    over = __any_sync(0xFFFFFFFFu, over);
    if (lane == 0) {
        dominant_id[i] = have ? best_id : empty_id;
        dominant_weight[i] = over ? -1 : (have ? best_w : 0);          // -1: more than FP_SLOTS distinct ids, vote unreliable
        footprint_pixels[i] = npix;
        q_max_out[i] = qmax;
        if (over) atomicAdd(overflow, 1);
    }
}

int launch_footprint_votes(int P, int W, int H, const GeomPtrs& g, const int32_t* sam_ids, int empty_id, float color,
                           int32_t* dominant_id, int32_t* dominant_weight, int32_t* footprint_pixels, int32_t* q_max,
                           int32_t* overflow, cudaStream_t s) {
    if (P <= 0) return 0;
    footprint_vote_kernel<<<(P + FP_WARPS - 1) / FP_WARPS, FP_WARPS * 32, 0, s>>>(P, W, H, g.rec0, g.rec1, g.tiles, sam_ids, empty_id, color,
                                                                                 dominant_id, dominant_weight, footprint_pixels,
                                                                                 q_max, overflow);
    return 0;
}

}  // namespace ogs
