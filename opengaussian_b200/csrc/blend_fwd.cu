// blend_fwd.cu -- per-tile front-to-back alpha compositing of C = 3 + n_extra channels plus depth
// and alpha in ONE pass.
//
// Replaces upstream renderCUDA<3> forward (ashawkey variant with depth/alpha outputs) that
// gaussian_renderer/__init__.py:104-112 calls, and -- through the extra channels -- the three
// additional passes of gaussian_renderer/__init__.py:129-163 (ins_feat[:, :3], ins_feat[:, 3:6],
// silhouette).  Semantics follow SURVEY.md section 8c / oracle/raster_oracle.c:
// power > 0 skip; alpha = min(0.99, o * exp(power)); alpha < 1/255 skip; T' = T (1 - alpha);
// T' < 1e-4 stops BEFORE applying; colour = sum c alpha T + T_final bg; depth = sum z alpha T;
// alpha_out = 1 - T_final.
//
// Layout: one CTA (256 threads) per 16x16 tile; warp w covers an 8x4 pixel block so that the
// warp-level "nobody contributes" vote is spatially tight.  Gaussians are staged 256 per round in
// shared memory as two float4 records + C colours (gathered once per tile, broadcast-read by all
// threads).  Bound: FP32 ALU + MUFU.EX2, not HBM.
#include "common.cuh"

namespace ogs {

#define BATCH 256

template <int C>
__global__ void __launch_bounds__(256) blend_fwd_kernel(BlendFwdArgs a) {
    __shared__ float4 s_r0[BATCH];
    __shared__ float4 s_r1[BATCH];
    __shared__ float s_col[BATCH * C];

    const int gx = (a.W + 15) / 16;
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int py = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;

    const uint2 range = a.ranges[tile];
    int todo = (int)(range.y - range.x);
    const int rounds = (todo + BATCH - 1) / BATCH;

    bool done = !inside;
    float T = 1.0f, D = 0.f;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0.f;
    uint32_t last = 0;
    const float bx0 = (float)(blockIdx.x * 16 + (warp & 1) * 8), by0 = (float)(blockIdx.y * 16 + (warp >> 1) * 4);

    for (int r = 0; r < rounds; r++, todo -= BATCH) {
        if (__syncthreads_count(done) == OGS_BLOCK) break;
        const int idx = r * BATCH + threadIdx.x;
        if (range.x + idx < range.y) {
            const uint32_t g = a.point_list[range.x + idx];
            s_r0[threadIdx.x] = __ldg(a.rec0 + g);
            s_r1[threadIdx.x] = __ldg(a.rec1 + g);
#pragma unroll
            for (int c = 0; c < 3; c++) s_col[threadIdx.x * C + c] = __ldg(a.base + 3 * (size_t)g + c);
#pragma unroll
            for (int c = 3; c < C; c++) s_col[threadIdx.x * C + c] = __ldg(a.extra + (size_t)(C - 3) * g + (c - 3));
        }
        __syncthreads();
        const int n = todo < BATCH ? todo : BATCH;
        if (!__all_sync(0xffffffffu, done)) {
            for (int grp = 0; grp < n; grp += 32) {
                // warp-level culling: lane l tests entry grp+l against this warp's 8x4 pixel block
                const int idx = grp + lane;
                bool hit = false;
                if (idx < n) hit = ogs_rect_hit(s_r0[idx], s_r1[idx], bx0, by0, bx0 + 7.0f, by0 + 3.0f);
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int j = grp + __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float4 r0 = s_r0[j];
                    const float4 r1 = s_r1[j];
                    const float dx = r0.x - pxf, dy = r0.y - pyf;
                    const float power = -0.5f * (r0.z * dx * dx + r1.x * dy * dy) - r0.w * dx * dy;
                    const float alpha = fminf(0.99f, r1.y * __expf(power));
                    if (done || power > 0.0f || alpha < (1.0f / 255.0f)) continue;
                    const float test_T = T * (1.0f - alpha);
                    if (test_T < 0.0001f) { done = true; continue; }
                    const float w = alpha * T;
#pragma unroll
                    for (int c = 0; c < C; c++) acc[c] = fmaf(s_col[j * C + c], w, acc[c]);
                    D = fmaf(r1.z, w, D);
                    T = test_T;
                    last = (uint32_t)(r * BATCH + j + 1);
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * a.W + px;
        const size_t HW = (size_t)a.H * a.W;
        a.final_T[pix] = T;
        a.n_contrib[pix] = last;
#pragma unroll
        for (int c = 0; c < C; c++) a.out_color[c * HW + pix] = acc[c] + T * __ldg(a.bg + c);
        a.out_depth[pix] = D;
        a.out_alpha[pix] = 1.0f - T;
    }
}

template <int C>
static int launch_c(const BlendFwdArgs& a, cudaStream_t s) {
    dim3 grid((a.W + 15) / 16, (a.H + 15) / 16);
    blend_fwd_kernel<C><<<grid, 256, 0, s>>>(a);
    return 0;
}

int launch_blend_forward(const BlendFwdArgs& a, cudaStream_t s) {
    switch (a.C) {
        case 3: return launch_c<3>(a, s);
        case 4: return launch_c<4>(a, s);
        case 6: return launch_c<6>(a, s);
        case 9: return launch_c<9>(a, s);
        case 12: return launch_c<12>(a, s);
        case 16: return launch_c<16>(a, s);
    }
    set_error("blend forward: unsupported channel count %d (supported: 3,4,6,9,12,16)", a.C);
    return -4;
}

}  // namespace ogs
