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
// Layout: one CTA of 128 threads per 16x16 tile (the tile size is part of the bit-exact binning
// contract).  Warp w owns an 8x8 pixel block, every lane TWO pixels (rows y and y+4): the shared
// memory reads of a Gaussian, the warp-level culling test and the loop overhead are amortised over
// 64 pixels and each thread carries two independent dependency chains.  Gaussians are staged 256
// per round in shared memory (two float4 records + C colours, gathered once per tile and
// broadcast-read).  Per group of 32 staged Gaussians, lane l tests Gaussian l against the warp's
// block (ogs_rect_hit) and only the ballot survivors are blended.
// Bound: FP32 ALU + MUFU.EX2 issue, not HBM.
#include "common.cuh"

namespace ogs {

#define BATCH 256
#define FWD_THREADS 128

template <int C>
__global__ void __launch_bounds__(FWD_THREADS) blend_fwd_kernel(BlendFwdArgs a) {
    __shared__ float4 s_r0[BATCH];
    __shared__ float4 s_r1[BATCH];
    __shared__ float s_col[BATCH * C];

    const int gx = (a.W + 15) / 16;
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bxi = blockIdx.x * 16 + (warp & 1) * 8, byi = blockIdx.y * 16 + (warp >> 1) * 8;
    const int px = bxi + (lane & 7);
    const int py0 = byi + (lane >> 3), py1 = py0 + 4;
    const bool in0 = px < a.W && py0 < a.H, in1 = px < a.W && py1 < a.H;
    const float pxf = (float)px, pyf0 = (float)py0, pyf1 = (float)py1;
    const float bx0 = (float)bxi, by0 = (float)byi;

    const uint2 range = a.ranges[tile];
    int todo = (int)(range.y - range.x);
    const int rounds = (todo + BATCH - 1) / BATCH;

    bool done0 = !in0, done1 = !in1;
    float T0 = 1.0f, T1 = 1.0f, D0 = 0.f, D1 = 0.f;
    float acc0[C], acc1[C];
#pragma unroll
    for (int c = 0; c < C; c++) { acc0[c] = 0.f; acc1[c] = 0.f; }
    uint32_t last0 = 0, last1 = 0;

    for (int r = 0; r < rounds; r++, todo -= BATCH) {
        if (__syncthreads_count(done0 && done1) == FWD_THREADS) break;
#pragma unroll
        for (int h = 0; h < BATCH / FWD_THREADS; h++) {
            const int slot = threadIdx.x + h * FWD_THREADS;
            const int idx = r * BATCH + slot;
            if (range.x + idx < range.y) {
                const uint32_t g = a.point_list[range.x + idx];
                s_r0[slot] = __ldg(a.rec0 + g);
                s_r1[slot] = __ldg(a.rec1 + g);
#pragma unroll
                for (int c = 0; c < 3; c++) s_col[slot * C + c] = __ldg(a.base + 3 * (size_t)g + c);
#pragma unroll
                for (int c = 3; c < C; c++) s_col[slot * C + c] = __ldg(a.extra + (size_t)(C - 3) * g + (c - 3));
            }
        }
        __syncthreads();
        const int n = todo < BATCH ? todo : BATCH;
        if (!__all_sync(0xffffffffu, done0 && done1)) {
            for (int grp = 0; grp < n; grp += 32) {
                const int idx = grp + lane;
                bool hit = false;
                if (idx < n) hit = ogs_rect_hit(s_r0[idx], s_r1[idx], bx0, by0, bx0 + 7.0f, by0 + 7.0f);
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int j = grp + __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float4 r0 = s_r0[j];
                    const float4 r1 = s_r1[j];
                    const float dx = r0.x - pxf;
                    const float dya = r0.y - pyf0, dyb = r0.y - pyf1;
                    const float adx = __fmul_rn(__fmul_rn(r0.z, dx), dx), bdx = __fmul_rn(r0.w, dx);
                    const float pw0 = ogs_power(adx, bdx, r1.x, dya);
                    const float pw1 = ogs_power(adx, bdx, r1.x, dyb);
                    const float al0 = fminf(0.99f, r1.y * __expf(pw0));
                    const float al1 = fminf(0.99f, r1.y * __expf(pw1));
                    const bool ok0 = !done0 && pw0 <= 0.0f && al0 >= (1.0f / 255.0f);
                    const bool ok1 = !done1 && pw1 <= 0.0f && al1 >= (1.0f / 255.0f);
                    const uint32_t pos = (uint32_t)(r * BATCH + j + 1);
                    if (ok0) {
                        const float test_T = T0 * (1.0f - al0);
                        if (test_T < 0.0001f) done0 = true;
                        else {
                            const float w = al0 * T0;
#pragma unroll
                            for (int c = 0; c < C; c++) acc0[c] = fmaf(s_col[j * C + c], w, acc0[c]);
                            D0 = fmaf(r1.z, w, D0);
                            T0 = test_T;
                            last0 = pos;
                        }
                    }
                    if (ok1) {
                        const float test_T = T1 * (1.0f - al1);
                        if (test_T < 0.0001f) done1 = true;
                        else {
                            const float w = al1 * T1;
#pragma unroll
                            for (int c = 0; c < C; c++) acc1[c] = fmaf(s_col[j * C + c], w, acc1[c]);
                            D1 = fmaf(r1.z, w, D1);
                            T1 = test_T;
                            last1 = pos;
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done0 && done1)) break;
            }
        }
    }
    const size_t HW = (size_t)a.H * a.W;
    if (in0) {
        const size_t pix = (size_t)py0 * a.W + px;
        a.final_T[pix] = T0;
        a.n_contrib[pix] = last0;
#pragma unroll
        for (int c = 0; c < C; c++) a.out_color[c * HW + pix] = acc0[c] + T0 * __ldg(a.bg + c);
        a.out_depth[pix] = D0;
        a.out_alpha[pix] = 1.0f - T0;
    }
    if (in1) {
        const size_t pix = (size_t)py1 * a.W + px;
        a.final_T[pix] = T1;
        a.n_contrib[pix] = last1;
#pragma unroll
        for (int c = 0; c < C; c++) a.out_color[c * HW + pix] = acc1[c] + T1 * __ldg(a.bg + c);
        a.out_depth[pix] = D1;
        a.out_alpha[pix] = 1.0f - T1;
    }
}

template <int C>
static int launch_c(const BlendFwdArgs& a, cudaStream_t s) {
    dim3 grid((a.W + 15) / 16, (a.H + 15) / 16);
    blend_fwd_kernel<C><<<grid, FWD_THREADS, 0, s>>>(a);
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
