// blend_bwd.cu -- per-tile back-to-front gradient pass of the alpha compositing.
//
// Replaces upstream renderCUDA<3> backward (ashawkey variant: colour + depth + alpha gradients) of
// the external rasterizer reached from loss.backward() at train.py:497-498.  Math follows
// SURVEY.md section 8c / oracle/raster_oracle.c (ogs_oracle_blend_backward).
//
// B200 design points (vs upstream's ~10 global float atomics per contributing (pixel, Gaussian)):
//  * the "colour behind" recursion is carried as ONE scalar per pixel: with d_i = <c_i, g> (dot of
//    the Gaussian's channel vector incl. depth and 1 with the pixel's incoming gradients),
//    S <- a_prev d_prev + (1 - a_prev) S and dL/dalpha = (d_i - S) T -- C-independent state;
//  * 128 threads per 16x16 tile, warp = 8x8 pixel block, TWO pixels per lane: the two pixels'
//    contributions are added in registers first, so the cross-lane reduction, the shared-memory
//    reads and the culling test are amortised over 64 pixels;
//  * per (warp, Gaussian) the lanes' contributions are summed with a recursive-halving
//    (transpose) reduction: V values cost ~V + 5 shuffles instead of 5 V;
//  * warp-level culling (ogs_rect_hit): lane l tests Gaussian l of a 32-group against the warp's
//    block; only ballot survivors are evaluated;
//  * each warp accumulates into its own shared-memory rows (no shared atomics); once per batch of
//    64 Gaussians the 4 warp rows are summed and ONE red.global.add per value per (tile, Gaussian)
//    is issued, skipping zeros;
//  * GEOM=false specialisation (OpenGaussian stages 1-2 detach geometry, train.py:431-436): only
//    w = alpha T and dL/dc are evaluated.
#include "common.cuh"

namespace ogs {

#define BB 64            // Gaussians per backward batch
#define BWD_THREADS 128
#define BWD_WARPS 4

template <int CUR, int M>
struct HalvingReduce {
    template <int VP>
    __device__ __forceinline__ static void run(float (&v)[VP], int lane) {
        if constexpr (M >= 1) {
            if constexpr (CUR > 1) {
                constexpr int HALF = CUR / 2;
                const bool upper = (lane & M) != 0;
#pragma unroll
                for (int k = 0; k < HALF; k++) {
                    const float send = upper ? v[k] : v[k + HALF];
                    const float keep = upper ? v[k + HALF] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, M);
                }
                HalvingReduce<HALF, M / 2>::run(v, lane);
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], M);
                HalvingReduce<1, M / 2>::run(v, lane);
            }
        }
    }
};

constexpr int next_pow2(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : (v <= 4 ? 4 : (v <= 8 ? 8 : (v <= 16 ? 16 : 32)))); }
constexpr int log2c(int v) { return v <= 1 ? 0 : 1 + log2c(v / 2); }

int blend_bwd_stride(int C, int geom) { return geom ? C + 7 : C; }

template <int C>
struct PixelState {
    float T, T_final, S, last_dot, last_alpha, gd, ga, bg_dot, pyf;
    float g[C];
    int last;
};

template <int C, bool GEOM, int VP>
__device__ __forceinline__ void pixel_contrib(PixelState<C>& p, bool ok, float alpha, float G, float dx, float dy,
                                              const float4& r0, const float4& r1, const float* __restrict__ col,
                                              float half_w, float half_h, float (&v)[VP]) {
    if (!ok) return;
    const float inv = __fdividef(1.0f, 1.0f - alpha);
    p.T = p.T * inv;
    const float w = alpha * p.T;
#pragma unroll
    for (int c = 0; c < C; c++) v[c] = fmaf(w, p.g[c], v[c]);
    if (GEOM) {
        float dot = fmaf(r1.z, p.gd, p.ga);
#pragma unroll
        for (int c = 0; c < C; c++) dot = fmaf(col[c], p.g[c], dot);
        p.S = fmaf(p.last_alpha, p.last_dot, (1.0f - p.last_alpha) * p.S);
        p.last_dot = dot;
        p.last_alpha = alpha;
        float dL_dalpha = (dot - p.S) * p.T;
        dL_dalpha = fmaf(-p.T_final * inv, p.bg_dot, dL_dalpha);
        const float dL_dG = r1.y * dL_dalpha;
        const float gdx = G * dx, gdy = G * dy;
        const float dG_ddelx = -gdx * r0.z - gdy * r0.w;
        const float dG_ddely = -gdy * r1.x - gdx * r0.w;
        v[C + 0] = fmaf(w, p.gd, v[C + 0]);
        v[C + 1] = fmaf(dL_dG * dG_ddelx, half_w, v[C + 1]);
        v[C + 2] = fmaf(dL_dG * dG_ddely, half_h, v[C + 2]);
        const float hq = -0.5f * dL_dG;
        v[C + 3] = fmaf(hq * gdx, dx, v[C + 3]);
        v[C + 4] = fmaf(hq * gdx, dy, v[C + 4]);
        v[C + 5] = fmaf(hq * gdy, dy, v[C + 5]);
        v[C + 6] = fmaf(G, dL_dalpha, v[C + 6]);
    }
}

template <int C, bool GEOM>
__global__ void __launch_bounds__(BWD_THREADS) blend_bwd_kernel(BlendBwdArgs a) {
    constexpr int V = GEOM ? C + 7 : C;
    constexpr int VP = next_pow2(V);
    constexpr int SHIFT = 5 - log2c(VP);  // lane >> SHIFT = value index held after the reduction
    extern __shared__ float s_dyn[];      // [BWD_WARPS][BB][V] per-warp accumulators
    __shared__ float4 s_r0[BB];
    __shared__ float4 s_r1[BB];
    __shared__ float s_col[BB * C];
    __shared__ uint32_t s_id[BB];
    __shared__ int s_max_last;

    const int gx = (a.W + 15) / 16;
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bxi = blockIdx.x * 16 + (warp & 1) * 8, byi = blockIdx.y * 16 + (warp >> 1) * 8;
    const int px = bxi + (lane & 7);
    const float pxf = (float)px;
    const float bx0 = (float)bxi, by0 = (float)byi;
    const size_t HW = (size_t)a.H * a.W;
    const uint2 range = a.ranges[tile];

    PixelState<C> ps[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int py = byi + (lane >> 3) + 4 * k;
        const bool inside = px < a.W && py < a.H;
        const size_t pix = (size_t)py * a.W + px;
        PixelState<C>& p = ps[k];
        p.pyf = (float)py;
        p.last = inside ? (int)a.n_contrib[pix] : 0;
        p.T_final = inside ? a.final_T[pix] : 0.f;
        p.T = p.T_final;
        p.S = 0.f; p.last_dot = 0.f; p.last_alpha = 0.f; p.gd = 0.f; p.ga = 0.f; p.bg_dot = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
            p.g[c] = inside ? __ldg(a.dL_dcolor + c * HW + pix) : 0.f;
            if (GEOM) p.bg_dot = fmaf(__ldg(a.bg + c), p.g[c], p.bg_dot);
        }
        if (GEOM) {
            p.gd = (inside && a.dL_ddepth) ? __ldg(a.dL_ddepth + pix) : 0.f;
            p.ga = (inside && a.dL_dalpha) ? __ldg(a.dL_dalpha + pix) : 0.f;
        }
    }
    const float half_w = 0.5f * (float)a.W, half_h = 0.5f * (float)a.H;

    // block / warp maxima of the last contributor
    if (threadIdx.x == 0) s_max_last = 0;
    for (int e = threadIdx.x; e < BWD_WARPS * BB * V; e += BWD_THREADS) s_dyn[e] = 0.f;
    __syncthreads();
    int wmax = max(ps[0].last, ps[1].last);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, m));
    if (lane == 0 && wmax > 0) atomicMax(&s_max_last, wmax);
    __syncthreads();
    const int max_last = s_max_last;
    if (max_last == 0) return;
    float* my_acc = s_dyn + (size_t)warp * BB * V;

    for (int b = (max_last - 1) / BB; b >= 0; b--) {
        const int start = b * BB;
        const int n = min(BB, max_last - start);
        // ---- stage the batch ----
        if (threadIdx.x < n) {
            const uint32_t gid = a.point_list[range.x + start + threadIdx.x];
            s_id[threadIdx.x] = gid;
            s_r0[threadIdx.x] = __ldg(a.rec0 + gid);
            s_r1[threadIdx.x] = __ldg(a.rec1 + gid);
        }
        for (int e = threadIdx.x; e < n * C; e += BWD_THREADS) {
            const int j = e / C, c = e - j * C;
            const uint32_t gid = a.point_list[range.x + start + j];
            s_col[e] = (c < 3) ? __ldg(a.base + 3 * (size_t)gid + c) : __ldg(a.extra + (size_t)(C - 3) * gid + (c - 3));
        }
        __syncthreads();
        // ---- traverse back to front ----
        if (start < wmax) {
            const int nw = min(n, wmax - start);
            for (int grp = ((nw - 1) >> 5) << 5; grp >= 0; grp -= 32) {
                const int idx = grp + lane;
                bool hit = false;
                if (idx < nw) hit = ogs_rect_hit(s_r0[idx], s_r1[idx], bx0, by0, bx0 + 7.0f, by0 + 7.0f);
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int bit = 31 - __clz(mask);
                    mask &= ~(1u << bit);
                    const int j = grp + bit;
                    const int pos = start + j;
                    const float4 r0 = s_r0[j];
                    const float4 r1 = s_r1[j];
                    const float dx = r0.x - pxf;
                    const float dy0 = r0.y - ps[0].pyf, dy1 = r0.y - ps[1].pyf;
                    const float adx = __fmul_rn(__fmul_rn(r0.z, dx), dx), bdx = __fmul_rn(r0.w, dx);
                    const float pw0 = ogs_power(adx, bdx, r1.x, dy0);
                    const float pw1 = ogs_power(adx, bdx, r1.x, dy1);
                    const float G0 = __expf(pw0), G1 = __expf(pw1);
                    const float al0 = fminf(0.99f, r1.y * G0), al1 = fminf(0.99f, r1.y * G1);
                    const bool ok0 = pos < ps[0].last && pw0 <= 0.0f && al0 >= (1.0f / 255.0f);
                    const bool ok1 = pos < ps[1].last && pw1 <= 0.0f && al1 >= (1.0f / 255.0f);
                    if (!__any_sync(0xffffffffu, ok0 || ok1)) continue;
                    float v[VP];
#pragma unroll
                    for (int k = 0; k < VP; k++) v[k] = 0.f;
                    pixel_contrib<C, GEOM, VP>(ps[0], ok0, al0, G0, dx, dy0, r0, r1, s_col + j * C, half_w, half_h, v);
                    pixel_contrib<C, GEOM, VP>(ps[1], ok1, al1, G1, dx, dy1, r0, r1, s_col + j * C, half_w, half_h, v);
                    HalvingReduce<VP, 16>::run(v, lane);
                    const int k = lane >> SHIFT;
                    if ((lane & ((1 << SHIFT) - 1)) == 0 && k < V) my_acc[j * V + k] += v[0];
                }
            }
        }
        __syncthreads();
        // ---- flush: sum the warp rows, one red per value per (tile, Gaussian) ----
        for (int e = threadIdx.x; e < n * V; e += BWD_THREADS) {
            float sum = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < BWD_WARPS; w8++) {
                sum += s_dyn[(size_t)w8 * BB * V + e];
                s_dyn[(size_t)w8 * BB * V + e] = 0.f;
            }
            if (sum != 0.f) {
                const int j = e / V, k = e - j * V;
                atomicAdd(a.acc + (size_t)s_id[j] * a.stride + k, sum);
            }
        }
        __syncthreads();
    }
}

template <int C, bool GEOM>
static int launch_cg(const BlendBwdArgs& a, cudaStream_t s) {
    constexpr int V = GEOM ? C + 7 : C;
    const size_t smem = (size_t)BWD_WARPS * BB * V * sizeof(float);
    dim3 grid((a.W + 15) / 16, (a.H + 15) / 16);
    blend_bwd_kernel<C, GEOM><<<grid, BWD_THREADS, smem, s>>>(a);
    return 0;
}

template <int C>
static int launch_c(const BlendBwdArgs& a, cudaStream_t s) {
    return a.geom ? launch_cg<C, true>(a, s) : launch_cg<C, false>(a, s);
}

int launch_blend_backward(const BlendBwdArgs& a, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(a.acc, 0, (size_t)a.P * a.stride * sizeof(float), s);
    if (e != cudaSuccess) return cuda_fail(e, "memset acc");
    switch (a.C) {
        case 3: return launch_c<3>(a, s);
        case 4: return launch_c<4>(a, s);
        case 6: return launch_c<6>(a, s);
        case 9: return launch_c<9>(a, s);
        case 12: return launch_c<12>(a, s);
        case 16: return launch_c<16>(a, s);
    }
    set_error("blend backward: unsupported channel count %d", a.C);
    return -4;
}

}  // namespace ogs
