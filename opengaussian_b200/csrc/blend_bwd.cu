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
//  * per (warp, Gaussian) the 32 pixels' contributions are summed with a recursive-halving
//    (transpose) reduction: V values cost ~V + 5 shuffles instead of 5 V;
//  * each warp accumulates into its own shared-memory rows (no shared atomics); once per batch of
//    64 Gaussians the 8 warp rows are summed and ONE red.global.add per value per (tile, Gaussian)
//    is issued, skipping zeros;
//  * GEOM=false specialisation (OpenGaussian stages 1-2 detach geometry, train.py:431-436): only
//    w = alpha T and dL/dc are evaluated.
#include "common.cuh"

namespace ogs {

#define BB 64   // Gaussians per backward batch

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

template <int C, bool GEOM>
__global__ void __launch_bounds__(256) blend_bwd_kernel(BlendBwdArgs a) {
    constexpr int V = GEOM ? C + 7 : C;
    constexpr int VP = next_pow2(V);
    constexpr int SHIFT = 5 - log2c(VP);  // lane >> SHIFT = value index held after the reduction
    extern __shared__ float s_dyn[];      // [8][BB][V] per-warp accumulators
    __shared__ float4 s_r0[BB];
    __shared__ float4 s_r1[BB];
    __shared__ float s_col[BB * C];
    __shared__ uint32_t s_id[BB];
    __shared__ int s_max_last;

    const int gx = (a.W + 15) / 16;
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int py = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix = (size_t)py * a.W + px;
    const uint2 range = a.ranges[tile];

    const int last = inside ? (int)a.n_contrib[pix] : 0;
    const float T_final = inside ? a.final_T[pix] : 0.f;
    float T = T_final;
    float g[C];
    float gd = 0.f, ga = 0.f, bg_dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) {
        g[c] = inside ? __ldg(a.dL_dcolor + c * HW + pix) : 0.f;
        if (GEOM) bg_dot = fmaf(__ldg(a.bg + c), g[c], bg_dot);
    }
    if (GEOM) {
        gd = (inside && a.dL_ddepth) ? __ldg(a.dL_ddepth + pix) : 0.f;
        ga = (inside && a.dL_dalpha) ? __ldg(a.dL_dalpha + pix) : 0.f;
    }
    float S = 0.f, last_dot = 0.f, last_alpha = 0.f;
    const float half_w = 0.5f * (float)a.W, half_h = 0.5f * (float)a.H;
    const float bx0 = (float)(blockIdx.x * 16 + (warp & 1) * 8), by0 = (float)(blockIdx.y * 16 + (warp >> 1) * 4);

    // block / warp maxima of the last contributor
    if (threadIdx.x == 0) s_max_last = 0;
    for (int e = threadIdx.x; e < 8 * BB * V; e += 256) s_dyn[e] = 0.f;
    __syncthreads();
    int wmax = last;
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
        for (int e = threadIdx.x; e < n * C; e += 256) {
            const int j = e / C, c = e - j * C;
            const uint32_t gid = a.point_list[range.x + start + j];
            s_col[e] = (c < 3) ? __ldg(a.base + 3 * (size_t)gid + c) : __ldg(a.extra + (size_t)(C - 3) * gid + (c - 3));
        }
        __syncthreads();
        // ---- traverse back to front ----
        if (start < wmax) {
          const int nw = min(n, wmax - start);
          for (int grp = ((nw - 1) >> 5) << 5; grp >= 0; grp -= 32) {
            // warp-level culling (see blend_fwd.cu): lane l tests entry grp+l against the warp's 8x4 block
            const int idx = grp + lane;
            bool hit = false;
            if (idx < nw) hit = ogs_rect_hit(s_r0[idx], s_r1[idx], bx0, by0, bx0 + 7.0f, by0 + 3.0f);
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int bit = 31 - __clz(mask);
                mask &= ~(1u << bit);
                const int j = grp + bit;
                const int pos = start + j;
                const float4 r0 = s_r0[j];
                const float4 r1 = s_r1[j];
                const float dx = r0.x - pxf, dy = r0.y - pyf;
                const float power = -0.5f * (r0.z * dx * dx + r1.x * dy * dy) - r0.w * dx * dy;
                const float G = __expf(power);
                const float alpha = fminf(0.99f, r1.y * G);
                const bool ok = pos < last && power <= 0.0f && alpha >= (1.0f / 255.0f);
                if (!__any_sync(0xffffffffu, ok)) continue;
                float v[VP];
#pragma unroll
                for (int k = 0; k < VP; k++) v[k] = 0.f;
                if (ok) {
                    const float inv = __fdividef(1.0f, 1.0f - alpha);
                    T = T * inv;
                    const float w = alpha * T;
#pragma unroll
                    for (int c = 0; c < C; c++) v[c] = w * g[c];
                    if (GEOM) {
                        float dot = fmaf(r1.z, gd, ga);
#pragma unroll
                        for (int c = 0; c < C; c++) dot = fmaf(s_col[j * C + c], g[c], dot);
                        S = fmaf(last_alpha, last_dot, (1.0f - last_alpha) * S);
                        last_dot = dot;
                        last_alpha = alpha;
                        float dL_dalpha = (dot - S) * T;
                        dL_dalpha = fmaf(-T_final * inv, bg_dot, dL_dalpha);
                        const float dL_dG = r1.y * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * r0.z - gdy * r0.w;
                        const float dG_ddely = -gdy * r1.x - gdx * r0.w;
                        v[C + 0] = w * gd;
                        v[C + 1] = dL_dG * dG_ddelx * half_w;
                        v[C + 2] = dL_dG * dG_ddely * half_h;
                        v[C + 3] = -0.5f * gdx * dx * dL_dG;
                        v[C + 4] = -0.5f * gdx * dy * dL_dG;
                        v[C + 5] = -0.5f * gdy * dy * dL_dG;
                        v[C + 6] = G * dL_dalpha;
                    }
                }
                HalvingReduce<VP, 16>::run(v, lane);
                const int k = lane >> SHIFT;
                if ((lane & ((1 << SHIFT) - 1)) == 0 && k < V) my_acc[j * V + k] += v[0];
            }
          }
        }
        __syncthreads();
        // ---- flush: sum the 8 warp rows, one red per value per (tile, Gaussian) ----
        for (int e = threadIdx.x; e < n * V; e += 256) {
            float sum = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; w8++) {
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
    const size_t smem = (size_t)8 * BB * V * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(blend_bwd_kernel<C, GEOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_done = true;
    }
    dim3 grid((a.W + 15) / 16, (a.H + 15) / 16);
    blend_bwd_kernel<C, GEOM><<<grid, 256, smem, s>>>(a);
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
