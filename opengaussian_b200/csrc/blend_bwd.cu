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
//  * 128 threads per 16x16 tile, warp = 8x8 pixel block, TWO pixels per lane evaluated as a packed
//    pair (FADD2/FMUL2/FFMA2, same staged records and exponent arithmetic as blend_fwd.cu, so the
//    skip decisions are bit-identical); non-contributing pixels are masked arithmetically
//    (alpha = G = 0 leaves T, S and every sum unchanged) so both halves share one instruction stream;
//  * geometry gradients are reduced as the raw moments of u = G dL/dalpha (sum u, u dx, u dy,
//    u dx^2, u dx dy, u dy^2): dL/dmean2D and dL/dconic are linear in them with per-Gaussian
//    factors, which preprocess_bwd applies once per Gaussian instead of once per (pixel, Gaussian);
//  * per (warp, Gaussian) the lanes' sums are combined with a recursive-halving (transpose)
//    reduction over power-of-two chunks of the value vector: V values cost ~V + 5 shuffles;
//  * warp-level culling (ogs_rect_hit_s): lane l tests Gaussian l of a 32-group against the warp's
//    block; only ballot survivors are evaluated;
//  * each warp stores into its own shared-memory rows (a (warp, Gaussian) pair is visited once per
//    batch: plain stores, no read-modify-write); once per batch of 64 Gaussians the 4 warp rows are
//    summed and ONE red.global.add per value per (tile, Gaussian) is issued, skipping zeros;
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

constexpr int floor_pow2(int v) { return v >= 16 ? 16 : (v >= 8 ? 8 : (v >= 4 ? 4 : (v >= 2 ? 2 : 1))); }
constexpr int log2c(int v) { return v <= 1 ? 0 : 1 + log2c(v / 2); }

// Sums v[BASE..BASE+N) over the 32 lanes, chunk by chunk (largest power of two first), and stores
// the totals to row[BASE..BASE+N): after a chunk of S values, lane l with (l % (32/S)) == 0 holds
// the total of value l / (32/S).
template <int N, int BASE>
struct ChunkReduceStore {
    template <int VT>
    __device__ __forceinline__ static void run(const float (&v)[VT], int lane, float* row) {
        if constexpr (N > 0) {
            constexpr int S = floor_pow2(N);
            constexpr int SH = 5 - log2c(S);
            float t[S];
#pragma unroll
            for (int k = 0; k < S; k++) t[k] = v[BASE + k];
            HalvingReduce<S, 16>::run(t, lane);
            if ((lane & ((1 << SH) - 1)) == 0) row[BASE + (lane >> SH)] = t[0];
            ChunkReduceStore<N - S, BASE + S>::run(v, lane, row);
        }
    }
};

int blend_bwd_stride(int C, int geom) { return geom ? C + 7 : C; }

template <int C, bool GEOM>
__global__ void __launch_bounds__(BWD_THREADS) blend_bwd_kernel(BlendBwdArgs a) {
    constexpr int V = GEOM ? C + 7 : C;
    constexpr int CH = (C + 1 + 3) & ~3;  // colours + depth, padded to float4
    extern __shared__ float s_dyn[];      // [BWD_WARPS][BB][V] per-warp sums
    __shared__ float4 s_a[BB];
    __shared__ float2 s_b[BB];
    __shared__ __align__(16) float s_ch[BB * CH];
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

    // per-lane state of the pixel pair (.x: row y, .y: row y + 4)
    float2 T2, S2 = s2(0.f), om2 = s2(1.f), prod2 = s2(0.f), gd2 = s2(0.f), ga2 = s2(0.f), tfbg2 = s2(0.f), npy;
    float2 g2[C];
    int last[2];
    {
        float Tf[2], gdv[2] = {0.f, 0.f}, gav[2] = {0.f, 0.f}, bgd[2] = {0.f, 0.f};
        float gv[2][C];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int py = byi + (lane >> 3) + 4 * k;
            const bool inside = px < a.W && py < a.H;
            const size_t pix = (size_t)py * a.W + px;
            last[k] = inside ? (int)a.n_contrib[pix] : 0;
            Tf[k] = inside ? a.final_T[pix] : 0.f;
#pragma unroll
            for (int c = 0; c < C; c++) {
                gv[k][c] = inside ? __ldg(a.dL_dcolor + c * HW + pix) : 0.f;
                if (GEOM) bgd[k] = fmaf(__ldg(a.bg + c), gv[k][c], bgd[k]);
            }
            if (GEOM) {
                gdv[k] = (inside && a.dL_ddepth) ? __ldg(a.dL_ddepth + pix) : 0.f;
                gav[k] = (inside && a.dL_dalpha) ? __ldg(a.dL_dalpha + pix) : 0.f;
            }
        }
        npy = make_float2(-(float)(byi + (lane >> 3)), -(float)(byi + (lane >> 3) + 4));
        T2 = make_float2(Tf[0], Tf[1]);
#pragma unroll
        for (int c = 0; c < C; c++) g2[c] = make_float2(gv[0][c], gv[1][c]);
        if (GEOM) {
            gd2 = make_float2(gdv[0], gdv[1]);
            ga2 = make_float2(gav[0], gav[1]);
            tfbg2 = make_float2(Tf[0] * bgd[0], Tf[1] * bgd[1]);
        }
    }

    // block / warp maxima of the last contributor
    if (threadIdx.x == 0) s_max_last = 0;
    for (int e = threadIdx.x; e < BWD_WARPS * BB * V; e += BWD_THREADS) s_dyn[e] = 0.f;
    __syncthreads();
    int wmax = max(last[0], last[1]);
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
        // ---- stage the batch (pre-scaled records, colours + depth) ----
        if (threadIdx.x < n) {
            const uint32_t gid = a.point_list[range.x + start + threadIdx.x];
            s_id[threadIdx.x] = gid;
            const float4 r1 = __ldg(a.rec1 + gid);
            ogs_stage(__ldg(a.rec0 + gid), r1, s_a[threadIdx.x], s_b[threadIdx.x]);
            s_ch[threadIdx.x * CH + C] = r1.z;
        }
        for (int e = threadIdx.x; e < n * C; e += BWD_THREADS) {
            const int j = e / C, c = e - j * C;
            const uint32_t gid = a.point_list[range.x + start + j];
            s_ch[j * CH + c] = (c < 3) ? __ldg(a.base + 3 * (size_t)gid + c) : __ldg(a.extra + (size_t)(C - 3) * gid + (c - 3));
        }
        __syncthreads();
        // ---- traverse back to front ----
        if (start < wmax) {
            const int nw = min(n, wmax - start);
            for (int grp = ((nw - 1) >> 5) << 5; grp >= 0; grp -= 32) {
                const int idx = grp + lane;
                bool hit = false;
                if (idx < nw) hit = ogs_rect_hit_s(s_a[idx], s_b[idx], bx0, by0, bx0 + 7.0f, by0 + 7.0f);
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int bit = 31 - __clz(mask);
                    mask &= ~(1u << bit);
                    const int j = grp + bit;
                    const int pos = start + j;
                    const float4 ra = s_a[j];
                    const float2 rb = s_b[j];
                    const float dx = ra.x - pxf;
                    const float adx = __fmul_rn(__fmul_rn(ra.z, dx), dx), bdx = __fmul_rn(ra.w, dx);
                    float2 dy;
                    const float2 pw = ogs_pair_power(adx, bdx, rb.x, ra.y, npy, dy);
                    const float2 G = make_float2(ogs_ex2(pw.x), ogs_ex2(pw.y));
                    float2 al = __fmul2_rn(s2(rb.y), G);
                    al.x = fminf(0.99f, al.x);
                    al.y = fminf(0.99f, al.y);
                    const bool oka = pos < last[0] && pw.x <= 0.0f && al.x >= (1.0f / 255.0f);
                    const bool okb = pos < last[1] && pw.y <= 0.0f && al.y >= (1.0f / 255.0f);
                    if (!__any_sync(0xffffffffu, oka || okb)) continue;
                    float chv[CH];
#pragma unroll
                    for (int q = 0; q < CH / 4; q++) {
                        const float4 t = reinterpret_cast<const float4*>(s_ch + j * CH)[q];
                        chv[4 * q] = t.x; chv[4 * q + 1] = t.y; chv[4 * q + 2] = t.z; chv[4 * q + 3] = t.w;
                    }
                    const float2 alm = make_float2(oka ? al.x : 0.f, okb ? al.y : 0.f);
                    const float2 om = __fadd2_rn(s2(1.0f), make_float2(-alm.x, -alm.y));
                    const float2 inv = make_float2(ogs_rcp(om.x), ogs_rcp(om.y));
                    T2 = __fmul2_rn(T2, inv);
                    const float2 w = __fmul2_rn(alm, T2);
                    float v[V];
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const float2 t = __fmul2_rn(w, g2[c]);
                        v[c] = t.x + t.y;
                    }
                    if (GEOM) {
                        float2 dot = __ffma2_rn(s2(chv[0]), g2[0], ga2);
#pragma unroll
                        for (int c = 1; c < C; c++) dot = __ffma2_rn(s2(chv[c]), g2[c], dot);
                        dot = __ffma2_rn(s2(chv[C]), gd2, dot);
                        S2 = __ffma2_rn(om2, S2, prod2);          // S <- a_prev d_prev + (1 - a_prev) S
                        prod2 = __fmul2_rn(alm, dot);
                        om2 = om;
                        const float2 dS = __fadd2_rn(dot, make_float2(-S2.x, -S2.y));
                        const float2 dLda = __ffma2_rn(make_float2(-tfbg2.x, -tfbg2.y), inv, __fmul2_rn(dS, T2));
                        const float2 Gm = make_float2(oka ? G.x : 0.f, okb ? G.y : 0.f);
                        const float2 u = __fmul2_rn(Gm, dLda);
                        const float2 uy = __fmul2_rn(u, dy);
                        const float2 uyy = __fmul2_rn(uy, dy);
                        const float2 wd = __fmul2_rn(w, gd2);
                        const float su = u.x + u.y, suy = uy.x + uy.y;
                        const float sux = su * dx;
                        v[C + 0] = wd.x + wd.y;
                        v[C + 1] = su;
                        v[C + 2] = sux;
                        v[C + 3] = suy;
                        v[C + 4] = sux * dx;
                        v[C + 5] = suy * dx;
                        v[C + 6] = uyy.x + uyy.y;
                    }
                    ChunkReduceStore<V, 0>::run(v, lane, my_acc + j * V);
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
