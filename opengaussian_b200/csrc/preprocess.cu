// preprocess.cu -- per-Gaussian forward preprocessing (cull, cov3D, EWA cov2D, conic, radius,
// pixel centre, tile rect, SH -> RGB) and markVisible.
//
// Replaces upstream preprocessCUDA<3> / markVisible of the external rasterizer package that
// gaussian_renderer/__init__.py:55-70,104-112 calls (source absent from /root/reference; the
// algorithm is restated in SURVEY.md section 8c and oracle/raster_oracle.c).
//
// ARITHMETIC CONTRACT: this translation unit is compiled with --fmad=false.  Every expression is
// associated exactly as in oracle/raster_oracle.c so that radii, tile rects, depth bits (and hence
// sort keys, tile ranges) are bit-identical to the CPU oracle.  IEEE division and sqrt are nvcc's
// defaults (-prec-div=true -prec-sqrt=true).
//
// Memory behaviour (HBM-bound): one thread per Gaussian; a warp reads
// contiguous 384 B (means3D/scales), 512 B (rotations) slabs.  Geometry comes first; only the SH rows
// (192 B/Gaussian, 80 % of the input bytes) of the Gaussians that SURVIVED culling are then staged
// through shared memory with coalesced 16-byte loads and consumed from a conflict-free padded
// layout -- a culled Gaussian costs 44 B, not 236 B.  CTAs are small (128 Gaussians, 25 KB slab) so
// that several per SM overlap their load and compute phases.  Outputs are two 16-byte records per
// Gaussian (the gather set of the blend kernels).
#include "common.cuh"

namespace ogs {

__constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                 0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                 -0.5900435899266435f};
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f

__device__ __forceinline__ int imin_(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax_(int a, int b) { return a > b ? a : b; }

__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy, int& x0,
                                          int& y0, int& x1, int& y1) {
    const float r = (float)radius;
    x0 = imin_(gx, imax_(0, (int)((px - r) / 16.0f)));
    y0 = imin_(gy, imax_(0, (int)((py - r) / 16.0f)));
    x1 = imin_(gx, imax_(0, (int)((px + r + 15.0f) / 16.0f)));
    y1 = imin_(gy, imax_(0, (int)((py + r + 15.0f) / 16.0f)));
}


#define PF 128   // Gaussians (= threads) per CTA

// (normalize(ins_feat) + 1) / 2 -- scene/gaussian_model.py:161-169 + gaussian_renderer/__init__.py:127.  One function for
// the preprocess kernel and the cached-view refresh, so that both produce the same bits (this file: --fmad=false).
__device__ __forceinline__ void unit_half_row(const float* __restrict__ extra, float* __restrict__ feat, int i, int n_extra) {
    float n2 = 0.f;
    for (int c = 0; c < n_extra; c++) { const float v = extra[(size_t)i * n_extra + c]; n2 += v * v; }
    const float nrm = fmaxf(sqrtf(n2), 1e-12f);
    for (int c = 0; c < n_extra; c++) feat[(size_t)i * n_extra + c] = (extra[(size_t)i * n_extra + c] / nrm + 1.0f) / 2.0f;
}

#ifndef SH_UB
#define SH_UB 6     // 16-byte SH loads in flight per thread while the block's slab is staged (12 per thread at degree 3)
#endif

template <bool HAS_SH>
__global__ void __launch_bounds__(PF, HAS_SH ? 8 : 1) preprocess_fwd_kernel(PreprocessArgs a) {
    extern __shared__ float s_sh[];  // [PF][M*3 + 1] when HAS_SH
    __shared__ uint8_t s_vis[PF];
    const int i = blockIdx.x * PF + threadIdx.x;
    const int P = a.P;

    // defaults for culled Gaussians
    int radius_out = 0;
    uint32_t tiles = 0;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    float rgb[3] = {0.f, 0.f, 0.f};
    uint8_t clamp_bits = 0;
    uint32_t dkey = 0xFFFFFFFFu;

    const float* v = a.view;
    const float* pm = a.proj;
    const bool active = i < P;
    const float p0 = active ? a.means3D[3 * (size_t)i + 0] : 0.f, p1 = active ? a.means3D[3 * (size_t)i + 1] : 0.f,
                p2 = active ? a.means3D[3 * (size_t)i + 2] : 0.f;
    bool visible = false;
    float pv0 = v[0] * p0 + v[4] * p1 + v[8] * p2 + v[12];
    float pv1 = v[1] * p0 + v[5] * p1 + v[9] * p2 + v[13];
    float pv2 = v[2] * p0 + v[6] * p1 + v[10] * p2 + v[14];
    do {
        if (!active || pv2 <= 0.2f) break;  // near cull
        float ph0 = pm[0] * p0 + pm[4] * p1 + pm[8] * p2 + pm[12];
        float ph1 = pm[1] * p0 + pm[5] * p1 + pm[9] * p2 + pm[13];
        float ph3 = pm[3] * p0 + pm[7] * p1 + pm[11] * p2 + pm[15];
        float pw = 1.0f / (ph3 + 0.0000001f);
        float ndc_x = ph0 * pw, ndc_y = ph1 * pw;

        float c6[6];
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = a.cov3D_precomp[6 * (size_t)i + k];
        } else {
            float4 q = reinterpret_cast<const float4*>(a.rotations)[i];
            if (a.act_flags & OGS_ACT_ROT_NORMALIZE) {   // torch.nn.functional.normalize: x / max(|x|_2, 1e-12)
                const float nrm = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
                q.x /= nrm; q.y /= nrm; q.z /= nrm; q.w /= nrm;
            }
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            float R[3][3];
            R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
            R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
            R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
            const float mod = a.scale_modifier;
            float s[3] = {a.scales[3 * (size_t)i + 0], a.scales[3 * (size_t)i + 1], a.scales[3 * (size_t)i + 2]};
            if (a.act_flags & OGS_ACT_SCALE_EXP) { s[0] = expf(s[0]); s[1] = expf(s[1]); s[2] = expf(s[2]); }
            s[0] = mod * s[0]; s[1] = mod * s[1]; s[2] = mod * s[2];
            float L[3][3];
#pragma unroll
            for (int ii = 0; ii < 3; ii++)
#pragma unroll
                for (int jj = 0; jj < 3; jj++) L[ii][jj] = s[jj] * R[ii][jj];
            c6[0] = L[0][0] * L[0][0] + L[0][1] * L[0][1] + L[0][2] * L[0][2];
            c6[1] = L[0][0] * L[1][0] + L[0][1] * L[1][1] + L[0][2] * L[1][2];
            c6[2] = L[0][0] * L[2][0] + L[0][1] * L[2][1] + L[0][2] * L[2][2];
            c6[3] = L[1][0] * L[1][0] + L[1][1] * L[1][1] + L[1][2] * L[1][2];
            c6[4] = L[1][0] * L[2][0] + L[1][1] * L[2][1] + L[1][2] * L[2][2];
            c6[5] = L[2][0] * L[2][0] + L[2][1] * L[2][1] + L[2][2] * L[2][2];
        }

        // EWA projection
        const float fx = (float)a.W / (2.0f * a.tanfovx);
        const float fy = (float)a.H / (2.0f * a.tanfovy);
        const float tz = pv2;
        const float limx = 1.3f * a.tanfovx, limy = 1.3f * a.tanfovy;
        const float txtz = pv0 / tz, tytz = pv1 / tz;
        const float tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        const float ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float J00 = fx / tz, J02 = -(fx * tx) / (tz * tz);
        const float J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
        float T0[3], T1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            T0[k] = v[4 * k + 0] * J00 + v[4 * k + 2] * J02;
            T1[k] = v[4 * k + 1] * J11 + v[4 * k + 2] * J12;
        }
        const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
        float A0[3], A1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            A0[k] = T0[0] * V[0][k] + T0[1] * V[1][k] + T0[2] * V[2][k];
            A1[k] = T1[0] * V[0][k] + T1[1] * V[1][k] + T1[2] * V[2][k];
        }
        float cxx = A0[0] * T0[0] + A0[1] * T0[1] + A0[2] * T0[2];
        float cxy = A1[0] * T0[0] + A1[1] * T0[1] + A1[2] * T0[2];
        float cyy = A1[0] * T1[0] + A1[1] * T1[1] + A1[2] * T1[2];
        cxx += 0.3f;
        cyy += 0.3f;
        const float det = cxx * cyy - cxy * cxy;
        if (det == 0.0f) break;
        const float det_inv = 1.f / det;
        const float conA = cyy * det_inv, conB = -cxy * det_inv, conC = cxx * det_inv;
        const float mid = 0.5f * (cxx + cyy);
        const float root = sqrtf(fmaxf(0.1f, mid * mid - det));
        const float lam1 = mid + root, lam2 = mid - root;
        const float rad_f = ceilf(3.f * sqrtf(fmaxf(lam1, lam2)));
        const int rad = (int)rad_f;
        const float px = ((ndc_x + 1.0f) * (float)a.W - 1.0f) * 0.5f;
        const float py = ((ndc_y + 1.0f) * (float)a.H - 1.0f) * 0.5f;
        const int gx = (a.W + 15) / 16, gy = (a.H + 15) / 16;
        int x0, y0, x1, y1;
        tile_rect(px, py, rad, gx, gy, x0, y0, x1, y1);
        if ((x1 - x0) * (y1 - y0) == 0) break;

        visible = true;
        radius_out = rad;
        tiles = (uint32_t)((y1 - y0) * (x1 - x0));
        r0 = make_float4(px, py, conA, conB);
        float opac = a.opacities[i];
        if (a.act_flags & OGS_ACT_OPACITY_SIGMOID) opac = 1.0f / (1.0f + expf(-opac));
        r1 = make_float4(conC, opac, pv2, __int_as_float(rad));
        dkey = __float_as_uint(pv2);
    } while (0);

    if (HAS_SH) {
        // cooperative, coalesced stage of the VISIBLE rows of this block's SH slab
        s_vis[threadIdx.x] = visible;
        __syncthreads();
        const int per = a.M * 3;
        if (a.shs_rest) {
            const SmallDiv by_pr(per - 3);
            // split SH: row = [_features_dc (3) | _features_rest ((M-1)*3)], two contiguous slabs
            const int pr = per - 3;
            const size_t base_dc = (size_t)blockIdx.x * PF * 3, base_r = (size_t)blockIdx.x * PF * pr;
            const size_t tot_dc = (size_t)P * 3, tot_r = (size_t)P * pr;
            for (int e = threadIdx.x; e < PF * 3; e += PF) {
                const int gi = e / 3, k = e - gi * 3;
                if (s_vis[gi] && base_dc + e < tot_dc) s_sh[gi * (per + 1) + k] = __ldg(a.shs + base_dc + e);
            }
            // the slab base is 16-byte aligned (PF * pr * 4 bytes per CTA); a float4 may straddle two rows
            const float4* src = reinterpret_cast<const float4*>(a.shs_rest + base_r);
            const int n4 = (PF * pr) / 4;
            // SH_UB loads are issued before the first shared-memory store that depends on one of them: the loop used to
            // alternate LDG.128 -> STS (in-order issue stalls on the store's operand), i.e. one DRAM round trip per
            // 16 bytes and thread -- ncu: 8.8 long-scoreboard stall cycles per issued instruction
            for (int e0 = threadIdx.x; e0 < n4; e0 += PF * SH_UB) {
                float4 q[SH_UB];
                int mode[SH_UB];                   // 0: nothing, 1: one 16-byte load, 2: ragged end of the tensor
#pragma unroll
                for (int u = 0; u < SH_UB; u++) {
                    const int e = e0 + u * PF, f = e * 4;
                    mode[u] = 0;
                    if (e < n4) {
                        const int g0 = by_pr(f), g3 = by_pr(f + 3);
                        if ((s_vis[g0] || s_vis[g3 < PF ? g3 : g0]) && base_r + f + 3 < tot_r) { mode[u] = 1; q[u] = __ldg(src + e); }
                        else if (base_r + f < tot_r) mode[u] = 2;
                    }
                }
#pragma unroll
                for (int u = 0; u < SH_UB; u++) {
                    const int f = (e0 + u * PF) * 4;
                    if (mode[u] == 1) {
                        const float v4[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const int gi = by_pr(f + t), k = (f + t) - gi * pr;
                            if (s_vis[gi]) s_sh[gi * (per + 1) + 3 + k] = v4[t];
                        }
                    } else if (mode[u] == 2) {
                        for (int t = 0; t < 4 && base_r + f + t < tot_r; t++) {
                            const int gi = by_pr(f + t), k = (f + t) - gi * pr;
                            if (s_vis[gi]) s_sh[gi * (per + 1) + 3 + k] = __ldg(a.shs_rest + base_r + f + t);
                        }
                    }
                }
            }
            for (int e = n4 * 4 + threadIdx.x; e < PF * pr; e += PF) {
                const int gi = by_pr(e), k = e - gi * pr;
                if (s_vis[gi] && base_r + e < tot_r) s_sh[gi * (per + 1) + 3 + k] = __ldg(a.shs_rest + base_r + e);
            }
        } else {
        const size_t base = (size_t)blockIdx.x * PF * per;
        const SmallDiv by_per(per);
        if ((per & 3) == 0) {
            const float4* src = reinterpret_cast<const float4*>(a.shs + base);
            const int total = PF / 4 * per;
            for (int e0 = threadIdx.x; e0 < total; e0 += PF * SH_UB) {       // loads first, stores after (see above)
                float4 q[SH_UB];
                int dsti[SH_UB];
#pragma unroll
                for (int u = 0; u < SH_UB; u++) {
                    const int e = e0 + u * PF, f = e * 4;
                    dsti[u] = -1;
                    if (e < total) {
                        const int gi = by_per(f);                             // per % 4 == 0: the 4 floats share gi
                        if (s_vis[gi]) { dsti[u] = gi * (per + 1) + (f - gi * per); q[u] = __ldg(src + e); }
                    }
                }
#pragma unroll
                for (int u = 0; u < SH_UB; u++) {
                    if (dsti[u] >= 0) {
                        float* d = s_sh + dsti[u];
                        d[0] = q[u].x; d[1] = q[u].y; d[2] = q[u].z; d[3] = q[u].w;
                    }
                }
            }
        } else {
            for (int e = threadIdx.x; e < PF * per; e += PF) {
                const int gi = by_per(e), k = e - gi * per;
                if (s_vis[gi]) s_sh[gi * (per + 1) + k] = __ldg(a.shs + base + e);
            }
        }
        }
        __syncthreads();
        if (visible) {
            const float* sh = s_sh + threadIdx.x * (a.M * 3 + 1);
            const float d0 = p0 - a.campos[0], d1 = p1 - a.campos[1], d2 = p2 - a.campos[2];
            const float len = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
            const float x = d0 / len, y = d1 / len, z = d2 / len;
            const int deg = a.D;
#pragma unroll
            for (int c = 0; c < 3; c++) {
#define S(k) sh[(k) * 3 + c]
                float res = SH_C0 * S(0);
                if (deg > 0) {
                    res = res - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
                    if (deg > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        res = res + c_SH_C2[0] * xy * S(4) + c_SH_C2[1] * yz * S(5) +
                              c_SH_C2[2] * (2.0f * zz - xx - yy) * S(6) + c_SH_C2[3] * xz * S(7) +
                              c_SH_C2[4] * (xx - yy) * S(8);
                        if (deg > 2) {
                            res = res + c_SH_C3[0] * y * (3.0f * xx - yy) * S(9) + c_SH_C3[1] * xy * z * S(10) +
                                  c_SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                                  c_SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                                  c_SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) +
                                  c_SH_C3[5] * z * (xx - yy) * S(14) + c_SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                        }
                    }
                }
#undef S
                res += 0.5f;
                if (res < 0.0f) clamp_bits |= (uint8_t)(1u << c);
                rgb[c] = fmaxf(res, 0.0f);
            }
        }
    }
    if (!active) return;

    if (visible && (a.act_flags & OGS_ACT_EXTRA_UNIT_HALF)) unit_half_row(a.extra, a.g.feat, i, a.n_extra);
    a.radii[i] = radius_out;
    a.g.rec0[i] = r0;
    a.g.rec1[i] = r1;
    a.g.tiles[i] = tiles;
    if (HAS_SH) {
        a.g.rgb[3 * (size_t)i + 0] = rgb[0];
        a.g.rgb[3 * (size_t)i + 1] = rgb[1];
        a.g.rgb[3 * (size_t)i + 2] = rgb[2];
        a.g.clamped[i] = clamp_bits;
    }
    a.depth_keys[i] = dkey;
    a.depth_vals[i] = (uint32_t)i;
}

int launch_preprocess_forward(const PreprocessArgs& a, cudaStream_t s) {
    if (a.P <= 0) return 0;
    const int blocks = (a.P + PF - 1) / PF;
    if (a.shs) {
        const size_t smem = (size_t)PF * (a.M * 3 + 1) * sizeof(float);
        static PerDeviceOnce attr_done;
        if (attr_done.todo()) {
            OGS_CUDA(cudaFuncSetAttribute(preprocess_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            attr_done.done();
        }
        if (smem > 100 * 1024) { set_error("preprocess: SH block too large (M=%d)", a.M); return -3; }
        preprocess_fwd_kernel<true><<<blocks, PF, smem, s>>>(a);
    } else {
        preprocess_fwd_kernel<false><<<blocks, PF, 0, s>>>(a);
    }
    return 0;
}

// Cached-view forward (ogs_raster_forward_cached): the only per-Gaussian quantity that changes between two calls on a
// frozen geometry.  Visible Gaussians only, like the preprocess kernel (rec1.w = radius as int bits).
__global__ void __launch_bounds__(256) feat_refresh_kernel(int P, int n_extra, const float4* __restrict__ rec1,
                                                           const float* __restrict__ extra, float* __restrict__ feat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    if (__float_as_int(__ldg(rec1 + i).w) > 0) unit_half_row(extra, feat, i, n_extra);
}

int launch_feat_refresh(int P, int n_extra, const float4* rec1, const float* extra, float* feat, cudaStream_t s) {
    if (P <= 0 || n_extra <= 0) return 0;
    feat_refresh_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, n_extra, rec1, extra, feat);
    return 0;
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ v,
                                    uint8_t* __restrict__ present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float p0 = means3D[3 * (size_t)i], p1 = means3D[3 * (size_t)i + 1], p2 = means3D[3 * (size_t)i + 2];
    const float z = v[2] * p0 + v[6] * p1 + v[10] * p2 + v[14];
    present[i] = (uint8_t)(z > 0.2f);
}

int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s) {
    if (P <= 0) return 0;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
    return 0;
}

}  // namespace ogs
