// preprocess_bwd.cu -- per-Gaussian chain rule from the blend accumulators to the inputs.
//
// Replaces upstream computeCov2DCUDA + preprocessCUDA<3> backward of the external rasterizer
// (SURVEY.md section 2b / 8a8): conic -> cov2D -> (cov3D, mean); mean2D -> mean3D through the
// projective division; depth -> mean3D; SH backward incl. direction normalisation and the clamp
// mask; cov3D -> (scale, quaternion).  Same formulas as oracle/raster_oracle.c
// (ogs_oracle_preprocess_backward), evaluated in fp32; cov3D is recomputed instead of being
// stored by the forward (saves 48 B/Gaussian of HBM traffic).
// HBM-bound: reads (C+7)*4 B accumulators + ~236 B parameters, writes the gradient rows.
#include "common.cuh"

namespace ogs {

#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
__constant__ float b_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__constant__ float b_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                 0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                 -0.5900435899266435f};

// out = v, or out += v in accumulate mode (the caller hands the parameter's existing .grad: summing the
// views of a step then costs one read of the old gradient here instead of a separate read-read-write pass)
// In accumulate mode the add is a fire-and-forget red.global (every address has exactly ONE writer, so the result
// is deterministic): no dependent load at the tail of the thread.
__device__ __forceinline__ void put(float* p, float v, bool accum) {
    if (accum) atomicAdd(p, v);
    else *p = v;
}

// STAGE_SH: the block's SH slab (visible rows only) is loaded with coalesced 16-byte loads into a
// padded shared-memory layout, each thread overwrites its row in place with dL/dsh, and the slab
// is written back coalesced -- instead of 48 strided 4-byte accesses per thread in each direction.
template <bool STAGE_SH>
__device__ __forceinline__ void preprocess_bwd_body(const PreprocessBwdArgs& a, int i, float* s_row, bool vis);

#ifndef PB
#define PB 128   // Gaussians (= threads) per CTA: 25 KB SH slab, several CTAs per SM overlap their load / compute / store phases
#endif
#ifndef PB_MINB
#define PB_MINB 8   // <= 64 registers: twice the resident warps outweighs ~70 B of spills (0.158 -> 0.144 ms)
#endif
#ifndef SHB_UB
#define SHB_UB 4    // 16-byte SH loads in flight per thread while the slab is staged (64-register cap)
#endif
template <bool STAGE_SH>
__global__ void __launch_bounds__(PB, PB_MINB) preprocess_bwd_kernel(PreprocessBwdArgs a) {
    extern __shared__ float s_sh[];
    __shared__ uint8_t s_vis[PB];
    const int i = blockIdx.x * PB + threadIdx.x;
    const bool active = i < a.P;
    bool vis = false;
    if (active) vis = __float_as_int(a.g.rec1[i].w) > 0;
    const int per = a.M * 3;
    const SmallDiv by_per(per > 0 ? per : 1), by_pr(per > 3 ? per - 3 : 1);
    if (STAGE_SH) {
        s_vis[threadIdx.x] = vis;
        __syncthreads();
        if (a.shs_rest) {   // split SH: row = [_features_dc (3) | _features_rest ((M-1)*3)]
            const int pr = per - 3;
            const size_t base_dc = (size_t)blockIdx.x * PB * 3, base_r = (size_t)blockIdx.x * PB * pr;
            for (int e = threadIdx.x; e < PB * 3; e += PB) {
                const int gi = e / 3, k = e - gi * 3;
                if (s_vis[gi]) s_sh[gi * (per + 1) + k] = __ldg(a.shs + base_dc + e);
            }
            // loads in batches of SHB_UB before the dependent shared-memory stores (in-order issue would otherwise make
            // every element its own DRAM round trip: ncu showed 8.3 long-scoreboard stall cycles per instruction)
            const size_t tot_r = (size_t)a.P * pr;
            for (int e0 = threadIdx.x; e0 < PB * pr; e0 += PB * SHB_UB * 2) {
                float q[SHB_UB * 2];
                int dsti[SHB_UB * 2];
#pragma unroll
                for (int u = 0; u < SHB_UB * 2; u++) {
                    const int e = e0 + u * PB;
                    dsti[u] = -1;
                    if (e < PB * pr) {
                        const int gi = by_pr(e);
                        if (s_vis[gi] && base_r + e < tot_r) { dsti[u] = gi * (per + 1) + 3 + (e - gi * pr); q[u] = __ldg(a.shs_rest + base_r + e); }
                    }
                }
#pragma unroll
                for (int u = 0; u < SHB_UB * 2; u++)
                    if (dsti[u] >= 0) s_sh[dsti[u]] = q[u];
            }
        } else {
        const size_t base = (size_t)blockIdx.x * PB * per;
        const float4* src = reinterpret_cast<const float4*>(a.shs + base);
        const int total = PB / 4 * per;                              // PB * per / 4 float4 (per % 4 == 0)
        for (int e0 = threadIdx.x; e0 < total; e0 += PB * SHB_UB) {
            float4 q[SHB_UB];
            int dsti[SHB_UB];
#pragma unroll
            for (int u = 0; u < SHB_UB; u++) {
                const int e = e0 + u * PB, f = e * 4;
                dsti[u] = -1;
                if (e < total) {
                    const int gi = by_per(f);
                    if (s_vis[gi]) { dsti[u] = gi * (per + 1) + (f - gi * per); q[u] = __ldg(src + e); }
                }
            }
#pragma unroll
            for (int u = 0; u < SHB_UB; u++) {
                if (dsti[u] >= 0) {
                    float* d = s_sh + dsti[u];
                    d[0] = q[u].x; d[1] = q[u].y; d[2] = q[u].z; d[3] = q[u].w;
                }
            }
        }
        }
        __syncthreads();
    }
    if (active) preprocess_bwd_body<STAGE_SH>(a, i, STAGE_SH ? s_sh + threadIdx.x * (per + 1) : nullptr, vis);
    if (STAGE_SH && a.dL_dshs) {
        __syncthreads();
        const int rows = min(PB, a.P - blockIdx.x * PB);
        const bool ac = (a.accumulate & 1) != 0;
        if (a.shs_rest) {
            const int pr = per - 3;
            float* dst_dc = a.dL_dshs + (size_t)blockIdx.x * PB * 3;
            for (int e = threadIdx.x; e < rows * 3; e += PB) {
                const int gi = e / 3, k = e - gi * 3;
                if (!ac) dst_dc[e] = s_sh[gi * (per + 1) + k];
                else if (s_vis[gi]) atomicAdd(dst_dc + e, s_sh[gi * (per + 1) + k]);
            }
            if (a.dL_dshs_rest) {
                float* dst_r = a.dL_dshs_rest + (size_t)blockIdx.x * PB * pr;
                for (int e = threadIdx.x; e < rows * pr; e += PB) {
                    const int gi = by_pr(e), k = e - gi * pr;
                    if (!ac) dst_r[e] = s_sh[gi * (per + 1) + 3 + k];
                    else if (s_vis[gi]) atomicAdd(dst_r + e, s_sh[gi * (per + 1) + 3 + k]);
                }
            }
        } else {
            const size_t base = (size_t)blockIdx.x * PB * per;
            float4* dst = reinterpret_cast<float4*>(a.dL_dshs + base);
            for (int e = threadIdx.x; e < rows * per / 4; e += PB) {
                const int f = e * 4;
                const int gi = by_per(f), k = f - gi * per;
                const float* d = s_sh + gi * (per + 1) + k;
                if (!ac) dst[e] = make_float4(d[0], d[1], d[2], d[3]);
                else if (s_vis[gi]) atomicAdd(dst + e, make_float4(d[0], d[1], d[2], d[3]));   // red.global.add.v4.f32
            }
        }
    }
}

template <bool STAGE_SH>
__device__ __forceinline__ void preprocess_bwd_body(const PreprocessBwdArgs& a, int i, float* s_row, bool vis) {
    const int C = a.C;
    const float* acc = a.acc + (size_t)i * a.stride;

    // colour / feature gradients pass straight through
    if (a.dL_dcolors_precomp) {
        if (vis || !(a.accumulate & 1))
            for (int c = 0; c < 3; c++) put(a.dL_dcolors_precomp + 3 * (size_t)i + c, vis ? acc[c] : 0.f, (a.accumulate & 1) != 0);
    }
    if (a.dL_dextra) {
        const int F = C - 3;
        if (vis && (a.act_flags & OGS_ACT_EXTRA_UNIT_HALF)) {
            // f = (x / n + 1) / 2, n = max(|x|, 1e-12):  dL/dx = (g - u <u, g>) / (2 n),  u = x / n
            float n2 = 0.f;
            for (int c = 0; c < F; c++) { const float v = a.extra[(size_t)F * i + c]; n2 += v * v; }
            const float nrm = fmaxf(sqrtf(n2), 1e-12f);
            float dotug = 0.f;
            for (int c = 0; c < F; c++) dotug += (a.extra[(size_t)F * i + c] / nrm) * acc[3 + c];
            for (int c = 0; c < F; c++)
                put(a.dL_dextra + (size_t)F * i + c, (acc[3 + c] - (a.extra[(size_t)F * i + c] / nrm) * dotug) / (2.0f * nrm), (a.accumulate & 1) != 0);
        } else {
            if (vis || !(a.accumulate & 1))
                for (int c = 3; c < C; c++) put(a.dL_dextra + (size_t)F * i + (c - 3), vis ? acc[c] : 0.f, (a.accumulate & 1) != 0);
        }
    }
    if (!a.geom) return;

    float dmean[3] = {0.f, 0.f, 0.f};
    float dscale[3] = {0.f, 0.f, 0.f};
    float drot[4] = {0.f, 0.f, 0.f, 0.f};
    float dc6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dm2x = 0.f, dm2y = 0.f, dop = 0.f;
    const int M = a.M;

    if (vis) {
        const float* v = a.view;
        const float* proj = a.proj;
        const float m0 = a.means3D[3 * (size_t)i], m1 = a.means3D[3 * (size_t)i + 1], m2 = a.means3D[3 * (size_t)i + 2];
        const float dL_ddepth = acc[C + 0];
        // blend_bwd delivers the raw moments of u = G dL/dalpha (common.cuh, acc layout); with
        // dL/dG = opacity * dL/dalpha: dL/dmean2D = -(W/2, H/2) * conic * (sum dL/dG G d),
        // dL/dconic = -1/2 sum dL/dG G d d^T  (b carries the half factor of the symmetric pair)
        float dLA, dLBh, dLC;
        {
            const float4 q0 = a.g.rec0[i], q1 = a.g.rec1[i];
            const float op = q1.y;
            const float e1 = op * acc[C + 2], e2 = op * acc[C + 3];
            dop = acc[C + 1];
            dm2x = -0.5f * (float)a.W * (q0.z * e1 + q0.w * e2);
            dm2y = -0.5f * (float)a.H * (q1.x * e2 + q0.w * e1);
            const float hop = -0.5f * op;
            dLA = hop * acc[C + 4]; dLBh = hop * acc[C + 5]; dLC = hop * acc[C + 6];
        }

        // cov3D (recomputed)
        float c6[6];
        float R[3][3];
        float s[3] = {0.f, 0.f, 0.f};
        float qr = 0.f, qx = 0.f, qy = 0.f, qz = 0.f, qnorm = 1.f;
        float sact[3] = {0.f, 0.f, 0.f};
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = a.cov3D_precomp[6 * (size_t)i + k];
        } else {
            float4 q = reinterpret_cast<const float4*>(a.rotations)[i];
            if (a.act_flags & OGS_ACT_ROT_NORMALIZE) {
                qnorm = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
                q.x /= qnorm; q.y /= qnorm; q.z /= qnorm; q.w /= qnorm;
            }
            qr = q.x; qx = q.y; qy = q.z; qz = q.w;
            const float r = qr, x = qx, y = qy, z = qz;
            R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
            R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
            R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                sact[k] = a.scales[3 * (size_t)i + k];
                if (a.act_flags & OGS_ACT_SCALE_EXP) sact[k] = expf(sact[k]);
                s[k] = a.scale_modifier * sact[k];
            }
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

        // ---- cov2D backward ----
        const float fx = (float)a.W / (2.0f * a.tanfovx);
        const float fy = (float)a.H / (2.0f * a.tanfovy);
        float t0 = v[0] * m0 + v[4] * m1 + v[8] * m2 + v[12];
        float t1 = v[1] * m0 + v[5] * m1 + v[9] * m2 + v[13];
        const float t2 = v[2] * m0 + v[6] * m1 + v[10] * m2 + v[14];
        const float limx = 1.3f * a.tanfovx, limy = 1.3f * a.tanfovy;
        const float txtz = t0 / t2, tytz = t1 / t2;
        t0 = fminf(limx, fmaxf(-limx, txtz)) * t2;
        t1 = fminf(limy, fmaxf(-limy, tytz)) * t2;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float J00 = fx / t2, J02 = -(fx * t0) / (t2 * t2);
        const float J11 = fy / t2, J12 = -(fy * t1) / (t2 * t2);
        float T0[3], T1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            T0[k] = v[4 * k + 0] * J00 + v[4 * k + 2] * J02;
            T1[k] = v[4 * k + 1] * J11 + v[4 * k + 2] * J12;
        }
        const float Vm[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
        float VT0[3], VT1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            VT0[k] = Vm[k][0] * T0[0] + Vm[k][1] * T0[1] + Vm[k][2] * T0[2];
            VT1[k] = Vm[k][0] * T1[0] + Vm[k][1] * T1[1] + Vm[k][2] * T1[2];
        }
        const float ca = T0[0] * VT0[0] + T0[1] * VT0[1] + T0[2] * VT0[2] + 0.3f;
        const float cb = T0[0] * VT1[0] + T0[1] * VT1[1] + T0[2] * VT1[2];
        const float cc = T1[0] * VT1[0] + T1[1] * VT1[1] + T1[2] * VT1[2] + 0.3f;
        const float denom = ca * cc - cb * cb;
        const float d2inv = 1.0f / (denom * denom + 0.0000001f);
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        if (d2inv != 0.f) {
            dL_da = d2inv * (-cc * cc * dLA + 2.f * cb * cc * dLBh + (denom - ca * cc) * dLC);
            dL_dc = d2inv * (-ca * ca * dLC + 2.f * ca * cb * dLBh + (denom - ca * cc) * dLA);
            dL_db = d2inv * 2.f * (cb * cc * dLA - (denom + 2.f * cb * cb) * dLBh + ca * cb * dLC);
            dc6[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
            dc6[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
            dc6[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
            dc6[1] = 2.f * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2.f * T1[0] * T1[1] * dL_dc;
            dc6[2] = 2.f * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2.f * T1[0] * T1[2] * dL_dc;
            dc6[4] = 2.f * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2.f * T1[1] * T1[2] * dL_dc;
        }
        float dT0[3], dT1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dT0[k] = 2.f * VT0[k] * dL_da + VT1[k] * dL_db;
            dT1[k] = 2.f * VT1[k] * dL_dc + VT0[k] * dL_db;
        }
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dJ00 += v[4 * k + 0] * dT0[k]; dJ02 += v[4 * k + 2] * dT0[k];
            dJ11 += v[4 * k + 1] * dT1[k]; dJ12 += v[4 * k + 2] * dT1[k];
        }
        const float tz = 1.0f / t2, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = x_grad_mul * -fx * tz2 * dJ02;
        const float dty = y_grad_mul * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * t0) * tz3 * dJ02 + (2.f * fy * t1) * tz3 * dJ12;
        dmean[0] = v[0] * dtx + v[1] * dty + v[2] * dtz;
        dmean[1] = v[4] * dtx + v[5] * dty + v[6] * dtz;
        dmean[2] = v[8] * dtx + v[9] * dty + v[10] * dtz;

        // ---- mean2D (NDC-scaled) -> mean3D ----
        const float mh0 = proj[0] * m0 + proj[4] * m1 + proj[8] * m2 + proj[12];
        const float mh1 = proj[1] * m0 + proj[5] * m1 + proj[9] * m2 + proj[13];
        const float mh3 = proj[3] * m0 + proj[7] * m1 + proj[11] * m2 + proj[15];
        const float m_w = 1.0f / (mh3 + 0.0000001f);
        const float mul1 = mh0 * m_w * m_w, mul2 = mh1 * m_w * m_w;
        dmean[0] += (proj[0] * m_w - proj[3] * mul1) * dm2x + (proj[1] * m_w - proj[3] * mul2) * dm2y;
        dmean[1] += (proj[4] * m_w - proj[7] * mul1) * dm2x + (proj[5] * m_w - proj[7] * mul2) * dm2y;
        dmean[2] += (proj[8] * m_w - proj[11] * mul1) * dm2x + (proj[9] * m_w - proj[11] * mul2) * dm2y;
        // ---- depth = p_view.z ----
        dmean[0] += v[2] * dL_ddepth;
        dmean[1] += v[6] * dL_ddepth;
        dmean[2] += v[10] * dL_ddepth;

        // ---- SH backward ----
        if (a.shs) {
            const float* sh = STAGE_SH ? s_row : a.shs + (size_t)i * M * 3;
            float* dsh = STAGE_SH ? s_row : (a.dL_dshs ? a.dL_dshs + (size_t)i * M * 3 : nullptr);
            const float d0 = m0 - a.campos[0], d1 = m1 - a.campos[1], d2 = m2 - a.campos[2];
            const float len = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
            const float x = d0 / len, y = d1 / len, z = d2 / len;
            const uint8_t cl = a.g.clamped[i];
            float dRGB[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = ((cl >> ch) & 1) ? 0.f : acc[ch];
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
            const int deg = a.D;
            auto term = [&](int k, float basis, float bx, float by, float bz) {
                float sx = 0.f;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    sx += sh[k * 3 + ch] * dRGB[ch];        // read before the in-place overwrite
                    if (dsh) dsh[k * 3 + ch] = basis * dRGB[ch];
                }
                ddx += bx * sx; ddy += by * sx; ddz += bz * sx;
            };
            term(0, SH_C0, 0.f, 0.f, 0.f);
            if (deg > 0) {
                term(1, -SH_C1 * y, 0.f, -SH_C1, 0.f);
                term(2, SH_C1 * z, 0.f, 0.f, SH_C1);
                term(3, -SH_C1 * x, -SH_C1, 0.f, 0.f);
                if (deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    term(4, b_SH_C2[0] * xy, b_SH_C2[0] * y, b_SH_C2[0] * x, 0.f);
                    term(5, b_SH_C2[1] * yz, 0.f, b_SH_C2[1] * z, b_SH_C2[1] * y);
                    term(6, b_SH_C2[2] * (2.f * zz - xx - yy), b_SH_C2[2] * -2.f * x, b_SH_C2[2] * -2.f * y, b_SH_C2[2] * 4.f * z);
                    term(7, b_SH_C2[3] * xz, b_SH_C2[3] * z, 0.f, b_SH_C2[3] * x);
                    term(8, b_SH_C2[4] * (xx - yy), b_SH_C2[4] * 2.f * x, b_SH_C2[4] * -2.f * y, 0.f);
                    if (deg > 2) {
                        term(9, b_SH_C3[0] * y * (3.f * xx - yy), b_SH_C3[0] * 6.f * xy, b_SH_C3[0] * (3.f * xx - 3.f * yy), 0.f);
                        term(10, b_SH_C3[1] * xy * z, b_SH_C3[1] * yz, b_SH_C3[1] * xz, b_SH_C3[1] * xy);
                        term(11, b_SH_C3[2] * y * (4.f * zz - xx - yy), b_SH_C3[2] * -2.f * xy, b_SH_C3[2] * (4.f * zz - xx - 3.f * yy), b_SH_C3[2] * 8.f * yz);
                        term(12, b_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy), b_SH_C3[3] * -6.f * xz, b_SH_C3[3] * -6.f * yz, b_SH_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy));
                        term(13, b_SH_C3[4] * x * (4.f * zz - xx - yy), b_SH_C3[4] * (4.f * zz - 3.f * xx - yy), b_SH_C3[4] * -2.f * xy, b_SH_C3[4] * 8.f * xz);
                        term(14, b_SH_C3[5] * z * (xx - yy), b_SH_C3[5] * 2.f * xz, b_SH_C3[5] * -2.f * yz, b_SH_C3[5] * (xx - yy));
                        term(15, b_SH_C3[6] * x * (xx - 3.f * yy), b_SH_C3[6] * (3.f * xx - 3.f * yy), b_SH_C3[6] * -6.f * xy, 0.f);
                    }
                }
            }
            if (dsh) {
                const int ncoef = (deg + 1) * (deg + 1);
                for (int k = ncoef * 3; k < M * 3; k++) dsh[k] = 0.f;
            }
            const float sum2 = len * len, invsum32 = 1.0f / (sum2 * len);
            dmean[0] += ((sum2 - d0 * d0) * ddx - d1 * d0 * ddy - d2 * d0 * ddz) * invsum32;
            dmean[1] += (-d0 * d1 * ddx + (sum2 - d1 * d1) * ddy - d2 * d1 * ddz) * invsum32;
            dmean[2] += (-d0 * d2 * ddx - d1 * d2 * ddy + (sum2 - d2 * d2) * ddz) * invsum32;
        }

        // ---- cov3D -> scale, quaternion ----
        if (!a.cov3D_precomp) {
            const float r = qr, x = qx, y = qy, z = qz;
            const float Gs[3][3] = {{dc6[0], 0.5f * dc6[1], 0.5f * dc6[2]}, {0.5f * dc6[1], dc6[3], 0.5f * dc6[4]}, {0.5f * dc6[2], 0.5f * dc6[4], dc6[5]}};
            float dLm[3][3], dR[3][3];
#pragma unroll
            for (int aa = 0; aa < 3; aa++)
#pragma unroll
                for (int bb = 0; bb < 3; bb++) {
                    float t = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; k++) t += Gs[aa][k] * R[k][bb] * s[bb];
                    dLm[aa][bb] = 2.f * t;
                }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                float ds = 0.f;
#pragma unroll
                for (int aa = 0; aa < 3; aa++) { ds += R[aa][j] * dLm[aa][j]; dR[aa][j] = dLm[aa][j] * s[j]; }
                dscale[j] = ds * a.scale_modifier;
            }
            drot[0] = 2.f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
            drot[1] = 2.f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.f * x * dR[1][1] - r * dR[1][2] + z * dR[2][0] + r * dR[2][1] - 2.f * x * dR[2][2]);
            drot[2] = 2.f * (-2.f * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] - r * dR[2][0] + z * dR[2][1] - 2.f * y * dR[2][2]);
            drot[3] = 2.f * (-2.f * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - 2.f * z * dR[1][1] + y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
            if (a.act_flags & OGS_ACT_SCALE_EXP) {         // d exp(x) = exp(x) dx
#pragma unroll
                for (int j = 0; j < 3; j++) dscale[j] *= sact[j];
            }
            if (a.act_flags & OGS_ACT_ROT_NORMALIZE) {     // q = x / n:  dL/dx = (g - q <q, g>) / n
                const float dq = r * drot[0] + x * drot[1] + y * drot[2] + z * drot[3];
                drot[0] = (drot[0] - r * dq) / qnorm; drot[1] = (drot[1] - x * dq) / qnorm;
                drot[2] = (drot[2] - y * dq) / qnorm; drot[3] = (drot[3] - z * dq) / qnorm;
            }
        }
        if (a.act_flags & OGS_ACT_OPACITY_SIGMOID) {       // o = sigmoid(x): dL/dx = dL/do o (1 - o)
            const float o = a.g.rec1[i].y;
            dop *= o * (1.0f - o);
        }
    } else if (a.shs && (STAGE_SH || a.dL_dshs)) {
        float* dsh = STAGE_SH ? s_row : a.dL_dshs + (size_t)i * M * 3;
        for (int k = 0; k < M * 3; k++) dsh[k] = 0.f;
    }

    // accumulate bit 0: the parameter gradients are added to; bit 1: dL_dmeans2D too (it is a per-view statistic
    // -- render() feeds a fresh tensor every call -- so it usually is NOT accumulated)
    const bool ac = (a.accumulate & 1) != 0, ac2 = (a.accumulate & 2) != 0;
    if (a.dL_dmeans2D) {
        if (vis || !ac2) {
            put(a.dL_dmeans2D + 3 * (size_t)i + 0, dm2x, ac2);
            put(a.dL_dmeans2D + 3 * (size_t)i + 1, dm2y, ac2);
            if (!ac2) a.dL_dmeans2D[3 * (size_t)i + 2] = 0.f;
        }
    }
    if (ac && !vis) return;                    // nothing to add
    if (a.dL_dmeans3D) {
#pragma unroll
        for (int k = 0; k < 3; k++) put(a.dL_dmeans3D + 3 * (size_t)i + k, dmean[k], ac);
    }
    if (a.dL_dopacities) put(a.dL_dopacities + i, dop, ac);
    if (a.dL_dscales) {
#pragma unroll
        for (int k = 0; k < 3; k++) put(a.dL_dscales + 3 * (size_t)i + k, dscale[k], ac);
    }
    if (a.dL_drotations) {
#pragma unroll
        for (int k = 0; k < 4; k++) put(a.dL_drotations + 4 * (size_t)i + k, drot[k], ac);
    }
    if (a.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) put(a.dL_dcov3D + 6 * (size_t)i + k, dc6[k], ac);
    }
}

// Colour-only backward of the raw-parameter feature channels (OpenGaussian stages 1-2: only `_ins_feat` trains,
// train.py:431-436): dL/dx of f = (x / n + 1) / 2, n = max(|x|, 1e-12), from the blend sums in columns [3, 3 + F) of the
// accumulator rows.  The generic kernel above walks its rows with 4-byte strided accesses under a 64-register cap
// (46 us at 1 M Gaussians); here a CTA moves its 256 rows of `extra` and of the accumulator through shared memory with
// 16-byte coalesced loads and stores (rows of culled Gaussians hold zeros after the memset, so no visibility lookup).
// Same expressions, in the same order, as preprocess_bwd_body.
#define FG_ROWS 256
template <int F>
__global__ void __launch_bounds__(FG_ROWS) feat_grad_kernel(int P, int stride, const float* __restrict__ extra,
                                                            const float* __restrict__ acc, float* __restrict__ dL_dextra,
                                                            int accumulate) {
    constexpr int XS = F | 1;                       // odd row strides: conflict-free column walks
    extern __shared__ float s_fg[];
    float* s_x = s_fg;                              // [FG_ROWS][XS]
    float* s_a = s_fg + FG_ROWS * XS;               // [FG_ROWS][stride | 1]
    const int AS = stride | 1;
    const int row0 = blockIdx.x * FG_ROWS;
    const int rows = min(FG_ROWS, P - row0);
    {
        const float4* src = reinterpret_cast<const float4*>(extra + (size_t)row0 * F);
        const int n = rows * F, n4 = n / 4;
        for (int e = threadIdx.x; e < n4; e += FG_ROWS) {
            const float4 v = __ldg(src + e);
            const float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) { const int f = 4 * e + k, r = f / F; s_x[r * XS + (f - r * F)] = t[k]; }
        }
        for (int f = 4 * n4 + threadIdx.x; f < n; f += FG_ROWS) { const int r = f / F; s_x[r * XS + (f - r * F)] = __ldg(extra + (size_t)row0 * F + f); }
    }
    {
        const float4* src = reinterpret_cast<const float4*>(acc + (size_t)row0 * stride);
        const int n = rows * stride, n4 = n / 4;
        for (int e = threadIdx.x; e < n4; e += FG_ROWS) {
            const float4 v = __ldg(src + e);
            const float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) { const int f = 4 * e + k, r = f / stride; s_a[r * AS + (f - r * stride)] = t[k]; }
        }
        for (int f = 4 * n4 + threadIdx.x; f < n; f += FG_ROWS) { const int r = f / stride; s_a[r * AS + (f - r * stride)] = __ldg(acc + (size_t)row0 * stride + f); }
    }
    __syncthreads();
    if (threadIdx.x < rows) {
        float* x = s_x + threadIdx.x * XS;
        const float* g = s_a + threadIdx.x * AS + 3;
        float n2 = 0.f;
#pragma unroll
        for (int c = 0; c < F; c++) n2 += x[c] * x[c];
        const float nrm = fmaxf(sqrtf(n2), 1e-12f);
        float dotug = 0.f;
#pragma unroll
        for (int c = 0; c < F; c++) dotug += (x[c] / nrm) * g[c];
#pragma unroll
        for (int c = 0; c < F; c++) x[c] = (g[c] - (x[c] / nrm) * dotug) / (2.0f * nrm);
    }
    __syncthreads();
    {
        float* dst = dL_dextra + (size_t)row0 * F;
        const int n = rows * F, n4 = n / 4;
        for (int e = threadIdx.x; e < n4; e += FG_ROWS) {
            float t[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { const int f = 4 * e + k, r = f / F; t[k] = s_x[r * XS + (f - r * F)]; }
            float4* d4 = reinterpret_cast<float4*>(dst) + e;
            if (accumulate) { const float4 o = *d4; t[0] += o.x; t[1] += o.y; t[2] += o.z; t[3] += o.w; }
            *d4 = make_float4(t[0], t[1], t[2], t[3]);
        }
        for (int f = 4 * n4 + threadIdx.x; f < n; f += FG_ROWS) {
            const int r = f / F;
            const float v = s_x[r * XS + (f - r * F)];
            dst[f] = accumulate ? dst[f] + v : v;
        }
    }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int launch_preprocess_backward(const PreprocessBwdArgs& a, cudaStream_t s) {
    if (a.P <= 0) return 0;
    const int F = a.C - 3;
    if (!a.geom && a.dL_dextra && !a.dL_dcolors_precomp && (a.act_flags & OGS_ACT_EXTRA_UNIT_HALF) && (F == 6 || F == 3) &&
        aligned16(a.extra) && aligned16(a.acc) && aligned16(a.dL_dextra)) {
        const size_t smem = (size_t)FG_ROWS * ((F | 1) + (a.stride | 1)) * sizeof(float);
        const int grid = (a.P + FG_ROWS - 1) / FG_ROWS;
        if (F == 6) feat_grad_kernel<6><<<grid, FG_ROWS, smem, s>>>(a.P, a.stride, a.extra, a.acc, a.dL_dextra, a.accumulate & 1);
        else feat_grad_kernel<3><<<grid, FG_ROWS, smem, s>>>(a.P, a.stride, a.extra, a.acc, a.dL_dextra, a.accumulate & 1);
        return 0;
    }
    const int per = a.M * 3;
    const size_t smem = (size_t)PB * (per + 1) * sizeof(float);
    if (a.shs_rest && !(a.geom && smem <= 100 * 1024)) {
        if (a.geom) { set_error("preprocess backward: split SH needs the staged path (M=%d too large)", a.M); return -3; }
    }
    const bool staged = a.geom && a.shs && (a.shs_rest || (per % 4) == 0) && smem <= 100 * 1024;
    if ((a.accumulate & 1) && a.geom && a.shs && a.dL_dshs && !staged) {
        set_error("preprocess backward: accumulate mode needs the staged SH path (M=%d)", a.M);
        return -3;
    }
    if (staged) {
        static PerDeviceOnce attr_done;
        if (attr_done.todo()) {
            OGS_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            attr_done.done();
        }
        preprocess_bwd_kernel<true><<<(a.P + PB - 1) / PB, PB, smem, s>>>(a);
    } else {
        preprocess_bwd_kernel<false><<<(a.P + PB - 1) / PB, PB, 0, s>>>(a);
    }
    return 0;
}

}  // namespace ogs
