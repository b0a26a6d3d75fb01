/*
 * oracle/raster_oracle.c -- CPU restatement of the tile-based differentiable Gaussian
 * rasterizer that OpenGaussian calls through `GaussianRasterizer` (package
 * `ashawkey_diff_gaussian_rasterization`, reference call sites
 * gaussian_renderer/__init__.py:55-70,104-163 and utils/sam_refinement_utils.py:347-403).
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (opengaussian_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED for binning and blending (PARTIALLY PINNED for preprocess: SH colours, 3D covariance and
 * camera conventions are checked against golden vectors produced by the reference's own Python code,
 * tests/golden/make_raster_golden.py): the rasterizer's CUDA source is a third-party dependency that is NOT
 * present under /root/reference (submodules/ashawkey-diff-gaussian-rasterization.zip is
 * listed in .MISSING_LARGE_BLOBS; only pin: the commented
 * `ashawkey-diff-gaussian-rasterization==0.0.0` at environment.yml:118).  The reference
 * holds no tests, golden vectors or fixtures for this path either.
 * What follows restates the published algorithm of ashawkey/diff-gaussian-rasterization
 * (a fork of graphdeco-inria/diff-gaussian-rasterization that additionally composites
 * depth and alpha), anchored on:
 *   - the call-site contract above (argument names, shapes, 4-tuple return),
 *   - matrix conventions of scene/cameras.py:71-78 + utils/graphics_utils.py:38-74
 *     (row-vector, i.e. transposed, 4x4 tensors; raw floats indexed column-major),
 *   - SH basis/constants of utils/sh_utils.py:26-112,
 *   - covariance packing (xx,xy,xz,yy,yz,zz) of utils/general_utils.py:64-73,
 *   - quaternion (r,x,y,z) -> R of utils/general_utils.py:78-99.
 *
 * Arithmetic contract (shared with the CUDA kernels so that integer artefacts are
 * bit-exact): every per-Gaussian quantity that feeds radii / tile rects / sort keys is
 * evaluated in IEEE fp32, expressions associated exactly as written here, with NO fused
 * multiply-add contraction (build with -ffp-contract=off; the CUDA side builds the
 * preprocess translation unit with --fmad=false).  float->int casts follow the CUDA
 * cvt.rzi.s32.f32 rule (NaN -> 0, saturating).  Pixel blending uses expf() here and
 * ex2.approx on the GPU: images agree to 1e-5 except where a skip/stop decision sits
 * within a relative margin of its threshold; such pixels are reported in `flags`.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OGS_TILE 16
/* The pixel loops are OpenMP-parallel when built with -fopenmp (bench.py's CPU baselines use all
 * host threads); per-Gaussian gradient sums then use atomic double adds. */
#ifdef _OPENMP
#include <omp.h>
#define OGS_ATOMIC _Pragma("omp atomic")
#else
#define OGS_ATOMIC
#endif
int ogs_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* CUDA-style float -> int32 conversion (round toward zero, saturating, NaN -> 0). */
static int32_t f2i_rz(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* m is the raw 16 floats of the row-vector (transposed) matrix tensor. */
static void xform4x3(const float* p, const float* m, float* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static void xform4x4(const float* p, const float* m, float* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* Sigma = (R diag(s)) (R diag(s))^T, quaternion q=(r,x,y,z) used as given (no re-normalisation). */
static void cov3d_from_scale_rot(const float* scale, float mod, const float* q, float* cov6) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[3][3];
    R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
    R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
    R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
    float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    float L[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) L[i][j] = s[j] * R[i][j];
    cov6[0] = L[0][0] * L[0][0] + L[0][1] * L[0][1] + L[0][2] * L[0][2];
    cov6[1] = L[0][0] * L[1][0] + L[0][1] * L[1][1] + L[0][2] * L[1][2];
    cov6[2] = L[0][0] * L[2][0] + L[0][1] * L[2][1] + L[0][2] * L[2][2];
    cov6[3] = L[1][0] * L[1][0] + L[1][1] * L[1][1] + L[1][2] * L[1][2];
    cov6[4] = L[1][0] * L[2][0] + L[1][1] * L[2][1] + L[1][2] * L[2][2];
    cov6[5] = L[2][0] * L[2][0] + L[2][1] * L[2][1] + L[2][2] * L[2][2];
}

/* EWA projection; returns (xx, xy, yy) with the 0.3 px low-pass added.  T0/T1 (rows of J*Rw) out. */
static void cov2d_project(const float* t_view, float fx, float fy, float tanx, float tany,
                          const float* c6, const float* v, float* cov3, float* T0, float* T1) {
    float tz = t_view[2];
    float limx = 1.3f * tanx, limy = 1.3f * tany;
    float txtz = t_view[0] / tz, tytz = t_view[1] / tz;
    float tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    float ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    float J00 = fx / tz, J02 = -(fx * tx) / (tz * tz);
    float J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
    for (int k = 0; k < 3; k++) {
        T0[k] = v[4 * k + 0] * J00 + v[4 * k + 2] * J02;
        T1[k] = v[4 * k + 1] * J11 + v[4 * k + 2] * J12;
    }
    float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
    float A0[3], A1[3];
    for (int k = 0; k < 3; k++) {
        A0[k] = T0[0] * V[0][k] + T0[1] * V[1][k] + T0[2] * V[2][k];
        A1[k] = T1[0] * V[0][k] + T1[1] * V[1][k] + T1[2] * V[2][k];
    }
    cov3[0] = A0[0] * T0[0] + A0[1] * T0[1] + A0[2] * T0[2];
    cov3[1] = A1[0] * T0[0] + A1[1] * T0[1] + A1[2] * T0[2];
    cov3[2] = A1[0] * T1[0] + A1[1] * T1[1] + A1[2] * T1[2];
    cov3[0] += 0.3f;
    cov3[2] += 0.3f;
}

static void tile_rect(float px, float py, int radius, int gx, int gy, int* rmin, int* rmax) {
    rmin[0] = imin(gx, imax(0, f2i_rz((px - (float)radius) / (float)OGS_TILE)));
    rmin[1] = imin(gy, imax(0, f2i_rz((py - (float)radius) / (float)OGS_TILE)));
    rmax[0] = imin(gx, imax(0, f2i_rz((px + (float)radius + (float)(OGS_TILE - 1)) / (float)OGS_TILE)));
    rmax[1] = imin(gy, imax(0, f2i_rz((py + (float)radius + (float)(OGS_TILE - 1)) / (float)OGS_TILE)));
}

/* SH (degree <= 3) -> RGB, +0.5, clamp at 0 with mask.  sh: [M][3] for this Gaussian. */
static void sh_to_rgb(int deg, const float* sh, const float* pos, const float* campos, float* rgb,
                      uint8_t* clamped) {
    float d[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
    float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    float x = d[0] / len, y = d[1] / len, z = d[2] / len;
    for (int c = 0; c < 3; c++) {
#define S(k) sh[(k) * 3 + c]
        float res = SH_C0 * S(0);
        if (deg > 0) {
            res = res - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                res = res + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) +
                      SH_C2[2] * (2.0f * zz - xx - yy) * S(6) + SH_C2[3] * xz * S(7) +
                      SH_C2[4] * (xx - yy) * S(8);
                if (deg > 2) {
                    res = res + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) +
                          SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                          SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                          SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) +
                          SH_C3[5] * z * (xx - yy) * S(14) + SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                }
            }
        }
#undef S
        res += 0.5f;
        clamped[c] = (uint8_t)(res < 0.0f);
        rgb[c] = fmaxf(res, 0.0f);
    }
}

/* ------------------------------------------------------------------------------------------
 * preprocess (per Gaussian): cull, cov3D, cov2D, conic, radius, pixel centre, tile rect, colour.
 * Outputs are zero-initialised for culled Gaussians (radii = 0, tiles_touched = 0).
 * rgb may be NULL when shs is NULL (precomputed colours are consumed directly by the blend).
 * ---------------------------------------------------------------------------------------- */
int ogs_oracle_preprocess(int P, int sh_degree, int M, const float* means3D, const float* scales,
                          const float* rotations, const float* cov3D_precomp,
                          const float* opacities, const float* shs, float scale_modifier,
                          const float* view, const float* proj, const float* campos, int W, int H,
                          float tanfovx, float tanfovy, int32_t* radii, float* xy, float* depth,
                          float* cov3D, float* conic_opacity, float* rgb, uint8_t* clamped,
                          uint32_t* tiles_touched) {
    const float fx = (float)W / (2.0f * tanfovx);
    const float fy = (float)H / (2.0f * tanfovy);
    const int gx = (W + OGS_TILE - 1) / OGS_TILE, gy = (H + OGS_TILE - 1) / OGS_TILE;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        xy[2 * i] = xy[2 * i + 1] = 0.f;
        depth[i] = 0.f;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = 0.f;
        for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0.f;
        if (rgb) for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.f; clamped[3 * i + k] = 0; }

        const float* p = means3D + 3 * i;
        float pv[3], ph[4];
        xform4x3(p, view, pv);
        if (pv[2] <= 0.2f) continue; /* near cull */
        xform4x4(p, proj, ph);
        float pw = 1.0f / (ph[3] + 0.0000001f);
        float ndc_x = ph[0] * pw, ndc_y = ph[1] * pw;

        float c6[6];
        if (cov3D_precomp) memcpy(c6, cov3D_precomp + 6 * i, sizeof c6);
        else cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rotations + 4 * i, c6);

        float cov[3], T0[3], T1[3];
        cov2d_project(pv, fx, fy, tanfovx, tanfovy, c6, view, cov, T0, T1);
        float det = cov[0] * cov[2] - cov[1] * cov[1];
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conic[3] = {cov[2] * det_inv, -cov[1] * det_inv, cov[0] * det_inv};
        float mid = 0.5f * (cov[0] + cov[2]);
        float root = sqrtf(fmaxf(0.1f, mid * mid - det));
        float lam1 = mid + root, lam2 = mid - root;
        float rad_f = ceilf(3.f * sqrtf(fmaxf(lam1, lam2)));
        int rad = f2i_rz(rad_f);
        float px = ((ndc_x + 1.0f) * (float)W - 1.0f) * 0.5f;
        float py = ((ndc_y + 1.0f) * (float)H - 1.0f) * 0.5f;
        int rmin[2], rmax[2];
        tile_rect(px, py, rad, gx, gy, rmin, rmax);
        if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;

        if (shs) sh_to_rgb(sh_degree, shs + (size_t)i * M * 3, p, campos, rgb + 3 * i, clamped + 3 * i);
        depth[i] = pv[2];
        radii[i] = rad;
        xy[2 * i] = px;
        xy[2 * i + 1] = py;
        memcpy(cov3D + 6 * i, c6, sizeof c6);
        conic_opacity[4 * i + 0] = conic[0];
        conic_opacity[4 * i + 1] = conic[1];
        conic_opacity[4 * i + 2] = conic[2];
        conic_opacity[4 * i + 3] = opacities[i];
        tiles_touched[i] = (uint32_t)((rmax[1] - rmin[1]) * (rmax[0] - rmin[0]));
    }
    return 0;
}

/* markVisible: the near-plane frustum test alone. */
int ogs_oracle_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
    for (int i = 0; i < P; i++) {
        float pv[3];
        xform4x3(means3D + 3 * i, view, pv);
        present[i] = (uint8_t)(pv[2] > 0.2f);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * binning: inclusive scan, (tile<<32 | depth bits, idx) duplicates, stable LSD radix sort on
 * the low 32+tile_bits bits, per-tile [first, last+1) ranges.
 * ---------------------------------------------------------------------------------------- */
int64_t ogs_oracle_scan(int P, const uint32_t* tiles_touched, uint32_t* offsets) {
    uint64_t acc = 0;
    for (int i = 0; i < P; i++) { acc += tiles_touched[i]; offsets[i] = (uint32_t)acc; }
    return (int64_t)acc;
}

static void radix_sort_pairs(uint64_t* k, uint32_t* v, size_t n, int bits) {
    uint64_t* k2 = (uint64_t*)malloc(n * sizeof *k2);
    uint32_t* v2 = (uint32_t*)malloc(n * sizeof *v2);
    uint64_t *ka = k, *kb = k2; uint32_t *va = v, *vb = v2;
    for (int sh = 0; sh < bits; sh += 8) {
        size_t cnt[257]; memset(cnt, 0, sizeof cnt);
        for (size_t i = 0; i < n; i++) cnt[((ka[i] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < n; i++) { size_t d = (ka[i] >> sh) & 255; size_t o = cnt[d]++; kb[o] = ka[i]; vb[o] = va[i]; }
        uint64_t* tk = ka; ka = kb; kb = tk; uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != k) { memcpy(k, ka, n * sizeof *k); memcpy(v, va, n * sizeof *v); }
    free(k2); free(v2);
}

int ogs_oracle_bin(int P, int W, int H, const float* xy, const float* depth, const int32_t* radii,
                   const uint32_t* offsets, int64_t N, uint64_t* keys, uint32_t* values,
                   uint32_t* ranges /* [tiles][2] */) {
    const int gx = (W + OGS_TILE - 1) / OGS_TILE, gy = (H + OGS_TILE - 1) / OGS_TILE;
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        size_t off = (i == 0) ? 0 : offsets[i - 1];
        int rmin[2], rmax[2];
        tile_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, rmin, rmax);
        for (int y = rmin[1]; y < rmax[1]; y++)
            for (int x = rmin[0]; x < rmax[0]; x++) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                key <<= 32;
                key |= f2u(depth[i]);
                keys[off] = key;
                values[off] = (uint32_t)i;
                off++;
            }
    }
    int tiles = gx * gy, tile_bits = 0;
    while ((1 << tile_bits) < tiles) tile_bits++; /* >= bits needed; sort order is unaffected */
    radix_sort_pairs(keys, values, (size_t)N, 32 + tile_bits + 1);
    memset(ranges, 0, (size_t)tiles * 2 * sizeof(uint32_t));
    for (int64_t i = 0; i < N; i++) {
        uint32_t t = (uint32_t)(keys[i] >> 32);
        if (i == 0) ranges[2 * t] = 0;
        else {
            uint32_t pt = (uint32_t)(keys[i - 1] >> 32);
            if (t != pt) { ranges[2 * pt + 1] = (uint32_t)i; ranges[2 * t] = (uint32_t)i; }
        }
        if (i == N - 1) ranges[2 * t + 1] = (uint32_t)N;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * forward blend.  colors: [P][C] (RGB from SH or precomputed, optionally followed by extra
 * per-Gaussian feature channels -- OpenGaussian's ins_feat); bg: [C].
 * out_color [C][H][W], out_depth [H][W] (sum depth*alpha*T, no bg, not normalised),
 * out_alpha [H][W] (sum alpha*T), final_T, n_contrib; flags[pix] != 0 where some skip/stop
 * decision lay within `margin` (relative) of its threshold.
 * ---------------------------------------------------------------------------------------- */
int ogs_oracle_blend_forward(int W, int H, int C, const uint32_t* ranges, const uint32_t* point_list,
                             const float* xy, const float* conic_opacity, const float* colors,
                             const float* depth, const float* bg, float margin, float* out_color,
                             float* out_depth, float* out_alpha, float* final_T,
                             uint32_t* n_contrib, uint8_t* flags) {
    const int gx = (W + OGS_TILE - 1) / OGS_TILE;
#pragma omp parallel
    {
    float* acc = (float*)malloc(sizeof(float) * (size_t)C);
#pragma omp for schedule(dynamic, 4)
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            int tile = (py / OGS_TILE) * gx + (px / OGS_TILE);
            uint32_t lo = ranges[2 * tile], hi = ranges[2 * tile + 1];
            float T = 1.0f, D = 0.f, wsum = 0.f;
            uint32_t contributor = 0, last = 0;
            uint8_t flag = 0;
            for (int c = 0; c < C; c++) acc[c] = 0.f;
            for (uint32_t k = lo; k < hi; k++) {
                contributor++;
                uint32_t g = point_list[k];
                float dx = xy[2 * g] - (float)px, dy = xy[2 * g + 1] - (float)py;
                const float* co = conic_opacity + 4 * g;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (fabsf(power) < 1e-6f) flag = 1;
                if (power > 0.0f) continue;
                float alpha = fminf(0.99f, co[3] * expf(power));
                if (fabsf(alpha - 1.0f / 255.0f) < margin * (1.0f / 255.0f)) flag = 1;
                if (alpha < 1.0f / 255.0f) continue;
                float test_T = T * (1 - alpha);
                if (fabsf(test_T - 0.0001f) < 10.f * margin * 0.0001f) flag = 1;
                if (test_T < 0.0001f) break; /* this Gaussian is NOT applied */
                float w = alpha * T;
                for (int c = 0; c < C; c++) acc[c] += colors[(size_t)g * C + c] * w;
                wsum += w;
                D += depth[g] * w;
                T = test_T;
                last = contributor;
            }
            size_t pix = (size_t)py * W + px;
            final_T[pix] = T;
            n_contrib[pix] = last;
            for (int c = 0; c < C; c++) out_color[(size_t)c * H * W + pix] = acc[c] + T * bg[c];
            out_depth[pix] = D;
            out_alpha[pix] = wsum;
            if (flags) flags[pix] = flag;
        }
    free(acc);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * backward blend: back-to-front from n_contrib, alpha recomputed, T recovered by division.
 * Accumulates in double.  dL_dmean2D [P][2] is in NDC-scaled units (x * 0.5 W, y * 0.5 H);
 * dL_dconic [P][3] holds (d/dA, HALF of d/dB, d/dC) -- the half convention matches the
 * chain rule used in the preprocess backward below.  The 0.99 clamp passes gradient through.
 * ---------------------------------------------------------------------------------------- */
int ogs_oracle_blend_backward(int P, int W, int H, int C, const uint32_t* ranges,
                              const uint32_t* point_list, const float* xy,
                              const float* conic_opacity, const float* colors, const float* depth,
                              const float* bg, const float* final_T, const uint32_t* n_contrib,
                              const float* dL_dpix /* [C][H][W] */, const float* dL_ddepth_pix,
                              const float* dL_dalpha_pix, double* dL_dmean2D, double* dL_dconic,
                              double* dL_dopacity, double* dL_dcolors, double* dL_ddepth) {
    const int gx = (W + OGS_TILE - 1) / OGS_TILE;
    memset(dL_dmean2D, 0, sizeof(double) * 2 * (size_t)P);
    memset(dL_dconic, 0, sizeof(double) * 3 * (size_t)P);
    memset(dL_dopacity, 0, sizeof(double) * (size_t)P);
    memset(dL_dcolors, 0, sizeof(double) * (size_t)C * P);
    memset(dL_ddepth, 0, sizeof(double) * (size_t)P);
    const double ddelx_dx = 0.5 * W, ddely_dy = 0.5 * H;
#pragma omp parallel
    {
    double* accum_rec = (double*)malloc(sizeof(double) * C);
    double* last_color = (double*)malloc(sizeof(double) * C);
    double* g = (double*)malloc(sizeof(double) * C);
#pragma omp for schedule(dynamic, 4)
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            size_t pix = (size_t)py * W + px;
            int tile = (py / OGS_TILE) * gx + (px / OGS_TILE);
            uint32_t lo = ranges[2 * tile];
            const double T_final = final_T[pix];
            double T = T_final;
            uint32_t last = n_contrib[pix];
            double bg_dot = 0;
            for (int c = 0; c < C; c++) { g[c] = dL_dpix[(size_t)c * H * W + pix]; accum_rec[c] = 0; last_color[c] = 0; bg_dot += (double)bg[c] * g[c]; }
            const double gd = dL_ddepth_pix ? dL_ddepth_pix[pix] : 0.0;
            const double ga = dL_dalpha_pix ? dL_dalpha_pix[pix] : 0.0;
            double accum_depth = 0, last_depth = 0, accum_alpha = 0, last_alpha = 0;
            for (uint32_t pos = last; pos-- > 0;) {
                uint32_t gi = point_list[lo + pos];
                float dx = xy[2 * gi] - (float)px, dy = xy[2 * gi + 1] - (float)py;
                const float* co = conic_opacity + 4 * gi;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                float Gf = expf(power);
                float alpha_f = fminf(0.99f, co[3] * Gf);
                if (alpha_f < 1.0f / 255.0f) continue;
                const double G = Gf, alpha = alpha_f;
                T = T / (1.0 - alpha);
                const double w = alpha * T;
                double dL_dalpha = 0;
                for (int c = 0; c < C; c++) {
                    double col = colors[(size_t)gi * C + c];
                    accum_rec[c] = last_alpha * last_color[c] + (1.0 - last_alpha) * accum_rec[c];
                    last_color[c] = col;
                    dL_dalpha += (col - accum_rec[c]) * g[c];
                    OGS_ATOMIC
                    dL_dcolors[(size_t)gi * C + c] += w * g[c];
                }
                double cd = depth[gi];
                accum_depth = last_alpha * last_depth + (1.0 - last_alpha) * accum_depth;
                last_depth = cd;
                dL_dalpha += (cd - accum_depth) * gd;
                OGS_ATOMIC
                dL_ddepth[gi] += w * gd;
                accum_alpha = last_alpha * 1.0 + (1.0 - last_alpha) * accum_alpha;
                dL_dalpha += (1.0 - accum_alpha) * ga;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.0 - alpha)) * bg_dot;

                const double dL_dG = (double)co[3] * dL_dalpha;
                const double gdx = G * dx, gdy = G * dy;
                const double dG_ddelx = -gdx * co[0] - gdy * co[1];
                const double dG_ddely = -gdy * co[2] - gdx * co[1];
                OGS_ATOMIC
                dL_dmean2D[2 * gi + 0] += dL_dG * dG_ddelx * ddelx_dx;
                OGS_ATOMIC
                dL_dmean2D[2 * gi + 1] += dL_dG * dG_ddely * ddely_dy;
                OGS_ATOMIC
                dL_dconic[3 * gi + 0] += -0.5 * gdx * dx * dL_dG;
                OGS_ATOMIC
                dL_dconic[3 * gi + 1] += -0.5 * gdx * dy * dL_dG;
                OGS_ATOMIC
                dL_dconic[3 * gi + 2] += -0.5 * gdy * dy * dL_dG;
                OGS_ATOMIC
                dL_dopacity[gi] += G * dL_dalpha;
            }
        }
    free(accum_rec); free(last_color); free(g);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * preprocess backward (per Gaussian with radii > 0): conic -> cov2D -> (cov3D, mean);
 * mean2D -> mean3D through the projective division; depth -> mean3D; SH backward (incl.
 * direction normalisation and the clamp mask); cov3D -> (scale, quaternion).  double math.
 * dL_dcolors_in: [P][C] from the blend; its first 3 channels drive the SH backward when shs
 * is given.  Outputs (all [P][*], zero for radii <= 0): dL_dmeans3D[3], dL_dscales[3],
 * dL_drotations[4], dL_dcov3D[6], dL_dshs[M][3].
 * ---------------------------------------------------------------------------------------- */
int ogs_oracle_preprocess_backward(int P, int sh_degree, int M, int C, const float* means3D,
                                   const float* scales, const float* rotations,
                                   const float* cov3D_precomp, const float* shs,
                                   float scale_modifier, const float* view, const float* proj,
                                   const float* campos, int W, int H, float tanfovx, float tanfovy,
                                   const int32_t* radii, const float* cov3D,
                                   const uint8_t* clamped, const double* dL_dmean2D,
                                   const double* dL_dconic, const double* dL_dcolors_in,
                                   const double* dL_ddepth, double* dL_dmeans3D,
                                   double* dL_dscales, double* dL_drotations, double* dL_dcov3D,
                                   double* dL_dshs) {
    const double fx = (double)((float)W / (2.0f * tanfovx));
    const double fy = (double)((float)H / (2.0f * tanfovy));
    memset(dL_dmeans3D, 0, sizeof(double) * 3 * (size_t)P);
    if (dL_dscales) memset(dL_dscales, 0, sizeof(double) * 3 * (size_t)P);
    if (dL_drotations) memset(dL_drotations, 0, sizeof(double) * 4 * (size_t)P);
    memset(dL_dcov3D, 0, sizeof(double) * 6 * (size_t)P);
    if (dL_dshs) memset(dL_dshs, 0, sizeof(double) * 3 * (size_t)M * P);
    const float* v = view;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue;
        const float* m = means3D + 3 * i;
        const float* c6 = cov3D + 6 * i;
        double* dmean = dL_dmeans3D + 3 * i;
        /* ---- cov2D backward ---- */
        float tvf[3];
        xform4x3(m, view, tvf);
        double t[3] = {tvf[0], tvf[1], tvf[2]};
        const double limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
        const double txtz = t[0] / t[2], tytz = t[1] / t[2];
        t[0] = fmin(limx, fmax(-limx, txtz)) * t[2];
        t[1] = fmin(limy, fmax(-limy, tytz)) * t[2];
        const double x_grad_mul = (txtz < -limx || txtz > limx) ? 0 : 1;
        const double y_grad_mul = (tytz < -limy || tytz > limy) ? 0 : 1;
        const double J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
        const double J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
        double T0[3], T1[3];
        for (int k = 0; k < 3; k++) {
            T0[k] = v[4 * k + 0] * J00 + v[4 * k + 2] * J02;
            T1[k] = v[4 * k + 1] * J11 + v[4 * k + 2] * J12;
        }
        const double V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
        double VT0[3], VT1[3];
        for (int k = 0; k < 3; k++) {
            VT0[k] = V[k][0] * T0[0] + V[k][1] * T0[1] + V[k][2] * T0[2];
            VT1[k] = V[k][0] * T1[0] + V[k][1] * T1[1] + V[k][2] * T1[2];
        }
        const double a = T0[0] * VT0[0] + T0[1] * VT0[1] + T0[2] * VT0[2] + 0.3f;
        const double b = T0[0] * VT1[0] + T0[1] * VT1[1] + T0[2] * VT1[2];
        const double c = T1[0] * VT1[0] + T1[1] * VT1[1] + T1[2] * VT1[2] + 0.3f;
        const double denom = a * c - b * b;
        const double dLA = dL_dconic[3 * i], dLBh = dL_dconic[3 * i + 1], dLC = dL_dconic[3 * i + 2];
        const double d2inv = 1.0 / (denom * denom + 0.0000001);
        double dL_da = 0, dL_db = 0, dL_dc = 0;
        if (d2inv != 0) {
            dL_da = d2inv * (-c * c * dLA + 2 * b * c * dLBh + (denom - a * c) * dLC);
            dL_dc = d2inv * (-a * a * dLC + 2 * a * b * dLBh + (denom - a * c) * dLA);
            dL_db = d2inv * 2 * (b * c * dLA - (denom + 2 * b * b) * dLBh + a * b * dLC);
            double* dc6 = dL_dcov3D + 6 * i;
            dc6[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
            dc6[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
            dc6[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
            dc6[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
            dc6[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
            dc6[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
        }
        double dT0[3], dT1[3];
        for (int k = 0; k < 3; k++) {
            dT0[k] = 2 * VT0[k] * dL_da + VT1[k] * dL_db;
            dT1[k] = 2 * VT1[k] * dL_dc + VT0[k] * dL_db;
        }
        double dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int k = 0; k < 3; k++) {
            dJ00 += v[4 * k + 0] * dT0[k]; dJ02 += v[4 * k + 2] * dT0[k];
            dJ11 += v[4 * k + 1] * dT1[k]; dJ12 += v[4 * k + 2] * dT1[k];
        }
        const double tz = 1.0 / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const double dtx = x_grad_mul * -fx * tz2 * dJ02;
        const double dty = y_grad_mul * -fy * tz2 * dJ12;
        const double dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
        dmean[0] = v[0] * dtx + v[1] * dty + v[2] * dtz;
        dmean[1] = v[4] * dtx + v[5] * dty + v[6] * dtz;
        dmean[2] = v[8] * dtx + v[9] * dty + v[10] * dtz;
        /* ---- mean2D (NDC-scaled) -> mean3D ---- */
        float mhf[4];
        xform4x4(m, proj, mhf);
        const double m_w = 1.0 / ((double)mhf[3] + 0.0000001);
        const double mul1 = (double)mhf[0] * m_w * m_w, mul2 = (double)mhf[1] * m_w * m_w;
        const double gx2 = dL_dmean2D[2 * i], gy2 = dL_dmean2D[2 * i + 1];
        dmean[0] += (proj[0] * m_w - proj[3] * mul1) * gx2 + (proj[1] * m_w - proj[3] * mul2) * gy2;
        dmean[1] += (proj[4] * m_w - proj[7] * mul1) * gx2 + (proj[5] * m_w - proj[7] * mul2) * gy2;
        dmean[2] += (proj[8] * m_w - proj[11] * mul1) * gx2 + (proj[9] * m_w - proj[11] * mul2) * gy2;
        /* ---- depth = p_view.z ---- */
        dmean[0] += v[2] * dL_ddepth[i];
        dmean[1] += v[6] * dL_ddepth[i];
        dmean[2] += v[10] * dL_ddepth[i];
        /* ---- SH backward ---- */
        if (shs) {
            const float* sh = shs + (size_t)i * M * 3;
            double* dsh = dL_dshs + (size_t)i * M * 3;
            double d0[3] = {(double)m[0] - campos[0], (double)m[1] - campos[1], (double)m[2] - campos[2]};
            /* direction as the forward computed it (fp32 difference) */
            d0[0] = (double)(m[0] - campos[0]); d0[1] = (double)(m[1] - campos[1]); d0[2] = (double)(m[2] - campos[2]);
            const double len = sqrt(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]);
            const double x = d0[0] / len, y = d0[1] / len, z = d0[2] / len;
            double dRGB[3];
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = clamped[3 * i + ch] ? 0.0 : dL_dcolors_in[(size_t)i * C + ch];
            double ddir[3] = {0, 0, 0};
            double basis[16], dbx[16], dby[16], dbz[16];
            memset(basis, 0, sizeof basis); memset(dbx, 0, sizeof dbx); memset(dby, 0, sizeof dby); memset(dbz, 0, sizeof dbz);
            basis[0] = SH_C0;
            if (sh_degree > 0) {
                basis[1] = -SH_C1 * y; dby[1] = -SH_C1;
                basis[2] = SH_C1 * z;  dbz[2] = SH_C1;
                basis[3] = -SH_C1 * x; dbx[3] = -SH_C1;
                if (sh_degree > 1) {
                    const double xx = x * x, yy = y * y, zz = z * z, xy_ = x * y, yz = y * z, xz = x * z;
                    basis[4] = SH_C2[0] * xy_; dbx[4] = SH_C2[0] * y; dby[4] = SH_C2[0] * x;
                    basis[5] = SH_C2[1] * yz;  dby[5] = SH_C2[1] * z; dbz[5] = SH_C2[1] * y;
                    basis[6] = SH_C2[2] * (2.0 * zz - xx - yy);
                    dbx[6] = SH_C2[2] * -2.0 * x; dby[6] = SH_C2[2] * -2.0 * y; dbz[6] = SH_C2[2] * 4.0 * z;
                    basis[7] = SH_C2[3] * xz; dbx[7] = SH_C2[3] * z; dbz[7] = SH_C2[3] * x;
                    basis[8] = SH_C2[4] * (xx - yy); dbx[8] = SH_C2[4] * 2.0 * x; dby[8] = SH_C2[4] * -2.0 * y;
                    if (sh_degree > 2) {
                        basis[9] = SH_C3[0] * y * (3.0 * xx - yy);
                        dbx[9] = SH_C3[0] * 6.0 * xy_; dby[9] = SH_C3[0] * (3.0 * xx - 3.0 * yy);
                        basis[10] = SH_C3[1] * xy_ * z;
                        dbx[10] = SH_C3[1] * yz; dby[10] = SH_C3[1] * xz; dbz[10] = SH_C3[1] * xy_;
                        basis[11] = SH_C3[2] * y * (4.0 * zz - xx - yy);
                        dbx[11] = SH_C3[2] * -2.0 * xy_; dby[11] = SH_C3[2] * (4.0 * zz - xx - 3.0 * yy); dbz[11] = SH_C3[2] * 8.0 * yz;
                        basis[12] = SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy);
                        dbx[12] = SH_C3[3] * -6.0 * xz; dby[12] = SH_C3[3] * -6.0 * yz; dbz[12] = SH_C3[3] * (6.0 * zz - 3.0 * xx - 3.0 * yy);
                        basis[13] = SH_C3[4] * x * (4.0 * zz - xx - yy);
                        dbx[13] = SH_C3[4] * (4.0 * zz - 3.0 * xx - yy); dby[13] = SH_C3[4] * -2.0 * xy_; dbz[13] = SH_C3[4] * 8.0 * xz;
                        basis[14] = SH_C3[5] * z * (xx - yy);
                        dbx[14] = SH_C3[5] * 2.0 * xz; dby[14] = SH_C3[5] * -2.0 * yz; dbz[14] = SH_C3[5] * (xx - yy);
                        basis[15] = SH_C3[6] * x * (xx - 3.0 * yy);
                        dbx[15] = SH_C3[6] * (3.0 * xx - 3.0 * yy); dby[15] = SH_C3[6] * -6.0 * xy_;
                    }
                }
            }
            const int ncoef = (sh_degree + 1) * (sh_degree + 1);
            for (int k = 0; k < ncoef; k++)
                for (int ch = 0; ch < 3; ch++) {
                    dsh[k * 3 + ch] = basis[k] * dRGB[ch];
                    const double s = (double)sh[k * 3 + ch] * dRGB[ch];
                    ddir[0] += dbx[k] * s; ddir[1] += dby[k] * s; ddir[2] += dbz[k] * s;
                }
            /* d(normalize(d0))/d(d0) applied to ddir */
            const double sum2 = len * len, invsum32 = 1.0 / (sum2 * len);
            dmean[0] += ((sum2 - d0[0] * d0[0]) * ddir[0] - d0[1] * d0[0] * ddir[1] - d0[2] * d0[0] * ddir[2]) * invsum32;
            dmean[1] += (-d0[0] * d0[1] * ddir[0] + (sum2 - d0[1] * d0[1]) * ddir[1] - d0[2] * d0[1] * ddir[2]) * invsum32;
            dmean[2] += (-d0[0] * d0[2] * ddir[0] - d0[1] * d0[2] * ddir[1] + (sum2 - d0[2] * d0[2]) * ddir[2]) * invsum32;
        }
        /* ---- cov3D -> scale, quaternion ---- */
        if (!cov3D_precomp && scales) {
            const float* q = rotations + 4 * i;
            const double r = q[0], x = q[1], y = q[2], z = q[3];
            double R[3][3];
            R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - r * z); R[0][2] = 2 * (x * z + r * y);
            R[1][0] = 2 * (x * y + r * z); R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - r * x);
            R[2][0] = 2 * (x * z - r * y); R[2][1] = 2 * (y * z + r * x); R[2][2] = 1 - 2 * (x * x + y * y);
            const double s[3] = {(double)scale_modifier * scales[3 * i], (double)scale_modifier * scales[3 * i + 1], (double)scale_modifier * scales[3 * i + 2]};
            const double* dc6 = dL_dcov3D + 6 * i;
            const double Gs[3][3] = {{dc6[0], 0.5 * dc6[1], 0.5 * dc6[2]}, {0.5 * dc6[1], dc6[3], 0.5 * dc6[4]}, {0.5 * dc6[2], 0.5 * dc6[4], dc6[5]}};
            /* Sigma = L L^T, L = R diag(s);  dL/dL = 2 Gs L */
            double dLm[3][3], dR[3][3];
            for (int a_ = 0; a_ < 3; a_++) for (int b_ = 0; b_ < 3; b_++) {
                double acc = 0;
                for (int k = 0; k < 3; k++) acc += Gs[a_][k] * R[k][b_] * s[b_];
                dLm[a_][b_] = 2 * acc;
            }
            for (int j = 0; j < 3; j++) {
                double ds = 0;
                for (int a_ = 0; a_ < 3; a_++) { ds += R[a_][j] * dLm[a_][j]; dR[a_][j] = dLm[a_][j] * s[j]; }
                dL_dscales[3 * i + j] = ds * (double)scale_modifier;
            }
            double* dq = dL_drotations + 4 * i;
            dq[0] = 2 * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
            dq[1] = 2 * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2 * x * dR[1][1] - r * dR[1][2] + z * dR[2][0] + r * dR[2][1] - 2 * x * dR[2][2]);
            dq[2] = 2 * (-2 * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2] - r * dR[2][0] + z * dR[2][1] - 2 * y * dR[2][2]);
            dq[3] = 2 * (-2 * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - 2 * z * dR[1][1] + y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
        }
    }
    return 0;
}
