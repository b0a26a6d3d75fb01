// upstream_structure.cu -- GPU COMPARATOR, not part of the product and never loaded by it.
//
// SURVEY.md section 8d, last row: the reference's CUDA rasterizer (`ashawkey/diff-gaussian-rasterization`) is an
// un-vendored dependency whose source is absent from /root/reference, so "x times the reference CUDA rasterizer"
// cannot be measured against the real thing.  This file is a from-the-specification restatement of that
// package's KERNEL STRUCTURE ("Rasterizer specification to restate"), labelled as such:
//   * binning: library inclusive scan, one (tile << 32 | depth) 64-bit key per (Gaussian, tile) duplicate, ONE
//     library radix sort over 32 + log2(tiles) bits (cub::DeviceRadixSort), tile ranges from key boundaries;
//   * forward blend: one thread per pixel, one 16x16 CTA per tile, cooperative 256-entry fetches into shared
//     memory, __syncthreads_count early exit, expf, 3 colour channels + depth + alpha;
//   * backward blend: same tiling back to front, and every contributing (pixel, Gaussian) pair issues its own
//     global atomicAdds (3 colour + 1 depth + 6 geometry = 10 per pair) -- no warp-level pre-reduction.
// Preprocess forward / backward are the product's own kernels (per-Gaussian threads in both designs; ~15 % of a
// frame), which is why the geometry atomics accumulate the same six moments the product's preprocess-backward
// consumes (common.cuh, BlendBwdArgs) -- the same COUNT of atomics per pair as the upstream layout
// (mean2D 2 + conic 3 + opacity 1).
// Same C signatures and state layout as ogs_raster_forward / ogs_raster_backward, so one driver script can time
// both (scripts/upstream_structure_bench.py).  Built by baseline/build_comparator.py into
// baseline/_build/libogs_upstream_structure.so together with the product's object files.
#include <cub/cub.cuh>
#include <string.h>

#include "../opengaussian_b200/csrc/common.cuh"

using namespace ogs;

namespace {

#define UPS_BLOCK 256

__device__ __forceinline__ void ups_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1, int& y1) {
    const float r = (float)radius;
    x0 = min(gx, max(0, (int)((px - r) / 16.0f)));
    y0 = min(gy, max(0, (int)((py - r) / 16.0f)));
    x1 = min(gx, max(0, (int)((px + r + 15.0f) / 16.0f)));
    y1 = min(gy, max(0, (int)((py + r + 15.0f) / 16.0f)));
}

__global__ void duplicate_with_keys(int P, const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                                    const uint32_t* __restrict__ tiles, const uint32_t* __restrict__ offsets,
                                    uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int gx, int gy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || tiles[i] == 0u) return;
    uint32_t off = (i == 0) ? 0u : offsets[i - 1];
    const float4 r0 = rec0[i], r1 = rec1[i];
    int x0, y0, x1, y1;
    ups_rect(r0.x, r0.y, __float_as_int(r1.w), gx, gy, x0, y0, x1, y1);
    const uint64_t depth_bits = (uint64_t)__float_as_uint(r1.z);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | depth_bits;
            vals[off] = (uint32_t)i;
            off++;
        }
}

__global__ void identify_tile_ranges(int64_t N, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N) return;
    const uint32_t tile = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) {
        ranges[tile].x = 0u;
    } else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (prev != tile) {
            ranges[prev].y = (uint32_t)idx;
            ranges[tile].x = (uint32_t)idx;
        }
    }
    if (idx == N - 1) ranges[tile].y = (uint32_t)N;
}

__global__ void __launch_bounds__(UPS_BLOCK) render_forward(int W, int H, int gx, const uint2* __restrict__ ranges,
                                                            const uint32_t* __restrict__ point_list,
                                                            const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                                                            const float* __restrict__ colors, const float* __restrict__ bg,
                                                            float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                                                            float* __restrict__ out_color, float* __restrict__ out_depth,
                                                            float* __restrict__ out_alpha) {
    __shared__ uint32_t s_id[UPS_BLOCK];
    __shared__ float2 s_xy[UPS_BLOCK];
    __shared__ float4 s_co[UPS_BLOCK];
    __shared__ float s_depth[UPS_BLOCK];
    const int rank = threadIdx.y * 16 + threadIdx.x;
    const int px = blockIdx.x * 16 + threadIdx.x, py = blockIdx.y * 16 + threadIdx.y;
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const size_t HW = (size_t)W * H, pix = (size_t)py * W + px;
    const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = (int)((range.y - range.x + UPS_BLOCK - 1) / UPS_BLOCK);
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    float T = 1.0f, C[3] = {0.f, 0.f, 0.f}, D = 0.f;
    uint32_t contributor = 0, last = 0;
    for (int r = 0; r < rounds; r++, todo -= UPS_BLOCK) {
        if (__syncthreads_count(done) == UPS_BLOCK) break;
        const uint32_t progress = (uint32_t)r * UPS_BLOCK + rank;
        if (range.x + progress < range.y) {
            const uint32_t id = point_list[range.x + progress];
            const float4 r0 = rec0[id], r1 = rec1[id];
            s_id[rank] = id;
            s_xy[rank] = make_float2(r0.x, r0.y);
            s_co[rank] = make_float4(r0.z, r0.w, r1.x, r1.y);
            s_depth[rank] = r1.z;
        }
        __syncthreads();
        for (int j = 0; !done && j < min(UPS_BLOCK, todo); j++) {
            contributor++;
            const float2 xy = s_xy[j];
            const float dx = xy.x - fx, dy = xy.y - fy;
            const float4 co = s_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, co.w * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float w = alpha * T;
            const float* c = colors + 3 * (size_t)s_id[j];
            C[0] += c[0] * w; C[1] += c[1] * w; C[2] += c[2] * w;
            D += s_depth[j] * w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        final_T[pix] = T;
        n_contrib[pix] = last;
        for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = C[ch] + T * bg[ch];
        out_depth[pix] = D;
        out_alpha[pix] = 1.0f - T;
    }
}

__global__ void __launch_bounds__(UPS_BLOCK) render_backward(int W, int H, int gx, const uint2* __restrict__ ranges,
                                                             const uint32_t* __restrict__ point_list,
                                                             const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                                                             const float* __restrict__ colors, const float* __restrict__ bg,
                                                             const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                                                             const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                                                             const float* __restrict__ dL_dalpha_img, float* __restrict__ acc,
                                                             int stride) {
    __shared__ uint32_t s_id[UPS_BLOCK];
    __shared__ float2 s_xy[UPS_BLOCK];
    __shared__ float4 s_co[UPS_BLOCK];
    __shared__ float s_depth[UPS_BLOCK];
    __shared__ float s_col[3 * UPS_BLOCK];
    const int rank = threadIdx.y * 16 + threadIdx.x;
    const int px = blockIdx.x * 16 + threadIdx.x, py = blockIdx.y * 16 + threadIdx.y;
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const size_t HW = (size_t)W * H, pix = (size_t)py * W + px;
    const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = (int)((range.y - range.x + UPS_BLOCK - 1) / UPS_BLOCK);
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    uint32_t contributor = (uint32_t)todo;
    const uint32_t last_contributor = inside ? n_contrib[pix] : 0u;
    float g[3] = {0.f, 0.f, 0.f}, g_depth = 0.f, g_alpha = 0.f;
    if (inside) {
        for (int ch = 0; ch < 3; ch++) g[ch] = dL_dcolor[ch * HW + pix];
        if (dL_ddepth) g_depth = dL_ddepth[pix];
        if (dL_dalpha_img) g_alpha = dL_dalpha_img[pix];
    }
    const float bg_dot = bg[0] * g[0] + bg[1] * g[1] + bg[2] * g[2];
    float rec_c[3] = {0.f, 0.f, 0.f}, rec_d = 0.f, rec_a = 0.f;          // "colour behind" recursions
    float last_alpha = 0.f, last_c[3] = {0.f, 0.f, 0.f}, last_d = 0.f;
    for (int r = 0; r < rounds; r++, todo -= UPS_BLOCK) {
        __syncthreads();
        const uint32_t progress = (uint32_t)r * UPS_BLOCK + rank;
        if (range.x + progress < range.y) {
            const uint32_t id = point_list[range.y - progress - 1];         // back to front
            const float4 r0 = rec0[id], r1 = rec1[id];
            s_id[rank] = id;
            s_xy[rank] = make_float2(r0.x, r0.y);
            s_co[rank] = make_float4(r0.z, r0.w, r1.x, r1.y);
            s_depth[rank] = r1.z;
            for (int ch = 0; ch < 3; ch++) s_col[ch * UPS_BLOCK + rank] = colors[3 * (size_t)id + ch];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(UPS_BLOCK, todo); j++) {
            contributor--;
            if (contributor >= last_contributor) continue;
            const float2 xy = s_xy[j];
            const float dx = xy.x - fx, dy = xy.y - fy;
            const float4 co = s_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.0f) continue;
            const float G = expf(power);
            const float alpha = fminf(0.99f, co.w * G);
            if (alpha < 1.0f / 255.0f) continue;
            T = T / (1.0f - alpha);
            const float w = alpha * T;
            float* a = acc + (size_t)s_id[j] * stride;
            float dL_dalpha = 0.f;
            for (int ch = 0; ch < 3; ch++) {
                const float c = s_col[ch * UPS_BLOCK + j];
                rec_c[ch] = last_alpha * last_c[ch] + (1.f - last_alpha) * rec_c[ch];
                last_c[ch] = c;
                dL_dalpha += (c - rec_c[ch]) * g[ch];
                atomicAdd(a + ch, w * g[ch]);
            }
            const float dep = s_depth[j];
            rec_d = last_alpha * last_d + (1.f - last_alpha) * rec_d;
            last_d = dep;
            dL_dalpha += (dep - rec_d) * g_depth;
            atomicAdd(a + 3, w * g_depth);
            rec_a = last_alpha + (1.f - last_alpha) * rec_a;                 // the alpha channel composites the constant 1
            dL_dalpha += (1.f - rec_a) * g_alpha;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
            const float u = G * dL_dalpha;
            atomicAdd(a + 4, u);
            atomicAdd(a + 5, u * dx);
            atomicAdd(a + 6, u * dy);
            atomicAdd(a + 7, u * dx * dx);
            atomicAdd(a + 8, u * dx * dy);
            atomicAdd(a + 9, u * dy * dy);
        }
    }
}

// stream-ordered scratch from the device's default pool, kept cached between frames (the upstream package gets its
// scratch from torch's caching allocator: no cudaMalloc on the timed path there either)
int pool_ready() {
    static thread_local int done_dev = -1;
    int dev = 0;
    OGS_CUDA(cudaGetDevice(&dev));
    if (done_dev == dev) return 0;
    cudaMemPool_t pool;
    OGS_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = UINT64_MAX;
    OGS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    done_dev = dev;
    return 0;
}

int check_inputs(const ogs_raster_inputs* in) {
    if (!in) { set_error("upstream-structure comparator: inputs is NULL"); return -1; }
    if (in->P <= 0 || in->W <= 0 || in->H <= 0) { set_error("upstream-structure comparator: needs P, W, H > 0"); return -1; }
    if (in->n_extra != 0 || in->act_flags != 0 || in->shs_rest) {
        set_error("upstream-structure comparator: 3 colour channels, activated inputs only (the upstream contract)");
        return -1;
    }
    if ((in->shs != nullptr) == (in->colors_precomp != nullptr)) { set_error("exactly one of shs / colors_precomp"); return -2; }
    if (!in->bg || !in->viewmatrix || !in->projmatrix || !in->campos || !in->means3D || !in->opacities) {
        set_error("upstream-structure comparator: missing inputs");
        return -1;
    }
    return 0;
}

}  // namespace

extern "C" {

int ups_raster_forward(const ogs_raster_inputs* in, const ogs_raster_outputs* out, ogs_alloc_fn alloc, void* alloc_user,
                       ogs_raster_state* st, void* stream_) {
    int rc = check_inputs(in);
    if (rc) return rc;
    if (!out || !out->color || !out->depth || !out->alpha || !out->radii || !alloc || !st) {
        set_error("outputs/alloc/state must be set");
        return -1;
    }
    cudaStream_t s = (cudaStream_t)stream_;
    if ((rc = pool_ready())) return rc;
    const int P = in->P, W = in->W, H = in->H;
    const int gx = (W + 15) / 16, gy = (H + 15) / 16, tiles = gx * gy;
    const bool has_sh = in->shs != nullptr;
    const GeomLayout gl = GeomLayout::make(P, has_sh, 0);
    const ImgLayout il = ImgLayout::make(W, H);
    memset(st, 0, sizeof *st);
    st->geom = alloc(alloc_user, gl.total, "geom");
    st->image = alloc(alloc_user, il.total, "image");
    st->geom_bytes = (int64_t)gl.total;
    st->image_bytes = (int64_t)il.total;
    if (!st->geom || !st->image) { set_error("allocation callback returned NULL"); return -6; }
    const GeomPtrs g = GeomPtrs::from(st->geom, gl);

    const size_t pw = align_up((size_t)P * 4, 256);
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, g.tiles, (uint32_t*)nullptr, P, s);
    scan_bytes = align_up(scan_bytes, 256);
    char* scratch1 = nullptr;
    OGS_CUDA(cudaMallocAsync((void**)&scratch1, 3 * pw + scan_bytes, s));
    uint32_t* dkeys = (uint32_t*)scratch1;
    uint32_t* dvals = (uint32_t*)(scratch1 + pw);
    uint32_t* offsets = (uint32_t*)(scratch1 + 2 * pw);
    void* scan_temp = scratch1 + 3 * pw;

    PreprocessArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.P = P; pa.D = in->sh_degree; pa.M = in->M; pa.W = W; pa.H = H;
    pa.means3D = in->means3D; pa.scales = in->scales; pa.rotations = in->rotations;
    pa.cov3D_precomp = in->cov3D_precomp; pa.opacities = in->opacities; pa.shs = in->shs;
    pa.scale_modifier = in->scale_modifier; pa.tanfovx = in->tanfovx; pa.tanfovy = in->tanfovy;
    pa.view = in->viewmatrix; pa.proj = in->projmatrix; pa.campos = in->campos;
    pa.radii = out->radii; pa.g = g; pa.depth_keys = dkeys; pa.depth_vals = dvals;
    rc = launch_preprocess_forward(pa, s);
    if (rc) { cudaFreeAsync(scratch1, s); return rc; }
    cub::DeviceScan::InclusiveSum(scan_temp, scan_bytes, g.tiles, offsets, P, s);
    uint32_t n32 = 0;
    cudaError_t e = cudaMemcpyAsync(&n32, offsets + (P - 1), 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);          // the upstream forward reads num_rendered here too
    if (e != cudaSuccess) { cudaFreeAsync(scratch1, s); return cuda_fail(e, "upstream-structure forward (scan)"); }
    const int64_t N = (int64_t)n32;
    st->num_rendered = N;

    const BinLayout bl = BinLayout::make(N, tiles);
    st->binning = alloc(alloc_user, bl.total, "binning");
    st->binning_bytes = (int64_t)bl.total;
    if (!st->binning) { cudaFreeAsync(scratch1, s); set_error("allocation callback returned NULL"); return -6; }
    uint32_t* point_list = (uint32_t*)((char*)st->binning + bl.point_list);
    uint2* ranges = (uint2*)((char*)st->binning + bl.ranges);
    e = cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), s);
    if (e != cudaSuccess) { cudaFreeAsync(scratch1, s); return cuda_fail(e, "upstream-structure forward (ranges)"); }

    char* scratch2 = nullptr;
    if (N > 0) {
        int bits = 0;
        while ((1u << bits) < (unsigned)tiles) ++bits;
        const size_t k8 = align_up((size_t)N * 8, 256), v4 = align_up((size_t)N * 4, 256);
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)N, 0, 32 + bits, s);
        sort_bytes = align_up(sort_bytes, 256);
        e = cudaMallocAsync((void**)&scratch2, 2 * k8 + v4 + sort_bytes, s);
        if (e != cudaSuccess) { cudaFreeAsync(scratch1, s); return cuda_fail(e, "upstream-structure forward (binning scratch)"); }
        uint64_t* keys_in = (uint64_t*)scratch2;
        uint64_t* keys_out = (uint64_t*)(scratch2 + k8);
        uint32_t* vals_in = (uint32_t*)(scratch2 + 2 * k8);
        void* sort_temp = scratch2 + 2 * k8 + v4;
        duplicate_with_keys<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec0, g.rec1, g.tiles, offsets, keys_in, vals_in, gx, gy);
        cub::DeviceRadixSort::SortPairs(sort_temp, sort_bytes, keys_in, keys_out, vals_in, point_list, (int)N, 0, 32 + bits, s);
        identify_tile_ranges<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(N, keys_out, ranges);
    }
    const float* colors = has_sh ? g.rgb : in->colors_precomp;
    render_forward<<<dim3(gx, gy), dim3(16, 16), 0, s>>>(W, H, gx, ranges, point_list, g.rec0, g.rec1, colors, in->bg,
                                                       (float*)((char*)st->image + il.final_T),
                                                       (uint32_t*)((char*)st->image + il.n_contrib), out->color, out->depth,
                                                       out->alpha);
    e = cudaGetLastError();
    cudaFreeAsync(scratch1, s);
    if (scratch2) cudaFreeAsync(scratch2, s);
    if (e != cudaSuccess) return cuda_fail(e, "upstream-structure forward");
    return 0;
}

int ups_raster_backward(const ogs_raster_inputs* in, const ogs_raster_state* st, const ogs_raster_grads_in* gin,
                        const ogs_raster_grads_out* go, void* stream_) {
    int rc = check_inputs(in);
    if (rc) return rc;
    if (!st || !gin || !go || !gin->dL_dcolor || gin->dL_dfeat || !go->scratch) { set_error("state/grads/scratch must be set"); return -1; }
    if (go->accumulate || go->dL_dshs_rest || go->dL_dextra) { set_error("upstream-structure comparator: plain gradients only"); return -1; }
    cudaStream_t s = (cudaStream_t)stream_;
    const int P = in->P, W = in->W, H = in->H, C = 3;
    const int gx = (W + 15) / 16, gy = (H + 15) / 16, tiles = gx * gy;
    const bool has_sh = in->shs != nullptr;
    const GeomLayout gl = GeomLayout::make(P, has_sh, 0);
    const ImgLayout il = ImgLayout::make(W, H);
    const BinLayout bl = BinLayout::make(st->num_rendered, tiles);
    const GeomPtrs g = GeomPtrs::from(st->geom, gl);
    const int stride = blend_bwd_stride(C, 1);                               // 3 colours + depth + 6 moments
    float* acc = (float*)go->scratch;
    OGS_CUDA(cudaMemsetAsync(acc, 0, (size_t)P * stride * sizeof(float), s));
    const float* colors = has_sh ? g.rgb : in->colors_precomp;
    render_backward<<<dim3(gx, gy), dim3(16, 16), 0, s>>>(
        W, H, gx, (const uint2*)((char*)st->binning + bl.ranges), (const uint32_t*)((char*)st->binning + bl.point_list), g.rec0, g.rec1,
        colors, in->bg, (const float*)((char*)st->image + il.final_T), (const uint32_t*)((char*)st->image + il.n_contrib),
        gin->dL_dcolor, gin->dL_ddepth, gin->dL_dalpha, acc, stride);
    PreprocessBwdArgs pb;
    memset(&pb, 0, sizeof pb);
    pb.P = P; pb.D = in->sh_degree; pb.M = in->M; pb.C = C; pb.W = W; pb.H = H;
    pb.opacities = in->opacities;
    pb.means3D = in->means3D; pb.scales = in->scales; pb.rotations = in->rotations;
    pb.cov3D_precomp = in->cov3D_precomp; pb.shs = in->shs;
    pb.scale_modifier = in->scale_modifier; pb.tanfovx = in->tanfovx; pb.tanfovy = in->tanfovy;
    pb.view = in->viewmatrix; pb.proj = in->projmatrix; pb.campos = in->campos;
    pb.g = g; pb.acc = acc; pb.stride = stride; pb.geom = 1;
    pb.dL_dmeans3D = go->dL_dmeans3D; pb.dL_dmeans2D = go->dL_dmeans2D; pb.dL_dopacities = go->dL_dopacities;
    pb.dL_dshs = go->dL_dshs; pb.dL_dcolors_precomp = go->dL_dcolors_precomp; pb.dL_dscales = go->dL_dscales;
    pb.dL_drotations = go->dL_drotations; pb.dL_dcov3D = go->dL_dcov3D;
    rc = launch_preprocess_backward(pb, s);
    if (rc) return rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "upstream-structure backward");
    return 0;
}

}  // extern "C"
