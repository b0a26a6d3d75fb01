// binning.cu -- builds the per-tile, depth-ordered Gaussian lists.
//
// Replaces upstream InclusiveSum + duplicateWithKeys + 64-bit RadixSort + identifyTileRanges of the
// external rasterizer.  The upstream order is: stable ascending sort
// of (tile << 32 | depth bits) with ties kept in Gaussian-index order.  The SAME order is produced
// here with ~6x less sort traffic (SURVEY.md section 7 "Sort cost"):
//   1. stable sort of the P Gaussians by depth bits (32-bit keys, P items; culled -> 0xFFFFFFFF),
//   2. inclusive scan of tiles_touched in that order (its last element is N, kept on the device),
//   3. emit one (tile id uint16, Gaussian idx) pair per touched tile, row-major inside the rect,
//   4. stable sort of the N pairs by tile id only (<= 13 bits => 2 digit passes),
//   5. tile ranges from the sorted tile ids.
// Inside one tile every entry belongs to a distinct Gaussian, so (depth, index) order inside a tile
// after step 4 is exactly the upstream order: point_list and ranges are bit-identical; the 64-bit
// keys are rebuilt on demand by ogs_raster_export for parity tests.
// Sorts and scan are the hand-written kernels of radix.cu.  Every kernel here takes the entry count
// N from device memory and is launched over a capacity, so no host round trip is needed.
//
// Compiled with --fmad=false (the tile-rect arithmetic must match preprocess.cu / the oracle).
#include "common.cuh"

namespace ogs {

int radix_sort_pairs_u32(uint32_t* k0, uint32_t* k1, uint32_t* v0, uint32_t* v1, const uint32_t* n_ptr, uint64_t capacity,
                         int begin_bit, int end_bit, void* scratch, cudaStream_t s, int* result_in_second);
int radix_sort_pairs_u16(uint16_t* k0, uint16_t* k1, uint32_t* v0, uint32_t* v1, const uint32_t* n_ptr, uint64_t capacity,
                         int begin_bit, int end_bit, void* scratch, cudaStream_t s, int* result_in_second);
size_t radix_scratch_bytes(uint64_t capacity, int passes);
size_t scan_scratch_bytes(int P);
int scan_gather(int P, const uint32_t* order, const uint32_t* tiles, uint32_t* out, void* scratch, cudaStream_t s);

size_t depth_sort_temp_bytes(int P) {
    const size_t a = radix_scratch_bytes((uint64_t)P, 4) + 256, b = scan_scratch_bytes(P);
    return (a > b ? a : b);
}

size_t tile_sort_temp_bytes(int64_t cap) { return radix_scratch_bytes((uint64_t)cap, 2); }

__global__ void set_scalar_kernel(uint32_t* p, uint32_t v) { *p = v; }

int depth_sort_and_scan(int P, const GeomPtrs& g, BinScratch& sc, cudaStream_t s, int debug) {
    // the item count of the depth sort is P itself: park it in the word after the sort scratch
    uint32_t* n_ptr = (uint32_t*)((char*)sc.cub_temp + radix_scratch_bytes((uint64_t)P, 4));
    int second = 0;
    set_scalar_kernel<<<1, 1, 0, s>>>(n_ptr, (uint32_t)P);   // (the sort's scratch memset stops short of n_ptr)
    int rc = radix_sort_pairs_u32(sc.dkeys_in, sc.dkeys_out, sc.dvals_in, sc.dvals_out, n_ptr, (uint64_t)P, 0, 32,
                                  sc.cub_temp, s, &second);
    if (rc) return rc;
    if (!second) {   // result landed in the *_in buffers: swap the roles
        uint32_t* t = sc.dkeys_in; sc.dkeys_in = sc.dkeys_out; sc.dkeys_out = t;
        t = sc.dvals_in; sc.dvals_in = sc.dvals_out; sc.dvals_out = t;
    }
    rc = scan_gather(P, sc.dvals_out, g.tiles, sc.offsets, sc.cub_temp, s);
    if (rc) return rc;
    OGS_KERNEL_CHECK("depth_sort_and_scan", debug, s);
    return 0;
}

__device__ __forceinline__ int imin_(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax_(int a, int b) { return a > b ? a : b; }

// Entry-balanced emit: CTA b writes output entries [b*EMIT_CHUNK, (b+1)*EMIT_CHUNK).  The
// depth-sorted Gaussians that own those entries are found by a binary search over the inclusive
// offsets, staged 256 at a time (start offset + rect) in shared memory, and the span is written
// entry-parallel (owner = 8-step binary search in shared memory).  Every store is coalesced and
// every CTA does the same amount of work, however skewed the per-Gaussian tile counts are (the
// nearest Gaussians -- first in depth order -- cover thousands of tiles each).
#define EMIT_CHUNK 2048
__global__ void __launch_bounds__(256) emit_kernel(int P, int gx, int gy, const uint32_t* __restrict__ n_ptr, uint32_t cap,
                                                   const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                                                   const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                                                   uint16_t* __restrict__ tkeys, uint32_t* __restrict__ tvals) {
    __shared__ uint32_t s_start[256];
    __shared__ uint32_t s_g[256];
    __shared__ uint32_t s_rect[256];   // x0 | y0 << 16
    __shared__ uint32_t s_w[256];      // rect width in tiles
    __shared__ uint32_t s_round_end;
    const int t = threadIdx.x;
    const uint32_t N = min(*n_ptr, cap);     // never write past the capacity (an overflow is re-run by the host)
    const uint32_t chunk_lo = blockIdx.x * (uint32_t)EMIT_CHUNK;
    if (chunk_lo >= N) return;
    const uint32_t chunk_hi = min(N, chunk_lo + (uint32_t)EMIT_CHUNK);
    // first sorted slot whose inclusive offset exceeds chunk_lo (it owns entry chunk_lo)
    int lo = 0, hi = P;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) > chunk_lo) hi = mid; else lo = mid + 1;
    }
    for (int round_start = lo; round_start < P; round_start += 256) {
        const int slot = round_start + t;
        uint32_t start = 0xFFFFFFFFu, g = 0, rect = 0, rw = 1;
        if (slot < P) {
            start = (slot == 0) ? 0u : __ldg(offsets + slot - 1);
            if (start < chunk_hi) {
                g = order[slot];
                const float4 b = rec1[g];
                const int radius = __float_as_int(b.w);
                if (radius > 0) {
                    const float4 a = rec0[g];
                    const float r = (float)radius;
                    const int x0 = imin_(gx, imax_(0, (int)((a.x - r) / 16.0f)));
                    const int y0 = imin_(gy, imax_(0, (int)((a.y - r) / 16.0f)));
                    const int x1 = imin_(gx, imax_(0, (int)((a.x + r + 15.0f) / 16.0f)));
                    rect = (uint32_t)x0 | ((uint32_t)y0 << 16);
                    rw = (uint32_t)(x1 - x0);
                }
            }
        }
        const int last_slot = min(P, round_start + 256) - 1;
        if (slot == last_slot) s_round_end = min(chunk_hi, __ldg(offsets + slot));
        s_start[t] = start;
        s_g[t] = g;
        s_rect[t] = rect;
        s_w[t] = rw;
        __syncthreads();
        const uint32_t e_lo = max(chunk_lo, s_start[0]);
        const uint32_t e_hi = s_round_end;
        for (uint32_t e = e_lo + t; e < e_hi; e += 256) {
            int j = 0;   // last j with s_start[j] <= e
#pragma unroll
            for (int step = 128; step >= 1; step >>= 1)
                if (s_start[j + step] <= e) j += step;
            const uint32_t rc = s_rect[j];
            const uint32_t local = e - s_start[j];
            const uint32_t w = s_w[j];
            const uint32_t dy = local / w, dx = local - dy * w;
            tkeys[e] = (uint16_t)(((rc >> 16) + dy) * gx + (rc & 0xFFFFu) + dx);
            tvals[e] = s_g[j];
        }
        __syncthreads();
        if (e_hi >= chunk_hi) break;
    }
}

__global__ void __launch_bounds__(256) ranges_kernel(const uint32_t* __restrict__ n_ptr, uint32_t cap,
                                                     const uint16_t* __restrict__ tk, uint2* __restrict__ ranges) {
    const uint32_t N = min(*n_ptr, cap);
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= N) return;
    const uint32_t t = tk[i];
    if (i == 0) ranges[t].x = 0;
    else {
        const uint32_t pt = tk[i - 1];
        if (t != pt) { ranges[pt].y = i; ranges[t].x = i; }
    }
    if (i == N - 1) ranges[t].y = N;
}

// n_ptr: device count of entries (last inclusive offset); cap: capacity of the key/value buffers.
// The tile sort ping-pongs between (tkeys_a, tvals_a) and (tkeys_b, point_list); emit writes into
// the side from which the sorted values end in point_list after `passes` passes (no final copy).
int emit_sort_ranges(int P, int W, int H, const uint32_t* n_ptr, int64_t cap, const GeomPtrs& g, BinScratch& sc,
                     uint16_t* tkeys_a, uint32_t* tvals_a, uint16_t* tkeys_b, uint32_t* point_list, uint2* ranges,
                     cudaStream_t s, int debug) {
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    const int tiles = gx * gy;
    OGS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), s));
    if (cap <= 0) return 0;
    int bits = 0;
    while ((1 << bits) < tiles) bits++;
    if (bits == 0) bits = 1;
    const int passes = (bits + 7) / 8;
    uint16_t *kX, *kY;
    uint32_t *vX, *vY;
    if (passes & 1) { kX = tkeys_a; vX = tvals_a; kY = tkeys_b; vY = point_list; }
    else { kX = tkeys_b; vX = point_list; kY = tkeys_a; vY = tvals_a; }
    prof_begin(PF_EMIT, s);
    emit_kernel<<<(unsigned)((cap + EMIT_CHUNK - 1) / EMIT_CHUNK), 256, 0, s>>>(P, gx, gy, n_ptr, (uint32_t)cap, sc.dvals_out,
                                                                                sc.offsets, g.rec0, g.rec1, kX, vX);
    prof_end(PF_EMIT, s);
    OGS_KERNEL_CHECK("emit_kernel", debug, s);
    int second = 0;
    prof_begin(PF_TILE_SORT, s);
    const int rc = radix_sort_pairs_u16(kX, kY, vX, vY, n_ptr, (uint64_t)cap, 0, bits, sc.cub_temp, s, &second);
    prof_end(PF_TILE_SORT, s);
    if (rc) return rc;
    OGS_KERNEL_CHECK("tile_sort", debug, s);
    const uint16_t* sorted_keys = second ? kY : kX;
    ProfScope ps(PF_RANGES, s);
    ranges_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(n_ptr, (uint32_t)cap, sorted_keys, ranges);
    OGS_KERNEL_CHECK("ranges_kernel", debug, s);
    return 0;
}

}  // namespace ogs
