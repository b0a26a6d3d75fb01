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
                         int begin_bit, int end_bit, void* scratch, cudaStream_t s, int* result_in_second, uint32_t* n_valid);
int tile_sort_pairs_u16(uint16_t* k0, uint16_t* k1, uint32_t* v0, uint32_t* v1, const uint32_t* n_ptr, uint64_t capacity,
                        int bits, void* scratch, cudaStream_t s);
size_t tile_sort_scratch_bytes(uint64_t capacity);
void tile_sort_plan(int bits, int* passes, int* bits0);
uint32_t* tile_sort_pass0_counts(void* scratch);
size_t radix_scratch_bytes(uint64_t capacity, int passes);
size_t scan_scratch_bytes(int P);
int scan_gather(int P, const uint32_t* n_ptr, const uint32_t* order, const uint32_t* tiles, uint32_t* out, void* scratch,
                cudaStream_t s);

size_t depth_sort_temp_bytes(int P) {
    const size_t a = radix_scratch_bytes((uint64_t)P, 4) + 256, b = scan_scratch_bytes(P);
    return (a > b ? a : b);
}

size_t tile_sort_temp_bytes(int64_t cap) { return tile_sort_scratch_bytes((uint64_t)cap); }

__global__ void set_scalar_kernel(uint32_t* p, uint32_t v) { p[0] = v; p[1] = 0u; }

// The culled Gaussians (about half of a typical frame) carry the key 0xFFFFFFFF: the first sort pass drops them, so
// passes 2-4, the scan and the emit kernel's searches run over the VISIBLE count only (kept on the device: sc.nvis_ptr).
int depth_sort_and_scan(int P, const GeomPtrs& g, BinScratch& sc, cudaStream_t s, int debug) {
    // the item count of the first pass is P itself: park it in the word after the sort scratch, the visible count behind it
    uint32_t* n_ptr = (uint32_t*)((char*)sc.cub_temp + radix_scratch_bytes((uint64_t)P, 4));
    uint32_t* nvis = n_ptr + 1;
    int second = 0;
    set_scalar_kernel<<<1, 1, 0, s>>>(n_ptr, (uint32_t)P);   // (the sort's scratch memset stops short of n_ptr)
    int rc = radix_sort_pairs_u32(sc.dkeys_in, sc.dkeys_out, sc.dvals_in, sc.dvals_out, n_ptr, (uint64_t)P, 0, 32,
                                  sc.cub_temp, s, &second, nvis);
    if (rc) return rc;
    if (!second) {   // result landed in the *_in buffers: swap the roles
        uint32_t* t = sc.dkeys_in; sc.dkeys_in = sc.dkeys_out; sc.dkeys_out = t;
        t = sc.dvals_in; sc.dvals_in = sc.dvals_out; sc.dvals_out = t;
    }
    // (the scan reuses the sort's scratch from its start; the two parked words lie behind the sort scratch, out of its way)
    sc.nvis_ptr = nvis;
    rc = scan_gather(P, nvis, sc.dvals_out, g.tiles, sc.offsets, sc.cub_temp, s);
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
// A chunk is one tile of the tile sort (radix.cu RS_TILE): the CTA also counts the first sort digit
// of its entries and writes its column of the [digit][chunk] histogram, so the sort's first pass
// needs no counting sweep.
#define EMIT_CHUNK 4096
__global__ void __launch_bounds__(256) emit_kernel(int P_cap, const uint32_t* __restrict__ nvis_ptr, int gx, int gy,
                                                   const uint32_t* __restrict__ n_ptr, uint32_t cap,
                                                   const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                                                   const float4* __restrict__ rec0, const float4* __restrict__ rec1,
                                                   uint16_t* __restrict__ tkeys, uint32_t* __restrict__ tvals,
                                                   int bits0, uint32_t* __restrict__ cnt0 /*[1<<bits0][gridDim.x]*/) {
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_start[256];
    __shared__ uint32_t s_g[256];
    __shared__ uint32_t s_rect[256];   // x0 | y0 << 16
    __shared__ uint32_t s_w[256];      // rect width in tiles
    __shared__ uint32_t s_round_end;
    const int t = threadIdx.x;
    const uint32_t N = min(*n_ptr, cap);     // never write past the capacity (an overflow is re-run by the host)
    const int P = (int)min(*nvis_ptr, (uint32_t)P_cap);   // sorted slots in use: the visible Gaussians
    const uint32_t chunk_lo = blockIdx.x * (uint32_t)EMIT_CHUNK;
    const uint32_t dmask = (1u << bits0) - 1u;
    if (chunk_lo >= N) {
        if (t < (1 << bits0)) cnt0[(size_t)t * gridDim.x + blockIdx.x] = 0u;
        return;
    }
    s_hist[t] = 0;
    const uint32_t chunk_hi = min(N, chunk_lo + (uint32_t)EMIT_CHUNK);
    // first sorted slot whose inclusive offset exceeds chunk_lo (it owns entry chunk_lo): a 32-ary search, one probe per
    // lane and one ballot per level -- 4 dependent L2 round trips for 500 k slots where a binary search makes 19
    // (every warp searches for itself: same addresses, no barrier)
    int lo = 0, hi = P;
    while (lo < hi) {
        const int step = (hi - lo + 31) >> 5;
        const int pos = lo + ((t & 31) + 1) * step - 1;
        const bool gt = pos >= hi || __ldg(offsets + pos) > chunk_lo;
        const unsigned m = __ballot_sync(0xffffffffu, gt);
        if (m == 0u) { lo = hi; break; }
        const int f = __ffs(m) - 1;
        hi = min(hi, lo + (f + 1) * step - 1);
        lo = lo + f * step;
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
            const uint32_t key = ((rc >> 16) + dy) * gx + (rc & 0xFFFFu) + dx;
            tkeys[e] = (uint16_t)key;
            tvals[e] = s_g[j];
            atomicAdd(&s_hist[key & dmask], 1u);
        }
        __syncthreads();
        if (e_hi >= chunk_hi) break;
    }
    __syncthreads();
    if (t < (1 << bits0)) cnt0[(size_t)t * gridDim.x + blockIdx.x] = s_hist[t];
}

// ranges[t] = [first, last+1) of tile t in the sorted key list; 8 keys (one 16-byte load) per thread
__global__ void __launch_bounds__(256) ranges_kernel(const uint32_t* __restrict__ n_ptr, uint32_t cap,
                                                     const uint16_t* __restrict__ tk, uint2* __restrict__ ranges) {
    const uint32_t N = min(*n_ptr, cap);
    const uint32_t i0 = (blockIdx.x * 256u + threadIdx.x) * 8u;
    if (i0 >= N) return;
    uint32_t k[8];
    if (i0 + 8 <= N) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(tk + i0));
        k[0] = q.x & 0xFFFFu; k[1] = q.x >> 16; k[2] = q.y & 0xFFFFu; k[3] = q.y >> 16;
        k[4] = q.z & 0xFFFFu; k[5] = q.z >> 16; k[6] = q.w & 0xFFFFu; k[7] = q.w >> 16;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) k[j] = (i0 + j < N) ? tk[i0 + j] : 0xFFFFFFFFu;
    }
    uint32_t prev = (i0 == 0) ? 0xFFFFFFFFu : tk[i0 - 1];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t i = i0 + j;
        if (i < N) {
            if (k[j] != prev) {
                ranges[k[j]].x = i;
                if (i > 0) ranges[prev].y = i;
            }
            if (i == N - 1) ranges[k[j]].y = N;
            prev = k[j];
        }
    }
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
    int passes = 1, bits0 = bits;
    tile_sort_plan(bits, &passes, &bits0);
    // emit writes (k0, v0); the sort leaves the result in (k1, v1) after one pass, (k0, v0) after two
    uint16_t* k0 = tkeys_a; uint16_t* k1 = tkeys_b;
    uint32_t* v0 = passes == 2 ? point_list : tvals_a;
    uint32_t* v1 = passes == 2 ? tvals_a : point_list;
    prof_begin(PF_EMIT, s);
    emit_kernel<<<(unsigned)((cap + EMIT_CHUNK - 1) / EMIT_CHUNK), 256, 0, s>>>(P, sc.nvis_ptr, gx, gy, n_ptr, (uint32_t)cap, sc.dvals_out,
                                                                                sc.offsets, g.rec0, g.rec1, k0, v0, bits0,
                                                                                tile_sort_pass0_counts(sc.cub_temp));
    prof_end(PF_EMIT, s);
    OGS_KERNEL_CHECK("emit_kernel", debug, s);
    prof_begin(PF_TILE_SORT, s);
    const int rc = tile_sort_pairs_u16(k0, k1, v0, v1, n_ptr, (uint64_t)cap, bits, sc.cub_temp, s);
    prof_end(PF_TILE_SORT, s);
    if (rc) return rc;
    OGS_KERNEL_CHECK("tile_sort", debug, s);
    const uint16_t* sorted_keys = passes == 2 ? k0 : k1;
    ProfScope ps(PF_RANGES, s);
    ranges_kernel<<<(unsigned)((cap + 2047) / 2048), 256, 0, s>>>(n_ptr, (uint32_t)cap, sorted_keys, ranges);
    OGS_KERNEL_CHECK("ranges_kernel", debug, s);
    return 0;
}

}  // namespace ogs
