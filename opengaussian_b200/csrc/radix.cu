// radix.cu -- hand-written stable LSD radix sort (pairs) and inclusive scan for the binning stage.
//
// Replaces the cub::DeviceRadixSort::SortPairs / cub::DeviceScan::InclusiveSum calls of the
// upstream rasterizer for the two sorts of binning.cu:
//   * depth sort: P items, 32-bit keys (depth bits), 32-bit values (Gaussian index), 4 digit passes;
//   * tile sort : N items, 16-bit keys (tile id), 32-bit values, 2 digit passes (<= 13 key bits).
// Design (one kernel per digit pass, each key/value read once and written once -- HBM-bound):
//   1. rs_hist_kernel: digit histograms of ALL passes in one sweep over the keys;
//   2. rs_pass_kernel ("onesweep"): a CTA takes a 4096-item tile through a ticket counter (so that
//      tile t only ever waits on tiles < t that have already started), ranks its keys stably
//      (peer groups from one ballot per digit bit + per-warp digit counters), publishes its digit counts and
//      resolves its global digit offsets by decoupled look-back over the preceding tiles' published
//      (aggregate | inclusive) words, reorders keys/values by local rank in shared memory and writes
//      each digit run with coalesced stores.
// The item count is read from DEVICE memory (n_ptr): the grid is sized for a capacity and CTAs
// past the end exit, which is what lets ogs_raster_forward run without a host round trip.
#include <stdlib.h>

#include "common.cuh"

namespace ogs {

#define RS_THREADS 256
#define RS_WARPS 8
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)   // 4096: tile of the N-entry tile sort
#define RS_ITEMS_SMALL 16
#define RS_TILE_SMALL (RS_THREADS * RS_ITEMS_SMALL)   // tile of the P-entry depth sort (245 tiles at P = 1 M: one wave)
#define RS_RADIX 256
#define RS_FLAG_AGG 0x40000000u
#define RS_FLAG_INC 0x80000000u
#define RS_VAL_MASK 0x3FFFFFFFu
#ifndef RS_LB
#define RS_LB 32   // predecessor tiles read per look-back round trip (independent loads in flight)
#endif

struct RsPasses {
    int num;
    int shift[4];
    int bits[4];
};

// n_valid != NULL ("drop" mode of the depth sort): keys equal to 0xFFFFFFFF -- culled Gaussians -- are neither counted
// in the digit histograms nor sorted; their complement is counted into *n_valid, the item count of passes 2..4.
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ n_ptr,
                                                             uint32_t cap, RsPasses ps, uint32_t* __restrict__ ghist /*[num][256]*/,
                                                             uint32_t* __restrict__ n_valid) {
    __shared__ uint32_t s_h[4 * RS_RADIX];
    for (int e = threadIdx.x; e < ps.num * RS_RADIX; e += RS_THREADS) s_h[e] = 0;
    __syncthreads();
    const uint32_t n = min(*n_ptr, cap);
    // drop mode: dropped keys are counted like any other (their digit is 0xFF in every pass) and taken out of the
    // 0xFF bins once per CTA at the end -- the hot loop only gains a compare and an add
    uint32_t ndrop = 0;
    auto count = [&](uint32_t k) {
        ndrop += (k == 0xFFFFFFFFu) ? 1u : 0u;
#pragma unroll
        for (int p = 0; p < 4; p++)
            if (p < ps.num) atomicAdd(&s_h[p * RS_RADIX + ((k >> ps.shift[p]) & ((1u << ps.bits[p]) - 1u))], 1u);
    };
    if (sizeof(KeyT) == 4) {   // four keys per 16-byte load, two loads in flight per thread
        const uint32_t n4 = n / 4;
        const uint4* k4 = reinterpret_cast<const uint4*>(keys);
        for (uint32_t i = blockIdx.x * RS_THREADS + threadIdx.x; i < n4; i += 2 * gridDim.x * RS_THREADS) {
            const uint32_t i2 = i + gridDim.x * RS_THREADS;
            const uint4 qa = __ldg(k4 + i);
            const uint4 qb = i2 < n4 ? __ldg(k4 + i2) : make_uint4(0, 0, 0, 0);
            count(qa.x); count(qa.y); count(qa.z); count(qa.w);
            if (i2 < n4) { count(qb.x); count(qb.y); count(qb.z); count(qb.w); }
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) count((uint32_t)keys[n4 * 4 + threadIdx.x]);
    } else {
        for (uint32_t i = blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += gridDim.x * RS_THREADS) count((uint32_t)keys[i]);
    }
    __syncthreads();
    if (n_valid) {
        __shared__ uint32_t s_drop;
        if (threadIdx.x == 0) s_drop = 0;
        __syncthreads();
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) ndrop += __shfl_xor_sync(0xffffffffu, ndrop, m);
        if ((threadIdx.x & 31) == 0 && ndrop) atomicAdd(&s_drop, ndrop);
        __syncthreads();
        if ((int)threadIdx.x < ps.num) s_h[threadIdx.x * RS_RADIX + ((1u << ps.bits[threadIdx.x]) - 1u)] -= s_drop;
        // *n_valid = n - (all dropped keys): CTA 0 contributes n, every CTA takes its dropped keys off (mod 2^32)
        if (threadIdx.x == 0) atomicAdd(n_valid, (blockIdx.x == 0 ? n : 0u) - s_drop);
        __syncthreads();
    }
    for (int e = threadIdx.x; e < ps.num * RS_RADIX; e += RS_THREADS)
        if (s_h[e]) atomicAdd(&ghist[e], s_h[e]);
}

// LOOKBACK = true : onesweep (ticketed tiles, decoupled look-back over tile_state[tile][256]).
// LOOKBACK = false: plain scatter; the tile's exclusive digit offsets were computed beforehand
//                   (rs_tile_hist_kernel / emit + rs_tile_scan_kernel) and are read from
//                   tile_state[digit * ntiles + tile]; ghist holds the digit totals.
#ifndef RS_MINB
#define RS_MINB 2
#endif
// DROP = true (first pass of the depth sort): items whose key is 0xFFFFFFFF are treated like items past the end -- they
// get no rank and are not written -- so the output holds the valid items only, in stable order.
template <typename KeyT, bool LOOKBACK, int ITEMS, bool DROP = false>
__global__ void __launch_bounds__(RS_THREADS, RS_MINB) rs_pass_kernel(const KeyT* __restrict__ kin, KeyT* __restrict__ kout,
                                                             const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                             const uint32_t* __restrict__ n_ptr, uint32_t cap, int shift, int bits,
                                                             const uint32_t* __restrict__ ghist /*[256] of this pass*/,
                                                             uint32_t* tile_state, uint32_t* ticket, uint32_t ntiles) {
    constexpr int TILE = RS_THREADS * ITEMS;
    __shared__ uint32_t s_wh[RS_WARPS * RS_RADIX];   // per-warp digit counters -> exclusive warp offsets
    __shared__ uint32_t s_dstart[RS_RADIX];          // local exclusive start of each digit in the tile
    __shared__ uint32_t s_gbase[RS_RADIX];           // global position of local sorted slot 0 of the digit
    __shared__ __align__(16) KeyT s_keys[TILE];
    __shared__ __align__(16) uint32_t s_vals[TILE];
    __shared__ uint32_t s_tile, s_tile_valid;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (LOOKBACK && tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int e = tid; e < RS_WARPS * RS_RADIX; e += RS_THREADS) s_wh[e] = 0;
    __syncthreads();
    const uint32_t tile = LOOKBACK ? s_tile : blockIdx.x;
    const uint32_t n = min(*n_ptr, cap);
    const uint32_t tile_start = tile * (uint32_t)TILE;
    if (tile_start >= n) return;
    const uint32_t tile_n = min((uint32_t)TILE, n - tile_start);
    const uint32_t dmask = (1u << bits) - 1u;

    // ---- load (warp-striped: item i of lane l = chunk[i*32 + l]) and rank ----
    const uint32_t wbase = tile_start + warp * (32 * ITEMS);
    uint32_t key[ITEMS];
    uint32_t rank[ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* wh = s_wh + warp * RS_RADIX;
    uint32_t val[ITEMS];
    // Whole tiles come in through the reorder buffers: every thread issues its share of the tile's keys AND values as
    // 16-byte loads back to back (one DRAM round trip for the CTA), stores them to shared memory and picks up its
    // warp-striped items from there.  The per-item form below compiled to 32 narrow loads per thread of which ptxas
    // (128-register budget) issued eleven key loads together, the other five each directly in front of its first use,
    // and the value loads only after the ranking -- seven exposed round trips per tile with two CTAs per SM to hide them
    // (ncu source page: the first use of each late key is the kernel's top stall site after the scatter).
    constexpr int NK = (int)(ITEMS * sizeof(KeyT) / 16), NV = ITEMS / 4;
    // (Not for the look-back variant: the depth sort's 245 tiles are one wave whose passes are bound by the chain
    // rank -> publish -> look back; the staging barrier in front of the ranking made it 8 us slower, measured.)
    const bool whole = !LOOKBACK && NK > 0 && NK * 16 == (int)(ITEMS * sizeof(KeyT)) && (ITEMS % 4) == 0 && tile_n == (uint32_t)TILE &&
                       (((uintptr_t)(kin + tile_start) | (uintptr_t)(vin + tile_start)) & 15) == 0;
    if (whole) {
        const uint4* ksrc = reinterpret_cast<const uint4*>(kin + tile_start);
        const uint4* vsrc = reinterpret_cast<const uint4*>(vin + tile_start);
        uint4 qk[NK > 0 ? NK : 1], qv[NV > 0 ? NV : 1];
#pragma unroll
        for (int u = 0; u < NK; u++) qk[u] = __ldg(ksrc + tid + u * RS_THREADS);
#pragma unroll
        for (int u = 0; u < NV; u++) qv[u] = __ldg(vsrc + tid + u * RS_THREADS);
#pragma unroll
        for (int u = 0; u < NK; u++) reinterpret_cast<uint4*>(s_keys)[tid + u * RS_THREADS] = qk[u];
#pragma unroll
        for (int u = 0; u < NV; u++) reinterpret_cast<uint4*>(s_vals)[tid + u * RS_THREADS] = qv[u];
        __syncthreads();
        const int lbase = warp * (32 * ITEMS) + lane;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            key[i] = (uint32_t)s_keys[lbase + i * 32];
            val[i] = s_vals[lbase + i * 32];
        }
        // the buffers are next written by the reorder phase, behind the barriers of the ranking
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t idx = wbase + i * 32 + lane;
            key[i] = idx < n ? (uint32_t)kin[idx] : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const uint32_t idx = wbase + i * 32 + lane;
            val[i] = idx < n ? vin[idx] : 0u;
        }
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t idx = wbase + i * 32 + lane;
        const bool valid = idx < n && (!DROP || key[i] != 0xFFFFFFFFu);
        const uint32_t d = valid ? ((key[i] >> shift) & dmask) : 0xFFFFFFFFu;
        // lanes holding the same digit: one ballot per digit bit.  (MATCH.ANY costs one round per
        // DISTINCT value in the warp, and here nearly all 32 digits differ.)
        uint32_t peers = __ballot_sync(0xffffffffu, valid);
        if (!valid) peers = ~peers;
#pragma unroll
        for (int bb = 0; bb < 8; bb++) {
            if (bb < bits) {
                const bool bit = (d >> bb) & 1u;
                const uint32_t m = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? m : ~m;
            }
        }
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && lane == leader) {
            pre = wh[d];
            wh[d] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[i] = pre + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: warp-exclusive offsets, tile count, local digit starts ----
    uint32_t count = 0;
    {
        const int d = tid;   // RS_THREADS == RS_RADIX
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t t = s_wh[w * RS_RADIX + d];
            s_wh[w * RS_RADIX + d] = count;
            count += t;
        }
    }
    __syncthreads();
    // exclusive scans over the 256 digits: local tile counts (s_dstart) and global totals (digit base)
    uint32_t dstart = 0, dbase = 0;
    {
        // warp-level scan of 256 values: each warp scans 32, then warp totals are combined
        const uint32_t gtot = (LOOKBACK || tid < (1 << bits)) ? ghist[tid] : 0u;
        uint32_t a = count, b = gtot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) { a += ta; b += tb; }
        }
        __shared__ uint32_t s_wa[RS_WARPS], s_wb[RS_WARPS];
        if (lane == 31) { s_wa[warp] = a; s_wb[warp] = b; }
        __syncthreads();
        uint32_t oa = 0, ob = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++)
            if (w < warp) { oa += s_wa[w]; ob += s_wb[w]; }
        dstart = oa + a - count;
        dbase = ob + b - gtot;
        if (DROP && tid == RS_THREADS - 1) s_tile_valid = dstart + count;   // items of this tile that take part
    }
    // ---- global offset of this tile's digit `tid`: precomputed, or decoupled look-back ----
    if (!LOOKBACK) {
        const int d = tid;
        const uint32_t excl = (d < (1 << bits)) ? tile_state[(size_t)d * ntiles + tile] : 0u;
        s_dstart[d] = dstart;
        s_gbase[d] = dbase + excl - dstart;
    } else {
        const int d = tid;
        volatile uint32_t* st = tile_state;
        uint32_t excl = 0;
        if (tile == 0) {
            st[(size_t)tile * RS_RADIX + d] = count | RS_FLAG_INC;
        } else {
            st[(size_t)tile * RS_RADIX + d] = count | RS_FLAG_AGG;
            __threadfence();
            // look back RS_LB predecessor tiles per round trip (independent loads in flight): a thread
            // walking one tile per ~700-cycle load would serialise a whole wave of CTAs
            int t = (int)tile - 1;
            while (t >= 0) {
                uint32_t v[RS_LB];
#pragma unroll
                for (int q = 0; q < RS_LB; q++)
                    v[q] = (t - q >= 0) ? st[(size_t)(t - q) * RS_RADIX + d] : RS_FLAG_INC;
#pragma unroll
                for (int q = 0; q < RS_LB; q++) {
                    if (v[q] & RS_FLAG_INC) { excl += v[q] & RS_VAL_MASK; t = -1; break; }
                    if (v[q] & RS_FLAG_AGG) { excl += v[q] & RS_VAL_MASK; t--; }
                    else break;
                }
            }
            st[(size_t)tile * RS_RADIX + d] = (excl + count) | RS_FLAG_INC;
        }
        s_dstart[d] = dstart;
        s_gbase[d] = dbase + excl - dstart;
    }
    __syncthreads();

    // ---- reorder by local rank in shared memory ----
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const uint32_t idx = wbase + i * 32 + lane;
        if (idx < n && (!DROP || key[i] != 0xFFFFFFFFu)) {
            const uint32_t d = (key[i] >> shift) & dmask;
            const uint32_t pos = s_dstart[d] + s_wh[warp * RS_RADIX + d] + rank[i];
            s_keys[pos] = (KeyT)key[i];
            s_vals[pos] = val[i];
        }
    }
    __syncthreads();
    // ---- coalesced write-out of the digit runs ----
    const uint32_t tile_out = DROP ? s_tile_valid : tile_n;
    for (uint32_t j = tid; j < tile_out; j += RS_THREADS) {
        const KeyT k = s_keys[j];
        const uint32_t d = ((uint32_t)k >> shift) & dmask;
        const uint32_t dst = s_gbase[d] + j;
        kout[dst] = k;
        vout[dst] = s_vals[j];
    }
}

// scratch layout for one sort: [ghist 4*256][ticket 4 (one per pass) + pad][tile_state passes * tiles * 256]
size_t radix_scratch_bytes(uint64_t capacity, int passes) {
    const size_t tiles = (size_t)((capacity + RS_THREADS * 4 - 1) / (RS_THREADS * 4)) + 1;      // smallest tile variant
    return align_up((size_t)4 * RS_RADIX * 4, 256) + 256 + (size_t)passes * tiles * RS_RADIX * 4;
}

// n_valid != NULL: drop mode -- keys equal to 0xFFFFFFFF are removed by the first pass; *n_valid (zero on entry, device
// memory) receives the number of remaining items, which is the item count of the later passes and of the result.
template <typename KeyT, int ITEMS>
static int radix_sort_pairs_t(KeyT* k0, KeyT* k1, uint32_t* v0, uint32_t* v1, const uint32_t* n_ptr, uint64_t capacity,
                              int begin_bit, int end_bit, void* scratch, cudaStream_t s, int* result_in_second,
                              uint32_t* n_valid = nullptr) {
    RsPasses ps;
    ps.num = 0;
    for (int b = begin_bit; b < end_bit && ps.num < 4; b += 8) {
        ps.shift[ps.num] = b;
        ps.bits[ps.num] = (end_bit - b) < 8 ? (end_bit - b) : 8;
        ps.num++;
    }
    // balance a short last digit with the one before (e.g. 13 bits -> 7 + 6 instead of 8 + 5)
    if (ps.num >= 2 && ps.bits[ps.num - 1] < 8) {
        const int tot = ps.bits[ps.num - 2] + ps.bits[ps.num - 1];
        ps.bits[ps.num - 2] = (tot + 1) / 2;
        ps.shift[ps.num - 1] = ps.shift[ps.num - 2] + ps.bits[ps.num - 2];
        ps.bits[ps.num - 1] = tot - ps.bits[ps.num - 2];
    }
    constexpr int TILE = RS_THREADS * ITEMS;
    const size_t tiles = (size_t)((capacity + TILE - 1) / TILE) + 1;
    char* base = (char*)scratch;
    uint32_t* ghist = (uint32_t*)base;
    uint32_t* tickets = (uint32_t*)(base + align_up((size_t)4 * RS_RADIX * 4, 256));
    uint32_t* states = (uint32_t*)(base + align_up((size_t)4 * RS_RADIX * 4, 256) + 256);
    OGS_CUDA(cudaMemsetAsync(scratch, 0, align_up((size_t)4 * RS_RADIX * 4, 256) + 256 + (size_t)ps.num * tiles * RS_RADIX * 4, s));
    int hist_grid = (int)((capacity + RS_THREADS * 8 - 1) / (RS_THREADS * 8));
    if (hist_grid > OGS_NUM_SMS * 4) hist_grid = OGS_NUM_SMS * 4;
    if (hist_grid < 1) hist_grid = 1;
    rs_hist_kernel<KeyT><<<hist_grid, RS_THREADS, 0, s>>>(k0, n_ptr, (uint32_t)capacity, ps, ghist, n_valid);
    KeyT* ka = k0; KeyT* kb = k1;
    uint32_t* va = v0; uint32_t* vb = v1;
    const unsigned grid = (unsigned)((capacity + TILE - 1) / TILE);
    for (int p = 0; p < ps.num; p++) {
        if (n_valid && p == 0)
            rs_pass_kernel<KeyT, true, ITEMS, true><<<grid, RS_THREADS, 0, s>>>(ka, kb, va, vb, n_ptr, (uint32_t)capacity, ps.shift[p],
                                                                                ps.bits[p], ghist + p * RS_RADIX,
                                                                                states + (size_t)p * tiles * RS_RADIX, tickets + p, 0u);
        else
            rs_pass_kernel<KeyT, true, ITEMS><<<grid, RS_THREADS, 0, s>>>(ka, kb, va, vb, n_valid ? n_valid : n_ptr, (uint32_t)capacity,
                                                                          ps.shift[p], ps.bits[p], ghist + p * RS_RADIX,
                                                                          states + (size_t)p * tiles * RS_RADIX, tickets + p, 0u);
        KeyT* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    *result_in_second = (ps.num & 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "radix_sort_pairs");
    return 0;
}

int radix_sort_pairs_u32(uint32_t* k0, uint32_t* k1, uint32_t* v0, uint32_t* v1, const uint32_t* n_ptr, uint64_t capacity,
                         int begin_bit, int end_bit, void* scratch, cudaStream_t s, int* result_in_second, uint32_t* n_valid) {
    static const int env_items = getenv("OGS_RS_ITEMS") ? atoi(getenv("OGS_RS_ITEMS")) : RS_ITEMS_SMALL;   // tuning knob
    if (env_items == 4)
        return radix_sort_pairs_t<uint32_t, 4>(k0, k1, v0, v1, n_ptr, capacity, begin_bit, end_bit, scratch, s, result_in_second, n_valid);
    if (env_items == 8)
        return radix_sort_pairs_t<uint32_t, 8>(k0, k1, v0, v1, n_ptr, capacity, begin_bit, end_bit, scratch, s, result_in_second, n_valid);
    return radix_sort_pairs_t<uint32_t, 16>(k0, k1, v0, v1, n_ptr, capacity, begin_bit, end_bit, scratch, s, result_in_second,
                                            n_valid);
}
// ---------------------------------------------------------------------------------------------
// Tile sort (N entries, 16-bit tile ids, <= 2 digit passes): counting passes WITHOUT look-back.
// With thousands of 4096-item tiles in flight a look-back walk is hundreds of predecessors long
// (every CTA re-reads their 1 KB state rows through L2); here each pass is
//   per-tile digit histogram [digit][tile]  (pass 1: written by the emit kernel itself,
//                                            pass 2: rs_tile_hist_kernel)
//   -> rs_tile_scan_kernel (one CTA per digit: exclusive scan along the tiles + digit total)
//   -> rs_pass_kernel<LOOKBACK=false> (rank in shared memory, coalesced scatter).
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_tile_hist_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ n_ptr,
                                                                  uint32_t cap, int shift, int bits, uint32_t ntiles,
                                                                  uint32_t* __restrict__ cnt /*[1<<bits][ntiles]*/) {
    __shared__ uint32_t s_h[RS_RADIX];
    const int tid = threadIdx.x;
    s_h[tid] = 0;
    __syncthreads();
    const uint32_t n = min(*n_ptr, cap);
    const uint32_t tile = blockIdx.x, start = tile * (uint32_t)RS_TILE;
    const uint32_t dmask = (1u << bits) - 1u;
    if (start < n) {
        const uint32_t end = min(n, start + (uint32_t)RS_TILE);
        if (sizeof(KeyT) == 2 && end - start == RS_TILE) {   // full tile of 16-bit keys: 16-byte loads
            const uint4* src = reinterpret_cast<const uint4*>(keys + start);
            for (int e = tid; e < RS_TILE / 8; e += RS_THREADS) {
                const uint4 q = __ldg(src + e);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    atomicAdd(&s_h[((w[k] & 0xFFFFu) >> shift) & dmask], 1u);
                    atomicAdd(&s_h[((w[k] >> 16) >> shift) & dmask], 1u);
                }
            }
        } else {
            for (uint32_t i = start + tid; i < end; i += RS_THREADS) atomicAdd(&s_h[((uint32_t)keys[i] >> shift) & dmask], 1u);
        }
    }
    __syncthreads();
    if (tid < (1 << bits)) cnt[(size_t)tid * ntiles + tile] = s_h[tid];
}

// cnt[digit][0..ntiles) -> exclusive scan in place; totals[digit] = sum.  One CTA per digit.
__global__ void __launch_bounds__(RS_THREADS) rs_tile_scan_kernel(uint32_t* __restrict__ cnt, uint32_t ntiles,
                                                                  uint32_t* __restrict__ totals /*[256]*/) {
    __shared__ uint32_t s_w[RS_WARPS];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t* row = cnt + (size_t)blockIdx.x * ntiles;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += RS_THREADS * 4) {
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t idx = base + tid * 4 + i;
            v[i] = idx < ntiles ? row[idx] : 0u;
            sum += v[i];
        }
        uint32_t a = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, a, o);
            if (lane >= o) a += t;
        }
        if (lane == 31) s_w[warp] = a;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            if (w < warp) woff += s_w[w];
            total += s_w[w];
        }
        uint32_t run = s_carry + woff + (a - sum);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t idx = base + tid * 4 + i;
            if (idx < ntiles) row[idx] = run;
            run += v[i];
        }
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = s_carry;
}

// scratch: [totals 2*256][cnt pass0: 256*ntiles][cnt pass1: 256*ntiles]
static size_t tile_sort_ntiles(uint64_t capacity) { return (size_t)((capacity + RS_TILE - 1) / RS_TILE); }
size_t tile_sort_scratch_bytes(uint64_t capacity) {
    return align_up((size_t)2 * RS_RADIX * 4, 256) + (size_t)2 * RS_RADIX * tile_sort_ntiles(capacity) * 4;
}
// digit split of a `bits`-wide tile id: one pass up to 8 bits, else two balanced passes
void tile_sort_plan(int bits, int* passes, int* bits0) {
    if (bits <= 8) { *passes = 1; *bits0 = bits; }
    else { *passes = 2; *bits0 = (bits + 1) / 2; }
}
uint32_t* tile_sort_pass0_counts(void* scratch) { return (uint32_t*)((char*)scratch + align_up((size_t)2 * RS_RADIX * 4, 256)); }

// Sorts (k0, v0)[0..*n_ptr) by the low `bits` bits of the 16-bit keys.  The per-tile histogram of the
// FIRST digit must already be in tile_sort_pass0_counts(scratch) ([digit][ntiles], written by the
// producer of k0 -- the emit kernel).  Result lands in (k1, v1) after 1 pass, (k0, v0) after 2.
int tile_sort_pairs_u16(uint16_t* k0, uint16_t* k1, uint32_t* v0, uint32_t* v1, const uint32_t* n_ptr, uint64_t capacity,
                        int bits, void* scratch, cudaStream_t s) {
    int passes, bits0;
    tile_sort_plan(bits, &passes, &bits0);
    const uint32_t ntiles = (uint32_t)tile_sort_ntiles(capacity);
    uint32_t* totals = (uint32_t*)scratch;
    uint32_t* cnt0 = tile_sort_pass0_counts(scratch);
    uint32_t* cnt1 = cnt0 + (size_t)RS_RADIX * ntiles;
    rs_tile_scan_kernel<<<1 << bits0, RS_THREADS, 0, s>>>(cnt0, ntiles, totals);
    rs_pass_kernel<uint16_t, false, RS_ITEMS><<<ntiles, RS_THREADS, 0, s>>>(k0, k1, v0, v1, n_ptr, (uint32_t)capacity, 0, bits0, totals,
                                                                  cnt0, nullptr, ntiles);
    if (passes == 2) {
        const int bits1 = bits - bits0;
        rs_tile_hist_kernel<uint16_t><<<ntiles, RS_THREADS, 0, s>>>(k1, n_ptr, (uint32_t)capacity, bits0, bits1, ntiles, cnt1);
        rs_tile_scan_kernel<<<1 << bits1, RS_THREADS, 0, s>>>(cnt1, ntiles, totals + RS_RADIX);
        rs_pass_kernel<uint16_t, false, RS_ITEMS><<<ntiles, RS_THREADS, 0, s>>>(k1, k0, v1, v0, n_ptr, (uint32_t)capacity, bits0, bits1,
                                                                      totals + RS_RADIX, cnt1, nullptr, ntiles);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "tile_sort_pairs");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Inclusive scan of tiles_touched gathered in depth order: out[i] = sum_{j<=i} tiles[order[j]].
// Single pass, decoupled look-back over 2048-item tiles (same ticket scheme as the sort).
#define SC_THREADS 256
#define SC_ITEMS 8
#define SC_TILE (SC_THREADS * SC_ITEMS)

// Inclusive scan of tiles[order[i]] over the first n = *n_ptr sorted slots (n <= P: the visible Gaussians).  The grand
// total -- N, the number of (Gaussian, tile) duplicates -- goes to out[P] and an overflow flag to out[P + 1]: the count
// and every offset are 32-bit (as upstream's), so a frame with more than 2^32 - 1 duplicates would wrap silently; the
// prefix is carried in 64 bits as well and the flag is raised instead.
__global__ void __launch_bounds__(SC_THREADS) scan_gather_kernel(int P, const uint32_t* __restrict__ n_ptr,
                                                                 const uint32_t* __restrict__ order,
                                                                 const uint32_t* __restrict__ tiles, uint32_t* __restrict__ out,
                                                                 uint32_t* tile_state, uint32_t* ticket) {
    __shared__ uint32_t s_tile, s_warp[SC_THREADS / 32], s_excl;
    __shared__ unsigned long long s_excl64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t n = min(*n_ptr, (uint32_t)P);
    if (n == 0) {
        if (tile == 0 && tid == 0) { out[P] = 0u; out[P + 1] = 0u; }
        return;
    }
    if (tile * (uint32_t)SC_TILE >= n) return;          // tiles are ticketed: nobody ever waits on a tile past the end
    const uint32_t base = tile * (uint32_t)SC_TILE + tid * SC_ITEMS;
    uint32_t v[SC_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t idx = base + i;
        v[i] = idx < n ? tiles[order[idx]] : 0u;
        sum += v[i];
        v[i] = sum;
    }
    uint32_t a = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, a, o);
        if (lane >= o) a += t;
    }
    if (lane == 31) s_warp[warp] = a;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; w++) {
        if (w < warp) woff += s_warp[w];
        total += s_warp[w];
    }
    // Every tile publishes its total; the tile's threads then sum ALL predecessors' totals
    // cooperatively (tiles are ticketed, so every predecessor is already running and will publish
    // without waiting on anyone): one round trip instead of a serial walk over hundreds of tiles.
    {
        volatile uint32_t* st = tile_state;
        if (tid == 0) {
            st[tile] = total | RS_FLAG_AGG;
            s_excl = 0;
            s_excl64 = 0ull;
        }
        __syncthreads();
        uint32_t part = 0;
        unsigned long long part64 = 0ull;
        for (uint32_t t = tid; t < tile; t += SC_THREADS) {
            uint32_t x;
            do { x = st[t]; } while (!(x & RS_FLAG_AGG));
            part += x & RS_VAL_MASK;
            part64 += x & RS_VAL_MASK;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            part += __shfl_xor_sync(0xffffffffu, part, m);
            part64 += __shfl_xor_sync(0xffffffffu, part64, m);
        }
        if (lane == 0 && part64) { atomicAdd(&s_excl, part); atomicAdd(&s_excl64, part64); }
    }
    __syncthreads();
    if (tid == 0 && tile == (n - 1) / (uint32_t)SC_TILE) {
        out[P] = s_excl + total;
        out[P + 1] = (s_excl64 + total > 0xFFFFFFFFull) ? 1u : 0u;
    }
    const uint32_t off = s_excl + woff + (a - sum);
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t idx = base + i;
        if (idx < n) out[idx] = off + v[i];
    }
}

size_t scan_scratch_bytes(int P) { return ((size_t)(P + SC_TILE - 1) / SC_TILE + 2) * 4 + 256; }

int scan_gather(int P, const uint32_t* n_ptr, const uint32_t* order, const uint32_t* tiles, uint32_t* out, void* scratch,
                cudaStream_t s) {
    if (P <= 0) return 0;
    const int ntiles = (P + SC_TILE - 1) / SC_TILE;
    OGS_CUDA(cudaMemsetAsync(scratch, 0, scan_scratch_bytes(P), s));
    uint32_t* ticket = (uint32_t*)scratch;
    uint32_t* states = (uint32_t*)((char*)scratch + 256);
    scan_gather_kernel<<<ntiles, SC_THREADS, 0, s>>>(P, n_ptr, order, tiles, out, states, ticket);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "scan_gather");
    return 0;
}

}  // namespace ogs
