// kmeans_seg.cu -- fine level of the two-level codebook for ALL coarse clusters in one launch
// (ogs_kmeans_assign_segmented).
//
// The reference runs its leaf mode once per coarse cluster (scene/kmeans_quantize.py:196-206 assign,
// :233-238 reassign; driven by train.py:322-332): every pass builds `self.cls_ids == idx_c` over all N
// points and gathers the members.  The per-cluster Lloyd iterations are independent of each other, so here
// every point picks the centre block of its own coarse cluster: rows [c*k2, c*k2 + seg_k[c]) of
// `seg_centers` (= leaf_centers, seg_k = iLeafSubNum), and the launch does the work of k1 leaf calls.
//
// Arithmetic contract for the ids: the same fmaf chain as kmeans.cu / oracle/kmeans_oracle.c
// (cn_j, s_j ascending in d, score = fmaf(-2, s_j, cn_j), strict '<': lowest index wins), so the ids are
// bit-identical to k1 per-cluster calls of the oracle's assign.
//
// Centroid sums are EXACT: every coordinate is converted to a 64-bit fixed-point integer
// (llrint(x * 2^fix_bits), |x| * 2^fix_bits < 2^32) and accumulated with integer atomics -- shared memory first,
// one row per (coarse, fine) pair, then integer reds into one global table.  Integer addition is
// associative, so the result does not depend on the order of the atomics, on the grid, or on how the points
// are sharded over GPUs: a sharded run reproduces the single-GPU centres bit for bit after the (integer)
// all-reduce.  The kernel is HBM-bound: 4 D + 8 (coarse id) + 8 (fine id) bytes per point.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "peer.cuh"

namespace ogs {

#define KS_THREADS 512
#define KS_PPT 4
#define KS_FLUSH_POINTS 16384    // a CTA folds its 32-bit shared accumulators into its 64-bit partial table this often

// Every CTA adds its sums into ONE global int64 table with red.global.add.u64 (zeros skipped) -- integer reds commute,
// so neither determinism nor exactness is lost -- about once per KS_FLUSH_POINTS points.
// Shared memory holds the centre blocks structure-of-arrays, s_c[d][c * k2p + j] with an ODD block stride k2p, so
// that lanes working on different coarse clusters spread over all 32 banks, and the accumulators as PAIRS of 32-bit
// words: a fixed-point value q (int64) is split as q = hi * 65536 + lo, lo in [0, 65535], and lo / hi are added
// with native 32-bit shared atomics (64-bit shared atomics are compare-and-swap loops).  Within KS_FLUSH_POINTS
// points neither word can overflow provided |x| * 2^fix_bits < 2^32 (the caller's contract); the CTA then folds
// the pairs into its own row of the int64 partial table in global memory (plain read-modify-write: the row is
// private to the CTA) and clears them.
// Optional tail (as in kmeans.cu): the LAST CTA to finish sums the per-CTA int64 partial tables, all-reduces the
// [k1*k2][D+1] integers over the peers' memory when the points are sharded, and applies the reference's centre update
// to every block -- one Lloyd pass of ALL coarse clusters' fine levels is then one launch.
struct KsTail {
    int mode;                    // 0: off (the sums are only ADDED into the caller's table)
    unsigned int* ticket;
    int has_peer;
    PeerDev peer;
    float* counts_state;         // [k1*k2]
    float eps_add, inv_scale;
    float* centers_out;          // [k1*k2][D]
};

template <int D>
__global__ void __launch_bounds__(KS_THREADS) kmeans_assign_seg_kernel(
    int64_t N, const float* __restrict__ a, const int64_t* __restrict__ coarse_ids, const float* __restrict__ seg_centers,
    const int32_t* __restrict__ seg_k, int k1, int k2, int64_t* __restrict__ ids_out,
    unsigned long long* __restrict__ table /* [k1*k2][D+1] global int64 sums, ADDED to; NULL = pure reassign */,
    float fix_scale, int vec2_ok, KsTail tail) {
    constexpr int ROW = D + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int rows = k1 * k2;
    const int k2p = k2 | 1;
    const int rows_p = k1 * k2p;
    // accumulators first (8-byte aligned: the fused tail re-reads the two 32-bit planes' storage as int64)
    unsigned int* s_lo = reinterpret_cast<unsigned int*>(smem_raw);       // [rows][ROW]      (if table)
    int* s_hi = reinterpret_cast<int*>(s_lo + (size_t)rows * ROW);        // [rows][ROW]
    float* s_c = reinterpret_cast<float*>(smem_raw + (table ? (size_t)rows * ROW * 8 : 0));   // [D + 1][rows_p] (last plane: ||c||^2)
    int* s_k = reinterpret_cast<int*>(s_c + (size_t)ROW * rows_p);        // [k1]

    for (int r = threadIdx.x; r < rows; r += KS_THREADS) {
        const int c = r / k2, j = r - c * k2;
        float cn = 0.f;
#pragma unroll
        for (int d = 0; d < D; d++) {
            const float v = seg_centers[(size_t)r * D + d];
            s_c[d * rows_p + c * k2p + j] = v;
            cn = __fmaf_rn(v, v, cn);
        }
        s_c[D * rows_p + c * k2p + j] = cn;
    }
    for (int j = threadIdx.x; j < k1; j += KS_THREADS) {
        const int v = seg_k[j];
        s_k[j] = v < 0 ? 0 : (v > k2 ? k2 : v);
    }
    if (table)
        for (int e = threadIdx.x; e < rows * ROW; e += KS_THREADS) { s_lo[e] = 0u; s_hi[e] = 0; }
    __syncthreads();

    // fold the CTA's 32-bit pairs into the global int64 table: integer reds, order-independent, zeros skipped
    auto flush = [&]() {
        __syncthreads();
        for (int e = threadIdx.x; e < rows * ROW; e += KS_THREADS) {
            const long long v = (long long)s_hi[e] * 65536ll + (long long)s_lo[e];
            if (v != 0) atomicAdd(table + e, (unsigned long long)v);
            s_lo[e] = 0u;
            s_hi[e] = 0;
        }
        __syncthreads();
    };

    const int64_t chunk = (int64_t)KS_THREADS * KS_PPT;
    int since_flush = 0;
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < N; base += (int64_t)gridDim.x * chunk) {
        float x[KS_PPT][D];
        int c[KS_PPT];
#pragma unroll
        for (int q = 0; q < KS_PPT; q++) {                 // all loads of the 4 points first: bytes in flight
            const int64_t i = base + q * KS_THREADS + threadIdx.x;
            c[q] = -1;
            if (i < N) {
                const int64_t cc = __ldg(coarse_ids + i);
                if (cc >= 0 && cc < k1 && s_k[cc] > 0) c[q] = (int)cc;
            }
            if ((D % 2 == 0) && vec2_ok) {                 // rows are 8-byte aligned: D/2 float2 loads per point
#pragma unroll
                for (int d = 0; d < D / 2; d++) {
                    const float2 v = (c[q] >= 0) ? __ldg(reinterpret_cast<const float2*>(a + i * D) + d) : make_float2(0.f, 0.f);
                    x[q][2 * d] = v.x;
                    x[q][2 * d + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int d = 0; d < D; d++) x[q][d] = (c[q] >= 0) ? __ldg(a + i * D + d) : 0.f;
            }
        }
#pragma unroll
        for (int q = 0; q < KS_PPT; q++) {
            if (c[q] < 0) continue;
            const int64_t i = base + q * KS_THREADS + threadIdx.x;
            const int n = s_k[c[q]];
            const float* blk = s_c + c[q] * k2p;
            float best = INFINITY;
            int best_j = 0;
            for (int j = 0; j < n; j++) {
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < D; d++) acc = __fmaf_rn(x[q][d], blk[d * rows_p + j], acc);
                acc = __fmaf_rn(-2.0f, acc, blk[D * rows_p + j]);
                if (acc < best) { best = acc; best_j = j; }
            }
            const int r = c[q] * k2 + best_j;
            ids_out[i] = (int64_t)r;
            if (table) {
#pragma unroll
                for (int d = 0; d < D; d++) {
                    const long long v = __float2ll_rn(__fmul_rn(x[q][d], fix_scale));
                    atomicAdd(s_lo + r * ROW + d, (unsigned int)(v & 0xFFFFll));
                    atomicAdd(s_hi + r * ROW + d, (int)(v >> 16));
                }
                atomicAdd(s_lo + r * ROW + D, 1u);
            }
        }
        since_flush += (int)chunk;
        if (table && since_flush >= KS_FLUSH_POINTS) { flush(); since_flush = 0; }
    }
    if (table) flush();
    if (tail.mode && table) {
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(tail.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        long long* vec = reinterpret_cast<long long*>(s_lo);        // [rows][ROW] int64: exactly the two 32-bit planes
        for (int e = threadIdx.x; e < rows * ROW; e += KS_THREADS) {
            vec[e] = (long long)__ldcg(table + e);
            table[e] = 0ull;                                          // ready for the next pass
        }
        __syncthreads();
        bool ok = true;
        if (tail.has_peer) ok = peer_allreduce_cta<long long>(tail.peer, vec, rows * ROW);
        if (ok) {
            for (int r = threadIdx.x; r < rows; r += KS_THREADS) {
                const float cnt = __fadd_rn(tail.counts_state[r], __fadd_rn((float)vec[r * ROW + D], tail.eps_add));
#pragma unroll
                for (int d = 0; d < D; d++)
                    tail.centers_out[(size_t)r * D + d] = __fdiv_rn((float)((double)vec[r * ROW + d] * (double)tail.inv_scale), cnt);
                tail.counts_state[r] = cnt > 0.1f ? 0.f : cnt;
            }
        }
        if (threadIdx.x == 0) *tail.ticket = 0u;
    }
}

static int seg_grid(int64_t N, size_t smem) {
    const int64_t want = (N + KS_THREADS * KS_PPT - 1) / (KS_THREADS * KS_PPT);
    // every CTA ends with up to k1*k2*(D+1) integer reds whatever its share of the points: few fat CTAs (512 threads,
    // 4 points each in flight per thread) keep that fixed cost small without starving the memory system
    static const int env_per_sm = getenv("OGS_KS_PER_SM") ? atoi(getenv("OGS_KS_PER_SM")) : 0;   // tuning knob
    const int per_sm = env_per_sm > 0 ? env_per_sm : 2;
    (void)smem;
    int grid = (int)(want < (int64_t)OGS_NUM_SMS * per_sm ? want : (int64_t)OGS_NUM_SMS * per_sm);
    return grid < 1 ? 1 : grid;
}

// one attribute cache per kernel instantiation, shared by both launchers (see kmeans.cu)
template <int D>
static int ensure_seg_smem(size_t smem) {
    static std::atomic<size_t> attr[OGS_MAX_DEVICES];
    std::atomic<size_t>& at = attr[current_device()];
    if (smem > 48 * 1024 && smem > at.load(std::memory_order_relaxed)) {
        OGS_CUDA(cudaFuncSetAttribute(kmeans_assign_seg_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        at.store(smem, std::memory_order_relaxed);
    }
    return 0;
}

template <int D>
static int launch_seg_d(int64_t N, const float* a, const int64_t* coarse_ids, const float* seg_centers, const int32_t* seg_k,
                        int k1, int k2, int64_t* ids_out, int64_t* acc, int fix_bits, cudaStream_t s) {
    const int rows = k1 * k2;
    const size_t smem = (acc ? (size_t)rows * (D + 1) * 8 : 0) + (size_t)k1 * (k2 | 1) * (D + 1) * 4 + (size_t)k1 * 4;
    if (smem > 200 * 1024) { set_error("kmeans_assign_segmented: k1*k2=%d D=%d needs %zu B shared memory", rows, D, smem); return -5; }
    { const int rc_attr = ensure_seg_smem<D>(smem); if (rc_attr) return rc_attr; }
    const int grid = seg_grid(N, smem);
    KsTail tail;
    memset(&tail, 0, sizeof tail);
    kmeans_assign_seg_kernel<D><<<grid, KS_THREADS, smem, s>>>(N, a, coarse_ids, seg_centers, seg_k, k1, k2, ids_out,
                                                               (unsigned long long*)acc, ldexpf(1.0f, fix_bits),
                                                               (((uintptr_t)a & 7) == 0) ? 1 : 0, tail);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "kmeans_assign_segmented");
    return 0;
}

int launch_kmeans_assign_segmented(int64_t N, const float* a, int D, const int64_t* coarse_ids, const float* seg_centers,
                                   const int32_t* seg_k, int k1, int k2, int64_t* ids_out, int64_t* acc, int fix_bits,
                                   cudaStream_t s) {
#define OGS_KS_CASE(DD) case DD: return launch_seg_d<DD>(N, a, coarse_ids, seg_centers, seg_k, k1, k2, ids_out, acc, fix_bits, s);
    switch (D) {
        OGS_KS_CASE(1) OGS_KS_CASE(2) OGS_KS_CASE(3) OGS_KS_CASE(4) OGS_KS_CASE(5) OGS_KS_CASE(6) OGS_KS_CASE(7)
        OGS_KS_CASE(8) OGS_KS_CASE(9) OGS_KS_CASE(10) OGS_KS_CASE(11) OGS_KS_CASE(12)
    }
#undef OGS_KS_CASE
    set_error("kmeans_assign_segmented: unsupported point dimension %d (1..12)", D);
    return -5;
}

// ---- one Lloyd pass of every coarse cluster's fine level in one launch ----
size_t kmeans_seg_lloyd_workspace_bytes(int k1, int k2, int D) { return 256 + (size_t)k1 * k2 * (D + 1) * 8; }

template <int D>
static int launch_seg_lloyd_d(int64_t N, const float* a, const int64_t* coarse_ids, float* seg_centers, const int32_t* seg_k,
                              int k1, int k2, int64_t* ids_out, int fix_bits, float* counts_state, float eps_add,
                              const ogs_peer_comm* comm, void* workspace, cudaStream_t s) {
    const int rows = k1 * k2;
    const size_t smem = (size_t)rows * (D + 1) * 8 + (size_t)k1 * (k2 | 1) * (D + 1) * 4 + (size_t)k1 * 4;
    if (smem > 200 * 1024) { set_error("kmeans_lloyd_pass_segmented: k1*k2=%d D=%d needs %zu B shared memory", rows, D, smem); return -5; }
    if (comm && (size_t)rows * (D + 1) * 8 > peer_comm_slot_bytes(comm)) {
        set_error("kmeans_lloyd_pass_segmented: the communicator's slots are too small");
        return -1;
    }
    { const int rc_attr = ensure_seg_smem<D>(smem); if (rc_attr) return rc_attr; }
    const int grid = seg_grid(N, smem);
    KsTail tail;
    memset(&tail, 0, sizeof tail);
    tail.mode = 1;
    tail.ticket = (unsigned int*)workspace;
    tail.has_peer = comm ? 1 : 0;
    if (comm) tail.peer = *peer_comm_dev(comm);
    tail.counts_state = counts_state; tail.eps_add = eps_add; tail.inv_scale = ldexpf(1.0f, -fix_bits);
    tail.centers_out = seg_centers;
    kmeans_assign_seg_kernel<D><<<grid, KS_THREADS, smem, s>>>(N, a, coarse_ids, seg_centers, seg_k, k1, k2, ids_out,
                                                               (unsigned long long*)((char*)workspace + 256),
                                                               ldexpf(1.0f, fix_bits), (((uintptr_t)a & 7) == 0) ? 1 : 0, tail);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "kmeans_lloyd_pass_segmented");
    return 0;
}

int launch_kmeans_seg_lloyd_pass(int64_t N, const float* a, int D, const int64_t* coarse_ids, float* seg_centers,
                                 const int32_t* seg_k, int k1, int k2, int64_t* ids_out, int fix_bits, float* counts_state,
                                 float eps_add, const ogs_peer_comm* comm, void* workspace, cudaStream_t s) {
#define OGS_KSL_CASE(DD) case DD: return launch_seg_lloyd_d<DD>(N, a, coarse_ids, seg_centers, seg_k, k1, k2, ids_out, fix_bits, counts_state, eps_add, comm, workspace, s);
    switch (D) {
        OGS_KSL_CASE(1) OGS_KSL_CASE(2) OGS_KSL_CASE(3) OGS_KSL_CASE(4) OGS_KSL_CASE(5) OGS_KSL_CASE(6) OGS_KSL_CASE(7)
        OGS_KSL_CASE(8) OGS_KSL_CASE(9) OGS_KSL_CASE(10) OGS_KSL_CASE(11) OGS_KSL_CASE(12)
    }
#undef OGS_KSL_CASE
    set_error("kmeans_lloyd_pass_segmented: unsupported point dimension %d (1..12)", D);
    return -5;
}

// One Lloyd update of every block from the exact sums, with the reference's count bookkeeping
// (scene/kmeans_quantize.py:167,186,208-214 per leaf call): counts_state += count + eps_add;
// centre = sum / counts_state; counts_state is zeroed where it exceeds 0.1.  Rows at or beyond seg_k[c] of a
// block have sum 0 and become 0 / eps = 0 (the reference rewrites all k2 rows, :211).
__global__ void kmeans_seg_finalize_kernel(int rows, int D, const long long* __restrict__ acc, float inv_scale,
                                           float eps_add, float* __restrict__ counts_state, float* __restrict__ centers) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const long long* row = acc + (size_t)r * (D + 1);
    const float cnt = __fadd_rn(counts_state[r], __fadd_rn((float)row[D], eps_add));
    for (int d = 0; d < D; d++) {
        const float sum = (float)((double)row[d] * (double)inv_scale);
        centers[(size_t)r * D + d] = __fdiv_rn(sum, cnt);
    }
    counts_state[r] = cnt > 0.1f ? 0.f : cnt;
}

int launch_kmeans_seg_finalize(int rows, int D, const int64_t* acc, int fix_bits, float eps_add, float* counts_state,
                               float* centers, cudaStream_t s) {
    if (rows <= 0) return 0;
    kmeans_seg_finalize_kernel<<<(rows + 127) / 128, 128, 0, s>>>(rows, D, (const long long*)acc, ldexpf(1.0f, -fix_bits),
                                                                  eps_add, counts_state, centers);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "kmeans_seg_finalize");
    return 0;
}

}  // namespace ogs
