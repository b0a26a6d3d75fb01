/*
 * oracle/kmeans_oracle.c -- CPU restatement of the codebook assign / centroid-sum steps of
 * OpenGaussian's two-level k-means (reference: scene/kmeans_quantize.py, class
 * Quantize_kMeans: get_dist :38-55, update_centers_ :82-87, cluster_assign :146-241).
 *
 * THIS FILE IS TEST INFRASTRUCTURE (see oracle/README.md).  The product never calls it.
 *
 * Pinning: the reference file itself is importable in the build container; the script
 * tests/golden/make_kmeans_golden.py runs it (torch CPU) and commits centres + ids as
 * fixtures; tests/test_kmeans_oracle.py checks this restatement against those fixtures.
 *
 * Arithmetic contract shared with the CUDA kernels (so ids are bit-exact) -- the matmul form
 * that torch.cdist (:51-54) itself takes for these sizes, with ||x||^2 (constant per point) dropped:
 *   cn_j    = fold_{d=0..D-1} fmaf(c_jd, c_jd, acc), acc0 = 0                  (fp32)
 *   s_j     = fold_{d=0..D-1} fmaf(x_d, c_jd, acc),  acc0 = 0
 *   score_j = fmaf(-2, s_j, cn_j)        ( = ||x - c_j||^2 - ||x||^2 )
 *   id = lowest index of the minimum score (strict '<' scan) -- torch.argmin tie rule (:182).
 *   The reference takes sqrt(clamp(.)) of the same expansion evaluated by a BLAS GEMM in an
 *   unspecified order; sqrt does not change the argmin except on near-ties, which the pin test counts.
 *   A point is the concatenation [a (Da floats) | b (Db floats) * scale_b] -- the reference
 *   builds cat(_ins_feat, _xyz * pos_weight) (:254-257); the product is a single fp32 multiply.
 *   Centroid sums: points are split into consecutive groups of `group` points; inside a
 *   group sums are accumulated sequentially in fp32 in index order; group partials are then
 *   accumulated sequentially in fp32 in group order.  (The reference sums by a one-hot GEMM per
 *   10000-point chunk, :84,:187 -- an unspecified BLAS order; equal to ~1e-6 relative.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float point_coord(const float* a, int Da, const float* b, int Db, float scale_b,
                                int64_t i, int d) {
    return d < Da ? a[i * Da + d] : b[i * Db + (d - Da)] * scale_b;
}

/*
 * select_ids/selected: when select_ids != NULL only points with select_ids[i] == selected are
 * assigned (leaf mode, :196-206); the others keep ids_out[i] untouched.  id_offset is added to
 * the winning index (leaf ids are start_id + argmin).
 */
int ogs_oracle_kmeans_assign(int64_t N, const float* a, int Da, const float* b, int Db,
                             float scale_b, const float* centers, int k,
                             const int64_t* select_ids, int64_t selected, int64_t id_offset,
                             int64_t* ids_out) {
    const int D = Da + Db;
    float x[64];
    if (D > 64) return -1;
    float* cn = (float*)malloc(sizeof(float) * (size_t)(k > 0 ? k : 1));
    for (int j = 0; j < k; j++) {
        float acc = 0.f;
        for (int d = 0; d < D; d++) acc = fmaf(centers[j * D + d], centers[j * D + d], acc);
        cn[j] = acc;
    }
    for (int64_t i = 0; i < N; i++) {
        if (select_ids && select_ids[i] != selected) continue;
        for (int d = 0; d < D; d++) x[d] = point_coord(a, Da, b, Db, scale_b, i, d);
        float best = INFINITY;
        int best_j = 0;
        for (int j = 0; j < k; j++) {
            float acc = 0.f;
            for (int d = 0; d < D; d++) acc = fmaf(x[d], centers[j * D + d], acc);
            acc = fmaf(-2.0f, acc, cn[j]);
            if (acc < best) { best = acc; best_j = j; }
        }
        ids_out[i] = id_offset + best_j;
    }
    free(cn);
    return 0;
}

/* sums [k][D], counts [k] (as float, exact integers) for ids in [id_offset, id_offset+k). */
int ogs_oracle_kmeans_accumulate(int64_t N, const float* a, int Da, const float* b, int Db,
                                 float scale_b, int k, const int64_t* ids,
                                 const int64_t* select_ids, int64_t selected, int64_t id_offset,
                                 int group, float* sums, float* counts) {
    const int D = Da + Db;
    float* part = (float*)malloc(sizeof(float) * (size_t)k * (D + 1));
    memset(sums, 0, sizeof(float) * (size_t)k * D);
    memset(counts, 0, sizeof(float) * (size_t)k);
    for (int64_t g0 = 0; g0 < N; g0 += group) {
        int64_t g1 = g0 + group < N ? g0 + group : N;
        memset(part, 0, sizeof(float) * (size_t)k * (D + 1));
        for (int64_t i = g0; i < g1; i++) {
            if (select_ids && select_ids[i] != selected) continue;
            int64_t j = ids[i] - id_offset;
            if (j < 0 || j >= k) continue;
            for (int d = 0; d < D; d++) part[j * (D + 1) + d] += point_coord(a, Da, b, Db, scale_b, i, d);
            part[j * (D + 1) + D] += 1.0f;
        }
        for (int j = 0; j < k; j++) {
            for (int d = 0; d < D; d++) sums[j * D + d] += part[j * (D + 1) + d];
            counts[j] += part[j * (D + 1) + D];
        }
    }
    free(part);
    return 0;
}
