/*
 * hclust_oracle.c -- CPU ORACLE for the "next" row SURVEY.md 8(f)-2: hierarchical
 * agglomerative clustering as bioshell-clustering does it.  TEST INFRASTRUCTURE (same rules
 * as bioshell_oracle.c: only tests/ and the benchmark's CPU-baseline leg may load it).
 *
 * Literal restatement of
 *   HierarchicalClusteringMatrix   bioshell-clustering/src/hierarchical/clustering_matrix.rs:1-75
 *   hierarchical_clustering        bioshell-clustering/src/hierarchical/hierarchical.rs:22-80
 *   linkage rules                  bioshell-clustering/src/hierarchical/strategies/mod.rs:25-92
 * including its quirks: the full O(n^2) `closest_elements` scan per merge with the
 * first-strict-minimum tie order (j outer, i inner), `sizes[j]` passed as size_k
 * (clustering_matrix.rs:66), row SWAP + column COPY in `replace_with_last`.
 * Parity status: PINNED by the reference's tests cluster_numbers / cluster_letters
 * (bioshell-clustering/tests/test_hierarchical.rs:9-41; tests/golden/ref_kats.json).
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_SINGLE = 0, ORC_COMPLETE = 1, ORC_AVERAGE = 2, ORC_MEDIAN = 3, ORC_CENTROID = 4, ORC_WARD = 5 };

/* strategies/mod.rs:25-92 -- f32 arithmetic in the reference's operation order */
static float merge_rule(int rule, size_t si, size_t sj, size_t sk, float dij, float dik, float djk) {
    switch (rule) {
        case ORC_SINGLE: return dik < djk ? dik : djk;            /* f32::min (no NaNs here) */
        case ORC_COMPLETE: return dik > djk ? dik : djk;
        case ORC_AVERAGE: {
            float d = 1.0f / (float)(si + sj);
            return d * (float)si * dik + d * (float)sj * djk;
        }
        case ORC_MEDIAN: return 0.5f * dik + 0.5f * djk - 0.25f * dij;
        case ORC_CENTROID: {
            float d = 1.0f / (float)(si + sj);
            return d * (float)si * dik + d * (float)sj * djk - (float)si * (float)sj * d * d * dij;
        }
        default: {
            float d = 1.0f / (float)(si + sj + sk);
            return d * (float)(si + sk) * dik + d * (float)(sj + sk) * djk - (float)sk * d * dij;
        }
    }
}

/*
 * dist: n x n row-major; like HierarchicalClusteringMatrix::new only dist[i*n+j] with i > j is
 * read and mirrored (clustering_matrix.rs:14-19).  Outputs per merge step s = 0..n-2:
 * mat_i/mat_j (the matrix indices closest_elements returned, i < j), id_i/id_j (cluster ids:
 * leaves 0..n-1, merge s creates id n+s), merge_dist.  Returns the number of merges.
 */
int64_t orc_hclust(uint32_t n, const float *dist, int rule, uint32_t *mat_i, uint32_t *mat_j,
                   uint32_t *id_i, uint32_t *id_j, float *merge_dist) {
    if (n == 0) return -1;
    float **m = (float **)malloc(sizeof(float *) * n);
    size_t *sizes = (size_t *)malloc(sizeof(size_t) * n);
    uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * n);   /* the `clusters` HashMap: index -> id */
    float *result = (float *)malloc(sizeof(float) * n);
    if (!m || !sizes || !ids || !result) return -2;
    for (uint32_t i = 0; i < n; ++i) {
        m[i] = (float *)calloc(n, sizeof(float));
        if (!m[i]) return -2;
        sizes[i] = 1;
        ids[i] = i;
    }
    for (uint32_t i = 1; i < n; ++i)
        for (uint32_t j = 0; j < i; ++j) {
            float v = dist[(size_t)i * n + j];
            m[i][j] = v;
            m[j][i] = v;
        }
    uint32_t order = n, next_id = n;
    int64_t step = 0;
    while (order > 1) {                                   /* hierarchical.rs:42 (clusters.len() > 1) */
        float best = FLT_MAX;                             /* clustering_matrix.rs:27-42 */
        uint32_t bi = 0, bj = 0;
        for (uint32_t j = 1; j < order; ++j)
            for (uint32_t i = 0; i < j; ++i)
                if (m[i][j] < best) { best = m[i][j]; bi = i; bj = j; }
        const uint32_t i = bi, j = bj;
        if (i == j) { step = -3; break; }                 /* clusters.remove(&j) would panic */
        const float dij = m[i][j];                        /* hierarchical.rs:52 */
        mat_i[step] = i; mat_j[step] = j;
        id_i[step] = ids[i]; id_j[step] = ids[j];
        merge_dist[step] = dij;
        ids[i] = next_id;                                 /* :61 clusters.insert(i, c) */
        /* update_distances(i, j, rule, i)  clustering_matrix.rs:55-74 */
        const size_t si = sizes[i], sj = sizes[j];
        for (uint32_t k = 0; k < order; ++k)
            result[k] = merge_rule(rule, si, sj, sizes[j], m[i][j], m[i][k], m[j][k]);
        for (uint32_t k = 0; k < order; ++k) { m[i][k] = result[k]; m[k][i] = result[k]; }
        m[i][i] = 0.0f;
        sizes[i] = sj + si;
        /* hierarchical.rs:65-74 */
        const uint32_t last = order - 1;
        if (j < last) {
            ids[j] = ids[last];
            order -= 1;                                   /* replace_with_last: clustering_matrix.rs:47-53 */
            float *t = m[j]; m[j] = m[order]; m[order] = t;
            for (uint32_t r = 0; r < order; ++r) m[r][j] = m[r][order];
            /* NOTE: `sizes` is NOT moved by the reference (clustering_matrix.rs:47-53) */
        } else {
            order -= 1;
        }
        ++next_id;
        ++step;
    }
    for (uint32_t i = 0; i < n; ++i) free(m[i]);
    free(m); free(sizes); free(ids); free(result);
    return step;
}
