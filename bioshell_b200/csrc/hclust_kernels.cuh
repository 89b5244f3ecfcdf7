// hclust_kernels.cuh -- hierarchical agglomerative clustering on the GPU, bit-compatible with
// bioshell-clustering (the consumer of the identity matrix).
//
// Replaces  HierarchicalClusteringMatrix::{new, closest_elements, update_distances,
//           replace_with_last}   bioshell-clustering/src/hierarchical/clustering_matrix.rs:11-74
//           and the merge loop of hierarchical_clustering   .../hierarchical.rs:42-77
//           with the linkage rules of                        .../strategies/mod.rs:25-92
//
// The reference rescans the whole upper triangle for every merge (O(n^3) on one core).  Here
// the n x n f32 matrix lives in HBM and each merge is two launches:
//   hclust_argmin_kernel  the closest_elements scan, all SMs, coalesced rows; HBM-bound:
//                         order^2/2 x 4 bytes per merge (n^3/6 x 4 bytes in total).  Ties resolve
//                         to the smallest j, then smallest i -- the reference's first strict
//                         minimum in (j outer, i inner) scan order.
//   hclust_merge_kernel   one CTA: final reduction, merge log, update_distances (results first,
//                         writes second, like the reference), row swap through an indirection
//                         table + column copy (replace_with_last).  The reference's quirks are
//                         kept: `sizes[j]` is passed as size_k and `sizes` is not moved by
//                         replace_with_last.
// No host round trip between merges: `order` and the step counter live in device memory.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace bsa {

struct HcBest { float v; uint32_t j, i; };

struct HcState {
    float* D;            // n x n, physical rows
    uint32_t* rmap;      // logical row -> physical row
    uint32_t* sizes;     // by logical index
    uint32_t* order;     // device scalars: order[0] = current order, order[1] = step
    HcBest* partial;     // one per argmin CTA
    float* result;       // n
    uint32_t* mat_i;     // merge log
    uint32_t* mat_j;
    float* mdist;
    uint32_t n;
    int rule;
};

__device__ __forceinline__ bool hc_better(float v, uint32_t j, uint32_t i, const HcBest& b) {
    // strict `<` in (j outer, i inner) scan order == lexicographic (v, j, i)
    return v < b.v || (v == b.v && (j < b.j || (j == b.j && i < b.i)));
}

// strategies/mod.rs:25-92 in f32 with the reference's operation order and NO fma contraction
__device__ __forceinline__ float hc_rule(int rule, uint32_t si, uint32_t sj, uint32_t sk, float dij, float dik,
                                         float djk) {
    switch (rule) {
        case 0: return dik < djk ? dik : djk;
        case 1: return dik > djk ? dik : djk;
        case 2: {
            const float d = __fdiv_rn(1.0f, (float)(si + sj));
            return __fadd_rn(__fmul_rn(__fmul_rn(d, (float)si), dik), __fmul_rn(__fmul_rn(d, (float)sj), djk));
        }
        case 3: return __fsub_rn(__fadd_rn(__fmul_rn(0.5f, dik), __fmul_rn(0.5f, djk)), __fmul_rn(0.25f, dij));
        case 4: {
            const float d = __fdiv_rn(1.0f, (float)(si + sj));
            const float a = __fadd_rn(__fmul_rn(__fmul_rn(d, (float)si), dik), __fmul_rn(__fmul_rn(d, (float)sj), djk));
            const float b = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn((float)si, (float)sj), d), d), dij);
            return __fsub_rn(a, b);
        }
        default: {
            const float d = __fdiv_rn(1.0f, (float)(si + sj + sk));
            const float a = __fadd_rn(__fmul_rn(__fmul_rn(d, (float)(si + sk)), dik),
                                      __fmul_rn(__fmul_rn(d, (float)(sj + sk)), djk));
            return __fsub_rn(a, __fmul_rn(__fmul_rn((float)sk, d), dij));
        }
    }
}

// HierarchicalClusteringMatrix::new (clustering_matrix.rs:11-22): only in[i][j], i > j is read
__global__ void hclust_init_kernel(HcState s) {
    const uint32_t n = s.n;
    const size_t total = (size_t)n * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(idx / n), j = (uint32_t)(idx - (size_t)i * n);
        if (i < j) s.D[idx] = s.D[(size_t)j * n + i];   // mirror the lower triangle upwards
        else if (i == j) s.D[idx] = 0.0f;
    }
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        s.rmap[k] = k;
        s.sizes[k] = 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { s.order[0] = n; s.order[1] = 0; }
}

// closest_elements (clustering_matrix.rs:27-42): rows are dealt round-robin over the CTAs
__global__ void __launch_bounds__(256) hclust_argmin_kernel(HcState s) {
    const uint32_t order = s.order[0];
    HcBest best{FLT_MAX, 0u, 0u};
    bool have = false;
    if (order >= 2) {
        for (uint32_t i = blockIdx.x; i + 1 < order; i += gridDim.x) {
            const float* row = s.D + (size_t)s.rmap[i] * s.n;
            for (uint32_t j = i + 1 + threadIdx.x; j < order; j += blockDim.x) {
                const float v = row[j];
                // the reference starts from f32::MAX with a strict `<`: MAX itself never wins
                if (v < FLT_MAX && (!have || hc_better(v, j, i, best))) { best = HcBest{v, j, i}; have = true; }
            }
        }
    }
    if (!have) best = HcBest{FLT_MAX, 0xffffffffu, 0xffffffffu};
    // block reduction
    __shared__ HcBest sm[256];
    sm[threadIdx.x] = best;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            const HcBest o = sm[threadIdx.x + off];
            if (hc_better(o.v, o.j, o.i, sm[threadIdx.x])) sm[threadIdx.x] = o;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) s.partial[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(1024) hclust_merge_kernel(HcState s, uint32_t n_partial) {
    __shared__ HcBest sm[1024];
    __shared__ uint32_t sh_i, sh_j, sh_order;
    const uint32_t tid = threadIdx.x;
    const uint32_t order = s.order[0];
    if (order < 2) return;
    HcBest best{FLT_MAX, 0xffffffffu, 0xffffffffu};
    for (uint32_t p = tid; p < n_partial; p += blockDim.x) {
        const HcBest o = s.partial[p];
        if (hc_better(o.v, o.j, o.i, best)) best = o;
    }
    sm[tid] = best;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if ((int)tid < off) {
            const HcBest o = sm[tid + off];
            if (hc_better(o.v, o.j, o.i, sm[tid])) sm[tid] = o;
        }
        __syncthreads();
    }
    if (tid == 0) {
        HcBest b = sm[0];
        if (b.j == 0xffffffffu) { b.i = 0; b.j = 0; }   // nothing below f32::MAX: the reference returns (0,0)
        sh_i = b.i; sh_j = b.j; sh_order = order;
    }
    __syncthreads();
    const uint32_t i = sh_i, j = sh_j;
    const uint32_t n = s.n;
    if (i == j) {   // the reference panics here (clusters.remove(&j)); stop and let the host report it
        if (tid == 0) { s.order[0] = 0; s.order[1] |= 0x80000000u; }
        return;
    }
    float* row_i = s.D + (size_t)s.rmap[i] * n;
    const float* row_j = s.D + (size_t)s.rmap[j] * n;
    const float dij = row_i[j];
    const uint32_t si = s.sizes[i], sj = s.sizes[j];
    const uint32_t step = s.order[1];
    // update_distances(i, j, rule, i): all results first (clustering_matrix.rs:64-67) ...
    for (uint32_t k = tid; k < order; k += blockDim.x)
        s.result[k] = hc_rule(s.rule, si, sj, s.sizes[j], dij, row_i[k], row_j[k]);
    __syncthreads();
    // ... then row i and column i (clustering_matrix.rs:68-71)
    for (uint32_t k = tid; k < order; k += blockDim.x) {
        const float r = s.result[k];
        row_i[k] = r;
        s.D[(size_t)s.rmap[k] * n + i] = r;
    }
    __syncthreads();
    if (tid == 0) {
        row_i[i] = 0.0f;                      // clustering_matrix.rs:72
        s.sizes[i] = sj + si;                 // :73
        s.mat_i[step] = i;
        s.mat_j[step] = j;
        s.mdist[step] = dij;
    }
    const uint32_t last = order - 1;
    if (j < last) {
        // replace_with_last(j) (clustering_matrix.rs:47-53): swap the ROWS, copy the column
        __syncthreads();
        if (tid == 0) {
            const uint32_t t = s.rmap[j];
            s.rmap[j] = s.rmap[last];
            s.rmap[last] = t;
        }
        __syncthreads();
        for (uint32_t r = tid; r < last; r += blockDim.x) {
            float* row = s.D + (size_t)s.rmap[r] * n;
            row[j] = row[last];
        }
    }
    __syncthreads();
    if (tid == 0) { s.order[0] = last; s.order[1] = step + 1; }
}

}  // namespace bsa
