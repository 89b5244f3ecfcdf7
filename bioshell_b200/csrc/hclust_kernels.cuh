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
//
// NEAREST-NEIGHBOUR CACHE (round 2).  The scan above is n^3/6 x 4 bytes over a clustering: 117 s for
// 100,000 points at 0.88 of the HBM roofline -- fast for what it does, but it does far too much.  The
// first strict minimum in (j outer, i inner) scan order is the lexicographic minimum of (value, j, i),
// and for a fixed row i that is the row's (value, j) minimum over j > i: so the scan is replaced by a
// per-row cache of that minimum.  closest_elements becomes a reduction over `order` cache entries
// (hclust_nn_argmin_kernel), and a merge of (i, j) touches, in every other row k, only the columns
// i, j (which receives column `last`) and `last` (which disappears): thread k of the merge kernel holds
// the new values of exactly those cells, so it can update its row's entry in O(1) -- the best new
// candidate wins if it is at least as good as the cached one; if the cached cell itself was one of
// the three and no candidate matches it, the row is queued and rebuilt (hclust_rowmin_kernel) together
// with rows i and j, which change completely.  Same merges, same ties, same log as the full scan
// (tests/test_gpu_hclust.py runs both against the oracle); 100,000 points: seconds instead of minutes.
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
    uint32_t pitch;      // floats per physical row (multiple of 32: rows are 128-byte aligned)
    int rule;
    // nearest-neighbour cache (null: every merge rescans the whole triangle): per logical row i the
    // lexicographic (value, j) minimum over the columns j > i, i.e. the row's best cell in scan order
    float* rmin_v;
    uint32_t* rmin_j;
    uint32_t* todo;      // rows whose cache entry must be rebuilt after the current merge
    uint32_t* todo_n;
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
__global__ void hclust_init_kernel(HcState s, const float* __restrict__ in) {
    const uint32_t n = s.n;
    const size_t total = (size_t)n * s.pitch;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(idx / s.pitch), j = (uint32_t)(idx - (size_t)i * s.pitch);
        float v = 0.0f;
        if (j < n && i != j) v = i > j ? in[(size_t)i * n + j] : in[(size_t)j * n + i];   // mirror the lower triangle
        s.D[idx] = v;
    }
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        s.rmap[k] = k;
        s.sizes[k] = 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { s.order[0] = n; s.order[1] = 0; s.order[2] = 0; s.order[3] = 0; }
}

// closest_elements (clustering_matrix.rs:27-42).  HBM-bound: every warp takes whole rows
// (round-robin), reads the part right of the diagonal with 16-byte loads, four of them in
// flight per lane, and keeps the lexicographic (value, j, i) minimum.
__device__ __forceinline__ void hc_consider(float v, uint32_t j, uint32_t i, HcBest& b) {
    // the reference starts from f32::MAX with a strict `<`: MAX itself never wins
    if (v < FLT_MAX && hc_better(v, j, i, b)) b = HcBest{v, j, i};
}

__global__ void __launch_bounds__(256) hclust_argmin_kernel(HcState s) {
    const uint32_t order = s.order[0];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    HcBest best{FLT_MAX, 0xffffffffu, 0xffffffffu};
    if (order >= 2) {
        for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + warp; i + 1 < order; i += nwarps) {
            const float* row = s.D + (size_t)s.rmap[i] * s.pitch;
            const uint32_t j0 = i + 1;
            const uint32_t ja = (j0 + 3u) & ~3u;                  // first 16-byte aligned column
            const uint32_t jb = order & ~3u;                      // end of the aligned body
            if (ja >= jb) {
                for (uint32_t j = j0 + lane; j < order; j += 32) hc_consider(row[j], j, i, best);
                continue;
            }
            if (j0 + lane < ja) hc_consider(row[j0 + lane], j0 + lane, i, best);
            if (jb + lane < order) hc_consider(row[jb + lane], jb + lane, i, best);
            const float4* r4 = reinterpret_cast<const float4*>(row);
            uint32_t q = (ja >> 2) + lane;
            const uint32_t qe = jb >> 2;
            for (; q + 96 < qe; q += 128) {
                const float4 a = r4[q], b = r4[q + 32], c = r4[q + 64], d = r4[q + 96];
                const float4 v[4] = {a, b, c, d};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t j = (q + 32 * u) << 2;
                    hc_consider(v[u].x, j, i, best); hc_consider(v[u].y, j + 1, i, best);
                    hc_consider(v[u].z, j + 2, i, best); hc_consider(v[u].w, j + 3, i, best);
                }
            }
            for (; q < qe; q += 32) {
                const float4 a = r4[q];
                const uint32_t j = q << 2;
                hc_consider(a.x, j, i, best); hc_consider(a.y, j + 1, i, best);
                hc_consider(a.z, j + 2, i, best); hc_consider(a.w, j + 3, i, best);
            }
        }
    }
    // warp, then block reduction
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        HcBest o;
        o.v = __shfl_down_sync(0xffffffffu, best.v, off);
        o.j = __shfl_down_sync(0xffffffffu, best.j, off);
        o.i = __shfl_down_sync(0xffffffffu, best.i, off);
        if (hc_better(o.v, o.j, o.i, best)) best = o;
    }
    __shared__ HcBest sm[8];
    if (lane == 0) sm[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        HcBest b = sm[0];
        for (uint32_t w = 1; w < (blockDim.x >> 5); ++w)
            if (hc_better(sm[w].v, sm[w].j, sm[w].i, b)) b = sm[w];
        s.partial[blockIdx.x] = b;
    }
}

__device__ __forceinline__ bool hc_key_less(float v, uint32_t j, float bv, uint32_t bj) {   // (v, j) < (bv, bj)
    return v < bv || (v == bv && j < bj);
}

// Row cache (re)build: one CTA per row, the part right of the diagonal with 16-byte loads, four in
// flight per lane.  all != 0: every row (once, after init); else the rows queued by the last merge.
__global__ void __launch_bounds__(1024) hclust_rowmin_kernel(HcState s, int all) {
    __shared__ float sv[32];
    __shared__ uint32_t sj[32];
    const uint32_t order = s.order[0];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const uint32_t count = all ? (order > 0 ? order : 0u) : *s.todo_n;
    for (uint32_t idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const uint32_t i = all ? idx : s.todo[idx];
        float bv = FLT_MAX;
        uint32_t bj = 0xffffffffu;
        if (i + 1 < order) {
            const float* row = s.D + (size_t)s.rmap[i] * s.pitch;
            const uint32_t j0 = i + 1;
            const uint32_t ja = (j0 + 3u) & ~3u;                  // first 16-byte aligned column
            const uint32_t jb = order & ~3u;                      // end of the aligned body
#define BSA_HC_SEE(V, J) { const float v_ = (V); const uint32_t j_ = (J); if (v_ < FLT_MAX && hc_key_less(v_, j_, bv, bj)) { bv = v_; bj = j_; } }
            if (ja >= jb) {
                for (uint32_t j = j0 + threadIdx.x; j < order; j += blockDim.x) BSA_HC_SEE(row[j], j)
            } else {
                if (j0 + threadIdx.x < ja) BSA_HC_SEE(row[j0 + threadIdx.x], j0 + threadIdx.x)
                if (jb + threadIdx.x < order) BSA_HC_SEE(row[jb + threadIdx.x], jb + threadIdx.x)
                const float4* r4 = reinterpret_cast<const float4*>(row);
                const uint32_t qe = jb >> 2, stride = blockDim.x;
                uint32_t q = (ja >> 2) + threadIdx.x;
                for (; q + 3 * stride < qe; q += 4 * stride) {
                    const float4 v[4] = {r4[q], r4[q + stride], r4[q + 2 * stride], r4[q + 3 * stride]};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t j = (q + u * stride) << 2;
                        BSA_HC_SEE(v[u].x, j) BSA_HC_SEE(v[u].y, j + 1) BSA_HC_SEE(v[u].z, j + 2) BSA_HC_SEE(v[u].w, j + 3)
                    }
                }
                for (; q < qe; q += stride) {
                    const float4 a = r4[q];
                    const uint32_t j = q << 2;
                    BSA_HC_SEE(a.x, j) BSA_HC_SEE(a.y, j + 1) BSA_HC_SEE(a.z, j + 2) BSA_HC_SEE(a.w, j + 3)
                }
            }
#undef BSA_HC_SEE
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_down_sync(0xffffffffu, bv, off);
            const uint32_t oj = __shfl_down_sync(0xffffffffu, bj, off);
            if (hc_key_less(ov, oj, bv, bj)) { bv = ov; bj = oj; }
        }
        __syncthreads();                       // the previous row's reduction is done with sv / sj
        if (lane == 0) { sv[warp] = bv; sj[warp] = bj; }
        __syncthreads();
        if (warp == 0) {
            bv = lane < nwarp ? sv[lane] : FLT_MAX;
            bj = lane < nwarp ? sj[lane] : 0xffffffffu;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_down_sync(0xffffffffu, bv, off);
                const uint32_t oj = __shfl_down_sync(0xffffffffu, bj, off);
                if (hc_key_less(ov, oj, bv, bj)) { bv = ov; bj = oj; }
            }
            if (lane == 0) { s.rmin_v[i] = bv; s.rmin_j[i] = bj; }
        }
    }
}

// closest_elements over the row cache: (value, j, i) minimum of the `order - 1` entries, one partial
// per CTA in the format hclust_merge_kernel reduces.  Also empties the rebuild queue for the merge that
// follows (the rebuild of the previous merge has run by now).
__global__ void __launch_bounds__(256) hclust_nn_argmin_kernel(HcState s) {
    const uint32_t order = s.order[0];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) *s.todo_n = 0;
    HcBest best{FLT_MAX, 0xffffffffu, 0xffffffffu};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < order; i += gridDim.x * blockDim.x)
        hc_consider(s.rmin_v[i], s.rmin_j[i], i, best);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        HcBest o;
        o.v = __shfl_down_sync(0xffffffffu, best.v, off);
        o.j = __shfl_down_sync(0xffffffffu, best.j, off);
        o.i = __shfl_down_sync(0xffffffffu, best.i, off);
        if (hc_better(o.v, o.j, o.i, best)) best = o;
    }
    __shared__ HcBest sm[8];
    if (lane == 0) sm[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        HcBest b = sm[0];
        for (uint32_t w = 1; w < (blockDim.x >> 5); ++w)
            if (hc_better(sm[w].v, sm[w].j, sm[w].i, b)) b = sm[w];
        s.partial[blockIdx.x] = b;
    }
}

// One merge: update_distances(i, j, rule, i) + replace_with_last(j) + bookkeeping
// (clustering_matrix.rs:47-74, hierarchical.rs:44-75), on several CTAs with no barrier between
// them.  Every CTA re-reduces the argmin partials to (i, j); thread k then owns element k of the
// new row: it computes rule(..., D[i][k], D[j][k]), writes D[i][k] and D[k][i], and performs the
// column copy of replace_with_last for row k.  The reference computes all results before it
// writes any; here the only elements another thread could clobber first are k = i and k = j,
// whose inputs are known without reading (the diagonal is 0 and the live matrix is symmetric,
// so D[j][i] = D[i][j] = dij).  The last CTA to finish does the scalar bookkeeping.
__global__ void __launch_bounds__(256) hclust_merge_kernel(HcState s, uint32_t n_partial, uint32_t* done_counter) {
    __shared__ HcBest sm[8];
    __shared__ uint32_t sh_i, sh_j, sh_ticket;
    __shared__ float sh_v;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t order = s.order[0];
    if (order < 2) return;
    HcBest best{FLT_MAX, 0xffffffffu, 0xffffffffu};
    for (uint32_t p = tid; p < n_partial; p += blockDim.x) {
        const HcBest o = s.partial[p];
        if (hc_better(o.v, o.j, o.i, best)) best = o;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        HcBest o;
        o.v = __shfl_down_sync(0xffffffffu, best.v, off);
        o.j = __shfl_down_sync(0xffffffffu, best.j, off);
        o.i = __shfl_down_sync(0xffffffffu, best.i, off);
        if (hc_better(o.v, o.j, o.i, best)) best = o;
    }
    if (lane == 0) sm[warp] = best;
    __syncthreads();
    if (tid == 0) {
        HcBest b = sm[0];
        for (uint32_t w = 1; w < (blockDim.x >> 5); ++w)
            if (hc_better(sm[w].v, sm[w].j, sm[w].i, b)) b = sm[w];
        if (b.j == 0xffffffffu) { b.i = 0; b.j = 0; }   // nothing below f32::MAX: the reference returns (0,0)
        sh_i = b.i; sh_j = b.j; sh_v = b.v;
    }
    __syncthreads();
    const uint32_t i = sh_i, j = sh_j;
    const uint32_t pitch = s.pitch;
    const uint32_t last = order - 1;
    const bool bad = (i == j);   // the reference panics here (clusters.remove(&j))
    const uint32_t step = s.order[1];
    // D[i][j] is the minimum the scan found; it must NOT be re-read from memory here: another CTA's
    // column copy may already have overwritten row i's entry j
    const float dij = sh_v;
    if (!bad) {
        const uint32_t ri = s.rmap[i], rj = s.rmap[j], rl = s.rmap[last];
        float* row_i = s.D + (size_t)ri * pitch;
        const float* row_j = s.D + (size_t)rj * pitch;
        const uint32_t si = s.sizes[i], sj = s.sizes[j], sk = s.sizes[j];   // sizes[j] as size_k: clustering_matrix.rs:66
        const bool moved = j < last;                                        // replace_with_last(j) follows
        for (uint32_t k = blockIdx.x * blockDim.x + tid; k < order; k += gridDim.x * blockDim.x) {
            const float dik = k == i ? 0.0f : (k == j ? dij : row_i[k]);
            const float djk = k == j ? 0.0f : (k == i ? dij : row_j[k]);
            float r = hc_rule(s.rule, si, sj, sk, dij, dik, djk);
            const uint32_t rk = k == i ? ri : (k == j ? rj : (k == last ? rl : s.rmap[k]));
            float* row_k = s.D + (size_t)rk * pitch;
            if (k == i) r = 0.0f;                              // matrix[new_index][new_index] = 0.0 (:72)
            if (!(moved && k == j)) row_i[k] = r;              // (k == j: the column copy below overwrites it)
            row_k[i] = r;
            if (s.rmin_v && k != i && k != j && k != last) {
                // this row keeps its logical index; of its cells right of the diagonal, column i now holds r,
                // column j receives column last (if a move follows) and column last disappears
                float cv = s.rmin_v[k];
                uint32_t cj = s.rmin_j[k];
                float bv = FLT_MAX;
                uint32_t bj = 0xffffffffu;
                if (i > k && r < FLT_MAX) { bv = r; bj = i; }
                if (moved && j > k) {
                    const float w = row_k[last];
                    if (w < FLT_MAX && hc_key_less(w, j, bv, bj)) { bv = w; bj = j; }
                }
                const bool touched = cj == i || cj == j || cj == last;      // the cached cell was one of the three
                const bool have = bj != 0xffffffffu;
                if (!touched) {
                    if (have && hc_key_less(bv, bj, cv, cj)) { s.rmin_v[k] = bv; s.rmin_j[k] = bj; }
                } else if (have && !hc_key_less(cv, cj, bv, bj)) {          // candidate <= cached: nothing else can beat it
                    s.rmin_v[k] = bv; s.rmin_j[k] = bj;
                } else {
                    s.todo[atomicAdd(s.todo_n, 1u)] = k;
                }
            }
            if (moved) {
                // replace_with_last(j): logical row j becomes the old last row; column j <- column last
                if (k == last) {
                    row_i[j] = r;                              // row i: its new [last] entry is this r
                } else if (k == j) {
                    float* nr = s.D + (size_t)rl * pitch;      // the row that moves into slot j
                    nr[j] = nr[last];
                } else if (k != i) {
                    row_k[j] = row_k[last];
                }
            }
        }
    }
    // the last CTA to get here owns the scalars (all others have finished reading them)
    __threadfence();
    __syncthreads();
    if (tid == 0) sh_ticket = atomicAdd(done_counter, 1u);
    __syncthreads();
    if (sh_ticket != gridDim.x - 1 || tid != 0) return;
    *done_counter = 0;
    if (bad) { s.order[0] = 0; s.order[1] |= 0x80000000u; return; }
    if (s.rmin_v) {
        // rows i and (after the move) j changed completely
        s.todo[atomicAdd(s.todo_n, 1u)] = i;
        if (j < last) s.todo[atomicAdd(s.todo_n, 1u)] = j;
    }
    s.sizes[i] = s.sizes[j] + s.sizes[i];                      // :73
    s.mat_i[step] = i;
    s.mat_j[step] = j;
    s.mdist[step] = dij;
    if (j < last) {
        const uint32_t t = s.rmap[j];
        s.rmap[j] = s.rmap[last];
        s.rmap[last] = t;
    }
    s.order[0] = last;
    s.order[1] = step + 1;
}

}  // namespace bsa
