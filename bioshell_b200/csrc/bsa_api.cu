// bsa_api.cu -- C ABI (include/bioshell_align.h) and host-side planning for the
// B200 global-alignment kernels in gotoh_kernels.cuh.
//
// Host work here is what the reference does in `align_all_pairs`
// (bioshell-seq/src/alignment/alignment_protocols.rs:83-115) around the aligner:
// pick the pairs, encode the sequences (similarity_score.rs:125-134), collect the
// results in t-major order.  There is no CPU alignment path in this file.
#include "../../include/bioshell_align.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <map>
#include <vector>

#include "gotoh_kernels.cuh"
#include "hclust_kernels.cuh"
#include "int_peak.cuh"

using namespace bsa;

namespace {

constexpr int kMaxSets = 8;
#ifndef BSA_STREAMS
#define BSA_STREAMS 8      // concurrent kernel groups (same-box A/B: 4 -> 8 streams +1.1 % on cfg2, +8 % on 1,000 sequences; 16 no further gain)
#endif
constexpr int kStreams = BSA_STREAMS;
constexpr int kMaxCodes = 128;
constexpr size_t kSmemBudget = 200 * 1024;   // profile bytes per CTA we are willing to use
constexpr size_t kSmemMax = 227 * 1024;      // dynamic shared memory one CTA can opt in to (sm_100)
constexpr int kKMax = 32;          // direction-store kernels
constexpr int kKStream = 20;       // streaming kernels: beyond 20 columns per lane the row state no
                                   // longer fits 128 registers (2 CTAs/SM); longer templates take passes

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct SeqSet {
    bool loaded = false;
    uint32_t n = 0;
    uint64_t total = 0;
    uint64_t maxlen = 0;
    std::vector<uint64_t> off;       // host copy, n+1
    std::vector<uint32_t> empties;   // indices of zero-length sequences (ascending)
    DevBuf codes;                    // kFrontPad + total + kBackPad bytes
    DevBuf doff;                     // n+1 uint64
    // even-aligned copy (a PAD byte ahead of every odd-length sequence), built on first use as the
    // stream of the two-row kernels (ensure_aligned); invalidated whenever the set is loaded again
    mutable DevBuf acodes, aoff;
    mutable bool aligned = false;
    SeqStoreDev adev() const {
        SeqStoreDev d;
        d.codes = acodes.as<uint8_t>() + kFrontPad;
        d.off = aoff.as<uint64_t>();
        d.n = n;
        return d;
    }
    SeqStoreDev dev() const {
        SeqStoreDev d;
        d.codes = codes.as<uint8_t>() + kFrontPad;
        d.off = doff.as<uint64_t>();
        d.n = n;
        return d;
    }
    uint64_t len(uint32_t i) const { return off[i + 1] - off[i]; }
};

}  // namespace

struct MultiState;   // multi_device.inl: the children and worker threads of a multi-device context

struct bsa_ctx {
    MultiState* multi = nullptr;   // non-null: this context only fans calls out to its children
    int device = 0;
    int sms = 0;
    std::string err;
    cudaStream_t streams[kStreams] = {};
    cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_s[kStreams] = {};
    cudaEvent_t ev_d0 = nullptr, ev_d1 = nullptr;   // run_pairs_dirs: around the kernels of one batch
    double dirs_kernel_ms = 0.0;                     // ... summed over the batches of the last call
    int wave_attr_smem = -1;    // wavefront kernel: occupancy checked for this shared-memory size
    std::mutex* gpu_gate = nullptr;   // child of a multi-device context: one tile's kernels at a time per GPU (planning and copies overlap)

    // residue alphabet: one code per distinct raw byte ever loaded; code 0 (kPadCode) is reserved for
    // the PAD rows of the even-aligned stores and never stands for a residue
    int code_of[256];
    uint8_t byte_of[kMaxCodes];
    int ncodes = 1;

    bool have_scoring = false;
    int32_t score[441];
    uint8_t aaidx[256];
    int go = 0, ge = 0;
    int max_m = 0, min_m = 0;
    int subst_codes = -1;          // alphabet size the device table was built for
    DevBuf d_subst, d_isgap;

    SeqSet sets[kMaxSets];

    DevBuf items, counters, scratch, out_scores, out_nid, fixes, pairs, dirs, path, pstart, status,
        raw, lut, presence, progress, wave_items, hc_matrix, hc_aux, items16, scratch16, items_pair,
        lt_codes, lt_off, lt_idx, lq_codes, lq_off, lq_idx, lutB, lutC, wave_bnd, gidx,
        lt_acodes, lt_aoff, lq_acodes, lq_aoff, items16q;
    bsa_stats stats;
    uint64_t pending_h2d = 0;   // bytes uploaded by bsa_load_sequences since the last alignment call
    uint32_t wave_epoch = 0;    // tag of the last wavefront launch on wave_bnd
};

namespace {

thread_local std::string g_create_err;

int fail(bsa_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}
int fail_cuda(bsa_ctx* c, cudaError_t e, const char* what) {
    cudaGetLastError();
    return fail(c, e == cudaErrorMemoryAllocation ? BSA_ERR_OOM : BSA_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call);      \
    } while (0)

inline int bitlen(uint64_t x) {
    int b = 0;
    while (x) { ++b; x >>= 1; }
    return b;
}

// ---------------- kernel tables ----------------
typedef void (*KernelFn)(const KArgs);
KernelFn g_stream_single[kKMax + 1], g_stream_multi[kKMax + 1], g_dirs[kKMax + 1];
KernelFn g_stream_single_tag[kKMax + 1], g_stream_multi_tag[kKMax + 1];   // TAG cell (4 ALU + 3 IMAD)
typedef void (*Kernel16Fn)(const KArgs16);
typedef void (*KernelLocalFn)(const KArgs, LocalOut*);
typedef void (*KernelPairFn)(const KArgsPair);
KernelPairFn g_pair[kKMax + 1], g_pair_tag[kKMax + 1];
KernelLocalFn g_local[kKMax + 1];
Kernel16Fn g_score16_single[kKMax + 1], g_score16_multi[kKMax + 1];
typedef void (*Kernel16QFn)(const KArgs16Q);
Kernel16QFn g_score16_quad[kKMax + 1];

template <int K>
struct Reg {
    static void run() {
        g_stream_single[K] = gotoh_stream_kernel<K, false>;
        g_stream_multi[K] = gotoh_stream_kernel<K, true>;
        g_score16_single[K] = gotoh_score16_kernel<K, false>;
        g_score16_multi[K] = gotoh_score16_kernel<K, true>;
        g_score16_quad[K] = gotoh_score16_quad_kernel<K>;
        g_pair[K] = gotoh_pair_kernel<K>;
        g_stream_single_tag[K] = gotoh_stream_kernel<K, false, true>;
        g_stream_multi_tag[K] = gotoh_stream_kernel<K, true, true>;
        g_pair_tag[K] = gotoh_pair_kernel<K, true>;
        Reg<K - 1>::run();
    }
};
template <>
struct Reg<0> {
    static void run() {}
};
const int kDirsK[] = {2, 4, 8, 12, 16, 24, 32};
void register_kernels_once() {
    Reg<kKStream>::run();
    g_dirs[2] = gotoh_dirs_kernel<2>;
    g_dirs[4] = gotoh_dirs_kernel<4>;
    g_dirs[8] = gotoh_dirs_kernel<8>;
    g_dirs[12] = gotoh_dirs_kernel<12>;
    g_dirs[16] = gotoh_dirs_kernel<16>;
    g_dirs[24] = gotoh_dirs_kernel<24>;
    g_dirs[32] = gotoh_dirs_kernel<32>;
    g_local[2] = gotoh_local_kernel<2>;
    g_local[4] = gotoh_local_kernel<4>;
    g_local[8] = gotoh_local_kernel<8>;
    g_local[12] = gotoh_local_kernel<12>;
    g_local[16] = gotoh_local_kernel<16>;
    g_local[24] = gotoh_local_kernel<24>;
    g_local[32] = gotoh_local_kernel<32>;
}
void register_kernels() {
    static std::once_flag once;
    std::call_once(once, register_kernels_once);
}
// profile + border vectors + the TMA staging area (gotoh_kernels.cuh, smem layout)
size_t smem_for(int K, int C) { return (size_t)(C + 2) * ((K + 3) / 4) * 32 * sizeof(uint4) + stage_bytes(K, C); }

// largest K whose profile fits the shared-memory budget for an alphabet of C codes
int k_cap(int C) {
    const size_t stage = stage_bytes(kKMax, C);
    size_t vmax = (kSmemBudget > stage ? kSmemBudget - stage : 0) / ((size_t)(C + 2) * 32 * sizeof(uint4));
    int kc = (int)std::min<size_t>(kKMax, vmax * 4);
    return kc;
}

// stream groups are indexed K + multi * (kKMax+1) + tag * 2 (kKMax+1)
constexpr int kGroupStride = kKMax + 1;
inline int group_index(int K, bool multi, bool tag) { return K + (multi ? kGroupStride : 0) + (tag ? 2 * kGroupStride : 0); }
inline int group_k(int g) { return g % kGroupStride; }
inline bool group_multi(int g) { return (g / kGroupStride) & 1; }
inline bool group_tag(int g) { return g >= 2 * kGroupStride; }
inline KernelFn group_kernel(int g) {
    const int K = group_k(g);
    if (group_tag(g)) return group_multi(g) ? g_stream_multi_tag[K] : g_stream_single_tag[K];
    return group_multi(g) ? g_stream_multi[K] : g_stream_single[K];
}

struct KChoice { int K; bool multi; uint32_t npass; };
KChoice choose_k(uint64_t m, int C) {
    const int kc = std::min(k_cap(C), kKStream);
    KChoice r;
    if (m <= (uint64_t)32 * kc) {
        r.K = (int)std::max<uint64_t>(1, (m + 31) / 32);
        r.multi = false;
        r.npass = 1;
    } else {
        r.npass = (uint32_t)((m + 32ull * kc - 1) / (32ull * kc));
        r.K = (int)((m + 32ull * r.npass - 1) / (32ull * r.npass));
        r.multi = true;
    }
    return r;
}
int choose_dirs_k(uint64_t m, int C) {
    const int kc = k_cap(C);
    int best = 0;
    for (int k : kDirsK) {
        if (k > kc) break;
        best = k;
        if ((uint64_t)32 * k >= m) return k;
    }
    return best;   // multi-pass with the largest K that fits
}

// The dynamic shared-memory limit of a kernel is a property of the (device, function), shared by every context of
// the process.  Contexts with different alphabets ask for different sizes, so the limit is only ever RAISED
// (a per-context "already set" flag let a second context lower it under the first one's feet).
cudaError_t raise_smem_limit(int device, const void* fn, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> limit;
    std::lock_guard<std::mutex> lk(mu);
    size_t& lim = limit[std::make_pair(device, fn)];
    if (smem <= lim) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) lim = smem;
    return e;
}

// persistent grid for `n_items` items: one wave of resident CTAs at most
int grid_for(bsa_ctx* ctx, KernelFn fn, int K, int C, uint32_t n_items, uint32_t* grid) {
    const size_t smem = smem_for(K, C);
    // the limit and the occupancy are asked once per (device, kernel, size) instead of once per launch
    CK(raise_smem_limit(ctx->device, (const void*)fn, smem));
    int nb = 0;
    {
        static std::mutex mu;
        static std::map<std::tuple<int, const void*, size_t>, int> occ;                  // (device, fn, smem) -> CTAs per SM
        std::lock_guard<std::mutex> lk(mu);
        auto key = std::make_tuple(ctx->device, (const void*)fn, smem);
        auto it = occ.find(key);
        if (it != occ.end()) {
            nb = it->second;
        } else {
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kThreads, smem));
            occ[key] = nb;
        }
    }
    if (nb < 1) return fail(ctx, BSA_ERR_CUDA, "kernel does not fit on an SM");
    *grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)nb * ctx->sms);
    return BSA_OK;
}

int launch(bsa_ctx* ctx, KernelFn fn, int K, const KArgs& a, cudaStream_t st) {
    const size_t smem = smem_for(K, a.C);
    uint32_t grid = 0;
    int rc = grid_for(ctx, fn, K, a.C, a.n_items, &grid);
    if (rc) return rc;
    if (grid == 0) return BSA_OK;
    fn<<<grid, kThreads, smem, st>>>(a);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    return BSA_OK;
}

// (re)build the C x C substitution table by residue code on the device
int sync_scoring(bsa_ctx* ctx) {
    if (!ctx->have_scoring) return fail(ctx, BSA_ERR_BAD_ARG, "bsa_set_scoring has not been called");
    if (ctx->subst_codes == ctx->ncodes) return BSA_OK;
    const int C = std::max(ctx->ncodes, 1);
    std::vector<int16_t> tab((size_t)C * C);
    std::vector<uint8_t> gap(C);
    for (int a = 0; a < C; ++a) {
        gap[a] = (ctx->byte_of[a] == '-' || ctx->byte_of[a] == '_') ? 1 : 0;   // msa.rs:264
        for (int b = 0; b < C; ++b)
            // similarity_score.rs:139-141 -> substitution_matrix.rs:81-83
            tab[(size_t)a * C + b] =
                (int16_t)ctx->score[ctx->aaidx[ctx->byte_of[a]] * 21 + ctx->aaidx[ctx->byte_of[b]]];
    }
    CK(ctx->d_subst.ensure(tab.size() * sizeof(int16_t)));
    CK(ctx->d_isgap.ensure(gap.size()));
    CK(cudaMemcpy(ctx->d_subst.p, tab.data(), tab.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_isgap.p, gap.data(), gap.size(), cudaMemcpyHostToDevice));
    ctx->subst_codes = ctx->ncodes;
    return BSA_OK;
}

// Even-aligned copy of a packed store for the two-row streaming kernels (gotoh_kernels.cuh,
// stream_block_tag2a): every sequence starts at an even position and takes an even number of bytes,
// an odd-length one behind a PAD byte (kPadCode); pads and the flagged byte ahead of sequence 0 as in
// the store itself.  `off` = the host copy of the store's offsets.
int build_aligned(bsa_ctx* ctx, const SeqStoreDev& src, const std::vector<uint64_t>& off, DevBuf& acodes,
                  DevBuf& aoff, cudaStream_t st) {
    const size_t n = off.size() - 1;
    std::vector<uint64_t> ao(n + 1, 0);
    for (size_t i = 0; i < n; ++i) ao[i + 1] = ao[i] + ((off[i + 1] - off[i] + 1) & ~(uint64_t)1);
    CK(acodes.ensure(kFrontPad + ao[n] + kBackPad));
    CK(aoff.ensure((n + 1) * 8));
    CK(cudaMemcpyAsync(aoff.p, ao.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(acodes.p, 0, kFrontPad, st));
    CK(cudaMemsetAsync(acodes.as<uint8_t>() + kFrontPad - 1, (int)kLastFlag, 1, st));
    CK(cudaMemsetAsync(acodes.as<uint8_t>() + kFrontPad + ao[n], 0, kBackPad, st));
    if (n) {
        align_seqs_kernel<<<(uint32_t)n, 128, 0, st>>>(src, aoff.as<uint64_t>(), acodes.as<uint8_t>() + kFrontPad);
        CK(cudaGetLastError());
        ctx->stats.launches++;
    }
    ctx->stats.h2d_bytes += (n + 1) * 8;
    CK(cudaStreamSynchronize(st));      // `ao` is a temporary
    return BSA_OK;
}
int ensure_aligned(bsa_ctx* ctx, const SeqSet& S) {
    if (S.aligned) return BSA_OK;
    int rc = build_aligned(ctx, S.dev(), S.off, S.acodes, S.aoff, ctx->streams[0]);
    if (rc) return rc;
    S.aligned = true;
    return BSA_OK;
}

struct PairReq { uint32_t q, t; uint64_t out; };

// Fill + traceback through the direction-store kernels for an explicit pair list.
// d_scores / d_nid are device arrays indexed by PairReq::out (either may be null).
// If path_buf != null, pair i's path (right-aligned in its len_q+len_t slot at
// slot_off[i]) and its length are returned on the host.
int run_pairs_dirs(bsa_ctx* ctx, const SeqSet& Q, const SeqSet& T, std::vector<PairReq>& reqs,
                   int32_t* d_scores, uint32_t* d_nid, uint8_t* path_buf,
                   const std::vector<uint64_t>* slot_off, std::vector<uint32_t>* path_len,
                   const std::vector<uint64_t>* req_index, LocalOut* d_lout = nullptr) {
    const bool local = d_lout != nullptr;   // LocalAlignment instead of GlobalAligner
    ctx->dirs_kernel_ms = 0.0;
    if (reqs.empty()) return BSA_OK;
    const int C = std::max(ctx->ncodes, 1);
    // order by template so a CTA shares one profile between its warps
    std::vector<uint32_t> order(reqs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return reqs[a].t < reqs[b].t; });
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t dir_budget = std::max<uint64_t>((uint64_t)(free_b * 0.6), 64ull << 20);

    size_t pos = 0;
    while (pos < order.size()) {
        // ---- take a batch that fits the direction budget ----
        std::vector<PairRec> recs;
        std::vector<uint32_t> rec_req;
        std::vector<uint32_t> wave_recs;    // K3 pairs of this batch
        uint64_t dir_words = 0, scr_entries = 0, bnd_entries = 0, path_bytes = 0;
        // K3: templates this long run their column blocks as a wavefront over many warps
        const size_t wave_fixed = subst_stage_bytes(C) + 2048;     // table staging + alignment slack + static
        const int wave_warps = (int)std::min<size_t>(kWaveWarps, (kSmemMax > wave_fixed ? kSmemMax - wave_fixed : 0) / wave_region_bytes(C));
        while (pos < order.size()) {
            const PairReq& r = reqs[order[pos]];
            const uint64_t n = Q.len(r.q), m = T.len(r.t);
            bool wave = !local && m >= 4096 && wave_warps >= 1;
            if (wave) {
                // K3 holds (score - (i + j) ge) * 8 + 3 flag bits in an int32: the frame adds up to (n + m) |ge|
                const uint64_t m_pad = (m + 255) / 256 * 256;
                const int64_t ub = (int64_t)std::max(ctx->max_m, 0) * (int64_t)std::min(n, m) + (int64_t)(n + m_pad + 8) * (-(int64_t)ctx->ge);
                const int64_t lb = 4 * (int64_t)(-(int64_t)ctx->go) + (int64_t)(n + m_pad + 8) * (-(int64_t)ctx->ge) +
                                   (int64_t)std::max(-ctx->min_m, 0);
                if (std::max(ub, lb) + 8 >= ((int64_t)1 << 27)) wave = false;      // the sequential passes below have 29 bits
            }
            const int K = wave ? kWaveK : choose_dirs_k(m, C);
            if (K < 1) return fail(ctx, BSA_ERR_ALPHABET, "alphabet too large for shared memory");
            const uint64_t W = (K + 7) / 8, npass = (m + 32ull * K - 1) / (32ull * K);
            // the per-template kernels hold score * 4 + priority in an int32 (cs = 0): every DP value of the pair,
            // the borders over the padded columns and one opening below them must stay inside 29 bits
            if (!wave) {
                const uint64_t m_pad = npass * 32ull * K;
                const int64_t ub = (int64_t)std::max(ctx->max_m, 0) * (int64_t)std::min(n, m);
                const int64_t lb = 4 * (int64_t)(-(int64_t)ctx->go) + (int64_t)(n + m_pad + 4) * (-(int64_t)ctx->ge) +
                                   (int64_t)std::max(-ctx->min_m, 0);
                if (std::max(ub, lb) + 8 >= ((int64_t)1 << 29))
                    return fail(ctx, BSA_ERR_RANGE, "pair too long for these gap penalties / scores: score * 4 leaves int32");
            }
            const uint64_t rpl = wave ? kWaveR : 1;   // rows per plane line
            const uint64_t words = npass * ((n + rpl - 1) / rpl + 32) * 32 * rpl * W;
            if (!recs.empty() && (dir_words + words) * 4 > dir_budget) break;
            PairRec pr;
            pr.q = r.q; pr.t = r.t; pr.out = r.out;
            pr.dir_off = dir_words;
            pr.scr_off = wave ? bnd_entries : scr_entries;
            pr.path_off = path_bytes;
            pr.k = (uint32_t)K | (rpl > 1 ? (uint32_t)rpl << kPlaneRowsShift : 0u);
            pr.prog_off = wave ? 0u : 0xffffffffu;      // marks the K3 pairs
            dir_words += words;
            if (wave) {
                bnd_entries += (npass - 1) * wave_bnd_stride(n);
                wave_recs.push_back((uint32_t)recs.size());
            } else {
                scr_entries += npass > 1 ? n : 0;
            }
            path_bytes += n + m;
            recs.push_back(pr);
            rec_req.push_back(order[pos]);
            ++pos;
        }
        // K3 items: (pair, column block), the pairs in order of their critical paths (rows / 4 + ~40 steps of
        // hand-off lag per block), the blocks of a pair left to right: a block's left neighbour always comes first
        std::vector<uint2> wave_items;
        {
            auto crit = [&](uint32_t ri) {
                const uint64_t n = Q.len(recs[ri].q), m = T.len(recs[ri].t);
                return n / 4 + 40 * ((m + 255) / 256);
            };
            std::stable_sort(wave_recs.begin(), wave_recs.end(), [&](uint32_t x, uint32_t y) { return crit(x) > crit(y); });
            for (uint32_t ri : wave_recs) {
                const uint64_t npass = (T.len(recs[ri].t) + 255) / 256;
                for (uint64_t ps = 0; ps < npass; ++ps) wave_items.push_back(make_uint2(ri, (uint32_t)ps));
            }
        }
        // ---- items: (template, K) runs of the batch ----
        std::vector<Item> items;
        std::vector<int> item_k;
        for (size_t i = 0; i < recs.size();) {
            size_t j = i;
            while (j < recs.size() && recs[j].t == recs[i].t) ++j;
            if (recs[i].prog_off != 0xffffffffu) { i = j; continue; }   // wavefront pairs: see below
            // keep CTAs busy: at most 4 pairs per warp per item
            for (size_t b = i; b < j; b += 4 * kWarpsPerCta) {
                Item it;
                it.t = recs[i].t;
                it.q_begin = (uint32_t)b;
                it.q_end = (uint32_t)std::min(j, b + 4 * kWarpsPerCta);
                it.cshift = 0;
                it.out_base = 0;
                items.push_back(it);
                item_k.push_back((int)(recs[i].k & kPlaneKMask));
            }
            i = j;
        }
        std::vector<size_t> perm(items.size());
        for (size_t i = 0; i < perm.size(); ++i) perm[i] = i;
        std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return item_k[a] > item_k[b]; });
        std::vector<Item> sorted(items.size());
        for (size_t i = 0; i < perm.size(); ++i) sorted[i] = items[perm[i]];

        CK(ctx->pairs.ensure(recs.size() * sizeof(PairRec)));
        CK(ctx->items.ensure(sorted.size() * sizeof(Item)));
        CK(ctx->dirs.ensure(std::max<uint64_t>(dir_words, 1) * 4));
        CK(ctx->scratch.ensure(std::max<uint64_t>(scr_entries, 1) * sizeof(uint2)));
        CK(ctx->pstart.ensure(recs.size() * 4));
        CK(ctx->status.ensure(4));
        CK(ctx->counters.ensure(64 * 4));
        if (path_buf) CK(ctx->path.ensure(std::max<uint64_t>(path_bytes, 1)));
        cudaStream_t st = ctx->streams[0];
        CK(cudaMemcpyAsync(ctx->pairs.p, recs.data(), recs.size() * sizeof(PairRec), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->items.p, sorted.data(), sorted.size() * sizeof(Item), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(ctx->counters.p, 0, 64 * 4, st));
        CK(cudaMemsetAsync(ctx->status.p, 0, 4, st));
        ctx->stats.h2d_bytes += recs.size() * sizeof(PairRec) + sorted.size() * sizeof(Item);

        // everything the host has to ask the runtime is asked before the first launch, so that the kernels of the
        // batch run back to back and the two events around them time the device, not the host
        const size_t wave_smem = wave_smem_bytes(wave_warps, C) + 1024;      // + slack to align the rings to 1 KB
        if (!wave_items.empty()) CK(raise_smem_limit(ctx->device, (const void*)gotoh_wave_kernel, wave_smem));
        if (!wave_items.empty() && ctx->wave_attr_smem != (int)wave_smem) {
            int nb = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gotoh_wave_kernel, wave_warps * 32, wave_smem));
            if (nb < 1) return fail(ctx, BSA_ERR_CUDA, "wavefront kernel does not fit on an SM");
            ctx->wave_attr_smem = (int)wave_smem;
        }
        if (!local) CK(raise_smem_limit(ctx->device, (const void*)traceback_kernel, 2 * kTraceWinLong));
        if (!wave_items.empty()) {
            CK(ctx->wave_items.ensure(wave_items.size() * sizeof(uint2)));
            CK(cudaMemcpyAsync(ctx->wave_items.p, wave_items.data(), wave_items.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
            const size_t need = std::max<uint64_t>(bnd_entries, 1) * sizeof(uint4);
            if (need > ctx->wave_bnd.cap || ctx->wave_epoch == 0xfffffffeu) {
                // the boundary buffer only ever holds {H, epoch, E, epoch} entries: zeroed when (re)allocated, epochs never repeat on it
                CK(ctx->wave_bnd.ensure(need));
                CK(cudaMemsetAsync(ctx->wave_bnd.p, 0, ctx->wave_bnd.cap, st));
                ctx->wave_epoch = 0;
            }
        }
        CK(cudaEventRecord(ctx->ev_d0, st));
        size_t gi = 0;
        int group = 0;
        while (gi < sorted.size()) {
            const int K = item_k[perm[gi]];
            size_t gj = gi;
            while (gj < sorted.size() && item_k[perm[gj]] == K) ++gj;
            KArgs a;
            memset(&a, 0, sizeof(a));
            a.Q = Q.dev(); a.T = T.dev();
            a.subst = ctx->d_subst.as<int16_t>(); a.isgap = ctx->d_isgap.as<uint8_t>();
            a.C = C; a.go = ctx->go; a.ge = ctx->ge; a.one = 1;
            a.items = ctx->items.as<Item>() + gi;
            a.n_items = (uint32_t)(gj - gi);
            a.item_counter = ctx->counters.as<uint32_t>() + group;
            a.scores = d_scores; a.nident = nullptr;
            a.scratch = ctx->scratch.as<uint2>(); a.scratch_stride = 0;
            a.pairs = ctx->pairs.as<PairRec>();
            a.dirs = ctx->dirs.as<uint32_t>();
            int rc = BSA_OK;
            if (local) {
                const size_t smem = smem_for(K, a.C);
                uint32_t grid = 0;
                rc = grid_for(ctx, (KernelFn)g_local[K], K, a.C, a.n_items, &grid);
                if (rc) return rc;
                g_local[K]<<<grid, kThreads, smem, st>>>(a, d_lout);
                CK(cudaGetLastError());
                ctx->stats.launches++;
            } else {
                rc = launch(ctx, g_dirs[K], K, a, st);
            }
            if (rc) return rc;
            gi = gj;
            ++group;
        }
        if (!wave_items.empty()) {
            KArgs a;
            memset(&a, 0, sizeof(a));
            a.Q = Q.dev(); a.T = T.dev();
            a.subst = ctx->d_subst.as<int16_t>(); a.isgap = ctx->d_isgap.as<uint8_t>();
            a.C = C; a.go = ctx->go; a.ge = ctx->ge; a.one = 1; a.one2 = 1; a.mone = -1;
            a.n_items = (uint32_t)wave_items.size();
            a.item_counter = ctx->counters.as<uint32_t>() + 63;
            a.scores = d_scores;
            a.pairs = ctx->pairs.as<PairRec>();
            a.dirs = ctx->dirs.as<uint32_t>();
            a.wave_items = ctx->wave_items.as<uint2>();
            a.wave_bnd = ctx->wave_bnd.as<uint4>();
            a.epoch = ++ctx->wave_epoch;
            const size_t smem = wave_smem;
            const int nb = 1;     // the wavefront is bound by its critical path: two warps per scheduler at most
            // one CTA per SM, as many as there are items: the kernel deals the items out round-robin over the CTAs
            uint32_t grid = (uint32_t)std::min<uint64_t>(wave_items.size(), (uint64_t)nb * ctx->sms);
            if (const char* e = getenv("BSA_WAVE_GRID")) grid = std::max(1u, std::min(grid, (uint32_t)atoi(e)));   // diagnostics
            const char* trace_path = getenv("BSA_WAVE_TRACE");     // diagnostics: per-item start / end times
            unsigned long long* d_trace = nullptr;
            if (trace_path) {
                CK(cudaMalloc(&d_trace, wave_items.size() * 32));
                CK(cudaMemsetAsync(d_trace, 0, wave_items.size() * 32, st));
                a.wave_trace = d_trace;
            }
            gotoh_wave_kernel<<<grid, wave_warps * 32, smem, st>>>(a);
            CK(cudaGetLastError());
            ctx->stats.launches++;
            if (trace_path) {
                std::vector<unsigned long long> tr(wave_items.size() * 4);
                CK(cudaMemcpyAsync(tr.data(), d_trace, tr.size() * 8, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                cudaFree(d_trace);
                if (FILE* f = fopen(trace_path, "w")) {
                    fprintf(f, "item,pair,block,n,m,start_ns,end_ns,sm,warp,steps\n");
                    for (size_t i = 0; i < wave_items.size(); ++i)
                        fprintf(f, "%zu,%u,%u,%llu,%llu,%llu,%llu,%llu,%llu,%llu\n", i, wave_items[i].x, wave_items[i].y,
                                (unsigned long long)Q.len(recs[wave_items[i].x].q), (unsigned long long)T.len(recs[wave_items[i].x].t),
                                tr[4 * i], tr[4 * i + 1], tr[4 * i + 2] >> 8, tr[4 * i + 2] & 255, tr[4 * i + 3]);
                    fclose(f);
                }
            }
        }
        TraceArgs ta;
        ta.Q = Q.dev(); ta.T = T.dev();
        ta.pairs = ctx->pairs.as<PairRec>();
        ta.n_pairs = (uint32_t)recs.size();
        ta.dirs = ctx->dirs.as<uint32_t>();
        ta.isgap = ctx->d_isgap.as<uint8_t>();
        ta.ncodes = (uint32_t)std::max(ctx->ncodes, 1);
        ta.path = path_buf ? ctx->path.as<uint8_t>() : nullptr;
        ta.path_start = ctx->pstart.as<uint32_t>();
        ta.nident = d_nid;
        ta.status = ctx->status.as<uint32_t>();
        if (local) {
            const uint32_t tb_blocks = (ta.n_pairs + kLocalTraceThreads / 32 - 1) / (kLocalTraceThreads / 32);   // one warp per pair
            traceback_local_kernel<<<tb_blocks, kLocalTraceThreads, 0, st>>>(ta, d_lout);
        } else {
            // one warp per pair; K3 pairs walk tens of thousands of steps each: one warp per CTA and large windows
            const bool longp = !wave_recs.empty();
            const uint32_t tb_warps = longp ? 1u : (uint32_t)kTraceMaxWarps;
            ta.win_bytes = longp ? kTraceWinLong : kTraceWinShort;
            const size_t tb_smem = (size_t)tb_warps * 2 * ta.win_bytes;
            const uint32_t tb_blocks = (ta.n_pairs + tb_warps - 1) / tb_warps;
            traceback_kernel<<<tb_blocks, tb_warps * 32, tb_smem, st>>>(ta);
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev_d1, st));
        ctx->stats.launches++;
        uint32_t status = 0;
        CK(cudaMemcpyAsync(&status, ctx->status.p, 4, cudaMemcpyDeviceToHost, st));
        std::vector<uint32_t> starts;
        std::vector<uint8_t> hpath;
        if (path_buf) {
            starts.resize(recs.size());
            hpath.resize(path_bytes);
            CK(cudaMemcpyAsync(starts.data(), ctx->pstart.p, recs.size() * 4, cudaMemcpyDeviceToHost, st));
            if (path_bytes)
                CK(cudaMemcpyAsync(hpath.data(), ctx->path.p, path_bytes, cudaMemcpyDeviceToHost, st));
            ctx->stats.d2h_bytes += recs.size() * 4 + path_bytes;
        }
        CK(cudaStreamSynchronize(st));
        {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ctx->ev_d0, ctx->ev_d1));
            ctx->dirs_kernel_ms += ms;
        }
        if (status) return fail(ctx, BSA_ERR_CUDA, "traceback met an invalid direction code");
        if (path_buf) {
            for (size_t i = 0; i < recs.size(); ++i) {
                const uint64_t ri = (*req_index)[rec_req[i]];
                const uint64_t slot = Q.len(recs[i].q) + T.len(recs[i].t);
                const uint32_t s0 = starts[i];
                memcpy(path_buf + (*slot_off)[ri] + s0, hpath.data() + recs[i].path_off + s0, slot - s0);
                (*path_len)[ri] = (uint32_t)(slot - s0);
            }
        }
    }
    return BSA_OK;
}

}  // namespace

#include "multi_device.inl"

// =====================================================================
//                               C ABI
// =====================================================================
extern "C" {

int bsa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

bsa_ctx* bsa_create(int device_id) {
    int n = bsa_device_count();
    if (n <= 0) { fail(nullptr, BSA_ERR_CUDA, "no CUDA device is visible (there is no CPU fallback)"); return nullptr; }
    if (device_id < 0 || device_id >= n) { fail(nullptr, BSA_ERR_BAD_ARG, "device_id out of range"); return nullptr; }
    if (cudaSetDevice(device_id) != cudaSuccess) { fail(nullptr, BSA_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { fail(nullptr, BSA_ERR_CUDA, "cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major != 10) {
        fail(nullptr, BSA_ERR_CUDA, "device is not sm_100 (this library carries sm_100a code only)");
        return nullptr;
    }
    bsa_ctx* c = new (std::nothrow) bsa_ctx();
    if (!c) return nullptr;
    c->device = device_id;
    c->sms = prop.multiProcessorCount;
    for (int i = 0; i < 256; ++i) c->code_of[i] = -1;
    c->byte_of[kPadCode] = 0xff;   // byte 255 is never a residue (bsa_load_sequences refuses it)
    memset(&c->stats, 0, sizeof(c->stats));
    for (int i = 0; i < kStreams; ++i) {
        cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&c->ev_s[i], cudaEventDisableTiming);
    }
    cudaEventCreate(&c->ev_start);
    cudaEventCreate(&c->ev_end);
    cudaEventCreate(&c->ev_d0);
    cudaEventCreate(&c->ev_d1);
    register_kernels();
    if (cudaGetLastError() != cudaSuccess) { delete c; fail(nullptr, BSA_ERR_CUDA, "stream/event creation failed"); return nullptr; }
    return c;
}

void bsa_destroy(bsa_ctx* c) {
    if (!c) return;
    if (c->multi) { multi_destroy(c); return; }
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto& s : c->sets) { s.codes.release(); s.doff.release(); s.acodes.release(); s.aoff.release(); }
    DevBuf* bufs[] = {&c->items, &c->counters, &c->scratch, &c->out_scores, &c->out_nid, &c->fixes,
                      &c->pairs, &c->dirs, &c->path, &c->pstart, &c->status, &c->raw, &c->lut,
                      &c->presence, &c->d_subst, &c->d_isgap, &c->progress, &c->wave_items, &c->hc_matrix, &c->hc_aux, &c->items16, &c->scratch16, &c->items_pair,
                      &c->lt_codes, &c->lt_off, &c->lt_idx, &c->lq_codes, &c->lq_off, &c->lq_idx, &c->lutB, &c->lutC,
                      &c->wave_bnd, &c->gidx, &c->lt_acodes, &c->lt_aoff, &c->lq_acodes, &c->lq_aoff, &c->items16q};
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < kStreams; ++i) {
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
        if (c->ev_s[i]) cudaEventDestroy(c->ev_s[i]);
    }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    if (c->ev_d0) cudaEventDestroy(c->ev_d0);
    if (c->ev_d1) cudaEventDestroy(c->ev_d1);
    delete c;
}

const char* bsa_last_error(const bsa_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

// substitution_matrix.rs:96-135
int bsa_parse_ncbi_matrix(const char* text, size_t len, int32_t score[441], uint8_t aa_index[256]) {
    if (!text || !score || !aa_index) return BSA_ERR_BAD_ARG;
    std::fill(score, score + 441, 0);
    std::fill(aa_index, aa_index + 256, (uint8_t)0);
    size_t row = 0, at = 0;
    while (at < len && row < 20) {
        size_t eol = at;
        while (eol < len && text[eol] != '\n') ++eol;
        std::string line(text + at, eol - at);
        at = eol + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && (line[0] == '#' || line[0] == ' ')) continue;   // :105
        std::vector<std::string> tok;
        size_t k = 0;
        while (k < line.size()) {
            while (k < line.size() && isspace((unsigned char)line[k])) ++k;
            size_t s = k;
            while (k < line.size() && !isspace((unsigned char)line[k])) ++k;
            if (k > s) tok.emplace_back(line, s, k - s);
        }
        if (tok.size() < 23) return BSA_ERR_FORMAT;                           // :108
        auto to_i32 = [](const std::string& s, int32_t* out) {
            size_t i = 0;
            bool neg = false;
            if (!s.empty() && (s[0] == '+' || s[0] == '-')) { neg = s[0] == '-'; i = 1; }
            if (i >= s.size()) return false;
            long long v = 0;
            for (; i < s.size(); ++i) {
                if (s[i] < '0' || s[i] > '9') return false;
                v = v * 10 + (s[i] - '0');
                if (v > 2147483648LL) return false;
            }
            v = neg ? -v : v;
            if (v > 2147483647LL) return false;
            *out = (int32_t)v;
            return true;
        };
        const unsigned char letter = (unsigned char)tok[0][0];
        if (letter == 255) return BSA_ERR_FORMAT;
        aa_index[letter] = (uint8_t)row;                                      // :110
        for (size_t j = 1; j <= 20; ++j) {                                    // :112-119
            int32_t v;
            if (!to_i32(tok[j], &v)) return BSA_ERR_FORMAT;
            score[row * 21 + (j - 1)] = v;
            score[(j - 1) * 21 + row] = v;
        }
        int32_t vx;                                                           // :121-126
        if (!to_i32(tok[tok.size() - 2], &vx)) return BSA_ERR_FORMAT;
        score[row * 21 + 20] = vx;
        score[20 * 21 + row] = vx;
        ++row;
    }
    aa_index[(unsigned char)'X'] = 20;                                        // :130
    score[20 * 21 + 20] = -1;                                                 // :132
    return BSA_OK;
}

int bsa_set_scoring(bsa_ctx* ctx, const int32_t score[441], const uint8_t aa_index[256],
                    int32_t gap_open, int32_t gap_extend) {
    if (!ctx || !score || !aa_index) return fail(ctx, BSA_ERR_BAD_ARG, "null argument");
    if (ctx->multi) return multi_set_scoring(ctx, score, aa_index, gap_open, gap_extend);
    // outside this domain the reference's sentinel can tie or its backtrace underflows
    // (SURVEY.md 8a note 1); the kernels drop the sentinel, so refuse.
    if (!(gap_open <= gap_extend && gap_extend <= 0 && gap_open < 0))
        return fail(ctx, BSA_ERR_UNSUPPORTED_GAPS, "need gap_open <= gap_extend <= 0 and gap_open < 0");
    // every kernel multiplies the gap penalties by the lane's score unit (>= 4); larger values could
    // never pass the per-template range checks anyway
    if (gap_open < -(1 << 20)) return fail(ctx, BSA_ERR_RANGE, "gap_open below -2^20");
    int mx = -1000000, mn = 1000000;
    for (int i = 0; i < 441; ++i) {
        if (score[i] > 32000 || score[i] < -32000) return fail(ctx, BSA_ERR_RANGE, "substitution score out of int16 range");
        mx = std::max(mx, score[i]);
        mn = std::min(mn, score[i]);
    }
    for (int i = 0; i < 256; ++i)
        if (aa_index[i] > 20) return fail(ctx, BSA_ERR_BAD_ARG, "aa_index entry > 20");
    memcpy(ctx->score, score, sizeof(ctx->score));
    memcpy(ctx->aaidx, aa_index, 256);
    ctx->go = gap_open;
    ctx->ge = gap_extend;
    ctx->max_m = mx;
    ctx->min_m = mn;
    ctx->have_scoring = true;
    ctx->subst_codes = -1;
    return BSA_OK;
}

int bsa_load_sequences(bsa_ctx* ctx, int set_id, const uint8_t* residues_raw, const uint64_t* offsets,
                       uint32_t n) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (ctx->multi) return multi_load_sequences(ctx, set_id, residues_raw, offsets, n);
    if (set_id < 0 || set_id >= kMaxSets || !offsets) return fail(ctx, BSA_ERR_BAD_ARG, "bad set_id or null offsets");
    if (n == 0) return fail(ctx, BSA_ERR_EMPTY, "empty sequence set");
    for (uint32_t i = 0; i < n; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, BSA_ERR_BAD_ARG, "offsets must be non-decreasing");
    const uint64_t base = offsets[0], total = offsets[n] - base;
    if (total && !residues_raw) return fail(ctx, BSA_ERR_BAD_ARG, "null residues");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->streams[0];
    SeqSet& S = ctx->sets[set_id];
    S.loaded = false;
    S.aligned = false;
    S.n = n;
    S.total = total;
    S.off.resize((size_t)n + 1);
    S.maxlen = 0;
    S.empties.clear();
    for (uint32_t i = 0; i <= n; ++i) S.off[i] = offsets[i] - base;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t l = S.off[i + 1] - S.off[i];
        S.maxlen = std::max(S.maxlen, l);
        if (l == 0) S.empties.push_back(i);
    }
    const size_t buf_bytes = kFrontPad + total + kBackPad;
    CK(S.codes.ensure(buf_bytes));
    CK(S.doff.ensure(((size_t)n + 1) * 8));
    CK(ctx->raw.ensure(total + 16));
    CK(ctx->lut.ensure(256));
    CK(ctx->presence.ensure(32));
    CK(cudaMemcpyAsync(S.doff.p, S.off.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    if (total) CK(cudaMemcpyAsync(ctx->raw.p, residues_raw + base, total, cudaMemcpyHostToDevice, st));
    ctx->pending_h2d += ((size_t)n + 1) * 8 + total;
    // which byte values occur?  (new ones get a residue code)
    CK(cudaMemsetAsync(ctx->presence.p, 0, 32, st));
    if (total) {
        const int blocks = (int)std::min<uint64_t>((total + 256 * 64 - 1) / (256 * 64), (uint64_t)ctx->sms * 8);
        byte_presence_kernel<<<blocks, 256, 0, st>>>(ctx->raw.as<uint8_t>(), total, ctx->presence.as<uint32_t>());
        CK(cudaGetLastError());
    }
    uint32_t pres[8];
    CK(cudaMemcpyAsync(pres, ctx->presence.p, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (pres[7] >> 31) return fail(ctx, BSA_ERR_ALPHABET, "residue byte 255 (the reference's [u8;255] lookup panics)");
    for (int b = 0; b < 255; ++b) {
        if (!((pres[b >> 5] >> (b & 31)) & 1u) || ctx->code_of[b] >= 0) continue;
        if (ctx->ncodes >= kMaxCodes) return fail(ctx, BSA_ERR_ALPHABET, "more than 127 distinct residue byte values");
        ctx->code_of[b] = ctx->ncodes;
        ctx->byte_of[ctx->ncodes++] = (uint8_t)b;
    }
    uint8_t lut[256];
    for (int b = 0; b < 256; ++b) lut[b] = (uint8_t)(ctx->code_of[b] < 0 ? 0 : ctx->code_of[b]);
    CK(cudaMemcpyAsync(ctx->lut.p, lut, 256, cudaMemcpyHostToDevice, st));
    // pads: zeros, with the byte just before sequence 0 flagged so every lane starts clean
    CK(cudaMemsetAsync(S.codes.p, 0, kFrontPad, st));
    CK(cudaMemsetAsync(S.codes.as<uint8_t>() + kFrontPad - 1, (int)kLastFlag, 1, st));
    CK(cudaMemsetAsync(S.codes.as<uint8_t>() + kFrontPad + total, 0, kBackPad, st));
    if (total) {
        const int blocks = (int)std::min<uint64_t>((total / 16 + 255) / 256 + 1, (uint64_t)ctx->sms * 16);
        encode_kernel<<<blocks, 256, 0, st>>>(ctx->raw.as<uint8_t>(), S.codes.as<uint8_t>() + kFrontPad, total,
                                             ctx->lut.as<uint8_t>());
        CK(cudaGetLastError());
        mark_last_kernel<<<(n + 255) / 256, 256, 0, st>>>(S.codes.as<uint8_t>() + kFrontPad, S.doff.as<uint64_t>(), n);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(st));
    S.loaded = true;
    return BSA_OK;
}

int bsa_gather_sequences(bsa_ctx* ctx, int src_set, int dst_set, const uint32_t* idx, uint32_t n) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (ctx->multi) return multi_gather_sequences(ctx, src_set, dst_set, idx, n);
    if (src_set < 0 || src_set >= kMaxSets || dst_set < 0 || dst_set >= kMaxSets || src_set == dst_set || !idx)
        return fail(ctx, BSA_ERR_BAD_ARG, "bad set id or null index list");
    if (n == 0) return fail(ctx, BSA_ERR_EMPTY, "empty sequence set");
    const SeqSet& S = ctx->sets[src_set];
    if (!S.loaded) return fail(ctx, BSA_ERR_EMPTY, "sequence set not loaded");
    for (uint32_t i = 0; i < n; ++i)
        if (idx[i] >= S.n) return fail(ctx, BSA_ERR_BAD_ARG, "sequence index out of range");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->streams[0];
    SeqSet& D = ctx->sets[dst_set];
    D.loaded = false;
    D.aligned = false;
    D.n = n;
    D.off.assign((size_t)n + 1, 0);
    D.maxlen = 0;
    D.empties.clear();
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t l = S.len(idx[i]);
        D.off[i + 1] = D.off[i] + l;
        D.maxlen = std::max(D.maxlen, l);
        if (l == 0) D.empties.push_back(i);
    }
    D.total = D.off[n];
    CK(D.codes.ensure(kFrontPad + D.total + kBackPad));
    CK(D.doff.ensure(((size_t)n + 1) * 8));
    CK(ctx->gidx.ensure((size_t)n * 4));
    CK(cudaMemcpyAsync(D.doff.p, D.off.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->gidx.p, idx, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    ctx->pending_h2d += ((size_t)n + 1) * 8 + (size_t)n * 4;
    // pads as in bsa_load_sequences; the residues never leave the device and keep their codes and last-residue flags
    CK(cudaMemsetAsync(D.codes.p, 0, kFrontPad, st));
    CK(cudaMemsetAsync(D.codes.as<uint8_t>() + kFrontPad - 1, (int)kLastFlag, 1, st));
    CK(cudaMemsetAsync(D.codes.as<uint8_t>() + kFrontPad + D.total, 0, kBackPad, st));
    gather_seqs_kernel<<<n, 128, 0, st>>>(S.dev(), ctx->gidx.as<uint32_t>(), D.doff.as<uint64_t>(),
                                         D.codes.as<uint8_t>() + kFrontPad);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));      // idx is the caller's
    D.loaded = true;
    return BSA_OK;
}

int bsa_plan_shards(bsa_ctx* ctx, int q_set, int t_set, const uint32_t* q_counts, uint32_t n_shards,
                    uint32_t* bounds) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (q_set < 0 || q_set >= kMaxSets || t_set < 0 || t_set >= kMaxSets || !bounds || n_shards == 0)
        return fail(ctx, BSA_ERR_BAD_ARG, "bad argument");
    const SeqSet &Q = ctx->sets[q_set], &T = ctx->sets[t_set];
    if (!Q.loaded || !T.loaded) return fail(ctx, BSA_ERR_BAD_ARG, "sequence set not loaded");
    std::vector<double> pre((size_t)T.n + 1, 0.0);
    for (uint32_t t = 0; t < T.n; ++t) {
        const uint32_t cnt = q_counts ? std::min(q_counts[t], Q.n) : Q.n;
        pre[t + 1] = pre[t] + (double)T.len(t) * (double)Q.off[cnt];
    }
    bounds[0] = 0;
    for (uint32_t r = 1; r < n_shards; ++r) {
        const double want = pre[T.n] * r / n_shards;
        bounds[r] = (uint32_t)(std::lower_bound(pre.begin(), pre.end(), want) - pre.begin());
        bounds[r] = std::max(bounds[r], bounds[r - 1]);
        bounds[r] = std::min(bounds[r], T.n);
    }
    bounds[n_shards] = T.n;
    return BSA_OK;
}

// A child of a multi-device context holds its GPU's gate from its first launch until its kernels have finished:
// the other child of that GPU plans its next tile meanwhile and copies its last results afterwards.
struct GateGuard {
    std::mutex* m = nullptr;
    void lock(std::mutex* g) { if (g) { g->lock(); m = g; } }
    void unlock() { if (m) { m->unlock(); m = nullptr; } }
    ~GateGuard() { unlock(); }
};

int bsa_align_all_pairs(bsa_ctx* ctx, int q_set, int t_set, const uint32_t* q_counts, uint32_t t_begin,
                        uint32_t t_end, uint32_t flags, int32_t* scores, uint32_t* n_identical,
                        uint64_t* n_results) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (ctx->multi)
        return multi_align_all_pairs(ctx, q_set, t_set, q_counts, t_begin, t_end, flags, scores, n_identical, n_results);
    const auto wall0 = std::chrono::steady_clock::now();
    if (q_set < 0 || q_set >= kMaxSets || t_set < 0 || t_set >= kMaxSets)
        return fail(ctx, BSA_ERR_BAD_ARG, "bad set id");
    const SeqSet &Q = ctx->sets[q_set], &T = ctx->sets[t_set];
    if (!Q.loaded || !T.loaded) return fail(ctx, BSA_ERR_EMPTY, "sequence set not loaded");
    if (t_begin > t_end || t_end > T.n) return fail(ctx, BSA_ERR_BAD_ARG, "bad template range");
    CK(cudaSetDevice(ctx->device));
    int rc = sync_scoring(ctx);
    if (rc) return rc;
    const bool want_s = (flags & BSA_WANT_SCORE) && scores;
    const bool want_i = (flags & BSA_WANT_IDENTICAL) && n_identical;
    const bool out_dev = (flags & BSA_OUT_DEVICE) != 0;
    const int C = std::max(ctx->ncodes, 1);
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->stats.h2d_bytes = ctx->pending_h2d;   // sequence uploads since the last call belong to this one
    ctx->pending_h2d = 0;

    // ---------------- plan ----------------
    std::vector<uint64_t> first((size_t)(t_end - t_begin) + 1, 0);
    double total_cells = 0.0;
    for (uint32_t t = t_begin; t < t_end; ++t) {
        const uint32_t cnt = q_counts ? q_counts[t] : Q.n;
        if (cnt > Q.n) return fail(ctx, BSA_ERR_BAD_ARG, "q_counts entry exceeds the query set size");
        first[t - t_begin + 1] = first[t - t_begin] + cnt;
        total_cells += (double)T.len(t) * (double)Q.off[cnt];
    }
    const uint64_t n_res = first.back();
    if (n_results) *n_results = n_res;
    ctx->stats.pairs = n_res;
    ctx->stats.cells = (uint64_t)total_cells;
    if (n_res == 0) return BSA_OK;

    // item size: ~40,000 items for a large problem; small problems get fewer, larger items (the ~100 kernel groups
    // run on concurrent streams and need not fill the GPU one by one, and the per-item costs -- profile build,
    // barriers, pipeline fill of short chunks -- are what such a run is made of)
    const char* min_item_env = getenv("BSA_MIN_ITEM_CELLS");        // A/B only
    const double min_item = min_item_env ? atof(min_item_env) : 4194304.0;
    const char* n_items_env = getenv("BSA_TARGET_ITEMS");           // A/B only
    const double n_items_target = n_items_env ? std::max(1000.0, atof(n_items_env)) : 40000.0;
    const double target_cells = std::min(std::max(total_cells / n_items_target, min_item), 268435456.0);
    struct Group { std::vector<Item> items; uint64_t stride = 0; uint64_t scr_off = 0; double cells = 0, swept = 0; };
    const bool use_tag = !getenv("BSA_NO_TAG");
    std::vector<Fix> fixes;
    std::vector<PairReq> fallback;
    double padded = 0.0;
    const int cs_cap = bitlen(Q.maxlen);

    // ---- score-only requests: pairs of templates in 16-bit packed lanes (gotoh_score16_kernel) ----
    struct Group16 { std::vector<Item16> items; uint64_t stride = 0; uint64_t scr_off = 0; double cells = 0; };
    std::vector<Group16> groups16(2 * (kKMax + 1));
    std::vector<uint8_t> done16((size_t)(t_end - t_begin), 0);
    // short templates (<= 16 x 20 columns): FOUR per warp on two 16-lane pipelines (gotoh_score16_quad_kernel)
    struct Group16Q { std::vector<Item16Q> items; double cells = 0; };
    std::vector<Group16Q> groups16Q(kKMax + 1);
    if (want_s && !want_i && Q.empties.empty() && !getenv("BSA_NO_S16")) {
        struct Cand { uint32_t t, cnt; uint64_t m; int key; int pkey; };     // pkey: key of the two-per-warp path, -1 = not eligible
        std::vector<Cand> cands, qcands;
        const bool quad_on = BSA_ALIGNED && !getenv("BSA_NO_QUAD16");
        for (uint32_t t = t_begin; t < t_end; ++t) {
            const uint32_t cnt = q_counts ? q_counts[t] : Q.n;
            const uint64_t m = T.len(t);
            if (cnt == 0 || m == 0) continue;
            // two templates per warp (32 lanes, gotoh_score16_kernel)
            int pkey = -1;
            const KChoice kc = choose_k(m, C);
            if (kc.K >= 1) {
                // every DP value (incl. borders, E/F one step below them, and the padded columns of the
                // shorter template of a pair) must fit a 16-bit lane; the biased low half must also stay
                // >= |go| so that the 32-bit `h + GO` never borrows from the high half
                const uint64_t m_pad16 = 32ull * kc.K * kc.npass;
                const int64_t ub = (int64_t)std::max(ctx->max_m, 0) * (int64_t)std::min<uint64_t>(m_pad16, Q.maxlen);
                const int64_t lb = 5 * (int64_t)(-ctx->go) + (int64_t)(Q.maxlen + m_pad16 + 68) * (int64_t)(-ctx->ge) +
                                   (int64_t)std::max(-ctx->min_m, 0);
                bool ok;
                if (BSA_ALIGNED && !kc.multi) {
                    // single-block templates run in the moving frame score - (i + j) ge (stream_block16_fa): the frame
                    // adds up to (n + m) |ge| at the top; at the bottom nothing falls below three openings under a
                    // substitution score
                    const int64_t ubf = ub + (int64_t)(Q.maxlen + m_pad16 + 8) * (int64_t)(-ctx->ge) + (int64_t)std::max(ctx->max_m, 0);
                    const int64_t lbf = 4 * (int64_t)(ctx->ge - ctx->go) + 4 * (int64_t)(-ctx->ge) + (int64_t)std::max(-ctx->min_m, 0);
                    ok = std::max(ubf, lbf) < 32000;
                } else ok = std::max(ub, lb) < 32000;
                if (ok) pkey = kc.K + (kc.multi ? kKMax + 1 : 0) + (int)kc.npass * 256;
            }
            // four short templates per warp (two 16-lane pipelines, gotoh_score16_quad_kernel)
            if (quad_on && m <= 16ull * kKStream) {
                const int K16 = (int)((m + 15) / 16);
                const uint64_t m_pad = 16ull * K16;
                const int64_t ubq = (int64_t)std::max(ctx->max_m, 0) * (int64_t)std::min<uint64_t>(m_pad, Q.maxlen) +
                                    (int64_t)(Q.maxlen + m_pad + 8) * (int64_t)(-ctx->ge) + (int64_t)std::max(ctx->max_m, 0);
                const int64_t lbq = 4 * (int64_t)(ctx->ge - ctx->go) + 4 * (int64_t)(-ctx->ge) + (int64_t)std::max(-ctx->min_m, 0);
                if (std::max(ubq, lbq) < 32000 && smem_for(K16, C) <= kSmemBudget) {
                    qcands.push_back(Cand{t, cnt, m, K16, pkey});
                    continue;
                }
            }
            if (pkey >= 0) cands.push_back(Cand{t, cnt, m, pkey, pkey});
        }
        // quads: templates of the same 16-column class and the same query count, four (or the last three) per item;
        // what is left over -- one or two of a kind, e.g. every template of a triangle, whose query counts all
        // differ -- goes two (or one) per warp as before
        std::stable_sort(qcands.begin(), qcands.end(), [](const Cand& a, const Cand& b) {
            if (a.key != b.key) return a.key < b.key;
            if (a.cnt != b.cnt) return a.cnt < b.cnt;
            return a.m < b.m;
        });
        for (size_t i = 0; i < qcands.size();) {
            const Cand& A = qcands[i];
            size_t nq = 1;
            while (nq < 4 && i + nq < qcands.size() && qcands[i + nq].key == A.key && qcands[i + nq].cnt == A.cnt) ++nq;
            if (nq < 3) {
                for (size_t k = 0; k < nq; ++k)
                    if (qcands[i + k].pkey >= 0) { Cand c = qcands[i + k]; c.key = c.pkey; cands.push_back(c); }
                i += nq;
                continue;
            }
            const int K16 = A.key;
            const uint64_t m_pad = 16ull * K16;
            uint64_t xb = (uint64_t)std::max(1.0, target_cells / (double)m_pad);
            xb = std::min<uint64_t>(xb, 1u << 18);
            double msum = 0;
            for (size_t k = 0; k < nq; ++k) msum += (double)qcands[i + k].m;
            uint32_t q = 0;
            while (q < A.cnt) {
                const uint64_t lim_off = Q.off[q] + xb;
                uint32_t q2 = (uint32_t)(std::upper_bound(Q.off.begin() + q + 1, Q.off.begin() + A.cnt + 1, lim_off) -
                                         Q.off.begin()) - 1;
                q2 = std::min(std::max(q2, q + 1), A.cnt);
                Item16Q it;
                for (size_t k = 0; k < 4; ++k) {
                    it.t[k] = k < nq ? qcands[i + k].t : 0xffffffffu;
                    it.out[k] = k < nq ? first[qcands[i + k].t - t_begin] + q : 0;
                }
                it.q_begin = q; it.q_end = q2;
                groups16Q[K16].items.push_back(it);
                const uint64_t x = Q.off[q2] - Q.off[q];
                padded += (double)(x + 15.0 * std::max<double>(kWarpsPerCta, (double)x / 3072.0)) * (double)m_pad * (double)nq;
                groups16Q[K16].cells += (double)x * msum;
                q = q2;
            }
            for (size_t k = 0; k < nq; ++k) done16[qcands[i + k].t - t_begin] = 1;
            i += nq;
        }
        // partners must agree on columns per lane, number of column blocks and query count
        std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
            if (a.key != b.key) return a.key < b.key;
            if (a.cnt != b.cnt) return a.cnt < b.cnt;
            return a.m < b.m;
        });
        for (size_t i = 0; i < cands.size();) {
            const Cand& A = cands[i];
            const bool pair = i + 1 < cands.size() && cands[i + 1].key == A.key && cands[i + 1].cnt == A.cnt;
            const Cand* B = pair ? &cands[i + 1] : nullptr;
            const KChoice kc = choose_k(A.m, C);
            const uint64_t m_pad = 32ull * kc.K * kc.npass;
            uint64_t xb = (uint64_t)std::max(1.0, 2.0 * target_cells / (double)m_pad);
            xb = std::min<uint64_t>(xb, 1u << 18);
            Group16& grp = groups16[kc.K + (kc.multi ? kKMax + 1 : 0)];
            uint32_t q = 0;
            while (q < A.cnt) {
                const uint64_t lim_off = Q.off[q] + xb;
                uint32_t q2 = (uint32_t)(std::upper_bound(Q.off.begin() + q + 1, Q.off.begin() + A.cnt + 1, lim_off) -
                                         Q.off.begin()) - 1;
                q2 = std::min(std::max(q2, q + 1), A.cnt);
                Item16 it;
                it.tA = A.t; it.tB = B ? B->t : 0xffffffffu;
                it.q_begin = q; it.q_end = q2;
                it.outA = first[A.t - t_begin] + q;
                it.outB = B ? first[B->t - t_begin] + q : 0;
                grp.items.push_back(it);
                const uint64_t x = Q.off[q2] - Q.off[q];
                padded += (double)(x + 31.0 * std::max<double>(kWarpsPerCta, (double)x / 3072.0)) * (double)m_pad * (B ? 2.0 : 1.0);
                grp.cells += (double)x * (double)(A.m + (B ? B->m : 0));
                if (kc.multi) grp.stride = std::max(grp.stride, x + 64);
                q = q2;
            }
            done16[A.t - t_begin] = 1;
            if (B) done16[B->t - t_begin] = 1;
            i += pair ? 2 : 1;
        }
    }

    // =====================================================================================
    // 32-bit lanes (score + identity, and score-only templates the 16-bit lanes cannot take).
    //
    // The streaming kernels keep ONE sequence -- the OWNER -- in registers as DP columns and stream
    // the others through the lanes as rows.  Whoever owns the columns, the pair (q, t) yields the
    // same result as long as the tie priorities follow the reference's orientation (rows = query):
    // with the roles swapped, E and F trade their H-max priorities (`flip`).  Three variants:
    //   A  owner = template t, stream = the query set                         (the reference's layout)
    //   B  owner = a SHORT query q, stream = the LONG templates with q_counts[t] > q, flipped:
    //      a template of more than one column block (> 32 x 20 columns) would sweep every query in
    //      several passes at ~70 % of the single-pass rate; as a ROW stream it costs nothing extra
    //   C  owner = a long template, stream = the LONG queries only (long x long, what is left)
    // B and C stream from derived stores that hold only the long sequences (gathered on the
    // device, original order kept), and scatter their results through a per-stream-sequence table:
    // k = item.out_base + lut[stream index].  B needs q_counts to be non-decreasing (a suffix of the
    // long templates per query); otherwise everything stays in A.
    // =====================================================================================
    struct GroupPair { std::vector<Item16> items; double cells = 0, swept = 0; };
    struct Variant {
        std::vector<Group> groups;          // K + multi * stride + tag * 2 stride
        std::vector<GroupPair> pairs;       // K + tag * stride
        const std::vector<uint64_t>* xoff = nullptr;   // offsets of the stream set (host copy)
        uint64_t xmax = 0;                  // its longest sequence
        int cs_cap = 0;                     // count-field cap: bitlen(xmax)
        SeqStoreDev qdev, tdev;             // stream / owner store on the device
        SeqStoreDev qadev;                  // even-aligned copy of the stream store (two-row TAG kernels)
        const uint64_t* lut = nullptr;      // device: result-index contribution per stream sequence
        int flip = 0;
        const char* name = "A";
    };
    Variant var[3];
    for (auto& V : var) { V.groups.resize(4 * kGroupStride); V.pairs.resize(2 * kGroupStride); }
    var[0].xoff = &Q.off; var[0].xmax = Q.maxlen; var[0].cs_cap = cs_cap;
    var[0].qdev = Q.dev(); var[0].tdev = T.dev();
    if (BSA_ALIGNED) {
        rc = ensure_aligned(ctx, Q);      // the two-row kernels (32-bit TAG and 16-bit frame) stream the aligned copy
        if (rc) return rc;
        var[0].qadev = Q.adev();
    }
    var[1].name = "B"; var[1].flip = 1;
    var[2].name = "C";

    const int kc_eff = std::min(k_cap(C), kKStream);
    const uint64_t Lc = 32ull * (uint64_t)std::max(kc_eff, 1);      // longest single-pass owner
    auto count_of = [&](uint32_t t) { return q_counts ? q_counts[t] : Q.n; };

    // ---- which templates leave variant A? ----
    std::vector<uint32_t> lt, lq;            // original indices of the long templates / long queries
    std::vector<uint64_t> ltoff, lqoff;      // offsets of the two derived stores
    std::vector<uint8_t> in_lt((size_t)(t_end - t_begin), 0);
    {
        bool hybrid = Q.empties.empty() && kc_eff >= 1 && !getenv("BSA_NO_HYBRID");
        for (uint32_t t = t_begin; hybrid && t + 1 < t_end; ++t)
            if (count_of(t + 1) < count_of(t)) hybrid = false;      // B needs a suffix per query
        if (hybrid) {
            for (uint32_t t = t_begin; t < t_end; ++t)
                if (T.len(t) > Lc && count_of(t) > 0 && !done16[t - t_begin]) lt.push_back(t);
            if (!lt.empty()) {
                const uint32_t max_cnt = count_of(lt.back());
                for (uint32_t q = 0; q < max_cnt; ++q)
                    if (Q.len(q) > Lc) lq.push_back(q);
                for (uint32_t t : lt) in_lt[t - t_begin] = 1;
            }
        }
    }
    std::vector<uint64_t> lutB, lutC;
    if (!lt.empty()) {
        // derived stores: the long sequences back to back, last residues keep their flags
        auto gather = [&](const SeqSet& S, const std::vector<uint32_t>& idx, std::vector<uint64_t>& off, DevBuf& codes,
                          DevBuf& doff, DevBuf& didx, uint64_t* maxlen) -> int {
            off.assign(idx.size() + 1, 0);
            *maxlen = 0;
            for (size_t i = 0; i < idx.size(); ++i) {
                off[i + 1] = off[i] + S.len(idx[i]);
                *maxlen = std::max(*maxlen, S.len(idx[i]));
            }
            CK(codes.ensure(kFrontPad + off.back() + kBackPad));
            CK(doff.ensure(off.size() * 8));
            CK(didx.ensure(std::max<size_t>(idx.size(), 1) * 4));
            cudaStream_t st = ctx->streams[0];
            CK(cudaMemcpyAsync(doff.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(codes.p, 0, kFrontPad, st));
            CK(cudaMemsetAsync(codes.as<uint8_t>() + kFrontPad - 1, (int)kLastFlag, 1, st));
            CK(cudaMemsetAsync(codes.as<uint8_t>() + kFrontPad + off.back(), 0, kBackPad, st));
            if (!idx.empty()) {
                CK(cudaMemcpyAsync(didx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, st));
                gather_seqs_kernel<<<(uint32_t)idx.size(), 128, 0, st>>>(S.dev(), didx.as<uint32_t>(), doff.as<uint64_t>(),
                                                                        codes.as<uint8_t>() + kFrontPad);
                CK(cudaGetLastError());
                ctx->stats.launches++;
            }
            ctx->stats.h2d_bytes += off.size() * 8 + idx.size() * 4;
            CK(cudaStreamSynchronize(st));     // idx / off are read by the copies above
            return BSA_OK;
        };
        uint64_t mxT = 0, mxQ = 0;
        rc = gather(T, lt, ltoff, ctx->lt_codes, ctx->lt_off, ctx->lt_idx, &mxT);
        if (rc) return rc;
        rc = gather(Q, lq, lqoff, ctx->lq_codes, ctx->lq_off, ctx->lq_idx, &mxQ);
        if (rc) return rc;
        lutB.resize(lt.size());
        for (size_t i = 0; i < lt.size(); ++i) lutB[i] = first[lt[i] - t_begin];
        lutC.resize(std::max<size_t>(lq.size(), 1), 0);
        for (size_t i = 0; i < lq.size(); ++i) lutC[i] = lq[i];
        CK(ctx->lutB.ensure(lutB.size() * 8));
        CK(ctx->lutC.ensure(lutC.size() * 8));
        CK(cudaMemcpy(ctx->lutB.p, lutB.data(), lutB.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->lutC.p, lutC.data(), lutC.size() * 8, cudaMemcpyHostToDevice));
        ctx->stats.h2d_bytes += (lutB.size() + lutC.size()) * 8;
        SeqStoreDev dT, dQ;
        dT.codes = ctx->lt_codes.as<uint8_t>() + kFrontPad; dT.off = ctx->lt_off.as<uint64_t>(); dT.n = (uint32_t)lt.size();
        dQ.codes = ctx->lq_codes.as<uint8_t>() + kFrontPad; dQ.off = ctx->lq_off.as<uint64_t>(); dQ.n = (uint32_t)lq.size();
        var[1].xoff = &ltoff; var[1].xmax = mxT; var[1].cs_cap = bitlen(mxT);
        if (use_tag) {
            rc = build_aligned(ctx, dT, ltoff, ctx->lt_acodes, ctx->lt_aoff, ctx->streams[0]);
            if (rc) return rc;
            rc = build_aligned(ctx, dQ, lqoff, ctx->lq_acodes, ctx->lq_aoff, ctx->streams[0]);
            if (rc) return rc;
            var[1].qadev = SeqStoreDev{ctx->lt_acodes.as<uint8_t>() + kFrontPad, ctx->lt_aoff.as<uint64_t>(), dT.n};
            var[2].qadev = SeqStoreDev{ctx->lq_acodes.as<uint8_t>() + kFrontPad, ctx->lq_aoff.as<uint64_t>(), dQ.n};
        }
        var[1].qdev = dT; var[1].tdev = Q.dev(); var[1].lut = ctx->lutB.as<uint64_t>();
        var[2].xoff = &lqoff; var[2].xmax = mxQ; var[2].cs_cap = bitlen(mxQ);
        var[2].qdev = dQ; var[2].tdev = T.dev(); var[2].lut = ctx->lutC.as<uint64_t>();
    }
    // original (query, template, result index) of stream sequence x of an owner
    auto orig_pair = [&](int v, uint32_t owner, uint32_t x, uint64_t out_base) {
        if (v == 0) return PairReq{x, owner, out_base + x};
        if (v == 1) return PairReq{owner, lt[x], out_base + lutB[x]};
        return PairReq{lq[x], owner, out_base + lutC[x]};
    };

    // ---- one owner against the stream sequences [x_lo, x_hi): items of at most xb stream residues ----
    auto plan_owner = [&](int v, uint32_t owner, uint64_t m, uint32_t x_lo, uint32_t x_hi, uint64_t out_base) -> int {
        Variant& V = var[v];
        const std::vector<uint64_t>& xo = *V.xoff;
        if (x_lo >= x_hi) return BSA_OK;
        const KChoice kc = choose_k(m, C);
        if (kc.K < 1) return fail(ctx, BSA_ERR_ALPHABET, "alphabet too large for shared memory");
        // integer-field check: score << (cs+2) must stay inside int32 for every cell
        const int cs = std::min(bitlen(m), V.cs_cap);
        const int64_t lim = cs <= 27 ? (int64_t)1 << (29 - cs) : 0;
        const int64_t lim_tag = cs + kTagBits <= 27 ? (int64_t)1 << (29 - cs - kTagBits) : 0;
        const uint64_t m_pad = 32ull * kc.K * kc.npass;
        const int64_t ub = (int64_t)std::max(ctx->max_m, 0) * (int64_t)std::min<uint64_t>(m, V.xmax);
        const int64_t lb = 4 * (int64_t)(-ctx->go) + (int64_t)(V.xmax + m + 4) * (int64_t)(-ctx->ge) +
                           (int64_t)std::max(-ctx->min_m, 0);
        const bool fits = std::max(ub, lb) + 8 < lim;
        // TAG cells live in the moving frame score - (i + j) ge: [-(4 |go| + ...), max(M) min(n, m) + (n + m) |ge|]
        const int64_t ubf = ub + (int64_t)(V.xmax + m_pad + 10) * (int64_t)(-ctx->ge) + (int64_t)std::max(ctx->max_m, 0);
        const int64_t lbf = 4 * (int64_t)(-ctx->go) + 8 * (int64_t)(-ctx->ge) + (int64_t)std::max(-ctx->min_m, 0);
        // single-block TAG kernels size the count field for their columns-per-lane class (a launch constant:
        // gotoh_stream_kernel, kUniformCs), which can be one bit more than this template needs
        const int cs_k = (BSA_ALIGNED && BSA_ETAG && !kc.multi) ? std::min(bitlen(32ull * kc.K), V.cs_cap) : cs;
        const int64_t lim_tag_k = cs_k + kTagBits <= 27 ? (int64_t)1 << (29 - cs_k - kTagBits) : 0;
        const bool fits_tag = use_tag && std::max(ubf, lbf) + 8 < std::min(lim_tag, lim_tag_k);   // room for the tag field too
        // short owners: keep items small enough that their group still fills the GPU;
        // multi-pass: this also bounds the per-CTA boundary slice (2 MiB)
        uint64_t xb = (uint64_t)std::max(1.0, target_cells / (double)m_pad);
        xb = std::min<uint64_t>(xb, 1u << 18);
        auto emit_items = [&](uint32_t lo, uint32_t hi, bool tag) {
            uint32_t q = lo;
            while (q < hi) {
                const uint64_t lim_off = xo[q] + xb;
                uint32_t q2 = (uint32_t)(std::upper_bound(xo.begin() + q + 1, xo.begin() + hi + 1, lim_off) - xo.begin()) - 1;
                q2 = std::max(q2, q + 1);
                q2 = std::min(q2, hi);
                Item it;
                it.t = owner; it.q_begin = q; it.q_end = q2; it.cshift = (uint32_t)cs;
                it.out_base = V.lut ? out_base : out_base + q;
                Group& grp = V.groups[group_index(kc.K, kc.multi, tag)];
                grp.items.push_back(it);
                const uint64_t x = xo[q2] - xo[q];
                const double sw = (double)(x + 31.0 * std::max<double>(kWarpsPerCta, (double)x / 3072.0)) * (double)m_pad;
                padded += sw;
                grp.swept += sw;
                grp.cells += (double)x * (double)m;
                if (kc.multi) grp.stride = std::max(grp.stride, x + 64);   // boundary column of one item
                q = q2;
            }
        };
        // runs of non-empty stream sequences (only the query set of variant A can hold empty ones)
        auto e_it = Q.empties.begin();
        uint32_t run_b = x_lo;
        while (run_b < x_hi) {
            uint32_t run_e = x_hi;
            if (v == 0) {
                while (e_it != Q.empties.end() && *e_it < run_b) ++e_it;
                if (e_it != Q.empties.end() && *e_it < x_hi) run_e = *e_it;
            }
            if (run_e == run_b) {
                // empty query: the top border value go + (m-1) ge (global.rs:81-88,143)
                fixes.push_back(Fix{out_base + run_b, (int32_t)(ctx->go + (int64_t)(m - 1) * ctx->ge), 0u});
                ++run_b;
                continue;
            }
            if (fits_tag) emit_items(run_b, run_e, true);
            else if (fits) emit_items(run_b, run_e, false);
            else for (uint32_t x = run_b; x < run_e; ++x) fallback.push_back(orig_pair(v, owner, x, out_base));
            run_b = run_e;
        }
        return BSA_OK;
    };

    // ---- short owners: two per warp on 16 lanes each (gotoh_pair_kernel); what cannot be paired goes to `rest` ----
    struct POwner { uint32_t idx; uint64_t m; uint32_t lo, hi; uint64_t out_base; };
    auto plan_pairs = [&](int v, const std::vector<POwner>& owners, std::vector<POwner>& rest) {
        Variant& V = var[v];
        const std::vector<uint64_t>& xo = *V.xoff;
        std::vector<std::vector<POwner>> by_k(2 * kGroupStride);
        const bool enabled = Q.empties.empty() && !getenv("BSA_NO_PAIR");
        for (const POwner& o : owners) {
            const uint64_t m = o.m;
            bool ok = enabled && m <= 16ull * kKStream && smem_for((int)((m + 15) / 16), C) <= kSmemBudget;
            bool tag = false;
            if (ok) {
                // the kernel sizes the count field for the longer owner of a pair: check with the
                // widest field any owner of this columns-per-lane class can meet
                const uint64_t m_class = 16ull * ((m + 15) / 16);
                const int cs = std::min(bitlen(m_class), V.cs_cap);
                const int64_t lim = cs <= 27 ? (int64_t)1 << (29 - cs) : 0;
                const int64_t lim_tag = cs + kTagBits <= 27 ? (int64_t)1 << (29 - cs - kTagBits) : 0;
                const int64_t ub = (int64_t)std::max(ctx->max_m, 0) * (int64_t)std::min<uint64_t>(m_class, V.xmax);
                const int64_t lb = 4 * (int64_t)(-ctx->go) + (int64_t)(V.xmax + m_class + 4) * (int64_t)(-ctx->ge) +
                                   (int64_t)std::max(-ctx->min_m, 0);
                const int64_t ubf = ub + (int64_t)(V.xmax + m_class + 10) * (int64_t)(-ctx->ge) + (int64_t)std::max(ctx->max_m, 0);
                const int64_t lbf = 4 * (int64_t)(-ctx->go) + 8 * (int64_t)(-ctx->ge) + (int64_t)std::max(-ctx->min_m, 0);
                ok = std::max(ub, lb) + 8 < lim;
                tag = use_tag && std::max(ubf, lbf) + 8 < lim_tag;
            }
            if (ok) by_k[(m + 15) / 16 + (tag ? kGroupStride : 0)].push_back(o);
            else rest.push_back(o);
        }
        for (int gk = 1; gk < 2 * kGroupStride; ++gk) {
            const int K = gk % kGroupStride;
            const auto& vv = by_k[gk];
            if (K < 1 || K > kKStream) { rest.insert(rest.end(), vv.begin(), vv.end()); continue; }
            size_t i = 0;
            for (; i + 1 < vv.size(); i += 2) {
                const POwner &A = vv[i], &B = vv[i + 1];
                const uint32_t clo = std::max(A.lo, B.lo), chi = std::min(A.hi, B.hi);
                if (clo >= chi) { rest.push_back(A); rest.push_back(B); continue; }
                const uint64_t m_pad = 32ull * K;
                uint64_t xb = (uint64_t)std::max(1.0, target_cells / (double)m_pad);
                xb = std::min<uint64_t>(xb, 1u << 18);
                uint32_t q = clo;
                while (q < chi) {
                    const uint64_t lim_off = xo[q] + xb;
                    uint32_t q2 = (uint32_t)(std::upper_bound(xo.begin() + q + 1, xo.begin() + chi + 1, lim_off) - xo.begin()) - 1;
                    q2 = std::min(std::max(q2, q + 1), chi);
                    Item16 it;
                    it.tA = A.idx; it.tB = B.idx; it.q_begin = q; it.q_end = q2;
                    it.outA = V.lut ? A.out_base : A.out_base + q;
                    it.outB = V.lut ? B.out_base : B.out_base + q;
                    V.pairs[gk].items.push_back(it);
                    const uint64_t x = xo[q2] - xo[q];
                    const double sw = (double)(x + 15.0 * std::max<double>(kWarpsPerCta, (double)x / 3072.0)) * (double)m_pad;
                    padded += sw;
                    V.pairs[gk].swept += sw;
                    V.pairs[gk].cells += (double)x * (double)(A.m + B.m);
                    q = q2;
                }
                // what is left of either stream range goes the one-owner way
                for (const POwner* o : {&A, &B}) {
                    if (o->lo < clo) rest.push_back(POwner{o->idx, o->m, o->lo, clo, o->out_base});
                    if (chi < o->hi) rest.push_back(POwner{o->idx, o->m, chi, o->hi, o->out_base});
                }
            }
            if (i < vv.size()) rest.push_back(vv[i]);
        }
    };

    {
        std::vector<POwner> ownA, ownB, rest;
        // variant A: every template that stays; empty ones are border values only
        for (uint32_t t = t_begin; t < t_end; ++t) {
            const uint32_t cnt = count_of(t);
            if (cnt == 0 || done16[t - t_begin] || in_lt[t - t_begin]) continue;
            const uint64_t m = T.len(t), kbase = first[t - t_begin];
            if (m == 0) {
                // empty template: the reference returns the row-0 border value... for an empty
                // query against it the loop never runs: H[0]=0 (global.rs:76,143); otherwise
                // the left border go + (n-1) ge after n rows (global.rs:99,140).
                for (uint32_t q = 0; q < cnt; ++q) {
                    const uint64_t n = Q.len(q);
                    fixes.push_back(Fix{kbase + q, n == 0 ? 0 : (int32_t)(ctx->go + (int64_t)(n - 1) * ctx->ge), 0u});
                }
                continue;
            }
            ownA.push_back(POwner{t, m, 0u, cnt, kbase});
        }
        plan_pairs(0, ownA, rest);
        for (const POwner& o : rest) { rc = plan_owner(0, o.idx, o.m, o.lo, o.hi, o.out_base); if (rc) return rc; }
        if (!lt.empty()) {
            // variant C: long template x its long queries (a prefix of the long-query store)
            for (size_t li = 0; li < lt.size(); ++li) {
                const uint32_t t = lt[li];
                const uint32_t hi = (uint32_t)(std::lower_bound(lq.begin(), lq.end(), count_of(t)) - lq.begin());
                rc = plan_owner(2, t, T.len(t), 0u, hi, first[t - t_begin]);
                if (rc) return rc;
            }
            // variant B: short query x the long templates that take it (a suffix of the long-template store)
            const uint32_t max_cnt = count_of(lt.back());
            size_t lo = 0;
            for (uint32_t q = 0; q < max_cnt; ++q) {
                while (lo < lt.size() && count_of(lt[lo]) <= q) ++lo;
                if (lo >= lt.size()) break;
                const uint64_t n = Q.len(q);
                if (n > Lc) continue;                    // long x long: variant C
                ownB.push_back(POwner{q, n, (uint32_t)lo, (uint32_t)lt.size(), (uint64_t)q});
            }
            rest.clear();
            plan_pairs(1, ownB, rest);
            for (const POwner& o : rest) { rc = plan_owner(1, o.idx, o.m, o.lo, o.hi, o.out_base); if (rc) return rc; }
        }
    }
    ctx->stats.padded_cells = (uint64_t)padded;
    ctx->stats.fallback_pairs = (uint32_t)std::min<size_t>(fallback.size(), 0xffffffffu);

    // ---------------- outputs ----------------
    int32_t* d_scores = nullptr;
    uint32_t* d_nid = nullptr;
    // Page-locked result buffers (bsa_host_alloc_pinned, or any cudaHostAlloc / pinned memory): the kernels store
    // every result straight into the caller's buffers over the host link -- 8 bytes per pair, fire-and-forget
    // stores -- so there is no staging copy in HBM and no device-to-host copy at the end of the call
    // (BSA_MAPPED_OUT=0 keeps the staged copy: A/B).  Pageable buffers are staged and copied as before.
    bool mapped_out = false;
    const char* mo_env = getenv("BSA_MAPPED_OUT");
    if (!out_dev && !(mo_env && mo_env[0] == '0')) {
        auto dev_alias = [](void* p) -> void* {
            if (!p) return nullptr;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
        };
        void* ds = want_s ? dev_alias(scores) : nullptr;
        void* dn = want_i ? dev_alias(n_identical) : nullptr;
        if ((!want_s || ds) && (!want_i || dn)) { mapped_out = true; d_scores = (int32_t*)ds; d_nid = (uint32_t*)dn; }
    }
    if (out_dev) {
        d_scores = want_s ? scores : nullptr;
        d_nid = want_i ? n_identical : nullptr;
    } else if (mapped_out) {
    } else {
        if (want_s) { CK(ctx->out_scores.ensure(n_res * 4)); d_scores = ctx->out_scores.as<int32_t>(); }
        if (want_i) { CK(ctx->out_nid.ensure(n_res * 4)); d_nid = ctx->out_nid.as<uint32_t>(); }
    }

    // ---------------- upload the plan, launch ----------------
    // one launch per non-empty group; long owners (large K, multi-pass) first: they are the expensive items
    struct Launch { int v; bool pair; int g; size_t item_off; uint32_t n_items; uint64_t scr_off; };
    std::vector<Launch> launches;
    std::vector<Item> all;
    std::vector<Item16> all_pair;
    for (int v : {2, 0, 1}) {
        std::vector<int> gorder;
        for (int g = (int)var[v].groups.size() - 1; g >= 0; --g) if (!var[v].groups[g].items.empty()) gorder.push_back(g);
        std::stable_sort(gorder.begin(), gorder.end(), [&](int a, int b) {
            return group_k(a) * (group_multi(a) ? 64 : 1) > group_k(b) * (group_multi(b) ? 64 : 1);
        });
        for (int g : gorder) {
            launches.push_back(Launch{v, false, g, all.size(), (uint32_t)var[v].groups[g].items.size(), 0});
            all.insert(all.end(), var[v].groups[g].items.begin(), var[v].groups[g].items.end());
        }
    }
    for (int v : {0, 1})
        for (int gk = (int)var[v].pairs.size() - 1; gk >= 0; --gk) {
            if (var[v].pairs[gk].items.empty()) continue;
            launches.push_back(Launch{v, true, gk, all_pair.size(), (uint32_t)var[v].pairs[gk].items.size(), 0});
            all_pair.insert(all_pair.end(), var[v].pairs[gk].items.begin(), var[v].pairs[gk].items.end());
        }
    ctx->stats.items = (uint32_t)(all.size() + all_pair.size());
    constexpr size_t kCounters16 = 1024;     // the 16-bit groups' counters follow those of the launches above
    const size_t n_counters = kCounters16 + groups16.size() + groups16Q.size();
    CK(ctx->counters.ensure(n_counters * 4));
    if (launches.size() > kCounters16) return fail(ctx, BSA_ERR_CUDA, "too many kernel groups");
    cudaStream_t s0 = ctx->streams[0];
    CK(cudaMemsetAsync(ctx->counters.p, 0, n_counters * 4, s0));
    // every item list goes up before the start event, so that every stream only has to wait for that one event
    std::vector<size_t> goff16(groups16.size(), 0), goff16Q(groups16Q.size(), 0);
    {
        if (!all.empty()) {
            CK(ctx->items.ensure(all.size() * sizeof(Item)));
            CK(cudaMemcpyAsync(ctx->items.p, all.data(), all.size() * sizeof(Item), cudaMemcpyHostToDevice, s0));
            ctx->stats.h2d_bytes += all.size() * sizeof(Item);
        }
        if (!all_pair.empty()) {
            CK(ctx->items_pair.ensure(all_pair.size() * sizeof(Item16)));
            CK(cudaMemcpyAsync(ctx->items_pair.p, all_pair.data(), all_pair.size() * sizeof(Item16), cudaMemcpyHostToDevice, s0));
            ctx->stats.h2d_bytes += all_pair.size() * sizeof(Item16);
        }
        std::vector<Item16> all16;
        std::vector<Item16Q> all16q;
        for (int g = (int)groups16.size() - 1; g >= 0; --g) {
            goff16[g] = all16.size();
            all16.insert(all16.end(), groups16[g].items.begin(), groups16[g].items.end());
        }
        if (!all16.empty()) {
            CK(ctx->items16.ensure(all16.size() * sizeof(Item16)));
            CK(cudaMemcpyAsync(ctx->items16.p, all16.data(), all16.size() * sizeof(Item16), cudaMemcpyHostToDevice, s0));
            ctx->stats.h2d_bytes += all16.size() * sizeof(Item16);
        }
        for (int g = (int)groups16Q.size() - 1; g >= 0; --g) {
            goff16Q[g] = all16q.size();
            all16q.insert(all16q.end(), groups16Q[g].items.begin(), groups16Q[g].items.end());
        }
        if (!all16q.empty()) {
            CK(ctx->items16q.ensure(all16q.size() * sizeof(Item16Q)));
            CK(cudaMemcpyAsync(ctx->items16q.p, all16q.data(), all16q.size() * sizeof(Item16Q), cudaMemcpyHostToDevice, s0));
            ctx->stats.h2d_bytes += all16q.size() * sizeof(Item16Q);
        }
        CK(cudaStreamSynchronize(s0));   // `all16` is a temporary
    }
    // every CTA of every concurrently running MULTI kernel owns its own boundary slice
    {
        uint64_t scr_total = 0;
        for (Launch& L : launches) {
            if (L.pair || !group_multi(L.g)) continue;
            uint32_t grid = 0;
            rc = grid_for(ctx, group_kernel(L.g), group_k(L.g), C, L.n_items, &grid);
            if (rc) return rc;
            L.scr_off = scr_total;
            scr_total += (uint64_t)grid * var[L.v].groups[L.g].stride;
        }
        if (scr_total) CK(ctx->scratch.ensure(scr_total * sizeof(uint2)));
    }
    GateGuard gate;
    gate.lock(ctx->gpu_gate);
    CK(cudaEventRecord(ctx->ev_start, s0));
    for (int i = 1; i < kStreams; ++i) CK(cudaStreamWaitEvent(ctx->streams[i], ctx->ev_start, 0));
    {
        // BSA_PROFILE_GROUPS=1: run the groups one after another and report each one's rate
        const bool prof_groups = getenv("BSA_PROFILE_GROUPS") != nullptr;
        int li = 0;
        for (const Launch& L : launches) {
            const Variant& V = var[L.v];
            const int K = L.pair ? L.g % kGroupStride : group_k(L.g);
            const bool tag = L.pair ? L.g >= kGroupStride : group_tag(L.g);
            cudaStream_t st = prof_groups ? s0 : ctx->streams[li % kStreams];
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (prof_groups) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
            double cells = 0, swept = 0;
            if (!L.pair) {
                const Group& G = V.groups[L.g];
                KArgs a;
                memset(&a, 0, sizeof(a));
                a.Q = V.qdev; a.T = V.tdev; a.QA = V.qadev; a.cs_cap = V.cs_cap;
                a.subst = ctx->d_subst.as<int16_t>(); a.isgap = ctx->d_isgap.as<uint8_t>();
                a.C = C; a.go = ctx->go; a.ge = ctx->ge; a.one = 1; a.one2 = 1;
                a.items = ctx->items.as<Item>() + L.item_off;
                a.n_items = L.n_items;
                a.item_counter = ctx->counters.as<uint32_t>() + li;
                a.scores = d_scores; a.nident = d_nid;
                a.scratch = ctx->scratch.as<uint2>() + L.scr_off;
                a.scratch_stride = (uint32_t)G.stride;
                a.out_lut = V.lut; a.flip = V.flip;
                rc = launch(ctx, group_kernel(L.g), K, a, st);
                if (rc) return rc;
                cells = G.cells; swept = G.swept;
            } else {
                const GroupPair& G = V.pairs[L.g];
                const KernelPairFn pfn = tag ? g_pair_tag[K] : g_pair[K];
                KArgsPair a;
                memset(&a, 0, sizeof(a));
                a.Q = V.qdev; a.T = V.tdev; a.QA = V.qadev;
                a.subst = ctx->d_subst.as<int16_t>(); a.isgap = ctx->d_isgap.as<uint8_t>();
                a.C = C; a.go = ctx->go; a.ge = ctx->ge; a.one = 1; a.one2 = 1; a.cs_cap = V.cs_cap;
                a.items = ctx->items_pair.as<Item16>() + L.item_off;
                a.n_items = L.n_items;
                a.item_counter = ctx->counters.as<uint32_t>() + li;
                a.scores = d_scores; a.nident = d_nid;
                a.out_lut = V.lut; a.flip = V.flip;
                const size_t smem = smem_for(K, C);
                uint32_t grid = 0;
                rc = grid_for(ctx, (KernelFn)pfn, K, C, a.n_items, &grid);
                if (rc) return rc;
                pfn<<<grid, kThreads, smem, st>>>(a);
                CK(cudaGetLastError());
                ctx->stats.launches++;
                cells = G.cells; swept = G.swept;
            }
            if (prof_groups) {
                cudaEventRecord(e1, st);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                fprintf(stderr, "[bsa %s %s] K=%2d multi=%d tag=%d items=%7u cells=%.4e swept/cells=%.3f ms=%9.3f GCUPS=%8.1f\n",
                        L.pair ? "pair " : "group", V.name, K, (!L.pair && group_multi(L.g)) ? 1 : 0, tag ? 1 : 0, L.n_items, cells,
                        cells > 0 ? swept / cells : 0.0, ms, ms > 0 ? cells / 1e6 / ms : 0.0);
                cudaEventDestroy(e0); cudaEventDestroy(e1);
            }
            ++li;
        }
    }
    // ---- 16-bit score-only quads: four short templates per warp ----
    {
        const bool prof_groups = getenv("BSA_PROFILE_GROUPS") != nullptr;
        int li = 0;
        for (int g = (int)groups16Q.size() - 1; g >= 1; --g) {
            Group16Q& G = groups16Q[g];
            if (G.items.empty()) continue;
            KArgs16Q a;
            memset(&a, 0, sizeof(a));
            a.Q = Q.dev(); a.T = T.dev(); a.QA = Q.adev();
            a.subst = ctx->d_subst.as<int16_t>();
            a.C = C; a.go = ctx->go; a.ge = ctx->ge;
            a.items = ctx->items16q.as<Item16Q>() + goff16Q[g];
            a.n_items = (uint32_t)G.items.size();
            a.item_counter = ctx->counters.as<uint32_t>() + kCounters16 + groups16.size() + g;
            a.scores = d_scores;
            const size_t smem = smem_for(g, C);
            uint32_t grid = 0;
            rc = grid_for(ctx, (KernelFn)g_score16_quad[g], g, C, a.n_items, &grid);
            if (rc) return rc;
            cudaStream_t st = prof_groups ? s0 : ctx->streams[(li + 3) % kStreams];
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (prof_groups) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
            g_score16_quad[g]<<<grid, kThreads, smem, st>>>(a);
            CK(cudaGetLastError());
            ctx->stats.launches++;
            ctx->stats.items += a.n_items;
            if (prof_groups) {
                cudaEventRecord(e1, st);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                fprintf(stderr, "[bsa quad16] K=%2d items=%7zu cells=%.4e ms=%9.3f GCUPS=%8.1f\n", g, G.items.size(), G.cells, ms,
                        ms > 0 ? G.cells / 1e6 / ms : 0.0);
                cudaEventDestroy(e0); cudaEventDestroy(e1);
            }
            ++li;
        }
    }
    // ---- 16-bit score-only groups (they follow on the same streams) ----
    {
        size_t n16 = 0;
        for (auto& g : groups16) n16 += g.items.size();
        if (n16) {
            std::vector<int> gorder;
            for (int g = (int)groups16.size() - 1; g >= 0; --g) if (!groups16[g].items.empty()) gorder.push_back(g);
            uint64_t scr_total = 0;
            for (int g : gorder) {
                if (g > kKMax) {
                    const int K = g - (kKMax + 1);
                    uint32_t grid = 0;
                    rc = grid_for(ctx, (KernelFn)g_score16_multi[K], K, C, (uint32_t)groups16[g].items.size(), &grid);
                    if (rc) return rc;
                    groups16[g].scr_off = scr_total;
                    scr_total += (uint64_t)grid * groups16[g].stride;
                }
            }
            ctx->stats.items += (uint32_t)n16;
            if (scr_total) CK(ctx->scratch16.ensure(scr_total * sizeof(uint2)));
            const bool prof_groups = getenv("BSA_PROFILE_GROUPS") != nullptr;
            int li = 0;
            for (int g : gorder) {
                const bool multi = g > kKMax;
                const int K = multi ? g - (kKMax + 1) : g;
                KArgs16 a;
                memset(&a, 0, sizeof(a));
                a.Q = Q.dev(); a.T = T.dev();
                a.QA = Q.adev();
                a.subst = ctx->d_subst.as<int16_t>();
                a.C = C; a.go = ctx->go; a.ge = ctx->ge; a.one = 1;
                a.items = ctx->items16.as<Item16>() + goff16[g];
                a.n_items = (uint32_t)groups16[g].items.size();
                a.item_counter = ctx->counters.as<uint32_t>() + kCounters16 + g;
                a.scores = d_scores;
                a.scratch = ctx->scratch16.as<uint2>() + groups16[g].scr_off;
                a.scratch_stride = (uint32_t)groups16[g].stride;
                Kernel16Fn fn = multi ? g_score16_multi[K] : g_score16_single[K];
                const size_t smem = smem_for(K, C);
                uint32_t grid = 0;
                rc = grid_for(ctx, (KernelFn)fn, K, C, a.n_items, &grid);
                if (rc) return rc;
                cudaStream_t st = prof_groups ? s0 : ctx->streams[li % kStreams];
                cudaEvent_t e0 = nullptr, e1 = nullptr;
                if (prof_groups) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
                fn<<<grid, kThreads, smem, st>>>(a);
                CK(cudaGetLastError());
                ctx->stats.launches++;
                if (prof_groups) {
                    cudaEventRecord(e1, st);
                    cudaEventSynchronize(e1);
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, e0, e1);
                    fprintf(stderr, "[bsa group16] K=%2d multi=%d items=%7zu cells=%.4e ms=%9.3f GCUPS=%8.1f\n", K, multi ? 1 : 0,
                            groups16[g].items.size(), groups16[g].cells, ms, ms > 0 ? groups16[g].cells / 1e6 / ms : 0.0);
                    cudaEventDestroy(e0); cudaEventDestroy(e1);
                }
                ++li;
            }
        }
    }
    // one join for everything that was launched on the side streams
    for (int i = 1; i < kStreams; ++i) {
        CK(cudaEventRecord(ctx->ev_s[i], ctx->streams[i]));
        CK(cudaStreamWaitEvent(s0, ctx->ev_s[i], 0));
    }
    // pairs whose score range does not fit the packed lanes: direction-store path
    if (!fallback.empty()) {
        rc = run_pairs_dirs(ctx, Q, T, fallback, d_scores, d_nid, nullptr, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    if (!fixes.empty() && (d_scores || d_nid)) {
        CK(ctx->fixes.ensure(fixes.size() * sizeof(Fix)));
        CK(cudaMemcpyAsync(ctx->fixes.p, fixes.data(), fixes.size() * sizeof(Fix), cudaMemcpyHostToDevice, s0));
        apply_fix_kernel<<<(uint32_t)((fixes.size() + 255) / 256), 256, 0, s0>>>(
            ctx->fixes.as<Fix>(), (uint32_t)fixes.size(), d_scores, d_nid);
        CK(cudaGetLastError());
        ctx->stats.launches++;
    }
    CK(cudaEventRecord(ctx->ev_end, s0));
    if (ctx->gpu_gate) {
        CK(cudaEventSynchronize(ctx->ev_end));      // the kernels are through: the GPU is the next tile's
        gate.unlock();
    }
    if (!out_dev && !mapped_out) {
        if (want_s) CK(cudaMemcpyAsync(scores, d_scores, n_res * 4, cudaMemcpyDeviceToHost, s0));
        if (want_i) CK(cudaMemcpyAsync(n_identical, d_nid, n_res * 4, cudaMemcpyDeviceToHost, s0));
    }
    if (!out_dev) ctx->stats.d2h_bytes += (want_s ? n_res * 4 : 0) + (want_i ? n_res * 4 : 0);
    CK(cudaStreamSynchronize(s0));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end));
    ctx->stats.kernel_ms = ms;
    ctx->stats.total_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return BSA_OK;
}

int bsa_all_vs_all(bsa_ctx* ctx, int set_id, uint32_t flags, int32_t* scores, uint32_t* n_identical) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (set_id < 0 || set_id >= kMaxSets || !ctx->sets[set_id].loaded)
        return fail(ctx, BSA_ERR_EMPTY, "sequence set not loaded");
    const uint32_t n = ctx->sets[set_id].n;
    std::vector<uint32_t> counts(n);
    for (uint32_t t = 0; t < n; ++t) counts[t] = t;   // q < t: alignment_protocols.rs:96-97
    return bsa_align_all_pairs(ctx, set_id, set_id, counts.data(), 0, n, flags, scores, n_identical, nullptr);
}

int bsa_one_vs_many(bsa_ctx* ctx, int q_set, int db_set, uint32_t flags, int32_t* scores,
                    uint32_t* n_identical) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (db_set < 0 || db_set >= kMaxSets || !ctx->sets[db_set].loaded)
        return fail(ctx, BSA_ERR_EMPTY, "sequence set not loaded");
    return bsa_align_all_pairs(ctx, q_set, db_set, nullptr, 0, ctx->sets[db_set].n, flags, scores,
                               n_identical, nullptr);
}

int bsa_align_pairs_paths(bsa_ctx* ctx, int q_set, int t_set, const uint32_t* q_idx, const uint32_t* t_idx,
                          uint64_t n_pairs, int32_t* scores, uint32_t* n_identical, uint8_t* path_buf,
                          uint64_t* path_off) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (ctx->multi)
        return multi_pair_list(ctx, q_set, t_set, q_idx, t_idx, n_pairs, path_buf, path_off,
                               [&](bsa_ctx* kid, uint64_t a, uint64_t cnt, uint8_t* pb, uint64_t* po) {
                                   return bsa_align_pairs_paths(kid, q_set, t_set, q_idx + a, t_idx + a, cnt,
                                                                scores ? scores + a : nullptr,
                                                                n_identical ? n_identical + a : nullptr, pb, po);
                               });
    const auto wall0 = std::chrono::steady_clock::now();
    if (q_set < 0 || q_set >= kMaxSets || t_set < 0 || t_set >= kMaxSets)
        return fail(ctx, BSA_ERR_BAD_ARG, "bad set id");
    const SeqSet &Q = ctx->sets[q_set], &T = ctx->sets[t_set];
    if (!Q.loaded || !T.loaded) return fail(ctx, BSA_ERR_EMPTY, "sequence set not loaded");
    if (n_pairs && (!q_idx || !t_idx)) return fail(ctx, BSA_ERR_BAD_ARG, "null pair list");
    if (path_buf && !path_off) return fail(ctx, BSA_ERR_BAD_ARG, "path_buf needs path_off");
    CK(cudaSetDevice(ctx->device));
    int rc = sync_scoring(ctx);
    if (rc) return rc;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->stats.h2d_bytes = ctx->pending_h2d;
    ctx->pending_h2d = 0;
    ctx->stats.pairs = n_pairs;
    if (path_off) path_off[0] = 0;
    if (n_pairs == 0) return BSA_OK;

    std::vector<PairReq> reqs;
    std::vector<uint64_t> req_index;          // request -> original pair index
    std::vector<uint64_t> slot_off(n_pairs + 1, 0);
    std::vector<uint32_t> plen(n_pairs, 0);
    std::vector<int32_t> h_scores(n_pairs, 0);
    std::vector<uint32_t> h_nid(n_pairs, 0);
    std::vector<uint8_t> degenerate(n_pairs, 0);
    double cells = 0;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        if (q_idx[p] >= Q.n || t_idx[p] >= T.n) return fail(ctx, BSA_ERR_BAD_ARG, "pair index out of range");
        const uint64_t n = Q.len(q_idx[p]), m = T.len(t_idx[p]);
        slot_off[p + 1] = slot_off[p] + n + m;
        cells += (double)n * (double)m;
        if (n == 0 || m == 0) {
            // only a border is walked: all '-' (row 0) or all '|' (column 0)
            degenerate[p] = 1;
            h_scores[p] = (n == 0 && m == 0) ? 0 : (int32_t)(ctx->go + (int64_t)(n + m - 1) * ctx->ge);
            plen[p] = (uint32_t)(n + m);
        } else {
            reqs.push_back(PairReq{q_idx[p], t_idx[p], p});
            req_index.push_back(p);
        }
    }
    ctx->stats.cells = (uint64_t)cells;
    CK(ctx->out_scores.ensure(n_pairs * 4));
    CK(ctx->out_nid.ensure(n_pairs * 4));
    cudaStream_t s0 = ctx->streams[0];
    CK(cudaMemsetAsync(ctx->out_scores.p, 0, n_pairs * 4, s0));
    CK(cudaMemsetAsync(ctx->out_nid.p, 0, n_pairs * 4, s0));
    CK(cudaEventRecord(ctx->ev_start, s0));
    rc = run_pairs_dirs(ctx, Q, T, reqs, ctx->out_scores.as<int32_t>(), ctx->out_nid.as<uint32_t>(), path_buf,
                        &slot_off, &plen, &req_index);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_end, s0));
    std::vector<int32_t> ds(n_pairs);
    std::vector<uint32_t> dn(n_pairs);
    CK(cudaMemcpyAsync(ds.data(), ctx->out_scores.p, n_pairs * 4, cudaMemcpyDeviceToHost, s0));
    CK(cudaMemcpyAsync(dn.data(), ctx->out_nid.p, n_pairs * 4, cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    ctx->stats.d2h_bytes += n_pairs * 8;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        if (!degenerate[p]) { h_scores[p] = ds[p]; h_nid[p] = dn[p]; }
        if (scores) scores[p] = h_scores[p];
        if (n_identical) n_identical[p] = h_nid[p];
    }
    if (path_buf) {
        // compact the right-aligned slots into back-to-back paths, in pair order
        uint64_t w = 0;
        for (uint64_t p = 0; p < n_pairs; ++p) {
            const uint64_t slot = slot_off[p + 1] - slot_off[p];
            if (degenerate[p]) {
                const uint64_t m = T.len(t_idx[p]);
                memset(path_buf + w, m ? '-' : '|', plen[p]);   // tmp space: w <= slot_off[p]
            } else {
                memmove(path_buf + w, path_buf + slot_off[p] + (slot - plen[p]), plen[p]);
            }
            w += plen[p];
            path_off[p + 1] = w;
        }
    } else if (path_off) {
        for (uint64_t p = 0; p < n_pairs; ++p) path_off[p + 1] = 0;
    }
    // device time of the fill and traceback kernels, batch by batch (run_pairs_dirs); total_ms is the wall clock of the call
    ctx->stats.kernel_ms = ctx->dirs_kernel_ms;
    ctx->stats.total_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return BSA_OK;
}

// LocalAlignment::align + backtrace + recent_end_point (bioshell-seq/src/alignment/local.rs:83-284)
int bsa_local_align_pairs(bsa_ctx* ctx, int q_set, int t_set, const uint32_t* q_idx, const uint32_t* t_idx,
                          uint64_t n_pairs, int32_t* scores, uint32_t* end_q, uint32_t* end_t, uint32_t* start_q,
                          uint32_t* start_t, uint8_t* path_buf, uint64_t* path_off) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (ctx->multi)
        return multi_pair_list(ctx, q_set, t_set, q_idx, t_idx, n_pairs, path_buf, path_off,
                               [&](bsa_ctx* kid, uint64_t a, uint64_t cnt, uint8_t* pb, uint64_t* po) {
                                   return bsa_local_align_pairs(kid, q_set, t_set, q_idx + a, t_idx + a, cnt,
                                                                scores ? scores + a : nullptr, end_q ? end_q + a : nullptr,
                                                                end_t ? end_t + a : nullptr, start_q ? start_q + a : nullptr,
                                                                start_t ? start_t + a : nullptr, pb, po);
                               });
    const auto wall0 = std::chrono::steady_clock::now();
    if (q_set < 0 || q_set >= kMaxSets || t_set < 0 || t_set >= kMaxSets)
        return fail(ctx, BSA_ERR_BAD_ARG, "bad set id");
    const SeqSet &Q = ctx->sets[q_set], &T = ctx->sets[t_set];
    if (!Q.loaded || !T.loaded) return fail(ctx, BSA_ERR_EMPTY, "sequence set not loaded");
    if (n_pairs && (!q_idx || !t_idx)) return fail(ctx, BSA_ERR_BAD_ARG, "null pair list");
    if (path_buf && !path_off) return fail(ctx, BSA_ERR_BAD_ARG, "path_buf needs path_off");
    CK(cudaSetDevice(ctx->device));
    int rc = sync_scoring(ctx);
    if (rc) return rc;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->stats.h2d_bytes = ctx->pending_h2d;
    ctx->pending_h2d = 0;
    ctx->stats.pairs = n_pairs;
    if (path_off) path_off[0] = 0;
    if (n_pairs == 0) return BSA_OK;
    std::vector<PairReq> reqs;
    std::vector<uint64_t> req_index;
    std::vector<uint64_t> slot_off(n_pairs + 1, 0);
    std::vector<uint32_t> plen(n_pairs, 0);
    double cells = 0;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        if (q_idx[p] >= Q.n || t_idx[p] >= T.n) return fail(ctx, BSA_ERR_BAD_ARG, "pair index out of range");
        const uint64_t n = Q.len(q_idx[p]), m = T.len(t_idx[p]);
        slot_off[p + 1] = slot_off[p] + n + m;
        cells += (double)n * (double)m;
        if (n && m) { reqs.push_back(PairReq{q_idx[p], t_idx[p], p}); req_index.push_back(p); }
    }
    ctx->stats.cells = (uint64_t)cells;
    DevBuf& lbuf = ctx->out_nid;   // reused as the LocalOut array of this call
    CK(lbuf.ensure(n_pairs * sizeof(LocalOut)));
    cudaStream_t s0 = ctx->streams[0];
    CK(cudaMemsetAsync(lbuf.p, 0, n_pairs * sizeof(LocalOut), s0));   // empty sequences: score 0, (0,0)
    CK(cudaEventRecord(ctx->ev_start, s0));
    rc = run_pairs_dirs(ctx, Q, T, reqs, nullptr, nullptr, path_buf, &slot_off, &plen, &req_index, lbuf.as<LocalOut>());
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_end, s0));
    std::vector<LocalOut> lo(n_pairs);
    CK(cudaMemcpyAsync(lo.data(), lbuf.p, n_pairs * sizeof(LocalOut), cudaMemcpyDeviceToHost, s0));
    CK(cudaStreamSynchronize(s0));
    ctx->stats.d2h_bytes += n_pairs * sizeof(LocalOut);
    for (uint64_t p = 0; p < n_pairs; ++p) {
        if (scores) scores[p] = lo[p].score;
        if (end_q) end_q[p] = lo[p].end_q;
        if (end_t) end_t[p] = lo[p].end_t;
        if (start_q) start_q[p] = lo[p].start_q;
        if (start_t) start_t[p] = lo[p].start_t;
    }
    if (path_buf) {
        uint64_t w = 0;
        for (uint64_t p = 0; p < n_pairs; ++p) {
            const uint64_t slot = slot_off[p + 1] - slot_off[p];
            if (plen[p]) memmove(path_buf + w, path_buf + slot_off[p] + (slot - plen[p]), plen[p]);
            w += plen[p];
            path_off[p + 1] = w;
        }
    } else if (path_off) {
        for (uint64_t p = 0; p < n_pairs; ++p) path_off[p + 1] = 0;
    }
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end));
    ctx->stats.kernel_ms = ms;
    ctx->stats.total_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return BSA_OK;
}

// hierarchical_clustering (bioshell-clustering/src/hierarchical/hierarchical.rs:22-80)
int bsa_hclust(bsa_ctx* ctx, uint32_t n, const float* dist, int linkage, uint32_t flags, uint32_t* mat_i,
               uint32_t* mat_j, float* merge_dist) {
    if (!ctx) return BSA_ERR_BAD_ARG;
    if (ctx->multi) {   // the clustering runs on the first device
        bsa_ctx* kid = ctx->multi->workers[0]->kid;
        const int rc = bsa_hclust(kid, n, dist, linkage, flags, mat_i, mat_j, merge_dist);
        ctx->err = kid->err;
        ctx->stats = kid->stats;
        return rc;
    }
    if (n == 0 || !dist || linkage < 0 || linkage > 5) return fail(ctx, BSA_ERR_BAD_ARG, "bad argument");
    const auto wall0 = std::chrono::steady_clock::now();
    CK(cudaSetDevice(ctx->device));
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    cudaStream_t st = ctx->streams[0];
    const size_t nn = (size_t)n * n;
    const uint32_t grid = (uint32_t)ctx->sms * 8;
    const uint32_t pitch = (n + 31u) & ~31u;
    CK(ctx->hc_matrix.ensure((size_t)n * pitch * sizeof(float)));
    // aux: rmap[n] sizes[n] mat_i[n] mat_j[n] mdist[n] result[n] rmin_v[n] rmin_j[n] todo[n] scalars[8] partial[grid]
    const size_t aux_bytes = (size_t)n * 4 * 9 + 32 + (size_t)grid * sizeof(HcBest) + 64;
    CK(ctx->hc_aux.ensure(aux_bytes));
    HcState hs;
    hs.D = ctx->hc_matrix.as<float>();
    uint8_t* a = ctx->hc_aux.as<uint8_t>();
    hs.rmap = (uint32_t*)a; a += (size_t)n * 4;
    hs.sizes = (uint32_t*)a; a += (size_t)n * 4;
    hs.mat_i = (uint32_t*)a; a += (size_t)n * 4;
    hs.mat_j = (uint32_t*)a; a += (size_t)n * 4;
    hs.mdist = (float*)a; a += (size_t)n * 4;
    hs.result = (float*)a; a += (size_t)n * 4;
    // nearest-neighbour cache (hclust_kernels.cuh): on unless BSA_HC_NN=0 asks for the full scan per merge
    const bool use_nn = !(getenv("BSA_HC_NN") && atoi(getenv("BSA_HC_NN")) == 0);
    hs.rmin_v = use_nn ? (float*)a : nullptr; a += (size_t)n * 4;
    hs.rmin_j = use_nn ? (uint32_t*)a : nullptr; a += (size_t)n * 4;
    hs.todo = (uint32_t*)a; a += (size_t)n * 4;
    hs.order = (uint32_t*)a; a += 32;
    uint32_t* done_counter = hs.order + 2;
    hs.todo_n = hs.order + 3;
    hs.partial = (HcBest*)a;
    hs.n = n;
    hs.pitch = pitch;
    hs.rule = linkage;
    const float* d_in = dist;
    if (!(flags & BSA_IN_DEVICE)) {
        CK(ctx->raw.ensure(nn * sizeof(float)));
        CK(cudaMemcpyAsync(ctx->raw.p, dist, nn * sizeof(float), cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += nn * sizeof(float);
        d_in = ctx->raw.as<float>();
    }
    CK(cudaEventRecord(ctx->ev_start, st));
    hclust_init_kernel<<<grid, 256, 0, st>>>(hs, d_in);
    CK(cudaGetLastError());
    ctx->stats.launches++;
    if (use_nn) {
        hclust_rowmin_kernel<<<std::min<uint32_t>(n, (uint32_t)ctx->sms * 2), 1024, 0, st>>>(hs, 1);
        CK(cudaGetLastError());
        ctx->stats.launches++;
    }
    const uint32_t nn_grid = std::max(1u, std::min(32u, (n + 2047u) / 2048u));     // cache entries: 12 bytes per row
    const uint32_t redo_grid = 16;                                                 // rows i, j and the odd queued row
    uint32_t merge_grid = std::max(1u, std::min(24u, (n + 255u) / 256u));
    if (const char* e = getenv("BSA_HC_MERGE_GRID")) merge_grid = (uint32_t)std::max(1, atoi(e));   // debugging aid
    // Every merge is the same two launches with the same arguments (the state lives on the device),
    // and late merges are launch-latency bound: replay them from CUDA graphs of kHcGraphMerges
    // merges.  The scan grid shrinks with the live matrix (one warp per ~2 rows, the full grid
    // from 4096 rows up), in power-of-two levels so that a handful of graphs covers the whole run.
    constexpr uint32_t kHcGraphMerges = 64;
    constexpr int kHcLevels = 7;
    auto level_for = [&](uint32_t order) {
        const uint64_t want = std::max<uint64_t>(16, (uint64_t)order * grid / 4096);
        int l = 0;
        while (l + 1 < kHcLevels && (grid >> (l + 1)) >= want) ++l;
        return l;
    };
    const bool use_graph = !getenv("BSA_HC_NO_GRAPH");
    cudaGraphExec_t execs[kHcLevels] = {nullptr};
    cudaError_t ge = cudaSuccess;
    const uint32_t n_merges = n - 1;
    uint32_t step = 0;
    while (step < n_merges && ge == cudaSuccess) {
        const uint32_t order = n - step;                 // rows alive before this merge
        const int l = use_nn ? 0 : level_for(order);          // the cache's grids do not depend on the order
        const uint32_t g = std::max(1u, grid >> l);
        if (use_graph && step + kHcGraphMerges <= n_merges) {
            if (!execs[l]) {
                cudaGraph_t graph = nullptr;
                ge = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
                if (ge != cudaSuccess) break;
                for (uint32_t k = 0; k < kHcGraphMerges; ++k) {
                    if (use_nn) {
                        hclust_nn_argmin_kernel<<<nn_grid, 256, 0, st>>>(hs);
                        hclust_merge_kernel<<<merge_grid, 256, 0, st>>>(hs, nn_grid, done_counter);
                        hclust_rowmin_kernel<<<redo_grid, 1024, 0, st>>>(hs, 0);
                    } else {
                        hclust_argmin_kernel<<<g, 256, 0, st>>>(hs);
                        hclust_merge_kernel<<<merge_grid, 256, 0, st>>>(hs, g, done_counter);
                    }
                }
                ge = cudaStreamEndCapture(st, &graph);
                if (ge != cudaSuccess) break;
                ge = cudaGraphInstantiate(&execs[l], graph, 0);
                cudaGraphDestroy(graph);
                if (ge != cudaSuccess) break;
            }
            ge = cudaGraphLaunch(execs[l], st);
            ctx->stats.launches += (use_nn ? 3 : 2) * kHcGraphMerges;
            step += kHcGraphMerges;
        } else if (use_nn) {
            hclust_nn_argmin_kernel<<<nn_grid, 256, 0, st>>>(hs);
            hclust_merge_kernel<<<merge_grid, 256, 0, st>>>(hs, nn_grid, done_counter);
            hclust_rowmin_kernel<<<redo_grid, 1024, 0, st>>>(hs, 0);
            ctx->stats.launches += 3;
            ++step;
        } else {
            hclust_argmin_kernel<<<g, 256, 0, st>>>(hs);
            hclust_merge_kernel<<<merge_grid, 256, 0, st>>>(hs, g, done_counter);
            ctx->stats.launches += 2;
            ++step;
        }
    }
    for (int l = 0; l < kHcLevels; ++l) if (execs[l]) cudaGraphExecDestroy(execs[l]);
    if (ge != cudaSuccess) return fail_cuda(ctx, ge, "hclust graph replay");
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_end, st));
    uint32_t fin[2] = {0, 0};
    CK(cudaMemcpyAsync(fin, hs.order, 8, cudaMemcpyDeviceToHost, st));
    if (n > 1) {
        if (mat_i) CK(cudaMemcpyAsync(mat_i, hs.mat_i, (size_t)(n - 1) * 4, cudaMemcpyDeviceToHost, st));
        if (mat_j) CK(cudaMemcpyAsync(mat_j, hs.mat_j, (size_t)(n - 1) * 4, cudaMemcpyDeviceToHost, st));
        if (merge_dist) CK(cudaMemcpyAsync(merge_dist, hs.mdist, (size_t)(n - 1) * 4, cudaMemcpyDeviceToHost, st));
        ctx->stats.d2h_bytes += (size_t)(n - 1) * 12;
    }
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end));
    ctx->stats.kernel_ms = ms;
    ctx->stats.pairs = n > 1 ? n - 1 : 0;
    ctx->stats.total_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    if (fin[1] & 0x80000000u)
        return fail(ctx, BSA_ERR_RANGE, "no finite distance left to merge (the reference panics here)");
    return BSA_OK;
}

void* bsa_host_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void bsa_host_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

int bsa_get_stats(const bsa_ctx* ctx, bsa_stats* out) {
    if (!ctx || !out) return BSA_ERR_BAD_ARG;
    *out = ctx->stats;
    return BSA_OK;
}

int bsa_measure_int_peak(bsa_ctx* ctx, int which, double* lane_ops_per_s, double* sm_clock_mhz) {
    if (!ctx || !lane_ops_per_s) return BSA_ERR_BAD_ARG;
    if (ctx->multi) ctx = ctx->multi->workers[0]->kid;
    CK(cudaSetDevice(ctx->device));
    double ops = 0, mhz = 0;
    cudaError_t e = bsa::measure_int_peak(which, ctx->sms, ctx->streams[0], &ops, &mhz);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "int-peak microbenchmark");
    *lane_ops_per_s = ops;
    if (sm_clock_mhz) *sm_clock_mhz = mhz;
    return BSA_OK;
}

}  // extern "C"
