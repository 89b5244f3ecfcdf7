// gotoh_kernels.cuh -- hand-written sm_100a kernels for Needleman-Wunsch/Gotoh global
// alignment with the BioShell reference's exact tie-breaking.
//
// Replaces (reference paths relative to the BioShell checkout):
//   GlobalAligner::align      bioshell-seq/src/alignment/global.rs:57-145
//   GlobalAligner::backtrace  bioshell-seq/src/alignment/global.rs:146-201
//   identity count            bioshell-seq/src/msa/msa.rs:261-269
//
// Design (DESIGN.md has the long version):
//  * TEMPLATE-STATIONARY, QUERY-STREAMING systolic warp.  A CTA owns one template
//    (the DP columns).  Lane l of a warp keeps K consecutive columns in registers;
//    the residues of MANY queries are streamed through the lanes back to back
//    (lane l works on stream position s-l at step s), so the pipeline fill/drain
//    of the 32-lane wavefront is paid once per work item, not once per pair.
//    Column boundaries (H and the eagerly-computed E) hop lanes by warp shuffle.
//  * ONE 32-bit lane carries (score, tie-priority, identical-count) packed as
//        v = score << (cs+2) | prio << cs | count          (cs = count bits)
//    so that a signed integer max IS the reference's selection rule:
//    higher score first, then the reference's priority among equal scores
//    (H: diagonal 3 > E 2 > F 1, global.rs:161-169; E/F: extend beats open,
//    global.rs:109,122), and the count of identical residues of the winning
//    path rides along for free.  This CLASSIC cell (templates whose scores leave
//    no room for a tag field, and the direction-store kernels) is 8 integer
//    instructions; the score + identity kernels run the TAG cell in the moving
//    frame score - (i + j) ge instead -- 5 instructions, see kTagBits below and
//    stream_block_tag2a (two rows per step over the even-aligned stream):
//        e  = eraw | PH                      LOP3
//        f  = fraw[c] | PV                   LOP3
//        d  = hdiag + T[q][t]                IADD3      (T carries prio 3 and the identity bit)
//        h  = max3(d, e, f)                  VIMNMX3    (DPX)
//        hc = h & ~prio                      LOP3
//        hg = hc + GO                        IADD3
//        eraw    = max(e + GE, hg)           VIADDMNMX  (DPX)
//        fraw[c] = max(f + GE, hg)           VIADDMNMX  (DPX)
//    The reference's third E/F term (e_from_f / f_from_e) and its capacity
//    dependent sentinel never win for gap_open <= gap_extend <= 0 (SURVEY.md 8a
//    note 1), so they are dropped; the borders are produced eagerly instead.
//  * The substitution scores come from a per-template PROFILE in shared memory,
//    laid out [code][vec][lane] as uint4 so that every lane reads its own 16-byte
//    bank group: conflict-free LDS.128 whatever residue each lane is on.
//  * DIRS variant: the same recurrences with cs = 0; the priority bits that fall
//    out of the max ARE the traceback directions, so each cell also emits a 4-bit
//    code (2-bit H source + E-extended + F-extended) to HBM, written coalesced in
//    step-major order, and a second kernel walks them (global.rs:146-201).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// tuning switches (tools/ab_build.sh builds variants of the library for same-box A/B runs)
#ifndef BSA_TAG2_U
#define BSA_TAG2_U 4        // double steps per loop iteration of the two-row blocks (same-box A/B on cfg2: 2 -> 3195, 4 -> 3284, 8 -> 3024, 1 -> 3026 GCUPS)
#endif
#ifndef BSA_TAG2_U_SMALLK
#define BSA_TAG2_U_SMALLK 0  // aligned two-row blocks with at most this many columns per lane take 8 double steps per iteration (measured with 6 and 10: no gain on cfg2 / one-vs-many, -4..-8 % on 1,000 sequences: off)
#endif
#ifndef BSA_TAIL16
#define BSA_TAIL16 2        // sixteenths of an item's stream that are cut into small chunks, the warps' run-in to the closing barrier (same-box A/B on cfg2: 4 -> 3,425, 3 -> 3,437, 2 -> 3,453, 1 -> 3,451 GCUPS)
#endif
#ifndef BSA_RING
#define BSA_RING 0          // warp-wide prefetch (measured -0.4 % on cfg2: off) of the boundary column in multi-pass kernels
#endif
#ifndef BSA_WAVE_RING
#define BSA_WAVE_RING 1     // K3: warp-wide batches of 32 boundary entries instead of lane 0's per-step load (one L2 round trip per row); 0: A/B only
#endif
#ifndef BSA_WAVE_P16
#define BSA_WAVE_P16 0      // K3: 16-bit profile entries (half the shared memory per warp -> 16 warps per SM): bit-exact, measured slower (28.0 vs 24.9 ms on cfg5), off
#endif
#ifndef BSA_WAVE_BATCH
#define BSA_WAVE_BATCH 32   // K3: boundary entries per publication / per warp-wide fetch (32 or 16; 16 halves the hand-off lag, doubles the release fences): NOT yet measured
#endif
#ifndef BSA_WAVE_PF
#define BSA_WAVE_PF 8       // K3: step inside a batch at which the next batch is fetched (even, < BSA_WAVE_BATCH); later = less lag: NOT yet measured
#endif
#ifndef BSA_WAVE_POLL
#define BSA_WAVE_POLL 0     // K3: poll the hand-off counter with relaxed loads (+ nanosleep) and acquire once, instead of an acquire (= L1 invalidate) per poll
#endif
#ifndef BSA_WAVE_PD
#define BSA_WAVE_PD 1       // K3: residue groups fetched ahead (1 = the next group only)
#endif
#ifndef BSA_TMA
#define BSA_TMA 1           // 0: A/B only -- the elected thread copies with plain loads instead of cp.async.bulk
#endif
#ifndef BSA_MB3_MAXK
#define BSA_MB3_MAXK 12     // largest K that still runs three CTAs per SM (80 registers per thread)
#endif
#ifndef BSA_CHUNK_BIG
#define BSA_CHUNK_BIG 4096  // stream residues per big chunk
#endif
#ifndef BSA_FLAG_LOOP
#define BSA_FLAG_LOOP 0      // two-row blocks: the flagged branch as a rolled loop over its double steps (half the code of the slow path): NOT yet measured
#endif
#if BSA_FLAG_LOOP
#define BSA_FLAG_UNROLL _Pragma("unroll 1")
#define BSA_B(u, k) b[k]
#else
#define BSA_FLAG_UNROLL _Pragma("unroll")
#define BSA_B(u, k) b[2 * (u) + (k)]
#endif
#ifndef BSA_FLAG_SPLIT
#define BSA_FLAG_SPLIT 0     // two-row blocks: a flagged double step whose flags all sit on its second row keeps the interleaved form
#endif
#ifndef BSA_ALIGNED
#define BSA_ALIGNED 1       // two-row blocks read the EVEN-ALIGNED copy of the stream store (a PAD row ahead of odd-length
                            // sequences): end-of-sequence flags only on the second row of a double step, one code path
#endif
#ifndef BSA_ETAG
#define BSA_ETAG 1          // aligned two-row blocks: E openings tagged by column (no addition in the E extension), the
                            // count-field width a launch constant so that every cell constant is warp-uniform
#endif
#ifndef BSA_PADSLOT
#define BSA_PADSLOT 1       // aligned two-row blocks, K % 4 != 0: lane 0's left-border value rides in the spare slot of the profile row
#endif
#ifndef BSA_TWO_ROWS
#define BSA_TWO_ROWS 1      // two-row step where TwoRows<K, HALF> says so
#endif

namespace bsa {

#ifndef BSA_WARPS_PER_CTA
#define BSA_WARPS_PER_CTA 8
#endif
constexpr int kWarpsPerCta = BSA_WARPS_PER_CTA;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr uint32_t kLastFlag = 0x80u;  // set on the last residue of every sequence in the store
constexpr uint32_t kCodeMask = 0x7Fu;
constexpr uint32_t kPadCode = 0u;      // residue code 0 is never a residue: the PAD row of the even-aligned stores
constexpr int kFrontPad = 64;          // bytes before the first sequence (last one flagged)
constexpr int kBackPad = 128;          // bytes after the last sequence

// One unit of work: template t against the contiguous query range [q_begin, q_end).
struct Item {
    uint32_t t;
    uint32_t q_begin;
    uint32_t q_end;
    uint32_t cshift;    // count-field width for this item (0 in DIRS mode)
    uint64_t out_base;  // result index of pair (q_begin, t)
};

// DIRS mode: one pair per entry; pairs of one item share the template.
struct PairRec {
    uint32_t q, t;
    uint64_t out;       // result index (scores / n_identical / path slot)
    uint64_t dir_off;   // word offset of this pair's direction planes
    uint64_t scr_off;   // entry offset of this pair's boundary column (multi-pass templates)
    uint64_t path_off;  // byte offset of this pair's path slot (len_q + len_t bytes)
    uint32_t k;         // columns per lane the fill kernel used
    uint32_t prog_off;  // wavefront mode: first progress counter of this pair (one per column block)
};

struct SeqStoreDev {
    const uint8_t* codes;   // points at the first residue of sequence 0 (pads around it)
    const uint64_t* off;    // n+1 offsets
    uint32_t n;
};

struct KArgs {
    SeqStoreDev Q, T;
    SeqStoreDev QA;         // even-aligned copy of the stream store Q (two-row TAG blocks; see stream_block_tag2a)
    int cs_cap;             // bitlen(longest stream sequence): cap of the count-field width (aligned two-row blocks)
    const int16_t* subst;   // C x C substitution scores by residue code
    const uint8_t* isgap;   // C flags: code is '-' or '_' (never identical, msa.rs:264)
    int C;
    int go, ge;
    int one;                // the constant 1, kept opaque to the compiler (see cell_row)
    int one2;               // a second one (TAG cell: keeps the two openings apart)
    int mone;               // -1, as opaque as `one` (K3: a subtraction on the FMA pipe)
    const Item* items;
    uint32_t n_items;
    uint32_t* item_counter;
    int32_t* scores;        // may be null
    uint32_t* nident;       // may be null
    uint2* scratch;         // MULTI: per-warp boundary columns
    uint32_t scratch_stride;  // entries per warp
    const PairRec* pairs;   // DIRS
    uint32_t* dirs;         // DIRS
    uint32_t* progress;     // WAVE: boundary hand-off counters
    const uint2* wave_items;  // WAVE: (pair index, column block), in dependency order
    // Owner swap (bsa_api.cu, "hybrid plan"): the stream set may be a derived store (the long
    // sequences only).  out_lut[i] is then what stream sequence i contributes to the result index,
    // k = item.out_base + out_lut[i]; flip = 1 when the ROWS are the reference's template and the
    // columns its query: the H-max priorities of E and F trade places (global.rs:161-169 seen
    // from the transposed matrix) and the substitution lookup is transposed.
    const uint64_t* out_lut;
    int flip;
    uint4* wave_bnd;        // WAVE (two-row blocks): boundary columns as self-validating entries {H, epoch, E, epoch}
    uint32_t epoch;         // WAVE: this launch's tag (nonzero, never reused on this buffer)
    unsigned long long* wave_trace;   // WAVE, diagnostics (BSA_WAVE_TRACE): per item {start ns, end ns, SM id, steps}, else null
};

__device__ __forceinline__ int max3_s32(int a, int b, int c) { return __vimax3_s32(a, b, c); }
// one residue code of the packed store, zero-extended into a full register (kept opaque so the
// compiler does not start juggling 8/16-bit sub-registers in the hot loop)
__device__ __forceinline__ uint32_t ld_code(const uint8_t* p) {
    uint32_t v;
    asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int addmax_s32(int a, int b, int c) { return __viaddmax_s32(a, b, c); }

struct Consts {
    int GE, GO, MASK, PH, PV, T_PAD;
    int GEB;  // border step gap_extend << sh (0 in TAG mode: the frame makes the borders constant)
    int GOE, GOF;  // TAG: (gap_open - gap_extend) with the E / F priority already in place (== GO otherwise)
    int XCLR;      // TAG: clears the streak field
    int X1;        // TAG: one unit of the streak field
    int XTOP;      // TAG: top bit of the streak field ("older than every opening of this block of rows")
    int TSUB;      // TAG: what the frame adds to every substitution score, in score units (-2 gap_extend)
    int ge;        // gap_extend (TAG: the emitted score is H* + (n + m) gap_extend)
    int hb0;  // packed H[1][0] = gap_open   (TAG: H*[i][0] = gap_open - gap_extend for every i >= 1)
    int cs;   // width of the count field
    int ps;   // position of the 2-bit priority field (cs, or cs + kTagBits in TAG mode)
    int sh;   // position of the score field (ps + 2)
    bool local;   // Smith-Waterman borders: everything starts at 0
};

// TAG cell (score + identity kernels): width of the tag field x that sits between the count
// and the priority,   v = score << (cs+7) | prio << (cs+5) | x << cs | count.
// E and F values are born with their H-max priority (2 / 1) already in place, so the two
// `| PH`, `| PV` of the classic cell disappear.
// MOVING FRAME (round 2): the score field of DP cell (i, j) holds  score - (i + j) * gap_extend.
// In that frame BOTH gap extensions are free (E*[j+1] = max(E*[j], H* + go - ge), the same for F
// down the rows), the diagonal adds (s - 2 ge) -- folded into the profile -- and the borders are
// the constant go - ge.  "Extend beats open on ties" (global.rs:109,122) comes from x:
//   E: an extension adds 1 to x and nothing else (the addend of its VIADDMNMX), an opening has
//      x = 0; x is cleared when E enters the next lane (<= K <= 20 increments);
//   F: no addition at all -- the OPENING of step s carries x = 15 - (s mod 16), so an older
//      opening outranks a newer one, and every kTagRows steps the stored F get the top bit of x
//      (an OR), which outranks every opening of the next 16 steps.
// The cell is 6 instructions: IMAD (diagonal), VIMNMX3, LOP3, IMAD (E opening), 2 VIADDMNMX.
// In the aligned two-row blocks (stream_block_tag2a, BSA_ETAG) E uses F's scheme along the columns of a
// lane instead -- the opening of column c carries x = K-1-c, an E that crosses a lane boundary gets the
// whole field -- so the E update is ONE VIADDMNMX with a warp-uniform per-column constant and the cell is
// 5 instructions: IMAD (diagonal), VIMNMX3, LOP3, 2 VIADDMNMX.
// tests/packed_model.py::frame_align is the scalar model of exactly this arithmetic (both E schemes).
constexpr int kTagBits = 5;
constexpr uint32_t kTagRows = 16;

template <int K>
struct KTraits {
    static constexpr int V = (K + 3) / 4;        // uint4 vectors per lane per profile row
    static constexpr int W = (K + 7) / 8;        // direction words per lane per step
    static constexpr int ROW = V * 32;           // uint4 per profile row
};

// Shared-memory layout (uint4 units): profile rows [C][V][32], then rsH [V][32], rsF [V][32], then
// the TMA staging area: the substitution table (C x C int16, rounded up to 16 B) and two template
// slices of 32 K + 32 bytes each (the 16-byte aligned superset of the columns of one block).
template <int K>
__host__ __device__ constexpr size_t smem_bytes(int C) {
    return (size_t)(C + 2) * KTraits<K>::ROW * sizeof(uint4);
}
__host__ __device__ constexpr uint32_t subst_stage_bytes(int C) { return ((uint32_t)(C * C * 2) + 15u) & ~15u; }
__host__ __device__ constexpr uint32_t slice_stage_bytes(int K) { return 32u * (uint32_t)K + 64u; }   // two of them also hold four 16 K + 32 byte slices (quad kernel)
__host__ __device__ constexpr size_t stage_bytes(int K, int C) {
    return (size_t)subst_stage_bytes(C) + 2u * slice_stage_bytes(K);
}

// ---------------------------------------------------------------------------------------------
// TMA staging.  The substitution table and the template residues a CTA builds its profile from
// reach shared memory through the bulk-copy engine (cp.async.bulk ... mbarrier::complete_tx,
// UBLKCP in SASS): one elected thread arms the mbarrier with the byte count and issues the copies,
// every thread waits on the barrier's phase.  Sources are the 16-byte aligned supersets of the
// wanted ranges (the packed store has 64/128 bytes of padding around it, the table buffer is
// over-allocated), sizes are multiples of 16.
// The barrier's phase bit lives in shared memory (flipped by the elected thread after the
// barrier that follows each use), so nothing of this occupies a register inside the DP loop.
struct TmaStage {
    static __device__ __forceinline__ uint32_t smem_u32(const void* p) {
        return (uint32_t)__cvta_generic_to_shared(p);
    }
    static __device__ __forceinline__ void init(uint64_t* bar) {   // elected thread
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // elected thread: the generic-proxy reads of the staging area that came before (ordered by the
    // caller's __syncthreads) must be visible to the async proxy before it overwrites the area
    static __device__ __forceinline__ void arm(uint64_t* bar, uint32_t bytes) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
    static __device__ __forceinline__ void copy(uint64_t* bar, void* dst_smem, const void* src, uint32_t bytes) {
#if !BSA_TMA
        for (uint32_t i = 0; i < bytes; i += 16)
            *reinterpret_cast<uint4*>(reinterpret_cast<char*>(dst_smem) + i) =
                *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(src) + i);
        asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        return;
#endif
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    }
    static __device__ __forceinline__ void wait(uint64_t* bar, uint32_t phase) {   // every thread
        uint32_t done;
        do {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        } while (!done);
    }
};

// Kernel preamble: set up the mbarrier and bring the substitution table in (once per CTA; it does
// not change between items).  BSA_TMA_PTRS carves the staging area behind the profile.
#define BSA_TMA_PTRS(KK, CC)                                                                      \
    uint8_t* stage = reinterpret_cast<uint8_t*>(rsF + KTraits<KK>::ROW);                          \
    const int16_t* s_subst = reinterpret_cast<const int16_t*>(stage);                             \
    uint8_t* s_tcA = stage + subst_stage_bytes(CC);                                               \
    uint8_t* s_tcB = s_tcA + slice_stage_bytes(KK);                                               \
    (void)s_subst; (void)s_tcB;
#define BSA_TMA_PREAMBLE(KK, CC, SUBST)                                                           \
    __shared__ uint64_t s_mbar;                                                                   \
    __shared__ uint32_t s_phase;                                                                  \
    if (threadIdx.x == 0) {                                                                       \
        TmaStage::init(&s_mbar);                                                                  \
        TmaStage::arm(&s_mbar, subst_stage_bytes(CC));                                            \
        TmaStage::copy(&s_mbar, rsF + KTraits<KK>::ROW, SUBST, subst_stage_bytes(CC));            \
        s_phase = 1;                                                                              \
    }                                                                                             \
    __syncthreads();                                                                              \
    TmaStage::wait(&s_mbar, 0);
// after the __syncthreads that follows a use: the elected thread flips the phase for the next one
#define BSA_TMA_FLIP() if (threadIdx.x == 0) s_phase ^= 1u;

// Stage the columns [colbase, colbase + ncols) of the template that starts at `tc` into `dst`;
// returns the pointer p with p[col] == tc[col] for those columns.  Called by the elected thread
// between arm() and wait(); `bytes_out` is what it asked the copy engine for.
__device__ __forceinline__ uint32_t slice_bytes(const uint8_t* tc, uint32_t colbase, uint32_t ncols) {
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(tc + colbase) & 15u);
    return (head + ncols + 15u) & ~15u;
}
__device__ __forceinline__ const uint8_t* slice_view(const uint8_t* dst, const uint8_t* tc, uint32_t colbase) {
    const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(tc + colbase) & 15u);
    return dst + head - colbase;
}
__device__ __forceinline__ const uint8_t* slice_src(const uint8_t* tc, uint32_t colbase) {
    return reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(tc + colbase) & ~(uintptr_t)15u);
}

// Build the profile of template columns [colbase, colbase + 32K) and the per-lane
// top-border reset vectors.  All threads of the CTA participate.
// `subst` and `tc` point into the TMA staging area (tc[col] valid for this block's columns).
template <int K>
__device__ __forceinline__ void build_profile(uint4* prof, uint4* rsH, uint4* rsF,
                                              const int16_t* subst, const uint8_t* tc, uint32_t m,
                                              uint32_t colbase, const KArgs& a, const Consts& cs) {
    constexpr int ROW = KTraits<K>::ROW;
    const int C = a.C;
    const int S = 1 << cs.sh;
    const int P3 = 3 << cs.ps;
    for (int idx = threadIdx.x; idx < C * ROW; idx += blockDim.x) {
        const int code = idx / ROW;
        const int r = idx - code * ROW;
        const int v = r >> 5, lane = r & 31;
        const bool gap = a.isgap[code] != 0;
        int o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * v + e;
            const uint32_t col = colbase + lane * K + c;
            int val = cs.T_PAD;
            // spare slot behind the K columns (load_vec_x): lane 0's left-border H* of a row with this residue
            if (BSA_PADSLOT && K % 4 != 0 && c == K && cs.XTOP && colbase == 0)
                val = lane == 0 ? (code == (int)kPadCode ? 0 : cs.hb0) : 0;
            if (c < K && col < m) {
                const int tcode = tc[col] & kCodeMask;
                val = ((int)subst[a.flip ? tcode * C + code : code * C + tcode] + cs.TSUB) * S + P3 +
                      ((cs.cs > 0 && tcode == code && !gap) ? 1 : 0);   // no count field when cs == 0
                // frame cell: the PAD row adds nothing on the diagonal of the frame (go - ge in DP column 1,
                // whose diagonal input is H*[0][0] = 0), so it reproduces the top border
                if (cs.XTOP && code == (int)kPadCode) val = (col == 0 ? cs.hb0 : 0) + P3;
            }
            o[e] = val;
        }
        prof[idx] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    // top border: H[0][j] = go + (j-1) ge  (global.rs:81-88); eager F for row 1 =
    // H[0][j] + go (opened, never extended from the sentinel: global.rs:118-128).
    // TAG (moving frame): both are constants, H*[0][j] = go - ge and F*[1][j] = 2 (go - ge), the
    // latter with the top tag bit (it is older than every opening that follows)
    for (int r = threadIdx.x; r < ROW; r += blockDim.x) {
        const int v = r >> 5, lane = r & 31;
        int h[4], f[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * v + e;
            const long long j = (long long)colbase + lane * K + c + 1;  // DP column
            const int hv = cs.local ? 0 : (cs.XTOP ? cs.hb0 : (int)((a.go + (j - 1) * a.ge) * S));
            h[e] = hv;
            f[e] = cs.local ? 0 : hv + cs.GOF + cs.XTOP;
        }
        rsH[r] = make_uint4(h[0], h[1], h[2], h[3]);
        rsF[r] = make_uint4(f[0], f[1], f[2], f[3]);
    }
}

template <int K>
__device__ __forceinline__ void load_vec(int (&dst)[K], const uint4* __restrict__ src) {
    constexpr int V = KTraits<K>::V;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint4 x = src[v * 32];
        if (4 * v + 0 < K) dst[4 * v + 0] = (int)x.x;
        if (4 * v + 1 < K) dst[4 * v + 1] = (int)x.y;
        if (4 * v + 2 < K) dst[4 * v + 2] = (int)x.z;
        if (4 * v + 3 < K) dst[4 * v + 3] = (int)x.w;
    }
}

// The same row plus the entry behind its K columns (K % 4 != 0: the last vector has a spare slot).  The
// aligned two-row blocks keep lane 0's left-border value of the row there -- go - ge, or H*[0][0] = 0 in the
// PAD row, 0 in every other lane -- so the PAD test costs no instruction (BSA_PADSLOT).
template <int K>
__device__ __forceinline__ void load_vec_x(int (&dst)[K], int& extra, const uint4* __restrict__ src) {
    constexpr int V = KTraits<K>::V;
    static_assert(K % 4 != 0, "no spare slot");
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint4 x = src[v * 32];
        const int w[4] = {(int)x.x, (int)x.y, (int)x.z, (int)x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (4 * v + e < K) dst[4 * v + e] = w[e];
            if (4 * v + e == K) extra = w[e];
        }
    }
}

// The same load, opaque to the compiler: the border vectors a lane reloads when it passes the end of a
// sequence are loop invariant, and hoisting them would pin 2 K registers for the whole DP loop.
template <int K>
__device__ __forceinline__ void load_vec_opaque(int (&dst)[K], const uint4* src) {
    constexpr int V = KTraits<K>::V;
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(src);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        int x, y, z, w;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr + (uint32_t)v * 512u));
        if (4 * v + 0 < K) dst[4 * v + 0] = x;
        if (4 * v + 1 < K) dst[4 * v + 1] = y;
        if (4 * v + 2 < K) dst[4 * v + 2] = z;
        if (4 * v + 3 < K) dst[4 * v + 3] = w;
    }
}
// a value the compiler can neither rematerialise nor move into a uniform register
__device__ __forceinline__ int opaque_reg(int v) {
    asm volatile("mov.b32 %0, %0;" : "+r"(v));
    return v;
}

// Same row from a 16-bit profile: 8 entries per uint4, [code][v16][lane]; entry c of a lane sits in
// half (c & 1) of word (c >> 1) & 3 of vector c >> 3.
template <int K>
__device__ __forceinline__ void load_vec16(int (&dst)[K], const uint4* __restrict__ src) {
    constexpr int V16 = (K + 7) / 8;
#pragma unroll
    for (int v = 0; v < V16; ++v) {
        const uint4 x = src[v * 32];
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (8 * v + 2 * e + 0 < K) dst[8 * v + 2 * e + 0] = (int)(short)(w[e] & 0xffffu);
            if (8 * v + 2 * e + 1 < K) dst[8 * v + 2 * e + 1] = (int)w[e] >> 16;
        }
    }
}

// One systolic step of one lane: a row of K cells.  Hold = H of the previous row in this
// lane's columns, Hnew = H of this row (the caller ping-pongs the two arrays so no register
// copies are needed).  `one` is the runtime constant 1: `x * one + y` keeps the two plain
// additions of the cell on the FMA pipe (IMAD) instead of the ALU pipe, which the six
// LOP3 / VIMNMX3 / VIADDMNMX instructions saturate.
// LOCAL (Smith-Waterman, bioshell-seq/src/alignment/local.rs:120-203): the same recurrences
// with the three maxima clamped at zero (the DPX *_relu forms).  A value whose score field is
// <= 0 is the reference's STOP: it carries direction 0 and clears to exactly 0 under MASK.
template <int K, bool DIRS, bool LOCAL = false, bool TAG = false>
__device__ __forceinline__ void cell_row(const int (&Hold)[K], int (&Hnew)[K], int (&Fr)[K],
                                         const int (&T)[K], int hd, int& er, const Consts& cs,
                                         const int one, const int cf, uint32_t (&dw)[KTraits<K>::W],
                                         int& rowmax) {
    if constexpr (TAG) {
        // the frame cell: 4 ALU-pipe instructions (VIMNMX3, LOP3, 2 VIADDMNMX) + 2 IMAD; `cf` is this
        // step's F opening addend ((go - ge) << sh | prio 1 | x = 15 - step mod 16)
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const int d = hd * one + T[c];
            const int h = max3_s32(d, er, Fr[c]);
            const int hc = h & cs.MASK;
            const int hge = hc * one + cs.GOE;
            er = addmax_s32(er, cs.X1, hge);
            Fr[c] = addmax_s32(hc, cf, Fr[c]);
            hd = Hold[c];
            Hnew[c] = hc;
        }
    } else {
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const int e = er | cs.PH;
        const int f = Fr[c] | cs.PV;
        const int d = hd * one + T[c];
        const int h = LOCAL ? __vimax3_s32_relu(d, e, f) : max3_s32(d, e, f);
        if (DIRS) {
            uint32_t hdir = (uint32_t)h & 3u;
            if (LOCAL) hdir = h >= 4 ? hdir : 0u;      // score <= 0: STOP (arrows = 0, local.rs:158-181)
            const uint32_t nib = hdir | ((((uint32_t)er | (uint32_t)Fr[c]) & 3u) << 2);
            dw[c >> 3] = (dw[c >> 3] << 4) | nib;
        }
        const int hc = h & cs.MASK;
        const int hg = hc * one + cs.GO;
        er = LOCAL ? __viaddmax_s32_relu(e, cs.GE, hg) : addmax_s32(e, cs.GE, hg);
        Fr[c] = LOCAL ? __viaddmax_s32_relu(f, cs.GE, hg) : addmax_s32(f, cs.GE, hg);
        if (LOCAL) rowmax = rowmax > hc ? rowmax : hc;
        hd = Hold[c];
        Hnew[c] = hc;
    }
    }
}

// best cell of a local alignment as one lane sees it (first strict maximum in row-major order)
struct LaneBest { int v; uint32_t row, col; };

// Stream the residues codes[g0 .. g1) (whole sequences, back to back, last residue
// of each flagged) through the 32 lanes for ONE column block of the template.
//   first: this block starts at template column 0 (left border is generated),
//          otherwise lane 0 reads the boundary column from `scratch`.
//   lastp: this block holds the template's last column (results are emitted),
//          otherwise lane 31 writes the boundary column to `scratch`.
// The step loop is unrolled U times (H ping-pongs between two register arrays, so no
// copies).  A group of U steps in which NO lane meets an end-of-sequence flag runs the
// fast body (no flag tests, no reconvergence points); otherwise the whole warp takes the
// checked body.  Extra steps past the end just run into the padding.
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Wait until *p >= need.  Every ld.acquire.gpu is followed by an L1 invalidate (CCTL.IVALL) that
// also evicts the co-resident warps' lines, and a waiting block polls thousands of times: with
// BSA_WAVE_POLL the polls are relaxed loads with a short sleep in between and ONE acquire follows.
__device__ __forceinline__ uint32_t wait_progress(const uint32_t* p, uint32_t need) {
    uint32_t v;
#if BSA_WAVE_POLL
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= need) break;
        __nanosleep(64);
    }
    v = ld_acquire_u32(p);
#else
    while ((v = ld_acquire_u32(p)) < need) {}
#endif
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr uint32_t kWavePub = (BSA_WAVE_RING && BSA_WAVE_BATCH == 16) ? 16u : 32u;   // rows per boundary publication
// WAVE: the column blocks of one long pair run CONCURRENTLY in different warps (K3, the
// intra-task wavefront).  The block to the left publishes how many boundary entries it has
// written (`prog_out`, release store every 32 rows); this block waits on `prog_in` (acquire)
// before it consumes them.  Without WAVE, scratch_out == scratch and the pointers are null.
// HALF: two templates share the warp (lanes 0-15 / 16-31, see gotoh_pair_kernel); lane_last,
// slot_last and out_idx0 are then per-lane values and positions count from the half's first lane.
template <int K, bool DIRS, bool MULTI, bool WAVE = false, bool LOCAL = false, bool HALF = false,
          bool TAG = false, bool P16 = false>
__device__ __forceinline__ void stream_block(const uint8_t* __restrict__ codes, uint64_t g0,
                                             uint64_t g1, const uint4* prof, const uint4* rsH,
                                             const uint4* rsF, const int lane, const bool first,
                                             const bool lastp, const int lane_last,
                                             const int slot_last, const int hdiag0,
                                             const Consts cs, const int one,
                                             uint2* scratch,
                                             int32_t* __restrict__ scores,
                                             uint32_t* __restrict__ nident, uint64_t out_idx0,
                                             uint32_t* __restrict__ dirs, uint2* scratch_out = nullptr,
                                             const uint32_t* prog_in = nullptr,
                                             uint32_t* prog_out = nullptr, LaneBest* lane_best = nullptr,
                                             const uint32_t colbase = 0, const int one2 = 1,
                                             const int m_emit = 0, const uint64_t* __restrict__ lutp = nullptr) {
    static_assert(!TAG || (!DIRS && !LOCAL && !WAVE), "the TAG cell carries no direction bits");
    constexpr int W = KTraits<K>::W;
    constexpr int ROWB = (P16 ? ((K + 7) / 8) * 32 : KTraits<K>::ROW) * (int)sizeof(uint4);
    if (!WAVE) scratch_out = scratch;
    uint32_t avail = 0;   // WAVE: boundary entries known to be published by the left block
    constexpr int U = (DIRS || MULTI) ? 2 : (K <= 10 ? 4 : 2);
    const uint32_t X = (uint32_t)(g1 - g0);
    const int span = HALF ? 15 : ((MULTI && !lastp) ? 31 : lane_last);
    const uint32_t nsteps = (X + (uint32_t)span + (U - 1)) / U * U;
    const int lrel = HALF ? (lane & 15) : lane;     // lane index inside its template

    int Ha[K], Hb[K], Fr[K], T[K];
    load_vec<K>(Ha, rsH + lane);
    load_vec<K>(Fr, rsF + lane);
    int hdiag = hdiag0;
    int hb = cs.hb0;
    int oh = 0, oe = 0;
    uint32_t emitted = 0;
    uint32_t qstart = 0;   // TAG: stream position at which this lane's current query began (the frame needs its length)
    int cf = cs.GOF;       // TAG: F opening addend of the current step
    const bool border = !MULTI || first;
    const bool lane0 = lrel == 0;
    const char* prof_lane = reinterpret_cast<const char*>(prof + lane);
    LaneBest lbest{0, 0u, 0u};

    const uint8_t* p = codes + g0 - lrel;   // lane's position at step 0 (may sit in the padding)
    // residues are fetched PD groups ahead (WAVE: a row is short next to an L2 round trip)
    constexpr int PD = WAVE ? BSA_WAVE_PD : 1;
    uint32_t b[U], nb[PD][U];
#pragma unroll
    for (int u = 0; u < U; ++u) b[u] = ld_code(p + u);
#pragma unroll
    for (int g = 0; g + 1 < PD; ++g) {
#pragma unroll
        for (int u = 0; u < U; ++u) nb[g][u] = ld_code(p + (g + 1) * U + u);
    }
    uint2 sc_next = make_uint2(0u, 0u);
    // Boundary column of the block to the left (multi-pass templates).  Without WAVE it is complete
    // before this pass starts, so the warp fetches it 32 entries at a time, one per lane and a whole
    // batch ahead (an L2 round trip is longer than a step), and lane 0 takes entry S by shuffle.
    // With WAVE the left block is still running: it publishes its boundary in batches of 32 entries
    // (prog_out), so this block takes a batch with ONE warp-wide load (L1 bypassed) as soon as it
    // is published, half a batch before its first entry is needed (kWavePf), and runs 56+ rows
    // behind its left neighbour.  (Lane 0 loading entry S+1 at step S made every row wait for an
    // L2 round trip: ~830 cycles per row on B200.)
    constexpr bool WRING = MULTI && WAVE && BSA_WAVE_RING;
    constexpr bool RING = MULTI && ((!WAVE && BSA_RING) || WRING);
    constexpr uint32_t kWavePf = BSA_WAVE_PF;   // step inside a batch at which the next batch is fetched (even)
    constexpr uint32_t WB = WRING ? BSA_WAVE_BATCH : 32u;   // entries per batch (the non-WAVE ring always uses 32)
    static_assert(BSA_WAVE_BATCH == 32 || BSA_WAVE_BATCH == 16, "batch of 32 or 16 boundary entries");
    static_assert(BSA_WAVE_PF % 2 == 0 && BSA_WAVE_PF < BSA_WAVE_BATCH, "fetch step: even, inside the batch");
    uint2 ring_cur = make_uint2(0u, 0u), ring_nxt = make_uint2(0u, 0u);
    if (WAVE && !first && X > 0) {
        const uint32_t need0 = WRING ? (X < WB ? X : WB) : 1u;
        if (lane0) avail = wait_progress(prog_in, need0);
        avail = __shfl_sync(0xffffffffu, avail, 0);
        __syncwarp();
    }
    if (RING && !first) {
        if (WRING) {
            if ((uint32_t)lane < X && (uint32_t)lane < WB) ring_cur = __ldcg(scratch + lane);
        } else {
            if ((uint32_t)lane < X) ring_cur = scratch[lane];
            if (32u + (uint32_t)lane < X) ring_nxt = scratch[32 + lane];
        }
    }
    if (MULTI && !RING && !first && lane0 && X > 0) sc_next = scratch[0];

    // one step; HO = previous row, HN = this row
#define BSA_STEP_CORE(HO, HN, B, S)                                                               \
        if (P16) load_vec16<K>(T, reinterpret_cast<const uint4*>(prof_lane + ((B)&kCodeMask) * ROWB)); \
        else load_vec<K>(T, reinterpret_cast<const uint4*>(prof_lane + ((B)&kCodeMask) * ROWB));  \
        int hin = __shfl_up_sync(0xffffffffu, oh, 1);                                             \
        int er = __shfl_up_sync(0xffffffffu, oe, 1);                                              \
        if (RING && !first) {                                                                     \
            sc_next.x = __shfl_sync(0xffffffffu, ring_cur.x, (S) & (WB - 1u));                    \
            sc_next.y = __shfl_sync(0xffffffffu, ring_cur.y, (S) & (WB - 1u));                    \
            if (((S) & (WB - 1u)) == WB - 1u) {                                                   \
                ring_cur = ring_nxt;                                                              \
                if (!WRING && (S) + 33u + (uint32_t)lane < X) ring_nxt = scratch[(S) + 33u + lane]; \
            }                                                                                     \
        }                                                                                         \
        if (lane0) {                                                                              \
            if (border) { /* H[i][0] = go + (i-1) ge; E[i][1] opens from it (global.rs:96-101) */ \
                hin = LOCAL ? 0 : hb;            /* local: H[i][0] = E[i][1] = 0 (local.rs:116) */ \
                er = LOCAL ? 0 : hb + cs.GOE;                                                     \
            } else {                                                                              \
                hin = (int)sc_next.x;                                                             \
                er = (int)sc_next.y;                                                              \
            }                                                                                     \
        }                                                                                         \
        if (TAG) er &= cs.XCLR;   /* the streak restarts in every lane */                         \
        if (MULTI && !RING && !first && lane0 && (S) + 1 < X) sc_next = scratch[(S) + 1];         \
        if (!TAG) hb += cs.GEB;          /* TAG: the frame keeps the left border constant */       \
        const int hd = hdiag;                                                                     \
        hdiag = hin;                                                                              \
        uint32_t dw[W];                                                                           \
        _Pragma("unroll") for (int w = 0; w < W; ++w) dw[w] = 0u;                                 \
        int rowmax = 0;                                                                           \
        cell_row<K, DIRS, LOCAL, TAG>(HO, HN, Fr, T, hd, er, cs, one, TAG ? cf : one2, dw, rowmax); \
        if (TAG) cf -= cs.X1;            /* the next step's openings rank below this step's */      \
        if (LOCAL && rowmax > lbest.v && (S) - (uint32_t)lane < X) {                              \
            /* a new best in this lane: remember the row and the first column that holds it */   \
            lbest.v = rowmax;                                                                     \
            lbest.row = (S) - (uint32_t)lane;                                                     \
            uint32_t cc = 0;                                                                      \
            _Pragma("unroll") for (int c = K - 1; c >= 0; --c) if (HN[c] == rowmax) cc = (uint32_t)c; \
            lbest.col = colbase + (uint32_t)lane * K + cc;                                        \
        }                                                                                         \
        oh = HN[K - 1];                                                                           \
        oe = er;                                                                                  \
        if (DIRS) {                                                                               \
            uint32_t* dp = dirs + ((size_t)(S)*32 + lane) * W;                                    \
            _Pragma("unroll") for (int w = 0; w < W; ++w) dp[w] = dw[w];                          \
        }                                                                                         \
        if (MULTI && !lastp && lane == 31) {                                                      \
            const uint32_t pos = (S)-31u; /* wraps for S < 31 -> fails the bound test */          \
            if (pos < X) {                                                                        \
                scratch_out[pos] = make_uint2((uint32_t)oh, (uint32_t)oe);                        \
                if (WAVE && ((pos & (kWavePub - 1u)) == kWavePub - 1u || pos + 1u == X)) st_release_u32(prog_out, pos + 1u); \
            }                                                                                     \
        }
#define BSA_STEP_FAST(HO, HN, B, S) { BSA_STEP_CORE(HO, HN, B, S) }
#define BSA_STEP(HO, HN, B, S)                                                                    \
    {                                                                                             \
        BSA_STEP_CORE(HO, HN, B, S)                                                               \
        if ((B)&kLastFlag) {                                                                      \
            /* end of a query: emit H[n][m] from the lane that owns column m, then put the */     \
            /* lane back on the top border for the next query of the stream.               */     \
            const uint32_t pos = (S) - (uint32_t)lrel;                                            \
            const bool valid = pos < X;                                                           \
            if (!LOCAL && lastp && valid && lane == lane_last) {                                  \
                int v = 0;                                                                        \
                _Pragma("unroll") for (int c = 0; c < K; ++c) if (c == slot_last) v = HN[c];      \
                const uint64_t k = out_idx0 + (lutp ? lutp[emitted] : (uint64_t)emitted);         \
                /* TAG: out of the frame, H = H* + (n + m) ge with n = rows of this query */       \
                if (scores) scores[k] = (v >> cs.sh) + (TAG ? (int)(pos + 1u - qstart + (uint32_t)m_emit) * cs.ge : 0); \
                if (nident) nident[k] = (uint32_t)v & ((1u << cs.cs) - 1u);                       \
            }                                                                                     \
            emitted += valid ? 1u : 0u;                                                           \
            qstart = pos + 1u;                                                                    \
            load_vec<K>(HN, rsH + lane);                                                          \
            load_vec<K>(Fr, rsF + lane);                                                          \
            hdiag = hdiag0;                                                                       \
            hb = cs.hb0;                                                                          \
        }                                                                                         \
    }

    for (uint32_t s = 0; s < nsteps; s += U) {
        if (WRING && !first) {
            if ((s & (WB - 1u)) == kWavePf) {
                const uint32_t base = (s & ~(WB - 1u)) + WB;       // first entry of the next batch
                if (base < X) {
                    const uint32_t need = base + WB < X ? base + WB : X;
                    if (avail < need) {
                        if (lane0) avail = wait_progress(prog_in, need);
                        avail = __shfl_sync(0xffffffffu, avail, 0);
                        __syncwarp();
                    }
                    if (base + (uint32_t)lane < X && (uint32_t)lane < WB) ring_nxt = __ldcg(scratch + base + lane);
                }
            }
        } else if (WAVE && !first) {
            // this group prefetches boundary entries up to index s + U
            const uint32_t need = s + U + 1u < X ? s + U + 1u : X;
            if (avail < need) {
                if (lane0) avail = wait_progress(prog_in, need);
                avail = __shfl_sync(0xffffffffu, avail, 0);
            }
        }
        if (TAG && (s & (kTagRows - 1u)) == 0u) {
            // a new block of 16 steps: what is stored outranks every opening of the block
#pragma unroll
            for (int c = 0; c < K; ++c) Fr[c] |= cs.XTOP;
            cf = cs.GOF + (int)(kTagRows - 1u) * cs.X1;
        }
        uint32_t any = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            nb[PD - 1][u] = ld_code(p + PD * U + u);   // prefetch the residues of the group PD ahead
            any |= b[u];
        }
        p += U;
        if (!__any_sync(0xffffffffu, any & kLastFlag)) {
#pragma unroll
            for (int u = 0; u < U; u += 2) {
                BSA_STEP_FAST(Ha, Hb, b[u], s + u)
                BSA_STEP_FAST(Hb, Ha, b[u + 1], s + u + 1)
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; u += 2) {
                BSA_STEP(Ha, Hb, b[u], s + u)
                BSA_STEP(Hb, Ha, b[u + 1], s + u + 1)
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) b[u] = nb[0][u];
#pragma unroll
        for (int g = 0; g + 1 < PD; ++g) {
#pragma unroll
            for (int u = 0; u < U; ++u) nb[g][u] = nb[g + 1][u];
        }
    }
#undef BSA_STEP
#undef BSA_STEP_FAST
#undef BSA_STEP_CORE
    if (LOCAL) *lane_best = lbest;
}

// ---------------------------------------------------------------------------------------------
// TWO ROWS PER STEP (TAG cell, one column block, score + identity).  The dependency chain of a
// row is four instructions per cell (VIMNMX3 -> LOP3 -> IMAD -> VIADDMNMX) and with 8-16 warps
// per SM the ALU pipe waits on it.  Here the lanes are skewed by TWO stream positions: at double
// step S lane l works on rows 2(S-l) and 2(S-l)+1, cell (r+1, c) right after cell (r, c), so the
// warp always has two independent chains in flight.  H is updated in place (one array: the
// diagonal values travel in two scalars), the two profile rows take the registers the ping-pong
// copy of H used to take.  A double step in which some lane passes an end-of-sequence flag runs
// its two rows one after the other with the flag handling in between.
// Which column counts take the two-row step: measured per K on B200 (gpurun_out/qb_groups_tag*.log);
// where the register budget forces spills (K = 11, 12 at three CTAs per SM, K = 19/20) the
// one-row step is as fast or faster.
#ifndef BSA_TR_FORCE
#define BSA_TR_FORCE (-1)   // A/B only: 1 = two-row step for every K, 0 = for none, -1 = the measured table below
#endif
// With the 6-instruction frame cell the two-row step wins (or ties) for every K, on 32 and on 16 lanes
// (same-box per-K A/B, profiles/r2_ab_two_rows_per_k.txt); round 1's 7-instruction cell spilled at some K.
template <int K, bool HALF>
struct TwoRows {
    static constexpr bool value = BSA_TR_FORCE >= 0 ? (BSA_TR_FORCE != 0) : (BSA_TWO_ROWS != 0);
};

template <int K, bool HALF>
__device__ __forceinline__ void stream_block_tag2(const uint8_t* __restrict__ codes, uint64_t g0, uint64_t g1,
                                                  const uint4* prof, const uint4* rsH, const uint4* rsF,
                                                  const int lane, const int lane_last, const int slot_last,
                                                  const int hdiag0, const Consts cs, const int one, const int one2,
                                                  int32_t* __restrict__ scores, uint32_t* __restrict__ nident,
                                                  uint64_t out_idx0, const int m_emit,
                                                  const uint64_t* __restrict__ lutp) {
    constexpr int ROWB = KTraits<K>::ROW * (int)sizeof(uint4);
    constexpr int U = BSA_TAG2_U;  // double steps per loop iteration (an even count lets H[] return to its registers)
    static_assert(8 % U == 0, "the stored F get the top tag bit every 8 double steps");
    const uint32_t X = (uint32_t)(g1 - g0);
    const int span = HALF ? 15 : lane_last;
    const uint32_t nd = ((X + 1u) / 2u + (uint32_t)span + (U - 1)) / U * U;
    const int lrel = HALF ? (lane & 15) : lane;
    const bool lane0 = lrel == 0;

    int H[K], Fr[K], T0[K], T1[K];
    load_vec<K>(H, rsH + lane);
    load_vec<K>(Fr, rsF + lane);
    int hdiag = hdiag0;
    const int hb = cs.hb0, eb = cs.hb0 + cs.GOE;   // the frame's constant left border: H*[i][0], E*[i][1]
    int cf = cs.GOF;                               // F opening addend of the first row of the current double step
    int oh0 = 0, oe0 = 0, oh1 = 0, oe1 = 0;
    uint32_t emitted = 0, qstart = 0;
    const char* prof_lane = reinterpret_cast<const char*>(prof + lane);
    const uint8_t* p = codes + g0 - 2 * lrel;   // lane's first row at double step 0 (may sit in the padding)
    uint32_t b[2 * U], nb[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) b[u] = ld_code(p + u);

    // one row, in place (the checked path)
#define BSA_ROW1(TT, HIN, ER, OH, OE, CF)                                                         \
    {                                                                                             \
        if (lane0) { HIN = hb; ER = eb; }                                                         \
        ER &= cs.XCLR;                                                                            \
        int hd = hdiag;                                                                           \
        hdiag = HIN;                                                                              \
        _Pragma("unroll") for (int c = 0; c < K; ++c) {                                           \
            const int d = hd * one + TT[c];                                                       \
            const int h = max3_s32(d, ER, Fr[c]);                                                 \
            const int hc = h & cs.MASK;                                                           \
            ER = addmax_s32(ER, cs.X1, hc * one + cs.GOE);                                        \
            Fr[c] = addmax_s32(hc, CF, Fr[c]);                                                    \
            hd = H[c];                                                                            \
            H[c] = hc;                                                                            \
        }                                                                                         \
        OH = H[K - 1];                                                                            \
        OE = ER;                                                                                  \
    }
#define BSA_FLAG1(B, POS)                                                                         \
    if ((B)&kLastFlag) {                                                                          \
        const bool valid = (POS) < X;                                                             \
        if (valid && lane == lane_last) {                                                         \
            int v = 0;                                                                            \
            _Pragma("unroll") for (int c = 0; c < K; ++c) if (c == slot_last) v = H[c];           \
            const uint64_t k = out_idx0 + (lutp ? lutp[emitted] : (uint64_t)emitted);             \
            /* out of the frame: H = H* + (n + m) ge, n = rows of this query */                   \
            if (scores) scores[k] = (v >> cs.sh) + (int)((POS) + 1u - qstart + (uint32_t)m_emit) * cs.ge; \
            if (nident) nident[k] = (uint32_t)v & ((1u << cs.cs) - 1u);                           \
        }                                                                                         \
        emitted += valid ? 1u : 0u;                                                               \
        qstart = (POS) + 1u;                                                                      \
        load_vec<K>(H, rsH + lane);                                                               \
        load_vec<K>(Fr, rsF + lane);                                                              \
        hdiag = hdiag0;                                                                           \
    }

    // two rows interleaved, in place (needs: no lane ends a sequence on the FIRST of the two rows)
#define BSA_PAIR2()                                                                               \
    {                                                                                             \
        if (lane0) {                                                                              \
            hin0 = hb;                                                                            \
            er0 = eb;                                                                             \
            hin1 = hb;                                                                            \
            er1 = eb;                                                                             \
        }                                                                                         \
        er0 &= cs.XCLR;                                                                           \
        er1 &= cs.XCLR;                                                                           \
        int hd0 = hdiag, hd1 = hin0;                                                              \
        hdiag = hin1;                                                                             \
        int hc1 = 0;                                                                              \
_Pragma("unroll")                                                                                 \
        for (int c = 0; c < K; ++c) {                                                             \
            const int d0 = hd0 * one + T0[c];                                                     \
            const int h0 = max3_s32(d0, er0, Fr[c]);                                              \
            const int hc0 = h0 & cs.MASK;                                                         \
            er0 = addmax_s32(er0, cs.X1, hc0 * one + cs.GOE);                                     \
            const int f1 = addmax_s32(hc0, cf0, Fr[c]);                                           \
            hd0 = H[c];                                                                           \
            const int d1 = hd1 * one2 + T1[c];                                                    \
            const int h1 = max3_s32(d1, er1, f1);                                                 \
            hc1 = h1 & cs.MASK;                                                                   \
            er1 = addmax_s32(er1, cs.X1, hc1 * one2 + cs.GOE);                                    \
            Fr[c] = addmax_s32(hc1, cf1, f1);                                                     \
            hd1 = hc0;                                                                            \
            H[c] = hc1;                                                                           \
        }                                                                                         \
        oh0 = hd1;                                                                                \
        oe0 = er0;                                                                                \
        oh1 = hc1;                                                                                \
        oe1 = er1;                                                                                \
    }

    for (uint32_t S = 0; S < nd; S += U) {
        if ((S & 7u) == 0u) {
            // a new block of 16 rows: what is stored outranks every opening of the block
#pragma unroll
            for (int c = 0; c < K; ++c) Fr[c] |= cs.XTOP;
            cf = cs.GOF + (int)(kTagRows - 1u) * cs.X1;
        }
        uint32_t any = 0;
#pragma unroll
        for (int u = 0; u < 2 * U; ++u) {
            nb[u] = ld_code(p + 2 * U + u);   // prefetch the next group's residues
            any |= b[u];
        }
        p += 2 * U;
        if (!__any_sync(0xffffffffu, any & kLastFlag)) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                load_vec<K>(T0, reinterpret_cast<const uint4*>(prof_lane + (b[2 * u] & kCodeMask) * ROWB));
                load_vec<K>(T1, reinterpret_cast<const uint4*>(prof_lane + (b[2 * u + 1] & kCodeMask) * ROWB));
                int hin0 = __shfl_up_sync(0xffffffffu, oh0, 1);
                int er0 = __shfl_up_sync(0xffffffffu, oe0, 1);
                int hin1 = __shfl_up_sync(0xffffffffu, oh1, 1);
                int er1 = __shfl_up_sync(0xffffffffu, oe1, 1);
                const int cf0 = cf, cf1 = cf - cs.X1;
                cf -= 2 * cs.X1;
                BSA_PAIR2()
            }
        } else {
            BSA_FLAG_UNROLL
            for (int u = 0; u < U; ++u) {
                const uint32_t pos0 = 2u * (S + u - (uint32_t)lrel);   // wraps below row 0: fails pos < X
                load_vec<K>(T0, reinterpret_cast<const uint4*>(prof_lane + (BSA_B(u, 0) & kCodeMask) * ROWB));
                load_vec<K>(T1, reinterpret_cast<const uint4*>(prof_lane + (BSA_B(u, 1) & kCodeMask) * ROWB));
                int hin0 = __shfl_up_sync(0xffffffffu, oh0, 1);
                int er0 = __shfl_up_sync(0xffffffffu, oe0, 1);
                int hin1 = __shfl_up_sync(0xffffffffu, oh1, 1);
                int er1 = __shfl_up_sync(0xffffffffu, oe1, 1);
                const int cf0 = cf, cf1 = cf - cs.X1;
                cf -= 2 * cs.X1;
                if (BSA_FLAG_SPLIT && !__any_sync(0xffffffffu, BSA_B(u, 0) & kLastFlag)) {
                    // flags only on the second row: the interleaved step stays valid, reset afterwards
                    BSA_PAIR2()
                    BSA_FLAG1(BSA_B(u, 1), pos0 + 1u)
                } else {
                    BSA_ROW1(T0, hin0, er0, oh0, oe0, cf0)
                    BSA_FLAG1(BSA_B(u, 0), pos0)
                    BSA_ROW1(T1, hin1, er1, oh1, oe1, cf1)
                    BSA_FLAG1(BSA_B(u, 1), pos0 + 1u)
                }
                if (BSA_FLAG_LOOP) {   // rolled loop: the next double step's residues move to b[0], b[1]
#pragma unroll
                    for (int k = 0; k + 2 < 2 * U; ++k) b[k] = b[k + 2];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2 * U; ++u) b[u] = nb[u];
    }
#undef BSA_ROW1
#undef BSA_FLAG1
#undef BSA_PAIR2
}


// ---------------------------------------------------------------------------------------------
// Two rows per step over the EVEN-ALIGNED copy of the stream store (BSA_ALIGNED).  Every sequence
// starts at an even stream position and takes an even number of bytes: an odd-length sequence is
// preceded by one PAD row (residue code 0), whose profile row adds nothing on the diagonal of the
// moving frame and therefore reproduces the top border (H* = go - ge, the stored F keep their top
// tag bit; tests/packed_model.py::frame_align(pad=True) is the scalar model, checked against the
// oracle).  The end-of-sequence flag then always sits on the SECOND row of a double step, so the
// interleaved two-row body is the only form of the cell code: no sequential-row path, half the
// code, and a flagged step costs the interleaved step plus the reset of the flagged lanes.
// The one special case: the diagonal input of the first real row in DP column 1 is H*[0][0] = 0,
// which lane 0 takes instead of the left border when the row above is the PAD row.
// `qoffp` = the ORIGINAL offsets of the chunk's first sequence (the emitted score leaves the frame
// with the real number of rows); the left border enters lane 0 without selects (IMAD / LOP3).
template <int K, bool HALF>
__device__ __forceinline__ void stream_block_tag2a(const uint8_t* __restrict__ codes, uint64_t g0, uint64_t g1,
                                                   const uint4* prof, const uint4* rsH, const uint4* rsF,
                                                   const int lane, const int lane_last, const int slot_last,
                                                   const int hdiag0, const Consts cs, const int one, const int one2,
                                                   int32_t* __restrict__ scores, uint32_t* __restrict__ nident,
                                                   uint64_t out_idx0, const int m_emit,
                                                   const uint64_t* __restrict__ lutp,
                                                   const uint64_t* __restrict__ qoffp) {
    constexpr int ROWB = KTraits<K>::ROW * (int)sizeof(uint4);
    constexpr int U = K <= BSA_TAG2_U_SMALLK ? 8 : BSA_TAG2_U;  // double steps per loop iteration
    static_assert(8 % U == 0, "the stored F get the top tag bit every 8 double steps");
    const uint32_t X = (uint32_t)(g1 - g0);        // even
    const int span = HALF ? 15 : lane_last;
    const uint32_t nd = (X / 2u + (uint32_t)span + (U - 1)) / U * U;
    const int lrel = HALF ? (lane & 15) : lane;
    const bool lane0 = lrel == 0;
    // the constant left border of the frame, H*[i][0] = go - ge and E*[i][1], enters lane 0 as
    // hin = shuffled * nl0 + hbl (IMAD) and er = (shuffled & keepx) | ebl (one LOP3, which also
    // clears the streak field of the E that crossed the lane boundary)
    const int nl0 = opaque_reg(lane0 ? 0 : one);
    const int hbl = opaque_reg(lane0 ? cs.hb0 : 0);
#if BSA_ETAG
    // E openings carry the tag K-1-c of their column, so an older opening outranks a newer one on ties and the
    // extension needs no addition at all: E = max(H + GOE_c, E) is ONE VIADDMNMX with a warp-uniform constant.
    // An E that crosses a lane boundary gets the full tag field (older than every opening of the next lane).
    const int xall = ~cs.XCLR;
    const int ebl = opaque_reg(lane0 ? (cs.hb0 + cs.GOE) | xall : xall);
    const int keepx = opaque_reg(lane0 ? 0 : -1);
#else
    const int ebl = opaque_reg(lane0 ? cs.hb0 + cs.GOE : 0);
    const int keepx = opaque_reg(lane0 ? 0 : cs.XCLR);
#endif
    const int one_a = opaque_reg(one), one_b = opaque_reg(one2);

    int H[K], Fr[K], T0[K], T1[K];
    load_vec<K>(H, rsH + lane);
    load_vec<K>(Fr, rsF + lane);
    int hdiag = hdiag0;
    int cf = cs.GOF;                               // F opening addend of the first row of the current double step
    int oh0 = 0, oe0 = 0, oh1 = 0, oe1 = 0;
    uint32_t emitted = 0;
    const char* prof_lane = reinterpret_cast<const char*>(prof + lane);
    const uint8_t* p = codes + g0 - 2 * lrel;   // lane's first row at double step 0 (may sit in the padding)
    uint32_t b[2 * U], nb[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) b[u] = ld_code(p + u);

    for (uint32_t S = 0; S < nd; S += U) {
        if ((S & 7u) == 0u) {
            // a new block of 16 rows: what is stored outranks every opening of the block
#pragma unroll
            for (int c = 0; c < K; ++c) Fr[c] |= cs.XTOP;
            cf = cs.GOF + (int)(kTagRows - 1u) * cs.X1;
        }
#pragma unroll
        for (int u = 0; u < 2 * U; ++u) nb[u] = ld_code(p + 2 * U + u);   // prefetch the next group's residues
        p += 2 * U;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t b0 = b[2 * u], b1 = b[2 * u + 1];     // b0 never carries a flag
            int hq = 0;
            if constexpr (BSA_PADSLOT && K % 4 != 0)
                load_vec_x<K>(T0, hq, reinterpret_cast<const uint4*>(prof_lane + b0 * ROWB));
            else {
                load_vec<K>(T0, reinterpret_cast<const uint4*>(prof_lane + b0 * ROWB));
                hq = b0 == kPadCode ? 0 : hbl;
            }
            load_vec<K>(T1, reinterpret_cast<const uint4*>(prof_lane + (b1 & kCodeMask) * ROWB));
            const int sh0 = __shfl_up_sync(0xffffffffu, oh0, 1);
            const int se0 = __shfl_up_sync(0xffffffffu, oe0, 1);
            const int sh1 = __shfl_up_sync(0xffffffffu, oh1, 1);
            const int se1 = __shfl_up_sync(0xffffffffu, oe1, 1);
            // row 1's diagonal input in this lane's first column = row 0's H to the left; for lane 0 the
            // border, or H*[0][0] = 0 when row 0 is the PAD row (hq: from the profile row's spare slot)
            const int hin0 = sh0 * nl0 + hq;
            const int hin1 = sh1 * nl0 + hbl;
            int er0 = (se0 & keepx) | ebl;
            int er1 = (se1 & keepx) | ebl;
            const int cf0 = cf, cf1 = cf - cs.X1;
            cf -= 2 * cs.X1;
            int hd0 = hdiag, hd1 = hin0;
            hdiag = hin1;
            int hc1 = 0;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                const int d0 = hd0 * one_a + T0[c];
                const int h0 = max3_s32(d0, er0, Fr[c]);
                const int hc0 = h0 & cs.MASK;
#if BSA_ETAG
                er0 = addmax_s32(hc0, cs.GOE + (K - 1 - c) * cs.X1, er0);
#else
                er0 = addmax_s32(er0, cs.X1, hc0 * one_a + cs.GOE);
#endif
                const int f1 = addmax_s32(hc0, cf0, Fr[c]);
                hd0 = H[c];
                const int d1 = hd1 * one_b + T1[c];
                const int h1 = max3_s32(d1, er1, f1);
                hc1 = h1 & cs.MASK;
#if BSA_ETAG
                er1 = addmax_s32(hc1, cs.GOE + (K - 1 - c) * cs.X1, er1);
#else
                er1 = addmax_s32(er1, cs.X1, hc1 * one_b + cs.GOE);
#endif
                Fr[c] = addmax_s32(hc1, cf1, f1);
                hd1 = hc0;
                H[c] = hc1;
            }
            oh0 = hd1;
            oe0 = er0;
            oh1 = hc1;
            oe1 = er1;
            if (b1 & kLastFlag) {
                // this lane has finished a sequence: emit (the lane that owns the last column), back to the top border
                const uint32_t pos1 = 2u * (S + u - (uint32_t)lrel) + 1u;   // wraps below row 0: fails pos1 < X
                const bool valid = pos1 < X;
                if (valid && lane == lane_last) {
                    int v = 0;
#pragma unroll
                    for (int c = 0; c < K; ++c) if (c == slot_last) v = H[c];
                    const uint32_t n = (uint32_t)(qoffp[emitted + 1] - qoffp[emitted]);
                    const uint64_t k = out_idx0 + (lutp ? lutp[emitted] : (uint64_t)emitted);
                    // out of the frame: H = H* + (n + m) ge, n = rows of this sequence
                    if (scores) scores[k] = (v >> cs.sh) + (int)(n + (uint32_t)m_emit) * cs.ge;
                    if (nident) nident[k] = (uint32_t)v & ((1u << cs.cs) - 1u);
                }
                emitted += valid ? 1u : 0u;
                load_vec_opaque<K>(H, rsH + lane);
                load_vec_opaque<K>(Fr, rsF + lane);
                hdiag = hdiag0;
            }
        }
#pragma unroll
        for (int u = 0; u < 2 * U; ++u) b[u] = nb[u];
    }
}

__device__ __forceinline__ uint32_t lower_bound_off(const uint64_t* __restrict__ off, uint32_t lo,
                                                    uint32_t hi, uint64_t x) {
    // first index i in [lo, hi] with off[i] >= x
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (off[mid] >= x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <int K>
__device__ __forceinline__ Consts make_consts(int go, int ge, int cshift, bool local = false,
                                              bool tag = false, bool flip = false) {
    Consts cs;
    cs.cs = cshift;
    cs.ps = cshift + (tag ? kTagBits : 0);
    cs.sh = cs.ps + 2;
    const int S = 1 << cs.sh;
    const int xmask = tag ? ((1 << kTagBits) - 1) << cshift : 0;
    cs.GEB = tag ? 0 : ge * S;
    cs.GE = tag ? 1 << cshift : ge * S;
    cs.GO = go * S;
    cs.MASK = ~((3 << cs.ps) | xmask);
    cs.XCLR = ~xmask;
    cs.X1 = 1 << cshift;
    cs.XTOP = tag ? 1 << (cshift + kTagBits - 1) : 0;
    cs.TSUB = tag ? -2 * ge : 0;
    cs.ge = ge;
    cs.PH = (flip ? 1 : 2) << cs.ps;   // E: horizontal, gap in the query; beats F on ties (global.rs:166-169)
    cs.PV = (flip ? 2 : 1) << cs.ps;   // F: vertical, gap in the template   (flip: rows and columns swapped)
    cs.GOE = tag ? (go - ge) * S + cs.PH : cs.GO;
    cs.GOF = tag ? (go - ge) * S + cs.PV : cs.GO;
    // padded columns: neutral in global mode; in local mode they must never score, so that no
    // cell outside the template can reach the best score (anything they inherit through E/F is
    // strictly below a real cell of the same row)
    cs.T_PAD = local ? -(1 << 24) : 3 << cs.ps;
    cs.hb0 = tag ? (go - ge) * S : go * S;
    cs.local = local;
    return cs;
}

// Score + identity for (template, query range) items.  Persistent CTAs pull items from an
// atomic counter and build the template's profile once; the item's residue stream is cut
// into chunks of whole queries that the CTA's warps pull from a shared counter (warps do
// not progress at the same rate -- the issue arbiter is not fair -- so a static split
// would leave the fast warps waiting at the item's closing barrier).
template <int K>
struct MinBlocks { static constexpr int value = K <= 4 ? 4 : (K <= BSA_MB3_MAXK ? 3 : 2); };

constexpr uint32_t kChunkBig = BSA_CHUNK_BIG;    // stream residues per chunk (pipeline fill is 31 steps)
#ifndef BSA_CHUNK_SMALL
#define BSA_CHUNK_SMALL 640
#endif
constexpr uint32_t kChunkSmall = BSA_CHUNK_SMALL;

// Chunk schedule of one item's stream: big chunks over the first 14/16, small ones over the rest, so the warps
// reach the item's closing barrier within half a small chunk of each other.  The big chunks shrink with the
// stream so that every warp gets at least two of them, and a short stream (small problems: items of a few
// thousand residues) is cut into uniform pieces -- with fixed 4,096-residue chunks one warp swept most of such
// an item while the other seven waited at the barrier.
struct ChunkPlan { uint64_t head; uint32_t nbig, nsmall; };
__device__ __forceinline__ ChunkPlan plan_chunks(uint64_t span) {
    constexpr uint64_t W = kWarpsPerCta;
    ChunkPlan p;
    if (span < 4u * W * kChunkSmall) {
        uint64_t lb = span / (2 * W);
        if (lb < 192) lb = 192;
        p.head = span;
        p.nbig = (uint32_t)((span + lb - 1) / lb);
        p.nsmall = 0;
    } else {
        p.head = span - span * BSA_TAIL16 / 16;
        uint64_t lb = p.head / (2 * W);
        lb = lb < kChunkSmall ? kChunkSmall : (lb > kChunkBig ? kChunkBig : lb);
        p.nbig = (uint32_t)((p.head + lb - 1) / lb);
        p.nsmall = (uint32_t)((span - p.head + kChunkSmall - 1) / kChunkSmall);
        if (p.nsmall < (uint32_t)W) p.nsmall = (uint32_t)W;
    }
    return p;
}

template <int K, bool MULTI, bool TAG = false>
__global__ void __launch_bounds__(kThreads, MinBlocks<K>::value) gotoh_stream_kernel(const KArgs a) {
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_item;
    __shared__ uint32_t s_chunk;
    constexpr int ROW = KTraits<K>::ROW;
    uint4* prof = smem;
    uint4* rsH = smem + (size_t)a.C * ROW;
    uint4* rsF = rsH + ROW;
    const int lane = threadIdx.x & 31;
    BSA_TMA_PREAMBLE(K, a.C, a.subst)

    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.item_counter, 1u);
        __syncthreads();
        const uint32_t ii = s_item;
        if (ii >= a.n_items) break;
        const Item it = a.items[ii];
        const uint64_t t0 = a.T.off[it.t];
        const uint32_t m = (uint32_t)(a.T.off[it.t + 1] - t0);
        const uint8_t* tc = a.T.codes + t0;
        // aligned two-row blocks: the count field is as wide as the columns-per-lane class needs (a launch constant,
        // so every constant of the cell is warp-uniform); the host's range check uses the same width
        constexpr bool kUniformCs = BSA_ALIGNED && BSA_ETAG && TAG && !MULTI && TwoRows<K, false>::value;
        const int cs_class = (32 - __clz(32 * K)) < a.cs_cap ? (32 - __clz(32 * K)) : a.cs_cap;
        const Consts cs = make_consts<K>(a.go, a.ge, kUniformCs ? cs_class : (int)it.cshift, false, TAG, a.flip != 0);

        // chunk schedule: big chunks over the first 13/16 of the stream, small ones over the rest,
        // so the warps reach the item's closing barrier within half a small chunk of each other
        // the two-row TAG blocks stream the even-aligned copy of the store (chunks are whole sequences either way)
        constexpr bool kAligned = BSA_ALIGNED && TAG && !MULTI && TwoRows<K, false>::value;
        const uint64_t* __restrict__ xoff = kAligned ? a.QA.off : a.Q.off;
        const uint64_t x0 = xoff[it.q_begin], x1 = xoff[it.q_end];
        const uint64_t span = x1 - x0;
        const ChunkPlan cp = plan_chunks(span);
        const uint64_t head = cp.head;
        const uint32_t nbig = cp.nbig, nsmall = cp.nsmall;
        const uint32_t nch = nbig + nsmall;
        const uint32_t npass = MULTI ? (m + 32 * K - 1) / (32 * K) : 1u;
        // MULTI: the boundary column of the whole item, addressed by stream position
        uint2* scratch = MULTI ? a.scratch + (size_t)blockIdx.x * a.scratch_stride : nullptr;

        for (uint32_t pass = 0; pass < npass; ++pass) {
            const uint32_t colbase = pass * 32 * K;
            __syncthreads();   // previous profile / chunk counter / staged slice are no longer in use
            {
                BSA_TMA_PTRS(K, a.C)
                if (threadIdx.x == 0) {
                    s_chunk = 0;
                    const uint32_t nb = slice_bytes(tc, colbase, min(m - colbase, 32u * K));
                    TmaStage::arm(&s_mbar, nb);
                    TmaStage::copy(&s_mbar, s_tcA, slice_src(tc, colbase), nb);
                }
                TmaStage::wait(&s_mbar, s_phase);
                build_profile<K>(prof, rsH, rsF, s_subst, slice_view(s_tcA, tc, colbase), m, colbase, a, cs);
            }
            __syncthreads();
            BSA_TMA_FLIP()
            const bool lastp = (pass + 1 == npass);
            const int lane_last = (int)((m - 1 - colbase) / K);
            const int slot_last = (int)((m - 1 - colbase) % K);
            const long long jl = (long long)colbase + (long long)lane * K;  // DP column left of the lane
            const int hdiag0 = jl == 0 ? 0 : (TAG ? cs.hb0 : (int)((a.go + (jl - 1) * a.ge) * (1 << cs.sh)));
            for (;;) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(&s_chunk, 1u);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= nch) break;
                // chunk c = the queries whose first residue falls in its share of the stream
                const uint64_t ca = c <= nbig ? head * c / nbig : head + (span - head) * (c - nbig) / nsmall;
                const uint64_t cb = c + 1 <= nbig ? head * (c + 1) / nbig
                                                  : head + (span - head) * (c + 1 - nbig) / nsmall;
                const uint32_t qa = c == 0 ? it.q_begin : lower_bound_off(xoff, it.q_begin, it.q_end, x0 + ca);
                const uint32_t qb = c + 1 == nch ? it.q_end
                                                 : lower_bound_off(xoff, it.q_begin, it.q_end, x0 + cb);
                const uint64_t g0 = xoff[qa], g1 = xoff[qb];
                if (g1 <= g0) continue;
                if constexpr (kAligned)
                    stream_block_tag2a<K, false>(a.QA.codes, g0, g1, prof, rsH, rsF, lane, lane_last, slot_last,
                                                 hdiag0, cs, a.one, a.one2, a.scores, a.nident,
                                                 a.out_lut ? it.out_base : it.out_base + (qa - it.q_begin), (int)m,
                                                 a.out_lut ? a.out_lut + qa : nullptr, a.Q.off + qa);
                else if constexpr (TAG && !MULTI && TwoRows<K, false>::value)
                    stream_block_tag2<K, false>(a.Q.codes, g0, g1, prof, rsH, rsF, lane, lane_last, slot_last,
                                                hdiag0, cs, a.one, a.one2, a.scores, a.nident,
                                                a.out_lut ? it.out_base : it.out_base + (qa - it.q_begin), (int)m,
                                                a.out_lut ? a.out_lut + qa : nullptr);
                else
                    stream_block<K, false, MULTI, false, false, false, TAG>(
                        a.Q.codes, g0, g1, prof, rsH, rsF, lane, pass == 0, lastp, lastp ? lane_last : 31,
                        slot_last, hdiag0, cs, a.one, MULTI ? scratch + (g0 - x0) : nullptr, a.scores, a.nident,
                        a.out_lut ? it.out_base : it.out_base + (qa - it.q_begin), nullptr, nullptr, nullptr, nullptr,
                        nullptr, 0, a.one2, (int)m, a.out_lut ? a.out_lut + qa : nullptr);
            }
        }
    }
}

// Direction-store variant: items are (template, pair range); each warp takes whole
// pairs.  cshift = 0, so the packed value is score*4 + prio and n_identical comes
// from the traceback kernel instead.
template <int K>
__global__ void __launch_bounds__(kThreads) gotoh_dirs_kernel(const KArgs a) {
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_item;
    constexpr int ROW = KTraits<K>::ROW;
    constexpr int W = KTraits<K>::W;
    uint4* prof = smem;
    uint4* rsH = smem + (size_t)a.C * ROW;
    uint4* rsF = rsH + ROW;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    BSA_TMA_PREAMBLE(K, a.C, a.subst)

    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.item_counter, 1u);
        __syncthreads();
        const uint32_t ii = s_item;
        if (ii >= a.n_items) break;
        const Item it = a.items[ii];   // q_begin/q_end index a.pairs here
        const uint64_t t0 = a.T.off[it.t];
        const uint32_t m = (uint32_t)(a.T.off[it.t + 1] - t0);
        const uint8_t* tc = a.T.codes + t0;
        const Consts cs = make_consts<K>(a.go, a.ge, 0);
        const uint32_t npass = (m + 32 * K - 1) / (32 * K);

        for (uint32_t pass = 0; pass < npass; ++pass) {
            const uint32_t colbase = pass * 32 * K;
            __syncthreads();
            {
                BSA_TMA_PTRS(K, a.C)
                if (threadIdx.x == 0) {
                    const uint32_t nb = slice_bytes(tc, colbase, min(m - colbase, 32u * K));
                    TmaStage::arm(&s_mbar, nb);
                    TmaStage::copy(&s_mbar, s_tcA, slice_src(tc, colbase), nb);
                }
                TmaStage::wait(&s_mbar, s_phase);
                build_profile<K>(prof, rsH, rsF, s_subst, slice_view(s_tcA, tc, colbase), m, colbase, a, cs);
            }
            __syncthreads();
            BSA_TMA_FLIP()
            const bool lastp = (pass + 1 == npass);
            const int lane_last = (int)((m - 1 - colbase) / K);
            const int slot_last = (int)((m - 1 - colbase) % K);
            const long long jl = (long long)colbase + (long long)lane * K;
            const int hdiag0 = jl == 0 ? 0 : (int)((a.go + (jl - 1) * a.ge) * 4);
            for (uint32_t pi = it.q_begin + warp; pi < it.q_end; pi += kWarpsPerCta) {
                const PairRec pr = a.pairs[pi];
                const uint64_t g0 = a.Q.off[pr.q], g1 = a.Q.off[pr.q + 1];
                const uint32_t n = (uint32_t)(g1 - g0);
                // plane of this pass: (n + 32) steps x 32 lanes x W words
                uint32_t* dirs = a.dirs + pr.dir_off + (size_t)pass * (size_t)(n + 32) * 32 * W;
                // the boundary column must survive until this pair's next pass (passes are
                // the outer loop because the CTA shares the profile), so it is per pair
                stream_block<K, true, true>(a.Q.codes, g0, g1, prof, rsH, rsF, lane, pass == 0,
                                            lastp, lastp ? lane_last : 31, slot_last, hdiag0, cs,
                                            a.one, a.scratch + pr.scr_off, a.scores, nullptr, pr.out, dirs);
                __syncwarp();
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// K1 for SHORT templates (<= 16 x 20 columns): two templates of the same columns-per-lane share
// a warp, template A on lanes 0-15 and template B on lanes 16-31, both fed by the same query
// stream.  Each lane then owns twice as many columns as it would with one template per warp, so
// the per-step bookkeeping (residue fetch, profile address, shuffles, border) is paid once per
// 2K cells instead of once per K, and the column padding is cut to a 16-column granularity.
struct KArgsPair {
    SeqStoreDev Q, T;
    SeqStoreDev QA;         // even-aligned copy of the stream store (see KArgs)
    const int16_t* subst;
    const uint8_t* isgap;
    int C, go, ge, one, one2;
    int cs_cap;             // bitlen(longest query): the count field never needs more bits
    const struct Item16* items;
    uint32_t n_items;
    uint32_t* item_counter;
    int32_t* scores;
    uint32_t* nident;
    const uint64_t* out_lut;   // see KArgs
    int flip;
};

// ---------------------------------------------------------------------------------------------
// Score-only, 16-bit packed lanes (one-vs-many, BASELINE configs[3]).  Two TEMPLATES of the
// same column count share a warp: the low and high halves of every 32-bit register carry the
// DP values of template A and template B, both fed by the same query stream, so one
// VIADDMNMX.S16x2 / VIMNMX.S16x2 / VIADD.16x2 advances two cells.  No tie priorities are needed
// for the score alone (SURVEY.md 8a note 5): the cell is 5 instructions per TWO cells
//     t  = max(hdiag + T, e)      VIADDMNMX.U16x2
//     h  = max(t, f)              VIMNMX.U16x2
//     hg = h + GO                 IMAD (FMA pipe: one 32-bit add does both halves, see below)
//     e  = max(e + GE, hg)        VIADDMNMX.U16x2
//     f  = max(f + GE, hg)        VIADDMNMX.U16x2
// The host only routes template pairs here whose scores provably stay inside 16 bits.
// Values are stored BIASED by 0x8000 per half (unsigned order == signed order of the true values).
// The only addition that is not fused into a DPX instruction, hg = h + gap_open, then is a plain
// 32-bit subtraction of |go| * 0x10001: the low half never borrows from the high half because
// every biased value in it is >= |go| (host range check), so it runs on the otherwise idle FMA
// pipe and the ALU pipe carries 4 instead of 5 instructions per two cells.
struct Item16 {
    uint32_t tA, tB;        // tB == 0xffffffff: single template (high half idle)
    uint32_t q_begin, q_end;
    uint64_t outA, outB;    // result index of (q_begin, tA) / (q_begin, tB)
};

__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }
constexpr uint32_t kBias2 = 0x80008000u;
__device__ __forceinline__ uint32_t pack2b(int v) { return pack2(v) ^ kBias2; }   // biased halves
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

template <int K, bool MULTI>
__device__ __forceinline__ void stream_block16(const uint8_t* __restrict__ codes, uint64_t g0, uint64_t g1,
                                               const uint4* prof, const uint4* rsH, const uint4* rsF,
                                               const int lane, const bool first, const bool lastp,
                                               const int lastA, const int slotA, const int lastB,
                                               const int slotB, const uint32_t hdiag0, const uint32_t GE,
                                               const uint32_t GO, const int GO32, const int one, uint2* scratch,
                                               int32_t* __restrict__ scores, uint64_t outA, uint64_t outB) {
    constexpr int ROWB = KTraits<K>::ROW * (int)sizeof(uint4);
    constexpr int U = MULTI ? 2 : 4;
    const uint32_t X = (uint32_t)(g1 - g0);
    const int span = (MULTI && !lastp) ? 31 : (lastA > lastB ? lastA : lastB);
    const uint32_t nsteps = (X + (uint32_t)span + (U - 1)) / U * U;

    int Ha[K], Hb[K], Fr[K], T[K];
    load_vec<K>(Ha, rsH + lane);
    load_vec<K>(Fr, rsF + lane);
    uint32_t hdiag = hdiag0;
    uint32_t hb = GO ^ kBias2;        // H[1][0] = gap_open, both halves (biased)
    uint32_t oh = 0, oe = 0, emitted = 0;
    const bool border = !MULTI || first;
    const bool lane0 = lane == 0;
    const char* prof_lane = reinterpret_cast<const char*>(prof + lane);
    const uint8_t* p = codes + g0 - lane;
    uint32_t b[U], nb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) b[u] = ld_code(p + u);
    uint2 sc_next = make_uint2(0u, 0u);
    uint2 ring_cur = make_uint2(0u, 0u), ring_nxt = make_uint2(0u, 0u);   // see stream_block
    constexpr bool RING = MULTI && BSA_RING;
    if (RING && !first) {
        if ((uint32_t)lane < X) ring_cur = scratch[lane];
        if (32u + (uint32_t)lane < X) ring_nxt = scratch[32 + lane];
    }
    if (MULTI && !RING && !first && lane0 && X > 0) sc_next = scratch[0];

#define BSA_STEP16_CORE(HO, HN, B, S)                                                             \
        load_vec<K>(T, reinterpret_cast<const uint4*>(prof_lane + ((B)&kCodeMask) * ROWB));       \
        uint32_t hin = __shfl_up_sync(0xffffffffu, oh, 1);                                        \
        uint32_t e = __shfl_up_sync(0xffffffffu, oe, 1);                                          \
        if (RING && !first) {                                                                     \
            sc_next.x = __shfl_sync(0xffffffffu, ring_cur.x, (S)&31u);                            \
            sc_next.y = __shfl_sync(0xffffffffu, ring_cur.y, (S)&31u);                            \
            if (((S)&31u) == 31u) {                                                               \
                ring_cur = ring_nxt;                                                              \
                if ((S) + 33u + (uint32_t)lane < X) ring_nxt = scratch[(S) + 33u + lane];         \
            }                                                                                     \
        }                                                                                         \
        if (lane0) {                                                                              \
            if (border) { hin = hb; e = add2(hb, GO); }                                           \
            else { hin = sc_next.x; e = sc_next.y; }                                              \
        }                                                                                         \
        if (MULTI && !RING && !first && lane0 && (S) + 1 < X) sc_next = scratch[(S) + 1];         \
        hb = add2(hb, GE);                                                                        \
        uint32_t hd = hdiag;                                                                      \
        hdiag = hin;                                                                              \
        _Pragma("unroll") for (int c = 0; c < K; ++c) {                                           \
            const uint32_t t = __viaddmax_u16x2(hd, (uint32_t)T[c], e);                           \
            const uint32_t h = __vmaxu2(t, (uint32_t)Fr[c]);                                      \
            const uint32_t hg = (uint32_t)((int)h * one + GO32);                                  \
            e = __viaddmax_u16x2(e, GE, hg);                                                      \
            Fr[c] = (int)__viaddmax_u16x2((uint32_t)Fr[c], GE, hg);                               \
            hd = (uint32_t)HO[c];                                                                 \
            HN[c] = (int)h;                                                                       \
        }                                                                                         \
        oh = (uint32_t)HN[K - 1];                                                                 \
        oe = e;                                                                                   \
        if (MULTI && !lastp && lane == 31) {                                                      \
            const uint32_t pos = (S)-31u;                                                         \
            if (pos < X) scratch[pos] = make_uint2(oh, oe);                                       \
        }
#define BSA_STEP16_FAST(HO, HN, B, S) { BSA_STEP16_CORE(HO, HN, B, S) }
#define BSA_STEP16(HO, HN, B, S)                                                                  \
    {                                                                                             \
        BSA_STEP16_CORE(HO, HN, B, S)                                                             \
        if ((B)&kLastFlag) {                                                                      \
            const uint32_t pos = (S) - (uint32_t)lane;                                            \
            const bool valid = pos < X;                                                           \
            if (lastp && valid && scores) {                                                       \
                if (lane == lastA) {                                                              \
                    int v = 0;                                                                    \
                    _Pragma("unroll") for (int c = 0; c < K; ++c) if (c == slotA) v = HN[c];      \
                    scores[outA + emitted] = (v & 0xffff) - 0x8000;                               \
                }                                                                                 \
                if (lane == lastB) {                                                              \
                    int v = 0;                                                                    \
                    _Pragma("unroll") for (int c = 0; c < K; ++c) if (c == slotB) v = HN[c];      \
                    scores[outB + emitted] = (int)((uint32_t)v >> 16) - 0x8000;                   \
                }                                                                                 \
            }                                                                                     \
            emitted += valid ? 1u : 0u;                                                           \
            load_vec<K>(HN, rsH + lane);                                                          \
            load_vec<K>(Fr, rsF + lane);                                                          \
            hdiag = hdiag0;                                                                       \
            hb = GO ^ kBias2;                                                                     \
        }                                                                                         \
    }

    for (uint32_t s = 0; s < nsteps; s += U) {
        uint32_t any = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            nb[u] = ld_code(p + U + u);
            any |= b[u];
        }
        p += U;
        if (!__any_sync(0xffffffffu, any & kLastFlag)) {
#pragma unroll
            for (int u = 0; u < U; u += 2) {
                BSA_STEP16_FAST(Ha, Hb, b[u], s + u)
                BSA_STEP16_FAST(Hb, Ha, b[u + 1], s + u + 1)
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; u += 2) {
                BSA_STEP16(Ha, Hb, b[u], s + u)
                BSA_STEP16(Hb, Ha, b[u + 1], s + u + 1)
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) b[u] = nb[u];
    }
#undef BSA_STEP16
#undef BSA_STEP16_FAST
#undef BSA_STEP16_CORE
}

// Two rows per step for the 16-bit score-only lanes (one column block): same idea as
// stream_block_tag2 -- lanes skewed by two stream positions, cell (r+1, c) right after cell (r, c),
// H in place.  The 16-bit cell is a chain of four dependent instructions per two cells and only
// 2 ALU-pipe instructions per cell, so with one row in flight the warp mostly waits on itself.
template <int K>
__device__ __forceinline__ void stream_block16_2r(const uint8_t* __restrict__ codes, uint64_t g0, uint64_t g1,
                                                  const uint4* prof, const uint4* rsH, const uint4* rsF,
                                                  const int lane, const int lastA, const int slotA,
                                                  const int lastB, const int slotB, const uint32_t hdiag0,
                                                  const uint32_t GE, const uint32_t GO, const int GO32,
                                                  const int one, int32_t* __restrict__ scores, uint64_t outA,
                                                  uint64_t outB) {
    constexpr int ROWB = KTraits<K>::ROW * (int)sizeof(uint4);
    constexpr int U = BSA_TAG2_U;
    const uint32_t X = (uint32_t)(g1 - g0);
    const int span = lastA > lastB ? lastA : lastB;
    const uint32_t nd = ((X + 1u) / 2u + (uint32_t)span + (U - 1)) / U * U;
    const bool lane0 = lane == 0;

    int H[K], Fr[K], T0[K], T1[K];
    load_vec<K>(H, rsH + lane);
    load_vec<K>(Fr, rsF + lane);
    uint32_t hdiag = hdiag0, hb = GO ^ kBias2;
    uint32_t oh0 = 0, oe0 = 0, oh1 = 0, oe1 = 0, emitted = 0;
    const char* prof_lane = reinterpret_cast<const char*>(prof + lane);
    const uint8_t* p = codes + g0 - 2 * lane;
    uint32_t b[2 * U], nb[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) b[u] = ld_code(p + u);

#define BSA_ROW16(TT, HIN, E, OH, OE)                                                             \
    {                                                                                             \
        if (lane0) { HIN = hb; E = add2(hb, GO); }                                                \
        hb = add2(hb, GE);                                                                        \
        uint32_t hd = hdiag;                                                                      \
        hdiag = HIN;                                                                              \
        _Pragma("unroll") for (int c = 0; c < K; ++c) {                                           \
            const uint32_t t = __viaddmax_u16x2(hd, (uint32_t)TT[c], E);                          \
            const uint32_t h = __vmaxu2(t, (uint32_t)Fr[c]);                                      \
            const uint32_t hg = (uint32_t)((int)h * one + GO32);                                  \
            E = __viaddmax_u16x2(E, GE, hg);                                                      \
            Fr[c] = (int)__viaddmax_u16x2((uint32_t)Fr[c], GE, hg);                               \
            hd = (uint32_t)H[c];                                                                  \
            H[c] = (int)h;                                                                        \
        }                                                                                         \
        OH = (uint32_t)H[K - 1];                                                                  \
        OE = E;                                                                                   \
    }
#define BSA_FLAG16(B, POS)                                                                        \
    if ((B)&kLastFlag) {                                                                          \
        const bool valid = (POS) < X;                                                             \
        if (valid && scores) {                                                                    \
            if (lane == lastA) {                                                                  \
                int v = 0;                                                                        \
                _Pragma("unroll") for (int c = 0; c < K; ++c) if (c == slotA) v = H[c];           \
                scores[outA + emitted] = (v & 0xffff) - 0x8000;                                   \
            }                                                                                     \
            if (lane == lastB) {                                                                  \
                int v = 0;                                                                        \
                _Pragma("unroll") for (int c = 0; c < K; ++c) if (c == slotB) v = H[c];           \
                scores[outB + emitted] = (int)((uint32_t)v >> 16) - 0x8000;                       \
            }                                                                                     \
        }                                                                                         \
        emitted += valid ? 1u : 0u;                                                               \
        load_vec<K>(H, rsH + lane);                                                               \
        load_vec<K>(Fr, rsF + lane);                                                              \
        hdiag = hdiag0;                                                                           \
        hb = GO ^ kBias2;                                                                         \
    }

    // two rows interleaved (needs: no lane ends a sequence on the FIRST of the two rows)
#define BSA_PAIR16()                                                                              \
    {                                                                                             \
        if (lane0) {                                                                              \
            hin0 = hb;                                                                            \
            e0 = add2(hb, GO);                                                                    \
            hin1 = add2(hb, GE);                                                                  \
            e1 = add2(hin1, GO);                                                                  \
        }                                                                                         \
        hb = add2(add2(hb, GE), GE);                                                              \
        uint32_t hd0 = hdiag, hd1 = hin0, h1 = 0;                                                 \
        hdiag = hin1;                                                                             \
_Pragma("unroll")                                                                                 \
        for (int c = 0; c < K; ++c) {                                                             \
            const uint32_t t0 = __viaddmax_u16x2(hd0, (uint32_t)T0[c], e0);                       \
            const uint32_t h0 = __vmaxu2(t0, (uint32_t)Fr[c]);                                    \
            const uint32_t hg0 = (uint32_t)((int)h0 * one + GO32);                                \
            e0 = __viaddmax_u16x2(e0, GE, hg0);                                                   \
            const uint32_t f1 = __viaddmax_u16x2((uint32_t)Fr[c], GE, hg0);                       \
            hd0 = (uint32_t)H[c];                                                                 \
            const uint32_t t1 = __viaddmax_u16x2(hd1, (uint32_t)T1[c], e1);                       \
            h1 = __vmaxu2(t1, f1);                                                                \
            const uint32_t hg1 = (uint32_t)((int)h1 * one + GO32);                                \
            e1 = __viaddmax_u16x2(e1, GE, hg1);                                                   \
            Fr[c] = (int)__viaddmax_u16x2(f1, GE, hg1);                                           \
            hd1 = h0;                                                                             \
            H[c] = (int)h1;                                                                       \
        }                                                                                         \
        oh0 = hd1;                                                                                \
        oe0 = e0;                                                                                 \
        oh1 = h1;                                                                                 \
        oe1 = e1;                                                                                 \
    }

    for (uint32_t S = 0; S < nd; S += U) {
        uint32_t any = 0;
#pragma unroll
        for (int u = 0; u < 2 * U; ++u) {
            nb[u] = ld_code(p + 2 * U + u);
            any |= b[u];
        }
        p += 2 * U;
        if (!__any_sync(0xffffffffu, any & kLastFlag)) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                load_vec<K>(T0, reinterpret_cast<const uint4*>(prof_lane + (b[2 * u] & kCodeMask) * ROWB));
                load_vec<K>(T1, reinterpret_cast<const uint4*>(prof_lane + (b[2 * u + 1] & kCodeMask) * ROWB));
                uint32_t hin0 = __shfl_up_sync(0xffffffffu, oh0, 1);
                uint32_t e0 = __shfl_up_sync(0xffffffffu, oe0, 1);
                uint32_t hin1 = __shfl_up_sync(0xffffffffu, oh1, 1);
                uint32_t e1 = __shfl_up_sync(0xffffffffu, oe1, 1);
                BSA_PAIR16()
            }
        } else {
            BSA_FLAG_UNROLL
            for (int u = 0; u < U; ++u) {
                const uint32_t pos0 = 2u * (S + u - (uint32_t)lane);
                load_vec<K>(T0, reinterpret_cast<const uint4*>(prof_lane + (BSA_B(u, 0) & kCodeMask) * ROWB));
                load_vec<K>(T1, reinterpret_cast<const uint4*>(prof_lane + (BSA_B(u, 1) & kCodeMask) * ROWB));
                uint32_t hin0 = __shfl_up_sync(0xffffffffu, oh0, 1);
                uint32_t e0 = __shfl_up_sync(0xffffffffu, oe0, 1);
                uint32_t hin1 = __shfl_up_sync(0xffffffffu, oh1, 1);
                uint32_t e1 = __shfl_up_sync(0xffffffffu, oe1, 1);
                if (BSA_FLAG_SPLIT && !__any_sync(0xffffffffu, BSA_B(u, 0) & kLastFlag)) {
                    BSA_PAIR16()
                    BSA_FLAG16(BSA_B(u, 1), pos0 + 1u)
                } else {
                    BSA_ROW16(T0, hin0, e0, oh0, oe0)
                    BSA_FLAG16(BSA_B(u, 0), pos0)
                    BSA_ROW16(T1, hin1, e1, oh1, oe1)
                    BSA_FLAG16(BSA_B(u, 1), pos0 + 1u)
                }
                if (BSA_FLAG_LOOP) {
#pragma unroll
                    for (int k = 0; k + 2 < 2 * U; ++k) b[k] = b[k + 2];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2 * U; ++u) b[u] = nb[u];
    }
#undef BSA_ROW16
#undef BSA_FLAG16
#undef BSA_PAIR16
}

// ---------------------------------------------------------------------------------------------
// Score only, 16-bit packed lanes, in the MOVING FRAME over the EVEN-ALIGNED stream (BSA_ALIGNED;
// single-block templates).  Every stored value of cell (i, j) is score - (i + j) * gap_extend, so
// both gap extensions are free and the cell is FOUR instructions per TWO cells, all of them DPX:
//     t = max(hdiag + T, e)            VIADDMNMX.U16x2     (T = s - 2 ge)
//     h = max(t, f)                    VIMNMX.U16x2
//     e = max(h + (go - ge), e)        VIADDMNMX.U16x2
//     f = max(h + (go - ge), f)        VIADDMNMX.U16x2
// The borders are the constants go - ge (H*) and 2 (go - ge) (the eager E* / F*), so there is no
// border arithmetic per row and the reset of a lane that passes the end of a sequence is a set of
// register moves.  Odd-length sequences are preceded by a PAD row (stream_block_tag2a explains the
// aligned stream): its profile row is 0 (go - ge in DP column 1) and reproduces the top border, so
// the end-of-sequence flag always sits on the second row of a double step and the interleaved
// two-row body is the only form of the cell code.  Values stay biased by 0x8000 per half.
// The score leaves the frame when it is emitted: H = H* + (n + m) ge with the sequence's real length.
template <int K, bool HALF = false>
__device__ __forceinline__ void stream_block16_fa(const uint8_t* __restrict__ codes, uint64_t g0, uint64_t g1,
                                                  const uint4* prof, const int lane, const int lastA,
                                                  const int slotA, const int lastB, const int slotB,
                                                  const uint32_t hdiag0, const uint32_t HB, const uint32_t GOF2,
                                                  const int ge, const int mA, const int mB,
                                                  int32_t* __restrict__ scores, uint64_t outA, uint64_t outB,
                                                  const uint64_t* __restrict__ qoffp) {
    constexpr int ROWB = KTraits<K>::ROW * (int)sizeof(uint4);
    constexpr int U = K <= BSA_TAG2_U_SMALLK ? 8 : BSA_TAG2_U;
    const uint32_t X = (uint32_t)(g1 - g0);        // even
    // HALF (gotoh_score16_quad_kernel): lanes 0-15 and 16-31 are two independent 16-lane pipelines (four templates
    // per warp); lastA/lastB, the template lengths and the output bases are then per-lane values
    const int span = HALF ? 15 : (lastA > lastB ? lastA : lastB);
    const uint32_t nd = (X / 2u + (uint32_t)span + (U - 1)) / U * U;
    const int lrel = HALF ? (lane & 15) : lane;
    const bool lane0 = lrel == 0;
    const uint32_t FB = add2(HB, GOF2);            // eager E*[i][1] and F*[1][j]: opened from the border
    // the constant left border enters lane 0 through one LOP3 per value: (shuffled & keep) | border
    const uint32_t keep = (uint32_t)opaque_reg(lane0 ? 0 : -1);
    const uint32_t hbl = (uint32_t)opaque_reg(lane0 ? (int)HB : 0);
    const uint32_t zbl = (uint32_t)opaque_reg(lane0 ? (int)kBias2 : 0);   // H*[0][0] = 0: the row above is the PAD row
    const uint32_t ebl = (uint32_t)opaque_reg(lane0 ? (int)FB : 0);

    uint32_t H[K], Fr[K];
    int T0[K], T1[K];
#pragma unroll
    for (int c = 0; c < K; ++c) { H[c] = HB; Fr[c] = FB; }
    uint32_t hdiag = hdiag0;
    uint32_t oh0 = 0, oe0 = 0, oh1 = 0, oe1 = 0, emitted = 0;
    const char* prof_lane = reinterpret_cast<const char*>(prof + lane);
    const uint8_t* p = codes + g0 - 2 * lrel;
    uint32_t b[2 * U], nb[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) b[u] = ld_code(p + u);

    for (uint32_t S = 0; S < nd; S += U) {
#pragma unroll
        for (int u = 0; u < 2 * U; ++u) nb[u] = ld_code(p + 2 * U + u);
        p += 2 * U;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t b0 = b[2 * u], b1 = b[2 * u + 1];     // b0 never carries a flag
            int hq = 0;
            if constexpr (BSA_PADSLOT && K % 4 != 0)
                load_vec_x<K>(T0, hq, reinterpret_cast<const uint4*>(prof_lane + b0 * ROWB));
            else {
                load_vec<K>(T0, reinterpret_cast<const uint4*>(prof_lane + b0 * ROWB));
                hq = (int)(b0 == kPadCode ? zbl : hbl);
            }
            load_vec<K>(T1, reinterpret_cast<const uint4*>(prof_lane + (b1 & kCodeMask) * ROWB));
            const uint32_t sh0 = __shfl_up_sync(0xffffffffu, oh0, 1);
            const uint32_t se0 = __shfl_up_sync(0xffffffffu, oe0, 1);
            const uint32_t sh1 = __shfl_up_sync(0xffffffffu, oh1, 1);
            const uint32_t se1 = __shfl_up_sync(0xffffffffu, oe1, 1);
            const uint32_t hin0 = (sh0 & keep) | (uint32_t)hq;      // hq: lane 0's border of row 0 (0 for the PAD row)
            const uint32_t hin1 = (sh1 & keep) | hbl;
            uint32_t e0 = (se0 & keep) | ebl;
            uint32_t e1 = (se1 & keep) | ebl;
            uint32_t hd0 = hdiag, hd1 = hin0, h1 = 0;
            hdiag = hin1;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                const uint32_t t0 = __viaddmax_u16x2(hd0, (uint32_t)T0[c], e0);
                const uint32_t h0 = __vmaxu2(t0, Fr[c]);
                e0 = __viaddmax_u16x2(h0, GOF2, e0);
                const uint32_t f1 = __viaddmax_u16x2(h0, GOF2, Fr[c]);
                hd0 = H[c];
                const uint32_t t1 = __viaddmax_u16x2(hd1, (uint32_t)T1[c], e1);
                h1 = __vmaxu2(t1, f1);
                e1 = __viaddmax_u16x2(h1, GOF2, e1);
                Fr[c] = __viaddmax_u16x2(h1, GOF2, f1);
                hd1 = h0;
                H[c] = h1;
            }
            oh0 = hd1;
            oe0 = e0;
            oh1 = h1;
            oe1 = e1;
            if (b1 & kLastFlag) {
                const uint32_t pos1 = 2u * (S + u - (uint32_t)lrel) + 1u;   // wraps below row 0: fails pos1 < X
                const bool valid = pos1 < X;
                if (valid && scores && (lane == lastA || lane == lastB)) {
                    const int n = (int)(qoffp[emitted + 1] - qoffp[emitted]);
                    if (lane == lastA) {
                        uint32_t v = 0;
#pragma unroll
                        for (int c = 0; c < K; ++c) if (c == slotA) v = H[c];
                        scores[outA + emitted] = (int)(v & 0xffffu) - 0x8000 + (n + mA) * ge;
                    }
                    if (lane == lastB) {
                        uint32_t v = 0;
#pragma unroll
                        for (int c = 0; c < K; ++c) if (c == slotB) v = H[c];
                        scores[outB + emitted] = (int)(v >> 16) - 0x8000 + (n + mB) * ge;
                    }
                }
                emitted += valid ? 1u : 0u;
                const uint32_t hbr = (uint32_t)opaque_reg((int)HB), fbr = (uint32_t)opaque_reg((int)FB);
#pragma unroll
                for (int c = 0; c < K; ++c) { H[c] = hbr; Fr[c] = fbr; }
                hdiag = hdiag0;
            }
        }
#pragma unroll
        for (int u = 0; u < 2 * U; ++u) b[u] = nb[u];
    }
}

template <int K, bool TAG = false>
__global__ void __launch_bounds__(kThreads, MinBlocks<K>::value) gotoh_pair_kernel(const KArgsPair a) {
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_item;
    __shared__ uint32_t s_chunk;
    constexpr int ROW = KTraits<K>::ROW;
    uint4* prof = smem;
    uint4* rsH = smem + (size_t)a.C * ROW;
    uint4* rsF = rsH + ROW;
    const int lane = threadIdx.x & 31;
    const int lrel = lane & 15;
    const bool isB = lane >= 16;
    BSA_TMA_PREAMBLE(K, a.C, a.subst)

    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.item_counter, 1u);
        __syncthreads();
        const uint32_t ii = s_item;
        if (ii >= a.n_items) break;
        const Item16 it = a.items[ii];
        const bool hasB = it.tB != 0xffffffffu;
        const uint64_t a0 = a.T.off[it.tA];
        const uint32_t mA = (uint32_t)(a.T.off[it.tA + 1] - a0);
        const uint64_t b0 = hasB ? a.T.off[it.tB] : 0;
        const uint32_t mB = hasB ? (uint32_t)(a.T.off[it.tB + 1] - b0) : 0u;
        const uint32_t mmax = mA > mB ? mA : mB;
        int cshift = 32 - __clz((int)mmax);
        if (BSA_ALIGNED && BSA_ETAG && TAG && TwoRows<K, true>::value) cshift = 32 - __clz(16 * K);   // launch constant
        cshift = cshift < a.cs_cap ? cshift : a.cs_cap;
        const Consts cs = make_consts<K>(a.go, a.ge, cshift, false, TAG, a.flip != 0);
        const int S = 1 << cs.sh, P3 = 3 << cs.ps;

        constexpr bool kAligned = BSA_ALIGNED && TAG && TwoRows<K, true>::value;
        const uint64_t* __restrict__ xoff = kAligned ? a.QA.off : a.Q.off;
        const uint64_t x0 = xoff[it.q_begin], x1 = xoff[it.q_end];
        const uint64_t span = x1 - x0;
        const ChunkPlan cp = plan_chunks(span);
        const uint64_t head = cp.head;
        const uint32_t nbig = cp.nbig, nsmall = cp.nsmall;
        const uint32_t nch = nbig + nsmall;

        __syncthreads();
        BSA_TMA_PTRS(K, a.C)
        if (threadIdx.x == 0) {
            s_chunk = 0;
            const uint32_t nbA = slice_bytes(a.T.codes + a0, 0, mA);
            const uint32_t nbB = hasB ? slice_bytes(a.T.codes + b0, 0, mB) : 0u;
            TmaStage::arm(&s_mbar, nbA + nbB);
            TmaStage::copy(&s_mbar, s_tcA, slice_src(a.T.codes + a0, 0), nbA);
            if (hasB) TmaStage::copy(&s_mbar, s_tcB, slice_src(a.T.codes + b0, 0), nbB);
        }
        TmaStage::wait(&s_mbar, s_phase);
        const uint8_t* vA = slice_view(s_tcA, a.T.codes + a0, 0);
        const uint8_t* vB = slice_view(s_tcB, a.T.codes + b0, 0);
        // profile: lanes 0-15 hold the columns of template A, lanes 16-31 those of template B
        for (int idx = threadIdx.x; idx < a.C * ROW; idx += blockDim.x) {
            const int code = idx / ROW, r = idx - code * ROW, v = r >> 5, ln = r & 31;
            const bool b = ln >= 16;
            const uint8_t* tc = b ? vB : vA;
            const uint32_t m = b ? mB : mA;
            const bool gap = a.isgap[code] != 0;
            int o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * v + e;
                const uint32_t col = (uint32_t)(ln & 15) * K + c;
                int val = cs.T_PAD;
                if (BSA_PADSLOT && TAG && K % 4 != 0 && c == K)      // spare slot: the left-border H* of lanes 0 and 16
                    val = (ln & 15) == 0 ? (code == (int)kPadCode ? 0 : cs.hb0) : 0;
                if (c < K && col < m) {
                    const int tcode = tc[col] & kCodeMask;
                    val = ((int)s_subst[a.flip ? tcode * a.C + code : code * a.C + tcode] + cs.TSUB) * S + P3 +
                          ((tcode == code && !gap) ? 1 : 0);
                    if (TAG && code == (int)kPadCode) val = (col == 0 ? cs.hb0 : 0) + P3;   // PAD row, see build_profile
                }
                o[e] = val;
            }
            prof[idx] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        for (int r = threadIdx.x; r < ROW; r += blockDim.x) {
            const int v = r >> 5, ln = r & 31;
            int h[4], f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long j = (long long)(ln & 15) * K + 4 * v + e + 1;
                h[e] = TAG ? cs.hb0 : (int)((a.go + (j - 1) * a.ge) * S);   // TAG: constant borders of the frame
                f[e] = h[e] + cs.GOF + cs.XTOP;
            }
            rsH[r] = make_uint4(h[0], h[1], h[2], h[3]);
            rsF[r] = make_uint4(f[0], f[1], f[2], f[3]);
        }
        __syncthreads();
        BSA_TMA_FLIP()
        const uint32_t mine = isB ? mB : mA;
        // a lane emits only for its own template; an absent template B never matches a lane
        const int my_last = (isB && !hasB) ? -1 : (int)((mine - 1) / K) + (isB ? 16 : 0);
        const int my_slot = mine ? (int)((mine - 1) % K) : 0;
        const long long jl = (long long)lrel * K;
        const int hdiag0 = jl == 0 ? 0 : (TAG ? cs.hb0 : (int)((a.go + (jl - 1) * a.ge) * S));
        for (;;) {
            uint32_t c = 0;
            if (lane == 0) c = atomicAdd(&s_chunk, 1u);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= nch) break;
            const uint64_t ca = c <= nbig ? head * c / nbig : head + (span - head) * (c - nbig) / nsmall;
            const uint64_t cb = c + 1 <= nbig ? head * (c + 1) / nbig : head + (span - head) * (c + 1 - nbig) / nsmall;
            const uint32_t qa = c == 0 ? it.q_begin : lower_bound_off(xoff, it.q_begin, it.q_end, x0 + ca);
            const uint32_t qb = c + 1 == nch ? it.q_end : lower_bound_off(xoff, it.q_begin, it.q_end, x0 + cb);
            const uint64_t g0 = xoff[qa], g1 = xoff[qb];
            if (g1 <= g0) continue;
            if constexpr (kAligned)
                stream_block_tag2a<K, true>(a.QA.codes, g0, g1, prof, rsH, rsF, lane, my_last, my_slot, hdiag0, cs,
                                            a.one, a.one2, a.scores, a.nident,
                                            (isB ? it.outB : it.outA) + (a.out_lut ? 0 : qa - it.q_begin), (int)mine,
                                            a.out_lut ? a.out_lut + qa : nullptr, a.Q.off + qa);
            else if constexpr (TAG && TwoRows<K, true>::value)
                stream_block_tag2<K, true>(a.Q.codes, g0, g1, prof, rsH, rsF, lane, my_last, my_slot, hdiag0, cs,
                                           a.one, a.one2, a.scores, a.nident,
                                           (isB ? it.outB : it.outA) + (a.out_lut ? 0 : qa - it.q_begin), (int)mine,
                                           a.out_lut ? a.out_lut + qa : nullptr);
            else
                stream_block<K, false, false, false, false, true, TAG>(
                    a.Q.codes, g0, g1, prof, rsH, rsF, lane, true, true, my_last, my_slot, hdiag0, cs, a.one, nullptr,
                    a.scores, a.nident, (isB ? it.outB : it.outA) + (a.out_lut ? 0 : qa - it.q_begin), nullptr, nullptr,
                    nullptr, nullptr, nullptr, 0, a.one2, (int)mine, a.out_lut ? a.out_lut + qa : nullptr);
        }
    }
}

struct KArgs16 {
    SeqStoreDev Q, T;
    SeqStoreDev QA;         // even-aligned copy of the stream store (single-block kernels, stream_block16_fa)
    const int16_t* subst;
    int C, go, ge;
    const Item16* items;
    uint32_t n_items;
    uint32_t* item_counter;
    int32_t* scores;
    uint2* scratch;
    uint32_t scratch_stride;
    int one;                // runtime 1: keeps hg = h + GO an IMAD
};

// the two-row step keeps two chains in flight per warp, so the single-block kernels trade the third
// resident CTA for a spill-free 128-register budget from K = 10 on
template <int K, bool MULTI>
struct MinBlocks16 { static constexpr int value = (!MULTI && K >= 10) ? 2 : MinBlocks<K>::value; };

template <int K, bool MULTI>
__global__ void __launch_bounds__(kThreads, MinBlocks16<K, MULTI>::value) gotoh_score16_kernel(const KArgs16 a) {
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_item;
    __shared__ uint32_t s_chunk;
    constexpr int ROW = KTraits<K>::ROW;
    uint4* prof = smem;
    uint4* rsH = smem + (size_t)a.C * ROW;
    uint4* rsF = rsH + ROW;
    const int lane = threadIdx.x & 31;
    const uint32_t GE = pack2(a.ge), GO = pack2(a.go);
    const int GO32 = a.go * 0x10001;   // h + GO32 subtracts |go| from both (biased) halves, no borrow
    BSA_TMA_PREAMBLE(K, a.C, a.subst)

    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.item_counter, 1u);
        __syncthreads();
        const uint32_t ii = s_item;
        if (ii >= a.n_items) break;
        const Item16 it = a.items[ii];
        const bool hasB = it.tB != 0xffffffffu;
        const uint64_t a0 = a.T.off[it.tA];
        const uint32_t mA = (uint32_t)(a.T.off[it.tA + 1] - a0);
        const uint64_t b0 = hasB ? a.T.off[it.tB] : 0;
        const uint32_t mB = hasB ? (uint32_t)(a.T.off[it.tB + 1] - b0) : 0u;
        const uint8_t* tcA = a.T.codes + a0;
        const uint8_t* tcB = a.T.codes + b0;
        const uint32_t mmax = mA > mB ? mA : mB;

        // single-block templates: moving frame over the even-aligned stream (stream_block16_fa)
        constexpr bool kFrame = BSA_ALIGNED && !MULTI;
        const uint64_t* __restrict__ xoff = kFrame ? a.QA.off : a.Q.off;
        const uint64_t x0 = xoff[it.q_begin], x1 = xoff[it.q_end];
        const uint64_t span = x1 - x0;
        const ChunkPlan cp = plan_chunks(span);
        const uint64_t head = cp.head;
        const uint32_t nbig = cp.nbig, nsmall = cp.nsmall;
        const uint32_t nch = nbig + nsmall;
        const uint32_t npass = MULTI ? (mmax + 32 * K - 1) / (32 * K) : 1u;
        uint2* scratch = MULTI ? a.scratch + (size_t)blockIdx.x * a.scratch_stride : nullptr;

        for (uint32_t pass = 0; pass < npass; ++pass) {
            const uint32_t colbase = pass * 32 * K;
            __syncthreads();
            BSA_TMA_PTRS(K, a.C)
            if (threadIdx.x == 0) {
                s_chunk = 0;
                // a template that ends before this block contributes no columns (zero-length copy is skipped)
                const uint32_t ncA = mA > colbase ? min(mA - colbase, 32u * K) : 0u;
                const uint32_t ncB = mB > colbase ? min(mB - colbase, 32u * K) : 0u;
                const uint32_t nbA = ncA ? slice_bytes(tcA, colbase, ncA) : 0u;
                const uint32_t nbB = ncB ? slice_bytes(tcB, colbase, ncB) : 0u;
                TmaStage::arm(&s_mbar, nbA + nbB);
                if (ncA) TmaStage::copy(&s_mbar, s_tcA, slice_src(tcA, colbase), nbA);
                if (ncB) TmaStage::copy(&s_mbar, s_tcB, slice_src(tcB, colbase), nbB);
            }
            TmaStage::wait(&s_mbar, s_phase);
            const uint8_t* vA = slice_view(s_tcA, tcA, colbase);
            const uint8_t* vB = slice_view(s_tcB, tcB, colbase);
            // packed profile: low half template A, high half template B
            for (int idx = threadIdx.x; idx < a.C * ROW; idx += blockDim.x) {
                const int code = idx / ROW, r = idx - code * ROW, v = r >> 5, ln = r & 31;
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * v + e;
                    const uint32_t col = colbase + ln * K + c;
                    int sa = (c < K && col < mA) ? (int)s_subst[code * a.C + (vA[col] & kCodeMask)] : 0;
                    int sb = (c < K && col < mB) ? (int)s_subst[code * a.C + (vB[col] & kCodeMask)] : 0;
                    if (kFrame) {
                        // frame: the diagonal adds s - 2 ge; the PAD row adds 0 (go - ge in DP column 1)
                        const int padv = col == 0 ? a.go - a.ge : 0;
                        if (c < K && col < mA) sa = code == (int)kPadCode ? padv : sa - 2 * a.ge;
                        if (c < K && col < mB) sb = code == (int)kPadCode ? padv : sb - 2 * a.ge;
                    }
                    o[e] = ((uint32_t)sa & 0xffffu) | ((uint32_t)sb << 16);
                    // spare slot behind the K columns: lane 0's left-border H* of a row with this residue (both halves)
                    if (BSA_PADSLOT && kFrame && K % 4 != 0 && c == K)
                        o[e] = ln == 0 ? (code == (int)kPadCode ? kBias2 : pack2b(a.go - a.ge)) : 0u;
                }
                prof[idx] = make_uint4(o[0], o[1], o[2], o[3]);
            }
            for (int r = threadIdx.x; r < ROW; r += blockDim.x) {
                const int v = r >> 5, ln = r & 31;
                uint32_t h[4], f[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const long long j = (long long)colbase + ln * K + 4 * v + e + 1;
                    const int hv = (int)(a.go + (j - 1) * a.ge);
                    h[e] = pack2b(hv);
                    f[e] = pack2b(hv + a.go);
                }
                rsH[r] = make_uint4(h[0], h[1], h[2], h[3]);
                rsF[r] = make_uint4(f[0], f[1], f[2], f[3]);
            }
            __syncthreads();
            BSA_TMA_FLIP()
            const bool lastp = (pass + 1 == npass);
            // a template that ends in an earlier block never emits from this kernel: the host only
            // pairs templates with the same number of blocks
            const int lastA = (int)((mA - 1 - colbase) / K), slotA = (int)((mA - 1 - colbase) % K);
            const int lastB = hasB ? (int)((mB - 1 - colbase) / K) : -1;
            const int slotB = hasB ? (int)((mB - 1 - colbase) % K) : 0;
            const long long jl = (long long)colbase + (long long)lane * K;
            const uint32_t hdiag0 = jl == 0 ? kBias2
                                            : pack2b(kFrame ? a.go - a.ge : (int)(a.go + (jl - 1) * a.ge));
            for (;;) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(&s_chunk, 1u);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= nch) break;
                const uint64_t ca = c <= nbig ? head * c / nbig : head + (span - head) * (c - nbig) / nsmall;
                const uint64_t cb = c + 1 <= nbig ? head * (c + 1) / nbig
                                                  : head + (span - head) * (c + 1 - nbig) / nsmall;
                const uint32_t qa = c == 0 ? it.q_begin : lower_bound_off(xoff, it.q_begin, it.q_end, x0 + ca);
                const uint32_t qb = c + 1 == nch ? it.q_end
                                                 : lower_bound_off(xoff, it.q_begin, it.q_end, x0 + cb);
                const uint64_t g0 = xoff[qa], g1 = xoff[qb];
                if (g1 <= g0) continue;
                if constexpr (kFrame)
                    stream_block16_fa<K>(a.QA.codes, g0, g1, prof, lane, lastA, slotA, lastB, slotB, hdiag0,
                                         pack2b(a.go - a.ge), pack2(a.go - a.ge), a.ge, (int)mA, (int)mB, a.scores,
                                         it.outA + (qa - it.q_begin), it.outB + (qa - it.q_begin), a.Q.off + qa);
                else if constexpr (!MULTI)
                    stream_block16_2r<K>(a.Q.codes, g0, g1, prof, rsH, rsF, lane, lastA, slotA, lastB, slotB, hdiag0,
                                         GE, GO, GO32, a.one, a.scores, it.outA + (qa - it.q_begin),
                                         it.outB + (qa - it.q_begin));
                else
                    stream_block16<K, MULTI>(a.Q.codes, g0, g1, prof, rsH, rsF, lane, pass == 0, lastp,
                                             lastp ? lastA : 31, slotA, lastp ? lastB : 31, slotB, hdiag0, GE, GO,
                                             GO32, a.one, MULTI ? scratch + (g0 - x0) : nullptr, a.scores,
                                             it.outA + (qa - it.q_begin), it.outB + (qa - it.q_begin));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Score only, FOUR short templates per warp (templates of at most 16 K <= 320 columns): lanes 0-15 carry
// templates A (low halves) and B (high halves), lanes 16-31 templates C and D, all four fed by the same
// stream (stream_block16_fa<K, HALF>: two independent 16-lane pipelines, lane 16 takes the left border
// instead of lane 15's shuffle).  Each lane then owns twice as many columns as with two templates per warp,
// so the per-step bookkeeping is paid once per 4 K cells, the column padding drops to 16 columns and the
// pipeline fill to 15 double steps.  Moving frame, even-aligned stream, PAD rows as in the two-template kernel.
struct Item16Q {
    uint32_t t[4];          // 0xffffffff: no template in this slot (A is always present)
    uint32_t q_begin, q_end;
    uint64_t out[4];        // result index of (q_begin, t[i])
};
struct KArgs16Q {
    SeqStoreDev Q, T, QA;
    const int16_t* subst;
    int C, go, ge;
    const Item16Q* items;
    uint32_t n_items;
    uint32_t* item_counter;
    int32_t* scores;
};

template <int K>
__global__ void __launch_bounds__(kThreads, MinBlocks16<K, false>::value) gotoh_score16_quad_kernel(const KArgs16Q a) {
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_item;
    __shared__ uint32_t s_chunk;
    constexpr int ROW = KTraits<K>::ROW;
    uint4* prof = smem;
    uint4* rsH = smem + (size_t)a.C * ROW;      // (border vectors unused: the frame's borders are constants)
    uint4* rsF = rsH + ROW;
    const int lane = threadIdx.x & 31;
    const int lrel = lane & 15;
    const int half = lane >> 4;
    BSA_TMA_PREAMBLE(K, a.C, a.subst)

    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.item_counter, 1u);
        __syncthreads();
        const uint32_t ii = s_item;
        if (ii >= a.n_items) break;
        const Item16Q it = a.items[ii];
        uint64_t t0[4];
        uint32_t m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool has = it.t[i] != 0xffffffffu;
            t0[i] = has ? a.T.off[it.t[i]] : 0;
            m[i] = has ? (uint32_t)(a.T.off[it.t[i] + 1] - t0[i]) : 0u;
        }
        const uint64_t* __restrict__ xoff = a.QA.off;
        const uint64_t x0 = xoff[it.q_begin], x1 = xoff[it.q_end];
        const uint64_t span = x1 - x0;
        const ChunkPlan cp = plan_chunks(span);
        const uint64_t head = cp.head;
        const uint32_t nbig = cp.nbig, nsmall = cp.nsmall;
        const uint32_t nch = nbig + nsmall;

        __syncthreads();
        BSA_TMA_PTRS(K, a.C)
        uint8_t* s_tc[4] = {s_tcA, s_tcA + (16 * K + 32), s_tcA + 2 * (16 * K + 32), s_tcA + 3 * (16 * K + 32)};
        (void)s_tcB;
        if (threadIdx.x == 0) {
            s_chunk = 0;
            uint32_t nb[4], tot = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { nb[i] = m[i] ? slice_bytes(a.T.codes + t0[i], 0, m[i]) : 0u; tot += nb[i]; }
            TmaStage::arm(&s_mbar, tot);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (nb[i]) TmaStage::copy(&s_mbar, s_tc[i], slice_src(a.T.codes + t0[i], 0), nb[i]);
        }
        TmaStage::wait(&s_mbar, s_phase);
        const uint8_t* v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = slice_view(s_tc[i], a.T.codes + t0[i], 0);
        // packed profile: lanes 0-15 templates A | B << 16, lanes 16-31 templates C | D << 16; frame values
        for (int idx = threadIdx.x; idx < a.C * ROW; idx += blockDim.x) {
            const int code = idx / ROW, r = idx - code * ROW, vv = r >> 5, ln = r & 31;
            const int h2 = (ln >> 4) * 2;
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * vv + e;
                const uint32_t col = (uint32_t)(ln & 15) * K + c;
                const int padv = col == 0 ? a.go - a.ge : 0;
                int sv[2];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int ti = h2 + hh;
                    int x = 0;
                    if (c < K && col < m[ti]) {
                        const int tcode = v[ti][col] & kCodeMask;
                        x = code == (int)kPadCode ? padv : (int)s_subst[code * a.C + tcode] - 2 * a.ge;
                    }
                    sv[hh] = x;
                }
                o[e] = ((uint32_t)sv[0] & 0xffffu) | ((uint32_t)sv[1] << 16);
                if (BSA_PADSLOT && K % 4 != 0 && c == K)
                    o[e] = (ln & 15) == 0 ? (code == (int)kPadCode ? kBias2 : pack2b(a.go - a.ge)) : 0u;
            }
            prof[idx] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        BSA_TMA_FLIP()
        // this lane's two templates (low / high half)
        const uint32_t mLo = half ? m[2] : m[0], mHi = half ? m[3] : m[1];
        const uint64_t outLo = half ? it.out[2] : it.out[0], outHi = half ? it.out[3] : it.out[1];
        const int lastLo = mLo ? (int)((mLo - 1) / K) + 16 * half : -1, slotLo = mLo ? (int)((mLo - 1) % K) : 0;
        const int lastHi = mHi ? (int)((mHi - 1) / K) + 16 * half : -1, slotHi = mHi ? (int)((mHi - 1) % K) : 0;
        const uint32_t hdiag0 = lrel == 0 ? kBias2 : pack2b(a.go - a.ge);
        for (;;) {
            uint32_t c = 0;
            if (lane == 0) c = atomicAdd(&s_chunk, 1u);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= nch) break;
            const uint64_t ca = c <= nbig ? head * c / nbig : head + (span - head) * (c - nbig) / nsmall;
            const uint64_t cb = c + 1 <= nbig ? head * (c + 1) / nbig : head + (span - head) * (c + 1 - nbig) / nsmall;
            const uint32_t qa = c == 0 ? it.q_begin : lower_bound_off(xoff, it.q_begin, it.q_end, x0 + ca);
            const uint32_t qb = c + 1 == nch ? it.q_end : lower_bound_off(xoff, it.q_begin, it.q_end, x0 + cb);
            const uint64_t g0 = xoff[qa], g1 = xoff[qb];
            if (g1 <= g0) continue;
            stream_block16_fa<K, true>(a.QA.codes, g0, g1, prof, lane, lastLo, slotLo, lastHi, slotHi, hdiag0,
                                       pack2b(a.go - a.ge), pack2(a.go - a.ge), a.ge, (int)mLo, (int)mHi, a.scores,
                                       outLo + (qa - it.q_begin), outHi + (qa - it.q_begin), a.Q.off + qa);
        }
    }
}

}  // namespace bsa
#include "wave_kernels.cuh"
namespace bsa {

// Local-alignment walk (traceback_local_kernel).
// One WARP walks one pair; all 32 lanes run the same (uniform) walk, lane 0 writes.  The walk mostly
// moves diagonally, i.e. one step (a new 128-byte line) back per glyph, and every load depends on
// the previous one: the warp therefore keeps the direction words of the current (pass, lane, word)
// column for the last 32 steps, one per lane -- a refill is ONE warp-wide load of 32 independent
// lines, and a lookup is a shuffle.  (With one thread per pair, the threads of a warp diverged on
// every refill and every state change and all waited for the slowest.)
constexpr int kTraceWin = 32;               // DirWindow (the local-alignment walk): register window of 32 steps
constexpr int kLocalTraceThreads = 256;     // 8 pairs per CTA

struct DirWindow {
    uint32_t key = 0xffffffffu, top = 0, v = 0;
    __device__ __forceinline__ uint32_t get(const uint32_t* __restrict__ dirs, size_t plane, uint32_t W,
                                            uint32_t pass, uint32_t step, uint32_t lane, uint32_t w,
                                            uint32_t me) {
        const uint32_t k = (pass * 32u + lane) * 4u + w;
        if (k != key || step > top || step + kTraceWin <= top) {      // warp-uniform
            key = k;
            top = step;
            const uint32_t* base = dirs + pass * plane + (size_t)lane * W + w;
            v = me <= step ? base[(size_t)(step - me) * 32 * W] : 0u;
        }
        return __shfl_sync(0xffffffffu, v, top - step);
    }
};

// Position of DP column j in the direction planes -- (pass, lane, slot) -- kept incrementally: the
// walk only ever moves one column to the left, so the two divisions by the runtime K are paid once.
struct ColCursor {
    uint32_t K, pass, lane, c;
    __device__ __forceinline__ void init(uint32_t K_, uint32_t j) {   // j >= 1
        K = K_;
        const uint32_t col = j - 1, BK = 32 * K;
        pass = col / BK;
        const uint32_t lc = col - pass * BK;
        lane = lc / K;
        c = lc - lane * K;
    }
    __device__ __forceinline__ void left() {
        if (c > 0) { --c; return; }
        c = K - 1;
        if (lane > 0) { --lane; return; }
        lane = 31;
        --pass;            // wraps at column 0; the walk stops there (j == 0)
    }
    __device__ __forceinline__ uint32_t word() const { return c >> 3; }
    __device__ __forceinline__ uint32_t shift() const {
        const uint32_t w = c >> 3;
        const uint32_t cnt = (K - 8 * w) < 8 ? (K - 8 * w) : 8;
        return 4 * (cnt - 1 - (c & 7));
    }
};


// Local alignment (LocalAlignment::align, bioshell-seq/src/alignment/local.rs:83-207): the
// direction-store kernel with LOCAL recurrences.  Every lane tracks the first strict maximum of
// H over its cells; the warp merges them by (score desc, row asc, column asc), which is the
// reference's row-major first-strict-maximum (E and F can never exceed H in the same cell, so
// best_state is always H, local.rs:185-202).
struct LocalOut {
    int32_t score;
    uint32_t end_q, end_t;      // recent_end_point(): 1-based DP indices of the best cell
    uint32_t start_q, start_t;  // backtrace(): where the walk stopped (0-based start offsets)
};

template <int K>
__global__ void __launch_bounds__(kThreads) gotoh_local_kernel(const KArgs a, LocalOut* __restrict__ lout) {
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_item;
    constexpr int ROW = KTraits<K>::ROW;
    constexpr int W = KTraits<K>::W;
    uint4* prof = smem;
    uint4* rsH = smem + (size_t)a.C * ROW;
    uint4* rsF = rsH + ROW;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    BSA_TMA_PREAMBLE(K, a.C, a.subst)

    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.item_counter, 1u);
        __syncthreads();
        const uint32_t ii = s_item;
        if (ii >= a.n_items) break;
        const Item it = a.items[ii];
        const uint64_t t0 = a.T.off[it.t];
        const uint32_t m = (uint32_t)(a.T.off[it.t + 1] - t0);
        const uint8_t* tc = a.T.codes + t0;
        const Consts cs = make_consts<K>(a.go, a.ge, 0, true);
        const uint32_t npass = (m + 32 * K - 1) / (32 * K);

        for (uint32_t pass = 0; pass < npass; ++pass) {
            const uint32_t colbase = pass * 32 * K;
            __syncthreads();
            {
                BSA_TMA_PTRS(K, a.C)
                if (threadIdx.x == 0) {
                    const uint32_t nb = slice_bytes(tc, colbase, min(m - colbase, 32u * K));
                    TmaStage::arm(&s_mbar, nb);
                    TmaStage::copy(&s_mbar, s_tcA, slice_src(tc, colbase), nb);
                }
                TmaStage::wait(&s_mbar, s_phase);
                build_profile<K>(prof, rsH, rsF, s_subst, slice_view(s_tcA, tc, colbase), m, colbase, a, cs);
            }
            __syncthreads();
            BSA_TMA_FLIP()
            const bool lastp = (pass + 1 == npass);
            for (uint32_t pi = it.q_begin + warp; pi < it.q_end; pi += kWarpsPerCta) {
                const PairRec pr = a.pairs[pi];
                const uint64_t g0 = a.Q.off[pr.q], g1 = a.Q.off[pr.q + 1];
                const uint32_t n = (uint32_t)(g1 - g0);
                uint32_t* dirs = a.dirs + pr.dir_off + (size_t)pass * (size_t)(n + 32) * 32 * W;
                LaneBest lb;
                stream_block<K, true, true, false, true>(a.Q.codes, g0, g1, prof, rsH, rsF, lane, pass == 0, lastp,
                                                         31, 0, 0, cs, a.one, a.scratch + pr.scr_off, nullptr,
                                                         nullptr, 0, dirs, nullptr, nullptr, nullptr, &lb, colbase);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    LaneBest o;
                    o.v = __shfl_down_sync(0xffffffffu, lb.v, off);
                    o.row = __shfl_down_sync(0xffffffffu, lb.row, off);
                    o.col = __shfl_down_sync(0xffffffffu, lb.col, off);
                    if (o.v > lb.v || (o.v == lb.v && (o.row < lb.row || (o.row == lb.row && o.col < lb.col)))) lb = o;
                }
                if (lane == 0) {
                    LocalOut cur;
                    if (pass == 0) cur = LocalOut{0, 0u, 0u, 0u, 0u};
                    else cur = lout[pr.out];
                    const int sc = lb.v >> 2;
                    // earlier blocks hold smaller columns: a later block wins ties only on an earlier row
                    if (sc > cur.score || (sc == cur.score && sc > 0 && lb.row + 1 < cur.end_q)) {
                        cur.score = sc;
                        cur.end_q = lb.row + 1;
                        cur.end_t = lb.col + 1;
                    }
                    lout[pr.out] = cur;
                }
                __syncwarp();
            }
        }
    }
}

// LocalAlignment::backtrace (local.rs:213-273): from the best cell, state H, until a STOP
__global__ void traceback_local_kernel(const TraceArgs a, LocalOut* __restrict__ lout) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t me = threadIdx.x & 31u;
    const bool writer = me == 0;
    if (p >= a.n_pairs) return;
    const PairRec pr = a.pairs[p];
    const uint32_t n = (uint32_t)(a.Q.off[pr.q + 1] - a.Q.off[pr.q]);
    const uint32_t m = (uint32_t)(a.T.off[pr.t + 1] - a.T.off[pr.t]);
    const uint32_t K = pr.k, W = (K + 7) / 8;
    const size_t plane = (size_t)(n + 32) * 32 * W;
    const uint32_t* dirs = a.dirs + pr.dir_off;
    uint8_t* out = (a.path && writer) ? a.path + pr.path_off : nullptr;
    LocalOut lo = lout[pr.out];
    uint32_t pos = n + m, i = lo.end_q, j = lo.end_t;
    int st = 0;
    DirWindow dw;
    ColCursor cur;
    if (j > 0) cur.init(K, j);
    while (i > 0 && j > 0) {
        const uint32_t step = (i - 1) + cur.lane;
        const uint32_t nib = (dw.get(dirs, plane, W, cur.pass, step, cur.lane, cur.word(), me) >> cur.shift()) & 15u;
        if (st == 0) {
            const uint32_t hd = nib & 3u;
            if (hd == 0u) break;                                   // arrows == 0: STOP
            if (hd == 3u) { --pos; if (out) out[pos] = '*'; --i; --j; cur.left(); }
            else if (hd == 2u) st = 1;
            else st = 2;
        } else if (st == 1) {
            --pos; if (out) out[pos] = '-';
            --j;
            cur.left();
            st = (nib & 8u) ? 1 : 0;
        } else {
            --pos; if (out) out[pos] = '|';
            --i;
            st = (nib & 4u) ? 2 : 0;
        }
    }
    if (writer) {
        a.path_start[p] = pos;
        lo.start_q = i;
        lo.start_t = j;
        lout[pr.out] = lo;
    }
}

// ---- sequence-store construction (K0) ----
// presence[8]: 256-bit mask of byte values that occur
__global__ void byte_presence_kernel(const uint8_t* __restrict__ raw, uint64_t total,
                                     uint32_t* __restrict__ presence) {
    __shared__ uint32_t sm[8];
    if (threadIdx.x < 8) sm[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t b = raw[i];
#pragma unroll
        for (int w = 0; w < 8; ++w)
            if ((b >> 5) == (uint32_t)w) loc[w] |= 1u << (b & 31);
    }
#pragma unroll
    for (int w = 0; w < 8; ++w)
        if (loc[w]) atomicOr(&sm[w], loc[w]);
    __syncthreads();
    if (threadIdx.x < 8 && sm[threadIdx.x]) atomicOr(&presence[threadIdx.x], sm[threadIdx.x]);
}

// raw bytes -> residue codes through a 256-entry LUT (similarity_score.rs:125-134 encodes
// per pair; here it happens once per set), 16 bytes per thread where aligned.
__global__ void encode_kernel(const uint8_t* __restrict__ raw, uint8_t* __restrict__ codes,
                              uint64_t total, const uint8_t* __restrict__ lut_g) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x & 255] = lut_g[threadIdx.x & 255];
    __syncthreads();
    const uint64_t nvec = total / 16;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint4* rv = reinterpret_cast<const uint4*>(raw);
    uint4* cv = reinterpret_cast<uint4*>(codes);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 x = rv[i];
        uint32_t* w = reinterpret_cast<uint32_t*>(&x);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = w[k];
            w[k] = (uint32_t)lut[v & 255] | ((uint32_t)lut[(v >> 8) & 255] << 8) |
                   ((uint32_t)lut[(v >> 16) & 255] << 16) | ((uint32_t)lut[v >> 24] << 24);
        }
        cv[i] = x;
    }
    for (uint64_t i = nvec * 16 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += stride)
        codes[i] = lut[raw[i]];
}

// Derived store of a subset of the sequences (the long ones, see the hybrid plan in bsa_api.cu):
// sequence idx[i] of `src` becomes sequence i of the new store; codes keep their last-residue flags.
__global__ void gather_seqs_kernel(const SeqStoreDev src, const uint32_t* __restrict__ idx,
                                   const uint64_t* __restrict__ dst_off, uint8_t* __restrict__ dst) {
    const uint32_t i = blockIdx.x;
    const uint64_t s0 = src.off[idx[i]], len = src.off[idx[i] + 1] - s0, d0 = dst_off[i];
    for (uint64_t k = threadIdx.x; k < len; k += blockDim.x) dst[d0 + k] = src.codes[s0 + k];
}

// Even-aligned copy of a store (bsa_api.cu, build_aligned): sequence i goes to dst_off[i] (even), an
// odd-length one behind a PAD byte; codes keep their last-residue flags, which end up on odd positions.
__global__ void align_seqs_kernel(const SeqStoreDev src, const uint64_t* __restrict__ dst_off,
                                  uint8_t* __restrict__ dst) {
    const uint32_t i = blockIdx.x;
    const uint64_t s0 = src.off[i], len = src.off[i + 1] - s0, d0 = dst_off[i] + (len & 1u);
    if ((len & 1u) && threadIdx.x == 0) dst[d0 - 1] = (uint8_t)kPadCode;
    for (uint64_t k = threadIdx.x; k < len; k += blockDim.x) dst[d0 + k] = src.codes[s0 + k];
}

__global__ void mark_last_kernel(uint8_t* __restrict__ codes, const uint64_t* __restrict__ off,
                                 uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && off[i + 1] > off[i]) codes[off[i + 1] - 1] |= (uint8_t)kLastFlag;
}

// writes a few (index, score, n_identical) triples computed on the host (empty sequences)
struct Fix { uint64_t k; int32_t score; uint32_t nid; };
__global__ void apply_fix_kernel(const Fix* __restrict__ fx, uint32_t n, int32_t* scores,
                                 uint32_t* nident) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (scores) scores[fx[i].k] = fx[i].score;
    if (nident) nident[fx[i].k] = fx[i].nid;
}

}  // namespace bsa
