// K3 -- long pairs (templates of 4,096+ residues): the intra-task wavefront fill and the
// warp-parallel traceback.  Included by gotoh_kernels.cuh (needs KArgs, PairRec, TmaStage, ld_code).
//
// FILL (`gotoh_wave_kernel`).  The DP matrix of ONE pair (GlobalAligner::align,
// bioshell-seq/src/alignment/global.rs:57-144) is cut into column blocks of 256 and every
// (pair, block) is a work item run by one warp: lane l owns 8 consecutive columns, the lanes are
// skewed by one STEP of 4 rows, and a step computes the 4 x 8 cells of a lane as four
// independent dependency chains.  The blocks of a pair run concurrently in different warps (on
// different SMs) and sweep the matrix as a staggered wavefront; a block hands its right boundary
// column to its neighbour through self-validating 16-byte entries {H, epoch, E, epoch} in global
// memory (no fence, no counter: the consumer re-reads what does not carry this launch's epoch).
// The kernel is bound by the critical path of the longest pair -- its rows plus the hand-off lag
// of its ~137 blocks -- i.e. by how many instructions a step ISSUES, so everything here is
// about a small step:
//
//  * DIRECTION-FRAME CELL, 11 instructions (8 ALU pipe + 3 FMA pipe).  A DP value is
//        v = (score - (i + j) gap_extend) << 3 | priority << 1 | m
//    In the moving frame both gap extensions are free and both borders are the constant
//    go - ge; E and F are born with their H-max priority (2 / 1), the diagonal carries 3
//    (folded into the profile), so ONE signed max3 applies global.rs:161-169 (diag > E > F).
//    m marks "this E/F value is an extension": the extension candidate is `prev | 1`, the
//    opening has m = 0, so on equal scores the extension wins (global.rs:109,122) and the bit
//    that falls out IS the traceback flag.  Per cell: IMAD (diagonal), VIMNMX3, LOP3 (clear the
//    low bits), 2 x LOP3 (| 1), 2 x VIADDMNMX, and for the directions one LOP3 + one funnel shift
//    (3 bits: H source, E flag) plus two IMAD (F flag, kept on the FMA pipe).
//  * TOP PADDING instead of a tail.  The rows are padded at the TOP to a multiple of 4 with rows
//    of a `pad` residue whose profile row scores 0 on the diagonal: in the frame such a row
//    reproduces the top border exactly (H* = go - ge), so the last row of the last step is row n,
//    the final score is simply a register after the loop, lanes that have not reached their first
//    row need no reset, and the loop body has no special cases at all.
//  * The query reaches the lanes as ready-made profile-row offsets in a 256-entry ring in shared
//    memory (one LDS.128 per step and lane); the left boundary column of lane 0 sits in a second
//    ring, refilled every 4 steps from the neighbour's entries (block 0: constants).
//  * Direction words: 32 bits per row of a lane's 8 columns = 8 x 3 bits {H source, E flag} + 8
//    F flags; a lane writes its 4 rows with one 16-byte store, a warp 512 contiguous bytes a step.
//
// TRACEBACK (`traceback_kernel`).  GlobalAligner::backtrace (global.rs:146-201) is a sequential
// walk, but its moves come in RUNS (diagonals between gaps, gap extensions): the 32 lanes look at
// the next 32 cells along the current direction at once, a ballot gives the length of the run,
// and the walk advances by the whole run -- glyphs and identity counts of the run are written /
// counted by the lanes in parallel.  The direction words come from two windows in shared memory
// that the bulk-copy engine (cp.async.bulk + mbarrier, UBLKCP) keeps ahead of the walk.
#pragma once

namespace bsa {

#ifndef BSA_WAVE_WARPS
#define BSA_WAVE_WARPS 8
#endif
#ifndef BSA_WAVE_SLEEP_MAX
#define BSA_WAVE_SLEEP_MAX 512  // ns: a block that waits for its left neighbour backs off up to this (a spinning warp takes issue slots from the working ones)
#endif
#ifndef BSA_WAVE_AFMA
#define BSA_WAVE_AFMA 0       // {H source, E flag} accumulated with IMADs (7 ALU + 5 FMA per cell) instead of a funnel shift (8 + 3)
#endif
#ifndef BSA_WAVE_HIPRIO
#define BSA_WAVE_HIPRIO 1     // first items (the longest pairs) start on the upper half of the warps of a CTA
#endif
constexpr int kWaveK = 8;                  // columns per lane: 256-column blocks
constexpr int kWaveR = 4;                  // rows per step
constexpr int kWaveWarps = BSA_WAVE_WARPS;
constexpr uint32_t kPlaneRowsShift = 28;   // PairRec::k = K | rows-per-line << 28 (0: classic planes, 4: K3 frame planes)
constexpr uint32_t kPlaneKMask = (1u << kPlaneRowsShift) - 1u;
constexpr uint32_t kWaveBndPad = 128;      // entries in front of a boundary column: lane 31's rows before the stream land there
constexpr uint32_t kWaveBatch = 16;        // boundary entries per refill (4 steps)

// shared memory of one wavefront worker (bytes): residue-offset ring (1 KB, 1 KB aligned), boundary
// ring (512 B), template slice staging (512 B), then C + 1 profile rows of 1 KB (row C: the pad residue)
__host__ __device__ constexpr size_t wave_region_bytes(int C) { return 2048u + (size_t)(C + 1) * 1024u; }
__host__ __device__ constexpr size_t wave_smem_bytes(int warps, int C) {
    return (size_t)warps * wave_region_bytes(C) + subst_stage_bytes(C);
}
// entries of one boundary column of a pair with n rows
__host__ __device__ constexpr uint64_t wave_bnd_stride(uint64_t n) { return n + 4u + kWaveBndPad; }

__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u4(uint4* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// shared memory by 32-bit address (the ring addresses wrap with one LOP3)
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t x) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(x) : "memory");
}
__device__ __forceinline__ void sts_u2(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}

// (a & m) | (b & ~m) in one LOP3 (the mask must sit in a register: two immediates would make it two)
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

struct WaveConsts {
    int GOE;   // (go - ge) << 3 | 2 << 1 : an E opening on top of a clean H (priority 2, m = 0)
    int GOF;   // (go - ge) << 3 | 1 << 1 : an F opening (priority 1, m = 0)
    int HB;    // (go - ge) << 3          : H* on both borders (clean low bits)
    int EB;    // HB + GOE                : E*[i][1], opened from the left border
    int FB;    // HB + GOF                : F*[1][j], opened from the top border
};
__device__ __forceinline__ WaveConsts make_wave_consts(int go, int ge) {
    WaveConsts w;
    w.HB = (go - ge) * 8;
    w.GOE = w.HB + 4;
    w.GOF = w.HB + 2;
    w.EB = w.HB + w.GOE;
    w.FB = w.HB + w.GOF;
    return w;
}

// One column block of one pair.  `region` is this warp's shared memory, its profile already built.
//   qc      codes of the query (row i is qc[i - 1])
//   n, pad  rows, and the top padding that makes n + pad a multiple of 4
//   bnd_in  boundary column written by the block to the left (entry t = padded row index), null for block 0
//   bnd_out this block's boundary column (null for the last block)
//   dirs    this block's direction plane: [step][lane][4 rows] words
// Returns H*[n][m-column of (lane_last, slot_last)] << 3 on lane `lane_last` (last block only).
__device__ __forceinline__ int wave_block_frame(const uint8_t* __restrict__ qc, const uint32_t n, const uint32_t pad,
                                                const uint32_t region_s, const int lane, const bool first,
                                                const bool lastp, const int lane_last, const int slot_last,
                                                const uint4* bnd_in, uint4* bnd_out, const uint32_t epoch,
                                                uint32_t* __restrict__ dirs, const uint32_t padoff,
                                                const WaveConsts w, const int one, const int two, const int mone,
                                                const uint32_t notone) {
    constexpr int K = kWaveK, R = kWaveR;
    const uint32_t X = n + pad;                     // padded rows, a multiple of 4
    const uint32_t tiles = X >> 2;
    const uint32_t nd = tiles + (uint32_t)(lastp ? lane_last : 31);   // steps until the last lane of interest is through
    const uint32_t ring_s = region_s, bring_s = region_s + 1024u;
    const uint32_t prof_s = region_s + 2048u + (uint32_t)lane * 16u;
    const bool writer = lane == 31 && !lastp;

    // profile-row offset of padded row t (t < 0: before the stream; t < pad: top padding; past the end: anything valid)
    auto fetch = [&](int t) -> uint32_t {
        const int qi = t - (int)pad;
        return (qi >= 0 && qi < (int)n) ? ld_code(qc + qi) : 0xffffffffu;
    };
    auto entry = [&](uint32_t code) -> uint32_t { return code == 0xffffffffu ? padoff : (code & kCodeMask) << 10; };

    __syncwarp();
    // residue ring: steps [-32, 16) to begin with (rows [-128, 64)), 6 rows per lane
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int t = -128 + 32 * k + lane;
        sts_u32(ring_s + (((uint32_t)t & 255u) << 2), entry(fetch(t)));
    }
    uint4 inflight = make_uint4(0u, 0u, 0u, 0u);
    if (first) {
        // the left border is constant in the frame: H*[i][0] = go - ge, E*[i][1] = 2 (go - ge); the one
        // exception is H*[0][0] = 0, the diagonal of cell (1, 1): it is the entry of the last padding row
        sts_u2(bring_s + (uint32_t)lane * 8u, (uint32_t)w.HB, (uint32_t)w.EB);
        sts_u2(bring_s + (uint32_t)(lane + 32) * 8u, (uint32_t)w.HB, (uint32_t)w.EB);
        __syncwarp();
        if (lane == 0 && pad) sts_u32(bring_s + (pad - 1u) * 8u, 0u);
    } else if ((uint32_t)lane < kWaveBatch && (uint32_t)lane < X) {
        inflight = ld_volatile_u4(bnd_in + lane);
    }
    __syncwarp();

    int H[K], Fr[K];
#pragma unroll
    for (int c = 0; c < K; ++c) { H[c] = w.HB; Fr[c] = w.FB; }
    int hdiag = (first && lane == 0 && pad == 0u) ? 0 : w.HB;
    // what enters from the left in the coming step (H of the column before this lane's, E of its first column):
    // shuffled over at the end of the step before; lane 0 takes the boundary ring's entries instead
    int hin[R], er[R];
    uint32_t dA[R], dF[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { hin[r] = w.HB; er[r] = w.EB; dA[r] = 0u; dF[r] = 0u; }
    uint32_t ra = ring_s + ((0u - 16u * (uint32_t)lane) & 1023u);      // ring entry of step S - lane
    uint32_t ba = bring_s;                                               // boundary entries of step S (lane 0's rows)
    char* dp = reinterpret_cast<char*>(dirs) + (size_t)lane * 16u;
    uint4* bo = bnd_out - 31 * R;                                        // lane 31's rows of step S: 4 (S - 31) ..
    uint32_t rf0 = 0xffffffffu, rf1 = 0xffffffffu;
    // profile rows of the coming step's 4 residues: replaced row by row inside the step, as soon as a row is through
    int T[R][K];
    uint4 offs = lds_u4(ra);
#define BSA_WAVE_LOAD_T(r, off)                                                                     \
    {                                                                                               \
        const uint4 v0 = lds_u4(prof_s + (off));                                                    \
        const uint4 v1 = lds_u4(prof_s + (off) + 512u);                                             \
        T[r][0] = (int)v0.x; T[r][1] = (int)v0.y; T[r][2] = (int)v0.z; T[r][3] = (int)v0.w;         \
        T[r][4] = (int)v1.x; T[r][5] = (int)v1.y; T[r][6] = (int)v1.z; T[r][7] = (int)v1.w;         \
    }
    BSA_WAVE_LOAD_T(0, offs.x) BSA_WAVE_LOAD_T(1, offs.y) BSA_WAVE_LOAD_T(2, offs.z) BSA_WAVE_LOAD_T(3, offs.w)
    uint4 b0 = make_uint4(0u, 0u, 0u, 0u), b1 = b0;

    for (uint32_t S0 = 0; S0 < nd; S0 += 4u) {
        // ---- every 4 steps: refill the rings ----
        __syncwarp();   // the ring entries rewritten below were last read (by every lane, fetching ahead) in the steps before
        if ((S0 & 15u) == 0u) {
            const int t = 4 * (int)(S0 + 16u) + 2 * lane;               // steps [S0 + 16, S0 + 32): asked for now,
            rf0 = fetch(t);
            rf1 = fetch(t + 1);
        } else if ((S0 & 15u) == 8u) {
            const int t = 4 * (int)(S0 + 8u) + 2 * lane;                // stored 8 steps later, read from step S0 + 7 on
            sts_u2(ring_s + (((uint32_t)t & 255u) << 2), entry(rf0), entry(rf1));
        }
        if (!first) {
            // boundary rows [4 S0, 4 S0 + 16): asked for 4 steps ago, must be complete now
            const uint32_t t = 4u * S0 + (uint32_t)lane;
            bool ok = (uint32_t)lane >= kWaveBatch || t >= X;
            for (uint32_t ns = 32u;;) {
                if (!ok) {
                    if (inflight.y == epoch && inflight.w == epoch) {
                        sts_u2(bring_s + ((t & 63u) << 3), inflight.x, inflight.z);
                        ok = true;
                    } else {
                        inflight = ld_volatile_u4(bnd_in + t);
                    }
                }
                if (__all_sync(0xffffffffu, ok)) break;
                __nanosleep(ns);
                ns = min(2u * ns, (uint32_t)BSA_WAVE_SLEEP_MAX);
            }
            if ((uint32_t)lane < kWaveBatch && t + kWaveBatch < X) inflight = ld_volatile_u4(bnd_in + t + kWaveBatch);
        } else if (S0 == 4u) {
            if (lane == 0 && pad) sts_u32(bring_s + (pad - 1u) * 8u, (uint32_t)w.HB);   // H*[0][0] has been used
        }
        __syncwarp();
        b0 = lds_u4(ba);                   // lane 0's entries of step S0 (what the last step fetched ahead was the old batch)
        b1 = lds_u4(ba + 16u);

        const uint32_t S1 = min(S0 + 4u, nd);
#pragma unroll 1
        for (uint32_t S = S0; S < S1; ++S) {
            // ---- lane 0: the boundary ring instead of a neighbour; then fetch ahead for step S + 1 ----
            if (lane == 0) {
                hin[0] = (int)b0.x; er[0] = (int)b0.y; hin[1] = (int)b0.z; er[1] = (int)b0.w;
                hin[2] = (int)b1.x; er[2] = (int)b1.y; hin[3] = (int)b1.z; er[3] = (int)b1.w;
            }
            int hd[R];
            hd[0] = hdiag;
#pragma unroll
            for (int r = 1; r < R; ++r) hd[r] = hin[r - 1];
            hdiag = hin[R - 1];
            ra = (ra & ~1023u) | ((ra + 16u) & 1023u);
            offs = lds_u4(ra);
            ba = (ba & ~511u) | ((ba + 32u) & 511u);
            b0 = lds_u4(ba);
            b1 = lds_u4(ba + 16u);
            const uint32_t o[R] = {offs.x, offs.y, offs.z, offs.w};
            // ---- the 4 x 8 cells, anti-diagonal by anti-diagonal: the cells of a diagonal are independent, and
            //      every operation is written for all of them before the next one (dependent instructions 4 apart) ----
#pragma unroll
            for (int dg = 0; dg < K + R - 1; ++dg) {
                int dd[R], hh[R], hc[R], f1[R];
#define BSA_DIAG(STMT)                                                                              \
                _Pragma("unroll") for (int r = 0; r < R; ++r) {                                     \
                    const int c = dg - r;                                                           \
                    if (c >= 0 && c < K) { STMT }                                                   \
                }
                BSA_DIAG(dd[r] = hd[r] * one + T[r][c];)
                BSA_DIAG(hh[r] = max3_s32(dd[r], er[r], Fr[c]);)
                BSA_DIAG(f1[r] = Fr[c] | 1;)
                BSA_DIAG(hc[r] = hh[r] & ~7;)
#if BSA_WAVE_AFMA
                // {H source, E flag} on the FMA pipe: the low bits of h are h - hc, the word shifts by a multiplication
                BSA_DIAG(dA[r] = dA[r] * 8u + bitsel((uint32_t)(hc[r] * mone + hh[r]), (uint32_t)er[r], notone);)
#else
                BSA_DIAG(dA[r] = __funnelshift_r(dA[r], bitsel((uint32_t)hh[r], (uint32_t)er[r], notone), 3);)
#endif
                BSA_DIAG(dF[r] = (uint32_t)(Fr[c] * mone + (int)(dF[r] * (uint32_t)two + (uint32_t)f1[r]));)   // 2 dF + (1 - m)
                BSA_DIAG(er[r] = addmax_s32(hc[r], w.GOE, er[r] | 1);)
                BSA_DIAG(Fr[c] = addmax_s32(hc[r], w.GOF, f1[r]);)
                BSA_DIAG(hd[r] = H[c]; H[c] = hc[r];)
#undef BSA_DIAG
                // the row that just did its last column gets the coming step's profile row
                if (dg == K - 1) BSA_WAVE_LOAD_T(0, o[0])
                if (dg == K) BSA_WAVE_LOAD_T(1, o[1])
                if (dg == K + 1) BSA_WAVE_LOAD_T(2, o[2])
                if (dg == K + 2) BSA_WAVE_LOAD_T(3, o[3])
            }
            // ---- out to the right (H of the last column, E of the column after it) and the directions ----
            int oh[R];
#pragma unroll
            for (int r = 0; r < R - 1; ++r) oh[r] = hd[r + 1];
            oh[R - 1] = H[K - 1];
            {
                uint4 wv;
                // BSA_WAVE_AFMA: column c's 3 bits at 3 (7 - c), F flags in byte 3; else 3 bits at 8 + 3 c, F flags in byte 0
                constexpr uint32_t SEL = BSA_WAVE_AFMA ? 0x4210u : 0x3214u;
                wv.x = __byte_perm(dA[0], dF[0], SEL);
                wv.y = __byte_perm(dA[1], dF[1], SEL);
                wv.z = __byte_perm(dA[2], dF[2], SEL);
                wv.w = __byte_perm(dA[3], dF[3], SEL);
                *reinterpret_cast<uint4*>(dp) = wv;
                dp += 512;
            }
            if (writer) {
#pragma unroll
                for (int r = 0; r < R; ++r) st_volatile_u4(bo + r, (uint32_t)oh[r], epoch, (uint32_t)er[r], epoch);
            }
            bo += R;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                hin[r] = __shfl_up_sync(0xffffffffu, oh[r], 1);
                er[r] = __shfl_up_sync(0xffffffffu, er[r], 1);
            }
        }
    }
#undef BSA_WAVE_LOAD_T
    int res = H[0];
#pragma unroll
    for (int c = 1; c < K; ++c) res = (c == slot_last) ? H[c] : res;
    return res;
}

__global__ void __launch_bounds__(kWaveWarps * 32) gotoh_wave_kernel(const KArgs a) {
    extern __shared__ uint4 smem_raw[];
    __shared__ uint64_t s_bar[kWaveWarps + 1];
    // the rings wrap by address arithmetic: every worker's region starts on a 1 KB boundary (the host adds the slack)
    char* const smem = reinterpret_cast<char*>(smem_raw) + ((1024u - (TmaStage::smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int K = kWaveK;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int C = a.C;
    const size_t RB = wave_region_bytes(C);
    char* region = smem + (size_t)warp * RB;
    const uint32_t region_s = TmaStage::smem_u32(region);
    uint8_t* s_slice = reinterpret_cast<uint8_t*>(region + 1536);                     // this warp's template columns (<= 288 B)
    uint4* prof = reinterpret_cast<uint4*>(region + 2048);
    int16_t* s_subst = reinterpret_cast<int16_t*>(smem + (size_t)nwarps * RB);
    const WaveConsts w = make_wave_consts(a.go, a.ge);
    const int one = a.one, two = a.one + a.one2, mone = a.mone;
    const uint32_t notone = ~(uint32_t)a.one2;      // 0xfffffffe, opaque to the compiler (bitsel)
    const uint32_t padoff = (uint32_t)C << 10;

    // the substitution table: once per CTA, through the bulk-copy engine
    if (threadIdx.x == 0) {
        TmaStage::init(&s_bar[kWaveWarps]);
        TmaStage::arm(&s_bar[kWaveWarps], subst_stage_bytes(C));
        TmaStage::copy(&s_bar[kWaveWarps], s_subst, a.subst, subst_stage_bytes(C));
    }
    if (lane == 0) TmaStage::init(&s_bar[warp]);
    __syncthreads();
    TmaStage::wait(&s_bar[kWaveWarps], 0);
    uint32_t phase = 0;

    // Initial items: item i starts on CTA i % G, in warp slot i / G counted from the HIGHEST warp id down.
    // The items are in order of their pairs' critical paths, so the blocks of the longest pair are spread
    // one per SM (when only that pair is left every one of its warps has an SM to itself -- four working
    // warps on one SM run a third slower than two) and sit in the warp the issue arbiter prefers (higher
    // warp ids first).  Everything else is claimed from the counter.  An item's left neighbour always has
    // a smaller index: it is either one of the initial items (all CTAs are resident) or was claimed before.
    const uint32_t G = gridDim.x;
#if BSA_WAVE_HIPRIO
    uint32_t wi = (uint32_t)(nwarps - 1 - warp) * G + blockIdx.x;
#else
    uint32_t wi = (uint32_t)warp * G + blockIdx.x;
#endif
    const uint32_t n_static = G * (uint32_t)nwarps;

    for (;;) {
        if (wi >= a.n_items) {
            if (wi < n_static) wi = n_static;          // nothing initial for this warp: fall through to the counter
            else break;
        }
        if (wi >= n_static || wi >= a.n_items) {
            uint32_t x = 0;
            if (lane == 0) x = atomicAdd(a.item_counter, 1u);
            wi = n_static + __shfl_sync(0xffffffffu, x, 0);
            if (wi >= a.n_items) break;
        }
        const uint2 item = a.wave_items[wi];
        const uint32_t item_index = wi;
        unsigned long long trace_t0 = 0;
        if (a.wave_trace) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(trace_t0));
        wi = n_static;                                   // the next one comes from the counter
        const PairRec pr = a.pairs[item.x];
        const uint32_t pass = item.y;
        const uint64_t t0 = a.T.off[pr.t];
        const uint32_t m = (uint32_t)(a.T.off[pr.t + 1] - t0);
        const uint8_t* tc = a.T.codes + t0;
        const uint32_t npass = (m + 32 * K - 1) / (32 * K);
        const uint32_t colbase = pass * 32 * K;
        const uint64_t g0 = a.Q.off[pr.q];
        const uint32_t n = (uint32_t)(a.Q.off[pr.q + 1] - g0);
        const uint32_t pad = (4u - (n & 3u)) & 3u;

        // this block's template columns, staged by the bulk-copy engine; then the warp-private profile:
        // rows [code][half][lane] of 4 columns each, entries (s - 2 ge) << 3 | 3 << 1, row C = the pad residue
        __syncwarp();
        if (lane == 0) {
            const uint32_t nb = slice_bytes(tc, colbase, min(m - colbase, 32u * K));
            TmaStage::arm(&s_bar[warp], nb);
            TmaStage::copy(&s_bar[warp], s_slice, slice_src(tc, colbase), nb);
        }
        TmaStage::wait(&s_bar[warp], phase);
        phase ^= 1u;
        const uint8_t* sl = slice_view(s_slice, tc, colbase);
        const int tsub = -2 * a.ge;
        for (int idx = lane; idx < (C + 1) * 64; idx += 32) {
            const int code = idx >> 6, r = idx & 63, v = r >> 5, ln = r & 31;
            int o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t col = colbase + ln * K + 4 * v + e;
                o[e] = (code < C && col < m) ? ((int)s_subst[code * C + (sl[col] & kCodeMask)] + tsub) * 8 + 6 : 6;
            }
            prof[idx] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();

        const bool lastp = (pass + 1 == npass);
        const int lane_last = (int)((m - 1 - colbase) / K);
        const int slot_last = (int)((m - 1 - colbase) % K);
        const uint64_t stride = wave_bnd_stride(n);
        uint4* bnd = a.wave_bnd + pr.scr_off;              // (npass - 1) columns of `stride` entries, kWaveBndPad in front of each
        uint32_t* dirs = a.dirs + pr.dir_off + (size_t)pass * (size_t)(((n + 3u) >> 2) + 32u) * 128u;
        const int res = wave_block_frame(a.Q.codes + g0, n, pad, region_s, lane, pass == 0, lastp, lastp ? lane_last : 31, slot_last,
                                         pass ? bnd + (size_t)(pass - 1) * stride + kWaveBndPad : nullptr,
                                         lastp ? nullptr : bnd + (size_t)pass * stride + kWaveBndPad, a.epoch, dirs, padoff,
                                         w, one, two, mone, notone);
        if (lastp && lane == lane_last && a.scores)
            a.scores[pr.out] = (res >> 3) + (int)(n + m) * a.ge;      // out of the frame
        if (a.wave_trace && lane == 0) {
            unsigned long long t1;
            uint32_t smid;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            unsigned long long* tr = a.wave_trace + 4ull * item_index;
            tr[0] = trace_t0; tr[1] = t1; tr[2] = ((unsigned long long)smid << 8) | (unsigned long long)warp; tr[3] = ((n + pad) >> 2) + (lastp ? lane_last : 31);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Traceback.  Walks the stored directions exactly as GlobalAligner::backtrace does
// (global.rs:146-201): state H follows diag > E > F, state E / F keeps going while the cell it
// leaves was an extension.  One warp per pair; the 32 lanes test the next 32 cells along the
// current direction and the walk advances by the whole run.  Writes the path glyphs backwards into
// the pair's slot and counts identical residues on the way (raw-byte equality == code equality;
// '-' / '_' never count, msa.rs:264).
struct TraceArgs {
    SeqStoreDev Q, T;
    const PairRec* pairs;
    uint32_t n_pairs;
    const uint32_t* dirs;
    const uint8_t* isgap;
    uint32_t ncodes;        // entries of isgap
    uint8_t* path;          // may be null (identity only)
    uint32_t* path_start;   // per pair (by position in `pairs`): first byte of the path in its slot
    uint32_t* nident;       // may be null; indexed by PairRec::out
    uint32_t* status;       // set to 1 if a direction code is invalid
    uint32_t win_bytes;     // bytes of one direction window (two per warp)
};

constexpr int kTraceMaxWarps = 4;
constexpr uint32_t kTraceWinShort = 8192;      // pairs of the per-template kernels: 4 warps (pairs) per CTA
constexpr uint32_t kTraceWinLong = 49152;      // K3 pairs: one warp per CTA, 96 lines of 4 rows per window

__global__ void __launch_bounds__(kTraceMaxWarps * 32) traceback_kernel(const TraceArgs a) {
    extern __shared__ uint4 tb_smem[];
    __shared__ uint64_t tb_bar[kTraceMaxWarps][2];
    const uint32_t warp = threadIdx.x >> 5, me = threadIdx.x & 31u, nw = blockDim.x >> 5;
    const uint32_t p = blockIdx.x * nw + warp;
    if (p >= a.n_pairs) return;                // whole warps leave; the barriers below are per warp
    const PairRec pr = a.pairs[p];
    const uint8_t* qc = a.Q.codes + a.Q.off[pr.q];
    const uint8_t* tc = a.T.codes + a.T.off[pr.t];
    const uint32_t n = (uint32_t)(a.Q.off[pr.q + 1] - a.Q.off[pr.q]);
    const uint32_t m = (uint32_t)(a.T.off[pr.t + 1] - a.T.off[pr.t]);
    const bool frame = (pr.k >> kPlaneRowsShift) == 4u;         // K3 planes: 4 rows per line, {3 bits x 8 | 8 F flags} words
    const uint32_t rpl = frame ? 4u : 1u, rsh = frame ? 2u : 0u;
    const uint32_t pad = frame ? ((4u - (n & 3u)) & 3u) : 0u;   // top padding rows of the K3 planes
    const uint32_t K = pr.k & kPlaneKMask, W = (K + 7u) / 8u, BK = 32u * K;
    const uint32_t invK = (65536u + K - 1u) / K;                // x / K == x * invK >> 16 for x < 1024 (K <= 32)
    const uint32_t LW = 32u * W * rpl;                          // words per line (all lanes' words of one step)
    const uint32_t nlines = ((n + pad) >> rsh) + 32u;
    const size_t plane = (size_t)nlines * LW;
    const uint32_t WB = a.win_bytes / 4u;                       // words per window
    const uint32_t RL = WB / LW;                                // lines per window
    const uint32_t* dirs = a.dirs + pr.dir_off;
    uint8_t* out = a.path ? a.path + pr.path_off : nullptr;
    uint32_t* const win0 = reinterpret_cast<uint32_t*>(tb_smem) + (size_t)(warp * 2u) * WB;
    uint64_t* const bar0 = &tb_bar[warp][0];
    if (me == 0) { TmaStage::init(bar0); TmaStage::init(bar0 + 1); }
    __syncwarp();
    uint32_t phases = 0u;                   // bit b: parity the next completion of buffer b will have
    // window state: buffer `cb` holds lines [lo, hi) of column block `wpass`; the other buffer has (or gets) lines [nlo, lo)
    uint32_t cb = 0, wpass = 0xffffffffu, lo = 0, hi = 0, nlo = 0;
    bool have_next = false, next_ready = false;
    auto request = [&](uint32_t buf, uint32_t pass, uint32_t l0, uint32_t l1) {   // lines [l0, l1) of `pass` into `buf`
        __syncwarp();                       // every lane is done reading what the buffer held
        if (me == 0) {
            const uint32_t bytes = (l1 - l0) * LW * 4u;
            TmaStage::arm(bar0 + buf, bytes);
            TmaStage::copy(bar0 + buf, win0 + (size_t)buf * WB, dirs + pass * plane + (size_t)l0 * LW, bytes);
        }
    };
    auto ready = [&](uint32_t buf) {
        TmaStage::wait(bar0 + buf, (phases >> buf) & 1u);
        phases ^= 1u << buf;
    };

    // '-' / '_' codes as bit masks (msa.rs:264), so the identity test is register-only
    unsigned long long gap_lo = 0ull, gap_hi = 0ull;
    for (uint32_t code = 0; code < a.ncodes; ++code)
        if (a.isgap[code]) {
            if (code < 64) gap_lo |= 1ull << code; else gap_hi |= 1ull << (code - 64);
        }
    uint32_t i = n, j = m, pos = n + m, nid = 0;
    int st = 0;
    bool bad = false;
    uint32_t pass = m ? (m - 1u) / BK : 0u;
    uint32_t lc0 = m ? (m - 1u) - pass * BK : 0u;        // column of cell (., j) inside its block
    // the residues of a diagonal run are only needed for the count: they are consumed one run later
    uint32_t px = 0xffu, py = 0xfeu;
#define BSA_COUNT_PENDING() \
    nid += (px == py && !((((px & 64u) ? gap_hi : gap_lo) >> (px & 63u)) & 1ull)) ? 1u : 0u;

    while (i > 0 && j > 0) {
        const uint32_t t0 = i - 1u + pad;
        const uint32_t line0 = (t0 >> rsh) + ((lc0 * invK) >> 16);
        if (pass != wpass || line0 < lo || line0 >= hi) {                 // warp-uniform
            if (pass == wpass && have_next && line0 >= nlo && line0 < lo) {
                // the walk moved up into the window that was requested ahead
                cb ^= 1u;
                if (!next_ready) ready(cb);
                hi = lo;
                lo = nlo;
            } else {
                if (have_next && !next_ready) ready(cb ^ 1u);             // drain the request in flight
                wpass = pass;
                hi = line0 + 1u;
                lo = hi > RL ? hi - RL : 0u;
                request(cb, wpass, lo, hi);
                ready(cb);
            }
            have_next = lo > 0u;
            next_ready = false;
            if (have_next) {
                nlo = lo > RL ? lo - RL : 0u;
                request(cb ^ 1u, wpass, nlo, lo);
            }
        }
        // a run may reach into the window above: take it in as soon as the walk comes near
        if (have_next && !next_ready && line0 < lo + 12u) { ready(cb ^ 1u); next_ready = true; }
        const uint32_t floor_line = next_ready ? nlo : lo;

        // lane `me` looks at the cell `me` moves ahead in the current direction
        const uint32_t di = st != 1 ? me : 0u, dj = st != 2 ? me : 0u;
        bool ok = di < i && dj <= lc0;                                    // inside the matrix and this column block
        const uint32_t lc = ok ? lc0 - dj : 0u;
        const uint32_t ln = (lc * invK) >> 16, c = lc - ln * K;
        const uint32_t t = ok ? t0 - di : t0;
        const uint32_t line = (t >> rsh) + ln;
        ok = ok && line >= floor_line;
        uint32_t hdir = 0u, eext = 0u, fext = 0u;
        if (ok) {
            const uint32_t inner = (ln * rpl + (t & (rpl - 1u))) * W + (c >> 3);
            const uint32_t wofs = line >= lo ? cb * WB + (line - lo) * LW + inner : (cb ^ 1u) * WB + (line - nlo) * LW + inner;
            const uint32_t wd = win0[wofs];
            if (frame) {
#if BSA_WAVE_AFMA
                const uint32_t a3 = (wd >> (3u * (7u - c))) & 7u;
                fext = ((wd >> (31u - c)) & 1u) ^ 1u;
#else
                const uint32_t a3 = (wd >> (8u + 3u * c)) & 7u;
                fext = ((wd >> (7u - c)) & 1u) ^ 1u;
#endif
                hdir = a3 >> 1;
                eext = a3 & 1u;
            } else {
                const uint32_t left = K - 8u * (c >> 3);
                const uint32_t cnt = left < 8u ? left : 8u;
                const uint32_t nib = (wd >> (4u * (cnt - 1u - (c & 7u)))) & 15u;
                hdir = nib & 3u;
                eext = (nib >> 3) & 1u;
                fext = (nib >> 2) & 1u;
            }
        }
        const bool flag = st == 0 ? hdir == 3u : (st == 1 ? eext != 0u : fext != 0u);
        const uint32_t okm = __ballot_sync(0xffffffffu, ok);
        const uint32_t mk = __ballot_sync(0xffffffffu, ok && flag);
        const uint32_t t1 = mk == 0xffffffffu ? 32u : (uint32_t)__ffs((int)~mk) - 1u;   // cells of the run
        const bool seen = t1 < 32u && ((okm >> t1) & 1u);               // the cell that ends the run is in view
        uint32_t run;
        if (st == 0) {
            run = t1;
            BSA_COUNT_PENDING()
            if (me < run) {
                if (out) out[pos - 1u - me] = '*';
                px = qc[i - 1u - me] & kCodeMask;
                py = tc[j - 1u - me] & kCodeMask;
            } else {
                px = 0xffu; py = 0xfeu;
            }
            i -= run;
            j -= run;
            if (seen) {
                const uint32_t nh = __shfl_sync(0xffffffffu, hdir, (int)t1);
                if (nh == 2u) st = 1;
                else if (nh == 1u) st = 2;
                else { bad = true; break; }
            }
            lc0 -= run;
        } else if (st == 1) {
            run = t1 + (seen ? 1u : 0u);                                 // the opening cell is walked too, then back to H
            if (out && me < run) out[pos - 1u - me] = '-';
            j -= run;
            lc0 -= run;
            if (seen) st = 0;
        } else {
            run = t1 + (seen ? 1u : 0u);
            if (out && me < run) out[pos - 1u - me] = '|';
            i -= run;
            if (seen) st = 0;
        }
        pos -= run;
        if ((int)lc0 < 0) { lc0 += BK; --pass; }                          // wraps at column 0; the walk stops there (j == 0)
    }
    BSA_COUNT_PENDING()
#undef BSA_COUNT_PENDING
    if (have_next && !next_ready) ready(cb ^ 1u);          // nothing of this warp stays in flight when it leaves
    nid = __reduce_add_sync(0xffffffffu, nid);
    if (bad && me == 0) *a.status = 1u;
    // borders: row 0 is all E-extensions, column 0 all F-extensions (global.rs:81-88,96-97)
    if (out) {
        for (uint32_t k = me; k < j; k += 32u) out[pos - 1u - k] = '-';
    }
    pos -= j;
    if (out) {
        for (uint32_t k = me; k < i; k += 32u) out[pos - 1u - k] = '|';
    }
    pos -= i;
    if (me == 0) {
        a.path_start[p] = pos;
        if (a.nident) a.nident[pr.out] = nid;
    }
}

}  // namespace bsa
