// pipe_mix.cu -- how do ALU-pipe (DPX / LOP3) and FMA-pipe (IMAD) instructions overlap on sm_100a?
// Independent chains of the TAG cell with a configurable number of IMADs moved to IADD (ALU) or dropped.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/pipe_mix tools/microbench/pipe_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 8;
constexpr int kIters = 4096;

// MODE: 0 = TAG cell (4 ALU + 3 IMAD)     1 = 4 ALU + 3 IADD (7 ALU)    2 = 4 ALU only (IMAD results replaced by moves of hc)
//       3 = 3 IMAD only                    4 = classic cell (6 ALU + 2 IMAD)   5 = 4 ALU + 1 IMAD   6 = 4 ALU + 2 IMAD
//       7 = TAG cell, IMADs with an immediate multiplier-free form (x + y via IADD3 with 3 operands = ALU) [same as 1]
//       8 = 2 ALU (VIMNMX3 + VIADDMNMX) + 2 IMAD     9 = VIMNMX3 only   10 = VIADDMNMX only   11 = LOP3 only  12 = IMAD only
template <int MODE>
__global__ void __launch_bounds__(256) k(int seed, int one, int one2, int* sink) {
    int a[kChains], b[kChains], c[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 5 + 1 + threadIdx.x; c[i] = seed - i - 3 * threadIdx.x; }
    const int ge = seed | 1, go = seed + 3, go2 = seed + 77, mask = ~(127 << 9), t = seed * 5 + 11, ph = 2 << 12, pv = 1 << 12;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (MODE == 0) {
                const int d = c[i] * one + t;
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(a[i], ge, hc * one + go);
                b[i] = __viaddmax_s32(b[i], ge, hc * one2 + go2);
                c[i] = hc;
            } else if (MODE == 1) {
                const int d = c[i] + t;
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(a[i], ge, hc + go);
                b[i] = __viaddmax_s32(b[i], ge, hc + go2);
                c[i] = hc;
            } else if (MODE == 2) {
                const int h = __vimax3_s32(c[i], a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(a[i], ge, hc);
                b[i] = __viaddmax_s32(b[i], go, hc);
                c[i] = hc;
            } else if (MODE == 3) {
                a[i] = a[i] * one + t;
                b[i] = b[i] * one + go;
                c[i] = c[i] * one2 + go2;
            } else if (MODE == 4) {
                const int e = a[i] | ph;
                const int f = b[i] | pv;
                const int d = c[i] * one + t;
                const int h = __vimax3_s32(d, e, f);
                const int hc = h & mask;
                const int hg = hc * one + go;
                a[i] = __viaddmax_s32(e, ge, hg);
                b[i] = __viaddmax_s32(f, ge, hg);
                c[i] = hc;
            } else if (MODE == 5) {
                const int d = c[i] * one + t;
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(a[i], ge, hc);
                b[i] = __viaddmax_s32(b[i], go, hc);
                c[i] = hc;
            } else if (MODE == 6) {
                const int d = c[i] * one + t;
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(a[i], ge, hc * one + go);
                b[i] = __viaddmax_s32(b[i], go, hc);
                c[i] = hc;
            } else if (MODE == 8) {
                const int d = c[i] * one + t;
                const int h = __vimax3_s32(d, a[i], b[i]);
                a[i] = __viaddmax_s32(a[i], ge, h * one2 + go);
                c[i] = h;
            } else if (MODE == 9) {
                a[i] = __vimax3_s32(a[i], b[i], c[i]);
                b[i] = __vimax3_s32(b[i], c[i], it);
            } else if (MODE == 10) {
                a[i] = __viaddmax_s32(a[i], ge, b[i]);
            } else if (MODE == 11) {
                a[i] = (a[i] & mask) | b[i];
            } else if (MODE == 13) {
                const int d = c[i] + t;
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(a[i], ge, hc + go);
                b[i] = __viaddmax_s32(b[i], ge, hc + go2);
                c[i] = hc;
            } else if (MODE == 14) {
                int d, x, y;
                asm("mad.lo.s32 %0, %1, 1, %2;" : "=r"(d) : "r"(c[i]), "r"(t));
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                asm("mad.lo.s32 %0, %1, 1, %2;" : "=r"(x) : "r"(hc), "r"(go));
                asm("mad.lo.s32 %0, %1, 1, %2;" : "=r"(y) : "r"(hc), "r"(go2));
                a[i] = __viaddmax_s32(a[i], ge, x);
                b[i] = __viaddmax_s32(b[i], ge, y);
                c[i] = hc;
            } else if (MODE == 12) {
                a[i] = a[i] * one + t;
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i] ^ b[i] ^ c[i];
    if (s == 0x7fffffff) *sink = s;
}

template <int MODE>
void run(const char* name, int ops, int sms, int warps_per_smsp, float mhz) {
    int* sink;
    cudaMalloc(&sink, 4);
    const int blocks = sms * warps_per_smsp / 2;   // 256 threads = 8 warps = 2 per SMSP
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) k<MODE><<<blocks, 256>>>(1, 1, 1, sink);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(1, 1, 1, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // cycles per warp-level "cell" per SMSP
    const double cells_per_smsp = (double)warps_per_smsp * kIters * kChains;
    const double cyc = best * 1e-3 * mhz * 1e6 / cells_per_smsp;
    printf("%-44s warps/SMSP=%2d  %7.3f ms  %6.2f cycles per warp-cell (%d instr: %.2f cyc/instr)\n", name, warps_per_smsp, best, cyc, ops, cyc / ops);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float mhz = khz / 1000.f;
    printf("%s, %d SMs, %.0f MHz (nominal max; cycles assume the GPU boosts to it)\n", p.name, p.multiProcessorCount, mhz);
    for (int w : {4, 8, 16}) {
        run<0>("TAG cell: 4 ALU + 3 IMAD", 7, p.multiProcessorCount, w, mhz);
        run<2>("4 ALU only", 4, p.multiProcessorCount, w, mhz);
        run<5>("4 ALU + 1 IMAD", 5, p.multiProcessorCount, w, mhz);
        run<6>("4 ALU + 2 IMAD", 6, p.multiProcessorCount, w, mhz);
        run<3>("3 IMAD only", 3, p.multiProcessorCount, w, mhz);
        run<13>("TAG cell, plain adds (compiler's choice)", 7, p.multiProcessorCount, w, mhz);
        run<14>("TAG cell, mad.lo imm 1", 7, p.multiProcessorCount, w, mhz);
        run<4>("classic cell: 6 ALU + 2 IMAD", 8, p.multiProcessorCount, w, mhz);
        run<8>("2 ALU + 2 IMAD", 4, p.multiProcessorCount, w, mhz);
        run<9>("2 VIMNMX3", 2, p.multiProcessorCount, w, mhz);
        run<10>("VIADDMNMX", 1, p.multiProcessorCount, w, mhz);
        run<11>("LOP3", 1, p.multiProcessorCount, w, mhz);
        run<12>("IMAD", 1, p.multiProcessorCount, w, mhz);
    }
    return 0;
}
