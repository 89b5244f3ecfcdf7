// int_peak.cuh -- integer / DPX issue-rate microbenchmark.
//
// MEASURED_PEAKS.json has no integer entry, so the roofline
// denominator for the alignment kernels is measured here, on the same GPU and in
// the same process as the numbers it normalises: independent dependency chains of
// one instruction kind (or of the kernel's own 8-instruction cell) on every SM, with
// enough warps and chains per thread that only the issue/pipe rate limits them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsa {

constexpr int kPeakChains = 8;
constexpr int kPeakIters = 4096;

template <int WHICH>
__global__ void __launch_bounds__(256) int_peak_kernel(int seed, int one, int one2, int* sink, unsigned long long* clk) {
    int a[kPeakChains], b[kPeakChains], c[kPeakChains];
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) {
        a[i] = seed + threadIdx.x * 7 + i;
        b[i] = seed * 3 + i * 5 + 1 + threadIdx.x;
        c[i] = seed - i - 3 * threadIdx.x;
    }
    const int ge = seed | 1, go = seed + 3, mask = ~(3 << 12), ph = 2 << 12, pv = 1 << 12;
    unsigned dirA[2] = {(unsigned)seed, 0u}, dirF[2] = {0u, (unsigned)seed};   // WHICH == 9: direction accumulators
    unsigned long long c0 = 0, t0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        c0 = clock64();
    }
#pragma unroll 1
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) {
            if (WHICH == 0) {
                // the alignment cell: 3 LOP3 + VIMNMX3 + 2 VIADDMNMX (ALU pipe) + 2 IMAD (FMA pipe)
                const int e = a[i] | ph;
                const int f = b[i] | pv;
                const int d = c[i] * one + ge;
                const int h = __vimax3_s32(d, e, f);
                const int hc = h & mask;
                const int hg = hc * one + go;
                a[i] = __viaddmax_s32(e, ge, hg);
                b[i] = __viaddmax_s32(f, ge, hg);
                c[i] = hc;
            } else if (WHICH == 7) {
                // the TAG cell (gotoh_kernels.cuh, cell_row<TAG>): VIMNMX3 + LOP3 + 2 VIADDMNMX (ALU pipe) + 3 IMAD
                const int d = c[i] * one + ge;
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(a[i], ge, hc * one + go);
                b[i] = __viaddmax_s32(b[i], ge, hc * one2 + ph);
                c[i] = hc;
            } else if (WHICH == 8) {
                // the FRAME cell (TAG mode since round 2): VIMNMX3 + LOP3 + 2 VIADDMNMX (ALU pipe) + 2 IMAD
                const int d = c[i] * one + ge;
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(a[i], ge, hc * one2 + go);
                b[i] = __viaddmax_s32(hc, ph, b[i]);
                c[i] = hc;
            } else if (WHICH == 10) {
                // the FRAME cell with column-tagged E openings (BSA_ETAG): VIMNMX3 + LOP3 + 2 VIADDMNMX (ALU pipe) + 1 IMAD
                const int d = c[i] * one + ge;
                const int h = __vimax3_s32(d, a[i], b[i]);
                const int hc = h & mask;
                a[i] = __viaddmax_s32(hc, go, a[i]);
                b[i] = __viaddmax_s32(hc, ph, b[i]);
                c[i] = hc;
            } else if (WHICH == 9) {
                // the K3 direction-frame cell (wave_kernels.cuh): VIMNMX3 + 4 LOP3 + 2 VIADDMNMX + SHF (ALU pipe) + 3 IMAD
                const int d = c[i] * one + ge;
                const int h = __vimax3_s32(d, a[i], b[i]);
                unsigned x;
                asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(x) : "r"(h), "r"(a[i]), "r"(mask));
                dirA[i & 1] = __funnelshift_r(dirA[i & 1], x, 3);
                const int f1 = b[i] | 1;
                dirF[i & 1] = (unsigned)(b[i] * one2 + (int)(dirF[i & 1] * (unsigned)go + (unsigned)f1));
                const int hc = h & ~7;
                a[i] = __viaddmax_s32(hc, ph, a[i] | 1);
                b[i] = __viaddmax_s32(hc, pv, f1);
                c[i] = hc;
            } else if (WHICH == 1) {
                a[i] = __viaddmax_s32(a[i], ge, b[i]);
            } else if (WHICH == 2) {
                a[i] = __vimax3_s32(a[i], b[i], c[i]);
                b[i] ^= it;   // keep the chain from collapsing; counted below
            } else if (WHICH == 3) {
                a[i] = (a[i] & mask) | (b[i] ^ it);   // one LOP3
            } else if (WHICH == 4) {
                a[i] = a[i] + b[i] + it;              // one IADD3
            } else if (WHICH == 5) {
                a[i] = a[i] * ge + b[i];              // one IMAD
            } else {
                a[i] = (int)__viaddmax_s16x2((unsigned)a[i], (unsigned)ge, (unsigned)b[i]);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long c1 = clock64(), t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        clk[0] = c1 - c0;
        clk[1] = t1 - t0;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) s += a[i] ^ b[i] ^ c[i];
    if (WHICH == 9) s += (int)(dirA[0] ^ dirA[1] ^ dirF[0] ^ dirF[1]);
    if (s == 0x7fffffff) *sink = s;
}

inline int peak_ops_per_iter(int which) {
    switch (which) {
        case 0: return 8;
        case 7: return 7;
        case 8: return 6;
        case 9: return 11;
        case 10: return 5;
        case 2: return 2;   // VIMNMX3 + the LOP3 that perturbs it
        default: return 1;
    }
}

inline cudaError_t measure_int_peak(int which, int sms, cudaStream_t st, double* lane_ops_per_s,
                                    double* sm_mhz) {
    int* sink = nullptr;
    unsigned long long* clk = nullptr;
    cudaError_t e = cudaMalloc(&sink, 4);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&clk, 16);
    if (e != cudaSuccess) { cudaFree(sink); return e; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 8;
    auto go = [&](void) {
        switch (which) {
            case 0: int_peak_kernel<0><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 1: int_peak_kernel<1><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 2: int_peak_kernel<2><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 3: int_peak_kernel<3><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 4: int_peak_kernel<4><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 5: int_peak_kernel<5><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 7: int_peak_kernel<7><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 8: int_peak_kernel<8><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 9: int_peak_kernel<9><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            case 10: int_peak_kernel<10><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
            default: int_peak_kernel<6><<<blocks, 256, 0, st>>>(1, 1, 1, sink, clk); break;
        }
    };
    for (int w = 0; w < 3; ++w) go();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0, st);
        go();
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    unsigned long long h[2] = {0, 1};
    if (e == cudaSuccess) e = cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
        const double lane_ops = (double)blocks * 256.0 * kPeakIters * kPeakChains * peak_ops_per_iter(which);
        *lane_ops_per_s = lane_ops / (best * 1e-3);
        *sm_mhz = h[1] ? (double)h[0] / (double)h[1] * 1e3 : 0.0;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(clk);
    return e;
}

}  // namespace bsa
