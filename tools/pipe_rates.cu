// Issue-rate microbenchmark for the instructions of the packed 16-bit DP cell (sm_100a):
// which of them share the ALU pipe (16 lanes / clk / SM sub-partition) and which run on the
// FMA pipes.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pipe_rates tools/pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP9(S) S(0) S(1) S(2) S(3) S(4) S(5) S(6) S(7)
template <int W>
__global__ void __launch_bounds__(256) k(int iters, uint32_t *sink) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 77u + i * 1234567u;
    const uint32_t b = blockIdx.x | 0x00010001u, c = threadIdx.x | 0x00030003u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (W == 0) a[i] = __vimax3_s16x2(a[i], b, c);
                else if (W == 1) a[i] = __viaddmax_s16x2(a[i], b, c);
                else if (W == 2) a[i] = __vadd2(a[i], b);
                else if (W == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                else if (W == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                else if (W == 5) asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(a[i]) : "r"(b));
                else if (W == 6) { if (i & 1) asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(a[i]) : "r"(b));
                                   else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
                else if (W == 7) asm volatile("shl.b32 %0, %0, 2;" : "+r"(a[i]));
                else if (W == 8) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                else if (W == 9) { if (i & 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                                   else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
                else if (W == 10) a[i] = __vimax3_s32(a[i], b, c);
                else if (W == 11) { if (i % 3 == 2) asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(a[i]) : "r"(b));
                                    else a[i] = __viaddmax_s16x2(a[i], b, c); }
                else if (W == 12) a[i] = __vmaxs2(a[i], b);
                else if (W == 13) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(a[i]) : "r"(b));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    if (r == 0x12345678u) sink[0] = r;
}

template <int W>
double run(const char *name, uint32_t *sink, int sms) {
    const int iters = 2048, grid = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); k<W><<<grid, 256>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double ops = (double)grid * 256 * iters * 16 * 8;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double per_clk_sm = ops / (best * 1e-3) / sms / (clk * 1e3);
    printf("{\"op\": \"%s\", \"thread_ops_per_s\": %.4g, \"per_clk_per_sm_at_max_clock\": %.1f}\n", name, ops / (best * 1e-3), per_clk_sm);
    return 0;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *sink; cudaMalloc(&sink, 256);
    run<0>("VIMNMX3.S16x2", sink, sms);
    run<1>("VIADDMNMX.S16x2", sink, sms);
    run<2>("VIADD.16x2", sink, sms);
    run<12>("VIMNMX.S16x2", sink, sms);
    run<10>("VIMNMX3.S32", sink, sms);
    run<3>("LOP3", sink, sms);
    run<4>("PRMT", sink, sms);
    run<5>("IMAD (x*5+b)", sink, sms);
    run<7>("SHL", sink, sms);
    run<13>("SHF", sink, sms);
    run<8>("IADD", sink, sms);
    run<6>("LOP3 + IMAD alternating", sink, sms);
    run<9>("LOP3 + IADD alternating", sink, sms);
    run<11>("2 VIADDMNMX.S16x2 : 1 IMAD", sink, sms);
    return 0;
}
