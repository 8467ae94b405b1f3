// Integer-pipe microbenchmark: the denominator of the DP roofline (SURVEY.md 8d says the
// int32 add/max rate must be measured on the box, it is not in MEASURED_PEAKS.json).
// which: 0 = add only, 1 = max only, 2 = the recurrence's 5 add : 4 max mix.
#include "kernels.cuh"

namespace dbga {

template <int WHICH>
__global__ void __launch_bounds__(256) int_peak_kernel(int iters, int *sink) {
    int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
        a7 = a0 + 7, a8 = a0 + 8;
    const int b = blockIdx.x | 1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (WHICH == 0) {
                asm volatile("add.s32 %0, %0, %9; add.s32 %1, %1, %9; add.s32 %2, %2, %9; add.s32 %3, %3, %9;"
                             "add.s32 %4, %4, %9; add.s32 %5, %5, %9; add.s32 %6, %6, %9; add.s32 %7, %7, %9;"
                             "add.s32 %8, %8, %9;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "+r"(a8)
                             : "r"(b));
            } else if (WHICH == 1) {
                asm volatile("max.s32 %0, %0, %1; max.s32 %1, %1, %2; max.s32 %2, %2, %3; max.s32 %3, %3, %4;"
                             "max.s32 %4, %4, %5; max.s32 %5, %5, %6; max.s32 %6, %6, %7; max.s32 %7, %7, %8;"
                             "max.s32 %8, %8, %9;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "+r"(a8)
                             : "r"(b));
            } else {
                asm volatile("add.s32 %0, %0, %9; max.s32 %1, %1, %0; add.s32 %2, %2, %9; max.s32 %3, %3, %2;"
                             "add.s32 %4, %4, %9; max.s32 %5, %5, %4; add.s32 %6, %6, %9; max.s32 %7, %7, %6;"
                             "add.s32 %8, %8, %9;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "+r"(a8)
                             : "r"(b));
            }
        }
    }
    const int r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ a8;
    if (r == 0x7fffffff) sink[0] = r;
}

// each thread executes iters * 16 * 9 integer ops
cudaError_t launch_int_peak(int which, int iters, int *sink, int num_sms, cudaStream_t st) {
    const int grid = num_sms * 8;
    if (which == 0) int_peak_kernel<0><<<grid, 256, 0, st>>>(iters, sink);
    else if (which == 1) int_peak_kernel<1><<<grid, 256, 0, st>>>(iters, sink);
    else int_peak_kernel<2><<<grid, 256, 0, st>>>(iters, sink);
    return cudaGetLastError();
}

}  // namespace dbga
