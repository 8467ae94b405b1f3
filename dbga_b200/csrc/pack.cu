// 2-bit packed input (dbga_align_batch_packed): the host link carries a quarter of the bytes, and
// one pass over HBM turns the words back into the ASCII layout every other kernel reads.
//
// Canonical packed layout of a batch (include/dbga_b200.h): sequence j of the batch starts at 32-bit
// word (offsets[j] >> 4) + j; base i of it sits in bits 2 (i & 15) .. of word i >> 4; A C G T = 0 1 2 3
// (the loader's alphabet, reference src/dbga/utils.py:95-96 -> cogent3 DNA).
#include "kernels.cuh"

namespace dbga {

// 4 codes (8 bits) -> 4 ASCII bases: one PRMT over "ACGT" with the codes spread into nibbles
__device__ __forceinline__ uint32_t expand4(uint32_t c8) {
    const uint32_t sel = (c8 & 3u) | ((c8 & 0xCu) << 2) | ((c8 & 0x30u) << 4) | ((c8 & 0xC0u) << 6);
    return __byte_perm(0x54474341u, 0u, sel);
}

// One warp per sequence; a lane owns a 16-byte aligned chunk of the destination, takes the 16
// codes that land there from two consecutive packed words (funnel shift: any base offset) and
// stores 128 bits.  Head and tail chunks store bytes.
__global__ void __launch_bounds__(256)
unpack_kernel(const uint32_t *__restrict__ packed, const PairDesc *__restrict__ pairs, int64_t num_seqs,
              int64_t base, int64_t first_seq, int64_t word_base, uint8_t *__restrict__ seqs) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < num_seqs; j += nwarps) {
        const PairDesc pd = pairs[j >> 1];
        const int64_t off = (j & 1) ? pd.off1 : pd.off0;
        const int len = (j & 1) ? pd.n1 : pd.n0;
        const uint32_t *src = packed + (((off + base) >> 4) + first_seq + j - word_base);
        uint8_t *dst = seqs + off;
        const int head = (int)(reinterpret_cast<uintptr_t>(dst) & 15);          // dst - head is 16-byte aligned
        for (int i0 = 16 * lane - head; i0 < len; i0 += 512) {
            // bases i0 .. i0 + 15 (clipped to [0, len)) go to dst + i0
            const int ib = i0 < 0 ? 0 : i0;
            const int wq = ib >> 4, sh = 2 * (ib & 15);
            const uint32_t w0 = __ldg(src + wq);
            const uint32_t w1 = (sh != 0 && 16 * (wq + 1) < len) ? __ldg(src + wq + 1) : 0u;
            const uint32_t codes = __funnelshift_r(w0, w1, sh);                  // codes of bases ib .. ib + 15
            if (i0 >= 0 && i0 + 16 <= len) {
                *reinterpret_cast<uint4 *>(dst + i0) = make_uint4(expand4(codes & 0xFFu), expand4((codes >> 8) & 0xFFu),
                                                                 expand4((codes >> 16) & 0xFFu), expand4(codes >> 24));
            } else {
                const int ie = i0 + 16 < len ? i0 + 16 : len;
                for (int i = ib; i < ie; ++i) dst[i] = (uint8_t)(expand4((codes >> (2 * (i - ib))) & 3u) & 0xFFu);
            }
        }
    }
}

cudaError_t launch_unpack(const uint32_t *packed, const PairDesc *pairs, int64_t num_pairs, int64_t base,
                          int64_t first_seq, int64_t word_base, uint8_t *seqs, int num_sms, cudaStream_t st) {
    if (num_pairs <= 0) return cudaSuccess;
    int64_t grid = (2 * num_pairs + 7) / 8;
    if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
    unpack_kernel<<<(int)grid, 256, 0, st>>>(packed, pairs, 2 * num_pairs, base, first_seq, word_base, seqs);
    return cudaGetLastError();
}

}  // namespace dbga
