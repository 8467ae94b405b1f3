// k-mer set statistics of sequence pairs: |K0|, |K1| (distinct k-mers of each sequence) and
// |K0 & K1|.  These are what the reference's k heuristics are made of:
//   predict_p(seqs, k) = (|K0 & K1| / |K0 | K1|) ** (1 / k)                src/dbga/utils.py:371-395
//   predict_final_p    = min of predict_p over a range of k                 src/dbga/utils.py:398-413
//   kmers_larger_than_threshold: (N - |K|) / N > threshold                  src/dbga/kmersize.py:10-14
// One CTA per pair, the same packing and bucketized shared-memory table as the fused stage 1+2
// kernel (kmer_table.cuh); one flag word per slot (bit 0: in seq1, bit 1: in seq2).
#include "kmer_table.cuh"

namespace dbga {

struct KsShared {
    int d0, d1, shared;
};

template <typename KeyT>
__global__ void __launch_bounds__(256)
kmer_stats_kernel(const uint8_t *__restrict__ seqs, const int64_t *__restrict__ offsets, int64_t off_base, int64_t count,
                  int k, KeyT kmask, uint32_t cap, uint32_t nb, int words, int64_t max_keys, int64_t max_n,
                  int32_t *__restrict__ status, int32_t *__restrict__ out /* 3 per pair */) {
    extern __shared__ __align__(16) uint8_t smem[];
    using B = Bucket<KeyT>;
    uint32_t *pk0 = reinterpret_cast<uint32_t *>(smem);
    uint32_t *pk1 = pk0 + words;
    KeyT *keys = reinterpret_cast<KeyT *>(pk1 + words);
    uint32_t *flag = reinterpret_cast<uint32_t *>(keys + cap);
    __shared__ KsShared sh;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31;

    for (int64_t pair = blockIdx.x; pair < count; pair += gridDim.x) {
        const int64_t o0 = offsets[2 * pair] - off_base, o1 = offsets[2 * pair + 1] - off_base, o2 = offsets[2 * pair + 2] - off_base;
        // a pair beyond the launch's table / packing size is counted by the global-memory kernel
        if (o1 - o0 > max_n || o2 - o1 > max_n || (o1 - o0 >= k && o2 - o1 >= k && (o1 - o0 - k + 1) + (o2 - o1 - k + 1) > max_keys)) continue;
        const int n0 = (int)(o1 - o0), n1 = (int)(o2 - o1);
        __syncthreads();                         // previous pair's shared state is no longer in use
        if (tid == 0) { sh.d0 = 0; sh.d1 = 0; sh.shared = 0; }
        if (k > n0 || k > n1) {                  // utils.py:129-130
            if (tid == 0) { status[pair] = DBGA_K_INVALID; out[3 * pair] = out[3 * pair + 1] = out[3 * pair + 2] = 0; }
            continue;
        }
        bool bad = false;
        for (int w = tid; w < words; w += T) {
            const int64_t r0 = (int64_t)n0 - 16 * w, r1 = (int64_t)n1 - 16 * w;
            pk0[w] = r0 > 0 ? pack16_fast(seqs + o0 + 16 * w, r0, bad) : 0u;
            pk1[w] = r1 > 0 ? pack16_fast(seqs + o1 + 16 * w, r1, bad) : 0u;
        }
        {
            uint4 *t4 = reinterpret_cast<uint4 *>(keys);
            const int n_ff = (int)((cap * sizeof(KeyT)) / 16), n_00 = (int)(cap * 4 / 16);
            for (int i = tid; i < n_ff; i += T) t4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
            uint4 *z4 = reinterpret_cast<uint4 *>(flag);
            for (int i = tid; i < n_00; i += T) z4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (__syncthreads_or(bad)) {
            if (tid == 0) { status[pair] = DBGA_BAD_CHAR; out[3 * pair] = out[3 * pair + 1] = out[3 * pair + 2] = 0; }
            continue;
        }
        const int N0 = n0 - k + 1, N1 = n1 - k + 1;
        int c0 = 0, c1 = 0, cs = 0;
        for (int p = tid; p < N0; p += T) {
            const KeyT key = B::kmer(pk0, p, kmask);
            bool claimed;
            const uint32_t h = bkt_insert<KeyT>(keys, nb, key, B::mix(key), claimed);
            if (claimed) { flag[h] = 1u; ++c0; }         // the claimer is the only writer in this phase
        }
        __syncthreads();
        for (int q = tid; q < N1; q += T) {
            const KeyT key = B::kmer(pk1, q, kmask);
            bool claimed;
            const uint32_t h = bkt_insert<KeyT>(keys, nb, key, B::mix(key), claimed);
            const uint32_t old = atomicOr(&flag[h], 2u);
            if (!(old & 2u)) { ++c1; cs += (int)(old & 1u); }   // first occurrence in seq2
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            cs += __shfl_xor_sync(0xffffffffu, cs, o);
        }
        if (lane == 0) { atomicAdd(&sh.d0, c0); atomicAdd(&sh.d1, c1); atomicAdd(&sh.shared, cs); }
        __syncthreads();
        if (tid == 0) { status[pair] = DBGA_OK; out[3 * pair] = sh.d0; out[3 * pair + 1] = sh.d1; out[3 * pair + 2] = sh.shared; }
    }
}

// Pairs whose table does not fit shared memory: the whole grid works on ONE pair, open
// addressing with linear probing in a global-memory table of 64-bit keys (k <= 31 always fits),
// k-mers read straight from the ASCII bases.  phase 0: seq1 (claim -> |K0|), phase 1: seq2.
// counts[0..2] = |K0|, |K1|, |K0 & K1|; counts[3] != 0: a character outside ACGT was seen.
__global__ void __launch_bounds__(256)
kmer_stats_large_kernel(const uint8_t *__restrict__ seq, int64_t n, int k, unsigned long long kmask, int phase,
                        unsigned long long *__restrict__ keys, uint32_t *__restrict__ flag, uint64_t cap,
                        unsigned int *__restrict__ counts) {
    const int64_t N = n - k + 1;
    int c_new = 0, c_sh = 0, bad = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long key = 0;
        for (int j = 0; j < k; ++j) {
            const uint8_t ch = seq[p + j];
            bad |= !is_acgt(ch);
            key |= (unsigned long long)base_code(ch) << (2 * j);
        }
        key &= kmask;
        uint64_t h = ((key * 0x9E3779B97F4A7C15ULL) >> 20) % cap;
        for (;;) {
            const unsigned long long old = atomicCAS(&keys[h], ~0ULL, key);
            if (old == ~0ULL || old == key) break;
            h = h + 1 == cap ? 0 : h + 1;
        }
        const uint32_t old = atomicOr(&flag[h], phase == 0 ? 1u : 2u);
        if (phase == 0) c_new += !(old & 1u);
        else if (!(old & 2u)) { ++c_new; c_sh += (int)(old & 1u); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_new += __shfl_xor_sync(0xffffffffu, c_new, o);
        c_sh += __shfl_xor_sync(0xffffffffu, c_sh, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c_new) atomicAdd(&counts[phase], (unsigned)c_new);
        if (c_sh) atomicAdd(&counts[2], (unsigned)c_sh);
        if (bad) atomicOr(&counts[3], 1u);
    }
}
// k >= 32: the k-mer does not fit the 64-bit key, so a slot names it by reference -- hash tag (high
// 32 bits) | position of the claiming occurrence (bit 31: it is a seq2 position) -- and a probe with a
// matching tag compares the k bases of the two occurrences (exact for every k; the reference's
// predict_final_p scans k up to ceil(log2(shortest)) + 9, src/dbga/utils.py:395-413).
__global__ void __launch_bounds__(256)
kmer_stats_wide_kernel(const uint8_t *__restrict__ seq0, const uint8_t *__restrict__ seq1, int64_t n, int k, int phase,
                       unsigned long long *__restrict__ keys, uint32_t *__restrict__ flag, uint64_t cap,
                       unsigned int *__restrict__ counts) {
    const uint8_t *seq = phase ? seq1 : seq0;
    const int64_t N = n - k + 1;
    int c_new = 0, c_sh = 0, bad = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long hv = 0x9E3779B97F4A7C15ULL;
        for (int j = 0; j < k; ++j) {
            const uint8_t ch = seq[p + j];
            bad |= !is_acgt(ch);
            hv = (hv ^ (unsigned long long)base_code(ch)) * 0x100000001B3ULL;
        }
        hv ^= hv >> 33; hv *= 0xff51afd7ed558ccdULL; hv ^= hv >> 33;
        const unsigned long long mine = (hv & 0xFFFFFFFF00000000ULL) | (phase ? 0x80000000ULL : 0ULL) | (unsigned long long)p;
        uint64_t h = (hv & 0xFFFFFFFFULL) % cap;
        for (;;) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&keys[h]);
            if (cur == ~0ULL) {
                cur = atomicCAS(&keys[h], ~0ULL, mine);
                if (cur == ~0ULL) break;
            }
            if ((cur >> 32) == (hv >> 32)) {
                const uint8_t *other = ((cur & 0x80000000ULL) ? seq1 : seq0) + (cur & 0x7FFFFFFFULL);
                bool same = true;
                for (int j = 0; j < k && same; ++j) same = other[j] == seq[p + j];
                if (same) break;
            }
            h = h + 1 == cap ? 0 : h + 1;
        }
        const uint32_t old = atomicOr(&flag[h], phase == 0 ? 1u : 2u);
        if (phase == 0) c_new += !(old & 1u);
        else if (!(old & 2u)) { ++c_new; c_sh += (int)(old & 1u); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_new += __shfl_xor_sync(0xffffffffu, c_new, o);
        c_sh += __shfl_xor_sync(0xffffffffu, c_sh, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c_new) atomicAdd(&counts[phase], (unsigned)c_new);
        if (c_sh) atomicAdd(&counts[2], (unsigned)c_sh);
        if (bad) atomicOr(&counts[3], 1u);
    }
}
__global__ void kmer_stats_large_init(unsigned long long *keys, uint32_t *flag, uint64_t cap, unsigned int *counts) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        keys[i] = ~0ULL; flag[i] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) counts[threadIdx.x] = 0u;
}
// one pair: table of `cap` slots (keys: 8 bytes, flag: 4 bytes each), counts: 4 words
cudaError_t launch_kmer_stats_large(const uint8_t *seq0, int64_t n0, const uint8_t *seq1, int64_t n1, int k,
                                    unsigned long long *keys, uint32_t *flag, uint64_t cap, unsigned int *counts,
                                    int num_sms, cudaStream_t st) {
    const unsigned long long kmask = (2 * k >= 64) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
    kmer_stats_large_init<<<num_sms * 4, 256, 0, st>>>(keys, flag, cap, counts);
    if (k >= 32) {
        kmer_stats_wide_kernel<<<num_sms * 8, 256, 0, st>>>(seq0, seq1, n0, k, 0, keys, flag, cap, counts);
        kmer_stats_wide_kernel<<<num_sms * 8, 256, 0, st>>>(seq0, seq1, n1, k, 1, keys, flag, cap, counts);
        return cudaGetLastError();
    }
    kmer_stats_large_kernel<<<num_sms * 8, 256, 0, st>>>(seq0, n0, k, kmask, 0, keys, flag, cap, counts);
    kmer_stats_large_kernel<<<num_sms * 8, 256, 0, st>>>(seq1, n1, k, kmask, 1, keys, flag, cap, counts);
    return cudaGetLastError();
}

size_t kmer_stats_smem(int key_bytes, uint32_t cap, int words) { return (size_t)words * 8 + (size_t)cap * (key_bytes + 4); }

// cap: table slots (multiple of 16, >= 1.5 x max_keys); words: packed words for max_n bases; pairs with more
// k-mers than max_keys or a sequence longer than max_n are left alone
cudaError_t launch_kmer_stats(const uint8_t *seqs, const int64_t *offsets, int64_t off_base, int64_t count, int k,
                              uint32_t cap, int words, int64_t max_keys, int64_t max_n, int32_t *status, int32_t *out,
                              int num_sms, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    const int key_bytes = k <= 15 ? 4 : 8;
    const size_t smem = kmer_stats_smem(key_bytes, cap, words);
    int occ = 1;
    cudaError_t e;
    if (k <= 15) {
        auto kern = kmer_stats_kernel<uint32_t>;
        if ((e = allow_max_dyn_smem(kern)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem)) != cudaSuccess) return e;
        const int64_t grid = std::min<int64_t>(count, (int64_t)num_sms * std::max(occ, 1));
        const uint32_t kmask = (2 * k >= 32) ? ~0u : ((1u << (2 * k)) - 1u);
        kern<<<(int)grid, 256, smem, st>>>(seqs, offsets, off_base, count, k, kmask, cap, cap / 4u, words, max_keys, max_n, status, out);
    } else {
        auto kern = kmer_stats_kernel<uint64_t>;
        if ((e = allow_max_dyn_smem(kern)) != cudaSuccess) return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem)) != cudaSuccess) return e;
        const int64_t grid = std::min<int64_t>(count, (int64_t)num_sms * std::max(occ, 1));
        const uint64_t kmask = (2 * k >= 64) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
        kern<<<(int)grid, 256, smem, st>>>(seqs, offsets, off_base, count, k, kmask, cap, cap / 2u, words, max_keys, max_n, status, out);
    }
    return cudaGetLastError();
}

}  // namespace dbga
