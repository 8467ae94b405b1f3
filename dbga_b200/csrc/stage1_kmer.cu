// Stage 1: k-mer extraction and hashing of both sequences of every pair.
//
// Replaces, for a whole batch, the reference's get_kmers / duplicate_kmers
// (src/dbga/utils.py:108-154) and the dict walk of _get_seq_kmer_indices / _add_node
// (src/dbga/debruijn_pairwise.py:78-204).  Output per pair: aof[q] = position in seq1 of
// the k-mer at seq2 position q when that k-mer occurs exactly once in each sequence (a
// "merge node" / anchor), else -1.  In graph mode also the per-position non-duplicate
// flags from which node ids follow by prefix sums.
//
// Two data paths:
//   small : one CTA per pair, sequences 2-bit packed into shared memory by 128-bit
//           loads, open-addressing table (linear probing) in shared memory;
//   large : table in global memory (L2 resident per wave of pairs), one thread per k-mer.
#include "kernels.cuh"

namespace dbga {

static constexpr uint32_t V_EMPTY = 0xFFFFFFFFu;
static constexpr uint32_t V_DUP = 0xFFFFFFFEu;
static constexpr uint16_t MEMO_NONE = 0xFFFFu;

// ---------------------------------------------------------------- packing helpers
// 16 ASCII bases -> one 32-bit word, base i at bits [2i, 2i+2).  `bad` is set when a byte
// inside [0, valid) is not one of ACGT.
__device__ __forceinline__ uint32_t pack16(const uint8_t *src, int64_t valid, bool aligned, bool &bad) {
    uint32_t w = 0;
    if (valid >= 16 && aligned) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src));
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint8_t c = (uint8_t)(q[t] >> (8 * b));
                bad |= !is_acgt(c);
                w |= (uint32_t)base_code(c) << (2 * (4 * t + b));
            }
        }
    } else {
        for (int i = 0; i < 16 && i < valid; ++i) {
            const uint8_t c = src[i];
            bad |= !is_acgt(c);
            w |= (uint32_t)base_code(c) << (2 * i);
        }
    }
    return w;
}

template <typename KeyT>
__device__ __forceinline__ KeyT kmer_at(const uint32_t *pk, int p, KeyT kmask);

template <>
__device__ __forceinline__ uint32_t kmer_at<uint32_t>(const uint32_t *pk, int p, uint32_t kmask) {
    const int w = p >> 4, sh = (p & 15) * 2;
    return __funnelshift_r(pk[w], pk[w + 1], sh) & kmask;
}
template <>
__device__ __forceinline__ uint64_t kmer_at<uint64_t>(const uint32_t *pk, int p, uint64_t kmask) {
    const int w = p >> 4, sh = (p & 15) * 2;
    const uint32_t a = pk[w], b = pk[w + 1], c = pk[w + 2];
    const uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
    return (((uint64_t)hi << 32) | lo) & kmask;
}

template <typename KeyT> struct KeyTraits;
template <> struct KeyTraits<uint32_t> {
    static constexpr uint32_t EMPTY = 0xFFFFFFFFu;
    static __device__ __forceinline__ uint32_t hash(uint32_t k) { return hash32(k); }
    static __device__ __forceinline__ uint32_t cas(uint32_t *a, uint32_t cmp, uint32_t v) { return atomicCAS(a, cmp, v); }
};
template <> struct KeyTraits<uint64_t> {
    static constexpr uint64_t EMPTY = ~0ULL;
    static __device__ __forceinline__ uint32_t hash(uint64_t k) { return hash64(k); }
    static __device__ __forceinline__ uint64_t cas(uint64_t *a, uint64_t cmp, uint64_t v) {
        return (uint64_t)atomicCAS(reinterpret_cast<unsigned long long *>(a), (unsigned long long)cmp, (unsigned long long)v);
    }
};

// find-or-insert with linear probing; returns the slot
template <typename KeyT>
__device__ __forceinline__ uint32_t tbl_insert(KeyT *keys, uint32_t mask, KeyT key) {
    uint32_t h = KeyTraits<KeyT>::hash(key) & mask;
    for (;;) {
        const KeyT cur = *reinterpret_cast<volatile KeyT *>(&keys[h]);
        if (cur == key) return h;
        if (cur == KeyTraits<KeyT>::EMPTY) {
            const KeyT old = KeyTraits<KeyT>::cas(&keys[h], KeyTraits<KeyT>::EMPTY, key);
            if (old == KeyTraits<KeyT>::EMPTY || old == key) return h;
        }
        h = (h + 1) & mask;
    }
}
// find only; returns mask+1 when absent
template <typename KeyT>
__device__ __forceinline__ uint32_t tbl_find(const KeyT *keys, uint32_t mask, KeyT key) {
    uint32_t h = KeyTraits<KeyT>::hash(key) & mask;
    for (;;) {
        const KeyT cur = keys[h];
        if (cur == key) return h;
        if (cur == KeyTraits<KeyT>::EMPTY) return mask + 1;
        h = (h + 1) & mask;
    }
}
// record an occurrence at position pos: first occurrence keeps pos, later ones mark DUP
__device__ __forceinline__ void occ_mark(uint32_t *v, uint32_t pos) {
    const uint32_t old = atomicCAS(v, V_EMPTY, pos);
    if (old != V_EMPTY) *reinterpret_cast<volatile uint32_t *>(v) = V_DUP;
}

// =============================================================== small path (one CTA/pair)
// Shared memory (dynamic): pk0[words] pk1[words] | keys[cap] | v0[cap] v1[cap] | memo[maxq]
template <typename KeyT, bool FULL>
__global__ void __launch_bounds__(256)
stage1_small_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs,
                    const int32_t *__restrict__ pair_list, int count, int k, uint32_t cap,
                    int words, int32_t *__restrict__ status, int32_t *__restrict__ aof,
                    uint8_t *__restrict__ nd0, uint8_t *__restrict__ nd1) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *pk0 = reinterpret_cast<uint32_t *>(smem);
    uint32_t *pk1 = pk0 + words;
    KeyT *keys = reinterpret_cast<KeyT *>(pk1 + words);
    uint32_t *v0 = reinterpret_cast<uint32_t *>(keys + cap);
    uint32_t *v1 = v0 + cap;
    uint16_t *memo = reinterpret_cast<uint16_t *>(v1 + cap);
    __shared__ int s_bad;

    const uint32_t mask = cap - 1;
    const KeyT kmask = (2 * k >= (int)(8 * sizeof(KeyT))) ? ~(KeyT)0 : (((KeyT)1 << (2 * k)) - 1);
    const int tid = threadIdx.x, T = blockDim.x;

    for (int it = blockIdx.x; it < count; it += gridDim.x) {
        const int pair = pair_list[it];
        const PairDesc pd = pairs[pair];
        if (tid == 0) s_bad = 0;
        if (k > pd.n0 || k > pd.n1) {          // utils.py:129-130
            if (tid == 0) status[pair] = DBGA_K_INVALID;
            continue;
        }
        __syncthreads();
        // ---- pack both sequences (coalesced 128-bit loads when 16-byte aligned)
        bool bad = false;
        {
            const uint8_t *s0 = seqs + pd.off0, *s1 = seqs + pd.off1;
            const bool al0 = (reinterpret_cast<uintptr_t>(s0) & 15) == 0;
            const bool al1 = (reinterpret_cast<uintptr_t>(s1) & 15) == 0;
            for (int w = tid; w < words; w += T) {
                const int64_t r0 = (int64_t)pd.n0 - 16 * w, r1 = (int64_t)pd.n1 - 16 * w;
                pk0[w] = r0 > 0 ? pack16(s0 + 16 * w, r0, al0, bad) : 0u;
                pk1[w] = r1 > 0 ? pack16(s1 + 16 * w, r1, al1, bad) : 0u;
            }
            // table init: keys = EMPTY, v0 = v1 = EMPTY (all-ones pattern)
            uint4 *t4 = reinterpret_cast<uint4 *>(keys);
            const int n4 = (int)((cap * (sizeof(KeyT) + 8)) / 16);
            for (int i = tid; i < n4; i += T) t4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        }
        if (bad) s_bad = 1;
        __syncthreads();
        if (s_bad) {
            if (tid == 0) status[pair] = DBGA_BAD_CHAR;
            __syncthreads();
            continue;
        }
        const int N0 = pd.n0 - k + 1, N1 = pd.n1 - k + 1;
        // ---- phase A: insert every k-mer of seq1
        for (int p = tid; p < N0; p += T) {
            const uint32_t h = tbl_insert<KeyT>(keys, mask, kmer_at<KeyT>(pk0, p, kmask));
            occ_mark(&v0[h], (uint32_t)p);
        }
        __syncthreads();
        // ---- phase B1: seq2 k-mers.  Batch mode only needs k-mers already in the table
        //      (an anchor must occur in seq1); graph mode inserts all of them so that
        //      duplicates private to seq2 are seen too (utils.py:148-154).
        for (int q = tid; q < N1; q += T) {
            const KeyT key = kmer_at<KeyT>(pk1, q, kmask);
            uint32_t h;
            if (FULL) {
                h = tbl_insert<KeyT>(keys, mask, key);
                occ_mark(&v1[h], (uint32_t)q);
            } else {
                h = tbl_find<KeyT>(keys, mask, key);
                if (h <= mask) {
                    if (v0[h] < V_DUP) occ_mark(&v1[h], (uint32_t)q);
                    else h = mask + 1;
                }
            }
            memo[q] = h <= mask ? (uint16_t)h : MEMO_NONE;
        }
        __syncthreads();
        // ---- phase B2: anchors = k-mers seen exactly once in each sequence
        int32_t *aof_p = aof + pd.q_base;
        for (int q = tid; q < N1; q += T) {
            const uint16_t h = memo[q];
            int32_t a = -1;
            if (h != MEMO_NONE) {
                const uint32_t a0 = v0[h], b1 = v1[h];
                if (a0 < V_DUP && b1 == (uint32_t)q) a = (int32_t)a0;
                if (FULL) nd1[pd.q_base + q] = !(a0 == V_DUP || b1 == V_DUP);
            }
            aof_p[q] = a;
        }
        if (FULL) {
            for (int p = tid; p < N0; p += T) {
                const uint32_t h = tbl_find<KeyT>(keys, mask, kmer_at<KeyT>(pk0, p, kmask));
                nd0[pd.p_base + p] = !(v0[h] == V_DUP || v1[h] == V_DUP);
            }
        }
        __syncthreads();
    }
}

size_t stage1_small_smem(int key_bytes, uint32_t cap, int words, int maxq) {
    return (size_t)words * 8 + (size_t)cap * (key_bytes + 8) + (size_t)((maxq + 7) / 8 * 8) * 2;
}

template <typename KeyT, bool FULL>
static cudaError_t launch_small_t(const uint8_t *seqs, const PairDesc *pairs, const int32_t *list, int count,
                                  int k, uint32_t cap, int words, int maxq, int32_t *status, int32_t *aof,
                                  uint8_t *nd0, uint8_t *nd1, int num_sms, cudaStream_t st) {
    const size_t smem = stage1_small_smem(sizeof(KeyT), cap, words, maxq);
    auto kern = stage1_small_kernel<KeyT, FULL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    int grid = num_sms * occ;
    if (grid > count) grid = count;
    kern<<<grid, 256, smem, st>>>(seqs, pairs, list, count, k, cap, words, status, aof, nd0, nd1);
    return cudaGetLastError();
}

cudaError_t launch_stage1_small(const uint8_t *seqs, const PairDesc *pairs, const int32_t *list, int count,
                                int k, uint32_t cap, int words, int maxq, bool full, int32_t *status,
                                int32_t *aof, uint8_t *nd0, uint8_t *nd1, int num_sms, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    if (k <= 15) {
        return full ? launch_small_t<uint32_t, true>(seqs, pairs, list, count, k, cap, words, maxq, status, aof, nd0, nd1, num_sms, st)
                    : launch_small_t<uint32_t, false>(seqs, pairs, list, count, k, cap, words, maxq, status, aof, nd0, nd1, num_sms, st);
    }
    return full ? launch_small_t<uint64_t, true>(seqs, pairs, list, count, k, cap, words, maxq, status, aof, nd0, nd1, num_sms, st)
                : launch_small_t<uint64_t, false>(seqs, pairs, list, count, k, cap, words, maxq, status, aof, nd0, nd1, num_sms, st);
}

// =============================================================== large path (global table)
struct __align__(16) GSlot {
    uint64_t key;
    uint32_t v0, v1;
};

// grid: (ceil(max_words / 256), 2 * count).  Packs ASCII into 2-bit words in HBM.
__global__ void __launch_bounds__(256)
pack_large_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs,
                  const LargeDesc *__restrict__ ld, const int32_t *__restrict__ pair_list, int k,
                  uint32_t *__restrict__ packed, int32_t *__restrict__ status) {
    const int it = blockIdx.y >> 1, sq = blockIdx.y & 1;
    const int pair = pair_list[it];
    const PairDesc pd = pairs[pair];
    const int n = sq ? pd.n1 : pd.n0;
    const int words = (n + 15) / 16 + 3;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    if (k > pd.n0 || k > pd.n1) {
        if (w == 0 && sq == 0) status[pair] = DBGA_K_INVALID;
        return;
    }
    const uint8_t *s = seqs + (sq ? pd.off1 : pd.off0);
    const bool al = (reinterpret_cast<uintptr_t>(s) & 15) == 0;
    bool bad = false;
    const int64_t r = (int64_t)n - 16 * (int64_t)w;
    const uint32_t word = r > 0 ? pack16(s + 16 * (int64_t)w, r, al, bad) : 0u;
    packed[(sq ? ld[it].pk1 : ld[it].pk0) + w] = word;
    if (bad) status[pair] = DBGA_BAD_CHAR;
}

__device__ __forceinline__ uint32_t g_insert(GSlot *tbl, uint32_t mask, uint64_t key) {
    uint32_t h = hash64(key) & mask;
    for (;;) {
        const uint64_t cur = *reinterpret_cast<volatile uint64_t *>(&tbl[h].key);
        if (cur == key) return h;
        if (cur == ~0ULL) {
            const uint64_t old = (uint64_t)atomicCAS(reinterpret_cast<unsigned long long *>(&tbl[h].key), ~0ULL, (unsigned long long)key);
            if (old == ~0ULL || old == key) return h;
        }
        h = (h + 1) & mask;
    }
}
__device__ __forceinline__ uint32_t g_find(const GSlot *tbl, uint32_t mask, uint64_t key) {
    uint32_t h = hash64(key) & mask;
    for (;;) {
        const uint64_t cur = tbl[h].key;
        if (cur == key) return h;
        if (cur == ~0ULL) return 0xFFFFFFFFu;
        h = (h + 1) & mask;
    }
}

// phase: 0 = insert seq1; 1 = seq2 lookup/claim (memo into aof); 2 = finalize aof (+flags)
template <int PHASE, bool FULL>
__global__ void __launch_bounds__(256)
hash_large_kernel(const PairDesc *__restrict__ pairs, const LargeDesc *__restrict__ ld,
                  const int32_t *__restrict__ pair_list, int k, const uint32_t *__restrict__ packed,
                  GSlot *__restrict__ table, const int32_t *__restrict__ status,
                  int32_t *__restrict__ aof, uint8_t *__restrict__ nd0, uint8_t *__restrict__ nd1) {
    const int it = blockIdx.y;
    const int pair = pair_list[it];
    if (status[pair] != DBGA_OK) return;
    const PairDesc pd = pairs[pair];
    const LargeDesc L = ld[it];
    const int N0 = pd.n0 - k + 1, N1 = pd.n1 - k + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t kmask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    GSlot *tbl = table + L.tbl;
    if (PHASE == 0) {
        if (i >= N0) return;
        const uint32_t h = g_insert(tbl, L.mask, kmer_at<uint64_t>(packed + L.pk0, i, kmask));
        occ_mark(&tbl[h].v0, (uint32_t)i);
    } else if (PHASE == 1) {
        if (i >= N1) return;
        const uint64_t key = kmer_at<uint64_t>(packed + L.pk1, i, kmask);
        uint32_t h;
        if (FULL) {
            h = g_insert(tbl, L.mask, key);
            occ_mark(&tbl[h].v1, (uint32_t)i);
        } else {
            h = g_find(tbl, L.mask, key);
            if (h != 0xFFFFFFFFu) {
                if (tbl[h].v0 < V_DUP) occ_mark(&tbl[h].v1, (uint32_t)i);
                else h = 0xFFFFFFFFu;
            }
        }
        aof[pd.q_base + i] = (int32_t)h;
    } else {
        if (i < N1) {
            const uint32_t h = (uint32_t)aof[pd.q_base + i];
            int32_t a = -1;
            if (h != 0xFFFFFFFFu) {
                const uint32_t a0 = tbl[h].v0, b1 = tbl[h].v1;
                if (a0 < V_DUP && b1 == (uint32_t)i) a = (int32_t)a0;
                if (FULL) nd1[pd.q_base + i] = !(a0 == V_DUP || b1 == V_DUP);
            }
            aof[pd.q_base + i] = a;
        }
        if (FULL && i < N0) {
            const uint32_t h = g_find(tbl, L.mask, kmer_at<uint64_t>(packed + L.pk0, i, kmask));
            nd0[pd.p_base + i] = !(tbl[h].v0 == V_DUP || tbl[h].v1 == V_DUP);
        }
    }
}

cudaError_t launch_stage1_large(const uint8_t *seqs, const PairDesc *pairs, const LargeDesc *ld,
                                const int32_t *list, int count, int max_n, int k, bool full,
                                uint32_t *packed, void *table, size_t table_bytes, int32_t *status,
                                int32_t *aof, uint8_t *nd0, uint8_t *nd1, cudaStream_t st, int *launches) {
    if (count <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(table, 0xFF, table_bytes, st);
    if (e != cudaSuccess) return e;
    const int words = (max_n + 15) / 16 + 3;
    dim3 gp((words + 255) / 256, 2 * count), gk((max_n + 255) / 256, count);
    pack_large_kernel<<<gp, 256, 0, st>>>(seqs, pairs, ld, list, k, packed, status);
    GSlot *tbl = reinterpret_cast<GSlot *>(table);
    if (full) {
        hash_large_kernel<0, true><<<gk, 256, 0, st>>>(pairs, ld, list, k, packed, tbl, status, aof, nd0, nd1);
        hash_large_kernel<1, true><<<gk, 256, 0, st>>>(pairs, ld, list, k, packed, tbl, status, aof, nd0, nd1);
        hash_large_kernel<2, true><<<gk, 256, 0, st>>>(pairs, ld, list, k, packed, tbl, status, aof, nd0, nd1);
    } else {
        hash_large_kernel<0, false><<<gk, 256, 0, st>>>(pairs, ld, list, k, packed, tbl, status, aof, nd0, nd1);
        hash_large_kernel<1, false><<<gk, 256, 0, st>>>(pairs, ld, list, k, packed, tbl, status, aof, nd0, nd1);
        hash_large_kernel<2, false><<<gk, 256, 0, st>>>(pairs, ld, list, k, packed, tbl, status, aof, nd0, nd1);
    }
    *launches += 4;
    return cudaGetLastError();
}

}  // namespace dbga
