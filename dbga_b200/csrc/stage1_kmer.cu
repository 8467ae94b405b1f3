// Stage 1: k-mer extraction and hashing of both sequences of every pair.
//
// Replaces, for a whole batch, the reference's get_kmers / duplicate_kmers
// (src/dbga/utils.py:108-154) and the dict walk of _get_seq_kmer_indices / _add_node
// (src/dbga/debruijn_pairwise.py:78-204).  Output per pair: aof[q] = position in seq1 of
// the k-mer at seq2 position q when that k-mer occurs exactly once in each sequence (a
// "merge node" / anchor), else -1.  In graph mode also the per-position non-duplicate
// flags from which node ids follow by prefix sums.
//
// This file holds the path for pairs whose table does not fit in shared memory: table in
// global memory (L2 resident per wave of pairs), one thread per k-mer.  Pairs that fit use
// the fused kernel in stage12_small.cu.
#include "kernels.cuh"

namespace dbga {

static constexpr uint32_t V_EMPTY = 0xFFFFFFFFu;
static constexpr uint32_t V_DUP = 0xFFFFFFFEu;

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

// record an occurrence at position pos: first occurrence keeps pos, later ones mark DUP
__device__ __forceinline__ void occ_mark(uint32_t *v, uint32_t pos) {
    const uint32_t old = atomicCAS(v, V_EMPTY, pos);
    if (old != V_EMPTY) *reinterpret_cast<volatile uint32_t *>(v) = V_DUP;
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

// ---- k >= 32 ("wide" k-mers; the reference accepts any k <= len, utils.py:129-131, and its own
// front-ends reach k = 33 and beyond on sequences of more than 2^21 bases, adaptive_debruijn.py:54-56).
// A k-mer no longer fits the 64-bit key, so the slot names it by reference instead: key = hash tag
// (high 32 bits) | position of the occurrence that claimed the slot (seq1 position p, or
// 0x80000000 | q for a seq2 position).  A probe compares the tag and, on a tag match, the 2k bits
// of the two occurrences word by word in the packed sequences -- exact for every k.
struct WideCtx { const uint32_t *pk0, *pk1; int k; };
__device__ __forceinline__ uint32_t wide_word(const uint32_t *pk, int p, int j) {      // bases p + 16 j .. p + 16 j + 15
    const int w = (p >> 4) + j, sh = (p & 15) * 2;
    return __funnelshift_r(pk[w], pk[w + 1], sh);
}
__device__ __forceinline__ uint64_t wide_hash(const uint32_t *pk, int p, int k) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    const int full = k >> 4, rem = k & 15;
    for (int j = 0; j < full; ++j) { h ^= wide_word(pk, p, j); h *= 0xff51afd7ed558ccdULL; h ^= h >> 29; }
    if (rem) { h ^= wide_word(pk, p, full) & ((1u << (2 * rem)) - 1u); h *= 0xff51afd7ed558ccdULL; h ^= h >> 29; }
    h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 32;
    return h;
}
__device__ __forceinline__ bool wide_same(const WideCtx &c, uint32_t ra, uint32_t rb) {
    const uint32_t *pa = (ra & 0x80000000u) ? c.pk1 : c.pk0, *pb = (rb & 0x80000000u) ? c.pk1 : c.pk0;
    const int a = (int)(ra & 0x7FFFFFFFu), b = (int)(rb & 0x7FFFFFFFu);
    const int full = c.k >> 4, rem = c.k & 15;
    for (int j = 0; j < full; ++j) if (wide_word(pa, a, j) != wide_word(pb, b, j)) return false;
    if (rem) { const uint32_t m = (1u << (2 * rem)) - 1u; if ((wide_word(pa, a, full) ^ wide_word(pb, b, full)) & m) return false; }
    return true;
}
// find-or-claim / find for the occurrence `me` (a position code) whose k-mer hashes to hv
__device__ __forceinline__ uint32_t gw_insert(GSlot *tbl, uint32_t mask, const WideCtx &c, uint64_t hv, uint32_t me) {
    uint32_t h = (uint32_t)hv & mask;
    const uint64_t mine = (hv & 0xFFFFFFFF00000000ULL) | me;
    for (;;) {
        uint64_t cur = *reinterpret_cast<volatile uint64_t *>(&tbl[h].key);
        if (cur == ~0ULL) {
            cur = (uint64_t)atomicCAS(reinterpret_cast<unsigned long long *>(&tbl[h].key), ~0ULL, (unsigned long long)mine);
            if (cur == ~0ULL) return h;
        }
        if ((cur >> 32) == (hv >> 32) && wide_same(c, (uint32_t)cur, me)) return h;
        h = (h + 1) & mask;
    }
}
__device__ __forceinline__ uint32_t gw_find(const GSlot *tbl, uint32_t mask, const WideCtx &c, uint64_t hv, uint32_t me) {
    uint32_t h = (uint32_t)hv & mask;
    for (;;) {
        const uint64_t cur = tbl[h].key;
        if (cur == ~0ULL) return 0xFFFFFFFFu;
        if ((cur >> 32) == (hv >> 32) && wide_same(c, (uint32_t)cur, me)) return h;
        h = (h + 1) & mask;
    }
}

// phase: 0 = insert seq1; 1 = seq2 lookup/claim (memo into aof); 2 = finalize aof (+flags)
template <int PHASE, bool FULL, bool WIDE>
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
    const WideCtx wc{packed + L.pk0, packed + L.pk1, k};
    if (PHASE == 0) {
        if (i >= N0) return;
        const uint32_t h = WIDE ? gw_insert(tbl, L.mask, wc, wide_hash(wc.pk0, i, k), (uint32_t)i)
                                : g_insert(tbl, L.mask, kmer_at<uint64_t>(packed + L.pk0, i, kmask));
        occ_mark(&tbl[h].v0, (uint32_t)i);
    } else if (PHASE == 1) {
        if (i >= N1) return;
        const uint64_t key = WIDE ? wide_hash(wc.pk1, i, k) : kmer_at<uint64_t>(packed + L.pk1, i, kmask);
        uint32_t h;
        if (FULL) {
            h = WIDE ? gw_insert(tbl, L.mask, wc, key, 0x80000000u | (uint32_t)i) : g_insert(tbl, L.mask, key);
            occ_mark(&tbl[h].v1, (uint32_t)i);
        } else {
            h = WIDE ? gw_find(tbl, L.mask, wc, key, 0x80000000u | (uint32_t)i) : g_find(tbl, L.mask, key);
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
            const uint32_t h = WIDE ? gw_find(tbl, L.mask, wc, wide_hash(wc.pk0, i, k), (uint32_t)i)
                                    : g_find(tbl, L.mask, kmer_at<uint64_t>(packed + L.pk0, i, kmask));
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
#define DBGA_HL(PH, F, W) hash_large_kernel<PH, F, W><<<gk, 256, 0, st>>>(pairs, ld, list, k, packed, tbl, status, aof, nd0, nd1)
    const bool wide = k >= 32;
    if (full && wide) { DBGA_HL(0, true, true); DBGA_HL(1, true, true); DBGA_HL(2, true, true); }
    else if (full) { DBGA_HL(0, true, false); DBGA_HL(1, true, false); DBGA_HL(2, true, false); }
    else if (wide) { DBGA_HL(0, false, true); DBGA_HL(1, false, true); DBGA_HL(2, false, true); }
    else { DBGA_HL(0, false, false); DBGA_HL(1, false, false); DBGA_HL(2, false, false); }
#undef DBGA_HL
    *launches += 4;
    return cudaGetLastError();
}

}  // namespace dbga
