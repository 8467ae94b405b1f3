// Shared-memory k-mer table pieces used by the fused stage 1+2 kernel (stage12_small.cu) and the
// k-mer statistics kernel (kmer_stats.cu): 2-bit packing of ASCII bases with validation, k-mer
// extraction from packed words, and the bucketized open-addressing table (one 128-bit LDS fetches
// a whole bucket of 4 x 32-bit or 2 x 64-bit keys; a slot is claimed with one atomicCAS).
// Reference: get_kmers / duplicate_kmers, src/dbga/utils.py:108-154.
#pragma once
#include "kernels.cuh"

namespace dbga {

static constexpr uint16_t NONE16 = 0xFFFFu;
static constexpr int RANK_SORT_MAX = 1024;

__device__ __forceinline__ uint4 lds_v4(const void *p) {
    uint4 v;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// ---- 4 ASCII bases (one 32-bit word) -> 8 bits of 2-bit codes; sets bad on non-ACGT
__device__ __forceinline__ uint32_t pack4(uint32_t w, bool &bad) {
    const uint32_t codes = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t t = codes | (codes >> 4);
    const uint32_t sel = __byte_perm(t, 0u, 0x4420u);            // nibble i = code of byte i
    const uint32_t expect = __byte_perm(0x54474341u, 0u, sel);  // "ACGT"[code]
    bad |= (expect != w);
    return (codes * 0x01041040u) >> 24;                          // c0 | c1<<2 | c2<<4 | c3<<6
}

// 16 bases starting at src (any alignment) -> one packed word.  Two aligned 128-bit loads
// and a byte-granular funnel shift replace 16 byte loads; `off` = src & 15 is the same for
// every word of a sequence, so the shift selection is warp-uniform.  The second load is
// only issued when it contains a byte of the sequence (never reads past the allocation's
// last 16-byte block).  Bytes beyond `valid` are replaced by 'A' (code 0, not an error).
__device__ __forceinline__ uint32_t pack16_fast(const uint8_t *src, int64_t valid, bool &bad) {
    const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
    const uint4 *p = reinterpret_cast<const uint4 *>(src - off);
    const int nv = valid < 16 ? (int)valid : 16;
    const uint4 a = __ldg(p);
    uint4 b = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
    if (off + nv > 16) b = __ldg(p + 1);
    uint32_t w0, w1, w2, w3;
    const uint32_t sh = (off & 3) * 8;
    switch (off >> 2) {
    case 0: w0 = __funnelshift_r(a.x, a.y, sh); w1 = __funnelshift_r(a.y, a.z, sh); w2 = __funnelshift_r(a.z, a.w, sh); w3 = __funnelshift_r(a.w, b.x, sh); break;
    case 1: w0 = __funnelshift_r(a.y, a.z, sh); w1 = __funnelshift_r(a.z, a.w, sh); w2 = __funnelshift_r(a.w, b.x, sh); w3 = __funnelshift_r(b.x, b.y, sh); break;
    case 2: w0 = __funnelshift_r(a.z, a.w, sh); w1 = __funnelshift_r(a.w, b.x, sh); w2 = __funnelshift_r(b.x, b.y, sh); w3 = __funnelshift_r(b.y, b.z, sh); break;
    default: w0 = __funnelshift_r(a.w, b.x, sh); w1 = __funnelshift_r(b.x, b.y, sh); w2 = __funnelshift_r(b.y, b.z, sh); w3 = __funnelshift_r(b.z, b.w, sh); break;
    }
    if (nv < 16) {                 // tail word: neutralise bytes past the end of the sequence
        uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = nv - 4 * j;                       // valid bytes in word j
            const uint32_t m = c >= 4 ? 0xFFFFFFFFu : (c <= 0 ? 0u : ((1u << (8 * c)) - 1u));
            w[j] = (w[j] & m) | (0x41414141u & ~m);
        }
        w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
    }
    return pack4(w0, bad) | (pack4(w1, bad) << 8) | (pack4(w2, bad) << 16) | (pack4(w3, bad) << 24);
}

template <typename KeyT> struct Bucket;
template <> struct Bucket<uint32_t> {
    static constexpr int W = 4;
    static constexpr uint32_t EMPTY = 0xFFFFFFFFu;
    static __device__ __forceinline__ uint32_t mix(uint32_t k) { return k * 0x9E3779B1u; }
    static __device__ __forceinline__ uint32_t kmer(const uint32_t *pk, int p, uint32_t kmask) {
        const int w = p >> 4, sh = (p & 15) * 2;
        return __funnelshift_r(pk[w], pk[w + 1], sh) & kmask;
    }
    // the k-mers at p and p + 1, p even: both come out of the same two packed words
    static __device__ __forceinline__ void kmer2(const uint32_t *pk, int p, uint32_t kmask, uint32_t &k0, uint32_t &k1) {
        const int w = p >> 4, sh = (p & 15) * 2;
        const uint32_t a = pk[w], b = pk[w + 1];
        k0 = __funnelshift_r(a, b, sh) & kmask;
        k1 = __funnelshift_r(a, b, sh + 2) & kmask;
    }
    // position of key / of the first EMPTY in the bucket (-1: none)
    static __device__ __forceinline__ void scan(const uint32_t *base, uint32_t key, int &hit, int &emp) {
        const uint4 kk = lds_v4(base);
        hit = kk.x == key ? 0 : (kk.y == key ? 1 : (kk.z == key ? 2 : (kk.w == key ? 3 : -1)));
        emp = kk.x == EMPTY ? 0 : (kk.y == EMPTY ? 1 : (kk.z == EMPTY ? 2 : (kk.w == EMPTY ? 3 : -1)));
    }
    // lookup only: slots fill in order, so the bucket has a free slot iff its last one is free
    static __device__ __forceinline__ void scan_find(const uint32_t *base, uint32_t key, int &hit, bool &open) {
        const uint4 kk = lds_v4(base);
        hit = kk.x == key ? 0 : (kk.y == key ? 1 : (kk.z == key ? 2 : (kk.w == key ? 3 : -1)));
        open = kk.w == EMPTY;
    }
    static __device__ __forceinline__ uint32_t cas(uint32_t *a, uint32_t v) { return atomicCAS(a, EMPTY, v); }
};
template <> struct Bucket<uint64_t> {
    static constexpr int W = 2;
    static constexpr uint64_t EMPTY = ~0ULL;
    static __device__ __forceinline__ uint32_t mix(uint64_t k) { return (uint32_t)((k * 0x9E3779B97F4A7C15ULL) >> 32); }
    static __device__ __forceinline__ uint64_t kmer(const uint32_t *pk, int p, uint64_t kmask) {
        const int w = p >> 4, sh = (p & 15) * 2;
        const uint32_t a = pk[w], b = pk[w + 1], c = pk[w + 2];
        return (((uint64_t)__funnelshift_r(b, c, sh) << 32) | __funnelshift_r(a, b, sh)) & kmask;
    }
    static __device__ __forceinline__ void kmer2(const uint32_t *pk, int p, uint64_t kmask, uint64_t &k0, uint64_t &k1) {
        const int w = p >> 4, sh = (p & 15) * 2;
        const uint32_t a = pk[w], b = pk[w + 1], c = pk[w + 2];
        k0 = (((uint64_t)__funnelshift_r(b, c, sh) << 32) | __funnelshift_r(a, b, sh)) & kmask;
        k1 = (((uint64_t)__funnelshift_r(b, c, sh + 2) << 32) | __funnelshift_r(a, b, sh + 2)) & kmask;
    }
    static __device__ __forceinline__ void scan(const uint64_t *base, uint64_t key, int &hit, int &emp) {
        const uint4 kk = lds_v4(base);
        const uint64_t k0 = ((uint64_t)kk.y << 32) | kk.x, k1 = ((uint64_t)kk.w << 32) | kk.z;
        hit = k0 == key ? 0 : (k1 == key ? 1 : -1);
        emp = k0 == EMPTY ? 0 : (k1 == EMPTY ? 1 : -1);
    }
    static __device__ __forceinline__ void scan_find(const uint64_t *base, uint64_t key, int &hit, bool &open) {
        const uint4 kk = lds_v4(base);
        const uint64_t k0 = ((uint64_t)kk.y << 32) | kk.x, k1 = ((uint64_t)kk.w << 32) | kk.z;
        hit = k0 == key ? 0 : (k1 == key ? 1 : -1);
        open = k1 == EMPTY;
    }
    static __device__ __forceinline__ uint64_t cas(uint64_t *a, uint64_t v) {
        return (uint64_t)atomicCAS(reinterpret_cast<unsigned long long *>(a), ~0ULL, (unsigned long long)v);
    }
};

// find-or-claim; `claimed` tells whether this call inserted the key.  Slots of a bucket fill in
// order; after a lost race for one slot the later slots that were empty at scan time are
// tried straight from registers (no second bucket load).
template <typename KeyT>
__device__ __forceinline__ uint32_t bkt_insert(KeyT *keys, uint32_t nb, KeyT key, uint32_t hv, bool &claimed) {
    using B = Bucket<KeyT>;
    uint32_t b = __umulhi(hv, nb);
    for (;;) {
        KeyT *base = keys + b * B::W;
        int hit, emp;
        B::scan(base, key, hit, emp);
        if (hit >= 0) { claimed = false; return b * B::W + hit; }
        if (emp >= 0) {
            // slots emp .. W-1 were empty when the bucket was read (slots fill in order)
            for (int e = emp; e < B::W; ++e) {
                const KeyT old = B::cas(base + e, key);
                if (old == B::EMPTY) { claimed = true; return b * B::W + e; }
                if (old == key) { claimed = false; return b * B::W + e; }
            }
        }
        b = b + 1 == nb ? 0 : b + 1;
    }
}
template <typename KeyT>
__device__ __forceinline__ uint32_t bkt_find(const KeyT *keys, uint32_t nb, KeyT key, uint32_t hv) {
    using B = Bucket<KeyT>;
    uint32_t b = __umulhi(hv, nb);
    for (;;) {
        int hit;
        bool open;
        B::scan_find(keys + b * B::W, key, hit, open);
        if (hit >= 0) return b * B::W + hit;
        if (open) return 0xFFFFFFFFu;
        b = b + 1 == nb ? 0 : b + 1;
    }
}

// ---- medium pairs: 4-byte slots that name their k-mer BY REFERENCE (half the bytes of a 32-bit key + word,
// a third of a 64-bit one: a 30 kb pair's table fits shared memory in one piece).  A slot holds the position
// (+ 1; 0 = empty) of the occurrence that claimed it, 11 bits of the hash (in a 12-bit tag that is never 0) and the occurrence flags; a probe
// whose tag matches compares the k-mer at that position of the packed sequence with its own.  The position
// never changes once claimed and the flags are only ever OR-ed in, so claimers and finders commute.
constexpr uint32_t MR_SRC1 = 1u << 28;    // the reference points into seq2 (graph mode: k-mers private to seq2)
constexpr uint32_t MR_SEEN1 = 1u << 29;   // occurs in seq2
constexpr uint32_t MR_SEEN2 = 1u << 30;   // ... more than once
constexpr uint32_t MR_DUP0 = 1u << 31;    // repeated in seq1
__device__ __forceinline__ uint32_t mr_tag(uint32_t hv) { return (((hv >> 8) & 0xFFFu) << 16) | 0x10000u; }   // never 0: an empty slot matches no tag
constexpr uint32_t MR_TAGM = 0x0FFF0000u;
template <typename KeyT>
__device__ __forceinline__ bool mr_same(uint32_t w, KeyT key, const uint32_t *pk0, const uint32_t *pk1, KeyT kmask) {
    return Bucket<KeyT>::kmer((w & MR_SRC1) ? pk1 : pk0, (int)(w & 0xFFFFu) - 1, kmask) == key;
}
// The slot of `key` among the filled slots of a bucket as read (kk), or -1.  Branch-free up to the one k-mer
// comparison almost every probe needs: the first slot whose tag matches is picked with selects; further tag
// matches (a 2^-11 event per slot) are looked at only when that comparison fails.
template <typename KeyT>
__device__ __forceinline__ int mr_match(const uint4 kk, uint32_t tag, KeyT key, const uint32_t *pk0, const uint32_t *pk1, KeyT kmask) {
    const bool t0 = (kk.x & MR_TAGM) == tag, t1 = (kk.y & MR_TAGM) == tag, t2 = (kk.z & MR_TAGM) == tag, t3 = (kk.w & MR_TAGM) == tag;
    if (!(t0 || t1 || t2 || t3)) return -1;
    const int ci = t0 ? 0 : (t1 ? 1 : (t2 ? 2 : 3));
    const uint32_t c = t0 ? kk.x : (t1 ? kk.y : (t2 ? kk.z : kk.w));
    if (mr_same<KeyT>(c, key, pk0, pk1, kmask)) return ci;
    if (ci < 1 && t1 && mr_same<KeyT>(kk.y, key, pk0, pk1, kmask)) return 1;
    if (ci < 2 && t2 && mr_same<KeyT>(kk.z, key, pk0, pk1, kmask)) return 2;
    if (ci < 3 && t3 && mr_same<KeyT>(kk.w, key, pk0, pk1, kmask)) return 3;
    return -1;
}
// find-or-claim (neww: the word a claimer leaves); slots of a bucket fill in order
// Probe sequence: double hashing over buckets -- the step (1, 3, 5 or 7 by two hash bits; `steps` holds the four
// choices, a step that divides the bucket count replaced by 1) keeps overflowing keys from piling up behind full
// buckets, and in a warp the longest probe sequence of 32 lanes is what every lane waits for.
__host__ __device__ inline uint32_t mr_steps(uint32_t nb) {
    return 1u | ((nb % 3u ? 3u : 1u) << 8) | ((nb % 5u ? 5u : 1u) << 16) | ((nb % 7u ? 7u : 1u) << 24);
}
__device__ __forceinline__ uint32_t mr_next(uint32_t b, uint32_t hv, uint32_t nb, uint32_t steps) {
    b += (steps >> (8u * ((hv >> 4) & 3u))) & 0xFFu;
    return b >= nb ? b - nb : b;
}
template <typename KeyT>
__device__ __forceinline__ uint32_t mr_insert(uint32_t *tab, uint32_t nb, uint32_t steps, KeyT key, uint32_t hv, uint32_t neww,
                                              const uint32_t *pk0, const uint32_t *pk1, KeyT kmask, bool &claimed) {
    const uint32_t tag = mr_tag(hv);
    uint32_t b = __umulhi(hv, nb);
    for (;;) {
        uint32_t *base = tab + 4 * b;
        const uint4 kk = lds_v4(base);
        const int hit = mr_match<KeyT>(kk, tag, key, pk0, pk1, kmask);
        if (hit >= 0) { claimed = false; return 4 * b + hit; }
        if (kk.w == 0u) {
            // slots emp .. 3 were empty when the bucket was read; a lost race shows who took the slot
            const int emp = kk.x == 0u ? 0 : (kk.y == 0u ? 1 : (kk.z == 0u ? 2 : 3));
            for (int e = emp; e < 4; ++e) {
                const uint32_t cur = atomicCAS(base + e, 0u, neww);
                if (cur == 0u) { claimed = true; return 4 * b + e; }
                if ((cur & MR_TAGM) == tag && mr_same<KeyT>(cur, key, pk0, pk1, kmask)) { claimed = false; return 4 * b + e; }
            }
        }
        b = mr_next(b, hv, nb, steps);
    }
}
template <typename KeyT>
__device__ __forceinline__ uint32_t mr_find(const uint32_t *tab, uint32_t nb, uint32_t steps, KeyT key, uint32_t hv, const uint32_t *pk0,
                                            const uint32_t *pk1, KeyT kmask) {
    const uint32_t tag = mr_tag(hv);
    uint32_t b = __umulhi(hv, nb);
    for (;;) {
        const uint4 kk = lds_v4(tab + 4 * b);
        const int hit = mr_match<KeyT>(kk, tag, key, pk0, pk1, kmask);
        if (hit >= 0) return 4 * b + hit;
        if (kk.w == 0u) return 0xFFFFFFFFu;          // the bucket never filled up: the key is not in the table
        b = mr_next(b, hv, nb, steps);
    }
}

}  // namespace dbga
