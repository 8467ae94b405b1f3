// Stages 1 + 2 fused for pairs whose k-mer table fits in shared memory: ONE CTA per pair,
// nothing but the sequences is read from HBM and nothing but runs / slots is written.
//
// Stage 1 (reference get_kmers / duplicate_kmers, src/dbga/utils.py:108-154, and the dict walk
// of debruijn_pairwise.py:78-204):
//   * 128-bit loads, 4 bases at a time turned into 2-bit codes with byte-SIMD integer ops
//     (validation included), packed 16 bases per word in shared memory;
//   * open-addressing table in shared memory, bucketized: one 128-bit LDS fetches a whole
//     bucket (4 x 32-bit or 2 x 64-bit keys), so almost every probe finishes in one
//     iteration and the warp does not diverge; a slot is claimed with one atomicCAS; the
//     occurrence bookkeeping is one 32-bit word per slot:
//         A   seq1: claim/find; the claimer stores pos0, later finders store cnt = DUP0
//         B1  seq2: find (graph mode: claim); cnt += 1 (shared-memory add, result unused)
//         B2  q is an anchor iff cnt == 1 (not repeated in seq1, exactly once in seq2) and
//             pos0 valid  ->  a = pos0[slot]
// Stage 2 (merge_node_idx order, extract_bubble :277-300, debruijn_merge_correctness
// utils.py:266-291): run starts/ends are rare (a few per hundred positions), so they are
// appended unordered with shared-memory atomics and then rank-sorted by seq-2 position
// (bitonic sort beyond 1024 runs); adjacent-difference cycle check; slots (slots.cuh).
#include "slots.cuh"
#include "kmer_table.cuh"

namespace dbga {

constexpr uint32_t DUP0 = 0x10000u;      // cnt[]: the k-mer occurs more than once in seq1 (seq-2 counts stay below 2^16)
struct S12Shared {
    unsigned int nSE;      // run starts appended (low 16 bits) | run ends appended (high 16 bits); both < 65 000
    int nA, big;
    int next[4];           // medium pairs: next block of positions of phase A / B1 / B2 (warps take 64 at a time)
    long long bcast[2];
    long long red[32];
};

// TPB = 1024 ("medium" pairs, one CTA per SM): 4-byte slots that name their k-mer by reference
// (kmer_table.cuh: mr_insert / mr_find), so that a 30 kb pair's table fits shared memory whole.  Longer
// pairs split the key space by hash into `parts` partitions that take turns in the table -- every round
// re-scans the packed sequences but inserts / looks up only its own partition's k-mers.  Medium pairs
// stop after stage 1: aof[] goes to HBM and the generic stage-2 kernel takes over (their run count is
// not bounded by the table size).
template <typename KeyT, bool FULL, int TPB>
__global__ void __launch_bounds__(TPB, TPB == 256 ? 5 : 1)      // five CTAs per SM for the small-pair case
stage12_small_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs,
                     const int32_t *__restrict__ pair_list, int count, int k, KeyT kmask, uint32_t cap, uint32_t nb, int parts,
                     int words, int maxq, int want_segments, PairOut *__restrict__ pout, Pools pools,
                     Counters *__restrict__ ctr, SlotCfg cfg, uint8_t *__restrict__ nd0, uint8_t *__restrict__ nd1,
                     int32_t *__restrict__ status, int32_t *__restrict__ aof, uint16_t *__restrict__ memo_g) {
    extern __shared__ __align__(16) uint8_t smem[];
    using B = Bucket<KeyT>;
    constexpr bool MEDIUM = TPB > 256;          // the small-pair instantiation compiles without any partition logic
    if (!MEDIUM) parts = 1;
    uint32_t *pk0 = reinterpret_cast<uint32_t *>(smem);
    uint32_t *pk1 = pk0 + words;
    KeyT *keys = reinterpret_cast<KeyT *>(pk1 + words);
    // small pairs: a 16-bit seq-1 position and one 32-bit word per slot (occurrences in seq2 | DUP0 if
    // repeated in seq1).  Medium pairs: the 4-byte by-reference slots and nothing else (cnt[] is the table).
    uint16_t *pos0 = reinterpret_cast<uint16_t *>(keys + cap);
    uint32_t *cnt = MEDIUM ? reinterpret_cast<uint32_t *>(pk1 + words) : reinterpret_cast<uint32_t *>(pos0 + cap);
    uint16_t *memo = reinterpret_cast<uint16_t *>(cnt + cap);       // later: aof in shared memory
    // aliases valid once the table is dead (after phase B2)
    uint32_t *ska = reinterpret_cast<uint32_t *>(keys);             // unordered run starts (b<<16 | a), cap entries
    uint16_t *ske = pos0;                                           // unordered run ends (b), cap entries
    uint32_t *srt_s = cnt;                                          // rank-sorted starts
    uint16_t *srt_e = memo;                                         // rank-sorted ends
    __shared__ S12Shared sh;

    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t steps = mr_steps(nb);        // medium pairs: probe steps (kmer_table.cuh)
    // nb = cap / B::W buckets (cap is a multiple of 16, not necessarily a power of two) and
    // kmask = 4^k - 1 come in as parameters: constant-bank operands instead of registers
    unsigned long long my_cells = 0, my_nseg = 0;

    for (int it = blockIdx.x; it < count; it += gridDim.x) {
        const int pair = pair_list[it];
        const PairDesc pd = pairs[pair];
        PairOut po;
        po.status = DBGA_OK; po.n_runs = 0; po.run_base = 0; po.slot_base = 0; po.n_slots = 0; po.pad = 0;
        po.n_anchors = 0; po.aln_len = 0;
        if (k > pd.n0 || k > pd.n1) {                                  // utils.py:129-130
            if (tid == 0) { po.status = DBGA_K_INVALID; pout[pair] = po; if (MEDIUM) status[pair] = DBGA_K_INVALID; }
            continue;
        }
        __syncthreads();            // previous pair's shared state is no longer in use
        if (tid == 0) { sh.nSE = 0u; sh.nA = 0; sh.big = 0; sh.next[0] = sh.next[1] = sh.next[2] = 0; }
        // ---- pack + table init
        bool bad = false;
        {
            const uint8_t *s0 = seqs + pd.off0, *s1 = seqs + pd.off1;
            for (int w = tid; w < words; w += T) {
                const int64_t r0 = (int64_t)pd.n0 - 16 * w, r1 = (int64_t)pd.n1 - 16 * w;
                pk0[w] = r0 > 0 ? pack16_fast(s0 + 16 * w, r0, bad) : 0u;
                pk1[w] = r1 > 0 ? pack16_fast(s1 + 16 * w, r1, bad) : 0u;
            }
        }
        auto init_table = [&]() {       // keys + pos0 = all ones, cnt = 0 (medium pairs: the table is cnt[], empty = 0)
            uint4 *t4 = reinterpret_cast<uint4 *>(keys);
            // (the 16-bit positions need no initial value unless seq-2-only k-mers can own a slot: graph mode)
            const int n_ff = MEDIUM ? 0 : (int)((cap * (sizeof(KeyT) + (!FULL ? 0 : 2))) / 16), n_00 = (int)(cap * 4 / 16);
            for (int i = tid; i < n_ff; i += T) t4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
            uint4 *z4 = reinterpret_cast<uint4 *>(cnt);
            for (int i = tid; i < n_00; i += T) z4[i] = make_uint4(0u, 0u, 0u, 0u);
        };
        init_table();                   // first partition's table, in the same phase as the packing
        if (__syncthreads_or(bad)) {
            if (tid == 0) { po.status = DBGA_BAD_CHAR; pout[pair] = po; if (MEDIUM) status[pair] = DBGA_BAD_CHAR; }
            continue;
        }
        const int N0 = pd.n0 - k + 1, N1 = pd.n1 - k + 1;
        if (MEDIUM) {
            // medium pairs keep memo[] (the slot each seq-2 position found) in a global scratch (L2): the whole
            // shared memory goes to the table.  Even start, two slack entries per pair: 32-bit accesses, disjoint regions.
            memo = memo_g + ((pd.q_base + 2 * (int64_t)pair + 1) & ~(int64_t)1);
        }
        for (int g = 0; g < parts; ++g) {
            if (g > 0) { __syncthreads(); init_table(); if (tid == 0) { sh.next[0] = sh.next[1] = sh.next[2] = 0; } __syncthreads(); }
            // partition of a k-mer: low 16 bits of its hash (the bucket uses the high bits)
            auto in_part = [&](uint32_t hv) { return !MEDIUM || parts == 1 || (int)(((hv & 0xFFFFu) * (uint32_t)parts) >> 16) == g; };
            // Every thread takes two adjacent positions per round: one pair of packed words
            // yields both k-mers and the two probes are independent (latency overlap).
            // ---- A: seq1 k-mers
            auto phase_a = [&](KeyT key, int p) {
                const uint32_t hv = B::mix(key);
                if (!in_part(hv)) return;
                bool claimed;
                if (MEDIUM) {
                    const uint32_t h = mr_insert<KeyT>(cnt, nb, steps, key, hv, ((uint32_t)p + 1u) | mr_tag(hv), pk0, pk1, kmask, claimed);
                    if (!claimed) atomicOr(&cnt[h], MR_DUP0);
                    return;
                }
                const uint32_t h = bkt_insert<KeyT>(keys, nb, key, hv, claimed);
                if (claimed) pos0[h] = (uint16_t)p;
                else cnt[h] = DUP0;
            };
            // Medium pairs: a warp takes the next 64 positions whenever it is done with its last ones (one shared-memory
            // add per warp and block): the number of buckets a probe needs is long-tailed, and with a fixed share per
            // warp the whole CTA waited at the phase's barrier for the unluckiest one.
            auto next_block = [&](int which) -> int {
                int base = 0;
                if (lane == 0) base = atomicAdd(&sh.next[which], 64);
                return __shfl_sync(0xffffffffu, base, 0);
            };
            if (MEDIUM) {
                for (int base = next_block(0); base < N0; base = next_block(0)) {
                    const int p = base + 2 * lane;
                    if (p < N0) {
                        KeyT k0, k1;
                        B::kmer2(pk0, p, kmask, k0, k1);
                        phase_a(k0, p);
                        if (p + 1 < N0) phase_a(k1, p + 1);
                    }
                }
            } else {
                for (int p = 2 * tid; p < N0; p += 2 * T) {
                    KeyT k0, k1;
                    B::kmer2(pk0, p, kmask, k0, k1);
                    phase_a(k0, p);
                    if (p + 1 < N0) phase_a(k1, p + 1);
                }
            }
            __syncthreads();
            // ---- B1: seq2 k-mers; memo[q] = slot (or, medium pairs, what an earlier partition left)
            auto phase_b1 = [&](KeyT key, uint32_t keep, int q) -> uint32_t {
                const uint32_t hv = B::mix(key);
                if (!in_part(hv)) return keep;
                uint32_t h;
                if (MEDIUM) {
                    bool claimed = false;
                    if (FULL) h = mr_insert<KeyT>(cnt, nb, steps, key, hv, ((uint32_t)q + 1u) | mr_tag(hv) | MR_SRC1 | MR_SEEN1, pk0, pk1, kmask, claimed);
                    else h = mr_find<KeyT>(cnt, nb, steps, key, hv, pk0, pk1, kmask);
                    if (h != 0xFFFFFFFFu && !claimed) {
                        const uint32_t old = atomicOr(&cnt[h], MR_SEEN1);
                        if ((old & (MR_SEEN1 | MR_SEEN2)) == MR_SEEN1) atomicOr(&cnt[h], MR_SEEN2);
                    }
                    return h != 0xFFFFFFFFu ? h : (uint32_t)NONE16;
                }
                if (FULL) {
                    bool claimed;
                    h = bkt_insert<KeyT>(keys, nb, key, hv, claimed);
                } else {
                    h = bkt_find<KeyT>(keys, nb, key, hv);
                }
                if (h != 0xFFFFFFFFu) atomicAdd(&cnt[h], 1u);      // result unused: a fire-and-forget shared-memory add
                return h != 0xFFFFFFFFu ? h : (uint32_t)NONE16;
            };
            uint32_t *memo2 = reinterpret_cast<uint32_t *>(memo);   // two positions per word
            auto b1_pair = [&](int q) {
                KeyT k0, k1;
                B::kmer2(pk1, q, kmask, k0, k1);
                // (medium pairs: a position outside this partition gets NONE16 here and is skipped by this round's B2)
                const uint32_t m0 = phase_b1(k0, NONE16, q);
                const uint32_t m1 = q + 1 < N1 ? phase_b1(k1, NONE16, q + 1) : (uint32_t)NONE16;
                memo2[q >> 1] = m0 | (m1 << 16);
            };
            if (MEDIUM) {
                for (int base = next_block(1); base < N1; base = next_block(1)) {
                    const int q = base + 2 * lane;
                    if (q < N1) b1_pair(q);
                }
            } else {
                for (int q = 2 * tid; q < N1; q += 2 * T) b1_pair(q);
            }
            __syncthreads();
            // ---- B2: anchors; memo[q] becomes aof[q] (position in seq1 or NONE16)
            auto phase_b2 = [&](uint32_t h, int q) -> uint32_t {
                uint32_t a = NONE16;
                if (h != NONE16) {
                    const uint32_t w = cnt[h];
                    const bool d = MEDIUM ? (w & (MR_DUP0 | MR_SEEN2)) != 0u : w != 1u;     // repeated in seq1 (DUP0) or seen more than once in seq2
                    if (!d) a = MEDIUM ? ((w & MR_SRC1) ? (uint32_t)NONE16 : (w & 0xFFFFu) - 1u) : pos0[h];   // NONE16 when the k-mer is private to seq2 (graph mode)
                    if (FULL) nd1[pd.q_base + q] = !d;
                }
                return a;
            };
            if (MEDIUM) {
                // this partition's positions get their final aof[] entry (position in seq1 or -1) straight in HBM
                int32_t *ga = aof + pd.q_base;
                for (int q = 2 * tid; q < N1; q += 2 * T) {
                    bool p0 = true, p1 = q + 1 < N1;
                    if (parts > 1) {
                        KeyT k0, k1;
                        B::kmer2(pk1, q, kmask, k0, k1);
                        p0 = in_part(B::mix(k0)); p1 = p1 && in_part(B::mix(k1));
                    }
                    const uint32_t hh = memo2[q >> 1];
                    if (p0) { const uint32_t a = phase_b2(hh & 0xFFFFu, q); ga[q] = a == NONE16 ? -1 : (int32_t)a; }
                    if (p1) { const uint32_t a = phase_b2(hh >> 16, q + 1); ga[q + 1] = a == NONE16 ? -1 : (int32_t)a; }
                }
            } else {
                for (int q = 2 * tid; q < N1; q += 2 * T) {
                    const uint32_t hh = memo2[q >> 1];
                    const uint32_t a0 = phase_b2(hh & 0xFFFFu, q);
                    const uint32_t a1 = q + 1 < N1 ? phase_b2(hh >> 16, q + 1) : (uint32_t)NONE16;
                    memo2[q >> 1] = a0 | (a1 << 16);
                }
            }
            if (FULL) {
                for (int p = tid; p < N0; p += T) {
                    const KeyT key = B::kmer(pk0, p, kmask);
                    const uint32_t hv = B::mix(key);
                    if (!in_part(hv)) continue;
                    if (MEDIUM) {
                        const uint32_t h = mr_find<KeyT>(cnt, nb, steps, key, hv, pk0, pk1, kmask);
                        nd0[pd.p_base + p] = (cnt[h] & (MR_DUP0 | MR_SEEN2)) == 0u;
                        continue;
                    }
                    const uint32_t h = bkt_find<KeyT>(keys, nb, key, hv);
                    nd0[pd.p_base + p] = cnt[h] <= 1u;      // not repeated in seq1, at most once in seq2
                }
            }
        }
        if (MEDIUM) continue;                         // medium pair: aof[] is complete, the generic stage 2 takes over
        __syncthreads();                              // table dead from here on
        // ---- stage 2: run starts / ends (rare) appended unordered.  Every lane flags what its two
        //      positions start / end; the warp's counts are scanned with shuffles and ONE shared-memory
        //      add per warp step reserves room for all of them (starts in the low half of the counter
        //      word, ends in the high half), instead of an atomic per start / end.
        int cA = 0;
        {
            const uint32_t *memo2 = reinterpret_cast<const uint32_t *>(memo);
            auto is_next = [](uint32_t x, uint32_t y) { return x != NONE16 && y != NONE16 && y == x + 1u; };   // y continues x's run
            for (int q0 = 0; q0 < N1; q0 += 2 * T) {
                const int q = q0 + 2 * tid;
                uint32_t a0 = NONE16, a1 = NONE16, pv = NONE16, nx = NONE16;
                if (q < N1) {
                    const uint32_t w = memo2[q >> 1];             // the odd half past N1 was left NONE16 by B2
                    a0 = w & 0xFFFFu; a1 = w >> 16;
                    if (w != 0xFFFFFFFFu) {
                        pv = q > 0 ? memo[q - 1] : NONE16;
                        nx = q + 2 < N1 ? memo[q + 2] : NONE16;
                    }
                }
                const bool h0 = a0 != NONE16, h1 = a1 != NONE16, link = is_next(a0, a1);
                const bool s0 = h0 && !is_next(pv, a0), e0 = h0 && !link, s1 = h1 && !link, e1 = h1 && !is_next(a1, nx);
                cA += (int)h0 + (int)h1;
                const uint32_t mine = (uint32_t)s0 + (uint32_t)s1 + (((uint32_t)e0 + (uint32_t)e1) << 16);
                if (__ballot_sync(0xffffffffu, mine != 0u) == 0u) continue;      // nothing starts or ends here (warp-uniform)
                uint32_t inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                uint32_t base = 0;
                if (lane == 31) base = atomicAdd(&sh.nSE, inc);
                base = __shfl_sync(0xffffffffu, base, 31) + inc - mine;
                uint32_t ps = base & 0xFFFFu, pe = base >> 16;
                if (s0) ska[ps++] = ((uint32_t)q << 16) | a0;
                if (s1) ska[ps] = ((uint32_t)(q + 1) << 16) | a1;
                if (e0) ske[pe++] = (uint16_t)q;
                if (e1) ske[pe] = (uint16_t)(q + 1);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cA += __shfl_xor_sync(0xffffffffu, cA, o);
        if (lane == 0 && cA) atomicAdd(&sh.nA, cA);
        __syncthreads();
        const int R = (int)(sh.nSE & 0xFFFFu);
        const uint32_t *S = srt_s;
        const uint16_t *E = srt_e;
        if (R <= RANK_SORT_MAX) {
            for (int t = tid; t < R; t += T) {
                const uint32_t ks = ska[t];
                const uint16_t ke = ske[t];
                int rs = 0, re = 0;
                for (int u = 0; u < R; ++u) { rs += ska[u] < ks; re += ske[u] < ke; }
                srt_s[rs] = ks; srt_e[re] = ke;
            }
        } else {                                       // in-place bitonic sort, padded to a power of two
            int n2 = 1;
            while (n2 < R) n2 <<= 1;
            for (int i = R + tid; i < n2; i += T) { ska[i] = 0xFFFFFFFFu; ske[i] = 0xFFFFu; }
            __syncthreads();
            for (int k2 = 2; k2 <= n2; k2 <<= 1) {
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < n2; i += T) {
                        const int x = i ^ j;
                        if (x > i) {
                            const bool asc = (i & k2) == 0;
                            const uint32_t a0 = ska[i], a1 = ska[x];
                            if ((a0 > a1) == asc) { ska[i] = a1; ska[x] = a0; }
                            const uint16_t e0 = ske[i], e1 = ske[x];
                            if ((e0 > e1) == asc) { ske[i] = e1; ske[x] = e0; }
                        }
                    }
                    __syncthreads();
                }
            }
            S = ska; E = ske;
        }
        __syncthreads();
        auto run = [&](int r, int &a, int &b, int &len) {
            const uint32_t key = S[r];
            a = (int)(key & 0xFFFFu); b = (int)(key >> 16); len = (int)E[r] - b + 1;
        };
        // ---- colinearity (utils.py:266-291), then room in the runs and slots pools: two
        //      threads fetch the two counters at once (one global round trip, one barrier)
        int my_cyc = 0;
        for (int r = tid + 1; r < R; r += T) {
            int a0, b0, l0, a1, b1, l1;
            run(r - 1, a0, b0, l0); run(r, a1, b1, l1);
            if (a1 <= a0 + l0 - 1) my_cyc = 1;
        }
        const bool cyc = __syncthreads_or(my_cyc) != 0;
        const bool need_slots = !cyc && want_segments;
        const int n_slots = R == 0 ? 1 : R + 1;
        if (tid == 0) {
            long long base = 0;
            if (R > 0) {
                base = (long long)atomicAdd(&ctr->runs_used, (unsigned long long)R);
                if (base + R > pools.runs_cap) { atomicExch(&ctr->overflow, 1u); base = -1; }
            }
            sh.bcast[0] = base;
        }
        if (tid == 32 && need_slots) {
            long long base = (long long)atomicAdd(&ctr->slots_used, (unsigned long long)n_slots);
            if (base + n_slots > pools.slots_cap) { atomicExch(&ctr->overflow, 1u); base = -1; }
            sh.bcast[1] = base;
        }
        __syncthreads();
        const long long run_base = sh.bcast[0];
        if (run_base < 0) {
            if (tid == 0) { po.status = DBGA_TOO_LARGE; pout[pair] = po; }
            continue;
        }
        {
            int32_t *gr = pools.runs + 3 * run_base;
            for (int r = tid; r < R; r += T) {
                int a, b, len;
                run(r, a, b, len);
                gr[3 * r] = a; gr[3 * r + 1] = b; gr[3 * r + 2] = len;
            }
        }
        po.n_runs = R; po.run_base = run_base; po.n_anchors = sh.nA;
        if (cyc) {
            if (tid == 0) { po.status = DBGA_CYCLE; pout[pair] = po; }
            continue;
        }
        if (!want_segments) {
            if (tid == 0) pout[pair] = po;
            continue;
        }
        const long long slot_base = sh.bcast[1];
        if (slot_base < 0) {
            if (tid == 0) { po.status = DBGA_TOO_LARGE; pout[pair] = po; }
            continue;
        }
        build_slots(pair, pd, R, run, slot_base, pools, ctr, cfg, &sh.big, my_cells, my_nseg);
        __syncthreads();
        po.slot_base = slot_base; po.n_slots = n_slots;
        if (tid == 0) {
            if (sh.big) po.status = DBGA_TOO_LARGE;
            pout[pair] = po;
        }
    }
    // per-CTA totals of DP work
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_cells += __shfl_xor_sync(0xffffffffu, my_cells, o);
        my_nseg += __shfl_xor_sync(0xffffffffu, my_nseg, o);
    }
    if (lane == 0) { sh.red[wid] = (long long)my_cells; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long c = 0;
        for (int w = 0; w < (T >> 5); ++w) c += (unsigned long long)sh.red[w];
        if (c) atomicAdd(&ctr->cells, c);
    }
    __syncthreads();
    if (lane == 0) { sh.red[wid] = (long long)my_nseg; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long c = 0;
        for (int w = 0; w < (T >> 5); ++w) c += (unsigned long long)sh.red[w];
        if (c) atomicAdd(&ctr->nseg, c);
    }
}

// medium: the partitioned-table instantiation (no 16-bit position array, no shared-memory memo[])
size_t stage12_small_smem(int key_bytes, uint32_t cap, int words, int maxq, bool medium) {
    if (medium) return (size_t)words * 8 + (size_t)cap * 4 + 16;       // packed sequences + by-reference slots
    return (size_t)words * 8 + (size_t)cap * (key_bytes + 6) + (size_t)((maxq + 7) / 8 * 8) * 2 + 16;
}

template <typename KeyT, bool FULL, int TPB>
static cudaError_t launch_t(const uint8_t *seqs, const PairDesc *pairs, const int32_t *list, int count, int k,
                            uint32_t cap, int parts, int words, int maxq, bool want_segments, PairOut *pout,
                            Pools pools, Counters *ctr, SlotCfg cfg, uint8_t *nd0, uint8_t *nd1, int32_t *status,
                            int32_t *aof, uint16_t *memo_g, int num_sms, cudaStream_t st) {
    const size_t smem = stage12_small_smem(sizeof(KeyT), cap, words, maxq, TPB > 256);
    auto kern = stage12_small_kernel<KeyT, FULL, TPB>;
    cudaError_t e = allow_max_dyn_smem(kern);
    if (e != cudaSuccess) return e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TPB, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    int grid = num_sms * occ;
    if (grid > count) grid = count;
    const KeyT kmask = (2 * k >= (int)(8 * sizeof(KeyT))) ? ~(KeyT)0 : (((KeyT)1 << (2 * k)) - 1);
    kern<<<grid, TPB, smem, st>>>(seqs, pairs, list, count, k, kmask, cap, cap / (uint32_t)(TPB > 256 ? 4 : Bucket<KeyT>::W), parts, words, maxq,
                                  want_segments ? 1 : 0, pout, pools, ctr, cfg, nd0, nd1, status, aof, memo_g);
    return cudaGetLastError();
}

template <typename KeyT>
static cudaError_t launch_k(bool full, bool medium, int parts, const uint8_t *seqs, const PairDesc *pairs, const int32_t *list,
                            int count, int k, uint32_t cap, int words, int maxq, bool want_segments, PairOut *pout,
                            Pools pools, Counters *ctr, SlotCfg cfg, uint8_t *nd0, uint8_t *nd1, int32_t *status,
                            int32_t *aof, uint16_t *memo_g, int num_sms, cudaStream_t st) {
#define DBGA_S12_ARGS seqs, pairs, list, count, k, cap, parts, words, maxq, want_segments, pout, pools, ctr, cfg, nd0, nd1, status, aof, memo_g, num_sms, st
    if (medium) return full ? launch_t<KeyT, true, 1024>(DBGA_S12_ARGS) : launch_t<KeyT, false, 1024>(DBGA_S12_ARGS);
    return full ? launch_t<KeyT, true, 256>(DBGA_S12_ARGS) : launch_t<KeyT, false, 256>(DBGA_S12_ARGS);
#undef DBGA_S12_ARGS
}

// !medium: stages 1+2 fused (small pairs).  medium: stage 1 only with by-reference slots, `parts` table
// partitions; status[] / aof[] feed the generic stage-2 kernel.
cudaError_t launch_stage12_small(const uint8_t *seqs, const PairDesc *pairs, const int32_t *list, int count, int k,
                                 uint32_t cap, bool medium, int parts, int words, int maxq, bool full, bool want_segments,
                                 PairOut *pout, Pools pools, Counters *ctr, int max_abs_score, int go_raw, int ge_raw,
                                 uint8_t *nd0, uint8_t *nd1, int32_t *status, int32_t *aof, uint16_t *memo_g, int num_sms,
                                 cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    const SlotCfg cfg{max_abs_score, go_raw, ge_raw};
    if (k <= 15)
        return launch_k<uint32_t>(full, medium, parts, seqs, pairs, list, count, k, cap, words, maxq, want_segments, pout, pools, ctr,
                                  cfg, nd0, nd1, status, aof, memo_g, num_sms, st);
    return launch_k<uint64_t>(full, medium, parts, seqs, pairs, list, count, k, cap, words, maxq, want_segments, pout, pools, ctr, cfg,
                              nd0, nd1, status, aof, memo_g, num_sms, st);
}

}  // namespace dbga
