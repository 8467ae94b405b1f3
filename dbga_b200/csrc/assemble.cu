// Result assembly: turns runs of anchors + per-segment alignments into the two gapped rows
// the reference builds by string concatenation in DeBruijnPairwise.alignment()
// (src/dbga/debruijn_pairwise.py:499-543): first segment whole, every merge node's first
// character except the last (:508-509), mid segments with their first column dropped
// (:499-500), tail segment whole (:521-538).  Two launches: layout (item offsets per pair and,
// by a single-pass scan across pairs, where every pair's rows / runs / copy tiles start -- the
// output is compacted so that only the used bytes travel back to the host) and an
// output-centric copy (a thread owns 16 bytes of both rows, 128-bit stores).
#include "kernels.cuh"

namespace dbga {

// ---------------------------------------------------------------- layout + scans (one pass)
// Items of a pair in output order: slot 0, run 0, slot 1, run 1, ..., run R-1, slot R (one slot when
// R == 0).  Item i of a pair lives at item_off[2 * slot_base + i] (slot ranges of different pairs
// are disjoint, and a pair with R runs has R + 1 slots).
//
// asm_layout_kernel: a CTA takes the next block of LAY_PAIRS pairs (ticket order); one warp per
// pair scans the item lengths (32 at a time) into pair-relative item offsets and the pair's
// totals; the block totals of (aligned columns, reported runs, copy tiles) are then scanned across
// CTAs with a decoupled look-back, so aln_off / run_off / tile_off come out of the same launch.
constexpr int ASM_TILE = 2048;        // output columns per copy tile
constexpr int ASM_THREADS = 128;
constexpr int LAY_PAIRS = 64;
constexpr int LAY_THREADS = 256;

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
    return *reinterpret_cast<const volatile unsigned int *>(p);
}

__global__ void __launch_bounds__(LAY_THREADS)
asm_layout_kernel(const PairOut *__restrict__ pout, int64_t num_pairs, const Slot *__restrict__ slots,
                  const int32_t *__restrict__ runs, int32_t *__restrict__ item_off, AsmScanRec *__restrict__ scan,
                  Counters *__restrict__ ctr, AsmOut out, int64_t *__restrict__ tile_off, int flags) {
    __shared__ long long s_v[3][LAY_PAIRS];
    __shared__ long long s_excl[3];
    __shared__ unsigned int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool compact = (flags & DBGA_FLAG_COMPACT_ROWS) != 0, no_runs = (flags & DBGA_FLAG_NO_RUNS) != 0 && !compact;
    if (tid == 0) s_ticket = atomicAdd(&ctr->scan_ticket, 1u);
    __syncthreads();
    const int64_t blk = s_ticket, pair0 = blk * LAY_PAIRS;
    for (int w = wid; w < LAY_PAIRS; w += LAY_THREADS / 32) {
        const int64_t pair = pair0 + w;
        long long carry = 0, sc = 0, ce = 0, nrep = 0;
        int ns = 0, ntile = 0, st = DBGA_OK;
        if (pair < num_pairs) {
            const PairOut po = pout[pair];
            st = po.status;
            bool ok = st == DBGA_OK;
            const bool report = ok || st == DBGA_CYCLE;
            const int R = po.n_runs;
            if (ok && po.n_slots > 0) {
                const int n_items = po.n_slots + R;
                int32_t *io = item_off + 2 * po.slot_base;
                for (int ib = 0; ib < n_items; ib += 32) {
                    const int it = ib + lane;
                    long long len = 0;
                    if (it < n_items) {
                        const int t = it >> 1;
                        if (R == 0 || (it & 1) == 0) {
                            const Slot *sp = slots + po.slot_base + t;
                            const int2 sl = *reinterpret_cast<const int2 *>(&sp->score);     // (score, len)
                            const int2 mn = make_int2(sp->m, sp->n);
                            const bool mid = R > 0 && t > 0 && t < po.n_slots - 1;
                            len = sl.y - ((mid && sl.y > 0) ? 1 : 0);                        // debruijn_pairwise.py:499-500
                            if (mn.x > 0 && mn.y > 0) { sc += sl.x; ce += (long long)mn.x * mn.y; ++ns; }
                        } else if (!compact) {
                            len = runs[3 * (po.run_base + t) + 2] - (t == R - 1 ? 1 : 0);   // last merge node is left to the tail
                        }
                    }
                    long long inc = len;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const long long v = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += v;
                    }
                    if (it < n_items) io[it] = (int32_t)(carry + inc - len);
                    carry += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (carry > 0x7fffffffLL) { ok = false; st = DBGA_TOO_LARGE; carry = 0; }      // 32-bit item offsets
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sc += __shfl_xor_sync(0xffffffffu, sc, o);
                ce += __shfl_xor_sync(0xffffffffu, ce, o);
                ns += __shfl_xor_sync(0xffffffffu, ns, o);
            }
            if (!ok) { carry = 0; sc = 0; ce = 0; ns = 0; }
            nrep = (report && !no_runs && st != DBGA_TOO_LARGE) ? R : 0;     // cycle pairs still report the anchors that cycle
            ntile = (int)((carry + ASM_TILE - 1) / ASM_TILE);
            if (ntile == 0 && (nrep > 0 || (compact && ok && po.n_slots > 0))) ntile = 1;   // runs / segment widths are copied by the pair's first tile
            if (lane == 0) {
                out.score[pair] = sc; out.cells[pair] = ce; out.nseg[pair] = ns;
                out.nanch[pair] = (ok || st == DBGA_CYCLE) ? po.n_anchors : 0;
                out.status[pair] = st;
                if (out.sop) *reinterpret_cast<int4 *>(out.sop + 4 * pair) = make_int4(0, 0, 0, 0);
            }
        }
        if (lane == 0) { s_v[0][w] = carry; s_v[1][w] = nrep; s_v[2][w] = ntile; }
    }
    __syncthreads();
    if (wid == 0) {
        long long a[3][2], inc[3], tot[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            a[q][0] = s_v[q][2 * lane]; a[q][1] = s_v[q][2 * lane + 1];
            inc[q] = a[q][0] + a[q][1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long v = __shfl_up_sync(0xffffffffu, inc[q], o);
                if (lane >= o) inc[q] += v;
            }
            tot[q] = __shfl_sync(0xffffffffu, inc[q], 31);
        }
        if (lane == 0) {
            AsmScanRec *me = scan + blk;
            long long ex[3] = {0, 0, 0};
            if (blk > 0) {
#pragma unroll
                for (int q = 0; q < 3; ++q) me->agg[q] = (unsigned long long)tot[q];
                __threadfence();
                atomicExch(&me->flag, 1u);
                for (int64_t j = blk - 1; j >= 0; --j) {        // decoupled look-back
                    const AsmScanRec *r = scan + j;
                    unsigned int f;
                    do { f = ld_volatile_u32(&r->flag); } while (f == 0u);
                    __threadfence();
                    const volatile unsigned long long *src = f == 2u ? r->pre : r->agg;
#pragma unroll
                    for (int q = 0; q < 3; ++q) ex[q] += (long long)src[q];
                    if (f == 2u) break;
                }
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) me->pre[q] = (unsigned long long)(ex[q] + tot[q]);
            __threadfence();
            atomicExch(&me->flag, 2u);
#pragma unroll
            for (int q = 0; q < 3; ++q) s_excl[q] = ex[q];
        }
        __syncwarp();
        int64_t *dst[3] = {out.aln_off, out.run_off, tile_off};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const long long e0 = s_excl[q] + inc[q] - a[q][0] - a[q][1];
            const int64_t p0 = pair0 + 2 * lane;
            if (p0 < num_pairs) dst[q][p0] = e0;
            if (p0 + 1 < num_pairs) dst[q][p0 + 1] = e0 + a[q][0];
            if (lane == 31 && pair0 + LAY_PAIRS >= num_pairs) {          // last block: grand totals
                const long long g = s_excl[q] + tot[q];
                dst[q][num_pairs] = g;
                if (q == 0) ctr->total_aln = (unsigned long long)g;
                else if (q == 1) ctr->total_runs = (unsigned long long)g;
                else ctr->total_tiles = (unsigned long long)g;
            }
        }
    }
}

// ---------------------------------------------------------------- output-centric copy
// Where the bytes of one item come from: column d of the item is p0[d * st0] / p1[d * st1].
// Inside a run both rows are the same sequence bytes; a DP segment's rows sit reversed in the
// scratch (stage 3 writes them while walking back); a one-sided segment is sequence bytes against
// a constant '-' (stride 0).
__device__ const uint8_t g_dash[16] = {'-', '-', '-', '-', '-', '-', '-', '-', '-', '-', '-', '-', '-', '-', '-', '-'};

struct AsmSrc { const uint8_t *p0, *p1; int st0, st1; };

__device__ __forceinline__ AsmSrc asm_resolve(int i, int d, int outlen, int R, int n_slots, const PairDesc &pd,
                                              const int32_t *__restrict__ gr, const uint8_t *__restrict__ seqs,
                                              const uint8_t *__restrict__ strbuf) {
    AsmSrc s;
    const int t = i >> 1;
    if (R > 0 && (i & 1)) {                         // run t: anchors (a + u, b + u)
        s.p0 = s.p1 = seqs + pd.off0 + __ldg(gr + 3 * t) + d;
        s.st0 = s.st1 = 1;
        return s;
    }
    int x0 = 0, x1 = pd.n0, y0 = 0, y1 = pd.n1;
    if (R > 0) {
        if (t > 0) { const int len = __ldg(gr + 3 * t - 1); x0 = __ldg(gr + 3 * t - 3) + len - 1; y0 = __ldg(gr + 3 * t - 2) + len - 1; }
        if (t < R) { x1 = __ldg(gr + 3 * t); y1 = __ldg(gr + 3 * t + 1); }
    }
    const int m = x1 - x0, n = y1 - y0;
    const int skip = (R > 0 && t > 0 && t < n_slots - 1) ? 1 : 0;          // (an item of this kind with len == 0 is never resolved)
    if (m > 0 && n > 0) {                           // reversed rows: column c of the segment is at [L - 1 - c]
        const int L = outlen + skip;
        s.p0 = strbuf + 2 * pd.off0 + x0 + y0 + (L - 1 - skip - d);
        s.p1 = s.p0 + ((pd.off1 - pd.off0) + pd.n1);
        s.st0 = s.st1 = -1;
    } else if (m > 0) {                             // utils.py:187-188
        s.p0 = seqs + pd.off0 + x0 + skip + d; s.st0 = 1;
        s.p1 = g_dash; s.st1 = 0;
    } else {                                        // utils.py:189-190
        s.p0 = g_dash; s.st0 = 0;
        s.p1 = seqs + pd.off1 + y0 + skip + d; s.st1 = 1;
    }
    return s;
}

// A CTA takes one tile of ASM_TILE output columns of one pair; a thread owns a 16-byte aligned
// chunk of the two output rows, finds the item its first column falls into (binary search over the
// pair's item offsets) and walks the items from there.  The four alignment-quality counts of
// pairwise_alignment_score (reference src/dbga/utils.py:440-504) fall out of the same walk: every
// column is classified (match / mismatch / gap in row 1 / gap in row 2) and a gap column extends
// iff the previous column is of the same kind.
__global__ void __launch_bounds__(ASM_THREADS)
asm_copy_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, int64_t num_pairs,
                const PairOut *__restrict__ pout, const int32_t *__restrict__ runs, const int32_t *__restrict__ item_off,
                const Counters *__restrict__ ctr, const uint8_t *__restrict__ strbuf, AsmOut out,
                const int64_t *__restrict__ tile_off, uint8_t *__restrict__ aln0, uint8_t *__restrict__ aln1,
                int32_t *__restrict__ runs_out, int32_t *__restrict__ seg_cols, int flags) {
    __shared__ long long s_pair;
    __shared__ int s_cnt[4];
    const int tid = threadIdx.x, lane = tid & 31;
    const bool compact = (flags & DBGA_FLAG_COMPACT_ROWS) != 0, no_runs = (flags & DBGA_FLAG_NO_RUNS) != 0 && !compact;
    const long long total_tiles = (long long)ctr->total_tiles;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        __syncthreads();
        if (tid == 0) {                 // the pair that owns this tile: last p with tile_off[p] <= tile
            long long lo = 0, hi = num_pairs;               // invariant: tile_off[lo] <= tile < tile_off[hi]
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (tile_off[mid] <= tile) lo = mid; else hi = mid;
            }
            s_pair = lo;
            s_cnt[0] = s_cnt[1] = s_cnt[2] = s_cnt[3] = 0;
        }
        __syncthreads();
        const int64_t pair = s_pair;
        const PairOut po = pout[pair];
        const PairDesc pd = pairs[pair];
        const int tp = (int)(tile - tile_off[pair]);
        const int64_t a_off = out.aln_off[pair];
        const int aln_len = (int)(out.aln_off[pair + 1] - a_off);
        const int R = po.n_runs;
        const int32_t *gr = runs + 3 * po.run_base;
        const bool ok = po.status == DBGA_OK && po.n_slots > 0;
        const int n_items = ok ? po.n_slots + R : 0;
        const int32_t *io = item_off + 2 * po.slot_base;
        if (tp == 0) {
            const int64_t r_off = out.run_off[pair];
            if (!no_runs && R > 0) for (int r = tid; r < 3 * R; r += ASM_THREADS) runs_out[3 * r_off + r] = __ldg(gr + r);
            if (compact && ok)           // columns of every segment, in output order (what dbga_expand_rows reads)
                for (int t = tid; t < po.n_slots; t += ASM_THREADS) {
                    const int i = R > 0 ? 2 * t : 0;
                    seg_cols[r_off + pair + t] = (i + 1 < n_items ? __ldg(io + i + 1) : aln_len) - __ldg(io + i);
                }
        }
        const int c_lo = tp * ASM_TILE, c_hi = min(c_lo + ASM_TILE, aln_len);
        if (c_hi <= c_lo) continue;
        const int64_t g_lo = a_off + c_lo, g_hi = a_off + c_hi;
        const int64_t base0 = g_lo & ~(int64_t)15;
        int cnt_m = 0, cnt_x = 0, cnt_o = 0, cnt_e = 0;
        for (int64_t cb = base0 + 16 * (int64_t)tid; cb < g_hi; cb += 16 * ASM_THREADS) {
            const int64_t b0 = cb < g_lo ? g_lo : cb, b1 = cb + 16 < g_hi ? cb + 16 : g_hi;
            const int nb = (int)(b1 - b0);
            const int c_first = (int)(b0 - a_off);
            int c = c_first > 0 ? c_first - 1 : 0;       // one column early: the previous column's kind
            bool warm = c_first > 0;
            int lo = 0, hi = n_items;                   // last item with start <= c
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(io + mid) <= c) lo = mid; else hi = mid;
            }
            int i = lo;
            int start = __ldg(io + i);
            int end = i + 1 < n_items ? __ldg(io + i + 1) : aln_len;
            AsmSrc s = asm_resolve(i, c - start, end - start, R, po.n_slots, pd, gr, seqs, strbuf);
            unsigned long long w0l = 0, w0h = 0, w1l = 0, w1h = 0;
            int prev = 0, k = 0;
            const int c_end = c_first + nb;
            for (; c < c_end; ++c) {
                if (c == end) {
                    do { ++i; start = end; end = i + 1 < n_items ? __ldg(io + i + 1) : aln_len; } while (end == start);
                    s = asm_resolve(i, 0, end - start, R, po.n_slots, pd, gr, seqs, strbuf);
                }
                const uint32_t v0 = *s.p0, v1 = *s.p1;
                s.p0 += s.st0; s.p1 += s.st1;
                const int kind = v0 == '-' ? (v1 == '-' ? 3 : 1) : (v1 == '-' ? 2 : 0);
                if (warm) { warm = false; prev = kind; continue; }
                cnt_m += (kind == 0 && v0 == v1); cnt_x += (kind == 0 && v0 != v1);
                const bool gap = kind == 1 || kind == 2;
                cnt_e += (gap && prev == kind); cnt_o += (gap && prev != kind);
                prev = kind;
                const int sh = (k & 7) * 8;
                if (k < 8) { w0l |= (unsigned long long)v0 << sh; w1l |= (unsigned long long)v1 << sh; }
                else { w0h |= (unsigned long long)v0 << sh; w1h |= (unsigned long long)v1 << sh; }
                ++k;
            }
            if (nb == 16) {
                *reinterpret_cast<uint4 *>(aln0 + b0) = make_uint4((uint32_t)w0l, (uint32_t)(w0l >> 32), (uint32_t)w0h, (uint32_t)(w0h >> 32));
                *reinterpret_cast<uint4 *>(aln1 + b0) = make_uint4((uint32_t)w1l, (uint32_t)(w1l >> 32), (uint32_t)w1h, (uint32_t)(w1h >> 32));
            } else {
                for (int q = 0; q < nb; ++q) {
                    aln0[b0 + q] = (uint8_t)((q < 8 ? w0l >> (8 * q) : w0h >> (8 * (q - 8))) & 0xFF);
                    aln1[b0 + q] = (uint8_t)((q < 8 ? w1l >> (8 * q) : w1h >> (8 * (q - 8))) & 0xFF);
                }
            }
        }
        if (out.sop && !compact) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                cnt_m += __shfl_xor_sync(0xffffffffu, cnt_m, o); cnt_x += __shfl_xor_sync(0xffffffffu, cnt_x, o);
                cnt_o += __shfl_xor_sync(0xffffffffu, cnt_o, o); cnt_e += __shfl_xor_sync(0xffffffffu, cnt_e, o);
            }
            if (lane == 0) { atomicAdd(&s_cnt[0], cnt_m); atomicAdd(&s_cnt[1], cnt_x); atomicAdd(&s_cnt[2], cnt_o); atomicAdd(&s_cnt[3], cnt_e); }
            __syncthreads();
            if (tid < 4 && s_cnt[tid]) atomicAdd(out.sop + 4 * pair + tid, s_cnt[tid]);
        }
    }
}

// Clears the per-run counters and the status array with a kernel: a cudaMemsetAsync would go
// through a copy engine and queue behind the (large) sequence uploads of other lanes.
__global__ void __launch_bounds__(256) clear_kernel(Counters *ctr, int32_t *status, int64_t n, AsmScanRec *scan, int64_t nscan) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
    if (i < (int64_t)(sizeof(Counters) / 4)) reinterpret_cast<uint32_t *>(ctr)[i] = 0u;
    for (int64_t j = i; j < n; j += step) status[j] = 0;
    for (int64_t j = i; j < nscan; j += step) scan[j].flag = 0u;        // look-back flags of the assembler's scan
}
cudaError_t launch_clear(Counters *ctr, int32_t *status, int64_t n, AsmScanRec *scan, int64_t nscan, cudaStream_t st) {
    int64_t grid = (n + 255) / 256;
    if (grid < 1) grid = 1;
    if (grid > 1024) grid = 1024;
    clear_kernel<<<(int)grid, 256, 0, st>>>(ctr, status, n, scan, nscan);
    return cudaGetLastError();
}

// Copies a few words to pinned host memory (device-accessible under unified addressing) with
// plain stores: unlike a cudaMemcpyAsync this does not queue behind the bulk downloads in
// the copy engine, so the host learns a chunk's counters the moment its kernels are done.
__global__ void publish_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst_host, int words) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst_host[i] = src[i];
    __threadfence_system();
}
cudaError_t launch_publish(const void *src, void *dst_host, int bytes, cudaStream_t st) {
    publish_kernel<<<1, 64, 0, st>>>(reinterpret_cast<const uint32_t *>(src), reinterpret_cast<uint32_t *>(dst_host), bytes / 4);
    return cudaGetLastError();
}

cudaError_t launch_assemble(const uint8_t *seqs, const PairDesc *pairs, int64_t num_pairs, const PairOut *pout,
                            const int32_t *runs, const Slot *slots, const uint8_t *strbuf, int32_t *item_off,
                            AsmScanRec *scan, int64_t *tile_off, AsmOut out, uint8_t *aln0, uint8_t *aln1,
                            int32_t *runs_out, int32_t *seg_cols, Counters *ctr, int flags, int num_sms,
                            cudaStream_t st, int *launches) {
    if (num_pairs <= 0) return cudaSuccess;
    const int64_t nblk = asm_scan_blocks(num_pairs);
    asm_layout_kernel<<<(unsigned)nblk, LAY_THREADS, 0, st>>>(pout, num_pairs, slots, runs, item_off, scan, ctr, out, tile_off, flags);
    asm_copy_kernel<<<num_sms * 16, ASM_THREADS, 0, st>>>(seqs, pairs, num_pairs, pout, runs, item_off, ctr, strbuf, out, tile_off,
                                                          aln0, aln1, runs_out, seg_cols, flags);
    *launches += 2;
    return cudaGetLastError();
}

}  // namespace dbga
