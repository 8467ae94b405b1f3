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
constexpr int ASM_THREADS = 64;
constexpr int LAY_PAIRS = 64;
constexpr int LAY_THREADS = 256;

struct LayMeta { long long run_base, slot_base, n_anchors, off0, off1; int status, R, n_slots, row_len; };

// flag hand-over of the look-back: release store / acquire load at GPU scope (the record's
// payload is written before the flag and read after it)
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(LAY_THREADS, 4)
asm_layout_kernel(const PairOut *__restrict__ pout, const PairDesc *__restrict__ pairs, int64_t num_pairs,
                  const Slot *__restrict__ slots, const int32_t *__restrict__ runs, const uint8_t *__restrict__ strbuf,
                  int32_t *__restrict__ item_off,
                  long long *__restrict__ item_src, AsmScanRec *__restrict__ scan, Counters *__restrict__ ctr, AsmOut out,
                  AsmTile *__restrict__ tiles, int flags) {
    __shared__ long long s_v[3][LAY_PAIRS];
    __shared__ LayMeta s_meta[LAY_PAIRS];
    __shared__ long long s_excl[3], s_tot[3];
    __shared__ unsigned int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool compact = (flags & DBGA_FLAG_COMPACT_ROWS) != 0, no_runs = (flags & DBGA_FLAG_NO_RUNS) != 0 && !compact;
    const bool want_sop = (flags & DBGA_FLAG_SOP) != 0 && out.sop != nullptr;
    if (tid == 0) s_ticket = atomicAdd(&ctr->scan_ticket, 1u);
    __syncthreads();
    const int64_t blk = s_ticket, pair0 = blk * LAY_PAIRS;
    // metadata of the block's pairs: one round trip for all of them (thread t fetches pair t), then
    // every warp reads its pairs' records from shared memory
    constexpr int PER_WARP = LAY_PAIRS / (LAY_THREADS / 32);
    if (tid < LAY_PAIRS) {
        LayMeta mt;
        mt.status = DBGA_K_INVALID; mt.R = 0; mt.n_slots = 0; mt.run_base = 0; mt.slot_base = 0; mt.n_anchors = 0; mt.off0 = 0; mt.off1 = 0; mt.row_len = 0;
        if (pair0 + tid < num_pairs) {
            const PairOut po = pout[pair0 + tid];
            mt.status = po.status; mt.R = po.n_runs; mt.n_slots = po.n_slots; mt.run_base = po.run_base; mt.slot_base = po.slot_base;
            mt.n_anchors = po.n_anchors;
            mt.off0 = pairs[pair0 + tid].off0; mt.off1 = pairs[pair0 + tid].off1;
            mt.row_len = (int)(mt.off1 - mt.off0) + pairs[pair0 + tid].n1;
        }
        s_meta[tid] = mt;
    }
    __syncthreads();
    for (int j = 0; j < PER_WARP; ++j) {
        const int w = wid + (LAY_THREADS / 32) * j;
        const int64_t pair = pair0 + w;
        long long sc = 0, ce = 0, nrep = 0;
        int c_m = 0, c_x = 0, c_o = 0, c_e = 0;            // match, mismatch, gap_open, gap_extend
        unsigned int carry = 0, carry_run = 1;
        int ns = 0, ntile = 0;
        int st = s_meta[w].status;
        if (pair < num_pairs) {
            const int R = s_meta[w].R, n_slots = s_meta[w].n_slots;
            const long long run_base = s_meta[w].run_base, slot_base = s_meta[w].slot_base;
            const long long off0 = s_meta[w].off0, off1 = s_meta[w].off1;
            const long long n_anchors = s_meta[w].n_anchors;
            bool ok = st == DBGA_OK;
            const bool report = ok || st == DBGA_CYCLE;
            if (ok && n_slots > 0) {
                int32_t *io = item_off + 2 * slot_base;
                long long *is = item_src + 2 * slot_base;
                bool big = false;
                int carry_kind = 0;                 // kind of the last column emitted so far (0: both bases)
                // lane handles slot t and the run that follows it (items 2t, 2t + 1)
                for (int tb = 0; tb < n_slots; tb += 32) {
                    const int t = tb + lane;
                    unsigned int len_s = 0, len_r = 0, run_emit = 0;
                    long long src_s = 0, src_r = 0;
                    // alignment-quality counts of the slot's emitted columns (pairwise_alignment_score,
                    // reference src/dbga/utils.py:440-504), its first / last column kind, and what to classify
                    int first_kind = 0, last_kind = 0, q_len = 0, q_skip = 0;
                    long long q_rows = -1;          // scratch offset of the reversed row 1 of a DP slot
                    if (t < n_slots) {
                        const Slot *sp = slots + slot_base + t;
                        const int4 xy = *reinterpret_cast<const int4 *>(sp);             // (x0, m, y0, n)
                        const int2 sl = *reinterpret_cast<const int2 *>(&sp->score);     // (score, len)
                        int3 rn = make_int3(0, 0, 0);
                        if (t < R) { const int32_t *rp = runs + 3 * (run_base + t); rn = make_int3(rp[0], rp[1], rp[2]); }
                        const bool mid = R > 0 && t > 0 && t < n_slots - 1;
                        const int skip = (mid && sl.y > 0) ? 1 : 0;                      // debruijn_pairwise.py:499-500
                        len_s = (unsigned int)(sl.y - skip);
                        if (xy.y > 0 && xy.w > 0) {
                            sc += sl.x; ce += (long long)xy.y * xy.w; ++ns;
                            src_s = ((2 * off0 + xy.x + xy.z + (sl.y - 1 - skip)) << 2) | 0;      // reversed rows in the scratch
                            q_rows = 2 * off0 + xy.x + xy.z; q_len = sl.y; q_skip = skip;
                        } else if (xy.y > 0) {
                            src_s = ((off0 + xy.x + skip) << 2) | 1;                     // utils.py:187-188
                            if (len_s) { c_o += 1; c_e += len_s - 1; first_kind = last_kind = 2; }
                        } else {
                            src_s = ((off1 + xy.z + skip) << 2) | 2;                     // utils.py:189-190
                            if (len_s) { c_o += 1; c_e += len_s - 1; first_kind = last_kind = 1; }
                        }
                        if (t < R) {
                            run_emit = (unsigned int)(rn.z - (t == R - 1 ? 1 : 0));     // last merge node is left to the tail
                            if (!compact) len_r = run_emit;
                            src_r = ((off0 + rn.x) << 2) | 3;
                            c_m += run_emit;                                             // inside a run every column is a match
                        }
                    }
                    // ---- classify the columns of the DP slots: a lane walks its own short slot; long slots
                    //      are taken one at a time by the whole warp (32 columns per step)
                    const int row_len = s_meta[w].row_len;
                    if (!want_sop) q_rows = -1;
                    if (q_rows >= 0 && q_len - q_skip <= 64) {
                        const uint8_t *rx = strbuf + q_rows, *ry = rx + row_len;
                        int prev = 0;
                        for (int c = q_skip; c < q_len; ++c) {
                            const uint32_t b0 = rx[q_len - 1 - c], b1 = ry[q_len - 1 - c];
                            const int kind = b0 == '-' ? (b1 == '-' ? 3 : 1) : (b1 == '-' ? 2 : 0);
                            c_m += (kind == 0 && b0 == b1); c_x += (kind == 0 && b0 != b1);
                            const bool gap = kind == 1 || kind == 2;
                            c_e += (gap && prev == kind); c_o += (gap && prev != kind);
                            if (c == q_skip) first_kind = kind;
                            prev = kind;
                        }
                        last_kind = prev;
                    }
                    for (unsigned int todo = __ballot_sync(0xffffffffu, q_rows >= 0 && q_len - q_skip > 64); todo; todo &= todo - 1) {
                        const int L = __ffs(todo) - 1;
                        const long long rows = __shfl_sync(0xffffffffu, q_rows, L);
                        const int ln = __shfl_sync(0xffffffffu, q_len, L), sk = __shfl_sync(0xffffffffu, q_skip, L);
                        const uint8_t *rx = strbuf + rows, *ry = rx + row_len;
                        int am = 0, ax = 0, ao2 = 0, ae = 0;
                        unsigned int p1 = 0, p2 = 0;            // bit 0: the previous column was a gap in row 1 / row 2 only
                        int fk = 0, lk = 0;
                        for (int c0 = sk; c0 < ln; c0 += 32) {
                            const int c = c0 + lane;
                            uint32_t b0 = 'A', b1 = 'C';
                            const bool have = c < ln;
                            if (have) { b0 = rx[ln - 1 - c]; b1 = ry[ln - 1 - c]; }
                            const unsigned int vm = __ballot_sync(0xffffffffu, have);
                            const unsigned int g0 = __ballot_sync(0xffffffffu, have && b0 == '-'), g1b = __ballot_sync(0xffffffffu, have && b1 == '-');
                            const unsigned int eq = __ballot_sync(0xffffffffu, have && b0 == b1);
                            const unsigned int nuc = vm & ~(g0 | g1b), k1 = g0 & ~g1b, k2 = g1b & ~g0;
                            const unsigned int ext = (k1 & ((k1 << 1) | p1)) | (k2 & ((k2 << 1) | p2));
                            am += __popc(nuc & eq); ax += __popc(nuc & ~eq); ae += __popc(ext); ao2 += __popc((k1 | k2) & ~ext);
                            const int top = 31 - __clz((int)vm);
                            p1 = (k1 >> top) & 1u; p2 = (k2 >> top) & 1u;
                            lk = p1 ? 1 : (p2 ? 2 : (((g0 & g1b) >> top) & 1u ? 3 : 0));
                            if (c0 == sk) fk = (k1 & 1u) ? 1 : ((k2 & 1u) ? 2 : ((g0 & g1b & 1u) ? 3 : 0));
                        }
                        if (lane == L) { c_m += am; c_x += ax; c_o += ao2; c_e += ae; first_kind = fk; last_kind = lk; }
                    }
                    // ---- a slot's first column follows the previous run's last column (both bases) -- or, when
                    //      that run emitted nothing (the last run of a pair may), the previous slot's last column
                    if (want_sop) {
                        int prev_last = __shfl_up_sync(0xffffffffu, last_kind, 1);
                        const unsigned int prev_run = __shfl_up_sync(0xffffffffu, run_emit, 1);
                        const unsigned int prev_len = __shfl_up_sync(0xffffffffu, len_s, 1);
                        if (lane == 0) prev_last = carry_kind;
                        const bool joined = t > 0 && t < n_slots && len_s > 0 && (lane == 0 ? carry_run == 0 : prev_run == 0) &&
                                            (lane == 0 ? true : prev_len > 0);
                        if (joined && (first_kind == 1 || first_kind == 2) && first_kind == prev_last) { c_o -= 1; c_e += 1; }
                        // what the next block of 32 slots follows
                        const int my_last = run_emit > 0 ? 0 : (len_s > 0 ? last_kind : -1);      // -1: nothing emitted here
                        int lastk = carry_kind;
                        unsigned int lastrun = carry_run;
#pragma unroll 1
                        for (int l2 = 0; l2 < 32; ++l2) {
                            const int k2 = __shfl_sync(0xffffffffu, my_last, l2);
                            const unsigned int r2 = __shfl_sync(0xffffffffu, run_emit, l2);
                            if (tb + l2 < n_slots && k2 >= 0) { lastk = k2; lastrun = r2; }
                        }
                        carry_kind = lastk; carry_run = lastrun;
                    }
                    const unsigned int len = len_s + len_r;
                    unsigned int inc = len;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += v;
                    }
                    const unsigned int ex = carry + inc - len;
                    if (t < n_slots) {
                        if (R > 0 && t < R) {
                            *reinterpret_cast<int2 *>(io + 2 * t) = make_int2((int)ex, (int)(ex + len_s));
                            *reinterpret_cast<longlong2 *>(is + 2 * t) = make_longlong2(src_s, src_r);
                        } else {
                            io[R > 0 ? 2 * t : 0] = (int)ex;
                            is[R > 0 ? 2 * t : 0] = src_s;
                        }
                    }
                    const unsigned int add = __shfl_sync(0xffffffffu, inc, 31);
                    big = big || carry + add < carry || carry + add > 0x7fffffffu;      // 32-bit item offsets
                    carry += add;
                }
                if (big) { ok = false; st = DBGA_TOO_LARGE; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sc += __shfl_xor_sync(0xffffffffu, sc, o);
                ce += __shfl_xor_sync(0xffffffffu, ce, o);
                ns += __shfl_xor_sync(0xffffffffu, ns, o);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c_m += __shfl_xor_sync(0xffffffffu, c_m, o); c_x += __shfl_xor_sync(0xffffffffu, c_x, o);
                c_o += __shfl_xor_sync(0xffffffffu, c_o, o); c_e += __shfl_xor_sync(0xffffffffu, c_e, o);
            }
            if (!ok) { carry = 0; sc = 0; ce = 0; ns = 0; c_m = c_x = c_o = c_e = 0; }
            nrep = (report && !no_runs && st != DBGA_TOO_LARGE) ? R : 0;     // cycle pairs still report the anchors that cycle
            ntile = (int)((carry + ASM_TILE - 1) / ASM_TILE);
            if (ntile == 0 && (nrep > 0 || (compact && ok && n_slots > 0))) ntile = 1;   // runs / segment widths are copied by the pair's first tile
            if (lane == 0) {
                out.score[pair] = sc; out.cells[pair] = ce; out.nseg[pair] = ns;
                out.nanch[pair] = (ok || st == DBGA_CYCLE) ? n_anchors : 0;
                out.status[pair] = st;
                if (want_sop) *reinterpret_cast<int4 *>(out.sop + 4 * pair) = make_int4(c_m, c_x, c_o, c_e);
            }
        }
        if (lane == 0) { s_v[0][w] = carry; s_v[1][w] = nrep; s_v[2][w] = ntile; s_meta[w].status = st; }
    }
    __syncthreads();
    if (wid == 0) {
        // block-local exclusive prefixes (in place) and block totals of the three quantities
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
            const long long a0 = s_v[q][2 * lane], a1 = s_v[q][2 * lane + 1];
            long long inc = a0 + a1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            s_v[q][2 * lane] = inc - a0 - a1;
            s_v[q][2 * lane + 1] = inc - a1;
            if (lane == 31) s_tot[q] = inc;
        }
        __syncwarp();
        if (lane == 0) {
            AsmScanRec *me = scan + blk;
            long long ex0 = 0, ex1 = 0, ex2 = 0;
            if (blk > 0) {
                me->agg[0] = (unsigned long long)s_tot[0]; me->agg[1] = (unsigned long long)s_tot[1]; me->agg[2] = (unsigned long long)s_tot[2];
                st_release_u32(&me->flag, 1u);
                for (int64_t j = blk - 1; j >= 0; --j) {        // decoupled look-back
                    const AsmScanRec *r = scan + j;
                    unsigned int f;
                    do { f = ld_acquire_u32(&r->flag); } while (f == 0u);
                    const volatile unsigned long long *src = f == 2u ? r->pre : r->agg;
                    ex0 += (long long)src[0]; ex1 += (long long)src[1]; ex2 += (long long)src[2];
                    if (f == 2u) break;
                }
            }
            me->pre[0] = (unsigned long long)(ex0 + s_tot[0]); me->pre[1] = (unsigned long long)(ex1 + s_tot[1]);
            me->pre[2] = (unsigned long long)(ex2 + s_tot[2]);
            st_release_u32(&me->flag, 2u);
            s_excl[0] = ex0; s_excl[1] = ex1; s_excl[2] = ex2;
        }
        __syncwarp();
        const bool last_blk = pair0 + LAY_PAIRS >= num_pairs;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {                                    // the lane's two pairs
            const int w = 2 * lane + h;
            const int64_t p = pair0 + w;
            if (p >= num_pairs) break;
            const long long a_off = s_excl[0] + s_v[0][w], r_off = s_excl[1] + s_v[1][w], t_off = s_excl[2] + s_v[2][w];
            out.aln_off[p] = a_off; out.run_off[p] = r_off;
            const long long aln_len = (w + 1 < LAY_PAIRS ? s_v[0][w + 1] : s_tot[0]) - s_v[0][w];
            const long long ntile = (w + 1 < LAY_PAIRS ? s_v[2][w + 1] : s_tot[2]) - s_v[2][w];
            const LayMeta &mt = s_meta[w];
            AsmTile tr;
            tr.a_off = a_off; tr.item_base = 2 * mt.slot_base; tr.run_base = mt.run_base; tr.run_off = r_off;
            tr.aln_len = (int)aln_len; tr.row_len = mt.row_len; tr.pair = (int)p; tr.R = mt.R; tr.n_slots = mt.n_slots;
            tr.ok = mt.status == DBGA_OK && mt.n_slots > 0;
            tr.n_items = tr.ok ? mt.n_slots + mt.R : 0;
            for (long long t = 0; t < ntile; ++t) { tr.tp = (int)t; tiles[t_off + t] = tr; }     // one record per copy tile
        }
        if (lane == 0 && last_blk) {                                     // last block: grand totals
            out.aln_off[num_pairs] = s_excl[0] + s_tot[0]; out.run_off[num_pairs] = s_excl[1] + s_tot[1];
            ctr->total_aln = (unsigned long long)(s_excl[0] + s_tot[0]);
            ctr->total_runs = (unsigned long long)(s_excl[1] + s_tot[1]);
            ctr->total_tiles = (unsigned long long)(s_excl[2] + s_tot[2]);
        }
    }
}

// ---------------------------------------------------------------- copy
// Where the bytes of an item come from is resolved once, by the layout pass (item_src: byte offset
// << 2 | kind).  kind 0: DP segment, rows reversed in the scratch (stage 3 writes them walking
// back; row 2 sits row_len bytes after row 1), offset = scratch position of the item's first
// emitted column, step -1.  kind 1 / 2: one side empty (utils.py:187-190): sequence bytes of seq1 /
// seq2 against '-'.  kind 3: a run of anchors, both rows are the same seq1 bytes.
//
// asm_copy_kernel: a CTA takes one tile of ASM_TILE output columns of one pair (one AsmTile record
// written by the layout pass says everything about it).  Phase 1 is item-centric and stages the two
// rows in shared memory at the alignment of their final addresses; phase 2 is output-centric: a
// thread owns a 16-byte aligned chunk of both rows (128-bit stores).
constexpr int ASM_G = 4;                      // lanes per item in phase 1
constexpr int ASM_SBUF = ASM_TILE + 32;       // tile + alignment slack (multiple of 16)
constexpr int ASM_WIN = 256;                  // items whose offsets / sources are staged at a time

// four bytes at any address: two aligned words and a funnel shift.  At least one of the four bytes
// belongs to the caller's data; the others may lie up to three bytes before / after it (the caller
// masks them off).  The first aligned word is read only if it starts inside the buffer; the second
// shares its word with a wanted byte or lies in the 8 bytes of slack every buffer of the library
// has behind its last byte (include/dbga_b200.h: device sequences must be readable up to there).
__device__ __forceinline__ uint32_t ld_u32_any(const uint8_t *p, const uint32_t *lo_w) {
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p - (sh >> 3));
    const uint32_t lo = q >= lo_w ? __ldg(q) : 0u;
    const uint32_t hi = sh ? __ldg(q + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

__global__ void __launch_bounds__(ASM_THREADS, 16)
asm_copy_kernel(const uint8_t *__restrict__ seqs, const int32_t *__restrict__ runs, const int32_t *__restrict__ item_off,
                const long long *__restrict__ item_src, const Counters *__restrict__ ctr,
                const uint8_t *__restrict__ strbuf, const AsmTile *__restrict__ tiles,
                uint8_t *__restrict__ aln0, uint8_t *__restrict__ aln1, int32_t *__restrict__ runs_out,
                int32_t *__restrict__ seg_cols, int64_t seq_bytes, int flags) {
    __shared__ __align__(16) uint8_t sbuf[2][ASM_SBUF];
    __shared__ int s_io[ASM_WIN + 1];
    __shared__ long long s_is[ASM_WIN];
    __shared__ int s_ilo;
    const int tid = threadIdx.x, lane = tid & 31;
    const bool compact = (flags & DBGA_FLAG_COMPACT_ROWS) != 0, no_runs = (flags & DBGA_FLAG_NO_RUNS) != 0 && !compact;
    const long long total_tiles = (long long)ctr->total_tiles;
    const uint32_t *seq_lo = reinterpret_cast<const uint32_t *>(seqs), *str_lo = reinterpret_cast<const uint32_t *>(strbuf);
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        AsmTile tr;
        {
            const uint4 *q = reinterpret_cast<const uint4 *>(tiles + tile);
            uint4 *d = reinterpret_cast<uint4 *>(&tr);
            d[0] = __ldg(q); d[1] = __ldg(q + 1); d[2] = __ldg(q + 2); d[3] = __ldg(q + 3);
        }
        const int aln_len = tr.aln_len, R = tr.R, n_items = tr.n_items;
        const int32_t *io = item_off + tr.item_base;
        const long long *is = item_src + tr.item_base;
        if (tr.tp == 0) {
            const int32_t *gr = runs + 3 * tr.run_base;
            if (!no_runs && R > 0) for (int r = tid; r < 3 * R; r += ASM_THREADS) runs_out[3 * tr.run_off + r] = __ldg(gr + r);
            if (compact && tr.ok)        // columns of every segment, in output order (what dbga_expand_rows reads)
                for (int t = tid; t < tr.n_slots; t += ASM_THREADS) {
                    const int i = R > 0 ? 2 * t : 0;
                    seg_cols[tr.run_off + tr.pair + t] = (i + 1 < n_items ? __ldg(io + i + 1) : aln_len) - __ldg(io + i);
                }
        }
        const int c_lo = tr.tp * ASM_TILE, c_hi = min(c_lo + ASM_TILE, aln_len);
        if (c_hi <= c_lo) continue;
        const int64_t g_lo = tr.a_off + c_lo, g_hi = tr.a_off + c_hi;
        const int64_t base0 = g_lo & ~(int64_t)15;          // sbuf[r][x] is output byte base0 + x of row r
        __syncthreads();                                    // the previous tile's phase 2 is done with sbuf
        if (aln_len > ASM_TILE) {
            // first item of the tile: last item whose start is <= c_lo (warp 0, 32 probes per round)
            if (tid < 32) {
                int lo = 0, hi = n_items;                   // start[lo] <= c_lo < start[hi]
                while (hi - lo > 1) {
                    const int step = (hi - lo + 31) >> 5;
                    const int idx = lo + (lane + 1) * step;
                    const bool le = idx < hi && __ldg(io + idx) <= c_lo;
                    const int cnt = __popc(__ballot_sync(0xffffffffu, le));
                    const int nlo = lo + cnt * step;
                    hi = min(hi, nlo + step);
                    lo = nlo;
                }
                if (lane == 0) s_ilo = lo;
            }
        } else if (tid == 0) {
            s_ilo = 0;
        }
        for (int z = tid; z < 2 * ASM_SBUF / 16; z += ASM_THREADS) reinterpret_cast<uint4 *>(&sbuf[0][0])[z] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        // ---- phase 1: items -> shared memory, ASM_WIN items per window: their offsets and sources are
        //      staged first (one round trip); then 4 lanes per item move whole 32-bit words of the staged
        //      rows (a run feeds both rows from one fetch; the scratch rows of a DP segment are read
        //      backwards), bytes only in an item's first and last word; two items per group in flight
        for (int i0 = s_ilo; i0 < n_items; i0 += ASM_WIN) {
            const int nw = min(ASM_WIN, n_items - i0);
            for (int u = tid; u <= nw; u += ASM_THREADS) {
                s_io[u] = i0 + u < n_items ? __ldg(io + i0 + u) : aln_len;
                if (u < nw) s_is[u] = __ldg(is + i0 + u);
            }
            __syncthreads();
            const bool more = s_io[nw] < c_hi && i0 + nw < n_items;      // uniform: another window follows
            const int grp = tid / ASM_G, gl = tid % ASM_G;
            constexpr int NG = ASM_THREADS / ASM_G;
            // Runs (odd items) and segments (even items) alternate; a warp that took them as they come
            // would run both code paths on every step with a quarter of its lanes each.  So: first all
            // runs of the window, then all segments, each pass with one straight-line word loop.
            // The bytes of a word that lie outside the item are masked off and the word is OR-ed into the
            // (zeroed) buffer: an item's first and last word take the same path as the others, and two
            // items may share a word.
            const int par_run = (R > 0) ? 1 - (i0 & 1) : nw;              // window index of the first run (none when R == 0)
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int first = pass == 0 ? par_run : (R > 0 ? (i0 & 1) : 0);
                for (int u0 = first; u0 < nw && s_io[u0] < c_hi; u0 += 4 * NG) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int u = u0 + 2 * (grp + h * NG);
                        if (u >= nw) continue;
                        const int start = s_io[u], end = s_io[u + 1];
                        if (start >= c_hi || end <= c_lo || end == start) continue;
                        const long long rec = s_is[u];
                        const int from = max(start, c_lo), to = min(end, c_hi);
                        const int q = (int)(tr.a_off + from - base0), n = to - from;      // bytes [q, q + n) of sbuf
                        if (pass == 0) {
                            const uint8_t *p = seqs + (rec >> 2) + (from - start);
                            for (int wq = (q >> 2) + gl; 4 * wq < q + n; wq += ASM_G) {
                                const int d = 4 * wq - q;                                  // item column of the word's byte 0
                                uint32_t mk = 0xFFFFFFFFu;
                                if (d < 0) mk <<= 8 * (-d);
                                if (d + 4 > n) mk &= 0xFFFFFFFFu >> (8 * (d + 4 - n));
                                const uint32_t v = ld_u32_any(p + d, seq_lo) & mk;       // inside a run both rows are the same bytes
                                atomicOr(reinterpret_cast<uint32_t *>(sbuf[0] + 4 * wq), v);
                                atomicOr(reinterpret_cast<uint32_t *>(sbuf[1] + 4 * wq), v);
                            }
                        } else {
                            const int kind = (int)(rec & 3);
                            // rows of a DP segment sit reversed in the scratch (column c at p - c), a one-sided
                            // segment is sequence bytes against '-'
                            const uint8_t *p = (kind == 0 ? strbuf - (from - start) : seqs + (from - start)) + (rec >> 2);
                            for (int wq = (q >> 2) + gl; 4 * wq < q + n; wq += ASM_G) {
                                const int d = 4 * wq - q;
                                uint32_t mk = 0xFFFFFFFFu;
                                if (d < 0) mk <<= 8 * (-d);
                                if (d + 4 > n) mk &= 0xFFFFFFFFu >> (8 * (d + 4 - n));
                                uint32_t v0, v1;
                                if (kind == 0) {
                                    v0 = __byte_perm(ld_u32_any(p - d - 3, str_lo), 0u, 0x0123);
                                    v1 = __byte_perm(ld_u32_any(p + tr.row_len - d - 3, str_lo), 0u, 0x0123);
                                } else if (kind == 1) { v0 = ld_u32_any(p + d, seq_lo); v1 = 0x2D2D2D2Du; }
                                else { v0 = 0x2D2D2D2Du; v1 = ld_u32_any(p + d, seq_lo); }
                                atomicOr(reinterpret_cast<uint32_t *>(sbuf[0] + 4 * wq), v0 & mk);
                                atomicOr(reinterpret_cast<uint32_t *>(sbuf[1] + 4 * wq), v1 & mk);
                            }
                        }
                    }
                }
            }
            if (!more) break;
            __syncthreads();                                // s_io / s_is are reused by the next window
        }
        __syncthreads();
        // ---- phase 2: shared memory -> both rows, 16 bytes per thread; the bytes of the pair's first and
        //      last chunk that belong to it go out one per thread
        const int64_t f_lo = (g_lo + 15) & ~(int64_t)15, f_hi = g_hi & ~(int64_t)15;     // whole chunks: [f_lo, f_hi)
        for (int64_t cb = f_lo + 16 * (int64_t)tid; cb < f_hi; cb += 16 * ASM_THREADS) {
            *reinterpret_cast<uint4 *>(aln0 + cb) = *reinterpret_cast<const uint4 *>(sbuf[0] + (cb - base0));
            *reinterpret_cast<uint4 *>(aln1 + cb) = *reinterpret_cast<const uint4 *>(sbuf[1] + (cb - base0));
        }
        {
            const int64_t h_hi = f_lo < g_hi ? f_lo : g_hi;                              // head bytes [g_lo, h_hi)
            const int64_t t_lo = f_hi > h_hi ? f_hi : h_hi;                              // tail bytes [t_lo, g_hi)
            const int64_t hb = g_lo + (tid & 15), tb2 = t_lo + (tid & 15);
            if (tid < 16 && hb < h_hi) { aln0[hb] = sbuf[0][hb - base0]; aln1[hb] = sbuf[1][hb - base0]; }
            if (tid >= 16 && tid < 32 && tb2 < g_hi) { aln0[tb2] = sbuf[0][tb2 - base0]; aln1[tb2] = sbuf[1][tb2 - base0]; }
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
                            long long *item_src, AsmScanRec *scan, AsmTile *tiles, AsmOut out,
                            uint8_t *aln0, uint8_t *aln1, int32_t *runs_out, int32_t *seg_cols, Counters *ctr, int64_t seq_bytes,
                            int flags, int num_sms, cudaStream_t st, int *launches) {
    if (num_pairs <= 0) return cudaSuccess;
    const int64_t nblk = asm_scan_blocks(num_pairs);
    asm_layout_kernel<<<(unsigned)nblk, LAY_THREADS, 0, st>>>(pout, pairs, num_pairs, slots, runs, strbuf, item_off, item_src, scan, ctr,
                                                              out, tiles, flags);
    asm_copy_kernel<<<num_sms * 16, ASM_THREADS, 0, st>>>(seqs, runs, item_off, item_src, ctr, strbuf, tiles, aln0, aln1,
                                                          runs_out, seg_cols, seq_bytes, flags);
    *launches += 2;
    return cudaGetLastError();
}

}  // namespace dbga
