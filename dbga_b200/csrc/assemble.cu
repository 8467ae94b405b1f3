// Result assembly: turns runs of anchors + per-segment alignments into the two gapped rows
// the reference builds by string concatenation in DeBruijnPairwise.alignment()
// (src/dbga/debruijn_pairwise.py:499-543): first segment whole, every merge node's first
// character except the last (:508-509), mid segments with their first column dropped
// (:499-500), tail segment whole (:521-538).  Output is compacted with exclusive scans
// over pairs so that only the used bytes travel back to the host.
#include "kernels.cuh"

namespace dbga {

// ---------------------------------------------------------------- per-pair layout
// One warp per pair walks the pair's items in output order (slot 0, run 0, slot 1, run 1, ...,
// run R-1, slot R) 32 at a time, scans their lengths and records where every slot and run
// starts inside the pair's rows.  The copy kernels below are then item-parallel, so a pair
// with a million columns is assembled by the whole chip, not by one warp.
__global__ void __launch_bounds__(256)
pair_layout_kernel(const PairOut *__restrict__ pout, int64_t num_pairs, Slot *__restrict__ slots,
                   const int32_t *__restrict__ runs, int2 *__restrict__ run_meta, int64_t *__restrict__ aln_len,
                   int64_t *__restrict__ run_cnt, int64_t *__restrict__ score, int64_t *__restrict__ cells,
                   int32_t *__restrict__ nseg, int64_t *__restrict__ n_anchors, int32_t *__restrict__ status_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t pair = warp; pair < num_pairs; pair += nwarps) {
        const PairOut po = pout[pair];
        const bool ok = po.status == DBGA_OK, report = ok || po.status == DBGA_CYCLE;
        const int R = po.n_runs;
        long long carry = 0, sc = 0, ce = 0;
        int ns = 0;
        if (ok && po.n_slots > 0) {
            const int n_items = po.n_slots + R;
            for (int ib = 0; ib < n_items; ib += 32) {
                const int it = ib + lane;
                long long len = 0;
                bool is_slot = false;
                int t = 0;
                if (it < n_items) {
                    is_slot = (R == 0) || (it % 2 == 0);
                    t = it / 2;
                    if (is_slot) {
                        const Slot s = slots[po.slot_base + t];
                        const bool mid = R > 0 && t > 0 && t < po.n_slots - 1;
                        const int skip = (mid && s.len > 0) ? 1 : 0;
                        len = s.len - skip;
                        slots[po.slot_base + t].skip = skip;
                        if (s.m > 0 && s.n > 0) { sc += s.score; ce += (long long)s.m * s.n; ++ns; }
                    } else {
                        len = runs[3 * (po.run_base + t) + 2] - (t == R - 1 ? 1 : 0);   // last merge node is left to the tail
                    }
                }
                long long inc = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                const long long off = carry + inc - len;
                carry += __shfl_sync(0xffffffffu, inc, 31);
                if (it < n_items) {
                    if (is_slot) slots[po.slot_base + t].out_off = (int32_t)off;
                    else run_meta[po.run_base + t] = make_int2((int)off, (int)pair);
                }
            }
        } else if (R > 0 && po.run_base >= 0) {
            // no alignment (cycle, or a limit was hit): runs are only reported / ignored
            for (int r = lane; r < R; r += 32) run_meta[po.run_base + r] = make_int2(-1, report ? (int)pair : -1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sc += __shfl_xor_sync(0xffffffffu, sc, o);
            ce += __shfl_xor_sync(0xffffffffu, ce, o);
            ns += __shfl_xor_sync(0xffffffffu, ns, o);
        }
        if (lane == 0) {
            aln_len[pair] = ok ? carry : 0;
            run_cnt[pair] = report ? R : 0;       // cycle pairs still report the anchors that cycle
            score[pair] = sc; cells[pair] = ce; nseg[pair] = ns;
            n_anchors[pair] = report ? po.n_anchors : 0;
            status_out[pair] = po.status;
        }
    }
}

// ---------------------------------------------------------------- exclusive scan (int64)
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ long long block_excl_scan(long long v, long long *total, long long *sm) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) sm[wid] = inc;
    __syncthreads();
    long long pre = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        if (w < wid) pre += sm[w];
        tot += sm[w];
    }
    *total = tot;
    return pre + inc - v;
}

__global__ void __launch_bounds__(SCAN_T)
scan_tile_sums(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ tile_sums) {
    __shared__ long long sm[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    long long v = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int64_t idx = base + (int64_t)i * SCAN_T + threadIdx.x;
        if (idx < n) v += in[idx];
    }
    long long tot;
    block_excl_scan(v, &tot, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_T)
scan_tiles(int64_t *__restrict__ tile_sums, int64_t ntiles, int64_t *__restrict__ grand_total) {
    __shared__ long long sm[32];
    long long carry = 0;
    for (int64_t base = 0; base < ntiles; base += SCAN_T) {
        const int64_t idx = base + threadIdx.x;
        const long long v = idx < ntiles ? tile_sums[idx] : 0;
        long long tot;
        const long long ex = block_excl_scan(v, &tot, sm);
        if (idx < ntiles) tile_sums[idx] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void __launch_bounds__(SCAN_T)
scan_apply(const int64_t *__restrict__ in, int64_t n, const int64_t *__restrict__ tile_sums,
           int64_t *__restrict__ out /* n+1 */) {
    __shared__ long long sm[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    long long vals[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { vals[i] = (base + i < n) ? in[base + i] : 0; sum += vals[i]; }
    long long tot;
    long long run = tile_sums[blockIdx.x] + block_excl_scan(sum, &tot, sm);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += vals[i];
        if (base + i == n - 1) out[n] = run;
    }
}

static cudaError_t exclusive_scan(const int64_t *in, int64_t n, int64_t *out, int64_t *tmp, int64_t *total_dev,
                                  cudaStream_t st, int *launches) {
    const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums<<<(int)ntiles, SCAN_T, 0, st>>>(in, n, tmp);
    scan_tiles<<<1, SCAN_T, 0, st>>>(tmp, ntiles, total_dev);
    scan_apply<<<(int)ntiles, SCAN_T, 0, st>>>(in, n, tmp, out);
    *launches += 3;
    return cudaGetLastError();
}

// ---------------------------------------------------------------- row assembly
// Item-parallel copies, 8 lanes per item (a slot is ~10 columns, a run ~40): one kernel
// over the slot pool, one over the run pool.  An entry is trusted only if the pair it names
// currently owns that pool index (pools are not cleared between runs).
constexpr int ASM_G = 8;

__global__ void __launch_bounds__(256)
assemble_slots_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, int64_t num_pairs,
                      const PairOut *__restrict__ pout, const Slot *__restrict__ slots, int64_t slots_cap,
                      const Counters *__restrict__ ctr, const uint8_t *__restrict__ strbuf,
                      const int64_t *__restrict__ aln_off, uint8_t *__restrict__ aln0, uint8_t *__restrict__ aln1) {
    const int gl = threadIdx.x & (ASM_G - 1);
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / ASM_G;
    const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / ASM_G;
    int64_t total = (int64_t)ctr->slots_used;
    if (total > slots_cap) total = slots_cap;
    for (int64_t idx = grp; idx < total; idx += ngrp) {
        const Slot s = slots[idx];
        if (s.pair < 0 || s.pair >= num_pairs) continue;
        const PairOut po = pout[s.pair];
        if (po.status != DBGA_OK || idx < po.slot_base || idx >= po.slot_base + po.n_slots) continue;
        const int len = s.len - s.skip;
        if (len <= 0) continue;
        const PairDesc pd = pairs[s.pair];
        uint8_t *o0 = aln0 + aln_off[s.pair] + s.out_off, *o1 = aln1 + aln_off[s.pair] + s.out_off;
        if (s.m > 0 && s.n > 0) {        // reversed rows in scratch: column c is at [len_full - 1 - c]
            const uint8_t *rx = strbuf + s.str_off, *ry = rx + ((pd.off1 - pd.off0) + pd.n1);
            for (int c = gl; c < len; c += ASM_G) {
                o0[c] = rx[s.len - 1 - (c + s.skip)];
                o1[c] = ry[s.len - 1 - (c + s.skip)];
            }
        } else {                         // one side empty (utils.py:187-190)
            const uint8_t *s0 = seqs + pd.off0 + s.x0, *s1 = seqs + pd.off1 + s.y0;
            for (int c = gl; c < len; c += ASM_G) {
                o0[c] = s.m > 0 ? s0[c + s.skip] : (uint8_t)'-';
                o1[c] = s.m > 0 ? (uint8_t)'-' : s1[c + s.skip];
            }
        }
    }
}

__global__ void __launch_bounds__(256)
assemble_runs_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, int64_t num_pairs,
                     const PairOut *__restrict__ pout, const int32_t *__restrict__ runs, const int2 *__restrict__ run_meta,
                     int64_t runs_cap, const Counters *__restrict__ ctr, const int64_t *__restrict__ aln_off,
                     const int64_t *__restrict__ run_off, uint8_t *__restrict__ aln0, uint8_t *__restrict__ aln1,
                     int32_t *__restrict__ runs_out) {
    const int gl = threadIdx.x & (ASM_G - 1);
    const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / ASM_G;
    const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / ASM_G;
    int64_t total = (int64_t)ctr->runs_used;
    if (total > runs_cap) total = runs_cap;
    for (int64_t idx = grp; idx < total; idx += ngrp) {
        const int2 meta = run_meta[idx];
        const int pair = meta.y;
        if (pair < 0 || pair >= num_pairs) continue;
        const PairOut po = pout[pair];
        if (idx < po.run_base || idx >= po.run_base + po.n_runs) continue;
        if (po.status != DBGA_OK && po.status != DBGA_CYCLE) continue;
        const int a = runs[3 * idx], b = runs[3 * idx + 1], len = runs[3 * idx + 2];
        if (gl < 3) runs_out[3 * (run_off[pair] + (idx - po.run_base)) + gl] = gl == 0 ? a : (gl == 1 ? b : len);
        if (po.status != DBGA_OK || meta.x < 0) continue;
        const int n = len - (idx == po.run_base + po.n_runs - 1 ? 1 : 0);
        const PairDesc pd = pairs[pair];
        const uint8_t *s0 = seqs + pd.off0 + a;
        uint8_t *o0 = aln0 + aln_off[pair] + meta.x, *o1 = aln1 + aln_off[pair] + meta.x;
        // inside a run the two sequences are identical (the anchors' k-mers overlap), so one load feeds both rows
        for (int c = gl; c < n; c += ASM_G) { const uint8_t v = s0[c]; o0[c] = v; o1[c] = v; }
    }
}

// Clears the per-run counters and the status array with a kernel: a cudaMemsetAsync would go
// through a copy engine and queue behind the (large) sequence uploads of other lanes.
__global__ void __launch_bounds__(256) clear_kernel(Counters *ctr, int32_t *status, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)(sizeof(Counters) / 4)) reinterpret_cast<uint32_t *>(ctr)[i] = 0u;
    for (int64_t j = i; j < n; j += (int64_t)gridDim.x * blockDim.x) status[j] = 0;
}
cudaError_t launch_clear(Counters *ctr, int32_t *status, int64_t n, cudaStream_t st) {
    int64_t grid = (n + 255) / 256;
    if (grid < 1) grid = 1;
    if (grid > 1024) grid = 1024;
    clear_kernel<<<(int)grid, 256, 0, st>>>(ctr, status, n);
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
                            const int32_t *status, const int32_t *runs, int64_t runs_cap, int2 *run_meta, Slot *slots,
                            int64_t slots_cap, const uint8_t *strbuf, int64_t *aln_len, int64_t *aln_off,
                            int64_t *run_cnt, int64_t *run_off, int64_t *scan_tmp, uint8_t *aln0, uint8_t *aln1,
                            int32_t *runs_out, int64_t *score, int64_t *cells, int32_t *nseg, int64_t *n_anchors,
                            Counters *ctr, bool want_rows, int num_sms, cudaStream_t st, int *launches) {
    if (num_pairs <= 0) return cudaSuccess;
    int64_t grid = (num_pairs + 7) / 8;
    if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
    pair_layout_kernel<<<(int)grid, 256, 0, st>>>(pout, num_pairs, slots, runs, run_meta, aln_len, run_cnt, score, cells,
                                                 nseg, n_anchors, const_cast<int32_t *>(status));
    *launches += 1;
    cudaError_t e = exclusive_scan(aln_len, num_pairs, aln_off, scan_tmp, reinterpret_cast<int64_t *>(&ctr->total_aln), st, launches);
    if (e != cudaSuccess) return e;
    e = exclusive_scan(run_cnt, num_pairs, run_off, scan_tmp, reinterpret_cast<int64_t *>(&ctr->total_runs), st, launches);
    if (e != cudaSuccess) return e;
    const int cgrid = num_sms * 8;
    if (want_rows) {
        assemble_slots_kernel<<<cgrid, 256, 0, st>>>(seqs, pairs, num_pairs, pout, slots, slots_cap, ctr, strbuf, aln_off, aln0, aln1);
        *launches += 1;
    }
    assemble_runs_kernel<<<cgrid, 256, 0, st>>>(seqs, pairs, num_pairs, pout, runs, run_meta, runs_cap, ctr, aln_off, run_off,
                                               aln0, aln1, runs_out);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace dbga
