// Result assembly: turns runs of anchors + per-segment alignments into the two gapped rows
// the reference builds by string concatenation in DeBruijnPairwise.alignment()
// (src/dbga/debruijn_pairwise.py:499-543): first segment whole, every merge node's first
// character except the last (:508-509), mid segments with their first column dropped
// (:499-500), tail segment whole (:521-538).  Output is compacted with exclusive scans
// over pairs so that only the used bytes travel back to the host.
#include "kernels.cuh"

namespace dbga {

// ---------------------------------------------------------------- per-pair lengths
__global__ void __launch_bounds__(256)
pair_len_kernel(const PairOut *__restrict__ pout, int64_t num_pairs, const Slot *__restrict__ slots,
                int64_t *__restrict__ aln_len, int64_t *__restrict__ run_cnt, int64_t *__restrict__ score,
                int64_t *__restrict__ cells, int32_t *__restrict__ nseg, int64_t *__restrict__ n_anchors,
                int32_t *__restrict__ status_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t pair = warp; pair < num_pairs; pair += nwarps) {
        const PairOut po = pout[pair];
        long long len = 0, sc = 0, ce = 0;
        int ns = 0;
        if (po.status == DBGA_OK) {
            for (int t = lane; t < po.n_slots; t += 32) {
                const Slot s = slots[po.slot_base + t];
                const bool mid = po.n_runs > 0 && t > 0 && t < po.n_slots - 1;
                len += s.len - (mid && s.len > 0 ? 1 : 0);
                if (s.m > 0 && s.n > 0) { sc += s.score; ce += (long long)s.m * s.n; ++ns; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            len += __shfl_xor_sync(0xffffffffu, len, o);
            sc += __shfl_xor_sync(0xffffffffu, sc, o);
            ce += __shfl_xor_sync(0xffffffffu, ce, o);
            ns += __shfl_xor_sync(0xffffffffu, ns, o);
        }
        if (lane == 0) {
            const bool ok = po.status == DBGA_OK;
            if (ok && po.n_slots > 0 && po.n_anchors > 0) len += po.n_anchors - 1;
            aln_len[pair] = ok ? len : 0;
            // runs are reported for aligned pairs and for cycle pairs (the anchors that cycle)
            run_cnt[pair] = (ok || po.status == DBGA_CYCLE) ? po.n_runs : 0;
            score[pair] = sc; cells[pair] = ce; nseg[pair] = ns;
            n_anchors[pair] = (ok || po.status == DBGA_CYCLE) ? po.n_anchors : 0;
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
// One warp per pair.  Items in output order: slot 0, run 0, slot 1, run 1, ..., run R-1,
// slot R (tail).  Lengths are scanned 32 items at a time; every item is then copied by the
// whole warp.
__global__ void __launch_bounds__(256)
assemble_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, int64_t num_pairs,
                const PairOut *__restrict__ pout, const int32_t *__restrict__ runs, const Slot *__restrict__ slots,
                const uint8_t *__restrict__ strbuf, const int64_t *__restrict__ aln_off,
                const int64_t *__restrict__ run_off, uint8_t *__restrict__ aln0, uint8_t *__restrict__ aln1,
                int32_t *__restrict__ runs_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t pair = warp; pair < num_pairs; pair += nwarps) {
        const PairOut po = pout[pair];
        if (po.status != DBGA_OK && po.status != DBGA_CYCLE) continue;
        const int R = po.n_runs;
        const int32_t *pr = runs + 3 * po.run_base;
        {   // compact copy of the runs (anchors)
            int32_t *dst = runs_out + 3 * run_off[pair];
            for (int i = lane; i < 3 * R; i += 32) dst[i] = pr[i];
        }
        if (po.status != DBGA_OK || po.n_slots == 0) continue;
        const PairDesc pd = pairs[pair];
        const uint8_t *s0 = seqs + pd.off0, *s1 = seqs + pd.off1;
        const int64_t row_len = (pd.off1 - pd.off0) + pd.n1;
        uint8_t *o0 = aln0 + aln_off[pair], *o1 = aln1 + aln_off[pair];
        const int n_items = po.n_slots + R;          // R == 0 -> just the one slot
        long long carry = 0;
        for (int ib = 0; ib < n_items; ib += 32) {
            const int it = ib + lane;
            // item description
            int kind = -1;       // 0: DP slot (reversed rows in scratch) 1: gap slot 2: run copy
            long long len = 0, src = 0;
            int skip = 0, sm_ = 0, sn_ = 0, sx0 = 0, sy0 = 0;
            if (it < n_items) {
                const bool is_slot = (R == 0) || (it % 2 == 0);
                const int t = it / 2;
                if (is_slot) {
                    const Slot s = slots[po.slot_base + t];
                    const bool mid = R > 0 && t > 0 && t < po.n_slots - 1;
                    skip = (mid && s.len > 0) ? 1 : 0;
                    len = s.len - skip;
                    sm_ = s.m; sn_ = s.n; sx0 = s.x0; sy0 = s.y0;
                    if (s.m > 0 && s.n > 0) { kind = 0; src = s.str_off; len = s.len - skip; sm_ = s.len; }
                    else kind = 1;
                } else {
                    kind = 2;
                    sx0 = pr[3 * t]; sy0 = pr[3 * t + 1];
                    len = pr[3 * t + 2] - (t == R - 1 ? 1 : 0);   // last merge node is left to the tail
                }
            }
            // exclusive scan of len over the 32 items
            long long inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const long long off = carry + inc - len;
            carry += __shfl_sync(0xffffffffu, inc, 31);
            const int cnt = min(32, n_items - ib);
            for (int u = 0; u < cnt; ++u) {
                const int kd = __shfl_sync(0xffffffffu, kind, u);
                const long long ln = __shfl_sync(0xffffffffu, len, u);
                const long long of = __shfl_sync(0xffffffffu, off, u);
                const long long sr = __shfl_sync(0xffffffffu, src, u);
                const int sk = __shfl_sync(0xffffffffu, skip, u);
                const int a0 = __shfl_sync(0xffffffffu, sx0, u), b0 = __shfl_sync(0xffffffffu, sy0, u);
                const int mm = __shfl_sync(0xffffffffu, sm_, u);
                if (kd == 0) {           // reversed rows: column c of the segment is at [full_len-1-c]
                    const uint8_t *rx = strbuf + sr, *ry = rx + row_len;
                    for (long long c = lane; c < ln; c += 32) {
                        o0[of + c] = rx[mm - 1 - (c + sk)];
                        o1[of + c] = ry[mm - 1 - (c + sk)];
                    }
                } else if (kd == 1) {    // one side empty (utils.py:187-190)
                    for (long long c = lane; c < ln; c += 32) {
                        o0[of + c] = mm > 0 ? s0[a0 + c + sk] : (uint8_t)'-';
                        o1[of + c] = mm > 0 ? (uint8_t)'-' : s1[b0 + c + sk];
                    }
                } else if (kd == 2) {
                    for (long long c = lane; c < ln; c += 32) {
                        o0[of + c] = s0[a0 + c];
                        o1[of + c] = s1[b0 + c];
                    }
                }
            }
        }
    }
}

cudaError_t launch_assemble(const uint8_t *seqs, const PairDesc *pairs, int64_t num_pairs, const PairOut *pout,
                            const int32_t *status, const int32_t *runs, const Slot *slots, const uint8_t *strbuf,
                            int64_t *aln_len, int64_t *aln_off, int64_t *run_cnt, int64_t *run_off,
                            int64_t *scan_tmp, uint8_t *aln0, uint8_t *aln1, int32_t *runs_out,
                            int64_t *score, int64_t *cells, int32_t *nseg, int64_t *n_anchors,
                            Counters *ctr, int num_sms, cudaStream_t st, int *launches) {
    if (num_pairs <= 0) return cudaSuccess;
    int64_t grid = (num_pairs + 7) / 8;
    if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
    pair_len_kernel<<<(int)grid, 256, 0, st>>>(pout, num_pairs, slots, aln_len, run_cnt, score, cells, nseg,
                                              n_anchors, const_cast<int32_t *>(status));
    *launches += 1;
    cudaError_t e = exclusive_scan(aln_len, num_pairs, aln_off, scan_tmp, reinterpret_cast<int64_t *>(&ctr->total_aln), st, launches);
    if (e != cudaSuccess) return e;
    e = exclusive_scan(run_cnt, num_pairs, run_off, scan_tmp, reinterpret_cast<int64_t *>(&ctr->total_runs), st, launches);
    if (e != cudaSuccess) return e;
    assemble_kernel<<<(int)grid, 256, 0, st>>>(seqs, pairs, num_pairs, pout, runs, slots, strbuf, aln_off, run_off,
                                              aln0, aln1, runs_out);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace dbga
