// Stage 2: merge-node (shared anchor) detection and ordering.
//
// Replaces the reference's Python graph walk: merge_node_idx in seq-2 order
// (src/dbga/debruijn_pairwise.py:94-98), extract_bubble (:277-300, the O(N*M) list scan),
// debruijn_merge_correctness (src/dbga/utils.py:266-291) and the segment bookkeeping of
// alignment() (:463-538).  Input: aof[q] from stage 1.  Per pair (one CTA):
//   * ordered stream compaction (warp ballot + block scan) of the *runs* of consecutive
//     anchors (a+u, b+u) -- inside a run every inter-anchor segment is the reference's
//     shortcut (:384-388), so only run boundaries produce DP work;
//   * colinearity ("cycle") check by adjacent difference over the compacted runs;
//   * the segment slots [first, mid_1 .. mid_{R-1}, tail] with their DP class, scratch
//     placement and work-list entries.
#include "slots.cuh"

namespace dbga {

struct ScanSmem {
    int warpS[32], warpE[32];
    int totS, totE;
    long long red[32];
    long long bcast[4];
    int flag_cycle, flag_big;
};

__device__ __forceinline__ long long block_sum(long long v, ScanSmem &sm) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sm.red[wid] = v;
    __syncthreads();
    long long t = 0;
    for (int w = 0; w < nw; ++w) t += sm.red[w];
    return t;
}

__global__ void stage2_kernel(const PairDesc *__restrict__ pairs, const int32_t *__restrict__ pair_list, int64_t num_pairs, int k, int no_anchors,
                              int want_segments, const int32_t *__restrict__ aof, const int32_t *__restrict__ status,
                              PairOut *__restrict__ pout, Pools pools, Counters *__restrict__ ctr,
                              int max_abs_score, int go_raw, int ge_raw) {
    __shared__ ScanSmem sm;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = (T + 31) >> 5;
    unsigned long long my_cells = 0, my_nseg = 0;

    for (int64_t it = blockIdx.x; it < num_pairs; it += gridDim.x) {
        const int64_t pair = pair_list ? pair_list[it] : it;
        const PairDesc pd = pairs[pair];
        int st = status[pair];
        PairOut po;
        po.status = st; po.n_runs = 0; po.run_base = 0; po.slot_base = 0; po.n_slots = 0; po.pad = 0;
        po.n_anchors = 0; po.aln_len = 0;
        if (st != DBGA_OK) {
            if (tid == 0) pout[pair] = po;
            continue;
        }
        const int N1 = no_anchors ? 0 : pd.n1 - k + 1;
        const int32_t *a_of = aof + pd.q_base;
        // ---- pass 1: every warp owns one contiguous chunk of seq-2 positions and counts its
        //      run starts, run ends and anchors (no block barrier inside the scan loops)
        const int chunk = (((N1 + nw - 1) / nw) + 31) & ~31;
        const int q_lo = min(N1, wid * chunk), q_hi = min(N1, q_lo + chunk);
        // A block of U x 32 consecutive positions per iteration: every lane loads its U values back to back
        // (independent loads: one memory round trip per block instead of one per 32 positions) and gets
        // its neighbours' values by shuffle; only the block's two edge neighbours are extra loads.
        constexpr int U = 4;
        auto load_block = [&](int base, int (&a)[U], bool (&isS)[U], bool (&isE)[U]) {
            int edge_lo = -1, edge_hi = -1;
            if (lane == 0 && base > 0) edge_lo = a_of[base - 1];
            if (lane == 31 && base + 32 * U < N1) edge_hi = a_of[base + 32 * U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + 32 * u + lane;
                a[u] = q < N1 ? a_of[q] : -1;          // values past the warp's own range (q >= q_hi) are neighbours only
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + 32 * u + lane;
                int pv = __shfl_up_sync(0xffffffffu, a[u], 1), nx = __shfl_down_sync(0xffffffffu, a[u], 1);
                const int pv0 = u > 0 ? __shfl_sync(0xffffffffu, a[u > 0 ? u - 1 : 0], 31) : edge_lo;
                const int nx31 = u + 1 < U ? __shfl_sync(0xffffffffu, a[u + 1 < U ? u + 1 : U - 1], 0) : edge_hi;
                if (lane == 0) pv = pv0;
                if (lane == 31) nx = nx31;
                const bool own = q < q_hi && a[u] >= 0;
                isS[u] = own && !(pv >= 0 && pv + 1 == a[u]);
                isE[u] = own && !(nx >= 0 && nx == a[u] + 1);
                if (q >= q_hi) a[u] = -1;
            }
        };
        int cS = 0, cE = 0;
        long long cA = 0;
        for (int base = q_lo; base < q_hi; base += 32 * U) {
            int a[U]; bool isS[U], isE[U];
            load_block(base, a, isS, isE);
#pragma unroll
            for (int u = 0; u < U; ++u) { cA += a[u] >= 0; cS += isS[u]; cE += isE[u]; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { cS += __shfl_xor_sync(0xffffffffu, cS, o); cE += __shfl_xor_sync(0xffffffffu, cE, o); }
        __syncthreads();
        if (lane == 0) { sm.warpS[wid] = cS; sm.warpE[wid] = cE; }
        __syncthreads();
        int R = 0, offS = 0, offE = 0;
        for (int w = 0; w < nw; ++w) {
            const int s = sm.warpS[w], e = sm.warpE[w];
            if (w < wid) { offS += s; offE += e; }
            R += s;
        }
        const long long M = block_sum(cA, sm);
        if (tid == 0) {
            sm.flag_cycle = 0; sm.flag_big = 0;
            long long base = 0;
            if (R > 0) {
                base = (long long)atomicAdd(&ctr->runs_used, (unsigned long long)R);
                if (base + R > pools.runs_cap) { atomicExch(&ctr->overflow, 1u); base = -1; }
            }
            sm.bcast[0] = base;
        }
        __syncthreads();
        const long long run_base = sm.bcast[0];
        if (run_base < 0) {      // pool exhausted: host grows the pool and re-runs stage 2
            if (tid == 0) { po.status = DBGA_TOO_LARGE; pout[pair] = po; }
            __syncthreads();
            continue;
        }
        int32_t *runs = pools.runs + 3 * run_base;
        // ---- pass 2: ordered compaction of run starts (a, b) and run ends (b), per warp chunk
        for (int base = q_lo; base < q_hi; base += 32 * U) {
            int a[U]; bool isS[U], isE[U];
            load_block(base, a, isS, isE);
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + 32 * u + lane;
                const unsigned bS = __ballot_sync(0xffffffffu, isS[u]), bE = __ballot_sync(0xffffffffu, isE[u]);
                if (isS[u]) {
                    const int idx = offS + __popc(bS & lt);
                    runs[3 * (int64_t)idx] = a[u];
                    runs[3 * (int64_t)idx + 1] = q;
                }
                if (isE[u]) runs[3 * (int64_t)(offE + __popc(bE & lt)) + 2] = q;   // end b, turned into len below
                offS += __popc(bS); offE += __popc(bE);
            }
        }
        __syncthreads();
        // ---- run lengths, then the colinearity check (utils.py:266-291): seq-1 positions of
        //      the merge nodes must be strictly increasing along seq 2
        for (int r = tid; r < R; r += T) runs[3 * (int64_t)r + 2] = runs[3 * (int64_t)r + 2] - runs[3 * (int64_t)r + 1] + 1;
        __syncthreads();
        for (int r = tid + 1; r < R; r += T) {
            const int prev_end = runs[3 * (int64_t)(r - 1)] + runs[3 * (int64_t)(r - 1) + 2] - 1;
            if (runs[3 * (int64_t)r] <= prev_end) sm.flag_cycle = 1;
        }
        __syncthreads();
        po.n_runs = R; po.run_base = run_base; po.n_anchors = M;
        if (sm.flag_cycle) {
            if (tid == 0) { po.status = DBGA_CYCLE; pout[pair] = po; }
            __syncthreads();
            continue;
        }
        if (!want_segments) {
            if (tid == 0) pout[pair] = po;
            __syncthreads();
            continue;
        }
        // ---- segment slots (debruijn_pairwise.py:453-457, 463-538)
        const int n_slots = R == 0 ? 1 : R + 1;
        if (tid == 0) {
            long long base = (long long)atomicAdd(&ctr->slots_used, (unsigned long long)n_slots);
            if (base + n_slots > pools.slots_cap) { atomicExch(&ctr->overflow, 1u); base = -1; }
            sm.bcast[1] = base;
        }
        __syncthreads();
        const long long slot_base = sm.bcast[1];
        if (slot_base < 0) {
            if (tid == 0) { po.status = DBGA_TOO_LARGE; pout[pair] = po; }
            __syncthreads();
            continue;
        }
        build_slots(pair, pd, R,
                    [&](int r, int &a, int &b, int &len) {
                        a = runs[3 * (int64_t)r]; b = runs[3 * (int64_t)r + 1]; len = runs[3 * (int64_t)r + 2];
                    },
                    slot_base, pools, ctr, SlotCfg{max_abs_score, go_raw, ge_raw}, &sm.flag_big, my_cells, my_nseg);
        __syncthreads();
        po.slot_base = slot_base; po.n_slots = n_slots;
        if (tid == 0) {
            if (sm.flag_big) po.status = DBGA_TOO_LARGE;
            pout[pair] = po;
        }
        __syncthreads();
    }
    // per-CTA totals
    const unsigned long long c = (unsigned long long)block_sum((long long)my_cells, sm);
    const unsigned long long s = (unsigned long long)block_sum((long long)my_nseg, sm);
    if (tid == 0 && (c | s)) { atomicAdd(&ctr->cells, c); atomicAdd(&ctr->nseg, s); }
}

cudaError_t launch_stage2(const PairDesc *pairs, const int32_t *pair_list, int64_t num_pairs, int k, bool no_anchors, bool want_segments,
                          const int32_t *aof, int32_t *status, PairOut *out, Pools pools, Counters *ctr,
                          int threads, int num_sms, int max_abs_score, int go_raw, int ge_raw, cudaStream_t st) {
    if (num_pairs <= 0) return cudaSuccess;
    int64_t grid = (int64_t)num_sms * (2048 / threads);
    if (grid > num_pairs) grid = num_pairs;
    stage2_kernel<<<(int)grid, threads, 0, st>>>(pairs, pair_list, num_pairs, k, no_anchors ? 1 : 0, want_segments ? 1 : 0, aof,
                                                status, out, pools, ctr, max_abs_score, go_raw, ge_raw);
    return cudaGetLastError();
}

}  // namespace dbga
