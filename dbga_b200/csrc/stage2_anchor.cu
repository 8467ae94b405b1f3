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
#include <algorithm>
#include <cooperative_groups.h>

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
                              int max_abs_score, int go_raw, int ge_raw, int cluster_above) {
    __shared__ ScanSmem sm;
    // one bit per seq-2 position: "a run starts here" / "a run ends here" (pass 1 -> pass 2)
    __shared__ uint32_t maskS[S2_CLUSTER_ABOVE / 32 + 1], maskE[S2_CLUSTER_ABOVE / 32 + 1];
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
        if (N1 > cluster_above) continue;              // stage2_cluster_kernel takes the very long pairs
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
        // (the ballots are the warp's counts and, kept as one bit per position, all that pass 2 needs: run boundaries
        // are rare -- a couple of hundred in 30 000 positions -- so pass 2 walks 32 positions per lane-word and
        // touches aof[] again only where a run starts)
        int cS = 0, cE = 0;
        long long cA = 0;
        for (int base = q_lo; base < q_hi; base += 32 * U) {
            int a[U]; bool isS[U], isE[U];
            load_block(base, a, isS, isE);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned bS = __ballot_sync(0xffffffffu, isS[u]), bE = __ballot_sync(0xffffffffu, isE[u]);
                cA += a[u] >= 0; cS += __popc(bS); cE += __popc(bE);
                if (lane == 0 && base + 32 * u < q_hi) { maskS[(base >> 5) + u] = bS; maskE[(base >> 5) + u] = bE; }
            }
        }
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
        // ---- pass 2: ordered compaction of run starts (a, b) and run ends (b), per warp chunk, from the bit masks:
        //      a lane takes one word of 32 positions, the warp scans the words' counts, every set bit is one store
        for (int w0 = q_lo >> 5; q_lo < q_hi && 32 * w0 < q_hi; w0 += 32) {      // (an empty chunk starts at N1, not at a word)
            const int w = w0 + lane;
            const bool have = 32 * w < q_hi;
            uint32_t mS = have ? maskS[w] : 0u, mE = have ? maskE[w] : 0u;
            int nS = __popc(mS), nE = __popc(mE), iS = nS, iE = nE;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int vS = __shfl_up_sync(0xffffffffu, iS, o), vE = __shfl_up_sync(0xffffffffu, iE, o);
                if (lane >= o) { iS += vS; iE += vE; }
            }
            int idxS = offS + iS - nS, idxE = offE + iE - nE;
            while (mS) {
                const int q = 32 * w + __ffs((int)mS) - 1;
                mS &= mS - 1;
                runs[3 * (int64_t)idxS] = a_of[q];
                runs[3 * (int64_t)idxS + 1] = q;
                ++idxS;
            }
            while (mE) {
                const int q = 32 * w + __ffs((int)mE) - 1;
                mE &= mE - 1;
                runs[3 * (int64_t)idxE + 2] = q;                   // end b, turned into len below
                ++idxE;
            }
            offS += __shfl_sync(0xffffffffu, iS, 31); offE += __shfl_sync(0xffffffffu, iE, 31);
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

// ---------------------------------------------------------------------------------------------
// The same stage 2 for a very long pair (more than S2_CLUSTER_ABOVE seq-2 positions: BASELINE's
// 1 Mb pair), spread over a thread-block cluster: 8 CTAs x 512 threads scan, compact and build
// slots together.  Per-CTA counts and flags are exchanged through distributed shared memory,
// cluster.sync() orders the phases (one CTA took 0.48 ms for 1 M positions).
constexpr int S2_CL = 8;
struct ClusterSmem {
    int warpS[32], warpE[32];
    int totS, totE, flag_cycle, flag_big;
    long long totA, bcast[2];
    long long red[32];
};

__global__ void __launch_bounds__(512)
stage2_cluster_kernel(const PairDesc *__restrict__ pairs, const int32_t *__restrict__ pair_list, int64_t num_pairs, int k,
                      int want_segments, const int32_t *__restrict__ aof, const int32_t *__restrict__ status,
                      PairOut *__restrict__ pout, Pools pools, Counters *__restrict__ ctr, int max_abs_score, int go_raw,
                      int ge_raw, int cluster_above) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ ClusterSmem sm;
    const int crank = (int)cluster.block_rank();
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    const int gtid = crank * T + tid, GT = S2_CL * T, gwid = crank * nw + wid, gnw = S2_CL * nw;
    auto peer_of = [&](int c) -> const ClusterSmem * { return cluster.map_shared_rank(&sm, c); };
    unsigned long long my_cells = 0, my_nseg = 0;
    const int64_t n_clusters = gridDim.x / S2_CL, cid = blockIdx.x / S2_CL;

    for (int64_t it = cid; it < num_pairs; it += n_clusters) {
        const int64_t pair = pair_list ? pair_list[it] : it;
        const PairDesc pd = pairs[pair];
        const int N1 = pd.n1 - k + 1;
        if (N1 <= cluster_above) continue;              // stage2_kernel takes it (uniform over the cluster)
        const int st = status[pair];
        PairOut po;
        po.status = st; po.n_runs = 0; po.run_base = 0; po.slot_base = 0; po.n_slots = 0; po.pad = 0;
        po.n_anchors = 0; po.aln_len = 0;
        if (st != DBGA_OK) {
            if (gtid == 0) pout[pair] = po;
            continue;
        }
        const int32_t *a_of = aof + pd.q_base;
        const int chunk = (((N1 + gnw - 1) / gnw) + 31) & ~31;
        const int q_lo = min(N1, gwid * chunk), q_hi = min(N1, q_lo + chunk);
        constexpr int U = 8;                             // (eight independent loads per lane: 128 warps sweep 1 M positions, latency-bound)
        auto load_block = [&](int base, int (&a)[U], bool (&isS)[U], bool (&isE)[U]) {      // as in stage2_kernel
            int edge_lo = -1, edge_hi = -1;
            if (lane == 0 && base > 0) edge_lo = a_of[base - 1];
            if (lane == 31 && base + 32 * U < N1) edge_hi = a_of[base + 32 * U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = base + 32 * u + lane;
                a[u] = q < N1 ? a_of[q] : -1;
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
        // ---- pass 1: counts per warp, per CTA, per cluster
        int cS = 0, cE = 0;
        long long cA = 0;
        for (int base = q_lo; base < q_hi; base += 32 * U) {
            int a[U]; bool isS[U], isE[U];
            load_block(base, a, isS, isE);
#pragma unroll
            for (int u = 0; u < U; ++u) { cA += a[u] >= 0; cS += isS[u]; cE += isE[u]; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cS += __shfl_xor_sync(0xffffffffu, cS, o); cE += __shfl_xor_sync(0xffffffffu, cE, o);
            cA += __shfl_xor_sync(0xffffffffu, cA, o);
        }
        if (lane == 0) { sm.warpS[wid] = cS; sm.warpE[wid] = cE; sm.red[wid] = cA; }
        __syncthreads();
        if (tid == 0) {
            int s = 0, e = 0; long long a = 0;
            for (int w = 0; w < nw; ++w) { s += sm.warpS[w]; e += sm.warpE[w]; a += sm.red[w]; }
            sm.totS = s; sm.totE = e; sm.totA = a; sm.flag_cycle = 0; sm.flag_big = 0;
        }
        cluster.sync();
        int R = 0, offS = 0, offE = 0;
        long long M = 0;
#pragma unroll
        for (int c = 0; c < S2_CL; ++c) {
            const int s = peer_of(c)->totS, e = peer_of(c)->totE;
            if (c < crank) { offS += s; offE += e; }
            R += s; M += peer_of(c)->totA;
        }
        for (int w = 0; w < wid; ++w) { offS += sm.warpS[w]; offE += sm.warpE[w]; }
        if (gtid == 0) {
            long long base = 0;
            if (R > 0) {
                base = (long long)atomicAdd(&ctr->runs_used, (unsigned long long)R);
                if (base + R > pools.runs_cap) { atomicExch(&ctr->overflow, 1u); base = -1; }
            }
            sm.bcast[0] = base;
        }
        cluster.sync();
        const long long run_base = peer_of(0)->bcast[0];
        if (run_base < 0) {      // pool exhausted: host grows the pool and re-runs stage 2
            if (gtid == 0) { po.status = DBGA_TOO_LARGE; pout[pair] = po; }
            cluster.sync();
            continue;
        }
        int32_t *runs = pools.runs + 3 * run_base;
        // ---- pass 2: ordered compaction
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
                if (isE[u]) runs[3 * (int64_t)(offE + __popc(bE & lt)) + 2] = q;
                offS += __popc(bS); offE += __popc(bE);
            }
        }
        __threadfence();
        cluster.sync();
        for (int r = gtid; r < R; r += GT) runs[3 * (int64_t)r + 2] = runs[3 * (int64_t)r + 2] - runs[3 * (int64_t)r + 1] + 1;
        __threadfence();
        cluster.sync();
        for (int r = gtid + 1; r < R; r += GT) {
            const int prev_end = runs[3 * (int64_t)(r - 1)] + runs[3 * (int64_t)(r - 1) + 2] - 1;
            if (runs[3 * (int64_t)r] <= prev_end) sm.flag_cycle = 1;
        }
        cluster.sync();
        bool cyc = false;
#pragma unroll
        for (int c = 0; c < S2_CL; ++c) cyc = cyc || peer_of(c)->flag_cycle != 0;
        po.n_runs = R; po.run_base = run_base; po.n_anchors = M;
        if (cyc || !want_segments) {
            if (gtid == 0) { if (cyc) po.status = DBGA_CYCLE; pout[pair] = po; }
            cluster.sync();                              // peers are done reading this pair's shared state
            continue;
        }
        const int n_slots = R == 0 ? 1 : R + 1;
        if (gtid == 0) {
            long long base = (long long)atomicAdd(&ctr->slots_used, (unsigned long long)n_slots);
            if (base + n_slots > pools.slots_cap) { atomicExch(&ctr->overflow, 1u); base = -1; }
            sm.bcast[1] = base;
        }
        cluster.sync();
        const long long slot_base = peer_of(0)->bcast[1];
        if (slot_base < 0) {
            if (gtid == 0) { po.status = DBGA_TOO_LARGE; pout[pair] = po; }
            cluster.sync();
            continue;
        }
        build_slots(pair, pd, R,
                    [&](int r, int &a, int &b, int &len) {
                        a = runs[3 * (int64_t)r]; b = runs[3 * (int64_t)r + 1]; len = runs[3 * (int64_t)r + 2];
                    },
                    slot_base, pools, ctr, SlotCfg{max_abs_score, go_raw, ge_raw}, &sm.flag_big, my_cells, my_nseg, gtid, GT);
        cluster.sync();
        bool big = false;
#pragma unroll
        for (int c = 0; c < S2_CL; ++c) big = big || peer_of(c)->flag_big != 0;
        po.slot_base = slot_base; po.n_slots = n_slots;
        if (gtid == 0) {
            if (big) po.status = DBGA_TOO_LARGE;
            pout[pair] = po;
        }
        cluster.sync();
    }
    // per-CTA totals
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_cells += __shfl_xor_sync(0xffffffffu, my_cells, o);
        my_nseg += __shfl_xor_sync(0xffffffffu, my_nseg, o);
    }
    if (lane == 0 && (my_cells | my_nseg)) { atomicAdd(&ctr->cells, my_cells); atomicAdd(&ctr->nseg, my_nseg); }
}

cudaError_t launch_stage2(const PairDesc *pairs, const int32_t *pair_list, int64_t num_pairs, int k, bool no_anchors, bool want_segments,
                          const int32_t *aof, int32_t *status, PairOut *out, Pools pools, Counters *ctr,
                          int threads, int num_sms, int max_abs_score, int go_raw, int ge_raw, int64_t cluster_pairs,
                          cudaStream_t st) {
    if (num_pairs <= 0) return cudaSuccess;
    int64_t grid = (int64_t)num_sms * (2048 / threads);
    if (grid > num_pairs) grid = num_pairs;
    // cluster_pairs > 0: that many pairs have more than S2_CLUSTER_ABOVE seq-2 positions and go to the cluster kernel
    const int above = cluster_pairs > 0 ? S2_CLUSTER_ABOVE : 0x7fffffff;
    stage2_kernel<<<(int)grid, threads, 0, st>>>(pairs, pair_list, num_pairs, k, no_anchors ? 1 : 0, want_segments ? 1 : 0, aof,
                                                status, out, pools, ctr, max_abs_score, go_raw, ge_raw, above);
    if (cluster_pairs > 0) {
        cudaLaunchConfig_t cfg = {};
        const int64_t ncl = std::min<int64_t>(cluster_pairs, std::max(1, num_sms / S2_CL));
        cfg.gridDim = dim3((unsigned)(ncl * S2_CL));
        cfg.blockDim = dim3(512);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = S2_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, stage2_cluster_kernel, pairs, pair_list, num_pairs, k, want_segments ? 1 : 0, aof,
                           (const int32_t *)status, out, pools, ctr, max_abs_score, go_raw, ge_raw, above);
    }
    return cudaGetLastError();
}

}  // namespace dbga
