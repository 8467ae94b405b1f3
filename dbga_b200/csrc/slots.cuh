// Segment-slot construction shared by the fused small-pair kernel and the generic stage-2
// kernel: slots [first, mid_1 .. mid_{R-1}, tail] of a pair with R runs of anchors
// (reference src/dbga/debruijn_pairwise.py:453-457, 463-538), their DP class, traceback
// placement and work-list entries.  Called by every thread of the CTA.
#pragma once
#include "kernels.cuh"

namespace dbga {

struct SlotCfg {
    int max_abs_score, go_raw, ge_raw;
};

// run(r, a, b, len) returns run r (ascending b).  flag_big: shared int, set on TOO_LARGE.
// tid / T: this thread's index among, and the number of, the threads that share the pair (a CTA, or all
// CTAs of a thread-block cluster); whole warps only.
template <class RunFn>
__device__ __forceinline__ void build_slots(int64_t pair, const PairDesc &pd, int R, RunFn run, int64_t slot_base,
                                            const Pools &pools, Counters *ctr, SlotCfg cfg, int *flag_big,
                                            unsigned long long &my_cells, unsigned long long &my_nseg,
                                            int tid = (int)threadIdx.x, int T = (int)blockDim.x) {
    const int lane = tid & 31;
    const int n_slots = R == 0 ? 1 : R + 1;
    const int64_t str_base = 2 * pd.off0;
    for (int base = 0; base < n_slots; base += T) {
        const int t = base + tid;
        int cls = CLS_NONE;
        Slot s;
        if (t < n_slots) {
            int x0 = 0, x1 = pd.n0, y0 = 0, y1 = pd.n1;
            if (R > 0) {
                if (t > 0) {
                    int a, b, len;
                    run(t - 1, a, b, len);
                    x0 = a + len - 1; y0 = b + len - 1;
                }
                if (t < R) {
                    int len;
                    run(t, x1, y1, len);
                }
            }
            s.x0 = x0; s.m = x1 - x0; s.y0 = y0; s.n = y1 - y0;
            s.score = 0; s.len = 0; s.pair = (int32_t)pair; s.tb_off = pd.off0;
            s.row_len = (int32_t)((pd.off1 - pd.off0) + pd.n1); s.n0 = (int32_t)(pd.off1 - pd.off0);
            s.str_off = str_base + x0 + y0;     // row x; row y at + row_len
            if (s.m > 0 && s.n > 0) {
                const long long cells = (long long)s.m * s.n;
                my_cells += (unsigned long long)cells; my_nseg += 1;
                const long long mag = (long long)cfg.max_abs_score * (s.m < s.n ? s.m : s.n) + cfg.go_raw +
                                      (long long)cfg.ge_raw * ((long long)s.m + s.n);
                if (2 * mag >= MAX_ABS_SCORE) { *flag_big = 1; }      // real score + the ge*(i+j) offset
                else if (s.m <= TINY_MAX && s.n <= TINY_MAX) {
                    const int mx = s.m > s.n ? s.m : s.n;
                    cls = NUM_TINY - 1;
#pragma unroll
                    for (int c = NUM_TINY - 2; c >= 0; --c) if (mx <= pools.tiny_bound[c]) cls = c;
                } else {
                    // one warp per segment, rows per lane chosen so that short segments still use
                    // all 32 lanes; huge segments (few pairs in the batch) get a whole CTA
                    cls = cells > pools.warp_max_cells ? CLS_CTA : (s.m <= 64 ? CLS_WARP2 : (s.m <= 128 ? CLS_WARP4 : (s.m <= 256 ? CLS_WARP8 : CLS_WARP16)));
                    if (cls == CLS_WARP16 && pools.strip_ok) {
                        // what the one-pass packed kernel cannot hold (wf16_ok: more than 1024 rows, or past the
                        // int16 ceiling) runs as packed row strips with the long class, not on the 32-bit warp kernel
                        const int r16 = s.m <= 512 ? 8 : 16;
                        const int mp = ((s.m + 2 * r16 - 1) / (2 * r16)) * (2 * r16);
                        const long long lim = mp < s.n + 1 ? mp : s.n + 1;
                        if (mp > 64 * r16 || (long long)pools.strip_per_step * lim + 64 > (long long)pools.strip_ulimit) cls = CLS_CTA;
                    }
                    if (cls == CLS_CTA) atomicMax(&ctr->long_max_m, (unsigned int)s.m);
                    const long long need = (tb_bytes_wavefront(s.m, s.n, wf_rows_per_lane(cls)) + 255) & ~255LL;
                    const long long off = (long long)atomicAdd(&ctr->tb_used, (unsigned long long)need);
                    if (off + need > pools.tb_cap) { *flag_big = 1; cls = CLS_NONE; }
                    s.tb_off = off;
                }
            } else {
                s.len = s.m + s.n;      // utils.py:187-192: one side empty -> all-gap row
            }
            s.cls = cls;
        }
        // warp-aggregated work-list append per DP class: the lanes of one class find each other with
        // one MATCH, the first of them reserves room for all
        {
            const unsigned peers = __match_any_sync(0xffffffffu, cls);
            const int leader = __ffs(peers) - 1;
            unsigned int pos = 0;
            if (cls != CLS_NONE && lane == leader) pos = atomicAdd(&ctr->n_cls[cls], (unsigned)__popc(peers));
            pos = __shfl_sync(0xffffffffu, pos, leader);
            if (cls != CLS_NONE)
                pools.lists[(int64_t)cls * pools.list_cap + pos + __popc(peers & ((1u << lane) - 1u))] = (int32_t)(slot_base + t);
        }
        if (t < n_slots) pools.slots[slot_base + t] = s;
    }
}

}  // namespace dbga
