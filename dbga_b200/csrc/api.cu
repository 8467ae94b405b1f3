// C-ABI host side: plan (sizing, classing, metadata upload), run (kernel pipeline on one
// stream, no host round trips), fetch (device -> pinned host).  See include/dbga_b200.h.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "kernels.cuh"

using namespace dbga;

namespace dbga { bool host_pack_seq(const uint8_t *src, int64_t len, uint32_t *dst); }   // hostpack.cpp

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};
struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct SmallClass {
    uint32_t cap;
    int parts;          // table partitions that take turns in shared memory (medium pairs; small pairs: 1)
    bool medium;        // stage 1 only, by-reference slots; false: fused stages 1+2 ("small" pairs)
    int words, maxq;
    int begin, count;   // range in the small pair list
};
struct LargeWave {
    int begin, count;   // range in the large list / LargeDesc array
    int max_n;
    size_t table_bytes;
};

}  // namespace

struct dbga_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int num_sms = 148;
    char err[512] = {0};
    int64_t scratch_limit = 8LL << 30;
    cudaEvent_t ev[8] = {};

    // ---- plan
    bool planned = false, ran = false;
    int64_t P = 0, total_bytes = 0, sumN0 = 0, sumN1 = 0;
    int64_t seq_extent = 0;       // bytes of the device sequence buffer the plan's offsets reach (relative to its base)
    int max_n0 = 0, max_n1 = 0;
    dbga_params prm{};
    bool no_anchors = false;
    DpParams dp{};
    int max_abs_score = 0;
    std::vector<PairDesc> h_pairs;
    std::vector<int32_t> h_list;          // small classes first, then large
    std::vector<SmallClass> classes;
    std::vector<LargeDesc> h_large;
    std::vector<LargeWave> waves;
    int large_begin = 0;
    int stage2_begin = 0;      // first list entry that needs the generic stage-2 kernel (medium + large pairs)
    int stage2_threads = 128;
    int64_t n_long_pairs = 0;  // pairs whose stage 2 runs on a thread-block cluster (more than S2_CLUSTER_ABOVE seq-2 positions)
    int64_t runs_cap = 0, slots_cap = 0, tb_cap = 0;
    int64_t bnd_stride_warp = 0, bnd_stride_cta = 0;
    int grid_warp = 0, grid_cta = 0, grid_cluster = 0;
    int64_t warp_max_cells = WARP_CLASS_MAX_CELLS_FEW;
    int tiny_bound[NUM_TINY] = {8, 16, 24, 32};
    bool seen_work[NUM_CLS] = {true, true, true, true, true, true, true, true, true};   // adaptive grids for the (often empty) wavefront launches
    int launches = 0;
    const uint8_t *d_seqs_run = nullptr;

    // ---- device buffers (grow only)
    DevBuf d_seqs, d_pairs, d_list, d_large, d_status, d_aof, d_nd0, d_nd1, d_pout, d_runs, d_item_off, d_slots, d_lists,
        d_tt, d_ctr, d_tb, d_bnd_w, d_bnd_c, d_str, d_packed, d_table, d_aln_off, d_tile_off,
        d_run_off, d_scan, d_aln0, d_aln1, d_runs_out, d_score, d_cells, d_nseg, d_nanch, d_status_out,
        d_sop, d_segcols, d_pk_in, d_item_src, d_tile_pair, d_sink;
    // ---- pinned host result buffers
    HostBuf h_status, h_score, h_cells, h_nseg, h_nanch, h_aln_off, h_run_off, h_aln0, h_aln1, h_runs, h_nd0,
        h_nd1, h_poff, h_qoff, h_ctr, h_meta, h_sop, h_segcols;
    float ms_plan = 0.f;
    // ---- pipelined dbga_align_batch: child contexts (own stream + workspaces) used round-robin
    static constexpr int NL = 4;
    dbga_ctx *lanes[NL] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_ready = nullptr;
    // sub-chunked upload of a lane's sequences: piece s is complete when ev_sub[s] fires;
    // stage 1+2 for the pairs of piece s is launched right behind it
    static constexpr int MAX_SUB = 16;
    cudaStream_t meta_stream = nullptr;   // lanes: upload plan metadata on the shared upload stream
    cudaEvent_t ev_meta = nullptr;
    // parent only: one download stream for all lanes.  A lane stream that has issued a copy keeps
    // a foot in the copy engine's queue: the next event / kernel on it was observed to wait behind
    // OTHER lanes' bulk downloads.  Lane streams therefore carry kernels only.
    cudaStream_t dl_stream = nullptr;
    cudaEvent_t ev_dl = nullptr;          // lanes: the chunk's download has left the lane's result buffers
    bool dl_pending = false;
    cudaStream_t copy_stream = nullptr;   // parent only: one upload stream shared by all lanes (uploads are serial on
                                          // the H2D engine anyway, and every extra stream risks aliasing onto one of the
                                          // CUDA_DEVICE_MAX_CONNECTIONS hardware queues, which creates false dependencies)
    cudaEvent_t ev_sub[MAX_SUB] = {};
    int n_sub = 0;
    int64_t sub_pair_lo[MAX_SUB + 1] = {};
    cudaEvent_t dbg_ev[4] = {};       // DBGA_TRACE: upload begin / end, download begin / end
    // Pipeline lanes write the per-pair result columns straight into the parent context's
    // whole-batch device arrays (at the chunk's pair offset), so they travel to the host in
    // seven copies per batch instead of seven small copies per chunk.
    struct MetaPtrs {
        int32_t *status = nullptr; int64_t *score = nullptr, *cells = nullptr; int32_t *nseg = nullptr;
        int64_t *nanch = nullptr, *aln_off = nullptr, *run_off = nullptr; int32_t *sop = nullptr;
    } ext;
    bool use_ext = false;
    int multi_peers = 1;      // contexts driven side by side by one dbga_align_batch_multi call
    // Host waits.  Spinning (CUDA's default on an idle box) answers fastest, but every rank of an
    // 8-GPU job keeps two threads spinning; with fewer than four host cores per GPU the waits
    // sleep instead (cudaEventBlockingSync events, condition variable between producer and consumer).
    bool spin = true;
    cudaEvent_t ev_block = nullptr;
    // long-segment DP (strip kernel + its traceback) runs beside the other DP classes: a few warps with a long
    // dependent chain on one stream, the throughput kernels on the other
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace {

int fail(dbga_ctx *ctx, int code, const char *msg) {
    snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    return code;
}

bool decide_spin(int peers) {
    if (const char *e = getenv("DBGA_SPIN")) return atoi(e) != 0;
    int local = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) local = std::max(1, atoi(e));
    const int cores = (int)std::max(1u, std::thread::hardware_concurrency());
    return cores / std::max(peers, local) >= 6;
}

// waits for everything queued on st so far
cudaError_t wait_stream(dbga_ctx *ctx, cudaStream_t st) {
    if (ctx->spin || !ctx->ev_block) return cudaStreamSynchronize(st);
    cudaError_t e = cudaEventRecord(ctx->ev_block, st);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->ev_block);
}

int dev_reserve(dbga_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return DBGA_SUCCESS;
    // (rare) growth: a lane's downloads run on the parent's download stream, so wait for the whole device
    if (b.p) { DBGA_CUDA_CHECK(cudaDeviceSynchronize()); DBGA_CUDA_CHECK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = std::max<size_t>(bytes, 256);
    want = (want + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        b.p = nullptr;
        return DBGA_ERR_NOMEM;
    }
    b.cap = want;
    return DBGA_SUCCESS;
}
int host_reserve(dbga_ctx *ctx, HostBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return DBGA_SUCCESS;
    if (b.p) { DBGA_CUDA_CHECK(cudaFreeHost(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = std::max<size_t>(bytes + bytes / 8, 4096);
    cudaError_t e = cudaHostAlloc(&b.p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "cudaHostAlloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        b.p = nullptr;
        return DBGA_ERR_NOMEM;
    }
    b.cap = want;
    return DBGA_SUCCESS;
}
#define RESERVE(buf, bytes)                                         \
    do {                                                            \
        int _r = dev_reserve(ctx, ctx->buf, (size_t)(bytes));       \
        if (_r != DBGA_SUCCESS) return _r;                          \
    } while (0)
#define HRESERVE(buf, bytes)                                        \
    do {                                                            \
        int _r = host_reserve(ctx, ctx->buf, (size_t)(bytes));      \
        if (_r != DBGA_SUCCESS) return _r;                          \
    } while (0)

uint32_t pow2_at_least(uint64_t v, uint32_t lo) {
    uint64_t c = lo;
    while (c < v) c <<= 1;
    return (uint32_t)c;
}

const int ORDER_PRIO[6][3] = {
    // priority (2 = wins ties) of {M, X, Y} for XMY, XYM, MXY, MYX, YXM, YMX
    {1, 2, 0}, {0, 2, 1}, {2, 1, 0}, {2, 0, 1}, {0, 1, 2}, {1, 0, 2}};

int make_dp_params(dbga_ctx *ctx, const dbga_params &p) {
    if (p.pred_order < 0 || p.pred_order > 5 || p.end_order < 0 || p.end_order > 5)
        return fail(ctx, DBGA_ERR_ARG, "pred_order / end_order must be in 0..5");
    if (p.gap_open < 0 || p.gap_extend < 0 || p.gap_open > 100000 || p.gap_extend > 100000)
        return fail(ctx, DBGA_ERR_ARG, "gap costs must be in 0..100000");
    DpParams &d = ctx->dp;
    d.go_raw = p.gap_len_mode ? p.gap_open + p.gap_extend : p.gap_open;
    d.ge_raw = p.gap_extend;
    d.go = SCALE * d.go_raw;
    d.ge = SCALE * d.ge_raw;
    d.tagM = TAGMUL * ORDER_PRIO[p.pred_order][0];
    d.tagX = TAGMUL * ORDER_PRIO[p.pred_order][1];
    d.tagY = TAGMUL * ORDER_PRIO[p.pred_order][2];
    d.endM = ORDER_PRIO[p.end_order][0];
    d.endX = ORDER_PRIO[p.end_order][1];
    d.endY = ORDER_PRIO[p.end_order][2];
    d.xy = p.xy_allowed ? 1 : 0;
    int mx = 0;
    for (int a = 0; a < 4; ++a) {
        uint32_t v[4];
        for (int c = 0; c < 4; ++c) {
            const int s = p.score[a * 4 + c];
            mx = std::max(mx, std::abs(s));
            if (std::abs(s) > 500) return fail(ctx, DBGA_ERR_ARG, "|substitution score| must be <= 500");
            v[c] = (uint32_t)(uint16_t)(int16_t)(SCALE * s + d.tagM);
        }
        d.lut_lo[a] = v[0] | (v[1] << 16);
        d.lut_hi[a] = v[2] | (v[3] << 16);
    }
    ctx->max_abs_score = mx;
    // packed 16-bit form (stage3_dp16.cu).  Constants: valid values V = H + ge (i + j) are
    // >= -(|smin| + 4 go) >= -300 under the limits below; stored = 4 (V - off) + priority;
    // "minus infinity" sits at -30968 and drifts by less than 1200 either way, which stays
    // below the valid floor 4 (-300 - off) = -29400 less one gap open and above -32768.
    Dp16 &h = d.h;
    h = Dp16{};
    int smax = 0, lo = 0, hi = 0;
    for (int i = 0; i < 16; ++i) {
        smax = std::max(smax, p.score[i]);
        lo = std::min(lo, p.score[i] + 2 * d.ge_raw);
        hi = std::max(hi, p.score[i] + 2 * d.ge_raw);
    }
    h.ok = d.go_raw <= 40 && d.ge_raw <= d.go_raw && lo >= -32 && hi <= 31 && getenv("DBGA_NO_WF16") == nullptr;
    if (h.ok) {
        auto dup = [](int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; };
        h.off = 7050;
        h.pair = getenv("DBGA_WF16_NOPAIR") == nullptr;
        h.G = dup(4 * (d.ge_raw - d.go_raw));
        h.tagM = dup(ORDER_PRIO[p.pred_order][0]);
        h.tagX = dup(ORDER_PRIO[p.pred_order][1]);
        h.tagY = dup(ORDER_PRIO[p.pred_order][2]);
        for (int a = 0; a < 4; ++a) {
            uint32_t w = 0;
            for (int c = 0; c < 4; ++c)
                w |= ((uint32_t)(4 * (p.score[a * 4 + c] + 2 * d.ge_raw) + ORDER_PRIO[p.pred_order][0]) & 0xFFu) << (8 * c);
            h.lut[a] = w;
        }
        h.mul[0] = 0xFFFFFFFFu; h.mul[1] = 4; h.mul[2] = 16; h.mul[3] = 256;
        {   // compact traceback fields: [X came from M] = ((u & 3) - pX) / (pM - pX), times 4; same for Y, times 8
            const int pM = ORDER_PRIO[p.pred_order][0], pX = ORDER_PRIO[p.pred_order][1], pY = ORDER_PRIO[p.pred_order][2];
            h.nmul[0] = (uint32_t)(4 / (pM - pX)); h.nmul[1] = (uint32_t)(8 / (pM - pY));
        }
        h.tbk = 4 * ORDER_PRIO[p.pred_order][1] + 16 * ORDER_PRIO[p.pred_order][2];
        h.q0 = 4 * (0 - h.off);
        h.qneg = -30968;
        h.qy0 = 4 * (d.ge_raw - d.go_raw - h.off);
        h.per_step = std::max(0, smax) + 2 * d.ge_raw;
        h.ulimit = 8191 + h.off - 41;
        // row strips in re-anchored frames (dp_wf16s_kernel): a frame must hold per_step * max(strip rows, block
        // columns) above its anchor, and a frame-to-frame step must fit 16 bits
        h.strip = !d.xy && (long long)h.per_step * (std::max(S16_ROWS, S16_CB) + 80) + S16_F0 + 300 <= h.ulimit &&
                  4LL * h.per_step * S16_CB + 4000 <= 32767 && getenv("DBGA_NO_STRIPS") == nullptr ? 8 : 0;
    }
    return DBGA_SUCCESS;
}

constexpr size_t SMEM_LIMIT = 227 * 1024;
constexpr size_t L2_TABLE_BUDGET = 64u << 20;

int do_plan(dbga_ctx *ctx, const int64_t *offsets, int64_t P, const dbga_params *params, bool no_anchors,
            int64_t off_base = 0) {
    ctx->planned = false; ctx->ran = false;
    const auto t_plan0 = std::chrono::steady_clock::now();
    if (!offsets || !params || P < 0) return fail(ctx, DBGA_ERR_ARG, "null argument");
    if (P > 0x7fffff00LL) return fail(ctx, DBGA_ERR_ARG, "too many pairs in one batch");
    if (!no_anchors && params->k < 1) return fail(ctx, DBGA_ERR_ARG, "k must be >= 1");
    if (params->stages != 2 && params->stages != 3) return fail(ctx, DBGA_ERR_ARG, "stages must be 2 or 3");
    ctx->prm = *params;
    ctx->no_anchors = no_anchors;
    int rc = make_dp_params(ctx, *params);
    if (rc != DBGA_SUCCESS) return rc;
    // strip height of the long-segment kernel: with few pairs a long segment is latency-bound (a lone warp per SM
    // sub-partition), so halve the step; with many the taller strip amortises the per-step overhead
    if (ctx->dp.h.strip) {
        // (decided below, once the lengths are known: P * ceil(longest / 512) estimates the strips in flight)
        ctx->dp.h.strip = 8;
        if (const char *e = getenv("DBGA_STRIP_R")) { const int r = atoi(e); if (r == 4 || r == 8) ctx->dp.h.strip = r; }
    }
    DBGA_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int k = no_anchors ? 1 : params->k;
    const bool full = params->want_graph != 0 && !no_anchors;
    ctx->P = P;
    ctx->h_pairs.resize((size_t)P);
    ctx->total_bytes = P > 0 ? offsets[2 * P] - off_base : 0;
    ctx->seq_extent = P > 0 ? offsets[2 * P] - off_base : 0;
    int64_t qb = 0, pb = 0;
    int mx0 = 0, mx1 = 0;
    for (int64_t i = 0; i < P; ++i) {
        const int64_t o0 = offsets[2 * i] - off_base, o1 = offsets[2 * i + 1] - off_base, o2 = offsets[2 * i + 2] - off_base;
        if (o0 < 0 || o1 < o0 || o2 < o1) return fail(ctx, DBGA_ERR_ARG, "offsets must be non-decreasing");
        if (o1 - o0 > 0x7ffffff0LL || o2 - o1 > 0x7ffffff0LL) return fail(ctx, DBGA_ERR_ARG, "sequence longer than 2^31");
        PairDesc &pd = ctx->h_pairs[(size_t)i];
        pd.off0 = o0; pd.off1 = o1; pd.n0 = (int32_t)(o1 - o0); pd.n1 = (int32_t)(o2 - o1);
        pd.q_base = qb; pd.p_base = pb;
        qb += std::max(0, pd.n1 - k + 1);
        pb += std::max(0, pd.n0 - k + 1);
        mx0 = std::max(mx0, pd.n0); mx1 = std::max(mx1, pd.n1);
    }
    ctx->sumN1 = qb; ctx->sumN0 = pb; ctx->max_n0 = mx0; ctx->max_n1 = mx1;

    // ---- stage-1 classing
    ctx->classes.clear(); ctx->h_list.clear(); ctx->h_large.clear(); ctx->waves.clear();
    if (!no_anchors) {
        const int key_bytes = k <= 15 ? 4 : 8;
        std::map<std::pair<int, uint32_t>, std::vector<int32_t>> by_cap;   // (partitions, capacity) -> pairs
        const uint64_t cap_lo = 150, cap_hi = 275;
        std::vector<int32_t> large;
        for (int64_t i = 0; i < P; ++i) {
            const PairDesc &pd = ctx->h_pairs[(size_t)i];
            const int64_t N0 = std::max(0, pd.n0 - k + 1), N1 = std::max(0, pd.n1 - k + 1);
            const uint64_t keys = full ? (uint64_t)(N0 + N1) : (uint64_t)N0;
            // Table capacity (measured on cfg3, tools/cap_sweep.sh): probing cost falls with the load
            // factor until the table costs occupancy, so take the largest table that still lets five
            // CTAs share an SM, within [1.5, 2.75] x the key count; never below the power of two the
            // run sort may need (>= min(N0, N1) runs).  Rounded to a granule so that similar pairs
            // share a launch.
            const int nmax0 = std::max(pd.n0, pd.n1);
            const int words0 = (((nmax0 + 15) / 16 + 3) + 1) & ~1;
            const int64_t fixed = (int64_t)words0 * 8 + ((N1 + 7) / 8 * 8) * 2 + 1024;
            const int64_t fit5 = ((int64_t)SMEM_LIMIT / 5 - fixed) / (key_bytes + 6);
            uint64_t want = (uint64_t)std::max<int64_t>((int64_t)(keys * cap_lo / 100 + 16), std::min<int64_t>((int64_t)(keys * cap_hi / 100), fit5));
            want = std::max<uint64_t>(want, pow2_at_least((uint64_t)std::min(N0, N1) + 1, 64));
            uint32_t gran = 256;
            while ((uint64_t)gran * 16 < want) gran <<= 1;             // <= 16 classes per octave
            const uint32_t cap = (uint32_t)(want / gran * gran >= keys * cap_lo / 100 + 16 && want / gran * gran >= pow2_at_least((uint64_t)std::min(N0, N1) + 1, 64)
                                                ? want / gran * gran : (want + gran - 1) / gran * gran);
            const int nmax = std::max(pd.n0, pd.n1);
            const int words = (((nmax + 15) / 16 + 3) + 1) & ~1;
            const bool fits = k <= 31 && cap <= 32768 && N0 < 65000 && N1 < 65000 &&   // 16-bit positions in shared memory; k >= 32: by-reference keys, global table
                              stage12_small_smem(key_bytes, cap, words, (int)N1, false) <= SMEM_LIMIT;
            if (fits) {
                by_cap[{1, cap}].push_back((int32_t)i);
            } else {
                // medium: 4-byte by-reference slots, the whole table in shared memory if it fits, else the keys
                // split over 2..8 partitions that take turns in it
                int parts = 0;
                uint32_t pcap = 0;
                // (16-bit positions + 1 in the slot, 16-bit slot numbers in memo[]; occurrences are flags, not counters)
                if (k <= 31 && N0 < 65000 && N1 < 65000) {
                    for (int g = 1; g <= 8 && !parts; ++g) {
                        uint64_t w = keys * 170 / 100 / g + 64;       // 1.7 x the expected keys per partition
                        uint32_t gr = 256;
                        while ((uint64_t)gr * 64 < w) gr <<= 1;        // fine granule: every slot counts against the partition count
                        const uint32_t c = (uint32_t)((w + gr - 1) / gr * gr);
                        // (medium pairs keep memo[] in global memory: shared memory holds the packed sequences and the slots)
                        if (c <= 65520 && stage12_small_smem(key_bytes, c, words0, 0, true) <= 224 * 1024) { parts = g; pcap = c; }
                    }
                }
                if (parts) by_cap[{100 + parts, pcap}].push_back((int32_t)i);       // (after every small class: the list's tail goes to the generic stage 2)
                else large.push_back((int32_t)i);
            }
        }
        int stage2_first = -1;
        for (auto &kv : by_cap) {
            const std::vector<int32_t> &members = kv.second;
            SmallClass c;
            c.medium = kv.first.first >= 100; c.parts = c.medium ? kv.first.first - 100 : 1;
            c.cap = kv.first.second; c.begin = (int)ctx->h_list.size(); c.count = (int)members.size();
            int nmax = 0, qmax = 0;
            for (int32_t i : members) {
                const PairDesc &pd = ctx->h_pairs[(size_t)i];
                nmax = std::max(nmax, std::max(pd.n0, pd.n1));
                qmax = std::max(qmax, std::max(0, pd.n1 - k + 1));
                ctx->h_list.push_back(i);
            }
            c.words = (((nmax + 15) / 16 + 3) + 1) & ~1;
            c.maxq = c.medium ? 0 : qmax;
            if (stage12_small_smem(key_bytes, c.cap, c.words, c.maxq, c.medium) > SMEM_LIMIT) {
                // mixed extremes inside one class: fall back to the large path for this class
                ctx->h_list.resize((size_t)c.begin);
                for (int32_t i : members) large.push_back(i);
                continue;
            }
            if (c.medium && stage2_first < 0) stage2_first = c.begin;
            ctx->classes.push_back(c);
        }
        ctx->large_begin = (int)ctx->h_list.size();
        ctx->stage2_begin = stage2_first >= 0 ? stage2_first : ctx->large_begin;
        // waves of large pairs whose tables stay L2 resident together
        size_t wave_tbl = 0, wave_pk = 0;
        LargeWave w{(int)ctx->h_large.size(), 0, 0, 0};
        auto flush = [&]() {
            if (w.count) { w.table_bytes = wave_tbl * 16; ctx->waves.push_back(w); }
            w = LargeWave{(int)ctx->h_large.size(), 0, 0, 0};
            wave_tbl = 0; wave_pk = 0;
        };
        for (int32_t i : large) {
            const PairDesc &pd = ctx->h_pairs[(size_t)i];
            const int64_t N0 = std::max(0, pd.n0 - k + 1), N1 = std::max(0, pd.n1 - k + 1);
            const uint64_t keys = full ? (uint64_t)(N0 + N1) : (uint64_t)N0;
            const uint64_t cap = pow2_at_least(keys + keys / 2 + 1, 1024);
            if (cap > (1ull << 31)) return fail(ctx, DBGA_ERR_ARG, "sequence too long for the k-mer table");
            if (w.count && ((wave_tbl + cap) * 16 > L2_TABLE_BUDGET || w.count >= 32768)) flush();
            LargeDesc ld;
            ld.tbl = (int64_t)wave_tbl; ld.mask = (uint32_t)(cap - 1); ld.pad = 0;
            ld.pk0 = (int64_t)wave_pk; wave_pk += (size_t)(pd.n0 + 15) / 16 + 4;
            ld.pk1 = (int64_t)wave_pk; wave_pk += (size_t)(pd.n1 + 15) / 16 + 4;
            wave_tbl += cap;
            ctx->h_large.push_back(ld);
            ctx->h_list.push_back(i);
            w.count++;
            w.max_n = std::max(w.max_n, std::max(pd.n0, pd.n1));
        }
        flush();
    }

    // ---- device metadata
    RESERVE(d_pairs, sizeof(PairDesc) * std::max<int64_t>(P, 1));
    RESERVE(d_list, sizeof(int32_t) * std::max<size_t>(ctx->h_list.size(), 1));
    RESERVE(d_large, sizeof(LargeDesc) * std::max<size_t>(ctx->h_large.size(), 1));
    {   // metadata goes through one pinned staging buffer so that the uploads are truly asynchronous
        const size_t b0 = sizeof(PairDesc) * (size_t)P, b1 = sizeof(int32_t) * ctx->h_list.size(),
                     b2 = sizeof(LargeDesc) * ctx->h_large.size();
        const size_t o1 = (b0 + 15) & ~(size_t)15, o2 = (o1 + b1 + 15) & ~(size_t)15;
        DBGA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));     // previous use of the staging buffer is over
        HRESERVE(h_meta, o2 + b2 + 16);
        char *m = (char *)ctx->h_meta.p;
        // a lane of the pipelined path uploads through the shared upload stream, so that its own
        // stream never touches the host-to-device copy engine (which is busy with sequence data)
        cudaStream_t ms = ctx->meta_stream ? ctx->meta_stream : ctx->stream;
        if (ctx->meta_stream) DBGA_CUDA_CHECK(cudaStreamWaitEvent(ms, ctx->ev_ready, 0));   // old metadata no longer in use
        if (b0) { memcpy(m, ctx->h_pairs.data(), b0); DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_pairs.p, m, b0, cudaMemcpyHostToDevice, ms)); }
        if (b1) { memcpy(m + o1, ctx->h_list.data(), b1); DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_list.p, m + o1, b1, cudaMemcpyHostToDevice, ms)); }
        if (b2) { memcpy(m + o2, ctx->h_large.data(), b2); DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_large.p, m + o2, b2, cudaMemcpyHostToDevice, ms)); }
        if (ctx->meta_stream) {
            DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev_meta, ms));
            DBGA_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_meta, 0));
        }
    }
    size_t max_tbl = 0, max_pk = 0;
    for (size_t wi = 0; wi < ctx->waves.size(); ++wi) {
        const LargeWave &w = ctx->waves[wi];
        max_tbl = std::max(max_tbl, w.table_bytes);
        const LargeDesc &last = ctx->h_large[(size_t)(w.begin + w.count - 1)];
        const PairDesc &pd = ctx->h_pairs[(size_t)ctx->h_list[(size_t)(ctx->large_begin + w.begin + w.count - 1)]];
        max_pk = std::max(max_pk, (size_t)last.pk1 + (size_t)(pd.n1 + 15) / 16 + 4);
    }
    RESERVE(d_table, max_tbl);
    RESERVE(d_packed, max_pk * 4);

    // ---- workspaces
    RESERVE(d_status, 4 * std::max<int64_t>(P, 1));
    RESERVE(d_status_out, 4 * std::max<int64_t>(P, 1));
    // (+ the medium pairs' 16-bit memo[] scratch behind the int32 entries: sumN1 + two slack entries per pair)
    RESERVE(d_aof, 4 * std::max<int64_t>(ctx->sumN1, 1) + 2 * (ctx->sumN1 + 2 * P + 2) + 16);
    if (full) { RESERVE(d_nd0, std::max<int64_t>(ctx->sumN0, 1)); RESERVE(d_nd1, std::max<int64_t>(ctx->sumN1, 1)); }
    RESERVE(d_pout, sizeof(PairOut) * std::max<int64_t>(P, 1));
    RESERVE(d_ctr, sizeof(Counters));
    const int64_t worst_runs = ctx->sumN1 + 1;
    ctx->runs_cap = std::max<int64_t>(ctx->runs_cap, std::min<int64_t>(worst_runs, ctx->sumN1 / 16 + 4 * P + 1024));
    ctx->slots_cap = std::max<int64_t>(ctx->slots_cap, ctx->runs_cap + P + 16);
    ctx->tb_cap = std::max<int64_t>(ctx->tb_cap, std::min<int64_t>(ctx->scratch_limit, 64LL << 20));
    ctx->n_long_pairs = 0;
    for (int64_t i = 0; i < P; ++i) ctx->n_long_pairs += (ctx->h_pairs[(size_t)i].n1 - k + 1) > S2_CLUSTER_ABOVE;
    // the very long pairs go to the cluster kernel; the CTA kernel is sized for the rest
    int mx1_cta = 0;
    for (int64_t i = 0; i < P; ++i) {
        const int n1 = ctx->h_pairs[(size_t)i].n1;
        if (no_anchors || n1 - k + 1 <= S2_CLUSTER_ABOVE) mx1_cta = std::max(mx1_cta, n1);
    }
    ctx->stage2_threads = mx1_cta <= 4096 ? 128 : (mx1_cta <= 65536 ? 256 : 1024);
    {   // tiny DP classes: a substitution bubble is (k+1) x (k+1), so bucket around multiples of it
        const int b0 = std::min(TINY_MAX, no_anchors ? 8 : k + 1);
        ctx->tiny_bound[0] = b0;
        ctx->tiny_bound[1] = std::min(TINY_MAX, std::max(b0, (3 * b0 + 1) / 2));
        ctx->tiny_bound[2] = std::min(TINY_MAX, std::max(b0, 2 * b0 + 2));
        ctx->tiny_bound[3] = TINY_MAX;
    }
    // boundary rows for multi-pass wavefronts
    ctx->warp_max_cells = P >= MANY_PAIRS ? WARP_CLASS_MAX_CELLS_MANY : WARP_CLASS_MAX_CELLS_FEW;
    ctx->bnd_stride_warp = std::min<int64_t>((int64_t)mx1 + 1, ctx->warp_max_cells / 65 + 2);
    ctx->bnd_stride_cta = (int64_t)mx1 + 2;         // + one entry where the end cell crosses the cluster
    ctx->grid_warp = (int)std::max<int64_t>(ctx->num_sms, std::min<int64_t>(ctx->num_sms * 16, (512LL << 20) / (32 * ctx->bnd_stride_warp)));
    {
        const int64_t per = 2 * ctx->bnd_stride_cta * 16;
        int64_t g = (1LL << 30) / std::max<int64_t>(per, 1);
        ctx->grid_cta = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->num_sms, g));
        // clusters in flight: WF_CLUSTER CTAs of WF_CL_WARPS warps each, two CTAs per SM
        ctx->grid_cluster = (int)std::max<int64_t>(1, std::min<int64_t>(2 * ctx->num_sms / WF_CLUSTER, g));
    }
    const bool any_non_tiny = (mx0 > TINY_MAX) || (mx1 > TINY_MAX);
    if (ctx->dp.h.strip && !getenv("DBGA_STRIP_R")) {
        // strips expected in flight: whole sequences are the segments without anchoring; between anchors long
        // segments are the exception (a few bubbles per pair at most), so count one per pair
        const int64_t est = no_anchors ? P * (((int64_t)mx0 + S16_ROWS - 1) / S16_ROWS) : P;
        ctx->dp.h.strip = est >= (int64_t)ctx->num_sms * 8 ? 8 : 4;
    }
    // (with the strip kernel the long class also takes what the one-pass packed kernel cannot hold)
    const bool any_cta = (int64_t)mx0 * mx1 > ctx->warp_max_cells || (ctx->dp.h.strip && any_non_tiny && mx0 > 256);
    if (!any_non_tiny) ctx->grid_warp = 0;
    if (!any_cta) ctx->grid_cta = ctx->grid_cluster = 0;
    if (params->stages == 3) {
        RESERVE(d_str, 2 * std::max<int64_t>(ctx->total_bytes, 1) + 16);      // + slack: the assembler reads whole words
        if (ctx->grid_warp) RESERVE(d_bnd_w, (size_t)ctx->grid_warp * 2 * ctx->bnd_stride_warp * 16);
        if (ctx->grid_cta && !ctx->dp.h.strip) RESERVE(d_bnd_c, (size_t)ctx->grid_cta * 2 * ctx->bnd_stride_cta * 16);
        RESERVE(d_aln0, std::max<int64_t>(ctx->total_bytes, 1));
        RESERVE(d_aln1, std::max<int64_t>(ctx->total_bytes, 1));
    }
    RESERVE(d_aln_off, 8 * (P + 2)); RESERVE(d_run_off, 8 * (P + 2)); RESERVE(d_tile_off, 8 * (P + 2));
    RESERVE(d_scan, sizeof(AsmScanRec) * (asm_scan_blocks(P) + 1));
    RESERVE(d_tile_pair, sizeof(AsmTile) * (ctx->total_bytes / ASM_TILE_COLS + P + 2));
    RESERVE(d_score, 8 * (P + 1)); RESERVE(d_cells, 8 * (P + 1)); RESERVE(d_nseg, 4 * (P + 1));
    RESERVE(d_nanch, 8 * (P + 1)); RESERVE(d_sop, 16 * (P + 1));
    ctx->ms_plan = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_plan0).count();
    ctx->planned = true;
    return DBGA_SUCCESS;
}

int reserve_pools(dbga_ctx *ctx) {
    RESERVE(d_runs, 12 * ctx->runs_cap);
    RESERVE(d_runs_out, 12 * ctx->runs_cap);
    if (ctx->prm.flags & DBGA_FLAG_COMPACT_ROWS) RESERVE(d_segcols, 4 * (ctx->runs_cap + ctx->P + 1));
    if (ctx->prm.stages == 3) {
        RESERVE(d_slots, sizeof(Slot) * ctx->slots_cap);
        RESERVE(d_item_off, 8 * ctx->slots_cap);
        RESERVE(d_item_src, 16 * ctx->slots_cap);
        RESERVE(d_lists, 4 * ctx->slots_cap * NUM_CLS);
        RESERVE(d_tt, dp_thread_scratch_bytes(TINY_MAX, ctx->num_sms));
        if (ctx->grid_warp || ctx->grid_cta) RESERVE(d_tb, ctx->tb_cap);
    }
    return DBGA_SUCCESS;
}

// the whole kernel pipeline on the context stream; re-run by do_fetch after growing pools
int run_pipeline(dbga_ctx *ctx) {
    cudaStream_t st = ctx->stream;
    const int64_t P = ctx->P;
    const bool seg = ctx->prm.stages == 3;
    const bool full = ctx->prm.want_graph != 0 && !ctx->no_anchors;
    const uint8_t *seqs = ctx->d_seqs_run;
    int rc = reserve_pools(ctx);
    if (rc != DBGA_SUCCESS) return rc;
    ctx->launches = 0;
    DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[1], st));
    DBGA_CUDA_CHECK(launch_clear((Counters *)ctx->d_ctr.p, (int32_t *)ctx->d_status.p, P, (AsmScanRec *)ctx->d_scan.p,
                                 asm_scan_blocks(P), st));
    ctx->launches += 1;
    Pools pools;
    pools.runs = (int32_t *)ctx->d_runs.p; pools.runs_cap = ctx->runs_cap;
    pools.slots = (Slot *)ctx->d_slots.p; pools.slots_cap = ctx->slots_cap;
    pools.tb_cap = (ctx->grid_warp || ctx->grid_cta) ? ctx->tb_cap : 0;
    pools.lists = (int32_t *)ctx->d_lists.p;
    for (int c = 0; c < NUM_TINY; ++c) pools.tiny_bound[c] = ctx->tiny_bound[c];
    pools.list_cap = ctx->slots_cap;
    pools.warp_max_cells = ctx->warp_max_cells;
    pools.strip_ok = ctx->dp.h.strip; pools.strip_per_step = ctx->dp.h.per_step; pools.strip_ulimit = ctx->dp.h.ulimit;
    Counters *ctr = (Counters *)ctx->d_ctr.p;
    PairDesc *pairs = (PairDesc *)ctx->d_pairs.p;
    PairOut *pout = (PairOut *)ctx->d_pout.p;
    // ---- stages 1+2 fused (shared-memory tables) and stage 1 with global tables
    if (!ctx->no_anchors) {
        const bool piecewise = ctx->n_sub > 1 && ctx->waves.empty();
        if (ctx->n_sub > 0 && !piecewise)
            for (int sb = 0; sb < ctx->n_sub; ++sb) DBGA_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_sub[sb], 0));
        for (int sb = 0; sb < (piecewise ? ctx->n_sub : 1); ++sb) {
            if (piecewise) DBGA_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_sub[sb], 0));
            for (const SmallClass &c : ctx->classes) {
                int lo = 0, hi = c.count;
                if (piecewise) {       // class lists are ascending in pair index
                    const int32_t *b = ctx->h_list.data() + c.begin;
                    lo = (int)(std::lower_bound(b, b + c.count, (int32_t)ctx->sub_pair_lo[sb]) - b);
                    hi = (int)(std::lower_bound(b, b + c.count, (int32_t)ctx->sub_pair_lo[sb + 1]) - b);
                }
                if (hi <= lo) continue;
                DBGA_CUDA_CHECK(launch_stage12_small(seqs, pairs, (int32_t *)ctx->d_list.p + c.begin + lo, hi - lo, ctx->prm.k,
                                                     c.cap, c.medium, c.parts, c.words, c.maxq, full, seg, pout, pools, ctr,
                                                     ctx->max_abs_score, ctx->dp.go_raw, ctx->dp.ge_raw, (uint8_t *)ctx->d_nd0.p,
                                                     (uint8_t *)ctx->d_nd1.p, (int32_t *)ctx->d_status.p, (int32_t *)ctx->d_aof.p,
                                                     (uint16_t *)((int32_t *)ctx->d_aof.p + std::max<int64_t>(ctx->sumN1, 1)),
                                                     ctx->num_sms, st));
                ctx->launches += 1;
            }
        }
        for (const LargeWave &w : ctx->waves) {
            DBGA_CUDA_CHECK(launch_stage1_large(seqs, pairs, (LargeDesc *)ctx->d_large.p + w.begin,
                                                (int32_t *)ctx->d_list.p + ctx->large_begin + w.begin, w.count, w.max_n,
                                                ctx->prm.k, full, (uint32_t *)ctx->d_packed.p, ctx->d_table.p, w.table_bytes,
                                                (int32_t *)ctx->d_status.p, (int32_t *)ctx->d_aof.p, (uint8_t *)ctx->d_nd0.p,
                                                (uint8_t *)ctx->d_nd1.p, st, &ctx->launches));
        }
    }
    if (ctx->no_anchors)
        for (int sb = 0; sb < ctx->n_sub; ++sb) DBGA_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_sub[sb], 0));
    DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[2], st));
    // ---- generic stage 2: pairs that went through the global-table path, or every pair
    //      when there is no anchoring at all (dbga_global_align_batch)
    {
        const int64_t n_generic = (int64_t)ctx->h_list.size() - ctx->stage2_begin;
        const int64_t cnt = ctx->no_anchors ? P : n_generic;
        const int32_t *lst = ctx->no_anchors ? nullptr : (int32_t *)ctx->d_list.p + ctx->stage2_begin;
        if (cnt > 0) {
            DBGA_CUDA_CHECK(launch_stage2(pairs, lst, cnt, ctx->no_anchors ? 1 : ctx->prm.k, ctx->no_anchors, seg,
                                          (int32_t *)ctx->d_aof.p, (int32_t *)ctx->d_status.p, pout, pools, ctr,
                                          ctx->stage2_threads, ctx->num_sms, ctx->max_abs_score, ctx->dp.go_raw,
                                          ctx->dp.ge_raw, ctx->no_anchors ? 0 : ctx->n_long_pairs, st));
            ctx->launches += (!ctx->no_anchors && ctx->n_long_pairs > 0) ? 2 : 1;
        }
    }
    DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[3], st));
    const bool strips = seg && ctx->grid_cta && ctx->dp.h.strip;
    if (strips) {
        // long segments as packed row strips, one warp per strip; any grid is correct (device-wide ticket).  Forked
        // onto the side stream: joined again before the assembler.
        static std::atomic<uint32_t> run_nonce{(uint32_t)std::chrono::steady_clock::now().time_since_epoch().count() | 1u};
        uint32_t nonce = run_nonce.fetch_add(1u);
        if (nonce == 0u) nonce = run_nonce.fetch_add(1u);
        cudaStream_t ss = ctx->side_stream;
        const bool seen = ctx->seen_work[CLS_CTA];
        DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, st));
        DBGA_CUDA_CHECK(cudaStreamWaitEvent(ss, ctx->ev_fork, 0));
        DBGA_CUDA_CHECK(launch_dp_strips16(seqs, pairs, pools.slots, pools.lists + (int64_t)CLS_CTA * pools.list_cap,
                                           &ctr->n_cls[CLS_CTA], ctr, seen ? ctx->num_sms * 16 : ctx->num_sms,
                                           ctx->dp, (uint8_t *)ctx->d_tb.p, nonce, ss));
        DBGA_CUDA_CHECK(launch_dp_traceback16(seqs, pairs, pools.slots, pools.lists, pools.list_cap, ctr->n_cls, CLS_CTA, CLS_CTA,
                                              seen ? ctx->num_sms * 4 : 8, ctx->dp, (uint8_t *)ctx->d_str.p,
                                              (const uint8_t *)ctx->d_tb.p, ss));
        DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev_join, ss));
        ctx->launches += 2;
    }
    if (seg) {
        {   // tiny classes: one launch per class that fits the register-resident kernel (bound <= 16),
            // one shared-memory launch for the rest
            TinyBounds tbnd;
            for (int c = 0; c < NUM_TINY; ++c) tbnd.b[c] = ctx->tiny_bound[c];
            int c = 0;
            for (; c < NUM_TINY && tbnd.b[c] <= 16; ++c) {
                if (c > 0 && tbnd.b[c] == tbnd.b[c - 1]) continue;        // empty class
                DBGA_CUDA_CHECK(launch_dp_thread(seqs, pairs, pools.slots, pools.lists, pools.list_cap, ctr->n_cls, c, c + 1, tbnd,
                                                 ctx->dp, (uint8_t *)ctx->d_str.p, (uint32_t *)ctx->d_tt.p, ctx->num_sms, st));
                ctx->launches += 1;
            }
            if (c < NUM_TINY && (c == 0 || tbnd.b[NUM_TINY - 1] > tbnd.b[c - 1])) {
                DBGA_CUDA_CHECK(launch_dp_thread(seqs, pairs, pools.slots, pools.lists, pools.list_cap, ctr->n_cls, c, NUM_TINY,
                                                 tbnd, ctx->dp, (uint8_t *)ctx->d_str.p, (uint32_t *)ctx->d_tt.p, ctx->num_sms, st));
                ctx->launches += 1;
            }
        }
        if (ctx->grid_warp) {
            for (int cls = CLS_WARP2; cls <= CLS_WARP16; ++cls) {
                const int g = ctx->seen_work[cls] ? ctx->grid_warp : std::min(ctx->grid_warp, ctx->num_sms);
                // segments the packed 16-bit form can hold go to dp_wf16_kernel, the rest of the
                // list to the 32-bit kernel; a class whose every member fits launches only the former
                const int rows16 = 32 * wf_rows_per_lane(cls);
                if (ctx->dp.h.ok) {
                    DBGA_CUDA_CHECK(launch_dp_wavefront16(seqs, pairs, pools.slots, pools.lists + (int64_t)cls * pools.list_cap,
                                                          &ctr->n_cls[cls], g, cls, ctx->dp, (uint8_t *)ctx->d_str.p,
                                                          (uint8_t *)ctx->d_tb.p, st, &ctx->launches));
                    if (cls != CLS_WARP16 && wf16_pick(ctx->dp.h, cls, rows16, 1 << 30)) continue;
                }
                DBGA_CUDA_CHECK(launch_dp_wavefront(seqs, pairs, pools.slots, pools.lists + (int64_t)cls * pools.list_cap,
                                                    &ctr->n_cls[cls], g, cls, ctx->dp, (uint8_t *)ctx->d_str.p,
                                                    (uint8_t *)ctx->d_tb.p, (int4 *)ctx->d_bnd_w.p, ctx->bnd_stride_warp, st));
                ctx->launches += 1;
            }
        }
        if (ctx->grid_warp && ctx->dp.h.ok) {
            bool seen = false;
            for (int cls = CLS_WARP2; cls <= CLS_WARP16; ++cls) seen = seen || ctx->seen_work[cls];
            DBGA_CUDA_CHECK(launch_dp_traceback16(seqs, pairs, pools.slots, pools.lists, pools.list_cap, ctr->n_cls, CLS_WARP2, CLS_WARP16,
                                                  seen ? ctx->num_sms * 16 : ctx->num_sms, ctx->dp, (uint8_t *)ctx->d_str.p,
                                                  (const uint8_t *)ctx->d_tb.p, st));
            ctx->launches += 1;
        }
        if (ctx->grid_cta && !ctx->dp.h.strip) {
            const bool seen = ctx->seen_work[CLS_CTA];
            DBGA_CUDA_CHECK(launch_dp_long(seqs, pairs, pools.slots, pools.lists + (int64_t)CLS_CTA * pools.list_cap,
                                           &ctr->n_cls[CLS_CTA], seen ? ctx->grid_cluster : std::min(ctx->grid_cluster, 2),
                                           seen ? ctx->grid_cta : std::min(ctx->grid_cta, 8), 2u * (unsigned)ctx->grid_cluster,
                                           ctx->dp, (uint8_t *)ctx->d_str.p, (uint8_t *)ctx->d_tb.p,
                                           (int4 *)ctx->d_bnd_c.p, ctx->bnd_stride_cta, st));
            ctx->launches += 2;
        }
    }
    if (strips) DBGA_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[4], st));
    dbga_ctx::MetaPtrs mp = ctx->ext;
    if (!ctx->use_ext) {
        mp.status = (int32_t *)ctx->d_status_out.p; mp.score = (int64_t *)ctx->d_score.p; mp.cells = (int64_t *)ctx->d_cells.p;
        mp.nseg = (int32_t *)ctx->d_nseg.p; mp.nanch = (int64_t *)ctx->d_nanch.p;
        mp.aln_off = (int64_t *)ctx->d_aln_off.p; mp.run_off = (int64_t *)ctx->d_run_off.p; mp.sop = (int32_t *)ctx->d_sop.p;
    }
    const AsmOut ao{mp.status, mp.score, mp.cells, mp.nseg, mp.nanch, mp.aln_off, mp.run_off, seg ? mp.sop : nullptr};
    DBGA_CUDA_CHECK(launch_assemble(seqs, pairs, P, pout, (int32_t *)ctx->d_runs.p, (Slot *)ctx->d_slots.p,
                                    (uint8_t *)ctx->d_str.p, (int32_t *)ctx->d_item_off.p, (long long *)ctx->d_item_src.p,
                                    (AsmScanRec *)ctx->d_scan.p, (AsmTile *)ctx->d_tile_pair.p, ao,
                                    (uint8_t *)ctx->d_aln0.p, (uint8_t *)ctx->d_aln1.p,
                                    (int32_t *)ctx->d_runs_out.p, (int32_t *)ctx->d_segcols.p, ctr, ctx->seq_extent, ctx->prm.flags,
                                    ctx->num_sms, st, &ctx->launches));
    DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[5], st));
    return DBGA_SUCCESS;
}

int do_run(dbga_ctx *ctx, const uint8_t *d_seqs, bool keep_sub = false) {
    if (!ctx->planned) return fail(ctx, DBGA_ERR_ARG, "dbga_plan_run without a plan");
    if (!keep_sub) ctx->n_sub = 0;
    DBGA_CUDA_CHECK(cudaSetDevice(ctx->device));
    ctx->d_seqs_run = d_seqs;
    int rc = run_pipeline(ctx);
    if (rc != DBGA_SUCCESS) return rc;
    ctx->ran = true;
    return DBGA_SUCCESS;
}

struct HostDst {
    int32_t *status; int64_t *score; int64_t *cells; int32_t *nseg; int64_t *nanch;
    int64_t *aln_off; int64_t *run_off; uint8_t *aln0; uint8_t *aln1; int32_t *runs; int32_t *sop; int32_t *segcols;
};

// counters -> host; on pool overflow grow to the measured need and re-run the pipeline
int fetch_resolve(dbga_ctx *ctx, bool counters_already_copied) {
    cudaStream_t st = ctx->stream;
    Counters *hc = (Counters *)ctx->h_ctr.p;
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (!(attempt == 0 && counters_already_copied)) {
            DBGA_CUDA_CHECK(cudaMemcpyAsync(hc, ctx->d_ctr.p, sizeof(Counters), cudaMemcpyDeviceToHost, st));
            DBGA_CUDA_CHECK(cudaStreamSynchronize(st));
        }
        if (hc->strip_timeout) return fail(ctx, DBGA_ERR_CUDA, "long-segment strip kernel: a boundary row never arrived (internal error)");
        bool grow = false;
        if ((int64_t)hc->runs_used > ctx->runs_cap) { ctx->runs_cap = (int64_t)hc->runs_used + 1024; grow = true; }
        if ((int64_t)hc->slots_used > ctx->slots_cap) { ctx->slots_cap = (int64_t)hc->slots_used + 1024; grow = true; }
        ctx->slots_cap = std::max(ctx->slots_cap, ctx->runs_cap + ctx->P + 16);
        if ((int64_t)hc->tb_used > ctx->tb_cap && ctx->tb_cap < ctx->scratch_limit) {
            ctx->tb_cap = std::min<int64_t>(ctx->scratch_limit, (int64_t)hc->tb_used + (1 << 20));
            grow = true;
        }
        if (!grow) break;
        int rc = run_pipeline(ctx);
        if (rc != DBGA_SUCCESS) return rc;
    }
    // the wavefront kernels are persistent, so any grid is correct; keep it small while the
    // workload has no segment for them (most closely related pairs)
    for (int c = 0; c < NUM_CLS; ++c) ctx->seen_work[c] = hc->n_cls[c] > 0;
    return DBGA_SUCCESS;
}

// asynchronous device -> host copies of one batch's results (offsets are batch-local)
int fetch_enqueue(dbga_ctx *ctx, const HostDst &d, int64_t total_aln, int64_t total_runs, cudaStream_t st) {
    const int64_t P = ctx->P;
    if (!P) return DBGA_SUCCESS;
    if (!ctx->use_ext) {          // pipeline lanes: the parent fetches these columns once per batch
    DBGA_CUDA_CHECK(cudaMemcpyAsync(d.status, ctx->d_status_out.p, 4 * P, cudaMemcpyDeviceToHost, st));
    DBGA_CUDA_CHECK(cudaMemcpyAsync(d.score, ctx->d_score.p, 8 * P, cudaMemcpyDeviceToHost, st));
    DBGA_CUDA_CHECK(cudaMemcpyAsync(d.cells, ctx->d_cells.p, 8 * P, cudaMemcpyDeviceToHost, st));
    DBGA_CUDA_CHECK(cudaMemcpyAsync(d.nseg, ctx->d_nseg.p, 4 * P, cudaMemcpyDeviceToHost, st));
    DBGA_CUDA_CHECK(cudaMemcpyAsync(d.nanch, ctx->d_nanch.p, 8 * P, cudaMemcpyDeviceToHost, st));
    DBGA_CUDA_CHECK(cudaMemcpyAsync(d.aln_off, ctx->d_aln_off.p, 8 * (P + 1), cudaMemcpyDeviceToHost, st));
    DBGA_CUDA_CHECK(cudaMemcpyAsync(d.run_off, ctx->d_run_off.p, 8 * (P + 1), cudaMemcpyDeviceToHost, st));
    if (ctx->prm.stages == 3 && (ctx->prm.flags & DBGA_FLAG_SOP))
        DBGA_CUDA_CHECK(cudaMemcpyAsync(d.sop, ctx->d_sop.p, 16 * P, cudaMemcpyDeviceToHost, st));
    }
    if (total_aln) {
        DBGA_CUDA_CHECK(cudaMemcpyAsync(d.aln0, ctx->d_aln0.p, total_aln, cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(cudaMemcpyAsync(d.aln1, ctx->d_aln1.p, total_aln, cudaMemcpyDeviceToHost, st));
    }
    if (total_runs) DBGA_CUDA_CHECK(cudaMemcpyAsync(d.runs, ctx->d_runs_out.p, 12 * total_runs, cudaMemcpyDeviceToHost, st));
    if ((ctx->prm.flags & DBGA_FLAG_COMPACT_ROWS) && ctx->prm.stages == 3)
        DBGA_CUDA_CHECK(cudaMemcpyAsync(d.segcols, ctx->d_segcols.p, 4 * (total_runs + P), cudaMemcpyDeviceToHost, st));
    return DBGA_SUCCESS;
}

int do_fetch(dbga_ctx *ctx, dbga_result *out) {
    if (!ctx->ran) return fail(ctx, DBGA_ERR_ARG, "dbga_plan_fetch before dbga_plan_run");
    DBGA_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t P = ctx->P;
    HRESERVE(h_ctr, sizeof(Counters));
    Counters *hc = (Counters *)ctx->h_ctr.p;
    int rc = fetch_resolve(ctx, false);
    if (rc != DBGA_SUCCESS) return rc;
    DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[6], st));
    const bool seg = ctx->prm.stages == 3;
    const bool full = ctx->prm.want_graph != 0 && !ctx->no_anchors;
    HRESERVE(h_status, 4 * (P + 1)); HRESERVE(h_score, 8 * (P + 1)); HRESERVE(h_cells, 8 * (P + 1));
    HRESERVE(h_nseg, 4 * (P + 1)); HRESERVE(h_nanch, 8 * (P + 1)); HRESERVE(h_aln_off, 8 * (P + 2));
    HRESERVE(h_run_off, 8 * (P + 2));
    const int64_t total_aln = seg ? (int64_t)hc->total_aln : 0, total_runs = (int64_t)hc->total_runs;
    HRESERVE(h_aln0, std::max<int64_t>(total_aln, 1)); HRESERVE(h_aln1, std::max<int64_t>(total_aln, 1));
    HRESERVE(h_runs, 12 * std::max<int64_t>(total_runs, 1));
    HRESERVE(h_sop, 16 * (P + 1));
    const bool compact = seg && (ctx->prm.flags & DBGA_FLAG_COMPACT_ROWS);
    if (compact) HRESERVE(h_segcols, 4 * (total_runs + P + 1));
    HostDst d{(int32_t *)ctx->h_status.p, (int64_t *)ctx->h_score.p, (int64_t *)ctx->h_cells.p, (int32_t *)ctx->h_nseg.p,
              (int64_t *)ctx->h_nanch.p, (int64_t *)ctx->h_aln_off.p, (int64_t *)ctx->h_run_off.p,
              (uint8_t *)ctx->h_aln0.p, (uint8_t *)ctx->h_aln1.p, (int32_t *)ctx->h_runs.p, (int32_t *)ctx->h_sop.p,
              (int32_t *)ctx->h_segcols.p};
    if (P) {
        rc = fetch_enqueue(ctx, d, total_aln, total_runs, st);
        if (rc != DBGA_SUCCESS) return rc;
    } else {
        ((int64_t *)ctx->h_aln_off.p)[0] = 0; ((int64_t *)ctx->h_run_off.p)[0] = 0;
    }
    if (full) {
        HRESERVE(h_nd0, std::max<int64_t>(ctx->sumN0, 1)); HRESERVE(h_nd1, std::max<int64_t>(ctx->sumN1, 1));
        HRESERVE(h_poff, 8 * (P + 1)); HRESERVE(h_qoff, 8 * (P + 1));
        if (ctx->sumN0) DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_nd0.p, ctx->d_nd0.p, ctx->sumN0, cudaMemcpyDeviceToHost, st));
        if (ctx->sumN1) DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_nd1.p, ctx->d_nd1.p, ctx->sumN1, cudaMemcpyDeviceToHost, st));
        int64_t *po = (int64_t *)ctx->h_poff.p, *qo = (int64_t *)ctx->h_qoff.p;
        for (int64_t i = 0; i < P; ++i) { po[i] = ctx->h_pairs[(size_t)i].p_base; qo[i] = ctx->h_pairs[(size_t)i].q_base; }
        po[P] = ctx->sumN0; qo[P] = ctx->sumN1;
    }
    DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[7], st));
    DBGA_CUDA_CHECK(cudaStreamSynchronize(st));
    memset(out, 0, sizeof(*out));
    out->num_pairs = P;
    out->status = d.status; out->score = d.score; out->dp_cells = d.cells; out->dp_segments = d.nseg;
    out->n_anchors = d.nanch; out->run_off = d.run_off; out->runs = d.runs; out->aln_off = d.aln_off;
    out->aln0 = d.aln0; out->aln1 = d.aln1;
    out->sop = (seg && (ctx->prm.flags & DBGA_FLAG_SOP)) ? d.sop : nullptr; out->seg_cols = compact ? d.segcols : nullptr;
    out->ms_plan = ctx->ms_plan;
    if (!seg && P) { for (int64_t i = 0; i <= P; ++i) d.aln_off[i] = 0; }
    if (full) {
        out->p_off = (int64_t *)ctx->h_poff.p; out->nondup0 = (uint8_t *)ctx->h_nd0.p;
        out->q_off = (int64_t *)ctx->h_qoff.p; out->nondup1 = (uint8_t *)ctx->h_nd1.p;
    }
    cudaEventElapsedTime(&out->ms_stage1, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&out->ms_stage2, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&out->ms_stage3, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&out->ms_assemble, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&out->ms_d2h, ctx->ev[6], ctx->ev[7]);
    cudaEventElapsedTime(&out->ms_total, ctx->ev[1], ctx->ev[7]);
    out->dp_cells_total = (int64_t)hc->cells;
    out->dp_segments_total = (int64_t)hc->nseg;
    out->kernel_launches = ctx->launches;
    return DBGA_SUCCESS;
}

int create_ctx(int device, dbga_ctx **out);

// Pipelined whole-batch call: the batch is cut into chunks that flow through NL child
// contexts (lanes).  Each lane's stream carries H2D -> kernels -> D2H of its chunk in order,
// so at any time one lane uploads, one computes and one downloads.  The host never waits
// for anything but the small counters of the previous chunk (needed to size its D2H).
int align_pipelined(dbga_ctx *ctx, const uint8_t *seqs, const uint32_t *packed, int64_t seq0, const int64_t *offsets, int64_t P,
                    const dbga_params *params, dbga_result *out, bool no_anchors, int64_t chunk_bytes) {
    const auto t_begin = std::chrono::steady_clock::now();
    const int64_t total_bytes = offsets[2 * P] - offsets[0];
    const bool seg = params->stages == 3;
    constexpr int NL = dbga_ctx::NL;
    int64_t sub_bytes = 16LL << 20;    // upload piece size
    if (const char *e = getenv("DBGA_SUB_MB")) sub_bytes = (int64_t)std::max(1, atoi(e)) << 20;
    if (!ctx->copy_stream && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(ctx, DBGA_ERR_CUDA, "could not create the upload stream");
    if (!ctx->dl_stream && cudaStreamCreateWithFlags(&ctx->dl_stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(ctx, DBGA_ERR_CUDA, "could not create the download stream");
    const bool spin = ctx->spin = decide_spin(ctx->multi_peers);
    for (int i = 0; i < NL; ++i) {
        if (!ctx->lanes[i]) {
            int rc = create_ctx(ctx->device, &ctx->lanes[i]);
            if (rc != DBGA_SUCCESS) return fail(ctx, rc, "could not create a pipeline lane");
            dbga_ctx *l = ctx->lanes[i];
            bool okc = cudaEventCreateWithFlags(&l->ev_ready, cudaEventDisableTiming | (spin ? 0 : cudaEventBlockingSync)) == cudaSuccess &&
                       cudaEventCreateWithFlags(&l->ev_meta, cudaEventDisableTiming) == cudaSuccess &&
                       cudaEventCreateWithFlags(&l->ev_dl, cudaEventDisableTiming) == cudaSuccess;
            l->meta_stream = ctx->copy_stream;
            for (int sb = 0; okc && sb < dbga_ctx::MAX_SUB; ++sb)
                okc = cudaEventCreateWithFlags(&l->ev_sub[sb], cudaEventDisableTiming) == cudaSuccess;
            for (auto &e : l->dbg_ev) okc = okc && cudaEventCreate(&e) == cudaSuccess;
            if (!okc) return fail(ctx, DBGA_ERR_CUDA, "could not create lane streams / events");
        }
        ctx->lanes[i]->scratch_limit = ctx->scratch_limit;
    }
    // Chunk boundaries.  The download engine is the busiest resource of the pipeline (results
    // are as large as the input and share the link with the uploads), so the first chunks are
    // small -- results start flowing back almost immediately -- and grow geometrically up to
    // the requested chunk size.
    std::vector<int64_t> lo;
    lo.push_back(0);
    {
        // sizes in bytes: 8, 16, 32, ... up to chunk_bytes, and back down to 8 MB at the end -- the last
        // chunk's kernels and download are the tail nothing can overlap with
        int64_t first_mb = packed ? 32 : 8;
        if (const char *e = getenv("DBGA_FIRST_MB")) first_mb = std::max(1, atoi(e));
        chunk_bytes = std::min<int64_t>(chunk_bytes, std::max<int64_t>(4LL << 20, total_bytes / 4));      // small batches still pipeline
        const int64_t small = std::min<int64_t>(std::min<int64_t>(first_mb << 20, chunk_bytes), std::max<int64_t>(2LL << 20, total_bytes / 8));
        std::vector<int64_t> head, tail;
        int64_t used = 0;
        for (int64_t w = small; w < chunk_bytes && used + 2 * w < total_bytes / 2; w *= 2) { head.push_back(w); tail.push_back(w); used += 2 * w; }
        const int64_t rest = total_bytes - used;
        const int64_t nmid = std::max<int64_t>(1, (rest + chunk_bytes - 1) / chunk_bytes);
        std::vector<int64_t> sizes = head;
        for (int64_t i = 0; i < nmid; ++i) sizes.push_back(rest / nmid + 1);
        for (size_t i = tail.size(); i-- > 0;) sizes.push_back(tail[i]);
        int64_t cut = 0;
        size_t si = 0;
        cut = sizes[0];
        for (int64_t i = 1; i < P && si + 1 < sizes.size(); ++i) {
            if (offsets[2 * i] - offsets[0] >= cut) {
                lo.push_back(i);
                ++si;
                cut = (offsets[2 * i] - offsets[0]) + sizes[si];
            }
        }
        // a tiny tail chunk is merged into its predecessor
        if (lo.size() > 1 && total_bytes - (offsets[2 * lo.back()] - offsets[0]) < (2LL << 20)) lo.pop_back();
        lo.push_back(P);
    }
    const int nchunks = (int)lo.size() - 1;
    HRESERVE(h_status, 4 * (P + 1)); HRESERVE(h_score, 8 * (P + 1)); HRESERVE(h_cells, 8 * (P + 1));
    HRESERVE(h_nseg, 4 * (P + 1)); HRESERVE(h_nanch, 8 * (P + 1));
    HRESERVE(h_aln_off, 8 * (P + 2 + nchunks)); HRESERVE(h_run_off, 8 * (P + 2 + nchunks));
    if (seg) { HRESERVE(h_aln0, total_bytes + 16); HRESERVE(h_aln1, total_bytes + 16); }
    const bool compact = seg && (params->flags & DBGA_FLAG_COMPACT_ROWS);
    const bool no_runs = (params->flags & DBGA_FLAG_NO_RUNS) && !compact;
    if (!no_runs) HRESERVE(h_runs, std::max<int64_t>((int64_t)ctx->h_runs.cap, total_bytes / 4 + 4096));
    if (compact) HRESERVE(h_segcols, std::max<int64_t>((int64_t)ctx->h_segcols.cap, total_bytes / 12 + 4 * P + 4096));
    HRESERVE(h_sop, 16 * (P + 1));
    RESERVE(d_status_out, 4 * (P + 1)); RESERVE(d_score, 8 * (P + 1)); RESERVE(d_cells, 8 * (P + 1));
    RESERVE(d_nseg, 4 * (P + 1)); RESERVE(d_nanch, 8 * (P + 1)); RESERVE(d_sop, 16 * (P + 1));
    RESERVE(d_aln_off, 8 * (P + 2 + nchunks)); RESERVE(d_run_off, 8 * (P + 2 + nchunks));
    int64_t *aln_off = (int64_t *)ctx->h_aln_off.p, *run_off = (int64_t *)ctx->h_run_off.p;
    std::vector<int64_t> aln_base((size_t)nchunks + 1, 0), run_base((size_t)nchunks + 1, 0);
    float ms[4] = {0, 0, 0, 0};
    std::vector<float> plan_ms((size_t)nchunks, 0.f);
    int64_t cells = 0, nsegs = 0, launches = 0;
    auto fail_lane = [&](dbga_ctx *l, int rc) { snprintf(ctx->err, sizeof(ctx->err), "%s", l->err); return rc; };

    const bool trace = getenv("DBGA_TRACE") != nullptr;
    // DBGA_TRACE: six events per chunk (upload begin / end, kernels begin / end, download begin / end),
    // read after the batch so that tracing does not serialise the pipeline
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize((size_t)nchunks * 6);
        for (auto &e : tev) cudaEventCreate(&e);
        cudaEventRecord(ctx->ev[0], ctx->stream); cudaEventSynchronize(ctx->ev[0]);
    }
    auto now_ms = [&]() { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };

    // ---- producer: plan chunk c, queue its upload pieces and its kernels (lane c % NL)
    auto issue = [&](int c) -> int {
        dbga_ctx *l = ctx->lanes[c % NL];
        const int64_t a = lo[(size_t)c], b = lo[(size_t)c + 1];
        const int64_t base = offsets[2 * a];
        if (cudaSetDevice(ctx->device) != cudaSuccess) return DBGA_ERR_CUDA;
        int rc = do_plan(l, offsets + 2 * a, b - a, params, no_anchors, base);
        if (rc != DBGA_SUCCESS) return rc;
        plan_ms[(size_t)c] = l->ms_plan;
        // chunk c's offsets land at [a + c, a + c + n + 1) (one extra slot per chunk), fixed up at the end
        l->use_ext = true;
        l->ext.status = (int32_t *)ctx->d_status_out.p + a; l->ext.score = (int64_t *)ctx->d_score.p + a;
        l->ext.cells = (int64_t *)ctx->d_cells.p + a; l->ext.nseg = (int32_t *)ctx->d_nseg.p + a;
        l->ext.nanch = (int64_t *)ctx->d_nanch.p + a; l->ext.sop = (int32_t *)ctx->d_sop.p + 4 * a;
        l->ext.aln_off = (int64_t *)ctx->d_aln_off.p + a + c; l->ext.run_off = (int64_t *)ctx->d_run_off.p + a + c;
        if ((size_t)std::max<int64_t>(l->total_bytes, 16) + 64 > l->d_seqs.cap) cudaStreamSynchronize(ctx->copy_stream);
        rc = dev_reserve(l, l->d_seqs, (size_t)std::max<int64_t>(l->total_bytes, 16) + 64);
        if (rc != DBGA_SUCCESS) return rc;
        // upload in pieces on the shared upload stream (after the lane's previous kernels are done with
        // d_seqs); stage 1+2 of each piece starts as soon as the piece has landed
        if (cudaStreamWaitEvent(ctx->copy_stream, l->ev_ready, 0) != cudaSuccess) return DBGA_ERR_CUDA;
        if (trace) cudaEventRecord(tev[(size_t)c * 6 + 0], ctx->copy_stream);
        const int64_t nb = l->total_bytes, np = b - a;
        if (packed) {
            // 2-bit input: the chunk's words go up in one piece and are expanded on the device
            const int64_t w_lo = (offsets[2 * a] >> 4) + seq0 + 2 * a, w_hi = (offsets[2 * b] >> 4) + seq0 + 2 * b;
            if ((size_t)(4 * (w_hi - w_lo) + 64) > l->d_pk_in.cap) cudaStreamSynchronize(ctx->copy_stream);
            rc = dev_reserve(l, l->d_pk_in, (size_t)(4 * (w_hi - w_lo) + 64));
            if (rc != DBGA_SUCCESS) return rc;
            if (w_hi > w_lo && cudaMemcpyAsync(l->d_pk_in.p, packed + w_lo, (size_t)(4 * (w_hi - w_lo)), cudaMemcpyHostToDevice,
                                               ctx->copy_stream) != cudaSuccess)
                return DBGA_ERR_CUDA;
            if (cudaEventRecord(l->ev_sub[0], ctx->copy_stream) != cudaSuccess) return DBGA_ERR_CUDA;
            if (trace) { cudaEventRecord(tev[(size_t)c * 6 + 1], ctx->copy_stream); cudaEventRecord(tev[(size_t)c * 6 + 2], l->stream); }
            if (l->dl_pending && cudaStreamWaitEvent(l->stream, l->ev_dl, 0) != cudaSuccess) return DBGA_ERR_CUDA;
            if (cudaStreamWaitEvent(l->stream, l->ev_sub[0], 0) != cudaSuccess) return DBGA_ERR_CUDA;
            if (launch_unpack((const uint32_t *)l->d_pk_in.p, (const PairDesc *)l->d_pairs.p, np, base, seq0 + 2 * a, w_lo,
                              (uint8_t *)l->d_seqs.p, l->num_sms, l->stream) != cudaSuccess)
                return DBGA_ERR_CUDA;
            rc = do_run(l, (const uint8_t *)l->d_seqs.p, false);
            if (rc != DBGA_SUCCESS) return rc;
            l->launches += 1;
            if (trace) cudaEventRecord(tev[(size_t)c * 6 + 3], l->stream);
            rc = host_reserve(l, l->h_ctr, sizeof(Counters));
            if (rc != DBGA_SUCCESS) return rc;
            if (launch_publish(l->d_ctr.p, l->h_ctr.p, (int)sizeof(Counters), l->stream) != cudaSuccess ||
                cudaEventRecord(l->ev_ready, l->stream) != cudaSuccess)
                return DBGA_ERR_CUDA;
            return DBGA_SUCCESS;
        }
        int ns = (int)std::max<int64_t>(1, std::min<int64_t>(dbga_ctx::MAX_SUB, (nb + sub_bytes - 1) / sub_bytes));
        if (np < ns) ns = (int)std::max<int64_t>(1, np);
        l->n_sub = ns;
        l->sub_pair_lo[0] = 0;
        int64_t prev_off = 0;
        for (int sb = 0; sb < ns; ++sb) {
            int64_t pe = np;                                   // first pair of the next piece (chunk-relative)
            if (sb + 1 < ns) {
                const int64_t target = base + nb * (sb + 1) / ns;
                const int64_t *o = offsets + 2 * a;
                int64_t lo2 = l->sub_pair_lo[sb], hi2 = np;
                while (lo2 < hi2) { const int64_t mid = (lo2 + hi2) / 2; if (o[2 * mid] < target) lo2 = mid + 1; else hi2 = mid; }
                pe = lo2;
            }
            l->sub_pair_lo[sb + 1] = pe;
            const int64_t end_off = (pe == np ? offsets[2 * b] : offsets[2 * (a + pe)]) - base;
            if (end_off > prev_off &&
                cudaMemcpyAsync((uint8_t *)l->d_seqs.p + prev_off, seqs + base + prev_off, (size_t)(end_off - prev_off),
                                cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess)
                return DBGA_ERR_CUDA;
            if (cudaEventRecord(l->ev_sub[sb], ctx->copy_stream) != cudaSuccess) return DBGA_ERR_CUDA;
            prev_off = end_off;
        }
        if (trace) { cudaEventRecord(tev[(size_t)c * 6 + 1], ctx->copy_stream); cudaEventRecord(tev[(size_t)c * 6 + 2], l->stream); }
        // the lane's previous download must have left its result buffers (committed before this chunk was issued)
        if (l->dl_pending && cudaStreamWaitEvent(l->stream, l->ev_dl, 0) != cudaSuccess) return DBGA_ERR_CUDA;
        rc = do_run(l, (const uint8_t *)l->d_seqs.p, true);
        if (rc != DBGA_SUCCESS) return rc;
        if (trace) cudaEventRecord(tev[(size_t)c * 6 + 3], l->stream);
        rc = host_reserve(l, l->h_ctr, sizeof(Counters));
        if (rc != DBGA_SUCCESS) return rc;
        if (launch_publish(l->d_ctr.p, l->h_ctr.p, (int)sizeof(Counters), l->stream) != cudaSuccess ||
            cudaEventRecord(l->ev_ready, l->stream) != cudaSuccess)
            return DBGA_ERR_CUDA;
        return DBGA_SUCCESS;
    };

    // ---- consumer: once chunk c's counters are on the host, queue its (exactly sized) download
    auto commit = [&](int c) -> int {
        dbga_ctx *l = ctx->lanes[c % NL];
        const int64_t a = lo[(size_t)c];
        if (cudaEventSynchronize(l->ev_ready) != cudaSuccess) return fail(ctx, DBGA_ERR_CUDA, "cudaEventSynchronize failed");
        int rc = fetch_resolve(l, true);
        if (rc != DBGA_SUCCESS) return fail_lane(l, rc);
        const Counters *hc = (const Counters *)l->h_ctr.p;
        int64_t ta = seg ? (int64_t)hc->total_aln : 0, tr = (int64_t)hc->total_runs;
        aln_base[(size_t)c + 1] = aln_base[(size_t)c] + ta;
        run_base[(size_t)c + 1] = run_base[(size_t)c] + tr;
        auto grow_keep = [&](HostBuf &hb, size_t need, size_t used) -> int {       // rare: grow a pinned result buffer
            if (need <= hb.cap) return DBGA_SUCCESS;
            for (int i = 0; i < NL; ++i) cudaStreamSynchronize(ctx->lanes[i]->stream);
            cudaStreamSynchronize(ctx->dl_stream);
            HostBuf old = hb;
            hb = HostBuf();
            const int r = host_reserve(ctx, hb, need * 2);
            if (r != DBGA_SUCCESS) return r;
            if (used) memcpy(hb.p, old.p, used);
            cudaFreeHost(old.p);
            return DBGA_SUCCESS;
        };
        rc = grow_keep(ctx->h_runs, (size_t)(12 * run_base[(size_t)c + 1]), (size_t)(12 * run_base[(size_t)c]));
        if (rc != DBGA_SUCCESS) return rc;
        if (compact) {
            rc = grow_keep(ctx->h_segcols, (size_t)(4 * (run_base[(size_t)c + 1] + lo[(size_t)c + 1])), (size_t)(4 * (run_base[(size_t)c] + a)));
            if (rc != DBGA_SUCCESS) return rc;
        }
        float t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        cudaEventElapsedTime(&t1, l->ev[1], l->ev[2]); cudaEventElapsedTime(&t2, l->ev[2], l->ev[3]);
        cudaEventElapsedTime(&t3, l->ev[3], l->ev[4]); cudaEventElapsedTime(&t4, l->ev[4], l->ev[5]);
        ms[0] += t1; ms[1] += t2; ms[2] += t3; ms[3] += t4;
        cells += (int64_t)hc->cells; nsegs += (int64_t)hc->nseg; launches += l->launches;
        // offsets of chunk c land at [a + c, a + c + n + 1) (one extra slot per chunk), fixed up below
        HostDst d{(int32_t *)ctx->h_status.p + a, (int64_t *)ctx->h_score.p + a, (int64_t *)ctx->h_cells.p + a,
                  (int32_t *)ctx->h_nseg.p + a, (int64_t *)ctx->h_nanch.p + a, aln_off + a + c, run_off + a + c,
                  (uint8_t *)ctx->h_aln0.p + aln_base[(size_t)c], (uint8_t *)ctx->h_aln1.p + aln_base[(size_t)c],
                  (int32_t *)ctx->h_runs.p + 3 * run_base[(size_t)c], (int32_t *)ctx->h_sop.p + 4 * a,
                  compact ? (int32_t *)ctx->h_segcols.p + run_base[(size_t)c] + a : nullptr};
        // ev_ready has fired (synchronised above), so the download needs no device-side wait
        if (trace) cudaEventRecord(tev[(size_t)c * 6 + 4], ctx->dl_stream);
        rc = fetch_enqueue(l, d, ta, tr, ctx->dl_stream);
        if (rc != DBGA_SUCCESS) return fail_lane(l, rc);
        if (cudaEventRecord(l->ev_dl, ctx->dl_stream) != cudaSuccess) return fail(ctx, DBGA_ERR_CUDA, "cudaEventRecord failed");
        l->dl_pending = true;
        if (trace) {
            cudaEventRecord(tev[(size_t)c * 6 + 5], ctx->dl_stream);
            fprintf(stderr, "[dbga] chunk %d committed at host %.2f ms (s12 %.2f s3 %.2f asm %.2f)\n", c, now_ms(), t1, t3, t4);
        }
        return DBGA_SUCCESS;
    };

    // The producer runs on its own host thread so that planning the next chunk never delays the
    // download of a finished one.  It may run at most NL - 1 chunks ahead of the consumer (a lane
    // is reused only after its previous chunk has been committed).
    // Two producer threads take alternate chunks (lanes are independent contexts; the shared upload
    // stream takes copies from both): planning + ~20 launches cost ~0.4 ms of host time per chunk,
    // which one thread cannot hide behind a 0.15 - 0.7 ms upload.
    int n_prod = 2;
    if (const char *e = getenv("DBGA_PRODUCERS")) n_prod = std::max(1, std::min(2, atoi(e)));
    if (!spin) n_prod = 1;                       // few host cores per GPU: do not add threads
    std::vector<std::atomic<int>> issued((size_t)nchunks);
    for (auto &f : issued) f.store(0);
    std::atomic<int> committed{0}, prod_rc{DBGA_SUCCESS};
    std::atomic<bool> stop{false};
    std::mutex mu;
    std::condition_variable cv;
    // spin: poll the atomics (fastest hand-over); otherwise sleep on the condition variable
    auto wait_until = [&](auto pred) {
        if (spin) { while (!pred()) std::this_thread::yield(); return; }
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, pred);
    };
    auto publish = [&](std::atomic<int> &var, int v) {
        if (spin) { var.store(v, std::memory_order_release); return; }
        { std::lock_guard<std::mutex> lk(mu); var.store(v, std::memory_order_release); }
        cv.notify_all();
    };
    auto produce = [&](int first) {
        for (int c = first; c < nchunks && !stop.load(); c += n_prod) {
            wait_until([&]() { return c - committed.load(std::memory_order_acquire) < NL || stop.load(); });
            if (stop.load()) break;
            const int rc = issue(c);
            if (rc != DBGA_SUCCESS) {
                { std::lock_guard<std::mutex> lk(mu); snprintf(ctx->err, sizeof(ctx->err), "pipeline producer: %.400s", ctx->lanes[c % NL]->err); }
                publish(prod_rc, rc);
                break;
            }
            publish(issued[(size_t)c], 1);
        }
    };
    std::vector<std::thread> producers;
    for (int t = 0; t < n_prod; ++t) producers.emplace_back(produce, t);
    int main_rc = DBGA_SUCCESS;
    for (int c = 0; c < nchunks; ++c) {
        wait_until([&]() { return issued[(size_t)c].load(std::memory_order_acquire) != 0 || prod_rc.load() != DBGA_SUCCESS; });
        if (issued[(size_t)c].load(std::memory_order_acquire) == 0) { main_rc = prod_rc.load(); break; }
        main_rc = commit(c);
        if (main_rc != DBGA_SUCCESS) break;
        publish(committed, c + 1);
    }
    if (spin) stop.store(true);
    else { { std::lock_guard<std::mutex> lk(mu); stop.store(true); } cv.notify_all(); }
    for (auto &t : producers) t.join();
    if (main_rc != DBGA_SUCCESS) {
        for (int i = 0; i < NL; ++i) cudaStreamSynchronize(ctx->lanes[i]->stream);
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->dl_stream);
        return main_rc;
    }
    // (every lane's last chunk was committed after its ev_ready fired: the lane streams are idle)
    {   // the per-pair columns of the whole batch, behind the last chunk's rows
        cudaStream_t st = ctx->dl_stream;
        DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_status.p, ctx->d_status_out.p, 4 * P, cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_score.p, ctx->d_score.p, 8 * P, cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_cells.p, ctx->d_cells.p, 8 * P, cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_nseg.p, ctx->d_nseg.p, 4 * P, cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_nanch.p, ctx->d_nanch.p, 8 * P, cudaMemcpyDeviceToHost, st));
        if (seg && (params->flags & DBGA_FLAG_SOP))
            DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_sop.p, ctx->d_sop.p, 16 * P, cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(cudaMemcpyAsync(aln_off, ctx->d_aln_off.p, 8 * (P + nchunks), cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(cudaMemcpyAsync(run_off, ctx->d_run_off.p, 8 * (P + nchunks), cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(wait_stream(ctx, st));
    }
    if (trace) {
        for (int c = 0; c < nchunks; ++c) {
            float t[6];
            for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&t[i], ctx->ev[0], tev[(size_t)c * 6 + i]);
            fprintf(stderr, "[dbga] chunk %d gpu: h2d %.2f-%.2f kernels %.2f-%.2f d2h %.2f-%.2f\n", c, t[0], t[1], t[2], t[3], t[4], t[5]);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    // compact the per-chunk offset arrays in place and add the chunk bases
    for (int c = 0; c < nchunks; ++c) {
        const int64_t a = lo[(size_t)c], n = lo[(size_t)c + 1] - a;
        for (int64_t i = 0; i < n; ++i) {
            aln_off[a + i] = (seg ? aln_off[a + c + i] : 0) + aln_base[(size_t)c];
            run_off[a + i] = run_off[a + c + i] + run_base[(size_t)c];
        }
    }
    aln_off[P] = aln_base[(size_t)nchunks];
    run_off[P] = run_base[(size_t)nchunks];
    memset(out, 0, sizeof(*out));
    out->num_pairs = P;
    out->status = (int32_t *)ctx->h_status.p; out->score = (int64_t *)ctx->h_score.p;
    out->dp_cells = (int64_t *)ctx->h_cells.p; out->dp_segments = (int32_t *)ctx->h_nseg.p;
    out->n_anchors = (int64_t *)ctx->h_nanch.p; out->run_off = run_off; out->runs = (int32_t *)ctx->h_runs.p;
    out->aln_off = aln_off; out->aln0 = (uint8_t *)ctx->h_aln0.p; out->aln1 = (uint8_t *)ctx->h_aln1.p;
    out->sop = (seg && (params->flags & DBGA_FLAG_SOP)) ? (int32_t *)ctx->h_sop.p : nullptr;
    out->seg_cols = compact ? (int32_t *)ctx->h_segcols.p : nullptr;
    out->ms_plan = 0.f;
    for (float v : plan_ms) out->ms_plan += v;
    if (no_runs) out->runs = nullptr;
    out->ms_stage1 = ms[0]; out->ms_stage2 = ms[1]; out->ms_stage3 = ms[2]; out->ms_assemble = ms[3];
    out->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    out->dp_cells_total = cells; out->dp_segments_total = nsegs; out->kernel_launches = launches;
    return DBGA_SUCCESS;
}

// packed != nullptr: 2-bit input in the canonical layout; sequence j of this call is sequence seq0 + j of
// that layout (a shard of a larger batch keeps the batch's word positions)
int align_common(dbga_ctx *ctx, const uint8_t *seqs, const uint32_t *packed, const int64_t *offsets, int64_t P,
                 const dbga_params *params, dbga_result *out, bool no_anchors, int64_t seq0 = 0) {
    if (!ctx || !out || !offsets || !params || (!seqs && !packed && P > 0)) return DBGA_ERR_ARG;
    DBGA_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (P >= 64 && !params->want_graph && ctx->own_stream) {
        const int64_t total = offsets[2 * P] - offsets[0];
        if (total >= (16LL << 20)) {
            // chunk sizes are in input bases; 2-bit input puts a quarter of the bytes on the wire, so its chunks
            // hold four times the bases (every chunk costs ~0.4 ms of host time and ~20 launches)
            int64_t chunk_mb = packed ? 128 : 32;
            if (const char *e = getenv("DBGA_CHUNK_MB")) chunk_mb = std::max(1, atoi(e));
            return align_pipelined(ctx, seqs, packed, seq0, offsets, P, params, out, no_anchors, std::max<int64_t>(chunk_mb << 20, total / 200));
        }
    }
    DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[0], ctx->stream));
    const int64_t base = P > 0 ? offsets[0] : 0;
    int rc = do_plan(ctx, offsets, P, params, no_anchors, base);
    if (rc != DBGA_SUCCESS) return rc;
    RESERVE(d_seqs, std::max<int64_t>(ctx->total_bytes, 16) + 16);
    int extra_launches = 0;
    if (packed && P > 0) {
        const int64_t w_lo = (offsets[0] >> 4) + seq0, w_hi = (offsets[2 * P] >> 4) + seq0 + 2 * P;
        RESERVE(d_pk_in, 4 * (w_hi - w_lo) + 64);
        DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_pk_in.p, packed + w_lo, (size_t)(4 * (w_hi - w_lo)), cudaMemcpyHostToDevice, ctx->stream));
        DBGA_CUDA_CHECK(launch_unpack((const uint32_t *)ctx->d_pk_in.p, (const PairDesc *)ctx->d_pairs.p, P, base, seq0, w_lo,
                                      (uint8_t *)ctx->d_seqs.p, ctx->num_sms, ctx->stream));
        extra_launches = 1;
    } else if (ctx->total_bytes) {
        DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_seqs.p, seqs + base, (size_t)ctx->total_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = do_run(ctx, (const uint8_t *)ctx->d_seqs.p);
    if (rc != DBGA_SUCCESS) return rc;
    rc = do_fetch(ctx, out);
    if (rc != DBGA_SUCCESS) return rc;
    out->kernel_launches += extra_launches;
    cudaEventElapsedTime(&out->ms_h2d, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&out->ms_total, ctx->ev[0], ctx->ev[7]);
    return DBGA_SUCCESS;
}

}  // namespace

extern "C" {

int dbga_version(void) { return 100; }

int dbga_create(int device, dbga_ctx **out) { return create_ctx(device, out); }

}  // extern "C"

namespace {
int create_ctx(int device, dbga_ctx **out) {
    if (!out) return DBGA_ERR_ARG;
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);   // only effective before the CUDA context exists
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DBGA_ERR_CUDA;
    dbga_ctx *ctx = new dbga_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DBGA_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return DBGA_ERR_CUDA; }
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return DBGA_ERR_CUDA; }
    for (auto &e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return DBGA_ERR_CUDA; }
    if (cudaEventCreateWithFlags(&ctx->ev_block, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) { delete ctx; return DBGA_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return DBGA_ERR_CUDA; }
    *out = ctx;
    return DBGA_SUCCESS;
}
}  // namespace

extern "C" {

void dbga_destroy(dbga_ctx *ctx) {
    if (!ctx) return;
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); ctx->copy_stream = nullptr; }
    if (ctx->dl_stream) { cudaStreamSynchronize(ctx->dl_stream); cudaStreamDestroy(ctx->dl_stream); ctx->dl_stream = nullptr; }
    for (auto &l : ctx->lanes) if (l) {
        if (l->ev_dl) cudaEventDestroy(l->ev_dl);
        if (l->ev_ready) cudaEventDestroy(l->ev_ready);
        if (l->ev_meta) cudaEventDestroy(l->ev_meta);
        for (auto &e : l->ev_sub) if (e) cudaEventDestroy(e);
        for (auto &e : l->dbg_ev) if (e) cudaEventDestroy(e);
        if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
        dbga_destroy(l); l = nullptr;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *dbs[] = {&ctx->d_seqs, &ctx->d_pairs, &ctx->d_list, &ctx->d_large, &ctx->d_status, &ctx->d_aof, &ctx->d_nd0,
                     &ctx->d_nd1, &ctx->d_pout, &ctx->d_runs, &ctx->d_item_off, &ctx->d_slots, &ctx->d_lists, &ctx->d_tt, &ctx->d_ctr,
                     &ctx->d_tb, &ctx->d_bnd_w, &ctx->d_bnd_c, &ctx->d_str, &ctx->d_packed, &ctx->d_table, &ctx->d_tile_off,
                     &ctx->d_aln_off, &ctx->d_run_off, &ctx->d_scan, &ctx->d_aln0, &ctx->d_aln1,
                     &ctx->d_runs_out, &ctx->d_score, &ctx->d_cells, &ctx->d_nseg, &ctx->d_nanch, &ctx->d_status_out,
                     &ctx->d_sop, &ctx->d_segcols, &ctx->d_pk_in, &ctx->d_item_src, &ctx->d_tile_pair, &ctx->d_sink};
    for (DevBuf *b : dbs) if (b->p) cudaFree(b->p);
    HostBuf *hbs[] = {&ctx->h_status, &ctx->h_score, &ctx->h_cells, &ctx->h_nseg, &ctx->h_nanch, &ctx->h_aln_off,
                      &ctx->h_run_off, &ctx->h_aln0, &ctx->h_aln1, &ctx->h_runs, &ctx->h_nd0, &ctx->h_nd1, &ctx->h_poff,
                      &ctx->h_qoff, &ctx->h_ctr, &ctx->h_meta, &ctx->h_sop, &ctx->h_segcols};
    for (HostBuf *b : hbs) if (b->p) cudaFreeHost(b->p);
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *dbga_last_error(dbga_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int dbga_set_stream(dbga_ctx *ctx, void *cuda_stream) {
    if (!ctx) return DBGA_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return DBGA_SUCCESS;
}

int dbga_set_scratch_limit(dbga_ctx *ctx, int64_t bytes) {
    if (!ctx || bytes < (1 << 20)) return DBGA_ERR_ARG;
    ctx->scratch_limit = bytes;
    return DBGA_SUCCESS;
}

void *dbga_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (bytes <= 0) bytes = 1;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void dbga_host_free(void *p) { if (p) cudaFreeHost(p); }

int dbga_align_batch(dbga_ctx *ctx, const uint8_t *seqs, const int64_t *offsets, int64_t num_pairs,
                     const dbga_params *params, dbga_result *out) {
    return align_common(ctx, seqs, nullptr, offsets, num_pairs, params, out, false);
}

int dbga_align_batch_packed(dbga_ctx *ctx, const uint32_t *packed, const int64_t *offsets, int64_t num_pairs,
                            const dbga_params *params, dbga_result *out) {
    if (!packed && num_pairs > 0) return DBGA_ERR_ARG;
    return align_common(ctx, nullptr, packed, offsets, num_pairs, params, out, false);
}

int dbga_global_align_batch(dbga_ctx *ctx, const uint8_t *seqs, const int64_t *offsets, int64_t num_pairs,
                            const dbga_params *params, dbga_result *out) {
    if (!params) return DBGA_ERR_ARG;
    dbga_params p = *params;
    p.stages = 3; p.want_graph = 0;
    return align_common(ctx, seqs, nullptr, offsets, num_pairs, &p, out, true);
}

int dbga_plan(dbga_ctx *ctx, const int64_t *offsets, int64_t num_pairs, const dbga_params *params) {
    if (!ctx) return DBGA_ERR_ARG;
    return do_plan(ctx, offsets, num_pairs, params, false);
}
int dbga_plan_global(dbga_ctx *ctx, const int64_t *offsets, int64_t num_pairs, const dbga_params *params) {
    if (!ctx || !params) return DBGA_ERR_ARG;
    dbga_params p = *params;
    p.stages = 3; p.want_graph = 0;
    return do_plan(ctx, offsets, num_pairs, &p, true);
}
int dbga_plan_run(dbga_ctx *ctx, const uint8_t *device_seqs) {
    if (!ctx) return DBGA_ERR_ARG;
    return do_run(ctx, device_seqs);
}
int dbga_plan_fetch(dbga_ctx *ctx, dbga_result *out) {
    if (!ctx || !out) return DBGA_ERR_ARG;
    return do_fetch(ctx, out);
}

int dbga_kmer_stats(dbga_ctx *ctx, const uint8_t *seqs, const int64_t *offsets, int64_t P, int32_t k,
                    int32_t *status, int64_t *distinct0, int64_t *distinct1, int64_t *shared) {
    if (!ctx) return DBGA_ERR_ARG;
    if (P < 0 || (P > 0 && (!seqs || !offsets || !status || !distinct0 || !distinct1 || !shared)))
        return fail(ctx, DBGA_ERR_ARG, "null argument");
    if (k < 1) return fail(ctx, DBGA_ERR_ARG, "k must be >= 1");
    if (P == 0) return DBGA_SUCCESS;
    if (P > 0x7fffff00LL) return fail(ctx, DBGA_ERR_ARG, "too many pairs in one batch");
    DBGA_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int64_t base = offsets[0], total = offsets[2 * P] - base;
    // pairs whose table fits shared memory share one launch (sized for the largest of them);
    // the others are counted one at a time in a global-memory table
    const int key_bytes = k <= 15 ? 4 : 8;
    auto cap_of = [](int64_t keys) { return ((uint64_t)keys * 3 / 2 + 64 + 15) / 16 * 16; };
    auto words_of = [](int64_t n) { return (int)((((n + 15) / 16 + 3) + 1) & ~1LL); };
    int64_t max_keys = 0, max_n = 0;
    std::vector<int64_t> large;
    for (int64_t i = 0; i < P; ++i) {
        const int64_t n0 = offsets[2 * i + 1] - offsets[2 * i], n1 = offsets[2 * i + 2] - offsets[2 * i + 1];
        if (n0 < 0 || n1 < 0) return fail(ctx, DBGA_ERR_ARG, "offsets must be non-decreasing");
        if (n0 < k || n1 < k) continue;            // K_INVALID, decided on the device
        const int64_t keys = (n0 - k + 1) + (n1 - k + 1), nmax = std::max(n0, n1);
        if (k > 31 || kmer_stats_smem(key_bytes, (uint32_t)std::min<uint64_t>(cap_of(keys), 1u << 30), words_of(nmax)) > SMEM_LIMIT - 1024) {
            large.push_back(i);
            continue;
        }
        max_keys = std::max(max_keys, keys); max_n = std::max(max_n, nmax);
    }
    cudaStream_t st = ctx->stream;
    RESERVE(d_seqs, (size_t)total + 64);
    RESERVE(d_aln_off, 8 * (2 * P + 1));           // offsets
    RESERVE(d_status, 4 * P);
    RESERVE(d_nseg, 12 * P);                        // three counts per pair
    HRESERVE(h_nseg, 12 * P); HRESERVE(h_status, 4 * P);
    DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_seqs.p, seqs + base, (size_t)total, cudaMemcpyHostToDevice, st));
    DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->d_aln_off.p, offsets, 8 * (size_t)(2 * P + 1), cudaMemcpyHostToDevice, st));
    if (large.size() < (size_t)P) {
        DBGA_CUDA_CHECK(launch_kmer_stats((const uint8_t *)ctx->d_seqs.p, (const int64_t *)ctx->d_aln_off.p, base, P, k,
                                          (uint32_t)cap_of(max_keys), words_of(max_n), max_keys, max_n,
                                          (int32_t *)ctx->d_status.p, (int32_t *)ctx->d_nseg.p, ctx->num_sms, st));
    }
    DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_nseg.p, ctx->d_nseg.p, 12 * (size_t)P, cudaMemcpyDeviceToHost, st));
    DBGA_CUDA_CHECK(cudaMemcpyAsync(ctx->h_status.p, ctx->d_status.p, 4 * (size_t)P, cudaMemcpyDeviceToHost, st));
    DBGA_CUDA_CHECK(cudaStreamSynchronize(st));
    const int32_t *c = (const int32_t *)ctx->h_nseg.p, *hs = (const int32_t *)ctx->h_status.p;
    for (int64_t i = 0; i < P; ++i) {
        const int64_t n0 = offsets[2 * i + 1] - offsets[2 * i], n1 = offsets[2 * i + 2] - offsets[2 * i + 1];
        if (n0 < k || n1 < k) {                    // utils.py:129-130 (such a pair may be outside the launch's size class)
            status[i] = DBGA_K_INVALID; distinct0[i] = distinct1[i] = shared[i] = 0;
            continue;
        }
        status[i] = hs[i]; distinct0[i] = c[3 * i]; distinct1[i] = c[3 * i + 1]; shared[i] = c[3 * i + 2];
    }
    for (int64_t i : large) {
        const int64_t o0 = offsets[2 * i] - base, o1 = offsets[2 * i + 1] - base, o2 = offsets[2 * i + 2] - base;
        const uint64_t lcap = cap_of((o1 - o0 - k + 1) + (o2 - o1 - k + 1));
        RESERVE(d_table, lcap * 12 + 64);
        unsigned long long *keys = (unsigned long long *)ctx->d_table.p;
        uint32_t *flag = (uint32_t *)(keys + lcap);
        unsigned int *counts = (unsigned int *)ctx->d_status.p;      // the small pairs' results are on the host already
        DBGA_CUDA_CHECK(launch_kmer_stats_large((const uint8_t *)ctx->d_seqs.p + o0, o1 - o0, (const uint8_t *)ctx->d_seqs.p + o1,
                                                o2 - o1, k, keys, flag, lcap, counts, ctx->num_sms, st));
        unsigned int hc[4];
        DBGA_CUDA_CHECK(cudaMemcpyAsync(hc, counts, sizeof(hc), cudaMemcpyDeviceToHost, st));
        DBGA_CUDA_CHECK(cudaStreamSynchronize(st));
        status[i] = hc[3] ? DBGA_BAD_CHAR : DBGA_OK;
        distinct0[i] = hc[3] ? 0 : hc[0]; distinct1[i] = hc[3] ? 0 : hc[1]; shared[i] = hc[3] ? 0 : hc[2];
    }
    ctx->planned = false; ctx->ran = false;          // the plan's device buffers were reused: no stale fetch
    return DBGA_SUCCESS;
}

int dbga_measure_int_peak(dbga_ctx *ctx, double out[3]) {
    if (!ctx || !out) return DBGA_ERR_ARG;
    DBGA_CUDA_CHECK(cudaSetDevice(ctx->device));
    RESERVE(d_sink, 256);
    const int iters = 4096;
    for (int which = 0; which < 3; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[0], ctx->stream));
            DBGA_CUDA_CHECK(launch_int_peak(which, iters, (int *)ctx->d_sink.p, ctx->num_sms, ctx->stream));
            DBGA_CUDA_CHECK(cudaEventRecord(ctx->ev[1], ctx->stream));
            DBGA_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
            if (rep > 0 && ms < best) best = ms;
        }
        const double ops = (double)ctx->num_sms * 8 * 256 * (double)iters * 16 * 9;
        out[which] = ops / (best * 1e-3);
    }
    return DBGA_SUCCESS;
}


// ---------------------------------------------------------------- host helpers (no GPU work)
int64_t dbga_packed_words(const int64_t *offsets, int64_t num_seqs) {
    if (!offsets || num_seqs < 0) return -1;
    return (offsets[num_seqs] >> 4) + num_seqs + 1;
}

}  // extern "C"

namespace {
template <class Fn> void parallel_ranges(int64_t n, int threads, Fn fn) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<int64_t>(threads, std::max<int64_t>(1, n / 256));
    if (threads <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([=]() { fn(n * t / threads, n * (t + 1) / threads); });
    for (auto &x : th) x.join();
}
}  // namespace

extern "C" {

int64_t dbga_pack_ascii(const uint8_t *seqs, const int64_t *offsets, int64_t num_seqs, uint32_t *packed, int threads) {
    if (!seqs || !offsets || !packed || num_seqs < 0) return -1;
    std::atomic<int64_t> bad{0};
    parallel_ranges(num_seqs, threads, [&](int64_t lo, int64_t hi) {
        int64_t nbad = 0;
        for (int64_t j = lo; j < hi; ++j) {
            const uint8_t *src = seqs + offsets[j];
            const int64_t len = offsets[j + 1] - offsets[j];
            uint32_t *dst = packed + (offsets[j] >> 4) + j;
            const bool b = host_pack_seq(src, len, dst);
            nbad += b;
        }
        bad += nbad;
    });
    return bad.load();
}

int dbga_expand_rows(const uint8_t *seqs, const int64_t *offsets, const dbga_result *res, uint8_t *row0, uint8_t *row1,
                     int64_t cap, int64_t *aln_off_full, int threads) {
    if (!seqs || !offsets || !res || !row0 || !row1 || !aln_off_full) return DBGA_ERR_ARG;
    const int64_t P = res->num_pairs;
    if (P > 0 && (!res->seg_cols || !res->runs || !res->run_off || !res->aln_off || !res->status)) return DBGA_ERR_ARG;
    // full length of every pair: its segment columns + the run bodies (the last merge node belongs to the tail)
    parallel_ranges(P, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            int64_t len = 0;
            if (res->status[i] == DBGA_OK) {
                const int64_t r0 = res->run_off[i], R = res->run_off[i + 1] - r0;
                len = res->aln_off[i + 1] - res->aln_off[i];
                for (int64_t r = 0; r < R; ++r) len += res->runs[3 * (r0 + r) + 2];
                if (R > 0) len -= 1;
            }
            aln_off_full[i + 1] = len;
        }
    });
    aln_off_full[0] = 0;
    for (int64_t i = 0; i < P; ++i) aln_off_full[i + 1] += aln_off_full[i];
    if (aln_off_full[P] > cap) return DBGA_ERR_ARG;
    parallel_ranges(P, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            if (res->status[i] != DBGA_OK) continue;
            const int64_t r0 = res->run_off[i], R = res->run_off[i + 1] - r0;
            const int32_t *sc = res->seg_cols + r0 + i;
            const uint8_t *c0 = res->aln0 + res->aln_off[i], *c1 = res->aln1 + res->aln_off[i];
            const uint8_t *s0 = seqs + offsets[2 * i];
            uint8_t *o0 = row0 + aln_off_full[i], *o1 = row1 + aln_off_full[i];
            for (int64_t t = 0; t <= R; ++t) {
                const int32_t cols = sc[t];
                memcpy(o0, c0, (size_t)cols); memcpy(o1, c1, (size_t)cols);
                o0 += cols; o1 += cols; c0 += cols; c1 += cols;
                if (t < R) {          // inside a run both rows are the input bytes (debruijn_pairwise.py:384-388, 508-509)
                    const int32_t a = res->runs[3 * (r0 + t)], len = res->runs[3 * (r0 + t) + 2] - (t == R - 1 ? 1 : 0);
                    memcpy(o0, s0 + a, (size_t)len); memcpy(o1, s0 + a, (size_t)len);
                    o0 += len; o1 += len;
                }
            }
        }
    });
    return DBGA_SUCCESS;
}

// ---------------------------------------------------------------- one call, several GPUs
struct dbga_multi {
    int n = 0;
    dbga_ctx *ctx[DBGA_MAX_SHARDS] = {};
    char err[512] = {0};
};

int dbga_multi_create(const int *devices, int num_devices, dbga_multi **out) {
    if (!out || !devices || num_devices < 1 || num_devices > DBGA_MAX_SHARDS) return DBGA_ERR_ARG;
    *out = nullptr;
    dbga_multi *m = new dbga_multi();
    for (int d = 0; d < num_devices; ++d) {
        const int rc = create_ctx(devices[d], &m->ctx[d]);
        if (rc != DBGA_SUCCESS) { for (int e = 0; e < d; ++e) dbga_destroy(m->ctx[e]); delete m; return rc; }
        m->ctx[d]->multi_peers = num_devices;
        ++m->n;
    }
    *out = m;
    return DBGA_SUCCESS;
}

void dbga_multi_destroy(dbga_multi *m) {
    if (!m) return;
    for (int d = 0; d < m->n; ++d) dbga_destroy(m->ctx[d]);
    delete m;
}

const char *dbga_multi_last_error(dbga_multi *m) { return m ? m->err : "null handle"; }

int dbga_align_batch_multi(dbga_multi *m, const uint8_t *seqs, const uint32_t *packed, const int64_t *offsets,
                           int64_t num_pairs, const dbga_params *params, dbga_multi_result *out) {
    if (!m || !out || !offsets || !params || num_pairs < 0 || (!seqs && !packed && num_pairs > 0)) return DBGA_ERR_ARG;
    const auto t0 = std::chrono::steady_clock::now();
    memset(out, 0, sizeof(*out));
    const int n = m->n;
    out->num_shards = n;
    // contiguous shards of (nearly) equal bytes: pair p goes to the shard its first byte falls into
    const int64_t b0 = offsets[0], total = num_pairs > 0 ? offsets[2 * num_pairs] - b0 : 0;
    out->pair_lo[0] = 0;
    for (int d = 1; d < n; ++d) {
        const int64_t target = b0 + total * d / n;
        int64_t lo = out->pair_lo[d - 1], hi = num_pairs;
        while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (offsets[2 * mid] < target) lo = mid + 1; else hi = mid; }
        out->pair_lo[d] = lo;
    }
    out->pair_lo[n] = num_pairs;
    int rcs[DBGA_MAX_SHARDS] = {};
    std::vector<std::thread> th;
    for (int d = 0; d < n; ++d) {
        th.emplace_back([&, d]() {
            const int64_t lo = out->pair_lo[d], cnt = out->pair_lo[d + 1] - lo;
            rcs[d] = align_common(m->ctx[d], seqs, packed, offsets + 2 * lo, cnt, params, &out->shard[d], false, 2 * lo);
        });
    }
    for (auto &x : th) x.join();
    for (int d = 0; d < n; ++d)
        if (rcs[d] != DBGA_SUCCESS) {
            snprintf(m->err, sizeof(m->err), "shard %d (device %d): %.400s", d, m->ctx[d]->device, m->ctx[d]->err);
            return rcs[d];
        }
    out->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return DBGA_SUCCESS;
}

}  // extern "C"
