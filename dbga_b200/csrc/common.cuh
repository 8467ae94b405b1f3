// Shared device/host definitions for the dbga_b200 kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/dbga_b200.h"

namespace dbga {

// ------------------------------------------------------------------ data layout in HBM
// One record per pair, built on the host from `offsets` (plan time).
struct PairDesc {
    int64_t off0, off1;   // byte offsets of seq1 / seq2 in the concatenated buffer
    int32_t n0, n1;       // lengths
    int64_t q_base;       // prefix of (n1-k+1) over pairs: base of this pair in aof[]
    int64_t p_base;       // prefix of (n0-k+1): base in nondup0 flags / large-path memo
};

// Per-pair outcome of stages 1-2 (device), consumed by stage 3 and the assembler.
struct PairOut {
    int32_t status;
    int32_t n_runs;       // maximal runs of consecutive anchors (a+u, b+u)
    int64_t run_base;     // first run in runs pool (int32 triples)
    int64_t slot_base;    // first segment slot; slots: [first, mid_1..mid_{R-1}, tail]
    int32_t n_slots;
    int32_t pad;
    int64_t n_anchors;
    int64_t aln_len;      // filled by the assembler's length pass
};

// One alignment segment (what the reference hands to dna_global_aln, utils.py:157-192).
struct Slot {
    int32_t x0, m;        // seq1[x0 : x0+m]
    int32_t y0, n;        // seq2[y0 : y0+n]
    int32_t score;        // DP score (0 when one side is empty)
    int32_t len;          // aligned columns produced (before any first-column drop)
    int64_t str_off;      // scratch: reversed rows; row x at [str_off, +m+n), row y after
    int64_t tb_off;       // warp / cta classes: traceback bytes in the scratch pool; tiny classes: pd.off0
    int32_t pair;
    int32_t cls;          // DP class (CLS_*)
    int32_t out_off;      // assembly: first output column of this slot, relative to the pair
    int32_t skip;         // assembly: 1 = drop the first column (mid segments, debruijn_pairwise.py:499-500)
    int32_t row_len;      // n0 + n1 of the pair: row y of the scratch sits row_len bytes after row x
    int32_t n0;           // with tb_off = off0 the tiny kernels find both sequences without reading PairDesc
};
static_assert(sizeof(Slot) == 64, "Slot is four 16-byte words");

// DP classes: 0..3 = thread-per-segment ("tiny", bucketed by max(m, n) so that the segments of
// a warp have similar loop bounds),
// 4..7 = warp wavefront with 2 / 4 / 8 / 16 rows per lane (segments of up to 64 / 128 / 256 / any rows;
// more rows per lane amortise the per-step shuffles, fewer keep all 32 lanes busy on short segments),
// 8 = CTA wavefront (8 rows per lane, 8 warps).
enum : int32_t { CLS_NONE = -1, CLS_TINY0 = 0, NUM_TINY = 4, CLS_WARP2 = 4, CLS_WARP4 = 5, CLS_WARP8 = 6, CLS_WARP16 = 7, CLS_CTA = 8, NUM_CLS = 9 };

// DP geometry (shared by host sizing and kernels)
constexpr int TINY_MAX = 32;     // thread-per-segment kernel: m, n <= 32
constexpr int TINY_THREADS = 128;
constexpr int WF_R = 8;          // rows per lane in the widest wavefront kernels
constexpr int WF_ROWS_PER_WARP = 32 * WF_R;   // 256
__host__ __device__ inline int wf_rows_per_lane(int cls) { return cls == CLS_WARP2 ? 2 : (cls == CLS_WARP4 ? 4 : (cls == CLS_WARP16 ? 16 : 8)); }
constexpr int WF_CTA_WARPS = 8;  // long segments, many of them: one CTA of 8 warps x 8 rows per lane each
constexpr int WF_CLUSTER = 8;    // long segments, few of them: a thread-block cluster of 8 CTAs each,
constexpr int WF_CL_WARPS = 4;   //   4 warps per CTA (one per SM sub-partition),
constexpr int WF_CL_R = 4;       //   4 rows per lane (traceback storage is sized for 8, which covers both)
constexpr int WF_D = 64;         // step delay between consecutive warps
constexpr int WF_C = 32;         // steps per barrier
constexpr int WF_RING = 128;     // ring entries (columns) per warp boundary
// A segment above warp_max_cells gets a whole CTA (8 pipelined warps), below it one warp.
// With few pairs in the batch a big segment needs the CTA's parallelism; with many pairs
// there are enough segments to fill the chip warp by warp, which wastes no pipeline fill.
constexpr int64_t WARP_CLASS_MAX_CELLS_FEW = 1 << 18;
constexpr int64_t WARP_CLASS_MAX_CELLS_MANY = 1 << 24;
constexpr int64_t MANY_PAIRS = 592;

// Scores are kept as 64*V + 21*tag(state): the three 2-bit tag fields make every max()
// break ties in the configured state order and leave the winner's identity in the low
// bits (bits 0-1 used for the M state, 2-3 for X, 4-5 for Y).
constexpr int SCALE = 64;
constexpr int TAGMUL = 21;
constexpr int32_t NEG_V = -(1 << 24);          // "minus infinity" in score units
constexpr int64_t MAX_ABS_SCORE = (1 << 24) - (1 << 20);

// Packed 16-bit wavefront (dp_wf16_kernel): two cells per 32-bit register, scores as
// 4 * (V - off) + priority in each int16 half.  Built by make_dp_params; ok = 0 when the
// scoring parameters do not fit the packed form (the 32-bit kernels then take everything).
struct Dp16 {
    uint32_t G;               // 4 * (ge - go) in both halves
    uint32_t tagM, tagX, tagY; // priority (0..2) in both halves
    uint32_t lut[4];          // per seq1 base: int8 4 * (S + 2 ge) + priority(M) for seq2 base A, C, G, T
    uint32_t mul[4];          // -1, 4, 16, 256: IMAD multipliers kept in registers
    uint32_t nmul[2];         // compact traceback (4 bits per cell): 4 / (pM - pX), 8 / (pM - pY) as 32-bit integers
    uint32_t tbk;             // 4 * priority(X) + 16 * priority(Y): bias of a raw traceback byte
    int32_t q0, qneg, qy0;    // 16-bit encodings (no tag) of M[0][0], "minus infinity", Y^[0][c]
    int32_t off;              // V offset: stored = 4 * (V - off) + priority
    int32_t per_step;         // max(0, largest substitution score) + 2 ge: growth of H^ per diagonal step
    int32_t ulimit;           // largest H^ the encoding can hold
    int32_t ok;
    int32_t pair;             // two segments per warp (dp_wf16p_kernel) for segments of up to 512 rows
    int32_t strip;            // long segments as row strips in per-strip, per-column-block frames (dp_wf16s_kernel): rows per half-lane (8 or 4), 0 = off
};

struct DpParams {
    int32_t go, ge;           // SCALE * gap open (first gap char), SCALE * extend
    int32_t tagM, tagX, tagY; // TAGMUL * priority (0..2); higher priority wins ties
    int32_t endM, endX, endY; // priorities at the terminal cell
    int32_t xy;               // allow X<->Y
    uint32_t lut_lo[4];       // per seq1 base: int16 (SCALE*S + tagM) for seq2 base A,C
    uint32_t lut_hi[4];       //                                             ...  G,T
    int32_t go_raw, ge_raw;
    Dp16 h;                   // packed 16-bit form of the same parameters
};

// Does segment (m, n) go to the packed 16-bit wavefront with R16 matrix rows per half-lane?
// One pass of 64 half-lanes covers 64 * R16 rows; rows are padded to whole lanes and one
// column past n is computed, and every such cell must stay below the encoding's ceiling.
__host__ __device__ inline bool wf16_ok(const Dp16 &h, int m, int n, int R16) {
    if (!h.ok) return false;
    const int mp = ((m + 2 * R16 - 1) / (2 * R16)) * (2 * R16);
    if (mp > 64 * R16) return false;
    const long long lim = mp < n + 1 ? mp : n + 1;
    return (long long)h.per_step * lim + 64 <= (long long)h.ulimit;
}
// Which packed kernel takes segment (m, n) of warp class cls: 0 = none (the class's 32-bit kernel
// runs it); R = the half-lane kernel with R rows per half-lane; WF16_PAIR + R = the two-segments-
// per-warp kernel with R rows per lane (h.pair: every segment of up to 32 R rows of the class,
// all of whose cells -- padding rows and the partner's extra columns included -- stay below
// the ceiling).  The open-ended class splits at 512 rows.
constexpr int WF16_PAIR = 64;
// Long segments (CLS_CTA) in the packed form: row strips of S16_ROWS rows, one warp per strip, strips of a
// segment pipelined through boundary rows in HBM/L2.  Values are kept relative to a frame base that is
// re-anchored per strip every S16_CB columns (on the strip's top boundary row), so the int16 range only has
// to hold per_step * max(S16_ROWS, S16_CB) whatever the segment's size.  S16_F0: where a frame puts its anchor.
constexpr int S16_ROWS = 512;                // the larger of the two strip heights (64 half-lanes x 8 or 4 rows)
constexpr int S16_CB = 256;
constexpr int S16_F0 = 300;
__host__ __device__ inline int wf16_pick(const Dp16 &h, int cls, int m, int n) {
    if (!h.ok) return 0;
    if (cls == CLS_CTA) return h.strip;              // rows per half-lane of the strip kernel (0: 32-bit kernels)
    if (h.pair) {
        const int R = wf_rows_per_lane(cls);
        // (not in the open-ended class: 16 rows per lane for two segments cost more registers than they save)
        if (cls != CLS_WARP16 && m <= 32 * R && (long long)h.per_step * (32 * R) + 64 <= (long long)h.ulimit) return WF16_PAIR + R;
    }
    const int r16 = cls == CLS_WARP2 ? 1 : (cls == CLS_WARP4 ? 2 : (cls == CLS_WARP8 ? 4 : (m <= 512 ? 8 : 16)));
    return wf16_ok(h, m, n, r16) ? r16 : 0;
}

// traceback bytes of one wavefront segment: one block of step rows x 32*R columns per 32*R
// matrix rows, whatever number of warps shares the segment.  The 32-bit kernels use n + 32
// step rows per block, the packed 16-bit kernel (same row width, 64 half-lanes) n + WF16_TB_PAD.
constexpr int WF16_TB_PAD = 82;       // 64 half-lanes of skew + whole chunks of 16 steps
__host__ __device__ inline int64_t tb_bytes_wavefront(int64_t m, int64_t n, int rows_per_lane) {
    const int64_t rows_per_warp = 32 * rows_per_lane;
    return ((m + rows_per_warp - 1) / rows_per_warp) * (n + WF16_TB_PAD) * rows_per_warp;
}

// Opt-in limit for dynamic shared memory, raised to the device maximum (less the kernel's static
// shared memory) at every launch site: the value a kernel is allowed never depends on the launch,
// so host threads that launch the same kernel with different sizes (pipeline producers, one thread
// per GPU) cannot undercut each other.
template <class Kernel> inline cudaError_t allow_max_dyn_smem(Kernel kern) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    int dev = 0, optin = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

// ------------------------------------------------------------------ small device utils
__device__ __forceinline__ int base_code(uint8_t c) {
    // A=0 C=1 G=2 T=3 for upper-case ACGT; callers validate separately
    return ((c >> 1) ^ (c >> 2)) & 3;
}
__device__ __forceinline__ bool is_acgt(uint8_t c) {
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t hash64(uint64_t z) {
    z ^= z >> 33; z *= 0xff51afd7ed558ccdULL; z ^= z >> 33;
    z *= 0xc4ceb9fe1a85ec53ULL; z ^= z >> 33;
    return (uint32_t)z;
}

#define DBGA_CUDA_CHECK(expr)                                                        \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d: %s: %s", __FILE__, __LINE__, \
                     #expr, cudaGetErrorString(_e));                                 \
            return DBGA_ERR_CUDA;                                                    \
        }                                                                            \
    } while (0)

}  // namespace dbga
