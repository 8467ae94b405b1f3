// Stage 3: affine-gap (Gotoh) global alignment of every inter-anchor segment, batched
// across segments and pairs.
//
// Replaces cogent3.align.global_pairwise as called from dna_global_aln
// (reference src/dbga/utils.py:157-192, call site :185) for every segment the reference's
// bubble_aln (src/dbga/debruijn_pairwise.py:345-395) hands to it.  Integer three-state
// recurrence, bit-exact against oracle/ (conventions: SURVEY.md 8c):
//     M[i][j] = S(x_i, y_j) + max(M, X, Y)[i-1][j-1]
//     X[i][j] = max(M[i-1][j] - go, X[i-1][j] - ge [, Y[i-1][j] - go])     (x_i vs gap)
//     Y[i][j] = max(M[i][j-1] - go, Y[i][j-1] - ge [, X[i][j-1] - go])     (gap vs y_j)
// All three states are stored with the same offset ge*(i+j) added (H^ = H + ge*(i+j)): a gap
// extension then costs nothing and a gap open costs go - ge, so X and Y are one fused
// add-max each (no separate subtraction); the match state absorbs 2*ge into its score add.
// Scores live in registers as 64*V + 21*prio(state): one integer max (VIMNMX3 /
// VIADDMNMX) both selects the best predecessor under the configured first-wins tie order
// and leaves the winner's identity in the low bits, which become the traceback byte
// (bits 0-1: predecessor of M, 2-3: of X, 4-5: of Y) with two LOP3s.
//
// Kernels:
//   dp_thread_kernel    one thread per tiny segment (m, n <= 32), previous row in shared
//                       memory, work lists bucketed by size;
//   dp_wf_kernel<W,R,XY,CL>  anti-diagonal wavefront: R rows per lane, 32 lanes per warp,
//                       W warps per CTA pipelined through shared-memory boundary rings, CL CTAs
//                       per thread-block cluster with the ring of a CTA's last warp living in
//                       the next CTA's shared memory (DSMEM).  W = CL = 1, R = 2/4/8/16: one
//                       warp per segment (throughput classes).  W = 4, R = 4, CL = 8: one
//                       cluster per long segment, 32 warps on 32 SM sub-partitions -- a single
//                       warp runs the recurrence at dependent-issue latency, so a lone big
//                       segment is spread as thin as the wavefront skew allows.  Row strips of
//                       32*R*W*CL rows, boundary rows between strips staged through HBM;
//                       traceback bytes in HBM in a step-major (skewed) layout so that every
//                       warp step writes one contiguous 32*R-byte row.  Traceback: one warp,
//                       shared-memory window over the table, 32-cell diagonal probes.
#include <cstdlib>

#include "kernels.cuh"
#include "dp_walk.cuh"
#include <cooperative_groups.h>

namespace dbga {
namespace cg = cooperative_groups;

__device__ __forceinline__ int strip(int v) { return v & ~63; }

// One cell.  (Ml,Xl,Yl): in = (i, j-1), out = (i, j).  d*: (i-1, j-1).  u*: (i-1, j).
template <bool XY>
__device__ __forceinline__ uint32_t dp_cell(int &Ml, int &Xl, int &Yl, int dM, int dX, int dY, int uM,
                                            int uX, int uY, int sprime, int gmo, int ge2, int tagX, int tagY) {
    const int t = __vimax3_s32(dM, dX, dY);
    int u = __viaddmax_s32(uM, gmo, uX);          // max(M_up + (ge - go), X_up)
    int v = __viaddmax_s32(Ml, gmo, Yl);
    if (XY) {
        u = max(u, uY + gmo);
        v = max(v, Xl + gmo);
    }
    Ml = strip(t) + sprime + ge2;
    Xl = strip(u) | tagX;
    Yl = strip(v) | tagY;
    // bit-select: low 2 bits of t, bits 2-3 of u, bits 4-5 of v
    uint32_t d = ((uint32_t)t & 3u) | ((uint32_t)u & ~3u);
    d = (d & 15u) | ((uint32_t)v & ~15u);
    return d & 0xFFu;
}

// int16 score LUT for the row's seq1 base, indexed by the seq2 base through PRMT
// (__byte_perm masks the selector with 0x7777, which would drop the sign-replicate bits)
__device__ __forceinline__ int lut_score(uint32_t lo, uint32_t hi, uint32_t sel) {
    int r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
    return r;
}
__device__ __forceinline__ uint32_t lut_sel(int ycode) { return 0x9910u + 0x2222u * (uint32_t)ycode; }

// terminal-cell choice under end_order (first wins, strict >)
__device__ __forceinline__ void pick_end(int Mv, int Xv, int Yv, const DpParams &p, int &state, int &score) {
    const long long cm = (long long)strip(Mv) + p.endM, cx = (long long)strip(Xv) + p.endX, cy = (long long)strip(Yv) + p.endY;
    long long b = cm; state = 0;
    if (cx > b) { b = cx; state = 1; }
    if (cy > b) { b = cy; state = 2; }
    score = (int)(b >> 6);          // still carries the offset ge*(m+n); the caller removes it
}

// ============================================================================ tiny segments
// One THREAD per segment (m, n <= 32).  Measured on cfg3 (round 1, profiles/r01_*): with 8
// lanes per ~10x10 segment the forward pass ran at 6 % lane efficiency and the serial
// traceback idled 30 of 32 lanes; a thread per segment keeps every lane busy in both phases.
// Row sweep: the previous row lives in shared memory as two values per column,
//   T[j] = max(M, X, Y)[i-1][j]   (tagged; the diagonal candidate of row i) and
//   U[j] = max(M[i-1][j] - go, X[i-1][j] - ge [, Y[i-1][j] - go])   (the X state of row i),
// laid out [column][thread] (bank = thread: conflict free).  Loops run to the warp-wide
// maximum of (m, n) and are branch-free: lanes outside their own segment compute and
// store garbage into their private columns.  Work lists are bucketed by max(m, n) in
// stage 2, so the bounds inside a warp are similar.  Traceback bytes go to a per-CTA
// scratch in HBM/L2, four columns per 32-bit word, [word][thread] (coalesced).
template <bool XY>
__global__ void __launch_bounds__(TINY_THREADS)
dp_thread_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, Slot *__restrict__ slots,
                 const int32_t *__restrict__ lists, int64_t list_cap, const unsigned int *__restrict__ counts_dev,
                 int c0, int c1, TinyBounds tb, DpParams prm, uint8_t *__restrict__ strbuf,
                 uint32_t *__restrict__ tb_scratch, int tb_in_smem) {
    extern __shared__ int smem_i[];
    constexpr int TH = TINY_THREADS;
    const int tid = threadIdx.x;
    int bmax = 0;
#pragma unroll
    for (int c = 0; c < NUM_TINY; ++c) if (c == c1 - 1) bmax = (tb.b[c] + 3) & ~3;   // shared memory is sized for the largest class
    int *Tc = smem_i + tid;                                   // Tc[j * TH], j = 0..bmax
    int *Uc = Tc + (bmax + 1) * TH;
    uint16_t *selc = reinterpret_cast<uint16_t *>(smem_i + 2 * (bmax + 1) * TH) + tid;   // selc[(j-1) * TH]
    const int gmo = prm.ge - prm.go, ge2 = 2 * prm.ge, tagM = prm.tagM, tagX = prm.tagX, tagY = prm.tagY;
    const int NEG = NEG_V * SCALE;
    // score LUT in shared memory: indexing kernel parameters with a run-time base code would
    // compile into a long predicated cascade
    __shared__ uint32_t s_lut[8];
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { s_lut[c] = prm.lut_lo[c]; s_lut[4 + c] = prm.lut_hi[c]; }
    }
    __syncthreads();

  // one launch walks the work lists of classes c0..c1-1 back to back: a CTA that runs out of
  // items in one class moves on to the next, so the classes' tails overlap
  for (int cls = c0; cls < c1; ++cls) {
    int bound = 0, prev_bound = -1;
#pragma unroll
    for (int c = 0; c < NUM_TINY; ++c) { if (c == cls) bound = tb.b[c]; if (c + 1 == cls) prev_bound = tb.b[c]; }
    if (bound == prev_bound) continue;                        // empty class
    const int32_t *list = lists + (int64_t)cls * list_cap;
    const unsigned int count = counts_dev[cls];
    const int wpr = (bound + 3) >> 2;                         // traceback words per row
    // traceback words: shared memory for the small classes (a dependent chain of L2 round trips
    // per traceback step was the top stall), per-CTA scratch in HBM/L2 otherwise
    uint32_t *tbw = tb_in_smem ? reinterpret_cast<uint32_t *>(smem_i + 2 * (bmax + 1) * TH) + (bmax * TH) / 2 + tid
                               : tb_scratch + (size_t)blockIdx.x * ((size_t)bmax * ((bmax + 3) >> 2) * TH) + tid;
    for (unsigned int base = blockIdx.x * TH; base < count; base += gridDim.x * TH) {
        const unsigned int item = base + tid;
        const bool have = item < count;
        int m = 0, n = 0, sid = 0;
        const uint8_t *xs = seqs, *ys = seqs;
        int64_t str_off = 0, row_len = 0;
        if (have) {
            sid = list[item];
            const Slot s = slots[sid];
            m = s.m; n = s.n;
            xs = seqs + s.tb_off + s.x0;                          // tiny classes: tb_off = pd.off0, n0 = pd.off1 - pd.off0
            ys = seqs + s.tb_off + s.n0 + s.y0;
            str_off = s.str_off;
            row_len = s.row_len;
        }
        // seq1 codes of the segment, 2 bits each, loaded back to back (independent loads)
        const unsigned long long xcodes = (unsigned long long)load_codes16(xs, m) | ((unsigned long long)load_codes16(xs + 16, m - 16) << 32);
        const unsigned long long ycodes = (unsigned long long)load_codes16(ys, n) | ((unsigned long long)load_codes16(ys + 16, n - 16) << 32);
        // row 0 and the per-column score selectors
        Tc[0] = 0 | tagM;
        Uc[0] = gmo | tagM;                                       // X^[1][0] comes from M[0][0]
        for (int j = 1; j <= n; ++j) {
            selc[(j - 1) * TH] = (uint16_t)lut_sel((int)(ycodes >> (2 * (j - 1))) & 3);
            const int y0 = gmo | tagY;                            // Y^[0][j] = -(go + (j-1) ge) + ge j
            Tc[j * TH] = y0;
            int u = __viaddmax_s32(NEG | tagM, gmo, NEG | tagX);
            if (XY) u = max(u, y0 + gmo);
            Uc[j * TH] = u;
        }
        int mmax = m, nmax = n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
            nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        }
        // columns are processed four at a time (one traceback word); columns past a lane's own
        // n, and rows past its m, compute garbage into the lane's private shared-memory columns
        const int ngrp = (nmax + 3) >> 2;
        int eM = 0, eX = 0, eY = 0;
        const int xi0 = gmo | tagX;                               // X^[i][0] = -(go + (i-1) ge) + ge i
        uint32_t *tbrow = tbw;
        for (int i = 1; i <= mmax; ++i) {
            const int xc = (int)(xcodes >> (2 * (i - 1))) & 3;
            const uint32_t slo = s_lut[xc], shi = s_lut[4 + xc];
            int Ml = NEG | tagM, Yl = NEG | tagY, Xl = xi0;
            int diag = Tc[0];
            Tc[0] = xi0;                                         // max(M, X, Y)[i][0]
            Uc[0] = xi0;                                         // X candidate of row i+1, column 0
            const bool last_row = (i == m);
            int *Tp = Tc + TH, *Up = Uc + TH;
            const uint16_t *sp = selc;
            for (int g = 0; g < ngrp; ++g) {
                uint32_t word = 0;
                const int capg = last_row ? n - 4 * g : -1;       // terminal cell is column capg of this group
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    const int Tj = Tp[u4 * TH], Uj = Up[u4 * TH];
                    const uint32_t sel = sp[u4 * TH];
                    const int Mn = strip(diag) + lut_score(slo, shi, sel) + ge2;
                    const int Xn = strip(Uj) | tagX;
                    int v = __viaddmax_s32(Ml, gmo, Yl);
                    if (XY) v = max(v, Xl + gmo);
                    const int Yn = strip(v) | tagY;
                    uint32_t d = ((uint32_t)diag & 3u) | ((uint32_t)Uj & ~3u);
                    d = (d & 15u) | ((uint32_t)v & ~15u);
                    word |= (d & 0xFFu) << (8 * u4);
                    int un = __viaddmax_s32(Mn, gmo, Xn);
                    if (XY) un = max(un, Yn + gmo);
                    Tp[u4 * TH] = __vimax3_s32(Mn, Xn, Yn);
                    Up[u4 * TH] = un;
                    if (capg == u4 + 1) { eM = Mn; eX = Xn; eY = Yn; }
                    diag = Tj; Ml = Mn; Yl = Yn; Xl = Xn;
                }
                tbrow[g * TH] = word;
                Tp += 4 * TH; Up += 4 * TH; sp += 4 * TH;
            }
            tbrow += wpr * TH;
        }
        // traceback: every lane walks its own segment; rows are written reversed
        if (have) {
            int st, score;
            pick_end(eM, eX, eY, prm, st, score);
            const uint32_t pmap = ((uint32_t)prio_to_state(0, prm)) | ((uint32_t)prio_to_state(1, prm) << 2) |
                                  ((uint32_t)prio_to_state(2, prm) << 4);
            int i = m, j = n, len = 0;
            uint8_t *ox = strbuf + str_off, *oy = ox + row_len;
            while (i > 0 || j > 0) {
                if (i == 0) st = 2;
                else if (j == 0) st = 1;
                uint32_t d = 0;
                if (i > 0 && j > 0) d = tbw[((size_t)(i - 1) * wpr + ((j - 1) >> 2)) * TH] >> (8 * ((j - 1) & 3));
                ox[len] = (uint8_t)(st == 2 ? (uint32_t)'-' : code_char((uint32_t)(xcodes >> (2 * (i - 1))) & 3u));
                oy[len] = (uint8_t)(st == 1 ? (uint32_t)'-' : code_char((uint32_t)(ycodes >> (2 * (j - 1))) & 3u));
                const uint32_t tag = (d >> (2 * st)) & 3u;
                i -= (st != 2); j -= (st != 1);
                st = (int)((pmap >> (2 * tag)) & 3u);
                ++len;
            }
            slots[sid].score = score - prm.ge_raw * (m + n);
            slots[sid].len = len;
        }
    }
  }
}

// Register-resident variant for the smallest classes (bound <= 4 * NG <= 16 columns): the
// previous row (T, U) and the per-column score selectors live in registers -- the column
// loop is fully unrolled, so all indexing is static -- and only the traceback words go to
// shared memory.  Saves the five shared-memory accesses per cell of the generic kernel.
// ENDT: the terminal cell is chosen with the same state priorities as the predecessors (end_order ==
// pred_order, the default convention): its winner and score are then max(M, X, Y)[m][n] itself, which
// the sweep keeps anyway (T[n]).  The row of T is parked in shared memory when a lane passes its own
// last row (stores on the idle LSU pipe) instead of picking M, X, Y out of every cell with a compare
// and three selects on the ALU pipe -- a quarter of the sweep's ALU instructions.
template <int NG, bool XY, bool ENDT>
__global__ void __launch_bounds__(TINY_THREADS)
dp_thread_reg_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, Slot *__restrict__ slots,
                     const int32_t *__restrict__ list, const unsigned int *__restrict__ count_dev, DpParams prm,
                     uint8_t *__restrict__ strbuf) {
    extern __shared__ int smem_i[];
    constexpr int TH = TINY_THREADS, NB = 4 * NG;
    const int tid = threadIdx.x;
    const int gmo = prm.ge - prm.go, ge2 = 2 * prm.ge, tagM = prm.tagM, tagX = prm.tagX, tagY = prm.tagY;
    const int NEG = NEG_V * SCALE;
    __shared__ uint32_t s_lut[8];
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { s_lut[c] = prm.lut_lo[c]; s_lut[4 + c] = prm.lut_hi[c]; }
    }
    __syncthreads();
    uint32_t *tbw = reinterpret_cast<uint32_t *>(smem_i) + tid;       // [row][word][thread]
    int *trow = smem_i + NB * NG * TH + tid;                          // ENDT: [column][thread], T of the lane's last row
    const unsigned int count = *count_dev;
    for (unsigned int base = blockIdx.x * TH; base < count; base += gridDim.x * TH) {
        const unsigned int item = base + tid;
        const bool have = item < count;
        int m = 0, n = 0, sid = 0;
        const uint8_t *xs = seqs, *ys = seqs;
        int64_t str_off = 0, row_len = 0;
        if (have) {
            sid = list[item];
            const Slot s = slots[sid];
            m = s.m; n = s.n;
            xs = seqs + s.tb_off + s.x0;                          // tiny classes: tb_off = pd.off0, n0 = pd.off1 - pd.off0
            ys = seqs + s.tb_off + s.n0 + s.y0;
            str_off = s.str_off;
            row_len = s.row_len;
        }
        const uint32_t xcodes = load_codes16(xs, m), ycodes = load_codes16(ys, n);
        int T[NB + 1], U[NB + 1];
        uint32_t sel[NB];
        T[0] = 0 | tagM;
        U[0] = gmo | tagM;                                        // X^[1][0] comes from M[0][0]
        {
            const int y0 = gmo | tagY;                            // Y^[0][j] = -(go + (j-1) ge) + ge j
            int u = __viaddmax_s32(NEG | tagM, gmo, NEG | tagX);
            if (XY) u = max(u, y0 + gmo);
#pragma unroll
            for (int j = 1; j <= NB; ++j) {
                sel[j - 1] = lut_sel((int)((ycodes >> (2 * (j - 1))) & 3u));
                T[j] = y0;
                U[j] = u;
            }
        }
        int mmax = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
        int eM = 0, eX = 0, eY = 0;
        const int xi0 = gmo | tagX;                               // X^[i][0] = -(go + (i-1) ge) + ge i
        uint32_t *tbrow = tbw;
        for (int i = 1; i <= mmax; ++i) {
            const int xc = (int)(xcodes >> (2 * (i - 1))) & 3;
            const uint32_t slo = s_lut[xc], shi = s_lut[4 + xc];
            int Ml = NEG | tagM, Yl = NEG | tagY, Xl = xi0;
            int diag = T[0];
            T[0] = xi0;                                          // max(M, X, Y)[i][0]
            U[0] = xi0;                                          // X candidate of row i+1, column 0
            const int ncap = (i == m) ? n : -1;                  // terminal cell: column ncap of this row
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                uint32_t word = 0;
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    const int j = 4 * g + u4 + 1;
                    const int Tj = T[j], Uj = U[j];
                    const int Mn = strip(diag) + lut_score(slo, shi, sel[j - 1]) + ge2;
                    const int Xn = strip(Uj) | tagX;
                    int v = __viaddmax_s32(Ml, gmo, Yl);
                    if (XY) v = max(v, Xl + gmo);
                    const int Yn = strip(v) | tagY;
                    uint32_t d = ((uint32_t)diag & 3u) | ((uint32_t)Uj & ~3u);
                    d = (d & 15u) | ((uint32_t)v & ~15u);
                    word |= (d & 0xFFu) << (8 * u4);
                    int un = __viaddmax_s32(Mn, gmo, Xn);
                    if (XY) un = max(un, Yn + gmo);
                    T[j] = __vimax3_s32(Mn, Xn, Yn);
                    U[j] = un;
                    if (!ENDT) { if (ncap == j) { eM = Mn; eX = Xn; eY = Yn; } }
                    diag = Tj; Ml = Mn; Yl = Yn; Xl = Xn;
                }
                tbrow[g * TH] = word;
            }
            tbrow += NG * TH;
            if (ENDT && i == m) {
#pragma unroll
                for (int j = 1; j <= NB; ++j) trow[(j - 1) * TH] = T[j];
            }
        }
        if (have) {
            int st, score;
            if (ENDT) {
                const int tn = trow[(n - 1) * TH];               // max(M, X, Y)[m][n], the winner's priority in its low bits
                st = prio_to_state(tn & 3, prm);
                score = strip(tn) >> 6;
            } else {
                pick_end(eM, eX, eY, prm, st, score);
            }
            const uint32_t pmap = ((uint32_t)prio_to_state(0, prm)) | ((uint32_t)prio_to_state(1, prm) << 2) |
                                  ((uint32_t)prio_to_state(2, prm) << 4);
            int i = m, j = n, len = 0;
            uint8_t *ox = strbuf + str_off, *oy = ox + row_len;
            while (i > 0 || j > 0) {
                if (i == 0) st = 2;
                else if (j == 0) st = 1;
                uint32_t d = 0;
                if (i > 0 && j > 0) d = tbw[((i - 1) * NG + ((j - 1) >> 2)) * TH] >> (8 * ((j - 1) & 3));
                ox[len] = (uint8_t)(st == 2 ? (uint32_t)'-' : code_char((xcodes >> (2 * (i - 1))) & 3u));
                oy[len] = (uint8_t)(st == 1 ? (uint32_t)'-' : code_char((ycodes >> (2 * (j - 1))) & 3u));
                const uint32_t tag = (d >> (2 * st)) & 3u;
                i -= (st != 2); j -= (st != 1);
                st = (int)((pmap >> (2 * tag)) & 3u);
                ++len;
            }
            slots[sid].score = score - prm.ge_raw * (m + n);
            slots[sid].len = len;
        }
    }
}

template <int NG>
static void launch_dp_thread_reg_t(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                                   const unsigned int *count_dev, DpParams prm, uint8_t *strbuf, int grid, cudaStream_t st) {
    const size_t smem = (size_t)(4 * NG) * (NG + 1) * 4 * TINY_THREADS;      // traceback words + one row of T per thread
    const bool endt = prm.endM * TAGMUL == prm.tagM && prm.endX * TAGMUL == prm.tagX && prm.endY * TAGMUL == prm.tagY;
#define DBGA_REG_LAUNCH(XYF, EF)                                                                                   \
    do {                                                                                                           \
        allow_max_dyn_smem(dp_thread_reg_kernel<NG, XYF, EF>);                                                     \
        dp_thread_reg_kernel<NG, XYF, EF><<<grid, TINY_THREADS, smem, st>>>(seqs, pairs, slots, list, count_dev, prm, strbuf); \
    } while (0)
    if (prm.xy) { if (endt) DBGA_REG_LAUNCH(true, true); else DBGA_REG_LAUNCH(true, false); }
    else { if (endt) DBGA_REG_LAUNCH(false, true); else DBGA_REG_LAUNCH(false, false); }
#undef DBGA_REG_LAUNCH
}

static size_t dp_thread_smem(int bound, bool tb_in_smem) {
    const int b4 = (bound + 3) & ~3;                   // columns are processed in groups of four
    size_t bytes = (size_t)(2 * (b4 + 1) * 4 + b4 * 2) * TINY_THREADS;
    if (tb_in_smem) bytes += (size_t)b4 * (b4 / 4) * 4 * TINY_THREADS;      // b4 rows x b4/4 words per thread
    return bytes;
}
static int dp_thread_grid(int num_sms) { return num_sms * 8; }
size_t dp_thread_scratch_bytes(int bound, int num_sms) {
    return (size_t)dp_thread_grid(num_sms) * bound * ((bound + 3) / 4) * TINY_THREADS * 4;
}

cudaError_t launch_dp_thread(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *lists,
                             int64_t list_cap, const unsigned int *counts_dev, int c0, int c1, TinyBounds tb,
                             DpParams prm, uint8_t *strbuf, uint32_t *tb_scratch, int num_sms, cudaStream_t st) {
    if (c1 == c0 + 1 && tb.b[c0] <= 16) {        // one small class: previous row in registers
        const int grid = dp_thread_grid(num_sms);
        const int32_t *list = lists + (int64_t)c0 * list_cap;
        if (tb.b[c0] <= 12) launch_dp_thread_reg_t<3>(seqs, pairs, slots, list, counts_dev + c0, prm, strbuf, grid, st);
        else launch_dp_thread_reg_t<4>(seqs, pairs, slots, list, counts_dev + c0, prm, strbuf, grid, st);
        return cudaGetLastError();
    }
    // traceback in shared memory when that leaves at least four CTAs per SM
    const bool tbs = dp_thread_smem(tb.b[c1 - 1], true) <= 56 * 1024;
    const size_t smem = dp_thread_smem(tb.b[c1 - 1], tbs);
    const int grid = dp_thread_grid(num_sms);
    if (prm.xy) {
        allow_max_dyn_smem(dp_thread_kernel<true>);
        dp_thread_kernel<true><<<grid, TINY_THREADS, smem, st>>>(seqs, pairs, slots, lists, list_cap, counts_dev, c0, c1, tb, prm, strbuf, tb_scratch, tbs ? 1 : 0);
    } else {
        allow_max_dyn_smem(dp_thread_kernel<false>);
        dp_thread_kernel<false><<<grid, TINY_THREADS, smem, st>>>(seqs, pairs, slots, lists, list_cap, counts_dev, c0, c1, tb, prm, strbuf, tb_scratch, tbs ? 1 : 0);
    }
    return cudaGetLastError();
}

// ======================================================================= wavefront kernel
struct WfShared {
    int endM, endX, endY, have_end;
};

template <int W, int R, bool XY, int CL>
__global__ void __launch_bounds__(W * 32)
dp_wf_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, Slot *__restrict__ slots,
             const int32_t *__restrict__ list, const unsigned int *__restrict__ count_dev, DpParams prm,
             uint8_t *__restrict__ strbuf, uint8_t *__restrict__ tb_pool, int4 *__restrict__ bnd_pool,
             int64_t bnd_stride, unsigned int cnt_lo, unsigned int cnt_hi) {
    constexpr int D = WF_D, C = WF_C, RING = WF_RING;
    constexpr int GW = W * CL;                                // warps working on one segment
    constexpr int ROWS_PASS = GW * 32 * R;
    __shared__ int4 ring[W + 1][RING];
    __shared__ WfShared sh;
    __shared__ __align__(16) uint8_t win[WALK_WIN_ROWS][64];      // traceback window
    __shared__ uint32_t s_lut[8];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { s_lut[c] = prm.lut_lo[c]; s_lut[4 + c] = prm.lut_hi[c]; }
    }
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // CL > 1: a thread-block cluster of CL CTAs shares one segment; the boundary row of a CTA's
    // last warp goes straight into the next CTA's ring[0] through distributed shared memory.
    int crank = 0;
    int4 *ring_next = nullptr;
    if constexpr (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        crank = (int)cluster.block_rank();
        if (crank + 1 < CL) ring_next = cluster.map_shared_rank(&ring[0][0], crank + 1);
    }
    const int gw = crank * W + w;
    auto bar = [&]() {
        if constexpr (CL > 1) cg::this_cluster().sync();
        else __syncthreads();
    };
    const unsigned int group = blockIdx.x / CL, n_groups = gridDim.x / CL;
    const unsigned int count = *count_dev;
    if (count < cnt_lo || count > cnt_hi) return;            // the long-segment list goes to one of two kernels
    const int gmo = prm.ge - prm.go, ge2 = 2 * prm.ge, tagM = prm.tagM, tagX = prm.tagX, tagY = prm.tagY;
    const int NEG = NEG_V * SCALE;
    int4 *bndA = bnd_pool + (int64_t)group * 2 * bnd_stride, *bndB = bndA + bnd_stride;
    int4 *endv = bndA + (bnd_stride - 1);                     // CL > 1: end-cell values cross CTAs here

    for (unsigned int item = group; item < count; item += n_groups) {
        const int sid = list[item];
        const Slot s = slots[sid];
        const PairDesc pd = pairs[s.pair];
        const int m = s.m, n = s.n;
        if constexpr (W == 1 && CL == 1) {
            if (wf16_pick(prm.h, s.cls, m, n)) continue;     // a packed 16-bit kernel of this class takes it
        }
        const uint8_t *xs = seqs + pd.off0 + s.x0, *ys = seqs + pd.off1 + s.y0;
        uint8_t *tb = tb_pool + s.tb_off;
        const int passes = (m + ROWS_PASS - 1) / ROWS_PASS;
        const int64_t tb_rows = (int64_t)n + 32;           // step rows per (pass, warp) block
        if (tid == 0) sh.have_end = 0;
        bar();

        for (int pass = 0; pass < passes; ++pass) {
            const int base = pass * ROWS_PASS;
            const int4 *bin = (pass & 1) ? bndB : bndA;
            int4 *bout = (pass & 1) ? bndA : bndB;
            const bool need_out = pass + 1 < passes;
            int Mr[R], Xr[R], Yr[R];
            uint32_t slo[R], shi[R];
            const int row0 = base + (gw * 32 + lane) * R;    // 0-based index of this lane's first row
            // warps whose rows all lie past m sit the pass out (they only keep the barriers)
            const int act = min(GW, (m - base + 32 * R - 1) / (32 * R));
            const bool active = gw < act;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = row0 + r + 1;
                const int xc = (i <= m) ? base_code(xs[i - 1]) : 0;
                slo[r] = s_lut[xc]; shi[r] = s_lut[4 + xc];
                Mr[r] = NEG | tagM;
                Xr[r] = gmo | tagX;                          // X^[i][0]
                Yr[r] = NEG | tagY;
            }
            int dgM = 0, dgX = 0, dgY = 0, my_yc = 0;
            const int own = m - 1 - base;                    // row m inside this pass?
            const bool i_own = own >= 0 && own < ROWS_PASS && (own / R) == (gw * 32 + lane);
            const int own_r = own >= 0 ? own % R : 0;
            uint8_t *tbw = tb + ((int64_t)(pass * GW + gw) * tb_rows) * (32 * R) + lane * R;
            const int off_last = (act - 1) * D + 31;
            const bool last_warp = gw == act - 1;
            int4 *ring_out = (w + 1 < W || ring_next == nullptr) ? &ring[w + 1][0] : ring_next;
            const int S = n + 1 + off_last;                  // steps in this pass
            const int NB = (S + C - 1) / C;

            auto fill = [&](int c) {                         // ring[0] entry for column c
                if (c < 0 || c > n) return;
                int4 e;
                if (pass == 0) {
                    e.x = (c == 0 ? 0 : NEG) | tagM; e.y = NEG | tagX;
                    e.z = (c == 0 ? NEG : gmo) | tagY;       // Y^[0][c]
                } else {
                    e = bin[c];
                }
                e.w = c >= 1 ? base_code(ys[c - 1]) : 0;
                ring[0][c & (RING - 1)] = e;
            };
            if (gw == 0) fill(lane);
            bar();

            for (int blk = 0; blk < NB; ++blk) {
                const int s0 = blk * C;
                if (gw == 0) fill(s0 + C + lane);            // next chunk, visible after the barrier
                if (need_out && last_warp) {                 // drain what lane 31 produced last block
                    const int c = s0 - C - off_last + lane;
                    if (c >= 0 && c <= n) bout[c] = ring[W][c & (RING - 1)];
                }
#pragma unroll 1
                for (int st = 0; active && st < C; ++st) {
                    const int sidx = s0 + st;
                    const int j = sidx - gw * D - lane;      // column (0 = init column)
                    int uM = __shfl_up_sync(0xffffffffu, Mr[R - 1], 1);
                    int uX = __shfl_up_sync(0xffffffffu, Xr[R - 1], 1);
                    int uY = __shfl_up_sync(0xffffffffu, Yr[R - 1], 1);
                    int yc = __shfl_up_sync(0xffffffffu, my_yc, 1);
                    if (lane == 0 && j >= 0 && j <= n) {
                        const int4 e = ring[w][j & (RING - 1)];
                        uM = e.x; uX = e.y; uY = e.z; yc = e.w;
                    }
                    if (j >= 1 && j <= n) {
                        const uint32_t sel = lut_sel(yc);
                        int dM = dgM, dX = dgX, dY = dgY, aM = uM, aX = uX, aY = uY;
                        uint32_t wv[(R + 3) / 4] = {};
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int oM = Mr[r], oX = Xr[r], oY = Yr[r];
                            const uint32_t d = dp_cell<XY>(Mr[r], Xr[r], Yr[r], dM, dX, dY, aM, aX, aY,
                                                           lut_score(slo[r], shi[r], sel), gmo, ge2, tagX, tagY);
                            wv[r / 4] |= d << (8 * (r % 4));
                            dM = oM; dX = oX; dY = oY; aM = Mr[r]; aX = Xr[r]; aY = Yr[r];
                        }
                        uint8_t *tbp = tbw + (int64_t)(j + lane) * (32 * R);
                        if constexpr (R == 16) *reinterpret_cast<uint4 *>(tbp) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                        else if constexpr (R == 8) *reinterpret_cast<uint2 *>(tbp) = make_uint2(wv[0], wv[1]);
                        else if constexpr (R == 4) *reinterpret_cast<uint32_t *>(tbp) = wv[0];
                        else *reinterpret_cast<uint16_t *>(tbp) = (uint16_t)wv[0];
                        my_yc = yc;
                        if (j == n && i_own) {
#pragma unroll
                            for (int r = 0; r < R; ++r)
                                if (r == own_r) {
                                    if constexpr (CL > 1) { *endv = make_int4(Mr[r], Xr[r], Yr[r], 1); __threadfence(); }
                                    else { sh.endM = Mr[r]; sh.endX = Xr[r]; sh.endY = Yr[r]; sh.have_end = 1; }
                                }
                        }
                    }
                    dgM = uM; dgX = uX; dgY = uY;
                    if (lane == 31 && j >= 0 && j <= n && (!last_warp || need_out))
                        ring_out[j & (RING - 1)] = make_int4(Mr[R - 1], Xr[R - 1], Yr[R - 1], my_yc);
                }
                bar();
            }
            if (need_out && last_warp) {                     // final drain
                for (int c = NB * C - C - off_last + lane; c <= n; c += 32)
                    if (c >= 0) bout[c] = ring[W][c & (RING - 1)];
            }
            if constexpr (CL > 1) __threadfence();
            bar();
        }
        // ---------------------------------------------------------------- traceback (warp 0)
        if (gw == 0) {
            int st = 0, score = 0;
            if constexpr (CL > 1) {
                const int4 e = __ldcg(endv);
                pick_end(e.x, e.y, e.z, prm, st, score);
            } else {
                pick_end(sh.endM, sh.endX, sh.endY, prm, st, score);
            }
            const int64_t row_len = (pd.off1 - pd.off0) + pd.n1;
            uint8_t *ox = strbuf + s.str_off, *oy = ox + row_len;
            // state reached through a traceback tag: packed table prio -> state (0 = M, 1 = X, 2 = Y)
            const uint32_t pmap = ((uint32_t)prio_to_state(0, prm)) | ((uint32_t)prio_to_state(1, prm) << 2) |
                                  ((uint32_t)prio_to_state(2, prm) << 4);
            constexpr int RSH = R == 2 ? 1 : (R == 4 ? 2 : (R == 8 ? 3 : 4));
            static_assert((1 << RSH) == R, "rows per lane must be 2, 4, 8 or 16");
            const int len = wf_traceback(win, tb, tb_rows, xs, ys, m, n, st, pmap, ox, oy, lane, WalkGeom{5 + RSH, RSH, -1, 0u});
            if (lane == 0) { slots[sid].score = score - prm.ge_raw * (m + n); slots[sid].len = len; }
        }
        __syncthreads();
    }
}

template <int W, int R, int CL = 1>
static void launch_wf_t(int grid, const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                        const unsigned int *count_dev, DpParams prm, uint8_t *strbuf, uint8_t *tb, int4 *bnd,
                        int64_t bnd_stride, cudaStream_t st, unsigned int cnt_lo = 0, unsigned int cnt_hi = 0xffffffffu) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid * CL);       // grid counts segments in flight (clusters when CL > 1)
    cfg.blockDim = dim3(W * 32);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (CL > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    if (prm.xy) cudaLaunchKernelEx(&cfg, dp_wf_kernel<W, R, true, CL>, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, bnd, bnd_stride, cnt_lo, cnt_hi);
    else cudaLaunchKernelEx(&cfg, dp_wf_kernel<W, R, false, CL>, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, bnd, bnd_stride, cnt_lo, cnt_hi);
}

// cls: CLS_WARP2 / 4 / 8 / 16 (one warp per segment)
cudaError_t launch_dp_wavefront(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                                const unsigned int *count_dev, int grid, int cls, DpParams prm,
                                uint8_t *strbuf, uint8_t *tb, int4 *bnd, int64_t bnd_stride, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    if (cls == CLS_WARP2) launch_wf_t<1, 2>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, bnd, bnd_stride, st);
    else if (cls == CLS_WARP4) launch_wf_t<1, 4>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, bnd, bnd_stride, st);
    else if (cls == CLS_WARP16) launch_wf_t<1, 16>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, bnd, bnd_stride, st);
    else launch_wf_t<1, 8>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, bnd, bnd_stride, st);
    return cudaGetLastError();
}

// CLS_CTA, the long segments.  Up to `few` of them: a cluster of WF_CLUSTER CTAs each (latency:
// a lone big segment is all there is to run).  More than that: one CTA of 8 warps x 8 rows per
// lane each (throughput: less pipeline fill, all SMs busy).  The count lives on the device, so
// both kernels are launched and one of them returns at once.
cudaError_t launch_dp_long(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                           const unsigned int *count_dev, int grid_clusters, int grid_ctas, unsigned int few, DpParams prm,
                           uint8_t *strbuf, uint8_t *tb, int4 *bnd, int64_t bnd_stride, cudaStream_t st) {
    if (grid_clusters <= 0 || grid_ctas <= 0) return cudaSuccess;
    launch_wf_t<WF_CL_WARPS, WF_CL_R, WF_CLUSTER>(grid_clusters, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, bnd, bnd_stride, st, 1u, few);
    launch_wf_t<WF_CTA_WARPS, 8, 1>(grid_ctas, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, bnd, bnd_stride, st, few + 1u, 0xffffffffu);
    return cudaGetLastError();
}

}  // namespace dbga
