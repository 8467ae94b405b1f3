// Stage 3, packed 16-bit wavefront: the same affine-gap (Gotoh) global alignment as
// stage3_dp.cu (replaces cogent3.align.global_pairwise as called from dna_global_aln,
// reference src/dbga/utils.py:157-192, call site :185), two cells per 32-bit register.
//
// A warp is 64 half-lanes: lane l owns 2R consecutive matrix rows, the first R in the low
// int16 halves of its registers, the next R in the high halves.  At step t the low half is
// at column t - 2l and the high half one column behind it (t - 2l - 1), so the high half's
// row above is the low half's last row one step earlier (same thread, no shuffle) and the
// low half's row above is the previous lane's high half one step earlier (one shuffle +
// PRMT per state).  All arithmetic is the sm_100a packed forms: VIMNMX3.S16x2,
// VIADDMNMX.S16x2, VIADD.16x2.
//
// Encoding per half: 4 * (V - off) + priority(state), V = H + ge * (i + j) (gap extension is
// free, as in the 32-bit kernels); the integer max breaks ties in the configured first-wins
// order and leaves the winner's tag in the low two bits, which become the traceback byte
// (bits 0-1: predecessor of M, 2-3: of X, 4-5: of Y).  Substitution scores are int8 LUT
// bytes picked and sign-extended for both halves by one PRMT.  `off` places the valid range
// asymmetrically: valid values are >= about -(|s| + 4 go), "minus infinity" sits in a band
// below that which a few additions cannot leave, and the rest of the int16 range is head
// room for per_step * min(m, n) (wf16_ok decides per segment; the 32-bit kernels take the
// rest).  Column 0 and the column before it are computed like any other from minus-infinity
// registers, which yields exactly X^[i][0] = ge - go and minus infinity for M and Y, so a
// half-lane needs no start-up masking.  One pass only (m <= 64 R); traceback bytes use the
// 32-bit kernels' step-major layout (row = step, 64 R bytes per row) and the shared walker.
#include <cstdlib>

#include "dp_walk.cuh"
#include "kernels.cuh"

namespace dbga {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t dup16(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }

// The per-lane constants of the packed cell.
struct Wf16Consts {
    uint32_t G, tagX, tagY, cm1, c4, c16;
};

// One step of one lane: R packed rows at the lane's current column(s).  (dM,dX,dY): the row above
// at the previous column, (aM,aX,aY): the row above at this column; Mr/Xr/Yr: the lane's rows at
// the previous column in, at this column out.  D[r]: raw traceback fields of row r's two cells
// (tt & 3, (u & 3) - pX, (v & 3) - pY per half as et + 4 eu + 16 ev, exact 32-bit arithmetic:
// the borrows between fields cancel once tbk is added to every byte).
// ALU pipe (the bottleneck: half issue rate): three maxes, three bit strips, the score PRMT and
// the packed add per row.  Everything else is IMAD on the FMA pipe, with multipliers the
// compiler cannot see through (a literal * 4 would become LEA / SHL on the ALU pipe).
template <int R, bool XY>
__device__ __forceinline__ void wf16_rows(uint32_t (&Mr)[R], uint32_t (&Xr)[R], uint32_t (&Yr)[R], uint32_t (&D)[R],
                                          const uint32_t (&lutA)[R], const uint32_t (&lutB)[R], uint32_t sel,
                                          uint32_t dM, uint32_t dX, uint32_t dY, uint32_t aM, uint32_t aX, uint32_t aY,
                                          const Wf16Consts &c) {
    constexpr uint32_t KS = 0xFFFCFFFCu;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t oM = Mr[r], oX = Xr[r], oY = Yr[r];
        const uint32_t tt = __vimax3_s16x2(dM, dX, dY);
        uint32_t u = __viaddmax_s16x2(aM, c.G, aX);
        uint32_t v = __viaddmax_s16x2(oM, c.G, oY);
        if (XY) {
            u = __viaddmax_s16x2(aY, c.G, u);
            v = __viaddmax_s16x2(oX, c.G, v);
        }
        const uint32_t ts = tt & KS;
        Xr[r] = (u & KS) | c.tagX;
        Yr[r] = (v & KS) | c.tagY;
        Mr[r] = __vadd2(ts, prmt(lutA[r], lutB[r], sel));       // LUT bytes carry M's tag
        const uint32_t et = imad(ts, c.cm1, tt), eu = imad(Xr[r], c.cm1, u), ev = imad(Yr[r], c.cm1, v);
        D[r] = imad(ev, c.c16, imad(eu, c.c4, et));
        dM = oM; dX = oX; dY = oY; aM = Mr[r]; aX = Xr[r]; aY = Yr[r];
    }
}

template <int R, bool XY>
__global__ void __launch_bounds__(32, R == 16 ? 12 : 20)
dp_wf16_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, Slot *__restrict__ slots,
               const int32_t *__restrict__ list, const unsigned int *__restrict__ count_dev, DpParams prm,
               uint8_t *__restrict__ strbuf, uint8_t *__restrict__ tb_pool) {
    __shared__ uint32_t s_lut[4];
    const int lane = threadIdx.x;
    const Dp16 &h = prm.h;
    if (lane < 4) s_lut[lane] = h.lut[lane];
    __syncwarp();
    constexpr uint32_t FULL = 0xffffffffu;
    // Compact traceback (no X <-> Y, R >= 4): FOUR bits per cell.  X can only come from M or X and Y from M
    // or Y, so (u & 3) - pX is 0 or pM - pX and the field "X came from M" is that difference times
    // 4 / (pM - pX) -- an integer multiplier (+-4 or +-2) in the same IMAD that the byte form runs with 4;
    // likewise 8 / (pM - pY) for Y.  A cell is then (t & 3) + 4 [X from M] + 8 [Y from M], no bias, two
    // matrix rows per byte, half the bytes per step (dp_walk.cuh: WalkGeom::nib_sh).
    constexpr bool NIB = !XY && R >= 4;
    const uint32_t c256 = h.mul[3], c16 = h.mul[2];
    const Wf16Consts cc{h.G, h.tagX, h.tagY, h.mul[0], NIB ? h.nmul[0] : h.mul[1], NIB ? h.nmul[1] : h.mul[2]};
    const uint32_t K1 = h.tbk * 0x00010001u;
    constexpr int ROWW = NIB ? 32 * R : 64 * R;               // bytes per step row
    const uint32_t tagM = h.tagM, tagX = h.tagX, tagY = h.tagY;
    const uint32_t negM = dup16(h.qneg) | tagM, negX = dup16(h.qneg) | tagX, negY = dup16(h.qneg) | tagY;
    // row 0 as seen by lane 0's low half (arrives where the shuffle would put it: high half)
    const uint32_t r0M0 = (dup16(h.q0) | tagM) << 16, r0Mn = negM << 16, r0X = negX << 16,
                   r0Y = (dup16(h.qy0) | tagY) << 16, r0Yn = negY << 16;
    const unsigned int count = *count_dev;

    for (unsigned int item = blockIdx.x; item < count; item += gridDim.x) {
        const int sid = list[item];
        const Slot s = slots[sid];
        const int m = s.m, n = s.n;
        if (wf16_pick(h, s.cls, m, n) != R) continue;        // another kernel of this class takes it
        const PairDesc pd = pairs[s.pair];
        const uint8_t *xs = seqs + pd.off0 + s.x0, *ys = seqs + pd.off1 + s.y0;
        uint8_t *tb = tb_pool + s.tb_off;
        uint32_t lutA[R], lutB[R], Mr[R], Xr[R], Yr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int iA = lane * 2 * R + r, iB = iA + R;    // 0-based matrix rows
            lutA[r] = s_lut[iA < m ? base_code(xs[iA]) : 0];
            lutB[r] = s_lut[iB < m ? base_code(xs[iB]) : 0];
            Mr[r] = negM; Xr[r] = negX; Yr[r] = negY;
        }
        uint32_t dgM = negM, dgX = negX, dgY = negY;
        // the terminal cell: half-lane ve, row re of it, reached at step te
        const int ve = (m - 1) / R, re = (m - 1) % R, te = n + ve, le = ve >> 1;
        // Steps run in whole chunks of 16 (one word of seq2 codes for lane 0) that end exactly at
        // step te, so the terminal cell is still in its registers after the loop; the first chunk
        // may start at a negative step.  No lane is masked.  Before its low half reaches column 0 a
        // lane sees the initial selector 0x8888 (score 0 or -1: the sign of a LUT byte), and
        // minus infinity in, minus infinity out: its registers only drift down by at most 4 per
        // step.  After its last column it computes garbage nobody reads.  Traceback rows are
        // offset by 16 so that negative steps have somewhere to write.
        const int nch = (te + 16) >> 4, t_start = te + 1 - 16 * nch;
        uint32_t yprev = 0, sel_cur = 0x8888u, sel_prev = 0x8888u;
        uint8_t *tbw = tb + 16 * ROWW + lane * (NIB ? R : 2 * R);
        // code of column t = bits 2 (t - t0) of the chunk's word; t0 <= 0 only in the first chunk,
        // whose word is built from ys[0..15] so that nothing before ys is read
        uint32_t wy_next = (uint32_t)((unsigned long long)load_codes16(ys, n) << (2 * (1 - t_start)));
        for (int t0 = t_start; t0 <= te; t0 += 16) {
            const uint32_t wy = wy_next;
            wy_next = load_codes16(ys + 15 + t0, n - 15 - t0);
#pragma unroll 2
            for (int sidx = 0; sidx < 16; ++sidx) {
                const int t = t0 + sidx;
                // lane 0: selector for columns (t, t - 1); lane l gets lane l-1's selector of two steps ago
                const uint32_t ycur = (wy >> (2 * sidx)) & 3u;
                const uint32_t sel0 = t < 0 ? 0x8888u : 0xC480u + 0x11u * ycur + 0x1100u * yprev;   // t < 0: lane 0 is not running yet either
                yprev = ycur;
                const uint32_t sel_in = __shfl_up_sync(FULL, sel_prev, 1);
                sel_prev = sel_cur;
                sel_cur = lane == 0 ? sel0 : sel_in;
                // the row above: low half <- previous lane's high half, high half <- own low half
                uint32_t sM = __shfl_up_sync(FULL, Mr[R - 1], 1);
                uint32_t sX = __shfl_up_sync(FULL, Xr[R - 1], 1);
                uint32_t sY = __shfl_up_sync(FULL, Yr[R - 1], 1);
                if (lane == 0) { sM = t == 0 ? r0M0 : r0Mn; sX = r0X; sY = t >= 1 ? r0Y : r0Yn; }
                const uint32_t uM = prmt(sM, Mr[R - 1], 0x5432u), uX = prmt(sX, Xr[R - 1], 0x5432u),
                               uY = prmt(sY, Yr[R - 1], 0x5432u);
                uint32_t D[R];
                wf16_rows<R, XY>(Mr, Xr, Yr, D, lutA, lutB, sel_cur, dgM, dgX, dgY, uM, uX, uY, cc);
                // traceback bytes of the lane, per pair of rows (r, r+1): [low r, low r+1, high r, high r+1]
                uint8_t *tbp = tbw + (int64_t)t * ROWW;
                if constexpr (NIB) {
                    // two rows per byte (r, r+1), then the bytes of rows (r, r+1) and (r+2, r+3) of both half-lanes in
                    // one word: [low (r, r+1), low (r+2, r+3), high (r, r+1), high (r+2, r+3)]
                    uint32_t F[R / 4];
#pragma unroll
                    for (int g = 0; g < R / 4; ++g)
                        F[g] = imad(imad(D[4 * g + 3], c16, D[4 * g + 2]), c256, imad(D[4 * g + 1], c16, D[4 * g]));
                    if constexpr (R == 4) {
                        *reinterpret_cast<uint32_t *>(tbp) = F[0];
                    } else if constexpr (R == 8) {
                        *reinterpret_cast<uint2 *>(tbp) = make_uint2(F[0], F[1]);
                    } else {
                        *reinterpret_cast<uint4 *>(tbp) = make_uint4(F[0], F[1], F[2], F[3]);
                    }
                } else if constexpr (R == 1) {
                    *reinterpret_cast<uint16_t *>(tbp) = (uint16_t)prmt(D[0] + K1, 0u, 0x4420u);
                } else {
                    uint32_t E[R / 2];
#pragma unroll
                    for (int g = 0; g < R / 2; ++g) E[g] = imad(D[2 * g + 1], c256, D[2 * g]);   // the walker adds the bias
                    if constexpr (R == 2) {
                        *reinterpret_cast<uint32_t *>(tbp) = E[0];
                    } else if constexpr (R == 4) {
                        *reinterpret_cast<uint2 *>(tbp) = make_uint2(E[0], E[1]);
                    } else if constexpr (R == 8) {
                        *reinterpret_cast<uint4 *>(tbp) = make_uint4(E[0], E[1], E[2], E[3]);
                    } else {
                        static_assert(R == 16, "R must be 1, 2, 4, 8 or 16");
                        *reinterpret_cast<uint4 *>(tbp) = make_uint4(E[0], E[1], E[2], E[3]);
                        *reinterpret_cast<uint4 *>(tbp + 16) = make_uint4(E[4], E[5], E[6], E[7]);
                    }
                }
                dgM = uM; dgX = uX; dgY = uY;
            }
        }
        // ------------------------------------------------------------------ terminal cell
        uint32_t eM = 0, eX = 0, eY = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r == re) { eM = Mr[r]; eX = Xr[r]; eY = Yr[r]; }
        eM = __shfl_sync(FULL, eM, le); eX = __shfl_sync(FULL, eX, le); eY = __shfl_sync(FULL, eY, le);
        const int sh = (ve & 1) ? 16 : 0;
        const int vM = (int)(int16_t)(eM >> sh), vX = (int)(int16_t)(eX >> sh), vY = (int)(int16_t)(eY >> sh);
        // end_order: first wins, strict > (oracle/dbga_oracle.py pick of the terminal state)
        const int cm = (vM & ~3) + prm.endM, cx = (vX & ~3) + prm.endX, cy = (vY & ~3) + prm.endY;
        int best = cm, st = 0;
        if (cx > best) { best = cx; st = 1; }
        if (cy > best) { best = cy; st = 2; }
        const int score = (best >> 2) + h.off - prm.ge_raw * (m + n);
        // the traceback runs in dp_tb16_kernel; the terminal state travels in slot.len
        if (lane == 0) { slots[sid].score = score; slots[sid].len = st; }
    }
}

// ---------------------------------------------------------------------- two segments per warp
// Throughput form for segments of up to 32 R rows: the low halves of a warp's registers carry one
// segment (A), the high halves another (B, the next entry of the class list), both at column
// t - lane.  Same cell, same encoding; no half-lane skew (32 lanes of skew instead of 64, plain
// shuffles for the row above), and the per-step overhead is shared by two segments.  Rows and
// columns run to the larger of the two segments; the smaller one computes cells nobody reads
// (seq2 codes past its end are 0, rows past its end use base code 0: all below the ceiling
// wf16_pick checks for 32 R rows).  Each segment has its own traceback table, in the 32-bit
// kernels' layout (row = step + 16, byte = matrix row).
template <int R, bool XY>
__global__ void __launch_bounds__(32, R == 16 ? 12 : 20)
dp_wf16p_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, Slot *__restrict__ slots,
                const int32_t *__restrict__ list, const unsigned int *__restrict__ count_dev, DpParams prm,
                uint8_t *__restrict__ tb_pool) {
    __shared__ uint32_t s_lut[4];
    const int lane = threadIdx.x;
    const Dp16 &h = prm.h;
    if (lane < 4) s_lut[lane] = h.lut[lane];
    __syncwarp();
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr bool NIB = !XY;                                 // compact traceback: four bits per cell, two matrix rows per byte
    constexpr int ROWW = NIB ? 16 * R : 32 * R;               // bytes per step row of a segment's table
    constexpr int LB = NIB ? R / 2 : R;                       // bytes per lane per step
    const uint32_t c256 = h.mul[3], c16 = h.mul[2], K4 = h.tbk * 0x01010101u;
    const Wf16Consts cc{h.G, h.tagX, h.tagY, h.mul[0], NIB ? h.nmul[0] : h.mul[1], NIB ? h.nmul[1] : h.mul[2]};
    const uint32_t tagM = h.tagM, tagX = h.tagX, tagY = h.tagY;
    const uint32_t negM = dup16(h.qneg) | tagM, negX = dup16(h.qneg) | tagX, negY = dup16(h.qneg) | tagY;
    const uint32_t r0M0 = dup16(h.q0) | tagM, r0Y = dup16(h.qy0) | tagY;
    const unsigned int count = *count_dev;

    for (unsigned int g = blockIdx.x; 2 * g < count; g += gridDim.x) {
        // the two list entries of this warp; an entry another kernel takes leaves its half idle
        int sid[2], m[2], n[2];
        const uint8_t *xs[2], *ys[2];
        uint8_t *tb[2];
        bool have[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const unsigned int item = 2 * g + e;
            have[e] = false; sid[e] = 0; m[e] = 0; n[e] = 0; xs[e] = seqs; ys[e] = seqs; tb[e] = tb_pool;
            if (item < count) {
                sid[e] = list[item];
                const Slot s = slots[sid[e]];
                if (wf16_pick(h, s.cls, s.m, s.n) == WF16_PAIR + R) {
                    const PairDesc pd = pairs[s.pair];
                    have[e] = true; m[e] = s.m; n[e] = s.n;
                    xs[e] = seqs + pd.off0 + s.x0; ys[e] = seqs + pd.off1 + s.y0;
                    tb[e] = tb_pool + s.tb_off + 16 * ROWW + lane * LB;
                }
            }
        }
        if (!have[0] && !have[1]) continue;
        uint32_t lutA[R], lutB[R], Mr[R], Xr[R], Yr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = lane * R + r;                       // 0-based matrix row
            lutA[r] = s_lut[i < m[0] ? base_code(xs[0][i]) : 0];
            lutB[r] = s_lut[i < m[1] ? base_code(xs[1][i]) : 0];
            Mr[r] = negM; Xr[r] = negX; Yr[r] = negY;
        }
        uint32_t dgM = negM, dgX = negX, dgY = negY;
        // a segment's terminal cell is reached at step te = n + (lane of row m)
        int te[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) te[e] = have[e] ? n[e] + (m[e] - 1) / R : -1;
        const int t_end = max(te[0], te[1]);
        const int nch = (t_end + 16) >> 4, t_start = t_end + 1 - 16 * nch;      // whole chunks of 16 steps ending at t_end
        uint32_t sel_cur = 0x8888u;                           // idle lanes: score 0 or -1, minus infinity stays put
        uint32_t wnA = (uint32_t)((unsigned long long)load_codes16(ys[0], n[0]) << (2 * (1 - t_start)));
        uint32_t wnB = (uint32_t)((unsigned long long)load_codes16(ys[1], n[1]) << (2 * (1 - t_start)));
        for (int t0 = t_start; t0 <= t_end; t0 += 16) {
            const uint32_t wA = wnA, wB = wnB;
            wnA = load_codes16(ys[0] + 15 + t0, n[0] - 15 - t0);
            wnB = load_codes16(ys[1] + 15 + t0, n[1] - 15 - t0);
#pragma unroll 2
            for (int sidx = 0; sidx < 16; ++sidx) {
                const int t = t0 + sidx;
                // lane 0: selector for column t of both segments; lane l gets lane l-1's of one step ago
                const uint32_t yA = (wA >> (2 * sidx)) & 3u, yB = (wB >> (2 * sidx)) & 3u;
                const uint32_t sel0 = t < 0 ? 0x8888u : 0xC480u + 0x11u * yA + 0x1100u * yB;
                const uint32_t sel_in = __shfl_up_sync(FULL, sel_cur, 1);
                sel_cur = lane == 0 ? sel0 : sel_in;
                uint32_t uM = __shfl_up_sync(FULL, Mr[R - 1], 1);
                uint32_t uX = __shfl_up_sync(FULL, Xr[R - 1], 1);
                uint32_t uY = __shfl_up_sync(FULL, Yr[R - 1], 1);
                if (lane == 0) { uM = t == 0 ? r0M0 : negM; uX = negX; uY = t >= 1 ? r0Y : negY; }
                uint32_t D[R];
                wf16_rows<R, XY>(Mr, Xr, Yr, D, lutA, lutB, sel_cur, dgM, dgX, dgY, uM, uX, uY, cc);
                dgM = uM; dgX = uX; dgY = uY;
                // traceback bytes: pairs of rows as one word [A r, A r+1, B r, B r+1] plus the bias, then split
                // steps past a segment's own last one (its partner is still running) all land in one spare row
                uint8_t *pa = tb[0] + (int64_t)min(t, te[0] + 1) * ROWW, *pb = tb[1] + (int64_t)min(t, te[1] + 1) * ROWW;
                uint32_t E[R / 2];
#pragma unroll
                for (int q = 0; q < R / 2; ++q) E[q] = NIB ? imad(D[2 * q + 1], c16, D[2 * q]) : imad(D[2 * q + 1], c256, D[2 * q]) + K4;
                if constexpr (NIB) {
                    // E[q]: byte 0 = A's rows (2q, 2q+1), byte 2 = B's rows (2q, 2q+1)
                    if constexpr (R == 2) {
                        if (have[0]) *pa = (uint8_t)E[0];
                        if (have[1]) *pb = (uint8_t)(E[0] >> 16);
                    } else if constexpr (R == 4) {
                        const uint32_t w2 = prmt(E[0], E[1], 0x6240u);            // [A01, A23, B01, B23]
                        if (have[0]) *reinterpret_cast<uint16_t *>(pa) = (uint16_t)w2;
                        if (have[1]) *reinterpret_cast<uint16_t *>(pb) = (uint16_t)(w2 >> 16);
                    } else {
                        static_assert(R == 8, "the paired kernel runs 2, 4 or 8 rows per lane");
                        const uint32_t w01 = prmt(E[0], E[1], 0x6240u), w23 = prmt(E[2], E[3], 0x6240u);
                        if (have[0]) *reinterpret_cast<uint32_t *>(pa) = prmt(w01, w23, 0x5410u);
                        if (have[1]) *reinterpret_cast<uint32_t *>(pb) = prmt(w01, w23, 0x7632u);
                    }
                } else if constexpr (R == 2) {
                    if (have[0]) *reinterpret_cast<uint16_t *>(pa) = (uint16_t)E[0];
                    if (have[1]) *reinterpret_cast<uint16_t *>(pb) = (uint16_t)(E[0] >> 16);
                } else {
                    uint32_t wa[R / 4], wb[R / 4];
#pragma unroll
                    for (int q = 0; q < R / 4; ++q) {
                        wa[q] = prmt(E[2 * q], E[2 * q + 1], 0x5410u);
                        wb[q] = prmt(E[2 * q], E[2 * q + 1], 0x7632u);
                    }
                    if constexpr (R == 4) {
                        if (have[0]) *reinterpret_cast<uint32_t *>(pa) = wa[0];
                        if (have[1]) *reinterpret_cast<uint32_t *>(pb) = wb[0];
                    } else if constexpr (R == 8) {
                        if (have[0]) *reinterpret_cast<uint2 *>(pa) = make_uint2(wa[0], wa[1]);
                        if (have[1]) *reinterpret_cast<uint2 *>(pb) = make_uint2(wb[0], wb[1]);
                    } else {
                        static_assert(R == 16, "R must be 2, 4, 8 or 16");
                        if (have[0]) *reinterpret_cast<uint4 *>(pa) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
                        if (have[1]) *reinterpret_cast<uint4 *>(pb) = make_uint4(wb[0], wb[1], wb[2], wb[3]);
                    }
                }
                // a lane that has just finished a segment's last column leaves its rows' three states in the
                // first rows of that segment's table (free by then: only steps < 0 ever wrote there); the
                // traceback kernel picks the terminal cell out of them
                const int jcol = t - lane;
                if (jcol == n[0] && have[0]) {
                    uint16_t *eb = reinterpret_cast<uint16_t *>(tb[0] - (int64_t)16 * ROWW - lane * LB) + lane * R;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        eb[r] = (uint16_t)Mr[r]; eb[32 * R + r] = (uint16_t)Xr[r]; eb[64 * R + r] = (uint16_t)Yr[r];
                    }
                }
                if (jcol == n[1] && have[1]) {
                    uint16_t *eb = reinterpret_cast<uint16_t *>(tb[1] - (int64_t)16 * ROWW - lane * LB) + lane * R;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        eb[r] = (uint16_t)(Mr[r] >> 16); eb[32 * R + r] = (uint16_t)(Xr[r] >> 16); eb[64 * R + r] = (uint16_t)(Yr[r] >> 16);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------- long segments: row strips
// Segments the one-pass kernels cannot hold (more than 64 R rows, or scores past the int16 ceiling): the
// matrix is cut into strips of 64 R rows, one warp sweeps a strip over all columns exactly like
// dp_wf16_kernel, and the strips of a segment run pipelined on different warps (anywhere on the chip).
// Work items (strip, segment) are handed out through one device-wide ticket in strip-major order, so a
// strip's producer always holds an earlier ticket (no deadlock whatever is resident), many segments fill
// the chip with no waiting at all, and a lone multi-kb segment spreads over as many SMs as it has strips.
//
// Boundary rows.  Half-lane 63 of strip s stores, per column, one 16-byte entry {M | X << 16, nonce,
// Y | delta << 16, nonce} with a plain volatile store; lanes 0-15 of strip s + 1 fetch 16 entries per chunk
// of 16 steps and poll until both nonces (unique per run, process-wide) are there: every 8-byte half
// validates itself, so there is no fence, no flag and no cleared memory.  (The one hardware assumption: an aligned
// 8-byte half of a 16-byte store becomes visible whole -- the same assumption the low-latency protocols of
// collective libraries make for their {data, flag} words.  Stale memory passes both 32-bit checks with probability
// 2^-64 per entry; an entry of an earlier run carries an earlier nonce.)
//
// Frames.  A half-lane keeps 4 (V - base - off) + priority with a base per (strip, block of S16_CB
// columns): block 0 uses base 0 (H^ <= per_step * min(i, j) is small there), block c >= 1 of strip s > 0
// puts max(M, X, Y) of its top boundary row at column c * S16_CB at S16_F0.  Every cell of the strip in that
// block then lies within [-(|smin| + 3 go), per_step * max(64 R, S16_CB) + 2 go] of the anchor (gap extension
// is free in H^, so a detour costs two gap opens), which fits the int16 range for any segment size.  When a
// half-lane crosses a block boundary it subtracts delta = 4 (base_c - base_{c-1}) from its live registers
// (its row at the previous column and the diagonal inputs); lane 0 derives delta and the producer-to-consumer
// offset from the boundary entry at the block start and the producer's own delta carried in that entry.
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// predicated in the instruction itself: the one storing lane does not split the warp
__device__ __forceinline__ void st_volatile_u4_if(uint4 *p, uint4 v, bool on) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((uint32_t)on));
}

template <int R, int OCC>
__global__ void __launch_bounds__(32, OCC)
dp_wf16s_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, Slot *__restrict__ slots,
                const int32_t *__restrict__ list, const unsigned int *__restrict__ count_dev, Counters *ctr,
                DpParams prm, uint8_t *__restrict__ tb_pool, uint32_t nonce) {
    constexpr int ROWS = 64 * R, CB = S16_CB, ROWW = 32 * R;      // four traceback bits per cell: 32 R bytes per step row
    static_assert(R == 4 || R == 8, "strip kernels: 4 or 8 rows per half-lane (prm.h.strip; dp_tb16_kernel reads the same value)");
    __shared__ uint32_t s_lut[4];
    __shared__ __align__(16) uint4 s_in[2][16];                   // lane 0's inputs of one chunk: row above (M, X, Y) and the selector
    const int lane = threadIdx.x;
    const Dp16 &h = prm.h;
    if (lane < 4) s_lut[lane] = h.lut[lane];
    __syncwarp();
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t c256 = h.mul[3], c16 = h.mul[2];
    const Wf16Consts cc{h.G, h.tagX, h.tagY, h.mul[0], h.nmul[0], h.nmul[1]};
    const uint32_t tagM = h.tagM, tagX = h.tagX, tagY = h.tagY;
    const uint32_t negM = dup16(h.qneg) | tagM, negX = dup16(h.qneg) | tagX, negY = dup16(h.qneg) | tagY;
    const uint32_t r0M0 = (dup16(h.q0) | tagM) << 16, r0Mn = negM << 16, r0X = negX << 16,
                   r0Y = (dup16(h.qy0) | tagY) << 16, r0Yn = negY << 16;
    const int AF = 4 * (S16_F0 - h.off);                          // encoded value of a frame's anchor
    const unsigned int count = *count_dev;
    if (count == 0) return;
    const unsigned long long total = (unsigned long long)((ctr->long_max_m + ROWS - 1) / ROWS) * count;

    for (;;) {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(&ctr->long_ticket, 1ULL);
        ticket = __shfl_sync(FULL, ticket, 0);
        if (ticket >= total) break;
        const int s = (int)(ticket / count);                      // strip-major: every segment's strip s before any strip s + 1
        const int sid = list[ticket % count];
        const Slot sl = slots[sid];
        const int m = sl.m, n = sl.n;
        const int S = (m + ROWS - 1) / ROWS;
        if (s >= S) continue;
        const PairDesc pd = pairs[sl.pair];
        const uint8_t *xs = seqs + pd.off0 + sl.x0, *ys = seqs + pd.off1 + sl.y0;
        const int i0 = s * ROWS;
        const bool last = s == S - 1, producer = !last && lane == 31;
        const int64_t tb_rows = (int64_t)n + WF16_TB_PAD;
        uint8_t *tbseg = tb_pool + sl.tb_off;
        uint8_t *tbw = tbseg + ((int64_t)s * tb_rows + 16) * ROWW + lane * R;
        // boundary rows live behind the segment's S traceback blocks (the allocation is sized for a byte per cell)
        uint4 *bnd = reinterpret_cast<uint4 *>(tbseg + (int64_t)S * tb_rows * ROWW);
        const int64_t bstride = (int64_t)n + 1;
        const uint4 *bin = bnd + (int64_t)(s - 1) * bstride;      // s > 0
        uint4 *bout = bnd + (int64_t)s * bstride;                 // s < S - 1
        uint32_t lutA[R], lutB[R], Mr[R], Xr[R], Yr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int iA = i0 + lane * 2 * R + r, iB = iA + R;    // 0-based matrix rows
            lutA[r] = s_lut[iA < m ? base_code(xs[iA]) : 0];
            lutB[r] = s_lut[iB < m ? base_code(xs[iB]) : 0];
            Mr[r] = negM; Xr[r] = negX; Yr[r] = negY;
        }
        uint32_t dgM = negM, dgX = negX, dgY = negY;
        // a strip that feeds another one runs all 64 half-lanes to column n; the last one stops at the terminal cell
        const int ms = min(ROWS, m - i0);
        const int ve = last ? (ms - 1) / R : 63, re = (ms - 1) % R, te = n + ve, le = ve >> 1;
        uint32_t sel_cur = 0x8888u, sel_prev = 0x8888u;
        int cross = CB;                                           // first column of the next block this lane's low half enters
        uint32_t dlt = 0, dhi = 0;                                // delta of the block being entered (both halves); of the high half's block
        int bsum = 0, rel = 0, a_prev = AF;                       // 4 * base of the newest block; producer -> consumer offset; last anchor
        uint4 E = make_uint4(0u, 0u, 0u, 0u);
        uint32_t yy = 0;
        auto fetch = [&](int t0) {                                // lanes 0-15: boundary entry and selector of column t0 + lane
            const int c = t0 + lane;
            E = make_uint4(0u, 0u, 0u, 0u); yy = 0;
            if (lane < 16) {
                if (s > 0 && c <= n) E = ld_volatile_u4(bin + c);
                const uint32_t yc = (c >= 1 && c <= n) ? (uint32_t)base_code(ys[c - 1]) : 0u;
                const uint32_t yp = (c >= 2 && c <= n + 1) ? (uint32_t)base_code(ys[c - 2]) : 0u;
                yy = 0xC480u + 0x11u * yc + 0x1100u * yp;         // columns (c, c - 1) for lane 0's (low, high) half
            }
        };
        fetch(0);
        int buf = 0;
        uint8_t *tbp = tbw;                                       // step t writes row t of the strip's block
        uint4 *boutp = bout - 63;
        for (int t0 = 0; t0 <= te; t0 += 16, buf ^= 1) {
            const int c = t0 + lane;
            if (s > 0) {
                const bool need = lane < 16 && c <= n;
                // (a consumer that has caught up with its producer backs off: its polls take issue slots from the
                // warps it shares the SM with -- possibly the producer itself)
                long long t_wait = 0;
                for (unsigned int nap = 32;;) {
                    const bool ok = !need || (E.y == nonce && E.w == nonce);
                    if (__all_sync(FULL, ok)) break;
                    // a producer is never more than a chunk's work away; seconds of waiting mean a bug -- say so and go on
                    // (the host turns the flag into an error) rather than hang the device
                    if (t_wait == 0) t_wait = clock64();
                    else if (__any_sync(FULL, clock64() - t_wait > 10000000000LL)) { if (lane == 0) atomicExch(&ctr->strip_timeout, 1u); break; }
                    __nanosleep(nap);
                    if (OCC > 8 && nap < 2048u) nap *= 2u;        // (the few-segments form runs a warp per SM: nothing to yield to)
                    if (!ok) E = ld_volatile_u4(bin + c);
                }
                if ((t0 & (CB - 1)) == 0) {                       // lane 0 enters a block: its frame
                    int dnew = 0;
                    if (t0 > 0 && t0 <= n) {
                        const int eM = (int)(int16_t)(E.x & 0xFFFFu), eX = (int)(int16_t)(E.x >> 16), eY = (int)(int16_t)(E.z & 0xFFFFu);
                        int a = max(eM, max(eX, eY)) & ~3;
                        int dP = (int)(int16_t)(E.z >> 16);
                        a = __shfl_sync(FULL, a, 0); dP = __shfl_sync(FULL, dP, 0);
                        dnew = dP + a - a_prev; a_prev = a; rel = AF - a; bsum += dnew;
                    }
                    dlt = dup16(dnew);
                }
            }
            if (lane < 16) {
                uint4 o;
                if (s == 0) {
                    o.x = c == 0 ? r0M0 : r0Mn; o.y = r0X; o.z = c >= 1 ? r0Y : r0Yn;
                } else {
                    const uint32_t rr = (uint32_t)rel;
                    o.x = (E.x + rr) << 16; o.y = ((E.x >> 16) + rr) << 16; o.z = (E.z + rr) << 16;
                }
                o.w = yy;
                s_in[buf][lane] = o;
            }
            __syncwarp();
            if (t0 + 16 <= te) fetch(t0 + 16);
            const int ns = min(16, te + 1 - t0);
            // some half-lane of the warp enters a block during this chunk (lane l at steps c CB + 2 l, + 1)
            const bool xing = s > 0 && t0 >= CB && (t0 & (CB - 1)) < 64;
#pragma unroll 2
            for (int sidx = 0; sidx < ns; ++sidx) {
                const int t = t0 + sidx;
                const uint4 in = s_in[buf][sidx];
                const uint32_t sel_in = __shfl_up_sync(FULL, sel_prev, 1);
                sel_prev = sel_cur;
                sel_cur = lane == 0 ? in.w : sel_in;
                uint32_t sM = __shfl_up_sync(FULL, Mr[R - 1], 1);
                uint32_t sX = __shfl_up_sync(FULL, Xr[R - 1], 1);
                uint32_t sY = __shfl_up_sync(FULL, Yr[R - 1], 1);
                if (lane == 0) { sM = in.x; sX = in.y; sY = in.z; }
                const uint32_t uM = prmt(sM, Mr[R - 1], 0x5432u), uX = prmt(sX, Xr[R - 1], 0x5432u),
                               uY = prmt(sY, Yr[R - 1], 0x5432u);
                // block boundary: the low half enters at d == 0, the high half one step later
                if (xing) {
                    const int d = t - 2 * lane - cross;
                    if ((unsigned)d < 2u) {
                        const uint32_t dd = d == 0 ? (dlt & 0xFFFFu) : (dlt & 0xFFFF0000u);
#pragma unroll
                        for (int r = 0; r < R; ++r) { Mr[r] = __vsub2(Mr[r], dd); Xr[r] = __vsub2(Xr[r], dd); Yr[r] = __vsub2(Yr[r], dd); }
                        dgM = __vsub2(dgM, dd); dgX = __vsub2(dgX, dd); dgY = __vsub2(dgY, dd);
                        if (d == 1) { cross += CB; dhi = dlt & 0xFFFF0000u; }
                    }
                }
                uint32_t D[R];
                wf16_rows<R, false>(Mr, Xr, Yr, D, lutA, lutB, sel_cur, dgM, dgX, dgY, uM, uX, uY, cc);
                uint32_t F[R / 4];
#pragma unroll
                for (int g = 0; g < R / 4; ++g)
                    F[g] = imad(imad(D[4 * g + 3], c16, D[4 * g + 2]), c256, imad(D[4 * g + 1], c16, D[4 * g]));
                if constexpr (R == 4) *reinterpret_cast<uint32_t *>(tbp) = F[0];
                else *reinterpret_cast<uint2 *>(tbp) = make_uint2(F[0], F[1]);
                tbp += ROWW;
                dgM = uM; dgX = uX; dgY = uY;
                // the strip's bottom row at column t - 63
                const uint4 ent = make_uint4((Mr[R - 1] >> 16) | (Xr[R - 1] & 0xFFFF0000u), nonce, (Yr[R - 1] >> 16) | dhi, nonce);
                st_volatile_u4_if(boutp, ent, producer && t >= 63);
                ++boutp;
            }
        }
        if (last) {
            uint32_t eM = 0, eX = 0, eY = 0;
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (r == re) { eM = Mr[r]; eX = Xr[r]; eY = Yr[r]; }
            eM = __shfl_sync(FULL, eM, le); eX = __shfl_sync(FULL, eX, le); eY = __shfl_sync(FULL, eY, le);
            const int sh = (ve & 1) ? 16 : 0;
            const int vM = (int)(int16_t)(eM >> sh), vX = (int)(int16_t)(eX >> sh), vY = (int)(int16_t)(eY >> sh);
            const int cm = (vM & ~3) + prm.endM, cx = (vX & ~3) + prm.endX, cy = (vY & ~3) + prm.endY;
            int best = cm, st = 0;
            if (cx > best) { best = cx; st = 1; }
            if (cy > best) { best = cy; st = 2; }
            const int score = (best >> 2) + h.off + (bsum >> 2) - prm.ge_raw * (m + n);
            if (lane == 0) { slots[sid].score = score; slots[sid].len = st; }    // dp_tb16_kernel walks from state st
        }
        __syncwarp();
    }
}

cudaError_t launch_dp_strips16(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                               const unsigned int *count_dev, Counters *ctr, int grid, DpParams prm, uint8_t *tb,
                               uint32_t nonce, cudaStream_t st) {
    if (grid <= 0 || !prm.h.strip) return cudaSuccess;
    // few strips in flight: 256-row strips (a lone warp's step is half as long, twice as many warps per segment), every
    // register the step wants (eight warps per SM at most), no poll back-off; many: 512-row strips, sixteen warps per SM
    if (prm.h.strip == 4) dp_wf16s_kernel<4, 8><<<grid / 2, 32, 0, st>>>(seqs, pairs, slots, list, count_dev, ctr, prm, tb, nonce);
    else dp_wf16s_kernel<8, 16><<<grid, 32, 0, st>>>(seqs, pairs, slots, list, count_dev, ctr, prm, tb, nonce);
    return cudaGetLastError();
}

// Traceback of every segment the packed forward kernels processed: one warp per segment over
// the four warp-class lists.  Kept out of the forward kernels because the walk is a chain of
// memory latencies (window loads, character loads): inside them it held a third of the warps
// of an SM sub-partition off the ALU pipe; here 64 warps per SM overlap their latencies.
constexpr int TB16_WARPS = 4;
__global__ void __launch_bounds__(TB16_WARPS * 32)
dp_tb16_kernel(const uint8_t *__restrict__ seqs, const PairDesc *__restrict__ pairs, Slot *__restrict__ slots,
               const int32_t *__restrict__ lists, int64_t list_cap, const unsigned int *__restrict__ counts_dev,
               int cls_lo, int cls_hi, DpParams prm, uint8_t *__restrict__ strbuf, const uint8_t *tb_pool) {
    __shared__ __align__(16) uint8_t win[TB16_WARPS][WALK_WIN_ROWS][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned int gw = blockIdx.x * TB16_WARPS + w, nw = gridDim.x * TB16_WARPS;
    const uint32_t pmap = ((uint32_t)prio_to_state(0, prm)) | ((uint32_t)prio_to_state(1, prm) << 2) |
                          ((uint32_t)prio_to_state(2, prm) << 4);
    for (int cls = cls_lo; cls <= cls_hi; ++cls) {             // warp classes; the long class: what dp_wf16s_kernel ran (wf16_pick)
        const unsigned int count = counts_dev[cls];
        const int32_t *list = lists + (int64_t)cls * list_cap;
        for (unsigned int item = gw; item < count; item += nw) {
            const int sid = list[item];
            const Slot s = slots[sid];
            const int m = s.m, n = s.n;
            const int pick = wf16_pick(prm.h, cls, m, n);
            if (!pick) continue;                             // a 32-bit kernel did forward pass and traceback
            const bool paired = pick >= WF16_PAIR;
            const int R = paired ? pick - WF16_PAIR : pick;
            const PairDesc pd = pairs[s.pair];
            const uint8_t *xs = seqs + pd.off0 + s.x0, *ys = seqs + pd.off1 + s.y0;
            const int64_t row_len = (pd.off1 - pd.off0) + pd.n1;
            uint8_t *ox = strbuf + s.str_off, *oy = ox + row_len;
            const int rsh = 31 - __clz(R);
            // paired kernel: the 32-bit kernels' table (32 R bytes per row, matrix-row order, true bytes);
            // half-lane kernel: 64 R bytes per row, row pairs interleaved, biased words
            // (compact form, four bits per cell: R >= 4 rows per half-lane and no X <-> Y, as in dp_wf16_kernel)
            const bool nibf = !prm.xy && (paired || R >= 4);
            const WalkGeom g = paired ? WalkGeom{5 + rsh, rsh, -1, 0u, nibf ? 0 : -1}
                                      : WalkGeom{6 + rsh, rsh, R >= 2 ? rsh : -1, R >= 2 ? prm.h.tbk * 0x01010101u : 0u, nibf ? rsh : -1};
            int st = s.len;                                  // half-lane kernel: the terminal state travels in slot.len
            if (paired) {                                    // paired kernel: last-column states at the head of the table
                const uint16_t *eb = reinterpret_cast<const uint16_t *>(tb_pool + s.tb_off);
                const int roww = 32 * R;
                const int vM = (int)(int16_t)eb[m - 1], vX = (int)(int16_t)eb[roww + m - 1], vY = (int)(int16_t)eb[2 * roww + m - 1];
                const int cm = (vM & ~3) + prm.endM, cx = (vX & ~3) + prm.endX, cy = (vY & ~3) + prm.endY;
                int best = cm; st = 0;
                if (cx > best) { best = cx; st = 1; }
                if (cy > best) { best = cy; st = 2; }
                if (lane == 0) slots[sid].score = (best >> 2) + prm.h.off - prm.ge_raw * (m + n);
            }
            const int len = wf_traceback(win[w], tb_pool + s.tb_off + (16 << (nibf ? g.roww_sh - 1 : g.roww_sh)), (int64_t)n + WF16_TB_PAD, xs, ys, m, n, st, pmap,
                                         ox, oy, lane, g);
            if (lane == 0) slots[sid].len = len;
            __syncwarp();
        }
    }
}

template <int R>
static void launch_wf16_t(int grid, const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                          const unsigned int *count_dev, const DpParams &prm, uint8_t *strbuf, uint8_t *tb, cudaStream_t st) {
    // grid comes sized for 16 warps per SM; these kernels fit 20 (R <= 8) or 12 (R = 16)
    grid = grid * (R == 16 ? 12 : 20) / 16;
    if (prm.xy) dp_wf16_kernel<R, true><<<grid, 32, 0, st>>>(seqs, pairs, slots, list, count_dev, prm, strbuf, tb);
    else dp_wf16_kernel<R, false><<<grid, 32, 0, st>>>(seqs, pairs, slots, list, count_dev, prm, strbuf, tb);
}

template <int R>
static void launch_wf16p_t(int grid, const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                           const unsigned int *count_dev, const DpParams &prm, uint8_t *tb, cudaStream_t st) {
    grid = grid * 20 / 16;
    if (prm.xy) dp_wf16p_kernel<R, true><<<grid, 32, 0, st>>>(seqs, pairs, slots, list, count_dev, prm, tb);
    else dp_wf16p_kernel<R, false><<<grid, 32, 0, st>>>(seqs, pairs, slots, list, count_dev, prm, tb);
}

// cls: CLS_WARP2 / 4 / 8 / 16; the packed kernel runs half as many rows per half-lane as the
// class's 32-bit kernel runs per lane (same 32 * R-byte traceback rows)
cudaError_t launch_dp_wavefront16(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                                  const unsigned int *count_dev, int grid, int cls, DpParams prm, uint8_t *strbuf,
                                  uint8_t *tb, cudaStream_t st, int *launches) {
    if (grid <= 0 || !prm.h.ok) return cudaSuccess;
    const int rpl = wf_rows_per_lane(cls);
    bool all_paired = false;       // every member of the class goes to the two-segments-per-warp kernel
    if (prm.h.pair) {
        if (cls == CLS_WARP2) { launch_wf16p_t<2>(grid, seqs, pairs, slots, list, count_dev, prm, tb, st); *launches += 1; }
        else if (cls == CLS_WARP4) { launch_wf16p_t<4>(grid, seqs, pairs, slots, list, count_dev, prm, tb, st); *launches += 1; }
        else if (cls == CLS_WARP8) { launch_wf16p_t<8>(grid, seqs, pairs, slots, list, count_dev, prm, tb, st); *launches += 1; }
        all_paired = wf16_pick(prm.h, cls, 32 * rpl, 1 << 30) == WF16_PAIR + rpl;
    }
    if (all_paired) return cudaGetLastError();
    // half-lane kernels: the whole class without pairing; with it, what the paired kernel cannot take
    // (the open-ended class's segments of more than 512 rows, or scoring too steep for 32 R rows)
    if (cls == CLS_WARP2) { launch_wf16_t<1>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, st); *launches += 1; }
    else if (cls == CLS_WARP4) { launch_wf16_t<2>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, st); *launches += 1; }
    else if (cls == CLS_WARP8) { launch_wf16_t<4>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, st); *launches += 1; }
    else {
        { launch_wf16_t<8>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, st); *launches += 1; }
        { launch_wf16_t<16>(grid, seqs, pairs, slots, list, count_dev, prm, strbuf, tb, st); *launches += 1; }
    }
    return cudaGetLastError();
}

cudaError_t launch_dp_traceback16(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *lists,
                                  int64_t list_cap, const unsigned int *counts_dev, int cls_lo, int cls_hi, int grid, DpParams prm,
                                  uint8_t *strbuf, const uint8_t *tb, cudaStream_t st) {
    if (grid <= 0 || !prm.h.ok) return cudaSuccess;
    dp_tb16_kernel<<<grid, TB16_WARPS * 32, 0, st>>>(seqs, pairs, slots, lists, list_cap, counts_dev, cls_lo, cls_hi, prm, strbuf, tb);
    return cudaGetLastError();
}

}  // namespace dbga
