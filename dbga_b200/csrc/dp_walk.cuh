// Traceback walk shared by the wavefront kernels (32-bit and packed 16-bit forward passes).
//
// The forward pass leaves one byte per cell (bits 0-1: predecessor tag of M, 2-3: of X, 4-5: of Y)
// in a step-major table: cell (i, j), 1-based, with i0 = i - 1, lives in block i0 / ROWW at
// step row j + (i0 % ROWW) / RDIV, byte column i0 % ROWW (RDIV = matrix rows per wavefront
// lane, ROWW = matrix rows per block = bytes per step row).  Reproduces the traceback of
// cogent3's global_pairwise as restated in oracle/ (reference call site src/dbga/utils.py:185).
#pragma once
#include "kernels.cuh"

namespace dbga {

__device__ __forceinline__ int prio_to_state(int prio, const DpParams &p) {
    // 0 = M, 1 = X, 2 = Y
    const int t = prio * TAGMUL;
    return t == p.tagM ? 0 : (t == p.tagX ? 1 : 2);
}

// 2-bit codes of the (up to) 16 bases at p, any alignment, code i in bits 2i..2i+1: two aligned
// 128-bit loads and a branch-free word/byte shift instead of 16 dependent-latency byte loads.
// The second block is only read when it holds one of the n bases; n == 0 reads nothing.
__device__ __forceinline__ uint32_t load_codes16(const uint8_t *p, int n) {
    if (n <= 0) return 0u;
    const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15);
    const uint4 *q = reinterpret_cast<const uint4 *>(p - off);
    const uint4 a = __ldg(q);
    uint4 b = make_uint4(0u, 0u, 0u, 0u);
    if (off + (uint32_t)min(n, 16) > 16u) b = __ldg(q + 1);
    const uint32_t ws = off >> 2, sh = (off & 3u) * 8u;
    // words ws .. ws+4 of (a, b)
    const uint32_t v0 = ws == 0 ? a.x : (ws == 1 ? a.y : (ws == 2 ? a.z : a.w));
    const uint32_t v1 = ws == 0 ? a.y : (ws == 1 ? a.z : (ws == 2 ? a.w : b.x));
    const uint32_t v2 = ws == 0 ? a.z : (ws == 1 ? a.w : (ws == 2 ? b.x : b.y));
    const uint32_t v3 = ws == 0 ? a.w : (ws == 1 ? b.x : (ws == 2 ? b.y : b.z));
    const uint32_t v4 = ws == 0 ? b.x : (ws == 1 ? b.y : (ws == 2 ? b.z : b.w));
    auto pack4 = [](uint32_t w) {                                  // 4 ASCII bases -> c0 | c1<<2 | c2<<4 | c3<<6
        const uint32_t codes = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
        return (codes * 0x01041040u) >> 24;
    };
    const uint32_t out = pack4(__funnelshift_r(v0, v1, sh)) | (pack4(__funnelshift_r(v1, v2, sh)) << 8) |
                         (pack4(__funnelshift_r(v2, v3, sh)) << 16) | (pack4(__funnelshift_r(v3, v4, sh)) << 24);
    return n >= 16 ? out : (out & ((1u << (2 * n)) - 1u));
}
__device__ __forceinline__ uint32_t code_char(uint32_t code) { return __byte_perm(0x54474341u, 0u, code) & 0xFFu; }   // "ACGT"[code]

// One warp walks; every lane runs the same walk (uniform control flow).  win: 32 x 64 bytes of
// shared memory owned by the warp.  st: state chosen at the terminal cell.  pmap: packed table
// tag -> state (0 = M, 1 = X, 2 = Y).  Rows are written reversed at ox / oy; returns the
// number of alignment columns.
// Table geometry.  roww_sh: log2 of the bytes per step row (= matrix rows per block); rdiv_sh:
// log2 of the matrix rows per wavefront lane (half-lane for the packed kernel); pr_sh >= 1: the
// packed kernel's byte order inside a lane's 2 * PR bytes, PR = 1 << pr_sh (pairs of rows
// (r, r+1) of the low and high half-lanes: [low r, low r+1, high r, high r+1]), each group of
// four bytes stored as one word less bias4 (exact 32-bit arithmetic, see stage3_dp16.cu);
// pr_sh < 0: bytes in matrix-row order, no bias.  The 32-bit kernels pass constants, which
// fold away after inlining.
// nib_sh == 0: the compact table of the two-segments-per-warp kernel (four bits per cell, byte = matrix row / 2).
// nib_sh >= 2: the compact table of the packed kernel (X <-> Y off, R = 1 << nib_sh >= 4 rows per half-lane):
// FOUR bits per cell -- bits 0-1 the predecessor tag of M, bit 2 "X came from M", bit 3 "Y came from M" --
// two matrix rows per byte; a lane's R bytes of a step row hold, per group of four rows r..r+3 of its two
// half-lanes, [low (r, r+1), low (r+2, r+3), high (r, r+1), high (r+2, r+3)].  A step row is then half as
// many bytes as it covers matrix rows.
constexpr int WALK_WIN_ROWS = 64;
struct WalkGeom {
    int roww_sh, rdiv_sh, pr_sh;
    uint32_t bias4;
    int nib_sh = -1;
};

__device__ __forceinline__ int wf_col(int c, int pr_sh) {
    if (pr_sh < 0) return c;
    const int pr = 1 << pr_sh;
    const int q = c & (2 * pr - 1), hh = q >> pr_sh, r = q & (pr - 1);
    return c - q + 4 * (r >> 1) + 2 * hh + (r & 1);
}
// byte column of matrix row c (within its block) in the compact table; the cell is nibble (c & 1) of it
__device__ __forceinline__ int wf_col_nib(int c, int nib_sh) {
    if (nib_sh == 0) return c >> 1;                  // two-segments-per-warp kernel: bytes in matrix-row order
    const int R = 1 << nib_sh;
    const int q = c & (2 * R - 1), hh = q >> nib_sh, r = q & (R - 1);
    return ((c - q) >> 1) + 4 * (r >> 2) + 2 * hh + ((r >> 1) & 1);
}

// One warp walks; every lane runs the same walk (uniform control flow).  win: WALK_WIN_ROWS x 64 bytes of
// shared memory owned by the warp.  st: state chosen at the terminal cell.  pmap: packed table
// tag -> state (0 = M, 1 = X, 2 = Y).  Rows are written reversed at ox / oy; returns the
// number of alignment columns.
__device__ __forceinline__ int wf_traceback(uint8_t (*win)[64], const uint8_t *__restrict__ tb, int64_t tb_rows,
                                            const uint8_t *__restrict__ xs, const uint8_t *__restrict__ ys, int m, int n,
                                            int st, uint32_t pmap, uint8_t *__restrict__ ox, uint8_t *__restrict__ oy, int lane,
                                            const WalkGeom g) {
    constexpr int WIN_W = 64;                        // traceback window width in bytes (every row is >= 64 bytes)
    constexpr int WIN_ROWS = WALK_WIN_ROWS;          // step rows per window
    const int roww_mask = (1 << g.roww_sh) - 1;
    const bool nib = g.nib_sh >= 0;
    const int rowb_sh = nib ? g.roww_sh - 1 : g.roww_sh;         // log2 of the bytes per step row
    int i = m, j = n, len = 0;
    // window: WIN_ROWS consecutive step rows x 64 bytes of one (pass, warp) block.  Every lane
    // runs the same walk (uniform control flow).  Only the kind of each step is kept, as
    // two bit masks per group of 32 steps (bit t of mY: step t consumes a seq-1 base, of
    // mX: a seq-2 base); a full group is turned into the two rows' characters by all 32
    // lanes at once, so the sequence loads stay off the walk's dependency chain: a group's
    // characters are loaded when it is full and stored when the next one is (or at the end).
    // The 32 lanes probe the next 32 cells in the current state's direction together and the walk
    // jumps over the whole run of steps that stay in the state (M -> M along the diagonal: most of
    // the path between similar sequences; a gap up a column or along a row) plus the step that leaves it.
    int win_blk = -1, win_row_hi = 0, win_col_lo = 0;
    uint32_t mY = 0, mX = 0;
    int pos = 0, gbase = 0;
    int ib = m, jb = n;                              // (i, j) before the first step of the group
    const uint32_t lt = (1u << lane) - 1u;
    uint8_t pend_x = 0, pend_y = 0;                  // characters of the previous group, not stored yet
    int pend_at = -1;
    auto drain = [&]() {
        if (pend_at >= 0) { ox[pend_at] = pend_x; oy[pend_at] = pend_y; }
        pend_at = -1;
    };
    auto flush = [&](int cnt) {
        drain();
        if (lane < cnt) {
            const int it = ib - __popc(mY & lt), jt = jb - __popc(mX & lt);
            pend_x = ((mY >> lane) & 1u) ? xs[it - 1] : (uint8_t)'-';
            pend_y = ((mX >> lane) & 1u) ? ys[jt - 1] : (uint8_t)'-';
            pend_at = gbase + lane;
        }
        ib -= __popc(mY); jb -= __popc(mX);
        gbase += cnt; mY = 0; mX = 0; pos = 0;
    };
    auto emit = [&](int state, int count) {
        len += count;
        while (count > 0) {
            const int take = min(count, 32 - pos);
            const uint32_t bits = (take == 32 ? 0xffffffffu : ((1u << take) - 1u)) << pos;
            if (state != 2) mY |= bits;
            if (state != 1) mX |= bits;
            pos += take; count -= take;
            if (pos == 32) flush(32);
        }
    };
    auto tb_byte = [&](int r, int c) -> uint32_t {
        if (g.pr_sh < 0) return win[r][c];
        return ((*reinterpret_cast<const uint32_t *>(&win[r][c & ~3]) + g.bias4) >> (8 * (c & 3))) & 0xFFu;
    };
    // table column (byte) of matrix row cb of a block
    auto col_of = [&](int cb) -> int { return nib ? wf_col_nib(cb, g.nib_sh) : wf_col(cb, g.pr_sh); };
    while (i > 0 || j > 0) {
        if (i == 0) { emit(2, j); break; }           // first row / column: one long gap
        if (j == 0) { emit(1, i); break; }
        const int i0 = i - 1;
        // Lane l looks at the cell l steps further along the current state's direction (M: the diagonal, X: up the
        // column, Y: along the row): does the path stay in this state there?  The walk takes the whole run of
        // cells that do, plus the cell that ends it (still consumed in this state; its tag names the next one).
        const int di = st != 2, dj = st != 1;
        const int il = i0 - lane * di, jl = j - lane * dj;
        uint32_t nxt = 3u;                           // 3: outside the matrix or the window
        if (il >= 0 && jl >= 1) {
            const int bl = il >> g.roww_sh, cl = il & roww_mask;
            const unsigned dr = (unsigned)(win_row_hi - (jl + (cl >> g.rdiv_sh)));
            const int wc = col_of(cl);
            const unsigned dc = (unsigned)((nib ? wc : cl) - win_col_lo);
            if (bl == win_blk && dr < (unsigned)WIN_ROWS && dc < (unsigned)WIN_W) {
                if (nib) {
                    // compact cell: bits 0-1 the predecessor tag of M, bit 2 "X came from M", bit 3 "Y came from M"
                    const uint32_t v = (uint32_t)win[dr][wc - win_col_lo] >> (4 * (cl & 1));
                    nxt = st == 0 ? (pmap >> (2 * (v & 3u))) & 3u : ((v & (st == 1 ? 4u : 8u)) ? 0u : (uint32_t)st);
                } else {
                    nxt = (pmap >> (2 * ((tb_byte(dr, wc - win_col_lo) >> (2 * st)) & 3u))) & 3u;
                }
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, nxt == (uint32_t)st);
        const int c = bal == 0xffffffffu ? 32 : __ffs((int)~bal) - 1;
        const uint32_t nx = __shfl_sync(0xffffffffu, nxt, c & 31);
        if (c == 0 && nx == 3u) {
            // the walk's own cell is outside the window (lane 0 is inside the matrix here): centre a new one on it
            const int blk = i0 >> g.roww_sh, cb = i0 & roww_mask;      // (pass * GW + warp), l * R + r
            const int row = j + (cb >> g.rdiv_sh);
            const int ccol = nib ? wf_col_nib(cb, g.nib_sh) : cb;       // window position of the cell's byte (byte form: by matrix row)
            __syncwarp();
            win_blk = blk; win_row_hi = row;
            win_col_lo = max(0, (ccol & ~31) - 32);
#pragma unroll
            for (int hrow = 0; hrow < WIN_ROWS; hrow += 32) {
                const int rr = row - hrow - lane;
                uint4 a = make_uint4(0, 0, 0, 0), b = a, c4 = a, d4 = a;
                if (rr >= 1) {
                    const uint4 *p = reinterpret_cast<const uint4 *>(tb + (((int64_t)blk * tb_rows + rr) << rowb_sh) + win_col_lo);
                    a = p[0]; b = p[1]; c4 = p[2]; d4 = p[3];
                }
                uint4 *dst = reinterpret_cast<uint4 *>(&win[hrow + lane][0]);
                dst[0] = a; dst[1] = b; dst[2] = c4; dst[3] = d4;
            }
            __syncwarp();
            continue;
        }
        const int take = c + ((c < 32 && nx != 3u) ? 1 : 0);
        emit(st, take);
        i -= take * di; j -= take * dj;
        if (c < 32 && nx != 3u) st = (int)nx;
    }
    if (pos) flush(pos);
    drain();
    return len;
}

}  // namespace dbga
