/* CPU oracle for DBGA's pairwise alignment hot path, plain C  --  TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker and the timed CPU baseline ("port"), never the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  Nothing under dbga_b200/ links or calls it.
 *
 * Restates (same semantics as oracle/dbga_oracle.py, which is itself machine-checked
 * against the reference's unmodified code, see oracle/make_golden.py):
 *   stages 1-2  reference src/dbga/utils.py:108-154 (get_kmers, duplicate_kmers),
 *               src/dbga/debruijn_pairwise.py:78-225 (node / merge-node ids),
 *               :277-300 + utils.py:266-291 (cycle check);
 *   stage 3     debruijn_pairwise.py:345-395, 418-543 (segments, shortcut, first-column
 *               strip, tail) and utils.py:157-192 (dna_global_aln, empty-side cases);
 *   DP          integer 3-state Gotoh standing in for cogent3.align.global_pairwise
 *               (utils.py:185; cogent3 pinned to the moving "develop" branch,
 *               pyproject.toml:15, absent from this image).  Parity with real cogent3 is
 *               UNPINNED beyond the reference's golden vectors; every convention is a
 *               parameter (struct oracle_conv).
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ST_OK 0
#define ST_CYCLE 1
#define ST_K_INVALID 2
#define ST_BAD_CHAR 3
#define ST_TOO_LARGE 4

#define NEGV (-(1 << 30))

typedef struct oracle_conv {
    int32_t gap_len_mode; /* 0: d+(L-1)e   1: d+L*e */
    int32_t xy_allowed;
    int32_t pred_order;   /* index into XMY,XYM,MXY,MYX,YXM,YMX */
    int32_t end_order;
} oracle_conv;

/* state ids */
enum { S_M = 0, S_X = 1, S_Y = 2 };
static const int ORDERS[6][3] = {
    {S_X, S_M, S_Y}, {S_X, S_Y, S_M}, {S_M, S_X, S_Y},
    {S_M, S_Y, S_X}, {S_Y, S_X, S_M}, {S_Y, S_M, S_X}};

static inline int code_of(unsigned char c) {
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
    }
}

/* ------------------------------------------------------------------ Gotoh (stage 3) */
/* ordered first-wins argmax over up to three candidates; absent = state not allowed */
static inline void pick3(const int ord[3], const int32_t v[3], const int present[3],
                         int32_t *best, int *who) {
    int have = 0;
    for (int t = 0; t < 3; ++t) {
        int st = ord[t];
        if (!present[st]) continue;
        if (!have || v[st] > *best) { *best = v[st]; *who = st; have = 1; }
    }
}

/* x: m codes, y: n codes (0..3).  out_x/out_y receive the gapped rows (ASCII, not
 * NUL-terminated), *out_len the column count.  Buffers must hold m+n bytes.
 * Returns the optimal score. */
int32_t oracle_gotoh(const uint8_t *x, int64_t m, const uint8_t *y, int64_t n,
                     const int32_t score[16], int32_t d, int32_t e, const oracle_conv *cv,
                     uint8_t *out_x, uint8_t *out_y, int64_t *out_len) {
    static const char L[4] = {'A', 'C', 'G', 'T'};
    const int *ord = ORDERS[cv->pred_order];
    const int32_t go = cv->gap_len_mode ? d + e : d, ge = e;
    const int64_t W = n + 1;
    int32_t *Mp = malloc(sizeof(int32_t) * W), *Xp = malloc(sizeof(int32_t) * W),
            *Yp = malloc(sizeof(int32_t) * W);
    int32_t *Mc = malloc(sizeof(int32_t) * W), *Xc = malloc(sizeof(int32_t) * W),
            *Yc = malloc(sizeof(int32_t) * W);
    /* one byte per cell: bits 0-1 pred of M, 2-3 pred of X, 4-5 pred of Y */
    uint8_t *tb = malloc((size_t)(m + 1) * W);
    Mp[0] = 0; Xp[0] = NEGV; Yp[0] = NEGV; tb[0] = 0;
    for (int64_t j = 1; j <= n; ++j) {
        Mp[j] = NEGV; Xp[j] = NEGV;
        Yp[j] = -(go + (int32_t)(j - 1) * ge);
        tb[j] = (uint8_t)((j > 1 ? S_Y : S_M) << 4);
    }
    for (int64_t i = 1; i <= m; ++i) {
        uint8_t *row = tb + i * W;
        Mc[0] = NEGV; Yc[0] = NEGV;
        Xc[0] = -(go + (int32_t)(i - 1) * ge);
        row[0] = (uint8_t)((i > 1 ? S_X : S_M) << 2);
        const int32_t *srow = score + 4 * x[i - 1];
        for (int64_t j = 1; j <= n; ++j) {
            int32_t v[3], best = 0; int who = 0, pres[3];
            v[S_M] = Mp[j - 1]; v[S_X] = Xp[j - 1]; v[S_Y] = Yp[j - 1];
            pres[0] = pres[1] = pres[2] = 1;
            pick3(ord, v, pres, &best, &who);
            Mc[j] = best + srow[y[j - 1]];
            uint8_t cell = (uint8_t)who;
            v[S_M] = Mp[j] - go; v[S_X] = Xp[j] - ge; v[S_Y] = Yp[j] - go;
            pres[S_M] = 1; pres[S_X] = 1; pres[S_Y] = cv->xy_allowed;
            pick3(ord, v, pres, &best, &who);
            Xc[j] = best; cell |= (uint8_t)(who << 2);
            v[S_M] = Mc[j - 1] - go; v[S_Y] = Yc[j - 1] - ge; v[S_X] = Xc[j - 1] - go;
            pres[S_M] = 1; pres[S_Y] = 1; pres[S_X] = cv->xy_allowed;
            pick3(ord, v, pres, &best, &who);
            Yc[j] = best; cell |= (uint8_t)(who << 4);
            row[j] = cell;
        }
        int32_t *t;
        t = Mp; Mp = Mc; Mc = t;
        t = Xp; Xp = Xc; Xc = t;
        t = Yp; Yp = Yc; Yc = t;
    }
    int32_t endv[3] = {Mp[n], Xp[n], Yp[n]}, best = 0; int st = 0, pres[3] = {1, 1, 1};
    pick3(ORDERS[cv->end_order], endv, pres, &best, &st);
    /* traceback, writing backwards into the tail of the buffers */
    int64_t i = m, j = n, pos = m + n;
    while (i > 0 || j > 0) {
        uint8_t cell = tb[i * W + j];
        --pos;
        if (st == S_M) {
            out_x[pos] = (uint8_t)L[x[i - 1]]; out_y[pos] = (uint8_t)L[y[j - 1]];
            st = cell & 3; --i; --j;
        } else if (st == S_X) {
            out_x[pos] = (uint8_t)L[x[i - 1]]; out_y[pos] = '-';
            st = (cell >> 2) & 3; --i;
        } else {
            out_x[pos] = '-'; out_y[pos] = (uint8_t)L[y[j - 1]];
            st = (cell >> 4) & 3; --j;
        }
    }
    int64_t len = m + n - pos;
    memmove(out_x, out_x + pos, (size_t)len);
    memmove(out_y, out_y + pos, (size_t)len);
    *out_len = len;
    free(Mp); free(Xp); free(Yp); free(Mc); free(Xc); free(Yc); free(tb);
    return best;
}

/* ------------------------------------------------------------------- stages 1 and 2 */
typedef struct {
    uint64_t key;     /* 2-bit packed k-mer (k <= 32) or a hash of it (k > 32); slot empty when c0 == 0 && c1 == 0 */
    int32_t c0, c1;   /* occurrence counts (saturate at 2) */
    int32_t p0, p1;   /* first position in s0 / s1 */
    const uint8_t *rep; /* k > 32: codes of the occurrence that claimed the slot (compared base by base) */
} slot_t;

static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 33; z *= 0xff51afd7ed558ccdULL; z ^= z >> 33;
    z *= 0xc4ceb9fe1a85ec53ULL; z ^= z >> 33; return z;
}

/* cur != NULL (k > 32): key is a hash; equal hashes are confirmed on the k codes themselves */
static slot_t *tbl_find(slot_t *t, uint64_t mask, uint64_t key, const uint8_t *cur, int32_t k) {
    uint64_t h = mix64(key) & mask;
    for (;;) {
        slot_t *s = &t[h];
        if ((s->c0 | s->c1) == 0) { s->rep = cur; return s; }
        if (s->key == key && (!cur || memcmp(s->rep, cur, (size_t)k) == 0)) return s;
        h = (h + 1) & mask;
    }
}
static uint64_t hash_codes(const uint8_t *c, int32_t k) {
    uint64_t h = 1469598103934665603ULL;
    for (int32_t i = 0; i < k; ++i) { h ^= c[i]; h *= 1099511628211ULL; }
    return h;
}

/* Result of one pair.  Arrays are malloc'd here and released with oracle_free_pair. */
typedef struct oracle_pair {
    int32_t status;
    int64_t n_anchors;
    int32_t *anchors;      /* 2*n_anchors: (pos in s0, pos in s1), ascending pos in s1 */
    int64_t score;         /* sum of DP scores over segments handed to the DP */
    int64_t aln_len;
    uint8_t *aln0, *aln1;
    int64_t dp_cells;      /* sum |x|*|y| over those segments */
    int64_t dp_segments;
    /* graph view (node ids), filled when want_graph != 0 */
    int64_t id_count;
    uint8_t *nondup0, *nondup1;   /* per k-mer position */
    int64_t nk0, nk1;
} oracle_pair;

void oracle_free_pair(oracle_pair *r) {
    free(r->anchors); free(r->aln0); free(r->aln1); free(r->nondup0); free(r->nondup1);
    memset(r, 0, sizeof(*r));
}

static void emit_seg(const uint8_t *c0, int64_t x0, int64_t x1, const uint8_t *c1, int64_t y0,
                     int64_t y1, int drop_first, const int32_t score[16], int32_t d, int32_t e,
                     const oracle_conv *cv, oracle_pair *r, uint8_t *tmpx, uint8_t *tmpy) {
    static const char L[4] = {'A', 'C', 'G', 'T'};
    int64_t m = x1 - x0, n = y1 - y0, len = 0;
    if (m > 0 && n > 0) {
        r->score += oracle_gotoh(c0 + x0, m, c1 + y0, n, score, d, e, cv, tmpx, tmpy, &len);
        r->dp_cells += m * n;
        r->dp_segments += 1;
    } else if (m > 0) {           /* utils.py:187-188 */
        for (int64_t i = 0; i < m; ++i) { tmpx[i] = (uint8_t)L[c0[x0 + i]]; tmpy[i] = '-'; }
        len = m;
    } else if (n > 0) {           /* utils.py:189-190 */
        for (int64_t i = 0; i < n; ++i) { tmpx[i] = '-'; tmpy[i] = (uint8_t)L[c1[y0 + i]]; }
        len = n;
    }
    int64_t skip = (drop_first && len > 0) ? 1 : 0;   /* debruijn_pairwise.py:499-500 */
    memcpy(r->aln0 + r->aln_len, tmpx + skip, (size_t)(len - skip));
    memcpy(r->aln1 + r->aln_len, tmpy + skip, (size_t)(len - skip));
    r->aln_len += len - skip;
}

int32_t oracle_align_pair(const uint8_t *s0, int64_t n0, const uint8_t *s1, int64_t n1, int32_t k,
                          const int32_t score[16], int32_t d, int32_t e, const oracle_conv *cv,
                          int32_t want_graph, int32_t stages, oracle_pair *r) {
    static const char L[4] = {'A', 'C', 'G', 'T'};
    memset(r, 0, sizeof(*r));
    if (k < 1 || k > n0 || k > n1) { r->status = ST_K_INVALID; return r->status; }   /* utils.py:129-131 */
    const int wide = k > 32;
    uint8_t *c0 = malloc((size_t)n0), *c1 = malloc((size_t)n1);
    int bad = 0;
    for (int64_t i = 0; i < n0; ++i) { int c = code_of(s0[i]); if (c < 0) bad = 1; c0[i] = (uint8_t)c; }
    for (int64_t i = 0; i < n1; ++i) { int c = code_of(s1[i]); if (c < 0) bad = 1; c1[i] = (uint8_t)c; }
    if (bad) { free(c0); free(c1); r->status = ST_BAD_CHAR; return r->status; }
    const int64_t N0 = n0 - k + 1, N1 = n1 - k + 1;
    uint64_t cap = 16;
    while (cap < (uint64_t)(2 * (N0 + N1))) cap <<= 1;
    slot_t *tbl = calloc(cap, sizeof(slot_t));
    const uint64_t kmask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    slot_t **sl0 = malloc(sizeof(slot_t *) * (size_t)N0), **sl1 = malloc(sizeof(slot_t *) * (size_t)N1);
    uint64_t key = 0;
    for (int64_t i = 0; i < n0; ++i) {
        key = ((key << 2) | c0[i]) & kmask;
        if (i >= k - 1) {
            const uint8_t *cur = wide ? c0 + (i - k + 1) : NULL;
            if (wide) key = hash_codes(cur, k);
            slot_t *s = tbl_find(tbl, cap - 1, key, cur, k);
            if (s->c0 == 0) s->p0 = (int32_t)(i - k + 1);
            s->key = key; if (s->c0 < 2) s->c0++;
            sl0[i - k + 1] = s;
        }
    }
    key = 0;
    for (int64_t i = 0; i < n1; ++i) {
        key = ((key << 2) | c1[i]) & kmask;
        if (i >= k - 1) {
            const uint8_t *cur = wide ? c1 + (i - k + 1) : NULL;
            if (wide) key = hash_codes(cur, k);
            slot_t *s = tbl_find(tbl, cap - 1, key, cur, k);
            if (s->c1 == 0) s->p1 = (int32_t)(i - k + 1);
            s->key = key; if (s->c1 < 2) s->c1++;
            sl1[i - k + 1] = s;
        }
    }
    /* duplicate = count > 1 in EITHER sequence (utils.py:148-154) */
    if (want_graph) {
        r->nondup0 = malloc((size_t)N0); r->nondup1 = malloc((size_t)N1);
        r->nk0 = N0; r->nk1 = N1;
        int64_t ids = 1;                                      /* 0 = start node of seq 0 */
        for (int64_t p = 0; p < N0; ++p) { r->nondup0[p] = !(sl0[p]->c0 > 1 || sl0[p]->c1 > 1); ids += r->nondup0[p]; }
        ids += 2;                                             /* end of seq 0, start of seq 1 */
        for (int64_t q = 0; q < N1; ++q) {
            r->nondup1[q] = !(sl1[q]->c0 > 1 || sl1[q]->c1 > 1);
            ids += (r->nondup1[q] && sl1[q]->c0 == 0);
        }
        r->id_count = ids + 1;                                /* end of seq 1 */
    }
    /* anchors: k-mers occurring exactly once in each sequence, in seq-1 order */
    int64_t M = 0;
    for (int64_t q = 0; q < N1; ++q) M += (sl1[q]->c0 == 1 && sl1[q]->c1 == 1);
    r->n_anchors = M;
    r->anchors = malloc(sizeof(int32_t) * 2 * (size_t)(M ? M : 1));
    int64_t t = 0;
    for (int64_t q = 0; q < N1; ++q)
        if (sl1[q]->c0 == 1 && sl1[q]->c1 == 1) { r->anchors[2 * t] = sl1[q]->p0; r->anchors[2 * t + 1] = (int32_t)q; ++t; }
    free(tbl); free(sl0); free(sl1);
    /* cycle check: merge ids (monotone in s0 position) strictly increasing along s1 */
    for (int64_t i = 1; i < M; ++i)
        if (!(r->anchors[2 * (i - 1)] < r->anchors[2 * i])) { r->status = ST_CYCLE; break; }
    if (r->status != ST_OK || stages < 3) { free(c0); free(c1); return r->status; }

    r->aln0 = malloc((size_t)(n0 + n1 + 1)); r->aln1 = malloc((size_t)(n0 + n1 + 1));
    uint8_t *tmpx = malloc((size_t)(n0 + n1 + 1)), *tmpy = malloc((size_t)(n0 + n1 + 1));
    if (M == 0) {                                  /* debruijn_pairwise.py:453-457 */
        emit_seg(c0, 0, n0, c1, 0, n1, 0, score, d, e, cv, r, tmpx, tmpy);
    } else {
        for (int64_t i = 0; i < M; ++i) {
            int64_t a = r->anchors[2 * i], b = r->anchors[2 * i + 1];
            if (i == 0) {
                emit_seg(c0, 0, a, c1, 0, b, 0, score, d, e, cv, r, tmpx, tmpy);
            } else {
                int64_t pa = r->anchors[2 * i - 2], pb = r->anchors[2 * i - 1];
                if (!(a - pa == 1 && b - pb == 1))   /* shortcut :384-388 */
                    emit_seg(c0, pa, a, c1, pb, b, 1, score, d, e, cv, r, tmpx, tmpy);
            }
            if (i != M - 1) {                        /* :508-509 */
                r->aln0[r->aln_len] = (uint8_t)L[c0[a]];
                r->aln1[r->aln_len] = (uint8_t)L[c1[b]];
                r->aln_len++;
            }
        }
        emit_seg(c0, r->anchors[2 * M - 2], n0, c1, r->anchors[2 * M - 1], n1, 0, score, d, e, cv, r, tmpx, tmpy);
    }
    free(tmpx); free(tmpy); free(c0); free(c1);
    return r->status;
}

/* Batch form used as the timed CPU baseline: seqs = concatenated ASCII, offsets[2P+1]
 * (sequence j of pair i is seqs[offsets[2i+j] : offsets[2i+j+1]]).  Keeps only per-pair
 * status / score / aligned length / cells so that the timing is the algorithm, not
 * result marshalling.  Pairs are fanned out over `threads` pthreads (dynamic chunks of
 * 16 pairs), which is the multiprocessing fan-out of BASELINE.md section 3. */
typedef struct {
    const uint8_t *seqs; const int64_t *offsets; int64_t P; int32_t k; const int32_t *score;
    int32_t d, e; const oracle_conv *cv;
    int32_t *status; int64_t *scores, *aln_len, *dp_cells, *n_anchors;
    /* optional: rows / anchor runs produced by the implementation under test, compared here
     * pair by pair (diff[i] bit 0: gapped rows differ, bit 1: anchors differ) */
    const int64_t *chk_aln_off; const uint8_t *chk_aln0, *chk_aln1;
    const int64_t *chk_run_off; const int32_t *chk_runs;
    int32_t *diff;
    atomic_long next;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (;;) {
        long lo = atomic_fetch_add(&j->next, 16);
        if (lo >= j->P) break;
        long hi = lo + 16 < j->P ? lo + 16 : j->P;
        for (long i = lo; i < hi; ++i) {
            oracle_pair r;
            const int64_t o0 = j->offsets[2 * i], o1 = j->offsets[2 * i + 1], o2 = j->offsets[2 * i + 2];
            oracle_align_pair(j->seqs + o0, o1 - o0, j->seqs + o1, o2 - o1, j->k, j->score, j->d,
                              j->e, j->cv, 0, 3, &r);
            j->status[i] = r.status; j->scores[i] = r.score; j->aln_len[i] = r.aln_len;
            j->dp_cells[i] = r.dp_cells; j->n_anchors[i] = r.n_anchors;
            if (j->diff) {
                int32_t df = 0;
                if (j->chk_aln_off && r.status == ST_OK) {
                    const int64_t a = j->chk_aln_off[i], len = j->chk_aln_off[i + 1] - a;
                    if (len != r.aln_len || memcmp(j->chk_aln0 + a, r.aln0, (size_t)len) != 0 ||
                        memcmp(j->chk_aln1 + a, r.aln1, (size_t)len) != 0)
                        df |= 1;
                }
                if (j->chk_run_off && (r.status == ST_OK || r.status == ST_CYCLE)) {
                    /* runs (a, b, len) expand to anchors (a+u, b+u), u < len, ascending b */
                    int64_t t = 0;
                    int bad = 0;
                    for (int64_t q = j->chk_run_off[i]; q < j->chk_run_off[i + 1] && !bad; ++q) {
                        const int32_t a = j->chk_runs[3 * q], b = j->chk_runs[3 * q + 1], len = j->chk_runs[3 * q + 2];
                        for (int32_t u = 0; u < len; ++u, ++t)
                            if (t >= r.n_anchors || r.anchors[2 * t] != a + u || r.anchors[2 * t + 1] != b + u) { bad = 1; break; }
                    }
                    if (bad || t != r.n_anchors) df |= 2;
                }
                j->diff[i] = df;
            }
            oracle_free_pair(&r);
        }
    }
    return NULL;
}

static void run_batch(batch_job *job, int32_t threads) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t *tid = malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    for (int t = 0; t < threads - 1; ++t)
        if (pthread_create(&tid[started], NULL, batch_worker, job) == 0) ++started;
    batch_worker(job);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
    free(tid);
}

void oracle_align_batch(const uint8_t *seqs, const int64_t *offsets, int64_t P, int32_t k,
                        const int32_t score[16], int32_t d, int32_t e, const oracle_conv *cv,
                        int32_t threads, int32_t *status, int64_t *scores, int64_t *aln_len,
                        int64_t *dp_cells, int64_t *n_anchors) {
    batch_job job = {seqs, offsets, P, k, score, d, e, cv, status, scores, aln_len, dp_cells, n_anchors,
                     NULL, NULL, NULL, NULL, NULL, NULL, 0};
    run_batch(&job, threads);
}

/* Same batch, and every pair's gapped rows / anchors are compared with the ones another
 * implementation produced (any of the chk_* groups may be NULL).  diff[P] receives the flags. */
void oracle_check_batch(const uint8_t *seqs, const int64_t *offsets, int64_t P, int32_t k,
                        const int32_t score[16], int32_t d, int32_t e, const oracle_conv *cv,
                        int32_t threads, int32_t *status, int64_t *scores, int64_t *aln_len,
                        int64_t *dp_cells, int64_t *n_anchors, const int64_t *chk_aln_off,
                        const uint8_t *chk_aln0, const uint8_t *chk_aln1, const int64_t *chk_run_off,
                        const int32_t *chk_runs, int32_t *diff) {
    batch_job job = {seqs, offsets, P, k, score, d, e, cv, status, scores, aln_len, dp_cells, n_anchors,
                     chk_aln_off, chk_aln0, chk_aln1, chk_run_off, chk_runs, diff, 0};
    run_batch(&job, threads);
}
