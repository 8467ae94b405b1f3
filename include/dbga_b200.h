/* dbga_b200 -- C ABI of the B200-native DBGA pairwise alignment hot path.
 *
 * The reference (xingjianleng/DBGA) is pure Python and has no FFI; these entry points are
 * what a maintainer would bind from `src/dbga/debruijn_pairwise.py` to replace, for a
 * batch of independent sequence pairs:
 *
 *   stages 1-2  DeBruijnPairwise.__init__            src/dbga/debruijn_pairwise.py:42-76
 *               (get_kmers / duplicate_kmers          src/dbga/utils.py:108-154,
 *                add_debruijn / extract_bubble        debruijn_pairwise.py:206-225, 277-300,
 *                debruijn_merge_correctness           utils.py:266-291)
 *   stage 3     DeBruijnPairwise.alignment            debruijn_pairwise.py:418-543
 *               (bubble_aln -> dna_global_aln ->      :345-395, utils.py:157-192,
 *                cogent3.align.global_pairwise         utils.py:185)
 *
 * Plain pointers and sizes only; no exceptions cross the boundary.  Per-pair outcomes are
 * status codes which the Python layer turns into the reference's exact exceptions
 * (INTEGRATION.md).  All calls on one context are synchronous on that context's stream;
 * use one context per GPU / per host thread.
 */
#ifndef DBGA_B200_H
#define DBGA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- per-pair status ---------------------------------------------------------------- */
#define DBGA_OK 0        /* aligned */
#define DBGA_CYCLE 1     /* ValueError("Cycles detected in de Bruijn graph, ...")  debruijn_pairwise.py:73-76 */
#define DBGA_K_INVALID 2 /* ValueError("Invalid k size for kmers")                 utils.py:129-130 */
#define DBGA_BAD_CHAR 3  /* a character outside ACGT (the DP alphabet, utils.py:184) */
#define DBGA_TOO_LARGE 4 /* exceeds a workspace or numeric limit of this build; never silent */

/* ---- call-level return codes -------------------------------------------------------- */
#define DBGA_SUCCESS 0
#define DBGA_ERR_CUDA (-1)
#define DBGA_ERR_ARG (-2)
#define DBGA_ERR_NOMEM (-3)

/* Tie-order encoding shared with oracle/: index into XMY, XYM, MXY, MYX, YXM, YMX
 * (first state wins ties; X consumes a seq1 character, Y a seq2 character). */
typedef struct dbga_params {
    int32_t k;            /* k-mer size >= 1 (k > a sequence's length: DBGA_K_INVALID); k >= 32 runs on
                             by-reference keys in a global-memory table (any k, slower per pair)    */
    int32_t score[16];    /* substitution scores, row-major over ACGT x ACGT            */
    int32_t gap_open;     /* d  (debruijn_pairwise.py:423)                              */
    int32_t gap_extend;   /* e  (:424)                                                  */
    int32_t gap_len_mode; /* 0: gap(L) = d + (L-1) e (default)   1: d + L e             */
    int32_t xy_allowed;   /* 0 (default): no X<->Y transitions                          */
    int32_t pred_order;   /* default 0 = X,M,Y                                          */
    int32_t end_order;    /* default 0 = X,M,Y                                          */
    int32_t stages;       /* 2: stop after stages 1-2 (constructor)   3: full alignment */
    int32_t want_graph;   /* 1: also return per-k-mer "non-duplicate" flags (node ids)  */
    int32_t flags;        /* DBGA_FLAG_* below, 0 = everything (round-1 behaviour)      */
} dbga_params;

/* dbga_params.flags: what travels back to the host (the end-to-end path is bound by the host link).
 * NO_RUNS       the anchor-run list is not returned (run_off all zero, runs NULL); n_anchors still is.
 *               DeBruijnPairwise.alignment() itself never returns anchors (debruijn_pairwise.py:543).
 * COMPACT_ROWS  aln0 / aln1 carry only the columns the DP (or an empty-side segment) produced; the
 *               columns inside a run of anchors are the caller's own input bytes and are not sent.
 *               seg_cols tells how many columns each segment has; dbga_expand_rows rebuilds the
 *               full gapped rows on the host.  Implies runs. */
#define DBGA_FLAG_NO_RUNS 1
#define DBGA_FLAG_COMPACT_ROWS 2
#define DBGA_FLAG_SOP 4          /* also count match / mismatch / gap_open / gap_extend per pair (dbga_result.sop) */

typedef struct dbga_ctx dbga_ctx;

/* Results of one batch.  Every array is host memory owned by the context (pinned) and
 * stays valid until the next dbga_align_batch / dbga_plan_fetch on the same context or
 * dbga_destroy. */
typedef struct dbga_result {
    int64_t num_pairs;
    const int32_t *status;      /* [P]                                                    */
    const int64_t *score;       /* [P] sum of DP scores of the segments handed to the DP  */
    const int64_t *dp_cells;    /* [P] sum |x|*|y| over those segments        */
    const int32_t *dp_segments; /* [P]                                                    */
    const int64_t *n_anchors;   /* [P] number of merge nodes M                            */
    const int64_t *run_off;     /* [P+1] offsets (in runs) into `runs`                    */
    const int32_t *runs;        /* [3*run_off[P]] (a, b, len): anchors (a+u, b+u), u<len,
                                   in ascending b -- the merge nodes in seq-2 order       */
    const int64_t *aln_off;     /* [P+1] byte offsets into aln0 / aln1                    */
    const uint8_t *aln0;        /* gapped seq1 rows, concatenated                         */
    const uint8_t *aln1;        /* gapped seq2 rows                                       */
    /* want_graph only: per k-mer position, 1 = not a duplicate k-mer (utils.py:148-154).
       Node ids follow by prefix sums (DESIGN.md "Node ids").                              */
    const int64_t *p_off;       /* [P+1] offsets into nondup0 (seq1 k-mer positions)      */
    const uint8_t *nondup0;
    const int64_t *q_off;       /* [P+1] offsets into nondup1 (seq2 k-mer positions)      */
    const uint8_t *nondup1;
    /* device-side timing of the last run, milliseconds (CUDA events on the ctx stream)  */
    float ms_h2d, ms_stage1, ms_stage2, ms_stage3, ms_assemble, ms_d2h, ms_total;
    int64_t dp_cells_total;
    int64_t dp_segments_total;
    int64_t kernel_launches;    /* kernels of this library launched for the batch         */
    /* [4P] match, mismatch, gap_open, gap_extend of every aligned pair: the reference's
       pairwise_alignment_score (src/dbga/utils.py:440-504), counted on the device while the rows
       are laid out (DBGA_FLAG_SOP only, NULL otherwise; zero for pairs that are not DBGA_OK)       */
    const int32_t *sop;
    /* COMPACT_ROWS only: [run_off[P] + P] columns of segment t of pair i at seg_cols[run_off[i] + i + t],
       t = 0 .. (number of runs of pair i)                                                        */
    const int32_t *seg_cols;
    float ms_plan;              /* host time spent planning (classing, metadata), milliseconds    */
} dbga_result;

/* Context: one CUDA device, one stream, grow-only workspaces. */
int dbga_create(int device, dbga_ctx **out);
void dbga_destroy(dbga_ctx *ctx);
const char *dbga_last_error(dbga_ctx *ctx);
/* Use an existing stream (e.g. torch's current stream) instead of the context's own. */
int dbga_set_stream(dbga_ctx *ctx, void *cuda_stream);
/* Upper bound on device scratch for traceback matrices, bytes (default 8 GiB). */
int dbga_set_scratch_limit(dbga_ctx *ctx, int64_t bytes);

/* Pinned host buffers for callers that want full-speed host<->device copies. */
void *dbga_host_alloc(int64_t bytes);
void dbga_host_free(void *p);

/* One call = the reference's  [DeBruijnPairwise(pair, k).alignment(...) for pair in batch].
 * seqs    : concatenated ASCII sequences (host memory)
 * offsets : [2P+1]; sequence j of pair i is seqs[offsets[2i+j] : offsets[2i+j+1]] */
int dbga_align_batch(dbga_ctx *ctx, const uint8_t *seqs, const int64_t *offsets,
                     int64_t num_pairs, const dbga_params *params, dbga_result *out);

/* 2-bit packed input: a quarter of the bytes on the host link.  Canonical layout of a batch of
 * num_seqs = 2P sequences described by the same offsets[] (in bases) as the ASCII entry: sequence j
 * starts at 32-bit word (offsets[j] >> 4) + j of `packed`; base i of it occupies bits 2 (i & 15) and up
 * of word i >> 4; A, C, G, T = 0, 1, 2, 3 (non-ACGT input cannot be represented: dbga_pack_ascii
 * reports it).  dbga_packed_words(offsets, 2P) is the size of the buffer in words.  Results are the
 * same ASCII gapped rows as dbga_align_batch. */
int64_t dbga_packed_words(const int64_t *offsets, int64_t num_seqs);
/* ASCII -> packed on the host (threads <= 0: all cores); returns the number of sequences that held a
 * character outside ACGT (packed as A), or -1 on a null argument. */
int64_t dbga_pack_ascii(const uint8_t *seqs, const int64_t *offsets, int64_t num_seqs, uint32_t *packed, int threads);
int dbga_align_batch_packed(dbga_ctx *ctx, const uint32_t *packed, const int64_t *offsets, int64_t num_pairs,
                            const dbga_params *params, dbga_result *out);

/* One call, several GPUs of one box (pairs shard by index, no collective, results
 * gathered on the host): the batch is cut into num_devices contiguous shards of equal bytes, every
 * shard runs dbga_align_batch on its own context / GPU from its own host thread, and the shards'
 * results come back side by side (shard d holds pairs pair_lo[d] .. pair_lo[d+1]-1; offsets inside a
 * shard's result are shard-local).  packed != NULL: 2-bit input (seqs may then be NULL). */
#define DBGA_MAX_SHARDS 16
typedef struct dbga_multi dbga_multi;
typedef struct dbga_multi_result {
    int32_t num_shards;
    int32_t pad;
    int64_t pair_lo[DBGA_MAX_SHARDS + 1];
    dbga_result shard[DBGA_MAX_SHARDS];
    float ms_total;             /* host wall clock of the whole call */
} dbga_multi_result;
int dbga_multi_create(const int *devices, int num_devices, dbga_multi **out);
void dbga_multi_destroy(dbga_multi *m);
const char *dbga_multi_last_error(dbga_multi *m);
int dbga_align_batch_multi(dbga_multi *m, const uint8_t *seqs, const uint32_t *packed, const int64_t *offsets,
                           int64_t num_pairs, const dbga_params *params, dbga_multi_result *out);

/* FASTA on the wire (reference: load_sequences -> cogent3.load_unaligned_seqs, src/dbga/utils.py:95-96;
 * output aln.to_fasta(), src/dbga/cli.py:135-136, rows in blocks of 60 columns).  dbga_fasta_read parses a
 * file of R records straight into the batch layout: upper-cased bases without line breaks in one pinned
 * buffer + offsets[R+1] (records 2i and 2i+1 are pair i), the labels (first word of each header line), and,
 * with want_packed, the canonical 2-bit words of the same records (bad_records counts those that hold a
 * character outside ACGT: use the ASCII buffer then).  dbga_fasta_write writes the two wrapped records of every
 * DBGA_OK pair of a result whose pair 0 is records first_record, first_record + 1 of `in`; a COMPACT_ROWS result
 * is expanded on the fly.  Returns the number of pairs written, -1 on error.  threads <= 0: all host cores.
 * dbga_fasta_free keeps the batch's pinned buffers (up to 4 GiB in all) for the next dbga_fasta_read of the process:
 * page-locking costs about ten times the parsing, and a tool that streams files pays it once. */
typedef struct dbga_fasta {
    int64_t num_records;
    const int64_t *offsets;     /* [R+1] into seqs                                  */
    const uint8_t *seqs;        /* pinned host memory                               */
    const uint32_t *packed;     /* pinned; NULL unless want_packed                  */
    int64_t packed_words;
    const int64_t *name_off;    /* [R+1] into names                                 */
    const char *names;
    int64_t bad_records;
} dbga_fasta;
int dbga_fasta_read(const char *path, int want_packed, int threads, dbga_fasta **out);
void dbga_fasta_free(dbga_fasta *f);
int64_t dbga_fasta_write(const char *path, const dbga_fasta *in, const dbga_result *res, int64_t first_record,
                         int width, int append, int threads);

/* Host-side expansion of a COMPACT_ROWS result into the full gapped rows (what the call returns
 * without the flag): row0 / row1 receive aln_off_full[P] bytes each, aln_off_full [P+1] the offsets.
 * seqs / offsets: the batch's input.  threads <= 0: all host cores.  Returns DBGA_SUCCESS, or
 * DBGA_ERR_ARG if a row would exceed cap bytes. */
int dbga_expand_rows(const uint8_t *seqs, const int64_t *offsets, const dbga_result *res, uint8_t *row0,
                     uint8_t *row1, int64_t cap, int64_t *aln_off_full, int threads);

/* Split form: plan (host-side sizing + metadata upload), run on device-resident
 * sequences (kernels only, nothing copied), fetch (device -> host). */
int dbga_plan(dbga_ctx *ctx, const int64_t *offsets, int64_t num_pairs, const dbga_params *params);
/* same, for dbga_global_align_batch's workload (every pair is one whole-sequence segment) */
int dbga_plan_global(dbga_ctx *ctx, const int64_t *offsets, int64_t num_pairs, const dbga_params *params);
int dbga_plan_run(dbga_ctx *ctx, const uint8_t *device_seqs);
int dbga_plan_fetch(dbga_ctx *ctx, dbga_result *out);

/* Measured integer-pipe peaks used as the DP roofline denominator:
 * out[0] = int32 add ops/s, out[1] = int32 max ops/s, out[2] = ops/s of the DP's own
 * 5 add : 4 max mix, all over the whole chip. */
int dbga_measure_int_peak(dbga_ctx *ctx, double out[3]);

/* Stand-alone stage-3 entry: P independent global alignments (the batched equivalent of
 * utils.py:157-192 dna_global_aln); same offsets convention; results in out->aln*. */
int dbga_global_align_batch(dbga_ctx *ctx, const uint8_t *seqs, const int64_t *offsets,
                            int64_t num_pairs, const dbga_params *params, dbga_result *out);

/* k-mer set statistics of P pairs (host buffers, same offsets convention, synchronous):
 * distinct0[i] = |K0|, distinct1[i] = |K1|, shared[i] = |K0 & K1| for the k-mer sets of pair i.
 * What the reference's k heuristics are computed from: predict_p / predict_final_p
 * (src/dbga/utils.py:371-413: (shared / (distinct0 + distinct1 - shared)) ** (1 / k)) and
 * kmers_larger_than_threshold (src/dbga/kmersize.py:10-14: (N - distinct) / N).
 * status[i]: DBGA_OK, DBGA_K_INVALID (k > a length), DBGA_BAD_CHAR.  Pairs whose k-mers fit a
 * shared-memory table (about 9 kb per sequence at k <= 15) share one launch; longer pairs are
 * counted one at a time, the whole GPU on a global-memory table. */
int dbga_kmer_stats(dbga_ctx *ctx, const uint8_t *seqs, const int64_t *offsets, int64_t num_pairs, int32_t k,
                    int32_t *status, int64_t *distinct0, int64_t *distinct1, int64_t *shared);

int dbga_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DBGA_B200_H */
