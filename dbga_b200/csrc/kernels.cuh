// Launchers of every kernel of the library (one translation unit per stage).
#pragma once
#include "common.cuh"

namespace dbga {

// per-pair placement for the global-memory (large) stage-1 path, one entry per list item
struct LargeDesc {
    int64_t tbl;     // first slot of this pair's table
    int64_t pk0, pk1; // word offsets of the packed sequences
    uint32_t mask;   // capacity - 1
    uint32_t pad;
};

// device-side bump allocators and work-list counters (one struct per context)
struct Counters {
    unsigned long long runs_used;      // in runs
    unsigned long long slots_used;     // in slots
    unsigned long long tb_used;        // bytes of traceback scratch
    unsigned int n_cls[NUM_CLS];       // work-list lengths per DP class
    unsigned int overflow;             // a pool was exhausted (host grows and re-runs)
    unsigned long long cells;          // sum of DP cells
    unsigned long long nseg;           // DP segments
    unsigned long long total_runs;
    unsigned long long total_aln;
    unsigned long long total_tiles;    // copy tiles of the assembler
    unsigned int scan_ticket;          // assembler's layout pass: next block of pairs
    unsigned int long_max_m;           // most rows of any long (CLS_CTA) segment
    unsigned long long long_ticket;    // strip kernel: next (strip, segment) work item
    unsigned int strip_timeout;        // strip kernel: a boundary entry never arrived (reported as an error, never a hang)
    unsigned int pad2;
};

struct Pools {
    int32_t *runs;   int64_t runs_cap;     // (a, b, len) triples
    Slot *slots;     int64_t slots_cap;
    int64_t tb_cap;                        // bytes
    int32_t *lists;                        // NUM_CLS work lists of list_cap slot ids each
    int64_t list_cap;
    int32_t tiny_bound[NUM_TINY];          // class c holds segments with max(m, n) <= tiny_bound[c]
    int64_t warp_max_cells;                // larger segments get a whole CTA
    int32_t strip_ok, strip_per_step, strip_ulimit;   // packed strip kernel available: what the one-pass packed kernel
                                           // cannot hold joins the long class instead of the 32-bit warp kernel
};

// ---- stages 1+2 fused (pairs whose table fits in shared memory)
size_t stage12_small_smem(int key_bytes, uint32_t cap, int words, int maxq, bool medium);
cudaError_t launch_stage12_small(const uint8_t *seqs, const PairDesc *pairs, const int32_t *list, int count, int k,
                                 uint32_t cap, bool medium, int parts, int words, int maxq, bool full, bool want_segments,
                                 PairOut *pout, Pools pools, Counters *ctr, int max_abs_score, int go_raw, int ge_raw,
                                 uint8_t *nd0, uint8_t *nd1, int32_t *status, int32_t *aof, uint16_t *memo_g, int num_sms,
                                 cudaStream_t st);
// ---- stage 1, global-memory table (large pairs)
cudaError_t launch_stage1_large(const uint8_t *seqs, const PairDesc *pairs, const LargeDesc *ld,
                                const int32_t *list, int count, int max_n, int k, bool full,
                                uint32_t *packed, void *table, size_t table_bytes, int32_t *status,
                                int32_t *aof, uint8_t *nd0, uint8_t *nd1, cudaStream_t st, int *launches);

// ---- stage 2
// pair_list == nullptr: all pairs 0..num_pairs-1
cudaError_t launch_stage2(const PairDesc *pairs, const int32_t *pair_list, int64_t num_pairs, int k, bool no_anchors, bool want_segments,
                          const int32_t *aof, int32_t *status, PairOut *out, Pools pools, Counters *ctr,
                          int threads, int num_sms, int max_abs_score, int go_raw, int ge_raw, int64_t cluster_pairs,
                          cudaStream_t st);
constexpr int S2_CLUSTER_ABOVE = 65536;   // seq-2 positions beyond which a pair's stage 2 runs on a thread-block cluster

// ---- stage 3
// count_dev: device pointer to the number of list entries (no host round trip needed)
// thread-per-segment kernel over the tiny classes c0..c1-1 (one launch)
struct TinyBounds { int b[NUM_TINY]; };
size_t dp_thread_scratch_bytes(int bound, int num_sms);
cudaError_t launch_dp_thread(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *lists,
                             int64_t list_cap, const unsigned int *counts_dev, int c0, int c1, TinyBounds tb,
                             DpParams prm, uint8_t *strbuf, uint32_t *tb_scratch, int num_sms, cudaStream_t st);
// bnd: per-CTA double-buffered boundary rows, int4 per column, bnd_stride columns each
cudaError_t launch_dp_wavefront(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                                const unsigned int *count_dev, int grid, int cls, DpParams prm,
                                uint8_t *strbuf, uint8_t *tb, int4 *bnd, int64_t bnd_stride, cudaStream_t st);
// packed 16-bit wavefront over the same class list (takes the segments wf16_ok accepts)
cudaError_t launch_dp_wavefront16(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                                  const unsigned int *count_dev, int grid, int cls, DpParams prm, uint8_t *strbuf,
                                  uint8_t *tb, cudaStream_t st, int *launches);
// traceback of everything the packed kernels processed in classes cls_lo..cls_hi, one launch
cudaError_t launch_dp_traceback16(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *lists,
                                  int64_t list_cap, const unsigned int *counts_dev, int cls_lo, int cls_hi, int grid, DpParams prm,
                                  uint8_t *strbuf, const uint8_t *tb, cudaStream_t st);
// packed row-strip kernel over the long class (prm.h.strip); traceback by launch_dp_traceback16
cudaError_t launch_dp_strips16(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                               const unsigned int *count_dev, Counters *ctr, int grid, DpParams prm, uint8_t *tb,
                               uint32_t nonce, cudaStream_t st);
cudaError_t launch_dp_long(const uint8_t *seqs, const PairDesc *pairs, Slot *slots, const int32_t *list,
                           const unsigned int *count_dev, int grid_clusters, int grid_ctas, unsigned int few, DpParams prm,
                           uint8_t *strbuf, uint8_t *tb, int4 *bnd, int64_t bnd_stride, cudaStream_t st);

// ---- assembly
// per-pair result columns (the context's own arrays, or a pipeline lane's window into its parent's)
struct AsmOut {
    int32_t *status; int64_t *score, *cells; int32_t *nseg; int64_t *nanch, *aln_off, *run_off; int32_t *sop;
};
// one record per block of pairs of the layout pass: block aggregate / inclusive prefix of
// (aligned columns, reported runs, copy tiles) + flag (0 none, 1 aggregate, 2 prefix)
struct AsmScanRec { unsigned long long agg[3]; unsigned long long pre[3]; unsigned int flag; unsigned int pad[3]; };
inline int64_t asm_scan_blocks(int64_t num_pairs) { return (num_pairs + 63) / 64; }
// everything the copy pass needs to know about one tile of ASM_TILE_COLS output columns of one pair
struct AsmTile {
    int64_t a_off;        // first output byte of the pair's rows
    int64_t item_base;    // the pair's items start at item_off / item_src + item_base
    int64_t run_base, run_off;   // the pair's runs in the pool / in the compacted output
    int32_t aln_len, n_items, row_len, tp;   // tp: index of the tile inside the pair
    int32_t pair, R, n_slots, ok;
};
static_assert(sizeof(AsmTile) == 64, "AsmTile is four 16-byte words");
cudaError_t launch_assemble(const uint8_t *seqs, const PairDesc *pairs, int64_t num_pairs, const PairOut *pout,
                            const int32_t *runs, const Slot *slots, const uint8_t *strbuf, int32_t *item_off,
                            long long *item_src, AsmScanRec *scan, AsmTile *tiles, AsmOut out,
                            uint8_t *aln0, uint8_t *aln1, int32_t *runs_out, int32_t *seg_cols, Counters *ctr, int64_t seq_bytes,
                            int flags, int num_sms, cudaStream_t st, int *launches);
// copy tiles of ASM_TILE_COLS output columns: a batch of P pairs and B input bytes has at most B / ASM_TILE_COLS + P of them
constexpr int ASM_TILE_COLS = 2048;

cudaError_t launch_publish(const void *src, void *dst_host, int bytes, cudaStream_t st);
cudaError_t launch_clear(Counters *ctr, int32_t *status, int64_t n, AsmScanRec *scan, int64_t nscan, cudaStream_t st);

// ---- k-mer set statistics (|K0|, |K1|, |K0 & K1| per pair), small pairs only
size_t kmer_stats_smem(int key_bytes, uint32_t cap, int words);
cudaError_t launch_kmer_stats(const uint8_t *seqs, const int64_t *offsets, int64_t off_base, int64_t count, int k,
                              uint32_t cap, int words, int64_t max_keys, int64_t max_n, int32_t *status, int32_t *out,
                              int num_sms, cudaStream_t st);

cudaError_t launch_kmer_stats_large(const uint8_t *seq0, int64_t n0, const uint8_t *seq1, int64_t n1, int k,
                                    unsigned long long *keys, uint32_t *flag, uint64_t cap, unsigned int *counts,
                                    int num_sms, cudaStream_t st);

// ---- 2-bit packed input -> the ASCII layout the kernels read (pack.cu)
cudaError_t launch_unpack(const uint32_t *packed, const PairDesc *pairs, int64_t num_pairs, int64_t base,
                          int64_t first_seq, int64_t word_base, uint8_t *seqs, int num_sms, cudaStream_t st);

// ---- integer-pipe microbenchmark
cudaError_t launch_int_peak(int which, int iters, int *sink, int num_sms, cudaStream_t st);

}  // namespace dbga
