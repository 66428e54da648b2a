/*
 * include/ppp_b200.h -- C ABI of libppp_b200.so, the B200-native PPP scan.
 *
 * The reference (malaybasu/ProPhylo) has no FFI/plugin interface; its seam for this path is the Perl method
 *     PhyloProf::Algorithm::ppp->new(-prof1,-prof2,[-prob],[-rank],[-taxonomy],[-dup])->get_match_score()
 * (lib/PhyloProf/Algorithm/ppp.pm:58-110, 130-245) called per (query, target) pair by bin/ppp.pl:559-582,
 * bin/ppp2d.pl:632, bin/ppp_cross_score.pl:681-699 and bin/ppp_hmmer.pl:222.  These entry points are what an
 * XS binding of that method binds (perl/PPPB200.xs; INTEGRATION.md shows the stub): plain pointers and sizes,
 * no C++ or torch types, caller-owned buffers, negative return codes instead of exceptions.
 *
 * Semantics fixed by the reference and therefore by this ABI:
 *   - one *pair* = one get_match_score evaluation -> (score, yes, number, depth)           ppp.pm:244
 *   - score starts at 1.0; a step improves only on strict `<`; the first minimum wins       ppp.pm:228-237
 *   - score = Math::Cephes::bdtrc(yes-1, number, p), incl. exact 0.0 underflow, 1-MACHEP saturation,
 *     "domain error" -> 0.0 for p outside [0,1]                                            ppp.pm:225
 *   - a pair with no common yes-genome returns (1, 0, 0, 0)
 *   - p is taken verbatim (the caller evaluates ppp.pm:138-142: -prob, else yes/total)
 *   - depth is the 1-based position in the target's list (raw position incl. skipped duplicates for ranked
 *     targets, BlastResult.pm:192-237; popcount prefix for bit-plane targets)
 * yes, number and depth are bit-exact; score agrees with the reference within 1e-9 relative on -log(score)
 * (bit-identical wherever the exact tier ran, see prophylo_b200/csrc/cephes_exact.cuh).
 *
 * There is no CPU fallback: every call fails with PPP_ERR_CUDA when no sm_100 device is usable.
 * Threading: a ppp_ctx is owned by one thread at a time.  Create it AFTER any fork() (bin/ppp.pl:910 forks).
 */
#ifndef PPP_B200_H
#define PPP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPP_B200_ABI_VERSION 2

typedef struct ppp_ctx ppp_ctx;

enum {
    PPP_OK = 0,
    PPP_ERR_INVALID = -1,  /* bad argument */
    PPP_ERR_CUDA = -2,     /* CUDA runtime / no usable device */
    PPP_ERR_NOMEM = -3,    /* host or device allocation failed */
    PPP_ERR_STATE = -4,    /* call order (e.g. score before set_targets) */
    PPP_ERR_CAPACITY = -5  /* caller's output buffer too small; *nout holds the needed count */
};

/* Result filter = what bin/ppp.pl:473-506 / bin/ppp_cutoff.pl:111-162 keep of one query's row. */
enum {
    PPP_FILTER_ALL = 0,          /* every pair, dense, out[q * nT + t]                                   */
    PPP_FILTER_THRESH_LOG10 = 1, /* pairs with log10(score) < thresh  (ppp_cutoff.pl:122,154: thresh=-0.0005
                                    keeps exactly the rows whose "%.3f" log-score is < 0)                */
    PPP_FILTER_TOPK = 2          /* the k smallest scores per query; ties by target index                */
};

typedef struct ppp_filter {
    uint32_t mode;
    uint32_t k;     /* PPP_FILTER_TOPK */
    double thresh;  /* PPP_FILTER_THRESH_LOG10 */
} ppp_filter;

/* One scored pair: the list get_match_score returns (ppp.pm:244), plus its coordinates. 32 bytes. */
typedef struct ppp_hit {
    uint32_t q;      /* query index  */
    uint32_t t;      /* target index */
    double score;    /* $score: best binomial upper-tail probability           */
    uint32_t yes;    /* $best_yes                                              */
    uint32_t number; /* $best_number                                           */
    uint32_t depth;  /* $best_depth: cut-off index, 1-based; 0 = no improvement */
    float margin;    /* ln(second-best) - ln(best) over the sweep (diagnostic; +inf if unique candidate;
                        0 where the exact tier decided)                          */
} ppp_hit;

typedef struct ppp_stats {
    uint64_t pairs;            /* pairs covered by the last run                                   */
    uint64_t candidates;       /* yes-increment steps examined (popc(T&Y) summed)                 */
    uint64_t accurate_evals;   /* FP64 series evaluations (tier B)                                */
    uint64_t exact_pairs;      /* pairs re-swept by the exact Cephes tier (tier C)                */
    uint64_t kernel_launches;  /* kernels launched by the last run                                */
    float sweep_kernel_ms;     /* CUDA-event time of the dominant sweep kernel(s), on the ctx stream */
    float total_device_ms;     /* CUDA-event time of the whole run on the ctx stream              */
    uint32_t sweep_launches;   /* number of sweep kernel launches in sweep_kernel_ms              */
    uint32_t reserved;
} ppp_stats;

/* Lifetime.  devices/ndev: CUDA ordinals this context may use (NULL / 0: device 0).
 *   ndev == 1  one GPU, one stream: the form a one-process-per-GPU launcher (torchrun + prophylo_b200/parallel.py) uses.
 *   ndev  > 1  ONE process drives ndev GPUs (replaces the reference's own fan-out, bin/ppp.pl:374-421, 895-922: chunk files
 *              of 50 targets and forked `ppp.pl --client` workers): the TARGET axis is sharded contiguously over the devices,
 *              every device keeps its shard resident and receives all queries, the devices score concurrently (one host
 *              thread and stream each, no exchange), and results are assembled in the caller's buffers exactly as a single
 *              device would return them -- dense rows by strided copies, threshold hits by a per-query merge, top-k lists
 *              merged on the devices by query slices over NVLink peer copies.  ppp_printed_cutoffs and ppp_topk_device need
 *              a single-device context.  A device may be listed more than once (each entry owns a shard).
 * Returns PPP_OK or a negative code; *out is NULL on failure. */
int ppp_init(ppp_ctx **out, const int *devices, int ndev);
void ppp_destroy(ppp_ctx *ctx);
/* ctx-owned string describing the last failure on ctx (valid until the next call); ctx may be NULL for
 * failures of ppp_init. */
const char *ppp_last_error(const ppp_ctx *ctx);
int ppp_abi_version(void);

/* Targets as bit planes in canonical genome order (replaces the -prof2 ARRAY of a target whose list is its set
 * genomes in universe order).  T is row-major, nT rows of ppp_words_per_row(G) uint32 words,
 * genome g = bit (g & 31) of word (g >> 5), padding bits zero.  The data is copied before returning. */
int ppp_set_targets_bits(ppp_ctx *ctx, const uint32_t *T, size_t nT, uint32_t G);
/* The same upload, streamed: returns without copying.  The next ppp_run / ppp_score queues the copies (pieces of 4, 16, then
 * 32 MiB on a copy stream of the context) at its first launch that reads targets and scans the first rows while the later
 * pieces are still crossing PCIe: every launch waits only for the pieces it reads.  T should be page-locked (cudaHostAlloc / cudaHostRegister; pageable memory works
 * but copies synchronously) and MUST stay valid and unchanged until a ppp_run / ppp_score over these targets has returned PPP_OK
 * or ppp_wait_targets() has returned.  No counterpart in the reference (its targets are Perl values in the worker's memory):
 * this is the hook for a host that keeps its packed target collection in pinned memory and re-scores it many times. */
int ppp_set_targets_bits_streamed(ppp_ctx *ctx, const uint32_t *T, size_t nT, uint32_t G);
/* blocks until a streamed upload is complete on the device (no-op otherwise) */
int ppp_wait_targets(ppp_ctx *ctx);
/* The same planes, already in DEVICE memory (of the context's GPU or of a peer): copied device to device before returning; all
 * work that produced dT must be complete (synchronise the producing stream first).  No counterpart in the reference.  This is
 * the hook for a query-partitioned N-GPU job, where every GPU needs ALL targets: each process uploads 1/N of the planes over
 * PCIe, the processes exchange their slices over NVLink (NCCL all-gather), and the gathered buffer is handed over here --
 * 1/N of the host-to-device traffic of N full uploads (bench.py's end-to-end leg at N > 1). */
int ppp_set_targets_bits_device(ppp_ctx *ctx, const uint32_t *dT, size_t nT, uint32_t G);

/* Targets as ordered hit lists (replaces -prof2 built by BlastResult::get_profile, BlastResult.pm:192-237, or
 * the value-sorted HASH of HMMERResult, ppp.pm:147-149).  For target t, entries row_off[t]..row_off[t+1]) give,
 * in list order and ALREADY de-duplicated by the host (ppp.pm:189-192), the genome index idx[] (0..G-1 in the
 * query universe; 0xFFFF = taxon foreign to the universe) and the 1-based raw list position rawdepth[].
 * At most 65535 list positions per target.  For -dup the host keeps the list's first repeated entry and ends the list
 * before the second (ppp_read_bla_flags / PhyloProf::B200 / prophylo_b200.packer) and sets the "dup" option. */
int ppp_set_targets_ranked(ppp_ctx *ctx, const uint16_t *idx, const uint16_t *rawdepth, const uint64_t *row_off,
                           size_t nT, uint32_t G);

/* Queries: presence plane P and yes plane Y (Y subset of P; same row layout as T) and p per query
 * (replaces -prof1 and -prob).  Copied before returning; stays resident until replaced. */
int ppp_set_queries(ppp_ctx *ctx, const uint32_t *P, const uint32_t *Y, const double *p, size_t nQ);

/* Score all resident queries against all resident targets; results stay on the device. */
int ppp_run(ppp_ctx *ctx, const ppp_filter *f);
/* Number of hits the last run produced, and copy them out (ALL: nQ*nT dense; otherwise grouped by query,
 * ascending score then target index). */
int ppp_result_count(ppp_ctx *ctx, size_t *nout);
int ppp_fetch(ppp_ctx *ctx, ppp_hit *out, size_t cap, size_t *nout);

/* One-call form with host buffers: ppp_set_queries + ppp_run + ppp_fetch. */
int ppp_score(ppp_ctx *ctx, const uint32_t *P, const uint32_t *Y, const double *p, size_t nQ, const ppp_filter *f,
              ppp_hit *out, size_t cap, size_t *nout);

/* Multi-GPU global top-k (the path's only exchange step).  ppp_topk_device exposes the DEVICE
 * pointers of the last top-k run ([nQ][k] hits, [nQ] counts; valid until the next run) so that the host can
 * all-gather them over NCCL without a round trip through host memory; ppp_merge_topk_device merges `nlists`
 * gathered lists (device memory, list l at lists + l*nQ*k, counts at counts + l*nQ), adds t_offsets[l] to the target
 * indices of list l and writes the global [nQ][k] lists and [nQ] counts to HOST buffers. */
int ppp_topk_device(ppp_ctx *ctx, const ppp_hit **lists, const uint32_t **counts, uint32_t *k, size_t *nQ);
int ppp_merge_topk_device(ppp_ctx *ctx, const ppp_hit *lists, const uint32_t *counts, uint32_t nlists,
                          const uint32_t *t_offsets, size_t nQ, uint32_t k, ppp_hit *out, uint32_t *out_counts);

/* Printed scores and slope cut-offs of the last PPP_FILTER_THRESH_LOG10 run, computed on the device from the ranked hits
 * (replaces the per-query Perl sort + scan of bin/ppp.pl:473-506 and bin/ppp_cutoff.pl:130-148).  Host buffers:
 *   row_off[nQ + 1]   rows of query q = hits row_off[q] .. row_off[q + 1] of ppp_fetch's order (may be NULL)
 *   milli[n hits]     printed score of each row, sprintf("%.3f", log($score) / log(10)), in thousandths (may be NULL)
 *   marker[nQ]        ppp.pl -m <slope>: index within the query's rows of the row the marker line precedes, -1 = none among
 *                     the rows present (rows the filter dropped would print 0.000 and follow them)
 *   cutoff_milli[nQ]  ppp_cutoff.pl -s <slope>: CUTOFF in thousandths, INT32_MIN = none
 *   status[nQ]        PPP_CUT_OK, or why the caller must treat the query itself: PPP_CUT_DOUBT (a printed score lies within
 *                     1e-6 of a rounding boundary, so the device's log cannot guarantee the reference's digit),
 *                     PPP_CUT_LOG0 (a score is 0: the reference dies "Can't take log of 0"), PPP_CUT_DIV0 (the best row
 *                     prints 0.000: ppp.pl -m dies dividing by it; cutoff_milli is still valid).
 * Needs a threshold run whose hits were ranked on the device (one launch); otherwise PPP_ERR_STATE. */
enum { PPP_CUT_OK = 0, PPP_CUT_DOUBT = 1, PPP_CUT_DIV0 = 2, PPP_CUT_LOG0 = 3 };
int ppp_printed_cutoffs(ppp_ctx *ctx, double slope, uint64_t *row_off, int32_t *milli, int32_t *marker, int32_t *cutoff_milli,
                        uint8_t *status);

int ppp_get_stats(const ppp_ctx *ctx, ppp_stats *out);
/* Shape of the resident data, for bindings that must size or check caller buffers: number of resident queries (what
 * ppp_printed_cutoffs writes per-query entries for) and words per plane row of the resident universe (0 before any
 * ppp_set_targets_* call). */
int ppp_query_count(const ppp_ctx *ctx, size_t *nQ);
int ppp_row_words(const ppp_ctx *ctx, size_t *words);

/* Options (tests / diagnostics): "force_exact" (0/1: every pair through the exact tier),
 * "check_bounds" (0/1: verify the fast tier's enclosures against the accurate tier, count violations),
 * "sweep_warps" (warps per CTA of the tile and refine kernels, default 16), "table_mode" (-1 auto / 0 off / 1 on: exact (yes, number) table per query for dense
 * runs of few queries against many targets), "dup" (0/1: the resident ranked lists were packed for -dup, see
 * ppp_read_bla_flags), "device_rank" (1/0: threshold results ranked on the device / on the host).
 * Top-k schedule (none of them changes a result, only how many pairs need a full evaluation): "speculate" (0/1: screen the
 * bulk of the targets against a speculative, verified threshold), "spec_first" (0/1: the first chunk as well, against the
 * seeding pass's bounds), "respec_mult" (a new, tighter speculation level every time the processed targets have grown this
 * many times, in runs of at least 2^"respec_min_pairs_log2" pairs; <= 1: one level), "spec_eps_ppm" (probability, in ppm, with which a speculative threshold may be too tight and
 * the query is re-scanned), "seed_mult" / "chunk0_mult" / "chunk1_mult" (targets of the seeding pass, of the first chunk's
 * head and of the first chunk, in units of k), "pf_overlap" (0/1: the pre-filter's class launches of a chunk on concurrent streams), "prefilter_tc" (bit 0: the chunk screen of the queries
 * with an all-ones presence plane takes its counts from tcgen05.mma instead of POPC, bit 1: that of the other queries too;
 * prophylo_b200/csrc/ppp_prefilter_tc.cu), "launch_pairs_log2" (pairs per launch of a
 * filtered run: 0 = automatic, the largest of 2^31 / 2^30 / 2^29 whose scratch buffers -- 40 bytes per pair -- stay below half of
 * the free device memory; or 12..31; tests lower it to exercise runs of several launches).  "spec_eps_ppm" is scaled by
 * min(1, 131072 / targets): a failed speculation costs a second scan of all targets.
 * Unknown names return PPP_ERR_INVALID. */
int ppp_set_option(ppp_ctx *ctx, const char *name, int64_t value);
/* Counters of the last run (diagnostics, bench.py): "candidates", "accurate_evals", "exact_pairs", "full_pairs",
 * "survivors", "prefilter_pairs", "bound_checks", "bound_violations", "spec_rank", "spec_failed", "used_table", "num_sms", "tc_launches",
 * and the CUDA-event times "us_sweep", "us_prefilter", "us_exact", "us_staircase", "us_merge" (microseconds). */
int ppp_get_counter(const ppp_ctx *ctx, const char *name, int64_t *value);

/* Row length in uint32 words of the bit-plane layout for G genomes: ceil(G/128)*4. */
uint32_t ppp_words_per_row(uint32_t G);

/* Host-side reader of the reference's per-gene BLAST caches (`.bla`: query, subject, e-value, subject taxon; written by
 * util/cache_blast_result.pl:86 and bin/ppp.pl:1143).  Replaces the per-file Perl parse of bin/ppp.pl:606-673
 * (get_profile) + BlastResult::get_profile (lib/PhyloProf/BlastResult.pm:192-237) for a whole genome's worth of files:
 * subjects in first-occurrence order, a subject's last taxon wins, Perl-false taxa ("" / "0") dropped, then taxa
 * de-duplicated with their raw 1-based list positions kept -- i.e. exactly the arrays ppp_set_targets_ranked() takes,
 * one target per path.  universe[g] = taxon string of genome index g (taxa outside it map to 0xFFFF).  Needs no GPU and
 * no ppp_ctx; nthreads <= 0 = all host threads.  On failure returns a negative code and a message in err.  The arrays
 * are malloc'ed by the library: release them with ppp_free_ranked_lists(). */
typedef struct ppp_ranked_lists {
    uint16_t *idx;      /* [n_entries] genome index, 0xFFFF = foreign taxon */
    uint16_t *rawdepth; /* [n_entries] 1-based raw list position */
    uint64_t *row_off;  /* [n_targets + 1] */
    size_t n_targets, n_entries;
} ppp_ranked_lists;
int ppp_read_bla(const char *const *paths, size_t n_paths, const char *const *universe, size_t G, int nthreads,
                 ppp_ranked_lists *out, char *err, size_t errlen);
/* Same with flags.  PPP_BLA_DUP packs the lists for the reference's -dup flag (bin/ppp.pl:282 -> ppp.pm:189-200, whose
 * `$seen{key}` is ONE counter for the whole list): the first repeated taxon of a list is kept as a second entry (counted
 * again), the list ends before the second repeated entry (`yes` is frozen from there on, so no later step can improve).
 * Score such lists with ppp_set_option(ctx, "dup", 1): the sweep then drops the candidate that would exceed |Y|
 * (`last if $yes > $max_yes`, ppp.pm:222-224). */
#define PPP_BLA_DUP 1u
int ppp_read_bla_flags(const char *const *paths, size_t n_paths, const char *const *universe, size_t G, int nthreads,
                       uint32_t flags, ppp_ranked_lists *out, char *err, size_t errlen);
void ppp_free_ranked_lists(ppp_ranked_lists *lists);
/* ppp_read_bla_flags with a persistent packed cache: the parsed lists of the whole path set are kept in
 * ONE binary file `cache_path`, keyed by the flags, the universe and every path's size and modification time.  A call whose key
 * matches reads the arrays back without opening a single `.bla` (*from_cache = 1); otherwise the files are parsed and the cache
 * is (re)written atomically.  cache_path NULL or "" = no cache.  A cache that cannot be written is not an error. */
int ppp_read_bla_cached(const char *const *paths, size_t n_paths, const char *const *universe, size_t G, int nthreads, uint32_t flags,
                        const char *cache_path, ppp_ranked_lists *out, int *from_cache, char *err, size_t errlen);

/* Host-side reader and bit-plane packer of a FAMILY of profile files (2 tab columns taxon, 0/1; '#' comments; optional header
 * line: lib/PhyloProf/Profile.pm:433-514, written by bin/hmm3profile.pl:175-181 / bin/blast2profile.pl): one shared universe --
 * the union of the files' taxa, numeric ids ascending (data/taxdis.txt order), other names after them -- and the packed presence
 * (P) and yes (Y) planes of every profile in the row layout of ppp_set_queries / ppp_set_targets_bits, with p = yes / total over
 * the file's own taxa (ppp.pm:138-142).  Replaces per-file Perl parsing + one hash probe per taxon for TIGRFAM-scale families
 * (BASELINE config 5).  Needs no GPU; nthreads <= 0 = all host threads; at most 8192 distinct taxa.  Release with
 * ppp_free_profile_set(). */
typedef struct ppp_profile_set {
    char *universe_buf;   /* G NUL-terminated taxon strings, back to back */
    size_t *universe_off; /* [G] offset of genome g's taxon string in universe_buf */
    uint32_t *P, *Y;      /* [n][wpr] presence / yes planes */
    double *p;            /* [n] yes / total */
    uint32_t *yes, *total; /* [n] Profile::get_yes / get_total */
    size_t n, G, wpr;
} ppp_profile_set;
int ppp_read_profiles(const char *const *paths, size_t n_paths, int header, int nthreads, ppp_profile_set *out, char *err, size_t errlen);
void ppp_free_profile_set(ppp_profile_set *set);

/* Device-side evaluation of the exact tier's bdtrc (diagnostic; parity tests of cephes_exact.cuh). */
int ppp_debug_bdtrc(ppp_ctx *ctx, const int32_t *k, const int32_t *n, const double *p, double *out, size_t m);

/* Device-side evaluation of the sweep's fast tiers at m points (yes, number, p) with 1 <= yes <= number <= 8192, 0 < p < 1
 * (diagnostic; parity tests): lo / hi = the float enclosure of ln P(X >= yes | number, p) that the staircase prune and the
 * candidate selection rely on, lnf = the FP64 series value.  The tests check lo <= ln bdtrc(yes-1, number, p) <= hi. */
int ppp_debug_enclosure(ppp_ctx *ctx, const int32_t *yes, const int32_t *number, const double *p, float *lo, float *hi, double *lnf,
                        size_t m);

/* Measured peak rate of one SM issue pipe on this device under its current clocks, in lane-operations per second over
 * the whole GPU (a microbenchmark of independent dependent-chains, a few milliseconds; prophylo_b200/csrc/ppp_debug.cu):
 * "popc" (XU pipe), "lop3", "iadd" (ALU), "dfma" (FP64).  bench.py's roofline divides by these. */
int ppp_pipe_peak(ppp_ctx *ctx, const char *pipe, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* PPP_B200_H */
