/*
 * prophylo_b200/csrc/ppp_capi.cu -- host side of libppp_b200.so: the C ABI of include/ppp_b200.h.
 *
 * Owns the device-resident state of one PPP scan (target planes, query planes, per-query constants, the
 * log-factorial table, hit buffers), launches the sweep kernel (ppp_sweep.cu) and the exact tier (ppp_exact.cu)
 * on one CUDA stream and times them with CUDA events on that stream.  No torch, no CPU scoring path: if the
 * device is missing every entry point fails with PPP_ERR_CUDA.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <new>
#include <thread>
#include <vector>

#include "../../include/ppp_b200.h"
#include "ppp_common.cuh"

extern "C" cudaError_t ppp_launch_sweep(const SweepParams *prm, int wpl, int nwarps, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_prefilter2(const SweepParams *prm, int wpl, int pall, const void *Tt, uint32_t ntt, const uint32_t *qorder,
                                             uint32_t nqk, uint32_t tile_q, uint32_t stair_cap, uint32_t *surv, unsigned long long *surv_count,
                                             unsigned long long surv_cap, unsigned long long *stats, int nsm, cudaStream_t st);
extern "C" uint32_t ppp_prefilter2_max_stair(int wpl, int pall);
extern "C" cudaError_t ppp_launch_prefilter_tc(const SweepParams *prm, int wpl, int pall, const void *Tt, uint32_t ntt, const uint32_t *qorder,
                                               uint32_t nqk, uint32_t stair_cap, uint32_t *surv, unsigned long long *surv_count,
                                               unsigned long long surv_cap, int exact, uint32_t *passmask, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_rescreen_masks(const SweepParams *prm, int wpl, int pall, const uint32_t *qorder, uint32_t nqk, uint32_t stair_cap,
                                                 const uint32_t *masks, uint32_t *out, unsigned long long *out_count, unsigned long long out_cap,
                                                 int nsm, cudaStream_t st);
extern "C" uint32_t ppp_rescreen_max_stair(int wpl, int pall);
extern "C" uint32_t ppp_prefilter_tc_tile_queries(int pall);
extern "C" uint32_t ppp_prefilter_tc_max_stair(int wpl, int pall);
extern "C" uint32_t ppp_prefilter2_tile_queries(int pall);
extern "C" cudaError_t ppp_launch_refine(const SweepParams *prm, int wpl, const uint32_t *surv, const unsigned long long *surv_count,
                                         unsigned long long surv_cap, int nwarps, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_transpose(const void *T, void *Tt, uint32_t nT, int wpl, cudaStream_t st);
extern "C" cudaError_t ppp_launch_transpose_rows(const void *T, void *Tt, uint32_t nT, int wpl, uint32_t t_begin, uint32_t t_end, cudaStream_t st);
extern "C" cudaError_t ppp_launch_exact(const SweepParams *prm, int W, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_sweep_ranked(const SweepParams *prm, int rwpl, int nwarps, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_exact_ranked(const SweepParams *prm, int rwpl, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_debug_bdtrc(const int32_t *k, const int32_t *n, const double *p, double *out, size_t m,
                                              cudaStream_t st);
extern "C" cudaError_t ppp_launch_staircase(const QConst *qc, const double *lf, const double *kth, int accept_mode, double thresh_ln,
                                            const uint32_t *static_off, uint16_t *nmax, uint32_t *active_off, double *last_tau,
                                            uint32_t nQ, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_topk_merge(const void *appended, uint32_t *qcnt, uint32_t qcap, void *topk, uint32_t *topk_cnt,
                                             double *kth, uint32_t k, uint32_t nQ, uint32_t t_offset, int nsm, cudaStream_t st);
extern "C" int ppp_topk_max_k(void);
extern "C" cudaError_t ppp_launch_table_build(const QConst *qc, const unsigned long long *toff, uint32_t nQ, double *tab, uint32_t *bad,
                                              int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_table_sweep(const SweepParams *prm, int wpl, const void *Tt, uint32_t ntt, const unsigned long long *toff,
                                              const double *tab, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_topk_merge_lists(const void *lists, const uint32_t *counts, uint32_t nlists, const uint32_t *t_offsets,
                                                   uint32_t nQ, uint32_t k, void *out, uint32_t *out_cnt, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_spec_init(const void *topk, const uint32_t *topk_cnt, const double *kth, const double *prev, uint32_t k,
                                            uint32_t r, double *kspec, double *kspec0, uint32_t nQ, cudaStream_t st);
extern "C" cudaError_t ppp_launch_spec_update(double *kspec, const double *kth, uint32_t nQ, cudaStream_t st);
extern "C" cudaError_t ppp_launch_spec_verify(const void *topk, const uint32_t *topk_cnt, const double *kspec0, uint32_t k, double *last_tau,
                                              uint32_t *failed, uint32_t *nfailed, uint32_t nQ, cudaStream_t st);
extern "C" void ppp_host_count_rows(const uint32_t *P, const uint32_t *Y, size_t nQ, size_t wpr, uint32_t *ny, uint32_t *np);
extern "C" cudaError_t ppp_launch_seed_kth(const void *scores, uint32_t S, uint32_t k, double *kth, uint32_t rs, double *kspec_seed, uint32_t nQ,
                                           int nsm, cudaStream_t st);
extern "C" size_t ppp_rank_scratch_words(size_t n);
extern "C" cudaError_t ppp_launch_rank_hits(const ppp_hit *rec, size_t n, uint32_t nQ, uint32_t nT, ppp_hit *out, uint32_t *scratch, int nsm,
                                            cudaStream_t st, uint32_t *launches);
extern "C" cudaError_t ppp_launch_rank_cutoffs(const ppp_hit *rec, size_t n, uint32_t nQ, double slope, unsigned long long *row_off,
                                               int32_t *milli, int32_t *marker, int32_t *cutoff_milli, uint8_t *status, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_debug_enclosure(const QConst *qc, const double *lf, const int32_t *y, const int32_t *n, float *lo, float *hi,
                                                  double *lnf, size_t m, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_pipe_peak(int kind, uint32_t iters, uint32_t *sink, int nsm, cudaStream_t st, double *lane_ops);
extern "C" cudaError_t ppp_launch_add_t(void *hits, size_t n, uint32_t off, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_accumulate(unsigned long long *totals, const unsigned long long *counters, int n, cudaStream_t st);

struct ppp_ctx {
    /* multi-device group (ppp_init with ndev > 1): this context owns one single-device context per GPU and shards the
     * TARGET axis over them; every entry point below dispatches to the group_* functions */
    std::vector<ppp_ctx *> subs;
    std::vector<size_t> sub_t0;          /* first global target index of each sub-context's shard, and nT at the end */
    uint32_t t_base = 0;                 /* (sub-context) added to the target index of every dense / threshold record */
    uint32_t group_mode = 0, group_k = 0; /* filter mode and k of the group's last run */
    std::vector<ppp_hit> group_hits;     /* threshold runs of a group: merged on the host */
    int device = 0;
    int nsm = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double phase_us[4] = {0, 0, 0, 0}; /* sweep, exact, staircase, merge */
    std::vector<cudaEvent_t> evpool; /* deferred per-chunk timing: events are read after the final synchronize */
    size_t ev_used = 0;
    struct Span { size_t a, b; int phase; }; /* phase: 0 sweep, 1 exact, 2 staircase, 3 merge, 4 prefilter */
    std::vector<Span> spans;
    unsigned long long *d_totals = nullptr; /* [CNT_N + 1] accumulated on the device across chunks */
    unsigned long long survivors = 0, prefilter_pairs = 0;
    double prefilter_us = 0;
    char err[512] = {0};

    /* targets */
    uint32_t *d_T = nullptr;
    size_t cap_T = 0;
    uint32_t *d_Tt = nullptr; /* chunk-major transposed planes for the thread-per-pair pre-filter */
    size_t cap_Tt = 0;
    bool Tt_valid = false;
    /* streamed upload (ppp_set_targets_bits_streamed): pieces of rows copied on copy_stream; a run waits for a piece -- and
     * transposes it -- right before the first launch that reads it, so the upload hides behind the scan of the earlier pieces */
    struct Piece { size_t t_end; cudaEvent_t ev; };
    std::vector<Piece> pieces;
    size_t pieces_waited = 0;
    const uint32_t *streamed_src = nullptr; /* host planes of a streamed upload whose copies are not queued yet */
    cudaStream_t copy_stream = nullptr;
    /* the pre-filter's class launches of one target chunk (all-ones presence planes / P == Y / general) are independent: the two
     * smaller ones go to side streams, so the tail of one launch -- its last items are 150 us each -- overlaps the next one */
    cudaStream_t pf_stream[2] = {nullptr, nullptr};
    cudaEvent_t pf_fork = nullptr, pf_join[2] = {nullptr, nullptr};
    int pf_overlap = 1;
    std::vector<cudaEvent_t> piece_evpool;
    uint32_t *d_surv = nullptr;
    size_t cap_surv = 0;
    uint32_t *d_surv1 = nullptr; /* pair-level column masks of the tensor-core screen ([query tile][target][words]), re-screened into d_surv */
    size_t cap_surv1 = 0;
    /* ranked targets */
    bool ranked = false;
    int rwpl = 0; /* list words per lane: 1 (<= 1024 entries) or 2 (<= 2048) */
    uint16_t *d_ridx = nullptr, *d_rraw = nullptr;
    size_t cap_ridx = 0, cap_rraw = 0;
    unsigned long long *d_roff = nullptr;
    size_t cap_roff = 0;
    int lf_wpl = 0;
    size_t nT = 0;
    uint32_t G = 0;
    int wpl = 0; /* device words per lane: device row = 32*wpl words */
    double *d_lf = nullptr;
    /* queries */
    uint32_t *d_Q = nullptr;
    QConst *d_qc = nullptr;
    void *h_stage = nullptr; /* pinned staging buffer of ppp_set_queries */
    size_t cap_stage = 0;
    uint32_t *d_qorder = nullptr; /* query ids by pre-filter class: the n_pall queries with an all-ones presence plane first, then the
                                   * n_pyy queries with P == Y (number == yes: one lookup per pair), then the others */
    size_t n_pall = 0, n_pyy = 0;
    uint32_t stair_need[3] = {0, 0, 0}; /* largest per-tile staircase footprint (entries) of the general / all-ones / P == Y launch */
    unsigned long long *d_pfstats = nullptr;
    uint32_t tile_q[3] = {0, 0, 0}; /* queries per pre-filter tile of each class: the kernel's 16 / 32 / 32 unless a tile's staircases would not fit */
    uint32_t tile_q_any = 0;        /* the same for a tile of arbitrary queries (re-scan launches) */
    uint32_t stair_need_any = 0; /* staircase footprint of the worst possible general tile (re-scan of arbitrary queries) */
    uint32_t stair_need_tc[2] = {0, 0}; /* largest staircase footprint of a tile of the tensor-core screen: general (64 queries) / all-ones (128) */
    int prefilter_tc = 0;        /* bit 0: the all-ones launch of the chunk screen runs on the tensor cores (ppp_prefilter_tc.cu); bit 1: the general one */
    int tc_exact = 1;            /* the tensor-core screen re-screens chunk-level passes exactly in the kernel (0: superset survivor list) */
    unsigned long long tc_launches = 0;
    int speculate = 1;           /* top-k: screen the bulk of the targets against a speculative, verified threshold */
    int spec_first = 1;          /* ... and the first chunk against the seeding pass's rs-th smallest bound */
    uint32_t spec_rank = 0, spec_failed = 0;
    int chunk0_mult = 4; /* head of the first chunk, in units of k (0 = no split) */
    int respec_min_pairs_log2 = 27; /* ... in runs of at least this many pairs */
    int respec_mult = 8; /* a new speculation level every time the processed targets have grown this many times (<= 1: one level) */
    int seed_mult = 4, chunk1_mult = 24, spec_eps_ppm = 5000; /* top-k schedule: seeding targets / k, first chunk / k, P(spec fails) */
    size_t nQ = 0, capQ = 0;
    /* results */
    ppp_hit *d_hits = nullptr;
    size_t cap_hits = 0;
    size_t n_results = 0;
    std::vector<ppp_hit> host_results; /* filtered modes: finalised on the host side of the ABI */
    bool results_on_host = false;
    bool results_topk = false;           /* results are the device top-k lists (d_topk, counts in topk_counts) */
    std::vector<uint32_t> topk_counts;
    ppp_hit *d_acc = nullptr; /* threshold runs of several launches: the launches' hits gathered before ranking */
    size_t cap_acc = 0;
    ppp_hit *d_ranked = nullptr; /* threshold runs: the hits in (q, score, t) order (ppp_rank.cu) */
    size_t cap_ranked = 0;
    uint32_t *d_rank_tmp = nullptr;
    size_t cap_rank_tmp = 0;
    bool results_ranked = false;
    int dup = 0; /* ranked lists were packed for -dup: the sweep drops the candidate that exceeds |Y| (ppp.pm:222-224) */
    int launch_pairs_log2 = 0; /* filtered runs: pairs per launch (0 = chosen from the free device memory; tests lower it to exercise
                                * runs of several launches) */
    int auto_pairs_log2 = 0;
    int device_rank = 1; /* 0: rank threshold results on the host (developer A/B) */
    uint32_t *d_exact = nullptr;
    size_t cap_exact = 0;
    unsigned long long *d_counters = nullptr;
    /* filtered runs */
    double *d_kth = nullptr;
    size_t cap_kth = 0;
    uint32_t *d_qcnt = nullptr;
    size_t cap_qcnt = 0;
    uint32_t *d_off = nullptr;
    size_t cap_off = 0;
    uint16_t *d_nmax = nullptr;
    size_t cap_nmax = 0;
    ppp_hit *d_topk = nullptr;
    size_t cap_topk = 0;
    uint32_t last_topk_k = 0; /* k of the resident device top-k lists (0 = none) */
    size_t last_topk_nq = 0;
    ppp_hit *d_mtopk = nullptr; /* scratch of ppp_merge_topk_device */
    size_t cap_mtopk = 0;
    double *d_mkth = nullptr;
    size_t cap_mkth = 0;
    uint32_t *d_mcnt = nullptr;
    size_t cap_mcnt = 0;
    std::vector<uint32_t> nmax_off_host; /* static staircase offsets: prefix sums of (ny + 1) */
    size_t nmax_entries = 0;

    /* exact table mode (dense runs, few queries x many targets) */
    std::vector<unsigned long long> toff_host; /* [nQ+1] table offsets: (ny+1) * (npres+1) cells per query */
    double *d_tab = nullptr;
    size_t cap_tab = 0;
    unsigned long long *d_toff = nullptr;
    size_t cap_toff = 0;
    uint32_t *d_tbad = nullptr;
    size_t cap_tbad = 0;
    bool table_valid = false, table_ok = false;
    int table_mode = -1; /* -1 auto, 0 off, 1 always (when it fits) */
    int last_used_table = 0;
    /* options */
    int force_exact = 0, check_bounds = 0, sweep_warps = 16; /* warps per CTA of the tile / refine kernels: 16 measured best (dense G = 4096: 104 ms against 160 ms with 8) */
    bool debug_chunks = false; /* PPP_B200_DEBUG_CHUNKS was set when the context was created */
    ppp_stats stats;
    unsigned long long last_counters[CNT_N];
};

static char g_init_err[512] = "";

static int fail(ppp_ctx *ctx, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_init_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? PPP_ERR_NOMEM : PPP_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e_));                                                       \
    } while (0)

template <typename T>
static int grow(ppp_ctx *ctx, T **ptr, size_t *cap, size_t need)
{
    if (need <= *cap) return PPP_OK;
    cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    CU(cudaMalloc(ptr, need * sizeof(T)));
    *cap = need;
    return PPP_OK;
}

extern "C" uint32_t ppp_words_per_row(uint32_t G) { return ((G + 127u) / 128u) * 4u; }
extern "C" int ppp_abi_version(void) { return PPP_B200_ABI_VERSION; }

extern "C" const char *ppp_last_error(const ppp_ctx *ctx) { return ctx ? ctx->err : g_init_err; }

static int group_init(ppp_ctx **out, const int *devices, int ndev);
static int group_set_targets_bits(ppp_ctx *g, const uint32_t *T, size_t nT, uint32_t G);
static int group_set_targets_ranked(ppp_ctx *g, const uint16_t *idx, const uint16_t *rawdepth, const uint64_t *row_off, size_t nT, uint32_t G);
static int group_set_queries(ppp_ctx *g, const uint32_t *P, const uint32_t *Y, const double *p, size_t nQ);
static int group_run(ppp_ctx *g, const ppp_filter *f);
static int group_fetch(ppp_ctx *g, ppp_hit *out, size_t cap, size_t *nout);
static int group_set_option(ppp_ctx *g, const char *name, int64_t value);
static int group_get_counter(const ppp_ctx *g, const char *name, int64_t *value);

extern "C" int ppp_init(ppp_ctx **out, const int *devices, int ndev)
{
    if (!out) return fail(nullptr, PPP_ERR_INVALID, "ppp_init: out is NULL");
    *out = nullptr;
    if (devices && ndev > 1) return group_init(out, devices, ndev);
    int dev = (devices && ndev > 0) ? devices[0] : 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, PPP_ERR_CUDA, "ppp_init: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (dev < 0 || dev >= count) return fail(nullptr, PPP_ERR_INVALID, "ppp_init: device %d out of range [0,%d)", dev, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess)
        return fail(nullptr, PPP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, PPP_ERR_CUDA, "ppp_init: device %d is sm_%d%d; this library is built for sm_100a only", dev,
                    prop.major, prop.minor);
    ppp_ctx *ctx = new (std::nothrow) ppp_ctx();
    if (!ctx) return fail(nullptr, PPP_ERR_NOMEM, "ppp_init: out of host memory");
    ctx->device = dev;
    ctx->nsm = prop.multiProcessorCount;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    memset(ctx->last_counters, 0, sizeof(ctx->last_counters));
    if ((e = cudaSetDevice(dev)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        fail(nullptr, PPP_ERR_CUDA, "ppp_init: %s", cudaGetErrorString(e));
        delete ctx;
        return PPP_ERR_CUDA;
    }
    for (int i = 0; i < 8; ++i) {
        if ((e = cudaEventCreate(&ctx->ev[i])) != cudaSuccess) {
            fail(nullptr, PPP_ERR_CUDA, "ppp_init: %s", cudaGetErrorString(e));
            ppp_destroy(ctx); /* releases the stream and the events created so far */
            return PPP_ERR_CUDA;
        }
    }
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaStreamCreateWithFlags(&ctx->pf_stream[i], cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&ctx->pf_join[i], cudaEventDisableTiming)) != cudaSuccess) {
            fail(nullptr, PPP_ERR_CUDA, "ppp_init: %s", cudaGetErrorString(e));
            ppp_destroy(ctx);
            return PPP_ERR_CUDA;
        }
    }
    if ((e = cudaEventCreateWithFlags(&ctx->pf_fork, cudaEventDisableTiming)) != cudaSuccess) {
        fail(nullptr, PPP_ERR_CUDA, "ppp_init: %s", cudaGetErrorString(e));
        ppp_destroy(ctx);
        return PPP_ERR_CUDA;
    }
    if ((e = cudaMalloc(&ctx->d_counters, CNT_SLOTS * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&ctx->d_totals, (CNT_N + 1) * sizeof(unsigned long long))) != cudaSuccess) {
        fail(nullptr, e == cudaErrorMemoryAllocation ? PPP_ERR_NOMEM : PPP_ERR_CUDA, "ppp_init: %s", cudaGetErrorString(e));
        ppp_destroy(ctx);
        return e == cudaErrorMemoryAllocation ? PPP_ERR_NOMEM : PPP_ERR_CUDA;
    }
    ctx->debug_chunks = getenv("PPP_B200_DEBUG_CHUNKS") != nullptr; /* developer aid, read once */
    *out = ctx;
    return PPP_OK;
}

extern "C" void ppp_destroy(ppp_ctx *ctx)
{
    if (!ctx) return;
    if (!ctx->subs.empty()) {
        for (ppp_ctx *s : ctx->subs) ppp_destroy(s);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_T);
    cudaFree(ctx->d_Tt);
    cudaFree(ctx->d_surv);
    cudaFree(ctx->d_surv1);
    cudaFree(ctx->d_ridx);
    cudaFree(ctx->d_rraw);
    cudaFree(ctx->d_roff);
    cudaFree(ctx->d_lf);
    cudaFree(ctx->d_Q);
    cudaFree(ctx->d_qc);
    cudaFree(ctx->d_qorder);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaFree(ctx->d_pfstats);
    cudaFree(ctx->d_hits);
    cudaFree(ctx->d_exact);
    cudaFree(ctx->d_ranked);
    cudaFree(ctx->d_acc);
    cudaFree(ctx->d_rank_tmp);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_totals);
    for (cudaEvent_t e : ctx->evpool) cudaEventDestroy(e);
    cudaFree(ctx->d_kth);
    cudaFree(ctx->d_qcnt);
    cudaFree(ctx->d_off);
    cudaFree(ctx->d_nmax);
    cudaFree(ctx->d_topk);
    cudaFree(ctx->d_tab);
    cudaFree(ctx->d_toff);
    cudaFree(ctx->d_tbad);
    cudaFree(ctx->d_mtopk);
    cudaFree(ctx->d_mkth);
    cudaFree(ctx->d_mcnt);
    for (int i = 0; i < 8; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    for (cudaEvent_t e : ctx->piece_evpool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (ctx->pf_stream[i]) {
            cudaStreamSynchronize(ctx->pf_stream[i]);
            cudaStreamDestroy(ctx->pf_stream[i]);
        }
        if (ctx->pf_join[i]) cudaEventDestroy(ctx->pf_join[i]);
    }
    if (ctx->pf_fork) cudaEventDestroy(ctx->pf_fork);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

/* log-factorial table lf[k] = lgamma(k+1), k = 0..8193, computed in long double once per context */
static int ensure_lf(ppp_ctx *ctx)
{
    if (ctx->d_lf) return PPP_OK;
    const size_t n = 8192 + 2;
    std::vector<double> lf(n);
    for (size_t k = 0; k < n; ++k) lf[k] = (double)lgammal((long double)k + 1.0L);
    CU(cudaMalloc(&ctx->d_lf, n * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->d_lf, lf.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PPP_OK;
}

static int pick_wpl(uint32_t G)
{
    if (G <= 1024) return 1;
    if (G <= 2048) return 2;
    if (G <= 4096) return 4;
    if (G <= 8192) return 8;
    return 0;
}

static void free_query_buffers(ppp_ctx *ctx);
static void reset_queries_if_layout_changed(ppp_ctx *ctx, int wpl, uint32_t G);
static int group_set_targets_bits_streamed(ppp_ctx *g, const uint32_t *T, size_t nT, uint32_t G);

/* a streamed upload that is still in flight is completed before the target buffers are touched again */
static int drain_streamed_upload(ppp_ctx *ctx)
{
    if (ctx->copy_stream && !ctx->pieces.empty()) CU(cudaStreamSynchronize(ctx->copy_stream));
    ctx->pieces.clear();
    ctx->pieces_waited = 0;
    ctx->streamed_src = nullptr;
    return PPP_OK;
}

/* Queue the pieces of a streamed upload on the copy stream: 4 MiB, 16 MiB, then 32 MiB each (the first launches of a top-k run
 * read only a few thousand rows).  Done at the first launch that reads targets, not inside ppp_set_targets_bits_streamed: the
 * host-to-device copy engine serves its queue in order, and the small uploads of ppp_set_queries and of the run's own set-up
 * (offsets, initial thresholds) would otherwise sit behind half a gigabyte of planes. */
static int issue_streamed_upload(ppp_ctx *ctx)
{
    const uint32_t *T = ctx->streamed_src;
    if (!T) return PPP_OK;
    ctx->streamed_src = nullptr;
    const size_t nT = ctx->nT, W = 32 * (size_t)ctx->wpl, wpr = ppp_words_per_row(ctx->G), row_bytes = W * 4;
    if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    /* everything queued on the compute stream so far may still read the old planes */
    CU(cudaEventRecord(ctx->ev[7], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[7], 0));
    size_t done = 0, piece_bytes = (size_t)4 << 20;
    while (done < nT) {
        const size_t rows = std::min(nT - done, std::max<size_t>(1024, piece_bytes / row_bytes));
        if (wpr != W) CU(cudaMemsetAsync(ctx->d_T + done * W, 0, rows * row_bytes, ctx->copy_stream));
        CU(cudaMemcpy2DAsync(ctx->d_T + done * W, row_bytes, T + done * wpr, wpr * 4, wpr * 4, rows, cudaMemcpyHostToDevice, ctx->copy_stream));
        const size_t i = ctx->pieces.size();
        if (i == ctx->piece_evpool.size()) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->piece_evpool.push_back(e);
        }
        CU(cudaEventRecord(ctx->piece_evpool[i], ctx->copy_stream));
        done += rows;
        ctx->pieces.push_back({done, ctx->piece_evpool[i]});
        piece_bytes = std::min<size_t>((size_t)32 << 20, piece_bytes * 4);
    }
    return PPP_OK;
}

/* make the targets below row t1 usable by the launches that follow on ctx->stream (no-op unless the upload was streamed) */
static int wait_targets(ppp_ctx *ctx, size_t t1)
{
    int rc;
    if ((rc = issue_streamed_upload(ctx))) return rc;
    while (ctx->pieces_waited < ctx->pieces.size()) {
        const size_t i = ctx->pieces_waited, begin = i ? ctx->pieces[i - 1].t_end : 0;
        if (begin >= t1) break;
        CU(cudaStreamWaitEvent(ctx->stream, ctx->pieces[i].ev, 0));
        CU(ppp_launch_transpose_rows(ctx->d_T, ctx->d_Tt, (uint32_t)ctx->nT, ctx->wpl, (uint32_t)begin, (uint32_t)ctx->pieces[i].t_end, ctx->stream));
        ctx->pieces_waited++;
    }
    return PPP_OK;
}

extern "C" int ppp_set_targets_bits_streamed(ppp_ctx *ctx, const uint32_t *T, size_t nT, uint32_t G)
{
    if (ctx && !ctx->subs.empty()) return group_set_targets_bits_streamed(ctx, T, nT, G);
    if (!ctx) return PPP_ERR_INVALID;
    if (!T && nT) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits_streamed: T is NULL");
    if (G == 0 || G > 8192) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits_streamed: G = %u outside [1, 8192]", G);
    if (nT > 0xffffffffull) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits_streamed: nT too large");
    CU(cudaSetDevice(ctx->device));
    const int wpl = pick_wpl(G);
    const size_t W = 32 * (size_t)wpl;
    ctx->nT = 0;
    ctx->Tt_valid = false;
    int rc;
    if ((rc = drain_streamed_upload(ctx))) return rc;
    if ((rc = grow(ctx, &ctx->d_T, &ctx->cap_T, std::max<size_t>(1, nT * W)))) return rc;
    if ((rc = grow(ctx, &ctx->d_Tt, &ctx->cap_Tt, std::max<size_t>(1, nT * W)))) return rc;
    if ((rc = ensure_lf(ctx))) return rc;
    ctx->ranked = false;
    reset_queries_if_layout_changed(ctx, wpl, G);
    ctx->nT = nT;
    ctx->G = G;
    ctx->wpl = wpl;
    ctx->streamed_src = nT ? T : nullptr;
    ctx->Tt_valid = true; /* built piece by piece in wait_targets */
    return PPP_OK;
}

extern "C" int ppp_wait_targets(ppp_ctx *ctx)
{
    if (!ctx) return PPP_ERR_INVALID;
    if (!ctx->subs.empty()) {
        for (ppp_ctx *s : ctx->subs) {
            const int rc = ppp_wait_targets(s);
            if (rc) return fail(ctx, rc, "%s", s->err);
        }
        return PPP_OK;
    }
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = wait_targets(ctx, ctx->nT))) return rc;
    if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PPP_OK;
}

extern "C" int ppp_set_targets_bits(ppp_ctx *ctx, const uint32_t *T, size_t nT, uint32_t G)
{
    if (ctx && !ctx->subs.empty()) return group_set_targets_bits(ctx, T, nT, G);
    if (!ctx) return PPP_ERR_INVALID;
    if (!T && nT) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits: T is NULL");
    if (G == 0 || G > 8192) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits: G = %u outside [1, 8192]", G);
    if (nT > 0xffffffffull) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits: nT too large");
    CU(cudaSetDevice(ctx->device));
    const int wpl = pick_wpl(G);
    const size_t W = 32 * (size_t)wpl, wpr = ppp_words_per_row(G);
    ctx->nT = 0;
    ctx->Tt_valid = false;
    int rc;
    if ((rc = drain_streamed_upload(ctx))) return rc;
    if ((rc = grow(ctx, &ctx->d_T, &ctx->cap_T, std::max<size_t>(1, nT * W)))) return rc; /* buffers are reused across calls */
    if (nT) {
        if (wpr != W) CU(cudaMemsetAsync(ctx->d_T, 0, nT * W * 4, ctx->stream));
        /* cudaMemcpyDefault: T may also be device memory (ppp_set_targets_bits_device) */
        CU(cudaMemcpy2DAsync(ctx->d_T, W * 4, T, wpr * 4, wpr * 4, nT, cudaMemcpyDefault, ctx->stream));
    }
    if ((rc = ensure_lf(ctx))) return rc;
    ctx->ranked = false;
    CU(cudaStreamSynchronize(ctx->stream));
    reset_queries_if_layout_changed(ctx, wpl, G);
    ctx->nT = nT;
    ctx->G = G;
    ctx->wpl = wpl;
    return PPP_OK;
}

extern "C" int ppp_set_targets_bits_device(ppp_ctx *ctx, const uint32_t *dT, size_t nT, uint32_t G)
{
    if (!ctx) return PPP_ERR_INVALID;
    if (nT) {
        cudaPointerAttributes at;
        if (!dT || cudaPointerGetAttributes(&at, dT) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
            (void)cudaGetLastError();
            return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits_device: dT is not device memory");
        }
    }
    return ppp_set_targets_bits(ctx, dT, nT, G);
}

static void free_query_buffers(ppp_ctx *ctx)
{
    cudaFree(ctx->d_Q);
    cudaFree(ctx->d_qc);
    cudaFree(ctx->d_qorder);
    ctx->d_Q = nullptr;
    ctx->d_qc = nullptr;
    ctx->d_qorder = nullptr;
    ctx->capQ = 0;
    ctx->nQ = 0;
}

static void reset_queries_if_layout_changed(ppp_ctx *ctx, int wpl, uint32_t G)
{
    if (ctx->wpl != wpl || ctx->G != G) free_query_buffers(ctx); /* queries must be re-uploaded: the row layout changed */
}

extern "C" int ppp_set_targets_ranked(ppp_ctx *ctx, const uint16_t *idx, const uint16_t *rawdepth, const uint64_t *row_off,
                                      size_t nT, uint32_t G)
{
    if (ctx && !ctx->subs.empty()) return group_set_targets_ranked(ctx, idx, rawdepth, row_off, nT, G);
    if (!ctx) return PPP_ERR_INVALID;
    if (!row_off || (nT && row_off[nT] && (!idx || !rawdepth))) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_ranked: NULL input");
    if (G == 0 || G > 8192) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_ranked: G = %u outside [1, 8192]", G);
    if (nT > 0xffffffffull) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_ranked: nT too large");
    size_t lmax = 0;
    for (size_t t = 0; t < nT; ++t) {
        if (row_off[t + 1] < row_off[t]) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_ranked: row_off not monotone at %zu", t);
        lmax = std::max<size_t>(lmax, row_off[t + 1] - row_off[t]);
    }
    if (lmax > 2048)
        return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_ranked: a list has %zu distinct taxa; at most 2048 are supported", lmax);
    const size_t total = nT ? row_off[nT] : 0;
    for (size_t i = 0; i < total; ++i)
        if (idx[i] != 0xFFFF && idx[i] >= G) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_ranked: idx[%zu] = %u >= G", i, idx[i]);
    CU(cudaSetDevice(ctx->device));
    ctx->nT = 0;
    ctx->Tt_valid = false;
    int rc;
    if ((rc = drain_streamed_upload(ctx))) return rc;
    if ((rc = ensure_lf(ctx))) return rc;
    if ((rc = grow(ctx, &ctx->d_ridx, &ctx->cap_ridx, total + 64))) return rc;
    if ((rc = grow(ctx, &ctx->d_rraw, &ctx->cap_rraw, total + 64))) return rc;
    if ((rc = grow(ctx, &ctx->d_roff, &ctx->cap_roff, nT + 1))) return rc;
    if (total) {
        CU(cudaMemcpyAsync(ctx->d_ridx, idx, total * 2, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_rraw, rawdepth, total * 2, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaMemcpyAsync(ctx->d_roff, row_off, (nT + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const int wpl = pick_wpl(G);
    reset_queries_if_layout_changed(ctx, wpl, G);
    ctx->nT = nT;
    ctx->G = G;
    ctx->wpl = wpl;
    ctx->ranked = true;
    ctx->rwpl = lmax <= 1024 ? 1 : 2;
    return PPP_OK;
}

/* positions [b, e) of pre-filter class `kind` in d_qorder: 1 = all-ones presence plane, 2 = P == Y, 0 = the others */
static void class_range(const ppp_ctx *ctx, int kind, size_t nQ, size_t *b, size_t *e)
{
    if (kind == 1) {
        *b = 0;
        *e = ctx->n_pall;
    } else if (kind == 2) {
        *b = ctx->n_pall;
        *e = ctx->n_pall + ctx->n_pyy;
    } else {
        *b = ctx->n_pall + ctx->n_pyy;
        *e = nQ;
    }
}

static void make_qconst(double p, uint32_t ny, uint32_t npres, QConst *qc)
{
    memset(qc, 0, sizeof(*qc));
    qc->p = p;
    qc->ny = ny;
    qc->npres = npres;
    if (isnan(p) || p == 0.0 || p == 1.0) {
        qc->flags = PPP_QF_TRIVIAL;
        return;
    }
    if (p < 0.0 || p > 1.0) {
        qc->flags = PPP_QF_DOMAIN;
        return;
    }
    const long double P = (long double)p, Q = 1.0L - P;
    const long double lnp = logl(P), lnq = log1pl(-P), theta = P / Q;
    qc->lnp = (double)lnp;
    qc->lnq = (double)lnq;
    qc->lntheta = (double)(lnp - lnq);
    qc->theta = (double)theta;
    qc->thetaf = (float)theta;
    qc->inv_thetaf = (float)(1.0L / theta);
    qc->pf = (float)p;
    qc->varf_unit = (float)(P * Q);
    if (p < 1e-30 || (double)Q < 1e-12) qc->flags = PPP_QF_EXACT;
}

extern "C" int ppp_set_queries(ppp_ctx *ctx, const uint32_t *P, const uint32_t *Y, const double *p, size_t nQ)
{
    if (ctx && !ctx->subs.empty()) return group_set_queries(ctx, P, Y, p, nQ);
    if (!ctx) return PPP_ERR_INVALID;
    if (!ctx->wpl) return fail(ctx, PPP_ERR_STATE, "ppp_set_queries: call ppp_set_targets_* first (it fixes G)");
    if (nQ && (!P || !Y || !p)) return fail(ctx, PPP_ERR_INVALID, "ppp_set_queries: NULL input");
    if (nQ > 0xffffffffull) return fail(ctx, PPP_ERR_INVALID, "ppp_set_queries: nQ too large");
    CU(cudaSetDevice(ctx->device));
    const size_t W = 32 * (size_t)ctx->wpl, wpr = ppp_words_per_row(ctx->G);
    if (nQ > ctx->capQ) {
        free_query_buffers(ctx);
        /* round up to a whole tile so that a tile's TMA bulk copy never reads past the allocation */
        const size_t cap = ((nQ + PPP_QTILE - 1) / PPP_QTILE) * PPP_QTILE;
        cudaError_t e;
        if ((e = cudaMalloc(&ctx->d_Q, cap * 2 * W * 4)) != cudaSuccess || (e = cudaMalloc(&ctx->d_qc, cap * sizeof(QConst))) != cudaSuccess ||
            (e = cudaMalloc(&ctx->d_qorder, cap * sizeof(uint32_t))) != cudaSuccess) {
            free_query_buffers(ctx); /* no live pointer is left behind a zero capacity */
            return fail(ctx, e == cudaErrorMemoryAllocation ? PPP_ERR_NOMEM : PPP_ERR_CUDA, "ppp_set_queries: %s", cudaGetErrorString(e));
        }
        ctx->capQ = cap; /* only once all three buffers exist */
    }
    ctx->nQ = 0;
    if (!nQ) return PPP_OK;
    std::vector<QConst> qc(nQ);
    /* device layout [nQ][2][W]: P and Y rows interleaved (zero padded to W words) in a pinned staging buffer and uploaded with
     * ONE contiguous copy -- two pitched 2-D copies of 512-byte rows cost ~6 ms for 10,000 queries */
    const size_t stage_need = ctx->capQ * 2 * W * 4;
    if (stage_need > ctx->cap_stage) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->cap_stage = 0;
        CU(cudaMallocHost(&ctx->h_stage, stage_need));
        ctx->cap_stage = stage_need;
    }
    uint32_t *st = reinterpret_cast<uint32_t *>(ctx->h_stage);
    if (nQ < ctx->capQ) memset(st + nQ * 2 * W, 0, (ctx->capQ - nQ) * 2 * W * 4); /* rows of the last partial tile */
    {
        /* per query: |Y| and |P| (POPCNT, ppp_bla.cpp), the constants (long double logarithms) and the staging copy -- a few
         * milliseconds for 10,000 queries on one core, so the queries are split over a few host threads */
        std::vector<uint32_t> cy(nQ), cp(nQ);
        auto work = [&](size_t q0, size_t q1) {
            ppp_host_count_rows(P + q0 * wpr, Y + q0 * wpr, q1 - q0, wpr, cy.data() + q0, cp.data() + q0);
            for (size_t q = q0; q < q1; ++q) {
                make_qconst(p[q], cy[q], cp[q], &qc[q]);
                uint32_t *row = st + q * 2 * W;
                if (wpr < W) memset(row, 0, 2 * W * 4); /* zero padding inside the rows */
                memcpy(row, P + q * wpr, wpr * 4);
                memcpy(row + W, Y + q * wpr, wpr * 4);
            }
        };
        const size_t nth = std::min<size_t>(std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency())), nQ / 1024 + 1);
        std::vector<std::thread> pool;
        for (size_t i = 1; i < nth; ++i) pool.emplace_back(work, nQ * i / nth, nQ * (i + 1) / nth);
        work(0, nQ / nth);
        for (std::thread &t : pool) t.join();
    }
    CU(cudaMemcpyAsync(ctx->d_Q, st, stage_need, cudaMemcpyHostToDevice, ctx->stream)); /* overlaps the host work below */
    ctx->nmax_off_host.resize(nQ);
    {
        size_t off = 0;
        for (size_t q = 0; q < nQ; ++q) {
            ctx->nmax_off_host[q] = (uint32_t)off;
            off += (size_t)qc[q].ny + 1;
        }
        if (off > 0xfffffff0ull) {
            cudaStreamSynchronize(ctx->stream); /* the plane upload queued above reads the staging buffer */
            return fail(ctx, PPP_ERR_INVALID, "ppp_set_queries: staircase tables exceed 4G entries");
        }
        ctx->nmax_entries = off;
    }
    ctx->toff_host.assign(nQ + 1, 0);
    for (size_t q = 0; q < nQ; ++q)
        ctx->toff_host[q + 1] = ctx->toff_host[q] + ((unsigned long long)qc[q].ny + 1) * ((unsigned long long)qc[q].npres + 1);
    ctx->table_valid = false;
    CU(cudaMemcpyAsync(ctx->d_qc, qc.data(), nQ * sizeof(QConst), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint32_t> qfl(nQ);
    for (size_t q = 0; q < nQ; ++q) qfl[q] = (qc[q].npres == ctx->G) ? 1u : 0u;
    {
        /* processing order of the pre-filter, one launch per class: all-ones presence planes (number == depth), P == Y (number ==
         * yes; Y is a subset of P, so equal counts mean equal planes), everything else */
        std::vector<uint32_t> order;
        order.reserve(nQ);
        auto is_pyy = [&](size_t q) { return !qfl[q] && qc[q].ny == qc[q].npres; };
        for (size_t q = 0; q < nQ; ++q)
            if (qfl[q]) order.push_back((uint32_t)q);
        ctx->n_pall = order.size();
        for (size_t q = 0; q < nQ; ++q)
            if (is_pyy(q)) order.push_back((uint32_t)q);
        ctx->n_pyy = order.size() - ctx->n_pall;
        for (size_t q = 0; q < nQ; ++q)
            if (!qfl[q] && !is_pyy(q)) order.push_back((uint32_t)q);
        /* deal the queries of each class over its tiles by descending |Y| (tile i takes the i-th, (i + ntiles)-th ...
         * largest): every tile then holds about the same staircase footprint, which is what sizes the kernel's shared
         * memory (and hence its occupancy), and about the same amount of slow-path work.  A tile normally has the kernel's
         * 32 (general launch: 16) queries; when the staircases of such a tile would not fit the kernel's shared memory (very
         * dense yes planes: 32 queries of |Y| > 1,000 each) the class runs with smaller tiles -- fewer queries share a target
         * row in registers, nothing else changes. */
        for (int kind = 0; kind < 3; ++kind) {
            size_t b, e;
            class_range(ctx, kind, nQ, &b, &e);
            const size_t n = e - b;
            const size_t cap = ppp_prefilter2_max_stair(ctx->wpl, kind);
            std::vector<uint32_t> sorted(order.begin() + b, order.begin() + e);
            std::stable_sort(sorted.begin(), sorted.end(), [&](uint32_t x, uint32_t y) { return qc[x].ny > qc[y].ny; });
            size_t qt = ppp_prefilter2_tile_queries(kind), worst = 1;
            for (;; --qt) {
                const size_t ntiles = n ? (n + qt - 1) / qt : 0;
                size_t o = b;
                for (size_t tile = 0; tile < ntiles; ++tile)
                    for (size_t j = 0; j * ntiles + tile < n && j < qt; ++j) order[o++] = sorted[j * ntiles + tile];
                /* what the kernel will stage: ITS tiles are the consecutive groups of qt entries of the order */
                worst = 1;
                for (size_t i = b; i < e; i += qt) {
                    size_t sum = 1;
                    for (size_t j = i; j < std::min(e, i + qt); ++j) sum += (size_t)qc[order[j]].ny + 1;
                    worst = std::max(worst, sum);
                }
                if (worst <= cap || qt == 1) break;
            }
            ctx->tile_q[kind] = (uint32_t)qt;
            ctx->stair_need[kind] = (uint32_t)std::min<size_t>(worst, 0xffffffffu);
        }
        CU(cudaMemcpyAsync(ctx->d_qorder, order.data(), nQ * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        for (int kind = 0; kind < 2; ++kind) {
            size_t b, e;
            class_range(ctx, kind, nQ, &b, &e);
            const size_t qt = ppp_prefilter_tc_tile_queries(kind);
            size_t worst = 0;
            for (size_t i = b; i < e; i += qt) {
                size_t sum = 1;
                for (size_t j = i; j < std::min(e, i + qt); ++j) sum += (size_t)qc[order[j]].ny + 1;
                worst = std::max(worst, sum);
            }
            ctx->stair_need_tc[kind] = (uint32_t)std::min<size_t>(worst, 0xffffffffu);
        }
        { /* a tile of ARBITRARY queries (the re-scan of the queries whose speculative threshold failed): the largest tile whose
           * worst case -- the queries with the longest staircases -- fits */
            std::vector<uint32_t> lens(nQ);
            for (size_t q = 0; q < nQ; ++q) lens[q] = qc[q].ny + 1;
            const size_t qt = std::min<size_t>(ppp_prefilter2_tile_queries(0), nQ), cap = ppp_prefilter2_max_stair(ctx->wpl, 0);
            std::partial_sort(lens.begin(), lens.begin() + qt, lens.end(), std::greater<uint32_t>());
            size_t sum = 1 + lens[0], t = 1;
            while (t < qt && sum + lens[t] <= cap) sum += lens[t++];
            ctx->tile_q_any = (uint32_t)t;
            ctx->stair_need_any = (uint32_t)std::min<size_t>(sum, 0xffffffffu);
        }
    }
    CU(cudaStreamSynchronize(ctx->stream)); /* inputs are copied before returning (Perl SV buffers may move) */
    ctx->nQ = nQ;
    return PPP_OK;
}

#define PPP_MAX_SPEC_LEVELS 6

static bool hit_less(const ppp_hit &a, const ppp_hit &b)
{
    if (a.q != b.q) return a.q < b.q;
    if (a.score != b.score) return a.score < b.score;
    return a.t < b.t;
}

static void base_params(const ppp_ctx *ctx, SweepParams *prm)
{
    memset(prm, 0, sizeof(*prm));
    prm->T = ctx->d_T;
    prm->Q = ctx->d_Q;
    prm->qc = ctx->d_qc;
    prm->lf = ctx->d_lf;
    prm->exact_list = ctx->d_exact;
    prm->counters = ctx->d_counters;
    prm->exact_cap = ctx->cap_exact;
    prm->nQ = (uint32_t)ctx->nQ;
    prm->nT = (uint32_t)ctx->nT;
    prm->G = ctx->G;
    prm->force_exact = (uint32_t)ctx->force_exact;
    prm->check_bounds = (uint32_t)ctx->check_bounds;
    prm->qwords = 32u * (uint32_t)ctx->wpl;
    prm->dup = (uint32_t)(ctx->dup && ctx->ranked);
    prm->ridx = ctx->d_ridx;
    prm->rraw = ctx->d_rraw;
    prm->roff = ctx->d_roff;
}

static int mark(ppp_ctx *ctx, size_t *idx)
{
    if (ctx->ev_used == ctx->evpool.size()) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        ctx->evpool.push_back(e);
    }
    *idx = ctx->ev_used++;
    CU(cudaEventRecord(ctx->evpool[*idx], ctx->stream));
    return PPP_OK;
}

/* pre-filter + refine of one target chunk (bit-plane targets): ppp_prefilter.cu screens every pair, the survivors go to the
 * refine kernel.  A tile's staircases live in shared memory; ppp_set_queries chose tile sizes that fit. */
static int launch_prefiltered(ppp_ctx *ctx, SweepParams *prm, cudaEvent_t between, const uint32_t *qlist = nullptr, uint32_t nlist = 0)
{
    unsigned long long *surv_count = ctx->d_counters + CNT_N;
    if (qlist) { /* only the listed queries (any presence plane): one general launch */
        CU(ppp_launch_prefilter2(prm, ctx->wpl, 0, ctx->d_Tt, (uint32_t)ctx->nT, qlist, nlist, ctx->tile_q_any, ctx->stair_need_any, ctx->d_surv,
                                 surv_count, ctx->cap_surv, ctx->d_pfstats, ctx->nsm, ctx->stream));
        if (between) CU(cudaEventRecord(between, ctx->stream));
        CU(ppp_launch_refine(prm, ctx->wpl, ctx->d_surv, surv_count, ctx->cap_surv, ctx->sweep_warps, ctx->nsm, ctx->stream));
        return PPP_OK;
    }
    /* all-ones presence planes on the run's stream, the general ones and P == Y on the side streams (the launches append to
     * the same survivor list with atomics and claim their items from separate counters) */
    const bool overlap = ctx->pf_overlap && !ctx->prefilter_tc;
    int nside = 0;
    if (overlap) CU(cudaEventRecord(ctx->pf_fork, ctx->stream));
    for (int kind : {1, 0, 2}) {
        size_t cb, ce;
        class_range(ctx, kind, ctx->nQ, &cb, &ce);
        const uint32_t *ql = ctx->d_qorder + cb;
        const uint32_t n = (uint32_t)(ce - cb);
        if (!n) continue;
        if (kind != 2 && ((ctx->prefilter_tc >> (kind ? 0 : 1)) & 1) && ctx->stair_need_tc[kind] <= ppp_prefilter_tc_max_stair(ctx->wpl, kind) &&
            (ctx->tc_exact || ctx->stair_need_tc[kind] <= ppp_rescreen_max_stair(ctx->wpl, kind))) {
            /* the chunk counts come from the tensor cores (ppp_prefilter_tc.cu); in pair-level mode the kernel writes one column
             * mask per (query tile, target) -- the pairs with a passing CHUNK -- which the re-screen kernel reduces to the pairs
             * with a passing candidate */
            const bool pl = !ctx->tc_exact;
            if (pl) {
                const size_t nqt = ppp_prefilter_tc_tile_queries(kind), tiles = (n + nqt - 1) / nqt;
                const int rc = grow(ctx, &ctx->d_surv1, &ctx->cap_surv1, tiles * (size_t)(prm->t1 - prm->t0) * (nqt / 32));
                if (rc) return rc;
            }
            CU(ppp_launch_prefilter_tc(prm, ctx->wpl, kind, ctx->d_Tt, (uint32_t)ctx->nT, ql, n, ctx->stair_need_tc[kind], ctx->d_surv, surv_count,
                                       ctx->cap_surv, ctx->tc_exact, ctx->d_surv1, ctx->nsm, ctx->stream));
            ctx->tc_launches++;
            if (pl)
                CU(ppp_launch_rescreen_masks(prm, ctx->wpl, kind, ql, n, ctx->stair_need_tc[kind], ctx->d_surv1, ctx->d_surv, surv_count, ctx->cap_surv,
                                             ctx->nsm, ctx->stream));
        } else {
            cudaStream_t st = ctx->stream;
            if (overlap && kind != 1) {
                st = ctx->pf_stream[nside];
                CU(cudaStreamWaitEvent(st, ctx->pf_fork, 0));
            }
            CU(ppp_launch_prefilter2(prm, ctx->wpl, kind, ctx->d_Tt, (uint32_t)ctx->nT, ql, n, ctx->tile_q[kind], ctx->stair_need[kind], ctx->d_surv,
                                     surv_count, ctx->cap_surv, ctx->d_pfstats, ctx->nsm, st));
            if (st != ctx->stream) {
                CU(cudaEventRecord(ctx->pf_join[nside], st));
                ++nside;
            }
        }
    }
    for (int i = 0; i < nside; ++i) CU(cudaStreamWaitEvent(ctx->stream, ctx->pf_join[i], 0));
    if (between) CU(cudaEventRecord(between, ctx->stream));
    CU(ppp_launch_refine(prm, ctx->wpl, ctx->d_surv, surv_count, ctx->cap_surv, ctx->sweep_warps, ctx->nsm, ctx->stream));
    return PPP_OK;
}

/* one target chunk without any host synchronisation (top-k runs): counters are accumulated on the device, event
 * spans are resolved after the run's final synchronize */
static int launch_chunk_deferred(ppp_ctx *ctx, SweepParams *prm, uint32_t *launches, bool prefiltered, const uint32_t *qlist = nullptr,
                                 uint32_t nlist = 0)
{
    const int W = 32 * ctx->wpl;
    int rc;
    size_t e0, e1, e2, em = 0;
    if ((rc = wait_targets(ctx, prm->t1))) return rc;
    CU(cudaMemsetAsync(ctx->d_counters, 0, CNT_SLOTS * sizeof(unsigned long long), ctx->stream));
    if ((rc = mark(ctx, &e0))) return rc;
    if (prefiltered) {
        if (ctx->ev_used == ctx->evpool.size()) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            ctx->evpool.push_back(e);
        }
        em = ctx->ev_used++;
        if ((rc = launch_prefiltered(ctx, prm, ctx->evpool[em], qlist, nlist))) return rc;
        *launches += 2;
        ctx->spans.push_back({e0, em, 4});
        ctx->prefilter_pairs += (unsigned long long)(qlist ? nlist : prm->nQ) * (prm->t1 - prm->t0);
    } else if (ctx->ranked) {
        CU(ppp_launch_sweep_ranked(prm, ctx->rwpl, ctx->sweep_warps, ctx->nsm, ctx->stream));
    } else {
        CU(ppp_launch_sweep(prm, ctx->wpl, ctx->sweep_warps, ctx->nsm, ctx->stream));
    }
    if ((rc = mark(ctx, &e1))) return rc;
    if (ctx->ranked) CU(ppp_launch_exact_ranked(prm, ctx->rwpl, ctx->nsm, ctx->stream));
    else CU(ppp_launch_exact(prm, W, ctx->nsm, ctx->stream));
    if ((rc = mark(ctx, &e2))) return rc;
    CU(ppp_launch_accumulate(ctx->d_totals, ctx->d_counters, CNT_N + 1, ctx->stream));
    if (ctx->debug_chunks) { /* developer aid: per-chunk counters (synchronises) */
        unsigned long long c[CNT_SLOTS];
        CU(cudaMemcpyAsync(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float m1 = 0.f, m2 = 0.f, mp = 0.f;
        cudaEventElapsedTime(&m1, ctx->evpool[e0], ctx->evpool[e1]);
        cudaEventElapsedTime(&m2, ctx->evpool[e1], ctx->evpool[e2]);
        if (prefiltered) cudaEventElapsedTime(&mp, ctx->evpool[e0], ctx->evpool[em]);
        fprintf(stderr, "[ppp chunk] t=[%u,%u) mode=%u bounds_only=%u prefiltered=%d: survivors %llu full %llu exact %llu cand %llu  sweep %.3f ms (pre-filter %.3f) exact %.3f ms\n",
                prm->t0, prm->t1, prm->accept_mode, prm->bounds_only, (int)prefiltered, c[CNT_N], c[CNT_FULL], c[CNT_EXACT], c[CNT_CAND], m1, mp, m2);
    }
    *launches += 3;
    ctx->spans.push_back({e0, e1, 0});
    ctx->spans.push_back({e1, e2, 1});
    return PPP_OK;
}

/* one target chunk: sweep + exact tier; accumulates counters and the sweep kernel's event time */
static int run_chunk(ppp_ctx *ctx, SweepParams *prm, unsigned long long *totals, float *sweep_ms, uint32_t *launches,
                     unsigned long long *chunk_counters, bool prefiltered)
{
    const int W = 32 * ctx->wpl;
    {
        const int rcw = wait_targets(ctx, prm->t1);
        if (rcw) return rcw;
    }
    CU(cudaMemsetAsync(ctx->d_counters, 0, CNT_SLOTS * sizeof(unsigned long long), ctx->stream));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (prefiltered) {
        {
            const int rc2 = launch_prefiltered(ctx, prm, ctx->ev[7]);
            if (rc2) return rc2;
        }
        *launches += 2;
    } else if (ctx->ranked) {
        CU(ppp_launch_sweep_ranked(prm, ctx->rwpl, ctx->sweep_warps, ctx->nsm, ctx->stream));
    } else {
        CU(ppp_launch_sweep(prm, ctx->wpl, ctx->sweep_warps, ctx->nsm, ctx->stream));
    }
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    if (ctx->ranked) CU(ppp_launch_exact_ranked(prm, ctx->rwpl, ctx->nsm, ctx->stream));
    else CU(ppp_launch_exact(prm, W, ctx->nsm, ctx->stream));
    CU(cudaEventRecord(ctx->ev[4], ctx->stream));
    *launches += 2;
    unsigned long long c[CNT_N + 1];
    CU(cudaMemcpyAsync(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->survivors += c[CNT_N];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    *sweep_ms += ms;
    ctx->phase_us[0] += 1e3 * ms;
    if (prefiltered) {
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[7]);
        ctx->prefilter_us += 1e3 * ms;
        ctx->prefilter_pairs += (unsigned long long)prm->nQ * (prm->t1 - prm->t0);
    }
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
    ctx->phase_us[1] += 1e3 * ms;
    for (int i = 0; i < CNT_N; ++i) totals[i] += c[i];
    if (chunk_counters) memcpy(chunk_counters, c, CNT_N * sizeof(unsigned long long));
    if (c[CNT_OVERFLOW]) return fail(ctx, PPP_ERR_CUDA, "ppp_run: internal buffer overflow (%llu records dropped)", c[CNT_OVERFLOW]);
    return PPP_OK;
}

extern "C" int ppp_run(ppp_ctx *ctx, const ppp_filter *f)
{
    if (ctx && !ctx->subs.empty()) return group_run(ctx, f);
    if (!ctx) return PPP_ERR_INVALID;
    if (!ctx->wpl || !ctx->nT) return fail(ctx, PPP_ERR_STATE, "ppp_run: no targets resident");
    if (!ctx->ranked && !ctx->d_T) return fail(ctx, PPP_ERR_STATE, "ppp_run: no targets resident");
    if (!ctx->nQ) return fail(ctx, PPP_ERR_STATE, "ppp_run: no queries resident");
    ppp_filter all = {PPP_FILTER_ALL, 0, 0.0};
    if (!f) f = &all;
    if (f->mode > PPP_FILTER_TOPK) return fail(ctx, PPP_ERR_INVALID, "ppp_run: unknown filter mode %u", f->mode);
    if (f->mode == PPP_FILTER_TOPK && (f->k == 0 || (int)f->k > ppp_topk_max_k()))
        return fail(ctx, PPP_ERR_INVALID, "ppp_run: top-k needs 1 <= k <= %d", ppp_topk_max_k());
    if (f->mode == PPP_FILTER_THRESH_LOG10 && isnan(f->thresh)) return fail(ctx, PPP_ERR_INVALID, "ppp_run: thresh is NaN");
    CU(cudaSetDevice(ctx->device));
    const size_t nQ = ctx->nQ, nT = ctx->nT;
    int rc;

    ctx->host_results.clear();
    ctx->results_on_host = (f->mode != PPP_FILTER_ALL);
    ctx->results_topk = false;
    ctx->results_ranked = false;
    ctx->last_topk_k = 0;
    unsigned long long totals[CNT_N];
    memset(totals, 0, sizeof(totals));
    float sweep_ms = 0.f;
    uint32_t sweep_launches = 0, launches = 0;

    for (int i = 0; i < 4; ++i) ctx->phase_us[i] = 0;
    ctx->survivors = 0;
    ctx->prefilter_pairs = 0;
    ctx->prefilter_us = 0;
    ctx->tc_launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    ctx->last_used_table = 0;
    bool use_table = false;
    if (f->mode == PPP_FILTER_ALL && !ctx->ranked && !ctx->force_exact && !ctx->check_bounds && ctx->table_mode != 0) {
        /* exact table mode pays when one table serves many targets: |Y| x |P| bit-exact evaluations per query against
         * popc(T&Y) enclosure + series evaluations per pair */
        const unsigned long long cells = ctx->toff_host[nQ];
        const bool fits = cells <= ((unsigned long long)1 << 27);
        const bool pays = nT >= 8192 && cells / 24 <= (unsigned long long)nQ * nT;
        use_table = fits && (ctx->table_mode == 1 || pays);
    }
    if (use_table) {
        if (!ctx->Tt_valid) {
            const size_t W = 32 * (size_t)ctx->wpl;
            if ((rc = grow(ctx, &ctx->d_Tt, &ctx->cap_Tt, nT * W))) return rc;
            CU(ppp_launch_transpose(ctx->d_T, ctx->d_Tt, (uint32_t)nT, ctx->wpl, ctx->stream));
            ctx->Tt_valid = true;
            launches++;
        }
        if (!ctx->table_valid) {
            if ((rc = grow(ctx, &ctx->d_tab, &ctx->cap_tab, (size_t)ctx->toff_host[nQ] + 1))) return rc;
            if ((rc = grow(ctx, &ctx->d_toff, &ctx->cap_toff, nQ + 1))) return rc;
            if ((rc = grow(ctx, &ctx->d_tbad, &ctx->cap_tbad, nQ))) return rc;
            CU(cudaMemcpyAsync(ctx->d_toff, ctx->toff_host.data(), (nQ + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemsetAsync(ctx->d_tbad, 0, nQ * sizeof(uint32_t), ctx->stream));
            CU(cudaEventRecord(ctx->ev[5], ctx->stream));
            CU(ppp_launch_table_build(ctx->d_qc, ctx->d_toff, (uint32_t)nQ, ctx->d_tab, ctx->d_tbad, ctx->nsm, ctx->stream));
            CU(cudaEventRecord(ctx->ev[6], ctx->stream));
            launches += 2;
            std::vector<uint32_t> bad(nQ);
            CU(cudaMemcpyAsync(bad.data(), ctx->d_tbad, nQ * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]);
            ctx->phase_us[2] += 1e3 * ms; /* reported under "staircase": per-query precomputation */
            ctx->table_ok = true;
            for (size_t q = 0; q < nQ; ++q)
                if (bad[q]) ctx->table_ok = false; /* Cephes rounding broke monotonicity somewhere: use the general sweep */
            ctx->table_valid = true;
        }
        use_table = ctx->table_ok;
    }
    if (use_table) {
        if ((rc = grow(ctx, &ctx->d_hits, &ctx->cap_hits, nQ * nT))) return rc;
        SweepParams prm;
        base_params(ctx, &prm);
        prm.hits = ctx->d_hits;
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        /* one launch over all targets -- or, while a streamed upload is arriving, one launch per piece: the sweep of a piece
         * overlaps the copy of the next one (config 2 end to end is then the PCIe time of the planes, not the sum) */
        if ((rc = issue_streamed_upload(ctx))) return rc;
        for (size_t tb = 0; tb < nT;) {
            const size_t te = ctx->pieces_waited < ctx->pieces.size() ? ctx->pieces[ctx->pieces_waited].t_end : nT;
            if ((rc = wait_targets(ctx, te))) return rc;
            prm.t0 = (uint32_t)tb;
            prm.t1 = (uint32_t)te;
            CU(ppp_launch_table_sweep(&prm, ctx->wpl, ctx->d_Tt, (uint32_t)nT, ctx->d_toff, ctx->d_tab, ctx->nsm, ctx->stream));
            launches++;
            sweep_launches++;
            tb = te;
        }
        CU(cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
        sweep_ms += ms;
        ctx->phase_us[0] += 1e3 * ms;
        ctx->n_results = nQ * nT;
        ctx->last_used_table = 1;
    } else if (f->mode == PPP_FILTER_ALL) {
        /* dense: chunks only bound the launch-local pair ids of the exact-tier list (31 bits) */
        size_t tchunk = std::max<size_t>(1, ((size_t)1 << 30) / nQ);
        if (tchunk > nT) tchunk = nT;
        if ((rc = grow(ctx, &ctx->d_hits, &ctx->cap_hits, nQ * nT))) return rc;
        if ((rc = grow(ctx, &ctx->d_exact, &ctx->cap_exact, nQ * tchunk))) return rc;
        for (size_t t0 = 0; t0 < nT; t0 += tchunk) {
            SweepParams prm;
            base_params(ctx, &prm);
            prm.hits = ctx->d_hits;
            prm.t0 = (uint32_t)t0;
            prm.t1 = (uint32_t)std::min(nT, t0 + tchunk);
            if ((rc = run_chunk(ctx, &prm, totals, &sweep_ms, &launches, nullptr, false))) return rc;
            sweep_launches++;
        }
        ctx->n_results = nQ * nT;
    } else {
        /* filtered: per-query staircases prune almost every pair inside the sweep kernel; survivors are appended */
        const bool topk = (f->mode == PPP_FILTER_TOPK);
        const uint32_t k = f->k;
        /* Pairs per launch.  Every launch pays for a staircase rebuild, a list merge and the tails of its kernels (about 0.5 ms at
         * config 3), so launches should be as large as the scratch buffers allow: 40 bytes per pair of a launch (append slots,
         * survivor and exact-tier ids; the pair ids of the exact tier have 31 bits).  Measured on the full config 3: 2^28 pairs
         * 395.1 ms, 2^29 377.8, 2^30 370.9, 2^31 365.5.  Unless the option fixes it, take the largest of 2^31 / 2^30 / 2^29 whose
         * scratch stays below half of the memory that is free when the context first needs it (86 GB for 2^31 on a B200). */
        if (!ctx->launch_pairs_log2 && !ctx->auto_pairs_log2) {
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            ctx->auto_pairs_log2 = 29;
            for (int l = 31; l > 29; --l)
                if (40.0 * ldexp(1.0, l) <= 0.5 * (double)free_b) {
                    ctx->auto_pairs_log2 = l;
                    break;
                }
        }
        const int pairs_log2 = ctx->launch_pairs_log2 ? ctx->launch_pairs_log2 : ctx->auto_pairs_log2;
        const size_t maxw = std::min<size_t>(nT, std::max<size_t>(256, ((size_t)1 << pairs_log2) / nQ));
        if ((rc = grow(ctx, &ctx->d_hits, &ctx->cap_hits, nQ * maxw))) return rc;
        if ((rc = grow(ctx, &ctx->d_exact, &ctx->cap_exact, nQ * maxw))) return rc;
        /* [nQ] k-th scores, [nQ] tau of the resident staircase, [nQ] speculative thresholds in use, [nQ] their initial values,
         * [nQ] speculative thresholds of the first chunk (from the seeding pass), [nQ] the same tightened by the k-th scores,
         * [PPP_MAX_SPEC_LEVELS][nQ] initial thresholds of the later speculation levels */
        if ((rc = grow(ctx, &ctx->d_kth, &ctx->cap_kth, (6 + PPP_MAX_SPEC_LEVELS) * nQ))) return rc;
        if ((rc = grow(ctx, &ctx->d_qcnt, &ctx->cap_qcnt, 2 * nQ))) return rc; /* [nQ] append counters, [nQ] top-k counts */
        /* [nQ] static offsets, [nQ] active offsets, [nQ] queries whose speculative threshold failed, [1] their count */
        if ((rc = grow(ctx, &ctx->d_off, &ctx->cap_off, 3 * nQ + 4))) return rc;
        if ((rc = grow(ctx, &ctx->d_nmax, &ctx->cap_nmax, ctx->nmax_entries + 8))) return rc;
        if (topk && (rc = grow(ctx, &ctx->d_topk, &ctx->cap_topk, nQ * (size_t)k))) return rc;
        if ((rc = grow(ctx, &ctx->d_surv, &ctx->cap_surv, nQ * maxw))) return rc;
        if (!ctx->ranked && !ctx->Tt_valid) {
            const size_t W = 32 * (size_t)ctx->wpl;
            if ((rc = grow(ctx, &ctx->d_Tt, &ctx->cap_Tt, nT * W))) return rc;
            CU(ppp_launch_transpose(ctx->d_T, ctx->d_Tt, (uint32_t)nT, ctx->wpl, ctx->stream));
            ctx->Tt_valid = true;
            launches++;
        }
        uint32_t *d_qcnt = ctx->d_qcnt, *d_topk_cnt = ctx->d_qcnt + nQ;
        uint32_t *d_off_static = ctx->d_off, *d_off_active = ctx->d_off + nQ;
        CU(cudaMemcpyAsync(d_off_static, ctx->nmax_off_host.data(), nQ * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(d_qcnt, 0, 2 * nQ * sizeof(uint32_t), ctx->stream));
        CU(cudaMemsetAsync(d_off_active, 0xff, nQ * sizeof(uint32_t), ctx->stream)); /* PPP_NO_TABLE */
        {
            std::vector<double> two(2 * nQ, 2.0);
            CU(cudaMemcpyAsync(ctx->d_kth, two.data(), 2 * nQ * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        const double thresh_ln = f->thresh * 2.302585092994045684;
        size_t processed = 0;
        bool seeded_pass = false;
        ctx->ev_used = 0;
        ctx->spans.clear();
        CU(cudaMemsetAsync(ctx->d_totals, 0, (CNT_N + 1) * sizeof(unsigned long long), ctx->stream));
        auto fill_filtered = [&](SweepParams &prm, size_t t0, size_t t1) {
            base_params(ctx, &prm);
            prm.hits = ctx->d_hits;
            prm.t0 = (uint32_t)t0;
            prm.t1 = (uint32_t)t1;
            prm.append = topk ? 1 : 2;
            prm.accept_mode = topk ? 0 : 1;
            prm.thresh_log10 = f->thresh;
            prm.kth = ctx->d_kth;
            prm.nmax = ctx->d_nmax;
            prm.nmax_off = d_off_active;
            prm.qcnt = d_qcnt;
            prm.qcap = (uint32_t)maxw;
        };
        /* smallest rank r with P(Poisson(k seen / nT) >= r) <= spec_eps: the r-th best score (or bound) over `seen` exchangeable
         * targets then lies above the final k-th best score except with that probability */
        /* a query whose speculation fails is scanned twice over ALL nT targets, while a looser threshold only costs a few more full
         * evaluations: the tolerated failure probability shrinks with the size of the run (measured at 10,000 x 1,000,000:
         * 5000 ppm 382.3 ms, 1000 ppm 379.7, 500 ppm 374.4 with no query re-scanned; at 10,000 x 125,000 5000 ppm is best) */
        const double spec_eps = 1e-6 * ctx->spec_eps_ppm * std::min(1.0, 131072.0 / (double)nT);
        auto poisson_rank = [&](size_t seen) {
            const double lam = (double)k * (double)seen / (double)nT;
            double term = exp(-lam), cdf = term; /* P(X <= 0) */
            uint32_t r = 1;
            while (1.0 - cdf > spec_eps && r < k) {
                term *= lam / (double)r;
                cdf += term;
                ++r;
            }
            return r;
        };
        if (topk) {
            /* ---- top-k: no host synchronisation until the lists are final ---- */
            /* very short lists (ppp2d keeps the best 2 targets): k-proportional sizes would seed from a few hundred targets, and
             * thresholds that loose send most chunks of a dense query down the slow path of the screen -- floors measured on
             * config 4 (2,000 prefix queries x 500,000 targets, k = 2: 55.8 -> 48.4 ms) */
            const size_t seed_floor = k <= 16 ? 1024 : 256, chunk1_floor = k <= 16 ? 4096 : 1024;
            const size_t chunk1 = std::max<size_t>(chunk1_floor, (size_t)ctx->chunk1_mult * k);      /* first exact chunk */
            const size_t w1 = std::min(std::min(maxw, nT), chunk1);
            const bool spec_ok = ctx->speculate && !ctx->ranked && !ctx->force_exact;
            /* The first chunk is speculative as well when the bulk will be (its verification covers both): its thresholds are
             * the rs-th smallest upper bounds of the seeding pass instead of the k-th -- at config 3 the guaranteed bound lets
             * 15 % of the first chunk's pairs through to a full evaluation, the speculative one 0.8 %. */
            const bool may_spec_first = spec_ok && ctx->spec_first && w1 >= 4 * (size_t)k && nT - w1 >= 8 * w1;
            /* targets of the seeding pass: when only the guaranteed k-th bound is used (no speculation) it pays to seed from more */
            const size_t first = std::min<size_t>(4096, std::max<size_t>(seed_floor, (size_t)(may_spec_first ? ctx->seed_mult : std::max(ctx->seed_mult, 9)) * k));
            const size_t w0 = std::min<size_t>(std::min(maxw, nT), first);
            double *d_kspec = ctx->d_kth + 2 * nQ, *d_kseed = ctx->d_kth + 4 * nQ, *d_kseed_use = ctx->d_kth + 5 * nQ;
            const bool chunk1_spec = may_spec_first && w0 >= (size_t)k;
            size_t chunk1_end = 0;
            /* the first chunk runs in two launches: a short head, after which queries whose k best scores are already final
             * (exact zeros: every later target loses the tie on its index) stop sending pairs to the exact tier */
            const size_t chunk0 = std::max<size_t>(256, (size_t)ctx->chunk0_mult * k);
            if (w0 >= (size_t)k) {
                /* Seeding pass: the k-th smallest UPPER bound (float enclosure only, no series, no exact tier) over the
                 * first targets is a valid threshold for the query's k-th best score; the provisional records are then
                 * dropped and every target -- these included -- goes through the screened path. */
                SweepParams prm;
                fill_filtered(prm, 0, w0);
                prm.bounds_only = 1;
                prm.append = 3; /* dense matrix of provisional scores in d_hits; nothing is appended, no record is kept */
                if ((rc = launch_chunk_deferred(ctx, &prm, &launches, false))) return rc;
                sweep_launches++;
                {
                    size_t ea, eb;
                    if ((rc = mark(ctx, &ea))) return rc;
                    CU(ppp_launch_seed_kth(ctx->d_hits, (uint32_t)w0, k, ctx->d_kth, chunk1_spec ? poisson_rank(w0) : 0u,
                                           chunk1_spec ? d_kseed : nullptr, (uint32_t)nQ, ctx->nsm, ctx->stream));
                    if ((rc = mark(ctx, &eb))) return rc;
                    ctx->spans.push_back({ea, eb, 3});
                }
                launches++;
                seeded_pass = true;
            }
            uint32_t *d_failed = ctx->d_off + 2 * nQ, *d_nfailed = ctx->d_off + 3 * nQ;
            bool spec_on = false;
            size_t spec_start = 0;
            /* speculation levels of the bulk: level l covers targets [lvl_begin[l], lvl_begin[l + 1]) and was screened against
             * thresholds whose initial values are kept in d_lvl(l) (the re-scan's lower acceptance bound for that range) */
            std::vector<size_t> lvl_begin;
            auto d_lvl = [&](size_t l) { return ctx->d_kth + (6 + l) * nQ; };
            ctx->spec_rank = 0;
            ctx->spec_failed = 0;
            while (processed < nT) {
                /* a new level: the first one as soon as the lists can support it, later ones every time the processed targets have
                 * grown respec_mult-fold (the r-th best of 5 x more targets is a much tighter threshold: at config 3 the pairs that
                 * need a full evaluation drop from 355 to 165 per query) while enough targets remain to pay for the staircases */
                const bool first_level = spec_ok && !spec_on && processed >= 4 * (size_t)k && nT - processed >= 8 * processed &&
                                         (!chunk1_spec || processed >= w1);
                const bool respec = ctx->respec_mult > 1 && (double)nQ * (double)nT >= ldexp(1.0, ctx->respec_min_pairs_log2); /* small runs: the extra launches cost more */
                const bool next_level = spec_on && respec && lvl_begin.size() < PPP_MAX_SPEC_LEVELS &&
                                        processed >= (size_t)ctx->respec_mult * lvl_begin.back() && 2 * (nT - processed) >= processed;
                if (first_level || next_level) {
                    /* speculative threshold (ppp_topk.cu): rank r of the list so far, never above the threshold in use */
                    const uint32_t r = poisson_rank(processed);
                    const double *prev = next_level ? d_kspec : (chunk1_end ? d_kseed_use : nullptr);
                    CU(ppp_launch_spec_init(ctx->d_topk, d_topk_cnt, ctx->d_kth, prev, k, r, d_kspec, d_lvl(lvl_begin.size()), (uint32_t)nQ,
                                            ctx->stream));
                    launches++;
                    if (first_level) spec_start = processed;
                    spec_on = true;
                    lvl_begin.push_back(processed);
                    ctx->spec_rank = r;
                }
                size_t w = std::min(maxw, nT - processed);
                if (spec_on && respec && lvl_begin.size() < PPP_MAX_SPEC_LEVELS) {
                    /* stop at the next level's boundary if there will be one */
                    const size_t b = (size_t)ctx->respec_mult * lvl_begin.back();
                    if (b > processed && b < nT && 2 * (nT - b) >= b) w = std::min(w, b - processed);
                }
                if (!spec_on) w = std::min(w, processed ? 4 * processed : (seeded_pass ? chunk1 : first));
                const bool first_spec = chunk1_spec && seeded_pass && processed < w1;
                if (first_spec) {
                    chunk1_end = w1;
                    w = w1 - processed;
                    if (!processed && ctx->chunk0_mult && chunk0 < w1) w = chunk0;
                    if (!processed) CU(cudaMemcpyAsync(d_kseed_use, d_kseed, nQ * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
                    else CU(ppp_launch_spec_update(d_kseed_use, ctx->d_kth, (uint32_t)nQ, ctx->stream));
                    if (processed) launches++;
                }
                const double *kth_used = spec_on ? d_kspec : first_spec ? d_kseed_use : ctx->d_kth;
                size_t ea, eb;
                if (processed || seeded_pass) {
                    if ((rc = mark(ctx, &ea))) return rc;
                    CU(ppp_launch_staircase(ctx->d_qc, ctx->d_lf, kth_used, 0, 0.0, d_off_static, ctx->d_nmax, d_off_active,
                                            ctx->d_kth + nQ, (uint32_t)nQ, ctx->nsm, ctx->stream));
                    if ((rc = mark(ctx, &eb))) return rc;
                    ctx->spans.push_back({ea, eb, 2});
                    launches++;
                }
                SweepParams prm;
                fill_filtered(prm, processed, processed + w);
                prm.kth = kth_used;
                if (spec_on || first_spec) prm.accept_mode = 3; /* speculative threshold: ties at the threshold are kept */
                /* without thresholds (fewer targets than k so far): tile kernel over every pair; otherwise thread-per-pair screen */
                if ((rc = launch_chunk_deferred(ctx, &prm, &launches, !ctx->ranked && (processed || seeded_pass)))) return rc;
                sweep_launches++;
                if ((rc = mark(ctx, &ea))) return rc;
                CU(ppp_launch_topk_merge(ctx->d_hits, d_qcnt, (uint32_t)maxw, ctx->d_topk, d_topk_cnt, ctx->d_kth, k, (uint32_t)nQ,
                                         0u, ctx->nsm, ctx->stream));
                if (spec_on) {
                    CU(ppp_launch_spec_update(d_kspec, ctx->d_kth, (uint32_t)nQ, ctx->stream));
                    launches++;
                }
                if ((rc = mark(ctx, &eb))) return rc;
                ctx->spans.push_back({ea, eb, 3});
                launches++;
                processed += w;
            }
            if (spec_on) {
                /* verify; re-scan the queries whose list is not provably exact with the guaranteed bound */
                uint32_t nfailed = 0;
                CU(cudaMemsetAsync(d_nfailed, 0, sizeof(uint32_t), ctx->stream));
                /* the last level's thresholds are the smallest: a full list whose k-th score is below them is exact */
                CU(ppp_launch_spec_verify(ctx->d_topk, d_topk_cnt, d_lvl(lvl_begin.size() - 1), k, ctx->d_kth + nQ, d_failed, d_nfailed,
                                          (uint32_t)nQ, ctx->stream));
                launches++;
                CU(cudaMemcpyAsync(&nfailed, d_nfailed, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
                CU(cudaStreamSynchronize(ctx->stream));
                ctx->spec_failed = nfailed;
                /* re-scan: every range against the thresholds it was screened with (pairs below them are already listed): the
                 * first chunk, then level by level */
                lvl_begin.push_back(nT);
                size_t lvl = 0;
                for (size_t t0 = chunk1_end ? 0 : spec_start; nfailed && t0 < nT;) {
                    const bool in_first = t0 < chunk1_end;
                    while (!in_first && t0 >= lvl_begin[lvl + 1]) ++lvl;
                    const size_t t1 = in_first ? chunk1_end : std::min(lvl_begin[lvl + 1], t0 + maxw);
                    size_t ea, eb;
                    if ((rc = mark(ctx, &ea))) return rc;
                    CU(ppp_launch_staircase(ctx->d_qc, ctx->d_lf, ctx->d_kth, 0, 0.0, d_off_static, ctx->d_nmax, d_off_active,
                                            ctx->d_kth + nQ, (uint32_t)nQ, ctx->nsm, ctx->stream));
                    if ((rc = mark(ctx, &eb))) return rc;
                    ctx->spans.push_back({ea, eb, 2});
                    SweepParams prm;
                    fill_filtered(prm, t0, t1);
                    prm.accept_mode = 2;
                    prm.klo = in_first ? d_kseed : d_lvl(lvl);
                    if ((rc = launch_chunk_deferred(ctx, &prm, &launches, true, d_failed, nfailed))) return rc;
                    sweep_launches++;
                    if ((rc = mark(ctx, &ea))) return rc;
                    CU(ppp_launch_topk_merge(ctx->d_hits, d_qcnt, (uint32_t)maxw, ctx->d_topk, d_topk_cnt, ctx->d_kth, k, (uint32_t)nQ,
                                             0u, ctx->nsm, ctx->stream));
                    if ((rc = mark(ctx, &eb))) return rc;
                    ctx->spans.push_back({ea, eb, 3});
                    launches += 2;
                    t0 = (in_first && spec_start > t1) ? spec_start : t1;
                }
            }
            unsigned long long c[CNT_N + 1];
            CU(cudaMemcpyAsync(c, ctx->d_totals, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < CNT_N; ++i) totals[i] += c[i];
            ctx->survivors = c[CNT_N];
            if (c[CNT_OVERFLOW]) return fail(ctx, PPP_ERR_CUDA, "ppp_run: internal buffer overflow (%llu records dropped)", c[CNT_OVERFLOW]);
            for (const ppp_ctx::Span &sp : ctx->spans) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ctx->evpool[sp.a], ctx->evpool[sp.b]);
                if (sp.phase == 4) ctx->prefilter_us += 1e3 * ms;
                else ctx->phase_us[sp.phase] += 1e3 * ms;
                if (sp.phase == 0) sweep_ms += ms;
            }
            /* the lists stay on the device (ppp_fetch copies them straight into the caller's buffer) */
            ctx->last_topk_k = k;
            ctx->last_topk_nq = nQ;
            ctx->topk_counts.resize(nQ);
            CU(cudaMemcpyAsync(ctx->topk_counts.data(), d_topk_cnt, nQ * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            size_t n = 0;
            for (size_t q = 0; q < nQ; ++q) n += ctx->topk_counts[q];
            ctx->n_results = n;
            ctx->results_on_host = false;
            ctx->results_topk = true;
        } else {
            /* ---- threshold: one staircase build, survivors of every chunk are copied out as they come ---- */
            CU(ppp_launch_staircase(ctx->d_qc, ctx->d_lf, ctx->d_kth, 1, thresh_ln, d_off_static, ctx->d_nmax, d_off_active,
                                    ctx->d_kth + nQ, (uint32_t)nQ, ctx->nsm, ctx->stream));
            launches++;
            /* The hits are ranked by (q, score, t) on the device (ppp_rank.cu) and stay there until ppp_fetch copies them straight
             * into the caller's buffer.  One launch: ranked where the kernels appended them.  Several launches (more than 2^29
             * pairs): each launch's hits are first gathered in a device buffer; should that not fit (or exceed the sort's 2^31
             * records) the chunks go to the host and are ordered there. */
            const bool one_chunk = nT <= maxw;
            bool on_device = ctx->device_rank != 0;
            size_t n_acc = 0;
            auto spill_to_host = [&](size_t n_dev, const ppp_hit *src) -> int {
                if (!n_dev) return PPP_OK;
                const size_t old = ctx->host_results.size();
                ctx->host_results.resize(old + n_dev);
                CU(cudaMemcpyAsync(ctx->host_results.data() + old, src, n_dev * sizeof(ppp_hit), cudaMemcpyDeviceToHost, ctx->stream));
                CU(cudaStreamSynchronize(ctx->stream));
                return PPP_OK;
            };
            while (processed < nT) {
                const size_t w = std::min(maxw, nT - processed);
                SweepParams prm;
                fill_filtered(prm, processed, processed + w);
                unsigned long long c[CNT_N];
                if ((rc = run_chunk(ctx, &prm, totals, &sweep_ms, &launches, c, !ctx->ranked))) return rc;
                sweep_launches++;
                const size_t n = c[CNT_APPEND];
                processed += w;
                if (on_device && one_chunk) {
                    n_acc = n;
                    break;
                }
                if (on_device && n) {
                    if (n_acc + n > ((size_t)1 << 30)) on_device = false;
                    else if (n_acc + n > ctx->cap_acc) {
                        /* grow the gathering buffer, keeping what it holds */
                        const size_t want = std::max(n_acc + n, 2 * ctx->cap_acc);
                        ppp_hit *bigger = nullptr;
                        if (cudaMalloc(&bigger, want * sizeof(ppp_hit)) != cudaSuccess) {
                            (void)cudaGetLastError();
                            on_device = false;
                        } else {
                            if (n_acc) CU(cudaMemcpyAsync(bigger, ctx->d_acc, n_acc * sizeof(ppp_hit), cudaMemcpyDeviceToDevice, ctx->stream));
                            CU(cudaStreamSynchronize(ctx->stream));
                            cudaFree(ctx->d_acc);
                            ctx->d_acc = bigger;
                            ctx->cap_acc = want;
                        }
                    }
                    if (!on_device) { /* what was gathered so far goes to the host with everything that follows */
                        if ((rc = spill_to_host(n_acc, ctx->d_acc))) return rc;
                        n_acc = 0;
                    }
                }
                if (on_device) {
                    if (n) CU(cudaMemcpyAsync(ctx->d_acc + n_acc, ctx->d_hits, n * sizeof(ppp_hit), cudaMemcpyDeviceToDevice, ctx->stream));
                    n_acc += n;
                } else if ((rc = spill_to_host(n, ctx->d_hits))) {
                    return rc;
                }
            }
            if (on_device) {
                const ppp_hit *src = one_chunk ? ctx->d_hits : ctx->d_acc;
                if (n_acc) {
                    if ((rc = grow(ctx, &ctx->d_ranked, &ctx->cap_ranked, n_acc))) return rc;
                    if ((rc = grow(ctx, &ctx->d_rank_tmp, &ctx->cap_rank_tmp, ppp_rank_scratch_words(n_acc)))) return rc;
                    CU(cudaEventRecord(ctx->ev[5], ctx->stream));
                    CU(ppp_launch_rank_hits(src, n_acc, (uint32_t)nQ, (uint32_t)nT, ctx->d_ranked, ctx->d_rank_tmp, ctx->nsm, ctx->stream,
                                            &launches));
                    CU(cudaEventRecord(ctx->ev[6], ctx->stream));
                    CU(cudaStreamSynchronize(ctx->stream));
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]);
                    ctx->phase_us[3] += 1e3 * ms;
                }
                ctx->n_results = n_acc;
                ctx->results_on_host = false;
                ctx->results_ranked = true;
            } else {
                std::sort(ctx->host_results.begin(), ctx->host_results.end(), hit_less);
                ctx->n_results = ctx->host_results.size();
            }
        }
    }
    if (ctx->t_base && f->mode != PPP_FILTER_TOPK) { /* sub-context of a group: records carry global target numbers (top-k: the merge adds them) */
        if (ctx->results_on_host) {
            for (ppp_hit &h : ctx->host_results) h.t += ctx->t_base;
        } else {
            CU(ppp_launch_add_t(ctx->results_ranked ? ctx->d_ranked : ctx->d_hits, ctx->n_results, ctx->t_base, ctx->nsm, ctx->stream));
            launches++;
        }
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float total_ms = 0.f;
    cudaEventElapsedTime(&total_ms, ctx->ev[0], ctx->ev[1]);

    memcpy(ctx->last_counters, totals, sizeof(totals));
    ctx->stats.pairs = (uint64_t)nQ * nT;
    ctx->stats.candidates = totals[CNT_CAND];
    ctx->stats.accurate_evals = totals[CNT_ACCURATE];
    ctx->stats.exact_pairs = totals[CNT_EXACT];
    ctx->stats.kernel_launches = launches;
    ctx->stats.sweep_kernel_ms = sweep_ms;
    ctx->stats.total_device_ms = total_ms;
    ctx->stats.sweep_launches = sweep_launches;
    return PPP_OK;
}

extern "C" int ppp_result_count(ppp_ctx *ctx, size_t *nout)
{
    if (!ctx || !nout) return PPP_ERR_INVALID;
    *nout = ctx->n_results;
    return PPP_OK;
}

extern "C" int ppp_fetch(ppp_ctx *ctx, ppp_hit *out, size_t cap, size_t *nout)
{
    if (ctx && !ctx->subs.empty()) return group_fetch(ctx, out, cap, nout);
    if (!ctx) return PPP_ERR_INVALID;
    if (nout) *nout = ctx->n_results;
    if (cap < ctx->n_results) return fail(ctx, PPP_ERR_CAPACITY, "ppp_fetch: need room for %zu hits, got %zu", ctx->n_results, cap);
    if (!ctx->n_results) return PPP_OK;
    if (!out) return fail(ctx, PPP_ERR_INVALID, "ppp_fetch: out is NULL");
    CU(cudaSetDevice(ctx->device));
    if (ctx->results_topk) {
        const size_t nQ = ctx->last_topk_nq, k = ctx->last_topk_k;
        if (ctx->n_results == nQ * k) { /* every list is full: one copy straight into the caller's buffer */
            CU(cudaMemcpyAsync(out, ctx->d_topk, nQ * k * sizeof(ppp_hit), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        } else {
            std::vector<ppp_hit> lists(nQ * k);
            CU(cudaMemcpyAsync(lists.data(), ctx->d_topk, lists.size() * sizeof(ppp_hit), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            size_t o = 0;
            for (size_t q = 0; q < nQ; ++q) {
                memcpy(out + o, lists.data() + q * k, ctx->topk_counts[q] * sizeof(ppp_hit));
                o += ctx->topk_counts[q];
            }
        }
    } else if (ctx->results_on_host) {
        memcpy(out, ctx->host_results.data(), ctx->n_results * sizeof(ppp_hit));
    } else {
        CU(cudaMemcpyAsync(out, ctx->results_ranked ? ctx->d_ranked : ctx->d_hits, ctx->n_results * sizeof(ppp_hit), cudaMemcpyDeviceToHost,
                           ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return PPP_OK;
}

extern "C" int ppp_score(ppp_ctx *ctx, const uint32_t *P, const uint32_t *Y, const double *p, size_t nQ, const ppp_filter *f,
                         ppp_hit *out, size_t cap, size_t *nout)
{
    int rc = ppp_set_queries(ctx, P, Y, p, nQ);
    if (rc) return rc;
    rc = ppp_run(ctx, f);
    if (rc) return rc;
    return ppp_fetch(ctx, out, cap, nout);
}

extern "C" int ppp_topk_device(ppp_ctx *ctx, const ppp_hit **lists, const uint32_t **counts, uint32_t *k, size_t *nQ)
{
    if (ctx && !ctx->subs.empty()) return fail(ctx, PPP_ERR_STATE, "ppp_topk_device: a multi-device context merges its lists itself (ppp_fetch)");
    if (!ctx || !lists || !counts || !k || !nQ) return PPP_ERR_INVALID;
    if (!ctx->last_topk_k) return fail(ctx, PPP_ERR_STATE, "ppp_topk_device: the last run was not a top-k run");
    *lists = ctx->d_topk;
    *counts = ctx->d_qcnt + ctx->last_topk_nq;
    *k = ctx->last_topk_k;
    *nQ = ctx->last_topk_nq;
    return PPP_OK;
}

/* merge `nlists` device-resident lists into ctx->d_mtopk [nQ][k] / ctx->d_mcnt [nQ] (on ctx's device and stream; no sync) */
static int merge_lists_on_device(ppp_ctx *ctx, const ppp_hit *lists, const uint32_t *counts, uint32_t nlists, const uint32_t *t_offsets,
                                 size_t nQ, uint32_t k)
{
    int rc;
    if ((rc = grow(ctx, &ctx->d_mtopk, &ctx->cap_mtopk, nQ * (size_t)k))) return rc;
    if ((rc = grow(ctx, &ctx->d_mkth, &ctx->cap_mkth, nQ))) return rc;
    if ((rc = grow(ctx, &ctx->d_mcnt, &ctx->cap_mcnt, nQ * ((size_t)nlists + 1)))) return rc;
    cudaError_t me = ppp_launch_topk_merge_lists(lists, counts, nlists, t_offsets, (uint32_t)nQ, k, ctx->d_mtopk, ctx->d_mcnt, ctx->nsm,
                                                 ctx->stream);
    if (me == cudaErrorNotSupported) {
        /* too many records for one sort: merge list by list; the merge kernel consumes (zeroes) its input counters,
         * so work on a copy */
        CU(cudaMemcpyAsync(ctx->d_mcnt + nQ, counts, nQ * nlists * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_mcnt, 0, nQ * sizeof(uint32_t), ctx->stream));
        for (uint32_t l = 0; l < nlists; ++l)
            CU(ppp_launch_topk_merge(lists + (size_t)l * nQ * k, ctx->d_mcnt + nQ * ((size_t)l + 1), k, ctx->d_mtopk, ctx->d_mcnt, ctx->d_mkth,
                                     k, (uint32_t)nQ, t_offsets[l], ctx->nsm, ctx->stream));
    } else {
        CU(me);
    }
    return PPP_OK;
}

extern "C" int ppp_merge_topk_device(ppp_ctx *ctx, const ppp_hit *lists, const uint32_t *counts, uint32_t nlists,
                                     const uint32_t *t_offsets, size_t nQ, uint32_t k, ppp_hit *out, uint32_t *out_counts)
{
    if (!ctx || !lists || !counts || !t_offsets || !out || !out_counts || !nlists || !nQ) return PPP_ERR_INVALID;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];
    if (k == 0 || (int)k > ppp_topk_max_k()) return fail(ctx, PPP_ERR_INVALID, "ppp_merge_topk_device: k out of range");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = merge_lists_on_device(ctx, lists, counts, nlists, t_offsets, nQ, k))) return rc;
    CU(cudaMemcpyAsync(out, ctx->d_mtopk, nQ * (size_t)k * sizeof(ppp_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_counts, ctx->d_mcnt, nQ * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PPP_OK;
}

/* developer statistics of the queued pre-filter (counted only in a -DPF2_STATS build): pushes, drains, drain
 * iterations, word-level passes.  The first call arms the counters. */
extern "C" int ppp_debug_pf2stats(ppp_ctx *ctx, unsigned long long *out8, int reset)
{
    if (ctx && !ctx->subs.empty()) ctx = ctx->subs[0];
    if (!ctx || !out8) return PPP_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d_pfstats) {
        CU(cudaMalloc(&ctx->d_pfstats, 8 * sizeof(unsigned long long)));
        CU(cudaMemset(ctx->d_pfstats, 0, 8 * sizeof(unsigned long long)));
    }
    CU(cudaMemcpy(out8, ctx->d_pfstats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) CU(cudaMemset(ctx->d_pfstats, 0, 8 * sizeof(unsigned long long)));
    return PPP_OK;
}

extern "C" int ppp_printed_cutoffs(ppp_ctx *ctx, double slope, uint64_t *row_off, int32_t *milli, int32_t *marker, int32_t *cutoff_milli,
                                   uint8_t *status)
{
    if (ctx && !ctx->subs.empty()) return fail(ctx, PPP_ERR_STATE, "ppp_printed_cutoffs: needs a single-device context (the hits of a group are ranked on the host)");
    if (!ctx || !marker || !cutoff_milli || !status) return PPP_ERR_INVALID;
    if (!ctx->results_ranked) return fail(ctx, PPP_ERR_STATE, "ppp_printed_cutoffs: the last run was not a threshold run ranked on the device");
    if (isnan(slope)) return fail(ctx, PPP_ERR_INVALID, "ppp_printed_cutoffs: slope is NaN");
    CU(cudaSetDevice(ctx->device));
    const size_t n = ctx->n_results, nQ = ctx->nQ;
    /* scratch (the sort's buffer is free again): [nQ + 1] u64 row offsets, [n] milli, [nQ] marker, [nQ] cutoff, [nQ] status */
    const size_t words = 2 * (nQ + 1) + n + 2 * nQ + (nQ + 3) / 4 + 8;
    int rc;
    if ((rc = grow(ctx, &ctx->d_rank_tmp, &ctx->cap_rank_tmp, words))) return rc;
    unsigned long long *d_off = reinterpret_cast<unsigned long long *>(ctx->d_rank_tmp);
    int32_t *d_milli = reinterpret_cast<int32_t *>(d_off + nQ + 1);
    int32_t *d_marker = d_milli + n, *d_cut = d_marker + nQ;
    uint8_t *d_status = reinterpret_cast<uint8_t *>(d_cut + nQ);
    CU(ppp_launch_rank_cutoffs(ctx->d_ranked, n, (uint32_t)nQ, slope, d_off, d_milli, d_marker, d_cut, d_status, ctx->nsm, ctx->stream));
    if (row_off) CU(cudaMemcpyAsync(row_off, d_off, (nQ + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (milli && n) CU(cudaMemcpyAsync(milli, d_milli, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(marker, d_marker, nQ * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(cutoff_milli, d_cut, nQ * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(status, d_status, nQ, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return PPP_OK;
}

extern "C" int ppp_get_stats(const ppp_ctx *ctx, ppp_stats *out)
{
    if (!ctx || !out) return PPP_ERR_INVALID;
    *out = ctx->stats;
    return PPP_OK;
}

extern "C" int ppp_query_count(const ppp_ctx *ctx, size_t *nQ)
{
    if (!ctx || !nQ) return PPP_ERR_INVALID;
    *nQ = ctx->nQ;
    return PPP_OK;
}

extern "C" int ppp_row_words(const ppp_ctx *ctx, size_t *words)
{
    if (!ctx || !words) return PPP_ERR_INVALID;
    *words = ctx->G ? ppp_words_per_row(ctx->G) : 0;
    return PPP_OK;
}

extern "C" int ppp_set_option(ppp_ctx *ctx, const char *name, int64_t value)
{
    if (ctx && !ctx->subs.empty()) return group_set_option(ctx, name, value);
    if (!ctx || !name) return PPP_ERR_INVALID;
    if (!strcmp(name, "force_exact")) ctx->force_exact = value != 0;
    else if (!strcmp(name, "check_bounds")) ctx->check_bounds = value != 0;
    else if (!strcmp(name, "table_mode")) {
        if (value < -1 || value > 1) return fail(ctx, PPP_ERR_INVALID, "table_mode must be -1 (auto), 0 (off) or 1 (on)");
        ctx->table_mode = (int)value;
    } else if (!strcmp(name, "seed_mult") || !strcmp(name, "chunk1_mult") || !strcmp(name, "spec_eps_ppm")) {
        if (value < 1 || value > 1000000) return fail(ctx, PPP_ERR_INVALID, "%s out of range", name);
        (name[1] == 'e' ? ctx->seed_mult : name[0] == 'c' ? ctx->chunk1_mult : ctx->spec_eps_ppm) = (int)value;
    } else if (!strcmp(name, "speculate")) {
        ctx->speculate = value != 0;
    } else if (!strcmp(name, "chunk0_mult")) {
        if (value < 0 || value > 1000000) return fail(ctx, PPP_ERR_INVALID, "chunk0_mult out of range");
        ctx->chunk0_mult = (int)value;
    } else if (!strcmp(name, "respec_min_pairs_log2")) {
        if (value < 0 || value > 62) return fail(ctx, PPP_ERR_INVALID, "respec_min_pairs_log2 out of range");
        ctx->respec_min_pairs_log2 = (int)value;
    } else if (!strcmp(name, "respec_mult")) {
        if (value < 0 || value > 1000000) return fail(ctx, PPP_ERR_INVALID, "respec_mult out of range");
        ctx->respec_mult = (int)value;
    } else if (!strcmp(name, "spec_first")) {
        ctx->spec_first = value != 0;
    } else if (!strcmp(name, "dup")) {
        ctx->dup = value != 0;
    } else if (!strcmp(name, "launch_pairs_log2")) {
        if (value != 0 && (value < 12 || value > 31)) return fail(ctx, PPP_ERR_INVALID, "launch_pairs_log2 must be 0 (automatic) or in [12, 31]");
        ctx->launch_pairs_log2 = (int)value;
    } else if (!strcmp(name, "device_rank")) {
        ctx->device_rank = value != 0;
    } else if (!strcmp(name, "prefilter_tc")) {
        if (value < 0 || value > 3) return fail(ctx, PPP_ERR_INVALID, "prefilter_tc must be 0 .. 3 (bit 0: all-ones launch, bit 1: general launch)");
        ctx->prefilter_tc = (int)value;
    } else if (!strcmp(name, "pf_overlap")) {
        ctx->pf_overlap = value != 0;
    } else if (!strcmp(name, "tc_exact")) {
        ctx->tc_exact = value != 0;
    } else if (!strcmp(name, "sweep_warps")) {
        if (value < 1 || value > 16) return fail(ctx, PPP_ERR_INVALID, "sweep_warps must be in [1,16]");
        ctx->sweep_warps = (int)value;
    } else
        return fail(ctx, PPP_ERR_INVALID, "unknown option '%s'", name);
    return PPP_OK;
}

extern "C" int ppp_get_counter(const ppp_ctx *ctx, const char *name, int64_t *value)
{
    if (ctx && !ctx->subs.empty()) return group_get_counter(ctx, name, value);
    if (!ctx || !name || !value) return PPP_ERR_INVALID;
    int i = -1;
    if (!strcmp(name, "exact_pairs")) i = CNT_EXACT;
    else if (!strcmp(name, "candidates")) i = CNT_CAND;
    else if (!strcmp(name, "accurate_evals")) i = CNT_ACCURATE;
    else if (!strcmp(name, "bound_violations")) i = CNT_VIOL;
    else if (!strcmp(name, "bound_checks")) i = CNT_CHECKS;
    else if (!strcmp(name, "full_pairs")) i = CNT_FULL;
    else if (!strcmp(name, "num_sms")) { *value = ctx->nsm; return PPP_OK; }
    else if (!strcmp(name, "used_table")) { *value = ctx->last_used_table; return PPP_OK; }
    else if (!strcmp(name, "pf_tile_q_gen")) { *value = ctx->tile_q[0]; return PPP_OK; }
    else if (!strcmp(name, "pf_tile_q_all")) { *value = ctx->tile_q[1]; return PPP_OK; }
    else if (!strcmp(name, "pf_tile_q_pyy")) { *value = ctx->tile_q[2]; return PPP_OK; }
    else if (!strcmp(name, "pf_tile_q_any")) { *value = ctx->tile_q_any; return PPP_OK; }
    else if (!strcmp(name, "spec_rank")) { *value = ctx->spec_rank; return PPP_OK; }
    else if (!strcmp(name, "spec_failed")) { *value = ctx->spec_failed; return PPP_OK; }
    else if (!strcmp(name, "tc_launches")) { *value = (int64_t)ctx->tc_launches; return PPP_OK; }
    else if (!strcmp(name, "us_prefilter")) { *value = (int64_t)ctx->prefilter_us; return PPP_OK; }
    else if (!strcmp(name, "prefilter_pairs")) { *value = (int64_t)ctx->prefilter_pairs; return PPP_OK; }
    else if (!strcmp(name, "survivors")) { *value = (int64_t)ctx->survivors; return PPP_OK; }
    else if (!strcmp(name, "us_sweep")) { *value = (int64_t)ctx->phase_us[0]; return PPP_OK; }
    else if (!strcmp(name, "us_exact")) { *value = (int64_t)ctx->phase_us[1]; return PPP_OK; }
    else if (!strcmp(name, "us_staircase")) { *value = (int64_t)ctx->phase_us[2]; return PPP_OK; }
    else if (!strcmp(name, "us_merge")) { *value = (int64_t)ctx->phase_us[3]; return PPP_OK; }
    if (i < 0) return PPP_ERR_INVALID;
    *value = (int64_t)ctx->last_counters[i];
    return PPP_OK;
}

extern "C" int ppp_debug_bdtrc(ppp_ctx *ctx, const int32_t *k, const int32_t *n, const double *p, double *out, size_t m)
{
    if (ctx && !ctx->subs.empty()) ctx = ctx->subs[0];
    if (!ctx || !k || !n || !p || !out) return PPP_ERR_INVALID;
    if (!m) return PPP_OK;
    CU(cudaSetDevice(ctx->device));
    int32_t *dk = nullptr, *dn = nullptr;
    double *dp = nullptr, *dout = nullptr;
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc(&dk, m * 4)) == cudaSuccess && (e = cudaMalloc(&dn, m * 4)) == cudaSuccess && (e = cudaMalloc(&dp, m * 8)) == cudaSuccess &&
        (e = cudaMalloc(&dout, m * 8)) == cudaSuccess &&
        (e = cudaMemcpyAsync(dk, k, m * 4, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(dn, n, m * 4, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(dp, p, m * 8, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess &&
        (e = ppp_launch_debug_bdtrc(dk, dn, dp, dout, m, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(out, dout, m * 8, cudaMemcpyDeviceToHost, ctx->stream)) == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dk);
    cudaFree(dn);
    cudaFree(dp);
    cudaFree(dout);
    CU(e);
    return PPP_OK;
}

extern "C" int ppp_debug_enclosure(ppp_ctx *ctx, const int32_t *y, const int32_t *n, const double *p, float *lo, float *hi, double *lnf,
                                   size_t m)
{
    if (ctx && !ctx->subs.empty()) ctx = ctx->subs[0];
    if (!ctx || !y || !n || !p || !lo || !hi || !lnf) return PPP_ERR_INVALID;
    if (!m) return PPP_OK;
    for (size_t i = 0; i < m; ++i)
        if (y[i] < 1 || n[i] < y[i] || n[i] > 8192 || !(p[i] > 0.0 && p[i] < 1.0))
            return fail(ctx, PPP_ERR_INVALID, "ppp_debug_enclosure: point %zu outside 1 <= yes <= number <= 8192, 0 < p < 1", i);
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_lf(ctx))) return rc;
    std::vector<QConst> qc(m);
    for (size_t i = 0; i < m; ++i) make_qconst(p[i], 0, 0, &qc[i]);
    QConst *dq = nullptr;
    int32_t *dy = nullptr, *dn = nullptr;
    float *dlo = nullptr, *dhi = nullptr;
    double *dl = nullptr;
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc(&dq, m * sizeof(QConst))) == cudaSuccess && (e = cudaMalloc(&dy, m * 4)) == cudaSuccess &&
        (e = cudaMalloc(&dn, m * 4)) == cudaSuccess && (e = cudaMalloc(&dlo, m * 4)) == cudaSuccess &&
        (e = cudaMalloc(&dhi, m * 4)) == cudaSuccess && (e = cudaMalloc(&dl, m * 8)) == cudaSuccess &&
        (e = cudaMemcpyAsync(dq, qc.data(), m * sizeof(QConst), cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(dy, y, m * 4, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(dn, n, m * 4, cudaMemcpyHostToDevice, ctx->stream)) == cudaSuccess &&
        (e = ppp_launch_debug_enclosure(dq, ctx->d_lf, dy, dn, dlo, dhi, dl, m, ctx->nsm, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(lo, dlo, m * 4, cudaMemcpyDeviceToHost, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(hi, dhi, m * 4, cudaMemcpyDeviceToHost, ctx->stream)) == cudaSuccess &&
        (e = cudaMemcpyAsync(lnf, dl, m * 8, cudaMemcpyDeviceToHost, ctx->stream)) == cudaSuccess)
        e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dq);
    cudaFree(dy);
    cudaFree(dn);
    cudaFree(dlo);
    cudaFree(dhi);
    cudaFree(dl);
    CU(e);
    return PPP_OK;
}

extern "C" int ppp_pipe_peak(ppp_ctx *ctx, const char *pipe, double *ops_per_s)
{
    if (ctx && !ctx->subs.empty()) ctx = ctx->subs[0];
    if (!ctx || !pipe || !ops_per_s) return PPP_ERR_INVALID;
    int kind = -1;
    if (!strcmp(pipe, "popc")) kind = 0;
    else if (!strcmp(pipe, "lop3")) kind = 1;
    else if (!strcmp(pipe, "iadd")) kind = 2;
    else if (!strcmp(pipe, "dfma")) kind = 3;
    if (kind < 0) return fail(ctx, PPP_ERR_INVALID, "ppp_pipe_peak: unknown pipe '%s' (popc, lop3, iadd, dfma)", pipe);
    CU(cudaSetDevice(ctx->device));
    /* FP64 and POPC run at 1/8 .. 1/4 of the ALU rate: size every kernel for a few milliseconds */
    const uint32_t iters = kind == 1 || kind == 2 ? 4096 : 1024;
    double best = 0.0, lane_ops = 0.0;
    for (int rep = 0; rep < 4; ++rep) { /* first repetition warms up; best of the rest */
        CU(cudaEventRecord(ctx->ev[5], ctx->stream));
        CU(ppp_launch_pipe_peak(kind, iters, reinterpret_cast<uint32_t *>(ctx->d_counters), ctx->nsm, ctx->stream, &lane_ops));
        CU(cudaEventRecord(ctx->ev[6], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]);
        if (rep && ms > 0.f) best = std::max(best, lane_ops / (1e-3 * (double)ms));
    }
    *ops_per_s = best;
    return PPP_OK;
}

/* ================================================================================================ multi-device group
 * ppp_init(devices, ndev > 1): ONE process drives ndev GPUs (replaces the reference's own fan-out over worker processes,
 * /root/reference/bin/ppp.pl:374-421, 895-922: chunk files of 50 targets and forked `ppp.pl --client` workers).  The group
 * shards the TARGET axis contiguously over its devices; every device holds its shard resident and receives all queries;
 * the devices run concurrently, each driven by its own host thread on its own stream, with no exchange while scoring.
 * Results: dense and threshold hits are assembled on the host (strided D2H copies / per-query merge of the devices' ranked
 * runs); top-k lists are merged on the devices by QUERY slices -- device d pulls every device's lists of its slice over
 * NVLink peer copies (cudaMemcpyPeerAsync), merges them with the list-merge kernel and copies its slice of the global
 * lists straight into the caller's buffer, so the ndev D2H copies use ndev PCIe links at once. */
template <typename F>
static int for_each_sub(ppp_ctx *g, F fn)
{
    const size_t n = g->subs.size();
    std::vector<int> rc(n, PPP_OK);
    std::vector<std::thread> th;
    for (size_t i = 1; i < n; ++i) th.emplace_back([&, i] { rc[i] = fn(g->subs[i], i); });
    rc[0] = fn(g->subs[0], 0);
    for (std::thread &t : th) t.join();
    for (size_t i = 0; i < n; ++i)
        if (rc[i]) {
            snprintf(g->err, sizeof(g->err), "device %d: %.480s", g->subs[i]->device, g->subs[i]->err);
            return rc[i];
        }
    return PPP_OK;
}

static int group_init(ppp_ctx **out, const int *devices, int ndev)
{
    if (ndev > 16) return fail(nullptr, PPP_ERR_INVALID, "ppp_init: at most 16 devices per context");
    /* a device may be listed more than once: every entry owns its own shard, stream and buffers (that is how the 1-GPU test
     * box exercises the group code) */
    ppp_ctx *g = new (std::nothrow) ppp_ctx();
    if (!g) return fail(nullptr, PPP_ERR_NOMEM, "ppp_init: out of host memory");
    memset(&g->stats, 0, sizeof(g->stats));
    memset(g->last_counters, 0, sizeof(g->last_counters));
    for (int i = 0; i < ndev; ++i) {
        ppp_ctx *s = nullptr;
        const int rc = ppp_init(&s, devices + i, 1);
        if (rc) {
            ppp_destroy(g);
            return rc; /* g_init_err holds the message */
        }
        g->subs.push_back(s);
    }
    /* NVLink / NVSwitch peer access for the list merge (cudaMemcpyPeerAsync falls back to staging through the host without it) */
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < ndev; ++j) {
            int can = 0;
            if (devices[i] != devices[j] && cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
                cudaSetDevice(devices[i]);
                if (cudaDeviceEnablePeerAccess(devices[j], 0) != cudaSuccess) (void)cudaGetLastError(); /* already enabled */
            }
        }
    g->device = devices[0];
    g->nsm = g->subs[0]->nsm;
    *out = g;
    return PPP_OK;
}

static void group_shard(ppp_ctx *g, size_t nT)
{
    const size_t n = g->subs.size(), base = nT / n, rem = nT % n;
    g->sub_t0.assign(n + 1, 0);
    for (size_t i = 0; i < n; ++i) g->sub_t0[i + 1] = g->sub_t0[i] + base + (i < rem ? 1 : 0);
}

static int group_set_targets_bits(ppp_ctx *g, const uint32_t *T, size_t nT, uint32_t G)
{
    if (!T && nT) return fail(g, PPP_ERR_INVALID, "ppp_set_targets_bits: T is NULL");
    if (nT > 0xffffffffull) return fail(g, PPP_ERR_INVALID, "ppp_set_targets_bits: nT too large");
    group_shard(g, nT);
    const size_t wpr = ppp_words_per_row(G);
    g->nT = 0;
    const int rc = for_each_sub(g, [&](ppp_ctx *s, size_t i) {
        s->t_base = (uint32_t)g->sub_t0[i];
        return ppp_set_targets_bits(s, T ? T + g->sub_t0[i] * wpr : nullptr, g->sub_t0[i + 1] - g->sub_t0[i], G);
    });
    if (rc) return rc;
    g->nT = nT;
    g->G = G;
    g->wpl = g->subs[0]->wpl;
    g->ranked = false;
    g->nQ = g->subs[0]->nQ; /* 0 if the layout changed */
    return PPP_OK;
}

static int group_set_targets_bits_streamed(ppp_ctx *g, const uint32_t *T, size_t nT, uint32_t G)
{
    if (!T && nT) return fail(g, PPP_ERR_INVALID, "ppp_set_targets_bits_streamed: T is NULL");
    if (nT > 0xffffffffull) return fail(g, PPP_ERR_INVALID, "ppp_set_targets_bits_streamed: nT too large");
    group_shard(g, nT);
    const size_t wpr = ppp_words_per_row(G);
    g->nT = 0;
    const int rc = for_each_sub(g, [&](ppp_ctx *s, size_t i) {
        s->t_base = (uint32_t)g->sub_t0[i];
        return ppp_set_targets_bits_streamed(s, T ? T + g->sub_t0[i] * wpr : nullptr, g->sub_t0[i + 1] - g->sub_t0[i], G);
    });
    if (rc) return rc;
    g->nT = nT;
    g->G = G;
    g->wpl = g->subs[0]->wpl;
    g->ranked = false;
    g->nQ = g->subs[0]->nQ;
    return PPP_OK;
}

static int group_set_targets_ranked(ppp_ctx *g, const uint16_t *idx, const uint16_t *rawdepth, const uint64_t *row_off, size_t nT, uint32_t G)
{
    if (!row_off) return fail(g, PPP_ERR_INVALID, "ppp_set_targets_ranked: NULL input");
    if (nT > 0xffffffffull) return fail(g, PPP_ERR_INVALID, "ppp_set_targets_ranked: nT too large");
    for (size_t t = 0; t < nT; ++t)
        if (row_off[t + 1] < row_off[t]) return fail(g, PPP_ERR_INVALID, "ppp_set_targets_ranked: row_off not monotone at %zu", t);
    group_shard(g, nT);
    g->nT = 0;
    const int rc = for_each_sub(g, [&](ppp_ctx *s, size_t i) {
        const size_t a = g->sub_t0[i], b = g->sub_t0[i + 1];
        std::vector<uint64_t> off(b - a + 1);
        for (size_t t = a; t <= b; ++t) off[t - a] = row_off[t] - row_off[a];
        s->t_base = (uint32_t)a;
        return ppp_set_targets_ranked(s, idx ? idx + row_off[a] : nullptr, rawdepth ? rawdepth + row_off[a] : nullptr, off.data(), b - a, G);
    });
    if (rc) return rc;
    g->nT = nT;
    g->G = G;
    g->wpl = g->subs[0]->wpl;
    g->ranked = true;
    g->nQ = g->subs[0]->nQ;
    return PPP_OK;
}

static int group_set_queries(ppp_ctx *g, const uint32_t *P, const uint32_t *Y, const double *p, size_t nQ)
{
    if (!g->wpl) return fail(g, PPP_ERR_STATE, "ppp_set_queries: call ppp_set_targets_* first (it fixes G)");
    g->nQ = 0;
    const int rc = for_each_sub(g, [&](ppp_ctx *s, size_t) { return ppp_set_queries(s, P, Y, p, nQ); });
    if (rc) return rc;
    g->nQ = nQ;
    return PPP_OK;
}

static int group_run(ppp_ctx *g, const ppp_filter *f)
{
    if (!g->wpl || !g->nT) return fail(g, PPP_ERR_STATE, "ppp_run: no targets resident");
    if (!g->nQ) return fail(g, PPP_ERR_STATE, "ppp_run: no queries resident");
    ppp_filter all = {PPP_FILTER_ALL, 0, 0.0};
    if (!f) f = &all;
    const size_t nQ = g->nQ, nd = g->subs.size();
    g->group_mode = f->mode;
    g->group_k = f->k;
    g->group_hits.clear();
    g->topk_counts.clear();
    ppp_filter fs = *f;
    int rc = for_each_sub(g, [&](ppp_ctx *s, size_t) {
        if (!s->nT) { /* fewer targets than devices: this device holds none */
            s->n_results = 0;
            s->last_topk_k = 0;
            memset(&s->stats, 0, sizeof(s->stats));
            return (int)PPP_OK;
        }
        return ppp_run(s, &fs);
    });
    if (rc) return rc;
    if (f->mode == PPP_FILTER_ALL) {
        g->n_results = nQ * g->nT;
    } else if (f->mode == PPP_FILTER_THRESH_LOG10) {
        /* every device's hits are ranked by (query, score, target) with global target numbers: merge the runs per query */
        std::vector<std::vector<ppp_hit>> part(nd);
        rc = for_each_sub(g, [&](ppp_ctx *s, size_t i) {
            part[i].resize(s->n_results);
            size_t n = 0;
            return s->n_results ? ppp_fetch(s, part[i].data(), part[i].size(), &n) : (int)PPP_OK;
        });
        if (rc) return rc;
        size_t total = 0;
        for (const auto &v : part) total += v.size();
        g->group_hits.resize(total);
        std::vector<size_t> pos(nd, 0);
        size_t o = 0;
        for (size_t q = 0; q < nQ; ++q) {
            const size_t q_begin = o;
            for (size_t i = 0; i < nd; ++i) { /* shards are in ascending target order: ties on the score stay ordered by target */
                const size_t a = pos[i];
                size_t b = a;
                while (b < part[i].size() && part[i][b].q == q) ++b;
                if (b > a) {
                    memcpy(g->group_hits.data() + o, part[i].data() + a, (b - a) * sizeof(ppp_hit));
                    std::inplace_merge(g->group_hits.begin() + q_begin, g->group_hits.begin() + o, g->group_hits.begin() + o + (b - a), hit_less);
                    o += b - a;
                }
                pos[i] = b;
            }
        }
        g->n_results = total;
    } else {
        /* top-k: device d merges the queries [q0, q1) of its slice from every device's lists (peer copies over NVLink) */
        const uint32_t k = f->k;
        g->topk_counts.assign(nQ, 0);
        std::vector<uint32_t> offs(nd);
        std::vector<size_t> src; /* devices that hold targets */
        for (size_t i = 0; i < nd; ++i)
            if (g->subs[i]->nT) {
                offs[src.size()] = (uint32_t)g->sub_t0[i];
                src.push_back(i);
            }
        rc = for_each_sub(g, [&](ppp_ctx *s, size_t d) {
            const size_t q0 = nQ * d / nd, q1 = nQ * (d + 1) / nd, ns = q1 - q0;
            if (!ns) return (int)PPP_OK;
            ppp_ctx *ctx = s;
            CU(cudaSetDevice(s->device));
            int r;
            /* gathered slices: [src][ns][k] records in d_acc, [src][ns] counts behind them */
            const size_t rec_words = src.size() * ns * (size_t)k;
            if ((r = grow(s, &s->d_acc, &s->cap_acc, rec_words + (src.size() * ns * sizeof(uint32_t) + sizeof(ppp_hit) - 1) / sizeof(ppp_hit)))) return r;
            uint32_t *g_cnt = reinterpret_cast<uint32_t *>(s->d_acc + rec_words);
            for (size_t j = 0; j < src.size(); ++j) {
                ppp_ctx *o = g->subs[src[j]];
                const uint32_t *o_cnt = o->d_qcnt + o->last_topk_nq;
                CU(cudaMemcpyPeerAsync(s->d_acc + j * ns * k, s->device, o->d_topk + q0 * k, o->device, ns * (size_t)k * sizeof(ppp_hit), s->stream));
                CU(cudaMemcpyPeerAsync(g_cnt + j * ns, s->device, o_cnt + q0, o->device, ns * sizeof(uint32_t), s->stream));
            }
            if ((r = merge_lists_on_device(s, s->d_acc, g_cnt, (uint32_t)src.size(), offs.data(), ns, k))) return r;
            CU(cudaMemcpyAsync(g->topk_counts.data() + q0, s->d_mcnt, ns * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            return (int)PPP_OK;
        });
        if (rc) return rc;
        size_t n = 0;
        for (size_t q = 0; q < nQ; ++q) n += g->topk_counts[q];
        g->n_results = n;
    }
    /* statistics: work summed over the devices, times = the slowest device */
    memset(&g->stats, 0, sizeof(g->stats));
    memset(g->last_counters, 0, sizeof(g->last_counters));
    for (ppp_ctx *s : g->subs) {
        if (!s->nT) continue;
        g->stats.candidates += s->stats.candidates;
        g->stats.accurate_evals += s->stats.accurate_evals;
        g->stats.exact_pairs += s->stats.exact_pairs;
        g->stats.kernel_launches += s->stats.kernel_launches;
        g->stats.sweep_launches += s->stats.sweep_launches;
        g->stats.sweep_kernel_ms = std::max(g->stats.sweep_kernel_ms, s->stats.sweep_kernel_ms);
        g->stats.total_device_ms = std::max(g->stats.total_device_ms, s->stats.total_device_ms);
        for (int i = 0; i < CNT_N; ++i) g->last_counters[i] += s->last_counters[i];
    }
    g->stats.pairs = (uint64_t)nQ * g->nT;
    return PPP_OK;
}

static int group_fetch(ppp_ctx *g, ppp_hit *out, size_t cap, size_t *nout)
{
    if (nout) *nout = g->n_results;
    if (cap < g->n_results) return fail(g, PPP_ERR_CAPACITY, "ppp_fetch: need room for %zu hits, got %zu", g->n_results, cap);
    if (!g->n_results) return PPP_OK;
    if (!out) return fail(g, PPP_ERR_INVALID, "ppp_fetch: out is NULL");
    const size_t nQ = g->nQ, nd = g->subs.size();
    if (g->group_mode == PPP_FILTER_THRESH_LOG10) {
        memcpy(out, g->group_hits.data(), g->n_results * sizeof(ppp_hit));
        return PPP_OK;
    }
    if (g->group_mode == PPP_FILTER_ALL) {
        /* dense out[q * nT + t]: device i's [nQ][nT_i] block goes to columns [t0_i, t0_i + nT_i) of every row */
        return for_each_sub(g, [&](ppp_ctx *s, size_t i) {
            ppp_ctx *ctx = s;
            const size_t w = g->sub_t0[i + 1] - g->sub_t0[i];
            if (!w) return (int)PPP_OK;
            CU(cudaSetDevice(s->device));
            CU(cudaMemcpy2DAsync(out + g->sub_t0[i], g->nT * sizeof(ppp_hit), s->d_hits, w * sizeof(ppp_hit), w * sizeof(ppp_hit), nQ,
                                 cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            return (int)PPP_OK;
        });
    }
    const size_t k = g->group_k;
    std::vector<size_t> prefix(nQ + 1, 0);
    for (size_t q = 0; q < nQ; ++q) prefix[q + 1] = prefix[q] + g->topk_counts[q];
    return for_each_sub(g, [&](ppp_ctx *s, size_t d) {
        ppp_ctx *ctx = s;
        const size_t q0 = nQ * d / nd, q1 = nQ * (d + 1) / nd, ns = q1 - q0;
        if (!ns) return (int)PPP_OK;
        CU(cudaSetDevice(s->device));
        if (prefix[q1] - prefix[q0] == ns * k) { /* every list of the slice is full: straight into the caller's buffer */
            CU(cudaMemcpyAsync(out + prefix[q0], s->d_mtopk, ns * k * sizeof(ppp_hit), cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
        } else {
            std::vector<ppp_hit> lists(ns * k);
            CU(cudaMemcpyAsync(lists.data(), s->d_mtopk, lists.size() * sizeof(ppp_hit), cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            for (size_t q = q0; q < q1; ++q) memcpy(out + prefix[q], lists.data() + (q - q0) * k, g->topk_counts[q] * sizeof(ppp_hit));
        }
        return (int)PPP_OK;
    });
}

static int group_set_option(ppp_ctx *g, const char *name, int64_t value)
{
    for (ppp_ctx *s : g->subs) {
        const int rc = ppp_set_option(s, name, value);
        if (rc) {
            snprintf(g->err, sizeof(g->err), "%s", s->err);
            return rc;
        }
    }
    return PPP_OK;
}

static int group_get_counter(const ppp_ctx *g, const char *name, int64_t *value)
{
    /* times ("us_*") and per-run settings = the maximum over the devices; work counters = the sum */
    const bool is_max = !strncmp(name, "us_", 3) || !strcmp(name, "num_sms") || !strcmp(name, "used_table") || !strcmp(name, "spec_rank");
    int64_t acc = 0;
    for (const ppp_ctx *s : g->subs) {
        int64_t v = 0;
        const int rc = ppp_get_counter(s, name, &v);
        if (rc) return rc;
        acc = is_max ? std::max(acc, v) : acc + v;
    }
    *value = acc;
    return PPP_OK;
}
