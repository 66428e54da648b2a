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
#include <new>
#include <vector>

#include "../../include/ppp_b200.h"
#include "ppp_common.cuh"

extern "C" cudaError_t ppp_launch_sweep(const SweepParams *prm, int wpl, int nwarps, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_exact(const SweepParams *prm, int W, int nsm, cudaStream_t st);
extern "C" cudaError_t ppp_launch_debug_bdtrc(const int32_t *k, const int32_t *n, const double *p, double *out, size_t m,
                                              cudaStream_t st);
extern "C" cudaError_t ppp_launch_filter(const void *hits, uint64_t npairs, double thresh_log10, int keep_all, void *out,
                                         uint64_t cap, unsigned long long *count, int nsm, cudaStream_t st);

struct ppp_ctx {
    int device = 0;
    int nsm = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    char err[512] = {0};

    /* targets */
    uint32_t *d_T = nullptr;
    size_t nT = 0;
    uint32_t G = 0;
    int wpl = 0; /* device words per lane: device row = 32*wpl words */
    double *d_lf = nullptr;
    /* queries */
    uint32_t *d_Q = nullptr;
    QConst *d_qc = nullptr;
    size_t nQ = 0, capQ = 0;
    /* results */
    ppp_hit *d_hits = nullptr;
    size_t cap_hits = 0;
    size_t n_results = 0;
    std::vector<ppp_hit> host_results; /* filtered modes: finalised on the host side of the ABI */
    bool results_on_host = false;
    uint32_t *d_exact = nullptr;
    size_t cap_exact = 0;
    unsigned long long *d_counters = nullptr;
    ppp_hit *d_filtered = nullptr;
    size_t cap_filtered = 0;

    /* options */
    int force_exact = 0, check_bounds = 0, sweep_warps = 8;
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

extern "C" uint32_t ppp_words_per_row(uint32_t G) { return ((G + 127u) / 128u) * 4u; }
extern "C" int ppp_abi_version(void) { return PPP_B200_ABI_VERSION; }

extern "C" const char *ppp_last_error(const ppp_ctx *ctx) { return ctx ? ctx->err : g_init_err; }

extern "C" int ppp_init(ppp_ctx **out, const int *devices, int ndev)
{
    if (!out) return fail(nullptr, PPP_ERR_INVALID, "ppp_init: out is NULL");
    *out = nullptr;
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
    for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
    if ((e = cudaMalloc(&ctx->d_counters, CNT_N * sizeof(unsigned long long))) != cudaSuccess) {
        fail(nullptr, PPP_ERR_CUDA, "ppp_init: %s", cudaGetErrorString(e));
        delete ctx;
        return PPP_ERR_CUDA;
    }
    *out = ctx;
    return PPP_OK;
}

extern "C" void ppp_destroy(ppp_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_T);
    cudaFree(ctx->d_lf);
    cudaFree(ctx->d_Q);
    cudaFree(ctx->d_qc);
    cudaFree(ctx->d_hits);
    cudaFree(ctx->d_exact);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_filtered);
    for (int i = 0; i < 4; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

static int pick_wpl(uint32_t G)
{
    if (G <= 1024) return 1;
    if (G <= 2048) return 2;
    if (G <= 4096) return 4;
    if (G <= 8192) return 8;
    return 0;
}

extern "C" int ppp_set_targets_bits(ppp_ctx *ctx, const uint32_t *T, size_t nT, uint32_t G)
{
    if (!ctx) return PPP_ERR_INVALID;
    if (!T && nT) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits: T is NULL");
    if (G == 0 || G > 8192) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits: G = %u outside [1, 8192]", G);
    if (nT > 0xffffffffull) return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_bits: nT too large");
    CU(cudaSetDevice(ctx->device));
    const int wpl = pick_wpl(G);
    const size_t W = 32 * (size_t)wpl, wpr = ppp_words_per_row(G);
    cudaFree(ctx->d_T);
    ctx->d_T = nullptr;
    cudaFree(ctx->d_lf);
    ctx->d_lf = nullptr;
    ctx->nT = 0;
    if (nT) {
        CU(cudaMalloc(&ctx->d_T, nT * W * 4));
        CU(cudaMemsetAsync(ctx->d_T, 0, nT * W * 4, ctx->stream));
        CU(cudaMemcpy2DAsync(ctx->d_T, W * 4, T, wpr * 4, wpr * 4, nT, cudaMemcpyHostToDevice, ctx->stream));
    }
    /* log-factorial table lf[k] = lgamma(k+1), k = 0..Gd+1, computed in long double */
    const size_t Gd = 32 * W;
    std::vector<double> lf(Gd + 2);
    for (size_t k = 0; k < Gd + 2; ++k) lf[k] = (double)lgammal((long double)k + 1.0L);
    CU(cudaMalloc(&ctx->d_lf, (Gd + 2) * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->d_lf, lf.data(), (Gd + 2) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->nT = nT;
    ctx->G = G;
    ctx->wpl = wpl;
    ctx->nQ = 0; /* queries must be re-uploaded: the device row length may have changed */
    cudaFree(ctx->d_Q);
    cudaFree(ctx->d_qc);
    ctx->d_Q = nullptr;
    ctx->d_qc = nullptr;
    ctx->capQ = 0;
    return PPP_OK;
}

extern "C" int ppp_set_targets_ranked(ppp_ctx *ctx, const uint16_t *idx, const uint16_t *rawdepth, const uint64_t *row_off,
                                      size_t nT, uint32_t G)
{
    (void)idx; (void)rawdepth; (void)row_off; (void)nT; (void)G;
    if (!ctx) return PPP_ERR_INVALID;
    return fail(ctx, PPP_ERR_INVALID, "ppp_set_targets_ranked: ranked target lists are not built in this round (DESIGN.md, 'next')");
}

static void make_qconst(double p, uint32_t ny, QConst *qc)
{
    memset(qc, 0, sizeof(*qc));
    qc->p = p;
    qc->ny = ny;
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
    if (!ctx) return PPP_ERR_INVALID;
    if (!ctx->wpl) return fail(ctx, PPP_ERR_STATE, "ppp_set_queries: call ppp_set_targets_* first (it fixes G)");
    if (nQ && (!P || !Y || !p)) return fail(ctx, PPP_ERR_INVALID, "ppp_set_queries: NULL input");
    if (nQ > 0xffffffffull) return fail(ctx, PPP_ERR_INVALID, "ppp_set_queries: nQ too large");
    CU(cudaSetDevice(ctx->device));
    const size_t W = 32 * (size_t)ctx->wpl, wpr = ppp_words_per_row(ctx->G);
    if (nQ > ctx->capQ) {
        cudaFree(ctx->d_Q);
        cudaFree(ctx->d_qc);
        ctx->d_Q = nullptr;
        ctx->d_qc = nullptr;
        ctx->capQ = 0;
        /* round up to a whole tile so that a tile's TMA bulk copy never reads past the allocation */
        const size_t cap = ((nQ + PPP_QTILE - 1) / PPP_QTILE) * PPP_QTILE;
        CU(cudaMalloc(&ctx->d_Q, cap * 2 * W * 4));
        CU(cudaMalloc(&ctx->d_qc, cap * sizeof(QConst)));
        ctx->capQ = cap;
    }
    ctx->nQ = 0;
    if (!nQ) return PPP_OK;
    std::vector<QConst> qc(nQ);
    for (size_t q = 0; q < nQ; ++q) {
        uint32_t ny = 0;
        const uint32_t *pr = P + q * wpr, *yr = Y + q * wpr;
        for (size_t w = 0; w < wpr; ++w) ny += (uint32_t)__builtin_popcount(pr[w] & yr[w]);
        make_qconst(p[q], ny, &qc[q]);
    }
    CU(cudaMemsetAsync(ctx->d_Q, 0, ctx->capQ * 2 * W * 4, ctx->stream));
    CU(cudaMemcpy2DAsync(ctx->d_Q, 2 * W * 4, P, wpr * 4, wpr * 4, nQ, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpy2DAsync(ctx->d_Q + W, 2 * W * 4, Y, wpr * 4, wpr * 4, nQ, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_qc, qc.data(), nQ * sizeof(QConst), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); /* inputs are copied before returning (Perl SV buffers may move) */
    ctx->nQ = nQ;
    return PPP_OK;
}

static int ensure_buffers(ppp_ctx *ctx, size_t dense_pairs, size_t launch_pairs)
{
    if (dense_pairs > ctx->cap_hits) {
        cudaFree(ctx->d_hits);
        ctx->d_hits = nullptr;
        ctx->cap_hits = 0;
        CU(cudaMalloc(&ctx->d_hits, dense_pairs * sizeof(ppp_hit)));
        ctx->cap_hits = dense_pairs;
    }
    if (launch_pairs > ctx->cap_exact) {
        cudaFree(ctx->d_exact);
        ctx->d_exact = nullptr;
        ctx->cap_exact = 0;
        CU(cudaMalloc(&ctx->d_exact, launch_pairs * sizeof(uint32_t)));
        ctx->cap_exact = launch_pairs;
    }
    return PPP_OK;
}

static bool hit_less(const ppp_hit &a, const ppp_hit &b)
{
    if (a.q != b.q) return a.q < b.q;
    if (a.score != b.score) return a.score < b.score;
    return a.t < b.t;
}

extern "C" int ppp_run(ppp_ctx *ctx, const ppp_filter *f)
{
    if (!ctx) return PPP_ERR_INVALID;
    if (!ctx->wpl || !ctx->nT) return fail(ctx, PPP_ERR_STATE, "ppp_run: no targets resident");
    if (!ctx->nQ) return fail(ctx, PPP_ERR_STATE, "ppp_run: no queries resident");
    ppp_filter all = {PPP_FILTER_ALL, 0, 0.0};
    if (!f) f = &all;
    if (f->mode > PPP_FILTER_TOPK) return fail(ctx, PPP_ERR_INVALID, "ppp_run: unknown filter mode %u", f->mode);
    if (f->mode == PPP_FILTER_TOPK && f->k == 0) return fail(ctx, PPP_ERR_INVALID, "ppp_run: top-k with k = 0");
    CU(cudaSetDevice(ctx->device));
    const size_t nQ = ctx->nQ, nT = ctx->nT;
    const int W = 32 * ctx->wpl;
    /* targets per launch: bounded so that launch-local pair ids fit 31 bits; filtered modes also bound the dense
     * scratch buffer (the dense records of a chunk are filtered on the device before the next chunk overwrites them) */
    size_t chunk_pairs_max = (size_t)1 << 30;
    if (f->mode != PPP_FILTER_ALL) chunk_pairs_max = (size_t)1 << 26;
    size_t tchunk = std::max<size_t>(1, chunk_pairs_max / nQ);
    if (tchunk > nT) tchunk = nT;
    const size_t dense_pairs = (f->mode == PPP_FILTER_ALL) ? nQ * nT : nQ * tchunk;
    int rc = ensure_buffers(ctx, dense_pairs, nQ * tchunk);
    if (rc) return rc;

    ctx->host_results.clear();
    ctx->results_on_host = (f->mode != PPP_FILTER_ALL);
    unsigned long long totals[CNT_N];
    memset(totals, 0, sizeof(totals));
    float sweep_ms = 0.f;
    uint32_t sweep_launches = 0, launches = 0;
    std::vector<std::vector<ppp_hit>> topk; /* per query running best-k (host side of the ABI) */
    if (f->mode == PPP_FILTER_TOPK) topk.resize(nQ);

    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    for (size_t t0 = 0; t0 < nT; t0 += tchunk) {
        const size_t t1 = std::min(nT, t0 + tchunk);
        SweepParams prm;
        memset(&prm, 0, sizeof(prm));
        prm.T = ctx->d_T;
        prm.Q = ctx->d_Q;
        prm.qc = ctx->d_qc;
        prm.lf = ctx->d_lf;
        prm.exact_list = ctx->d_exact;
        prm.counters = ctx->d_counters;
        prm.exact_cap = ctx->cap_exact;
        prm.nQ = (uint32_t)nQ;
        prm.G = ctx->G;
        prm.force_exact = (uint32_t)ctx->force_exact;
        prm.check_bounds = (uint32_t)ctx->check_bounds;
        prm.t0 = (uint32_t)t0;
        prm.t1 = (uint32_t)t1;
        if (f->mode == PPP_FILTER_ALL) {
            prm.hits = ctx->d_hits;
            prm.nT = (uint32_t)nT;
        } else {
            /* chunk-local dense scratch: row stride = chunk width, rows start at column t0 */
            prm.hits = ctx->d_hits - t0; /* kernels index q * nT + t with nT = chunk width below */
            prm.nT = (uint32_t)(t1 - t0);
        }
        CU(cudaMemsetAsync(ctx->d_counters, 0, CNT_N * sizeof(unsigned long long), ctx->stream));
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        CU(ppp_launch_sweep(&prm, ctx->wpl, ctx->sweep_warps, ctx->nsm, ctx->stream));
        CU(cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(ppp_launch_exact(&prm, W, ctx->nsm, ctx->stream));
        launches += 2;
        sweep_launches += 1;
        unsigned long long c[CNT_N];
        CU(cudaMemcpyAsync(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
        if (f->mode != PPP_FILTER_ALL) {
            /* device-side threshold compaction of the chunk's dense records */
            const size_t np = nQ * (t1 - t0);
            if (np > ctx->cap_filtered) {
                cudaFree(ctx->d_filtered);
                ctx->d_filtered = nullptr;
                ctx->cap_filtered = 0;
                CU(cudaMalloc(&ctx->d_filtered, np * sizeof(ppp_hit)));
                ctx->cap_filtered = np;
            }
            CU(cudaMemsetAsync(ctx->d_counters + CNT_N - 1, 0, sizeof(unsigned long long), ctx->stream));
            CU(ppp_launch_filter(ctx->d_hits, np, f->thresh, f->mode == PPP_FILTER_TOPK, ctx->d_filtered, ctx->cap_filtered,
                                 ctx->d_counters + CNT_N - 1, ctx->nsm, ctx->stream));
            launches += 1;
            unsigned long long nf = 0;
            CU(cudaMemcpyAsync(&nf, ctx->d_counters + CNT_N - 1, sizeof(nf), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            std::vector<ppp_hit> part(nf);
            if (nf) CU(cudaMemcpy(part.data(), ctx->d_filtered, nf * sizeof(ppp_hit), cudaMemcpyDeviceToHost));
            if (f->mode == PPP_FILTER_THRESH_LOG10) {
                ctx->host_results.insert(ctx->host_results.end(), part.begin(), part.end());
            } else {
                for (const ppp_hit &h : part) topk[h.q].push_back(h);
                for (size_t q = 0; q < nQ; ++q) {
                    std::vector<ppp_hit> &v = topk[q];
                    if (v.size() > f->k) {
                        std::partial_sort(v.begin(), v.begin() + f->k, v.end(), hit_less);
                        v.resize(f->k);
                    }
                }
            }
        } else {
            CU(cudaStreamSynchronize(ctx->stream));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
        sweep_ms += ms;
        for (int i = 0; i < CNT_N; ++i) totals[i] += c[i];
        if (c[CNT_OVERFLOW])
            return fail(ctx, PPP_ERR_CUDA, "ppp_run: exact-tier list overflow (%llu pairs dropped)", c[CNT_OVERFLOW]);
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float total_ms = 0.f;
    cudaEventElapsedTime(&total_ms, ctx->ev[0], ctx->ev[1]);

    if (f->mode == PPP_FILTER_TOPK) {
        for (size_t q = 0; q < nQ; ++q) {
            std::sort(topk[q].begin(), topk[q].end(), hit_less);
            ctx->host_results.insert(ctx->host_results.end(), topk[q].begin(), topk[q].end());
        }
    } else if (f->mode == PPP_FILTER_THRESH_LOG10) {
        std::sort(ctx->host_results.begin(), ctx->host_results.end(), hit_less);
    }
    ctx->n_results = (f->mode == PPP_FILTER_ALL) ? nQ * nT : ctx->host_results.size();
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
    if (!ctx) return PPP_ERR_INVALID;
    if (nout) *nout = ctx->n_results;
    if (cap < ctx->n_results) return fail(ctx, PPP_ERR_CAPACITY, "ppp_fetch: need room for %zu hits, got %zu", ctx->n_results, cap);
    if (!ctx->n_results) return PPP_OK;
    if (!out) return fail(ctx, PPP_ERR_INVALID, "ppp_fetch: out is NULL");
    CU(cudaSetDevice(ctx->device));
    if (ctx->results_on_host) {
        memcpy(out, ctx->host_results.data(), ctx->n_results * sizeof(ppp_hit));
    } else {
        CU(cudaMemcpyAsync(out, ctx->d_hits, ctx->n_results * sizeof(ppp_hit), cudaMemcpyDeviceToHost, ctx->stream));
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

extern "C" int ppp_get_stats(const ppp_ctx *ctx, ppp_stats *out)
{
    if (!ctx || !out) return PPP_ERR_INVALID;
    *out = ctx->stats;
    return PPP_OK;
}

extern "C" int ppp_set_option(ppp_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return PPP_ERR_INVALID;
    if (!strcmp(name, "force_exact")) ctx->force_exact = value != 0;
    else if (!strcmp(name, "check_bounds")) ctx->check_bounds = value != 0;
    else if (!strcmp(name, "sweep_warps")) {
        if (value < 1 || value > 16) return fail(ctx, PPP_ERR_INVALID, "sweep_warps must be in [1,16]");
        ctx->sweep_warps = (int)value;
    } else
        return fail(ctx, PPP_ERR_INVALID, "unknown option '%s'", name);
    return PPP_OK;
}

extern "C" int ppp_get_counter(const ppp_ctx *ctx, const char *name, int64_t *value)
{
    if (!ctx || !name || !value) return PPP_ERR_INVALID;
    int i = -1;
    if (!strcmp(name, "exact_pairs")) i = CNT_EXACT;
    else if (!strcmp(name, "candidates")) i = CNT_CAND;
    else if (!strcmp(name, "accurate_evals")) i = CNT_ACCURATE;
    else if (!strcmp(name, "bound_violations")) i = CNT_VIOL;
    else if (!strcmp(name, "bound_checks")) i = CNT_CHECKS;
    else if (!strcmp(name, "num_sms")) { *value = ctx->nsm; return PPP_OK; }
    if (i < 0) return PPP_ERR_INVALID;
    *value = (int64_t)ctx->last_counters[i];
    return PPP_OK;
}

extern "C" int ppp_debug_bdtrc(ppp_ctx *ctx, const int32_t *k, const int32_t *n, const double *p, double *out, size_t m)
{
    if (!ctx || !k || !n || !p || !out) return PPP_ERR_INVALID;
    if (!m) return PPP_OK;
    CU(cudaSetDevice(ctx->device));
    int32_t *dk = nullptr, *dn = nullptr;
    double *dp = nullptr, *dout = nullptr;
    CU(cudaMalloc(&dk, m * 4));
    CU(cudaMalloc(&dn, m * 4));
    CU(cudaMalloc(&dp, m * 8));
    CU(cudaMalloc(&dout, m * 8));
    CU(cudaMemcpyAsync(dk, k, m * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dn, n, m * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dp, p, m * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(ppp_launch_debug_bdtrc(dk, dn, dp, dout, m, ctx->stream));
    CU(cudaMemcpyAsync(out, dout, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(dk);
    cudaFree(dn);
    cudaFree(dp);
    cudaFree(dout);
    return PPP_OK;
}
