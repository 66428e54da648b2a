/*
 * prophylo_b200/csrc/ppp_topk.cu -- device side of the result filters:
 *
 *  ppp_build_staircase_kernel   per query (one CTA of 2, 4 or 8 warps), the exact integer pre-filter of a score threshold tau (ln domain):
 *        nmax[yes] = largest `number` for which P(X >= yes | number, p) can still be <= exp(tau).
 *     The binomial tail increases with number and decreases with yes, so a sweep candidate (yes, number) can
 *     reach the threshold iff number <= nmax[yes]; the sweep kernel rejects a whole lane of words with one
 *     lookup.  tau is the query's current k-th best score (top-k, bin/ppp.pl:473-506 keeps the best rows) or the
 *     fixed ppp_cutoff.pl:122 threshold.  Built with the FP64 ratio series and a safety margin, so it never
 *     rejects a pair the reference would keep.
 *
 *  ppp_topk_merge_kernel        per query, merge the hits appended by the last target chunk into the running
 *     top-k list (ascending score, ties by target index) with a shared-memory bitonic sort, publish the new
 *     k-th score and reset the append counter.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ppp_b200.h"
#include "ppp_common.cuh"

/* Can ln P(X >= y | n, p) be <= tau (tau < ln 1/2)?
 *   - lower side (y + 1 <= (n+1) p): no -- the binomial median is >= floor(n p), so the tail is >= 1/2;
 *   - otherwise the sweep kernel's own float enclosure (tier_a) decides: its lower bound `lo` never exceeds the
 *     true value, so "lo <= tau" holds wherever the true tail is <= tau.
 * The answer may err on the side of "yes" only (by the width of the enclosure, a few percent of a nat), which makes
 * the staircase slightly looser, never wrong.  The predicate need not be monotone beyond the true boundary: a
 * bisection that only moves its upper end past a FALSE probe still ends at or above the true boundary. */
__device__ static bool tail_may_be_le(int y, int n, const QConst &qc, const double *__restrict__ lf, double tau)
{
    const int m = n - y;
    if (!((double)m * qc.theta < (double)(y + 1))) return false;
    float lo, hi;
    tier_a(y, n, qc, lf, lo, hi);
    return !(lo > (float)tau + 1e-3f + 1e-6f * fabsf((float)tau)); /* NaN-safe; margin covers the float cast of tau */
}

/* One staircase: tab[y] (y = 0..ny) = largest number for which P(X >= y | number, p) can still be <= exp(tau).  A TEAM of
 * warps (one CTA) works on it: thread i searches y = 1 + i, 1 + i + NT, ... (NT = 32 * TEAM), then warp 0 applies the monotone
 * fix-up.  Requires tau < -0.70 and an ordinary query.
 *
 * The search for one y is a bisection on `number` over [y, |P|] with the predicate tail_may_be_le (13 enclosure evaluations at
 * G = 4096).  From its third value on a thread extrapolates its two previous results linearly (the staircase is smooth: its
 * second difference over NT entries is a few units), probes there and gallops -- steps of 2, 4, 8 ... -- to a bracket before
 * bisecting: about 5 evaluations.  Only FALSE probes ever lower the upper end, so the result stays at or above the true boundary
 * whatever the predicate does beyond it (see tail_may_be_le). */
template <int TEAM>
__device__ static void build_staircase(const QConst &c, const double *__restrict__ lf, double tau, uint16_t *__restrict__ tab, int tid)
{
    constexpr int NT = 32 * TEAM;
    const int lane = tid & 31;
    const double tau_m = tau + 1e-5 + 1e-9 * fabs(tau); /* safety margin >> series + table error */
    const int ny = (int)c.ny, np_ = (int)c.npres;
    if (tid == 0) tab[0] = 0;
    int p1 = 0, p2 = 0; /* this thread's two previous results (0: none, or no valid entry) */
    for (int y = 1 + tid; y <= ny; y += NT) {
        int lo = -1, hi = np_; /* lo: a number known to pass (-1: none yet); every number above hi is known to fail or out of range */
        if (isfinite(tau_m)) {
            if (p1 > 0 && p2 > 0) {
                const int g = min(max(p1 + (p1 - p2), y), np_);
                int d = 2;
                if (tail_may_be_le(y, g, c, lf, tau_m)) {
                    lo = g;
                    while (lo < hi) { /* gallop up */
                        const int t = min(lo + d, hi);
                        if (tail_may_be_le(y, t, c, lf, tau_m)) {
                            lo = t;
                            d <<= 1;
                        } else {
                            hi = t - 1;
                            break;
                        }
                    }
                } else {
                    hi = g - 1;
                    while (hi >= y) { /* gallop down */
                        const int t = max(hi - d + 1, y);
                        if (tail_may_be_le(y, t, c, lf, tau_m)) {
                            lo = t;
                            break;
                        }
                        hi = t - 1;
                        d <<= 1;
                    }
                }
            } else if (tail_may_be_le(y, y, c, lf, tau_m)) {
                lo = y;
            }
        }
        int res = 0;
        if (lo >= 0) {
            while (lo < hi) { /* invariant: lo passes */
                const int mid = (lo + hi + 1) >> 1;
                if (tail_may_be_le(y, mid, c, lf, tau_m)) lo = mid;
                else hi = mid - 1;
            }
            res = lo;
        }
        tab[y] = (uint16_t)res;
        p2 = p1;
        p1 = res;
    }
    __syncthreads();
    /* enforce what the true staircase satisfies (independent searches may differ by rounding; only ever loosens):
     * non-decreasing in yes, and nmax[y] >= y  =>  nmax[y+1] >= nmax[y] + 1 (one more trial that is a success cannot
     * raise the tail), i.e. the slack nmax[y] - y is non-decreasing from the first valid entry on -- the chunk-level
     * screen of ppp_prefilter.cu tests only a chunk's last candidate on the strength of this.  Two warp prefix-max
     * scans, 32 entries a round: v1 = running max of the table, then slack = running max of (v1[i] - i) over the
     * valid entries (v1[i] >= i); result = y + slack where a valid entry precedes, v1 otherwise. */
    if (tid < 32) {
        uint32_t carry_v = 0;
        int carry_s = -1;
        for (int y0 = 1; y0 <= ny; y0 += 32) {
            const int y = y0 + lane;
            const bool in = y <= ny;
            uint32_t v = in ? (uint32_t)tab[y] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v = max(v, u);
            }
            v = max(v, carry_v);
            carry_v = __shfl_sync(0xffffffffu, v, 31);
            int sl = (in && (int)v >= y) ? (int)v - y : -1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, sl, o);
                if (lane >= o) sl = max(sl, u);
            }
            sl = max(sl, carry_s);
            carry_s = __shfl_sync(0xffffffffu, sl, 31);
            if (in) tab[y] = (uint16_t)min((sl >= 0) ? (uint32_t)(y + sl) : v, 65535u);
        }
    }
}

/* one CTA of TEAM warps per query */
template <int TEAM>
__global__ void __launch_bounds__(32 * TEAM) ppp_build_staircase_kernel(const QConst *__restrict__ qc, const double *__restrict__ lf,
                                                                        const double *__restrict__ kth, int accept_mode, double thresh_ln,
                                                                        const uint32_t *__restrict__ static_off, uint16_t *__restrict__ nmax,
                                                                        uint32_t *__restrict__ active_off, double *__restrict__ last_tau, uint32_t nQ)
{
    const int tid = threadIdx.x;
    for (uint32_t q = blockIdx.x; q < nQ; q += gridDim.x) {
        const QConst c = qc[q];
        double tau;
        if (accept_mode == 0) {
            const double kq = kth[q];
            tau = (kq < 1.0) ? ((kq > 0.0) ? log(kq) : -INFINITY) : INFINITY;
        } else {
            tau = thresh_ln;
        }
        /* no pruning near or above the median, for odd p, or for empty queries */
        const bool none = !(tau < -0.70) || (c.flags & (PPP_QF_TRIVIAL | PPP_QF_DOMAIN | PPP_QF_EXACT)) || c.ny == 0;
        /* the k-th score only ever decreases: a staircase built for a slightly larger tau stays valid (just looser) */
        const bool keep = !none && active_off[q] != PPP_NO_TABLE && tau > last_tau[q] - 0.05;
        __syncthreads(); /* every thread has read the query's state before thread 0 changes it */
        if (none) {
            if (tid == 0) active_off[q] = PPP_NO_TABLE;
            continue;
        }
        if (keep) continue;
        build_staircase<TEAM>(c, lf, tau, nmax + static_off[q], tid);
        if (tid == 0) {
            last_tau[q] = tau;
            active_off[q] = static_off[q];
        }
    }
}

extern "C" cudaError_t ppp_launch_staircase(const QConst *qc, const double *lf, const double *kth, int accept_mode, double thresh_ln,
                                            const uint32_t *static_off, uint16_t *nmax, uint32_t *active_off, double *last_tau,
                                            uint32_t nQ, int nsm, cudaStream_t st)
{
    /* The kernel is latency-bound (dependent enclosure evaluations), so what matters is that every warp slot of the chip has
     * work: the fewer queries, the more warps share one (the query slice of one GPU of an 8-GPU job has 1,250 queries for
     * 9,472 warp slots).  Two warps per query at least: 32 CTAs per SM of 64 threads fill an SM's 64 warp slots. */
    const uint32_t slots = (uint32_t)nsm * 64u;
    if (!nQ) return cudaSuccess;
    if ((uint64_t)nQ * 8u <= slots) {
        ppp_build_staircase_kernel<8><<<nQ, 256, 0, st>>>(qc, lf, kth, accept_mode, thresh_ln, static_off, nmax, active_off, last_tau, nQ);
    } else if ((uint64_t)nQ * 4u <= slots) {
        ppp_build_staircase_kernel<4><<<nQ, 128, 0, st>>>(qc, lf, kth, accept_mode, thresh_ln, static_off, nmax, active_off, last_tau, nQ);
    } else {
        const uint32_t grid = nQ < slots ? nQ : slots; /* grid-stride over the queries beyond the resident CTAs */
        ppp_build_staircase_kernel<2><<<grid, 64, 0, st>>>(qc, lf, kth, accept_mode, thresh_ln, static_off, nmax, active_off, last_tau, nQ);
    }
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------ top-k merge */
#define TOPK_M 4096 /* records sorted at once (128 KB of shared memory); k <= TOPK_M / 2 */

__device__ __forceinline__ bool rec_less(const ppp_hit &a, const ppp_hit &b)
{
    if (a.score != b.score) return a.score < b.score;
    return a.t < b.t;
}

/* sort keys only (16 bytes: score, target, position in the concatenated input) and gather the k winners' 32-byte records
 * afterwards: half the shared-memory traffic of sorting records, and 64 KB instead of 128 KB per block */
struct __align__(16) MKey {
    double score;
    uint32_t t, idx;
};
__device__ __forceinline__ bool key_less(const MKey &a, const MKey &b)
{
    if (a.score != b.score) return a.score < b.score;
    return a.t < b.t;
}

/* M = keys sorted at once (a power of two >= 2 k, the kernel's shared memory), PER = winners per thread (256 PER >= k).  Short
 * lists get a small M and PER = 1: the sort of a few hundred keys is a chain of barrier-separated stages, i.e. latency, so what
 * counts is how many queries are in flight -- 16 KB and 40 registers let 8 CTAs share an SM where 64 KB and the 8 winners'
 * records in registers allowed 2. */
template <int PER>
__global__ void __launch_bounds__(256) ppp_topk_merge_kernel(const ppp_hit *__restrict__ appended, uint32_t *__restrict__ qcnt,
                                                             uint32_t qcap, ppp_hit *__restrict__ topk, uint32_t *__restrict__ topk_cnt,
                                                             double *__restrict__ kth, uint32_t k, uint32_t nQ, uint32_t t_offset, uint32_t M)
{
    extern __shared__ __align__(16) unsigned char smem_topk[];
    MKey *key = reinterpret_cast<MKey *>(smem_topk);
    for (uint32_t q = blockIdx.x; q < nQ; q += gridDim.x) {
        const uint32_t nnew = min(qcnt[q], qcap);
        uint32_t cur = topk_cnt[q];
        if (nnew == 0) continue; /* block-uniform */
        ppp_hit *list = topk + (size_t)q * k;
        const ppp_hit *app = appended + (size_t)q * qcap;
        uint32_t done = 0;
        double kth_new = 0.0;
        while (done < nnew) {
            const uint32_t batch = min(nnew - done, M - cur);
            for (uint32_t i = threadIdx.x; i < cur; i += blockDim.x) {
                MKey m;
                m.score = list[i].score;
                m.t = list[i].t;
                m.idx = i;
                key[i] = m;
            }
            for (uint32_t i = threadIdx.x; i < batch; i += blockDim.x) {
                MKey m;
                m.score = app[done + i].score;
                m.t = app[done + i].t + t_offset; /* shard-local -> global target index (multi-GPU merge); 0 otherwise */
                m.idx = cur + i;
                key[cur + i] = m;
            }
            const uint32_t tot = cur + batch;
            uint32_t n2 = 1;
            while (n2 < tot) n2 <<= 1;
            for (uint32_t i = tot + threadIdx.x; i < n2; i += blockDim.x) {
                MKey m;
                m.score = INFINITY;
                m.t = 0xffffffffu;
                m.idx = 0xffffffffu;
                key[i] = m;
            }
            __syncthreads();
            for (uint32_t size = 2; size <= n2; size <<= 1) {
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    for (uint32_t i = threadIdx.x; i < (n2 >> 1); i += blockDim.x) {
                        const uint32_t lo = 2 * i - (i & (stride - 1));
                        const uint32_t hi = lo + stride;
                        const bool asc = ((lo & size) == 0);
                        const MKey a = key[lo], b = key[hi];
                        if (key_less(b, a) == asc) {
                            key[lo] = b;
                            key[hi] = a;
                        }
                    }
                    __syncthreads();
                }
            }
            /* gather the winners (the list is rewritten in place: read everything first) */
            const uint32_t keep = min(tot, k);
            ppp_hit h[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const uint32_t i = threadIdx.x + 256u * j;
                if (i < keep) {
                    const uint32_t src = key[i].idx;
                    if (src < cur) {
                        h[j] = list[src];
                    } else {
                        h[j] = app[done + (src - cur)];
                        h[j].t += t_offset;
                    }
                }
            }
            if (keep >= k) kth_new = key[k - 1].score;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const uint32_t i = threadIdx.x + 256u * j;
                if (i < keep) list[i] = h[j];
            }
            cur = keep;
            done += batch;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            topk_cnt[q] = cur;
            if (cur >= k) kth[q] = kth_new; /* otherwise keep the previous (seeded) bound; 2.0 initially */
            qcnt[q] = 0;
        }
        __syncthreads();
    }
}

extern "C" cudaError_t ppp_launch_topk_merge(const void *appended, uint32_t *qcnt, uint32_t qcap, void *topk, uint32_t *topk_cnt,
                                             double *kth, uint32_t k, uint32_t nQ, uint32_t t_offset, int nsm, cudaStream_t st)
{
    const ppp_hit *app = reinterpret_cast<const ppp_hit *>(appended);
    ppp_hit *lists = reinterpret_cast<ppp_hit *>(topk);
    cudaError_t e;
    if (k <= 256) {
        const uint32_t M = 1024;
        e = cudaFuncSetAttribute(ppp_topk_merge_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_M * (int)sizeof(MKey));
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)min((uint32_t)(nsm * 6), nQ); /* 38 registers x 256 threads: 6 CTAs per SM */
        ppp_topk_merge_kernel<1><<<grid, 256, M * sizeof(MKey), st>>>(app, qcnt, qcap, lists, topk_cnt, kth, k, nQ, t_offset, M);
    } else {
        const uint32_t M = TOPK_M;
        e = cudaFuncSetAttribute(ppp_topk_merge_kernel<TOPK_M / 2 / 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_M * (int)sizeof(MKey));
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)min((uint32_t)(nsm * 2), nQ);
        ppp_topk_merge_kernel<TOPK_M / 2 / 256><<<grid, 256, M * sizeof(MKey), st>>>(app, qcnt, qcap, lists, topk_cnt, kth, k, nQ, t_offset, M);
    }
    return cudaGetLastError();
}

extern "C" int ppp_topk_max_k(void) { return TOPK_M / 2; }

/* multi-GPU exchange step: merge up to 16 gathered per-rank lists of one query in ONE sort (nlists * k <= TOPK_M) */
struct MergeOffsets {
    uint32_t t[16];
};

__global__ void __launch_bounds__(256) ppp_topk_merge_lists_kernel(const ppp_hit *__restrict__ lists, const uint32_t *__restrict__ counts,
                                                                   uint32_t nlists, MergeOffsets offs, uint32_t nQ, uint32_t k,
                                                                   ppp_hit *__restrict__ out, uint32_t *__restrict__ out_cnt)
{
    extern __shared__ __align__(16) unsigned char smem_topk[];
    MKey *key = reinterpret_cast<MKey *>(smem_topk); /* 16-byte keys are sorted, the winners' records gathered afterwards */
    __shared__ uint32_t s_base[17];
    for (uint32_t q = blockIdx.x; q < nQ; q += gridDim.x) {
        if (threadIdx.x == 0) {
            uint32_t b = 0;
            for (uint32_t l = 0; l < nlists; ++l) {
                s_base[l] = b;
                b += min(counts[(size_t)l * nQ + q], k);
            }
            s_base[nlists] = b;
        }
        __syncthreads();
        const uint32_t tot = s_base[nlists];
        for (uint32_t l = 0; l < nlists; ++l) {
            const uint32_t c = s_base[l + 1] - s_base[l];
            const ppp_hit *src = lists + ((size_t)l * nQ + q) * k;
            for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
                MKey m;
                m.score = src[i].score;
                m.t = src[i].t + offs.t[l];
                m.idx = (l << 16) | i; /* k <= 2048 */
                key[s_base[l] + i] = m;
            }
        }
        uint32_t n2 = 1;
        while (n2 < tot) n2 <<= 1;
        for (uint32_t i = tot + threadIdx.x; i < n2; i += blockDim.x) {
            MKey m;
            m.score = INFINITY;
            m.t = 0xffffffffu;
            m.idx = 0xffffffffu;
            key[i] = m;
        }
        __syncthreads();
        for (uint32_t size = 2; size <= n2; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t i = threadIdx.x; i < (n2 >> 1); i += blockDim.x) {
                    const uint32_t lo = 2 * i - (i & (stride - 1));
                    const uint32_t hi = lo + stride;
                    const bool asc = ((lo & size) == 0);
                    const MKey a = key[lo], b = key[hi];
                    if (key_less(b, a) == asc) {
                        key[lo] = b;
                        key[hi] = a;
                    }
                }
                __syncthreads();
            }
        }
        const uint32_t keep = min(tot, k);
        for (uint32_t i = threadIdx.x; i < keep; i += blockDim.x) {
            const MKey m = key[i];
            ppp_hit h = lists[((size_t)(m.idx >> 16) * nQ + q) * k + (m.idx & 0xffffu)];
            h.t = m.t;
            out[(size_t)q * k + i] = h;
        }
        if (threadIdx.x == 0) out_cnt[q] = keep;
        __syncthreads();
    }
}

/* returns cudaErrorNotSupported when the lists do not fit one sort (caller falls back to sequential merges) */
extern "C" cudaError_t ppp_launch_topk_merge_lists(const void *lists, const uint32_t *counts, uint32_t nlists, const uint32_t *t_offsets,
                                                   uint32_t nQ, uint32_t k, void *out, uint32_t *out_cnt, int nsm, cudaStream_t st)
{
    if (nlists > 16 || (size_t)nlists * k > TOPK_M) return cudaErrorNotSupported;
    MergeOffsets offs;
    for (uint32_t l = 0; l < 16; ++l) offs.t[l] = l < nlists ? t_offsets[l] : 0u;
    const int smem = TOPK_M * (int)sizeof(MKey);
    cudaError_t e = cudaFuncSetAttribute(ppp_topk_merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)min((uint32_t)(nsm * 3), nQ);
    ppp_topk_merge_lists_kernel<<<grid, 256, smem, st>>>(reinterpret_cast<const ppp_hit *>(lists), counts, nlists, offs, nQ, k,
                                                         reinterpret_cast<ppp_hit *>(out), out_cnt);
    return cudaGetLastError();
}

/* shift the target index of n records (a sub-context of a multi-device group reports global target numbers) */
__global__ void ppp_add_t_kernel(ppp_hit *__restrict__ hits, size_t n, uint32_t off)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) hits[i].t += off;
}

extern "C" cudaError_t ppp_launch_add_t(void *hits, size_t n, uint32_t off, int nsm, cudaStream_t st)
{
    if (!n || !off) return cudaSuccess;
    ppp_add_t_kernel<<<nsm * 8, 256, 0, st>>>(reinterpret_cast<ppp_hit *>(hits), n, off);
    return cudaGetLastError();
}

__global__ void ppp_accumulate_kernel(unsigned long long *totals, const unsigned long long *counters, int n)
{
    if ((int)threadIdx.x < n) totals[threadIdx.x] += counters[threadIdx.x];
}

extern "C" cudaError_t ppp_launch_accumulate(unsigned long long *totals, const unsigned long long *counters, int n, cudaStream_t st)
{
    ppp_accumulate_kernel<<<1, 32, 0, st>>>(totals, counters, n);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------ speculative thresholds
 * After the first target chunks of a top-k run a query's list holds its k best pairs of `processed` targets.  Its final
 * k-th best score over all nT targets is, for exchangeable targets, close to the (k * processed / nT)-th best seen so
 * far -- far below the guaranteed bound (the k-th best so far).  The remaining targets are therefore screened against
 * the r-th best score so far (r chosen on the host so that a Poisson(k processed / nT) count reaches r with probability
 * <= spec_eps_ppm, 5e-3 by default), which cuts the pairs that need a full evaluation several-fold, and the result is VERIFIED: every pair at
 * or below the speculative threshold was found exactly (acceptance is INCLUSIVE in speculative launches, accept_mode 3), so the
 * list is exact iff it is full and its k-th score does not exceed that threshold.  Inclusive matters: the scores of a sparse
 * query are heavily tied (thousands of targets share the score of "one yes at the first position"), the threshold -- an entry
 * of the list -- is typically one of the tied values, and with a strict rule every such query failed the verification (5 % of
 * config 3's queries, whatever the Poisson tail) and was scanned twice.  Queries that fail are re-scanned over the same targets
 * with the guaranteed bound and klo < score <= kth acceptance (ppp_capi.cu). */
/* prev != nullptr: the targets processed so far were themselves screened against the speculative thresholds prev[q] (taken
 * from the seeding pass), so a list holds EVERY processed pair below prev[q] and its r-th entry is the exact r-th best of
 * the processed targets even when the list is not full; the new threshold never exceeds prev[q] and is always verified. */
__global__ void ppp_spec_init_kernel(const ppp_hit *__restrict__ topk, const uint32_t *__restrict__ topk_cnt, const double *__restrict__ kth,
                                     const double *__restrict__ prev, uint32_t k, uint32_t r, double *__restrict__ kspec,
                                     double *__restrict__ kspec0, uint32_t nQ)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nQ) return;
    const double g = kth[q];
    double use = g, s0 = INFINITY;
    if (topk_cnt[q] >= (prev ? r : k) && r >= 1 && r <= k) {
        const double v = topk[(size_t)q * k + (r - 1)].score;
        if (v > 0.0 && v < g) { /* exact zeros (Cephes underflow) leave nothing below them: keep the guaranteed bound */
            use = v;
            s0 = v;
        }
    }
    if (prev) {
        use = fmin(use, prev[q]);
        s0 = use;
    }
    kspec[q] = use;
    kspec0[q] = s0;
}

__global__ void ppp_spec_update_kernel(double *__restrict__ kspec, const double *__restrict__ kth, uint32_t nQ)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nQ) kspec[q] = fmin(kspec[q], kth[q]);
}

__global__ void ppp_spec_verify_kernel(const ppp_hit *__restrict__ topk, const uint32_t *__restrict__ topk_cnt, const double *__restrict__ kspec0,
                                       uint32_t k, double *__restrict__ last_tau, uint32_t *__restrict__ failed, uint32_t *__restrict__ nfailed,
                                       uint32_t nQ)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nQ) return;
    const double s0 = kspec0[q];
    if (!(s0 < INFINITY)) return; /* not speculated */
    const bool ok = topk_cnt[q] >= k && topk[(size_t)q * k + (k - 1)].score <= s0;
    if (!ok) {
        failed[atomicAdd(nfailed, 1u)] = q;
        last_tau[q] = INFINITY; /* force the staircase of this query to be rebuilt for the (looser) guaranteed bound */
    }
}

extern "C" cudaError_t ppp_launch_spec_init(const void *topk, const uint32_t *topk_cnt, const double *kth, const double *prev, uint32_t k,
                                            uint32_t r, double *kspec, double *kspec0, uint32_t nQ, cudaStream_t st)
{
    ppp_spec_init_kernel<<<(nQ + 255) / 256, 256, 0, st>>>(reinterpret_cast<const ppp_hit *>(topk), topk_cnt, kth, prev, k, r, kspec, kspec0, nQ);
    return cudaGetLastError();
}
extern "C" cudaError_t ppp_launch_spec_update(double *kspec, const double *kth, uint32_t nQ, cudaStream_t st)
{
    ppp_spec_update_kernel<<<(nQ + 255) / 256, 256, 0, st>>>(kspec, kth, nQ);
    return cudaGetLastError();
}
extern "C" cudaError_t ppp_launch_spec_verify(const void *topk, const uint32_t *topk_cnt, const double *kspec0, uint32_t k, double *last_tau,
                                              uint32_t *failed, uint32_t *nfailed, uint32_t nQ, cudaStream_t st)
{
    ppp_spec_verify_kernel<<<(nQ + 255) / 256, 256, 0, st>>>(reinterpret_cast<const ppp_hit *>(topk), topk_cnt, kspec0, k, last_tau, failed, nfailed, nQ);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------ top-k seeding
 * k-th smallest of the S provisional upper bounds of every query (dense double[nQ][S] written by the bounds-only pass):
 * a valid bound on the query's final k-th best score.  Block per query, bitonic sort of S <= 4096 doubles. */
__global__ void __launch_bounds__(256) ppp_seed_kth_kernel(const double *__restrict__ scores, uint32_t S, uint32_t k, double *__restrict__ kth,
                                                           uint32_t rs, double *__restrict__ kspec_seed, uint32_t nQ)
{
    extern __shared__ __align__(16) unsigned char smem_topk[];
    double *v = reinterpret_cast<double *>(smem_topk);
    uint32_t n2 = 1;
    while (n2 < S) n2 <<= 1;
    for (uint32_t q = blockIdx.x; q < nQ; q += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) v[i] = i < S ? scores[(size_t)q * S + i] : INFINITY;
        __syncthreads();
        for (uint32_t size = 2; size <= n2; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t i = threadIdx.x; i < (n2 >> 1); i += blockDim.x) {
                    const uint32_t lo = 2 * i - (i & (stride - 1));
                    const uint32_t hi = lo + stride;
                    const bool asc = ((lo & size) == 0);
                    const double a = v[lo], b = v[hi];
                    if ((b < a) == asc) {
                        v[lo] = b;
                        v[hi] = a;
                    }
                }
                __syncthreads();
            }
        }
        if (threadIdx.x == 0 && S >= k) {
            kth[q] = v[k - 1];
            if (kspec_seed) kspec_seed[q] = v[rs - 1]; /* rs-th smallest bound, rs <= k: speculative threshold of the first chunk */
        }
        __syncthreads();
    }
}

extern "C" cudaError_t ppp_launch_seed_kth(const void *scores, uint32_t S, uint32_t k, double *kth, uint32_t rs, double *kspec_seed, uint32_t nQ,
                                           int nsm, cudaStream_t st)
{
    if (S > 4096 || (kspec_seed && (rs < 1 || rs > k))) return cudaErrorInvalidValue;
    uint32_t n2 = 1;
    while (n2 < S) n2 <<= 1;
    const int smem = (int)(n2 * sizeof(double));
    const unsigned grid = (unsigned)min((uint32_t)(nsm * 4), nQ);
    ppp_seed_kth_kernel<<<grid, 256, smem, st>>>(reinterpret_cast<const double *>(scores), S, k, kth, rs, kspec_seed, nQ);
    return cudaGetLastError();
}
