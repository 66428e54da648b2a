/*
 * prophylo_b200/csrc/ppp_rank.cu -- device-side ranking of filtered results.
 *
 * Replaces, for threshold runs, the host sort that put the appended hits into the order the drivers print them in:
 * per query, ascending raw score (bin/ppp.pl:473 `sort {$a->[1] <=> $b->[1]}`), ties by target index -- i.e. the key
 * (q, score, t).  The appended 32-byte records arrive in atomic order; here they are ranked by a stable least-
 * significant-digit radix sort of a (key word, record index) pair array, one 32-bit word of the composite key at a time
 * (t, score low, score high, q), 8 bits per pass, and gathered once at the end.  Integer work, HBM-bound:
 * per pass 2 reads + 1 write of 8 bytes per hit; the final gather moves 64 bytes per hit.
 *
 * Kernels per pass: digit histogram per 4096-hit tile; per-digit exclusive scan over the tiles (one CTA per digit) and a
 * 256-entry scan of the digit totals; stable scatter (per-warp match_any ranks, warps own contiguous, ordered segments).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ppp_b200.h"

#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)

/* word w of the composite key of record r: 0 = t, 1 = score low, 2 = score high, 3 = q.  Scores are non-negative
 * doubles (probabilities); the usual sign transform keeps the order total should a negative zero ever appear. */
__device__ __forceinline__ uint32_t rank_key_word(const ppp_hit &h, int w)
{
    if (w == 0) return h.t;
    if (w == 3) return h.q;
    unsigned long long b = (unsigned long long)__double_as_longlong(h.score);
    b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    return w == 1 ? (uint32_t)b : (uint32_t)(b >> 32);
}

__global__ void __launch_bounds__(256) ppp_rank_gather_key_kernel(const ppp_hit *__restrict__ rec, const uint32_t *__restrict__ idx_in,
                                                                  uint32_t n, int word, uint32_t *__restrict__ key, uint32_t *__restrict__ idx_out)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t r = idx_in ? idx_in[i] : i;
        key[i] = rank_key_word(rec[r], word);
        if (!idx_in) idx_out[i] = i;
    }
}

__global__ void __launch_bounds__(RS_THREADS) ppp_rank_hist_kernel(const uint32_t *__restrict__ key, uint32_t n, int shift, uint32_t ntiles,
                                                                   uint32_t *__restrict__ hist)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(key[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = sh[threadIdx.x];
}

/* CTA d: exclusive scan of hist[d][0..ntiles) in place; total[d] = the digit's count */
__global__ void __launch_bounds__(1024) ppp_rank_scan_tiles_kernel(uint32_t *__restrict__ hist, uint32_t ntiles, uint32_t *__restrict__ total)
{
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_sh;
    uint32_t *row = hist + (size_t)blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < ntiles ? row[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        if (lane == 31) wsum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = wsum[lane], ws = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += u;
            }
            wsum[lane] = ws - w;
        }
        __syncthreads();
        const uint32_t carry = carry_sh;
        if (i < ntiles) row[i] = carry + wsum[warp] + s - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_sh = carry + wsum[warp] + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[blockIdx.x] = carry_sh;
}

/* exclusive scan of the 256 digit totals, in place */
__global__ void __launch_bounds__(256) ppp_rank_scan_digits_kernel(uint32_t *__restrict__ total)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = total[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int d = 0; d < 256; ++d) {
            const uint32_t v = sh[d];
            sh[d] = run;
            run += v;
        }
    }
    __syncthreads();
    total[threadIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) ppp_rank_scatter_kernel(const uint32_t *__restrict__ key_in, const uint32_t *__restrict__ idx_in,
                                                                      uint32_t n, int shift, uint32_t ntiles, const uint32_t *__restrict__ hist,
                                                                      const uint32_t *__restrict__ digit_base, uint32_t *__restrict__ key_out,
                                                                      uint32_t *__restrict__ idx_out)
{
    __shared__ uint32_t whist[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    /* warp w owns items [tile + w*512, +512) in order: round r covers 32 consecutive items */
    const uint32_t seg = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS], val[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = seg + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? key_in[i] : 0u;
        val[r] = ok ? idx_in[i] : 0u;
        const uint32_t d = ok ? ((key[r] >> shift) & 255u) : (256u + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t c = 0;
        if (ok) c = whist[warp][d];
        __syncwarp();
        if (ok && (peers & lt) == 0) whist[warp][d] = c + __popc(peers);
        __syncwarp();
        rank[r] = (uint16_t)(c + __popc(peers & lt));
    }
    __syncthreads();
    {
        const uint32_t d = threadIdx.x;
        uint32_t run = hist[(size_t)d * ntiles + blockIdx.x] + digit_base[d];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = whist[w][d];
            whist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = seg + r * 32 + lane;
        if (i < n) {
            const uint32_t o = whist[warp][(key[r] >> shift) & 255u] + rank[r];
            key_out[o] = key[r];
            idx_out[o] = val[r];
        }
    }
}

__global__ void __launch_bounds__(256) ppp_rank_gather_records_kernel(const ppp_hit *__restrict__ rec, const uint32_t *__restrict__ idx, uint32_t n,
                                                                      ppp_hit *__restrict__ out)
{
    /* two lanes per record: 16 bytes each */
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < 2u * n; j += stride) {
        const uint32_t i = j >> 1, h = j & 1u;
        reinterpret_cast<uint4 *>(out)[j] = reinterpret_cast<const uint4 *>(rec)[2u * (size_t)idx[i] + h];
    }
}

static int bits_of(uint32_t maxval)
{
    int b = 0;
    while (b < 32 && (maxval >> b)) ++b;
    return b;
}

extern "C" size_t ppp_rank_scratch_words(size_t n)
{
    const size_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    return 4 * n + 256 * ntiles + 256 + 64;
}

/* Ranks rec[0..n) by (q, score, t) into out[0..n).  scratch: ppp_rank_scratch_words(n) uint32 words.  nQ / nT bound the
 * digits of q and t that need a pass.  *launches is incremented by the kernels launched. */
extern "C" cudaError_t ppp_launch_rank_hits(const ppp_hit *rec, size_t n, uint32_t nQ, uint32_t nT, ppp_hit *out, uint32_t *scratch,
                                            int nsm, cudaStream_t st, uint32_t *launches)
{
    if (!n) return cudaSuccess;
    if (n > 0xffffffffull / 2) return cudaErrorInvalidValue;
    const uint32_t N = (uint32_t)n;
    const uint32_t ntiles = (N + RS_TILE - 1) / RS_TILE;
    uint32_t *key[2] = {scratch, scratch + n};
    uint32_t *idx[2] = {scratch + 2 * n, scratch + 3 * n};
    uint32_t *hist = scratch + 4 * n;
    uint32_t *total = hist + (size_t)256 * ntiles;
    const int word_bits[4] = {bits_of(nT ? nT - 1 : 0), 32, 32, bits_of(nQ ? nQ - 1 : 0)};
    const uint32_t ggrid = (uint32_t)nsm * 8;
    int cur = 0;
    bool have_idx = false;
    for (int w = 0; w < 4; ++w) {
        if (word_bits[w] == 0 && have_idx) continue;
        ppp_rank_gather_key_kernel<<<ggrid, 256, 0, st>>>(rec, have_idx ? idx[cur] : nullptr, N, w, key[cur], idx[cur]);
        have_idx = true;
        ++*launches;
        for (int shift = 0; shift < word_bits[w]; shift += 8) {
            ppp_rank_hist_kernel<<<ntiles, RS_THREADS, 0, st>>>(key[cur], N, shift, ntiles, hist);
            ppp_rank_scan_tiles_kernel<<<256, 1024, 0, st>>>(hist, ntiles, total);
            ppp_rank_scan_digits_kernel<<<1, 256, 0, st>>>(total);
            ppp_rank_scatter_kernel<<<ntiles, RS_THREADS, 0, st>>>(key[cur], idx[cur], N, shift, ntiles, hist, total, key[cur ^ 1], idx[cur ^ 1]);
            cur ^= 1;
            *launches += 4;
        }
    }
    ppp_rank_gather_records_kernel<<<ggrid, 256, 0, st>>>(rec, idx[cur], N, out);
    ++*launches;
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------ printed scores and cut-offs
 * What the drivers compute from a query's rows once they are sorted, on the ranked hits:
 *   printed score   sprintf("%.3f", log($score) / log(10))                                  bin/ppp.pl:476-503
 *   ppp.pl -m       marker before the first row with (|best| - |s|) / |best| * 100 > slope, best = the first row's printed
 *                   score, never updated                                                      bin/ppp.pl:477-489
 *   ppp_cutoff.pl   printed scores <= -1 in descending order; the first adjacent pair with |(last - s) / last * 100| > slope
 *                   gives CUTOFF = last                                                       bin/ppp_cutoff.pl:130-148
 * Printed scores are kept as integer thousandths.  The device's log is not the host libm's and the reference's clients hand
 * the score to the server through a %.15g string (bin/ppp.pl:582), so a value is only trusted when it is not within 1e-6
 * of a rounding boundary (those perturbations move it by < 1e-12); otherwise the query is flagged PPP_CUT_DOUBT and the
 * caller formats that query's rows itself.  The comparisons use the same IEEE double operations as Perl (no contraction). */
__device__ __forceinline__ int32_t printed_milli(double score, int *status)
{
    if (!(score > 0.0)) {
        *status = PPP_CUT_LOG0;
        return 0;
    }
    const double t = __dmul_rn(__ddiv_rn(log(score), log(10.0)), 1000.0);
    const double r = rint(t);
    if (fabs(t - r) > 0.5 - 1e-6) *status = max(*status, (int)PPP_CUT_DOUBT);
    return (int32_t)r;
}

__global__ void __launch_bounds__(256) ppp_rank_row_off_kernel(const ppp_hit *__restrict__ rec, uint32_t n, uint32_t nQ,
                                                               unsigned long long *__restrict__ row_off)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        const uint32_t qlo = i ? rec[i - 1].q + 1 : 0u;       /* queries whose rows start at i: (previous row's q, this row's q] */
        const uint32_t qhi = i < n ? rec[i].q : nQ;           /* i == n: every remaining query, and the end marker row_off[nQ] */
        for (uint32_t q = qlo; q <= qhi && q <= nQ; ++q) row_off[q] = i;
    }
}

/* block per query */
__global__ void __launch_bounds__(256) ppp_rank_cutoff_kernel(const ppp_hit *__restrict__ rec, const unsigned long long *__restrict__ row_off,
                                                              uint32_t nQ, double slope, int32_t *__restrict__ milli, int32_t *__restrict__ marker,
                                                              int32_t *__restrict__ cutoff_milli, uint8_t *__restrict__ status)
{
    __shared__ int s_status;
    __shared__ unsigned long long s_first, s_last;
    for (uint32_t q = blockIdx.x; q < nQ; q += gridDim.x) {
        const unsigned long long r0 = row_off[q], r1 = row_off[q + 1];
        if (threadIdx.x == 0) {
            s_status = PPP_CUT_OK;
            s_first = ~0ull;
            s_last = 0ull;
        }
        __syncthreads();
        int st = PPP_CUT_OK;
        for (unsigned long long i = r0 + threadIdx.x; i < r1; i += blockDim.x) milli[i] = printed_milli(rec[i].score, &st);
        if (st) atomicMax(&s_status, st);
        __syncthreads(); /* milli[] of this query is visible to the block (same-block global writes after a barrier) */
        if (r1 > r0 && s_status != PPP_CUT_LOG0) {
            const double pl = fabs(__ddiv_rn((double)milli[r0], 1000.0));
            if (pl == 0.0) {
                if (threadIdx.x == 0) s_status = max(s_status, (int)PPP_CUT_DIV0); /* ppp.pl -m: "Illegal division by zero" */
            } else {
                /* ppp.pl -m: first row i with (pl - |s_i|) / pl * 100 > slope */
                for (unsigned long long i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
                    const double ps = fabs(__ddiv_rn((double)milli[i], 1000.0));
                    if (__dmul_rn(__ddiv_rn(__dsub_rn(pl, ps), pl), 100.0) > slope) {
                        atomicMin(&s_first, i - r0);
                        break; /* later rows of this thread are later in the list */
                    }
                }
            }
            /* ppp_cutoff.pl: rows with printed score <= -1, weakest first; rows are ranked strongest first, so walking the
             * list backwards the first violating pair is the one with the LARGEST index: last = row j + 1, s = row j */
            for (unsigned long long j = r0 + threadIdx.x; j + 1 < r1; j += blockDim.x) {
                const int32_t ml = milli[j + 1], ms = milli[j];
                if (ml > -1000) break; /* ranked: every later row is weaker still */
                const double last = __ddiv_rn((double)ml, 1000.0), s = __ddiv_rn((double)ms, 1000.0);
                if (fabs(__dmul_rn(__ddiv_rn(__dsub_rn(last, s), last), 100.0)) > slope) atomicMax(&s_last, j - r0 + 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            status[q] = (uint8_t)s_status;
            marker[q] = s_first == ~0ull ? -1 : (int32_t)s_first;
            cutoff_milli[q] = s_last ? milli[r0 + s_last] : INT32_MIN; /* s_last = (index of `last`), >= 1 when found */
        }
        __syncthreads();
    }
}

extern "C" cudaError_t ppp_launch_rank_cutoffs(const ppp_hit *rec, size_t n, uint32_t nQ, double slope, unsigned long long *row_off,
                                               int32_t *milli, int32_t *marker, int32_t *cutoff_milli, uint8_t *status, int nsm, cudaStream_t st)
{
    ppp_rank_row_off_kernel<<<nsm * 4, 256, 0, st>>>(rec, (uint32_t)n, nQ, row_off);
    ppp_rank_cutoff_kernel<<<nsm * 8, 256, 0, st>>>(rec, row_off, nQ, slope, milli, marker, cutoff_milli, status);
    return cudaGetLastError();
}
