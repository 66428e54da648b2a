/*
 * oracle/ppp_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked by the product).
 *
 * CPU restatement of the reference hot loop
 *   PhyloProf::Algorithm::ppp::get_match_score   lib/PhyloProf/Algorithm/ppp.pm:130-245
 * over integer taxon ids.  Every rule below cites the line of ppp.pm it follows.
 * The binomial tail is cp_bdtrc() from cephes_port.c (bit-identical to the
 * reference's bundled Math::Cephes 0.47 binary).  Pinned against the reference
 * itself (perl + oracle/run_reference.pl) by oracle/gen_golden.py; the committed
 * vectors live in tests/golden/.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "cephes_port.h"
#include "ppp_oracle.h"

/*
 * Generic list form.
 *   present[u], isyes[u]  query membership for taxon id u in [0,U)  (Profile::is_present / is_yes,
 *                         lib/PhyloProf/Profile.pm:284-330); ids >= U are taxa foreign to the query.
 *   max_yes               Profile::get_yes  (ppp.pm:136)
 *   list[0..L)            target taxon ids in target order, duplicates kept (ppp.pm:151-153)
 *   p                     probability, taken verbatim (ppp.pm:138-142 is evaluated by the caller)
 *   dup                   the -dup flag, with the reference's literal-key quirk (ppp.pm:189-200,214-217)
 * Returns the number of sweep steps that evaluated bdtrc (work counter for benchmarks).
 */
/* memo != NULL: bdtrc is a pure function of (yes, number) for a fixed p, so a caller that sweeps MANY targets against
 * one query may hand in a per-query cache memo[yes * memo_stride + number] (all bytes 0xFF = empty; a stored value is
 * never that NaN).  The loop, its order of evaluation and every result are unchanged -- the cache only avoids
 * re-evaluating cp_bdtrc at a point it has already been evaluated at.  Used by the full-size parity tests; the timed
 * CPU baseline of bench.py never passes a memo (the reference does not cache). */
static long oracle_list_core(const uint8_t *present, const uint8_t *isyes, int32_t U, int32_t max_yes, const int32_t *list,
                             int32_t L, double p, int dup, ppp_oracle_result *out, double *memo, int64_t memo_stride)
{
    double score = 1.0; /* ppp.pm:132 */
    int number = 0, yes = 0;
    int best_yes = 0, best_number = 0, best_depth = 0; /* ppp.pm:157-159 */
    int seen_literal = 0;                              /* $seen{key}: one counter for ALL duplicates (quirk) */
    long steps = 0;
    int32_t maxid = U;
    for (int32_t i = 0; i < L; ++i)
        if (list[i] >= maxid)
            maxid = list[i] + 1;
    uint8_t *seen = (uint8_t *)calloc((size_t)maxid + 1, 1);

    for (int32_t i = 0; i < L; ++i) { /* ppp.pm:169 */
        int32_t key = list[i];
        if (seen[key]) { /* ppp.pm:189 */
            if (!dup)
                continue; /* ppp.pm:190-192 */
            if (seen_literal >= 2)
                continue;   /* ppp.pm:194-196 */
            seen_literal++; /* ppp.pm:198-200 */
        } else {
            seen[key] = 1; /* ppp.pm:202-204 */
        }
        int is_p = key < U && present[key];
        int is_y = key < U && present[key] && isyes[key];
        if (is_p)
            number++; /* ppp.pm:207-210 */
        if (is_y) {   /* ppp.pm:212-220 */
            if (dup && seen_literal == 2) {
            } else
                yes++;
        }
        if (yes > max_yes)
            break;                                     /* ppp.pm:222-224 */
        double answer;
        if (memo) {
            double *cell = &memo[(int64_t)yes * memo_stride + number];
            uint64_t bits;
            memcpy(&bits, cell, 8);
            if (bits == ~(uint64_t)0) {
                *cell = cp_bdtrc(yes - 1, number, p);
            }
            answer = *cell;
        } else {
            answer = cp_bdtrc(yes - 1, number, p); /* ppp.pm:225 */
        }
        steps++;
        if (answer < score) { /* ppp.pm:228-237: strict <, first minimum wins */
            best_depth = i + 1;
            score = answer;
            best_yes = yes;
            best_number = number;
            continue;
        }
        if (answer <= 0)
            break; /* ppp.pm:239-241 */
    }
    free(seen);
    out->score = score;
    out->yes = best_yes;
    out->number = best_number;
    out->depth = best_depth;
    out->p = p;
    return steps;
}

long ppp_oracle_list(const uint8_t *present, const uint8_t *isyes, int32_t U, int32_t max_yes,
                     const int32_t *list, int32_t L, double p, int dup, ppp_oracle_result *out)
{
    return oracle_list_core(present, isyes, U, max_yes, list, L, p, dup, out, NULL, 0);
}

static inline int getbit(const uint32_t *w, int g) { return (w[g >> 5] >> (g & 31)) & 1; }

/*
 * Canonical-order bit form (SURVEY.md section 8d embedding): genome g is bit (g & 31) of word g >> 5;
 * the target's list is its set bits in ascending genome order (no duplicates), so
 * depth = popc(T[0..g]), number = popc((T&P)[0..g]), yes = popc((T&Y)[0..g]).
 */
long ppp_oracle_bits(const uint32_t *P, const uint32_t *Y, const uint32_t *T, int32_t G, double p,
                     ppp_oracle_result *out)
{
    uint8_t *present = (uint8_t *)malloc((size_t)G);
    uint8_t *isyes = (uint8_t *)malloc((size_t)G);
    int32_t *list = (int32_t *)malloc(sizeof(int32_t) * (size_t)G);
    int32_t L = 0, max_yes = 0;
    for (int g = 0; g < G; ++g) {
        present[g] = (uint8_t)getbit(P, g);
        isyes[g] = (uint8_t)(getbit(Y, g) & getbit(P, g));
        max_yes += isyes[g];
        if (getbit(T, g))
            list[L++] = g;
    }
    long steps = ppp_oracle_list(present, isyes, G, max_yes, list, L, p, 0, out);
    free(present);
    free(isyes);
    free(list);
    return steps;
}

/*
 * All-pairs batch over bit planes: query q (planes P,Y; row stride wpr words) x target t.
 * out is row-major [nq][nt].  nthreads pthreads over interleaved pair indices
 * (used only as bench.py's cpu_baseline / --impl reference leg and by tests).
 */
typedef struct {
    const uint32_t *P, *Y, *T;
    const double *p;
    int64_t nq, nt;
    int32_t G, wpr;
    ppp_oracle_result *out;
    int tid, nthreads;
    long steps;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    int64_t npairs = j->nq * j->nt;
    long steps = 0;
    for (int64_t i = j->tid; i < npairs; i += j->nthreads) {
        int64_t q = i / j->nt, t = i % j->nt;
        steps += ppp_oracle_bits(j->P + q * j->wpr, j->Y + q * j->wpr, j->T + t * j->wpr, j->G, j->p[q], &j->out[i]);
    }
    j->steps = steps;
    return NULL;
}

long ppp_oracle_bits_batch(const uint32_t *P, const uint32_t *Y, const double *p, int64_t nq, const uint32_t *T,
                           int64_t nt, int32_t G, int32_t wpr, ppp_oracle_result *out, int nthreads)
{
    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > 256)
        nthreads = 256;
    pthread_t th[256];
    batch_job jobs[256];
    for (int i = 0; i < nthreads; ++i) {
        batch_job j = {P, Y, T, p, nq, nt, G, wpr, out, i, nthreads, 0};
        jobs[i] = j;
        pthread_create(&th[i], NULL, batch_worker, &jobs[i]);
    }
    long steps = 0;
    for (int i = 0; i < nthreads; ++i) {
        pthread_join(th[i], NULL);
        steps += jobs[i].steps;
    }
    return steps;
}

/*
 * The same all-pairs batch with the per-query bdtrc cache of oracle_list_core (see there): identical results, about
 * 50x faster when one query meets >= 10^5 targets, which is what makes complete top-k lists over BASELINE-size target
 * sets checkable against the oracle.  Work is split by (query, target block) so that few queries still use all threads.
 */
typedef struct {
    const uint32_t *P, *Y, *T;
    const double *p;
    int64_t nq, nt, nblocks;
    int32_t G, wpr;
    ppp_oracle_result *out;
    int64_t *next; /* shared job counter */
    long steps;
} memo_job;

static void *memo_worker(void *arg)
{
    memo_job *j = (memo_job *)arg;
    const int32_t G = j->G;
    uint8_t *present = (uint8_t *)malloc((size_t)G);
    uint8_t *isyes = (uint8_t *)malloc((size_t)G);
    int32_t *list = (int32_t *)malloc(sizeof(int32_t) * (size_t)G);
    long steps = 0;
    for (;;) {
        const int64_t job = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (job >= j->nq * j->nblocks)
            break;
        const int64_t q = job / j->nblocks, b = job % j->nblocks;
        const int64_t t0 = j->nt * b / j->nblocks, t1 = j->nt * (b + 1) / j->nblocks;
        const uint32_t *P = j->P + q * j->wpr, *Y = j->Y + q * j->wpr;
        int32_t max_yes = 0, npres = 0;
        for (int g = 0; g < G; ++g) {
            present[g] = (uint8_t)getbit(P, g);
            isyes[g] = (uint8_t)(getbit(Y, g) & getbit(P, g));
            max_yes += isyes[g];
            npres += present[g];
        }
        const int64_t stride = (int64_t)npres + 1;
        const size_t cells = (size_t)(max_yes + 2) * (size_t)stride;
        double *memo = (double *)malloc(cells * sizeof(double));
        memset(memo, 0xFF, cells * sizeof(double));
        for (int64_t t = t0; t < t1; ++t) {
            const uint32_t *T = j->T + t * j->wpr;
            int32_t L = 0;
            for (int w = 0; w * 32 < G; ++w) /* the target's list = its set genomes in index order */
                for (uint32_t bits = T[w]; bits; bits &= bits - 1) {
                    const int g = w * 32 + __builtin_ctz(bits);
                    if (g < G)
                        list[L++] = g;
                }
            steps += oracle_list_core(present, isyes, G, max_yes, list, L, j->p[q], 0, &j->out[q * j->nt + t], memo, stride);
        }
        free(memo);
    }
    free(present);
    free(isyes);
    free(list);
    j->steps = steps;
    return NULL;
}

long ppp_oracle_bits_batch_memo(const uint32_t *P, const uint32_t *Y, const double *p, int64_t nq, const uint32_t *T,
                                int64_t nt, int32_t G, int32_t wpr, ppp_oracle_result *out, int nthreads)
{
    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > 256)
        nthreads = 256;
    pthread_t th[256];
    memo_job jobs[256];
    int64_t next = 0;
    int64_t nblocks = nq >= 4 * (int64_t)nthreads ? 1 : (4 * (int64_t)nthreads + nq - 1) / (nq > 0 ? nq : 1);
    if (nblocks > nt)
        nblocks = nt > 0 ? nt : 1;
    for (int i = 0; i < nthreads; ++i) {
        memo_job j = {P, Y, T, p, nq, nt, nblocks, G, wpr, out, &next, 0};
        jobs[i] = j;
        pthread_create(&th[i], NULL, memo_worker, &jobs[i]);
    }
    long steps = 0;
    for (int i = 0; i < nthreads; ++i) {
        pthread_join(th[i], NULL);
        steps += jobs[i].steps;
    }
    return steps;
}
