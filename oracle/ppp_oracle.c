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
long ppp_oracle_list(const uint8_t *present, const uint8_t *isyes, int32_t U, int32_t max_yes,
                     const int32_t *list, int32_t L, double p, int dup, ppp_oracle_result *out)
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
        double answer = cp_bdtrc(yes - 1, number, p); /* ppp.pm:225 */
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
