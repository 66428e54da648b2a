/* oracle/ppp_oracle.h -- TEST INFRASTRUCTURE ONLY. See ppp_oracle.c. */
#ifndef PPP_ORACLE_H
#define PPP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    double score;   /* ppp.pm:244 $score       */
    int32_t yes;    /*            $best_yes    */
    int32_t number; /*            $best_number */
    int32_t depth;  /*            $best_depth  */
    int32_t pad_;
    double p;       /*            $p           */
} ppp_oracle_result;

long ppp_oracle_list(const uint8_t *present, const uint8_t *isyes, int32_t U, int32_t max_yes,
                     const int32_t *list, int32_t L, double p, int dup, ppp_oracle_result *out);
long ppp_oracle_bits(const uint32_t *P, const uint32_t *Y, const uint32_t *T, int32_t G, double p,
                     ppp_oracle_result *out);
long ppp_oracle_bits_batch(const uint32_t *P, const uint32_t *Y, const double *p, int64_t nq, const uint32_t *T,
                           int64_t nt, int32_t G, int32_t wpr, ppp_oracle_result *out, int nthreads);
/* same results as ppp_oracle_bits_batch, with a per-query cache of the pure function bdtrc(yes-1, number, p) */
long ppp_oracle_bits_batch_memo(const uint32_t *P, const uint32_t *Y, const double *p, int64_t nq, const uint32_t *T,
                                int64_t nt, int32_t G, int32_t wpr, ppp_oracle_result *out, int nthreads);
#ifdef __cplusplus
}
#endif
#endif
