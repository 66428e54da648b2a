/* oracle/cephes_port.h -- TEST INFRASTRUCTURE ONLY. See cephes_port.c. */
#ifndef CEPHES_PORT_H
#define CEPHES_PORT_H
#ifdef __cplusplus
extern "C" {
#endif
double cp_log(double x);
double cp_exp(double x);
double cp_log1p(double x);
double cp_expm1(double x);
double cp_pow(double x, double y);
double cp_gamma(double x);
double cp_lgam(double x);
double cp_incbet(double a, double b, double x);
double cp_bdtrc(int k, int n, double p);
void cp_bdtrc_batch(const int *k, const int *n, const double *p, double *out, long m);
extern int cp_domain_errors;
#ifdef __cplusplus
}
#endif
#endif
