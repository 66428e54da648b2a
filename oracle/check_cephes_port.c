/*
 * oracle/check_cephes_port.c -- TEST INFRASTRUCTURE ONLY.
 * Pins oracle/cephes_port.c bit-for-bit against the reference's bundled
 * Math::Cephes 0.47 binary (oracle/_ref/Cephes.so, extracted from
 * /root/reference/cpan/x86_64-linux-thread-multi/Math-Cephes-0.47-...par).
 * usage: check_cephes_port <Cephes.so> <npoints> [seed]
 * exit status 0 iff every point is bit-identical.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cephes_port.h"

typedef double (*bdtrc_t)(int, int, double);
typedef double (*incbet_t)(double, double, double);

static uint64_t s[2];
static uint64_t rnd(void)
{ /* xorshift128+ */
    uint64_t x = s[0], y = s[1];
    s[0] = y;
    x ^= x << 23;
    s[1] = x ^ y ^ (x >> 17) ^ (y >> 26);
    return s[1] + y;
}
static double unif(void) { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
static uint64_t bits(double x)
{
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s Cephes.so npoints [seed]\n", argv[0]);
        return 2;
    }
    void *h = dlopen(argv[1], RTLD_LAZY | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 2;
    }
    bdtrc_t rb = (bdtrc_t)dlsym(h, "bdtrc");
    incbet_t ri = (incbet_t)dlsym(h, "incbet");
    long N = atol(argv[2]);
    s[0] = argc > 3 ? strtoull(argv[3], 0, 10) : 12345;
    s[1] = 0x9E3779B97F4A7C15ull;
    /* silence the binary's mtherr printf */
    FILE *devnull = freopen("/dev/null", "w", stdout);
    (void)devnull;
    long bad = 0, bad_i = 0;
    for (long i = 0; i < N; ++i) {
        int mode = i % 6;
        int n = (mode == 5) ? 1 + (int)(rnd() % 200) : 1 + (int)(rnd() % 8192);
        double p;
        switch (mode) {
        case 0: p = unif(); break;
        case 1: p = exp(log(1e-5) * unif()); break;
        case 2: { int y = 1 + (int)(rnd() % n); p = (double)y / n; if (rnd() & 1) p *= 0.3 + 1.2 * unif(); if (p > 1) p = 1; } break;
        case 3: p = (double)(1 + rnd() % 1459) / 1459.0; break;
        case 4: p = 1.0 - exp(log(1e-6) * unif()); break;
        default: p = unif(); break;
        }
        int k;
        switch ((i / 6) % 4) {
        case 0: k = (int)(rnd() % (n + 1)); break;
        case 1: { double sd = sqrt(n * p * (1 - p)) + 1; k = (int)(n * p + sd * 6 * (unif() - 0.5)); } break;
        case 2: { double sd = sqrt(n * p * (1 - p)) + 1; k = (int)(n * p + sd * 40 * unif()); } break;
        default: k = (rnd() % 5 == 0) ? 0 : (int)(rnd() % 8); break;
        }
        if (k < -1) k = -1;
        if (k > n) k = n;
        double r = rb(k, n, p), c = cp_bdtrc(k, n, p);
        if (bits(r) != bits(c)) {
            if (bad < 10)
                fprintf(stderr, "bdtrc(%d,%d,%.17g): ref %.17g port %.17g\n", k, n, p, r, c);
            bad++;
        }
        /* raw incbet with non-integer shape parameters too */
        double a = 0.5 + 4000 * unif() * unif(), b = 0.5 + 4000 * unif() * unif(), x = unif();
        r = ri(a, b, x);
        c = cp_incbet(a, b, x);
        if (bits(r) != bits(c)) {
            if (bad_i < 10)
                fprintf(stderr, "incbet(%.17g,%.17g,%.17g): ref %.17g port %.17g\n", a, b, x, r, c);
            bad_i++;
        }
    }
    fprintf(stderr, "bdtrc: %ld mismatches / %ld; incbet: %ld mismatches / %ld\n", bad, N, bad_i, N);
    return (bad || bad_i) ? 1 : 0;
}
