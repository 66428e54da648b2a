/*
 * oracle/cephes_port.c -- TEST INFRASTRUCTURE ONLY (never linked by the product).
 *
 * CPU restatement, in plain C / IEEE double, of the arithmetic that the
 * reference's hot path calls at lib/PhyloProf/Algorithm/ppp.pm:225
 *     my $answer = bdtrc( $yes - 1, $number, $p );
 * `bdtrc` is Math::Cephes 0.47 (CPAN; wraps S. Moshier's Cephes 2.x math
 * library: bdtr.c, incbet.c, gamma.c, pow.c, exp.c, log.c, unity.c, floor.c,
 * polevl.c, const.c).  Its source is NOT vendored in /root/reference -- only the
 * prebuilt binary cpan/x86_64-linux-thread-multi/Math-Cephes-0.47-...-5.10.1.par
 * (arch/auto/Math/Cephes/Cephes.so).  This file restates the published Cephes
 * algorithms; it is pinned bit-for-bit against that binary by
 * oracle/check_cephes_port.c (`make -C oracle check`: dlopen of the extracted
 * .so, 2e6 random bdtrc + 2e6 random incbet points plus every edge branch);
 * tests/test_oracle.py checks the committed golden subset
 * (tests/golden/bdtrc_points.json).
 *
 * Build with -O2 -ffp-contract=off (no FMA contraction: the 2011 x86-64
 * binary is plain SSE2).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "cephes_port.h"

/* ---- const.c (IEEE, DENORMAL defined: MINLOG = log(2^-1075)) ---- */
static const double MACHEP = 1.11022302462515654042E-16; /* 2**-53 */
static const double MAXLOG = 7.09782712893383996843E2;   /* log(2**1024) */
static const double MINLOG = -7.451332191019412076235E2; /* log(2**-1075) */
static const double MAXNUM = 1.79769313486231570815E308;
static const double LOG2E = 1.4426950408889634073599;    /* 1/log(2) */

#define MAXGAM 171.624376956302725

int cp_domain_errors = 0; /* count of mtherr(DOMAIN) events, for tests */

/* ---- polevl.c ---- */
static double polevl(double x, const double *coef, int N)
{
    const double *p = coef;
    double ans = *p++;
    int i = N;
    do
        ans = ans * x + *p++;
    while (--i);
    return ans;
}

static double p1evl(double x, const double *coef, int N)
{
    const double *p = coef;
    double ans = x + *p++;
    int i = N - 1;
    do
        ans = ans * x + *p++;
    while (--i);
    return ans;
}

/* ---- log.c ---- */
static const double LOG_P[] = {
    1.01875663804580931796E-4, 4.97494994976747001425E-1,
    4.70579119878881725854E0,  1.44989225341610930846E1,
    1.79368678507819816313E1,  7.70838733755885391666E0,
};
static const double LOG_Q[] = {
    /* 1.0 */
    1.12873587189167450590E1, 4.52279145837532221105E1,
    8.29875266912776603211E1, 7.11544750618563894466E1,
    2.31251620126765340583E1,
};
static const double LOG_R[3] = {
    -7.89580278884799154124E-1,
    1.63866645699558079767E1,
    -6.41409952958715622951E1,
};
static const double LOG_S[3] = {
    /* 1.0 */
    -3.56722798256324312549E1,
    3.12093766372244180303E2,
    -7.69691943550460008604E2,
};
#define SQRTH 0.70710678118654752440

double cp_log(double x)
{
    int e;
    double y, z;

    if (isnan(x))
        return x;
    if (x == INFINITY)
        return x;
    if (x <= 0.0) {
        if (x == 0.0)
            return -INFINITY;
        return NAN;
    }
    x = frexp(x, &e);

    if ((e > 2) || (e < -2)) {
        if (x < SQRTH) {
            e -= 1;
            z = x - 0.5;
            y = 0.5 * z + 0.5;
        } else {
            z = x - 0.5;
            z -= 0.5;
            y = 0.5 * x + 0.5;
        }
        x = z / y;
        z = x * x;
        z = x * (z * polevl(z, LOG_R, 2) / p1evl(z, LOG_S, 3));
        y = e;
        z = z - y * 2.121944400546905827679e-4;
        z = z + x;
        z = z + e * 0.693359375;
        return z;
    }

    if (x < SQRTH) {
        e -= 1;
        x = ldexp(x, 1) - 1.0;
    } else {
        x = x - 1.0;
    }
    z = x * x;
    y = x * (z * polevl(x, LOG_P, 5) / p1evl(x, LOG_Q, 5));
    if (e)
        y = y - e * 2.121944400546905827679e-4;
    y = y - ldexp(z, -1);
    z = x + y;
    if (e)
        z = z + e * 0.693359375;
    return z;
}

/* ---- exp.c ---- */
static const double EXP_P[] = {
    1.26177193074810590878E-4,
    3.02994407707441961300E-2,
    9.99999999999999999910E-1,
};
static const double EXP_Q[] = {
    3.00198505138664455042E-6,
    2.52448340349684104192E-3,
    2.27265548208155028766E-1,
    2.00000000000000000009E0,
};
static const double EXP_C1 = 6.93145751953125E-1;
static const double EXP_C2 = 1.42860682030941723212E-6;

double cp_exp(double x)
{
    double px, xx;
    int n;

    if (isnan(x))
        return x;
    if (x > MAXLOG)
        return INFINITY;
    if (x < MINLOG)
        return 0.0;

    px = floor(LOG2E * x + 0.5);
    n = (int)px;
    x -= px * EXP_C1;
    x -= px * EXP_C2;

    xx = x * x;
    px = x * polevl(xx, EXP_P, 2);
    x = px / (polevl(xx, EXP_Q, 3) - px);
    x = 1.0 + 2.0 * x;
    x = ldexp(x, n);
    return x;
}

/* ---- unity.c ---- */
static const double LP[] = {
    4.5270000862445199635215E-5, 4.9854102823193375972212E-1,
    6.5787325942061044846969E0,  2.9911919328553073277375E1,
    6.0949667980987787057556E1,  5.7112963590585538103336E1,
    2.0039553499201281259648E1,
};
static const double LQ[] = {
    /* 1.0 */
    1.5062909083469192043167E1, 8.3047565967967209469434E1,
    2.2176239823732856465394E2, 3.0909872225312059774938E2,
    2.1642788614495947685003E2, 6.0118660497603843919306E1,
};
#define SQRT2 1.41421356237309504880

double cp_log1p(double x)
{
    double z = 1.0 + x;
    if ((z < SQRTH) || (z > SQRT2))
        return cp_log(z);
    z = x * x;
    z = -0.5 * z + x * (z * polevl(x, LP, 6) / p1evl(x, LQ, 6));
    return x + z;
}

static const double EP[3] = {
    1.2617719307481059087798E-4,
    3.0299440770744196129956E-2,
    9.9999999999999999991025E-1,
};
static const double EQ[4] = {
    3.0019850513866445504159E-6,
    2.5244834034968410419224E-3,
    2.2726554820815502876593E-1,
    2.0000000000000000000897E0,
};

double cp_expm1(double x)
{
    double r, xx;
    if (isnan(x))
        return x;
    if (x == INFINITY)
        return INFINITY;
    if (x == -INFINITY)
        return -1.0;
    if ((x < -0.5) || (x > 0.5))
        return cp_exp(x) - 1.0;
    xx = x * x;
    r = x * polevl(xx, EP, 2);
    r = r / (polevl(xx, EQ, 3) - r);
    return r + r;
}

/* ---- pow.c ---- */
static const double POW_P[] = {
    4.97778295871696322025E-1,
    3.73336776063286838734E0,
    7.69994162726912503298E0,
    4.66651806774358464979E0,
};
static const double POW_Q[] = {
    /* 1.0 */
    9.33340916416696166113E0,
    2.79999886606328401649E1,
    3.35994905342304405431E1,
    1.39995542032307539578E1,
};
/* 2^(-i/16), i = 0..16 */
static const double POW_A[] = {
    1.00000000000000000000E0, 9.57603280698573700036E-1,
    9.17004043204671215328E-1, 8.78126080186649726755E-1,
    8.40896415253714502036E-1, 8.05245165974627141736E-1,
    7.71105412703970372057E-1, 7.38413072969749673113E-1,
    7.07106781186547572737E-1, 6.77127773468446325644E-1,
    6.48419777325504820276E-1, 6.20928906036742001007E-1,
    5.94603557501360513449E-1, 5.69394317378345782288E-1,
    5.45253866332628844837E-1, 5.22136891213706877402E-1,
    5.00000000000000000000E-1,
};
static const double POW_B[] = {
    0.00000000000000000000E0,  1.64155361212281360176E-17,
    4.09950501029074826006E-17, 3.97491740484881042808E-17,
    -4.83364665672645672553E-17, 1.26912513974441574796E-17,
    1.99100761573282305549E-17, -1.52339103990623557348E-17,
    0.00000000000000000000E0,
};
static const double POW_R[] = {
    1.49664108433729301083E-5, 1.54010762792771901396E-4,
    1.33335476964097721140E-3, 9.61812908476554225149E-3,
    5.55041086645832347466E-2, 2.40226506959099779976E-1,
    6.93147180559945308821E-1,
};
#define MEXP 16383.0
#define MNEXP -17183.0 /* DENORMAL */
#define LOG2EA 0.44269504088896340736

static double reduc(double x)
{
    double t = ldexp(x, 4);
    t = floor(t);
    t = ldexp(t, -4);
    return t;
}

/* integer power (powi.c), only reached when base is integer-valued */
static double cp_powi(double x, int nn)
{
    int n, e, sign, asign, lx;
    double w, y, s;

    if (x == 0.0) {
        if (nn == 0)
            return 1.0;
        else if (nn < 0)
            return INFINITY;
        else
            return 0.0;
    }
    if (nn == 0)
        return 1.0;
    if (nn == -1)
        return 1.0 / x;
    if (x < 0.0) {
        asign = -1;
        x = -x;
    } else
        asign = 0;
    if (nn < 0) {
        sign = -1;
        n = -nn;
    } else {
        sign = 1;
        n = nn;
    }
    if ((n & 1) == 0)
        asign = 0;
    s = frexp(x, &lx);
    e = (lx - 1) * n;
    if ((e == 0) || (e > 64) || (e < -64)) {
        s = (s - 7.0710678118654752e-1) / (s + 7.0710678118654752e-1);
        s = (2.9142135623730950 * s - 0.5 + lx) * nn * 0.693147180559945309417;
    } else {
        s = 0.693147180559945309417 * e;
    }
    if (s > MAXLOG)
        return asign ? -INFINITY : INFINITY;
    if (s < MINLOG)
        return 0.0;
    if ((s < (-MAXLOG + 2.0)) && (sign < 0)) {
        x = 1.0 / x;
        sign = -sign;
    }
    if (n & 1)
        y = x;
    else
        y = 1.0;
    w = x;
    n >>= 1;
    while (n) {
        w = w * w;
        if (n & 1)
            y *= w;
        n >>= 1;
    }
    if (sign < 0)
        y = 1.0 / y;
    if (asign)
        y = -y;
    return y;
}

double cp_pow(double x, double y)
{
    double w, z, W, Wa, Wb, ya, yb, u;
    double aw, ay, wy;
    int e, i, nflg, iyflg, yoddint;

    if (y == 0.0)
        return 1.0;
    if (isnan(x))
        return x;
    if (isnan(y))
        return y;
    if (y == 1.0)
        return x;
    if (!isfinite(y) && (x == 1.0 || x == -1.0))
        return NAN;
    if (x == 1.0)
        return 1.0;
    if (y >= MAXNUM) {
        if (x > 1.0)
            return INFINITY;
        if (x > 0.0 && x < 1.0)
            return 0.0;
        if (x < -1.0)
            return INFINITY;
        if (x > -1.0 && x < 0.0)
            return 0.0;
    }
    if (y <= -MAXNUM) {
        if (x > 1.0)
            return 0.0;
        if (x > 0.0 && x < 1.0)
            return INFINITY;
        if (x < -1.0)
            return 0.0;
        if (x > -1.0 && x < 0.0)
            return INFINITY;
    }
    if (x >= MAXNUM) {
        if (y > 0.0)
            return INFINITY;
        return 0.0;
    }
    iyflg = 0;
    w = floor(y);
    if (w == y)
        iyflg = 1;
    yoddint = 0;
    if (iyflg) {
        ya = fabs(y);
        ya = floor(0.5 * ya);
        yb = 0.5 * fabs(w);
        if (ya != yb)
            yoddint = 1;
    }
    if (x <= -MAXNUM) {
        if (y > 0.0) {
            if (yoddint)
                return -INFINITY;
            return INFINITY;
        }
        if (y < 0.0) {
            if (yoddint)
                return -0.0;
            return 0.0;
        }
    }
    nflg = 0;
    if (x <= 0.0) {
        if (x == 0.0) {
            if (y < 0.0) {
                if (signbit(x) && yoddint)
                    return -INFINITY;
                return INFINITY;
            }
            if (y > 0.0) {
                if (signbit(x) && yoddint)
                    return -0.0;
                return 0.0;
            }
            return 1.0;
        } else {
            if (iyflg == 0)
                return NAN;
            nflg = 1;
        }
    }
    if (iyflg) {
        i = (int)w;
        w = floor(x);
        if ((w == x) && (fabs(y) < 32768.0)) {
            w = cp_powi(x, (int)y);
            return w;
        }
    }
    if (nflg)
        x = fabs(x);

    /* results close to 1: series */
    w = x - 1.0;
    aw = fabs(w);
    ay = fabs(y);
    wy = w * y;
    ya = fabs(wy);
    if ((aw <= 1.0e-3 && ay <= 1.0) || (ya <= 1.0e-3 && ay >= 1.0)) {
        z = (((((w * (y - 5.) / 720. + 1. / 120.) * w * (y - 4.) + 1. / 24.) * w * (y - 3.) + 1. / 6.) * w *
                  (y - 2.) +
              0.5) *
             w * (y - 1.)) *
                wy +
            wy + 1.;
        goto done;
    }
    x = frexp(x, &e);

    i = 1;
    if (x <= POW_A[9])
        i = 9;
    if (x <= POW_A[i + 4])
        i += 4;
    if (x <= POW_A[i + 2])
        i += 2;
    if (x >= POW_A[1])
        i = -1;
    i += 1;

    x -= POW_A[i];
    x -= POW_B[i / 2];
    x /= POW_A[i];

    z = x * x;
    w = x * (z * polevl(x, POW_P, 3) / p1evl(x, POW_Q, 4));
    w = w - ldexp(z, -1);

    w = w + LOG2EA * w;
    z = w + LOG2EA * x;
    z = z + x;

    w = -i;
    w = ldexp(w, -4);
    w += e;

    ya = reduc(y);
    yb = y - ya;

    W = z * y + w * yb;
    Wa = reduc(W);
    Wb = W - Wa;

    W = Wa + w * ya;
    Wa = reduc(W);
    u = W - Wa;

    W = Wb + u;
    Wb = reduc(W);
    w = ldexp(Wa + Wb, 4);

    if (w > MEXP)
        return (nflg && yoddint) ? -INFINITY : INFINITY;
    if (w < (MNEXP - 1.0))
        return (nflg && yoddint) ? -0.0 : 0.0;

    e = (int)w;
    Wb = W - Wb;
    if (Wb > 0.0) {
        e += 1;
        Wb -= 0.0625;
    }
    z = Wb * polevl(Wb, POW_R, 6);
    if (e < 0)
        i = 0;
    else
        i = 1;
    i = e / 16 + i;
    e = 16 * i - e;
    w = POW_A[e];
    z = w + w * z;
    z = ldexp(z, i);

done:
    if (nflg && yoddint) {
        if (z == 0.0)
            z = -0.0;
        else
            z = -z;
    }
    return z;
}

/* ---- gamma.c ---- */
static const double GAM_P[] = {
    1.60119522476751861407E-4, 1.19135147006586384913E-3,
    1.04213797561761569935E-2, 4.76367800457137231464E-2,
    2.07448227648435975150E-1, 4.94214826801497100753E-1,
    9.99999999999999996796E-1,
};
static const double GAM_Q[] = {
    -2.31581873324120129819E-5, 5.39605580493303397842E-4,
    -4.45641913851797240494E-3, 1.18139785222060435552E-2,
    3.58236398605498653373E-2,  -2.34591795718243348568E-1,
    7.14304917030273074085E-2,  1.00000000000000000320E0,
};
static const double STIR[5] = {
    7.87311395793093628397E-4,  -2.29549961613378126380E-4,
    -2.68132617805781232825E-3, 3.47222221605458667310E-3,
    8.33333333333482257126E-2,
};
#define MAXSTIR 143.01608
static const double SQTPI = 2.50662827463100050242E0;

static double stirf(double x)
{
    double y, w, v;
    w = 1.0 / x;
    w = 1.0 + w * polevl(w, STIR, 4);
    y = cp_exp(x);
    if (x > MAXSTIR) {
        v = cp_pow(x, 0.5 * x - 0.25);
        y = v * (v / y);
    } else {
        y = cp_pow(x, x - 0.5) / y;
    }
    y = SQTPI * y * w;
    return y;
}

/* positive arguments only are reachable from bdtrc (a, b, a+b >= 1) */
double cp_gamma(double x)
{
    double p, q, z;

    if (isnan(x))
        return x;
    if (x == INFINITY)
        return x;
    if (x == -INFINITY)
        return NAN;
    q = fabs(x);
    if (q > 33.0) {
        if (x < 0.0)
            return NAN; /* reflection branch: unreachable on this path */
        return stirf(x);
    }
    z = 1.0;
    while (x >= 3.0) {
        x -= 1.0;
        z *= x;
    }
    while (x < 0.0) {
        if (x > -1.E-9)
            goto small;
        z /= x;
        x += 1.0;
    }
    while (x < 2.0) {
        if (x < 1.e-9)
            goto small;
        z /= x;
        x += 1.0;
    }
    if (x == 2.0)
        return z;
    x -= 2.0;
    p = polevl(x, GAM_P, 6);
    q = polevl(x, GAM_Q, 7);
    return z * p / q;
small:
    if (x == 0.0)
        return INFINITY;
    return z / ((1.0 + 0.5772156649015329 * x) * x);
}

static const double LGAM_A[] = {
    8.11614167470508450300E-4,  -5.95061904284301438324E-4,
    7.93650340457716943945E-4,  -2.77777777730099687205E-3,
    8.33333333333331927722E-2,
};
static const double LGAM_B[] = {
    -1.37825152569120859100E3, -3.88016315134637840924E4,
    -3.31612992738871184744E5, -1.16237097492762307383E6,
    -1.72173700820839662146E6, -8.53555664245765465627E5,
};
static const double LGAM_C[] = {
    /* 1.0 */
    -3.51815701436523470549E2, -1.70642106651881159223E4,
    -2.20528590553854454839E5, -1.13933444367982507207E6,
    -2.53252307177582951285E6, -2.01889141433532773231E6,
};
static const double LS2PI = 0.91893853320467274178;
#define MAXLGM 2.556348e305

double cp_lgam(double x)
{
    double p, q, u, w, z;

    if (isnan(x))
        return x;
    if (!isfinite(x))
        return INFINITY;
    if (x < -34.0)
        return NAN; /* reflection branch: unreachable on this path */
    if (x < 13.0) {
        z = 1.0;
        p = 0.0;
        u = x;
        while (u >= 3.0) {
            p -= 1.0;
            u = x + p;
            z *= u;
        }
        while (u < 2.0) {
            if (u == 0.0)
                return INFINITY;
            z /= u;
            p += 1.0;
            u = x + p;
        }
        if (z < 0.0)
            z = -z;
        if (u == 2.0)
            return cp_log(z);
        p -= 2.0;
        x = x + p;
        p = x * polevl(x, LGAM_B, 5) / p1evl(x, LGAM_C, 6);
        return cp_log(z) + p;
    }
    if (x > MAXLGM)
        return INFINITY;
    q = (x - 0.5) * cp_log(x) - x + LS2PI;
    if (x > 1.0e8)
        return q;
    p = 1.0 / (x * x);
    if (x >= 1000.0)
        q += ((7.9365079365079365079365e-4 * p - 2.7777777777777777777778e-3) * p + 0.0833333333333333333333) / x;
    else
        q += polevl(p, LGAM_A, 4) / x;
    (void)w;
    return q;
}

/* ---- incbet.c ---- */
static const double big = 4.503599627370496e15;
static const double biginv = 2.22044604925031308085e-16;

static double incbcf(double a, double b, double x)
{
    double xk, pk, pkm1, pkm2, qk, qkm1, qkm2;
    double k1, k2, k3, k4, k5, k6, k7, k8;
    double r, t, ans, thresh;
    int n;

    k1 = a;
    k2 = a + b;
    k3 = a;
    k4 = a + 1.0;
    k5 = 1.0;
    k6 = b - 1.0;
    k7 = k4;
    k8 = a + 2.0;

    pkm2 = 0.0;
    qkm2 = 1.0;
    pkm1 = 1.0;
    qkm1 = 1.0;
    ans = 1.0;
    r = 1.0;
    n = 0;
    thresh = 3.0 * MACHEP;
    do {
        xk = -(x * k1 * k2) / (k3 * k4);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1;
        pkm1 = pk;
        qkm2 = qkm1;
        qkm1 = qk;

        xk = (x * k5 * k6) / (k7 * k8);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1;
        pkm1 = pk;
        qkm2 = qkm1;
        qkm1 = qk;

        if (qk != 0)
            r = pk / qk;
        if (r != 0) {
            t = fabs((ans - r) / r);
            ans = r;
        } else
            t = 1.0;

        if (t < thresh)
            goto cdone;

        k1 += 1.0;
        k2 += 1.0;
        k3 += 2.0;
        k4 += 2.0;
        k5 += 1.0;
        k6 -= 1.0;
        k7 += 2.0;
        k8 += 2.0;

        if ((fabs(qk) + fabs(pk)) > big) {
            pkm2 *= biginv;
            pkm1 *= biginv;
            qkm2 *= biginv;
            qkm1 *= biginv;
        }
        if ((fabs(qk) < biginv) || (fabs(pk) < biginv)) {
            pkm2 *= big;
            pkm1 *= big;
            qkm2 *= big;
            qkm1 *= big;
        }
    } while (++n < 300);
cdone:
    return ans;
}

static double incbd(double a, double b, double x)
{
    double xk, pk, pkm1, pkm2, qk, qkm1, qkm2;
    double k1, k2, k3, k4, k5, k6, k7, k8;
    double r, t, ans, z, thresh;
    int n;

    k1 = a;
    k2 = b - 1.0;
    k3 = a;
    k4 = a + 1.0;
    k5 = 1.0;
    k6 = a + b;
    k7 = a + 1.0;
    k8 = a + 2.0;

    pkm2 = 0.0;
    qkm2 = 1.0;
    pkm1 = 1.0;
    qkm1 = 1.0;
    z = x / (1.0 - x);
    ans = 1.0;
    r = 1.0;
    n = 0;
    thresh = 3.0 * MACHEP;
    do {
        xk = -(z * k1 * k2) / (k3 * k4);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1;
        pkm1 = pk;
        qkm2 = qkm1;
        qkm1 = qk;

        xk = (z * k5 * k6) / (k7 * k8);
        pk = pkm1 + pkm2 * xk;
        qk = qkm1 + qkm2 * xk;
        pkm2 = pkm1;
        pkm1 = pk;
        qkm2 = qkm1;
        qkm1 = qk;

        if (qk != 0)
            r = pk / qk;
        if (r != 0) {
            t = fabs((ans - r) / r);
            ans = r;
        } else
            t = 1.0;

        if (t < thresh)
            goto cdone;

        k1 += 1.0;
        k2 -= 1.0;
        k3 += 2.0;
        k4 += 2.0;
        k5 += 1.0;
        k6 += 1.0;
        k7 += 2.0;
        k8 += 2.0;

        if ((fabs(qk) + fabs(pk)) > big) {
            pkm2 *= biginv;
            pkm1 *= biginv;
            qkm2 *= biginv;
            qkm1 *= biginv;
        }
        if ((fabs(qk) < biginv) || (fabs(pk) < biginv)) {
            pkm2 *= big;
            pkm1 *= big;
            qkm2 *= big;
            qkm1 *= big;
        }
    } while (++n < 300);
cdone:
    return ans;
}

static double pseries(double a, double b, double x)
{
    double s, t, u, v, n, t1, z, ai;

    ai = 1.0 / a;
    u = (1.0 - b) * x;
    v = u / (a + 1.0);
    t1 = v;
    t = u;
    n = 2.0;
    s = 0.0;
    z = MACHEP * ai;
    while (fabs(v) > z) {
        u = (n - b) * x / n;
        t *= u;
        v = t / (a + n);
        s += v;
        n += 1.0;
    }
    s += t1;
    s += ai;

    u = a * cp_log(x);
    if ((a + b) < MAXGAM && fabs(u) < MAXLOG) {
        t = cp_gamma(a + b) / (cp_gamma(a) * cp_gamma(b));
        s = s * t * cp_pow(x, a);
    } else {
        t = cp_lgam(a + b) - cp_lgam(a) - cp_lgam(b) + u + cp_log(s);
        if (t < MINLOG)
            s = 0.0;
        else
            s = cp_exp(t);
    }
    return s;
}

double cp_incbet(double aa, double bb, double xx)
{
    double a, b, t, x, xc, w, y;
    int flag;

    if (aa <= 0.0 || bb <= 0.0)
        goto domerr;

    if ((xx <= 0.0) || (xx >= 1.0)) {
        if (xx == 0.0)
            return 0.0;
        if (xx == 1.0)
            return 1.0;
    domerr:
        cp_domain_errors++;
        return 0.0;
    }

    flag = 0;
    if ((bb * xx) <= 1.0 && xx <= 0.95) {
        t = pseries(aa, bb, xx);
        goto done;
    }

    w = 1.0 - xx;

    if (xx > (aa / (aa + bb))) {
        flag = 1;
        a = bb;
        b = aa;
        xc = xx;
        x = w;
    } else {
        a = aa;
        b = bb;
        xc = w;
        x = xx;
    }

    if (flag == 1 && (b * x) <= 1.0 && x <= 0.95) {
        t = pseries(a, b, x);
        goto done;
    }

    y = x * (a + b - 2.0) - (a - 1.0);
    if (y < 0.0)
        w = incbcf(a, b, x);
    else
        w = incbd(a, b, x) / xc;

    y = a * cp_log(x);
    t = b * cp_log(xc);
    if ((a + b) < MAXGAM && fabs(y) < MAXLOG && fabs(t) < MAXLOG) {
        t = cp_pow(xc, b);
        t *= cp_pow(x, a);
        t /= a;
        t *= w;
        t *= cp_gamma(a + b) / (cp_gamma(a) * cp_gamma(b));
        goto done;
    }
    y += t + cp_lgam(a + b) - cp_lgam(a) - cp_lgam(b);
    y += cp_log(w / a);
    if (y < MINLOG)
        t = 0.0;
    else
        t = cp_exp(y);

done:
    if (flag == 1) {
        if (t <= MACHEP)
            t = 1.0 - MACHEP;
        else
            t = 1.0 - t;
    }
    return t;
}

/*
 * Runtime-linkage fact (verified on the binary, oracle/check_cephes_port.c):
 * Cephes.so calls `expm1` through its PLT (the symbol was not md_-renamed and
 * the .so is not -Bsymbolic), so inside any process that has libm loaded --
 * perl always has -- the call binds to glibc's expm1, NOT to Cephes' own
 * unity.c expm1.  The reference's effective k==0, p<0.01 branch is therefore
 *     -expm1_glibc(n * md_log1p(-p)).
 * cp_expm1 (Cephes' own) is kept for completeness and differs from glibc in
 * the last ulp on ~14% of arguments.
 */
#ifndef CP_BDTRC_EXPM1
#define CP_BDTRC_EXPM1 expm1
#endif

/* ---- bdtr.c: complemented binomial distribution, sum_{j=k+1..n} C(n,j) p^j (1-p)^(n-j) ---- */
double cp_bdtrc(int k, int n, double p)
{
    double dk, dn;

    if ((p < 0.0) || (p > 1.0))
        goto domerr;
    if (k < 0)
        return 1.0;
    if (n < k) {
    domerr:
        cp_domain_errors++;
        return 0.0;
    }
    if (k == n)
        return 0.0;
    dn = n - k;
    if (k == 0) {
        if (p < .01)
            dk = -CP_BDTRC_EXPM1(dn * cp_log1p(-p));
        else
            dk = 1.0 - cp_pow(1.0 - p, dn);
    } else {
        dk = k + 1;
        dk = cp_incbet(dk, dn, p);
    }
    return dk;
}

/* batch helper for ctypes callers */
void cp_bdtrc_batch(const int *k, const int *n, const double *p, double *out, long m)
{
    for (long i = 0; i < m; ++i)
        out[i] = cp_bdtrc(k[i], n[i], p[i]);
}
