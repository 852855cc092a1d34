/*
 * sabc_oracle.c -- CPU ORACLE (test infrastructure; see sabc_oracle.h for the scope statement and
 * the "parity unpinned" note).  Plain C11 + OpenMP; compile with -O2 -ffp-contract=off -mfma.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * Where the reference leaves the arithmetic to Julia's RNG / libm (which cannot be reproduced
 * bit-for-bit on a GPU) the oracle follows the build's written spec instead:
 * counter-based Philox4x32-10 streams, deterministic log/exp/sincos, exact fixed-point means of
 * u, fixed-point resampling weights, radix-256 tree sums.  The CUDA product implements the same
 * spec independently; tests compare the two bit-for-bit.
 */
#define _GNU_SOURCE
#include "sabc_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static _Thread_local char g_err[512];
const char* orc_last_error(void) { return g_err; }
static int fail(int code, const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Deterministic math (spec: DESIGN.md §3.2).  Only + - * / sqrt fma and integer ops are used, */
/* so host and device agree bit-for-bit.  Algorithms: fdlibm-style log/exp (public domain,     */
/* Sun Microsystems 1993), Taylor sin/cos of pi*r on [-1/4,1/4], Stirling log-factorial.        */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define FMA(a, b, c) __builtin_fma((a), (b), (c))

static const double LN2_HI = 0x1.62e42fee00000p-1, LN2_LO = 0x1.a39ef35793c76p-33;
static const double INV_LN2 = 0x1.71547652b82fep+0;

double orc_log(double x) {
    static const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    if (x != x) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    int k = 0;
    uint64_t ix = d2u(x);
    if (ix < 0x0010000000000000ULL) { x = x * 0x1p54; k -= 54; ix = d2u(x); }
    uint32_t hx = (uint32_t)(ix >> 32);
    k += (int)(hx >> 20) - 1023;
    hx &= 0x000fffff;
    uint32_t i = (hx + 0x95f64) & 0x100000;
    ix = ((uint64_t)(hx | (i ^ 0x3ff00000)) << 32) | (ix & 0xffffffffULL);
    k += (int)(i >> 20);
    x = u2d(ix);
    double f = x - 1.0;
    double s = f / (2.0 + f);
    double dk = (double)k;
    double z = s * s;
    double w = z * z;
    double t1 = w * FMA(w, FMA(w, Lg6, Lg4), Lg2);
    double t2 = z * FMA(w, FMA(w, FMA(w, Lg7, Lg5), Lg3), Lg1);
    double R = t2 + t1;
    double hfsq = (0.5 * f) * f;
    double a = FMA(s, hfsq + R, dk * LN2_LO);
    return dk * LN2_HI - ((hfsq - a) - f);
}

double orc_exp(double x) {
    static const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                        P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                        P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 709.782712893383973096) return INFINITY;
    if (x < -745.13321910194110842) return 0.0;
    double kf = floor(FMA(x, INV_LN2, 0.5));
    int k = (int)kf;
    double hi = FMA(-kf, LN2_HI, x);
    double lo = kf * LN2_LO;
    double r = hi - lo;
    double t = r * r;
    double c = r - t * FMA(t, FMA(t, FMA(t, FMA(t, P5, P4), P3), P2), P1);
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    if (k >= -1021 && k <= 1023) return u2d(d2u(y) + ((uint64_t)(int64_t)k << 52));
    if (k > 1023) return y * 0x1p1023 * u2d((uint64_t)(k - 1023 + 1023) << 52);
    return u2d(d2u(y) + ((uint64_t)(int64_t)(k + 1000) << 52)) * 0x1p-1000;
}

/* sin(2 pi u), cos(2 pi u) for u in [0,1). */
void orc_sincos2pi(double u, double* sn, double* cs) {
    static const double S[10] = { 0x1.921fb54442d18p+1, -0x1.4abbce625be53p+2, 0x1.466bc6775aae2p+1,
        -0x1.32d2cce62bd86p-1, 0x1.50783487ee782p-4, -0x1.e3074fde8871fp-8, 0x1.e8f434d018d63p-12,
        -0x1.6fadb9f155744p-16, 0x1.aaec32af93359p-21, -0x1.8a404211f9547p-26 };
    static const double C[10] = { 1.0, -0x1.3bd3cc9be45dep+2, 0x1.03c1f081b5ac4p+2, -0x1.55d3c7e3cbffap+0,
        0x1.e1f506891babbp-3, -0x1.a6d1f2a204a8cp-6, 0x1.f9d38a3763cc3p-10, -0x1.b6e24f44b128fp-14,
        0x1.20c62c2f2d7f5p-18, -0x1.2a0c591af8314p-23 };
    double t = 2.0 * u;                    /* angle = pi * t, t in [0,2) */
    double qf = floor(FMA(2.0, t, 0.5));   /* nearest multiple of 1/2 */
    int q = (int)qf;
    double r = t - 0.5 * qf;               /* exact, |r| <= 1/4 */
    double r2 = r * r;
    double ps = S[9], pc = C[9];
    for (int k = 8; k >= 0; --k) { ps = FMA(ps, r2, S[k]); pc = FMA(pc, r2, C[k]); }
    double sr = r * ps, cr = pc;
    switch (q & 3) {
        case 0: *sn = sr;  *cs = cr;  break;
        case 1: *sn = cr;  *cs = -sr; break;
        case 2: *sn = -sr; *cs = -cr; break;
        default: *sn = -cr; *cs = sr; break;
    }
}

/* log(k!) for integer-valued k >= 0: table to 16, Stirling series of lgamma(k+1) beyond. */
double orc_logfact(double k) {
    static const double LF[17] = { 0.0, 0.0, 0x1.62e42fefa39efp-1, 0x1.cab0bfa2a2002p+0, 0x1.96ca77c922cf9p+1,
        0x1.326643c4479c9p+2, 0x1.a51273acf01cap+2, 0x1.10ce1f32dcc30p+3, 0x1.5358e82fcb70dp+3,
        0x1.99a8921a7f7cfp+3, 0x1.e357590954d15p+3, 0x1.180973f3a8d74p+4, 0x1.3fcba16d50143p+4,
        0x1.68d5a9c3b32cep+4, 0x1.930f3df162a42p+4, 0x1.be636a63fd346p+4, 0x1.eabff061f1a84p+4 };
    static const double HALF_LOG2PI = 0x1.d67f1c864beb5p-1;
    if (k <= 16.0) return LF[(int)k];
    double x = k + 1.0;
    double lx = orc_log(x);
    double r = 1.0 / x, r2 = r * r;
    double p = FMA(-r2, 1.0 / 1680.0, 1.0 / 1260.0);
    p = FMA(-r2, p, 1.0 / 360.0);
    p = FMA(-r2, p, 1.0 / 12.0);
    double t = (x - 0.5) * lx;
    t = t - x;
    t = t + HALF_LOG2PI;
    return t + r * p;
}

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; Random123).  Streams replace Julia's task-local        */
/* Xoshiro (src/SimulatedAnnealingABC.jl:172-179,308-331 draw through rand()/randn()).          */
/* ------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { KIND_CTRL = 0, KIND_MODEL = 1, KIND_RW = 3, KIND_PRIOR = 4, KIND_RESAMPLE = 6 };

typedef struct { uint64_t seed; uint32_t particle; uint64_t sweep; uint32_t kind; uint32_t block; } stream_t;

/* one 128-bit block of a stream -> two 64-bit words */
static void stream_block(const stream_t* st, uint32_t j, uint64_t* a, uint64_t* b) {
    uint32_t ctr[4] = { st->particle, (uint32_t)st->sweep, j, st->kind | ((uint32_t)(st->sweep >> 32) << 4) };
    uint32_t key[2] = { (uint32_t)st->seed, (uint32_t)(st->seed >> 32) };
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    *a = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    *b = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
}
static void stream_next(stream_t* st, uint64_t* a, uint64_t* b) { stream_block(st, st->block, a, b); st->block++; }

static inline double u53(uint64_t x) { return (double)(x >> 11) * 0x1p-53; }            /* [0,1)  */
static inline double u53_open0(uint64_t x) { return (double)((x >> 11) + 1) * 0x1p-53; } /* (0,1]  */
static inline double u53_mid(uint64_t x) { return ((double)(x >> 11) + 0.5) * 0x1p-53; }   /* (0,1)  */
static inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }

/* Box-Muller pair from one block (53-bit uniforms): the normal behind the PRIOR draws (initialization) */
void orc_normal_pair(uint64_t a, uint64_t b, double* z0, double* z1) {
    double u1 = u53_open0(a), u2 = u53(b);
    double r = sqrt(-2.0 * orc_log(u1));
    double sn, cs;
    orc_sincos2pi(u2, &sn, &cs);
    *z0 = r * cs; *z1 = r * sn;
}
/* randn() on the hot path (src/proposals.jl:42,54,110 and the models' noise; Julia's randn() is a ziggurat as well):
   256-layer ziggurat (Marsaglia & Tsang 2000) on ONE 64-bit word -- spec DESIGN.md section 3.3:
     bits 0..7 layer i, bit 8 sign, bits 11..63 the 53-bit integer m; x = m * W[i]; m < K[i] -> +-x (98.5 %).
   Otherwise one extra block e of the same stream: layer 0 -> tail beyond R (xx = -log(U(e.a))/R, yy = -log(U(e.b)), accept R + xx
   when 2 yy > xx^2, else the next block); layer >= 1 -> wedge (accept x when F[i] + U(e.a)(F[i+1] - F[i]) < exp(-x^2/2), else start
   again with the word e.b).  Tables: tools/gen_ziggurat.py. */
#include "zig_tables.inc"
static double zig_normal(uint64_t w, stream_t* st) {
    for (;;) {
        uint32_t i = (uint32_t)w & 255u;
        uint64_t m = w >> 11;
        int neg = (int)((w >> 8) & 1u);
        double x = (double)m * ZIG_KW[i].w;
        if (m < ZIG_KW[i].k) return neg ? -x : x;
        uint64_t ea, eb;
        stream_next(st, &ea, &eb);
        if (i == 0) {
            for (;;) {
                double xx = (-orc_log(u53_open0(ea))) * ZIG_INV_R;
                double yy = -orc_log(u53_open0(eb));
                if (yy + yy > xx * xx) return neg ? -(ZIG_R + xx) : ZIG_R + xx;
                stream_next(st, &ea, &eb);
            }
        }
        double y = ZIG_F[i] + u53(ea) * (ZIG_F[i + 1] - ZIG_F[i]);
        if (y < orc_exp((-0.5 * x) * x)) return neg ? -x : x;
        w = eb;
    }
}
/* two / one normals from the next block of a stream */
static void normal2(stream_t* st, double* z0, double* z1) {
    uint64_t a, b; stream_next(st, &a, &b);
    *z0 = zig_normal(a, st); *z1 = zig_normal(b, st);
}
/* the normals that n_pairs consecutive normal2() calls on the model stream (seed, particle, sweep) return */
void orc_normal_stream(uint64_t seed, uint32_t particle, uint64_t sweep, int32_t n_pairs, double* out) {
    stream_t st = { seed, particle, sweep, KIND_MODEL, 0 };
    for (int32_t k = 0; k < n_pairs; ++k) normal2(&st, out + 2 * k, out + 2 * k + 1);
}
double orc_zig_normal(uint64_t seed, uint32_t particle, uint64_t sweep, uint32_t* block_io) {
    stream_t st = { seed, particle, sweep, KIND_MODEL, *block_io };
    uint64_t a, b; stream_next(&st, &a, &b);
    double z = zig_normal(a, &st);
    *block_io = st.block;
    return z;
}

/* Poisson sampler (spec §3.3): one-uniform sequential-search inversion below 10, Hoermann's PTRS (1993) above,
   with the acceptance tests rearranged to one reciprocal and one logarithm of a quotient.  One ATTEMPT consumes one
   Philox block (none when lam <= 0); poisson() repeats attempts until one accepts. */
static int poisson_attempt(double lam, stream_t* st, int64_t* k_out) {
    uint64_t a, b;
    if (!(lam > 0.0)) { *k_out = 0; return 1; }
    stream_next(st, &a, &b);
    if (lam < 10.0) {
        double U = u53(a);
        double p = orc_exp(-lam), F = p;
        int64_t k = 0;
        while (U > F && k < 1024) { k++; p = (p * lam) * (1.0 / (double)k); F = F + p; }
        *k_out = k;
        return 1;
    }
    double slam = sqrt(lam);
    double bb = 0.931 + 2.53 * slam;
    double aa = -0.059 + 0.02483 * bb;
    double U = u53(a) - 0.5, V = u53(b);
    double us = 0.5 - fabs(U);
    double r = 1.0 / us;
    double kf = floor(((2.0 * aa) * r + bb) * U + lam + 0.43);
    if (us >= 0.07 && (0.9277 - V) * (bb - 2.0) >= 3.6224) { *k_out = (int64_t)kf; return 1; }
    if (kf < 0.0 || (us < 0.013 && V > us)) return 0;
    double bm = bb - 3.4;
    double num = V * (1.1239 * bm + 1.1328);
    double den = bm * ((aa * r) * r + bb);
    double lhs = orc_log(num / den);
    double rhs = (-lam + kf * orc_log(lam)) - orc_logfact(kf);
    if (lhs <= rhs) { *k_out = (int64_t)kf; return 1; }
    return 0;
}
static int64_t poisson(double lam, stream_t* st) {
    int64_t k;
    while (!poisson_attempt(lam, st, &k)) {}
    return k;
}
int64_t orc_poisson(double lam, uint64_t seed, uint32_t particle, uint64_t sweep, uint32_t* block_io) {
    stream_t st = { seed, particle, sweep, KIND_MODEL, *block_io };
    int64_t k = poisson(lam, &st);
    *block_io = st.block;
    return k;
}

/* radix-256 tree sum (spec §3.5): groups of 256 = 8 x 32-lane butterflies, then 8 sequential adds */
static double group256(const double* v, int64_t n) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) {
        double l[32];
        for (int i = 0; i < 32; ++i) { int64_t idx = 32 * w + i; l[i] = idx < n ? v[idx] : 0.0; }
        for (int off = 16; off >= 1; off >>= 1)
            for (int i = 0; i < off; ++i) l[i] = l[i] + l[i + off];
        tot = (w == 0) ? l[0] : tot + l[0];
    }
    return tot;
}
double orc_treesum(const double* x, int64_t n) {
    if (n <= 0) return 0.0;
    if (n <= 256) return group256(x, n);
    int64_t g = (n + 255) / 256;
    double* part = (double*)malloc((size_t)g * sizeof(double));
    #pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < g; ++c) {
        int64_t rem = n - 256 * c;
        part[c] = group256(x + 256 * c, rem < 256 ? rem : 256);
    }
    double r = orc_treesum(part, g);
    free(part);
    return r;
}

/* exact order-independent mean of u in [0,1] (spec §3.4): u*2^62 split into 31-bit limbs */
static void u_limbs(double u, uint64_t* hi, uint64_t* lo) {
    double c = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    if (c != c) c = 0.0;
    uint64_t q = (uint64_t)(c * 0x1p62);
    *hi = q >> 31; *lo = q & 0x7fffffffULL;
}
static double limbs_to_sum(uint64_t hi, uint64_t lo) { return ((double)hi * 2147483648.0 + (double)lo) * 0x1p-62; }
void orc_exact_mean_u(const double* u, int64_t n, double* mean_out) {
    uint64_t H = 0, L = 0;
    for (int64_t i = 0; i < n; ++i) { uint64_t h, l; u_limbs(u[i], &h, &l); H += h; L += l; }
    *mean_out = limbs_to_sum(H, L) / (double)n;
}

/* ------------------------------------------------------------------------------------------ */
/* ECDF  (src/cdf_estimators.jl:23-44 build, :68-70 evaluate; Interpolations.jl ^0.15         */
/* LinearMonotonicInterpolation + Flat extrapolation, restated from SURVEY.md App. A1/B1)      */
/* ------------------------------------------------------------------------------------------ */
static int cmp_dbl(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }

int64_t orc_ecdf_build(const double* x, int64_t n, double* K) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) if (x[i] > 0.0) K[1 + m++] = x[i];   /* cdf_estimators.jl:29 filter(e -> e > 0) */
    if (m == 0) return -1;                                               /* maximum(x) of empty throws */
    qsort(K + 1, (size_t)m, sizeof(double), cmp_dbl);                    /* :33 sort(x) */
    K[0] = 0.0;                                                          /* :33 [0; ...] */
    K[m + 1] = K[m] * 1.5;                                               /* :33 maximum(x)*a, a = 1.5 (:32) */
    return m + 2;
}

static double ecdf_eval1(const double* K, int64_t L, double rho) {
    double x = rho > K[L - 1] ? K[L - 1] : (rho < K[0] ? K[0] : rho);    /* Flat(): clamp to the knot range */
    int64_t lo = 0, hi = L;                                              /* searchsortedfirst: first K[j] >= x */
    while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (K[mid] < x) lo = mid + 1; else hi = mid; }
    int64_t j = lo;
    if (j > 0) j -= 1;                                                   /* k > 1 && (k -= 1) */
    if (j > L - 2) j = L - 2;                                            /* m[L] = m[L-1]; unreachable after the clamp */
    double Lm1 = (double)(L - 1);
    double y0 = (double)j / Lm1, y1 = (double)(j + 1) / Lm1;            /* :36 range(0, stop=1, length=L) */
    double m = (y1 - y0) / (K[j + 1] - K[j]);                            /* calcTangents: Δ_k */
    double xd = x - K[j];
    return y0 + m * xd;                                                  /* A[k] + m[k]*xdiff (c = d = 0) */
}
void orc_ecdf_eval(const double* K, int64_t L, const double* rho, int64_t m, double* u) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; ++i) u[i] = ecdf_eval1(K, L, rho[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* Accept rule  (src/SimulatedAnnealingABC.jl:314-329)                                         */
/* ------------------------------------------------------------------------------------------ */
static int accept1(int32_t s, const double* u_old, int64_t ldo, const double* u_new, int64_t ldn,
                   const double* eps, int32_t n_eps, double dlp, double log_factor, double U) {
    double S = 0.0;
    for (int32_t j = 0; j < s; ++j) {
        double e = eps[n_eps == 1 ? 0 : j];
        double t = (u_old[j * ldo] - u_new[j * ldn]) / e;                /* (u[i,:] .- u_proposal) ./ ϵ */
        S = (j == 0) ? t : S + t;                                        /* sum(...) left to right */
    }
    double Lacc = (dlp + S) + log_factor;                                /* :318-319 */
    return orc_log(U) < Lacc;                                            /* :324 log(rand()) < log_accept_prob */
}
void orc_accept_step(int64_t m, int32_t s, const double* u_old, const double* u_new, const double* eps, int32_t n_eps,
                     const double* dlogprior, const double* log_factor, const double* uniform, uint8_t* acc) {
    for (int64_t i = 0; i < m; ++i) {
        if (dlogprior[i] == -INFINITY) { acc[i] = (uint8_t)(orc_log(uniform[i]) < -INFINITY); continue; } /* :320-322 */
        acc[i] = (uint8_t)accept1(s, u_old + i, m, u_new + i, m, eps, n_eps, dlogprior[i], log_factor[i], uniform[i]);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* epsilon updates  (src/SimulatedAnnealingABC.jl:92-95 single, :100-117 multi)                */
/* ------------------------------------------------------------------------------------------ */
/* literal restatement: Roots.find_zero(ϵ -> ϵ^2 + v*ϵ^(3/2) - ū^2, (0, ū)) = bisection to adjacent floats */
double orc_eps_single_bisect(double ubar, double v) {
    if (ubar <= 2.220446049250313e-16) return 0.0;
    double lo = 0.0, hi = ubar, u2 = ubar * ubar;
    for (int it = 0; it < 4000; ++it) {
        double mid = lo + (hi - lo) * 0.5;
        if (mid <= lo || mid >= hi) break;
        double g = mid * mid + v * (mid * sqrt(mid)) - u2;
        if (g == 0.0) return mid;
        if (g < 0.0) lo = mid; else hi = mid;
    }
    return hi;
}
/* engine spec (§3.6): monotone Newton on s = sqrt(ϵ): s^4 + v s^3 - ū^2 = 0 from s0 = sqrt(ū) */
double orc_eps_single(double ubar, double v) {
    if (ubar <= 2.220446049250313e-16) return 0.0;
    double u2 = ubar * ubar;
    double s = sqrt(ubar);
    for (int it = 0; it < 64; ++it) {
        double s2 = s * s, s3 = s2 * s;
        double f = (s3 * s + v * s3) - u2;
        double fp = 4.0 * s3 + (3.0 * v) * s2;
        double sn = s - f / fp;
        if (!(sn < s)) break;
        s = sn;
    }
    return s * s;
}

static double ipow(double x, int m) { double p = 1.0; for (int i = 0; i < m; ++i) p = p * x; return p; }
static double powhalf(double x, int m) { return (m & 1) ? sqrt(x) * ipow(x, (m - 1) / 2) : ipow(x, m / 2); } /* x^(m/2) */

/* mean of Exp(β) truncated to [0,1] minus ū, and its derivative (stable form of :113) */
static void trunc_exp_mean(double beta, double* g, double* dg) {
    double ab = fabs(beta);
    if (ab < 0.01) {
        double b2 = beta * beta;
        *g = 0.5 - beta * (1.0 / 12.0 - b2 * (1.0 / 720.0 - b2 * (1.0 / 30240.0)));
        *dg = -(1.0 / 12.0) + b2 * (1.0 / 240.0 - b2 * (1.0 / 6048.0));
        return;
    }
    double t = orc_exp(-beta);
    double omt = 1.0 - t;
    *g = 1.0 / beta - t / omt;
    *dg = t / (omt * omt) - 1.0 / (beta * beta);
}
int orc_eps_multi(const double* ubar, int32_t n, double v, double* eps_out) {
    double cn = 1.0;                                   /* (2n+2)!/((n+1)!(n+2)!) = Catalan(n+1) */
    for (int k = 2; k <= n + 1; ++k) cn = cn * (double)(n + 1 + k) / (double)k;
    for (int32_t i = 0; i < n; ++i) {
        double ui = ubar[i];
        if (ui <= 2.220446049250313e-16) return fail(-5, "Division by zero - Mean u for a statistic <= eps()"); /* :107-109 */
        double sumq = 0.0, prodq = 1.0;
        for (int32_t j = 0; j < n; ++j) {
            double q = ubar[j] / ui;                                      /* :110 */
            double t = powhalf(q, n);
            sumq = (j == 0) ? t : sumq + t;
            prodq = (j == 0) ? q : prodq * q;
        }
        double num = 1.0 + sumq;                                          /* :111 */
        double den = ((cn * (double)(n + 1)) * powhalf(ui, n + 2)) * prodq; /* :112 */
        double beta = 1.0 / ui;                                           /* :113 start value */
        for (int it = 0; it < 100; ++it) {
            double g, dg;
            trunc_exp_mean(beta, &g, &dg);
            double bn = beta - (g - ui) / dg;
            double diff = fabs(bn - beta);
            beta = bn;
            if (diff <= 4.0e-16 * fabs(bn)) break;
        }
        eps_out[i] = 1.0 / (beta + (v * num) / den);                      /* :114 */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Resampling  (src/SimulatedAnnealingABC.jl:124-137)                                          */
/* ------------------------------------------------------------------------------------------ */
void orc_resample_weights(const double* u, int64_t n, int32_t s, const double* ubar, double delta,
                          double* w_out, uint64_t* q_out) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int32_t j = 0; j < s; ++j) {
            double t = (u[i + j * n] * delta) / ubar[j];                  /* u[:,i] .* δ ./ ū[i]  (:127) */
            acc = (j == 0) ? t : acc + t;
        }
        double w = orc_exp(-acc);
        if (w_out) w_out[i] = w;
        q_out[i] = (uint64_t)(w * 4294967296.0);                          /* fixed-point weight (spec §3.7) */
    }
}
/* N iid categorical draws ∝ q (sample(1:n, weights(w), n, replace=true), :129) by inversion of the
   exact integer prefix sums; draw k uses the Philox block (k, resample_count, kind RESAMPLE). */
void orc_resample_indices(const uint64_t* q, int64_t n, uint64_t seed, uint64_t resample_count, int64_t* idx) {
    uint64_t* P = (uint64_t*)malloc((size_t)n * sizeof(uint64_t));
    uint64_t run = 0;
    for (int64_t i = 0; i < n; ++i) { run += q[i]; P[i] = run; }
    #pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        uint32_t ctr[4] = { (uint32_t)k, (uint32_t)resample_count, (uint32_t)((uint64_t)k >> 32), KIND_RESAMPLE };
        uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, o[4];
        orc_philox4x32_10(ctr, key, o);
        uint64_t r = mulhi64((uint64_t)o[0] | ((uint64_t)o[1] << 32), run);
        int64_t lo = 0, hi = n;                                           /* first i with P[i] > r */
        while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (P[mid] <= r) lo = mid + 1; else hi = mid; }
        idx[k] = lo < n ? lo : n - 1;
    }
    free(P);
}

/* ------------------------------------------------------------------------------------------ */
/* Prior  (logpdf / rand call sites src/SimulatedAnnealingABC.jl:163,174,314,318;              */
/* Distributions ^0.25 formulas)                                            */
/* ------------------------------------------------------------------------------------------ */
static const double LOG2PI = 0x1.d67f1c864beb5p+0;
typedef struct { int32_t kind; double p0, p1, c; } prior1_t;   /* c = log σ  |  -log(b-a) | Gamma / Beta: log normaliser */

/* log Gamma(x), x > 0 (spec §3.2): recurrence up to x >= 16, then Stirling's series through 1/(1188 x^9) */
double orc_lgamma(double x) {
    if (!(x > 0.0)) return INFINITY;
    double prod = 1.0;
    while (x < 16.0) { prod = prod * x; x = x + 1.0; }
    double lx = orc_log(x);
    double r = 1.0 / x, r2 = r * r;
    double p = fma(-r2, 1.0 / 1188.0, 1.0 / 1680.0);
    p = fma(-r2, p, 1.0 / 1260.0);
    p = fma(-r2, p, 1.0 / 360.0);
    p = fma(-r2, p, 1.0 / 12.0);
    double t = (x - 0.5) * lx;
    t = t - x;
    t = t + 0x1.d67f1c864beb5p-1;
    t = t + r * p;
    return t - orc_log(prod);
}

static void prior_prepare(int32_t d, const int32_t* kind, const double* par, prior1_t* out) {
    for (int32_t c = 0; c < d; ++c) {
        out[c].kind = kind[c]; out[c].p0 = par[2 * c]; out[c].p1 = par[2 * c + 1];
        if (kind[c] == ORC_PRIOR_UNIFORM) out[c].c = -orc_log(par[2 * c + 1] - par[2 * c]);
        else if (kind[c] == ORC_PRIOR_EXPONENTIAL) out[c].c = orc_log(par[2 * c]);
        else if (kind[c] == ORC_PRIOR_GAMMA) out[c].c = orc_lgamma(par[2 * c]) + par[2 * c] * orc_log(par[2 * c + 1]);
        else if (kind[c] == ORC_PRIOR_BETA) out[c].c = (orc_lgamma(par[2 * c]) + orc_lgamma(par[2 * c + 1])) - orc_lgamma(par[2 * c] + par[2 * c + 1]);
        else if (kind[c] == ORC_PRIOR_LAPLACE) out[c].c = orc_log(2.0 * par[2 * c + 1]);
        else if (kind[c] == ORC_PRIOR_WEIBULL) out[c].c = orc_log(par[2 * c] / par[2 * c + 1]);
        else if (kind[c] == ORC_PRIOR_INVGAMMA) out[c].c = orc_lgamma(par[2 * c]) - par[2 * c] * orc_log(par[2 * c + 1]);
        else out[c].c = orc_log(par[2 * c + 1]);
    }
}
static double prior_logpdf(int32_t d, const prior1_t* pr, const double* th) {
    double lp = 0.0;
    for (int32_t c = 0; c < d; ++c) {
        double t, x = th[c];
        switch (pr[c].kind) {
        case ORC_PRIOR_NORMAL: { double z = (x - pr[c].p0) / pr[c].p1; t = -((z * z + LOG2PI) * 0.5) - pr[c].c; break; }
        case ORC_PRIOR_UNIFORM: t = (x >= pr[c].p0 && x <= pr[c].p1) ? pr[c].c : -INFINITY; break;
        case ORC_PRIOR_EXPONENTIAL: t = x >= 0.0 ? (-(x / pr[c].p0)) - pr[c].c : -INFINITY; break;
        case ORC_PRIOR_GAMMA: {                  /* Distributions: xlogy(α-1, x) - x/θ - (lgamma(α) + α log θ) */
            if (!(x >= 0.0)) { t = -INFINITY; break; }
            double a1 = pr[c].p0 - 1.0;
            double tt = a1 == 0.0 ? 0.0 : a1 * orc_log(x);
            t = (tt - x / pr[c].p1) - pr[c].c; break; }
        case ORC_PRIOR_BETA: {                   /* xlogy(α-1, x) + xlog1py(β-1, -x) - logbeta(α, β) on [0, 1] */
            if (!(x >= 0.0 && x <= 1.0)) { t = -INFINITY; break; }
            double a1 = pr[c].p0 - 1.0, b1 = pr[c].p1 - 1.0;
            double t0 = a1 == 0.0 ? 0.0 : a1 * orc_log(x);
            double t1 = b1 == 0.0 ? 0.0 : b1 * orc_log(1.0 - x);
            t = (t0 + t1) - pr[c].c; break; }
        case ORC_PRIOR_CAUCHY: {                 /* -(log1psq(z) + logπ + log σ) */
            double z = (x - pr[c].p0) / pr[c].p1;
            t = -((orc_log(1.0 + z * z) + 0x1.250d048e7a1bdp+0) + pr[c].c); break; }
        case ORC_PRIOR_LAPLACE: t = -(fabs(x - pr[c].p0) / pr[c].p1 + pr[c].c); break;   /* -(|x-μ|/θ + log 2θ) */
        case ORC_PRIOR_WEIBULL: {                /* log(α/θ) + xlogy(α-1, z) - z^α, z = x/θ */
            if (!(x >= 0.0)) { t = -INFINITY; break; }
            double lz = orc_log(x / pr[c].p1), a1 = pr[c].p0 - 1.0;
            double tt = a1 == 0.0 ? 0.0 : a1 * lz;
            t = (pr[c].c + tt) - orc_exp(pr[c].p0 * lz); break; }
        case ORC_PRIOR_INVGAMMA:                 /* α log θ - lgamma(α) - (α+1) log x - θ/x */
            if (!(x > 0.0)) { t = -INFINITY; break; }
            t = (-((pr[c].p0 + 1.0) * orc_log(x)) - pr[c].p1 / x) - pr[c].c; break;
        default:                                                          /* LogNormal */
            if (!(x > 0.0)) { t = -INFINITY; break; }
            { double lx = orc_log(x), z = (lx - pr[c].p0) / pr[c].p1; t = (-((z * z + LOG2PI) * 0.5) - pr[c].c) - lx; }
        }
        lp = (c == 0) ? t : lp + t;
    }
    return lp;
}
double orc_prior_logpdf(int32_t d, const int32_t* kind, const double* par, const double* theta) {
    prior1_t pr[16];
    prior_prepare(d, kind, par, pr);
    return prior_logpdf(d, pr, theta);
}
/* Gamma(a, 1), Marsaglia & Tsang 2000 (spec §3.3): attempt t of component comp reads blocks comp + 256 (base + 2t + 1)
 * (normal) and comp + 256 (base + 2t + 2) (word a: acceptance uniform, word b: U^(1/a) scaling when a < 1) */
static double gamma_std(double a, stream_t* st, uint32_t comp, uint32_t base) {
    double ae = a < 1.0 ? a + 1.0 : a;
    double d = ae - 1.0 / 3.0;
    double cc = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0; t < 100000u; ++t) {
        uint64_t a0, b0, a1, b1;
        double z, z1;
        stream_block(st, comp + 256u * (base + 2u * t + 1u), &a0, &b0);
        orc_normal_pair(a0, b0, &z, &z1);
        stream_block(st, comp + 256u * (base + 2u * t + 2u), &a1, &b1);
        double v = 1.0 + cc * z;
        if (!(v > 0.0)) continue;
        v = (v * v) * v;
        double rhs = ((0.5 * z) * z + d) - d * v + d * orc_log(v);
        if (orc_log(u53_open0(a1)) < rhs) {
            double g = d * v;
            if (a < 1.0) g = g * orc_exp(orc_log(u53_open0(b1)) / a);
            return g;
        }
    }
    return d;
}
static void prior_rand(int32_t d, const prior1_t* pr, uint64_t seed, uint32_t particle, double* th) {
    stream_t st = { seed, particle, 0, KIND_PRIOR, 0 };
    for (int32_t c = 0; c < d; ++c) {
        uint64_t a, b;
        stream_block(&st, (uint32_t)c, &a, &b);
        double z0, z1;
        switch (pr[c].kind) {
        case ORC_PRIOR_NORMAL: orc_normal_pair(a, b, &z0, &z1); th[c] = pr[c].p0 + pr[c].p1 * z0; break;
        case ORC_PRIOR_UNIFORM: th[c] = pr[c].p0 + (pr[c].p1 - pr[c].p0) * u53(a); break;
        case ORC_PRIOR_EXPONENTIAL: th[c] = pr[c].p0 * (-orc_log(u53_open0(a))); break;
        case ORC_PRIOR_GAMMA: th[c] = pr[c].p1 * gamma_std(pr[c].p0, &st, (uint32_t)c, 0u); break;
        case ORC_PRIOR_BETA: {
            double g1 = gamma_std(pr[c].p0, &st, (uint32_t)c, 0u), g2 = gamma_std(pr[c].p1, &st, (uint32_t)c, 1u << 20);
            th[c] = g1 / (g1 + g2); break; }
        case ORC_PRIOR_CAUCHY: {                 /* quantile μ + σ tan(π(u - 1/2)) = μ - σ cos(πu)/sin(πu), u in (0,1) */
            double sn, cs; orc_sincos2pi(0.5 * u53_mid(a), &sn, &cs);
            th[c] = pr[c].p0 - pr[c].p1 * (cs / sn); break; }
        case ORC_PRIOR_LAPLACE: {                /* quantile */
            double u = u53_mid(a);
            th[c] = u < 0.5 ? pr[c].p0 + pr[c].p1 * orc_log(2.0 * u) : pr[c].p0 - pr[c].p1 * orc_log(2.0 * (1.0 - u)); break; }
        case ORC_PRIOR_WEIBULL: th[c] = pr[c].p1 * orc_exp(orc_log(-orc_log(u53_open0(a))) / pr[c].p0); break;   /* θ E^(1/α) */
        case ORC_PRIOR_INVGAMMA: th[c] = pr[c].p1 / gamma_std(pr[c].p0, &st, (uint32_t)c, 0u); break;
        default: orc_normal_pair(a, b, &z0, &z1); th[c] = orc_exp(pr[c].p0 + pr[c].p1 * z0); break;
        }
    }
}
void orc_prior_rand(int32_t d, const int32_t* kind, const double* par, uint64_t seed, uint32_t particle, double* theta_out) {
    prior1_t pr[16];
    prior_prepare(d, kind, par, pr);
    prior_rand(d, pr, seed, particle, theta_out);
}

/* ------------------------------------------------------------------------------------------ */
/* Models -- the device "f_dist" plugins (contract: src/SimulatedAnnealingABC.jl:421; shapes   */
/* from test/runtests.jl:35,86,128-131,167-170, docs/src/usage.md:16-35, docs/src/example.md).  */
/* ------------------------------------------------------------------------------------------ */
static int model_sim(int32_t id, int32_t d, int32_t s, const double* mp, const double* th,
                     uint64_t seed, uint32_t particle, uint64_t sweep, double* rho) {
    stream_t st = { seed, particle, sweep, KIND_MODEL, 0 };
    uint64_t a, b;
    switch (id) {
    case ORC_MODEL_GAUSS_MEAN: {        /* par: ybar_obs, sd_mean */
        stream_next(&st, &a, &b);
        double z0 = zig_normal(a, &st);
        double ysim = th[0] + mp[1] * z0;
        rho[0] = fabs(ysim - mp[0]);
        return 0; }
    case ORC_MODEL_GAUSS_SAMPLE: {      /* par: n, sigma_fixed, obs1, obs2, stat2_kind */
        int n = (int)mp[0];
        double sig = d >= 2 ? th[1] : mp[1];
        double s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < n; k += 2) {
            double z0, z1;
            normal2(&st, &z0, &z1);
            double y = th[0] + sig * z0;
            s1 = s1 + y; s2 = s2 + y * y;
            if (k + 1 < n) { y = th[0] + sig * z1; s1 = s1 + y; s2 = s2 + y * y; }
        }
        rho[0] = fabs(mp[2] - s1 / (double)n);
        if (s >= 2) rho[1] = fabs(mp[3] - (mp[4] != 0.0 ? s2 : s2 / (double)n));
        return 0; }
    case ORC_MODEL_LOGISTIC: {          /* par: x0, T, obs[T]; theta = (r, K, sigma) */
        int T = (int)mp[1];
        double x = mp[0];
        for (int t = 0; t < T; t += 2) {
            double z[2];
            normal2(&st, &z[0], &z[1]);
            for (int h = 0; h < 2 && t + h < T; ++h) {
                double grow = (th[0] * x) * (1.0 - x / th[1]);
                double noise = (th[2] * x) * z[h];
                x = (x + grow) + noise;
                if (!(x > 0.0)) x = 0.0;
                rho[t + h] = fabs(x - mp[2 + t + h]);
            }
        }
        return 0; }
    case ORC_MODEL_SIR: {               /* par: pop, T, tau, obs_total, obs_peak, obs_tpeak; theta = (β, γ, ι, φ) */
        double pop = mp[0]; int T = (int)mp[1]; double tau = mp[2];
        double inv_pop = 1.0 / pop;
        int64_t I = (int64_t)floor(th[2] * pop + 0.5);
        if (I < 0) I = 0;
        if (I > (int64_t)pop) I = (int64_t)pop;
        int64_t S = (int64_t)pop - I;
        int64_t total = 0, peak = -1, tpeak = 0;
        for (int t = 1; t <= T; ++t) {
            double li = (((th[0] * (double)S) * (double)I) * inv_pop) * tau;
            int64_t ninf = poisson(li, &st); if (ninf > S) ninf = S;
            double lr = (th[1] * (double)I) * tau;
            int64_t nrec = poisson(lr, &st); if (nrec > I) nrec = I;
            S -= ninf; I += ninf - nrec;
            int64_t c = poisson(th[3] * (double)ninf, &st);
            total += c;
            if (c > peak) { peak = c; tpeak = t; }
        }
        double d0 = (double)total - mp[3], d1 = (double)peak - mp[4], d2 = (double)tpeak - mp[5];
        rho[0] = d0 * d0; rho[1] = d1 * d1; rho[2] = d2 * d2;
        return 0; }
    case ORC_MODEL_SIR_GILLESPIE: {     /* docs/src/example.md:75-122,143-147; par: S0, I0, R0, t_max, obs_total, obs_peak, obs_tpeak */
        double S = mp[0], I = mp[1], R = mp[2], tmax = mp[3];
        double Npop = (S + I) + R, t = 0.0, peak = I, tpeak = 0.0;
        for (int ev = 0; ev < 65536 && t < tmax && I > 0.0; ++ev) {           /* while t < t_max && I > 0  (:91) */
            double inf = ((th[0] * S) * I) / Npop;                               /* :93 */
            double rec = th[1] * I;                                              /* :94 */
            double tot = inf + rec;
            stream_next(&st, &a, &b);
            t = t + (-orc_log(u53_open0(a))) / tot;                              /* :98-99 rand(Exponential(1/total_rate)) */
            if (u53(b) < inf / tot) { S -= 1.0; I += 1.0; }                      /* :102-105 */
            else { I -= 1.0; R += 1.0; }                                         /* :107-109 */
            if (I > peak) { peak = I; tpeak = t; }                               /* maximum(sim.I), sim.time[argmax(sim.I)] */
        }
        double d0 = R - mp[4], d1 = peak - mp[5], d2 = tpeak - mp[6];
        d0 = d0 * d0; d1 = d1 * d1; d2 = d2 * d2;                                /* abs2 (:143-147) */
        if (s >= 3) { rho[0] = d0; rho[1] = d1; rho[2] = d2; }
        else rho[0] = (d0 + d1) + d2;                                            /* f_dist_single_stat = sum (:152) */
        return 0; }
    default: return -1;
    }
}
int orc_model_simulate(int32_t id, int32_t d, int32_t s, const double* mp, int32_t nmp, const double* th,
                       uint64_t seed, uint32_t particle, uint64_t sweep, double* rho) {
    (void)nmp;
    return model_sim(id, d, s, mp, th, seed, particle, sweep, rho);
}

/* ------------------------------------------------------------------------------------------ */
/* Proposals  (src/proposals.jl:40-43,52-55 RandomWalk; :101-114 DifferentialEvolution;        */
/* :137-148 StretchMove).  P = inactive half, M rows of d.                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t A, B, C, D; } ctrl_t;
static void ctrl_blocks(uint64_t seed, uint32_t particle, uint64_t sweep, ctrl_t* c) {
    stream_t st = { seed, particle, sweep, KIND_CTRL, 0 };
    stream_block(&st, 0, &c->A, &c->B);
    stream_block(&st, 1, &c->C, &c->D);
}
static void propose(int32_t proposal, const double* pp, int32_t d, const double* th, const double* P, int64_t M,
                    const double* chol, const ctrl_t* cb, uint64_t seed, uint32_t particle, uint64_t sweep,
                    double* out, double* log_factor) {
    if (proposal == ORC_PROP_DE) {
        int64_t i1 = (int64_t)mulhi64(cb->A, (uint64_t)M);               /* :103-107: uniform pair i1 != i2 */
        int64_t i2 = (int64_t)mulhi64(cb->B, (uint64_t)(M - 1));
        if (i2 >= i1) i2++;
        stream_t cs = { seed, particle, sweep, KIND_CTRL, 2 };           /* slow path of the jitter normal: control blocks 2, 3, ... */
        double g = pp[0] * (1.0 + pp[1] * zig_normal(cb->C, &cs));       /* :110 γ0*(1 + σ_γ*randn()) */
        for (int32_t c = 0; c < d; ++c) out[c] = th[c] + g * (P[i1 * d + c] - P[i2 * d + c]);   /* :113 */
        *log_factor = 0.0;
    } else if (proposal == ORC_PROP_STRETCH) {
        int64_t i = (int64_t)mulhi64(cb->A, (uint64_t)M);                /* :141 */
        double t = (pp[0] - 1.0) * u53(cb->B) + 1.0;
        double z = (t * t) / pp[0];                                      /* :144 */
        *log_factor = orc_log(z) * (double)(d - 1);                      /* :146 */
        for (int32_t c = 0; c < d; ++c) out[c] = P[i * d + c] + z * (th[c] - P[i * d + c]);     /* :147 */
    } else {
        stream_t st = { seed, particle, sweep, KIND_RW, 0 };
        double zs[16];
        for (int32_t c = 0; c < d; c += 2) normal2(&st, &zs[c], &zs[c + 1]);
        if (d == 1) out[0] = th[0] + chol[0] * zs[0];                    /* :54 θ + rand(Normal(0, sqrt(Σ))) */
        else for (int32_t r = 0; r < d; ++r) {                           /* :42 θ .+ rand(MvNormal(0, Σ)) = θ + L z */
            double acc = 0.0;
            for (int32_t c = 0; c <= r; ++c) { double t = chol[r * d + c] * zs[c]; acc = (c == 0) ? t : acc + t; }
            out[r] = th[r] + acc;
        }
        *log_factor = 0.0;
    }
}
int orc_propose(int32_t proposal, const double* pp, int32_t d, const double* th, const double* P, int64_t M,
                const double* chol, uint64_t seed, uint32_t particle, uint64_t sweep, double* out, double* lf) {
    ctrl_t cb; ctrl_blocks(seed, particle, sweep, &cb);
    propose(proposal, pp, d, th, P, M, chol, &cb, seed, particle, sweep, out, lf);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Engine                                                                                      */
/* ------------------------------------------------------------------------------------------ */
#define MAXD 16
#define MAXS 32
struct orc_engine {
    orc_config cfg;
    double* model_par; int32_t prior_kind[MAXD]; double prior_par[2 * MAXD]; prior1_t prior[MAXD];
    int64_t N; int32_t d, s, n_eps;
    double *theta, *u, *rho, *lp;            /* row-major per particle: theta[i*d+c], u[i*s+j] */
    double* knots[MAXS]; int64_t L[MAXS];
    double eps[MAXS];
    double chol[MAXD * MAXD];                /* RW: Cholesky factor (d>1) or sd (d==1) */
    int64_t n_simulation, n_accept, n_resampling, n_population_updates;
    double *eps_h, *u_h, *rho_h; int64_t n_rec, cap_rec;
    int initialised;
};

static void push_history(orc_engine* e, const double* um, const double* rm) {
    if (e->n_rec == e->cap_rec) {
        e->cap_rec = e->cap_rec ? 2 * e->cap_rec : 64;
        e->eps_h = realloc(e->eps_h, (size_t)e->cap_rec * e->n_eps * sizeof(double));
        e->u_h = realloc(e->u_h, (size_t)e->cap_rec * e->s * sizeof(double));
        e->rho_h = realloc(e->rho_h, (size_t)e->cap_rec * e->s * sizeof(double));
    }
    memcpy(e->eps_h + e->n_rec * e->n_eps, e->eps, e->n_eps * sizeof(double));
    memcpy(e->u_h + e->n_rec * e->s, um, e->s * sizeof(double));
    memcpy(e->rho_h + e->n_rec * e->s, rm, e->s * sizeof(double));
    e->n_rec++;
}

int orc_create(orc_engine** out, const orc_config* c) {
    if (c->n_para < 1 || c->n_para > MAXD || c->n_stats < 1 || c->n_stats > MAXS) return fail(-1, "bad dimensions");
    if (c->algorithm != ORC_ALG_SINGLE_EPS && c->algorithm != ORC_ALG_MULTI_EPS)
        return fail(-6, "Argument `algorithm` must be :multi_eps or :single_eps");      /* :462-464 */
    if (c->proposal == ORC_PROP_RW && !(c->prop_par[0] > 0.0 && c->prop_par[0] <= 1.0))
        return fail(-7, "Mixing parameter `β` must be between zero and one.");          /* proposals.jl:30 */
    orc_engine* e = calloc(1, sizeof *e);
    e->cfg = *c; e->N = c->n_particles; e->d = c->n_para; e->s = c->n_stats;
    e->n_eps = c->algorithm == ORC_ALG_MULTI_EPS ? c->n_stats : 1;
    e->model_par = malloc((size_t)(c->n_model_par > 0 ? c->n_model_par : 1) * sizeof(double));
    memcpy(e->model_par, c->model_par, (size_t)c->n_model_par * sizeof(double));
    memcpy(e->prior_kind, c->prior_kind, (size_t)e->d * sizeof(int32_t));
    memcpy(e->prior_par, c->prior_par, (size_t)2 * e->d * sizeof(double));
    prior_prepare(e->d, e->prior_kind, e->prior_par, e->prior);
    e->theta = malloc((size_t)e->N * e->d * sizeof(double));
    e->u = malloc((size_t)e->N * e->s * sizeof(double));
    e->rho = malloc((size_t)e->N * e->s * sizeof(double));
    e->lp = malloc((size_t)e->N * sizeof(double));
    *out = e;
    return 0;
}
int orc_destroy(orc_engine* e) {
    if (!e) return 0;
    free(e->model_par); free(e->theta); free(e->u); free(e->rho); free(e->lp);
    for (int j = 0; j < MAXS; ++j) free(e->knots[j]);
    free(e->eps_h); free(e->u_h); free(e->rho_h); free(e);
    return 0;
}

/* exact per-statistic means of u over the population */
static void mean_u_cols(const orc_engine* e, double* um, double* grand) {
    uint64_t GH = 0, GL = 0;
    for (int32_t j = 0; j < e->s; ++j) {
        uint64_t H = 0, L = 0;
        #pragma omp parallel for schedule(static) reduction(+ : H, L)
        for (int64_t i = 0; i < e->N; ++i) { uint64_t h, l; u_limbs(e->u[i * e->s + j], &h, &l); H += h; L += l; }
        um[j] = limbs_to_sum(H, L) / (double)e->N;
        GH += H; GL += L;
    }
    if (grand) *grand = limbs_to_sum(GH, GL) / (double)(e->N * e->s);   /* mean(u) over all N*s entries (:203,353) */
}
static double col_treesum(const double* a, int64_t lo, int64_t hi, int32_t ld, int32_t j) {
    int64_t n = hi - lo;
    double* tmp = malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    for (int64_t i = 0; i < n; ++i) tmp[i] = a[(lo + i) * ld + j];
    double r = orc_treesum(tmp, n);
    free(tmp);
    return r;
}

/* resample_population(population, u, δ)  (:124-137); ρ is NOT resampled (:197,341) */
static void resample(orc_engine* e) {
    int64_t N = e->N; int32_t s = e->s, d = e->d;
    double um[MAXS]; mean_u_cols(e, um, NULL);                            /* :126 */
    uint64_t* q = malloc((size_t)N * sizeof(uint64_t));
    int64_t* idx = malloc((size_t)N * sizeof(int64_t));
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        double acc = 0.0;
        for (int32_t j = 0; j < s; ++j) { double t = (e->u[i * s + j] * e->cfg.delta) / um[j]; acc = (j == 0) ? t : acc + t; }
        q[i] = (uint64_t)(orc_exp(-acc) * 4294967296.0);                  /* :127 */
    }
    orc_resample_indices(q, N, e->cfg.seed, (uint64_t)e->n_resampling, idx);   /* :129 */
    double* th2 = malloc((size_t)N * d * sizeof(double)); double* u2 = malloc((size_t)N * s * sizeof(double));
    double* lp2 = malloc((size_t)N * sizeof(double));
    #pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < N; ++k) {                                      /* :131-132 */
        memcpy(th2 + k * d, e->theta + idx[k] * d, d * sizeof(double));
        memcpy(u2 + k * s, e->u + idx[k] * s, s * sizeof(double));
        lp2[k] = e->lp[idx[k]];
    }
    free(e->theta); free(e->u); free(e->lp); e->theta = th2; e->u = u2; e->lp = lp2;
    free(q); free(idx);
    e->n_resampling += 1;
}

static int update_eps(orc_engine* e) {
    double um[MAXS], grand; mean_u_cols(e, um, &grand);
    if (e->cfg.algorithm == ORC_ALG_MULTI_EPS) return orc_eps_multi(um, e->s, e->cfg.v, e->eps);   /* :200-201,350-351 */
    e->eps[0] = orc_eps_single(grand, e->cfg.v);                                                  /* :202-203,352-353 */
    return 0;
}

/* update_proposal!  (src/proposals.jl:46-48 n-D, :58-60 1-D; no-op for DE/Stretch :116,150) */
static void update_proposal(orc_engine* e) {
    if (e->cfg.proposal != ORC_PROP_RW) return;
    int64_t N = e->N; int32_t d = e->d;
    double mean[MAXD], cov[MAXD * MAXD];
    double* tmp = malloc((size_t)N * sizeof(double));
    for (int32_t c = 0; c < d; ++c) mean[c] = col_treesum(e->theta, 0, N, d, c) / (double)N;
    for (int32_t a = 0; a < d; ++a) for (int32_t b = 0; b <= a; ++b) {
        for (int64_t i = 0; i < N; ++i) tmp[i] = (e->theta[i * d + a] - mean[a]) * (e->theta[i * d + b] - mean[b]);
        cov[a * d + b] = cov[b * d + a] = orc_treesum(tmp, N) / (double)(N - 1);        /* corrected cov */
    }
    free(tmp);
    double beta = e->cfg.prop_par[0];
    if (d == 1) { e->chol[0] = sqrt(beta * cov[0]); return; }                             /* :59, :54 */
    double Sg[MAXD * MAXD];
    for (int32_t a = 0; a < d; ++a) for (int32_t b = 0; b < d; ++b)
        Sg[a * d + b] = beta * (a == b ? cov[a * d + b] + 1e-8 : cov[a * d + b]);         /* :47 */
    for (int32_t r = 0; r < d; ++r) for (int32_t c = 0; c <= r; ++c) {                    /* Cholesky, row by row */
        double sum = Sg[r * d + c];
        for (int32_t k = 0; k < c; ++k) sum = sum - e->chol[r * d + k] * e->chol[c * d + k];
        e->chol[r * d + c] = (r == c) ? sqrt(sum) : sum / e->chol[c * d + c];
    }
}

/* initialization()  (src/SimulatedAnnealingABC.jl:151-227) */
int orc_init(orc_engine* e) {
    int64_t N = e->N; int32_t d = e->d, s = e->s;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {                                                       /* :172-179 */
        prior_rand(d, e->prior, e->cfg.seed, (uint32_t)i, e->theta + i * d);
        e->lp[i] = prior_logpdf(d, e->prior, e->theta + i * d);
        model_sim(e->cfg.model_id, d, s, e->model_par, e->theta + i * d, e->cfg.seed, (uint32_t)i, 0, e->rho + i * s);
    }
    double rm[MAXS], um[MAXS];
    for (int32_t j = 0; j < s; ++j) rm[j] = col_treesum(e->rho, 0, N, s, j) / (double)N;   /* :180 */
    for (int64_t i = 0; i < N * s; ++i) if (e->rho[i] < 0.0) return fail(-4, "Negative distances are not allowed!"); /* :185 */
    double* col = malloc((size_t)N * sizeof(double));
    for (int32_t j = 0; j < s; ++j) {                                                       /* :187 build_cdf */
        for (int64_t i = 0; i < N; ++i) col[i] = e->rho[i * s + j];
        free(e->knots[j]); e->knots[j] = malloc((size_t)(N + 2) * sizeof(double));
        e->L[j] = orc_ecdf_build(col, N, e->knots[j]);
        if (e->L[j] < 0) { free(col); return fail(-8, "build_cdf: no positive prior distance for a statistic"); }
        int64_t K = e->cfg.ecdf_max_knots, m = e->L[j] - 2;
        if (K >= 2 && m > K) {                 /* compressed mode: x[floor(i (m-1)/(K-1))], i = 0..K-1 */
            double* kn = e->knots[j];
            double* small = malloc((size_t)(K + 2) * sizeof(double));
            small[0] = 0.0;
            for (int64_t i = 0; i < K; ++i) small[1 + i] = kn[1 + (i * (m - 1)) / (K - 1)];
            small[K + 1] = kn[m] * 1.5;
            free(kn); e->knots[j] = small; e->L[j] = K + 2;
        }
    }
    free(col);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)                                                         /* :190-192 */
        for (int32_t j = 0; j < s; ++j) e->u[i * s + j] = ecdf_eval1(e->knots[j], e->L[j], e->rho[i * s + j]);
    e->n_resampling = 0;
    resample(e);                                                                            /* :197 -> n_resampling = 1 */
    int rc = update_eps(e); if (rc) return rc;                                              /* :200-204 */
    mean_u_cols(e, um, NULL);
    e->n_rec = 0;
    push_history(e, um, rm);                                                                /* :180,207-208 */
    e->n_simulation = N; e->n_accept = 0; e->n_population_updates = 0;                      /* :213-223 */
    e->initialised = 1;
    return 0;
}

/* update_population!()  (src/SimulatedAnnealingABC.jl:251-402) */
int orc_update(orc_engine* e, int64_t n_simulation, int64_t checkpoint_history, double* seconds_out) {
    if (!(e->cfg.v > 0.0)) return fail(-2, "Annealing speed `v` must be positive.");      /* :261 */
    if (!(e->cfg.delta > 0.0)) return fail(-3, "Resamping intensity `δ` must be positive."); /* :262 */
    int64_t N = e->N; int32_t d = e->d, s = e->s;
    int64_t n_pop = n_simulation / N;                                                       /* :275 */
    int64_t last_cp = 0;
    if (checkpoint_history < 1) checkpoint_history = 1;
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    update_proposal(e);                                                                     /* :284 */
    double um[MAXS], rm[MAXS];
    int64_t h0 = N / 2;                                                                     /* :300-301 */
    for (int64_t ix = 1; ix <= n_pop; ++ix) {
        uint64_t t = (uint64_t)(e->n_population_updates + ix);
        int64_t nacc = 0;
        double rsum[2][MAXS];
        for (int half = 0; half < 2; ++half) {                                              /* :304 */
            int64_t a0 = half == 0 ? 0 : h0, a1 = half == 0 ? h0 : N;
            int64_t i0 = half == 0 ? h0 : 0, M = half == 0 ? N - h0 : h0;
            const double* P = e->theta + i0 * d;
            uint64_t sweep = 2 * t + (uint64_t)half;
            #pragma omp parallel for schedule(dynamic, 256) reduction(+ : nacc)
            for (int64_t i = a0; i < a1; ++i) {                                             /* :308-331 */
                ctrl_t cb; ctrl_blocks(e->cfg.seed, (uint32_t)i, sweep, &cb);
                double thp[MAXD], rp[MAXS], up[MAXS], lf, Lacc;
                propose(e->cfg.proposal, e->cfg.prop_par, d, e->theta + i * d, P, M, e->chol, &cb,
                        e->cfg.seed, (uint32_t)i, sweep, thp, &lf);                        /* :311 */
                double lpp = prior_logpdf(d, e->prior, thp);
                int acc;
                if (lpp > -INFINITY) {                                                      /* :314 */
                    model_sim(e->cfg.model_id, d, s, e->model_par, thp, e->cfg.seed, (uint32_t)i, sweep, rp); /* :315 */
                    for (int32_t j = 0; j < s; ++j) up[j] = ecdf_eval1(e->knots[j], e->L[j], rp[j]);          /* :316 */
                    acc = accept1(s, e->u + i * s, 1, up, 1, e->eps, e->n_eps, lpp - e->lp[i], lf, u53(cb.D)); /* :318-324 */
                } else {
                    Lacc = -INFINITY;                                                       /* :320-322 */
                    acc = orc_log(u53(cb.D)) < Lacc;
                }
                if (acc) {                                                                  /* :325-328 */
                    memcpy(e->theta + i * d, thp, d * sizeof(double));
                    memcpy(e->u + i * s, up, s * sizeof(double));
                    memcpy(e->rho + i * s, rp, s * sizeof(double));
                    e->lp[i] = lpp;
                    nacc += 1;
                }
            }
            for (int32_t j = 0; j < s; ++j) rsum[half][j] = col_treesum(e->rho, a0, a1, s, j);
        }
        e->n_accept += nacc;                                                                /* :334 */
        if (e->n_accept >= (e->n_resampling + 1) * e->cfg.resample) resample(e);            /* :340-343 */
        update_proposal(e);                                                                 /* :348 */
        int rc = update_eps(e); if (rc) return rc;                                          /* :350-354 */
        if (ix % checkpoint_history == 0) {                                                 /* :367-372 */
            mean_u_cols(e, um, NULL);
            for (int32_t j = 0; j < s; ++j) rm[j] = (rsum[0][j] + rsum[1][j]) / (double)N;
            push_history(e, um, rm);
            last_cp = ix;
        }
    }
    if (last_cp != n_pop) {                                                                 /* :378-382 */
        mean_u_cols(e, um, NULL);
        for (int32_t j = 0; j < s; ++j) rm[j] = (col_treesum(e->rho, 0, h0, s, j) + col_treesum(e->rho, h0, N, s, j)) / (double)N;
        push_history(e, um, rm);
    }
    e->n_simulation += n_pop * N;                                                           /* :391 */
    e->n_population_updates += n_pop;                                                       /* :394 */
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds_out) *seconds_out = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return 0;
}

int orc_get_population(orc_engine* e, double* theta, double* u, double* rho) {
    int64_t N = e->N;
    for (int64_t i = 0; i < N; ++i) {
        if (theta) for (int32_t c = 0; c < e->d; ++c) theta[c * N + i] = e->theta[i * e->d + c];
        if (u) for (int32_t j = 0; j < e->s; ++j) u[j * N + i] = e->u[i * e->s + j];
        if (rho) for (int32_t j = 0; j < e->s; ++j) rho[j * N + i] = e->rho[i * e->s + j];
    }
    return 0;
}
int orc_set_population(orc_engine* e, const double* theta, const double* u, const double* rho,
                       const double* eps, const int64_t counters[4]) {
    int64_t N = e->N;
    for (int64_t i = 0; i < N; ++i) {
        for (int32_t c = 0; c < e->d; ++c) e->theta[i * e->d + c] = theta[c * N + i];
        for (int32_t j = 0; j < e->s; ++j) { e->u[i * e->s + j] = u[j * N + i]; e->rho[i * e->s + j] = rho[j * N + i]; }
        e->lp[i] = prior_logpdf(e->d, e->prior, e->theta + i * e->d);
    }
    memcpy(e->eps, eps, e->n_eps * sizeof(double));
    e->n_simulation = counters[0]; e->n_accept = counters[1]; e->n_resampling = counters[2]; e->n_population_updates = counters[3];
    return 0;
}
int orc_get_state(orc_engine* e, double* eps, int64_t counters[4]) {
    if (eps) memcpy(eps, e->eps, e->n_eps * sizeof(double));
    if (counters) { counters[0] = e->n_simulation; counters[1] = e->n_accept; counters[2] = e->n_resampling; counters[3] = e->n_population_updates; }
    return 0;
}
int64_t orc_history_len(orc_engine* e) { return e->n_rec; }
int orc_get_history(orc_engine* e, double* eps_h, double* u_h, double* rho_h) {
    if (eps_h) memcpy(eps_h, e->eps_h, (size_t)e->n_rec * e->n_eps * sizeof(double));
    if (u_h) memcpy(u_h, e->u_h, (size_t)e->n_rec * e->s * sizeof(double));
    if (rho_h) memcpy(rho_h, e->rho_h, (size_t)e->n_rec * e->s * sizeof(double));
    return 0;
}
int64_t orc_get_ecdf(orc_engine* e, int32_t stat, double* knots_out) {
    if (stat < 0 || stat >= e->s || !e->knots[stat]) return -1;
    if (knots_out) memcpy(knots_out, e->knots[stat], (size_t)e->L[stat] * sizeof(double));
    return e->L[stat];
}
int orc_set_ecdf(orc_engine* e, int32_t stat, const double* knots, int64_t L) {
    if (stat < 0 || stat >= e->s || L < 3) return fail(-1, "bad ecdf");
    free(e->knots[stat]); e->knots[stat] = malloc((size_t)L * sizeof(double));
    memcpy(e->knots[stat], knots, (size_t)L * sizeof(double)); e->L[stat] = L;
    return 0;
}
