// detmath.cuh -- deterministic FP64 elementary functions for host and device.
//
// The SABC accept test compares log(U) with a sum of u-differences (reference
// src/SimulatedAnnealingABC.jl:318-324) and the model plugins draw normals / Poisson counts.
// libm, Julia and libdevice each round log/exp/sin/cos differently in the last ulp, so the
// product carries its own implementations built only from + - * / sqrt fma and integer ops;
// compiled with -fmad=false they give the same bits on sm_100a and on any IEEE-754 host.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define SABC_HD __host__ __device__ __forceinline__
#define SABC_D __device__ __forceinline__
#else
#define SABC_HD inline
#define SABC_D inline
#endif

namespace sabc {

SABC_HD uint64_t f64_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; __builtin_memcpy(&u, &x, 8); return u;
#endif
}
SABC_HD double bits_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; __builtin_memcpy(&x, &u, 8); return x;
#endif
}
SABC_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
SABC_HD double dinf() { return bits_f64(0x7ff0000000000000ULL); }
SABC_HD double dnan() { return bits_f64(0x7ff8000000000000ULL); }

// natural logarithm, <= 1 ulp: x = 2^k (1+f), log(1+f) by the atanh-type series in s = f/(2+f)
SABC_HD double det_log(double x) {
    const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33;
    if (x != x) return x;
    if (x < 0.0) return dnan();
    if (x == 0.0) return -dinf();
    if (x == dinf()) return x;
    int k = 0;
    uint64_t b = f64_bits(x);
    if (b < 0x0010000000000000ULL) { x = x * 0x1p54; k = -54; b = f64_bits(x); }
    uint32_t top = (uint32_t)(b >> 32);
    k += (int)(top >> 20) - 1023;
    top &= 0x000fffffu;
    const uint32_t bump = (top + 0x95f64u) & 0x100000u;       // mantissa above sqrt(2): halve it
    k += (int)(bump >> 20);
    b = ((uint64_t)(top | (bump ^ 0x3ff00000u)) << 32) | (b & 0xffffffffULL);
    const double f = bits_f64(b) - 1.0;
    const double s = f / (2.0 + f);
    const double dk = (double)k;
    const double z = s * s, w = z * z;
    const double even = w * dfma(w, dfma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double odd = z * dfma(w, dfma(w, dfma(w, 1.479819860511658591e-01, 1.818357216161805012e-01),
                                       2.857142874366239149e-01), 6.666666666666735130e-01);
    const double R = odd + even;
    const double hfsq = (0.5 * f) * f;
    const double corr = dfma(s, hfsq + R, dk * ln2_lo);
    return dk * ln2_hi - ((hfsq - corr) - f);
}

// det_exp's constants with a non-zero low mantissa word: from the constant bank on the device (one LDCU.64 / half an
// LDCU.128 each instead of two UMOV per use); the same literals on the host
#if defined(__CUDACC__)
__constant__ double c_exp[8] = {0x1.62e42fee00000p-1, 0x1.a39ef35793c76p-33, 0x1.71547652b82fep+0, 4.13813679705723846039e-08,
                                -1.65339022054652515390e-06, 6.61375632143793436117e-05, -2.77777777770155933842e-03,
                                1.66666666666666019037e-01};
#endif
#if defined(__CUDA_ARCH__) && !defined(SABC_NO_CONST_BANK)
#define SABC_EC(i, v) c_exp[i]
#else
#define SABC_EC(i, v) (v)
#endif

// exponential, <= 1 ulp: x = k ln2 + r, exp(r) = 1 + r + r c/(2-c)
SABC_HD double det_exp(double x) {
    const double ln2_hi = SABC_EC(0, 0x1.62e42fee00000p-1), ln2_lo = SABC_EC(1, 0x1.a39ef35793c76p-33), inv_ln2 = SABC_EC(2, 0x1.71547652b82fep+0);
    if (x != x) return x;
    if (x > 709.782712893383973096) return dinf();
    if (x < -745.13321910194110842) return 0.0;
    const double kf = floor(dfma(x, inv_ln2, 0.5));
    const int k = (int)kf;
    const double hi = dfma(-kf, ln2_hi, x);
    const double lo = kf * ln2_lo;
    const double r = hi - lo;
    const double t = r * r;
    const double poly = dfma(t, dfma(t, dfma(t, dfma(t, SABC_EC(3, 4.13813679705723846039e-08), SABC_EC(4, -1.65339022054652515390e-06)),
                                            SABC_EC(5, 6.61375632143793436117e-05)), SABC_EC(6, -2.77777777770155933842e-03)),
                             SABC_EC(7, 1.66666666666666019037e-01));
    const double c = r - t * poly;
    const double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    if (k >= -1021 && k <= 1023) return bits_f64(f64_bits(y) + ((uint64_t)(int64_t)k << 52));
    if (k > 1023) return y * 0x1p1023 * bits_f64((uint64_t)k << 52);
    return bits_f64(f64_bits(y) + ((uint64_t)(int64_t)(k + 1000) << 52)) * 0x1p-1000;
}

// sin(2 pi u), cos(2 pi u), u in [0,1): quadrant reduction is exact, then Taylor in r on [-1/4,1/4]
SABC_HD void det_sincos2pi(double u, double& sn, double& cs) {
    const double t = 2.0 * u;
    const double qf = floor(dfma(2.0, t, 0.5));
    const int q = (int)qf;
    const double r = t - 0.5 * qf;
    const double r2 = r * r;
    double ps = -0x1.8a404211f9547p-26, pc = -0x1.2a0c591af8314p-23;
    ps = dfma(ps, r2, 0x1.aaec32af93359p-21);   pc = dfma(pc, r2, 0x1.20c62c2f2d7f5p-18);
    ps = dfma(ps, r2, -0x1.6fadb9f155744p-16);  pc = dfma(pc, r2, -0x1.b6e24f44b128fp-14);
    ps = dfma(ps, r2, 0x1.e8f434d018d63p-12);   pc = dfma(pc, r2, 0x1.f9d38a3763cc3p-10);
    ps = dfma(ps, r2, -0x1.e3074fde8871fp-8);   pc = dfma(pc, r2, -0x1.a6d1f2a204a8cp-6);
    ps = dfma(ps, r2, 0x1.50783487ee782p-4);    pc = dfma(pc, r2, 0x1.e1f506891babbp-3);
    ps = dfma(ps, r2, -0x1.32d2cce62bd86p-1);   pc = dfma(pc, r2, -0x1.55d3c7e3cbffap+0);
    ps = dfma(ps, r2, 0x1.466bc6775aae2p+1);    pc = dfma(pc, r2, 0x1.03c1f081b5ac4p+2);
    ps = dfma(ps, r2, -0x1.4abbce625be53p+2);   pc = dfma(pc, r2, -0x1.3bd3cc9be45dep+2);
    ps = dfma(ps, r2, 0x1.921fb54442d18p+1);    pc = dfma(pc, r2, 1.0);
    const double sr = r * ps, cr = pc;
    const int qq = q & 3;
    sn = (qq == 0) ? sr : (qq == 1) ? cr : (qq == 2) ? -sr : -cr;
    cs = (qq == 0) ? cr : (qq == 1) ? -sr : (qq == 2) ? -cr : sr;
}

// cos(2 pi u) only
SABC_HD double det_cos2pi(double u) {
    double s, c; det_sincos2pi(u, s, c); return c;
}

// log(k!) for integer-valued k >= 0
SABC_HD double det_logfact(double k) {
    if (k <= 16.0) {
        switch ((int)k) {
            case 0: case 1: return 0.0;
            case 2: return 0x1.62e42fefa39efp-1;   case 3: return 0x1.cab0bfa2a2002p+0;
            case 4: return 0x1.96ca77c922cf9p+1;   case 5: return 0x1.326643c4479c9p+2;
            case 6: return 0x1.a51273acf01cap+2;   case 7: return 0x1.10ce1f32dcc30p+3;
            case 8: return 0x1.5358e82fcb70dp+3;   case 9: return 0x1.99a8921a7f7cfp+3;
            case 10: return 0x1.e357590954d15p+3;  case 11: return 0x1.180973f3a8d74p+4;
            case 12: return 0x1.3fcba16d50143p+4;  case 13: return 0x1.68d5a9c3b32cep+4;
            case 14: return 0x1.930f3df162a42p+4;  case 15: return 0x1.be636a63fd346p+4;
            default: return 0x1.eabff061f1a84p+4;
        }
    }
    const double x = k + 1.0;
    const double lx = det_log(x);
    const double r = 1.0 / x, r2 = r * r;
    double p = dfma(-r2, 1.0 / 1680.0, 1.0 / 1260.0);
    p = dfma(-r2, p, 1.0 / 360.0);
    p = dfma(-r2, p, 1.0 / 12.0);
    double t = (x - 0.5) * lx;
    t = t - x;
    t = t + 0x1.d67f1c864beb5p-1;
    return t + r * p;
}

// log Gamma(x) for x > 0: argument shifted to x >= 16 by the recurrence, then Stirling's series through 1/(1188 x^9)
// (truncation < 1e-16); the normalisation constants of the Gamma and Beta priors
SABC_HD double det_lgamma(double x) {
    if (!(x > 0.0)) return dinf();
    double prod = 1.0;
    while (x < 16.0) { prod = prod * x; x = x + 1.0; }
    const double lx = det_log(x);
    const double r = 1.0 / x, r2 = r * r;
    double p = dfma(-r2, 1.0 / 1188.0, 1.0 / 1680.0);
    p = dfma(-r2, p, 1.0 / 1260.0);
    p = dfma(-r2, p, 1.0 / 360.0);
    p = dfma(-r2, p, 1.0 / 12.0);
    double t = (x - 0.5) * lx;
    t = t - x;
    t = t + 0x1.d67f1c864beb5p-1;
    t = t + r * p;
    return t - det_log(prod);
}

}  // namespace sabc
