// ptrs2_experimental.cuh -- NOT PART OF THE PRODUCT BUILD.  Compiled only with -DSABC_EXPERIMENTAL_PTRS2 (see
// tools/exp_ptrs2.sh); written at the end of round 1 when no GPU time was left, so it is compile-checked and its decision
// logic is checked by the CPU emulation tools/ptrs_candidate_study.cpp, but it has never run on a GPU.  Before it may be
// enabled it has to pass sabc_ptrs2_check (hooks.cu, same macro) on >= 1e10 attempts and the whole GPU parity suite.
//
// Two changes to the PTRS attempt of poisson_attempt_d, both in the spirit of the acceptance filters: only DECISIONS are
// approximated, each against a propagated error bound, and every undecided attempt takes today's exact path, so no draw changes.
//  (1) ptrs_candidate_mufu: sqrt(lam), 1/us, b, a and (2a/us + b) U in FP32 on the MUFU unit, lam + 0.43 added in FP64;
//      floor() is trusted when the fractional part is farther than 1e-6 |t| + 1e-9 from an integer, the squeeze when it clears
//      6e-7 b; num and den (FP32) go to the MUFU acceptance filter, whose bound absorbs their 1.5e-6 relative perturbation.
//  (2) ptrs_filter_mufu2: k ln(1 + D/lam) by the series of log1p for |D/lam| < 1/16 instead of MUFU.LG2, which removes the
//      k * 1e-6 term of the bound (E = 3e-4 + 1.5e-6 |D|): at lam = 25000 the bound drops from 2.5e-2 to 7e-4, so the FP64
//      filter runs for ~0.1 % instead of ~4 % of the candidates that reach the acceptance test.
#pragma once

namespace sabc {

SABC_D float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// acceptance filter on FP32 inputs: +1 accept, -1 reject, 0 undecided
SABC_D int ptrs_filter_mufu2(double lam, double kf, float numf, float denf, float& T, float& E) {
    T = 0.0f; E = 0.0f;
    if (!(kf >= 2.0) || !(kf < 1e7) || !(numf > 0x1p-100f) || !(denf > 0x1p-100f) || !(denf < 0x1p100f)) return 0;
    const double x = kf + 1.0;
    const double D = x - lam;
    const float xf = __double2float_rn(x), kff = __double2float_rn(kf);
    const float dl = __fmul_rn(__double2float_rn(D), mufu_rcp(__double2float_rn(lam)));      // delta = D / lam, relative error < 2e-7
    const float q = __fadd_rn(1.0f, dl);
    if (!(q >= 0.5f) || !(q <= 256.0f)) return 0;
    const float rx = mufu_rcp(xf);
    const float corr = __fmul_rn(rx, __fmaf_rn(__fmul_rn(rx, rx), -1.0f / 360.0f, 1.0f / 12.0f));
    const bool small = fabsf(dl) < 0.0625f;
    float klnq;                                                                               // k ln(q)
    if (small) {
        // log1p(d) = d (1 - d/2 + d^2/3 - d^3/4 + d^4/5 - d^5/6), truncation < d^7/7 <= 5.3e-10
        float p = __fmaf_rn(dl, -1.0f / 6.0f, 0.2f);
        p = __fmaf_rn(dl, p, -0.25f);
        p = __fmaf_rn(dl, p, 1.0f / 3.0f);
        p = __fmaf_rn(dl, p, -0.5f);
        p = __fmaf_rn(dl, p, 1.0f);
        klnq = __fmul_rn(kff, __fmul_rn(dl, p));
    } else {
        klnq = __fmul_rn(kff, __fmul_rn(0.693147180559945f, mufu_lg2(q)));
    }
    const float rest = __fmul_rn(0.693147180559945f, __fadd_rn(__fmul_rn(0.5f, mufu_lg2(xf)), __fsub_rn(mufu_lg2(numf), mufu_lg2(denf))));
    const float Df = __double2float_rn(D - c_ptrs[12]);
    T = __fsub_rn(__fsub_rn(__fsub_rn(Df, klnq), corr), rest);
    // small: k |delta| (2e-7 + 3 ulp) <= 1.1 |D| * 4e-7, FP32 sums of two terms of size |D|: 2e-7 |D|  ->  1.5e-6 |D| with margin
    // else : the bound of ptrs_filter_mufu; both + 1e-4 for the FP32 num, den of the approximate candidate
    E = small ? __fmaf_rn(fabsf(Df), 1.5e-6f, 3e-4f)
              : __fmaf_rn(kff, q <= 2.0f ? 1e-6f : 3e-6f, __fmaf_rn(fabsf(Df), 5e-7f, 3e-4f));
    return T > E ? 1 : (T < -E ? -1 : 0);
}

// 1 = accept kf, 0 = reject, 2 = acceptance test on (kf, numf, denf), 3 = undecided: run the exact candidate
SABC_D int ptrs_candidate_mufu(double lam, const U64x2 w, double& kf, float& numf, float& denf) {
    const double U = u53(w.a) - 0.5, V = u53(w.b);
    const double us = 0.5 - fabs(U);                                     // exact
    const float slam = mufu_sqrt(__double2float_rn(lam));
    const float b = __fmaf_rn(2.53f, slam, 0.931f);
    const float a = __fmaf_rn(0.02483f, b, -0.059f);
    const float r = mufu_rcp(__double2float_rn(us));
    const float g = __fmaf_rn(__fadd_rn(a, a), r, b);
    const float t = __fmul_rn(g, __double2float_rn(U));
    const double Et = fma(1e-6, fabs((double)t), 1e-9);
    const double arg = (double)t + (lam + c_ptrs[4]);
    kf = floor(arg);
    const double frac = arg - kf;
    const bool kf_sure = frac > Et && frac < 1.0 - Et;                  // false for NaN / Inf as well
    const double bd = (double)b;
    const double sq = (c_ptrs[5] - V) * (bd - 2.0), Esq = fma(6e-7, bd, 1e-9);
    if (us >= c_ptrs[7]) {
        if (sq >= c_ptrs[6] + Esq) return kf_sure ? 1 : 3;
        if (sq > c_ptrs[6] - Esq) return 3;
    }
    if (!kf_sure) return 3;
    if (kf < 0.0 || (us < c_ptrs[8] && V > us)) return 0;
    const float bm = __fsub_rn(b, 3.4f);
    numf = __fmul_rn(__double2float_rn(V), __fmaf_rn(1.1239f, bm, 1.1328f));
    denf = __fmul_rn(bm, __fmaf_rn(__fmul_rn(a, r), r, b));
    return 2;
}

// the experimental attempt: 1 accept (kf valid), 0 reject, 3 = fall back to the exact attempt
SABC_D int ptrs_attempt2(double lam, const U64x2 w, double& kf) {
    float numf = 0.0f, denf = 0.0f;
    int s = ptrs_candidate_mufu(lam, w, kf, numf, denf);
    if (s == 2) {
        float T, E;
        const int dec = ptrs_filter_mufu2(lam, kf, numf, denf, T, E);
        s = dec > 0 ? 1 : (dec < 0 ? 0 : 3);
    }
    return s;
}

}  // namespace sabc
