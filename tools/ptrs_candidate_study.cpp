// ptrs_candidate_study.cpp -- CPU study for the next step of the SIR kernel (DEVELOPMENT.md): an approximate-first PTRS
// candidate.  The exact candidate (csrc/philox.cuh, ptrs_candidate) computes sqrt(lam), 1/us and ~25 FP64 operations per
// attempt; all of it feeds DECISIONS (the squeeze, the quick reject, floor of the candidate value).  Here the same quantities
// are computed in single precision with the special-function results perturbed by a relative error of up to 2^-22 (random
// sign and size: a stand-in for MUFU.SQRT / MUFU.RCP, which cannot be reproduced on a CPU), and a decision is taken only when
// it clears a propagated error bound; otherwise the attempt is "undecided" and would take the exact path on the device.
// The FP32 acceptance filter with the log1p series (ptrs_filter_mufu2 of csrc/ptrs2_experimental.cuh) is emulated the same way.
// The program counts undecided attempts and -- the point of the study -- final decisions that differ from the spec'd attempt.
//   g++ -O2 -std=c++17 -ffp-contract=off -mfma -fopenmp tools/ptrs_candidate_study.cpp -o /tmp/ptrs_study && /tmp/ptrs_study [attempts per lambda] [error mode 0..4]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../simulatedannealingabc.jl_b200/csrc/philox.cuh"
using namespace sabc;

static int g_mode = 0;   // 0: random error in [-2^-22, 2^-22]; 1..4: always at an extreme, the four sign combinations
static inline float perturb(float x, uint32_t bits, int which) {   // x (1 + d), |d| <= 2^-22
    float d = ((float)(bits & 0xffffu) * (1.0f / 65536.0f) * 2.0f - 1.0f) * 0x1p-22f;
    if (g_mode) d = (((g_mode - 1) >> which) & 1) ? 0x1p-22f : -0x1p-22f;
    return x * (1.0f + d);
}

// MUFU.LG2 stand-in: absolute error 2^-21.41 on [0.5, 2], 2 ulp elsewhere (the documented bounds), at an extreme or random
static inline float lg2_emul(float x, uint32_t bits, int which) {
    const double v = log2((double)x);
    double e = (x >= 0.5f && x <= 2.0f) ? 3.6e-7 : fmax(3.6e-7, 2.4e-7 * fabs(v));
    double d = ((double)(bits & 0xffffu) / 65536.0 * 2.0 - 1.0) * e;
    if (g_mode) d = (((g_mode - 1) >> which) & 1) ? e : -e;
    return (float)(v + d);
}

// emulation of ptrs_filter_mufu2 (csrc/ptrs2_experimental.cuh): +1 / -1 / 0
static int filter2_emul(double lam, double kf, float numf, float denf, uint32_t noise) {
    if (!(kf >= 2.0) || !(kf < 1e7) || !(numf > 0x1p-100f) || !(denf > 0x1p-100f) || !(denf < 0x1p100f)) return 0;
    const double x = kf + 1.0, D = x - lam;
    const float xf = (float)x, kff = (float)kf;
    const float dl = (float)D * perturb(1.0f / (float)lam, noise, 1);
    const float q = 1.0f + dl;
    if (!(q >= 0.5f) || !(q <= 256.0f)) return 0;
    const float rx = perturb(1.0f / xf, noise >> 3, 0);
    const float corr = rx * fmaf(rx * rx, -1.0f / 360.0f, 1.0f / 12.0f);
    const bool small = fabsf(dl) < 0.0625f;
    float klnq;
    if (small) {
        float p = fmaf(dl, -1.0f / 6.0f, 0.2f);
        p = fmaf(dl, p, -0.25f); p = fmaf(dl, p, 1.0f / 3.0f); p = fmaf(dl, p, -0.5f); p = fmaf(dl, p, 1.0f);
        klnq = kff * (dl * p);
    } else {
        klnq = kff * (0.693147180559945f * lg2_emul(q, noise >> 5, 0));
    }
    const float rest = 0.693147180559945f * ((0.5f * lg2_emul(xf, noise >> 7, 1)) + (lg2_emul(numf, noise >> 9, 0) - lg2_emul(denf, noise >> 11, 1)));
    const float Df = (float)(D - 0x1.d67f1c864beb5p-1);
    const float T = ((Df - klnq) - corr) - rest;
    const float E = small ? fmaf(fabsf(Df), 1.5e-6f, 3e-4f) : fmaf(kff, q <= 2.0f ? 1e-6f : 3e-6f, fmaf(fabsf(Df), 5e-7f, 3e-4f));
    return T > E ? 1 : (T < -E ? -1 : 0);
}

// returns 1 accept kf, 0 reject, 2 exact test needed (kf valid), 3 undecided (take the exact candidate)
static int approx_candidate(double lam, const U64x2 w, double& kf, uint32_t noise, float& numf, float& denf) {
    const double U = u53(w.a) - 0.5, V = u53(w.b);
    const double us = 0.5 - fabs(U);                                   // exact in FP64
    const float slam = perturb(sqrtf((float)lam), noise, 0);              // MUFU.SQRT
    const float b = fmaf(2.53f, slam, 0.931f);
    const float a = fmaf(0.02483f, b, -0.059f);
    const float r = perturb(1.0f / (float)us, noise >> 16, 1);            // MUFU.RCP
    const float t = fmaf(2.0f * a, r, b) * (float)U;
    // relative error of t: sqrt 2^-22 (+conversion), a and b inherit it, rcp 2^-22 + conversion, four roundings: < 1e-6
    const double Et = 1e-6 * fabs((double)t) + 1e-9;
    const double arg = (double)t + (lam + 0.43);
    kf = floor(arg);
    const double frac = arg - kf;
    const bool kf_sure = frac > Et && frac < 1.0 - Et;
    // squeeze: (0.9277 - V)(b - 2) >= 3.6224 with b known to 4e-7 relative
    const double bd = (double)b, sq = (0.9277 - V) * (bd - 2.0), Esq = 6e-7 * bd + 1e-9;
    if (us >= 0.07) {
        if (sq >= 3.6224 + Esq) return kf_sure ? 1 : 3;
        if (sq > 3.6224 - Esq) return 3;
    }
    if (!kf_sure) return 3;
    if (kf < 0.0 || (us < 0.013 && V > us)) return 0;
    const float bm = b - 3.4f;
    numf = (float)V * fmaf(1.1239f, bm, 1.1328f);
    denf = bm * fmaf(a * r, r, b);
    return 2;
}

int main(int argc, char** argv) {
    const long long per = argc > 1 ? atoll(argv[1]) : 4000000LL;
    g_mode = argc > 2 ? atoi(argv[2]) : 0;
    std::vector<double> lams;
    for (double l = 10.0; l < 2e7; l *= 1.31) lams.push_back(l);
    long long tot = 0, und = 0, bad = 0, und2 = 0, slow = 0;
    printf("%12s %10s %10s %10s %10s %8s\n", "lambda", "attempts", "undecided", "acc.tests", "undecided2", "wrong");
    for (size_t li = 0; li < lams.size(); ++li) {
        const double lam = lams[li];
        long long u = 0, b = 0, u2 = 0, sl = 0;
        #pragma omp parallel for reduction(+ : u, b, u2, sl) schedule(static)
        for (long long i = 0; i < per; ++i) {
            Stream st(0x57d7ULL + li, (uint32_t)i, 0, KIND_MODEL);
            const U64x2 w = st.draw();
            const uint32_t noise = (uint32_t)(st.block(1u << 20).a);
            double kf0, num = 0.0, den = 0.0, kf1;
            float numf = 0.0f, denf = 0.0f;
            int s0 = ptrs_candidate(lam, w, kf0, num, den);
            if (s0 == 2) s0 = (int)ptrs_exact(lam, kf0, num, den);       // the decision of the spec'd attempt
            int s1 = approx_candidate(lam, w, kf1, noise, numf, denf);
            if (s1 == 3) { u++; continue; }
            if (s1 == 2) {
                sl++;
                const int dec = filter2_emul(lam, kf1, numf, denf, noise >> 13);
                if (dec == 0) { u2++; if (kf1 != kf0) b++; continue; }
                s1 = dec > 0;
            }
            if (s1 != s0 || (s0 == 1 && kf1 != kf0)) b++;
        }
        if (li % 4 == 0 || b) printf("%12.5g %10lld %10lld %10lld %10lld %8lld\n", lam, per, u, sl, u2, b);
        tot += per; und += u; bad += b; und2 += u2; slow += sl;
    }
    printf("total %lld attempts: %.3f %% undecided by the candidate, %.2f %% of the %lld acceptance tests undecided by filter 2, %lld wrong decisions\n",
           tot, 100.0 * und / tot, 100.0 * und2 / (slow ? slow : 1), slow, bad);
    return bad != 0;
}
