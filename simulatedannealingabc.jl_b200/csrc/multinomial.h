// multinomial.h -- host side of the sharded resampling: the N iid categorical draws of
// resample_population (src/SimulatedAnnealingABC.jl:129) are split over the ranks by ONE multinomial draw of the per-rank counts
// c ~ Multinomial(N; w_g / W) from a seed every rank shares; rank g then makes c_g draws among its own particles.  Given the
// counts, the draws inside a rank's weight range are iid, so the result is exactly the N-fold categorical draw of the reference
// while each rank only touches its own N/G weights.
//
// The counts come from G-1 conditional binomials.  Binomial(n, p): Hoermann's BTRS (transformed rejection with squeeze, 1993)
// when n min(p, 1-p) >= 10, sequential-search inversion below; uniforms from a Philox stream, logarithms from detmath, so every
// rank (and every host) computes the same vector.
#pragma once
#include "philox.cuh"
#include <vector>

namespace sabc {

struct HostRng {
    uint32_t k0, k1, c1, c3, next = 0;
    U64x2 cur{0, 0};
    int have = 0;
    HostRng(uint64_t seed, uint32_t resample_count, uint32_t tag) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c1(resample_count), c3(tag) {}
    double uniform() {                                  // (0, 1)
        if (!have) { cur = philox4x32_10(next++, c1, 0xC0117u, c3, k0, k1); have = 2; }
        const uint64_t w = have == 2 ? cur.a : cur.b;
        --have;
        return u53_mid(w);
    }
};

// log(k!) - [ log(sqrt(2 pi)) + (k + 1/2) log(k + 1) - (k + 1) ]
inline double stirling_tail(double k) {
    static const double t[10] = {0.0810614667953272, 0.0413406959554092, 0.0276779256849983, 0.02079067210376509, 0.0166446911898211,
                                 0.0138761288230707, 0.0118967099458917, 0.0104112652619720, 0.00925546218271273, 0.00833056343336287};
    if (k <= 9.0) return t[(int)k];
    const double kp1 = k + 1.0, kp1sq = kp1 * kp1;
    return (1.0 / 12.0 - (1.0 / 360.0 - 1.0 / 1260.0 / kp1sq) / kp1sq) / kp1;
}

inline int64_t binomial_draw(int64_t n, double p, HostRng& rng) {
    if (n <= 0 || !(p > 0.0)) return 0;
    if (p >= 1.0) return n;
    if (p > 0.5) return n - binomial_draw(n, 1.0 - p, rng);
    const double nd = (double)n;
    if (nd * p < 10.0) {                                 // inversion: geometric waiting times between successes
        const double lq = det_log(1.0 - p);
        double sum = 0.0;
        int64_t k = 0;
        for (;;) {
            sum += floor(det_log(rng.uniform()) / lq) + 1.0;   // trials up to and including the next success
            if (sum > nd) return k;
            ++k;
        }
    }
    const double q = 1.0 - p, sd = sqrt(nd * p * q);
    const double b = 1.15 + 2.53 * sd, a = -0.0873 + 0.0248 * b + 0.01 * p, c = nd * p + 0.5, vr = 0.92 - 4.2 / b, r = p / q;
    const double alpha = (2.83 + 5.1 / b) * sd, m = floor((nd + 1.0) * p);
    for (;;) {
        const double u = rng.uniform() - 0.5;
        double v = rng.uniform();
        const double us = 0.5 - fabs(u);
        const double k = floor((2.0 * a / us + b) * u + c);
        if (us >= 0.07 && v <= vr) return (int64_t)k;
        if (k < 0.0 || k > nd) continue;
        v = det_log(v * alpha / (a / (us * us) + b));
        const double ub = (m + 0.5) * det_log((m + 1.0) / (r * (nd - m + 1.0))) + (nd + 1.0) * det_log((nd - m + 1.0) / (nd - k + 1.0)) +
                          (k + 0.5) * det_log(r * (nd - k + 1.0) / (k + 1.0)) + stirling_tail(m) + stirling_tail(nd - m) - stirling_tail(k) -
                          stirling_tail(nd - k);
        if (v <= ub) return (int64_t)k;
    }
}

// counts[g] ~ Multinomial(N; w[g] / sum w), the same on every rank for the same (w, N, seed, resample_count)
inline void multinomial_split(int64_t N, const unsigned long long* w, int G, uint64_t seed, uint32_t resample_count, int64_t* counts) {
    HostRng rng(seed, resample_count, KIND_RESAMPLE | 0x100u);
    unsigned __int128 rest = 0;                          // exact integer weight of the ranks not yet served
    for (int g = 0; g < G; ++g) rest += w[g];
    int64_t left = N;
    for (int g = 0; g + 1 < G; ++g) {
        int64_t c = 0;
        if (left > 0 && w[g] > 0) c = (unsigned __int128)w[g] == rest ? left : binomial_draw(left, (double)w[g] / (double)rest, rng);
        counts[g] = c; left -= c; rest -= w[g];
    }
    counts[G - 1] = left;
}

}  // namespace sabc
