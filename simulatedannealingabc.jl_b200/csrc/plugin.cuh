// plugin.cuh -- operator plug-ins of the SABC hot path: prior, proposals, ECDF lookup, models.
//
// These are the device forms of the reference's plug points:
//   prior     : rand(prior) / logpdf(prior, θ)      src/SimulatedAnnealingABC.jl:163,174,314,318
//   proposal  : (p::Proposal)(θ, population_inactive)  src/proposals.jl:40-55,101-114,137-148
//   G (ECDF)  : cdfs_dist_prior(ρ)                  src/cdf_estimators.jl:68-70
//   f_dist    : user model + distance               src/SimulatedAnnealingABC.jl:315,421
// Everything is header-only so that a model plug-in compiled out of tree instantiates the same
// fused kernels (kernels.cuh) and registers its launchers (sabc_register_model).
#pragma once
#include "philox.cuh"

namespace sabc {

constexpr int MAX_D = 8;            // max parameters per particle
constexpr int MAX_S = 32;           // max summary statistics
constexpr int MAX_MODEL_PAR = 40;   // doubles in a model parameter blob
constexpr int ECDF_MAX_LEVELS = 10;
constexpr int ECDF_PAD = 16;        // +inf entries appended to a knot table
constexpr int CHUNK = 256;          // particles per tree-sum group == threads per CTA

enum Algorithm : int32_t { ALG_SINGLE_EPS = 0, ALG_MULTI_EPS = 1 };
enum ProposalKind : int32_t { PROP_DE = 0, PROP_STRETCH = 1, PROP_RW = 2 };
enum PriorKind : int32_t { PRIOR_UNIFORM = 0, PRIOR_NORMAL = 1, PRIOR_EXPONENTIAL = 2, PRIOR_LOGNORMAL = 3, PRIOR_GAMMA = 4, PRIOR_BETA = 5,
                           PRIOR_CAUCHY = 6, PRIOR_LAPLACE = 7, PRIOR_WEIBULL = 8, PRIOR_INVGAMMA = 9, PRIOR_KIND_END = 10 };

// ------------------------------------------------------------------------------------------------
// Prior: product of independent univariates (Distributions ^0.25 formulas)
// ------------------------------------------------------------------------------------------------
struct PriorSpec {
    int32_t n;
    int32_t kind[MAX_D];
    double p0[MAX_D], p1[MAX_D];   // Uniform(a,b) | Normal(mu,sigma) | Exponential(theta,-) | LogNormal(mu,sigma) | Gamma(alpha,theta) | Beta(alpha,beta)
    double c[MAX_D];               // -log(b-a)    | log(sigma)       | log(theta)           | log(sigma)   (det_log)
                                   // Gamma: lgamma(alpha) + alpha log(theta) | Beta: lgamma(alpha) + lgamma(beta) - lgamma(alpha+beta)
                                   // Cauchy(mu,sigma): log(sigma) | Laplace(mu,theta): log(2 theta) | Weibull(alpha,theta): log(alpha/theta)
                                   // InverseGamma(alpha,theta): lgamma(alpha) - alpha log(theta)
};

// parameter domains of Distributions.jl's constructors
inline bool prior_params_valid(int kind, double p0, double p1) {
    switch (kind) {
        case PRIOR_UNIFORM: return p0 < p1;
        case PRIOR_NORMAL: case PRIOR_LOGNORMAL: case PRIOR_CAUCHY: case PRIOR_LAPLACE: return p1 > 0.0 && p0 == p0;
        case PRIOR_EXPONENTIAL: return p0 > 0.0;
        case PRIOR_GAMMA: case PRIOR_BETA: case PRIOR_WEIBULL: case PRIOR_INVGAMMA: return p0 > 0.0 && p1 > 0.0;
        default: return false;
    }
}

inline void prior_prepare(PriorSpec& p) {
    for (int i = 0; i < p.n; ++i) {
        switch (p.kind[i]) {
            case PRIOR_UNIFORM: p.c[i] = -det_log(p.p1[i] - p.p0[i]); break;
            case PRIOR_EXPONENTIAL: p.c[i] = det_log(p.p0[i]); break;
            case PRIOR_GAMMA: p.c[i] = det_lgamma(p.p0[i]) + p.p0[i] * det_log(p.p1[i]); break;
            case PRIOR_BETA: p.c[i] = (det_lgamma(p.p0[i]) + det_lgamma(p.p1[i])) - det_lgamma(p.p0[i] + p.p1[i]); break;
            case PRIOR_LAPLACE: p.c[i] = det_log(2.0 * p.p1[i]); break;
            case PRIOR_WEIBULL: p.c[i] = det_log(p.p0[i] / p.p1[i]); break;
            case PRIOR_INVGAMMA: p.c[i] = det_lgamma(p.p0[i]) - p.p0[i] * det_log(p.p1[i]); break;
            default: p.c[i] = det_log(p.p1[i]); break;
        }
    }
}

// log-density of one univariate component (Distributions ^0.25 formulas)
SABC_HD double prior_logpdf1(int kind, double p0, double p1, double c, double x) {
    const double LOG2PI = 0x1.d67f1c864beb5p+0;
    if (kind == PRIOR_NORMAL) {
        const double z = (x - p0) / p1;
        return -((z * z + LOG2PI) * 0.5) - c;
    }
    if (kind == PRIOR_UNIFORM) return (x >= p0 && x <= p1) ? c : -dinf();
    if (kind == PRIOR_EXPONENTIAL) return x >= 0.0 ? (-(x / p0)) - c : -dinf();
    if (kind == PRIOR_GAMMA) {                               // (alpha-1) log x - x/theta - c, xlogy(0, 0) = 0 at the edge
        if (!(x >= 0.0)) return -dinf();
        const double a1 = p0 - 1.0;
        const double t = a1 == 0.0 ? 0.0 : a1 * det_log(x);
        return (t - x / p1) - c;
    }
    if (kind == PRIOR_BETA) {                                // (alpha-1) log x + (beta-1) log(1-x) - c on [0, 1]
        if (!(x >= 0.0 && x <= 1.0)) return -dinf();
        const double a1 = p0 - 1.0, b1 = p1 - 1.0;
        const double t0 = a1 == 0.0 ? 0.0 : a1 * det_log(x);
        const double t1 = b1 == 0.0 ? 0.0 : b1 * det_log(1.0 - x);
        return (t0 + t1) - c;
    }
    if (kind == PRIOR_CAUCHY) {                              // -(log1p(z^2) + log(pi) + log(sigma))
        const double z = (x - p0) / p1;
        return -((det_log(1.0 + z * z) + 0x1.250d048e7a1bdp+0) + c);
    }
    if (kind == PRIOR_LAPLACE) return -(fabs(x - p0) / p1 + c);
    if (kind == PRIOR_WEIBULL) {                             // log(alpha/theta) + (alpha-1) log z - z^alpha, z = x/theta
        if (!(x >= 0.0)) return -dinf();
        const double lz = det_log(x / p1);
        const double a1 = p0 - 1.0;
        const double t = a1 == 0.0 ? 0.0 : a1 * lz;
        return (c + t) - det_exp(p0 * lz);
    }
    if (kind == PRIOR_INVGAMMA) {                            // alpha log(theta) - lgamma(alpha) - (alpha+1) log x - theta/x
        if (!(x > 0.0)) return -dinf();
        return (-((p0 + 1.0) * det_log(x)) - p1 / x) - c;
    }
    if (!(x > 0.0)) return -dinf();                         // LogNormal
    const double lx = det_log(x);
    const double z = (lx - p0) / p1;
    return (-((z * z + LOG2PI) * 0.5) - c) - lx;
}

template <int D>
SABC_HD double prior_logpdf(const PriorSpec& p, const double (&th)[D]) {
    double lp = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const double t = prior_logpdf1(p.kind[c], p.p0[c], p.p1[c], p.c[c], th[c]);
        lp = (c == 0) ? t : lp + t;                          // product distribution: sum of the components
    }
    return lp;
}

// Gamma(a, 1) by Marsaglia & Tsang (2000): d = a - 1/3 (a + 1 - 1/3 below 1, then the draw is scaled by U^(1/a)),
// v = (1 + z / sqrt(9 d))^3, accept when log(U) < z^2/2 + d - d v + d log v.  Attempt t of component `comp` reads the
// prior stream's blocks comp + 256 (base + 2t + 1) [normal pair, first output] and comp + 256 (base + 2t + 2)
// [word a: acceptance uniform, word b: the scaling uniform for a < 1]; block `comp` itself stays with the one-block families.
SABC_HD double gamma_std(double a, const Stream& st, uint32_t comp, uint32_t base) {
    const double ae = a < 1.0 ? a + 1.0 : a;
    const double d = ae - 1.0 / 3.0;
    const double cc = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0; t < 100000u; ++t) {
        double z, z1;
        normal_pair(st.block(comp + 256u * (base + 2u * t + 1u)), z, z1);
        const U64x2 w = st.block(comp + 256u * (base + 2u * t + 2u));
        double v = 1.0 + cc * z;
        if (!(v > 0.0)) continue;
        v = (v * v) * v;
        const double rhs = ((0.5 * z) * z + d) - d * v + d * det_log(v);
        if (det_log(u53_open0(w.a)) < rhs) {
            double g = d * v;
            if (a < 1.0) g = g * det_exp(det_log(u53_open0(w.b)) / a);
            return g;
        }
    }
    return d;
}

template <int D>
SABC_HD void prior_rand(const PriorSpec& p, uint64_t seed, uint32_t particle, double (&th)[D]) {
    const Stream st(seed, particle, 0, KIND_PRIOR);
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const U64x2 w = st.block((uint32_t)c);
        double z0, z1;
        switch (p.kind[c]) {
            case PRIOR_NORMAL: normal_pair(w, z0, z1); th[c] = p.p0[c] + p.p1[c] * z0; break;
            case PRIOR_UNIFORM: th[c] = p.p0[c] + (p.p1[c] - p.p0[c]) * u53(w.a); break;
            case PRIOR_EXPONENTIAL: th[c] = p.p0[c] * (-det_log(u53_open0(w.a))); break;
            case PRIOR_GAMMA: th[c] = p.p1[c] * gamma_std(p.p0[c], st, (uint32_t)c, 0u); break;
            case PRIOR_BETA: {
                const double g1 = gamma_std(p.p0[c], st, (uint32_t)c, 0u), g2 = gamma_std(p.p1[c], st, (uint32_t)c, 1u << 20);
                th[c] = g1 / (g1 + g2);
                break;
            }
            case PRIOR_CAUCHY: {                              // quantile: mu + sigma tan(pi (u - 1/2)) = mu - sigma cos(pi u) / sin(pi u)
                double sn, cs; det_sincos2pi(0.5 * u53_mid(w.a), sn, cs);
                th[c] = p.p0[c] - p.p1[c] * (cs / sn);
                break;
            }
            case PRIOR_LAPLACE: {                             // quantile: mu + theta log(2u) below 1/2, mu - theta log(2(1-u)) above
                const double u = u53_mid(w.a);
                th[c] = u < 0.5 ? p.p0[c] + p.p1[c] * det_log(2.0 * u) : p.p0[c] - p.p1[c] * det_log(2.0 * (1.0 - u));
                break;
            }
            case PRIOR_WEIBULL: th[c] = p.p1[c] * det_exp(det_log(-det_log(u53_open0(w.a))) / p.p0[c]); break;   // theta E^(1/alpha), E ~ Exp(1)
            case PRIOR_INVGAMMA: th[c] = p.p1[c] / gamma_std(p.p0[c], st, (uint32_t)c, 0u); break;
            default: normal_pair(w, z0, z1); th[c] = det_exp(p.p0[c] + p.p1[c] * z0); break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Proposals.  P(c, i) reads coordinate c of particle i of the inactive half.
// ------------------------------------------------------------------------------------------------
struct CtrlWords { uint64_t A, B, C, D; };   // blocks 0 and 1 of the control stream

SABC_HD CtrlWords ctrl_words(uint64_t seed, uint32_t particle, uint64_t sweep, const uint32_t* rk = nullptr) {
    const Stream st(seed, particle, sweep, KIND_CTRL, rk);
    const U64x2 b0 = st.block(0), b1 = st.block(1);
    return CtrlWords{b0.a, b0.b, b1.a, b1.b};
}

// `ctrl` = the particle's control stream; the jitter normal comes from word C, its rare slow path from blocks 2, 3, ...
template <int D, class Gather>
SABC_HD void propose_de(const double (&th)[D], const Gather& P, int64_t M, double gamma0, double sigma_gamma,
                        const CtrlWords& cw, Stream ctrl, double (&out)[D], double& log_factor) {
    int64_t i1 = (int64_t)mulhi64(cw.A, (uint64_t)M);               // proposals.jl:103-107
    int64_t i2 = (int64_t)mulhi64(cw.B, (uint64_t)(M - 1));
    if (i2 >= i1) i2++;
    ctrl.next = 2;
    const double g = gamma0 * (1.0 + sigma_gamma * zig_normal(cw.C, ctrl)); // :110
#pragma unroll
    for (int c = 0; c < D; ++c) out[c] = th[c] + g * (P(c, i1) - P(c, i2));   // :113
    log_factor = 0.0;
}

template <int D, class Gather>
SABC_HD void propose_stretch(const double (&th)[D], const Gather& P, int64_t M, double a, const CtrlWords& cw,
                             double (&out)[D], double& log_factor) {
    const int64_t i = (int64_t)mulhi64(cw.A, (uint64_t)M);          // proposals.jl:141
    const double t = (a - 1.0) * u53(cw.B) + 1.0;
    const double z = (t * t) / a;                                   // :144
    log_factor = det_log(z) * (double)(D - 1);                      // :146
#pragma unroll
    for (int c = 0; c < D; ++c) { const double p = P(c, i); out[c] = p + z * (th[c] - p); }   // :147
}

// chol: row-major lower Cholesky factor of Σ (D>1) or sqrt(Σ) (D==1)   proposals.jl:42,54
template <int D>
SABC_HD void propose_rw(const double (&th)[D], const double* chol, uint64_t seed, uint32_t particle, uint64_t sweep,
                        double (&out)[D], double& log_factor, uint32_t zig_smem = 0) {
    Stream st(seed, particle, sweep, KIND_RW, nullptr, zig_smem);
    double z[D + 1];
#pragma unroll
    for (int c = 0; c < D; c += 2) normal2(st, z[c], z[c + 1 < D ? c + 1 : D]);
    if (D == 1) {
        out[0] = th[0] + chol[0] * z[0];
    } else {
#pragma unroll
        for (int r = 0; r < D; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c <= r; ++c) { const double t = chol[r * D + c] * z[c]; acc = (c == 0) ? t : acc + t; }
            out[r] = th[r] + acc;
        }
    }
    log_factor = 0.0;
}

// ------------------------------------------------------------------------------------------------
// ECDF: sorted knots K[0..L) in HBM (level 0) plus a sampled index: level k+1 = every 9th entry of level k, until a level has
// at most 2048 / 1024 / ... entries (power of two, by the number of statistics); that top level is staged in shared memory.
// Evaluation follows Interpolations.jl's LinearMonotonicInterpolation + Flat() exactly: bracket by
// searchsortedfirst-1, slope by division, one rounded multiply then one rounded add.
//
// What bounds the lookup on a B200 (profiles/r2_c5_*): every lane of a warp searches a different region, so every load
// instruction costs one L1 wavefront PER LANE whatever its width up to 16 bytes, and the fused Gaussian kernels spend their time
// on exactly those wavefronts (78 % of the L1 data-pipe peak with a 16-ary index searched by scalar probes: 20 loads per lookup).
// The index is therefore laid out for two 16-byte probes per level: below entry m of level k+1 sit the eight entries
// e1..e8 = A_k[9m+1 .. 9m+8], stored as one 64-byte node in PROBE ORDER [e3 e6 | e1 e2 | e4 e5 | e7 e8]; the first probe
// (e3, e6) selects the pair of the second, and 3 c1 + c2 is the count of entries below x.  Every probed value also tightens
// a running bracket (largest value < x, smallest value >= x), which by induction over the levels ends as (K[j], K[j+1]):
// the knots themselves are never read again.  8 wavefronts per lookup at L = 10^7 instead of 20, a third of the instructions.
// ------------------------------------------------------------------------------------------------
constexpr int ECDF_NODE = 8;        // entries per node (64 bytes)
constexpr int ECDF_STRIDE = 9;      // level k+1 samples every 9th entry of level k
struct EcdfStat {
    const double* knots;                      // K[0..L), natural order (returned by sabc_get_ecdf; not read by the lookup)
    const double* node[ECDF_MAX_LEVELS];      // node[k]: probe-ordered nodes for the search inside level k, k = 0 .. nlev-2
    const double* top;                        // top level, natural order, top_pow2 entries (+inf padded), stored SPLIT: the high 32-bit
                                              // words of all entries, then the low words (same bytes; see ecdf_eval)
    int64_t L;
    double kmax;                              // K[L-1]
    int32_t nlev;                             // levels incl. the knots; level nlev-1 is the staged one
    int32_t top_cnt, top_pow2;                // real / padded (power of two) entries of the top level
    int32_t top_off;                          // offset (doubles) of the staged level in shared memory
};

// number of entries of sorted a[0..n) that are < x
SABC_HD int64_t count_less(const double* a, int64_t n, double x) {
    int64_t lo = 0;
    while (n > 0) {
        const int64_t half = n >> 1;
        if (a[lo + half] < x) { lo += half + 1; n -= half + 1; } else { n = half; }
    }
    return lo;
}

#if defined(__CUDACC__)
// the Metropolis test log(U) < L  (src/SimulatedAnnealingABC.jl:324)
SABC_D bool log_u_less(double U, double L) { return det_log(U) < L; }

// The lookup in three stages, so that the lookups of several statistics of one particle can be interleaved (ecdf_eval_all):
// the dependent loads of one lookup (12 shared-memory steps, then two global probes per level) are what a simulation-heavy
// model with many statistics waits for -- the logistic model's 20 lookups per particle ran at 9 % issue-slot utilisation as a
// chain of 160 dependent loads; interleaved, CH independent chains are in flight per thread.
struct EcdfCursor { double x, lo, hi; int64_t lb; };

SABC_D void ecdf_top_search(const EcdfStat& e, const double* s_top, double rho, EcdfCursor& q) {
    double x = rho > e.kmax ? e.kmax : (rho < 0.0 ? 0.0 : rho);              // Flat(): clamp to [K_1, K_L]
    q.x = x; q.lb = 0; q.lo = 0.0; q.hi = 0.0;
    if (!(x == x)) return;                                                    // NaN distance: handled by ecdf_finish
    x = x + 0.0;                                                              // -0.0 -> +0.0: from here on x orders like its bit pattern
    q.x = x;
    // searchsortedfirst on the staged top level: bisection over a power-of-two table padded with +inf.  The table is staged as
    // 32-bit words (all high words, then all low words) and compared as integers -- non-negative doubles order like their bit
    // patterns: 32 lanes reading random 8-byte entries cost 10-20 shared-memory wavefronts per step (bank conflicts), random 4-byte
    // words 3-4; the low word is only read when the high words tie.
    const uint32_t* Th = reinterpret_cast<const uint32_t*>(s_top + e.top_off);
    const uint32_t* Tl = Th + e.top_pow2;
    const uint32_t xh = (uint32_t)(f64_bits(x) >> 32), xl = (uint32_t)f64_bits(x);
    int c = 0;
    for (int step = e.top_pow2 >> 1; step >= 1; step >>= 1) {
        const uint32_t th = Th[c + step - 1];
        bool lt = th < xh;
        if (th == xh) lt = Tl[c + step - 1] < xl;
        c += lt ? step : 0;
    }
    {
        const uint32_t th = Th[c];
        bool lt = th < xh;
        if (th == xh) lt = Tl[c] < xl;
        c += lt ? 1 : 0;
    }
    if (c == 0) return;                                                       // no knot below x: x is K_1 = 0
    q.lo = bits_f64(((uint64_t)Th[c - 1] << 32) | Tl[c - 1]);
    q.hi = c < e.top_pow2 ? bits_f64(((uint64_t)Th[c] << 32) | Tl[c]) : dinf();   // A[c-1] < x <= A[c]
    q.lb = c;
}
// one level down: the two 16-byte probes of the node below entry lb - 1
SABC_D void ecdf_descend(const EcdfStat& e, int lv, EcdfCursor& q) {
    const double2* nd = reinterpret_cast<const double2*>(e.node[lv] + (q.lb - 1) * ECDF_NODE);
    const double x = q.x;
    const double2 s = __ldg(nd);                                              // (e3, e6)
    const int c1 = (s.x < x ? 1 : 0) + (s.y < x ? 1 : 0);
    const double2 p = __ldg(nd + 1 + c1);                                     // (e1,e2) | (e4,e5) | (e7,e8)
    const int c2 = (p.x < x ? 1 : 0) + (p.y < x ? 1 : 0);
    const double lo_s = c1 == 0 ? q.lo : (c1 == 1 ? s.x : s.y), hi_s = c1 == 0 ? s.x : (c1 == 1 ? s.y : q.hi);
    q.lo = c2 == 0 ? lo_s : (c2 == 1 ? p.x : p.y);
    q.hi = c2 == 0 ? p.x : (c2 == 1 ? p.y : hi_s);
    q.lb = (q.lb - 1) * ECDF_STRIDE + 1 + 3 * c1 + c2;
}
SABC_D double ecdf_finish(const EcdfStat& e, const EcdfCursor& q) {
    if (!(q.x == q.x)) return q.x;                                            // u = 0 + m (NaN - 0)
    if (q.lb == 0) return 0.0;                                                // first interval: u = 0 + m (0 - 0)
    const int64_t j = q.lb - 1;                                               // k > 1 && (k -= 1); lo = K[j], hi = K[j+1]
    const double Lm1 = (double)(e.L - 1);
    const double y0 = (double)j / Lm1, y1 = (double)(j + 1) / Lm1;           // range(0, stop=1, length=L)
    const double m = (y1 - y0) / (q.hi - q.lo);
    return y0 + m * (q.x - q.lo);
}
SABC_D double ecdf_eval(const EcdfStat& e, const double* s_top, double rho) {
    EcdfCursor q;
    ecdf_top_search(e, s_top, rho, q);
    if (q.lb > 0)
        for (int lv = e.nlev - 2; lv >= 0; --lv) ecdf_descend(e, lv, q);
    return ecdf_finish(e, q);
}
// u[j] = G_j(rho[j]) for the S statistics of one particle, CH lookups at a time with their loads interleaved
template <int S>
SABC_D void ecdf_eval_all(const EcdfStat* e, const double* s_top, const double (&rho)[S], double (&u)[S]) {
    constexpr int CH = S < 4 ? S : 4;
#pragma unroll
    for (int j0 = 0; j0 < S; j0 += CH) {
        EcdfCursor q[CH];
        int top = 0;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (j0 + k < S) { ecdf_top_search(e[j0 + k], s_top, rho[j0 + k], q[k]); top = e[j0 + k].nlev > top ? e[j0 + k].nlev : top; }
        for (int lv = top - 2; lv >= 0; --lv) {
#pragma unroll
            for (int k = 0; k < CH; ++k)
                if (j0 + k < S && lv <= e[j0 + k].nlev - 2 && q[k].lb > 0) ecdf_descend(e[j0 + k], lv, q[k]);
        }
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (j0 + k < S) u[j0 + k] = ecdf_finish(e[j0 + k], q[k]);
    }
}
#endif

// host/CPU-free reference form used by the unit hook kernel (plain binary search over the knots)
SABC_HD double ecdf_eval_flat(const double* K, int64_t L, double rho) {
    const double x = rho > K[L - 1] ? K[L - 1] : (rho < K[0] ? K[0] : rho);
    int64_t j = count_less(K, L, x);
    if (j > 0) j -= 1;
    if (j > L - 2) j = L - 2;
    const double Lm1 = (double)(L - 1);
    const double y0 = (double)j / Lm1, y1 = (double)(j + 1) / Lm1;
    const double m = (y1 - y0) / (K[j + 1] - K[j]);
    return y0 + m * (x - K[j]);
}

// ------------------------------------------------------------------------------------------------
// Models (device f_dist).  A model is a struct with
//   static constexpr int D, S;                         parameters, statistics
//   static constexpr int FUSED_MIN_BLOCKS, SIM_MIN_BLOCKS, KEY_BITS;   occupancy requests of the fused / simulation kernels, work-list key
//   static __device__ void sim(th, par, stream, rho)   simulate + distance(s) >= 0
// ------------------------------------------------------------------------------------------------
struct ModelPar { double v[MAX_MODEL_PAR]; };

// C1/C5: 1-D Gaussian mean, sufficient-statistic form.  par: ybar_obs, sd_mean (= sigma/sqrt(n))
#ifndef SABC_GAUSSMEAN_MIN_BLOCKS
#define SABC_GAUSSMEAN_MIN_BLOCKS 6
#endif
#ifndef SABC_GAUSSSAMPLE_MIN_BLOCKS
#define SABC_GAUSSSAMPLE_MIN_BLOCKS 4
#endif
struct GaussMean {
    static constexpr int D = 1, S = 1;
    static constexpr int FUSED_MIN_BLOCKS = SABC_GAUSSMEAN_MIN_BLOCKS;   // resident CTAs per SM requested for the fused kernel (register cap)
    static constexpr int SIM_MIN_BLOCKS = 1;   // resident CTAs per SM requested for the simulation kernel
    static constexpr int KEY_BITS = 0;         // no work-list bucketing
    SABC_HD static void sim(const double (&th)[1], const ModelPar& mp, Stream& st, double (&rho)[1]) {
        const double z0 = normal1(st);
        const double ysim = th[0] + mp.v[1] * z0;
        rho[0] = fabs(ysim - mp.v[0]);
    }
};

// n iid draws N(θ1, θ2 or fixed σ); statistics |obs1 - mean|, |obs2 - mean(y²) or Σy²|
// (test/runtests.jl:35,86,128-131,167-170; docs/src/usage.md:16-35).  par: n, sigma_fixed, obs1, obs2, stat2_kind
template <int D_, int S_>
struct GaussSample {
    static constexpr int D = D_, S = S_;
    static constexpr int FUSED_MIN_BLOCKS = SABC_GAUSSSAMPLE_MIN_BLOCKS;
    static constexpr int SIM_MIN_BLOCKS = 1;   // resident CTAs per SM requested for the simulation kernel
    static constexpr int KEY_BITS = 0;         // no work-list bucketing
    SABC_HD static void sim(const double (&th)[D_], const ModelPar& mp, Stream& st, double (&rho)[S_]) {
        const int n = (int)mp.v[0];
        const double sig = D_ >= 2 ? th[D_ - 1] : mp.v[1];
        double s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < n; k += 2) {
            double z0, z1; normal2(st, z0, z1);
            double y = th[0] + sig * z0;
            s1 = s1 + y; s2 = s2 + y * y;
            if (k + 1 < n) { y = th[0] + sig * z1; s1 = s1 + y; s2 = s2 + y * y; }
        }
        rho[0] = fabs(mp.v[2] - s1 / (double)n);
        if (S_ >= 2) rho[S_ - 1] = fabs(mp.v[3] - (mp.v[4] != 0.0 ? s2 : s2 / (double)n));
    }
};

// C3: stochastic logistic growth, θ = (r, K, σ), T = 20 points.  par: x0, T, obs[T]
#ifndef SABC_LOGISTIC_SIM_MIN_BLOCKS
#define SABC_LOGISTIC_SIM_MIN_BLOCKS 4   // 64 registers: 4 CTAs per SM hide the latency of the 20 ECDF lookups (1 / 2 / 3 / 4 / 5 CTAs: 1.07 / 0.65 / 0.50 / 0.47 / 0.70 ms per half-sweep)
#endif
struct Logistic {
    static constexpr int D = 3, S = 20;
    static constexpr int FUSED_MIN_BLOCKS = 1;
    static constexpr int SIM_MIN_BLOCKS = SABC_LOGISTIC_SIM_MIN_BLOCKS;   // resident CTAs per SM requested for the simulation kernel
    static constexpr int KEY_BITS = 0;         // no work-list bucketing
    SABC_HD static void sim(const double (&th)[3], const ModelPar& mp, Stream& st, double (&rho)[20]) {
        double x = mp.v[0];
#pragma unroll
        for (int t = 0; t < S; t += 2) {
            double z[2]; normal2<false>(st, z[0], z[1]);      // ten unrolled call sites: slow path out of line
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double grow = (th[0] * x) * (1.0 - x / th[1]);
                const double noise = (th[2] * x) * z[h];
                x = (x + grow) + noise;
                if (!(x > 0.0)) x = 0.0;
                rho[t + h] = fabs(x - mp.v[2 + t + h]);
            }
        }
    }
};

#ifndef SABC_SIR_MIN_BLOCKS
#define SABC_SIR_MIN_BLOCKS 4
#endif
// C4: SIR tau-leap, θ = (β, γ, ι, φ).  par: pop, T, tau, obs_total, obs_peak, obs_tpeak
struct SirTauLeap {
    static constexpr int D = 4, S = 3;
    static constexpr int FUSED_MIN_BLOCKS = 1;
    static constexpr int NO_NORMALS = 1;
    static constexpr int SIM_MIN_BLOCKS = SABC_SIR_MIN_BLOCKS;   // 4: cap at 64 registers, 32 warps per SM hide the FP64 latencies
    // similarity key of a proposal for the work-list bucketing: growth rate β−γ (6 bits), initial fraction ι (3), γ (3).
    // Particles of one bucket have similar epidemic curves, so the lanes of a warp meet the same sampler regimes and
    // finish together.  The key only orders the work; it never enters the arithmetic.
    static constexpr int KEY_BITS = 12;
    SABC_HD static uint32_t work_key(const double (&th)[4], const ModelPar&) {
        auto q = [](double x, double lo, double hi, int n) { int v = (int)((x - lo) / (hi - lo) * (double)n); return (uint32_t)(v < 0 ? 0 : (v >= n ? n - 1 : v)); };
        return (q(th[0] - th[1], -0.4, 0.95, 64) << 6) | (q(th[2], 0.001, 0.05, 8) << 3) | q(th[1], 0.05, 0.5, 8);
    }
    // Per step: n_inf ~ Poisson(β S I/pop τ) ∧ S, n_rec ~ Poisson(γ I τ) ∧ I, cases ~ Poisson(φ n_inf).  Written as a
    // per-lane state machine over sampler ATTEMPTS (phase 0/1/2 = the three draws of a step): every loop trip each
    // lane makes one attempt on its own current draw, so a PTRS rejection costs that lane one trip instead of
    // stalling the whole warp.  The blocks a particle consumes, and hence its result, are those of the plain
    // sequential loop over steps and draws.  (Measured and rejected: parking the lanes that need the exact PTRS test or the
    // inversion loop, r1 notes; persistent lanes that take the next particle as soon as theirs is finished, with the ECDF /
    // accept epilogue as a separate convergent kernel -- 0.845 ms against 0.783 per half-sweep, the live lanes per instruction
    // did not rise: the idle lanes sit inside a trip, in the divergent sampler branches, not behind finished particles;
    // profiles/r2_notes.md.)
    SABC_HD static void sim(const double (&th)[4], const ModelPar& mp, Stream& st, double (&rho)[3]) {
        const double pop = mp.v[0], tau = mp.v[2];
        const double inv_pop = drcp(pop);
        const int T = (int)mp.v[1];
        const double popi = (double)(int64_t)pop;
        double I = floor(th[2] * pop + 0.5);
        if (I < 0.0) I = 0.0;
        if (I > popi) I = popi;
        double Sc = popi - I;
        double total = 0.0, peak = -1.0, ninf = 0.0;
        int tpeak = 0;
        int t = 1, phase = 0;
        double lam = (((th[0] * Sc) * I) * inv_pop) * tau;
        auto advance = [&](double k) {       // one accepted draw k moves the particle to its next draw
            if (phase == 0) {
                ninf = k > Sc ? Sc : k;
                lam = (th[1] * I) * tau;
                phase = 1;
            } else if (phase == 1) {
                const double nrec = k > I ? I : k;
                Sc -= ninf; I += ninf - nrec;
                lam = th[3] * ninf;
                phase = 2;
            } else {
                total += k;
                if (k > peak) { peak = k; tpeak = t; }
                lam = (((th[0] * Sc) * I) * inv_pop) * tau;
                phase = 0; t++;
            }
        };
        while (t <= T) {
            double k;
            if (!poisson_attempt_d(lam, st, k)) continue;
            advance(k);
        }
        const double d0 = total - mp.v[3], d1 = peak - mp.v[4], d2 = (double)tpeak - mp.v[5];
        rho[0] = d0 * d0; rho[1] = d1 * d1; rho[2] = d2 * d2;
    }
};

// The reference's documented example (docs/src/example.md:75-122): event-driven (Gillespie) SIR, θ = (β, γ), S0=99, I0=1,
// t_max=160; statistics abs2 of total infected, peak infected, time of the peak (:143-147), or their sum (S_ == 1, :152).
// par: S0, I0, R0, t_max, obs_total, obs_peak, obs_tpeak.  One Philox block per event: waiting time and event type.
template <int S_>
struct SirGillespie {
    static constexpr int D = 2, S = S_;
    static constexpr int FUSED_MIN_BLOCKS = 1;
    static constexpr int NO_NORMALS = 1;
    static constexpr int SIM_MIN_BLOCKS = 4;
    static constexpr int KEY_BITS = 0;
    SABC_HD static void sim(const double (&th)[2], const ModelPar& mp, Stream& st, double (&rho)[S_]) {
        double Sc = mp.v[0], I = mp.v[1], R = mp.v[2];
        const double tmax = mp.v[3], Npop = (Sc + I) + R;
        double t = 0.0, peak = I, tpeak = 0.0;
        for (int ev = 0; ev < 65536 && t < tmax && I > 0.0; ++ev) {
            const double inf = ((th[0] * Sc) * I) / Npop;
            const double rec = th[1] * I;
            const double tot = inf + rec;
            const U64x2 w = st.draw();
            t = t + (-det_log(u53_open0(w.a))) / tot;
            if (u53(w.b) < inf / tot) { Sc -= 1.0; I += 1.0; } else { I -= 1.0; R += 1.0; }
            if (I > peak) { peak = I; tpeak = t; }
        }
        double d0 = R - mp.v[4], d1 = peak - mp.v[5], d2 = tpeak - mp.v[6];
        d0 = d0 * d0; d1 = d1 * d1; d2 = d2 * d2;
        if (S_ >= 3) { rho[0] = d0; rho[S_ >= 3 ? 1 : 0] = d1; rho[S_ >= 3 ? 2 : 0] = d2; }
        else rho[0] = (d0 + d1) + d2;
    }
};

}  // namespace sabc
