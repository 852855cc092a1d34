// philox.cuh -- counter-based random streams and samplers.
//
// The reference draws from Julia's task-local Xoshiro (rand()/randn() inside the @threads loops,
// src/SimulatedAnnealingABC.jl:172-179,308-331), which is neither reproducible across thread
// counts nor portable to a GPU.  Here every (particle, sweep, purpose) owns a Philox4x32-10
// stream addressed by its counter, so a particle update needs no RNG state in HBM and the result
// does not depend on the launch geometry.
#pragma once
#include "detmath.cuh"

namespace sabc {

enum StreamKind : uint32_t { KIND_CTRL = 0, KIND_MODEL = 1, KIND_RW = 3, KIND_PRIOR = 4, KIND_RESAMPLE = 6 };

struct U64x2 { uint64_t a, b; };

SABC_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
SABC_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11)
SABC_HD U64x2 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    U64x2 o;
    o.a = (uint64_t)c0 | ((uint64_t)c1 << 32);
    o.b = (uint64_t)c2 | ((uint64_t)c3 << 32);
    return o;
}

// The ten round keys of a seed, precomputed on the host and handed to a kernel in its parameter block: inside a hot
// loop they are then five LDCU.128 per Philox call instead of eighteen key-schedule additions.
struct RoundKeys { uint32_t k[20]; };
inline RoundKeys make_round_keys(uint64_t seed) {
    RoundKeys rk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { rk.k[2 * r] = k0; rk.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return rk;
}
SABC_HD U64x2 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* rk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ rk[2 * r]; c1 = l1; c2 = h0 ^ c3 ^ rk[2 * r + 1]; c3 = l0;
    }
    U64x2 o;
    o.a = (uint64_t)c0 | ((uint64_t)c1 << 32);
    o.b = (uint64_t)c2 | ((uint64_t)c3 << 32);
    return o;
}

// A stream = (seed, particle, sweep, kind); block j of it is one Philox call.
struct Stream {
    uint32_t k0, k1, particle, sweep_lo, tag, next;
    uint32_t warp_mask;   // lanes known to call the model together (0: the model asks __activemask())
    uint32_t zig_smem;    // shared-space address of a CTA-staged copy of the ziggurat table (stage_zig), 0: read it through L1
    const uint32_t* rk;   // optional precomputed round keys of the same seed (RoundKeys::k), else nullptr
    SABC_HD Stream(uint64_t seed, uint32_t particle_, uint64_t sweep, uint32_t kind, const uint32_t* rk_ = nullptr, uint32_t zig_smem_ = 0)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), particle(particle_), sweep_lo((uint32_t)sweep),
          tag(kind | ((uint32_t)(sweep >> 32) << 4)), next(0), warp_mask(0), zig_smem(zig_smem_), rk(rk_) {}
    SABC_HD U64x2 block(uint32_t j) const {
        return rk ? philox4x32_10_rk(particle, sweep_lo, j, tag, rk) : philox4x32_10(particle, sweep_lo, j, tag, k0, k1);
    }
    SABC_HD U64x2 draw() { return block(next++); }
};

SABC_HD double u53(uint64_t x) { return (double)(x >> 11) * 0x1p-53; }              // [0,1)
SABC_HD double u53_open0(uint64_t x) { return (double)((x >> 11) + 1) * 0x1p-53; }  // (0,1]
SABC_HD double u53_mid(uint64_t x) { return ((double)(x >> 11) + 0.5) * 0x1p-53; }   // (0,1), for quantile transforms

// Box-Muller pair from one block: the normal behind the PRIOR draws (initialization only; addressed by explicit block numbers)
SABC_HD void normal_pair(const U64x2 w, double& z0, double& z1) {
    const double r = sqrt(-2.0 * det_log(u53_open0(w.a)));
    double sn, cs;
    det_sincos2pi(u53(w.b), sn, cs);
    z0 = r * cs; z1 = r * sn;
}

// ------------------------------------------------------------------------------------------------
// randn() on the hot path (proposals.jl:110 DE jitter, :42,54 RandomWalk, the models' noise): 256-layer ziggurat
// (Marsaglia & Tsang 2000; Julia's own randn() is a ziggurat too) on ONE 64-bit Philox word:
//   bits 0..7 layer i, bit 8 sign, bits 11..63 the 53-bit integer m;  x = m * ZIG_W[i]  (one exact conversion, one rounded product)
//   m < ZIG_K[i]  (98.5 %)  ->  +-x.   Otherwise the slow path takes one extra block e of the SAME stream:
//   i == 0: tail beyond R: xx = -log(U(e.a)) / R, yy = -log(U(e.b)), accept R + xx when 2 yy > xx^2, else next block, ...
//   i >= 1: wedge: accept x when ZIG_F[i] + U(e.a) (ZIG_F[i+1] - ZIG_F[i]) < exp(-x^2/2), else start again with the word e.b.
// Tables: tools/gen_ziggurat.py -> csrc/zig_tables.cuh.  A stream's blocks are consumed in order,
// so the result is a pure function of (seed, particle, sweep, kind) like every other draw.
// ------------------------------------------------------------------------------------------------
struct alignas(16) ZigEntry { unsigned long long k; double w; };
#if defined(__CUDACC__)
namespace zig_dev {
#define ZIG_TABLE_QUALIFIER static __device__ const
#include "zig_tables.cuh"
#undef ZIG_TABLE_QUALIFIER
}
#endif
namespace zig_host {
#define ZIG_TABLE_QUALIFIER static const
#include "zig_tables.cuh"
#undef ZIG_TABLE_QUALIFIER
}
#if defined(__CUDA_ARCH__)
#define SABC_ZIG_F(i) __ldg(&zig_dev::ZIG_F[i])
#else
#define SABC_ZIG_F(i) zig_host::ZIG_F[i]
#endif

struct ZigSlow { double z; uint32_t next; };
// the 1.5 % that miss the fast test, every argument by value so that the caller's Stream stays in registers
SABC_HD ZigSlow zig_slow(uint64_t w, uint32_t k0, uint32_t k1, uint32_t particle, uint32_t sweep_lo, uint32_t tag, uint32_t next) {
    for (;;) {
        const uint32_t idx = (uint32_t)w & 255u;
        const uint64_t m = w >> 11;
        const uint64_t sign = (w & 256u) << 55;
#if defined(__CUDA_ARCH__)
        const unsigned long long kk = zig_dev::ZIG_KW[idx].k; const double ww = zig_dev::ZIG_KW[idx].w;
#else
        const unsigned long long kk = zig_host::ZIG_KW[idx].k; const double ww = zig_host::ZIG_KW[idx].w;
#endif
        const double x = (double)m * ww;
        if (m < kk) return ZigSlow{bits_f64(f64_bits(x) | sign), next};
        U64x2 e = philox4x32_10(particle, sweep_lo, next++, tag, k0, k1);
        if (idx == 0) {
            for (;;) {
                const double xx = (-det_log(u53_open0(e.a))) * ZIG_INV_R;
                const double yy = -det_log(u53_open0(e.b));
                if (yy + yy > xx * xx) return ZigSlow{bits_f64(f64_bits(ZIG_R + xx) | sign), next};
                e = philox4x32_10(particle, sweep_lo, next++, tag, k0, k1);
            }
        }
        const double f0 = SABC_ZIG_F(idx), f1 = SABC_ZIG_F(idx + 1);
        const double y = f0 + u53(e.a) * (f1 - f0);
        if (y < det_exp((-0.5 * x) * x)) return ZigSlow{bits_f64(f64_bits(x) | sign), next};
        w = e.b;
    }
}
// out-of-line form for models with many call sites (an ABI call saves the live registers to local memory: measured on the fused
// Gaussian kernel as 800 MB of extra L2 traffic per half-sweep, so the default is the inline form)
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
ZigSlow zig_slow_call(uint64_t w, uint32_t k0, uint32_t k1, uint32_t particle, uint32_t sweep_lo, uint32_t tag, uint32_t next) {
    return zig_slow(w, k0, k1, particle, sweep_lo, tag, next);
}
// one normal from the word w; the slow path continues on st
template <bool INLINE_SLOW = true>
SABC_HD double zig_normal(uint64_t w, Stream& st) {
    const uint32_t idx = (uint32_t)w & 255u;
    const uint64_t m = w >> 11;
#if defined(__CUDA_ARCH__)
    // the layer is random per lane: from a shared-memory copy a warp's 32 lookups cost a few bank conflicts, through L1 up to 32 tag
    // wavefronts -- the fused Gaussian kernels are bound by exactly those (profiles/), so their CTAs stage the 4 KB table
    ulonglong2 e;
    if (st.zig_smem) asm("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(e.x), "=l"(e.y) : "r"(st.zig_smem + idx * 16u));
    else e = __ldg(reinterpret_cast<const ulonglong2*>(&zig_dev::ZIG_KW[idx]));
    const unsigned long long kk = e.x; const double ww = __longlong_as_double((long long)e.y);
#else
    const unsigned long long kk = zig_host::ZIG_KW[idx].k; const double ww = zig_host::ZIG_KW[idx].w;
#endif
    const double x = (double)m * ww;
    if (m < kk) return bits_f64(f64_bits(x) | ((w & 256u) << 55));
    const ZigSlow s = INLINE_SLOW ? zig_slow(w, st.k0, st.k1, st.particle, st.sweep_lo, st.tag, st.next)
                                  : zig_slow_call(w, st.k0, st.k1, st.particle, st.sweep_lo, st.tag, st.next);
    st.next = s.next;
    return s.z;
}
#if defined(__CUDACC__)
// copy the layer table into the CTA's shared memory (all threads of the CTA; the caller synchronises before the first draw)
SABC_D uint32_t stage_zig(ZigEntry* s_zig) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_zig[i] = zig_dev::ZIG_KW[i];
    return (uint32_t)__cvta_generic_to_shared(s_zig);
}
#endif
// two normals from the next block of st (stands in for two randn() calls)
template <bool INLINE_SLOW = true>
SABC_HD void normal2(Stream& st, double& z0, double& z1) {
    const U64x2 w = st.draw();
    z0 = zig_normal<INLINE_SLOW>(w.a, st);
    z1 = zig_normal<INLINE_SLOW>(w.b, st);
}
SABC_HD double normal1(Stream& st) {
    const U64x2 w = st.draw();
    return zig_normal(w.a, st);
}

// correctly rounded reciprocal: 1.0/x on the host, the cheaper MUFU-seeded __drcp_rn on the device (same bits)
SABC_HD double drcp(double x) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}

// 1/k, correctly rounded, for the inversion loop: a 64-entry table on the device (the loop runs with few live lanes,
// so a broadcast-style constant read beats the ~10-instruction reciprocal), the division itself elsewhere
#if defined(__CUDACC__)
__constant__ double c_rcp_int[64] = {
    0.0, 1.0 / 1, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13,
    1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25,
    1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37,
    1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41, 1.0 / 42, 1.0 / 43, 1.0 / 44, 1.0 / 45, 1.0 / 46, 1.0 / 47, 1.0 / 48, 1.0 / 49,
    1.0 / 50, 1.0 / 51, 1.0 / 52, 1.0 / 53, 1.0 / 54, 1.0 / 55, 1.0 / 56, 1.0 / 57, 1.0 / 58, 1.0 / 59, 1.0 / 60, 1.0 / 61,
    1.0 / 62, 1.0 / 63};
#endif
SABC_HD double rcp_int(int k) {
#if defined(__CUDA_ARCH__)
    return k < 64 ? c_rcp_int[k] : __drcp_rn((double)k);
#else
    return 1.0 / (double)k;
#endif
}

// ln(y) for a positive normal y with |error| <= 2e-9: exponent split, mantissa moved to [sqrt(1/2), sqrt(2)), then a
// degree-10 polynomial in t = m - 1 (Chebyshev interpolant of ln(1+t)/t; measured maximum error 1.62e-9).  No division:
// about a third of det_log's instructions.  NOT part of the numerical specification -- it only feeds ptrs_filter().
SABC_HD double approx_log(double y) {
    uint64_t b = f64_bits(y);
    uint32_t top = (uint32_t)(b >> 32);
    int k = (int)(top >> 20) - 1023;
    top &= 0x000fffffu;
    const uint32_t bump = (top + 0x95f64u) & 0x100000u;
    k += (int)(bump >> 20);
    b = ((uint64_t)(top | (bump ^ 0x3ff00000u)) << 32) | (b & 0xffffffffULL);
    const double t = bits_f64(b) - 1.0;
    double q = -0x1.313359b9a714ep-4;
    q = dfma(q, t, 0x1.0647858f47021p-3);  q = dfma(q, t, -0x1.0fb035e36d8d6p-3);
    q = dfma(q, t, 0x1.22cf0fbed0be2p-3);  q = dfma(q, t, -0x1.5423b015df288p-3);
    q = dfma(q, t, 0x1.999e867edab4ap-3);  q = dfma(q, t, -0x1.000423b72b7d5p-2);
    q = dfma(q, t, 0x1.55555d65bf746p-2);  q = dfma(q, t, -0x1.fffff847a7034p-2);
    q = dfma(q, t, 0x1.fffffffab3c37p-1);
    return dfma((double)k, 0x1.62e42fefa39efp-1, t * q);
}

// Cheap decision of the exact PTRS acceptance test  log(num/den) <= -lam + k log(lam) - log(k!)  (ptrs_exact below).
// With x = k+1 and Stirling's series for log Gamma(x) the difference rhs - lhs is
//   T = (x - lam) - log(sqrt(2 pi)) - [1/(12x) - 1/(360x^3) + 1/(1260x^5)] - k (ln x - ln lam) - ln(x)/2 - ln(num) + ln(den),
// which approx_log evaluates to within (2k + 2.5) * 1.62e-9 + 1/(1680 x^7) + rounding; the spec'd test itself carries
// about 1e-15 * k * ln(lam) of rounding.  Returns +1 (the exact test accepts) or -1 (it rejects) when |T| exceeds a bound
// E = 1e-6 + 5e-9 k that covers all of these with margin, 0 (undecided: run the exact test) otherwise.  About one slow-path
// attempt in 10^3..10^4 is undecided, so warps almost never execute the three det_log + det_logfact of the exact test;
// the decisions -- and therefore every Poisson draw -- are unchanged (tools/check_ptrs_filter.cpp compares them on
// >10^9 attempts on the CPU; the GPU parity tests compare the draws with the CPU restatement, which has no filter).
SABC_HD int ptrs_filter(double lam, double kf, double num, double den, double& T, double& E) {
    T = 0.0; E = 0.0;
    if (!(kf >= 2.0) || !(kf < 1e12) || !(num > 0x1p-1000) || !(den > 0x1p-1000) || !(den < 0x1p1000)) return 0;
    const double x = kf + 1.0;
    const double lx = approx_log(x), ll = approx_log(lam);
    const double rx = drcp(x), rx2 = rx * rx;
    const double corr = rx * dfma(-rx2, dfma(-rx2, 1.0 / 1260.0, 1.0 / 360.0), 1.0 / 12.0);
    T = (x - lam) - 0x1.d67f1c864beb5p-1;
    T = T - corr;
    T = dfma(-kf, lx - ll, T);
    T = dfma(-0.5, lx, T);
    T = T - (approx_log(num) - approx_log(den));
    E = dfma(kf, 5e-9, 1e-6);
    return T > E ? 1 : (T < -E ? -1 : 0);
}
SABC_HD int ptrs_filter(double lam, double kf, double num, double den) {
    double T, E; return ptrs_filter(lam, kf, num, den, T, E);
}

// The PTRS constants whose low mantissa word is not zero cost two UMOV each as immediates, every loop trip; from the
// constant bank they are one LDCU.64 (or half an LDCU.128).  Same literals on the host.
#if defined(__CUDACC__)
__constant__ double c_ptrs[14] = {0.931, 2.53, -0.059, 0.02483, 0.43, 0.9277, 3.6224, 0.07, 0.013, 3.4, 1.1239, 1.1328,
                                  0x1.d67f1c864beb5p-1, 0.0};
#endif
#if defined(__CUDA_ARCH__) && !defined(SABC_NO_CONST_BANK)
#define SABC_PC(i, v) c_ptrs[i]
#else
#define SABC_PC(i, v) (v)
#endif

#if defined(__CUDACC__)
// log2 on the special-function unit: ONE MUFU.LG2 (arguments here are normal floats, so the flush-to-zero form needs no
// pre-scaling).  Documented accuracy of lg2.approx: absolute error <= 2^-21.41 on [0.5, 2], <= 2 ulp elsewhere.
SABC_D float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SABC_D float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// First-level filter, device only: the same T in single precision with the logarithms on the MUFU unit -- about 35
// instructions against ~180 of ptrs_filter and ~330 of the exact test.  The k ln(x/lam) term is taken as k ln(1 + D/lam)
// with D = x - lam exact in FP64, so its error is k times that of ONE logarithm near 1:
//   MUFU 2^-21.41 ln2 = 2.5e-7, the argument (three roundings and a 1-ulp reciprocal) <= 3.6e-7 on [0.5, 2]  ->  6.1e-7 k
//   (q in (2, 256]: 2 ulp of |log2 q| <= 8 plus the argument: 1.5e-6 k); the FP32 products and sums of the two large
//   terms, each at most 2|D| on q <= 2: < 4.1e-7 |D|; everything else (ln x / 2, ln num - ln den with |log2| < 128,
//   Stirling truncated after 1/(360 x^3) with x >= 3) < 5e-5.
// Bound used: E = 2e-4 + k (q <= 2 ? 1e-6 : 3e-6) + 5e-7 |D|.  Outside the guarded ranges, or when |T| <= E, it returns 0 and the
// FP64 filter above decides (or passes on to the exact test).  A NaN anywhere fails both comparisons and falls through.
// tests/test_gpu_hooks.py::test_ptrs_filters_agree_with_exact_test checks decisions and the margin |T - T_exact| / E on the GPU.
SABC_D int ptrs_filter_mufu(double lam, double kf, double num, double den, float& T, float& E) {
    T = 0.0f; E = 0.0f;
    if (!(kf >= 2.0) || !(kf < 1e7) || !(num > 0x1p-100) || !(den > 0x1p-100) || !(den < 0x1p100)) return 0;
    const double x = kf + 1.0;
    const double D = x - lam;
    const float xf = __double2float_rn(x), kff = __double2float_rn(kf);
    const float q = __fadd_rn(1.0f, __fmul_rn(__double2float_rn(D), mufu_rcp(__double2float_rn(lam))));
    if (!(q >= 0.5f) || !(q <= 256.0f)) return 0;
    const float rx = mufu_rcp(xf);
    const float corr = __fmul_rn(rx, __fmaf_rn(__fmul_rn(rx, rx), -1.0f / 360.0f, 1.0f / 12.0f));
    // ln2 * [ k log2 q + log2(x)/2 + log2 num - log2 den ]
    float L = __fmaf_rn(kff, mufu_lg2(q), __fmul_rn(0.5f, mufu_lg2(xf)));
    L = __fadd_rn(L, __fsub_rn(mufu_lg2(__double2float_rn(num)), mufu_lg2(__double2float_rn(den))));
    const float Df = __double2float_rn(D - c_ptrs[12]);
    T = __fsub_rn(__fsub_rn(Df, corr), __fmul_rn(0.693147180559945f, L));
    E = __fmaf_rn(kff, q <= 2.0f ? 1e-6f : 3e-6f, __fmaf_rn(fabsf(Df), 5e-7f, 2e-4f));
    return T > E ? 1 : (T < -E ? -1 : 0);
}
#endif

// Hoermann's PTRS (1993) for lam >= 10, split into the candidate with its cheap tests and the exact test.
// ptrs_candidate: 1 = accept kf, 0 = reject, 2 = the exact test on (kf, num, den) decides.
SABC_HD int ptrs_candidate(double lam, const U64x2 w, double& kf, double& num, double& den) {
    const double slam = sqrt(lam);
    const double b = SABC_PC(0, 0.931) + SABC_PC(1, 2.53) * slam;
    const double a = SABC_PC(2, -0.059) + SABC_PC(3, 0.02483) * b;
    const double U = u53(w.a) - 0.5, V = u53(w.b);
    const double us = 0.5 - fabs(U);
    const double r = drcp(us);
    kf = floor(((2.0 * a) * r + b) * U + lam + SABC_PC(4, 0.43));
    if (us >= SABC_PC(7, 0.07) && (SABC_PC(5, 0.9277) - V) * (b - 2.0) >= SABC_PC(6, 3.6224)) return 1;
    if (kf < 0.0 || (us < SABC_PC(8, 0.013) && V > us)) return 0;
    const double bm = b - SABC_PC(9, 3.4);
    num = V * (SABC_PC(10, 1.1239) * bm + SABC_PC(11, 1.1328));
    den = bm * ((a * r) * r + b);
    return 2;
}
SABC_HD bool ptrs_exact(double lam, double kf, double num, double den) {
    const double lhs = det_log(num / den);
    const double rhs = (-lam + kf * det_log(lam)) - det_logfact(kf);
    return lhs <= rhs;
}

// Poisson(lam): one-uniform sequential-search inversion below 10, PTRS above, with the acceptance tests
// rearranged to one reciprocal and one logarithm of a quotient.  One ATTEMPT consumes one Philox block (none when
// lam <= 0) and either returns a count or rejects; this is the unit the SIR kernel interleaves across lanes so that a
// rejection in one lane does not stall the accepted lanes of its warp.  The two filters only shortcut the exact test.
// poisson_attempt_d returns the count as an integer-valued double (what PTRS computes anyway); poisson_attempt converts.
SABC_HD bool poisson_attempt_d(double lam, Stream& st, double& k_out) {
    if (!(lam > 0.0)) { k_out = 0.0; return true; }
    const U64x2 w = st.draw();
    if (lam < 10.0) {
        const double U = u53(w.a);
        double p = det_exp(-lam), F = p;
        int k = 0;
#if defined(__CUDA_ARCH__)
        // same arithmetic as the loop below while the 1/k table covers k (lam < 10: in practice always); no branch on k inside
#pragma unroll 4
        while (U > F && k < 63) { k++; p = (p * lam) * c_rcp_int[k]; F = F + p; }
#endif
        while (U > F && k < 1024) { k++; p = (p * lam) * rcp_int(k); F = F + p; }         // p_k = p_{k-1} λ (1/k)
        k_out = (double)k;
        return true;
    }
    double kf, num = 0.0, den = 0.0;
    int s = ptrs_candidate(lam, w, kf, num, den);
    if (s == 2) {
#if !defined(SABC_NO_PTRS_FILTER)
        int dec = 0;
#if defined(__CUDA_ARCH__) && !defined(SABC_NO_MUFU_FILTER)
        { float T1, E1; dec = ptrs_filter_mufu(lam, kf, num, den, T1, E1); }
        if (dec == 0)
#endif
            dec = ptrs_filter(lam, kf, num, den);
#if defined(SABC_FILTER_STAT)
        SABC_FILTER_STAT(dec);
#endif
        s = dec != 0 ? (dec > 0) : (int)ptrs_exact(lam, kf, num, den);
#else
        s = (int)ptrs_exact(lam, kf, num, den);
#endif
    }
    if (s == 1) k_out = kf;
    return s == 1;
}
SABC_HD bool poisson_attempt(double lam, Stream& st, int64_t& k_out) {
    double kd;
    if (!poisson_attempt_d(lam, st, kd)) return false;
    k_out = (int64_t)kd;
    return true;
}
SABC_HD int64_t poisson(double lam, Stream& st) {
    int64_t k;
    while (!poisson_attempt(lam, st, k)) {}
    return k;
}

// exact, order-independent accumulation of u in [0,1]: u*2^62 split into two 31-bit limbs (§3.4)
SABC_HD void u_limbs(double u, uint32_t& hi, uint32_t& lo) {
    double c = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    if (c != c) c = 0.0;
    const uint64_t q = (uint64_t)(c * 0x1p62);
    hi = (uint32_t)(q >> 31); lo = (uint32_t)(q & 0x7fffffffULL);
}
SABC_HD double limbs_to_sum(uint64_t hi, uint64_t lo) { return ((double)hi * 2147483648.0 + (double)lo) * 0x1p-62; }

}  // namespace sabc
