// check_ptrs_filter.cpp -- CPU check that ptrs_filter() (csrc/philox.cuh) never changes a decision of the exact PTRS
// acceptance test, and how often it is undecided.  Built twice from this one file (tools/check_ptrs_filter.sh):
//   g++ ... -DSABC_NO_PTRS_FILTER -Dsabc=sabc_exact -c   -> the spec'd sampler (own namespace: inline functions differ)
//   g++ -O2 -ffp-contract=off -mfma -fopenmp                           -> the sampler with the filter + main()
// Usage: check_ptrs_filter [attempts per lambda, default 2e6]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#ifndef SABC_NO_PTRS_FILTER
static thread_local long long g_stat[3];
#define SABC_FILTER_STAT(dec) g_stat[(dec) + 1]++
#endif
#include "../simulatedannealingabc.jl_b200/csrc/philox.cuh"

#ifdef SABC_NO_PTRS_FILTER
extern "C" int attempt_exact(double lam, uint64_t seed, uint32_t particle, long long* k) {
    sabc::Stream st(seed, particle, 0, sabc::KIND_MODEL);
    int64_t kk = -1; const bool ok = sabc::poisson_attempt(lam, st, kk); *k = ok ? kk : -1; return ok;
}
#else
extern "C" int attempt_exact(double lam, uint64_t seed, uint32_t particle, long long* k);
static int attempt_filtered(double lam, uint64_t seed, uint32_t particle, long long* k) {
    sabc::Stream st(seed, particle, 0, sabc::KIND_MODEL);
    int64_t kk = -1; const bool ok = sabc::poisson_attempt(lam, st, kk); *k = ok ? kk : -1; return ok;
}
int main(int argc, char** argv) {
    const long long per = argc > 1 ? atoll(argv[1]) : 2000000LL;
    std::vector<double> lams;
    for (double l = 10.0; l < 3e9; l *= 1.37) { lams.push_back(l); lams.push_back(floor(l) + 0.5); lams.push_back(nextafter(l, 0.0)); }
    lams.push_back(10.0); lams.push_back(16.0); lams.push_back(17.0); lams.push_back(1e12); lams.push_back(1e15);
    long long bad = 0, tot = 0, slow = 0, und = 0;
    printf("%14s %12s %12s %12s %10s\n", "lambda", "attempts", "slow-path", "undecided", "mismatch");
    for (size_t li = 0; li < lams.size(); ++li) {
        const double lam = lams[li];
        long long b = 0, s = 0, u = 0;
        #pragma omp parallel reduction(+ : b, s, u)
        {
            g_stat[0] = g_stat[1] = g_stat[2] = 0;
            #pragma omp for schedule(static)
            for (long long i = 0; i < per; ++i) {
                long long k0, k1;
                const int a0 = attempt_exact(lam, 0x5abc0000ULL + li, (uint32_t)i, &k0);
                const int a1 = attempt_filtered(lam, 0x5abc0000ULL + li, (uint32_t)i, &k1);
                if (a0 != a1 || k0 != k1) b++;
            }
            s += g_stat[0] + g_stat[1] + g_stat[2]; u += g_stat[1];
        }
        if (li % 6 == 0 || b) printf("%14.6g %12lld %12lld %12lld %10lld\n", lam, per, s, u, b);
        bad += b; tot += per; slow += s; und += u;
    }
    printf("total: %lld attempts, %lld reached the exact test (%.1f %%), %lld undecided by the filter (%.4f %% of those), %lld mismatches\n",
           tot, slow, 100.0 * slow / tot, und, 100.0 * und / (slow ? slow : 1), bad);
    return bad ? 1 : 0;
}
#endif
