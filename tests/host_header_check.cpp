// host_header_check.cpp -- TEST CODE (links the oracle).  The product's sampler and model headers are written for host and
// device (SABC_HD); compiled here with g++ they must reproduce the oracle bit for bit: Poisson draws (inversion, PTRS with
// the FP64 acceptance filter in front of the exact test) and whole SIR tau-leap simulations.  This pins the header logic
// on the CPU; the GPU parity tests pin the device build of the same headers (which adds the MUFU filter and the tables).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../simulatedannealingabc.jl_b200/csrc/plugin.cuh"
extern "C" {
int orc_model_simulate(int32_t model_id, int32_t d, int32_t s, const double* model_par, int32_t n_model_par, const double* theta,
                       uint64_t seed, uint32_t particle, uint64_t sweep, double* rho_out);
int64_t orc_poisson(double lam, uint64_t seed, uint32_t particle, uint64_t sweep, uint32_t* block_io);
}
using namespace sabc;

int main(int argc, char** argv) {
    const int n_sim = argc > 1 ? atoi(argv[1]) : 4000, n_draw = argc > 2 ? atoi(argv[2]) : 400000;
    long bad = 0;
    // Poisson draws over lambda in (0, 1e6]
    for (int i = 0; i < n_draw; ++i) {
        Stream r(9, (uint32_t)i, 0, KIND_PRIOR);
        const U64x2 w = r.draw();
        const double e = 7.0 * u53(w.a) - 1.0;                       // log10(lambda) in [-1, 6)
        double lam = 1.0; for (int k = 0; k < 40; ++k) lam = lam * (1.0 + e * 0.0575646273248511);   // ~10^e, any positive value will do
        if (i % 50 == 0) lam = 10.0 + 1e-9 * (double)(i % 7);
        Stream st(77, (uint32_t)i, 5, KIND_MODEL);
        const int64_t k = poisson(lam, st);
        uint32_t blk = 0;
        const int64_t ko = orc_poisson(lam, 77, (uint32_t)i, 5, &blk);
        if (k != ko || blk != st.next) { if (bad < 5) printf("poisson mismatch lam=%g: %lld vs %lld\n", lam, (long long)k, (long long)ko); bad++; }
    }
    // SIR tau-leap simulations: prior draws, a concentrated cloud, clamped initial conditions
    const double par[6] = {1e5, 50, 1.0, 20000.0, 1500.0, 25.0};
    ModelPar mp{};
    for (int i = 0; i < 6; ++i) mp.v[i] = par[i];
    for (int i = 0; i < n_sim; ++i) {
        Stream r(123, (uint32_t)i, 0, KIND_PRIOR);
        const U64x2 a = r.draw(), b = r.draw();
        double th[4] = {0.1 + 0.9 * u53(a.a), 0.05 + 0.45 * u53(a.b), 0.001 + 0.049 * u53(b.a), 0.2 + 0.8 * u53(b.b)};
        if (i % 3 == 0) { th[0] = 0.3 * (0.9 + 0.2 * u53(a.a)); th[1] = 0.1 * (0.9 + 0.2 * u53(a.b)); th[2] = 0.01; }
        if (i % 7 == 0) th[2] = 1.2;
        if (i % 11 == 0) th[2] = -0.1;
        double rho[3], ro[3];
        Stream st(77, (uint32_t)i, 5, KIND_MODEL);
        const double(&thr)[4] = th;
        SirTauLeap::sim(thr, mp, st, rho);
        orc_model_simulate(3, 4, 3, par, 6, th, 77, (uint32_t)i, 5, ro);
        if (memcmp(rho, ro, sizeof rho) != 0) { if (bad < 10) printf("sir mismatch %d\n", i); bad++; }
    }
    printf("%d poisson draws, %d sir simulations, %ld mismatches\n", n_draw, n_sim, bad);
    return bad != 0;
}
