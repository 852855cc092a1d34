// host_header_check.cpp -- TEST CODE (links the oracle).  The product's sampler and model headers are written for host and
// device (SABC_HD); compiled here with g++ they must reproduce the oracle bit for bit: Poisson draws (inversion, PTRS with
// the FP64 acceptance filter in front of the exact test), whole SIR tau-leap simulations, prior draws and log-densities of
// every family.  This pins the header logic
// on the CPU; the GPU parity tests pin the device build of the same headers (which adds the MUFU filter and the tables).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../simulatedannealingabc.jl_b200/csrc/plugin.cuh"
extern "C" {
int orc_model_simulate(int32_t model_id, int32_t d, int32_t s, const double* model_par, int32_t n_model_par, const double* theta,
                       uint64_t seed, uint32_t particle, uint64_t sweep, double* rho_out);
int64_t orc_poisson(double lam, uint64_t seed, uint32_t particle, uint64_t sweep, uint32_t* block_io);
void orc_prior_rand(int32_t d, const int32_t* kind, const double* par, uint64_t seed, uint32_t particle, double* theta_out);
double orc_prior_logpdf(int32_t d, const int32_t* kind, const double* par, const double* theta);
int orc_propose(int32_t proposal, const double* pp, int32_t d, const double* th, const double* P, int64_t M, const double* chol, uint64_t seed,
                uint32_t particle, uint64_t sweep, double* out, double* lf);
}
struct RowMajor { const double* p; int d; double operator()(int c, int64_t i) const { return p[i * d + c]; } };
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
    // ziggurat normals through the models that draw them (Gaussian mean, Gaussian sample, logistic growth) and through the DE / RandomWalk
    // proposals: fast path, wedge and tail of the sampler on the host build of the product headers against the oracle
    {
        const double pg[2] = {1.0, 0.31622776601683794}, ps[5] = {10, 1.0, 2.0, 42.5, 1.0};
        double pl[22] = {10.0, 20.0};
        for (int t = 0; t < 20; ++t) pl[2 + t] = 15.0 + 11.0 * t;
        ModelPar mg{}, ms{}, ml{};
        for (int i = 0; i < 2; ++i) mg.v[i] = pg[i];
        for (int i = 0; i < 5; ++i) ms.v[i] = ps[i];
        for (int i = 0; i < 22; ++i) ml.v[i] = pl[i];
        const int n_norm = 60 * n_sim;
        for (int i = 0; i < n_norm; ++i) {
            Stream r(321, (uint32_t)i, 0, KIND_PRIOR);
            const U64x2 a = r.draw(), b = r.draw();
            { double th[1] = {4.0 * u53(a.a) - 2.0}, rho[1], ro[1]; Stream st(55, (uint32_t)i, 9, KIND_MODEL); const double(&t1)[1] = th;
              GaussMean::sim(t1, mg, st, rho); orc_model_simulate(0, 1, 1, pg, 2, th, 55, (uint32_t)i, 9, ro);
              if (memcmp(rho, ro, sizeof rho) != 0) { if (bad < 10) printf("gauss_mean mismatch %d\n", i); bad++; } }
            { double th[2] = {6.0 * u53(a.a) - 3.0, 0.1 + 2.0 * u53(a.b)}, rho[2], ro[2]; Stream st(55, (uint32_t)i, 9, KIND_MODEL); const double(&t2)[2] = th;
              GaussSample<2, 2>::sim(t2, ms, st, rho); orc_model_simulate(1, 2, 2, ps, 5, th, 55, (uint32_t)i, 9, ro);
              if (memcmp(rho, ro, sizeof rho) != 0) { if (bad < 10) printf("gauss_sample mismatch %d\n", i); bad++; } }
            if (i % 4 == 0) { double th[3] = {u53(a.a), 50.0 + 450.0 * u53(a.b), 0.5 * u53(b.a)}, rho[20], ro[20]; Stream st(55, (uint32_t)i, 9, KIND_MODEL);
              const double(&t3)[3] = th;
              Logistic::sim(t3, ml, st, rho); orc_model_simulate(2, 3, 20, pl, 22, th, 55, (uint32_t)i, 9, ro);
              if (memcmp(rho, ro, sizeof rho) != 0) { if (bad < 10) printf("logistic mismatch %d\n", i); bad++; } }
            if (i % 8 == 0) {                       // proposals over a small inactive half
                double P[16 * 3], th[3] = {u53(a.a), u53(a.b), u53(b.a)}, out[3], oo[3], lf, lo;
                for (int k = 0; k < 48; ++k) P[k] = u53(Stream(7, (uint32_t)k, 0, KIND_PRIOR).block(0).a);
                const double pp[2] = {2.38 / sqrt(6.0), 1e-5}, chol[9] = {0.5, 0, 0, 0.1, 0.4, 0, -0.2, 0.05, 0.3};
                const double(&t3)[3] = th; double(&o3)[3] = out;
                const CtrlWords cw = ctrl_words(99, (uint32_t)i, 4);
                propose_de<3>(t3, RowMajor{P, 3}, 16, pp[0], pp[1], cw, Stream(99, (uint32_t)i, 4, KIND_CTRL), o3, lf);
                orc_propose(0, pp, 3, th, P, 16, chol, 99, (uint32_t)i, 4, oo, &lo);
                if (memcmp(out, oo, sizeof out) != 0 || lf != lo) { if (bad < 10) printf("propose_de mismatch %d\n", i); bad++; }
                propose_rw<3>(t3, chol, 99, (uint32_t)i, 4, o3, lf);
                orc_propose(2, pp, 3, th, P, 16, chol, 99, (uint32_t)i, 4, oo, &lo);
                if (memcmp(out, oo, sizeof out) != 0 || lf != lo) { if (bad < 10) printf("propose_rw mismatch %d\n", i); bad++; }
            }
        }
        printf("%d gaussian / logistic simulations and proposals checked\n", n_norm);
    }
    // priors: every family, draws and log-densities (also off the support)
    const int n_prior = 20000;
    const int32_t kinds[5][6] = {{4, 5, 0, 1, 2, 3}, {5, 4, 4, 5, 4, 5}, {4, 4, 5, 5, 1, 0}, {6, 7, 8, 9, 8, 9}, {9, 8, 7, 6, 8, 9}};
    const double pars[5][12] = {{2.5, 0.8, 0.7, 3.0, -1.0, 3.0, 0.5, 2.0, 1.5, 0.0, 0.5, 0.8},
                                {4.0, 2.0, 0.4, 1.5, 1.0, 3.0, 1.0, 1.0, 30.0, 0.01, 0.05, 0.05},
                                {0.05, 1.0, 100.0, 2.0, 0.5, 0.5, 50.0, 60.0, 0.0, 1.0, 0.0, 1.0},
                                {0.5, 2.0, -1.0, 0.7, 1.7, 2.5, 3.0, 2.0, 0.6, 1.0, 0.5, 0.1},
                                {20.0, 5.0, 1.0, 3.0, 0.0, 10.0, -3.0, 0.01, 5.0, 0.2, 1.0, 1.0}};
    for (int set = 0; set < 5; ++set) {
        PriorSpec ps{};
        ps.n = 6;
        for (int c = 0; c < 6; ++c) { ps.kind[c] = kinds[set][c]; ps.p0[c] = pars[set][2 * c]; ps.p1[c] = pars[set][2 * c + 1]; }
        prior_prepare(ps);
        for (int i = 0; i < n_prior; ++i) {
            double th[6], to[6];
            prior_rand<6>(ps, 4242, (uint32_t)i, th);
            orc_prior_rand(6, kinds[set], pars[set], 4242, (uint32_t)i, to);
            if (memcmp(th, to, sizeof th) != 0) { if (bad < 10) printf("prior_rand mismatch set %d particle %d\n", set, i); bad++; continue; }
            if (i % 5 == 1) th[i % 6] = -th[i % 6];
            if (i % 5 == 2) th[i % 6] = th[i % 6] + 1.0;
            if (i % 97 == 0) th[i % 6] = 0.0;
            if (i % 89 == 0) th[i % 6] = 1.0;
            const double(&thr)[6] = th;
            const double lp = prior_logpdf<6>(ps, thr), lo = orc_prior_logpdf(6, kinds[set], pars[set], th);
            if (memcmp(&lp, &lo, 8) != 0 && !(lp != lp && lo != lo)) { if (bad < 10) printf("prior_logpdf mismatch set %d particle %d: %.17g vs %.17g\n", set, i, lp, lo); bad++; }
        }
    }
    printf("%d poisson draws, %d sir simulations, %d x 5 prior draws, %ld mismatches\n", n_draw, n_sim, n_prior, bad);
    return bad != 0;
}
