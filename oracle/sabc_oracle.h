/*
 * sabc_oracle.h -- CPU ORACLE for the SABC population-update path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithm
 * (SimulatedAnnealingABC.jl v0.4.0: src/SimulatedAnnealingABC.jl, src/cdf_estimators.jl,
 * src/proposals.jl).  It is used by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py as the CHECKER and the timed CPU baseline.  Nothing under
 * simulatedannealingabc.jl_b200/ (the product) includes, links or calls it.
 *
 * PARITY UNPINNED: the reference is Julia and cannot run in this image (no julia binary, no
 * network); its own tests hold no golden values for this path (test/runtests.jl:9-29 are
 * inequalities only).  The oracle is pinned by (i) those inequality tests, (ii) the published
 * Random123 Philox4x32-10 known-answer vectors, (iii) restatement-derived known answers
 * (SURVEY.md App. F), (iv) libm/mpmath agreement of the deterministic math, and (v) the
 * analytic conjugate posterior of the 1-D Gaussian config.  Third-party arithmetic that is not
 * under /root/reference (Interpolations.jl ^0.15 monotonic linear interpolation + Flat
 * extrapolation, Roots.jl ^2.1 find_zero, StatsBase ^0.34 sample/mean/cov, Distributions ^0.25
 * logpdf/rand) is restated from its published semantics; see DESIGN.md.
 */
#ifndef SABC_ORACLE_H
#define SABC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_ALG_SINGLE_EPS = 0, ORC_ALG_MULTI_EPS = 1 };
enum { ORC_PROP_DE = 0, ORC_PROP_STRETCH = 1, ORC_PROP_RW = 2 };
enum { ORC_PRIOR_UNIFORM = 0, ORC_PRIOR_NORMAL = 1, ORC_PRIOR_EXPONENTIAL = 2, ORC_PRIOR_LOGNORMAL = 3, ORC_PRIOR_GAMMA = 4, ORC_PRIOR_BETA = 5,
       ORC_PRIOR_CAUCHY = 6, ORC_PRIOR_LAPLACE = 7, ORC_PRIOR_WEIBULL = 8, ORC_PRIOR_INVGAMMA = 9 };
enum { ORC_MODEL_GAUSS_MEAN = 0, ORC_MODEL_GAUSS_SAMPLE = 1, ORC_MODEL_LOGISTIC = 2, ORC_MODEL_SIR = 3, ORC_MODEL_SIR_GILLESPIE = 4 };

typedef struct orc_config {
    int64_t n_particles;
    int32_t n_para;          /* d = length(prior) */
    int32_t n_stats;         /* s = length(f_dist(theta)) */
    int32_t algorithm;       /* ORC_ALG_* */
    int32_t proposal;        /* ORC_PROP_* */
    double  prop_par[2];     /* DE: gamma0, sigma_gamma | Stretch: a | RW: beta */
    double  v;
    double  delta;
    int64_t resample;
    uint64_t seed;
    int32_t model_id;
    int32_t n_model_par;
    const double* model_par;
    const int32_t* prior_kind;   /* n_para entries */
    const double* prior_par;     /* 2*n_para entries */
    int32_t ecdf_max_knots;      /* 0: full ECDF; K >= 2: K rank-uniform quantiles (compressed mode of the product) */
} orc_config;

typedef struct orc_engine orc_engine;

/* ---- unit functions (each restates one step; file:line of the reference in the .c) ---- */
void   orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_log(double x);
double orc_exp(double x);
void   orc_sincos2pi(double u, double* s, double* c);
double orc_logfact(double k);
void   orc_normal_pair(uint64_t a, uint64_t b, double* z0, double* z1);
int64_t orc_poisson(double lam, uint64_t seed, uint32_t particle, uint64_t sweep, uint32_t* block_io);
/* one ziggurat normal from the next block(s) of the model stream (seed, particle, sweep); *block_io advances */
void   orc_normal_stream(uint64_t seed, uint32_t particle, uint64_t sweep, int32_t n_pairs, double* out);
double orc_zig_normal(uint64_t seed, uint32_t particle, uint64_t sweep, uint32_t* block_io);
double orc_treesum(const double* x, int64_t n);

int64_t orc_ecdf_build(const double* x, int64_t n, double* knots_out /* n+2 */);
void   orc_ecdf_eval(const double* knots, int64_t L, const double* rho, int64_t m, double* u_out);
void   orc_accept_step(int64_t m, int32_t s, const double* u_old /* m*s col-major */, const double* u_new,
                       const double* eps, int32_t n_eps, const double* dlogprior, const double* log_factor,
                       const double* uniform, uint8_t* accept_out);
double orc_eps_single(double ubar, double v);
double orc_eps_single_bisect(double ubar, double v);
int    orc_eps_multi(const double* ubar, int32_t s, double v, double* eps_out);
void   orc_resample_weights(const double* u /* n*s col-major */, int64_t n, int32_t s, const double* ubar,
                            double delta, double* w_out, uint64_t* q_out);
void   orc_resample_indices(const uint64_t* q, int64_t n, uint64_t seed, uint64_t resample_count, int64_t* idx_out);
void   orc_exact_mean_u(const double* u, int64_t n, double* mean_out);
double orc_lgamma(double x);
void   orc_prior_rand(int32_t d, const int32_t* kind, const double* par, uint64_t seed, uint32_t particle, double* theta_out);
double orc_prior_logpdf(int32_t d, const int32_t* kind, const double* par, const double* theta);
int    orc_model_simulate(int32_t model_id, int32_t d, int32_t s, const double* model_par, int32_t n_model_par,
                          const double* theta, uint64_t seed, uint32_t particle, uint64_t sweep, double* rho_out);
int    orc_propose(int32_t proposal, const double* prop_par, int32_t d, const double* theta_i,
                   const double* inactive /* M*d row-major */, int64_t M, const double* chol /* d*d row-major or sd */,
                   uint64_t seed, uint32_t particle, uint64_t sweep, double* theta_out, double* log_factor_out);

/* ---- engine (restates initialization() and update_population!()) ---- */
int  orc_create(orc_engine** out, const orc_config* cfg);
int  orc_destroy(orc_engine* e);
int  orc_init(orc_engine* e);
int  orc_update(orc_engine* e, int64_t n_simulation, int64_t checkpoint_history, double* seconds_out);
int  orc_get_population(orc_engine* e, double* theta /* N*d col-major */, double* u /* N*s */, double* rho /* N*s */);
int  orc_set_population(orc_engine* e, const double* theta, const double* u, const double* rho,
                        const double* eps, const int64_t counters[4]);
int  orc_get_state(orc_engine* e, double* eps, int64_t counters[4]);
int64_t orc_history_len(orc_engine* e);
int  orc_get_history(orc_engine* e, double* eps_h, double* u_h, double* rho_h);
int64_t orc_get_ecdf(orc_engine* e, int32_t stat, double* knots_out);
int  orc_set_ecdf(orc_engine* e, int32_t stat, const double* knots, int64_t L);
int  orc_num_threads(void);
void orc_set_num_threads(int n);   /* launchers such as torchrun export OMP_NUM_THREADS=1 */
const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
