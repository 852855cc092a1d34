/*
 * sabc_b200.h -- C ABI of the B200-native SABC population engine (libsabc_b200.so).
 *
 * Drop-in boundary for the hot path of SimulatedAnnealingABC.jl v0.4.0.  Each entry point names
 * the reference interface it replaces (paths relative to the reference repository).  Julia binds
 * these with `ccall` (INTEGRATION.md shows the stubs); tests and bench.py bind them with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative SABC_ERR_* code; the message is available
 *     from sabc_last_error() (thread-local).  No C++ exception crosses the boundary.
 *   - all matrices are column-major FP64 exactly like Julia's: theta is N x d, u and rho are N x s.
 *   - the caller owns every host buffer it passes and keeps it alive for the duration of the call;
 *     the library owns all device memory, streams and communicators inside the handle.
 *   - calls on one handle are not re-entrant.  Calls block until their results are host-visible.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     SABC_ERR_CUDA.
 */
#ifndef SABC_B200_H
#define SABC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SABC_ABI_VERSION 2   /* 2: sabc_config ends with n_gpus, gpu_ids */

/* error classes; the first block mirrors the reference's error() sites */
#define SABC_OK                 0
#define SABC_ERR_NSIM_TOO_SMALL (-1)  /* src/SimulatedAnnealingABC.jl:155-156 */
#define SABC_ERR_BAD_V          (-2)  /* :261  */
#define SABC_ERR_BAD_DELTA      (-3)  /* :262  */
#define SABC_ERR_NEG_DISTANCE   (-4)  /* :185  */
#define SABC_ERR_UBAR_ZERO      (-5)  /* :107-109 */
#define SABC_ERR_BAD_ALGORITHM  (-6)  /* :462-464 */
#define SABC_ERR_BAD_PROPOSAL   (-7)  /* src/proposals.jl:30,96 */
#define SABC_ERR_NO_POSITIVE    (-8)  /* src/cdf_estimators.jl:33 maximum() of an empty column */
#define SABC_ERR_INVALID        (-20) /* bad argument / unknown model / dimension mismatch */
#define SABC_ERR_STATE          (-21) /* call order (e.g. update before init) */
#define SABC_ERR_CUDA           (-30)
#define SABC_ERR_NCCL           (-31)

enum { SABC_ALG_SINGLE_EPS = 0, SABC_ALG_MULTI_EPS = 1 };          /* algorithm = :single_eps | :multi_eps  (:453) */
enum { SABC_PROP_DE = 0, SABC_PROP_STRETCH = 1, SABC_PROP_RW = 2 }; /* src/proposals.jl:85,132,24 */
enum { SABC_PRIOR_UNIFORM = 0, SABC_PRIOR_NORMAL = 1, SABC_PRIOR_EXPONENTIAL = 2, SABC_PRIOR_LOGNORMAL = 3,
       SABC_PRIOR_GAMMA = 4 /* (shape, scale) */, SABC_PRIOR_BETA = 5, SABC_PRIOR_CAUCHY = 6 /* (mu, sigma) */,
       SABC_PRIOR_LAPLACE = 7 /* (mu, theta) */, SABC_PRIOR_WEIBULL = 8 /* (shape, scale) */,
       SABC_PRIOR_INVERSEGAMMA = 9 /* (shape, scale) */ };                                                         /* Distributions.jl */

typedef struct sabc_engine sabc_engine;

/* Configuration = the keyword arguments of sabc() (src/SimulatedAnnealingABC.jl:451-460) plus the
 * device-side plug-ins that replace the closure f_dist and the Distribution object. */
typedef struct sabc_config {
    int64_t  n_particles;     /* n_particles: GLOBAL particle count over all ranks */
    int32_t  n_para;          /* length(prior) */
    int32_t  n_stats;         /* length(f_dist(theta)) */
    int32_t  algorithm;       /* SABC_ALG_* */
    int32_t  proposal;        /* SABC_PROP_* */
    double   prop_par[2];     /* DE: gamma0, sigma_gamma | Stretch: a | RW: beta   (src/proposals.jl:29,87-99,135) */
    double   v;               /* annealing speed (:455) */
    double   delta;           /* resampling intensity (:455) */
    int64_t  resample;        /* resample after this many accepts (:454) */
    uint64_t seed;            /* Philox key */
    const char* model_name;   /* registered device model ("gauss_mean", "gauss_sample", "logistic", "sir_tauleap", ...) */
    const double* model_par;  /* model parameter blob (y_obs etc.; the args.../kwargs... of f_dist) */
    int32_t  n_model_par;
    int32_t  device;          /* CUDA device ordinal; -1 = current device */
    const int32_t* prior_kind;/* n_para entries, SABC_PRIOR_* */
    const double*  prior_par; /* 2*n_para entries: Uniform (a,b) | Normal (mu,sigma) | Exponential (theta,0) | LogNormal (mu,sigma) | Gamma (alpha,theta) | Beta (alpha,beta) | Cauchy (mu,sigma) | Laplace (mu,theta) | Weibull (alpha,theta) | InverseGamma (alpha,theta) */
    /* multi-GPU, one process per GPU: this rank owns a contiguous slice of n_particles/world_size (one process for all GPUs: n_gpus below) */
    int32_t  rank, world_size;
    const void* nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks; NULL when world_size == 1 */
    uint32_t flags;           /* SABC_FLAG_* */
    /* 0 = the reference's ECDF over the whole prior sample (src/cdf_estimators.jl:23-44).  K >= 2: compressed mode, the table
     * keeps K rank-uniform quantiles of the positive prior distances (plus 0 and 1.5 max) and fits into shared memory as a whole;
     * an approximation of order 1/K that the oracle reproduces when given the same K. */
    int32_t  ecdf_max_knots;
    /* ABI 2.  n_gpus > 1: ONE process drives n_gpus devices through this handle (the reference's caller is one Julia session,
     * src/SimulatedAnnealingABC.jl:451-460): the library forms an in-process communicator, GPU r owns the particle slice
     * [r N/n_gpus, (r+1) N/n_gpus), host arrays are the GLOBAL N x d / N x s matrices.  gpu_ids: n_gpus device ordinals, NULL = 0 .. n_gpus-1.
     * rank / world_size / nccl_unique_id / device are ignored then.  Same bits as the process-per-GPU mode with the same seed. */
    int32_t  n_gpus;
    const int32_t* gpu_ids;
} sabc_config;

#define SABC_FLAG_NO_GRAPH      1u  /* launch kernels directly instead of replaying a CUDA graph */
#define SABC_FLAG_TIME_KERNELS  2u  /* record CUDA events around every update_half / simulate_accept launch (implies NO_GRAPH) */
#define SABC_FLAG_NO_PIPELINE   8u  /* sabc_update_host: upload, update, download strictly one after the other */
#define SABC_FLAG_SORT_WORK    16u  /* split path: bucket the work list by the model's similarity key (if it has one) */
#define SABC_FLAG_GENERIC_TAIL 32u  /* never use the single-CTA tail kernel of small populations */
#define SABC_FLAG_MG_REPLICATED 64u /* world_size > 1 or n_gpus > 1: every rank holds the whole population and simulates a share of each half-sweep;
                                       bit-identical to one GPU (strict mode for parity studies, memory does not scale) */
#define SABC_FLAG_MG_STRICT_RESAMPLE 128u /* world_size > 1 or n_gpus > 1, sharded: every rank walks all N global resampling draws, so that the resampled
                                       multiset equals the single-GPU one for the same seed (O(N_global) work and memory per rank);
                                       default: per-rank counts from ONE shared-seed multinomial draw, O(N / world_size) per rank */
#define SABC_FLAG_FUSED         4u  /* always use the fused update_half kernel, also for simulation-heavy models */

/* timing of the last sabc_update(), measured with CUDA events on the engine's stream */
typedef struct sabc_timing {
    double  update_ms;         /* whole update loop */
    double  kernel_ms;         /* sum over update_half launches (only with SABC_FLAG_TIME_KERNELS) */
    int64_t kernel_launches;   /* number of update_half launches */
    int64_t total_launches;    /* all kernels launched by the loop */
    double  h2d_ms, d2h_ms;    /* host-buffer call: duration of the upload / of the final download (they overlap the updates) */
    double  host_ms;           /* host-buffer call: first byte up to last byte down, CUDA events */
    double  resample_ms;       /* multi-GPU: host wall time inside the global resampling exchanges */
    int64_t resample_events;
} sabc_timing;

/* ---- lifetime ---- */
int  sabc_abi_version(void);
int  sabc_create(sabc_engine** out, const sabc_config* cfg);
int  sabc_destroy(sabc_engine* e);                        /* idempotent on NULL; Julia finalizer */
const char* sabc_last_error(void);
int  sabc_device_count(int* n);
int  sabc_nccl_unique_id(void* out128);                   /* rank 0 creates, the host broadcasts */

/* ---- the path ---- */
/* initialization(): prior sample, ECDF build, transform, first resampling, eps_0
 * (src/SimulatedAnnealingABC.jl:151-227) */
int  sabc_init(sabc_engine* e);
/* update_population!(): n_simulation / n_particles population updates on the device-resident state
 * (src/SimulatedAnnealingABC.jl:251-402) */
int  sabc_update(sabc_engine* e, int64_t n_simulation, int64_t checkpoint_history);
/* same call with the SABCresult held in HOST buffers, as Julia's update_population!(::SABCresult)
 * would make it: uploads (theta,u,rho,eps,counters), runs the updates, downloads the result into the
 * same buffers.  world_size > 1: this rank's slice (n_particles / world_size rows).  n_gpus > 1: the GLOBAL arrays (leading
 * dimension n_particles); every GPU of the handle copies its own rows. */
int  sabc_update_host(sabc_engine* e, double* theta, double* u, double* rho, double* eps, int64_t counters[4],
                      int64_t n_simulation, int64_t checkpoint_history);

/* keyword arguments update_population! accepts anew on every call: v, δ, resample, proposal
 * (src/SimulatedAnnealingABC.jl:251-259).  prop_par as in sabc_config. */
int  sabc_set_tuning(sabc_engine* e, double v, double delta, int64_t resample, int32_t proposal, const double* prop_par);

/* ---- state in / out (SABCresult / SABCstate, src/SimulatedAnnealingABC.jl:28-60) ---- */
int  sabc_local_particles(sabc_engine* e, int64_t* n_local, int64_t* offset);
int  sabc_get_population(sabc_engine* e, double* theta, double* u, double* rho);      /* any pointer may be NULL */
int  sabc_set_population(sabc_engine* e, const double* theta, const double* u, const double* rho,
                         const double* eps, const int64_t counters[4]);
/* counters = { n_simulation, n_accept, n_resampling, n_population_updates } */
int  sabc_get_state(sabc_engine* e, double* eps, int64_t counters[4]);
int  sabc_history_len(sabc_engine* e, int64_t* n_records);
int  sabc_get_history(sabc_engine* e, double* eps_h /* n_rec x n_eps */, double* u_h /* n_rec x s */, double* rho_h);
int  sabc_get_ecdf(sabc_engine* e, int32_t stat, double* knots_out /* may be NULL */, int64_t* L);
int  sabc_set_ecdf(sabc_engine* e, int32_t stat, const double* knots, int64_t L);
int  sabc_get_timing(sabc_engine* e, sabc_timing* out);
int  sabc_update_kernel_info(sabc_engine* e, int* grid, int* block, int* smem_bytes, int* blocks_per_sm);

/* pinned host memory for the host-buffer calls */
int  sabc_host_alloc(void** out, int64_t bytes);
int  sabc_host_free(void* p);

/* ---- parity hooks: the engine's own device code on caller-supplied arrays ---- */
/* build_cdf(::AbstractVector)  src/cdf_estimators.jl:23-44; knots_out holds n+2 doubles */
int  sabc_ecdf_build(const double* dist, int64_t n, double* knots_out, int64_t* L);
/* cdfs_dist_prior(rho)  src/cdf_estimators.jl:68-70, through the staged multi-level index */
int  sabc_ecdf_transform(const double* knots, int64_t L, const double* rho, int64_t m, double* u_out);
/* accept rule  src/SimulatedAnnealingABC.jl:314-329; u_old/u_new are m x s column-major, uniform in [0,1) */
int  sabc_accept_step(int64_t m, int32_t s, const double* u_old, const double* u_new, const double* eps, int32_t n_eps,
                      const double* dlogprior, const double* log_factor, const double* uniform, uint8_t* accept_out);
int  sabc_update_epsilon_single(double ubar, double v, double* eps_out);               /* :92-95  */
int  sabc_update_epsilon_multi(const double* ubar, int32_t s, double v, double* eps_out); /* :100-117 */
/* resample_population  :124-137: fixed-point weights and the N categorical draws */
int  sabc_resample_weights(const double* u, int64_t n, int32_t s, const double* ubar, double delta, uint64_t* q_out);
int  sabc_resample_indices(const uint64_t* q, int64_t n, uint64_t seed, uint64_t resample_count, int64_t* idx_out);
int  sabc_exact_mean_u(const double* u, int64_t n, double* mean_out);
int  sabc_treesum(const double* x, int64_t n, double* sum_out);
/* deterministic math / samplers on the device: op 0 log, 1 exp, 2 sin(2 pi x), 3 cos(2 pi x), 4 log(x!) */
int  sabc_detmath(int32_t op, const double* x, int64_t n, double* out);
int  sabc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int  sabc_poisson(const double* lam, int64_t n, uint64_t seed, uint64_t sweep, int64_t* k_out, uint32_t* blocks_out);
/* the two cheap filters in front of the exact PTRS acceptance test (csrc/philox.cuh) against that test on `attempts`
 * candidates, lambda cycling through lam[]: counts_out = {reached the exact test, undecided by the MUFU filter, undecided
 * by both, wrong decisions of the MUFU filter, wrong decisions of the FP64 filter}; ratio_out = max |T - T_exact| / E */
int  sabc_ptrs_filter_check(const double* lam, int32_t n_lam, int64_t attempts, uint64_t seed, int64_t counts_out[5],
                            double ratio_out[2]);
int  sabc_prior_logpdf(int32_t d, const int32_t* kind, const double* par, const double* theta /* n x d */, int64_t n,
                       double* lp_out);
/* f_dist on the device: theta is n x d column-major, rho_out n x s; particle i uses Philox counter particle_base+i */
int  sabc_model_simulate(const char* model_name, const double* model_par, int32_t n_model_par, const double* theta,
                         int64_t n, uint64_t seed, uint32_t particle_base, uint64_t sweep, double* rho_out);
int  sabc_model_info(const char* model_name, int32_t* n_para, int32_t* n_stats);
/* proposal(theta_i, population_inactive)  src/proposals.jl: proposes for n active particles against M inactive ones */
int  sabc_propose(int32_t proposal, const double* prop_par, int32_t d, const double* theta_active /* n x d */, int64_t n,
                  const double* theta_inactive /* M x d */, int64_t M, const double* chol /* d x d row-major or sd */,
                  uint64_t seed, uint32_t particle_base, uint64_t sweep, double* theta_out, double* log_factor_out);

/* host-side plan of the surplus exchange of the multi-GPU resampling (pure function, needs no device): counts[g] = number
 * of the N global draws that selected a particle of rank g; outputs are offsets/counts into this rank's packed selection
 * (send) and into its particle slice (recv), per peer.  Entry `me` of both describes the part that stays local. */
int  sabc_mg_exchange_plan(const int64_t* counts, int32_t world, int64_t n_local, int32_t me, int64_t* send_off,
                           int64_t* send_cnt, int64_t* recv_off, int64_t* recv_cnt);

/* host-side split of the N resampling draws over the ranks of a sharded population (pure function, needs no device): counts[g] ~
 * Multinomial(n_draws; w[g] / sum w) from the shared seed and the resampling count, the same vector on every rank */
int  sabc_multinomial_split(int64_t n_draws, const uint64_t* w, int32_t world, uint64_t seed, uint32_t resample_count, int64_t* counts_out);

/* ---- device model plug-in registry (new; the GPU form of f_dist) ---- */
/* `vtable` points to a sabc::ModelVTable (csrc/kernels.cuh) built by an out-of-tree .cu that includes
 * the header-only kernel templates; typically called from a static initialiser at dlopen time. */
int  sabc_register_model(const void* vtable);
int  sabc_model_count(void);
const char* sabc_model_name(int index);

#ifdef __cplusplus
}
#endif
#endif
