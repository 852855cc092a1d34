/* Plain-C host of the C ABI (what a Julia `ccall` or any FFI does): C1 = 1-D Gaussian mean, N = 1000, n_simulation = 100000.
 *   gcc -O2 -I include examples/c_host/sabc_c_host.c -o sabc_c_host -L simulatedannealingabc.jl_b200 -l:libsabc_b200.so \
 *       -Wl,-rpath,$PWD/simulatedannealingabc.jl_b200
 * Exit code 0 and one line "ok ..." on success; on a box without a CUDA device it prints the library's error and exits 3.
 * `sabc_c_host G` (G > 1): the same call sequence with cfg.n_gpus = G -- this one process drives G GPUs through the one handle, the
 * arrays it reads back are the global ones. */
#include <stdio.h>
#include <stdlib.h>
#include "sabc_b200.h"

int main(int argc, char** argv) {
    const double model_par[2] = {1.0, 0.31622776601683794};       /* ybar_obs, sigma/sqrt(n) */
    const int32_t prior_kind[1] = {SABC_PRIOR_NORMAL};
    const double prior_par[2] = {0.0, 1.0};
    sabc_config cfg = {0};
    cfg.n_particles = 1000; cfg.n_para = 1; cfg.n_stats = 1;
    cfg.algorithm = SABC_ALG_SINGLE_EPS; cfg.proposal = SABC_PROP_DE;
    cfg.prop_par[0] = 2.38 / 1.4142135623730951; cfg.prop_par[1] = 1e-5;
    cfg.v = 1.0; cfg.delta = 0.1; cfg.resample = 1000; cfg.seed = 0x5ABC;   /* the configuration of tests/golden/trajectories.json[0] */
    cfg.model_name = "gauss_mean"; cfg.model_par = model_par; cfg.n_model_par = 2; cfg.device = -1;
    cfg.prior_kind = prior_kind; cfg.prior_par = prior_par; cfg.rank = 0; cfg.world_size = 1;
    cfg.n_gpus = argc > 1 ? atoi(argv[1]) : 0; cfg.gpu_ids = NULL;         /* ABI 2: single-process multi-GPU handle */

    sabc_engine* e = NULL;
    int rc = sabc_create(&e, &cfg);
    if (rc == 0) rc = sabc_init(e);                                /* initialization() */
    if (rc == 0) rc = sabc_update(e, 100000 - 1000, 1);            /* update_population!() */
    if (rc != 0) { fprintf(stderr, "sabc error %d: %s\n", rc, sabc_last_error()); sabc_destroy(e); return 3; }
    double eps[1]; int64_t cnt[4], n_rec = 0;
    sabc_get_state(e, eps, cnt);
    sabc_history_len(e, &n_rec);
    double* theta = malloc(1000 * sizeof(double));
    sabc_get_population(e, theta, NULL, NULL);
    double mean = 0.0;
    for (int i = 0; i < 1000; ++i) mean += theta[i];
    printf("ok eps=%a n_simulation=%lld n_accept=%lld n_resampling=%lld n_population_updates=%lld records=%lld mean=%.6f\n", eps[0],
           (long long)cnt[0], (long long)cnt[1], (long long)cnt[2], (long long)cnt[3], (long long)n_rec, mean / 1000.0);
    free(theta);
    sabc_destroy(e);
    return 0;
}
