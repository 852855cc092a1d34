// group.inl -- single-process multi-GPU handle (sabc_config.n_gpus > 1).  Included by engine.cu.
//
// The caller described by the reference's surface is ONE host process: `sabc(f_dist, prior; n_particles = 10^7)` from one Julia
// session (src/SimulatedAnnealingABC.jl:451-460).  Such a caller gets all GPUs of the box through one handle: the group engine
// owns one per-GPU engine per device -- the ranks of an in-process NCCL communicator, each with the contiguous particle slice
// [r N/G, (r+1) N/G) -- and every C-ABI call fans out to them on one host thread per GPU (SURVEY.md section 8b "one thread per
// GPU inside the library").  Host arrays are the caller's GLOBAL column-major arrays; each GPU copies its own rows.  The per-GPU
// engines run exactly the code of the process-per-GPU mode, so both modes give the same bits for the same seed.
#pragma once

// run f(child, rank) on every child, one thread each; the first failure's code and message become the caller's error
template <class F>
static int group_run(sabc_engine* e, F f) {
    const int G = (int)e->children.size();
    std::vector<int> rc(G, 0);
    std::vector<std::string> msg(G);
    auto body = [&](int r) {
        rc[r] = f(e->children[r], r);
        if (rc[r] != 0) msg[r] = last_error_buf();
    };
    std::vector<std::thread> th;
    for (int r = 1; r < G; ++r) th.emplace_back(body, r);
    body(0);
    for (auto& t : th) t.join();
    for (int r = 0; r < G; ++r)
        if (rc[r] != 0) return set_error(rc[r], "GPU %d of the group: %s", r, msg[r].c_str());
    return 0;
}

static int group_create(sabc_engine** out, const sabc_config* c) {
    const int G = c->n_gpus;
    if (c->world_size > 1) return set_error(SABC_ERR_INVALID, "n_gpus > 1 (one process drives the GPUs) excludes world_size > 1 (one process per GPU)");
    int ndev = 0;
    SABC_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < G) return set_error(SABC_ERR_CUDA, "n_gpus = %d but only %d CUDA device(s) are visible", G, ndev);
    if (c->n_particles % G != 0) return set_error(SABC_ERR_INVALID, "n_particles must be divisible by n_gpus");
    std::vector<int> dev(G);
    for (int r = 0; r < G; ++r) {
        dev[r] = c->gpu_ids ? c->gpu_ids[r] : r;
        if (dev[r] < 0 || dev[r] >= ndev) return set_error(SABC_ERR_INVALID, "gpu_ids[%d] = %d is not a visible device", r, dev[r]);
        for (int q = 0; q < r; ++q) if (dev[q] == dev[r]) return set_error(SABC_ERR_INVALID, "gpu_ids names device %d twice", dev[r]);
    }
    unsigned char uid[128];
    SABC_TRY(nccl_get_unique_id(uid));
    auto* g = new sabc_engine();
    g->children.assign(G, nullptr);
    g->N = c->n_particles; g->D = c->n_para; g->S = c->n_stats; g->world = 1; g->rank = 0;
    g->n_eps = c->algorithm == SABC_ALG_MULTI_EPS ? c->n_stats : 1;
    g->replicated = (c->flags & SABC_FLAG_MG_REPLICATED) != 0;
    g->n_local = c->n_particles; g->offset = 0;
    // the communicator forms only when all ranks call in: the children are created concurrently
    const int rc = group_run(g, [&](sabc_engine*&, int r) -> int {
        sabc_config cc = *c;
        cc.n_gpus = 0; cc.gpu_ids = nullptr;
        cc.device = dev[r]; cc.rank = r; cc.world_size = G; cc.nccl_unique_id = uid;
        return sabc_create(&g->children[r], &cc);
    });
    if (rc != 0) {
        for (auto*& ch : g->children) { if (ch) sabc_destroy(ch); ch = nullptr; }
        g->children.clear();
        delete g;
        return rc;
    }
    *out = g;
    return 0;
}

static int group_destroy(sabc_engine* e) {
    group_run(e, [](sabc_engine* ch, int) { return sabc_destroy(ch); });
    e->children.clear();
    delete e;
    return 0;
}

static void group_pull_state(sabc_engine* e) {           // every rank holds the same global scalars: mirror rank 0
    const sabc_engine* c0 = e->children[0];
    for (int k = 0; k < e->n_eps; ++k) e->eps[k] = c0->eps[k];
    e->n_simulation = c0->n_simulation; e->n_accept = c0->n_accept; e->n_resampling = c0->n_resampling;
    e->n_population_updates = c0->n_population_updates; e->initialised = c0->initialised;
    e->timing = c0->timing;
    for (auto* ch : e->children) {
        e->timing.update_ms = std::max(e->timing.update_ms, ch->timing.update_ms);
        e->timing.host_ms = std::max(e->timing.host_ms, ch->timing.host_ms);
        e->timing.h2d_ms = std::max(e->timing.h2d_ms, ch->timing.h2d_ms);
        e->timing.d2h_ms = std::max(e->timing.d2h_ms, ch->timing.d2h_ms);
        e->timing.resample_ms = std::max(e->timing.resample_ms, ch->timing.resample_ms);
        if (ch != c0) { e->timing.total_launches += ch->timing.total_launches; e->timing.kernel_launches += ch->timing.kernel_launches; }
    }
}

static int group_init(sabc_engine* e) {
    SABC_TRY(group_run(e, [](sabc_engine* ch, int) { return sabc_init(ch); }));
    group_pull_state(e);
    return 0;
}
static int group_update(sabc_engine* e, int64_t n_simulation, int64_t checkpoint_history) {
    SABC_TRY(group_run(e, [=](sabc_engine* ch, int) { return sabc_update(ch, n_simulation, checkpoint_history); }));
    group_pull_state(e);
    return 0;
}
static int group_set_tuning(sabc_engine* e, double v, double delta, int64_t resample, int32_t proposal, const double* prop_par) {
    return group_run(e, [=](sabc_engine* ch, int) { return sabc_set_tuning(ch, v, delta, resample, proposal, prop_par); });
}
// slice r of a global column-major array with leading dimension N starts at row r n
static int group_get_population(sabc_engine* e, double* theta, double* u, double* rho) {
    if (e->replicated) return sabc_get_population(e->children[0], theta, u, rho);
    const int64_t N = e->N;
    return group_run(e, [=](sabc_engine* ch, int r) {
        const int64_t o = (int64_t)r * ch->n_local;
        return get_population_ld(ch, theta ? theta + o : nullptr, u ? u + o : nullptr, rho ? rho + o : nullptr, N);
    });
}
static int group_set_population(sabc_engine* e, const double* theta, const double* u, const double* rho, const double* eps,
                                const int64_t counters[4]) {
    if (!theta || !u || !rho || !eps || !counters) return set_error(SABC_ERR_INVALID, "null argument");
    const int64_t N = e->N;
    const bool rep = e->replicated;
    SABC_TRY(group_run(e, [=](sabc_engine* ch, int r) {
        const int64_t o = rep ? 0 : (int64_t)r * ch->n_local;
        return set_population_ld(ch, theta + o, u + o, rho + o, N, eps, counters);
    }));
    group_pull_state(e);
    return 0;
}
static int group_update_host(sabc_engine* e, double* theta, double* u, double* rho, double* eps, int64_t counters[4],
                             int64_t n_simulation, int64_t checkpoint_history) {
    if (!theta || !u || !rho || !eps || !counters) return set_error(SABC_ERR_INVALID, "null argument");
    const int G = (int)e->children.size();
    const int64_t N = e->N;
    const bool rep = e->replicated;
    std::vector<std::vector<double>> eps_r(G, std::vector<double>(eps, eps + e->n_eps));
    std::vector<std::vector<int64_t>> cnt_r(G, std::vector<int64_t>(counters, counters + 4));
    std::vector<double> scratch;
    SABC_TRY(group_run(e, [&](sabc_engine* ch, int r) -> int {
        if (rep && r > 0) {                                 // every rank needs the input, only rank 0 writes the result back
            SABC_TRY(set_population_ld(ch, theta, u, rho, N, eps_r[r].data(), cnt_r[r].data()));
            return sabc_update(ch, n_simulation, checkpoint_history);
        }
        const int64_t o = rep ? 0 : (int64_t)r * ch->n_local;
        return update_host_ld(ch, theta + o, u + o, rho + o, N, eps_r[r].data(), cnt_r[r].data(), n_simulation, checkpoint_history);
    }));
    for (int k = 0; k < e->n_eps; ++k) eps[k] = eps_r[0][k];
    for (int k = 0; k < 4; ++k) counters[k] = cnt_r[0][k];
    group_pull_state(e);
    return 0;
}
static int group_set_ecdf(sabc_engine* e, int32_t stat, const double* knots, int64_t L) {
    SABC_TRY(group_run(e, [=](sabc_engine* ch, int) { return sabc_set_ecdf(ch, stat, knots, L); }));
    e->top_doubles = e->children[0]->top_doubles;
    return 0;
}
