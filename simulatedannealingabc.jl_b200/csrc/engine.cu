// engine.cu -- the SABC population engine behind the C ABI (include/sabc_b200.h).
//
// Restates, B200-first, initialization() and update_population!() of the reference
// (src/SimulatedAnnealingABC.jl:151-227, 251-402): all particle state is structure-of-arrays FP64
// in HBM, one population update is a fixed sequence of kernels whose control state (iteration
// number, ε, accept counters, resampling trigger, history cursor) lives in device memory, so the
// sequence is captured once in a CUDA graph and replayed without host round trips.
#include "ecdf_index.cuh"
#include "nccl_dyn.h"
#include "multinomial.h"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace sabc {

// ---------------------------------------------------------------------------------------------
// error text + model registry
// ---------------------------------------------------------------------------------------------
char* last_error_buf() { static thread_local char buf[1024] = {0}; return buf; }
int set_error(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(last_error_buf(), 1024, fmt, ap);
    va_end(ap);
    return code;
}

// Entries have stable addresses (std::deque never moves an element on push_back): engines keep a pointer to their model's entry, and a
// plug-in loaded later (dlopen of an out-of-tree model) must not invalidate it.  The name is copied: the caller's string may go away.
static std::mutex g_reg_mutex;
static std::deque<ModelVTable>& registry() { static std::deque<ModelVTable> r; return r; }
static std::deque<std::string>& registry_names() { static std::deque<std::string> r; return r; }
static void registry_add(const ModelVTable& v) {
    registry_names().emplace_back(v.name);
    registry().push_back(v);
    registry().back().name = registry_names().back().c_str();
}
static void ensure_builtin_models() {
    static std::once_flag once;
    std::call_once(once, [] {
        registry_add(ModelLaunchers<GaussMean>::vtable("gauss_mean"));
        registry_add(ModelLaunchers<GaussSample<1, 1>>::vtable("gauss_sample_d1s1"));
        registry_add(ModelLaunchers<GaussSample<1, 2>>::vtable("gauss_sample_d1s2"));
        registry_add(ModelLaunchers<GaussSample<2, 1>>::vtable("gauss_sample_d2s1"));
        registry_add(ModelLaunchers<GaussSample<2, 2>>::vtable("gauss_sample_d2s2"));
        registry_add(ModelLaunchers<Logistic>::vtable("logistic", 1));
        registry_add(ModelLaunchers<SirTauLeap>::vtable("sir_tauleap", 1));
        registry_add(ModelLaunchers<SirGillespie<3>>::vtable("sir_gillespie_s3", 1));
        registry_add(ModelLaunchers<SirGillespie<1>>::vtable("sir_gillespie_s1", 1));
    });
}
const ModelVTable* find_model(const char* name) {
    ensure_builtin_models();
    std::lock_guard<std::mutex> lk(g_reg_mutex);
    for (auto& m : registry()) if (std::strcmp(m.name, name) == 0) return &m;
    return nullptr;
}

}  // namespace sabc

using namespace sabc;

// ---------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------
constexpr int MG_LOOKAHEAD = 4;   // updates the host keeps enqueued ahead of the one whose resampling decision it is waiting for
struct MgScratch {   // multi-GPU work space
    DevBuf<unsigned long long> wall, F, tsum, toff, scalar;
    DevBuf<unsigned long long> stats_send, stats_all;   // per-update statistics: this rank's packed record, all ranks' records
    DevBuf<int64_t> src;
    DevBuf<double> sb;
    DevBuf<int> flag;
    int* hold_slots = nullptr;                          // pinned: `hold` after each of the last MG_LOOKAHEAD enqueued updates
    cudaEvent_t hold_ev[MG_LOOKAHEAD] = {};
    bool graph_failed = false;
    ~MgScratch() {
        if (hold_slots) cudaFreeHost(hold_slots);
        for (auto ev : hold_ev) if (ev) cudaEventDestroy(ev);
    }
};

struct sabc_engine {
    // single-process multi-GPU handle: the engine is a group of per-GPU engines (ranks of an in-process communicator), every call
    // fans out to them on one host thread per GPU (group.inl); a leaf engine has no children
    std::vector<sabc_engine*> children;
    bool is_group() const { return !children.empty(); }
    MgScratch mg;
    // configuration
    int64_t N = 0, n_local = 0, offset = 0;
    int D = 0, S = 0, n_eps = 1, algorithm = 0, proposal = 0;
    double prop_par[2] = {0, 0}, v = 1.0, delta = 0.1;
    int64_t resample = 0;
    uint64_t seed = 0;
    uint32_t flags = 0;
    int ecdf_max_knots = 0;
    int device = 0, rank = 0, world = 1, n_sm = 148;
    bool replicated = false;   // world > 1 with the whole population on every GPU (SABC_FLAG_MG_REPLICATED)
    bool sharded() const { return world > 1 && !replicated; }
    const ModelVTable* model = nullptr;
    ModelPar mp{};
    PriorSpec prior{};
    cudaStream_t stream = nullptr;
    NcclComm comm;

    // device state
    DevBuf<double> b_theta, b_u, b_rho, b_lp, b_ttheta, b_tu, b_tlp;
    PopView pop{}, tmp{};
    DevBuf<DevState> b_ds;
    DevBuf<EcdfStat> b_ecdf;
    EcdfStat h_ecdf[MAX_S];
    std::vector<DevBuf<double>*> ecdf_bufs;
    int top_doubles = 0;
    DevBuf<unsigned long long> b_q, b_tile_sum, b_tile_off;
    DevBuf<double> b_rho_part, b_scratch, b_rw_part, b_rw_sums, b_hist;
    DevBuf<double> b_sp_theta, b_sp_lp, b_sp_lf;      // split path work list
    DevBuf<uint32_t> b_sp_idx, b_sp_key, b_sp_perm;
    DevBuf<unsigned int> b_sp_hist, b_sp_off;
    bool sort_work = false;
    bool split = false;
    int grid_simacc = 0, bps_simacc = 0;

    // per-call hooks of the sweeps: kernel timing events and the copy/compute pipeline of the host-buffer call
    std::vector<cudaEvent_t>* kev = nullptr;           // event pairs around the dominant kernel of each (sub-)sweep
    int pipe_nsub = 1;                                 // sub-ranges per half-sweep (1 = plain half-sweeps)
    bool pipe_first = false, pipe_last = false;        // this update is the first / last one of a pipelined host call
    std::function<int(int, int)> pipe_before;          // (half, sub): stream waits for the rows' upload, log-prior of the rows
    std::function<int(int, int)> pipe_after;           // (half, sub): rows are final, enqueue their download
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int64_t part_ld = 0, scratch_ld = 0, hist_cap = 0;
    int grid_update = 0, grid_aux = 0, bps_update = 0;
    size_t smem_update = 0;

    // graph
    cudaGraphExec_t graph_exec = nullptr;
    bool mg_warm = false;      // sharded: one update has run with direct launches (NCCL connections exist), the graph may be captured

    // host mirror of SABCstate
    bool initialised = false;
    double eps[MAX_S] = {0};
    int64_t n_simulation = 0, n_accept = 0, n_resampling = 0, n_population_updates = 0;
    std::vector<double> eps_h, u_h, rho_h;
    sabc_timing timing{};

    ~sabc_engine() {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        for (auto* b : ecdf_bufs) delete b;
        comm.destroy();
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        if (stream) cudaStreamDestroy(stream);
    }
};

static int halves(const sabc_engine* e, int half, int64_t& act_off, int64_t& act_n, int64_t& ina_off, int64_t& ina_n) {
    const int64_t h0 = e->n_local / 2;                               // :300-301 batch_1 = 1:(N÷2)
    if (half == 0) { act_off = 0; act_n = h0; ina_off = h0; ina_n = e->n_local - h0; }
    else { act_off = h0; act_n = e->n_local - h0; ina_off = 0; ina_n = h0; }
    return 0;
}

static int free_ecdf(sabc_engine* e) {
    for (auto* b : e->ecdf_bufs) delete b;
    e->ecdf_bufs.clear();
    return 0;
}

// Build the index levels of statistic j over knots already resident in `knots` (device, L entries).
static int ecdf_attach(sabc_engine* e, int j, DevBuf<double>* knots, int64_t L, int top_max) {
    return ecdf_build_index(e->h_ecdf[j], knots->p, L, top_max, e->ecdf_bufs, e->stream);
}

static int ecdf_finalize(sabc_engine* e) {
    int off = 0;
    for (int j = 0; j < e->S; ++j) { e->h_ecdf[j].top_off = off; off += e->h_ecdf[j].top_pow2; }
    e->top_doubles = off;
    SABC_CUDA(e->b_ecdf.ensure(MAX_S));
    SABC_CUDA(cudaMemcpyAsync(e->b_ecdf.p, e->h_ecdf, sizeof(EcdfStat) * e->S, cudaMemcpyHostToDevice, e->stream));
    e->smem_update = (size_t)off * sizeof(double);
    int bps = 0;
    SABC_CUDA(e->model->update_occupancy(e->proposal, e->smem_update, &bps));
    if (bps < 1) return set_error(SABC_ERR_CUDA, "update kernel does not fit on an SM (smem %zu B)", e->smem_update);
    e->bps_update = bps;
    const int64_t groups = (e->n_local - e->n_local / 2 + CHUNK - 1) / CHUNK;
    e->grid_update = (int)std::max<int64_t>(1, std::min<int64_t>(groups, (int64_t)bps * e->n_sm));
    if (e->split) {
        int b2 = 0;
        SABC_CUDA(e->model->simacc_occupancy(e->smem_update, &b2));
        if (b2 < 1) return set_error(SABC_ERR_CUDA, "simulate_accept kernel does not fit on an SM");
        e->bps_simacc = b2;
        e->grid_simacc = (int)std::max<int64_t>(1, std::min<int64_t>(groups, (int64_t)b2 * e->n_sm));
    }
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    return 0;
}

static int top_max_for(int S) { return ecdf_top_max(S); }

// build_cdf for column j from `d_col` (n values on the device)   src/cdf_estimators.jl:23-44
static int ecdf_build_column(sabc_engine* e, int j, const double* d_col, int64_t n, DevBuf<double>& keys,
                             DevBuf<unsigned char>& cub_tmp, DevBuf<unsigned long long>& cnt) {
    SABC_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), e->stream));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 8192);
    k_mark_positive<<<grid, 256, 0, e->stream>>>(d_col, n, keys.p, cnt.p);
    SABC_CUDA(cudaGetLastError());
    auto* knots = new DevBuf<double>();
    e->ecdf_bufs.push_back(knots);
    SABC_CUDA(knots->alloc((size_t)n + 2 + ECDF_PAD));
    k_fill_inf<<<1, 32, 0, e->stream>>>(knots->p + n + 2, ECDF_PAD);   // entries L..n+1 are +inf sentinels of the sort already
    size_t tmp_bytes = 0;
    SABC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys.p, knots->p + 1, (int64_t)n, 0, 64, e->stream));
    SABC_CUDA(cub_tmp.ensure(tmp_bytes));
    SABC_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp.p, tmp_bytes, keys.p, knots->p + 1, (int64_t)n, 0, 64, e->stream));
    unsigned long long n_pos = 0;
    SABC_CUDA(cudaMemcpyAsync(&n_pos, cnt.p, sizeof n_pos, cudaMemcpyDeviceToHost, e->stream));
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    if (n_pos == 0) return set_error(SABC_ERR_NO_POSITIVE, "build_cdf: statistic %d has no positive prior distance", j + 1);
    if (e->ecdf_max_knots >= 2 && (int64_t)n_pos > e->ecdf_max_knots) {      // compressed mode: keep K quantiles only
        const int K = e->ecdf_max_knots;
        auto* small = new DevBuf<double>();
        SABC_CUDA(small->alloc((size_t)K + 2 + ECDF_PAD));
        k_ecdf_subsample<<<(K + 2 + ECDF_PAD + 255) / 256, 256, 0, e->stream>>>(knots->p + 1, (int64_t)n_pos, K, small->p);
        SABC_CUDA(cudaGetLastError());
        SABC_CUDA(cudaStreamSynchronize(e->stream));
        e->ecdf_bufs.pop_back(); delete knots;                                // the full sorted sample is not kept
        e->ecdf_bufs.push_back(small);
        return ecdf_attach(e, j, small, (int64_t)K + 2, std::max(top_max_for(e->S), pow2_ceil(K + 2)));
    }
    k_ecdf_ends<<<1, 1, 0, e->stream>>>(knots->p, (int64_t)n_pos);
    SABC_CUDA(cudaGetLastError());
    return ecdf_attach(e, j, knots, (int64_t)n_pos + 2, top_max_for(e->S));
}

// rows [r0, r1) (relative to the half start) of sub-range `sub` of `nsub`; boundaries are multiples of the 256-group
static void sub_range(int64_t act_n, int sub, int nsub, int64_t& r0, int64_t& r1) {
    const int64_t per = (((act_n + nsub - 1) / nsub + CHUNK - 1) / CHUNK) * CHUNK;
    r0 = std::min<int64_t>(act_n, (int64_t)sub * per);
    r1 = std::min<int64_t>(act_n, r0 + per);
}
static UpdateArgs make_update_args(sabc_engine* e, int half, int sub = 0, int nsub = 1) {
    UpdateArgs a{};
    a.pop = e->pop;
    halves(e, half, a.act_off, a.act_n, a.ina_off, a.ina_n);
    a.particle_base = (uint32_t)e->offset;
    a.half = half; a.seed = e->seed; a.rk = make_round_keys(e->seed); a.ds = e->b_ds.p; a.ecdf = e->b_ecdf.p;
    a.rho_part = e->b_rho_part.p + (int64_t)half * e->S * e->part_ld;
    a.part_ld = e->part_ld; a.n_eps = e->n_eps; a.top_doubles = e->top_doubles;
    a.prop0 = e->prop_par[0]; a.prop1 = e->prop_par[1];
    a.prior = e->prior; a.mp = e->mp;
    a.slot = half * MAX_SUB + sub;
    if (nsub > 1) {
        int64_t r0, r1;
        sub_range(a.act_n, sub, nsub, r0, r1);
        a.act_off += r0; a.act_n = r1 - r0; a.rho_part += r0 / CHUNK;
    }
    return a;
}
static SplitScratch make_split(sabc_engine* e) {
    SplitScratch w{};
    w.theta = e->b_sp_theta.p; w.lp = e->b_sp_lp.p; w.lf = e->b_sp_lf.p; w.idx = e->b_sp_idx.p;
    w.count = &e->b_ds.p->list_count[0]; w.cursor = &e->b_ds.p->list_cursor[0];
    w.cap = e->n_local - e->n_local / 2;
    if (e->sort_work) { w.key = e->b_sp_key.p; w.perm = e->b_sp_perm.p; w.hist = e->b_sp_hist.p; w.off = e->b_sp_off.p; }
    return w;
}
static int launch_update_half(sabc_engine* e, int half, int sub = 0, int nsub = 1) {
    const UpdateArgs a = make_update_args(e, half, sub, nsub);
    if (a.act_n <= 0) return 0;
    SABC_CUDA(e->model->launch_update(e->proposal, a, e->grid_update, e->smem_update, e->stream));
    return 0;
}
// split path, one (sub-range of a) half: propose + compact, then one simulation per lane over the work list
static int launch_split_propose(sabc_engine* e, int half, int sub = 0, int nsub = 1) {
    const UpdateArgs a = make_update_args(e, half, sub, nsub);
    if (a.act_n <= 0) return 0;
    const int64_t groups = (a.act_n + CHUNK - 1) / CHUNK;
    const SplitScratch w = make_split(e);
    SABC_CUDA(e->model->launch_propose(e->proposal, a, w, (int)std::min<int64_t>(groups, (int64_t)e->n_sm * 8), e->stream));
    if (e->sort_work) {                // group the work list by the model's similarity key
        bucket_scan_kernel<<<1, 1024, 0, e->stream>>>(w.hist, w.off);
        bucket_scatter_kernel<<<(int)std::min<int64_t>(groups, (int64_t)e->n_sm * 8), CHUNK, 0, e->stream>>>(w.key, w.count, a.slot, w.hist, w.off, w.perm);
        bucket_clear_kernel<<<1, 1024, 0, e->stream>>>(w.hist);
        SABC_CUDA(cudaGetLastError());
    }
    return 0;
}
static int launch_split_simacc(sabc_engine* e, int half, int sub = 0, int nsub = 1) {
    const UpdateArgs a = make_update_args(e, half, sub, nsub);
    if (a.act_n <= 0) return 0;
    SABC_CUDA(e->model->launch_simacc(a, make_split(e), e->grid_simacc, e->smem_update, e->stream));
    return 0;
}
static int launch_split_stats(sabc_engine* e) {
    for (int half = 0; half < 2; ++half) {
        int64_t off, n, t0, t1;
        halves(e, half, off, n, t0, t1);
        if (n <= 0) continue;
        const int64_t groups = (n + CHUNK - 1) / CHUNK;
        stats_kernel<<<(int)std::min<int64_t>(groups, (int64_t)e->n_sm * 8), CHUNK, 0, e->stream>>>(
            e->pop, off, n, e->S, e->b_ds.p, e->b_rho_part.p + (int64_t)half * e->S * e->part_ld, e->part_ld);
    }
    SABC_CUDA(cudaGetLastError());
    return 0;
}

static int launch_post1(sabc_engine* e, int decide) {
    Post1Args a{};
    a.ds = e->b_ds.p; a.rho_part = e->b_rho_part.p; a.part_ld = e->part_ld;
    int64_t o, n0, n1, t0, t1;
    halves(e, 0, o, n0, t0, t1); halves(e, 1, o, n1, t0, t1);
    a.groups0 = (n0 + CHUNK - 1) / CHUNK; a.groups1 = (n1 + CHUNK - 1) / CHUNK;
    a.scratch = e->b_scratch.p; a.scratch_ld = e->scratch_ld;
    a.S = e->S; a.n_global = e->N; a.resample = e->resample; a.decide = decide;
    k_post1<<<2 * e->S + 1, CHUNK, 0, e->stream>>>(a);
    SABC_CUDA(cudaGetLastError());
    return 0;
}

// resample_population on one GPU (:124-137).  force = 1 at initialization (:197).
static int launch_resample_local(sabc_engine* e, int force) {
    const int64_t n = e->n_local;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    const int g_tiles = (int)std::min<int64_t>(n_tiles, (int64_t)e->n_sm * 8);
    const int g_grp = (int)std::min<int64_t>((n + CHUNK - 1) / CHUNK, (int64_t)e->n_sm * 8);
    DevState* ds = e->b_ds.p;
    k_weights<<<g_tiles, CHUNK, 0, e->stream>>>(e->pop, n, e->S, e->delta, ds, e->b_q.p, e->b_tile_sum.p, force);
    k_scan_tiles<<<1, 1024, 0, e->stream>>>(e->b_tile_sum.p, n_tiles, e->b_tile_off.p, &ds->w_total, ds, force);
    k_prefix<<<g_tiles, CHUNK, 0, e->stream>>>(e->b_q.p, n, e->b_tile_off.p, ds, force);
    k_draw_gather<<<g_grp, CHUNK, 0, e->stream>>>(e->pop, e->tmp, n, e->D, e->S, e->b_q.p, e->seed, ds, force);
    k_copyback<<<g_grp, CHUNK, 0, e->stream>>>(e->pop, e->tmp, n, e->D, e->S, ds, force);
    SABC_CUDA(cudaGetLastError());
    return 0;
}

// update_proposal!(RandomWalk)  src/proposals.jl:46-48,58-60
static int launch_update_proposal(sabc_engine* e) {
    if (e->proposal != PROP_RW) return 0;
    const int64_t n = e->n_local, groups = (n + CHUNK - 1) / CHUNK;
    const int grid = (int)std::min<int64_t>(groups, (int64_t)e->n_sm * 8);
    const int npair = e->D * (e->D + 1) / 2;
    DevState* ds = e->b_ds.p;
    k_group_sums<<<grid, CHUNK, 0, e->stream>>>(e->pop.theta, e->pop.ld, n, e->D, e->b_rw_part.p, e->part_ld * 2);
    k_treesum_cols<<<e->D, CHUNK, 0, e->stream>>>(e->b_rw_part.p, e->part_ld * 2, groups, e->b_scratch.p, e->scratch_ld, e->b_rw_sums.p);
    k_rw_means<<<1, 32, 0, e->stream>>>(ds, e->b_rw_sums.p, e->D, e->N);
    k_rw_cross_sums<<<grid, CHUNK, 0, e->stream>>>(e->pop.theta, e->pop.ld, n, e->D, ds, e->b_rw_part.p, e->part_ld * 2);
    k_treesum_cols<<<npair, CHUNK, 0, e->stream>>>(e->b_rw_part.p, e->part_ld * 2, groups, e->b_scratch.p, e->scratch_ld, e->b_rw_sums.p);
    k_rw_chol<<<1, 32, 0, e->stream>>>(ds, e->b_rw_sums.p, e->D, e->N, e->prop_par[0]);
    SABC_CUDA(cudaGetLastError());
    return 0;
}

static int launch_finish(sabc_engine* e) {
    FinishArgs a{};
    a.ds = e->b_ds.p; a.hist = e->b_hist.p; a.S = e->S; a.n_eps = e->n_eps; a.algorithm = e->algorithm;
    a.n_global = e->N; a.v = e->v;
    k_finish<<<1, 32, 0, e->stream>>>(a);
    SABC_CUDA(cudaGetLastError());
    return 0;
}

static bool small_tail(const sabc_engine* e);
static int kernels_per_iteration(const sabc_engine* e) {
    return (e->split ? 6 : 2) + (e->sort_work ? 6 : 0) + (small_tail(e) ? 1 : 1 + 5 + (e->proposal == PROP_RW ? 6 : 0) + 1);
}

// the two half-sweeps (:304-332) in the fused or the split form.  A pipelined host call cuts each half into sub-ranges
// (identical results: the particles of a half-sweep are independent) so that transfers overlap at a finer grain.
static int enqueue_sweeps(sabc_engine* e) {
    const int nsub = e->pipe_nsub;
    for (int half = 0; half < 2; ++half) {
        for (int sub = 0; sub < nsub; ++sub) {
            if (e->pipe_first && e->pipe_before) SABC_TRY(e->pipe_before(half, sub));
            if (e->split) SABC_TRY(launch_split_propose(e, half, sub, nsub));
            cudaEvent_t a = nullptr, b = nullptr;
            if (e->kev) {
                SABC_CUDA(cudaEventCreate(&a)); SABC_CUDA(cudaEventCreate(&b));
                SABC_CUDA(cudaEventRecord(a, e->stream));
            }
            if (e->split) SABC_TRY(launch_split_simacc(e, half, sub, nsub)); else SABC_TRY(launch_update_half(e, half, sub, nsub));
            if (e->kev) { SABC_CUDA(cudaEventRecord(b, e->stream)); e->kev->push_back(a); e->kev->push_back(b); }
            if (e->pipe_last && e->pipe_after) SABC_TRY(e->pipe_after(half, sub));
        }
    }
    return e->split ? launch_split_stats(e) : 0;
}

// one population update, single GPU: every launch is unconditional, the resampling kernels
// return immediately unless the device-side trigger fired
static bool small_tail(const sabc_engine* e) { return !(e->flags & SABC_FLAG_GENERIC_TAIL) && !e->sharded() && e->proposal != PROP_RW && e->n_local <= 16384; }

static int launch_tail_small(sabc_engine* e) {
    TailArgs a{};
    Post1Args& p = a.post;
    p.ds = e->b_ds.p; p.rho_part = e->b_rho_part.p; p.part_ld = e->part_ld;
    int64_t o, n0, n1, t0, t1;
    halves(e, 0, o, n0, t0, t1); halves(e, 1, o, n1, t0, t1);
    p.groups0 = (n0 + CHUNK - 1) / CHUNK; p.groups1 = (n1 + CHUNK - 1) / CHUNK;
    p.scratch = e->b_scratch.p; p.scratch_ld = e->scratch_ld;
    p.S = e->S; p.n_global = e->N; p.resample = e->resample; p.decide = 1;
    FinishArgs& f = a.fin;
    f.ds = e->b_ds.p; f.hist = e->b_hist.p; f.S = e->S; f.n_eps = e->n_eps; f.algorithm = e->algorithm; f.n_global = e->N; f.v = e->v;
    a.pop = e->pop; a.tmp = e->tmp; a.n = e->n_local; a.D = e->D; a.S = e->S; a.delta = e->delta; a.seed = e->seed;
    a.q = e->b_q.p; a.tile_sum = e->b_tile_sum.p; a.tile_off = e->b_tile_off.p;
    k_tail_small<<<1, CHUNK, 0, e->stream>>>(a);
    SABC_CUDA(cudaGetLastError());
    return 0;
}

static int enqueue_iteration(sabc_engine* e) {
    SABC_TRY(enqueue_sweeps(e));
    if (small_tail(e)) return launch_tail_small(e);
    SABC_TRY(launch_post1(e, 1));
    SABC_TRY(launch_resample_local(e, 0));
    SABC_TRY(launch_update_proposal(e));
    SABC_TRY(launch_finish(e));
    return 0;
}

static int sync_state_from_device(sabc_engine* e, int64_t* n_rec_out) {
    DevState h;
    SABC_CUDA(cudaMemcpyAsync(&h, e->b_ds.p, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < e->n_eps; ++k) e->eps[k] = h.eps[k];
    e->n_accept = h.n_accept; e->n_resampling = h.n_resampling;
    if (n_rec_out) *n_rec_out = h.rec;
    if (h.error_flag & 1) return set_error(SABC_ERR_NEG_DISTANCE, "Negative distances are not allowed!");
    if (h.error_flag & 2) return set_error(SABC_ERR_UBAR_ZERO, "Division by zero - Mean u for a statistic <= eps()");
    if (h.error_flag & 4) return set_error(SABC_ERR_INVALID, "all resampling weights are zero: delta * u / mean(u) is beyond the range of the 32.32 fixed-point weights for every particle");
    return 0;
}

static int append_history(sabc_engine* e, int64_t n_rec) {
    if (n_rec <= 0) return 0;
    const int w = e->n_eps + 2 * e->S;
    std::vector<double> h((size_t)n_rec * w);
    SABC_CUDA(cudaMemcpy(h.data(), e->b_hist.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r < n_rec; ++r) {
        const double* rec = h.data() + r * w;
        e->eps_h.insert(e->eps_h.end(), rec, rec + e->n_eps);
        e->u_h.insert(e->u_h.end(), rec + e->n_eps, rec + e->n_eps + e->S);
        e->rho_h.insert(e->rho_h.end(), rec + e->n_eps + e->S, rec + w);
    }
    return 0;
}

static int ensure_hist(sabc_engine* e, int64_t n_rec) {
    const int w = e->n_eps + 2 * e->S;
    if (n_rec > e->hist_cap) {
        const int64_t cap = std::max<int64_t>(n_rec, 1024);
        SABC_CUDA(e->b_hist.alloc((size_t)cap * w));
        e->hist_cap = cap;
        // the captured graph holds the old buffer address in its kernel arguments
        if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    }
    return 0;
}

// Host-side plan of the surplus exchange of the multi-GPU resampling.  The selected particles of rank g occupy the
// global slots [C_g, C_g + c_g) (C = exclusive prefix of counts); rank d owns the slots [d n, (d+1) n).
extern "C" int sabc_mg_exchange_plan(const int64_t* counts, int32_t world, int64_t n_local, int32_t me, int64_t* send_off,
                                     int64_t* send_cnt, int64_t* recv_off, int64_t* recv_cnt) {
    if (!counts || world < 1 || me < 0 || me >= world) return set_error(SABC_ERR_INVALID, "bad exchange-plan argument");
    std::vector<int64_t> C(world + 1, 0);
    for (int g = 0; g < world; ++g) { if (counts[g] < 0) return set_error(SABC_ERR_INVALID, "negative count"); C[g + 1] = C[g] + counts[g]; }
    if (C[world] != n_local * world) return set_error(SABC_ERR_INVALID, "counts do not sum to the global particle number");
    for (int d = 0; d < world; ++d) {          // what I send to d (d == me: the part I keep)
        const int64_t lo = std::max(C[me], (int64_t)d * n_local), hi = std::min(C[me + 1], (int64_t)(d + 1) * n_local);
        send_off[d] = hi > lo ? lo - C[me] : 0; send_cnt[d] = hi > lo ? hi - lo : 0;
    }
    for (int g = 0; g < world; ++g) {          // what I receive from g, and where it lands in my slice
        const int64_t lo = std::max(C[g], (int64_t)me * n_local), hi = std::min(C[g + 1], (int64_t)(me + 1) * n_local);
        recv_off[g] = hi > lo ? lo - (int64_t)me * n_local : 0; recv_cnt[g] = hi > lo ? hi - lo : 0;
    }
    return 0;
}

static int ecdf_attach(sabc_engine* e, int j, DevBuf<double>* knots, int64_t L, int top_max);
static int top_max_for(int S);
extern "C" int sabc_multinomial_split(int64_t n_draws, const uint64_t* w, int32_t world, uint64_t seed, uint32_t resample_count,
                                      int64_t* counts_out) {
    if (!w || !counts_out || world < 1 || n_draws < 0) return set_error(SABC_ERR_INVALID, "bad multinomial-split argument");
    multinomial_split(n_draws, (const unsigned long long*)w, world, seed, resample_count, counts_out);
    return 0;
}

#include "multi_gpu.inl"

extern "C" {
static int get_population_ld(sabc_engine* e, double* theta, double* u, double* rho, int64_t ld);
static int set_population_ld(sabc_engine* e, const double* theta, const double* u, const double* rho, int64_t ld, const double* eps,
                             const int64_t counters[4]);
static int update_host_ld(sabc_engine* e, double* theta, double* u, double* rho, int64_t ld, double* eps, int64_t counters[4],
                          int64_t n_simulation, int64_t checkpoint_history);
}
#include "group.inl"

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int sabc_abi_version(void) { return SABC_ABI_VERSION; }
const char* sabc_last_error(void) { return last_error_buf(); }

int sabc_device_count(int* n) {
    SABC_CUDA(cudaGetDeviceCount(n));
    return 0;
}

int sabc_nccl_unique_id(void* out128) { return nccl_get_unique_id(out128); }

int sabc_host_alloc(void** out, int64_t bytes) {
    SABC_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return 0;
}
int sabc_host_free(void* p) {
    if (p) SABC_CUDA(cudaFreeHost(p));
    return 0;
}

int sabc_register_model(const void* vtable) {
    if (!vtable) return set_error(SABC_ERR_INVALID, "null vtable");
    ensure_builtin_models();
    const ModelVTable* vt = (const ModelVTable*)vtable;
    if (vt->struct_size != (uint32_t)sizeof(ModelVTable) || vt->abi_version != SABC_MODEL_VTABLE_VERSION)
        return set_error(SABC_ERR_INVALID, "model launch table was built against other kernel headers (size %u, version %u; this library: %u, %u)",
                         vt->struct_size, vt->abi_version, (unsigned)sizeof(ModelVTable), (unsigned)SABC_MODEL_VTABLE_VERSION);
    if (!vt->name || !vt->name[0]) return set_error(SABC_ERR_INVALID, "model launch table without a name");
    std::lock_guard<std::mutex> lk(g_reg_mutex);
    for (auto& m : registry())
        if (std::strcmp(m.name, vt->name) == 0) { const char* keep = m.name; m = *vt; m.name = keep; return 0; }   // same entry, same address
    registry_add(*vt);
    return 0;
}
int sabc_model_count(void) { ensure_builtin_models(); std::lock_guard<std::mutex> lk(g_reg_mutex); return (int)registry().size(); }
const char* sabc_model_name(int i) {
    ensure_builtin_models();
    std::lock_guard<std::mutex> lk(g_reg_mutex);
    return (i >= 0 && i < (int)registry().size()) ? registry()[i].name : nullptr;
}
int sabc_model_info(const char* name, int32_t* n_para, int32_t* n_stats) {
    const ModelVTable* m = find_model(name);
    if (!m) return set_error(SABC_ERR_INVALID, "unknown device model '%s'", name);
    if (n_para) *n_para = m->n_para;
    if (n_stats) *n_stats = m->n_stats;
    return 0;
}

int sabc_create(sabc_engine** out, const sabc_config* c) {
    if (!out || !c) return set_error(SABC_ERR_INVALID, "null argument");
    *out = nullptr;
    if (c->n_gpus > 1) return group_create(out, c);
    if (!(c->algorithm == SABC_ALG_SINGLE_EPS || c->algorithm == SABC_ALG_MULTI_EPS))
        return set_error(SABC_ERR_BAD_ALGORITHM, "Argument `algorithm` must be :multi_eps or :single_eps");
    if (c->proposal < 0 || c->proposal > 2) return set_error(SABC_ERR_BAD_PROPOSAL, "unknown proposal %d", c->proposal);
    if (c->proposal == SABC_PROP_RW && !(c->prop_par[0] > 0.0 && c->prop_par[0] <= 1.0))
        return set_error(SABC_ERR_BAD_PROPOSAL, "Mixing parameter `β` must be between zero and one.");
    if (!c->model_name) return set_error(SABC_ERR_INVALID, "model_name is null");
    const ModelVTable* model = find_model(c->model_name);
    if (!model) return set_error(SABC_ERR_INVALID, "unknown device model '%s'", c->model_name);
    if (model->n_para != c->n_para || model->n_stats != c->n_stats)
        return set_error(SABC_ERR_INVALID, "model '%s' has %d parameters and %d statistics, config says %d and %d",
                         c->model_name, model->n_para, model->n_stats, c->n_para, c->n_stats);
    if (c->n_para > MAX_D || c->n_stats > MAX_S || c->n_model_par > MAX_MODEL_PAR || c->n_model_par < 0)
        return set_error(SABC_ERR_INVALID, "dimension limits exceeded");
    const int world = c->world_size > 1 ? c->world_size : 1;
    if (c->n_particles % world != 0) return set_error(SABC_ERR_INVALID, "n_particles must be divisible by world_size");
    const int64_t n_local = c->n_particles / world;
    if (n_local < 4) return set_error(SABC_ERR_INVALID, "need at least 4 particles per GPU (each half >= 2)");
    if (c->n_particles > 0xffffffffLL) return set_error(SABC_ERR_INVALID, "n_particles exceeds 2^32-1");
    if (c->resample <= 0) return set_error(SABC_ERR_INVALID, "resample must be positive");
    if (c->ecdf_max_knots != 0 && (c->ecdf_max_knots < 2 || c->ecdf_max_knots > (1 << 20) || (int64_t)pow2_ceil(c->ecdf_max_knots + 2) * c->n_stats * 8 > 200 * 1024))
        return set_error(SABC_ERR_INVALID, "ecdf_max_knots must be 0 or at least 2 with pow2(K + 2) * %d statistics * 8 bytes <= 200 KB (the compressed tables live in shared memory, "
                         "padded to a power of two: K = 2^k - 2 wastes nothing)", c->n_stats);

    int ndev = 0;
    SABC_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return set_error(SABC_ERR_CUDA, "no CUDA device");
    int dev = c->device;
    if (dev < 0) SABC_CUDA(cudaGetDevice(&dev));
    SABC_CUDA(cudaSetDevice(dev));

    auto* e = new sabc_engine();
    e->N = c->n_particles; e->n_local = n_local; e->rank = world > 1 ? c->rank : 0; e->world = world;
    e->offset = (int64_t)e->rank * n_local;
    e->replicated = world > 1 && (c->flags & SABC_FLAG_MG_REPLICATED);
    if (e->replicated) {
        if (world > MAX_SUB) { delete e; return set_error(SABC_ERR_INVALID, "replicated mode supports at most %d ranks", MAX_SUB); }
        e->n_local = c->n_particles; e->offset = 0;          // every rank holds the whole population and simulates a share
    }
    e->D = c->n_para; e->S = c->n_stats; e->algorithm = c->algorithm; e->proposal = c->proposal;
    e->n_eps = c->algorithm == SABC_ALG_MULTI_EPS ? c->n_stats : 1;
    e->prop_par[0] = c->prop_par[0]; e->prop_par[1] = c->prop_par[1];
    e->v = c->v; e->delta = c->delta; e->resample = c->resample; e->seed = c->seed; e->flags = c->flags;
    e->ecdf_max_knots = c->ecdf_max_knots;
    if (e->flags & SABC_FLAG_TIME_KERNELS) e->flags |= SABC_FLAG_NO_GRAPH;
    if (e->replicated) e->flags |= SABC_FLAG_NO_GRAPH;
    e->split = model->heavy && !(e->flags & SABC_FLAG_FUSED);
    e->sort_work = e->split && model->key_bits > 0 && (e->flags & SABC_FLAG_SORT_WORK);
    e->device = dev; e->model = model;
    for (int k = 0; k < c->n_model_par; ++k) e->mp.v[k] = c->model_par[k];
    e->prior.n = e->D;
    for (int k = 0; k < e->D; ++k) {
        e->prior.kind[k] = c->prior_kind[k]; e->prior.p0[k] = c->prior_par[2 * k]; e->prior.p1[k] = c->prior_par[2 * k + 1];
        if (c->prior_kind[k] < SABC_PRIOR_UNIFORM || c->prior_kind[k] >= PRIOR_KIND_END) {
            delete e; return set_error(SABC_ERR_INVALID, "unknown prior kind %d", c->prior_kind[k]);
        }
        if (!prior_params_valid(c->prior_kind[k], c->prior_par[2 * k], c->prior_par[2 * k + 1])) {
            delete e; return set_error(SABC_ERR_INVALID, "prior component %d (kind %d): parameters (%g, %g) are outside the distribution's domain",
                                       k, c->prior_kind[k], c->prior_par[2 * k], c->prior_par[2 * k + 1]);
        }
    }
    prior_prepare(e->prior);

    auto fail = [&](int rc) { delete e; return rc; };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return fail(set_error(SABC_ERR_CUDA, "cudaGetDeviceProperties failed"));
    e->n_sm = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(set_error(SABC_ERR_CUDA, "stream creation failed"));

    const size_t n = (size_t)e->n_local;
    cudaError_t ce = cudaSuccess;
    auto A = [&](cudaError_t r) { if (ce == cudaSuccess) ce = r; };
    A(e->b_theta.alloc(n * e->D)); A(e->b_u.alloc(n * e->S)); A(e->b_rho.alloc(n * e->S)); A(e->b_lp.alloc(n));
    A(e->b_ttheta.alloc(n * e->D)); A(e->b_tu.alloc(n * e->S)); A(e->b_tlp.alloc(n));
    A(e->b_ds.alloc(1)); A(e->b_ecdf.alloc(MAX_S));
    A(e->b_q.alloc(n));
    { const size_t cap = n - n / 2; A(e->b_sp_theta.alloc(cap * e->D)); A(e->b_sp_lp.alloc(cap)); A(e->b_sp_lf.alloc(cap)); A(e->b_sp_idx.alloc(cap));
      A(e->b_sp_key.alloc(cap)); A(e->b_sp_perm.alloc(cap)); A(e->b_sp_hist.alloc(WORK_BUCKETS)); A(e->b_sp_off.alloc(WORK_BUCKETS));
      if (ce == cudaSuccess) ce = cudaMemsetAsync(e->b_sp_hist.p, 0, WORK_BUCKETS * sizeof(unsigned int), e->stream); }
    const int64_t n_tiles = ((int64_t)n + TILE - 1) / TILE;
    A(e->b_tile_sum.alloc((size_t)n_tiles)); A(e->b_tile_off.alloc((size_t)n_tiles));
    e->part_ld = ((int64_t)n + CHUNK - 1) / CHUNK + 1;
    A(e->b_rho_part.alloc((size_t)2 * e->S * e->part_ld));
    e->scratch_ld = 8;                                   // cta_treesum writes every level's partials: ceil(g/256) + ceil(.../256) + ...
    for (int64_t g = (e->part_ld + CHUNK - 1) / CHUNK; ; g = (g + CHUNK - 1) / CHUNK) { e->scratch_ld += g; if (g <= 1) break; }
    const int ncol = std::max(2 * e->S + 1, e->D * (e->D + 1) / 2 + e->D);
    A(e->b_scratch.alloc((size_t)ncol * e->scratch_ld));
    A(e->b_rw_part.alloc((size_t)(e->D * (e->D + 1) / 2) * e->part_ld * 2)); A(e->b_rw_sums.alloc(MAX_D * MAX_D));
    if (ce != cudaSuccess) return fail(set_error(SABC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(ce)));
    if (cudaMemsetAsync(e->b_ds.p, 0, sizeof(DevState), e->stream) != cudaSuccess) return fail(set_error(SABC_ERR_CUDA, "memset failed"));
    e->pop = PopView{e->b_theta.p, e->b_u.p, e->b_rho.p, e->b_lp.p, (int64_t)n};
    e->tmp = PopView{e->b_ttheta.p, e->b_tu.p, nullptr, e->b_tlp.p, (int64_t)n};
    e->grid_aux = e->n_sm * 8;

    if (world > 1) {
        if (!c->nccl_unique_id) return fail(set_error(SABC_ERR_INVALID, "world_size > 1 needs nccl_unique_id"));
        int rc = e->comm.init(c->nccl_unique_id, e->rank, world);
        if (rc == 0) rc = mg_warm_p2p(e);
        if (rc) return fail(rc);
        // work space of the global resampling, allocated once (a cudaMalloc inside the update loop costs tens of ms): O(N / G),
        // except for the strict variant, which walks all N global draws on every rank
        if (e->replicated) { *out = e; return 0; }
        cudaError_t ce2 = cudaSuccess;
        if (e->flags & SABC_FLAG_MG_STRICT_RESAMPLE) {
            const size_t Ng = (size_t)e->N, Nt = (Ng + TILE - 1) / TILE;
            ce2 = e->mg.F.alloc(Ng);
            if (ce2 == cudaSuccess) ce2 = e->mg.src.alloc(Ng);
            if (ce2 == cudaSuccess) ce2 = e->mg.tsum.alloc(Nt);
            if (ce2 == cudaSuccess) ce2 = e->mg.toff.alloc(Nt);
        }
        if (ce2 == cudaSuccess) ce2 = e->mg.scalar.alloc(1);
        if (ce2 == cudaSuccess) ce2 = e->mg.flag.alloc(1);
        if (ce2 == cudaSuccess) ce2 = e->mg.stats_send.alloc((size_t)4 * MAX_S + 1);
        if (ce2 == cudaSuccess) ce2 = e->mg.stats_all.alloc((size_t)(4 * MAX_S + 1) * world);
        if (ce2 == cudaSuccess) ce2 = cudaHostAlloc((void**)&e->mg.hold_slots, sizeof(int) * MG_LOOKAHEAD, cudaHostAllocDefault);
        for (int k = 0; k < MG_LOOKAHEAD && ce2 == cudaSuccess; ++k) ce2 = cudaEventCreateWithFlags(&e->mg.hold_ev[k], cudaEventDisableTiming);
        if (ce2 == cudaSuccess) ce2 = e->mg.sb.alloc((size_t)(e->D + e->S + 1) * (n + n / 4 + 4096));
        if (ce2 != cudaSuccess) return fail(set_error(SABC_ERR_CUDA, "multi-GPU work space allocation failed: %s", cudaGetErrorString(ce2)));
    }
    *out = e;
    return 0;
}

int sabc_destroy(sabc_engine* e) {
    if (!e) return 0;
    if (e->is_group()) return group_destroy(e);
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    delete e;
    return 0;
}

int sabc_set_tuning(sabc_engine* e, double v, double delta, int64_t resample, int32_t proposal, const double* prop_par) {
    if (!e || !prop_par) return set_error(SABC_ERR_INVALID, "null argument");
    if (e->is_group()) return group_set_tuning(e, v, delta, resample, proposal, prop_par);
    if (proposal < 0 || proposal > 2) return set_error(SABC_ERR_BAD_PROPOSAL, "unknown proposal %d", proposal);
    if (proposal == SABC_PROP_RW && !(prop_par[0] > 0.0 && prop_par[0] <= 1.0))
        return set_error(SABC_ERR_BAD_PROPOSAL, "Mixing parameter `β` must be between zero and one.");
    if (resample <= 0) return set_error(SABC_ERR_INVALID, "resample must be positive");
    const bool same = e->v == v && e->delta == delta && e->resample == resample && e->proposal == proposal &&
                      e->prop_par[0] == prop_par[0] && e->prop_par[1] == prop_par[1];
    if (same) return 0;                                       // nothing to re-capture: update_population! passes its keywords on every call
    const bool new_kernel = e->proposal != proposal;
    e->v = v; e->delta = delta; e->resample = resample;       // v, δ are validated by sabc_update like the reference (:261-262)
    e->proposal = proposal; e->prop_par[0] = prop_par[0]; e->prop_par[1] = prop_par[1];
    if (e->top_doubles > 0 && new_kernel) SABC_TRY(ecdf_finalize(e));       // occupancy / grid of the newly selected kernel
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }   // the captured arguments hold the old values
    return 0;
}

int sabc_local_particles(sabc_engine* e, int64_t* n_local, int64_t* offset) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (n_local) *n_local = e->n_local;
    if (offset) *offset = e->offset;
    return 0;
}

// initialization()  src/SimulatedAnnealingABC.jl:151-227
int sabc_init(sabc_engine* e) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (e->is_group()) return group_init(e);
    SABC_CUDA(cudaSetDevice(e->device));
    const int64_t n = e->n_local;
    DevState* ds = e->b_ds.p;
    SABC_CUDA(cudaMemsetAsync(ds, 0, sizeof(DevState), e->stream));
    k_begin<<<1, 1, 0, e->stream>>>(ds, 0, 1, 1);
    SABC_TRY(ensure_hist(e, 4));
    e->eps_h.clear(); e->u_h.clear(); e->rho_h.clear();
    e->n_resampling = 0;

    // prior sample + first simulations (:172-179), negative check (:185), Σρ for ρ_history[1] (:180)
    InitArgs ia{};
    ia.pop = e->pop; ia.n = n; ia.particle_base = (uint32_t)e->offset; ia.seed = e->seed; ia.ds = ds;
    ia.rho_part = e->b_rho_part.p; ia.part_ld = e->part_ld; ia.prior = e->prior; ia.mp = e->mp;
    const int64_t groups = (n + CHUNK - 1) / CHUNK;
    SABC_CUDA(e->model->launch_init(ia, (int)std::min<int64_t>(groups, (int64_t)e->n_sm * 8), e->stream));
    k_treesum_cols<<<e->S, CHUNK, 0, e->stream>>>(e->b_rho_part.p, e->part_ld, groups, e->b_scratch.p, e->scratch_ld,
                                                  &ds->rho_sum[0][0]);
    SABC_CUDA(cudaGetLastError());
    if (e->sharded()) SABC_TRY(mg_allreduce_f64(e, &ds->rho_sum[0][0], e->S));
    {
        int err = 0;
        SABC_CUDA(cudaMemcpyAsync(&err, &ds->error_flag, sizeof err, cudaMemcpyDeviceToHost, e->stream));
        SABC_CUDA(cudaStreamSynchronize(e->stream));
        if (e->sharded()) SABC_TRY(mg_any_flag(e, &err));
        if (err & 1) return set_error(SABC_ERR_NEG_DISTANCE, "Negative distances are not allowed!");
    }

    // build_cdf per statistic (:187) from the GLOBAL prior sample
    free_ecdf(e);
    {
        DevBuf<double> keys, gathered;
        DevBuf<unsigned char> cub_tmp;
        DevBuf<unsigned long long> cnt;
        SABC_CUDA(cnt.alloc(1));
        for (int j = 0; j < e->S; ++j) {
            const double* col = e->pop.rho + (int64_t)j * e->pop.ld;
            if (e->sharded() && e->ecdf_max_knots >= 2) {
                // compressed ECDF over a sharded population: the K global quantiles by a distributed selection, O(N / G) per rank
                int done = 0;
                SABC_TRY(mg_ecdf_quantiles(e, j, col, keys, cub_tmp, cnt, &done));
                if (done) continue;
            }
            SABC_CUDA(keys.ensure((size_t)e->N));
            if (e->sharded()) { SABC_CUDA(gathered.ensure((size_t)e->N)); SABC_TRY(mg_allgather_f64(e, col, gathered.p, n)); col = gathered.p; }
            SABC_TRY(ecdf_build_column(e, j, col, e->N, keys, cub_tmp, cnt));
        }
    }
    SABC_TRY(ecdf_finalize(e));

    // u = G(ρ) (:190-192) + exact Σu
    if (e->smem_update > 40 * 1024)
        SABC_CUDA(cudaFuncSetAttribute(k_transform, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_update));
    k_transform<<<(int)std::min<int64_t>(groups, (int64_t)e->n_sm * 8), CHUNK, e->smem_update, e->stream>>>(e->pop, n, e->S, e->b_ecdf.p, ds);
    SABC_CUDA(cudaGetLastError());
    // first resampling (:197), ε_0 (:200-204), history record 0 (:180,207-208)
    if (e->sharded()) SABC_TRY(mg_reduce_iteration_sums(e));
    k_decide<<<1, 32, 0, e->stream>>>(ds, e->S, e->N, e->resample);          // ū for the weights
    SABC_CUDA(cudaGetLastError());
    if (e->sharded()) SABC_TRY(mg_resample(e)); else SABC_TRY(launch_resample_local(e, 1));
    k_force_flag<<<1, 1, 0, e->stream>>>(ds, 1);
    SABC_TRY(launch_finish(e));
    int64_t n_rec = 0;
    SABC_TRY(sync_state_from_device(e, &n_rec));
    SABC_TRY(append_history(e, n_rec));
    // counters (:213-223)
    e->n_simulation = e->N; e->n_accept = 0; e->n_population_updates = 0;
    k_set_counters<<<1, 1, 0, e->stream>>>(ds, 0, e->n_resampling);
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    e->initialised = true;
    return 0;
}

// update_population!()  src/SimulatedAnnealingABC.jl:251-402
int sabc_update(sabc_engine* e, int64_t n_simulation, int64_t checkpoint_history) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (e->is_group()) return group_update(e, n_simulation, checkpoint_history);
    if (!e->initialised) return set_error(SABC_ERR_STATE, "sabc_update before sabc_init / sabc_set_population");
    if (!(e->v > 0.0)) return set_error(SABC_ERR_BAD_V, "Annealing speed `v` must be positive.");
    if (!(e->delta > 0.0)) return set_error(SABC_ERR_BAD_DELTA, "Resamping intensity `δ` must be positive.");
    SABC_CUDA(cudaSetDevice(e->device));
    if (checkpoint_history < 1) checkpoint_history = 1;
    const int64_t n_pop = n_simulation / e->N;                       // :275
    e->timing = sabc_timing{};
    if (n_pop <= 0) return 0;
    DevState* ds = e->b_ds.p;
    SABC_TRY(ensure_hist(e, n_pop / checkpoint_history + 2));
    k_begin<<<1, 1, 0, e->stream>>>(ds, (long long)(e->n_population_updates + 1), (long long)n_pop, (long long)checkpoint_history);
    SABC_CUDA(cudaGetLastError());
    if (e->sharded()) SABC_TRY(launch_update_proposal_mg(e)); else SABC_TRY(launch_update_proposal(e));   // :284

    std::vector<cudaEvent_t> kev;
    struct EventGuard {                                   // every exit path below, early error returns included, destroys the events
        cudaEvent_t a = nullptr, b = nullptr; std::vector<cudaEvent_t>* k;
        ~EventGuard() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); for (auto ev : *k) cudaEventDestroy(ev); }
    } guard{nullptr, nullptr, &kev};
    SABC_CUDA(cudaEventCreate(&guard.a)); SABC_CUDA(cudaEventCreate(&guard.b));
    const cudaEvent_t ev0 = guard.a, ev1 = guard.b;
    const bool time_kernels = (e->flags & SABC_FLAG_TIME_KERNELS) != 0;
    int rc = 0;
    const bool piped = (bool)e->pipe_before || (bool)e->pipe_after;
    if (e->replicated) {
        SABC_CUDA(cudaEventRecord(ev0, e->stream));
        if (time_kernels) e->kev = &kev;
        for (int64_t ix = 0; ix < n_pop && rc == 0; ++ix) rc = rep_iteration(e);
        e->kev = nullptr;
        SABC_CUDA(cudaEventRecord(ev1, e->stream));
    } else if (e->world > 1) {
        // Sharded: the host keeps up to MG_LOOKAHEAD updates enqueued and only looks at the `hold` word each of them leaves
        // behind (asynchronous copy into pinned memory).  A raised hold means that update's decision asked for a resampling: the
        // updates behind it ran as no-ops, the host runs the exchange, finishes that update and enqueues the others again.
        MgScratch& mg = e->mg;
        const bool use_graph = !(e->flags & SABC_FLAG_NO_GRAPH) && !piped && !mg.graph_failed;
        SABC_CUDA(cudaEventRecord(ev0, e->stream));
        if (time_kernels) e->kev = &kev;
        int64_t ix = 0, done = 0;
        while (done < n_pop && rc == 0) {
            while (ix < n_pop && ix - done < MG_LOOKAHEAD && rc == 0) {
                e->pipe_first = piped && ix == 0; e->pipe_last = piped && ix == n_pop - 1;
                if (use_graph && e->graph_exec) {
                    SABC_CUDA(cudaGraphLaunch(e->graph_exec, e->stream));
                } else if (use_graph && e->mg_warm) {           // NCCL has set its channels up: capture the update once
                    cudaGraph_t graph = nullptr;
                    cudaError_t ce = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
                    if (ce == cudaSuccess) {
                        rc = mg_enqueue_iteration(e);
                        ce = cudaStreamEndCapture(e->stream, &graph);
                        if (rc == 0 && ce == cudaSuccess) ce = cudaGraphInstantiate(&e->graph_exec, graph, 0);
                        if (graph) cudaGraphDestroy(graph);
                    }
                    if (rc != 0 || ce != cudaSuccess || !e->graph_exec) {   // capture refused (e.g. by the NCCL build): direct launches from now on
                        cudaGetLastError(); rc = 0; mg.graph_failed = true;
                        if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
                        rc = mg_enqueue_iteration(e);
                    } else {
                        SABC_CUDA(cudaGraphLaunch(e->graph_exec, e->stream));
                    }
                } else {
                    rc = mg_enqueue_iteration(e);
                    e->mg_warm = true;
                }
                if (rc) break;
                SABC_CUDA(cudaMemcpyAsync(&mg.hold_slots[ix % MG_LOOKAHEAD], &ds->hold, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
                SABC_CUDA(cudaEventRecord(mg.hold_ev[ix % MG_LOOKAHEAD], e->stream));
                ++ix;
            }
            if (rc) break;
            SABC_CUDA(cudaEventSynchronize(mg.hold_ev[done % MG_LOOKAHEAD]));
            if (mg.hold_slots[done % MG_LOOKAHEAD] != 0) {
                rc = mg_complete_held_iteration(e);
                ix = done + 1;
            }
            ++done;
        }
        e->kev = nullptr; e->pipe_first = e->pipe_last = false;
        SABC_CUDA(cudaEventRecord(ev1, e->stream));
    } else if ((e->flags & SABC_FLAG_NO_GRAPH) || piped) {
        SABC_CUDA(cudaEventRecord(ev0, e->stream));
        if (time_kernels) e->kev = &kev;
        for (int64_t ix = 0; ix < n_pop && rc == 0; ++ix) {
            e->pipe_first = piped && ix == 0; e->pipe_last = piped && ix == n_pop - 1;   // upload overlap / early download
            rc = enqueue_iteration(e);
        }
        e->kev = nullptr; e->pipe_first = e->pipe_last = false;
        SABC_CUDA(cudaEventRecord(ev1, e->stream));
    } else {
        if (!e->graph_exec) {
            cudaGraph_t graph = nullptr;
            SABC_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_iteration(e);
            cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            SABC_CUDA(ce);
            ce = cudaGraphInstantiate(&e->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            SABC_CUDA(ce);
        }
        SABC_CUDA(cudaEventRecord(ev0, e->stream));
        for (int64_t ix = 0; ix < n_pop; ++ix) SABC_CUDA(cudaGraphLaunch(e->graph_exec, e->stream));
        SABC_CUDA(cudaEventRecord(ev1, e->stream));
    }
    if (rc) return rc;
    int64_t n_rec = 0;
    rc = sync_state_from_device(e, &n_rec);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    e->timing.update_ms = ms;
    e->timing.kernel_launches = 2 * n_pop * e->pipe_nsub;
    e->timing.total_launches = ((int64_t)kernels_per_iteration(e) + (int64_t)(e->split ? 4 : 2) * (e->pipe_nsub - 1)) * n_pop;
    for (size_t k = 0; k + 1 < kev.size(); k += 2) {
        float t = 0.f; cudaEventElapsedTime(&t, kev[k], kev[k + 1]); e->timing.kernel_ms += t;
    }
    if (rc) return rc;
    SABC_TRY(append_history(e, n_rec));
    e->n_simulation += n_pop * e->N;                                 // :391
    e->n_population_updates += n_pop;                                // :394
    return 0;
}

// Copies of the row range [r0, r1) of a column-major matrix with ncol columns between host and device; the two sides may have
// different leading dimensions (a single-process multi-GPU handle passes slices of the caller's global arrays)
static int copy_rows(double* dst, int64_t dst_ld, const double* src, int64_t src_ld, int ncol, int64_t r0, int64_t r1,
                     cudaMemcpyKind kind, cudaStream_t st) {
    if (r1 <= r0) return 0;
    if (dst_ld == src_ld && r0 == 0 && r1 == dst_ld) { SABC_CUDA(cudaMemcpyAsync(dst, src, (size_t)dst_ld * ncol * sizeof(double), kind, st)); return 0; }
    // one strided copy: ncol rows of (r1-r0) doubles, pitch = one column
    SABC_CUDA(cudaMemcpy2DAsync(dst + r0, (size_t)dst_ld * sizeof(double), src + r0, (size_t)src_ld * sizeof(double),
                                (size_t)(r1 - r0) * sizeof(double), (size_t)ncol, kind, st));
    return 0;
}

static int get_population_ld(sabc_engine* e, double* theta, double* u, double* rho, int64_t ld) {
    SABC_CUDA(cudaSetDevice(e->device));
    const int64_t n = e->n_local;
    if (theta) SABC_TRY(copy_rows(theta, ld, e->pop.theta, n, e->D, 0, n, cudaMemcpyDeviceToHost, e->stream));
    if (u) SABC_TRY(copy_rows(u, ld, e->pop.u, n, e->S, 0, n, cudaMemcpyDeviceToHost, e->stream));
    if (rho) SABC_TRY(copy_rows(rho, ld, e->pop.rho, n, e->S, 0, n, cudaMemcpyDeviceToHost, e->stream));
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

static int set_state_scalars(sabc_engine* e, const double* eps, const int64_t counters[4]) {
    for (int k = 0; k < e->n_eps; ++k) e->eps[k] = eps[k];
    e->n_simulation = counters[0]; e->n_accept = counters[1]; e->n_resampling = counters[2]; e->n_population_updates = counters[3];
    SABC_CUDA(cudaMemcpyAsync(&e->b_ds.p->eps[0], e->eps, sizeof(double) * e->n_eps, cudaMemcpyHostToDevice, e->stream));
    k_set_counters<<<1, 1, 0, e->stream>>>(e->b_ds.p, e->n_accept, e->n_resampling);
    SABC_CUDA(cudaGetLastError());
    return 0;
}

static int set_population_ld(sabc_engine* e, const double* theta, const double* u, const double* rho, int64_t ld, const double* eps,
                             const int64_t counters[4]) {
    if (!theta || !u || !rho || !eps || !counters) return set_error(SABC_ERR_INVALID, "null argument");
    if (e->top_doubles == 0) return set_error(SABC_ERR_STATE, "no ECDF tables: call sabc_init or sabc_set_ecdf for every statistic first");
    SABC_CUDA(cudaSetDevice(e->device));
    const int64_t n = e->n_local;
    SABC_TRY(copy_rows(e->pop.theta, n, theta, ld, e->D, 0, n, cudaMemcpyHostToDevice, e->stream));
    SABC_TRY(copy_rows(e->pop.u, n, u, ld, e->S, 0, n, cudaMemcpyHostToDevice, e->stream));
    SABC_TRY(copy_rows(e->pop.rho, n, rho, ld, e->S, 0, n, cudaMemcpyHostToDevice, e->stream));
    k_recompute_lp<<<e->grid_aux, 256, 0, e->stream>>>(e->pop, 0, n, e->D, e->prior);
    SABC_TRY(set_state_scalars(e, eps, counters));
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    e->initialised = true;
    return 0;
}

int sabc_get_population(sabc_engine* e, double* theta, double* u, double* rho) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (e->is_group()) return group_get_population(e, theta, u, rho);
    return get_population_ld(e, theta, u, rho, e->n_local);
}

int sabc_set_population(sabc_engine* e, const double* theta, const double* u, const double* rho, const double* eps,
                        const int64_t counters[4]) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (e->is_group()) return group_set_population(e, theta, u, rho, eps, counters);
    return set_population_ld(e, theta, u, rho, e->n_local, eps, counters);
}

// update_population!(::SABCresult) with the result held in host buffers (leading dimension ld).  For large slices the transfers
// are pipelined with the two half-sweeps: the upload is issued in order of need, the first sweep starts as soon as theta of the
// second half (its partners) and the first sub-range of the first half have arrived, and every sub-range goes home as soon as
// its last sweep is done while the next one is simulated.
// (Measured and rejected in round 2: downloading only the rows that accepted, packed on the device and scattered by host
// threads -- 14-30 % of the rows change per update, nearly every cache line of the host arrays is touched, and the CPU
// scatter is slower than the DMA of the whole matrices: 4.56 ms per C4 step against 3.5; profiles/r2_notes.md.)
static int update_host_ld(sabc_engine* e, double* theta, double* u, double* rho, int64_t ld, double* eps, int64_t counters[4],
                          int64_t n_simulation, int64_t checkpoint_history) {
    if (!theta || !u || !rho || !eps || !counters) return set_error(SABC_ERR_INVALID, "null argument");
    SABC_CUDA(cudaSetDevice(e->device));
    const int64_t n = e->n_local, h0 = n / 2, n_pop = n_simulation / e->N;
    const bool pipe = n_pop >= 1 && n >= 32768 && !(e->flags & SABC_FLAG_NO_PIPELINE) && !e->replicated && e->top_doubles > 0 &&
                      e->v > 0.0 && e->delta > 0.0;
    cudaEvent_t a, b, c, d;
    SABC_CUDA(cudaEventCreate(&a)); SABC_CUDA(cudaEventCreate(&b)); SABC_CUDA(cudaEventCreate(&c)); SABC_CUDA(cudaEventCreate(&d));
    std::vector<cudaEvent_t> evs;
    auto cleanup = [&] { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); cudaEventDestroy(d); for (auto ev : evs) cudaEventDestroy(ev); };
    int rc = 0;
    if (!pipe) {
        SABC_CUDA(cudaEventRecord(a, e->stream));
        rc = set_population_ld(e, theta, u, rho, ld, eps, counters);
        if (!rc) { SABC_CUDA(cudaEventRecord(b, e->stream)); rc = sabc_update(e, n_simulation, checkpoint_history); }
        if (!rc) { SABC_CUDA(cudaEventRecord(c, e->stream)); rc = get_population_ld(e, theta, u, rho, ld); }
        if (!rc) rc = sabc_get_state(e, eps, counters);
        if (!rc) {
            SABC_CUDA(cudaEventRecord(d, e->stream));
            SABC_CUDA(cudaEventSynchronize(d));
            float t1 = 0, t2 = 0, t3 = 0;
            cudaEventElapsedTime(&t1, a, b); cudaEventElapsedTime(&t2, c, d); cudaEventElapsedTime(&t3, a, d);
            e->timing.h2d_ms = t1; e->timing.d2h_ms = t2; e->timing.host_ms = t3;
        }
        cleanup();
        return rc;
    }
    if (!e->s_in) SABC_CUDA(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    if (!e->s_out) SABC_CUDA(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
    // 2 sub-ranges per half measured best on B200 (4.70 ms vs 4.85 at 1, 5.9 at 3, 6.8 at 4 per C4 step): the simulation
    // kernel needs ~150 k items to fill the GPU, smaller sub-sweeps only add latency-bound tails.  SABC_PIPE_NSUB overrides.
    const int nsub_env = getenv("SABC_PIPE_NSUB") ? atoi(getenv("SABC_PIPE_NSUB")) : 2;
    const int nsub = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(nsub_env, (int)MAX_SUB), h0 / 32768));
    cudaEvent_t evT1, evUp[2][MAX_SUB], evDn[2][MAX_SUB];
    auto mk = [&](cudaEvent_t& ev) { cudaError_t r = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); if (r == cudaSuccess) evs.push_back(ev); return r; };
    bool ev_ok = mk(evT1) == cudaSuccess;
    for (int h = 0; h < 2; ++h) for (int j = 0; j < nsub; ++j) ev_ok = ev_ok && mk(evUp[h][j]) == cudaSuccess && mk(evDn[h][j]) == cudaSuccess;
    if (!ev_ok) { cleanup(); return set_error(SABC_ERR_CUDA, "event creation failed"); }
    const auto H2D = cudaMemcpyHostToDevice; const auto D2H = cudaMemcpyDeviceToHost;
    auto rows = [&](int half, int sub, int64_t& r0, int64_t& r1) {       // absolute row range of a sub-range
        int64_t off, an, t0, t1;
        halves(e, half, off, an, t0, t1);
        sub_range(an, sub, nsub, r0, r1);
        r0 += off; r1 += off;
    };
    // upload order = order of need: theta of the second half (partners of the first sweep), then per sub-range of the first
    // half its rows of theta, u, rho; then the rows of u, rho of the second half
    SABC_CUDA(cudaEventRecord(a, e->s_in));
    rc = copy_rows(e->pop.theta, n, theta, ld, e->D, h0, n, H2D, e->s_in);
    if (!rc) { SABC_CUDA(cudaEventRecord(evT1, e->s_in)); }
    for (int half = 0; half < 2 && !rc; ++half)
        for (int j = 0; j < nsub && !rc; ++j) {
            int64_t r0, r1; rows(half, j, r0, r1);
            if (half == 0) rc = copy_rows(e->pop.theta, n, theta, ld, e->D, r0, r1, H2D, e->s_in);
            if (!rc) rc = copy_rows(e->pop.u, n, u, ld, e->S, r0, r1, H2D, e->s_in);
            if (!rc) rc = copy_rows(e->pop.rho, n, rho, ld, e->S, r0, r1, H2D, e->s_in);
            if (!rc) { SABC_CUDA(cudaEventRecord(evUp[half][j], e->s_in)); }
        }
    if (rc) { cudaStreamSynchronize(e->s_in); cleanup(); return rc; }
    SABC_CUDA(cudaEventRecord(b, e->s_in));
    // state scalars; cached log-prior of the second half as soon as its theta is there
    SABC_CUDA(cudaStreamWaitEvent(e->stream, evT1, 0));
    k_recompute_lp<<<e->grid_aux, 256, 0, e->stream>>>(e->pop, h0, n, e->D, e->prior);
    rc = set_state_scalars(e, eps, counters);
    if (rc) { cudaStreamSynchronize(e->s_in); cleanup(); return rc; }
    if (e->proposal == PROP_RW)        // update_proposal! (:284) reads the whole population before the first sweep
        SABC_CUDA(cudaStreamWaitEvent(e->stream, evUp[1][nsub - 1], 0));
    e->initialised = true;
    const int64_t n_res_before = e->n_resampling;
    e->pipe_nsub = nsub;
    e->pipe_before = [&](int half, int sub) -> int {
        SABC_CUDA(cudaStreamWaitEvent(e->stream, evUp[half][sub], 0));
        if (half == 0) {
            int64_t r0, r1; rows(half, sub, r0, r1);
            if (r1 > r0) k_recompute_lp<<<e->grid_aux, 256, 0, e->stream>>>(e->pop, r0, r1, e->D, e->prior);
        }
        return 0;
    };
    // after its (sub-)sweep of the last update a row range is final (unless a resampling follows): it goes home while
    // the next sub-ranges are being simulated
    e->pipe_after = [&](int half, int sub) -> int {
        int64_t r0, r1; rows(half, sub, r0, r1);
        SABC_CUDA(cudaEventRecord(evDn[half][sub], e->stream));
        SABC_CUDA(cudaStreamWaitEvent(e->s_out, evDn[half][sub], 0));
        SABC_TRY(copy_rows(theta, ld, e->pop.theta, n, e->D, r0, r1, D2H, e->s_out));
        SABC_TRY(copy_rows(u, ld, e->pop.u, n, e->S, r0, r1, D2H, e->s_out));
        return copy_rows(rho, ld, e->pop.rho, n, e->S, r0, r1, D2H, e->s_out);
    };
    rc = sabc_update(e, n_simulation, checkpoint_history);          // blocks until the updates are done
    e->pipe_before = nullptr; e->pipe_after = nullptr; e->pipe_nsub = 1;
    const sabc_timing t_upd = e->timing;
    if (!rc) {
        SABC_CUDA(cudaEventRecord(c, e->s_out));
        if (e->n_resampling != n_res_before) {                      // theta, u of every row changed after the sweeps
            rc = copy_rows(theta, ld, e->pop.theta, n, e->D, 0, n, D2H, e->s_out);
            if (!rc) rc = copy_rows(u, ld, e->pop.u, n, e->S, 0, n, D2H, e->s_out);
        }
        if (!rc) rc = sabc_get_state(e, eps, counters);
    }
    if (!rc) {
        SABC_CUDA(cudaEventRecord(d, e->s_out));
        SABC_CUDA(cudaEventSynchronize(d));
        SABC_CUDA(cudaStreamSynchronize(e->s_in));
        float t1 = 0, t2 = 0, t3 = 0;
        cudaEventElapsedTime(&t1, a, b); cudaEventElapsedTime(&t2, c, d); cudaEventElapsedTime(&t3, a, d);
        e->timing = t_upd;
        e->timing.h2d_ms = t1; e->timing.d2h_ms = t2; e->timing.host_ms = t3;
    } else {
        cudaStreamSynchronize(e->s_in); cudaStreamSynchronize(e->s_out);
    }
    cleanup();
    return rc;
}

int sabc_update_host(sabc_engine* e, double* theta, double* u, double* rho, double* eps, int64_t counters[4],
                     int64_t n_simulation, int64_t checkpoint_history) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (e->is_group()) return group_update_host(e, theta, u, rho, eps, counters, n_simulation, checkpoint_history);
    return update_host_ld(e, theta, u, rho, e->n_local, eps, counters, n_simulation, checkpoint_history);
}

int sabc_get_state(sabc_engine* e, double* eps, int64_t counters[4]) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (eps) for (int k = 0; k < e->n_eps; ++k) eps[k] = e->eps[k];
    if (counters) { counters[0] = e->n_simulation; counters[1] = e->n_accept; counters[2] = e->n_resampling; counters[3] = e->n_population_updates; }
    return 0;
}
int sabc_history_len(sabc_engine* e, int64_t* n_records) {
    if (!e || !n_records) return set_error(SABC_ERR_INVALID, "null argument");
    if (e->is_group()) e = e->children[0];
    *n_records = (int64_t)(e->eps_h.size() / (size_t)e->n_eps);
    return 0;
}
int sabc_get_history(sabc_engine* e, double* eps_h, double* u_h, double* rho_h) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (e->is_group()) e = e->children[0];
    if (eps_h) std::memcpy(eps_h, e->eps_h.data(), e->eps_h.size() * sizeof(double));
    if (u_h) std::memcpy(u_h, e->u_h.data(), e->u_h.size() * sizeof(double));
    if (rho_h) std::memcpy(rho_h, e->rho_h.data(), e->rho_h.size() * sizeof(double));
    return 0;
}
int sabc_get_ecdf(sabc_engine* e, int32_t stat, double* knots_out, int64_t* L) {
    if (e && e->is_group()) e = e->children[0];
    if (!e || stat < 0 || stat >= e->S || e->h_ecdf[stat].L == 0) return set_error(SABC_ERR_INVALID, "no such ECDF");
    if (L) *L = e->h_ecdf[stat].L;
    if (knots_out) {
        SABC_CUDA(cudaSetDevice(e->device));
        SABC_CUDA(cudaMemcpy(knots_out, e->h_ecdf[stat].knots, (size_t)e->h_ecdf[stat].L * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
}
int sabc_set_ecdf(sabc_engine* e, int32_t stat, const double* knots, int64_t L) {
    if (!e || stat < 0 || stat >= e->S || !knots || L < 3) return set_error(SABC_ERR_INVALID, "bad ECDF argument");
    if (knots[0] != 0.0) return set_error(SABC_ERR_INVALID, "knots[0] must be 0 (values = [0; sort(x); ...], src/cdf_estimators.jl:33)");
    if (e->is_group()) return group_set_ecdf(e, stat, knots, L);
    SABC_CUDA(cudaSetDevice(e->device));
    auto* b = new DevBuf<double>();
    e->ecdf_bufs.push_back(b);
    SABC_CUDA(b->alloc((size_t)L + ECDF_PAD));
    SABC_CUDA(cudaMemcpyAsync(b->p, knots, (size_t)L * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    k_fill_inf<<<1, 32, 0, e->stream>>>(b->p + L, ECDF_PAD);
    SABC_TRY(ecdf_attach(e, stat, b, L, top_max_for(e->S)));
    bool all = true;
    for (int j = 0; j < e->S; ++j) all = all && e->h_ecdf[j].L > 0;
    if (all) SABC_TRY(ecdf_finalize(e));
    return 0;
}
int sabc_get_timing(sabc_engine* e, sabc_timing* out) {
    if (!e || !out) return set_error(SABC_ERR_INVALID, "null argument");
    *out = e->timing;
    return 0;
}
int sabc_update_kernel_info(sabc_engine* e, int* grid, int* block, int* smem_bytes, int* blocks_per_sm) {
    if (!e) return set_error(SABC_ERR_INVALID, "null engine");
    if (e->is_group()) e = e->children[0];
    if (grid) *grid = e->split ? e->grid_simacc : e->grid_update;
    if (block) *block = CHUNK;
    if (smem_bytes) *smem_bytes = (int)e->smem_update;
    if (blocks_per_sm) *blocks_per_sm = e->split ? e->bps_simacc : e->bps_update;
    return 0;
}

}  // extern "C"
