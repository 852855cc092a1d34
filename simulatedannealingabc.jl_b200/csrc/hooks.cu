// hooks.cu -- parity hooks of the C ABI: each runs the ENGINE'S OWN device functions on
// caller-supplied host arrays (copy in, one kernel, copy out), so tests can compare single steps
// of the path bit-for-bit with the oracle.  No CPU computation happens here.
#include "ecdf_index.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cstring>
#include <vector>

using namespace sabc;

namespace {

template <class T>
int upload(DevBuf<T>& b, const T* h, size_t n) {
    SABC_CUDA(b.alloc(n));
    if (n) SABC_CUDA(cudaMemcpy(b.p, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
template <class T>
int download(T* h, const DevBuf<T>& b, size_t n) {
    SABC_CUDA(cudaDeviceSynchronize());
    if (n) SABC_CUDA(cudaMemcpy(h, b.p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return 0;
}
int grid_for(int64_t n, int block = 256) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, 148 * 16)); }

__global__ void k_detmath(int op, const double* x, int64_t n, double* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s, c;
        switch (op) {
            case 0: out[i] = det_log(x[i]); break;
            case 1: out[i] = det_exp(x[i]); break;
            case 2: det_sincos2pi(x[i], s, c); out[i] = s; break;
            case 3: det_sincos2pi(x[i], s, c); out[i] = c; break;
            default: out[i] = det_logfact(x[i]); break;
        }
    }
}
__global__ void k_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    const U64x2 w = philox4x32_10(c0, c1, c2, c3, k0, k1);
    out[0] = (uint32_t)w.a; out[1] = (uint32_t)(w.a >> 32); out[2] = (uint32_t)w.b; out[3] = (uint32_t)(w.b >> 32);
}
__global__ void k_poisson(const double* lam, int64_t n, uint64_t seed, uint64_t sweep, int64_t* k_out, uint32_t* blocks) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        Stream st(seed, (uint32_t)i, sweep, KIND_MODEL);
        k_out[i] = poisson(lam[i], st);
        blocks[i] = st.next;
    }
}
// statistics of the two PTRS acceptance filters against the exact test (philox.cuh), over `attempts` candidates
__global__ void k_ptrs_filter_check(const double* lam, int n_lam, int64_t attempts, uint64_t seed, unsigned long long* counts,
                                    unsigned long long* ratio_bits) {
    unsigned long long c[5] = {0, 0, 0, 0, 0};
    double r1 = 0.0, r2 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < attempts; i += (int64_t)gridDim.x * blockDim.x) {
        const double l = lam[i % n_lam];
        Stream st(seed, (uint32_t)i, (uint64_t)(i >> 32), KIND_MODEL);
        double kf, num = 0.0, den = 0.0;
        if (ptrs_candidate(l, st.draw(), kf, num, den) != 2) continue;
        const bool ex = ptrs_exact(l, kf, num, den);
        const double Tex = ((-l + kf * det_log(l)) - det_logfact(kf)) - det_log(num / den);
        float T1, E1; double T2, E2;
        const int d1 = ptrs_filter_mufu(l, kf, num, den, T1, E1);
        const int d2 = ptrs_filter(l, kf, num, den, T2, E2);
        c[0]++;
        c[1] += d1 == 0; c[2] += d1 == 0 && d2 == 0;
        c[3] += d1 != 0 && (d1 > 0) != ex; c[4] += d2 != 0 && (d2 > 0) != ex;
        if (E1 > 0.0f) { const double r = fabs((double)T1 - Tex) / (double)E1; if (r > r1) r1 = r; }
        if (E2 > 0.0) { const double r = fabs(T2 - Tex) / E2; if (r > r2) r2 = r; }
    }
    for (int j = 0; j < 5; ++j) if (c[j]) atomicAdd(counts + j, c[j]);
    atomicMax(ratio_bits + 0, (unsigned long long)__double_as_longlong(r1));   // non-negative doubles order like integers
    atomicMax(ratio_bits + 1, (unsigned long long)__double_as_longlong(r2));
}
__global__ void k_accept(int64_t m, int s, const double* uo, const double* un, const double* eps, int n_eps, const double* dlp,
                         const double* lf, const double* U, uint8_t* acc) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        if (dlp[i] == -dinf()) { acc[i] = (uint8_t)(det_log(U[i]) < -dinf()); continue; }     // :320-322
        acc[i] = (uint8_t)accept_rule(s, uo + i, m, un + i, m, eps, n_eps, dlp[i], lf[i], U[i]);
    }
}
__global__ void k_eps_single(double ubar, double v, double* out) { out[0] = eps_single(ubar, v); }
__global__ void k_eps_multi(const double* ubar, int s, double v, double* out, int* err) {
    const int i = threadIdx.x;
    if (i < s) { double e = 0.0; if (!eps_multi_one(ubar, s, i, v, e)) atomicOr(err, 1); out[i] = e; }
}
__global__ void k_set_ubar(DevState* ds, const double* ubar, int s) { if (threadIdx.x < s) ds->ubar[threadIdx.x] = ubar[threadIdx.x]; }
__global__ void k_mean_from_limbs(const unsigned long long* hi, const unsigned long long* lo, int64_t n, double* out) {
    out[0] = limbs_to_sum(hi[0], lo[0]) / (double)n;
}
template <int D, int PROP>
__global__ void k_propose(const double* act, int64_t n, const double* ina, int64_t M, const double* chol, double p0, double p1,
                          uint64_t seed, uint32_t pbase, uint64_t sweep, double* out, double* lf_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double th[D], thp[D], lf;
    for (int c = 0; c < D; ++c) th[c] = act[c * n + i];
    const uint32_t pid = pbase + (uint32_t)i;
    const CtrlWords cw = ctrl_words(seed, pid, sweep);
    const InactiveGather<D> P{ina, M};
    if (PROP == PROP_DE) propose_de<D>(th, P, M, p0, p1, cw, Stream(seed, pid, sweep, KIND_CTRL), thp, lf);
    else if (PROP == PROP_STRETCH) propose_stretch<D>(th, P, M, p0, cw, thp, lf);
    else propose_rw<D>(th, chol, seed, pid, sweep, thp, lf);
    for (int c = 0; c < D; ++c) out[c * n + i] = thp[c];
    lf_out[i] = lf;
}
template <int D>
int propose_dispatch(int proposal, const double* act, int64_t n, const double* ina, int64_t M, const double* chol, double p0,
                     double p1, uint64_t seed, uint32_t pbase, uint64_t sweep, double* out, double* lf) {
    const int grid = (int)((n + 127) / 128);
    switch (proposal) {
        case PROP_DE: k_propose<D, PROP_DE><<<grid, 128>>>(act, n, ina, M, chol, p0, p1, seed, pbase, sweep, out, lf); break;
        case PROP_STRETCH: k_propose<D, PROP_STRETCH><<<grid, 128>>>(act, n, ina, M, chol, p0, p1, seed, pbase, sweep, out, lf); break;
        default: k_propose<D, PROP_RW><<<grid, 128>>>(act, n, ina, M, chol, p0, p1, seed, pbase, sweep, out, lf); break;
    }
    SABC_CUDA(cudaGetLastError());
    return 0;
}

// build the staged index for a knot table already on the device
struct HookEcdf {
    EcdfStat st{};
    std::vector<DevBuf<double>*> levels;
    DevBuf<EcdfStat> d_st;
    ~HookEcdf() { for (auto* b : levels) delete b; }
    int attach(const double* d_knots, int64_t L, int top_max) {
        SABC_TRY(ecdf_build_index(st, d_knots, L, top_max, levels, nullptr));
        st.top_off = 0;
        SABC_CUDA(d_st.alloc(1));
        SABC_CUDA(cudaMemcpy(d_st.p, &st, sizeof st, cudaMemcpyHostToDevice));
        return 0;
    }
};

}  // namespace

extern "C" {

int sabc_detmath(int32_t op, const double* x, int64_t n, double* out) {
    DevBuf<double> dx, dout;
    SABC_TRY(upload(dx, x, (size_t)n)); SABC_CUDA(dout.alloc((size_t)n));
    k_detmath<<<grid_for(n), 256>>>(op, dx.p, n, dout.p);
    SABC_CUDA(cudaGetLastError());
    return download(out, dout, (size_t)n);
}

int sabc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    DevBuf<uint32_t> d;
    SABC_CUDA(d.alloc(4));
    k_philox<<<1, 1>>>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], d.p);
    SABC_CUDA(cudaGetLastError());
    return download(out, d, 4);
}

int sabc_poisson(const double* lam, int64_t n, uint64_t seed, uint64_t sweep, int64_t* k_out, uint32_t* blocks_out) {
    DevBuf<double> dl; DevBuf<int64_t> dk; DevBuf<uint32_t> db;
    SABC_TRY(upload(dl, lam, (size_t)n)); SABC_CUDA(dk.alloc((size_t)n)); SABC_CUDA(db.alloc((size_t)n));
    k_poisson<<<grid_for(n, 128), 128>>>(dl.p, n, seed, sweep, dk.p, db.p);
    SABC_CUDA(cudaGetLastError());
    SABC_TRY(download(k_out, dk, (size_t)n));
    if (blocks_out) SABC_TRY(download(blocks_out, db, (size_t)n));
    return 0;
}

int sabc_ptrs_filter_check(const double* lam, int32_t n_lam, int64_t attempts, uint64_t seed, int64_t counts_out[5],
                           double ratio_out[2]) {
    if (!lam || n_lam < 1 || attempts < 0 || !counts_out || !ratio_out) return set_error(SABC_ERR_INVALID, "bad argument");
    DevBuf<double> dl; DevBuf<unsigned long long> dc;
    SABC_TRY(upload(dl, lam, (size_t)n_lam)); SABC_CUDA(dc.alloc(7));
    SABC_CUDA(cudaMemset(dc.p, 0, 7 * sizeof(unsigned long long)));
    k_ptrs_filter_check<<<148 * 8, 256>>>(dl.p, n_lam, attempts, seed, dc.p, dc.p + 5);
    SABC_CUDA(cudaGetLastError());
    unsigned long long h[7];
    SABC_TRY(download(h, dc, 7));
    for (int j = 0; j < 5; ++j) counts_out[j] = (int64_t)h[j];
    for (int j = 0; j < 2; ++j) { double r; memcpy(&r, &h[5 + j], 8); ratio_out[j] = r; }
    return 0;
}


// build_cdf(::AbstractVector)  src/cdf_estimators.jl:23-44
int sabc_ecdf_build(const double* dist, int64_t n, double* knots_out, int64_t* L) {
    if (!dist || !knots_out || !L || n < 1) return set_error(SABC_ERR_INVALID, "bad argument");
    DevBuf<double> dx, keys, knots; DevBuf<unsigned long long> cnt; DevBuf<unsigned char> tmp;
    SABC_TRY(upload(dx, dist, (size_t)n));
    SABC_CUDA(keys.alloc((size_t)n)); SABC_CUDA(knots.alloc((size_t)n + 2)); SABC_CUDA(cnt.alloc(1));
    SABC_CUDA(cudaMemset(cnt.p, 0, sizeof(unsigned long long)));
    k_mark_positive<<<grid_for(n), 256>>>(dx.p, n, keys.p, cnt.p);
    SABC_CUDA(cudaGetLastError());
    size_t bytes = 0;
    SABC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys.p, knots.p + 1, n));
    SABC_CUDA(tmp.alloc(bytes));
    SABC_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, keys.p, knots.p + 1, n));
    unsigned long long n_pos = 0;
    SABC_CUDA(cudaMemcpy(&n_pos, cnt.p, sizeof n_pos, cudaMemcpyDeviceToHost));
    if (n_pos == 0) return set_error(SABC_ERR_NO_POSITIVE, "build_cdf: no positive distance");
    k_ecdf_ends<<<1, 1>>>(knots.p, (int64_t)n_pos);
    SABC_CUDA(cudaGetLastError());
    *L = (int64_t)n_pos + 2;
    return download(knots_out, knots, (size_t)*L);
}

// cdfs_dist_prior(rho)  src/cdf_estimators.jl:68-70, through the same staged multi-level index as the engine
int sabc_ecdf_transform(const double* knots, int64_t L, const double* rho, int64_t m, double* u_out) {
    if (!knots || !rho || !u_out || L < 3) return set_error(SABC_ERR_INVALID, "bad argument");
    if (knots[0] != 0.0) return set_error(SABC_ERR_INVALID, "knots[0] must be 0 (values = [0; sort(x); ...], src/cdf_estimators.jl:33)");
    DevBuf<double> dk, dr, du;
    SABC_CUDA(dk.alloc((size_t)L + ECDF_PAD));
    SABC_CUDA(cudaMemcpy(dk.p, knots, (size_t)L * sizeof(double), cudaMemcpyHostToDevice));
    k_fill_inf<<<1, 32>>>(dk.p + L, ECDF_PAD);
    SABC_TRY(upload(dr, rho, (size_t)m)); SABC_CUDA(du.alloc((size_t)m));
    HookEcdf h;
    SABC_TRY(h.attach(dk.p, L, 2048));
    const size_t smem = (size_t)h.st.top_pow2 * sizeof(double);
    k_transform1<<<grid_for(m), CHUNK, smem>>>(dr.p, m, h.d_st.p, du.p);
    SABC_CUDA(cudaGetLastError());
    return download(u_out, du, (size_t)m);
}

int sabc_accept_step(int64_t m, int32_t s, const double* u_old, const double* u_new, const double* eps, int32_t n_eps,
                     const double* dlogprior, const double* log_factor, const double* uniform, uint8_t* accept_out) {
    DevBuf<double> a, b, e, d, l, u; DevBuf<uint8_t> o;
    SABC_TRY(upload(a, u_old, (size_t)m * s)); SABC_TRY(upload(b, u_new, (size_t)m * s)); SABC_TRY(upload(e, eps, (size_t)n_eps));
    SABC_TRY(upload(d, dlogprior, (size_t)m)); SABC_TRY(upload(l, log_factor, (size_t)m)); SABC_TRY(upload(u, uniform, (size_t)m));
    SABC_CUDA(o.alloc((size_t)m));
    k_accept<<<grid_for(m), 256>>>(m, s, a.p, b.p, e.p, n_eps, d.p, l.p, u.p, o.p);
    SABC_CUDA(cudaGetLastError());
    return download(accept_out, o, (size_t)m);
}

int sabc_update_epsilon_single(double ubar, double v, double* eps_out) {
    DevBuf<double> o;
    SABC_CUDA(o.alloc(1));
    k_eps_single<<<1, 1>>>(ubar, v, o.p);
    SABC_CUDA(cudaGetLastError());
    return download(eps_out, o, 1);
}
int sabc_update_epsilon_multi(const double* ubar, int32_t s, double v, double* eps_out) {
    if (s < 1 || s > MAX_S) return set_error(SABC_ERR_INVALID, "bad s");
    DevBuf<double> u, o; DevBuf<int> err;
    SABC_TRY(upload(u, ubar, (size_t)s)); SABC_CUDA(o.alloc((size_t)s)); SABC_CUDA(err.alloc(1));
    SABC_CUDA(cudaMemset(err.p, 0, sizeof(int)));
    k_eps_multi<<<1, 32>>>(u.p, s, v, o.p, err.p);
    SABC_CUDA(cudaGetLastError());
    int herr = 0;
    SABC_TRY(download(&herr, err, 1));
    if (herr) return set_error(SABC_ERR_UBAR_ZERO, "Division by zero - Mean u for a statistic <= eps()");
    return download(eps_out, o, (size_t)s);
}

int sabc_resample_weights(const double* u, int64_t n, int32_t s, const double* ubar, double delta, uint64_t* q_out) {
    DevBuf<double> du, dub; DevBuf<unsigned long long> q, ts; DevBuf<DevState> ds;
    SABC_TRY(upload(du, u, (size_t)n * s)); SABC_TRY(upload(dub, ubar, (size_t)s));
    SABC_CUDA(q.alloc((size_t)n)); SABC_CUDA(ts.alloc((size_t)(n + TILE - 1) / TILE)); SABC_CUDA(ds.alloc(1));
    SABC_CUDA(cudaMemset(ds.p, 0, sizeof(DevState)));
    k_set_ubar<<<1, 32>>>(ds.p, dub.p, s);
    PopView pop{nullptr, du.p, nullptr, nullptr, n};
    k_weights<<<grid_for(n, TILE), CHUNK>>>(pop, n, s, delta, ds.p, q.p, ts.p, 1);
    SABC_CUDA(cudaGetLastError());
    SABC_CUDA(cudaDeviceSynchronize());
    SABC_CUDA(cudaMemcpy(q_out, q.p, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return 0;
}
int sabc_resample_indices(const uint64_t* qh, int64_t n, uint64_t seed, uint64_t resample_count, int64_t* idx_out) {
    DevBuf<unsigned long long> q, ts, to, tot; DevBuf<int64_t> idx; DevBuf<DevState> ds;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    SABC_CUDA(q.alloc((size_t)n));
    SABC_CUDA(cudaMemcpy(q.p, qh, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice));
    SABC_CUDA(ts.alloc((size_t)n_tiles)); SABC_CUDA(to.alloc((size_t)n_tiles)); SABC_CUDA(tot.alloc(1));
    SABC_CUDA(idx.alloc((size_t)n)); SABC_CUDA(ds.alloc(1));
    k_tile_sums<<<grid_for(n, TILE), CHUNK>>>(q.p, n, ts.p);
    k_scan_tiles<<<1, 1024>>>(ts.p, n_tiles, to.p, tot.p, ds.p, 1);
    k_prefix<<<grid_for(n, TILE), CHUNK>>>(q.p, n, to.p, ds.p, 1);
    SABC_CUDA(cudaGetLastError());
    unsigned long long W = 0;
    SABC_CUDA(cudaMemcpy(&W, tot.p, sizeof W, cudaMemcpyDeviceToHost));
    k_draw_indices<<<grid_for(n), 256>>>(q.p, n, W, seed, (uint32_t)resample_count, idx.p);
    SABC_CUDA(cudaGetLastError());
    return download(idx_out, idx, (size_t)n);
}
int sabc_exact_mean_u(const double* u, int64_t n, double* mean_out) {
    DevBuf<double> du, out; DevBuf<unsigned long long> acc;
    SABC_TRY(upload(du, u, (size_t)n)); SABC_CUDA(out.alloc(1)); SABC_CUDA(acc.alloc(2));
    SABC_CUDA(cudaMemset(acc.p, 0, 2 * sizeof(unsigned long long)));
    k_sum_u<<<grid_for(n), CHUNK>>>(du.p, n, n, 1, acc.p, acc.p + 1);
    k_mean_from_limbs<<<1, 1>>>(acc.p, acc.p + 1, n, out.p);
    SABC_CUDA(cudaGetLastError());
    return download(mean_out, out, 1);
}
int sabc_treesum(const double* x, int64_t n, double* sum_out) {
    DevBuf<double> dx, part, scratch, out;
    SABC_TRY(upload(dx, x, (size_t)n));
    const int64_t groups = (n + CHUNK - 1) / CHUNK;
    SABC_CUDA(part.alloc((size_t)groups)); SABC_CUDA(scratch.alloc((size_t)groups / CHUNK + 8)); SABC_CUDA(out.alloc(1));
    k_group_sums<<<grid_for(n), CHUNK>>>(dx.p, n, n, 1, part.p, groups);
    k_treesum_cols<<<1, CHUNK>>>(part.p, groups, groups, scratch.p, 0, out.p);
    SABC_CUDA(cudaGetLastError());
    return download(sum_out, out, 1);
}

int sabc_prior_logpdf(int32_t d, const int32_t* kind, const double* par, const double* theta, int64_t n, double* lp_out) {
    if (d < 1 || d > MAX_D) return set_error(SABC_ERR_INVALID, "bad d");
    PriorSpec p{};
    p.n = d;
    for (int c = 0; c < d; ++c) { p.kind[c] = kind[c]; p.p0[c] = par[2 * c]; p.p1[c] = par[2 * c + 1]; }
    prior_prepare(p);
    DevBuf<double> th, lp;
    SABC_TRY(upload(th, theta, (size_t)n * d)); SABC_CUDA(lp.alloc((size_t)n));
    PopView pop{th.p, nullptr, nullptr, lp.p, n};
    k_recompute_lp<<<grid_for(n), 256>>>(pop, 0, n, d, p);
    SABC_CUDA(cudaGetLastError());
    return download(lp_out, lp, (size_t)n);
}

int sabc_model_simulate(const char* model_name, const double* model_par, int32_t n_model_par, const double* theta, int64_t n,
                        uint64_t seed, uint32_t particle_base, uint64_t sweep, double* rho_out) {
    const ModelVTable* m = find_model(model_name);
    if (!m) return set_error(SABC_ERR_INVALID, "unknown device model '%s'", model_name);
    if (n_model_par < 0 || n_model_par > MAX_MODEL_PAR) return set_error(SABC_ERR_INVALID, "bad n_model_par");
    ModelPar mp{};
    for (int k = 0; k < n_model_par; ++k) mp.v[k] = model_par[k];
    DevBuf<double> th, rho;
    SABC_TRY(upload(th, theta, (size_t)n * m->n_para)); SABC_CUDA(rho.alloc((size_t)n * m->n_stats));
    SABC_CUDA(m->simulate(th.p, n, n, mp, seed, particle_base, sweep, rho.p, 0));
    return download(rho_out, rho, (size_t)n * m->n_stats);
}

int sabc_propose(int32_t proposal, const double* prop_par, int32_t d, const double* theta_active, int64_t n,
                 const double* theta_inactive, int64_t M, const double* chol, uint64_t seed, uint32_t particle_base,
                 uint64_t sweep, double* theta_out, double* log_factor_out) {
    if (d < 1 || d > 4) return set_error(SABC_ERR_INVALID, "sabc_propose hook supports d = 1..4");
    if (proposal < 0 || proposal > 2) return set_error(SABC_ERR_BAD_PROPOSAL, "unknown proposal");
    DevBuf<double> a, p, c, o, lf;
    SABC_TRY(upload(a, theta_active, (size_t)n * d)); SABC_TRY(upload(p, theta_inactive, (size_t)M * d));
    std::vector<double> ch(MAX_D * MAX_D, 0.0);
    if (chol) for (int k = 0; k < d * d; ++k) ch[k] = chol[k];
    SABC_TRY(upload(c, ch.data(), ch.size())); SABC_CUDA(o.alloc((size_t)n * d)); SABC_CUDA(lf.alloc((size_t)n));
    int rc = 0;
    switch (d) {
        case 1: rc = propose_dispatch<1>(proposal, a.p, n, p.p, M, c.p, prop_par[0], prop_par[1], seed, particle_base, sweep, o.p, lf.p); break;
        case 2: rc = propose_dispatch<2>(proposal, a.p, n, p.p, M, c.p, prop_par[0], prop_par[1], seed, particle_base, sweep, o.p, lf.p); break;
        case 3: rc = propose_dispatch<3>(proposal, a.p, n, p.p, M, c.p, prop_par[0], prop_par[1], seed, particle_base, sweep, o.p, lf.p); break;
        default: rc = propose_dispatch<4>(proposal, a.p, n, p.p, M, c.p, prop_par[0], prop_par[1], seed, particle_base, sweep, o.p, lf.p); break;
    }
    if (rc) return rc;
    SABC_TRY(download(theta_out, o, (size_t)n * d));
    return download(log_factor_out, lf, (size_t)n);
}

}  // extern "C"
