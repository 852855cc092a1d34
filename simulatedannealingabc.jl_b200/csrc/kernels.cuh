// kernels.cuh -- the fused per-particle kernels of the SABC population update, templated on the
// model plug-in and the proposal kind, plus the small device functions shared with the parity
// hooks (accept rule, epsilon solvers, radix-256 tree sums).
//
// Reference path: the particle body of update_population! (src/SimulatedAnnealingABC.jl:308-331)
// and the prior-sample loop of initialization() (:172-179).  One thread owns one particle; a CTA
// of 256 threads owns one 256-particle group of the tree-sum spec.
#pragma once
#include "plugin.cuh"
#include <cuda_runtime.h>

namespace sabc {

// Device-resident algorithm state: everything that changes between population updates lives
// here, so one captured CUDA graph can be replayed for every iteration without host round trips.
constexpr int MAX_SUB = 8;                 // sub-ranges of a half-sweep (transfer pipelining of the host-buffer call)
constexpr int MAX_SLOTS = 2 * MAX_SUB;

struct DevState {
    // [u_hi, u_lo, n_acc_iter] and [r_hi, r_lo] are contiguous: each is one integer all-reduce on multi-GPU
    unsigned long long u_hi[MAX_S], u_lo[MAX_S];   // exact Σu limbs of the running iteration
    unsigned long long n_acc_iter;                 // accepts of the running iteration
    unsigned long long r_hi[MAX_S], r_lo[MAX_S];   // exact Σu limbs of the resampled population
    unsigned long long w_total;                    // Σ fixed-point weights (this GPU)
    double eps[MAX_S];
    double chol[MAX_D * MAX_D];                    // RandomWalk: Cholesky factor / sd
    double ubar[MAX_S];                            // column means of u (before resampling)
    double rho_sum[2][MAX_S];                      // tree sums of ρ per half
    double mom[MAX_D + MAX_D * MAX_D];             // RandomWalk: means, centred cross products
    long long t;                                   // global population-update number being executed
    long long ix, n_pop, checkpoint, rec, last_cp; // position inside the current update() call
    long long n_accept, n_resampling;
    int resample_flag, error_flag;
    // sharded multi-GPU: a global resampling is due.  The host runs the exchange; until it has, the rest of this update and every
    // update enqueued behind it are no-ops (each kernel of the iteration returns at once), so the host never has to wait for the
    // trigger before it enqueues the next update (multi_gpu.inl)
    int hold, hold_pad;
    unsigned int list_count[MAX_SLOTS], list_cursor[MAX_SLOTS];   // split path: work-list length / fetch cursor per (half, sub-range)
};

// ------------------------------------------------------------------------------------------------
// shared device functions
// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
// every kernel of the update sequence starts with this test: after an error (e.g. mean u <= eps() in update_epsilon_multi_eps,
// :107-109) the state stays as it was at the failing update, like the reference's error() aborting there, although the captured
// graph keeps replaying; `hold` is the multi-GPU resampling request (DevState)
SABC_D bool halted(const DevState* ds) { return (ds->hold | ds->error_flag) != 0; }
#endif

// Accept rule, src/SimulatedAnnealingABC.jl:318-324.  uo/un stride through column-major rows.
SABC_HD bool accept_rule(int s, const double* uo, int64_t ldo, const double* un, int64_t ldn, const double* eps,
                         int n_eps, double dlp, double log_factor, double U) {
    double S = 0.0;
    for (int j = 0; j < s; ++j) {
        const double t = (uo[j * ldo] - un[j * ldn]) / eps[n_eps == 1 ? 0 : j];
        S = (j == 0) ? t : S + t;
    }
    const double Lacc = (dlp + S) + log_factor;
    return det_log(U) < Lacc;
}

// update_epsilon_single_eps, :92-95: root of ε² + v ε^{3/2} − ū² by monotone Newton on s = sqrt(ε)
SABC_HD double eps_single(double ubar, double v) {
    if (ubar <= 2.220446049250313e-16) return 0.0;
    const double u2 = ubar * ubar;
    double s = sqrt(ubar);
    for (int it = 0; it < 64; ++it) {
        const double s2 = s * s, s3 = s2 * s;
        const double f = (s3 * s + v * s3) - u2;
        const double fp = 4.0 * s3 + (3.0 * v) * s2;
        const double sn = s - f / fp;
        if (!(sn < s)) break;
        s = sn;
    }
    return s * s;
}

SABC_HD double ipow(double x, int m) { double p = 1.0; for (int i = 0; i < m; ++i) p = p * x; return p; }
SABC_HD double powhalf(double x, int m) { return (m & 1) ? sqrt(x) * ipow(x, (m - 1) / 2) : ipow(x, m / 2); }

// update_epsilon_multi_eps, :100-117, component i.  Returns false if ū_i <= eps() (:107-109).
SABC_HD bool eps_multi_one(const double* ubar, int n, int i, double v, double& eps_out) {
    double cn = 1.0;
    for (int k = 2; k <= n + 1; ++k) cn = cn * (double)(n + 1 + k) / (double)k;
    const double ui = ubar[i];
    if (ui <= 2.220446049250313e-16) return false;
    double sumq = 0.0, prodq = 1.0;
    for (int j = 0; j < n; ++j) {
        const double q = ubar[j] / ui;
        const double t = powhalf(q, n);
        sumq = (j == 0) ? t : sumq + t;
        prodq = (j == 0) ? q : prodq * q;
    }
    const double num = 1.0 + sumq;
    const double den = ((cn * (double)(n + 1)) * powhalf(ui, n + 2)) * prodq;
    double beta = 1.0 / ui;
    for (int it = 0; it < 100; ++it) {
        double g, dg;
        if (fabs(beta) < 0.01) {
            const double b2 = beta * beta;
            g = 0.5 - beta * (1.0 / 12.0 - b2 * (1.0 / 720.0 - b2 * (1.0 / 30240.0)));
            dg = -(1.0 / 12.0) + b2 * (1.0 / 240.0 - b2 * (1.0 / 6048.0));
        } else {
            const double t = det_exp(-beta), omt = 1.0 - t;
            g = 1.0 / beta - t / omt;
            dg = t / (omt * omt) - 1.0 / (beta * beta);
        }
        const double bn = beta - (g - ui) / dg;
        const double diff = fabs(bn - beta);
        beta = bn;
        if (diff <= 4.0e-16 * fabs(bn)) break;
    }
    eps_out = 1.0 / (beta + (v * num) / den);
    return true;
}

#if defined(__CUDACC__)
// 32-lane butterfly of the tree-sum spec; lane 0 holds the group value
SABC_D double warp_tree(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, off);
    return v;
}
// one 256-group: called by all 256 threads of a CTA; result valid in thread 0
SABC_D double group256(double v, double* s_w /* 8 doubles */) {
    const double w = warp_tree(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = w;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0) { tot = s_w[0]; for (int k = 1; k < 8; ++k) tot = tot + s_w[k]; }
    __syncthreads();
    return tot;
}
// radix-256 tree sum of x[0..n) by ONE CTA of 256 threads; scratch needs ceil(n/256) doubles.
// Result valid in thread 0.
SABC_D double cta_treesum(const double* x, int64_t n, double* scratch, double* s_w) {
    if (n <= 0) return 0.0;
    const double* in = x;
    double* out = scratch;
    for (;;) {
        const int64_t g = (n + CHUNK - 1) / CHUNK;
        double last = 0.0;
        for (int64_t c = 0; c < g; ++c) {
            const int64_t i = c * CHUNK + threadIdx.x;
            const double r = group256(i < n ? in[i] : 0.0, s_w);
            if (threadIdx.x == 0) { if (g > 1) out[c] = r; last = r; }
        }
        if (g == 1) return last;
        __syncthreads();
        // ping-pong inside scratch: next level reads what was just written
        in = out; out = out + g; n = g;
    }
}
// warp sum of 31-bit limbs into 64 bits via two 16-bit hardware reductions
SABC_D unsigned long long warp_sum_u32(uint32_t x) {
    const uint32_t lo = __reduce_add_sync(0xffffffffu, x & 0xffffu);
    const uint32_t hi = __reduce_add_sync(0xffffffffu, x >> 16);
    return ((unsigned long long)hi << 16) + lo;
}

// ------------------------------------------------------------------------------------------------
// kernel arguments
// ------------------------------------------------------------------------------------------------
struct PopView {                 // structure-of-arrays particle state of one GPU (FP64, leading dim ld)
    double* theta;               // [D][ld]
    double* u;                   // [S][ld]
    double* rho;                 // [S][ld]
    double* lp;                  // [ld] cached logpdf(prior, θ_i)   (reference recomputes it, :318)
    int64_t ld;
};

struct UpdateArgs {
    PopView pop;
    int64_t act_off, act_n, ina_off, ina_n;   // active / inactive half (local indices)
    uint32_t particle_base;                   // global index of local particle 0 (Philox counter)
    int32_t half;
    uint64_t seed;
    RoundKeys rk;                             // Philox round keys of `seed` (simulation kernel's hot loop)
    int32_t slot;                             // work-list counter slot = half * MAX_SUB + sub-range
    DevState* ds;
    const EcdfStat* ecdf;                     // [S] descriptors in HBM
    double* rho_part;                         // [S][part_ld] per-group ρ sums of this half
    int64_t part_ld;
    int32_t n_eps;
    int32_t top_doubles;                      // staged ECDF index size (doubles)
    double prop0, prop1;                      // DE: γ0, σ_γ | Stretch: a | RW: β
    PriorSpec prior;
    ModelPar mp;
};

// Work list of the split ("heavy model") path: proposals that passed the prior check, compacted so that every lane
// of the simulation kernel has a simulation to run.
struct SplitScratch {
    double* theta;            // [D][cap] proposed θ′ in list order
    double* lp;               // [cap] logpdf(prior, θ′)
    double* lf;               // [cap] log_factor of the proposal
    uint32_t* idx;            // [cap] local index (within the active half) of the particle
    unsigned int* count;      // [MAX_SLOTS] entries per (half, sub-range)
    unsigned int* cursor;     // [MAX_SLOTS] dynamic work cursor of the simulation kernel
    int64_t cap;
    // optional bucketing of the work list by a model-supplied similarity key (lanes of a warp then simulate alike)
    uint32_t* key;            // [cap] bucket of entry q, or nullptr
    uint32_t* perm;           // [cap] simulation order: position -> entry
    unsigned int* hist;       // [WORK_BUCKETS] bucket sizes, then scatter cursors
    unsigned int* off;        // [WORK_BUCKETS] bucket offsets
};
constexpr int WORK_BUCKETS = 4096;

struct InitArgs {
    PopView pop;
    int64_t n;
    uint32_t particle_base;
    uint64_t seed;
    DevState* ds;
    double* rho_part;                         // [S][part_ld] per-group ρ sums
    int64_t part_ld;
    PriorSpec prior;
    ModelPar mp;
};

// a model that never draws normals declares `static constexpr int NO_NORMALS = 1` and its simulation kernel skips the table
template <class M, class = void> struct model_draws_normals { static constexpr bool value = true; };
template <class M> struct model_draws_normals<M, decltype((void)M::NO_NORMALS)> { static constexpr bool value = false; };

template <int D>
struct InactiveGather {
    const double* base; int64_t ld;
    SABC_D double operator()(int c, int64_t i) const { return base[c * ld + i]; }
};

// Stage the top index level of every statistic in shared memory with TMA bulk copies (cp.async.bulk, SASS UBLKCP): one
// elected thread posts the expected byte count on an mbarrier and issues one bulk copy per statistic; every thread then
// waits on the barrier's phase.  The staged levels are padded to a power of two >= 2 entries: sizes are multiples of 16 bytes.
SABC_D void stage_ecdf_top(const EcdfStat* ecdf, int S, double* s_top) {
#if defined(SABC_NO_TMA)
    for (int j = 0; j < S; ++j) {
        const EcdfStat& e = ecdf[j];
        double* dst = s_top + e.top_off;
        for (int i = threadIdx.x; i < e.top_pow2; i += blockDim.x) dst[i] = e.top[i];
    }
    __syncthreads();
#else
    __shared__ __align__(8) unsigned long long s_mbar;
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int j = 0; j < S; ++j) total += (uint32_t)ecdf[j].top_pow2 * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(total) : "memory");
        for (int j = 0; j < S; ++j) {
            const EcdfStat& e = ecdf[j];
            const uint32_t bytes = (uint32_t)e.top_pow2 * 8u;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_top + e.top_off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(e.top), "r"(bytes), "r"(mbar) : "memory");
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar) : "memory");
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// K4 update_half: propose -> prior -> simulate+distance -> ECDF -> Metropolis accept, fused with
// the partial reductions the ε update, the resampling weights and the history need.
// ------------------------------------------------------------------------------------------------
template <class M, int PROP>
__global__ void __launch_bounds__(CHUNK, M::FUSED_MIN_BLOCKS) update_half_kernel(const UpdateArgs a) {
    constexpr int D = M::D, S = M::S;
    extern __shared__ __align__(128) double s_top[];
    __shared__ unsigned long long s_acc[2 * S + 1];
    __shared__ double s_w[S][8];
    __shared__ double s_chol[PROP == PROP_RW ? D * D : 1];
    __shared__ ZigEntry s_zig[256];

    const int tid = threadIdx.x;
    if (halted(a.ds)) return;
    for (int k = tid; k < 2 * S + 1; k += CHUNK) s_acc[k] = 0ull;
    if (PROP == PROP_RW) for (int k = tid; k < D * D; k += CHUNK) s_chol[k] = a.ds->chol[k];
    const uint32_t zig = stage_zig(s_zig);
    stage_ecdf_top(a.ecdf, S, s_top);
    __syncthreads();

    const uint64_t sweep = 2ull * (uint64_t)a.ds->t + (uint64_t)a.half;
    double eps[S];
#pragma unroll
    for (int j = 0; j < S; ++j) eps[j] = a.ds->eps[a.n_eps == 1 ? 0 : j];

    const int64_t ld = a.pop.ld;
    const int64_t n_groups = (a.act_n + CHUNK - 1) / CHUNK;
    const InactiveGather<D> P{a.pop.theta + a.ina_off, ld};

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t il = grp * CHUNK + tid;
        const bool valid = il < a.act_n;
        const int64_t gi = a.act_off + il;
        double rp[S], up[S];
        bool acc = false;
        if (valid) {
            const uint32_t pid = a.particle_base + (uint32_t)gi;
            double th[D], thp[D], lf;
#pragma unroll
            for (int c = 0; c < D; ++c) th[c] = a.pop.theta[c * ld + gi];
            const CtrlWords cw = ctrl_words(a.seed, pid, sweep);
            if (PROP == PROP_DE) propose_de<D>(th, P, a.ina_n, a.prop0, a.prop1, cw, Stream(a.seed, pid, sweep, KIND_CTRL, nullptr, zig), thp, lf);
            else if (PROP == PROP_STRETCH) propose_stretch<D>(th, P, a.ina_n, a.prop0, cw, thp, lf);
            else propose_rw<D>(th, s_chol, a.seed, pid, sweep, thp, lf, zig);
            const double lpp = prior_logpdf<D>(a.prior, thp);
            if (lpp > -dinf()) {                                       // :314 (no simulation outside the support)
                Stream st(a.seed, pid, sweep, KIND_MODEL, nullptr, zig);
                M::sim(thp, a.mp, st, rp);                              // :315
                double Ssum = 0.0;
                ecdf_eval_all<S>(a.ecdf, s_top, rp, up);                // :316
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const double t = (a.pop.u[j * ld + gi] - up[j]) / eps[j];
                    Ssum = (j == 0) ? t : Ssum + t;
                }
                const double Lacc = ((lpp - a.pop.lp[gi]) + Ssum) + lf; // :318-319
                acc = log_u_less(u53(cw.D), Lacc);                      // :324
            }
            if (acc) {                                                  // :325-328
#pragma unroll
                for (int c = 0; c < D; ++c) a.pop.theta[c * ld + gi] = thp[c];
#pragma unroll
                for (int j = 0; j < S; ++j) { a.pop.u[j * ld + gi] = up[j]; a.pop.rho[j * ld + gi] = rp[j]; }
                a.pop.lp[gi] = lpp;
            } else {
#pragma unroll
                for (int j = 0; j < S; ++j) { up[j] = a.pop.u[j * ld + gi]; rp[j] = a.pop.rho[j * ld + gi]; }
            }
        }
        // epilogue: exact Σu limbs, accept count, per-group ρ tree sums
#pragma unroll
        for (int j = 0; j < S; ++j) {
            uint32_t hi = 0, lo = 0;
            if (valid) u_limbs(up[j], hi, lo);
            const unsigned long long sh = warp_sum_u32(hi), sl = warp_sum_u32(lo);
            if ((tid & 31) == 0) { atomicAdd(&s_acc[2 * j], sh); atomicAdd(&s_acc[2 * j + 1], sl); }
            const double w = warp_tree(valid ? rp[j] : 0.0);
            if ((tid & 31) == 0) s_w[j][tid >> 5] = w;
        }
        const uint32_t na = __reduce_add_sync(0xffffffffu, acc ? 1u : 0u);
        if ((tid & 31) == 0 && na) atomicAdd(&s_acc[2 * S], (unsigned long long)na);
        __syncthreads();
        if (tid < S) {
            double tot = s_w[tid][0];
#pragma unroll
            for (int k = 1; k < 8; ++k) tot = tot + s_w[tid][k];
            a.rho_part[tid * a.part_ld + grp] = tot;
        }
        __syncthreads();
    }
    if (tid < S) { atomicAdd(&a.ds->u_hi[tid], s_acc[2 * tid]); atomicAdd(&a.ds->u_lo[tid], s_acc[2 * tid + 1]); }
    if (tid == 0 && s_acc[2 * S]) atomicAdd(&a.ds->n_acc_iter, s_acc[2 * S]);
}

// ------------------------------------------------------------------------------------------------
// Split form of K4 for models whose simulation dominates (SIR, logistic): with DE/Stretch moves and bounded priors a
// large share of the proposals fails the prior check (:314) and runs no simulation; inside the fused kernel those
// lanes idle for the whole simulation of their warp.  propose_kernel therefore compacts the surviving proposals
// into a work list and simulate_accept_kernel runs one simulation per lane over that list (warps fetch 32 items
// at a time from a device-side cursor).  The Σu / Σρ / accept statistics are then taken by stats_kernel.
// Results are identical to the fused kernel: every particle uses the same Philox streams and arithmetic.
// ------------------------------------------------------------------------------------------------
template <class M, int PROP>
__global__ void __launch_bounds__(CHUNK) propose_kernel(const UpdateArgs a, const SplitScratch w) {
    constexpr int D = M::D;
    __shared__ double s_chol[PROP == PROP_RW ? D * D : 1];
    __shared__ ZigEntry s_zig[PROP == PROP_STRETCH ? 1 : 256];
    const int tid = threadIdx.x, lane = tid & 31;
    if (halted(a.ds)) return;
    if (PROP == PROP_RW) for (int k = tid; k < D * D; k += CHUNK) s_chol[k] = a.ds->chol[k];
    const uint32_t zig = PROP == PROP_STRETCH ? 0u : stage_zig(s_zig);
    __syncthreads();
    const uint64_t sweep = 2ull * (uint64_t)a.ds->t + (uint64_t)a.half;
    const int64_t ld = a.pop.ld;
    const int64_t n_groups = (a.act_n + CHUNK - 1) / CHUNK;
    const InactiveGather<D> P{a.pop.theta + a.ina_off, ld};
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t il = grp * CHUNK + tid;
        bool ok = false;
        double thp[D], lf = 0.0, lpp = 0.0;
        if (il < a.act_n) {
            const int64_t gi = a.act_off + il;
            const uint32_t pid = a.particle_base + (uint32_t)gi;
            double th[D];
#pragma unroll
            for (int c = 0; c < D; ++c) th[c] = a.pop.theta[c * ld + gi];
            const CtrlWords cw = ctrl_words(a.seed, pid, sweep);
            if (PROP == PROP_DE) propose_de<D>(th, P, a.ina_n, a.prop0, a.prop1, cw, Stream(a.seed, pid, sweep, KIND_CTRL, nullptr, zig), thp, lf);
            else if (PROP == PROP_STRETCH) propose_stretch<D>(th, P, a.ina_n, a.prop0, cw, thp, lf);
            else propose_rw<D>(th, s_chol, a.seed, pid, sweep, thp, lf, zig);
            lpp = prior_logpdf<D>(a.prior, thp);
            ok = lpp > -dinf();                                         // :314
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        unsigned base = 0;
        if (lane == 0 && mask) base = atomicAdd(&w.count[a.slot], (unsigned)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (ok) {
            const int64_t q = base + __popc(mask & ((1u << lane) - 1u));
#pragma unroll
            for (int c = 0; c < D; ++c) w.theta[c * w.cap + q] = thp[c];
            w.lp[q] = lpp; w.lf[q] = lf; w.idx[q] = (uint32_t)il;
            if constexpr (M::KEY_BITS > 0) {
                if (w.key) {
                    const uint32_t key = M::work_key(thp, a.mp) & (WORK_BUCKETS - 1);
                    w.key[q] = key;
                    atomicAdd(&w.hist[key], 1u);
                }
            }
        }
    }
}

template <class M>
__global__ void __launch_bounds__(CHUNK, M::SIM_MIN_BLOCKS) simulate_accept_kernel(const __grid_constant__ UpdateArgs a, const SplitScratch w) {
    constexpr int D = M::D, S = M::S;
    extern __shared__ __align__(128) double s_top[];
    __shared__ ZigEntry s_zig[model_draws_normals<M>::value ? 256 : 1];
    if (halted(a.ds)) return;
    const uint32_t zig = model_draws_normals<M>::value ? stage_zig(s_zig) : 0u;
    stage_ecdf_top(a.ecdf, S, s_top);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t sweep = 2ull * (uint64_t)a.ds->t + (uint64_t)a.half;
    double eps[S];
#pragma unroll
    for (int j = 0; j < S; ++j) eps[j] = a.ds->eps[a.n_eps == 1 ? 0 : j];
    const int64_t ld = a.pop.ld;
    const unsigned n_items = w.count[a.slot];
    unsigned n_acc = 0;
    for (;;) {
        unsigned q0 = 0;
        if (lane == 0) q0 = atomicAdd(&w.cursor[a.slot], 32u);
        q0 = __shfl_sync(0xffffffffu, q0, 0);
        if (q0 >= n_items) break;
        const unsigned pos = q0 + lane;
        const unsigned live = __ballot_sync(0xffffffffu, pos < n_items);
        if (pos < n_items) {
            const unsigned q = (M::KEY_BITS > 0 && w.perm) ? w.perm[pos] : pos;
            const int64_t gi = a.act_off + (int64_t)w.idx[q];
            const uint32_t pid = a.particle_base + (uint32_t)gi;
            double thp[D], rp[S], up[S];
#pragma unroll
            for (int c = 0; c < D; ++c) thp[c] = w.theta[c * w.cap + q];
            Stream st(a.seed, pid, sweep, KIND_MODEL, a.rk.k, zig);
            st.warp_mask = live;
            M::sim(thp, a.mp, st, rp);                                  // :315
            double Ssum = 0.0;
            ecdf_eval_all<S>(a.ecdf, s_top, rp, up);                    // :316
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const double t = (a.pop.u[j * ld + gi] - up[j]) / eps[j];
                Ssum = (j == 0) ? t : Ssum + t;
            }
            const double lpp = w.lp[q];
            const double Lacc = ((lpp - a.pop.lp[gi]) + Ssum) + w.lf[q]; // :318-319
            const uint64_t wD = Stream(a.seed, pid, sweep, KIND_CTRL).block(1).b;   // accept uniform: word D of the control stream
            if (log_u_less(u53(wD), Lacc)) {                            // :324-328
#pragma unroll
                for (int c = 0; c < D; ++c) a.pop.theta[c * ld + gi] = thp[c];
#pragma unroll
                for (int j = 0; j < S; ++j) { a.pop.u[j * ld + gi] = up[j]; a.pop.rho[j * ld + gi] = rp[j]; }
                a.pop.lp[gi] = lpp;
                n_acc++;
            }
        }
    }
    n_acc = __reduce_add_sync(0xffffffffu, n_acc);
    if (lane == 0 && n_acc) atomicAdd(&a.ds->n_acc_iter, (unsigned long long)n_acc);
}

// bucket offsets (exclusive scan of the histogram) by one CTA; the histogram is cleared for its second use as cursors
static __global__ void __launch_bounds__(1024) bucket_scan_kernel(unsigned int* hist, unsigned int* off) {
    __shared__ unsigned int s_warp[32];
    __shared__ unsigned int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < WORK_BUCKETS; base += 1024) {
        const unsigned int v = hist[base + threadIdx.x];
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned int x = s_warp[lane], xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
            s_warp[lane] = xi - x;
        }
        __syncthreads();
        const unsigned int excl = s_carry + s_warp[wid] + inc - v;
        off[base + threadIdx.x] = excl;
        hist[base + threadIdx.x] = 0u;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}
// simulation order: entries grouped by bucket (the order inside a bucket is irrelevant to the result)
static __global__ void bucket_scatter_kernel(const uint32_t* key, const unsigned int* count, int slot, unsigned int* cursor,
                                             const unsigned int* off, uint32_t* perm) {
    const unsigned int n = count[slot];
    for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t k = key[q];
        perm[off[k] + atomicAdd(&cursor[k], 1u)] = q;
    }
}
static __global__ void bucket_clear_kernel(unsigned int* hist) {
    for (int i = threadIdx.x; i < WORK_BUCKETS; i += blockDim.x) hist[i] = 0u;
}

// Σu limbs and per-group ρ tree sums of one half after the split update (generic in S)
static __global__ void __launch_bounds__(CHUNK) stats_kernel(PopView pop, int64_t off, int64_t n, int S, DevState* ds, double* rho_part,
                                                      int64_t part_ld) {
    __shared__ unsigned long long s_acc[2 * MAX_S];
    __shared__ double s_w[8];
    const int tid = threadIdx.x;
    if (halted(ds)) return;
    for (int k = tid; k < 2 * S; k += CHUNK) s_acc[k] = 0ull;
    __syncthreads();
    const int64_t n_groups = (n + CHUNK - 1) / CHUNK;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t il = grp * CHUNK + tid;
        const bool valid = il < n;
        for (int j = 0; j < S; ++j) {
            uint32_t hi = 0, lo = 0;
            double r = 0.0;
            if (valid) { u_limbs(pop.u[j * pop.ld + off + il], hi, lo); r = pop.rho[j * pop.ld + off + il]; }
            const unsigned long long sh = warp_sum_u32(hi), sl = warp_sum_u32(lo);
            if ((tid & 31) == 0) { atomicAdd(&s_acc[2 * j], sh); atomicAdd(&s_acc[2 * j + 1], sl); }
            const double g = group256(r, s_w);
            if (tid == 0) rho_part[j * part_ld + grp] = g;
        }
    }
    __syncthreads();
    if (tid < S) { atomicAdd(&ds->u_hi[tid], s_acc[2 * tid]); atomicAdd(&ds->u_lo[tid], s_acc[2 * tid + 1]); }
}

// ------------------------------------------------------------------------------------------------
// K1 init_prior_sim: θ ~ prior, ρ = f_dist(θ)  (src/SimulatedAnnealingABC.jl:172-179), negative
// distance check (:185) and per-group ρ sums for ρ_history[1] (:180).
// ------------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(CHUNK) init_prior_sim_kernel(const InitArgs a) {
    constexpr int D = M::D, S = M::S;
    __shared__ double s_w[S][8];
    const int tid = threadIdx.x;
    const int64_t ld = a.pop.ld;
    const int64_t n_groups = (a.n + CHUNK - 1) / CHUNK;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t i = grp * CHUNK + tid;
        const bool valid = i < a.n;
        double rho[S];
#pragma unroll
        for (int j = 0; j < S; ++j) rho[j] = 0.0;
        if (valid) {
            const uint32_t pid = a.particle_base + (uint32_t)i;
            double th[D];
            prior_rand<D>(a.prior, a.seed, pid, th);
            Stream st(a.seed, pid, 0, KIND_MODEL);
            M::sim(th, a.mp, st, rho);
            bool neg = false;
#pragma unroll
            for (int c = 0; c < D; ++c) a.pop.theta[c * ld + i] = th[c];
#pragma unroll
            for (int j = 0; j < S; ++j) { a.pop.rho[j * ld + i] = rho[j]; neg |= rho[j] < 0.0; }
            a.pop.lp[i] = prior_logpdf<D>(a.prior, th);
            if (neg) atomicOr(&a.ds->error_flag, 1);
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const double w = warp_tree(rho[j]);
            if ((tid & 31) == 0) s_w[j][tid >> 5] = w;
        }
        __syncthreads();
        if (tid < S) {
            double tot = s_w[tid][0];
#pragma unroll
            for (int k = 1; k < 8; ++k) tot = tot + s_w[tid][k];
            a.rho_part[tid * a.part_ld + grp] = tot;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Model plug-in registry: one launcher table per model (the template instantiations above).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t SABC_MODEL_VTABLE_VERSION = 2;
struct ModelVTable {
    uint32_t struct_size, abi_version;   // sizeof(ModelVTable), SABC_MODEL_VTABLE_VERSION of the headers the plug-in was compiled with
    const char* name;
    int32_t n_para, n_stats;
    cudaError_t (*launch_init)(const InitArgs&, int grid, cudaStream_t);
    cudaError_t (*launch_update)(int proposal, const UpdateArgs&, int grid, size_t smem, cudaStream_t);
    cudaError_t (*update_occupancy)(int proposal, size_t smem, int* blocks_per_sm);
    int32_t heavy;            // 1: use the split path (propose -> compacted simulate+accept -> stats)
    int32_t key_bits;         // > 0: the model supplies a similarity key for work-list bucketing
    cudaError_t (*launch_propose)(int proposal, const UpdateArgs&, const SplitScratch&, int grid, cudaStream_t);
    cudaError_t (*launch_simacc)(const UpdateArgs&, const SplitScratch&, int grid, size_t smem, cudaStream_t);
    cudaError_t (*simacc_occupancy)(size_t smem, int* blocks_per_sm);
    cudaError_t (*simulate)(const double* d_theta, int64_t n, int64_t ld, const ModelPar&, uint64_t seed,
                            uint32_t particle_base, uint64_t sweep, double* d_rho, cudaStream_t);
};

template <class M>
__global__ void simulate_kernel(const double* theta, int64_t n, int64_t ld, const ModelPar mp, uint64_t seed,
                                uint32_t particle_base, uint64_t sweep, double* rho) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double th[M::D], r[M::S];
    for (int c = 0; c < M::D; ++c) th[c] = theta[c * ld + i];
    Stream st(seed, particle_base + (uint32_t)i, sweep, KIND_MODEL);
    M::sim(th, mp, st, r);
    for (int j = 0; j < M::S; ++j) rho[j * ld + i] = r[j];
}

template <class M>
struct ModelLaunchers {
    static cudaError_t init(const InitArgs& a, int grid, cudaStream_t s) {
        init_prior_sim_kernel<M><<<grid, CHUNK, 0, s>>>(a);
        return cudaGetLastError();
    }
    template <int PROP>
    static cudaError_t upd(const UpdateArgs& a, int grid, size_t smem, cudaStream_t s) {
        if (smem > 40 * 1024)
            cudaFuncSetAttribute(update_half_kernel<M, PROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        update_half_kernel<M, PROP><<<grid, CHUNK, smem, s>>>(a);
        return cudaGetLastError();
    }
    static cudaError_t update(int proposal, const UpdateArgs& a, int grid, size_t smem, cudaStream_t s) {
        switch (proposal) {
            case PROP_DE: return upd<PROP_DE>(a, grid, smem, s);
            case PROP_STRETCH: return upd<PROP_STRETCH>(a, grid, smem, s);
            case PROP_RW: return upd<PROP_RW>(a, grid, smem, s);
            default: return cudaErrorInvalidValue;
        }
    }
    static cudaError_t occupancy(int proposal, size_t smem, int* b) {
        if (smem > 40 * 1024) {      // the opt-in must precede the query, which otherwise reports 0 resident CTAs
            cudaFuncSetAttribute(update_half_kernel<M, PROP_DE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(update_half_kernel<M, PROP_STRETCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(update_half_kernel<M, PROP_RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        switch (proposal) {
            case PROP_DE: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(b, update_half_kernel<M, PROP_DE>, CHUNK, smem);
            case PROP_STRETCH: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(b, update_half_kernel<M, PROP_STRETCH>, CHUNK, smem);
            case PROP_RW: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(b, update_half_kernel<M, PROP_RW>, CHUNK, smem);
            default: return cudaErrorInvalidValue;
        }
    }
    static cudaError_t propose(int proposal, const UpdateArgs& a, const SplitScratch& w, int grid, cudaStream_t s) {
        switch (proposal) {
            case PROP_DE: propose_kernel<M, PROP_DE><<<grid, CHUNK, 0, s>>>(a, w); break;
            case PROP_STRETCH: propose_kernel<M, PROP_STRETCH><<<grid, CHUNK, 0, s>>>(a, w); break;
            case PROP_RW: propose_kernel<M, PROP_RW><<<grid, CHUNK, 0, s>>>(a, w); break;
            default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    static cudaError_t simacc(const UpdateArgs& a, const SplitScratch& w, int grid, size_t smem, cudaStream_t s) {
        if (smem > 40 * 1024)
            cudaFuncSetAttribute(simulate_accept_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        simulate_accept_kernel<M><<<grid, CHUNK, smem, s>>>(a, w);
        return cudaGetLastError();
    }
    static cudaError_t simacc_occ(size_t smem, int* b) {
        if (smem > 40 * 1024)
            cudaFuncSetAttribute(simulate_accept_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(b, simulate_accept_kernel<M>, CHUNK, smem);
    }
    static cudaError_t simulate(const double* th, int64_t n, int64_t ld, const ModelPar& mp, uint64_t seed,
                                uint32_t pb, uint64_t sweep, double* rho, cudaStream_t s) {
        simulate_kernel<M><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(th, n, ld, mp, seed, pb, sweep, rho);
        return cudaGetLastError();
    }
    static ModelVTable vtable(const char* name, int heavy = 0) {
        ModelVTable v{};
        v.struct_size = (uint32_t)sizeof(ModelVTable); v.abi_version = SABC_MODEL_VTABLE_VERSION;
        v.name = name; v.n_para = M::D; v.n_stats = M::S;
        v.launch_init = &init; v.launch_update = &update; v.update_occupancy = &occupancy; v.simulate = &simulate;
        v.heavy = heavy; v.key_bits = M::KEY_BITS; v.launch_propose = &propose; v.launch_simacc = &simacc; v.simacc_occupancy = &simacc_occ;
        return v;
    }
};
#endif  // __CUDACC__

}  // namespace sabc
