// aux_kernels.cuh -- the O(N) kernels around the fused update: ECDF build helpers and transform,
// resampling (weights, scan, draws, gather), per-iteration finalisation (ρ tree sums, resampling
// trigger, ε solve, history) and the RandomWalk covariance.  Generic in D and S (runtime).
//
// Reference steps: build_cdf src/cdf_estimators.jl:23-44; transform src/SimulatedAnnealingABC.jl:190-192;
// resample_population :124-137; trigger :340-343; ε :350-354; history :367-372; update_proposal!
// src/proposals.jl:46-48,58-60.
#pragma once
#include "kernels.cuh"

namespace sabc {

constexpr int TILE = 2048;   // items per scan tile = 256 threads x 8

// ---- ECDF build ----
static __global__ void k_mark_positive(const double* x, int64_t n, double* keys, unsigned long long* n_pos) {
    unsigned int cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        const bool pos = v > 0.0;                       // filter(e -> e > 0, x): drops zeros, negatives, NaN
        keys[i] = pos ? v : dinf();
        cnt += pos ? 1u : 0u;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_pos, (unsigned long long)cnt);
}
static __global__ void k_ecdf_ends(double* knots, int64_t n_pos) {   // values = [0; sort(x); maximum(x)*1.5]
    knots[0] = 0.0;
    knots[n_pos + 1] = knots[n_pos] * 1.5;
}
// one index level (plugin.cuh, "ECDF"): from the natural array `in` (cnt entries) the probe-ordered nodes of the blocks below
// every 9th entry, and the next natural level (those 9th entries), padded with +inf up to next_len
static __global__ void k_ecdf_level(const double* in, int64_t cnt, int64_t n_nodes, double* nodes, double* next, int64_t next_len) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < n_nodes * ECDF_NODE; i += stride) {
        const int s = (int)(i & 7);
        const int within = s == 0 ? 3 : s == 1 ? 6 : s == 2 ? 1 : s == 3 ? 2 : s == 4 ? 4 : s == 5 ? 5 : s == 6 ? 7 : 8;   // [e3 e6 | e1 e2 | e4 e5 | e7 e8]
        const int64_t idx = (i >> 3) * ECDF_STRIDE + within;
        nodes[i] = idx < cnt ? in[idx] : dinf();
    }
    for (int64_t i = t0; i < next_len; i += stride) next[i] = i < n_nodes ? in[i * ECDF_STRIDE] : dinf();
}
// the staged top level: `len` entries (+inf beyond cnt) stored as len high words followed by len low words
static __global__ void k_top_split(const double* in, int64_t cnt, double* out, int64_t len) {
    uint32_t* hi = reinterpret_cast<uint32_t*>(out);
    uint32_t* lo = hi + len;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = f64_bits(i < cnt ? in[i] : dinf());
        hi[i] = (uint32_t)(b >> 32); lo[i] = (uint32_t)b;
    }
}
// compressed ECDF: K rank-uniform quantiles x[floor(i (m-1)/(K-1))], i = 0..K-1, of the m sorted positive distances
static __global__ void k_ecdf_subsample(const double* sorted_pos, int64_t m, int K, double* out /* K + 2 + ECDF_PAD */) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K + 2 + ECDF_PAD; i += gridDim.x * blockDim.x) {
        double v;
        if (i == 0) v = 0.0;
        else if (i <= K) v = sorted_pos[((int64_t)(i - 1) * (m - 1)) / (int64_t)(K - 1)];
        else if (i == K + 1) v = sorted_pos[m - 1] * 1.5;
        else v = dinf();
        out[i] = v;
    }
}
static __global__ void k_fill_inf(double* p, int n) {
    if (threadIdx.x < n) p[threadIdx.x] = dinf();
}

// u = G(ρ) for every particle and statistic + exact Σu limbs
static __global__ void __launch_bounds__(CHUNK) k_transform(PopView pop, int64_t n, int S, const EcdfStat* ecdf, DevState* ds) {
    extern __shared__ __align__(128) double s_top[];
    __shared__ unsigned long long s_acc[2 * MAX_S];
    for (int k = threadIdx.x; k < 2 * S; k += CHUNK) s_acc[k] = 0ull;
    stage_ecdf_top(ecdf, S, s_top);
    const int64_t n_groups = (n + CHUNK - 1) / CHUNK;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t i = grp * CHUNK + threadIdx.x;
        for (int j = 0; j < S; ++j) {
            uint32_t hi = 0, lo = 0;
            if (i < n) {
                const double u = ecdf_eval(ecdf[j], s_top, pop.rho[j * pop.ld + i]);
                pop.u[j * pop.ld + i] = u;
                u_limbs(u, hi, lo);
            }
            const unsigned long long sh = warp_sum_u32(hi), sl = warp_sum_u32(lo);
            if ((threadIdx.x & 31) == 0) { atomicAdd(&s_acc[2 * j], sh); atomicAdd(&s_acc[2 * j + 1], sl); }
        }
    }
    __syncthreads();
    if (threadIdx.x < S) {
        atomicAdd(&ds->u_hi[threadIdx.x], s_acc[2 * threadIdx.x]);
        atomicAdd(&ds->u_lo[threadIdx.x], s_acc[2 * threadIdx.x + 1]);
    }
}

// standalone transform used by the parity hook (one statistic, u only)
static __global__ void __launch_bounds__(CHUNK) k_transform1(const double* rho, int64_t m, const EcdfStat* ecdf, double* u) {
    extern __shared__ __align__(128) double s_top[];
    stage_ecdf_top(ecdf, 1, s_top);
    for (int64_t i = (int64_t)blockIdx.x * CHUNK + threadIdx.x; i < m; i += (int64_t)gridDim.x * CHUNK)
        u[i] = ecdf_eval(ecdf[0], s_top, rho[i]);
}

// ---- column tree sums: block b reduces column b of part (leading dim part_ld) ----
static __global__ void __launch_bounds__(CHUNK) k_treesum_cols(const double* part, int64_t part_ld, int64_t n, double* scratch,
                                                        int64_t scratch_ld, double* out) {
    __shared__ double s_w[8];
    const double r = cta_treesum(part + blockIdx.x * part_ld, n, scratch + blockIdx.x * scratch_ld, s_w);
    if (threadIdx.x == 0) out[blockIdx.x] = r;
}
// level-1 group sums of a raw column-major array: grid-stride over groups, block handles all columns
static __global__ void __launch_bounds__(CHUNK) k_group_sums(const double* x, int64_t ld, int64_t n, int ncols, double* part,
                                                      int64_t part_ld) {
    __shared__ double s_w[8];
    const int64_t n_groups = (n + CHUNK - 1) / CHUNK;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t i = grp * CHUNK + threadIdx.x;
        for (int c = 0; c < ncols; ++c) {
            const double r = group256(i < n ? x[c * ld + i] : 0.0, s_w);
            if (threadIdx.x == 0) part[c * part_ld + grp] = r;
        }
    }
}

// ---- per-iteration finalisation, first part: ρ tree sums (blocks 0..2S-1) and the resampling trigger ----
struct Post1Args {
    DevState* ds;
    const double* rho_part;      // [2][S][part_ld]
    int64_t part_ld, groups0, groups1;
    double* scratch; int64_t scratch_ld;
    int S; int64_t n_global; int64_t resample; int decide;
};
// accept counter, resampling trigger and the column means the weights need
SABC_D void post1_decide(DevState* ds, int S, int64_t n_global, int64_t resample) {
    ds->n_accept += (long long)ds->n_acc_iter;                                      // :334
    ds->resample_flag = ds->n_accept >= (ds->n_resampling + 1) * resample ? 1 : 0;  // :340
    for (int j = 0; j < S; ++j) {
        ds->ubar[j] = limbs_to_sum(ds->u_hi[j], ds->u_lo[j]) / (double)n_global;    // :126 mean(u, dims=1)
        ds->r_hi[j] = 0ull; ds->r_lo[j] = 0ull;
    }
    ds->w_total = 0ull;
}
SABC_D void d_post1_block(const Post1Args& a, int b, double* s_w) {
    if (b < 2 * a.S) {
        const int half = b / a.S, j = b % a.S;
        const double r = cta_treesum(a.rho_part + ((int64_t)half * a.S + j) * a.part_ld, half == 0 ? a.groups0 : a.groups1,
                                     a.scratch + (int64_t)b * a.scratch_ld, s_w);
        if (threadIdx.x == 0) a.ds->rho_sum[half][j] = r;
    } else if (a.decide && threadIdx.x == 0) {
        post1_decide(a.ds, a.S, a.n_global, a.resample);
    }
}
static __global__ void __launch_bounds__(CHUNK) k_post1(const Post1Args a) {
    __shared__ double s_w[8];
    if (halted(a.ds)) return;
    d_post1_block(a, blockIdx.x, s_w);
}
static __global__ void k_decide(DevState* ds, int S, int64_t n_global, int64_t resample) {
    if (threadIdx.x == 0) post1_decide(ds, S, n_global, resample);
}
static __global__ void k_force_flag(DevState* ds, int flag) { ds->resample_flag = flag; }
static __global__ void k_set_counters(DevState* ds, long long n_accept, long long n_resampling) {
    ds->n_accept = n_accept; ds->n_resampling = n_resampling;
}
static __global__ void k_recompute_lp(PopView pop, int64_t r0, int64_t r1, int D, const PriorSpec prior) {
    for (int64_t i = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += (int64_t)gridDim.x * blockDim.x) {
        double lp = 0.0;
        for (int c = 0; c < D; ++c) {
            const double t = prior_logpdf1(prior.kind[c], prior.p0[c], prior.p1[c], prior.c[c], pop.theta[c * pop.ld + i]);
            lp = (c == 0) ? t : lp + t;
        }
        pop.lp[i] = lp;
    }
}

// ---- resampling ----
SABC_D unsigned long long cta_sum_u64(unsigned long long v, unsigned long long* s8) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) s8[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long tot = 0;
    if (threadIdx.x == 0) for (int k = 0; k < 8; ++k) tot += s8[k];
    __syncthreads();
    return tot;   // valid in thread 0
}

// w_i = exp(-Σ_j u_ij δ / ū_j) as 32.32 fixed point, plus per-tile sums  (:127)
SABC_D void d_weights(const PopView& pop, int64_t n, int S, double delta, const DevState* ds, unsigned long long* q,
                      unsigned long long* tile_sum, int vb, int vg, unsigned long long* s8, double* s_ubar) {
    if (threadIdx.x < S) s_ubar[threadIdx.x] = ds->ubar[threadIdx.x];
    __syncthreads();
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = vb; tile < n_tiles; tile += vg) {
        const int64_t base = tile * TILE + (int64_t)threadIdx.x * 8;
        unsigned long long local = 0;
        for (int k = 0; k < 8; ++k) {
            const int64_t i = base + k;
            if (i < n) {
                double acc = 0.0;
                for (int j = 0; j < S; ++j) {
                    const double t = (pop.u[j * pop.ld + i] * delta) / s_ubar[j];
                    acc = (j == 0) ? t : acc + t;
                }
                const unsigned long long qi = (unsigned long long)(det_exp(-acc) * 4294967296.0);
                q[i] = qi; local += qi;
            }
        }
        const unsigned long long tot = cta_sum_u64(local, s8);
        if (threadIdx.x == 0) tile_sum[tile] = tot;
    }
}
static __global__ void __launch_bounds__(CHUNK) k_weights(PopView pop, int64_t n, int S, double delta, const DevState* ds,
                                                   unsigned long long* q, unsigned long long* tile_sum, int force) {
    if (!force && !ds->resample_flag) return;
    __shared__ unsigned long long s8[8];
    __shared__ double s_ubar[MAX_S];
    d_weights(pop, n, S, delta, ds, q, tile_sum, blockIdx.x, gridDim.x, s8, s_ubar);
}
// standalone per-tile sums of an existing integer array (multi-GPU selection flags, hooks)
static __global__ void __launch_bounds__(CHUNK) k_tile_sums(const unsigned long long* q, int64_t n, unsigned long long* tile_sum) {
    __shared__ unsigned long long s8[8];
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE + (int64_t)threadIdx.x * 8;
        unsigned long long local = 0;
        for (int k = 0; k < 8; ++k) if (base + k < n) local += q[base + k];
        const unsigned long long tot = cta_sum_u64(local, s8);
        if (threadIdx.x == 0) tile_sum[tile] = tot;
    }
}
// exclusive scan of the tile sums by one CTA of 1024 threads; total -> *total_out
static __global__ void __launch_bounds__(1024) k_scan_tiles(const unsigned long long* tile_sum, int64_t n_tiles,
                                                     unsigned long long* tile_off, unsigned long long* total_out,
                                                     const DevState* ds, int force) {
    if (!force && !ds->resample_flag) return;
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_sum[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = s_warp[lane], winc = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, off);
                if (lane >= off) winc += t;
            }
            s_warp[lane] = winc - w;   // exclusive warp offsets
        }
        __syncthreads();
        const unsigned long long excl = s_carry + s_warp[wid] + inc - v;
        if (i < n_tiles) tile_off[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total_out = s_carry;
        // every 32.32 fixed-point weight flushed to zero (delta u / mean u beyond ~22 for all particles): fail like the sharded
        // path does instead of collapsing the population onto one particle
        // (force == 2: a rank of a sharded population, whose slice alone may legitimately weigh nothing)
        if (s_carry == 0ull && total_out == &ds->w_total && force != 2) atomicOr(const_cast<int*>(&ds->error_flag), 4);
    }
}
// inclusive prefix sums in place: q[i] <- tile_off + Σ_{k<=i in tile} q[k]
SABC_D void d_prefix(unsigned long long* q, int64_t n, const unsigned long long* tile_off, int vb, int vg, unsigned long long* s_warp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = vb; tile < n_tiles; tile += vg) {
        const int64_t base = tile * TILE + (int64_t)threadIdx.x * 8;
        unsigned long long v[8], run = 0;
        for (int k = 0; k < 8; ++k) { v[k] = base + k < n ? q[base + k] : 0ull; run += v[k]; v[k] = run; }
        unsigned long long inc = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        unsigned long long woff = 0;
        for (int k = 0; k < wid; ++k) woff += s_warp[k];
        const unsigned long long excl = tile_off[tile] + woff + inc - run;
        for (int k = 0; k < 8; ++k) if (base + k < n) q[base + k] = excl + v[k];
        __syncthreads();
    }
}
static __global__ void __launch_bounds__(CHUNK) k_prefix(unsigned long long* q, int64_t n, const unsigned long long* tile_off,
                                                  const DevState* ds, int force) {
    if (!force && !ds->resample_flag) return;
    __shared__ unsigned long long s_warp[8];
    d_prefix(q, n, tile_off, blockIdx.x, gridDim.x, s_warp);
}

// first index i in [0,n) with P[i] > r
SABC_HD int64_t upper_bound_u64(const unsigned long long* P, int64_t n, unsigned long long r) {
    int64_t lo = 0, len = n;
    while (len > 0) {
        const int64_t half = len >> 1;
        if (P[lo + half] <= r) { lo += half + 1; len -= half + 1; } else { len = half; }
    }
    return lo;
}

// N iid categorical draws by inversion of the exact prefix sums, fused with the gather of
// population[idx] and u[idx,:] (:129-132; ρ is not resampled, :197,341) and the exact Σu of the result.
SABC_D void d_draw_gather(const PopView& pop, const PopView& tmp, int64_t n, int D, int S, const unsigned long long* P, uint64_t seed,
                          DevState* ds, int vb, int vg, unsigned long long* s_acc) {
    for (int k = threadIdx.x; k < 2 * S; k += CHUNK) s_acc[k] = 0ull;
    __syncthreads();
    const unsigned long long W = ds->w_total;
    const uint32_t rc = (uint32_t)ds->n_resampling;
    const int64_t n_groups = (n + CHUNK - 1) / CHUNK;
    for (int64_t grp = vb; grp < n_groups; grp += vg) {
        const int64_t k = grp * CHUNK + threadIdx.x;
        int64_t src = 0;
        if (k < n) {
            const U64x2 w = philox4x32_10((uint32_t)k, rc, (uint32_t)((uint64_t)k >> 32), KIND_RESAMPLE,
                                          (uint32_t)seed, (uint32_t)(seed >> 32));
            src = upper_bound_u64(P, n, mulhi64(w.a, W));
            if (src >= n) src = n - 1;
            for (int c = 0; c < D; ++c) tmp.theta[c * tmp.ld + k] = pop.theta[c * pop.ld + src];
            tmp.lp[k] = pop.lp[src];
        }
        for (int j = 0; j < S; ++j) {
            uint32_t hi = 0, lo = 0;
            if (k < n) {
                const double u = pop.u[j * pop.ld + src];
                tmp.u[j * tmp.ld + k] = u;
                u_limbs(u, hi, lo);
            }
            const unsigned long long sh = warp_sum_u32(hi), sl = warp_sum_u32(lo);
            if ((threadIdx.x & 31) == 0) { atomicAdd(&s_acc[2 * j], sh); atomicAdd(&s_acc[2 * j + 1], sl); }
        }
    }
    __syncthreads();
    if (threadIdx.x < S) {
        atomicAdd(&ds->r_hi[threadIdx.x], s_acc[2 * threadIdx.x]);
        atomicAdd(&ds->r_lo[threadIdx.x], s_acc[2 * threadIdx.x + 1]);
    }
}
static __global__ void __launch_bounds__(CHUNK) k_draw_gather(PopView pop, PopView tmp, int64_t n, int D, int S,
                                                       const unsigned long long* P, uint64_t seed, DevState* ds, int force) {
    if ((!force && !ds->resample_flag) || ds->error_flag) return;
    __shared__ unsigned long long s_acc[2 * MAX_S];
    d_draw_gather(pop, tmp, n, D, S, P, seed, ds, blockIdx.x, gridDim.x, s_acc);
}
SABC_D void d_copyback(const PopView& pop, const PopView& tmp, int64_t n, int D, int S, int64_t first, int64_t stride) {
    for (int64_t i = first; i < n; i += stride) {
        for (int c = 0; c < D; ++c) pop.theta[c * pop.ld + i] = tmp.theta[c * tmp.ld + i];
        for (int j = 0; j < S; ++j) pop.u[j * pop.ld + i] = tmp.u[j * tmp.ld + i];
        pop.lp[i] = tmp.lp[i];
    }
}
static __global__ void k_copyback(PopView pop, PopView tmp, int64_t n, int D, int S, const DevState* ds, int force) {
    if ((!force && !ds->resample_flag) || ds->error_flag) return;
    d_copyback(pop, tmp, n, D, S, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}
// indices only (parity hook)
static __global__ void k_draw_indices(const unsigned long long* P, int64_t n, unsigned long long W, uint64_t seed, uint32_t rc,
                               int64_t* idx) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const U64x2 w = philox4x32_10((uint32_t)k, rc, (uint32_t)((uint64_t)k >> 32), KIND_RESAMPLE, (uint32_t)seed,
                                      (uint32_t)(seed >> 32));
        int64_t src = upper_bound_u64(P, n, mulhi64(w.a, W));
        idx[k] = src < n ? src : n - 1;
    }
}

// ---- per-iteration finalisation, second part: ε update, history record, counters ----
struct FinishArgs {
    DevState* ds;
    double* hist;            // records of (n_eps + 2 S) doubles: ε, mean u, mean ρ
    int S, n_eps, algorithm;
    int64_t n_global;
    double v;
};
SABC_D void d_finish(const FinishArgs& a, double* s_um, unsigned long long* s_hi, unsigned long long* s_lo) {
    DevState* ds = a.ds;
    const int tid = threadIdx.x;
    const int flag = ds->resample_flag;
    if (tid < a.S) {
        s_hi[tid] = flag ? ds->r_hi[tid] : ds->u_hi[tid];
        s_lo[tid] = flag ? ds->r_lo[tid] : ds->u_lo[tid];
        s_um[tid] = limbs_to_sum(s_hi[tid], s_lo[tid]) / (double)a.n_global;
    }
    __syncthreads();
    if (a.algorithm == ALG_MULTI_EPS) {                                        // :350-351
        if (tid < a.S) {
            double e = 0.0;
            if (!eps_multi_one(s_um, a.S, tid, a.v, e)) atomicOr(&ds->error_flag, 2);
            ds->eps[tid] = e;
        }
    } else if (tid == 0) {                                                     // :352-353 mean(u) over all N*s entries
        unsigned long long gh = 0, gl = 0;
        for (int j = 0; j < a.S; ++j) { gh += s_hi[j]; gl += s_lo[j]; }
        ds->eps[0] = eps_single(limbs_to_sum(gh, gl) / (double)(a.n_global * a.S), a.v);
    }
    __syncthreads();
    if (tid == 0) {
        if (flag) ds->n_resampling += 1;                                       // :342
        if (ds->ix % ds->checkpoint == 0 || ds->ix == ds->n_pop) {            // :367-372, :378-382
            double* rec = a.hist + ds->rec * (a.n_eps + 2 * a.S);
            for (int k = 0; k < a.n_eps; ++k) rec[k] = ds->eps[k];
            for (int j = 0; j < a.S; ++j) {
                rec[a.n_eps + j] = s_um[j];
                rec[a.n_eps + a.S + j] = (ds->rho_sum[0][j] + ds->rho_sum[1][j]) / (double)a.n_global;
            }
            ds->rec += 1;
            ds->last_cp = ds->ix;
        }
        for (int j = 0; j < a.S; ++j) { ds->u_hi[j] = 0ull; ds->u_lo[j] = 0ull; }
        ds->n_acc_iter = 0ull;
        ds->resample_flag = 0;
        for (int k = 0; k < MAX_SLOTS; ++k) { ds->list_count[k] = 0u; ds->list_cursor[k] = 0u; }
        ds->t += 1; ds->ix += 1;
    }
}
static __global__ void k_finish(const FinishArgs a) {
    __shared__ double s_um[MAX_S];
    __shared__ unsigned long long s_hi[MAX_S], s_lo[MAX_S];
    if (halted(a.ds)) return;
    d_finish(a, s_um, s_hi, s_lo);
}

// ---- small populations: everything after the two half-sweeps in ONE single-CTA kernel (rho tree sums, trigger, resampling,
// eps, history).  For n <= 16384 the nine tiny launches of the generic tail cost more than their work; results are identical.
struct TailArgs {
    Post1Args post;
    FinishArgs fin;
    PopView pop, tmp;
    int64_t n; int D, S;
    double delta; uint64_t seed;
    unsigned long long *q, *tile_sum, *tile_off;
};
static __global__ void __launch_bounds__(CHUNK) k_tail_small(const TailArgs a) {
    __shared__ double s_w[8];
    __shared__ unsigned long long s8[8], s_acc[2 * MAX_S], s_hi[MAX_S], s_lo[MAX_S];
    __shared__ double s_ubar[MAX_S], s_um[MAX_S];
    DevState* ds = a.post.ds;
    for (int b = 0; b <= 2 * a.S; ++b) d_post1_block(a.post, b, s_w);
    __syncthreads();
    if (halted(ds)) return;
    if (ds->resample_flag) {
        d_weights(a.pop, a.n, a.S, a.delta, ds, a.q, a.tile_sum, 0, 1, s8, s_ubar);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long run = 0;
            const int64_t n_tiles = (a.n + TILE - 1) / TILE;
            for (int64_t t = 0; t < n_tiles; ++t) { a.tile_off[t] = run; run += a.tile_sum[t]; }
            ds->w_total = run;
            if (run == 0ull) atomicOr(&ds->error_flag, 4);
        }
        __syncthreads();
        d_prefix(a.q, a.n, a.tile_off, 0, 1, s8);
        __syncthreads();
        if (ds->error_flag) return;                       // all weights zero (uniform across the CTA)
        d_draw_gather(a.pop, a.tmp, a.n, a.D, a.S, a.q, a.seed, ds, 0, 1, s_acc);
        __syncthreads();
        d_copyback(a.pop, a.tmp, a.n, a.D, a.S, threadIdx.x, blockDim.x);
        __syncthreads();
    }
    d_finish(a.fin, s_um, s_hi, s_lo);
}

// ---- RandomWalk covariance (update_proposal!, src/proposals.jl:46-48,58-60) ----
// centred cross products, level-1 group sums: column index p = a(a+1)/2 + b, b <= a
static __global__ void __launch_bounds__(CHUNK) k_rw_cross_sums(const double* theta, int64_t ld, int64_t n, int D, const DevState* ds,
                                                         double* part, int64_t part_ld) {
    __shared__ double s_w[8];
    const int64_t n_groups = (n + CHUNK - 1) / CHUNK;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t i = grp * CHUNK + threadIdx.x;
        int p = 0;
        for (int a = 0; a < D; ++a) for (int b = 0; b <= a; ++b, ++p) {
            double v = 0.0;
            if (i < n) v = (theta[a * ld + i] - ds->mom[a]) * (theta[b * ld + i] - ds->mom[b]);
            const double r = group256(v, s_w);
            if (threadIdx.x == 0) part[p * part_ld + grp] = r;
        }
    }
}
static __global__ void k_rw_means(DevState* ds, const double* sums, int D, int64_t n_global) {
    if (halted(ds)) return;
    if (threadIdx.x < D) ds->mom[threadIdx.x] = sums[threadIdx.x] / (double)n_global;
}
static __global__ void k_rw_chol(DevState* ds, const double* sums, int D, int64_t n_global, double beta) {
    if (threadIdx.x != 0 || halted(ds)) return;
    double Sg[MAX_D * MAX_D];
    int p = 0;
    for (int a = 0; a < D; ++a) for (int b = 0; b <= a; ++b, ++p) {
        const double cov = sums[p] / (double)(n_global - 1);
        if (D == 1) { ds->chol[0] = sqrt(beta * cov); return; }                     // :59, :54
        const double s = beta * (a == b ? cov + 1e-8 : cov);                        // :47
        Sg[a * D + b] = s; Sg[b * D + a] = s;
    }
    for (int r = 0; r < D; ++r) for (int c = 0; c <= r; ++c) {
        double sum = Sg[r * D + c];
        for (int k = 0; k < c; ++k) sum = sum - ds->chol[r * D + k] * ds->chol[c * D + k];
        ds->chol[r * D + c] = (r == c) ? sqrt(sum) : sum / ds->chol[c * D + c];
    }
}

// start of an update() call / of init(): position counters
static __global__ void k_begin(DevState* ds, long long t, long long n_pop, long long checkpoint) {
    ds->t = t; ds->ix = 1; ds->n_pop = n_pop; ds->checkpoint = checkpoint; ds->rec = 0; ds->last_cp = 0;
    for (int j = 0; j < MAX_S; ++j) { ds->u_hi[j] = 0ull; ds->u_lo[j] = 0ull; }
    ds->n_acc_iter = 0ull; ds->resample_flag = 0; ds->hold = 0; ds->error_flag = 0;    // an earlier call's error does not stick
    for (int k = 0; k < MAX_SLOTS; ++k) { ds->list_count[k] = 0u; ds->list_cursor[k] = 0u; }
}
// Σu limbs of an existing u array (set_population, multi-GPU resampling)
static __global__ void __launch_bounds__(CHUNK) k_sum_u(const double* u, int64_t ld, int64_t n, int S, unsigned long long* hi_out,
                                                 unsigned long long* lo_out) {
    __shared__ unsigned long long s_acc[2 * MAX_S];
    for (int k = threadIdx.x; k < 2 * S; k += CHUNK) s_acc[k] = 0ull;
    __syncthreads();
    const int64_t n_groups = (n + CHUNK - 1) / CHUNK;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t i = grp * CHUNK + threadIdx.x;
        for (int j = 0; j < S; ++j) {
            uint32_t hi = 0, lo = 0;
            if (i < n) u_limbs(u[j * ld + i], hi, lo);
            const unsigned long long sh = warp_sum_u32(hi), sl = warp_sum_u32(lo);
            if ((threadIdx.x & 31) == 0) { atomicAdd(&s_acc[2 * j], sh); atomicAdd(&s_acc[2 * j + 1], sl); }
        }
    }
    __syncthreads();
    if (threadIdx.x < S) { atomicAdd(&hi_out[threadIdx.x], s_acc[2 * threadIdx.x]); atomicAdd(&lo_out[threadIdx.x], s_acc[2 * threadIdx.x + 1]); }
}

}  // namespace sabc
