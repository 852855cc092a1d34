// multi_gpu.inl -- ranks of one communicator, contiguous particle slices.  Included by engine.cu.  A rank is a
// process (one per GPU) or one of the per-GPU engines of a single-process handle (group.inl): the code below is the same.
//
// Per population update ONE collective crosses NVLink: an all-gather of every rank's packed statistics (exact Σu limbs, accept
// count, FP64 ρ sums), reduced in rank order on every rank, which then solves ε redundantly.  The update sequence is enqueued
// without host round trips (look-ahead with a device-side `hold` flag, replayed as a CUDA graph).  DE/Stretch partners come from
// the rank's own inactive half.  Resampling is the exact global multinomial of the reference (:129), split by rank: the per-rank
// counts are one multinomial draw from a shared seed, each rank draws its particles locally and ships only the surplus to the
// ranks that own the destination slots (grouped ncclSend/ncclRecv).  The strict variant (all ranks walk the same N global
// variates; multiset equal to one GPU's) and the replicated mode (bit-identical to one GPU) are the parity instruments.

static __global__ void __launch_bounds__(CHUNK) k_mg_mark(int64_t n_global, unsigned long long w_total, unsigned long long my_off,
                                                   unsigned long long my_w, const unsigned long long* P, int64_t n_local,
                                                   uint64_t seed, uint32_t rc, unsigned long long* F, int64_t* src) {
    for (int64_t k = (int64_t)blockIdx.x * CHUNK + threadIdx.x; k < n_global; k += (int64_t)gridDim.x * CHUNK) {
        const U64x2 w = philox4x32_10((uint32_t)k, rc, (uint32_t)((uint64_t)k >> 32), KIND_RESAMPLE, (uint32_t)seed,
                                      (uint32_t)(seed >> 32));
        const unsigned long long r = mulhi64(w.a, w_total);
        const bool mine = r >= my_off && r - my_off < my_w;
        int64_t s = -1;
        if (mine) { s = upper_bound_u64(P, n_local, r - my_off); if (s >= n_local) s = n_local - 1; }
        F[k] = mine ? 1ull : 0ull;
        src[k] = s;
    }
}
// pack the selected particles field-major into the send buffer, in draw order
static __global__ void __launch_bounds__(CHUNK) k_mg_pack(PopView pop, int D, int S, int64_t n_global, const unsigned long long* Fincl,
                                                   const int64_t* src, double* sb, int64_t sb_ld) {
    for (int64_t k = (int64_t)blockIdx.x * CHUNK + threadIdx.x; k < n_global; k += (int64_t)gridDim.x * CHUNK) {
        const int64_t s = src[k];
        if (s < 0) continue;
        const int64_t pos = (int64_t)Fincl[k] - 1;
        for (int c = 0; c < D; ++c) sb[c * sb_ld + pos] = pop.theta[c * pop.ld + s];
        for (int j = 0; j < S; ++j) sb[(D + j) * sb_ld + pos] = pop.u[j * pop.ld + s];
        sb[(D + S) * sb_ld + pos] = pop.lp[s];
    }
}

static int mg_allreduce_f64(sabc_engine* e, double* d, int n) {
    SABC_NCCL(nccl_api()->AllReduce(d, d, (size_t)n, ncclFloat64, ncclSum, e->comm.comm, e->stream));
    return 0;
}
static int mg_allreduce_u64(sabc_engine* e, unsigned long long* d, int n) {
    SABC_NCCL(nccl_api()->AllReduce(d, d, (size_t)n, ncclUint64, ncclSum, e->comm.comm, e->stream));
    return 0;
}
static int mg_allgather_f64(sabc_engine* e, const double* src, double* dst, int64_t n_each) {
    SABC_NCCL(nccl_api()->AllGather(src, dst, (size_t)n_each, ncclFloat64, e->comm.comm, e->stream));
    return 0;
}
static int mg_allgather_u64_host(sabc_engine* e, const unsigned long long* d_one, unsigned long long* h_all) {
    MgScratch& s = e->mg;
    SABC_CUDA(s.wall.ensure((size_t)e->world));
    SABC_NCCL(nccl_api()->AllGather(d_one, s.wall.p, 1, ncclUint64, e->comm.comm, e->stream));
    SABC_CUDA(cudaMemcpyAsync(h_all, s.wall.p, sizeof(unsigned long long) * e->world, cudaMemcpyDeviceToHost, e->stream));
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}
static int mg_any_flag(sabc_engine* e, int* host_flag) {
    MgScratch& s = e->mg;
    SABC_CUDA(s.flag.ensure(1));
    SABC_CUDA(cudaMemcpyAsync(s.flag.p, host_flag, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    SABC_NCCL(nccl_api()->AllReduce(s.flag.p, s.flag.p, 1, ncclInt32, ncclMax, e->comm.comm, e->stream));
    SABC_CUDA(cudaMemcpyAsync(host_flag, s.flag.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}
// Σu limbs + accept count of the running iteration: one integer all-reduce (exact, order-free)
static int mg_reduce_iteration_sums(sabc_engine* e) {
    return mg_allreduce_u64(e, &e->b_ds.p->u_hi[0], 2 * MAX_S + 1);
}

static int launch_update_proposal_mg(sabc_engine* e) {
    if (e->proposal != PROP_RW) return 0;
    const int64_t n = e->n_local, groups = (n + CHUNK - 1) / CHUNK;
    const int grid = (int)std::min<int64_t>(groups, (int64_t)e->n_sm * 8);
    const int npair = e->D * (e->D + 1) / 2;
    DevState* ds = e->b_ds.p;
    k_group_sums<<<grid, CHUNK, 0, e->stream>>>(e->pop.theta, e->pop.ld, n, e->D, e->b_rw_part.p, e->part_ld * 2);
    k_treesum_cols<<<e->D, CHUNK, 0, e->stream>>>(e->b_rw_part.p, e->part_ld * 2, groups, e->b_scratch.p, e->scratch_ld, e->b_rw_sums.p);
    SABC_TRY(mg_allreduce_f64(e, e->b_rw_sums.p, e->D));
    k_rw_means<<<1, 32, 0, e->stream>>>(ds, e->b_rw_sums.p, e->D, e->N);
    k_rw_cross_sums<<<grid, CHUNK, 0, e->stream>>>(e->pop.theta, e->pop.ld, n, e->D, ds, e->b_rw_part.p, e->part_ld * 2);
    k_treesum_cols<<<npair, CHUNK, 0, e->stream>>>(e->b_rw_part.p, e->part_ld * 2, groups, e->b_scratch.p, e->scratch_ld, e->b_rw_sums.p);
    SABC_TRY(mg_allreduce_f64(e, e->b_rw_sums.p, npair));
    k_rw_chol<<<1, 32, 0, e->stream>>>(ds, e->b_rw_sums.p, e->D, e->N, e->prop_par[0]);
    SABC_CUDA(cudaGetLastError());
    return 0;
}

// Establish every send/recv connection once, at communicator creation: NCCL sets point-to-point channels up lazily per
// direction, which otherwise costs hundreds of milliseconds inside the first resampling that needs a new direction.
static int mg_warm_p2p(sabc_engine* e) {
    NcclApi* nc = nccl_api();
    MgScratch& s = e->mg;
    SABC_CUDA(s.wall.ensure((size_t)2 * e->world));
    SABC_CUDA(cudaMemsetAsync(s.wall.p, 0, sizeof(unsigned long long) * 2 * e->world, e->stream));
    SABC_NCCL(nc->GroupStart());
    for (int g = 0; g < e->world; ++g) {
        if (g == e->rank) continue;
        SABC_NCCL(nc->Send(s.wall.p + e->rank, 1, ncclUint64, g, e->comm.comm, e->stream));
        SABC_NCCL(nc->Recv(s.wall.p + e->world + g, 1, ncclUint64, g, e->comm.comm, e->stream));
    }
    SABC_NCCL(nc->GroupEnd());
    SABC_TRY(mg_allreduce_u64(e, s.wall.p, 1));
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Compressed ECDF over a sharded population (multi-GPU half): the table keeps the K rank-uniform
// quantiles x[floor(i (m-1) / (K-1))] of the m GLOBAL positive prior distances.  No rank ever holds the global column: every
// rank sorts its own slice, and the K order statistics are found by K simultaneous bisections on the 64-bit pattern of the value
// (monotone for positive doubles) -- each step counts the local entries <= the K candidates by binary search and all-reduces the
// K counts.  64 steps of a K-word all-reduce instead of an all-gather of 8 N bytes and a sort of N values on every rank; the
// knots equal the single-GPU compressed table for the same K bit for bit (tests/test_gpu_multi.py).
// ------------------------------------------------------------------------------------------------
static __global__ void k_q_init(int K, int64_t m, unsigned long long* lo, unsigned long long* hi, long long* target) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) {
        lo[i] = 0ull; hi[i] = 0x7fefffffffffffffull;                       // [+0, largest finite]
        target[i] = (long long)(((int64_t)i * (m - 1)) / (int64_t)(K - 1)) + 1;   // rank (1-based) of quantile i
    }
}
// number of local sorted positives <= the mid-point candidate of every bisection
static __global__ void k_q_count(int K, const double* sorted, int64_t n_pos, const unsigned long long* lo, const unsigned long long* hi,
                                 long long* cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) {
        const double v = __longlong_as_double((long long)(lo[i] + (hi[i] - lo[i]) / 2));
        int64_t a = 0, len = n_pos;                                           // first index with sorted[idx] > v
        while (len > 0) { const int64_t h = len >> 1; if (sorted[a + h] <= v) { a += h + 1; len -= h + 1; } else len = h; }
        cnt[i] = (long long)a;
    }
}
static __global__ void k_q_step(int K, unsigned long long* lo, unsigned long long* hi, const long long* cnt, const long long* target) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) {
        const unsigned long long mid = lo[i] + (hi[i] - lo[i]) / 2;
        if (cnt[i] >= target[i]) hi[i] = mid; else lo[i] = mid + 1;
    }
}
static __global__ void k_q_knots(int K, const unsigned long long* lo, double* out /* K + 2 + ECDF_PAD */) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K + 2 + ECDF_PAD; i += gridDim.x * blockDim.x) {
        double v;
        if (i == 0) v = 0.0;
        else if (i <= K) v = __longlong_as_double((long long)lo[i - 1]);
        else if (i == K + 1) v = __longlong_as_double((long long)lo[K - 1]) * 1.5;
        else v = dinf();
        out[i] = v;
    }
}
static int mg_ecdf_quantiles(sabc_engine* e, int j, const double* d_col, DevBuf<double>& keys, DevBuf<unsigned char>& cub_tmp,
                             DevBuf<unsigned long long>& cnt, int* done) {
    *done = 0;
    const int64_t n = e->n_local;
    const int K = e->ecdf_max_knots;
    SABC_CUDA(keys.ensure((size_t)2 * n));                                      // marked keys | sorted keys
    double* sorted = keys.p + n;
    SABC_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), e->stream));
    k_mark_positive<<<(int)std::min<int64_t>((n + 255) / 256, 8192), 256, 0, e->stream>>>(d_col, n, keys.p, cnt.p);
    SABC_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    SABC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys.p, sorted, (int64_t)n, 0, 64, e->stream));
    SABC_CUDA(cub_tmp.ensure(tmp_bytes));
    SABC_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp.p, tmp_bytes, keys.p, sorted, (int64_t)n, 0, 64, e->stream));
    unsigned long long n_pos_local = 0, m = 0;
    SABC_CUDA(cudaMemcpyAsync(&n_pos_local, cnt.p, sizeof n_pos_local, cudaMemcpyDeviceToHost, e->stream));
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    std::vector<unsigned long long> all(e->world);
    SABC_TRY(mg_allgather_u64_host(e, cnt.p, all.data()));
    for (auto v : all) m += v;
    if (m == 0) return set_error(SABC_ERR_NO_POSITIVE, "build_cdf: statistic %d has no positive prior distance", j + 1);
    if ((int64_t)m <= K) return 0;                                               // the whole sample is kept: the general path builds it
    DevBuf<unsigned long long> lo, hi;
    DevBuf<long long> target, c;
    SABC_CUDA(lo.alloc(K)); SABC_CUDA(hi.alloc(K)); SABC_CUDA(target.alloc(K)); SABC_CUDA(c.alloc(K));
    const int grid = (K + 255) / 256;
    k_q_init<<<grid, 256, 0, e->stream>>>(K, (int64_t)m, lo.p, hi.p, target.p);
    for (int it = 0; it < 64; ++it) {
        k_q_count<<<grid, 256, 0, e->stream>>>(K, sorted, (int64_t)n_pos_local, lo.p, hi.p, c.p);
        SABC_CUDA(cudaGetLastError());
        SABC_NCCL(nccl_api()->AllReduce(c.p, c.p, (size_t)K, ncclInt64, ncclSum, e->comm.comm, e->stream));
        k_q_step<<<grid, 256, 0, e->stream>>>(K, lo.p, hi.p, c.p, target.p);
    }
    auto* small = new DevBuf<double>();
    e->ecdf_bufs.push_back(small);
    SABC_CUDA(small->alloc((size_t)K + 2 + ECDF_PAD));
    k_q_knots<<<(K + 2 + ECDF_PAD + 255) / 256, 256, 0, e->stream>>>(K, lo.p, small->p);
    SABC_CUDA(cudaGetLastError());
    SABC_CUDA(cudaStreamSynchronize(e->stream));
    SABC_TRY(ecdf_attach(e, j, small, (int64_t)K + 2, std::max(top_max_for(e->S), pow2_ceil(K + 2))));
    *done = 1;
    return 0;
}

// the surplus exchange: rank g's packed selection occupies the global slots [C_g, C_g + c_g), rank d owns [d n, (d+1) n);
// only what crosses a slice boundary travels (grouped ncclSend/ncclRecv), then the statistics of the new population
static int mg_exchange_selection(sabc_engine* e, const std::vector<int64_t>& counts, int64_t sb_ld) {
    MgScratch& s = e->mg;
    NcclApi* nc = nccl_api();
    const int G = e->world, me = e->rank;
    const int64_t n = e->n_local;
    DevState* ds = e->b_ds.p;
    const int nf = e->D + e->S + 1;
    std::vector<int64_t> s_off(G), s_cnt(G), r_off(G), r_cnt(G);
    SABC_TRY(sabc_mg_exchange_plan(counts.data(), G, n, me, s_off.data(), s_cnt.data(), r_off.data(), r_cnt.data()));
    auto field_dst = [&](int f) -> double* {
        if (f < e->D) return e->tmp.theta + (int64_t)f * e->tmp.ld;
        if (f < e->D + e->S) return e->tmp.u + (int64_t)(f - e->D) * e->tmp.ld;
        return e->tmp.lp;
    };
    for (int f = 0; f < nf; ++f)               // slots I fill from my own selection
        if (s_cnt[me] > 0)
            SABC_CUDA(cudaMemcpyAsync(field_dst(f) + r_off[me], s.sb.p + (int64_t)f * sb_ld + s_off[me],
                                      (size_t)s_cnt[me] * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    SABC_NCCL(nc->GroupStart());
    for (int f = 0; f < nf; ++f) {
        for (int d = 0; d < G; ++d)
            if (d != me && s_cnt[d] > 0)
                SABC_NCCL(nc->Send(s.sb.p + (int64_t)f * sb_ld + s_off[d], (size_t)s_cnt[d], ncclFloat64, d, e->comm.comm, e->stream));
        for (int g = 0; g < G; ++g)
            if (g != me && r_cnt[g] > 0)
                SABC_NCCL(nc->Recv(field_dst(f) + r_off[g], (size_t)r_cnt[g], ncclFloat64, g, e->comm.comm, e->stream));
    }
    SABC_NCCL(nc->GroupEnd());
    const int g_grp = (int)std::min<int64_t>((n + CHUNK - 1) / CHUNK, (int64_t)e->n_sm * 8);
    k_copyback<<<g_grp, CHUNK, 0, e->stream>>>(e->pop, e->tmp, n, e->D, e->S, ds, 1);
    k_sum_u<<<g_grp, CHUNK, 0, e->stream>>>(e->pop.u, e->pop.ld, n, e->S, &ds->r_hi[0], &ds->r_lo[0]);
    SABC_CUDA(cudaGetLastError());
    SABC_TRY(mg_allreduce_u64(e, &ds->r_hi[0], 2 * MAX_S));
    e->n_resampling += 1;
    return 0;
}

// strict variant (SABC_FLAG_MG_STRICT_RESAMPLE): every rank evaluates the same N global variates against the all-gathered weight
// totals and keeps the draws that land in its own range, so the resampled MULTISET equals the single-GPU one for the same seed
// (tests/test_gpu_multi.py); O(N_global) work and memory per rank
static int mg_resample_strict(sabc_engine* e, const std::vector<unsigned long long>& w) {
    MgScratch& s = e->mg;
    const int G = e->world, me = e->rank;
    const int64_t n = e->n_local, N = e->N;
    DevState* ds = e->b_ds.p;
    unsigned long long W = 0, my_off = 0;
    for (int g = 0; g < G; ++g) { if (g == me) my_off = W; W += w[g]; }
    std::vector<unsigned long long> cnt(G);
    SABC_CUDA(s.F.ensure((size_t)N)); SABC_CUDA(s.src.ensure((size_t)N));
    const int64_t N_tiles = (N + TILE - 1) / TILE;
    SABC_CUDA(s.tsum.ensure((size_t)N_tiles)); SABC_CUDA(s.toff.ensure((size_t)N_tiles)); SABC_CUDA(s.scalar.ensure(1));
    const int g_all = (int)std::min<int64_t>((N + CHUNK - 1) / CHUNK, (int64_t)e->n_sm * 8);
    const int g_alltiles = (int)std::min<int64_t>(N_tiles, (int64_t)e->n_sm * 8);
    k_mg_mark<<<g_all, CHUNK, 0, e->stream>>>(N, W, my_off, w[me], e->b_q.p, n, e->seed, (uint32_t)e->n_resampling, s.F.p, s.src.p);
    k_tile_sums<<<g_alltiles, CHUNK, 0, e->stream>>>(s.F.p, N, s.tsum.p);
    k_scan_tiles<<<1, 1024, 0, e->stream>>>(s.tsum.p, N_tiles, s.toff.p, s.scalar.p, ds, 1);
    k_prefix<<<g_alltiles, CHUNK, 0, e->stream>>>(s.F.p, N, s.toff.p, ds, 1);
    SABC_CUDA(cudaGetLastError());
    SABC_TRY(mg_allgather_u64_host(e, s.scalar.p, cnt.data()));
    const int64_t sb_ld = std::max<int64_t>((int64_t)cnt[me], 1);
    SABC_CUDA(s.sb.ensure((size_t)(e->D + e->S + 1) * sb_ld));
    k_mg_pack<<<g_all, CHUNK, 0, e->stream>>>(e->pop, e->D, e->S, N, s.F.p, s.src.p, s.sb.p, sb_ld);
    SABC_CUDA(cudaGetLastError());
    return mg_exchange_selection(e, std::vector<int64_t>(cnt.begin(), cnt.end()), sb_ld);
}

// c_me draws among this rank's own particles, fused with the packing of the selected particles (field-major, draw order).
// Draw t of this rank is global draw C_me + t of the resampling: its own Philox block.
static __global__ void __launch_bounds__(CHUNK) k_mg_draw_pack(PopView pop, int D, int S, int64_t c_me, int64_t C_me, unsigned long long w_me,
                                                        const unsigned long long* P, int64_t n_local, uint64_t seed, uint32_t rc,
                                                        double* sb, int64_t sb_ld) {
    for (int64_t t = (int64_t)blockIdx.x * CHUNK + threadIdx.x; t < c_me; t += (int64_t)gridDim.x * CHUNK) {
        const uint64_t k = (uint64_t)(C_me + t);
        const U64x2 w = philox4x32_10((uint32_t)k, rc, (uint32_t)(k >> 32), KIND_RESAMPLE, (uint32_t)seed, (uint32_t)(seed >> 32));
        int64_t src = upper_bound_u64(P, n_local, mulhi64(w.a, w_me));
        if (src >= n_local) src = n_local - 1;
        for (int c = 0; c < D; ++c) sb[c * sb_ld + t] = pop.theta[c * pop.ld + src];
        for (int j = 0; j < S; ++j) sb[(D + j) * sb_ld + t] = pop.u[j * pop.ld + src];
        sb[(D + S) * sb_ld + t] = pop.lp[src];
    }
}

// exact global multinomial resampling over all ranks (resample_population, :124-137): the per-rank counts are ONE multinomial
// draw from a shared seed (multinomial.h), each rank then draws its c_g particles locally -- O(N / G) work and memory per rank
static int mg_resample(sabc_engine* e) {
    MgScratch& s = e->mg;
    const int G = e->world, me = e->rank;
    const int64_t n = e->n_local, N = e->N;
    DevState* ds = e->b_ds.p;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    const int g_tiles = (int)std::min<int64_t>(n_tiles, (int64_t)e->n_sm * 8);
    // local weights and prefix sums
    k_weights<<<g_tiles, CHUNK, 0, e->stream>>>(e->pop, n, e->S, e->delta, ds, e->b_q.p, e->b_tile_sum.p, 1);
    k_scan_tiles<<<1, 1024, 0, e->stream>>>(e->b_tile_sum.p, n_tiles, e->b_tile_off.p, &ds->w_total, ds, 2);
    k_prefix<<<g_tiles, CHUNK, 0, e->stream>>>(e->b_q.p, n, e->b_tile_off.p, ds, 1);
    SABC_CUDA(cudaGetLastError());
    std::vector<unsigned long long> w(G);
    SABC_TRY(mg_allgather_u64_host(e, &ds->w_total, w.data()));
    unsigned long long W = 0;
    for (int g = 0; g < G; ++g) W += w[g];
    if (W == 0) return set_error(SABC_ERR_INVALID, "all resampling weights are zero");
    if (e->flags & SABC_FLAG_MG_STRICT_RESAMPLE) return mg_resample_strict(e, w);
    std::vector<int64_t> counts(G);
    multinomial_split(N, w.data(), G, e->seed, (uint32_t)e->n_resampling, counts.data());
    int64_t C_me = 0;
    for (int g = 0; g < me; ++g) C_me += counts[g];
    const int64_t c_me = counts[me], sb_ld = std::max<int64_t>(c_me, 1);
    SABC_CUDA(s.sb.ensure((size_t)(e->D + e->S + 1) * sb_ld));
    if (c_me > 0) {
        const int grid = (int)std::min<int64_t>((c_me + CHUNK - 1) / CHUNK, (int64_t)e->n_sm * 8);
        k_mg_draw_pack<<<grid, CHUNK, 0, e->stream>>>(e->pop, e->D, e->S, c_me, C_me, w[me], e->b_q.p, n, e->seed, (uint32_t)e->n_resampling,
                                                      s.sb.p, sb_ld);
        SABC_CUDA(cudaGetLastError());
    }
    return mg_exchange_selection(e, counts, sb_ld);
}

// the per-update statistics of this rank, packed for ONE all-gather: [u_hi[S] | u_lo[S] | n_acc | rho_sum half 0 [S] | half 1 [S]]
static __global__ void k_mg_pack_stats(const DevState* ds, int S, unsigned long long* out) {
    const int t = threadIdx.x;
    if (t < S) {
        out[t] = ds->u_hi[t]; out[S + t] = ds->u_lo[t];
        out[2 * S + 1 + t] = (unsigned long long)__double_as_longlong(ds->rho_sum[0][t]);
        out[3 * S + 1 + t] = (unsigned long long)__double_as_longlong(ds->rho_sum[1][t]);
    }
    if (t == 0) out[2 * S] = ds->n_acc_iter;
}
// every rank reduces the gathered statistics in rank order (integer sums exact, rho sums in one fixed order: identical on all
// ranks), then takes the decision of :334-343.  A due resampling raises `hold`: see DevState.
static __global__ void k_mg_decide(DevState* ds, const unsigned long long* all, int G, int S, int64_t n_global, int64_t resample) {
    if (halted(ds)) return;
    const int W = 4 * S + 1, t = threadIdx.x;
    if (t < S) {
        unsigned long long hi = 0, lo = 0;
        double r0 = 0.0, r1 = 0.0;
        for (int g = 0; g < G; ++g) {
            hi += all[g * W + t]; lo += all[g * W + S + t];
            const double a = __longlong_as_double((long long)all[g * W + 2 * S + 1 + t]), b = __longlong_as_double((long long)all[g * W + 3 * S + 1 + t]);
            r0 = g == 0 ? a : r0 + a; r1 = g == 0 ? b : r1 + b;
        }
        ds->u_hi[t] = hi; ds->u_lo[t] = lo; ds->rho_sum[0][t] = r0; ds->rho_sum[1][t] = r1;
    }
    __syncthreads();
    if (t == 0) {
        unsigned long long na = 0;
        for (int g = 0; g < G; ++g) na += all[g * W + 2 * S];
        ds->n_acc_iter = na;
        post1_decide(ds, S, n_global, resample);
        ds->hold = ds->resample_flag;
    }
}
static __global__ void k_mg_release(DevState* ds) { ds->hold = 0; }

// one population update on this rank's slice, enqueue only: no host round trip, so a run of updates is queued ahead of the GPU
// (and replayed as a CUDA graph).  One collective per update; RandomWalk adds the two moment all-reduces of update_proposal!.
static int mg_enqueue_iteration(sabc_engine* e) {
    DevState* ds = e->b_ds.p;
    MgScratch& s = e->mg;
    const int W = 4 * e->S + 1;
    SABC_TRY(enqueue_sweeps(e));
    SABC_TRY(launch_post1(e, 0));
    k_mg_pack_stats<<<1, 32, 0, e->stream>>>(ds, e->S, s.stats_send.p);
    SABC_CUDA(cudaGetLastError());
    SABC_NCCL(nccl_api()->AllGather(s.stats_send.p, s.stats_all.p, (size_t)W, ncclUint64, e->comm.comm, e->stream));
    k_mg_decide<<<1, 32, 0, e->stream>>>(ds, s.stats_all.p, e->world, e->S, e->N, e->resample);
    SABC_CUDA(cudaGetLastError());
    SABC_TRY(launch_update_proposal_mg(e));
    SABC_TRY(launch_finish(e));
    return 0;
}
// the host half of an update whose decision raised `hold`: the global resampling, then the rest of that update.  Everything is
// stream-ordered behind the squashed updates; the only host wait is the one inside mg_resample (the per-rank weight totals the
// multinomial split needs), so the next updates are enqueued while the exchange is still running.
static int mg_complete_held_iteration(sabc_engine* e) {
    const auto t0 = std::chrono::steady_clock::now();
    SABC_TRY(mg_resample(e));
    k_mg_release<<<1, 1, 0, e->stream>>>(e->b_ds.p);
    SABC_CUDA(cudaGetLastError());
    SABC_TRY(launch_update_proposal_mg(e));
    SABC_TRY(launch_finish(e));
    e->timing.resample_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    e->timing.resample_events += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Replicated multi-GPU mode (SABC_FLAG_MG_REPLICATED): every rank keeps the WHOLE population and simulates a contiguous
// share of each half-sweep; the updated rows are broadcast after the sweep and everything else (statistics, resampling,
// eps) runs replicated.  Halves, partners and slots are those of the single-GPU algorithm, so the result is bit-identical
// to one GPU -- the "strict" mode of SURVEY.md section 8e, used for parity studies (memory does not scale).
// ------------------------------------------------------------------------------------------------
static __global__ void k_zero_u_sums(DevState* ds) {
    if (threadIdx.x < MAX_S) { ds->u_hi[threadIdx.x] = 0ull; ds->u_lo[threadIdx.x] = 0ull; }
}
static int rep_iteration(sabc_engine* e) {
    NcclApi* nc = nccl_api();
    DevState* ds = e->b_ds.p;
    const int G = e->world;
    for (int half = 0; half < 2; ++half) {
        if (e->split) { SABC_TRY(launch_split_propose(e, half, e->rank, G)); SABC_TRY(launch_split_simacc(e, half, e->rank, G)); }
        else SABC_TRY(launch_update_half(e, half, e->rank, G));
        int64_t off, an, t0, t1;
        halves(e, half, off, an, t0, t1);
        SABC_NCCL(nc->GroupStart());
        for (int g = 0; g < G; ++g) {
            int64_t r0, r1;
            sub_range(an, g, G, r0, r1);
            if (r1 <= r0) continue;
            const size_t cnt = (size_t)(r1 - r0);
            for (int c = 0; c < e->D; ++c) { double* p = e->pop.theta + (int64_t)c * e->pop.ld + off + r0; SABC_NCCL(nc->Broadcast(p, p, cnt, ncclFloat64, g, e->comm.comm, e->stream)); }
            for (int j = 0; j < e->S; ++j) {
                double* pu = e->pop.u + (int64_t)j * e->pop.ld + off + r0; SABC_NCCL(nc->Broadcast(pu, pu, cnt, ncclFloat64, g, e->comm.comm, e->stream));
                double* pr = e->pop.rho + (int64_t)j * e->pop.ld + off + r0; SABC_NCCL(nc->Broadcast(pr, pr, cnt, ncclFloat64, g, e->comm.comm, e->stream));
            }
            double* pl = e->pop.lp + off + r0; SABC_NCCL(nc->Broadcast(pl, pl, cnt, ncclFloat64, g, e->comm.comm, e->stream));
        }
        SABC_NCCL(nc->GroupEnd());
    }
    // accept count of all shares; statistics over the replicated arrays (the sweeps only saw a share)
    SABC_TRY(mg_allreduce_u64(e, &ds->n_acc_iter, 1));
    k_zero_u_sums<<<1, 32, 0, e->stream>>>(ds);
    SABC_CUDA(cudaGetLastError());
    SABC_TRY(launch_split_stats(e));
    if (small_tail(e)) return launch_tail_small(e);
    SABC_TRY(launch_post1(e, 1));
    SABC_TRY(launch_resample_local(e, 0));
    SABC_TRY(launch_update_proposal(e));
    SABC_TRY(launch_finish(e));
    return 0;
}
