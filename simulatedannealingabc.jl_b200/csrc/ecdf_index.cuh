// ecdf_index.cuh -- host side of the ECDF index (layout and lookup: plugin.cuh, "ECDF"): builds the probe-ordered levels over a
// knot table that is already on the device.  Shared by the engine and the parity hooks.
#pragma once
#include "aux_kernels.cuh"
#include "common.h"
#include <algorithm>
#include <cstring>
#include <vector>

namespace sabc {

inline int pow2_ceil(int64_t n) { int p = 2; while (p < n) p <<= 1; return p; }
// entries the staged top level may have: a power of two, 2048 for up to three statistics, less for more (shared memory)
inline int ecdf_top_max(int S) { int p = 2048; while (p > 64 && (int64_t)p * S > 6144) p >>= 1; return p; }

// `bufs` receives every device buffer allocated here (the caller owns them).  top_max must be a power of two.
inline int ecdf_build_index(EcdfStat& st, const double* d_knots, int64_t L, int top_max, std::vector<DevBuf<double>*>& bufs,
                            cudaStream_t stream) {
    std::memset(&st, 0, sizeof st);
    st.L = L; st.knots = d_knots; st.nlev = 1;
    SABC_CUDA(cudaMemcpyAsync(&st.kmax, d_knots + (L - 1), sizeof(double), cudaMemcpyDeviceToHost, stream));
    const double* cur = d_knots;
    int64_t cnt = L;
    auto grid_of = [](int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 4096)); };
    while (cnt > top_max) {
        if (st.nlev >= ECDF_MAX_LEVELS) return set_error(SABC_ERR_INVALID, "ECDF table too large for %d index levels", ECDF_MAX_LEVELS);
        const int64_t n_nodes = (cnt + ECDF_STRIDE - 1) / ECDF_STRIDE;
        auto* nodes = new DevBuf<double>(); bufs.push_back(nodes);
        auto* next = new DevBuf<double>(); bufs.push_back(next);
        SABC_CUDA(nodes->alloc((size_t)n_nodes * ECDF_NODE));
        SABC_CUDA(next->alloc((size_t)n_nodes));
        k_ecdf_level<<<grid_of(n_nodes * ECDF_NODE), 256, 0, stream>>>(cur, cnt, n_nodes, nodes->p, next->p, n_nodes);
        SABC_CUDA(cudaGetLastError());
        st.node[st.nlev - 1] = nodes->p;
        cur = next->p; cnt = n_nodes; st.nlev++;
    }
    st.top_cnt = (int)cnt; st.top_pow2 = pow2_ceil(cnt);
    auto* top = new DevBuf<double>(); bufs.push_back(top);
    SABC_CUDA(top->alloc((size_t)st.top_pow2));
    k_top_split<<<grid_of(st.top_pow2), 256, 0, stream>>>(cur, cnt, top->p, st.top_pow2);
    SABC_CUDA(cudaGetLastError());
    st.top = top->p;
    SABC_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

}  // namespace sabc
