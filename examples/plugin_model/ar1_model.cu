// Out-of-tree device model plug-in: the GPU form of a user-written f_dist (contract src/SimulatedAnnealingABC.jl:421).
//
//   θ = (φ, σ);  x_0 = 0, x_t = φ x_{t-1} + σ z_t  for t = 1..T (T = par[0] <= 64);
//   statistics: |lag-1 autocovariance − obs1|, |variance − obs2|         par = [T, obs1, obs2]
//
// Build (see build.sh) into a shared object, dlopen it AFTER libsabc_b200.so; the static initialiser below registers the
// launch table of the header-only kernel templates under the name "ar1".
#include "../../simulatedannealingabc.jl_b200/csrc/kernels.cuh"
#include "../../include/sabc_b200.h"

struct Ar1Model {
    static constexpr int D = 2, S = 2;
    static constexpr int FUSED_MIN_BLOCKS = 1; // resident CTAs per SM requested for the fused kernel
    static constexpr int SIM_MIN_BLOCKS = 1;   // resident CTAs per SM requested for the simulation kernel (split path)
    static constexpr int KEY_BITS = 0;         // no work-list bucketing
    SABC_HD static void sim(const double (&th)[2], const sabc::ModelPar& mp, sabc::Stream& st, double (&rho)[2]) {
        const int T = (int)mp.v[0];
        double x = 0.0, s1 = 0.0, s2 = 0.0, sx = 0.0;
        for (int t = 0; t < T; t += 2) {
            double z[2];
            sabc::normal2(st, z[0], z[1]);                     // one Philox block -> two normals (ziggurat)
            for (int h = 0; h < 2 && t + h < T; ++h) {
                const double xn = th[0] * x + th[1] * z[h];
                s1 = s1 + xn * x;                                // lag-1 cross product
                s2 = s2 + xn * xn;
                sx = sx + xn;
                x = xn;
            }
        }
        const double m = sx / (double)T;
        rho[0] = fabs(s1 / (double)T - mp.v[1]);
        rho[1] = fabs((s2 / (double)T - m * m) - mp.v[2]);
    }
};

static const int ar1_registered = [] {
    static sabc::ModelVTable vt = sabc::ModelLaunchers<Ar1Model>::vtable("ar1", /*heavy=*/0);
    return sabc_register_model(&vt);
}();
