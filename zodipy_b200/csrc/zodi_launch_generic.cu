// zodi_launch_generic.cu - instantiations of zodi_los_generic_kernel for one arithmetic type
// (compiled once per type: -DZODI_TU_REAL=float|double -DZODI_TU_SUFFIX=f32|f64).
#include "zodi_launch.hpp"

namespace zodi {

namespace {
template <typename Real, int L>
cudaError_t launch_generic_L(const DevModel<Real>& M, const LaunchArgs& a, const Pair<Real>* tab,
                             const Pair<Real>* nodes, cudaStream_t stream) {
    const int per_cta = kThreads / L;
    const int64_t grid = (a.n + per_cta - 1) / per_cta;
    const size_t smem = (size_t)(M.n_temps + M.n_nodes) * sizeof(Pair<Real>);
    // static fp64 math tables (17 KB) + this: above the 48 KB default for the largest descriptors
    // the ABI admits (1024 knots + 1024 nodes in double = 32 KB)
    if (smem > 16 * 1024) {  // per device and cheap: no caching
        cudaError_t e = cudaFuncSetAttribute(zodi_los_generic_kernel<Real, L>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
    }
    zodi_los_generic_kernel<Real, L><<<(unsigned)grid, kThreads, smem, stream>>>(M, a, tab, nodes);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}
}  // namespace

#define ZODI_CAT2(a, b) a##b
#define ZODI_CAT(a, b) ZODI_CAT2(a, b)

cudaError_t ZODI_CAT(launch_generic_, ZODI_TU_SUFFIX)(const DevModel<ZODI_TU_REAL>& M, const LaunchArgs& a,
                                                      const Pair<ZODI_TU_REAL>* tab, const Pair<ZODI_TU_REAL>* nodes,
                                                      cudaStream_t stream) {
    using Real = ZODI_TU_REAL;
    switch (pick_lanes(a.shape_n > 0 ? a.shape_n : a.n, M.n_nodes)) {
        case 1: return launch_generic_L<Real, 1>(M, a, tab, nodes, stream);
        case 2: return launch_generic_L<Real, 2>(M, a, tab, nodes, stream);
        case 4: return launch_generic_L<Real, 4>(M, a, tab, nodes, stream);
        case 8: return launch_generic_L<Real, 8>(M, a, tab, nodes, stream);
        case 16: return launch_generic_L<Real, 16>(M, a, tab, nodes, stream);
        default: return launch_generic_L<Real, 32>(M, a, tab, nodes, stream);
    }
}

}  // namespace zodi
