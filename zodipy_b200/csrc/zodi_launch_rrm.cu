// zodi_launch_rrm.cu - instantiations of the fused RRM kernel for one arithmetic type
// (compiled once per type: -DZODI_TU_REAL=float|double -DZODI_TU_SUFFIX=f32|f64).
#include "zodi_launch.hpp"

namespace zodi {

namespace {
template <typename Real, int L>
cudaError_t launch_rrm_L(const RrmModel<Real>& R, const LaunchArgs& a, const Pair<Real>* tab,
                         const Pair<Real>* nodes, cudaStream_t stream) {
    const int per_cta = kThreads / L;
    const int64_t grid = (a.n + per_cta - 1) / per_cta;
    zodi_los_rrm_kernel<Real, L><<<(unsigned)grid, kThreads, 0, stream>>>(R, a, tab, nodes);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}
}  // namespace

#if defined(ZODI_TU_IS_F32)
// packed fp32 kernel: 128-thread CTAs, 8 per SM (64 registers)
cudaError_t launch_rrm_packed(const RrmModelX2& X, const LaunchArgs& a, const Pair<float>* tab, const Pair<float>* nodes,
                              cudaStream_t stream) {
    constexpr int kT = 128;
    const int64_t grid = (a.n + 2 * kT - 1) / (2 * kT);
    zodi_los_rrm_x2_kernel<kT, 8><<<(unsigned)grid, kT, 0, stream>>>(X, a, tab, nodes);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}
#endif

#define ZODI_CAT2(a, b) a##b
#define ZODI_CAT(a, b) ZODI_CAT2(a, b)

cudaError_t ZODI_CAT(launch_rrm_, ZODI_TU_SUFFIX)(const RrmModel<ZODI_TU_REAL>& R, const LaunchArgs& a,
                                                  const Pair<ZODI_TU_REAL>* tab, const Pair<ZODI_TU_REAL>* nodes,
                                                  cudaStream_t stream) {
    using Real = ZODI_TU_REAL;
    switch (pick_lanes(a.shape_n > 0 ? a.shape_n : a.n, R.n_nodes)) {
        case 1: return launch_rrm_L<Real, 1>(R, a, tab, nodes, stream);
        case 2: return launch_rrm_L<Real, 2>(R, a, tab, nodes, stream);
        case 4: return launch_rrm_L<Real, 4>(R, a, tab, nodes, stream);
        default: return launch_rrm_L<Real, 8>(R, a, tab, nodes, stream);
    }
}

}  // namespace zodi
