// zodi_launch_kelsall.cu - instantiations of the scalar fused Kelsall-family kernel for one arithmetic
// type (compiled once per type: -DZODI_TU_REAL=float|double -DZODI_TU_SUFFIX=f32|f64).
#include "zodi_launch.hpp"

namespace zodi {

namespace {
template <typename Real, bool HAS_RF, bool SCATTER, bool SHARE13, int L>
cudaError_t launch_kelsall_L(const KelsallModel<Real>& K, const LaunchArgs& a, const Pair<Real>* tab,
                             const Pair<Real>* nodes, cudaStream_t stream) {
    // fp64 variants of 128 threads x 9 CTAs/SM (56 registers) and 256 x 5 (48 registers, spills)
    // measured within 1 % of this shape on B200: the kernel is bound by issue slots, not occupancy.
    const int per_cta = kThreads / L;
    const int64_t grid = (a.n + per_cta - 1) / per_cta;
    zodi_los_kelsall_kernel<Real, HAS_RF, SCATTER, SHARE13, L>
        <<<(unsigned)grid, kThreads, 0, stream>>>(K, a, tab, nodes);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <typename Real, bool HAS_RF, bool SCATTER, bool SHARE13>
cudaError_t launch_kelsall_RSS(const KelsallModel<Real>& K, const LaunchArgs& a, const Pair<Real>* tab,
                               const Pair<Real>* nodes, cudaStream_t stream) {
    if (pick_lanes(a.shape_n > 0 ? a.shape_n : a.n, K.n_nodes) == 1)
        return launch_kelsall_L<Real, HAS_RF, SCATTER, SHARE13, 1>(K, a, tab, nodes, stream);
    return launch_kelsall_L<Real, HAS_RF, SCATTER, SHARE13, 8>(K, a, tab, nodes, stream);
}

template <typename Real, bool HAS_RF, bool SCATTER>
cudaError_t launch_kelsall_RS(const KelsallModel<Real>& K, const LaunchArgs& a, const Pair<Real>* tab,
                              const Pair<Real>* nodes, cudaStream_t stream) {
    if (K.share13) return launch_kelsall_RSS<Real, HAS_RF, SCATTER, true>(K, a, tab, nodes, stream);
    return launch_kelsall_RSS<Real, HAS_RF, SCATTER, false>(K, a, tab, nodes, stream);
}
}  // namespace

#define ZODI_CAT2(a, b) a##b
#define ZODI_CAT(a, b) ZODI_CAT2(a, b)

cudaError_t ZODI_CAT(launch_kelsall_, ZODI_TU_SUFFIX)(const KelsallModel<ZODI_TU_REAL>& K, const LaunchArgs& a,
                                                      const Pair<ZODI_TU_REAL>* tab, const Pair<ZODI_TU_REAL>* nodes,
                                                      cudaStream_t stream) {
    using Real = ZODI_TU_REAL;
    if (K.n_comps == 6) {
        if (K.scatter) return launch_kelsall_RS<Real, true, true>(K, a, tab, nodes, stream);
        return launch_kelsall_RS<Real, true, false>(K, a, tab, nodes, stream);
    }
    if (K.scatter) return launch_kelsall_RS<Real, false, true>(K, a, tab, nodes, stream);
    return launch_kelsall_RS<Real, false, false>(K, a, tab, nodes, stream);
}

}  // namespace zodi
