// zodi_launch_multiband.cu - instantiations of the multi-band kernel for one arithmetic type
// (compiled once per type: -DZODI_TU_REAL=float|double -DZODI_TU_SUFFIX=f32|f64).
#include "zodi_launch.hpp"

namespace zodi {

namespace {
template <typename Real, int NB>
cudaError_t launch_multiband_NB(const MultiBandModel<Real>& MB, const LaunchArgs& a, const Pair<Real>* tabs,
                                const Pair<Real>* nodes, cudaStream_t stream) {
    const unsigned grid = (unsigned)((a.n + kThreads - 1) / kThreads);
    const bool rf = MB.base.n_comps == 6, sc = MB.base.scatter != 0;
    if (rf && sc) zodi_los_multiband_kernel<Real, NB, true, true><<<grid, kThreads, 0, stream>>>(MB, a, tabs, nodes);
    else if (rf) zodi_los_multiband_kernel<Real, NB, true, false><<<grid, kThreads, 0, stream>>>(MB, a, tabs, nodes);
    else if (sc) zodi_los_multiband_kernel<Real, NB, false, true><<<grid, kThreads, 0, stream>>>(MB, a, tabs, nodes);
    else zodi_los_multiband_kernel<Real, NB, false, false><<<grid, kThreads, 0, stream>>>(MB, a, tabs, nodes);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

#if defined(ZODI_TU_IS_F32)
template <int NB>
cudaError_t launch_multiband_x2_NB(const MultiBandModel<float>& MB, const LaunchArgs& a, const Pair<float>* tabs,
                                   const Pair<float>* nodes, cudaStream_t stream) {
    constexpr int per_cta = 2 * kPackedDefaultThreads;
    const unsigned grid = (unsigned)((a.n + per_cta - 1) / per_cta);
    const bool rf = MB.base.n_comps == 6, sc = MB.base.scatter != 0;
    if (rf && sc) zodi_los_multiband_x2_kernel<NB, true, true><<<grid, kPackedDefaultThreads, 0, stream>>>(MB, a, tabs, nodes);
    else if (rf) zodi_los_multiband_x2_kernel<NB, true, false><<<grid, kPackedDefaultThreads, 0, stream>>>(MB, a, tabs, nodes);
    else if (sc) zodi_los_multiband_x2_kernel<NB, false, true><<<grid, kPackedDefaultThreads, 0, stream>>>(MB, a, tabs, nodes);
    else zodi_los_multiband_x2_kernel<NB, false, false><<<grid, kPackedDefaultThreads, 0, stream>>>(MB, a, tabs, nodes);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}
#endif
}  // namespace

#define ZODI_CAT2(a, b) a##b
#define ZODI_CAT(a, b) ZODI_CAT2(a, b)

cudaError_t ZODI_CAT(launch_multiband_, ZODI_TU_SUFFIX)(const MultiBandModel<ZODI_TU_REAL>& MB, const LaunchArgs& a,
                                                        const Pair<ZODI_TU_REAL>* tabs,
                                                        const Pair<ZODI_TU_REAL>* nodes, cudaStream_t stream) {
    using Real = ZODI_TU_REAL;
    if (MB.n_bands <= 4) return launch_multiband_NB<Real, 4>(MB, a, tabs, nodes, stream);
    if (MB.n_bands <= 8) return launch_multiband_NB<Real, 8>(MB, a, tabs, nodes, stream);
    return launch_multiband_NB<Real, 16>(MB, a, tabs, nodes, stream);
}

#if defined(ZODI_TU_IS_F32)
cudaError_t launch_multiband_packed(const MultiBandModel<float>& MB, const LaunchArgs& a, const Pair<float>* tabs,
                                    const Pair<float>* nodes, cudaStream_t stream) {
    if (MB.n_bands <= 4) return launch_multiband_x2_NB<4>(MB, a, tabs, nodes, stream);
    if (MB.n_bands <= 8) return launch_multiband_x2_NB<8>(MB, a, tabs, nodes, stream);
    return launch_multiband_x2_NB<16>(MB, a, tabs, nodes, stream);
}
#endif

}  // namespace zodi
