// zodi_launch.hpp - launch entry points of the integrator kernel families.
//
// The kernels are instantiated in separate translation units (zodi_launch_*.cu) so that the library
// builds in parallel; zodi_capi.cu (the C ABI) only sees these functions.  Every function enqueues
// ONE kernel on `stream` and returns the launch status.
#pragma once

#include <atomic>
#include <cstdint>

#include <cuda_runtime.h>

#include "zodi_kernels.cuh"

namespace zodi {

extern std::atomic<int64_t> g_launches;  // kernels launched by this library (zodi_kernel_launch_count)

// Number of SMs of the current device (cached per device ordinal).
int sm_count();

// Lanes per line of sight of the scalar kernels: enough threads to fill the machine (SMs x 2048
// resident threads), never more lanes than quadrature nodes.
int pick_lanes(int64_t n, int n_nodes);

// Shape of a packed-kernel launch: lanes per PAIR of lines of sight and threads per CTA.
struct PackedShape { int lanes; int threads; };
PackedShape pick_packed_shape(int64_t n, int n_nodes, int n_comps);

cudaError_t launch_generic_f32(const DevModel<float>& M, const LaunchArgs& a, const Pair<float>* tab,
                               const Pair<float>* nodes, cudaStream_t stream);
cudaError_t launch_generic_f64(const DevModel<double>& M, const LaunchArgs& a, const Pair<double>* tab,
                               const Pair<double>* nodes, cudaStream_t stream);
cudaError_t launch_kelsall_f32(const KelsallModel<float>& K, const LaunchArgs& a, const Pair<float>* tab,
                               const Pair<float>* nodes, cudaStream_t stream);
cudaError_t launch_kelsall_f64(const KelsallModel<double>& K, const LaunchArgs& a, const Pair<double>* tab,
                               const Pair<double>* nodes, cudaStream_t stream);
// packed fp32 kernels (zodi_kelsall_x2.cuh); lanes in {1, 2, 4, 8}, threads in {128, 256}
cudaError_t launch_kelsall_packed(const KelsallModel<float>& K, const LaunchArgs& a, const Pair<float>* tab,
                                  const Pair<float>* nodes, PackedShape shape, cudaStream_t stream);
cudaError_t launch_multiband_f32(const MultiBandModel<float>& MB, const LaunchArgs& a, const Pair<float>* tabs,
                                 const Pair<float>* nodes, cudaStream_t stream);
// packed fp32 multi-band kernel (zodi_multiband_x2.cuh): two lines of sight per thread
cudaError_t launch_multiband_packed(const MultiBandModel<float>& MB, const LaunchArgs& a, const Pair<float>* tabs,
                                    const Pair<float>* nodes, cudaStream_t stream);
cudaError_t launch_multiband_f64(const MultiBandModel<double>& MB, const LaunchArgs& a, const Pair<double>* tabs,
                                 const Pair<double>* nodes, cudaStream_t stream);
// fused RRM kernels (zodi_rrm.cuh)
cudaError_t launch_rrm_f32(const RrmModel<float>& R, const LaunchArgs& a, const Pair<float>* tab,
                           const Pair<float>* nodes, cudaStream_t stream);
cudaError_t launch_rrm_packed(const RrmModelX2& X, const LaunchArgs& a, const Pair<float>* tab, const Pair<float>* nodes,
                              cudaStream_t stream);
cudaError_t launch_rrm_f64(const RrmModel<double>& R, const LaunchArgs& a, const Pair<double>* tab,
                           const Pair<double>* nodes, cudaStream_t stream);

}  // namespace zodi
