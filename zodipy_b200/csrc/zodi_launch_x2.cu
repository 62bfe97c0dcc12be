// zodi_launch_x2.cu - instantiations of the packed fp32 kernel for ONE lane count
// (compiled once per count: -DZODI_TU_LANES=1|2|4|8).
#include "zodi_launch.hpp"

namespace zodi {

namespace {
constexpr int L = ZODI_TU_LANES;

// 5 CTAs of 256 threads per SM (48 registers): cloud+bands only measured 5 % faster than 4 CTAs/SM (60
// registers) and 0.4 % faster than 6 (40 registers).  With the ring/feature loops: 4 CTAs/SM (64 registers)
// and those two loops unrolled x5 - their per-node chains (sqrt / lg2 / ex2 / table / ex2) then overlap - is
// 2.2 % (nside 1024) to 6.8 % (nside 64) faster than 5 CTAs/SM without unrolling; without unrolling the two
// occupancies measured equal.  The scattering terms need more registers (3 CTAs/SM; 4 measured equal).  128-thread CTAs
// keep the same number of resident warps in twice as many, half as long CTAs (shorter tail of a launch);
// 64-thread CTAs measured the same again (profiles/r2_packed_cta_size_sweep.jsonl, r2_bench_n8_t128.json)
// and were dropped.
#ifndef ZODI_X2_CTAS_THERMAL
#define ZODI_X2_CTAS_THERMAL 5
#endif
#ifndef ZODI_X2_CTAS_RF
#define ZODI_X2_CTAS_RF 4  // 64 registers: room for the ring / feature loops unrolled x5 (ZODI_X2_RF_UNROLL)
#endif

template <bool HAS_RF, bool SHARE13, bool SCATTER, int THREADS>
cudaError_t launch_T(const KelsallModel<float>& K, const LaunchArgs& a, const Pair<float>* tab,
                     const Pair<float>* nodes, cudaStream_t stream) {
    constexpr int per_cta = 2 * THREADS / L;
    const int64_t grid = (a.n + per_cta - 1) / per_cta;
    constexpr int kCtas256 = SCATTER ? 3 : (HAS_RF ? ZODI_X2_CTAS_RF : ZODI_X2_CTAS_THERMAL);
    constexpr int kMinCtas = kCtas256 * (256 / THREADS);
    // persistent tiles (LaunchArgs::tile_counter; instantiated for the large-N shape only: L = 1, 128-thread
    // CTAs): one resident wave of CTAs, each claiming tiles until none is left
    bool persistent = false;
    if constexpr (L == 1 && THREADS == 128) {
        if (a.tile_counter != nullptr && grid > (int64_t)sm_count() * kMinCtas) {
            zodi_los_kelsall_x2_kernel<HAS_RF, SHARE13, SCATTER, 1, 128, kMinCtas, true>
                <<<(unsigned)(sm_count() * kMinCtas), 128, 0, stream>>>(K, a, tab, nodes);
            persistent = true;
        }
    }
    if (!persistent)
        zodi_los_kelsall_x2_kernel<HAS_RF, SHARE13, SCATTER, L, THREADS, kMinCtas>
            <<<(unsigned)grid, THREADS, 0, stream>>>(K, a, tab, nodes);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

template <bool HAS_RF, bool SHARE13, bool SCATTER>
cudaError_t launch_S(const KelsallModel<float>& K, const LaunchArgs& a, const Pair<float>* tab,
                     const Pair<float>* nodes, int threads, cudaStream_t stream) {
    return threads == 128 ? launch_T<HAS_RF, SHARE13, SCATTER, 128>(K, a, tab, nodes, stream)
                          : launch_T<HAS_RF, SHARE13, SCATTER, 256>(K, a, tab, nodes, stream);
}

template <bool HAS_RF, bool SHARE13>
cudaError_t launch_RS(const KelsallModel<float>& K, const LaunchArgs& a, const Pair<float>* tab,
                      const Pair<float>* nodes, int threads, cudaStream_t stream) {
    return K.scatter ? launch_S<HAS_RF, SHARE13, true>(K, a, tab, nodes, threads, stream)
                     : launch_S<HAS_RF, SHARE13, false>(K, a, tab, nodes, threads, stream);
}
}  // namespace

#define ZODI_CAT2(a, b) a##b
#define ZODI_CAT(a, b) ZODI_CAT2(a, b)

cudaError_t ZODI_CAT(launch_kelsall_packed_l, ZODI_TU_LANES)(const KelsallModel<float>& K, const LaunchArgs& a,
                                                             const Pair<float>* tab, const Pair<float>* nodes,
                                                             int threads, cudaStream_t stream) {
    if (K.n_comps == 6)
        return K.share13 ? launch_RS<true, true>(K, a, tab, nodes, threads, stream)
                         : launch_RS<true, false>(K, a, tab, nodes, threads, stream);
    return K.share13 ? launch_RS<false, true>(K, a, tab, nodes, threads, stream)
                     : launch_RS<false, false>(K, a, tab, nodes, threads, stream);
}

}  // namespace zodi
