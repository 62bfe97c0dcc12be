// zodi_kernels.cuh - __global__ kernels of the line-of-sight integrator (sm_100a).
//
// Mapping: L consecutive lanes (L in {1,2,4,8,16,32}) share one line of sight and split its
// quadrature nodes (node k -> lane k mod L); partial sums are combined with warp shuffles.
// L = 1 is thread-per-line-of-sight (no lane idles when n_nodes is not a multiple of L; used for
// large N), larger L fills the machine when N is small.  Parameters live in the kernel-parameter
// constant bank (__grid_constant__), the blackbody table and the quadrature nodes are staged in
// shared memory once per CTA, inputs are read SoA/coalesced and outputs written coalesced.
#pragma once

#include <cuda_runtime.h>

#include "zodi_device.cuh"
#include "zodi_kelsall.cuh"
#include "zodi_kelsall_x2.cuh"
#include "zodi_multiband.cuh"
#include "zodi_multiband_x2.cuh"
#include "zodi_rrm.cuh"
#include "zodi_rrm_x2.cuh"

namespace zodi {

constexpr int kThreads = 256;
#ifndef ZODI_X2_DEFAULT_THREADS
#define ZODI_X2_DEFAULT_THREADS 128
#endif
constexpr int kPackedDefaultThreads = ZODI_X2_DEFAULT_THREADS;  // CTA size of the packed kernels

// Launch-time arguments (plain pointers; device memory).
struct LaunchArgs {
    int64_t n;
    // lines of sight of the whole C-ABI call this launch belongs to (host-memory calls are cut into
    // chunks): the launch shape (lanes per line of sight) is chosen from THIS number, so that the summation
    // order of the lane partials - and with it the result, bit for bit - does not depend on the chunking
    int64_t shape_n;
    const double* u;      int64_t u_stride;
    const double* obs;    int64_t obs_stride;   int obs_per_sample;    // 0: one observer
    const double* earth;  int64_t earth_stride; int earth_per_sample;
    uint32_t outside_mask;  // bit 2c: outside inner cutoff of comp c; bit 2c+1: outer
    int return_comps;
    int out_f32;
    void* out;            int64_t out_stride;
    // fused all-gather: store to every peer's map (see zodi_eval_args.peer_out)
    int n_peers;
    void* peer_out[ZODI_MAX_PEERS];
    int64_t peer_offset;  int64_t peer_stride;
    // on-device directions: HEALPix RING pixel hp_start + j (u is ignored when hp_nside > 0)
    int64_t hp_nside;     int64_t hp_start;
    int hp_nest;          // 1: hp_start + j is a NESTED index (converted to RING on the fly)
    int hp_rotate;        double hp_rot[9];   // row-major 3x3 applied to the pixel / lon-lat vector
    // on-device directions from spherical sky coordinates [rad] (u is ignored when lon != NULL)
    const double* lon;    const double* lat;
    // block-cyclic shard layout (see zodi_eval_args.cyclic_block); 0 = identity
    int64_t cyc_block;    int cyc_parts;      int cyc_rank;
    int cyc_shift = -1;   // log2(cyc_block) when it is a power of two (shift / mask instead of a 64-bit division)
    // on-device ephemeris (time-ordered data): positions from cubic splines at obstime[j]
    const double* eph_coef;   // [n_knots-1][3][4] Earth (highest power first); NULL = use obs/earth arrays
    const double* eph_obs_coef;  // same for the observer, or NULL: observer = eph_scale * Earth
    const double* obstime;
    int64_t eph_nseg;     double eph_t0, eph_dt, eph_scale;
    // persistent tiles (packed Kelsall kernel): two zero-initialised counters {claimed, CTAs done}; the CTAs of
    // a grid sized to the machine take further tiles from `claimed`, the last CTA to leave zeroes both again.
    // NULL: one tile per CTA.
    unsigned int* tile_counter = nullptr;
};

// Spline position at time t: interval i = floor((t - t0)/dt) clamped to the spline (scipy's PPoly
// extrapolates with the end polynomials), value = ((c0 s + c1) s + c2) s + c3 with s = t - knot_i.
__device__ __forceinline__ void spline_position(const double* __restrict__ coef, int64_t nseg, double t0,
                                                double dt, double t, double& x, double& y, double& z) {
    int64_t i = (int64_t)floor((t - t0) / dt);
    i = i < 0 ? 0 : (i > nseg - 1 ? nseg - 1 : i);
    const double s = t - (t0 + (double)i * dt);
    const double* c = coef + i * 12;
    x = fma(fma(fma(c[0], s, c[1]), s, c[2]), s, c[3]);
    y = fma(fma(fma(c[4], s, c[5]), s, c[6]), s, c[7]);
    z = fma(fma(fma(c[8], s, c[9]), s, c[10]), s, c[11]);
}

// Observer and Earth position of line of sight jj: arrays of the reference seam, or the splines.
__device__ __forceinline__ void load_positions(const LaunchArgs& a, int64_t jj, bool need_earth, double& ox,
                                               double& oy, double& oz, double& ex, double& ey) {
    if (a.eph_coef != nullptr) {
        const double t = a.obstime[jj];
        double ez;
        spline_position(a.eph_coef, a.eph_nseg, a.eph_t0, a.eph_dt, t, ex, ey, ez);
        if (a.eph_obs_coef != nullptr) spline_position(a.eph_obs_coef, a.eph_nseg, a.eph_t0, a.eph_dt, t, ox, oy, oz);
        else { ox = a.eph_scale * ex; oy = a.eph_scale * ey; oz = a.eph_scale * ez; }
        return;
    }
    const int64_t jo = a.obs_per_sample ? jj : 0;
    ox = a.obs[jo]; oy = a.obs[a.obs_stride + jo]; oz = a.obs[2 * a.obs_stride + jo];
    ex = 0.0; ey = 0.0;
    if (need_earth) {
        const int64_t je = a.earth_per_sample ? jj : 0;
        ex = a.earth[je];
        ey = a.earth[a.earth_stride + je];
    }
}

// Global index of local line of sight j under the (optional) block-cyclic layout.
__device__ __forceinline__ int64_t global_index(const LaunchArgs& a, int64_t j) {
    if (a.cyc_block <= 0) return j;
    if (a.cyc_shift >= 0) {
        const int64_t lb = j >> a.cyc_shift;
        return (((lb * a.cyc_parts + a.cyc_rank)) << a.cyc_shift) + (j & (a.cyc_block - 1));
    }
    const int64_t lb = j / a.cyc_block;
    return (lb * a.cyc_parts + a.cyc_rank) * a.cyc_block + (j - lb * a.cyc_block);
}

// Direction of line of sight j: loaded (reference array seam) or generated from the pixel index.
__device__ __forceinline__ void load_direction(const LaunchArgs& a, int64_t jj, double& ux, double& uy,
                                               double& uz) {
    if (a.hp_nside <= 0 && a.lon == nullptr) {
        ux = a.u[jj]; uy = a.u[a.u_stride + jj]; uz = a.u[2 * a.u_stride + jj];
        return;
    }
    double x, y, z;
    if (a.hp_nside > 0) {
        long long ipix = a.hp_start + global_index(a, jj);
        if (a.hp_nest) ipix = healpix_nest2ring(a.hp_nside, ipix);
        healpix_ring_pix2vec(a.hp_nside, ipix, x, y, z);
    } else {
        // UnitSphericalRepresentation -> cartesian, as SkyCoord.cartesian does on the host
        double sl, cl, sb, cb;
        sincos(a.lon[jj], &sl, &cl);
        sincos(a.lat[jj], &sb, &cb);
        x = cb * cl; y = cb * sl; z = sb;
    }
    if (a.hp_rotate) {
        ux = a.hp_rot[0] * x + a.hp_rot[1] * y + a.hp_rot[2] * z;
        uy = a.hp_rot[3] * x + a.hp_rot[4] * y + a.hp_rot[5] * z;
        uz = a.hp_rot[6] * x + a.hp_rot[7] * y + a.hp_rot[8] * z;
    } else {
        ux = x; uy = y; uz = z;
    }
}

// Store one result: row `ci` (component, or 0 for the summed map), column j of this call.
// With peers: the same element goes to every GPU's full map over NVLink (P2P stores, coalesced
// across the warp), which is the all-gather of the reference's `np.concatenate` done in the
// epilogue of the compute kernel.
template <typename Real>
__device__ __forceinline__ void store_out(const LaunchArgs& a, int ci, int64_t j, Real v) {
    if (a.n_peers > 0) {
        const int64_t idx = (int64_t)ci * a.peer_stride + a.peer_offset + global_index(a, j);
        for (int p = 0; p < a.n_peers; ++p) {
            if (a.out_f32) reinterpret_cast<float*>(a.peer_out[p])[idx] = (float)v;
            else reinterpret_cast<double*>(a.peer_out[p])[idx] = (double)v;
        }
        return;
    }
    const int64_t idx = (int64_t)ci * a.out_stride + j;
    if (a.out_f32) reinterpret_cast<float*>(a.out)[idx] = (float)v;
    else reinterpret_cast<double*>(a.out)[idx] = (double)v;
}

template <typename Real, int L>
__device__ __forceinline__ Real lane_group_sum(Real v) {
#pragma unroll
    for (int off = L / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Generic kernel: any component list (all reference models incl. RRM and user-edited ones).
template <typename Real, int L>
__global__ void __launch_bounds__(kThreads)
zodi_los_generic_kernel(const __grid_constant__ DevModel<Real> model,
                        const __grid_constant__ LaunchArgs args,
                        const Pair<Real>* __restrict__ g_table,
                        const Pair<Real>* __restrict__ g_nodes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Pair<Real>* s_table = reinterpret_cast<Pair<Real>*>(smem_raw);
    Pair<Real>* s_nodes = s_table + model.n_temps;
    for (int i = threadIdx.x; i < model.n_temps; i += blockDim.x) s_table[i] = g_table[i];
    for (int i = threadIdx.x; i < model.n_nodes; i += blockDim.x) s_nodes[i] = g_nodes[i];
    if (sizeof(Real) == sizeof(double)) fp64_tables_stage();
    __syncthreads();

    constexpr int kLosPerCta = kThreads / L;
    const int sub = threadIdx.x % L;
    const int64_t j = (int64_t)blockIdx.x * kLosPerCta + threadIdx.x / L;
    const bool active = j < args.n;
    const int64_t jj = active ? j : args.n - 1;  // keep whole warps converged for the shuffles

    double ux, uy, uz;
    load_direction(args, jj, ux, uy, uz);
    double ox, oy, oz, ex, ey;
    load_positions(args, jj, model.has_feature != 0, ox, oy, oz, ex, ey);

    Real total = Real(0);
    integrate_line_of_sight<Real>(
        model, s_table, s_nodes, ux, uy, uz, ox, oy, oz, ex, ey, args.outside_mask, sub, L,
        [&](int ci, Real part) {
            const Real v = lane_group_sum<Real, L>(part);
            total += v;  // component order = model order (emission.sum(axis=0), model.py:203)
            if (args.return_comps && active && sub == 0)
                store_out<Real>(args, ci, j, v);
        });
    if (!args.return_comps && active && sub == 0) store_out<Real>(args, 0, j, total);
}

// Fused Kelsall-family kernel (zodi_kelsall.cuh): cloud + 3 bands on one shared grid, then ring
// and feature.  Same thread mapping and staging as the generic kernel.
// Static shared capacity of the fused kernel (the reference uses 100 knots and 50 nodes); larger
// tables / rules take the generic kernel, whose staging area is sized dynamically.
constexpr int kFastMaxTemps = 128;
// multi-band kernels keep one table per band in (static, <= 48 KB) shared memory next to the 16 KB of
// fp64 log2 / exp2 tables: 16 bands x 104 knots x 16 B = 26 KB.  The reference's table has 100 knots.
constexpr int kMultiBandMaxTemps = 104;
constexpr int kFastMaxNodes = 128;

template <typename Real, bool HAS_RF, bool SCATTER, bool SHARE13, int L>
__global__ void __launch_bounds__(kThreads)
zodi_los_kelsall_kernel(const __grid_constant__ KelsallModel<Real> model,
                        const __grid_constant__ LaunchArgs args,
                        const Pair<Real>* __restrict__ g_table,
                        const Pair<Real>* __restrict__ g_nodes) {
    // static arrays: their shared-window addresses are compile-time constants, which keeps the
    // per-node address arithmetic out of the (issue-bound) hot loop
    __shared__ Pair<Real> s_table[kFastMaxTemps];
    __shared__ Pair<Real> s_nodes[kFastMaxNodes];
    for (int i = threadIdx.x; i < model.n_temps; i += blockDim.x) s_table[i] = g_table[i];
    for (int i = threadIdx.x; i < model.n_nodes; i += blockDim.x) s_nodes[i] = g_nodes[i];
    if (sizeof(Real) == sizeof(double)) fp64_tables_stage();
    __syncthreads();

    constexpr int kLosPerCta = kThreads / L;
    const int sub = threadIdx.x % L;
    const int64_t j = (int64_t)blockIdx.x * kLosPerCta + threadIdx.x / L;
    const bool active = j < args.n;
    const int64_t jj = active ? j : args.n - 1;

    double ux, uy, uz;
    load_direction(args, jj, ux, uy, uz);
    double ox, oy, oz, ex, ey;
    load_positions(args, jj, HAS_RF, ox, oy, oz, ex, ey);

    Real total = Real(0);
    integrate_kelsall<Real, HAS_RF, SCATTER, SHARE13>(
        model, s_table, s_nodes, ux, uy, uz, ox, oy, oz, ex, ey, args.outside_mask, sub, L,
        [&](int ci, Real part) {
            const Real v = lane_group_sum<Real, L>(part);
            total += v;
            if (args.return_comps && active && sub == 0)
                store_out<Real>(args, ci, j, v);
        });
    if (!args.return_comps && active && sub == 0) store_out<Real>(args, 0, j, total);
}

// Packed-fp32 fused kernel: every thread works on TWO lines of sight with FFMA2/FMUL2/FADD2
// (zodi_kelsall_x2.cuh): cloud + bands, then ring | ring and feature | feature.  L lanes share a pair of
// lines of sight and split its quadrature nodes (L = 1 for large N - the benchmarked path; 2 / 4 / 8 fill
// the machine for the map sizes the reference documents, nside 64-256).  A CTA of THREADS threads covers
// 2 * THREADS / L lines of sight: pair group g takes lines g and g + THREADS / L of the CTA's block
// (coalesced loads and stores).  With L >= 2 the fp64 prologue of a line of sight is computed once (even
// lanes: first line of the pair, odd lanes: second) and exchanged by warp shuffles.
template <int L>
__device__ __forceinline__ void exchange_pre(const LosPre& mine, int sub, LosPre& P0, LosPre& P1) {
    const float* m = reinterpret_cast<const float*>(&mine);
    float* p0 = reinterpret_cast<float*>(&P0);
    float* p1 = reinterpret_cast<float*>(&P1);
    const int base = (threadIdx.x & 31) - sub;  // first lane of this pair group
#pragma unroll
    for (int i = 0; i < kLosPreFloats; ++i) {
        p0[i] = __shfl_sync(0xffffffffu, m[i], base);
        p1[i] = __shfl_sync(0xffffffffu, m[i], base + 1);
    }
}

template <bool HAS_RF, bool SHARE13, bool SCATTER, int L, int THREADS, int MIN_CTAS, bool PERSIST = false>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
zodi_los_kelsall_x2_kernel(const __grid_constant__ KelsallModel<float> model,
                           const __grid_constant__ LaunchArgs args,
                           const Pair<float>* __restrict__ g_table,
                           const Pair<float>* __restrict__ g_nodes) {
    __shared__ Pair<float> s_table[kFastMaxTemps];
    __shared__ Pair<float> s_nodes[kFastMaxNodes];
    for (int i = threadIdx.x; i < model.n_temps; i += blockDim.x) s_table[i] = g_table[i];
    for (int i = threadIdx.x; i < model.n_nodes; i += blockDim.x) s_nodes[i] = g_nodes[i];
    __syncthreads();

    constexpr int kGroups = THREADS / L;  // pair groups per CTA
    const int sub = threadIdx.x % L;
    // A tile = the 2 * kGroups lines of sight one CTA pass covers.  Default: tile = CTA.  PERSIST (tile counter
    // given): the grid only fills the machine and every CTA claims further tiles as it finishes one: a CTA's slot is
    // not released until its stores have been acknowledged, which for peer (NVLink) stores costs 1-2 us per
    // CTA - 3-4 % of a 45 us tile at 8 GPUs; a persistent CTA pays that once, and its remote stores of one
    // tile drain behind the next tile's arithmetic.
    __shared__ unsigned int s_claim;
    const unsigned int n_tiles = (unsigned int)((args.n + 2 * kGroups - 1) / (2 * kGroups));
    for (unsigned int tile = blockIdx.x;;) {
    const int64_t j0 = (int64_t)tile * (2 * kGroups) + threadIdx.x / L, j1 = j0 + kGroups;
    const bool act0 = j0 < args.n, act1 = j1 < args.n;
    const int64_t jj0 = act0 ? j0 : args.n - 1, jj1 = act1 ? j1 : args.n - 1;

    auto prologue = [&](int64_t jj) {
        double ux, uy, uz, ox, oy, oz, ex, ey;
        load_direction(args, jj, ux, uy, uz);
        load_positions(args, jj, HAS_RF, ox, oy, oz, ex, ey);
        return los_pre<HAS_RF>(model, ux, uy, uz, ox, oy, oz, ex, ey, args.outside_mask);
    };
    LosPre P0, P1;
    if (L == 1) {
        P0 = prologue(jj0);
        P1 = prologue(jj1);
    } else {
        const LosPre mine = prologue((sub & 1) ? jj1 : jj0);
        exchange_pre<L>(mine, sub, P0, P1);
    }

    float tot0 = 0.f, tot1 = 0.f;
    auto emit2 = [&](int ci, float va, float vb) {
        va = lane_group_sum<float, L>(va);
        vb = lane_group_sum<float, L>(vb);
        tot0 += va;
        tot1 += vb;
        if (args.return_comps && sub == 0) {
            if (act0) store_out<float>(args, ci, j0, va);
            if (act1) store_out<float>(args, ci, j1, vb);
        }
    };
#if ZODI_X2_RF_FIRST
    // ring / feature first: during the cloud + bands loop only their four results stay live instead of the
    // twelve interval / rotation values; components are still emitted (and summed) in model order
    float ra = 0.f, rb = 0.f, fa = 0.f, fb = 0.f;
    if (HAS_RF) {
        kelsall_ring_x2<SCATTER, L>(model, s_table, s_nodes, P0, P1, sub, [&](float a, float b) { ra = a; rb = b; });
        kelsall_feature_x2<SCATTER, L>(model, s_table, s_nodes, P0, P1, sub, [&](float a, float b) { fa = a; fb = b; });
    }
    kelsall_group_a_x2<SHARE13, SCATTER, L>(model, s_table, s_nodes, P0, P1, sub, emit2);
    if (HAS_RF) {
        emit2(4, ra, rb);
        emit2(5, fa, fb);
    }
#else
    kelsall_group_a_x2<SHARE13, SCATTER, L>(model, s_table, s_nodes, P0, P1, sub, emit2);
    if (HAS_RF) {
        kelsall_ring_x2<SCATTER, L>(model, s_table, s_nodes, P0, P1, sub, [&](float a, float b) { emit2(4, a, b); });
        kelsall_feature_x2<SCATTER, L>(model, s_table, s_nodes, P0, P1, sub, [&](float a, float b) { emit2(5, a, b); });
    }
#endif
    if (!args.return_comps && sub == 0) {
        if (act0) store_out<float>(args, 0, j0, tot0);
        if (act1) store_out<float>(args, 0, j1, tot1);
    }
    if (!PERSIST) break;
    __syncthreads();  // every thread has read the previous claim
    if (threadIdx.x == 0) s_claim = atomicAdd(args.tile_counter, 1u);
    __syncthreads();
    tile = s_claim + gridDim.x;
    if (tile >= n_tiles) break;
    }
    if (PERSIST && threadIdx.x == 0) {
        // all claims of this CTA precede its `done` increment, so the CTA that completes the count may reset
        if (atomicAdd(args.tile_counter + 1, 1u) == gridDim.x - 1) {
            args.tile_counter[0] = 0u;
            args.tile_counter[1] = 0u;
        }
    }
}

// Fused RRM kernel (zodi_rrm.cuh): six node loops for the eight components of the shipped layout.
// Same thread mapping and staging as the scalar Kelsall kernel.
template <typename Real, int L>
__global__ void __launch_bounds__(kThreads)
zodi_los_rrm_kernel(const __grid_constant__ RrmModel<Real> model,
                    const __grid_constant__ LaunchArgs args,
                    const Pair<Real>* __restrict__ g_table,
                    const Pair<Real>* __restrict__ g_nodes) {
    __shared__ Pair<Real> s_table[kFastMaxTemps];
    __shared__ Pair<Real> s_nodes[kFastMaxNodes];
    for (int i = threadIdx.x; i < model.n_temps; i += blockDim.x) s_table[i] = g_table[i];
    for (int i = threadIdx.x; i < model.n_nodes; i += blockDim.x) s_nodes[i] = g_nodes[i];
    if (sizeof(Real) == sizeof(double)) fp64_tables_stage();
    __syncthreads();

    constexpr int kLosPerCta = kThreads / L;
    const int sub = threadIdx.x % L;
    const int64_t j = (int64_t)blockIdx.x * kLosPerCta + threadIdx.x / L;
    const bool active = j < args.n;
    const int64_t jj = active ? j : args.n - 1;

    double ux, uy, uz;
    load_direction(args, jj, ux, uy, uz);
    double ox, oy, oz, ex, ey;
    load_positions(args, jj, true, ox, oy, oz, ex, ey);

    Real total = Real(0);
    integrate_rrm<Real>(
        model, s_table, s_nodes, ux, uy, uz, ox, oy, oz, ex, ey, args.outside_mask, sub, L,
        [&](int ci, Real part) {
            const Real v = lane_group_sum<Real, L>(part);
            total += v;
            if (args.return_comps && active && sub == 0)
                store_out<Real>(args, ci, j, v);
        });
    if (!args.return_comps && active && sub == 0) store_out<Real>(args, 0, j, total);
}

// Packed-fp32 fused RRM kernel (zodi_rrm_x2.cuh): two lines of sight per thread (j and j + THREADS of the
// CTA's block), large-N fp32 path of the RRM model.
template <int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
zodi_los_rrm_x2_kernel(const __grid_constant__ RrmModelX2 model,
                       const __grid_constant__ LaunchArgs args,
                       const Pair<float>* __restrict__ g_table,
                       const Pair<float>* __restrict__ g_nodes) {
    __shared__ Pair<float> s_table[kFastMaxTemps];
    __shared__ Pair<float> s_nodes[kFastMaxNodes];
    for (int i = threadIdx.x; i < model.r.n_temps; i += blockDim.x) s_table[i] = g_table[i];
    for (int i = threadIdx.x; i < model.r.n_nodes; i += blockDim.x) s_nodes[i] = g_nodes[i];
    __syncthreads();

    const int64_t j0 = (int64_t)blockIdx.x * (2 * THREADS) + threadIdx.x, j1 = j0 + THREADS;
    const bool act0 = j0 < args.n, act1 = j1 < args.n;
    const int64_t jj0 = act0 ? j0 : args.n - 1, jj1 = act1 ? j1 : args.n - 1;
    LosPre P[2];
    RrmIntervals I[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int64_t jj = q ? jj1 : jj0;
        double ux, uy, uz, ox, oy, oz, ex, ey;
        load_direction(args, jj, ux, uy, uz);
        load_positions(args, jj, true, ox, oy, oz, ex, ey);
        rrm_pre(model, ux, uy, uz, ox, oy, oz, ex, ey, args.outside_mask, P[q], I[q]);
    }
    float tot0 = 0.f, tot1 = 0.f;
    integrate_rrm_x2(model, s_table, s_nodes, P[0], P[1], I[0], I[1], [&](int ci, float va, float vb) {
        tot0 += va;
        tot1 += vb;
        if (args.return_comps) {
            if (act0) store_out<float>(args, ci, j0, va);
            if (act1) store_out<float>(args, ci, j1, vb);
        }
    });
    if (!args.return_comps) {
        if (act0) store_out<float>(args, 0, j0, tot0);
        if (act1) store_out<float>(args, 0, j1, tot1);
    }
}

// Multi-band kernel (zodi_multiband.cuh): NB bands of one Kelsall-family model per pass; output
// row b of `out` (row stride out_stride) is band b's component-summed emission.
template <typename Real, int NB, bool HAS_RF, bool SCATTER>
__global__ void __launch_bounds__(kThreads)
zodi_los_multiband_kernel(const __grid_constant__ MultiBandModel<Real> model,
                          const __grid_constant__ LaunchArgs args,
                          const Pair<Real>* __restrict__ g_tables,   // [n_bands][n_temps]
                          const Pair<Real>* __restrict__ g_nodes) {
    __shared__ Pair<Real> s_tables[NB * kMultiBandMaxTemps];
    __shared__ Pair<Real> s_nodes[kFastMaxNodes];
    const int nt = model.base.n_temps;
    for (int i = threadIdx.x; i < NB * nt; i += blockDim.x) {
        const Pair<Real> zero = {Real(0), Real(0)};
        s_tables[i] = (i < model.n_bands * nt) ? g_tables[i] : zero;  // padded bands: zero source
    }
    for (int i = threadIdx.x; i < model.base.n_nodes; i += blockDim.x) s_nodes[i] = g_nodes[i];
    if (sizeof(Real) == sizeof(double)) fp64_tables_stage();
    __syncthreads();

    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool active = j < args.n;
    const int64_t jj = active ? j : args.n - 1;
    double ux, uy, uz, ox, oy, oz, ex, ey;
    load_direction(args, jj, ux, uy, uz);
    load_positions(args, jj, HAS_RF, ox, oy, oz, ex, ey);
    integrate_kelsall_multiband<Real, NB, HAS_RF, SCATTER>(
        model, s_tables, s_nodes, ux, uy, uz, ox, oy, oz, ex, ey, args.outside_mask, 0, 1,
        [&](int b, Real v) {
            if (active && b < model.n_bands) store_out<Real>(args, b, j, v);
        });
}

// Packed-fp32 multi-band kernel (zodi_multiband_x2.cuh): two lines of sight per thread (lines t and
// t + THREADS of the CTA's block), tables staged knot-major.  Resident 128-thread CTAs per SM: 8 (64
// registers) up to 8 bands, 6 (80 registers) for 16 - the node loop is a long dependent chain, so warps in
// flight count for more than the handful of spilled prologue values (4 CTAs: 74 - 110 registers, no spills).
#ifndef ZODI_MBX2_CTAS
#define ZODI_MBX2_CTAS 8
#endif
#ifndef ZODI_MBX2_CTAS_16
#define ZODI_MBX2_CTAS_16 6
#endif
template <int NB, bool HAS_RF, bool SCATTER>
__global__ void __launch_bounds__(kPackedDefaultThreads, (NB <= 8 ? ZODI_MBX2_CTAS : ZODI_MBX2_CTAS_16))
zodi_los_multiband_x2_kernel(const __grid_constant__ MultiBandModel<float> model,
                             const __grid_constant__ LaunchArgs args,
                             const Pair<float>* __restrict__ g_rows,     // [n_temps][NB + 2], knot-major
                             const Pair<float>* __restrict__ g_nodes) {
    constexpr int kRow = MbRows<NB>::kRow;
    __shared__ __align__(16) Pair<float> s_rows[kMultiBandMaxTemps * kRow];
    __shared__ Pair<float> s_nodes[kFastMaxNodes];
    {   // rows are staged as they lie in HBM (the host transposed them): 16-byte copies, kRow is even
        const float4* src = reinterpret_cast<const float4*>(g_rows);
        float4* dst = reinterpret_cast<float4*>(s_rows);
        const int n16 = model.base.n_temps * (kRow / 2);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i < model.base.n_nodes; i += blockDim.x) s_nodes[i] = g_nodes[i];
    __syncthreads();

    constexpr int T = kPackedDefaultThreads;
    const int64_t j0 = (int64_t)blockIdx.x * (2 * T) + threadIdx.x, j1 = j0 + T;
    const bool act0 = j0 < args.n, act1 = j1 < args.n;
    const int64_t jj0 = act0 ? j0 : args.n - 1, jj1 = act1 ? j1 : args.n - 1;
    auto prologue = [&](int64_t jj) {
        double ux, uy, uz, ox, oy, oz, ex, ey;
        load_direction(args, jj, ux, uy, uz);
        load_positions(args, jj, HAS_RF, ox, oy, oz, ex, ey);
        return los_pre<HAS_RF>(model.base, ux, uy, uz, ox, oy, oz, ex, ey, args.outside_mask);
    };
    const LosPre P0 = prologue(jj0), P1 = prologue(jj1);
    integrate_multiband_x2<NB, HAS_RF, SCATTER>(model, s_rows, s_nodes, P0, P1, [&](int b, float va, float vb) {
        if (b < model.n_bands) {
            if (act0) store_out<float>(args, b, j0, va);
            if (act1) store_out<float>(args, b, j1, vb);
        }
    });
}

}  // namespace zodi
