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

namespace zodi {

constexpr int kThreads = 256;

// Launch-time arguments (plain pointers; device memory).
struct LaunchArgs {
    int64_t n;
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
    // on-device ephemeris (time-ordered data): positions from cubic splines at obstime[j]
    const double* eph_coef;   // [n_knots-1][3][4] Earth (highest power first); NULL = use obs/earth arrays
    const double* eph_obs_coef;  // same for the observer, or NULL: observer = eph_scale * Earth
    const double* obstime;
    int64_t eph_nseg;     double eph_t0, eph_dt, eph_scale;
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

// Packed-fp32 fused kernel: every thread integrates TWO lines of sight (j and j + 256 of its CTA's
// 512) with FFMA2/FMUL2/FADD2 for the cloud + bands group (zodi_kelsall_x2.cuh); ring and feature
// reuse the scalar routines.  fp32, thermal-only, large-N (L = 1) case = the benchmarked path.
template <bool HAS_RF, bool SHARE13, bool SCATTER, int MIN_CTAS>
__global__ void __launch_bounds__(kThreads, MIN_CTAS)
zodi_los_kelsall_x2_kernel(const __grid_constant__ KelsallModel<float> model,
                           const __grid_constant__ LaunchArgs args,
                           const Pair<float>* __restrict__ g_table,
                           const Pair<float>* __restrict__ g_nodes) {
    __shared__ Pair<float> s_table[kFastMaxTemps];
    __shared__ Pair<float> s_nodes[kFastMaxNodes];
    for (int i = threadIdx.x; i < model.n_temps; i += blockDim.x) s_table[i] = g_table[i];
    for (int i = threadIdx.x; i < model.n_nodes; i += blockDim.x) s_nodes[i] = g_nodes[i];
    __syncthreads();

    const int64_t j0 = (int64_t)blockIdx.x * (2 * kThreads) + threadIdx.x, j1 = j0 + kThreads;
    const bool act0 = j0 < args.n, act1 = j1 < args.n;
    const int64_t jj0 = act0 ? j0 : args.n - 1, jj1 = act1 ? j1 : args.n - 1;

    LosGeometry<float> G[2];
    double ex[2] = {0.0, 0.0}, ey[2] = {0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int64_t jj = q ? jj1 : jj0;
        double ux, uy, uz;
        load_direction(args, jj, ux, uy, uz);
        double ox, oy, oz;
        load_positions(args, jj, HAS_RF, ox, oy, oz, ex[q], ey[q]);
        G[q] = los_geometry<float>(ux, uy, uz, ox, oy, oz);
    }

    float tot0 = 0.f, tot1 = 0.f;
    auto emit2 = [&](int ci, float va, float vb) {
        tot0 += va;
        tot1 += vb;
        if (args.return_comps) {
            if (act0) store_out<float>(args, ci, j0, va);
            if (act1) store_out<float>(args, ci, j1, vb);
        }
    };
    kelsall_group_a_x2<SHARE13, SCATTER>(model, s_table, s_nodes, G[0], G[1], args.outside_mask, emit2);
    if (HAS_RF) {
        float ring[2], feat[2];
#pragma unroll
        for (int q = 0; q < 2; ++q)
            kelsall_ring_feature_packed<SCATTER>(model, s_table, s_nodes, G[q], ex[q], ey[q], args.outside_mask,
                                        [&](float r, float f) { ring[q] = r; feat[q] = f; });
        emit2(4, ring[0], ring[1]);
        emit2(5, feat[0], feat[1]);
    }
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

// Generated directions only (same device routine the integrators use in their prologue).
__global__ void zodi_healpix_vectors_kernel(const __grid_constant__ LaunchArgs args, double* __restrict__ out,
                                            int64_t out_stride) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= args.n) return;
    double ux, uy, uz;
    load_direction(args, j, ux, uy, uz);
    out[j] = ux; out[out_stride + j] = uy; out[2 * out_stride + j] = uz;
}

// Number density of every component at n points (grid_number_density, number_density.py:482-536).
__global__ void zodi_number_density_kernel(const __grid_constant__ DevModel<double> model,
                                           const double* __restrict__ xyz, int64_t n, int64_t stride,
                                           double ex, double ey, double* __restrict__ out, int64_t out_stride) {
    fp64_tables_stage();
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double x = xyz[j], y = xyz[stride + j], z = xyz[2 * stride + j];
    for (int ci = 0; ci < model.n_comps; ++ci) {
        const DevComp<double>& c = model.comps[ci];
        double theta_earth = 0.0;
        if (c.type == D_FEATURE) theta_earth = atan2(ey - c.y0, ex - c.x0);
        out[(int64_t)ci * out_stride + j] = density<double>(c, x - c.x0, y - c.y0, z - c.z0, theta_earth);
    }
}

// Element-wise device math for tests (zodi_device_math): the routines the integrators are built from,
// run on the GPU itself (MUFU seeds, shared-memory tables) rather than in the host emulation.
__global__ void zodi_device_math_kernel(int op, int64_t n, const double* __restrict__ x, double aux,
                                        double* __restrict__ y) {
    fp64_tables_stage();
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        double r;
        switch (op) {
            case ZODI_MATH_LOG2_F64: r = Math<double>::log2_(v); break;
            case ZODI_MATH_EXP2_F64: r = Math<double>::exp2_(v); break;
            case ZODI_MATH_RSQRT_F64: r = Math<double>::rsqrt_(v); break;
            case ZODI_MATH_ATAN2_ABS_F64: r = Math<double>::atan2_abs_(v, aux); break;
            case ZODI_MATH_ASIN_F32: r = (double)asin_unit((float)v); break;
            case ZODI_MATH_ATAN2_ABS_F32: r = (double)Math<float>::atan2_abs_((float)v, (float)aux); break;
            case ZODI_MATH_ONE_MINUS_EXP2_NEG_F32: r = (double)Math<float>::one_minus_exp2_neg((float)v); break;
            case ZODI_MATH_EXP2_F32: r = (double)Math<float>::exp2_((float)v); break;
            case ZODI_MATH_LOG2_F32: r = (double)Math<float>::log2_((float)v); break;
            default: r = 0.0;
        }
        y[i] = r;
    }
}

// Spline positions at n times (tests / users): earth_out, obs_out (3, n) or NULL.
__global__ void zodi_ephemeris_positions_kernel(const __grid_constant__ LaunchArgs a, double* __restrict__ earth_out,
                                                double* __restrict__ obs_out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n) return;
    const double t = a.obstime[j];
    double ex, ey, ez, ox, oy, oz;
    spline_position(a.eph_coef, a.eph_nseg, a.eph_t0, a.eph_dt, t, ex, ey, ez);
    if (a.eph_obs_coef != nullptr) spline_position(a.eph_obs_coef, a.eph_nseg, a.eph_t0, a.eph_dt, t, ox, oy, oz);
    else { ox = a.eph_scale * ex; oy = a.eph_scale * ey; oz = a.eph_scale * ez; }
    if (earth_out) { earth_out[j] = ex; earth_out[a.n + j] = ey; earth_out[2 * a.n + j] = ez; }
    if (obs_out) { obs_out[j] = ox; obs_out[a.n + j] = oy; obs_out[2 * a.n + j] = oz; }
}

// Reductions over the samples: stats[0] += sum |earth|^2, stats[1] = max |earth|^2 (as bits),
// stats[2] = max |observer|^2 (as bits; only when the observer has its own spline).
__global__ void zodi_ephemeris_stats_kernel(const __grid_constant__ LaunchArgs a, double* __restrict__ sum_out,
                                            unsigned long long* __restrict__ max_bits) {
    double sum = 0.0, me = 0.0, mo = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += (int64_t)gridDim.x * blockDim.x) {
        const double t = a.obstime[j];
        double x, y, z;
        spline_position(a.eph_coef, a.eph_nseg, a.eph_t0, a.eph_dt, t, x, y, z);
        const double r2 = x * x + y * y + z * z;
        sum += r2;
        me = fmax(me, r2);
        if (a.eph_obs_coef != nullptr) {
            spline_position(a.eph_obs_coef, a.eph_nseg, a.eph_t0, a.eph_dt, t, x, y, z);
            mo = fmax(mo, x * x + y * y + z * z);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        me = fmax(me, __shfl_xor_sync(0xffffffffu, me, off));
        mo = fmax(mo, __shfl_xor_sync(0xffffffffu, mo, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sum_out, sum);
        atomicMax(max_bits, (unsigned long long)__double_as_longlong(me));
        atomicMax(max_bits + 1, (unsigned long long)__double_as_longlong(mo));
    }
}

// max over observers of r^2 = x^2+y^2+z^2 (for the global early-out flags, quirk Q1).
// Non-negative doubles order like their bit patterns, so atomicMax on the 64-bit pattern works.
__global__ void zodi_max_r2_kernel(const double* __restrict__ obs, int64_t n, int64_t stride,
                                   unsigned long long* __restrict__ out_bits) {
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double x = obs[i], y = obs[stride + i], z = obs[2 * stride + i];
        m = fmax(m, x * x + y * y + z * z);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));
}

// ---- pipe-peak microbenchmarks (roofline denominators measured on the box) -----------------
template <typename Real>
__global__ void zodi_peak_fma_kernel(Real* out, int iters) {
    Real a0 = Real(threadIdx.x) * Real(1e-3), a1 = a0 + Real(1), a2 = a0 + Real(2), a3 = a0 + Real(3);
    Real a4 = a0 + Real(4), a5 = a0 + Real(5), a6 = a0 + Real(6), a7 = a0 + Real(7);
    const Real m = Real(0.9999), c = Real(1e-4);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
    }
    const Real s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == Real(-1)) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void zodi_peak_mufu_kernel(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
    float a4 = a0 + 0.4f, a5 = a0 + 0.5f, a6 = a0 + 0.6f, a7 = a0 + 0.7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
        }
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == -1.0f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void zodi_peak_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                      int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace zodi
