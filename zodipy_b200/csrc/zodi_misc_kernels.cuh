// zodi_misc_kernels.cuh - the non-integrator kernels (directions, densities, ephemeris reductions,
// device-math test hook, pipe-peak probes).  Included by zodi_capi.cu ONLY: they are ordinary
// (non-template) __global__ functions and must be defined in exactly one translation unit.
#pragma once

#include "zodi_kernels.cuh"

namespace zodi {

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

// Completion rendezvous of the fused all-gather (zodi_peer_rendezvous): thread p publishes this rank's
// epoch to peer p and waits for peer p's epoch in the own array.  Runs behind the integrator kernel on
// the same stream, so the kernel's peer stores are complete; the system-scope fence + release order
// them before the flag for the peers' acquire loads.
struct PeerFlagPtrs { uint32_t* p[ZODI_MAX_PEERS]; };
__global__ void zodi_peer_rendezvous_kernel(PeerFlagPtrs flags, int n_peers, int rank, uint32_t epoch) {
    const int p = threadIdx.x;
    if (p >= n_peers) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[p] + rank), "r"(epoch) : "memory");
    const uint32_t* mine = flags.p[rank] + p;
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (clock64() - t0 > (20LL << 30)) {  // ~10 s: a peer died - report instead of hanging the GPU
            flags.p[rank][ZODI_MAX_PEERS] = 1u;
            break;
        }
    }
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
