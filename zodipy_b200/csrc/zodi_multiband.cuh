// zodi_multiband.cuh - several bands of one Kelsall-family model in a single pass.
//
// The reference evaluates one wavelength / bandpass per Model (zodipy/model.py:36-117); pipelines
// that need every DIRBE or Planck channel for the same pointings (e.g. the per-band loop of the
// reference's own tests/test_evaluate.py:54-66) repeat the whole line-of-sight integration per band.
// Between bands of one model only the source function changes - the blackbody table values
// (zodipy/blackbody.py:33-49) and the spectral scalars emissivity / albedo / C1-C3 / solar
// irradiance (zodipy/unpack_model.py:34-105) - while positions, grain temperature (T_0, delta are
// band-independent), the table INDEX and every number density are identical.  This routine does
// that shared work once per quadrature node and adds, per band, one table read and one small
// weighted sum.
//
// Output: the component-summed emission per band (n_bands, N).  Per-component output would need
// n_bands x n_comps accumulators per line of sight; use the single-band kernels for that.
#pragma once

#include "zodi_kelsall.cuh"

namespace zodi {

constexpr int kMaxBands = 16;

template <typename Real>
struct MultiBandModel {
    KelsallModel<Real> base;  // geometry, density constants, temperature law, cutoffs (aB/aS unused)
    int n_bands;
    int n_bands_padded;       // compile-time band count of the kernel instance (4, 8 or 16)
    Real aB[kMaxBands][6];    // (1 - albedo) * emissivity * amplitude   per band and component
    Real aS[kMaxBands][6];    // albedo * F_sun * N_phase * amplitude
    Real C1p[kMaxBands], C2p[kMaxBands], C3l[kMaxBands];  // phase function per band (C3 * log2e)
    uint32_t scatter_bands;   // bit b: band b has a non-zero albedo (packed kernel: scattering term per band)
};

// Table coordinate -> (segment index, fraction): table_coord() of zodi_device.cuh.
template <typename Real>
ZODI_HD void table_locate(Real t, Real t_top, int& idx, Real& frac) {
    table_coord<Real>(t, t_top, idx, frac);
}

// Per-node quantities shared by all bands.
template <typename Real>
struct BandNode {
    Real xh, yh, zh, Rh2;
    int idx;      // table segment
    Real frac;    // position inside the segment
    Real th;      // scattering angle (SCATTER only)
    Real rh2inv;  // 1 / R_h^2      (SCATTER only)
};

template <typename Real, bool SCATTER>
ZODI_HD BandNode<Real> band_node(const KelsallModel<Real>& K, Real R_los, const LosGeometry<Real>& G) {
    using M = Math<Real>;
    BandNode<Real> s;
    s.xh = M::fma_(R_los, G.ux, G.ox);
    s.yh = M::fma_(R_los, G.uy, G.oy);
    s.zh = M::fma_(R_los, G.uz, G.oz);
    s.Rh2 = M::fma_(s.xh, s.xh, M::fma_(s.yh, s.yh, s.zh * s.zh));
    const Real t = M::fma_(K.t_scale, M::exp2_bounded_(K.mhd * M::log2_(s.Rh2)), K.t_ofs);
    table_locate<Real>(t, K.t_top, s.idx, s.frac);
    s.th = Real(0);
    s.rh2inv = Real(0);
    if (SCATTER) {
        const Real rh_inv = M::rsqrt_(s.Rh2);
        Real ct = M::fma_(G.ux, s.xh, M::fma_(G.uy, s.yh, G.uz * s.zh)) * rh_inv;
        ct = M::max_(Real(-1), M::min_(Real(1), ct));
        s.th = M::acos_(-ct);
        s.rh2inv = rh_inv * rh_inv;
    }
    return s;
}

// acc[b] += w * (B_b * sum_c aB[b][c] n_c  +  F_b * sum_c aS[b][c] n_c) for components c0 .. c0+NC-1
template <typename Real, int NB, int NC, bool SCATTER>
ZODI_HD void band_accumulate_all(const MultiBandModel<Real>& MB, const Pair<Real>* tabs, int n_temps,
                                 const BandNode<Real>& s, Real w, const Real (&n)[NC], int c0, Real (&acc)[NB]) {
    using M = Math<Real>;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < NB; ++b) {
        const Pair<Real> e = tabs[b * n_temps + s.idx];
        const Real B = M::fma_(e.b, s.frac, e.a);  // brightness.py:48 for band b
        Real sB = MB.aB[b][c0] * n[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 1; c < NC; ++c) sB = M::fma_(MB.aB[b][c0 + c], n[c], sB);
        Real src = B * sB;
        if (SCATTER) {
            Real sS = MB.aS[b][c0] * n[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int c = 1; c < NC; ++c) sS = M::fma_(MB.aS[b][c0 + c], n[c], sS);
            const Real F = (MB.C1p[b] + MB.C2p[b] * s.th + M::exp2_(MB.C3l[b] * s.th)) * s.rh2inv;
            src = M::fma_(F, sS, src);
        }
        acc[b] = M::fma_(w, src, acc[b]);
    }
}

// One line of sight, NB bands.  emit(b, partial) receives band b's emission summed over components
// (partial over this lane's nodes; the caller reduces over the L lanes).
template <typename Real, int NB, bool HAS_RF, bool SCATTER, typename Emit>
ZODI_HD void integrate_kelsall_multiband(const MultiBandModel<Real>& MB, const Pair<Real>* tabs,
                                         const Pair<Real>* nodes, double dux, double duy, double duz,
                                         double dox, double doy, double doz, double dex, double dey,
                                         uint32_t outside_mask, int sub, int L, Emit emit) {
    using M = Math<Real>;
    const KelsallModel<Real>& K = MB.base;
    const LosGeometry<Real> G = los_geometry<Real>(dux, duy, duz, dox, doy, doz);
    Real total[NB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < NB; ++b) total[b] = Real(0);

    // ---- cloud + band1..3 on one grid ----
    {
        Real h, mid;
        los_interval<Real>(G, K.cutA_in, K.cutA_out, outside_mask & 1u, (outside_mask >> 1) & 1u, h, mid);
        Real acc[NB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int b = 0; b < NB; ++b) acc[b] = Real(0);
        for (int k0 = 0; k0 < K.n_nodes; k0 += L) {  // warp-uniform trip count (votes inside)
            const int k = k0 + sub;
            Pair<Real> nw = nodes[k < K.n_nodes ? k : K.n_nodes - 1];
            if (k >= K.n_nodes) nw.b = Real(0);
            const BandNode<Real> s = band_node<Real, SCATTER>(K, M::fma_(h, nw.a, mid), G);
            Real n[4] = {Real(0), Real(0), Real(0), Real(0)};
            const Real rinv = M::rsqrt_(s.Rh2);
            const Real rad1 = band_radial<Real>(s.Rh2, K.b_y[0]);
            const Real rad2 = band_radial<Real>(s.Rh2, K.b_y[1]);
            const Real rad3 = K.share13 ? rad1 : band_radial<Real>(s.Rh2, K.b_y[2]);
            Real unused = Real(0);
            // band_accumulate with unit weight leaves the (possibly skipped) density in n[.]
            band_accumulate<Real, false>(n[1], unused, Real(1), Real(0), s.xh, s.yh, s.zh, rinv, rinv * rad1, K.bnx[0], K.bny[0], K.bnz[0], K.b_c3[0]);
            band_accumulate<Real, false>(n[2], unused, Real(1), Real(0), s.xh, s.yh, s.zh, rinv, rinv * rad2, K.bnx[1], K.bny[1], K.bnz[1], K.b_c3[1]);
            band_accumulate<Real, false>(n[3], unused, Real(1), Real(0), s.xh, s.yh, s.zh, rinv, rinv * rad3, K.bnx[2], K.bny[2], K.bnz[2], K.b_c3[2]);
            const Real xc = s.xh - K.cx0, yc = s.yh - K.cy0, zc = s.zh - K.cz0;
            const Real Rc2 = M::fma_(xc, xc, M::fma_(yc, yc, zc * zc));
            const Real zeta = M::abs_(M::fma_(xc, K.cnx, M::fma_(yc, K.cny, zc * K.cnz))) * M::rsqrt_(Rc2);
            const Real g = (zeta < K.c_mu) ? zeta * zeta * K.c_inv2mu : zeta - K.c_halfmu;
            const Real gp = M::exp2_(K.c_gamma * M::log2_(g));
            n[0] = M::exp2_(M::fma_(K.c_mha, M::log2_(Rc2), K.c_mbl * gp));
            band_accumulate_all<Real, NB, 4, SCATTER>(MB, tabs, K.n_temps, s, nw.b, n, 0, acc);
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int b = 0; b < NB; ++b) total[b] = h * acc[b];
    }
    if (HAS_RF) {
        // ---- ring (own grid) ----
        {
            Real h, mid;
            los_interval<Real>(G, K.cutR_in, K.cutR_out, (outside_mask >> 8) & 1u, (outside_mask >> 9) & 1u, h, mid);
            Real acc[NB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int b = 0; b < NB; ++b) acc[b] = Real(0);
            for (int k = sub; k < K.n_nodes; k += L) {
                const Pair<Real> nw = nodes[k];
                const BandNode<Real> s = band_node<Real, SCATTER>(K, M::fma_(h, nw.a, mid), G);
                const Real d = M::sqrt_(s.Rh2) - K.r_R;
                const Real Zc = M::fma_(s.xh, K.rnx, M::fma_(s.yh, K.rny, s.zh * K.rnz));
                const Real n[1] = {M::exp2_neg_(-M::fma_(d * d, K.r_c2, M::abs_(Zc) * K.r_c3))};
                band_accumulate_all<Real, NB, 1, SCATTER>(MB, tabs, K.n_temps, s, nw.b, n, 4, acc);
            }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int b = 0; b < NB; ++b) total[b] = M::fma_(h, acc[b], total[b]);
        }
        // ---- feature (own grid) ----
        {
            Real h, mid;
            los_interval<Real>(G, K.cutF_in, K.cutF_out, (outside_mask >> 10) & 1u, (outside_mask >> 11) & 1u, h, mid);
            Real cr, sr;
            feature_rotation<Real>(dex, dey, K.f_cos0, K.f_sin0, cr, sr);
            Real acc[NB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int b = 0; b < NB; ++b) acc[b] = Real(0);
            for (int k = sub; k < K.n_nodes; k += L) {
                const Pair<Real> nw = nodes[k];
                const BandNode<Real> s = band_node<Real, SCATTER>(K, M::fma_(h, nw.a, mid), G);
                const Real d = M::sqrt_(s.Rh2) - K.f_R;
                const Real Zc = M::fma_(s.xh, K.fnx, M::fma_(s.yh, K.fny, s.zh * K.fnz));
                const Real xr = M::fma_(s.xh, cr, s.yh * sr), yr = M::fma_(s.yh, cr, -(s.xh * sr));
                const Real dth = M::atan2_abs_(yr, xr);
                const Real e = M::fma_(d * d, K.f_c2, M::fma_(M::abs_(Zc), K.f_c3, dth * dth * K.f_c5));
                const Real n[1] = {M::exp2_neg_(-e)};
                band_accumulate_all<Real, NB, 1, SCATTER>(MB, tabs, K.n_temps, s, nw.b, n, 5, acc);
            }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int b = 0; b < NB; ++b) total[b] = M::fma_(h, acc[b], total[b]);
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < NB; ++b) emit(b, total[b]);
}

}  // namespace zodi
