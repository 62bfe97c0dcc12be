// zodi_kelsall.cuh - fused integrator specialised for the Kelsall model family.
//
// All five Kelsall-type models the reference ships (dirbe, planck13, planck15, planck18, odegard;
// zodipy/model_registry.py:4-62) have the component list
//     cloud, band1, band2, band3 [, ring, feature]
// where cloud and the three bands share one line-of-sight range (cutoffs (eps, 5.2 AU),
// zodipy/line_of_sight.py:19-26), the bands are centred on the Sun (X_0 = 0,
// zodipy/component_params.py:32-67) and have p = 4.  For that layout the work per quadrature node
// that the reference repeats for every component (position, heliocentric distance, grain
// temperature, blackbody interpolation; zodipy/brightness.py:41-49) is done ONCE and shared by
// the four densities, the bands share 1/R, constants are merged on the host (log2(e) factors,
// reciprocals, emissivities), and powers are multiplication chains.  This cuts the executed work
// from ~60 flops + 7 SFU ops per (pixel x component x node) evaluation (SURVEY.md 8(d) canonical
// count) to ~27 issue slots + 3.25 SFU ops.  Ring and Feature own their ranges (cutoffs
// (0.8, 1.2) and (0.7, 1.3) AU) and are integrated in two further node loops.
//
// Models that do not fit (user-edited band offsets or p != 4, RRM, ...) use the generic kernel
// (zodi_device.cuh); eligibility is decided once per model in build_kelsall_model().
#pragma once

#include <string.h>

#include "zodi_device.cuh"

namespace zodi {

template <typename Real>
struct KelsallModel {
    int n_comps;   // 4 or 6
    int n_nodes;
    int n_temps;
    int scatter;   // any albedo != 0
    int share13;   // band3 uses the same delta_r as band1 -> shares the radial cutoff term
    // source function
    Real t_scale;  // T0 / dT           : t = t_scale * 2^(mhd * log2 R^2) + t_ofs  (table coordinate)
    Real t_ofs;    // -T_min / dT
    Real t_top;    // n_temps - 1
    Real mhd;      // -delta / 2
    Real C1p, C2p, C3l;   // phase function: C1, C2, C3*log2e
    int phase_poly_ok;    // fp32: cancellation-free polynomial form (zodi_device.cuh: phase_of_cos)
    int phase_terms;      // 8 or kPhaseTerms: coefficients that are not zero padding
    Real phase_poly[kPhaseTerms];
    Real aB[6];    // (1 - albedo_c) * emissivity_c * amplitude_c     (thermal)
    Real aS[6];    // albedo_c * F_sun * N_phase * amplitude_c        (scattering)
    // cloud (number_density.py:47-73)
    Real cx0, cy0, cz0, cnx, cny, cnz;
    Real c_mu, c_inv2mu, c_halfmu, c_mha, c_mbl, c_gamma;
    // bands (number_density.py:76-110); normals pre-scaled by log2e^(1/6) / delta_zeta
    Real bnx[3], bny[3], bnz[3];
    Real b_c3[3];  // 1 / (v * log2e^(2/3))
    Real b_y[3];   // log2e^(1/10) / delta_r^2
    // ring / feature (number_density.py:113-181), both centred on the Sun
    Real rnx, rny, rnz, r_R, r_c2, r_c3;
    Real fnx, fny, fnz, f_R, f_c2, f_c3, f_c5, f_theta0;
    // ranges (always double)
    double cutA_in, cutA_out, cutR_in, cutR_out, cutF_in, cutF_out;
    double f_cos0, f_sin0;  // cos / sin of the feature's theta_0 (see feature_rotation)
};

// --- table lookup -----------------------------------------------------------------------------
// t = table coordinate (knot units); see table_coord() in zodi_device.cuh.
template <typename Real>
ZODI_HD Real table_at(const Pair<Real>* tab, Real t, Real t_top) {
    int idx;
    Real frac;
    table_coord<Real>(t, t_top, idx, frac);
    const Pair<Real> e = tab[idx];  // idx may equal t_top (last knot): its stored delta is 0
    return Math<Real>::fma_(e.b, frac, e.a);
}

// Shared per-node source quantities.
template <typename Real>
struct NodeSource {
    Real xh, yh, zh, Rh2, lgR, B, F;  // F = Phi(Theta) / R_h^2 (scattering only)
};

template <typename Real, bool SCATTER>
ZODI_HD NodeSource<Real> node_source(const KelsallModel<Real>& K, const Pair<Real>* tab, Real R_los,
                                     Real ux, Real uy, Real uz, Real ox, Real oy, Real oz) {
    using M = Math<Real>;
    NodeSource<Real> s;
    s.xh = M::fma_(R_los, ux, ox);
    s.yh = M::fma_(R_los, uy, oy);
    s.zh = M::fma_(R_los, uz, oz);
    s.Rh2 = M::fma_(s.xh, s.xh, M::fma_(s.yh, s.yh, s.zh * s.zh));
    s.lgR = M::log2_(s.Rh2);
    // |mhd lg2 R^2| < 1020 for any distance between 1e-300 and 1e300 AU: no underflow test needed
    const Real t = M::fma_(K.t_scale, M::exp2_bounded_(K.mhd * s.lgR), K.t_ofs);  // blackbody.py:30
    s.B = table_at<Real>(tab, t, K.t_top);                                 // brightness.py:48
    s.F = Real(0);
    if (SCATTER) {  // brightness.py:50-54, scattering.py:29-50
        const Real rh_inv = M::rsqrt_(s.Rh2);
        Real ct = M::fma_(ux, s.xh, M::fma_(uy, s.yh, uz * s.zh)) * rh_inv;
        ct = M::max_(Real(-1), M::min_(Real(1), ct));
        s.F = phase_of_cos<Real>(ct, K.C1p, K.C2p, K.C3l, K.phase_poly_ok, K.phase_terms, K.phase_poly) * rh_inv * rh_inv;
    }
    return s;
}

// Adds wB * n_band (and wF * n_band when scattering) to the accumulators, with
// n_band = exp(-s^6) (1 + s^4/v) (rinv rad); (bx,by,bz) is the pre-scaled plane normal.  Skipped
// (n_band == 0 exactly) when every lane of the warp is so far from the band plane that exp(-s^6)
// underflows to zero.
template <typename Real, bool SCATTER>
ZODI_HD void band_accumulate(Real& accB, Real& accS, Real wB, Real wF, Real xh, Real yh, Real zh,
                             Real rinv, Real rinv_rad, Real bx, Real by, Real bz, Real c3) {
    using M = Math<Real>;
    const Real sz = M::fma_(xh, bx, M::fma_(yh, by, zh * bz)) * rinv;  // sign irrelevant (even powers)
    const Real s2 = sz * sz;
    // decided on s^2 (kS2Underflow): lanes with s^6 beyond the underflow point inside an executing warp get
    // their exact 0 from exp2_neg_ itself
    if (warp_any(s2 <= M::kS2Underflow)) {
        const Real s4 = s2 * s2, s6 = s4 * s2;
        const Real n = (M::exp2_neg_(s6) * M::fma_(s4, c3, Real(1))) * rinv_rad;
        accB = M::fma_(wB, n, accB);
        if (SCATTER) accS = M::fma_(wF, n, accS);
    }
}

// band_accumulate with the plane distance dot = n . X already formed and the skip already decided by the
// caller: same operations on the lanes that are computed.
template <typename Real>
ZODI_HD void band_accumulate_dot(Real& accB, Real wB, Real dot, Real rinv, Real rinv_rad, Real c3) {
    using M = Math<Real>;
    const Real sz = dot * rinv;
    const Real s2 = sz * sz, s4 = s2 * s2, s6 = s4 * s2;
    const Real n = (M::exp2_neg_(s6) * M::fma_(s4, c3, Real(1))) * rinv_rad;
    accB = M::fma_(wB, n, accB);
}

// 1 - exp(-(R/delta_r)^20) with y = R^2 * log2e^(1/10) / delta_r^2  (y^10 = log2e * (R/delta_r)^20).
// Beyond kRadialOne (R > ~1.2 delta_r: most of a line of sight that runs out to 5.2 AU) the term is
// exactly 1; when the whole warp is there the power chain and the exponential are skipped.
template <typename Real>
ZODI_HD Real band_radial(Real Rh2, Real by) {
    using M = Math<Real>;
    const Real y = Rh2 * by;
    Real rad = Real(1);
    if (warp_any(y < M::kRadialOne)) {
        const Real y2 = y * y, y4 = y2 * y2, y5 = y4 * y;
        rad = M::one_minus_exp2_neg(y5 * y5);
    }
    return rad;
}

// Cloud density without its amplitude: Rc^-alpha exp(-beta g^gamma), g = zeta^2 / (2 mu) below the break
// zeta = |Z_c| / R_c < mu and zeta - mu / 2 above it (number_density.py:69-73).
// (Tried and dropped, profiles/r2_ab_cloud_logform.jsonl: lg2 g = 2 lg2|Z_c| - lg2 R_c^2 - lg2 2mu for lanes
// below the break, which spares warps that are entirely below it the reciprocal square root - 4 instead of
// 5 MUFU - measured 6 % SLOWER: two more votes and the mixed warps cost more than the MUFU pair saves.)
template <typename Real>
ZODI_HD Real cloud_density(const KelsallModel<Real>& K, Real Rc2, Real Zc) {
    using M = Math<Real>;
    const Real zeta = M::abs_(Zc) * M::rsqrt_(Rc2);
    const Real g = (zeta < K.c_mu) ? zeta * zeta * K.c_inv2mu : zeta - K.c_halfmu;
    const Real gp = M::exp2_(K.c_gamma * M::log2_(g));
    return M::exp2_(M::fma_(K.c_mha, M::log2_(Rc2), K.c_mbl * gp));
}

// Per-line-of-sight quantities shared by the component groups (prologue in double).
template <typename Real>
struct LosGeometry {
    double r_obs2, bq;       // observer distance^2 and b/2 of the ray/sphere quadratic
    Real ux, uy, uz, ox, oy, oz;
};

template <typename Real>
ZODI_HD LosGeometry<Real> los_geometry(double dux, double duy, double duz, double dox, double doy,
                                       double doz) {
    LosGeometry<Real> g;
    g.r_obs2 = dox * dox + doy * doy + doz * doz;
    g.bq = ray_bq(dux, duy, duz, dox, doy);
    g.ux = Real(dux); g.uy = Real(duy); g.uz = Real(duz);
    g.ox = Real(dox); g.oy = Real(doy); g.oz = Real(doz);
    return g;
}

// Half-range and mid-point of the quadrature interval between two cutoff spheres.
template <typename Real>
ZODI_HD void los_interval(const LosGeometry<Real>& g, double cut_in, double cut_out, bool out_in,
                          bool out_out, Real& h, Real& mid) {
    const double start = sphere_distance(g.bq, g.r_obs2, cut_in, out_in);
    const double stop = sphere_distance(g.bq, g.r_obs2, cut_out, out_out);
    h = Real(0.5 * (stop - start));    // brightness.py:41
    mid = Real(0.5 * (stop + start));
}

// ---------------- group A: cloud + band1..3 on one grid ---------------------------------------
template <typename Real, bool SCATTER, bool SHARE13, typename Emit>
ZODI_HD void kelsall_group_a(const KelsallModel<Real>& K, const Pair<Real>* tab, const Pair<Real>* nodes,
                             const LosGeometry<Real>& G, uint32_t outside_mask, int sub, int L, Emit emit) {
    using M = Math<Real>;
    Real h, mid;
    los_interval<Real>(G, K.cutA_in, K.cutA_out, outside_mask & 1u, (outside_mask >> 1) & 1u, h, mid);
    Real aB0 = 0, aB1 = 0, aB2 = 0, aB3 = 0, aS0 = 0, aS1 = 0, aS2 = 0, aS3 = 0;
    const Real by_min = M::min_(K.b_y[0], M::min_(K.b_y[1], K.b_y[2]));
    // Warp-uniform trip count: the body contains warp votes (warp_any), so every lane must run the
    // same number of iterations even when n_nodes is not a multiple of L; surplus iterations
    // re-evaluate the last node with weight 0.
    for (int k0 = 0; k0 < K.n_nodes; k0 += L) {
        const int k = k0 + sub;
        Pair<Real> nw = nodes[k < K.n_nodes ? k : K.n_nodes - 1];
        if (k >= K.n_nodes) nw.b = Real(0);
        const NodeSource<Real> s = node_source<Real, SCATTER>(K, tab, M::fma_(h, nw.a, mid), G.ux, G.uy,
                                                              G.uz, G.ox, G.oy, G.oz);
        // bands: centred on the Sun -> share R
        const Real wB = nw.b * s.B, wF = SCATTER ? nw.b * s.F : Real(0);
        if (!SCATTER) {
            // Band skips decided on (n . X)^2 > c R^2 BEFORE forming 1/R (see kelsall_group_a_x2): when every
            // band of the (warp, node) is skipped the reciprocal square root is dead too.  The margin over
            // kS2Underflow covers the rounding of rsqrt; lanes in between are computed and get an exact 0.
            const Real d1 = M::fma_(s.xh, K.bnx[0], M::fma_(s.yh, K.bny[0], s.zh * K.bnz[0]));
            const Real d2 = M::fma_(s.xh, K.bnx[1], M::fma_(s.yh, K.bny[1], s.zh * K.bnz[1]));
            const Real d3 = M::fma_(s.xh, K.bnx[2], M::fma_(s.yh, K.bny[2], s.zh * K.bnz[2]));
            const Real thr = s.Rh2 * (M::kS2Underflow * Real(1.00001));
            const bool need1 = warp_any(d1 * d1 <= thr), need2 = warp_any(d2 * d2 <= thr), need3 = warp_any(d3 * d3 <= thr);
            if (need1 || need2 || need3) {
                const Real rinv = M::rsqrt_(s.Rh2);
                Real rr1 = rinv, rr2 = rinv, rr3 = rinv;
                if (warp_any(s.Rh2 * by_min < M::kRadialOne)) {
                    if (need1 || (SHARE13 && need3)) rr1 = rinv * band_radial<Real>(s.Rh2, K.b_y[0]);
                    if (need2) rr2 = rinv * band_radial<Real>(s.Rh2, K.b_y[1]);
                    if (SHARE13) rr3 = rr1;
                    else if (need3) rr3 = rinv * band_radial<Real>(s.Rh2, K.b_y[2]);
                }
                if (need1) band_accumulate_dot<Real>(aB1, wB, d1, rinv, rr1, K.b_c3[0]);
                if (need2) band_accumulate_dot<Real>(aB2, wB, d2, rinv, rr2, K.b_c3[1]);
                if (need3) band_accumulate_dot<Real>(aB3, wB, d3, rinv, rr3, K.b_c3[2]);
            }
        } else {
        const Real rinv = M::rsqrt_(s.Rh2);
        // all radial cut-off factors are exactly 1 once R^2 min(b_y) >= kRadialOne in every lane (b_y > 0, the
        // products are monotonic in b_y): one vote then replaces the per-band ones and rinv * 1 is rinv
        if (!warp_any(s.Rh2 * by_min < M::kRadialOne)) {
            band_accumulate<Real, SCATTER>(aB1, aS1, wB, wF, s.xh, s.yh, s.zh, rinv, rinv, K.bnx[0], K.bny[0], K.bnz[0], K.b_c3[0]);
            band_accumulate<Real, SCATTER>(aB2, aS2, wB, wF, s.xh, s.yh, s.zh, rinv, rinv, K.bnx[1], K.bny[1], K.bnz[1], K.b_c3[1]);
            band_accumulate<Real, SCATTER>(aB3, aS3, wB, wF, s.xh, s.yh, s.zh, rinv, rinv, K.bnx[2], K.bny[2], K.bnz[2], K.b_c3[2]);
        } else {
            const Real rad1 = band_radial<Real>(s.Rh2, K.b_y[0]);
            const Real rad2 = band_radial<Real>(s.Rh2, K.b_y[1]);
            const Real rad3 = SHARE13 ? rad1 : band_radial<Real>(s.Rh2, K.b_y[2]);
            band_accumulate<Real, SCATTER>(aB1, aS1, wB, wF, s.xh, s.yh, s.zh, rinv, rinv * rad1, K.bnx[0], K.bny[0], K.bnz[0], K.b_c3[0]);
            band_accumulate<Real, SCATTER>(aB2, aS2, wB, wF, s.xh, s.yh, s.zh, rinv, rinv * rad2, K.bnx[1], K.bny[1], K.bnz[1], K.b_c3[1]);
            band_accumulate<Real, SCATTER>(aB3, aS3, wB, wF, s.xh, s.yh, s.zh, rinv, rinv * rad3, K.bnx[2], K.bny[2], K.bnz[2], K.b_c3[2]);
        }
        }
        // cloud
        const Real xc = s.xh - K.cx0, yc = s.yh - K.cy0, zc = s.zh - K.cz0;
        const Real Rc2 = M::fma_(xc, xc, M::fma_(yc, yc, zc * zc));
        const Real n0 = cloud_density<Real>(K, Rc2, M::fma_(xc, K.cnx, M::fma_(yc, K.cny, zc * K.cnz)));

        aB0 = M::fma_(wB, n0, aB0);
        if (SCATTER) aS0 = M::fma_(wF, n0, aS0);
    }
    emit(0, h * M::fma_(K.aB[0], aB0, K.aS[0] * aS0));
    emit(1, h * M::fma_(K.aB[1], aB1, K.aS[1] * aS1));
    emit(2, h * M::fma_(K.aB[2], aB2, K.aS[2] * aS2));
    emit(3, h * M::fma_(K.aB[3], aB3, K.aS[3] * aS3));
}

// ---------------- ring (own grid) -------------------------------------------------------------
// Phi(Theta) / R_h^2 at a node (brightness.py:50-54, scattering.py:29-50); rh_inv = 1 / R_h.
template <typename Real>
ZODI_HD Real scatter_term(const KelsallModel<Real>& K, Real ux, Real uy, Real uz, Real xh, Real yh, Real zh,
                          Real rh_inv) {
    using M = Math<Real>;
    Real ct = M::fma_(ux, xh, M::fma_(uy, yh, uz * zh)) * rh_inv;
    ct = M::max_(Real(-1), M::min_(Real(1), ct));
    return phase_of_cos<Real>(ct, K.C1p, K.C2p, K.C3l, K.phase_poly_ok, K.phase_terms, K.phase_poly) * rh_inv * rh_inv;
}

template <typename Real, bool SCATTER>
ZODI_HD Real kelsall_ring(const KelsallModel<Real>& K, const Pair<Real>* tab, const Pair<Real>* nodes,
                          const LosGeometry<Real>& G, uint32_t outside_mask, int sub, int L) {
    using M = Math<Real>;
    Real h, mid;
    los_interval<Real>(G, K.cutR_in, K.cutR_out, (outside_mask >> 8) & 1u, (outside_mask >> 9) & 1u, h, mid);
    Real aB = 0, aS = 0;
    for (int k = sub; k < K.n_nodes; k += L) {
        const Pair<Real> nw = nodes[k];
        const Real R_los = M::fma_(h, nw.a, mid);
        const Real xh = M::fma_(R_los, G.ux, G.ox), yh = M::fma_(R_los, G.uy, G.oy), zh = M::fma_(R_los, G.uz, G.oz);
        const Real Rh2 = M::fma_(xh, xh, M::fma_(yh, yh, zh * zh));
        const Real d = M::sqrt_(Rh2) - K.r_R;
        const Real B = table_at<Real>(tab, M::fma_(K.t_scale, M::exp2_bounded_(K.mhd * M::log2_(Rh2)), K.t_ofs), K.t_top);
        const Real Zc = M::fma_(xh, K.rnx, M::fma_(yh, K.rny, zh * K.rnz));
        const Real n = M::exp2_neg_(-M::fma_(d * d, K.r_c2, M::abs_(Zc) * K.r_c3));
        aB = M::fma_(nw.b * B, n, aB);
        if (SCATTER) aS = M::fma_(nw.b * scatter_term<Real>(K, G.ux, G.uy, G.uz, xh, yh, zh, M::rsqrt_(Rh2)), n, aS);
    }
    return h * M::fma_(K.aB[4], aB, K.aS[4] * aS);
}

// cos / sin of  theta = atan2(dey, dex) + theta_0  (Earth's longitude about the feature's centre plus
// the trailing offset, number_density.py:163-170): the nodes are rotated by -theta so that atan2 returns
// the wrapped longitude offset directly.  cos(atan2(y, x)) = x / r and sin = y / r, so the angle sum
// needs one reciprocal square root instead of atan2 + cos + sin in double per line of sight.
template <typename Real>
ZODI_HD void feature_rotation(double dex, double dey, double cos0, double sin0, Real& cr, Real& sr) {
    const double r2 = dex * dex + dey * dey;
    double c = 1.0, s = 0.0;  // atan2(0, 0) = 0
    if (r2 > 0.0) {
        const double inv = Math<double>::rsqrt_(r2);
        c = dex * inv;
        s = dey * inv;
    }
    cr = Real(c * cos0 - s * sin0);
    sr = Real(s * cos0 + c * sin0);
}

// ---------------- feature (own grid) ----------------------------------------------------------
template <typename Real, bool SCATTER>
ZODI_HD Real kelsall_feature(const KelsallModel<Real>& K, const Pair<Real>* tab, const Pair<Real>* nodes,
                             const LosGeometry<Real>& G, double dex, double dey, uint32_t outside_mask,
                             int sub, int L) {
    using M = Math<Real>;
    Real h, mid;
    los_interval<Real>(G, K.cutF_in, K.cutF_out, (outside_mask >> 10) & 1u, (outside_mask >> 11) & 1u, h, mid);
    // rotate by -(theta_earth + theta_0): then atan2 gives the wrapped longitude offset directly
    // (number_density.py:163-170; [-pi,pi) vs (-pi,pi] only differs at |delta| = pi where the
    // squared offset is identical)
    Real cr, sr;
    feature_rotation<Real>(dex, dey, K.f_cos0, K.f_sin0, cr, sr);
    Real aB = 0, aS = 0;
    for (int k = sub; k < K.n_nodes; k += L) {
        const Pair<Real> nw = nodes[k];
        const NodeSource<Real> s = node_source<Real, SCATTER>(K, tab, M::fma_(h, nw.a, mid), G.ux, G.uy,
                                                              G.uz, G.ox, G.oy, G.oz);
        const Real d = M::sqrt_(s.Rh2) - K.f_R;
        const Real Zc = M::fma_(s.xh, K.fnx, M::fma_(s.yh, K.fny, s.zh * K.fnz));
        const Real xr = M::fma_(s.xh, cr, s.yh * sr), yr = M::fma_(s.yh, cr, -(s.xh * sr));
        const Real dth = M::atan2_abs_(yr, xr);  // only dth^2 is used
        const Real e = M::fma_(d * d, K.f_c2, M::fma_(M::abs_(Zc), K.f_c3, dth * dth * K.f_c5));
        const Real n = M::exp2_neg_(-e);
        aB = M::fma_(nw.b * s.B, n, aB);
        if (SCATTER) aS = M::fma_(nw.b * s.F, n, aS);
    }
    return h * M::fma_(K.aB[5], aB, K.aS[5] * aS);
}

template <typename Real, bool HAS_RF, bool SCATTER, bool SHARE13, typename Emit>
ZODI_HD void integrate_kelsall(const KelsallModel<Real>& K, const Pair<Real>* tab,
                               const Pair<Real>* nodes, double dux, double duy, double duz,
                               double dox, double doy, double doz, double dex, double dey,
                               uint32_t outside_mask, int sub, int L, Emit emit) {
    const LosGeometry<Real> G = los_geometry<Real>(dux, duy, duz, dox, doy, doz);
    kelsall_group_a<Real, SCATTER, SHARE13>(K, tab, nodes, G, outside_mask, sub, L, emit);
    if (!HAS_RF) return;
    emit(4, kelsall_ring<Real, SCATTER>(K, tab, nodes, G, outside_mask, sub, L));
    emit(5, kelsall_feature<Real, SCATTER>(K, tab, nodes, G, dex, dey, outside_mask, sub, L));
}

}  // namespace zodi
