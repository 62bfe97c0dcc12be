// zodi_model_build.hpp - pure C++ (no CUDA calls): raw model descriptor -> device-form constants.
// Shared by the C-ABI implementation (zodi_capi.cu) and the host-emulation debug harness under
// tests/host_emu (test tool only).
#pragma once

#include <cmath>
#include <cstring>
#include <vector>

#include "zodi_device.cuh"
#include "zodi_kelsall.cuh"
#include "zodi_rrm.cuh"
#include "zodi_rrm_x2.cuh"
#include "zodi_multiband.cuh"

namespace zodi {

inline int n_shape_params(int type) {
    switch (type) {
        case ZODI_CLOUD: return 5;
        case ZODI_BAND: return 5;
        case ZODI_RING: return 4;
        case ZODI_FEATURE: return 6;
        case ZODI_FAN: return 5;
        case ZODI_COMET: return 6;
        case ZODI_INTERSTELLAR: return 1;
        case ZODI_NARROW_BAND: return 6;
        case ZODI_BROAD_BAND: return 6;
        case ZODI_RING_RRM: return 5;
        case ZODI_FEATURE_RRM: return 7;
        default: return -1;
    }
}

// Raw reference parameters -> constants consumed by density<Real>() (zodi_device.cuh).
inline void derive_component(const zodi_model_desc& d, const zodi_component_desc& c, double phase_norm,
                      DevComp<double>& o) {
    std::memset(&o, 0, sizeof(o));
    const double* p = c.shape;
    o.x0 = c.x0[0]; o.y0 = c.x0[1]; o.z0 = c.x0[2];
    // Z_c = X_c.x sinO sini - X_c.y cosO sini + X_c.z cosi  (e.g. number_density.py:63-67)
    o.nx = c.sin_Omega * c.sin_i;
    o.ny = -c.cos_Omega * c.sin_i;
    o.nz = c.cos_i;
    o.cut_in = c.cutoff_inner;
    o.cut_out = c.cutoff_outer;
    if (d.kind == ZODI_KELSALL) {
        o.scatter = (c.albedo != 0.0);                       // brightness.py:50
        o.e1 = (1.0 - c.albedo) * c.emissivity;               // brightness.py:49
        o.sc = c.albedo * d.solar_irradiance * phase_norm;    // brightness.py:51-54
        o.T0 = d.T_0;
        o.mhd = -0.5 * d.delta;
    } else {
        o.scatter = 0;
        o.e1 = d.calibration;                                 // brightness.py:83
        o.sc = 0.0;
        o.T0 = c.T_0;                                         // unpack_model.py:123-126
        o.mhd = -0.5 * c.delta;
    }
    switch (c.type) {
        case ZODI_CLOUD:  // n_0, alpha, beta, gamma, mu
            o.type = D_CLOUD;
            o.s[0] = p[0]; o.s[1] = p[4]; o.s[2] = 1.0 / (2.0 * p[4]); o.s[3] = 0.5 * p[4];
            o.s[4] = -0.5 * p[1]; o.s[5] = -p[2] * kLog2e; o.s[6] = p[3];
            break;
        case ZODI_BAND:  // n_0, delta_zeta_rad, v, p, delta_r
            o.type = D_BAND;
            o.s[0] = 3.0 * p[0]; o.s[1] = 1.0 / p[1]; o.s[2] = p[3]; o.s[3] = 1.0 / p[2];
            o.s[4] = 1.0 / p[4]; o.s[5] = (p[3] == 4.0) ? 1.0 : 0.0; o.s[6] = kLog2e;
            break;
        case ZODI_RING_RRM:
        case ZODI_RING:  // n_0, R, sigma_r, sigma_z [, A]
            o.type = D_RING;
            o.s[0] = (c.type == ZODI_RING_RRM) ? p[4] * p[0] : p[0];
            o.s[1] = p[1]; o.s[2] = -kLog2e / (p[2] * p[2]); o.s[3] = -kLog2e / p[3];
            break;
        case ZODI_FEATURE_RRM:
        case ZODI_FEATURE:  // n_0, R, sigma_r, sigma_z, theta_rad, sigma_theta_rad [, A]
            o.type = D_FEATURE;
            o.s[0] = (c.type == ZODI_FEATURE_RRM) ? p[6] * p[0] : p[0];
            o.s[1] = p[1]; o.s[2] = -kLog2e / (p[2] * p[2]); o.s[3] = -kLog2e / p[3];
            o.s[4] = p[4]; o.s[5] = -kLog2e / (p[5] * p[5]);
            break;
        case ZODI_FAN:  // Q, P, gamma, Z_0, R_outer
            o.type = D_FAN;
            o.s[0] = -0.5 * p[2]; o.s[1] = 1.0 / p[3]; o.s[2] = -p[1] * kLog2e; o.s[3] = 1.0;
            o.s[4] = 0.0; o.s[5] = p[4] * p[4]; o.s[6] = p[3]; o.s[7] = p[0];
            break;
        case ZODI_COMET:  // gamma, Z_0, P, amp, R_inner, R_outer
            o.type = D_COMET;
            o.s[0] = -0.5 * p[0]; o.s[1] = 1.0 / p[1]; o.s[2] = -p[2] * kLog2e; o.s[3] = p[3];
            o.s[4] = p[4] * p[4]; o.s[5] = p[5] * p[5]; o.s[6] = p[1]; o.s[7] = 0.0;
            break;
        case ZODI_INTERSTELLAR:
            o.type = D_INTERSTELLAR;
            o.s[0] = p[0];
            break;
        case ZODI_NARROW_BAND:  // beta_nb, G, gamma, A, R_inner, R_outer
            o.type = D_NARROW;
            o.s[0] = p[0]; o.s[1] = p[1] * kLog2e; o.s[2] = -0.5 * p[2];
            o.s[3] = p[3] * std::pow(p[5], p[2]); o.s[4] = p[4] * p[4]; o.s[5] = p[5] * p[5];
            break;
        case ZODI_BROAD_BAND:  // beta_bb, sigma_bb, gamma, A, R_inner, R_outer
            o.type = D_BROAD;
            o.s[0] = p[0]; o.s[1] = 1.0 / p[1]; o.s[2] = -0.5 * p[2];
            o.s[3] = p[3] * std::pow(p[5], p[2]); o.s[4] = p[4] * p[4]; o.s[5] = p[5] * p[5];
            o.s[6] = -0.5 * kLog2e;
            o.s[7] = (p[2] == 1.0) ? 1.0 : 0.0;  // gamma == 1: R^-gamma is 1 / R (fused RRM kernel)
            break;
    }
}

// Taylor coefficients of C1 + C2 T + exp(C3 T) about T = pi/2 (see phase_of_cos()).  Returns the
// number of terms kept (8 or kPhaseTerms, the rest zero-padded) such that the truncation error on
// |T - pi/2| <= pi/2 is below 1e-9 of the leading exponential term, or 0 if kPhaseTerms do not
// reach that (then the device evaluates the literal form).
inline int phase_polynomial(double C1, double C2, double C3, double* a) {
    const double t0 = 0.5 * kPi, E = std::exp(C3 * t0);
    a[0] = C1 + C2 * t0 + E;
    a[1] = C2 + C3 * E;
    double term = C3 * E;
    int terms = 0;
    for (int k = 2; k <= kPhaseTerms; ++k) {
        term *= C3 / k;  // coefficient of t^k
        // remainder after k terms <= |a_k| t0^k e^{|C3| t0}
        if (!terms && (k == 8 || k == kPhaseTerms) &&
            std::fabs(term) * std::pow(t0, k) * std::exp(std::fabs(C3) * t0) <= 1e-9 * E)
            terms = k;
        if (k < kPhaseTerms) a[k] = terms ? 0.0 : term;
    }
    return terms;
}

template <typename To, typename From>
inline void narrow_model(const DevModel<From>& a, DevModel<To>& b) {
    std::memset(&b, 0, sizeof(b));
    b.n_comps = a.n_comps; b.n_nodes = a.n_nodes; b.n_temps = a.n_temps;
    b.has_feature = a.has_feature;
    b.t_min = (To)a.t_min; b.inv_dt = (To)a.inv_dt;
    b.C1 = (To)a.C1; b.C2 = (To)a.C2; b.C3 = (To)a.C3;
    b.phase_poly_ok = a.phase_poly_ok; b.phase_terms = a.phase_terms;
    for (int k = 0; k < kPhaseTerms; ++k) b.phase_poly[k] = (To)a.phase_poly[k];
    for (int i = 0; i < a.n_comps; ++i) {
        const DevComp<From>& s = a.comps[i];
        DevComp<To>& t = b.comps[i];
        t.type = s.type; t.scatter = s.scatter;
        t.x0 = (To)s.x0; t.y0 = (To)s.y0; t.z0 = (To)s.z0;
        t.nx = (To)s.nx; t.ny = (To)s.ny; t.nz = (To)s.nz;
        for (int k = 0; k < 8; ++k) t.s[k] = (To)s.s[k];
        t.e1 = (To)s.e1; t.sc = (To)s.sc; t.T0 = (To)s.T0; t.mhd = (To)s.mhd;
        t.cut_in = s.cut_in; t.cut_out = s.cut_out;
    }
}


// Whole model: shared scalars + per-component constants (double).
inline void build_dev_model(const zodi_model_desc& d, DevModel<double>& M) {
    std::memset(&M, 0, sizeof(M));
    M.n_comps = d.n_comps; M.n_nodes = d.n_nodes; M.n_temps = d.n_temps;
    M.t_min = d.temps[0];
    M.inv_dt = (d.n_temps - 1) / (d.temps[d.n_temps - 1] - d.temps[0]);
    M.C1 = d.C1; M.C2 = d.C2; M.C3 = d.C3 * kLog2e;
    M.phase_terms = phase_polynomial(d.C1, d.C2, d.C3, M.phase_poly);
    M.phase_poly_ok = M.phase_terms ? 1 : 0;
    // _get_phase_normalization, scattering.py:53-59
    const double phase_norm =
        1.0 / (2.0 * kPi * (2.0 * d.C1 + kPi * d.C2 + (std::exp(d.C3 * kPi) + 1.0) / (d.C3 * d.C3 + 1.0)));
    M.has_feature = 0;
    for (int i = 0; i < d.n_comps; ++i) {
        derive_component(d, d.comps[i], phase_norm, M.comps[i]);
        if (M.comps[i].type == D_FEATURE) M.has_feature = 1;
    }
}

// Blackbody table as (B_i, B_{i+1}-B_i) pairs, quadrature as (x_k, w_k) pairs.
inline void build_pairs(const zodi_model_desc& d, std::vector<Pair<double>>& t64,
                        std::vector<Pair<double>>& n64, std::vector<Pair<float>>& t32,
                        std::vector<Pair<float>>& n32) {
    t64.resize(d.n_temps); t32.resize(d.n_temps);
    n64.resize(d.n_nodes); n32.resize(d.n_nodes);
    for (int i = 0; i < d.n_temps; ++i) {
        const double b0 = d.bnu[i], b1 = d.bnu[i + 1 < d.n_temps ? i + 1 : i];
        t64[i] = {b0, b1 - b0};
        t32[i] = {(float)b0, (float)(b1 - b0)};
    }
    for (int i = 0; i < d.n_nodes; ++i) {
        n64[i] = {d.nodes[i], d.weights[i]};
        n32[i] = {(float)d.nodes[i], (float)d.weights[i]};
    }
}

// Kelsall-family fast path: returns false if the model does not have the layout the fused
// kernel assumes (then the generic kernel is used).  See zodi_kelsall.cuh.
inline bool build_kelsall_model(const zodi_model_desc& d, KelsallModel<double>& K) {
    std::memset(&K, 0, sizeof(K));
    if (d.kind != ZODI_KELSALL) return false;
    if (d.n_comps != 4 && d.n_comps != 6) return false;
    const zodi_component_desc* c = d.comps;
    if (c[0].type != ZODI_CLOUD) return false;
    for (int b = 1; b <= 3; ++b) {
        if (c[b].type != ZODI_BAND) return false;
        if (c[b].x0[0] != 0.0 || c[b].x0[1] != 0.0 || c[b].x0[2] != 0.0) return false;  // share R
        if (c[b].shape[3] != 4.0) return false;                                          // p == 4
        if (c[b].cutoff_inner != c[0].cutoff_inner || c[b].cutoff_outer != c[0].cutoff_outer)
            return false;                                                                // one grid
        if (!(c[b].shape[1] > 0.0) || !(c[b].shape[2] != 0.0) || !(c[b].shape[4] > 0.0)) return false;
    }
    if (d.n_comps == 6) {
        if (c[4].type != ZODI_RING || c[5].type != ZODI_FEATURE) return false;
        for (int r = 4; r <= 5; ++r)
            if (c[r].x0[0] != 0.0 || c[r].x0[1] != 0.0 || c[r].x0[2] != 0.0) return false;
    }
    K.n_comps = d.n_comps; K.n_nodes = d.n_nodes; K.n_temps = d.n_temps;
    const double dt = (d.temps[d.n_temps - 1] - d.temps[0]) / (d.n_temps - 1);
    K.t_scale = d.T_0 / dt;
    K.t_ofs = -d.temps[0] / dt;
    K.t_top = d.n_temps - 1;
    K.mhd = -0.5 * d.delta;
    K.C1p = d.C1; K.C2p = d.C2; K.C3l = d.C3 * kLog2e;
    K.phase_terms = phase_polynomial(d.C1, d.C2, d.C3, K.phase_poly);
    K.phase_poly_ok = K.phase_terms ? 1 : 0;
    const double phase_norm =
        1.0 / (2.0 * kPi * (2.0 * d.C1 + kPi * d.C2 + (std::exp(d.C3 * kPi) + 1.0) / (d.C3 * d.C3 + 1.0)));
    K.scatter = 0;
    for (int i = 0; i < d.n_comps; ++i) {
        const double amp = (c[i].type == ZODI_BAND ? 3.0 : 1.0) * c[i].shape[0];  // n_0 (3 n_0 for bands)
        K.aB[i] = (1.0 - c[i].albedo) * c[i].emissivity * amp;   // brightness.py:49
        K.aS[i] = c[i].albedo * d.solar_irradiance * phase_norm * amp;  // brightness.py:51-54
        if (c[i].albedo != 0.0) K.scatter = 1;
    }
    // cloud: n_0, alpha, beta, gamma, mu
    K.cx0 = c[0].x0[0]; K.cy0 = c[0].x0[1]; K.cz0 = c[0].x0[2];
    K.cnx = c[0].sin_Omega * c[0].sin_i; K.cny = -c[0].cos_Omega * c[0].sin_i; K.cnz = c[0].cos_i;
    K.c_mu = c[0].shape[4]; K.c_inv2mu = 1.0 / (2.0 * c[0].shape[4]); K.c_halfmu = 0.5 * c[0].shape[4];
    K.c_mha = -0.5 * c[0].shape[1]; K.c_mbl = -c[0].shape[2] * kLog2e; K.c_gamma = c[0].shape[3];
    // bands: n_0, delta_zeta_rad, v, p, delta_r
    const double l6 = std::pow(kLog2e, 1.0 / 6.0), l23 = std::pow(kLog2e, 2.0 / 3.0),
                 l10 = std::pow(kLog2e, 0.1);
    for (int b = 0; b < 3; ++b) {
        const zodi_component_desc& q = c[b + 1];
        const double sc = l6 / q.shape[1];
        K.bnx[b] = q.sin_Omega * q.sin_i * sc; K.bny[b] = -q.cos_Omega * q.sin_i * sc; K.bnz[b] = q.cos_i * sc;
        K.b_c3[b] = 1.0 / (q.shape[2] * l23);
        K.b_y[b] = l10 / (q.shape[4] * q.shape[4]);
    }
    K.share13 = (c[1].shape[4] == c[3].shape[4]);
    K.cutA_in = c[0].cutoff_inner; K.cutA_out = c[0].cutoff_outer;
    if (d.n_comps == 6) {
        const zodi_component_desc& r = c[4];  // n_0, R, sigma_r, sigma_z
        K.rnx = r.sin_Omega * r.sin_i; K.rny = -r.cos_Omega * r.sin_i; K.rnz = r.cos_i;
        K.r_R = r.shape[1]; K.r_c2 = -kLog2e / (r.shape[2] * r.shape[2]); K.r_c3 = -kLog2e / r.shape[3];
        K.cutR_in = r.cutoff_inner; K.cutR_out = r.cutoff_outer;
        const zodi_component_desc& f = c[5];  // n_0, R, sigma_r, sigma_z, theta_rad, sigma_theta_rad
        K.fnx = f.sin_Omega * f.sin_i; K.fny = -f.cos_Omega * f.sin_i; K.fnz = f.cos_i;
        K.f_R = f.shape[1]; K.f_c2 = -kLog2e / (f.shape[2] * f.shape[2]); K.f_c3 = -kLog2e / f.shape[3];
        K.f_theta0 = f.shape[4]; K.f_c5 = -kLog2e / (f.shape[5] * f.shape[5]);
        K.f_cos0 = std::cos(f.shape[4]); K.f_sin0 = std::sin(f.shape[4]);
        K.cutF_in = f.cutoff_inner; K.cutF_out = f.cutoff_outer;
    }
    return true;
}

template <typename To, typename From>
inline void narrow_kelsall(const KelsallModel<From>& a, KelsallModel<To>& b) {
    std::memset(&b, 0, sizeof(b));
    b.n_comps = a.n_comps; b.n_nodes = a.n_nodes; b.n_temps = a.n_temps;
    b.scatter = a.scatter; b.share13 = a.share13;
#define ZN(f) b.f = (To)a.f
    ZN(t_scale); ZN(t_ofs); ZN(t_top); ZN(mhd); ZN(C1p); ZN(C2p); ZN(C3l);
    b.phase_poly_ok = a.phase_poly_ok; b.phase_terms = a.phase_terms;
    for (int i = 0; i < kPhaseTerms; ++i) ZN(phase_poly[i]);
    for (int i = 0; i < 6; ++i) { ZN(aB[i]); ZN(aS[i]); }
    ZN(cx0); ZN(cy0); ZN(cz0); ZN(cnx); ZN(cny); ZN(cnz);
    ZN(c_mu); ZN(c_inv2mu); ZN(c_halfmu); ZN(c_mha); ZN(c_mbl); ZN(c_gamma);
    for (int i = 0; i < 3; ++i) { ZN(bnx[i]); ZN(bny[i]); ZN(bnz[i]); ZN(b_c3[i]); ZN(b_y[i]); }
    ZN(rnx); ZN(rny); ZN(rnz); ZN(r_R); ZN(r_c2); ZN(r_c3);
    ZN(fnx); ZN(fny); ZN(fnz); ZN(f_R); ZN(f_c2); ZN(f_c3); ZN(f_c5); ZN(f_theta0);
#undef ZN
    b.cutA_in = a.cutA_in; b.cutA_out = a.cutA_out; b.cutR_in = a.cutR_in; b.cutR_out = a.cutR_out;
    b.cutF_in = a.cutF_in; b.cutF_out = a.cutF_out;
    b.f_cos0 = a.f_cos0; b.f_sin0 = a.f_sin0;
}

// Multi-band form (zodi_multiband.cuh) of n_bands descriptors that differ only in their spectral parameters:
// band 0's geometry plus the per-band source-function scalars and one blackbody table per band
// ([band][knot]).  Returns the index of the first band that is not eligible for the fused layout, -1 if all are.
inline int build_multiband_model(const zodi_model_desc* descs, int n_bands, MultiBandModel<double>& MB,
                                 MultiBandModel<float>& MF, std::vector<Pair<double>>& t64,
                                 std::vector<Pair<float>>& t32) {
    std::memset(&MB, 0, sizeof(MB));
    std::memset(&MF, 0, sizeof(MF));
    if (!build_kelsall_model(descs[0], MB.base)) return 0;
    MB.base.scatter = 0;
    MB.n_bands = n_bands;
    MB.n_bands_padded = n_bands <= 4 ? 4 : (n_bands <= 8 ? 8 : 16);
    const int nt = descs[0].n_temps;
    t64.assign((size_t)n_bands * nt, Pair<double>{0.0, 0.0});
    t32.assign((size_t)n_bands * nt, Pair<float>{0.f, 0.f});
    for (int b = 0; b < n_bands; ++b) {
        KelsallModel<double> kb;
        if (!build_kelsall_model(descs[b], kb)) return b;
        for (int c = 0; c < 6; ++c) { MB.aB[b][c] = kb.aB[c]; MB.aS[b][c] = kb.aS[c]; }
        MB.C1p[b] = kb.C1p; MB.C2p[b] = kb.C2p; MB.C3l[b] = kb.C3l;
        if (kb.scatter) { MB.base.scatter = 1; MB.scatter_bands |= 1u << b; }
        std::vector<Pair<double>> tb, nb;
        std::vector<Pair<float>> tbf, nbf;
        build_pairs(descs[b], tb, nb, tbf, nbf);
        for (int i = 0; i < nt; ++i) { t64[(size_t)b * nt + i] = tb[i]; t32[(size_t)b * nt + i] = tbf[i]; }
    }
    narrow_kelsall(MB.base, MF.base);
    MF.n_bands = MB.n_bands; MF.n_bands_padded = MB.n_bands_padded; MF.scatter_bands = MB.scatter_bands;
    for (int b = 0; b < kMaxBands; ++b) {
        for (int c = 0; c < 6; ++c) { MF.aB[b][c] = (float)MB.aB[b][c]; MF.aS[b][c] = (float)MB.aS[b][c]; }
        MF.C1p[b] = (float)MB.C1p[b]; MF.C2p[b] = (float)MB.C2p[b]; MF.C3l[b] = (float)MB.C3l[b];
    }
    return -1;
}

// RRM fast path: returns false unless the model has the shipped rrm-experimental layout (zodi_rrm.cuh):
// the eight component types in registry order, every component centred on the Sun, the three bands on
// one range with one temperature law.  M is the generic device form of the same descriptor.
inline bool build_rrm_model(const zodi_model_desc& d, const DevModel<double>& M, RrmModel<double>& R) {
    std::memset(&R, 0, sizeof(R));
    static const int kTypes[R_NCOMPS] = {ZODI_FAN, ZODI_COMET, ZODI_NARROW_BAND, ZODI_NARROW_BAND, ZODI_BROAD_BAND,
                                         ZODI_INTERSTELLAR, ZODI_RING_RRM, ZODI_FEATURE_RRM};
    if (d.kind != ZODI_RRM || d.n_comps != R_NCOMPS) return false;
    const zodi_component_desc* c = d.comps;
    for (int i = 0; i < R_NCOMPS; ++i) {
        if (c[i].type != kTypes[i]) return false;
        if (i != R_INTERSTELLAR && (c[i].x0[0] != 0.0 || c[i].x0[1] != 0.0 || c[i].x0[2] != 0.0)) return false;
    }
    for (int b = R_NB_OUT; b <= R_BROAD; ++b)
        if (c[b].cutoff_inner != c[R_NB_IN].cutoff_inner || c[b].cutoff_outer != c[R_NB_IN].cutoff_outer ||
            c[b].T_0 != c[R_NB_IN].T_0 || c[b].delta != c[R_NB_IN].delta)
            return false;
    R.n_nodes = d.n_nodes; R.n_temps = d.n_temps;
    const double dt = (d.temps[d.n_temps - 1] - d.temps[0]) / (d.n_temps - 1);
    R.t_ofs = -d.temps[0] / dt;
    R.t_top = d.n_temps - 1;
    R.e1 = d.calibration;
    for (int i = 0; i < R_NCOMPS; ++i) {
        R.t_scale[i] = c[i].T_0 / dt;
        R.mhd[i] = -0.5 * c[i].delta;
        R.c[i] = M.comps[i];
    }
    const DevComp<double>&a = R.c[R_NB_IN], &b = R.c[R_NB_OUT];
    R.nb_share_plane = (a.nx == b.nx && a.ny == b.ny && a.nz == b.nz);
    R.f_cos0 = std::cos(c[R_FEATURE].shape[4]);
    R.f_sin0 = std::sin(c[R_FEATURE].shape[4]);
    return true;
}

template <typename To, typename From>
inline void narrow_rrm(const RrmModel<From>& a, const DevModel<To>& m, RrmModel<To>& b) {
    std::memset(&b, 0, sizeof(b));
    b.n_nodes = a.n_nodes; b.n_temps = a.n_temps; b.nb_share_plane = a.nb_share_plane;
    b.t_ofs = (To)a.t_ofs; b.t_top = (To)a.t_top; b.e1 = (To)a.e1;
    for (int i = 0; i < R_NCOMPS; ++i) {
        b.t_scale[i] = (To)a.t_scale[i];
        b.mhd[i] = (To)a.mhd[i];
        b.c[i] = m.comps[i];  // already narrowed by narrow_model()
    }
    b.f_cos0 = a.f_cos0; b.f_sin0 = a.f_sin0;
}

// Packed fp32 form (zodi_rrm_x2.cuh): the narrowed RRM block plus the ring / feature constants in the
// KelsallModel layout the packed ring / feature loops read.  Returns false when ring and feature do not
// share one temperature law (then the scalar fused kernel is used).
inline bool build_rrm_x2(const RrmModel<double>& R64, const RrmModel<float>& R32, RrmModelX2& X) {
    std::memset(&X, 0, sizeof(X));
    if (R64.t_scale[R_RING] != R64.t_scale[R_FEATURE] || R64.mhd[R_RING] != R64.mhd[R_FEATURE]) return false;
    X.r = R32;
    KelsallModel<float>& K = X.rf;
    K.n_comps = 6; K.n_nodes = R64.n_nodes; K.n_temps = R64.n_temps;
    K.t_scale = (float)R64.t_scale[R_RING]; K.t_ofs = (float)R64.t_ofs; K.t_top = (float)R64.t_top;
    K.mhd = (float)R64.mhd[R_RING];
    const DevComp<double>&r = R64.c[R_RING], &f = R64.c[R_FEATURE];
    K.rnx = (float)r.nx; K.rny = (float)r.ny; K.rnz = (float)r.nz;
    K.r_R = (float)r.s[1]; K.r_c2 = (float)r.s[2]; K.r_c3 = (float)r.s[3];
    K.fnx = (float)f.nx; K.fny = (float)f.ny; K.fnz = (float)f.nz;
    K.f_R = (float)f.s[1]; K.f_c2 = (float)f.s[2]; K.f_c3 = (float)f.s[3]; K.f_theta0 = (float)f.s[4];
    K.f_c5 = (float)f.s[5];
    K.aB[4] = (float)(r.s[0] * R64.e1);  // A n_0 calibration
    K.aB[5] = (float)(f.s[0] * R64.e1);
    K.f_cos0 = R64.f_cos0; K.f_sin0 = R64.f_sin0;
    return true;
}

// Not-a-knot cubic spline through uniformly or non-uniformly spaced knots, one axis: the same
// linear system scipy.interpolate.CubicSpline (bc_type="not-a-knot", n >= 4) solves for the knot
// derivatives, then its PPoly coefficients c[k][i] (highest power first) on interval i.
inline void cubic_spline_not_a_knot(const std::vector<double>& x, const double* y, std::vector<double>& c0,
                                    std::vector<double>& c1, std::vector<double>& c2, std::vector<double>& c3) {
    const int n = (int)x.size();
    std::vector<double> dx(n - 1), slope(n - 1);
    for (int i = 0; i < n - 1; ++i) { dx[i] = x[i + 1] - x[i]; slope[i] = (y[i + 1] - y[i]) / dx[i]; }
    // tridiagonal system lo[i] s[i-1] + di[i] s[i] + up[i] s[i+1] = b[i]
    std::vector<double> lo(n, 0.0), di(n, 0.0), up(n, 0.0), b(n, 0.0), s(n, 0.0);
    for (int i = 1; i < n - 1; ++i) {
        lo[i] = dx[i];
        di[i] = 2.0 * (dx[i - 1] + dx[i]);
        up[i] = dx[i - 1];
        b[i] = 3.0 * (dx[i] * slope[i - 1] + dx[i - 1] * slope[i]);
    }
    {   // not-a-knot at the first interior knot
        const double d = x[2] - x[0];
        di[0] = dx[1];
        up[0] = d;
        b[0] = ((dx[0] + 2.0 * d) * dx[1] * slope[0] + dx[0] * dx[0] * slope[1]) / d;
    }
    {   // not-a-knot at the last interior knot
        const double d = x[n - 1] - x[n - 3];
        di[n - 1] = dx[n - 3];
        lo[n - 1] = d;
        b[n - 1] = (dx[n - 2] * dx[n - 2] * slope[n - 3] + (2.0 * d + dx[n - 2]) * dx[n - 3] * slope[n - 2]) / d;
    }
    // Thomas algorithm with partial safety (the matrix is diagonally dominant in the interior)
    std::vector<double> cp(n, 0.0), dp(n, 0.0);
    cp[0] = up[0] / di[0];
    dp[0] = b[0] / di[0];
    for (int i = 1; i < n; ++i) {
        const double m = di[i] - lo[i] * cp[i - 1];
        cp[i] = up[i] / m;
        dp[i] = (b[i] - lo[i] * dp[i - 1]) / m;
    }
    s[n - 1] = dp[n - 1];
    for (int i = n - 2; i >= 0; --i) s[i] = dp[i] - cp[i] * s[i + 1];
    c0.resize(n - 1); c1.resize(n - 1); c2.resize(n - 1); c3.resize(n - 1);
    for (int i = 0; i < n - 1; ++i) {
        const double t = (s[i] + s[i + 1] - 2.0 * slope[i]) / dx[i];
        c0[i] = t / dx[i];
        c1[i] = (slope[i] - s[i]) / dx[i] - t;
        c2[i] = s[i];
        c3[i] = y[i];
    }
}

}  // namespace zodi
