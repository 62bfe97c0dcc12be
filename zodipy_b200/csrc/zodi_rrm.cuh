// zodi_rrm.cuh - fused integrator for the RRM model layout (rrm-experimental).
//
// The reference integrates each of the eight RRM components on its own quadrature grid and, per node
// and component, repeats position, heliocentric distance, grain temperature and the blackbody
// interpolation (zodipy/brightness.py:59-83) before calling the density (zodipy/number_density.py:
// 184-404).  The shipped layout (zodipy/component_params.py:102-206, zodipy/model_registry.py:64-73)
//     fan, comet, inner narrow band, outer narrow band, broad band, interstellar, ring, feature
// has every component centred on the Sun, so the density's own distance IS the heliocentric distance
// of the source function (one log2 R^2 serves the temperature law, the radial power laws and the
// cos^Q term), and the three asteroidal bands share one range (cutoffs (1.5237, 3.137) AU,
// zodipy/line_of_sight.py:28-48): they are integrated on ONE grid with shared position / temperature /
// table lookup / 1/R, the two narrow bands also sharing their latitude when their symmetry planes
// coincide.  Six node loops instead of eight, no per-node dispatch on the density type, ~2.5x fewer
// instructions per evaluation than the generic kernel.  Models that do not have this layout
// (user-edited component lists, off-centre components) keep the generic kernel; build_rrm_model()
// decides per model.
#pragma once

#include "zodi_device.cuh"
#include "zodi_kelsall.cuh"

namespace zodi {

enum RrmSlot : int { R_FAN = 0, R_COMET = 1, R_NB_IN = 2, R_NB_OUT = 3, R_BROAD = 4, R_INTERSTELLAR = 5,
                     R_RING = 6, R_FEATURE = 7, R_NCOMPS = 8 };

template <typename Real>
struct RrmModel {
    int n_nodes;
    int n_temps;
    int nb_share_plane;   // inner / outer narrow band have the same symmetry plane: one latitude serves both
    Real t_ofs, t_top;    // table coordinate t = t_scale * 2^(mhd * log2 R^2) + t_ofs
    Real t_scale[R_NCOMPS], mhd[R_NCOMPS];  // per component (unpack_model.py:123-126); equal inside the band group
    Real e1;              // calibration (brightness.py:83)
    DevComp<Real> c[R_NCOMPS];  // density constants as derive_component() lays them out
    double f_cos0, f_sin0;      // cos / sin of the feature's theta_0 (feature_rotation)
};

// Shared per-node source terms: position, R^2, log2 R^2, w * B(T).
template <typename Real>
struct RrmNode { Real xh, yh, zh, R2, lgR2, wB; };

template <typename Real>
ZODI_HD RrmNode<Real> rrm_node(const RrmModel<Real>& R, int slot, const Pair<Real>* tab, Pair<Real> nw, Real h,
                               Real mid, const LosGeometry<Real>& G) {
    using M = Math<Real>;
    RrmNode<Real> s;
    const Real R_los = M::fma_(h, nw.a, mid);                      // brightness.py:74
    s.xh = M::fma_(R_los, G.ux, G.ox);
    s.yh = M::fma_(R_los, G.uy, G.oy);
    s.zh = M::fma_(R_los, G.uz, G.oz);
    s.R2 = M::fma_(s.xh, s.xh, M::fma_(s.yh, s.yh, s.zh * s.zh));
    s.lgR2 = M::log2_(s.R2);
    const Real t = M::fma_(R.t_scale[slot], M::exp2_bounded_(R.mhd[slot] * s.lgR2), R.t_ofs);  // blackbody.py:30
    s.wB = nw.b * table_at<Real>(tab, t, R.t_top);                  // brightness.py:81
    return s;
}

// Node k of a rule split over L lanes with a warp-uniform trip count: surplus iterations re-evaluate the
// last node with weight 0 (loops whose bodies contain warp votes).
template <typename Real>
ZODI_HD Pair<Real> rrm_lane_node(const Pair<Real>* nodes, int n_nodes, int k) {
    Pair<Real> nw = nodes[k < n_nodes ? k : n_nodes - 1];
    if (k >= n_nodes) nw.b = Real(0);
    return nw;
}

// Latitude of the node above the component's symmetry plane: asin(Z_c / R_c), clipped like the
// reference's arcsin argument can only be by rounding.
template <typename Real>
ZODI_HD Real rrm_latitude(const DevComp<Real>& c, const RrmNode<Real>& s, Real rinv, Real& Zc) {
    using M = Math<Real>;
    Zc = M::fma_(s.xh, c.nx, M::fma_(s.yh, c.ny, s.zh * c.nz));
    return M::asin_(M::max_(Real(-1), M::min_(Real(1), Zc * rinv)));
}

// Fan (WITH_Q) / comet density, number_density.py:184-256; constants c.s[] as in density<Real>().
template <typename Real, bool WITH_Q>
ZODI_HD Real rrm_fan_like(const DevComp<Real>& c, const RrmNode<Real>& s) {
    using M = Math<Real>;
    // outside [R_inner, R_outer] the density is 0 (boolean-mask gather / scatter of the reference); the value
    // is selected at the end so that every lane reaches the warp vote below
    const bool in_range = s.R2 >= c.s[4] && s.R2 <= c.s[5];
    Real Zc;
    const Real beta = rrm_latitude<Real>(c, s, M::rsqrt_(s.R2), Zc);
    const Real za = M::abs_(Zc);
    const Real ab = M::abs_(beta);
    // |beta|^epsilon with epsilon = 2 - |Z_c| / Z_0 inside the slab |Z_c| < Z_0 and 1 outside it, where the
    // power is the identity (np.power(x, 1.0) == x): the log2 / exp2 pair runs only for warps with a lane
    // inside the slab (a few per cent of the nodes)
    Real bp = ab;
    const bool in_slab = za < c.s[6];
    if (warp_any(in_slab)) {
        const Real pw = (ab > Real(0)) ? M::exp2_((Real(2) - za * c.s[1]) * M::log2_(ab)) : Real(0);
        bp = in_slab ? pw : ab;
    }
    Real lg = c.s[2] * M::sin_(bp);
    if (WITH_Q) {
        // cos(beta)^Q = (rho / R)^Q with rho the in-plane distance (no cancellation near the poles)
        const Real px = M::fma_(-Zc, c.nx, s.xh), py = M::fma_(-Zc, c.ny, s.yh), pz = M::fma_(-Zc, c.nz, s.zh);
        const Real rho2 = M::fma_(px, px, M::fma_(py, py, pz * pz));
        lg = M::fma_(Real(0.5) * c.s[7], M::log2_(rho2) - s.lgR2, lg);
    }
    return in_range ? c.s[3] * M::exp2_(M::fma_(c.s[0], s.lgR2, lg)) : Real(0);
}

// Narrow band given the latitude in degrees (absolute value), number_density.py:267-304.
template <typename Real>
ZODI_HD Real rrm_narrow(const DevComp<Real>& c, const RrmNode<Real>& s, Real abs_lat_deg) {
    using M = Math<Real>;
    const bool inside = (s.R2 >= c.s[4] && s.R2 <= c.s[5]) && (abs_lat_deg < c.s[0]);
    Real n = Real(0);
    // the band is a few degrees wide: most warps have no lane inside it and skip the exponential
    if (warp_any(inside)) n = inside ? c.s[3] * M::exp2_(M::fma_(c.s[2], s.lgR2, c.s[1] * (abs_lat_deg - c.s[0]))) : Real(0);
    return n;
}

// Broad band given the signed latitude in degrees, number_density.py:307-342.
template <typename Real>
ZODI_HD Real rrm_broad(const DevComp<Real>& c, const RrmNode<Real>& s, Real lat_deg, Real rinv) {
    using M = Math<Real>;
    if (!(s.R2 >= c.s[4] && s.R2 <= c.s[5])) return Real(0);  // no vote below: lanes may leave early
    const Real a = (lat_deg - c.s[0]) * c.s[1], b = (lat_deg + c.s[0]) * c.s[1];
    const Real f = M::exp2_(c.s[6] * a * a) + M::exp2_(c.s[6] * b * b);
    // (R / R_outer)^-gamma: gamma == 1 in the shipped model, where it is 1 / R itself
    const Real rp = (c.s[7] != Real(0)) ? rinv : M::exp2_(c.s[2] * s.lgR2);
    return c.s[3] * f * rp;
}

template <typename Real, typename Emit>
ZODI_HD void integrate_rrm(const RrmModel<Real>& R, const Pair<Real>* tab, const Pair<Real>* nodes, double dux,
                           double duy, double duz, double dox, double doy, double doz, double dex, double dey,
                           uint32_t outside_mask, int sub, int L, Emit emit) {
    using M = Math<Real>;
    const LosGeometry<Real> G = los_geometry<Real>(dux, duy, duz, dox, doy, doz);
    auto interval = [&](int slot, Real& h, Real& mid) {
        los_interval<Real>(G, R.c[slot].cut_in, R.c[slot].cut_out, (outside_mask >> (2 * slot)) & 1u,
                           (outside_mask >> (2 * slot + 1)) & 1u, h, mid);
    };
    const Real deg = Real(180.0 / kPi);
    Real h, mid;

    // ---- fan, comet: own grids ----
    interval(R_FAN, h, mid);
    Real acc = Real(0);
    for (int k0 = 0; k0 < R.n_nodes; k0 += L) {  // warp-uniform trip count: the densities contain warp votes
        const RrmNode<Real> s = rrm_node<Real>(R, R_FAN, tab, rrm_lane_node<Real>(nodes, R.n_nodes, k0 + sub), h, mid, G);
        acc = M::fma_(s.wB, rrm_fan_like<Real, true>(R.c[R_FAN], s), acc);
    }
    emit(R_FAN, acc * (R.e1 * h));

    interval(R_COMET, h, mid);
    acc = Real(0);
    for (int k0 = 0; k0 < R.n_nodes; k0 += L) {
        const RrmNode<Real> s = rrm_node<Real>(R, R_COMET, tab, rrm_lane_node<Real>(nodes, R.n_nodes, k0 + sub), h, mid, G);
        acc = M::fma_(s.wB, rrm_fan_like<Real, false>(R.c[R_COMET], s), acc);
    }
    emit(R_COMET, acc * (R.e1 * h));

    // ---- the three asteroidal bands: one grid, shared source terms and 1/R ----
    interval(R_NB_IN, h, mid);
    Real a_in = Real(0), a_out = Real(0), a_bb = Real(0);
    for (int k0 = 0; k0 < R.n_nodes; k0 += L) {
        const RrmNode<Real> s = rrm_node<Real>(R, R_NB_IN, tab, rrm_lane_node<Real>(nodes, R.n_nodes, k0 + sub), h, mid, G);
        const Real rinv = M::rsqrt_(s.R2);
        Real Zc;
        const Real lat_in = M::abs_(rrm_latitude<Real>(R.c[R_NB_IN], s, rinv, Zc)) * deg;
        const Real lat_out = R.nb_share_plane ? lat_in : M::abs_(rrm_latitude<Real>(R.c[R_NB_OUT], s, rinv, Zc)) * deg;
        const Real lat_bb = rrm_latitude<Real>(R.c[R_BROAD], s, rinv, Zc) * deg;
        a_in = M::fma_(s.wB, rrm_narrow<Real>(R.c[R_NB_IN], s, lat_in), a_in);
        a_out = M::fma_(s.wB, rrm_narrow<Real>(R.c[R_NB_OUT], s, lat_out), a_out);
        a_bb = M::fma_(s.wB, rrm_broad<Real>(R.c[R_BROAD], s, lat_bb, rinv), a_bb);
    }
    emit(R_NB_IN, a_in * (R.e1 * h));
    emit(R_NB_OUT, a_out * (R.e1 * h));
    emit(R_BROAD, a_bb * (R.e1 * h));

    // ---- interstellar: constant density (number_density.py:259-264), own temperature law ----
    interval(R_INTERSTELLAR, h, mid);
    acc = Real(0);
    for (int k = sub; k < R.n_nodes; k += L) acc += rrm_node<Real>(R, R_INTERSTELLAR, tab, nodes[k], h, mid, G).wB;
    emit(R_INTERSTELLAR, acc * (R.c[R_INTERSTELLAR].s[0] * R.e1 * h));

    // ---- circumsolar ring and Earth-trailing feature (number_density.py:345-404 = A x the Kelsall ones) ----
    interval(R_RING, h, mid);
    acc = Real(0);
    {
        const DevComp<Real>& c = R.c[R_RING];
        for (int k = sub; k < R.n_nodes; k += L) {
            const RrmNode<Real> s = rrm_node<Real>(R, R_RING, tab, nodes[k], h, mid, G);
            const Real d = M::sqrt_(s.R2) - c.s[1];
            const Real Zc = M::fma_(s.xh, c.nx, M::fma_(s.yh, c.ny, s.zh * c.nz));
            acc = M::fma_(s.wB, M::exp2_neg_(-M::fma_(d * d, c.s[2], M::abs_(Zc) * c.s[3])), acc);
        }
        emit(R_RING, acc * (c.s[0] * R.e1 * h));
    }

    interval(R_FEATURE, h, mid);
    acc = Real(0);
    {
        const DevComp<Real>& c = R.c[R_FEATURE];
        // rotate by -(theta_earth + theta_0): atan2 then returns the wrapped longitude offset (see kelsall_feature)
        Real cr, sr;
        feature_rotation<Real>(dex, dey, R.f_cos0, R.f_sin0, cr, sr);
        for (int k = sub; k < R.n_nodes; k += L) {
            const RrmNode<Real> s = rrm_node<Real>(R, R_FEATURE, tab, nodes[k], h, mid, G);
            const Real d = M::sqrt_(s.R2) - c.s[1];
            const Real Zc = M::fma_(s.xh, c.nx, M::fma_(s.yh, c.ny, s.zh * c.nz));
            const Real xr = M::fma_(s.xh, cr, s.yh * sr), yr = M::fma_(s.yh, cr, -(s.xh * sr));
            const Real dth = M::atan2_abs_(yr, xr);  // only dth^2 is used
            const Real e = M::fma_(d * d, c.s[2], M::fma_(M::abs_(Zc), c.s[3], dth * dth * c.s[5]));
            acc = M::fma_(s.wB, M::exp2_neg_(-e), acc);
        }
        emit(R_FEATURE, acc * (c.s[0] * R.e1 * h));
    }
}

}  // namespace zodi
