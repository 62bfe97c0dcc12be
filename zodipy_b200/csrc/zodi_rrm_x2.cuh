// zodi_rrm_x2.cuh - packed-fp32 form of the fused RRM integrator (zodi_rrm.cuh): every thread works on
// TWO lines of sight and keeps each per-node quantity as a register pair, so the multiply / add / fma work
// issues as FFMA2 / FMUL2 / FADD2 (one issue slot for two operations).  The scalar fused RRM kernel is
// bound by issue slots (88 % issue utilisation, XU 66 %, profiles/r2_ncu_rrm_nside512.md); transcendentals,
// compares and selects stay scalar per half.  Same operations per line of sight as integrate_rrm<float>.
// Ring and feature reuse the packed Kelsall loops (number_density.py:345-404: A x the Kelsall densities)
// through a KelsallModel<float> that carries their constants.
#pragma once

#include "zodi_kelsall_x2.cuh"
#include "zodi_rrm.cuh"

namespace zodi {

struct RrmModelX2 {
    RrmModel<float> r;
    KelsallModel<float> rf;  // ring / feature constants in the layout kelsall_ring_x2 / kelsall_feature_x2 read
};

// Intervals of the four grids that are not in LosPre (which carries ring / feature and the directions).
struct RrmIntervals { float h_fan, mid_fan, h_comet, mid_comet, h_bands, mid_bands, h_is, mid_is; };

ZODI_HD void rrm_pre(const RrmModelX2& M, double ux, double uy, double uz, double ox, double oy, double oz, double ex,
                     double ey, uint32_t mask, LosPre& P, RrmIntervals& I) {
    const LosGeometry<float> G = los_geometry<float>(ux, uy, uz, ox, oy, oz);
    P.ux = G.ux; P.uy = G.uy; P.uz = G.uz; P.ox = G.ox; P.oy = G.oy; P.oz = G.oz;
    P.hA = P.midA = 0.f;
    auto interval = [&](int slot, float& h, float& mid) {
        los_interval<float>(G, M.r.c[slot].cut_in, M.r.c[slot].cut_out, (mask >> (2 * slot)) & 1u,
                            (mask >> (2 * slot + 1)) & 1u, h, mid);
    };
    interval(R_FAN, I.h_fan, I.mid_fan);
    interval(R_COMET, I.h_comet, I.mid_comet);
    interval(R_NB_IN, I.h_bands, I.mid_bands);
    interval(R_INTERSTELLAR, I.h_is, I.mid_is);
    interval(R_RING, P.hR, P.midR);
    interval(R_FEATURE, P.hF, P.midF);
    feature_rotation<float>(ex, ey, M.r.f_cos0, M.r.f_sin0, P.cr, P.sr);
}

struct RrmNode2 { F2 xh, yh, zh, R2, lgR2, wB; };

struct RrmLos2 { F2 ux, uy, uz, ox, oy, oz; };

ZODI_HD RrmNode2 rrm_node2(const RrmModel<float>& R, int slot, TableRef tref, Pair<float> nw, F2 h, F2 mid,
                           const RrmLos2& G) {
    RrmNode2 s;
    const F2 R_los = fma2(h, nw.a, mid);
    s.xh = fma2(R_los, G.ux, G.ox);
    s.yh = fma2(R_los, G.uy, G.oy);
    s.zh = fma2(R_los, G.uz, G.oz);
    s.R2 = fma2(s.xh, s.xh, fma2(s.yh, s.yh, mul2(s.zh, s.zh)));
    s.lgR2 = lg2_2(s.R2);
    const F2 t = fma2(ex2_2(mul2(s.lgR2, R.mhd[slot])), R.t_scale[slot], R.t_ofs);
    s.wB = mul2(table_at2(tref, t, R.t_top), nw.b);
    return s;
}

ZODI_HD F2 abs2(F2 v) { return f2(fabsf(v.x), fabsf(v.y)); }

// asin(Z_c / R_c) for both halves (rrm_latitude<float>).
ZODI_HD F2 rrm_latitude2(const DevComp<float>& c, const RrmNode2& s, F2 rinv, F2& Zc) {
    Zc = fma2(s.xh, c.nx, fma2(s.yh, c.ny, mul2(s.zh, c.nz)));
    return asin_unit2(clamp1_2(mul2(Zc, rinv)));
}

// rrm_fan_like<float, WITH_Q> for both halves.
template <bool WITH_Q>
ZODI_HD F2 rrm_fan_like2(const DevComp<float>& c, const RrmNode2& s) {
    const bool in_x = s.R2.x >= c.s[4] && s.R2.x <= c.s[5], in_y = s.R2.y >= c.s[4] && s.R2.y <= c.s[5];
    F2 Zc;
    const F2 beta = rrm_latitude2(c, s, rsq_2(s.R2), Zc);
    const F2 za = abs2(Zc), ab = abs2(beta);
    F2 bp = ab;
    const bool slab_x = za.x < c.s[6], slab_y = za.y < c.s[6];
    if (warp_any(slab_x || slab_y)) {
        const F2 pw = ex2_2(mul2(fma2(za, -c.s[1], 2.0f), lg2_2(ab)));
        bp = f2(slab_x ? (ab.x > 0.0f ? pw.x : 0.0f) : ab.x, slab_y ? (ab.y > 0.0f ? pw.y : 0.0f) : ab.y);
    }
    F2 lg = mul2(f2(Math<float>::sin_(bp.x), Math<float>::sin_(bp.y)), c.s[2]);
    if (WITH_Q) {
        const F2 px = fma2(Zc, -c.nx, s.xh), py = fma2(Zc, -c.ny, s.yh), pz = fma2(Zc, -c.nz, s.zh);
        const F2 rho2 = fma2(px, px, fma2(py, py, mul2(pz, pz)));
        lg = fma2(fma2(s.lgR2, -1.0f, lg2_2(rho2)), 0.5f * c.s[7], lg);
    }
    const F2 n = mul2(ex2_2(fma2(s.lgR2, c.s[0], lg)), c.s[3]);
    return f2(in_x ? n.x : 0.0f, in_y ? n.y : 0.0f);
}

// rrm_narrow<float> for both halves; lat = |latitude| in degrees.
ZODI_HD F2 rrm_narrow2(const DevComp<float>& c, const RrmNode2& s, F2 lat) {
    const bool in_x = (s.R2.x >= c.s[4] && s.R2.x <= c.s[5]) && (lat.x < c.s[0]);
    const bool in_y = (s.R2.y >= c.s[4] && s.R2.y <= c.s[5]) && (lat.y < c.s[0]);
    F2 n = f2(0.0f);
    if (warp_any(in_x || in_y)) {
        const F2 v = mul2(ex2_2(fma2(s.lgR2, c.s[2], mul2(add2(lat, -c.s[0]), c.s[1]))), c.s[3]);
        n = f2(in_x ? v.x : 0.0f, in_y ? v.y : 0.0f);
    }
    return n;
}

// rrm_broad<float> for both halves; lat = signed latitude in degrees.
ZODI_HD F2 rrm_broad2(const DevComp<float>& c, const RrmNode2& s, F2 lat, F2 rinv) {
    const bool in_x = s.R2.x >= c.s[4] && s.R2.x <= c.s[5], in_y = s.R2.y >= c.s[4] && s.R2.y <= c.s[5];
    const F2 a = mul2(add2(lat, -c.s[0]), c.s[1]), b = mul2(add2(lat, c.s[0]), c.s[1]);
    const F2 f = add2(ex2_2(mul2(mul2(a, a), c.s[6])), ex2_2(mul2(mul2(b, b), c.s[6])));
    const F2 rp = (c.s[7] != 0.0f) ? rinv : ex2_2(mul2(s.lgR2, c.s[2]));
    const F2 n = mul2(mul2(f, c.s[3]), rp);
    return f2(in_x ? n.x : 0.0f, in_y ? n.y : 0.0f);
}

// All eight components of two lines of sight; emit(ci, value_a, value_b) in model order.
template <typename Emit>
ZODI_HD void integrate_rrm_x2(const RrmModelX2& M, const Pair<float>* tab, const Pair<float>* nodes, const LosPre& Pa,
                              const LosPre& Pb, const RrmIntervals& Ia, const RrmIntervals& Ib, Emit emit) {
    const RrmModel<float>& R = M.r;
    const RrmLos2 G = {f2(Pa.ux, Pb.ux), f2(Pa.uy, Pb.uy), f2(Pa.uz, Pb.uz),
                       f2(Pa.ox, Pb.ox), f2(Pa.oy, Pb.oy), f2(Pa.oz, Pb.oz)};
    const TableRef tref = table_ref(tab);
    const float deg = 57.295779513082323f;

    {   // fan
        const F2 h = f2(Ia.h_fan, Ib.h_fan), mid = f2(Ia.mid_fan, Ib.mid_fan);
        F2 acc = f2(0.f);
        for (int k = 0; k < R.n_nodes; ++k) {
            const RrmNode2 s = rrm_node2(R, R_FAN, tref, nodes[k], h, mid, G);
            acc = fma2(s.wB, rrm_fan_like2<true>(R.c[R_FAN], s), acc);
        }
        const F2 v = mul2(acc, mul2(h, R.e1));
        emit(R_FAN, v.x, v.y);
    }
    {   // comet
        const F2 h = f2(Ia.h_comet, Ib.h_comet), mid = f2(Ia.mid_comet, Ib.mid_comet);
        F2 acc = f2(0.f);
        for (int k = 0; k < R.n_nodes; ++k) {
            const RrmNode2 s = rrm_node2(R, R_COMET, tref, nodes[k], h, mid, G);
            acc = fma2(s.wB, rrm_fan_like2<false>(R.c[R_COMET], s), acc);
        }
        const F2 v = mul2(acc, mul2(h, R.e1));
        emit(R_COMET, v.x, v.y);
    }
    {   // the three asteroidal bands on one grid
        const F2 h = f2(Ia.h_bands, Ib.h_bands), mid = f2(Ia.mid_bands, Ib.mid_bands);
        F2 a_in = f2(0.f), a_out = f2(0.f), a_bb = f2(0.f);
        for (int k = 0; k < R.n_nodes; ++k) {
            const RrmNode2 s = rrm_node2(R, R_NB_IN, tref, nodes[k], h, mid, G);
            const F2 rinv = rsq_2(s.R2);
            F2 Zc;
            const F2 lat_in = mul2(abs2(rrm_latitude2(R.c[R_NB_IN], s, rinv, Zc)), deg);
            const F2 lat_out = R.nb_share_plane ? lat_in : mul2(abs2(rrm_latitude2(R.c[R_NB_OUT], s, rinv, Zc)), deg);
            const F2 lat_bb = mul2(rrm_latitude2(R.c[R_BROAD], s, rinv, Zc), deg);
            a_in = fma2(s.wB, rrm_narrow2(R.c[R_NB_IN], s, lat_in), a_in);
            a_out = fma2(s.wB, rrm_narrow2(R.c[R_NB_OUT], s, lat_out), a_out);
            a_bb = fma2(s.wB, rrm_broad2(R.c[R_BROAD], s, lat_bb, rinv), a_bb);
        }
        const F2 sc = mul2(h, R.e1);
        const F2 v_in = mul2(a_in, sc), v_out = mul2(a_out, sc), v_bb = mul2(a_bb, sc);
        emit(R_NB_IN, v_in.x, v_in.y);
        emit(R_NB_OUT, v_out.x, v_out.y);
        emit(R_BROAD, v_bb.x, v_bb.y);
    }
    {   // interstellar
        const F2 h = f2(Ia.h_is, Ib.h_is), mid = f2(Ia.mid_is, Ib.mid_is);
        F2 acc = f2(0.f);
        for (int k = 0; k < R.n_nodes; ++k) acc = add2(acc, rrm_node2(R, R_INTERSTELLAR, tref, nodes[k], h, mid, G).wB);
        const F2 v = mul2(acc, mul2(h, R.c[R_INTERSTELLAR].s[0] * R.e1));
        emit(R_INTERSTELLAR, v.x, v.y);
    }
    kelsall_ring_x2<false, 1>(M.rf, tab, nodes, Pa, Pb, 0, [&](float a, float b) { emit(R_RING, a, b); });
    kelsall_feature_x2<false, 1>(M.rf, tab, nodes, Pa, Pb, 0, [&](float a, float b) { emit(R_FEATURE, a, b); });
}

}  // namespace zodi
