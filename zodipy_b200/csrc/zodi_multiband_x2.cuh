// zodi_multiband_x2.cuh - packed-fp32 form of the multi-band integrator (zodi_multiband.cuh): every thread
// works on TWO lines of sight and keeps each per-node quantity and each band accumulator as a register pair
// (FFMA2 / FMUL2 / FADD2, see zodi_kelsall_x2.cuh).  The work shared by all bands - positions, grain
// temperature, table row, every number density (zodipy/number_density.py:47-181) - is the packed Kelsall
// arithmetic; per band and node pair remain one 16-byte table read per half and two bands, two scalar FMAs
// for the interpolated B_nu (zodipy/brightness.py:48), the weighted density sum and one accumulate.
//
// Differences from the scalar multi-band kernel, all exact or within fp32 rounding:
//   * bands past n_bands are skipped by a uniform branch instead of integrating a zero table;
//   * asteroidal-band densities a whole warp skips (exp(-s^6) == 0, the pretest of kelsall_group_a_x2) drop
//     out of the weighted sum as well; the scattering term is formed only for bands with a non-zero albedo;
//   * the quadrature weight and the interval half-width multiply the densities once per node, so ONE
//     accumulator per band serves the three node loops (cloud + bands, ring, feature).
// Table layout in shared memory: row = temperature knot, NB + 2 (a, delta) pairs per row - the two lines of
// sight of a thread, and the lanes of a warp, read different rows; the padding spreads rows over the banks.
#pragma once

#include "zodi_kelsall_x2.cuh"
#include "zodi_multiband.cuh"

namespace zodi {

template <int NB>
struct MbRows { static constexpr int kRow = NB + 2; };  // Pair<float> per table row (16-byte aligned rows)

struct alignas(16) BandPair { Pair<float> p, q; };  // table entries of bands 2i, 2i + 1 at one knot

// Node quantities shared by all bands, both halves.
struct MbNode2 {
    F2 xh, yh, zh, Rh2, frac;
    int ix, iy;      // table rows
    F2 th, rh2inv;   // scattering angle and 1 / R_h^2 (SCATTER)
};

template <bool SCATTER>
ZODI_HD MbNode2 mb_node2(const KelsallModel<float>& K, F2 R_los, F2 ux, F2 uy, F2 uz, F2 ox, F2 oy, F2 oz) {
    MbNode2 s;
    s.xh = fma2(R_los, ux, ox);
    s.yh = fma2(R_los, uy, oy);
    s.zh = fma2(R_los, uz, oz);
    s.Rh2 = fma2(s.xh, s.xh, fma2(s.yh, s.yh, mul2(s.zh, s.zh)));
    F2 t = fma2(ex2_2(mul2(lg2_2(s.Rh2), K.mhd)), K.t_scale, K.t_ofs);
    // table_coord<float> for both halves
    t = f2(fminf(fmaxf(t.x, 0.0f), K.t_top), fminf(fmaxf(t.y, 0.0f), K.t_top));
    const float magic = 12582912.0f;
    const F2 r = add2(add2(t, -0.5f), magic);
#if defined(__CUDA_ARCH__)
    s.ix = __float_as_int(r.x) - 0x4B400000;
    s.iy = __float_as_int(r.y) - 0x4B400000;
#else
    int b0, b1;
    memcpy(&b0, &r.x, 4);
    memcpy(&b1, &r.y, 4);
    s.ix = b0 - 0x4B400000;
    s.iy = b1 - 0x4B400000;
#endif
    s.frac = fma2(add2(r, -magic), -1.0f, t);
    s.th = f2(0.f);
    s.rh2inv = f2(0.f);
    if (SCATTER) {
        const F2 rh_inv = rsq_2(s.Rh2);
        const F2 ct = clamp1_2(mul2(fma2(ux, s.xh, fma2(uy, s.yh, mul2(uz, s.zh))), rh_inv));
        s.th = add2(asin_unit2(ct), 1.57079637f);  // acos(-ct)
        s.rh2inv = mul2(rh_inv, rh_inv);
    }
    return s;
}

// acc[b] += B_b * sum_c aB[b][c0 + c] wn[c]  +  F_b * sum_c aS[b][c0 + c] wn[c]   for the NC densities wn
// (already multiplied by quadrature weight x interval half-width).
template <int NB, int NC, bool SCATTER>
ZODI_HD void mb_accumulate2(const MultiBandModel<float>& MB, const Pair<float>* rows, const MbNode2& s,
                            const F2 (&wn)[NC], int c0, F2 (&acc)[NB]) {
    const BandPair* rx = reinterpret_cast<const BandPair*>(rows + s.ix * MbRows<NB>::kRow);
    const BandPair* ry = reinterpret_cast<const BandPair*>(rows + s.iy * MbRows<NB>::kRow);
    const int nb = MB.n_bands;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b2 = 0; b2 < NB / 2; ++b2) {
        if (2 * b2 >= nb) break;  // uniform
        const BandPair ex = rx[b2], ey = ry[b2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 2; ++q) {
            const int b = 2 * b2 + q;
            if (b >= nb) break;  // uniform
            const Pair<float> tx = q ? ex.q : ex.p, ty = q ? ey.q : ey.p;
            const F2 B = f2(fmaf(tx.b, s.frac.x, tx.a), fmaf(ty.b, s.frac.y, ty.a));
            F2 sB = mul2(wn[0], MB.aB[b][c0]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int c = 1; c < NC; ++c) sB = fma2(wn[c], MB.aB[b][c0 + c], sB);
            acc[b] = fma2(B, sB, acc[b]);
            if (SCATTER && ((MB.scatter_bands >> b) & 1u)) {  // uniform
                F2 sS = mul2(wn[0], MB.aS[b][c0]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int c = 1; c < NC; ++c) sS = fma2(wn[c], MB.aS[b][c0 + c], sS);
                const F2 F = mul2(add2(fma2(s.th, MB.C2p[b], MB.C1p[b]), ex2_2(mul2(s.th, MB.C3l[b]))), s.rh2inv);
                acc[b] = fma2(F, sS, acc[b]);
            }
        }
    }
}

// exp(-s^6) (1 + s^4 / v) * (rinv * rad) for both halves, plane distance dot = n . X given.
ZODI_HD F2 band_density2_dot(F2 dot, F2 rinv, F2 rinv_rad, float c3) {
    const F2 sz = mul2(dot, rinv);
    const F2 s2 = mul2(sz, sz), s4 = mul2(s2, s2), s6 = mul2(s4, s2);
    return mul2(mul2(ex2_neg2(s6), fma2(s4, c3, 1.0f)), rinv_rad);
}

// Two lines of sight, NB bands; emit(b, value_a, value_b) receives band b's component-summed emission.
template <int NB, bool HAS_RF, bool SCATTER, typename Emit>
ZODI_HD void integrate_multiband_x2(const MultiBandModel<float>& MB, const Pair<float>* rows, const Pair<float>* nodes,
                                    const LosPre& Pa, const LosPre& Pb, Emit emit) {
    const KelsallModel<float>& K = MB.base;
    const F2 ux = f2(Pa.ux, Pb.ux), uy = f2(Pa.uy, Pb.uy), uz = f2(Pa.uz, Pb.uz);
    const F2 ox = f2(Pa.ox, Pb.ox), oy = f2(Pa.oy, Pb.oy), oz = f2(Pa.oz, Pb.oz);
    F2 acc[NB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < NB; ++b) acc[b] = f2(0.f);

    if (HAS_RF) {
        {   // ring (own grid; number_density.py:113-139)
            const F2 h = f2(Pa.hR, Pb.hR), mid = f2(Pa.midR, Pb.midR);
            for (int k = 0; k < K.n_nodes; ++k) {
                const Pair<float> nw = nodes[k];
                const MbNode2 s = mb_node2<SCATTER>(K, fma2(h, nw.a, mid), ux, uy, uz, ox, oy, oz);
                const F2 d = add2(sqrt_2(s.Rh2), -K.r_R);
                const F2 Zc = fma2(s.xh, K.rnx, fma2(s.yh, K.rny, mul2(s.zh, K.rnz)));
                const F2 n = ex2_2(fma2(mul2(d, d), K.r_c2, mul2(f2(fabsf(Zc.x), fabsf(Zc.y)), K.r_c3)));
                const F2 wn[1] = {mul2(n, mul2(h, nw.b))};
                mb_accumulate2<NB, 1, SCATTER>(MB, rows, s, wn, 4, acc);
            }
        }
        {   // feature (own grid; number_density.py:142-181)
            const F2 h = f2(Pa.hF, Pb.hF), mid = f2(Pa.midF, Pb.midF);
            const F2 cr = f2(Pa.cr, Pb.cr), sr = f2(Pa.sr, Pb.sr), msr = f2(-Pa.sr, -Pb.sr);
            for (int k = 0; k < K.n_nodes; ++k) {
                const Pair<float> nw = nodes[k];
                const MbNode2 s = mb_node2<SCATTER>(K, fma2(h, nw.a, mid), ux, uy, uz, ox, oy, oz);
                const F2 d = add2(sqrt_2(s.Rh2), -K.f_R);
                const F2 Zc = fma2(s.xh, K.fnx, fma2(s.yh, K.fny, mul2(s.zh, K.fnz)));
                const F2 xr = fma2(s.xh, cr, mul2(s.yh, sr)), yr = fma2(s.yh, cr, mul2(s.xh, msr));
                const F2 dth = atan2_abs2(yr, xr);
                const F2 e = fma2(mul2(d, d), K.f_c2,
                                  fma2(f2(fabsf(Zc.x), fabsf(Zc.y)), K.f_c3, mul2(mul2(dth, dth), K.f_c5)));
                const F2 wn[1] = {mul2(ex2_2(e), mul2(h, nw.b))};
                mb_accumulate2<NB, 1, SCATTER>(MB, rows, s, wn, 5, acc);
            }
        }
    }
    {   // cloud + band1..3 on one grid
        const F2 h = f2(Pa.hA, Pb.hA), mid = f2(Pa.midA, Pb.midA);
        const float by_min = fminf(K.b_y[0], fminf(K.b_y[1], K.b_y[2]));
        for (int k = 0; k < K.n_nodes; ++k) {
            const Pair<float> nw = nodes[k];
            const MbNode2 s = mb_node2<SCATTER>(K, fma2(h, nw.a, mid), ux, uy, uz, ox, oy, oz);
            const F2 hw = mul2(h, nw.b);
            const F2 xc = add2(s.xh, -K.cx0), yc = add2(s.yh, -K.cy0), zc = add2(s.zh, -K.cz0);
            const F2 Rc2 = fma2(xc, xc, fma2(yc, yc, mul2(zc, zc)));
            const F2 Zc = fma2(xc, K.cnx, fma2(yc, K.cny, mul2(zc, K.cnz)));
            const F2 n0 = cloud_density2(K, Rc2, Zc);
            // band skips decided on (n . X)^2 > c R^2 before 1 / R is formed (kelsall_group_a_x2)
            const F2 d1 = fma2(s.xh, K.bnx[0], fma2(s.yh, K.bny[0], mul2(s.zh, K.bnz[0])));
            const F2 d2 = fma2(s.xh, K.bnx[1], fma2(s.yh, K.bny[1], mul2(s.zh, K.bnz[1])));
            const F2 d3 = fma2(s.xh, K.bnx[2], fma2(s.yh, K.bny[2], mul2(s.zh, K.bnz[2])));
            const F2 thr = mul2(s.Rh2, Math<float>::kS2Underflow * 1.00001f);
            const F2 q1 = mul2(d1, d1), q2 = mul2(d2, d2), q3 = mul2(d3, d3);
            const bool need1 = warp_any(q1.x <= thr.x || q1.y <= thr.y);
            const bool need2 = warp_any(q2.x <= thr.x || q2.y <= thr.y);
            const bool need3 = warp_any(q3.x <= thr.x || q3.y <= thr.y);
            if (need1 || need2 || need3) {
                const F2 rinv = rsq_2(s.Rh2);
                F2 rr1 = rinv, rr2 = rinv, rr3 = rinv;
                const F2 ymin = mul2(s.Rh2, by_min);
                if (warp_any(ymin.x < Math<float>::kRadialOne || ymin.y < Math<float>::kRadialOne)) {
                    if (need1 || (K.share13 && need3)) band_radial2(s.Rh2, K.b_y[0], rinv, rr1);
                    if (need2) band_radial2(s.Rh2, K.b_y[1], rinv, rr2);
                    if (K.share13) rr3 = rr1;
                    else if (need3) band_radial2(s.Rh2, K.b_y[2], rinv, rr3);
                }
                F2 wn[4] = {mul2(n0, hw), f2(0.f), f2(0.f), f2(0.f)};
                if (need1) wn[1] = mul2(band_density2_dot(d1, rinv, rr1, K.b_c3[0]), hw);
                if (need2) wn[2] = mul2(band_density2_dot(d2, rinv, rr2, K.b_c3[1]), hw);
                if (need3) wn[3] = mul2(band_density2_dot(d3, rinv, rr3, K.b_c3[2]), hw);
                mb_accumulate2<NB, 4, SCATTER>(MB, rows, s, wn, 0, acc);
            } else {
                const F2 wn[1] = {mul2(n0, hw)};
                mb_accumulate2<NB, 1, SCATTER>(MB, rows, s, wn, 0, acc);
            }
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < NB; ++b) emit(b, acc[b].x, acc[b].y);
}

}  // namespace zodi
