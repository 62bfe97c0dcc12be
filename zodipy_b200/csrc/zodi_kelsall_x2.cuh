// zodi_kelsall_x2.cuh - packed-fp32 variant of the fused Kelsall group (cloud + 3 bands).
//
// The scalar fused kernel is bound by the SM's issue slots (ncu: 88.7 % issue utilisation, XU pipe
// 76.6 %, FMA pipe 65.6 %; profiles/r1_ncu_kelsall_fp32_nside1024.md): of ~117 warp-instructions
// per quadrature node, 82 are plain fp32 multiply/add/fma.  Blackwell's packed FFMA2 / FMUL2 /
// FADD2 (PTX `fma.rn.f32x2`; sm_100+) execute two fp32 operations per lane in ONE issue slot (at
// half the instruction rate, so the FMA pipe load is unchanged).  Here every thread integrates
// TWO lines of sight and keeps each per-node quantity as an (a, b) register pair, which halves
// the issue slots of the arithmetic and leaves the XU (MUFU) pipe as the limiter.
// Transcendentals stay scalar MUFU ops (there is no packed form); compares/selects stay scalar.
//
// Arithmetic is the same sequence of operations as kelsall_group_a<float> (zodi_kelsall.cuh), so
// both kernels produce bit-identical results (tests/test_gpu_parity.py checks this).
#pragma once

#include "zodi_kelsall.cuh"

#ifndef ZODI_X2_BAND_PRETEST
#define ZODI_X2_BAND_PRETEST 1  // band skips decided before 1/R (see kelsall_group_a_x2); 0: round-1 flow (A/B)
#endif
#ifndef ZODI_X2_RF_FIRST
#define ZODI_X2_RF_FIRST 1  // ring / feature loops before the cloud + bands loop (register pressure); 0: after (A/B)
#endif
#ifndef ZODI_X2_UNROLL
#define ZODI_X2_UNROLL 1  // node-loop unroll factor of the packed cloud + bands loop (measured: see DESIGN.md)
#endif

#ifndef ZODI_X2_RF_UNROLL
// node-loop unroll factor of the packed ring and feature loops: 5 (with 64 registers, see zodi_launch_x2.cu)
// measured 2 - 7 % faster than 1 over nside 32 ... 1024 (profiles/r2_ab_ring_feature_unroll_sweep.jsonl)
#define ZODI_X2_RF_UNROLL 5
#endif

namespace zodi {

constexpr int kX2Unroll = ZODI_X2_UNROLL;
constexpr int kX2RfUnroll = ZODI_X2_RF_UNROLL;

struct alignas(8) F2 { float x, y; };

ZODI_HD F2 f2(float a, float b) { F2 r; r.x = a; r.y = b; return r; }
ZODI_HD F2 f2(float a) { return f2(a, a); }

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float2 as_f2(F2 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ F2 from_f2(float2 v) { return f2(v.x, v.y); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { return from_f2(__ffma2_rn(as_f2(a), as_f2(b), as_f2(c))); }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { return from_f2(__fmul2_rn(as_f2(a), as_f2(b))); }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { return from_f2(__fadd2_rn(as_f2(a), as_f2(b))); }
#else
inline F2 fma2(F2 a, F2 b, F2 c) { return f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline F2 mul2(F2 a, F2 b) { return f2(a.x * b.x, a.y * b.y); }
inline F2 add2(F2 a, F2 b) { return f2(a.x + b.x, a.y + b.y); }
#endif
ZODI_HD F2 fma2(F2 a, float b, F2 c) { return fma2(a, f2(b), c); }
ZODI_HD F2 fma2(F2 a, float b, float c) { return fma2(a, f2(b), f2(c)); }
ZODI_HD F2 fma2(F2 a, F2 b, float c) { return fma2(a, b, f2(c)); }
ZODI_HD F2 mul2(F2 a, float b) { return mul2(a, f2(b)); }
ZODI_HD F2 add2(F2 a, float b) { return add2(a, f2(b)); }

// scalar (MUFU / select) helpers applied to both halves
ZODI_HD F2 ex2_2(F2 v) { return f2(Math<float>::exp2_(v.x), Math<float>::exp2_(v.y)); }
ZODI_HD F2 ex2_neg2(F2 v) { return f2(Math<float>::exp2_neg_(v.x), Math<float>::exp2_neg_(v.y)); }
ZODI_HD F2 lg2_2(F2 v) { return f2(Math<float>::log2_(v.x), Math<float>::log2_(v.y)); }
ZODI_HD F2 rsq_2(F2 v) { return f2(Math<float>::rsqrt_(v.x), Math<float>::rsqrt_(v.y)); }
ZODI_HD F2 sqrt_2(F2 v) { return f2(Math<float>::sqrt_(v.x), Math<float>::sqrt_(v.y)); }
ZODI_HD F2 clamp1_2(F2 v) {
    return f2(fmaxf(-1.0f, fminf(1.0f, v.x)), fmaxf(-1.0f, fminf(1.0f, v.y)));
}
// asin_unit() for both halves: the polynomial runs packed, the selects per half.
ZODI_HD F2 asin_unit2(F2 c) {
    const F2 a = f2(fabsf(c.x), fabsf(c.y));
    const bool bx = a.x > 0.5f, by = a.y > 0.5f;
    const F2 zb = fma2(a, f2(-0.5f), f2(0.5f)), zs = mul2(a, a);
    const F2 z = f2(bx ? zb.x : zs.x, by ? zb.y : zs.y);
    const F2 s = f2(bx ? Math<float>::sqrt_(z.x) : a.x, by ? Math<float>::sqrt_(z.y) : a.y);
    F2 p = fma2(f2(kAsinP4), z, f2(kAsinP3));
    p = fma2(p, z, f2(kAsinP2));
    p = fma2(p, z, f2(kAsinP1));
    p = fma2(p, z, f2(kAsinP0));
    p = fma2(mul2(s, z), p, s);
    const F2 r = fma2(p, f2(-2.0f), f2(1.57079637f));
    return f2(copysignf(bx ? r.x : p.x, c.x), copysignf(by ? r.y : p.y, c.y));
}
// phase_of_cos<float>() for both halves.
ZODI_HD F2 phase_of_cos2(const KelsallModel<float>& K, F2 c) {
    const F2 t = asin_unit2(c);
    if (!K.phase_poly_ok) {
        const F2 th = add2(t, 1.57079637f);
        return f2(K.C1p + K.C2p * th.x + Math<float>::exp2_(K.C3l * th.x),
                  K.C1p + K.C2p * th.y + Math<float>::exp2_(K.C3l * th.y));
    }
    F2 p;
    if (K.phase_terms <= 8) {
        p = f2(K.phase_poly[7]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 6; k >= 0; --k) p = fma2(p, t, f2(K.phase_poly[k]));
    } else {
        p = f2(K.phase_poly[kPhaseTerms - 1]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = kPhaseTerms - 2; k >= 0; --k) p = fma2(p, t, f2(K.phase_poly[k]));
    }
    return p;
}
// Phi(Theta) / R_h^2 for both halves (node_source<float, true>'s scattering block).
ZODI_HD F2 scatter_term2(const KelsallModel<float>& K, F2 ux, F2 uy, F2 uz, F2 xh, F2 yh, F2 zh, F2 rh_inv) {
    const F2 ct = clamp1_2(mul2(fma2(ux, xh, fma2(uy, yh, mul2(uz, zh))), rh_inv));
    return mul2(mul2(phase_of_cos2(K, ct), rh_inv), rh_inv);
}

// Loop-invariant handle of the staged table for table_at2.  Device: the shared-window byte address of
// the table minus 8 * 0x4B400000 (mod 2^32), so that entry i is at handle + 8 * bits(s) with s = magic + i
// and the index extraction costs one shift-add per lookup; the value passes through a warp shuffle
// (once per line of sight), otherwise ptxas re-associates the constant back into the loop (one more add
// per lookup).
struct TableRef {
#if defined(__CUDA_ARCH__)
    uint32_t base;
#else
    const Pair<float>* tab;
#endif
};
ZODI_HD TableRef table_ref(const Pair<float>* tab) {
    TableRef r;
#if defined(__CUDA_ARCH__)
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(tab) - (0x4B400000u << 3);
    r.base = __shfl_sync(0xffffffffu, b, 0);  // opaque to ptxas, unlike a move
#else
    r.tab = tab;
#endif
    return r;
}

// Table lookup for two temperatures (same arithmetic as table_at<float>).
ZODI_HD F2 table_at2(TableRef ref, F2 t, float t_top) {
    t = f2(fminf(fmaxf(t.x, 0.0f), t_top), fminf(fmaxf(t.y, 0.0f), t_top));
    const float magic = 12582912.0f;
    const F2 s = add2(add2(t, -0.5f), magic);
    Pair<float> e0, e1;
#if defined(__CUDA_ARCH__)
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e0.a), "=f"(e0.b) : "r"(ref.base + (__float_as_uint(s.x) << 3)));
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e1.a), "=f"(e1.b) : "r"(ref.base + (__float_as_uint(s.y) << 3)));
#else
    int b0, b1;
    memcpy(&b0, &s.x, 4);
    memcpy(&b1, &s.y, 4);
    e0 = ref.tab[b0 - 0x4B400000];
    e1 = ref.tab[b1 - 0x4B400000];
#endif
    const F2 frac = fma2(add2(s, -magic), -1.0f, t);  // t - (s - magic)
    // two scalar FMAs: the table entries arrive as (a, b) pairs per half, so a packed FMA would need
    // four register moves to regroup them into (b0, b1) / (a0, a1) (3 issue slots more; same FMA-pipe load)
    return f2(fmaf(e0.b, frac.x, e0.a), fmaf(e1.b, frac.y, e1.a));
}

// 1 - 2^-y for both halves (same switch and polynomial as Math<float>::one_minus_exp2_neg); the
// MUFU pair is skipped when no lane of the warp is in the window where it matters.
ZODI_HD F2 one_minus_exp2_neg2(F2 y) {
    const F2 small = mul2(y, fma2(y, fma2(y, 0.05550411f, -0.24022651f), 0.69314718f));
    F2 direct = f2(1.0f);  // y > 25.05: 1 - 2^-y rounds to exactly 1 (2^-y < 2^-25)
    const bool need = (y.x >= 0.04508422f && y.x <= 25.05f) || (y.y >= 0.04508422f && y.y <= 25.05f);
    if (warp_any(need)) direct = fma2(ex2_neg2(y), -1.0f, 1.0f);  // 1 - e
    return f2(y.x < 0.04508422f ? small.x : direct.x, y.y < 0.04508422f ? small.y : direct.y);
}

// y = R^2 log2e^(1/10) / delta_r^2, so y^10 = log2e (R/delta_r)^20.  Beyond y = kRadialOne = 1.38
// (R > 1.15 delta_r, most of a line of sight that runs out to 5.2 AU) 2^(-y^10) < 2^-25 and the term
// rounds to exactly 1: when the whole warp is there the power chain, the polynomial and the MUFU pair
// are skipped.
// rr = rinv * (1 - 2^(-y^10)) and true, or false when the term is skipped for the whole warp: the factor is
// then rinv itself (x * 1 == x exactly) and the caller passes rinv on, so neither the constant 1, the
// product nor a register copy is formed.
ZODI_HD bool band_radial2(F2 Rh2, float by, F2 rinv, F2& rr) {
    const F2 y = mul2(Rh2, by);
    const bool cut = warp_any(y.x < Math<float>::kRadialOne || y.y < Math<float>::kRadialOne);  // warp-uniform
    if (cut) {
        const F2 y2 = mul2(y, y), y4 = mul2(y2, y2), y5 = mul2(y4, y);
        rr = mul2(rinv, one_minus_exp2_neg2(mul2(y5, y5)));
    }
    return cut;
}

// Adds w*B * n_band to acc, n_band = exp(-s^6) (1 + s^4/v) * (rinv * rad); skipped (n_band == 0)
// when every lane of the warp is far enough from the band plane for exp(-s^6) to flush to zero.
template <bool SCATTER>
ZODI_HD void band_accumulate2(F2& acc, F2& accS, F2 wB, F2 wF, F2 xh, F2 yh, F2 zh, F2 rinv, F2 rinv_rad,
                              float bx, float by, float bz, float c3) {
    const F2 sz = mul2(fma2(xh, bx, fma2(yh, by, mul2(zh, bz))), rinv);
    const F2 s2 = mul2(sz, sz);
    // a warp whose lanes all have s^2 > kS2Underflow has exp(-s^6) == 0 in every lane (the scalar kernel
    // tests s^6 <= 126; lanes between the two thresholds get an exact 0 from the flushed ex2 instead)
    if (warp_any(s2.x <= Math<float>::kS2Underflow || s2.y <= Math<float>::kS2Underflow)) {
        const F2 s4 = mul2(s2, s2), s6 = mul2(s4, s2);
        const F2 n = mul2(mul2(ex2_neg2(s6), fma2(s4, c3, 1.0f)), rinv_rad);
        acc = fma2(wB, n, acc);
        if (SCATTER) accS = fma2(wF, n, accS);
    }
}

// Everything the packed node loops need of one line of sight, in single precision: the product of the
// fp64 per-line-of-sight prologue (direction, positions, ray / sphere ranges, Earth longitude).  With
// L > 1 lanes per pair of lines of sight the prologue of a line of sight is computed by ONE lane and
// handed to the others by warp shuffles (it used to be replicated in every lane).
struct LosPre {
    float ux, uy, uz, ox, oy, oz;
    float hA, midA;            // cloud + bands: half-range and mid-point of the interval (brightness.py:41)
    float hR, midR, hF, midF;  // ring, feature
    float cr, sr;              // feature: cos / sin of -(theta_earth + theta_0), see feature_rotation()
};
constexpr int kLosPreFloats = 14;
static_assert(sizeof(LosPre) == kLosPreFloats * sizeof(float), "LosPre is shuffled float by float");

template <bool HAS_RF>
ZODI_HD LosPre los_pre(const KelsallModel<float>& K, double ux, double uy, double uz, double ox, double oy,
                       double oz, double ex, double ey, uint32_t outside_mask) {
    const LosGeometry<float> G = los_geometry<float>(ux, uy, uz, ox, oy, oz);
    LosPre P;
    P.ux = G.ux; P.uy = G.uy; P.uz = G.uz; P.ox = G.ox; P.oy = G.oy; P.oz = G.oz;
    los_interval<float>(G, K.cutA_in, K.cutA_out, outside_mask & 1u, (outside_mask >> 1) & 1u, P.hA, P.midA);
    P.hR = P.midR = P.hF = P.midF = 0.f;
    P.cr = 1.f; P.sr = 0.f;
    if (HAS_RF) {
        los_interval<float>(G, K.cutR_in, K.cutR_out, (outside_mask >> 8) & 1u, (outside_mask >> 9) & 1u, P.hR, P.midR);
        los_interval<float>(G, K.cutF_in, K.cutF_out, (outside_mask >> 10) & 1u, (outside_mask >> 11) & 1u, P.hF, P.midF);
        feature_rotation<float>(ex, ey, K.f_cos0, K.f_sin0, P.cr, P.sr);  // ring / feature are Sun-centred
    }
    return P;
}

// Node k0 + sub of a rule split over L lanes, with a warp-uniform trip count (the loop bodies contain
// warp votes): surplus iterations re-evaluate the last node with weight 0.
template <int L>
ZODI_HD Pair<float> lane_node(const Pair<float>* nodes, int n_nodes, int k0, int sub) {
    if (L == 1) return nodes[k0];
    const int k = k0 + sub;
    Pair<float> nw = nodes[k < n_nodes ? k : n_nodes - 1];
    if (k >= n_nodes) nw.b = 0.f;
    return nw;
}

// cloud_density<float>() for both halves.
ZODI_HD F2 cloud_density2(const KelsallModel<float>& K, F2 Rc2, F2 Zc) {
    const F2 zeta = mul2(f2(fabsf(Zc.x), fabsf(Zc.y)), rsq_2(Rc2));
    const F2 g_in = mul2(mul2(zeta, zeta), K.c_inv2mu), g_out = add2(zeta, -K.c_halfmu);
    const F2 g = f2(zeta.x < K.c_mu ? g_in.x : g_out.x, zeta.y < K.c_mu ? g_in.y : g_out.y);
    const F2 gp = ex2_2(mul2(lg2_2(g), K.c_gamma));
    return ex2_2(fma2(lg2_2(Rc2), K.c_mha, mul2(gp, K.c_mbl)));
}

// band_accumulate2 with the plane distance dot = n . X already formed and the skip already decided by the
// caller (ZODI_X2_BAND_PRETEST): same operations on the lanes that are computed.
ZODI_HD void band_accumulate2_dot(F2& acc, F2 wB, F2 dot, F2 rinv, F2 rinv_rad, float c3) {
    const F2 sz = mul2(dot, rinv);
    const F2 s2 = mul2(sz, sz), s4 = mul2(s2, s2), s6 = mul2(s4, s2);
    const F2 n = mul2(mul2(ex2_neg2(s6), fma2(s4, c3, 1.0f)), rinv_rad);
    acc = fma2(wB, n, acc);
}

// Group A (cloud + band1..3) for two lines of sight; lane `sub` of L takes nodes sub, sub + L, ...
// emit(ci, partial_a, partial_b): partial quadrature sums times the half-range (the caller adds the
// L partials of a line of sight).
template <bool SHARE13, bool SCATTER, int L, typename Emit>
ZODI_HD void kelsall_group_a_x2(const KelsallModel<float>& K, const Pair<float>* tab,
                                const Pair<float>* nodes, const LosPre& Pa, const LosPre& Pb, int sub, Emit emit) {
    const F2 h = f2(Pa.hA, Pb.hA), mid = f2(Pa.midA, Pb.midA);
    const F2 ux = f2(Pa.ux, Pb.ux), uy = f2(Pa.uy, Pb.uy), uz = f2(Pa.uz, Pb.uz);
    const F2 ox = f2(Pa.ox, Pb.ox), oy = f2(Pa.oy, Pb.oy), oz = f2(Pa.oz, Pb.oz);
    F2 a0 = f2(0.f), a1 = f2(0.f), a2 = f2(0.f), a3 = f2(0.f);
    F2 s0 = f2(0.f), s1 = f2(0.f), s2 = f2(0.f), s3 = f2(0.f);  // scattering accumulators
    const TableRef tref = table_ref(tab);
    const float by_min = fminf(K.b_y[0], fminf(K.b_y[1], K.b_y[2]));
#if defined(__CUDA_ARCH__)
#pragma unroll kX2Unroll
#endif
    for (int k0 = 0; k0 < K.n_nodes; k0 += L) {
        const Pair<float> nw = lane_node<L>(nodes, K.n_nodes, k0, sub);
        // shared source quantities (node_source<float, false>)
        const F2 R_los = fma2(h, nw.a, mid);
        const F2 xh = fma2(R_los, ux, ox), yh = fma2(R_los, uy, oy), zh = fma2(R_los, uz, oz);
        const F2 Rh2 = fma2(xh, xh, fma2(yh, yh, mul2(zh, zh)));
        const F2 lgR = lg2_2(Rh2);
        const F2 t = fma2(ex2_2(mul2(lgR, K.mhd)), K.t_scale, K.t_ofs);
        const F2 B = table_at2(tref, t, K.t_top);
        // cloud: independent of the bands, placed next to the source terms so that its chain of ten MUFU ops
        // overlaps their arithmetic (the band blocks below are separated by warp votes, which fence scheduling)
        const F2 xc = add2(xh, -K.cx0), yc = add2(yh, -K.cy0), zc = add2(zh, -K.cz0);
        const F2 Rc2 = fma2(xc, xc, fma2(yc, yc, mul2(zc, zc)));
        const F2 Zc = fma2(xc, K.cnx, fma2(yc, K.cny, mul2(zc, K.cnz)));
        const F2 n0 = cloud_density2(K, Rc2, Zc);
        // bands
        const F2 wB = mul2(B, nw.b);
#if ZODI_X2_BAND_PRETEST
        if (!SCATTER) {
            // Decide the three band skips on (n . X)^2 > c R^2 (FMA pipe only) BEFORE forming 1/R: when every
            // band of the (warp, node) is skipped - exp(-s^6) == 0 in all lanes - the MUFU.RSQ pair is dead
            // too.  The threshold carries a 1e-5 margin over kS2Underflow, so a skipped band has s^2 > 5.02
            // whatever the rounding of rsqrt; lanes in between are computed and get their exact 0 from ex2.
            const F2 d1 = fma2(xh, K.bnx[0], fma2(yh, K.bny[0], mul2(zh, K.bnz[0])));
            const F2 d2 = fma2(xh, K.bnx[1], fma2(yh, K.bny[1], mul2(zh, K.bnz[1])));
            const F2 d3 = fma2(xh, K.bnx[2], fma2(yh, K.bny[2], mul2(zh, K.bnz[2])));
            const F2 thr = mul2(Rh2, Math<float>::kS2Underflow * 1.00001f);
            const F2 q1 = mul2(d1, d1), q2 = mul2(d2, d2), q3 = mul2(d3, d3);
            const bool need1 = warp_any(q1.x <= thr.x || q1.y <= thr.y);
            const bool need2 = warp_any(q2.x <= thr.x || q2.y <= thr.y);
            const bool need3 = warp_any(q3.x <= thr.x || q3.y <= thr.y);
            if (need1 || need2 || need3) {
                const F2 rinv = rsq_2(Rh2);
                const F2 ymin = mul2(Rh2, by_min);
                if (!warp_any(ymin.x < Math<float>::kRadialOne || ymin.y < Math<float>::kRadialOne)) {
                    // common case: every radial cut-off factor is exactly 1 (separate code, so that no
                    // register copies of 1/R are formed for it)
                    if (need1) band_accumulate2_dot(a1, wB, d1, rinv, rinv, K.b_c3[0]);
                    if (need2) band_accumulate2_dot(a2, wB, d2, rinv, rinv, K.b_c3[1]);
                    if (need3) band_accumulate2_dot(a3, wB, d3, rinv, rinv, K.b_c3[2]);
                } else {
                    F2 rr1 = rinv, rr2 = rinv, rr3 = rinv;
                    if (need1 || (SHARE13 && need3)) band_radial2(Rh2, K.b_y[0], rinv, rr1);
                    if (need2) band_radial2(Rh2, K.b_y[1], rinv, rr2);
                    if (SHARE13) rr3 = rr1;
                    else if (need3) band_radial2(Rh2, K.b_y[2], rinv, rr3);
                    if (need1) band_accumulate2_dot(a1, wB, d1, rinv, rr1, K.b_c3[0]);
                    if (need2) band_accumulate2_dot(a2, wB, d2, rinv, rr2, K.b_c3[1]);
                    if (need3) band_accumulate2_dot(a3, wB, d3, rinv, rr3, K.b_c3[2]);
                }
            }
            a0 = fma2(wB, n0, a0);
            continue;
        }
#endif
        const F2 rinv = rsq_2(Rh2);
        F2 wF = f2(0.f);
        if (SCATTER) wF = mul2(scatter_term2(K, ux, uy, uz, xh, yh, zh, rinv), nw.b);  // rinv == 1/R_h (bands are Sun-centred)
        // 1 - 2^(-y^10) is exactly 1 for every band once R^2 * min(b_y) >= kRadialOne in all lanes (b_y > 0:
        // the products are monotonic in b_y): one test then replaces the per-band ones - the common case
        F2 rr = rinv;
        const F2 ymin = mul2(Rh2, by_min);
        if (!warp_any(ymin.x < Math<float>::kRadialOne || ymin.y < Math<float>::kRadialOne)) {
            band_accumulate2<SCATTER>(a1, s1, wB, wF, xh, yh, zh, rinv, rinv, K.bnx[0], K.bny[0], K.bnz[0], K.b_c3[0]);
            band_accumulate2<SCATTER>(a2, s2, wB, wF, xh, yh, zh, rinv, rinv, K.bnx[1], K.bny[1], K.bnz[1], K.b_c3[1]);
            band_accumulate2<SCATTER>(a3, s3, wB, wF, xh, yh, zh, rinv, rinv, K.bnx[2], K.bny[2], K.bnz[2], K.b_c3[2]);
        } else {
            // the two sides of each branch are the same calls with rr or rinv as the radial factor
            if (band_radial2(Rh2, K.b_y[0], rinv, rr)) {
                band_accumulate2<SCATTER>(a1, s1, wB, wF, xh, yh, zh, rinv, rr, K.bnx[0], K.bny[0], K.bnz[0], K.b_c3[0]);
                if (SHARE13) band_accumulate2<SCATTER>(a3, s3, wB, wF, xh, yh, zh, rinv, rr, K.bnx[2], K.bny[2], K.bnz[2], K.b_c3[2]);
            } else {
                band_accumulate2<SCATTER>(a1, s1, wB, wF, xh, yh, zh, rinv, rinv, K.bnx[0], K.bny[0], K.bnz[0], K.b_c3[0]);
                if (SHARE13) band_accumulate2<SCATTER>(a3, s3, wB, wF, xh, yh, zh, rinv, rinv, K.bnx[2], K.bny[2], K.bnz[2], K.b_c3[2]);
            }
            if (band_radial2(Rh2, K.b_y[1], rinv, rr))
                band_accumulate2<SCATTER>(a2, s2, wB, wF, xh, yh, zh, rinv, rr, K.bnx[1], K.bny[1], K.bnz[1], K.b_c3[1]);
            else
                band_accumulate2<SCATTER>(a2, s2, wB, wF, xh, yh, zh, rinv, rinv, K.bnx[1], K.bny[1], K.bnz[1], K.b_c3[1]);
            if (!SHARE13) {
                if (band_radial2(Rh2, K.b_y[2], rinv, rr))
                    band_accumulate2<SCATTER>(a3, s3, wB, wF, xh, yh, zh, rinv, rr, K.bnx[2], K.bny[2], K.bnz[2], K.b_c3[2]);
                else
                    band_accumulate2<SCATTER>(a3, s3, wB, wF, xh, yh, zh, rinv, rinv, K.bnx[2], K.bny[2], K.bnz[2], K.b_c3[2]);
            }
        }
        a0 = fma2(wB, n0, a0);
        if (SCATTER) s0 = fma2(wF, n0, s0);
    }
    // scalar kernel: h * fma(aB, accB, aS * accS)   (accS == 0 without scattering)
    const F2 r0 = mul2(h, fma2(a0, K.aB[0], mul2(s0, K.aS[0]))), r1 = mul2(h, fma2(a1, K.aB[1], mul2(s1, K.aS[1])));
    const F2 r2 = mul2(h, fma2(a2, K.aB[2], mul2(s2, K.aS[2]))), r3 = mul2(h, fma2(a3, K.aB[3], mul2(s3, K.aS[3])));
    emit(0, r0.x, r0.y); emit(1, r1.x, r1.y); emit(2, r2.x, r2.y); emit(3, r3.x, r3.y);
}

// Ring of TWO lines of sight per loop (ring | ring in the two halves of every register pair).  Same
// operations per line of sight as kelsall_ring<float, SCATTER> (bit-identical results).
template <bool SCATTER, int L, typename Emit>
ZODI_HD void kelsall_ring_x2(const KelsallModel<float>& K, const Pair<float>* tab, const Pair<float>* nodes,
                             const LosPre& Pa, const LosPre& Pb, int sub, Emit emit) {
    const F2 h = f2(Pa.hR, Pb.hR), mid = f2(Pa.midR, Pb.midR);
    const F2 ux = f2(Pa.ux, Pb.ux), uy = f2(Pa.uy, Pb.uy), uz = f2(Pa.uz, Pb.uz);
    const F2 ox = f2(Pa.ox, Pb.ox), oy = f2(Pa.oy, Pb.oy), oz = f2(Pa.oz, Pb.oz);
    F2 acc = f2(0.f), accS = f2(0.f);
    const TableRef tref = table_ref(tab);
#if defined(__CUDA_ARCH__)
#pragma unroll kX2RfUnroll
#endif
    for (int k = sub; k < K.n_nodes; k += L) {
        const Pair<float> nw = nodes[k];
        const F2 R_los = fma2(h, nw.a, mid);
        const F2 xh = fma2(R_los, ux, ox), yh = fma2(R_los, uy, oy), zh = fma2(R_los, uz, oz);
        const F2 Rh2 = fma2(xh, xh, fma2(yh, yh, mul2(zh, zh)));
        const F2 d = add2(sqrt_2(Rh2), -K.r_R);
        const F2 B = table_at2(tref, fma2(ex2_2(mul2(lg2_2(Rh2), K.mhd)), K.t_scale, K.t_ofs), K.t_top);
        const F2 Zc = fma2(xh, K.rnx, fma2(yh, K.rny, mul2(zh, K.rnz)));
        const F2 n = ex2_2(fma2(mul2(d, d), K.r_c2, mul2(f2(fabsf(Zc.x), fabsf(Zc.y)), K.r_c3)));
        acc = fma2(mul2(B, nw.b), n, acc);
        if (SCATTER) {
            const F2 F = scatter_term2(K, ux, uy, uz, xh, yh, zh, rsq_2(Rh2));
            accS = fma2(mul2(F, nw.b), n, accS);
        }
    }
    const F2 r = mul2(h, fma2(acc, K.aB[4], mul2(accS, K.aS[4])));
    emit(r.x, r.y);
}

// |atan2(y, x)| for both halves: Math<float>::atan2_abs_ with the polynomial, the quotient and the two
// reflections on packed instructions (min / max / selects per half).
ZODI_HD F2 atan2_abs2(F2 y, F2 x) {
    const F2 ax = f2(fabsf(x.x), fabsf(x.y)), ay = f2(fabsf(y.x), fabsf(y.y));
    const F2 mn = f2(fminf(ax.x, ay.x), fminf(ax.y, ay.y));
    const F2 mx = f2(fmaxf(fmaxf(ax.x, ay.x), 1e-30f), fmaxf(fmaxf(ax.y, ay.y), 1e-30f));
    const F2 a = mul2(mn, f2(atan_rcp(mx.x), atan_rcp(mx.y))), s = mul2(a, a);
    F2 p = f2(kAtanC8);
    p = fma2(p, s, kAtanC7); p = fma2(p, s, kAtanC6); p = fma2(p, s, kAtanC5); p = fma2(p, s, kAtanC4);
    p = fma2(p, s, kAtanC3); p = fma2(p, s, kAtanC2); p = fma2(p, s, kAtanC1); p = fma2(p, s, kAtanC0);
    F2 r = mul2(a, p);
    const F2 rc = fma2(r, -1.0f, 1.57079637f);  // pi/2 - r
    r = f2(ay.x > ax.x ? rc.x : r.x, ay.y > ax.y ? rc.y : r.y);
    const F2 rs = fma2(r, -1.0f, 3.14159274f);  // pi - r
    return f2(x.x < 0.0f ? rs.x : r.x, x.y < 0.0f ? rs.y : r.y);
}

// Feature of TWO lines of sight per loop (feature | feature): rotation, atan polynomial and exponent run
// packed.  Same operations per line of sight as kelsall_feature<float, SCATTER>.
template <bool SCATTER, int L, typename Emit>
ZODI_HD void kelsall_feature_x2(const KelsallModel<float>& K, const Pair<float>* tab, const Pair<float>* nodes,
                                const LosPre& Pa, const LosPre& Pb, int sub, Emit emit) {
    const F2 h = f2(Pa.hF, Pb.hF), mid = f2(Pa.midF, Pb.midF);
    const F2 ux = f2(Pa.ux, Pb.ux), uy = f2(Pa.uy, Pb.uy), uz = f2(Pa.uz, Pb.uz);
    const F2 ox = f2(Pa.ox, Pb.ox), oy = f2(Pa.oy, Pb.oy), oz = f2(Pa.oz, Pb.oz);
    const F2 cr = f2(Pa.cr, Pb.cr), sr = f2(Pa.sr, Pb.sr), msr = f2(-Pa.sr, -Pb.sr);
    F2 acc = f2(0.f), accS = f2(0.f);
    const TableRef tref = table_ref(tab);
#if defined(__CUDA_ARCH__)
#pragma unroll kX2RfUnroll
#endif
    for (int k = sub; k < K.n_nodes; k += L) {
        const Pair<float> nw = nodes[k];
        const F2 R_los = fma2(h, nw.a, mid);
        const F2 xh = fma2(R_los, ux, ox), yh = fma2(R_los, uy, oy), zh = fma2(R_los, uz, oz);
        const F2 Rh2 = fma2(xh, xh, fma2(yh, yh, mul2(zh, zh)));
        const F2 t = fma2(ex2_2(mul2(lg2_2(Rh2), K.mhd)), K.t_scale, K.t_ofs);
        const F2 B = table_at2(tref, t, K.t_top);
        const F2 d = add2(sqrt_2(Rh2), -K.f_R);
        const F2 Zc = fma2(xh, K.fnx, fma2(yh, K.fny, mul2(zh, K.fnz)));
        const F2 xr = fma2(xh, cr, mul2(yh, sr)), yr = fma2(yh, cr, mul2(xh, msr));  // x * (-s) == -(x * s) exactly
        const F2 dth = atan2_abs2(yr, xr);  // only dth^2 is used
        const F2 e = fma2(mul2(d, d), K.f_c2, fma2(f2(fabsf(Zc.x), fabsf(Zc.y)), K.f_c3, mul2(mul2(dth, dth), K.f_c5)));
        const F2 n = ex2_2(e);
        acc = fma2(mul2(B, nw.b), n, acc);
        if (SCATTER) {
            const F2 F = scatter_term2(K, ux, uy, uz, xh, yh, zh, rsq_2(Rh2));
            accS = fma2(mul2(F, nw.b), n, accS);
        }
    }
    const F2 r = mul2(h, fma2(acc, K.aB[5], mul2(accS, K.aS[5])));
    emit(r.x, r.y);
}

}  // namespace zodi
