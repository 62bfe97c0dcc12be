// zodi_emu.cpp - HOST build of the device routines in zodipy_b200/csrc/zodi_device.cuh.
//
// TEST TOOL ONLY (lives under tests/): lets the CPU-only test-suite exercise the exact arithmetic
// the CUDA kernels run - descriptor -> device-form constants -> integrate_line_of_sight<Real>() -
// without a GPU, so formula/constant mistakes are caught before GPU time is spent.  It is never
// built into, loaded by, or reachable from the product library (zodipy_b200/libzodi_b200.so).
// fp32 MUFU intrinsics are replaced by libm here, so fp32 results are a lower bound on the GPU's
// rounding error, not a bit-exact emulation.
#include <cstdint>
#include <vector>

#include "../../zodipy_b200/csrc/zodi_model_build.hpp"
#include "../../zodipy_b200/csrc/zodi_kelsall_x2.cuh"
#include "../../zodipy_b200/csrc/zodi_rrm_x2.cuh"
#include "../../zodipy_b200/csrc/zodi_multiband_x2.cuh"

using namespace zodi;

template <typename Real>
static void run(const zodi_model_desc* d, const DevModel<Real>& M, const std::vector<Pair<Real>>& tab,
                const std::vector<Pair<Real>>& nodes, int64_t n, const double* u, const double* obs,
                int64_t n_obs, const double* earth, int64_t n_earth, const uint8_t* flags,
                int lanes, double* out /* (n_comps, n) */) {
    uint32_t mask = 0;
    for (int c = 0; c < d->n_comps; ++c) {
        if (flags[2 * c]) mask |= 1u << (2 * c);
        if (flags[2 * c + 1]) mask |= 1u << (2 * c + 1);
    }
    for (int64_t j = 0; j < n; ++j) {
        const int64_t jo = n_obs == n ? j : 0, je = n_earth == n ? j : 0;
        for (int c = 0; c < d->n_comps; ++c) out[c * n + j] = 0.0;
        for (int sub = 0; sub < lanes; ++sub)  // emulate the L lanes of a line of sight serially
            integrate_line_of_sight<Real>(
                M, tab.data(), nodes.data(), u[j], u[n + j], u[2 * n + j], obs[jo], obs[n_obs + jo],
                obs[2 * n_obs + jo], earth[je], earth[n_earth + je], mask, sub, lanes,
                [&](int ci, Real part) { out[ci * n + j] += (double)part; });
    }
}

template <typename Real>
static void run_kelsall(const KelsallModel<Real>& K, const std::vector<Pair<Real>>& tab,
                        const std::vector<Pair<Real>>& nodes, int64_t n, const double* u,
                        const double* obs, int64_t n_obs, const double* earth, int64_t n_earth,
                        const uint8_t* flags, int lanes, double* out) {
    uint32_t mask = 0;
    for (int c = 0; c < K.n_comps; ++c) {
        if (flags[2 * c]) mask |= 1u << (2 * c);
        if (flags[2 * c + 1]) mask |= 1u << (2 * c + 1);
    }
    for (int64_t j = 0; j < n; ++j) {
        const int64_t jo = n_obs == n ? j : 0, je = n_earth == n ? j : 0;
        for (int c = 0; c < K.n_comps; ++c) out[c * n + j] = 0.0;
        auto emit = [&](int ci, Real part) { out[ci * n + j] += (double)part; };
        for (int sub = 0; sub < lanes; ++sub) {
#define ZCALL(RF, SC)                                                                             \
    if (K.share13)                                                                                \
        integrate_kelsall<Real, RF, SC, true>(K, tab.data(), nodes.data(), u[j], u[n + j],         \
                                    u[2 * n + j], obs[jo], obs[n_obs + jo], obs[2 * n_obs + jo],   \
                                    earth[je], earth[n_earth + je], mask, sub, lanes, emit);       \
    else                                                                                          \
    integrate_kelsall<Real, RF, SC, false>(K, tab.data(), nodes.data(), u[j], u[n + j], u[2 * n + j],    \
                                    obs[jo], obs[n_obs + jo], obs[2 * n_obs + jo], earth[je],     \
                                    earth[n_earth + je], mask, sub, lanes, emit)
            if (K.n_comps == 6) { if (K.scatter) ZCALL(true, true); else ZCALL(true, false); }
            else { if (K.scatter) ZCALL(false, true); else ZCALL(false, false); }
#undef ZCALL
        }
    }
}

// Packed (two lines of sight per "thread") routines of zodi_kelsall_x2.cuh (fp32), L lanes per pair
// emulated serially; partial sums are added in lane order.
template <int L>
static void run_kelsall_x2(const KelsallModel<float>& K, const std::vector<Pair<float>>& tab,
                           const std::vector<Pair<float>>& nodes, int64_t n, const double* u,
                           const double* obs, int64_t n_obs, const double* earth, int64_t n_earth,
                           const uint8_t* flags, double* out) {
    uint32_t mask = 0;
    for (int c = 0; c < K.n_comps; ++c) {
        if (flags[2 * c]) mask |= 1u << (2 * c);
        if (flags[2 * c + 1]) mask |= 1u << (2 * c + 1);
    }
    for (int64_t j0 = 0; j0 < n; j0 += 2) {
        const int64_t jj[2] = {j0, j0 + 1 < n ? j0 + 1 : j0};
        LosPre P[2];
        for (int q = 0; q < 2; ++q) {
            const int64_t j = jj[q], jo = n_obs == n ? j : 0, je = n_earth == n ? j : 0;
            if (K.n_comps == 6)
                P[q] = los_pre<true>(K, u[j], u[n + j], u[2 * n + j], obs[jo], obs[n_obs + jo], obs[2 * n_obs + jo],
                                     earth[je], earth[n_earth + je], mask);
            else
                P[q] = los_pre<false>(K, u[j], u[n + j], u[2 * n + j], obs[jo], obs[n_obs + jo], obs[2 * n_obs + jo],
                                      earth[je], earth[n_earth + je], mask);
        }
        double acc[6][2] = {{0.0}};  // like run_kelsall: lane partials are added in double
        for (int sub = 0; sub < L; ++sub) {
            auto emit2 = [&](int ci, float a, float b) { acc[ci][0] += (double)a; acc[ci][1] += (double)b; };
#define ZX2(SH, SC) kelsall_group_a_x2<SH, SC, L>(K, tab.data(), nodes.data(), P[0], P[1], sub, emit2)
            if (K.share13) { if (K.scatter) ZX2(true, true); else ZX2(true, false); }
            else { if (K.scatter) ZX2(false, true); else ZX2(false, false); }
#undef ZX2
            if (K.n_comps == 6) {
                auto ring = [&](float a, float b) { emit2(4, a, b); };
                auto feat = [&](float a, float b) { emit2(5, a, b); };
                if (K.scatter) {
                    kelsall_ring_x2<true, L>(K, tab.data(), nodes.data(), P[0], P[1], sub, ring);
                    kelsall_feature_x2<true, L>(K, tab.data(), nodes.data(), P[0], P[1], sub, feat);
                } else {
                    kelsall_ring_x2<false, L>(K, tab.data(), nodes.data(), P[0], P[1], sub, ring);
                    kelsall_feature_x2<false, L>(K, tab.data(), nodes.data(), P[0], P[1], sub, feat);
                }
            }
        }
        for (int c = 0; c < K.n_comps; ++c)
            for (int q = 1; q >= 0; --q) out[c * n + jj[q]] = acc[c][q];
    }
}

// Fused RRM routine (zodi_rrm.cuh).
template <typename Real>
static void run_rrm(const RrmModel<Real>& R, const std::vector<Pair<Real>>& tab, const std::vector<Pair<Real>>& nodes,
                    int64_t n, const double* u, const double* obs, int64_t n_obs, const double* earth,
                    int64_t n_earth, const uint8_t* flags, int lanes, double* out) {
    uint32_t mask = 0;
    for (int c = 0; c < R_NCOMPS; ++c) {
        if (flags[2 * c]) mask |= 1u << (2 * c);
        if (flags[2 * c + 1]) mask |= 1u << (2 * c + 1);
    }
    for (int64_t j = 0; j < n; ++j) {
        const int64_t jo = n_obs == n ? j : 0, je = n_earth == n ? j : 0;
        for (int c = 0; c < R_NCOMPS; ++c) out[c * n + j] = 0.0;
        for (int sub = 0; sub < lanes; ++sub)
            integrate_rrm<Real>(R, tab.data(), nodes.data(), u[j], u[n + j], u[2 * n + j], obs[jo], obs[n_obs + jo],
                                obs[2 * n_obs + jo], earth[je], earth[n_earth + je], mask, sub, lanes,
                                [&](int ci, Real part) { out[ci * n + j] += (double)part; });
    }
}

// Packed fused RRM routines (zodi_rrm_x2.cuh), two lines of sight per "thread".
static void run_rrm_x2(const RrmModelX2& X, const std::vector<Pair<float>>& tab, const std::vector<Pair<float>>& nodes,
                       int64_t n, const double* u, const double* obs, int64_t n_obs, const double* earth,
                       int64_t n_earth, const uint8_t* flags, double* out) {
    uint32_t mask = 0;
    for (int c = 0; c < R_NCOMPS; ++c) {
        if (flags[2 * c]) mask |= 1u << (2 * c);
        if (flags[2 * c + 1]) mask |= 1u << (2 * c + 1);
    }
    for (int64_t j0 = 0; j0 < n; j0 += 2) {
        const int64_t jj[2] = {j0, j0 + 1 < n ? j0 + 1 : j0};
        LosPre P[2];
        RrmIntervals I[2];
        for (int q = 0; q < 2; ++q) {
            const int64_t j = jj[q], jo = n_obs == n ? j : 0, je = n_earth == n ? j : 0;
            rrm_pre(X, u[j], u[n + j], u[2 * n + j], obs[jo], obs[n_obs + jo], obs[2 * n_obs + jo], earth[je],
                    earth[n_earth + je], mask, P[q], I[q]);
        }
        integrate_rrm_x2(X, tab.data(), nodes.data(), P[0], P[1], I[0], I[1], [&](int ci, float a, float b) {
            out[ci * n + jj[1]] = b;
            out[ci * n + jj[0]] = a;
        });
    }
}

// fast: 0 generic routine, 1 scalar fused routine, 2 packed fused routines (fp32, no scattering).
// Returns 1 / 2 if the Kelsall fast path was eligible and used, 3 for the fused RRM routine, 0 if the
// generic routine ran.
extern "C" int zodi_emu_evaluate_mode(const zodi_model_desc* d, int precision, int lanes, int fast,
                                      int64_t n, const double* u, const double* obs, int64_t n_obs,
                                      const double* earth, int64_t n_earth, const uint8_t* flags,
                                      double* out);

extern "C" int zodi_emu_evaluate(const zodi_model_desc* d, int precision, int lanes, int64_t n,
                                 const double* u, const double* obs, int64_t n_obs,
                                 const double* earth, int64_t n_earth, const uint8_t* flags,
                                 double* out) {
    DevModel<double> m64;
    DevModel<float> m32;
    build_dev_model(*d, m64);
    narrow_model(m64, m32);
    std::vector<Pair<double>> t64, n64;
    std::vector<Pair<float>> t32, n32;
    build_pairs(*d, t64, n64, t32, n32);
    if (precision == ZODI_FP32) run<float>(d, m32, t32, n32, n, u, obs, n_obs, earth, n_earth, flags, lanes, out);
    else run<double>(d, m64, t64, n64, n, u, obs, n_obs, earth, n_earth, flags, lanes, out);
    return 0;
}

extern "C" int zodi_emu_evaluate_mode(const zodi_model_desc* d, int precision, int lanes, int fast,
                                      int64_t n, const double* u, const double* obs, int64_t n_obs,
                                      const double* earth, int64_t n_earth, const uint8_t* flags,
                                      double* out) {
    KelsallModel<double> k64;
    if (fast && d->kind == ZODI_RRM) {
        DevModel<double> m64;
        DevModel<float> m32;
        build_dev_model(*d, m64);
        narrow_model(m64, m32);
        RrmModel<double> r64;
        if (build_rrm_model(*d, m64, r64)) {
            std::vector<Pair<double>> t64, n64;
            std::vector<Pair<float>> t32, n32;
            build_pairs(*d, t64, n64, t32, n32);
            if (precision == ZODI_FP32) {
                RrmModel<float> r32;
                narrow_rrm(r64, m32, r32);
                RrmModelX2 x2;
                if (fast == 2 && build_rrm_x2(r64, r32, x2)) {
                    run_rrm_x2(x2, t32, n32, n, u, obs, n_obs, earth, n_earth, flags, out);
                    return 4;
                }
                run_rrm<float>(r32, t32, n32, n, u, obs, n_obs, earth, n_earth, flags, lanes, out);
            } else {
                run_rrm<double>(r64, t64, n64, n, u, obs, n_obs, earth, n_earth, flags, lanes, out);
            }
            return 3;
        }
    }
    if (!fast || !build_kelsall_model(*d, k64)) {
        zodi_emu_evaluate(d, precision, lanes, n, u, obs, n_obs, earth, n_earth, flags, out);
        return 0;
    }
    std::vector<Pair<double>> t64, n64;
    std::vector<Pair<float>> t32, n32;
    build_pairs(*d, t64, n64, t32, n32);
    if (precision == ZODI_FP32) {
        KelsallModel<float> k32;
        narrow_kelsall(k64, k32);
        if (fast == 2) {
            switch (lanes) {
                case 2: run_kelsall_x2<2>(k32, t32, n32, n, u, obs, n_obs, earth, n_earth, flags, out); break;
                case 4: run_kelsall_x2<4>(k32, t32, n32, n, u, obs, n_obs, earth, n_earth, flags, out); break;
                case 8: run_kelsall_x2<8>(k32, t32, n32, n, u, obs, n_obs, earth, n_earth, flags, out); break;
                default: run_kelsall_x2<1>(k32, t32, n32, n, u, obs, n_obs, earth, n_earth, flags, out);
            }
            return 2;
        }
        run_kelsall<float>(k32, t32, n32, n, u, obs, n_obs, earth, n_earth, flags, lanes, out);
    } else {
        run_kelsall<double>(k64, t64, n64, n, u, obs, n_obs, earth, n_earth, flags, lanes, out);
    }
    return 1;
}

// Multi-band routines (zodi_multiband.cuh / zodi_multiband_x2.cuh): out is (n_bands, n), component-summed.
// packed = 0: scalar routine (fp64 or fp32); 1: packed fp32 routine with the knot-major table rows the kernel
// stages.  Returns the padded band count of the instance that ran, -1 when the descriptors are not eligible.
template <typename Real, int NB>
static void run_multiband(const MultiBandModel<Real>& MB, const std::vector<Pair<Real>>& tabs,
                          const std::vector<Pair<Real>>& nodes, int64_t n, const double* u, const double* obs,
                          int64_t n_obs, const double* earth, int64_t n_earth, uint32_t mask, double* out) {
    for (int64_t j = 0; j < n; ++j) {
        const int64_t jo = n_obs == n ? j : 0, je = n_earth == n ? j : 0;
        auto emit = [&](int b, Real v) { if (b < MB.n_bands) out[b * n + j] = (double)v; };
#define ZMB(RF, SC) integrate_kelsall_multiband<Real, NB, RF, SC>(MB, tabs.data(), nodes.data(), u[j], u[n + j], \
        u[2 * n + j], obs[jo], obs[n_obs + jo], obs[2 * n_obs + jo], earth[je], earth[n_earth + je], mask, 0, 1, emit)
        if (MB.base.n_comps == 6) { if (MB.base.scatter) ZMB(true, true); else ZMB(true, false); }
        else { if (MB.base.scatter) ZMB(false, true); else ZMB(false, false); }
#undef ZMB
    }
}

template <int NB>
static void run_multiband_x2(const MultiBandModel<float>& MB, const std::vector<Pair<float>>& tabs,
                             const std::vector<Pair<float>>& nodes, int64_t n, const double* u, const double* obs,
                             int64_t n_obs, const double* earth, int64_t n_earth, uint32_t mask, double* out) {
    constexpr int kRow = MbRows<NB>::kRow;
    const int nt = MB.base.n_temps;
    // the kernel's staging loop: knot-major rows, bands past n_bands zero
    std::vector<Pair<float>> rows_store((size_t)nt * kRow + 2, Pair<float>{0.f, 0.f});
    Pair<float>* rows = rows_store.data();
    if (reinterpret_cast<uintptr_t>(rows) % 16) ++rows;  // BandPair reads are 16-byte aligned
    for (int i = 0; i < nt * kRow; ++i) {
        const int knot = i / kRow, b = i - knot * kRow;
        rows[i] = b < MB.n_bands ? tabs[(size_t)b * nt + knot] : Pair<float>{0.f, 0.f};
    }
    for (int64_t j0 = 0; j0 < n; j0 += 2) {
        const int64_t jj[2] = {j0, j0 + 1 < n ? j0 + 1 : j0};
        LosPre P[2];
        for (int q = 0; q < 2; ++q) {
            const int64_t j = jj[q], jo = n_obs == n ? j : 0, je = n_earth == n ? j : 0;
            if (MB.base.n_comps == 6)
                P[q] = los_pre<true>(MB.base, u[j], u[n + j], u[2 * n + j], obs[jo], obs[n_obs + jo],
                                     obs[2 * n_obs + jo], earth[je], earth[n_earth + je], mask);
            else
                P[q] = los_pre<false>(MB.base, u[j], u[n + j], u[2 * n + j], obs[jo], obs[n_obs + jo],
                                      obs[2 * n_obs + jo], earth[je], earth[n_earth + je], mask);
        }
        auto emit = [&](int b, float va, float vb) {
            if (b < MB.n_bands) { out[b * n + jj[1]] = (double)vb; out[b * n + jj[0]] = (double)va; }
        };
#define ZMB(RF, SC) integrate_multiband_x2<NB, RF, SC>(MB, rows, nodes.data(), P[0], P[1], emit)
        if (MB.base.n_comps == 6) { if (MB.base.scatter) ZMB(true, true); else ZMB(true, false); }
        else { if (MB.base.scatter) ZMB(false, true); else ZMB(false, false); }
#undef ZMB
    }
}

extern "C" int zodi_emu_multiband(const zodi_model_desc* descs, int n_bands, int precision, int packed, int64_t n,
                                  const double* u, const double* obs, int64_t n_obs, const double* earth,
                                  int64_t n_earth, const uint8_t* flags, double* out) {
    MultiBandModel<double> MB;
    MultiBandModel<float> MF;
    std::vector<Pair<double>> t64, n64, tb;
    std::vector<Pair<float>> t32, n32, tbf;
    if (n_bands < 1 || n_bands > kMaxBands || build_multiband_model(descs, n_bands, MB, MF, t64, t32) >= 0) return -1;
    build_pairs(descs[0], tb, n64, tbf, n32);
    uint32_t mask = 0;
    for (int c = 0; c < MB.base.n_comps; ++c) {
        if (flags[2 * c]) mask |= 1u << (2 * c);
        if (flags[2 * c + 1]) mask |= 1u << (2 * c + 1);
    }
#define ZNB(NB)                                                                                                  \
    do {                                                                                                         \
        if (precision != ZODI_FP32) run_multiband<double, NB>(MB, t64, n64, n, u, obs, n_obs, earth, n_earth, mask, out); \
        else if (packed) run_multiband_x2<NB>(MF, t32, n32, n, u, obs, n_obs, earth, n_earth, mask, out);         \
        else run_multiband<float, NB>(MF, t32, n32, n, u, obs, n_obs, earth, n_earth, mask, out);                 \
    } while (0)
    if (MB.n_bands_padded == 4) ZNB(4);
    else if (MB.n_bands_padded == 8) ZNB(8);
    else ZNB(16);
#undef ZNB
    return MB.n_bands_padded;
}

// Not-a-knot cubic spline coefficients (the routine zodi_ephemeris_create uses), scipy layout
// c[k * (n - 1) + i], k = 0..3 highest power first.
extern "C" int zodi_emu_spline(int n, const double* x, const double* y, double* c) {
    std::vector<double> xs(x, x + n), c0, c1, c2, c3;
    cubic_spline_not_a_knot(xs, y, c0, c1, c2, c3);
    for (int i = 0; i < n - 1; ++i) {
        c[0 * (n - 1) + i] = c0[i]; c[1 * (n - 1) + i] = c1[i];
        c[2 * (n - 1) + i] = c2[i]; c[3 * (n - 1) + i] = c3[i];
    }
    return 0;
}

// Element-wise access to the device math for unit tests: op 0 = Math<double>::log2_, 1 = exp2_,
// 2 = asin_unit (fp32), 3 = table_coord<double> fraction, 4 = table_coord<double> index,
// 5 = Math<double>::atan2_abs_(x[i], aux).
extern "C" int zodi_emu_math(int op, int64_t n, const double* x, double aux, double* y) {
    for (int64_t i = 0; i < n; ++i) {
        switch (op) {
            case 0: y[i] = Math<double>::log2_(x[i]); break;
            case 1: y[i] = Math<double>::exp2_(x[i]); break;
            case 2: y[i] = (double)asin_unit((float)x[i]); break;
            case 3: case 4: {
                int idx; double frac;
                table_coord<double>(x[i], aux, idx, frac);
                y[i] = op == 3 ? frac : (double)idx;
                break;
            }
            case 5: y[i] = Math<double>::atan2_abs_(x[i], aux); break;
            default: return 1;
        }
    }
    return 0;
}

// phase_of_cos<float> with the polynomial the host builds for (C1, C2, C3); returns terms kept.
extern "C" int zodi_emu_phase(double C1, double C2, double C3, int64_t n, const double* c, double* y) {
    double poly64[kPhaseTerms];
    float poly[kPhaseTerms];
    const int terms = phase_polynomial(C1, C2, C3, poly64);
    for (int k = 0; k < kPhaseTerms; ++k) poly[k] = (float)poly64[k];
    for (int64_t i = 0; i < n; ++i)
        y[i] = (double)phase_of_cos<float>((float)c[i], (float)C1, (float)C2, (float)(C3 * kLog2e), terms ? 1 : 0, terms, poly);
    return terms;
}
