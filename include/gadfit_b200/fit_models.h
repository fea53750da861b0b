// Fitting functions compiled into the library ("model registry").
//
// A fitting function is user code: in the reference it is the lambda handed
// to gadfit::LMsolver (c++/tests/example.cpp:11-19).  Here it is a functor
// templated on the evaluation context (fit_function.h), so that ONE source
// of each function is compiled three ways:
//   * into the sm_100a kernels (SDual / Dir2 / passive),
//   * against the UNMODIFIED reference's gadfit::AdVar (oracle/ref_harness.cpp),
//   * against the oracle's own dense forward-mode dual (oracle/oracle_solver.cpp),
// which is what makes "identical inputs" in the parity tests literal.
//
// Functions and parameter layouts (BASELINE.json configs):
//   exponential          I0*exp(-x/tau)+bgr              [I0, tau, bgr]
//   gauss_hermite3       sum of 3 distorted Gaussians    [A,mu,w,a,b,c] x3, b0, b1
//   gauss_linear_bg      A e^{-u^2}(1+s u)+b0+b1 x       [A, mu, w, s, b0, b1]
//   conv_exp_gauss       single-integral model           [A, tau, sg, t0, bgr]
//   lorentz_double_int   double-integral model           [A, x0, gamma, b, c, k1, k2]
//   single_integral_*    the reference's adaptive-integration test functions [p0, p1]
#pragma once

#include "fit_function.h"

#include <tuple>
#include <utility>

namespace gadfit_b200::models {

// c++/tests/lm_solver.cpp:11-19 (also c++/tests/example.cpp:11-19)
struct Exponential
{
    static constexpr int n_pars { 3 };
    static constexpr const char* name { "exponential" };
    static constexpr int chi2_unroll { 8 }; // tuning hint, see csrc/sweep_kernels.cuh
    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        const auto I0 { par<0>(c) };
        const auto tau { par<1>(c) };
        const auto bgr { par<2>(c) };
        return I0 * exp(-x / tau) + bgr;
    }
};

// BASELINE config 2: "sum-of-3-Gaussians model, 20 shared params".  Each
// peak is a Gaussian with a cubic Gauss-Hermite-style distortion,
//   A e^{-u^2} (1 + a u + b u^2 + c u^3),  u = (x - mu) / w,
// plus a linear background b0 + b1 x: 3 x 6 + 2 = 20 parameters.
struct GaussHermite3
{
    static constexpr int n_pars { 20 };
    static constexpr const char* name { "gauss_hermite3" };
    static constexpr int chi2_unroll { 8 };

    // 1 / w_k depends on the parameters only: once per data set, not per point
    // (the value of (x - mu) / w is (x - mu) * (1.0 / w) in the reference's AD
    // arithmetic too, automatic_differentiation.cpp:207-208)
    template <class Ctx>
    GADFIT_HD auto prepare(const Ctx& c) const
    {
        return pack(1.0 / par<2>(c), 1.0 / par<8>(c), 1.0 / par<14>(c));
    }

    template <int K, class Ctx, class Pre>
    GADFIT_HD static auto peak(const Ctx& c, const Pre& pre, const double x)
    {
        const auto u { (x - par<6 * K + 1>(c)) * pre.template get<K>() };
        const auto e { exp(-(u * u)) };
        const auto poly {
            1.0 + u * (par<6 * K + 3>(c) + u * (par<6 * K + 4>(c) + u * par<6 * K + 5>(c)))
        };
        return par<6 * K>(c) * e * poly;
    }

    template <class Ctx, class Pre>
    GADFIT_HD auto operator()(const Ctx& c, const Pre& pre, const double x) const
    {
        return peak<0>(c, pre, x) + peak<1>(c, pre, x) + peak<2>(c, pre, x) + par<18>(c)
               + par<19>(c) * x;
    }
};

// BASELINE config 3: "Gaussian peak + linear background: 4 local + 2 shared".
// Shared (w, s); local (A, mu, b0, b1).
struct GaussLinearBg
{
    static constexpr int n_pars { 6 };
    static constexpr const char* name { "gauss_linear_bg" };
    static constexpr int chi2_unroll { 8 };
    static constexpr int jacobian_blocks { 4 }; // tuning hints, see csrc/sweep_kernels.cuh
    static constexpr bool jacobian_in_registers { true };
    template <class Ctx>
    GADFIT_HD auto prepare(const Ctx& c) const
    {
        return pack(1.0 / par<2>(c));
    }
    template <class Ctx, class Pre>
    GADFIT_HD auto operator()(const Ctx& c, const Pre& pre, const double x) const
    {
        const auto u { (x - par<1>(c)) * pre.template get<0>() };
        return par<0>(c) * exp(-(u * u)) * (1.0 + par<3>(c) * u) + par<4>(c)
               + par<5>(c) * x;
    }
};

// BASELINE config 4: "model with per-point single integral (Gaussian-
// convolved exponential via fixed Gauss-Legendre[-Kronrod])":
//   A * int_0^T e^{-t/tau} e^{-((x-t-t0)/sg)^2} / (sg sqrt(pi)) dt + bgr,
// T = 10, one 15-node Kronrod panel.  Shared (tau, sg, t0); local (A, bgr).
struct ConvExpGauss
{
    static constexpr int n_pars { 5 };
    static constexpr const char* name { "conv_exp_gauss" };
    // chi2 sweep: 4 points per thread with 128 registers (measured, r02r: 528 ->
    // 471 us per sweep on config 4; the 225-node function of config 5 is fastest
    // with one point and 64 registers)
    static constexpr int chi2_unroll { 4 };
    static constexpr int chi2_blocks { 2 };
    // Jacobian sweep with 4 CTAs per SM (64 registers: 891 -> 848 us per sweep),
    // omega sweep with 3 (80 registers: 1028 -> 990 us), measured r02u; the
    // 225-node function of config 5 is fastest with the default 2 (128 registers)
    static constexpr int jacobian_blocks { 4 };
    static constexpr int omega_blocks { 3 };
    static constexpr double upper_bound { 10.0 };
    static constexpr double sqrt_pi { 1.7724538509055160272981674833411451 };

    // Parameter-only factors of the integrand: e^{-t_i/tau} at the 15 nodes
    // (get<0>), 1/sg (get<1>) and the normalisation 1/(sg sqrt(pi)) (get<2>)
    template <class Ctx>
    GADFIT_HD auto prepare(const Ctx& c) const
    {
        const auto tau { par<1>(c) };
        const auto sg { par<2>(c) };
        return pack(tabulate<k15_size>(
                      c, [&](const int i) { return exp(-kronrod_node(i, 0.0, upper_bound) / tau); }),
                    1.0 / sg,
                    1.0 / (sg * sqrt_pi));
    }
    template <class Ctx, class Pre>
    GADFIT_HD auto operator()(const Ctx& c, const Pre& pre, const double x) const
    {
        const auto t0 { par<3>(c) };
        const auto integral { c.kronrod_indexed(
          0.0, upper_bound, [&](const auto& t, const int i) {
              const auto z { (x - t - t0) * pre.template get<1>() };
              return pre.template get<0>()[i] * exp(-(z * z)) * pre.template get<2>();
          }) };
        return par<0>(c) * integral + par<4>(c);
    }
};

// BASELINE config 5: "Lorentzian + background with double-integral term,
// 3 local + 4 shared":
//   A g^2/((x-x0)^2+g^2) + b + c * int_0^1 int_0^1 e^{-k1 u^2 - k2 v^2}
//                                       / (1 + (x-u-v)^2) du dv,
// 15 x 15 Kronrod.  Shared (gamma, c, k1, k2); local (A, x0, b).
struct LorentzDoubleInt
{
    static constexpr int n_pars { 7 };
    static constexpr const char* name { "lorentz_double_int" };

    // e^{-k1 u_i^2 - k2 v_j^2} depends on the parameters and the node pair
    // (i, j) only: a table of 15 x 15 exponentials per DATA SET (one warp fills
    // it, seven entries per lane) instead of 225 exponentials -- or, with the
    // factorised form used until round 2, 225 products of two table entries
    // and their tangents -- per data POINT.
    // get<0>: e^{-k1 u_i^2 - k2 v_j^2} at [15 i + j] (inner node i, outer node
    // j), get<1>: gamma^2
    template <class Ctx>
    GADFIT_HD auto prepare(const Ctx& c) const
    {
        const auto g { par<2>(c) };
        const auto k1 { par<5>(c) };
        const auto k2 { par<6>(c) };
        const auto node_sq { [](const int i) {
            const double node { kronrod_node(i, 0.0, 1.0) };
            return node * node;
        } };
        return pack(tabulate<k15_size * k15_size>(c,
                                                  [&](const int ij) {
                                                      const int i { ij / k15_size };
                                                      const int j { ij - i * k15_size };
                                                      return exp(-(k1 * node_sq(i) + k2 * node_sq(j)));
                                                  }),
                    g * g);
    }
    template <class Ctx, class Pre>
    GADFIT_HD auto operator()(const Ctx& c, const Pre& pre, const double x) const
    {
        const auto dx { x - par<1>(c) };
        const auto& g2 { pre.template get<1>() };
        const auto lorentz { par<0>(c) * g2 / (dx * dx + g2) };
        const auto integral { c.kronrod2d_indexed(
          0.0, 1.0, 0.0, 1.0, [&](const auto& u, const auto& v, const int i, const int j) {
              const auto s { x - u - v };
              return pre.template get<0>()[i * k15_size + j] / (1.0 + s * s);
          }) };
        return lorentz + par<3>(c) + par<4>(c) * integral;
    }
};

// The reference's own adaptive-integration test functions ("Single integral",
// reference c++/tests/numerical_integration.cpp:25-223): eight combinations of
// constant / data-dependent / parameter-dependent integration bounds, with and
// without fitting parameters inside the integrand.  Parameters [p0, p1]; the
// constants are fix_d[0..2] of c++/tests/fixtures.h:9-11.
//   BOUND codes: 0 = 0.0, 1 = p0 / fix_d[0], 2 = p1, 3 = the data point x
//   INTEGRAND 0: t^p0 exp(-p1 t^2)          (:29-31)
//             1: x t^p0 exp(-p1 t^2)        (:46-48)
//             2: x t^fix_d[2] exp(-t^2)     (:74-76)
template <int ID, int LOWER, int UPPER, int INTEGRAND, int SIGN>
struct SingleIntegral
{
    static constexpr int n_pars { 2 };
    static constexpr double f0 { 6.1360420701563498 };
    static constexpr double f1 { 2.9606444748278875 };
    static constexpr double f2 { 9.9253736972586246 };
    static constexpr const char* names[8] {
        "single_integral_no_bounds",        "single_integral_lower",
        "single_integral_lower_nopar",      "single_integral_upper",
        "single_integral_upper_nopar",      "single_integral_both",
        "single_integral_both_lower_fixed", "single_integral_both_nopar"
    };
    static constexpr const char* name { names[ID] };

    template <int CODE, class Ctx>
    GADFIT_HD static auto bound(const Ctx& c, const double x)
    {
        if constexpr (CODE == 0) {
            return 0.0;
        } else if constexpr (CODE == 1) {
            return par<0>(c) / f0;
        } else if constexpr (CODE == 2) {
            return par<1>(c);
        } else {
            return x;
        }
    }

    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        const auto integrand { [x](const auto& cc, const auto& t) {
            if constexpr (INTEGRAND == 0) {
                return pow(t, par<0>(cc)) * exp(-par<1>(cc) * t * t);
            } else if constexpr (INTEGRAND == 1) {
                return x * pow(t, par<0>(cc)) * exp(-par<1>(cc) * t * t);
            } else {
                return x * pow(t, f2) * exp(-t * t);
            }
        } };
        return (SIGN * f1) * c.integrate(bound<LOWER>(c, x), bound<UPPER>(c, x), integrand, 1e-12);
    }
};
using SingleIntegral1 = SingleIntegral<0, 0, 3, 0, 1>;  // :25-45
using SingleIntegral2 = SingleIntegral<1, 1, 0, 1, -1>; // :46-72
using SingleIntegral3 = SingleIntegral<2, 1, 0, 2, -1>; // :73-95
using SingleIntegral4 = SingleIntegral<3, 0, 1, 1, 1>;  // :96-123
using SingleIntegral5 = SingleIntegral<4, 0, 1, 2, 1>;  // :124-147
using SingleIntegral6 = SingleIntegral<5, 1, 2, 1, -1>; // :148-175
using SingleIntegral7 = SingleIntegral<6, 2, 1, 1, -1>; // :176-199
using SingleIntegral8 = SingleIntegral<7, 1, 2, 2, -1>; // :200-222


// Calls f.template operator()<Model>() for every registered model; the
// library, the reference harness and the oracle all dispatch through this
// one list.  The GROUP variants exist so that the CUDA library can compile
// the (template-heavy) groups in separate translation units in parallel.
template <int GROUP, class F>
inline void for_each_model_in_group(F&& f)
{
    if constexpr (GROUP == 0) { // BASELINE workloads
        f.template operator()<Exponential>();
        f.template operator()<GaussHermite3>();
        f.template operator()<GaussLinearBg>();
        f.template operator()<ConvExpGauss>();
        f.template operator()<LorentzDoubleInt>();
    } else if constexpr (GROUP == 1) { // adaptive single integrals
        f.template operator()<SingleIntegral1>();
        f.template operator()<SingleIntegral2>();
        f.template operator()<SingleIntegral3>();
        f.template operator()<SingleIntegral4>();
    } else if constexpr (GROUP == 2) {
        f.template operator()<SingleIntegral5>();
        f.template operator()<SingleIntegral6>();
        f.template operator()<SingleIntegral7>();
        f.template operator()<SingleIntegral8>();
    }
}
constexpr int n_model_groups { 3 };

template <class F, int... GROUPS>
inline void for_each_model_impl(F&& f, std::integer_sequence<int, GROUPS...>)
{
    (for_each_model_in_group<GROUPS>(f), ...);
}
template <class F>
inline void for_each_model(F&& f)
{
    for_each_model_impl(f, std::make_integer_sequence<int, n_model_groups> {});
}

} // namespace gadfit_b200::models
