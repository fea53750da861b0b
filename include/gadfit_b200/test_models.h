// Fitting functions of the reference's nested and direct double-integral tests
// (reference c++/tests/numerical_integration.cpp:240-1440).  They exist to pin
// the adaptive integrators against the reference's own known answers; they are
// NOT compiled into libgadfit_b200.so but into a model plug-in
// (gadfit_b200/csrc/test_models_plugin.cu -> libgadfit_b200_testmodels.so,
// loaded with gadfit_b200_load_models), which doubles as the worked example of
// how user-defined fitting functions are added.
#pragma once

#include "fit_function.h"

#include <tuple>
#include <utility>

namespace gadfit_b200::models {

// The reference's nested double-integral test function ("Double integral
// (nested)", reference c++/tests/numerical_integration.cpp:240-918): an
// adaptive integral whose integrand contains another adaptive integral, with
// fitting parameters in all four bounds.  Parameters [p0 .. p5]:
//   f(x) = 1/x * int_{OL}^{OH} e^{-u} [ int_{IL}^{IH} log((e^t - 0.9) q(u) + 1)/t dt ] du,
//   q(u) = 1 + p0 p1 erf(u)
// Bound codes -- inner: 0 = p4, 1 = p5 x / p1, 2 = p5 fix_d[2], 3 = fix_d[3],
// 4 = 1 / fix_d[16], 5 = p5;  outer: 0 = p4 (p1 - p2), 1 = p3, 2 = p4 p2,
// 3 = p3 / p5, 4 = p4.  SIGN -1: the outer integrand is -e^{-u} [...] (the
// sections that swap the inner bounds); OSIGN -1: f = -1/x * int (the
// sections that swap the outer bounds).  (In the reference test the outer
// integrand sees the renamed vector pars2 = [p0, p1, x, p4, p5].)
template <int ID, int IL, int IH, int OL, int OH, int SIGN = 1, int OSIGN = 1>
struct NestedIntegral
{
    static constexpr int n_pars { 6 };
    static constexpr double tol_inner { 1e-3 };
    static constexpr double tol_outer { 1e-2 };
    static constexpr double f2 { 9.9253736972586246 };  // fix_d[2], fixtures.h:9-17
    static constexpr double f3 { 0.5356792380861322 };  // fix_d[3]
    static constexpr double f16 { 32.898784649467867 }; // fix_d[16]
    static constexpr const char* names[12] {
        "nested_integral_y1y2x1x2", "nested_integral_x1x2", "nested_integral_no_active_bounds",
        "nested_integral_y1y2x1",   "nested_integral_y1y2x2", "nested_integral_y2x1x2",
        "nested_integral_y1y2",     "nested_integral_y2x1", "nested_integral_y1x1",
        "nested_integral_y1",       "nested_integral_y2",   "nested_integral_x1"
    };
    static constexpr const char* name { names[ID] };

    template <int CODE, class Ctx>
    GADFIT_HD static auto inner_bound(const Ctx& c, const double x)
    {
        if constexpr (CODE == 0) {
            return par<4>(c);
        } else if constexpr (CODE == 1) {
            return par<5>(c) * x / par<1>(c);
        } else if constexpr (CODE == 2) {
            return par<5>(c) * f2;
        } else if constexpr (CODE == 3) {
            return f3;
        } else if constexpr (CODE == 4) {
            return 1 / f16;
        } else {
            return par<5>(c);
        }
    }
    template <int CODE, class Ctx>
    GADFIT_HD static auto outer_bound(const Ctx& c)
    {
        if constexpr (CODE == 0) {
            return par<4>(c) * (par<1>(c) - par<2>(c));
        } else if constexpr (CODE == 1) {
            return par<3>(c);
        } else if constexpr (CODE == 2) {
            return par<4>(c) * par<2>(c);
        } else if constexpr (CODE == 3) {
            return par<3>(c) / par<5>(c);
        } else {
            return par<4>(c);
        }
    }

    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        const auto outer { [x](const auto& co, const auto& u) {
            // :258-266: the inner integrand's only parameter is q(u)
            const auto inner { [&u](const auto& ci, const auto& t) {
                const auto q { 1 + par<0>(ci) * par<1>(ci) * erf(u) };
                return log((exp(t) - 0.9) * q + 1.0) / t;
            } };
            const auto inner_integral { co.integrate(
              inner_bound<IL>(co, x), inner_bound<IH>(co, x), inner, tol_inner) };
            if constexpr (SIGN > 0) {
                return exp(-u) * inner_integral;
            } else {
                return -exp(-u) * inner_integral;
            }
        } };
        const auto outer_integral { c.integrate(
          outer_bound<OL>(c), outer_bound<OH>(c), outer, tol_outer) };
        if constexpr (OSIGN > 0) {
            return outer_integral / x;
        } else {
            return -outer_integral / x;
        }
    }
};
using NestedIntegral1 = NestedIntegral<0, 0, 1, 0, 1>; // :249-291 y1 y2 x1 x2, :379-421 y1 x1 x2
using NestedIntegral2 = NestedIntegral<1, 0, 1, 2, 1>; // :505-547 x1 x2
using NestedIntegral3 = NestedIntegral<2, 0, 1, 0, 3>; // :874-917 no active bounds
using NestedIntegral4 = NestedIntegral<3, 0, 2, 0, 1>; // :292-335 y1 y2 x1
using NestedIntegral5 = NestedIntegral<4, 3, 1, 0, 1>; // :336-378 y1 y2 x2, :548-590 y1 x2
using NestedIntegral6 = NestedIntegral<5, 0, 1, 1, 0, 1, -1>; // :422-464 y2 x1 x2, :671-713 y2 x2
using NestedIntegral7 = NestedIntegral<6, 4, 5, 0, 1>; // :465-504 y1 y2
using NestedIntegral8 = NestedIntegral<7, 5, 0, 4, 1, -1>;  // :591-627 y2 x1
using NestedIntegral9 = NestedIntegral<8, 1, 0, 0, 1, -1>;  // :628-670 y1 x1
using NestedIntegral10 = NestedIntegral<9, 0, 5, 0, 1>;     // :714-753 y1, :834-873 x2
using NestedIntegral11 = NestedIntegral<10, 0, 5, 1, 0, 1, -1>; // :754-793 y2
using NestedIntegral12 = NestedIntegral<11, 5, 0, 0, 1, -1>; // :794-833 x1

// The reference's direct double-integral test function ("Double integral
// (direct)", reference c++/tests/numerical_integration.cpp:930-1440): one
// 2-D adaptive integral over a rectangle whose four edges depend on the
// fitting parameters [p0 .. p6]:
//   f(x) = 1/x * int_{y1}^{y2} int_{x1}^{x2} e^{-v} log((e^u - 0.9) q(v) + 1)/u du dv,
//   q(v) = 1 + p0 p1 erf(v)
// Bound codes (".val" = the bound enters as a plain number, as in the
// reference sections that switch individual Leibniz terms off):
//   y1: 0 = p4 (p1 - p2), 1 = same .val, 2 = (p1 - p2).val
//   y2: 0 = p3 p6,        1 = same .val, 2 = p3
//   x1: 0 = p4 p6,        1 = same .val
//   x2: 0 = p5 / p1,      1 = same .val, 2 = p5.val
template <int ID, int Y1, int Y2, int X1, int X2>
struct DirectIntegral
{
    static constexpr int n_pars { 7 };
    static constexpr double tolerance { 1e-4 };
    static constexpr const char* names[15] {
        "direct_integral_y1y2x1x2", "direct_integral_y1y2x1", "direct_integral_y1y2x2",
        "direct_integral_y1x1x2",   "direct_integral_y2x1x2", "direct_integral_y1y2",
        "direct_integral_x1x2",     "direct_integral_y1x2",   "direct_integral_y2x1",
        "direct_integral_y1x1",     "direct_integral_y2x2",   "direct_integral_y1",
        "direct_integral_y2",       "direct_integral_x1",     "direct_integral_x2"
    };
    static constexpr const char* name { names[ID] };

    template <class Ctx>
    GADFIT_HD static auto y1(const Ctx& c)
    {
        if constexpr (Y1 == 0) {
            return par<4>(c) * (par<1>(c) - par<2>(c));
        } else if constexpr (Y1 == 1) {
            return (par<4>(c) * (par<1>(c) - par<2>(c))).val;
        } else {
            return (par<1>(c) - par<2>(c)).val;
        }
    }
    template <class Ctx>
    GADFIT_HD static auto y2(const Ctx& c)
    {
        if constexpr (Y2 == 0) {
            return par<3>(c) * par<6>(c);
        } else if constexpr (Y2 == 1) {
            return (par<3>(c) * par<6>(c)).val;
        } else {
            return par<3>(c);
        }
    }
    template <class Ctx>
    GADFIT_HD static auto x1(const Ctx& c)
    {
        if constexpr (X1 == 0) {
            return par<4>(c) * par<6>(c);
        } else {
            return (par<4>(c) * par<6>(c)).val;
        }
    }
    template <class Ctx>
    GADFIT_HD static auto x2(const Ctx& c)
    {
        if constexpr (X2 == 0) {
            return par<5>(c) / par<1>(c);
        } else if constexpr (X2 == 1) {
            return (par<5>(c) / par<1>(c)).val;
        } else {
            return par<5>(c).val;
        }
    }

    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        const auto integrand { [](const auto& cc, const auto& u, const auto& v) {
            const auto q { 1 + par<0>(cc) * par<1>(cc) * erf(v) };
            return exp(-v) * log((exp(u) - 0.9) * q + 1.0) / u;
        } };
        return c.integrate2d(y1(c), y2(c), x1(c), x2(c), integrand, tolerance) / x;
    }
};
using DirectIntegral1 = DirectIntegral<0, 0, 0, 0, 0>;   // :951-971 and :1409-1440
using DirectIntegral2 = DirectIntegral<1, 0, 0, 0, 1>;   // :972-1002
using DirectIntegral3 = DirectIntegral<2, 0, 0, 1, 0>;   // :1003-1033
using DirectIntegral4 = DirectIntegral<3, 0, 1, 0, 0>;   // :1034-1064
using DirectIntegral5 = DirectIntegral<4, 1, 0, 0, 0>;   // :1065-1095
using DirectIntegral6 = DirectIntegral<5, 0, 0, 1, 2>;   // :1096-1126
using DirectIntegral7 = DirectIntegral<6, 2, 1, 0, 0>;   // :1127-1157
using DirectIntegral8 = DirectIntegral<7, 0, 1, 1, 0>;   // :1158-1188
using DirectIntegral9 = DirectIntegral<8, 1, 2, 0, 1>;   // :1189-1219
using DirectIntegral10 = DirectIntegral<9, 0, 1, 0, 1>;  // :1220-1250
using DirectIntegral11 = DirectIntegral<10, 1, 0, 1, 0>; // :1251-1281
using DirectIntegral12 = DirectIntegral<11, 0, 1, 1, 2>; // :1282-1312
using DirectIntegral13 = DirectIntegral<12, 1, 0, 1, 1>; // :1313-1344
using DirectIntegral14 = DirectIntegral<13, 2, 1, 0, 1>; // :1345-1376
using DirectIntegral15 = DirectIntegral<14, 1, 1, 1, 0>; // :1377-1408

// The fitting functions of the reference's Fortran tests
// (fortran/tests/1_gaussian.F90:29-33, 2_integral_single.F90:27-46); they pin
// the gadf_fit driver (gadfit_b200_gadf_options) on the Fortran version's own
// known answers.
struct FortranGaussian // [fmax, x0, a, bgr]
{
    static constexpr int n_pars { 4 };
    static constexpr const char* name { "fortran_gaussian" };
    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        const auto u { (x - par<1>(c)) / par<2>(c) };
        return par<0>(c) * exp(-(u * u)) + par<3>(c);
    }
};
struct FortranIntegralSingle // [a, b]:  pi int_0^x t^a e^{-b t^2} dt
{
    static constexpr int n_pars { 2 };
    static constexpr const char* name { "fortran_integral_single" };
    static constexpr double pi { 3.14159265358979323846264338327950 };
    static constexpr double rel_error { 1e-12 }; // gadf_init(f, rel_error=1e-12_kp), :57
    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        const auto integrand { [](const auto& cc, const auto& t) {
            return pow(t, par<0>(cc)) * exp(-par<1>(cc) * (t * t));
        } };
        return pi * c.integrate(0.0, x, integrand, rel_error);
    }
};
// fortran/tests/3_integral_double.F90:27-61, [a, b]:
//   f(x) = 1/x int_0^inf e^{-t} [ int_0^{x/b} log((e^u - 1) q(t) + 1)/u du ] dt,
//   q(t) = 1 + b a erf(t);  gadf_init(rel_error_inner=1e-6, rel_error=1e-5), :75-76
struct FortranIntegralDouble
{
    static constexpr int n_pars { 2 };
    static constexpr const char* name { "fortran_integral_double" };
    static constexpr double rel_error_inner { 1e-6 };
    static constexpr double rel_error_outer { 1e-5 };
    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        const auto outer { [x](const auto& co, const auto& t) {
            const auto inner { [&t](const auto& ci, const auto& u) {
                const auto q { 1 + par<1>(ci) * par<0>(ci) * erf(t) };
                return log((exp(u) - 1.0) * q + 1.0) / u;
            } };
            return exp(-t) * co.integrate(0.0, x / par<1>(co), inner, rel_error_inner);
        } };
        return integrate_to_infinity(c, 0.0, outer, rel_error_outer) / x;
    }
};

using TestModels = std::tuple<NestedIntegral1,
                              NestedIntegral2,
                              NestedIntegral3,
                              NestedIntegral4,
                              NestedIntegral5,
                              NestedIntegral6,
                              NestedIntegral7,
                              NestedIntegral8,
                              NestedIntegral9,
                              NestedIntegral10,
                              NestedIntegral11,
                              NestedIntegral12,
                              DirectIntegral1,
                              DirectIntegral2,
                              DirectIntegral3,
                              DirectIntegral4,
                              DirectIntegral5,
                              DirectIntegral6,
                              DirectIntegral7,
                              DirectIntegral8,
                              DirectIntegral9,
                              DirectIntegral10,
                              DirectIntegral11,
                              DirectIntegral12,
                              DirectIntegral13,
                              DirectIntegral14,
                              DirectIntegral15,
                              FortranGaussian,
                              FortranIntegralSingle,
                              FortranIntegralDouble>;
constexpr int n_test_models { static_cast<int>(std::tuple_size_v<TestModels>) };

template <class F, int... I>
inline void for_each_test_model_impl(F&& f, std::integer_sequence<int, I...>)
{
    (f.template operator()<std::tuple_element_t<I, TestModels>>(), ...);
}
template <class F>
inline void for_each_test_model(F&& f)
{
    for_each_test_model_impl(f, std::make_integer_sequence<int, n_test_models> {});
}

} // namespace gadfit_b200::models
