// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.
//
// CPU restatement of the reference's forward-mode AD variable: value, first
// and second directional derivative, with the activity decided at RUN time
// per operand exactly as gadfit::AdVar does with its `idx` field (reference
// c++/gadfit/automatic_differentiation.h:29-83).  Only the forward-mode and
// passive branches of every elemental operation are restated
// (automatic_differentiation.cpp:59-680, automatic_differentiation.h:174-425);
// the reverse tape is not: the oracle obtains Jacobian column k as the first
// directional derivative along the unit vector e_k (one forward pass per
// active parameter), which the reference's own tests show to be the same
// numbers as its reverse sweep (c++/tests/ad_forward_mode.cpp:43 ==
// c++/tests/ad_reverse_mode.cpp:15).
//
// Independent of the product's include/gadfit_b200/dual.h on purpose.
#pragma once

#include <cmath>
#include <numbers>
#include <type_traits>
#include <vector>

namespace gadfit_oracle {

template <class T>
concept Number = std::is_arithmetic_v<T>;

struct FVar
{
    double val { 0.0 };
    double d { 0.0 };
    double dd { 0.0 };
    bool on { false }; // forward active (AdVar::idx == forward_active_idx)
};

inline auto passive(const double val) -> FVar
{
    return { val, 0.0, 0.0, false };
}

// Comparisons, automatic_differentiation.h:148-171
inline auto operator>(const FVar& x, const FVar& y) -> bool { return x.val > y.val; }
inline auto operator>(const FVar& x, const Number auto y) -> bool { return x.val > static_cast<double>(y); }
inline auto operator>(const Number auto y, const FVar& x) -> bool { return static_cast<double>(y) > x.val; }
inline auto operator<(const FVar& x, const FVar& y) -> bool { return x.val < y.val; }
inline auto operator<(const FVar& x, const Number auto y) -> bool { return x.val < static_cast<double>(y); }
inline auto operator<(const Number auto y, const FVar& x) -> bool { return static_cast<double>(y) < x.val; }

// Addition, automatic_differentiation.cpp:59-105
inline auto operator+(const FVar& x1, const FVar& x2) -> FVar
{
    FVar y { passive(x1.val + x2.val) };
    if (x1.on && x2.on) {
        y = { y.val, x1.d + x2.d, x1.dd + x2.dd, true };
    } else if (x1.on) {
        y = { y.val, x1.d, x1.dd, true };
    } else if (x2.on) {
        y = { y.val, x2.d, x2.dd, true };
    }
    return y;
}
// automatic_differentiation.h:174-214
inline auto operator+(const FVar& x1, const Number auto x2) -> FVar
{
    return { x1.val + x2, x1.on ? x1.d : 0.0, x1.on ? x1.dd : 0.0, x1.on };
}
inline auto operator+(const Number auto x1, const FVar& x2) -> FVar
{
    return { x1 + x2.val, x2.on ? x2.d : 0.0, x2.on ? x2.dd : 0.0, x2.on };
}
inline auto operator+(const FVar& x) -> FVar { return x; }

// Subtraction, automatic_differentiation.cpp:107-149, .h:226-267
inline auto operator-(const FVar& x1, const FVar& x2) -> FVar
{
    FVar y { passive(x1.val - x2.val) };
    if (x1.on && x2.on) {
        y = { y.val, x1.d - x2.d, x1.dd - x2.dd, true };
    } else if (x1.on) {
        y = { y.val, x1.d, x1.dd, true };
    } else if (x2.on) {
        y = { y.val, -x2.d, -x2.dd, true };
    }
    return y;
}
inline auto operator-(const FVar& x1, const Number auto x2) -> FVar
{
    return { x1.val - x2, x1.on ? x1.d : 0.0, x1.on ? x1.dd : 0.0, x1.on };
}
inline auto operator-(const Number auto x1, const FVar& x2) -> FVar
{
    return { x1 - x2.val, x2.on ? -x2.d : 0.0, x2.on ? -x2.dd : 0.0, x2.on };
}

// Multiplication, automatic_differentiation.cpp:151-201, .h:279-323
inline auto operator*(const FVar& x1, const FVar& x2) -> FVar
{
    FVar y { passive(x1.val * x2.val) };
    if (x1.on && x2.on) {
        y.d = x1.d * x2.val + x1.val * x2.d;
        y.dd = x1.dd * x2.val + 2 * x1.d * x2.d + x1.val * x2.dd;
        y.on = true;
    } else if (x1.on) {
        y = { y.val, x1.d * x2.val, x1.dd * x2.val, true };
    } else if (x2.on) {
        y = { y.val, x1.val * x2.d, x1.val * x2.dd, true };
    }
    return y;
}
inline auto operator*(const FVar& x1, const Number auto x2) -> FVar
{
    return { x1.val * x2, x1.on ? x1.d * x2 : 0.0, x1.on ? x1.dd * x2 : 0.0, x1.on };
}
inline auto operator*(const Number auto x1, const FVar& x2) -> FVar
{
    return { x1 * x2.val, x2.on ? x1 * x2.d : 0.0, x2.on ? x1 * x2.dd : 0.0, x2.on };
}
// automatic_differentiation.cpp:39-42
inline auto operator-(const FVar& x) -> FVar { return -1 * x; }

// Division, automatic_differentiation.cpp:203-248, .h:329-373
inline auto operator/(const FVar& x1, const FVar& x2) -> FVar
{
    const double inv_x2 { 1.0 / x2.val };
    FVar y { passive(x1.val * inv_x2) };
    if (x1.on && x2.on) {
        y.d = (x1.d - x2.d * y.val) * inv_x2;
        y.dd = (x1.dd - x2.dd * y.val - 2 * y.d * x2.d) * inv_x2;
        y.on = true;
    } else if (x1.on) {
        y = { y.val, x1.d * inv_x2, x1.dd * inv_x2, true };
    } else if (x2.on) {
        y.d = -y.val * x2.d * inv_x2;
        y.dd = (-x2.dd * y.val - 2 * y.d * x2.d) * inv_x2;
        y.on = true;
    }
    return y;
}
inline auto operator/(const FVar& x1, const Number auto x2) -> FVar
{
    const double inv_x2 { 1.0 / x2 };
    return { x1.val * inv_x2, x1.on ? x1.d * inv_x2 : 0.0, x1.on ? x1.dd * inv_x2 : 0.0, x1.on };
}
inline auto operator/(const Number auto x1, const FVar& x2) -> FVar
{
    const double inv_x2 { 1.0 / x2.val };
    FVar y { passive(x1 * inv_x2) };
    if (x2.on) {
        y.d = -y.val * x2.d * inv_x2;
        y.dd = (-x2.dd * y.val - 2 * y.d * x2.d) * inv_x2;
        y.on = true;
    }
    return y;
}

// Exponentiation, automatic_differentiation.cpp:250-305, .h:380-425
inline auto pow(const FVar& x1, const FVar& x2) -> FVar
{
    FVar y { passive(std::pow(x1.val, x2.val)) };
    if (x1.on && x2.on) {
        const double log_value { std::log(x1.val) };
        y.d = x1.d * x2.val * std::pow(x1.val, x2.val - 1) + x2.d * log_value * y.val;
        const double inv_x1 { 1.0 / x1.val };
        y.dd = y.d * y.d / y.val
               + y.val
                   * (x2.dd * log_value
                      + (2 * x1.d * x2.d + x2.val * (x1.dd - x1.d * x1.d * inv_x1)) * inv_x1);
        y.on = true;
    } else if (x1.on) {
        y.d = x1.d * x2.val * std::pow(x1.val, x2.val - 1);
        const double inv_x1 { 1.0 / x1.val };
        y.dd = y.d * y.d / y.val + y.val * x2.val * (x1.dd - x1.d * x1.d * inv_x1) * inv_x1;
        y.on = true;
    } else if (x2.on) {
        const double log_value { std::log(x1.val) };
        y.d = y.val * x2.d * log_value;
        y.dd = y.d * y.d / y.val + y.val * x2.dd * log_value;
        y.on = true;
    }
    return y;
}
inline auto pow(const FVar& x1, const Number auto x2) -> FVar
{
    FVar y { passive(std::pow(x1.val, x2)) };
    if (x1.on) {
        y.d = x1.d * x2 * std::pow(x1.val, x2 - 1);
        const double inv_x1 { 1.0 / x1.val };
        y.dd = y.d * y.d / y.val + y.val * x2 * (x1.dd - x1.d * x1.d * inv_x1) * inv_x1;
        y.on = true;
    }
    return y;
}
inline auto pow(const Number auto x1, const FVar& x2) -> FVar
{
    FVar y { passive(std::pow(x1, x2.val)) };
    if (x2.on) {
        const double log_value { std::log(x1) };
        y.d = y.val * x2.d * log_value;
        y.dd = y.d * y.d / y.val + y.val * x2.dd * log_value;
        y.on = true;
    }
    return y;
}

// y = g(x) with g' and g'' folded in by the caller: helper for the unary
// functions whose (d, dd) the reference writes out one by one
#define GADFIT_ORACLE_UNARY(NAME, VALUE, BODY)                                 \
    inline auto NAME(const FVar& x) -> FVar                                   \
    {                                                                          \
        FVar y { passive(VALUE) };                                             \
        if (x.on) {                                                            \
            BODY y.on = true;                                                  \
        }                                                                      \
        return y;                                                              \
    }

// automatic_differentiation.cpp:307-391
GADFIT_ORACLE_UNARY(log, std::log(x.val), const double inv_x { 1.0 / x.val }; y.d = x.d * inv_x;
                    y.dd = (x.dd - x.d * y.d) * inv_x;)
GADFIT_ORACLE_UNARY(exp, std::exp(x.val), y.d = x.d * y.val; y.dd = x.dd * y.val + x.d * y.d;)
GADFIT_ORACLE_UNARY(sqrt, std::sqrt(x.val), const double inv_y { 1.0 / y.val };
                    y.d = x.d / 2 * inv_y;
                    y.dd = (x.dd * inv_y - y.d * x.d / x.val) / 2;)
GADFIT_ORACLE_UNARY(abs, std::abs(x.val), y.d = x.d * (std::signbit(x.val) ? -1 : 1);
                    y.dd = x.dd * (std::signbit(x.val) ? -1 : 1);)
// automatic_differentiation.cpp:393-657
GADFIT_ORACLE_UNARY(sin, std::sin(x.val), const double c { std::cos(x.val) }; y.d = x.d * c;
                    y.dd = x.dd * c - x.d * x.d * y.val;)
GADFIT_ORACLE_UNARY(cos, std::cos(x.val), const double s { -std::sin(x.val) }; y.d = x.d * s;
                    y.dd = x.dd * s - x.d * x.d * y.val;)
GADFIT_ORACLE_UNARY(tan, std::tan(x.val), const double ic { 1.0 / std::cos(x.val) };
                    const double ic2 { ic * ic };
                    y.d = x.d * ic2;
                    y.dd = x.dd * ic2 + 2 * x.d * y.val * y.d;)
GADFIT_ORACLE_UNARY(asin, std::asin(x.val), const double t { 1.0 / std::sqrt(1 - x.val * x.val) };
                    y.d = x.d * t;
                    y.dd = t * (x.dd + x.val * y.d * y.d);)
GADFIT_ORACLE_UNARY(acos, std::acos(x.val), const double t { -1.0 / std::sqrt(1 - x.val * x.val) };
                    y.d = x.d * t;
                    y.dd = t * (x.dd + x.val * y.d * y.d);)
GADFIT_ORACLE_UNARY(atan, std::atan(x.val), const double t { 1.0 / (1 + x.val * x.val) };
                    y.d = x.d * t;
                    y.dd = x.dd * t - 2 * x.val * y.d * y.d;)
GADFIT_ORACLE_UNARY(sinh, std::sinh(x.val), const double c { std::cosh(x.val) }; y.d = x.d * c;
                    y.dd = x.dd * c + x.d * x.d * y.val;)
GADFIT_ORACLE_UNARY(cosh, std::cosh(x.val), const double s { std::sinh(x.val) }; y.d = x.d * s;
                    y.dd = x.dd * s + x.d * x.d * y.val;)
GADFIT_ORACLE_UNARY(tanh, std::tanh(x.val), const double ic { 1.0 / std::cosh(x.val) };
                    const double ic2 { ic * ic };
                    y.d = x.d * ic2;
                    y.dd = x.dd * ic2 - 2 * x.d * y.val * y.d;)
GADFIT_ORACLE_UNARY(asinh, std::asinh(x.val), const double t { 1.0 / std::sqrt(x.val * x.val + 1) };
                    y.d = x.d * t;
                    y.dd = (x.dd - x.val * y.d * y.d) * t;)
GADFIT_ORACLE_UNARY(acosh, std::acosh(x.val), const double t { 1.0 / std::sqrt(x.val * x.val - 1) };
                    y.d = x.d * t;
                    y.dd = (x.dd - x.val * y.d * y.d) * t;)
GADFIT_ORACLE_UNARY(atanh, std::atanh(x.val), const double t { 1.0 / (1 - x.val * x.val) };
                    y.d = x.d * t;
                    y.dd = (x.dd + 2 * x.val * x.d * y.d) * t;)
// automatic_differentiation.cpp:659-680
GADFIT_ORACLE_UNARY(erf, std::erf(x.val),
                    const double t { 2 * std::numbers::inv_sqrtpi * std::exp(-x.val * x.val) };
                    y.d = x.d * t;
                    y.dd = (x.dd - 2 * x.d * x.d * x.val) * t;)

#undef GADFIT_ORACLE_UNARY

// Fixed 15-point Kronrod panel, numerical_integration.cpp:133-148 (1-D) and
// :189-215 (2-D); nodes and weights are those of the rule, in the reference's
// order (gauss_kronrod_parameters.h:20-47: 7 negative, 0, 7 positive).
inline constexpr double k15_roots[15] {
    -0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
    -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
    -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
    -0.207784955007898467600689403773245, 0.000000000000000000000000000000000,
    0.991455371120812639206854697526329,  0.949107912342758524526189684047851,
    0.864864423359769072789712788640926,  0.741531185599394439863864773280788,
    0.586087235467691130294144838258730,  0.405845151377397166906606412076961,
    0.207784955007898467600689403773245
};
inline constexpr double k15_weights[15] {
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649, 0.209482141084727828012999174891714,
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
    0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
    0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
    0.204432940075298892414161999234649
};

template <class F>
auto kronrod(const double lower, const double upper, F&& f) -> FVar
{
    const double scale { (upper - lower) / 2 };
    const double shift { (lower + upper) / 2 };
    FVar sum { passive(0.0) };
    for (int i = 0; i < 15; ++i) {
        sum = sum + k15_weights[i] * f(passive(scale * k15_roots[i] + shift));
    }
    return scale * sum;
}

template <class F>
auto kronrod2d(const double outer_lower,
               const double outer_upper,
               const double inner_lower,
               const double inner_upper,
               F&& f) -> FVar
{
    const double outer_scale { (outer_upper - outer_lower) / 2 };
    const double outer_shift { (outer_lower + outer_upper) / 2 };
    const double inner_scale { (inner_upper - inner_lower) / 2 };
    const double inner_shift { (inner_lower + inner_upper) / 2 };
    FVar sum { passive(0.0) };
    for (int i = 0; i < 15; ++i) {
        const FVar arg_x { passive(inner_scale * k15_roots[i] + inner_shift) };
        for (int j = 0; j < 15; ++j) {
            const FVar arg_y { passive(outer_scale * k15_roots[j] + outer_shift) };
            sum = sum + (k15_weights[i] * k15_weights[j]) * f(arg_x, arg_y);
        }
    }
    return outer_scale * inner_scale * sum;
}

inline constexpr double g7_weights[7] {
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
    0.381830050505118944950369775488975, 0.417959183673469387755102040816327,
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
    0.381830050505118944950369775488975
};

// G7/K15 of one panel with passive parameters, numerical_integration.cpp:108-131
template <class F>
auto gaussKronrod(F&& f, const double lower, const double upper, double& abs_error) -> double
{
    const double scale { (upper - lower) / 2 };
    const double shift { (lower + upper) / 2 };
    double sum_kronrod {};
    double sum_gauss {};
    for (int i = 0; i < 15; ++i) {
        const double val { f(passive(scale * k15_roots[i] + shift)).val };
        sum_kronrod += k15_weights[i] * val;
        if (i % 2 == 1) {
            sum_gauss += g7_weights[i / 2] * val;
        }
    }
    sum_kronrod *= scale;
    abs_error = std::abs(sum_kronrod - scale * sum_gauss);
    return sum_kronrod;
}

// Evaluation context over a parameter vector of FVar (the role the
// std::vector<AdVar> argument plays in the reference's fitSignature,
// fit_function.h:27-28)
struct Ctx
{
    const FVar* pars;
    int n_pars { 0 };
    template <int K>
    auto get() const -> const FVar&
    {
        return pars[K];
    }
    // Adaptive integration, numerical_integration.cpp:242-306, and the
    // forward-mode Leibniz terms of AD-valued bounds, :336-343 / :369-376.
    // f(ctx, t) receives the passive copy of the parameters during refinement
    // (setInactiveParameters, :221-240) and the caller's for the final pass.
    template <class F>
    auto integrate(const double lower,
                   const double upper,
                   F&& f,
                   const double rel_error = 2.220446049250313e-14,
                   const double abs_error = 0.0) const -> FVar
    {
        FVar inactive[64];
        for (int i = 0; i < n_pars; ++i) {
            inactive[i] = passive(pars[i].val);
        }
        const Ctx pc { inactive, n_pars };
        const auto fp { [&](const FVar& t) { return f(pc, t); } };
        constexpr int workspace { 1000 }; // numerical_integration.h:40
        double lo[workspace + 1];
        double hi[workspace + 1];
        double sums[workspace + 1];
        double errs[workspace + 1];
        lo[0] = lower;
        hi[0] = upper;
        sums[0] = gaussKronrod(fp, lower, upper, errs[0]);
        for (int iter = 0; iter < workspace; ++iter) {
            int worst { 0 };
            for (int i = 1; i <= iter; ++i) {
                if (errs[worst] < errs[i]) {
                    worst = i;
                }
            }
            const double a { lo[worst] };
            const double b { hi[worst] };
            const double middle { (b + a) / 2 };
            sums[worst] = gaussKronrod(fp, a, middle, errs[worst]);
            sums[iter + 1] = gaussKronrod(fp, middle, b, errs[iter + 1]);
            hi[worst] = middle;
            lo[iter + 1] = middle;
            hi[iter + 1] = b;
            double errors_sum {};
            double sums_sum {};
            for (int i = 0; i < iter + 2; ++i) {
                errors_sum += errs[i];
                sums_sum += sums[i];
            }
            if (errors_sum < abs_error || errors_sum / sums_sum < rel_error) {
                FVar result { passive(0.0) };
                for (int i = 0; i < iter + 2; ++i) {
                    result = result
                             + gadfit_oracle::kronrod(
                               lo[i], hi[i], [&](const FVar& t) { return f(*this, t); });
                }
                return result;
            }
        }
        return passive(std::nan(""));
    }
    template <class F>
    auto bound_term(F&& f, const FVar& bound, const double sign, FVar& result) const -> void
    {
        if (!bound.on) {
            return;
        }
        const FVar tmp { f(*this, passive(bound.val)) };
        result.d += sign * bound.d * tmp.val;
        result.dd += sign * (bound.dd * tmp.val + bound.d * (f(*this, bound).d + tmp.d));
        result.on = true;
    }
    template <class F>
    auto integrate(const FVar& lower, const double upper, F&& f, const double rel = 2.220446049250313e-14,
                   const double abs = 0.0) const -> FVar
    {
        FVar result { integrate(lower.val, upper, f, rel, abs) };
        bound_term(f, lower, -1.0, result);
        return result;
    }
    template <class F>
    auto integrate(const double lower, const FVar& upper, F&& f, const double rel = 2.220446049250313e-14,
                   const double abs = 0.0) const -> FVar
    {
        FVar result { integrate(lower, upper.val, f, rel, abs) };
        bound_term(f, upper, 1.0, result);
        return result;
    }
    template <class F>
    auto integrate(const FVar& lower, const FVar& upper, F&& f, const double rel = 2.220446049250313e-14,
                   const double abs = 0.0) const -> FVar
    {
        FVar result { integrate(lower.val, upper.val, f, rel, abs) };
        bound_term(f, lower, -1.0, result);
        bound_term(f, upper, 1.0, result);
        return result;
    }
    // Direct 2-D adaptive integration, numerical_integration.cpp:426-526, with
    // the forward-mode terms of traceRecordY1/Y2/X1/X2 (:533-811) for AD-valued
    // bounds.  f(ctx, x, y); y1, y2 bound the outer variable, x1, x2 the inner.
    static auto as_var(const double b) -> FVar { return passive(b); }
    static auto as_var(const FVar& b) -> FVar { return b; }
    // G7xG7 / K15xK15 on one rectangle, :150-187
    template <class F>
    static auto gaussKronrod2d(F&& fp,
                               const double y1,
                               const double y2,
                               const double x1,
                               const double x2,
                               double& abs_error) -> double
    {
        const double outer_scale { (y2 - y1) / 2 };
        const double outer_shift { (y1 + y2) / 2 };
        const double inner_scale { (x2 - x1) / 2 };
        const double inner_shift { (x1 + x2) / 2 };
        double sum_kronrod {};
        double sum_gauss {};
        for (int i = 0; i < 15; ++i) {
            const FVar arg_x { passive(inner_scale * k15_roots[i] + inner_shift) };
            for (int j = 0; j < 15; ++j) {
                const FVar arg_y { passive(outer_scale * k15_roots[j] + outer_shift) };
                const double val { fp(arg_x, arg_y).val };
                sum_kronrod += (k15_weights[i] * k15_weights[j]) * val;
                if (i % 2 == 1 && j % 2 == 1) {
                    sum_gauss += (g7_weights[i / 2] * g7_weights[j / 2]) * val;
                }
            }
        }
        sum_kronrod *= outer_scale * inner_scale;
        abs_error = std::abs(sum_kronrod - outer_scale * inner_scale * sum_gauss);
        return sum_kronrod;
    }
    // One moving edge.  y_edge: `bound` is y1 or y2 and the edge integral runs
    // over x in [o1, o2]; otherwise `bound` is x1 or x2 and it runs over y.
    template <class F>
    auto edge_term(F&& f,
                   const bool y_edge,
                   const FVar& bound,
                   const FVar& o1,
                   const FVar& o2,
                   const double sign,
                   const double rel,
                   const double abs,
                   FVar& result) const -> void
    {
        if (!bound.on) {
            return;
        }
        FVar inactive[64];
        for (int i = 0; i < n_pars; ++i) {
            inactive[i] = passive(pars[i].val);
        }
        const Ctx pc { inactive, n_pars };
        const auto along { [&](const Ctx& cc, const FVar& t, const FVar& edge) {
            return y_edge ? f(cc, t, edge) : f(cc, edge, t);
        } };
        const FVar b_passive { passive(bound.val) };
        const auto fixed { [&](const Ctx& cc, const FVar& t) { return along(cc, t, b_passive); } };
        const auto fixed_active { [&](const Ctx& cc, const FVar& t) { return along(cc, t, bound); } };
        const double edge_integral { pc.integrate(o1.val, o2.val, fixed, rel, abs).val };
        const double f_hi { along(pc, passive(o2.val), b_passive).val };
        const double f_lo { along(pc, passive(o1.val), b_passive).val };
        const FVar f_chi { integrate(o1.val, o2.val, fixed, rel, abs) };
        const FVar f_full { integrate(o1.val, o2.val, fixed_active, rel, abs) };
        result.d += sign * bound.d * edge_integral;
        result.dd += sign
                     * (bound.dd * edge_integral
                        + bound.d * (o2.d * f_hi - o1.d * f_lo + f_chi.d + f_full.d));
        result.on = true;
    }
    template <class BY1, class BY2, class BX1, class BX2, class F>
    auto integrate2d(const BY1& y1_bound,
                     const BY2& y2_bound,
                     const BX1& x1_bound,
                     const BX2& x2_bound,
                     F&& f,
                     const double rel_error = 2.220446049250313e-14,
                     const double abs_error = 0.0) const -> FVar
    {
        const FVar by1 { as_var(y1_bound) };
        const FVar by2 { as_var(y2_bound) };
        const FVar bx1 { as_var(x1_bound) };
        const FVar bx2 { as_var(x2_bound) };
        FVar inactive[64];
        for (int i = 0; i < n_pars; ++i) {
            inactive[i] = passive(pars[i].val);
        }
        const Ctx pc { inactive, n_pars };
        const auto fp { [&](const FVar& x, const FVar& y) { return f(pc, x, y); } };
        struct Rect
        {
            double y1, y2, x1, x2;
        };
        constexpr int workspace { 1000 };
        std::vector<Rect> rect(workspace + 1);
        std::vector<double> sums(workspace + 1);
        std::vector<double> errs(workspace + 1);
        const double y_scale_inv { 1.0 / (by2.val - by1.val) };
        const double x_scale_inv { 1.0 / (bx2.val - bx1.val) };
        rect[0] = { by1.val, by2.val, bx1.val, bx2.val };
        sums[0] = gaussKronrod2d(fp, by1.val, by2.val, bx1.val, bx2.val, errs[0]);
        FVar result { passive(std::nan("")) };
        for (int iter = 0; iter < workspace; ++iter) {
            const int next { iter + 1 };
            int worst { 0 };
            for (int i = 1; i <= iter; ++i) {
                if (errs[worst] < errs[i]) {
                    worst = i;
                }
            }
            const Rect cur { rect[worst] };
            if (y_scale_inv * (cur.y2 - cur.y1) > x_scale_inv * (cur.x2 - cur.x1)) {
                const double middle { (cur.y1 + cur.y2) / 2 };
                sums[worst] = gaussKronrod2d(fp, cur.y1, middle, cur.x1, cur.x2, errs[worst]);
                sums[next] = gaussKronrod2d(fp, middle, cur.y2, cur.x1, cur.x2, errs[next]);
                rect[worst].y2 = middle;
                rect[next] = { middle, cur.y2, cur.x1, cur.x2 };
            } else {
                const double middle { (cur.x1 + cur.x2) / 2 };
                sums[worst] = gaussKronrod2d(fp, cur.y1, cur.y2, cur.x1, middle, errs[worst]);
                sums[next] = gaussKronrod2d(fp, cur.y1, cur.y2, middle, cur.x2, errs[next]);
                rect[worst].x2 = middle;
                rect[next] = { cur.y1, cur.y2, middle, cur.x2 };
            }
            double errors_sum {};
            double sums_sum {};
            for (int i = 0; i < iter + 2; ++i) {
                errors_sum += errs[i];
                sums_sum += sums[i];
            }
            if (errors_sum < abs_error || errors_sum / sums_sum < rel_error) {
                result = passive(0.0);
                for (int i = 0; i < iter + 2; ++i) {
                    result = result
                             + gadfit_oracle::kronrod2d(
                               rect[i].y1,
                               rect[i].y2,
                               rect[i].x1,
                               rect[i].x2,
                               [&](const FVar& x, const FVar& y) { return f(*this, x, y); });
                }
                break;
            }
        }
        edge_term(f, true, by1, bx1, bx2, -1.0, rel_error, abs_error, result);
        edge_term(f, true, by2, bx1, bx2, 1.0, rel_error, abs_error, result);
        edge_term(f, false, bx1, by1, by2, -1.0, rel_error, abs_error, result);
        edge_term(f, false, bx2, by1, by2, 1.0, rel_error, abs_error, result);
        return result;
    }
    template <class F>
    auto kronrod(const double lower, const double upper, F&& f) const -> FVar
    {
        return gadfit_oracle::kronrod(lower, upper, f);
    }
    template <class F>
    auto kronrod2d(const double ol, const double ou, const double il, const double iu, F&& f) const
      -> FVar
    {
        return gadfit_oracle::kronrod2d(ol, ou, il, iu, f);
    }
    // the integrand also receives the node indices (fit_function.h)
    template <class F>
    auto kronrod_indexed(const double lower, const double upper, F&& f) const -> FVar
    {
        int i { 0 };
        return gadfit_oracle::kronrod(lower, upper, [&](const FVar& t) { return f(t, i++); });
    }
    template <class F>
    auto kronrod2d_indexed(const double ol, const double ou, const double il, const double iu,
                           F&& f) const -> FVar
    {
        int n { 0 }; // numerical_integration.cpp:204-212: inner node i outer loop, outer node j inner loop
        return gadfit_oracle::kronrod2d(ol, ou, il, iu, [&](const FVar& u, const FVar& v) {
            const int i { n / 15 };
            const int j { n % 15 };
            ++n;
            return f(u, v, i, j);
        });
    }
};

} // namespace gadfit_oracle
