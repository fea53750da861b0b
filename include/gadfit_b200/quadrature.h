// Fixed-order 15-point Kronrod quadrature in registers, 1-D and 2-D.
//
// Counterpart of gadfit::kronrod() of the reference: one 15-node panel,
// AD-aware (c++/gadfit/numerical_integration.cpp:133-148 for 1-D, :189-215
// for 2-D).  Node order, weights and the order of the floating point
// operations follow the reference: nodes are visited as 7 negative, 0,
// 7 positive (gauss_kronrod_parameters.h:20-31); the sum starts from a
// passive zero, so its first term is exactly w0 * f(t0); the result is
// scale * sum (1-D) and (outer_scale * inner_scale) * sum (2-D).
//
// The integrand is any callable on a passive node (and whatever parameters
// it captured); its return type carries the AD mode (SDual<M> or Dir2).
#pragma once

#include "dual.h"

namespace gadfit_b200 {

constexpr int k15_size { 15 };

// Nodes and weights of the 15-point Kronrod rule (the numerical values are
// those of the rule itself; layout as in gauss_kronrod_parameters.h:20-47)
#define GADFIT_K15_ROOTS                                                       \
    -0.991455371120812639206854697526329, -0.949107912342758524526189684047851, \
      -0.864864423359769072789712788640926,                                    \
      -0.741531185599394439863864773280788,                                    \
      -0.586087235467691130294144838258730,                                    \
      -0.405845151377397166906606412076961,                                    \
      -0.207784955007898467600689403773245,                                    \
      0.000000000000000000000000000000000,                                     \
      0.991455371120812639206854697526329,                                     \
      0.949107912342758524526189684047851,                                     \
      0.864864423359769072789712788640926,                                     \
      0.741531185599394439863864773280788,                                     \
      0.586087235467691130294144838258730,                                     \
      0.405845151377397166906606412076961,                                     \
      0.207784955007898467600689403773245
#define GADFIT_K15_WEIGHTS                                                     \
    0.022935322010529224963732008058970, 0.063092092629978553290700663189204,  \
      0.104790010322250183839876322541518,                                     \
      0.140653259715525918745189590510238,                                     \
      0.169004726639267902826583426598550,                                     \
      0.190350578064785409913256402421014,                                     \
      0.204432940075298892414161999234649,                                     \
      0.209482141084727828012999174891714,                                     \
      0.022935322010529224963732008058970,                                     \
      0.063092092629978553290700663189204,                                     \
      0.104790010322250183839876322541518,                                     \
      0.140653259715525918745189590510238,                                     \
      0.169004726639267902826583426598550,                                     \
      0.190350578064785409913256402421014,                                     \
      0.204432940075298892414161999234649

#if defined(__CUDACC__)
// Warp-uniform index -> constant-bank load (LDC), no local memory
static __constant__ double k15_roots_dev[k15_size] { GADFIT_K15_ROOTS };
static __constant__ double k15_weights_dev[k15_size] { GADFIT_K15_WEIGHTS };
#endif
static constexpr double k15_roots_host[k15_size] { GADFIT_K15_ROOTS };
static constexpr double k15_weights_host[k15_size] { GADFIT_K15_WEIGHTS };

GADFIT_HD double k15_root(const int i)
{
#if defined(__CUDA_ARCH__)
    return k15_roots_dev[i];
#else
    return k15_roots_host[i];
#endif
}

GADFIT_HD double k15_weight(const int i)
{
#if defined(__CUDA_ARCH__)
    return k15_weights_dev[i];
#else
    return k15_weights_host[i];
#endif
}

// numerical_integration.cpp:189-215.  f(arg_x, arg_y): arg_x runs over the
// inner bounds in the OUTER loop, arg_y over the outer bounds in the inner
// loop, exactly as the reference does.
template <class F>
GADFIT_HD auto kronrod2d(const double outer_lower,
                         const double outer_upper,
                         const double inner_lower,
                         const double inner_upper,
                         F&& f)
{
    const double outer_scale { (outer_upper - outer_lower) / 2 };
    const double outer_shift { (outer_lower + outer_upper) / 2 };
    const double inner_scale { (inner_upper - inner_lower) / 2 };
    const double inner_shift { (inner_lower + inner_upper) / 2 };
    auto sum { (k15_weight(0) * k15_weight(0))
               * f(Passive { inner_scale * k15_root(0) + inner_shift },
                   Passive { outer_scale * k15_root(0) + outer_shift }) };
#pragma unroll 1
    for (int ij = 1; ij < k15_size * k15_size; ++ij) {
        const int i { ij / k15_size };
        const int j { ij - i * k15_size };
        sum = sum
              + (k15_weight(i) * k15_weight(j))
                  * f(Passive { inner_scale * k15_root(i) + inner_shift },
                      Passive { outer_scale * k15_root(j) + outer_shift });
    }
    return outer_scale * inner_scale * sum;
}

// Node i of the 15-point panel over [lower, upper], computed exactly as the
// fixed rules (and the reference, numerical_integration.cpp:139-144) do: what
// a fitting function's prepare() tabulates per node matches the node the
// integrand is later called with.
GADFIT_HD double kronrod_node(const int i, const double lower, const double upper)
{
    const double scale { (upper - lower) / 2 };
    const double shift { (lower + upper) / 2 };
    return scale * k15_root(i) + shift;
}

// The 2-D rule with the node indices handed to the integrand: f(u, v, i, j),
// u = inner node i, v = outer node j (same nodes, weights and summation order
// as kronrod2d above)
template <class F>
GADFIT_HD auto kronrod2d_indexed(const double outer_lower,
                                 const double outer_upper,
                                 const double inner_lower,
                                 const double inner_upper,
                                 F&& f)
{
    const double outer_scale { (outer_upper - outer_lower) / 2 };
    const double outer_shift { (outer_lower + outer_upper) / 2 };
    const double inner_scale { (inner_upper - inner_lower) / 2 };
    const double inner_shift { (inner_lower + inner_upper) / 2 };
    // Partially unrolled: the node evaluations of one group are independent
    // (only the additions into `sum` are ordered), which is what lets a
    // thread overlap their dependent FP64 chains (reciprocal, exp)
    const Passive u0 { inner_scale * k15_root(0) + inner_shift };
    const double w0 { k15_weight(0) };
    auto sum { (w0 * w0) * f(u0, Passive { outer_scale * k15_root(0) + outer_shift }, 0, 0) };
#pragma unroll 7
    for (int j = 1; j < k15_size; ++j) {
        sum = sum
              + (w0 * k15_weight(j))
                  * f(u0, Passive { outer_scale * k15_root(j) + outer_shift }, 0, j);
    }
#pragma unroll 1
    for (int i = 1; i < k15_size; ++i) {
        const Passive u { inner_scale * k15_root(i) + inner_shift };
        const double wi { k15_weight(i) };
#pragma unroll 5
        for (int j = 0; j < k15_size; ++j) {
            sum = sum
                  + (wi * k15_weight(j))
                      * f(u, Passive { outer_scale * k15_root(j) + outer_shift }, i, j);
        }
    }
    return outer_scale * inner_scale * sum;
}


// ===========================================================================
// Adaptive Gauss-Kronrod (G7/K15) integration, 1-D
// ===========================================================================
//
// Counterpart of gadfit::integrate() (reference
// c++/gadfit/numerical_integration.cpp:242-306): G7/K15 on [a, b], then
// repeatedly bisect the panel with the largest error estimate (always at
// least once) until  sum(err) < abs_error  or  sum(err)/sum(K15) < rel_error,
// with PASSIVE parameters; finally every panel is re-evaluated with AD on by
// the fixed rule above, in panel order (:292-298).  AD-valued bounds add the
// Leibniz terms of traceRecordLower / traceRecordUpper (:313-376).
//
// The integrand is a callable f(ctx, t): it receives the evaluation context
// (passive during refinement, the caller's during the final pass) exactly as
// the reference hands its integrand either `inactive_parameters` or
// `parameters`.
//
// The per-thread panel workspace is a fixed array (the reference's default
// is 1000 panels on the heap, numerical_integration.h:40); overflow returns
// NaN -- which the LM loop treats as a rejected step -- instead of throwing
// InsufficientIntegrationWorkspace.

constexpr int k_max_panels { 96 };

// Sets bit 0 of the context's status word, if it carries one (the device
// contexts of fit_function.h do; host-side test contexts need not)
template <class Ctx>
GADFIT_HD void report_workspace_exhausted(const Ctx& c)
{
    if constexpr (requires { c.status; }) {
        if (c.status != nullptr) {
#if defined(__CUDA_ARCH__)
            atomicOr(c.status, 1U);
#else
            *c.status |= 1U;
#endif
        }
    }
}
constexpr double k_default_rel_error { 2.220446049250313e-14 }; // 1e2 * epsilon, numerical_integration.h:37-38

#define GADFIT_G7_WEIGHTS                                                      \
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,  \
      0.381830050505118944950369775488975,                                     \
      0.417959183673469387755102040816327,                                     \
      0.129484966168869693270611432679082,                                     \
      0.279705391489276667901467771423780,                                     \
      0.381830050505118944950369775488975
#if defined(__CUDACC__)
static __constant__ double g7_weights_dev[7] { GADFIT_G7_WEIGHTS };
#endif
static constexpr double g7_weights_host[7] { GADFIT_G7_WEIGHTS };

GADFIT_HD double g7_weight(const int i)
{
#if defined(__CUDA_ARCH__)
    return g7_weights_dev[i];
#else
    return g7_weights_host[i];
#endif
}

// ---------------------------------------------------------------------------
// Rule traits.  KronrodRule<15> is the rule of the C++ reference; the Fortran
// version also offers 21, 31, 41, 51 and 61 points (set_integration_rule,
// fortran/gadfit/numerical_integration.F90:139-171) -- quadrature_rules.h.
// A rule is a template argument here (`c.template integrate<31>(...)`,
// `kronrod<21>(...)`): the node loop bound is a compile-time constant and the
// tables sit in the constant bank.  Every rule lists its nodes in the order of
// the reference's tables; the Gauss nodes are the odd positions.
// ---------------------------------------------------------------------------
template <int PTS>
struct KronrodRule;

template <>
struct KronrodRule<15>
{
    static constexpr int size { k15_size };
    GADFIT_HD static double root(const int i) { return k15_root(i); }
    GADFIT_HD static double weight(const int i) { return k15_weight(i); }
    GADFIT_HD static double gauss_weight(const int i) { return g7_weight(i); }
};

#define GADFIT_LIST(...) __VA_ARGS__
#if defined(__CUDACC__)
#define GADFIT_RULE_DEVICE_TABLE(name, n, ...) static __constant__ double name##_dev[n] { __VA_ARGS__ };
#else
#define GADFIT_RULE_DEVICE_TABLE(name, n, ...)
#endif
#if defined(__CUDA_ARCH__)
#define GADFIT_RULE_ENTRY(name, i) name##_dev[i]
#else
#define GADFIT_RULE_ENTRY(name, i) name##_host[i]
#endif
#define GADFIT_DEFINE_KRONROD_RULE(PTS, NG, ROOTS, KRONROD, GAUSS)                          \
    GADFIT_RULE_DEVICE_TABLE(k##PTS##_roots, PTS, ROOTS)                                    \
    GADFIT_RULE_DEVICE_TABLE(k##PTS##_weights, PTS, KRONROD)                                \
    GADFIT_RULE_DEVICE_TABLE(k##PTS##_gauss, NG, GAUSS)                                     \
    static constexpr double k##PTS##_roots_host[PTS] { ROOTS };                             \
    static constexpr double k##PTS##_weights_host[PTS] { KRONROD };                         \
    static constexpr double k##PTS##_gauss_host[NG] { GAUSS };                              \
    template <>                                                                             \
    struct KronrodRule<PTS>                                                                 \
    {                                                                                       \
        static constexpr int size { PTS };                                                  \
        GADFIT_HD static double root(const int i) { return GADFIT_RULE_ENTRY(k##PTS##_roots, i); }      \
        GADFIT_HD static double weight(const int i) { return GADFIT_RULE_ENTRY(k##PTS##_weights, i); }  \
        GADFIT_HD static double gauss_weight(const int i) { return GADFIT_RULE_ENTRY(k##PTS##_gauss, i); } \
    };

// numerical_integration.cpp:133-148 (any rule)
template <int PTS = 15, class F>
GADFIT_HD auto kronrod(const double lower, const double upper, F&& f)
{
    using Rule = KronrodRule<PTS>;
    const double scale { (upper - lower) / 2 };
    const double shift { (lower + upper) / 2 };
    auto sum { Rule::weight(0) * f(Passive { scale * Rule::root(0) + shift }) };
#pragma unroll 1
    for (int i = 1; i < Rule::size; ++i) {
        sum = sum + Rule::weight(i) * f(Passive { scale * Rule::root(i) + shift });
    }
    return scale * sum;
}

// The same with the node index handed to the integrand: f(t, i)
template <int PTS = 15, class F>
GADFIT_HD auto kronrod_indexed(const double lower, const double upper, F&& f)
{
    using Rule = KronrodRule<PTS>;
    const double scale { (upper - lower) / 2 };
    const double shift { (lower + upper) / 2 };
    auto sum { Rule::weight(0) * f(Passive { scale * Rule::root(0) + shift }, 0) };
    // partially unrolled: independent node evaluations overlap their FP64 chains
#pragma unroll 7
    for (int i = 1; i < Rule::size; ++i) {
        sum = sum + Rule::weight(i) * f(Passive { scale * Rule::root(i) + shift }, i);
    }
    return scale * sum;
}

// numerical_integration.cpp:108-131: K15 value and |K15 - G7| of one panel
template <int PTS = 15, class Ctx, class F>
GADFIT_HD double gauss_kronrod(const Ctx& passive_ctx,
                               F&& f,
                               const double lower,
                               const double upper,
                               double& abs_error)
{
    using Rule = KronrodRule<PTS>;
    const double scale { (upper - lower) / 2 };
    const double shift { (lower + upper) / 2 };
    double sum_kronrod { 0.0 };
    double sum_gauss { 0.0 };
#pragma unroll 1
    for (int i = 0; i < Rule::size; ++i) {
        const double val { value_of(f(passive_ctx, Passive { scale * Rule::root(i) + shift })) };
        sum_kronrod += Rule::weight(i) * val;
        if (i % 2 == 1) {
            sum_gauss += Rule::gauss_weight(i / 2) * val;
        }
    }
    sum_kronrod *= scale;
    abs_error = ::fabs(sum_kronrod - scale * sum_gauss);
    return sum_kronrod;
}

GADFIT_HD double d_of(const double) { return 0.0; }
GADFIT_HD double d_of(const Passive&) { return 0.0; }
GADFIT_HD double d_of(const Dir2& x) { return x.d; }
GADFIT_HD Dir2 as_dir2(const Passive& x) { return Dir2 { x.val, 0.0, 0.0 }; }
GADFIT_HD Dir2 as_dir2(const Dir2& x) { return x; }

template <class T>
struct is_passive_bound : std::bool_constant<std::is_arithmetic_v<T> || std::is_same_v<T, Passive>>
{};

GADFIT_HD double bound_value(const double x) { return x; }
template <uint32_t M>
GADFIT_HD double bound_value(const SDual<M>& x) { return x.val; }
GADFIT_HD double bound_value(const Dir2& x) { return x.val; }

// Leibniz term of one AD-valued bound; sign = +1 upper, -1 lower
template <class Ctx, class F, class R, class B>
GADFIT_HD auto leibniz(const Ctx& c, F&& f, const R& result, const B& bound, const double sign)
{
    if constexpr (is_passive_bound<B>::value) {
        return result;
    } else if constexpr (std::is_same_v<B, Dir2>) {
        // numerical_integration.cpp:336-343 / :369-376
        const auto tmp { f(c, Passive { bound.val }) };
        Dir2 r { as_dir2(result) };
        r.d += sign * bound.d * value_of(tmp);
        r.dd += sign * (bound.dd * value_of(tmp) + bound.d * (d_of(f(c, bound)) + d_of(tmp)));
        return r;
    } else {
        // first derivatives only: d/dp int^{b(p)} f = b'(p) f(b)
        // (the adjoint rule of Op::integration_bound,
        // automatic_differentiation.cpp:867-871)
        const double fb { value_of(f(c.passive(), Passive { bound.val })) };
        return result + (bound - bound.val) * (sign * fb);
    }
}

template <int PTS = 15, class Ctx, class F, class BL, class BU>
GADFIT_HD auto integrate(const Ctx& c,
                         const BL& lower_bound,
                         const BU& upper_bound,
                         F&& f,
                         const double rel_error = k_default_rel_error,
                         const double abs_error = 0.0)
{
    const double lower { bound_value(lower_bound) };
    const double upper { bound_value(upper_bound) };
    const auto pc { c.passive() };
    double lo[k_max_panels];
    double hi[k_max_panels];
    double sums[k_max_panels];
    double errs[k_max_panels];
    lo[0] = lower;
    hi[0] = upper;
    sums[0] = gauss_kronrod<PTS>(pc, f, lower, upper, errs[0]);
    int n_panels { 0 };
    for (int iter = 0; iter + 1 < k_max_panels; ++iter) {
        // first panel with the largest error (std::max_element)
        int worst { 0 };
        for (int i = 1; i <= iter; ++i) {
            if (errs[worst] < errs[i]) {
                worst = i;
            }
        }
        const double a { lo[worst] };
        const double b { hi[worst] };
        const double middle { (b + a) / 2 };
        sums[worst] = gauss_kronrod<PTS>(pc, f, a, middle, errs[worst]);
        sums[iter + 1] = gauss_kronrod<PTS>(pc, f, middle, b, errs[iter + 1]);
        hi[worst] = middle;
        lo[iter + 1] = middle;
        hi[iter + 1] = b;
        double errors_sum { 0.0 };
        double sums_sum { 0.0 };
        for (int i = 0; i < iter + 2; ++i) {
            errors_sum += errs[i];
            sums_sum += sums[i];
        }
        if (errors_sum < abs_error || errors_sum / sums_sum < rel_error) {
            n_panels = iter + 2;
            break;
        }
    }
    // Final pass with AD on (:292-298); the sum starts from a passive zero
    auto result { Passive { 0.0 } + kronrod<PTS>(lo[0], hi[0], [&](const Passive& t) { return f(c, t); }) };
    for (int i = 1; i < n_panels; ++i) {
        result = result + kronrod<PTS>(lo[i], hi[i], [&](const Passive& t) { return f(c, t); });
    }
    if (n_panels == 0) {
        // workspace exhausted: the status word of the context is set -- fit()
        // then returns GADFIT_B200_INSUFFICIENT_INTEGRATION_WORKSPACE, the
        // counterpart of the reference's exception -- and the value is
        // poisoned so that no step computed from it can be accepted
        report_workspace_exhausted(c);
        result = result * (0.0 / 0.0);
    }
    return leibniz(c, f, leibniz(c, f, result, lower_bound, -1.0), upper_bound, 1.0);
}

// ===========================================================================
// Adaptive Gauss-Kronrod integration, direct 2-D
// ===========================================================================
//
// Counterpart of the 2-D gadfit::integrate() (reference
// numerical_integration.cpp:426-526): 15 x 15 Kronrod / 7 x 7 Gauss on a
// rectangle, bisect the rectangle with the largest error estimate along its
// (relatively) longer side, final pass with AD on; AD-valued bounds add the
// Leibniz terms of traceRecordY1/Y2/X1/X2 (:533-811), which are 1-D adaptive
// integrals along the moving edge.  The integrand is f(ctx, x, y); y is the
// outer variable (bounds y1, y2), x the inner one (bounds x1, x2).

constexpr int k_max_panels_2d { 64 };

// numerical_integration.cpp:150-187
template <class Ctx, class F>
GADFIT_HD double gauss_kronrod2d(const Ctx& passive_ctx,
                                 F&& f,
                                 const double y1,
                                 const double y2,
                                 const double x1,
                                 const double x2,
                                 double& abs_error)
{
    const double outer_scale { (y2 - y1) / 2 };
    const double outer_shift { (y1 + y2) / 2 };
    const double inner_scale { (x2 - x1) / 2 };
    const double inner_shift { (x1 + x2) / 2 };
    double sum_kronrod { 0.0 };
    double sum_gauss { 0.0 };
#pragma unroll 1
    for (int i = 0; i < k15_size; ++i) {
        const double arg_x { inner_scale * k15_root(i) + inner_shift };
#pragma unroll 1
        for (int j = 0; j < k15_size; ++j) {
            const double arg_y { outer_scale * k15_root(j) + outer_shift };
            const double val { value_of(f(passive_ctx, Passive { arg_x }, Passive { arg_y })) };
            sum_kronrod += (k15_weight(i) * k15_weight(j)) * val;
            if (i % 2 == 1 && j % 2 == 1) {
                sum_gauss += (g7_weight(i / 2) * g7_weight(j / 2)) * val;
            }
        }
    }
    sum_kronrod *= outer_scale * inner_scale;
    abs_error = ::fabs(sum_kronrod - outer_scale * inner_scale * sum_gauss);
    return sum_kronrod;
}

// Leibniz term of one AD-valued bound of the rectangle.  AXIS 0: `bound` is a
// y bound (the edge integral runs over x in [o1, o2]); AXIS 1: an x bound (the
// edge integral runs over y).  sign = +1 upper, -1 lower.
template <int AXIS, class Ctx, class F, class R, class B, class O1, class O2>
GADFIT_HD auto leibniz2d(const Ctx& c,
                         F&& f,
                         const R& result,
                         const B& bound,
                         const O1& o1,
                         const O2& o2,
                         const double sign,
                         const double rel_error,
                         const double abs_error)
{
    if constexpr (is_passive_bound<B>::value) {
        return result;
    } else {
        const double lo { bound_value(o1) };
        const double hi { bound_value(o2) };
        // f along the edge; `edge` is the fixed coordinate (passive or active)
        const auto along { [&f](const auto& cc, const auto& t, const auto& edge) {
            if constexpr (AXIS == 0) {
                return f(cc, t, edge);
            } else {
                return f(cc, edge, t);
            }
        } };
        const Passive b_passive { bound.val };
        const auto pc { c.passive() };
        const double edge_integral { value_of(integrate(
          pc, lo, hi, [&](const auto& cc, const auto& t) { return along(cc, t, b_passive); },
          rel_error, abs_error)) };
        if constexpr (std::is_same_v<B, Dir2>) {
            // numerical_integration.cpp:561-590 (and the three siblings)
            const double f_hi { value_of(along(pc, Passive { hi }, b_passive)) };
            const double f_lo { value_of(along(pc, Passive { lo }, b_passive)) };
            const auto f_chi { integrate(
              c, lo, hi, [&](const auto& cc, const auto& t) { return along(cc, t, b_passive); },
              rel_error, abs_error) };
            const auto f_full { integrate(
              c, lo, hi, [&](const auto& cc, const auto& t) { return along(cc, t, bound); },
              rel_error, abs_error) };
            Dir2 r { as_dir2(result) };
            r.d += sign * bound.d * edge_integral;
            r.dd += sign
                    * (bound.dd * edge_integral
                       + bound.d * (d_of(o2) * f_hi - d_of(o1) * f_lo + d_of(f_chi) + d_of(f_full)));
            return r;
        } else {
            return result + (bound - bound.val) * (sign * edge_integral);
        }
    }
}

template <class Ctx, class F, class BY1, class BY2, class BX1, class BX2>
GADFIT_HD auto integrate2d(const Ctx& c,
                           const BY1& y1_bound,
                           const BY2& y2_bound,
                           const BX1& x1_bound,
                           const BX2& x2_bound,
                           F&& f,
                           const double rel_error = k_default_rel_error,
                           const double abs_error = 0.0)
{
    const double y1 { bound_value(y1_bound) };
    const double y2 { bound_value(y2_bound) };
    const double x1 { bound_value(x1_bound) };
    const double x2 { bound_value(x2_bound) };
    const auto pc { c.passive() };
    const double y_scale_inv { 1.0 / (y2 - y1) };
    const double x_scale_inv { 1.0 / (x2 - x1) };
    double by1[k_max_panels_2d];
    double by2[k_max_panels_2d];
    double bx1[k_max_panels_2d];
    double bx2[k_max_panels_2d];
    double sums[k_max_panels_2d];
    double errs[k_max_panels_2d];
    by1[0] = y1;
    by2[0] = y2;
    bx1[0] = x1;
    bx2[0] = x2;
    sums[0] = gauss_kronrod2d(pc, f, y1, y2, x1, x2, errs[0]);
    int n_panels { 0 };
    for (int iter = 0; iter + 1 < k_max_panels_2d; ++iter) {
        const int next { iter + 1 };
        int worst { 0 };
        for (int i = 1; i <= iter; ++i) {
            if (errs[worst] < errs[i]) {
                worst = i;
            }
        }
        const double cy1 { by1[worst] };
        const double cy2 { by2[worst] };
        const double cx1 { bx1[worst] };
        const double cx2 { bx2[worst] };
        if (y_scale_inv * (cy2 - cy1) > x_scale_inv * (cx2 - cx1)) {
            const double middle { (cy1 + cy2) / 2 };
            sums[worst] = gauss_kronrod2d(pc, f, cy1, middle, cx1, cx2, errs[worst]);
            sums[next] = gauss_kronrod2d(pc, f, middle, cy2, cx1, cx2, errs[next]);
            by2[worst] = middle;
            by1[next] = middle;
            by2[next] = cy2;
            bx1[next] = cx1;
            bx2[next] = cx2;
        } else {
            const double middle { (cx1 + cx2) / 2 };
            sums[worst] = gauss_kronrod2d(pc, f, cy1, cy2, cx1, middle, errs[worst]);
            sums[next] = gauss_kronrod2d(pc, f, cy1, cy2, middle, cx2, errs[next]);
            bx2[worst] = middle;
            by1[next] = cy1;
            by2[next] = cy2;
            bx1[next] = middle;
            bx2[next] = cx2;
        }
        double errors_sum { 0.0 };
        double sums_sum { 0.0 };
        for (int i = 0; i < iter + 2; ++i) {
            errors_sum += errs[i];
            sums_sum += sums[i];
        }
        if (errors_sum < abs_error || errors_sum / sums_sum < rel_error) {
            n_panels = iter + 2;
            break;
        }
    }
    const auto active { [&](const Passive& x, const Passive& y) { return f(c, x, y); } };
    auto result { Passive { 0.0 } + kronrod2d(by1[0], by2[0], bx1[0], bx2[0], active) };
    for (int i = 1; i < n_panels; ++i) {
        result = result + kronrod2d(by1[i], by2[i], bx1[i], bx2[i], active);
    }
    if (n_panels == 0) {
        report_workspace_exhausted(c); // see integrate()
        result = result * (0.0 / 0.0);
    }
    // traceRecordY1, Y2, X1, X2 in the reference's order (:1131-1144)
    auto r1 { leibniz2d<0>(c, f, result, y1_bound, x1_bound, x2_bound, -1.0, rel_error, abs_error) };
    auto r2 { leibniz2d<0>(c, f, r1, y2_bound, x1_bound, x2_bound, 1.0, rel_error, abs_error) };
    auto r3 { leibniz2d<1>(c, f, r2, x1_bound, y1_bound, y2_bound, -1.0, rel_error, abs_error) };
    return leibniz2d<1>(c, f, r3, x2_bound, y1_bound, y2_bound, 1.0, rel_error, abs_error);
}

} // namespace gadfit_b200

#include "quadrature_rules.h"
