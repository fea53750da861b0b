// Evaluation contexts: how a fitting function sees its parameters.
//
// Counterpart of gadfit::FitFunction (reference c++/gadfit/fit_function.h:
// 29-62, fit_function.cpp:17-61): there a std::function body receives a
// std::vector<AdVar> whose `idx` fields were flipped by activateParReverse /
// activateParForward / deactivatePar before each sweep.  A std::function
// cannot cross to the device, so a fitting function here is a functor with
//
//     static constexpr int n_pars;
//     template <class Ctx> GADFIT_HD auto operator()(const Ctx& c, double x) const;
//
// and the sweep being run is selected by the TYPE of the context:
//
//   ValueCtx  par<K>(c) is passive            -> chi2 sweep (lm_solver.cpp:513-529)
//   GradCtx   par<K>(c) is SDual<1 << K>      -> Jacobian sweep (lm_solver.cpp:286-322)
//   Dir2Ctx   par<K>(c) is Dir2 seeded with   -> omega sweep (lm_solver.cpp:362-380,
//             d = delta1[column], dd = 0         fit_function.cpp:45-51)
//
// PREPARE HOOK (optional).  Sub-expressions that depend on the parameters only
// -- 1/w, a table of exp(-k u_i^2) at the quadrature nodes -- need not be
// recomputed for every data point: a function may add
//
//     template <class Ctx> GADFIT_HD auto prepare(const Ctx& c) const;   // any aggregate of AD values
//     template <class Ctx, class Pre>
//     GADFIT_HD auto operator()(const Ctx& c, const Pre& pre, double x) const;
//
// The sweeps call prepare() once per DATA SET (parameters are per data set)
// and keep the result in shared memory; evaluate(model, ctx, x) below is the
// one entry point and, for contexts without a cache (the reference's AdVar in
// oracle/ref_harness.cpp, the oracle's dual), calls prepare() at every
// evaluation -- the same arithmetic, so the source of a function stays single.
//
// TUNING HINTS (optional static members of a fitting function; defaults are
// fine for any function, every hint below is an A/B measurement on a BASELINE
// shape, see csrc/sweep_kernels.cuh):
//     static constexpr int  chi2_unroll;            // points per thread and loop iteration of the chi2 sweep (1; 8 for cheap closed forms)
//     static constexpr int  chi2_blocks;            // resident CTAs per SM the chi2 sweep is compiled for (4; 2 from chi2_unroll 8)
//     static constexpr int  jacobian_blocks;        // the same for the software-pipelined Jacobian sweep (2 = 128 registers)
//     static constexpr bool jacobian_in_registers;  // few parameters, cheap function: M_s accumulated in registers, no tensor tiles
//     static constexpr int  omega_blocks;           // resident CTAs per SM of the omega sweep (no cap)
//
// c.kronrod(...) / c.kronrod2d(...) give the fixed 15-point Kronrod panel of
// the reference (numerical_integration.cpp:133-148,189-215); the *_indexed
// variants also hand the integrand the node indices, so that it can look up
// what prepare() tabulated at the nodes (kronrod_node());
// c.integrate(lower, upper, f(ctx, t), rel_error) its adaptive G7/K15 (or, as
// c.template integrate<PTS>(...), the Fortran version's G10/K21 ... G30/K61)
// integrator incl. AD-valued bounds (numerical_integration.cpp:242-423);
// c.integrate2d(y1, y2, x1, x2, f(ctx, x, y), rel_error) the direct 2-D
// adaptive integrator (numerical_integration.cpp:426-1146).
#pragma once

#include "dual.h"
#include "quadrature.h"

#include <utility>

namespace gadfit_b200 {

struct QuadratureMixin
{
    // PTS: 15 (default), 21, 31, 41, 51 or 61 points -- c.template kronrod<21>(...)
    template <int PTS = 15, class F>
    GADFIT_HD auto kronrod(const double lower, const double upper, F&& f) const
    {
        return gadfit_b200::kronrod<PTS>(lower, upper, f);
    }
    template <class F>
    GADFIT_HD auto kronrod2d(const double outer_lower,
                             const double outer_upper,
                             const double inner_lower,
                             const double inner_upper,
                             F&& f) const
    {
        return gadfit_b200::kronrod2d(
          outer_lower, outer_upper, inner_lower, inner_upper, f);
    }
    // f(t, i): i = index of the node t = kronrod_node(i, lower, upper)
    template <class F>
    GADFIT_HD auto kronrod_indexed(const double lower, const double upper, F&& f) const
    {
        return gadfit_b200::kronrod_indexed(lower, upper, f);
    }
    // f(u, v, i, j): u = inner node i, v = outer node j
    template <class F>
    GADFIT_HD auto kronrod2d_indexed(const double outer_lower,
                                     const double outer_upper,
                                     const double inner_lower,
                                     const double inner_upper,
                                     F&& f) const
    {
        return gadfit_b200::kronrod2d_indexed(
          outer_lower, outer_upper, inner_lower, inner_upper, f);
    }
};

// Bit 0 of the status word a context may carry: an adaptive integrator ran
// out of panels (the reference throws InsufficientIntegrationWorkspace,
// numerical_integration.cpp:304,525)
constexpr unsigned k_status_workspace { 1U };

struct ValueCtx : QuadratureMixin
{
    const double* p;
    unsigned* status; // sweep-wide status word (device memory) or nullptr
    GADFIT_HD explicit ValueCtx(const double* p, unsigned* status = nullptr)
      : p { p }, status { status }
    {}
    GADFIT_HD ValueCtx passive() const { return *this; }
    template <int PTS = 15, class BL, class BU, class F>
    GADFIT_HD auto integrate(const BL& lower,
                             const BU& upper,
                             F&& f,
                             const double rel_error = k_default_rel_error,
                             const double abs_error = 0.0) const
    {
        return gadfit_b200::integrate<PTS>(*this, lower, upper, f, rel_error, abs_error);
    }
    template <class BY1, class BY2, class BX1, class BX2, class F>
    GADFIT_HD auto integrate2d(const BY1& y1,
                               const BY2& y2,
                               const BX1& x1,
                               const BX2& x2,
                               F&& f,
                               const double rel_error = k_default_rel_error,
                               const double abs_error = 0.0) const
    {
        return gadfit_b200::integrate2d(*this, y1, y2, x1, x2, f, rel_error, abs_error);
    }
    template <int K>
    GADFIT_HD Passive get() const
    {
        return Passive { p[K] };
    }
};

struct GradCtx : QuadratureMixin
{
    const double* p;
    unsigned* status;
    GADFIT_HD explicit GradCtx(const double* p, unsigned* status = nullptr)
      : p { p }, status { status }
    {}
    GADFIT_HD ValueCtx passive() const { return ValueCtx { p, status }; }
    template <int PTS = 15, class BL, class BU, class F>
    GADFIT_HD auto integrate(const BL& lower,
                             const BU& upper,
                             F&& f,
                             const double rel_error = k_default_rel_error,
                             const double abs_error = 0.0) const
    {
        return gadfit_b200::integrate<PTS>(*this, lower, upper, f, rel_error, abs_error);
    }
    template <class BY1, class BY2, class BX1, class BX2, class F>
    GADFIT_HD auto integrate2d(const BY1& y1,
                               const BY2& y2,
                               const BX1& x1,
                               const BX2& x2,
                               F&& f,
                               const double rel_error = k_default_rel_error,
                               const double abs_error = 0.0) const
    {
        return gadfit_b200::integrate2d(*this, y1, y2, x1, x2, f, rel_error, abs_error);
    }
    template <int K>
    GADFIT_HD SDual<(1U << K)> get() const
    {
        SDual<(1U << K)> y;
        y.val = p[K];
        y.d[0] = 1.0;
        return y;
    }
};

struct Dir2Ctx : QuadratureMixin
{
    const double* p;
    // Direction: delta1 of the column this parameter maps to, 0 if passive
    const double* seed;
    unsigned* status;
    GADFIT_HD Dir2Ctx(const double* p, const double* seed, unsigned* status = nullptr)
      : p { p }, seed { seed }, status { status }
    {}
    GADFIT_HD ValueCtx passive() const { return ValueCtx { p, status }; }
    template <int PTS = 15, class BL, class BU, class F>
    GADFIT_HD auto integrate(const BL& lower,
                             const BU& upper,
                             F&& f,
                             const double rel_error = k_default_rel_error,
                             const double abs_error = 0.0) const
    {
        return gadfit_b200::integrate<PTS>(*this, lower, upper, f, rel_error, abs_error);
    }
    template <class BY1, class BY2, class BX1, class BX2, class F>
    GADFIT_HD auto integrate2d(const BY1& y1,
                               const BY2& y2,
                               const BX1& x1,
                               const BX2& x2,
                               F&& f,
                               const double rel_error = k_default_rel_error,
                               const double abs_error = 0.0) const
    {
        return gadfit_b200::integrate2d(*this, y1, y2, x1, x2, f, rel_error, abs_error);
    }
    template <int K>
    GADFIT_HD Dir2 get() const
    {
        return Dir2 { p[K], seed[K], 0.0 };
    }
};

// par<K>(ctx): parameter K in the AD type of the current sweep
template <int K, class Ctx>
GADFIT_HD auto par(const Ctx& c)
{
    return c.template get<K>();
}

// ---- prepare hook ----------------------------------------------------------
//
// Tables inside what prepare() returns: tabulate<N>(c, f) = { f(0), ..., f(N-1) }.
// On the device the 256 evaluation threads of a sweep's CTA run prepare()
// together, with the context wrapped in Cooperative<Ctx>: tabulate() then
// evaluates and returns only the entries i = lane (mod 256) of every table -- a
// LaneTable of ceil(N / 256) entries that lives in registers, so a table of
// exp() at 15 quadrature nodes, or at 15 x 15 node pairs, costs one exp() of
// latency -- and store_prepared() below scatters the lanes' parts into the
// full Table in shared memory.  Plain contexts (the reference's AdVar, the oracle's
// dual, a device context outside a sweep) fill the whole table.
constexpr int k_prepare_lanes { 256 };

template <class Ctx>
struct Cooperative : Ctx
{
    int lane;
    GADFIT_HD Cooperative(const Ctx& c, const int lane) : Ctx { c }, lane { lane } {}
};
template <class T>
struct is_cooperative : std::false_type
{};
template <class Ctx>
struct is_cooperative<Cooperative<Ctx>> : std::true_type
{};

template <class T, int N>
struct Table
{
    T v[N];
    GADFIT_HD const T& operator[](const int i) const { return v[i]; }
};
// One lane's entries of a Table<T, N>: v[k] = entry lane + 256 k
template <class T, int N>
struct LaneTable
{
    static constexpr int per_lane { (N + k_prepare_lanes - 1) / k_prepare_lanes };
    T v[per_lane];
};

template <int N, class Ctx, class F>
GADFIT_HD auto tabulate(const Ctx& c, F&& f)
{
    using T = decltype(f(0));
    if constexpr (is_cooperative<Ctx>::value) {
        LaneTable<T, N> t;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < LaneTable<T, N>::per_lane; ++k) {
            const int i { c.lane + k_prepare_lanes * k };
            t.v[k] = f(i < N ? i : N - 1); // (past the end: a valid entry, not stored)
        }
        return t;
    } else {
        Table<T, N> t;
        for (int i = 0; i < N; ++i) {
            t.v[i] = f(i);
        }
        return t;
    }
}

// dst <- the part of `mine` this lane is responsible for: its entries of every
// Table, everything else from lane 0 (every lane computed the same value)
template <class... Ts>
struct Pack;
template <class T>
struct is_pack : std::false_type
{};
template <class... Ts>
struct is_pack<Pack<Ts...>> : std::true_type
{};

template <class T>
    requires(!is_pack<T>::value)
GADFIT_HD void store_prepared(const T& mine, T& dst, const int lane)
{
    if (lane == 0) {
        dst = mine;
    }
}
template <class T, int N>
GADFIT_HD void store_prepared(const LaneTable<T, N>& mine, Table<T, N>& dst, const int lane)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < LaneTable<T, N>::per_lane; ++k) {
        const int i { lane + k_prepare_lanes * k };
        if (i < N) {
            dst.v[i] = mine.v[k];
        }
    }
}
// A heterogeneous aggregate for what prepare() returns: pack(a, b, c) keeps
// each value in its own AD type (SDual masks differ), get<I>() reads it back.
template <class... Ts>
struct Pack;
template <>
struct Pack<>
{};
template <class T, class... Ts>
struct Pack<T, Ts...>
{
    T head;
    Pack<Ts...> tail;
    template <int I>
    GADFIT_HD const auto& get() const
    {
        if constexpr (I == 0) {
            return head;
        } else {
            return tail.template get<I - 1>();
        }
    }
};
GADFIT_HD Pack<> pack()
{
    return {};
}
template <class T, class... Ts>
GADFIT_HD Pack<T, Ts...> pack(const T& v, const Ts&... rest)
{
    return { v, pack(rest...) };
}

// (element by element: a LaneTable of `mine` goes into the Table of `dst`)
template <class A, class... As, class B, class... Bs>
GADFIT_HD void store_prepared(const Pack<A, As...>& mine, Pack<B, Bs...>& dst, const int lane)
{
    static_assert(sizeof...(As) == sizeof...(Bs));
    store_prepared(mine.head, dst.head, lane);
    if constexpr (sizeof...(As) > 0) {
        store_prepared(mine.tail, dst.tail, lane);
    }
}

template <class Model, class Ctx>
concept HasPrepare = requires(const Model& m, const Ctx& c) { m.prepare(c); };

struct NoPrepared
{};

template <class Model, class Ctx, bool = HasPrepare<Model, Ctx>>
struct PreparedOf
{
    using type = NoPrepared;
};
template <class Model, class Ctx>
struct PreparedOf<Model, Ctx, true>
{
    using type = decltype(std::declval<const Model&>().prepare(std::declval<const Ctx&>()));
};
template <class Model, class Ctx>
using prepared_t = typename PreparedOf<Model, Ctx>::type;

// What prepare() gives for this context (an empty struct without the hook)
template <class Model, class Ctx>
GADFIT_HD auto prepare_of(const Model& m, const Ctx& c) -> prepared_t<Model, Ctx>
{
    if constexpr (HasPrepare<Model, Ctx>) {
        return m.prepare(c);
    } else {
        return {};
    }
}

// f(x) given what prepare_of() returned for the same context
template <class Model, class Ctx, class Pre>
GADFIT_HD auto evaluate_prepared(const Model& m, const Ctx& c, const Pre& pre, const double x)
{
    if constexpr (HasPrepare<Model, Ctx>) {
        return m(c, pre, x);
    } else {
        return m(c, x);
    }
}

// f(x) from scratch: the entry point of contexts that keep no cache
template <class Model, class Ctx>
GADFIT_HD auto evaluate(const Model& m, const Ctx& c, const double x)
{
    if constexpr (HasPrepare<Model, Ctx>) {
        return m(c, m.prepare(c), x);
    } else {
        return m(c, x);
    }
}

// Semi-infinite integration ranges, as in the Fortran version
// (fortran/gadfit/numerical_integration.F90:286-352): the integral over
// (lower, +inf) is mapped to (0, 1] by x = lower + (1 - t)/t, the one over
// (-inf, upper) by x = upper - (1 - t)/t, and the adaptive rule is applied to
// the transformed integrand f(x)/t^2.  Written against the context interface
// only, so the same code serves the device contexts, the reference's AdVar
// and the oracle's dual.
template <class Ctx, class F>
GADFIT_HD auto integrate_to_infinity(const Ctx& c,
                                     const double lower,
                                     F&& f,
                                     const double rel_error,
                                     const double abs_error = 0.0)
{
    return c.integrate(
      0.0,
      1.0,
      [&f, lower](const auto& cc, const auto& t) { return f(cc, lower - 1 + 1 / t) / (t * t); },
      rel_error,
      abs_error);
}
template <class Ctx, class F>
GADFIT_HD auto integrate_from_minus_infinity(const Ctx& c,
                                             const double upper,
                                             F&& f,
                                             const double rel_error,
                                             const double abs_error = 0.0)
{
    return c.integrate(
      0.0,
      1.0,
      [&f, upper](const auto& cc, const auto& t) { return f(cc, upper + 1 - 1 / t) / (t * t); },
      rel_error,
      abs_error);
}

// Value of any AD type
GADFIT_HD double value_of(const double x)
{
    return x;
}
template <uint32_t M>
GADFIT_HD double value_of(const SDual<M>& x)
{
    return x.val;
}
GADFIT_HD double value_of(const Dir2& x)
{
    return x.val;
}

} // namespace gadfit_b200
