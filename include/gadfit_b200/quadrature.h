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

// numerical_integration.cpp:133-148
template <class F>
GADFIT_HD auto kronrod(const double lower, const double upper, F&& f)
{
    const double scale { (upper - lower) / 2 };
    const double shift { (lower + upper) / 2 };
    auto sum { k15_weight(0) * f(Passive { scale * k15_root(0) + shift }) };
#pragma unroll 1
    for (int i = 1; i < k15_size; ++i) {
        sum = sum + k15_weight(i) * f(Passive { scale * k15_root(i) + shift });
    }
    return scale * sum;
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

} // namespace gadfit_b200
