// The expressions of the reference's forward-mode known-answer tests
// (reference c++/tests/ad_forward_mode.cpp:16-256) and reverse-mode gradients
// (c++/tests/ad_reverse_mode.cpp:13-281) over include/gadfit_b200/dual.h,
// written once for the host build (tests/host_dual_check.cpp) and for the
// kernel of tests/device_dual_check.cu.  Activity patterns are chosen by the
// TYPES of the arguments (Dir2 = forward active, Passive = passive,
// SDual<bit> = one gradient seed).  dual_check_fill() writes, per entry of
// dual_check_names, (val, d, dd) for the forward-mode entries and
// (val, d/da, d/db, d/dc) for the gradient entries: dual_check_width doubles
// per entry.
#pragma once

#include "gadfit_b200/dual.h"

namespace dual_check {

using namespace gadfit_b200;

// c++/tests/fixtures.h:9-11
constexpr double fix_d0 { 6.1360420701563498 };
constexpr double fix_d1 { 2.9606444748278875 };
constexpr double fix_d3 { 0.5356792380861322 };
constexpr double fix_d4 { -1.4727205961479033 };
constexpr double fix_d5 { 4.2512661200877879 };
constexpr double fix_d6 { -2.9410316453781444 };
constexpr double par_d { fix_d0 };
constexpr int par_i { static_cast<int>(fix_d1) };

// ad_forward_mode.cpp:21-26
template <class A, class B, class Cc>
GADFIT_HD auto arithmetic(const A& a, const B& b, const Cc& c)
{
    return par_d * (a / par_d / par_i + par_d) + b * (par_d - c)
           - (c - par_d) / (par_d + a - b + par_d + par_i) + par_d / (-b) * par_d;
}
// ad_forward_mode.cpp:124-128
template <class A, class B>
GADFIT_HD auto exponentiation(const A& a, const B& b)
{
    return pow(b, a) + pow(b, par_d) / pow(b, par_i) - pow(par_d, +a) / pow(par_i, a) * abs(a) * abs(b)
           + exp(sqrt(abs(a)) + log(b)) / sqrt(log(-b / a) * par_d);
}
// ad_forward_mode.cpp:176-181: sin cos tan asin acos atan sinh cosh tanh asinh
// acosh atanh in one expression
template <class A, class B>
GADFIT_HD auto trigonometric(const A& a, const B& b)
{
    return sin(a * b) * cos(a) / cos(b) + tan(cos(a)) / atan(b * asin(1 / a) / acos(a / b))
           + sinh(a / b) * pow(cosh(b / a), tanh(b / a))
           + asinh(a / b) * pow(acosh(abs(b / a)), atanh(abs(a / b)));
}

constexpr int width { 4 };
constexpr const char* names[] {
    "arith_none", "arith_a",   "arith_b",    "arith_c",   "arith_ab",   "arith_ac",
    "arith_bc",   "arith_abc", "exp_none",   "exp_a",     "exp_b",      "exp_ab",
    "trig_none",  "trig_a",    "trig_b",     "trig_ab",   "erf_none",   "erf_a",
    "grad_arith", "grad_exp",  "grad_trig",  "grad_erf",  "grad_arith_b_only",
    // every elemental operation on its own (value, first and second
    // directional derivative through Dir2; gradient through SDual)
    "op_sin", "op_cos", "op_tan", "op_asin", "op_acos", "op_atan", "op_sinh", "op_cosh",
    "op_tanh", "op_asinh", "op_acosh", "op_atanh", "op_exp", "op_log", "op_sqrt", "op_abs",
    "op_erf", "op_pow_aa", "op_pow_ar", "op_pow_ra", "op_pow_ai", "op_add", "op_sub", "op_mul",
    "op_div",
    "gop_sin", "gop_cos", "gop_tan", "gop_asin", "gop_acos", "gop_atan", "gop_sinh", "gop_cosh",
    "gop_tanh", "gop_asinh", "gop_acosh", "gop_atanh", "gop_exp", "gop_log", "gop_sqrt", "gop_abs",
    "gop_erf", "gop_pow_aa", "gop_pow_ar", "gop_pow_ra", "gop_pow_ai", "gop_add", "gop_sub",
    "gop_mul", "gop_div"
};
constexpr int n_entries { sizeof(names) / sizeof(names[0]) };

GADFIT_HD void put(double*& out, const Dir2& r)
{
    out[0] = r.val;
    out[1] = r.d;
    out[2] = r.dd;
    out[3] = 0.0;
    out += width;
}
GADFIT_HD void put(double*& out, const Passive& r)
{
    out[0] = r.val;
    out[1] = out[2] = out[3] = 0.0;
    out += width;
}
template <uint32_t M>
GADFIT_HD void put_grad(double*& out, const SDual<M>& r)
{
    out[0] = r.val;
    out[1] = r.template tangent<0>();
    out[2] = r.template tangent<1>();
    out[3] = r.template tangent<2>();
    out += width;
}

// The single-operation entries: arguments inside every function's domain
// (|x| < 1 for asin/acos/atanh, x > 1 for acosh, x > 0 for log/sqrt/pow)
template <class A, class B, class Put>
GADFIT_HD void single_ops(const A& s, const A& g, const B& t, Put&& emit)
{
    // s in (0, 1), g > 1, t: a second active variable in (0, 1)
    emit(sin(s));
    emit(cos(s));
    emit(tan(s));
    emit(asin(s));
    emit(acos(s));
    emit(atan(g));
    emit(sinh(g));
    emit(cosh(g));
    emit(tanh(s));
    emit(asinh(g));
    emit(acosh(g));
    emit(atanh(s));
    emit(exp(s));
    emit(log(g));
    emit(sqrt(g));
    emit(abs(-1 * g));
    emit(erf(s));
    emit(pow(g, t));
    emit(pow(g, par_d));
    emit(pow(par_d, t));
    emit(pow(g, par_i));
    emit(s + t);
    emit(s - t);
    emit(s * t);
    emit(s / t);
}

GADFIT_HD void fill(double* out)
{
    const Passive pa { fix_d4 }, pb { fix_d5 }, pc { fix_d6 };
    const Dir2 a { fix_d4, 1.0, 1.0 }, b { fix_d5, 1.0, 1.0 }, c { fix_d6, 1.0, 1.0 };
    put(out, arithmetic(pa, pb, pc));
    put(out, arithmetic(a, pb, pc));
    put(out, arithmetic(pa, b, pc));
    put(out, arithmetic(pa, pb, c));
    put(out, arithmetic(a, b, pc));
    put(out, arithmetic(a, pb, c));
    put(out, arithmetic(pa, b, c));
    put(out, arithmetic(a, b, c));
    put(out, exponentiation(pa, pb));
    put(out, exponentiation(a, pb));
    put(out, exponentiation(pa, b));
    put(out, exponentiation(a, b));
    put(out, trigonometric(pa, pb));
    put(out, trigonometric(a, pb));
    put(out, trigonometric(pa, b));
    put(out, trigonometric(a, b));
    put(out, erf(Passive { fix_d3 }));
    put(out, erf(Dir2 { fix_d3, 1.0, 1.0 }));
    // Gradients with static sparsity (one tangent per parameter bit)
    SDual<1U> ga;
    ga.val = fix_d4;
    ga.d[0] = 1.0;
    SDual<2U> gb;
    gb.val = fix_d5;
    gb.d[0] = 1.0;
    SDual<4U> gc;
    gc.val = fix_d6;
    gc.d[0] = 1.0;
    put_grad(out, arithmetic(ga, gb, gc));
    put_grad(out, exponentiation(ga, gb));
    put_grad(out, trigonometric(ga, gb));
    SDual<1U> ge;
    ge.val = fix_d3;
    ge.d[0] = 1.0;
    put_grad(out, erf(ge));
    // Partial activity through the types: only b carries a tangent
    put_grad(out, arithmetic(pa, gb, pc));
    // One operation at a time, second order (directions d = 0.7 / -0.4 with a
    // non-zero dd seed so that every term of the rules is exercised) ...
    const Dir2 s2 { fix_d3, 0.7, 0.3 }, g2 { fix_d1, -0.4, 0.2 }, t2 { 0.25, 1.0, -0.5 };
    single_ops(s2, g2, t2, [&](const Dir2& r) { put(out, r); });
    // ... and first order with two independent seeds
    SDual<1U> s1, g1;
    s1.val = fix_d3;
    s1.d[0] = 1.0;
    g1.val = fix_d1;
    g1.d[0] = 1.0;
    SDual<2U> t1;
    t1.val = 0.25;
    t1.d[0] = 1.0;
    single_ops(s1, g1, t1, [&](const auto& r) { put_grad(out, r); });
}

} // namespace dual_check
