// Host-side check of include/gadfit_b200/dual.h against the reference's
// forward-mode known-answer tests (reference c++/tests/ad_forward_mode.cpp) and
// its reverse-mode gradients (c++/tests/ad_reverse_mode.cpp).  The expressions
// are the ones those files test; activity patterns are chosen by the TYPES of
// the arguments (Dir2 = forward active, Passive = passive, SDual<bit> = one
// gradient seed).  Prints "name val d dd" lines for tests/test_dual_host.py.
#include "gadfit_b200/dual.h"

#include <cstdio>

using namespace gadfit_b200;

constexpr double fix_d[] { 6.1360420701563498,  2.9606444748278875,  9.9253736972586246,
                           0.5356792380861322,  -1.4727205961479033, 4.2512661200877879,
                           -2.9410316453781444 };
constexpr double par_d { fix_d[0] };
constexpr int par_i { static_cast<int>(fix_d[1]) };

// ad_forward_mode.cpp:21-26
template <class A, class B, class Cc>
auto arithmetic(const A& a, const B& b, const Cc& c)
{
    return par_d * (a / par_d / par_i + par_d) + b * (par_d - c)
           - (c - par_d) / (par_d + a - b + par_d + par_i) + par_d / (-b) * par_d;
}
// ad_forward_mode.cpp:124-128
template <class A, class B>
auto exponentiation(const A& a, const B& b)
{
    return pow(b, a) + pow(b, par_d) / pow(b, par_i) - pow(par_d, +a) / pow(par_i, a) * abs(a) * abs(b)
           + exp(sqrt(abs(a)) + log(b)) / sqrt(log(-b / a) * par_d);
}
// ad_forward_mode.cpp:176-181
template <class A, class B>
auto trigonometric(const A& a, const B& b)
{
    return sin(a * b) * cos(a) / cos(b) + tan(cos(a)) / atan(b * asin(1 / a) / acos(a / b))
           + sinh(a / b) * pow(cosh(b / a), tanh(b / a))
           + asinh(a / b) * pow(acosh(abs(b / a)), atanh(abs(a / b)));
}

void print(const char* name, const Dir2& r) { std::printf("%s %.17g %.17g %.17g\n", name, r.val, r.d, r.dd); }
void print(const char* name, const Passive& r) { std::printf("%s %.17g 0 0\n", name, r.val); }
template <uint32_t M>
void print_grad(const char* name, const SDual<M>& r)
{
    std::printf("%s %.17g", name, r.val);
    std::printf(" %.17g %.17g %.17g\n", r.template tangent<0>(), r.template tangent<1>(),
                r.template tangent<2>());
}

int main()
{
    const Passive pa { fix_d[4] }, pb { fix_d[5] }, pc { fix_d[6] };
    const Dir2 a { fix_d[4], 1.0, 1.0 }, b { fix_d[5], 1.0, 1.0 }, c { fix_d[6], 1.0, 1.0 };
    print("arith_none", arithmetic(pa, pb, pc));
    print("arith_a", arithmetic(a, pb, pc));
    print("arith_b", arithmetic(pa, b, pc));
    print("arith_c", arithmetic(pa, pb, c));
    print("arith_ab", arithmetic(a, b, pc));
    print("arith_ac", arithmetic(a, pb, c));
    print("arith_bc", arithmetic(pa, b, c));
    print("arith_abc", arithmetic(a, b, c));
    print("exp_none", exponentiation(pa, pb));
    print("exp_a", exponentiation(a, pb));
    print("exp_b", exponentiation(pa, b));
    print("exp_ab", exponentiation(a, b));
    print("trig_none", trigonometric(pa, pb));
    print("trig_a", trigonometric(a, pb));
    print("trig_b", trigonometric(pa, b));
    print("trig_ab", trigonometric(a, b));
    print("erf_none", erf(Passive { fix_d[3] }));
    print("erf_a", erf(Dir2 { fix_d[3], 1.0, 1.0 }));
    // Gradients with static sparsity (one tangent per parameter bit)
    SDual<1U> ga;
    ga.val = fix_d[4];
    ga.d[0] = 1.0;
    SDual<2U> gb;
    gb.val = fix_d[5];
    gb.d[0] = 1.0;
    SDual<4U> gc;
    gc.val = fix_d[6];
    gc.d[0] = 1.0;
    print_grad("grad_arith", arithmetic(ga, gb, gc));
    print_grad("grad_exp", exponentiation(ga, gb));
    print_grad("grad_trig", trigonometric(ga, gb));
    SDual<1U> ge;
    ge.val = fix_d[3];
    ge.d[0] = 1.0;
    print_grad("grad_erf", erf(ge));
    // Partial activity through the types: only b carries a tangent
    print_grad("grad_arith_b_only", arithmetic(pa, gb, pc));
    return 0;
}
