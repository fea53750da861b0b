// Host-side check of the Gauss-Kronrod rules of include/gadfit_b200/quadrature.h
// (the same header the kernels compile): the 15-point rule of the C++ reference
// and the 21 ... 61-point rules of the Fortran version.  Prints "name values..."
// lines for tests/test_quadrature_rules.py.
#include "gadfit_b200/fit_function.h"

#include <cmath>
#include <cstdio>

using namespace gadfit_b200;

template <int PTS>
void rule_properties()
{
    using Rule = KronrodRule<PTS>;
    constexpr int n { (PTS - 1) / 2 }; // Gauss order
    double sum_k { 0.0 };
    double sum_g { 0.0 };
    for (int i = 0; i < PTS; ++i) {
        sum_k += Rule::weight(i);
        if (i % 2 == 1) {
            sum_g += Rule::gauss_weight(i / 2);
        }
    }
    std::printf("weights_%d %.17g %.17g\n", PTS, sum_k, sum_g);
    // Exactness: Kronrod integrates x^k exactly up to k = 3n + 1, Gauss up to 2n - 1
    double worst_k { 0.0 };
    double worst_g { 0.0 };
    for (int k = 0; k <= 3 * n + 1; ++k) {
        double ik { 0.0 };
        double ig { 0.0 };
        for (int i = 0; i < PTS; ++i) {
            const double v { std::pow(Rule::root(i), k) };
            ik += Rule::weight(i) * v;
            if (i % 2 == 1) {
                ig += Rule::gauss_weight(i / 2) * v;
            }
        }
        const double exact { k % 2 == 0 ? 2.0 / (k + 1) : 0.0 };
        worst_k = std::fmax(worst_k, std::fabs(ik - exact));
        if (k <= 2 * n - 1) {
            worst_g = std::fmax(worst_g, std::fabs(ig - exact));
        }
    }
    std::printf("exactness_%d %.3e %.3e\n", PTS, worst_k, worst_g);
    // One fixed panel and the adaptive integrator with first derivatives:
    // F(a, b) = int_0^b exp(-a t) cos(3 t) dt, b AD-valued (Leibniz term)
    const double par[2] { 0.7, 2.3 };
    const GradCtx c { par };
    const auto a { c.get<0>() };
    const auto b { c.get<1>() };
    const auto integrand { [&](const auto& ctx, const auto& t) {
        return exp(-1 * (ctx.template get<0>() * t)) * cos(3.0 * t);
    } };
    const auto fixed { kronrod<PTS>(0.0, par[1], [&](const Passive& t) { return integrand(c, t); }) };
    const auto adaptive { c.template integrate<PTS>(0.0, b, integrand, 1e-13) };
    std::printf("fixed_%d %.17g %.17g\n", PTS, fixed.val, fixed.d[0]);
    std::printf("adaptive_%d %.17g %.17g %.17g\n", PTS, adaptive.val, adaptive.d[0], adaptive.d[1]);
    // second directional derivative along (1, 0.5)
    const double seed[2] { 1.0, 0.5 };
    const Dir2Ctx c2 { par, seed };
    const auto second { c2.template integrate<PTS>(0.0, c2.get<1>(), integrand, 1e-13) };
    std::printf("dir2_%d %.17g %.17g %.17g\n", PTS, second.val, second.d, second.dd);
    (void)a;
}

int main()
{
    rule_properties<15>();
    rule_properties<21>();
    rule_properties<31>();
    rule_properties<41>();
    rule_properties<51>();
    rule_properties<61>();
    return 0;
}
