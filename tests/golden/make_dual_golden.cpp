// TEST INFRASTRUCTURE.  Evaluates the expressions of tests/dual_check_expr.h
// with the UNMODIFIED reference's gadfit::AdVar in forward mode
// (c++/gadfit/automatic_differentiation.cpp, compiled where it lies) and prints
// "name val d dd" lines: the golden values of tests/golden/dual_ops_ref.json.
// Built and run by tests/golden/make_dual_golden.py in the build container
// (the reference tree does not exist on the GPU box).
#include "../dual_check_expr.h"

#include <gadfit/automatic_differentiation.h>

#include <cstdio>

using gadfit::AdVar;

static auto active(const double val, const double d, const double dd) -> AdVar
{
    return AdVar { val, d, dd, gadfit::forward_active_idx };
}

int main()
{
    using namespace dual_check;
    const AdVar s2 { active(fix_d3, 0.7, 0.3) };
    const AdVar g2 { active(fix_d1, -0.4, 0.2) };
    const AdVar t2 { active(0.25, 1.0, -0.5) };
    int e { 23 }; // first single-operation entry of dual_check::names
    single_ops(s2, g2, t2, [&](const AdVar& r) {
        std::printf("%s %.17g %.17g %.17g\n", names[e++], r.val, r.d, r.dd);
    });
    // first order, one seed at a time (d = 1, dd = 0): the gradient entries
    const AdVar zs { AdVar { fix_d3, 0.0, 0.0, gadfit::passive_idx } };
    const AdVar zg { AdVar { fix_d1, 0.0, 0.0, gadfit::passive_idx } };
    const AdVar zt { AdVar { 0.25, 0.0, 0.0, gadfit::passive_idx } };
    double first[25][3] {};
    int k { 0 };
    // seed on (s, g): both are "parameter 0" of their own entries
    single_ops(active(fix_d3, 1.0, 0.0), active(fix_d1, 1.0, 0.0), zt, [&](const AdVar& r) {
        first[k][0] = r.val;
        first[k++][1] = r.d;
    });
    k = 0;
    single_ops(zs, zg, active(0.25, 1.0, 0.0), [&](const AdVar& r) { first[k++][2] = r.d; });
    for (int i = 0; i < 25; ++i) {
        std::printf("%s %.17g %.17g %.17g\n", names[e++], first[i][0], first[i][1], first[i][2]);
    }
    return 0;
}
