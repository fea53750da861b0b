// Host-side check of include/gadfit_b200/dual.h against the reference's
// forward-mode known-answer tests (reference c++/tests/ad_forward_mode.cpp) and
// its reverse-mode gradients (c++/tests/ad_reverse_mode.cpp).  The expressions
// live in dual_check_expr.h, which the device check (device_dual_check.cu)
// compiles into a kernel.  Prints "name v0 v1 v2 v3" lines for
// tests/test_dual_host.py.
#include "dual_check_expr.h"

#include <cstdio>

int main()
{
    double out[dual_check::n_entries * dual_check::width];
    dual_check::fill(out);
    for (int e = 0; e < dual_check::n_entries; ++e) {
        const double* v { out + e * dual_check::width };
        std::printf("%s %.17g %.17g %.17g %.17g\n", dual_check::names[e], v[0], v[1], v[2], v[3]);
    }
    return 0;
}
