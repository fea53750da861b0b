// The reference's usage example (reference c++/tests/example.cpp:21-41) and its
// first known-answer test (c++/tests/lm_solver.cpp:21-47) written against the
// header-only mirror include/gadfit_b200/lm_solver.h.  Compiled with g++ and
// linked against libgadfit_b200.so by tests/test_cpp_wrapper.py.
//   argv[1] = "host": only exercise the call-order rules (no GPU needed)
//   argv[1] = "fit" : run the known-answer fit on the GPU and print the result
#include "gadfit_b200/lm_solver.h"

#include <cstdio>
#include <cstring>
#include <vector>

int main(int argc, char** argv)
{
    const std::vector<double> x { 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0 };
    const std::vector<double> y1 { 56.0, 51.0, 54.0, 46.0, 50.0, 48.0, 31.0, 45.0 };
    const std::vector<double> y2 { 153.0, 157.0, 137.0, 111.0, 126.0, 108.0, 113.0, 128.0 };
    gadfit::LMsolver solver { "exponential" };
    solver.addDataset(x, y1);
    solver.addDataset(x, y2);
    solver.setPar(0, 50.0, true, 0);
    solver.setPar(2, 1.0, true, 0);
    solver.setPar(0, 150.0, true, 1);
    solver.setPar(2, 5.5, false, 1);
    solver.setPar(1, 10.0, true, "tau");
    int caught { 0 };
    try {
        solver.addDataset(x, y1); // lm_solver.cpp:35-37
    } catch (const gadfit::LateAddDatasetCall&) {
        ++caught;
    }
    try {
        solver.setPar(1, 1.0, true, 7); // lm_solver.cpp:65-68
    } catch (const gadfit::SetParInvalidIndex&) {
        ++caught;
    }
    try {
        gadfit::LMsolver unknown { "no_such_function" };
    } catch (const gadfit::UnknownModel&) {
        ++caught;
    }
    std::printf("caught %d\n", caught);
    if (argc > 2) { // argv[2]: a model plug-in; its functions become constructible
        try {
            gadfit::LMsolver early { "direct_integral_y1" };
            std::printf("plugin function known before loading\n");
        } catch (const gadfit::UnknownModel&) {
            gadfit::loadModels(argv[2]);
            gadfit::LMsolver late { "direct_integral_y1" };
            std::printf("plugin loaded\n");
        }
        try {
            gadfit::loadModels("/nonexistent/libmodels.so");
        } catch (const gadfit::InvalidArgument&) {
            std::printf("bad plugin rejected\n");
        }
    }
    if (argc > 1 && std::strcmp(argv[1], "fit") == 0) {
        // iteration results go to stdout by default, as in the reference; the
        // last fit below shows them, this one is quiet
        solver.settings.verbosity = gadfit::io::hide_all;
        solver.settings.iteration_limit = 6;
        solver.settings.acceleration_threshold = 0.9;
        solver.fit(1.0);
        std::printf("chi2 %.17g\n", solver.chi2());
        std::printf("tau %.17g\n", solver.getParValue(1));
        std::printf("I0_0 %.17g\n", solver.getParValue(0, 0));
        std::printf("bgr_0 %.17g\n", solver.getParValue(2, 0));
        std::printf("I0_1 %.17g\n", solver.getParValue(0, 1));
        std::printf("iterations %d dof %d trials %zu\n", solver.iterations(), solver.degreesOfFreedom(),
                    solver.trace().size());
        const auto& jtj { solver.getJTJ() };
        std::printf("jtj_size %zu\n", jtj.size());
        // default verbosity: the reference's per-iteration report on stdout
        solver.settings.verbosity = gadfit::io::final_only;
        solver.settings.iteration_limit = 1;
        std::printf("report_begin\n");
        solver.fit(1.0);
        std::printf("report_end\n");
    }
    return caught == 3 ? 0 : 1;
}
