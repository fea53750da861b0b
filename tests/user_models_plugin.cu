// A user's model plug-in as INTEGRATION.md describes it: two fitting functions
// that are not in libgadfit_b200.so.  tests/test_model_plugin.py builds it with
// `python -m gadfit_b200.build --plugin`, loads it and fits with it.
#include "gadfit_b200/model_plugin.h"

// The reference's example function (c++/tests/example.cpp:11-19) under a
// user's own name: must reproduce the known answers of the built-in one.
struct UserDecay
{
    static constexpr int n_pars { 3 };
    static constexpr const char* name { "user_decay" };
    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        using namespace gadfit_b200;
        const auto I0 { par<0>(c) };
        const auto tau { par<1>(c) };
        const auto bgr { par<2>(c) };
        return I0 * exp(-x / tau) + bgr;
    }
};
GADFIT_B200_REGISTER_MODEL(UserDecay)

// Stretched exponential with an adaptive integral:  A int_0^x e^{-(t/tau)^beta} dt + c
struct UserStretched
{
    static constexpr int n_pars { 4 };
    static constexpr const char* name { "user_stretched_integral" };
    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        using namespace gadfit_b200;
        const auto integrand { [](const auto& cc, const auto& t) {
            return exp(-pow(t / par<1>(cc), par<2>(cc)));
        } };
        return par<0>(c) * c.integrate(0.0, x, integrand, 1e-10) + par<3>(c);
    }
};
GADFIT_B200_REGISTER_MODEL(UserStretched)

// The same function with the Fortran version's other Gauss-Kronrod rules
// (set_integration_rule, fortran/gadfit/numerical_integration.F90:139-171):
// the rule is a template argument of c.integrate
template <int PTS>
struct UserStretchedRule
{
    static constexpr int n_pars { 4 };
    static constexpr const char* name { PTS == 21   ? "user_stretched_integral_k21"
                                        : PTS == 31 ? "user_stretched_integral_k31"
                                        : PTS == 41 ? "user_stretched_integral_k41"
                                        : PTS == 51 ? "user_stretched_integral_k51"
                                                    : "user_stretched_integral_k61" };
    template <class Ctx>
    GADFIT_HD auto operator()(const Ctx& c, const double x) const
    {
        using namespace gadfit_b200;
        const auto integrand { [](const auto& cc, const auto& t) {
            return exp(-pow(t / par<1>(cc), par<2>(cc)));
        } };
        return par<0>(c) * c.template integrate<PTS>(0.0, x, integrand, 1e-10) + par<3>(c);
    }
};
GADFIT_B200_REGISTER_MODEL(UserStretchedRule<21>)
GADFIT_B200_REGISTER_MODEL(UserStretchedRule<31>)
GADFIT_B200_REGISTER_MODEL(UserStretchedRule<41>)
GADFIT_B200_REGISTER_MODEL(UserStretchedRule<51>)
GADFIT_B200_REGISTER_MODEL(UserStretchedRule<61>)
