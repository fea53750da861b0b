// User-defined fitting functions: model plug-ins.
//
// In the reference a fitting function is a C++ lambda compiled by the user's
// compiler and handed to gadfit::LMsolver (lm_solver.h:146,
// c++/tests/example.cpp:11-19).  Device code cannot be handed over as a
// std::function, so here the user's functor is compiled by nvcc into a small
// shared library -- a model plug-in -- which libgadfit_b200.so loads at run
// time (gadfit_b200_load_models, include/gadfit_b200.h); libgadfit_b200.so
// itself is not rebuilt.
//
//     // my_models.cu
//     #include "gadfit_b200/model_plugin.h"
//     struct Decay {
//         static constexpr int n_pars { 3 };
//         static constexpr const char* name { "decay" };
//         template <class Ctx>
//         GADFIT_HD auto operator()(const Ctx& c, const double x) const
//         {
//             using namespace gadfit_b200;
//             return par<0>(c) * exp(-x / par<1>(c)) + par<2>(c);
//         }
//     };
//     GADFIT_B200_REGISTER_MODEL(Decay)
//
//     $ python -m gadfit_b200.build --plugin my_models.cu -o libmy_models.so
//
// A plug-in may consist of several translation units, each registering any
// number of functors.  The functor is written exactly like the built-in ones
// (include/gadfit_b200/fit_models.h): a call operator templated on the
// evaluation context, parameters read with par<K>(c), the overloaded
// elementary functions of dual.h, c.kronrod / c.integrate / c.integrate2d for
// quadrature.
#pragma once

#include "plugin_abi.cuh" // gadfit_b200/csrc: the kernel templates (model_ops.cuh)

#include <vector>

namespace gadfit_b200::plugin {

struct Entry
{
    const char* name;
    int n_pars;
    ModelOps* (*make)();
};

// One registry per plug-in library (hidden visibility: not shared with
// libgadfit_b200.so or with other plug-ins)
inline auto registry() -> std::vector<Entry>&
{
    static std::vector<Entry> entries;
    return entries;
}

template <class Model>
struct Registrar
{
    Registrar()
    {
        registry().push_back({ Model::name, Model::n_pars, []() -> ModelOps* {
                                  // more than 32 parameters: listed, but the kernels cannot
                                  // be instantiated (32-bit role / sparsity masks);
                                  // gadfit_b200_create answers GADFIT_B200_TOO_MANY_PARAMETERS
                                  if constexpr (Model::n_pars <= kMaxShared) {
                                      return new ModelOpsImpl<Model> {};
                                  } else {
                                      return nullptr;
                                  }
                              } });
    }
};

} // namespace gadfit_b200::plugin

#define GADFIT_B200_PLUGIN_CAT2(a, b) a##b
#define GADFIT_B200_PLUGIN_CAT(a, b) GADFIT_B200_PLUGIN_CAT2(a, b)
#define GADFIT_B200_REGISTER_MODEL(Model)                                                         \
    namespace {                                                                                   \
    const gadfit_b200::plugin::Registrar<Model> GADFIT_B200_PLUGIN_CAT(gadfit_b200_registrar_,    \
                                                                       __COUNTER__) {};           \
    }

// Entry points looked up by gadfit_b200_load_models.  `inline` + `used`: every
// translation unit of the plug-in may include this header; the linker keeps
// one copy.
#define GADFIT_B200_PLUGIN_EXPORT extern "C" __attribute__((visibility("default"), used)) inline

GADFIT_B200_PLUGIN_EXPORT long long gadfit_b200_plugin_abi(void)
{
    return gadfit_b200::plugin::abi_fingerprint();
}
GADFIT_B200_PLUGIN_EXPORT int gadfit_b200_plugin_n_models(void)
{
    return static_cast<int>(gadfit_b200::plugin::registry().size());
}
GADFIT_B200_PLUGIN_EXPORT int gadfit_b200_plugin_model_info(int i, const char** name, int* n_pars)
{
    const auto& r { gadfit_b200::plugin::registry() };
    if (i < 0 || i >= static_cast<int>(r.size())) {
        return 1;
    }
    *name = r[i].name;
    *n_pars = r[i].n_pars;
    return 0;
}
GADFIT_B200_PLUGIN_EXPORT void* gadfit_b200_plugin_make(int i)
{
    const auto& r { gadfit_b200::plugin::registry() };
    if (i < 0 || i >= static_cast<int>(r.size())) {
        return nullptr;
    }
    return r[i].make();
}
