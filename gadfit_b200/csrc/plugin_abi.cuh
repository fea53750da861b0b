// What a model plug-in (include/gadfit_b200/model_plugin.h) and the library
// must agree on.
#pragma once

#include "model_ops.cuh"

// Bumped whenever ModelOps, SweepArgs, SmallSolveArgs, LmState or LocalFactor change
#define GADFIT_B200_PLUGIN_ABI_VERSION 2

namespace gadfit_b200::plugin {

// Version number plus a layout fingerprint, compared by the loader
constexpr auto abi_fingerprint() -> long long
{
    return GADFIT_B200_PLUGIN_ABI_VERSION * 1000000LL
           + static_cast<long long>(sizeof(SweepArgs)) * 1000
           + static_cast<long long>(sizeof(LmState));
}

} // namespace gadfit_b200::plugin
