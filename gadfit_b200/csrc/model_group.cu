// Kernel instantiations of one group of fitting functions (see
// include/gadfit_b200/fit_models.h, for_each_model_in_group).  Compiled once
// per group with -DGADFIT_MODEL_GROUP=k by gadfit_b200/build.py.
#include "model_ops.cuh"

#include <string>

#ifndef GADFIT_MODEL_GROUP
#error "compile with -DGADFIT_MODEL_GROUP=<k>"
#endif

namespace gadfit_b200 {

template <>
auto make_model_ops_group<GADFIT_MODEL_GROUP>(const char* model_name) -> std::unique_ptr<ModelOps>
{
    std::unique_ptr<ModelOps> ops;
    models::for_each_model_in_group<GADFIT_MODEL_GROUP>([&]<class Model>() {
        if (!ops && std::string { model_name } == Model::name) {
            ops = std::make_unique<ModelOpsImpl<Model>>();
        }
    });
    return ops;
}

} // namespace gadfit_b200
