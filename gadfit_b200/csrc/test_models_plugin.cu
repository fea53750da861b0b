// Model plug-in with the fitting functions of the reference's nested and
// direct double-integral tests (include/gadfit_b200/test_models.h).  One
// translation unit per function (-DGADFIT_TEST_MODEL=k, they are template
// heavy), linked into libgadfit_b200_testmodels.so by gadfit_b200/build.py.
#include "gadfit_b200/model_plugin.h"
#include "gadfit_b200/test_models.h"

#ifndef GADFIT_TEST_MODEL
#error "compile with -DGADFIT_TEST_MODEL=<k>"
#endif

using TestModel = std::tuple_element_t<GADFIT_TEST_MODEL, gadfit_b200::models::TestModels>;
GADFIT_B200_REGISTER_MODEL(TestModel)
