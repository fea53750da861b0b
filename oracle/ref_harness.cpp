// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.
//
// C-ABI around the UNMODIFIED reference solver gadfit::LMsolver.  This file
// is compiled TOGETHER with the reference's own sources, taken where they lie
// under /root/reference (see oracle/Makefile; nothing is copied), into
// oracle/_ref/libgadfit_ref.so.  It exports the same call shapes as the
// product's C-ABI (include/gadfit_b200.h) under the prefix gadfit_ref_, so a
// parity test is the same script run against two libraries.
//
// What is the reference here: c++/gadfit/{automatic_differentiation,
// fit_function,lm_solver,exceptions,numerical_integration,timer}.cpp plus
// lapack_fallback.cpp (in-tree linear algebra, the default FALLBACK_LINALG=ON
// of the reference's CMakeLists.txt:31-32) or lapack.cpp + a BLAS.
// The fitting functions are the shared functors of
// include/gadfit_b200/fit_models.h instantiated with gadfit::AdVar, i.e. they
// run through the reference's reverse-mode tape (Jacobian) and forward mode
// (omega) exactly like a user lambda would.
//
// The accept/reject trace is the reference's own debug output
// (lm_solver.cpp:457-460), captured as doubles by oracle/spdlog_stub.

#include <gadfit/lm_solver.h>
#include <gadfit/numerical_integration.h>
#include <spdlog/spdlog.h>

#include "gadfit_b200.h"
#include "gadfit_b200/fit_models.h"
#include "gadfit_b200/test_models.h"

#include <algorithm>
#include <cstring>
#include <memory>
#include <omp.h>
#include <string>
#include <vector>

namespace {

// Evaluation context over the reference's AD variable
struct RefCtx
{
    const std::vector<gadfit::AdVar>& pars;

    template <int K>
    auto get() const -> const gadfit::AdVar&
    {
        return pars[K];
    }
    template <class F>
    auto kronrod(const double lower, const double upper, F&& f) const
      -> gadfit::AdVar
    {
        return gadfit::kronrod(
          [&f](const std::vector<gadfit::AdVar>&, const gadfit::AdVar& t) {
              return f(t);
          },
          pars,
          lower,
          upper);
    }
    // The integrand also receives the index of the node it is called at
    // (include/gadfit_b200/fit_function.h): gadfit::kronrod visits its nodes in
    // table order (numerical_integration.cpp:141-145, 204-212), so counting
    // the calls gives the index.
    template <class F>
    auto kronrod_indexed(const double lower, const double upper, F&& f) const -> gadfit::AdVar
    {
        int i { 0 };
        return kronrod(lower, upper, [&](const gadfit::AdVar& t) { return f(t, i++); });
    }
    template <class F>
    auto kronrod2d_indexed(const double outer_lower,
                           const double outer_upper,
                           const double inner_lower,
                           const double inner_upper,
                           F&& f) const -> gadfit::AdVar
    {
        int n { 0 };
        return kronrod2d(outer_lower, outer_upper, inner_lower, inner_upper,
                         [&](const gadfit::AdVar& u, const gadfit::AdVar& v) {
                             const int i { n / 15 };
                             const int j { n % 15 };
                             ++n;
                             return f(u, v, i, j);
                         });
    }
    // gadfit::integrate, numerical_integration.h:83-113 (double or AdVar bounds)
    template <class BL, class BU, class F>
    auto integrate(const BL& lower,
                   const BU& upper,
                   F&& f,
                   const double rel_error = gadfit::default_rel_error,
                   const double abs_error = 0.0) const -> gadfit::AdVar
    {
        return gadfit::integrate(
          [&f](const std::vector<gadfit::AdVar>& p, const gadfit::AdVar& t) {
              return gadfit::AdVar { f(RefCtx { p }, t) };
          },
          pars,
          lower,
          upper,
          rel_error,
          abs_error);
    }
    // direct 2-D gadfit::integrate, numerical_integration.h:116-260 (every
    // combination of double and AdVar bounds)
    template <class BY1, class BY2, class BX1, class BX2, class F>
    auto integrate2d(const BY1& y1,
                     const BY2& y2,
                     const BX1& x1,
                     const BX2& x2,
                     F&& f,
                     const double rel_error = gadfit::default_rel_error,
                     const double abs_error = 0.0) const -> gadfit::AdVar
    {
        return gadfit::integrate(
          [&f](const std::vector<gadfit::AdVar>& p,
               const gadfit::AdVar& x,
               const gadfit::AdVar& y) {
              return gadfit::AdVar { f(RefCtx { p }, x, y) };
          },
          pars,
          y1,
          y2,
          x1,
          x2,
          rel_error,
          abs_error);
    }
    template <class F>
    auto kronrod2d(const double outer_lower,
                   const double outer_upper,
                   const double inner_lower,
                   const double inner_upper,
                   F&& f) const -> gadfit::AdVar
    {
        return gadfit::kronrod(
          [&f](const std::vector<gadfit::AdVar>&,
               const gadfit::AdVar& u,
               const gadfit::AdVar& v) { return f(u, v); },
          pars,
          outer_lower,
          outer_upper,
          inner_lower,
          inner_upper);
    }
};

template <class Model>
auto makeSignature() -> gadfit::fitSignature
{
    return [](const std::vector<gadfit::AdVar>& parameters, const double x) {
        return gadfit::AdVar { gadfit_b200::evaluate(Model {}, RefCtx { parameters }, x) };
    };
}

struct RefSolver
{
    std::unique_ptr<gadfit::LMsolver> solver;
    std::vector<gadfit_b200_trial> trace;
    gadfit_b200_timings timings {};
    int n_iterations {};
    std::vector<double> DTD_min;
    std::string last_error;
    std::string log; // info-level text of the last fit, one line per message
};

template <class F>
auto guarded(RefSolver* h, F&& f) -> int
{
    try {
        f();
        return GADFIT_B200_OK;
    } catch (const gadfit::UnknownOperation& e) {
        h->last_error = e.what();
        return GADFIT_B200_UNKNOWN_OPERATION;
    } catch (const gadfit::LateAddDatasetCall& e) {
        h->last_error = e.what();
        return GADFIT_B200_LATE_ADD_DATASET_CALL;
    } catch (const gadfit::SetParInvalidIndex& e) {
        h->last_error = e.what();
        return GADFIT_B200_SET_PAR_INVALID_INDEX;
    } catch (const gadfit::UninitializedParameter& e) {
        h->last_error = e.what();
        return GADFIT_B200_UNINITIALIZED_PARAMETER;
    } catch (const gadfit::NegativeDegreesOfFreedom& e) {
        h->last_error = e.what();
        return GADFIT_B200_NEGATIVE_DEGREES_OF_FREEDOM;
    } catch (const gadfit::NoFittingParameters& e) {
        h->last_error = e.what();
        return GADFIT_B200_NO_FITTING_PARAMETERS;
    } catch (const gadfit::NoGlobalParameters& e) {
        h->last_error = e.what();
        return GADFIT_B200_NO_GLOBAL_PARAMETERS;
    } catch (const gadfit::InsufficientIntegrationWorkspace& e) {
        h->last_error = e.what();
        return GADFIT_B200_INSUFFICIENT_INTEGRATION_WORKSPACE;
    } catch (const std::exception& e) {
        h->last_error = e.what();
        return GADFIT_B200_INVALID_ARGUMENT;
    }
}

} // namespace

extern "C" {

int gadfit_ref_create(const char* model_name, RefSolver** out)
{
    static bool integration_ready { false };
    if (!integration_ready) {
        // Two workspaces: nested integrals need the nesting level stated
        // (c++/tests/numerical_integration.cpp:226-231)
        gadfit::initIntegration(gadfit::default_workspace_size, 2);
        integration_ready = true;
    }
    std::unique_ptr<gadfit::LMsolver> solver;
    const auto match { [&]<class Model>() {
        if (std::strcmp(model_name, Model::name) == 0) {
            solver =
              std::make_unique<gadfit::LMsolver>(makeSignature<Model>());
        }
    } };
    gadfit_b200::models::for_each_model(match);
    gadfit_b200::models::for_each_test_model(match);
    if (!solver) {
        return GADFIT_B200_UNKNOWN_MODEL;
    }
    *out = new RefSolver {};
    (*out)->solver = std::move(solver);
    return GADFIT_B200_OK;
}

int gadfit_ref_destroy(RefSolver* h)
{
    delete h;
    return GADFIT_B200_OK;
}

int gadfit_ref_add_dataset(RefSolver* h,
                           int n,
                           const double* x,
                           const double* y,
                           const double* errors)
{
    return guarded(h, [&] { h->solver->addDataset(n, x, y, errors); });
}

int gadfit_ref_set_par(RefSolver* h,
                       int i_par,
                       double val,
                       int active,
                       int i_dataset,
                       const char* name)
{
    return guarded(h, [&] {
        h->solver->setPar(
          i_par, val, active != 0, i_dataset, name ? std::string { name } : "");
    });
}

int gadfit_ref_default_settings(gadfit_b200_settings* s)
{
    // lm_solver.h:88-96
    *s = { 1000, 3, 10.0, 10.0, 1, nullptr, 0, -1.0, 0U, 0, 1 };
    return GADFIT_B200_OK;
}

int gadfit_ref_set_settings(RefSolver* h, const gadfit_b200_settings* s)
{
    auto& dst { h->solver->settings };
    dst.iteration_limit = s->iteration_limit;
    dst.lambda_incs = s->lambda_incs;
    dst.lambda_down = s->lambda_down;
    dst.lambda_up = s->lambda_up;
    dst.damp_max = s->damp_max != 0;
    if (s->DTD_min != nullptr && s->n_DTD_min > 0) {
        dst.DTD_min.assign(s->DTD_min, s->DTD_min + s->n_DTD_min);
    } else {
        dst.DTD_min = { 0.0 };
    }
    dst.acceleration_threshold = s->acceleration_threshold;
    // Always request the timing table: the stub turns it into numbers
    dst.verbosity = gadfit::io::flag { s->verbosity } | gadfit::io::timings;
    dst.loss = static_cast<gadfit::Loss>(s->loss);
    dst.n_threads = s->n_threads > 0 ? s->n_threads : omp_get_max_threads();
    return GADFIT_B200_OK;
}

// The reference's own info-level output of the last fit (what spdlog would
// have printed with the pattern "%v"), lines separated by '\n'
int gadfit_ref_get_log(RefSolver* h, char* buf, size_t cap, size_t* need)
{
    *need = h->log.size() + 1;
    if (buf != nullptr && cap > 0) {
        const size_t n { std::min(cap - 1, h->log.size()) };
        std::memcpy(buf, h->log.data(), n);
        buf[n] = '\0';
    }
    return GADFIT_B200_OK;
}

int gadfit_ref_fit(RefSolver* h, double lambda)
{
    auto& events { spdlog::oracle_events() };
    events.clear();
    spdlog::oracle_log().clear();
    h->solver->settings.verbosity |= gadfit::io::timings;
    const int status { guarded(h, [&] { h->solver->fit(lambda); }) };
    h->trace.clear();
    h->n_iterations = 0;
    h->timings = {};
    int i_timer {};
    // Replay the reference's log (lm_solver.cpp:454-501): a trial record is
    // followed by "Iteration: N" iff it was accepted; "Lambda increased"
    // marks the last rejected trial as the give-up.
    h->log.clear();
    for (const auto& line : spdlog::oracle_log()) {
        h->log += line;
        h->log += '\n';
    }
    for (const auto& e : events) {
        switch (e.kind) {
        case 0:
            h->trace.push_back(
              { e.a, e.b, e.c, h->n_iterations + 1, e.b < e.c ? 0 : 1 });
            break;
        case 1:
            h->n_iterations = static_cast<int>(e.a);
            break;
        case 2:
            if (!h->trace.empty()) {
                h->trace.back().decision = 2;
            }
            break;
        case 4:
            if (i_timer < 5) {
                h->timings.seconds[i_timer] = e.a;
                h->timings.calls[i_timer] = static_cast<int>(e.c);
                ++i_timer;
            }
            break;
        default:
            break;
        }
    }
    return status;
}

int gadfit_ref_get_par(RefSolver* h, int i_par, int i_dataset, double* out)
{
    return guarded(h,
                   [&] { *out = h->solver->getParValue(i_par, i_dataset); });
}

int gadfit_ref_get_value(RefSolver* h, double arg, int i_dataset, double* out)
{
    return guarded(h, [&] { *out = h->solver->getValue(arg, i_dataset); });
}

int gadfit_ref_chi2(RefSolver* h, double* out)
{
    return guarded(h, [&] { *out = h->solver->chi2(); });
}

int gadfit_ref_dof(RefSolver* h, int* out)
{
    *out = h->solver->degreesOfFreedom();
    return GADFIT_B200_OK;
}

int gadfit_ref_get_array(RefSolver* h,
                         int which,
                         double* buf,
                         size_t cap,
                         size_t* need)
{
    return guarded(h, [&] {
        const std::vector<double>* src { nullptr };
        switch (which) {
        case GADFIT_B200_JACOBIAN:
            src = &h->solver->getJacobian();
            break;
        case GADFIT_B200_JTJ:
            src = &h->solver->getJTJ();
            break;
        case GADFIT_B200_DTD:
            src = &h->solver->getDTD();
            break;
        case GADFIT_B200_LEFT_SIDE:
            src = &h->solver->getLeftSide();
            break;
        case GADFIT_B200_RIGHT_SIDE:
            src = &h->solver->getRightSide();
            break;
        case GADFIT_B200_RESIDUALS:
            src = &h->solver->getResiduals();
            break;
        case GADFIT_B200_INV_JTJ:
            src = &h->solver->getInvJTJ();
            break;
        default:
            throw std::invalid_argument { "array not exposed by the reference" };
        }
        *need = src->size();
        const size_t n { cap < src->size() ? cap : src->size() };
        if (buf != nullptr && n > 0) {
            std::memcpy(buf, src->data(), n * sizeof(double));
        }
    });
}

int gadfit_ref_get_trace(RefSolver* h,
                         gadfit_b200_trial* buf,
                         size_t cap,
                         size_t* n)
{
    *n = h->trace.size();
    const size_t m { cap < h->trace.size() ? cap : h->trace.size() };
    if (buf != nullptr && m > 0) {
        std::memcpy(buf, h->trace.data(), m * sizeof(gadfit_b200_trial));
    }
    return GADFIT_B200_OK;
}

int gadfit_ref_n_iterations(RefSolver* h, int* out)
{
    *out = h->n_iterations;
    return GADFIT_B200_OK;
}

int gadfit_ref_get_timings(RefSolver* h, gadfit_b200_timings* out)
{
    *out = h->timings;
    return GADFIT_B200_OK;
}

const char* gadfit_ref_last_error(RefSolver* h)
{
    return h->last_error.c_str();
}

int gadfit_ref_max_threads(void)
{
    return omp_get_max_threads();
}

} // extern "C"
