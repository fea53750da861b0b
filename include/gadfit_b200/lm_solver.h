// Header-only C++ mirror of the reference solver interface.
//
// Same class name, method names, defaults, settings struct and exception
// types as gadfit::LMsolver (reference c++/gadfit/lm_solver.h:84-227,
// exceptions.h:21-80), implemented as a thin layer over the C-ABI of
// include/gadfit_b200.h.  A program written against the reference,
//
//     gadfit::LMsolver solver { exponential };      // std::function body
//     solver.addDataset(x_data_1, y_data_1);
//     solver.setPar(0, 1.0, true, 0);
//     solver.setPar(1, 1.0, true);
//     solver.fit();
//
// changes in exactly one place: the fitting function is named instead of
// passed as a std::function, because it has to be compiled for the device
// (built in: include/gadfit_b200/fit_models.h; the user's own: a model
// plug-in, include/gadfit_b200/model_plugin.h):
//
//     gadfit::LMsolver solver { "exponential" };
//
//     gadfit::loadModels("libmy_models.so");        // user-defined functions
//     gadfit::LMsolver solver2 { "my_function" };
//
// Link with -lgadfit_b200.  Nothing here computes: every call forwards.
#pragma once

#include "../gadfit_b200.h"

#include <bitset>
#include <cassert>
#include <cstdio>
#include <exception>
#include <string>
#include <vector>

namespace gadfit {

// ---- exceptions, same names as the reference (exceptions.h:21-80) ---------
class GADfitException : public std::exception
{
protected:
    std::string message;

public:
    GADfitException() = default;
    explicit GADfitException(std::string msg) : message { std::move(msg) } {}
    [[nodiscard]] auto what() const noexcept -> const char* override
    {
        return message.c_str();
    }
};
#define GADFIT_B200_DEFINE_EXCEPTION(NAME)                                     \
    class NAME : public GADfitException                                        \
    {                                                                          \
    public:                                                                    \
        using GADfitException::GADfitException;                                \
    }
GADFIT_B200_DEFINE_EXCEPTION(UnknownOperation);
GADFIT_B200_DEFINE_EXCEPTION(LateAddDatasetCall);
GADFIT_B200_DEFINE_EXCEPTION(SetParInvalidIndex);
GADFIT_B200_DEFINE_EXCEPTION(UninitializedParameter);
GADFIT_B200_DEFINE_EXCEPTION(NegativeDegreesOfFreedom);
GADFIT_B200_DEFINE_EXCEPTION(NoFittingParameters);
GADFIT_B200_DEFINE_EXCEPTION(NoGlobalParameters);
GADFIT_B200_DEFINE_EXCEPTION(InsufficientIntegrationWorkspace);
// conditions without a reference counterpart
GADFIT_B200_DEFINE_EXCEPTION(UnknownModel);
GADFIT_B200_DEFINE_EXCEPTION(DeviceError);
GADFIT_B200_DEFINE_EXCEPTION(InvalidArgument);
GADFIT_B200_DEFINE_EXCEPTION(NotAvailable);
#undef GADFIT_B200_DEFINE_EXCEPTION

// Makes the fitting functions of a model plug-in available by name
// (gadfit_b200_load_models): the user's functor takes the place of the
// reference's fitSignature lambda, lm_solver.h:146-148.
inline auto loadModels(const std::string& plugin_path) -> void
{
    if (gadfit_b200_load_models(plugin_path.c_str()) != GADFIT_B200_OK) {
        throw InvalidArgument { plugin_path
                                + ": not loadable, not a model plug-in built against these "
                                  "headers, or a function name is already registered" };
    }
}

// ---- IO flags and loss functions (lm_solver.h:60-82) -----------------------
namespace io {
constexpr int n_io_flags { 8 };
using flag = std::bitset<n_io_flags>;
static constexpr flag all { 0b0000'0001 };
static constexpr flag delta1 { 0b0000'0010 };
static constexpr flag delta2 { 0b0000'0100 };
static constexpr flag timings { 0b0000'1000 };
static constexpr flag hide_local { 0b0001'0000 };
static constexpr flag hide_global { 0b0010'0000 };
static constexpr flag hide_all { 0b0100'0000 };
static constexpr flag final_only { 0b1000'0000 };
} // namespace io

enum class Loss
{
    linear,
    cauchy,
    huber,
};

class LMsolver
{
private:
    static auto printLine(const char* line, void*) -> void { std::puts(line); }
    static constexpr int global_dataset_idx { -1 };
    gadfit_b200_solver* handle { nullptr };
    std::vector<double> buffer; // backing store of the getter that ran last

    auto check(const int status) const -> void
    {
        if (status == GADFIT_B200_OK) {
            return;
        }
        const std::string msg { handle != nullptr ? gadfit_b200_last_error(handle) : "" };
        switch (status) {
        case GADFIT_B200_UNKNOWN_OPERATION:
            throw UnknownOperation { msg };
        case GADFIT_B200_LATE_ADD_DATASET_CALL:
            throw LateAddDatasetCall { msg };
        case GADFIT_B200_SET_PAR_INVALID_INDEX:
            throw SetParInvalidIndex { msg };
        case GADFIT_B200_UNINITIALIZED_PARAMETER:
            throw UninitializedParameter { msg };
        case GADFIT_B200_NEGATIVE_DEGREES_OF_FREEDOM:
            throw NegativeDegreesOfFreedom { msg };
        case GADFIT_B200_NO_FITTING_PARAMETERS:
            throw NoFittingParameters { msg };
        case GADFIT_B200_NO_GLOBAL_PARAMETERS:
            throw NoGlobalParameters { msg };
        case GADFIT_B200_INSUFFICIENT_INTEGRATION_WORKSPACE:
            throw InsufficientIntegrationWorkspace { msg };
        case GADFIT_B200_UNKNOWN_MODEL:
            throw UnknownModel { msg };
        case GADFIT_B200_CUDA_ERROR:
        case GADFIT_B200_NCCL_ERROR:
            throw DeviceError { msg };
        case GADFIT_B200_NOT_AVAILABLE:
            throw NotAvailable { msg };
        default:
            throw InvalidArgument { msg };
        }
    }

    auto array(const int which) -> const std::vector<double>&
    {
        size_t need { 0 };
        check(gadfit_b200_get_array(handle, which, nullptr, 0, &need));
        buffer.resize(need);
        if (need > 0) {
            check(gadfit_b200_get_array(handle, which, buffer.data(), need, &need));
        }
        return buffer;
    }

    auto pushSettings() -> void
    {
        gadfit_b200_settings s {};
        s.iteration_limit = settings.iteration_limit;
        s.lambda_incs = settings.lambda_incs;
        s.lambda_down = settings.lambda_down;
        s.lambda_up = settings.lambda_up;
        s.damp_max = settings.damp_max ? 1 : 0;
        s.DTD_min = settings.DTD_min.data();
        s.n_DTD_min = static_cast<int>(settings.DTD_min.size());
        s.acceleration_threshold = settings.acceleration_threshold;
        s.verbosity = static_cast<unsigned>(settings.verbosity.to_ulong());
        s.loss = static_cast<int>(settings.loss);
        s.n_threads = settings.n_threads;
        check(gadfit_b200_set_settings(handle, &s));
    }

public:
    LMsolver() = delete;
    // The one API delta: a registered device function instead of fitSignature
    explicit LMsolver(const std::string& model_name)
    {
        check(gadfit_b200_create(model_name.c_str(), &handle));
        // The reference logs its iteration results through spdlog by default
        // (lm_solver.cpp:465-467,587-654, pattern "%v" in its example): same
        // lines, to stdout, unless the caller installs another sink.
        try {
            setLog(&LMsolver::printLine, nullptr);
        } catch (...) {
            gadfit_b200_destroy(handle); // no destructor runs for a throwing constructor
            handle = nullptr;
            throw;
        }
    }
    // Where the text output of fit() goes (gadfit_b200_set_log); nullptr = nowhere
    auto setLog(gadfit_b200_log_fn sink, void* user) -> void
    {
        check(gadfit_b200_set_log(handle, sink, user));
    }
    LMsolver(const LMsolver&) = delete;
    auto operator=(const LMsolver&) -> LMsolver& = delete;
    LMsolver(LMsolver&& other) noexcept
      : handle { other.handle }, buffer { std::move(other.buffer) },
        settings { std::move(other.settings) }
    {
        other.handle = nullptr;
    }
    ~LMsolver()
    {
        if (handle != nullptr) {
            gadfit_b200_destroy(handle);
        }
    }

    // The data arrays are BORROWED, as in the reference -- but the reference
    // reads them at every fit() (it keeps spans, lm_solver.h:105-113), whereas
    // here they are copied to HBM once, at the first fit() / chi2() or at
    // upload().  After changing x / y / errors in place call invalidateData():
    // the next fit() copies them again.
    auto upload() -> void { check(gadfit_b200_upload(handle)); }
    auto invalidateData() -> void { check(gadfit_b200_invalidate_data(handle)); }
    // per-phase event timers (default on, as the reference's Timers); off for latency-bound fits
    auto setPhaseTimers(const bool on) -> void { check(gadfit_b200_set_phase_timers(handle, on ? 1 : 0)); }

    // lm_solver.h:158-167.  Arrays are borrowed, as in the reference.
    auto addDataset(const auto& x_data, const auto& y_data, const auto& errors) -> void
    {
        assert(x_data.size() == y_data.size());
        assert(x_data.size() == errors.size());
        addDataset(static_cast<int>(x_data.size()), x_data.data(), y_data.data(), errors.data());
    }
    auto addDataset(const auto& x_data, const auto& y_data) -> void
    {
        assert(x_data.size() == y_data.size());
        addDataset(static_cast<int>(x_data.size()), x_data.data(), y_data.data(), nullptr);
    }
    auto addDataset(const int n_datapoints,
                    const double* x_data,
                    const double* y_data,
                    const double* errors = nullptr) -> void
    {
        check(gadfit_b200_add_dataset(handle, n_datapoints, x_data, y_data, errors));
    }

    // lm_solver.h:171-180
    auto setPar(const int i_par,
                const double val,
                const bool active = false,
                const int i_dataset = global_dataset_idx,
                const std::string& parameter_name = "") -> void
    {
        check(gadfit_b200_set_par(handle,
                                  i_par,
                                  val,
                                  active ? 1 : 0,
                                  i_dataset,
                                  parameter_name.empty() ? nullptr : parameter_name.c_str()));
    }
    auto setPar(const int i_par, const double val, const bool active, const std::string& parameter_name)
      -> void
    {
        setPar(i_par, val, active, global_dataset_idx, parameter_name);
    }

    // lm_solver.h:181
    auto fit(double lambda = 10.0) -> void
    {
        pushSettings();
        check(gadfit_b200_fit(handle, lambda));
    }

    // lm_solver.h:182-195
    [[nodiscard]] auto getParValue(const int i_par, const int i_dataset = 0) const -> double
    {
        double out { 0.0 };
        check(gadfit_b200_get_par(handle, i_par, i_dataset, &out));
        return out;
    }
    [[nodiscard]] auto getValue(const double arg, const int i_dataset = 0) const -> double
    {
        double out { 0.0 };
        check(gadfit_b200_get_value(handle, arg, i_dataset, &out));
        return out;
    }
    [[nodiscard]] auto degreesOfFreedom() const -> int
    {
        int out { 0 };
        check(gadfit_b200_dof(handle, &out));
        return out;
    }
    [[nodiscard]] auto getJacobian() -> const std::vector<double>& { return array(GADFIT_B200_JACOBIAN); }
    [[nodiscard]] auto getJTJ() -> const std::vector<double>& { return array(GADFIT_B200_JTJ); }
    [[nodiscard]] auto getDTD() -> const std::vector<double>& { return array(GADFIT_B200_DTD); }
    [[nodiscard]] auto getLeftSide() -> const std::vector<double>& { return array(GADFIT_B200_LEFT_SIDE); }
    [[nodiscard]] auto getRightSide() -> const std::vector<double>& { return array(GADFIT_B200_RIGHT_SIDE); }
    [[nodiscard]] auto getResiduals() -> const std::vector<double>& { return array(GADFIT_B200_RESIDUALS); }
    [[nodiscard]] auto getInvJTJ() -> const std::vector<double>& { return array(GADFIT_B200_INV_JTJ); }
    // diag((J^T J)^-1) from the block-arrow factors: any n (getInvJTJ() is dense, n^2 <= 2^27)
    [[nodiscard]] auto getInvJTJDiagonal() -> const std::vector<double>&
    {
        return array(GADFIT_B200_INV_JTJ_DIAGONAL);
    }
    [[nodiscard]] auto chi2() -> double
    {
        double out { 0.0 };
        check(gadfit_b200_chi2(handle, &out));
        return out;
    }

    // Additions: the accept/reject trace the reference prints at debug level
    // (lm_solver.cpp:457-460) and the number of successful iterations
    [[nodiscard]] auto trace() const -> std::vector<gadfit_b200_trial>
    {
        size_t n { 0 };
        check(gadfit_b200_get_trace(handle, nullptr, 0, &n));
        std::vector<gadfit_b200_trial> out(n);
        check(gadfit_b200_get_trace(handle, out.data(), n, &n));
        return out;
    }
    [[nodiscard]] auto iterations() const -> int
    {
        int out { 0 };
        check(gadfit_b200_n_iterations(handle, &out));
        return out;
    }
    [[nodiscard]] auto native() const -> gadfit_b200_solver* { return handle; }

    // lm_solver.h:198-210, same names and defaults (:88-96)
    struct
    {
        int iteration_limit { 1000 };
        int lambda_incs { 3 };
        double lambda_down { 10.0 };
        double lambda_up { 10.0 };
        bool damp_max { true };
        std::vector<double> DTD_min { 0.0 };
        double acceleration_threshold { -1.0 };
        io::flag verbosity {};
        Loss loss { Loss::linear };
        int n_threads { 1 };
    } settings;
};

} // namespace gadfit
