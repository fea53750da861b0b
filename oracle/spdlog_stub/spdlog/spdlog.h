// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.
//
// Print-only stand-in for the spdlog dependency of the reference
// (pinned there as gabime/spdlog@bae78f7b, c++/gadfit/CMakeLists.txt:50-53,
// fetched from the network, which this image does not have).  spdlog is used
// by the reference for printing only (lm_solver.cpp:141,457,485,498,591-593,
// 624,661-679); no arithmetic lives in it.  This stub lets the UNMODIFIED
// reference sources compile, and -- instead of formatting text -- records the
// numeric arguments of the messages that make up the reference's
// accept/reject trace:
//
//   debug("Current lambda: {}, new chi2 = {}, old chi2: {}", l, n, o)
//                                              lm_solver.cpp:457-460
//   info("Iteration: {}", i)                   lm_solver.cpp:591
//   info("Lambda increased {} times in a row") lm_solver.cpp:485
//   info("Iteration limit reached")            lm_solver.cpp:498
//
// The records are bit-exact doubles, not parsed text.  In addition every
// info-level message is formatted (std::vformat has the semantics of the fmt
// library spdlog uses) and kept as a line of text, so that the product's text
// output can be compared with the reference's own (oracle_log).
#pragma once

#include <format>
#include <string>
#include <string_view>
#include <type_traits>
#include <vector>

namespace spdlog {

namespace level {
enum level_enum { trace, debug, info, warn, err, critical, off };
}

struct OracleEvent
{
    // 0 = trial step (lambda, new chi2, old chi2), 1 = accepted iteration
    // (iteration number), 2 = gave up after lambda_incs rejections,
    // 3 = iteration limit reached, 4 = row of the timing table (wall seconds,
    // cpu seconds, number of calls)
    int kind;
    double a, b, c;
};

inline auto oracle_events() -> std::vector<OracleEvent>&
{
    static std::vector<OracleEvent> events {};
    return events;
}

inline auto oracle_log() -> std::vector<std::string>&
{
    static std::vector<std::string> lines {};
    return lines;
}

namespace detail {
template <class T>
auto as_double(const T& x) -> double
{
    if constexpr (std::is_arithmetic_v<T>) {
        return static_cast<double>(x);
    } else {
        return 0.0;
    }
}
} // namespace detail

template <class... Args>
auto debug(std::string_view fmt, const Args&... args) -> void
{
    if constexpr (sizeof...(Args) == 3) {
        if (fmt.starts_with("Current lambda")) {
            const double v[3] { detail::as_double(args)... };
            oracle_events().push_back({ 0, v[0], v[1], v[2] });
        }
    }
}

template <class... Args>
auto info(std::string_view fmt, const Args&... args) -> void
{
    oracle_log().push_back(std::vformat(fmt, std::make_format_args(args...)));
    if constexpr (sizeof...(Args) == 1) {
        const double v[1] { detail::as_double(args)... };
        if (fmt.starts_with("Iteration: ")) {
            oracle_events().push_back({ 1, v[0], 0.0, 0.0 });
        } else if (fmt.starts_with("Lambda increased")) {
            oracle_events().push_back({ 2, v[0], 0.0, 0.0 });
        }
    } else if constexpr (sizeof...(Args) == 5) {
        // Timing table row (lm_solver.cpp:661-667): term, wall, cpu, rel, calls
        const double v[5] { detail::as_double(args)... };
        oracle_events().push_back({ 4, v[1], v[2], v[4] });
    } else if constexpr (sizeof...(Args) == 0) {
        if (fmt.starts_with("Iteration limit reached")) {
            oracle_events().push_back({ 3, 0.0, 0.0, 0.0 });
        }
    }
}

template <class... Args>
auto warn(std::string_view, const Args&...) -> void
{}

template <class... Args>
auto error(std::string_view, const Args&...) -> void
{}

inline auto set_pattern(const std::string&) -> void {}
inline auto set_level(level::level_enum) -> void {}

} // namespace spdlog
