// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
// this library; the product never does.
//
// CPU restatement of the hot path of the reference: LMsolver::fit and what it
// calls (reference c++/gadfit/lm_solver.cpp), each function citing the lines
// it follows.  Differences to the reference, all deliberate:
//   * derivatives come from the forward mode restated in oracle_ad.h (one
//     pass per active parameter) instead of the reverse tape;
//   * only the structural non-zeros of the Jacobian are stored (the
//     reference keeps a dense N x n array, lm_solver.cpp:411), so the
//     restatement also runs BASELINE configs 3-5 at sizes the reference
//     cannot;
//   * the normal equations are solved either DENSE, with the loops of the
//     reference's in-tree lapack_fallback.cpp (default for n <= 400), or by
//     the block-arrow Schur complement (any n).  Both are validated against
//     the unmodified reference (oracle/_ref) by tests/test_oracle.py.
// PARITY STATUS: pinned -- tests/test_oracle.py checks this restatement
// against every LM known-answer value of c++/tests/lm_solver.cpp (via
// tests/golden/) and, when oracle/_ref is present, against the compiled
// reference on the scaled-down configs.
//
// Exports the call shapes of include/gadfit_b200.h under gadfit_oracle_.

#include "gadfit_b200.h"
#include "gadfit_b200/fit_models.h"
#include "gadfit_b200/test_models.h"
#include "oracle_ad.h"

#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>
#include <numeric>
#include <omp.h>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

using gadfit_oracle::FVar;

struct StatusError : std::runtime_error
{
    int status;
    StatusError(int status, const std::string& msg) : std::runtime_error { msg }, status { status } {}
};

using Body = std::function<FVar(const FVar*, double)>;

struct OracleSolver
{
    // lm_solver.h:101-138
    Body body;
    int NP { 0 };
    struct Set
    {
        int n;
        const double* x;
        const double* y;
        const double* errors;
    };
    std::vector<Set> sets;
    std::vector<std::vector<double>> errors_dummy;
    std::vector<std::vector<FVar>> pars; // one parameter vector per data set
    std::vector<std::set<int>> active;   // Indices::active
    std::set<int> global;                // Indices::global
    bool set_par_called { false };
    gadfit_b200_settings settings {};
    std::vector<double> DTD_min { 0.0 };
    int mode { 0 }; // 0 auto, 1 dense, 2 block-arrow
    bool extended { false }; // long double accumulation of the shared block (rounding probe)

    // Indices
    long long n_active { 0 };
    long long n_datapoints { 0 };
    int dof { 0 };
    std::vector<std::vector<long long>> column; // Indices::jacobian
    int pg { 0 };

    // work arrays (structural non-zeros only)
    std::vector<double> jac;       // [N][max_row] compact rows
    int row_stride { 0 };
    std::vector<double> residuals; // [N]
    std::vector<double> omega;     // [N]
    std::vector<long long> set_begin;
    // block-arrow storage of JTJ: G (pg x pg), per set C (pg x pl), L (pl x pl)
    std::vector<double> G;
    std::vector<std::vector<double>> C, L;
    std::vector<double> DTD, right_side, delta1, delta2; // dense vectors of length n
    std::vector<double> dense_lhs, dense_chol;           // dense mode
    // block mode factorisation
    std::vector<std::vector<double>> R, Y;
    std::vector<double> S_chol;
    bool use_dense { true };
    double lambda_lhs { 0.0 };

    std::vector<gadfit_b200_trial> trace;
    int n_iterations { 0 };
    double seconds[5] {};
    int calls[5] {};
    std::string last_error;

    // ------------------------------------------------------------ set-up
    // lm_solver.cpp:30-56
    void addDataset(int n, const double* x, const double* y, const double* e)
    {
        if (set_par_called) {
            throw StatusError { GADFIT_B200_LATE_ADD_DATASET_CALL, "late addDataset" };
        }
        if (e == nullptr) {
            errors_dummy.emplace_back(n, 1.0);
            e = errors_dummy.back().data();
        }
        sets.push_back({ n, x, y, e });
        pars.emplace_back(pars.empty() ? std::vector<FVar> {} : pars.back());
        active.emplace_back();
    }

    FVar& par(int s, int i)
    {
        if (i >= static_cast<int>(pars[s].size())) {
            pars[s].resize(i + 1); // fit_function.cpp:17-23
        }
        return pars[s][i];
    }

    // lm_solver.cpp:58-105
    void setPar(int i_par, double val, bool is_active, int i_dataset)
    {
        if (i_dataset >= static_cast<int>(sets.size()) || sets.empty()) {
            throw StatusError { GADFIT_B200_SET_PAR_INVALID_INDEX, "invalid data set index" };
        }
        if (i_par < 0 || i_par >= NP || i_dataset < -1) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "parameter index out of range" };
        }
        set_par_called = true;
        if (i_dataset == -1) {
            if (is_active) {
                global.insert(i_par);
            } else {
                global.erase(i_par);
            }
            for (int s = 0; s < static_cast<int>(sets.size()); ++s) {
                par(s, i_par) = gadfit_oracle::passive(val);
                if (is_active) {
                    active[s].insert(i_par);
                } else {
                    active[s].erase(i_par);
                }
            }
        } else {
            global.erase(i_par);
            par(i_dataset, i_par) = gadfit_oracle::passive(val);
            if (is_active) {
                active[i_dataset].insert(i_par);
            } else {
                active[i_dataset].erase(i_par);
            }
        }
    }

    // lm_solver.cpp:119-190
    void prepareIndexing()
    {
        const int S { static_cast<int>(sets.size()) };
        if (S == 0) {
            throw StatusError { GADFIT_B200_NO_FITTING_PARAMETERS, "no data sets" };
        }
        n_active = 0;
        for (const auto& a : active) {
            n_active += static_cast<long long>(a.size());
        }
        n_active -= static_cast<long long>(S - 1) * static_cast<long long>(global.size());
        if (n_active > 0 && S > 1 && global.empty()) {
            throw StatusError { GADFIT_B200_NO_GLOBAL_PARAMETERS, "no global parameters" };
        }
        n_datapoints = 0;
        for (const auto& set : sets) {
            n_datapoints += set.n;
        }
        const long long degrees { n_datapoints - n_active };
        if (degrees < 0) {
            throw StatusError { GADFIT_B200_NEGATIVE_DEGREES_OF_FREEDOM, "negative dof" };
        }
        dof = degrees == 0 ? 1 : static_cast<int>(std::min<long long>(degrees, INT32_MAX));
        for (const auto& p : pars) {
            if (p.size() != pars.front().size()) {
                throw StatusError { GADFIT_B200_UNINITIALIZED_PARAMETER, "uninitialized parameter" };
            }
        }
        if (static_cast<int>(pars.front().size()) != NP) {
            throw StatusError { GADFIT_B200_UNINITIALIZED_PARAMETER, "uninitialized parameter" };
        }
        pg = static_cast<int>(global.size());
        column.assign(S, {});
        long long next_idx { 0 };
        for (int s = 0; s < S; ++s) {
            auto& cur { column[s] };
            cur.resize(active[s].size());
            long long next_global_pos { 0 };
            auto it { active[s].cbegin() };
            for (size_t i = 0; i < cur.size(); ++i, ++it) {
                if (global.count(*it) != 0) {
                    cur[i] = next_global_pos++;
                } else if (s == 0) {
                    cur[i] = pg + next_idx++;
                } else {
                    cur[i] = next_idx++;
                }
            }
            if (s == 0) {
                next_idx = static_cast<long long>(cur.size());
            }
        }
        if (n_active == 0) {
            throw StatusError { GADFIT_B200_NO_FITTING_PARAMETERS, "no fitting parameters" };
        }
    }

    // ------------------------------------------------------------- sweeps
    double eval(const std::vector<FVar>& p, double x) const { return body(p.data(), x).val; }

    // lm_solver.cpp:513-529
    double chi2()
    {
        const double t0 { omp_get_wtime() };
        if (extended) {
            long double sum_x { 0.0L };
            for (size_t s = 0; s < sets.size(); ++s) {
                const auto& set { sets[s] };
                const auto& p { pars[s] };
#pragma omp parallel for reduction(+ : sum_x) num_threads(threads())
                for (int i = 0; i < set.n; ++i) {
                    const double diff { (set.y[i] - eval(p, set.x[i])) / set.errors[i] };
                    sum_x += static_cast<long double>(diff) * diff;
                }
            }
            seconds[1] += omp_get_wtime() - t0;
            ++calls[1];
            return static_cast<double>(sum_x);
        }
        double sum { 0.0 };
        for (size_t s = 0; s < sets.size(); ++s) {
            const auto& set { sets[s] };
            const auto& p { pars[s] };
#pragma omp parallel for reduction(+ : sum) num_threads(threads())
            for (int i = 0; i < set.n; ++i) {
                const double diff { (set.y[i] - eval(p, set.x[i])) / set.errors[i] };
                sum += diff * diff;
            }
        }
        seconds[1] += omp_get_wtime() - t0;
        ++calls[1];
        return sum;
    }

    int threads() const { return settings.n_threads > 0 ? settings.n_threads : 1; }

    // lm_solver.cpp:257-284
    static double loss(int type, double res)
    {
        switch (type) {
        case GADFIT_B200_LOSS_CAUCHY:
            return std::sqrt(1.0 / (1.0 + res * res));
        case GADFIT_B200_LOSS_HUBER:
            return res * res > 1.0 ? std::sqrt(1.0 / std::abs(res)) : 1.0;
        case GADFIT_B200_LOSS_LINEAR:
            return 1.0;
        default:
            return 0.0;
        }
    }

    // lm_solver.cpp:286-346: residuals, Jacobian rows, JTJ, DTD, J^T r
    void computeLeftAndRightHandSide()
    {
        double t0 { omp_get_wtime() };
        const int S { static_cast<int>(sets.size()) };
        for (int s = 0; s < S; ++s) {
            const auto& set { sets[s] };
            const std::vector<int> idx(active[s].cbegin(), active[s].cend());
            const int ps { static_cast<int>(idx.size()) };
#pragma omp parallel num_threads(threads())
            {
                std::vector<FVar> p { pars[s] };
#pragma omp for
                for (int i = 0; i < set.n; ++i) {
                    const long long row { set_begin[s] + i };
                    // lm_solver.cpp:303-310
                    const double res { (set.y[i] - body(p.data(), set.x[i]).val) / set.errors[i] };
                    const double drho_sqrt { loss(settings.loss, res) };
                    residuals[row] = drho_sqrt * res;
                    for (int a = 0; a < ps; ++a) {
                        // activateParForward(idx, 1.0), fit_function.cpp:45-51
                        p[idx[a]].d = 1.0;
                        p[idx[a]].dd = 0.0;
                        p[idx[a]].on = true;
                        const double grad { body(p.data(), set.x[i]).d };
                        p[idx[a]].on = false;
                        // lm_solver.cpp:315-317
                        jac[row * row_stride + a] = drho_sqrt * grad / set.errors[i];
                    }
                }
            }
        }
        seconds[0] += omp_get_wtime() - t0;
        ++calls[0];
        t0 = omp_get_wtime();
        // dsyrk, lapack_fallback.cpp:25-41 (sum over points in order) on the
        // structural blocks, and dgemv, lapack_fallback.cpp:43-57
        std::fill(G.begin(), G.end(), 0.0);
        std::fill(right_side.begin(), right_side.end(), 0.0);
        // Rounding probe (mode bit 4): the shared block and J^T r accumulated
        // in long double, to tell which of two double-precision summation
        // orders is closer to the exact sums (tools/full_size_parity.py)
        std::vector<long double> Gx(extended ? G.size() : 0, 0.0L);
        std::vector<long double> rx(extended ? static_cast<size_t>(pg) : 0, 0.0L);
        for (int s = 0; s < S; ++s) {
            const auto& cols { column[s] };
            const int ps { static_cast<int>(cols.size()) };
            const int pl { ps - pg_of(s) };
            std::fill(C[s].begin(), C[s].end(), 0.0);
            std::fill(L[s].begin(), L[s].end(), 0.0);
            std::vector<int> gpos, lpos; // positions within the compact row
            for (int a = 0; a < ps; ++a) {
                (cols[a] < pg ? gpos : lpos).push_back(a);
            }
            for (int i = 0; i < sets[s].n; ++i) {
                const double* row { &jac[(set_begin[s] + i) * row_stride] };
                const double r { residuals[set_begin[s] + i] };
                for (size_t a = 0; a < gpos.size(); ++a) {
                    const long long ca { cols[gpos[a]] };
                    for (size_t b = 0; b < gpos.size(); ++b) {
                        if (extended) {
                            Gx[ca * pg + cols[gpos[b]]] +=
                              static_cast<long double>(row[gpos[a]]) * row[gpos[b]];
                        } else {
                            G[ca * pg + cols[gpos[b]]] += row[gpos[a]] * row[gpos[b]];
                        }
                    }
                    for (int b = 0; b < pl; ++b) {
                        C[s][ca * pl + b] += row[gpos[a]] * row[lpos[b]];
                    }
                    if (extended) {
                        rx[ca] += static_cast<long double>(row[gpos[a]]) * r;
                    } else {
                        right_side[ca] += row[gpos[a]] * r;
                    }
                }
                for (int a = 0; a < pl; ++a) {
                    for (int b = 0; b < pl; ++b) {
                        L[s][a * pl + b] += row[lpos[a]] * row[lpos[b]];
                    }
                    right_side[cols[lpos[a]]] += row[lpos[a]] * r;
                }
            }
        }
        if (extended) {
            for (size_t i = 0; i < G.size(); ++i) {
                G[i] = static_cast<double>(Gx[i]);
            }
            for (int a = 0; a < pg; ++a) {
                right_side[a] = static_cast<double>(rx[a]);
            }
        }
        // lm_solver.cpp:326-329
        for (long long c = 0; c < n_active; ++c) {
            const double jtj { diag(c) };
            DTD[c] = settings.damp_max != 0 ? std::max(DTD[c], jtj) : jtj;
        }
        seconds[2] += omp_get_wtime() - t0;
        ++calls[2];
    }

    int pg_of(int s) const
    {
        int n { 0 };
        for (const int k : active[s]) {
            n += static_cast<int>(global.count(k));
        }
        return n;
    }

    std::vector<long long> local_begin; // first local column of each set
    double diag(long long c) const
    {
        if (c < pg) {
            return G[c * pg + c];
        }
        const auto it { std::upper_bound(local_begin.cbegin(), local_begin.cend(), c) };
        const int s { static_cast<int>(it - local_begin.cbegin()) - 1 };
        const int pl { static_cast<int>(column[s].size()) - pg_of(s) };
        const long long a { c - local_begin[s] };
        return L[s][a * pl + a];
    }

    // Dense JTJ (the array the reference keeps), from the blocks
    std::vector<double> denseJTJ() const
    {
        const long long n { n_active };
        std::vector<double> A(n * n, 0.0);
        for (int a = 0; a < pg; ++a) {
            for (int b = 0; b < pg; ++b) {
                A[a * n + b] = G[a * pg + b];
            }
        }
        for (size_t s = 0; s < sets.size(); ++s) {
            const int pl { static_cast<int>(column[s].size()) - pg_of(static_cast<int>(s)) };
            const long long o { local_begin[s] };
            for (int a = 0; a < pg; ++a) {
                for (int b = 0; b < pl; ++b) {
                    A[a * n + o + b] = C[s][a * pl + b];
                    A[(o + b) * n + a] = C[s][a * pl + b];
                }
            }
            for (int a = 0; a < pl; ++a) {
                for (int b = 0; b < pl; ++b) {
                    A[(o + a) * n + o + b] = L[s][a * pl + b];
                }
            }
        }
        return A;
    }

    // lapack_fallback.cpp:59-81 (dpptrf) on a dense n x n
    static void cholesky(int n, const std::vector<double>& A, std::vector<double>& Lo)
    {
        Lo.assign(A.size(), 0.0);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < i; ++j) {
                double v { A[i * n + j] };
                for (int k = 0; k < j; ++k) {
                    v -= Lo[i * n + k] * Lo[j * n + k];
                }
                Lo[i * n + j] = v / Lo[j * n + j];
            }
            double v { A[i * n + i] };
            for (int k = 0; k < i; ++k) {
                v -= Lo[i * n + k] * Lo[i * n + k];
            }
            Lo[i * n + i] = std::sqrt(v);
        }
    }
    // lapack_fallback.cpp:83-105 (dpptrs)
    static void cholSolve(int n, const std::vector<double>& Lo, double* B)
    {
        std::vector<double> y(n);
        for (int i = 0; i < n; ++i) {
            y[i] = B[i];
            for (int j = 0; j < i; ++j) {
                y[i] -= Lo[i * n + j] * y[j];
            }
            y[i] /= Lo[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            B[i] = y[i];
            for (int j = n - 1; j > i; --j) {
                B[i] -= Lo[j * n + i] * B[j];
            }
            B[i] /= Lo[i * n + i];
        }
    }

    // left_side = JTJ + lambda * DTD and its factorisation
    // (lm_solver.cpp:330-332,351-353)
    void factorise(double lambda)
    {
        lambda_lhs = lambda;
        const int S { static_cast<int>(sets.size()) };
        if (use_dense) {
            dense_lhs = denseJTJ();
            const long long n { n_active };
            for (long long c = 0; c < n; ++c) {
                dense_lhs[c * n + c] += lambda * DTD[c];
            }
            cholesky(static_cast<int>(n), dense_lhs, dense_chol);
            return;
        }
        // Block-arrow: R_s R_s' = L_s + lambda D_s, Y_s = R_s^-1 C_s',
        // S = G + lambda D_g - sum_s Y_s' Y_s
        std::vector<double> Sm(G);
        for (int a = 0; a < pg; ++a) {
            Sm[a * pg + a] += lambda * DTD[a];
        }
        for (int s = 0; s < S; ++s) {
            const int pl { static_cast<int>(column[s].size()) - pg_of(s) };
            std::vector<double> A(L[s]);
            for (int a = 0; a < pl; ++a) {
                A[a * pl + a] += lambda * DTD[local_begin[s] + a];
            }
            cholesky(pl, A, R[s]);
            Y[s].assign(static_cast<size_t>(pl) * pg, 0.0);
            for (int g = 0; g < pg; ++g) {
                for (int i = 0; i < pl; ++i) {
                    double v { C[s][g * pl + i] };
                    for (int j = 0; j < i; ++j) {
                        v -= R[s][i * pl + j] * Y[s][j * pg + g];
                    }
                    Y[s][i * pg + g] = v / R[s][i * pl + i];
                }
            }
            for (int a = 0; a < pg; ++a) {
                for (int b = 0; b < pg; ++b) {
                    double v { 0.0 };
                    for (int i = 0; i < pl; ++i) {
                        v += Y[s][i * pg + a] * Y[s][i * pg + b];
                    }
                    Sm[a * pg + b] -= v;
                }
            }
        }
        cholesky(pg, Sm, S_chol);
    }

    // delta = (JTJ + lambda DTD)^-1 rhs with the current factorisation
    // (lm_solver.cpp:354,384)
    void solve(std::vector<double>& rhs_to_delta)
    {
        if (use_dense) {
            cholSolve(static_cast<int>(n_active), dense_chol, rhs_to_delta.data());
            return;
        }
        const int S { static_cast<int>(sets.size()) };
        std::vector<double> rg(rhs_to_delta.begin(), rhs_to_delta.begin() + pg);
        std::vector<std::vector<double>> z(S);
        for (int s = 0; s < S; ++s) {
            const int pl { static_cast<int>(column[s].size()) - pg_of(s) };
            z[s].resize(pl);
            for (int i = 0; i < pl; ++i) {
                double v { rhs_to_delta[local_begin[s] + i] };
                for (int j = 0; j < i; ++j) {
                    v -= R[s][i * pl + j] * z[s][j];
                }
                z[s][i] = v / R[s][i * pl + i];
            }
            for (int a = 0; a < pg; ++a) {
                for (int i = 0; i < pl; ++i) {
                    rg[a] -= Y[s][i * pg + a] * z[s][i];
                }
            }
        }
        cholSolve(pg, S_chol, rg.data());
        std::copy(rg.begin(), rg.end(), rhs_to_delta.begin());
        for (int s = 0; s < S; ++s) {
            const int pl { static_cast<int>(column[s].size()) - pg_of(s) };
            std::vector<double> d(z[s]);
            for (int i = 0; i < pl; ++i) {
                for (int a = 0; a < pg; ++a) {
                    d[i] -= Y[s][i * pg + a] * rg[a];
                }
            }
            for (int i = pl - 1; i >= 0; --i) {
                for (int j = pl - 1; j > i; --j) {
                    d[i] -= R[s][j * pl + i] * d[j];
                }
                d[i] /= R[s][i * pl + i];
                rhs_to_delta[local_begin[s] + i] = d[i];
            }
        }
    }

    // lm_solver.cpp:348-399
    void computeDeltas(double lambda)
    {
        double t0 { omp_get_wtime() };
        factorise(lambda);
        delta1 = right_side;
        solve(delta1);
        seconds[2] += omp_get_wtime() - t0;
        ++calls[2];
        if (!(settings.acceleration_threshold > 0.0)) {
            return;
        }
        t0 = omp_get_wtime();
        const int S { static_cast<int>(sets.size()) };
        for (int s = 0; s < S; ++s) {
            const auto& set { sets[s] };
            std::vector<FVar> p { pars[s] };
            size_t a { 0 };
            for (const int k : active[s]) {
                // lm_solver.cpp:366-370
                p[k].d = delta1[column[s][a++]];
                p[k].dd = 0.0;
                p[k].on = true;
            }
#pragma omp parallel for num_threads(threads())
            for (int i = 0; i < set.n; ++i) {
                omega[set_begin[s] + i] = body(p.data(), set.x[i]).dd / set.errors[i];
            }
        }
        seconds[3] += omp_get_wtime() - t0;
        ++calls[3];
        t0 = omp_get_wtime();
        // dgemv('t', J, omega), lm_solver.cpp:383
        std::fill(delta2.begin(), delta2.end(), 0.0);
        for (int s = 0; s < S; ++s) {
            const auto& cols { column[s] };
            for (int i = 0; i < sets[s].n; ++i) {
                const long long row { set_begin[s] + i };
                for (size_t a = 0; a < cols.size(); ++a) {
                    delta2[cols[a]] += jac[row * row_stride + a] * omega[row];
                }
            }
        }
        solve(delta2);
        // lm_solver.cpp:388-397
        double D_delta2 { 0.0 };
        double D_delta1 { 0.0 };
        for (long long c = 0; c < n_active; ++c) {
            D_delta2 += delta2[c] * (DTD[c] * delta2[c]);
        }
        for (long long c = 0; c < n_active; ++c) {
            D_delta1 += delta1[c] * (DTD[c] * delta1[c]);
        }
        const double acc_ratio { std::sqrt(D_delta2 / D_delta1) };
        if (acc_ratio > settings.acceleration_threshold) {
            std::fill(delta2.begin(), delta2.end(), 0.0);
        }
        seconds[2] += omp_get_wtime() - t0;
        ++calls[2];
    }

    // lm_solver.cpp:230-253
    void updateParameters()
    {
        for (size_t s = 0; s < sets.size(); ++s) {
            size_t a { 0 };
            for (const int k : active[s]) {
                pars[s][k].val += delta1[column[s][a++]];
            }
        }
        if (settings.acceleration_threshold > 0.0) {
            for (size_t s = 0; s < sets.size(); ++s) {
                size_t a { 0 };
                for (const int k : active[s]) {
                    pars[s][k].val -= 0.5 * delta2[column[s][a++]];
                }
            }
        }
    }

    // work arrays and initial damping matrix (lm_solver.cpp:411-428)
    void prepareWorkArrays()
    {
        const int S { static_cast<int>(sets.size()) };
        std::fill(std::begin(seconds), std::end(seconds), 0.0);
        std::fill(std::begin(calls), std::end(calls), 0);
        use_dense = mode == 1 || (mode == 0 && n_active <= 400);
        set_begin.assign(S + 1, 0);
        local_begin.assign(S, 0);
        row_stride = 1;
        for (int s = 0; s < S; ++s) {
            set_begin[s + 1] = set_begin[s] + sets[s].n;
            row_stride = std::max(row_stride, static_cast<int>(column[s].size()));
            const int pgs { pg_of(s) };
            local_begin[s] = s == 0 ? pg
                                    : local_begin[s - 1]
                                        + static_cast<long long>(column[s - 1].size())
                                        - pg_of(s - 1);
            (void)pgs;
        }
        jac.assign(static_cast<size_t>(n_datapoints) * row_stride, 0.0);
        residuals.assign(n_datapoints, 0.0);
        omega.assign(n_datapoints, 0.0);
        G.assign(static_cast<size_t>(pg) * pg, 0.0);
        C.assign(S, {});
        L.assign(S, {});
        R.assign(S, {});
        Y.assign(S, {});
        for (int s = 0; s < S; ++s) {
            const int pl { static_cast<int>(column[s].size()) - pg_of(s) };
            C[s].assign(static_cast<size_t>(pg) * pl, 0.0);
            L[s].assign(static_cast<size_t>(pl) * pl, 0.0);
        }
        DTD.assign(n_active, 0.0);
        right_side.assign(n_active, 0.0);
        delta1.assign(n_active, 0.0);
        delta2.assign(n_active, 0.0);
        // lm_solver.cpp:424-428
        if (DTD_min.size() > 1) {
            if (static_cast<long long>(DTD_min.size()) < n_active) {
                throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "DTD_min too short" };
            }
            for (long long c = 0; c < n_active; ++c) {
                DTD[c] = DTD_min[c];
            }
        }
    }

    // lm_solver.cpp:401-511
    void fit(double lambda)
    {
        if (gadf.enabled != 0) {
            fitGadf(lambda);
            return;
        }
        prepareIndexing();
        prepareWorkArrays();
        lambda_exhausted = false;
        std::vector<std::vector<FVar>> old_parameters;
        trace.clear();
        const double t_main { omp_get_wtime() };
        double old_chi2 { chi2() };
        int i_iteration { 0 };
        bool finished { settings.iteration_limit == 0 };
        while (!finished) {
            ++i_iteration;
            computeLeftAndRightHandSide();
            computeDeltas(lambda);
            old_parameters = pars; // saveParameters
            updateParameters();
            for (int i_lambda = 0; i_lambda <= settings.lambda_incs; ++i_lambda) {
                const double new_chi2 { chi2() };
                trace.push_back({ lambda, new_chi2, old_chi2, i_iteration, 0 });
                if (new_chi2 < old_chi2) {
                    lambda /= settings.lambda_down;
                    old_chi2 = new_chi2;
                    break;
                }
                if (i_lambda < settings.lambda_incs) {
                    trace.back().decision = 1;
                    lambda *= settings.lambda_up;
                    pars = old_parameters; // revertParameters
                    computeDeltas(lambda);
                    updateParameters();
                } else {
                    trace.back().decision = 2;
                    pars = old_parameters;
                    --i_iteration;
                    finished = true;
                    lambda_exhausted = true;
                }
            }
            if (i_iteration == settings.iteration_limit) {
                finished = true;
            }
        }
        seconds[4] = omp_get_wtime() - t_main;
        calls[4] = 1;
        n_iterations = i_iteration;
        stop_reason = lambda_exhausted ? GADFIT_B200_STOP_LAMBDA_INCS
                                       : GADFIT_B200_STOP_ITERATION_LIMIT;
    }

    // ---- the Fortran driver, fortran/gadfit/gadfit.F90:678-915 ------------
    gadfit_b200_gadf_options gadf { 0, 0, 0, 0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0 };
    double umnigh_a { 0.5 }; // gadfit.F90:515 (a saved variable)
    int stop_reason { 0 };
    bool lambda_exhausted { false };

    double dotD(const std::vector<double>& a, const std::vector<double>& b) const
    {
        double sum { 0.0 };
        for (long long c = 0; c < n_active; ++c) {
            sum += a[c] * (DTD[c] * b[c]);
        }
        return sum;
    }

    void updateDelta1Only()
    {
        for (size_t s = 0; s < sets.size(); ++s) {
            size_t a { 0 };
            for (const int k : active[s]) {
                pars[s][k].val += delta1[column[s][a++]];
            }
        }
    }

    void fitGadf(double lambda)
    {
        if (settings.loss != 0) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "gadf_fit has no loss functions" };
        }
        prepareIndexing();
        prepareWorkArrays();
        const int S { static_cast<int>(sets.size()) };
        const double umnigh_m { std::exp(-0.2) }; // :516
        std::vector<double> old_delta1(n_active, 0.0);
        std::vector<std::vector<FVar>> old_parameters { pars };
        trace.clear();
        const double t_main { omp_get_wtime() };
        double old_chi2 { chi2() };
        double old_old_chi2 { old_chi2 };
        int iterations { 0 };
        stop_reason = GADFIT_B200_STOP_ITERATION_LIMIT;
        bool done { false };
        while (!done) {
            // STEP 1-3 (:680-743): Jacobian, JTJ, DTD rule, delta1, omega, delta2
            computeLeftAndRightHandSide();
            computeDeltas(lambda);
            updateParameters(); // val + delta1 + 0.5 delta2_F, delta2_F = -delta2 (:744-750)
            // STEP 4 (:751-818)
            double new_chi2 { 0.0 };
            double beta { 0.0 };
            for (int i = 1; i <= settings.lambda_incs + 1; ++i) {
                new_chi2 = chi2();
                trace.push_back({ lambda, new_chi2, old_chi2, iterations + 1, 0 });
                beta = iterations == 0 ? 0.0
                                       : dotD(delta1, old_delta1) / std::sqrt(dotD(delta1, delta1))
                                           / std::sqrt(dotD(old_delta1, old_delta1));
                double factor { 1.0 };
                for (int k = 0; k < gadf.uphill; ++k) {
                    factor *= 1.0 - beta;
                }
                if (factor * new_chi2 < old_chi2) {
                    if (gadf.nielsen != 0) {
                        // rho = (old - new) / 2 / delta1^T (JTJ + lambda DTD) delta1, :766-769
                        const auto A { denseJTJ() };
                        double quad { 0.0 };
                        for (long long r = 0; r < n_active; ++r) {
                            double row { 0.0 };
                            for (long long c = 0; c < n_active; ++c) {
                                const double d { r == c ? lambda * DTD[c] : 0.0 };
                                row += (A[r * n_active + c] + d) * delta1[c];
                            }
                            quad += delta1[r] * row;
                        }
                        const double rho { (old_chi2 - new_chi2) / 2 / quad };
                        const double t { 2 * rho - 1 };
                        lambda *= std::max(1 / settings.lambda_down, 1 - t * t * t);
                    }
                    if (gadf.umnigh != 0) { // :771-783
                        if (new_chi2 < old_chi2 && beta >= 0.0) {
                            umnigh_a = umnigh_a * umnigh_m + 1.0 - umnigh_m;
                            const double t { 1.0 - std::abs(2.0 * umnigh_a - 1.0) };
                            lambda *= std::min(1.0, std::max(1e-2, t * t));
                        } else {
                            umnigh_a = umnigh_a * umnigh_m + (1.0 - umnigh_m) / 2.0;
                            if (new_chi2 >= old_chi2) {
                                lambda /= std::min(
                                  10.0, std::max(1.0, 1.0 - std::abs(2.0 * umnigh_a - 1.0)));
                            }
                        }
                    }
                    if (gadf.nielsen == 0 && gadf.umnigh == 0) {
                        lambda /= settings.lambda_down;
                    }
                    break;
                }
                if (i <= settings.lambda_incs) { // :790-808
                    trace.back().decision = 1;
                    if (gadf.umnigh != 0) {
                        umnigh_a *= umnigh_m;
                        const double t { 1.0 - std::abs(2.0 * umnigh_a - 1.0) };
                        lambda *= beta < 0.0 ? std::min(1.0, std::max(1e-2, t * t))
                                             : std::min(1.0, std::max(0.1, t));
                    } else {
                        lambda = settings.lambda_up * lambda;
                    }
                    pars = old_parameters;
                    factorise(lambda);
                    delta1 = right_side;
                    solve(delta1);
                    updateDelta1Only();
                } else { // :809-816
                    trace.back().decision = 2;
                    pars = old_parameters;
                    stop_reason = GADFIT_B200_STOP_LAMBDA_INCS;
                    done = true;
                }
            }
            if (done) {
                break;
            }
            // :819-824
            old_parameters = pars;
            old_delta1 = delta1;
            old_old_chi2 = old_chi2;
            old_chi2 = std::min(old_chi2, new_chi2);
            ++iterations;
            // STEP 5 (:832-909)
            if (gadf.chi2_abs >= 0.0 && old_chi2 / dof < gadf.chi2_abs) {
                stop_reason = GADFIT_B200_STOP_CHI2_ABS;
                break;
            }
            if (gadf.chi2_rel >= 0.0 && (old_old_chi2 - old_chi2) / old_chi2 < gadf.chi2_rel) {
                stop_reason = GADFIT_B200_STOP_CHI2_REL;
                break;
            }
            if (gadf.grad_chi2 >= 0.0) {
                double norm2 { 0.0 };
                for (long long c = 0; c < n_active; ++c) {
                    norm2 += right_side[c] * right_side[c];
                }
                if (2 * std::sqrt(norm2) < gadf.grad_chi2) {
                    stop_reason = GADFIT_B200_STOP_GRAD_CHI2;
                    break;
                }
            }
            if (gadf.cos_phi >= 0.0) {
                double res_Jd { 0.0 };
                double res_res { 0.0 };
                double Jd_Jd { 0.0 };
                for (int s = 0; s < S; ++s) {
                    const auto& cols { column[s] };
                    for (int p = 0; p < sets[s].n; ++p) {
                        const long long row { set_begin[s] + p };
                        double Jd { 0.0 };
                        for (size_t a = 0; a < cols.size(); ++a) {
                            Jd += jac[row * row_stride + a] * delta1[cols[a]];
                        }
                        res_Jd += residuals[row] * Jd;
                        res_res += residuals[row] * residuals[row];
                        Jd_Jd += Jd * Jd;
                    }
                }
                if (std::abs(res_Jd) / std::sqrt(res_res) / std::sqrt(Jd_Jd) < gadf.cos_phi) {
                    stop_reason = GADFIT_B200_STOP_COS_PHI;
                    break;
                }
            }
            if (gadf.rel_error >= 0.0) {
                bool any_large { false };
                for (int s = 0; s < S; ++s) {
                    size_t a { 0 };
                    for (const int k : active[s]) {
                        any_large = any_large
                                    || std::abs(delta1[column[s][a++]] / pars[s][k].val)
                                         > gadf.rel_error;
                    }
                }
                if (!any_large) {
                    stop_reason = GADFIT_B200_STOP_REL_ERROR;
                    break;
                }
            }
            if (gadf.rel_error_global >= 0.0) {
                bool any_large { false };
                size_t a { 0 };
                for (const int k : active[0]) {
                    if (global.count(k) != 0) {
                        any_large = any_large
                                    || std::abs(delta1[column[0][a]] / pars[0][k].val)
                                         > gadf.rel_error_global;
                    }
                    ++a;
                }
                if (!any_large) {
                    stop_reason = GADFIT_B200_STOP_REL_ERROR_GLOBAL;
                    break;
                }
            }
            if (iterations >= settings.iteration_limit) {
                stop_reason = GADFIT_B200_STOP_ITERATION_LIMIT;
                break;
            }
        }
        seconds[4] = omp_get_wtime() - t_main;
        calls[4] = 1;
        n_iterations = iterations;
    }
};

template <class F>
int guarded(OracleSolver* h, F&& f)
{
    try {
        f();
        return GADFIT_B200_OK;
    } catch (const StatusError& e) {
        h->last_error = e.what();
        return e.status;
    } catch (const std::exception& e) {
        h->last_error = e.what();
        return GADFIT_B200_INVALID_ARGUMENT;
    }
}

} // namespace

extern "C" {

int gadfit_oracle_create(const char* model_name, OracleSolver** out)
{
    auto solver { std::make_unique<OracleSolver>() };
    const auto match { [&]<class Model>() {
        if (std::strcmp(model_name, Model::name) == 0) {
            solver->NP = Model::n_pars;
            solver->body = [](const FVar* p, double x) {
                return FVar { gadfit_b200::evaluate(Model {}, gadfit_oracle::Ctx { p, Model::n_pars }, x) };
            };
        }
    } };
    gadfit_b200::models::for_each_model(match);
    gadfit_b200::models::for_each_test_model(match);
    if (!solver->body) {
        return GADFIT_B200_UNKNOWN_MODEL;
    }
    solver->settings = { 1000, 3, 10.0, 10.0, 1, nullptr, 0, -1.0, 0U, 0, 1 };
    *out = solver.release();
    return GADFIT_B200_OK;
}

int gadfit_oracle_destroy(OracleSolver* h)
{
    delete h;
    return GADFIT_B200_OK;
}

int gadfit_oracle_add_dataset(OracleSolver* h, int n, const double* x, const double* y, const double* e)
{
    return guarded(h, [&] { h->addDataset(n, x, y, e); });
}

int gadfit_oracle_set_par(OracleSolver* h, int i_par, double val, int active, int i_dataset, const char*)
{
    return guarded(h, [&] { h->setPar(i_par, val, active != 0, i_dataset); });
}

int gadfit_oracle_default_settings(gadfit_b200_settings* s)
{
    *s = { 1000, 3, 10.0, 10.0, 1, nullptr, 0, -1.0, 0U, 0, 1 };
    return GADFIT_B200_OK;
}

int gadfit_oracle_set_settings(OracleSolver* h, const gadfit_b200_settings* s)
{
    h->settings = *s;
    if (s->DTD_min != nullptr && s->n_DTD_min > 0) {
        h->DTD_min.assign(s->DTD_min, s->DTD_min + s->n_DTD_min);
    } else {
        h->DTD_min = { 0.0 };
    }
    h->settings.DTD_min = nullptr;
    if (h->settings.n_threads <= 0) {
        h->settings.n_threads = omp_get_max_threads();
    }
    return GADFIT_B200_OK;
}

// 0 = automatic (dense for n <= 400), 1 = dense, 2 = block-arrow; +4 = accumulate the
// shared block, J^T r and chi2 in long double (rounding probe)
int gadfit_oracle_set_mode(OracleSolver* h, int mode)
{
    h->mode = mode & 3;
    h->extended = (mode & 4) != 0;
    return GADFIT_B200_OK;
}

int gadfit_oracle_default_gadf_options(gadfit_b200_gadf_options* out)
{
    *out = { 0, 0, 0, 0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0 };
    return GADFIT_B200_OK;
}

int gadfit_oracle_set_gadf_options(OracleSolver* h, const gadfit_b200_gadf_options* o)
{
    return guarded(h, [&] { h->gadf = *o; });
}

int gadfit_oracle_get_stop_reason(OracleSolver* h, int* out)
{
    *out = h->stop_reason;
    return GADFIT_B200_OK;
}

int gadfit_oracle_fit(OracleSolver* h, double lambda)
{
    return guarded(h, [&] { h->fit(lambda); });
}

int gadfit_oracle_get_par(OracleSolver* h, int i_par, int i_dataset, double* out)
{
    return guarded(h, [&] { *out = h->pars.at(i_dataset).at(i_par).val; });
}

int gadfit_oracle_get_value(OracleSolver* h, double arg, int i_dataset, double* out)
{
    return guarded(h, [&] { *out = h->eval(h->pars.at(i_dataset), arg); });
}

int gadfit_oracle_chi2(OracleSolver* h, double* out)
{
    return guarded(h, [&] { *out = h->chi2(); });
}

int gadfit_oracle_dof(OracleSolver* h, int* out)
{
    *out = h->dof;
    return GADFIT_B200_OK;
}

int gadfit_oracle_get_array(OracleSolver* h, int which, double* buf, size_t cap, size_t* need)
{
    return guarded(h, [&] {
        std::vector<double> data;
        const long long n { h->n_active };
        switch (which) {
        case GADFIT_B200_JACOBIAN: {
            data.assign(static_cast<size_t>(h->n_datapoints) * n, 0.0);
            for (size_t s = 0; s < h->sets.size(); ++s) {
                for (int i = 0; i < h->sets[s].n; ++i) {
                    const long long row { h->set_begin[s] + i };
                    for (size_t a = 0; a < h->column[s].size(); ++a) {
                        data[row * n + h->column[s][a]] = h->jac[row * h->row_stride + a];
                    }
                }
            }
            break;
        }
        case GADFIT_B200_JTJ:
            data = h->denseJTJ();
            break;
        case GADFIT_B200_DTD:
        case GADFIT_B200_LEFT_SIDE: {
            data = which == GADFIT_B200_DTD ? std::vector<double>(n * n, 0.0) : h->denseJTJ();
            const double f { which == GADFIT_B200_DTD ? 1.0 : h->lambda_lhs };
            for (long long c = 0; c < n; ++c) {
                data[c * n + c] += f * h->DTD[c];
            }
            break;
        }
        case GADFIT_B200_RIGHT_SIDE:
            data = h->right_side;
            break;
        case GADFIT_B200_RESIDUALS:
            data = h->residuals;
            break;
        case GADFIT_B200_INV_JTJ: {
            // lm_solver.cpp:578-585 with lapack_fallback.cpp:107-125
            const auto A { h->denseJTJ() };
            std::vector<double> Lo;
            OracleSolver::cholesky(static_cast<int>(n), A, Lo);
            data.assign(n * n, 0.0);
            for (long long j = 0; j < n; ++j) {
                std::vector<double> e(n, 0.0);
                e[j] = 1.0;
                OracleSolver::cholSolve(static_cast<int>(n), Lo, e.data());
                for (long long i = 0; i < n; ++i) {
                    data[i * n + j] = e[i];
                }
            }
            break;
        }
        case GADFIT_B200_DELTA1:
            data = h->delta1;
            break;
        case GADFIT_B200_DELTA2:
            data = h->delta2;
            break;
        case GADFIT_B200_OMEGA:
            data = h->omega;
            break;
        default:
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "unknown array" };
        }
        *need = data.size();
        const size_t m { std::min(cap, data.size()) };
        if (buf != nullptr && m > 0) {
            std::memcpy(buf, data.data(), m * sizeof(double));
        }
    });
}

int gadfit_oracle_get_trace(OracleSolver* h, gadfit_b200_trial* buf, size_t cap, size_t* n)
{
    *n = h->trace.size();
    const size_t m { std::min(cap, h->trace.size()) };
    if (buf != nullptr && m > 0) {
        std::memcpy(buf, h->trace.data(), m * sizeof(gadfit_b200_trial));
    }
    return GADFIT_B200_OK;
}

int gadfit_oracle_n_iterations(OracleSolver* h, int* out)
{
    *out = h->n_iterations;
    return GADFIT_B200_OK;
}

int gadfit_oracle_get_timings(OracleSolver* h, gadfit_b200_timings* out)
{
    for (int i = 0; i < 5; ++i) {
        out->seconds[i] = h->seconds[i];
        out->calls[i] = h->calls[i];
    }
    return GADFIT_B200_OK;
}

const char* gadfit_oracle_last_error(OracleSolver* h)
{
    return h->last_error.c_str();
}

} // extern "C"
