// Host side of the B200-native LM solver + the C-ABI of include/gadfit_b200.h.
//
// Thin by design: bookkeeping that the reference does in LMsolver
// (c++/gadfit/lm_solver.cpp) -- data set / parameter registration, the
// Jacobian column map, the order of the phases of fit() -- lives here; every
// floating point operation of the fit runs in the kernels of
// sweep_kernels.cuh and block_arrow_kernels.cuh.  The LM state (lambda, chi2,
// accept/reject decision) is device resident; the host reads one small record
// per trial step to know which phase to launch next and to keep the trace.

#include "model_ops.cuh"
#include "plugin_abi.cuh"
#include "peer_reduce.cuh"

#include <algorithm>
#include <cmath>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <limits>
#include <sstream>
#include <dlfcn.h>
#include <mutex>
#include <string_view>
#include <numeric>

namespace gadfit_b200 {

namespace {

constexpr size_t kDenseLimit { size_t { 1 } << 27 };

struct Partition
{
    int first { 0 };
    int end { 0 };
    bool split { false };
};

// Whole data sets per rank, balanced by points; with fewer data sets than
// ranks every rank takes a slice of the points of every data set (the
// img_bounds idea of fortran/gadfit/gadfit.F90:539-551)
auto compute_partition(const int n_sets, const int* n_points, const int rank, const int world)
  -> Partition
{
    if (world <= 1) {
        return { 0, n_sets, false };
    }
    if (n_sets < world) {
        return { 0, n_sets, true };
    }
    long long total { 0 };
    for (int s = 0; s < n_sets; ++s) {
        total += n_points[s];
    }
    Partition part { n_sets, n_sets, false };
    long long prefix { 0 };
    bool found { false };
    for (int s = 0; s < n_sets; ++s) {
        int owner { 0 };
        if (total > 0) {
            const long double mid { static_cast<long double>(prefix)
                                    + 0.5L * static_cast<long double>(n_points[s]) };
            owner = static_cast<int>(mid * world / static_cast<long double>(total));
        } else {
            owner = static_cast<int>(static_cast<long long>(s) * world / n_sets);
        }
        owner = std::min(owner, world - 1);
        if (owner == rank) {
            if (!found) {
                part.first = s;
                found = true;
            }
            part.end = s + 1;
        }
        prefix += n_points[s];
    }
    if (!found) {
        part.first = part.end = 0;
    }
    return part;
}

__global__ void k_fill(double* p, const long long n, const double v)
{
    const long long i { static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x };
    if (i < n) {
        p[i] = v;
    }
}

// sigma -> 1 / sigma, once at upload: the sweeps multiply
__global__ void k_invert(double* p, const long long n)
{
    const long long i { static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x };
    if (i < n) {
        p[i] = 1.0 / p[i];
    }
}

__global__ void k_copy_chi2_to_old(LmState* state)
{
    state->old_chi2 = state->new_chi2;
}

// Dense symmetric inverse through the Cholesky factor (LMsolver::getInvJTJ,
// lm_solver.cpp:578-585 = dpptrf + dpotri).  Small n only, one thread.
__global__ void k_dense_inverse(const int n, const double* A, double* L, double* inv)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) {
        return;
    }
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < i; ++j) {
            double v { A[i * n + j] };
            for (int k = 0; k < j; ++k) {
                v -= L[i * n + k] * L[j * n + k];
            }
            L[i * n + j] = v / L[j * n + j];
        }
        double v { A[i * n + i] };
        for (int k = 0; k < i; ++k) {
            v -= L[i * n + k] * L[i * n + k];
        }
        L[i * n + i] = ::sqrt(v);
    }
    for (int c = 0; c < n; ++c) {
        // Solve L L' x = e_c; y is stored in column c of inv
        for (int i = 0; i < n; ++i) {
            double v { i == c ? 1.0 : 0.0 };
            for (int j = 0; j < i; ++j) {
                v -= L[i * n + j] * inv[j * n + c];
            }
            inv[i * n + c] = v / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double v { inv[i * n + c] };
            for (int j = n - 1; j > i; --j) {
                v -= L[j * n + i] * inv[j * n + c];
            }
            inv[i * n + c] = v / L[i * n + i];
        }
    }
}

// ---- block-arrow inverse of J^T J (LMsolver::getInvJTJ) ---------------------
//
// The reference inverts the dense n x n matrix (dpptrf + dpotri,
// lm_solver.cpp:578-585, lapack.cpp:113-125).  Here J^T J = [G C'; C L] with L
// block diagonal, and its inverse follows from the factors the Schur solve
// already uses (lambda = 0):  L_s = R_s R_s',  Y_s = R_s^-1 C_s',
// S = G - sum_s Y_s' Y_s,
//     inv_gg       = S^-1
//     W_s          = L_s^-1 C_s' = R_s^-T Y_s
//     inv_sg       = -W_s inv_gg
//     inv_ss       = L_s^-1 + W_s inv_gg W_s'
//     inv_st (s/=t) = W_s inv_gg W_t' = -inv_sg(s) W_t'   (dense assembly only)
// One thread per data set; the layout of a block record is
// [ss (NP x NP) | sg (NP x NP, row i = local slot, column a = global slot) | W].
__global__ void k_schur_matrix(const int pg,
                               const double* __restrict__ shared_sums,
                               const double* __restrict__ schur_sums,
                               double* __restrict__ A)
{
    const int e { static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x) };
    if (e < pg * pg) {
        const int i { e / pg };
        const int j { e % pg };
        const int o { i <= j ? tri_index(pg, i, j) : tri_index(pg, j, i) };
        A[e] = shared_sums[o] - schur_sums[o];
    }
}

__global__ void k_inverse_local(const int NP,
                                const int n_sets,
                                const unsigned* __restrict__ active,
                                const unsigned global_mask,
                                const double* __restrict__ factors, // LocalFactor<NP> records
                                const double* __restrict__ inv_gg,  // [pg][pg]
                                const int pg,
                                double* __restrict__ blocks,        // [n_sets][3 NP NP]
                                double* __restrict__ work)          // [n_sets][NP NP]
{
    const int s { static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x) };
    if (s >= n_sets) {
        return;
    }
    const int pl { popcount32(active[s] & ~global_mask) };
    const long long NN { static_cast<long long>(NP) * NP };
    const double* R { factors + s * (2 * NN + NP) };
    const double* Y { R + NN };
    double* ss { blocks + s * 3 * NN };
    double* sg { ss + NN };
    double* W { sg + NN };
    double* T { work + s * NN }; // R^-1, lower triangular
    for (int c = 0; c < pl; ++c) {
        for (int i = 0; i < pl; ++i) {
            double v { i == c ? 1.0 : 0.0 };
            for (int k = c; k < i; ++k) {
                v -= R[i * NP + k] * T[k * NP + c];
            }
            T[i * NP + c] = i < c ? 0.0 : v / R[i * NP + i];
        }
    }
    // W = R^-T Y by back substitution
    for (int a = 0; a < pg; ++a) {
        for (int i = pl - 1; i >= 0; --i) {
            double v { Y[i * NP + a] };
            for (int k = i + 1; k < pl; ++k) {
                v -= R[k * NP + i] * W[k * NP + a];
            }
            W[i * NP + a] = v / R[i * NP + i];
        }
    }
    for (int i = 0; i < pl; ++i) {
        for (int a = 0; a < pg; ++a) {
            double v { 0.0 };
            for (int b = 0; b < pg; ++b) {
                v += W[i * NP + b] * inv_gg[b * pg + a];
            }
            sg[i * NP + a] = -v;
        }
    }
    for (int i = 0; i < pl; ++i) {
        for (int j = 0; j < pl; ++j) {
            double v { 0.0 };
            for (int k = (i > j ? i : j); k < pl; ++k) {
                v += T[k * NP + i] * T[k * NP + j]; // L_s^-1 = R^-T R^-1
            }
            for (int a = 0; a < pg; ++a) {
                v -= sg[i * NP + a] * W[j * NP + a];
            }
            ss[i * NP + j] = v;
        }
    }
}

// Dense n x n view of the block-arrow inverse; col_set[c] = data set of column
// c (-1: shared), col_idx[c] = its slot among the shared / that set's local columns
__global__ void k_inverse_assemble(const long long n,
                                   const int NP,
                                   const int pg,
                                   const int* __restrict__ col_set,
                                   const int* __restrict__ col_idx,
                                   const double* __restrict__ inv_gg,
                                   const double* __restrict__ blocks,
                                   double* __restrict__ out)
{
    const long long e { static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x };
    if (e >= n * n) {
        return;
    }
    // (r, c) and (c, r) take the same path: the output is symmetric to the bit
    const long long r { e / n <= e % n ? e / n : e % n };
    const long long c { e / n <= e % n ? e % n : e / n };
    const int sr { col_set[r] };
    const int sc { col_set[c] };
    const int ir { col_idx[r] };
    const int ic { col_idx[c] };
    const long long NN { static_cast<long long>(NP) * NP };
    double v { 0.0 };
    if (sr < 0 && sc < 0) {
        v = inv_gg[ir * pg + ic];
    } else if (sr < 0) {
        v = blocks[sc * 3 * NN + NN + ic * NP + ir];
    } else if (sc < 0) {
        v = blocks[sr * 3 * NN + NN + ir * NP + ic];
    } else if (sr == sc) {
        v = blocks[sr * 3 * NN + ir * NP + ic];
    } else {
        const double* sg { blocks + sr * 3 * NN + NN + ir * NP };
        const double* W { blocks + sc * 3 * NN + 2 * NN + ic * NP };
        for (int b = 0; b < pg; ++b) {
            v -= sg[b] * W[b];
        }
    }
    out[e] = v;
}

// Tile range and first data set of every CTA of one sweep (SweepArgs::cta_*)
struct CtaTable
{
    DeviceBuffer<int> tile_begin; // [grid + 1]
    DeviceBuffer<int> first_set;  // [grid]
};

struct SlotTable
{
    int grid { 0 };
    int n_slots { 0 };
    DeviceBuffer<int> base;  // [grid]  first slot of each CTA
    DeviceBuffer<int> first; // [n_sets] first slot of each data set
    DeviceBuffer<int> count; // [n_sets]
    CtaTable cta;
};

} // namespace

// State and phases of one solver; the opaque C handle derives from it
// GADFIT_B200_TRACE_HOST=1: wall-clock time of the host-side phases of fit()
// to stderr (developer aid for the multi-GPU set-up / gather costs)
struct HostPhase
{
    const char* name;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit HostPhase(const char* name)
      : name { name }, t0 { std::chrono::steady_clock::now() },
        on { std::getenv("GADFIT_B200_TRACE_HOST") != nullptr }
    {}
    ~HostPhase()
    {
        if (on) {
            const double ms { std::chrono::duration<double, std::milli>(
                                std::chrono::steady_clock::now() - t0)
                                .count() };
            std::fprintf(stderr, "[gadfit_b200 host] %-28s %9.3f ms\n", name, ms);
        }
    }
};

struct SolverImpl
{
    // ---- problem description (host; mirrors the private state of LMsolver,
    // lm_solver.h:101-123)
    struct DataSet
    {
        int n;
        const double* x;
        const double* y;
        const double* errors;
    };
    std::unique_ptr<ModelOps> ops;
    int NP { 0 };
    std::vector<DataSet> sets;
    std::vector<double> par;         // [S][NP]
    bool indexing_dirty { true };    // data sets or parameter roles changed since prepare_indexing()
    // page-locked view of `par` for the gather of all ranks' parameters at the
    // end of a multi-GPU fit (pageable: ~7 GB/s, 45 MB for config 5 on 8 GPUs)
    void* par_registered { nullptr };
    size_t par_registered_bytes { 0 };
    std::vector<int> n_set_pars;     // FitFunction::getNumPars per data set
    std::vector<unsigned> active;    // Indices::active as bit masks
    unsigned global_mask { 0U };     // Indices::global
    std::vector<std::string> names;
    bool set_par_called { false };
    gadfit_b200_settings settings {};
    std::vector<double> DTD_min;
    std::string last_error;

    // ---- indexing (prepareIndexing, lm_solver.cpp:119-190)
    long long n_active { 0 };
    long long n_datapoints { 0 };
    int dof { 0 };
    int pg { 0 };
    std::vector<int> global_pars;       // model parameter of global column a
    std::vector<long long> local_off;   // first local column of data set s

    // ---- device
    int device { 0 };
    cudaStream_t stream { nullptr };
    bool device_ready { false };
    bool data_uploaded { false };
    long long n_launches { 0 };

    // communicator
    int rank { 0 };
    int world { 1 };
    ncclComm_t comm { nullptr };
    PeerReduce peer {};
    long long n_peer_reductions { 0 };
    long long n_nccl_reductions { 0 };

    // this rank's share of the problem
    Partition part {};
    int n_local_sets { 0 };
    long long n_local_points { 0 };
    std::vector<long long> local_begin; // first owned point of each local set
    std::vector<long long> local_count; // owned points of each local set

    DeviceBuffer<double> d_x, d_y, d_sig; // d_sig holds 1 / sigma
    DeviceBuffer<int> d_set_tile_begin;        // [n_local_sets + 1]
    DeviceBuffer<long long> d_set_point_begin; // [n_local_sets + 1]
    std::vector<int> h_tile_set;               // data set of each tile (host, for the slots)
    int n_tiles { 0 };
    SlotTable slots_jac, slots_omega;
    int grid_chi2 { 0 };
    CtaTable cta_chi2;
    std::vector<int> h_set_tile_begin; // host copy of d_set_tile_begin

    DeviceBuffer<double> d_par, d_par_old, d_par_jac;
    DeviceBuffer<unsigned> d_active;
    DeviceBuffer<double> d_chi2_partials, d_jac_partials, d_omega_partials;
    DeviceBuffer<double> d_M;      // [n_local_sets][NT]
    DeviceBuffer<double> d_vsets;  // [n_local_sets][NP]  J^T omega by parameter
    DeviceBuffer<double> d_dtd_local, d_dtd_global;
    DeviceBuffer<double> d_contrib, d_stage;
    DeviceBuffer<double> d_shared_sums, d_schur_sums, d_rhs2_sums, d_ratio_sums;
    DeviceBuffer<double> d_chol, d_delta1_g, d_delta2_g;
    DeviceBuffer<double> d_delta1_par, d_delta2_par;
    DeviceBuffer<unsigned char> d_factors;
    DeviceBuffer<double> d_J, d_resid, d_omega;
    DeviceBuffer<LmState> d_state;  // [0] live state, [1] record, [2] copy
    DeviceBuffer<gadfit_b200_trial> d_trace; // device-driven loop: one record per trial
    DeviceBuffer<unsigned> d_chi2_counter;   //   CTAs of the chi2 sweep that are done
    gadfit_b200_trial* h_trace { nullptr };  // pinned, kTraceBatch records
    // pipelined upload (upload_data(true)): copies on their own stream, the
    // number of leading points that are complete published after every chunk
    cudaStream_t copy_stream { nullptr };
    cudaEvent_t upload_done { nullptr };
    DeviceBuffer<unsigned long long> d_arrived;
    unsigned long long* h_arrived { nullptr }; // pinned, one entry per chunk
    size_t h_arrived_capacity { 0 };
    bool upload_in_flight { false }; // first Jacobian sweep must wait / invert sigma
    static constexpr int kTraceBatch { 64 };
    DeviceBuffer<double> d_scratch;
    LmState* h_mirror { nullptr };  // pinned, 2 entries: record, state
    double* h_scalar { nullptr };   // pinned

    // ---- results of the last fit
    std::vector<gadfit_b200_trial> trace;
    int n_iterations { 0 };
    bool fitted { false };
    double lambda_lhs { 0.0 };      // lambda of the last left-hand side
    bool have_delta2 { false };
    // text output of the reference (printIterationResults / printTimings)
    gadfit_b200_log_fn log_fn { nullptr };
    void* log_user { nullptr };
    // the Fortran driver (gadfit_b200_gadf_options)
    gadfit_b200_gadf_options gadf { 0, 0, 0, 0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0 };
    double umnigh_a { 0.5 }; // gadfit.F90:515, a saved variable: persists across fits
    int stop_reason { GADFIT_B200_STOP_ITERATION_LIMIT };
    DeviceBuffer<double> d_old_delta1_par;
    DeviceBuffer<double> d_old_delta1_g;
    DeviceBuffer<double> d_gadf_sums;
    PhaseTimer t_jacobian, t_chi2, t_linalg, t_omega, t_main, t_jac_kernel;

    ~SolverImpl()
    {
        if (comm != nullptr) {
            Nccl::instance().commDestroy(comm);
        }
        if (par_registered != nullptr) {
            cudaHostUnregister(par_registered);
        }
        if (h_mirror != nullptr) {
            cudaFreeHost(h_mirror);
        }
        if (h_scalar != nullptr) {
            cudaFreeHost(h_scalar);
        }
        if (h_trace != nullptr) {
            cudaFreeHost(h_trace);
        }
        if (h_arrived != nullptr) {
            cudaFreeHost(h_arrived);
        }
        if (upload_done != nullptr) {
            cudaEventDestroy(upload_done);
        }
        if (copy_stream != nullptr) {
            cudaStreamDestroy(copy_stream);
        }
        if (stream != nullptr) {
            cudaStreamDestroy(stream);
        }
    }

    // ------------------------------------------------------------ host logic

    auto add_dataset(const int n, const double* x, const double* y, const double* e) -> void
    {
        // lm_solver.cpp:35-37
        if (set_par_called) {
            throw StatusError { GADFIT_B200_LATE_ADD_DATASET_CALL,
                                "addDataset called after setPar" };
        }
        if (n < 0) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "negative number of points" };
        }
        sets.push_back({ n, x, y, e });
        indexing_dirty = true;
        if (par_registered != nullptr) { // the vector may move
            cudaHostUnregister(par_registered);
            par_registered = nullptr;
        }
        par.resize(sets.size() * NP, 0.0);
        n_set_pars.push_back(0);
        // A new data set starts as a copy of the fitting function of the
        // last one (lm_solver.cpp:52-54); parameters cannot have been set yet
        active.push_back(0U);
        data_uploaded = false;
    }

    auto set_par(const int i_par, const double val, const bool is_active, const int i_dataset)
      -> void
    {
        // lm_solver.cpp:65-68
        const int S { static_cast<int>(sets.size()) };
        if (i_dataset >= S || S == 0) {
            throw StatusError { GADFIT_B200_SET_PAR_INVALID_INDEX,
                                "invalid data set index " + std::to_string(i_dataset) };
        }
        if (i_par < 0 || i_par >= NP || i_dataset < -1) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT,
                                "parameter index outside the fitting function" };
        }
        set_par_called = true;
        const unsigned bit { 1U << i_par };
        // (the column map is rebuilt at the next fit() only if a role changed)
        const unsigned old_global { global_mask };
        const auto assign { [&](const int s) {
            par[static_cast<size_t>(s) * NP + i_par] = val;
            const int n_pars_new { std::max(n_set_pars[s], i_par + 1) };
            const unsigned active_new { is_active ? (active[s] | bit) : (active[s] & ~bit) };
            if (n_pars_new != n_set_pars[s] || active_new != active[s]) {
                indexing_dirty = true;
            }
            n_set_pars[s] = n_pars_new;
            active[s] = active_new;
        } };
        if (i_dataset == -1) {
            // lm_solver.cpp:73-88
            if (is_active) {
                global_mask |= bit;
            } else {
                global_mask &= ~bit;
            }
            for (int s = 0; s < S; ++s) {
                assign(s);
            }
        } else {
            // lm_solver.cpp:89-98
            global_mask &= ~bit;
            assign(i_dataset);
        }
        if (global_mask != old_global) {
            indexing_dirty = true;
        }
    }

    // lm_solver.cpp:119-190, same order of checks
    auto prepare_indexing() -> void
    {
        const int S { static_cast<int>(sets.size()) };
        if (S == 0) {
            throw StatusError { GADFIT_B200_NO_FITTING_PARAMETERS, "no data sets" };
        }
        if (!indexing_dirty) {
            return; // same data sets and parameter roles as at the last successful call
        }
        n_active = 0;
        for (const unsigned a : active) {
            n_active += popcount32(a);
        }
        n_active -= static_cast<long long>(S - 1) * popcount32(global_mask);
        if (n_active > 0 && S > 1 && global_mask == 0U) {
            throw StatusError { GADFIT_B200_NO_GLOBAL_PARAMETERS,
                                "several data sets but no global fitting parameter" };
        }
        n_datapoints = 0;
        for (const auto& set : sets) {
            n_datapoints += set.n;
        }
        const long long degrees { n_datapoints - n_active };
        if (degrees < 0) {
            throw StatusError { GADFIT_B200_NEGATIVE_DEGREES_OF_FREEDOM,
                                "more fitting parameters than data points" };
        }
        // No degrees of freedom: chi2/DOF has no meaning (lm_solver.cpp:140-144)
        dof = degrees == 0 ? 1 : static_cast<int>(std::min<long long>(degrees, INT32_MAX));
        for (const int n : n_set_pars) {
            if (n != n_set_pars.front()) {
                throw StatusError { GADFIT_B200_UNINITIALIZED_PARAMETER,
                                    "a parameter is not set for every data set" };
            }
        }
        if (n_set_pars.front() != NP) {
            throw StatusError { GADFIT_B200_UNINITIALIZED_PARAMETER,
                                "the fitting function has " + std::to_string(NP)
                                  + " parameters" };
        }
        // Column map: globals first in ascending parameter index, then the
        // locals of data set 0, 1, ... (lm_solver.cpp:155-186)
        global_pars.clear();
        for (int k = 0; k < NP; ++k) {
            if (((global_mask >> k) & 1U) != 0U) {
                global_pars.push_back(k);
            }
        }
        pg = static_cast<int>(global_pars.size());
        local_off.resize(S);
        long long next { pg };
        for (int s = 0; s < S; ++s) {
            local_off[s] = next;
            next += popcount32(active[s] & ~global_mask);
        }
        if (n_active == 0) {
            throw StatusError { GADFIT_B200_NO_FITTING_PARAMETERS,
                                "no active fitting parameters" };
        }
        indexing_dirty = false;
    }

    // Column of parameter k of data set s, -1 if passive
    [[nodiscard]] auto column(const int s, const int k) const -> long long
    {
        const unsigned bit { 1U << k };
        if ((active[s] & bit) == 0U) {
            return -1;
        }
        if ((global_mask & bit) != 0U) {
            return slot_of(global_mask, k);
        }
        return local_off[s] + slot_of(active[s] & ~global_mask, k);
    }

    // ---------------------------------------------------------- device setup

    auto launched(const int n = 1) -> void
    {
        n_launches += n;
        GADFIT_CUDA_TRY(cudaGetLastError());
    }

    auto ensure_device() -> void
    {
        GADFIT_CUDA_TRY(cudaSetDevice(device));
        if (device_ready) {
            return;
        }
        GADFIT_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        GADFIT_CUDA_TRY(cudaMallocHost(&h_mirror, 2 * sizeof(LmState)));
        GADFIT_CUDA_TRY(cudaMallocHost(&h_scalar, 24 * sizeof(double)));
        GADFIT_CUDA_TRY(cudaMallocHost(&h_trace, kTraceBatch * sizeof(gadfit_b200_trial)));
        d_chi2_counter.zero(1, stream);
        ops->init_device();
        d_state.zero(3, stream);
        d_scratch.ensure(8);
        device_ready = true;
    }

    auto sm_count() const -> int
    {
        int n { 0 };
        GADFIT_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
        return n;
    }

    auto all_reduce(double* buf, const size_t n) -> void
    {
        if (world <= 1 || n == 0) {
            return;
        }
        // Small reductions (everything inside the LM loop): one-shot all-reduce
        // over NVLink peer memory, peer_reduce.cuh.  Large, rare ones (the
        // point-split M, the final parameter gather): NCCL.
        if (peer.active && n <= static_cast<size_t>(kPeerMaxDoubles)) {
            peer.all_reduce(buf, static_cast<int>(n), stream);
            launched();
            ++n_peer_reductions;
            return;
        }
        auto& nccl { Nccl::instance() };
        nccl.check(nccl.allReduce(buf, buf, n, ncclDouble, ncclSum, comm, stream),
                   "ncclAllReduce");
        ++n_nccl_reductions;
    }

    // unroll > 1: the cut points are moved up to the next multiple of `unroll`
    // tiles counted from the start of the data set they fall into (or to its
    // end), as k_chi2 would compute them itself
    auto build_cta_table(CtaTable& table, const int grid, const int unroll) -> void
    {
        std::vector<int> begin(grid + 1, n_tiles);
        std::vector<int> first(std::max(grid, 1), 0);
        for (int b = 0; b < grid; ++b) {
            int t { static_cast<int>(static_cast<long long>(n_tiles) * b / grid) };
            if (unroll > 1 && b > 0 && t < n_tiles) {
                const int s { h_tile_set[t] };
                const int groups { (t - h_set_tile_begin[s] + unroll - 1) / unroll };
                t = std::min(h_set_tile_begin[s] + groups * unroll, h_set_tile_begin[s + 1]);
            }
            begin[b] = t;
            first[b] = t < n_tiles ? h_tile_set[t] : 0;
        }
        table.tile_begin.upload(begin, stream);
        table.first_set.upload(first, stream);
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream)); // pageable sources die at scope exit
    }

    auto build_slots(SlotTable& table, const int grid) -> void
    {
        table.grid = grid;
        build_cta_table(table.cta, grid, 1);
        std::vector<int> base(std::max(grid, 1), 0);
        std::vector<int> first(std::max(n_local_sets, 1), 0);
        std::vector<int> count(std::max(n_local_sets, 1), 0);
        int slot { 0 };
        for (int b = 0; b < grid; ++b) {
            const int t0 { static_cast<int>(static_cast<long long>(n_tiles) * b / grid) };
            const int t1 { static_cast<int>(static_cast<long long>(n_tiles) * (b + 1)
                                            / grid) };
            base[b] = slot;
            int cur { -1 };
            for (int t = t0; t < t1; ++t) {
                const int s { h_tile_set[t] };
                if (s != cur) {
                    if (count[s] == 0) {
                        first[s] = slot;
                    }
                    ++count[s];
                    ++slot;
                    cur = s;
                }
            }
        }
        table.n_slots = slot;
        table.base.upload(base, stream);
        table.first.upload(first, stream);
        table.count.upload(count, stream);
        // The host vectors die at scope exit; the copies are from pageable
        // memory and therefore already staged by the runtime
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
    }

    // x, y, sigma of the owned points in chunks of ~4M points, chunk-major
    // (x_c, y_c, sigma_c, then the arrival counter), on the copy stream
    auto enqueue_chunked_copies() -> void
    {
        if (copy_stream == nullptr) {
            GADFIT_CUDA_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
            GADFIT_CUDA_TRY(cudaEventCreateWithFlags(&upload_done, cudaEventDisableTiming));
        }
        constexpr long long chunk_points { 1LL << 22 };
        // pieces: (device offset, length, host pointers)
        struct Piece
        {
            long long off;
            long long n;
            const double* x;
            const double* y;
            const double* e;
        };
        std::vector<Piece> pieces;
        long long off { 0 };
        for (int ls = 0; ls < n_local_sets; ++ls) {
            const auto& set { sets[part.first + ls] };
            for (long long done = 0; done < local_count[ls];) {
                const long long n { std::min(chunk_points, local_count[ls] - done) };
                const long long h { local_begin[ls] + done };
                // extend the previous piece if host and device ranges continue it
                if (!pieces.empty() && pieces.back().n + n <= chunk_points
                    && pieces.back().x + pieces.back().n == set.x + h
                    && pieces.back().y + pieces.back().n == set.y + h
                    && pieces.back().e + pieces.back().n == set.errors + h) {
                    pieces.back().n += n;
                } else {
                    pieces.push_back({ off, n, set.x + h, set.y + h, set.errors + h });
                }
                off += n;
                done += n;
            }
        }
        if (pieces.size() > h_arrived_capacity) {
            if (h_arrived != nullptr) {
                // the previous upload's counter copies have long completed
                GADFIT_CUDA_TRY(cudaStreamSynchronize(copy_stream));
                cudaFreeHost(h_arrived);
                h_arrived = nullptr;
            }
            GADFIT_CUDA_TRY(cudaMallocHost(&h_arrived, pieces.size() * sizeof(unsigned long long)));
            h_arrived_capacity = pieces.size();
        }
        d_arrived.zero(1, stream);
        // the copies may start once the counter is zero and the buffers exist
        GADFIT_CUDA_TRY(cudaEventRecord(upload_done, stream));
        GADFIT_CUDA_TRY(cudaStreamWaitEvent(copy_stream, upload_done, 0));
        for (size_t k = 0; k < pieces.size(); ++k) {
            const Piece& p { pieces[k] };
            const size_t bytes { static_cast<size_t>(p.n) * sizeof(double) };
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_x.ptr + p.off, p.x, bytes, cudaMemcpyHostToDevice,
                                            copy_stream));
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_y.ptr + p.off, p.y, bytes, cudaMemcpyHostToDevice,
                                            copy_stream));
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_sig.ptr + p.off, p.e, bytes, cudaMemcpyHostToDevice,
                                            copy_stream));
            h_arrived[k] = static_cast<unsigned long long>(p.off + p.n);
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_arrived.ptr, h_arrived + k, sizeof(unsigned long long),
                                            cudaMemcpyHostToDevice, copy_stream));
        }
        GADFIT_CUDA_TRY(cudaEventRecord(upload_done, copy_stream));
    }

    // After the first Jacobian sweep of a pipelined upload has been enqueued:
    // everything later on the compute stream may assume the data are complete
    auto finish_pipelined_upload() -> void
    {
        if (upload_in_flight) {
            GADFIT_CUDA_TRY(cudaStreamWaitEvent(stream, upload_done, 0));
            upload_in_flight = false;
        }
    }

    // Copy this rank's share of the data sets to HBM, concatenated data-set
    // major, and lay out the tiles.
    // pipelined (only from fit()): the copies go chunk by chunk through a copy
    // stream of their own and are NOT waited for; the first Jacobian sweep
    // follows the arrival counter (SweepArgs::arrived) and inverts sigma as it
    // goes, so the sweep -- and all host-side set-up below -- overlaps the PCIe
    // transfer.  Needs explicit errors for every owned data set.
    auto upload_data(const bool pipelined = false) -> void
    {
        ensure_device();
        const int S { static_cast<int>(sets.size()) };
        std::vector<int> n_points(S);
        for (int s = 0; s < S; ++s) {
            n_points[s] = sets[s].n;
        }
        part = compute_partition(S, n_points.data(), rank, world);
        n_local_sets = part.end - part.first;
        local_begin.assign(n_local_sets, 0);
        local_count.assign(n_local_sets, 0);
        n_local_points = 0;
        for (int ls = 0; ls < n_local_sets; ++ls) {
            const long long n { sets[part.first + ls].n };
            if (part.split) {
                local_begin[ls] = n * rank / world;
                local_count[ls] = n * (rank + 1) / world - local_begin[ls];
            } else {
                local_count[ls] = n;
            }
            n_local_points += local_count[ls];
        }
        d_x.ensure(n_local_points);
        d_y.ensure(n_local_points);
        d_sig.ensure(n_local_points);
        // Merge host-contiguous data sets into single copies
        const auto copy_field { [&](const int field, double* dst) {
            long long off { 0 };
            int ls { 0 };
            while (ls < n_local_sets) {
                const auto field_ptr { [&](const int l) -> const double* {
                    const auto& set { sets[part.first + l] };
                    const double* p { field == 0 ? set.x
                                                 : (field == 1 ? set.y : set.errors) };
                    return p == nullptr ? nullptr : p + local_begin[l];
                } };
                const double* start { field_ptr(ls) };
                long long run { local_count[ls] };
                int next { ls + 1 };
                if (start != nullptr) {
                    while (next < n_local_sets && field_ptr(next) == start + run) {
                        run += local_count[next];
                        ++next;
                    }
                }
                if (run > 0) {
                    if (start == nullptr) {
                        if (field != 2) {
                            throw StatusError {
                                GADFIT_B200_INVALID_ARGUMENT,
                                "data set " + std::to_string(part.first + ls)
                                  + " is owned by this rank but has no data"
                            };
                        }
                        // Equal weights by default (lm_solver.cpp:43-48)
                        k_fill<<<blocks_for(run, 256), 256, 0, stream>>>(dst + off, run, 1.0);
                        launched();
                    } else {
                        GADFIT_CUDA_TRY(cudaMemcpyAsync(dst + off,
                                                        start,
                                                        run * sizeof(double),
                                                        cudaMemcpyHostToDevice,
                                                        stream));
                    }
                }
                off += run;
                ls = next;
            }
        } };
        bool all_have_errors { n_local_points > 0 };
        for (int ls = 0; ls < n_local_sets; ++ls) {
            const auto& set { sets[part.first + ls] };
            all_have_errors = all_have_errors
                              && (local_count[ls] == 0
                                  || (set.x != nullptr && set.y != nullptr && set.errors != nullptr));
        }
        upload_in_flight = false;
        if (pipelined && all_have_errors && n_local_points >= (1LL << 21)) {
            enqueue_chunked_copies();
            upload_in_flight = true;
        } else {
            copy_field(0, d_x.ptr);
            copy_field(1, d_y.ptr);
            copy_field(2, d_sig.ptr);
            if (n_local_points > 0) {
                k_invert<<<blocks_for(n_local_points, 256), 256, 0, stream>>>(d_sig.ptr,
                                                                              n_local_points);
                launched();
            }
        }

        // Tiles
        h_tile_set.clear();
        std::vector<int> set_tile_begin(n_local_sets + 1, 0);
        std::vector<long long> set_point_begin(n_local_sets + 1, 0);
        for (int ls = 0; ls < n_local_sets; ++ls) {
            const int tiles { static_cast<int>((local_count[ls] + kTile - 1) / kTile) };
            h_tile_set.insert(h_tile_set.end(), tiles, ls);
            set_tile_begin[ls + 1] = set_tile_begin[ls] + tiles;
            set_point_begin[ls + 1] = set_point_begin[ls] + local_count[ls];
        }
        n_tiles = static_cast<int>(h_tile_set.size());
        h_set_tile_begin = set_tile_begin;
        d_set_tile_begin.upload(set_tile_begin, stream);
        d_set_point_begin.upload(set_point_begin, stream);
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));

        const int sms { sm_count() };
        const auto grid_for { [&](const int which) {
            return std::max(1, std::min(n_tiles, sms * ops->ctas_per_sm(which)));
        } };
        grid_chi2 = grid_for(0);
        d_chi2_partials.ensure(grid_chi2);
        build_cta_table(cta_chi2, grid_chi2, ops->chi2_unroll());
        build_slots(slots_jac, grid_for(1));
        build_slots(slots_omega, grid_for(2));
        const int NA { NP + 1 };
        const int NT { NA * (NA + 1) / 2 };
        d_jac_partials.ensure(static_cast<size_t>(std::max(slots_jac.n_slots, 1)) * NT);
        d_omega_partials.ensure(static_cast<size_t>(std::max(slots_omega.n_slots, 1)) * NP);
        d_M.zero(static_cast<size_t>(std::max(n_local_sets, 1)) * NT, stream);
        d_vsets.zero(static_cast<size_t>(std::max(n_local_sets, 1)) * NP, stream);
        data_uploaded = true;
    }

    // Parameters and roles of the local data sets -> device
    auto upload_parameters(const bool pipelined = false) -> void
    {
        if (!data_uploaded) {
            upload_data(pipelined);
        }
        const size_t n { static_cast<size_t>(n_local_sets) * NP };
        d_par.upload(par.data() + static_cast<size_t>(part.first) * NP, n, stream);
        d_par_old.ensure(n);
        d_par_jac.ensure(n);
        d_active.upload(active.data() + part.first, n_local_sets, stream);
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
    }

    auto download_parameters() -> void
    {
        const HostPhase phase { "download_parameters" };
        const size_t n { static_cast<size_t>(n_local_sets) * NP };
        if (world > 1 && !part.split) {
            // Every rank ends up with the parameters of every data set: each
            // fills its block of a zeroed [S][NP] array, summed over ranks
            const size_t all { sets.size() * static_cast<size_t>(NP) };
            d_stage.zero(all, stream);
            if (n > 0) {
                GADFIT_CUDA_TRY(cudaMemcpyAsync(d_stage.ptr
                                                  + static_cast<size_t>(part.first) * NP,
                                                d_par.ptr,
                                                n * sizeof(double),
                                                cudaMemcpyDeviceToDevice,
                                                stream));
            }
            all_reduce(d_stage.ptr, all);
            const size_t bytes { all * sizeof(double) };
            if (bytes >= (size_t { 1 } << 20)
                && (par_registered != par.data() || par_registered_bytes != bytes)) {
                if (par_registered != nullptr) {
                    cudaHostUnregister(par_registered);
                    par_registered = nullptr;
                }
                if (cudaHostRegister(par.data(), bytes, cudaHostRegisterDefault) == cudaSuccess) {
                    par_registered = par.data();
                    par_registered_bytes = bytes;
                } else {
                    cudaGetLastError(); // not fatal: the copy is then a pageable one
                }
            }
            GADFIT_CUDA_TRY(cudaMemcpyAsync(par.data(),
                                            d_stage.ptr,
                                            all * sizeof(double),
                                            cudaMemcpyDeviceToHost,
                                            stream));
        } else if (n > 0) {
            GADFIT_CUDA_TRY(cudaMemcpyAsync(par.data() + static_cast<size_t>(part.first) * NP,
                                            d_par.ptr,
                                            n * sizeof(double),
                                            cudaMemcpyDeviceToHost,
                                            stream));
        }
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
    }

    [[nodiscard]] auto sweep_args(const CtaTable* cta = nullptr) const -> SweepArgs
    {
        return { d_x.ptr, d_y.ptr,      d_sig.ptr, d_set_tile_begin.ptr, d_set_point_begin.ptr,
                 n_local_sets, n_tiles, d_par.ptr, &d_state.ptr->flags, nullptr, 0,
                 upload_in_flight ? d_arrived.ptr : nullptr, upload_in_flight ? 1 : 0,
                 cta != nullptr ? cta->tile_begin.ptr : nullptr,
                 cta != nullptr ? cta->first_set.ptr : nullptr };
    }

    // --------------------------------------------------------------- phases

    // chi2 at the current parameters -> *dst (device), summed over ranks
    auto chi2_phase(double* dst) -> void
    {
        t_chi2.start(stream);
        if (n_tiles > 0) {
            ops->chi2(grid_chi2, stream, sweep_args(&cta_chi2), d_chi2_partials.ptr, TrialFinish {});
            launched();
            if (world > 1 && peer.active) {
                // final sum of the CTA partials and the all-reduce in one kernel
                peer.sum_partials_all_reduce(d_chi2_partials.ptr, grid_chi2, dst, stream);
                launched();
                ++n_peer_reductions;
                t_chi2.stop(stream);
                return;
            }
            k_sum_partials<<<1, kThreads, 0, stream>>>(d_chi2_partials.ptr, grid_chi2, dst);
            launched();
        } else {
            GADFIT_CUDA_TRY(cudaMemsetAsync(dst, 0, sizeof(double), stream));
        }
        all_reduce(dst, 1);
        t_chi2.stop(stream);
    }

    // Deterministic column sums of contrib[n_local_sets][V] -> dst[V];
    // reduce: followed by the sum over ranks -- over peer memory the last
    // stage and the all-reduce are one kernel (peer_reduce.cuh)
    auto column_sums(const double* contrib, const int V, double* dst, const bool reduce = false)
      -> void
    {
        if (V == 0) {
            return;
        }
        if (n_local_sets == 0) {
            GADFIT_CUDA_TRY(cudaMemsetAsync(dst, 0, V * sizeof(double), stream));
            if (reduce) {
                all_reduce(dst, V);
            }
            return;
        }
        const bool fused { reduce && world > 1 && peer.active && V <= kPeerMaxDoubles };
        constexpr int chunk { 64 };
        const int blocks { (n_local_sets + chunk - 1) / chunk };
        const int threads { std::min(256, ((V + 31) / 32) * 32) };
        const double* last_in { contrib };
        int last_rows { n_local_sets };
        if (blocks > 1) {
            d_stage.ensure(std::max(d_stage.capacity, static_cast<size_t>(blocks) * V));
            k_column_sums<<<blocks, threads, 0, stream>>>(
              contrib, n_local_sets, V, chunk, d_stage.ptr);
            launched();
            last_in = d_stage.ptr;
            last_rows = blocks;
        }
        if (fused) {
            peer.column_sums_all_reduce(last_in, last_rows, V, dst, threads, stream);
            launched();
            ++n_peer_reductions;
            return;
        }
        k_column_sums<<<1, threads, 0, stream>>>(last_in, last_rows, V, last_rows, dst);
        launched();
        if (reduce) {
            all_reduce(dst, V);
        }
    }

    [[nodiscard]] auto shared_len() const -> int { return pg * (pg + 1) / 2 + pg; }

    // computeLeftHandSide + computeRightHandSide (lm_solver.cpp:286-346)
    auto jacobian_phase(const bool store_rows, const bool store_residuals) -> void
    {
        const int NA { NP + 1 };
        const int NT { NA * (NA + 1) / 2 };
        t_jacobian.start(stream);
        const size_t n_par { static_cast<size_t>(n_local_sets) * NP };
        if (n_par > 0) {
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_par_jac.ptr,
                                            d_par.ptr,
                                            n_par * sizeof(double),
                                            cudaMemcpyDeviceToDevice,
                                            stream));
        }
        if (n_tiles > 0) {
            if (store_rows) {
                d_J.ensure(static_cast<size_t>(NP) * n_local_points);
            }
            if (store_residuals) {
                d_resid.ensure(n_local_points);
            }
            t_jac_kernel.start(stream);
            ops->jacobian(slots_jac.grid,
                          stream,
                          sweep_args(&slots_jac.cta),
                          settings.loss,
                          slots_jac.base.ptr,
                          d_jac_partials.ptr,
                          store_rows ? d_J.ptr : nullptr,
                          store_residuals ? d_resid.ptr : nullptr,
                          n_local_points);
            t_jac_kernel.stop(stream);
            launched();
            k_reduce_slots<<<blocks_for(static_cast<long long>(n_local_sets) * NT, 256),
                             256,
                             0,
                             stream>>>(d_jac_partials.ptr,
                                       slots_jac.first.ptr,
                                       slots_jac.count.ptr,
                                       n_local_sets,
                                       NT,
                                       d_M.ptr);
            launched();
        }
        if (part.split) {
            if (n_tiles == 0) {
                // a rank without points (every data set shorter than the number
                // of ranks) contributes zeros, not last iteration's global sums
                d_M.zero(static_cast<size_t>(std::max(n_local_sets, 1)) * NT, stream);
            }
            all_reduce(d_M.ptr, static_cast<size_t>(n_local_sets) * NT);
        }
        t_jacobian.stop(stream);

        t_linalg.start(stream);
        const int VG { shared_len() };
        d_contrib.ensure(static_cast<size_t>(std::max(n_local_sets, 1)) * std::max(VG, 2));
        d_shared_sums.ensure(std::max(VG, 1));
        if (n_local_sets > 0) {
            ops->local_prepare(stream,
                               d_M.ptr,
                               d_active.ptr,
                               global_mask,
                               n_local_sets,
                               settings.damp_max,
                               d_dtd_local.ptr,
                               d_contrib.ptr,
                               VG);
            launched();
        }
        column_sums(d_contrib.ptr, VG, d_shared_sums.ptr, !part.split);
        t_linalg.stop(stream);
    }

    // computeDeltas (lm_solver.cpp:348-399) followed by updateParameters
    // delta1_only: the Fortran driver's re-solve after a rejected step
    // (gadfit.F90:801-808): no omega sweep, no delta2, update with delta1 alone
    auto solve_phase(const bool first_trial, const bool delta1_only = false) -> void
    {
        const int VG { shared_len() };
        const bool accelerate { !delta1_only && settings.acceleration_threshold > 0.0 };
        t_linalg.start(stream);
        // No data set has a local parameter (n = p_g, e.g. BASELINE config 2):
        // the Schur complement is empty on every rank, so its kernels and its
        // all-reduce are skipped and the (zeroed) sums are used as they are
        const bool has_locals { n_active > pg };
        if (!has_locals) {
            if (first_trial) {
                d_schur_sums.zero(std::max(VG, 1), stream);
            }
        } else {
            d_schur_sums.ensure(std::max(VG, 1));
            if (n_local_sets > 0) {
                ops->local_schur(stream,
                                 d_M.ptr,
                                 d_active.ptr,
                                 global_mask,
                                 n_local_sets,
                                 d_dtd_local.ptr,
                                 d_state.ptr,
                                 d_factors.ptr,
                                 d_contrib.ptr,
                                 VG);
                launched();
            }
            column_sums(d_contrib.ptr, VG, d_schur_sums.ptr, !part.split);
        }
        if (pg > 0) {
            k_shared_solve<<<1, 32, 0, stream>>>(pg,
                                                 1,
                                                 first_trial ? 1 : 0,
                                                 settings.damp_max,
                                                 d_shared_sums.ptr,
                                                 d_schur_sums.ptr,
                                                 nullptr,
                                                 d_state.ptr,
                                                 d_dtd_global.ptr,
                                                 d_chol.ptr,
                                                 d_delta1_g.ptr);
            launched();
        }
        if (n_local_sets > 0) {
            ops->local_backsolve(stream,
                                 d_active.ptr,
                                 global_mask,
                                 n_local_sets,
                                 d_factors.ptr,
                                 d_delta1_g.ptr,
                                 d_delta1_par.ptr);
            launched();
        }
        t_linalg.stop(stream);

        if (accelerate) {
            // omega sweep along delta1 (lm_solver.cpp:361-381)
            t_omega.start(stream);
            if (n_tiles > 0) {
                ops->omega(slots_omega.grid,
                           stream,
                           sweep_args(&slots_omega.cta),
                           d_delta1_par.ptr,
                           d_J.ptr,
                           n_local_points,
                           slots_omega.base.ptr,
                           d_omega_partials.ptr,
                           nullptr);
                launched();
                k_reduce_slots<<<blocks_for(static_cast<long long>(n_local_sets) * NP, 256),
                                 256,
                                 0,
                                 stream>>>(d_omega_partials.ptr,
                                           slots_omega.first.ptr,
                                           slots_omega.count.ptr,
                                           n_local_sets,
                                           NP,
                                           d_vsets.ptr);
                launched();
            }
            if (part.split) {
                if (n_tiles == 0) {
                    d_vsets.zero(static_cast<size_t>(std::max(n_local_sets, 1)) * NP, stream);
                }
                all_reduce(d_vsets.ptr, static_cast<size_t>(n_local_sets) * NP);
            }
            t_omega.stop(stream);

            // delta2 with the same factor (lm_solver.cpp:383-384)
            t_linalg.start(stream);
            d_rhs2_sums.ensure(std::max(2 * pg, 1));
            d_ratio_sums.ensure(2);
            if (n_local_sets > 0) {
                ops->local_schur_rhs(stream,
                                     d_vsets.ptr,
                                     d_active.ptr,
                                     global_mask,
                                     n_local_sets,
                                     d_factors.ptr,
                                     d_contrib.ptr,
                                     2 * pg);
                launched();
            }
            column_sums(d_contrib.ptr, 2 * pg, d_rhs2_sums.ptr, !part.split);
            if (pg > 0) {
                k_shared_solve<<<1, 32, 0, stream>>>(pg,
                                                     2,
                                                     0,
                                                     settings.damp_max,
                                                     d_shared_sums.ptr,
                                                     d_schur_sums.ptr,
                                                     d_rhs2_sums.ptr,
                                                     d_state.ptr,
                                                     d_dtd_global.ptr,
                                                     d_chol.ptr,
                                                     d_delta2_g.ptr);
                launched();
            }
            if (n_local_sets > 0) {
                ops->local_backsolve(stream,
                                     d_active.ptr,
                                     global_mask,
                                     n_local_sets,
                                     d_factors.ptr,
                                     d_delta2_g.ptr,
                                     d_delta2_par.ptr);
                ops->acc_ratio_local(stream,
                                     d_active.ptr,
                                     global_mask,
                                     n_local_sets,
                                     d_dtd_local.ptr,
                                     d_delta1_par.ptr,
                                     d_delta2_par.ptr,
                                     d_contrib.ptr);
                launched(2);
            }
            column_sums(d_contrib.ptr, 2, d_ratio_sums.ptr, !part.split);
            k_acc_ratio_final<<<1, 32, 0, stream>>>(pg,
                                                    settings.acceleration_threshold,
                                                    d_ratio_sums.ptr,
                                                    d_dtd_global.ptr,
                                                    d_delta1_g.ptr,
                                                    d_delta2_g.ptr,
                                                    d_state.ptr);
            launched();
            t_linalg.stop(stream);
            have_delta2 = true;
        }

        // saveParameters (first trial only) + updateParameters
        const long long n_par { static_cast<long long>(n_local_sets) * NP };
        if (n_par > 0) {
            k_update_parameters<<<blocks_for(n_par, 256), 256, 0, stream>>>(
              n_par,
              first_trial ? 1 : 0,
              accelerate ? 1 : 0,
              d_state.ptr,
              d_delta1_par.ptr,
              d_delta2_par.ptr,
              d_par.ptr,
              d_par_old.ptr);
            launched();
        }
    }

    // ---- latency-bound regime (block_arrow_kernels.cuh, "fused kernels") ----

    // One GPU, at most kSmallMaxSets data sets, no acceleration: a trial step
    // is Jacobian sweep | k_solve_small | chi2 sweep | k_finish_trial.
    // GADFIT_B200_FUSED=0 keeps the one-kernel-per-stage path (A/B runs, and the
    // parity test that both paths give bit-identical results).
    [[nodiscard]] auto small_path() const -> bool
    {
        const char* e { std::getenv("GADFIT_B200_FUSED") }; // read at every fit()
        const bool allowed { e == nullptr || std::strcmp(e, "0") != 0 };
        return allowed && world == 1 && !part.split && n_tiles > 0 && n_local_sets > 0
               && n_local_sets <= kSmallMaxSets && settings.acceleration_threshold <= 0.0;
    }

    // The Jacobian sweep alone; its partial slots are reduced by k_solve_small
    auto jacobian_sweep_small() -> void
    {
        t_jacobian.start(stream);
        ops->jacobian(slots_jac.grid,
                      stream,
                      sweep_args(&slots_jac.cta),
                      settings.loss,
                      slots_jac.base.ptr,
                      d_jac_partials.ptr,
                      nullptr,
                      nullptr,
                      n_local_points);
        t_jacobian.stop(stream);
        launched();
    }

    // Work arrays of k_solve_small (once per iteration; once per fit in the
    // device-driven loop)
    auto prepare_small_buffers() -> void
    {
        const int VG { shared_len() };
        const int chunks { (n_local_sets + kColumnChunk - 1) / kColumnChunk };
        d_contrib.ensure(2 * static_cast<size_t>(n_local_sets) * std::max(VG, 2));
        d_stage.ensure(
          std::max(d_stage.capacity, 2 * static_cast<size_t>(chunks) * std::max(VG, 1)));
        d_shared_sums.ensure(std::max(VG, 1));
        if (n_active > pg) {
            d_schur_sums.ensure(std::max(VG, 1));
        } else {
            d_schur_sums.zero(std::max(VG, 1), stream);
        }
    }

    auto solve_small(const bool first_trial, const bool gated = false) -> void
    {
        const int VG { shared_len() };
        const int chunks { (n_local_sets + kColumnChunk - 1) / kColumnChunk };
        t_linalg.start(stream);
        if (first_trial && !gated) {
            prepare_small_buffers();
        }
        SmallSolveArgs a {};
        a.n_sets = n_local_sets;
        a.pg = pg;
        a.VG = VG;
        a.global_mask = global_mask;
        a.prepare = first_trial ? 1 : 0;
        a.has_locals = n_active > pg ? 1 : 0;
        a.update = 1;
        a.damp_max = settings.damp_max;
        a.gated = gated ? 1 : 0;
        a.jac_partials = d_jac_partials.ptr;
        a.slot_first = slots_jac.first.ptr;
        a.slot_count = slots_jac.count.ptr;
        a.active = d_active.ptr;
        a.M = d_M.ptr;
        a.dtd_local = d_dtd_local.ptr;
        a.contrib = d_contrib.ptr;
        a.stage = d_stage.ptr;
        a.shared_sums = d_shared_sums.ptr;
        a.schur_sums = d_schur_sums.ptr;
        a.state = d_state.ptr;
        a.dtd_global = d_dtd_global.ptr;
        a.chol = d_chol.ptr;
        a.delta_g = d_delta1_g.ptr;
        a.factors = d_factors.ptr;
        a.delta_par = d_delta1_par.ptr;
        a.par = d_par.ptr;
        a.par_old = d_par_old.ptr;
        a.par_jac = d_par_jac.ptr;
        ops->solve_small(stream, a);
        launched();
        t_linalg.stop(stream);
    }

    // chi2 sweep of the trial step, then the last stage of its sum, the
    // accept / reject decision and the revert in one launch
    auto chi2_decide_small() -> void
    {
        t_chi2.start(stream);
        ops->chi2(grid_chi2, stream, sweep_args(&cta_chi2), d_chi2_partials.ptr, TrialFinish {});
        k_finish_trial<<<1, kThreads, 0, stream>>>(d_chi2_partials.ptr,
                                                   grid_chi2,
                                                   settings.lambda_incs,
                                                   settings.lambda_down,
                                                   settings.lambda_up,
                                                   d_state.ptr,
                                                   static_cast<long long>(n_local_sets) * NP,
                                                   d_par_old.ptr,
                                                   d_par.ptr);
        launched(2);
        t_chi2.stop(stream);
    }

    // ---- device-driven loop (latency-bound regime) ---------------------------
    //
    // The host-driven loop reads one small record per trial step to know what
    // to enqueue next: a stream synchronisation (~15 us) and an idle GPU until
    // the next launch arrives, per trial.  Here the decision stays on the
    // device: a TRIAL SLOT is the fixed launch sequence
    //     Jacobian sweep | k_solve_small | chi2 sweep (+ finish in its last CTA)
    // whose kernels look at LmState::phase and skip themselves (Jacobian sweep
    // unless a new iteration starts, everything once the fit is done), the
    // last CTA of the chi2 sweep takes the accept / reject decision, reverts
    // the parameters, writes the trace record and sets the phase of the next
    // slot.  The host enqueues as many slots as there are iterations left (the
    // number needed if every step is accepted) and only then reads the state.
    // Same kernels, same sums: results are bit-identical to the host-driven
    // loop (GADFIT_B200_DEVICE_LOOP=0).
    [[nodiscard]] auto device_loop_allowed() const -> bool
    {
        const char* e { std::getenv("GADFIT_B200_DEVICE_LOOP") };
        if (e != nullptr && std::strcmp(e, "0") == 0) {
            return false;
        }
        // per-iteration text output needs the parameters on the host after
        // every accepted step: host-driven loop
        const bool chatty { log_fn != nullptr && !io_test(io_hide_all) && !io_test(io_final_only) };
        return !chatty;
    }

    auto enqueue_trial_slot(const bool skip_jacobian = false) -> void
    {
        SweepArgs sweep { sweep_args(&slots_jac.cta) };
        sweep.gate = d_state.ptr;
        if (!skip_jacobian) {
            // Jacobian sweep: only when a new iteration starts
            sweep.gate_skip = (1 << kPhaseRetry) | (1 << kPhaseDone);
            t_jacobian.start(stream);
            ops->jacobian(slots_jac.grid, stream, sweep, settings.loss, slots_jac.base.ptr,
                          d_jac_partials.ptr, nullptr, nullptr, n_local_points);
            t_jacobian.stop(stream);
            launched();
        }
        sweep = sweep_args(&cta_chi2);
        sweep.gate = d_state.ptr;
        solve_small(true, true);
        sweep.gate_skip = 1 << kPhaseDone;
        TrialFinish fin {};
        fin.state = d_state.ptr;
        fin.counter = d_chi2_counter.ptr;
        fin.lambda_incs = settings.lambda_incs;
        fin.iteration_limit = settings.iteration_limit;
        fin.lambda_down = settings.lambda_down;
        fin.lambda_up = settings.lambda_up;
        fin.n_par = static_cast<long long>(n_local_sets) * NP;
        fin.par_old = d_par_old.ptr;
        fin.par = d_par.ptr;
        fin.trace = d_trace.ptr;
        fin.trace_cap = static_cast<int>(d_trace.capacity);
        t_chi2.start(stream);
        ops->chi2(grid_chi2, stream, sweep, d_chi2_partials.ptr, fin);
        t_chi2.stop(stream);
        launched();
    }

    // Returns the final state; fills `trace`
    // first_jacobian_done: the sweep of the first slot already ran (pipelined upload)
    auto run_device_loop(const bool first_jacobian_done) -> LmState
    {
        const long long worst { static_cast<long long>(settings.iteration_limit)
                                * (settings.lambda_incs + 1) };
        d_trace.ensure(static_cast<size_t>(std::min<long long>(std::max<long long>(worst, 1), 1 << 20)));
        prepare_small_buffers();
        int n_read { 0 };
        LmState now {};
        int iterations_left { settings.iteration_limit };
        while (true) {
            const int n_slots { std::max(1, std::min(iterations_left, kTraceBatch)) };
            for (int k = 0; k < n_slots; ++k) {
                enqueue_trial_slot(first_jacobian_done && n_read == 0 && trace.empty() && k == 0);
            }
            GADFIT_CUDA_TRY(cudaMemcpyAsync(
              h_mirror, d_state.ptr, sizeof(LmState), cudaMemcpyDeviceToHost, stream));
            const int n_copy { static_cast<int>(std::min<size_t>(
              n_slots, d_trace.capacity - std::min<size_t>(n_read, d_trace.capacity))) };
            if (n_copy > 0) {
                GADFIT_CUDA_TRY(cudaMemcpyAsync(h_trace,
                                                d_trace.ptr + n_read,
                                                n_copy * sizeof(gadfit_b200_trial),
                                                cudaMemcpyDeviceToHost,
                                                stream));
            }
            GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
            now = h_mirror[0];
            const int n_new { std::min(now.n_trials - n_read, n_copy) };
            for (int k = 0; k < n_new; ++k) {
                trace.push_back(h_trace[k]);
            }
            n_read += std::max(n_new, 0);
            if (now.flags != 0U || now.phase == kPhaseDone) {
                break;
            }
            iterations_left = settings.iteration_limit - (now.iteration - 1);
        }
        return now;
    }

    // The initial values of the damping matrix (lm_solver.cpp:414,424-428)
    auto init_damping() -> void
    {
        const size_t n_loc { static_cast<size_t>(std::max(n_local_sets, 1)) * NP };
        std::vector<double> dtd_local(n_loc, 0.0);
        std::vector<double> dtd_global(std::max(pg, 1), 0.0);
        if (DTD_min.size() > 1) {
            if (static_cast<long long>(DTD_min.size()) < n_active) {
                throw StatusError { GADFIT_B200_INVALID_ARGUMENT,
                                    "DTD_min is shorter than the number of active "
                                    "parameters" };
            }
            for (int a = 0; a < pg; ++a) {
                dtd_global[a] = DTD_min[a];
            }
            for (int ls = 0; ls < n_local_sets; ++ls) {
                const int s { part.first + ls };
                for (int k = 0; k < NP; ++k) {
                    const long long col { column(s, k) };
                    if (col >= pg) {
                        dtd_local[static_cast<size_t>(ls) * NP + k] = DTD_min[col];
                    }
                }
            }
        }
        d_dtd_local.upload(dtd_local, stream);
        d_dtd_global.upload(dtd_global, stream);
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
    }

    auto reset_timers() -> void
    {
        for (PhaseTimer* t :
             { &t_jacobian, &t_chi2, &t_linalg, &t_omega, &t_main, &t_jac_kernel }) {
            t->reset();
        }
    }

    // ---- text output, lm_solver.cpp:587-686 ---------------------------------
    // Emitted line by line to the sink installed with gadfit_b200_set_log (the
    // library itself never writes to stdout); which lines appear is decided
    // by settings.verbosity exactly as in the reference.
    enum : unsigned
    {
        io_all = 1U,
        io_delta1 = 2U,
        io_delta2 = 4U,
        io_timings = 8U,
        io_hide_local = 16U,
        io_hide_global = 32U,
        io_hide_all = 64U,
        io_final_only = 128U
    };
    [[nodiscard]] auto io_test(const unsigned flag) const -> bool // lm_solver.cpp:682-686
    {
        return (settings.verbosity & io_all) != 0U || (settings.verbosity & flag) != 0U;
    }
    auto log_line(const std::string& line) const -> void
    {
        if (log_fn != nullptr) {
            log_fn(line.c_str(), log_user);
        }
    }
    // "{}" of a double in the reference's spdlog/fmt: shortest round-trip form
    static auto shortest(const double v) -> std::string
    {
        char buf[40];
        const auto res { std::to_chars(buf, buf + sizeof(buf), v) };
        return std::string { buf, res.ptr };
    }
    // delta1 / delta2 as dense vectors over the Jacobian columns
    auto dense_delta(const bool second) -> std::vector<double>
    {
        std::vector<double> out(static_cast<size_t>(n_active), 0.0);
        const int S { static_cast<int>(sets.size()) };
        if (world > 1 && !part.split) {
            return out; // remote local blocks are not gathered for printing
        }
        const auto d { download(second ? d_delta2_par.ptr : d_delta1_par.ptr,
                                static_cast<size_t>(S) * NP) };
        for (int set = 0; set < S; ++set) {
            for (int k = 0; k < NP; ++k) {
                const long long c { column(set, k) };
                if (c >= 0) {
                    out[c] = d[static_cast<size_t>(set) * NP + k];
                }
            }
        }
        return out;
    }
    // printIterationResults, lm_solver.cpp:587-654
    auto print_iteration_results(const int i_iteration, const double lambda, const double chi2)
      -> void
    {
        if (log_fn == nullptr) {
            return;
        }
        download_parameters();
        const bool show_d1 { io_test(io_delta1) };
        const bool show_d2 { io_test(io_delta2) && have_delta2 };
        const auto delta1 { show_d1 ? dense_delta(false) : std::vector<double> {} };
        auto delta2 { show_d2 ? dense_delta(true) : std::vector<double> {} };
        if (show_d2 && download(d_state.ptr, 1).front().use_delta2 == 0) {
            std::fill(delta2.begin(), delta2.end(), 0.0); // lm_solver.cpp:394-396
        }
        log_line("Iteration: " + std::to_string(i_iteration));
        log_line("Lambda: " + shortest(lambda));
        log_line("Chi2/DOF: " + shortest(chi2 / dof));
        const int S { static_cast<int>(sets.size()) };
        const auto parameter_line { [&](const int i_set, const int i_par) {
            std::ostringstream line {};
            line << std::setprecision(std::numeric_limits<double>::digits10);
            if (i_par < static_cast<int>(names.size()) && !names[i_par].empty()) {
                line << std::setw(15) << names[i_par] << ": ";
            } else {
                line << "    Parameter " << i_par << ": ";
            }
            line << par[static_cast<size_t>(i_set) * NP + i_par];
            const long long c { column(i_set, i_par) };
            if (c >= 0) {
                if (show_d1) {
                    line << " (" << delta1[c] << ')';
                }
                if (io_test(io_delta2) && !delta2.empty()) {
                    line << " (" << delta2[c] << ')';
                }
            } else {
                line << " (fixed)";
            }
            log_line(line.str());
        } };
        const bool single_dataset { S == 1 };
        const auto is_global { [&](const int i_par) { // Indices::global
            return ((global_mask >> i_par) & 1U) != 0U;
        } };
        if (!single_dataset && !io_test(io_hide_global)) {
            log_line("  Global parameters");
            for (int i_par = 0; i_par < NP; ++i_par) {
                if (is_global(i_par)) {
                    parameter_line(0, i_par);
                }
            }
        }
        if (!io_test(io_hide_local)) {
            for (int i_set = 0; i_set < S; ++i_set) {
                if (S > 1) {
                    log_line("  Data set: " + std::to_string(i_set));
                }
                for (int i_par = 0; i_par < NP; ++i_par) {
                    if (single_dataset || !is_global(i_par)) {
                        parameter_line(i_set, i_par);
                    }
                }
            }
        }
        log_line("");
    }
    // printTimings, lm_solver.cpp:656-680 (device seconds in both time columns)
    auto print_timings() -> void
    {
        if (log_fn == nullptr) {
            return;
        }
        const auto row { [&](const char* term, const PhaseTimer& t) {
            char buf[96];
            std::snprintf(buf, sizeof(buf), "%-14s %10.2f %10.2f  %6.2f%% %5d", term, t.seconds,
                          t.seconds, 100 * t.seconds / t_main.seconds, t.calls);
            log_line(buf);
        } };
        log_line("");
        log_line("Timings          Wall (s)    CPU (s)  CPU rel  Calls");
        log_line("====================================================");
        row("Jacobian", t_jacobian);
        row("Chi2", t_chi2);
        row("Linear algebra", t_linalg);
        row("Omega", t_omega);
        log_line("----------------------------------------------------");
        row("Main loop", t_main);
        log_line("====================================================");
        log_line("");
    }

    // Status bits the kernels raised (LmState::flags) -> the status the
    // reference reports by exception (numerical_integration.cpp:304,525:
    // InsufficientIntegrationWorkspace) / a collective that timed out
    auto throw_if_flagged(const unsigned flags) -> void
    {
        if ((flags & kFlagPeerTimeout) != 0U) {
            throw StatusError { GADFIT_B200_NCCL_ERROR,
                                "a peer-memory all-reduce timed out: another rank did not "
                                "take part in the collective" };
        }
        if ((flags & kFlagWorkspace) != 0U) {
            throw StatusError {
                GADFIT_B200_INSUFFICIENT_INTEGRATION_WORKSPACE,
                "an adaptive integral needed more than " + std::to_string(k_max_panels)
                  + " panels (1-D) / " + std::to_string(k_max_panels_2d)
                  + " rectangles (2-D) at some data point: loosen rel_error / abs_error of "
                    "integrate() or narrow the range"
            };
        }
    }
    // Reads the live flags (one 48-byte copy + sync) and clears them
    auto check_device_flags() -> void
    {
        GADFIT_CUDA_TRY(cudaMemcpyAsync(
          h_mirror, d_state.ptr, sizeof(LmState), cudaMemcpyDeviceToHost, stream));
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        const unsigned flags { h_mirror[0].flags };
        if (flags != 0U) {
            GADFIT_CUDA_TRY(cudaMemsetAsync(&d_state.ptr->flags, 0, sizeof(unsigned), stream));
            throw_if_flagged(flags);
        }
    }

    // LMsolver::fit, lm_solver.cpp:401-511
    auto fit(const double lambda0) -> void
    {
        if (gadf.enabled != 0) {
            fit_gadf(lambda0);
            return;
        }
        stop_reason = GADFIT_B200_STOP_ITERATION_LIMIT;
        {
            const HostPhase phase { "prepare_indexing" };
            prepare_indexing();
        }
        ensure_device();
        // Data not in HBM yet: the copies are enqueued chunk by chunk on a copy
        // stream and not waited for; the FIRST Jacobian sweep (launched ahead of
        // the initial chi2 sweep: both only read the start parameters) follows
        // their arrival, so that sweep and all the set-up below overlap the
        // PCIe transfer.
        const bool fresh_upload { !data_uploaded && settings.iteration_limit != 0 };
        {
            const HostPhase phase { "upload_parameters" };
            upload_parameters(fresh_upload);
        }
        reset_timers();
        {
            const HostPhase phase { "init_damping" };
            init_damping();
        }
        const size_t n_par { static_cast<size_t>(std::max(n_local_sets, 1)) * NP };
        d_delta1_par.zero(n_par, stream);
        d_delta2_par.zero(n_par, stream);
        d_delta1_g.zero(kMaxShared, stream);
        d_delta2_g.zero(kMaxShared, stream);
        d_chol.ensure(kMaxShared * kMaxShared);
        d_factors.ensure(static_cast<size_t>(std::max(n_local_sets, 1)) * ops->factor_bytes());
        const bool accelerate { settings.acceleration_threshold > 0.0 };
        have_delta2 = false;
        trace.clear();
        n_iterations = 0;
        struct UploadGuard // an exception before the first sweep: never leave copies unwaited
        {
            SolverImpl* self;
            ~UploadGuard()
            {
                if (self->upload_in_flight) {
                    cudaStreamSynchronize(self->copy_stream);
                    self->upload_in_flight = false;
                    self->data_uploaded = false; // sigma was not inverted: upload again
                }
            }
        } upload_guard { this };

        LmState init {};
        init.lambda = lambda0;
        init.use_delta2 = 0;
        init.phase = kPhaseNewIteration;
        init.iteration = 1;
        GADFIT_CUDA_TRY(cudaMemcpyAsync(
          d_state.ptr, &init, sizeof(LmState), cudaMemcpyHostToDevice, stream));
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        lambda_lhs = lambda0;
        double lambda { lambda0 };
        double last_old_chi2 { 0.0 };

        const bool small { small_path() };
        t_main.start(stream);
        bool first_jacobian_done { false };
        if (fresh_upload) {
            small ? jacobian_sweep_small() : jacobian_phase(accelerate, false);
            finish_pipelined_upload();
            first_jacobian_done = true;
        }
        chi2_phase(&d_state.ptr->new_chi2);
        k_copy_chi2_to_old<<<1, 1, 0, stream>>>(d_state.ptr);
        launched();
        int i_iteration { 0 };
        bool finished { settings.iteration_limit == 0 };
        bool finite_start { true };
        int device_loop_trials { -1 };
        int device_loop_jacobians { 0 };
        if (small && !finished && device_loop_allowed()) {
            const LmState end { run_device_loop(first_jacobian_done) };
            device_loop_trials = end.n_trials;
            device_loop_jacobians = end.n_jacobians;
            i_iteration = end.iteration;
            lambda = end.lambda;
            last_old_chi2 = end.old_chi2;
            lambda_lhs = trace.empty() ? lambda0 : trace.back().lambda;
            if (end.flags != 0U) {
                t_main.stop(stream);
                n_iterations = std::max(i_iteration - 1, 0);
                download_parameters();
                fitted = true;
                throw_if_flagged(end.flags);
            }
            if (!trace.empty() && trace.back().decision == 2) { // lm_solver.cpp:482-494
                --i_iteration;
                stop_reason = std::isfinite(trace.back().old_chi2)
                                ? GADFIT_B200_STOP_LAMBDA_INCS
                                : GADFIT_B200_STOP_NON_FINITE_CHI2;
            } else if (!io_test(io_hide_all)) { // lm_solver.cpp:497-499
                log_line("Iteration limit reached");
            }
            finished = true;
        }
        while (!finished) {
            ++i_iteration;
            if (first_jacobian_done && i_iteration == 1) {
                // already swept (ahead of the initial chi2, following the upload)
            } else if (small) {
                jacobian_sweep_small();
            } else {
                jacobian_phase(accelerate, false);
            }
            lambda_lhs = lambda;
            small ? solve_small(true) : solve_phase(true);
            for (int i_lambda = 0; i_lambda <= settings.lambda_incs; ++i_lambda) {
                if (small) {
                    chi2_decide_small();
                } else {
                    chi2_phase(&d_state.ptr->new_chi2);
                    k_decide<<<1, 1, 0, stream>>>(settings.lambda_incs,
                                                  settings.lambda_down,
                                                  settings.lambda_up,
                                                  d_state.ptr,
                                                  d_state.ptr + 1);
                    const long long n { static_cast<long long>(n_local_sets) * NP };
                    if (n > 0) {
                        k_revert_if_rejected<<<blocks_for(n, 256), 256, 0, stream>>>(
                          n, d_state.ptr, d_par_old.ptr, d_par.ptr);
                        launched();
                    }
                    launched();
                }
                // One small record per trial step: [1] = state before the
                // decision (+ the decision), [0] = state after it
                GADFIT_CUDA_TRY(cudaMemcpyAsync(h_mirror,
                                                d_state.ptr,
                                                2 * sizeof(LmState),
                                                cudaMemcpyDeviceToHost,
                                                stream));
                GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
                const LmState& now { h_mirror[0] };
                const LmState& rec { h_mirror[1] };
                trace.push_back(
                  { rec.lambda, rec.new_chi2, rec.old_chi2, i_iteration, rec.decision });
                lambda = now.lambda;
                last_old_chi2 = now.old_chi2;
                if (now.flags != 0U) {
                    // as the reference's exception: the fit ends here; what has
                    // been accepted so far stays in the solver
                    t_main.stop(stream);
                    n_iterations = std::max(i_iteration - 1, 0);
                    download_parameters();
                    fitted = true;
                    throw_if_flagged(now.flags);
                }
                const bool chatty { !io_test(io_hide_all) && !io_test(io_final_only) };
                if (rec.decision == 0) {
                    if (chatty) { // lm_solver.cpp:465-467
                        print_iteration_results(i_iteration, lambda, rec.new_chi2);
                    }
                    break;
                }
                if (rec.decision == 1) {
                    lambda_lhs = lambda;
                    small ? solve_small(false) : solve_phase(false);
                } else {
                    if (chatty) { // lm_solver.cpp:484-488
                        log_line("Lambda increased " + std::to_string(settings.lambda_incs)
                                 + " times in a row");
                        log_line("");
                    }
                    --i_iteration;
                    finished = true;
                    // NaN/inf at the start: every step is rejected (new < old is
                    // false, as in the reference) -- tell it apart from a
                    // converged fit
                    finite_start = finite_start && std::isfinite(rec.old_chi2);
                    stop_reason = finite_start ? GADFIT_B200_STOP_LAMBDA_INCS
                                               : GADFIT_B200_STOP_NON_FINITE_CHI2;
                }
            }
            if (i_iteration == settings.iteration_limit) {
                if (!io_test(io_hide_all)) { // lm_solver.cpp:497-499
                    log_line("Iteration limit reached");
                }
                finished = true;
            }
        }
        t_main.stop(stream);
        n_iterations = i_iteration;
        download_parameters();
        if (settings.iteration_limit == 0) {
            check_device_flags(); // the initial chi2 sweep is all that ran
        }
        for (PhaseTimer* t :
             { &t_jacobian, &t_chi2, &t_linalg, &t_omega, &t_main, &t_jac_kernel }) {
            t->collect();
        }
        if (device_loop_trials >= 0) {
            // slots that skipped themselves were timed too (a few microseconds
            // each); the call counts are those of the work that ran
            t_jacobian.calls = device_loop_jacobians;
            t_linalg.calls = device_loop_trials;
            t_chi2.calls = device_loop_trials + 1;
        }
        fitted = true;
        // lm_solver.cpp:504-510
        if ((!io_test(io_hide_all) && io_test(io_final_only)) || settings.iteration_limit == 0) {
            if (settings.iteration_limit == 0 && log_fn != nullptr) {
                GADFIT_CUDA_TRY(cudaMemcpyAsync(h_mirror, d_state.ptr, sizeof(LmState),
                                                cudaMemcpyDeviceToHost, stream));
                GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
                last_old_chi2 = h_mirror[0].old_chi2;
            }
            print_iteration_results(i_iteration, lambda, last_old_chi2);
        }
        if (!io_test(io_hide_all) && io_test(io_timings)) {
            print_timings();
        }
    }

    // gadf_fit, fortran/gadfit/gadfit.F90:678-915: the same sweeps and solves,
    // driven by the Fortran version's accept test, lambda updates, re-solve
    // rule and convergence tests.  The scalars they need are reduced on the
    // device (k_gadf_local / k_gadf_global); the decisions are taken on the
    // host, which reads one small record per trial step anyway.
    auto fit_gadf(const double lambda0) -> void
    {
        if (settings.loss != GADFIT_B200_LOSS_LINEAR) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "gadf_fit has no loss functions" };
        }
        prepare_indexing();
        ensure_device();
        upload_parameters();
        reset_timers();
        init_damping();
        const size_t n_par { static_cast<size_t>(std::max(n_local_sets, 1)) * NP };
        d_delta1_par.zero(n_par, stream);
        d_delta2_par.zero(n_par, stream);
        d_old_delta1_par.zero(n_par, stream);
        d_delta1_g.zero(kMaxShared, stream);
        d_delta2_g.zero(kMaxShared, stream);
        d_old_delta1_g.zero(kMaxShared, stream);
        d_gadf_sums.zero(kGadfSums + 1, stream);
        d_chol.ensure(kMaxShared * kMaxShared);
        d_factors.ensure(static_cast<size_t>(std::max(n_local_sets, 1)) * ops->factor_bytes());
        const bool accelerate { settings.acceleration_threshold > 0.0 };
        have_delta2 = false;
        trace.clear();
        n_iterations = 0;
        stop_reason = GADFIT_B200_STOP_ITERATION_LIMIT;

        LmState init {};
        init.lambda = lambda0;
        GADFIT_CUDA_TRY(cudaMemcpyAsync(
          d_state.ptr, &init, sizeof(LmState), cudaMemcpyHostToDevice, stream));
        double lambda { lambda0 };
        const auto set_lambda { [&](const double value) {
            lambda = value;
            h_scalar[0] = value;
            GADFIT_CUDA_TRY(cudaMemcpyAsync(&d_state.ptr->lambda,
                                            h_scalar,
                                            sizeof(double),
                                            cudaMemcpyHostToDevice,
                                            stream));
            GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        } };
        const auto revert { [&] {
            if (n_local_sets > 0) {
                GADFIT_CUDA_TRY(cudaMemcpyAsync(d_par.ptr,
                                                d_par_old.ptr,
                                                static_cast<size_t>(n_local_sets) * NP
                                                  * sizeof(double),
                                                cudaMemcpyDeviceToDevice,
                                                stream));
            }
        } };
        // chi2 of the trial parameters + the reduced scalars -> host
        double* const h_gadf { h_scalar + 8 }; // pinned
        double sums[kGadfSums + 1] {};
        const auto measure { [&] {
            chi2_phase(d_gadf_sums.ptr + kGadfSums);
            t_linalg.start(stream);
            d_contrib.ensure(static_cast<size_t>(std::max(n_local_sets, 1))
                             * std::max(shared_len(), kGadfSums));
            if (n_local_sets > 0) {
                k_gadf_local<<<blocks_for(n_local_sets, 128), 128, 0, stream>>>(
                  NP,
                  d_M.ptr,
                  d_active.ptr,
                  global_mask,
                  n_local_sets,
                  d_dtd_local.ptr,
                  d_delta1_par.ptr,
                  d_old_delta1_par.ptr,
                  d_par.ptr,
                  gadf.rel_error,
                  d_contrib.ptr);
                launched();
            }
            column_sums(d_contrib.ptr, kGadfSums, d_gadf_sums.ptr);
            const bool owns_set0 { n_local_sets > 0 && part.first == 0 };
            if (pg > 0 && owns_set0) {
                k_gadf_global<<<1, 1, 0, stream>>>(NP,
                                                   pg,
                                                   global_mask,
                                                   d_shared_sums.ptr,
                                                   d_dtd_global.ptr,
                                                   d_delta1_g.ptr,
                                                   d_old_delta1_g.ptr,
                                                   d_par.ptr,
                                                   gadf.rel_error,
                                                   gadf.rel_error_global,
                                                   d_gadf_sums.ptr);
                launched();
            }
            if (!part.split) {
                all_reduce(d_gadf_sums.ptr, kGadfSums);
            }
            t_linalg.stop(stream);
            GADFIT_CUDA_TRY(cudaMemcpyAsync(h_gadf,
                                            d_gadf_sums.ptr,
                                            (kGadfSums + 1) * sizeof(double),
                                            cudaMemcpyDeviceToHost,
                                            stream));
            GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
            std::copy(h_gadf, h_gadf + kGadfSums + 1, sums);
            if (!std::isfinite(sums[kGadfSums])) {
                check_device_flags(); // a poisoned integral, not an ordinary bad step?
            }
        } };

        const double umnigh_m { std::exp(-0.2) }; // :516
        t_main.start(stream);
        chi2_phase(d_gadf_sums.ptr + kGadfSums);
        GADFIT_CUDA_TRY(cudaMemcpyAsync(h_gadf,
                                        d_gadf_sums.ptr + kGadfSums,
                                        sizeof(double),
                                        cudaMemcpyDeviceToHost,
                                        stream));
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        double old_chi2 { h_gadf[0] };
        double old_old_chi2 { old_chi2 };
        int iterations { 0 };
        bool done { false };
        while (!done) {
            // STEP 1-3 (:680-750)
            jacobian_phase(accelerate, false);
            lambda_lhs = lambda;
            double lambda_solved { lambda };
            solve_phase(true);
            // STEP 4 (:751-818)
            double new_chi2 { 0.0 };
            for (int i = 1; i <= settings.lambda_incs + 1; ++i) {
                measure();
                new_chi2 = sums[kGadfSums];
                trace.push_back({ lambda, new_chi2, old_chi2, iterations + 1, 0 });
                const double beta { iterations == 0 ? 0.0
                                                    : sums[1] / std::sqrt(sums[0])
                                                        / std::sqrt(sums[2]) };
                double factor { 1.0 };
                for (int k = 0; k < gadf.uphill; ++k) {
                    factor *= 1.0 - beta;
                }
                if (factor * new_chi2 < old_chi2) {
                    if (gadf.nielsen != 0) {
                        // delta1'(JTJ + lambda D) delta1 = delta1' J^T r, :766-769
                        const double rho { (old_chi2 - new_chi2) / 2 / sums[3] };
                        const double t { 2 * rho - 1 };
                        set_lambda(lambda * std::max(1 / settings.lambda_down, 1 - t * t * t));
                    }
                    if (gadf.umnigh != 0) { // :771-783
                        if (new_chi2 < old_chi2 && beta >= 0.0) {
                            umnigh_a = umnigh_a * umnigh_m + 1.0 - umnigh_m;
                            const double t { 1.0 - std::abs(2.0 * umnigh_a - 1.0) };
                            set_lambda(lambda * std::min(1.0, std::max(1e-2, t * t)));
                        } else {
                            umnigh_a = umnigh_a * umnigh_m + (1.0 - umnigh_m) / 2.0;
                            if (new_chi2 >= old_chi2) {
                                set_lambda(lambda
                                           / std::min(10.0,
                                                      std::max(1.0,
                                                               1.0
                                                                 - std::abs(2.0 * umnigh_a
                                                                            - 1.0))));
                            }
                        }
                    }
                    if (gadf.nielsen == 0 && gadf.umnigh == 0) {
                        set_lambda(lambda / settings.lambda_down);
                    }
                    break;
                }
                if (i <= settings.lambda_incs) { // :790-808
                    trace.back().decision = 1;
                    if (gadf.umnigh != 0) {
                        umnigh_a *= umnigh_m;
                        const double t { 1.0 - std::abs(2.0 * umnigh_a - 1.0) };
                        set_lambda(lambda
                                   * (beta < 0.0 ? std::min(1.0, std::max(1e-2, t * t))
                                                 : std::min(1.0, std::max(0.1, t))));
                    } else {
                        set_lambda(settings.lambda_up * lambda);
                    }
                    revert();
                    lambda_lhs = lambda;
                    lambda_solved = lambda;
                    solve_phase(false, true);
                } else { // :809-816
                    trace.back().decision = 2;
                    revert();
                    stop_reason = GADFIT_B200_STOP_LAMBDA_INCS;
                    done = true;
                }
            }
            if (done) {
                break;
            }
            // :819-824
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_old_delta1_par.ptr,
                                            d_delta1_par.ptr,
                                            n_par * sizeof(double),
                                            cudaMemcpyDeviceToDevice,
                                            stream));
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_old_delta1_g.ptr,
                                            d_delta1_g.ptr,
                                            kMaxShared * sizeof(double),
                                            cudaMemcpyDeviceToDevice,
                                            stream));
            old_old_chi2 = old_chi2;
            old_chi2 = std::min(old_chi2, new_chi2);
            ++iterations;
            // STEP 5 (:832-909), in the Fortran version's order
            if (gadf.chi2_abs >= 0.0 && old_chi2 / dof < gadf.chi2_abs) {
                stop_reason = GADFIT_B200_STOP_CHI2_ABS;
            } else if (gadf.chi2_rel >= 0.0
                       && (old_old_chi2 - old_chi2) / old_chi2 < gadf.chi2_rel) {
                stop_reason = GADFIT_B200_STOP_CHI2_REL;
            } else if (gadf.grad_chi2 >= 0.0 && 2 * std::sqrt(sums[4]) < gadf.grad_chi2) {
                stop_reason = GADFIT_B200_STOP_GRAD_CHI2;
            } else if (gadf.cos_phi >= 0.0
                       // |r'J delta1| / |r| / |J delta1|,  |J delta1|^2 =
                       // delta1'J^T r - lambda delta1'D delta1
                       && std::abs(sums[3]) / std::sqrt(sums[7])
                              / std::sqrt(sums[3] - lambda_solved * sums[0])
                            < gadf.cos_phi) {
                stop_reason = GADFIT_B200_STOP_COS_PHI;
            } else if (gadf.rel_error >= 0.0 && sums[5] == 0.0) {
                stop_reason = GADFIT_B200_STOP_REL_ERROR;
            } else if (gadf.rel_error_global >= 0.0 && sums[6] == 0.0) {
                stop_reason = GADFIT_B200_STOP_REL_ERROR_GLOBAL;
            } else if (iterations >= settings.iteration_limit) {
                stop_reason = GADFIT_B200_STOP_ITERATION_LIMIT;
            } else {
                continue;
            }
            break;
        }
        t_main.stop(stream);
        n_iterations = iterations;
        download_parameters();
        check_device_flags();
        for (PhaseTimer* t :
             { &t_jacobian, &t_chi2, &t_linalg, &t_omega, &t_main, &t_jac_kernel }) {
            t->collect();
        }
        fitted = true;
    }

    auto public_chi2() -> double
    {
        if (sets.empty()) {
            return 0.0;
        }
        ensure_device();
        upload_parameters();
        chi2_phase(d_scratch.ptr);
        GADFIT_CUDA_TRY(cudaMemcpyAsync(
          h_scalar, d_scratch.ptr, sizeof(double), cudaMemcpyDeviceToHost, stream));
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        if (!std::isfinite(h_scalar[0])) {
            check_device_flags();
        }
        return h_scalar[0];
    }

    auto public_value(const double arg, const int i_dataset) -> double
    {
        if (i_dataset < 0 || i_dataset >= static_cast<int>(sets.size())) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "invalid data set index" };
        }
        ensure_device();
        d_stage.ensure(std::max<size_t>(d_stage.capacity, NP));
        GADFIT_CUDA_TRY(cudaMemcpyAsync(d_stage.ptr,
                                        par.data() + static_cast<size_t>(i_dataset) * NP,
                                        NP * sizeof(double),
                                        cudaMemcpyHostToDevice,
                                        stream));
        ops->value(stream, d_stage.ptr, arg, d_scratch.ptr);
        launched();
        GADFIT_CUDA_TRY(cudaMemcpyAsync(
          h_scalar, d_scratch.ptr, sizeof(double), cudaMemcpyDeviceToHost, stream));
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        return h_scalar[0];
    }

    // -------------------------------------------------------------- getters

    template <class T>
    auto download(const T* src, const size_t n) -> std::vector<T>
    {
        std::vector<T> out(n);
        if (n > 0) {
            GADFIT_CUDA_TRY(cudaMemcpyAsync(
              out.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
            GADFIT_CUDA_TRY(cudaStreamSynchronize(stream));
        }
        return out;
    }

    // Block-arrow inverse of J^T J of the last Jacobian evaluation (kernels at
    // the top of this file): shared block + per data set [ss | sg | W]
    struct InverseBlocks
    {
        std::vector<double> gg;     // [pg][pg]
        std::vector<double> blocks; // [S][3 NP NP]
    };
    DeviceBuffer<double> d_inv_gg, d_inv_blocks;

    auto inverse_blocks_device() -> void
    {
        const int S { static_cast<int>(sets.size()) };
        const int VG { shared_len() };
        const size_t NN { static_cast<size_t>(NP) * NP };
        // factors and Schur sums at lambda = 0
        DeviceBuffer<LmState> zero_state;
        DeviceBuffer<unsigned char> factors;
        DeviceBuffer<double> contrib, schur, A, L, work;
        zero_state.zero(1, stream); // lambda = 0
        factors.ensure(static_cast<size_t>(S) * ops->factor_bytes());
        GADFIT_CUDA_TRY(cudaMemsetAsync(factors.ptr, 0, static_cast<size_t>(S) * ops->factor_bytes(), stream));
        contrib.ensure(static_cast<size_t>(S) * std::max(VG, 2));
        schur.zero(std::max(VG, 1), stream);
        if (n_active > pg) {
            ops->local_schur(stream, d_M.ptr, d_active.ptr, global_mask, S, d_dtd_local.ptr,
                             zero_state.ptr, factors.ptr, contrib.ptr, VG);
            launched();
            column_sums(contrib.ptr, VG, schur.ptr);
        }
        d_inv_gg.zero(std::max<size_t>(static_cast<size_t>(pg) * pg, 1), stream);
        if (pg > 0) {
            A.ensure(static_cast<size_t>(pg) * pg);
            L.zero(static_cast<size_t>(pg) * pg, stream);
            k_schur_matrix<<<blocks_for(pg * pg, 128), 128, 0, stream>>>(
              pg, d_shared_sums.ptr, schur.ptr, A.ptr);
            k_dense_inverse<<<1, 1, 0, stream>>>(pg, A.ptr, L.ptr, d_inv_gg.ptr);
            launched(2);
        }
        d_inv_blocks.zero(static_cast<size_t>(S) * 3 * NN, stream);
        work.zero(static_cast<size_t>(S) * NN, stream);
        k_inverse_local<<<blocks_for(S, 128), 128, 0, stream>>>(
          NP, S, d_active.ptr, global_mask, reinterpret_cast<const double*>(factors.ptr),
          d_inv_gg.ptr, pg, d_inv_blocks.ptr, work.ptr);
        launched();
        GADFIT_CUDA_TRY(cudaStreamSynchronize(stream)); // the temporaries die here
    }

    auto inverse_blocks() -> InverseBlocks
    {
        inverse_blocks_device();
        const size_t NN { static_cast<size_t>(NP) * NP };
        return { download(d_inv_gg.ptr, static_cast<size_t>(pg) * pg),
                 download(d_inv_blocks.ptr, sets.size() * 3 * NN) };
    }

    auto inverse_dense() -> std::vector<double>
    {
        const int S { static_cast<int>(sets.size()) };
        const size_t n { static_cast<size_t>(n_active) };
        inverse_blocks_device();
        std::vector<int> col_set(n, -1);
        std::vector<int> col_idx(n, 0);
        for (int a = 0; a < pg; ++a) {
            col_idx[a] = a;
        }
        for (int s = 0; s < S; ++s) {
            const unsigned local { active[s] & ~global_mask };
            for (int k = 0; k < NP; ++k) {
                const long long c { column(s, k) };
                if (c >= pg) {
                    col_set[c] = s;
                    col_idx[c] = slot_of(local, k);
                }
            }
        }
        DeviceBuffer<int> d_set, d_idx;
        DeviceBuffer<double> d_out;
        d_set.upload(col_set, stream);
        d_idx.upload(col_idx, stream);
        d_out.ensure(n * n);
        k_inverse_assemble<<<blocks_for(static_cast<long long>(n * n), 256), 256, 0, stream>>>(
          static_cast<long long>(n), NP, pg, d_set.ptr, d_idx.ptr, d_inv_gg.ptr, d_inv_blocks.ptr,
          d_out.ptr);
        launched();
        return download(d_out.ptr, n * n);
    }

    // Dense views of the block-arrow storage (single rank only)
    auto dense_array(const int which) -> std::vector<double>
    {
        if (!fitted) {
            throw StatusError { GADFIT_B200_NOT_AVAILABLE, "fit() has not been called" };
        }
        if (world > 1) {
            throw StatusError { GADFIT_B200_NOT_AVAILABLE,
                                "dense arrays are only assembled on a single rank" };
        }
        const int S { static_cast<int>(sets.size()) };
        const size_t n { static_cast<size_t>(n_active) };
        const int NA { NP + 1 };
        const int NT { NA * (NA + 1) / 2 };
        const auto need_dense { [&](const size_t rows) {
            if (rows * n > kDenseLimit) {
                throw StatusError { GADFIT_B200_NOT_AVAILABLE,
                                    "dense array exceeds 2^27 entries" };
            }
        } };
        GADFIT_CUDA_TRY(cudaSetDevice(device));
        switch (which) {
        case GADFIT_B200_JTJ:
        case GADFIT_B200_LEFT_SIDE:
        case GADFIT_B200_DTD:
        case GADFIT_B200_INV_JTJ: {
            need_dense(n);
            std::vector<double> out(n * n, 0.0);
            if (settings.iteration_limit == 0 && which != GADFIT_B200_DTD) {
                return out; // fit() never computed a Jacobian
            }
            if (which == GADFIT_B200_INV_JTJ) {
                return inverse_dense();
            }
            const bool want_jtj { which != GADFIT_B200_DTD };
            std::vector<double> M;
            std::vector<double> sums;
            if (want_jtj) {
                M = download(d_M.ptr, static_cast<size_t>(S) * NT);
                sums = download(d_shared_sums.ptr,
                                static_cast<size_t>(std::max(shared_len(), 1)));
            }
            const auto dl { download(d_dtd_local.ptr, static_cast<size_t>(S) * NP) };
            const auto dg { download(d_dtd_global.ptr, static_cast<size_t>(std::max(pg, 1))) };
            const double lam { which == GADFIT_B200_LEFT_SIDE ? lambda_lhs : 0.0 };
            if (want_jtj) {
                for (int a = 0; a < pg; ++a) {
                    for (int b = a; b < pg; ++b) {
                        const double v { sums[tri_index(pg, a, b)] };
                        out[a * n + b] = v;
                        out[b * n + a] = v;
                    }
                }
                for (int s = 0; s < S; ++s) {
                    const double* Ms { M.data() + static_cast<size_t>(s) * NT };
                    for (int k = 0; k < NP; ++k) {
                        const long long ck { column(s, k) };
                        if (ck < pg) {
                            continue; // passive or shared-shared (already done)
                        }
                        for (int m = 0; m < NP; ++m) {
                            const long long cm { column(s, m) };
                            if (cm < 0) {
                                continue;
                            }
                            const double v { k <= m ? Ms[tri_index(NA, k, m)]
                                                    : Ms[tri_index(NA, m, k)] };
                            out[ck * n + cm] = v;
                            out[cm * n + ck] = v;
                        }
                    }
                }
            }
            if (which == GADFIT_B200_DTD || which == GADFIT_B200_LEFT_SIDE) {
                const double f { which == GADFIT_B200_DTD ? 1.0 : lam };
                for (int a = 0; a < pg; ++a) {
                    out[a * n + a] += f * dg[a];
                }
                for (int s = 0; s < S; ++s) {
                    for (int k = 0; k < NP; ++k) {
                        const long long c { column(s, k) };
                        if (c >= pg) {
                            out[c * n + c] += f * dl[static_cast<size_t>(s) * NP + k];
                        }
                    }
                }
            }
            return out;
        }
        case GADFIT_B200_INV_JTJ_DIAGONAL: {
            std::vector<double> out(n, 0.0);
            if (settings.iteration_limit == 0) {
                return out;
            }
            const InverseBlocks inv { inverse_blocks() };
            const size_t NN { static_cast<size_t>(NP) * NP };
            for (int a = 0; a < pg; ++a) {
                out[a] = inv.gg[static_cast<size_t>(a) * pg + a];
            }
            for (int s = 0; s < S; ++s) {
                const unsigned local { active[s] & ~global_mask };
                for (int k = 0; k < NP; ++k) {
                    const long long c { column(s, k) };
                    if (c >= pg) {
                        const int i { slot_of(local, k) };
                        out[c] = inv.blocks[static_cast<size_t>(s) * 3 * NN + static_cast<size_t>(i) * NP + i];
                    }
                }
            }
            return out;
        }
        case GADFIT_B200_RIGHT_SIDE: {
            std::vector<double> out(n, 0.0);
            if (settings.iteration_limit == 0) {
                return out;
            }
            const auto M { download(d_M.ptr, static_cast<size_t>(S) * NT) };
            const auto sums { download(d_shared_sums.ptr, static_cast<size_t>(std::max(shared_len(), 1))) };
            for (int a = 0; a < pg; ++a) {
                out[a] = sums[pg * (pg + 1) / 2 + a];
            }
            for (int s = 0; s < S; ++s) {
                for (int k = 0; k < NP; ++k) {
                    const long long c { column(s, k) };
                    if (c >= pg) {
                        out[c] = M[static_cast<size_t>(s) * NT + tri_index(NA, k, NP)];
                    }
                }
            }
            return out;
        }
        case GADFIT_B200_DELTA1:
        case GADFIT_B200_DELTA2: {
            std::vector<double> out(n, 0.0);
            if (settings.iteration_limit == 0
                || (which == GADFIT_B200_DELTA2 && !have_delta2)) {
                return out;
            }
            if (which == GADFIT_B200_DELTA2
                && download(d_state.ptr, 1).front().use_delta2 == 0) {
                return out; // ratio test failed: zero-filled, lm_solver.cpp:394-396
            }
            const auto d { download(which == GADFIT_B200_DELTA1 ? d_delta1_par.ptr
                                                                 : d_delta2_par.ptr,
                                    static_cast<size_t>(S) * NP) };
            for (int s = 0; s < S; ++s) {
                for (int k = 0; k < NP; ++k) {
                    const long long c { column(s, k) };
                    if (c >= 0) {
                        out[c] = d[static_cast<size_t>(s) * NP + k];
                    }
                }
            }
            return out;
        }
        case GADFIT_B200_OMEGA: {
            // The second directional derivative along delta1 at the parameters
            // of the last Jacobian evaluation (lm_solver.cpp:362-380)
            const size_t N { static_cast<size_t>(n_datapoints) };
            if (!have_delta2 || settings.iteration_limit == 0) {
                throw StatusError { GADFIT_B200_NOT_AVAILABLE,
                                    "omega exists only with geodesic acceleration" };
            }
            DeviceBuffer<double> keep;
            DeviceBuffer<double> scratch;
            const size_t n_par { static_cast<size_t>(S) * NP };
            keep.ensure(n_par);
            scratch.ensure(static_cast<size_t>(std::max(slots_omega.n_slots, 1)) * NP);
            d_omega.ensure(N);
            GADFIT_CUDA_TRY(cudaMemcpyAsync(keep.ptr, d_par.ptr, n_par * sizeof(double),
                                            cudaMemcpyDeviceToDevice, stream));
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_par.ptr, d_par_jac.ptr, n_par * sizeof(double),
                                            cudaMemcpyDeviceToDevice, stream));
            ops->omega(slots_omega.grid, stream, sweep_args(&slots_omega.cta), d_delta1_par.ptr, d_J.ptr,
                       n_local_points, slots_omega.base.ptr, scratch.ptr, d_omega.ptr);
            launched();
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_par.ptr, keep.ptr, n_par * sizeof(double),
                                            cudaMemcpyDeviceToDevice, stream));
            return download(d_omega.ptr, N);
        }
        case GADFIT_B200_JACOBIAN:
        case GADFIT_B200_RESIDUALS: {
            const size_t N { static_cast<size_t>(n_datapoints) };
            if (which == GADFIT_B200_JACOBIAN) {
                need_dense(N);
            }
            std::vector<double> out(which == GADFIT_B200_JACOBIAN ? N * n : N, 0.0);
            if (settings.iteration_limit == 0) {
                return out;
            }
            // Re-run the Jacobian sweep at the parameters of the last
            // Jacobian evaluation with row / residual output switched on
            DeviceBuffer<double> keep;
            const size_t n_par { static_cast<size_t>(S) * NP };
            keep.ensure(n_par);
            GADFIT_CUDA_TRY(cudaMemcpyAsync(keep.ptr, d_par.ptr, n_par * sizeof(double),
                                            cudaMemcpyDeviceToDevice, stream));
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_par.ptr, d_par_jac.ptr, n_par * sizeof(double),
                                            cudaMemcpyDeviceToDevice, stream));
            d_J.ensure(static_cast<size_t>(NP) * N);
            d_resid.ensure(N);
            DeviceBuffer<double> scratch;
            scratch.ensure(static_cast<size_t>(std::max(slots_jac.n_slots, 1)) * NT);
            ops->jacobian(slots_jac.grid, stream, sweep_args(&slots_jac.cta), settings.loss,
                          slots_jac.base.ptr, scratch.ptr, d_J.ptr, d_resid.ptr,
                          n_local_points);
            launched();
            GADFIT_CUDA_TRY(cudaMemcpyAsync(d_par.ptr, keep.ptr, n_par * sizeof(double),
                                            cudaMemcpyDeviceToDevice, stream));
            if (which == GADFIT_B200_RESIDUALS) {
                return download(d_resid.ptr, N);
            }
            const auto J { download(d_J.ptr, static_cast<size_t>(NP) * N) };
            size_t row { 0 };
            for (int s = 0; s < S; ++s) {
                for (int p = 0; p < sets[s].n; ++p, ++row) {
                    for (int k = 0; k < NP; ++k) {
                        const long long c { column(s, k) };
                        if (c >= 0) {
                            out[row * n + c] = J[static_cast<size_t>(k) * N + row];
                        }
                    }
                }
            }
            return out;
        }
        default:
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "unknown array" };
        }
    }
};

} // namespace gadfit_b200

struct gadfit_b200_solver : gadfit_b200::SolverImpl
{};

// ------------------------------------------------------------------- C-ABI
//
// Defined inside the namespace for name lookup only: functions with C
// language linkage are the same functions as the ones the header declares.

namespace gadfit_b200 {

namespace {

inline auto require_output(const void* p) -> void
{
    if (p == nullptr) {
        throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "null output pointer" };
    }
}

template <class F>
auto guarded(gadfit_b200_solver* h, F&& f) -> int
{
    if (h == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
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

struct ModelInfo
{
    const char* name; // static storage of the library / the plug-in
    int n_pars;
    void* (*make_indexed)(int); // plug-in factory; nullptr for built-in functions
    int plugin_index;
};

// Built-in functions first, then those of the loaded plug-ins
// (gadfit_b200_load_models) in loading order
struct ModelRegistry
{
    std::mutex mutex;
    std::vector<ModelInfo> table;
    std::vector<void*> libraries;
    ModelRegistry()
    {
        models::for_each_model([&]<class Model>() {
            table.push_back({ Model::name, Model::n_pars, nullptr, -1 });
        });
    }
};

auto model_registry() -> ModelRegistry&
{
    static ModelRegistry registry;
    return registry;
}

// include/gadfit_b200/model_plugin.h
auto load_models(const char* path) -> int
{
    auto& reg { model_registry() };
    const std::lock_guard lock { reg.mutex };
    void* lib { dlopen(path, RTLD_NOW | RTLD_LOCAL) };
    if (lib == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
    for (void* known : reg.libraries) {
        if (known == lib) {
            dlclose(lib); // already loaded: drop the extra reference
            return GADFIT_B200_OK;
        }
    }
    const auto abi { reinterpret_cast<long long (*)()>(dlsym(lib, "gadfit_b200_plugin_abi")) };
    const auto count { reinterpret_cast<int (*)()>(dlsym(lib, "gadfit_b200_plugin_n_models")) };
    const auto info { reinterpret_cast<int (*)(int, const char**, int*)>(
      dlsym(lib, "gadfit_b200_plugin_model_info")) };
    const auto make { reinterpret_cast<void* (*)(int)>(dlsym(lib, "gadfit_b200_plugin_make")) };
    if (abi == nullptr || count == nullptr || info == nullptr || make == nullptr
        || abi() != plugin::abi_fingerprint()) {
        dlclose(lib);
        return GADFIT_B200_INVALID_ARGUMENT; // not a plug-in, or built against other headers
    }
    std::vector<ModelInfo> added;
    for (int i = 0; i < count(); ++i) {
        const char* name { nullptr };
        int n_pars { 0 };
        // (functions with more than kMaxShared parameters are listed, so that
        // create() can say why it refuses them)
        if (info(i, &name, &n_pars) != 0 || name == nullptr || n_pars < 1) {
            dlclose(lib);
            return GADFIT_B200_INVALID_ARGUMENT;
        }
        for (const auto& m : reg.table) {
            if (std::string_view { m.name } == name) {
                dlclose(lib);
                return GADFIT_B200_INVALID_ARGUMENT; // name already taken
            }
        }
        added.push_back({ name, n_pars, make, i });
    }
    reg.table.insert(reg.table.end(), added.begin(), added.end());
    reg.libraries.push_back(lib);
    return GADFIT_B200_OK;
}

} // namespace

#pragma GCC visibility push(default)
extern "C" {

int gadfit_b200_n_models(void)
{
    auto& reg { model_registry() };
    const std::lock_guard lock { reg.mutex };
    return static_cast<int>(reg.table.size());
}

int gadfit_b200_load_models(const char* path)
{
    if (path == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
    try {
        return load_models(path);
    } catch (const std::exception&) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
}

int gadfit_b200_model_info(int i_model, const char** name, int* n_pars)
{
    auto& reg { model_registry() };
    const std::lock_guard lock { reg.mutex };
    if (i_model < 0 || i_model >= static_cast<int>(reg.table.size()) || name == nullptr
        || n_pars == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
    *name = reg.table[i_model].name;
    *n_pars = reg.table[i_model].n_pars;
    return GADFIT_B200_OK;
}

int gadfit_b200_create(const char* model_name, gadfit_b200_solver** out)
{
    if (model_name == nullptr || out == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
    try {
        // kernels of each model group live in their own translation unit
        std::unique_ptr<ModelOps> ops;
        static_for<0, models::n_model_groups>([&](auto G) {
            if (!ops) {
                ops = make_model_ops_group<G>(model_name);
            }
        });
        if (!ops) { // a plug-in function?
            auto& reg { model_registry() };
            const std::lock_guard lock { reg.mutex };
            for (const auto& m : reg.table) {
                if (m.make_indexed != nullptr && std::string_view { m.name } == model_name) {
                    if (m.n_pars > kMaxShared) {
                        return GADFIT_B200_TOO_MANY_PARAMETERS;
                    }
                    ops.reset(static_cast<ModelOps*>(m.make_indexed(m.plugin_index)));
                    break;
                }
            }
        }
        if (!ops) {
            return GADFIT_B200_UNKNOWN_MODEL;
        }
        auto* h { new gadfit_b200_solver {} };
        h->NP = ops->n_pars();
        h->ops = std::move(ops);
        gadfit_b200_default_settings(&h->settings);
        int dev { 0 };
        if (cudaGetDevice(&dev) == cudaSuccess) {
            h->device = dev;
        } else {
            (void)cudaGetLastError();
        }
        *out = h;
        return GADFIT_B200_OK;
    } catch (const std::exception&) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
}

int gadfit_b200_destroy(gadfit_b200_solver* solver)
{
    delete solver;
    return GADFIT_B200_OK;
}

int gadfit_b200_add_dataset(gadfit_b200_solver* solver,
                            int n_datapoints,
                            const double* x_data,
                            const double* y_data,
                            const double* errors_or_null)
{
    return guarded(solver, [&] {
        solver->add_dataset(n_datapoints, x_data, y_data, errors_or_null);
    });
}

int gadfit_b200_set_par(gadfit_b200_solver* solver,
                        int i_par,
                        double val,
                        int active,
                        int i_dataset,
                        const char* name_or_null)
{
    return guarded(solver, [&] {
        solver->set_par(i_par, val, active != 0, i_dataset);
        // lm_solver.cpp:99-104
        if (i_par >= static_cast<int>(solver->names.size())) {
            solver->names.resize(i_par + 1);
        }
        if (name_or_null != nullptr && name_or_null[0] != '\0') {
            solver->names[i_par] = name_or_null;
        }
    });
}

int gadfit_b200_default_settings(gadfit_b200_settings* out)
{
    if (out == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
    // lm_solver.h:88-96
    *out = { 1000, 3, 10.0, 10.0, 1, nullptr, 0, -1.0, 0U, GADFIT_B200_LOSS_LINEAR, 1 };
    return GADFIT_B200_OK;
}

int gadfit_b200_set_settings(gadfit_b200_solver* solver, const gadfit_b200_settings* s)
{
    return guarded(solver, [&] {
        if (s == nullptr) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "null settings" };
        }
        solver->settings = *s;
        if (s->DTD_min != nullptr && s->n_DTD_min > 0) {
            solver->DTD_min.assign(s->DTD_min, s->DTD_min + s->n_DTD_min);
        } else {
            solver->DTD_min = { 0.0 };
        }
        solver->settings.DTD_min = nullptr;
    });
}

int gadfit_b200_set_log(gadfit_b200_solver* solver, gadfit_b200_log_fn fn, void* user)
{
    return guarded(solver, [&] {
        solver->log_fn = fn;
        solver->log_user = user;
    });
}

int gadfit_b200_default_gadf_options(gadfit_b200_gadf_options* out)
{
    if (out == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
    *out = { 0, 0, 0, 0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0 };
    return GADFIT_B200_OK;
}

int gadfit_b200_set_gadf_options(gadfit_b200_solver* solver, const gadfit_b200_gadf_options* o)
{
    return guarded(solver, [&] {
        if (o == nullptr || o->uphill < 0) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "invalid gadf options" };
        }
        solver->gadf = *o;
    });
}

int gadfit_b200_fit(gadfit_b200_solver* solver, double lambda)
{
    return guarded(solver, [&] { solver->fit(lambda); });
}

int gadfit_b200_get_stop_reason(gadfit_b200_solver* solver, int* out)
{
    return guarded(solver, [&] {
        if (out == nullptr) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "null output" };
        }
        *out = solver->stop_reason;
    });
}

int gadfit_b200_get_par(gadfit_b200_solver* solver, int i_par, int i_dataset, double* out)
{
    return guarded(solver, [&] {
        if (i_par < 0 || i_par >= solver->NP || i_dataset < 0
            || i_dataset >= static_cast<int>(solver->sets.size())) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "index out of range" };
        }
        require_output(out);
        *out = solver->par[static_cast<size_t>(i_dataset) * solver->NP + i_par];
    });
}

int gadfit_b200_get_value(gadfit_b200_solver* solver, double arg, int i_dataset, double* out)
{
    return guarded(solver, [&] {
        require_output(out);
        *out = solver->public_value(arg, i_dataset);
    });
}

int gadfit_b200_chi2(gadfit_b200_solver* solver, double* out)
{
    return guarded(solver, [&] {
        require_output(out);
        *out = solver->public_chi2();
    });
}

int gadfit_b200_dof(gadfit_b200_solver* solver, int* out)
{
    return guarded(solver, [&] {
        require_output(out);
        *out = solver->dof;
    });
}

int gadfit_b200_get_array(gadfit_b200_solver* solver,
                          int which,
                          double* buf,
                          size_t cap,
                          size_t* need)
{
    return guarded(solver, [&] {
        const auto data { solver->dense_array(which) };
        if (need != nullptr) {
            *need = data.size();
        }
        const size_t n { std::min(cap, data.size()) };
        if (buf != nullptr && n > 0) {
            std::memcpy(buf, data.data(), n * sizeof(double));
        }
    });
}

int gadfit_b200_get_trace(gadfit_b200_solver* solver,
                          gadfit_b200_trial* buf,
                          size_t cap,
                          size_t* n)
{
    return guarded(solver, [&] {
        if (n != nullptr) {
            *n = solver->trace.size();
        }
        const size_t m { std::min(cap, solver->trace.size()) };
        if (buf != nullptr && m > 0) {
            std::memcpy(buf, solver->trace.data(), m * sizeof(gadfit_b200_trial));
        }
    });
}

int gadfit_b200_n_iterations(gadfit_b200_solver* solver, int* out)
{
    return guarded(solver, [&] {
        require_output(out);
        *out = solver->n_iterations;
    });
}

int gadfit_b200_get_timings(gadfit_b200_solver* solver, gadfit_b200_timings* out)
{
    return guarded(solver, [&] {
        require_output(out);
        const PhaseTimer* timers[5] { &solver->t_jacobian,
                                      &solver->t_chi2,
                                      &solver->t_linalg,
                                      &solver->t_omega,
                                      &solver->t_main };
        for (int i = 0; i < 5; ++i) {
            out->seconds[i] = timers[i]->seconds;
            out->calls[i] = timers[i]->calls;
        }
    });
}

int gadfit_b200_n_launches(gadfit_b200_solver* solver, long long* out)
{
    return guarded(solver, [&] {
        require_output(out);
        *out = solver->n_launches;
    });
}

int gadfit_b200_comm_stats(gadfit_b200_solver* solver,
                           int* peer_reduce_active,
                           long long* n_peer_reductions,
                           long long* n_nccl_reductions)
{
    return guarded(solver, [&] {
        require_output(peer_reduce_active);
        require_output(n_peer_reductions);
        require_output(n_nccl_reductions);
        *peer_reduce_active = solver->peer.active ? 1 : 0;
        *n_peer_reductions = solver->n_peer_reductions;
        *n_nccl_reductions = solver->n_nccl_reductions;
    });
}

int gadfit_b200_jacobian_kernel_time(gadfit_b200_solver* solver,
                                     double* seconds_per_launch,
                                     int* launches)
{
    return guarded(solver, [&] {
        require_output(seconds_per_launch);
        require_output(launches);
        // (fused small-problem path: the Jacobian phase is exactly the kernel)
        const auto& t { solver->t_jac_kernel.calls > 0 ? solver->t_jac_kernel
                                                        : solver->t_jacobian };
        *launches = t.calls;
        *seconds_per_launch = t.calls > 0 ? t.seconds / t.calls : 0.0;
    });
}

const char* gadfit_b200_last_error(gadfit_b200_solver* solver)
{
    return solver != nullptr ? solver->last_error.c_str() : "null solver";
}

const char* gadfit_b200_status_name(int status)
{
    static const char* names[] { "OK",
                                 "UnknownOperation",
                                 "LateAddDatasetCall",
                                 "SetParInvalidIndex",
                                 "UninitializedParameter",
                                 "NegativeDegreesOfFreedom",
                                 "NoFittingParameters",
                                 "NoGlobalParameters",
                                 "InsufficientIntegrationWorkspace",
                                 "UnknownModel",
                                 "CudaError",
                                 "NcclError",
                                 "InvalidArgument",
                                 "NotAvailable",
                                 "TooManyParameters" };
    if (status < 0 || status > GADFIT_B200_TOO_MANY_PARAMETERS) {
        return "?";
    }
    return names[status];
}

int gadfit_b200_set_device(gadfit_b200_solver* solver, int device)
{
    return guarded(solver, [&] {
        if (solver->device_ready && device != solver->device) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT,
                                "the device cannot change after first use" };
        }
        solver->device = device;
    });
}

int gadfit_b200_upload(gadfit_b200_solver* solver)
{
    return guarded(solver, [&] {
        if (!solver->data_uploaded) {
            solver->upload_data();
        }
    });
}

int gadfit_b200_invalidate_data(gadfit_b200_solver* solver)
{
    return guarded(solver, [&] { solver->data_uploaded = false; });
}

int gadfit_b200_set_phase_timers(gadfit_b200_solver* solver, const int on)
{
    return guarded(solver, [&] {
        for (PhaseTimer* t : { &solver->t_jacobian, &solver->t_chi2, &solver->t_linalg,
                               &solver->t_omega, &solver->t_jac_kernel }) {
            t->enabled = on != 0;
        }
    });
}

int gadfit_b200_comm_unique_id(void* id_128_bytes)
{
    if (id_128_bytes == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
    try {
        auto& nccl { Nccl::instance() };
        ncclUniqueId id {};
        static_assert(sizeof(ncclUniqueId) == 128);
        nccl.check(nccl.getUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(id_128_bytes, &id, sizeof(id));
        return GADFIT_B200_OK;
    } catch (const std::exception&) {
        return GADFIT_B200_NCCL_ERROR;
    }
}

int gadfit_b200_set_comm(gadfit_b200_solver* solver,
                         int rank,
                         int world_size,
                         const void* id_128_bytes)
{
    return guarded(solver, [&] {
        if (world_size < 1 || rank < 0 || rank >= world_size) {
            throw StatusError { GADFIT_B200_INVALID_ARGUMENT, "bad rank / world size" };
        }
        solver->ensure_device();
        if (solver->comm != nullptr) {
            // re-attaching replaces the previous communicator
            Nccl::instance().commDestroy(solver->comm);
            solver->comm = nullptr;
        }
        if (world_size > 1) {
            auto& nccl { Nccl::instance() };
            ncclUniqueId id {};
            std::memcpy(&id, id_128_bytes, sizeof(id));
            nccl.check(nccl.commInitRank(&solver->comm, world_size, id, rank),
                       "ncclCommInitRank");
            solver->peer.setup(rank, world_size, solver->comm, solver->stream);
            solver->peer.table.status = &solver->d_state.ptr->flags;
        }
        solver->rank = rank;
        solver->world = world_size;
        solver->data_uploaded = false;
    });
}

int gadfit_b200_partition(int n_datasets,
                          const int* n_points,
                          int rank,
                          int world_size,
                          int* first_dataset,
                          int* end_dataset,
                          int* split)
{
    if (n_datasets < 0 || n_points == nullptr || world_size < 1 || rank < 0
        || rank >= world_size || first_dataset == nullptr || end_dataset == nullptr
        || split == nullptr) {
        return GADFIT_B200_INVALID_ARGUMENT;
    }
    const Partition part { compute_partition(n_datasets, n_points, rank, world_size) };
    *first_dataset = part.first;
    *end_dataset = part.end;
    *split = part.split ? 1 : 0;
    return GADFIT_B200_OK;
}

} // extern "C"
#pragma GCC visibility pop

} // namespace gadfit_b200
