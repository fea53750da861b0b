// Type-erased launchers: one instantiation of every kernel per registered
// fitting function (the "model registry" of include/gadfit_b200/fit_models.h).
#pragma once

#include "block_arrow_kernels.cuh"
#include "device_util.cuh"

#include <cstdlib>
#include <cstring>
#include <memory>

namespace gadfit_b200 {

struct ModelOps
{
    virtual ~ModelOps() = default;
    [[nodiscard]] virtual auto name() const -> const char* = 0;
    [[nodiscard]] virtual auto n_pars() const -> int = 0;
    // Resident CTAs per SM of each sweep kernel (0: chi2, 1: Jacobian, 2: omega)
    [[nodiscard]] virtual auto ctas_per_sm(int which) const -> int = 0;
    [[nodiscard]] virtual auto factor_bytes() const -> size_t = 0;
    // Tiles per unrolled group of the chi2 sweep (its CTA ranges are aligned to it)
    [[nodiscard]] virtual auto chi2_unroll() const -> int = 0;
    // First device use: opt in to the shared memory the Jacobian sweep needs
    // and query occupancies (kept out of the constructor so that the host-side
    // bookkeeping of the solver works without a GPU)
    virtual auto init_device() -> void = 0;

    // fin.state != nullptr: the last CTA also finishes the trial step
    // (device-driven loop, sweep_kernels.cuh: TrialFinish)
    virtual auto chi2(int grid,
                      cudaStream_t st,
                      const SweepArgs& a,
                      double* partials,
                      const TrialFinish& fin) const -> void = 0;
    virtual auto jacobian(int grid,
                          cudaStream_t st,
                          const SweepArgs& a,
                          int loss,
                          const int* slot_base,
                          double* partials,
                          double* J_out,
                          double* r_out,
                          long long n_points) const -> void = 0;
    virtual auto omega(int grid,
                       cudaStream_t st,
                       const SweepArgs& a,
                       const double* seed,
                       const double* J,
                       long long n_points,
                       const int* slot_base,
                       double* partials,
                       double* omega_out) const -> void = 0;
    virtual auto value(cudaStream_t st, const double* par, double arg, double* out) const
      -> void = 0;

    virtual auto local_prepare(cudaStream_t st,
                               const double* M,
                               const unsigned* active,
                               unsigned global_mask,
                               int n_sets,
                               int damp_max,
                               double* dtd_local,
                               double* contrib,
                               int VG) const -> void = 0;
    virtual auto local_schur(cudaStream_t st,
                             const double* M,
                             const unsigned* active,
                             unsigned global_mask,
                             int n_sets,
                             const double* dtd_local,
                             const LmState* state,
                             void* factors,
                             double* contrib,
                             int VG) const -> void = 0;
    virtual auto local_schur_rhs(cudaStream_t st,
                                 const double* v_sets,
                                 const unsigned* active,
                                 unsigned global_mask,
                                 int n_sets,
                                 void* factors,
                                 double* contrib,
                                 int VG2) const -> void = 0;
    virtual auto local_backsolve(cudaStream_t st,
                                 const unsigned* active,
                                 unsigned global_mask,
                                 int n_sets,
                                 const void* factors,
                                 const double* delta_g,
                                 double* delta_par) const -> void = 0;
    virtual auto acc_ratio_local(cudaStream_t st,
                                 const unsigned* active,
                                 unsigned global_mask,
                                 int n_sets,
                                 const double* dtd_local,
                                 const double* delta1_par,
                                 const double* delta2_par,
                                 double* contrib) const -> void = 0;
    // Latency-bound regime: everything between the Jacobian sweep and the chi2
    // sweep of a trial step in one launch of one cluster (block_arrow_kernels.cuh)
    virtual auto solve_small(cudaStream_t st, const SmallSolveArgs& a) const -> void = 0;
};

inline auto blocks_for(const long long n, const int threads) -> int
{
    return static_cast<int>((n + threads - 1) / threads);
}


template <class Model>
struct ModelOpsImpl final : ModelOps
{
    static constexpr int NP { Model::n_pars };
    using Shape = JacobianShape<Model>;
    using WsShape = JacobianWsShape<Model>;
    int occupancy[3] { 1, 1, 1 };
    bool device_ready { false };
    // true: the software-pipelined Jacobian sweep (every warp evaluates tile
    // t + 1 while it contracts tile t); false: the warp-specialised one
    bool pipelined_jacobian { false };
    // the register-accumulating sweep (cheap functions with few parameters)
    bool register_jacobian { false };
    // instantiated only for the functions that ask for it (hint
    // `jacobian_in_registers`): four more kernels per function otherwise
    static constexpr bool kHasRegisterJacobian { JacobianInRegisters<Model>::value };
    bool chi2_staged_fits { false };
    // GADFIT_B200_CHI2=staged|direct (read at the first device use of a solver)
    // overrides the choice by size: 1 staged, -1 direct, 0 by size (A/B runs, tests)
    int chi2_force { 0 };
    bool small_smem_ready { false };
    static constexpr size_t kSmallSmemLimit { 200 * 1024 };

    auto init_device() -> void override
    {
        static_assert(NP >= 1 && NP <= kMaxShared, "1..32 parameters supported");
        if (device_ready) {
            return;
        }
        // Shared-memory carve-out: the maximum, so that kJacobianMinBlocks CTAs
        // (staging + prefetch buffers each) are resident.  Measured: a smaller
        // carve-out that still fits them (86-90 %, more L1 for the spilled
        // accumulators) changes the sweep time by < 0.1 %.
        const auto configure { [](auto* kernel) {
            GADFIT_CUDA_TRY(cudaFuncSetAttribute(kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(Shape::smem_bytes)));
            GADFIT_CUDA_TRY(cudaFuncSetAttribute(kernel,
                                                 cudaFuncAttributePreferredSharedMemoryCarveout,
                                                 cudaSharedmemCarveoutMaxShared));
        } };
        configure(k_jacobian<Model, false, false>);
        configure(k_jacobian<Model, false, true>);
        configure(k_jacobian<Model, true, false>);
        configure(k_jacobian<Model, true, true>);
        const auto configure_ws { [](auto* kernel) {
            GADFIT_CUDA_TRY(cudaFuncSetAttribute(kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(WsShape::smem_bytes)));
            GADFIT_CUDA_TRY(cudaFuncSetAttribute(kernel,
                                                 cudaFuncAttributePreferredSharedMemoryCarveout,
                                                 cudaSharedmemCarveoutMaxShared));
        } };
        configure_ws(k_jacobian_ws<Model, false, false>);
        configure_ws(k_jacobian_ws<Model, false, true>);
        configure_ws(k_jacobian_ws<Model, true, false>);
        configure_ws(k_jacobian_ws<Model, true, true>);
        if constexpr (kHasRegisterJacobian) {
            const auto configure_reg { [](auto* kernel) {
                GADFIT_CUDA_TRY(cudaFuncSetAttribute(
                  kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                  static_cast<int>(JacobianRegShape<Model>::smem_bytes)));
            } };
            configure_reg(k_jacobian_reg<Model, false, false>);
            configure_reg(k_jacobian_reg<Model, false, true>);
            configure_reg(k_jacobian_reg<Model, true, false>);
            configure_reg(k_jacobian_reg<Model, true, true>);
        }
        // Which Jacobian sweep (sweep_kernels.cuh).  Measured on B200 (r02c):
        // with several 8-column blocks the contraction dominates and dedicated
        // contraction warps win (config 2, 21 columns: 2.52 vs 2.62 ms); with a
        // single block the evaluation dominates and the software-pipelined
        // kernel, whose warps all evaluate with 128 registers, wins (config 4:
        // 0.90 vs 1.25 ms; config 5: 8.7 vs 8.9 ms).  GADFIT_B200_JAC=ws|sp
        // overrides the choice (A/B runs).
        const char* jac { std::getenv("GADFIT_B200_JAC") };
        pipelined_jacobian = Shape::NB < 2;
        const char* chi2_variant { std::getenv("GADFIT_B200_CHI2") };
        chi2_force = chi2_variant == nullptr
                       ? 0
                       : (std::strcmp(chi2_variant, "staged") == 0
                            ? 1
                            : (std::strcmp(chi2_variant, "direct") == 0 ? -1 : 0));
        register_jacobian = JacobianInRegisters<Model>::value;
        if (jac != nullptr && std::strcmp(jac, "sp") == 0) {
            pipelined_jacobian = true;
            register_jacobian = false;
        } else if (jac != nullptr && std::strcmp(jac, "ws") == 0) {
            pipelined_jacobian = false;
            register_jacobian = false;
        } else if (jac != nullptr && std::strcmp(jac, "reg") == 0) {
            register_jacobian = kHasRegisterJacobian;
        }
        if constexpr (Chi2Shape<Model>::staged) {
            GADFIT_CUDA_TRY(cudaFuncSetAttribute(k_chi2<Model, true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(Chi2Shape<Model>::smem_bytes)));
            GADFIT_CUDA_TRY(cudaFuncSetAttribute(k_chi2<Model, true>,
                                                 cudaFuncAttributePreferredSharedMemoryCarveout,
                                                 cudaSharedmemCarveoutMaxShared));
            int staged_occupancy { 0 };
            GADFIT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
              &staged_occupancy, k_chi2<Model, true>, kThreads, Chi2Shape<Model>::smem_bytes));
            chi2_staged_fits = staged_occupancy >= 1;
        }
        // (the grid is sized for the direct-load variant; the staged one has the
        // same or a lower occupancy and then simply runs more than one wave)
        GADFIT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &occupancy[0], k_chi2<Model, false>, kThreads, 0));
        if (register_jacobian) {
            if constexpr (kHasRegisterJacobian) {
                GADFIT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                  &occupancy[1], k_jacobian_reg<Model, true, false>, kThreads,
                  JacobianRegShape<Model>::smem_bytes));
            }
        } else if (pipelined_jacobian) {
            GADFIT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
              &occupancy[1], k_jacobian<Model, true, false>, kThreads, Shape::smem_bytes));
        } else {
            GADFIT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
              &occupancy[1], k_jacobian_ws<Model, true, false>, kJacThreads,
              WsShape::smem_bytes));
        }
        GADFIT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &occupancy[2], k_omega<Model>, kThreads, 0));
        for (int& o : occupancy) {
            o = o < 1 ? 1 : o;
        }
        small_smem_ready =
          cudaFuncSetAttribute(k_solve_small<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(kSmallSmemLimit))
          == cudaSuccess;
        if (!small_smem_ready) {
            cudaGetLastError(); // not fatal: the global-memory stages are used
        }
        device_ready = true;
    }

    [[nodiscard]] auto name() const -> const char* override { return Model::name; }
    [[nodiscard]] auto n_pars() const -> int override { return NP; }
    [[nodiscard]] auto ctas_per_sm(const int which) const -> int override
    {
        return occupancy[which];
    }
    [[nodiscard]] auto factor_bytes() const -> size_t override
    {
        return sizeof(LocalFactor<NP>);
    }
    [[nodiscard]] auto chi2_unroll() const -> int override { return Chi2Tuning<Model>::unroll; }

    auto chi2(const int grid,
              cudaStream_t st,
              const SweepArgs& a,
              double* partials,
              const TrialFinish& fin) const -> void override
    {
        if constexpr (Chi2Shape<Model>::staged) {
            const bool by_size { a.n_tiles >= static_cast<long long>(kChi2StagedMinTiles) * grid };
            const bool staged { chi2_force == 1 ? true : (chi2_force == -1 ? false : by_size) };
            if (chi2_staged_fits && staged) {
                k_chi2<Model, true><<<grid, kThreads, Chi2Shape<Model>::smem_bytes, st>>>(
                  a, partials, fin);
                return;
            }
        }
        k_chi2<Model, false><<<grid, kThreads, 0, st>>>(a, partials, fin);
    }
    auto jacobian(const int grid,
                  cudaStream_t st,
                  const SweepArgs& a,
                  const int loss,
                  const int* slot_base,
                  double* partials,
                  double* J_out,
                  double* r_out,
                  const long long n_points) const -> void override
    {
        const bool store { J_out != nullptr || r_out != nullptr };
        const bool linear { loss == GADFIT_B200_LOSS_LINEAR };
        if constexpr (kHasRegisterJacobian) {
            if (register_jacobian) {
                const auto launch { [&](auto* kernel) {
                    kernel<<<grid, kThreads, JacobianRegShape<Model>::smem_bytes, st>>>(
                      a, loss, slot_base, partials, J_out, r_out, n_points);
                } };
                if (store) {
                    linear ? launch(k_jacobian_reg<Model, true, true>)
                           : launch(k_jacobian_reg<Model, true, false>);
                } else {
                    linear ? launch(k_jacobian_reg<Model, false, true>)
                           : launch(k_jacobian_reg<Model, false, false>);
                }
                return;
            }
        }
        if (!pipelined_jacobian) {
            const auto launch { [&](auto* kernel) {
                kernel<<<grid, kJacThreads, WsShape::smem_bytes, st>>>(
                  a, loss, slot_base, partials, J_out, r_out, n_points);
            } };
            if (store) {
                linear ? launch(k_jacobian_ws<Model, true, true>)
                       : launch(k_jacobian_ws<Model, true, false>);
            } else {
                linear ? launch(k_jacobian_ws<Model, false, true>)
                       : launch(k_jacobian_ws<Model, false, false>);
            }
            return;
        }
        const auto launch { [&](auto* kernel) {
            kernel<<<grid, kThreads, Shape::smem_bytes, st>>>(
              a, loss, slot_base, partials, J_out, r_out, n_points);
        } };
        if (store) {
            linear ? launch(k_jacobian<Model, true, true>) : launch(k_jacobian<Model, true, false>);
        } else {
            linear ? launch(k_jacobian<Model, false, true>)
                   : launch(k_jacobian<Model, false, false>);
        }
    }
    auto omega(const int grid,
               cudaStream_t st,
               const SweepArgs& a,
               const double* seed,
               const double* J,
               const long long n_points,
               const int* slot_base,
               double* partials,
               double* omega_out) const -> void override
    {
        k_omega<Model><<<grid, kThreads, 0, st>>>(
          a, seed, J, n_points, slot_base, partials, omega_out);
    }
    auto value(cudaStream_t st, const double* par, const double arg, double* out) const
      -> void override
    {
        k_value<Model><<<1, 1, 0, st>>>(par, arg, out);
    }

    static constexpr int kSetThreads { 128 };

    auto local_prepare(cudaStream_t st,
                       const double* M,
                       const unsigned* active,
                       const unsigned global_mask,
                       const int n_sets,
                       const int damp_max,
                       double* dtd_local,
                       double* contrib,
                       const int VG) const -> void override
    {
        // one warp per data set
        k_local_prepare<NP><<<blocks_for(static_cast<long long>(n_sets) * 32, kSetThreads),
                              kSetThreads,
                              0,
                              st>>>(
          M, active, global_mask, n_sets, damp_max, dtd_local, contrib, VG);
    }
    auto local_schur(cudaStream_t st,
                     const double* M,
                     const unsigned* active,
                     const unsigned global_mask,
                     const int n_sets,
                     const double* dtd_local,
                     const LmState* state,
                     void* factors,
                     double* contrib,
                     const int VG) const -> void override
    {
        k_local_schur<NP><<<blocks_for(n_sets, kSetThreads), kSetThreads, 0, st>>>(
          M,
          active,
          global_mask,
          n_sets,
          dtd_local,
          state,
          static_cast<LocalFactor<NP>*>(factors),
          contrib,
          VG);
    }
    auto local_schur_rhs(cudaStream_t st,
                         const double* v_sets,
                         const unsigned* active,
                         const unsigned global_mask,
                         const int n_sets,
                         void* factors,
                         double* contrib,
                         const int VG2) const -> void override
    {
        k_local_schur_rhs<NP><<<blocks_for(n_sets, kSetThreads), kSetThreads, 0, st>>>(
          v_sets,
          active,
          global_mask,
          n_sets,
          static_cast<LocalFactor<NP>*>(factors),
          contrib,
          VG2);
    }
    auto local_backsolve(cudaStream_t st,
                         const unsigned* active,
                         const unsigned global_mask,
                         const int n_sets,
                         const void* factors,
                         const double* delta_g,
                         double* delta_par) const -> void override
    {
        k_local_backsolve<NP><<<blocks_for(n_sets, kSetThreads), kSetThreads, 0, st>>>(
          active,
          global_mask,
          n_sets,
          static_cast<const LocalFactor<NP>*>(factors),
          delta_g,
          delta_par);
    }
    auto acc_ratio_local(cudaStream_t st,
                         const unsigned* active,
                         const unsigned global_mask,
                         const int n_sets,
                         const double* dtd_local,
                         const double* delta1_par,
                         const double* delta2_par,
                         double* contrib) const -> void override
    {
        k_acc_ratio_local<NP><<<blocks_for(n_sets, kSetThreads), kSetThreads, 0, st>>>(
          active, global_mask, n_sets, dtd_local, delta1_par, delta2_par, contrib);
    }
    auto solve_small(cudaStream_t st, const SmallSolveArgs& a) const -> void override
    {
        // shared-memory residency: one pass of the cluster and the records fit
        SmallSolveArgs b { a };
        const size_t bytes { SmallSolveSmem<NP>::bytes(a.VG) };
        b.in_smem = (small_smem_ready && a.n_sets <= kSmallCtas * kSmallSetThreads
                     && bytes <= kSmallSmemLimit)
                      ? 1
                      : 0;
        k_solve_small<NP><<<kSmallCtas, kSmallThreads, b.in_smem != 0 ? bytes : 0, st>>>(b); // one cluster
    }
};

// One translation unit per model group (model_group.cu compiled with
// -DGADFIT_MODEL_GROUP=k) defines make_model_ops_group<k>; solver.cu only
// dispatches, so the kernel instantiations build in parallel.
template <int GROUP>
auto make_model_ops_group(const char* model_name) -> std::unique_ptr<ModelOps>;

} // namespace gadfit_b200
