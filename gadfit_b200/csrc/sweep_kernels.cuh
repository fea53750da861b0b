// Data-parallel sweeps over all data points of all data sets (sm_100a).
//
//   k_chi2      value-only sweep                 reference lm_solver.cpp:513-529
//   k_jacobian  residual + gradient + normal eq. reference lm_solver.cpp:286-346
//   k_omega     second directional derivative    reference lm_solver.cpp:362-383
//
// LAYOUT IN HBM.  x, y, sigma of every data set this GPU owns are concatenated
// data-set-major into three arrays; a TILE is up to kTile consecutive points of
// ONE data set (geometry per data set: set_tile_begin / set_point_begin).  A CTA
// owns a contiguous range of tiles, one thread per point, loads are fully
// coalesced.  Parameters of
// the current data set sit in shared memory.
//
// NORMAL EQUATIONS.  The reference stores the dense N x n Jacobian and calls
// dsyrk/dgemv on it (lm_solver.cpp:325,339).  Here each point contributes its
// AUGMENTED row a = [w g_0/sigma, ..., w g_{NP-1}/sigma, w r] (NA = NP+1
// entries, g over ALL model parameters of this data set, r the residual,
// w the loss weight) and the sweep accumulates the packed upper triangle of
// M_s = sum_i a_i a_i^T per DATA SET.  M_s holds every block the block-arrow
// system needs: J^T J (shared-shared, shared-local, local-local), J^T r and
// the weighted residual norm.  Partial sums are written per (CTA, data-set
// run) into `partials` at host-assigned slots and reduced in fixed order, so
// results are bitwise reproducible run to run.
#pragma once

#include "gadfit_b200.h"
#include "gadfit_b200/fit_models.h"

#include <cuda_runtime.h>

namespace gadfit_b200 {

// Device-resident state of the LM loop (lm_solver.cpp:441-502: lambda, the
// chi2 pair of the accept test, the decision of the last trial step)
struct LmState
{
    double lambda;
    double old_chi2;
    double new_chi2;
    double acc_ratio;
    int decision;   // 0 accept, 1 retry, 2 give up (gadfit_b200_trial)
    int i_lambda;   // consecutive rejections in this iteration
    int use_delta2; // acceleration passed the ratio test
    // Status bits raised by kernels during the fit (never cleared by them):
    // bit 0 (fit_function.h k_status_workspace): an adaptive integrator ran
    // out of panels; bit 1: a peer-memory collective timed out
    unsigned flags;
    // Device-driven loop (latency-bound regime): what the next trial slot has
    // to do -- the kernels of a slot read it and skip themselves
    int phase;       // kPhaseNewIteration / kPhaseRetry / kPhaseDone
    int iteration;   // 1-based number of the iteration in progress
    int n_trials;    // trial records written to the device trace
    int n_jacobians; // Jacobian sweeps that actually ran
};
constexpr unsigned kFlagWorkspace { 1U };
constexpr unsigned kFlagPeerTimeout { 2U };
constexpr int kPhaseNewIteration { 0 }; // Jacobian sweep + first solve of an iteration
constexpr int kPhaseRetry { 1 };        // re-solve with the raised lambda
constexpr int kPhaseDone { 2 };         // nothing left to do

// The accept / reject step of lm_solver.cpp:455-494; `rec` = the state before
// the decision + the decision (what the reference prints, :457)
__device__ __forceinline__ void decide_trial(LmState& s,
                                             LmState& rec,
                                             const int lambda_incs,
                                             const double lambda_down,
                                             const double lambda_up)
{
    rec = s;
    if (s.new_chi2 < s.old_chi2) {
        s.lambda /= lambda_down;
        s.old_chi2 = s.new_chi2;
        s.decision = 0;
        s.i_lambda = 0;
    } else if (s.i_lambda < lambda_incs) {
        s.lambda *= lambda_up;
        s.decision = 1;
        s.i_lambda += 1;
    } else {
        s.decision = 2;
        s.i_lambda = 0;
    }
    rec.decision = s.decision;
}

// What the LAST CTA of a chi2 sweep needs to finish the trial step itself
// (device-driven loop): final sum of the CTA partials in fixed order, accept /
// reject, revert of the parameters, trace record, phase of the next slot.
// state == nullptr: plain sweep, partials only.
struct TrialFinish
{
    LmState* state;
    unsigned* counter; // CTAs of this sweep that have written their partial
    int lambda_incs;
    int iteration_limit;
    double lambda_down;
    double lambda_up;
    long long n_par;
    const double* par_old;
    double* par;
    gadfit_b200_trial* trace;
    int trace_cap;
};

constexpr int kTile { 256 };    // points per tile == threads per CTA
constexpr int kThreads { 256 };
// Resident CTAs per SM the Jacobian sweep is compiled for (register cap 85)
#ifndef GADFIT_JAC_MIN_BLOCKS
#define GADFIT_JAC_MIN_BLOCKS 3
#endif
constexpr int kJacobianMinBlocks { GADFIT_JAC_MIN_BLOCKS };
// Tiles of one data set the chi2 sweep evaluates per loop iteration: a
// fitting function may state `static constexpr int chi2_unroll` (cheap
// closed-form functions profit from 8: instruction-level parallelism across
// points and shared parameter-only subexpressions; measured on BASELINE
// config 2: 0.77 / 0.53 / 0.48 ms per sweep for 1 / 4 / 8).  Functions with
// quadrature loops keep 1 -- their work per point is already long.
template <class Model, class = void>
struct Chi2Tuning
{
    static constexpr int unroll { 1 };
    static constexpr int min_blocks { 4 };
};
template <class Model>
struct Chi2Tuning<Model, std::void_t<decltype(Model::chi2_unroll)>>
{
    static constexpr int unroll { Model::chi2_unroll };
    // resident CTAs per SM the sweep is compiled for: `chi2_blocks` if the
    // function states it, else 2 (128 registers) from 8 points per thread
    static constexpr int min_blocks { [] {
        if constexpr (requires { Model::chi2_blocks; }) {
            return Model::chi2_blocks;
        } else {
            return Model::chi2_unroll >= 8 ? 2 : 4;
        }
    }() };
};
// Unrolled sweeps can stage x, y, 1/sigma of the NEXT group of tiles in shared
// memory (LDGSTS, two buffers of `unroll` tiles) while the current group is
// evaluated: the DRAM latency no longer sits between a group's loads and its
// arithmetic (r02b: 30 % of the stall samples of config 2 were long-scoreboard).
// Measured (r02i): config 2 (845 tiles per CTA) 486 -> 435 us per sweep;
// config 3 (54 tiles per CTA, a sweep of 27 us) 27 -> 33 us -- the exposed
// first fetch and the 96 KiB of shared memory per CTA cost more than the
// overlap gives.  Both variants are compiled; the launcher takes the staged
// one from kChi2StagedMinTiles tiles per CTA.
#ifndef GADFIT_CHI2_STAGED
#define GADFIT_CHI2_STAGED 1
#endif
constexpr int kChi2StagedMinTiles { 128 };
template <class Model>
struct Chi2Shape
{
    static constexpr int unroll { Chi2Tuning<Model>::unroll };
    static constexpr bool staged { unroll >= 4 && GADFIT_CHI2_STAGED != 0 };
    static constexpr size_t group_doubles { static_cast<size_t>(unroll) * 3 * 256 };
    static constexpr size_t smem_bytes { staged ? 2 * group_doubles * sizeof(double) : 0 };
};

struct SweepArgs
{
    const double* __restrict__ x;
    const double* __restrict__ y;
    const double* __restrict__ inv_sig; // 1 / sigma (inverted once, at upload)
    // Tile geometry per DATA SET (not per tile): data set s owns the tiles
    // [set_tile_begin[s], set_tile_begin[s+1]) and the points
    // [set_point_begin[s], set_point_begin[s+1]) of the concatenated arrays
    const int* __restrict__ set_tile_begin;
    const long long* __restrict__ set_point_begin;
    int n_sets;
    int n_tiles;
    const double* __restrict__ par; // [n_sets][NP] current parameters
    // Status word of the sweep (LmState::flags): bit 0 = an adaptive
    // integrator ran out of panels (fit_function.h, k_status_workspace)
    unsigned* status;
    // Device-driven loop: the sweep returns at once if bit (gate->phase) of
    // gate_skip is set (nullptr: never)
    const LmState* gate;
    int gate_skip;
    // Pipelined upload (first Jacobian sweep of a fit() whose data are still
    // arriving over PCIe): `arrived` = number of leading points of x / y /
    // sigma that are complete in HBM, written by the copy engine after every
    // chunk (nullptr: everything is there); first_touch: `inv_sig` still holds
    // sigma -- this sweep inverts it in place.
    const unsigned long long* arrived;
    int first_touch;
    // Host-computed range of every CTA of THIS launch: tiles
    // [cta_tile_begin[b], cta_tile_begin[b + 1]) and the data set of the first
    // one.  Replaces up to three binary searches per CTA over set_tile_begin --
    // 10 dependent global loads each for 1000 data sets, a visible share of a
    // 30 us sweep.  nullptr: the kernel computes the range itself.
    const int* cta_tile_begin;
    const int* cta_first_set;
};

// Spin (one thread's view) until the first `need` points have arrived
__device__ __forceinline__ void wait_arrived(const SweepArgs& a,
                                             const long long need,
                                             unsigned long long& seen)
{
    if (a.arrived == nullptr || static_cast<unsigned long long>(need) <= seen) {
        return;
    }
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.arrived) : "memory");
    } while (seen < static_cast<unsigned long long>(need));
}

__device__ __forceinline__ bool sweep_is_gated_off(const SweepArgs& a)
{
    return a.gate != nullptr && ((a.gate_skip >> a.gate->phase) & 1) != 0;
}

// Walks the tiles of a CTA in ascending order.  Offsets and lengths are
// arithmetic; memory is touched only when the data set changes.
struct TileWalker
{
    int set { -1 };
    int tile_begin { 0 };
    int tile_end { 0 };
    long long point_begin { 0 };
    long long point_end { 0 };

    __device__ __forceinline__ void load(const SweepArgs& a, const int s)
    {
        set = s;
        tile_begin = a.set_tile_begin[s];
        tile_end = a.set_tile_begin[s + 1];
        point_begin = a.set_point_begin[s];
        point_end = a.set_point_begin[s + 1];
    }
    // Data set that contains tile t (binary search, once per CTA)
    __device__ __forceinline__ void seek(const SweepArgs& a, const int t)
    {
        int lo { 0 };
        int hi { a.n_sets };
        while (hi - lo > 1) {
            const int mid { (lo + hi) >> 1 };
            if (a.set_tile_begin[mid] <= t) {
                lo = mid;
            } else {
                hi = mid;
            }
        }
        load(a, lo);
    }
    // t must not decrease and must be < n_tiles; empty data sets are skipped
    __device__ __forceinline__ void advance_to(const SweepArgs& a, const int t)
    {
        while (t >= tile_end) {
            load(a, set + 1);
        }
    }
    __device__ __forceinline__ long long off(const int t) const
    {
        return point_begin + static_cast<long long>(t - tile_begin) * kTile;
    }
    __device__ __forceinline__ int len(const int t) const
    {
        const long long rest { point_end - off(t) };
        return rest < kTile ? static_cast<int>(rest) : kTile;
    }
};

// Asynchronous 8-byte copy global -> shared (SASS: LDGSTS): no register is
// tied up while the DRAM request is in flight
__device__ __forceinline__ void cp_async_8(double* smem_dst, const double* gmem_src)
{
    const unsigned dst { static_cast<unsigned>(__cvta_generic_to_shared(smem_dst)) };
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// all but the most recent commit group have landed
__device__ __forceinline__ void cp_async_wait_all_but_one()
{
    asm volatile("cp.async.wait_group 1;" ::: "memory");
}

// Packed upper triangle, row major: (i, j), i <= j, of an NA x NA matrix
__host__ __device__ constexpr int tri_index(const int NA, const int i, const int j)
{
    return i * NA - (i * (i - 1)) / 2 + (j - i);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_down_sync(0xffffffffU, v, o);
    }
    return v;
}

// Sum over the CTA in a fixed order; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* s_red)
{
    v = warp_sum(v);
    const int lane { static_cast<int>(threadIdx.x) & 31 };
    const int warp { static_cast<int>(threadIdx.x) >> 5 };
    __syncthreads();
    if (lane == 0) {
        s_red[warp] = v;
    }
    __syncthreads();
    double total { 0.0 };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            total += s_red[w];
        }
    }
    return total;
}

// Square root of the derivative of the loss function, lm_solver.cpp:257-284
__device__ __forceinline__ double loss_weight(const int loss_type, const double res)
{
    switch (loss_type) {
    case 1: // cauchy
        return ::sqrt(1.0 / (1.0 + res * res));
    case 2: // huber
        if (res * res > 1.0) {
            return ::sqrt(1.0 / ::fabs(res));
        }
        return 1.0;
    case 0: // linear
        return 1.0;
    default:
        return 0.0;
    }
}

// Thread 0 of the CTA that finishes a trial: decision, trace record, phase of
// the next slot.  Returns the decision.
__device__ __forceinline__ int finish_trial_record(const TrialFinish& fin, const double new_chi2)
{
    LmState s { *fin.state };
    s.new_chi2 = new_chi2;
    LmState rec;
    decide_trial(s, rec, fin.lambda_incs, fin.lambda_down, fin.lambda_up);
    if (fin.trace != nullptr) {
        if (s.n_trials < fin.trace_cap) {
            fin.trace[s.n_trials] = { rec.lambda, rec.new_chi2, rec.old_chi2, s.iteration, s.decision };
        }
        s.n_trials += 1;
        if (s.decision == 0) { // lm_solver.cpp:462-468, 496-501
            if (s.iteration >= fin.iteration_limit) {
                s.phase = kPhaseDone;
            } else {
                s.iteration += 1;
                s.phase = kPhaseNewIteration;
            }
        } else {
            s.phase = s.decision == 1 ? kPhaseRetry : kPhaseDone;
        }
    }
    fin.state[0] = s;
    fin.state[1] = rec;
    return s.decision;
}

// Tail of the chi2 sweep in the device-driven loop: the CTA that writes the
// last partial sums them all in the order of k_sum_partials / k_finish_trial
// (so chi2 is bit-identical to the stage-wise path), decides and reverts.
__device__ __forceinline__ void finish_trial_in_last_block(const TrialFinish& fin,
                                                           const double* block_partials,
                                                           double* s_red)
{
    __shared__ int s_last;
    __shared__ int s_decision;
    const int tid { static_cast<int>(threadIdx.x) };
    if (tid == 0) {
        __threadfence(); // my partial before my ticket
        s_last = atomicAdd(fin.counter, 1U) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last == 0) {
        return;
    }
    __threadfence();
    double sum { 0.0 };
    for (int i = tid; i < static_cast<int>(gridDim.x); i += kThreads) {
        sum += __ldcg(block_partials + i);
    }
    const double total { block_sum(sum, s_red) };
    if (tid == 0) {
        *fin.counter = 0U; // ready for the next sweep
        s_decision = finish_trial_record(fin, total);
    }
    __syncthreads();
    if (s_decision != 0) { // revertParameters, lm_solver.cpp:206-218
        for (long long i = tid; i < fin.n_par; i += kThreads) {
            fin.par[i] = fin.par_old[i];
        }
    }
}

// ------------------------------------------------------------------- chi2
//
// sum_i ((y_i - f(x_i)) / sigma_i)^2 with passive parameters; the loss
// function is NOT applied (lm_solver.cpp:521-524).  One partial per CTA.
// `sig` holds 1 / sigma_i (inverted once, at upload: the division leaves the
// sweeps).
template <class Model, bool kStaged>
__global__ void __launch_bounds__(kThreads, Chi2Tuning<Model>::min_blocks)
  k_chi2(const SweepArgs a, double* __restrict__ block_partials, const TrialFinish fin)
{
    if (sweep_is_gated_off(a)) {
        return;
    }
    constexpr int NP { Model::n_pars };
    constexpr int kChi2Unroll { Chi2Tuning<Model>::unroll };
    using Pre = prepared_t<Model, ValueCtx>;
    __shared__ double s_par[NP];
    __shared__ double s_red[kThreads / 32];
    __shared__ Pre s_prep; // prepare() of the current data set (empty without the hook)
    // Contiguous tile range of this CTA.  With kChi2Unroll > 1 the cut points
    // are moved up to the next multiple of kChi2Unroll tiles counted from the
    // start of the data set they fall into (or to its end): a range then never
    // starts or stops in the middle of an unrolled group, which would leave up
    // to 2 (kChi2Unroll - 1) tiles per CTA to the one-tile-at-a-time path
    // (config 3: 54 tiles per CTA, 16 per data set).
    const auto first_tile { [&](const int b) -> int {
        if (b >= static_cast<int>(gridDim.x)) {
            return a.n_tiles;
        }
        int t { static_cast<int>(static_cast<long long>(a.n_tiles) * b / gridDim.x) };
        if (kChi2Unroll > 1 && b > 0) {
            TileWalker w {};
            w.seek(a, t);
            w.advance_to(a, t);
            const int groups { (t - w.tile_begin + kChi2Unroll - 1) / kChi2Unroll };
            t = min(w.tile_begin + groups * kChi2Unroll, w.tile_end);
        }
        return t;
    } };
    const bool tabled { a.cta_tile_begin != nullptr };
    const int t0 { tabled ? a.cta_tile_begin[blockIdx.x] : first_tile(static_cast<int>(blockIdx.x)) };
    const int t1 { tabled ? a.cta_tile_begin[blockIdx.x + 1]
                          : first_tile(static_cast<int>(blockIdx.x) + 1) };
    double sum { 0.0 };
    const int tid { static_cast<int>(threadIdx.x) };
    TileWalker walk {};
    int cur_set { -1 };
    int t { t0 };
    if (t0 < t1) {
        if (tabled) {
            walk.load(a, a.cta_first_set[blockIdx.x]);
        } else {
            walk.seek(a, t0);
        }
    }
    // parameters (and prepare()) of the data set the walker stands on
    const auto enter_set { [&]() {
        __syncthreads();
        if (tid < NP) {
            s_par[tid] = a.par[static_cast<long long>(walk.set) * NP + tid];
        }
        __syncthreads();
        if constexpr (HasPrepare<Model, ValueCtx>) {
            // every evaluation thread fills the table entries tid, tid + 256, ...
            const Cooperative<ValueCtx> pc { ValueCtx { s_par, a.status }, tid };
            store_prepared(prepare_of(Model {}, pc), s_prep, tid);
            __syncthreads();
        }
        cur_set = walk.set;
    } };
    // W consecutive tiles of one data set (contiguous in memory; only the
    // last may be partial): the unrolled evaluations share every
    // parameter-only subexpression (1/w, ...) and constant.  Groups of
    // kChi2Unroll tiles, then 4, 2, 1 for what is left of the data set; a
    // thread adds its terms in ascending tile order whatever the grouping, so
    // the sum does not depend on it.
    const auto group_width { [&](const TileWalker& w, const int tile) -> int {
        const int avail { min(t1, w.tile_end) - tile };
        if (kChi2Unroll > 4 && avail >= kChi2Unroll) {
            return kChi2Unroll;
        }
        if (kChi2Unroll >= 4 && avail >= 4) {
            return 4;
        }
        if (kChi2Unroll >= 2 && avail >= 2) {
            return 2;
        }
        return 1;
    } };
    if constexpr (kStaged) {
        static_assert(Chi2Shape<Model>::staged);
        extern __shared__ double s_chi2_stage[]; // [2][kChi2Unroll][3][kTile]
        constexpr int kGroup { static_cast<int>(Chi2Shape<Model>::group_doubles) };
        // every thread copies, and later reads back, only its own elements
        const auto prefetch { [&](const TileWalker& w, const int tile, const int width, const int buf) {
            const int last_len { w.len(tile + width - 1) };
            const long long first { w.off(tile) + tid };
            double* dst { s_chi2_stage + buf * kGroup + tid };
#pragma unroll
            for (int u = 0; u < kChi2Unroll; ++u) {
                if (u < width && (u + 1 < width || tid < last_len)) {
                    const long long i { first + u * kTile };
                    cp_async_8(dst + (3 * u) * kTile, a.x + i);
                    cp_async_8(dst + (3 * u + 1) * kTile, a.y + i);
                    cp_async_8(dst + (3 * u + 2) * kTile, a.inv_sig + i);
                }
            }
            cp_async_commit();
        } };
        const auto evaluate_group { [&](auto width, const int buf) {
            constexpr int W { decltype(width)::value };
            const ValueCtx ctx { s_par, a.status };
            const int last_len { walk.len(t + W - 1) };
            const double* src { s_chi2_stage + buf * kGroup + tid };
            double x[W];
            double y[W];
            double sg[W];
#pragma unroll
            for (int u = 0; u < W; ++u) {
                const bool in { u + 1 < W || tid < last_len };
                x[u] = in ? src[(3 * u) * kTile] : 0.0;
                y[u] = in ? src[(3 * u + 1) * kTile] : 0.0;
                sg[u] = in ? src[(3 * u + 2) * kTile] : 1.0;
            }
#pragma unroll
            for (int u = 0; u < W; ++u) {
                const double f { value_of(evaluate_prepared(Model {}, ctx, s_prep, x[u])) };
                const double diff { (y[u] - f) * sg[u] };
                if (u + 1 < W || tid < last_len) {
                    sum += diff * diff;
                }
            }
        } };
        int width { 0 };
        int buf { 0 };
        if (t0 < t1) {
            walk.advance_to(a, t0);
            width = group_width(walk, t0);
            prefetch(walk, t0, width, 0);
        }
        while (t < t1) {
            walk.advance_to(a, t);
            if (walk.set != cur_set) {
                enter_set();
            }
            const int t_next { t + width };
            int width_next { 0 };
            if (t_next < t1) {
                TileWalker next { walk };
                next.advance_to(a, t_next);
                width_next = group_width(next, t_next);
                prefetch(next, t_next, width_next, buf ^ 1);
                cp_async_wait_all_but_one();
            } else {
                cp_async_wait_all();
            }
            if (width == kChi2Unroll) {
                evaluate_group(std::integral_constant<int, kChi2Unroll> {}, buf);
            } else if (width == 4) {
                evaluate_group(std::integral_constant<int, 4> {}, buf);
            } else if (width == 2) {
                evaluate_group(std::integral_constant<int, 2> {}, buf);
            } else {
                evaluate_group(std::integral_constant<int, 1> {}, buf);
            }
            t = t_next;
            width = width_next;
            buf ^= 1;
        }
    } else {
        while (t < t1) {
            walk.advance_to(a, t);
            if (walk.set != cur_set) {
                enter_set();
            }
            const ValueCtx ctx { s_par, a.status };
            const long long base { walk.off(t) + tid };
            const auto group { [&](auto width) {
                constexpr int W { decltype(width)::value };
                const int last_len { walk.len(t + W - 1) };
                if constexpr (W == 1) {
                    // threads past the end of the data set evaluate nothing
                    if (tid < last_len) {
                        const double f { value_of(evaluate_prepared(Model {}, ctx, s_prep, a.x[base])) };
                        const double diff { (a.y[base] - f) * a.inv_sig[base] };
                        sum += diff * diff;
                    }
                    ++t;
                    return;
                }
                double x[W];
                double y[W];
                double sg[W];
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const bool in { u + 1 < W || tid < last_len };
                    const long long i { base + u * kTile };
                    x[u] = in ? a.x[i] : 0.0;
                    y[u] = in ? a.y[i] : 0.0;
                    sg[u] = in ? a.inv_sig[i] : 1.0;
                }
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const double f { value_of(evaluate_prepared(Model {}, ctx, s_prep, x[u])) };
                    const double diff { (y[u] - f) * sg[u] };
                    if (u + 1 < W || tid < last_len) {
                        sum += diff * diff;
                    }
                }
                t += W;
            } };
            const int width { group_width(walk, t) };
            if (width == 1) {
                group(std::integral_constant<int, 1> {});
            } else if (width == 2) {
                group(std::integral_constant<int, 2> {});
            } else if (width == 4) {
                group(std::integral_constant<int, 4> {});
            } else {
                group(std::integral_constant<int, kChi2Unroll> {});
            }
        }
    }
    const double total { block_sum(sum, s_red) };
    if (threadIdx.x == 0) {
        block_partials[blockIdx.x] = total;
    }
    if (fin.state != nullptr) {
        finish_trial_in_last_block(fin, block_partials, s_red);
    }
}

// --------------------------------------------------------------- Jacobian
//
// Extracts the dense gradient over all NP model parameters from an SDual of
// any mask (parameters the function does not depend on give 0).
template <int NP, uint32_t M>
__device__ __forceinline__ void dense_gradient(const SDual<M>& f, double (&g)[NP])
{
    static_for<0, NP>([&](auto K) { g[K] = f.template tangent<K>(); });
}

// D(8x8) += A(8x4) * B(4x8), FP64 tensor path (SASS: DMMA.884)
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], const double a, const double b)
{
    asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// --------------------------------- Jacobian, software-pipelined per warp
//
// One thread per point evaluates f and its gradient (static-sparsity forward
// mode); each WARP contracts its own 32 rows on the FP64 tensor path
// (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4): for every k-step of 4 points, lane l
// loads J[point l%4][column 8b + l/4] for the NB column blocks -- the same
// register is the A fragment (J^T, 8 x 4) of block row b and the B fragment
// (J, 4 x 8) of block column b -- and issues one DMMA per upper-triangle block.
//
// The two jobs of a warp are SOFTWARE-PIPELINED: while the rows of tile t are
// contracted out of staging buffer `cur`, the same warp evaluates its point of
// tile t+1 in registers and stores the new row into buffer `cur ^ 1` afterwards.
// The evaluation (dependent DFMA chains: exp, polynomials, tangents) and the
// contraction (48 DMMA in 6 independent accumulator chains for 21 columns)
// are independent instruction streams inside one basic block, so the hardware
// always has something from the other stream to issue while one waits on a
// result -- both share the FP64 datapath, which is the resource to keep busy.
// A warp only reads rows it wrote itself: the steady state has __syncwarp()
// only, no block-wide or named barrier.  Accumulators (2 doubles per block
// and lane) stay in registers across all tiles of a data set; the 8 warps are
// summed in fixed order when the data set changes.  128 registers per thread,
// two CTAs per SM.
//
// If J_out != nullptr the weighted Jacobian rows are stored parameter-major
// (J_out[k * n_points + i]) for the omega sweep and getJacobian(); if
// r_out != nullptr so are the weighted residuals (kStore instantiation only).
// Resident CTAs per SM the software-pipelined Jacobian sweep is compiled for
// (register cap 65536 / (256 * blocks)).  Default 2 (128 registers: the
// quadrature models want them); a cheap closed-form function may state
// `static constexpr int jacobian_blocks` -- with ~100 FP64 instructions per
// point the sweep is latency bound and more resident warps pay.
template <class Model, class = void>
struct JacobianTuning
{
    static constexpr int min_blocks { 2 };
};
#ifndef GADFIT_NO_JACOBIAN_HINT
template <class Model>
struct JacobianTuning<Model, std::void_t<decltype(Model::jacobian_blocks)>>
{
    static constexpr int min_blocks { Model::jacobian_blocks };
};
#endif

template <class Model>
struct JacobianShape
{
    static constexpr int NP { Model::n_pars };
    static constexpr int NA { NP + 1 };
    static constexpr int NT { NA * (NA + 1) / 2 };
    // The augmented row is cut into NB blocks of 8 columns (zero padded); the
    // upper triangle of M has NBT blocks of 8 x 8
    static constexpr int NB { (NA + 7) / 8 };
    static constexpr int NBT { NB * (NB + 1) / 2 };
    // Staging buffers, column major: A[column][point], leading dimension
    // kTile + 4 so that the DMMA fragment loads are bank-conflict free
    // ((row * LDP + point) mod 16 distinct over a half warp)
    static constexpr int LDP { kTile + 4 };
    static constexpr int kWarps { kThreads / 32 };
    using Pre = prepared_t<Model, GradCtx>;
    // dynamic shared memory, in doubles: two staging buffers of NA column
    // lines, one all-zero line (padding columns of the last block), the
    // scratch of the cross-warp sum, input prefetch, parameters, prepare()
    static constexpr size_t stage { static_cast<size_t>(NA) * LDP };
    static constexpr size_t zero_line { LDP };
    static constexpr size_t flush { static_cast<size_t>(kWarps) * NBT * 64 };
    static constexpr size_t prefetch { 2 * 3 * kTile };
    static constexpr size_t par { (NP + 1) & ~size_t { 1 } };
    static constexpr size_t prep { (sizeof(Pre) + 7) / 8 };
    // the flush scratch aliases the staging buffers (used only between runs)
    static constexpr size_t work { 2 * stage > flush ? 2 * stage : flush };
    static constexpr size_t smem_bytes {
        sizeof(double) * (work + zero_line + prefetch + par + prep)
    };
};

template <class Model, bool kStore, bool kLinearLoss>
__global__ void __launch_bounds__(kThreads, JacobianTuning<Model>::min_blocks)
  k_jacobian(const SweepArgs a,
             const int loss_type,
             const int* __restrict__ slot_base,
             double* __restrict__ partials,
             double* __restrict__ J_out,
             double* __restrict__ r_out,
             const long long n_points)
{
    if (sweep_is_gated_off(a)) {
        return;
    }
    using Shape = JacobianShape<Model>;
    using Pre = typename Shape::Pre;
    constexpr int NP { Shape::NP };
    constexpr int NA { Shape::NA };
    constexpr int NT { Shape::NT };
    constexpr int NB { Shape::NB };
    constexpr int NBT { Shape::NBT };
    constexpr int LDP { Shape::LDP };
    constexpr int kWarps { Shape::kWarps };
    extern __shared__ double s_dyn[];
    double* const s_stage { s_dyn };                       // [2][NA][LDP], reused by flush()
    double* const s_zero { s_stage + Shape::work };        // [LDP]
    double* const s_pre { s_zero + Shape::zero_line };     // [2][3][kTile]
    double* const s_par { s_pre + Shape::prefetch };       // [NP]
    Pre* const s_prep { reinterpret_cast<Pre*>(s_par + Shape::par) };

    const int tid { static_cast<int>(threadIdx.x) };
    const int lane { tid & 31 };
    const int warp { tid >> 5 };
    const int frag_row { lane >> 2 }; // column within a block of 8
    const int frag_pt { lane & 3 };   // point within a k-step of 4
    for (int i = tid; i < LDP; i += kThreads) {
        s_zero[i] = 0.0;
    }
    const bool tabled { a.cta_tile_begin != nullptr };
    const int t0 { tabled ? a.cta_tile_begin[blockIdx.x]
                          : static_cast<int>(static_cast<long long>(a.n_tiles) * blockIdx.x
                                             / gridDim.x) };
    const int t1 { tabled ? a.cta_tile_begin[blockIdx.x + 1]
                          : static_cast<int>(static_cast<long long>(a.n_tiles)
                                             * (blockIdx.x + 1) / gridDim.x) };
    if (t0 >= t1) {
        return;
    }
    int slot { slot_base[blockIdx.x] };

    double acc[NBT][2];
#pragma unroll
    for (int b = 0; b < NBT; ++b) {
        acc[b][0] = 0.0;
        acc[b][1] = 0.0;
    }
    // Sum the warps in fixed order and write one partial (packed triangle)
    const auto flush { [&]() {
        __syncthreads(); // every warp is done with the staging buffers
#pragma unroll
        for (int b = 0; b < NBT; ++b) {
            double* dst { s_stage + (warp * NBT + b) * 64 + frag_row * 8 + 2 * frag_pt };
            dst[0] = acc[b][0];
            dst[1] = acc[b][1];
            acc[b][0] = 0.0;
            acc[b][1] = 0.0;
        }
        __syncthreads();
        for (int e = tid; e < NT; e += kThreads) {
            int i { 0 };
            int rem { e };
            while (rem >= NA - i) {
                rem -= NA - i;
                ++i;
            }
            const int j { i + rem };
            const int blk { tri_index(NB, i >> 3, j >> 3) };
            const int within { (i & 7) * 8 + (j & 7) };
            double total { 0.0 };
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                total += s_stage[(w * NBT + blk) * 64 + within];
            }
            partials[static_cast<long long>(slot) * NT + e] = total;
        }
        __syncthreads();
        ++slot;
    } };

    // x, y, 1/sigma of the tile after next are copied asynchronously (LDGSTS)
    // into a double-buffered staging area: the DRAM latency hides behind a
    // whole tile of work without holding registers.  Every thread reads back
    // only what it requested itself.  (A ring of three slots -- two tiles in
    // flight -- was measured, r02k: config 3 57.6 vs 56.7 us per sweep, config 5
    // 8.91 vs 8.67 ms; the walker look-ahead costs what the deeper queue gives.)
    // Threads past the end of a data set's last tile fetch its last point: in
    // the pipelined block they evaluate a VALID point (an adaptive integral
    // over a degenerate dummy range would never converge) and zero their row.
    unsigned long long seen { 0ULL };
    const auto prefetch { [&](const TileWalker& w, const int t, const int slot_in) {
        wait_arrived(a, w.off(t) + w.len(t), seen);
        const long long i { w.off(t) + min(tid, w.len(t) - 1) };
        double* dst { s_pre + slot_in * 3 * kTile + tid };
        cp_async_8(dst, a.x + i);
        cp_async_8(dst + kTile, a.y + i);
        cp_async_8(dst + 2 * kTile, a.inv_sig + i);
        cp_async_commit();
    } };
    // The augmented row of this thread's point of tile t, in registers
    struct Row
    {
        double v[NA];
    };
    const auto evaluate_row { [&](const TileWalker& w, const int t, const int in_slot) -> Row {
        const bool in_range { tid < w.len(t) };
        const double* in { s_pre + in_slot * 3 * kTile + tid };
        // (threads past the end of the data set: the last point again, row
        // zeroed below -- no divergence inside the pipelined block)
        const double px { in[0] };
        const double py { in[kTile] };
        double inv_sig { in[2 * kTile] };
        if (a.first_touch != 0) { // pipelined upload: sigma -> 1 / sigma, in place
            inv_sig = 1.0 / inv_sig;
            if (in_range) {
                const_cast<double*>(a.inv_sig)[w.off(t) + tid] = inv_sig;
            }
        }
        const auto f { evaluate_prepared(Model {}, GradCtx { s_par, a.status }, *s_prep, px) };
        // lm_solver.cpp:303-310,315-317 with 1/sigma hoisted out of the row
        const double res { (py - f.val) * inv_sig };
        const double wgt { kLinearLoss ? 1.0 : loss_weight(loss_type, res) };
        const double scale { kLinearLoss ? inv_sig : wgt * inv_sig };
        double g[NP];
        dense_gradient<NP>(f, g);
        Row row;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            row.v[k] = in_range ? g[k] * scale : 0.0;
        }
        row.v[NP] = in_range ? wgt * res : 0.0;
        if constexpr (kStore) {
            if (in_range) {
                const long long i { w.off(t) + tid };
                if (J_out != nullptr) {
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        J_out[static_cast<long long>(k) * n_points + i] = row.v[k];
                    }
                }
                if (r_out != nullptr) {
                    r_out[i] = row.v[NP];
                }
            }
        }
        return row;
    } };
    const auto store_row { [&](const Row& row, const int buf) {
        double* column { s_stage + buf * Shape::stage + tid };
#pragma unroll
        for (int k = 0; k < NA; ++k) {
            column[k * LDP] = row.v[k];
        }
    } };
    // 8 k-steps of 4 points over this warp's 32 rows of staging buffer `buf`
    const int frag_in_line { warp * 32 + frag_pt };
    const bool last_is_padding { 8 * (NB - 1) + frag_row >= NA };
    const auto contract { [&](const int buf) {
        const double* stage { s_stage + buf * Shape::stage };
        const double* frag_src[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            frag_src[b] = stage + (8 * b + frag_row) * LDP + frag_in_line;
        }
        if (last_is_padding) {
            frag_src[NB - 1] = s_zero + frag_in_line;
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            double frag[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                frag[b] = frag_src[b][ks * 4];
            }
#pragma unroll
            for (int bi = 0; bi < NB; ++bi) {
#pragma unroll
                for (int bj = bi; bj < NB; ++bj) {
                    dmma_m8n8k4(acc[tri_index(NB, bi, bj)], frag[bi], frag[bj]);
                }
            }
        }
    } };

    TileWalker walk {};
    if (tabled) {
        walk.load(a, a.cta_first_set[blockIdx.x]);
    } else {
        walk.seek(a, t0);
    }
    prefetch(walk, t0, 0);
    int t { t0 };
    while (t < t1) {
        // ---- a run of tiles of one data set
        walk.advance_to(a, t);
        __syncthreads(); // everyone is done with the previous parameters
        if (tid < NP) {
            s_par[tid] = a.par[static_cast<long long>(walk.set) * NP + tid];
        }
        __syncthreads();
        if constexpr (HasPrepare<Model, GradCtx>) {
            // every evaluation thread fills the table entries tid, tid + 256, ...
            const Cooperative<GradCtx> pc { GradCtx { s_par, a.status }, tid };
            store_prepared(prepare_of(Model {}, pc), *s_prep, tid);
            __syncthreads();
        }
        const int run_begin { t };
        const int run_end { min(t1, walk.tile_end) };
        // Step k evaluates tile k + 1 (registers only) while tile k is
        // contracted, then publishes the new row; the first step of a run
        // (k = run_begin - 1) has nothing to contract, the last one nothing to
        // evaluate.  ONE call site of evaluate_row: with two (a prologue of
        // its own) the compiler turns the function body into a subroutine,
        // and the adaptive-integral models misbehaved in it (r02e).
        int cur { 1 }; // staging buffer being contracted; the new row goes to cur ^ 1
        for (int k = run_begin - 1; k < run_end; ++k) {
            const bool more { k + 1 < run_end };
            Row row;
            if (more) {
                const int in_slot { (k + 1 - t0) & 1 };
                cp_async_wait_all(); // the inputs of tile k + 1 have landed
                if (k + 2 < t1) {
                    TileWalker next { walk };
                    next.advance_to(a, k + 2);
                    prefetch(next, k + 2, in_slot ^ 1);
                }
                row = evaluate_row(walk, k + 1, in_slot);
            }
            if (k >= run_begin) {
                contract(cur);
            }
            if (more) {
                store_row(row, cur ^ 1);
            }
            __syncwarp();
            cur ^= 1;
        }
        t = run_end;
        flush();
    }
}

// --------------------------------- Jacobian, accumulated in registers
//
// For a function with few parameters and a cheap closed form (BASELINE config
// 3: 6 parameters, ~50 FP64 instructions per point) the 8 x 8 tensor tile is
// mostly waste -- NA = 7 columns need 28 products per point, one DMMA block
// computes 64, and the FP64 datapath then spends more cycles on the
// contraction (8 DMMA x 16 cycles per 32 points) than on the evaluation.  A
// thread already holds the augmented row of its point in registers: this
// kernel keeps the NT = NA (NA + 1) / 2 entries of M_s as per-lane
// accumulators and adds a_i a_j with one DFMA each, straight from the row --
// no staging buffers, no shared-memory traffic, no barrier in the steady
// state.  The lanes are summed by a shuffle tree and the warps in fixed order
// when the data set changes (deterministic).  Chosen by the model hint
// `static constexpr bool jacobian_in_registers` (the accumulators cost 2 NT
// registers: 56 for NA = 7).
template <class Model>
struct JacobianRegShape
{
    static constexpr int NP { Model::n_pars };
    static constexpr int NA { NP + 1 };
    static constexpr int NT { NA * (NA + 1) / 2 };
    static constexpr int kWarps { kThreads / 32 };
    // tiles per group (one batch of copies / one pass of the bookkeeping)
    static constexpr int kGroup { 4 };
    static constexpr size_t smem_bytes { 2 * kGroup * 3 * kTile * sizeof(double) };
};

template <class Model, class = void>
struct JacobianInRegisters : std::false_type
{};
template <class Model>
struct JacobianInRegisters<Model, std::void_t<decltype(Model::jacobian_in_registers)>>
  : std::bool_constant<Model::jacobian_in_registers>
{};

template <class Model, bool kStore, bool kLinearLoss>
__global__ void __launch_bounds__(kThreads, 2)
  k_jacobian_reg(const SweepArgs a,
                 const int loss_type,
                 const int* __restrict__ slot_base,
                 double* __restrict__ partials,
                 double* __restrict__ J_out,
                 double* __restrict__ r_out,
                 const long long n_points)
{
    if (sweep_is_gated_off(a)) {
        return;
    }
    using Shape = JacobianRegShape<Model>;
    using Pre = prepared_t<Model, GradCtx>;
    constexpr int NP { Shape::NP };
    constexpr int NA { Shape::NA };
    constexpr int NT { Shape::NT };
    constexpr int kWarps { Shape::kWarps };
    constexpr int U { Shape::kGroup };
    extern __shared__ double s_dyn[];
    double* const s_pre { s_dyn }; // [2][U][3][kTile]: x, y, 1/sigma of two groups of tiles
    __shared__ double s_par[NP];
    __shared__ double s_red[kWarps][NT];
    __shared__ Pre s_prep;
    const int tid { static_cast<int>(threadIdx.x) };
    const int lane { tid & 31 };
    const int warp { tid >> 5 };
    const bool tabled { a.cta_tile_begin != nullptr };
    const int t0 { tabled ? a.cta_tile_begin[blockIdx.x]
                          : static_cast<int>(static_cast<long long>(a.n_tiles) * blockIdx.x
                                             / gridDim.x) };
    const int t1 { tabled ? a.cta_tile_begin[blockIdx.x + 1]
                          : static_cast<int>(static_cast<long long>(a.n_tiles)
                                             * (blockIdx.x + 1) / gridDim.x) };
    if (t0 >= t1) {
        return;
    }
    int slot { slot_base[blockIdx.x] };
    double acc[NT];
#pragma unroll
    for (int e = 0; e < NT; ++e) {
        acc[e] = 0.0;
    }
    // lanes by a shuffle tree, warps in fixed order, one partial (packed triangle)
    const auto flush { [&]() {
#pragma unroll
        for (int e = 0; e < NT; ++e) {
            const double total { warp_sum(acc[e]) };
            acc[e] = 0.0;
            if (lane == 0) {
                s_red[warp][e] = total;
            }
        }
        __syncthreads();
        for (int e = tid; e < NT; e += kThreads) {
            double total { 0.0 };
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                total += s_red[w][e];
            }
            partials[static_cast<long long>(slot) * NT + e] = total;
        }
        __syncthreads();
        ++slot;
    } };
    // A group = up to U consecutive tiles of one data set (contiguous in
    // memory, only the last may be partial): one batch of asynchronous copies,
    // one pass of the walker / loop bookkeeping for U points per thread
    const auto group_width { [&](const TileWalker& w, const int tile) -> int {
        return min(U, min(t1, w.tile_end) - tile);
    } };
    unsigned long long seen { 0ULL };
    const auto prefetch { [&](const TileWalker& w, const int tile, const int width, const int buf) {
        const int last_len { w.len(tile + width - 1) };
        const long long first { w.off(tile) + tid };
        wait_arrived(a, w.off(tile + width - 1) + last_len, seen);
        double* dst { s_pre + buf * (U * 3 * kTile) + tid };
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u < width && (u + 1 < width || tid < last_len)) {
                const long long i { first + u * kTile };
                cp_async_8(dst + (3 * u) * kTile, a.x + i);
                cp_async_8(dst + (3 * u + 1) * kTile, a.y + i);
                cp_async_8(dst + (3 * u + 2) * kTile, a.inv_sig + i);
            }
        }
        cp_async_commit();
    } };
    TileWalker walk {};
    if (tabled) {
        walk.load(a, a.cta_first_set[blockIdx.x]);
    } else {
        walk.seek(a, t0);
    }
    walk.advance_to(a, t0);
    int width { group_width(walk, t0) };
    int buf { 0 };
    prefetch(walk, t0, width, 0);
    int t { t0 };
    int cur_set { -1 };
    while (t < t1) {
        walk.advance_to(a, t);
        if (walk.set != cur_set) {
            if (cur_set >= 0) {
                flush();
            }
            __syncthreads(); // everyone is done with the previous parameters
            if (tid < NP) {
                s_par[tid] = a.par[static_cast<long long>(walk.set) * NP + tid];
            }
            __syncthreads();
            if constexpr (HasPrepare<Model, GradCtx>) {
                // every evaluation thread fills the table entries tid, tid + 256, ...
                const Cooperative<GradCtx> pc { GradCtx { s_par, a.status }, tid };
                store_prepared(prepare_of(Model {}, pc), s_prep, tid);
                __syncthreads();
            }
            cur_set = walk.set;
        }
        // the next group goes on its way, this one has landed
        const int t_next { t + width };
        int width_next { 0 };
        if (t_next < t1) {
            TileWalker next { walk };
            next.advance_to(a, t_next);
            width_next = group_width(next, t_next);
            prefetch(next, t_next, width_next, buf ^ 1);
            cp_async_wait_all_but_one();
        } else {
            cp_async_wait_all();
        }
        const int last_len { walk.len(t + width - 1) };
        const double* src { s_pre + buf * (U * 3 * kTile) + tid };
        const long long first { walk.off(t) + tid };
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u < width && (u + 1 < width || tid < last_len)) {
                const double px { src[(3 * u) * kTile] };
                const double py { src[(3 * u + 1) * kTile] };
                double inv_sig { src[(3 * u + 2) * kTile] };
                const long long i { first + u * kTile };
                if (a.first_touch != 0) { // pipelined upload: sigma -> 1 / sigma, in place
                    inv_sig = 1.0 / inv_sig;
                    const_cast<double*>(a.inv_sig)[i] = inv_sig;
                }
                const auto f {
                    evaluate_prepared(Model {}, GradCtx { s_par, a.status }, s_prep, px)
                };
                // lm_solver.cpp:303-310,315-317 with 1/sigma hoisted out of the row
                const double res { (py - f.val) * inv_sig };
                const double wgt { kLinearLoss ? 1.0 : loss_weight(loss_type, res) };
                const double scale { kLinearLoss ? inv_sig : wgt * inv_sig };
                double g[NP];
                dense_gradient<NP>(f, g);
                double row[NA];
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    row[k] = g[k] * scale;
                }
                row[NP] = wgt * res;
                if constexpr (kStore) {
                    if (J_out != nullptr) {
#pragma unroll
                        for (int k = 0; k < NP; ++k) {
                            J_out[static_cast<long long>(k) * n_points + i] = row[k];
                        }
                    }
                    if (r_out != nullptr) {
                        r_out[i] = row[NP];
                    }
                }
                // M_s += a a' (upper triangle, packed row major)
                static_for<0, NA>([&](auto I) {
                    static_for<I, NA>([&](auto J) {
                        constexpr int e { tri_index(NA, I, J) };
                        acc[e] = ::fma(row[I], row[J], acc[e]);
                    });
                });
            }
        }
        t = t_next;
        width = width_next;
        buf ^= 1;
    }
    flush();
}

// ------------------------------------------- Jacobian, warp-specialised
//
// k_jacobian above makes every warp alternate between evaluating its 32
// points and contracting them, so all 256 threads hold the 2 * NBT accumulator
// registers (24 of 80 for 21 columns) through the evaluation -- which then
// spills -- and a warp's two phases can only overlap with OTHER warps' phases.
// Here the two jobs belong to different warps of a 384-thread CTA:
//
//   warps 0-7   EVALUATION   one thread per point of the current tile:
//               f, grad f (static-sparsity forward mode), augmented row into
//               a double-buffered staging area [column][point]; no
//               accumulators, the whole register budget serves the function.
//   warps 8-11  CONTRACTION  warp m owns the rows of evaluation warps 2m and
//               2m+1 (64 points = 16 k-steps of mma.sync.m8n8k4.f64 per tile)
//               and keeps the accumulators of M_s for the whole run of tiles
//               of a data set; the four warps are summed in fixed order when
//               the data set changes.
//
// Hand-over through named barriers (bar.sync / bar.arrive, ids 1-4): FULL[b]
// = evaluation warps arrive after writing stage b, contraction warps wait;
// EMPTY[b] = the reverse.  The evaluation runs up to two tiles ahead.  The
// (warp_id % 4) placement of warps on the SM's four sub-partitions gives each
// sub-partition two evaluation warps and one contraction warp per CTA, two
// CTAs per SM.
//
// Zero padding of the last 8-column block is never stored: lanes whose
// fragment row lies beyond NA read from a shared all-zero line instead.
constexpr int kJacEvalWarps { 8 };
constexpr int kJacMmaWarps { 4 };
constexpr int kJacThreads { (kJacEvalWarps + kJacMmaWarps) * 32 };
static_assert(kJacEvalWarps * 32 == kTile);

template <class Model>
struct JacobianWsShape
{
    static constexpr int NP { Model::n_pars };
    static constexpr int NA { NP + 1 };
    static constexpr int NT { NA * (NA + 1) / 2 };
    static constexpr int NB { (NA + 7) / 8 };
    static constexpr int NBT { NB * (NB + 1) / 2 };
    static constexpr int LDP { kTile + 4 };
    using Pre = prepared_t<Model, GradCtx>;
    // dynamic shared memory, in doubles (config 2, 21 columns: 111.5 KiB, two
    // CTAs per SM with their 1 KiB system reserve = 225 of 228 KiB)
    static constexpr size_t stage { static_cast<size_t>(NA) * LDP };   // x 2
    static constexpr size_t zero_line { LDP };
    static constexpr size_t flush { static_cast<size_t>(kJacMmaWarps) * NBT * 64 };
    static constexpr size_t prefetch { 2 * 3 * kTile };
    static constexpr size_t par { (NP + 1) & ~size_t { 1 } };
    static constexpr size_t prep { (sizeof(Pre) + 7) / 8 };
    static constexpr size_t smem_bytes {
        sizeof(double) * (2 * stage + zero_line + flush + prefetch + par + prep)
    };
};

__device__ __forceinline__ void named_bar_sync(const int id, const int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(const int id, const int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <class Model, bool kStore, bool kLinearLoss>
__global__ void __launch_bounds__(kJacThreads, 2)
  k_jacobian_ws(const SweepArgs a,
                const int loss_type,
                const int* __restrict__ slot_base,
                double* __restrict__ partials,
                double* __restrict__ J_out,
                double* __restrict__ r_out,
                const long long n_points)
{
    if (sweep_is_gated_off(a)) {
        return;
    }
    using Shape = JacobianWsShape<Model>;
    using Pre = typename Shape::Pre;
    constexpr int NP { Shape::NP };
    constexpr int NA { Shape::NA };
    constexpr int NT { Shape::NT };
    constexpr int NB { Shape::NB };
    constexpr int NBT { Shape::NBT };
    constexpr int LDP { Shape::LDP };
    constexpr int kBarFull { 1 };  // + buffer
    constexpr int kBarEmpty { 3 }; // + buffer
    constexpr int kBarEval { 5 };
    constexpr int kBarMma { 6 };
    extern __shared__ double s_dyn[];
    double* const s_stage { s_dyn };                              // [2][NA][LDP]
    double* const s_zero { s_stage + 2 * Shape::stage };          // [LDP]
    double* const s_flush { s_zero + Shape::zero_line };          // [4][NBT][64]
    double* const s_pre { s_flush + Shape::flush };               // [2][3][kTile]
    double* const s_par { s_pre + Shape::prefetch };              // [NP]
    Pre* const s_prep { reinterpret_cast<Pre*>(s_par + Shape::par) };

    const int tid { static_cast<int>(threadIdx.x) };
    const int warp { tid >> 5 };
    for (int i = tid; i < LDP; i += kJacThreads) {
        s_zero[i] = 0.0;
    }
    __syncthreads();

    const bool tabled { a.cta_tile_begin != nullptr };
    const int t0 { tabled ? a.cta_tile_begin[blockIdx.x]
                          : static_cast<int>(static_cast<long long>(a.n_tiles) * blockIdx.x
                                             / gridDim.x) };
    const int t1 { tabled ? a.cta_tile_begin[blockIdx.x + 1]
                          : static_cast<int>(static_cast<long long>(a.n_tiles)
                                             * (blockIdx.x + 1) / gridDim.x) };
    if (t0 >= t1) {
        return;
    }
    TileWalker walk {};
    if (tabled) {
        walk.load(a, a.cta_first_set[blockIdx.x]);
    } else {
        walk.seek(a, t0);
    }
    int cur_set { -1 };

    if (warp < kJacEvalWarps) {
        // ------------------------------------------------------ evaluation
        unsigned long long seen { 0ULL };
        const auto prefetch { [&](const TileWalker& w, const int t, const int buf) {
            wait_arrived(a, w.off(t) + w.len(t), seen);
            if (tid < w.len(t)) {
                const long long i { w.off(t) + tid };
                double* dst { s_pre + buf * 3 * kTile + tid };
                cp_async_8(dst, a.x + i);
                cp_async_8(dst + kTile, a.y + i);
                cp_async_8(dst + 2 * kTile, a.inv_sig + i);
            }
            cp_async_commit();
        } };
        prefetch(walk, t0, 0);
        for (int t = t0; t < t1; ++t) {
            const int buf { (t - t0) & 1 };
            walk.advance_to(a, t);
            if (walk.set != cur_set) {
                named_bar_sync(kBarEval, kTile); // everyone is done with the old parameters
                if (tid < NP) {
                    s_par[tid] = a.par[static_cast<long long>(walk.set) * NP + tid];
                }
                named_bar_sync(kBarEval, kTile);
                if constexpr (HasPrepare<Model, GradCtx>) {
                    // every evaluation thread fills the table entries tid, tid + 256, ...
                    const Cooperative<GradCtx> pc { GradCtx { s_par, a.status }, tid };
                    store_prepared(prepare_of(Model {}, pc), *s_prep, tid);
                    named_bar_sync(kBarEval, kTile);
                }
                cur_set = walk.set;
            }
            const int len { walk.len(t) };
            const long long tile_off { walk.off(t) };
            cp_async_wait_all();
            const double px { s_pre[buf * 3 * kTile + tid] };
            const double py { s_pre[buf * 3 * kTile + kTile + tid] };
            double inv_sig { s_pre[buf * 3 * kTile + 2 * kTile + tid] };
            if (a.first_touch != 0 && tid < len) { // pipelined upload: invert sigma in place
                inv_sig = 1.0 / inv_sig;
                const_cast<double*>(a.inv_sig)[tile_off + tid] = inv_sig;
            }
            if (t + 1 < t1) {
                TileWalker next { walk };
                next.advance_to(a, t + 1);
                prefetch(next, t + 1, buf ^ 1);
            }
            if (t - t0 >= 2) {
                named_bar_sync(kBarEmpty + buf, kJacThreads); // stage `buf` has been contracted
            }
            double* column { s_stage + buf * Shape::stage + tid };
            if (tid < len) {
                const long long i { tile_off + tid };
                const auto f { evaluate_prepared(
                  Model {}, GradCtx { s_par, a.status }, *s_prep, px) };
                // lm_solver.cpp:303-310,315-317 with 1/sigma hoisted out of the row
                const double res { (py - f.val) * inv_sig };
                const double w { kLinearLoss ? 1.0 : loss_weight(loss_type, res) };
                const double scale { kLinearLoss ? inv_sig : w * inv_sig };
                double g[NP];
                dense_gradient<NP>(f, g);
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    const double jk { g[k] * scale };
                    column[k * LDP] = jk;
                    if constexpr (kStore) {
                        if (J_out != nullptr) {
                            J_out[static_cast<long long>(k) * n_points + i] = jk;
                        }
                    }
                }
                column[NP * LDP] = w * res;
                if constexpr (kStore) {
                    if (r_out != nullptr) {
                        r_out[i] = w * res;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < NA; ++k) {
                    column[k * LDP] = 0.0;
                }
            }
            named_bar_arrive(kBarFull + buf, kJacThreads);
        }
    } else {
        // ----------------------------------------------------- contraction
        const int lane { tid & 31 };
        const int m { warp - kJacEvalWarps }; // rows [64 m, 64 m + 64) of every tile
        const int mtid { tid - kJacEvalWarps * 32 };
        const int frag_row { lane >> 2 }; // column within a block of 8
        const int frag_pt { lane & 3 };   // point within a k-step of 4
        // fragment source of this lane within a column line; the lanes of the
        // padding columns of the last block read the all-zero line instead
        const int frag_in_line { 64 * m + frag_pt };
        const bool last_is_padding { 8 * (NB - 1) + frag_row >= NA };
        int slot { slot_base[blockIdx.x] };
        double acc[NBT][2];
#pragma unroll
        for (int b = 0; b < NBT; ++b) {
            acc[b][0] = 0.0;
            acc[b][1] = 0.0;
        }
        // Sum the contraction warps in fixed order, one partial (packed triangle)
        const auto flush { [&]() {
#pragma unroll
            for (int b = 0; b < NBT; ++b) {
                double* dst { s_flush + (m * NBT + b) * 64 + frag_row * 8 + 2 * frag_pt };
                dst[0] = acc[b][0];
                dst[1] = acc[b][1];
                acc[b][0] = 0.0;
                acc[b][1] = 0.0;
            }
            named_bar_sync(kBarMma, kJacMmaWarps * 32);
            for (int e = mtid; e < NT; e += kJacMmaWarps * 32) {
                int i { 0 };
                int rem { e };
                while (rem >= NA - i) {
                    rem -= NA - i;
                    ++i;
                }
                const int j { i + rem };
                const int blk { tri_index(NB, i >> 3, j >> 3) };
                const int within { (i & 7) * 8 + (j & 7) };
                double total { 0.0 };
#pragma unroll
                for (int w = 0; w < kJacMmaWarps; ++w) {
                    total += s_flush[(w * NBT + blk) * 64 + within];
                }
                partials[static_cast<long long>(slot) * NT + e] = total;
            }
            named_bar_sync(kBarMma, kJacMmaWarps * 32);
            ++slot;
        } };
        for (int t = t0; t < t1; ++t) {
            const int buf { (t - t0) & 1 };
            walk.advance_to(a, t);
            if (walk.set != cur_set) {
                if (cur_set >= 0) {
                    flush();
                }
                cur_set = walk.set;
            }
            named_bar_sync(kBarFull + buf, kJacThreads);
            const double* stage { s_stage + buf * Shape::stage };
            const double* frag_src[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                frag_src[b] = stage + (8 * b + frag_row) * LDP + frag_in_line;
            }
            if (last_is_padding) {
                frag_src[NB - 1] = s_zero + frag_in_line;
            }
            // rows beyond the end of a partial tile are zero; whole 32-row
            // halves beyond it are skipped
            const int rows { walk.len(t) - 64 * m };
            const int ksteps { rows >= 64 ? 16 : (rows > 32 ? 16 : (rows > 0 ? 8 : 0)) };
#pragma unroll 4
            for (int ks = 0; ks < ksteps; ++ks) {
                double frag[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    frag[b] = frag_src[b][ks * 4];
                }
#pragma unroll
                for (int bi = 0; bi < NB; ++bi) {
#pragma unroll
                    for (int bj = bi; bj < NB; ++bj) {
                        dmma_m8n8k4(acc[tri_index(NB, bi, bj)], frag[bi], frag[bj]);
                    }
                }
            }
            if (t + 2 < t1) {
                named_bar_arrive(kBarEmpty + buf, kJacThreads);
            }
        }
        flush();
    }
}

// ------------------------------------------------------------------ omega
//
// omega_i = f.dd / sigma_i along delta1 (lm_solver.cpp:366-377; no loss
// weight), immediately contracted with the stored weighted Jacobian rows:
// v_s[k] = sum_i J_ik omega_i (the dgemv of lm_solver.cpp:383), one partial
// vector of NP entries per (CTA, data-set run).
// Resident CTAs per SM the sweep is compiled for: `omega_blocks` if the
// function states it (config 4's single-integral function: 3, i.e. 80
// registers instead of 100 -- 1028 -> 990 us per sweep, r02u), else no cap
template <class Model>
constexpr int omega_min_blocks()
{
    if constexpr (requires { Model::omega_blocks; }) {
        return Model::omega_blocks;
    } else {
        return 1;
    }
}
template <class Model>
__global__ void __launch_bounds__(kThreads, omega_min_blocks<Model>())
  k_omega(const SweepArgs a,
          const double* __restrict__ seed, // [n_sets][NP] delta1 by parameter
          const double* __restrict__ J,    // [NP][n_points]
          const long long n_points,
          const int* __restrict__ slot_base,
          double* __restrict__ partials,   // [n_slots][NP]
          double* __restrict__ omega_out)
{
    constexpr int NP { Model::n_pars };
    using Pre = prepared_t<Model, Dir2Ctx>;
    __shared__ double s_par[NP];
    __shared__ double s_seed[NP];
    __shared__ double s_red[kThreads / 32];
    __shared__ Pre s_prep;
    const int tid { static_cast<int>(threadIdx.x) };
    const bool tabled { a.cta_tile_begin != nullptr };
    const int t0 { tabled ? a.cta_tile_begin[blockIdx.x]
                          : static_cast<int>(static_cast<long long>(a.n_tiles) * blockIdx.x
                                             / gridDim.x) };
    const int t1 { tabled ? a.cta_tile_begin[blockIdx.x + 1]
                          : static_cast<int>(static_cast<long long>(a.n_tiles)
                                             * (blockIdx.x + 1) / gridDim.x) };
    int slot { slot_base[blockIdx.x] };
    int cur_set { -1 };
    double v[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        v[k] = 0.0;
    }
    const auto flush { [&]() {
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const double total { block_sum(v[k], s_red) };
            if (tid == 0) {
                partials[static_cast<long long>(slot) * NP + k] = total;
            }
            v[k] = 0.0;
        }
        ++slot;
    } };
    TileWalker walk {};
    if (t0 < t1) {
        if (tabled) {
            walk.load(a, a.cta_first_set[blockIdx.x]);
        } else {
            walk.seek(a, t0);
        }
    }
    for (int t = t0; t < t1; ++t) {
        walk.advance_to(a, t);
        if (walk.set != cur_set) {
            if (cur_set >= 0) {
                flush();
            }
            __syncthreads();
            if (tid < NP) {
                s_par[tid] = a.par[static_cast<long long>(walk.set) * NP + tid];
                s_seed[tid] = seed[static_cast<long long>(walk.set) * NP + tid];
            }
            __syncthreads();
            if constexpr (HasPrepare<Model, Dir2Ctx>) {
                // every evaluation thread fills the table entries tid, tid + 256, ...
                const Cooperative<Dir2Ctx> pc { Dir2Ctx { s_par, s_seed, a.status }, tid };
                store_prepared(prepare_of(Model {}, pc), s_prep, tid);
                __syncthreads();
            }
            cur_set = walk.set;
        }
        if (tid < walk.len(t)) {
            const long long i { walk.off(t) + tid };
            const Dir2 f { evaluate_prepared(
              Model {}, Dir2Ctx { s_par, s_seed, a.status }, s_prep, a.x[i]) };
            const double omega { f.dd * a.inv_sig[i] };
            if (omega_out != nullptr) {
                omega_out[i] = omega;
            }
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                v[k] = ::fma(J[static_cast<long long>(k) * n_points + i], omega, v[k]);
            }
        }
    }
    if (cur_set >= 0) {
        flush();
    }
}

// Single evaluation f(arg) with the parameters of one data set
// (LMsolver::getValue, lm_solver.cpp:693-696)
template <class Model>
__global__ void k_value(const double* __restrict__ par, const double arg, double* out)
{
    *out = value_of(evaluate(Model {}, ValueCtx { par }, arg));
}

} // namespace gadfit_b200
