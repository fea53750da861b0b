// Data-parallel sweeps over all data points of all data sets (sm_100a).
//
//   k_chi2      value-only sweep                 reference lm_solver.cpp:513-529
//   k_jacobian  residual + gradient + normal eq. reference lm_solver.cpp:286-346
//   k_omega     second directional derivative    reference lm_solver.cpp:362-383
//
// LAYOUT IN HBM.  x, y, sigma of every data set this GPU owns are concatenated
// data-set-major into three arrays; a TILE is up to kTile consecutive points of
// ONE data set (tile_set / tile_off / tile_len).  A CTA owns a contiguous range
// of tiles, one thread per point, loads are fully coalesced.  Parameters of
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

#include "gadfit_b200/fit_models.h"

#include <cuda_runtime.h>

namespace gadfit_b200 {

constexpr int kTile { 256 };    // points per tile == threads per CTA
constexpr int kThreads { 256 };

struct SweepArgs
{
    const double* __restrict__ x;
    const double* __restrict__ y;
    const double* __restrict__ sig;
    const int* __restrict__ tile_set;       // data set (local index) per tile
    const long long* __restrict__ tile_off; // first point of the tile
    const int* __restrict__ tile_len;       // points in the tile (<= kTile)
    int n_tiles;
    const double* __restrict__ par; // [n_sets][NP] current parameters
};

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

// ------------------------------------------------------------------- chi2
//
// sum_i ((y_i - f(x_i)) / sigma_i)^2 with passive parameters; the loss
// function is NOT applied (lm_solver.cpp:521-524).  One partial per CTA.
template <class Model>
__global__ void __launch_bounds__(kThreads)
  k_chi2(const SweepArgs a, double* __restrict__ block_partials)
{
    constexpr int NP { Model::n_pars };
    __shared__ double s_par[NP];
    __shared__ double s_red[kThreads / 32];
    const int t0 { static_cast<int>(static_cast<long long>(a.n_tiles) * blockIdx.x
                                    / gridDim.x) };
    const int t1 { static_cast<int>(static_cast<long long>(a.n_tiles)
                                    * (blockIdx.x + 1) / gridDim.x) };
    double sum { 0.0 };
    int cur_set { -1 };
    for (int t = t0; t < t1; ++t) {
        const int s { a.tile_set[t] };
        if (s != cur_set) {
            __syncthreads();
            if (threadIdx.x < NP) {
                s_par[threadIdx.x] = a.par[static_cast<long long>(s) * NP + threadIdx.x];
            }
            __syncthreads();
            cur_set = s;
        }
        if (static_cast<int>(threadIdx.x) < a.tile_len[t]) {
            const long long i { a.tile_off[t] + threadIdx.x };
            const double f { value_of(Model {}(ValueCtx { s_par }, a.x[i])) };
            const double diff { (a.y[i] - f) / a.sig[i] };
            sum += diff * diff;
        }
    }
    const double total { block_sum(sum, s_red) };
    if (threadIdx.x == 0) {
        block_partials[blockIdx.x] = total;
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

template <class Model>
struct JacobianShape
{
    static constexpr int NP { Model::n_pars };
    static constexpr int NA { NP + 1 };
    static constexpr int NT { NA * (NA + 1) / 2 };
    // Odd leading dimension: conflict-free 64-bit stores of one row per thread
    static constexpr int LD { (NA % 2 == 0) ? NA + 1 : NA };
    // Triangle entries per thread / point groups per entry
    static constexpr int EPT { (NT + kThreads - 1) / kThreads };
    static constexpr int PG { NT >= kThreads ? 1 : kThreads / NT };
    static constexpr size_t smem_bytes { sizeof(double)
                                         * (static_cast<size_t>(kTile) * LD
                                            + (PG > 1 ? static_cast<size_t>(PG) * NT : 0)) };
};

// One thread per point evaluates f and its gradient (static-sparsity forward
// mode), forms the augmented row and stages it in shared memory; the CTA then
// accumulates the packed triangle of sum a a^T with every thread owning one
// triangle entry (and one of PG point groups when the triangle is smaller
// than the CTA).  If J_out != nullptr the weighted Jacobian rows are stored
// parameter-major (J_out[k * n_points + i]) for the omega sweep and
// getJacobian(); if r_out != nullptr so are the weighted residuals.
template <class Model>
__global__ void __launch_bounds__(kThreads)
  k_jacobian(const SweepArgs a,
             const int loss_type,
             const int* __restrict__ slot_base,
             double* __restrict__ partials,
             double* __restrict__ J_out,
             double* __restrict__ r_out,
             const long long n_points)
{
    using Shape = JacobianShape<Model>;
    constexpr int NP { Shape::NP };
    constexpr int NA { Shape::NA };
    constexpr int NT { Shape::NT };
    constexpr int LD { Shape::LD };
    constexpr int EPT { Shape::EPT };
    constexpr int PG { Shape::PG };
    extern __shared__ double s_dyn[];
    double* s_A { s_dyn };                 // [kTile][LD]
    double* s_grp { s_dyn + kTile * LD };  // [PG][NT] (PG > 1 only)
    __shared__ double s_par[NP];

    const int tid { static_cast<int>(threadIdx.x) };
    // Triangle entries owned by this thread
    int ei[EPT];
    int ej[EPT];
    double acc[EPT];
    const int q { PG > 1 ? tid / NT : 0 }; // point group
    const bool worker { PG > 1 ? (q < PG) : true };
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
        const int e { PG > 1 ? tid % NT : tid + m * kThreads };
        int i { 0 };
        int rem { e };
        while (i < NA - 1 && rem >= NA - i) {
            rem -= NA - i;
            ++i;
        }
        ei[m] = (e < NT) ? i : 0;
        ej[m] = (e < NT) ? i + rem : 0;
        acc[m] = 0.0;
    }

    const int t0 { static_cast<int>(static_cast<long long>(a.n_tiles) * blockIdx.x
                                    / gridDim.x) };
    const int t1 { static_cast<int>(static_cast<long long>(a.n_tiles)
                                    * (blockIdx.x + 1) / gridDim.x) };
    int slot { slot_base[blockIdx.x] };
    int cur_set { -1 };

    const auto flush { [&]() {
        // Sum the point groups in fixed order, then write one partial
        if constexpr (PG > 1) {
            __syncthreads();
            if (worker) {
                s_grp[q * NT + tid % NT] = acc[0];
            }
            __syncthreads();
            if (tid < NT) {
                double total { 0.0 };
#pragma unroll
                for (int g = 0; g < PG; ++g) {
                    total += s_grp[g * NT + tid];
                }
                partials[static_cast<long long>(slot) * NT + tid] = total;
            }
        } else {
#pragma unroll
            for (int m = 0; m < EPT; ++m) {
                const int e { tid + m * kThreads };
                if (e < NT) {
                    partials[static_cast<long long>(slot) * NT + e] = acc[m];
                }
            }
        }
#pragma unroll
        for (int m = 0; m < EPT; ++m) {
            acc[m] = 0.0;
        }
        ++slot;
    } };

    for (int t = t0; t < t1; ++t) {
        const int s { a.tile_set[t] };
        if (s != cur_set) {
            if (cur_set >= 0) {
                flush();
            }
            __syncthreads();
            if (tid < NP) {
                s_par[tid] = a.par[static_cast<long long>(s) * NP + tid];
            }
            cur_set = s;
        }
        __syncthreads(); // s_par visible; previous tile's s_A fully consumed
        const int len { a.tile_len[t] };
        if (tid < len) {
            const long long i { a.tile_off[t] + tid };
            const double inv_sig_den { a.sig[i] };
            const auto f { Model {}(GradCtx { s_par }, a.x[i]) };
            // lm_solver.cpp:303-310
            const double res { (a.y[i] - f.val) / inv_sig_den };
            const double w { loss_weight(loss_type, res) };
            double g[NP];
            dense_gradient<NP>(f, g);
            double* row { s_A + tid * LD };
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                // lm_solver.cpp:315-317
                const double jk { w * g[k] / inv_sig_den };
                row[k] = jk;
                if (J_out != nullptr) {
                    J_out[static_cast<long long>(k) * n_points + i] = jk;
                }
            }
            row[NP] = w * res;
            if (r_out != nullptr) {
                r_out[i] = w * res;
            }
        }
        __syncthreads();
        if (worker) {
#pragma unroll
            for (int m = 0; m < EPT; ++m) {
                const int i { ei[m] };
                const int j { ej[m] };
                double sum { acc[m] };
                for (int p = q; p < len; p += PG) {
                    sum = ::fma(s_A[p * LD + i], s_A[p * LD + j], sum);
                }
                acc[m] = sum;
            }
        }
    }
    if (cur_set >= 0) {
        flush();
    }
}

// ------------------------------------------------------------------ omega
//
// omega_i = f.dd / sigma_i along delta1 (lm_solver.cpp:366-377; no loss
// weight), immediately contracted with the stored weighted Jacobian rows:
// v_s[k] = sum_i J_ik omega_i (the dgemv of lm_solver.cpp:383), one partial
// vector of NP entries per (CTA, data-set run).
template <class Model>
__global__ void __launch_bounds__(kThreads)
  k_omega(const SweepArgs a,
          const double* __restrict__ seed, // [n_sets][NP] delta1 by parameter
          const double* __restrict__ J,    // [NP][n_points]
          const long long n_points,
          const int* __restrict__ slot_base,
          double* __restrict__ partials,   // [n_slots][NP]
          double* __restrict__ omega_out)
{
    constexpr int NP { Model::n_pars };
    __shared__ double s_par[NP];
    __shared__ double s_seed[NP];
    __shared__ double s_red[kThreads / 32];
    const int tid { static_cast<int>(threadIdx.x) };
    const int t0 { static_cast<int>(static_cast<long long>(a.n_tiles) * blockIdx.x
                                    / gridDim.x) };
    const int t1 { static_cast<int>(static_cast<long long>(a.n_tiles)
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
    for (int t = t0; t < t1; ++t) {
        const int s { a.tile_set[t] };
        if (s != cur_set) {
            if (cur_set >= 0) {
                flush();
            }
            __syncthreads();
            if (tid < NP) {
                s_par[tid] = a.par[static_cast<long long>(s) * NP + tid];
                s_seed[tid] = seed[static_cast<long long>(s) * NP + tid];
            }
            __syncthreads();
            cur_set = s;
        }
        if (tid < a.tile_len[t]) {
            const long long i { a.tile_off[t] + tid };
            const Dir2 f { Model {}(Dir2Ctx { s_par, s_seed }, a.x[i]) };
            const double omega { f.dd / a.sig[i] };
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
    *out = value_of(Model {}(ValueCtx { par }, arg));
}

} // namespace gadfit_b200
