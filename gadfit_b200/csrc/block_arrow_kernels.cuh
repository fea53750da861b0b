// Block-arrow normal equations: batched per-data-set Cholesky, Schur
// complement onto the shared parameters, damped solve, step bookkeeping
// (sm_100a).
//
// The reference assembles the dense n x n matrix J^T J, adds lambda * DTD and
// Cholesky-factorises it (lm_solver.cpp:325-332,348-354,383-384 via
// lapack.cpp / lapack_fallback.cpp).  With n = p_g + sum_s p_l(s) this is
// infeasible for many data sets (n = 300 004 in BASELINE config 5).  The
// matrix is block-arrow:
//
//       [ G    C_1  C_2 ... ]   G   shared x shared   (sum over data sets)
//   A = [ C_1' L_1          ]   C_s shared x local_s
//       [ C_2'      L_2     ]   L_s local_s x local_s
//
// so (A + lambda D) delta = rhs is solved exactly by
//   L~_s = L_s + lambda D_s = R_s R_s'            (one thread per data set)
//   Y_s = R_s^-1 C_s',  z_s = R_s^-1 b_s
//   (G + lambda D_g - sum_s Y_s' Y_s) delta_g = b_g - sum_s Y_s' z_s
//   delta_s = R_s^-T (z_s - Y_s delta_g)
// Every block is read from the per-data-set augmented matrix M_s produced by
// the Jacobian sweep (sweep_kernels.cuh).
//
// Parameter roles of data set s are two bit masks over the NP model
// parameters: `active[s]` (lm_solver.h:37, Indices::active) and the solver-wide
// `global_mask` (lm_solver.h:42, Indices::global); local = active & ~global.
#pragma once

#include "sweep_kernels.cuh"
#ifdef GADFIT_SOLVE_SMALL_TIMING
#include <cstdio>
#endif

namespace gadfit_b200 {

// Per-data-set factorisation kept between the Schur step and the
// back-substitution of the same trial
template <int NP>
struct LocalFactor
{
    double R[NP * NP]; // lower Cholesky factor of L_s + lambda D_s (pl x pl)
    double Y[NP * NP]; // R^-1 C_s'  (pl x pg)
    double z[NP];      // R^-1 b_s   (pl)
};

template <int NP>
__device__ __forceinline__ void role_lists(const unsigned active,
                                           const unsigned global_mask,
                                           int (&idx_g)[NP],
                                           int& pg,
                                           int (&idx_l)[NP],
                                           int& pl)
{
    pg = 0;
    pl = 0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        if (((active >> k) & 1U) != 0U) {
            if (((global_mask >> k) & 1U) != 0U) {
                idx_g[pg++] = k;
            } else {
                idx_l[pl++] = k;
            }
        }
    }
}

__device__ __forceinline__ double m_entry(const double* __restrict__ M,
                                          const int NA,
                                          const int a,
                                          const int b)
{
    return a <= b ? M[tri_index(NA, a, b)] : M[tri_index(NA, b, a)];
}

// Sum the partial rows of each data set in slot order (fixed => deterministic)
// out[s][v] = sum_{slot in [first[s], first[s]+count[s])} partials[slot][v]
static __global__ void k_reduce_slots(const double* __restrict__ partials,
                               const int* __restrict__ slot_first,
                               const int* __restrict__ slot_count,
                               const int n_sets,
                               const int V,
                               double* __restrict__ out)
{
    const long long idx { static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x };
    if (idx >= static_cast<long long>(n_sets) * V) {
        return;
    }
    const int s { static_cast<int>(idx / V) };
    const int v { static_cast<int>(idx - static_cast<long long>(s) * V) };
    double sum { 0.0 };
    const int first { slot_first[s] };
    const int count { slot_count[s] };
    for (int c = 0; c < count; ++c) {
        sum += partials[static_cast<long long>(first + c) * V + v];
    }
    out[idx] = sum;
}

// sum_{r in [r0, r1)} in[r * stride + v] in ascending order.  The loads of 16
// rows are issued together and then added one by one: a plain loop is a chain
// of (L2 latency + add) per row.
__device__ __forceinline__ double ordered_column_sum(const double* in,
                                                     const int r0,
                                                     const int r1,
                                                     const long long stride,
                                                     const int v)
{
    constexpr int batch { 16 };
    double sum { 0.0 };
    int r { r0 };
    for (; r + batch <= r1; r += batch) {
        double x[batch];
#pragma unroll
        for (int u = 0; u < batch; ++u) {
            x[u] = in[static_cast<long long>(r + u) * stride + v];
        }
#pragma unroll
        for (int u = 0; u < batch; ++u) {
            sum += x[u];
        }
    }
    for (; r < r1; ++r) {
        sum += in[static_cast<long long>(r) * stride + v];
    }
    return sum;
}

// Column sums over rows in two fixed stages: stage 1, block b sums rows
// [b*chunk, (b+1)*chunk) for every column; stage 2 (gridDim == 1, chunk ==
// number of stage-1 blocks) sums the block results.
static __global__ void k_column_sums(const double* __restrict__ in,
                              const int n_rows,
                              const int V,
                              const int chunk,
                              double* __restrict__ out)
{
    const int r0 { static_cast<int>(blockIdx.x) * chunk };
    const int r1 { min(n_rows, r0 + chunk) };
    for (int v = static_cast<int>(threadIdx.x); v < V; v += blockDim.x) {
        out[static_cast<long long>(blockIdx.x) * V + v] = ordered_column_sum(in, r0, r1, V, v);
    }
}

// Final stage of the chi2 sweep: sum the CTA partials in order
static __global__ void k_sum_partials(const double* __restrict__ partials,
                               const int n,
                               double* __restrict__ out)
{
    __shared__ double s_red[kThreads / 32];
    double sum { 0.0 };
    for (int i = static_cast<int>(threadIdx.x); i < n; i += kThreads) {
        sum += partials[i];
    }
    const double total { block_sum(sum, s_red) };
    if (threadIdx.x == 0) {
        *out = total;
    }
}

// Once per iteration (lm_solver.cpp:326-329 restricted to the local columns
// + the shared-block rows of dsyrk/dgemv, :325,:339): update the local part of
// the damping matrix and emit this data set's contribution to the shared
// block: [G_s (pg x pg packed upper), b_g,s (pg)].
template <int NP>
__device__ __forceinline__ void local_prepare_set(const int s,
                                                  const int lane,
                                                  const int stride,
                                                  const double* __restrict__ M,
                                                  const unsigned* __restrict__ active,
                                                  const unsigned global_mask,
                                                  const int damp_max,
                                                  double* __restrict__ dtd_local, // [n_sets][NP]
                                                  double* __restrict__ contrib,   // [n_sets][VG]
                                                  const int VG)
{
    // `stride` lanes per data set over the output entries (pure gather)
    constexpr int NA { NP + 1 };
    constexpr int NT { NA * (NA + 1) / 2 };
    const double* Ms { M + static_cast<long long>(s) * NT };
    int idx_g[NP];
    int idx_l[NP];
    int pg;
    int pl;
    role_lists<NP>(active[s], global_mask, idx_g, pg, idx_l, pl);
    for (int a = lane; a < pl; a += stride) {
        const int k { idx_l[a] };
        const double jtj { Ms[tri_index(NA, k, k)] };
        double* d { dtd_local + static_cast<long long>(s) * NP + k };
        // std::max(DTD, JTJ), lm_solver.cpp:328
        *d = damp_max != 0 ? ((*d < jtj) ? jtj : *d) : jtj;
    }
    double* out { contrib + static_cast<long long>(s) * VG };
    const int ntri { pg * (pg + 1) / 2 };
    for (int o = lane; o < VG; o += stride) {
        if (o < ntri) {
            int a { 0 };
            int rem { o };
            while (rem >= pg - a) {
                rem -= pg - a;
                ++a;
            }
            out[o] = Ms[tri_index(NA, idx_g[a], idx_g[a + rem])];
        } else {
            out[o] = Ms[tri_index(NA, idx_g[o - ntri], NP)];
        }
    }
}

template <int NP>
__global__ void k_local_prepare(const double* __restrict__ M,
                                const unsigned* __restrict__ active,
                                const unsigned global_mask,
                                const int n_sets,
                                const int damp_max,
                                double* __restrict__ dtd_local,
                                double* __restrict__ contrib,
                                const int VG)
{
    const int s { static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) >> 5) };
    if (s < n_sets) {
        local_prepare_set<NP>(s, static_cast<int>(threadIdx.x) & 31, 32, M, active, global_mask,
                              damp_max, dtd_local, contrib, VG);
    }
}

// Per trial step: factorise the damped local block and emit the Schur
// contribution [Y'Y (pg x pg packed upper), Y'z (pg)].  The Cholesky loop
// order is that of the reference's in-tree dpptrf (lapack_fallback.cpp:59-81).
template <int NP, bool kBurst = true>
__device__ __forceinline__ void local_schur_into(LocalFactor<NP>& F,
                                                 const int s,
                                                 const double* __restrict__ M,
                                                 const unsigned* __restrict__ active,
                                                 const unsigned global_mask,
                                                 const double* __restrict__ dtd_local,
                                                 const double lambda,
                                                 double* __restrict__ contrib,
                                                 const int VG)
{
    constexpr int NA { NP + 1 };
    constexpr int NT { NA * (NA + 1) / 2 };
    const double* Ms { M + static_cast<long long>(s) * NT };
    // (small NP: all of M_s in one burst of loads instead of on demand inside
    // the dependent chains below)
    double m_local[NP <= 8 && kBurst ? NT : 1];
    if constexpr (NP <= 8 && kBurst) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            m_local[t] = Ms[t];
        }
        Ms = m_local;
    }
    int idx_g[NP];
    int idx_l[NP];
    int pg;
    int pl;
    role_lists<NP>(active[s], global_mask, idx_g, pg, idx_l, pl);
    // R R' = L_s + lambda D_s
#pragma unroll 1
    for (int i = 0; i < pl; ++i) {
#pragma unroll 1
        for (int j = 0; j < i; ++j) {
            double v { m_entry(Ms, NA, idx_l[i], idx_l[j]) };
#pragma unroll 1
            for (int k = 0; k < j; ++k) {
                v -= F.R[i * NP + k] * F.R[j * NP + k];
            }
            F.R[i * NP + j] = v / F.R[j * NP + j];
        }
        double v { Ms[tri_index(NA, idx_l[i], idx_l[i])]
                   + lambda * dtd_local[static_cast<long long>(s) * NP + idx_l[i]] };
#pragma unroll 1
        for (int k = 0; k < i; ++k) {
            v -= F.R[i * NP + k] * F.R[i * NP + k];
        }
        F.R[i * NP + i] = ::sqrt(v);
    }
    // Y = R^-1 C_s' and z = R^-1 b_s (forward substitution,
    // lapack_fallback.cpp:88-95)
#pragma unroll 1
    for (int g = 0; g <= pg; ++g) {
        const int col { g < pg ? idx_g[g] : NP };
#pragma unroll 1
        for (int i = 0; i < pl; ++i) {
            double v { m_entry(Ms, NA, idx_l[i], col) };
#pragma unroll 1
            for (int j = 0; j < i; ++j) {
                v -= F.R[i * NP + j] * (g < pg ? F.Y[j * NP + g] : F.z[j]);
            }
            v /= F.R[i * NP + i];
            if (g < pg) {
                F.Y[i * NP + g] = v;
            } else {
                F.z[i] = v;
            }
        }
    }
    double* out { contrib + static_cast<long long>(s) * VG };
    int o { 0 };
#pragma unroll 1
    for (int a = 0; a < pg; ++a) {
#pragma unroll 1
        for (int b = a; b < pg; ++b) {
            double v { 0.0 };
#pragma unroll 1
            for (int i = 0; i < pl; ++i) {
                v += F.Y[i * NP + a] * F.Y[i * NP + b];
            }
            out[o++] = v;
        }
    }
#pragma unroll 1
    for (int a = 0; a < pg; ++a) {
        double v { 0.0 };
#pragma unroll 1
        for (int i = 0; i < pl; ++i) {
            v += F.Y[i * NP + a] * F.z[i];
        }
        out[o++] = v;
    }
}

// The factorisation is a chain of dependent read-modify-writes of R, Y, z.
// Up to 8 parameters it runs in thread-local storage (L1-resident) and is
// stored once; in global memory every link of the chain is an L2 round trip.
template <int NP>
__device__ __forceinline__ void local_schur_set(const int s,
                                                const double* __restrict__ M,
                                                const unsigned* __restrict__ active,
                                                const unsigned global_mask,
                                                const double* __restrict__ dtd_local,
                                                const double lambda,
                                                LocalFactor<NP>* __restrict__ factors,
                                                double* __restrict__ contrib,
                                                const int VG,
                                                const bool in_shared = false)
{
    if (in_shared) {
        // M_s and the factor live in shared memory: the chain of dependent
        // read-modify-writes runs there directly (no burst copies)
        local_schur_into<NP, false>(
          factors[s], s, M, active, global_mask, dtd_local, lambda, contrib, VG);
        return;
    }
    if constexpr (NP <= 8) {
        LocalFactor<NP> F;
        local_schur_into<NP>(F, s, M, active, global_mask, dtd_local, lambda, contrib, VG);
        factors[s] = F;
    } else {
        local_schur_into<NP>(
          factors[s], s, M, active, global_mask, dtd_local, lambda, contrib, VG);
    }
}

template <int NP>
__global__ void k_local_schur(const double* __restrict__ M,
                              const unsigned* __restrict__ active,
                              const unsigned global_mask,
                              const int n_sets,
                              const double* __restrict__ dtd_local,
                              const LmState* __restrict__ state,
                              LocalFactor<NP>* __restrict__ factors,
                              double* __restrict__ contrib,
                              const int VG)
{
    const int s { static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x) };
    if (s < n_sets) {
        local_schur_set<NP>(
          s, M, active, global_mask, dtd_local, state->lambda, factors, contrib, VG);
    }
}

// Same factor, new right-hand side J^T omega (lm_solver.cpp:383-384):
// z'_s = R^-1 v_s[local]; contribution [v_s[global] (pg), Y' z' (pg)].
template <int NP>
__global__ void k_local_schur_rhs(const double* __restrict__ v_sets, // [n_sets][NP]
                                  const unsigned* __restrict__ active,
                                  const unsigned global_mask,
                                  const int n_sets,
                                  LocalFactor<NP>* __restrict__ factors,
                                  double* __restrict__ contrib, // [n_sets][2 pg]
                                  const int VG2)
{
    const int s { static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x) };
    if (s >= n_sets) {
        return;
    }
    int idx_g[NP];
    int idx_l[NP];
    int pg;
    int pl;
    role_lists<NP>(active[s], global_mask, idx_g, pg, idx_l, pl);
    LocalFactor<NP>& F { factors[s] };
    const double* vs { v_sets + static_cast<long long>(s) * NP };
#pragma unroll 1
    for (int i = 0; i < pl; ++i) {
        double v { vs[idx_l[i]] };
#pragma unroll 1
        for (int j = 0; j < i; ++j) {
            v -= F.R[i * NP + j] * F.z[j];
        }
        F.z[i] = v / F.R[i * NP + i];
    }
    double* out { contrib + static_cast<long long>(s) * VG2 };
#pragma unroll 1
    for (int a = 0; a < pg; ++a) {
        out[a] = vs[idx_g[a]];
    }
#pragma unroll 1
    for (int a = 0; a < pg; ++a) {
        double v { 0.0 };
#pragma unroll 1
        for (int i = 0; i < pl; ++i) {
            v += F.Y[i * NP + a] * F.z[i];
        }
        out[pg + a] = v;
    }
}

// Shared-parameter system on ONE WARP (pg <= 32), matrix in shared memory.
//   shared_sums = [G (packed upper), b_g]            (once per iteration)
//   schur_sums  = [sum Y'Y (packed upper), sum Y'z]  (per trial)
// which == 1: factorise G + lambda D_g - sum Y'Y and solve for delta1_g
//             (update_dtd != 0: first apply lm_solver.cpp:326-329 to D_g)
// which == 2: reuse the factor, rhs = rhs2_sums[0..pg) - rhs2_sums[pg..2pg)
// Lane i owns row i.  Column j of the factor is finished by all lanes at
// once (left-looking Cholesky); every inner sum runs over k in ascending
// order, i.e. exactly the operation order of the reference's sequential
// dpptrf / dpptrs loops (lapack_fallback.cpp:59-105), so the result is
// bit-identical to a one-thread implementation.
constexpr int kMaxShared { 32 };

// Called by exactly one full warp
__device__ __forceinline__ void shared_solve_warp(const int pg,
                                                  const int which,
                                                  const int update_dtd,
                                                  const int damp_max,
                                                  const double* __restrict__ shared_sums,
                                                  const double* __restrict__ schur_sums,
                                                  const double* __restrict__ rhs2_sums,
                                                  const double lambda,
                                                  double* __restrict__ dtd_global, // [pg]
                                                  double* __restrict__ chol,       // [pg*pg]
                                                  double* __restrict__ delta_g,    // [pg]
                                                  double (*s_L)[kMaxShared + 1])
{
    const int lane { static_cast<int>(threadIdx.x) & 31 };
    const int ntri { pg * (pg + 1) / 2 };
    const bool row { lane < pg };
    if (which == 1) {
        double dtd { 0.0 };
        if (row) {
            dtd = dtd_global[lane];
            if (update_dtd != 0) {
                const double jtj { shared_sums[tri_index(pg, lane, lane)] };
                dtd = damp_max != 0 ? ((dtd < jtj) ? jtj : dtd) : jtj;
                dtd_global[lane] = dtd;
            }
            // Row `lane` of the lower triangle of the damped, Schur-reduced matrix
            for (int j = 0; j <= lane; ++j) {
                const int o { tri_index(pg, j, lane) };
                double v { shared_sums[o] - schur_sums[o] };
                if (j == lane) {
                    v = shared_sums[o] + lambda * dtd - schur_sums[o];
                }
                s_L[lane][j] = v;
            }
        }
        __syncwarp();
        for (int j = 0; j < pg; ++j) {
            double v { 0.0 };
            if (row && lane >= j) {
                v = s_L[lane][j];
                for (int k = 0; k < j; ++k) {
                    v -= s_L[lane][k] * s_L[j][k];
                }
            }
            const double diag { ::sqrt(__shfl_sync(0xffffffffU, v, j)) };
            if (row && lane >= j) {
                s_L[lane][j] = lane == j ? diag : v / diag;
            }
            __syncwarp();
        }
        if (row) {
            for (int j = 0; j <= lane; ++j) {
                chol[lane * pg + j] = s_L[lane][j];
            }
        }
    } else {
        if (row) {
            for (int j = 0; j <= lane; ++j) {
                s_L[lane][j] = chol[lane * pg + j];
            }
        }
        __syncwarp();
    }
    // Forward substitution L y = rhs
    double v { 0.0 };
    if (row) {
        v = which == 1 ? shared_sums[ntri + lane] - schur_sums[ntri + lane]
                       : rhs2_sums[lane] - rhs2_sums[pg + lane];
    }
    for (int j = 0; j < pg; ++j) {
        const double yj { __shfl_sync(0xffffffffU, v, j) / s_L[j][j] };
        if (lane == j) {
            v = yj;
        } else if (row && lane > j) {
            v -= s_L[lane][j] * yj;
        }
    }
    // Backward substitution L' x = y
    for (int i = pg - 1; i >= 0; --i) {
        const double xi { __shfl_sync(0xffffffffU, v, i) / s_L[i][i] };
        if (lane == i) {
            v = xi;
        } else if (lane < i) {
            v -= s_L[i][lane] * xi;
        }
    }
    if (row) {
        delta_g[lane] = v;
    }
}

static __global__ void __launch_bounds__(32)
  k_shared_solve(const int pg,
                 const int which,
                 const int update_dtd,
                 const int damp_max,
                 const double* __restrict__ shared_sums,
                 const double* __restrict__ schur_sums,
                 const double* __restrict__ rhs2_sums,
                 const LmState* __restrict__ state,
                 double* __restrict__ dtd_global,
                 double* __restrict__ chol,
                 double* __restrict__ delta_g)
{
    __shared__ double s_L[kMaxShared][kMaxShared + 1];
    shared_solve_warp(pg, which, update_dtd, damp_max, shared_sums, schur_sums, rhs2_sums,
                      state->lambda, dtd_global, chol, delta_g, s_L);
}

// delta_s = R^-T (z - Y delta_g), scattered by model parameter index:
// delta_par[s][k] = step of parameter k of data set s (0 if inactive);
// global parameters get delta_g.
template <int NP>
__device__ __forceinline__ void local_backsolve_with(const LocalFactor<NP>& F,
                                                     const int s,
                                                     const unsigned* __restrict__ active,
                                                     const unsigned global_mask,
                                                     const double* __restrict__ delta_g,
                                                     double* __restrict__ delta_par) // [n_sets][NP]
{
    int idx_g[NP];
    int idx_l[NP];
    int pg;
    int pl;
    role_lists<NP>(active[s], global_mask, idx_g, pg, idx_l, pl);
    double d[NP];
#pragma unroll 1
    for (int i = 0; i < pl; ++i) {
        double v { F.z[i] };
#pragma unroll 1
        for (int a = 0; a < pg; ++a) {
            v -= F.Y[i * NP + a] * delta_g[a];
        }
        d[i] = v;
    }
    // Backward substitution, lapack_fallback.cpp:96-104
#pragma unroll 1
    for (int i = pl - 1; i >= 0; --i) {
        double v { d[i] };
#pragma unroll 1
        for (int j = pl - 1; j > i; --j) {
            v -= F.R[j * NP + i] * d[j];
        }
        d[i] = v / F.R[i * NP + i];
    }
    double* out { delta_par + static_cast<long long>(s) * NP };
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        out[k] = 0.0;
    }
    for (int a = 0; a < pg; ++a) {
        out[idx_g[a]] = delta_g[a];
    }
    for (int i = 0; i < pl; ++i) {
        out[idx_l[i]] = d[i];
    }
}

template <int NP>
__device__ __forceinline__ void local_backsolve_set(const int s,
                                                    const unsigned* __restrict__ active,
                                                    const unsigned global_mask,
                                                    const LocalFactor<NP>* __restrict__ factors,
                                                    const double* __restrict__ delta_g,
                                                    double* __restrict__ delta_par,
                                                    const bool in_shared = false)
{
    if (in_shared) {
        local_backsolve_with<NP>(factors[s], s, active, global_mask, delta_g, delta_par);
        return;
    }
    if constexpr (NP <= 8) {
        const LocalFactor<NP> F { factors[s] }; // one burst of independent loads
        local_backsolve_with<NP>(F, s, active, global_mask, delta_g, delta_par);
    } else {
        local_backsolve_with<NP>(factors[s], s, active, global_mask, delta_g, delta_par);
    }
}

template <int NP>
__global__ void k_local_backsolve(const unsigned* __restrict__ active,
                                  const unsigned global_mask,
                                  const int n_sets,
                                  const LocalFactor<NP>* __restrict__ factors,
                                  const double* __restrict__ delta_g,
                                  double* __restrict__ delta_par)
{
    const int s { static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x) };
    if (s < n_sets) {
        local_backsolve_set<NP>(s, active, global_mask, factors, delta_g, delta_par);
    }
}

// Local part of delta' D delta for delta1 and delta2 (lm_solver.cpp:388-393):
// contrib[s] = [sum_local d2^2 D, sum_local d1^2 D]
template <int NP>
__global__ void k_acc_ratio_local(const unsigned* __restrict__ active,
                                  const unsigned global_mask,
                                  const int n_sets,
                                  const double* __restrict__ dtd_local,
                                  const double* __restrict__ delta1_par,
                                  const double* __restrict__ delta2_par,
                                  double* __restrict__ contrib)
{
    const int s { static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x) };
    if (s >= n_sets) {
        return;
    }
    const unsigned local { active[s] & ~global_mask };
    double a2 { 0.0 };
    double a1 { 0.0 };
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        if (((local >> k) & 1U) != 0U) {
            const long long o { static_cast<long long>(s) * NP + k };
            const double D { dtd_local[o] };
            a2 += delta2_par[o] * (D * delta2_par[o]);
            a1 += delta1_par[o] * (D * delta1_par[o]);
        }
    }
    contrib[2 * s] = a2;
    contrib[2 * s + 1] = a1;
}

// acc_ratio = sqrt(d2'Dd2 / d1'Dd1) > threshold => delta2 := 0
// (lm_solver.cpp:394-397).  local_sums = all-reduced [a2, a1] of the local
// blocks; the shared block is added here.
static __global__ void k_acc_ratio_final(const int pg,
                                  const double threshold,
                                  const double* __restrict__ local_sums,
                                  const double* __restrict__ dtd_global,
                                  const double* __restrict__ delta1_g,
                                  const double* __restrict__ delta2_g,
                                  LmState* __restrict__ state)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) {
        return;
    }
    double a2 { 0.0 };
    double a1 { 0.0 };
    for (int a = 0; a < pg; ++a) {
        a2 += delta2_g[a] * (dtd_global[a] * delta2_g[a]);
        a1 += delta1_g[a] * (dtd_global[a] * delta1_g[a]);
    }
    a2 += local_sums[0];
    a1 += local_sums[1];
    const double ratio { ::sqrt(a2 / a1) };
    state->acc_ratio = ratio;
    state->use_delta2 = ratio > threshold ? 0 : 1;
}

// saveParameters + updateParameters (lm_solver.cpp:192-204,230-253): beta +=
// delta1, then (acceleration) beta -= 0.5 * delta2 as a second rounding step.
static __global__ void k_update_parameters(const long long n, // n_sets * NP
                                    const int save,
                                    const int accelerate,
                                    const LmState* __restrict__ state,
                                    const double* __restrict__ delta1_par,
                                    const double* __restrict__ delta2_par,
                                    double* __restrict__ par,
                                    double* __restrict__ par_old)
{
    const long long i { static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x };
    if (i >= n) {
        return;
    }
    double p { par[i] };
    if (save != 0) {
        par_old[i] = p;
    }
    p += delta1_par[i];
    if (accelerate != 0 && state->use_delta2 != 0) {
        p -= 0.5 * delta2_par[i];
    }
    par[i] = p;
}

// The accept / reject step of lm_solver.cpp:455-494 on the device
static __global__ void k_decide(const int lambda_incs,
                         const double lambda_down,
                         const double lambda_up,
                         LmState* __restrict__ state,
                         LmState* __restrict__ record) // pre-decision copy
{
    if (threadIdx.x != 0 || blockIdx.x != 0) {
        return;
    }
    LmState s { *state };
    LmState rec;
    decide_trial(s, rec, lambda_incs, lambda_down, lambda_up);
    *state = s;
    *record = rec;
}

// revertParameters (lm_solver.cpp:206-218), predicated on the device-side
// decision so no host round trip sits between k_decide and the revert
static __global__ void k_revert_if_rejected(const long long n,
                                     const LmState* __restrict__ state,
                                     const double* __restrict__ par_old,
                                     double* __restrict__ par)
{
    if (state->decision == 0) {
        return;
    }
    const long long i { static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x };
    if (i < n) {
        par[i] = par_old[i];
    }
}

// ---- fused kernels of the latency-bound regime ------------------------------
//
// With a few thousand data sets on one GPU (BASELINE config 3: 1000 x 4096)
// the sweeps take tens of microseconds and a trial step was a chain of
// sixteen few-microsecond launches.  Everything between the Jacobian sweep and
// the chi2 sweep of a trial is ONE launch here (k_solve_small), everything
// after the chi2 sweep another (k_finish_trial): four launches per iteration.
//
// k_solve_small is one thread-block cluster of 8 CTAs: the per-data-set work
// (one thread per data set, strided 200-byte records) is spread over eight
// SMs' load/store units, the stages are separated by the hardware cluster
// barrier (release/acquire at cluster scope, so global-memory results of one
// stage are visible to the next) instead of kernel boundaries.  A single CTA
// was tried first: 179 us per solve on config 3 against 62 us for the launch
// chain -- one SM cannot keep that many scattered loads in flight.
//
// Each stage is the per-data-set body of the kernel it replaces and the sums
// run in the same order (slots; 64-row chunks, then chunk results), so the
// results are bit-identical to the multi-kernel path.
constexpr int kSmallThreads { 512 };    // per CTA: slot reduction, column sums
constexpr int kSmallSetThreads { 128 }; // of which own one data set each per pass
constexpr int kSmallCtas { 8 };
// Beyond two passes of the cluster's 1024 set-threads the launch chain, which
// spreads over all SMs, wins again (config-3 shape, 1024-point data sets:
// 2000 data sets 0.183 vs 0.202 ms per iteration, 4000 data sets 0.306 vs 0.264)
constexpr int kSmallMaxSets { 2048 };
constexpr int kColumnChunk { 64 };
static_assert(kSmallSetThreads % kColumnChunk == 0, "a CTA owns whole chunks");

struct SmallSolveArgs
{
    int n_sets;
    int pg;
    int VG;             // pg (pg + 1) / 2 + pg
    unsigned global_mask;
    int prepare;        // first trial of an iteration: slots -> M, damping, shared sums
    int has_locals;     // some data set has local parameters (Schur step needed)
    int update;         // apply delta1 to the parameters (no acceleration)
    int damp_max;
    int gated;          // device-driven loop: `prepare` and whether to run at all come from state->phase
    const double* jac_partials;
    const int* slot_first;
    const int* slot_count;
    const unsigned* active;
    double* M;
    double* dtd_local;
    double* contrib;    // [2][n_sets][VG]: shared-block rows, Schur rows
    double* stage;      // [2][ceil(n_sets / 64)][VG]
    double* shared_sums;
    double* schur_sums;
    LmState* state;
    double* dtd_global;
    double* chol;
    double* delta_g;
    void* factors;
    double* delta_par;
    double* par;
    double* par_old;
    double* par_jac;
    int in_smem;        // one pass (n_sets <= 1024) and the per-data-set records fit in shared memory
};

// Shared-memory residency of k_solve_small (in_smem): every CTA keeps M_s, the
// local damping diagonal, both contribution rows and the local factors of ITS
// 128 data sets in shared memory between the stages, so the dependent chain
// slot sums -> prepare -> factorisation -> column sums -> back-substitution
// runs on ~30-cycle shared-memory accesses instead of L2 round trips (r02j:
// 44 us per solve on config 3 for 0.2 M warp instructions, 56 % of the stall
// samples at the stage barriers, 30 % long scoreboard).  Global memory still
// receives M and the damping diagonal (a rejected step re-solves from them
// with a new lambda) -- written, not read back.
template <int NP>
struct SmallSolveSmem
{
    static constexpr int NA { NP + 1 };
    static constexpr int NT { NA * (NA + 1) / 2 };
    // doubles per CTA for VG contribution entries per data set
    static constexpr size_t doubles(const int VG)
    {
        return static_cast<size_t>(kSmallSetThreads)
               * (NT + NP + 2 * static_cast<size_t>(VG) + sizeof(LocalFactor<NP>) / sizeof(double));
    }
    static constexpr size_t bytes(const int VG) { return doubles(VG) * sizeof(double); }
    static_assert(sizeof(LocalFactor<NP>) % sizeof(double) == 0);
};

// Stage 1 of k_column_sums for the chunks whose rows this CTA produced
// (data set s belongs to CTA (s / 128) % 8, so chunk b to CTA (b % 16) / 2)
__device__ __forceinline__ void column_sums_own_chunks(const int cta,
                                                       const double* in,
                                                       const int n_rows,
                                                       const int V,
                                                       double* stage)
{
    constexpr int per_cta { kSmallSetThreads / kColumnChunk };
    const int blocks { (n_rows + kColumnChunk - 1) / kColumnChunk };
    if (blocks <= 1) {
        return;
    }
    for (int b0 = per_cta * cta; b0 < blocks; b0 += per_cta * kSmallCtas) {
        for (int item = static_cast<int>(threadIdx.x); item < per_cta * V; item += kSmallThreads) {
            const int j { item / V };
            const int v { item - j * V };
            const int b { b0 + j };
            if (b < blocks) {
                const int r1 { min(n_rows, (b + 1) * kColumnChunk) };
                stage[static_cast<long long>(b) * V + v] =
                  ordered_column_sum(in, b * kColumnChunk, r1, V, v);
            }
        }
    }
}

// Stage 2 of k_column_sums by one CTA
// (by the threads [first, first + count) of the CTA; the result also goes to
// shared memory, where the shared solve reads it without an L2 round trip)
__device__ __forceinline__ void column_sums_final(const double* in,
                                                  const double* stage,
                                                  const int n_rows,
                                                  const int V,
                                                  double* out,
                                                  double* out_shared,
                                                  const int first,
                                                  const int count)
{
    const int blocks { (n_rows + kColumnChunk - 1) / kColumnChunk };
    const double* last_in { blocks > 1 ? stage : in };
    const int last_rows { blocks > 1 ? blocks : n_rows };
    const int t { static_cast<int>(threadIdx.x) - first };
    if (t < 0 || t >= count) {
        return;
    }
    for (int v = t; v < V; v += count) {
        const double sum { ordered_column_sum(last_in, 0, last_rows, V, v) };
        out[v] = sum;
        out_shared[v] = sum;
    }
}

#ifdef GADFIT_SOLVE_SMALL_TIMING
// Developer build (gadfit_b200.build.build_variant): wall-clock stamps of the
// stages of k_solve_small, printed by thread 0 of CTA 0
#define GADFIT_STAMP(k)                                                                   \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                            \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamp[k]));                      \
    }
#else
#define GADFIT_STAMP(k)
#endif

__device__ __forceinline__ void cluster_barrier()
{
    // release / acquire at cluster scope (the PTX defaults, spelled out)
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::
                   : "memory");
}

// k_reduce_slots + k_local_prepare + column sums (first trial), k_local_schur
// + column sums, k_shared_solve, k_local_backsolve, k_update_parameters
template <int NP>
__global__ void __cluster_dims__(kSmallCtas, 1, 1) __launch_bounds__(kSmallThreads, 1)
  k_solve_small(const SmallSolveArgs a_in)
{
    constexpr int NA { NP + 1 };
    constexpr int NT { NA * (NA + 1) / 2 };
    constexpr int cluster_threads { kSmallCtas * kSmallThreads };
    constexpr int sets_per_pass { kSmallCtas * kSmallSetThreads };
    SmallSolveArgs a { a_in };
    if (a.gated != 0) {
        const int phase { a.state->phase };
        if (phase == kPhaseDone) {
            return; // the whole cluster, before any barrier
        }
        a.prepare = phase == kPhaseNewIteration ? 1 : 0;
    }
    __shared__ double s_L[kMaxShared][kMaxShared + 1];
    constexpr int kMaxVG { kMaxShared * (kMaxShared + 1) / 2 + kMaxShared };
    __shared__ double s_sums[2 * kMaxVG]; // [G, b_g] and the Schur sums, as the shared solve reads them
    extern __shared__ double s_small[];
#ifdef GADFIT_SOLVE_SMALL_TIMING
    unsigned long long stamp[10] {};
#endif
    GADFIT_STAMP(0)
    const int tid { static_cast<int>(threadIdx.x) };
    const int cta { static_cast<int>(blockIdx.x) }; // grid == one cluster
    const int g { cta * kSmallThreads + tid };
    const int n_par { a.n_sets * NP };
    if (a.gated != 0 && a.prepare != 0 && g == 0) {
        a.state->n_jacobians += 1; // a Jacobian sweep ran in this slot
    }
    const int chunks { (a.n_sets + kColumnChunk - 1) / kColumnChunk };
    // the data set this thread owns in the first pass (threads >= 128: none)
    const int my_set { tid < kSmallSetThreads ? cta * kSmallSetThreads + tid : a.n_sets };
    LocalFactor<NP>* factors { static_cast<LocalFactor<NP>*>(a.factors) };
    double* contrib_shared { a.contrib };
    double* contrib_schur { a.contrib + static_cast<long long>(a.n_sets) * a.VG };
    double* stage_schur { a.stage + static_cast<long long>(chunks) * a.VG };
    // Where the stages read and write the per-data-set records: global memory,
    // or (in_smem) this CTA's shared-memory copies, addressed by the global
    // data-set index like the global arrays (base pointers shifted by set0)
    const int set0 { cta * kSmallSetThreads };
    const int n_own { max(0, min(kSmallSetThreads, a.n_sets - set0)) };
    const double* M_in { a.M };
    double* dtd { a.dtd_local };
    if (a.in_smem != 0) {
        double* s_M { s_small };
        double* s_dtd { s_M + kSmallSetThreads * NT };
        double* s_contrib { s_dtd + kSmallSetThreads * NP };
        double* s_contrib2 { s_contrib + kSmallSetThreads * a.VG };
        LocalFactor<NP>* s_F { reinterpret_cast<LocalFactor<NP>*>(s_contrib2 + kSmallSetThreads * a.VG) };
        // local damping diagonal of the own data sets (read-modify-write below)
        for (int i = tid; i < n_own * NP; i += kSmallThreads) {
            s_dtd[i] = a.dtd_local[static_cast<long long>(set0) * NP + i];
        }
        if (a.prepare == 0) {
            // a retry: M of the last Jacobian sweep
            for (int i = tid; i < n_own * NT; i += kSmallThreads) {
                s_M[i] = a.M[static_cast<long long>(set0) * NT + i];
            }
        }
        M_in = s_M - static_cast<long long>(set0) * NT;
        dtd = s_dtd - static_cast<long long>(set0) * NP;
        contrib_shared = s_contrib - static_cast<long long>(set0) * a.VG;
        contrib_schur = s_contrib2 - static_cast<long long>(set0) * a.VG;
        factors = s_F - set0;
    }
    if (a.prepare != 0) {
        // k_reduce_slots, four entries per thread at a time so that their
        // (dependent) index and partial loads overlap; each entry still adds
        // its slots in ascending order.  in_smem: every CTA reduces the slots
        // of its own data sets (no cluster barrier before the next stage).
        const int item0 { a.in_smem != 0 ? set0 * NT : 0 };
        const int n_items { a.in_smem != 0 ? (set0 + n_own) * NT : a.n_sets * NT };
        const int first_item { item0 + (a.in_smem != 0 ? tid : g) };
        const int item_stride { a.in_smem != 0 ? kSmallThreads : cluster_threads };
        double* M_out { const_cast<double*>(M_in) };
        for (int base = first_item; base < n_items; base += 4 * item_stride) {
            int first[4];
            int count[4];
            int col[4];
            double sum[4];
            int max_count { 0 };
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx { base + u * item_stride };
                const bool valid { idx < n_items };
                const int s { valid ? idx / NT : 0 };
                col[u] = idx - s * NT;
                first[u] = a.slot_first[s];
                count[u] = valid ? a.slot_count[s] : 0;
                max_count = max(max_count, count[u]);
                sum[u] = 0.0;
            }
            // (two slots per step: the loads of slot c + 1 do not wait for the
            // addition of slot c; the additions stay in ascending slot order)
            for (int c = 0; c < max_count; c += 2) {
                double p0[4];
                double p1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    p0[u] = c < count[u]
                              ? a.jac_partials[static_cast<long long>(first[u] + c) * NT + col[u]]
                              : 0.0;
                    p1[u] = c + 1 < count[u]
                              ? a.jac_partials[static_cast<long long>(first[u] + c + 1) * NT + col[u]]
                              : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (c < count[u]) {
                        sum[u] += p0[u];
                    }
                    if (c + 1 < count[u]) {
                        sum[u] += p1[u];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx { base + u * item_stride };
                if (idx < n_items) {
                    a.M[idx] = sum[u];
                    if (a.in_smem != 0) {
                        M_out[idx] = sum[u];
                    }
                }
            }
        }
        if (a.in_smem != 0) {
            __syncthreads();
        } else {
            cluster_barrier();
        }
        GADFIT_STAMP(1)
        // k_local_prepare.  More than one chunk: 4 lanes per data set, inside
        // the CTA that owns it (which then sums the rows of its chunks);
        // a single chunk (all of it owned by CTA 0, no stage 1): 8 lanes per data set
        if (chunks > 1) {
            constexpr int lanes { kSmallThreads / kSmallSetThreads };
            for (int s = cta * kSmallSetThreads + tid / lanes; s < a.n_sets; s += sets_per_pass) {
                local_prepare_set<NP>(s, tid % lanes, lanes, M_in, a.active, a.global_mask,
                                      a.damp_max, dtd, contrib_shared, a.VG);
            }
        } else if (cta == 0) {
            constexpr int lanes { kSmallThreads / kColumnChunk };
            if (tid / lanes < a.n_sets) {
                local_prepare_set<NP>(tid / lanes, tid % lanes, lanes, M_in, a.active,
                                      a.global_mask, a.damp_max, dtd, contrib_shared, a.VG);
            }
        }
        __syncthreads();
        if (a.in_smem != 0) {
            // the updated damping diagonal goes back to global memory (retries,
            // the next iteration's max rule, getDTD)
            for (int i = tid; i < n_own * NP; i += kSmallThreads) {
                a.dtd_local[static_cast<long long>(set0) * NP + i] = dtd[static_cast<long long>(set0) * NP + i];
            }
            if (chunks == 1 && cta == 0) {
                // column_sums_final reads the rows of a single chunk directly
                for (int i = tid; i < n_own * a.VG; i += kSmallThreads) {
                    a.contrib[i] = contrib_shared[i];
                }
            }
        }
        GADFIT_STAMP(2)
        column_sums_own_chunks(cta, contrib_shared, a.n_sets, a.VG, a.stage);
        GADFIT_STAMP(3)
    } else if (a.in_smem != 0) {
        __syncthreads(); // s_M, s_dtd of a retry are in place
    }
    if (a.prepare != 0 && tid >= kSmallSetThreads) {
        // the parameters of this Jacobian, copied by the threads that own no
        // data set (they idle through the factorisation below)
        constexpr int idle { kSmallThreads - kSmallSetThreads };
        for (int i = cta * idle + (tid - kSmallSetThreads); i < n_par; i += kSmallCtas * idle) {
            a.par_jac[i] = a.par[i];
        }
    }
    const double lambda { a.state->lambda };
    // (one pass: the current parameters of the own data set are fetched now,
    // not after the second cluster barrier)
    double p_now[NP];
    if (a.in_smem != 0 && a.update != 0 && my_set < a.n_sets) {
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            p_now[k] = a.par[my_set * NP + k];
        }
    }
    if (a.has_locals != 0) {
        for (int s = my_set; s < a.n_sets; s += sets_per_pass) {
            local_schur_set<NP>(s, M_in, a.active, a.global_mask, dtd, lambda, factors,
                                contrib_schur, a.VG, a.in_smem != 0);
        }
        __syncthreads();
        GADFIT_STAMP(4)
        if (a.in_smem != 0 && chunks == 1 && cta == 0) {
            for (int i = tid; i < n_own * a.VG; i += kSmallThreads) {
                a.contrib[static_cast<long long>(a.n_sets) * a.VG + i] = contrib_schur[i];
            }
        }
        column_sums_own_chunks(cta, contrib_schur, a.n_sets, a.VG, stage_schur);
    }
    GADFIT_STAMP(5)
    cluster_barrier();
    GADFIT_STAMP(6)
    if (cta == 0) {
        // The two sums by different halves of the CTA (their loads overlap),
        // kept in shared memory for the solve as well
        constexpr int half { kSmallThreads / 2 };
        if (a.prepare != 0) {
            column_sums_final(a.contrib, a.stage, a.n_sets, a.VG, a.shared_sums, s_sums, 0, half);
        } else if (tid < half) {
            for (int v = tid; v < a.VG; v += half) {
                s_sums[v] = a.shared_sums[v]; // a retry: the sums of this iteration
            }
        }
        if (a.has_locals != 0) {
            column_sums_final(a.contrib + static_cast<long long>(a.n_sets) * a.VG, stage_schur,
                              a.n_sets, a.VG, a.schur_sums, s_sums + kMaxVG, half, half);
        } else if (tid >= half) {
            for (int v = tid - half; v < a.VG; v += half) {
                s_sums[kMaxVG + v] = a.schur_sums[v]; // zeros
            }
        }
        __syncthreads();
        if (tid < 32 && a.pg > 0) {
            shared_solve_warp(a.pg, 1, a.prepare, a.damp_max, s_sums, s_sums + kMaxVG, nullptr,
                              lambda, a.dtd_global, a.chol, a.delta_g, s_L);
        }
    }
    GADFIT_STAMP(7)
    cluster_barrier();
    GADFIT_STAMP(8)
    for (int s = my_set; s < a.n_sets; s += sets_per_pass) {
        local_backsolve_set<NP>(s, a.active, a.global_mask, factors, a.delta_g, a.delta_par,
                                a.in_smem != 0);
        if (a.update != 0) {
            // saveParameters (first trial) + updateParameters,
            // lm_solver.cpp:192-204,230-253
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int i { s * NP + k };
                const double p { a.in_smem != 0 ? p_now[k] : a.par[i] };
                if (a.prepare != 0) {
                    a.par_old[i] = p;
                }
                a.par[i] = p + a.delta_par[i];
            }
        }
    }
#ifdef GADFIT_SOLVE_SMALL_TIMING
    GADFIT_STAMP(9)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        printf("solve_small stages (ns): slots %llu prepare %llu sums1 %llu schur %llu sums2 %llu "
               "barrier %llu shared_solve %llu barrier %llu backsolve %llu total %llu\n",
               stamp[1] - stamp[0], stamp[2] - stamp[1], stamp[3] - stamp[2], stamp[4] - stamp[3],
               stamp[5] - stamp[4], stamp[6] - stamp[5], stamp[7] - stamp[6], stamp[8] - stamp[7],
               stamp[9] - stamp[8], stamp[9] - stamp[0]);
    }
#endif
}

// k_sum_partials + k_decide + k_revert_if_rejected
static __global__ void __launch_bounds__(kThreads)
  k_finish_trial(const double* partials,
                 const int n_partials,
                 const int lambda_incs,
                 const double lambda_down,
                 const double lambda_up,
                 LmState* state,
                 const long long n_par,
                 const double* par_old,
                 double* par)
{
    __shared__ double s_red[kThreads / 32];
    __shared__ int s_decision;
    double sum { 0.0 };
    for (int i = static_cast<int>(threadIdx.x); i < n_partials; i += kThreads) {
        sum += partials[i];
    }
    const double total { block_sum(sum, s_red) };
    if (threadIdx.x == 0) {
        LmState s { *state };
        s.new_chi2 = total;
        LmState rec;
        decide_trial(s, rec, lambda_incs, lambda_down, lambda_up);
        state[0] = s;
        state[1] = rec;
        s_decision = s.decision;
    }
    __syncthreads();
    if (s_decision != 0) {
        for (long long i = threadIdx.x; i < n_par; i += kThreads) {
            par[i] = par_old[i];
        }
    }
}

// ---- the Fortran driver gadf_fit (fortran/gadfit/gadfit.F90:751-909) --------
//
// Scalars its accept test, lambda updates and convergence tests need, as sums
// over the active parameters (every global parameter counted once):
//   [0] delta1' D delta1       [1] delta1' D old_delta1   [2] old_delta1' D old_delta1
//   [3] delta1' J^T r          [4] |J^T r|^2
//   [5] #{ |delta1 / parameter| > rel_error }             (:882-894)
//   [6] #{ the same over global parameters, rel_error_global }  (:895-903)
//   [7] |r|^2 at the Jacobian point
constexpr int kGadfSums { 8 };

// One thread per data set: the local parameters' share
static __global__ void k_gadf_local(const int NP,
                                    const double* __restrict__ M,
                                    const unsigned* __restrict__ active,
                                    const unsigned global_mask,
                                    const int n_sets,
                                    const double* __restrict__ dtd_local,
                                    const double* __restrict__ delta1_par,
                                    const double* __restrict__ old_delta1_par,
                                    const double* __restrict__ par,
                                    const double rel_error,
                                    double* __restrict__ contrib)
{
    const int s { static_cast<int>(blockIdx.x * blockDim.x + threadIdx.x) };
    if (s >= n_sets) {
        return;
    }
    const int NA { NP + 1 };
    const int NT { NA * (NA + 1) / 2 };
    const double* Ms { M + static_cast<long long>(s) * NT };
    const unsigned local { active[s] & ~global_mask };
    double sums[kGadfSums] {};
    for (int k = 0; k < NP; ++k) {
        if (((local >> k) & 1U) != 0U) {
            const long long o { static_cast<long long>(s) * NP + k };
            const double D { dtd_local[o] };
            const double d1 { delta1_par[o] };
            const double d0 { old_delta1_par[o] };
            const double rhs { Ms[tri_index(NA, k, NP)] };
            sums[0] += d1 * (D * d1);
            sums[1] += d1 * (D * d0);
            sums[2] += d0 * (D * d0);
            sums[3] += d1 * rhs;
            sums[4] += rhs * rhs;
            if (rel_error >= 0.0 && fabs(d1 / par[o]) > rel_error) {
                sums[5] += 1.0;
            }
        }
    }
    sums[7] = Ms[tri_index(NA, NP, NP)];
    for (int j = 0; j < kGadfSums; ++j) {
        contrib[static_cast<long long>(s) * kGadfSums + j] = sums[j];
    }
}

// The global parameters' share, added by one thread on the rank that owns
// data set 0 (par_set0 = its parameter vector)
static __global__ void k_gadf_global(const int NP,
                                     const int pg,
                                     const unsigned global_mask,
                                     const double* __restrict__ shared_sums,
                                     const double* __restrict__ dtd_global,
                                     const double* __restrict__ delta1_g,
                                     const double* __restrict__ old_delta1_g,
                                     const double* __restrict__ par_set0,
                                     const double rel_error,
                                     const double rel_error_global,
                                     double* __restrict__ sums)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) {
        return;
    }
    int a { 0 };
    for (int k = 0; k < NP; ++k) {
        if (((global_mask >> k) & 1U) == 0U) {
            continue;
        }
        const double D { dtd_global[a] };
        const double d1 { delta1_g[a] };
        const double d0 { old_delta1_g[a] };
        const double rhs { shared_sums[pg * (pg + 1) / 2 + a] };
        sums[0] += d1 * (D * d1);
        sums[1] += d1 * (D * d0);
        sums[2] += d0 * (D * d0);
        sums[3] += d1 * rhs;
        sums[4] += rhs * rhs;
        const double rel { fabs(d1 / par_set0[k]) };
        if (rel_error >= 0.0 && rel > rel_error) {
            sums[5] += 1.0;
        }
        if (rel_error_global >= 0.0 && rel > rel_error_global) {
            sums[6] += 1.0;
        }
        ++a;
    }
}

} // namespace gadfit_b200
