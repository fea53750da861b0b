/* gadfit_b200 -- C-ABI of the B200-native Levenberg-Marquardt inner loop.
 *
 * Drop-in boundary for the hot path of GADfit's global multi-dataset fit:
 * every entry point below replaces one piece of the reference solver
 * interface gadfit::LMsolver (reference c++/gadfit/lm_solver.h:146-211) and
 * is what a cgo / JNI / ISO_C_BINDING / ctypes binding of that interface
 * would bind.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Call order and error behaviour are those of the reference:
 *   create -> add_dataset (all of them) -> set_par (...) -> [set_settings]
 *   -> fit -> get_par / chi2 / get_array ...
 * Every function returns a gadfit_b200_status; a non-zero status corresponds
 * to the reference exception of the same name (exceptions.h:21-80) or to a
 * device-side failure.  Nothing here ever falls back to a CPU path: without a
 * usable CUDA device fit()/chi2()/get_value() return GADFIT_B200_CUDA_ERROR.
 */
#ifndef GADFIT_B200_H
#define GADFIT_B200_H

#include <stddef.h>

#if defined(__GNUC__)
#define GADFIT_B200_API __attribute__((visibility("default")))
#else
#define GADFIT_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gadfit_b200_solver gadfit_b200_solver;

typedef enum gadfit_b200_status
{
    GADFIT_B200_OK = 0,
    /* reference exceptions, c++/gadfit/exceptions.h:24-78 */
    GADFIT_B200_UNKNOWN_OPERATION = 1,
    GADFIT_B200_LATE_ADD_DATASET_CALL = 2,    /* lm_solver.cpp:35-37 */
    GADFIT_B200_SET_PAR_INVALID_INDEX = 3,    /* lm_solver.cpp:65-68 */
    GADFIT_B200_UNINITIALIZED_PARAMETER = 4,  /* lm_solver.cpp:145-149 */
    GADFIT_B200_NEGATIVE_DEGREES_OF_FREEDOM = 5, /* lm_solver.cpp:137-139 */
    GADFIT_B200_NO_FITTING_PARAMETERS = 6,    /* lm_solver.cpp:187-189 */
    GADFIT_B200_NO_GLOBAL_PARAMETERS = 7,     /* lm_solver.cpp:129-131 */
    /* numerical_integration.cpp:304,525: an adaptive integrate() ran out of
     * panels at some data point.  The device workspace is 96 panels (1-D) /
     * 64 rectangles (2-D) per evaluation against the reference's default of
     * 1000 (numerical_integration.h:40); fit() / chi2() return this status,
     * the steps accepted before it stay in the solver. */
    GADFIT_B200_INSUFFICIENT_INTEGRATION_WORKSPACE = 8,
    /* conditions that have no reference counterpart */
    GADFIT_B200_UNKNOWN_MODEL = 9,
    GADFIT_B200_CUDA_ERROR = 10,
    GADFIT_B200_NCCL_ERROR = 11,
    GADFIT_B200_INVALID_ARGUMENT = 12,
    GADFIT_B200_NOT_AVAILABLE = 13,
    /* create(): the fitting function has more than 32 parameters (the
     * parameter roles and the static-sparsity masks are 32-bit words) */
    GADFIT_B200_TOO_MANY_PARAMETERS = 14
} gadfit_b200_status;

/* Loss functions, lm_solver.h:77-82 */
enum
{
    GADFIT_B200_LOSS_LINEAR = 0,
    GADFIT_B200_LOSS_CAUCHY = 1,
    GADFIT_B200_LOSS_HUBER = 2
};

/* POD mirror of LMsolver::settings, lm_solver.h:198-210 (defaults :88-96) */
typedef struct gadfit_b200_settings
{
    int iteration_limit;           /* 1000 */
    int lambda_incs;               /* 3 */
    double lambda_down;            /* 10 */
    double lambda_up;              /* 10 */
    int damp_max;                  /* 1 */
    const double* DTD_min;         /* NULL; used only if n_DTD_min > 1, */
    int n_DTD_min;                 /*   as in lm_solver.cpp:424-428 */
    double acceleration_threshold; /* -1 (off) */
    unsigned verbosity;            /* io::flag bits, lm_solver.h:60-72 */
    int loss;                      /* GADFIT_B200_LOSS_LINEAR */
    int n_threads;                 /* accepted for source compatibility; the
                                      data-parallel width is the GPU's */
} gadfit_b200_settings;

/* Arrays readable after fit(), lm_solver.h:188-194 (+ the step vectors) */
typedef enum gadfit_b200_array
{
    GADFIT_B200_JACOBIAN = 0,  /* dense N x n, row major; small problems only */
    GADFIT_B200_JTJ = 1,       /* dense n x n */
    GADFIT_B200_DTD = 2,       /* dense n x n (diagonal) */
    GADFIT_B200_LEFT_SIDE = 3, /* dense n x n, JTJ + lambda DTD */
    GADFIT_B200_RIGHT_SIDE = 4, /* n */
    GADFIT_B200_RESIDUALS = 5, /* N */
    GADFIT_B200_INV_JTJ = 6,   /* dense n x n, assembled from the block-arrow inverse */
    GADFIT_B200_DELTA1 = 7,    /* n */
    GADFIT_B200_DELTA2 = 8,    /* n (zeros unless acceleration is on) */
    GADFIT_B200_OMEGA = 9,     /* N (acceleration only) */
    /* diagonal of (J^T J)^-1, n entries: the parameter variances (times
     * chi2/dof).  From the block-arrow factors, so available for any n --
     * the dense GADFIT_B200_INV_JTJ is limited to n^2 <= 2^27 entries */
    GADFIT_B200_INV_JTJ_DIAGONAL = 10
} gadfit_b200_array;

/* One record per trial step: the reference prints exactly this at debug
 * level, "Current lambda: {}, new chi2 = {}, old chi2: {}"
 * (lm_solver.cpp:457-460).  decision: 0 = accepted (:462-468), 1 = rejected,
 * lambda raised and step recomputed (:470-481), 2 = rejected and given up
 * (:482-494). */
typedef struct gadfit_b200_trial
{
    double lambda;
    double new_chi2;
    double old_chi2;
    int iteration;
    int decision;
} gadfit_b200_trial;

/* The options that exist only in the Fortran driver gadf_fit
 * (fortran/gadfit/gadfit.F90:502-511; theory doc/user_guide.tex:125-216).
 * With enabled = 1, fit() follows gadf_fit (gadfit.F90:678-915) instead of
 * LMsolver::fit:
 *   - a step is accepted if (1 - beta)^uphill * new_chi2 < old_chi2, beta the
 *     D-weighted cosine between this and the previous delta1 (:756-764);
 *   - on acceptance lambda is divided by lambda_down, or updated by Nielsen's
 *     gain ratio (:765-770) or by Umrigar-Nightingale (:771-783);
 *   - on rejection lambda is raised, delta1 alone is re-solved and applied
 *     without delta2 (:785-808; LMsolver::fit recomputes both, lm_solver.cpp:476);
 *   - old_chi2 = min(old_chi2, new_chi2) (:823), at least one iteration runs,
 *     and after every iteration the convergence tests below are applied in
 *     this order (:832-909); a negative tolerance means "argument absent".
 * settings.iteration_limit is max_iter, lambda_incs / lambda_up / lambda_down /
 * damp_max / DTD_min / acceleration_threshold keep their meaning (the Fortran
 * defaults lambda = 1, lam_incs = 2 are the caller's to pass).  Error models
 * (gadf_set_errors, :445-470) are sigma arrays: SQRT_Y sqrt(y), PROPTO_Y y,
 * INVERSE_Y 1/y, USER as given. */
typedef struct gadfit_b200_gadf_options
{
    int enabled;
    int uphill;  /* exponent b; 0 = no uphill steps */
    int nielsen; /* 1 = on */
    int umnigh;  /* 1 = on */
    double chi2_abs;         /* stop if chi2/dof < chi2_abs */
    double chi2_rel;         /* stop if the relative decrease of chi2 < chi2_rel */
    double grad_chi2;        /* stop if 2 |J^T r| < grad_chi2 */
    double cos_phi;          /* stop if |cos(r, J delta1)| < cos_phi */
    double rel_error;        /* stop if all |delta1 / parameter| <= rel_error */
    double rel_error_global; /* the same over the global parameters only */
} gadfit_b200_gadf_options;

/* Why the last fit() ended */
typedef enum gadfit_b200_stop_reason
{
    GADFIT_B200_STOP_ITERATION_LIMIT = 0, /* lm_solver.cpp:496 / gadfit.F90:904-908 */
    GADFIT_B200_STOP_LAMBDA_INCS = 1,     /* lm_solver.cpp:482-494 / gadfit.F90:809-816 */
    GADFIT_B200_STOP_CHI2_ABS = 2,        /* gadfit.F90:832-837 */
    GADFIT_B200_STOP_CHI2_REL = 3,        /* :838-844 */
    GADFIT_B200_STOP_GRAD_CHI2 = 4,       /* :845-857 */
    GADFIT_B200_STOP_COS_PHI = 5,         /* :858-881 */
    GADFIT_B200_STOP_REL_ERROR = 6,       /* :882-894 */
    GADFIT_B200_STOP_REL_ERROR_GLOBAL = 7, /* :895-903 */
    /* chi2 was NaN or infinite at the start: every trial step is rejected
     * (new < old is false, lm_solver.cpp:462) and fit() gives up after
     * lambda_incs + 1 trials exactly as the reference does -- but the caller
     * can tell it apart from a converged fit */
    GADFIT_B200_STOP_NON_FINITE_CHI2 = 8
} gadfit_b200_stop_reason;

/* The five phase timers of the reference (lm_solver.h:140-144, table printed
 * by lm_solver.cpp:656-680): seconds of device time (CUDA events) and number
 * of calls, in the order Jacobian, chi2, linear algebra, omega, main loop. */
typedef struct gadfit_b200_timings
{
    double seconds[5];
    int calls[5];
} gadfit_b200_timings;

/* LMsolver::LMsolver(fitSignature), lm_solver.h:148.  The fitting function is
 * selected by name from the functions compiled into the library
 * (include/gadfit_b200/fit_models.h) or into a loaded model plug-in; a
 * std::function cannot run on the device. */
GADFIT_B200_API int gadfit_b200_create(const char* model_name, gadfit_b200_solver** out);
GADFIT_B200_API int gadfit_b200_destroy(gadfit_b200_solver* solver);

/* Number of registered fitting functions / name and parameter count of one */
GADFIT_B200_API int gadfit_b200_n_models(void);
GADFIT_B200_API int gadfit_b200_model_info(int i_model, const char** name, int* n_pars);

/* User-defined fitting functions -- the counterpart of handing the reference
 * an arbitrary lambda (lm_solver.h:146-148, c++/tests/example.cpp:11-19).
 * Loads a model plug-in: a shared library built by nvcc from functors
 * registered with GADFIT_B200_REGISTER_MODEL
 * (include/gadfit_b200/model_plugin.h; `python -m gadfit_b200.build --plugin`).
 * Its functions become available to gadfit_b200_create under their names.
 * GADFIT_B200_INVALID_ARGUMENT: not loadable, not a plug-in, built against
 * different headers, or a name is already registered.  Loading the same
 * library twice is a no-op.  Needs no GPU. */
GADFIT_B200_API int gadfit_b200_load_models(const char* plugin_path);

/* LMsolver::addDataset(n, x, y, errors = nullptr), lm_solver.h:164-167,
 * lm_solver.cpp:30-56.  Arrays are BORROWED until fit() has uploaded them
 * (first fit() or gadfit_b200_upload()).  With a communicator attached, x/y
 * may be NULL on ranks that do not own the dataset (see
 * gadfit_b200_owned_datasets). */
GADFIT_B200_API int gadfit_b200_add_dataset(gadfit_b200_solver* solver,
                            int n_datapoints,
                            const double* x_data,
                            const double* y_data,
                            const double* errors_or_null);

/* LMsolver::setPar(i_par, val, active, i_dataset = -1, name = ""),
 * lm_solver.h:171-180, lm_solver.cpp:58-113.  i_dataset = -1: global. */
GADFIT_B200_API int gadfit_b200_set_par(gadfit_b200_solver* solver,
                        int i_par,
                        double val,
                        int active,
                        int i_dataset,
                        const char* name_or_null);

/* LMsolver::settings, lm_solver.h:198-210 */
GADFIT_B200_API int gadfit_b200_default_settings(gadfit_b200_settings* out);
GADFIT_B200_API int gadfit_b200_set_settings(gadfit_b200_solver* solver,
                             const gadfit_b200_settings* settings);

/* Text output of LMsolver::fit -- printIterationResults / printTimings and
 * the termination messages, lm_solver.cpp:485,498,587-680.  The library never
 * writes to stdout itself: with a sink installed, fit() hands it the lines the
 * reference would log at info level (one call per line, no newline), selected
 * by settings.verbosity exactly as in the reference (io::flag bits,
 * lm_solver.h:60-72; 0 = results of every iteration).  Numbers are formatted
 * like the reference's ("{}" shortest round-trip, parameter lines with 15
 * significant digits).  Without a sink nothing is formatted or downloaded.
 * NULL removes the sink. */
typedef void (*gadfit_b200_log_fn)(const char* line, void* user);
GADFIT_B200_API int gadfit_b200_set_log(gadfit_b200_solver* solver,
                                        gadfit_b200_log_fn sink,
                                        void* user);

/* The Fortran-only arguments of gadf_fit (see gadfit_b200_gadf_options);
 * default: disabled, every tolerance absent */
GADFIT_B200_API int gadfit_b200_default_gadf_options(gadfit_b200_gadf_options* out);
GADFIT_B200_API int gadfit_b200_set_gadf_options(gadfit_b200_solver* solver,
                                 const gadfit_b200_gadf_options* options);

/* LMsolver::fit(lambda = 10), lm_solver.h:181, lm_solver.cpp:401-511; with
 * gadf options enabled: gadf_fit, fortran/gadfit/gadfit.F90:502-915 */
GADFIT_B200_API int gadfit_b200_fit(gadfit_b200_solver* solver, double lambda);
/* Which criterion ended the last fit() (gadfit_b200_stop_reason) */
GADFIT_B200_API int gadfit_b200_get_stop_reason(gadfit_b200_solver* solver, int* out);

/* getParValue / getValue / chi2 / degreesOfFreedom,
 * lm_solver.h:182-187,195, lm_solver.cpp:513-534,688-696 */
GADFIT_B200_API int gadfit_b200_get_par(gadfit_b200_solver* solver,
                        int i_par,
                        int i_dataset,
                        double* out);
GADFIT_B200_API int gadfit_b200_get_value(gadfit_b200_solver* solver,
                          double arg,
                          int i_dataset,
                          double* out);
GADFIT_B200_API int gadfit_b200_chi2(gadfit_b200_solver* solver, double* out);
GADFIT_B200_API int gadfit_b200_dof(gadfit_b200_solver* solver, int* out);

/* getJacobian/getJTJ/getDTD/getLeftSide/getRightSide/getResiduals/getInvJTJ,
 * lm_solver.h:188-194, lm_solver.cpp:536-585.  Writes min(cap, need) doubles;
 * *need is the full length.  Dense n x n / N x n outputs are materialised on
 * demand from the block-arrow storage and refused with
 * GADFIT_B200_NOT_AVAILABLE beyond 2^27 entries. */
GADFIT_B200_API int gadfit_b200_get_array(gadfit_b200_solver* solver,
                          int which,
                          double* buf,
                          size_t cap,
                          size_t* need);

/* Accept/reject history of the last fit() (see gadfit_b200_trial) */
GADFIT_B200_API int gadfit_b200_get_trace(gadfit_b200_solver* solver,
                          gadfit_b200_trial* buf,
                          size_t cap,
                          size_t* n);
/* Number of successful iterations of the last fit(), lm_solver.cpp:439-501 */
GADFIT_B200_API int gadfit_b200_n_iterations(gadfit_b200_solver* solver, int* out);
GADFIT_B200_API int gadfit_b200_get_timings(gadfit_b200_solver* solver,
                            gadfit_b200_timings* out);
/* Kernels launched by this solver since creation */
GADFIT_B200_API int gadfit_b200_n_launches(gadfit_b200_solver* solver, long long* out);
/* Multi-GPU: whether the small per-trial reductions run as one-shot
 * all-reduces over NVLink peer memory (csrc/peer_reduce.cuh) and how many
 * reductions went that way / through NCCL since creation */
GADFIT_B200_API int gadfit_b200_comm_stats(gadfit_b200_solver* solver,
                                           int* peer_reduce_active,
                                           long long* n_peer_reductions,
                                           long long* n_nccl_reductions);
/* Average device time [s] of the Jacobian-sweep kernel alone in the last
 * fit() (CUDA events on the solver's stream) and the number of launches */
GADFIT_B200_API int gadfit_b200_jacobian_kernel_time(gadfit_b200_solver* solver,
                                     double* seconds_per_launch,
                                     int* launches);

GADFIT_B200_API const char* gadfit_b200_last_error(gadfit_b200_solver* solver);
GADFIT_B200_API const char* gadfit_b200_status_name(int status);

/* ---- device / multi-GPU plumbing (no reference counterpart) ------------- */

/* CUDA device of this solver (default: the current device at create) */
GADFIT_B200_API int gadfit_b200_set_device(gadfit_b200_solver* solver, int device);
/* Copy the borrowed data sets to HBM now instead of at the first fit() */
GADFIT_B200_API int gadfit_b200_upload(gadfit_b200_solver* solver);
/* The reference reads the caller's arrays at every fit() (it holds spans,
 * lm_solver.h:105-113); here they are copied to HBM once.  After changing
 * x / y / errors in place, call this: the next fit() / chi2() / upload()
 * copies them again (the arrays must then be alive again until it has). */
GADFIT_B200_API int gadfit_b200_invalidate_data(gadfit_b200_solver* solver);
/* The five phase timers of the reference (lm_solver.h:140-144) are CUDA event
 * pairs around each phase of every trial step.  They are on by default, as in
 * the reference; in the latency-bound regime (a trial step of ~100 us) the
 * twelve event records per trial are a measurable share of the step, so a
 * caller that does not read gadfit_b200_get_timings / the Jacobian kernel
 * time can switch them off (on != 0: on).  The whole-fit timer ("main")
 * always runs. */
GADFIT_B200_API int gadfit_b200_set_phase_timers(gadfit_b200_solver* solver, int on);

/* One process per GPU.  Rank 0 obtains an id, the caller broadcasts the 128
 * bytes (torch.distributed / MPI / sockets) and every rank attaches.  All
 * ranks then describe the SAME global problem (same add_dataset / set_par
 * calls, as the images of the Fortran coarray version do,
 * fortran/gadfit/gadfit.F90:539-551); each rank uploads and sweeps only the
 * data sets (or the slice of points) it owns, and only the shared-parameter
 * normal block, its right-hand side and chi2 are all-reduced per trial step
 * (the reference's co_sum of the dense JTJ, gadfit.F90:700-701,735,1032). */
GADFIT_B200_API int gadfit_b200_comm_unique_id(void* id_128_bytes);
GADFIT_B200_API int gadfit_b200_set_comm(gadfit_b200_solver* solver,
                         int rank,
                         int world_size,
                         const void* id_128_bytes);
/* Partition used for (n_datasets, points per data set) on (rank, world):
 * first/last+1 owned data set, or -- when there are fewer data sets than
 * ranks -- every data set with the point slice [*point_begin, *point_end)
 * expressed as fractions k/world of each data set (then *split = 1). */
GADFIT_B200_API int gadfit_b200_partition(int n_datasets,
                          const int* n_points,
                          int rank,
                          int world_size,
                          int* first_dataset,
                          int* end_dataset,
                          int* split);

/* FP64 FMA peak of the current device measured by a register-resident DFMA
 * loop on every SM [FLOP/s]; the roofline denominator of bench.py */
GADFIT_B200_API int gadfit_b200_measure_fp64_peak(double* flops_per_second);
/* Diagnostics behind DESIGN.md's kernel choices: which = 0 DFMA vector pipe,
 * 1 DMMA (mma.sync.m8n8k4.f64), 2 both interleaved in one warp [FLOP/s] */
GADFIT_B200_API int gadfit_b200_microbench(int which, double* flops_per_second);

#ifdef __cplusplus
}
#endif

#endif /* GADFIT_B200_H */
