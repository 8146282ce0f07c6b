/* nonlin_batch.h — C ABI of the B200 batched nonlinear-solver engine (libnonlin_b200.so).
 *
 * This is the drop-in boundary for the reference's "M equations / N unknowns" solver path.
 * The reference (jchristopherson/nonlin, Fortran) exposes type-bound procedures that are not
 * C-interoperable; a maintainer binds these entry points with iso_c_binding
 * (fortran/nonlin_batch.f90) next to the same-named Fortran types.  Each
 * entry point cites the reference interface it replaces (paths relative to the reference
 * repository root).
 *
 * Batch layout: B independent systems, structure-of-arrays with the system index fastest:
 *     x[j*B + b]    j = 0..n-1      (a Fortran array declared  x(B, n))
 *     fvec[i*B + b] i = 0..m-1      (                          fvec(B, m))
 *     sys[k*B + b]  k = 0..sys_len-1  per-system data of the residual (the `args` analogue)
 *     shared[i]     data common to all systems (e.g. abscissae)
 * Every data pointer may be a host pointer or a device pointer (detected with
 * cudaPointerGetAttributes); host buffers (pageable or pinned) are copied to and from the handle's
 * grow-only device workspace on the call's stream, and the call then synchronises that stream.
 * There is no CPU fallback: every entry point that computes returns NLB_ERR_NO_DEVICE if no
 * CUDA device is usable.
 *
 * Return value of every function: 0 or an NLB_ERR_* API-level error.  Algorithmic outcomes
 * are per system: status[b] is 0 or the NL_* code the reference would have executed
 * `error stop` with (src/nonlin_error_handling.f90:10-38); x, fvec and ib hold the state at
 * that point, as the reference fills `ib` before stopping (src/nonlin_least_squares.f90:378-390).
 */
#ifndef NONLIN_BATCH_H
#define NONLIN_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- per-system status codes: src/nonlin_error_handling.f90:10-38 ------------------- */
#define NLB_NO_ERROR 0
#define NLB_INVALID_INPUT_ERROR 201
#define NLB_ARRAY_SIZE_ERROR 202
/* NL_CONVERGENCE_ERROR aliases linalg's LA_CONVERGENCE_ERROR (linalg is not vendored in the
 * reference tree; 106 is the value in linalg's published C header). */
#define NLB_CONVERGENCE_ERROR 106
#define NLB_DIVERGENT_BEHAVIOR_ERROR 206
#define NLB_SPURIOUS_CONVERGENCE_ERROR 207
#define NLB_TOLERANCE_TOO_SMALL_ERROR 208
#define NLB_UNDEFINED_FUNCTION_ERROR 211
#define NLB_UNDERDEFINED_PROBLEM_ERROR 212

/* ---- API-level errors (function return values) -------------------------------------- */
#define NLB_OK 0
#define NLB_ERR_INVALID_ARGUMENT 1   /* null pointer, B < 0, bad handle                          */
#define NLB_ERR_UNKNOWN_FCN 2        /* fcn id not registered  (ref: NL_UNDEFINED_FUNCTION_ERROR)  */
#define NLB_ERR_SIZE 3               /* m/n do not match the registered residual, or m < n for LM,
                                        or m != n for Newton / quasi-Newton
                                        (ref: src/nonlin_least_squares.f90:189, src/nonlin_solve.f90:237,519) */
#define NLB_ERR_UNSUPPORTED 4        /* no kernel instantiated for this (solver, fcn) pair         */
#define NLB_ERR_CUDA 5               /* a CUDA runtime call failed: see nlb_last_error()           */
#define NLB_ERR_NO_DEVICE 6          /* no usable CUDA device: the engine has no CPU fallback      */

/* Solver settings = the private members behind the reference's getters/setters. */
typedef struct nlb_params {
    int32_t max_fcn_evals;        /* equation_solver%set_max_fcn_evals      default 100    src/nonlin_multi_eqn_mult_var.f90:69,313 */
    double fcn_tol;               /* equation_solver%set_fcn_tolerance      default 1e-8   :71,334  */
    double var_tol;               /* equation_solver%set_var_tolerance      default 1e-12  :73,354  */
    double grad_tol;              /* equation_solver%set_gradient_tolerance default 1e-12  :75,375  */
    double lm_factor;             /* least_squares_solver%set_step_scaling_factor, default 100, clamped to [0.1, 100]
                                     src/nonlin_least_squares.f90:25,96-115 */
    int32_t jacobian_interval;    /* quasi_newton_solver%set_jacobian_interval  default 5  src/nonlin_solve.f90:51,439 */
    int32_t use_line_search;      /* line_search_solver%set_use_line_search     default 1  src/nonlin_solve.f90:30,144 */
    int32_t ls_max_fcn_evals;     /* line_search%set_max_fcn_evals      default 100   src/nonlin_linesearch.f90:35,82  */
    double ls_alpha;              /* line_search%set_scaling_factor     default 1e-4  src/nonlin_linesearch.f90:38,107 */
    double ls_factor;             /* line_search%set_distance_factor    default 0.1, (0,1) else 0.1 / 0.99
                                     src/nonlin_linesearch.f90:46,133-149 */
    int32_t use_analytic_jacobian;/* vecfcn_helper%set_jacobian was called (registered Jacobian of the residual is
                                     used instead of forward differences)  src/nonlin_multi_eqn_mult_var.f90:143,241 */
    int32_t max_iter_guard;       /* not in the reference: bound on the quasi-Newton iteration counter, which the
                                     reference can spin forever on uphill restarts (src/nonlin_solve.f90:330-337) */
} nlb_params;

/* iteration_behavior, src/nonlin_types.f90:8-29 (gfortran default LOGICAL = 4 bytes). */
typedef struct nlb_iteration_behavior {
    int32_t iter_count;
    int32_t fcn_count;
    int32_t jacobian_count;
    int32_t gradient_count;
    int32_t converge_on_fcn;
    int32_t converge_on_chng;
    int32_t converge_on_zero_diff;
} nlb_iteration_behavior;

/* Batch statistics (the one quantity that is reduced across GPUs). */
enum {
    NLB_STAT_SYSTEMS = 0,          /* systems solved                                   */
    NLB_STAT_CONVERGED = 1,        /* status == 0                                      */
    NLB_STAT_CONVERGED_FCN = 2,    /* converge_on_fcn                                  */
    NLB_STAT_CONVERGED_CHNG = 3,   /* converge_on_chng                                 */
    NLB_STAT_CONVERGED_ZERO_DIFF = 4,
    NLB_STAT_FAILED = 5,           /* status != 0                                      */
    NLB_STAT_SUM_ITER = 6,
    NLB_STAT_SUM_FCN = 7,
    NLB_STAT_SUM_JAC = 8,
    NLB_STAT_MAX_ITER = 9,         /* combine across ranks with MAX, the others with SUM */
    NLB_STAT_COUNT = 16
};

typedef struct nlb_handle nlb_handle;

/* Engine handle: owns a stream and the device staging workspace; one per GPU, thread-safe
 * per handle.  Replaces the per-solve allocate/deallocate of the reference
 * (src/nonlin_least_squares.f90:199-208, src/nonlin_solve.f90:247-254,529-534). */
int nlb_create(nlb_handle** handle, int device);
int nlb_destroy(nlb_handle* handle);
const char* nlb_last_error(const nlb_handle* handle);
/* number of engine kernels launched through this handle since creation */
int64_t nlb_kernel_launch_count(const nlb_handle* handle);

/* Defaults of the reference's solver objects (file:line in nlb_params above). */
void nlb_params_default(nlb_params* p);

/* Residual registry: the analogue of vecfcn_helper%set_fcn(fcn, nfcn, nvar)
 * (src/nonlin_multi_eqn_mult_var.f90:126-140) for compiled-in __device__ residuals.
 * m or n reported as 0 mean "taken from the call"; sys_len / shared_len of -1 mean m. */
int nlb_vecfcn_count(void);
int nlb_vecfcn_lookup(const char* name);            /* id or -1 */
const char* nlb_vecfcn_name(int fcn_id);            /* NULL if unknown */
int nlb_vecfcn_info(int fcn_id, int* m, int* n, int* sys_len, int* shared_len, int* has_jacobian);

/* Residuals compiled outside the engine (plug-ins): vecfcn_helper%set_fcn takes any procedure at run time
 * (src/nonlin_multi_eqn_mult_var.f90:126-140); a __device__ function cannot cross a library boundary, so a new residual
 * is a small library built from nonlin_b200/csrc/nlb_plugin.cuh (see there) that hands the engine host-side launchers
 * of the engine's own kernel templates instantiated for it.  No rebuild of the engine is involved.
 *   nlb_load_plugin      dlopen the library and let it register its residuals; returns how many, or -1
 *   nlb_register_vecfcn  what the plug-in calls for each of them; returns the new id (>= the built-in count) or -1
 * Registered residuals resolve through nlb_vecfcn_lookup / _name / _info like the built-in ones and are served by the
 * least-squares, Newton, quasi-Newton, evaluation and Jacobian entry points (thread-per-system kernels; fixed m, n). */
typedef int (*nlb_user_solve_fn)(int solver, const struct nlb_params* params, int64_t nsys, int64_t B, double* x,
                                 double* fvec, const double* sys, const double* shared,
                                 struct nlb_iteration_behavior* ib, int32_t* status, void* stream);
typedef int (*nlb_user_eval_fn)(int what /* 0 residual, 1 Jacobian */, int analytic, int64_t B, const double* x, double* out,
                                const double* sys, const double* shared, void* stream);
typedef int (*nlb_register_vecfcn_fn)(const char* name, int m, int n, int sys_len, int shared_len, int has_jacobian,
                                      nlb_user_solve_fn solve, nlb_user_eval_fn eval);
int nlb_register_vecfcn(const char* name, int m, int n, int sys_len, int shared_len, int has_jacobian,
                        nlb_user_solve_fn solve, nlb_user_eval_fn eval);
int nlb_load_plugin(const char* path);

/* least_squares_solver%solve  (lss_solve, src/nonlin_least_squares.f90:118-391) over B systems.
 * x in/out, fvec out, ib / status out (fvec, ib and status may each be NULL: an output that is not asked for is not
 * copied back - for host buffers the call is PCIe-bound and x + status are 20 of the 64 bytes per 2x2 system). stream: cudaStream_t, or NULL for the
 * handle's own stream (pass cudaStreamLegacy to name the legacy default stream).  Asynchronous when every pointer is a device pointer; synchronous otherwise. */
int nlb_least_squares_solve_batch(nlb_handle* handle, const nlb_params* params, int fcn_id, int64_t B, int m, int n,
                                  double* x, double* fvec, const double* sys, const double* shared,
                                  nlb_iteration_behavior* ib, int32_t* status, void* stream);

/* newton_solver%solve  (ns_solve, src/nonlin_solve.f90:452-638) over B systems. */
int nlb_newton_solve_batch(nlb_handle* handle, const nlb_params* params, int fcn_id, int64_t B, int m, int n,
                           double* x, double* fvec, const double* sys, const double* shared,
                           nlb_iteration_behavior* ib, int32_t* status, void* stream);

/* quasi_newton_solver%solve  (qns_solve, src/nonlin_solve.f90:156-425) over B systems. */
int nlb_quasi_newton_solve_batch(nlb_handle* handle, const nlb_params* params, int fcn_id, int64_t B, int m, int n,
                                 double* x, double* fvec, const double* sys, const double* shared,
                                 nlb_iteration_behavior* ib, int32_t* status, void* stream);

/* constrained_least_squares_solver: its own settings and the limits of constrained_equation_solver.
 *   trust_region_radius   set_trust_region_radius  (cls_set_radius, src/nonlin_least_squares.f90:898-910; default 1, :64)
 *   step_scaling_factor   set_step_scaling_factor  (cls_set_factor, :923-935; default 1, :66)
 *   lower / upper         set_lower_limits / set_upper_limits (:812-855): n doubles each in HOST memory, shared by
 *                         every system of the batch as they are shared by every solve of one solver object;
 *                         NULL = the -huge / +huge arrays cls_solve installs when none were set (:1014-1024).
 * As in the setters, a non-positive radius or factor selects 1. */
typedef struct nlb_constrained_options {
    double trust_region_radius;
    double step_scaling_factor;
    const double* lower;
    const double* upper;
} nlb_constrained_options;
void nlb_constrained_options_default(nlb_constrained_options* o);

/* constrained_least_squares_solver%solve  (cls_solve, src/nonlin_least_squares.f90:938-1176; dogleg :1301-1403,
 * coleman_li_scaling :1222-1260, alpha_box :1181-1219, ces_apply_limits :858-883) over B systems.  Arguments as
 * nlb_least_squares_solve_batch.  Available for the built-in residuals with n <= 16: the fixed-size ones (m, n
 * reported non-zero by nlb_vecfcn_info) and the curve-fit families with a run-time number of observations m (their
 * m-sized state lives in a device workspace capped at 4 GB; a batch x m that cannot be served within it returns
 * NLB_ERR_UNSUPPORTED, as plug-in residuals do).  status[b] = 0 or NLB_CONVERGENCE_ERROR (:1173-1175);
 * a system whose start (after clamping to the limits) or first residual is NaN or +-huge returns status 0 with an
 * all-zero iteration_behavior, as the reference's early `return` does (:1043-1045). */
int nlb_constrained_least_squares_solve_batch(nlb_handle* handle, const nlb_params* params,
                                              const nlb_constrained_options* options, int fcn_id, int64_t B, int m,
                                              int n, double* x, double* fvec, const double* sys, const double* shared,
                                              nlb_iteration_behavior* ib, int32_t* status, void* stream);

/* polynomial%fit  (poly_fit, src/nonlin_polynomials.f90:146-199) and, with thru_zero != 0, polynomial%fit_thru_zero
 * (poly_fit_thru_zero, :202-253) over B data sets of npts points each: the least-squares polynomial of the given
 * order through (x, y), solved as the reference does (Vandermonde matrix, linalg solve_least_squares = LAPACK DGELS).
 *   x       abscissae: npts doubles shared by every data set (x_is_shared != 0) or x[i*B + b]
 *   y       ordinates y[i*B + b]; NOT overwritten (the reference's y is intent(inout) scratch)
 *   coeffs  out, coeffs[k*B + b] = c_k of data set b, k = 0..order (c0 first, as polynomial%get(k+1));
 *           c0 = 0 for thru_zero
 *   status  out (may be NULL): 0, or NLB_LA_INVALID_OPERATION_ERROR where the matrix is exactly rank deficient
 *           (linalg reports that instead of returning coefficients)
 * `order >= npts` or `order < 1` is the reference's `error stop 4` (:166-169): NLB_ERR_SIZE.  At most 8 fitted
 * coefficients (order <= 7, or <= 8 through zero); more returns NLB_ERR_UNSUPPORTED. */
#define NLB_LA_INVALID_OPERATION_ERROR 107
int nlb_polynomial_fit_batch(nlb_handle* handle, int64_t B, int npts, int order, int thru_zero, int x_is_shared,
                             const double* x, const double* y, double* coeffs, int32_t* status, void* stream);

/* polynomial%evaluate  (poly_eval_double, src/nonlin_polynomials.f90:256-283) for B polynomials at npts points:
 * y[i*B + b] = c_order x^order + ... + c_0, Horner from the highest coefficient.  order 0..8. */
int nlb_polynomial_evaluate_batch(nlb_handle* handle, int64_t B, int order, int npts, int x_is_shared,
                                  const double* coeffs, const double* x, double* y, void* stream);

/* One-variable solvers.  Settings = the members of equation_solver_1var
 * (src/nonlin_single_var.f90:44-54: set_max_fcn_evals :227, set_fcn_tolerance :248, set_var_tolerance :268,
 * set_diff_tolerance :311) plus "fcn1var_helper%set_diff was called" (:203-211). */
typedef struct nlb_params_1var {
    int32_t max_fcn_evals;      /* 100   */
    double fcn_tol;             /* 1e-8  */
    double var_tol;             /* 1e-12 */
    double diff_tol;            /* 1e-12 */
    int32_t use_analytic_diff;  /* 0: forward difference of f1h_diff_fcn (:154-200) */
} nlb_params_1var;
void nlb_params_1var_default(nlb_params_1var* p);

/* Registry of compiled-in one-variable functions: the analogue of fcn1var_helper%set_fcn (:132-140). */
int nlb_fcn1var_count(void);
int nlb_fcn1var_lookup(const char* name);            /* id or -1 */
const char* nlb_fcn1var_name(int fcn_id);            /* NULL if unknown */
int nlb_fcn1var_info(int fcn_id, int* args_len, int* has_derivative);

/* brent_solver%solve  (brent_solve, src/nonlin_solve.f90:643-835) and newton_1var_solver%solve  (newt1var_solve,
 * src/nonlin_solve.f90:840-1032) over B independent equations f(x; args_b) = 0.
 *   lim1, lim2   the value_pair of each equation (search limits, either order), B doubles each
 *   x            in/out, B doubles: the root.  Brent ignores the input and returns 0 where it fails, as the
 *                reference does; Newton leaves x untouched where the limits are rejected
 *   f            out, B doubles, or NULL = the reference's optional `f` absent (Newton then counts one evaluation less)
 *   args         per-equation data args[k*B + b] (`class(*) args`), NULL if the function takes none
 *   status       0, NLB_INVALID_INPUT_ERROR (|lim1 - lim2| < epsilon, :713 / :899) or NLB_CONVERGENCE_ERROR (:813, :1004)
 * ib%jacobian_count carries Newton's derivative-evaluation count, as in the reference (:1022). */
int nlb_brent_solve_batch(nlb_handle* handle, const nlb_params_1var* params, int fcn_id, int64_t B, const double* lim1,
                          const double* lim2, double* x, double* f, const double* args, nlb_iteration_behavior* ib,
                          int32_t* status, void* stream);
int nlb_newton_1var_solve_batch(nlb_handle* handle, const nlb_params_1var* params, int fcn_id, int64_t B,
                                const double* lim1, const double* lim2, double* x, double* f, const double* args,
                                nlb_iteration_behavior* ib, int32_t* status, void* stream);

/* vecfcn_helper%fcn  (vfh_fcn, src/nonlin_multi_eqn_mult_var.f90:178-195) over B points. */
int nlb_vecfcn_eval_batch(nlb_handle* handle, int fcn_id, int64_t B, int m, int n, const double* x, double* fvec,
                          const double* sys, const double* shared, void* stream);

/* vecfcn_helper%jacobian  (vfh_jac_fcn, src/nonlin_multi_eqn_mult_var.f90:198-277) over B points:
 * forward differences, or the registered Jacobian when params->use_analytic_jacobian.
 * jac[(i + j*m)*B + b] = d f_i / d x_j of system b (column-major m x n per system). */
int nlb_jacobian_batch(nlb_handle* handle, const nlb_params* params, int fcn_id, int64_t B, int m, int n,
                       const double* x, double* jac, const double* sys, const double* shared, void* stream);

/* Convergence statistics of a finished batch: stats[NLB_STAT_COUNT] (host or device pointer).
 * Multi-GPU runs sum these across ranks (MAX for NLB_STAT_MAX_ITER) — the only collective of
 * the path. */
int nlb_reduce_stats(nlb_handle* handle, int64_t B, const nlb_iteration_behavior* ib, const int32_t* status,
                     int64_t* stats, void* stream);

/* One batch over several GPUs of ONE process (the caller shape of a Fortran host: a single `solve` call, reference
 * src/nonlin_multi_eqn_mult_var.f90:94-119; no launcher, no MPI).  handles[ndev]: one handle per device.  The batch
 * shards by contiguous system ranges (sizes differ by at most one); every device solves its range on its own host
 * thread - there is no data-path collective - and the convergence statistics are combined with one NCCL all-reduce
 * (SUM; MAX for NLB_STAT_MAX_ITER) over NVLink into stats[NLB_STAT_COUNT] (host; may be NULL: no collective at all).
 * All data pointers are HOST buffers laid out as for the single-device calls; fvec, ib and status may be NULL.
 * solver: one of NLB_SOLVER_*.  NCCL is loaded at run time (libnccl.so.2) and only when ndev > 1 and stats != NULL. */
#define NLB_SOLVER_LEAST_SQUARES 0
#define NLB_SOLVER_NEWTON 1
#define NLB_SOLVER_QUASI_NEWTON 2
int nlb_solve_sharded(nlb_handle* const* handles, int ndev, int solver, const nlb_params* params, int fcn_id, int64_t B,
                      int m, int n, double* x, double* fvec, const double* sys, const double* shared,
                      nlb_iteration_behavior* ib, int32_t* status, int64_t* stats);

/* Measured FP64 throughput of this GPU (roofline denominator): dependent-chain-free DFMA and
 * DADD/DMUL micro-kernels, TFLOP/s.  The parity build issues no DFMA, so its ceiling is the
 * second number. */
int nlb_measure_fp64_peak(nlb_handle* handle, double* dfma_tflops, double* dadd_dmul_tflops);

/* Dependent-chain latencies on this GPU, in SM cycles per operation: cycles4[0] = DADD, [1] = IEEE division,
 * [2] = sqrt (+ one add), [3] = shared-memory load + DADD.  These bound the serial chains (ordered sums, Givens
 * chains, triangular solves) that the reference's summation order imposes. */
int nlb_measure_fp64_latency(nlb_handle* handle, double* cycles4);

#ifdef __cplusplus
}
#endif
#endif /* NONLIN_BATCH_H */
