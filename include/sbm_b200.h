/* sbm_b200.h - C ABI of libsbm_b200.so, the B200-native batched ODE-with-sensitivities engine behind
 * SysBio_Modeling's OdeModel / Project hot path.
 *
 * The reference (FedericoV/SysBio_Modeling) is pure Python and has no FFI; its boundary for this path
 * is two duck-typed Python interfaces.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference checkout).  INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; sbm_last_error() gives a thread-local message;
 *   - device pointers are plain pointers to memory owned by the caller (e.g. torch.Tensor.data_ptr());
 *     the library never frees caller memory; opaque handles are library-owned (sbm_*_destroy);
 *   - all launches go on the caller's stream (a cudaStream_t passed as void*); the device-pointer
 *     entry points never synchronise; the *_host entry points copy in/out and synchronise the stream;
 *   - all arithmetic is IEEE double.  There is NO CPU fallback: without a CUDA device every compute
 *     entry point fails with SBM_E_CUDA.
 */
#ifndef SBM_B200_H
#define SBM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SBM_E_INVALID  (-1)   /* bad argument                        */
#define SBM_E_NVRTC    (-2)   /* model/kernels failed to compile     */
#define SBM_E_CUDA     (-3)   /* CUDA driver/runtime error, no GPU   */
#define SBM_E_NOMEM    (-4)

/* solvers */
#define SBM_SOLVER_DP5    0   /* adaptive Dormand-Prince 5(4), non-stiff                           */
#define SBM_SOLVER_RODAS4 1   /* Rosenbrock RODAS4 with per-trajectory dense LU, stiff; at most 12 states.  With a Jacobian
                                 buffer: forward sensitivities through the same LU (autonomous models only) */

/* loss functions (reference project/loss_functions/squared_loss/) */
#define SBM_LOSS_SQUARE     0 /* squared_loss_function.py:27-80                                     */
#define SBM_LOSS_LOG        1 /* log_squared_loss_function.py:23-98 + log_scale_factor.py           */
#define SBM_LOSS_NORMALIZED 2 /* normalized_squared_loss_function.py:23-46 (sigma <- sigma*mean)    */

/* flags for sbm_loss / sbm_jacobian */
#define SBM_FLAG_PRIOR_JAC_UNIT 1 /* parameter-prior Jacobian rows = unit vectors (intended maths);
                                     default 0 reproduces the reference, whose rows stay zero because
                                     base_project.py:519-530 is never called by :764-771            */

#define SBM_FLAG_LOSS_GENERIC    2 /* force the CTA-per-set loss kernel (J stays in global memory) even when the
                                     warp-per-set kernel with the TMA-staged tile fits (tests compare the two)        */

/* row kinds in sbm_plan_desc_t.row_kind */
#define SBM_ROW_MEASUREMENT 0
#define SBM_ROW_PARAM_PRIOR 1 /* row_aux = project column (base_project.py:427-437)                 */
#define SBM_ROW_SF_PRIOR    2 /* row_aux = scale-factor group (abstract_loss_function.py:93-120)    */
#define SBM_ROW_PLAIN       3 /* residual (sim - data)/sigma on the value as given: no scale factor, no log */

typedef struct sbm_model sbm_model;
typedef struct sbm_plan  sbm_plan;

/* ---- model: generated RHS / Jacobian / sensitivity device functions + integrator kernels ------------
 * Replaces the numba-compiled callables `model(y,t,yout,p)` / `sens_model(y,t,yout,p)` that the reference
 * builds in model/ode_model.py:46-57 from symbolic/sympy_tools.py:162-269.  `model_src` is the text
 * produced by sysbio_modeling_b200.symbolic.emit_cuda; it is compiled with NVRTC for `arch` together with
 * the integrator kernel templates embedded in the library.                                            */
typedef struct sbm_model_desc_t {
    int         n_vars;          /* N                                                  */
    int         n_params;        /* all model parameters, fixed ones included           */
    int         n_free;          /* sensitivity columns (non-fixed parameters)          */
    int         group_lanes;     /* lanes per trajectory for the sensitivity kernel, 0 = auto */
    int         cols_per_lane;   /* columns per lane, 0 = auto                          */
    int         block_threads;   /* CTA size of the sensitivity kernel, 0 = auto        */
    int         sens_layout;     /* sensitivity kernel: <=0 auto, 1 lane-per-column warp groups (sbm_dp5_sens),
                                    2 CTA-transposed: lane = trajectory, warp = column group (sbm_dp5_ct)  */
    const char* arch;            /* NULL -> "sm_100a"                                   */
    const char* cache_dir;       /* NULL -> no on-disk cubin cache                      */
} sbm_model_desc_t;

typedef struct sbm_model_info_t {
    int n_vars, n_params, n_free;
    int group_lanes, cols_per_lane, block_threads, stage_smem, dyn_smem_bytes;
    int regs_dp5_plain, regs_dp5_sens, regs_rodas4;          /* registers per thread (after module load; -1 before) */
    int local_dp5_plain, local_dp5_sens, local_rodas4;       /* local-memory (spill) bytes per thread               */
    int smem_dp5_plain, smem_dp5_sens, smem_rodas4;          /* static shared memory bytes per CTA                  */
    int cubin_bytes;
    int from_cache;
    int sens_layout;                                         /* layout sbm_solve uses for sensitivity solves (1 | 2) */
    int ct_warps, ct_cols, ct_cols_state, ct_dyn_smem_bytes, ct_ctas_per_sm; /* CTA-transposed kernel: warps per CTA,
                                                                columns per warp / on the last warp */
    int regs_dp5_ct, local_dp5_ct, smem_dp5_ct;
} sbm_model_info_t;

int         sbm_model_compile(const char* model_src, const sbm_model_desc_t* desc, sbm_model** out);
int         sbm_model_info(sbm_model* m, sbm_model_info_t* out);   /* loads the module if a device is present */
const char* sbm_model_log(const sbm_model* m);                       /* NVRTC / ptxas -v log                   */
void        sbm_model_destroy(sbm_model* m);

/* ---- plan: the flat numeric form of a Project ---------------------------------------------------------
 * Replaces the per-call Python bookkeeping of project/base_project.py:343-391 (parameter gather,
 * linspace/searchsorted time snapping, row offsets) and :443-488 (column scatter): built once per
 * project on the host, copied to the device here.  All arrays are HOST pointers and are copied.       */
typedef struct sbm_plan_desc_t {
    int E;                 /* experiments (sorted by name, base_project.py:624)                         */
    int n_rows;            /* residual rows incl. prior rows                                           */
    int n_proj;            /* project parameters (columns)                                             */
    int n_groups;          /* scale-factor groups                                                      */
    int n_tout;            /* total snapped output times over all experiments                          */
    int n_orow;            /* total (output time, measurement) pairs                                   */
    int jac_atomic;        /* 1 if Jacobian entries can collide (duplicate columns / reference row mode) */
    const int*                pmap;      /* [E][n_params]                                              */
    const double*             pfix;      /* [E][n_params]                                              */
    const double*             y0;        /* [E][n_vars]                                                */
    const int*                tout_off;  /* [E+1]                                                      */
    const double*             tout;      /* [n_tout]                                                   */
    const int*                orow_off;  /* [n_tout+1]                                                 */
    const int*                row_sim;   /* [n_orow]                                                   */
    const int*                row_jac;   /* [n_orow]                                                   */
    const unsigned long long* row_mask;  /* [n_orow]                                                   */
    const double*             data;      /* [n_rows] measurement mean / prior centre                   */
    const double*             sigma;     /* [n_rows] measurement sigma / prior sigma                   */
    const int*                row_group; /* [n_rows] scale-factor group or -1                          */
    const int*                row_kind;  /* [n_rows] SBM_ROW_*                                         */
    const int*                row_aux;   /* [n_rows]                                                   */
    const int*                grp_off;   /* [n_groups+1]                                               */
    const int*                grp_rows;  /* rows of each group                                         */
} sbm_plan_desc_t;

/* m may be NULL: a loss-only plan (row tables for sbm_loss; E and the integrator tables are ignored). */
int  sbm_plan_create(sbm_model* m, const sbm_plan_desc_t* desc, sbm_plan** out);
void sbm_plan_destroy(sbm_plan* p);

typedef struct sbm_solver_opts_t {
    int    solver;        /* SBM_SOLVER_*                                                              */
    int    max_steps;     /* per trajectory; <=0 -> 500000                                             */
    int    jac_raw;       /* 1: jac = raw dy/dp_free, [B][n_rows][n_free], no log-parameter chain rule
                             (the OdeModel.calc_jacobian convention, model/ode_model.py:83-126)        */
    int    dense_output;  /* states-only kernels (Dormand-Prince, RODAS4): 1 = output times are read off the continuous
                             extension of the step that passes them (no step is clipped except onto the last time), 2 = every
                             step lands on its output time, 0 = automatic: dense when a trajectory has >= 256 output times (the
                             reference's 1000-node grids, base_project.py:419-420) */
    double rtol, atol;    /* reference hard-codes 1e-10 / 1e-10 (model/ode_model.py:122-123,167-168)   */
    double h0;            /* initial step, 0 = automatic                                               */
} sbm_solver_opts_t;

/* ---- solve: OdeModel.simulate / OdeModel.calc_jacobian for every (parameter set, experiment) ----------
 * Replaces model/ode_model.py:128-169 (plain) and :83-126 (sensitivity-augmented) plus the mapping
 * functions project/utils.py:10-89 and the chain rule/scatter of base_project.py:443-488.
 *   theta   device, element (b, c) at theta[b*theta_sb + c*theta_sp]          (log-parameters)
 *   sims    device [B][n_rows]            observables at the snapped times (prior rows untouched)
 *   jac     device [B][n_rows][n_proj]    d sims / d theta, or NULL for a plain solve; zeroed here
 *           (with opts->jac_raw the plan's n_proj must equal the model's n_free)
 *   status  device int [B][E]             0 ok, else SBM_ERR_* of sbm_kernel_args.h (rows then hold NaN)
 *   counters device u64 [8]               accumulated (steps, rejects, rhs evals, LU, trajectories)   */
int sbm_solve(sbm_plan* p, const sbm_solver_opts_t* opts, long long B,
              const double* theta, long long theta_sb, long long theta_sp,
              double* sims, double* jac, int* status, unsigned long long* counters, void* stream);

/* ---- loss epilogue: scale factors, residuals, scaled Jacobian, SSR, gradient ---------------------------
 * Replaces project/loss_functions/abstract_loss_function.py:46-120, squared_loss/linear_scale_factor.py:27-53,
 * squared_loss/squared_loss_function.py:27-106 and the reductions of base_project.py:773-827.
 *   sims    device [B][n_rows] in: observables from sbm_solve; out: prior rows filled (theta / log B)
 *   jac     device [B][n_rows][n_proj] in: model Jacobian; out: scaled Jacobian B*J + sim*dB (NULL: skip)
 *   resid   device [B][n_rows]   (scaled_sim - data)/sigma          ssr device [B] = 0.5*sum resid^2
 *   sf      device [B][n_groups] optimal scale factors               sfgrad device [B][n_groups][n_proj] (workspace/out; NULL iff jac NULL)
 *   grad    device [B][n_proj] = J_scaled^T resid (NULL: skip)                                        */
int sbm_loss(sbm_plan* p, int loss_kind, int flags, long long B,
             const double* theta, long long theta_sb, long long theta_sp,
             double* sims, double* jac, double* resid, double* ssr, double* sf, double* sfgrad,
             double* grad, void* stream);

/* ---- the loss epilogue in two halves: the hook of the EXPERIMENT shard (experiments of one parameter set spread over
 * several GPUs).  A scale factor couples the rows of its group across all experiments (abstract_loss_function.py:62-91),
 * so every rank forms the group sums of ITS rows, the sums are all-reduced, and every rank finishes its rows with the
 * global sums.  sums is device [B][n_groups][3 + 2 n_proj]:
 *   linear_scale_factor.py:27-42   p1 = sum s d / sigma^2, p2 = sum s^2 / sigma^2, (p3 unused),
 *                                  j1[c] = sum J_c d / sigma^2, j2[c] = sum J_c s / sigma^2
 *   log_scale_factor.py:17-36      p1 = sum 1/sr^2, p2 = sum log(d)/sr^2, p3 = sum log(s)/sr^2, j1[c] = sum J_c/(s sr^2)
 * sbm_loss == sbm_loss_sums + sbm_loss_finish without the exchange.  jac may be NULL (residual path: p1..p3 only).   */
int sbm_loss_sums(sbm_plan* p, int loss_kind, int flags, long long B,
                  const double* theta, long long theta_sb, long long theta_sp,
                  double* sims, double* jac, double* sums, void* stream);
int sbm_loss_finish(sbm_plan* p, int loss_kind, int flags, long long B,
                    const double* theta, long long theta_sb, long long theta_sp,
                    double* sims, double* jac, const double* sums, double* resid, double* ssr, double* sf,
                    double* sfgrad, double* grad, void* stream);

/* ---- per-set normal equations: what the Jacobian-block gather of the multi-GPU path ships ("normally
 * only J^T J / J^T r [22 x 22 + 22] per set are needed") and what scipy.optimize.leastsq's consumer forms from
 * Project.calc_project_jacobian / residuals (tests/test_Project.py:202-213).
 *   out device [B][n_proj (n_proj + 1) / 2 + n_proj + 1]: lower triangle of J^T J (row-major), J^T r, r^T r.  n_proj <= 32. */
int sbm_normal_equations(long long B, int n_rows, int n_proj, const double* jac, const double* resid, double* out,
                         void* stream);

/* ---- fused objective calls with HOST buffers (what a reference-side binding calls) ---------------------
 * Project.residuals / calc_sum_square_residuals (base_project.py:708-729, 807-827) and
 * Project.calc_project_jacobian / calc_rss_gradient (base_project.py:731-805) for B parameter sets.
 * theta_host is [B][n_proj] row-major.  Any output pointer may be NULL.  Copies are made on `stream`
 * and the stream is synchronised before returning.                                                     */
int sbm_residuals_host(sbm_plan* p, const sbm_solver_opts_t* opts, int loss_kind, long long B,
                       const double* theta_host, double* resid_host, double* ssr_host, double* sf_host,
                       double* sims_host, int* status_host, unsigned long long* counters_host, void* stream);
int sbm_jacobian_host(sbm_plan* p, const sbm_solver_opts_t* opts, int loss_kind, int flags, long long B,
                      const double* theta_host, double* jac_host, double* model_jac_host /* unscaled, before the loss */,
                      double* resid_host, double* ssr_host,
                      double* grad_host, double* sf_host, double* sims_host, int* status_host,
                      unsigned long long* counters_host, void* stream);

/* ---- pipelined form of sbm_jacobian_host: the copy-out of one evaluation (C4: 494 MB of scaled Jacobians, ~9 ms over PCIe)
 * cannot hide inside its own call (the persistent integrator fills every SM), but it hides behind the NEXT evaluation.
 * sbm_jacobian_host_begin enqueues copy-in, integrator and loss epilogue on the caller's stream and the copy-out on the
 * plan's own copy stream, using workspace slot 0 or 1, and returns without waiting; sbm_host_end(slot) waits until that
 * slot's host buffers are complete.  A consumer of Project.calc_project_jacobian that evaluates points back to back
 * (multistart LM, an ensemble sampler) keeps two evaluations in flight.  Host buffers should be pinned.                  */
int sbm_jacobian_host_begin(sbm_plan* p, const sbm_solver_opts_t* opts, int loss_kind, int flags, long long B,
                            const double* theta_host, double* jac_host, double* resid_host, double* ssr_host,
                            double* grad_host, double* sf_host, double* sims_host, int* status_host,
                            unsigned long long* counters_host, void* stream, int slot);
int sbm_host_end(sbm_plan* p, int slot);

/* ---- batched Levenberg-Marquardt step (the consumer of residuals + Jacobian) ---------------------------
 * Replaces the inner linear algebra of scipy.optimize.leastsq(proj.residuals, x0, Dfun=proj.calc_project_jacobian)
 * (reference call sites tests/test_Project.py:202-213, 351-357) for B parameter sets at once, on the device:
 *   (J^T J + lambda D) delta = -J^T r,  pred = 0.5 delta^T (lambda D delta - J^T r),
 *   D_j = max(dscale_j, (J^T J)_jj) written back to dscale (MINPACK's running column scaling; 1 where still zero)
 *   jac    device [B][n_rows][n_proj]  (as written by sbm_loss)      resid  device [B][n_rows]
 *   lambda device [B]                   dscale device [B][n_proj] in/out, zero before the first step (NULL: no memory)
 *   delta  device [B][n_proj]           pred, gmax device [B]   (gmax = max |J^T r|)
 *   info   device int [B]: 0 ok, 1 system not positive definite or input not finite (delta = 0: raise lambda)
 * n_proj <= 32.                                                                                            */
int sbm_lm_step(long long B, int n_rows, int n_proj, const double* jac, const double* resid, const double* lambda,
                double* dscale, double* delta, double* pred, double* gmax, int* info, void* stream);

/* ---- batched Metropolis ensemble sampler: the two steps around the objective evaluation ------------------------------------
 * W lock-step copies of the reference's ensemble_log_params (Ensembles.py:55-178).  Per Monte-Carlo step: sbm_mc_propose,
 * sbm_solve + sbm_loss on the candidates, sbm_mc_accept - four launches, no host synchronisation.
 *   sbm_mc_propose   nxt[w] = cur[w] + samp . z  (Ensembles.py:266-270 with the sampling matrix of :226-264);
 *                    z = normals[W][P] when given, else Philox4x32-10(seed; walker, index, step) + Box-Muller
 *   sbm_mc_accept    F_new = ssr_new - T x scale-factor entropy (base_project.py:854-892; linear_scale_factor.py:13-18, 63-80
 *                    by a composite 16-point Gauss-Legendre rule; only groups with a log prior: sf_mu / sf_sig [n_groups],
 *                    NaN = none, both NULL = no entropy term); accept if u < exp(-(F_new - curF) / T) (Ensembles.py:193-198);
 *                    accepted walkers: cur <- nxt, curF <- F_new, accepted += 1 (the reference never updates curr_F,
 *                    :151-154 - not reproduced).  init = 1: no move, curF <- F(cur) from ssr_new / sims_new.
 *                    ens_out [W][P] / fes_out [W]: where to record this step's positions / free energies (NULL: not recorded).
 * All pointers are device pointers; u = uniforms[W] when given, else the device generator.                                   */
int sbm_mc_propose(long long W, int P, const double* samp, const double* cur, double* nxt, const double* normals,
                   unsigned long long seed, unsigned step, void* stream);
int sbm_mc_accept(sbm_plan* p, long long W, int init, double temperature, double* cur, const double* nxt, double* curF,
                  const double* ssr_new, const double* sims_new, const double* sf_mu, const double* sf_sig,
                  const double* uniforms, unsigned long long seed, unsigned step, double* accepted,
                  double* ens_out, double* fes_out, void* stream);

/* ---- utilities ----------------------------------------------------------------------------------------- */
/* DFMA-throughput microbenchmark on the current device: the measured FP64 roofline denominator. */
int         sbm_fp64_peak(int iters, double* tflops_out, double* ms_out);
int         sbm_device_count(void);
const char* sbm_last_error(void);
int         sbm_version(void);

#ifdef __cplusplus
}
#endif
#endif
