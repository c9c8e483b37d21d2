// Launch-argument block shared by the host library (sbm_api.cu) and the NVRTC-compiled integrator
// kernels (kernels/sbm_ode_kernels.cuh).  Plain C, fixed-width via (unsigned) long long only so
// that NVRTC needs no system headers.  Layout must stay identical on both sides.
#ifndef SBM_KERNEL_ARGS_H
#define SBM_KERNEL_ARGS_H

// per-trajectory status codes written to status[b*E + e]
#define SBM_OK            0
#define SBM_ERR_MAXSTEPS  1   /* step budget exhausted before t_end            */
#define SBM_ERR_UNDERFLOW 2   /* step size collapsed below resolution of t     */
#define SBM_ERR_NONFINITE 3   /* NaN/Inf in the state that step halving cannot cure */

// global counters (unsigned long long each)
#define SBM_CNT_STEPS   0     /* attempted steps                               */
#define SBM_CNT_REJECT  1     /* rejected steps                                */
#define SBM_CNT_RHS     2     /* augmented right-hand-side evaluations         */
#define SBM_CNT_LU      3     /* LU factorisations (Rosenbrock)                */
#define SBM_CNT_TRAJ    4     /* trajectories finished                         */
#define SBM_CNT_QUEUE   7     /* work-queue cursor of the persistent kernels (zeroed by sbm_solve) */
#define SBM_N_COUNTERS  8

typedef struct SbmSolveArgs {
    // ---- flat plan (device pointers; built once per project by sbm_plan_create) ----
    const int*                pmap;      // [E][NP]  project column feeding model parameter m, or -1 (fixed)
    const double*             pfix;      // [E][NP]  value used when pmap < 0
    const double*             y0;        // [E][N]   initial state (reference: zeros)
    const int*                tout_off;  // [E+1]    offsets into tout
    const double*             tout;      // snapped output times, strictly increasing per experiment
    const int*                orow_off;  // [n_tout+1] offsets into the row tables, per output time
    const int*                row_sim;   // [n_orow] residual-row index receiving the observable
    const int*                row_jac;   // [n_orow] Jacobian-row index (differs only in reference-compat mode)
    const unsigned long long* row_mask;  // [n_orow] observable = sum of states whose bit is set
    // ---- batch of parameter sets ----
    const double*             theta;     // log-parameters; element (b, c) at theta[b*theta_sb + c*theta_sp]
    long long                 theta_sb;
    long long                 theta_sp;
    double*                   sims;      // [B][n_rows]   unscaled observables at the snapped times
    double*                   jac;       // [B][n_rows][n_proj] d(observable)/d(theta), or NULL
    int*                      status;    // [B][E]
    unsigned long long*       counters;  // [SBM_N_COUNTERS]
    long long                 B;
    int                       E;
    int                       n_rows;
    int                       n_proj;
    int                       max_steps;
    int                       jac_atomic; // 1: accumulate with atomicAdd (duplicate rows/columns possible)
    int                       jac_raw;    // 1: jac holds raw dy/dp_free (column j = free parameter j), no chain rule
    double                    rtol;
    double                    atol;
    double                    h0;         // initial step; 0 = automatic
    int                       n_tout;     // lengths of tout / of the row tables (small plans are staged in shared memory)
    int                       n_orow;
    int                       dense;      // 1: states-only Dormand-Prince reads output times off the continuous extension
} SbmSolveArgs;

#endif
