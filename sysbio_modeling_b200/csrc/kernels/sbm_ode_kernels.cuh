// =================================================================================================
// sbm_ode_kernels.cuh - batched FP64 ODE(+forward sensitivity) integrators for NVIDIA B200 (sm_100a)
//
// This text is compiled at run time by NVRTC, AFTER the model-specific device functions emitted by
// sysbio_modeling_b200.symbolic.emit_cuda (sbm_f, sbm_fJG, sbm_col, sbm_jac_dense, sbm_dfdt and the
// SBM_N / SBM_NP / SBM_NF / SBM_NJ / SBM_NG macros) and sbm_kernel_args.h.  It replaces, for every
// (experiment, parameter set) pair at once, what the reference does one experiment at a time:
//   OdeModel.simulate / OdeModel.calc_jacobian   (reference model/ode_model.py:83-169, SciPy LSODA)
//   the mapping functions                         (reference project/utils.py:10-89)
//   the log-parameter chain rule + column scatter (reference project/base_project.py:343-363, 443-488)
//
// Work decomposition ("group" = G lanes of one warp, G in {1,2,4,8,16,32}):
//   * the augmented system  [y | S_1 ... S_NF]  (S_j = dy/dp_j, N doubles each) is split by COLUMN:
//     lane l of the group owns columns l, l+G, ... (C slots); column 0 is the state y itself.
//     Column dynamics are independent given y:  dS_j/dt = J(y) S_j + df/dp_j, so a stage needs one
//     broadcast of the stage value of y (shuffles from the group leader) and nothing else.
//   * the state x of a lane's columns lives in registers, the six Runge-Kutta stage vectors in shared
//     memory (thread-minor, conflict-free); nothing but the final observable rows ever goes to HBM -
//     full trajectories never exist.
//   * all lanes of a group share t, h and the accept/reject decision (error norm by xor-shuffle
//     butterfly), so there is no divergence inside a group; groups in one warp may diverge.
//   * output: the integrator steps exactly onto each snapped measurement time (no dense output) and
//     forms the observable (sum over a state bit-mask) and its log-parameter derivative row in
//     place; lanes write consecutive project columns -> coalesced row segments of the Jacobian.
// =================================================================================================

#ifndef SBM_BLOCK_SENS
#define SBM_BLOCK_SENS 128
#endif
#ifndef SBM_BLOCK_PLAIN
#define SBM_BLOCK_PLAIN 128
#endif

#define SBM_NPP (SBM_NP | 1)      /* odd stride: conflict-free shared-memory parameter rows */

#ifndef SBM_HAS_FREE_IDX
#error "model source must define sbm_free_idx"
#endif

// ---------------------------------------------------------------------------------------------
// Dormand-Prince 5(4) coefficients
// ---------------------------------------------------------------------------------------------
#define DP_C2 (1.0/5.0)
#define DP_C3 (3.0/10.0)
#define DP_C4 (4.0/5.0)
#define DP_C5 (8.0/9.0)
#define DP_A21 (1.0/5.0)
#define DP_A31 (3.0/40.0)
#define DP_A32 (9.0/40.0)
#define DP_A41 (44.0/45.0)
#define DP_A42 (-56.0/15.0)
#define DP_A43 (32.0/9.0)
#define DP_A51 (19372.0/6561.0)
#define DP_A52 (-25360.0/2187.0)
#define DP_A53 (64448.0/6561.0)
#define DP_A54 (-212.0/729.0)
#define DP_A61 (9017.0/3168.0)
#define DP_A62 (-355.0/33.0)
#define DP_A63 (46732.0/5247.0)
#define DP_A64 (49.0/176.0)
#define DP_A65 (-5103.0/18656.0)
#define DP_A71 (35.0/384.0)
#define DP_A73 (500.0/1113.0)
#define DP_A74 (125.0/192.0)
#define DP_A75 (-2187.0/6784.0)
#define DP_A76 (11.0/84.0)
#define DP_E1 (71.0/57600.0)
#define DP_E3 (-71.0/16695.0)
#define DP_E4 (71.0/1920.0)
#define DP_E5 (-17253.0/339200.0)
#define DP_E6 (22.0/525.0)
#define DP_E7 (-1.0/40.0)
// continuous extension of order 4 (Hairer, Norsett, Wanner II.6; dopri5.f CONTD5)
#define DP_D1 (-12715105075.0/11282082432.0)
#define DP_D3 (87487479700.0/32700410799.0)
#define DP_D4 (-10690763975.0/1880347072.0)
#define DP_D5 (701980252875.0/199316789632.0)
#define DP_D6 (-1453857185.0/822651844.0)
#define DP_D7 (69997945.0/29380423.0)

// ---------------------------------------------------------------------------------------------
// group helpers
// ---------------------------------------------------------------------------------------------
// xor-butterfly over the G lanes of a group; the whole warp takes part (warp-uniform control flow), so the
// mask is the constant full mask and offsets < G keep the exchange inside each group
template <int G>
__device__ __forceinline__ double sbm_group_sum(double v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Observable rows for one output time: value (state column) and d/dtheta row (sensitivity columns).
template <int G, int C, int SENS>
__device__ __forceinline__ void sbm_emit_rows(const SbmSolveArgs& a, const long long b, const int kt,
                                              const double (&x)[C][SBM_N], const int (&cid)[C],
                                              const int (&pcol)[C], const double (&pscale)[C],
                                              const bool poison)
{
    const int r0 = a.orow_off[kt], r1 = a.orow_off[kt + 1];
    for (int r = r0; r < r1; ++r) {
        const unsigned long long mask = a.row_mask[r];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < SBM_N; ++i) acc += ((mask >> i) & 1ull) ? x[c][i] : 0.0;
            if (poison) acc = __longlong_as_double(0x7ff8000000000000ll);
            if (cid[c] == 0) {
                a.sims[b * a.n_rows + a.row_sim[r]] = acc;
            } else if (SENS && a.jac != nullptr && pcol[c] >= 0) {
                // chain rule to log-parameters: d/dtheta_c = d/dp * exp(theta_c) = d/dp * p
                double* dst = a.jac + ((b * a.n_rows + a.row_jac[r]) * (long long)a.n_proj + pcol[c]);
                const double v = acc * pscale[c];
                if (a.jac_atomic) atomicAdd(dst, v); else *dst = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stage-vector storage: the six Runge-Kutta k-vectors of this lane's columns live in dynamic shared
// memory, element-major / thread-minor (stride = CTA size), so a warp's accesses are consecutive
// doubles - conflict-free LDS.64/STS.64.  Only x (and one transient column of the stage input) stay in
// registers, which is what lets a lane own C > 1 columns of a large system: the model's primal part
// (f, J, G) is then evaluated once per lane for C columns and once per warp instruction for 32/G
// trajectories.  Slots are addressed at run time, which lets ONE copy of the stage code serve all
// stages (the model code is large; seven inlined copies thrash the instruction cache - measured).
// ---------------------------------------------------------------------------------------------
template <int C, int BLOCK>
struct SbmStages {
    double* s;
    __device__ __forceinline__ double get(const int v, const int c, const int i) const
    {
        return s[((v * C + c) * SBM_N + i) * BLOCK];
    }
    __device__ __forceinline__ void set(const int v, const int c, const int i, const double val)
    {
        s[((v * C + c) * SBM_N + i) * BLOCK] = val;
    }
};

// Butcher tableau by stage id.  Stage ids 0 and 1 are the start-up evaluations (k1 = F(x0) and the
// explicit-Euler probe of Hairer's initial-step rule); 2..6 are the Dormand-Prince stages and 7 is the
// 5th-order solution whose derivative is both the FSAL k1 of the next step and the 7th error stage.
__constant__ double sbm_dp_a[8][6] = {
    {0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {1.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {DP_A21, 0.0, 0.0, 0.0, 0.0, 0.0},
    {DP_A31, DP_A32, 0.0, 0.0, 0.0, 0.0},
    {DP_A41, DP_A42, DP_A43, 0.0, 0.0, 0.0},
    {DP_A51, DP_A52, DP_A53, DP_A54, 0.0, 0.0},
    {DP_A61, DP_A62, DP_A63, DP_A64, DP_A65, 0.0},
    {DP_A71, 0.0, DP_A73, DP_A74, DP_A75, DP_A76}};
__constant__ double sbm_dp_c[8] = {0.0, 1.0, DP_C2, DP_C3, DP_C4, DP_C5, 1.0, 1.0};
__constant__ int sbm_dp_nt[8] = {0, 1, 1, 2, 3, 4, 5, 6};      // number of k-vectors entering the stage input
__constant__ int sbm_dp_out[8] = {0, 1, 1, 2, 3, 4, 5, 1};     // logical k-slot receiving the stage derivative

// ---------------------------------------------------------------------------------------------
// Adaptive Dormand-Prince 5(4), FSAL, PI step controller (Hairer's DOPRI5 constants)
//
// Control flow is WARP-UNIFORM: every lane of a warp runs the same flat loop - one Runge-Kutta stage
// per iteration - until all groups of the warp are done; a finished (or out-of-range) group keeps
// executing with zero step and its writes predicated off.  That costs nothing under SIMT (the warp
// would wait for its slowest group anyway), lets every shuffle use the constant full mask, and keeps
// a single copy of the generated model code in the loop body.
// ---------------------------------------------------------------------------------------------
template <int G, int C, int SENS, int BLOCK, int DENSE = 0>
__device__ __forceinline__ void sbm_dp5_impl(const SbmSolveArgs& a)
{
    constexpr int N = SBM_N;
    constexpr int NCOL = SENS ? (SBM_NF + 1) : 1;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(G * C >= NCOL, "group too small for the augmented system");
    constexpr int GPB = BLOCK / G;                 // groups (trajectories) per block
    __shared__ double sp[GPB * SBM_NPP];
    extern __shared__ double sbm_dyn[];            // [6][C][N][BLOCK] stage vectors

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int lig = (G == 1) ? 0 : (lane & (G - 1));
    const int leader = lane - lig;
    const long long ntraj = a.B * (long long)a.E;
    const long long traj_raw = ((long long)blockIdx.x * BLOCK + tid) / G;
    const bool active = traj_raw < ntraj;
    const long long traj = active ? traj_raw : 0;  // idle groups shadow trajectory 0 and write nothing
    const int e = (int)(traj % a.E);
    const long long b = traj / a.E;

    // ---- model parameters: p_m = exp(theta[col]) or the experiment's fixed value (base_project.py:343-363)
    double* p = sp + (tid / G) * SBM_NPP;
    for (int m = lig; m < SBM_NP; m += G) {
        const int c = a.pmap[e * SBM_NP + m];
        p[m] = (c >= 0) ? exp(a.theta[b * a.theta_sb + c * a.theta_sp]) : a.pfix[e * SBM_NP + m];
    }
    __syncwarp();

    int cid[C], pcol[C];
    double pscale[C];
    double x[C][N];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        cid[c] = lig + G * c;
        pcol[c] = -1;
        pscale[c] = 0.0;
        if (SENS && cid[c] >= 1 && cid[c] <= SBM_NF) {
            const int m = sbm_free_idx[cid[c] - 1];
            if (a.jac_raw) { pcol[c] = cid[c] - 1; pscale[c] = 1.0; }
            else { pcol[c] = a.pmap[e * SBM_NP + m]; pscale[c] = p[m]; }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[c][i] = (cid[c] == 0) ? a.y0[e * N + i] : 0.0;
    }

    const double rtol = a.rtol, atol = a.atol;
    const double inv_naug = 1.0 / (double)(N * NCOL);
    int kt = a.tout_off[e];
    const int kt1 = a.tout_off[e + 1];
    const double tend = (kt1 > kt) ? a.tout[kt1 - 1] : 1.0;
    bool done = !active || kt >= kt1;
    double tk = done ? 0.0 : a.tout[kt];

    SbmStages<C, BLOCK> K;
    K.s = sbm_dyn + tid;
    // slots 0/1 swap roles on every accepted step (FSAL: k7 becomes the next k1 without a copy)
    int s0 = 0, s1 = 1;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) { K.set(0, c, i, 0.0); K.set(1, c, i, 0.0); }

    unsigned nstep = 0, nrej = 0, nrhs = 0;
    int status = SBM_OK;
    double t = 0.0, h = 0.0, hprop = 0.0, h0 = 0.0, d1 = 0.0;
    float facold = 1e-4f;
    bool last_rejected = false, clipped = false, go = false;
    int stage = 0;                                  // warp-uniform

    while (__any_sync(FULL, !done)) {
        if (stage == 2) {
            // ---- step start: arrived at an output time (or failed: flush the remaining rows as NaN) ----
            if (!done && (t >= tk || status != SBM_OK)) {
                sbm_emit_rows<G, C, SENS>(a, b, kt, x, cid, pcol, pscale, status != SBM_OK);
                kt++;
                done = kt >= kt1;
                if (!done) tk = a.tout[kt];
            }
            const bool stepping = !done && status == SBM_OK && t < tk;
            // clip onto the output time; remember the unclipped proposal
            hprop = h;
            clipped = false;
            if (stepping) {
                // dense output (states only): steps are clipped onto the LAST output time only; the others are read off the
                // continuous extension of the accepted step that passes them
                const double tclip = DENSE ? tend : tk;
                if (t + 1.01 * h >= tclip) { h = tclip - t; clipped = true; }
                if (nstep >= (unsigned)a.max_steps) status = SBM_ERR_MAXSTEPS;
                else if (!clipped && !(h > 1e-15 * fmax(t, 1e-300))) status = SBM_ERR_UNDERFLOW;
            }
            go = stepping && status == SBM_OK;
        }
        // idle groups take a harmless zero-length step
        const double hs = (stage == 0) ? 0.0 : (stage == 1) ? h0 : (go ? h : 0.0);

        // ================= one stage: input, model, derivative, (error) =================
        const int nt = sbm_dp_nt[stage];
        const int vout_l = sbm_dp_out[stage];
        const int vout = (vout_l == 0) ? s0 : (vout_l == 1) ? s1 : vout_l;
        const double ts = t + sbm_dp_c[stage] * hs;
        double cf[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) cf[j] = hs * sbm_dp_a[stage][j];

        // stage input: loads are unconditional (every slot always holds finite-or-garbage data that is
        // simply not selected), so the common path of a stage has no branch at all
        double w[N], Y[N], out[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = x[0][i];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const double kv = K.get((j == 0) ? s0 : (j == 1) ? s1 : j, 0, i);
                const double nx = fma(cf[j], kv, acc);
                acc = (j < nt) ? nx : acc;
            }
            w[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) Y[i] = (G == 1) ? w[i] : __shfl_sync(FULL, w[i], leader);

        if (!SENS) {
            sbm_f(ts, Y, p, out);
#pragma unroll
            for (int i = 0; i < N; ++i) K.set(vout, 0, i, out[i]);
        } else {
            double f[N], Jv[SBM_NJ], Gv[SBM_NG];
            sbm_fJG(ts, Y, p, f, Jv, Gv);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (c > 0) {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        double acc = x[c][i];
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const double kv = K.get((j == 0) ? s0 : (j == 1) ? s1 : j, c, i);
                            const double nx = fma(cf[j], kv, acc);
                            acc = (j < nt) ? nx : acc;
                        }
                        w[i] = acc;
                    }
                }
                sbm_col(f, Jv, Gv, w, cid[c], out);
#pragma unroll
                for (int i = 0; i < N; ++i) K.set(vout, c, i, out[i]);
            }
        }

        // ================= what the stage was for =================
        if (stage == 0) {
            // k1 = F(t0, x0) is in slot s0.  Initial step size (Hairer, Norsett, Wanner II.4), part 1.
            double d0 = 0.0;
            d1 = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double isk = sbm_rcp(atol + rtol * fabs(x[c][i]));
                    const double kv = K.get(s0, c, i);
                    d0 = fma(x[c][i] * isk, x[c][i] * isk, d0);
                    d1 = fma(kv * isk, kv * isk, d1);
                }
            d0 = sqrt(sbm_group_sum<G>(d0) * inv_naug);
            d1 = sqrt(sbm_group_sum<G>(d1) * inv_naug);
            h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            h0 = fmin(h0, tend);
            nrhs++;
            stage = 1;
        } else if (stage == 1) {
            // slot s1 holds F(x0 + h0*k1): second-derivative estimate, part 2 of the initial step rule
            double d2 = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double q = (K.get(s1, c, i) - K.get(s0, c, i)) * sbm_rcp(atol + rtol * fabs(x[c][i]));
                    d2 = fma(q, q, d2);
                }
            d2 = sqrt(sbm_group_sum<G>(d2) * inv_naug) / h0;
            const double dm = fmax(d1, d2);
            const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.2);
            h = fmin(fmin(100.0 * h0, h1), tend);
            if (a.h0 > 0.0) h = fmin(a.h0, tend);
            nrhs++;
            stage = 2;
        } else if (stage < 7) {
            stage++;
        } else {
            // ---- end of step: 5th-order solution and embedded error estimate in one pass over the stage
            //      vectors (k7 sits in slot s1), scaled RMS over the whole augmented state ----
            double xn[C][N];
            double sq = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double k1v = K.get(s0, c, i), k3v = K.get(2, c, i), k4v = K.get(3, c, i);
                    const double k5v = K.get(4, c, i), k6v = K.get(5, c, i), k7v = K.get(s1, c, i);
                    xn[c][i] = fma(cf[5], k6v, fma(cf[4], k5v, fma(cf[3], k4v, fma(cf[2], k3v, fma(cf[0], k1v, x[c][i])))));
                    const double er = hs * fma(DP_E7, k7v, fma(DP_E6, k6v, fma(DP_E5, k5v, fma(DP_E4, k4v,
                                      fma(DP_E3, k3v, DP_E1 * k1v)))));
                    const double q = er * sbm_rcp(atol + rtol * fmax(fabs(x[c][i]), fabs(xn[c][i])));
                    sq = fma(q, q, sq);
                }
            const double err = sqrt(sbm_group_sum<G>(sq) * inv_naug);
            if (go) {
                nstep++;
                nrhs += 6;
                if (err <= 1.0) {
                    // accept; PI controller (DOPRI5: beta = 0.04, safe = 0.9, fac in [0.2, 10])
                    const float ferr = fmaxf((float)err, 1e-30f);
                    float fac = __powf(ferr, 0.2f - 0.04f * 0.75f) / __powf(facold, 0.04f);
                    fac = fmaxf(0.1f, fminf(5.0f, fac / 0.9f));
                    facold = fmaxf(ferr, 1e-4f);
                    if (DENSE) {
                        const double tn = clipped ? tend : t + h;
                        while (kt < kt1 - 1 && tk <= tn) {
                            const double th = (tk - t) / h, th1 = 1.0 - th;
                            double xi[C][N];
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                const double k1v = K.get(s0, 0, i), k3v = K.get(2, 0, i), k4v = K.get(3, 0, i);
                                const double k5v = K.get(4, 0, i), k6v = K.get(5, 0, i), k7v = K.get(s1, 0, i);
                                const double yd = xn[0][i] - x[0][i];
                                const double bs = fma(h, k1v, -yd);
                                const double r4 = yd - h * k7v - bs;
                                const double r5 = h * fma(DP_D7, k7v, fma(DP_D6, k6v, fma(DP_D5, k5v, fma(DP_D4, k4v,
                                                  fma(DP_D3, k3v, DP_D1 * k1v)))));
                                xi[0][i] = fma(th, fma(th1, fma(th, fma(th1, r5, r4), bs), yd), x[0][i]);
#pragma unroll
                                for (int c = 1; c < C; ++c) xi[c][i] = 0.0;
                            }
                            sbm_emit_rows<G, C, SENS>(a, b, kt, xi, cid, pcol, pscale, false);
                            kt++;
                            tk = a.tout[kt];
                        }
                    }
                    t = clipped ? (DENSE ? tend : tk) : t + h;
#pragma unroll
                    for (int c = 0; c < C; ++c)
#pragma unroll
                        for (int i = 0; i < N; ++i) x[c][i] = xn[c][i];
                    // FSAL: k7 (slot s1) is the next step's k1
                    const int tmp = s0; s0 = s1; s1 = tmp;
                    double hnew = h / (double)fac;
                    if (last_rejected) hnew = fmin(hnew, h);
                    if (clipped) hnew = fmax(hnew, hprop);   // a clipped step must not shrink the cruise step
                    h = hnew;
                    last_rejected = false;
                } else {
                    nrej++;
                    last_rejected = true;
                    if (err == err && err < 1e300) {
                        const float ferr = (float)fmin(err, 1e30);
                        h = h / (double)fminf(5.0f, __powf(ferr, 0.2f - 0.04f * 0.75f) / 0.9f);
                    } else {
                        h *= 0.1;                            // NaN/Inf in the trial state: retreat hard
                        if (h < 1e-300) status = SBM_ERR_NONFINITE;
                    }
                }
            }
            stage = 2;
        }
    }

    if (active && lig == 0) {
        a.status[traj] = status;
        atomicAdd(a.counters + SBM_CNT_STEPS, (unsigned long long)nstep);
        atomicAdd(a.counters + SBM_CNT_REJECT, (unsigned long long)nrej);
        atomicAdd(a.counters + SBM_CNT_RHS, (unsigned long long)nrhs);
        atomicAdd(a.counters + SBM_CNT_TRAJ, 1ull);
    }
}

// ---------------------------------------------------------------------------------------------
// Rosenbrock RODAS4 (Hairer & Wanner, stiff; L-stable, order 4(3)), one thread per trajectory,
// dense LU with partial pivoting held in registers.  States only (no sensitivity columns).
// ---------------------------------------------------------------------------------------------
#define RO_GAMMA_V 0.25
#define RO_C2_V 0.386
#define RO_C3_V 0.21
#define RO_C4_V 0.63
#define RO_D1_V 0.25
#define RO_D2_V (-0.1043)
#define RO_D3_V 0.1035
#define RO_D4_V (-0.03620000000000023)
#define RO_A21_V 1.544
#define RO_A31_V 0.9466785280815826
#define RO_A32_V 0.2557011698983284
#define RO_A41_V 3.314825187068521
#define RO_A42_V 2.896124015972201
#define RO_A43_V 0.9986419139977817
#define RO_A51_V 1.221224509226641
#define RO_A52_V 6.019134481288629
#define RO_A53_V 12.53708332932087
#define RO_A54_V (-0.6878860361058950)
#define RO_C21_V (-5.6688)
#define RO_C31_V (-2.430093356833875)
#define RO_C32_V (-0.2063599157091915)
#define RO_C41_V (-0.1073529058151375)
#define RO_C42_V (-9.594562251023355)
#define RO_C43_V (-20.47028614809616)
#define RO_C51_V 7.496443313967647
#define RO_C52_V (-10.24680431464352)
#define RO_C53_V (-33.99990352819905)
#define RO_C54_V 11.70890893206160
#define RO_C61_V 8.083246795921522
#define RO_C62_V (-7.981132988064893)
#define RO_C63_V (-31.52159432874371)
#define RO_C64_V 16.31930543123136
#define RO_C65_V (-6.058818238834054)
// dense output (rodas.f, RODAS4: D21..D25, D31..D35)
#define RO_D21_V 10.12623508344586
#define RO_D22_V (-7.487995877610167)
#define RO_D23_V (-34.80091861555747)
#define RO_D24_V (-7.992771707568823)
#define RO_D25_V 1.025137723295662
#define RO_D31_V (-0.6762803392801253)
#define RO_D32_V 6.087714651680015
#define RO_D33_V 16.43084320892478
#define RO_D34_V 24.76722511418386
#define RO_D35_V (-6.594389125716872)
// the Rosenbrock coefficients as FMA operands straight from the constant bank (literal doubles are materialised with two
// moves each: UMOV + IMAD were 20 % of the executed instructions of sbm_rodas4_plain, profiles/r02_rodas4_plain_final.txt)
__constant__ double sbm_roc[43] = {RO_GAMMA_V, RO_C2_V, RO_C3_V, RO_C4_V, RO_D1_V, RO_D2_V, RO_D3_V, RO_D4_V, RO_A21_V, RO_A31_V, RO_A32_V, RO_A41_V, RO_A42_V, RO_A43_V, RO_A51_V, RO_A52_V, RO_A53_V, RO_A54_V, RO_C21_V, RO_C31_V, RO_C32_V, RO_C41_V, RO_C42_V, RO_C43_V, RO_C51_V, RO_C52_V, RO_C53_V, RO_C54_V, RO_C61_V, RO_C62_V, RO_C63_V, RO_C64_V, RO_C65_V, RO_D21_V, RO_D22_V, RO_D23_V, RO_D24_V, RO_D25_V, RO_D31_V, RO_D32_V, RO_D33_V, RO_D34_V, RO_D35_V};
#define RO_GAMMA sbm_roc[0]
#define RO_C2 sbm_roc[1]
#define RO_C3 sbm_roc[2]
#define RO_C4 sbm_roc[3]
#define RO_D1 sbm_roc[4]
#define RO_D2 sbm_roc[5]
#define RO_D3 sbm_roc[6]
#define RO_D4 sbm_roc[7]
#define RO_A21 sbm_roc[8]
#define RO_A31 sbm_roc[9]
#define RO_A32 sbm_roc[10]
#define RO_A41 sbm_roc[11]
#define RO_A42 sbm_roc[12]
#define RO_A43 sbm_roc[13]
#define RO_A51 sbm_roc[14]
#define RO_A52 sbm_roc[15]
#define RO_A53 sbm_roc[16]
#define RO_A54 sbm_roc[17]
#define RO_C21 sbm_roc[18]
#define RO_C31 sbm_roc[19]
#define RO_C32 sbm_roc[20]
#define RO_C41 sbm_roc[21]
#define RO_C42 sbm_roc[22]
#define RO_C43 sbm_roc[23]
#define RO_C51 sbm_roc[24]
#define RO_C52 sbm_roc[25]
#define RO_C53 sbm_roc[26]
#define RO_C54 sbm_roc[27]
#define RO_C61 sbm_roc[28]
#define RO_C62 sbm_roc[29]
#define RO_C63 sbm_roc[30]
#define RO_C64 sbm_roc[31]
#define RO_C65 sbm_roc[32]
#define RO_D21 sbm_roc[33]
#define RO_D22 sbm_roc[34]
#define RO_D23 sbm_roc[35]
#define RO_D24 sbm_roc[36]
#define RO_D25 sbm_roc[37]
#define RO_D31 sbm_roc[38]
#define RO_D32 sbm_roc[39]
#define RO_D33 sbm_roc[40]
#define RO_D34 sbm_roc[41]
#define RO_D35 sbm_roc[42]

// ---------------------------------------------------------------------------------------------
// States-only Dormand-Prince 5(4) in PUSH form, one thread per trajectory (round 2; C5: 25 states, 131,072 walkers).
//
// sbm_dp5_impl keeps the six stage derivatives in shared memory and forms every stage input from all of them: per stage
// 6 N loads + 6 N FMAs + 12 N selects around a model of ~6 N flops, and 6 N doubles of shared memory per thread (C5: 38 KB
// per warp, one warp per scheduler).  Here every stage derivative is pushed into the accumulators of the later stages as
// soon as it exists (the tableau of sbm_dp5_ct.cuh): four vectors V1..V4 in shared memory, the state in registers -
//   stage 2: w = y + h a21 V1 (V1 = k1)            V2 <- k2
//   stage 3: w = y + h (a31 V1 + a32 V2)            V3 <- k3
//   stage 4: w = y + h (a41 V1 + a42 V2 + a43 V3)   V1 <- input of stage 5, V2 / V3 <- partial inputs of stages 6 / 7,
//                                                   V4 <- partial error estimate (all from k1..k4)
//   stage 5: w = V1    V2, V3, V4 += (a65, a75, e5) h k5
//   stage 6: w = V2    V3, V4 += (a76, e6) h k6
//   stage 7: w = V3 = the 5th-order solution;  V1 <- k7 (FSAL),  V4 += e7 h k7 = the error estimate
// 17 vector loads and 14 vector stores per step instead of 36 + 6, no selects, 4 N doubles of shared memory per thread.
// After a rejected step k1 is evaluated again (V1 holds k7 of the rejected trial).  Output times: every step lands on
// its output time (the dense variant needs all stage derivatives: sbm_dp5_plain_dense).
// ---------------------------------------------------------------------------------------------
template <int BLOCK>
__device__ __forceinline__ void sbm_dp5_push_impl(const SbmSolveArgs& a)
{
    constexpr int N = SBM_N;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double sp[BLOCK * SBM_NPP];
    extern __shared__ double sbm_dyn[];            // [4 or 5][N][BLOCK]: V1..V4 (and the state y when it is not in registers)
    const int tid = threadIdx.x, lane = tid & 31;
    const long long ntraj = a.B * (long long)a.E;
    unsigned long long* queue = a.counters + SBM_CNT_QUEUE;
    double* p = sp + tid * SBM_NPP;
    double* V = sbm_dyn + tid;
    // keeps ptxas from hoisting all 3 N loads of an update above its FMA chains (the live ranges cost 1.7 KB of spills)
#ifndef SBM_PUSH_FENCE_EVERY
#define SBM_PUSH_FENCE_EVERY 2
#endif
#define SBM_PUSH_FENCE(i) do { if (SBM_PUSH_FENCE_EVERY > 0 && ((i) % SBM_PUSH_FENCE_EVERY) == SBM_PUSH_FENCE_EVERY - 1) asm volatile("" ::: "memory"); } while (0)
#define SBM_PV(v, i) V[((v) * N + (i)) * BLOCK]
    const int cid[1] = {0};
    const int pcol[1] = {-1};
    const double pscale[1] = {0.0};
    // the state in registers while three state-sized arrays fit the register file, in shared memory beyond
#ifndef SBM_PUSH_YREG_MAX
#define SBM_PUSH_YREG_MAX 25
#endif
#if SBM_N <= SBM_PUSH_YREG_MAX
    double yreg[N];
#define SBM_PY(i) yreg[i]
#else
#define SBM_PY(i) SBM_PV(4, i)
#endif
    double w[N], out[N];
    const double rtol = a.rtol, atol = a.atol;

    // A persistent grid: a lane that finishes its trajectory takes the next one from a global queue (one warp-aggregated
    // atomic), so the lanes of a warp stay busy although their step counts differ (a one-trajectory-per-thread launch ran at
    // 26.6 of 32 lanes on C5, profiles/r02_dp5_push.txt).  One loop: [fetch + start-up of the lanes that need it] + [one step
    // attempt of every lane that has a trajectory, its output rows when the step lands on an output time].
    long long traj = -1, b = 0;
    int e = 0, kt = 0, kt1 = 0, status = SBM_OK;
    double t = 0.0, h = 0.0, tk = 0.0;
    float facold = 1e-4f;
    bool last_rejected = false, k1_stale = false, exhausted = false;
    unsigned nstep = 0;                            // of the trajectory in flight (step budget)
    unsigned long long c_step = 0, c_rej = 0, c_rhs = 0, c_traj = 0;

    for (;;) {
        // ---- fetch ----
        const bool need = (traj < 0) && !exhausted;
        const unsigned want = __ballot_sync(FULL, need);
        if (want) {
            const int leader = __ffs(want) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(queue, (unsigned long long)__popc(want));
            base = __shfl_sync(FULL, base, leader);
            if (need) {
                const long long nt = (long long)base + __popc(want & ((1u << lane) - 1u));
                if (nt >= ntraj) {
                    exhausted = true;
                } else {
                    traj = nt;
                    e = (int)(traj % a.E);
                    b = traj / a.E;
                    for (int m = 0; m < SBM_NP; ++m) {
                        const int c = a.pmap[e * SBM_NP + m];
                        p[m] = (c >= 0) ? exp(a.theta[b * a.theta_sb + c * a.theta_sp]) : a.pfix[e * SBM_NP + m];
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) { SBM_PY(i) = a.y0[e * N + i]; w[i] = a.y0[e * N + i]; }
                    kt = a.tout_off[e];
                    kt1 = a.tout_off[e + 1];
                    status = SBM_OK;
                    t = 0.0;
                    nstep = 0;
                    facold = 1e-4f;
                    last_rejected = false;
                    k1_stale = false;
                    if (kt >= kt1) {                          // no output times: nothing to integrate
                        a.status[traj] = SBM_OK;
                        c_traj++;
                        traj = -1;
                    } else {
                        tk = a.tout[kt];
                        const double tend = a.tout[kt1 - 1];
                        // ---- k1 and the initial step size (Hairer, Norsett, Wanner II.4) ----
                        sbm_f(t, w, p, out);
                        double d0 = 0.0, d1 = 0.0;
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            SBM_PV(0, i) = out[i];
                            const double isk = sbm_rcp(atol + rtol * fabs(w[i]));
                            d0 = fma(w[i] * isk, w[i] * isk, d0);
                            d1 = fma(out[i] * isk, out[i] * isk, d1);
                        }
                        d0 = sqrt(d0 * (1.0 / N)); d1 = sqrt(d1 * (1.0 / N));
                        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
                        h0 = fmin(h0, tend);
#pragma unroll
                        for (int i = 0; i < N; ++i) w[i] = fma(h0, out[i], w[i]);
                        sbm_f(t + h0, w, p, out);
                        double d2 = 0.0;
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            const double q = (out[i] - SBM_PV(0, i)) * sbm_rcp(atol + rtol * fabs(SBM_PY(i)));
                            d2 = fma(q, q, d2);
                        }
                        d2 = sqrt(d2 * (1.0 / N)) / h0;
                        const double dm = fmax(d1, d2);
                        const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.2);
                        h = fmin(fmin(100.0 * h0, h1), tend);
                        if (a.h0 > 0.0) h = fmin(a.h0, tend);
                        c_rhs += 2;
                    }
                }
            }
        }
        if (__all_sync(FULL, traj < 0 && exhausted)) break;
        if (traj < 0) continue;

        // ---- one step attempt ----
        const double hprop = h;
        bool clipped = false;
        if (t + 1.01 * h >= tk) { h = tk - t; clipped = true; }
        if (nstep >= (unsigned)a.max_steps) status = SBM_ERR_MAXSTEPS;
        else if (!clipped && !(h > 1e-15 * fmax(t, 1e-300))) status = SBM_ERR_UNDERFLOW;
        bool landed = false;
        if (status == SBM_OK) {
            nstep++;
            c_step++;
            if (k1_stale) {                                   // after a rejection: V1 holds k7 of the rejected trial
#pragma unroll
                for (int i = 0; i < N; ++i) w[i] = SBM_PY(i);
                sbm_f(t, w, p, out);
#pragma unroll
                for (int i = 0; i < N; ++i) SBM_PV(0, i) = out[i];
                c_rhs++;
                k1_stale = false;
            }
            double sq = 0.0;
#pragma unroll 1
            for (int S = 2; S <= 7; ++S) {
                // ---- stage input ----
                if (S == 2) {
#pragma unroll
                    for (int i = 0; i < N; ++i) w[i] = fma(h * DP_A21, SBM_PV(0, i), SBM_PY(i));
                } else if (S == 3) {
#pragma unroll
                    for (int i = 0; i < N; ++i) w[i] = fma(h, fma(DP_A32, SBM_PV(1, i), DP_A31 * SBM_PV(0, i)), SBM_PY(i));
                } else if (S == 4) {
#pragma unroll
                    for (int i = 0; i < N; ++i)
                        w[i] = fma(h, fma(DP_A43, SBM_PV(2, i), fma(DP_A42, SBM_PV(1, i), DP_A41 * SBM_PV(0, i))), SBM_PY(i));
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) w[i] = SBM_PV(S - 5, i);
                }
                sbm_f(fma(sbm_dp_c[S], h, t), w, p, out);
                // ---- what the stage does with its derivative ----
                if (S == 2) {
#pragma unroll
                    for (int i = 0; i < N; ++i) SBM_PV(1, i) = out[i];
                } else if (S == 3) {
#pragma unroll
                    for (int i = 0; i < N; ++i) SBM_PV(2, i) = out[i];
                } else if (S == 4) {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double k1v = SBM_PV(0, i), k2v = SBM_PV(1, i), k3v = SBM_PV(2, i), k4v = out[i];
                        SBM_PV(0, i) = fma(h, fma(DP_A54, k4v, fma(DP_A53, k3v, fma(DP_A52, k2v, DP_A51 * k1v))), SBM_PY(i));
                        SBM_PV(1, i) = fma(h, fma(DP_A64, k4v, fma(DP_A63, k3v, fma(DP_A62, k2v, DP_A61 * k1v))), SBM_PY(i));
                        SBM_PV(2, i) = fma(h, fma(DP_A74, k4v, fma(DP_A73, k3v, DP_A71 * k1v)), SBM_PY(i));
                        SBM_PV(3, i) = h * fma(DP_E4, k4v, fma(DP_E3, k3v, DP_E1 * k1v));
                        SBM_PUSH_FENCE(i);
                    }
                } else if (S == 5) {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double hk = h * out[i];
                        SBM_PV(1, i) = fma(DP_A65, hk, SBM_PV(1, i));
                        SBM_PV(2, i) = fma(DP_A75, hk, SBM_PV(2, i));
                        SBM_PV(3, i) = fma(DP_E5, hk, SBM_PV(3, i));
                        SBM_PUSH_FENCE(i);
                    }
                } else if (S == 6) {
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double hk = h * out[i];
                        SBM_PV(2, i) = fma(DP_A76, hk, SBM_PV(2, i));
                        SBM_PV(3, i) = fma(DP_E6, hk, SBM_PV(3, i));
                        SBM_PUSH_FENCE(i);
                    }
                } else {
                    // w = the 5th-order solution, out = k7
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        SBM_PV(0, i) = out[i];
                        const double er = fma(h * DP_E7, out[i], SBM_PV(3, i));
                        const double q = er * sbm_rcp(atol + rtol * fmax(fabs(SBM_PY(i)), fabs(w[i])));
                        sq = fma(q, q, sq);
                    }
                }
            }
            c_rhs += 6;
            const double err = sqrt(sq * (1.0 / N));
            if (err <= 1.0) {
                // accept; PI controller (DOPRI5: beta = 0.04, safe = 0.9, fac in [0.2, 10])
                const float ferr = fmaxf((float)err, 1e-30f);
                float fac = __powf(ferr, 0.2f - 0.04f * 0.75f) / __powf(facold, 0.04f);
                fac = fmaxf(0.1f, fminf(5.0f, fac / 0.9f));
                facold = fmaxf(ferr, 1e-4f);
                t = clipped ? tk : t + h;
                landed = clipped;
#pragma unroll
                for (int i = 0; i < N; ++i) SBM_PY(i) = w[i];
                double hnew = h / (double)fac;
                if (last_rejected) hnew = fmin(hnew, h);
                if (clipped) hnew = fmax(hnew, hprop);       // a clipped step must not shrink the cruise step
                h = hnew;
                last_rejected = false;
            } else {
                c_rej++;
                last_rejected = true;
                k1_stale = true;
                if (err == err && err < 1e300) {
                    const float ferr = (float)fmin(err, 1e30);
                    h = h / (double)fminf(5.0f, __powf(ferr, 0.2f - 0.04f * 0.75f) / 0.9f);
                } else {
                    h *= 0.1;                                // NaN/Inf in the trial state: retreat hard
                    if (h < 1e-300) status = SBM_ERR_NONFINITE;
                }
            }
        }
        // ---- output rows: the step landed on an output time, or the trajectory failed (the remaining rows become NaN) ----
        if (landed || status != SBM_OK) {
            double y[1][N];
#pragma unroll
            for (int i = 0; i < N; ++i) y[0][i] = SBM_PY(i);
            do {
                sbm_emit_rows<1, 1, 0>(a, b, kt, y, cid, pcol, pscale, status != SBM_OK);
                kt++;
            } while (status != SBM_OK && kt < kt1);
            if (kt >= kt1) {
                a.status[traj] = status;
                c_traj++;
                traj = -1;
            } else {
                tk = a.tout[kt];
            }
        }
    }
#undef SBM_PV
#undef SBM_PY
    // ---- counters: one atomic per warp ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c_step += __shfl_xor_sync(FULL, c_step, o);
        c_rej += __shfl_xor_sync(FULL, c_rej, o);
        c_rhs += __shfl_xor_sync(FULL, c_rhs, o);
        c_traj += __shfl_xor_sync(FULL, c_traj, o);
    }
    if (lane == 0) {
        atomicAdd(a.counters + SBM_CNT_STEPS, c_step);
        atomicAdd(a.counters + SBM_CNT_REJECT, c_rej);
        atomicAdd(a.counters + SBM_CNT_RHS, c_rhs);
        atomicAdd(a.counters + SBM_CNT_TRAJ, c_traj);
    }
}

#define SBM_DEFINE_DP5_PUSH(NAME, BLOCK)                                                    \
    extern "C" __global__ void __launch_bounds__(BLOCK) NAME(const SbmSolveArgs a)          \
    { sbm_dp5_push_impl<BLOCK>(a); }

// x^(1/4) for the Rosenbrock step-size controller: single precision is plenty for a step-size factor that is clamped to
// [1/6, 5] (the double-precision pow() was ~250 FP64 instructions per step, a fifth of the kernel's FP64 work on C3)
__device__ __forceinline__ double sbm_root4(const double x)
{
    const float v = fminf(fmaxf((float)x, 1e-36f), 1e36f);
    return (double)sqrtf(sqrtf(v));
}

// in-register LU, partial pivoting by predicated row swaps (all indices compile-time after unrolling)
__device__ __forceinline__ bool sbm_lu_factor(double (&A)[SBM_N * SBM_N], int (&piv)[SBM_N])
{
    constexpr int N = SBM_N;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        int r = k;
        double best = fabs(A[k * N + k]);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double v = fabs(A[i * N + k]);
            if (v > best) { best = v; r = i; }
        }
        piv[k] = r;
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            if (N <= 6) {
                const bool sw = (r == i);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double ak = A[k * N + j], ai = A[i * N + j];
                    A[k * N + j] = sw ? ai : ak;
                    A[i * N + j] = sw ? ak : ai;
                }
            } else if (r == i) {
#pragma unroll
                for (int j = 0; j < N; ++j) { const double tmp = A[k * N + j]; A[k * N + j] = A[i * N + j]; A[i * N + j] = tmp; }
            }
        }
        if (!(best > 0.0)) ok = false;
        const double inv = sbm_rcp(A[k * N + k]);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double l = A[i * N + k] * inv;
            A[i * N + k] = l;
#pragma unroll
            for (int j = k + 1; j < N; ++j) A[i * N + j] = fma(-l, A[k * N + j], A[i * N + j]);
        }
        A[k * N + k] = inv;          // store the reciprocal pivot
    }
    return ok;
}

__device__ __forceinline__ void sbm_lu_solve(const double (&A)[SBM_N * SBM_N], const int (&piv)[SBM_N],
                                             double (&b)[SBM_N])
{
    constexpr int N = SBM_N;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        // (small systems: the swap as selects - written as a branch, ptxas turns the chain into a dynamically indexed copy of b
        //  in local memory: 8 % of the executed instructions of C3 were STL / LDL, profiles/r02_rodas4_plain.txt)
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            if (N <= 6) {
                const bool sw = (piv[k] == i);
                const double bk = b[k], bi = b[i];
                b[k] = sw ? bi : bk;
                b[i] = sw ? bk : bi;
            } else if (piv[k] == i) { const double tmp = b[k]; b[k] = b[i]; b[i] = tmp; }
        }
#pragma unroll
        for (int i = k + 1; i < N; ++i) b[i] = fma(-A[i * N + k], b[k], b[i]);
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        double s = b[k];
#pragma unroll
        for (int j = k + 1; j < N; ++j) s = fma(-A[k * N + j], b[j], s);
        b[k] = s * A[k * N + k];
    }
}

template <int BLOCK, int DENSE = 0>
__device__ __forceinline__ void sbm_rodas4_impl(const SbmSolveArgs& a)
{
    constexpr int N = SBM_N;
    __shared__ double sp[BLOCK * SBM_NPP];
    const int tid = threadIdx.x;
    const long long traj = (long long)blockIdx.x * BLOCK + tid;
    if (traj >= a.B * (long long)a.E) return;
    const int e = (int)(traj % a.E);
    const long long b = traj / a.E;

    double* p = sp + tid * SBM_NPP;
    for (int m = 0; m < SBM_NP; ++m) {
        const int c = a.pmap[e * SBM_NP + m];
        p[m] = (c >= 0) ? exp(a.theta[b * a.theta_sb + c * a.theta_sp]) : a.pfix[e * SBM_NP + m];
    }

    const int cid[1] = {0};
    const int pcol[1] = {-1};
    const double pscale[1] = {0.0};
    double y[1][N], yn[1][N];
#pragma unroll
    for (int i = 0; i < N; ++i) y[0][i] = a.y0[e * N + i];

    const double rtol = a.rtol, atol = a.atol;
    const int kt0 = a.tout_off[e], kt1 = a.tout_off[e + 1];
    const double tend = (kt1 > kt0) ? a.tout[kt1 - 1] : 0.0;
    unsigned nstep = 0, nrej = 0, nrhs = 0, nlu = 0;
    int status = SBM_OK;
    double t = 0.0;

    double f0[N], fs[N], ft[N];
    double k1[N], k2[N], k3[N], k4[N], k5[N], k6[N];
    double A[N * N];
    int piv[N];

    // initial step from the local rate of change (order-4 variant of Hairer's rule)
    double h;
    {
        sbm_f(t, y[0], p, f0);
        nrhs++;
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = atol + rtol * fabs(y[0][i]);
            d0 = fma(y[0][i] / sk, y[0][i] / sk, d0);
            d1 = fma(f0[i] / sk, f0[i] / sk, d1);
        }
        d0 = sqrt(d0 / N); d1 = sqrt(d1 / N);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, tend);
#pragma unroll
        for (int i = 0; i < N; ++i) yn[0][i] = fma(h0, f0[i], y[0][i]);
        sbm_f(t + h0, yn[0], p, fs);
        nrhs++;
        double d2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = atol + rtol * fabs(y[0][i]);
            const double q = (fs[i] - f0[i]) / sk;
            d2 = fma(q, q, d2);
        }
        d2 = sqrt(d2 / N) / h0;
        const double dm = fmax(d1, d2);
        const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.25);
        h = fmin(fmin(100.0 * h0, h1), tend);
        if (a.h0 > 0.0) h = fmin(a.h0, tend);
    }

    bool last_rejected = false, first = true, poison = false;
    double hacc = h, erracc = 1e-2;
    unsigned naccept = 0;

    // DENSE: one pass to the last output time; the others are read off the continuous extension of the step that passes
    // them (Hairer & Wanner's RODAS, CONTRO) - kd is the next output time to be written
    int kd = kt0;
    for (int kt = DENSE ? (kt1 > kt0 ? kt1 - 1 : kt1) : kt0; kt < kt1; ++kt) {
        const double tk = a.tout[kt];
        while (status == SBM_OK && t < tk) {
            const double hprop = h;
            bool clipped = false;
            if (t + 1.01 * h >= tk) { h = tk - t; clipped = true; }
            if (nstep >= (unsigned)a.max_steps) { status = SBM_ERR_MAXSTEPS; break; }
            if (!clipped && !(h > 1e-15 * fmax(t, 1e-300))) { status = SBM_ERR_UNDERFLOW; break; }
            nstep++;

            // E = I/(h*gamma) - J, factored once per step
            sbm_f(t, y[0], p, f0);
            sbm_jac_dense(t, y[0], p, A);
#if !SBM_AUTONOMOUS
            sbm_dfdt(t, y[0], p, ft);
#else
#pragma unroll
            for (int i = 0; i < N; ++i) ft[i] = 0.0;
#endif
            const double hi = sbm_rcp(h);
            const double fac = hi * (1.0 / RO_GAMMA);
#pragma unroll
            for (int q = 0; q < N * N; ++q) A[q] = -A[q];
#pragma unroll
            for (int i = 0; i < N; ++i) A[i * N + i] += fac;
            const bool lu_ok = sbm_lu_factor(A, piv);
            nlu++;

            // stage 1
#pragma unroll
            for (int i = 0; i < N; ++i) k1[i] = fma(h * RO_D1, ft[i], f0[i]);
            sbm_lu_solve(A, piv, k1);
            // stage 2
#pragma unroll
            for (int i = 0; i < N; ++i) yn[0][i] = fma(RO_A21, k1[i], y[0][i]);
            sbm_f(t + RO_C2 * h, yn[0], p, fs);
#pragma unroll
            for (int i = 0; i < N; ++i) k2[i] = fma(h * RO_D2, ft[i], fma(hi * RO_C21, k1[i], fs[i]));
            sbm_lu_solve(A, piv, k2);
            // stage 3
#pragma unroll
            for (int i = 0; i < N; ++i) yn[0][i] = fma(RO_A32, k2[i], fma(RO_A31, k1[i], y[0][i]));
            sbm_f(t + RO_C3 * h, yn[0], p, fs);
#pragma unroll
            for (int i = 0; i < N; ++i)
                k3[i] = fma(h * RO_D3, ft[i], fma(hi * RO_C32, k2[i], fma(hi * RO_C31, k1[i], fs[i])));
            sbm_lu_solve(A, piv, k3);
            // stage 4
#pragma unroll
            for (int i = 0; i < N; ++i)
                yn[0][i] = fma(RO_A43, k3[i], fma(RO_A42, k2[i], fma(RO_A41, k1[i], y[0][i])));
            sbm_f(t + RO_C4 * h, yn[0], p, fs);
#pragma unroll
            for (int i = 0; i < N; ++i)
                k4[i] = fma(h * RO_D4, ft[i], fma(hi * RO_C43, k3[i], fma(hi * RO_C42, k2[i],
                        fma(hi * RO_C41, k1[i], fs[i]))));
            sbm_lu_solve(A, piv, k4);
            // stage 5
#pragma unroll
            for (int i = 0; i < N; ++i)
                yn[0][i] = fma(RO_A54, k4[i], fma(RO_A53, k3[i], fma(RO_A52, k2[i], fma(RO_A51, k1[i], y[0][i]))));
            sbm_f(t + h, yn[0], p, fs);
#pragma unroll
            for (int i = 0; i < N; ++i)
                k5[i] = fma(hi * RO_C54, k4[i], fma(hi * RO_C53, k3[i], fma(hi * RO_C52, k2[i],
                        fma(hi * RO_C51, k1[i], fs[i]))));
            sbm_lu_solve(A, piv, k5);
            // stage 6 (embedded 3rd-order solution is yn + k5)
#pragma unroll
            for (int i = 0; i < N; ++i) yn[0][i] += k5[i];
            sbm_f(t + h, yn[0], p, fs);
#pragma unroll
            for (int i = 0; i < N; ++i)
                k6[i] = fma(hi * RO_C65, k5[i], fma(hi * RO_C64, k4[i], fma(hi * RO_C63, k3[i],
                        fma(hi * RO_C62, k2[i], fma(hi * RO_C61, k1[i], fs[i])))));
            sbm_lu_solve(A, piv, k6);
#pragma unroll
            for (int i = 0; i < N; ++i) yn[0][i] += k6[i];
            nrhs += 7;

            double sq = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double sk = atol + rtol * fmax(fabs(y[0][i]), fabs(yn[0][i]));
                const double q = k6[i] * sbm_rcp(sk);
                sq = fma(q, q, sq);
            }
            const double err = sqrt(sq * (1.0 / N));

            if (lu_ok && err <= 1.0) {
                // accept; Gustafsson predictive controller as in Hairer's RODAS
                double facr = fmax(1.0 / 6.0, fmin(5.0, sbm_root4(err) / 0.9));
                if (naccept >= 1) {
                    double facgus = (hacc / h) * sbm_root4(err * err / erracc) / 0.9;
                    facgus = fmax(1.0 / 6.0, fmin(5.0, facgus));
                    facr = fmax(facr, facgus);
                }
                hacc = h;
                erracc = fmax(1e-2, err);
                naccept++;
                if (DENSE) {
                    const double tn = clipped ? tk : t + h;
                    while (kd < kt1 - 1 && a.tout[kd] <= tn) {
                        const double th = (a.tout[kd] - t) / h, th1 = 1.0 - th;
                        double yi[1][N];
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            const double c1 = fma(RO_D25, k5[i], fma(RO_D24, k4[i], fma(RO_D23, k3[i], fma(RO_D22, k2[i], RO_D21 * k1[i]))));
                            const double c2 = fma(RO_D35, k5[i], fma(RO_D34, k4[i], fma(RO_D33, k3[i], fma(RO_D32, k2[i], RO_D31 * k1[i]))));
                            yi[0][i] = fma(th, fma(th1, fma(th, c2, c1), yn[0][i]), th1 * y[0][i]);
                        }
                        sbm_emit_rows<1, 1, 0>(a, b, kd, yi, cid, pcol, pscale, false);
                        kd++;
                    }
                }
                t = clipped ? tk : t + h;
#pragma unroll
                for (int i = 0; i < N; ++i) y[0][i] = yn[0][i];
                double hnew = h / facr;
                if (last_rejected) hnew = fmin(hnew, h);
                if (clipped) hnew = fmax(hnew, hprop);
                h = hnew;
                last_rejected = false;
                first = false;
            } else {
                nrej++;
                if (lu_ok && err == err && err < 1e300) {
                    const double facr = fmax(1.0 / 6.0, fmin(5.0, sbm_root4(err) / 0.9));
                    h = first ? 0.1 * h : h / facr;
                } else {
                    h *= 0.5;
                    if (h < 1e-300) { status = SBM_ERR_NONFINITE; break; }
                }
                last_rejected = true;
            }
        }
        if (status != SBM_OK) poison = true;
        if (DENSE) {
            // a failed trajectory flushes what it has not written as NaN
            for (; poison && kd < kt; ++kd) sbm_emit_rows<1, 1, 0>(a, b, kd, y, cid, pcol, pscale, true);
        }
        sbm_emit_rows<1, 1, 0>(a, b, kt, y, cid, pcol, pscale, poison);
    }

    a.status[traj] = status;
    atomicAdd(a.counters + SBM_CNT_STEPS, (unsigned long long)nstep);
    atomicAdd(a.counters + SBM_CNT_REJECT, (unsigned long long)nrej);
    atomicAdd(a.counters + SBM_CNT_RHS, (unsigned long long)nrhs);
    atomicAdd(a.counters + SBM_CNT_LU, (unsigned long long)nlu);
    atomicAdd(a.counters + SBM_CNT_TRAJ, 1ull);
}

// ---------------------------------------------------------------------------------------------
// RODAS4 with forward sensitivities (stiff models whose Jacobian the optimiser needs: the reference's LSODA switches to BDF
// on the augmented system, ode_model.py:122-123).  One thread per trajectory, autonomous models of at most 12 states.
//
// The augmented system z = [y | S_1 .. S_NF], S_j' = J(y) S_j + G_j(y), has the block lower-triangular Jacobian
//     dF/dz = [ J  0 ... ; B_1  J ... ; ... ],   B_j = d(J S_j + G_j)/dy,
// so every Rosenbrock stage is solved by block forward substitution with ONE LU of E = I/(h gamma) - J per step:
//     E u_i(y)   = rhs_i(y)
//     E u_i(S_j) = rhs_i(S_j) + B_j u_i(y),       B_j u = DJ(u) S_j + DG(u)[:, j]     (sbm_dJG: exact, emitted from SymPy)
// A Rosenbrock method (unlike a W-method) keeps its order only with the exact Jacobian, hence the exact B_j.  The error
// estimate (u_6) and the step-size control run over the whole augmented state, as in the Dormand-Prince kernels.
// The stage loop is not unrolled: one copy of the model code, coefficient tables in constant memory; the increments of the
// sensitivity columns live in local memory (indexed by column at run time).
// ---------------------------------------------------------------------------------------------
#if SBM_AUTONOMOUS && defined(SBM_HAS_DJG)
__constant__ double sbm_ro_a[6][5] = {
    {0.0, 0.0, 0.0, 0.0, 0.0},
    {RO_A21_V, 0.0, 0.0, 0.0, 0.0},
    {RO_A31_V, RO_A32_V, 0.0, 0.0, 0.0},
    {RO_A41_V, RO_A42_V, RO_A43_V, 0.0, 0.0},
    {RO_A51_V, RO_A52_V, RO_A53_V, RO_A54_V, 0.0},
    {RO_A51_V, RO_A52_V, RO_A53_V, RO_A54_V, 1.0}};
__constant__ double sbm_ro_c[6][5] = {
    {0.0, 0.0, 0.0, 0.0, 0.0},
    {RO_C21_V, 0.0, 0.0, 0.0, 0.0},
    {RO_C31_V, RO_C32_V, 0.0, 0.0, 0.0},
    {RO_C41_V, RO_C42_V, RO_C43_V, 0.0, 0.0},
    {RO_C51_V, RO_C52_V, RO_C53_V, RO_C54_V, 0.0},
    {RO_C61_V, RO_C62_V, RO_C63_V, RO_C64_V, RO_C65_V}};

// small systems (Robertson: 3 states x 3 parameters): stages and columns unrolled, so that the stage increments of the state
// and of the sensitivity columns are register arrays with compile-time indices instead of local memory
#if SBM_N * SBM_NF <= 12
#define SBM_RO_UNROLL_ST 6
#define SBM_RO_UNROLL_J SBM_NF
#define SBM_RO_UNROLL_L 5
#else
#define SBM_RO_UNROLL_ST 1
#define SBM_RO_UNROLL_J 1
#define SBM_RO_UNROLL_L 1
#endif
template <int BLOCK>
__device__ __forceinline__ void sbm_rodas4_sens_impl(const SbmSolveArgs& a)
{
    constexpr int N = SBM_N, NF = SBM_NF;
    __shared__ double sp[BLOCK * SBM_NPP];
    const int tid = threadIdx.x;
    const long long traj = (long long)blockIdx.x * BLOCK + tid;
    if (traj >= a.B * (long long)a.E) return;
    const int e = (int)(traj % a.E);
    const long long b = traj / a.E;

    double* p = sp + tid * SBM_NPP;
    for (int m = 0; m < SBM_NP; ++m) {
        const int c = a.pmap[e * SBM_NP + m];
        p[m] = (c >= 0) ? exp(a.theta[b * a.theta_sb + c * a.theta_sp]) : a.pfix[e * SBM_NP + m];
    }

    double y[N], Y[N], f0[N], fi[N];
    double ky[6][N];
    double S[NF][N], U[6][NF][N];
    double Jv0[SBM_NJ], Gv0[SBM_NG], Jvi[SBM_NJ], Gvi[SBM_NG], DJv[SBM_NJ], DGv[SBM_NG];
    double A[N * N];
    int piv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = a.y0[e * N + i];
#pragma unroll SBM_RO_UNROLL_J
    for (int j = 0; j < NF; ++j)
#pragma unroll
        for (int i = 0; i < N; ++i) S[j][i] = 0.0;

    const double rtol = a.rtol, atol = a.atol;
    const int kt0 = a.tout_off[e], kt1 = a.tout_off[e + 1];
    const double tend = (kt1 > kt0) ? a.tout[kt1 - 1] : 0.0;
    unsigned nstep = 0, nrej = 0, nrhs = 0, nlu = 0;
    int status = SBM_OK;
    double t = 0.0;

    // initial step from the local rate of change of the states (as sbm_rodas4_impl)
    double h;
    {
        sbm_f(t, y, p, f0);
        nrhs++;
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = atol + rtol * fabs(y[i]);
            d0 = fma(y[i] / sk, y[i] / sk, d0);
            d1 = fma(f0[i] / sk, f0[i] / sk, d1);
        }
        d0 = sqrt(d0 / N); d1 = sqrt(d1 / N);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, tend);
#pragma unroll
        for (int i = 0; i < N; ++i) Y[i] = fma(h0, f0[i], y[i]);
        sbm_f(t + h0, Y, p, fi);
        nrhs++;
        double d2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = atol + rtol * fabs(y[i]);
            const double q = (fi[i] - f0[i]) / sk;
            d2 = fma(q, q, d2);
        }
        d2 = sqrt(d2 / N) / h0;
        const double dm = fmax(d1, d2);
        const double h1 = (dm <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / dm, 0.25);
        // the sensitivities start at zero and grow from G: a cautious first step (the controller opens it up quickly)
        h = 0.01 * fmin(fmin(100.0 * h0, h1), tend);
        if (a.h0 > 0.0) h = fmin(a.h0, tend);
    }

    bool last_rejected = false, first = true, poison = false;
    double hacc = h, erracc = 1e-2;
    unsigned naccept = 0;
    const double inv_naug = 1.0 / (double)(N * (NF + 1));

    for (int kt = kt0; kt < kt1; ++kt) {
        const double tk = a.tout[kt];
        while (status == SBM_OK && t < tk) {
            const double hprop = h;
            bool clipped = false;
            if (t + 1.01 * h >= tk) { h = tk - t; clipped = true; }
            if (nstep >= (unsigned)a.max_steps) { status = SBM_ERR_MAXSTEPS; break; }
            if (!clipped && !(h > 1e-15 * fmax(t, 1e-300))) { status = SBM_ERR_UNDERFLOW; break; }
            nstep++;

            sbm_fJG<1>(t, y, p, f0, Jv0, Gv0);
            sbm_jac_dense(t, y, p, A);
            const double hi = sbm_rcp(h);
            const double fac = hi * (1.0 / RO_GAMMA);
#pragma unroll
            for (int q = 0; q < N * N; ++q) A[q] = -A[q];
#pragma unroll
            for (int i = 0; i < N; ++i) A[i * N + i] += fac;
            const bool lu_ok = sbm_lu_factor(A, piv);
            nlu++;

#pragma unroll SBM_RO_UNROLL_ST
            for (int st = 0; st < 6; ++st) {
                double ca[5], cc[5];
#pragma unroll
                for (int l = 0; l < 5; ++l) { ca[l] = sbm_ro_a[st][l]; cc[l] = hi * sbm_ro_c[st][l]; }
                // ---- state block ----
                double r[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    double v = y[i], w = 0.0;
#pragma unroll SBM_RO_UNROLL_L
                    for (int l = 0; l < st; ++l) { v = fma(ca[l], ky[l][i], v); w = fma(cc[l], ky[l][i], w); }
                    Y[i] = v;
                    r[i] = w;
                }
                const double* Jp = Jv0;
                const double* Gp = Gv0;
                const double* fp = f0;
                if (st > 0) {
                    sbm_fJG<1>(t, Y, p, fi, Jvi, Gvi);
                    Jp = Jvi; Gp = Gvi; fp = fi;
                }
#pragma unroll
                for (int i = 0; i < N; ++i) r[i] += fp[i];
                sbm_lu_solve(A, piv, r);
#pragma unroll
                for (int i = 0; i < N; ++i) ky[st][i] = r[i];
                // ---- sensitivity blocks: same LU, right-hand side + B_j u_i(y) ----
                sbm_dJG(t, y, p, r, DJv, DGv);
#pragma unroll SBM_RO_UNROLL_J
                for (int j = 0; j < NF; ++j) {
                    double Sj[N], w[N], r1[N], r2[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        double v = S[j][i], ww = 0.0;
#pragma unroll SBM_RO_UNROLL_L
                    for (int l = 0; l < st; ++l) { v = fma(ca[l], U[l][j][i], v); ww = fma(cc[l], U[l][j][i], ww); }
                        Sj[i] = v;
                        w[i] = ww;
                    }
                    sbm_col(fp, Jp, Gp, Sj, j + 1, r1);
                    sbm_col(fp, DJv, DGv, S[j], j + 1, r2);
#pragma unroll
                    for (int i = 0; i < N; ++i) w[i] += r1[i] + r2[i];
                    sbm_lu_solve(A, piv, w);
#pragma unroll
                    for (int i = 0; i < N; ++i) U[st][j][i] = w[i];
                }
            }
            nrhs += 6;

            // new state = stage value of stage 6 + u_6; u_6 is the error estimate.  (Y holds the stage value of stage 6.)
            double sq = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                Y[i] += ky[5][i];
                const double sk = atol + rtol * fmax(fabs(y[i]), fabs(Y[i]));
                const double q = ky[5][i] * sbm_rcp(sk);
                sq = fma(q, q, sq);
            }
#pragma unroll SBM_RO_UNROLL_J
            for (int j = 0; j < NF; ++j) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    double v = S[j][i];
#pragma unroll
                    for (int l = 0; l < 5; ++l) v = fma(sbm_ro_a[5][l], U[l][j][i], v);
                    const double u6 = U[5][j][i];
                    v += u6;
                    const double sk = atol + rtol * fmax(fabs(S[j][i]), fabs(v));
                    const double q = u6 * sbm_rcp(sk);
                    sq = fma(q, q, sq);
                    U[5][j][i] = v;                    // the slot now holds the candidate S_j(t + h)
                }
            }
            const double err = sqrt(sq * inv_naug);

            if (lu_ok && err <= 1.0) {
                double facr = fmax(1.0 / 6.0, fmin(5.0, sbm_root4(err) / 0.9));
                if (naccept >= 1) {
                    double facgus = (hacc / h) * sbm_root4(err * err / erracc) / 0.9;
                    facgus = fmax(1.0 / 6.0, fmin(5.0, facgus));
                    facr = fmax(facr, facgus);
                }
                hacc = h;
                erracc = fmax(1e-2, err);
                naccept++;
                t = clipped ? tk : t + h;
#pragma unroll
                for (int i = 0; i < N; ++i) y[i] = Y[i];
#pragma unroll SBM_RO_UNROLL_J
                for (int j = 0; j < NF; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) S[j][i] = U[5][j][i];
                double hnew = h / facr;
                if (last_rejected) hnew = fmin(hnew, h);
                if (clipped) hnew = fmax(hnew, hprop);
                h = hnew;
                last_rejected = false;
                first = false;
            } else {
                nrej++;
                if (lu_ok && err == err && err < 1e300) {
                    const double facr = fmax(1.0 / 6.0, fmin(5.0, sbm_root4(err) / 0.9));
                    h = first ? 0.1 * h : h / facr;
                } else {
                    h *= 0.5;
                    if (h < 1e-300) { status = SBM_ERR_NONFINITE; break; }
                }
                last_rejected = true;
            }
        }
        if (status != SBM_OK) poison = true;
        // ---- observable rows of this output time: value and d/dtheta row (chain rule to log-parameters, base_project.py:443-488) ----
        const int r0 = a.orow_off[kt], r1 = a.orow_off[kt + 1];
        for (int r = r0; r < r1; ++r) {
            const unsigned long long mask = a.row_mask[r];
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) acc += ((mask >> i) & 1ull) ? y[i] : 0.0;
            if (poison) acc = __longlong_as_double(0x7ff8000000000000ll);
            a.sims[b * a.n_rows + a.row_sim[r]] = acc;
            if (a.jac != nullptr) {
#pragma unroll SBM_RO_UNROLL_J
                for (int j = 0; j < NF; ++j) {
                    const int m = sbm_free_idx[j];
                    const int pc = a.jac_raw ? j : a.pmap[e * SBM_NP + m];
                    if (pc < 0) continue;
                    double v = 0.0;
#pragma unroll
                    for (int i = 0; i < N; ++i) v += ((mask >> i) & 1ull) ? S[j][i] : 0.0;
                    v *= a.jac_raw ? 1.0 : p[m];
                    if (poison) v = __longlong_as_double(0x7ff8000000000000ll);
                    double* dst = a.jac + ((b * a.n_rows + a.row_jac[r]) * (long long)a.n_proj + pc);
                    if (a.jac_atomic) atomicAdd(dst, v); else *dst = v;
                }
            }
        }
    }

    a.status[traj] = status;
    atomicAdd(a.counters + SBM_CNT_STEPS, (unsigned long long)nstep);
    atomicAdd(a.counters + SBM_CNT_REJECT, (unsigned long long)nrej);
    atomicAdd(a.counters + SBM_CNT_RHS, (unsigned long long)nrhs);
    atomicAdd(a.counters + SBM_CNT_LU, (unsigned long long)nlu);
    atomicAdd(a.counters + SBM_CNT_TRAJ, 1ull);
}

#define SBM_DEFINE_RODAS4_SENS(NAME, BLOCK)                                                 \
    extern "C" __global__ void __launch_bounds__(BLOCK) NAME(const SbmSolveArgs a)          \
    { sbm_rodas4_sens_impl<BLOCK>(a); }
#endif  /* SBM_AUTONOMOUS && SBM_HAS_DJG */

// ---------------------------------------------------------------------------------------------
// kernel entry points are instantiated by the host (sbm_api.cu appends SBM_DEFINE_* lines)
// ---------------------------------------------------------------------------------------------
#define SBM_DEFINE_DP5(NAME, G, C, SENS, BLOCK)                                             \
    extern "C" __global__ void __launch_bounds__(BLOCK) NAME(const SbmSolveArgs a)          \
    { sbm_dp5_impl<G, C, SENS, BLOCK>(a); }

// states only, output times read off the continuous extension (a kernel of its own: the interpolation keeps three state
// vectors live, which costs the 25-state model of C5 its spill-free register allocation)
#define SBM_DEFINE_DP5_DENSE(NAME, BLOCK)                                                   \
    extern "C" __global__ void __launch_bounds__(BLOCK) NAME(const SbmSolveArgs a)          \
    { sbm_dp5_impl<1, 1, 0, BLOCK, 1>(a); }

#define SBM_DEFINE_RODAS4(NAME, BLOCK)                                                      \
    extern "C" __global__ void __launch_bounds__(BLOCK) NAME(const SbmSolveArgs a)          \
    { sbm_rodas4_impl<BLOCK>(a); }
#define SBM_DEFINE_RODAS4_DENSE(NAME, BLOCK)                                                \
    extern "C" __global__ void __launch_bounds__(BLOCK) NAME(const SbmSolveArgs a)          \
    { sbm_rodas4_impl<BLOCK, 1>(a); }
