// =================================================================================================
// sbm_dp5_ct.cuh - "CTA-transposed" Dormand-Prince 5(4) integrator for sensitivity-augmented systems
//
// Compiled by NVRTC after the model functions and sbm_ode_kernels.cuh (which defines the DP_* tableau).
// Same job as sbm_dp5_sens (reference model/ode_model.py:83-126 + project/utils.py:29-89 +
// project/base_project.py:443-488 for every (experiment, parameter set) pair), different mapping:
//
//   lane  = trajectory     32 trajectories per CTA, each with its own t, h and accept/reject decision
//   warp  = column group   warps 0..W-2 ("sensitivity warps") own C sensitivity columns S_j = dy/dp_j each, of
//                          ALL 32 trajectories; warp W-1 (the "state warp") owns the state column y, evaluates
//                          the model, runs the step controller and carries the last CS sensitivity columns
//                          (W = 8 keeps two warps per scheduler, i.e. 255 registers per thread)
//
// Why: with lane = column (sbm_dp5_sens) every lane of a group re-evaluates the model's primal part
// (f, J = df/dy, G = df/dp: rate laws, reciprocals, exp/log) - measured to be more than half of the
// FP64 instructions.  Here the primal part is evaluated ONCE per trajectory and stage by the state
// warp (thread = trajectory), published through shared memory ([entry][lane], conflict-free), and the
// sensitivity warps only run  dS_j/dt = J S_j + G_j  on their own columns.  The state column does not
// depend on the sensitivities, so the state warp runs ONE STAGE AHEAD: while the sensitivity warps
// work on stage s, the state warp evaluates the model at stage s+1 (double-buffered J/G).  One CTA
// barrier per stage.
//
// Storage per (trajectory, column): the 5 vectors  y, S1..S4  of a "push" formulation of the tableau
//   stages 2-4 pull   Y_s = y + h sum a_sj k_j           (k1..k3 live in S1..S3)
//   after k4          S1 <- Y5, S2 <- Y6 - h a65 k5, S3 <- Y7 - h (a75 k5 + a76 k6), S4 <- h (e1 k1 + e3 k3 + e4 k4)
//   stages 5-7 push   k5, k6, k7 are added into S2, S3, S4 as soon as they exist; S1 <- k7 (FSAL)
// which needs two vectors fewer than keeping k1..k6.  y, S2, S3 live in registers, S1 and S4 in shared
// memory ([element][lane]).  Full trajectories never exist; only the observable rows go to HBM.
//
// A rejected step has destroyed k1 (S1 holds k7 of the rejected trial), so the next "step" of that lane
// is a zero-length step, which re-evaluates F(y) in every stage and leaves k1 in S1 again.  The same
// zero-length step starts every trajectory (k1 = F(y0) and the norms for the initial step size).
// Lanes fetch new trajectories from a global queue as they finish (persistent CTAs).
// =================================================================================================

#define SBM_CT_F_ACCEPT 1
#define SBM_CT_F_EMIT   2
#define SBM_CT_F_POISON 4
#define SBM_CT_F_RESET  8

#define SBM_CT_ALIVE    1
#define SBM_CT_ANYEMIT  2
#define SBM_CT_ANYINIT  4

#define SBM_CT_IDLE   0
#define SBM_CT_ZERO   1
#define SBM_CT_NORMAL 2

__device__ __forceinline__ void sbm_ct_bar(const int nthreads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// reciprocal for error weights: hardware seed + one Newton step (relative error ~1e-10, plenty for a norm)
__device__ __forceinline__ double sbm_rcp_fast(const double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
}

// stage input from the vectors of one element
template <int S>
__device__ __forceinline__ double sbm_ct_input(const double h, const double y, const double v1, const double s2,
                                               const double s3)
{
    if (S == 2) return fma(h * DP_A21, v1, y);
    if (S == 3) return fma(h, fma(DP_A32, s2, DP_A31 * v1), y);
    if (S == 4) return fma(h, fma(DP_A43, s3, fma(DP_A42, s2, DP_A41 * v1)), y);
    if (S == 5) return v1;
    if (S == 6) return s2;
    return s3;
}

// what stage S does with its derivative kk (one element)
template <int S>
__device__ __forceinline__ void sbm_ct_update(const double kk, const double h, const double y, double& v1,
                                              double& s2, double& s3, double& v4)
{
    if (S == 2) {
        s2 = kk;
    } else if (S == 3) {
        s3 = kk;
    } else if (S == 4) {
        const double k1 = v1, k2 = s2, k3 = s3;
        v1 = fma(h, fma(DP_A54, kk, fma(DP_A53, k3, fma(DP_A52, k2, DP_A51 * k1))), y);
        s2 = fma(h, fma(DP_A64, kk, fma(DP_A63, k3, fma(DP_A62, k2, DP_A61 * k1))), y);
        s3 = fma(h, fma(DP_A74, kk, fma(DP_A73, k3, DP_A71 * k1)), y);
        v4 = h * fma(DP_E4, kk, fma(DP_E3, k3, DP_E1 * k1));
    } else if (S == 5) {
        const double hk = h * kk;
        s2 = fma(DP_A65, hk, s2);
        s3 = fma(DP_A75, hk, s3);
        v4 = fma(DP_E5, hk, v4);
    } else if (S == 6) {
        const double hk = h * kk;
        s3 = fma(DP_A76, hk, s3);
        v4 = fma(DP_E6, hk, v4);
    } else {
        v1 = kk;                               // FSAL: k7 is the next k1
        v4 = fma(h * DP_E7, kk, v4);           // embedded error estimate of this element
    }
}

// -------------------------------------------------------------------------------------------------
// sensitivity warps: one stage on C columns of 32 trajectories
// -------------------------------------------------------------------------------------------------
template <int S, int C>
__device__ __forceinline__ void sbm_ct_sens_stage(const double h, const double* __restrict__ Jp,
                                                  const double* __restrict__ Gp, const int col0,
                                                  double (&y)[C][SBM_N], double (&s2)[C][SBM_N],
                                                  double (&s3)[C][SBM_N], double* __restrict__ V1,
                                                  double* __restrict__ V4, const double rtol, const double atol,
                                                  const bool anyinit, double& sq, double& d0, double& d1)
{
    constexpr int N = SBM_N;
    double w[C][N];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double v1 = (S <= 5) ? V1[(c * N + i) * 32] : 0.0;
            w[c][i] = sbm_ct_input<S>(h, y[c][i], v1, s2[c][i], s3[c][i]);
        }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int q = sbm_gtab[(col0 + c) * N + i];       // warp-uniform
            acc[c] = (q >= 0) ? Gp[q * 32] : 0.0;
        }
        sbm_jrow<C>(i, Jp, w, acc);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            double v1 = 0.0, v4 = 0.0;
            if (S == 4) v1 = V1[(c * N + i) * 32];
            if (S >= 5) v4 = V4[(c * N + i) * 32];
            sbm_ct_update<S>(acc[c], h, y[c][i], v1, s2[c][i], s3[c][i], v4);
            if (S == 4 || S == 7) V1[(c * N + i) * 32] = v1;
            if (S >= 4 && S <= 6) V4[(c * N + i) * 32] = v4;
            if (S == 7) {
                const double isk = sbm_rcp_fast(fma(rtol, fmax(fabs(y[c][i]), fabs(s3[c][i])), atol));
                const double q = v4 * isk;
                sq = fma(q, q, sq);
                if (anyinit) {
                    const double q0 = y[c][i] * isk, q1 = acc[c] * isk;
                    d0 = fma(q0, q0, d0);
                    d1 = fma(q1, q1, d1);
                }
            }
        }
    }
}

// start of a step for a block of CW sensitivity columns: apply the accept decision, write the Jacobian
// rows of an output time, restart on a new trajectory, and clear k1 before a zero-length step
template <int CW>
__device__ __forceinline__ void sbm_ct_cols_begin(const SbmSolveArgs& a, const int info, const int cflags,
                                                  const double h, const int col0, const int lane,
                                                  double (&y)[CW][SBM_N], const double (&s3)[CW][SBM_N],
                                                  double* __restrict__ V1, const double* __restrict__ pb,
                                                  const long long* __restrict__ trj, const int* __restrict__ ktb)
{
    constexpr int N = SBM_N;
    if (info & SBM_CT_F_ACCEPT) {
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int i = 0; i < N; ++i) y[c][i] = s3[c][i];
    }
    if (cflags & SBM_CT_ANYEMIT) {
        if (info & SBM_CT_F_EMIT) {
            // observable rows of this output time: d(observable)/d(theta) for these columns
            // (mapping: project/utils.py:29-45, 71-89; chain rule + scatter: base_project.py:443-488)
            const long long traj = trj[lane];
            const int e = (int)(traj % a.E);
            const long long b = traj / a.E;
            const int kt = ktb[lane];
            const bool poison = (info & SBM_CT_F_POISON) != 0;
            const int r0 = a.orow_off[kt], r1 = a.orow_off[kt + 1];
            for (int r = r0; r < r1; ++r) {
                const unsigned long long mask = a.row_mask[r];
                double* row = a.jac + (b * a.n_rows + a.row_jac[r]) * (long long)a.n_proj;
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                    if (col0 + c < SBM_NF) {
                        double acc = 0.0;
#pragma unroll
                        for (int i = 0; i < N; ++i) acc += ((mask >> i) & 1ull) ? y[c][i] : 0.0;
                        if (poison) acc = __longlong_as_double(0x7ff8000000000000ll);
                        const int m = sbm_free_idx[col0 + c];
                        int pc;
                        double sc;
                        if (a.jac_raw) { pc = col0 + c; sc = 1.0; }
                        else { pc = a.pmap[e * SBM_NP + m]; sc = pb[m * 32 + lane]; }   // d/dtheta = p d/dp
                        if (pc >= 0) {
                            const double v = acc * sc;
                            if (a.jac_atomic) atomicAdd(row + pc, v); else row[pc] = v;
                        }
                    }
                }
            }
        }
    }
    if (info & SBM_CT_F_RESET) {
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int i = 0; i < N; ++i) y[c][i] = 0.0;
    }
    if (h == 0.0) {
        // zero-length step (start-up, refresh after a rejection, idle lane): the stale k1 must not leak
        // into the stage inputs (0 * Inf = NaN after a rejected non-finite trial)
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int i = 0; i < N; ++i) V1[(c * N + i) * 32] = 0.0;
    }
}

template <int CW>
__device__ __forceinline__ void sbm_ct_cols_stage(const int s, const double h, const double* __restrict__ Jb,
                                                  const double* __restrict__ Gb, const int lane, const int col0,
                                                  double (&y)[CW][SBM_N], double (&s2)[CW][SBM_N],
                                                  double (&s3)[CW][SBM_N], double* __restrict__ V1,
                                                  double* __restrict__ V4, const double rtol, const double atol,
                                                  const bool anyinit, double& sq, double& d0, double& d1)
{
    const double* Jp = Jb + (s & 1) * SBM_NJ * 32 + lane;
    const double* Gp = Gb + (s & 1) * SBM_NG * 32 + lane;
    switch (s) {
    case 2: sbm_ct_sens_stage<2, CW>(h, Jp, Gp, col0, y, s2, s3, V1, V4, rtol, atol, anyinit, sq, d0, d1); break;
    case 3: sbm_ct_sens_stage<3, CW>(h, Jp, Gp, col0, y, s2, s3, V1, V4, rtol, atol, anyinit, sq, d0, d1); break;
    case 4: sbm_ct_sens_stage<4, CW>(h, Jp, Gp, col0, y, s2, s3, V1, V4, rtol, atol, anyinit, sq, d0, d1); break;
    case 5: sbm_ct_sens_stage<5, CW>(h, Jp, Gp, col0, y, s2, s3, V1, V4, rtol, atol, anyinit, sq, d0, d1); break;
    case 6: sbm_ct_sens_stage<6, CW>(h, Jp, Gp, col0, y, s2, s3, V1, V4, rtol, atol, anyinit, sq, d0, d1); break;
    default: sbm_ct_sens_stage<7, CW>(h, Jp, Gp, col0, y, s2, s3, V1, V4, rtol, atol, anyinit, sq, d0, d1); break;
    }
}

// -------------------------------------------------------------------------------------------------
// pure sensitivity warp
// -------------------------------------------------------------------------------------------------
template <int NT, int C, int CM>
__device__ __forceinline__ void sbm_ct_sens_role(const SbmSolveArgs& a, const int warp, const int nwarps,
                                                 const int lane, double* __restrict__ Jb, double* __restrict__ Gb,
                                                 double* __restrict__ sqb, const double* __restrict__ hb,
                                                 const double* __restrict__ pb, const long long* __restrict__ trj,
                                                 const int* __restrict__ fl, const int* __restrict__ ktb,
                                                 const int* __restrict__ cta, double* __restrict__ V)
{
    constexpr int N = SBM_N;
    const int col0 = warp * C;                     // first sensitivity column (free-parameter index) of this warp
    double* V1 = V + ((warp * 2 + 0) * CM * N) * 32 + lane;
    double* V4 = V + ((warp * 2 + 1) * CM * N) * 32 + lane;
    double y[C][N], s2[C][N], s3[C][N];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            y[c][i] = 0.0; s2[c][i] = 0.0; s3[c][i] = 0.0;
            V1[(c * N + i) * 32] = 0.0;
            V4[(c * N + i) * 32] = 0.0;
        }
    const double rtol = a.rtol, atol = a.atol;

    for (;;) {
        sbm_ct_bar(NT);                            // control words + J/G of stage 2 are published
        const int cflags = cta[0];
        if (!(cflags & SBM_CT_ALIVE)) break;
        const int info = fl[lane];
        const double h = hb[lane];
        sbm_ct_cols_begin<C>(a, info, cflags, h, col0, lane, y, s3, V1, pb, trj, ktb);
        const bool anyinit = (cflags & SBM_CT_ANYINIT) != 0;
        double sq = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll 1
        for (int s = 2; s <= 7; ++s) {
            sbm_ct_cols_stage<C>(s, h, Jb, Gb, lane, col0, y, s2, s3, V1, V4, rtol, atol, anyinit, sq, d0, d1);
            if (s == 7) {
                sqb[(0 * nwarps + warp) * 32 + lane] = sq;
                if (anyinit) {
                    sqb[(1 * nwarps + warp) * 32 + lane] = d0;
                    sqb[(2 * nwarps + warp) * 32 + lane] = d1;
                }
            }
            sbm_ct_bar(NT);
        }
    }
}

// -------------------------------------------------------------------------------------------------
// state warp: step control, trajectory queue, the state column, the model's primal part, and the last
// CS sensitivity columns
// -------------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void sbm_ct_state_inputs(const double h, const double (&y)[SBM_N], const double (&s1)[SBM_N],
                                                    const double (&s2)[SBM_N], const double (&s3)[SBM_N],
                                                    double (&Y)[SBM_N])
{
#pragma unroll
    for (int i = 0; i < SBM_N; ++i) Y[i] = sbm_ct_input<S>(h, y[i], s1[i], s2[i], s3[i]);
}

template <int S>
__device__ __forceinline__ void sbm_ct_state_update(const double h, const double (&f)[SBM_N], const double (&y)[SBM_N],
                                                    double (&s1)[SBM_N], double (&s2)[SBM_N], double (&s3)[SBM_N],
                                                    double (&s4)[SBM_N])
{
#pragma unroll
    for (int i = 0; i < SBM_N; ++i) sbm_ct_update<S>(f[i], h, y[i], s1[i], s2[i], s3[i], s4[i]);
}

template <int W, int C, int CS, int CM>
__device__ __forceinline__ void sbm_ct_state_role(const SbmSolveArgs& a, const int lane, double* __restrict__ Jb,
                                                  double* __restrict__ Gb, double* __restrict__ sqb,
                                                  double* __restrict__ hb, double* __restrict__ pb,
                                                  long long* __restrict__ trj, int* __restrict__ fl,
                                                  int* __restrict__ ktb, int* __restrict__ cta, double* __restrict__ V)
{
    constexpr int N = SBM_N, NT = W * 32, CSD = (CS > 0) ? CS : 1;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int warp = W - 1;
    constexpr int col0 = (W - 1) * C;
    const long long ntraj = a.B * (long long)a.E;
    const double rtol = a.rtol, atol = a.atol;
    const double inv_naug = 1.0 / (double)(N * (SBM_NF + 1));
    unsigned long long* queue = a.counters + SBM_CNT_QUEUE;

    double y[N], s1[N], s2[N], s3[N], s4[N];
    for (int m = 0; m < SBM_NP; ++m) pb[m * 32 + lane] = 1.0;   // parameters live in shared memory ([param][lane])
#pragma unroll
    for (int i = 0; i < N; ++i) { y[i] = 0.0; s1[i] = 0.0; s2[i] = 0.0; s3[i] = 0.0; s4[i] = 0.0; }
    // this warp's own sensitivity columns
    double* V1 = V + ((warp * 2 + 0) * CM * N) * 32 + lane;
    double* V4 = V + ((warp * 2 + 1) * CM * N) * 32 + lane;
    double ys[CSD][N], ss2[CSD][N], ss3[CSD][N];
#pragma unroll
    for (int c = 0; c < CSD; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ys[c][i] = 0.0; ss2[c][i] = 0.0; ss3[c][i] = 0.0;
            if (CS > 0) { V1[(c * N + i) * 32] = 0.0; V4[(c * N + i) * 32] = 0.0; }
        }

    double t = 0.0, h = 0.0, hstep = 0.0, hprop = 0.0, tk = 0.0, tend = 1.0;
    float facold = 1e-4f;
    int kt = 0, kt1 = 0, status = SBM_OK, mode = SBM_CT_IDLE, e = 0;
    long long traj = -1, b = 0;
    bool lastrej = false, clipped = false, refresh = false, initf = false, exhausted = false;
    unsigned nstep_traj = 0;
    unsigned long long tot_steps = 0, tot_rej = 0, tot_rhs = 0, tot_traj = 0;

    for (;;) {
        // =========================== control (per lane = per trajectory) ===========================
        int info = 0;
        int ktemit = 0;
        if (mode != SBM_CT_IDLE) {
            double sq = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) sq += sqb[(0 * W + w) * 32 + lane];
            const double err = sqrt(sq * inv_naug);
            if (mode == SBM_CT_NORMAL) {
                nstep_traj++;
                tot_steps++;
                tot_rhs += 6;
                if (err <= 1.0) {
                    // accept; PI controller (Hairer's DOPRI5: beta = 0.04, safe = 0.9, factor in [0.2, 10])
                    const float ferr = fmaxf((float)err, 1e-30f);
                    float fac = __powf(ferr, 0.2f - 0.04f * 0.75f) / __powf(facold, 0.04f);
                    fac = fmaxf(0.1f, fminf(5.0f, fac / 0.9f));
                    facold = fmaxf(ferr, 1e-4f);
                    t = clipped ? tk : t + h;
#pragma unroll
                    for (int i = 0; i < N; ++i) y[i] = s3[i];
                    double hnew = h / (double)fac;
                    if (lastrej) hnew = fmin(hnew, h);
                    if (clipped) hnew = fmax(hnew, hprop);     // a clipped step must not shrink the cruise step
                    h = hnew;
                    lastrej = false;
                    info |= SBM_CT_F_ACCEPT;
                } else {
                    tot_rej++;
                    tot_rhs++;                                   // k1 has to be re-evaluated
                    lastrej = true;
                    refresh = true;
                    if (err == err && err < 1e300) {
                        const float ferr = (float)fmin(err, 1e30);
                        h = h / (double)fminf(5.0f, __powf(ferr, 0.2f - 0.04f * 0.75f) / 0.9f);
                    } else {
                        h *= 0.1;                                // NaN/Inf in the trial state: retreat hard
                        if (h < 1e-300) status = SBM_ERR_NONFINITE;
                    }
                }
            } else {
                // a zero-length step just re-evaluated k1 = F(y)
                refresh = false;
                if (initf) {
                    // initial step size (Hairer, Norsett, Wanner II.4): h0 = 0.01 |y|/|f|; start at the rule's
                    // upper bound 100 h0 and let the controller (and, if needed, one rejection) settle it
                    double q0 = 0.0, q1 = 0.0;
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        q0 += sqb[(1 * W + w) * 32 + lane];
                        q1 += sqb[(2 * W + w) * 32 + lane];
                    }
                    const double n0 = sqrt(q0 * inv_naug), n1 = sqrt(q1 * inv_naug);
                    const double h0 = (n0 < 1e-5 || n1 < 1e-5) ? 1e-6 : 0.01 * n0 / n1;
                    h = fmin(100.0 * h0, tend);
                    if (a.h0 > 0.0) h = fmin(a.h0, tend);
                    initf = false;
                    tot_rhs++;
                }
            }
            // arrived at an output time (or failed: flush the remaining rows as NaN, one time per pass)
            if (!refresh && (t >= tk || status != SBM_OK)) {
                info |= SBM_CT_F_EMIT;
                if (status != SBM_OK) info |= SBM_CT_F_POISON;
                ktemit = kt;
                const int r0 = a.orow_off[kt], r1 = a.orow_off[kt + 1];
                for (int r = r0; r < r1; ++r) {
                    const unsigned long long mask = a.row_mask[r];
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < N; ++i) acc += ((mask >> i) & 1ull) ? y[i] : 0.0;   // utils.py:10-26, 48-68
                    if (status != SBM_OK) acc = __longlong_as_double(0x7ff8000000000000ll);
                    a.sims[b * a.n_rows + a.row_sim[r]] = acc;
                }
                kt++;
                if (kt >= kt1) {
                    a.status[traj] = status;
                    tot_traj++;
                    mode = SBM_CT_IDLE;
                } else {
                    tk = a.tout[kt];
                }
            }
        } else if (!exhausted) {
            // ---- fetch a new trajectory (warp-aggregated) ----
            const unsigned want = __activemask();
            const int leader = __ffs(want) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(queue, (unsigned long long)__popc(want));
            base = __shfl_sync(want, base, leader);
            const long long nt = (long long)base + __popc(want & ((1u << lane) - 1u));
            if (nt < ntraj) {
                traj = nt;
                e = (int)(traj % a.E);
                b = traj / a.E;
                for (int m = 0; m < SBM_NP; ++m) {
                    const int c = a.pmap[e * SBM_NP + m];        // base_project.py:343-363
                    pb[m * 32 + lane] = (c >= 0) ? exp(a.theta[b * a.theta_sb + c * a.theta_sp]) : a.pfix[e * SBM_NP + m];
                }
#pragma unroll
                for (int i = 0; i < N; ++i) y[i] = a.y0[e * N + i];
                kt = a.tout_off[e];
                kt1 = a.tout_off[e + 1];
                trj[lane] = traj;
                t = 0.0;
                status = SBM_OK;
                facold = 1e-4f;
                lastrej = false;
                nstep_traj = 0;
                if (kt1 > kt) {
                    tk = a.tout[kt];
                    tend = a.tout[kt1 - 1];
                    mode = SBM_CT_ZERO;
                    refresh = true;
                    initf = true;
                    info |= SBM_CT_F_RESET;
                } else {
                    a.status[traj] = SBM_OK;                   // no output times: nothing to integrate
                    tot_traj++;
                }
            } else {
                exhausted = true;
            }
        }
        // ---- next step of this lane ----
        hstep = 0.0;
        if (mode != SBM_CT_IDLE) {
            mode = SBM_CT_ZERO;
            if (!refresh && status == SBM_OK) {
                hprop = h;
                clipped = false;
                if (t + 1.01 * h >= tk) { h = tk - t; clipped = true; }
                if (nstep_traj >= (unsigned)a.max_steps) status = SBM_ERR_MAXSTEPS;
                else if (!clipped && !(h > 1e-15 * fmax(t, 1e-300))) status = SBM_ERR_UNDERFLOW;
                if (status == SBM_OK) { hstep = h; mode = SBM_CT_NORMAL; }
            }
        }
        hb[lane] = hstep;
        fl[lane] = info;
        ktb[lane] = ktemit;
        const bool alive = __any_sync(FULL, mode != SBM_CT_IDLE) || !__all_sync(FULL, exhausted);
        const bool anyemit = __any_sync(FULL, (info & SBM_CT_F_EMIT) != 0);
        const bool anyinit = __any_sync(FULL, initf);
        const int cflags = (alive ? SBM_CT_ALIVE : 0) | (anyemit ? SBM_CT_ANYEMIT : 0) | (anyinit ? SBM_CT_ANYINIT : 0);
        if (lane == 0) cta[0] = cflags;
        if (!alive) {
            sbm_ct_bar(NT);
            break;
        }
        if (hstep == 0.0) {
#pragma unroll
            for (int i = 0; i < N; ++i) s1[i] = 0.0;   // no stale k1 in a zero-length step (see sbm_ct_cols_begin)
        }
        if (CS > 0) {
            sbm_ct_cols_begin<CSD>(a, info, cflags, hstep, col0, lane, ys, ss3, V1, pb, trj, ktb);
        }
        double sqs = 0.0, d0s = 0.0, d1s = 0.0;

        // ============ stages: the model at stage s (one ahead), own sensitivity columns at stage s-1 ============
#pragma unroll 1
        for (int s = 2; s <= 7; ++s) {
            {
                double Y[N], f[N], Jv[SBM_NJ], Gv[SBM_NG];
                double cs;
                switch (s) {
                case 2: sbm_ct_state_inputs<2>(hstep, y, s1, s2, s3, Y); cs = DP_C2; break;
                case 3: sbm_ct_state_inputs<3>(hstep, y, s1, s2, s3, Y); cs = DP_C3; break;
                case 4: sbm_ct_state_inputs<4>(hstep, y, s1, s2, s3, Y); cs = DP_C4; break;
                case 5: sbm_ct_state_inputs<5>(hstep, y, s1, s2, s3, Y); cs = DP_C5; break;
                case 6: sbm_ct_state_inputs<6>(hstep, y, s1, s2, s3, Y); cs = 1.0; break;
                default: sbm_ct_state_inputs<7>(hstep, y, s1, s2, s3, Y); cs = 1.0; break;
                }
                sbm_fJG<32>(fma(cs, hstep, t), Y, pb + lane, f, Jv, Gv);
                double* Jo = Jb + (s & 1) * SBM_NJ * 32 + lane;
                double* Go = Gb + (s & 1) * SBM_NG * 32 + lane;
#pragma unroll
                for (int q = 0; q < SBM_NJ; ++q) Jo[q * 32] = Jv[q];
#pragma unroll
                for (int q = 0; q < SBM_NG; ++q) Go[q * 32] = Gv[q];
                switch (s) {
                case 2: sbm_ct_state_update<2>(hstep, f, y, s1, s2, s3, s4); break;
                case 3: sbm_ct_state_update<3>(hstep, f, y, s1, s2, s3, s4); break;
                case 4: sbm_ct_state_update<4>(hstep, f, y, s1, s2, s3, s4); break;
                case 5: sbm_ct_state_update<5>(hstep, f, y, s1, s2, s3, s4); break;
                case 6: sbm_ct_state_update<6>(hstep, f, y, s1, s2, s3, s4); break;
                default: {
                    sbm_ct_state_update<7>(hstep, f, y, s1, s2, s3, s4);
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double isk = sbm_rcp_fast(fma(rtol, fmax(fabs(y[i]), fabs(s3[i])), atol));
                        const double q = s4[i] * isk, q0 = y[i] * isk, q1 = f[i] * isk;
                        sqs = fma(q, q, sqs);
                        d0s = fma(q0, q0, d0s);
                        d1s = fma(q1, q1, d1s);
                    }
                } break;
                }
            }
            if (CS > 0 && s >= 3)
                sbm_ct_cols_stage<CSD>(s - 1, hstep, Jb, Gb, lane, col0, ys, ss2, ss3, V1, V4, rtol, atol, anyinit, sqs, d0s, d1s);
            sbm_ct_bar(NT);
        }
        if (CS > 0)
            sbm_ct_cols_stage<CSD>(7, hstep, Jb, Gb, lane, col0, ys, ss2, ss3, V1, V4, rtol, atol, anyinit, sqs, d0s, d1s);
        sqb[(0 * W + warp) * 32 + lane] = sqs;
        sqb[(1 * W + warp) * 32 + lane] = d0s;
        sqb[(2 * W + warp) * 32 + lane] = d1s;
        sbm_ct_bar(NT);                            // every warp has finished stage 7
    }

    // ---- counters: one atomic per warp ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tot_steps += __shfl_xor_sync(FULL, tot_steps, o);
        tot_rej += __shfl_xor_sync(FULL, tot_rej, o);
        tot_rhs += __shfl_xor_sync(FULL, tot_rhs, o);
        tot_traj += __shfl_xor_sync(FULL, tot_traj, o);
    }
    if (lane == 0) {
        atomicAdd(a.counters + SBM_CNT_STEPS, tot_steps);
        atomicAdd(a.counters + SBM_CNT_REJECT, tot_rej);
        atomicAdd(a.counters + SBM_CNT_RHS, tot_rhs);
        atomicAdd(a.counters + SBM_CNT_TRAJ, tot_traj);
    }
}

// W warps per CTA: warps 0..W-2 carry C sensitivity columns each, warp W-1 the state column, the model and
// the last CS columns ((W-1)*C + CS >= NF).  Shared-memory carve-up in doubles (sbm_api.cu computes the
// same size on the host), CM = max(C, CS):
//   2*NJ*32 + 2*NG*32 + 3*W*32 + 32 + NP*32 + 32 + 48 + W*2*CM*N*32
template <int W, int C, int CS>
__device__ __forceinline__ void sbm_dp5_ct_impl(const SbmSolveArgs& a)
{
    constexpr int CM = (C > CS) ? C : CS;
    extern __shared__ double sbm_dyn[];
    double* Jb = sbm_dyn;                                   // [2][NJ][32]
    double* Gb = Jb + 2 * SBM_NJ * 32;                      // [2][NG][32]
    double* sqb = Gb + 2 * SBM_NG * 32;                     // [3][W][32]    error / |y| / |f| partial sums
    double* hb = sqb + 3 * W * 32;                          // [32]          step of each lane (0: idle)
    double* pb = hb + 32;                                   // [NP][32]      model parameters (chain rule)
    long long* trj = (long long*)(pb + SBM_NP * 32);        // [32]          trajectory id
    int* fl = (int*)(trj + 32);                             // [32]          per-lane flags
    int* ktb = fl + 32;                                     // [32]          output-time index to emit
    int* cta = ktb + 32;                                    // [32]          CTA-wide flags in cta[0]
    double* V = (double*)(cta + 32);                        // [W][2][CM*N][32]  S1 and S4 of the sensitivity columns

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (warp == W - 1) sbm_ct_state_role<W, C, CS, CM>(a, lane, Jb, Gb, sqb, hb, pb, trj, fl, ktb, cta, V);
    else if (W > 1) sbm_ct_sens_role<W * 32, (W > 1 ? C : 1), CM>(a, warp, W, lane, Jb, Gb, sqb, hb, pb, trj, fl, ktb, cta, V);
}

#define SBM_DEFINE_DP5_CT(NAME, W, C, CS)                                                               \
    extern "C" __global__ void __launch_bounds__((W) * 32, 1) NAME(const SbmSolveArgs a)                \
    { sbm_dp5_ct_impl<W, C, CS>(a); }
