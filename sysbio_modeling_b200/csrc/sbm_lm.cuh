// sbm_lm.cuh - batched Levenberg-Marquardt step, compiled statically into libsbm_b200.so.
//
// The reference fits a Project with scipy.optimize.leastsq(proj.residuals, x0, Dfun=proj.calc_project_jacobian)
// (tests/test_Project.py:202-213, 351-357): MINPACK's LM on ONE parameter vector, every trial a host round trip.  Here the
// residuals r [n_rows] and the scaled Jacobian J [n_rows][n_proj] of B parameter sets are already in HBM (sbm_solve +
// sbm_loss), so the step is taken there for all sets at once:
//
//     A = J^T J,  g = J^T r,  D_j = max(D_j of the previous iterations, A_jj)   (1 where still zero),
//     (A + lambda D) delta = -g   by Cholesky,      pred = 0.5 delta^T (lambda D delta - g)
//
// D is Marquardt's scaling with MINPACK's running maximum (lmder: diag(j) = max(diag(j), column norm)): without it a
// parameter that drifts into a flat direction loses its damping together with its column norm and the step explodes
// (the Michaelis-Menten 'sum' fixture of the reference is not identifiable and does exactly that).
//
// pred is the reduction of 0.5 |r|^2 predicted by the linear model (what the gain ratio divides by).
//
// One warp per parameter set (n_proj <= 32): every row of J is read once, coalesced, into a small shared-memory tile; the
// lower triangle of A is spread over the lanes in registers; the factorisation runs in shared memory (n_proj^2 doubles per
// warp).  HBM-bound in principle (8 (n_rows (n_proj + 1) + 1) bytes in, 8 (n_proj + 2) + 4 out per set); at C4's 40 x 22 the
// shared loads + FMAs per entry and the dependent steps of the factorisation make it issue-bound instead.
#ifndef SBM_LM_CUH
#define SBM_LM_CUH

#include <cuda_runtime.h>

struct SbmLmArgs {
    const double* jac;      // [B][n_rows][n_proj]
    const double* resid;    // [B][n_rows]
    const double* lambda;   // [B] damping of each parameter set
    double* dscale;         // [B][n_proj] in/out: running maximum of diag(J^T J) (NULL: this iteration's diagonal only)
    double* delta;          // [B][n_proj] step (0 where the system is not positive definite or not finite)
    double* pred;           // [B] predicted reduction of 0.5 |r|^2
    double* gmax;           // [B] max_j |g_j|
    int* info;              // [B] 0 ok, 1 not positive definite / non-finite input
    long long B;
    int n_rows, n_proj;
};

#define SBM_LM_WARPS 4
#define SBM_LM_RC 8            /* rows of J staged per pass */

// One warp per parameter set.  The lower triangle of A = J^T J (n_proj (n_proj + 1) / 2 entries) is spread over the 32 lanes,
// EPL entries each in registers; J goes through a shared-memory tile of SBM_LM_RC rows (coalesced loads, then two
// shared loads + one FMA per entry and row).  Cholesky is left-looking with lane = row: L_ik = (A_ik - sum_j L_ij L_kj) / L_kk,
// the reciprocal square root of the pivot broadcast by a shuffle.
template <int EPL>
__global__ void __launch_bounds__(SBM_LM_WARPS * 32) sbm_lm_step_kernel(const SbmLmArgs a)
{
    extern __shared__ double lm_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = a.n_proj, R = a.n_rows, LD = P + 1;           // LD: odd stride against bank conflicts
    const int per_warp = LD * P + 4 * P + SBM_LM_RC * (P + 1);
    double* M = lm_dyn + (size_t)warp * per_warp;               // [P][LD] lower triangle of A + lambda D, then L
    double* gs = M + LD * P;                                    // [P] -g, then the right-hand side / solution
    double* ds = gs + P;                                        // [P] D
    double* xs = ds + P;                                        // [P] delta
    double* iv = xs + P;                                        // [P] 1 / L_kk (the solves multiply)
    double* tile = iv + P;                                      // [RC][P] rows of J
    double* rt = tile + SBM_LM_RC * P;                          // [RC] residuals of those rows
    constexpr unsigned FULL = 0xffffffffu;
    const int NE = P * (P + 1) / 2;
    // entry idx = lane + 32 e of the lower triangle, row-major: (i, j), j <= i
    int ei[EPL], ej[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int idx = lane + 32 * e;
        int i = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while (i * (i + 1) / 2 > idx) --i;
        while ((i + 1) * (i + 2) / 2 <= idx) ++i;
        ei[e] = idx < NE ? i : 0;
        ej[e] = idx < NE ? idx - i * (i + 1) / 2 : 0;
    }
    for (long long b = (long long)blockIdx.x * SBM_LM_WARPS + warp; b < a.B; b += (long long)gridDim.x * SBM_LM_WARPS) {
        const double* __restrict__ J = a.jac + b * (long long)R * P;
        const double* __restrict__ r = a.resid + b * (long long)R;
        double acc[EPL], g = 0.0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = 0.0;
        for (int q0 = 0; q0 < R; q0 += SBM_LM_RC) {
            const int nq = min(SBM_LM_RC, R - q0);
            __syncwarp();
            for (int t = lane; t < nq * P; t += 32) tile[t] = J[(long long)q0 * P + t];
            if (lane < nq) rt[lane] = r[q0 + lane];
            __syncwarp();
            for (int q = 0; q < nq; ++q) {
                const double* __restrict__ row = tile + q * P;
#pragma unroll
                for (int e = 0; e < EPL; ++e) acc[e] = fma(row[ei[e]], row[ej[e]], acc[e]);
                if (lane < P) g = fma(row[lane], rt[q], g);
            }
        }
        const double lam = a.lambda[b];
        bool finite = (g == g) && (fabs(g) < 1e300) && (lam == lam);
#pragma unroll
        for (int e = 0; e < EPL; ++e)
            if (lane + 32 * e < NE) {
                finite = finite && (acc[e] == acc[e]) && (fabs(acc[e]) < 1e300);
                M[ei[e] * LD + ej[e]] = acc[e];
            }
        finite = __all_sync(FULL, finite);
        __syncwarp();
        // ---- D_j = max(running maximum, A_jj), 1 where still zero; M = A + lambda D ----
        if (lane < P) {
            double dj = M[lane * LD + lane];
            const double ajj = dj;
            if (a.dscale) {
                const double dprev = a.dscale[b * P + lane];
                if (dprev > dj) dj = dprev;
                if (dj == dj && dj < 1e300) a.dscale[b * P + lane] = dj;
            }
            if (!(dj > 0.0)) dj = 1.0;                            // a parameter without influence: plain Levenberg damping
            M[lane * LD + lane] = ajj + lam * dj;
            gs[lane] = -g;
            ds[lane] = dj;
        }
        double gm = lane < P ? fabs(g) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(FULL, gm, o));
        __syncwarp();
        // ---- Cholesky M = L L^T in place, left-looking, lane = row ----
        bool ok = finite;
        for (int k = 0; k < P; ++k) {
            double sden = 0.0;
            if (lane >= k && lane < P) {
                sden = M[lane * LD + k];
                for (int j = 0; j < k; ++j) sden = fma(-M[lane * LD + j], M[k * LD + j], sden);
            }
            const double d = __shfl_sync(FULL, sden, k);
            ok = ok && d > 0.0 && d < 1e300;
            const double rs = rsqrt(d);
            if (lane == k) { M[k * LD + k] = d * rs; iv[k] = rs; }
            if (lane > k && lane < P) M[lane * LD + k] = sden * rs;
            __syncwarp();
        }
        double pr = 0.0;
        if (ok) {
            // ---- L y = -g, L^T delta = y ----
            for (int k = 0; k < P; ++k) {
                const double yk = gs[k] * iv[k];
                __syncwarp();
                if (lane == k) gs[k] = yk;
                if (lane > k && lane < P) gs[lane] -= M[lane * LD + k] * yk;
                __syncwarp();
            }
            for (int k = P - 1; k >= 0; --k) {
                const double xk = gs[k] * iv[k];
                __syncwarp();
                if (lane == k) xs[k] = xk;
                if (lane < k) gs[lane] -= M[k * LD + lane] * xk;
                __syncwarp();
            }
            if (lane < P) {
                const double d = xs[lane];
                ok = (d == d) && fabs(d) < 1e300;
                pr = 0.5 * d * (lam * ds[lane] * d - g);          // g_j is still in the register
            }
            ok = __all_sync(FULL, ok);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pr += __shfl_xor_sync(FULL, pr, o);
        }
        if (lane < P) a.delta[b * P + lane] = ok ? xs[lane] : 0.0;
        if (lane == 0) {
            a.pred[b] = ok ? pr : 0.0;
            a.gmax[b] = gm;
            a.info[b] = ok ? 0 : 1;
        }
        __syncwarp();
    }
}

static inline int sbm_launch_lm_step(const SbmLmArgs& a, cudaStream_t st)
{
    const int P = a.n_proj;
    if (P > 32) return -1;
    const size_t smem = (size_t)SBM_LM_WARPS * ((size_t)(P + 1) * P + 4 * P + SBM_LM_RC * (P + 1)) * sizeof(double);
    long long blocks = (a.B + SBM_LM_WARPS - 1) / SBM_LM_WARPS;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const int epl = (P * (P + 1) / 2 + 31) / 32;                // 1 .. 17
    if (epl <= 2) sbm_lm_step_kernel<2><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    else if (epl <= 5) sbm_lm_step_kernel<5><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    else if (epl <= 8) sbm_lm_step_kernel<8><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    else if (epl <= 12) sbm_lm_step_kernel<12><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    else sbm_lm_step_kernel<17><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    return 0;
}

#endif
