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
// One warp per parameter set, lane j owns column j of A (n_proj <= 32): every row of J is read once, coalesced, and
// its entries are exchanged by shuffles - J never goes through shared memory.  The factorisation runs in shared memory
// (n_proj^2 doubles per warp).  HBM-bound in principle (8 (n_rows (n_proj + 1) + 1) bytes in, 8 (n_proj + 2) + 4 out per
// set); at C4's 40 x 22 the n_proj shuffles + FMAs per row make it issue-bound instead.
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

template <int PMAX>
__global__ void __launch_bounds__(SBM_LM_WARPS * 32) sbm_lm_step_kernel(const SbmLmArgs a)
{
    extern __shared__ double lm_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = a.n_proj, R = a.n_rows, LD = P + 1;           // LD: odd stride against bank conflicts
    double* M = lm_dyn + (size_t)warp * (LD * P + 4 * P);        // [P][LD] lower triangle of A + lambda D, then L
    double* gs = M + LD * P;                                    // [P] g, then the right-hand side / solution
    double* ds = gs + P;                                        // [P] D
    double* xs = ds + P;                                        // [P] delta
    double* iv = xs + P;                                        // [P] 1 / L_kk (the solves multiply)
    constexpr unsigned FULL = 0xffffffffu;
    for (long long b = (long long)blockIdx.x * SBM_LM_WARPS + warp; b < a.B; b += (long long)gridDim.x * SBM_LM_WARPS) {
        const double* __restrict__ J = a.jac + b * (long long)R * P;
        const double* __restrict__ r = a.resid + b * (long long)R;
        // ---- A = J^T J (column `lane` in registers), g = J^T r ----
        double acc[PMAX], g = 0.0;
#pragma unroll
        for (int i = 0; i < PMAX; ++i) acc[i] = 0.0;
        for (int q = 0; q < R; ++q) {
            const double x = lane < P ? J[q * P + lane] : 0.0;
            const double rq = r[q];
            g = fma(x, rq, g);
#pragma unroll
            for (int i = 0; i < PMAX; ++i) acc[i] = fma(__shfl_sync(FULL, x, i), x, acc[i]);   // A[i][lane]
        }
        const double lam = a.lambda[b];
        // ---- M = A + lambda D (lower triangle: row i >= column lane) ----
        double dj = 0.0;
#pragma unroll
        for (int i = 0; i < PMAX; ++i) if (i == lane) dj = acc[i];
        if (a.dscale && lane < P) {
            const double dprev = a.dscale[b * P + lane];
            if (dprev > dj) dj = dprev;
            if (dj == dj && dj < 1e300) a.dscale[b * P + lane] = dj;
        }
        if (!(dj > 0.0)) dj = 1.0;                                // a parameter without influence: plain Levenberg damping
        bool finite = (g == g) && (fabs(g) < 1e300) && (lam == lam);
        if (lane < P) {
#pragma unroll
            for (int i = 0; i < PMAX; ++i)
                if (i < P) {
                    finite = finite && (acc[i] == acc[i]) && (fabs(acc[i]) < 1e300);
                    M[i * LD + lane] = acc[i] + (i == lane ? lam * dj : 0.0);
                }
            gs[lane] = -g;
            ds[lane] = dj;
        }
        finite = __all_sync(FULL, finite);
        double gm = lane < P ? fabs(g) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(FULL, gm, o));
        __syncwarp();
        // ---- Cholesky M = L L^T in place (right-looking; lane i owns row i of the trailing update) ----
        bool ok = finite;
        for (int k = 0; k < P && ok; ++k) {
            const double mkk = M[k * LD + k];
            ok = mkk > 0.0 && mkk < 1e300;
            if (!ok) break;
            const double lkk = sqrt(mkk), inv = 1.0 / lkk;
            __syncwarp();
            if (lane == k) { M[k * LD + k] = lkk; iv[k] = inv; }
            if (lane > k && lane < P) M[lane * LD + k] *= inv;
            __syncwarp();
            if (lane > k && lane < P) {
                const double lik = M[lane * LD + k];
                for (int j = k + 1; j <= lane; ++j) M[lane * LD + j] -= lik * M[j * LD + k];
            }
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
    const size_t smem = (size_t)SBM_LM_WARPS * ((size_t)(P + 1) * P + 4 * P) * sizeof(double);
    long long blocks = (a.B + SBM_LM_WARPS - 1) / SBM_LM_WARPS;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (P <= 8) sbm_lm_step_kernel<8><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    else if (P <= 16) sbm_lm_step_kernel<16><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    else if (P <= 24) sbm_lm_step_kernel<24><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    else if (P <= 32) sbm_lm_step_kernel<32><<<(unsigned)blocks, SBM_LM_WARPS * 32, smem, st>>>(a);
    else return -1;
    return 0;
}

#endif
