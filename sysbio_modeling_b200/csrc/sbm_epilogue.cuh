// sbm_epilogue.cuh - model-independent device code compiled statically into libsbm_b200.so:
//   * the loss epilogue: optimal scale factors and their gradients, residuals, SSR, scaled Jacobian
//     and J^T r  (reference project/loss_functions/**; formulas cited at each step);
//   * the DFMA-throughput microbenchmark that gives the measured FP64 roofline denominator.
//
// One CTA per parameter set: a scale factor couples every row of its group across all experiments of
// that parameter set, so the reduction cannot live inside the per-trajectory integrator.  The pass is
// HBM-bound (reads the model Jacobian twice, writes it once); its traffic is stated in DESIGN.md.
#ifndef SBM_EPILOGUE_CUH
#define SBM_EPILOGUE_CUH

#include <cuda_runtime.h>
#include "../../include/sbm_b200.h"

struct SbmLossArgs {
    const double* data;
    const double* sigma;
    const int* row_group;
    const int* row_kind;
    const int* row_aux;
    const int* grp_off;
    const int* grp_rows;
    const double* theta;
    long long theta_sb, theta_sp;
    double* sims;
    double* jac;
    double* resid;
    double* ssr;
    double* sf;
    double* sfgrad;
    double* grad;
    long long B;
    int n_rows, n_proj, n_groups, loss_kind, flags;
};

#define SBM_LOSS_THREADS 128

__device__ __forceinline__ double sbm_block_sum(double v, double* red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    __syncthreads();                       // protect red[] from the previous use
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < SBM_LOSS_THREADS / 32; ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(SBM_LOSS_THREADS) sbm_loss_kernel(const SbmLossArgs a)
{
    extern __shared__ double dyn[];        // [3][n_groups]: sf, aux1 (sde | inv), aux2 (sds)
    __shared__ double red[SBM_LOSS_THREADS / 32];
    __shared__ int s_bad;
    double* s_sf = dyn;
    double* s_a1 = dyn + a.n_groups;
    double* s_a2 = dyn + 2 * a.n_groups;
    const int tid = threadIdx.x;
    const bool is_log = (a.loss_kind == SBM_LOSS_LOG);

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        double* sims = a.sims + b * a.n_rows;
        double* resid = a.resid + b * a.n_rows;
        double* jac = a.jac ? a.jac + b * (long long)a.n_rows * a.n_proj : nullptr;
        if (tid == 0) s_bad = 0;
        __syncthreads();

        // ---- prior rows carry theta itself (base_project.py:427-437); NaN scan (squared_loss_function.py:28-32)
        int bad = 0;
        for (int r = tid; r < a.n_rows; r += SBM_LOSS_THREADS) {
            const int kind = a.row_kind[r];
            if (kind == SBM_ROW_PARAM_PRIOR) sims[r] = a.theta[b * a.theta_sb + a.row_aux[r] * a.theta_sp];
            if (kind != SBM_ROW_SF_PRIOR) { const double s = sims[r]; if (s != s) bad = 1; }
        }
        if (bad) s_bad = 1;
        __syncthreads();

        // ---- optimal scale factor per group
        for (int g = 0; g < a.n_groups; ++g) {
            double p1 = 0.0, p2 = 0.0, p3 = 0.0;
            for (int q = a.grp_off[g] + tid; q < a.grp_off[g + 1]; q += SBM_LOSS_THREADS) {
                const int r = a.grp_rows[q];
                const double s = sims[r], d = a.data[r], sg = a.sigma[r];
                if (!is_log) {
                    // linear_scale_factor.py:27-31   B = sum(s d / sigma^2) / sum(s^2 / sigma^2)
                    const double w = 1.0 / (sg * sg);
                    p1 += s * d * w;
                    p2 += s * s * w;
                } else {
                    // log_scale_factor.py:17-27 with sigma rescaled by the data (s_i = sigma_i / d_i)
                    const double sr = sg / d;
                    const double w = 1.0 / (sr * sr);
                    p1 += w;
                    p2 += log(d) * w;
                    p3 += log(s) * w;
                }
            }
            p1 = sbm_block_sum(p1, red);
            p2 = sbm_block_sum(p2, red);
            if (is_log) p3 = sbm_block_sum(p3, red);
            if (tid == 0) {
                double B;
                if (!is_log) { B = p1 / p2; s_a1[g] = p1; s_a2[g] = p2; }
                else { B = exp(p2 / p1) * exp(-p3 / p1); s_a1[g] = p1; s_a2[g] = 0.0; }
                s_sf[g] = B;
                a.sf[b * a.n_groups + g] = B;
            }
            __syncthreads();
        }

        // ---- gradient of the scale factors w.r.t. theta (needs the unscaled model Jacobian)
        if (jac) {
            for (int g = 0; g < a.n_groups; ++g) {
                double* out = a.sfgrad + (b * a.n_groups + g) * (long long)a.n_proj;
                for (int c = tid; c < a.n_proj; c += SBM_LOSS_THREADS) {
                    double j1 = 0.0, j2 = 0.0;
                    for (int q = a.grp_off[g]; q < a.grp_off[g + 1]; ++q) {
                        const int r = a.grp_rows[q];
                        const double J = jac[(long long)r * a.n_proj + c];
                        const double s = sims[r], d = a.data[r], sg = a.sigma[r];
                        if (!is_log) {
                            const double w = 1.0 / (sg * sg);
                            j1 += J * d * w;           // jac_dot_exp
                            j2 += J * s * w;           // jac_dot_sim
                        } else {
                            const double sr = sg / d;
                            j1 += J / (s * sr * sr);
                        }
                    }
                    double gB;
                    if (!is_log) {
                        // linear_scale_factor.py:33-42 (note the factor 2)
                        const double sde = s_a1[g], sds = s_a2[g];
                        gB = j1 / sds - 2.0 * sde * j2 / (sds * sds);
                    } else {
                        // log_scale_factor.py:29-36
                        gB = s_sf[g] * (-1.0 / s_a1[g]) * j1;
                    }
                    out[c] = gB;
                }
            }
            __syncthreads();
        }

        // ---- residuals (squared_loss_function.py:34-40, log_squared_loss_function.py:23-64)
        const bool allbad = (s_bad != 0);
        double acc = 0.0;
        for (int r = tid; r < a.n_rows; r += SBM_LOSS_THREADS) {
            const int kind = a.row_kind[r];
            double s = sims[r];
            double res;
            if (kind == SBM_ROW_MEASUREMENT) {
                const int g = a.row_group[r];
                if (g >= 0) s *= s_sf[g];
                res = is_log ? (log(s) - log(a.data[r])) / a.sigma[r] : (s - a.data[r]) / a.sigma[r];
            } else {
                if (kind == SBM_ROW_SF_PRIOR) { s = log(s_sf[a.row_aux[r]]); sims[r] = s; }   // abstract_loss_function.py:93-105
                res = (s - a.data[r]) / a.sigma[r];
            }
            if (allbad) res = __longlong_as_double(0x7ff0000000000000ll);
            resid[r] = res;
            acc += res * res;
        }
        acc = sbm_block_sum(acc, red);
        if (tid == 0) a.ssr[b] = 0.5 * acc;            // base_project.py:807-827
        __syncthreads();

        // ---- scaled Jacobian  B*J + sim*dB  (squared_loss_function.py:58-80) and  J^T r  (base_project.py:805)
        if (jac) {
            for (int c = tid; c < a.n_proj; c += SBM_LOSS_THREADS) {
                double gacc = 0.0;
                for (int r = 0; r < a.n_rows; ++r) {
                    const int kind = a.row_kind[r];
                    double J = jac[(long long)r * a.n_proj + c];
                    if (kind == SBM_ROW_MEASUREMENT) {
                        const int g = a.row_group[r];
                        const double s = sims[r];
                        if (g >= 0) {
                            const double gB = a.sfgrad[(b * a.n_groups + g) * (long long)a.n_proj + c];
                            J = is_log ? (J / s + gB / s_sf[g]) : (J * s_sf[g] + s * gB);
                        } else if (is_log) {
                            J = J / s;
                        }
                    } else if (kind == SBM_ROW_SF_PRIOR) {
                        const int g = a.row_aux[r];     // abstract_loss_function.py:107-120, linear_scale_factor.py:44-53
                        J = a.sfgrad[(b * a.n_groups + g) * (long long)a.n_proj + c] / s_sf[g];
                    } else if (kind == SBM_ROW_PARAM_PRIOR && (a.flags & SBM_FLAG_PRIOR_JAC_UNIT)) {
                        J = (a.row_aux[r] == c) ? 1.0 : 0.0;
                    }
                    if (allbad) J = __longlong_as_double(0x7ff0000000000000ll);
                    jac[(long long)r * a.n_proj + c] = J;
                    gacc += J * resid[r];
                }
                if (a.grad) a.grad[b * a.n_proj + c] = gacc;
            }
        }
        __syncthreads();
    }
}

static inline void sbm_launch_loss(const SbmLossArgs& a, cudaStream_t st)
{
    long long blocks = a.B < 1048576 ? a.B : 1048576;
    size_t smem = (size_t)(3 * (a.n_groups > 0 ? a.n_groups : 1)) * sizeof(double);
    sbm_loss_kernel<<<(unsigned)blocks, SBM_LOSS_THREADS, smem, st>>>(a);
}

// ------------------------------------------------------------------------------------------------
// DFMA throughput: 16 independent FMA chains per thread, 4 CTAs of 256 threads per SM
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sbm_fp64_peak_kernel(double* out, int iters, double seed)
{
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = seed + (double)(threadIdx.x + i);
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma(x[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[0] = s;       // never true; keeps the chains alive
}

static inline cudaError_t sbm_run_fp64_peak(int iters, double* tflops, double* ms_out)
{
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    double* d = nullptr;
    if ((e = cudaMalloc(&d, 8)) != cudaSuccess) return e;
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    sbm_fp64_peak_kernel<<<blocks, threads>>>(d, iters / 8 + 1, 1.0);      // warm-up
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        sbm_fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    *ms_out = best;
    return cudaGetLastError();
}

#endif
