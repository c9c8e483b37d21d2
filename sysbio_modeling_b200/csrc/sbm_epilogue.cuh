// sbm_epilogue.cuh - model-independent device code compiled statically into libsbm_b200.so:
//   * the loss epilogue: optimal scale factors and their gradients, residuals, SSR, scaled Jacobian
//     and J^T r  (reference project/loss_functions/**; formulas cited at each step);
//   * per-set normal equations J^T J / J^T r (what the optimisers and the Jacobian-block gather consume);
//   * the DFMA-throughput microbenchmark that gives the measured FP64 roofline denominator.
//
// A scale factor couples every row of its group across all experiments of a parameter set, so the reduction
// cannot live inside the per-trajectory integrator.  Two kernels, same arithmetic:
//
//   sbm_loss_tile_kernel  one WARP per parameter set.  The model Jacobian of the set ([n_rows][n_proj], contiguous)
//                         is staged in shared memory by one TMA bulk copy (cp.async.bulk, completion on an mbarrier),
//                         double-buffered per warp; the scaled Jacobian goes back by a TMA bulk store.  HBM traffic:
//                         J read ONCE and written once (the previous kernel read it twice).  Used whenever the tile
//                         fits (C4: 40 x 22 doubles = 7 KB), and for residual-only evaluations (no tile at all).
//   sbm_loss_kernel       one CTA per parameter set, J stays in global memory (two reads, one write): any shape
//   sbm_loss_wide_*       few sets with big Jacobians (C2 through the Project: 65,536 rows x 16,384 project columns per
//                         set): every set spread over the whole grid, six kernels in stream order.
//
// Both can stop after the group sums (SBM_LOSS_PHASE_SUMS) or start from given sums (SBM_LOSS_PHASE_FINISH): that is
// the hook of the experiment shard (distributed.py), which all-reduces the sums over ranks in between
// (abstract_loss_function.py:62-91, linear_scale_factor.py:27-42).
#ifndef SBM_EPILOGUE_CUH
#define SBM_EPILOGUE_CUH

#include <cuda_runtime.h>
#include <cstdlib>
#include "../../include/sbm_b200.h"

#define SBM_LOSS_PHASE_ALL    0
#define SBM_LOSS_PHASE_SUMS   1   /* write sums[b][g][3 + 2 n_proj] and stop                     */
#define SBM_LOSS_PHASE_FINISH 2   /* read the (reduced) sums instead of forming them              */

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
    double* sums;          // [B][n_groups][3 + 2 n_proj]: p1, p2, p3, j1[n_proj], j2[n_proj] (phases SUMS / FINISH)
    long long B;
    int n_rows, n_proj, n_groups, loss_kind, flags, phase;
};

#define SBM_LOSS_THREADS 128
#define SBM_INF_BITS 0x7ff0000000000000ll

// ---- the arithmetic shared by both kernels -----------------------------------------------------------------
// weights of one row of a scale-factor group
//   linear (linear_scale_factor.py:27-42):  w = 1/sigma^2;  p1 += s d w, p2 += s s w;  j1 += J d w, j2 += J s w
//   log    (log_scale_factor.py:17-36), sigma rescaled by the data (sr = sigma / d):  w = 1/sr^2;
//          p1 += w, p2 += log(d) w, p3 += log(s) w;  j1 += J / (s sr^2)
__device__ __forceinline__ void sbm_sf_terms(const bool is_log, const double s, const double d, const double sg,
                                             double& p1, double& p2, double& p3)
{
    if (!is_log) {
        const double w = 1.0 / (sg * sg);
        p1 += s * d * w;
        p2 += s * s * w;
    } else {
        const double sr = sg / d;
        const double w = 1.0 / (sr * sr);
        p1 += w;
        p2 += log(d) * w;
        p3 += log(s) * w;
    }
}

__device__ __forceinline__ void sbm_sfgrad_terms(const bool is_log, const double J, const double s, const double d,
                                                 const double sg, double& j1, double& j2)
{
    if (!is_log) {
        const double w = 1.0 / (sg * sg);
        j1 += J * d * w;           // jac_dot_exp
        j2 += J * s * w;           // jac_dot_sim
    } else {
        const double sr = sg / d;
        j1 += J / (s * sr * sr);
    }
}

__device__ __forceinline__ double sbm_sf_value(const bool is_log, const double p1, const double p2, const double p3)
{
    return is_log ? exp(p2 / p1) * exp(-p3 / p1) : p1 / p2;
}

// dB/dtheta_c: linear_scale_factor.py:33-42 (note the factor 2), log_scale_factor.py:29-36
__device__ __forceinline__ double sbm_sfgrad_value(const bool is_log, const double B, const double p1, const double p2,
                                                   const double j1, const double j2)
{
    return is_log ? B * (-1.0 / p1) * j1 : j1 / p2 - 2.0 * p1 * j2 / (p2 * p2);
}

// scaled Jacobian entry (squared_loss_function.py:58-80, log_squared_loss_function.py:66-98; prior rows:
// abstract_loss_function.py:107-120, linear_scale_factor.py:44-53, base_project.py:519-530)
__device__ __forceinline__ double sbm_scaled_entry(const SbmLossArgs& a, const bool is_log, const int c,
                                                   const int kind, const int g, const int aux, double J, const double s,
                                                   const double* __restrict__ sfv, const double* __restrict__ sfg)
{
    if (kind == SBM_ROW_MEASUREMENT) {
        if (g >= 0) {
            const double gB = sfg[g * a.n_proj + c];
            J = is_log ? (J / s + gB / sfv[g]) : (J * sfv[g] + s * gB);
        } else if (is_log) {
            J = J / s;
        }
    } else if (kind == SBM_ROW_SF_PRIOR) {
        J = sfg[aux * a.n_proj + c] / sfv[aux];
    } else if (kind == SBM_ROW_PARAM_PRIOR && (a.flags & SBM_FLAG_PRIOR_JAC_UNIT)) {
        J = (aux == c) ? 1.0 : 0.0;
    }
    return J;
}

__device__ __forceinline__ double sbm_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- TMA bulk copies (1-D) and mbarriers ----------------------------------------------------------------------
__device__ __forceinline__ unsigned sbm_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sbm_mbar_init(unsigned long long* mb, const int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sbm_smem_addr(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void sbm_mbar_expect_tx(unsigned long long* mb, const unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sbm_smem_addr(mb)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbm_mbar_wait(unsigned long long* mb, const unsigned parity)
{
    const unsigned addr = sbm_smem_addr(mb);
    unsigned ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 28)) __trap();          // a lost copy must not hang the GPU
    }
}
__device__ __forceinline__ void sbm_tma_load(void* dst_smem, const void* src_gmem, const unsigned bytes, unsigned long long* mb)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sbm_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(sbm_smem_addr(mb)) : "memory");
}
__device__ __forceinline__ void sbm_tma_store(void* dst_gmem, const void* src_smem, const unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst_gmem), "r"(sbm_smem_addr(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// =================================================================================================================
// warp-per-set kernel.  Dynamic shared memory: per CTA the row tables of the plan (data, 1/sigma, sigma-derived weights,
// kind | group | aux, rows of every group: constant over the parameter sets, read from global memory once per CTA), then
// per warp (doubles): 2 * tile | sims[n_rows] | resid[n_rows] | w2[n_rows] | sf[G] | p1[G] | p2[G] | sfgrad[G * n_proj];
// then 2 mbarriers per warp.  Everything a row contributes to the column loops is formed once per set by the lanes over
// the rows; the column loops (lane = project column) are shared-memory loads and FMAs only.
// =================================================================================================================
struct SbmLossTileShape {
    int tile_doubles, warp_doubles, cta_doubles, n_grp_rows, warps, nbuf;
    size_t smem;
};

template <bool HAS_JAC>
__global__ void __launch_bounds__(384) sbm_loss_tile_kernel(const SbmLossArgs a, const SbmLossTileShape sp)
{
    extern __shared__ __align__(16) double dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool is_log = (a.loss_kind == SBM_LOSS_LOG);
    const int R = a.n_rows, P = a.n_proj, G = a.n_groups;
    // ---- CTA-wide row tables ----
    double* c_data = dyn;                          // data (log loss: log(data))
    double* c_isig = c_data + R;                   // 1 / sigma
    double* c_wd = c_isig + R;                     // weight of J in j1: d / sigma^2 (log loss: 1 / sr^2, completed by 1 / s)
    double* c_w = c_wd + R;                        // 1 / sigma^2 (log loss: unused)
    int* c_meta = reinterpret_cast<int*>(c_w + R); // kind | (group + 1) << 4 | aux << 16
    int* c_grows = c_meta + R + (R & 1);
    int* c_goff = c_grows + sp.n_grp_rows + (sp.n_grp_rows & 1);
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const double d = a.data[r], sg = a.sigma[r];
        const int kind = a.row_kind[r];
        c_meta[r] = kind | ((a.row_group[r] + 1) << 4) | (a.row_aux[r] << 16);
        c_isig[r] = 1.0 / sg;
        if (!is_log || kind != SBM_ROW_MEASUREMENT) { c_data[r] = d; const double w = 1.0 / (sg * sg); c_w[r] = w; c_wd[r] = d * w; }
        else { c_data[r] = log(d); const double sr = sg / d; c_w[r] = 0.0; c_wd[r] = 1.0 / (sr * sr); }
    }
    for (int q = threadIdx.x; q < sp.n_grp_rows; q += blockDim.x) c_grows[q] = a.grp_rows[q];
    for (int g = threadIdx.x; g <= G; g += blockDim.x) c_goff[g] = a.grp_off[g];
    double* base = dyn + sp.cta_doubles + (size_t)warp * sp.warp_doubles;
    double* tile0 = base;
    double* s_sims = base + sp.nbuf * sp.tile_doubles;
    double* s_res = s_sims + R;
    double* s_w1 = s_res + R;                      // per set: weight of J in j1 (log loss) / unused
    double* s_w2 = s_w1 + R;                       // per set: weight of J in j2: s / sigma^2
    double* s_sf = s_w2 + R;
    double* s_p1 = s_sf + G;
    double* s_p2 = s_p1 + G;
    double* s_sfg = s_p2 + G;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(dyn + sp.cta_doubles + (size_t)nwarps * sp.warp_doubles) + 2 * warp;
    const unsigned tile_bytes = (unsigned)(R * P) * 8u;
    const long long stride = (long long)gridDim.x * nwarps;
    long long b = (long long)blockIdx.x * nwarps + warp;

    if (HAS_JAC) {
        if (lane == 0) { sbm_mbar_init(mbar, 1); sbm_mbar_init(mbar + 1, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
        if (lane == 0 && b < a.B) {
            sbm_mbar_expect_tx(mbar, tile_bytes);
            sbm_tma_load(tile0, a.jac + b * (long long)R * P, tile_bytes, mbar);
        }
    }
    __syncthreads();                               // row tables complete
    unsigned it = 0;
    const int nbuf = sp.nbuf;
    for (; b < a.B; b += stride, ++it) {
        const int buf = (nbuf == 2) ? (it & 1) : 0;
        double* tile = tile0 + buf * sp.tile_doubles;
        double* sims = a.sims + b * R;
        // ---- two buffers: the next set's Jacobian into the other one (its previous contents have been read by the bulk
        //      store).  One buffer (more warps per SM instead: their loads overlap): issued at the end of the iteration ----
        if (HAS_JAC && nbuf == 2 && lane == 0 && b + stride < a.B) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            sbm_mbar_expect_tx(mbar + (buf ^ 1), tile_bytes);
            sbm_tma_load(tile0 + (buf ^ 1) * sp.tile_doubles, a.jac + (b + stride) * (long long)R * P, tile_bytes, mbar + (buf ^ 1));
        }
        // ---- prior rows carry theta itself (base_project.py:427-437); NaN scan (squared_loss_function.py:28-32) ----
        int bad = 0;
        for (int r = lane; r < R; r += 32) {
            const int meta = c_meta[r], kind = meta & 15;
            double s = sims[r];
            if (kind == SBM_ROW_PARAM_PRIOR) { s = a.theta[b * a.theta_sb + (meta >> 16) * a.theta_sp]; sims[r] = s; }
            if (kind != SBM_ROW_SF_PRIOR && s != s) bad = 1;
            s_sims[r] = s;
            if (HAS_JAC) {
                if (!is_log) s_w2[r] = s * c_w[r];
                else s_w1[r] = c_wd[r] / s;
            }
        }
        const bool allbad = __any_sync(0xffffffffu, bad) != 0;
        __syncwarp();
        // ---- optimal scale factor per group ----
        double* sums = a.sums ? a.sums + b * (long long)G * (3 + 2 * P) : nullptr;
        for (int g = 0; g < G; ++g) {
            double p1 = 0.0, p2 = 0.0, p3 = 0.0;
            if (a.phase != SBM_LOSS_PHASE_FINISH) {
                for (int q = c_goff[g] + lane; q < c_goff[g + 1]; q += 32) {
                    const int r = c_grows[q];
                    const double sv = s_sims[r];
                    if (!is_log) {                 // linear_scale_factor.py:27-31
                        p1 = fma(sv, c_wd[r], p1);
                        p2 = fma(sv * sv, c_w[r], p2);
                    } else {                       // log_scale_factor.py:17-27
                        p1 += c_wd[r];
                        p2 = fma(c_data[r], c_wd[r], p2);
                        p3 = fma(log(sv), c_wd[r], p3);
                    }
                }
                p1 = sbm_warp_sum(p1);
                p2 = sbm_warp_sum(p2);
                if (is_log) p3 = sbm_warp_sum(p3);
                if (a.phase == SBM_LOSS_PHASE_SUMS && lane == 0) {
                    sums[g * (3 + 2 * P) + 0] = p1; sums[g * (3 + 2 * P) + 1] = p2; sums[g * (3 + 2 * P) + 2] = p3;
                }
            } else {
                p1 = sums[g * (3 + 2 * P) + 0]; p2 = sums[g * (3 + 2 * P) + 1]; p3 = sums[g * (3 + 2 * P) + 2];
            }
            if (lane == 0) {
                const double Bv = sbm_sf_value(is_log, p1, p2, p3);
                s_sf[g] = Bv; s_p1[g] = p1; s_p2[g] = p2;
                if (a.phase != SBM_LOSS_PHASE_SUMS) a.sf[b * G + g] = Bv;
            }
        }
        __syncwarp();
        if (HAS_JAC) {
            sbm_mbar_wait(mbar + buf, ((nbuf == 2) ? (it >> 1) : it) & 1u);
            // ---- gradient of the scale factors w.r.t. theta (needs the unscaled model Jacobian) ----
            for (int g = 0; g < G; ++g) {
                for (int c = lane; c < P; c += 32) {
                    double j1 = 0.0, j2 = 0.0;
                    if (a.phase != SBM_LOSS_PHASE_FINISH) {
                        // j1 += J d / sigma^2, j2 += J s / sigma^2 (log loss: j1 += J / (s sr^2)); the weights come staged
                        const int q1 = c_goff[g + 1];
                        const double* w1 = is_log ? s_w1 : c_wd;
#pragma unroll 4
                        for (int q = c_goff[g]; q < q1; ++q) {
                            const int r = c_grows[q];
                            const double J = tile[r * P + c];
                            j1 = fma(J, w1[r], j1);
                            j2 = fma(J, s_w2[r], j2);
                        }
                        if (a.phase == SBM_LOSS_PHASE_SUMS) {
                            sums[g * (3 + 2 * P) + 3 + c] = j1; sums[g * (3 + 2 * P) + 3 + P + c] = j2;
                        }
                    } else {
                        j1 = sums[g * (3 + 2 * P) + 3 + c]; j2 = sums[g * (3 + 2 * P) + 3 + P + c];
                    }
                    if (a.phase != SBM_LOSS_PHASE_SUMS) {
                        const double gB = sbm_sfgrad_value(is_log, s_sf[g], s_p1[g], s_p2[g], j1, j2);
                        s_sfg[g * P + c] = gB;
                        a.sfgrad[(b * G + g) * (long long)P + c] = gB;
                    }
                }
            }
            __syncwarp();
        }
        if (a.phase == SBM_LOSS_PHASE_SUMS) {               // (the tile is untouched: nothing to store back)
            if (HAS_JAC && nbuf == 1 && lane == 0 && b + stride < a.B) {
                sbm_mbar_expect_tx(mbar, tile_bytes);
                sbm_tma_load(tile0, a.jac + (b + stride) * (long long)R * P, tile_bytes, mbar);
            }
            continue;
        }
        // ---- residuals (squared_loss_function.py:34-40, log_squared_loss_function.py:23-64) and SSR ----
        double acc = 0.0;
        for (int r = lane; r < R; r += 32) {
            const int meta = c_meta[r];
            const int kind = meta & 15;
            double s = s_sims[r];
            double res;
            if (kind == SBM_ROW_MEASUREMENT) {
                const int g = ((meta >> 4) & 0xfff) - 1;
                if (g >= 0) s *= s_sf[g];
                res = is_log ? (log(s) - c_data[r]) * c_isig[r] : (s - c_data[r]) * c_isig[r];
            } else {
                if (kind == SBM_ROW_SF_PRIOR) { s = log(s_sf[meta >> 16]); sims[r] = s; s_sims[r] = s; }
                res = (s - c_data[r]) * c_isig[r];
            }
            if (allbad) res = __longlong_as_double(SBM_INF_BITS);
            a.resid[b * R + r] = res;
            s_res[r] = res;
            acc += res * res;
        }
        acc = sbm_warp_sum(acc);
        if (lane == 0) a.ssr[b] = 0.5 * acc;               // base_project.py:807-827
        __syncwarp();
        // ---- scaled Jacobian  B*J + sim*dB  and  J^T r  (base_project.py:805), in the tile; bulk store back ----
        if (HAS_JAC) {
            for (int c = lane; c < P; c += 32) {
                double gacc = 0.0;
#pragma unroll 4
                for (int r = 0; r < R; ++r) {
                    const int meta = c_meta[r];
                    double J = sbm_scaled_entry(a, is_log, c, meta & 15, ((meta >> 4) & 0xfff) - 1, meta >> 16, tile[r * P + c],
                                                s_sims[r], s_sf, s_sfg);
                    if (allbad) J = __longlong_as_double(SBM_INF_BITS);
                    tile[r * P + c] = J;
                    gacc = fma(J, s_res[r], gacc);
                }
                if (a.grad) a.grad[b * P + c] = gacc;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                sbm_tma_store(a.jac + b * (long long)R * P, tile, tile_bytes);
                if (nbuf == 1 && b + stride < a.B) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    sbm_mbar_expect_tx(mbar, tile_bytes);
                    sbm_tma_load(tile0, a.jac + (b + stride) * (long long)R * P, tile_bytes, mbar);
                }
            }
        }
    }
    if (HAS_JAC && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// =================================================================================================================
// CTA-per-set kernel: any shape, J stays in global memory
// =================================================================================================================
__device__ __forceinline__ double sbm_block_sum(double v, double* red)
{
    v = sbm_warp_sum(v);
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
    extern __shared__ double dyn[];        // [3][n_groups]: sf, p1, p2
    __shared__ double red[SBM_LOSS_THREADS / 32];
    __shared__ int s_bad;
    double* s_sf = dyn;
    double* s_a1 = dyn + a.n_groups;
    double* s_a2 = dyn + 2 * a.n_groups;
    const int tid = threadIdx.x;
    const bool is_log = (a.loss_kind == SBM_LOSS_LOG);
    const int P = a.n_proj, G = a.n_groups;

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        double* sims = a.sims + b * a.n_rows;
        double* resid = a.resid + b * a.n_rows;
        double* jac = a.jac ? a.jac + b * (long long)a.n_rows * P : nullptr;
        double* sums = a.sums ? a.sums + b * (long long)G * (3 + 2 * P) : nullptr;
        double* sfgrad = a.sfgrad ? a.sfgrad + b * (long long)G * P : nullptr;
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
        for (int g = 0; g < G; ++g) {
            double p1 = 0.0, p2 = 0.0, p3 = 0.0;
            if (a.phase != SBM_LOSS_PHASE_FINISH) {
                for (int q = a.grp_off[g] + tid; q < a.grp_off[g + 1]; q += SBM_LOSS_THREADS) {
                    const int r = a.grp_rows[q];
                    sbm_sf_terms(is_log, sims[r], a.data[r], a.sigma[r], p1, p2, p3);
                }
                p1 = sbm_block_sum(p1, red);
                p2 = sbm_block_sum(p2, red);
                if (is_log) p3 = sbm_block_sum(p3, red);
                if (a.phase == SBM_LOSS_PHASE_SUMS && tid == 0) {
                    sums[g * (3 + 2 * P) + 0] = p1; sums[g * (3 + 2 * P) + 1] = p2; sums[g * (3 + 2 * P) + 2] = p3;
                }
            } else {
                p1 = sums[g * (3 + 2 * P) + 0]; p2 = sums[g * (3 + 2 * P) + 1]; p3 = sums[g * (3 + 2 * P) + 2];
            }
            if (tid == 0) {
                const double Bv = sbm_sf_value(is_log, p1, p2, p3);
                s_sf[g] = Bv; s_a1[g] = p1; s_a2[g] = p2;
                if (a.phase != SBM_LOSS_PHASE_SUMS) a.sf[b * G + g] = Bv;
            }
            __syncthreads();
        }

        // ---- gradient of the scale factors w.r.t. theta (needs the unscaled model Jacobian)
        if (jac) {
            for (int g = 0; g < G; ++g) {
                for (int c = tid; c < P; c += SBM_LOSS_THREADS) {
                    double j1 = 0.0, j2 = 0.0;
                    if (a.phase != SBM_LOSS_PHASE_FINISH) {
                        for (int q = a.grp_off[g]; q < a.grp_off[g + 1]; ++q) {
                            const int r = a.grp_rows[q];
                            sbm_sfgrad_terms(is_log, jac[(long long)r * P + c], sims[r], a.data[r], a.sigma[r], j1, j2);
                        }
                        if (a.phase == SBM_LOSS_PHASE_SUMS) {
                            sums[g * (3 + 2 * P) + 3 + c] = j1; sums[g * (3 + 2 * P) + 3 + P + c] = j2;
                        }
                    } else {
                        j1 = sums[g * (3 + 2 * P) + 3 + c]; j2 = sums[g * (3 + 2 * P) + 3 + P + c];
                    }
                    if (a.phase != SBM_LOSS_PHASE_SUMS)
                        sfgrad[g * (long long)P + c] = sbm_sfgrad_value(is_log, s_sf[g], s_a1[g], s_a2[g], j1, j2);
                }
            }
            __syncthreads();
        }
        if (a.phase == SBM_LOSS_PHASE_SUMS) continue;

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
            if (allbad) res = __longlong_as_double(SBM_INF_BITS);
            resid[r] = res;
            acc += res * res;
        }
        acc = sbm_block_sum(acc, red);
        if (tid == 0) a.ssr[b] = 0.5 * acc;            // base_project.py:807-827
        __syncthreads();

        // ---- scaled Jacobian  B*J + sim*dB  (squared_loss_function.py:58-80) and  J^T r  (base_project.py:805)
        if (jac) {
            for (int c = tid; c < P; c += SBM_LOSS_THREADS) {
                double gacc = 0.0;
                for (int r = 0; r < a.n_rows; ++r) {
                    double J = sbm_scaled_entry(a, is_log, c, a.row_kind[r], a.row_group[r], a.row_aux[r], jac[(long long)r * P + c], sims[r], s_sf, sfgrad);
                    if (allbad) J = __longlong_as_double(SBM_INF_BITS);
                    jac[(long long)r * P + c] = J;
                    gacc += J * resid[r];
                }
                if (a.grad) a.grad[b * P + c] = gacc;
            }
        }
        __syncthreads();
    }
}

// =================================================================================================================
// row-parallel ("wide") path: few parameter sets, big Jacobians (C2 through the Project: 65,536 rows x 16,384 project
// columns per set, 8.6 GB) - the whole grid works on one set's rows and columns.  Six small kernels in stream order;
// every sum has a fixed order (per-chunk partials + a final pass), so results do not depend on the schedule.
//   rows     prior rows <- theta, NaN scan                                          grid (row chunks, B)
//   sf       group sums p1 p2 p3, scale factors (one warp per group)                grid (group chunks, B)
//   sfgrad   j1 j2 sums over a group's rows, dB/dtheta                              grid (column blocks, groups, B)
//   resid    residuals, SSR partials                                                grid (row chunks, B)
//   jac      scaled Jacobian in place, J^T r partials                               grid (column blocks, row chunks, B)
//   reduce   SSR and J^T r from the partials                                        grid (B)
// A CTA of 128 threads is CW columns x (128 / CW) rows, CW = min(128, next power of two >= n_proj): row-major J is read
// and written coalesced along the columns.
// =================================================================================================================
struct SbmLossWide {
    int cw, ry;              // columns per CTA, rows in parallel per CTA
    int n_rchunks, rchunk;   // row chunks and their length
    double* psum;            // [B][G][psum_sg]: p1, p2, p3 (scratch, or the caller's sums buffer)
    long long psum_sb;       // stride of a set / of a group in psum
    int psum_sg;
    double* ssrp;            // [B][n_rchunks]
    double* gradp;           // [B][n_rchunks][n_proj]
    int* bad;                // [B]
};

__global__ void __launch_bounds__(SBM_LOSS_THREADS) sbm_loss_wide_rows(const SbmLossArgs a, const SbmLossWide w)
{
    const long long b = blockIdx.y;
    double* sims = a.sims + b * a.n_rows;
    const int r0 = blockIdx.x * w.rchunk, r1 = min(a.n_rows, r0 + w.rchunk);
    int bad = 0;
    for (int r = r0 + threadIdx.x; r < r1; r += SBM_LOSS_THREADS) {
        const int kind = a.row_kind[r];
        if (kind == SBM_ROW_PARAM_PRIOR) sims[r] = a.theta[b * a.theta_sb + a.row_aux[r] * a.theta_sp];
        if (kind != SBM_ROW_SF_PRIOR) { const double s = sims[r]; if (s != s) bad = 1; }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(w.bad + b, 1);
}

__global__ void __launch_bounds__(SBM_LOSS_THREADS) sbm_loss_wide_sf(const SbmLossArgs a, const SbmLossWide w)
{
    const long long b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const bool is_log = (a.loss_kind == SBM_LOSS_LOG);
    const double* sims = a.sims + b * a.n_rows;
    for (int g = blockIdx.x * (SBM_LOSS_THREADS / 32) + (threadIdx.x >> 5); g < a.n_groups; g += gridDim.x * (SBM_LOSS_THREADS / 32)) {
        double* ps = w.psum + b * w.psum_sb + (long long)g * w.psum_sg;
        double p1 = 0.0, p2 = 0.0, p3 = 0.0;
        if (a.phase != SBM_LOSS_PHASE_FINISH) {
            for (int q = a.grp_off[g] + lane; q < a.grp_off[g + 1]; q += 32) {
                const int r = a.grp_rows[q];
                sbm_sf_terms(is_log, sims[r], a.data[r], a.sigma[r], p1, p2, p3);
            }
            p1 = sbm_warp_sum(p1);
            p2 = sbm_warp_sum(p2);
            p3 = sbm_warp_sum(p3);
            if (lane == 0) { ps[0] = p1; ps[1] = p2; ps[2] = p3; }
        } else {
            p1 = ps[0]; p2 = ps[1]; p3 = ps[2];
        }
        if (lane == 0 && a.phase != SBM_LOSS_PHASE_SUMS) a.sf[b * a.n_groups + g] = sbm_sf_value(is_log, p1, p2, p3);
    }
}

__global__ void __launch_bounds__(SBM_LOSS_THREADS) sbm_loss_wide_sfgrad(const SbmLossArgs a, const SbmLossWide w)
{
    __shared__ double red1[SBM_LOSS_THREADS], red2[SBM_LOSS_THREADS];
    const long long b = blockIdx.z;
    const int P = a.n_proj, G = a.n_groups;
    const int tx = threadIdx.x % w.cw, ty = threadIdx.x / w.cw;
    const int c = blockIdx.x * w.cw + tx;
    const bool is_log = (a.loss_kind == SBM_LOSS_LOG);
    const double* sims = a.sims + b * a.n_rows;
    const double* jac = a.jac + b * (long long)a.n_rows * P;
    for (int g = blockIdx.y; g < G; g += gridDim.y) {
        const double* ps = w.psum + b * w.psum_sb + (long long)g * w.psum_sg;
        double* sums = a.sums ? a.sums + (b * G + g) * (long long)(3 + 2 * P) : nullptr;
        double j1 = 0.0, j2 = 0.0;
        if (a.phase != SBM_LOSS_PHASE_FINISH) {
            if (c < P) {
                const int q1 = a.grp_off[g + 1];
#pragma unroll 4
                for (int q = a.grp_off[g] + ty; q < q1; q += w.ry) {
                    const int r = a.grp_rows[q];
                    sbm_sfgrad_terms(is_log, jac[(long long)r * P + c], sims[r], a.data[r], a.sigma[r], j1, j2);
                }
            }
            if (w.ry > 1) {
                __syncthreads();
                red1[threadIdx.x] = j1; red2[threadIdx.x] = j2;
                __syncthreads();
                if (ty == 0) { j1 = 0.0; j2 = 0.0; for (int k = 0; k < w.ry; ++k) { j1 += red1[k * w.cw + tx]; j2 += red2[k * w.cw + tx]; } }
            }
            if (a.phase == SBM_LOSS_PHASE_SUMS && ty == 0 && c < P) { sums[3 + c] = j1; sums[3 + P + c] = j2; }
        } else if (ty == 0 && c < P) {
            j1 = sums[3 + c]; j2 = sums[3 + P + c];
        }
        if (a.phase != SBM_LOSS_PHASE_SUMS && ty == 0 && c < P)
            a.sfgrad[(b * G + g) * (long long)P + c] = sbm_sfgrad_value(is_log, a.sf[b * G + g], ps[0], ps[1], j1, j2);
    }
}

__global__ void __launch_bounds__(SBM_LOSS_THREADS) sbm_loss_wide_resid(const SbmLossArgs a, const SbmLossWide w)
{
    __shared__ double red[SBM_LOSS_THREADS / 32];
    const long long b = blockIdx.y;
    const bool is_log = (a.loss_kind == SBM_LOSS_LOG);
    double* sims = a.sims + b * a.n_rows;
    const double* sf = a.sf + b * a.n_groups;
    const bool allbad = w.bad[b] != 0;
    const int r0 = blockIdx.x * w.rchunk, r1 = min(a.n_rows, r0 + w.rchunk);
    double acc = 0.0;
    for (int r = r0 + threadIdx.x; r < r1; r += SBM_LOSS_THREADS) {
        const int kind = a.row_kind[r];
        double s = sims[r];
        double res;
        if (kind == SBM_ROW_MEASUREMENT) {
            const int g = a.row_group[r];
            if (g >= 0) s *= sf[g];
            res = is_log ? (log(s) - log(a.data[r])) / a.sigma[r] : (s - a.data[r]) / a.sigma[r];
        } else {
            if (kind == SBM_ROW_SF_PRIOR) { s = log(sf[a.row_aux[r]]); sims[r] = s; }
            res = (s - a.data[r]) / a.sigma[r];
        }
        if (allbad) res = __longlong_as_double(SBM_INF_BITS);
        a.resid[b * a.n_rows + r] = res;
        acc += res * res;
    }
    acc = sbm_block_sum(acc, red);
    if (threadIdx.x == 0) w.ssrp[b * w.n_rchunks + blockIdx.x] = acc;
}

__global__ void __launch_bounds__(SBM_LOSS_THREADS) sbm_loss_wide_jac(const SbmLossArgs a, const SbmLossWide w)
{
    __shared__ double red[SBM_LOSS_THREADS];
    const long long b = blockIdx.z;
    const int P = a.n_proj;
    const int tx = threadIdx.x % w.cw, ty = threadIdx.x / w.cw;
    const int c = blockIdx.x * w.cw + tx;
    const bool is_log = (a.loss_kind == SBM_LOSS_LOG);
    const double* sims = a.sims + b * a.n_rows;
    const double* resid = a.resid + b * a.n_rows;
    const double* sf = a.sf + b * a.n_groups;
    const double* sfgrad = a.sfgrad + b * (long long)a.n_groups * P;
    double* jac = a.jac + b * (long long)a.n_rows * P;
    const bool allbad = w.bad[b] != 0;
    const int r0 = blockIdx.y * w.rchunk, r1 = min(a.n_rows, r0 + w.rchunk);
    double gacc = 0.0;
    if (c < P) {
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += w.ry) {
            double J = sbm_scaled_entry(a, is_log, c, a.row_kind[r], a.row_group[r], a.row_aux[r], jac[(long long)r * P + c], sims[r], sf, sfgrad);
            if (allbad) J = __longlong_as_double(SBM_INF_BITS);
            jac[(long long)r * P + c] = J;
            gacc = fma(J, resid[r], gacc);
        }
    }
    if (w.ry > 1) {
        red[threadIdx.x] = gacc;
        __syncthreads();
        if (ty == 0) { gacc = 0.0; for (int k = 0; k < w.ry; ++k) gacc += red[k * w.cw + tx]; }
    }
    if (ty == 0 && c < P) w.gradp[(b * w.n_rchunks + blockIdx.y) * (long long)P + c] = gacc;
}

__global__ void __launch_bounds__(SBM_LOSS_THREADS) sbm_loss_wide_reduce(const SbmLossArgs a, const SbmLossWide w)
{
    const long long b = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double acc = 0.0;
        for (int k = 0; k < w.n_rchunks; ++k) acc += w.ssrp[b * w.n_rchunks + k];
        a.ssr[b] = 0.5 * acc;                          // base_project.py:807-827
    }
    if (a.jac && a.grad) {
        const int c = blockIdx.x * SBM_LOSS_THREADS + threadIdx.x;
        if (c < a.n_proj) {
            double acc = 0.0;
            for (int k = 0; k < w.n_rchunks; ++k) acc += w.gradp[(b * w.n_rchunks + k) * (long long)a.n_proj + c];
            a.grad[b * a.n_proj + c] = acc;            // base_project.py:805
        }
    }
}

static inline cudaError_t sbm_launch_loss_wide(const SbmLossArgs& a, cudaStream_t st, int sm_count)
{
    SbmLossWide w;
    const int P = a.n_proj, G = a.n_groups, R = a.n_rows;
    w.cw = 1;
    while (w.cw < P && w.cw < SBM_LOSS_THREADS) w.cw <<= 1;
    w.ry = SBM_LOSS_THREADS / w.cw;
    const int colblocks = (P + w.cw - 1) / w.cw;
    // row chunks: enough CTAs for ~8 per SM over (column blocks x sets), chunks of at least 4 row groups
    long long want = (8ll * (sm_count > 0 ? sm_count : 148) + colblocks * a.B - 1) / (colblocks * a.B);
    if (want < 1) want = 1;
    if (want > 1024) want = 1024;
    w.rchunk = (int)((R + want - 1) / want);
    if (w.rchunk < 4 * w.ry) w.rchunk = 4 * w.ry;
    if (w.rchunk > R) w.rchunk = R;
    w.n_rchunks = (R + w.rchunk - 1) / w.rchunk;
    const bool own_psum = (a.sums == nullptr);
    const size_t n_ps = own_psum ? (size_t)a.B * (G > 0 ? G : 0) * 3 : 0;
    const size_t n_ssrp = (size_t)a.B * w.n_rchunks;
    const size_t n_gradp = a.jac ? (size_t)a.B * w.n_rchunks * P : 0;
    const size_t n_bad = ((size_t)a.B + 1) / 2;
    double* scratch = nullptr;
    cudaError_t e = cudaMallocAsync(&scratch, (n_ps + n_ssrp + n_gradp + n_bad + 1) * sizeof(double), st);
    if (e != cudaSuccess) return e;
    w.psum = own_psum ? scratch : a.sums;
    w.psum_sg = own_psum ? 3 : 3 + 2 * P;
    w.psum_sb = (long long)G * w.psum_sg;
    w.ssrp = scratch + n_ps;
    w.gradp = w.ssrp + n_ssrp;
    w.bad = reinterpret_cast<int*>(w.gradp + n_gradp);
    cudaMemsetAsync(w.bad, 0, n_bad * sizeof(double), st);
    const unsigned Bz = (unsigned)a.B;
    sbm_loss_wide_rows<<<dim3(w.n_rchunks, Bz), SBM_LOSS_THREADS, 0, st>>>(a, w);
    if (G > 0) {
        const int gblocks = (G + SBM_LOSS_THREADS / 32 - 1) / (SBM_LOSS_THREADS / 32);
        sbm_loss_wide_sf<<<dim3(gblocks < 4096 ? gblocks : 4096, Bz), SBM_LOSS_THREADS, 0, st>>>(a, w);
        if (a.jac) sbm_loss_wide_sfgrad<<<dim3(colblocks, G < 16384 ? G : 16384, Bz), SBM_LOSS_THREADS, 0, st>>>(a, w);
    }
    if (a.phase != SBM_LOSS_PHASE_SUMS) {
        sbm_loss_wide_resid<<<dim3(w.n_rchunks, Bz), SBM_LOSS_THREADS, 0, st>>>(a, w);
        if (a.jac) sbm_loss_wide_jac<<<dim3(colblocks, w.n_rchunks, Bz), SBM_LOSS_THREADS, 0, st>>>(a, w);
        sbm_loss_wide_reduce<<<dim3((a.jac && a.grad) ? (P + SBM_LOSS_THREADS - 1) / SBM_LOSS_THREADS : 1, Bz), SBM_LOSS_THREADS, 0, st>>>(a, w);
    }
    e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    return e;
}

// tile kernel launch shape: warps per CTA so that two CTAs share an SM's shared memory
static inline bool sbm_loss_tile_shape(const SbmLossArgs& a, int n_grp_rows, SbmLossTileShape& sp)
{
    const long long tile = a.jac ? (long long)a.n_rows * a.n_proj : 0;
    if (a.jac && ((tile & 1) || tile * 8 > 40 * 1024)) return false;      // bulk copies move multiples of 16 bytes
    if ((reinterpret_cast<size_t>(a.jac) & 15) != 0) return false;
    if (a.n_rows > 2048 || a.n_groups > 4000 || a.n_proj > 32000) return false;
    sp.tile_doubles = (int)tile;
    sp.n_grp_rows = n_grp_rows;
    long long cd = 4ll * a.n_rows + (a.n_rows + 1) / 2 + (n_grp_rows + 1) / 2 + (a.n_groups + 2) / 2 + 2;
    cd += (cd & 1);
    // one tile buffer per warp when two would leave fewer than 8 warps per CTA (two CTAs per SM): with 16+ warps per SM the
    // other warps cover a load; measured on C4 (7 KB tiles): 2 buffers x 7 warps 0.41 ms, see DESIGN.md
    const long long budget = 110 * 1024 - cd * 8;
    const long long rest = 4ll * a.n_rows + 3ll * a.n_groups + (long long)a.n_groups * a.n_proj;
    sp.nbuf = 2;
    if (const char* e = getenv("SBM_LOSS_NBUF")) { int v = atoi(e); if (v == 1 || v == 2) sp.nbuf = v; }
    else if (tile > 0 && budget / ((2 * tile + rest + 1) * 8 + 16) < 8) sp.nbuf = 1;
    long long wd = sp.nbuf * tile + rest;
    wd += (wd & 1);                                                      // keep every warp's block 16-byte aligned
    if (budget < wd * 8 + 16) return false;
    sp.cta_doubles = (int)cd;
    sp.warp_doubles = (int)wd;
    sp.warps = (int)(budget / (wd * 8 + 16));
    const int wmax = getenv("SBM_LOSS_WARPS") ? atoi(getenv("SBM_LOSS_WARPS")) : 12;
    if (sp.warps > wmax) sp.warps = wmax;
    sp.smem = (size_t)cd * 8 + (size_t)sp.warps * ((size_t)wd * 8 + 16);
    return true;
}

static inline cudaError_t sbm_launch_loss(const SbmLossArgs& a, cudaStream_t st, int sm_count, int n_grp_rows)
{
    SbmLossTileShape sp;
    static bool attr_set = false;
    if (!(a.flags & SBM_FLAG_LOSS_GENERIC) && sbm_loss_tile_shape(a, n_grp_rows, sp)) {
        if (!attr_set) {
            cudaFuncSetAttribute(sbm_loss_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
            cudaFuncSetAttribute(sbm_loss_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
            attr_set = true;
        }
        long long blocks = (a.B + sp.warps - 1) / sp.warps;
        const long long cap = (long long)(sm_count > 0 ? sm_count : 148) * (a.jac ? 2 : 8);
        if (blocks > cap) blocks = cap;
        if (a.jac) sbm_loss_tile_kernel<true><<<(unsigned)blocks, sp.warps * 32, sp.smem, st>>>(a, sp);
        else sbm_loss_tile_kernel<false><<<(unsigned)blocks, sp.warps * 32, sp.smem, st>>>(a, sp);
        return cudaGetLastError();
    }
    // few sets with big Jacobians: spread each set over the grid (SBM_LOSS_WIDE=0 / 1 overrides the choice)
    bool wide = a.B < 2ll * (sm_count > 0 ? sm_count : 148) && (long long)a.n_rows * a.n_proj >= 32768 && a.B < 65536;
    if (const char* e = getenv("SBM_LOSS_WIDE")) wide = (atoi(e) != 0) && a.B < 65536;
    if (wide) return sbm_launch_loss_wide(a, st, sm_count);
    long long blocks = a.B < 1048576 ? a.B : 1048576;
    size_t gs = (size_t)(3 * (a.n_groups > 0 ? a.n_groups : 1)) * sizeof(double);
    sbm_loss_kernel<<<(unsigned)blocks, SBM_LOSS_THREADS, gs, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// per-set normal equations: out[b] = [ lower triangle of J^T J (row-major, P (P+1) / 2) | J^T r (P) | r^T r ]
// One warp per set, J rows read coalesced (lane = column), the triangle of a set spread over the lanes.
// What the Jacobian-block gather of the multi-GPU path ships instead of the full [n_rows x n_proj] blocks
//  and what an optimiser's normal-equation step consumes.
// ------------------------------------------------------------------------------------------------
struct SbmGramArgs {
    const double* jac;
    const double* resid;
    double* out;
    long long B;
    int n_rows, n_proj;
};

__global__ void __launch_bounds__(256) sbm_gram_kernel(const SbmGramArgs a)
{
    extern __shared__ double dyn[];        // per warp: one row of J (n_proj) + its residual
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int P = a.n_proj, R = a.n_rows, T = P * (P + 1) / 2, OUT = T + P + 1;
    double* row = dyn + warp * (P + 1);
    constexpr int MAXE = 17;               // entries of the triangle per lane: P <= 32 -> T <= 528 = 16.5 * 32
    for (long long b = (long long)blockIdx.x * nwarps + warp; b < a.B; b += (long long)gridDim.x * nwarps) {
        const double* J = a.jac + b * (long long)R * P;
        const double* res = a.resid + b * R;
        double accT[MAXE], accG = 0.0, accR = 0.0;
        int ei[MAXE], ej[MAXE];
#pragma unroll
        for (int k = 0; k < MAXE; ++k) {
            accT[k] = 0.0;
            int e = lane + 32 * k;
            if (e >= T) e = 0;
            // (i, j) of the e-th entry of the row-major lower triangle: i = floor((sqrt(8e + 1) - 1) / 2)
            int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (i * (i + 1) / 2 > e) --i;
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            ei[k] = i;
            ej[k] = e - i * (i + 1) / 2;
        }
        for (int r = 0; r < R; ++r) {
            const double rr = res[r];
            __syncwarp();
            if (lane < P) row[lane] = J[(long long)r * P + lane];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < MAXE; ++k)
                if (lane + 32 * k < T) accT[k] = fma(row[ei[k]], row[ej[k]], accT[k]);
            if (lane < P) accG = fma(row[lane], rr, accG);
            if (lane == 0) accR = fma(rr, rr, accR);
        }
        double* out = a.out + b * OUT;
#pragma unroll
        for (int k = 0; k < MAXE; ++k)
            if (lane + 32 * k < T) out[lane + 32 * k] = accT[k];
        if (lane < P) out[T + lane] = accG;
        if (lane == 0) out[T + P] = accR;
    }
}

static inline cudaError_t sbm_launch_gram(const SbmGramArgs& a, cudaStream_t st, int sm_count)
{
    const int warps = 8;
    long long blocks = (a.B + warps - 1) / warps;
    const long long cap = (long long)(sm_count > 0 ? sm_count : 148) * 8;
    if (blocks > cap) blocks = cap;
    sbm_gram_kernel<<<(unsigned)blocks, warps * 32, (size_t)warps * (a.n_proj + 1) * sizeof(double), st>>>(a);
    return cudaGetLastError();
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
