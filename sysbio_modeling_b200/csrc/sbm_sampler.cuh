// sbm_sampler.cuh - the two device steps of the batched Metropolis ensemble sampler,
// compiled statically into libsbm_b200.so.  W independent walkers advance in lock-step; per Monte-Carlo step
//
//   sbm_mc_propose_kernel   nxt = cur + samp_mat . z            reference Ensembles.py:266-270 (_trial_move), :134
//   (integrator + loss epilogue of the candidates: the hot path)
//   sbm_mc_accept_kernel    F = SSR - T x entropy of the scale factors      base_project.py:854-892
//                           accept if u < exp(-(F_new - F_cur) / T)         Ensembles.py:193-198 (_accept_move)
//                           cur <- nxt, F_cur <- F_new for accepted walkers; every (skip_elems + 1)-th step is recorded
//
// so a step is four launches and no host synchronisation; the ensemble stays in HBM until the run ends.
//   * z and u come from a counter-based generator in the kernel (Philox4x32-10: counter = (walker, index, step), key =
//     seed; Box-Muller for the normals) or, for tests that share random numbers with the CPU restatement, from buffers.
//   * The entropy integral (linear_scale_factor.py:13-18, 63-80; QUADPACK over the whole line in the reference) is a
//     composite 16-point Gauss-Legendre rule over a window that covers the likelihood peak and the prior peak (12 widths
//     either side), panels no wider than the narrower of the two - the rule of the host implementation
//     (LinearScaleFactor.calc_scale_factor_entropy), with the panel count chosen per walker.
#ifndef SBM_SAMPLER_CUH
#define SBM_SAMPLER_CUH

#include <cuda_runtime.h>

struct SbmMcArgs {
    long long W;
    int P, n_rows, n_groups;
    const double* samp;       // [P][P] row-major sampling matrix
    double* cur;              // [W][P]
    double* nxt;              // [W][P]
    const double* normals;    // [W][P] or NULL (device generator)
    const double* uniforms;   // [W] or NULL
    unsigned long long seed;
    unsigned step;
    double* curF;             // [W]
    const double* ssr_new;    // [W]
    const double* sims_new;   // [W][n_rows] unscaled simulations of the candidates
    double temperature;
    const double* data;       // plan tables (device)
    const double* sigma;
    const int* grp_off;
    const int* grp_rows;
    const double* sf_mu;      // [G] log prior of each group's scale factor (NaN: none - the group has no entropy term)
    const double* sf_sig;     // [G]
    double* accepted;         // [W] number of accepted moves
    double* ens_out;          // [W][P] slot of the recorded ensemble, or NULL
    double* fes_out;          // [W]
    int init;                 // 1: no move - F_cur <- F(cur) from ssr_new / sims_new (the starting point)
};

__constant__ double sbm_gl16_x[8] = {0.095012509837637441, 0.28160355077925892, 0.45801677765722737, 0.61787624440264377,
                                     0.755404408355003, 0.86563120238783176, 0.9445750230732326, 0.98940093499164994};
__constant__ double sbm_gl16_w[8] = {0.18945061045506864, 0.18260341504492364, 0.16915651939500265, 0.14959598881657671,
                                     0.12462897125553407, 0.095158511682492605, 0.062253523938647456, 0.027152459411754176};

// Philox4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ void sbm_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double sbm_u01(const unsigned hi, const unsigned lo)      // uniform in (0, 1), 53 bits
{
    const unsigned long long v = ((unsigned long long)hi << 21) ^ (unsigned long long)(lo >> 11);
    return ((double)(v & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

#define SBM_MC_WARPS 8

// nxt = cur + samp . z, one warp per walker
__global__ void __launch_bounds__(SBM_MC_WARPS * 32) sbm_mc_propose_kernel(const SbmMcArgs a)
{
    extern __shared__ double dyn[];            // z of every warp: [warps][P]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* z = dyn + warp * a.P;
    const int P = a.P;
    for (long long w = (long long)blockIdx.x * SBM_MC_WARPS + warp; w < a.W; w += (long long)gridDim.x * SBM_MC_WARPS) {
        __syncwarp();
        for (int k = lane; k < P; k += 32) {
            double v;
            if (a.normals) {
                v = a.normals[w * P + k];
            } else {
                unsigned r[4];
                sbm_philox((unsigned)w, (unsigned)(w >> 32), (unsigned)k, a.step, (unsigned)a.seed, (unsigned)(a.seed >> 32) ^ 0x5eed0001u, r);
                const double u1 = sbm_u01(r[0], r[1]), u2 = sbm_u01(r[2], r[3]);
                v = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);                 // Box-Muller
            }
            z[k] = v;
        }
        __syncwarp();
        for (int j = lane; j < P; j += 32) {
            double acc = 0.0;
            for (int k = 0; k < P; ++k) acc = fma(a.samp[j * P + k], z[k], acc);
            a.nxt[w * P + j] = a.cur[w * P + j] + acc;
        }
    }
}

__device__ __forceinline__ double sbm_warp_sum_mc(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// T x sum over the groups of log integral: the scale-factor entropy of one walker (warp-cooperative)
__device__ __forceinline__ double sbm_mc_entropy(const SbmMcArgs& a, const double* __restrict__ sims, const int lane)
{
    const double T = a.temperature;
    double ent = 0.0;
    for (int g = 0; g < a.n_groups; ++g) {
        const double mu = a.sf_mu[g], sg = a.sf_sig[g];
        if (!(mu == mu) || !(sg == sg)) continue;
        double sds = 0.0, sde = 0.0;
        for (int q = a.grp_off[g] + lane; q < a.grp_off[g + 1]; q += 32) {
            const int r = a.grp_rows[q];
            const double s = sims[r], d = a.data[r], w = 1.0 / (a.sigma[r] * a.sigma[r]);
            sds += s * s * w;
            sde += s * d * w;
        }
        sds = sbm_warp_sum_mc(sds);
        sde = sbm_warp_sum_mc(sde);
        const double Bf = sde / sds, lB = log(Bf);
        const double w_like = sqrt(T / (sds * Bf * Bf));
        const double wide = fmax(w_like, sg), narrow = fmax(fmin(w_like, sg), 1e-9);
        const double lo = fmin(0.0, mu - lB) - 12.0 * wide, hi = fmax(0.0, mu - lB) + 12.0 * wide;
        double np_d = ceil((hi - lo) / narrow);
        if (!(np_d >= 8.0)) np_d = 8.0;                      // also NaN
        if (np_d > 4096.0) np_d = 4096.0;
        const int npan = (int)np_d;
        const double half = 0.5 * (hi - lo) / (double)npan;
        double tot = 0.0;
        for (int p = lane; p < npan; p += 32) {
            const double mid = lo + (2.0 * p + 1.0) * half;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int sgn = -1; sgn <= 1; sgn += 2) {
                    const double u = mid + sgn * half * sbm_gl16_x[k];
                    const double e1 = Bf * exp(u) - Bf, e2 = u + lB - mu;
                    s += sbm_gl16_w[k] * exp(-sds / (2.0 * T) * e1 * e1 - e2 * e2 / (2.0 * sg * sg));     // linear_scale_factor.py:13-18
                }
            }
            tot += half * s;
        }
        tot = sbm_warp_sum_mc(tot);
        ent += T * log(tot);                                 // abstract_loss_function.py:132
    }
    return ent;
}

// free energy of the candidates, Metropolis rule, bookkeeping; one warp per walker
__global__ void __launch_bounds__(SBM_MC_WARPS * 32) sbm_mc_accept_kernel(const SbmMcArgs a)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = a.P;
    for (long long w = (long long)blockIdx.x * SBM_MC_WARPS + warp; w < a.W; w += (long long)gridDim.x * SBM_MC_WARPS) {
        double F = a.ssr_new[w];                              // reference: rss - entropy (base_project.py:888-891)
        if (a.n_groups > 0 && a.sf_mu) F -= sbm_mc_entropy(a, a.sims_new + w * a.n_rows, lane);
        if (!(F == F) || F > 1e300 || F < -1e300) F = __longlong_as_double(0x7ff0000000000000ll);
        bool acc;
        if (a.init) {
            acc = false;
            if (lane == 0) { a.curF[w] = F; a.accepted[w] = 0.0; }
        } else {
            double u;
            if (a.uniforms) u = a.uniforms[w];
            else {
                unsigned r[4];
                sbm_philox((unsigned)w, (unsigned)(w >> 32), 0xffffffffu, a.step, (unsigned)a.seed, (unsigned)(a.seed >> 32) ^ 0x5eed0002u, r);
                u = sbm_u01(r[0], r[1]);
            }
            const double Fc = a.curF[w];
            acc = u < exp(-(F - Fc) / a.temperature);         // _accept_move (Ensembles.py:193-198)
            __syncwarp();
            if (acc) {
                for (int j = lane; j < P; j += 32) a.cur[w * P + j] = a.nxt[w * P + j];
                if (lane == 0) { a.curF[w] = F; a.accepted[w] += 1.0; }   // (the reference forgets to update curr_F, :151-154)
            }
        }
        if (a.ens_out) {
            __syncwarp();
            for (int j = lane; j < P; j += 32) a.ens_out[w * P + j] = acc ? a.nxt[w * P + j] : a.cur[w * P + j];
            if (lane == 0) a.fes_out[w] = (acc || a.init) ? F : a.curF[w];
        }
    }
}

static inline cudaError_t sbm_launch_mc_propose(const SbmMcArgs& a, cudaStream_t st, int sm_count)
{
    long long blocks = (a.W + SBM_MC_WARPS - 1) / SBM_MC_WARPS;
    const long long cap = (long long)(sm_count > 0 ? sm_count : 148) * 8;
    if (blocks > cap) blocks = cap;
    sbm_mc_propose_kernel<<<(unsigned)blocks, SBM_MC_WARPS * 32, (size_t)SBM_MC_WARPS * a.P * sizeof(double), st>>>(a);
    return cudaGetLastError();
}

static inline cudaError_t sbm_launch_mc_accept(const SbmMcArgs& a, cudaStream_t st, int sm_count)
{
    long long blocks = (a.W + SBM_MC_WARPS - 1) / SBM_MC_WARPS;
    const long long cap = (long long)(sm_count > 0 ? sm_count : 148) * 8;
    if (blocks > cap) blocks = cap;
    sbm_mc_accept_kernel<<<(unsigned)blocks, SBM_MC_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

#endif
