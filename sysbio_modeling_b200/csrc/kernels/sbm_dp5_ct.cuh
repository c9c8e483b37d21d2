// =================================================================================================
// sbm_dp5_ct.cuh - "CTA-transposed" Dormand-Prince 5(4) integrator for sensitivity-augmented systems
//
// Compiled by NVRTC after the model functions and sbm_ode_kernels.cuh (which defines the DP_* tableau).
// Same job as sbm_dp5_sens (reference model/ode_model.py:83-126 + project/utils.py:29-89 +
// project/base_project.py:443-488 for every (experiment, parameter set) pair), different mapping:
//
//   lane = trajectory     32 trajectories per CTA, each with its own t, h and accept/reject decision
//   warp = column group   warp w owns C sensitivity columns S_j = dy/dp_j (the last warp CL <= C) of ALL 32
//                         trajectories
//   model warp(s)         XW warps without columns evaluate the model's primal part and run the step controller
//                         (default: one; small systems run everything in one warp, XW = 0)
//
// Why: with lane = column (sbm_dp5_sens) every lane of a group re-evaluates the model's primal part
// (f, J = df/dy, G = df/dp: rate laws, reciprocals, exp/log) - measured to be more than half of the FP64
// instructions.  Here the primal part is evaluated ONCE per trajectory and stage: the emitter
// (symbolic/sympy_tools.py:_emit_primal_shares) splits it into straight-line shares (one for the model warp in the
// default layout; at the head of a step, where nothing else can run, three); the results go to shared memory
// ([entry][lane], conflict-free), from where the column warps read J and their own columns of G for
// dS_j/dt = J S_j + G_j.
//
// The state column y does not depend on the sensitivities, so the model runs ONE STAGE AHEAD: while the column warps
// work on stage s, the model warp evaluates the model at the input of stage s+1 (double-buffered J and G) - one
// CTA barrier per stage.  For that the state column lives in shared memory in "pull" form (y and the stage
// derivatives f_1..f_7); the partial stage values are prepared one iteration ahead by the column warps.
// A step is eight barrier intervals: controller | head (model at the input of stage 2; column warps: rows of an
// output time) | stages 2..7.  Where the clocks go: DESIGN.md section 6, -DSBM_CT_PROF.
//
// Storage per (trajectory, sensitivity column): the 5 vectors  y, S1..S4  of a "push" formulation
//   stages 2-4 pull   Y_s = y + h sum a_sj k_j           (k1..k3 live in S1..S3)
//   after k4          S1 <- Y5, S2 <- Y6 - h a65 k5, S3 <- Y7 - h (a75 k5 + a76 k6), S4 <- h (e1 k1 + e3 k3 + e4 k4)
//   stages 5-7 push   k5, k6, k7 are added into S2, S3, S4 as soon as they exist; S1 <- k7 (FSAL)
// which needs two vectors fewer than keeping k1..k6.  y, S2, S3 live in registers, S1 and S4 in tensor memory
// (registers for small systems).  Full trajectories never exist; only the observable rows go to HBM.
//
// A rejected step has destroyed k1 of the sensitivity columns (S1 holds k7 of the rejected trial), so the
// next "step" of that lane is a zero-length step, which re-evaluates F(y) in every stage and leaves k1 in S1
// again.  The same zero-length step starts every trajectory (k1 = F(y0) and the norms for the initial step
// size).  Lanes fetch new trajectories from a global queue as they finish (persistent CTAs).
// =================================================================================================

#if defined(SBM_CT_W) && SBM_CT_W > 0      /* the emitter found a layout for this model */

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

// per-lane controller state kept in shared memory between steps (only warp 0 touches it)
#define SBM_CT_D_T      0
#define SBM_CT_D_H      1
#define SBM_CT_D_HPROP  2
#define SBM_CT_D_TK     3
#define SBM_CT_D_TEND   4
#define SBM_CT_D_FACOLD 5
#define SBM_CT_ND       12    /* (the lagged controller keeps a second, speculative set) */
#define SBM_CT_I_KT     0
#define SBM_CT_I_KT1    1
#define SBM_CT_I_STATUS 2
#define SBM_CT_I_MODE   3
#define SBM_CT_I_BITS   4     /* 1 lastrej, 2 clipped, 4 refresh, 8 initf, 16 exhausted */
#define SBM_CT_I_NSTEP  5
#define SBM_CT_NI       12

// -DSBM_CT_JITTER=<n>: random delays (up to ~n x 32 ns, different in every launch, warp and call site) before every CTA
// barrier, every block of tensor-memory transfers and every share of the model.  A missing barrier / fence between a
// producer and a consumer shows up as results that change from launch to launch (tests/test_gpu_determinism.py).
#ifdef SBM_CT_JITTER
__device__ __forceinline__ void sbm_ct_jitter(const unsigned site)
{
    unsigned z = (unsigned)clock() * 2654435761u ^ (threadIdx.x >> 5) * 40503u ^ site * 2246822519u ^ blockIdx.x * 3266489917u;
    z ^= z >> 15; z *= 2246822519u; z ^= z >> 13;
    z = __shfl_sync(0xffffffffu, z, 0);             // one decision per warp: the delay must not split it
    if ((z & 3u) == 0u) __nanosleep(((z >> 8) % (unsigned)(SBM_CT_JITTER)) * 32u);
}
#define SBM_CT_JIT(site) sbm_ct_jitter(site)
#else
#define SBM_CT_JIT(site)
#endif
#define SBM_CT_BAR(nt) do { SBM_CT_JIT(1u); asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory"); } while (0)
#ifndef SBM_CT_BUF_ARG
#define SBM_CT_BUF_ARG 0
#endif
#ifndef SBM_CT_LAG
#define SBM_CT_LAG 0
#endif
#ifndef SBM_CTL_SAFETY
#define SBM_CTL_SAFETY 0.9f
#endif
#ifndef SBM_CTL_FACMAX
#define SBM_CTL_FACMAX 5.0f
#endif
#ifndef SBM_CTL_BOOK_WARP
#define SBM_CTL_BOOK_WARP 0          /* column warp that runs part B of the bookkeeping (a scheduler without the model warp) */
#endif
#define SBM_CTL_TAB_T 64
#define SBM_CTL_TAB_R 128
#ifndef SBM_CT_ACCEPT_AT_BEGIN
#define SBM_CT_ACCEPT_AT_BEGIN 0
#endif
#ifndef SBM_CT_NO_STAGE_ROWS
#define SBM_CT_NO_STAGE_ROWS 0
#endif
#define SBM_CT_EM 4           /* rows of one output time staged in shared memory (more rows take the slow path) */

#ifdef SBM_CT_PROF
// experiment (-DSBM_CT_PROF=<column warp>): clocks of the eight barrier intervals of a step and the busy part of each
// for the model warp and one column warp, printed by CTA 0.  The barrier is "deferred blocking": a clock read right after it
// may execute before the barrier completes, so the read is made to depend on a shared-memory load issued after the barrier.
__shared__ long long sbm_ctl_prof[8];          // clocks of the parts of the controller (lane 0)
__device__ __forceinline__ long long sbm_ct_clock_after_barrier(const int* cta)
{
    const int dep = *reinterpret_cast<const volatile int*>(cta);
    long long t;
    asm volatile("{\n\t.reg .b32 d;\n\tmov.b32 d, %1;\n\tmov.u64 %0, %%clock64;\n\t}" : "=l"(t) : "r"(dep) : "memory");
    return t;
}
#endif

// reciprocal for error weights: the hardware seed alone (relative error <= 1e-6, measured in tests/cuda/check_device_math.cu):
// a weight of the error norm needs no more, and stage 7 of the column warps is on the critical path
__device__ __forceinline__ double sbm_rcp_fast(const double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// max(|a|, |b|) for an error weight, decided by the upper 32 bits only: values that agree in sign-less exponent and
// the upper 20 mantissa bits may give either one (a relative difference <= 2^-20 in the weight, nothing for a norm).
// There is no FP64 max instruction: fmax(fabs(a), fabs(b)) expands to DSETP + two selects + ~5 moves per element; this is
// 2 LOP3 + ISETP + 2 SEL on the integer pipe.  NaN propagates (its upper word is the largest).
__device__ __forceinline__ double sbm_absmax_hi(const double a, const double b)
{
    const int ha = __double2hiint(a) & 0x7fffffff, hb = __double2hiint(b) & 0x7fffffff;
    const bool ta = ha >= hb;
    return __hiloint2double(ta ? ha : hb, ta ? __double2loint(a) : __double2loint(b));
}

#ifndef SBM_CT_JTM
#define SBM_CT_JTM 0
#endif
#ifndef SBM_NJU
#define SBM_NJU SBM_NJ
#endif
#define SBM_NJS (SBM_CT_JTM ? SBM_NJP : SBM_NJU)       /* doubles per lane of one J buffer in shared memory */

#if SBM_CT_JTM
// ---- J through tensor memory ---------------------------------------------------------------------------------
// The model warp stores J of a stage to shared memory as [entry pair][lane][2 doubles] (16 bytes per lane and pair);
// one tcgen05.cp ...32x128b.warpx4 per entry pair copies the 32 x 16 bytes to 4 tensor-memory columns of lanes 0-31
// AND of the other three lane quadrants (measured with tools/tcp_probe.cu: lane l receives the 16 bytes at
// base + (l / 8) * SBO + (l % 8) * 16, SBO = 128: contiguous), so every column warp reads the J entries of its 32
// trajectories from its own quadrant with one tcgen05.ld per row pair instead of one shared-memory load per entry:
// 11 warps x 26 loads x 256 B = 73 KB per stage were 60 % of the kernel's shared-memory traffic, which bounds a stage.
__device__ __forceinline__ void sbm_ct_j_copy(const double* jsm, const unsigned tdst, unsigned long long* bar)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(jsm);
    // shared-memory matrix descriptor (no swizzle): start address >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
    unsigned long long desc = (unsigned long long)((addr & 0x3ffffu) >> 4) | (1ull << 16) | (8ull << 32) | (1ull << 46);
#pragma unroll
    for (int k = 0; k < SBM_NJP / 2; ++k) {
        asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" :: "r"(tdst + 4u * k), "l"(desc) : "memory");
        desc += 32ull;                                 // next [lane][16 B] block: 512 bytes on
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void sbm_ct_j_wait_bar(unsigned long long* bar, const unsigned parity)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();      // a protocol error must not hang the GPU
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

struct SbmCtJRegs { unsigned r[2 * SBM_JRP]; };
__device__ __forceinline__ void sbm_ct_j_issue(const unsigned ta, SbmCtJRegs& q)
{
#if SBM_JRP == 8
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3]), "=r"(q.r[4]), "=r"(q.r[5]), "=r"(q.r[6]), "=r"(q.r[7]),
                   "=r"(q.r[8]), "=r"(q.r[9]), "=r"(q.r[10]), "=r"(q.r[11]), "=r"(q.r[12]), "=r"(q.r[13]), "=r"(q.r[14]), "=r"(q.r[15])
                 : "r"(ta));
#elif SBM_JRP == 4
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3]), "=r"(q.r[4]), "=r"(q.r[5]), "=r"(q.r[6]), "=r"(q.r[7])
                 : "r"(ta));
#else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3]) : "r"(ta));
#endif
}
__device__ __forceinline__ void sbm_ct_j_wait(SbmCtJRegs& q, double (&jr)[SBM_JRP])
{
    // the registers are valid only after the wait: the doubles are formed inside the same asm block (see SbmCtStoreTmem)
#if SBM_JRP == 8
    asm volatile("tcgen05.wait::ld.sync.aligned;\n\tmov.b64 %0, {%8, %9};\n\tmov.b64 %1, {%10, %11};\n\tmov.b64 %2, {%12, %13};\n\t"
                 "mov.b64 %3, {%14, %15};\n\tmov.b64 %4, {%16, %17};\n\tmov.b64 %5, {%18, %19};\n\tmov.b64 %6, {%20, %21};\n\t"
                 "mov.b64 %7, {%22, %23};"
                 : "=d"(jr[0]), "=d"(jr[1]), "=d"(jr[2]), "=d"(jr[3]), "=d"(jr[4]), "=d"(jr[5]), "=d"(jr[6]), "=d"(jr[7])
                 : "r"(q.r[0]), "r"(q.r[1]), "r"(q.r[2]), "r"(q.r[3]), "r"(q.r[4]), "r"(q.r[5]), "r"(q.r[6]), "r"(q.r[7]),
                   "r"(q.r[8]), "r"(q.r[9]), "r"(q.r[10]), "r"(q.r[11]), "r"(q.r[12]), "r"(q.r[13]), "r"(q.r[14]), "r"(q.r[15]));
#elif SBM_JRP == 4
    asm volatile("tcgen05.wait::ld.sync.aligned;\n\tmov.b64 %0, {%4, %5};\n\tmov.b64 %1, {%6, %7};\n\tmov.b64 %2, {%8, %9};\n\t"
                 "mov.b64 %3, {%10, %11};"
                 : "=d"(jr[0]), "=d"(jr[1]), "=d"(jr[2]), "=d"(jr[3])
                 : "r"(q.r[0]), "r"(q.r[1]), "r"(q.r[2]), "r"(q.r[3]), "r"(q.r[4]), "r"(q.r[5]), "r"(q.r[6]), "r"(q.r[7]));
#else
    asm volatile("tcgen05.wait::ld.sync.aligned;\n\tmov.b64 %0, {%2, %3};\n\tmov.b64 %1, {%4, %5};"
                 : "=d"(jr[0]), "=d"(jr[1]) : "r"(q.r[0]), "r"(q.r[1]), "r"(q.r[2]), "r"(q.r[3]));
#endif
}
#endif  /* SBM_CT_JTM */

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
// Storage of the two vectors that do not fit in registers (S1 and S4 of every sensitivity column).
//
// Blackwell tensor memory used as a register-file extension.  TMEM is 128 lanes x 512
// 32-bit columns per SM; a warp reaches the 32 lanes of its quadrant (warp % 4) and with the 32x32b shape thread t
// reads / writes lane t - i.e. every thread owns a private row, exactly what lane = trajectory needs.  An element
// (a double) is two columns; a row pair of one column of S is one tcgen05.ld/st ...x4.  This takes 43 % of the
// kernel's shared-memory bytes off the shared-memory pipe and frees 90 KB of shared memory (MAPK).
// -------------------------------------------------------------------------------------------------
struct SbmCtPair { unsigned r0, r1, r2, r3; };      // two doubles in flight (from TMEM) or at hand (registers)

// S1 / S4 in tensor memory
struct SbmCtStoreTmem {
    unsigned t1, t4;                                // TMEM addresses (lane quadrant | column) of S1[0] and S4[0]
    __device__ __forceinline__ void issue(const int slot, const int e0, SbmCtPair& v) const
    {
        const unsigned ta = (slot == 1 ? t1 : t4) + 2u * e0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.r0), "=r"(v.r1), "=r"(v.r2), "=r"(v.r3) : "r"(ta));
    }
    __device__ __forceinline__ void wait(SbmCtPair& v, double& a, double& b) const
    {
        // the 32-bit registers are valid only after the wait: the doubles are formed inside the same asm block, so
        // nothing can read them earlier (and ptxas coalesces the moves of adjacent registers away)
        asm volatile("tcgen05.wait::ld.sync.aligned;\n\tmov.b64 %0, {%2, %3};\n\tmov.b64 %1, {%4, %5};"
                     : "=d"(a), "=d"(b) : "r"(v.r0), "r"(v.r1), "r"(v.r2), "r"(v.r3));
    }
    __device__ __forceinline__ void store(const int slot, const int e0, const double a, const double b)
    {
        const unsigned ta = (slot == 1 ? t1 : t4) + 2u * e0;
        asm volatile("{\n\t.reg .b32 q0, q1, q2, q3;\n\tmov.b64 {q0, q1}, %1;\n\tmov.b64 {q2, q3}, %2;\n\t"
                     "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {q0, q1, q2, q3};\n\t}"
                     :: "r"(ta), "d"(a), "d"(b) : "memory");
    }
    __device__ __forceinline__ void fence() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
};

// experiment (-DSBM_WHATIF_NO_TMEM): no storage at all for S1 / S4 - what the tensor-memory transfers cost (results are wrong)
struct SbmCtStoreNone {
    unsigned t1, t4;
    __device__ __forceinline__ void issue(const int, const int, SbmCtPair& v) const { v.r0 = 0u; v.r1 = 0x39b4484bu; v.r2 = 0u; v.r3 = 0x39b4484bu; }
    __device__ __forceinline__ void wait(SbmCtPair& v, double& a, double& b) const
    {
        a = __hiloint2double((int)v.r1, (int)v.r0);
        b = __hiloint2double((int)v.r3, (int)v.r2);
    }
    __device__ __forceinline__ void store(const int, const int, const double, const double) {}
    __device__ __forceinline__ void fence() const {}
};

// S1 / S4 in registers: small systems (one warp per CTA = one thread per trajectory) whose five vectors fit the
// register file - every index is a compile-time constant after unrolling
template <int NEL>
struct SbmCtStoreReg {
    double v1[NEL], v4[NEL];
    __device__ __forceinline__ void issue(const int slot, const int e0, SbmCtPair& v) const
    {
        const double a = slot == 1 ? v1[e0] : v4[e0], b = slot == 1 ? v1[e0 + 1] : v4[e0 + 1];
        v.r0 = (unsigned)__double2loint(a); v.r1 = (unsigned)__double2hiint(a);
        v.r2 = (unsigned)__double2loint(b); v.r3 = (unsigned)__double2hiint(b);
    }
    __device__ __forceinline__ void wait(SbmCtPair& v, double& a, double& b) const
    {
        a = __hiloint2double((int)v.r1, (int)v.r0);
        b = __hiloint2double((int)v.r3, (int)v.r2);
    }
    __device__ __forceinline__ void store(const int slot, const int e0, const double a, const double b)
    {
        if (slot == 1) { v1[e0] = a; v1[e0 + 1] = b; } else { v4[e0] = a; v4[e0 + 1] = b; }
    }
    __device__ __forceinline__ void fence() const {}
};

// -------------------------------------------------------------------------------------------------
// one stage on CW sensitivity columns of 32 trajectories.  Gw points at this warp's first column of the
// dense G buffer: the entry of (column c, state i) sits at Gw[SBM_GIDX(c*NE + i)], structural zeros hold 0.
// -------------------------------------------------------------------------------------------------
template <int S, int CW, class ST>
__device__ __forceinline__ void sbm_ct_sens_stage(const double h, const double* __restrict__ Jp, const unsigned jt,
                                                  const double* __restrict__ Gw,
                                                  double (&y)[CW][SBM_N], double (&s2)[CW][SBM_N],
                                                  double (&s3)[CW][SBM_N], ST& st,
                                                  const double rtol, const double atol,
                                                  const bool anyinit, double& sq, double& d0, double& d1)
{
    constexpr int N = SBM_N, NE = N + (N & 1), NP2 = NE / 2;
    // tcgen05.ld / st are .sync.aligned: every thread of the warp must execute the same instruction together.  The stage
    // follows per-lane branches (accept copy, rows of an output time, garbage of a rejected lane) after which the compiler
    // is free to leave the warp diverged; without this reconvergence the transfers below are undefined behaviour.  That is
    // what round 1 met as a "code-generation hazard" (results changing with the way the buffer index reached
    // sbm_ct_primal: a different code layout left a different branch unconverged); experiment: -DSBM_WHATIF_NO_SYNCWARP.
#ifndef SBM_WHATIF_NO_SYNCWARP
    __syncwarp();
#endif
    SBM_CT_JIT(2u + S);
    // Tensor-memory transfers are pipelined by hand: tcgen05.wait::ld waits for EVERY outstanding load of the thread, so
    // the vectors a row pair needs in its update (k1 in stage 4, the error accumulator from stage 5 on) are requested one
    // row pair ahead - right AFTER the wait for the current one - and their latency hides behind the FMA chains of the
    // current row pair (measured with -DSBM_WHATIF_NO_TMEM: exposed TMEM latency was 1,700 clocks of stage 4's 4,150).
    constexpr int USLOT = (S == 4) ? 1 : 4;        // vector read by the updates of this stage (S >= 4)
    SbmCtPair pu[2][CW];                           // [row pair parity][column]
    if (S >= 4) {
#pragma unroll
        for (int c = 0; c < CW; ++c) st.issue(USLOT, c * NE, pu[0][c]);       // joins the batch of the stage inputs
    }
#if SBM_CT_JTM
    // J of this stage: tensor-memory columns jt .. jt + 2 NJP of this warp's lane quadrant, one load per row pair.
    // Stages 2 and 3 (no update work to hide a load behind, registers to spare) keep two row pairs in flight.
    constexpr bool JDBL = (S <= 3);
    SbmCtJRegs jq[2];
    sbm_ct_j_issue(jt, jq[0]);
#endif
    // ---- stage input of every element (element e = c*NE + i; S1 comes in row pairs) ----
    double w[CW][N];
    if (S <= 5) {
        SbmCtPair pv[CW][NP2];
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int ip = 0; ip < NP2; ++ip) st.issue(1, c * NE + 2 * ip, pv[c][ip]);
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int ip = 0; ip < NP2; ++ip) {
                double va, vb;
                st.wait(pv[c][ip], va, vb);
                w[c][2 * ip] = sbm_ct_input<S>(h, y[c][2 * ip], va, s2[c][2 * ip], s3[c][2 * ip]);
                if (2 * ip + 1 < N)
                    w[c][2 * ip + 1] = sbm_ct_input<S>(h, y[c][2 * ip + 1], vb, s2[c][2 * ip + 1], s3[c][2 * ip + 1]);
            }
    } else {
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int i = 0; i < N; ++i) w[c][i] = sbm_ct_input<S>(h, y[c][i], 0.0, s2[c][i], s3[c][i]);
    }
    // ---- rows in pairs: two independent FMA chains per column, one 128-bit load of G, one transfer of S1 / S4 ----
#pragma unroll
    for (int i0 = 0; i0 < N; i0 += 2) {
        const bool two = (i0 + 1 < N);
        const int par = (i0 >> 1) & 1;
        double acc[2][CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) {
#ifdef SBM_WHATIF_NO_G         /* experiment: no loads of G (results are wrong) */
            acc[0][c] = 1e-30;
            acc[1][c] = 2e-30;
#else
            // dense G, both states of the row pair in one 128-bit load (layout: SBM_GIDX in the model source)
            const double2 g2 = *reinterpret_cast<const double2*>(Gw + SBM_GIDX(c * NE + i0));
            acc[0][c] = g2.x;
            acc[1][c] = g2.y;
#endif
        }
#if SBM_CT_JTM
        {
            double jr[SBM_JRP];
            sbm_ct_j_wait(jq[JDBL ? par : 0], jr);
            if (JDBL && i0 + 2 < N) sbm_ct_j_issue(jt + (unsigned)((i0 >> 1) + 1) * (2u * SBM_JRP), jq[par ^ 1]);
            sbm_jrow_tm<CW>(i0 >> 1, jr, w, acc[0], acc[1]);
        }
#else
        sbm_jrow<CW>(i0, Jp, w, acc[0]);
        if (two) sbm_jrow<CW>(i0 + 1, Jp, w, acc[1]);
#endif
        double ua[CW], ub[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) {
            ua[c] = 0.0; ub[c] = 0.0;
            if (S >= 4) st.wait(pu[par][c], ua[c], ub[c]);
        }
        if (S >= 4 && i0 + 2 < N) {
#pragma unroll
            for (int c = 0; c < CW; ++c) st.issue(USLOT, c * NE + i0 + 2, pu[par ^ 1][c]);
        }
#if SBM_CT_JTM
        if (!JDBL && i0 + 2 < N) sbm_ct_j_issue(jt + (unsigned)((i0 >> 1) + 1) * (2u * SBM_JRP), jq[0]);
#endif
#pragma unroll
        for (int c = 0; c < CW; ++c) {
            double v1a = 0.0, v1b = 0.0, v4a = 0.0, v4b = 0.0;
            if (S == 4) { v1a = ua[c]; v1b = ub[c]; }
            if (S >= 5) { v4a = ua[c]; v4b = ub[c]; }
            double dummy2 = 0.0, dummy3 = 0.0;
            sbm_ct_update<S>(acc[0][c], h, y[c][i0], v1a, s2[c][i0], s3[c][i0], v4a);
            if (two) sbm_ct_update<S>(acc[1][c], h, y[c][i0 + 1], v1b, s2[c][i0 + 1], s3[c][i0 + 1], v4b);
            else sbm_ct_update<S>(0.0, h, 0.0, v1b, dummy2, dummy3, v4b);
            if (S == 4 || S == 7) st.store(1, c * NE + i0, v1a, v1b);
            if (S >= 4 && S <= 6) st.store(4, c * NE + i0, v4a, v4b);
            if (S == 7) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (r == 0 || two) {
                        const int i = i0 + r;
                        const double isk = sbm_rcp_fast(fma(rtol, sbm_absmax_hi(y[c][i], s3[c][i]), atol));
                        const double q = (r == 0 ? v4a : v4b) * isk;
                        sq = fma(q, q, sq);
                        if (anyinit) {
                            const double q0 = y[c][i] * isk, q1 = acc[r][c] * isk;
                            d0 = fma(q0, q0, d0);
                            d1 = fma(q1, q1, d1);
                        }
                    }
                }
            }
        }
    }
    st.fence();
}

// shared-memory views of one CTA
struct SbmCtShared {
    double* Jb;          // [2][NJ][32]        J of the stage being consumed / produced
    double* Gd;          // [2][NF*NE/2][32][2] dense G (column-major per parameter, row pairs adjacent), zeros stay 0
    double* fh;          // [7][N][32]         stage derivatives f_1..f_7 of the state column
    double* yb;          // [N][32]            state y at the start of the step
    double* ynb;         // [N][32]            5th-order state of the step just taken (prepared by the model warp)
    double* Pb;          // [2][N][32]         partial stage values y + h sum_{j<S-1} a_Sj f_j, prepared one iteration ahead
    double* pb;          // [NP][32]           model parameters of each lane's trajectory
    double* sqb;         // [W+3][32]          partial sums of the squared error norm (per column warp), then the state
                         //                    column's error / |y| / |f| sums
    double* hb;          // [32]               step of each lane (0: zero-length step / idle)
    double* cd;          // [SBM_CT_ND][32]    controller doubles
    unsigned long long* cnt;   // [4][32]      steps, rejects, rhs evaluations, trajectories
    long long* trj;      // [32]               trajectory id
    int* ci;             // [SBM_CT_NI][32]    controller ints
    int* fl;             // [32]               per-lane flags of this step
    int* ktb;            // [32]               output-time index to emit
    int* cta;            // [32]               CTA-wide flags in cta[0]
    double* park;        // [2][W][32]         start-up norms |y|, |f| (per column warp)
    unsigned long long* em_mask;  // [SBM_CT_EM][32]  rows of the output time being emitted: state mask ...
    long long* em_off;   // [SBM_CT_EM][32]    ... and offset of the row in the Jacobian output (staged by the controller)
    int* em_n;           // [32]               number of rows of that output time
    int* pcb;            // [NP][32]           project column of each model parameter of the lane's experiment (-1: fixed)
    long long* trow;     // [32]               first output row of the lane's parameter set (b * n_rows)
    long long* em_soff;  // [SBM_CT_EM][32]    rows staged ahead by the model warp: position of the observable ...
    double* em_val;      // [SBM_CT_EM][32]    ... and its value at the trial state
    double* em_tk;       // [32]               the output time after the staged one
    int* em_kt;          // [32]               output-time index the staging belongs to (-1: none)
    unsigned tmem_base;  // TMEM allocation holding S1 and S4 of the sensitivity columns
    unsigned jt_base;    // first tensor-memory column of the two J buffers (SBM_CT_JTM)
    unsigned long long* jbar;  // [2] mbarriers: the tcgen05.cp transfers of J buffer b have completed
    unsigned long long* mbar;  // [2*NB + 2] mbarriers of the asynchronous hand-over: full[NB], empty[NB], done, go
};

// start of a step for a block of CW sensitivity columns: apply the accept decision, write the Jacobian
// rows of an output time, restart on a new trajectory, and clear k1 before a zero-length step
template <int CW>
__device__ __forceinline__ void sbm_ct_pick_state(const int i, const double (&y)[CW][SBM_N], double (&acc)[CW])
{
    // acc[c] = y[c][i] for a run-time i: a jump table instead of N compare-and-select pairs per column
#define SBM_CT_PICK(k) case k: if (k < SBM_N) { _Pragma("unroll") for (int c = 0; c < CW; ++c) acc[c] = y[c][k < SBM_N ? k : 0]; } break;
    switch (i) {
        SBM_CT_PICK(0) SBM_CT_PICK(1) SBM_CT_PICK(2) SBM_CT_PICK(3) SBM_CT_PICK(4) SBM_CT_PICK(5) SBM_CT_PICK(6) SBM_CT_PICK(7)
        SBM_CT_PICK(8) SBM_CT_PICK(9) SBM_CT_PICK(10) SBM_CT_PICK(11) SBM_CT_PICK(12) SBM_CT_PICK(13) SBM_CT_PICK(14)
        SBM_CT_PICK(15) SBM_CT_PICK(16) SBM_CT_PICK(17) SBM_CT_PICK(18) SBM_CT_PICK(19) SBM_CT_PICK(20) SBM_CT_PICK(21)
        SBM_CT_PICK(22) SBM_CT_PICK(23)
    default: break;
    }
#undef SBM_CT_PICK
}

// one observable row of an output time for CW columns: d(observable)/d(theta)
// (mapping: project/utils.py:29-45, 71-89; chain rule + scatter: base_project.py:443-488)
template <int CW>
__device__ __forceinline__ void sbm_ct_emit_row(const SbmSolveArgs& a, const SbmCtShared& sh, const unsigned long long mask,
                                                double* __restrict__ row, const bool poison, const int col0, const int lane,
                                                const double (&y)[CW][SBM_N])
{
    constexpr int N = SBM_N;
    double acc[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[c] = 0.0;
    if ((mask & (mask - 1ull)) == 0ull) {
        sbm_ct_pick_state<CW>(__ffsll((long long)mask) - 1, y, acc);           // 'direct' rows: one state
    } else {
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int i = 0; i < N; ++i) acc[c] += ((mask >> i) & 1ull) ? y[c][i] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < CW; ++c) {
        if (poison) acc[c] = __longlong_as_double(0x7ff8000000000000ll);
        const int m = sbm_free_idx[col0 + c];
        int pc;
        double sc;
        if (a.jac_raw) { pc = col0 + c; sc = 1.0; }
        else { pc = sh.pcb[m * 32 + lane]; sc = sh.pb[m * 32 + lane]; }        // d/dtheta = p d/dp
        if (pc >= 0) {
            const double v = acc[c] * sc;
            if (a.jac_atomic) atomicAdd(row + pc, v); else row[pc] = v;
        }
    }
}

template <int CW, class ST>
__device__ __forceinline__ void sbm_ct_cols_begin(const SbmSolveArgs& a, const SbmCtShared& sh, const int info,
                                                  const int cflags, const double h, const int col0, const int lane,
                                                  double (&y)[CW][SBM_N], const double (&s3)[CW][SBM_N], ST& st)
{
    constexpr int N = SBM_N;
#if SBM_CT_ACCEPT_AT_BEGIN
    if (info & SBM_CT_F_ACCEPT) {
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int i = 0; i < N; ++i) y[c][i] = s3[c][i];
    }
    const double (&ynow)[CW][SBM_N] = y;
#else
    // a lane emits only right after an accepted step (the controller raises EMIT in the pass that advanced t) or to flush
    // NaN rows of a failed trajectory: the rows are those of the accepted solution, still in s3 (copied into y at stage 2)
    const double (&ynow)[CW][SBM_N] = s3;
#endif
#ifndef SBM_WHATIF_NO_COL_EMIT   /* experiment: how much of a step is the emission of Jacobian rows (results are wrong without) */
    if (cflags & SBM_CT_ANYEMIT) {
        if (info & SBM_CT_F_EMIT) {
            // every column warp stalls the whole CTA while one of its lanes writes rows (one step in five on the MAPK
            // benchmark has an emitting lane): masks and row positions come staged from the controller
            const bool poison = (info & SBM_CT_F_POISON) != 0;
            const int n = sh.em_n[lane];
#ifdef SBM_CT_EMIT_UNROLL
#pragma unroll SBM_CT_EMIT_UNROLL
#endif
            for (int q = 0; q < n && q < SBM_CT_EM; ++q)
                sbm_ct_emit_row<CW>(a, sh, sh.em_mask[q * 32 + lane], a.jac + sh.em_off[q * 32 + lane], poison, col0, lane, ynow);
            if (n > SBM_CT_EM) {
                const long long rowbase = sh.trow[lane];
                const int r0 = a.orow_off[sh.ktb[lane]];
                for (int r = r0 + SBM_CT_EM; r < r0 + n; ++r)
                    sbm_ct_emit_row<CW>(a, sh, a.row_mask[r], a.jac + (rowbase + a.row_jac[r]) * (long long)a.n_proj,
                                        poison, col0, lane, ynow);
            }
        }
    }
#endif
    if (info & SBM_CT_F_RESET) {
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int i = 0; i < N; ++i) y[c][i] = 0.0;
    }
    if (__any_sync(0xffffffffu, h == 0.0)) {
        // zero-length step (start-up, refresh after a rejection, idle lane): the stale k1 must not leak into the
        // stage inputs (0 * Inf = NaN after a rejected non-finite trial).  Warp-uniform: TMEM transfers are collective.
        constexpr int NE = N + (N & 1);
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int ip = 0; ip < NE / 2; ++ip) {
                SbmCtPair pv;
                double va, vb;
                st.issue(1, c * NE + 2 * ip, pv);
                st.wait(pv, va, vb);
                st.store(1, c * NE + 2 * ip, h == 0.0 ? 0.0 : va, h == 0.0 ? 0.0 : vb);
            }
        st.fence();
    }
}

template <int CW, class ST>
__device__ __forceinline__ void sbm_ct_cols_stage(const int s, const int info, const double h, const double* __restrict__ Jp,
                                                  const unsigned jt, const double* __restrict__ Gw,
                                                  double (&y)[CW][SBM_N], double (&s2)[CW][SBM_N],
                                                  double (&s3)[CW][SBM_N], ST& st,
                                                  const double rtol, const double atol,
                                                  const bool anyinit, double& sq, double& d0, double& d1)
{
    switch (s) {
    case 2:
#if !SBM_CT_ACCEPT_AT_BEGIN
        // the accept decision (y = the 5th-order solution of the step just taken) is applied here, not at the start of
        // the step: there the head shares wait for it, here the column warps wait for the model warp anyway
        if (info & SBM_CT_F_ACCEPT) {
#pragma unroll
            for (int c = 0; c < CW; ++c)
#pragma unroll
                for (int i = 0; i < SBM_N; ++i) y[c][i] = s3[c][i];
        }
#endif
        sbm_ct_sens_stage<2, CW, ST>(h, Jp, jt, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1); break;
    case 3: sbm_ct_sens_stage<3, CW, ST>(h, Jp, jt, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1); break;
    case 4: sbm_ct_sens_stage<4, CW, ST>(h, Jp, jt, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1); break;
    case 5: sbm_ct_sens_stage<5, CW, ST>(h, Jp, jt, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1); break;
    case 6: sbm_ct_sens_stage<6, CW, ST>(h, Jp, jt, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1); break;
    default: sbm_ct_sens_stage<7, CW, ST>(h, Jp, jt, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1); break;
    }
}

// -------------------------------------------------------------------------------------------------
// step controller, trajectory queue and the state column (warp 0, one lane per trajectory)
// -------------------------------------------------------------------------------------------------
// 5th-order solution, embedded error and start-up norms of the state column from its stage derivatives f_1..f_7
__device__ __forceinline__ void sbm_ct_state_column(const SbmCtShared& sh, const int lane, const double hdone,
                                                    const double rtol, const double atol, double (&ynew)[SBM_N],
                                                    double (&f7)[SBM_N], double& sq, double& q0, double& q1)
{
    constexpr int N = SBM_N;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double f1 = sh.fh[(0 * N + i) * 32 + lane], f3 = sh.fh[(2 * N + i) * 32 + lane];
        const double f4 = sh.fh[(3 * N + i) * 32 + lane], f5 = sh.fh[(4 * N + i) * 32 + lane];
        const double f6 = sh.fh[(5 * N + i) * 32 + lane];
        f7[i] = sh.fh[(6 * N + i) * 32 + lane];
        const double yi = sh.yb[i * 32 + lane];
        ynew[i] = fma(hdone, fma(DP_A76, f6, fma(DP_A75, f5, fma(DP_A74, f4, fma(DP_A73, f3, DP_A71 * f1)))), yi);
        const double er = hdone * fma(DP_E7, f7[i], fma(DP_E6, f6, fma(DP_E5, f5, fma(DP_E4, f4,
                                  fma(DP_E3, f3, DP_E1 * f1)))));
        const double isk = sbm_rcp_fast(fma(rtol, sbm_absmax_hi(yi, ynew[i]), atol));
        const double q = er * isk, u0 = yi * isk, u1 = f7[i] * isk;
        sq = fma(q, q, sq);
        q0 = fma(u0, u0, q0);
        q1 = fma(u1, u1, q1);
    }
}

// Rows of the output time a clipped step lands on, prepared from the TRIAL state while the column warps are still busy
// with stage 7 (the model warp idles there): if the controller accepts the step it only has to store them.
__device__ __forceinline__ void sbm_ct_stage_rows(const SbmSolveArgs& a, const SbmCtShared& sh, const int lane,
                                                  const double (&ynew)[SBM_N])
{
    constexpr int N = SBM_N;
    const int kt = sh.ci[SBM_CT_I_KT * 32 + lane], kt1 = sh.ci[SBM_CT_I_KT1 * 32 + lane];
    const long long rowbase = sh.trow[lane];
    const int r0 = __ldg(a.orow_off + kt), r1 = __ldg(a.orow_off + kt + 1);
    sh.em_tk[lane] = (kt + 1 < kt1) ? __ldg(a.tout + kt + 1) : 0.0;
    sh.em_n[lane] = r1 - r0;
    for (int r = r0; r < r1 && r - r0 < SBM_CT_EM; ++r) {
        const unsigned long long mask = __ldg(a.row_mask + r);
        const int rs = __ldg(a.row_sim + r), rj = __ldg(a.row_jac + r);
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) acc += ((mask >> i) & 1ull) ? ynew[i] : 0.0;                  // utils.py:10-26, 48-68
        sh.em_mask[(r - r0) * 32 + lane] = mask;
        sh.em_off[(r - r0) * 32 + lane] = (rowbase + rj) * (long long)a.n_proj;
        sh.em_soff[(r - r0) * 32 + lane] = rowbase + rs;
        sh.em_val[(r - r0) * 32 + lane] = acc;
    }
    sh.em_kt[lane] = kt;
}

// what the model warp has computed for the controller while the column warps finished stage 7; stays in registers across
// the barrier (PRE == 2), so the controller - on the critical path of the whole CTA - starts with no loads for it
struct SbmCtCarry {
    double ynew[SBM_N], f7[SBM_N], sqs, q0, q1;
};

template <int W, int C, int CL, int PRE>
__device__ __forceinline__ void sbm_ct_control(const SbmSolveArgs& a, const SbmCtShared& sh, const int lane,
                                               const SbmCtCarry& cy)
{
    constexpr int N = SBM_N;
    constexpr unsigned FULL = 0xffffffffu;
    const long long ntraj = a.B * (long long)a.E;
    const double rtol = a.rtol, atol = a.atol;
    const double inv_naug = 1.0 / (double)(N * (SBM_NF + 1));
    unsigned long long* queue = a.counters + SBM_CNT_QUEUE;

#ifdef SBM_CT_PROF
    long long cp0 = clock64();
#define SBM_CTL_MARK(i) { const long long tn = clock64(); if (lane == 0) sbm_ctl_prof[i] += tn - cp0; cp0 = tn; }
#else
#define SBM_CTL_MARK(i)
#endif
    double t = sh.cd[SBM_CT_D_T * 32 + lane], h = sh.cd[SBM_CT_D_H * 32 + lane];
    double hprop = sh.cd[SBM_CT_D_HPROP * 32 + lane], tk = sh.cd[SBM_CT_D_TK * 32 + lane];
    double tend = sh.cd[SBM_CT_D_TEND * 32 + lane];
    float facold = (float)sh.cd[SBM_CT_D_FACOLD * 32 + lane];
    int kt = sh.ci[SBM_CT_I_KT * 32 + lane], kt1 = sh.ci[SBM_CT_I_KT1 * 32 + lane];
    int status = sh.ci[SBM_CT_I_STATUS * 32 + lane], mode = sh.ci[SBM_CT_I_MODE * 32 + lane];
    const int bits = sh.ci[SBM_CT_I_BITS * 32 + lane];
    unsigned nstep_traj = (unsigned)sh.ci[SBM_CT_I_NSTEP * 32 + lane];
    bool lastrej = bits & 1, clipped = bits & 2, refresh = bits & 4, initf = bits & 8, exhausted = bits & 16;
    long long traj = sh.trj[lane];
    const double hdone = sh.hb[lane];

    int info = 0, ktemit = 0;
    bool fetched = false;
    if (mode != SBM_CT_IDLE) {
        // ---- the state column of the step just taken: 5th-order solution, embedded error, norms ----
        // (everything between the partial sums and the new step size is on the critical path of the whole CTA: the
        //  column warps wait for it.  Hence four interleaved sums, no square root and no FP64 division below.)
        double sq = 0.0, q0 = 0.0, q1 = 0.0;
        {
            double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int w = 0; w < W; ++w) part[w & 3] += sh.sqb[w * 32 + lane];
            sq = (part[0] + part[1]) + (part[2] + part[3]);
        }
        double ynew[N], f7[N];
        if (PRE == 2) {
            sq += cy.sqs;
            q0 = cy.q0;
            q1 = cy.q1;
#pragma unroll
            for (int i = 0; i < N; ++i) { ynew[i] = cy.ynew[i]; f7[i] = cy.f7[i]; }
        } else if (PRE) {
            // prepared by the model warp while the column warps were busy with stage 7
            sq += sh.sqb[(W + 0) * 32 + lane];
            q0 = sh.sqb[(W + 1) * 32 + lane];
            q1 = sh.sqb[(W + 2) * 32 + lane];
#pragma unroll
            for (int i = 0; i < N; ++i) { ynew[i] = sh.ynb[i * 32 + lane]; f7[i] = sh.fh[(6 * N + i) * 32 + lane]; }
        } else {
            double sqs = 0.0;
            sbm_ct_state_column(sh, lane, hdone, rtol, atol, ynew, f7, sqs, q0, q1);
            sq += sqs;
        }
        // err2 = err^2; the controller works on the square (half the exponents) in single precision: the step factor
        // needs no more, and the result is rounded the same way for every lane and every launch
        const double err2 = sq * inv_naug;
        SBM_CTL_MARK(0)                            // state and partial sums loaded
        bool fsal = false;
        if (mode == SBM_CT_NORMAL) {
            // PI controller (Hairer's DOPRI5: beta = 0.04, safe = 0.9, step factor in [0.2, 10]); facold holds the SQUARE
            // of the previous accepted error (floor 1e-4 ^ 2).  Accepted and rejected lanes share one instruction stream
            // (selects, predicated stores): only the clean-up after a rejection is a branch.
            nstep_traj++;
            const bool acc = err2 <= 1.0;                                      // false for NaN
            const bool finite = (err2 == err2) && (err2 < 1e300);
            const float ferr2 = fmaxf((float)err2, 1e-37f);                    // +Inf when err2 exceeds the float range
            const float e17 = __powf(ferr2, 0.5f * (0.2f - 0.04f * 0.75f));    // err ^ 0.17
            float ra = fminf(10.0f, fmaxf(0.2f, 0.9f * __fdividef(__powf(facold, 0.5f * 0.04f), e17)));
            if (lastrej) ra = fminf(ra, 1.0f);
            const float rr = finite ? fmaxf(0.2f, __fdividef(0.9f, e17)) : 0.1f;   // NaN/Inf in the trial state: retreat hard
            double hnew = h * (double)(acc ? ra : rr);
            if (acc && clipped && hnew < hprop) hnew = hprop;                  // a clipped step must not shrink the cruise step
            t = acc ? (clipped ? tk : t + h) : t;
            facold = acc ? fmaxf(ferr2, 1e-8f) : facold;
            h = hnew;
            sh.cnt[0 * 32 + lane] += 1;
            sh.cnt[2 * 32 + lane] += acc ? 6 : 7;               // after a rejection k1 of the sensitivity columns is re-evaluated
            if (acc) {
#pragma unroll
                for (int i = 0; i < N; ++i) sh.yb[i * 32 + lane] = ynew[i];
                info |= SBM_CT_F_ACCEPT;
            }
            fsal = acc;
            lastrej = !acc;
            if (!acc) {
                sh.cnt[1 * 32 + lane] += 1;
                refresh = true;
                if (!finite && h < 1e-300) status = SBM_ERR_NONFINITE;
                // stage derivatives of the rejected trial may be non-finite: they must not meet the zero
                // coefficients of the refresh step (0 * Inf)
#pragma unroll
                for (int j = 1; j < 7; ++j)
#pragma unroll
                    for (int i = 0; i < N; ++i) sh.fh[(j * N + i) * 32 + lane] = 0.0;
            }
        } else {
            // a zero-length step just re-evaluated F(y) in every stage
            refresh = false;
            fsal = true;
            if (initf) {
                // initial step size (Hairer, Norsett, Wanner II.4): h0 = 0.01 |y|/|f|; start at the rule's
                // upper bound 100 h0 and let the controller (and, if needed, one rejection) settle it
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    q0 += sh.park[w * 32 + lane];
                    q1 += sh.park[(W + w) * 32 + lane];
                }
                const double n0 = sqrt(q0 * inv_naug), n1 = sqrt(q1 * inv_naug);
                const double h0 = (n0 < 1e-5 || n1 < 1e-5) ? 1e-6 : 0.01 * n0 / n1;
                h = fmin(100.0 * h0, tend);
                if (a.h0 > 0.0) h = fmin(a.h0, tend);
                initf = false;
                sh.cnt[2 * 32 + lane] += 1;
            }
        }
        SBM_CTL_MARK(1)                            // decision and step size
        if (fsal) {
#pragma unroll
            for (int i = 0; i < N; ++i) sh.fh[(0 * N + i) * 32 + lane] = f7[i];
        }
        SBM_CTL_MARK(2)
        // arrived at an output time (or failed: flush the remaining rows as NaN, one time per pass)
        if (!refresh && (t >= tk || status != SBM_OK)) {
            info |= SBM_CT_F_EMIT;
            if (status != SBM_OK) info |= SBM_CT_F_POISON;
            ktemit = kt;
            // (this block stalls the whole CTA once in five steps on the MAPK benchmark: read-only table loads through the
            //  L1 path, issued before they are needed; the row base of the parameter set was kept at fetch time)
            const long long rowbase = sh.trow[lane];
            double tknext;
            int rfirst, r1;
            if (PRE && status == SBM_OK && sh.em_kt[lane] == kt) {
                // staged from the trial state during stage 7 (sbm_ct_stage_rows): store the observables, done
                const int n = sh.em_n[lane];
                for (int q = 0; q < n && q < SBM_CT_EM; ++q) a.sims[sh.em_soff[q * 32 + lane]] = sh.em_val[q * 32 + lane];
                tknext = sh.em_tk[lane];
                rfirst = r1 = 0;
                if (n > SBM_CT_EM) {                                   // rows beyond the staging take the loop below
                    r1 = __ldg(a.orow_off + kt) + n;
                    rfirst = r1 - n + SBM_CT_EM;
                }
            } else {
                rfirst = __ldg(a.orow_off + kt);
                r1 = __ldg(a.orow_off + kt + 1);
                tknext = (kt + 1 < kt1) ? __ldg(a.tout + kt + 1) : tk;
                sh.em_n[lane] = r1 - rfirst;
            }
            sh.em_kt[lane] = -1;
            const int r0 = r1 - sh.em_n[lane];
#pragma unroll 2
            for (int r = rfirst; r < r1; ++r) {
                const unsigned long long mask = __ldg(a.row_mask + r);
                const int rs = __ldg(a.row_sim + r), rj = __ldg(a.row_jac + r);
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) acc += ((mask >> i) & 1ull) ? sh.yb[i * 32 + lane] : 0.0;   // utils.py:10-26, 48-68
                if (status != SBM_OK) acc = __longlong_as_double(0x7ff8000000000000ll);
                a.sims[rowbase + rs] = acc;
                // the column warps write the Jacobian rows of this output time at the start of the next step: hand them
                // the mask and the position of each row (they would otherwise all repeat these index computations)
                if (r - r0 < SBM_CT_EM) {
                    sh.em_mask[(r - r0) * 32 + lane] = mask;
                    sh.em_off[(r - r0) * 32 + lane] = (rowbase + rj) * (long long)a.n_proj;
                }
            }
            kt++;
            if (kt >= kt1) {
                a.status[traj] = status;
                sh.cnt[3 * 32 + lane] += 1;
                mode = SBM_CT_IDLE;
            } else {
                tk = tknext;
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
            fetched = true;                                 // its model parameters are gathered by the whole warp below
            const int e = (int)(traj % a.E);
            const long long b = traj / a.E;
#pragma unroll
            for (int i = 0; i < N; ++i) sh.yb[i * 32 + lane] = a.y0[e * N + i];
            for (int q = 0; q < 7 * N; ++q) sh.fh[q * 32 + lane] = 0.0;
            kt = a.tout_off[e];
            kt1 = a.tout_off[e + 1];
            sh.trj[lane] = traj;
            sh.trow[lane] = b * a.n_rows;
            t = 0.0;
            status = SBM_OK;
            facold = 1e-8f;                                 // (1e-4)^2: the controller keeps squared errors
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
                sh.cnt[3 * 32 + lane] += 1;
            }
        } else {
            exhausted = true;
        }
    }
    // ---- model parameters of the new trajectories: exp(theta) for optimised, the raw value for fixed parameters
    //      (base_project.py:343-363).  One parameter per lane instead of a loop of NP exponentials in the one lane that
    //      fetched (the whole CTA waits for this code) ----
    for (unsigned need = __ballot_sync(FULL, fetched); need; need &= need - 1u) {
        const int L = __ffs(need) - 1;
        const long long trajL = __shfl_sync(FULL, traj, L);
        const int e = (int)(trajL % a.E);
        const long long b = trajL / a.E;
        for (int m = lane; m < SBM_NP; m += 32) {
            const int c = a.pmap[e * SBM_NP + m];
            sh.pb[m * 32 + L] = (c >= 0) ? exp(a.theta[b * a.theta_sb + c * a.theta_sp]) : a.pfix[e * SBM_NP + m];
            sh.pcb[m * 32 + L] = c;
        }
    }
    SBM_CTL_MARK(3)                                // output time / new trajectory
    // ---- next step of this lane ----
    double hstep = 0.0;
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
    sh.hb[lane] = hstep;
    sh.fl[lane] = info;
    sh.ktb[lane] = ktemit;
    const bool alive = __any_sync(FULL, mode != SBM_CT_IDLE) || !__all_sync(FULL, exhausted);
    const bool anyemit = __any_sync(FULL, (info & SBM_CT_F_EMIT) != 0);
    const bool anyinit = __any_sync(FULL, initf);
    if (lane == 0)
        sh.cta[0] = (alive ? SBM_CT_ALIVE : 0) | (anyemit ? SBM_CT_ANYEMIT : 0) | (anyinit ? SBM_CT_ANYINIT : 0);

    sh.cd[SBM_CT_D_T * 32 + lane] = t;
    sh.cd[SBM_CT_D_H * 32 + lane] = h;
    sh.cd[SBM_CT_D_HPROP * 32 + lane] = hprop;
    sh.cd[SBM_CT_D_TK * 32 + lane] = tk;
    sh.cd[SBM_CT_D_TEND * 32 + lane] = tend;
    sh.cd[SBM_CT_D_FACOLD * 32 + lane] = (double)facold;
    sh.ci[SBM_CT_I_KT * 32 + lane] = kt;
    sh.ci[SBM_CT_I_KT1 * 32 + lane] = kt1;
    sh.ci[SBM_CT_I_STATUS * 32 + lane] = status;
    sh.ci[SBM_CT_I_MODE * 32 + lane] = mode;
    sh.ci[SBM_CT_I_BITS * 32 + lane] = (lastrej ? 1 : 0) | (clipped ? 2 : 0) | (refresh ? 4 : 0) | (initf ? 8 : 0) |
                                       (exhausted ? 16 : 0);
    sh.ci[SBM_CT_I_NSTEP * 32 + lane] = (int)nstep_traj;
    SBM_CTL_MARK(4)                                // clipping, flags, state saved
#ifdef SBM_CT_PROF
    if (lane == 0 && blockIdx.x == 0 && !(sh.cta[0] & SBM_CT_ALIVE))
        printf("SBM_CT_PROF controller clocks (total over steps): load %lld decide %lld fsal %lld emit/fetch %lld finish %lld\n",
               sbm_ctl_prof[0], sbm_ctl_prof[1], sbm_ctl_prof[2], sbm_ctl_prof[3], sbm_ctl_prof[4]);
#endif
}

// Tableau rows without their last entry: the stage value of stage S is prepared one iteration ahead as
// P_S = y + h sum_{j <= S-2} a_Sj f_j (any warp, see sbm_ct_prepare) and completed by the shares with the newest f
__constant__ double sbm_dp_ap[8][5] = {
    {0.0, 0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0, 0.0},
    {DP_A31, 0.0, 0.0, 0.0, 0.0},
    {DP_A41, DP_A42, 0.0, 0.0, 0.0},
    {DP_A51, DP_A52, DP_A53, 0.0, 0.0},
    {DP_A61, DP_A62, DP_A63, DP_A64, 0.0},
    {DP_A71, 0.0, DP_A73, DP_A74, DP_A75}};

// this warp's share of the model at the input of stage S (S = 2..7), written into the buffers of stage S
__device__ __forceinline__ void sbm_ct_primal(const SbmCtShared& sh, const int warp, const int lane, const int S,
                                              const double h, const int buf)
{
    // with one CTA barrier per stage the buffer is the stage's parity (the compiler then sees one expression for the
    // producer and the consumers; passing it in was measured to change the generated code's results - see DESIGN.md)
    const int b = (SBM_CT_ASYNC || SBM_CT_BUF_ARG) ? buf : (S & 1);
    SBM_CT_JIT(16u + S);
    const double ts = fma(sbm_dp_c[S], h, sh.cd[SBM_CT_D_T * 32 + lane]);
    const double* Pp = (S == 2) ? sh.yb + lane : sh.Pb + (S & 1) * SBM_N * 32 + lane;
    // stage 2 is the head of the step: its step size exists only after the previous step has been judged, so it
    // overlaps with nothing and is split over more warps (sbm_primal_head) than the stages that run one ahead
    if (S == 2 && !SBM_CT_ASYNC)
        sbm_primal_head<32>(warp, ts, Pp, sh.fh + lane, h * DP_A21, sh.pb + lane, sh.Jb + b * SBM_NJS * 32 + (SBM_CT_JTM ? 2 : 1) * lane,
                            sh.Gd + b * SBM_NF * (SBM_N + (SBM_N & 1)) * 32 + 2 * lane, sh.fh + SBM_N * 32 + lane);
    else
        sbm_primal_share<32>(warp, ts, Pp, sh.fh + (S - 2) * SBM_N * 32 + lane, h * sbm_dp_a[S][S - 2], sh.pb + lane,
                             sh.Jb + b * SBM_NJS * 32 + (SBM_CT_JTM ? 2 : 1) * lane,
                             sh.Gd + b * SBM_NF * (SBM_N + (SBM_N & 1)) * 32 + 2 * lane,
                             sh.fh + (S - 1) * SBM_N * 32 + lane);
#if SBM_CT_JTM
    // the J entries just stored are read by the tensor-memory copy engine (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// partial stage values of stage S (3..7) for the state components this warp looks after (any warp can: the state
// column lives in shared memory); runs one iteration before the shares need them
template <int WT>
__device__ __forceinline__ void sbm_ct_prepare(const SbmCtShared& sh, const int warp, const int lane, const int S,
                                               const double h)
{
    double cf[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) cf[j] = h * sbm_dp_ap[S][j];
    for (int i = warp; i < SBM_N; i += WT) {
        double v = sh.yb[i * 32 + lane];
#pragma unroll
        for (int j = 0; j < 5; ++j) v = fma(cf[j], sh.fh[(j * SBM_N + i) * 32 + lane], v);
        sh.Pb[((S & 1) * SBM_N + i) * 32 + lane] = v;
    }
}

// one warp: CW sensitivity columns + primal share `warp`
template <int W, int C, int CW, int XW>
__device__ __forceinline__ void sbm_ct_warp(const SbmSolveArgs& a, const SbmCtShared& sh, const int warp, const int lane)
{
    constexpr int N = SBM_N, NT = (W + XW) * 32;
    const int col0 = warp * C;                     // first sensitivity column (free-parameter index) of this warp
    constexpr int NE = N + (N & 1);
#if SBM_CT_REGV
    typedef SbmCtStoreReg<CW * NE> ST;
    ST st;
#else
#ifdef SBM_WHATIF_NO_TMEM
    typedef SbmCtStoreNone ST;
#else
    typedef SbmCtStoreTmem ST;
#endif
    ST st;
    // this warp's TMEM window: its lane quadrant, 4*C*NE columns per warp of the quadrant (S1 then S4)
    st.t1 = sh.tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * 4 * C * NE);
    st.t4 = st.t1 + 2u * CW * NE;
#endif
    static_assert(CW * N >= 2, "at least one row pair per warp");
    double y[CW][N], s2[CW][N], s3[CW][N];
#pragma unroll
    for (int c = 0; c < CW; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) { y[c][i] = 0.0; s2[c][i] = 0.0; s3[c][i] = 0.0; }
#pragma unroll
    for (int c = 0; c < CW; ++c)
#pragma unroll
        for (int ip = 0; ip < NE / 2; ++ip) {
            st.store(1, c * NE + 2 * ip, 0.0, 0.0);
            st.store(4, c * NE + 2 * ip, 0.0, 0.0);
        }
    st.fence();
    const double rtol = a.rtol, atol = a.atol;

#ifdef SBM_CT_PROF
    long long cbusy[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ctprev = 0;
#endif
    unsigned jk = 0;                               // stages consumed: J buffer (jk & 1), fill number jk >> 1
    for (;;) {
        if (XW == 0 && warp == 0) { SbmCtCarry none; sbm_ct_control<W, C, SBM_NF - (W - 1) * C, 0>(a, sh, lane, none); }
        SBM_CT_BAR(NT);                            // controller output published
#ifdef SBM_CT_PROF
        ctprev = sbm_ct_clock_after_barrier(sh.cta);
        cbusy[8] += 1;
#endif
        const int cflags = sh.cta[0];
        if (!(cflags & SBM_CT_ALIVE)) break;
        const int info = sh.fl[lane];
        const double h = sh.hb[lane];
        sbm_ct_cols_begin<CW, ST>(a, sh, info, cflags, h, col0, lane, y, s3, st);
#ifdef SBM_CT_PROF
        cbusy[0] += clock64() - ctprev;            // cols_begin alone
#endif
        const bool anyinit = (cflags & SBM_CT_ANYINIT) != 0;
        double sq = 0.0, d0 = 0.0, d1 = 0.0;
        // iteration s: stage s of this warp's columns (s >= 2) and its share of the model at the input of
        // stage s+1 (s <= 6) - the model runs one stage ahead; one barrier per iteration
#pragma unroll 1
        for (int s = 1; s <= 7; ++s) {
#ifndef SBM_WHATIF_COLS_SKIP   /* experiment: the model warp alone (results are wrong) */
            if (s >= 2) {
                const double* Jp = sh.Jb + (s & 1) * SBM_NJS * 32 + lane;
                const double* Gw = sh.Gd + ((s & 1) * SBM_NF + col0) * (N + (N & 1)) * 32 + 2 * lane;
                unsigned jt = 0;
#if SBM_CT_JTM
                sbm_ct_j_wait_bar(sh.jbar + (s & 1), (jk >> 1) & 1u);
                ++jk;
                jt = sh.tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + sh.jt_base + (unsigned)(s & 1) * (2u * SBM_NJP);
#endif
                sbm_ct_cols_stage<CW, ST>(s, info, h, Jp, jt, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1);
            }
#endif
            if (s < 6) sbm_ct_prepare<W + XW>(sh, warp, lane, s + 2, h);
#ifdef SBM_CT_PRIMAL_REPEAT
            // experiment: evaluate the share several times to measure how much of the critical path it is
            if (s < 7) for (int rep = 0; rep < SBM_CT_PRIMAL_REPEAT; ++rep) sbm_ct_primal(sh, warp, lane, s + 1, h, (s + 1) & 1);
#else
            if (s < 7) sbm_ct_primal(sh, warp, lane, s + 1, h, (s + 1) & 1);
#endif
            else {
                sh.sqb[warp * 32 + lane] = sq;
                if (anyinit) { sh.park[warp * 32 + lane] = d0; sh.park[(W + warp) * 32 + lane] = d1; }
            }
#ifdef SBM_CT_PROF
            cbusy[s] += clock64() - ctprev;
#endif
            SBM_CT_BAR(NT);
#ifdef SBM_CT_PROF
            ctprev = sbm_ct_clock_after_barrier(sh.cta);
#endif
        }
    }
#ifdef SBM_CT_PROF
    if (blockIdx.x == 0 && (warp == SBM_CT_PROF || warp == W - 1) && lane == 0)
        printf("SBM_CT_PROF column warp %d busy: begin %lld s1 %lld s2 %lld s3 %lld s4 %lld s5 %lld s6 %lld s7 %lld\n", warp,
               cbusy[0] / cbusy[8], cbusy[1] / cbusy[8], cbusy[2] / cbusy[8], cbusy[3] / cbusy[8], cbusy[4] / cbusy[8], cbusy[5] / cbusy[8],
               cbusy[6] / cbusy[8], cbusy[7] / cbusy[8]);
#endif
}

// a warp without columns: evaluates its share of the model, runs the step controller
template <int W, int C, int XW>
__device__ __forceinline__ void sbm_ct_model_warp(const SbmSolveArgs& a, const SbmCtShared& sh, const int warp, const int lane)
{
    constexpr int NT = (W + XW) * 32;
#ifdef SBM_CT_PROF
    // [0] controller, [s] iteration s (columns: stage s; model: input of stage s+1), [8] steps
    long long prof[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, busy[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define SBM_CT_MARK(i) { const long long tnow = sbm_ct_clock_after_barrier(sh.cta); prof[i] += tnow - tprev; tprev = tnow; }
#define SBM_CT_BUSY(i) { busy[i] += clock64() - tprev; }
#else
#define SBM_CT_MARK(i)
#define SBM_CT_BUSY(i)
#endif
    SbmCtCarry cy;
#pragma unroll
    for (int i = 0; i < SBM_N; ++i) { cy.ynew[i] = 0.0; cy.f7[i] = 0.0; }
    cy.sqs = cy.q0 = cy.q1 = 0.0;
    for (;;) {
        if (warp == W) sbm_ct_control<W, C, SBM_NF - (W - 1) * C, 2>(a, sh, lane, cy);
        SBM_CT_BUSY(0)
        SBM_CT_BAR(NT);
        SBM_CT_MARK(0)
        if (!(sh.cta[0] & SBM_CT_ALIVE)) break;
        const double h = sh.hb[lane];
#ifdef SBM_CT_PROF
        prof[8] += 1;
#endif
#pragma unroll 1
        for (int s = 1; s <= 7; ++s) {
#if SBM_CT_JTM
            // J at the input of stage 2 is written by several warps (head shares): copied once they have all arrived
            if (s == 2 && warp == W && lane == 0)
                sbm_ct_j_copy(sh.Jb, sh.tmem_base + sh.jt_base, sh.jbar);
#endif
            if (s < 6) sbm_ct_prepare<W + XW>(sh, warp, lane, s + 2, h);
            if (s < 7) sbm_ct_primal(sh, warp, lane, s + 1, h, (s + 1) & 1);
#if SBM_CT_JTM
            if (s >= 2 && s < 7 && warp == W) {
                __syncwarp();
                if (lane == 0)
                    sbm_ct_j_copy(sh.Jb + ((s + 1) & 1) * SBM_NJS * 32, sh.tmem_base + sh.jt_base + (unsigned)((s + 1) & 1) * (2u * SBM_NJP),
                                  sh.jbar + ((s + 1) & 1));
            }
#endif
            if (s == 7 && warp == W) {
                // while the column warps finish stage 7: the state column's part of the controller's input
                cy.sqs = cy.q0 = cy.q1 = 0.0;
                sbm_ct_state_column(sh, lane, h, a.rtol, a.atol, cy.ynew, cy.f7, cy.sqs, cy.q0, cy.q1);
#if !SBM_CT_NO_STAGE_ROWS
                // a clipped step ends on an output time: prepare its rows now, off the controller's critical path
                if (sh.ci[SBM_CT_I_MODE * 32 + lane] == SBM_CT_NORMAL && (sh.ci[SBM_CT_I_BITS * 32 + lane] & 2))
                    sbm_ct_stage_rows(a, sh, lane, cy.ynew);
#endif
            }
            SBM_CT_BUSY(s)
            SBM_CT_BAR(NT);
            SBM_CT_MARK(s)
        }
    }
#ifdef SBM_CT_PROF
    if (blockIdx.x == 0 && warp == W && lane == 0) {
        printf("SBM_CT_PROF steps %lld clocks/step: ctl %lld s1 %lld s2 %lld s3 %lld s4 %lld s5 %lld s6 %lld s7 %lld\n", prof[8],
               prof[0] / prof[8], prof[1] / prof[8], prof[2] / prof[8], prof[3] / prof[8], prof[4] / prof[8],
               prof[5] / prof[8], prof[6] / prof[8], prof[7] / prof[8]);
        printf("SBM_CT_PROF model warp busy: ctl %lld s1 %lld s2 %lld s3 %lld s4 %lld s5 %lld s6 %lld s7 %lld\n",
               busy[0] / prof[8], busy[1] / prof[8], busy[2] / prof[8], busy[3] / prof[8], busy[4] / prof[8],
               busy[5] / prof[8], busy[6] / prof[8], busy[7] / prof[8]);
    }
#endif
}

// -------------------------------------------------------------------------------------------------
// SBM_CT_ASYNC: producer / consumer hand-over of J and G through mbarriers instead of one CTA barrier per stage.
// The model warp (it evaluates the whole model: one share) runs up to NB - 1 stages ahead of the slowest column
// warp; a column warp waits only for the data of its next stage, never for the other column warps, so the warps
// of a scheduler drift apart and stop competing for the FP64 pipe in the same phase of a stage.  Two CTA-wide
// hand-shakes per step remain: "done" (all error-norm partial sums written) and "go" (controller output published).
// -------------------------------------------------------------------------------------------------
#ifndef SBM_CT_NB
#define SBM_CT_NB 3
#endif

__device__ __forceinline__ void sbm_mb_init(unsigned long long* mb, const int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void sbm_mb_arrive(unsigned long long* mb)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n}"
                 :: "r"((unsigned)__cvta_generic_to_shared(mb)) : "memory");
}
__device__ __forceinline__ void sbm_mb_wait(unsigned long long* mb, const unsigned parity)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(mb);
    unsigned ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();      // a protocol error must not hang the GPU
    }
}

template <int W, int C, int CW>
__device__ __forceinline__ void sbm_ct_warp_async(const SbmSolveArgs& a, const SbmCtShared& sh, const int warp, const int lane)
{
    constexpr int N = SBM_N, NE = N + (N & 1), NB = SBM_CT_NB;
    const int col0 = warp * C;
    typedef SbmCtStoreTmem ST;
    ST st;
    st.t1 = sh.tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * 4 * C * NE);
    st.t4 = st.t1 + 2u * CW * NE;
    double y[CW][N], s2[CW][N], s3[CW][N];
#pragma unroll
    for (int c = 0; c < CW; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) { y[c][i] = 0.0; s2[c][i] = 0.0; s3[c][i] = 0.0; }
#pragma unroll
    for (int c = 0; c < CW; ++c)
#pragma unroll
        for (int ip = 0; ip < NE / 2; ++ip) { st.store(1, c * NE + 2 * ip, 0.0, 0.0); st.store(4, c * NE + 2 * ip, 0.0, 0.0); }
    st.fence();
    const double rtol = a.rtol, atol = a.atol;
    unsigned long long* full = sh.mbar;
    unsigned long long* empty = sh.mbar + NB;
    unsigned long long* done = sh.mbar + 2 * NB;
    unsigned long long* go = sh.mbar + 2 * NB + 1;
    unsigned gc = 0, step = 0;                    // stages consumed, steps started
    __syncwarp();
    if (lane == 0) sbm_mb_arrive(done);            // "step -1 finished": lets the controller start

    for (;;) {
        sbm_mb_wait(go, step & 1u);                // controller output of this step published
        ++step;
        const int cflags = sh.cta[0];
        if (!(cflags & SBM_CT_ALIVE)) break;
        const int info = sh.fl[lane];
        const double h = sh.hb[lane];
        sbm_ct_cols_begin<CW, ST>(a, sh, info, cflags, h, col0, lane, y, s3, st);
        const bool anyinit = (cflags & SBM_CT_ANYINIT) != 0;
        double sq = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll 1
        for (int s = 2; s <= 7; ++s) {
            const unsigned b = gc % NB;
            sbm_mb_wait(full + b, (gc / NB) & 1u); // J, G of this stage are in buffer b
            const double* Jp = sh.Jb + b * SBM_NJ * 32 + lane;
            const double* Gw = sh.Gd + (b * SBM_NF + col0) * NE * 32 + 2 * lane;
            sbm_ct_cols_stage<CW, ST>(s, info, h, Jp, 0u, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1);
            __syncwarp();
            if (lane == 0) sbm_mb_arrive(empty + b);   // this warp is done with buffer b
            ++gc;
        }
        sh.sqb[warp * 32 + lane] = sq;
        if (anyinit) { sh.park[warp * 32 + lane] = d0; sh.park[(W + warp) * 32 + lane] = d1; }
        __syncwarp();
        if (lane == 0) sbm_mb_arrive(done);
    }
}

template <int W, int C>
__device__ __forceinline__ void sbm_ct_model_warp_async(const SbmSolveArgs& a, const SbmCtShared& sh, const int lane)
{
    constexpr int NB = SBM_CT_NB;
    unsigned long long* full = sh.mbar;
    unsigned long long* empty = sh.mbar + NB;
    unsigned long long* done = sh.mbar + 2 * NB;
    unsigned long long* go = sh.mbar + 2 * NB + 1;
    unsigned gp = 0, step = 0;                    // stages produced, steps finished by the column warps
    for (;;) {
        sbm_mb_wait(done, step & 1u);              // all partial sums of the previous step are written
        ++step;
        { SbmCtCarry none; sbm_ct_control<W, C, SBM_NF - (W - 1) * C, 1>(a, sh, lane, none); }
        __syncwarp();
        const bool alive = (sh.cta[0] & SBM_CT_ALIVE) != 0;
        if (lane == 0) sbm_mb_arrive(go);
        if (!alive) break;
        const double h = sh.hb[lane];
#pragma unroll 1
        for (int S = 2; S <= 7; ++S) {
            const unsigned b = gp % NB;
            if (gp >= (unsigned)NB) sbm_mb_wait(empty + b, ((gp / NB) - 1u) & 1u);   // every column warp has left buffer b
#ifndef SBM_WHATIF_ASYNC_NOMODEL   /* experiment: how fast are the column warps when the model costs nothing (results are wrong) */
            if (S >= 3) sbm_ct_prepare<1>(sh, 0, lane, S, h);
            sbm_ct_primal(sh, W, lane, S, h, (int)b);
#endif
            __syncwarp();
            if (lane == 0) sbm_mb_arrive(full + b);
            ++gp;
        }
        // the state column's part of the controller's input (f_7 is this warp's own)
        double ynew[SBM_N], f7[SBM_N], sqs = 0.0, q0 = 0.0, q1 = 0.0;
        sbm_ct_state_column(sh, lane, h, a.rtol, a.atol, ynew, f7, sqs, q0, q1);
#pragma unroll
        for (int i = 0; i < SBM_N; ++i) sh.ynb[i * 32 + lane] = ynew[i];
        sh.sqb[(W + 0) * 32 + lane] = sqs;
        sh.sqb[(W + 1) * 32 + lane] = q0;
        sh.sqb[(W + 2) * 32 + lane] = q1;
    }
}

// =================================================================================================================
// SBM_CT_LAG: lagged step-size control - the controller and the head of a step leave the critical path.
//
// In the synchronous schedule above a step is EIGHT barrier intervals: controller | head (model at the input of stage 2) |
// stages 2..7, and in the first two (19 % of a step, measured) the column warps have nothing to do, because the step
// size of step k+1 is only known once the error of step k is.  Here the step size of step k+1 is fixed BEFORE the error
// of step k is known - from the error of step k-1 (an I-controller on a one-step-old error constant; measured on the
// C4 distribution with a Python model of the kernel's controller: 3,051 steps per solve and 1.9 % rejections against
// 3,162 steps and 0.3 % with the PI controller) - and the error of step k only decides whether step k is accepted:
//
//   interval 7 of step k    columns: stage 7 (k7, error norm)   model warp: state column of step k, set-up of step k+1
//                                                                ASSUMING step k is accepted, model at the input of
//                                                                stage 2 of step k+1 (speculative)
//   interval 2 of step k+1  columns: every warp sums the partial error norms itself and decides (same arithmetic in
//                           every warp); accepted: y <- Y7, rows of an output time, stage 2.  model warp: the same
//                           decision, then the bookkeeping of step k (commit / reject, FSAL, step size of step k+2, output
//                           values, trajectory queue), model at the input of stage 3
//   intervals 3..6          as before
//
// A step is SIX barrier intervals.  A rejected lane turns step k+1 into the zero-length refresh step the push form needs
// anyway (k1 of its columns was destroyed in stage 4): h = 0 for that lane; what the model warp evaluated speculatively
// for its stage 2 belongs to the wrong state, so k2 of that lane is cleared (it only ever meets a zero coefficient).
// The model warp owns the state column: it forms every stage value itself (the column warps no longer prepare them).
// =================================================================================================================
#if SBM_CT_LAG

#define SBM_CTL_F_NORMAL 16           /* sh.fl: the step in flight is a normal step: its error decides */

// controller state (per lane, shared memory, model warp only): [cur] describes the step in flight, [spec] what the
// set-up computed for the next step assuming the one in flight is accepted
#define SBM_CTL_D_T       0
#define SBM_CTL_D_TK      1
#define SBM_CTL_D_TEND    2
#define SBM_CTL_D_HCUR    3
#define SBM_CTL_D_HPROP   4           /* cruise step before clipping of the step in flight */
#define SBM_CTL_D_HCRUISE 5
#define SBM_CTL_D_FACOLD  6
#define SBM_CTL_D_TS      7
#define SBM_CTL_D_TKS     8
#define SBM_CTL_D_HPROPS  9
#define SBM_CTL_I_KT      0
#define SBM_CTL_I_KT1     1
#define SBM_CTL_I_STATUS  2
#define SBM_CTL_I_MODE    3
#define SBM_CTL_I_BITS    4           /* 1 lastrej, 2 clipped, 8 initf, 16 exhausted */
#define SBM_CTL_I_NSTEP   5
#define SBM_CTL_I_KTS     6
#define SBM_CTL_I_MODES   7
#define SBM_CTL_I_BITSS   8
#define SBM_CTL_I_STATUSS 9

// err^2 of the step in flight from the partial sums of the column warps and of the state column: EVERY warp runs this
// (same loads, same order of additions), so every warp takes the same decision without a hand-over
template <int W>
__device__ __forceinline__ bool sbm_ctl_accepted(const SbmCtShared& sh, const int lane, const int flags, double& err2)
{
    double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int w = 0; w < W; ++w) part[w & 3] += sh.sqb[w * 32 + lane];
    const double sq = ((part[0] + part[1]) + (part[2] + part[3])) + sh.sqb[W * 32 + lane];
    err2 = sq * (1.0 / (double)(SBM_N * (SBM_NF + 1)));
    return !(flags & SBM_CTL_F_NORMAL) || err2 <= 1.0;                 // false for NaN
}

template <int W>
__device__ __forceinline__ void sbm_ctl_book_b(const SbmSolveArgs& a, const SbmCtShared& sh, const int lane);

// ---- column warp ------------------------------------------------------------------------------------------------
template <int W, int C, int CW, int XW>
__device__ __forceinline__ void sbm_ctl_warp(const SbmSolveArgs& a, const SbmCtShared& sh, const int warp, const int lane)
{
    constexpr int N = SBM_N, NT = (W + XW) * 32, NE = N + (N & 1);
    const int col0 = warp * C;
    typedef SbmCtStoreTmem ST;
    ST st;
    st.t1 = sh.tmem_base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)((warp >> 2) * 4 * C * NE);
    st.t4 = st.t1 + 2u * CW * NE;
    double y[CW][N], s2[CW][N], s3[CW][N];
#pragma unroll
    for (int c = 0; c < CW; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) { y[c][i] = 0.0; s2[c][i] = 0.0; s3[c][i] = 0.0; }
#pragma unroll
    for (int c = 0; c < CW; ++c)
#pragma unroll
        for (int ip = 0; ip < NE / 2; ++ip) { st.store(1, c * NE + 2 * ip, 0.0, 0.0); st.store(4, c * NE + 2 * ip, 0.0, 0.0); }
    st.fence();
    const double rtol = a.rtol, atol = a.atol;
#ifdef SBM_CT_PROF
    long long cbusy[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ctprev = 0;
#endif
    SBM_CT_BAR(NT);                                // the first set-up is published
#ifdef SBM_CT_PROF
    ctprev = sbm_ct_clock_after_barrier(sh.cta);
#endif
    for (;;) {
        const int cflags = sh.cta[0];
        if (!(cflags & SBM_CT_ALIVE)) break;
        const int fl = sh.fl[lane];
        double err2;
        const bool acc = sbm_ctl_accepted<W>(sh, lane, fl, err2);
        const double h = acc ? sh.hb[lane] : 0.0;  // a rejected lane: zero-length refresh step
        int info = fl & (SBM_CT_F_POISON | SBM_CT_F_RESET);
        if (acc && (fl & SBM_CTL_F_NORMAL)) info |= SBM_CT_F_ACCEPT;
        if (acc && (fl & SBM_CT_F_EMIT)) info |= SBM_CT_F_EMIT;
        sbm_ct_cols_begin<CW, ST>(a, sh, info, cflags, h, col0, lane, y, s3, st);
#ifdef SBM_CT_PROF
        cbusy[0] += clock64() - ctprev;
        cbusy[8] += 1;
#endif
        const bool anyinit = (cflags & SBM_CT_ANYINIT) != 0;
        double sq = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll 1
        for (int s = 2; s <= 7; ++s) {
            const double* Jp = sh.Jb + (s & 1) * SBM_NJS * 32 + lane;
            const double* Gw = sh.Gd + ((s & 1) * SBM_NF + col0) * NE * 32 + 2 * lane;
            sbm_ct_cols_stage<CW, ST>(s, info, h, Jp, 0u, Gw, y, s2, s3, st, rtol, atol, anyinit, sq, d0, d1);
            if (s == 2 && !acc) {
                // the model at the input of stage 2 was evaluated for the trial state of the rejected step: k2 is garbage
#pragma unroll
                for (int c = 0; c < CW; ++c)
#pragma unroll
                    for (int i = 0; i < N; ++i) s2[c][i] = 0.0;
            }
            if (s == 7) {
                sh.sqb[warp * 32 + lane] = sq;
                if (anyinit) { sh.park[warp * 32 + lane] = d0; sh.park[(W + warp) * 32 + lane] = d1; }
            }
            if (s == 3 && warp == SBM_CTL_BOOK_WARP) sbm_ctl_book_b<W>(a, sh, lane);
#ifdef SBM_CT_PROF
            cbusy[s] += clock64() - ctprev;
#endif
            SBM_CT_BAR(NT);
#ifdef SBM_CT_PROF
            ctprev = sbm_ct_clock_after_barrier(sh.cta);
#endif
        }
    }
#ifdef SBM_CT_PROF
    if (blockIdx.x == 0 && (warp == SBM_CT_PROF || warp == W - 1) && lane == 0)
        printf("SBM_CT_PROF column warp %d busy: begin %lld s2 %lld s3 %lld s4 %lld s5 %lld s6 %lld s7 %lld\n", warp,
               cbusy[0] / cbusy[8], cbusy[2] / cbusy[8], cbusy[3] / cbusy[8], cbusy[4] / cbusy[8], cbusy[5] / cbusy[8],
               cbusy[6] / cbusy[8], cbusy[7] / cbusy[8]);
#endif
}

// ---- model warp ---------------------------------------------------------------------------------------------------
// stage value Y_S = y + h sum_j a_Sj f_j of the state column, S = 3..7, in registers: the model reads it from there (through
// shared memory the hand-over cost a store, a load and their latency at the head of the critical chain of every evaluation)
__device__ __forceinline__ void sbm_ctl_form_y(const SbmCtShared& sh, const int lane, const int S, const double h,
                                               double (&Y)[SBM_N])
{
    double cf[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) cf[j] = sbm_dp_a[S][j];
#pragma unroll
    for (int i = 0; i < SBM_N; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) v = fma(cf[j], sh.fh[(j * SBM_N + i) * 32 + lane], v);
        Y[i] = fma(h, v, sh.yb[i * 32 + lane]);
    }
}

// the model at the stage value Y -> J, G of stage S (buffer S & 1), f_S
__device__ __forceinline__ void sbm_ctl_model(const SbmCtShared& sh, const int share, const int lane, const int S, const double ts,
                                              const double (&Y)[SBM_N])
{
    const int b = S & 1;
    SBM_CT_JIT(16u + S);
    sbm_primal_share<32, 1>(share, ts, Y, Y, 0.0, sh.pb + lane, sh.Jb + b * SBM_NJS * 32 + lane,
                            sh.Gd + b * SBM_NF * (SBM_N + (SBM_N & 1)) * 32 + 2 * lane, sh.fh + (S - 1) * SBM_N * 32 + lane);
}

struct SbmCtlRegs {
    double ynew[SBM_N], f7[SBM_N];     // trial state and its derivative (the FSAL k1 of the next step) of the step in flight
    double q0, q1;                     // start-up norms of the state column
};

// set-up of the next step, assuming the step in flight is accepted; publishes hb / fl / cta flags for the column warps
template <int W>
__device__ __forceinline__ void sbm_ctl_setup(const SbmSolveArgs& a, const SbmCtShared& sh, const int lane, SbmCtlRegs& r)
{
    constexpr int N = SBM_N;
    constexpr unsigned FULL = 0xffffffffu;
    const long long ntraj = a.B * (long long)a.E;
    unsigned long long* queue = a.counters + SBM_CNT_QUEUE;
    const double t = sh.cd[SBM_CTL_D_T * 32 + lane], hcur = sh.cd[SBM_CTL_D_HCUR * 32 + lane];
    const double tk = sh.cd[SBM_CTL_D_TK * 32 + lane], hcruise = sh.cd[SBM_CTL_D_HCRUISE * 32 + lane];
    const int kt = sh.ci[SBM_CTL_I_KT * 32 + lane], kt1 = sh.ci[SBM_CTL_I_KT1 * 32 + lane];
    const int status = sh.ci[SBM_CTL_I_STATUS * 32 + lane], mode = sh.ci[SBM_CTL_I_MODE * 32 + lane];
    const int bits = sh.ci[SBM_CTL_I_BITS * 32 + lane];
    const bool clipped = bits & 2, initf = bits & 8;
    bool exhausted = bits & 16;

    double ts = t, tks = tk, hs = 0.0, hprops = 0.0;
    int kts = kt, modes = mode, statuss = status, flags = 0;
    bool clippeds = false, initfs = false, fetched = false, landed = false;
    long long traj = sh.trj[lane];
    if (mode == SBM_CT_NORMAL) {
        flags |= SBM_CTL_F_NORMAL;
        ts = clipped ? tk : t + hcur;
        landed = clipped;
    } else if (mode == SBM_CT_ZERO) {
        if (status != SBM_OK) landed = true;        // failed trajectory: flush the remaining rows as NaN, one time per step
        else if (initf) modes = SBM_CT_ZERO;        // the step after the start-up step is a filler: h0 arrives one step late
        else modes = SBM_CT_NORMAL;
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
            fetched = true;
            const int e = (int)(traj % a.E);
            const long long b = traj / a.E;
#pragma unroll
            for (int i = 0; i < N; ++i) { sh.yb[i * 32 + lane] = a.y0[e * N + i]; r.ynew[i] = a.y0[e * N + i]; r.f7[i] = 0.0; }
            for (int q = 0; q < 7 * N; ++q) sh.fh[q * 32 + lane] = 0.0;
            sh.trj[lane] = traj;
            sh.trow[lane] = b * a.n_rows;
            sh.ci[SBM_CTL_I_KT1 * 32 + lane] = a.tout_off[e + 1];
            sh.ci[SBM_CTL_I_NSTEP * 32 + lane] = 0;
            sh.cd[SBM_CTL_D_FACOLD * 32 + lane] = 1e-8;             // (1e-4)^2: the controller keeps squared errors
            kts = a.tout_off[e];
            ts = 0.0;
            statuss = SBM_OK;
            if (a.tout_off[e + 1] > kts) {
                tks = a.tout[kts];
                sh.cd[SBM_CTL_D_TEND * 32 + lane] = a.tout[a.tout_off[e + 1] - 1];
                modes = SBM_CT_ZERO;
                initfs = true;
                flags |= SBM_CT_F_RESET;
            } else {
                a.status[traj] = SBM_OK;                            // no output times: nothing to integrate
                sh.cnt[3 * 32 + lane] += 1;
            }
        } else {
            exhausted = true;
        }
    }
    // ---- model parameters of the new trajectories: exp(theta) for optimised, the raw value for fixed parameters
    //      (base_project.py:343-363), gathered by the whole warp ----
    for (unsigned need = __ballot_sync(FULL, fetched); need; need &= need - 1u) {
        const int L = __ffs(need) - 1;
        const long long trajL = __shfl_sync(FULL, traj, L);
        const int e = (int)(trajL % a.E);
        const long long b = trajL / a.E;
        for (int m = lane; m < SBM_NP; m += 32) {
            const int c = a.pmap[e * SBM_NP + m];
            sh.pb[m * 32 + L] = (c >= 0) ? exp(a.theta[b * a.theta_sb + c * a.theta_sp]) : a.pfix[e * SBM_NP + m];
            sh.pcb[m * 32 + L] = c;
        }
    }
    // ---- an output time is reached (or flushed): stage its rows; the trajectory may end there ----
    if (landed) {
        flags |= SBM_CT_F_EMIT;
        if (status != SBM_OK) flags |= SBM_CT_F_POISON;
        const long long rowbase = sh.trow[lane];
        const int r0 = a.orow_off[kt], r1 = a.orow_off[kt + 1];
        sh.em_n[lane] = r1 - r0;
        sh.ktb[lane] = kt;
        for (int q = r0; q < r1 && q - r0 < SBM_CT_EM; ++q) {
            const unsigned long long mask = a.row_mask[q];
            sh.em_mask[(q - r0) * 32 + lane] = mask;
            sh.em_off[(q - r0) * 32 + lane] = (rowbase + a.row_jac[q]) * (long long)a.n_proj;
            sh.em_soff[(q - r0) * 32 + lane] = rowbase + a.row_sim[q];
        }
        kts = kt + 1;
        if (kts >= kt1) modes = SBM_CT_IDLE;
        else tks = a.tout[kts];
    }
    // ---- step size of a normal next step: the cruise step is one step old (lagged), clipped onto the output time ----
    if (modes == SBM_CT_NORMAL && statuss == SBM_OK) {
        double h = hcruise;
        hprops = h;
        if (ts + 1.01 * h >= tks) { h = tks - ts; clippeds = true; }
        if ((unsigned)sh.ci[SBM_CTL_I_NSTEP * 32 + lane] >= (unsigned)a.max_steps) statuss = SBM_ERR_MAXSTEPS;
        else if (!clippeds && !(h > 1e-15 * fmax(ts, 1e-300))) statuss = SBM_ERR_UNDERFLOW;
        if (statuss == SBM_OK) hs = h;
        else { modes = SBM_CT_ZERO; clippeds = false; }
    } else if (modes == SBM_CT_NORMAL) {
        modes = SBM_CT_ZERO;                       // failed: zero-length steps that flush the rows
    }
    sh.hb[lane] = hs;
    sh.fl[lane] = flags;
    sh.cd[SBM_CTL_D_TS * 32 + lane] = ts;
    sh.cd[SBM_CTL_D_TKS * 32 + lane] = tks;
    sh.cd[SBM_CTL_D_HPROPS * 32 + lane] = hprops;
    sh.ci[SBM_CTL_I_KTS * 32 + lane] = kts;
    sh.ci[SBM_CTL_I_MODES * 32 + lane] = modes;
    sh.ci[SBM_CTL_I_STATUSS * 32 + lane] = statuss;
    sh.ci[SBM_CTL_I_BITSS * 32 + lane] = (clippeds ? 2 : 0) | (initfs ? 8 : 0) | (exhausted ? 16 : 0);
    sh.ci[SBM_CTL_I_BITS * 32 + lane] = (bits & ~16) | (exhausted ? 16 : 0);
    // a lane whose step in flight is a normal step is alive whatever the set-up assumed (the step may be rejected)
    const bool alive = __any_sync(FULL, mode != SBM_CT_IDLE || modes != SBM_CT_IDLE) || !__all_sync(FULL, exhausted);
    const bool anyemit = __any_sync(FULL, (flags & SBM_CT_F_EMIT) != 0);
    const bool anyinit = __any_sync(FULL, initfs);
    if (lane == 0)
        sh.cta[0] = (alive ? SBM_CT_ALIVE : 0) | (anyemit ? SBM_CT_ANYEMIT : 0) | (anyinit ? SBM_CT_ANYINIT : 0);
}

// Bookkeeping of the step that has just finished (its error is known now), in two parts.  Part A is what the model warp
// needs before it can evaluate the model at the input of stage 3: the decision, the state column's commit (or the clean-up
// after a rejection), the step size of the step in flight.  Part B - the controller arithmetic, the output values, the
// trajectory's bookkeeping - is only needed by the next set-up: column warp 0 runs it in interval 3, where the column
// warps wait for the model warp anyway (measured: 1.4 k clocks of column work against 2.4 k of the model warp).
template <int W>
__device__ __forceinline__ double sbm_ctl_book_a(const SbmCtShared& sh, const int lane, const SbmCtlRegs& r, double& t)
{
    constexpr int N = SBM_N;
    const int flags = sh.fl[lane];
    double err2;
    const bool acc = sbm_ctl_accepted<W>(sh, lane, flags, err2);
    const int mode = sh.ci[SBM_CTL_I_MODE * 32 + lane];
    if (mode == SBM_CT_NORMAL && !acc) {
        // the stage derivatives of the rejected trial (and of the speculative stage 2) may be non-finite: they must not
        // meet the zero coefficients of the refresh step (0 * Inf)
#pragma unroll
        for (int j = 1; j < 7; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) sh.fh[(j * N + i) * 32 + lane] = 0.0;
    } else if (mode != SBM_CT_IDLE) {
        // accepted step: y <- trial state; accepted or zero-length step: FSAL, k1 <- k7
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (mode == SBM_CT_NORMAL) sh.yb[i * 32 + lane] = r.ynew[i];
            sh.fh[i * 32 + lane] = r.f7[i];
        }
    }
    t = acc ? sh.cd[SBM_CTL_D_TS * 32 + lane] : sh.cd[SBM_CTL_D_T * 32 + lane];
    return acc ? sh.hb[lane] : 0.0;
}

template <int W>
__device__ __forceinline__ void sbm_ctl_book_b(const SbmSolveArgs& a, const SbmCtShared& sh, const int lane)
{
    constexpr int N = SBM_N;
    const int flags = sh.fl[lane];
    double err2;
    const bool acc = sbm_ctl_accepted<W>(sh, lane, flags, err2);
    const int mode = sh.ci[SBM_CTL_I_MODE * 32 + lane];
    int bits = sh.ci[SBM_CTL_I_BITS * 32 + lane];
    int status = sh.ci[SBM_CTL_I_STATUS * 32 + lane];
    const double hcur = sh.cd[SBM_CTL_D_HCUR * 32 + lane];
    bool commit = true;
    if (mode == SBM_CT_NORMAL) {
        // I-controller on the error of the step just finished; the result is the cruise step of the step AFTER the one
        // in flight.  facold holds the SQUARE of the previous accepted error (Hairer's PI term, beta = 0.04).
        float facold = (float)sh.cd[SBM_CTL_D_FACOLD * 32 + lane];
        const bool finite = (err2 == err2) && (err2 < 1e300);
        const float ferr2 = fmaxf((float)err2, 1e-37f);
        const float e17 = __powf(ferr2, 0.5f * (0.2f - 0.04f * 0.75f));
        float ra = fminf(SBM_CTL_FACMAX, fmaxf(0.2f, SBM_CTL_SAFETY * __fdividef(__powf(facold, 0.5f * 0.04f), e17)));
        if (bits & 1) ra = fminf(ra, 1.0f);
        const float rr = finite ? fmaxf(0.2f, __fdividef(SBM_CTL_SAFETY, e17)) : 0.1f;
        double hnew = hcur * (double)(acc ? ra : rr);
        const double hprop = sh.cd[SBM_CTL_D_HPROP * 32 + lane];
        if (acc && (bits & 2) && hnew < hprop) hnew = hprop;         // a clipped step must not shrink the cruise step
        sh.cd[SBM_CTL_D_HCRUISE * 32 + lane] = hnew;
        sh.ci[SBM_CTL_I_NSTEP * 32 + lane] += 1;
        sh.cnt[0 * 32 + lane] += 1;
        sh.cnt[2 * 32 + lane] += acc ? 6 : 7;        // after a rejection k1 of the sensitivity columns is re-evaluated
        if (acc) {
            sh.cd[SBM_CTL_D_FACOLD * 32 + lane] = (double)fmaxf(ferr2, 1e-8f);
            bits &= ~1;
        } else {
            commit = false;
            sh.cnt[1 * 32 + lane] += 1;
            bits = (bits | 1) & ~(2 | 8);
            if (!finite && hnew < 1e-300) status = SBM_ERR_NONFINITE;
            sh.ci[SBM_CTL_I_MODE * 32 + lane] = SBM_CT_ZERO;
            sh.ci[SBM_CTL_I_BITS * 32 + lane] = bits;
            sh.ci[SBM_CTL_I_STATUS * 32 + lane] = status;
            sh.cd[SBM_CTL_D_HCUR * 32 + lane] = 0.0;
        }
    } else if (mode == SBM_CT_ZERO && (bits & 8)) {
        // initial step size (Hairer, Norsett, Wanner II.4): h0 = 0.01 |y|/|f|; start at the rule's upper bound 100 h0
        double q0 = sh.sqb[(W + 1) * 32 + lane], q1 = sh.sqb[(W + 2) * 32 + lane];
#pragma unroll
        for (int w = 0; w < W; ++w) { q0 += sh.park[w * 32 + lane]; q1 += sh.park[(W + w) * 32 + lane]; }
        const double inv_naug = 1.0 / (double)(N * (SBM_NF + 1));
        const double n0 = sqrt(q0 * inv_naug), n1 = sqrt(q1 * inv_naug);
        const double h0 = (n0 < 1e-5 || n1 < 1e-5) ? 1e-6 : 0.01 * n0 / n1;
        double h = fmin(100.0 * h0, sh.cd[SBM_CTL_D_TEND * 32 + lane]);
        if (a.h0 > 0.0) h = fmin(a.h0, sh.cd[SBM_CTL_D_TEND * 32 + lane]);
        sh.cd[SBM_CTL_D_HCRUISE * 32 + lane] = h;
        sh.cnt[2 * 32 + lane] += 1;
    }
    if (commit) {
        // the set-up's assumption holds: its results become the state of the step in flight
        if (flags & SBM_CT_F_EMIT) {
            const int n = sh.em_n[lane];
            for (int q = 0; q < n && q < SBM_CT_EM; ++q) a.sims[sh.em_soff[q * 32 + lane]] = sh.em_val[q * 32 + lane];
            if (n > SBM_CT_EM) {
                const long long rowbase = sh.trow[lane];
                const int r0 = a.orow_off[sh.ktb[lane]];
                for (int q = r0 + SBM_CT_EM; q < r0 + n; ++q) {
                    const unsigned long long mask = a.row_mask[q];
                    double v = 0.0;
#pragma unroll
                    for (int i = 0; i < N; ++i) v += ((mask >> i) & 1ull) ? sh.yb[i * 32 + lane] : 0.0;
                    if (flags & SBM_CT_F_POISON) v = __longlong_as_double(0x7ff8000000000000ll);
                    a.sims[rowbase + a.row_sim[q]] = v;
                }
            }
            if (sh.ci[SBM_CTL_I_MODES * 32 + lane] == SBM_CT_IDLE) {
                a.status[sh.trj[lane]] = status;
                sh.cnt[3 * 32 + lane] += 1;
            }
        }
        const int bs = sh.ci[SBM_CTL_I_BITSS * 32 + lane];
        sh.cd[SBM_CTL_D_T * 32 + lane] = sh.cd[SBM_CTL_D_TS * 32 + lane];
        sh.cd[SBM_CTL_D_TK * 32 + lane] = sh.cd[SBM_CTL_D_TKS * 32 + lane];
        sh.cd[SBM_CTL_D_HPROP * 32 + lane] = sh.cd[SBM_CTL_D_HPROPS * 32 + lane];
        sh.cd[SBM_CTL_D_HCUR * 32 + lane] = sh.hb[lane];
        sh.ci[SBM_CTL_I_KT * 32 + lane] = sh.ci[SBM_CTL_I_KTS * 32 + lane];
        sh.ci[SBM_CTL_I_MODE * 32 + lane] = sh.ci[SBM_CTL_I_MODES * 32 + lane];
        sh.ci[SBM_CTL_I_STATUS * 32 + lane] = sh.ci[SBM_CTL_I_STATUSS * 32 + lane];
        sh.ci[SBM_CTL_I_BITS * 32 + lane] = (bits & 1) | (bs & (2 | 8 | 16));
    }
}

// values of the rows of an output time the step in flight lands on, from its trial state (their masks were staged by the
// set-up, which runs before the trial state exists)
__device__ __forceinline__ void sbm_ctl_stage_values(const SbmCtShared& sh, const int lane, const SbmCtlRegs& r)
{
    const int flags = sh.fl[lane];
    if (flags & SBM_CT_F_EMIT) {
        const int n = sh.em_n[lane];
        for (int q = 0; q < n && q < SBM_CT_EM; ++q) {
            const unsigned long long mask = sh.em_mask[q * 32 + lane];
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < SBM_N; ++i) acc += ((mask >> i) & 1ull) ? r.ynew[i] : 0.0;                  // utils.py:10-26, 48-68
            if (flags & SBM_CT_F_POISON) acc = __longlong_as_double(0x7ff8000000000000ll);
            sh.em_val[q * 32 + lane] = acc;
        }
    }
}

template <int W, int C, int XW>
__device__ __forceinline__ void sbm_ctl_model_warp(const SbmSolveArgs& a, const SbmCtShared& sh, const int warp, const int lane)
{
    constexpr int N = SBM_N, NT = (W + XW) * 32;
    SbmCtlRegs r;
#pragma unroll
    for (int i = 0; i < N; ++i) { r.ynew[i] = 0.0; r.f7[i] = 0.0; }
    r.q0 = r.q1 = 0.0;
#ifdef SBM_CT_PROF
    long long prof[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, busy[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    // prologue: every lane is idle; fetch trajectories, publish the first step, model at the input of its stage 2
    sbm_ctl_setup<W>(a, sh, lane, r);
    __syncwarp();
    sbm_ctl_model(sh, warp, lane, 2, sh.cd[SBM_CTL_D_TS * 32 + lane], r.ynew);
    SBM_CT_BAR(NT);
#ifdef SBM_CT_PROF
    tprev = sbm_ct_clock_after_barrier(sh.cta);
#endif
    for (;;) {
        if (!(sh.cta[0] & SBM_CT_ALIVE)) break;
        // ---- interval 2: the error of the finished step is known ----
        double t;
        const double h = sbm_ctl_book_a<W>(sh, lane, r, t);
#ifdef SBM_CT_PROF
        prof[8] += 1;
        busy[0] += clock64() - tprev;
#endif
#pragma unroll 1
        for (int S = 3; S <= 7; ++S) {
            double Y[N];
            sbm_ctl_form_y(sh, lane, S, h, Y);
            if (S == 7) {
#pragma unroll
                for (int i = 0; i < N; ++i) r.ynew[i] = Y[i];
            }
            sbm_ctl_model(sh, warp, lane, S, fma(sbm_dp_c[S], h, t), Y);
            // interval 4 is bound by the column warps (the conversion of k1..k4): the set-up of the next step fits here -
            // column warp 0 finished the bookkeeping of the previous step in interval 3
            if (S == 5) sbm_ctl_setup<W>(a, sh, lane, r);
#ifdef SBM_CT_PROF
            busy[S - 1] += clock64() - tprev;
#endif
            SBM_CT_BAR(NT);
#ifdef SBM_CT_PROF
            { const long long tn = sbm_ct_clock_after_barrier(sh.cta); prof[S - 1] += tn - tprev; tprev = tn; }
#endif
        }
        // ---- interval 7: error of the state column, values of the rows of an output time, model at the input of stage 2 of
        //      the next step (speculative: assumes that the step in flight is accepted) ----
        {
            double yn[N], sqs = 0.0;
            r.q0 = r.q1 = 0.0;
            sbm_ct_state_column(sh, lane, h, a.rtol, a.atol, yn, r.f7, sqs, r.q0, r.q1);
            sh.sqb[W * 32 + lane] = sqs;
            sh.sqb[(W + 1) * 32 + lane] = r.q0;
            sh.sqb[(W + 2) * 32 + lane] = r.q1;
        }
        sbm_ctl_stage_values(sh, lane, r);
        {
            const double hs = sh.hb[lane];
            // (a lane that has just fetched starts from y0: its committed state, which the stage value of stage 7 repeats)
            double Y[N];
#pragma unroll
            for (int i = 0; i < N; ++i) Y[i] = fma(hs * DP_A21, r.f7[i], r.ynew[i]);
            sbm_ctl_model(sh, warp, lane, 2, fma(DP_C2, hs, sh.cd[SBM_CTL_D_TS * 32 + lane]), Y);
        }
#ifdef SBM_CT_PROF
        busy[7] += clock64() - tprev;
#endif
        SBM_CT_BAR(NT);
#ifdef SBM_CT_PROF
        { const long long tn = sbm_ct_clock_after_barrier(sh.cta); prof[7] += tn - tprev; tprev = tn; }
#endif
    }
#ifdef SBM_CT_PROF
    if (blockIdx.x == 0 && lane == 0) {
        printf("SBM_CT_PROF steps %lld clocks/step: s2 %lld s3 %lld s4 %lld s5 %lld s6 %lld s7 %lld\n", prof[8],
               prof[2] / prof[8], prof[3] / prof[8], prof[4] / prof[8], prof[5] / prof[8], prof[6] / prof[8], prof[7] / prof[8]);
        printf("SBM_CT_PROF model warp busy: book %lld s2 %lld s3 %lld s4 %lld s5 %lld s6 %lld s7 %lld\n", busy[0] / prof[8],
               busy[2] / prof[8], busy[3] / prof[8], busy[4] / prof[8], busy[5] / prof[8], busy[6] / prof[8], busy[7] / prof[8]);
    }
#endif
}
#endif  /* SBM_CT_LAG */

// W column warps per CTA (warps 0..W-2 carry C sensitivity columns each, warp W-1 the remaining
// CL = NF - (W-1)*C) plus XW warps without columns.  Shared-memory carve-up in doubles (sbm_api.cu computes the same
// size on the host):
//   2*NJ*32 + 2*NF*NE*32 + 7*N*32 + 2*N*32 + 2*N*32 + NP*32 + (W+3)*32 + 32 + ND*32 + 4*32 + 32 + (NI*32 + 3*32)/2
//   + 2*W*32 + 16  (S1 and S4 of the sensitivity columns live in tensor memory; 16 = mbarriers)
// with NBUF = 2 buffers of J and G (SBM_CT_NB = 3 with the asynchronous hand-over) in place of the factor 2 above
template <int W, int C, int XW>
__device__ __forceinline__ void sbm_dp5_ct_impl(const SbmSolveArgs& a)
{
    constexpr int N = SBM_N;
    constexpr int CL = SBM_NF - (W - 1) * C;
    constexpr int WT = W + XW;
    static_assert(CL >= 1 && CL <= C, "column split does not cover the free parameters");
    extern __shared__ double sbm_dyn[];
    SbmCtShared sh;
    sh.Jb = sbm_dyn;
#if SBM_CT_ASYNC
    constexpr int NBUF = SBM_CT_NB;
#else
    constexpr int NBUF = 2;
#endif
    sh.Gd = sh.Jb + NBUF * SBM_NJS * 32;
    sh.fh = sh.Gd + NBUF * SBM_NF * (N + (N & 1)) * 32;
    sh.yb = sh.fh + 7 * N * 32;
    sh.ynb = sh.yb + N * 32;
    sh.Pb = sh.ynb + N * 32;
    sh.pb = sh.Pb + 2 * N * 32;
    sh.sqb = sh.pb + SBM_NP * 32;
    sh.hb = sh.sqb + (W + 3) * 32;
    sh.cd = sh.hb + 32;
    sh.cnt = (unsigned long long*)(sh.cd + SBM_CT_ND * 32);
    sh.trj = (long long*)(sh.cnt + 4 * 32);
    sh.ci = (int*)(sh.trj + 32);
    sh.fl = sh.ci + SBM_CT_NI * 32;
    sh.ktb = sh.fl + 32;
    sh.cta = sh.ktb + 32;
    sh.park = (double*)(sh.cta + 32);
    sh.mbar = (unsigned long long*)(sh.park + 2 * W * 32);
    sh.jbar = sh.mbar + 12;                          // (the asynchronous hand-over uses at most 2 * 3 + 2 of the 16 slots)
    sh.jt_base = 0;
    sh.em_mask = sh.mbar + 16;
    sh.em_off = (long long*)(sh.em_mask + SBM_CT_EM * 32);
    sh.em_n = (int*)(sh.em_off + SBM_CT_EM * 32);
    sh.pcb = sh.em_n + 32;
    sh.trow = (long long*)(sh.pcb + SBM_NP * 32 + (SBM_NP & 1) * 32);
    sh.em_soff = sh.trow + 32;
    sh.em_val = (double*)(sh.em_soff + SBM_CT_EM * 32);
    sh.em_tk = sh.em_val + SBM_CT_EM * 32;
    sh.em_kt = (int*)(sh.em_tk + 32);
    sh.tmem_base = 0;
#ifdef SBM_CT_PROF
    if (threadIdx.x < 8) sbm_ctl_prof[threadIdx.x] = 0;
#endif

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // ---- one-time initialisation of the shared state ----
    for (int q = threadIdx.x; q < NBUF * SBM_NF * (N + (N & 1)) * 32; q += WT * 32) sh.Gd[q] = 0.0;   // structural zeros of G
    for (int q = threadIdx.x; q < NBUF * SBM_NJS * 32; q += WT * 32) sh.Jb[q] = 0.0;
    for (int q = threadIdx.x; q < 7 * N * 32; q += WT * 32) sh.fh[q] = 0.0;
    for (int q = threadIdx.x; q < 2 * N * 32; q += WT * 32) sh.Pb[q] = 0.0;
    for (int q = threadIdx.x; q < (W + 3) * 32; q += WT * 32) sh.sqb[q] = 0.0;
    for (int q = threadIdx.x; q < N * 32; q += WT * 32) sh.ynb[q] = 0.0;
    if (warp == 0) {
        for (int i = 0; i < N; ++i) sh.yb[i * 32 + lane] = 0.0;
        for (int m = 0; m < SBM_NP; ++m) sh.pb[m * 32 + lane] = 1.0;
        for (int q = 0; q < SBM_CT_ND; ++q) sh.cd[q * 32 + lane] = 0.0;
        sh.cd[SBM_CT_D_TEND * 32 + lane] = 1.0;
        for (int q = 0; q < SBM_CT_NI; ++q) sh.ci[q * 32 + lane] = 0;
        for (int q = 0; q < 4; ++q) sh.cnt[q * 32 + lane] = 0ull;
        sh.trj[lane] = -1;
        sh.em_kt[lane] = -1;
        sh.hb[lane] = 0.0;
        sh.fl[lane] = 0;
        sh.ktb[lane] = 0;
        if (lane != 1) sh.cta[lane] = 0;             // cta[1] receives the tensor-memory base address from tcgen05.alloc
    }
    for (int q = threadIdx.x; q < 2 * W * 32; q += WT * 32) sh.park[q] = 0.0;
#if SBM_CT_ASYNC
    if (threadIdx.x == 0) {
        for (int b = 0; b < SBM_CT_NB; ++b) { sbm_mb_init(sh.mbar + b, 1); sbm_mb_init(sh.mbar + SBM_CT_NB + b, W); }
        sbm_mb_init(sh.mbar + 2 * SBM_CT_NB, W);         // done: one arrival per column warp
        sbm_mb_init(sh.mbar + 2 * SBM_CT_NB + 1, 1);     // go: the controller
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
#if SBM_CT_REGV
    SBM_CT_BAR(WT * 32);
#else
    // tensor memory for S1 / S4: ceil(W/4) warps per lane quadrant, 4*C*NE columns each (power of two >= 32)
    constexpr int TM_S = ((W + 3) / 4) * 4 * C * (N + (N & 1));
#if SBM_CT_JTM
    constexpr int TM_NEED = TM_S + 2 * 2 * SBM_NJP;   // + two buffers of J (2 columns per double)
    sh.jt_base = TM_S;
    if (threadIdx.x == 0) {
        sbm_mb_init(sh.jbar, 1);
        sbm_mb_init(sh.jbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#else
    constexpr int TM_NEED = TM_S;
#endif
    constexpr int TM_COLS = TM_NEED <= 32 ? 32 : TM_NEED <= 64 ? 64 : TM_NEED <= 128 ? 128 : TM_NEED <= 256 ? 256 : 512;
    static_assert(TM_NEED <= 512, "sensitivity columns do not fit in tensor memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((unsigned)__cvta_generic_to_shared(sh.cta + 1)), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    SBM_CT_BAR(WT * 32);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    sh.tmem_base = (unsigned)sh.cta[1];
#endif

#if SBM_CT_ASYNC
    static_assert(XW == 1, "the asynchronous hand-over needs the dedicated model warp");
    if (warp >= W) sbm_ct_model_warp_async<W, C>(a, sh, lane);
    else if (warp == W - 1 && CL != C) sbm_ct_warp_async<W, C, CL>(a, sh, warp, lane);
    else sbm_ct_warp_async<W, C, C>(a, sh, warp, lane);
#elif defined(SBM_CT_REGSPLIT)
    // experiment: 16 warps = 4 warp groups launched with 128 registers per thread; the last warp group (model shares only)
    // gives registers back, the other three (column warps and the main model warp) take them: 3 x 128 x 152 + 128 x 56 = 64 K
    static_assert(WT == 16 && W <= 12 && W >= 9, "register split: three warp groups of column warps + one of model warps");
    if (warp >= 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
        sbm_ct_model_warp<W, C, XW>(a, sh, warp, lane);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;" ::: "memory");
        if (warp >= W) sbm_ct_model_warp<W, C, XW>(a, sh, warp, lane);
        else if (warp == W - 1 && CL != C) sbm_ct_warp<W, C, CL, XW>(a, sh, warp, lane);
        else sbm_ct_warp<W, C, C, XW>(a, sh, warp, lane);
    }
#elif SBM_CT_LAG
    static_assert(XW == 1, "the lagged controller needs the dedicated model warp");
    // The tables of a small plan (output times, rows of every output time: C4 has 20 times x 2 rows) are staged in shared
    // memory: the set-up of a step that lands on an output time walks orow_off -> row_mask / row_sim / row_jac -> tout, three
    // dependent loads that cost the model warp - the critical path of the CTA - an L2 round trip each.
    __shared__ double t_tout[SBM_CTL_TAB_T];
    __shared__ unsigned long long t_mask[SBM_CTL_TAB_R];
    __shared__ int t_orow_off[SBM_CTL_TAB_T + 2], t_rsim[SBM_CTL_TAB_R], t_rjac[SBM_CTL_TAB_R];
    SbmSolveArgs al = a;
    if (a.n_tout <= SBM_CTL_TAB_T && a.n_orow <= SBM_CTL_TAB_R) {
        for (int q = threadIdx.x; q < a.n_tout; q += WT * 32) t_tout[q] = a.tout[q];
        for (int q = threadIdx.x; q <= a.n_tout; q += WT * 32) t_orow_off[q] = a.orow_off[q];
        for (int q = threadIdx.x; q < a.n_orow; q += WT * 32) { t_mask[q] = a.row_mask[q]; t_rsim[q] = a.row_sim[q]; t_rjac[q] = a.row_jac[q]; }
        al.tout = t_tout; al.orow_off = t_orow_off; al.row_mask = t_mask; al.row_sim = t_rsim; al.row_jac = t_rjac;
        SBM_CT_BAR(WT * 32);
    }
    if (warp >= W) sbm_ctl_model_warp<W, C, XW>(al, sh, warp, lane);
    else if (warp == W - 1 && CL != C) sbm_ctl_warp<W, C, CL, XW>(al, sh, warp, lane);
    else sbm_ctl_warp<W, C, C, XW>(al, sh, warp, lane);
#else
    if (XW > 0 && warp >= W) sbm_ct_model_warp<W, C, XW>(a, sh, warp, lane);
    else if (warp == W - 1 && CL != C) sbm_ct_warp<W, C, CL, XW>(a, sh, warp, lane);
    else sbm_ct_warp<W, C, C, XW>(a, sh, warp, lane);
#endif

    // ---- counters: one atomic per CTA ----
#if SBM_CT_REGV
    SBM_CT_BAR(WT * 32);
#else
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    SBM_CT_BAR(WT * 32);
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(sh.tmem_base), "r"(TM_COLS) : "memory");
    }
#endif
    if (warp == 0) {
        unsigned long long v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] = sh.cnt[q * 32 + lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        }
        if (lane == 0) {
            atomicAdd(a.counters + SBM_CNT_STEPS, v[0]);
            atomicAdd(a.counters + SBM_CNT_REJECT, v[1]);
            atomicAdd(a.counters + SBM_CNT_RHS, v[2]);
            atomicAdd(a.counters + SBM_CNT_TRAJ, v[3]);
        }
    }
}

#define SBM_DEFINE_DP5_CT(NAME, W, C, XW)                                                               \
    extern "C" __global__ void __launch_bounds__(((W) + (XW)) * 32, 1) NAME(const SbmSolveArgs a)       \
    { sbm_dp5_ct_impl<W, C, XW>(a); }

#endif  /* SBM_CT_W > 0 */
