// sbm_device_math.h - branch-free FP64 helpers for the generated model code and the integrators.
//
// nvcc/NVRTC expand every `1.0/x` and `a/b` into a MUFU seed + Newton steps + a range check that
// branches to a slow-path subroutine.  The generated right-hand sides contain a reciprocal per rate law
// (13 in the MAPK model) and the error norm one per state, so those branches cut the code into ~20 basic
// blocks per stage and the scheduler cannot overlap the independent rate-law chains.  sbm_rcp() has no
// branch: hardware seed (rcp.approx.ftz.f64, relative error <= 2^-20) + two Newton-Raphson steps in
// FMA form (error 2^-20 -> 2^-40 -> 2^-80: the result is the rounding of the last FMA) -> within 1 ulp for normal arguments
// (tests/cuda/check_device_math.cu: identical to IEEE division on 4 M random arguments; a third step changes nothing).  Zero, subnormal, Inf and NaN arguments give Inf/NaN/0
// without special handling (a trajectory that produces them is rejected by the step controller anyway).
#ifndef SBM_DEVICE_MATH_H
#define SBM_DEVICE_MATH_H

__device__ __forceinline__ double sbm_rcp(const double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// Branch-free log / exp for the generated model code (x**a with a symbolic exponent is printed as
// sbm_exp(a * sbm_log(x)); the logarithm is shared with d(x**a)/da).  CUDA's log() / exp() are ~170 / ~120
// SASS instructions each in the MAPK model (every polynomial coefficient is materialised with two moves, and
// special cases branch to slow paths, which cuts the model code into basic blocks); these are ~45 / ~30:
// coefficients come straight from the constant bank and nothing branches.
//   sbm_log: fdlibm's e_log algorithm (x = 2^k m, m in [sqrt(1/2), sqrt(2)), s = f / (2 + f), degree-14 minimax in s),
//            < 1 ulp for positive normal x; zero, negative, subnormal, Inf and NaN arguments all give NaN (the step
//            controller rejects the trajectory), never a wrong finite number.
//   sbm_exp: k = rint(x / ln 2), Cody-Waite reduction in FMA form, degree-13 Taylor polynomial on |r| <= ln2 / 2 (Estrin's scheme),
//            scaling by 2^k in two factors (overflow -> Inf, underflow -> 0 through subnormals), < 1 ulp.
// Both measured against libm on the device: tests/cuda/check_device_math.cu.
// ---------------------------------------------------------------------------------------------------------
__constant__ double sbm_mc_log[9] = {
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,
    6.93147180369123816490e-01 /* ln2_hi */, 1.90821492927058770002e-10 /* ln2_lo */};

__device__ __forceinline__ double sbm_log(const double x)
{
    const double* __restrict__ C = sbm_mc_log;
    int hx = __double2hiint(x);
    const bool bad = (unsigned)(hx - 0x00100000) >= 0x7fe00000u;        // not a positive normal number
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    const double m = __hiloint2double(hx | (i ^ 0x3ff00000), __double2loint(x));   // m or m / 2
    k += i >> 20;
    const double f = m - 1.0;
    const double s = f * sbm_rcp(2.0 + f);
    const double dk = (double)k;
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, C[5], C[3]), C[1]);
    const double t2 = z * fma(w, fma(w, fma(w, C[6], C[4]), C[2]), C[0]);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double r = dk * C[7] - ((hfsq - fma(s, hfsq + R, dk * C[8])) - f);
    return __hiloint2double(bad ? 0x7ff80000 : __double2hiint(r), __double2loint(r));
}

__constant__ double sbm_mc_exp[16] = {
    1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0,
    1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0,
    1.4426950408889634 /* log2(e) */, 6755399441055744.0 /* 1.5 * 2^52 */,
    -6.93147180559945286227e-01 /* -ln2_hi */, -2.31904681384629955842e-17 /* -ln2_lo */};

__device__ __forceinline__ double sbm_exp(const double x)
{
    const double* __restrict__ C = sbm_mc_exp;
    const double t = fma(x, C[12], C[13]);
    const double dk = t - C[13];                                        // rint(x / ln 2)
    double r = fma(dk, C[14], x);
    r = fma(dk, C[15], r);
    // exp(r) = 1 + (r + r^2 q(r)), q of degree 11 by Estrin's scheme: 6 dependent operations instead of 13 (the single model
    // warp of sbm_dp5_ct is bound by the latency of its longest chain, ~8 clocks per dependent FP64 instruction)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double e0 = fma(fma(C[3], r, C[2]), r2, fma(C[1], r, C[0]));
    const double e1 = fma(fma(C[7], r, C[6]), r2, fma(C[5], r, C[4]));
    const double e2 = fma(fma(C[11], r, C[10]), r2, fma(C[9], r, C[8]));
    const double q = fma(r8, e2, fma(r4, e1, e0));
    const double p = 1.0 + fma(r2, q, r);
    int k = __double2int_rn(dk);                                        // saturates; NaN -> 0 (p is NaN then)
    k = max(-2044, min(2044, k));
    const int k1 = k >> 1, k2 = k - k1;
    return p * __hiloint2double((k1 + 1023) << 20, 0) * __hiloint2double((k2 + 1023) << 20, 0);
}

#endif
