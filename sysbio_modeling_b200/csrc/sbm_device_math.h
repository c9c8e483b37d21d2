// sbm_device_math.h - branch-free FP64 helpers for the generated model code and the integrators.
//
// nvcc/NVRTC expand every `1.0/x` and `a/b` into a MUFU seed + Newton steps + a range check that
// branches to a slow-path subroutine.  The generated right-hand sides contain a reciprocal per rate law
// (13 in the MAPK model) and the error norm one per state, so those branches cut the code into ~20 basic
// blocks per stage and the scheduler cannot overlap the independent rate-law chains.  sbm_rcp() has no
// branch: hardware seed (rcp.approx.ftz.f64, relative error <= 2^-20) + three Newton-Raphson steps in
// FMA form -> within 1 ulp for normal arguments.  Zero, subnormal, Inf and NaN arguments give Inf/NaN/0
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
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

#endif
