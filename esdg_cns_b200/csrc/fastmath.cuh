// FP64 reciprocal / log / exp / sqrt for the kernels' argument ranges (densities, inverse temperatures,
// internal energies: normal, positive, moderate).  The FP64 pipe is the bottleneck of every kernel here and
// the CUDA library versions (which also handle denormals / inf / nan, and whose pow() is ~200 instructions)
// dominated the instruction count; these are MUFU-seeded and <= 1 ulp (tests: test_log_exp_accuracy).
#pragma once
#include <cuda_runtime.h>

namespace esdg {

__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// log / exp for normal, positive, moderate arguments (densities, inverse temperatures, -v_E, log
// of rho_e).  The classic fdlibm argument reductions and minimax polynomials (e_log.c / e_exp.c,
// error < 1 ulp) with the division replaced by fast_rcp; coefficients live in constant memory so
// that the DFMAs take them as constant-bank operands instead of materialising 64-bit immediates.
// ~2.5x fewer instructions than the library calls, which also handle denormals/inf/nan.
static __constant__ double c_lg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                               2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                               1.479819860511658591e-01};
static __constant__ double c_ex[5] = {1.66666666666666019037e-01, -2.77777777770155933842e-03, 6.61375632143793436117e-05,
                               -1.65339022054652515390e-06, 4.13813679705723846039e-08};

__device__ __forceinline__ double log_pos(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;   // mantissa in (sqrt(2)/2, sqrt(2)]
    hi |= (i ^ 0x3ff00000);
    k += (i >> 20);
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double s = f * fast_rcp(2.0 + f);
    const double dk = (double)k;
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, c_lg[5], c_lg[3]), c_lg[1]);
    const double t2 = z * fma(w, fma(w, fma(w, c_lg[6], c_lg[4]), c_lg[2]), c_lg[0]);
    const double R = t2 + t1;
    const double hfsq = .5 * f * f;
    return dk * 6.93147180369123816490e-01 - ((hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f);
}

__device__ __forceinline__ double exp_mod(double x) {
    const double kd = rint(x * 1.44269504088896338700e+00);
    const double hi = fma(kd, -6.93147180369123816490e-01, x);
    const double lo = kd * 1.90821492927058770002e-10;
    const double r = hi - lo;
    const double t = r * r;
    const double c = r - t * fma(t, fma(t, fma(t, fma(t, c_ex[4], c_ex[3]), c_ex[2]), c_ex[1]), c_ex[0]);
    const double y = 1.0 - ((lo - (r * c) * fast_rcp(2.0 - c)) - hi);
    return __hiloint2double(__double2hiint(y) + ((int)kd << 20), __double2loint(y));
}


// sqrt(x), x >= 0: MUFU rsqrt seed, two Newton steps on y ~ 1/sqrt(x), one correction of s = x*y (<= 1 ulp)
__device__ __forceinline__ double fast_sqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);
    y = fma(.5 * y, e, y);
    e = fma(-x * y, y, 1.0);
    y = fma(.5 * y, e, y);
    double s = x * y;
    s = fma(.5 * y, fma(-s, s, x), s);
    return x > 0.0 ? s : 0.0;
}

}  // namespace esdg
