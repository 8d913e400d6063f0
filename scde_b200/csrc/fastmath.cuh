// fastmath.cuh -- exp() for non-positive arguments, used where only a normalising sum or a soft-max weight is needed.
#pragma once

namespace scde {

// exp(a) for a <= 0 (any magnitude): round-to-nearest split a = n ln2 + r, |r| <= 0.3466, Taylor polynomial of degree 10
// (remainder < 3e-13 relative), scaling by an exponent-field add.  Below -700 the scaling is done in two steps so the
// result is a correctly scaled denormal (the last bit of a denormal may differ from libm's); below -745.2 it is exactly 0.
// No special cases for +-Inf / NaN: -Inf gives 0, NaN gives garbage -- callers pass finite or -Inf arguments only.
// About 22 instructions instead of the ~45 of the library routine.
// FULL_RANGE = false: the caller guarantees a >= -700 (no clamp, no two-step scaling).
template <bool FULL_RANGE = true>
__device__ __forceinline__ double exp_nonpos(double a) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: the low word of (t + MAGIC) is rint(t)
    if (FULL_RANGE) a = fmax(a, -746.0);
    const bool tiny = FULL_RANGE && a < -700.0;
    const double t = fma(a, 1.44269504088896340736, MAGIC);
    int n = __double2loint(t);
    const double nf = t - MAGIC;
    double r = fma(nf, -6.93147180369123816490e-01, a);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 2.75573192239858906526e-07;  // 1/10!
    p = fma(p, r, 2.75573192239858906526e-06);
    p = fma(p, r, 2.48015873015873015873e-05);
    p = fma(p, r, 1.98412698412698412698e-04);
    p = fma(p, r, 1.38888888888888888889e-03);
    p = fma(p, r, 8.33333333333333333333e-03);
    p = fma(p, r, 4.16666666666666666667e-02);
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    n += tiny ? 64 : 0;
    const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return tiny ? res * 5.42101086242752217004e-20 /* 2^-64 */ : res;
}

// log(1 + w) for 0 <= w <= 1: z = w / (2 + w) <= 1/3 (reciprocal from the FP32 unit + one Newton step, 4e-15),
// log1p(w) = 2 atanh(z) = 2 z (1 + z^2/3 + ... + z^26/27), truncation below 2e-14.  About 22 instructions.
__device__ __forceinline__ double log1p_unit(double w) {
    const double t = 2.0 + w;
    double r = (double)__frcp_rn((float)t);
    r = r * fma(-t, r, 2.0);
    const double z = w * r, y = z * z;
    double p = 1.0 / 27.0;
    p = fma(p, y, 1.0 / 25.0);
    p = fma(p, y, 1.0 / 23.0);
    p = fma(p, y, 1.0 / 21.0);
    p = fma(p, y, 1.0 / 19.0);
    p = fma(p, y, 1.0 / 17.0);
    p = fma(p, y, 1.0 / 15.0);
    p = fma(p, y, 1.0 / 13.0);
    p = fma(p, y, 1.0 / 11.0);
    p = fma(p, y, 1.0 / 9.0);
    p = fma(p, y, 1.0 / 7.0);
    p = fma(p, y, 1.0 / 5.0);
    p = fma(p, y, 1.0 / 3.0);
    p = fma(p, y, 1.0);
    return (z + z) * p;
}

// log(y) for 0 < y < 2^-900 (normal or denormal): y is scaled by 2^64 (exact), split into m 2^e with m in
// [sqrt(1/2), sqrt 2), and log m = 2 atanh((m - 1) / (m + 1)) is summed to z^18 / 19 (z^2 <= 0.0295: truncation 1e-15).
// Used on exp(a) for a in [-746, -708], where the reference's log(exp(a)) sees the denormal rounding of exp().
__device__ __forceinline__ double log_tiny(double y) {
    const double ys = y * 18446744073709551616.0;  // 2^64
    int hi = __double2hiint(ys);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000FFFFF) | 0x3FF00000;  // mantissa in [1, 2)
    if (hi >= 0x3FF6A09F) {               // above sqrt 2 (to within 1e-6: the series covers both sides)
        hi -= 0x00100000;
        e += 1;
    }
    const double m = __hiloint2double(hi, __double2loint(ys));
    const double t = m + 1.0;
    double r = (double)__frcp_rn((float)t);
    r = r * fma(-t, r, 2.0);
    const double z = (m - 1.0) * r, w = z * z;
    double p = 1.0 / 19.0;
    p = fma(p, w, 1.0 / 17.0);
    p = fma(p, w, 1.0 / 15.0);
    p = fma(p, w, 1.0 / 13.0);
    p = fma(p, w, 1.0 / 11.0);
    p = fma(p, w, 1.0 / 9.0);
    p = fma(p, w, 1.0 / 7.0);
    p = fma(p, w, 1.0 / 5.0);
    p = fma(p, w, 1.0 / 3.0);
    p = fma(p, w, 1.0);
    const double E = (double)(e - 64);
    return fma(E, 6.93147180369123816490e-01, fma(E, 1.90821492927058770002e-10, (z + z) * p));
}

}  // namespace scde
