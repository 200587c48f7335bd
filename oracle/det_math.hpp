// oracle/det_math.hpp -- TEST INFRASTRUCTURE ONLY (part of the CPU oracle).
#pragma once
#include <cstdint>
#include <cstring>
#include <limits>
#define DETMATH_FN static inline
#define DETMATH_INF std::numeric_limits<double>::infinity()
#define DETMATH_NAN std::numeric_limits<double>::quiet_NaN()
static inline uint64_t detmath_bits(double v) { uint64_t b; std::memcpy(&b, &v, 8); return b; }
static inline double detmath_from_bits(uint64_t b) { double v; std::memcpy(&v, &b, 8); return v; }
// Deterministic natural logarithm and exponential: the classic argument-reduction + minimax
// polynomial scheme (as in Sun's freely distributable fdlibm e_log.c / e_exp.c), written with
// explicit IEEE double operations only, so that the CPU oracle (compiled with -ffp-contract=off)
// and the CUDA kernels (compiled with --fmad=false) produce bit-identical results.  Both are
// accurate to < 1 ulp; neither is guaranteed to agree bit-for-bit with a platform libm
// (the Rust reference calls the platform libm through f64::ln / f64::exp).
DETMATH_FN double det_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 two54 = 1.80143985094819840000e+16,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    uint64_t bits = detmath_bits(x);
    int32_t hx = (int32_t)(bits >> 32);
    uint32_t lx = (uint32_t)bits;
    int32_t k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | (int32_t)lx) == 0) return -DETMATH_INF;  // log(+-0) = -inf
        if (hx < 0) return DETMATH_NAN;                                   // log(negative)
        k -= 54;
        x *= two54;  // subnormal: scale up
        bits = detmath_bits(x);
        hx = (int32_t)(bits >> 32);
        lx = (uint32_t)bits;
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int32_t i0 = (hx + 0x95f64) & 0x100000;
    x = detmath_from_bits(((uint64_t)(uint32_t)(hx | (i0 ^ 0x3ff00000)) << 32) | lx);  // normalise x or x/2
    k += (i0 >> 20);
    const double f = x - 1.0;
    const double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {  // |f| < 2**-20
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return dk * ln2_hi + dk * ln2_lo;
        }
        const double R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    const double s = f / (2.0 + f);
    const double z = s * s;
    int32_t i = hx - 0x6147a;
    const double w = z * z;
    const int32_t j = 0x6b851 - hx;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    const double R = t2 + t1;
    if (i > 0) {
        const double hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

// exp(x) for finite x in [-700, 700]
DETMATH_FN double det_exp(double x) {
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
                 invln2 = 1.44269504088896338700e+00,
                 P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    double hi = x, lo = 0.0;
    int32_t k = 0;
    const double ax = x < 0.0 ? -x : x;
    if (ax > 0.34657359027997264) {  // |x| > 0.5 ln2
        k = (int32_t)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
        const double t = (double)k;
        hi = x - t * ln2HI;
        lo = t * ln2LO;
        x = hi - lo;
    } else if (ax < 3.725290298461914e-09) {  // |x| < 2**-28
        return 1.0 + x;
    }
    const double t = x * x;
    const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    return detmath_from_bits(detmath_bits(y) + ((uint64_t)(int64_t)k << 52));  // y * 2^k (no over/underflow in range)
}

// expm1(x) for 0 <= x < 1.5 ln2 (the only range ExpRestricted01 asks for: lambda (1 - x) with lambda = ln(m / (m - 1)) <= ln 2):
// the k = 0 and k = 1 branches of the classic scheme (fdlibm s_expm1.c), explicit IEEE double operations only.  Used by BOTH
// the CPU oracle and the kernels for the last rejection test of ExpRestricted01, so that the two sides cannot differ by the
// one ulp that separates CUDA's expm1 from glibc's; agreement with a platform libm is < 1 ulp, not bit-for-bit.
DETMATH_FN double det_expm1(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Q1 = -3.33333333333331316428e-02, Q2 = 1.58730158725481460165e-03, Q3 = -7.93650757867487942473e-05,
                 Q4 = 4.00821782732936239552e-06, Q5 = -2.01099218183624371326e-07;
    int k = 0;
    double c = 0.0;
    if (x > 0.34657359027997264) {  // x > 0.5 ln2: x = hi - lo + ln2
        const double hi = x - ln2_hi, lo = ln2_lo;
        k = 1;
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (x < 5.551115123125783e-17) {  // x < 2**-54
        return x;
    }
    const double hfx = 0.5 * x, hxs = x * hfx;
    const double r1 = 1.0 + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    const double t = 3.0 - r1 * hfx;
    double e = hxs * ((r1 - t) / (6.0 - x * t));
    if (k == 0) return x - (x * e - hxs);
    e = (x * (e - c) - c);
    e -= hxs;
    if (x < -0.25) return -2.0 * (e - (x + 0.5));
    return 1.0 + 2.0 * (x - e);
}
