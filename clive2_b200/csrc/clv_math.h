/* clv_math.h — float32 elementary functions built only from IEEE-754 correctly rounded operations.
 *
 * Why this exists: the CUDA kernels and the CPU oracle have to produce the same random walks to be
 * diffed vertex by vertex, but glibc's sinf/cosf/acosf/atanf/expf and CUDA's libdevice versions differ in the last
 * bits, after which the two walks diverge chaotically.  Every function here is a fixed sequence of float
 * add/sub/mul/div/sqrt/floor and integer bit operations, so it rounds identically on an x86 host (gcc
 * -ffp-contract=off) and on an sm_100a device (explicit __fmul_rn/__fadd_rn, which ptxas never fuses into FFMA).
 *
 * The product (the .cu files next to this header) always uses these.  The oracle is built twice (oracle/Makefile): once against
 * glibc libm (the canonical "C transcription", images compared within tolerance) and once with -DCLV_SHARED_MATH
 * against this header (every buffer compared bit for bit).  tests/test_clv_math.py bounds the error of these
 * functions against libm so a mistake here cannot hide in both sides at once.
 *
 * Polynomial coefficients are the classic single-precision Cephes minimax sets (public domain, S. Moshier);
 * only the domains the path tracer uses are range-reduced: sin/cos for |x| up to a few thousand, acos on [-1,1],
 * atan on the whole line, exp on [-87,88].
 */
#ifndef CLV_MATH_H
#define CLV_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CLV_MATH_FN __host__ __device__ __forceinline__
#define CLV_MATH_API __host__ __device__ __forceinline__
#else
#define CLV_MATH_FN static inline
#define CLV_MATH_API static inline
#endif

#if defined(__CUDA_ARCH__)
#define CLV_MUL(a, b) __fmul_rn((a), (b))
#define CLV_ADD(a, b) __fadd_rn((a), (b))
#define CLV_SUB(a, b) __fsub_rn((a), (b))
#define CLV_DIV(a, b) __fdiv_rn((a), (b))
#define CLV_SQRT(a) __fsqrt_rn((a))
#else
#define CLV_MUL(a, b) ((a) * (b))
#define CLV_ADD(a, b) ((a) + (b))
#define CLV_SUB(a, b) ((a) - (b))
#define CLV_DIV(a, b) ((a) / (b))
#define CLV_SQRT(a) sqrtf((a))
#endif

CLV_MATH_FN float clv_bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

/* Horner step a*x + b with two roundings */
#define CLV_HORNER(a, x, b) CLV_ADD(CLV_MUL((a), (x)), (b))

/* ---- sin / cos ------------------------------------------------------------------------------------------- */
/* Reduce x = k*(pi/2) + r with |r| <= pi/4 (three-term Cody-Waite), return r and k mod 4. */
CLV_MATH_FN float clv_reduce_pio2(float x, int *quadrant) {
    const float two_over_pi = 0.636619772367581343f;
    const float p1 = 1.5703125f;                 /* pi/2 split: 1.5703125 + 4.837512969970703125e-4 + 7.549789948768648e-8 */
    const float p2 = 4.837512969970703125e-4f;
    const float p3 = 7.549789948768648e-8f;
    float kf = floorf(CLV_ADD(CLV_MUL(x, two_over_pi), 0.5f));
    float r = CLV_SUB(x, CLV_MUL(kf, p1));
    r = CLV_SUB(r, CLV_MUL(kf, p2));
    r = CLV_SUB(r, CLV_MUL(kf, p3));
    *quadrant = ((int)kf) & 3;
    return r;
}

CLV_MATH_FN float clv_sin_poly(float r) {
    float z = CLV_MUL(r, r);
    float p = CLV_HORNER(-1.9515295891e-4f, z, 8.3321608736e-3f);
    p = CLV_HORNER(p, z, -1.6666654611e-1f);
    return CLV_ADD(CLV_MUL(CLV_MUL(p, z), r), r);
}

CLV_MATH_FN float clv_cos_poly(float r) {
    float z = CLV_MUL(r, r);
    float p = CLV_HORNER(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    p = CLV_HORNER(p, z, 4.166664568298827e-2f);
    /* 1 - z/2 + z*z*p */
    return CLV_ADD(CLV_SUB(1.0f, CLV_MUL(0.5f, z)), CLV_MUL(CLV_MUL(z, z), p));
}

CLV_MATH_API float clv_sinf(float x) {
    int q;
    float r = clv_reduce_pio2(x, &q);
    float v = (q & 1) ? clv_cos_poly(r) : clv_sin_poly(r);
    return (q & 2) ? -v : v;
}

CLV_MATH_API float clv_cosf(float x) {
    int q;
    float r = clv_reduce_pio2(x, &q);
    float v = (q & 1) ? clv_sin_poly(r) : clv_cos_poly(r);
    return ((q + 1) & 2) ? -v : v;
}

/* ---- asin / acos ------------------------------------------------------------------------------------------ */
CLV_MATH_FN float clv_asin_poly(float x) { /* |x| <= 0.5 */
    float z = CLV_MUL(x, x);
    float p = CLV_HORNER(4.2163199048e-2f, z, 2.4181311049e-2f);
    p = CLV_HORNER(p, z, 4.5470025998e-2f);
    p = CLV_HORNER(p, z, 7.4953002686e-2f);
    p = CLV_HORNER(p, z, 1.6666752422e-1f);
    return CLV_ADD(CLV_MUL(CLV_MUL(p, z), x), x);
}

CLV_MATH_API float clv_acosf(float x) {
    const float pi = 3.14159265358979323846f;
    const float pio2 = 1.57079632679489661923f;
    float a = fabsf(x);
    if (!(a <= 1.0f)) return clv_bits_to_float(0x7fc00000u); /* NaN for |x| > 1 and for NaN */
    if (a > 0.5f) {
        float s = CLV_SQRT(CLV_MUL(0.5f, CLV_SUB(1.0f, a)));
        float r = CLV_MUL(2.0f, clv_asin_poly(s));
        return (x < 0.0f) ? CLV_SUB(pi, r) : r;
    }
    return CLV_SUB(pio2, clv_asin_poly(x));
}

/* ---- atan ------------------------------------------------------------------------------------------------ */
CLV_MATH_API float clv_atanf(float x) {
    const float pio2 = 1.57079632679489661923f;
    const float pio4 = 0.78539816339744830962f;
    float a = fabsf(x);
    float y0, t;
    if (a > 2.414213562373095f) { /* tan(3pi/8) */
        y0 = pio2;
        t = -CLV_DIV(1.0f, a);
    } else if (a > 0.4142135623730950f) { /* tan(pi/8) */
        y0 = pio4;
        t = CLV_DIV(CLV_SUB(a, 1.0f), CLV_ADD(a, 1.0f));
    } else {
        y0 = 0.0f;
        t = a;
    }
    float z = CLV_MUL(t, t);
    float p = CLV_HORNER(8.05374449538e-2f, z, -1.38776856032e-1f);
    p = CLV_HORNER(p, z, 1.99777106478e-1f);
    p = CLV_HORNER(p, z, -3.33329491539e-1f);
    float r = CLV_ADD(y0, CLV_ADD(CLV_MUL(CLV_MUL(p, z), t), t));
    return (x < 0.0f) ? -r : r;
}

/* ---- exp ------------------------------------------------------------------------------------------------- */
CLV_MATH_API float clv_expf(float x) {
    if (!(x == x)) return x;
    if (x > 88.7f) return clv_bits_to_float(0x7f800000u);
    if (x < -87.3f) return 0.0f;
    float nf = floorf(CLV_ADD(CLV_MUL(x, 1.44269504088896341f), 0.5f));
    float r = CLV_SUB(x, CLV_MUL(nf, 0.693359375f));
    r = CLV_SUB(r, CLV_MUL(nf, -2.12194440e-4f));
    float p = CLV_HORNER(1.9875691500e-4f, r, 1.3981999507e-3f);
    p = CLV_HORNER(p, r, 8.3334519073e-3f);
    p = CLV_HORNER(p, r, 4.1665795894e-2f);
    p = CLV_HORNER(p, r, 1.6666665459e-1f);
    p = CLV_HORNER(p, r, 5.0000001201e-1f);
    float e = CLV_ADD(CLV_ADD(CLV_MUL(p, CLV_MUL(r, r)), r), 1.0f);
    int n = (int)nf;                       /* in [-126, 128] after the clamps above */
    /* scale by 2^n in two exact steps so n = 128 and n = -126 stay representable */
    int n1 = n / 2, n2 = n - n1;
    float s1 = clv_bits_to_float((uint32_t)(n1 + 127) << 23);
    float s2 = clv_bits_to_float((uint32_t)(n2 + 127) << 23);
    return CLV_MUL(CLV_MUL(e, s1), s2);
}

#endif /* CLV_MATH_H */
