// sfe_libm.cuh - logf / sinf / cosf with the bit patterns of the host C library the reference links against.
//
// vis_draw() (visualization.c:503-683) turns libm results into INTEGER pixel coordinates and colours, so "close" is not
// enough for a bit-identical accumulation buffer: over the arguments vis_draw() draws (uniform random numbers times tau)
// a correctly rounded cosf / sinf differs from glibc's in 1.3 % of the cases and logf in 0.7 % (0.08 % / 0.13 % / 0.02 %
// over all floats, tools/libm_probe.cpp).  The reference has no arithmetic of its own here - it calls the
// platform libm (types.h:52-60) - and this image's is glibc 2.39, whose single-precision routines are the ARM
// optimized-routines algorithms evaluated in double precision:
//   logf    16-entry table of (1/c, log c), degree-3 polynomial of r = z/c - 1          (sysdeps/ieee754/flt-32/e_logf.c)
//   sinf/cosf  quadrant reduction by 2/pi scaled 2^24, degree-7/8 polynomials           (sysdeps/ieee754/flt-32/s_sincosf.h)
// x86-64 glibc dispatches to its FMA build on any CPU with AVX2 + FMA, i.e. every a*b + c below is one fused operation.
// The restatement is checked against the host libm for EVERY float of the ranges vis_draw() can produce
// (tools/libm_probe.cpp, run by tests/test_host_logic.py), and the device evaluates the same IEEE double operations
// (__fma_rn / __dmul_rn / __dadd_rn), so device and host agree bit for bit.
// Arguments outside the fast paths (|x| >= 120 for sin/cos; zero, negative, subnormal, inf, nan for log) take the
// documented fall-backs at the bottom; vis_draw() reaches them only for absurd potentials or a zero random number.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)          /* device code of the engine (sfe_vis.cu) */
#define SFE_LM_HD __device__ __forceinline__
#define SFE_LM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SFE_LM_MUL(a, b) __dmul_rn((a), (b))
#define SFE_LM_ADD(a, b) __dadd_rn((a), (b))
#define SFE_LM_TABLE static __device__ const
#else                            /* host build of the probe: g++ -ffp-contract=off */
#define SFE_LM_HD static inline
#define SFE_LM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define SFE_LM_MUL(a, b) ((a) * (b))
#define SFE_LM_ADD(a, b) ((a) + (b))
#define SFE_LM_TABLE static const
#endif

namespace sfe_libm {

// The double constants of the three routines.  On the device they are read as constant-bank operands of DFMA / DMUL
// (a 64-bit literal costs two UMOV instructions per use instead); the host build uses the literals directly.
#define SFE_LM_LN2     0x1.62e42fefa39efp-1
#define SFE_LM_A0      -0x1.00ea348b88334p-2
#define SFE_LM_A1      0x1.5575b0be00b6ap-2
#define SFE_LM_A2      -0x1.ffffef20a4123p-2
#define SFE_LM_S1      -0x1.555545995a603p-3
#define SFE_LM_S2      0x1.1107605230bc4p-7
#define SFE_LM_S3      -0x1.994eb3774cf24p-13
#define SFE_LM_C1      -0x1.ffffffd0c621cp-2
#define SFE_LM_C2      0x1.55553e1068f19p-5
#define SFE_LM_C3      -0x1.6c087e89a359dp-10
#define SFE_LM_C4      0x1.99343027bf8c3p-16
#define SFE_LM_HPI_INV 0x1.45f306dc9c883p+23
#define SFE_LM_HPI     0x1.921fb54442d18p+0
#if defined(__CUDACC__)
static __constant__ double KTAB[13] = {SFE_LM_LN2, SFE_LM_A0, SFE_LM_A1, SFE_LM_A2, SFE_LM_S1, SFE_LM_S2, SFE_LM_S3,
                                       SFE_LM_C1, SFE_LM_C2, SFE_LM_C3, SFE_LM_C4, SFE_LM_HPI_INV, SFE_LM_HPI};
#define SFE_LM_K(i, v) KTAB[i]
#else
#define SFE_LM_K(i, v) (v)
#endif

SFE_LM_HD uint32_t as_u32(float x) {
#if defined(__CUDACC__)
    return __float_as_uint(x);
#else
    uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
SFE_LM_HD float as_f32(uint32_t u) {
#if defined(__CUDACC__)
    return __uint_as_float(u);
#else
    float x; memcpy(&x, &u, 4); return x;
#endif
}

// ---------------------------------------------------------------------------- logf
// __logf_data (LOGF_TABLE_BITS 4): tab[i] = {invc, logc}, ln2, poly[3]
SFE_LM_TABLE double LOGF_TAB[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010bp+0,  -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8eap+0,  -0x1.1aa2bc79c81p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1p+0,               0x0p+0,
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aap-1,  0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3,
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};

// Valid for positive normal x (the caller handles the rest).
SFE_LM_HD float logf_glibc_normal(float x) {
    const double LN2 = SFE_LM_K(0, SFE_LM_LN2);
    const double A0 = SFE_LM_K(1, SFE_LM_A0), A1 = SFE_LM_K(2, SFE_LM_A1), A2 = SFE_LM_K(3, SFE_LM_A2);
    const uint32_t ix = as_u32(x);
    if (ix == 0x3f800000u) return 0.0f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = LOGF_TAB[2 * i], logc = LOGF_TAB[2 * i + 1];
    const double z = (double)as_f32(iz);
    const double r = SFE_LM_FMA(z, invc, -1.0);                 // z * invc - 1
    const double y0 = SFE_LM_FMA((double)k, LN2, logc);         // logc + k * Ln2
    const double r2 = SFE_LM_MUL(r, r);
    double y = SFE_LM_FMA(A1, r, A2);                           // A[1] * r + A[2]
    y = SFE_LM_FMA(A0, r2, y);                                  // A[0] * r2 + y
    y = SFE_LM_FMA(y, r2, SFE_LM_ADD(y0, r));                   // y * r2 + (y0 + r)
    return (float)y;
}

// logf for normal x below 1 (what a uniform random number in (0, 1) is): the same evaluation without the x == 1 test
SFE_LM_HD float logf_glibc_below_one(float x) {
    const double LN2 = SFE_LM_K(0, SFE_LM_LN2);
    const double A0 = SFE_LM_K(1, SFE_LM_A0), A1 = SFE_LM_K(2, SFE_LM_A1), A2 = SFE_LM_K(3, SFE_LM_A2);
    const uint32_t ix = as_u32(x);
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double z = (double)as_f32(iz);
    const double r = SFE_LM_FMA(z, LOGF_TAB[2 * i], -1.0);
    const double y0 = SFE_LM_FMA((double)k, LN2, LOGF_TAB[2 * i + 1]);
    const double r2 = SFE_LM_MUL(r, r);
    double y = SFE_LM_FMA(A1, r, A2);
    y = SFE_LM_FMA(A0, r2, y);
    y = SFE_LM_FMA(y, r2, SFE_LM_ADD(y0, r));
    return (float)y;
}

// ---------------------------------------------------------------------------- sinf / cosf
// the two polynomials of s_sincosf.h: sinf_poly(x, x2, p, n) evaluates the sine one for even n, the cosine one for odd n
SFE_LM_HD double sin_poly(double x, double x2) {
    const double s1 = SFE_LM_K(4, SFE_LM_S1), s2 = SFE_LM_K(5, SFE_LM_S2), s3 = SFE_LM_K(6, SFE_LM_S3);
    const double x3 = SFE_LM_MUL(x, x2);
    const double t1 = SFE_LM_FMA(x2, s3, s2);                   // s2 + x2 * s3
    const double x7 = SFE_LM_MUL(x3, x2);
    const double s = SFE_LM_FMA(x3, s1, x);                     // x + x3 * s1
    return SFE_LM_FMA(x7, t1, s);                               // s + x7 * s1'
}
SFE_LM_HD double cos_poly(double x2) {
    // __sincosf_table[0]; table[1] (used when n & 2) holds the same coefficients NEGATED, and since IEEE rounding is
    // symmetric in sign every intermediate of that evaluation is the exact negation of this one: the caller negates
    const double c0 = 0x1p0, c1 = SFE_LM_K(7, SFE_LM_C1), c2 = SFE_LM_K(8, SFE_LM_C2),
                 c3 = SFE_LM_K(9, SFE_LM_C3), c4 = SFE_LM_K(10, SFE_LM_C4);
    const double x4 = SFE_LM_MUL(x2, x2);
    const double q2 = SFE_LM_FMA(x2, c4, c3);                   // c3 + x2 * c4
    const double q1 = SFE_LM_FMA(x2, c1, c0);                   // c0 + x2 * c1
    const double x6 = SFE_LM_MUL(x4, x2);
    const double c = SFE_LM_FMA(x4, c2, q1);                    // c1' + x4 * c2
    return SFE_LM_FMA(x6, q2, c);                               // c + x6 * c2'
}

SFE_LM_HD uint32_t abstop12(float x) { return (as_u32(x) >> 20) & 0x7ffu; }

// sinf(y) and cosf(y) together, valid for |y| < 120 (the caller handles the rest).  glibc's sinf and cosf reduce the
// argument the same way and then call sinf_poly with n and n ^ 1: the SAME two polynomial values, swapped by the quadrant.
// Evaluating each once gives both results with the bits of the separate calls and without a data-dependent branch.
SFE_LM_HD void sincosf_glibc_fast(float y, float* sn, float* cs) {
    const double HPI_INV = SFE_LM_K(11, SFE_LM_HPI_INV), HPI = SFE_LM_K(12, SFE_LM_HPI);
    double x = (double)y;
    // glibc branches three ways: |y| < 2^-12 returns {y, 1}, |y| < pi/4 evaluates the polynomials on y itself, the rest
    // reduces first.  The reduction of a |y| < pi/4 finds n = 0 and leaves x untouched (fma(-0, hpi, x) = x), and for
    // |y| < 2^-12 the polynomials round to y and 1 in single precision - except that -0 would come out as +0 - so only
    // that (rare) range keeps its early return and one path covers the rest with the same bits (tools/libm_probe.cpp).
    if (abstop12(y) < 0x398u) { *sn = y; *cs = 1.0f; return; }
    const double r = SFE_LM_MUL(x, HPI_INV);
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = SFE_LM_FMA(-(double)n, HPI, x);                          // x - n * hpi
    // glibc evaluates sinf_poly(x * sign[n & 3], x * x, &table[(n >> 1) & 1], .) with sign = {1, -1, -1, 1}: the sine
    // polynomial is odd in its first argument and the cosine one linear in the table's sign, both exactly (symmetric
    // rounding), so the signs are applied to the results instead
    const double x2 = SFE_LM_MUL(x, x);
    const float a0 = (float)sin_poly(x, x2), b0 = (float)cos_poly(x2);
    const float a = (((n + 1) & 2) != 0) ? -a0 : a0;             // sign[n & 3] < 0 for n & 3 in {1, 2}
    const float b = ((n & 2) != 0) ? -b0 : b0;
    *sn = (n & 1) ? b : a;                                       // sinf: sinf_poly(x * s, x * x, p, n)
    *cs = (n & 1) ? a : b;                                       // cosf: sinf_poly(x * s, x * x, p, n ^ 1)
}

// ---------------------------------------------------------------------------- entry points with the fall-backs
SFE_LM_HD float logf_glibc(float x) {
    const uint32_t ix = as_u32(x);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2u == 0u) return -INFINITY;                     // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                         // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return NAN;
        // subnormal: normalise (e_logf.c does the same)
        const uint32_t jx = as_u32(x * 0x1p23f) - (23u << 23);
        // the body only looks at the bits, so a biased exponent below 1 is fine
        const uint32_t tmp = jx - 0x3f330000u;
        const int i = (int)((tmp >> 19) & 15u);
        const int k = (int32_t)tmp >> 23;
        const uint32_t iz = jx - (tmp & 0xff800000u);
        const double LN2 = SFE_LM_K(0, SFE_LM_LN2);
        const double A0 = SFE_LM_K(1, SFE_LM_A0), A1 = SFE_LM_K(2, SFE_LM_A1), A2 = SFE_LM_K(3, SFE_LM_A2);
        const double z = (double)as_f32(iz);
        const double r = SFE_LM_FMA(z, LOGF_TAB[2 * i], -1.0);
        const double y0 = SFE_LM_FMA((double)k, LN2, LOGF_TAB[2 * i + 1]);
        const double r2 = SFE_LM_MUL(r, r);
        double y = SFE_LM_FMA(A1, r, A2);
        y = SFE_LM_FMA(A0, r2, y);
        y = SFE_LM_FMA(y, r2, SFE_LM_ADD(y0, r));
        return (float)y;
    }
    return logf_glibc_normal(x);
}

// Fall-back for |y| >= 120 (not bit-pinned): glibc reduces large arguments with a 192-bit 4/pi table; here the
// double-precision routine is rounded once, which differs from glibc in ~1 % of such arguments by one ulp.
SFE_LM_HD void sincosf_glibc(float y, float* sn, float* cs) {
    if (abstop12(y) < 0x42fu) { sincosf_glibc_fast(y, sn, cs); return; }   // |y| < 120
    *sn = (float)sin((double)y); *cs = (float)cos((double)y);
}
SFE_LM_HD float cosf_glibc(float y) { float s, c; sincosf_glibc(y, &s, &c); return c; }
SFE_LM_HD float sinf_glibc(float y) { float s, c; sincosf_glibc(y, &s, &c); return s; }

}  // namespace sfe_libm
