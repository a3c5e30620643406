// libm_probe.cpp - pins starflood_b200/csrc/sfe_libm.cuh against the host C library: every float of the ranges that
// vis_draw() (visualization.c:503-683) can feed to logf / sinf / cosf is evaluated by the restatement and by libm and
// the bit patterns are compared.  Also counts how often a correctly rounded result (double routine rounded once) differs
// from libm - the reason the restatement exists.   build: g++ -O2 -ffp-contract=off -fopenmp tools/libm_probe.cpp -o tools/libm_probe
// usage: libm_probe [stride]   (stride 1 = exhaustive, ~1 min on 8 threads; the CPU test suite uses a prime stride)
#include "../starflood_b200/csrc/sfe_libm.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 10) : 1u;
    unsigned long long bad_log = 0, bad_cos = 0, bad_sin = 0, n_log = 0, n_trig = 0, cr_log = 0, cr_cos = 0, cr_sin = 0;
    // logf: every positive finite float (subnormals included), plus the special values
    #pragma omp parallel for reduction(+ : bad_log, n_log, cr_log) schedule(static)
    for (long long b = 1; b < 0x7f800000ll; b += stride) {
        const float x = sfe_libm::as_f32((uint32_t)b);
        const float ref = logf(x), got = sfe_libm::logf_glibc(x);
        n_log++;
        if (sfe_libm::as_u32(ref) != sfe_libm::as_u32(got)) bad_log++;
        // the variant the splat kernel calls for its uniform random numbers: normal x below 1
        if (b >= 0x00800000ll && b < 0x3f800000ll && sfe_libm::as_u32(ref) != sfe_libm::as_u32(sfe_libm::logf_glibc_below_one(x))) bad_log++;
        if (sfe_libm::as_u32(ref) != sfe_libm::as_u32((float)log((double)x))) cr_log++;
    }
    const float specials[] = {0.0f, -0.0f, INFINITY, -1.0f, NAN};
    for (float x : specials) {
        const float ref = logf(x), got = sfe_libm::logf_glibc(x);
        if (!((ref != ref && got != got) || sfe_libm::as_u32(ref) == sfe_libm::as_u32(got))) bad_log++;
    }
    // sinf / cosf: every float with |y| < 120, both signs
    #pragma omp parallel for reduction(+ : bad_cos, bad_sin, n_trig, cr_cos, cr_sin) schedule(static)
    for (long long b = 0; b < 0x42f00000ll; b += stride) {
        for (uint32_t sign = 0; sign < 2; sign++) {
            const float y = sfe_libm::as_f32((uint32_t)b | (sign << 31));
            const float rc = cosf(y), rs = sinf(y);
            float gs, gc;
            sfe_libm::sincosf_glibc(y, &gs, &gc);
            n_trig++;
            if (sfe_libm::as_u32(rc) != sfe_libm::as_u32(gc)) bad_cos++;
            if (sfe_libm::as_u32(rs) != sfe_libm::as_u32(gs)) bad_sin++;
            if (sfe_libm::as_u32(rc) != sfe_libm::as_u32((float)cos((double)y))) cr_cos++;
            if (sfe_libm::as_u32(rs) != sfe_libm::as_u32((float)sin((double)y))) cr_sin++;
        }
    }
    printf("{\"stride\": %u, \"logf\": {\"checked\": %llu, \"mismatches\": %llu, \"correctly_rounded_differs\": %llu}, "
           "\"cosf\": {\"checked\": %llu, \"mismatches\": %llu, \"correctly_rounded_differs\": %llu}, "
           "\"sinf\": {\"checked\": %llu, \"mismatches\": %llu, \"correctly_rounded_differs\": %llu}}\n",
           stride, n_log, bad_log, cr_log, n_trig, bad_cos, cr_cos, n_trig, bad_sin, cr_sin);
    return (bad_log || bad_cos || bad_sin) ? 1 : 0;
}
