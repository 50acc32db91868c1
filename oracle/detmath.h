/*
 * oracle/detmath.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Deterministic single-precision sin/cos/atan2/acos used on the *decision path* of the
 * environment-light sampler (texel selection, sampled directions).  The reference calls
 * CUDA libm (sincos / atan2f / acosf under -use_fast_math,
 * render/optixutils/c_src/envsampling/kernel.cu:124-138, optix_wrapper.cpp:36); CUDA libm and
 * glibc disagree in the last ulp, and a one-ulp difference in `_dir_to_tc` flips the selected
 * env-map texel (kernel.cu:177-178,198-199).  To make "same seed => same texels, same rays,
 * same visibility bits" a testable property, DESIGN.md fixes these four functions to the
 * algorithms below (Cephes single-precision kernels: Cody-Waite reduction + minimax polynomial,
 * evaluated with separate IEEE multiplies and adds in exactly this order, no FMA contraction).
 * The CUDA product has its own, independently written, implementation of the same
 * specification (nvdiffrecmc_b200/csrc/detmath.cuh).  Max error vs correctly rounded results
 * is < 2 ulp (tests/test_oracle_math.py).
 *
 * Must be compiled with -ffp-contract=off.
 */
#ifndef MCORACLE_DETMATH_H
#define MCORACLE_DETMATH_H
#include <math.h>

#define DET_PI_F    3.14159265358979323846f
#define DET_PIO2_F  1.57079632679489661923f
#define DET_PIO4_F  0.78539816339744830962f

/* sin and cos of a (|a| < ~1e4), float, deterministic. */
static inline void det_sincosf(float a, float *s, float *c)
{
    float k = rintf(a * 0.636619772367581343f);           /* nearest multiple of pi/2 */
    int   q = (int)k;
    float r = a - k * 1.5703125f;                         /* Cody-Waite, 3 terms */
    r = r - k * 4.837512969970703125e-4f;
    r = r - k * 7.54978995489188216e-8f;
    float z  = r * r;
    float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z
               - 0.5f * z + 1.0f;
    switch (q & 3) {
    case 0:  *s =  sp; *c =  cp; break;
    case 1:  *s =  cp; *c = -sp; break;
    case 2:  *s = -sp; *c = -cp; break;
    default: *s = -cp; *c =  sp; break;
    }
}

static inline float det_sinf(float a) { float s, c; det_sincosf(a, &s, &c); return s; }
static inline float det_cosf(float a) { float s, c; det_sincosf(a, &s, &c); return c; }

/* atan(t) for t >= 0 */
static inline float det_atan_pos(float t)
{
    float y0;
    if (t > 2.414213562373095f)      { y0 = DET_PIO2_F; t = -(1.0f / t); }
    else if (t > 0.4142135623730950f) { y0 = DET_PIO4_F; t = (t - 1.0f) / (t + 1.0f); }
    else                               { y0 = 0.0f; }
    float z = t * t;
    float y = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
               - 3.33329491539e-1f) * z * t + t;
    return y0 + y;
}

static inline float det_atan2f(float y, float x)
{
    if (x == 0.0f && y == 0.0f) return 0.0f;
    float a = det_atan_pos(fabsf(y) / fabsf(x));
    if (x < 0.0f) a = DET_PI_F - a;
    return y < 0.0f ? -a : a;
}

/* asin kernel for |a| <= 0.5 */
static inline float det_asin_kernel(float a)
{
    float z = a * a;
    return ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z
             + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * a + a;
}

/* acos(x), x in [-1, 1] */
static inline float det_acosf(float x)
{
    if (x < -0.5f) return DET_PI_F - 2.0f * det_asin_kernel(sqrtf(0.5f * (1.0f + x)));
    if (x >  0.5f) return 2.0f * det_asin_kernel(sqrtf(0.5f * (1.0f - x)));
    return DET_PIO2_F - det_asin_kernel(x);
}

#endif
