// exact.cuh -- "decision path" arithmetic for the env-light sampler.
//
// Everything that decides WHICH texel is read, WHICH direction is traced and WHICH lobe is sampled
// must round exactly like the CPU oracle (oracle/mcoracle.c, compiled with -ffp-contract=off),
// otherwise a one-ulp difference flips a whole Monte-Carlo sample (DESIGN.md "determinism").
// `xf` is a float whose operators map to the IEEE round-to-nearest intrinsics, which nvcc never
// contracts into FMAs and never replaces by approximate division / sqrt.  `xd` is the same for the
// handful of places where the reference promotes to double through CUDART_PI
// (render/optixutils/c_src/envsampling/kernel.cu:65,74,126-127,134-136).
//
// The four transcendental functions follow the algorithms fixed in DESIGN.md (Cephes single
// precision kernels, fixed evaluation order); this file is an independent implementation of that
// specification -- the oracle has its own (oracle/detmath.h).
#pragma once
#include "common.cuh"

struct xf {
    float v;
    __device__ __forceinline__ xf() {}
    __device__ __forceinline__ xf(float f) : v(f) {}
};
__device__ __forceinline__ xf operator+(xf a, xf b) { return xf(__fadd_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator-(xf a, xf b) { return xf(__fsub_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator*(xf a, xf b) { return xf(__fmul_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator/(xf a, xf b) { return xf(__fdiv_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator-(xf a) { return xf(-a.v); }
__device__ __forceinline__ bool operator<(xf a, xf b) { return a.v < b.v; }
__device__ __forceinline__ bool operator>(xf a, xf b) { return a.v > b.v; }
__device__ __forceinline__ bool operator<=(xf a, xf b) { return a.v <= b.v; }
__device__ __forceinline__ bool operator>=(xf a, xf b) { return a.v >= b.v; }
__device__ __forceinline__ xf xsqrt(xf a) { return xf(__fsqrt_rn(a.v)); }
__device__ __forceinline__ xf xmin(xf a, xf b) { return a.v < b.v ? a : b; }
__device__ __forceinline__ xf xmax(xf a, xf b) { return a.v > b.v ? a : b; }
__device__ __forceinline__ xf xclamp(xf x, xf lo, xf hi) { return xmin(hi, xmax(lo, x)); }

struct xf3 { xf x, y, z; };
__device__ __forceinline__ xf3 X3(xf x, xf y, xf z) { xf3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ xf3 X3(f3 a) { return X3(xf(a.x), xf(a.y), xf(a.z)); }
__device__ __forceinline__ f3 toF3(xf3 a) { return F3(a.x.v, a.y.v, a.z.v); }
__device__ __forceinline__ xf3 operator+(xf3 a, xf3 b) { return X3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ xf3 operator-(xf3 a, xf3 b) { return X3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ xf3 operator*(xf3 a, xf s) { return X3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ xf3 operator/(xf3 a, xf s) { return X3(a.x / s, a.y / s, a.z / s); }
// (a.x*b.x + a.y*b.y) + a.z*b.z, every product and sum rounded
__device__ __forceinline__ xf xdot(xf3 a, xf3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ xf3 xcross(xf3 a, xf3 b)
{
    return X3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ xf3 xnormalize(xf3 v)      // math_utils.h:135-139
{
    xf l = xsqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return l.v > 0.0f ? v / l : X3(xf(0.0f), xf(0.0f), xf(0.0f));
}
// Pixar branchless orthonormal basis, math_utils.h:156-163
__device__ __forceinline__ void xONB(xf3 n, xf3 &b1, xf3 &b2)
{
    xf sign = xf(copysignf(1.0f, n.z.v));
    xf a = xf(-1.0f) / (sign + n.z);
    xf b = n.x * n.y * a;
    b1 = X3(xf(1.0f) + sign * n.x * n.x * a, sign * b, -sign * n.x);
    b2 = X3(b, sign + n.y * n.y * a, -n.y);
}

// ---- double, round-to-nearest, never contracted ------------------------------------------------
#define XD_PI 3.14159265358979323846
__device__ __forceinline__ double xd_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xd_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xd_div(double a, double b) { return __ddiv_rn(a, b); }

// ---- deterministic transcendentals (spec: DESIGN.md; oracle twin: oracle/detmath.h) -------------
#ifndef MCS_DET_INLINE
#define MCS_DET_INLINE __forceinline__
#endif
__device__ MCS_DET_INLINE void det_sincos(xf a, xf &s, xf &c)
{
    xf k = xf(rintf((a * xf(0.636619772367581343f)).v));
    int q = (int)k.v;
    xf r = a - k * xf(1.5703125f);
    r = r - k * xf(4.837512969970703125e-4f);
    r = r - k * xf(7.54978995489188216e-8f);
    xf z = r * r;
    xf sp = ((xf(-1.9515295891e-4f) * z + xf(8.3321608736e-3f)) * z - xf(1.6666654611e-1f)) * z * r + r;
    xf cp = ((xf(2.443315711809948e-5f) * z - xf(1.388731625493765e-3f)) * z + xf(4.166664568298827e-2f)) * z * z - xf(0.5f) * z + xf(1.0f);
    switch (q & 3) {
    case 0:  s = sp;  c = cp;  break;
    case 1:  s = cp;  c = -sp; break;
    case 2:  s = -sp; c = -cp; break;
    default: s = -cp; c = sp;  break;
    }
}
// Range reduction of atan: t > tan(3pi/8): pi/2 + atan(-1/t); t > tan(pi/8): pi/4 + atan((t-1)/(t+1)); else atan(t).  The three ranges
// share ONE division with selected operands -- (-1)/t == -(1/t) and t/1 == t exactly -- so a warp whose lanes fall into different
// ranges (nearly always) executes one IEEE division instead of two divergent ones; the values are bit-identical to the branchy form.
__device__ __forceinline__ xf det_atan_pos(xf t)
{
    const bool hi = t.v > 2.414213562373095f, mid = !hi && t.v > 0.4142135623730950f;
    const xf y0 = xf(hi ? 1.57079632679489661923f : (mid ? 0.78539816339744830962f : 0.0f));
    const xf num = hi ? xf(-1.0f) : (mid ? t - xf(1.0f) : t);
    const xf den = hi ? t : (mid ? t + xf(1.0f) : xf(1.0f));
    t = num / den;
    xf z = t * t;
    xf y = (((xf(8.05374449538e-2f) * z - xf(1.38776856032e-1f)) * z + xf(1.99777106478e-1f)) * z - xf(3.33329491539e-1f)) * z * t + t;
    return y0 + y;
}
__device__ MCS_DET_INLINE xf det_atan2(xf y, xf x)
{
    if (x.v == 0.0f && y.v == 0.0f) return xf(0.0f);
    xf a = det_atan_pos(xf(fabsf(y.v)) / xf(fabsf(x.v)));
    if (x.v < 0.0f) a = xf(3.14159265358979323846f) - a;
    return y.v < 0.0f ? -a : a;
}
__device__ __forceinline__ xf det_asin_kernel(xf a)
{
    xf z = a * a;
    return ((((xf(4.2163199048e-2f) * z + xf(2.4181311049e-2f)) * z + xf(4.5470025998e-2f)) * z + xf(7.4953002686e-2f)) * z + xf(1.6666752422e-1f)) * z * a + a;
}
// acos through ONE evaluation of the asin kernel: |x| > 0.5 uses 2 asin(sqrt((1 - |x|) / 2)) (1 - |x| is 1 + x for negative x, exactly),
// otherwise pi/2 - asin(x).  Same operations on the same values as the three-branch form (oracle/detmath.h), without executing the
// polynomial up to three times per warp.
__device__ MCS_DET_INLINE xf det_acos(xf x)
{
    const bool big = x.v < -0.5f || x.v > 0.5f;
    const xf om = x.v < 0.0f ? xf(1.0f) + x : xf(1.0f) - x;
    const xf a = big ? xsqrt(xf(0.5f) * om) : x;
    const xf k = det_asin_kernel(a);
    if (x.v < -0.5f) return xf(3.14159265358979323846f) - xf(2.0f) * k;
    if (x.v > 0.5f) return xf(2.0f) * k;
    return xf(1.57079632679489661923f) - k;
}
