// ref_env_shade.cpp -- TEST INFRASTRUCTURE (oracle/_ref): compiles the UNMODIFIED reference raygen program
// render/optixutils/c_src/envsampling/kernel.cu (with the headers it includes: params.h, ../common.h, ../math_utils.h, ../bsdf.h,
// ../accessor.h) for the host, from where it lies under /root/reference, and drives it pixel by pixel.  Nothing of the reference is
// copied: the file is #included by path at build time (oracle/__init__.py:build_ref, only when /root/reference exists).
//
// What this file adds around it: the CUDA qualifiers defined away, host versions of the few CUDA math overloads / atomicAdd the source
// uses, the OptiX stand-in of ./optix.h whose trace call asks the oracle's visibility predicate, and an extern "C" entry that fills
// the reference's own `params` struct from raw arrays and loops `__raygen__rg()` over the launch grid.
// Differences from the GPU build of the same source that remain: host libm instead of CUDA libm (last-ulp differences can pick a
// neighbouring env texel for a few rays per million), and the compiler's own FMA contraction (disabled: -ffp-contract=off).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector_types.h>
#include <vector_functions.h>
#include <math_constants.h>

#undef __device__
#undef __global__
#undef __constant__
#undef __host__
#undef __forceinline__
#define __device__
#define __global__
#define __constant__
#define __host__
#define __forceinline__ inline
#ifndef __CUDACC__
#define __CUDACC__ 1            // selects the device-side branches of ../common.h and ../accessor.h
#endif

// ---- CUDA's mixed-precision min / max overloads and the device-only names the source relies on ----
// CUDA resolves abs(float) to the float overload; make the host do the same (plain ::abs would be int abs(int))
using std::abs;
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline double max(double a, float b) { return fmax(a, (double)b); }
static inline double max(float a, double b) { return fmax((double)a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline double min(double a, float b) { return fmin(a, (double)b); }
static inline double min(float a, double b) { return fmin((double)a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline float atomicAdd(float *p, float v)
{
    float old;
#pragma omp atomic capture
    { old = *p; *p += v; }
    return old;
}
static inline void sincos(double x, float *s, float *c) { double ds, dc; ::sincos(x, &ds, &dc); *s = (float)ds; *c = (float)dc; }
static inline void sincos(float x, float *s, float *c) { ::sincosf(x, s, c); }

#include "optix.h"
thread_local RefShimState g_shim;

// ---- the reference source, verbatim, from its own location (REF_KERNEL is passed by the build recipe) ----
#include REF_KERNEL

// ---- visibility: the oracle's predicate (mcoracle.c: orc_occluded1, fp32 build) through a function pointer set by the caller.
//      The reference traces [0, 1e16) through the closed-source OptiX runtime; the hit/miss decision itself is therefore not part
//      of this source and is supplied by the same fixed-order Moeller-Trumbore predicate the product uses. ----
typedef int (*ref_occ_fn)(const void *scene, int mode, const float *ro, const float *rd);
static ref_occ_fn g_occ = nullptr;
static const void *g_scene = nullptr;
static int g_vis_mode = 0;
bool ref_shim_occluded(float3 o, float3 d, float, float)
{
    const float ro[3] = {o.x, o.y, o.z}, rd[3] = {d.x, d.y, d.z};
    return g_occ(g_scene, g_vis_mode, ro, rd) != 0;
}

// ---- optional ray log: direction of every optixTrace call, per pixel in call order (= sample slot order: light i, BSDF i, ...) ----
static float *g_log = nullptr;
static int *g_log_cnt = nullptr;
static int g_log_cap = 0;
void ref_shim_log_ray(float3, float3 d)
{
    if (!g_log) return;
    const size_t pix = ((size_t)g_shim.idx.z * g_shim.dim.y + g_shim.idx.y) * g_shim.dim.x + g_shim.idx.x;
    const int k = g_log_cnt[pix]++;                                    // one OS thread per pixel: no race
    if (k < g_log_cap) { float *o = g_log + (pix * g_log_cap + k) * 3; o[0] = d.x; o[1] = d.y; o[2] = d.z; }
}

namespace {
template <int N> struct Raw { void *p; int32_t sizes[N]; int32_t strides[N]; };
template <int N, class A> void fill(A &dst, const void *p, const int32_t *sizes)
{
    static_assert(sizeof(A) == sizeof(Raw<N>), "accessor layout changed");
    Raw<N> r; r.p = const_cast<void *>(p);
    int32_t st = 1;
    for (int i = N - 1; i >= 0; --i) { r.sizes[i] = sizes[i]; r.strides[i] = sizes[i] == 1 ? 0 : st; st *= sizes[i]; }
    std::memcpy((void *)&dst, &r, sizeof(r));
}
}  // namespace

extern "C" {

void ref_set_ray_log(float *dirs, int *counts, int cap) { g_log = dirs; g_log_cnt = counts; g_log_cap = cap; }

void ref_set_visibility(void *fn, const void *scene, int mode) { g_occ = (ref_occ_fn)fn; g_scene = scene; g_vis_mode = mode; }

// All arrays contiguous fp32 (perms int32).  view_pos is [B,1,1,3] (render/render.py passes the broadcast camera position).
// backward != 0: diff_grad / spec_grad are inputs and the *_grad outputs (zero-initialised by the caller) are accumulated.
void ref_env_shade(int B, int H, int W, int Hl, int Wl, int n_perms, int n_samples_x, unsigned int bsdf, unsigned int rnd_seed, float shadow_scale,
                   int backward, const float *mask, const float *ro, const float *gb_pos, const float *gb_normal, const float *gb_view_pos,
                   const float *gb_kd, const float *gb_ks, const float *light, const float *pdf, const float *rows, const float *cols, const int *perms,
                   float *diff, float *spec, float *diff_grad, float *spec_grad, float *gb_pos_grad, float *gb_normal_grad, float *gb_kd_grad,
                   float *gb_ks_grad, float *light_grad)
{
    const int32_t s4[4] = {B, H, W, 3}, s3[3] = {B, H, W}, v4[4] = {B, 1, 1, 3}, l3[3] = {Hl, Wl, 3}, l2[2] = {Hl, Wl}, l1[1] = {Hl};
    const int32_t pm[2] = {n_perms, n_samples_x * n_samples_x};
    fill<4>(params.ro, ro, s4); fill<3>(params.mask, mask, s3);
    fill<4>(params.gb_pos, gb_pos, s4); fill<4>(params.gb_pos_grad, gb_pos_grad, s4);
    fill<4>(params.gb_normal, gb_normal, s4); fill<4>(params.gb_normal_grad, gb_normal_grad, s4);
    fill<4>(params.gb_view_pos, gb_view_pos, v4);
    fill<4>(params.gb_kd, gb_kd, s4); fill<4>(params.gb_kd_grad, gb_kd_grad, s4);
    fill<4>(params.gb_ks, gb_ks, s4); fill<4>(params.gb_ks_grad, gb_ks_grad, s4);
    fill<3>(params.light, light, l3); fill<3>(params.light_grad, light_grad, l3);
    fill<2>(params.pdf, pdf, l2); fill<1>(params.rows, rows, l1); fill<2>(params.cols, cols, l2);
    fill<4>(params.diff, diff, s4); fill<4>(params.diff_grad, diff_grad, s4); fill<4>(params.spec, spec, s4); fill<4>(params.spec_grad, spec_grad, s4);
    fill<2>(params.perms, perms, pm);
    params.handle = 0; params.BSDF = bsdf; params.n_samples_x = (unsigned int)n_samples_x; params.rnd_seed = rnd_seed;
    params.backward = (unsigned int)backward; params.shadow_scale = shadow_scale;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int z = 0; z < B; ++z)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                g_shim.idx = make_uint3((unsigned)x, (unsigned)y, (unsigned)z);
                g_shim.dim = make_uint3((unsigned)W, (unsigned)H, (unsigned)B);
                __raygen__rg();
            }
}

}  // extern "C"
