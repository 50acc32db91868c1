/* optix.h -- host stand-in for the handful of OptiX device API names that the reference's raygen program uses
 * (render/optixutils/c_src/envsampling/kernel.cu:101-118, 463-467, 544-547).  NOT the OptiX SDK header: this file belongs to the
 * test-only recipe oracle/ref_shim/ that compiles the UNMODIFIED reference kernel source for the CPU (see ref_env_shade.cpp).
 * A trace call asks the oracle's brute-force visibility predicate; a miss runs the reference's own __miss__ms program. */
#pragma once
#include <stdint.h>

typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
enum { OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1, OPTIX_RAY_FLAG_DISABLE_CLOSESTHIT = 2, OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 4 };

struct RefShimState {                      /* per-thread launch state */
    uint3 idx, dim;
    unsigned int payload0;
};
extern thread_local RefShimState g_shim;
bool ref_shim_occluded(float3 o, float3 d, float tmin, float tmax);      /* defined in ref_env_shade.cpp */
void ref_shim_log_ray(float3 o, float3 d);                              /* optional per-pixel ray log (test hook) */
extern "C" void __miss__ms();

static inline uint3 optixGetLaunchIndex() { return g_shim.idx; }
static inline uint3 optixGetLaunchDimensions() { return g_shim.dim; }
static inline void optixSetPayload_0(unsigned int v) { g_shim.payload0 = v; }
static inline void optixTrace(OptixTraversableHandle, float3 origin, float3 dir, float tmin, float tmax, float, OptixVisibilityMask, unsigned int,
                              unsigned int, unsigned int, unsigned int, unsigned int &p0)
{
    g_shim.payload0 = p0;
    ref_shim_log_ray(origin, dir);
    if (!ref_shim_occluded(origin, dir, tmin, tmax)) __miss__ms();      /* any-hit and closest-hit programs are disabled by the ray flags */
    p0 = g_shim.payload0;
}
