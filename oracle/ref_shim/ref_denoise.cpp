// ref_denoise.cpp -- TEST INFRASTRUCTURE (oracle/_ref): compiles the UNMODIFIED reference denoiser kernels
// render/optixutils/c_src/denoising.cu (bilateral_denoiser_fwd_kernel / _bwd_kernel, with denoising.h, common.h, accessor.h,
// math_utils.h) for the host from where the file lies under /root/reference, and runs them one "thread" per pixel.
// Nothing of the reference is copied; see ref_env_shade.cpp for the conventions of this recipe.
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
#define __CUDACC__ 1
#endif

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

// the launch geometry a CUDA thread sees: one pixel per "thread", blocks of 1x1x1
static thread_local uint3 blockIdx, threadIdx;
static thread_local dim3 blockDim;

#include REF_DENOISE

namespace {
struct Raw4 { void *p; int32_t sizes[4]; int32_t strides[4]; };
template <class A> void fill4(A &dst, const void *p, int B, int H, int W, int Cn)
{
    static_assert(sizeof(A) == sizeof(Raw4), "accessor layout changed");
    Raw4 r; r.p = const_cast<void *>(p);
    r.sizes[0] = B; r.sizes[1] = H; r.sizes[2] = W; r.sizes[3] = Cn;
    r.strides[3] = 1; r.strides[2] = Cn; r.strides[1] = Cn * W; r.strides[0] = Cn * W * H;
    std::memcpy((void *)&dst, &r, sizeof(r));
}
template <class K> void launch(K kernel, const BilateralDenoiserParams &p, int B, int H, int W)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < B; ++z)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                blockDim = dim3(1, 1, 1); threadIdx = make_uint3(0, 0, 0); blockIdx = make_uint3((unsigned)x, (unsigned)y, (unsigned)z);
                kernel(p);
            }
}
}  // namespace

extern "C" {

// col [B,H,W,3], nrm [B,H,W,3], zdz [B,H,W,2] -> out [B,H,W,4] (rgb weighted sum, weight): torch_bindings.cpp:274-295
void ref_bilateral_fwd(int B, int H, int W, float sigma, const float *col, const float *nrm, const float *zdz, float *out)
{
    BilateralDenoiserParams p;
    std::memset((void *)&p, 0, sizeof(p));
    fill4(p.col, col, B, H, W, 3); fill4(p.nrm, nrm, B, H, W, 3); fill4(p.zdz, zdz, B, H, W, 2); fill4(p.out, out, B, H, W, 4);
    p.sigma = sigma;
    launch(bilateral_denoiser_fwd_kernel, p, B, H, W);
}

// out_grad [B,H,W,4] -> col_grad [B,H,W,3]: torch_bindings.cpp:297-319
void ref_bilateral_bwd(int B, int H, int W, float sigma, const float *col, const float *nrm, const float *zdz, const float *out_grad, float *col_grad)
{
    BilateralDenoiserParams p;
    std::memset((void *)&p, 0, sizeof(p));
    fill4(p.col, col, B, H, W, 3); fill4(p.nrm, nrm, B, H, W, 3); fill4(p.zdz, zdz, B, H, W, 2); fill4(p.out_grad, out_grad, B, H, W, 4);
    fill4(p.col_grad, col_grad, B, H, W, 3);
    p.sigma = sigma;
    launch(bilateral_denoiser_bwd_kernel, p, B, H, W);
}

}  // extern "C"
