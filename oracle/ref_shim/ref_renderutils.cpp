// ref_renderutils.cpp -- TEST INFRASTRUCTURE (oracle/_ref): compiles the UNMODIFIED kernels of the reference's renderutils plugin,
// render/renderutils/c_src/bsdf.cu (16 kernels), normal.cu (2), loss.cu (2) and mesh.cu (2) with bsdf.h, normal.h, loss.h, mesh.h,
// common.h, tensor.h, vec3f.h, vec4f.h, for the host from where they lie under /root/reference.  Nothing of the reference is copied;
// conventions as in ref_env_shade.cpp.  The per-pixel kernels run one "thread" per pixel in 1x1x1 blocks.  loss.cu reduces over a warp
// with __shfl_xor_sync: in a 1x1x1 block getWarpSize() is (1,1,1), the other 31 lanes hold 0, so the kernel writes the per-pixel loss
// (the Python side sums the partial tensor either way, renderutils/ops.py:494).  mesh.cu stages the matrix in __shared__ memory behind a
// __syncthreads(): a block of 16 threads is run twice on one OS thread (first pass fills the staging array, second pass reads it; the
// outputs are plain stores, so the second pass overwrites the first).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
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

static thread_local uint3 blockIdx, threadIdx;
static thread_local dim3 blockDim;
#define __shared__ static thread_local
static inline void __syncthreads() {}
static inline float __shfl_xor_sync(unsigned, float, int) { return 0.0f; }          // lanes outside the 1-thread block contribute nothing
dim3 getLaunchBlockSize(int, int, dim3) { return dim3(1, 1, 1); }                     // declared (not defined) by common.h
dim3 getLaunchGridSize(dim3, dim3 d) { return d; }

#include REF_RU_BSDF
#include REF_RU_NORMAL
#include REF_RU_LOSS
#include REF_RU_MESH

namespace {

struct Desc { float *val; float *d_val; int dims[4]; };

// every parameter struct of bsdf.h / normal.h starts with its Tensor members back to back, followed by `dim3 gridSize`
template <class P> void fill(P &p, int nt, const Desc *d, int gx, int gy, int gz)
{
    std::memset((void *)&p, 0, sizeof(P));
    Tensor *t = reinterpret_cast<Tensor *>(&p);
    for (int i = 0; i < nt; ++i) {
        t[i].val = d[i].val; t[i].d_val = d[i].d_val; t[i].fp16 = false;
        int st = 1;
        for (int k = 3; k >= 0; --k) { t[i].dims[k] = d[i].dims[k]; t[i].strides[k] = st; st *= d[i].dims[k]; }
        t[i]._dims[0] = gz; t[i]._dims[1] = gy; t[i]._dims[2] = gx; t[i]._dims[3] = d[i].dims[3];     // make_cuda_tensor(val, outDims, grad), torch_bindings.cpp:108-121
    }
    p.gridSize = dim3(gx, gy, gz);
}
template <class P> void launch(void (*kernel)(P), const P &p)
{
    const int gx = p.gridSize.x, gy = p.gridSize.y, gz = p.gridSize.z;
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < gz; ++z)
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x) {
                blockDim = dim3(1, 1, 1); threadIdx = make_uint3(0, 0, 0); blockIdx = make_uint3((unsigned)x, (unsigned)y, (unsigned)z);
                kernel(p);
            }
}
template <class P> int run(void (*kernel)(P), int want, int nt, const Desc *d, int gx, int gy, int gz)
{
    if (nt != want) return 2;
    P p; fill(p, nt, d, gx, gy, gz);
    launch(kernel, p);
    return 0;
}

}  // namespace

extern "C" int ref_ru_run(const char *name, int nt, const Desc *d, int gx, int gy, int gz, float f0, int i0, int i1)
{
    const std::string k(name);
#define SIMPLE(NAME, KERNEL, P, N) if (k == NAME) return run<P>(KERNEL, N, nt, d, gx, gy, gz)
    SIMPLE("lambert_fwd", LambertFwdKernel, LambertKernelParams, 3); SIMPLE("lambert_bwd", LambertBwdKernel, LambertKernelParams, 3);
    SIMPLE("frostbite_fwd", FrostbiteDiffuseFwdKernel, FrostbiteDiffuseKernelParams, 5); SIMPLE("frostbite_bwd", FrostbiteDiffuseBwdKernel, FrostbiteDiffuseKernelParams, 5);
    SIMPLE("fresnel_fwd", FresnelShlickFwdKernel, FresnelShlickKernelParams, 4); SIMPLE("fresnel_bwd", FresnelShlickBwdKernel, FresnelShlickKernelParams, 4);
    SIMPLE("ndf_fwd", ndfGGXFwdKernel, NdfGGXParams, 3); SIMPLE("ndf_bwd", ndfGGXBwdKernel, NdfGGXParams, 3);
    SIMPLE("lambda_fwd", lambdaGGXFwdKernel, NdfGGXParams, 3); SIMPLE("lambda_bwd", lambdaGGXBwdKernel, NdfGGXParams, 3);
    SIMPLE("masking_fwd", maskingSmithFwdKernel, MaskingSmithParams, 4); SIMPLE("masking_bwd", maskingSmithBwdKernel, MaskingSmithParams, 4);
#undef SIMPLE
    if (k == "specular_fwd" || k == "specular_bwd") {
        if (nt != 6) return 2;
        PbrSpecular p; fill(p, nt, d, gx, gy, gz); p.min_roughness = f0;
        launch(k == "specular_fwd" ? pbrSpecularFwdKernel : pbrSpecularBwdKernel, p);
        return 0;
    }
    if (k == "bsdf_fwd" || k == "bsdf_bwd") {
        if (nt != 7) return 2;
        PbrBSDF p; fill(p, nt, d, gx, gy, gz); p.min_roughness = f0; p.BSDF = i0;
        launch(k == "bsdf_fwd" ? pbrBSDFFwdKernel : pbrBSDFBwdKernel, p);
        return 0;
    }
    if (k == "loss_fwd" || k == "loss_bwd") {        // tensors: img, target, out (fwd: per-pixel loss [N,H,W,1]; bwd: its upstream gradient)
        if (nt != 3) return 2;
        LossKernelParams p; fill(p, nt, d, gx, gy, gz); p.tonemapper = (TonemapperType)i0; p.loss = (LossType)i1;
        launch(k == "loss_fwd" ? imgLossFwdKernel : imgLossBwdKernel, p);
        return 0;
    }
    if (k == "xfm_fwd" || k == "xfm_bwd") {          // tensors: points [B|1,V,3(,1)], matrix [B,4,4(,1)], out [B,V,4|3(,1)]; gx = V, gz = B
        if (nt != 3) return 2;
        XfmKernelParams p; std::memset((void *)&p, 0, sizeof(p));
        Tensor *t[3] = {&p.points, &p.matrix, &p.out};
        for (int i = 0; i < 3; ++i) {
            t[i]->val = d[i].val; t[i]->d_val = d[i].d_val; t[i]->fp16 = false;
            int st = 1;
            for (int q = 3; q >= 0; --q) { t[i]->dims[q] = d[i].dims[q]; t[i]->strides[q] = st; st *= d[i].dims[q]; }
            t[i]->_dims[0] = gz; t[i]->_dims[1] = gx; t[i]->_dims[2] = d[i].dims[2]; t[i]->_dims[3] = 1;        // the 3-D branch of make_cuda_tensor, torch_bindings.cpp:121
        }
        p.isPoints = i0 != 0; p.gridSize = dim3(gx, 1, gz);
        const int nb = (gx + 15) / 16;
#pragma omp parallel for collapse(2) schedule(static)
        for (int z = 0; z < gz; ++z)
            for (int b = 0; b < nb; ++b)
                for (int pass = 0; pass < 2; ++pass)
                    for (int tx = 0; tx < 16; ++tx) {
                        blockDim = dim3(16, 1, 1); threadIdx = make_uint3((unsigned)tx, 0, 0); blockIdx = make_uint3((unsigned)b, 0, (unsigned)z);
                        if (k == "xfm_fwd") xfmPointsFwdKernel(p); else xfmPointsBwdKernel(p);
                    }
        return 0;
    }
    if (k == "psn_fwd" || k == "psn_bwd") {
        if (nt != 7) return 2;
        PrepareShadingNormalKernelParams p; fill(p, nt, d, gx, gy, gz); p.two_sided_shading = i0 != 0; p.opengl = i1 != 0;
        launch(k == "psn_fwd" ? PrepareShadingNormalFwdKernel : PrepareShadingNormalBwdKernel, p);
        return 0;
    }
    return 1;
}
