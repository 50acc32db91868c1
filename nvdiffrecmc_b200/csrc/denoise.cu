// denoise.cu -- cross-bilateral (SVGF-style) denoiser, forward and transposed backward, for sm_100a.
// Replaces bilateral_denoiser_fwd_kernel / _bwd_kernel, render/optixutils/c_src/denoising.cu:14-130
// (8x8 blocks, every tap re-fetched from global memory with 2 expf + powf(.,128) + sqrtf per tap).
//
// B200 design (compute-bound: (2r+1)^2 = 529 taps/px at sigma = 2, only 48 B/px of compulsory HBM
// traffic):
//   * one CTA = 32x16 output pixels, the (32+2r)x(16+2r) halo tile of guides (normal, depth,
//     depth-gradient) and signals staged ONCE in shared memory as SoA planes (conflict-free: a warp
//     reads 32 consecutive floats of a plane row);
//   * each thread produces two vertically adjacent outputs so every tap value read from shared
//     memory is used twice (halves LDS traffic, the co-limiter next to the FP32/MUFU pipes);
//   * the spatial gaussian exponent is one FMA on a running tap offset; gaussian and depth term are merged into ONE
//     ex2.approx (exp(a)*exp(b) = exp2((a+b)*log2 e)); 1/max(dz*dist, eps) = min(inv_dz * rsqrt(dist^2), 1/eps) with the
//     guarded 1/dz staged per pixel; pow(x,128) is 7 squarings -- 2 MUFU and no table look-up per tap;
//   * out-of-image taps are stored as zero normals => clamp(dot, 1e-4, 1)^128 underflows to exactly
//     0, which reproduces the reference's `continue` without a branch;
//   * the diffuse and specular signals, which render.py:120-121 filters with identical guides, can
//     share one pass (NSIG = 2): weights are computed once.
#include "common.cuh"

namespace {

constexpr int TILE_W = 32;
constexpr int TILE_H = 16;
constexpr float FLT_EPS_ = 0.0001f;      // denoising.cu:12
constexpr float LOG2E = 1.4426950408889634f;

struct BilateralParams {
    TView nrm, zdz;
    TView sig[2];          // fwd: col ; bwd: out_grad (first 3 channels used)
    float *out[2];         // fwd: [B,H,W,4] ; bwd: [B,H,W,3]
    int B, H, W;
    int r;
    float neg_inv_2var_log2e;   // -log2(e) / (2 sigma^2)
};

// MUFU approximations without the range-handling wrappers of exp2f / __fdividef (their operands are bounded here: exponents
// <= 0, where a flushed denormal weight is indistinguishable from the reference's 1e-38; reciprocal arguments >= 1e-4)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 1 / max(dz * dist, 1e-4) = min(inv_dz * rsqrt(dist^2), 1e4) with inv_dz = dz > 0 ? 1/dz : +inf  (dist = 0 -> inf -> 1e4 as well)
__device__ __forceinline__ float guarded_inv(float dz) { return dz > 0.0f ? 1.0f / dz : INFINITY; }

__device__ __forceinline__ float pow128(float x)
{
    x *= x; x *= x; x *= x; x *= x; x *= x; x *= x; x *= x;
    return x;
}

template <int NSIG, bool BWD>
__global__ void __launch_bounds__(256) bilateral_kernel(BilateralParams p)
{
    extern __shared__ float smem[];
    const int r = p.r;
    const int tw = TILE_W + 2 * r, th = TILE_H + 2 * r;
    const int plane = tw * th;
    float *s_nx = smem, *s_ny = s_nx + plane, *s_nz = s_ny + plane, *s_z = s_nz + plane, *s_dz = s_z + plane;
    float *s_sig = s_dz + plane;                       // NSIG * 3 planes

    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * TILE_W - r, y0 = blockIdx.y * TILE_H - r;

    for (int i = tid; i < plane; i += 256) {
        int ty = i / tw, tx = i - ty * tw;
        int gy = y0 + ty, gx = x0 + tx;
        bool in = gy >= 0 && gx >= 0 && gy < p.H && gx < p.W;
        f3 n = F3(0.0f); float z = 0.0f, dz = 0.0f;
        if (in) {
            n = p.nrm.ld3(b, gy, gx);
            const float *q = p.zdz.p + p.zdz.off(b, gy, gx);
            z = __ldg(q); dz = __ldg(q + p.zdz.s3);
        }
        s_nx[i] = n.x; s_ny[i] = n.y; s_nz[i] = n.z; s_z[i] = z; s_dz[i] = guarded_inv(dz);     // plane holds 1/dz (guarded)
#pragma unroll
        for (int s = 0; s < NSIG; ++s) {
            f3 c = F3(0.0f);
            if (in) c = p.sig[s].ld3(b, gy, gx);
            s_sig[(3 * s + 0) * plane + i] = c.x; s_sig[(3 * s + 1) * plane + i] = c.y; s_sig[(3 * s + 2) * plane + i] = c.z;
        }
    }
    __syncthreads();

    // two vertically adjacent outputs per thread: rows 2*ty and 2*ty+1 of the tile
    const int lx = threadIdx.x, lyA = 2 * threadIdx.y;
    const int cxs = lx + r;
    f3 cn[2]; float cz[2], cdz[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        int ci = (lyA + o + r) * tw + cxs;
        cn[o] = F3(s_nx[ci], s_ny[ci], s_nz[ci]); cz[o] = s_z[ci]; cdz[o] = s_dz[ci];      // cdz = guarded 1/dz of the centre
    }
    float acc[2][NSIG][3]; float accw[2] = {0.0f, 0.0f};
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int s = 0; s < NSIG; ++s) acc[o][s][0] = acc[o][s][1] = acc[o][s][2] = 0.0f;

    const float k2 = p.neg_inv_2var_log2e;
    for (int rr = 0; rr <= 2 * r + 1; ++rr) {
        // tap row rr of the tile serves output A at vertical offset rr - r and output B at rr - r - 1; a row outside an output's
        // window gets exponent -inf (weight exactly 0)
        const float fyA = (float)(rr - r), fyB = (float)(rr - r - 1);
        const float fy2[2] = {fyA * fyA, fyB * fyB};
        const float gy[2] = {rr <= 2 * r ? fy2[0] * k2 : -INFINITY, rr >= 1 ? fy2[1] * k2 : -INFINITY};
        const int rowoff = (lyA + rr) * tw + lx;
        float fx = (float)(-r);
        for (int cx = 0; cx <= 2 * r; ++cx, fx += 1.0f) {
            const int i = rowoff + cx;
            const float fx2 = fx * fx;
            f3 tn = F3(s_nx[i], s_ny[i], s_nz[i]);
            float tz = s_z[i];
            float tinv = BWD ? s_dz[i] : 0.0f;
            float sg[NSIG][3];
#pragma unroll
            for (int s = 0; s < NSIG; ++s) {
                sg[s][0] = s_sig[(3 * s + 0) * plane + i]; sg[s][1] = s_sig[(3 * s + 1) * plane + i]; sg[s][2] = s_sig[(3 * s + 2) * plane + i];
            }
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                float wn = pow128(fminf(fmaxf(dot(tn, cn[o]), FLT_EPS_), 1.0f));
                // fwd: centre's dz (denoising.cu:59); bwd: tap's dz (denoising.cu:118)
                const float inv_den = fminf((BWD ? tinv : cdz[o]) * rsqrt_approx(fx2 + fy2[o]), 1.0f / FLT_EPS_);
                const float e = fmaf(fx2, k2, gy[o]) - LOG2E * (fabsf(tz - cz[o]) * inv_den);
                float w = wn * ex2_approx(e);
#pragma unroll
                for (int s = 0; s < NSIG; ++s) {
                    acc[o][s][0] = fmaf(sg[s][0], w, acc[o][s][0]);
                    acc[o][s][1] = fmaf(sg[s][1], w, acc[o][s][1]);
                    acc[o][s][2] = fmaf(sg[s][2], w, acc[o][s][2]);
                }
                accw[o] += w;
            }
        }
    }

    const int gx = blockIdx.x * TILE_W + lx;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const int gy = blockIdx.y * TILE_H + lyA + o;
        if (gx < p.W && gy < p.H) {
            int64_t px = ((int64_t)b * p.H + gy) * p.W + gx;
#pragma unroll
            for (int s = 0; s < NSIG; ++s) {
                if (!BWD) {
                    reinterpret_cast<float4 *>(p.out[s])[px] = make_float4(acc[o][s][0], acc[o][s][1], acc[o][s][2], fmaxf(accw[o], 0.0001f));
                } else {
                    float *d = p.out[s] + px * 3;
                    d[0] = acc[o][s][0]; d[1] = acc[o][s][1]; d[2] = acc[o][s][2];
                }
            }
        }
    }
}

static int check_guides(const mcs_tensor *nrm, const mcs_tensor *zdz, const mcs_tensor *sig, int sig_c, const char *sig_name)
{
    MCS_REQUIRE(view_ok(nrm) && view_ok(zdz) && view_ok(sig), "bilateral: null / empty tensor argument");
    MCS_REQUIRE(nrm->sizes[3] == 3, "bilateral: nrm must have 3 channels");
    MCS_REQUIRE(zdz->sizes[3] == 2, "bilateral: zdz must have 2 channels");
    MCS_REQUIRE(sig->sizes[3] == sig_c, "bilateral: %s must have %d channels", sig_name, sig_c);
    for (int d = 0; d < 3; ++d)
        MCS_REQUIRE(nrm->sizes[d] == sig->sizes[d] && zdz->sizes[d] == sig->sizes[d], "bilateral: shape mismatch in dim %d", d);
    return 0;
}

template <int NSIG, bool BWD>
static int launch_bilateral(BilateralParams &p, float sigma, cudaStream_t stream)
{
    MCS_REQUIRE(sigma > 0.0f, "bilateral: sigma must be > 0");
    p.r = 2 * (int)ceilf(sigma * 2.5f) + 1;                 // denoising.cu:28 filter_rad
    p.neg_inv_2var_log2e = -LOG2E / (2.0f * sigma * sigma);
    int tw = TILE_W + 2 * p.r, th = TILE_H + 2 * p.r;
    size_t smem = sizeof(float) * ((size_t)(5 + 3 * NSIG) * tw * th);
    MCS_REQUIRE(smem <= 227 * 1024, "bilateral: sigma %.3f needs a %zu-byte tile (> 227 KB shared memory)", sigma, smem);
    auto kern = bilateral_kernel<NSIG, BWD>;
    MCS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.W + TILE_W - 1) / TILE_W, (p.H + TILE_H - 1) / TILE_H, p.B), block(32, 8, 1);
    kern<<<grid, block, smem, stream>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" {

int mcs_bilateral_fwd(const mcs_tensor *col, const mcs_tensor *nrm, const mcs_tensor *zdz, float sigma, float *out, mcs_stream stream)
{
    if (int e = check_guides(nrm, zdz, col, 3, "col")) return e;
    BilateralParams p{};
    p.nrm = make_view(nrm); p.zdz = make_view(zdz); p.sig[0] = make_view(col); p.out[0] = out;
    p.B = col->sizes[0]; p.H = col->sizes[1]; p.W = col->sizes[2];
    return launch_bilateral<1, false>(p, sigma, (cudaStream_t)stream);
}

int mcs_bilateral_bwd(const mcs_tensor *nrm, const mcs_tensor *zdz, float sigma, const mcs_tensor *out_grad, float *col_grad, mcs_stream stream)
{
    if (int e = check_guides(nrm, zdz, out_grad, 4, "out_grad")) return e;
    BilateralParams p{};
    p.nrm = make_view(nrm); p.zdz = make_view(zdz); p.sig[0] = make_view(out_grad); p.out[0] = col_grad;
    p.B = nrm->sizes[0]; p.H = nrm->sizes[1]; p.W = nrm->sizes[2];
    return launch_bilateral<1, true>(p, sigma, (cudaStream_t)stream);
}

int mcs_bilateral_fwd2(const mcs_tensor *colA, const mcs_tensor *colB, const mcs_tensor *nrm, const mcs_tensor *zdz, float sigma,
                       float *outA, float *outB, mcs_stream stream)
{
    if (int e = check_guides(nrm, zdz, colA, 3, "colA")) return e;
    if (int e = check_guides(nrm, zdz, colB, 3, "colB")) return e;
    BilateralParams p{};
    p.nrm = make_view(nrm); p.zdz = make_view(zdz); p.sig[0] = make_view(colA); p.sig[1] = make_view(colB); p.out[0] = outA; p.out[1] = outB;
    p.B = nrm->sizes[0]; p.H = nrm->sizes[1]; p.W = nrm->sizes[2];
    return launch_bilateral<2, false>(p, sigma, (cudaStream_t)stream);
}

int mcs_bilateral_bwd2(const mcs_tensor *nrm, const mcs_tensor *zdz, float sigma, const mcs_tensor *out_gradA, const mcs_tensor *out_gradB,
                       float *col_gradA, float *col_gradB, mcs_stream stream)
{
    if (int e = check_guides(nrm, zdz, out_gradA, 4, "out_gradA")) return e;
    if (int e = check_guides(nrm, zdz, out_gradB, 4, "out_gradB")) return e;
    BilateralParams p{};
    p.nrm = make_view(nrm); p.zdz = make_view(zdz); p.sig[0] = make_view(out_gradA); p.sig[1] = make_view(out_gradB);
    p.out[0] = col_gradA; p.out[1] = col_gradB;
    p.B = nrm->sizes[0]; p.H = nrm->sizes[1]; p.W = nrm->sizes[2];
    return launch_bilateral<2, true>(p, sigma, (cudaStream_t)stream);
}

}  // extern "C"
