// raster.cu -- SURVEY section 8 row f2: primary visibility + attribute interpolation on the LBVH that the shadow rays already use, so the
// G-buffer of render/render.py:208-234 (dr.rasterize + dr.interpolate, nvdiffrast -- absent here) can be produced without a rasteriser.
//
//   k_rasterize   : one thread per pixel.  The pixel centre is un-projected with the inverse clip matrix to a near and a far point
//                   (NDC z = -1 / +1), the closest hit along that segment's line is found with bvh_closest (same fixed-order
//                   Moeller-Trumbore predicate as the shadow rays, ties broken by triangle id), and the result is written in
//                   nvdiffrast's `rast` convention: (u, v, z/w, triangle_id + 1), u / v = barycentric weights of vertex 0 / 1,
//                   0 in all channels for background.  Image row iy maps to NDC y = (iy + 0.5) / H * 2 - 1 (no flip, like dr.rasterize).
//   k_interpolate : out[b,y,x,:] = u * A[i0] + v * A[i1] + (1 - u - v) * A[i2]; backward scatters into dA with float atomics.
#include "ctx.h"
#include "bvh_traverse.cuh"

namespace {

struct RasterParams {
    BvhView bvh;
    const float *mtx, *inv;      // [B,4,4] row-major, clip = mtx * (p, 1)
    int B, H, W;
    float *rast;
};

__device__ __forceinline__ void mul4(const float *__restrict__ m, float x, float y, float z, float w, float (&o)[4])
{
#pragma unroll
    for (int r = 0; r < 4; ++r) o[r] = fmaf(__ldg(m + 4 * r), x, fmaf(__ldg(m + 4 * r + 1), y, fmaf(__ldg(m + 4 * r + 2), z, __ldg(m + 4 * r + 3) * w)));
}

__global__ void __launch_bounds__(128) k_rasterize(const RasterParams p)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)p.B * p.H * p.W;
    if (i >= n) return;
    const int ix = (int)(i % p.W); const int64_t t_ = i / p.W; const int iy = (int)(t_ % p.H), b = (int)(t_ / p.H);
    const float x = ((float)ix + 0.5f) / (float)p.W * 2.0f - 1.0f, y = ((float)iy + 0.5f) / (float)p.H * 2.0f - 1.0f;
    float a[4], c[4];
    mul4(p.inv + 16 * b, x, y, -1.0f, 1.0f, a);
    mul4(p.inv + 16 * b, x, y, 1.0f, 1.0f, c);
    const f3 o = F3(a[0] / a[3], a[1] / a[3], a[2] / a[3]);
    const f3 f = F3(c[0] / c[3], c[1] / c[3], c[2] / c[3]);
    const f3 d = f - o;
    float t, u, v;
    const int id = bvh_closest(p.bvh, o, d, t, u, v);
    float4 out = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (id >= 0) {
        const f3 h = o + d * t;
        float q[4];
        mul4(p.mtx + 16 * b, h.x, h.y, h.z, 1.0f, q);
        out = make_float4(1.0f - u - v, u, q[2] / q[3], (float)(id + 1));
    }
    reinterpret_cast<float4 *>(p.rast)[i] = out;
}

// inverse of B row-major 4x4 matrices, Gauss-Jordan with partial pivoting in fp64 (one thread per matrix; a singular matrix gives NaNs,
// i.e. an all-background image)
__global__ void k_invert4(const float *__restrict__ m, float *__restrict__ inv, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { a[r][c] = (double)m[16 * b + 4 * r + c]; a[r][4 + c] = r == c ? 1.0 : 0.0; }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r) if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        for (int c = 0; c < 8; ++c) { const double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
        const double d = 1.0 / a[col][col];
        for (int c = 0; c < 8; ++c) a[col][c] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = a[r][col];
            for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[16 * b + 4 * r + c] = (float)a[r][4 + c];
}

struct InterpParams {
    const float *attr; int64_t attr_bs; int V, C;
    const int32_t *tris; int T;
    const float4 *rast;
    const float *dout; float *out, *dattr;
    int64_t npx, px_per_batch;
};

template <bool BWD>
__global__ void __launch_bounds__(256) k_interpolate(const InterpParams p)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.npx) return;
    const float4 r = __ldg(p.rast + i);
    const int id = (int)r.w - 1;
    if (id < 0 || id >= p.T) {
        if (!BWD) for (int c = 0; c < p.C; ++c) p.out[i * p.C + c] = 0.0f;
        return;
    }
    const int i0 = __ldg(p.tris + 3 * (size_t)id), i1 = __ldg(p.tris + 3 * (size_t)id + 1), i2 = __ldg(p.tris + 3 * (size_t)id + 2);
    const int64_t base = (i / p.px_per_batch) * p.attr_bs;
    const float w0 = r.x, w1 = r.y, w2 = 1.0f - r.x - r.y;
    for (int c = 0; c < p.C; ++c) {
        if (!BWD) {
            const float *A = p.attr + base;
            p.out[i * p.C + c] = fmaf(w0, __ldg(A + (size_t)i0 * p.C + c), fmaf(w1, __ldg(A + (size_t)i1 * p.C + c), w2 * __ldg(A + (size_t)i2 * p.C + c)));
        } else {
            const float g = __ldg(p.dout + i * p.C + c);
            float *D = p.dattr + base;
            atomicAdd(D + (size_t)i0 * p.C + c, w0 * g); atomicAdd(D + (size_t)i1 * p.C + c, w1 * g); atomicAdd(D + (size_t)i2 * p.C + c, w2 * g);
        }
    }
}

// Nearest-texel material fetch (stand-in for dr.texture with filter_mode='nearest', render/texture.py:66-75): out[i,:] = tex[idx[i],:];
// the backward pass scatters with float atomics (torch's index backward sorts the 2 M indices first and is ~8x slower).
template <bool BWD>
__global__ void __launch_bounds__(256) k_texel_fetch(const float *__restrict__ tex, const int64_t *__restrict__ idx, int64_t n, int C, int64_t T,
                                                     float *__restrict__ out, const float *__restrict__ dout, float *__restrict__ dtex)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t t = __ldg(idx + i);
    const bool ok = t >= 0 && t < T;
    for (int c = 0; c < C; ++c) {
        if (!BWD) out[i * C + c] = ok ? __ldg(tex + t * C + c) : 0.0f;
        else if (ok) atomicAdd(dtex + t * C + c, __ldg(dout + i * C + c));
    }
}

}  // namespace

extern "C" {

int mcs_rasterize(mcs_ctx *c, const float *mtx, int32_t B, int32_t H, int32_t W, float *rast, mcs_stream stream)
{
    MCS_REQUIRE(c && c->T > 0, "mcs_rasterize: no acceleration structure built (call mcs_bvh_build first)");
    MCS_REQUIRE(mtx && rast && B > 0 && H > 0 && W > 0, "mcs_rasterize: bad arguments");
    if (int e = mcs_buf_reserve(c->mtx_inv, sizeof(float) * 16 * (size_t)B, (cudaStream_t)stream)) return e;
    float *inv_mtx = (float *)c->mtx_inv.p;
    k_invert4<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(mtx, inv_mtx, B);
    MCS_LAUNCH_CHECK();
    RasterParams p{};
    p.bvh = BvhView{(const float4 *)c->nodes.p, (const float4 *)c->tris.p, nullptr, nullptr};
    p.mtx = mtx; p.inv = inv_mtx; p.B = B; p.H = H; p.W = W; p.rast = rast;
    const int64_t n = (int64_t)B * H * W;
    k_rasterize<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}

static int interp_common(InterpParams &p, const float *attr, int64_t attr_batch_stride, int32_t V, int32_t C, const int32_t *tris, int32_t T,
                         const float *rast, int32_t B, int32_t H, int32_t W)
{
    MCS_REQUIRE(attr && tris && rast && V > 0 && C > 0 && T > 0 && B > 0 && H > 0 && W > 0, "mcs_interpolate: bad arguments");
    p.attr = attr; p.attr_bs = attr_batch_stride; p.V = V; p.C = C; p.tris = tris; p.T = T; p.rast = (const float4 *)rast;
    p.px_per_batch = (int64_t)H * W; p.npx = p.px_per_batch * B;
    return 0;
}

int mcs_interpolate_fwd(const float *attr, int64_t attr_batch_stride, int32_t V, int32_t C, const int32_t *tris, int32_t T, const float *rast,
                        int32_t B, int32_t H, int32_t W, float *out, mcs_stream stream)
{
    InterpParams p{};
    if (int e = interp_common(p, attr, attr_batch_stride, V, C, tris, T, rast, B, H, W)) return e;
    MCS_REQUIRE(out != nullptr, "mcs_interpolate_fwd: null output");
    p.out = out;
    k_interpolate<false><<<(unsigned)((p.npx + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}

int mcs_interpolate_bwd(const float *attr, int64_t attr_batch_stride, int32_t V, int32_t C, const int32_t *tris, int32_t T, const float *rast,
                        int32_t B, int32_t H, int32_t W, const float *d_out, float *d_attr, mcs_stream stream)
{
    InterpParams p{};
    if (int e = interp_common(p, attr, attr_batch_stride, V, C, tris, T, rast, B, H, W)) return e;
    MCS_REQUIRE(d_out && d_attr, "mcs_interpolate_bwd: null gradient pointer");
    p.dout = d_out; p.dattr = d_attr;
    k_interpolate<true><<<(unsigned)((p.npx + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}

int mcs_texel_fetch_fwd(const float *tex, int64_t T, int32_t C, const int64_t *idx, int64_t n, float *out, mcs_stream stream)
{
    MCS_REQUIRE(tex && idx && out && T > 0 && C > 0 && n >= 0, "mcs_texel_fetch_fwd: bad arguments");
    if (n == 0) return 0;
    k_texel_fetch<false><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tex, idx, n, C, T, out, nullptr, nullptr);
    MCS_LAUNCH_CHECK();
    return 0;
}

int mcs_texel_fetch_bwd(int64_t T, int32_t C, const int64_t *idx, int64_t n, const float *d_out, float *d_tex, mcs_stream stream)
{
    MCS_REQUIRE(idx && d_out && d_tex && T > 0 && C > 0 && n >= 0, "mcs_texel_fetch_bwd: bad arguments");
    if (n == 0) return 0;
    k_texel_fetch<true><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nullptr, idx, n, C, T, nullptr, d_out, d_tex);
    MCS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
