// lossmesh.cu -- the two cheap brackets of the hot path, B200 versions:
//   * image_loss  : tonemap (log-sRGB) + {L1, MSE, RELMSE, SMAPE, N2N} per pixel with a deterministic two-level reduction
//                   (replaces imgLossFwdKernel / imgLossBwdKernel, render/renderutils/c_src/loss.cu:105-227, and
//                   image_loss_fwd/_bwd, torch_bindings.cpp:739-800).  FIX: "n2n" is honoured (the reference's strToLoss,
//                   torch_bindings.cpp:727-737, has no "n2n" case and silently computes L1 on the CUDA path).
//   * xfm_points / xfm_vectors : batched 4x4 transform of [1|B, V, 3] points (replaces xfmPointsFwd/BwdKernel,
//                   render/renderutils/c_src/mesh.cu:19-90, and xfm_fwd/_bwd, torch_bindings.cpp:803-864).
// Both are HBM-streaming: image_loss reads 24 B/px (fwd) / writes 24 B/px more (bwd); xfm reads 12 B and writes 16 B per vertex.
#include "common.cuh"

namespace {

enum { LOSS_L1 = 0, LOSS_MSE = 1, LOSS_RELMSE = 2, LOSS_SMAPE = 3, LOSS_N2N = 4 };
constexpr int LOSS_BLOCK = 256;

__device__ __forceinline__ float bwd_abs(float x) { return x == 0.0f ? 0.0f : (x < 0.0f ? -1.0f : 1.0f); }        // loss.cu:17
// loss.cu:28-41
__device__ __forceinline__ float fwd_srgb(float x) { return x > 0.0031308f ? powf(fmaxf(x, 0.0031308f), 1.0f / 2.4f) * 1.055f - 0.055f : 12.92f * fmaxf(x, 0.0f); }
__device__ __forceinline__ float bwd_srgb(float x, float d_out)
{
    if (x > 0.0031308f) return d_out * 0.439583f / powf(x, 0.583333f);
    if (x > 0.0f) return d_out * 12.92f;
    return 0.0f;
}
__device__ __forceinline__ float fwd_tonemap(float x) { return fwd_srgb(logf(x + 1.0f)); }                          // loss.cu:43-46
__device__ __forceinline__ float bwd_tonemap(float x, float d_out)                                                  // loss.cu:48-65
{
    if (x > 0.0f && x < 65535.0f) return bwd_srgb(logf(x + 1.0f), d_out) * (1.0f / (x + 1.0f));
    return 0.0f;
}

__device__ __forceinline__ float loss_fwd1(int loss, float img, float tgt)
{
    const float eps = 0.01f;
    const float d = img - tgt;
    switch (loss) {
    case LOSS_MSE: return d * d;
    case LOSS_RELMSE: return d * d / (img * img + tgt * tgt + eps);             // loss.cu:67-70
    case LOSS_N2N: return d * d / (img * img + eps);                            // loss.cu:79-82
    case LOSS_SMAPE: return fabsf(d) / (img + tgt + eps);                       // loss.cu:92-95
    default: return fabsf(d);
    }
}
__device__ __forceinline__ void loss_bwd1(int loss, float img, float tgt, float dv, float &d_img, float &d_tgt)
{
    const float eps = 0.01f;
    const float d = img - tgt;
    switch (loss) {
    case LOSS_MSE: d_img = dv * 2.0f * d; d_tgt = -d_img; break;                                            // loss.cu:181-185
    case LOSS_RELMSE: {                                                                                     // loss.cu:72-77
        const float den = tgt * tgt + img * img + eps;
        d_img = dv * 2.0f * d * (tgt * (tgt + img) + eps) / (den * den);
        d_tgt = -dv * 2.0f * d * (img * (tgt + img) + eps) / (den * den);
    } break;
    case LOSS_N2N: {                                                                                        // loss.cu:84-89
        const float den = img * img + eps;
        d_img = dv * 2.0f * d / den; d_tgt = -d_img;
    } break;
    case LOSS_SMAPE: {                                                                                      // loss.cu:97-102
        const float den = tgt + img + eps;
        d_img = dv * bwd_abs(d) * (2.0f * tgt + eps) / (den * den);
        d_tgt = -dv * bwd_abs(d) * (2.0f * img + eps) / (den * den);
    } break;
    default: d_img = dv * bwd_abs(d); d_tgt = -d_img; break;
    }
}

struct LossParams {
    TView img, tgt;
    int N, H, W; int64_t npx;
    int loss, tonemap;
    float *partial;              // fwd: [nblocks] per-CTA sums of (sum_c loss)/3
    TView dout; int dout_n;      // bwd: upstream gradient per CTA partial (or a single broadcast value)
    float *d_img, *d_tgt;        // bwd: contiguous [N,H,W,3]
};

__global__ void __launch_bounds__(LOSS_BLOCK) k_image_loss_fwd(LossParams p)
{
    __shared__ float s_part[LOSS_BLOCK / 32];
    const int64_t px = (int64_t)blockIdx.x * LOSS_BLOCK + threadIdx.x;
    float v = 0.0f;
    if (px < p.npx) {
        const int w = (int)(px % p.W); const int64_t t = px / p.W; const int h = (int)(t % p.H), n = (int)(t / p.H);
        f3 a = p.img.ld3(n, h, w), b = p.tgt.ld3(n, h, w);
        float ia[3] = {a.x, a.y, a.z}, tb[3] = {b.x, b.y, b.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float x = clampf(ia[c], 0.0f, 65535.0f), y = clampf(tb[c], 0.0f, 65535.0f);       // loss.cu:118-119 (always, not only when tonemapping)
            if (p.tonemap) { x = fwd_tonemap(x); y = fwd_tonemap(y); }
            v += loss_fwd1(p.loss, x, y);
        }
        v *= (1.0f / 3.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < LOSS_BLOCK / 32; ++i) s += s_part[i];
        p.partial[blockIdx.x] = s;              // fixed order: deterministic
    }
}

__global__ void __launch_bounds__(LOSS_BLOCK) k_image_loss_bwd(LossParams p)
{
    const int64_t px = (int64_t)blockIdx.x * LOSS_BLOCK + threadIdx.x;
    if (px >= p.npx) return;
    const int w = (int)(px % p.W); const int64_t t = px / p.W; const int h = (int)(t % p.H), n = (int)(t / p.H);
    const float d_out = __ldg(p.dout.p + (p.dout_n == 1 ? 0 : (int64_t)blockIdx.x * p.dout.s0));
    f3 a = p.img.ld3(n, h, w), b = p.tgt.ld3(n, h, w);
    float ia[3] = {a.x, a.y, a.z}, tb[3] = {b.x, b.y, b.z}, gi[3], gt[3];
    const float dv = d_out * (1.0f / 3.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = ia[c], y = tb[c];
        if (p.tonemap) { x = fwd_tonemap(x); y = fwd_tonemap(y); }                 // loss.cu:163-167 (no clamp in the replay, as in the reference)
        float di, dt;
        loss_bwd1(p.loss, x, y, dv, di, dt);
        if (p.tonemap) { di = bwd_tonemap(ia[c], di); dt = bwd_tonemap(tb[c], dt); }
        if (ia[c] <= 0.0f || ia[c] >= 65535.0f) di = 0.0f;                          // loss.cu:217-222
        if (tb[c] <= 0.0f || tb[c] >= 65535.0f) dt = 0.0f;
        gi[c] = di; gt[c] = dt;
    }
    float *o1 = p.d_img + px * 3, *o2 = p.d_tgt + px * 3;
    o1[0] = gi[0]; o1[1] = gi[1]; o1[2] = gi[2];
    o2[0] = gt[0]; o2[1] = gt[1]; o2[2] = gt[2];
}

struct XfmParams {
    const float *points; int p_s0, p_s1, p_s2; int p_b;     // [1|B, V, 3]
    const float *matrix; int m_s0, m_s1, m_s2;              // [B, 4, 4]
    const float *dout; int d_s0, d_s1, d_s2;                // bwd: [B, V, 4|3]
    float *out;                                             // fwd: contiguous [B,V,4] (points) or [B,V,3] (vectors); bwd: [B,V,3]
    int B, V, is_points;
};

__global__ void __launch_bounds__(256) k_xfm_fwd(XfmParams p)
{
    __shared__ float m[4][4];       // m[r][c] = matrix[b][r][c]
    const int b = blockIdx.y;
    if (threadIdx.x < 16) m[threadIdx.x / 4][threadIdx.x % 4] = __ldg(p.matrix + (int64_t)b * p.m_s0 + (threadIdx.x / 4) * p.m_s1 + (threadIdx.x % 4) * p.m_s2);
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.V) return;
    const float *q = p.points + (int64_t)(p.p_b == 1 ? 0 : b) * p.p_s0 + (int64_t)v * p.p_s1;
    const float x = __ldg(q), y = __ldg(q + p.p_s2), z = __ldg(q + 2 * p.p_s2);
    // out = [x y z w] * M^T  (ops.py:515: matmul(pad(points), transpose(matrix)))
    if (p.is_points) {
        float4 o;
        o.x = x * m[0][0] + y * m[0][1] + z * m[0][2] + m[0][3];
        o.y = x * m[1][0] + y * m[1][1] + z * m[1][2] + m[1][3];
        o.z = x * m[2][0] + y * m[2][1] + z * m[2][2] + m[2][3];
        o.w = x * m[3][0] + y * m[3][1] + z * m[3][2] + m[3][3];
        reinterpret_cast<float4 *>(p.out)[(int64_t)b * p.V + v] = o;
    } else {
        float *o = p.out + ((int64_t)b * p.V + v) * 3;
        o[0] = x * m[0][0] + y * m[0][1] + z * m[0][2];
        o[1] = x * m[1][0] + y * m[1][1] + z * m[1][2];
        o[2] = x * m[2][0] + y * m[2][1] + z * m[2][2];
    }
}

__global__ void __launch_bounds__(256) k_xfm_bwd(XfmParams p)
{
    __shared__ float m[4][4];
    const int b = blockIdx.y;
    if (threadIdx.x < 16) m[threadIdx.x / 4][threadIdx.x % 4] = __ldg(p.matrix + (int64_t)b * p.m_s0 + (threadIdx.x / 4) * p.m_s1 + (threadIdx.x % 4) * p.m_s2);
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.V) return;
    const float *g = p.dout + (int64_t)b * p.d_s0 + (int64_t)v * p.d_s1;
    const float gx = __ldg(g), gy = __ldg(g + p.d_s2), gz = __ldg(g + 2 * p.d_s2), gw = p.is_points ? __ldg(g + 3 * p.d_s2) : 0.0f;
    float *o = p.out + ((int64_t)b * p.V + v) * 3;                   // full-batch gradient; a broadcast input is reduced by the caller
    o[0] = gx * m[0][0] + gy * m[1][0] + gz * m[2][0] + gw * m[3][0];
    o[1] = gx * m[0][1] + gy * m[1][1] + gz * m[2][1] + gw * m[3][1];
    o[2] = gx * m[0][2] + gy * m[1][2] + gz * m[2][2] + gw * m[3][2];
}

static int loss_common(LossParams &p, const mcs_tensor *img, const mcs_tensor *target, int32_t loss, int32_t tonemapper)
{
    MCS_REQUIRE(view_ok(img) && view_ok(target), "image_loss: null / empty tensor argument");
    MCS_REQUIRE((img->sizes[3] == 3 || img->sizes[3] == 1) && (target->sizes[3] == 3 || target->sizes[3] == 1), "image_loss: img/target must have 3 channels");
    MCS_REQUIRE(loss >= 0 && loss <= 4, "image_loss: unknown loss id %d (0 l1, 1 mse, 2 relmse, 3 smape, 4 n2n)", loss);
    p.N = img->sizes[0] > target->sizes[0] ? img->sizes[0] : target->sizes[0];
    p.H = img->sizes[1] > target->sizes[1] ? img->sizes[1] : target->sizes[1];
    p.W = img->sizes[2] > target->sizes[2] ? img->sizes[2] : target->sizes[2];
    for (int d = 0; d < 3; ++d) {
        const int full = d == 0 ? p.N : (d == 1 ? p.H : p.W);
        MCS_REQUIRE((img->sizes[d] == full || img->sizes[d] == 1) && (target->sizes[d] == full || target->sizes[d] == 1), "image_loss: shapes are not broadcastable");
    }
    p.npx = (int64_t)p.N * p.H * p.W;
    p.img = make_view(img); p.tgt = make_view(target);
    p.loss = loss; p.tonemap = tonemapper;
    return 0;
}

}  // namespace

extern "C" {

int mcs_image_loss_num_partials(int32_t N, int32_t H, int32_t W) { return (int)(((int64_t)N * H * W + LOSS_BLOCK - 1) / LOSS_BLOCK); }

int mcs_image_loss_fwd(const mcs_tensor *img, const mcs_tensor *target, int32_t loss, int32_t tonemapper, float *partials, mcs_stream s)
{
    LossParams p{};
    if (int e = loss_common(p, img, target, loss, tonemapper)) return e;
    MCS_REQUIRE(partials != nullptr, "image_loss_fwd: null output");
    p.partial = partials;
    const int nb = mcs_image_loss_num_partials(p.N, p.H, p.W);
    k_image_loss_fwd<<<nb, LOSS_BLOCK, 0, (cudaStream_t)s>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}

int mcs_image_loss_bwd(const mcs_tensor *img, const mcs_tensor *target, int32_t loss, int32_t tonemapper, const mcs_tensor *d_partials,
                       float *d_img, float *d_target, mcs_stream s)
{
    LossParams p{};
    if (int e = loss_common(p, img, target, loss, tonemapper)) return e;
    MCS_REQUIRE(view_ok(d_partials) && d_img && d_target, "image_loss_bwd: null argument");
    const int nb = mcs_image_loss_num_partials(p.N, p.H, p.W);
    MCS_REQUIRE(d_partials->sizes[0] == nb || d_partials->sizes[0] == 1, "image_loss_bwd: gradient must have one value per partial sum (%d)", nb);
    p.dout = make_view(d_partials); p.dout_n = d_partials->sizes[0];
    p.d_img = d_img; p.d_tgt = d_target;
    k_image_loss_bwd<<<nb, LOSS_BLOCK, 0, (cudaStream_t)s>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}

static int xfm_common(XfmParams &p, const mcs_tensor *points, const mcs_tensor *matrix)
{
    MCS_REQUIRE(view_ok(points) && view_ok(matrix), "xfm: null / empty tensor argument");
    // points: sizes (1|B, V, 3, 1); matrix: (B, 4, 4, 1)
    MCS_REQUIRE(points->sizes[2] == 3 && matrix->sizes[1] == 4 && matrix->sizes[2] == 4, "xfm: points must be [1|B,V,3] and matrix [B,4,4]");
    MCS_REQUIRE(points->sizes[0] == 1 || points->sizes[0] == matrix->sizes[0], "xfm: points batch must be 1 or match the matrix batch");
    p.points = (const float *)points->ptr; p.p_s0 = points->strides[0]; p.p_s1 = points->strides[1]; p.p_s2 = points->strides[2]; p.p_b = points->sizes[0];
    p.matrix = (const float *)matrix->ptr; p.m_s0 = matrix->strides[0]; p.m_s1 = matrix->strides[1]; p.m_s2 = matrix->strides[2];
    p.B = matrix->sizes[0]; p.V = points->sizes[1];
    MCS_REQUIRE(p.B <= 65535, "xfm: batch too large");
    return 0;
}

int mcs_xfm_fwd(const mcs_tensor *points, const mcs_tensor *matrix, int32_t is_points, float *out, mcs_stream s)
{
    XfmParams p{};
    if (int e = xfm_common(p, points, matrix)) return e;
    MCS_REQUIRE(out != nullptr, "xfm_fwd: null output");
    p.out = out; p.is_points = is_points;
    k_xfm_fwd<<<dim3((p.V + 255) / 256, p.B), 256, 0, (cudaStream_t)s>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}

int mcs_xfm_bwd(const mcs_tensor *points, const mcs_tensor *matrix, const mcs_tensor *d_out, int32_t is_points, float *d_points, mcs_stream s)
{
    XfmParams p{};
    if (int e = xfm_common(p, points, matrix)) return e;
    MCS_REQUIRE(view_ok(d_out) && d_points, "xfm_bwd: null argument");
    MCS_REQUIRE(d_out->sizes[0] == p.B && d_out->sizes[1] == p.V && d_out->sizes[2] == (is_points ? 4 : 3), "xfm_bwd: upstream gradient shape mismatch");
    p.dout = (const float *)d_out->ptr; p.d_s0 = d_out->strides[0]; p.d_s1 = d_out->strides[1]; p.d_s2 = d_out->strides[2];
    p.out = d_points; p.is_points = is_points;
    k_xfm_bwd<<<dim3((p.V + 255) / 256, p.B), 256, 0, (cudaStream_t)s>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
