// light.cu -- EnvironmentLight.update_pdf (render/light.py:46-59) as two launches instead of ~12 tiny torch kernels.
//
//   pdf  = max_c(base) * sin(pi * (y + 0.5) / H), normalised by its grand total
//   cols = row-wise inclusive cumsum of pdf, each row divided by its last element (if > 0)
//   rows = inclusive cumsum of the row totals divided by its last element (if > 0)          [returned as a vector of H]
//
// k_light_rows: one CTA per probe row.  Each thread owns a contiguous chunk of the row, the chunk prefix is sequential, chunk
//   offsets come from a shuffle scan; all sums are carried in fp64 and rounded once, so the fp32 CDF is monotone and independent
//   of the scan shape (the reference's torch.cumsum is an fp32 tree scan; the oracle is an fp32 sequential loop: both are within
//   a few 1e-7 of this).  Writes the un-normalised pdf, the normalised cols row and the fp64 row total.
// k_light_finish: every CTA reduces the H row totals in the same fixed order (bitwise identical grand total in every CTA), scales its
//   slice of pdf; CTA 0 additionally scans the row totals into `rows`.
// HBM traffic: reads 12 B/texel, writes 8 B/texel (+ 8 B/texel for the pdf rescale): 7 MB at 512x512, latency-bound at 256x256.
#include "common.cuh"

namespace {

constexpr int LB = 256;

struct LightParams {
    TView base;
    int H, W;
    float *pdf, *rows, *cols;
    double *rowtot;
};

__device__ __forceinline__ double shfl_up_d(double v, int d)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xFFFFFFFFu, lo, d); hi = __shfl_up_sync(0xFFFFFFFFu, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_d(double v, int d)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xFFFFFFFFu, lo, d); hi = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
    return __hiloint2double(hi, lo);
}

// exclusive prefix of one value per thread over the CTA (LB threads); `total` = sum over all threads
__device__ __forceinline__ double block_exclusive(double v, double *sm /* [LB/32] */, double &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double o = shfl_up_d(inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    double woff = 0.0, tot = 0.0;
#pragma unroll
    for (int w = 0; w < LB / 32; ++w) {
        const double t = sm[w];
        if (w < warp) woff += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return woff + inc - v;
}

__global__ void __launch_bounds__(LB) k_light_rows(const LightParams p)
{
    __shared__ double sm[LB / 32];
    __shared__ float s_sin;
    const int y = blockIdx.x;
    if (threadIdx.x == 0) {
        const float Y = __fdiv_rn((float)y + 0.5f, (float)p.H);             // util.pixel_grid, render/util.py:62-66
        s_sin = (float)sin((double)__fmul_rn(Y, 3.14159274101257324f));    // torch.sin(Y * np.pi) on an fp32 tensor
    }
    __syncthreads();
    const float sn = s_sin;
    const int per = (p.W + LB - 1) / LB;
    const int x0 = threadIdx.x * per, x1 = min(p.W, x0 + per);
    double local = 0.0;
    for (int x = x0; x < x1; ++x) {
        const f3 b = p.base.ld3(0, y, x);
        const float v = __fmul_rn(fmaxf(b.x, fmaxf(b.y, b.z)), sn);
        p.pdf[(size_t)y * p.W + x] = v;
        local += (double)v;
    }
    double total;
    double run = block_exclusive(local, sm, total);
    const double dn = total > 0.0 ? total : 1.0;
    for (int x = x0; x < x1; ++x) {
        run += (double)p.pdf[(size_t)y * p.W + x];
        p.cols[(size_t)y * p.W + x] = (float)(run / dn);
    }
    if (threadIdx.x == 0) p.rowtot[y] = total;
}

__global__ void __launch_bounds__(LB) k_light_finish(const LightParams p)
{
    __shared__ double sm[LB / 32];
    const int per = (p.H + LB - 1) / LB;
    const int y0 = threadIdx.x * per, y1 = min(p.H, y0 + per);
    double local = 0.0;
    for (int y = y0; y < y1; ++y) local += p.rowtot[y];
    double total;
    double run = block_exclusive(local, sm, total);
    if (blockIdx.x == 0) {
        const double dn = total > 0.0 ? total : 1.0;
        for (int y = y0; y < y1; ++y) {
            run += p.rowtot[y];
            p.rows[y] = (float)(run / dn);
        }
    }
    // pdf / sum(pdf): an all-black probe divides by zero exactly like the reference (NaN pdf, light.py:51)
    const float tf = (float)total;
    const size_t n = (size_t)p.H * p.W;
    for (size_t i = (size_t)blockIdx.x * LB + threadIdx.x; i < n; i += (size_t)gridDim.x * LB) p.pdf[i] = __fdiv_rn(p.pdf[i], tf);
}

}  // namespace

int mcs_update_pdf(const mcs_tensor *base, float *pdf, float *rows, float *cols, double *row_totals, mcs_stream s)
{
    MCS_REQUIRE(view_ok(base) && pdf && rows && cols && row_totals, "update_pdf: null / empty argument");
    MCS_REQUIRE(base->sizes[0] == 1 && base->sizes[3] == 3, "update_pdf: base must be a [1,H,W,3] view");
    LightParams p{};
    p.base = make_view(base); p.H = base->sizes[1]; p.W = base->sizes[2];
    MCS_REQUIRE(p.H >= 1 && p.W >= 1 && p.H < 32768 && p.W < 65536, "update_pdf: probe resolution out of range");
    p.pdf = pdf; p.rows = rows; p.cols = cols; p.rowtot = row_totals;
    k_light_rows<<<p.H, LB, 0, (cudaStream_t)s>>>(p);
    MCS_LAUNCH_CHECK();
    const size_t n = (size_t)p.H * p.W;
    const int grid = (int)((n + (size_t)LB * 8 - 1) / ((size_t)LB * 8));
    k_light_finish<<<grid < 1 ? 1 : (grid > 1184 ? 1184 : grid), LB, 0, (cudaStream_t)s>>>(p);
    MCS_LAUNCH_CHECK();
    return 0;
}
