/*
 * mcshade.h -- C ABI of libmcshade.so, the B200-native (sm_100a) replacement for the nvdiffrecmc
 * per-iteration hot path.  No torch / pybind types cross this boundary: plain device pointers,
 * sizes, element strides and a CUDA stream handle.  Every entry point returns 0 on success and a
 * non-zero code on failure (mcs_last_error() then holds a message); unlike the reference
 * (render/optixutils/c_src/common.h:37-61, errors formatted then dropped) nothing fails silently.
 * All work is enqueued on `stream` and NO entry point synchronises the host (the reference forces a
 * cudaStreamSynchronize after every env_shade launch, optixutils/c_src/torch_bindings.cpp:185,269).
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference
 * repository root).
 */
#ifndef MCSHADE_H
#define MCSHADE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCS_ABI_VERSION 2

/* Strided NHWC view of fp32 (or int32) device memory.  sizes/strides are in ELEMENTS, dims are
 * (N, H, W, C); a size-1 dimension broadcasts (stride ignored), exactly like the reference's
 * accessors (optixutils/c_src/common.h:13-27 fetch3; renderutils/c_src/tensor.h:32 nhwcIndex). */
typedef struct mcs_tensor {
    const void *ptr;
    int32_t sizes[4];
    int32_t strides[4];
} mcs_tensor;

typedef struct mcs_ctx mcs_ctx;       /* opaque; replaces OptiXStateWrapper (optixutils/c_src/optix_wrapper.h:17-37) */
typedef void *mcs_stream;             /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
int mcs_abi_version(void);
const char *mcs_last_error(void);     /* thread-local message of the last failing call */

/* ---- context: replaces OptiXStateWrapper(path, cuda_home) ctor/dtor,
 *      optixutils/c_src/optix_wrapper.cpp:306-348 (no NVRTC, no OptiX: nothing to compile at run time) */
int mcs_ctx_create(mcs_ctx **out);
int mcs_ctx_destroy(mcs_ctx *ctx);

/* ---- acceleration structure: replaces optix_build_bvh(state, verts, tris, rebuild),
 *      optixutils/c_src/torch_bindings.cpp:37-116 (optixAccelBuild).
 *      verts: V x 3 fp32 contiguous, tris: T x 3 int32 contiguous (device).  rebuild != 0: full LBVH
 *      build (Morton sort + Karras topology + refit); rebuild == 0: refit boxes on the existing
 *      topology (OPTIX_BUILD_OPERATION_UPDATE).  Runs on `stream` (the reference uses stream 0). */
int mcs_bvh_build(mcs_ctx *ctx, const float *verts, int32_t V, const int32_t *tris, int32_t T, uint32_t rebuild, mcs_stream stream);

/* Test / inspection hook: copies the binary LBVH (sorted Morton keys, sorted->original triangle
 * ids, Karras children, padded node boxes) into caller-provided DEVICE buffers of sizes
 * T, T, T-1, T-1, (2T-1)*3, (2T-1)*3.  Node ids: internal 0..T-2, leaf j = T-1+j. */
int mcs_bvh_export(mcs_ctx *ctx, uint32_t *morton, int32_t *prim, int32_t *left, int32_t *right, float *lo, float *hi, mcs_stream stream);

/* Any-hit visibility of n rays (origin, direction; t in (0, 1e16)), the "integer visibility mask":
 * vis[i] = 1 if nothing is hit.  Same predicate as the shadow rays inside env_shade; replaces
 * shadow_test()/optixTrace, optixutils/c_src/envsampling/kernel.cu:101-118. */
int mcs_trace_visibility(mcs_ctx *ctx, const float *ro, const float *rd, int64_t n, uint8_t *vis, mcs_stream stream);

/* Closest hit (primary visibility for the synthetic G-buffer producer):
 * tri_id[i] = original triangle id or -1; tuv[i] = (t, u, v). */
int mcs_trace_closest(mcs_ctx *ctx, const float *ro, const float *rd, int64_t n, int32_t *tri_id, float *tuv, mcs_stream stream);

/* ---- fused env-light importance sampling + shadow rays + BSDF:
 *      replaces env_shade_fwd / env_shade_bwd, optixutils/c_src/torch_bindings.cpp:123-188 / 190-272
 *      (optixLaunch of __raygen__rg, envsampling/kernel.cu:463-542).
 *      mask [B,H,W,1]; ro, gb_pos, gb_normal, gb_kd, gb_ks [B,H,W,3]; gb_view_pos broadcastable
 *      [B|1,H|1,W|1,3]; light [1,Hl,Wl,3]; pdf [1,Hl,Wl,1]; rows [1,Hl,1,1]; cols [1,Hl,Wl,1];
 *      perms int32 [1,P,1,N*N] (all as mcs_tensor views, arbitrary strides).
 *      bsdf: 0 'pbr', 1 'diffuse', 2 'white' (optixutils/ops.py:136).
 *      batch_offset is added to the batch index inside the per-pixel RNG hash so a rank that holds
 *      views [o, o+B) of a larger batch reproduces the single-GPU random stream (kernel.cu:504).
 *      Outputs are contiguous [B,H,W,3] fp32 and are fully written (masked pixels = 0).
 *      hit_record (optional, may be NULL): uint32 [B,H,W,ceil(2*n_samples_x^2/32)], receives one bit per sample slot
 *      w (w < N^2: light sample of stratum w; else BSDF sample of stratum w-N^2): 1 = shadow ray occluded.  Passing the
 *      same buffer to mcs_env_shade_bwd (same seed, same inputs) lets the backward pass REPLAY visibility instead of
 *      re-tracing every ray as the reference does (torch_bindings.cpp:266-267 launches the same program with backward=1).
 *      rec_count / rec_rays (optional, together): the full RAY RECORD -- rec_count uint32 [B,H,W] = number of rays that were
 *      evaluated for the pixel (visible, or occluded with shadow_scale < 1), rec_rays fp32 [B,H,W,5,rec_slots] =
 *      (dx, dy, dz, MIS weight, env texel | occluded << 31) in evaluation order, rec_slots >= 2*n_samples_x^2.  With it the
 *      backward pass needs neither sampling nor traversal: see mcs_env_shade_bwd_replay.
 *      seed_offset_dev (optional, may be NULL): DEVICE pointer to one uint32 that the kernel adds to rnd_seed when it starts.  The
 *      reference passes the seed by value from a host counter (render/render.py:19,112-116); a host value is frozen into a captured
 *      CUDA graph, a device value is not -- the training step can be captured once and replayed with an advancing seed. */
int mcs_env_shade_fwd(mcs_ctx *ctx,
                      const mcs_tensor *mask, const mcs_tensor *ro, const mcs_tensor *gb_pos, const mcs_tensor *gb_normal,
                      const mcs_tensor *gb_view_pos, const mcs_tensor *gb_kd, const mcs_tensor *gb_ks,
                      const mcs_tensor *light, const mcs_tensor *pdf, const mcs_tensor *rows, const mcs_tensor *cols,
                      const mcs_tensor *perms,
                      uint32_t bsdf, uint32_t n_samples_x, uint32_t rnd_seed, const uint32_t *seed_offset_dev, float shadow_scale, int32_t batch_offset,
                      float *diff, float *spec, uint32_t *hit_record, uint32_t *rec_count, float *rec_rays, int32_t rec_slots, mcs_stream stream);

/* Gradient outputs: gb_pos_grad, gb_normal_grad, gb_kd_grad, gb_ks_grad contiguous [B,H,W,3]
 * (fully written), light_grad contiguous [Hl,Wl,3] (zeroed by the call, then accumulated). */
int mcs_env_shade_bwd(mcs_ctx *ctx,
                      const mcs_tensor *mask, const mcs_tensor *ro, const mcs_tensor *gb_pos, const mcs_tensor *gb_normal,
                      const mcs_tensor *gb_view_pos, const mcs_tensor *gb_kd, const mcs_tensor *gb_ks,
                      const mcs_tensor *light, const mcs_tensor *pdf, const mcs_tensor *rows, const mcs_tensor *cols,
                      const mcs_tensor *perms,
                      uint32_t bsdf, uint32_t n_samples_x, uint32_t rnd_seed, const uint32_t *seed_offset_dev, float shadow_scale, int32_t batch_offset,
                      const mcs_tensor *diff_grad, const mcs_tensor *spec_grad,
                      float *gb_pos_grad, float *gb_normal_grad, float *gb_kd_grad, float *gb_ks_grad, float *light_grad,
                      const uint32_t *hit_record /* NULL = re-trace */, mcs_stream stream);

/* Backward from the ray record written by mcs_env_shade_fwd (same G-buffer, light and shadow_scale): one warp per pixel walks
 * the recorded rays and runs only the adjoint BSDF + env-map gradient scatter.  Outputs as mcs_env_shade_bwd. */
int mcs_env_shade_bwd_replay(const mcs_tensor *gb_pos, const mcs_tensor *gb_normal, const mcs_tensor *gb_view_pos, const mcs_tensor *gb_kd,
                             const mcs_tensor *gb_ks, const mcs_tensor *light, uint32_t bsdf, uint32_t n_samples_x, float shadow_scale,
                             const mcs_tensor *diff_grad, const mcs_tensor *spec_grad, const uint32_t *rec_count, const float *rec_rays, int32_t rec_slots,
                             float *gb_pos_grad, float *gb_normal_grad, float *gb_kd_grad, float *gb_ks_grad, float *light_grad, mcs_stream stream);

/* Debug/parity hook: forward pass that also records, per pixel and per ray slot
 * (slot = 2*i for the light sample of stratum i, 2*i+1 for the BSDF sample), the env texel read
 * ((y<<16)|x) and the shadow-ray result (1 visible, 0 occluded; 255/-1 for masked pixels).
 * rec_texel int32 [B,H,W,2N^2], rec_vis uint8 [B,H,W,2N^2].  Rays whose contribution is provably
 * zero are not traced by the product; their rec_vis is 2 ("skipped"). */
int mcs_env_shade_records(mcs_ctx *ctx,
                          const mcs_tensor *mask, const mcs_tensor *ro, const mcs_tensor *gb_pos, const mcs_tensor *gb_normal,
                          const mcs_tensor *gb_view_pos, const mcs_tensor *gb_kd, const mcs_tensor *gb_ks,
                          const mcs_tensor *light, const mcs_tensor *pdf, const mcs_tensor *rows, const mcs_tensor *cols,
                          const mcs_tensor *perms,
                          uint32_t bsdf, uint32_t n_samples_x, uint32_t rnd_seed, const uint32_t *seed_offset_dev, float shadow_scale, int32_t batch_offset,
                          float *diff, float *spec, int32_t *rec_texel, uint8_t *rec_vis, mcs_stream stream);

/* ---- bilateral denoiser: replaces bilateral_denoiser_fwd / _bwd,
 *      optixutils/c_src/torch_bindings.cpp:274-319 (denoising.cu:14-130).
 *      col, nrm [B,H,W,3], zdz [B,H,W,2] strided views; out contiguous [B,H,W,4]
 *      = (sum w*col, max(sum w, 1e-4)); col_grad contiguous [B,H,W,3]; out_grad [B,H,W,4] view. */
int mcs_bilateral_fwd(const mcs_tensor *col, const mcs_tensor *nrm, const mcs_tensor *zdz, float sigma, float *out, mcs_stream stream);
int mcs_bilateral_bwd(const mcs_tensor *nrm, const mcs_tensor *zdz, float sigma, const mcs_tensor *out_grad, float *col_grad, mcs_stream stream);
/* Fused fast path for the two calls render/render.py:120-121 makes with identical guides
 * (diffuse + specular): weights are computed once.  outA/outB as above. */
int mcs_bilateral_fwd2(const mcs_tensor *colA, const mcs_tensor *colB, const mcs_tensor *nrm, const mcs_tensor *zdz, float sigma,
                       float *outA, float *outB, mcs_stream stream);
int mcs_bilateral_bwd2(const mcs_tensor *nrm, const mcs_tensor *zdz, float sigma, const mcs_tensor *out_gradA, const mcs_tensor *out_gradB,
                       float *col_gradA, float *col_gradB, mcs_stream stream);

/* ---- renderutils elementwise ops: replace the *_fwd / *_bwd functions of renderutils_plugin,
 *      renderutils/c_src/torch_bindings.cpp:866-888.  Inputs are broadcastable NHWC views; the launch
 *      grid (N,H,W) is the max over inputs (update_grid, torch_bindings.cpp:87-101); outputs and
 *      gradients are contiguous full-grid fp32 (gradients of broadcast inputs are NOT reduced, as in
 *      the reference, tensor.h:61,75 -- the autograd wrapper sums them). */
int mcs_lambert_fwd(const mcs_tensor *nrm, const mcs_tensor *wi, float *out, mcs_stream s);                                   /* torch_bindings.cpp:255 */
int mcs_lambert_bwd(const mcs_tensor *nrm, const mcs_tensor *wi, const mcs_tensor *d_out, float *d_nrm, float *d_wi, mcs_stream s);
int mcs_frostbite_fwd(const mcs_tensor *nrm, const mcs_tensor *wi, const mcs_tensor *wo, const mcs_tensor *lin_rough, float *out, mcs_stream s);
int mcs_frostbite_bwd(const mcs_tensor *nrm, const mcs_tensor *wi, const mcs_tensor *wo, const mcs_tensor *lin_rough, const mcs_tensor *d_out,
                      float *d_nrm, float *d_wi, float *d_wo, float *d_lin_rough, mcs_stream s);
int mcs_fresnel_shlick_fwd(const mcs_tensor *f0, const mcs_tensor *f90, const mcs_tensor *cos_theta, float *out, mcs_stream s);
int mcs_fresnel_shlick_bwd(const mcs_tensor *f0, const mcs_tensor *f90, const mcs_tensor *cos_theta, const mcs_tensor *d_out,
                           float *d_f0, float *d_f90, float *d_cos, mcs_stream s);
int mcs_ndf_ggx_fwd(const mcs_tensor *alpha_sqr, const mcs_tensor *cos_theta, float *out, mcs_stream s);
int mcs_ndf_ggx_bwd(const mcs_tensor *alpha_sqr, const mcs_tensor *cos_theta, const mcs_tensor *d_out, float *d_alpha_sqr, float *d_cos, mcs_stream s);
int mcs_lambda_ggx_fwd(const mcs_tensor *alpha_sqr, const mcs_tensor *cos_theta, float *out, mcs_stream s);
int mcs_lambda_ggx_bwd(const mcs_tensor *alpha_sqr, const mcs_tensor *cos_theta, const mcs_tensor *d_out, float *d_alpha_sqr, float *d_cos, mcs_stream s);
int mcs_masking_smith_fwd(const mcs_tensor *alpha_sqr, const mcs_tensor *cos_i, const mcs_tensor *cos_o, float *out, mcs_stream s);
int mcs_masking_smith_bwd(const mcs_tensor *alpha_sqr, const mcs_tensor *cos_i, const mcs_tensor *cos_o, const mcs_tensor *d_out,
                          float *d_alpha_sqr, float *d_cos_i, float *d_cos_o, mcs_stream s);
int mcs_pbr_specular_fwd(const mcs_tensor *col, const mcs_tensor *nrm, const mcs_tensor *wo, const mcs_tensor *wi, const mcs_tensor *alpha,
                         float min_roughness, float *out, mcs_stream s);
int mcs_pbr_specular_bwd(const mcs_tensor *col, const mcs_tensor *nrm, const mcs_tensor *wo, const mcs_tensor *wi, const mcs_tensor *alpha,
                         float min_roughness, const mcs_tensor *d_out,
                         float *d_col, float *d_nrm, float *d_wo, float *d_wi, float *d_alpha, mcs_stream s);
/* pbr_bsdf_fwd / pbr_bsdf_bwd, renderutils/c_src/torch_bindings.cpp:653-722; bsdf: 0 lambert, 1 frostbite */
int mcs_pbr_bsdf_fwd(const mcs_tensor *kd, const mcs_tensor *arm, const mcs_tensor *pos, const mcs_tensor *nrm, const mcs_tensor *view_pos,
                     const mcs_tensor *light_pos, float min_roughness, int32_t bsdf, float *out, mcs_stream s);
int mcs_pbr_bsdf_bwd(const mcs_tensor *kd, const mcs_tensor *arm, const mcs_tensor *pos, const mcs_tensor *nrm, const mcs_tensor *view_pos,
                     const mcs_tensor *light_pos, float min_roughness, int32_t bsdf, const mcs_tensor *d_out,
                     float *d_kd, float *d_arm, float *d_pos, float *d_nrm, float *d_view_pos, float *d_light_pos, mcs_stream s);
/* prepare_shading_normal_fwd / _bwd, renderutils/c_src/torch_bindings.cpp:170-250 (normal.cu:95-178) */
int mcs_prepare_shading_normal_fwd(const mcs_tensor *pos, const mcs_tensor *view_pos, const mcs_tensor *perturbed_nrm, const mcs_tensor *smooth_nrm,
                                   const mcs_tensor *smooth_tng, const mcs_tensor *geom_nrm, int32_t two_sided_shading, int32_t opengl,
                                   float *out, mcs_stream s);
int mcs_prepare_shading_normal_bwd(const mcs_tensor *pos, const mcs_tensor *view_pos, const mcs_tensor *perturbed_nrm, const mcs_tensor *smooth_nrm,
                                   const mcs_tensor *smooth_tng, const mcs_tensor *geom_nrm, int32_t two_sided_shading, int32_t opengl,
                                   const mcs_tensor *d_out,
                                   float *d_pos, float *d_view_pos, float *d_perturbed_nrm, float *d_smooth_nrm, float *d_smooth_tng, float *d_geom_nrm,
                                   mcs_stream s);

/* ---- image loss: replaces image_loss_fwd / image_loss_bwd, renderutils/c_src/torch_bindings.cpp:739-800
 *      (loss.cu:105-227).  loss: 0 l1, 1 mse, 2 relmse, 3 smape, 4 n2n (FIX: the reference's strToLoss maps "n2n" to l1);
 *      tonemapper: 0 none, 1 log_srgb.  Forward writes mcs_image_loss_num_partials(N,H,W) deterministic per-CTA partial sums of
 *      mean_c(loss) -- the caller sums them and divides by N*H*W exactly like renderutils/ops.py:494.  Backward takes the upstream
 *      gradient of those partials ([P,1,1,1] view, or one broadcast value) and writes contiguous [N,H,W,3] gradients. */
int mcs_image_loss_num_partials(int32_t N, int32_t H, int32_t W);
int mcs_image_loss_fwd(const mcs_tensor *img, const mcs_tensor *target, int32_t loss, int32_t tonemapper, float *partials, mcs_stream s);
int mcs_image_loss_bwd(const mcs_tensor *img, const mcs_tensor *target, int32_t loss, int32_t tonemapper, const mcs_tensor *d_partials,
                       float *d_img, float *d_target, mcs_stream s);

/* ---- batched 4x4 transform (row f3): replaces xfm_fwd / xfm_bwd, renderutils/c_src/torch_bindings.cpp:803-864 (mesh.cu:19-90).
 *      points as a (1|B, V, 3, 1) view, matrix as (B, 4, 4, 1); out contiguous [B,V,4] (is_points) or [B,V,3] (vectors);
 *      d_out (B, V, 4|3, 1) view; d_points contiguous [B,V,3] (a broadcast input's gradient is reduced by the caller). */
int mcs_xfm_fwd(const mcs_tensor *points, const mcs_tensor *matrix, int32_t is_points, float *out, mcs_stream s);
int mcs_xfm_bwd(const mcs_tensor *points, const mcs_tensor *matrix, const mcs_tensor *d_out, int32_t is_points, float *d_points, mcs_stream s);

/* ---- tail of render.shade() (row f3): normalise the denoiser outputs and recombine the demodulated signals, render/render.py:119-131.
 *      a4 / b4: [B,H,W,4] raw bilateral outputs (rgb weighted sum, weight) of the diffuse / specular signal; kd, ks [B,H,W,3];
 *      pbr != 0: out = a.rgb/a.w * kd * (1 - ks.z) + b.rgb/b.w ;  pbr == 0 ('diffuse' / 'white'): out = a.rgb/a.w * kd (b4, ks ignored but
 *      must be valid views).  Backward writes contiguous gradients for a4, kd (and b4, ks when pbr). */
int mcs_shade_combine_fwd(const mcs_tensor *a4, const mcs_tensor *b4, const mcs_tensor *kd, const mcs_tensor *ks, int32_t pbr, float *out, mcs_stream s);
int mcs_shade_combine_bwd(const mcs_tensor *a4, const mcs_tensor *b4, const mcs_tensor *kd, const mcs_tensor *ks, int32_t pbr, const mcs_tensor *d_out,
                          float *d_a4, float *d_b4, float *d_kd, float *d_ks, mcs_stream s);

/* ---- primary visibility + attribute interpolation: stands in for dr.rasterize / dr.interpolate of nvdiffrast at
 *      the call sites render/render.py:208-234 (closest hit on the context's LBVH instead of rasterisation).
 *      mtx: [B,4,4] row-major fp32 device array, clip = mtx * (p, 1) (inverted on the device, no host round trip).  rast: [B,H,W,4] contiguous, nvdiffrast
 *      convention (u, v, z/w, triangle_id + 1) with u / v the barycentric weights of vertex 0 / 1; all zero = background.
 *      interpolate: attr [V,C] (attr_batch_stride 0) or [B,V,C] (stride V*C), tris int32 [T,3], out / d_out [B,H,W,C];
 *      d_attr (same layout as attr) must be zeroed by the caller and receives float atomics. */
int mcs_rasterize(mcs_ctx *ctx, const float *mtx, int32_t B, int32_t H, int32_t W, float *rast, mcs_stream stream);
int mcs_interpolate_fwd(const float *attr, int64_t attr_batch_stride, int32_t V, int32_t C, const int32_t *tris, int32_t T, const float *rast,
                        int32_t B, int32_t H, int32_t W, float *out, mcs_stream stream);
int mcs_interpolate_bwd(const float *attr, int64_t attr_batch_stride, int32_t V, int32_t C, const int32_t *tris, int32_t T, const float *rast,
                        int32_t B, int32_t H, int32_t W, const float *d_out, float *d_attr, mcs_stream stream);

/* Nearest-texel fetch out[i,:] = tex[idx[i],:] (tex [T,C] contiguous, idx int64 [n]; out-of-range indices give zeros) and its scatter-add
 * backward into a caller-zeroed d_tex [T,C] (float atomics) -- the material look-up of the synthetic G-buffer producer. */
int mcs_texel_fetch_fwd(const float *tex, int64_t T, int32_t C, const int64_t *idx, int64_t n, float *out, mcs_stream stream);
int mcs_texel_fetch_bwd(int64_t T, int32_t C, const int64_t *idx, int64_t n, const float *d_out, float *d_tex, mcs_stream stream);

/* ---- env-light pdf / CDF tables: replaces the torch-op chain of EnvironmentLight.update_pdf,
 *      render/light.py:46-59.  base: (1, Hl, Wl, 3) view.  Outputs (caller-allocated, contiguous): pdf [Hl,Wl] normalised to sum 1,
 *      rows [Hl] (what the call site passes as lgt.rows[:,0], render/render.py:114), cols [Hl,Wl]; row_totals: Hl doubles of scratch
 *      (holds the un-normalised row sums on return).  Sums are carried in fp64 and rounded once. */
int mcs_update_pdf(const mcs_tensor *base, float *pdf, float *rows, float *cols, double *row_totals, mcs_stream s);

#ifdef __cplusplus
}
#endif
#endif /* MCSHADE_H */
