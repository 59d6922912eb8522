/*
 * sunerf_b200.h - C ABI of the B200-native SuNeRF ray-render hot path.
 *
 * The reference (FrontierDevelopmentLab/2024-HL-SPI3S-SuNeRF) has no FFI: its boundary is a Python class
 * surface.  This header is what a ctypes binding on the reference side would load;
 * each entry point cites the reference code it replaces (paths relative to the upstream tree).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch), unless it says "host";
 *   - every entry is stream-ordered on `stream` (a cudaStream_t) of the CURRENT device, never synchronises, keeps no
 *     pointer after return, allocates nothing and is safe to call from several host threads and on several devices
 *     of one process (per-device one-time setup is internal and thread-safe; there is no other mutable state);
 *   - return value: 0 ok, <0 bad argument (SNF_E_*), >0 a cudaError_t;
 *   - float32 everywhere unless a name says bf16; tensors are contiguous row-major.
 */
#ifndef SUNERF_B200_H
#define SUNERF_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SNF_VERSION 100
#define SNF_E_ARG (-1)      /* null pointer / non-positive size */
#define SNF_E_SHAPE (-2)    /* size outside what the kernels are instantiated for */
#define SNF_E_ALIGN (-3)    /* pointer not aligned as documented */
#define SNF_N_AIA 7         /* AIA channels 94,131,171,193,211,304,335 (model.py:154-162) */
#define SNF_TABLE_LEN 101   /* logT grid of aia_temp_resp.genx */

int snf_version(void);
const char *snf_error_string(int code);
/* number of kernel launches issued through this library by the calling process (bench.py gpu_launches) */
int64_t snf_launch_count(void);
/* a CUDA-graph replay launches the captured kernels without passing through this library: the host adds them here */
void snf_count_launches(int64_t n);

/* Configuration of the CURRENT device (thread-safe, takes effect for later launches): the persistent field-network grids
 * leave `n` SMs (rounded down to even, 0..64) free for a concurrent collective.  RayTrainer sets it when world_size > 1:
 * it replaces nothing in the reference (Lightning 'dp' has no such notion), it is what lets the gradient exchange of
 * run_emission.py:64-69 overlap the backward instead of queueing behind it.  Returns 0 or SNF_E_ARG. */
int snf_config_reserve_sms(int n);

/* ---- a1: StratifiedSampler.forward, sunerf/train/sampling.py:68-102 -------------------------------
 * t_vals[S] is the sampler buffer (linspace(0,1,S)); t_rand[N,S] is the torch.rand draw of :97, or NULL
 * for perturb=False.  Writes z_vals[N,S] and, if non-NULL, points[N,S,3].  Same fp32 operation order as
 * the reference (no FMA contraction). */
int snf_stratified_sample(const float *rays_o, const float *rays_d, const float *t_vals, const float *t_rand,
                          int64_t N, int S, float distance, float solar_R, float *z_vals, float *points,
                          void *stream);

/* SphericalSampler.forward, sunerf/train/sampling.py:16-54 (selectable through sampling_config {'type': 'spherical'},
 * base_tracing.py:27-28): bins between the ray's entry and exit of the sphere of radius `distance`, the far end clipped
 * at the solar surface; NaN rows for rays that miss the sphere, as in the reference.  Same arguments as above. */
int snf_spherical_sample(const float *rays_o, const float *rays_d, const float *t_vals, const float *t_rand,
                         int64_t N, int S, float distance, float solar_R, float *z_vals, float *points,
                         void *stream);

/* ---- a2: HierarchicalSampler.forward + sample_pdf, sampling.py:111-169 (perturb=False) --------------
 * z_vals[N,S], weights[N,S] (coarse weights), u[n_new] = linspace(0,1,n_new).
 * cdf_in[N,S-1]: optional externally supplied CDF (stage-boundary parity test); NULL -> built here.
 * Outputs: new_z[N,n_new], z_comb[N,S+n_new] (sorted merge); optional inds[N,n_new] (int64, the
 * searchsorted(right=True) result) and cdf_out[N,S-1].  S<=256, n_new<=512. */
int snf_hier_resample(const float *z_vals, const float *weights, const float *u, const float *cdf_in, int64_t N,
                      int S, int n_new, float *new_z, float *z_comb, int64_t *inds, float *cdf_out, void *stream);
/* HierarchicalSampler(perturb=True), sampling.py:144-146: u_rand[N,n_new] is the torch.rand draw of :145 (one row of
 * unordered uniforms per ray, drawn by the host exactly as the reference does); everything else as above. */
int snf_hier_resample_perturb(const float *z_vals, const float *weights, const float *u_rand, const float *cdf_in,
                              int64_t N, int S, int n_new, float *new_z, float *z_comb, int64_t *inds, float *cdf_out,
                              void *stream);

/* ---- a3: query = cat(o + d*z, t), sunerf/rendering/base_tracing.py:64-65, 83-84; sampling.py:100 ---- */
/* N1: observer image rays on the device.  Replaces get_rays (sunerf/data/ray_sampling.py:7-36) for the
 * regular pixel grid Tx = (j - cx) * plate_arcsec * asec, Ty = (i - cy) * plate_arcsec * asec; pixels [first, first+count)
 * of the row-major H x W image.  c2w_host: the pose_spherical matrix (coordinate_transformation.py:36-54), HOST memory,
 * row-major with a row stride of 4 floats (3x4 or 4x4). */
int snf_image_rays(const float *c2w_host, int H, int W, double plate_arcsec, double asec, double cx, double cy,
                   int64_t first, int64_t count, float *rays_o, float *rays_d, void *stream);

int snf_make_query(const float *rays_o, const float *rays_d, const float *z, const float *times, int64_t N, int S,
                   float *query /*[N,S,4]*/, void *stream);

/* ---- a4-a6: PositionalEncoding + NeRF / NeRF_DT forward, sunerf/model/model.py:44-57,123-132,169-187 -
 * Network: enc(4->84) -> n_hidden x [Linear(.,d_filter)+sin] -> Linear(d_filter,2) (+ out_offset).
 * W, B: HOST arrays of n_hidden+1 device pointers in nn.Linear layout (weight[out,in]).
 * fp32 entry points: FFMA SIMT, the 1e-5 parity mode.  ws: workspace of snf_mlp_ws_bytes() bytes.
 * train!=0 keeps every layer's activation and cosine in ws for snf_mlp_bwd_f32. */
int64_t snf_mlp_ws_bytes(int64_t M, int n_hidden, int d_filter, int mode /*0 fp32 SIMT, 1 16-bit tensor cores, 2 split-precision tensor cores*/, int train);
int snf_mlp_fwd_f32(const float *x /*[M,4]*/, int64_t M, const float *const *W, const float *const *B,
                    int n_hidden, int d_filter, float out_offset0, float out_offset1, float *out /*[M,2]*/,
                    void *ws, int train, void *stream);
/* gW/gB: HOST arrays of device pointers, same shapes as W/B, OVERWRITTEN with the gradients. */
int snf_mlp_bwd_f32(const float *x, int64_t M, const float *const *W, int n_hidden, int d_filter,
                    const float *grad_out /*[M,2]*/, void *ws, float *const *gW, float *const *gB, void *stream);

/* 16-bit tensor-core entry points ("bf16-MLP mode" of the task; tcgen05 + TMEM, weights streamed by TMA bulk copies).
 * Operands are fp16 since round 2 (weights, activations, scaled gradients; fp32 accumulation): the 1e-3 per-parameter
 * gradient gate needs 11-bit significands on the weights.  Raw coordinates must be finite and
 * below 65504 in magnitude.  d_filter==512, n_hidden==8 only.  `packed`: snf_mlp_pack_bytes() bytes, 1024-byte aligned,
 * holding the UMMA-layout fp16 copy of the weights; call snf_mlp_pack_bf16 whenever the fp32 weights changed.
 * ws (train): snf_mlp_ws_bytes(M, 8, 512, 1, 1) bytes, 1024-byte aligned. */
int64_t snf_mlp_pack_bytes(void);
int snf_mlp_pack_bf16(const float *const *W, const float *const *B, void *packed, void *stream);
int snf_mlp_fwd_bf16(const float *x, int64_t M, const void *packed, float out_offset0, float out_offset1,
                     float *out, void *ws, int train, void *stream);
int snf_mlp_bwd_bf16(const float *x, int64_t M, const void *packed, const float *grad_out, void *ws,
                     float *const *gW, float *const *gB, void *stream);

/* Split-precision tensor-core entry points: the 1e-5 "fp32 mode" of the default 8 x 512 network on tcgen05.  Every
 * operand of the forward is the pair (fp16(v), fp16(v - fp16(v))) and every product three MMAs into one fp32 accumulator
 * (csrc/snf_mlp_x3.cu); the backward is the 16-bit one with the W^T operand of the dgrad chain split in two.  Same
 * `packed` buffer as above (snf_mlp_pack_bf16 also writes the low halves); ws: snf_mlp_ws_bytes(M, 8, 512, 2, train)
 * bytes, 1024-byte aligned.  Intensities ~3e-7, per-parameter gradients ~1e-4 of the reference's. */
int snf_mlp_fwd_x3(const float *x, int64_t M, const void *packed, float out_offset0, float out_offset1,
                   float *out, void *ws, int train, void *stream);
int snf_mlp_bwd_x3(const float *x, int64_t M, const void *packed, const float *grad_out, void *ws,
                   float *const *gW, float *const *gB, void *stream);

/* ---- a7: SimpleStar.forward, sunerf/model/stellar_model.py:53-102 --------------------------------- */
int snf_simple_star_fwd(const float *x /*[M,4]*/, int64_t M, float rho_0, float h0, float T0, float R_s,
                        float t_photosphere, float *out /*[M,2]*/, void *stream);

/* ---- a8: EmissionRadiativeTransfer.raw2outputs, sunerf/rendering/emission.py:14-54 ----------------
 * + cumprod_exclusive, base_tracing.py:135-156.  image[N], weights[N,S], absorption[N,S].  S<=256. */
int snf_composite_emission_fwd(const float *raw /*[N,S,2]*/, const float *z /*[N,S]*/, const float *rays_d,
                               int64_t N, int S, float *image, float *weights, float *absorption, void *stream);
/* analytic backward: g_image[N], g_absorption[N,S] (or NULL) -> g_raw[N,S,2] (overwritten) */
int snf_composite_emission_bwd(const float *raw, const float *z, const float *rays_d, int64_t N, int S,
                               const float *g_image, const float *g_absorption, float *g_raw, void *stream);

/* ---- a9: DensityTemperatureRadiativeTransfer.raw2outputs, rendering/density_temperature.py:192-271 -
 * wavelengths[N,C] float (0 = channel absent); log_abs[7], vol_c[1], table_x[101], table_y[7,101] device.
 * image[N,C], weights[N,S], regq[N,S].  S<=256, C<=8. */
int snf_composite_dt_fwd(const float *inferences /*[N,S,2]*/, const float *z, const float *wavelengths, int64_t N,
                         int S, int C, const float *log_abs, const float *vol_c, const float *table_x,
                         const float *table_y, float pixel_intensity_factor, float *image, float *weights,
                         float *regq, void *stream);
/* g_image[N,C], g_regq[N,S] (or NULL) -> g_inferences[N,S,2] (overwritten); g_log_abs[7] and g_vol_c[1]
 * are ACCUMULATED into (caller zeroes them). */
int snf_composite_dt_bwd(const float *inferences, const float *z, const float *wavelengths, int64_t N, int S, int C,
                         const float *log_abs, const float *vol_c, const float *table_x, const float *table_y,
                         float pixel_intensity_factor, const float *g_image, const float *g_regq,
                         float *g_inferences, float *g_log_abs, float *g_vol_c, void *stream);

/* ---- a10: SuNeRFRendering.forward epilogue, base_tracing.py:91-111; regularization :43-44 and
 * density_temperature.py:273-274.  kind 0 = emission: reg = relu(dist-r0)*(1-q); 1 = DT: relu(dist-r0)*relu(q).
 * g_q (optional, [N,S]): d(lambda*mean(reg))/dq for the backward of the fine compositing. */
int snf_render_epilogue(const float *rays_o, const float *rays_d, const float *z_comb, const float *weights,
                        const float *q, int64_t N, int S, float r0, int kind, float *height_map,
                        float *absorption_map, float *reg, float reg_grad_scale, float *g_q, void *stream);

/* ---- a11: training_step losses, sunerf/model/sunerf.py:98-131 (asinh-MSE, train/scaling.py:17-28) and
 * :173-206 (plain MSE).  images [N,C]; losses[4] = {total, coarse, fine, reg}; g_coarse/g_fine [N,C];
 * finite_flag[1] is OR-ed with 1 when any of the inputs is NaN/Inf (the assert of :105-107). */
int snf_train_loss(const float *coarse, const float *fine, const float *target, const float *reg, int64_t N, int C,
                   int64_t n_reg, int asinh_scaling, float asinh_a, float lambda_image, float lambda_reg,
                   float *losses, float *g_coarse, float *g_fine, int *finite_flag, void *stream);

/* ---- a12: clip-by-global-norm (run_emission.py:72) + Adam (sunerf.py:30-35) on one flat buffer ------
 * grads are first scaled by grad_scale (1/world_size after the allreduce).  scratch: >= 1024 floats.
 * step is 1-based.  norm_out[1] receives the pre-clip global norm. */
int snf_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int64_t step, float clip_norm, float grad_scale,
                  float *scratch, float *norm_out, void *stream);

/* Same step with the schedule resident on the device (CUDA-graph replay: no host scalar changes between launches).
 * sched = {lr, step (starts at 1), gamma, lr_floor} as doubles; after the update the step is advanced and lr is multiplied
 * by gamma while lr > lr_floor - the per-batch ExponentialLR rule of sunerf/model/sunerf.py:36-40. */
int snf_adam_step_sched(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, double *sched,
                        float beta1, float beta2, float eps, float clip_norm, float grad_scale, float *scratch,
                        float *norm_out, void *stream);

/* ---- whole chain: SuNeRFRendering.forward, sunerf/rendering/base_tracing.py:46-111, as ONE call ------------------
 * (stratified sampling -> query -> coarse field -> compositing -> hierarchical resampling -> query -> fine field ->
 * compositing -> epilogue), and its backward into the parameter gradients.  This is the entry a non-Python host binds;
 * the Python classes launch the same kernels stage by stage.  Everything the descriptor points to is DEVICE memory except
 * the W_x / B_x arrays themselves (HOST arrays of n_hidden+1 device pointers, as in snf_mlp_fwd_f32). */
typedef struct snf_render_desc {
  int kind;                 /* 0: EmissionRadiativeTransfer (emission.py:14-54), 1: DensityTemperatureRadiativeTransfer */
  int mode;                 /* 0: fp32 SIMT field networks (W_x / B_x); 1: 16-bit tensor cores, 2: split-precision tensor cores
                             * (both: packed_x from snf_mlp_pack_bf16) */
  int n_hidden, d_filter;   /* 8, 512 */
  const float *const *W_coarse, *const *B_coarse, *const *W_fine, *const *B_fine;
  const void *packed_coarse, *packed_fine;
  float out_offset0, out_offset1;   /* NeRF_DT base_log_density / base_log_temperature (model.py:180-181), else 0 */
  const float *t_vals;      /* [S]     StratifiedSampler buffer linspace(0,1,S), sampling.py:64 */
  const float *u;           /* [n_new] HierarchicalSampler linspace(0,1,n_new), sampling.py:141 */
  int S, n_new;             /* 64, 128; S + n_new <= 256 */
  float distance, solar_R;  /* sampling.py:62-63 */
  float reg_radius;         /* 1.2 / Rs_per_ds (base_tracing.py:43) or 1.25 / Rs_per_ds (density_temperature.py:273) */
  int C;                    /* kind 1: wavelength channels per ray (<= 8) */
  float pixel_intensity_factor;
  const float *log_abs_coarse, *vol_c_coarse, *log_abs_fine, *vol_c_fine;   /* kind 1: [7], [1] per model */
  const float *table_x, *table_y;                                           /* kind 1: [101], [7,101] */
} snf_render_desc;

int64_t snf_render_ws_bytes(const snf_render_desc *desc, int64_t N, int train);
/* times[N], wavelengths[N,C] (kind 1, else NULL), t_rand[N,S] or NULL (perturb=False).  ws: 1024-byte aligned workspace.
 * Outputs are the keys of the reference's output dict: z_vals_stratified[N,S] (may be NULL), coarse_image[N,C],
 * z_vals_hierarchical[N,n_new], fine_image[N,C], height_map[N], absorption_map[N], regularization[N,S+n_new]
 * (C = 1 for kind 0).  train != 0 keeps in ws what snf_render_fused_bwd needs, including
 * d(reg_grad_scale * sum(regularization)) / d(regularizing_quantity) of the fine pass. */
int snf_render_fused_fwd(const snf_render_desc *desc, const float *rays_o, const float *rays_d, const float *times,
                         const float *wavelengths, const float *t_rand, int64_t N, void *ws, int train,
                         float reg_grad_scale, float *z_vals_stratified, float *coarse_image,
                         float *z_vals_hierarchical, float *fine_image, float *height_map, float *absorption_map,
                         float *regularization, void *stream);
/* Backward of the call above (same desc, N and ws): g_coarse_image / g_fine_image [N,C] are dL/d(image) of the two
 * passes; with_reg_grad != 0 adds the regulariser term stored by the forward.  gW_x / gB_x: HOST arrays of device
 * pointers, OVERWRITTEN with the gradients; g_log_abs_x[7] / g_vol_c_x[1] (kind 1) are ACCUMULATED into. */
int snf_render_fused_bwd(const snf_render_desc *desc, const float *rays_d, const float *wavelengths, int64_t N, void *ws,
                         const float *g_coarse_image, const float *g_fine_image, int with_reg_grad,
                         float *const *gW_coarse, float *const *gB_coarse, float *const *gW_fine, float *const *gB_fine,
                         float *g_log_abs_coarse, float *g_vol_c_coarse, float *g_log_abs_fine, float *g_vol_c_fine,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif
