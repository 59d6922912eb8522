// Whole-chain entry points: SuNeRFRendering.forward (sunerf/rendering/base_tracing.py:46-111) for one coarse/fine model
// pair as ONE C call, and the matching backward into the parameter gradients.  Host-side orchestration only: every
// step is one of the kernels behind the single-stage entry points of this library, launched in the reference's order on
// the caller's stream (no allocation, no synchronisation, graph-capturable).  Intermediates live in a caller-owned
// workspace of snf_render_ws_bytes() bytes; a training forward leaves there what snf_render_fused_bwd needs.
#include "snf_common.cuh"

namespace snf {
namespace {

struct RenderWs {
  float *z, *z_comb, *query_c, *query_f, *raw_c, *raw_f, *w_c, *w_f, *q_c, *q_f, *g_q, *g_raw_c, *g_raw_f;
  void *mlp_c, *mlp_f;
  int64_t bytes;
};

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// returns <0 on a bad descriptor
int64_t render_layout(const snf_render_desc *d, int64_t N, int train, void *base, RenderWs *out) {
  const int S = d->S, Sf = d->S + d->n_new;
  const int64_t mc = snf_mlp_ws_bytes(N * S, d->n_hidden, d->d_filter, d->mode, train);
  const int64_t mf = snf_mlp_ws_bytes(N * Sf, d->n_hidden, d->d_filter, d->mode, train);
  if (mc < 0 || mf < 0) return SNF_E_ARG;
  uint8_t *p = reinterpret_cast<uint8_t *>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { uint8_t *r = p ? p + off : nullptr; off += align_up(bytes, 1024); return r; };
  RenderWs w{};
  w.mlp_c = take(mc);
  w.mlp_f = take(mf);
  w.z = reinterpret_cast<float *>(take(N * S * 4));
  w.z_comb = reinterpret_cast<float *>(take(N * Sf * 4));
  w.query_c = reinterpret_cast<float *>(take(N * S * 16));
  w.query_f = reinterpret_cast<float *>(take(N * Sf * 16));
  w.raw_c = reinterpret_cast<float *>(take(N * S * 8));
  w.raw_f = reinterpret_cast<float *>(take(N * Sf * 8));
  w.w_c = reinterpret_cast<float *>(take(N * S * 4));
  w.w_f = reinterpret_cast<float *>(take(N * Sf * 4));
  w.q_c = reinterpret_cast<float *>(take(N * S * 4));
  w.q_f = reinterpret_cast<float *>(take(N * Sf * 4));
  if (train) {
    w.g_q = reinterpret_cast<float *>(take(N * Sf * 4));
    w.g_raw_c = reinterpret_cast<float *>(take(N * S * 8));
    w.g_raw_f = reinterpret_cast<float *>(take(N * Sf * 8));
  }
  w.bytes = off > 0 ? off : 1024;
  if (out) *out = w;
  return w.bytes;
}

int check_desc(const snf_render_desc *d) {
  SNF_CHECK_PTR(d);
  if (d->kind != 0 && d->kind != 1) return SNF_E_ARG;
  if (d->mode < 0 || d->mode > 2) return SNF_E_ARG;
  if (d->S < 3 || d->n_new <= 0 || d->S > 256 || d->S + d->n_new > 256) return SNF_E_SHAPE;
  SNF_CHECK_PTR(d->t_vals); SNF_CHECK_PTR(d->u);
  if (d->mode == 0) { SNF_CHECK_PTR(d->W_coarse); SNF_CHECK_PTR(d->B_coarse); SNF_CHECK_PTR(d->W_fine); SNF_CHECK_PTR(d->B_fine); }
  else { SNF_CHECK_PTR(d->packed_coarse); SNF_CHECK_PTR(d->packed_fine); }
  if (d->kind == 1) {
    if (d->C <= 0 || d->C > 8) return SNF_E_SHAPE;
    SNF_CHECK_PTR(d->log_abs_coarse); SNF_CHECK_PTR(d->vol_c_coarse); SNF_CHECK_PTR(d->log_abs_fine);
    SNF_CHECK_PTR(d->vol_c_fine); SNF_CHECK_PTR(d->table_x); SNF_CHECK_PTR(d->table_y);
  }
  return 0;
}

int field_fwd(const snf_render_desc *d, bool fine, const float *query, int64_t M, float *raw, void *ws, int train, void *st) {
  if (d->mode == 1)
    return snf_mlp_fwd_bf16(query, M, fine ? d->packed_fine : d->packed_coarse, d->out_offset0, d->out_offset1, raw, ws, train, st);
  if (d->mode == 2)
    return snf_mlp_fwd_x3(query, M, fine ? d->packed_fine : d->packed_coarse, d->out_offset0, d->out_offset1, raw, ws, train, st);
  return snf_mlp_fwd_f32(query, M, fine ? d->W_fine : d->W_coarse, fine ? d->B_fine : d->B_coarse, d->n_hidden, d->d_filter,
                         d->out_offset0, d->out_offset1, raw, ws, train, st);
}

int composite_fwd(const snf_render_desc *d, bool fine, const float *raw, const float *z, const float *rays_d,
                  const float *wavelengths, int64_t N, int S, float *image, float *weights, float *q, void *st) {
  if (d->kind == 0) return snf_composite_emission_fwd(raw, z, rays_d, N, S, image, weights, q, st);
  return snf_composite_dt_fwd(raw, z, wavelengths, N, S, d->C, fine ? d->log_abs_fine : d->log_abs_coarse,
                              fine ? d->vol_c_fine : d->vol_c_coarse, d->table_x, d->table_y, d->pixel_intensity_factor,
                              image, weights, q, st);
}

}  // namespace
}  // namespace snf

using namespace snf;

extern "C" int64_t snf_render_ws_bytes(const snf_render_desc *d, int64_t N, int train) {
  if (int e = check_desc(d)) return e;
  if (N < 0) return SNF_E_ARG;
  return render_layout(d, N, train, nullptr, nullptr);
}

#define SNF_TRY(call) \
  do { if (int e_ = (call)) return e_; } while (0)

extern "C" int snf_render_fused_fwd(const snf_render_desc *d, const float *rays_o, const float *rays_d, const float *times,
                                    const float *wavelengths, const float *t_rand, int64_t N, void *ws, int train,
                                    float reg_grad_scale, float *z_vals_stratified, float *coarse_image,
                                    float *z_vals_hierarchical, float *fine_image, float *height_map,
                                    float *absorption_map, float *regularization, void *stream) {
  if (int e = check_desc(d)) return e;
  if (N == 0) return 0;
  if (N < 0) return SNF_E_ARG;
  SNF_CHECK_PTR(rays_o); SNF_CHECK_PTR(rays_d); SNF_CHECK_PTR(times); SNF_CHECK_PTR(ws);
  SNF_CHECK_PTR(coarse_image); SNF_CHECK_PTR(fine_image); SNF_CHECK_PTR(z_vals_hierarchical);
  SNF_CHECK_PTR(height_map); SNF_CHECK_PTR(absorption_map); SNF_CHECK_PTR(regularization);
  SNF_CHECK_ALIGN(ws, 1024);
  if (d->kind == 1) SNF_CHECK_PTR(wavelengths);
  RenderWs w;
  if (render_layout(d, N, train, ws, &w) < 0) return SNF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = d->S, Sf = d->S + d->n_new;
  // coarse pass (base_tracing.py:53-70)
  SNF_TRY(snf_stratified_sample(rays_o, rays_d, d->t_vals, t_rand, N, S, d->distance, d->solar_R, w.z, nullptr, stream));
  SNF_TRY(snf_make_query(rays_o, rays_d, w.z, times, N, S, w.query_c, stream));
  SNF_TRY(field_fwd(d, false, w.query_c, N * S, w.raw_c, w.mlp_c, train, stream));
  SNF_TRY(composite_fwd(d, false, w.raw_c, w.z, rays_d, wavelengths, N, S, coarse_image, w.w_c, w.q_c, stream));
  // fine pass (:72-89)
  SNF_TRY(snf_hier_resample(w.z, w.w_c, d->u, nullptr, N, S, d->n_new, z_vals_hierarchical, w.z_comb, nullptr, nullptr, stream));
  SNF_TRY(snf_make_query(rays_o, rays_d, w.z_comb, times, N, Sf, w.query_f, stream));
  SNF_TRY(field_fwd(d, true, w.query_f, N * Sf, w.raw_f, w.mlp_f, train, stream));
  SNF_TRY(composite_fwd(d, true, w.raw_f, w.z_comb, rays_d, wavelengths, N, Sf, fine_image, w.w_f, w.q_f, stream));
  // epilogue (:91-111)
  SNF_TRY(snf_render_epilogue(rays_o, rays_d, w.z_comb, w.w_f, w.q_f, N, Sf, d->reg_radius, d->kind, height_map, absorption_map,
                              regularization, reg_grad_scale, train ? w.g_q : nullptr, stream));
  if (z_vals_stratified != nullptr) {
    cudaError_t e = cudaMemcpyAsync(z_vals_stratified, w.z, (size_t)N * S * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
  }
  return launch_status();
}

extern "C" int snf_render_fused_bwd(const snf_render_desc *d, const float *rays_d, const float *wavelengths, int64_t N,
                                    void *ws, const float *g_coarse_image, const float *g_fine_image, int with_reg_grad,
                                    float *const *gW_coarse, float *const *gB_coarse, float *const *gW_fine,
                                    float *const *gB_fine, float *g_log_abs_coarse, float *g_vol_c_coarse,
                                    float *g_log_abs_fine, float *g_vol_c_fine, void *stream) {
  if (int e = check_desc(d)) return e;
  if (N == 0) return 0;
  if (N < 0) return SNF_E_ARG;
  SNF_CHECK_PTR(rays_d); SNF_CHECK_PTR(ws); SNF_CHECK_PTR(g_coarse_image); SNF_CHECK_PTR(g_fine_image);
  SNF_CHECK_PTR(gW_coarse); SNF_CHECK_PTR(gB_coarse); SNF_CHECK_PTR(gW_fine); SNF_CHECK_PTR(gB_fine);
  SNF_CHECK_ALIGN(ws, 1024);
  if (d->kind == 1) {
    SNF_CHECK_PTR(wavelengths); SNF_CHECK_PTR(g_log_abs_coarse); SNF_CHECK_PTR(g_vol_c_coarse);
    SNF_CHECK_PTR(g_log_abs_fine); SNF_CHECK_PTR(g_vol_c_fine);
  }
  RenderWs w;
  if (render_layout(d, N, 1, ws, &w) < 0) return SNF_E_ARG;
  const int S = d->S, Sf = d->S + d->n_new;
  const float *gq = with_reg_grad ? w.g_q : nullptr;
  for (int fine = 1; fine >= 0; --fine) {   // fine first: its gradient bucket can be exchanged while the coarse pass runs
    const int s = fine ? Sf : S;
    const float *raw = fine ? w.raw_f : w.raw_c, *z = fine ? w.z_comb : w.z;
    float *g_raw = fine ? w.g_raw_f : w.g_raw_c;
    const float *g_img = fine ? g_fine_image : g_coarse_image;
    if (d->kind == 0)
      SNF_TRY(snf_composite_emission_bwd(raw, z, rays_d, N, s, g_img, fine ? gq : nullptr, g_raw, stream));
    else
      SNF_TRY(snf_composite_dt_bwd(raw, z, wavelengths, N, s, d->C, fine ? d->log_abs_fine : d->log_abs_coarse,
                                   fine ? d->vol_c_fine : d->vol_c_coarse, d->table_x, d->table_y, d->pixel_intensity_factor,
                                   g_img, fine ? gq : nullptr, g_raw, fine ? g_log_abs_fine : g_log_abs_coarse,
                                   fine ? g_vol_c_fine : g_vol_c_coarse, stream));
    const float *query = fine ? w.query_f : w.query_c;
    void *mws = fine ? w.mlp_f : w.mlp_c;
    if (d->mode == 1)
      SNF_TRY(snf_mlp_bwd_bf16(query, N * s, fine ? d->packed_fine : d->packed_coarse, g_raw, mws, fine ? gW_fine : gW_coarse,
                               fine ? gB_fine : gB_coarse, stream));
    else if (d->mode == 2)
      SNF_TRY(snf_mlp_bwd_x3(query, N * s, fine ? d->packed_fine : d->packed_coarse, g_raw, mws, fine ? gW_fine : gW_coarse,
                             fine ? gB_fine : gB_coarse, stream));
    else
      SNF_TRY(snf_mlp_bwd_f32(query, N * s, fine ? d->W_fine : d->W_coarse, d->n_hidden, d->d_filter, g_raw, mws,
                              fine ? gW_fine : gW_coarse, fine ? gB_fine : gB_coarse, stream));
  }
  return launch_status();
}
