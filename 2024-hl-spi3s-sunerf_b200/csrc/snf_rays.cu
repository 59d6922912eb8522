// Observer image rays generated on the device: replaces the host numpy path
//   pose_spherical  sunerf/train/coordinate_transformation.py:36-54   (4x4 pose: stays on the host, 16 floats)
//   get_rays        sunerf/data/ray_sampling.py:7-36                  (this kernel)
// for the regular helioprojective pixel grid Tx = (j - cx) p, Ty = (i - cy) p.  Same arithmetic as numpy: the
// direction cosines in double, rounded to float32, then a float32 3x3 product with ((a+b)+c) association.
#include "snf_common.cuh"

namespace snf {

struct Pose { float r[9]; float o[3]; };

__global__ void __launch_bounds__(256) image_rays_kernel(const Pose pose, int W, double plate_arcsec, double asec,
                                                         double cx, double cy, int64_t first, int64_t count,
                                                         float *__restrict__ rays_o, float *__restrict__ rays_d) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int64_t idx = first + t;
  const int64_t i = idx / W, j = idx - i * W;
  const double Tx = __dmul_rn(__dmul_rn((double)j - cx, plate_arcsec), asec);      // (jj - c) * plate * asec, left to right
  const double Ty = __dmul_rn(__dmul_rn((double)i - cy, plate_arcsec), asec);
  double sx, cxx, sy, cyy;
  sincos(Tx, &sx, &cxx);
  sincos(Ty, &sy, &cyy);
  const float dx = (float)sx;                                                       // ray_sampling.py:15-17
  const float dy = (float)__dmul_rn(-sy, cxx);
  const float dz = (float)__dmul_rn(-cxx, cyy);
#pragma unroll
  for (int k = 0; k < 3; ++k) {                                                     // :29  sum(dir * c2w[k,:3])
    rays_d[3 * t + k] = fadd(fadd(fmul(dx, pose.r[3 * k]), fmul(dy, pose.r[3 * k + 1])), fmul(dz, pose.r[3 * k + 2]));
    rays_o[3 * t + k] = pose.o[k];                                                  // :35
  }
}

}  // namespace snf

using namespace snf;

extern "C" int snf_image_rays(const float *c2w_host /*[3][4] or [4][4] row-major, HOST*/, int H, int W,
                              double plate_arcsec, double asec, double cx, double cy, int64_t first, int64_t count,
                              float *rays_o, float *rays_d, void *stream) {
  SNF_CHECK_PTR(c2w_host); SNF_CHECK_PTR(rays_o); SNF_CHECK_PTR(rays_d);
  if (H <= 0 || W <= 0 || first < 0 || count < 0 || first + count > (int64_t)H * W) return SNF_E_ARG;
  if (count == 0) return 0;
  Pose p;
  for (int k = 0; k < 3; ++k) {
    for (int c = 0; c < 3; ++c) p.r[3 * k + c] = c2w_host[4 * k + c];
    p.o[k] = c2w_host[4 * k + 3];
  }
  image_rays_kernel<<<(unsigned)ceil_div64(count, 256), 256, 0, (cudaStream_t)stream>>>(p, W, plate_arcsec, asec, cx, cy, first,
                                                                                    count, rays_o, rays_d);
  count_launch();
  return launch_status();
}
