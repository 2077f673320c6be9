// The callers either side of the hot path, on the device:
//
//  * ray generation + batch gather: the reference precomputes every ray of every training image on
//    the host (nerfmlp/data.py:76-97, float64 numpy), then serves them one by one through
//    Dataset.__getitem__ (:99-104) and DataLoader collate (scripts/train.py:219,368-371).  Here a
//    ray is rebuilt from (pose, pixel) when it is needed -- 64 B of pose per image instead of 24 B
//    per ray -- and the target colour is fetched (and, for raw RGBA bytes, alpha-composited and
//    sRGB-decoded, data.py:43-62 / :8-22) in the same pass.  The same kernel generates the
//    contiguous pixel range of one view for rendering (scripts/render_example.py:245-250).
//  * image post-processing: brightness, linear->sRGB, clip, 8-bit quantisation
//    (scripts/render_example.py:12-26,256-271).
//
// Arithmetic follows numpy's: directions are computed in float64 ((i - W/2)/focal etc., the 3x3
// rotation applied as float64 products summed left to right) and rounded once to float32, which
// is what `torch.from_numpy(...).float()` does to the reference's float64 tables.
#include "nerf_common.cuh"

namespace nerf {

struct RayGenArgs {
  const float* poses;      // [n_poses, 4, 4] row-major camera-to-world (transforms_*.json 'transform_matrix')
  const int64_t* idx;      // [n] flat ray ids (img * H * W + j * W + i), or null -> first + k
  int64_t first, n;
  int n_poses, H, W;
  double focal;
  float* rays_o;           // [n, 3]
  float* rays_d;           // [n, 3]
  const uint8_t* rgba;     // [n_poses, H, W, 4] raw 8-bit RGBA (nullable)
  const float* rgb_lin;    // [n_poses, H, W, 3] preprocessed linear RGB (nullable)
  int white_bkgd;
  float* rgb_out;          // [n, 3] (nullable)
};

// data.py:8-22 (float32 where/power, like numpy on a float32 array)
__device__ __forceinline__ float srgb_to_linear_f(float v) {
  return v <= 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
}

__global__ void raygen_kernel(const RayGenArgs a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n) return;
  const int64_t id = a.idx != nullptr ? a.idx[k] : a.first + k;
  const int64_t hw = (int64_t)a.H * a.W;
  if (id < 0 || id >= hw * a.n_poses) {                                  // bad index: poison, never read out of bounds
    const float q = __int_as_float(0x7fc00000);
    for (int c = 0; c < 3; ++c) {
      a.rays_o[k * 3 + c] = q; a.rays_d[k * 3 + c] = q;
      if (a.rgb_out != nullptr) a.rgb_out[k * 3 + c] = q;
    }
    return;
  }
  const int64_t img = id / hw, pix = id - img * hw;
  const int j = (int)(pix / a.W), i = (int)(pix - (int64_t)j * a.W);
  const float* P = a.poses + img * 16;
  // dirs = [(i - W/2)/focal, -(j - H/2)/focal, -1]                     data.py:80 (float64)
  const double dx = ((double)i - (double)a.W / 2.0) / a.focal;
  const double dy = -((double)j - (double)a.H / 2.0) / a.focal;
  const double dz = -1.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // rays_d = dirs @ pose[:3,:3].T                                     data.py:86
    const double v = __dadd_rn(__dadd_rn(__dmul_rn(dx, (double)P[4 * c + 0]), __dmul_rn(dy, (double)P[4 * c + 1])),
                               __dmul_rn(dz, (double)P[4 * c + 2]));
    a.rays_d[k * 3 + c] = (float)v;
    a.rays_o[k * 3 + c] = P[4 * c + 3];                                  // data.py:87
  }
  if (a.rgb_out == nullptr) return;
  if (a.rgb_lin != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) a.rgb_out[k * 3 + c] = a.rgb_lin[id * 3 + c];
  } else {
    const uchar4 px = reinterpret_cast<const uchar4*>(a.rgba)[id];
    const double al = (double)px.w / 255.0;                              // data.py:47 (float64 /255)
    const uint8_t ch[3] = {px.x, px.y, px.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double v = (double)ch[c] / 255.0;
      if (a.white_bkgd) v = v * al + (1.0 - al);                         // data.py:55 (float64)
      a.rgb_out[k * 3 + c] = srgb_to_linear_f((float)v);                 // data.py:62 -> :16 astype(float32)
    }
  }
}

// render_example.py:256-271: rgb * boost; optional linear_to_srgb (:12-26, float32); clip(0,1)*255 -> uint8
__global__ void postprocess_kernel(const float* __restrict__ rgb, int64_t n, float boost, int to_srgb,
                                   uint8_t* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float v = rgb[k];
  if (boost != 1.0f) v = __fmul_rn(v, boost);
  if (to_srgb) v = v <= 0.0031308f ? __fmul_rn(v, 12.92f) : __fsub_rn(__fmul_rn(1.055f, powf(v, (float)(1.0 / 2.4))), 0.055f);
  v = fminf(fmaxf(v, 0.f), 1.f);                                         // np.clip (NaN -> NaN -> 0 below)
  const float s = __fmul_rn(v, 255.f);
  out[k] = (s == s) ? (uint8_t)s : (uint8_t)0;                           // astype(uint8): truncation
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_generate_rays(const float* poses, int n_poses, int H, int W, double focal, const int64_t* idx,
                                  int64_t first, int64_t n, float* rays_o, float* rays_d, const uint8_t* rgba,
                                  const float* rgb_lin, int white_bkgd, float* rgb_out, void* stream) {
  NERF_CHECK_ARG(n >= 0 && n_poses >= 1 && H >= 1 && W >= 1 && focal > 0.0, "nerf_generate_rays: bad shape n=%lld n_poses=%d H=%d W=%d focal=%g",
                 (long long)n, n_poses, H, W, focal);
  if (n == 0) return 0;
  NERF_CHECK_ARG(poses && rays_o && rays_d, "nerf_generate_rays: null pointer");
  NERF_CHECK_ARG(idx != nullptr || (first >= 0 && first + n <= (int64_t)n_poses * H * W),
                 "nerf_generate_rays: ray range [%lld, %lld) outside %d images of %dx%d", (long long)first,
                 (long long)(first + n), n_poses, H, W);
  NERF_CHECK_ARG(rgb_out == nullptr || rgba != nullptr || rgb_lin != nullptr, "nerf_generate_rays: rgb_out needs an image source");
  RayGenArgs a{poses, idx, first, n, n_poses, H, W, focal, rays_o, rays_d, rgba, rgb_lin, white_bkgd, rgb_out};
  raygen_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a);
  NERF_LAUNCH_CHECK("raygen_kernel");
  return 0;
}

extern "C" int nerf_postprocess_rgb8(const float* rgb, int64_t n, float brightness, int to_srgb, uint8_t* out,
                                     void* stream) {
  NERF_CHECK_ARG(n >= 0, "nerf_postprocess_rgb8: bad n=%lld", (long long)n);
  if (n == 0) return 0;
  NERF_CHECK_ARG(rgb && out, "nerf_postprocess_rgb8: null pointer");
  postprocess_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(rgb, n, brightness, to_srgb, out);
  NERF_LAUNCH_CHECK("postprocess_kernel");
  return 0;
}
