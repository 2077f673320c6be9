// extern "C" surface of libnerf_b200.so (see include/nerf_b200.h): argument checks, error
// plumbing, precision dispatch.  No device allocation, no synchronisation.
#include "nerf_common.cuh"
#include <stdarg.h>

namespace nerf {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;   // kernels launched by this library (bench.py's gpu_launches)
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e > 0 ? (int)e : 1;
}

int current_device(DeviceProps* out) {
  static DeviceProps cache[kMaxDevices];
  static bool have[kMaxDevices] = {};
  int dev = 0;
  NERF_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < kMaxDevices && __atomic_load_n(&have[dev], __ATOMIC_ACQUIRE)) { *out = cache[dev]; return 0; }
  cudaDeviceProp p;
  NERF_CUDA(cudaGetDeviceProperties(&p, dev));
  DeviceProps d{dev, p.major, p.minor, p.multiProcessorCount};
  if (dev >= 0 && dev < kMaxDevices) { cache[dev] = d; __atomic_store_n(&have[dev], true, __ATOMIC_RELEASE); }
  *out = d;
  return 0;
}

}  // namespace nerf

using namespace nerf;

extern "C" const char* nerf_last_error(void) { return g_err; }
extern "C" unsigned long long nerf_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" const char* nerf_version(void) { return "nerf_b200 0.1 (sm_100a)"; }

extern "C" int nerf_device_info(int* sm, int* sm_count) {
  DeviceProps p;
  int rc = current_device(&p);
  if (rc) return rc;
  if (sm) *sm = p.sm_major * 10 + p.sm_minor;
  if (sm_count) *sm_count = p.sm_count;
  return 0;
}

extern "C" size_t nerf_packed_weight_bytes(void) { return mlp_tc_packed_bytes(); }

extern "C" int nerf_pack_weights(const float* flat_params, void* packed, void* stream) {
  NERF_CHECK_ARG(flat_params && packed, "nerf_pack_weights: null pointer");
  return mlp_tc_pack(flat_params, packed, (cudaStream_t)stream);
}

extern "C" size_t nerf_mlp_workspace_bytes(int64_t M, int precision, int save) {
  if (M <= 0) return 0;
  if (save == NERF_FWD_DENSITY_ONLY) save = 0;
  if (precision == NERF_PREC_FP32) {
    const int64_t rows = save ? M : (M < 262144 ? M : 262144);
    return (size_t)rows * mlp_fp32_workspace_floats_per_row(save) * sizeof(float);
  }
  return mlp_tc_workspace_bytes(M, save);
}

static int check_prec(int precision, const void* packed, const char* who) {
  NERF_CHECK_ARG(precision == NERF_PREC_BF16 || precision == NERF_PREC_FP32, "%s: unknown precision %d", who, precision);
  NERF_CHECK_ARG(precision == NERF_PREC_FP32 || packed != nullptr, "%s: bf16 mode needs the packed weight image", who);
  return 0;
}

extern "C" int nerf_mlp_fwd_rays(const float* rays_o, const float* rays_d, const float* z_vals, int R, int S,
                                 float coord_scale, const float* params, const void* packed, float* raw,
                                 void* workspace, size_t workspace_bytes, int precision, int save, void* stream) {
  NERF_CHECK_ARG(R >= 0 && S >= 1, "nerf_mlp_fwd_rays: bad shape R=%d S=%d", R, S);
  if (check_prec(precision, packed, "nerf_mlp_fwd_rays")) return -1;
  if (R == 0) return 0;
  NERF_CHECK_ARG(rays_o && rays_d && z_vals && params && raw, "nerf_mlp_fwd_rays: null pointer");
  NERF_CHECK_ARG(save >= 0 && save <= 2, "nerf_mlp_fwd_rays: save must be 0, 1 or NERF_FWD_DENSITY_ONLY (got %d)", save);
  const int64_t M = (int64_t)R * S;
  if (precision == NERF_PREC_FP32)     // check mode evaluates the whole network in every mode
    return mlp_fp32_forward(rays_o, rays_d, z_vals, R, S, coord_scale, nullptr, nullptr, M, params, raw,
                            (float*)workspace, workspace_bytes, save == 1, (cudaStream_t)stream);
  return mlp_tc_forward(rays_o, rays_d, z_vals, R, S, coord_scale, nullptr, nullptr, M, params, packed, raw,
                        workspace, workspace_bytes, save, (cudaStream_t)stream);
}

extern "C" int nerf_mlp_fwd_encoded(const float* x_enc, const float* d_enc, int64_t M, const float* params,
                                    const void* packed, float* out, void* workspace, size_t workspace_bytes,
                                    int precision, int save, void* stream) {
  NERF_CHECK_ARG(M >= 0, "nerf_mlp_fwd_encoded: bad M=%lld", (long long)M);
  if (check_prec(precision, packed, "nerf_mlp_fwd_encoded")) return -1;
  if (M == 0) return 0;
  NERF_CHECK_ARG(x_enc && d_enc && params && out, "nerf_mlp_fwd_encoded: null pointer");
  NERF_CHECK_ARG(save == 0 || save == 1, "nerf_mlp_fwd_encoded: save must be 0 or 1 (got %d)", save);
  if (precision == NERF_PREC_FP32)
    return mlp_fp32_forward(nullptr, nullptr, nullptr, 0, 1, 1.f, x_enc, d_enc, M, params, out,
                            (float*)workspace, workspace_bytes, save, (cudaStream_t)stream);
  return mlp_tc_forward(nullptr, nullptr, nullptr, 0, 1, 1.f, x_enc, d_enc, M, params, packed, out,
                        workspace, workspace_bytes, save, (cudaStream_t)stream);
}

extern "C" int nerf_mlp_bwd_stage(const float* d_raw, int64_t M, int rows_per_dir, const float* params, const void* packed,
                                  float* flat_grads, void* workspace, size_t workspace_bytes, int precision, int stage,
                                  void* stream) {
  NERF_CHECK_ARG(M >= 0, "nerf_mlp_bwd: bad M=%lld", (long long)M);
  NERF_CHECK_ARG(stage == NERF_BWD_ALL || stage == NERF_BWD_DGRAD || stage == NERF_BWD_WGRAD, "nerf_mlp_bwd: unknown stage %d", stage);
  if (check_prec(precision, packed, "nerf_mlp_bwd")) return -1;
  if (M == 0) return 0;
  NERF_CHECK_ARG(d_raw && params && flat_grads && workspace, "nerf_mlp_bwd: null pointer");
  if (precision == NERF_PREC_FP32) {
    if (stage == NERF_BWD_WGRAD) return 0;
    return mlp_fp32_backward(d_raw, M, params, flat_grads, (float*)workspace, workspace_bytes, (cudaStream_t)stream);
  }
  return mlp_tc_backward(d_raw, M, rows_per_dir, params, packed, flat_grads, workspace, workspace_bytes, stage,
                         (cudaStream_t)stream);
}

extern "C" int nerf_mlp_bwd(const float* d_raw, int64_t M, int rows_per_dir, const float* params, const void* packed,
                            float* flat_grads, void* workspace, size_t workspace_bytes, int precision,
                            void* stream) {
  return nerf_mlp_bwd_stage(d_raw, M, rows_per_dir, params, packed, flat_grads, workspace, workspace_bytes, precision,
                            NERF_BWD_ALL, stream);
}
