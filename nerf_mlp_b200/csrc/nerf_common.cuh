// Shared definitions for libnerf_b200: flat parameter layout, error plumbing, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nerf_b200.h"

namespace nerf {

// ---- flat fp32 parameter layout: state_dict order of NeRFMLP (reference model.py:39-53) -------
// layer ids: 0..7 trunk, 8 sigma, 9 bottleneck, 10 view, 11 rgb
constexpr int kNumLayers = 12;
constexpr int kOut[kNumLayers] = {256, 256, 256, 256, 256, 256, 256, 256, 1, 256, 128, 3};
constexpr int kIn[kNumLayers]  = {63, 256, 256, 256, 256, 319, 256, 256, 256, 256, 283, 128};

constexpr int64_t w_off(int l) {
  int64_t o = 0;
  for (int i = 0; i < l; ++i) o += (int64_t)kOut[i] * kIn[i] + kOut[i];
  return o;
}
constexpr int64_t b_off(int l) { return w_off(l) + (int64_t)kOut[l] * kIn[l]; }
static_assert(w_off(kNumLayers) == NERF_N_PARAMS, "parameter count");

enum Layer { L_SIGMA = 8, L_BOTT = 9, L_VIEW = 10, L_RGB = 11 };

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();
int cuda_fail(cudaError_t e, const char* what);

#define NERF_CHECK_ARG(cond, ...)                         \
  do {                                                    \
    if (!(cond)) {                                        \
      nerf::set_error(__VA_ARGS__);                       \
      return -1;                                          \
    }                                                     \
  } while (0)

#define NERF_CUDA(expr)                                   \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return nerf::cuda_fail(_e, #expr); \
  } while (0)

#define NERF_LAUNCH_CHECK(name)                           \
  do {                                                    \
    cudaError_t _e = cudaGetLastError();                  \
    if (_e != cudaSuccess) return nerf::cuda_fail(_e, name); \
    nerf::count_launch();                                 \
  } while (0)

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- per-device state ------------------------------------------------------------------------
// Function attributes (cudaFuncSetAttribute), the SM count and the constant-memory schedules are
// per DEVICE, not per process: one process may drive several GPUs.  `current_device`
// returns the calling thread's device ordinal with its cached properties; `DeviceOnce` is a
// one-time flag per device for "set this kernel's attributes once".
constexpr int kMaxDevices = 64;
struct DeviceProps { int ordinal, sm_major, sm_minor, sm_count; };
int current_device(DeviceProps* out);
struct DeviceOnce {
  bool done[kMaxDevices] = {};
  bool needed(int dev) const { return dev < 0 || dev >= kMaxDevices || !done[dev]; }
  void mark(int dev) { if (dev >= 0 && dev < kMaxDevices) done[dev] = true; }
};

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Internal entry points implemented in the other translation units.
int mlp_fp32_workspace_floats_per_row(int save);
int mlp_fp32_forward(const float* rays_o, const float* rays_d, const float* z_vals, int R, int S,
                     float coord_scale, const float* x_enc, const float* d_enc, int64_t M,
                     const float* params, float* out, float* ws, size_t ws_bytes, int save,
                     cudaStream_t st);
int mlp_fp32_backward(const float* d_raw, int64_t M, const float* params, float* grads, float* ws,
                      size_t ws_bytes, cudaStream_t st);

size_t mlp_tc_packed_bytes();
int mlp_tc_pack(const float* params, void* packed, cudaStream_t st);
size_t mlp_tc_workspace_bytes(int64_t M, int save);
int mlp_tc_forward(const float* rays_o, const float* rays_d, const float* z_vals, int R, int S,
                   float coord_scale, const float* x_enc, const float* d_enc, int64_t M,
                   const float* params, const void* packed, float* out, void* ws, size_t ws_bytes,
                   int save, cudaStream_t st);
int mlp_tc_backward(const float* d_raw, int64_t M, int rows_per_dir, const float* params, const void* packed,
                    float* grads, void* ws, size_t ws_bytes, int stage, cudaStream_t st);

}  // namespace nerf
