// Training-step glue that keeps the whole optimisation step on the device (SURVEY.md section 8f
// row 1): the reference reads loss / PSNR / gradient norm back to the host every step
// (scripts/train.py:376-390: .item(), .cpu().numpy(), 24x .item()) and passes the Adam scalars from
// Python.  Here the step counter, the learning rate and the derived Adam scalars live in a small
// device-resident state block, so that the step can be captured once as a CUDA graph and replayed
// while the schedule advances, and the three metrics are produced without a host sync.
//
// state block (doubles): [0] lr  [1] beta1  [2] beta2  [3] eps  [4] grad_scale  [5] step
//   derived per step:    [6] 1-beta1  [7] 1-beta2  [8] -(lr/bias_correction1)  [9] sqrt(bias_correction2)
//   metrics:             [10] loss  [11] psnr  [12] grad_norm
//   scratch:             [15] block counter (as u32), [16..16+kNormBlocks) partial sums
#include "nerf_common.cuh"

namespace nerf {

constexpr int kNormBlocks = 64;
static_assert(NERF_TRAIN_STATE_DOUBLES >= 16 + kNormBlocks, "state block too small");

// One launch, kNormBlocks blocks: sum of squares of the gradient (fp64 partials, combined in block
// order by the last block to finish => deterministic), then the per-step scalars.
__global__ void train_prepare_kernel(double* __restrict__ st, const float* __restrict__ loss,
                                     const float* __restrict__ g, int64_t n) {
  __shared__ double part[8];
  __shared__ bool last;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)g[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    st[16 + blockIdx.x] = s;
    __threadfence();
    unsigned int* counter = reinterpret_cast<unsigned int*>(st + 15);
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  double ss = 0.0;
  for (int b = 0; b < (int)gridDim.x; ++b) ss += *(volatile double*)(st + 16 + b);
  *reinterpret_cast<unsigned int*>(st + 15) = 0u;             // ready for the next launch / graph replay
  const double lr = st[0], b1 = st[1], b2 = st[2], scale = st[4];
  const double step = st[5] + 1.0;
  st[5] = step;
  // python-double scalar arithmetic of torch/optim/adam.py (_single_tensor_adam)
  const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
  st[6] = 1.0 - b1;
  st[7] = 1.0 - b2;
  st[8] = -(lr / bc1);
  st[9] = sqrt(bc2);
  const double l = loss != nullptr ? (double)*loss : 0.0;
  st[10] = l;
  st[11] = l > 0.0 ? 10.0 * log10(1.0 / l) : INFINITY;        // skimage psnr, data_range = 1 (scripts/train.py:33-37)
  st[12] = sqrt(ss) * fabs(scale);                            // |grad|_2 of the (averaged) gradient (scripts/train.py:60-67)
}

// Same update as adam_kernel (nerf_render.cu), scalars read from the state block.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const double* __restrict__ st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float one_minus_b1 = (float)st[6], b2 = (float)st[2], one_minus_b2 = (float)st[7];
  const float neg_step_size = (float)st[8], bc2_sqrt = (float)st[9], eps = (float)st[3], grad_scale = (float)st[4];
  const float gi = g[i] * grad_scale;
  float mi = m[i], vi = v[i];
  mi = __fadd_rn(mi, __fmul_rn(__fsub_rn(gi, mi), one_minus_b1));
  vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(__fmul_rn(gi, gi), one_minus_b2));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
  p[i] = __fadd_rn(p[i], __fdiv_rn(__fmul_rn(neg_step_size, mi), denom));
  m[i] = mi;
  v[i] = vi;
}

// train_prepare_kernel + adam_dev_kernel in one launch: every block derives the step's scalars from the
// state block itself (thread 0, double arithmetic), updates its 256 parameters and leaves the sum of
// squares of its gradients in `scratch`; the last block to finish adds the partials in block order,
// fills the metrics and advances the step counter (no other block reads the state after that).
__global__ void __launch_bounds__(256) adam_fused_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t n, double* __restrict__ st,
                                                         const float* __restrict__ loss, double* __restrict__ scratch) {
  __shared__ float sc[7];
  __shared__ double part[8];
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const double lr = st[0], b1 = st[1], b2 = st[2], step = st[5] + 1.0;
    const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    sc[0] = (float)(1.0 - b1); sc[1] = (float)b2; sc[2] = (float)(1.0 - b2);
    sc[3] = (float)(-(lr / bc1)); sc[4] = (float)sqrt(bc2); sc[5] = (float)st[3]; sc[6] = (float)st[4];
  }
  __syncthreads();
  double sq = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float graw = g[i];
    sq += (double)graw * (double)graw;
    const float gi = graw * sc[6];
    float mi = m[i], vi = v[i];
    mi = __fadd_rn(mi, __fmul_rn(__fsub_rn(gi, mi), sc[0]));
    vi = __fadd_rn(__fmul_rn(vi, sc[1]), __fmul_rn(__fmul_rn(gi, gi), sc[2]));
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), sc[4]), sc[5]);
    p[i] = __fadd_rn(p[i], __fdiv_rn(__fmul_rn(sc[3], mi), denom));
    m[i] = mi;
    v[i] = vi;
  }
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += part[w];
    scratch[1 + blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned int*>(scratch), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // last block: all 256 threads add the per-block partials (fixed assignment and tree => deterministic)
  __threadfence();
  double ss = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) ss += *(volatile double*)(scratch + 1 + b);
  ss = warp_sum(ss);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    ss = 0.0;
    for (int w = 0; w < 8; ++w) ss += part[w];
    *reinterpret_cast<unsigned int*>(scratch) = 0u;
    const double l = loss != nullptr ? (double)*loss : 0.0;
    st[5] = st[5] + 1.0;
    st[10] = l;
    st[11] = l > 0.0 ? 10.0 * log10(1.0 / l) : INFINITY;
    st[12] = sqrt(ss) * fabs(st[4]);
  }
}

}  // namespace nerf

using namespace nerf;

extern "C" size_t nerf_adam_fused_scratch_bytes(int64_t n) { return (size_t)(1 + ceil_div(n > 0 ? n : 1, 256)) * sizeof(double); }

extern "C" int nerf_adam_step_fused(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                    double* state, const float* loss, void* scratch, void* stream) {
  NERF_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state && scratch && n >= 1, "nerf_adam_step_fused: bad arguments");
  NERF_CHECK_ARG((((uintptr_t)state | (uintptr_t)scratch) & 7) == 0, "nerf_adam_step_fused: state/scratch must be 8-byte aligned");
  // one wave of blocks: every block pays the two double-precision pow() of the bias corrections once
  const int blocks = ceil_div(n, 256) < 592 ? ceil_div(n, 256) : 592;
  adam_fused_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, state, loss,
                                                              (double*)scratch);
  NERF_LAUNCH_CHECK("adam_fused_kernel");
  return 0;
}

extern "C" int nerf_train_prepare(double* state, const float* loss, const float* flat_grads, int64_t n, void* stream) {
  NERF_CHECK_ARG(state != nullptr && flat_grads != nullptr && n >= 1, "nerf_train_prepare: bad arguments");
  NERF_CHECK_ARG(((uintptr_t)state & 7) == 0, "nerf_train_prepare: state must be 8-byte aligned");
  train_prepare_kernel<<<kNormBlocks, 256, 0, (cudaStream_t)stream>>>(state, loss, flat_grads, n);
  NERF_LAUNCH_CHECK("train_prepare_kernel");
  return 0;
}

extern "C" int nerf_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                  const double* state, void* stream) {
  NERF_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state && n >= 0, "nerf_adam_step_dev: bad arguments");
  if (n == 0) return 0;
  adam_dev_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, state);
  NERF_LAUNCH_CHECK("adam_dev_kernel");
  return 0;
}
