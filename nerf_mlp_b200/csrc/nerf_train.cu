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

// ------------------------------------------------------------------------------------------
// Data-parallel step, gradient exchange + Adam in ONE kernel over NVLink peer memory (no NCCL call):
// every rank's flat gradient lives in a symmetric (peer-mapped) buffer; each rank reads all `world` gradients
// straight out of the peers' HBM (one-shot all-reduce: ld.volatile.v4 over NVLink, all `world` loads of an element in
// flight before the first add), sums them IN RANK ORDER -- so every rank computes bit-identical sums and
// parameters -- and applies the update.  What NCCL's all-reduce + the separate Adam launch did in two kernels and
// three passes over the gradient is one pass here.
//
// Hand-shake (u32 flags in each rank's symmetric block, epoch e = 1, 2, ... kept in the block itself so that the
// kernel can be replayed from a CUDA graph):
//   flags[r]          rank r's gradient of epoch >= value is complete              (written by rank r, st.release.sys)
//   flags[kPeerMax+r] rank r has finished reading MY gradient of epoch >= value     (written by rank r)
//   flags[2 kPeerMax] my epoch (advanced by the last block, after everything else)
// Start: block 0 raises "my gradient is complete" at every peer; every block waits until all `world` gradients are.
// End: the last block of the grid raises "I have read yours" at every peer and waits for theirs, so the kernel (and
// with it the stream) does not pass until nobody reads this rank's gradient buffer any more -- the next step may
// zero it.  Waits are bounded (~18 s) and trap.
// ------------------------------------------------------------------------------------------
constexpr int kPeerMax = NERF_PEER_MAX;
struct PeerArgs {
  const float* grads[kPeerMax];
  float* red[kPeerMax];                 // two-shot only: every rank's buffer for the reduced gradient
  uint32_t* flags[kPeerMax];
  int rank, world;
};
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {        // not cached in L1: the peer rewrites it every step
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_wait(const uint32_t* flag, uint32_t e, int what, int r) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < e) {
    __nanosleep(40);
    if (clock64() - t0 > (1ll << 35)) {                  // ~18 s: ranks may be seconds apart in their first (lazy-init) step
      printf("nerf_b200: peer gradient exchange: wait timeout (%s of rank %d, epoch %u, have %u)\n",
             what == 0 ? "gradient-ready flag" : "read-done flag", r, e, ld_acquire_sys(flag));
      __trap();
    }
  }
}

__global__ void __launch_bounds__(256) adam_fused_peer_kernel(float* __restrict__ p, const PeerArgs pa, float* __restrict__ m,
                                                              float* __restrict__ v, int64_t n, double* __restrict__ st,
                                                              const float* __restrict__ loss, double* __restrict__ scratch) {
  __shared__ float sc[7];
  __shared__ double part[8];
  __shared__ bool last;
  __shared__ uint32_t s_epoch;
  uint32_t* mine = pa.flags[pa.rank];
  if (threadIdx.x == 0) {
    const double lr = st[0], b1 = st[1], b2 = st[2], step = st[5] + 1.0;
    const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    sc[0] = (float)(1.0 - b1); sc[1] = (float)b2; sc[2] = (float)(1.0 - b2);
    sc[3] = (float)(-(lr / bc1)); sc[4] = (float)sqrt(bc2); sc[5] = (float)st[3]; sc[6] = (float)st[4];
    s_epoch = *reinterpret_cast<volatile uint32_t*>(mine + 2 * kPeerMax) + 1u;
  }
  __syncthreads();
  const uint32_t e = s_epoch;
  if (blockIdx.x == 0 && (int)threadIdx.x < pa.world) st_release_sys(pa.flags[threadIdx.x] + pa.rank, e);
  if ((int)threadIdx.x < pa.world) peer_wait(mine + threadIdx.x, e, 0, (int)threadIdx.x);
  __syncthreads();

  auto update = [&](float graw, float& pi, float& mi, float& vi) {
    const float gi = graw * sc[6];
    mi = __fadd_rn(mi, __fmul_rn(__fsub_rn(gi, mi), sc[0]));
    vi = __fadd_rn(__fmul_rn(vi, sc[1]), __fmul_rn(__fmul_rn(gi, gi), sc[2]));
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), sc[4]), sc[5]);
    pi = __fadd_rn(pi, __fdiv_rn(__fmul_rn(sc[3], mi), denom));
  };
  double sq = 0.0;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 t[kPeerMax];
#pragma unroll
    for (int r = 0; r < kPeerMax; ++r)
      if (r < pa.world) t[r] = ld_peer_v4(pa.grads[r] + 4 * i);
    float4 g = t[0];
#pragma unroll
    for (int r = 1; r < kPeerMax; ++r)
      if (r < pa.world) { g.x = __fadd_rn(g.x, t[r].x); g.y = __fadd_rn(g.y, t[r].y); g.z = __fadd_rn(g.z, t[r].z); g.w = __fadd_rn(g.w, t[r].w); }
    sq += (double)g.x * (double)g.x + (double)g.y * (double)g.y + (double)g.z * (double)g.z + (double)g.w * (double)g.w;
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    update(g.x, pp.x, mm.x, vv.x); update(g.y, pp.y, mm.y, vv.y); update(g.z, pp.z, mm.z, vv.z); update(g.w, pp.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && (int64_t)threadIdx.x < n - 4 * n4) {                      // tail (n not a multiple of 4)
    const int64_t i = 4 * n4 + threadIdx.x;
    float g = ld_peer(pa.grads[0] + i);
    for (int r = 1; r < pa.world; ++r) g = __fadd_rn(g, ld_peer(pa.grads[r] + i));
    sq += (double)g * (double)g;
    update(g, p[i], m[i], v[i]);
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
  // last block: every block of this rank has read the peers' gradients -> tell them, and wait until they have read ours
  __threadfence();
  if ((int)threadIdx.x < pa.world) {
    st_release_sys(pa.flags[threadIdx.x] + kPeerMax + pa.rank, e);
    peer_wait(mine + kPeerMax + threadIdx.x, e, 1, (int)threadIdx.x);
  }
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
    *reinterpret_cast<volatile uint32_t*>(mine + 2 * kPeerMax) = e;               // epoch done
  }
}

// Two-shot variant (world >= 4): reduce-scatter + all-gather over peer memory, then the update.  Rank r sums slice r of the
// `world` gradients (in rank order) and STORES the result into every rank's `red` buffer (fire-and-forget remote
// stores); once all slices have landed, every rank applies Adam to the whole reduced gradient out of its own HBM.
// NVLink traffic per rank: 2 (world-1)/world x n floats instead of (world-1) x n (at 8 GPUs 4.2 MB instead of 16.7 MB);
// parameters and moments stay replicated and bit-identical, so nothing else in the framework changes.
// Flags: [r] gradient ready (as above); [kPeerMax + r] "rank r's slice is in your red buffer" -- which also means rank r
// has finished reading everybody's gradient, so the one-shot kernel's closing hand-shake is not needed: a kernel that
// has passed the slice wait may exit.  [2 kPeerMax + 1]: block counter of the scatter phase (self-resetting).
__device__ __forceinline__ void st_peer_v4(float* p, float4 v) {
  asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_peer(float* p, float v) { asm volatile("st.volatile.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

__global__ void __launch_bounds__(256, 4) adam_fused_peer2_kernel(float* __restrict__ p, const PeerArgs pa, float* __restrict__ m,
                                                               float* __restrict__ v, int64_t n, double* __restrict__ st,
                                                               const float* __restrict__ loss, double* __restrict__ scratch) {
  __shared__ float sc[7];
  __shared__ double part[8];
  __shared__ bool last;
  __shared__ uint32_t s_epoch;
  uint32_t* mine = pa.flags[pa.rank];
  if (threadIdx.x == 0) {
    const double lr = st[0], b1 = st[1], b2 = st[2], step = st[5] + 1.0;
    const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    sc[0] = (float)(1.0 - b1); sc[1] = (float)b2; sc[2] = (float)(1.0 - b2);
    sc[3] = (float)(-(lr / bc1)); sc[4] = (float)sqrt(bc2); sc[5] = (float)st[3]; sc[6] = (float)st[4];
    s_epoch = *reinterpret_cast<volatile uint32_t*>(mine + 2 * kPeerMax) + 1u;
  }
  __syncthreads();
  const uint32_t e = s_epoch;
  if (blockIdx.x == 0 && (int)threadIdx.x < pa.world) st_release_sys(pa.flags[threadIdx.x] + pa.rank, e);
  if ((int)threadIdx.x < pa.world) peer_wait(mine + threadIdx.x, e, 0, (int)threadIdx.x);
  __syncthreads();

  // ---- reduce-scatter: this rank's slice, result broadcast into every rank's `red` ----
  const int64_t n4 = n >> 2;
  const int64_t per = (n4 + pa.world - 1) / pa.world;
  const int64_t lo = per * pa.rank, hi = (lo + per < n4) ? lo + per : n4;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    float4 t[kPeerMax];
#pragma unroll
    for (int r = 0; r < kPeerMax; ++r)
      if (r < pa.world) t[r] = ld_peer_v4(pa.grads[r] + 4 * i);
    float4 g = t[0];
#pragma unroll
    for (int r = 1; r < kPeerMax; ++r)
      if (r < pa.world) { g.x = __fadd_rn(g.x, t[r].x); g.y = __fadd_rn(g.y, t[r].y); g.z = __fadd_rn(g.z, t[r].z); g.w = __fadd_rn(g.w, t[r].w); }
#pragma unroll
    for (int r = 0; r < kPeerMax; ++r)
      if (r < pa.world) st_peer_v4(pa.red[r] + 4 * i, g);
  }
  if (pa.rank == pa.world - 1 && blockIdx.x == 0 && (int64_t)threadIdx.x < n - 4 * n4) {   // tail (n not a multiple of 4)
    const int64_t i = 4 * n4 + threadIdx.x;
    float g = ld_peer(pa.grads[0] + i);
    for (int r = 1; r < pa.world; ++r) g = __fadd_rn(g, ld_peer(pa.grads[r] + i));
    for (int r = 0; r < pa.world; ++r) st_peer(pa.red[r] + i, g);
  }
  __threadfence_system();                                  // this thread's remote stores are performed before the block is counted
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* ctr = mine + 2 * kPeerMax + 1;
    last = atomicAdd(ctr, 1u) == gridDim.x - 1;
    if (last) *reinterpret_cast<volatile uint32_t*>(ctr) = 0u;
  }
  __syncthreads();
  if (last) {                                              // the whole grid of this rank has scattered its slice
    __threadfence_system();
    if ((int)threadIdx.x < pa.world) st_release_sys(pa.flags[threadIdx.x] + kPeerMax + pa.rank, e);
  }
  if ((int)threadIdx.x < pa.world) peer_wait(mine + kPeerMax + threadIdx.x, e, 1, (int)threadIdx.x);
  __syncthreads();

  // ---- the update, on the reduced gradient in this rank's own memory (same arithmetic as adam_fused_kernel) ----
  const float* red = pa.red[pa.rank];
  auto update = [&](float graw, float& pi, float& mi, float& vi) {
    const float gi = graw * sc[6];
    mi = __fadd_rn(mi, __fmul_rn(__fsub_rn(gi, mi), sc[0]));
    vi = __fadd_rn(__fmul_rn(vi, sc[1]), __fmul_rn(__fmul_rn(gi, gi), sc[2]));
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), sc[4]), sc[5]);
    pi = __fadd_rn(pi, __fdiv_rn(__fmul_rn(sc[3], mi), denom));
  };
  double sq = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = ld_peer_v4(red + 4 * i);               // written by remote stores: not through L1
    sq += (double)g.x * (double)g.x + (double)g.y * (double)g.y + (double)g.z * (double)g.z + (double)g.w * (double)g.w;
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    update(g.x, pp.x, mm.x, vv.x); update(g.y, pp.y, mm.y, vv.y); update(g.z, pp.z, mm.z, vv.z); update(g.w, pp.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && (int64_t)threadIdx.x < n - 4 * n4) {
    const int64_t i = 4 * n4 + threadIdx.x;
    const float g = ld_peer(red + i);
    sq += (double)g * (double)g;
    update(g, p[i], m[i], v[i]);
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
    *reinterpret_cast<volatile uint32_t*>(mine + 2 * kPeerMax) = e;               // epoch done
  }
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_adam_step_fused_peer(float* params, const float* const* peer_grads, float* const* peer_red,
                                         uint32_t* const* peer_flags, int rank, int world, float* exp_avg,
                                         float* exp_avg_sq, int64_t n, double* state, const float* loss, void* scratch,
                                         void* stream) {
  NERF_CHECK_ARG(params && peer_grads && peer_flags && exp_avg && exp_avg_sq && state && scratch && n >= 1,
                 "nerf_adam_step_fused_peer: bad arguments");
  NERF_CHECK_ARG(world >= 1 && world <= kPeerMax && rank >= 0 && rank < world,
                 "nerf_adam_step_fused_peer: rank %d / world %d (at most %d peers)", rank, world, kPeerMax);
  NERF_CHECK_ARG((((uintptr_t)state | (uintptr_t)scratch) & 7) == 0, "nerf_adam_step_fused_peer: state/scratch must be 8-byte aligned");
  NERF_CHECK_ARG((((uintptr_t)params | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                 "nerf_adam_step_fused_peer: params / moments must be 16-byte aligned");
  PeerArgs pa{};
  pa.rank = rank; pa.world = world;
  for (int r = 0; r < world; ++r) {
    NERF_CHECK_ARG(peer_grads[r] && peer_flags[r] && ((uintptr_t)peer_grads[r] & 15) == 0, "nerf_adam_step_fused_peer: peer %d pointers", r);
    pa.grads[r] = peer_grads[r];
    pa.flags[r] = peer_flags[r];
    if (peer_red != nullptr) {
      NERF_CHECK_ARG(peer_red[r] && ((uintptr_t)peer_red[r] & 15) == 0, "nerf_adam_step_fused_peer: peer %d reduced-gradient buffer", r);
      pa.red[r] = peer_red[r];
    }
  }
  const int64_t n4 = n >> 2;
  int blocks = (int)(ceil_div(n4 > 0 ? n4 : 1, 256) < 592 ? ceil_div(n4 > 0 ? n4 : 1, 256) : 592);
  // The two-shot kernel's blocks wait (for the peers' slices) while other blocks of the same grid still have to run
  // their scatter phase: its whole grid must be co-resident, so the grid is capped at what the occupancy calculator
  // says fits (the kernel is bounded to 64 registers: 4 blocks per SM).  The one-shot kernel's blocks never wait for
  // another block of their own grid, but the same cap costs nothing (both kernels are grid-stride loops).
  DeviceProps dp;
  int rc = current_device(&dp);
  if (rc) return rc;
  static int occ[2][kMaxDevices];
  static DeviceOnce occ_done;
  if (occ_done.needed(dp.ordinal)) {
    int o1 = 0, o2 = 0;
    NERF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, adam_fused_peer_kernel, 256, 0));
    NERF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, adam_fused_peer2_kernel, 256, 0));
    if (dp.ordinal >= 0 && dp.ordinal < kMaxDevices) { occ[0][dp.ordinal] = o1 < 1 ? 1 : o1; occ[1][dp.ordinal] = o2 < 1 ? 1 : o2; }
    occ_done.mark(dp.ordinal);
  }
  const int per_sm = (dp.ordinal >= 0 && dp.ordinal < kMaxDevices) ? occ[peer_red != nullptr ? 1 : 0][dp.ordinal] : 1;
  if (blocks > per_sm * dp.sm_count) blocks = per_sm * dp.sm_count;
  if (peer_red != nullptr)
    adam_fused_peer2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, pa, exp_avg, exp_avg_sq, n, state, loss, (double*)scratch);
  else
    adam_fused_peer_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, pa, exp_avg, exp_avg_sq, n, state, loss, (double*)scratch);
  NERF_LAUNCH_CHECK("adam_fused_peer_kernel");
  return 0;
}

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
