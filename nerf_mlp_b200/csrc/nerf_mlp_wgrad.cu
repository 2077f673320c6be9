// Weight / bias gradients of the bf16 tensor-core backward (the autograd of reference
// model.py:57-81 w.r.t. its 24 parameters, triggered at scripts/train.py:382).
//
//   dW_l[out,in] = sum_rows dY_l[row,out] * X_l[row,in]      (X_l, dY_l: bf16 [M, features] in HBM,
//                                                             written by the forward / dgrad kernels)
//
// is a GEMM whose contraction runs over the sample rows, so BOTH operands are "MN-major" for the
// tensor core (features contiguous, rows strided).  One CTA owns one (layer, row-slab) job:
//   loaders  (4 warps)  cp.async 16-byte chunks of a 64-row chunk of dY and X into a 3-stage ring,
//                       in the 128B-swizzled MN-major layout UMMA expects
//   issuer   (1 warp)   per chunk 4 K-steps x (1|2 M-halves) tcgen05.mma 128 x N x 16, fp32
//                       accumulators for the whole dW block stay in TMEM (2 x 256 columns)
//   bias     (4 warps)  column sums of the dY chunk straight from shared memory (db_l), then, after
//                       the last chunk, the epilogue: TMEM -> registers -> red.global.add.f32 into
//                       the flat fp32 gradient buffer (split-K reduction across the slabs).
// The traffic is the HBM roofline of this stage: every dY and X row is read once.
// The tiny heads (rgb 128->3, sigma 256->1) and the 27 view-direction columns of view_linear run
// on CUDA cores.
#include "tc_common.cuh"

namespace nerf {
using namespace ptx;

constexpr int kWgChunkRows = 64;
constexpr int kWgStages = 3;
constexpr int kWgStageBytes = 65536;                       // A: 64 x 256 bf16 (32 KB) + B: 64 x 256 bf16 (32 KB)
constexpr int kWgOffB = 32768;
constexpr int kWgLoaderWarps = 4, kWgBiasWarps = 4;
constexpr int kWgThreads = 32 * (kWgLoaderWarps + 1 + kWgBiasWarps);
constexpr int kWgOffBar = kWgStages * kWgStageBytes;
constexpr int kWgSmemBytes = kWgOffBar + 128;
constexpr int kMaxJobs = 16;

struct WgradJob {
  const __nv_bfloat16* A; int lda;      // dY [M, lda]; columns [0, 128*mh) are used
  const __nv_bfloat16* B; int ldb;      // X  [M, ldb]; columns [0, n) are used
  int mh;                               // M halves of 128 output features (1 or 2)
  int n;                                // 64 or 256 input features
  float* out; int ld_out; int ncols;    // dW block (column offset applied) and its valid width
  float* bias;                          // db or nullptr
  int first_block, nslabs;              // grid mapping
};
struct WgradJobs { WgradJob j[kMaxJobs]; int n; };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Instruction descriptor: bf16 x bf16 -> f32, M=128, both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int N) {
  return make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradJobs jobs, int64_t M) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kWgOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kWgStages + s); };
  const uint32_t bar_acc = bar0 + 8u * (2 * kWgStages);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + kWgOffBar + 64);

  // job / slab of this CTA
  int ji = 0;
  for (int j = 1; j < jobs.n; ++j)
    if ((int)blockIdx.x >= jobs.j[j].first_block) ji = j;
  const WgradJob& job = jobs.j[ji];
  const int slab = blockIdx.x - job.first_block;
  const int64_t total_chunks = (M + kWgChunkRows - 1) / kWgChunkRows;
  const int64_t c_beg = total_chunks * slab / job.nslabs, c_end = total_chunks * (slab + 1) / job.nslabs;
  const int nchunks = (int)(c_end - c_beg);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full(s), 32 * kWgLoaderWarps); mbar_init(bar_empty(s), 1 + kWgBiasWarps); }
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == kWgLoaderWarps) { tmem_alloc(sbase + kWgOffBar + 64, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kWgLoaderWarps) {
    // ================= loaders =================
    const int tid = threadIdx.x;                            // 0..127
    const int a_cpr = 16 * job.mh, b_cpr = job.n >> 3;      // 16-byte chunks per row
    for (int c = 0; c < nchunks + kWgStages - 1; ++c) {
      if (c < nchunks) {
        const int s = c % kWgStages;
        if (c >= kWgStages) mbar_wait(bar_empty(s), ((c / kWgStages) - 1) & 1, 600 + s);
        const int64_t row0 = (c_beg + c) * kWgChunkRows;
        const uint32_t sa = sbase + s * kWgStageBytes, sb = sa + kWgOffB;
        for (int idx = tid; idx < kWgChunkRows * a_cpr; idx += 32 * kWgLoaderWarps) {
          const int r = idx / a_cpr, cc = idx % a_cpr;
          const int64_t row = row0 + r;
          const bool ok = row < M;
          // MN-major SW128: feature block (64 feats) -> [row][128 B], 16-byte chunk XOR (row & 7)
          cp_async16(sa + (cc >> 3) * (kWgChunkRows * 128) + r * 128 + (((cc & 7) ^ (r & 7)) << 4),
                     job.A + (ok ? row : 0) * job.lda + cc * 8, ok ? 16u : 0u);
        }
        for (int idx = tid; idx < kWgChunkRows * b_cpr; idx += 32 * kWgLoaderWarps) {
          const int r = idx / b_cpr, cc = idx % b_cpr;
          const int64_t row = row0 + r;
          const bool ok = row < M;
          cp_async16(sb + (cc >> 3) * (kWgChunkRows * 128) + r * 128 + (((cc & 7) ^ (r & 7)) << 4),
                     job.B + (ok ? row : 0) * job.ldb + cc * 8, ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      const int done = c - (kWgStages - 1);                 // the group issued kWgStages-1 iterations ago has landed
      if (done >= 0) {
        cp_async_wait<kWgStages - 1>();
        fence_proxy_async();                                // generic-proxy writes -> visible to the tensor core
        mbar_arrive(bar_full(done % kWgStages));
      }
    }
  } else if (warp == kWgLoaderWarps) {
    // ================= MMA issuer =================
    constexpr uint32_t kHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO = 1024 (8-row group), SW128
    const uint32_t lbo = ((uint32_t)(kWgChunkRows * 128) >> 4) << 16;          // LBO = next 64-feature block
    const uint32_t idesc = make_idesc_mn(job.n);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % kWgStages;
      mbar_wait(bar_full(s), (c / kWgStages) & 1, 700 + s);
      tc_fence_after();
      const uint32_t sa = sbase + s * kWgStageBytes, sb = sa + kWgOffB;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kWgChunkRows / 16; ++ks) {
          const uint32_t b_lo = ((sb + ks * 2048) >> 4) | lbo;
          for (int h = 0; h < job.mh; ++h) {
            const uint32_t a_lo = ((sa + h * (2 * kWgChunkRows * 128) + ks * 2048) >> 4) | lbo;
            mma_bf16_ss(tmem_base + h * 256, ((uint64_t)kHiMn << 32) | a_lo, ((uint64_t)kHiMn << 32) | b_lo, idesc,
                        (c == 0 && ks == 0) ? 0u : 1u);
          }
        }
        tc_commit(bar_empty(s));
        if (c == nchunks - 1) tc_commit(bar_acc);
      }
      __syncwarp();
    }
  } else {
    // ================= bias column sums, then the epilogue =================
    const int bw = warp - kWgLoaderWarps - 1;              // 0..3
    const int b = bw * 32 + lane;                          // 0..127 -> dY features 2b, 2b+1
    const bool bias_on = job.bias != nullptr && (2 * b) < 128 * job.mh;
    float s0 = 0.f, s1 = 0.f;
    const int f = 2 * b, fb = f >> 6, ch = (f & 63) >> 3, e = f & 7;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % kWgStages;
      mbar_wait(bar_full(s), (c / kWgStages) & 1, 800 + s);
      if (bias_on) {
        const uint8_t* sa = smem + s * kWgStageBytes + fb * (kWgChunkRows * 128) + e * 2;
#pragma unroll 8
        for (int r = 0; r < kWgChunkRows; ++r) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(sa + r * 128 + ((ch ^ (r & 7)) << 4));
          s0 += __uint_as_float(w << 16);
          s1 += __uint_as_float(w & 0xFFFF0000u);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty(s));
    }
    if (bias_on && nchunks > 0) {
      atomicAdd(job.bias + f, s0);
      atomicAdd(job.bias + f + 1, s1);
    }
    if (nchunks > 0) {
      mbar_wait(bar_acc, 0, 900);
      tc_fence_after();
      const int q = warp & 3;                              // TMEM lane quadrant
      for (int h = 0; h < job.mh; ++h) {
        const int mrow = h * 128 + q * 32 + lane;          // output feature
        float* orow = job.out + (int64_t)mrow * job.ld_out;
        for (int c0 = 0; c0 < job.n; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * 256 + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < job.ncols) atomicAdd(orow + c0 + j, __uint_as_float(r[j]));
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == kWgLoaderWarps) tmem_dealloc(tmem_base, 512);
}

// ---- small CUDA-core pieces ----------------------------------------------------------------------
// rgb_linear (dW[3,128], db[3]) and sigma_linear (dW[1,256], db[1]): thread = input feature
__global__ void __launch_bounds__(256) heads_wgrad_kernel(const float* __restrict__ d_raw, const __nv_bfloat16* __restrict__ hv,
                                                         const __nv_bfloat16* __restrict__ h7, int64_t M, int64_t rows_per_block,
                                                         float* __restrict__ grads) {
  const int t = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float ar = 0.f, ag = 0.f, ab = 0.f, as = 0.f, br = 0.f, bg = 0.f, bb = 0.f, bs = 0.f;
  for (int64_t row = r0; row < r1; ++row) {
    const float4 d = __ldg(reinterpret_cast<const float4*>(d_raw) + row);
    as = fmaf(d.w, __bfloat162float(h7[row * 256 + t]), as);
    if (t < 128) {
      const float h = __bfloat162float(hv[row * 128 + t]);
      ar = fmaf(d.x, h, ar); ag = fmaf(d.y, h, ag); ab = fmaf(d.z, h, ab);
    }
    if (t == 0) { br += d.x; bg += d.y; bb += d.z; bs += d.w; }
  }
  atomicAdd(grads + w_off(L_SIGMA) + t, as);
  if (t < 128) {
    atomicAdd(grads + w_off(L_RGB) + t, ar);
    atomicAdd(grads + w_off(L_RGB) + 128 + t, ag);
    atomicAdd(grads + w_off(L_RGB) + 256 + t, ab);
  }
  if (t == 0) {
    atomicAdd(grads + b_off(L_RGB), br); atomicAdd(grads + b_off(L_RGB) + 1, bg); atomicAdd(grads + b_off(L_RGB) + 2, bb);
    atomicAdd(grads + b_off(L_SIGMA), bs);
  }
}

// view_linear: the 27 direction columns and the bias.  d_hv_pre rows of one direction (ray) are
// summed first, then dW[n][256+j] += g[n]*de[j], db[n] += g[n].   thread = output feature n
__global__ void __launch_bounds__(128) view_dirs_wgrad_kernel(const __nv_bfloat16* __restrict__ dhv, const float* __restrict__ de,
                                                             int64_t M, int rows_per_dir, int64_t dirs_per_block,
                                                             float* __restrict__ grads) {
  const int n = threadIdx.x;
  const int64_t ndirs = (M + rows_per_dir - 1) / rows_per_dir;
  const int64_t v0 = (int64_t)blockIdx.x * dirs_per_block, v1 = min(ndirs, v0 + dirs_per_block);
  float acc[27], accb = 0.f;
#pragma unroll
  for (int j = 0; j < 27; ++j) acc[j] = 0.f;
  for (int64_t v = v0; v < v1; ++v) {
    float g = 0.f;
    const int64_t r0 = v * rows_per_dir, r1 = min(M, r0 + rows_per_dir);
    for (int64_t row = r0; row < r1; ++row) g += __bfloat162float(dhv[row * 128 + n]);
    accb += g;
    const float* d = de + v * 32;
#pragma unroll
    for (int j = 0; j < 27; ++j) acc[j] = fmaf(g, __ldg(d + j), acc[j]);
  }
  float* w = grads + w_off(L_VIEW) + (int64_t)n * 283 + 256;
#pragma unroll
  for (int j = 0; j < 27; ++j) atomicAdd(w + j, acc[j]);
  atomicAdd(grads + b_off(L_VIEW) + n, accb);
}

int mlp_tc_wgrad(const void* ws, const WsLayout& L, const float* d_raw, int64_t M, int rows_per_dir, float* grads,
                 cudaStream_t st) {
  const uint8_t* b = (const uint8_t*)ws;
  const __nv_bfloat16* act = (const __nv_bfloat16*)(b + L.act);
  const __nv_bfloat16* hv = (const __nv_bfloat16*)(b + L.hv);
  const __nv_bfloat16* xenc = (const __nv_bfloat16*)(b + L.xenc);
  const __nv_bfloat16* dpre = (const __nv_bfloat16*)(b + L.dpre);
  const __nv_bfloat16* dhv = (const __nv_bfloat16*)(b + L.dhv);
  const float* de = (const float*)(b + L.de);
  auto ACT = [&](int l) { return act + (int64_t)l * M * 256; };     // h_l (l = 8: bottleneck)
  auto DPRE = [&](int l) { return dpre + (int64_t)l * M * 256; };   // d(pre-activation) of layer l (8: d_bottleneck)

  // ---- tensor-core jobs ----
  WgradJobs jb{};
  WgradJob* jobs = jb.j;
  int nj = 0;
  auto add = [&](const __nv_bfloat16* A, int lda, int mh, const __nv_bfloat16* B, int ldb, int n, int layer, int col0,
                 int ncols, bool bias, int weight) {
    WgradJob& j = jobs[nj++];
    j.A = A; j.lda = lda; j.mh = mh; j.B = B; j.ldb = ldb; j.n = n;
    j.out = grads + w_off(layer) + col0; j.ld_out = kIn[layer]; j.ncols = ncols;
    j.bias = bias ? grads + b_off(layer) : nullptr;
    j.nslabs = weight;                                               // relative cost, turned into slabs below
  };
  add(DPRE(0), 256, 2, xenc, 64, 64, 0, 0, 63, true, 5);                                   // layer 0: X = x_enc
  for (int l = 1; l <= 7; ++l) add(DPRE(l), 256, 2, ACT(l - 1), 256, 256, l, l == 5 ? 63 : 0, 256, true, 8);
  add(DPRE(5), 256, 2, xenc, 64, 64, 5, 0, 63, false, 5);                                  // layer 5, x part of [x,h]
  add(DPRE(8), 256, 2, ACT(7), 256, 256, L_BOTT, 0, 256, true, 8);                         // bottleneck_linear
  add(dhv, 128, 1, ACT(8), 256, 256, L_VIEW, 0, 256, false, 6);                            // view_linear, bottleneck columns
  static int sm_count = 0;
  static bool attr_done = false;
  if (!attr_done) {
    int dev = 0;
    NERF_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    NERF_CUDA(cudaGetDeviceProperties(&p, dev));
    sm_count = p.multiProcessorCount;
    NERF_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    attr_done = true;
  }
  int wsum = 0;
  for (int j = 0; j < nj; ++j) wsum += jobs[j].nslabs;
  const int64_t total_chunks = (M + kWgChunkRows - 1) / kWgChunkRows;
  int nblocks = 0;
  for (int j = 0; j < nj; ++j) {
    int64_t s = (int64_t)jobs[j].nslabs * sm_count / wsum;
    if (s < 1) s = 1;
    if (s > total_chunks) s = total_chunks;
    jobs[j].first_block = nblocks;
    jobs[j].nslabs = (int)s;
    nblocks += (int)s;
  }
  jb.n = nj;
  wgrad_tc_kernel<<<nblocks, kWgThreads, kWgSmemBytes, st>>>(jb, M);
  NERF_LAUNCH_CHECK("wgrad_tc_kernel");

  // ---- CUDA-core pieces ----
  {
    const int64_t rpb = 512;
    heads_wgrad_kernel<<<ceil_div(M, rpb), 256, 0, st>>>(d_raw, hv, ACT(7), M, rpb, grads);
    NERF_LAUNCH_CHECK("heads_wgrad_kernel");
    const int64_t ndirs = (M + rows_per_dir - 1) / rows_per_dir;
    int64_t dpb = (ndirs + 4 * sm_count - 1) / (4 * sm_count);
    if (dpb < 1) dpb = 1;
    view_dirs_wgrad_kernel<<<ceil_div(ndirs, dpb), 128, 0, st>>>(dhv, de, M, rows_per_dir, dpb, grads);
    NERF_LAUNCH_CHECK("view_dirs_wgrad_kernel");
  }
  return 0;
}

}  // namespace nerf
