// Weight / bias gradients of the bf16 tensor-core backward (the autograd of reference
// model.py:57-81 w.r.t. its 24 parameters, triggered at scripts/train.py:382).
//
//   dW_l[out,in] = sum_rows dY_l[row,out] * X_l[row,in]      (X_l, dY_l: bf16 [M, features] in HBM,
//                                                             written by the forward / dgrad kernels)
//
// is a GEMM whose contraction runs over the sample rows, so BOTH operands are "MN-major" for the
// tensor core (features contiguous, rows strided).  One CTA owns one (layer, row-slab) job:
//   loader   (1 lane)   TMA bulk copies of a 64-row chunk of dY and X into a 3-stage ring: the
//                       operands are stored in HBM as shared-memory tile images (tc_common.cuh), i.e.
//                       already in the 128B-swizzled MN-major layout UMMA expects
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
#ifndef NERF_WG_PREFETCH
#define NERF_WG_PREFETCH 0
#endif
constexpr int kWgPrefetch = NERF_WG_PREFETCH;               // L2 prefetch distance in chunks (0 = off)
constexpr int kWgStageBytes = 65536;                       // A: 64 x 256 bf16 (32 KB) + B: 64 x 256 bf16 (32 KB)
constexpr int kWgOffB = 32768;
constexpr int kWgLoaderWarps = 1, kWgBiasWarps = 4;
constexpr int kWgThreads = 32 * (kWgLoaderWarps + 1 + kWgBiasWarps);
constexpr int kWgOffBar = kWgStages * kWgStageBytes;
constexpr int kWgSmemBytes = kWgOffBar + 128;
constexpr int kMaxJobs = 16;

struct WgradJob {
  const uint8_t* A; int a_tile_bytes;   // dY tile images; feature blocks [0, 2*mh) are used
  const uint8_t* B; int b_tile_bytes;   // X  tile images; feature blocks [0, n/64) are used
  int mh;                               // M halves of 128 output features (1 or 2)
  int n;                                // 64, 128 or 256 input features
  float* out; int ld_out; int ncols;    // dW block (column offset applied) and its valid width
  float* bias;                          // db or nullptr
  int first_block, nslabs;              // grid mapping
};
struct WgradJobs { WgradJob j[kMaxJobs]; int n; };

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Instruction descriptor: bf16 x bf16 -> f32, M=128, both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int N) {
  return make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradJobs jobs, int64_t Mp) {
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
  const int64_t total_chunks = Mp / kWgChunkRows;
  // Slab s takes the 64-row chunks  total-1-s, total-1-s-nslabs, ...  (interleaved, descending): the
  // dgrad kernel wrote the highest tiles last, so every CTA starts on dY data that is still in L2.
  const int nchunks = slab < total_chunks ? (int)((total_chunks - 1 - slab) / job.nslabs) + 1 : 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1 + kWgBiasWarps); }
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == kWgLoaderWarps) { tmem_alloc(sbase + kWgOffBar + 64, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kWgLoaderWarps) {
    // ================= loader: one lane, TMA bulk copies of tile-image blocks =================
    if (lane == 0) {
      const int a_fb = 2 * job.mh, b_fb = job.n >> 6;
      const uint32_t bytes = (uint32_t)(a_fb + b_fb) * 8192u;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % kWgStages;
        if (c >= kWgStages) mbar_wait(bar_empty(s), ((c / kWgStages) - 1) & 1, 600 + s);
        const int64_t chunk = total_chunks - 1 - slab - (int64_t)c * job.nslabs;
        const int64_t tile = chunk >> 1;
        const uint32_t half = (uint32_t)(chunk & 1) * 8192u;      // rows 0-63 / 64-127 of the tile
        const uint32_t sa = sbase + s * kWgStageBytes, sb = sa + kWgOffB;
        mbar_expect_tx(bar_full(s), bytes);
        for (int fb = 0; fb < a_fb; ++fb)
          bulk_g2s(sa + fb * 8192, job.A + tile * job.a_tile_bytes + fb * 16384 + half, 8192, bar_full(s));
        for (int fb = 0; fb < b_fb; ++fb)
          bulk_g2s(sb + fb * 8192, job.B + tile * job.b_tile_bytes + fb * 16384 + half, 8192, bar_full(s));
        // Optional L2 prefetch kWgPrefetch chunks ahead of the ring (-DNERF_WG_PREFETCH=n).  Measured on the
        // 196 608-row step: 0.40 ms without, 0.44 / 0.52 / 0.67 ms at distance 3 / 6 / 12 -- the kernel is at the
        // HBM read rate this access pattern reaches (5.3 TB/s), not latency-bound, so it stays off.
        if (kWgPrefetch > 0 && c + kWgPrefetch < nchunks) {
          const int64_t pchunk = total_chunks - 1 - slab - (int64_t)(c + kWgPrefetch) * job.nslabs;
          const int64_t ptile = pchunk >> 1;
          const uint32_t phalf = (uint32_t)(pchunk & 1) * 8192u;
          for (int fb = 0; fb < a_fb; ++fb) bulk_prefetch_l2(job.A + ptile * job.a_tile_bytes + fb * 16384 + phalf, 8192);
          for (int fb = 0; fb < b_fb; ++fb) bulk_prefetch_l2(job.B + ptile * job.b_tile_bytes + fb * 16384 + phalf, 8192);
        }
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
        // rows of dW that start 16-byte aligned with a multiple-of-4 width take red.global.add.v4.f32
        // (a quarter of the L2 reduction operations of the split-K tail)
        const bool vec = ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) && (job.ncols & 3) == 0;
        for (int c0 = 0; c0 < job.n; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * 256 + c0, r);
          tmem_ld_wait();
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (c0 + j < job.ncols)
                red_add_v4(orow + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                           __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < job.ncols) atomicAdd(orow + c0 + j, __uint_as_float(r[j]));
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == kWgLoaderWarps) tmem_dealloc(tmem_base, 512);
}

// ---- small CUDA-core piece: rgb_linear (dW[3,128], db[3]) and sigma_linear (dW[1,256], db[1]) ----
// Reads the hv / h7 tile images with 16 bytes per lane: a warp covers one 512-byte h7 row (or two
// 256-byte hv rows) per load instruction; lane l owns logical 16-byte chunk l of the row.
constexpr int kHeadsRowsPerWarp = 64, kHeadsWarps = 8;
__global__ void __launch_bounds__(32 * kHeadsWarps) heads_wgrad_kernel(const float* __restrict__ d_raw,
                                                                       const uint8_t* __restrict__ hv_img,
                                                                       const uint8_t* __restrict__ h7_img, int64_t M,
                                                                       float* __restrict__ grads) {
  __shared__ float red[256 + 384 + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 644; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int64_t r0 = ((int64_t)blockIdx.x * kHeadsWarps + warp) * kHeadsRowsPerWarp;
  float as[8], av[3][8], bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) { as[j] = 0.f; av[0][j] = av[1][j] = av[2][j] = 0.f; }
  // sigma: h7 rows, lane = chunk (fb = lane>>3, cl = lane&7)
  for (int i = 0; i < kHeadsRowsPerWarp; i += 4) {
    uint4 x[4]; float4 d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t row = r0 + i + u;
      const bool ok = row < M;
      const int rt = (int)(row & 127);
      d[u] = ok ? __ldg(reinterpret_cast<const float4*>(d_raw) + row) : make_float4(0.f, 0.f, 0.f, 0.f);
      x[u] = ok ? __ldg(reinterpret_cast<const uint4*>(h7_img + (row >> 7) * 65536 + (lane >> 3) * 16384 + rt * 128 +
                                                       (((lane & 7) ^ (rt & 7)) << 4)))
                : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        as[2 * j] = fmaf(d[u].w, __uint_as_float(w[j] << 16), as[2 * j]);
        as[2 * j + 1] = fmaf(d[u].w, __uint_as_float(w[j] & 0xFFFF0000u), as[2 * j + 1]);
      }
      bsum[0] += d[u].x; bsum[1] += d[u].y; bsum[2] += d[u].z; bsum[3] += d[u].w;
    }
  }
  // rgb: hv rows, two rows per instruction: lanes 0-15 -> row 2i, 16-31 -> row 2i+1; chunk = lane & 15
  const int hc = lane & 15;
  for (int i = 0; i < kHeadsRowsPerWarp; i += 8) {
    uint4 x[4]; float4 d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t row = r0 + i + 2 * u + (lane >> 4);
      const bool ok = row < M;
      const int rt = (int)(row & 127);
      d[u] = ok ? __ldg(reinterpret_cast<const float4*>(d_raw) + row) : make_float4(0.f, 0.f, 0.f, 0.f);
      x[u] = ok ? __ldg(reinterpret_cast<const uint4*>(hv_img + (row >> 7) * 32768 + (hc >> 3) * 16384 + rt * 128 +
                                                       (((hc & 7) ^ (rt & 7)) << 4)))
                : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xFFFF0000u);
        av[0][2 * j] = fmaf(d[u].x, lo, av[0][2 * j]); av[0][2 * j + 1] = fmaf(d[u].x, hi, av[0][2 * j + 1]);
        av[1][2 * j] = fmaf(d[u].y, lo, av[1][2 * j]); av[1][2 * j + 1] = fmaf(d[u].y, hi, av[1][2 * j + 1]);
        av[2][2 * j] = fmaf(d[u].z, lo, av[2][2 * j]); av[2][2 * j + 1] = fmaf(d[u].z, hi, av[2][2 * j + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&red[lane * 8 + j], as[j]);
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&red[256 + c * 128 + hc * 8 + j], av[c][j]);   // lanes l and l+16 share features
  }
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < 4; ++c) atomicAdd(&red[640 + c], bsum[c]);
  __syncthreads();
  for (int i = threadIdx.x; i < 644; i += blockDim.x) {
    float* dst = i < 256 ? grads + w_off(L_SIGMA) + i
               : i < 640 ? grads + w_off(L_RGB) + (i - 256)
               : i < 643 ? grads + b_off(L_RGB) + (i - 640)
                         : grads + b_off(L_SIGMA);
    atomicAdd(dst, red[i]);
  }
}

// Internal side stream (one per device, created on first use) for the two tiny heads: they read tensors the big
// backward kernel does not (hv, and h7 a second time) and fill its ramp-up / tail instead of adding 40 us after it.
// Fork and join are event dependencies (capturable in a CUDA graph, no host synchronisation).
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static int side_stream(SideStream** out) {
  static SideStream side[kMaxDevices];
  DeviceProps dp;
  int rc = current_device(&dp);
  if (rc) return rc;
  NERF_CHECK_ARG(dp.ordinal >= 0 && dp.ordinal < kMaxDevices, "device ordinal %d out of range", dp.ordinal);
  SideStream& x = side[dp.ordinal];
  if (x.s == nullptr) {
    NERF_CUDA(cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking));
    NERF_CUDA(cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming));
    NERF_CUDA(cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming));
  }
  *out = &x;
  return 0;
}

// rgb / sigma head gradients beside the fused backward kernel: call fork BEFORE launching the big kernel on `st`
// (the heads only depend on what precedes it) and launch_join AFTER (the big kernel's CTAs take the SMs first; the
// heads' blocks run wherever one of them has finished).
int mlp_tc_heads_fork(cudaStream_t st) {
  SideStream* x;
  int rc = side_stream(&x);
  if (rc) return rc;
  NERF_CUDA(cudaEventRecord(x->fork, st));
  NERF_CUDA(cudaStreamWaitEvent(x->s, x->fork, 0));
  return 0;
}
int mlp_tc_heads_wgrad(const void* ws, const WsLayout& L, const float* d_raw, int64_t M, float* grads, cudaStream_t st) {
  SideStream* x;
  int rc = side_stream(&x);
  if (rc) return rc;
  const uint8_t* b = (const uint8_t*)ws;
  const int64_t ntiles = L.Mp / kTileM;
  heads_wgrad_kernel<<<ceil_div(M, (int64_t)kHeadsWarps * kHeadsRowsPerWarp), 32 * kHeadsWarps, 0, x->s>>>(
      d_raw, b + L.hv, b + L.act + (int64_t)7 * ntiles * 65536, M, grads);
  NERF_LAUNCH_CHECK("heads_wgrad_kernel");
  NERF_CUDA(cudaEventRecord(x->join, x->s));
  NERF_CUDA(cudaStreamWaitEvent(st, x->join, 0));
  return 0;
}

int mlp_tc_wgrad(const void* ws, const WsLayout& L, const float* d_raw, int64_t M, int rows_per_dir, float* grads,
                 cudaStream_t st) {
  (void)rows_per_dir;   // the per-sample direction encodings were saved by the forward (de16)
  const uint8_t* b = (const uint8_t*)ws;
  const int64_t ntiles = L.Mp / kTileM;
  auto ACT = [&](int l) { return b + L.act + (int64_t)l * ntiles * 65536; };     // h_l (l = 8: bottleneck)
  auto DPRE = [&](int l) { return b + L.dpre + (int64_t)l * ntiles * 65536; };   // d(pre-act) of layer l (8: d_bottleneck)
  const uint8_t* hv = b + L.hv;
  const uint8_t* xenc = b + L.xenc;
  const uint8_t* de16 = b + L.de16;
  const uint8_t* dhv = b + L.dhv;

  // ---- tensor-core jobs ----
  WgradJobs jb{};
  WgradJob* jobs = jb.j;
  int nj = 0;
  auto add = [&](const uint8_t* A, int a_tile_bytes, int mh, const uint8_t* B, int b_tile_bytes, int n, int layer,
                 int col0, int ncols, bool bias, int weight) {
    WgradJob& j = jobs[nj++];
    j.A = A; j.a_tile_bytes = a_tile_bytes; j.mh = mh; j.B = B; j.b_tile_bytes = b_tile_bytes; j.n = n;
    j.out = grads + w_off(layer) + col0; j.ld_out = kIn[layer]; j.ncols = ncols;
    j.bias = bias ? grads + b_off(layer) : nullptr;
    j.nslabs = weight;                                               // relative cost, turned into slabs below
  };
  add(DPRE(0), 65536, 2, xenc, 16384, 64, 0, 0, 63, true, 5);                              // layer 0: X = x_enc
  for (int l = 1; l <= 7; ++l) add(DPRE(l), 65536, 2, ACT(l - 1), 65536, 256, l, l == 5 ? 63 : 0, 256, true, 8);
  add(DPRE(5), 65536, 2, xenc, 16384, 64, 5, 0, 63, false, 5);                             // layer 5, x part of [x,h]
  add(DPRE(8), 65536, 2, ACT(7), 65536, 256, L_BOTT, 0, 256, true, 8);                     // bottleneck_linear
  add(dhv, 32768, 1, ACT(8), 65536, 256, L_VIEW, 0, 256, true, 6);                         // view_linear: bottleneck columns + bias
  add(dhv, 32768, 1, de16, 16384, 64, L_VIEW, 256, 27, false, 3);                          // view_linear: direction columns
  static DeviceOnce attr_done;
  DeviceProps dp;
  int rc = current_device(&dp);
  if (rc) return rc;
  const int sm_count = dp.sm_count;
  if (attr_done.needed(dp.ordinal)) {
    NERF_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    attr_done.mark(dp.ordinal);
  }
  int wsum = 0;
  for (int j = 0; j < nj; ++j) wsum += jobs[j].nslabs;
  const int64_t total_chunks = L.Mp / kWgChunkRows;
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
  int rc2 = mlp_tc_heads_fork(st);
  if (rc2) return rc2;
  wgrad_tc_kernel<<<nblocks, kWgThreads, kWgSmemBytes, st>>>(jb, L.Mp);
  NERF_LAUNCH_CHECK("wgrad_tc_kernel");
  if ((rc2 = mlp_tc_heads_wgrad(ws, L, d_raw, M, grads, st))) return rc2;
  return 0;
}

}  // namespace nerf
