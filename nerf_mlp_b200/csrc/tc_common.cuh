// Definitions shared by the tensor-core MLP kernels (forward, dgrad chain, wgrad): swizzled
// shared-memory layouts, the weight-slot schedules and the bf16-mode workspace layout.
#pragma once
#include "nerf_common.cuh"
#include "tc_ptx.cuh"

#ifndef NERF_TC_NK
#define NERF_TC_NK 2          // K-steps (16 wide) per weight slot
#endif
#ifndef NERF_TC_RING
#define NERF_TC_RING 4        // ring slots
#endif
#ifndef NERF_TC_SKEW
#define NERF_TC_SKEW 1        // slots by which tile B's issuer starts behind tile A's
#endif

namespace nerf {

constexpr int kNK = NERF_TC_NK;
constexpr int kRing = NERF_TC_RING;
constexpr int kSkew = NERF_TC_SKEW;
static_assert(kNK == 1 || kNK == 2 || kNK == 4, "slot = 1, 2 or 4 K-steps");
static_assert(kSkew >= 1 && kSkew + 1 < 2 * kRing, "ring must hold the skew plus at least one prefetch slot");
constexpr int kSlotBytes = 8192 * kNK;            // 256 rows x 32 B x NK
constexpr int kTileM = 128;
#ifndef NERF_TC_CTAS
#define NERF_TC_CTAS 2        // 2 = CTA pairs (cta_group::2): B operand split across two SMs
#endif
constexpr int kCtasDefault = NERF_TC_CTAS;

// ---- swizzled K-major element offsets (bytes): 16-byte chunk index XOR row bits, as applied by
// TMA / UMMA for SWIZZLE_{32,64,128}B ------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t sw128_off(int row, int k) {   // rows of 128 B (64 bf16)
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ row) & 7) << 4) + (k & 7) * 2);
}
__host__ __device__ __forceinline__ uint32_t sw64_off(int row, int k) {    // rows of 64 B (32 bf16)
  return (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((((k >> 3) ^ (row >> 1)) & 3) << 4) + (k & 7) * 2);
}
__host__ __device__ __forceinline__ uint32_t sw32_off(int row, int k) {    // rows of 32 B (16 bf16)
  return (uint32_t)((row >> 3) * 256 + (row & 7) * 32 + ((((k >> 3) ^ (row >> 2)) & 1) << 4) + (k & 7) * 2);
}

// ---- weight-slot schedule ----------------------------------------------------------------------
enum AKind : uint32_t { A_X = 0, A_ACT = 1, A_ONES = 2 };
struct Slot {            // consumed by the kernels: one 16-byte constant-bank load per slot
  uint32_t goff, bytes;
  uint32_t a_add;        // (byte offset of the first K-step inside the A tile) >> 4
  uint32_t flags;        // bits 0-1 a_kind | 2 nk==2 | 3 first | 4 last | 5 N==128 | 6 nk==4
};
constexpr uint32_t kFlagNk2 = 4, kFlagFirst = 8, kFlagLast = 16, kFlagN128 = 32, kFlagNk4 = 64;
struct PackSlot {        // consumed by the pack kernel: value(n,kk) = params[w_base + n*n_stride + kk*k_stride]
  uint32_t goff;
  int32_t w_base, n_stride, k_stride, kvalid, b_off, n, nk, sw, is_bias;
};
constexpr int kMaxSlots = 192;

// ---- bf16-mode workspace (byte offsets) -----------------------------------------------------------
// Saved activations / gradients live in HBM as SHARED-MEMORY TILE IMAGES: per 128-row tile and
// per 64-feature block a contiguous 16 KB block [128 rows][128 B] with the 16-byte chunks XOR-
// swizzled by (row & 7) -- byte for byte what the kernels hold in shared memory.  A save is then a
// per-warp coalesced 4 KB block copy (LDS.128 -> STG.128), and the wgrad kernel's operand loads are plain
// bulk-async copies (cp.async.bulk, SASS UBLKCP -- contiguous blocks, no tensor map needed).
//   element (row r, feature f) of a tensor with F features (F/64 blocks per tile):
//     tile = r / 128, rt = r % 128, fb = f / 64
//     byte = tile * (F/64) * 16384 + fb * 16384 + rt * 128 + ((((f % 64) / 8) ^ (rt & 7)) * 16) + (f % 8) * 2
// forward:  vb    fp32 [nvb][128]    view bias (one row per ray, or per sample for the encoded entry)
//           de    fp32 [nvb][32]     encoded view direction (27 used)                     (save only)
//           act   img  [9][Mp x 256] h0..h7 (post-ReLU), bottleneck                       (save only)
//           hv    img  [Mp x 128]    view-layer output (post-ReLU)                        (save only)
//           xenc  img  [Mp x 64]     encoded position (63 used)                           (save only)
//           de16  img  [Mp x 64]     encoded view direction per sample (27 used)          (save only)
//           mask  u32  [8][M][8]     ReLU masks of h0..h7; hvmask u32 [M][4]   (row-major, save only)
// backward: dpre  img  [9][Mp x 256] d(pre-activation) of layers 0..7, d(bottleneck); dhv img [Mp x 128]
//           flags u32  [10][Mp/128]   fused backward: number of warps that have published (d_pre_0..7, d_bott, d_hv) of a tile
//                      + 32 words (unit counter) + 128 words (wgrad roles' progress, read by the unit throttle)
//           -- Mp = rows padded to a multiple of 512 (whole tile quads per CTA pair)
struct WsLayout {
  size_t vb, de, act, hv, xenc, de16, mask, hvmask, dpre, dhv, flags, total;
  int64_t Mp;
};
inline WsLayout ws_layout(int64_t M, int save) {
  WsLayout w{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 1023) & ~(size_t)1023; return r; };
  w.Mp = (M + 4 * kTileM - 1) / (4 * kTileM) * (4 * kTileM);      // whole tile pairs per CTA, whole quads per CTA pair
  w.vb = take((size_t)M * 128 * 4);
  if (save) {
    w.de = take((size_t)M * 32 * 4);
    w.act = take((size_t)9 * w.Mp * 256 * 2);
    w.hv = take((size_t)w.Mp * 128 * 2);
    w.xenc = take((size_t)w.Mp * 64 * 2);
    w.de16 = take((size_t)w.Mp * 64 * 2);
    w.mask = take((size_t)8 * M * 8 * 4);
    w.hvmask = take((size_t)M * 4 * 4);
    w.dpre = take((size_t)9 * w.Mp * 256 * 2);
    w.dhv = take((size_t)w.Mp * 128 * 2);
    w.flags = take(((size_t)10 * (w.Mp / kTileM) + 160) * 4);  // fused backward: publish counters [10 tensors][tiles] + unit counter (32 words) + wgrad progress per SM pair (128 words)
  }
  w.total = o;
  return w;
}

int mlp_tc_wgrad(const void* ws, const WsLayout& L, const float* d_raw, int64_t M, int rows_per_dir, float* grads,
                 cudaStream_t st);
int mlp_tc_heads_fork(cudaStream_t st);
int mlp_tc_heads_wgrad(const void* ws, const WsLayout& L, const float* d_raw, int64_t M, float* grads, cudaStream_t st);

}  // namespace nerf
