// Fused NeRF MLP on the 5th-gen tensor cores (tcgen05, sm_100a): forward and the dgrad chain of
// the backward.  (The weight gradients are in nerf_mlp_wgrad.cu.)
//
// One persistent CTA per SM processes PAIRS of 128-row sample tiles.  Per tile the whole network
// (reference model.py:57-81) runs on-chip:
//
//   prologue   points o+d*z, positional encoding (L=10) in registers -> bf16 x-tile in SMEM
//   10 GEMMs   tcgen05.mma 128xNx16 (bf16 in, fp32 accumulate in TMEM); A = activation tile in SMEM
//              (128B-swizzled, K-major), B = weight slots streamed from L2 by bulk-async (TMA) copies
//              into a ring; bias enters as one extra K-step against a constant "ones" tile (hi+lo
//              bf16 split), so the epilogue is a pure convert
//   epilogue   tcgen05.ld -> ReLU -> bf16 -> SMEM (next layer's A operand); sigma head (256->1) and
//              rgb head (128->3) on CUDA cores from the un-rounded fp32 accumulators; the view-
//              direction part of view_linear is a per-ray fp32 vector ("view bias") added here.
//
// The backward dgrad chain (d_raw -> d(pre-activation) of every layer) is the same machine run
// over the transposed weight image: 9 GEMMs per tile, epilogue = ReLU mask (bit masks saved by the
// forward) -> bf16 -> SMEM + HBM (operands of the wgrad kernel).
//
// Warp roles: warp 0 = weight producer (one lane), warps 1 / 2 = MMA issuers of tile A / tile B
// (one lane each; a single issuing thread was measured to be the bottleneck), warps 3-6 / 7-10 =
// prologue + epilogue of tile A / tile B.  The two tiles share every weight slot (a slot is freed
// when both issuers have committed it), which halves the L2->SMEM weight traffic; tile B starts
// kSkew slots behind tile A so that each tile's epilogue overlaps the other tile's MMAs.
// TMEM: 2 x 256 fp32 columns (all 512).
//
// kCtas = 2 (cta_group::2): two CTAs on the SMs of one TPC form a pair.  Each keeps its own two
// 128-row tiles and HALF of every weight slot (B is split along N across the pair), so the weight
// bytes per SM and the B-operand shared-memory reads halve and the ring is twice as deep.  The
// leader CTA's issuers issue M = 256 MMAs for both CTAs; commits are multicast to both CTAs'
// barriers; the peer's epilogue threads arrive remotely on the leader's "activations ready"
// barrier and a relay warp forwards the peer's "slot full" events.
#include "tc_common.cuh"
#include <mutex>
#include <stdlib.h>
#include <vector>

namespace nerf {
using namespace ptx;

// Measured and dropped: TMA bulk stores for the tile-image saves (bulk S2G 0.34 vs 0.31 ms on the 196 608-row save
// pass: they queue behind the same SM's bulk loads).  Dedicated LSU copy warps (below) measure equal to the epilogue
// warps copying their own blocks (0.306 vs 0.316 / 0.245 vs 0.242 ms) and are kept because they let all eight
// compute warps share every epilogue; NERF_TC_COPY_WARPS=0 restores the round-1 arrangement.
constexpr int kThreads = 352;                     // producer, 2 mma issuers, 8 prologue/epilogue warps
// Kernels that write tile images to HBM (save-mode forward, dgrad chain) get 4 more warps that do nothing but the
// copy-out: the tile in shared memory is byte for byte its HBM image, so a copy warp moves one whole 16 KB feature
// block with coalesced LDS.128 / STG.128 while the eight compute warps -- which then share every epilogue exactly as in
// the inference kernel -- are already waiting for the next accumulator.  (Round 1 had the epilogue warps copy their
// own blocks out: ~2 000 clk per tile-layer on the dependent chain MMA -> epilogue -> MMA.)
#ifndef NERF_TC_COPY_WARPS
#define NERF_TC_COPY_WARPS 4
#endif
constexpr int kCopyWarps = NERF_TC_COPY_WARPS;
constexpr int kThreadsStore = kThreads + 32 * kCopyWarps;
template <bool kStores> constexpr int tc_threads() { return (kStores && kCopyWarps > 0) ? kThreadsStore : kThreads; }
#ifndef NERF_TC_ARRIVE_LANES
#define NERF_TC_ARRIVE_LANES 1                    // 1 = every thread arrives on "activations ready" (default); 32 = one arrive per warp (measured: no gain, 78.7 vs 79.4 % of peak)
#endif
constexpr int kArriveLanes = NERF_TC_ARRIVE_LANES;
// Separate weight streams per tile ("split" mode).  With ONE shared stream both tiles must consume every slot
// within ring distance of each other, which locks their MMA phases together: in the save / dgrad kernels (each
// tile's epilogue on its own four warps) the SM then alternates between "both tiles issue MMAs" and "both
// tiles run their epilogue" and the tensor pipe idles through every epilogue (cycle counters: 6.2k clk of MMAs,
// then 5k clk of commit + epilogue + arrive per layer).  With a private half-ring per tile the two chains are
// decoupled and tile B is started half a period behind tile A, so one tile's MMAs run under the other tile's
// epilogue.  Costs 2x the L2->SMEM weight bytes (16 B/clk/SM) and half the ring depth per stream.
#ifndef NERF_TC_SPLIT_TRAIN
#define NERF_TC_SPLIT_TRAIN 0                     // save-mode forward + dgrad kernels (measured slower: 0.322 / 0.255 vs 0.311 / 0.243 ms)
#endif
// Forward kernels: the x-tile (positional encoding) of the NEXT work unit is computed while the compute warps wait
// for the layer-6 accumulators of the current one (the x-tile's last reader is layer 5), so the unit boundary costs
// one ordinary layer transition instead of  drain + last epilogue + prologue  (timeline: 6.6k clk of idle tensor
// pipe per unit of 62k in the inference kernel, 13k of ~100k in the save-mode kernel).
// Pair mode: the peer's "my half of slot s landed" relay arrives directly on the leader's full(s) barrier (count 2)
// instead of a second barrier, so the issue loop makes one barrier wait per slot instead of two.
#ifndef NERF_TC_RELAY_FULL
#define NERF_TC_RELAY_FULL 1
#endif
// The issuer warps' barrier polls are made by lane 0 alone (then __syncwarp) instead of by all 32 lanes.
#ifndef NERF_TC_POLL_LANE0
#define NERF_TC_POLL_LANE0 0                      // measured: 70.6 vs 84.4 % of peak -- the extra __syncwarp per slot costs far more than the 31 polls
#endif
#ifndef NERF_TC_PIPE_PROLOGUE
#define NERF_TC_PIPE_PROLOGUE 1
#endif
#ifndef NERF_TC_VB_PREFETCH
#define NERF_TC_VB_PREFETCH 0                     // L1 prefetch of the view-bias row under the view layer's accumulator wait (measured: no change, 84.9 % both)
#endif
#ifndef NERF_TC_SHARED_TRAIN
#define NERF_TC_SHARED_TRAIN 0                    // 1: the save / dgrad kernels share every epilogue among all eight warps too
#endif
#ifndef NERF_TC_SPLIT_INFER
#define NERF_TC_SPLIT_INFER 0                     // inference forward (shared epilogues)
#endif
static_assert(kArriveLanes == 1 || kArriveLanes == 32, "arrive per thread or per warp");
constexpr int kFirstComputeWarp = 3;
constexpr int kNumGemmsFwd = 10, kNumGemmsBwd = 9;

// ---- shared memory map (bytes) ------------------------------------------------------------------
constexpr int kOffAct = 0;                                   // 2 x [128 x 256] bf16, SW128 K-blocks of 64
constexpr int kActBytes = 65536;
constexpr int kOffX = kOffAct + 2 * kActBytes;               // 2 x [128 x 64] bf16, SW128
constexpr int kXBytes = 16384;
constexpr int kOffRing = kOffX + 2 * kXBytes;
constexpr int kOffOnes = kOffRing + kRing * kSlotBytes;      // [8 x 16] bf16, SW32; all 16 row groups alias it (SBO = 0)
constexpr int kOnesBytes = 256;
constexpr int kOffHead = kOffOnes + kOnesBytes;              // w_sigma[256] w_rgb[3][128] b_sigma b_rgb[3]
constexpr int kHeadFloats = 256 + 384 + 4;
constexpr int kOffBar = kOffHead + kHeadFloats * 4;
constexpr int kBarSets = (NERF_TC_RELAY_FULL != 0) ? 2 : 3;   // full, empty (+ the peer-landed barriers when the relay does not use `full`)
constexpr int kNumBars = kBarSets * (2 * kRing) + 5 + 4;     // sized for the 2-CTA variant (ring twice as deep); + copy-out handshake
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 8;
static_assert(kSmemBytes <= 232448, "exceeds 227 KB of shared memory");

#ifdef NERF_DBG_TIMING
__device__ unsigned long long g_dbg_clk[8];
#define DBG_T(var) const long long var = clock64()
#define DBG_ACC(i, v) do { if (blockIdx.x == 0 && warp == kFirstComputeWarp && lane == 0) g_dbg_clk[i] += (unsigned long long)(v); } while (0)
#else
#define DBG_T(var)
#define DBG_ACC(i, v)
#endif
#ifdef NERF_DBG_TRACE
// timeline of ONE work unit of CTA 0 (development): writer 0/1 = MMA issuer of tile A/B, 2/3 = warp 3 / warp 7 lane 0
__device__ long long g_trace[4][160];
__device__ int g_trace_n[4];
#define TRACE(w, ok, tag) do { if ((ok) && blockIdx.x == 0 && g_trace_n[w] < 160) g_trace[w][g_trace_n[w]++] = (clock64() << 4) | (tag); } while (0)
#else
#define TRACE(w, ok, tag)
#endif
__constant__ Slot c_slots[kMaxSlots];          // forward schedule, then the dgrad schedule
__constant__ PackSlot c_pack[kMaxSlots];
__constant__ int c_nslots_fwd, c_nslots_bwd, c_nslots_fwd_sigma;

struct Schedule {
  std::vector<Slot> slots;
  std::vector<PackSlot> pack;
  int n_fwd = 0, n_bwd = 0;
  int n_fwd_sigma = 0;             // forward slots through layer 7 (density-only forward: sigma head, no view branch)
  size_t bytes = 0;
};

static const Schedule& schedule() {
  static Schedule s;
  static std::once_flag once;
  std::call_once(once, [] {
    // One GEMM of a schedule: `parts` of the A operand (tile kind, K length), the weight block it
    // multiplies (value(n,k) = params[w0 + n*ns + k*ks], k < kvalid), N, and an optional bias K-step.
    struct Part { int kind, klen, w0, ns, ks, kvalid; };
    struct G { int n, b_off; std::vector<Part> parts; };
    std::vector<G> fwd, bwd;
    auto W = [](int l) { return (int)w_off(l); };
    // ---- forward (reference model.py:57-81)
    fwd.push_back({256, (int)b_off(0), {{A_X, 64, W(0), 63, 1, 63}}});
    for (int l = 1; l <= 4; ++l) fwd.push_back({256, (int)b_off(l), {{A_ACT, 256, W(l), 256, 1, 256}}});
    fwd.push_back({256, (int)b_off(5), {{A_X, 64, W(5), 319, 1, 63}, {A_ACT, 256, W(5) + 63, 319, 1, 256}}});  // [x,h] :62-63
    fwd.push_back({256, (int)b_off(6), {{A_ACT, 256, W(6), 256, 1, 256}}});
    fwd.push_back({256, (int)b_off(7), {{A_ACT, 256, W(7), 256, 1, 256}}});
    fwd.push_back({256, (int)b_off(L_BOTT), {{A_ACT, 256, W(L_BOTT), 256, 1, 256}}});
    fwd.push_back({128, -1, {{A_ACT, 256, W(L_VIEW), 283, 1, 256}}});     // dirs part + bias live in the view bias
    // ---- dgrad chain: dX = dY . W, i.e. B(n = in-feature, k = out-feature) = W[k][n]
    bwd.push_back({256, -1, {{A_ACT, 128, W(L_VIEW), 1, 283, 128}}});     // d_bott   = d_hv_pre . W_view[:, :256]
    bwd.push_back({256, -1, {{A_ACT, 256, W(L_BOTT), 1, 256, 256}}});     // d_h7     = d_bott . W_bott (+ sigma term)
    for (int l = 7; l >= 1; --l)                                          // d_h{l-1} = d_pre_l . W_l[:, h part]
      bwd.push_back({256, -1, {{A_ACT, 256, W(l) + (l == 5 ? 63 : 0), 1, kIn[l], 256}}});
    uint32_t goff = 0;
    auto emit = [&](const std::vector<G>& gs) {
      int gi = 0;
      for (const G& g : gs) {
        const size_t first_idx = s.slots.size();
        for (const Part& p : g.parts) {
          for (int k0 = 0; k0 < p.klen; k0 += 16 * kNK) {
            const int nk = (p.klen - k0) / 16 < kNK ? (p.klen - k0) / 16 : kNK;
            Slot sl{};
            sl.goff = goff;
            sl.bytes = (uint32_t)g.n * 32 * nk;
            sl.a_add = (uint32_t)((k0 >> 6) * 16384 + ((k0 & 63) >> 4) * 32) >> 4;
            if (nk == 3) { fprintf(stderr, "nerf_b200: schedule: K tail of 3 steps is not supported\n"); abort(); }
            sl.flags = (uint32_t)p.kind | (nk == 2 ? kFlagNk2 : 0) | (nk == 4 ? kFlagNk4 : 0) | (g.n == 128 ? kFlagN128 : 0);
            PackSlot ps{};
            ps.goff = goff; ps.w_base = p.w0 + k0 * p.ks; ps.n_stride = p.ns; ps.k_stride = p.ks;
            ps.kvalid = p.kvalid - k0; ps.n = g.n; ps.nk = nk; ps.sw = (nk == 4) ? 128 : (nk == 2) ? 64 : 32;
            s.slots.push_back(sl); s.pack.push_back(ps);
            goff += kSlotBytes;          // fixed stride keeps every slot 8 KB aligned in the image
          }
        }
        if (g.b_off >= 0) {
          Slot sl{};
          sl.goff = goff; sl.bytes = (uint32_t)g.n * 32; sl.a_add = 0;
          sl.flags = (uint32_t)A_ONES | (g.n == 128 ? kFlagN128 : 0);
          PackSlot ps{};
          ps.goff = goff; ps.n = g.n; ps.nk = 1; ps.sw = 32; ps.is_bias = 1; ps.b_off = g.b_off;
          s.slots.push_back(sl); s.pack.push_back(ps);
          goff += kSlotBytes;
        }
        s.slots[first_idx].flags |= kFlagFirst;
        s.slots.back().flags |= kFlagLast;
        if (&gs == &fwd && ++gi == 8) s.n_fwd_sigma = (int)s.slots.size();
      }
    };
    emit(fwd);
    s.n_fwd = (int)s.slots.size();
    emit(bwd);
    s.n_bwd = (int)s.slots.size() - s.n_fwd;
    s.bytes = goff;
  });
  return s;
}

static int upload_schedule() {
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  NERF_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 64 && done[dev]) return 0;
  const Schedule& s = schedule();
  const int n = (int)s.slots.size();
  NERF_CHECK_ARG(n <= kMaxSlots, "slot table overflow (%d)", n);
  NERF_CUDA(cudaMemcpyToSymbol(c_slots, s.slots.data(), n * sizeof(Slot)));
  NERF_CUDA(cudaMemcpyToSymbol(c_pack, s.pack.data(), n * sizeof(PackSlot)));
  NERF_CUDA(cudaMemcpyToSymbol(c_nslots_fwd, &s.n_fwd, sizeof(int)));
  NERF_CUDA(cudaMemcpyToSymbol(c_nslots_bwd, &s.n_bwd, sizeof(int)));
  NERF_CUDA(cudaMemcpyToSymbol(c_nslots_fwd_sigma, &s.n_fwd_sigma, sizeof(int)));
  if (dev < 64) done[dev] = true;
  return 0;
}

size_t mlp_tc_packed_bytes() { return schedule().bytes; }

// ---- pack: flat fp32 parameters -> bf16 slot images (forward + transposed) ----------------------
__global__ void pack_kernel(const float* __restrict__ params, uint8_t* __restrict__ packed, int nslots) {
  const int si = blockIdx.y;
  if (si >= nslots) return;
  const PackSlot ps = c_pack[si];
  const int kw = 16 * ps.nk;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ps.n * kw; e += gridDim.x * blockDim.x) {
    const int n = e / kw, kk = e % kw;
    float v = 0.f;
    if (ps.is_bias) {
      const float b = params[ps.b_off + n];
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      if (kk == 0) v = hi;
      else if (kk == 1) v = b - hi;          // lo part: bias reaches the accumulator with ~16 mantissa bits
    } else if (kk < ps.kvalid) {
      v = params[ps.w_base + (int64_t)n * ps.n_stride + (int64_t)kk * ps.k_stride];
    }
    const uint32_t off = (ps.sw == 128) ? sw128_off(n, kk) : (ps.sw == 64) ? sw64_off(n, kk) : sw32_off(n, kk);
    *reinterpret_cast<__nv_bfloat16*>(packed + ps.goff + off) = __float2bfloat16_rn(v);
  }
}

int mlp_tc_pack(const float* params, void* packed, cudaStream_t st) {
  int rc = upload_schedule();
  if (rc) return rc;
  const int n = (int)schedule().slots.size();
  dim3 grid(8, (unsigned)n);
  pack_kernel<<<grid, 256, 0, st>>>(params, (uint8_t*)packed, n);
  NERF_LAUNCH_CHECK("pack_kernel");
  return 0;
}

// ---- view bias: vb[row][n] = b_view[n] + sum_j W_view[n][256+j] * dir_enc[j]   (fp32) -------------
// rays entry: one row per ray, direction normalised d/(|d|+1e-8) and encoded with L=4 here
// (reference renderer.py:72-74); encoded entry: one row per sample from d_enc.
// kVbRows rows per block: thread n keeps its 27 direction weights of view_linear in registers across the rows
// (one row per block re-read the strided weight columns for every ray: 7.7 us per 1024 rays).
constexpr int kVbRows = 8;
__global__ void __launch_bounds__(128) view_bias_kernel(const float* __restrict__ rays_d, const float* __restrict__ d_enc,
                                                       int64_t nrows, const float* __restrict__ params,
                                                       float* __restrict__ vb, float* __restrict__ de_out) {
  __shared__ float de[kVbRows][28];
  const int64_t row0 = (int64_t)blockIdx.x * kVbRows;
  const int tid = threadIdx.x;
  if (d_enc != nullptr) {
    for (int i = tid; i < kVbRows * 27; i += 128) {
      const int r = i / 27, j = i - r * 27;
      if (row0 + r < nrows) de[r][j] = d_enc[(row0 + r) * 27 + j];
    }
  } else if (tid < kVbRows * 3) {
    const int r = tid / 3, c = tid - r * 3;
    const int64_t row = row0 + r;
    if (row < nrows) {
      const float dx = rays_d[3 * row], dy = rays_d[3 * row + 1], dz = rays_d[3 * row + 2];
      const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      const float v = __fdiv_rn(rays_d[3 * row + c], __fadd_rn(nrm, 1e-8f));                  // renderer.py:72
      de[r][c] = v;
      float f = 1.f;
      for (int k = 0; k < 4; ++k, f *= 2.f) {
        float sn, cs;
        sincosf(f * v, &sn, &cs);
        de[r][3 + 6 * k + c] = sn;                                                            // model.py:24
        de[r][3 + 6 * k + 3 + c] = cs;                                                        // model.py:25
      }
    }
  }
  __syncthreads();
  const int n = tid;
  const float* w = params + w_off(L_VIEW) + (int64_t)n * 283 + 256;
  float wr[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) wr[j] = w[j];
  const float b = params[b_off(L_VIEW) + n];
#pragma unroll 1
  for (int r = 0; r < kVbRows; ++r) {
    const int64_t row = row0 + r;
    if (row >= nrows) break;
    float acc = b;
#pragma unroll
    for (int j = 0; j < 27; ++j) acc = fmaf(wr[j], de[r][j], acc);
    vb[row * 128 + n] = acc;
    if (de_out != nullptr && n < 32) de_out[row * 32 + n] = n < 27 ? de[r][n] : 0.f;
  }
}

// ---- kernel arguments ---------------------------------------------------------------------------------
struct TcArgs {
  // forward inputs
  const float* rays_o; const float* rays_d; const float* z_vals; int S; float coord_scale;
  const float* x_enc;
  const float* vb; int vb_div;
  float* out;                       // [M,4]
  // backward input
  const float* d_raw;               // [M,4]
  // common
  int64_t M;
  const uint8_t* packed;
  const float* params;
  // saved tensors (forward writes when kSave, dgrad reads the masks); *_img = tile images (tc_common.cuh)
  uint8_t* act_img; uint8_t* hv_img; uint8_t* xenc_img; uint8_t* de16_img; uint32_t* mask; uint32_t* hvmask;
  const float* de;                  // fp32 [nvb][32] encoded directions (written by view_bias_kernel)
  // dgrad outputs
  uint8_t* dpre_img; uint8_t* dhv_img;
  int64_t Mp;                       // rows padded to whole tile pairs
  int num_pairs;
  // work units: units [0, n_full_units) hold two tiles per CTA; the units after them ("half units", at most one per
  // CTA, always a CTA's last) hold ONE tile per CTA -- the tiles of a last partial wave are spread over twice as many
  // SMs instead of leaving most of them idle for a whole unit (196 608 rows = 5.19 waves of 148 x 2 tiles)
  int n_full_units, n_units;
};

// bf16 x-tile row: 63 encoded channels (+ a zero pad column) -> 8 swizzled 16-byte chunks
__device__ __forceinline__ void store_x_row(uint8_t* xt, int m, const float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 q;
    q.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]);
    q.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
    q.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
    q.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
    *reinterpret_cast<uint4*>(xt + sw128_off(m, 8 * c)) = q;
  }
}

__device__ __forceinline__ void fwd_prologue(const TcArgs& a, int64_t row, int m, uint8_t* xt) {
  float v[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) v[j] = 0.f;
  if (row < a.M) {
    if (a.x_enc != nullptr) {
      const float* x = a.x_enc + row * 63;
#pragma unroll
      for (int j = 0; j < 63; ++j) v[j] = __ldg(x + j);
    } else {
      const int64_t r = row / a.S;
      const float z = __ldg(a.z_vals + row);
      float p[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        p[i] = __fadd_rn(__ldg(a.rays_o + 3 * r + i), __fmul_rn(__ldg(a.rays_d + 3 * r + i), z));   // renderer.py:63
        if (a.coord_scale != 1.f) p[i] = __fmul_rn(p[i], a.coord_scale);                               // :67-68
        v[i] = p[i];
      }
      // sin/cos(2^k p): accurate sincosf at k = 0 and k = 5, double-angle steps in between
      // (4 doublings amplify a ~1e-7 error to ~2e-6, far below bf16 resolution).
      float s[3], c[3];
#pragma unroll
      for (int k = 0; k < 10; ++k) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (k == 0) sincosf(p[i], &s[i], &c[i]);
          else if (k == 5) sincosf(32.f * p[i], &s[i], &c[i]);
          else {
            const float s2 = 2.f * s[i] * c[i];
            const float c2 = (c[i] - s[i]) * (c[i] + s[i]);
            s[i] = s2; c[i] = c2;
          }
          v[3 + 6 * k + i] = s[i];                                                                     // model.py:24
          v[3 + 6 * k + 3 + i] = c[i];                                                                 // model.py:25
        }
      }
    }
  }
  store_x_row(xt, m, v);
}

// ReLU mask of 32 fp32 accumulators, one funnel shift per element: bit j = !sign(r[j]).
// Four independent 8-bit chains (a single 32-long dependent chain was the top stall of the
// save-mode epilogue), merged with two byte permutes.
// (An accumulator that is exactly +0 counts as active; its gradient contribution is multiplied by
// a zero activation downstream only in the weight gradient, and the event has measure zero.)
__device__ __forceinline__ uint32_t relu_mask32(const uint32_t (&r)[32]) {
  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
  for (int j = 7; j >= 0; --j) {                                     // w = (w << 1) | (r[j] >> 31)
    w0 = __funnelshift_l(r[j], w0, 1);
    w1 = __funnelshift_l(r[8 + j], w1, 1);
    w2 = __funnelshift_l(r[16 + j], w2, 1);
    w3 = __funnelshift_l(r[24 + j], w3, 1);
  }
  return ~(__byte_perm(w0, w1, 0x1140) | __byte_perm(w2, w3, 0x4011));   // bytes: w0, w1, w2, w3 (upper bytes of each are 0)
}

// kBwd = false: forward (kSave: also write the tensors the backward needs); kBwd = true: dgrad chain
// kSigmaOnly (inference forward only): stop after layer 7 and the sigma head -- the coarse pass of a
// render / training step only feeds the hierarchical resampling, which needs the density alone
// (renderer.py:79-87 uses `weights`; the coarse colour maps are dropped by render(), renderer.py:44,
// and never reach the loss, scripts/train.py:374-376).  Output rows are (0, 0, 0, sigma).
template <bool kBwd, bool kSave, int kCtas, bool kSigmaOnly = false>
__global__ void __launch_bounds__((tc_threads<kBwd || kSave>()), 1) mlp_tc_kernel(const TcArgs a) {
  static_assert(!kSigmaOnly || (!kBwd && !kSave), "density-only mode is an inference-forward mode");
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr bool kCopy = (kBwd || kSave) && kCopyWarps > 0;                     // dedicated copy-out warps (11 .. 11 + kCopyWarps - 1)
  constexpr bool kShared = (!kBwd && !kSave) || kCopy || (NERF_TC_SHARED_TRAIN != 0);   // all eight compute warps share every epilogue
  constexpr bool kSplit = (!kBwd && !kSave) ? (NERF_TC_SPLIT_INFER != 0) : (NERF_TC_SPLIT_TRAIN != 0);   // private weight stream per tile
  constexpr int kRingK = kRing * kCtas;                 // ring slots per CTA (same bytes, half-size slots in pair mode)
  constexpr int kRingT = kSplit ? kRingK / 2 : kRingK;  // slots of one stream
  static_assert(!kSplit || (kRingK % 2 == 0 && kRingT >= 2), "split mode needs two slots per stream");
  constexpr int kSlotK = kSlotBytes / kCtas;            // bytes of a slot held by one CTA
  const uint32_t rank = (kCtas == 2) ? cluster_ctarank() : 0u;
  const int unit0 = (kCtas == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // first work unit of this CTA (pair)
  const int unit_step = (kCtas == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_units = a.n_units;                                                  // unit = tile pair (1 CTA) / quad (CTA pair); half units last
  // first tile of this CTA in `unit`, and how many tiles (2, or 1 in a half unit) it has there
  auto unit_tiles = [&](int unit, int64_t& tile0) -> int {
    if (unit < a.n_full_units) { tile0 = ((int64_t)unit * kCtas + rank) * 2; return 2; }
    tile0 = (int64_t)a.n_full_units * kCtas * 2 + (int64_t)(unit - a.n_full_units) * kCtas + rank;
    return 1;
  };
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kRingK + s); };
  auto bar_pfull = [&](int s) { return bar0 + 8u * (2 * kRingK + s); };    // pair mode, leader: the peer's half of slot s landed
  constexpr int kB0 = kBarSets * kRingK;
  auto bar_act = [&](int t) { return bar0 + 8u * (kB0 + t); };             // activations of tile t ready (epilogue -> MMA)
  auto bar_acc = [&](int t) { return bar0 + 8u * (kB0 + 2 + t); };         // accumulator of tile t ready (MMA -> epilogue)
  const uint32_t bar_skew = bar0 + 8u * (kB0 + 4);                         // one-shot: tile A's issuer is kSkew slots in
  auto bar_cp = [&](int t) { return bar0 + 8u * (kB0 + 5 + t); };          // tile t complete in shared memory: copy it out (compute -> copy warps)
  auto bar_cpfree = [&](int t) { return bar0 + 8u * (kB0 + 7 + t); };      // the copy warps have read tile t: it may be overwritten
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);
  float* head = reinterpret_cast<float*>(smem + kOffHead);
  constexpr int kNumGemms = kBwd ? kNumGemmsBwd : (kSigmaOnly ? 8 : kNumGemmsFwd);

  if (warp == 0 && lane == 0) {
    constexpr bool kRelayFull = NERF_TC_RELAY_FULL != 0;
    for (int s = 0; s < kRingK; ++s) {
      mbar_init(bar_full(s), (kCtas == 2 && kRelayFull && rank == 0) ? 2 : 1);      // leader: own bytes + the peer's relay
      mbar_init(bar_empty(s), kSplit ? 1 : 2);
      if (kBarSets == 3) mbar_init(bar_pfull(s), 1);
    }
    for (int t = 0; t < 2; ++t) { mbar_init(bar_act(t), (kShared ? 256 : 128) / kArriveLanes * kCtas); mbar_init(bar_acc(t), 1); }
    mbar_init(bar_skew, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(bar_cp(t), 8); mbar_init(bar_cpfree(t), kCopyWarps > 0 ? kCopyWarps : 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kCtas == 2) { tmem_alloc2(sbase + kOffTmemPtr, 512); tmem_relinquish2(); }
    else { tmem_alloc(sbase + kOffTmemPtr, 512); tmem_relinquish(); }
  }
  if (warp >= kFirstComputeWarp) {
    const int tid = threadIdx.x - 32 * kFirstComputeWarp;   // 0..255
    // constant A operand of the bias K-step: columns 0,1 = 1, rest 0 (SW32 layout)
    if (tid < 8) {
      const uint32_t one2 = 0x3F803F80u;              // bf16 (1.0, 1.0)
      uint8_t* ones = smem + kOffOnes;
      *reinterpret_cast<uint4*>(ones + sw32_off(tid, 0)) = make_uint4(one2, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(ones + sw32_off(tid, 8)) = make_uint4(0u, 0u, 0u, 0u);
    }
    // fp32 head weights: sigma_linear (model.py:69) and rgb_linear (:75)
    for (int i = tid; i < 256; i += 256) head[i] = a.params[w_off(L_SIGMA) + i];
    for (int i = tid; i < 384; i += 256) head[256 + i] = a.params[w_off(L_RGB) + i];
    if (tid == 0) head[640] = a.params[b_off(L_SIGMA)];
    if (tid < 3) head[641 + tid] = a.params[b_off(L_RGB) + tid];
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  if (kCtas == 2) cluster_sync();                     // both CTAs' barriers exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int slot0 = kBwd ? c_nslots_fwd : 0;
  const int nslots = kBwd ? c_nslots_bwd : (kSigmaOnly ? c_nslots_fwd_sigma : c_nslots_fwd);

  if (warp == 0) {
    // ================= weight producer =================
    // shared stream: lane 0 fills the whole ring; split mode: lane t fills tile t's half-ring
    if (lane < (kSplit ? 2 : 1)) {
      const uint32_t base = kSplit ? (uint32_t)lane * kRingT : 0u;
      uint32_t g = 0;
      for (int unit = unit0; unit < num_units; unit += unit_step) {
        for (int i = 0; i < nslots; ++i, ++g) {
          const uint32_t s = base + g % kRingT, ph = (g / kRingT) & 1;
          const uint2 rec = *reinterpret_cast<const uint2*>(&c_slots[slot0 + i]);      // goff, bytes
          const uint32_t bytes = rec.y / kCtas;          // pair mode: this CTA's N-half of the slot
          mbar_wait(bar_empty(s), ph ^ 1, 100 + (int)s);
          mbar_expect_tx(bar_full(s), bytes);
          bulk_g2s(sbase + kOffRing + s * kSlotK, a.packed + rec.x + rank * bytes, bytes, bar_full(s));
        }
      }
    }
  } else if (warp < kFirstComputeWarp) {
    // ================= MMA issuers: warp 1 -> tile A, warp 2 -> tile B =================
    // The whole warp walks the schedule (uniform control flow); one elected lane issues.
    if (kCtas == 2 && rank != 0) {
      // peer CTA of a pair: its issuer warps do not issue; warp 1 relays "my half of slot s landed"
      // to the leader, whose MMAs read both halves.
      if (warp == 1 || kSplit) {                       // split mode: warp 1 + t relays tile t's stream
        const uint32_t s_lo = kSplit ? (uint32_t)(warp - 1) * kRingT : 0u, s_hi = s_lo + kRingT;
        uint32_t s = s_lo, ph = 0;
        for (int unit = unit0; unit < num_units; unit += unit_step) {
          for (int i = 0; i < nslots; ++i) {
            mbar_wait(bar_full(s), ph, 250 + (int)s);
            if (lane == 0) mbar_arrive_cluster(mapa((NERF_TC_RELAY_FULL != 0) ? bar_full(s) : bar_pfull(s), 0));
            __syncwarp();
            if (++s == s_hi) { s = s_lo; ph ^= 1; }
          }
        }
      }
    } else {
      const int t = warp - 1;
      // descriptor words that never change: hi = SBO | version | layout, lo = (addr >> 4) | LBO
      constexpr uint32_t kHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t kHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);
      constexpr uint32_t kHiSw32 = (256u >> 4) | (1u << 14) | (6u << 29);
      constexpr uint32_t kHiOnes = (0u >> 4) | (1u << 14) | (6u << 29);      // SBO = 0: every 8-row group reads the same 8 rows
      const uint32_t a_lo_x = ((sbase + kOffX + t * kXBytes) >> 4) | (1u << 16);
      const uint32_t a_lo_act = ((sbase + kOffAct + t * kActBytes) >> 4) | (1u << 16);
      const uint32_t a_lo_ones = ((sbase + kOffOnes) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((sbase + kOffRing) >> 4) | (1u << 16);
      const uint32_t d_tmem = tmem_base + (uint32_t)t * 256;
      constexpr uint32_t kIdesc256 = make_idesc_bf16(kTileM * kCtas, 256), kIdesc128 = make_idesc_bf16(kTileM * kCtas, 128);
      const uint32_t s_lo = kSplit ? (uint32_t)t * kRingT : 0u, s_hi = s_lo + kRingT;
      uint32_t s = s_lo, ph = 0, act_ph = 0;
      // one-shot start offset behind tile A: kSkew slots (shared stream), or -- split mode -- until tile A's
      // first epilogue has finished, i.e. half a period of the (now independent) per-tile chains
      if (t == 1) mbar_wait(bar_skew, 0, 500);
      for (int unit = unit0; unit < num_units; unit += unit_step) {
        const bool idle = (t == 1) && unit >= a.n_full_units;      // half unit: tile B does not exist
        for (int i = 0; i < nslots; ++i) {
          const uint4 rec = *reinterpret_cast<const uint4*>(&c_slots[slot0 + i]);
          const uint32_t a_add = rec.z, fl = rec.w;
          if (idle) {
            // keep the ring protocol (a slot is freed by BOTH issuers) without issuing: an empty commit arrives at once
            mbar_wait(bar_full(s), ph, 200 + (int)s);
            if (kCtas == 2 && NERF_TC_RELAY_FULL == 0) mbar_wait(bar_pfull(s), ph, 220 + (int)s);
            tc_fence_after();
            if (elect_one()) {
              if (kCtas == 2) tc_commit_mc2(bar_empty(s), 3); else tc_commit(bar_empty(s));
            }
            __syncwarp();
            if (++s == s_hi) { s = s_lo; ph ^= 1; }
            continue;
          }
#ifdef NERF_DBG_TIMING
          const long long i0_ = clock64();
#endif
          // plain (cta-scope) waits, as CUTLASS's 2-SM pipelines do: the remote arrivals are
          // release.cluster and the data they publish is consumed by the async proxy (the MMA)
          if (NERF_TC_POLL_LANE0 == 0 || lane == 0) {
            mbar_wait(bar_full(s), ph, 200 + (int)s);
            if (kCtas == 2 && NERF_TC_RELAY_FULL == 0) mbar_wait(bar_pfull(s), ph, 220 + (int)s);
          }
#ifdef NERF_DBG_TIMING
          const long long i1_ = clock64();
#endif
          if (fl & kFlagFirst) {
            TRACE(t, lane == 0 && unit == unit0 + 2 * unit_step, 1);       // layer start: weights of the first slot are there
            if (NERF_TC_POLL_LANE0 == 0 || lane == 0) mbar_wait(bar_act(t), act_ph, 300 + t);
            act_ph ^= 1;
            TRACE(t, lane == 0 && unit == unit0 + 2 * unit_step, 2);       // activations ready
          }
          if (NERF_TC_POLL_LANE0 != 0) __syncwarp();                       // lane 0's acquire is ordered before the elected lane's issue
          tc_fence_after();
#ifdef NERF_DBG_TIMING
          if (blockIdx.x == 0 && warp == 1 && lane == 0) {
            g_dbg_clk[4] += (unsigned long long)(i1_ - i0_);
            g_dbg_clk[5] += (unsigned long long)(clock64() - i1_);
            g_dbg_clk[6] += 1;
          }
#endif
          const uint32_t kind = fl & 3u;
          const uint32_t a_lo = (kind == A_X ? a_lo_x : (kind == A_ACT ? a_lo_act : a_lo_ones)) + a_add;
          const uint32_t a_hi = (kind == A_ONES) ? kHiOnes : kHiSw128;
          const uint32_t b_lo = b_lo0 + s * (kSlotK >> 4);
          const uint32_t b_hi = (fl & kFlagNk4) ? kHiSw128 : (fl & kFlagNk2) ? kHiSw64 : kHiSw32;
          const uint32_t idesc = (fl & kFlagN128) ? kIdesc128 : kIdesc256;
          if (elect_one()) {
            if (kCtas == 2) {
              mma_bf16_ss_2cta(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc,
                               (fl & kFlagFirst) ? 0u : 1u);
              if (fl & (kFlagNk2 | kFlagNk4))
                mma_bf16_ss_2cta(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 2), ((uint64_t)b_hi << 32) | (b_lo + 2), idesc, 1u);
              if (fl & kFlagNk4) {
                mma_bf16_ss_2cta(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 4), ((uint64_t)b_hi << 32) | (b_lo + 4), idesc, 1u);
                mma_bf16_ss_2cta(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 6), ((uint64_t)b_hi << 32) | (b_lo + 6), idesc, 1u);
              }
              tc_commit_mc2(bar_empty(s), 3);               // both CTAs' producers may refill the slot
              if (fl & kFlagLast) tc_commit_mc2(bar_acc(t), 3);   // both CTAs' epilogues may read their accumulators
            } else {
              mma_bf16_ss(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc,
                          (fl & kFlagFirst) ? 0u : 1u);
              if (fl & (kFlagNk2 | kFlagNk4))
                mma_bf16_ss(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 2), ((uint64_t)b_hi << 32) | (b_lo + 2), idesc, 1u);
              if (fl & kFlagNk4) {
                mma_bf16_ss(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 4), ((uint64_t)b_hi << 32) | (b_lo + 4), idesc, 1u);
                mma_bf16_ss(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 6), ((uint64_t)b_hi << 32) | (b_lo + 6), idesc, 1u);
              }
              tc_commit(bar_empty(s));                      // slot is free once both issuers' MMAs on it retired
              if (fl & kFlagLast) tc_commit(bar_acc(t));
            }
            if (!kSplit && t == 0 && i == kSkew - 1 && unit == unit0) mbar_arrive(bar_skew);
          }
          __syncwarp();
          if (fl & kFlagLast) TRACE(t, lane == 0 && unit == unit0 + 2 * unit_step, 3);   // last MMA of the layer issued + committed
          if (++s == s_hi) { s = s_lo; ph ^= 1; }
        }
      }
    }
  } else if (warp >= kFirstComputeWarp + 8) {
    // ================= copy-out warps (save-mode forward / dgrad only) =================
    if constexpr (kCopy) {
      const int w = warp - (kFirstComputeWarp + 8);     // feature block of the tile this warp moves
      const int64_t ntiles = a.Mp / kTileM;
      uint32_t cp_ph = 0u;                              // bit t = phase of bar_cp(t)
      auto copy_tile = [&](int t, uint8_t* dst, int nfb) {
        mbar_wait(bar_cp(t), (cp_ph >> t) & 1u, 960 + t);
        cp_ph ^= 1u << t;
#ifdef NERF_DBG_NOCOPY       // timing experiment only (nothing is saved): what does the copy-out traffic cost?
        if (false) {
#else
        if (w < nfb) {
#endif
          const uint8_t* sp = smem + kOffAct + t * kActBytes + w * 16384 + lane * 16;
          uint8_t* dp = dst + w * 16384 + lane * 16;
#pragma unroll 1
          for (int i0 = 0; i0 < 32; i0 += 8) {
            uint4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4*>(sp + (i0 + i) * 512);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // forward: saved for a reader a whole pass (> 1 GB of traffic) away -> streaming store (evict-first), so
              // that the tiles do not push the weights and the next kernels' working set out of L2 (step 0.976 ->
              // 0.959 ms).  dgrad kernel: its tiles are the next kernel's operands -> normal store.
              if (!kBwd) __stcs(reinterpret_cast<uint4*>(dp + (i0 + i) * 512), v[i]);
              else *reinterpret_cast<uint4*>(dp + (i0 + i) * 512) = v[i];
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_cpfree(t));
      };
      for (int unit = unit0; unit < num_units; unit += unit_step) {
        int64_t tile0;
        const int ntl = unit_tiles(unit, tile0);
        if constexpr (kBwd) {
          for (int t = 0; t < ntl; ++t) copy_tile(t, a.dhv_img + (tile0 + t) * 32768, 2);
          for (int g = 0; g < kNumGemmsBwd; ++g) {
            const int dst = (g == 0) ? 8 : 8 - g;
            for (int t = 0; t < ntl; ++t) copy_tile(t, a.dpre_img + ((int64_t)dst * ntiles + tile0 + t) * 65536, 4);
          }
        } else {
          for (int g = 0; g < 9; ++g)
            for (int t = 0; t < ntl; ++t)
              copy_tile(t, a.act_img + ((int64_t)g * ntiles + tile0 + t) * 65536, 4);
        }
      }
    }
  } else {
    // ================= prologue + epilogue warps =================
    // Group h (warps 3-6 / 7-10) owns the PROLOGUE of tile h (one row per thread).  Every EPILOGUE is
    // shared by all eight warps: warp (q, h) drains TMEM lane quadrant q, columns [128h, 128h+128) of
    // whichever tile's accumulator is ready, so one epilogue takes half as long and the dependent
    // chain  MMA(l) -> epilogue(l) -> MMA(l+1)  of a tile -- which was measured to pace the kernel,
    // above all in the save modes -- shortens accordingly; the two warps of a scheduler partition
    // now overlap each other's TMEM-load / shared-store latencies.
    const int h = (warp - kFirstComputeWarp) >> 2;    // prologue: own tile; shared epilogues: column half
    // Epilogue ownership (kShared is a kernel-level constant, see bar_act's arrival count):
    //   shared   (inference forward)  : every warp works on BOTH tiles, column half h
    //   separate (save forward, dgrad): warp group h works on tile h only, both column halves in turn.
    //     Those kernels are bound by the SM's shared-memory / L1 data path (operand reads + tile
    //     writes + the tile-image copy-out), not by the epilogue latency; sharing measured 7 % slower.
    const int t_lo = kShared ? 0 : h, t_hi = kShared ? 2 : h + 1;
    const int ch_lo = kShared ? h : 0, ch_hi = kShared ? h + 1 : 2;
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;                      // row within a tile
    constexpr bool kStores = kBwd || kSave;           // does this kernel write tile images to HBM?
    // Tile-image saves: this warp's 32 rows of one 64-feature block are ONE contiguous 4 KB block, in
    // shared memory and in the HBM tile image alike, and are written by this warp only -> a
    // coalesced per-warp copy with warp-level synchronisation.
    // (Ownership rule: a warp only ever copies out blocks it wrote itself, and the next writer of
    // those bytes is the same warp, or a warp that has since passed a barrier this warp reached
    // after the copy.)
    auto store_blocks = [&](uint8_t* dst, int dst_fb0, const uint8_t* src, int src_fb0, int nfb) {
      __syncwarp();
      const uint8_t* sp = src + src_fb0 * 16384 + q * 4096 + lane * 16;
      uint8_t* dp = dst + dst_fb0 * 16384 + q * 4096 + lane * 16;
      for (int fb = 0; fb < nfb; ++fb) {
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4*>(sp + fb * 16384 + i * 512);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!kBwd) __stcs(reinterpret_cast<uint4*>(dp + fb * 16384 + i * 512), v[i]);      // forward saves: streaming (see copy_tile)
          else *reinterpret_cast<uint4*>(dp + fb * 16384 + i * 512) = v[i];
        }
      }
      __syncwarp();     // lanes read each other's rows: nobody may overwrite them before all are done
    };
    // "activations of tile t ready": local, or (peer CTA of a pair) remote at the leader.  kArriveLanes == 32 is
    // the fence / __syncwarp / elected-arrive form (one arrive per warp instead of 32); it was measured to make
    // no difference, so the per-thread arrives are not what stretches the chain epilogue(l) -> MMA(l+1).
    auto act_arrive = [&](int t) {
      if (kArriveLanes == 32) { __syncwarp(); if (lane != 0) return; }
      if (kCtas == 2 && rank != 0) mbar_arrive_cluster(mapa(bar_act(t), 0)); else mbar_arrive(bar_act(t));
    };
    // copy-out handshake (kCopy): request after the tile is complete + fenced, wait before the tile is written again
    uint32_t cp_pend = 0u, cpfree_ph = 0u;            // bit t: a copy of tile t is outstanding / phase of bar_cpfree(t)
    auto cp_request = [&](int t) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_cp(t));
      cp_pend |= 1u << t;
    };
    auto cp_wait_free = [&](int t) {
      if (cp_pend & (1u << t)) {
        mbar_wait(bar_cpfree(t), (cpfree_ph >> t) & 1u, 970 + t);
        cpfree_ph ^= 1u << t;
        cp_pend &= ~(1u << t);
      }
    };
    uint32_t acc_ph = 0u;                             // bit t = phase of bar_acc(t)
    const int64_t ntiles = a.Mp / kTileM;
    // cross-warp scratch for the head partial sums (forward, last layer of tile t): 512 B per lane quadrant in
    // feature block 3 of tile t's OWN activation tile -- its last reader (the MMAs of the last layer) has
    // completed when the accumulator barrier fires, the save mode stages hv / dir_enc in blocks 0-2 only, and the
    // next writer (tile t's layer-0 epilogue of the next unit) needs tile t's next "activations ready" phase,
    // which every thread joins only after it has read the scratch.
    // (Not the x-tile: with NERF_TC_PIPE_PROLOGUE it already holds the next unit's encoding by then.)
    auto xchg_of = [&](int t) { return reinterpret_cast<float4*>(smem + kOffAct + t * kActBytes + 3 * 16384 + q * 1024); };
    for (int unit = unit0; unit < num_units; unit += unit_step) {
      int64_t tile0;
      const int ntl = unit_tiles(unit, tile0);               // 2 tiles, or 1 in a half unit (tile B and its group's prologue idle)
      const bool own = h < ntl;                              // this warp group's own tile (prologue) exists
#ifdef NERF_DBG_TRACE
      const int tw = (warp == kFirstComputeWarp) ? 2 : 3;
      const bool tr_ok = (warp == kFirstComputeWarp || warp == kFirstComputeWarp + 4) && lane == 0 && unit == unit0 + 2 * unit_step;
#endif
      if constexpr (!kBwd) {
        // ------------------------------ forward ------------------------------
        uint8_t* xt_own = smem + kOffX + h * kXBytes;
        // inference kernels only: in the save-mode kernel the compute warps are the bottleneck and it measured slower (0.352 vs 0.346 ms)
        constexpr bool kPipe = (NERF_TC_PIPE_PROLOGUE != 0) && !kSave;
        {                                                  // prologue of the own tile (pipelined mode: first unit only)
          const int64_t row = (tile0 + h) * kTileM + m;
          // (pipelined mode, later units: the x-tile was written under the previous unit's layer-6 wait and the
          //  arrives were made right after each tile's last epilogue, so tile A's layer 0 runs under tile B's
          //  last epilogue)
          if (!kPipe || unit == unit0) {
            if (kShared && 1 - h < ntl) act_arrive(1 - h); // nothing to write for the other tile
            if (own) {
              fwd_prologue(a, row, m, xt_own);
              fence_proxy_async();
              act_arrive(h);
            }
          }
          if (kSave && own && (!kPipe || unit == unit0)) store_blocks(a.xenc_img + (tile0 + h) * 16384, 0, xt_own, 0, 1);
        }
        float sigma0 = 0.f, sigma1 = 0.f;                  // partial sigma head of (tile A / B, own column half)
        for (int g = 0; g < kNumGemms; ++g) {
#pragma unroll 1
          for (int t = t_lo; t < t_hi && t < ntl; ++t) {
            float sigma = t ? sigma1 : sigma0;
            const int64_t tile = tile0 + t;
            const int64_t row = tile * kTileM + m;
            const bool valid = row < a.M;
            uint8_t* at = smem + kOffAct + t * kActBytes;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256;
            DBG_T(t0_);
            if (NERF_TC_VB_PREFETCH != 0 && !kSigmaOnly && g == 9) {
              // the view-layer epilogue adds this row's 512-byte view-bias vector (one per ray, L2-resident): pull its
              // four lines into L1 while waiting for the accumulator instead of paying the L2 latency twice inside the
              // epilogue (the timeline showed 2 800 clk per view epilogue against ~1 050 for a trunk layer)
              const char* vbp = reinterpret_cast<const char*>(a.vb + ((valid ? row : (a.M - 1)) / a.vb_div) * 128);
#pragma unroll
              for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(vbp + 128 * i));
            }
            mbar_wait(bar_acc(t), (acc_ph >> t) & 1u, 400 + t);
            acc_ph ^= 1u << t;
            tc_fence_after();
            if (kCopy) cp_wait_free(t);                        // the previous tensor of this tile has left shared memory
            TRACE(tw, tr_ok, 4 + 8 * t);
            DBG_T(t1_);
            DBG_ACC(0, t1_ - t0_);
            if (kSigmaOnly && g == 7) {
              // density-only forward, last layer: sigma head straight from the fp32 accumulators; nothing is
              // written back to the activation tile and no MMA follows (the next arrive is the next prologue's)
#pragma unroll 1
              for (int c0 = 128 * h; c0 < 128 * h + 128; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) sigma = fmaf(fmaxf(__uint_as_float(r[j]), 0.f), head[c0 + j], sigma);
              }
              tc_fence_before();
              float4* xchg = xchg_of(t);
              if (h == 1) xchg[lane] = make_float4(0.f, 0.f, 0.f, sigma);
              named_bar_sync(1, 256);
              if (h == 0 && valid)
                *reinterpret_cast<float4*>(a.out + row * 4) = make_float4(0.f, 0.f, 0.f, sigma + xchg[lane].w + head[640]);
              if (kPipe && unit + unit_step < num_units && (t == 0 || unit + unit_step < a.n_full_units))
                act_arrive(t);                                                // next unit's layer 0 of this tile may start
            } else if (g < 9) {
#pragma unroll 1
             for (int ch = ch_lo; ch < ch_hi; ++ch) {
              uint32_t mkw[4];
              // one 32-column chunk: sigma head, ReLU mask, convert, swizzled store
              auto chunk = [&](const uint32_t (&r)[32], int c0, int mi) {
                if (g == 7) {                              // sigma head from fp32 post-ReLU activations (model.py:69)
#pragma unroll
                  for (int j = 0; j < 32; ++j) sigma = fmaf(fmaxf(__uint_as_float(r[j]), 0.f), head[c0 + j], sigma);
                }
#ifndef NERF_DBG_NOMASK      // timing experiment only (no ReLU masks are saved)
                if (kSave) mkw[mi] = relu_mask32(r);
#endif
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  uint4 o;
                  if (g != 8) {                            // ReLU on every trunk layer (model.py:65); bottleneck has none (:70)
                    o.x = pack_bf16x2_relu(__uint_as_float(r[8 * c + 0]), __uint_as_float(r[8 * c + 1]));
                    o.y = pack_bf16x2_relu(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
                    o.z = pack_bf16x2_relu(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]));
                    o.w = pack_bf16x2_relu(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
                  } else {
                    o.x = pack_bf16x2(__uint_as_float(r[8 * c + 0]), __uint_as_float(r[8 * c + 1]));
                    o.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
                    o.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]));
                    o.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
                  }
                  const int k = c0 + 8 * c;
                  *reinterpret_cast<uint4*>(at + (k >> 6) * 16384 + sw128_off(m, k & 63)) = o;
                }
              };
              {                                            // two register buffers: the next chunk's tcgen05.ld is in
                uint32_t ra[32], rb[32];                   // flight while the current chunk is converted and stored
                const int cb = 128 * ch;
                tmem_ld32(taddr + cb, ra);
#pragma unroll                                             // static offsets: the mask words stay in registers
                for (int c0 = 0; c0 < 128; c0 += 64) {
                  tmem_ld_wait();
                  tmem_ld32(taddr + cb + c0 + 32, rb);
                  chunk(ra, cb + c0, c0 >> 5);
                  tmem_ld_wait();
                  if (c0 + 64 < 128) tmem_ld32(taddr + cb + c0 + 64, ra);
                  chunk(rb, cb + c0 + 32, (c0 >> 5) + 1);
                }
              }
#ifndef NERF_DBG_NOMASK
              if (kSave && g < 8 && valid)
                *reinterpret_cast<uint4*>(a.mask + ((int64_t)g * a.M + row) * 8 + 4 * ch) =
                    make_uint4(mkw[0], mkw[1], mkw[2], mkw[3]);
#endif
             }
              tc_fence_before();
              fence_proxy_async();
              act_arrive(t);
              if (kSplit && g == 0 && t == 0 && unit == unit0 && warp == kFirstComputeWarp && lane == 0) mbar_arrive(bar_skew);   // start tile B's chain
              TRACE(tw, tr_ok, 5 + 8 * t);
              DBG_T(t2_);
              DBG_ACC(1, t2_ - t1_);
              DBG_ACC(3, 1);
              if (g == 7) { if (t) sigma1 = sigma; else sigma0 = sigma; }
              if (kSave && kCopy) cp_request(t);          // the copy warps move the tile to the workspace
              if (kSave && !kCopy) {                       // off the critical path: the MMAs are already released
#pragma unroll 1
                for (int ch = ch_lo; ch < ch_hi; ++ch)
                  store_blocks(a.act_img + ((int64_t)g * ntiles + tile) * 65536, 2 * ch, at, 2 * ch, 2);
                TRACE(tw, tr_ok, 6 + 8 * t);
                DBG_T(t3_);
                DBG_ACC(2, t3_ - t2_);
              }
            } else {
              // view layer epilogue: + view bias, ReLU (model.py:73-74), rgb head (:75), output [rgb, sigma] (:77)
              // this thread: 64 of the 128 columns
              const int64_t vrow = (valid ? row : (a.M - 1)) / a.vb_div;
              const float4* vb4 = reinterpret_cast<const float4*>(a.vb + vrow * 128);
              float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 1
             for (int ch = ch_lo; ch < ch_hi; ++ch) {
              uint32_t hmw[2];
#pragma unroll 1
              for (int ci = 0; ci < 2; ++ci) {
                const int c0 = 64 * ch + 32 * ci;
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                float hh[32];
                uint32_t mw = 0;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 b = __ldg(vb4 + (c0 >> 2) + j4);
                  hh[4 * j4 + 0] = fmaxf(__uint_as_float(r[4 * j4 + 0]) + b.x, 0.f);
                  hh[4 * j4 + 1] = fmaxf(__uint_as_float(r[4 * j4 + 1]) + b.y, 0.f);
                  hh[4 * j4 + 2] = fmaxf(__uint_as_float(r[4 * j4 + 2]) + b.z, 0.f);
                  hh[4 * j4 + 3] = fmaxf(__uint_as_float(r[4 * j4 + 3]) + b.w, 0.f);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  o0 = fmaf(hh[j], head[256 + c0 + j], o0);
                  o1 = fmaf(hh[j], head[384 + c0 + j], o1);
                  o2 = fmaf(hh[j], head[512 + c0 + j], o2);
                  if (kSave) mw |= (hh[j] > 0.f) ? (1u << j) : 0u;
                }
                if (kSave) {
                  if (ci) hmw[1] = mw; else hmw[0] = mw;
#pragma unroll
                  for (int c = 0; c < 4; ++c) {           // stage in this warp's own block 2h of the (now free) activation tile
                    uint4 o;
                    o.x = pack_bf16x2(hh[8 * c + 0], hh[8 * c + 1]);
                    o.y = pack_bf16x2(hh[8 * c + 2], hh[8 * c + 3]);
                    o.z = pack_bf16x2(hh[8 * c + 4], hh[8 * c + 5]);
                    o.w = pack_bf16x2(hh[8 * c + 6], hh[8 * c + 7]);
                    *reinterpret_cast<uint4*>(at + 2 * ch * 16384 + sw128_off(m, 32 * ci + 8 * c)) = o;
                  }
                }
              }
              if (kSave && valid) *reinterpret_cast<uint2*>(a.hvmask + row * 4 + 2 * ch) = make_uint2(hmw[0], hmw[1]);
             }
              tc_fence_before();
              // shared mode: combine the two column halves of the heads: half 1 hands its partial sums to half 0
              if (kShared) {
                if (h == 1) xchg_of(t)[lane] = make_float4(o0, o1, o2, sigma);
                named_bar_sync(1, 256);
              }
              if (!kShared || h == 0) {
                const float4 p = kShared ? xchg_of(t)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid)
                  *reinterpret_cast<float4*>(a.out + row * 4) =
                      make_float4(o0 + p.x + head[641], o1 + p.y + head[642], o2 + p.z + head[643], sigma + p.w + head[640]);
              }
              if (kPipe && unit + unit_step < num_units && (t == 0 || unit + unit_step < a.n_full_units))
                act_arrive(t);                                                // next unit's layer 0 of this tile may start
              if (kSave) {
#pragma unroll 1
                for (int ch = ch_lo; ch < ch_hi; ++ch) store_blocks(a.hv_img + tile * 32768, ch, at, 2 * ch, 1);
                if (!kShared || h == 0) {
                  // encoded view direction of this sample as bf16 (operand of view_linear's direction columns), own block 1
                  const float* dep = a.de + vrow * 32;
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (c < 4 && valid) {
                      const float4 u0 = __ldg(reinterpret_cast<const float4*>(dep) + 2 * c);
                      const float4 u1 = __ldg(reinterpret_cast<const float4*>(dep) + 2 * c + 1);
                      o.x = pack_bf16x2(u0.x, u0.y); o.y = pack_bf16x2(u0.z, u0.w);
                      o.z = pack_bf16x2(u1.x, u1.y); o.w = pack_bf16x2(u1.z, u1.w);
                    }
                    *reinterpret_cast<uint4*>(at + 1 * 16384 + sw128_off(m, 8 * c)) = o;
                  }
                  store_blocks(a.de16_img + tile * 16384, 0, at, 1, 1);
                }
              }
            }
          }
          if (kPipe && g == 5 && unit + unit_step < num_units) {
            // layer 5 was the x-tile's last reader (its MMAs completed before the accumulator barrier fired):
            // encode the next unit's points now, under the wait for the layer-6 accumulators
            int64_t tile0n;
            if (h < unit_tiles(unit + unit_step, tile0n)) {
              fwd_prologue(a, (tile0n + h) * kTileM + m, m, xt_own);
              fence_proxy_async();
              if (kSave) store_blocks(a.xenc_img + (tile0n + h) * 16384, 0, xt_own, 0, 1);
            }
          }
        }
      } else {
        // ------------------------------ dgrad chain ------------------------------
        float dsig0, dsig1;                                // d_sigma of this thread's row in tile A / tile B
        {
          // prologue (own tile): d_hv_pre = (d_rgb . W_rgb) * [hv > 0]  (reference autograd of model.py:73-75)
          const int64_t row = (tile0 + h) * kTileM + m;
          const bool valid = own && row < a.M;
          uint8_t* at = smem + kOffAct + h * kActBytes;
          const float4 dr = valid ? __ldg(reinterpret_cast<const float4*>(a.d_raw) + row) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (kCopy) { cp_wait_free(0); cp_wait_free(1); }     // the previous unit's last tensors have left shared memory
          if (kShared) {
            if (1 - h < ntl) act_arrive(1 - h);
            // blocks 0-1 of tile 1 were copied out by the group-0 warps in the previous unit's last epilogue
            if (!kCopy) named_bar_sync(1, 256);
          }
          if (own) {
          uint4 mw = valid ? __ldg(reinterpret_cast<const uint4*>(a.hvmask) + row) : make_uint4(0u, 0u, 0u, 0u);
          const uint32_t mws[4] = {mw.x, mw.y, mw.z, mw.w};
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int n = 8 * c + j;
              const float gg = fmaf(dr.x, head[256 + n], fmaf(dr.y, head[384 + n], dr.z * head[512 + n]));
              v[j] = ((mws[n >> 5] >> (n & 31)) & 1u) ? gg : 0.f;
            }
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            const int k = 8 * c;
            *reinterpret_cast<uint4*>(at + (k >> 6) * 16384 + sw128_off(m, k & 63)) = o;
          }
          fence_proxy_async();
          // copy d_hv out BEFORE releasing the MMAs: in the first epilogue warp (q,0) overwrites blocks 0-1 of
          // both tiles, and for tile 1 that is not the warp that copies here
          if (kShared && !kCopy) store_blocks(a.dhv_img + (tile0 + h) * 32768, 0, at, 0, 2);
          act_arrive(h);
          }
          if (kCopy) { cp_request(0); if (ntl > 1) cp_request(1); }   // d_hv of the unit's tiles (every warp arrives on each barrier)
          if (!kShared && own) store_blocks(a.dhv_img + (tile0 + h) * 32768, 0, at, 0, 2);
        }
        {
          const int64_t r0 = tile0 * kTileM + m, r1 = r0 + kTileM;
          dsig0 = r0 < a.M ? __ldg(a.d_raw + r0 * 4 + 3) : 0.f;
          dsig1 = (ntl > 1 && r1 < a.M) ? __ldg(a.d_raw + r1 * 4 + 3) : 0.f;
        }
        for (int g = 0; g < kNumGemms; ++g) {
          // g = 0: d_bott (no mask) | g = 1: d_h7 (+ sigma term, mask 7) | g >= 2: d_pre_{8-g} (mask 8-g)
          const int ml = 8 - g;                                    // mask layer (g >= 1)
          const int dst = (g == 0) ? 8 : ml;                        // dpre slot: 8 = d_bott, else layer index
#pragma unroll 1
          for (int t = t_lo; t < t_hi && t < ntl; ++t) {
            const float dsig = t ? dsig1 : dsig0;
            const int64_t tile = tile0 + t;
            const int64_t row = tile * kTileM + m;
            const bool valid = row < a.M;
            uint8_t* at = smem + kOffAct + t * kActBytes;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256;
            uint4 m0 = make_uint4(~0u, ~0u, ~0u, ~0u), m1 = m0;       // masks of the column halves this thread handles
            if (g >= 1) {                                          // prefetched before the accumulator wait
              const uint4* mp = reinterpret_cast<const uint4*>(a.mask + ((int64_t)ml * a.M + (valid ? row : 0)) * 8);
              m0 = __ldg(mp + ch_lo);
              if (!kShared) m1 = __ldg(mp + 1);
            }
            DBG_T(t0_);
            mbar_wait(bar_acc(t), (acc_ph >> t) & 1u, 400 + t);
            acc_ph ^= 1u << t;
            tc_fence_after();
            if (kCopy) cp_wait_free(t);
            TRACE(tw, tr_ok, 4 + 8 * t);
            DBG_T(t1_);
            DBG_ACC(0, t1_ - t0_);
            // one 32-column chunk: (+ sigma term) -> ReLU mask -> bf16 -> swizzled store
            auto bchunk = [&](const uint32_t (&r)[32], int c0, uint32_t mw, bool with_sigma) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(r[j]);
                if (with_sigma) x = fmaf(dsig, head[c0 + j], x);   // + d_sigma * w_sigma  (model.py:69)
                v[j] = ((mw >> j) & 1u) ? x : 0.f;
              }
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint4 o;
                o.x = pack_bf16x2(v[8 * c + 0], v[8 * c + 1]); o.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
                o.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]); o.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
                const int k = c0 + 8 * c;
                *reinterpret_cast<uint4*>(at + (k >> 6) * 16384 + sw128_off(m, k & 63)) = o;
              }
            };
            auto bpass = [&](bool with_sigma, int ch, const uint4& mq) {   // two register buffers, as in the forward epilogue
              const uint32_t mws[4] = {mq.x, mq.y, mq.z, mq.w};
              uint32_t ra[32], rb[32];
              const int cb = 128 * ch;
              tmem_ld32(taddr + cb, ra);
#pragma unroll
              for (int c0 = 0; c0 < 128; c0 += 64) {
                tmem_ld_wait();
                tmem_ld32(taddr + cb + c0 + 32, rb);
                bchunk(ra, cb + c0, mws[c0 >> 5], with_sigma);
                tmem_ld_wait();
                if (c0 + 64 < 128) tmem_ld32(taddr + cb + c0 + 64, ra);
                bchunk(rb, cb + c0 + 32, mws[(c0 >> 5) + 1], with_sigma);
              }
            };
#pragma unroll 1
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
              const uint4 mq = (ch == ch_lo) ? m0 : m1;
              if (g == 1) bpass(true, ch, mq); else bpass(false, ch, mq);
            }
            tc_fence_before();
            fence_proxy_async();
            if (g < kNumGemms - 1) act_arrive(t);
            if (kSplit && g == 0 && t == 0 && unit == unit0 && warp == kFirstComputeWarp && lane == 0) mbar_arrive(bar_skew);   // start tile B's chain
            TRACE(tw, tr_ok, 5 + 8 * t);
            DBG_T(t2_);
            DBG_ACC(1, t2_ - t1_);
            DBG_ACC(3, 1);
            if (kCopy) cp_request(t);
#pragma unroll 1
            for (int ch = ch_lo; ch < ch_hi && !kCopy; ++ch)
              store_blocks(a.dpre_img + ((int64_t)dst * ntiles + tile) * 65536, 2 * ch, at, 2 * ch, 2);
            TRACE(tw, tr_ok, 6 + 8 * t);
            DBG_T(t3_);
            DBG_ACC(2, t3_ - t2_);
          }
        }
      }
    }
  }
#ifdef NERF_DBG_TIMING
  if (blockIdx.x == 0 && warp == kFirstComputeWarp && lane == 0 && g_dbg_clk[3] > 0) {
    const double n = (double)g_dbg_clk[3];
    printf("DBG kBwd=%d kSave=%d kCtas=%d: per tile-layer epilogue (warp 3, CTA 0): wait_acc %.0f clk, critical %.0f clk, copy-out %.0f clk, n=%.0f\n",
           (int)kBwd, (int)kSave, kCtas, g_dbg_clk[0] / n, g_dbg_clk[1] / n, g_dbg_clk[2] / n, n);
    printf("DBG issuer A (CTA 0): per slot: weight wait %.0f clk, activation wait %.0f clk (x slots per layer ~8.7), slots=%.0f\n",
           g_dbg_clk[4] / (double)g_dbg_clk[6], g_dbg_clk[5] / (double)g_dbg_clk[6], (double)g_dbg_clk[6]);
    g_dbg_clk[0] = g_dbg_clk[1] = g_dbg_clk[2] = g_dbg_clk[3] = g_dbg_clk[4] = g_dbg_clk[5] = g_dbg_clk[6] = 0;
  }
#endif
#ifdef NERF_DBG_TRACE
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0 && g_trace_n[0] > 0) {
    long long t0 = g_trace[0][0] >> 4;
    printf("TRACE kBwd=%d kSave=%d kSplit=%d kShared=%d\n", (int)kBwd, (int)kSave, (int)kSplit, (int)kShared);
    for (int w = 0; w < 4; ++w) {
      for (int i = 0; i < g_trace_n[w]; ++i) printf("T %d %d %lld\n", w, (int)(g_trace[w][i] & 15), (g_trace[w][i] >> 4) - t0);
      g_trace_n[w] = 0;
    }
  }
#endif
  __syncthreads();
  if (kCtas == 2) cluster_sync();                     // the peer may still be signalling this CTA's barriers
  if (warp == 1) {
    if (kCtas == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

size_t mlp_tc_workspace_bytes(int64_t M, int save) { return ws_layout(M, save).total; }

static int num_ctas() {         // NERF_TC_CTAS_RT=1|2 overrides the compiled default (kernel-variant sweeps)
  static int n = [] {
    const char* e = getenv("NERF_TC_CTAS_RT");
    const int v = e ? atoi(e) : kCtasDefault;
    return v == 2 ? 2 : 1;
  }();
  return n;
}
template <bool kBwd, bool kSave, int kCtas, bool kSigmaOnly = false>
static int launch_tc_impl(const TcArgs& a, cudaStream_t st) {
  static DeviceOnce attr_done;                      // per kernel instantiation AND per device
  DeviceProps dp;
  int rc = current_device(&dp);
  if (rc) return rc;
  if (attr_done.needed(dp.ordinal)) {
    NERF_CHECK_ARG(dp.sm_major == 10, "libnerf_b200 needs an sm_100 device (found sm_%d%d); there is no fallback", dp.sm_major, dp.sm_minor);
    NERF_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<kBwd, kSave, kCtas, kSigmaOnly>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done.mark(dp.ordinal);
  }
  const int max_units = dp.sm_count / kCtas;                  // CTAs (or CTA pairs) that run at once
  const int units_all = (kCtas == 2) ? (a.num_pairs + 1) / 2 : a.num_pairs;   // as full units (two tiles per CTA)
  // the last partial wave: if at most half of the CTAs would get a (two-tile) unit, hand out one-tile half units instead
  const int rem = units_all % max_units;
  TcArgs ah = a;
  static const bool half_ok = [] { const char* e = getenv("NERF_TC_HALF_UNITS"); return !(e && e[0] == '0'); }();
  if (half_ok && rem > 0 && 2 * rem <= max_units) { ah.n_full_units = units_all - rem; ah.n_units = units_all + rem; }
  else { ah.n_full_units = units_all; ah.n_units = units_all; }
  const int units = ah.n_units;
  const int grid = (units < max_units ? units : max_units) * kCtas;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)tc_threads<kBwd || kSave>());
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (kCtas == 2) ? 1 : 0;
  NERF_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_kernel<kBwd, kSave, kCtas, kSigmaOnly>, ah));
  NERF_LAUNCH_CHECK(kBwd ? "mlp_tc_kernel<dgrad>" : "mlp_tc_kernel<fwd>");
  return 0;
}
template <bool kBwd, bool kSave, bool kSigmaOnly = false>
static int launch_tc(const TcArgs& a, cudaStream_t st) {
  return num_ctas() == 2 ? launch_tc_impl<kBwd, kSave, 2, kSigmaOnly>(a, st) : launch_tc_impl<kBwd, kSave, 1, kSigmaOnly>(a, st);
}

static void fill_saved(TcArgs& a, void* ws, const WsLayout& L) {
  uint8_t* b = (uint8_t*)ws;
  a.act_img = b + L.act; a.hv_img = b + L.hv; a.xenc_img = b + L.xenc; a.de16_img = b + L.de16;
  a.mask = (uint32_t*)(b + L.mask); a.hvmask = (uint32_t*)(b + L.hvmask);
  a.de = (const float*)(b + L.de);
  a.dpre_img = b + L.dpre; a.dhv_img = b + L.dhv;
  a.Mp = L.Mp;
}

int mlp_tc_forward(const float* rays_o, const float* rays_d, const float* z_vals, int R, int S, float coord_scale,
                   const float* x_enc, const float* d_enc, int64_t M, const float* params, const void* packed,
                   float* out, void* ws, size_t ws_bytes, int save, cudaStream_t st) {
  const bool sigma_only = (save == 2);                  // density-only inference forward (see mlp_tc_kernel)
  if (sigma_only) save = 0;
  const WsLayout L = ws_layout(M, save);
  NERF_CHECK_ARG(ws_bytes >= L.total, "mlp tc forward: workspace too small (%zu < %zu)", ws_bytes, L.total);
  NERF_CHECK_ARG((((uintptr_t)packed | (uintptr_t)out | (uintptr_t)ws) & 15) == 0, "mlp tc forward: packed/out/workspace must be 16-byte aligned");
  int rc = upload_schedule();
  if (rc) return rc;
  float* vb = (float*)((uint8_t*)ws + L.vb);
  float* de = save ? (float*)((uint8_t*)ws + L.de) : nullptr;
  const int64_t nvb = (x_enc != nullptr) ? M : R;
  if (!sigma_only) {                                    // the view branch is not evaluated in density-only mode
    view_bias_kernel<<<(unsigned)ceil_div(nvb, kVbRows), 128, 0, st>>>(rays_d, d_enc, nvb, params, vb, de);
    NERF_LAUNCH_CHECK("view_bias_kernel");
  }
  TcArgs a{};
  a.rays_o = rays_o; a.rays_d = rays_d; a.z_vals = z_vals; a.S = S; a.coord_scale = coord_scale;
  a.x_enc = x_enc; a.M = M; a.packed = (const uint8_t*)packed; a.params = params;
  a.vb = vb; a.vb_div = (x_enc != nullptr) ? 1 : S;
  a.out = out;
  if (save) fill_saved(a, ws, L);
  a.num_pairs = (int)(L.Mp / (2 * kTileM));   // the whole padded tile range: every tile image the wgrad kernel reads is written (single-CTA mode too)
  if (sigma_only) return launch_tc<false, false, true>(a, st);
  return save ? launch_tc<false, true>(a, st) : launch_tc<false, false>(a, st);
}

}  // namespace nerf
#include "nerf_mlp_bwd_fused.cuh"
namespace nerf {

// NERF_BWD_FUSED=0 selects the two-kernel backward (dgrad chain, then wgrad) for A/B measurements
static bool bwd_fused_enabled() {
  static const bool on = [] { const char* e = getenv("NERF_BWD_FUSED"); return !(e && e[0] == '0'); }();
  return on;
}

int mlp_tc_backward(const float* d_raw, int64_t M, int rows_per_dir, const float* params, const void* packed,
                    float* grads, void* ws, size_t ws_bytes, int stage, cudaStream_t st) {
  const WsLayout L = ws_layout(M, 1);
  NERF_CHECK_ARG(ws_bytes >= L.total, "mlp tc backward: workspace too small (%zu < %zu)", ws_bytes, L.total);
  NERF_CHECK_ARG(rows_per_dir >= 1, "mlp tc backward: rows_per_dir must be >= 1");
  NERF_CHECK_ARG((((uintptr_t)d_raw | (uintptr_t)ws) & 15) == 0, "mlp tc backward: d_raw/workspace must be 16-byte aligned");
  int rc = upload_schedule();
  if (rc) return rc;
  TcArgs a{};
  a.d_raw = d_raw; a.M = M; a.packed = (const uint8_t*)packed; a.params = params;
  fill_saved(a, ws, L);
  a.num_pairs = (int)(L.Mp / (2 * kTileM));   // the whole padded tile range: every tile image the wgrad kernel reads is written (single-CTA mode too)
  if (stage == NERF_BWD_ALL && bwd_fused_enabled()) {
    // one launch: dgrad chain + tensor-core weight gradients; the two tiny heads (rgb, sigma) follow on CUDA cores
    if ((rc = mlp_tc_heads_fork(st))) return rc;                 // the heads run beside the big kernel (side stream)
    if ((rc = launch_bwd_fused(a, ws, L, grads, st))) return rc;
    return mlp_tc_heads_wgrad(ws, L, d_raw, M, grads, st);
  }
  if (stage != NERF_BWD_WGRAD && (rc = launch_tc<true, false>(a, st))) return rc;     // d(pre-activations) -> workspace
  if (stage == NERF_BWD_DGRAD) return 0;

  return mlp_tc_wgrad(ws, L, d_raw, M, rows_per_dir, grads, st);            // weight / bias gradients
}

}  // namespace nerf
