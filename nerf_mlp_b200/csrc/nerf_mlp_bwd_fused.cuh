// Fused backward of the NeRF MLP on tcgen05: the dgrad chain AND the weight gradients in ONE launch.
// (Included at the end of nerf_mlp_tc.cu: shares its weight-slot schedule, TcArgs and epilogue helpers.)
//
// Why: run as two kernels, the dgrad chain leaves the tensor pipe ~50 % idle (MMA -> drain -> epilogue -> MMA is one
// dependent chain per tile) and writes 4.9 KB of d(pre-activation) per sample row to HBM, which the wgrad kernel then
// streams back together with the saved activations (9.7 KB per row, 128 FLOP/B: it can never exceed half of the
// tensor peak from HBM).  Here every CTA pair does both at once:
//
//   dgrad role   one 128-row tile per CTA (M = 256 per pair, cta_group::2), the 9-GEMM chain over the transposed
//                weight image exactly as mlp_tc_kernel<dgrad>; all 8 epilogue warps share the tile's epilogue; each
//                d(pre-activation) tile image is copied out to the workspace and PUBLISHED with a per-(tensor, tile)
//                counter (release at gpu scope).
//   wgrad role   the pair is stationary on ONE weight-gradient job (a layer): its dW block lives in the other 256
//                TMEM columns of both CTAs (128 output features each) for the whole kernel.  A loader lane waits for
//                the counter of the next tile (acquire), then bulk-copies 64-row chunks of dY (just written by some
//                other pair: L2 hits) and X (saved by the forward) into a 3-stage ring; the wgrad issuer's MMAs
//                (M = 256, N = 256, both operands MN-major) fill the tensor-pipe time the dgrad chain leaves idle.
//                The pair's K-slab is every npairs-th tile pair, in the order the dgrad roles produce them.
//   epilogue     after its last dgrad unit the 8 epilogue warps reduce the pair's dW block into the flat fp32
//                gradient (red.global.add.v4.f32, split-K over the job's pairs); 2 warps sum the dY chunk columns for db.
//
// The dgrad roles never wait for a wgrad role, all CTAs are co-resident (one per SM), so the flag waits cannot
// deadlock; like every mbarrier wait they are bounded and trap instead of hanging the GPU.
#pragma once

namespace nerf {

constexpr int kFzThreads = 512;                   // 4 control warps, 8 epilogue warps, 2 db warps, 2 copy-out / publish warps
#ifndef NERF_FZ_RING
#define NERF_FZ_RING 8
#endif
#ifndef NERF_FZ_STAGES
#define NERF_FZ_STAGES 3
#endif
constexpr int kFzRing = NERF_FZ_RING;             // dgrad weight ring: half-slots of 8 KB per CTA (a layer is 8 of them)
constexpr int kFzSlotK = kSlotBytes / 2;
constexpr int kFzStages = NERF_FZ_STAGES;         // wgrad operand ring
constexpr int kFzStageBytes = 32768;              // A: 2 feature blocks x [64 rows][128 B] (16 KB) | B: the same
constexpr int kFzOffA = 0, kFzOffB = 16384;
constexpr int kFzBiasWarps = 2;
constexpr int kFzCopyWarps = 2;
#ifndef NERF_FZ_PREFETCH
#define NERF_FZ_PREFETCH 0     // measured: 0.625 ms without, 0.641 ms with distance 6 (1024 rays x 192): the loads are not latency-bound
#endif
constexpr int kFzPrefetch = NERF_FZ_PREFETCH;     // L2 prefetch distance of the X operand, in 64-row chunks
constexpr int kFzKindsC = 10;
// shared-memory map; kWgOnly = the same kernel run as a stand-alone weight-gradient kernel (no dgrad role: the
// activation tile and the weight ring give their space to a deeper operand ring)
struct FzL {
  static constexpr bool kWgOnly = false;
  static constexpr int kStages = kWgOnly ? 6 : kFzStages;
  static constexpr int kOffAct = 0;                                        // [128 x 256] bf16, SW128 K-blocks of 64
  static constexpr int kOffRing = kWgOnly ? 0 : kOffAct + kActBytes;
  static constexpr int kOffWg = kWgOnly ? 0 : kOffRing + kFzRing * kFzSlotK;
  static constexpr int kOffHead = kOffWg + kStages * kFzStageBytes;
  static constexpr int kOffBar = kOffHead + kHeadFloats * 4;
  static constexpr int kNumBars = 2 * kFzRing + 2 * kStages + 7;
  static constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr int kOffUnit = kOffTmemPtr + 8;                         // int[2]: unit index handed to the roles (double-buffered)
  static constexpr int kSmemBytes = kOffUnit + 8;
  static_assert(kOffWg % 1024 == 0, "SW128 operand stages need 1024-byte alignment");
  static_assert(kSmemBytes <= 232448, "exceeds 227 KB of shared memory");
};
static_assert(kNK == 2, "the fused backward assumes 16 KB weight slots (two K-steps)");

constexpr int kFzKinds = 10;                      // published tensors: d(pre-act) of layers 0..7, d_bottleneck (8), d_hv (9)
constexpr int kFzMaxJobs = 12;
struct FzJob {
  const uint8_t* A; const uint8_t* B;   // dY / X tile images
  float* out; float* bias;              // dW block (column offset applied), db (or nullptr)
  int a_tile_bytes, b_tile_bytes;
  int a_fb[2], b_fb[2];                 // first feature block loaded by CTA rank 0 / 1 (A: 2 blocks; B: b_nfb blocks)
  int b_nfb;                            // 2: N = 256 (rank r loads input features 128r..); 1: N = 128, both ranks load the same block
  int ld_out, ncols;                    // row stride of dW, valid columns (D column c <-> dW column c)
  int out_row0[2], out_rows[2];         // first output feature of rank r's 128 TMEM lanes; rows it reduces (128, or 0 = padding half)
  int kind, need;                       // flag row of the dY tensor; publishing warps per tile
  int first_pair, npairs;
  int single_reader;                    // this job is the only reader of its dY tensor (its chunks may be discarded from L2 after use)
};
struct FzArgs {
  TcArgs t;
  FzJob jobs[kFzMaxJobs];
  int njobs;
  uint32_t* flags;                      // [kFzKinds][ntiles], zeroed before the launch
  int ntiles;
  int dbg;                              // development (NERF_FZ_MODE): 1 = no dgrad role, 2 = no wgrad role, 4 = print role timings
  float* grads;                         // flat fp32 gradient buffer (bias gradients are reduced straight from the copy-out)
  int db_off[kFzKinds];                 // offset of the bias gradient fed by published tensor `kind`
  uint32_t* unit_ctr;                   // next dgrad unit (tile pair) to hand out; zeroed with the flags
  int stagger_clk;                      // start offset between consecutive pairs' first units (clocks)
  uint32_t* progress;                   // [pairs]: unit index the pair's wgrad role has reached (0x7fffffff when done)
  int window;                           // dgrad units may start at most this far ahead of the slowest wgrad role (0 = unthrottled)
  int prefetch;                         // L2 prefetch distance of the X operand (saved by the forward, comes from HBM), in 64-row chunks; 0 = off
  int discard;                          // 1: single-reader dY chunks are dropped from L2 after their wgrad MMAs (no write-back)
};
__device__ long long g_fz_t[148][13];

// Cross-proxy fence for GLOBAL memory: the dY tiles are written with generic-proxy stores (another SM's copy warps) and
// read by bulk copies (async proxy).  The ordering against the writer is the ld.acquire.gpu on the publish counter; this
// fence only has to make the async proxy see what the generic proxy has observed.  `.global` compiles to a single
// FENCE.VIEW.ASYNC.G -- the unqualified form adds a MEMBAR.ALL.GPU (~1 000+ clk), which with the unit throttle (small
// batches of newly published tiles, so one fence every few chunks) sat on the wgrad loaders' critical path.
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// the d(pre-activation) tile was written with generic-proxy stores by another SM; the loads that follow are async-proxy
__device__ __forceinline__ void wait_flag(const uint32_t* f, uint32_t need, int tag) {
  if (ld_acquire_gpu(f) < need) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(f) < need) {
      __nanosleep(64);
      if (clock64() - t0 > (1ll << 32)) {
        printf("nerf_b200: fused backward: flag wait timeout tag=%d block=%d have=%u need=%u\n", tag, (int)blockIdx.x, ld_acquire_gpu(f), need);
        __trap();
      }
    }
  }
  fence_proxy_async_all();
}
__device__ __forceinline__ void red_add_v4f(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kFzThreads, 1) bwd_fused_kernel(const __grid_constant__ FzArgs fa) {
  using L = FzL;
  constexpr bool kWgOnly = false;
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcArgs& a = fa.t;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + L::kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kFzRing + s); };
  auto bar_wfull = [&](int s) { return bar0 + 8u * (2 * kFzRing + s); };
  auto bar_wempty = [&](int s) { return bar0 + 8u * (2 * kFzRing + L::kStages + s); };
  const uint32_t bar_act = bar0 + 8u * (2 * kFzRing + 2 * L::kStages);   // activations of the tile ready (epilogue -> MMA), at the leader
  const uint32_t bar_acc = bar_act + 8;                                   // dgrad accumulator ready (MMA -> epilogue)
  const uint32_t bar_wacc = bar_act + 16;                                 // dW block complete
  const uint32_t bar_cp = bar_act + 24;                                   // tile (or d_hv) written to shared memory: copy it out (epilogue -> publisher)
  const uint32_t bar_cpfree = bar_act + 32;                               // the copy warps have read the tile: it may be overwritten
  auto bar_unit = [&](int k) { return bar_act + 40 + 8u * (k & 1); };     // unit index k (k-th unit of this pair) is in its box
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::kOffTmemPtr);
  float* head = reinterpret_cast<float*>(smem + L::kOffHead);

  // wgrad job / K-slab of this pair
  int ji = 0;
  for (int j = 1; j < fa.njobs; ++j)
    if (pair >= fa.jobs[j].first_pair) ji = j;
  const FzJob& job = fa.jobs[ji];
  const int slab = pair - job.first_pair;
  const int all_units = a.num_pairs;                         // tile pairs: unit u = tiles 2u (leader CTA) and 2u+1 (peer)
  const int num_units = (kWgOnly || (fa.dbg & 1)) ? 0 : all_units;        // units of the dgrad role
  const int wg_units = (!(fa.dbg & 2) && slab < job.npairs && slab < all_units) ? (all_units - 1 - slab) / job.npairs + 1 : 0;
  const long long t_begin = clock64();
  long long w_a = 0, w_b = 0;                                // development counters (per role: two kinds of waiting)
  long long w_thr = 0;                                       // leader's producer: time units were held back by the throttle
  const int wg_chunks = wg_units * 4;                        // 64-row chunks: two per tile

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kFzRing; ++s) { mbar_init(bar_full(s), rank == 0 ? 2 : 1); mbar_init(bar_empty(s), 1); }
    for (int s = 0; s < L::kStages; ++s) { mbar_init(bar_wfull(s), rank == 0 ? 2 : 1); mbar_init(bar_wempty(s), 1 + kFzBiasWarps); }
    mbar_init(bar_act, 256 * 2);
    mbar_init(bar_acc, 1);
    mbar_init(bar_wacc, 1);
    mbar_init(bar_cp, 8);
    mbar_init(bar_cpfree, kFzCopyWarps);
    mbar_init(bar_unit(0), 1);
    mbar_init(bar_unit(1), 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc2(sbase + L::kOffTmemPtr, 512); tmem_relinquish2(); }
  if (warp >= 4 && warp < 12) {
    const int tid = threadIdx.x - 128;                   // 0..255
    for (int i = tid; i < 256; i += 256) head[i] = a.params[w_off(L_SIGMA) + i];
    for (int i = tid; i < 384; i += 256) head[256 + i] = a.params[w_off(L_RGB) + i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                          // both CTAs' barriers exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int slot0 = c_nslots_fwd, nslots = c_nslots_bwd;
  const int64_t ntiles = fa.ntiles;

  // ---- dynamic distribution of the dgrad units (tile pairs) ----
  // The leader's weight producer draws the pair's next unit from a global counter and posts it to both CTAs; every
  // dgrad-side role picks it up from its CTA's box.  Units are therefore STARTED in increasing order over time (the
  // wgrad roles consume each tensor's tiles in exactly that order), the first units are staggered so that the pairs
  // do not all produce the same layer's tiles at the same moment (the wgrad roles of a layer then see a steady
  // stream instead of a burst per round, and read dY while it is still in L2), and a pair whose wgrad job is heavy
  // simply draws fewer units.
  volatile int* unit_box = reinterpret_cast<volatile int*>(smem + L::kOffUnit);
  auto unit_get = [&](int k) -> int {
    if (kWgOnly) return -1;
    mbar_wait_cluster(bar_unit(k), (uint32_t)(k >> 1) & 1u, 970);
    return unit_box[k & 1];
  };
  auto unit_draw = [&](int k) -> int {                       // leader, warp 0, lane 0: next unit of the global sequence
    if (kWgOnly) return -1;
    if (k == 0 && fa.stagger_clk > 0) {
      const long long wait_clk = (long long)pair * fa.stagger_clk;
      while (clock64() - t_begin < wait_clk) __nanosleep(256);
    }
    const int u = (int)atomicAdd(fa.unit_ctr, 1u);
    return u >= num_units ? -1 : u;
  };
  auto unit_post = [&](int k, int u) {                       // leader, warp 0, lane 0: hand unit u to both CTAs
    if (kWgOnly) return;
    unit_box[k & 1] = u;
    const uint32_t rbox = mapa(sbase + L::kOffUnit + 4u * (k & 1), 1);
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(rbox), "r"(u) : "memory");
    mbar_arrive(bar_unit(k));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(bar_unit(k), 1)) : "memory");
  };

  if (warp < 4) {
    if (warp == 0) {
      // ================= dgrad weight producer (+ unit hand-out and throttle at the leader) =================
      // Throttle: the dgrad chains outrun the wgrad roles (they would be done at ~80 % of the kernel), and dY that is
      // not consumed soon leaves L2 and is re-read from HBM.  The leader therefore starts a unit only when the slowest
      // wgrad role is within `window` units of it.  The whole warp polls the roles' progress words (one L2 round
      // trip per poll).  No deadlock: units are handed out in increasing order and a unit, once started, never waits
      // for a wgrad role, so every unit below (slowest + window) gets published.
      uint32_t g = 0;
      int prog_min = 0;                                        // last known minimum of the roles' progress (monotone)
      for (int k = 0;; ++k) {
        if (rank == 0) {
          int u = 0;
          if (lane == 0) u = unit_draw(k);
          u = __shfl_sync(0xffffffffu, u, 0);
          if (u >= 0 && fa.window > 0 && u > prog_min + fa.window) {
            const long long t0 = clock64();
            for (;;) {
              uint32_t mn = 0x3fffffffu;
              for (int i = lane; i < npairs; i += 32) {
                const uint32_t v = *reinterpret_cast<volatile const uint32_t*>(fa.progress + i);
                mn = v < mn ? v : mn;
              }
              mn = __reduce_min_sync(0xffffffffu, mn);
              prog_min = (int)mn;
              if (u <= prog_min + fa.window) break;
              __nanosleep(200);
              if (clock64() - t0 > (1ll << 32)) {
                if (lane == 0) printf("nerf_b200: fused backward: throttle wait timeout block=%d unit=%d slowest=%d\n", (int)blockIdx.x, u, prog_min);
                __trap();
              }
            }
            w_thr += clock64() - t0;
          }
          if (lane == 0) unit_post(k, u);
        }
        __syncwarp();
        int unit = 0;
        if (lane == 0) unit = unit_get(k);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit < 0) break;
        if (lane == 0) {
          for (int i = 0; i < nslots; ++i, ++g) {
            const uint32_t s = g % kFzRing, ph = (g / kFzRing) & 1;
            const uint2 rec = *reinterpret_cast<const uint2*>(&c_slots[slot0 + i]);      // goff, bytes
            const uint32_t bytes = rec.y >> 1;                                            // this CTA's N-half of the slot
            mbar_wait(bar_empty(s), ph ^ 1, 100 + (int)s);
            mbar_expect_tx(bar_full(s), bytes);
            bulk_g2s(sbase + L::kOffRing + s * kFzSlotK, a.packed + rec.x + rank * bytes, bytes, bar_full(s));
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      if (rank != 0) {
        // peer: relay "my half of slot s landed" to the leader, whose MMAs read both halves
        uint32_t s = 0, ph = 0;
        for (int k = 0; unit_get(k) >= 0; ++k) {
          for (int i = 0; i < nslots; ++i) {
            mbar_wait(bar_full(s), ph, 250 + (int)s);
            if (lane == 0) mbar_arrive_cluster(mapa(bar_full(s), 0));
            __syncwarp();
            if (++s == kFzRing) { s = 0; ph ^= 1; }
          }
        }
      } else {
        // ================= dgrad MMA issuer (leader CTA; M = 256 over the pair) =================
        constexpr uint32_t kHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t kHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);
        constexpr uint32_t kHiSw32 = (256u >> 4) | (1u << 14) | (6u << 29);
        const uint32_t a_lo_act = ((sbase + L::kOffAct) >> 4) | (1u << 16);
        const uint32_t b_lo0 = ((sbase + L::kOffRing) >> 4) | (1u << 16);
        constexpr uint32_t kIdesc256 = make_idesc_bf16(256, 256);
        uint32_t s = 0, ph = 0, act_ph = 0;
        for (int k = 0; unit_get(k) >= 0; ++k) {
          for (int i = 0; i < nslots; ++i) {
            const uint4 rec = *reinterpret_cast<const uint4*>(&c_slots[slot0 + i]);
            const uint32_t a_add = rec.z, fl = rec.w;
            { const long long t_ = clock64(); mbar_wait(bar_full(s), ph, 200 + (int)s); w_a += clock64() - t_; }
            if (fl & kFlagFirst) { const long long t_ = clock64(); mbar_wait(bar_act, act_ph, 300); act_ph ^= 1; w_b += clock64() - t_; }
            tc_fence_after();
            const uint32_t a_lo = a_lo_act + a_add;
            const uint32_t b_lo = b_lo0 + s * (kFzSlotK >> 4);
            const uint32_t b_hi = (fl & kFlagNk2) ? kHiSw64 : kHiSw32;
            if (elect_one()) {
              mma_bf16_ss_2cta(tmem_base, ((uint64_t)kHiSw128 << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, kIdesc256,
                               (fl & kFlagFirst) ? 0u : 1u);
              if (fl & kFlagNk2)
                mma_bf16_ss_2cta(tmem_base, ((uint64_t)kHiSw128 << 32) | (a_lo + 2), ((uint64_t)b_hi << 32) | (b_lo + 2), kIdesc256, 1u);
              tc_commit_mc2(bar_empty(s), 3);
              if (fl & kFlagLast) tc_commit_mc2(bar_acc, 3);
            }
            __syncwarp();
            if (++s == kFzRing) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp == 2) {
      // ================= wgrad loader: bulk copies of 64-row chunks of the tile images =================
      // The whole warp walks the chunk list: lane 0 issues; all 32 lanes look AHEAD at the publish counters of the next
      // 32 tiles of the slab (one L2 round trip per batch instead of one per tile on the loader's critical path).
      const uint32_t bytes = (uint32_t)(2 + job.b_nfb) * 8192u;
      const int a_fb = job.a_fb[rank], b_fb = job.b_fb[rank];
      const uint32_t* flags = fa.flags + (int64_t)job.kind * ntiles;
      const int slab_tiles = 2 * wg_units;
      auto tile_of = [&](int ti) { return 2 * (int64_t)(slab + (ti >> 1) * job.npairs) + (ti & 1); };   // ti-th tile of the slab
      int ready = kWgOnly ? slab_tiles : 0;                 // tiles [0, ready) of the slab are known to be published
      // L2 policy: X (saved by the forward) is read exactly once -> evict-first, so that it does not push the freshly
      // written dY tiles out of L2 before their wgrad role gets to them; dY is dead after its last reader (layer 5 and
      // d_hv have two readers: normal priority there)
      const uint64_t pol_x = l2_policy_evict_first();
      for (int c = 0; c < wg_chunks; ++c) {
        const int s = c % L::kStages;
        const int ti = c >> 1;
        if (ti >= ready) {
          const long long t0 = clock64();
          for (;;) {
            const int idx = ti + lane;
            const bool ok = idx < slab_tiles && ld_acquire_gpu(flags + tile_of(idx)) >= (uint32_t)job.need;
            const uint32_t okm = __ballot_sync(0xffffffffu, ok);
            const int n = (okm == 0xffffffffu) ? 32 : (__ffs(~okm) - 1);      // leading run of published tiles
            if (n > 0) { ready = ti + n; break; }
            __nanosleep(100);
            if (clock64() - t0 > (1ll << 32)) {
              if (lane == 0) printf("nerf_b200: fused backward: publish-counter wait timeout block=%d kind=%d tile=%lld\n", (int)blockIdx.x, job.kind, (long long)tile_of(ti));
              __trap();
            }
          }
          w_b += clock64() - t0;
          fence_proxy_async_all();                          // the tile was written by generic-proxy stores; the copies below are async-proxy reads
        }
        // L2 prefetch kFzPrefetch chunks ahead, on the LSU path (all 32 lanes, one 128-byte line each per instruction) so
        // that it costs the TMA unit nothing: X always (it comes from HBM, no dependency), dY if already published.
        // The 3-stage ring then only has to cover the L2 latency.
        if (fa.prefetch > 0 && c + fa.prefetch < wg_chunks) {
          const int cp = c + fa.prefetch;
          const int64_t ptile = tile_of(cp >> 1);
          const uint32_t phalf = (uint32_t)(cp & 1) * 8192u;
          for (int fb = 0; fb < job.b_nfb; ++fb) {
            const uint8_t* src = job.B + ptile * job.b_tile_bytes + (b_fb + fb) * 16384 + phalf + lane * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 4096));
          }
        }
        if (lane == 0) {
          if (rank == 0 && (c & 3) == 0)                    // progress of this role, read by the leaders' unit throttle
            *reinterpret_cast<volatile uint32_t*>(fa.progress + pair) = (uint32_t)(slab + (c >> 2) * job.npairs);
          if (c >= L::kStages) { const long long t_ = clock64(); mbar_wait(bar_wempty(s), ((c / L::kStages) - 1) & 1, 600 + s); w_a += clock64() - t_; }
          const int64_t tile = tile_of(ti);
          const uint32_t half = (uint32_t)(c & 1) * 8192u;          // rows 0-63 / 64-127 of the tile
          const uint32_t sa = sbase + L::kOffWg + s * kFzStageBytes + kFzOffA, sb = sa + (kFzOffB - kFzOffA);
          const uint8_t* a_src = job.A + tile * job.a_tile_bytes;
          mbar_expect_tx(bar_wfull(s), bytes);
          for (int fb = 0; fb < 2; ++fb) {
            const uint8_t* src = a_src + (a_fb + fb) * 16384 + half;
            bulk_g2s(sa + fb * 8192, src, 8192, bar_wfull(s));
          }
          for (int fb = 0; fb < job.b_nfb; ++fb) {
            const uint8_t* src = job.B + tile * job.b_tile_bytes + (b_fb + fb) * 16384 + half;
            bulk_g2s_hint(sb + fb * 8192, src, 8192, bar_wfull(s), pol_x);
          }
        }
        __syncwarp();
        // The MMAs of chunk c - kStages have completed (lane 0 waited for its stage above): its dY rows are dead if
        // this job is their only reader.  Dropping the lines from L2 spares the write-back of dY to HBM.
        if (fa.discard && job.single_reader && c >= L::kStages) {
          const int cd = c - L::kStages;
          const uint8_t* d_src = job.A + tile_of(cd >> 1) * job.a_tile_bytes + (uint32_t)(cd & 1) * 8192u + lane * 128;
#pragma unroll
          for (int fb = 0; fb < 2; ++fb) {
            const uint8_t* q = d_src + (a_fb + fb) * 16384;
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(q) : "memory");
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(q + 4096) : "memory");
          }
        }
      }
      if (lane == 0 && rank == 0) *reinterpret_cast<volatile uint32_t*>(fa.progress + pair) = 0x7fffffffu;
    } else {
      if (rank != 0) {
        // peer: relay "my halves of stage s landed"
        for (int c = 0; c < wg_chunks; ++c) {
          const int s = c % L::kStages;
          mbar_wait(bar_wfull(s), (c / L::kStages) & 1, 750 + s);
          if (lane == 0) mbar_arrive_cluster(mapa(bar_wfull(s), 0));
          __syncwarp();
        }
      } else {
        // ================= wgrad MMA issuer: dW[256 out x N in] += dY_chunk^T . X_chunk (K = 64 rows) =================
        constexpr uint32_t kHiMn = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO = 1024 (8-row group), SW128
        const uint32_t lbo = ((uint32_t)(64 * 128) >> 4) << 16;                     // next 64-feature block
        const uint32_t idesc = make_idesc_bf16(256, 128 * job.b_nfb) | (1u << 15) | (1u << 16);   // both operands MN-major
        for (int c = 0; c < wg_chunks; ++c) {
          const int s = c % L::kStages;
          { const long long t_ = clock64(); mbar_wait(bar_wfull(s), (c / L::kStages) & 1, 700 + s); w_a += clock64() - t_; }
          tc_fence_after();
          const uint32_t sa = sbase + L::kOffWg + s * kFzStageBytes + kFzOffA, sb = sa + (kFzOffB - kFzOffA);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t a_lo = ((sa + ks * 2048) >> 4) | lbo;
              const uint32_t b_lo = ((sb + ks * 2048) >> 4) | lbo;
              mma_bf16_ss_2cta(tmem_base + 256, ((uint64_t)kHiMn << 32) | a_lo, ((uint64_t)kHiMn << 32) | b_lo, idesc,
                               (c == 0 && ks == 0) ? 0u : 1u);
            }
            tc_commit_mc2(bar_wempty(s), 3);
            if (c == wg_chunks - 1) tc_commit_mc2(bar_wacc, 3);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 12) {
    // ================= dgrad prologue + epilogue (8 warps share the CTA's tile), copy-out, publish, final dW reduction ===========
    const int h = (warp - 4) >> 2;                    // column half [128h, 128h+128) = feature blocks 2h, 2h+1
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;                      // row within the tile
    uint8_t* at = smem + L::kOffAct;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    // Copy-out and publish are NOT done here: the epilogue warps only tell the publisher warp that the tile (or d_hv)
    // is complete in shared memory (bar_cp) and, before they overwrite it, make sure its bulk store has finished
    // reading (bar_cpfree, normally long since).  A gpu-scope release on these warps cost ~1 800 clk per layer.
    long long ph_b = 0, ph_c = 0, ph_d = 0, ph_e = 0;  // development counters
    uint32_t cpfree_ph = 0u;
    bool cp_pending = false;                          // a copy-out of the activation tile has been requested and not yet awaited
    auto cp_request = [&]() {                         // after this thread's st.shared + fence.proxy.async
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_cp);
      cp_pending = true;
    };
    auto cp_wait_free = [&]() {
      if (cp_pending) {
        const long long t_ = clock64();
        mbar_wait(bar_cpfree, cpfree_ph, 450);
        cpfree_ph ^= 1u;
        cp_pending = false;
        ph_b += clock64() - t_;
      }
    };
    auto act_arrive = [&]() {
      if (rank != 0) mbar_arrive_cluster(mapa(bar_act, 0)); else mbar_arrive(bar_act);
    };
    uint32_t acc_ph = 0u;
    for (int k = 0;; ++k) {
      const int unit = unit_get(k);
      if (unit < 0) break;
      const int64_t tile = 2 * (int64_t)unit + rank;
      const int64_t row = tile * kTileM + m;
      const bool valid = row < a.M;
      const float4 dr = valid ? __ldg(reinterpret_cast<const float4*>(a.d_raw) + row) : make_float4(0.f, 0.f, 0.f, 0.f);
      // prologue: d_hv_pre = (d_rgb . W_rgb) * [hv > 0]  (autograd of model.py:73-75) -> feature blocks 0-1, owned by the h = 0 warps
      cp_wait_free();                                  // the previous unit's last tensor has left shared memory
      if (h == 0) {
        const uint4 mw = valid ? __ldg(reinterpret_cast<const uint4*>(a.hvmask) + row) : make_uint4(0u, 0u, 0u, 0u);
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
      }
      act_arrive();
      cp_request();                                    // d_hv (blocks 0-1) -> workspace
      const float dsig = dr.w;
      for (int g = 0; g < kNumGemmsBwd; ++g) {
        // g = 0: d_bott (no mask) | g = 1: d_h7 (+ sigma term, mask 7) | g >= 2: d_pre_{8-g} (mask 8-g)
        const int ml = 8 - g;
        const int dst = (g == 0) ? 8 : ml;
        uint4 mq = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (g >= 1) mq = __ldg(reinterpret_cast<const uint4*>(a.mask + ((int64_t)ml * a.M + (valid ? row : 0)) * 8) + h);
        { const long long t_ = clock64(); mbar_wait(bar_acc, acc_ph, 400); w_a += clock64() - t_; }
        acc_ph ^= 1u;
        tc_fence_after();
        cp_wait_free();
        const long long p1_ = clock64();
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
        auto bpass = [&](bool with_sigma) {             // two register buffers: the next chunk's tcgen05.ld is in flight
          const uint32_t mws[4] = {mq.x, mq.y, mq.z, mq.w};
          uint32_t ra[32], rb[32];
          const int cb = 128 * h;
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
        if (g == 1) bpass(true); else bpass(false);
        const long long p2_ = clock64();
        tc_fence_before();
        fence_proxy_async();
        if (g < kNumGemmsBwd - 1) act_arrive();
        const long long p3_ = clock64();
        cp_request();                                  // d(pre-activation) tile `dst` -> workspace
        const long long p4_ = clock64();
        ph_c += p2_ - p1_; ph_d += p3_ - p2_; ph_e += p4_ - p3_;
      }
    }
    // ---- reduce the pair's dW block into the flat gradient (split-K over the job's pairs) ----
    const long long t_dgrad_done = clock64();
    if ((fa.dbg & 4) && warp == 4 && lane == 0 && blockIdx.x < 148) {
      long long* o = g_fz_t[blockIdx.x];
      o[8] = ph_b; o[9] = ph_c; o[10] = ph_d; o[11] = ph_e;
    }
    if (wg_chunks > 0) {
      mbar_wait(bar_wacc, 0, 900);
      w_b = clock64() - t_dgrad_done;
      tc_fence_after();
      if (job.out_rows[rank] > 0) {
        // Each thread holds one dW ROW (TMEM lane) x 32 columns; reduced straight from there a warp-wide RED touches 32
        // different rows (32 separate L2 atomics, and scalar ones wherever a row of the flat gradient is not 16-byte
        // aligned: layer 5's [x, h] split, ld = 63 / 283 / 319, bottleneck_linear behind the 257-float sigma head --
        // those jobs' flush took ~100 k clk and set the kernel's duration).  So every 32 x 32 block is transposed
        // through shared memory (the dgrad weight ring, idle by now: every MMA that read it has completed) and reduced
        // row by row with lane = column: one coalesced 128-byte reduction per instruction, whatever the alignment.
        const int ncol_half = 64 * job.b_nfb;            // D columns handled by this warp: [h * ncol_half, (h + 1) * ncol_half)
        float* tile = reinterpret_cast<float*>(smem + L::kOffRing) + (warp - 4) * (32 * 33);
        float* obase = job.out + (int64_t)(job.out_row0[rank] + q * 32) * job.ld_out;
        for (int c0 = h * ncol_half; c0 < (h + 1) * ncol_half; c0 += 32) {
          if (c0 >= job.ncols) break;
          uint32_t r[32];
          tmem_ld32(taddr + 256 + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          if (c0 + lane < job.ncols) {
            float* ocol = obase + c0 + lane;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) atomicAdd(ocol + (int64_t)rr * job.ld_out, tile[rr * 33 + lane]);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
    }
  } else if (warp >= 14) {
    // ================= copy-out + publish: 2 warps, LSU path =================
    // Sequence per unit: d_hv (feature blocks 0-1), then the 9 d(pre-activation) tiles in chain order.  The tile in
    // shared memory is byte for byte its HBM image, so copy warp w moves whole 16 KB feature blocks (w, or 2w and 2w+1)
    // with coalesced LDS.128 / STG.128 -- the per-SM TMA unit is the busiest resource of this kernel (weight ring +
    // wgrad operand ring), so the stores stay off it -- then releases the tile's counter at gpu scope (a ~1 000 clk
    // MEMBAR that must not sit on the epilogue warps).
    const int w = warp - 14;
    uint32_t cp_ph = 0u;
    for (int k = 0;; ++k) {
      const int unit = unit_get(k);
      if (unit < 0) break;
      const int64_t tile = 2 * (int64_t)unit + rank;
      for (int step = 0; step <= kNumGemmsBwd; ++step) {
        const int g = step - 1;
        const int kind = (step == 0) ? 9 : (g == 0 ? 8 : 8 - g);
        mbar_wait(bar_cp, cp_ph, 950);
        cp_ph ^= 1u;
        uint8_t* dst = (step == 0) ? a.dhv_img + tile * 32768 : a.dpre_img + ((int64_t)kind * ntiles + tile) * 65536;
        const int nfb = (step == 0) ? 1 : 2, fb0 = w * nfb;
        const uint8_t* sp = smem + L::kOffAct + fb0 * 16384 + lane * 16;
        uint8_t* dp = dst + fb0 * 16384 + lane * 16;
        for (int i0 = 0; i0 < nfb * 32; i0 += 8) {
          uint4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4*>(sp + (i0 + i) * 512);
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(dp + (i0 + i) * 512) = v[i];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_cpfree);            // the tile has been read: the epilogue may overwrite it
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(fa.flags + (int64_t)kind * ntiles + tile, 1u);
      }
    }
  } else {
    // ================= db: column sums of this CTA's 128 dY features, straight from the operand ring (2 warps) =================
    const int b = (warp - 12) * 32 + lane;             // 0..63 -> features 2b, 2b+1 of the CTA's A operand
    const bool bias_on = job.bias != nullptr && job.out_rows[rank] > 0;
    float s0 = 0.f, s1 = 0.f;
    const int f = 2 * b, fb = f >> 6, ch = (f & 63) >> 3, e = f & 7;
    for (int c = 0; c < wg_chunks; ++c) {
      const int s = c % L::kStages;
      mbar_wait(bar_wfull(s), (c / L::kStages) & 1, 800 + s);
      if (bias_on) {
        const uint8_t* sa = smem + L::kOffWg + s * kFzStageBytes + kFzOffA + fb * 8192 + e * 2;
#pragma unroll 8
        for (int r = 0; r < 64; ++r) {
          const uint32_t w2 = *reinterpret_cast<const uint32_t*>(sa + r * 128 + ((ch ^ (r & 7)) << 4));
          s0 += __uint_as_float(w2 << 16);
          s1 += __uint_as_float(w2 & 0xFFFF0000u);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_wempty(s));
    }
    if (bias_on && wg_chunks > 0) {
      atomicAdd(job.bias + job.out_row0[rank] + f, s0);
      atomicAdd(job.bias + job.out_row0[rank] + f + 1, s1);
    }
  }
  if ((fa.dbg & 4) && lane == 0 && blockIdx.x < 148) {
    const long long t_end = clock64() - t_begin;
    long long* o = g_fz_t[blockIdx.x];
    if (warp == 0) { o[12] = w_thr; }                        // leader's producer: throttle waits
    if (warp == 1) { o[0] = w_a; o[1] = w_b; }               // dgrad issuer: weight-slot waits, activation waits
    if (warp == 2) { o[2] = w_a; o[3] = w_b; }               // wgrad loader: stage-free waits, flag waits
    if (warp == 3) { o[4] = w_a; }                           // wgrad issuer: stage-full waits
    if (warp == 4) { o[5] = w_a; o[6] = w_b; o[7] = t_end; } // epilogue: accumulator waits, wait for the dW block after the last dgrad unit, total
  }
  __syncthreads();
  cluster_sync();                                      // the peer may still be signalling this CTA's barriers
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

// One launch: dgrad chain + all tensor-core weight gradients.  (rgb / sigma heads: heads_wgrad_kernel, nerf_mlp_wgrad.cu)
static int launch_bwd_fused(const TcArgs& ta, void* ws, const WsLayout& L, float* grads, cudaStream_t st) {
  using FL = FzL;
  constexpr bool kWgOnly = false;
  static DeviceOnce attr_done;
  DeviceProps dp;
  int rc = current_device(&dp);
  if (rc) return rc;
  NERF_CHECK_ARG(dp.sm_major == 10, "libnerf_b200 needs an sm_100 device (found sm_%d%d); there is no fallback", dp.sm_major, dp.sm_minor);
  if (attr_done.needed(dp.ordinal)) {
    NERF_CUDA(cudaFuncSetAttribute(bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FL::kSmemBytes));
    attr_done.mark(dp.ordinal);
  }
  const int npairs = dp.sm_count / 2;
  NERF_CHECK_ARG(npairs >= kFzMaxJobs, "fused backward needs at least %d SM pairs (found %d)", kFzMaxJobs, npairs);
  const uint8_t* b = (const uint8_t*)ws;
  const int64_t ntiles = L.Mp / kTileM;
  auto ACT = [&](int l) { return b + L.act + (int64_t)l * ntiles * 65536; };     // h_l (l = 8: bottleneck)
  auto DPRE = [&](int l) { return b + L.dpre + (int64_t)l * ntiles * 65536; };   // d(pre-act) of layer l (8: d_bottleneck)
  FzArgs fa{};
  fa.t = ta;
  fa.flags = (uint32_t*)((uint8_t*)ws + L.flags);
  fa.ntiles = (int)ntiles;
  int nj = 0;
  double weight[kFzMaxJobs];
  // full = both CTAs hold 128 real output features; otherwise rank 1 duplicates rank 0's dY blocks and reduces nothing
  auto add = [&](const uint8_t* A, int a_tile_bytes, bool full, const uint8_t* B, int b_tile_bytes, int b_nfb, int layer,
                 int col0, int ncols, bool bias, int kind, double w) {
    FzJob& j = fa.jobs[nj];
    j.A = A; j.B = B; j.a_tile_bytes = a_tile_bytes; j.b_tile_bytes = b_tile_bytes;
    j.a_fb[0] = 0; j.a_fb[1] = full ? 2 : 0;
    j.b_fb[0] = 0; j.b_fb[1] = (b_nfb == 2) ? 2 : 0;
    j.b_nfb = b_nfb;
    j.out = grads + w_off(layer) + col0; j.ld_out = kIn[layer]; j.ncols = ncols;
    j.bias = bias ? grads + b_off(layer) : nullptr;
    j.out_row0[0] = 0; j.out_row0[1] = full ? 128 : 0;
    j.out_rows[0] = 128; j.out_rows[1] = full ? 128 : 0;
    j.kind = kind; j.need = kFzCopyWarps;
    j.single_reader = (kind != 5 && kind != 9) ? 1 : 0;
    weight[nj++] = w;
  };
  const uint8_t* xenc = b + L.xenc;
  const uint8_t* de16 = b + L.de16;
  const uint8_t* dhv = b + L.dhv;
  // job weights (SM pairs are apportioned in proportion): measured finish times x pairs of each job, 1024 rays x 192
  // (profiles/r02_fused_role_timings.txt).  A small job re-reads dY for a 64-wide X, so it costs most of a big one;
  // layer 0's dY is the last tensor of every unit's chain, so its roles start (and end) last.
  auto envw = [](const char* name, double dflt) { const char* e = getenv(name); return e ? atoi(e) / 100.0 : dflt; };
  static const double kSmall = envw("NERF_FZ_SMALL", 0.85), kW0 = envw("NERF_FZ_W0", 0.85), kW5 = envw("NERF_FZ_W5", 1.0);
  constexpr double kBig = 1.0;
  add(DPRE(0), 65536, true, xenc, 16384, 1, 0, 0, 63, true, 0, kW0);                         // layer 0: X = x_enc
  for (int l = 1; l <= 7; ++l) add(DPRE(l), 65536, true, ACT(l - 1), 65536, 2, l, l == 5 ? 63 : 0, 256, true, l, l == 5 ? kW5 : kBig);
  add(DPRE(5), 65536, true, xenc, 16384, 1, 5, 0, 63, false, 5, kSmall);                     // layer 5, x part of [x,h]
  add(DPRE(8), 65536, true, ACT(7), 65536, 2, L_BOTT, 0, 256, true, 8, kBig);                // bottleneck_linear
  add(dhv, 32768, false, ACT(8), 65536, 2, L_VIEW, 0, 256, true, 9, kBig);                   // view_linear: bottleneck columns + bias
  add(dhv, 32768, false, de16, 16384, 1, L_VIEW, 256, 27, false, 9, kSmall);                 // view_linear: direction columns
  fa.njobs = nj;
  // apportion the SM pairs to the jobs by weight (largest remainder), at least one each
  double wsum = 0;
  for (int j = 0; j < nj; ++j) wsum += weight[j];
  int alloc[kFzMaxJobs], used = 0;
  double frac[kFzMaxJobs];
  for (int j = 0; j < nj; ++j) {
    const double ideal = weight[j] * npairs / wsum;
    alloc[j] = (int)ideal < 1 ? 1 : (int)ideal;
    frac[j] = ideal - alloc[j];
    used += alloc[j];
  }
  while (used < npairs) {
    int best = 0;
    for (int j = 1; j < nj; ++j) if (frac[j] > frac[best]) best = j;
    ++alloc[best]; frac[best] -= 1.0; ++used;
  }
  int first = 0;
  for (int j = 0; j < nj; ++j) { fa.jobs[j].first_pair = first; fa.jobs[j].npairs = alloc[j]; first += alloc[j]; }
  static const int dbg = [] { const char* e = getenv("NERF_FZ_MODE"); return e ? atoi(e) : 0; }();
  static const int stagger = [] { const char* e = getenv("NERF_FZ_STAGGER"); return e ? atoi(e) : 0; }();   // measured: no gain from 1000 / 2000 clk
  fa.dbg = dbg;
  fa.grads = grads;
  for (int k = 0; k < 8; ++k) fa.db_off[k] = (int)b_off(k);
  fa.db_off[8] = (int)b_off(L_BOTT);
  fa.db_off[9] = (int)b_off(L_VIEW);
  fa.stagger_clk = stagger;
  fa.unit_ctr = fa.flags + (size_t)kFzKinds * ntiles;
  fa.progress = fa.unit_ctr + 32;
  // default: pairs + 20 % (88 on 148 SMs; sweep: profiles/r02_fused_throttle_sweep.txt) -- `npairs` units are in flight
  // by construction, so anything below that serialises the chains
  static const int window_env = [] { const char* e = getenv("NERF_FZ_WINDOW"); return e ? atoi(e) : -1; }();
  const int window = window_env >= 0 ? window_env : npairs + npairs / 5;
  static const int discard = [] { const char* e = getenv("NERF_FZ_DISCARD"); return e ? atoi(e) : 1; }();
  fa.window = window;
  fa.discard = discard;
  static const int prefetch = [] { const char* e = getenv("NERF_FZ_PREFETCH"); return e ? atoi(e) : kFzPrefetch; }();
  fa.prefetch = prefetch;

  if (!kWgOnly) {
    NERF_CUDA(cudaMemsetAsync(fa.flags, (dbg & 1) ? 0xFF : 0, (size_t)kFzKinds * ntiles * sizeof(uint32_t), st));
    NERF_CUDA(cudaMemsetAsync(fa.unit_ctr, 0, 160 * sizeof(uint32_t), st));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * npairs));
  cfg.blockDim = dim3(kFzThreads);
  cfg.dynamicSmemBytes = FL::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NERF_CUDA(cudaLaunchKernelEx(&cfg, bwd_fused_kernel, fa));
  NERF_LAUNCH_CHECK(kWgOnly ? "bwd_fused_kernel<wgrad only>" : "bwd_fused_kernel");
  if (dbg & 4) {                                          // development: role timings of a few CTAs (synchronises!)
    static int printed = 0;
    if (printed++ == 3) {
      long long h[148][13];
      NERF_CUDA(cudaStreamSynchronize(st));
      NERF_CUDA(cudaMemcpyFromSymbol(h, g_fz_t, sizeof(h)));
      long long tmin = 1ll << 60, tmax = 0; int bmax = 0;
      for (int bq = 0; bq < 2 * npairs && bq < 148; ++bq) { if (h[bq][7] < tmin) tmin = h[bq][7]; if (h[bq][7] > tmax) { tmax = h[bq][7]; bmax = bq; } }
      fprintf(stderr, "FZ CTA totals (epilogue warp 4): min %lld max %lld clk (CTA %d)\n", tmin, tmax, bmax);
      for (int j = 0; j < nj; ++j) {
        const int blk = 2 * fa.jobs[j].first_pair;
        fprintf(stderr, "FZ job %2d kind %d npairs %d (CTA %3d): total %7lld clk | dgrad issuer: wait weights %7lld, wait act %7lld | "
                "wgrad loader: wait stage %7lld, wait flag %7lld | wgrad issuer: wait full %7lld | epilogue: wait acc %7lld, tail wait dW %7lld | "
                "epilogue phases: copy-free wait %lld, tmem->smem %lld, fence+arrive %lld, copy request %lld | throttle %lld\n",
                j, fa.jobs[j].kind, fa.jobs[j].npairs, blk, h[blk][7], h[blk][0], h[blk][1], h[blk][2], h[blk][3], h[blk][4], h[blk][5], h[blk][6],
                h[blk][8], h[blk][9], h[blk][10], h[blk][11], h[blk][12]);
      }
    }
  }
  return 0;
}

}  // namespace nerf
