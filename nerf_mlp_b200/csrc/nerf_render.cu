// Ray-side kernels of the NeRF hot path: stratified depths, volume compositing (fwd + analytic
// bwd), hierarchical resampling (+ sorted merge), Adam, MSE.  All warp-per-ray, HBM-bound:
// coalesced 16-byte loads of raw[R,S,4], fp32 math in the reference's operation order, and the
// two scans (transmittance cumprod, cdf cumsum) accumulated in fp64 and rounded once -- which is
// what torch's CPU cumprod/cumsum do and what oracle/nerf_oracle.py pins.
#include "nerf_common.cuh"
#include <math_constants.h>

namespace nerf {

constexpr int kWarpsPerBlock = 8;

// ------------------------------------------------------------------------------------------
// z_vals = near*(1-t) + far*t, optional stratified jitter            (reference renderer.py:52-61)
// ------------------------------------------------------------------------------------------
__global__ void stratified_z_kernel(const float* __restrict__ t_vals, const float* __restrict__ t_rand,
                                    int R, int S, float near_, float far_, float* __restrict__ z_vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * S) return;
  int s = (int)(i % S);
  auto zc = [&](int k) {
    float t = t_vals[k];
    return __fadd_rn(__fmul_rn(near_, __fsub_rn(1.f, t)), __fmul_rn(far_, t));  // :53, no FMA contraction
  };
  float z = zc(s);
  if (t_rand != nullptr) {
    float lower = (s == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, zc(s - 1)));       // :57-59
    float upper = (s == S - 1) ? z : __fmul_rn(0.5f, __fadd_rn(zc(s + 1), z));
    z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[i]));         // :61
  }
  z_vals[i] = z;
}

// PositionalEncoding.forward, materialised (drop-in surface only; the MLP kernels encode in-register)
__global__ void posenc_kernel(const float* __restrict__ x, int64_t n, int d, const float* __restrict__ freqs, int L,
                              int include_input, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  const int64_t row = i / d;
  const int c = (int)(i % d);
  const int width = d * (include_input + 2 * L);
  float* o = out + row * width;
  const float v = x[i];
  int off = 0;
  if (include_input) { o[c] = v; off = d; }
  for (int k = 0; k < L; ++k) {
    float s, co;
    sincosf(__fmul_rn(freqs[k], v), &s, &co);                     // model.py:24-25
    o[off + 2 * d * k + c] = s;
    o[off + 2 * d * k + d + c] = co;
  }
}

// ------------------------------------------------------------------------------------------
// Volume rendering integral                                   (reference renderer.py:114-163)
// One warp per ray; lane l owns samples l, l+32, ... (coalesced float4 loads of raw).
// ------------------------------------------------------------------------------------------
struct SampleTerms {
  float r, g, b;      // sigmoid(raw rgb)
  float alpha, om;    // alpha, (1-alpha)+1e-10
  float dist, e;      // dists*|d|, exp(-sigma' dist)
  float sig;          // raw sigma + noise
};

__device__ __forceinline__ SampleTerms sample_terms(const float4 rw, float nz, float z_cur, float z_nxt,
                                                   bool last, float dnorm) {
  SampleTerms t;
  t.r = 1.f / (1.f + expf(-rw.x));                                   // :130
  t.g = 1.f / (1.f + expf(-rw.y));
  t.b = 1.f / (1.f + expf(-rw.z));
  t.dist = __fmul_rn(last ? 1e10f : __fsub_rn(z_nxt, z_cur), dnorm);  // :120-127
  t.sig = rw.w + nz;                                                 // :134-136
  t.e = expf(-__fmul_rn(fmaxf(t.sig, 0.f), t.dist));                 // :140
  t.alpha = 1.f - t.e;
  t.om = __fadd_rn(__fsub_rn(1.f, t.alpha), 1e-10f);                 // :147 association
  return t;
}

// inclusive warp product scan in fp64
__device__ __forceinline__ double warp_scan_mul(double p, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, p, o);
    if (lane >= o) p *= t;
  }
  return p;
}
__device__ __forceinline__ double warp_scan_add(double p, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, p, o);
    if (lane >= o) p += t;
  }
  return p;
}

// One ray per warp: returns (all lanes) the composited rgb / depth / acc; writes weights if asked.
struct RayMaps { float r, g, b, depth, acc; };
__device__ __forceinline__ RayMaps composite_fwd_ray(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                                                     const float* __restrict__ rays_d, const float* __restrict__ noise,
                                                     int r, int S, int white_bkgd, float* __restrict__ weights, int lane) {
  const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const int64_t base = (int64_t)r * S;
  double carry = 1.0;
  float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
  float z_cur = lane < S ? z_vals[base + lane] : 0.f;
  for (int b0 = 0; b0 < S; b0 += 32) {
    const int s = b0 + lane;
    const bool valid = s < S;
    const float z_nb = (s + 32 < S) ? z_vals[base + s + 32] : 0.f;   // next block's depth (prefetch)
    const float4 rw = valid ? __ldg(raw + base + s) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float nz = (noise != nullptr && valid) ? noise[base + s] : 0.f;
    float z_nxt = __shfl_down_sync(0xffffffffu, z_cur, 1);
    const float z_n0 = __shfl_sync(0xffffffffu, z_nb, 0);
    if (lane == 31) z_nxt = z_n0;
    SampleTerms t = sample_terms(rw, nz, z_cur, z_nxt, s == S - 1, dnorm);
    double p = warp_scan_mul(valid ? (double)t.om : 1.0, lane);
    double incl = carry * p;
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = carry;
    carry = __shfl_sync(0xffffffffu, incl, 31);
    const float T = (float)excl;                  // fp32(cumprod in double), exclusive       :147
    const float w = valid ? __fmul_rn(t.alpha, T) : 0.f;                                   // :148
    if (weights != nullptr && valid) weights[base + s] = w;
    ar += __fmul_rn(w, t.r);                                                               // :151
    ag += __fmul_rn(w, t.g);
    ab += __fmul_rn(w, t.b);
    ad += __fmul_rn(w, z_cur);                                                             // :154
    aa += w;                                                                               // :157
    z_cur = z_nb;
  }
  ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); ad = warp_sum(ad); aa = warp_sum(aa);
  if (white_bkgd) {                                                                        // :160-161
    const float bg = 1.f - aa;
    ar += bg; ag += bg; ab += bg;
  }
  return RayMaps{ar, ag, ab, ad, aa};
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                     const float* __restrict__ rays_d, const float* __restrict__ noise, int R, int S,
                     int white_bkgd, float* __restrict__ rgb_map, float* __restrict__ depth_map,
                     float* __restrict__ acc_map, float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  const RayMaps m = composite_fwd_ray(raw, z_vals, rays_d, noise, r, S, white_bkgd, weights, lane);
  if (lane == 0) {
    rgb_map[3 * r] = m.r; rgb_map[3 * r + 1] = m.g; rgb_map[3 * r + 2] = m.b;
    depth_map[r] = m.depth;
    acc_map[r] = m.acc;
  }
}

// Analytic backward.  Two sweeps over the ray: (1) total = sum_k w_k g_k,
// (2) prefix sums so that suffix_{>i} = total - prefix_i; both in fp64 (the kernel is HBM-bound).
// gA is the upstream gradient of acc_map BEFORE the white-background term is folded in.
__device__ __forceinline__ void composite_bwd_ray(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                                                  const float* __restrict__ rays_d, const float* __restrict__ noise,
                                                  int r, int S, int white_bkgd, float gR, float gG, float gB, float gD,
                                                  float gA, const float* __restrict__ d_weights,
                                                  float4* __restrict__ d_raw, int lane) {
  const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const int64_t base = (int64_t)r * S;
  if (white_bkgd) gA -= (gR + gG + gB);   // rgb_map += 1 - acc
  double total = 0.0;
  for (int sweep = 0; sweep < 2; ++sweep) {
    double carry = 1.0, prefix = 0.0;
    float z_cur = lane < S ? z_vals[base + lane] : 0.f;
    for (int b0 = 0; b0 < S; b0 += 32) {
      const int s = b0 + lane;
      const bool valid = s < S;
      const float z_nb = (s + 32 < S) ? z_vals[base + s + 32] : 0.f;
      const float4 rw = valid ? __ldg(raw + base + s) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float nz = (noise != nullptr && valid) ? noise[base + s] : 0.f;
      float z_nxt = __shfl_down_sync(0xffffffffu, z_cur, 1);
      const float z_n0 = __shfl_sync(0xffffffffu, z_nb, 0);
      if (lane == 31) z_nxt = z_n0;
      SampleTerms t = sample_terms(rw, nz, z_cur, z_nxt, s == S - 1, dnorm);
      double p = warp_scan_mul(valid ? (double)t.om : 1.0, lane);
      double incl = carry * p;
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = carry;
      carry = __shfl_sync(0xffffffffu, incl, 31);
      const float T = (float)excl;
      const float w = valid ? t.alpha * T : 0.f;
      float g = gR * t.r + gG * t.g + gB * t.b + gD * z_cur + gA;
      if (d_weights != nullptr && valid) g += d_weights[base + s];
      const double wg = valid ? (double)w * (double)g : 0.0;
      if (sweep == 0) {
        total += wg;
      } else {
        double pin = warp_scan_add(wg, lane) + prefix;   // inclusive prefix of w*g
        prefix = __shfl_sync(0xffffffffu, pin, 31);
        const double suffix = total - pin;               // sum_{k>i} w_k g_k
        const float d_alpha = (float)((double)T * (double)g - suffix / (double)t.om);
        const float d_sig = (t.sig > 0.f) ? d_alpha * t.dist * t.e : 0.f;
        if (valid) {
          float4 o;
          o.x = w * gR * t.r * (1.f - t.r);
          o.y = w * gG * t.g * (1.f - t.g);
          o.z = w * gB * t.b * (1.f - t.b);
          o.w = d_sig;
          d_raw[base + s] = o;
        }
      }
      z_cur = z_nb;
    }
    if (sweep == 0) total = warp_sum(total);
  }
}


__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                     const float* __restrict__ rays_d, const float* __restrict__ noise, int R, int S,
                     int white_bkgd, const float* __restrict__ d_rgb, const float* __restrict__ d_depth,
                     const float* __restrict__ d_acc, const float* __restrict__ d_weights,
                     float4* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  composite_bwd_ray(raw, z_vals, rays_d, noise, r, S, white_bkgd, d_rgb[3 * r], d_rgb[3 * r + 1], d_rgb[3 * r + 2],
                    d_depth != nullptr ? d_depth[r] : 0.f, d_acc != nullptr ? d_acc[r] : 0.f, d_weights, d_raw, lane);
}

// ------------------------------------------------------------------------------------------
// Register-resident compositing for S <= 32 * NB (NB <= 8, i.e. every sampling of BASELINE.json except the 256+256
// stress): the ray's samples are loaded ONCE -- all NB coalesced float4 loads of a lane are issued before the first one
// is used, so a warp keeps NB * 512 B in flight instead of one block -- and the per-sample terms stay in registers,
// which lets the analytic backward run from them without reading `raw` again (the generic backward above sweeps the
// ray twice and the fused training kernel used to read it three times).  Arithmetic, operation order and the fp64
// scans are those of composite_fwd_ray / composite_bwd_ray: the forward is bit-identical to the generic path, the
// backward differs only in the association of d_sigma = d_alpha * (dist * e).
// ------------------------------------------------------------------------------------------
template <int NB>
struct RayRegs {
  float cr[NB], cg[NB], cb[NB];     // sigmoid(raw rgb)
  float alpha[NB], om[NB];          // alpha, (1 - alpha) + 1e-10
  float de[NB];                     // dist * exp(-sigma' dist), 0 where sigma' <= 0 (d alpha / d sigma)
  float T[NB], z[NB];               // exclusive transmittance, sample depth
};

template <int NB, bool kKeep>
__device__ __forceinline__ RayMaps composite_fwd_regs(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                                                      const float* __restrict__ rays_d, const float* __restrict__ noise,
                                                      int r, int S, int white_bkgd, float* __restrict__ weights, int lane,
                                                      RayRegs<NB>& st) {
  const int64_t base = (int64_t)r * S;
  float4 rw[NB];
  float zz[NB], nz[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) {                         // every load of the ray in flight before the first use
    const int s = 32 * k + lane;
    const bool valid = s < S;
    rw[k] = valid ? __ldg(raw + base + s) : make_float4(0.f, 0.f, 0.f, 0.f);
    zz[k] = valid ? z_vals[base + s] : 0.f;
    nz[k] = (noise != nullptr && valid) ? noise[base + s] : 0.f;
  }
  const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
  const float dnorm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  double carry = 1.0;
  float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int s = 32 * k + lane;
    const bool valid = s < S;
    float z_nxt = __shfl_down_sync(0xffffffffu, zz[k], 1);
    const float z_n0 = __shfl_sync(0xffffffffu, (k + 1 < NB) ? zz[k + 1] : 0.f, 0);
    if (lane == 31) z_nxt = z_n0;
    const SampleTerms t = sample_terms(rw[k], nz[k], zz[k], z_nxt, s == S - 1, dnorm);
    const double p = warp_scan_mul(valid ? (double)t.om : 1.0, lane);
    const double incl = carry * p;
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = carry;
    carry = __shfl_sync(0xffffffffu, incl, 31);
    const float T = (float)excl;                  // fp32(cumprod in double), exclusive       :147
    const float w = valid ? __fmul_rn(t.alpha, T) : 0.f;                                   // :148
    if (weights != nullptr && valid) weights[base + s] = w;
    ar += __fmul_rn(w, t.r);                                                               // :151
    ag += __fmul_rn(w, t.g);
    ab += __fmul_rn(w, t.b);
    ad += __fmul_rn(w, zz[k]);                                                             // :154
    aa += w;                                                                               // :157
    if (kKeep) {
      st.cr[k] = t.r; st.cg[k] = t.g; st.cb[k] = t.b;
      st.alpha[k] = valid ? t.alpha : 0.f; st.om[k] = t.om;
      st.de[k] = (t.sig > 0.f) ? t.dist * t.e : 0.f;
      st.T[k] = T; st.z[k] = zz[k];
    }
  }
  ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); ad = warp_sum(ad); aa = warp_sum(aa);
  if (white_bkgd) {                                                                        // :160-161
    const float bg = 1.f - aa;
    ar += bg; ag += bg; ab += bg;
  }
  return RayMaps{ar, ag, ab, ad, aa};
}

// analytic backward from the kept terms (same formulas and fp64 sums as composite_bwd_ray)
template <int NB>
__device__ __forceinline__ void composite_bwd_regs(const RayRegs<NB>& st, int r, int S, int white_bkgd, float gR, float gG,
                                                   float gB, float gD, float gA, const float* __restrict__ d_weights,
                                                   float4* __restrict__ d_raw, int lane) {
  const int64_t base = (int64_t)r * S;
  if (white_bkgd) gA -= (gR + gG + gB);   // rgb_map += 1 - acc
  float g[NB];
  double total = 0.0;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int s = 32 * k + lane;
    const bool valid = s < S;
    const float w = st.alpha[k] * st.T[k];
    g[k] = gR * st.cr[k] + gG * st.cg[k] + gB * st.cb[k] + gD * st.z[k] + gA;
    if (d_weights != nullptr && valid) g[k] += d_weights[base + s];
    total += valid ? (double)w * (double)g[k] : 0.0;
  }
  total = warp_sum(total);
  double prefix = 0.0;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int s = 32 * k + lane;
    const bool valid = s < S;
    const float w = st.alpha[k] * st.T[k];
    const double wg = valid ? (double)w * (double)g[k] : 0.0;
    const double pin = warp_scan_add(wg, lane) + prefix;   // inclusive prefix of w*g
    prefix = __shfl_sync(0xffffffffu, pin, 31);
    const double suffix = total - pin;                     // sum_{k>i} w_k g_k
    // suffix / om without a double-precision division (~30 instructions per sample): fp32 reciprocal + one fp64
    // Newton step (relative error ~1e-14, far below the final rounding to fp32)
    const float rf = __frcp_rn(st.om[k]);
    const double q0 = suffix * (double)rf;
    const double q = fma(q0, fma(-(double)st.om[k], (double)rf, 1.0), q0);
    const float d_alpha = (float)((double)st.T[k] * (double)g[k] - q);
    if (valid) {
      float4 o;
      o.x = w * gR * st.cr[k] * (1.f - st.cr[k]);
      o.y = w * gG * st.cg[k] * (1.f - st.cg[k]);
      o.z = w * gB * st.cb[k] * (1.f - st.cb[k]);
      o.w = d_alpha * st.de[k];
      d_raw[base + s] = o;
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)      // 64 registers at NB = 6: 4 blocks per SM; forcing 6 spills and is slower
composite_fwd_regs_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                          const float* __restrict__ rays_d, const float* __restrict__ noise, int R, int S,
                          int white_bkgd, float* __restrict__ rgb_map, float* __restrict__ depth_map,
                          float* __restrict__ acc_map, float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  RayRegs<NB> st;
  const RayMaps m = composite_fwd_regs<NB, false>(raw, z_vals, rays_d, noise, r, S, white_bkgd, weights, lane, st);
  if (lane == 0) {
    rgb_map[3 * r] = m.r; rgb_map[3 * r + 1] = m.g; rgb_map[3 * r + 2] = m.b;
    depth_map[r] = m.depth;
    acc_map[r] = m.acc;
  }
}

template <int NB>
// 4 resident blocks per SM (<= 64 registers, ~60 B of spill at NB = 6): the kernel is bound by latency, not by its
// register-resident working set -- 648 us at 2 blocks per SM (98 registers), 524 at 3, 490 at 4 for 262 144 x 192
__global__ void __launch_bounds__(kWarpsPerBlock * 32, NB <= 6 ? 4 : 2)
composite_bwd_regs_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals,
                          const float* __restrict__ rays_d, const float* __restrict__ noise, int R, int S,
                          int white_bkgd, const float* __restrict__ d_rgb, const float* __restrict__ d_depth,
                          const float* __restrict__ d_acc, const float* __restrict__ d_weights,
                          float4* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  RayRegs<NB> st;
  composite_fwd_regs<NB, true>(raw, z_vals, rays_d, noise, r, S, white_bkgd, nullptr, lane, st);
  composite_bwd_regs<NB>(st, r, S, white_bkgd, d_rgb[3 * r], d_rgb[3 * r + 1], d_rgb[3 * r + 2],
                         d_depth != nullptr ? d_depth[r] : 0.f, d_acc != nullptr ? d_acc[r] : 0.f, d_weights, d_raw, lane);
}

// Training pass of the FINE samples in one launch (scripts/train.py:374-382 around renderer.py:106-107):
// composite -> rgb_map; d_rgb = 2 (rgb - target) / (3R) (the gradient of mean((rgb-target)^2), :376);
// analytic backward -> d_raw; loss as a deterministic two-level fp64 reduction (per-block partials in
// `scratch`, summed in block order by the last block to finish); and, folded in because this is the
// last launch before the weight-gradient kernels, optimizer.zero_grad() of the flat gradient buffer.
template <int NB>      // NB > 0: register-resident ray (S <= 32 NB, one read of raw); NB = 0: generic three-sweep path
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_train_kernel(const float4* __restrict__ raw, const float* __restrict__ z_vals, const float* __restrict__ rays_d,
                       const float* __restrict__ noise, int R, int S, int white_bkgd, const float* __restrict__ target,
                       float* __restrict__ rgb_map, float* __restrict__ depth_map, float* __restrict__ acc_map,
                       float4* __restrict__ d_raw, float* __restrict__ loss, double* __restrict__ scratch,
                       float4* __restrict__ zero_buf, int64_t zero_n4) {
  __shared__ double part[kWarpsPerBlock];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_n4; i += (int64_t)gridDim.x * blockDim.x)
    zero_buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int r = blockIdx.x * kWarpsPerBlock + warp;
  double sq = 0.0;
  if (r < R) {
    RayRegs<(NB > 0 ? NB : 1)> st;
    const RayMaps m = (NB > 0) ? composite_fwd_regs<(NB > 0 ? NB : 1), true>(raw, z_vals, rays_d, noise, r, S, white_bkgd, nullptr, lane, st)
                               : composite_fwd_ray(raw, z_vals, rays_d, noise, r, S, white_bkgd, nullptr, lane);
    const float scale = 2.f / (float)(3 * (int64_t)R);
    const float e0 = m.r - target[3 * r], e1 = m.g - target[3 * r + 1], e2 = m.b - target[3 * r + 2];
    sq = (double)e0 * (double)e0 + (double)e1 * (double)e1 + (double)e2 * (double)e2;
    if (lane == 0) {
      rgb_map[3 * r] = m.r; rgb_map[3 * r + 1] = m.g; rgb_map[3 * r + 2] = m.b;
      depth_map[r] = m.depth;
      acc_map[r] = m.acc;
    }
    if (NB > 0)
      composite_bwd_regs<(NB > 0 ? NB : 1)>(st, r, S, white_bkgd, scale * e0, scale * e1, scale * e2, 0.f, 0.f, nullptr, d_raw, lane);
    else
      composite_bwd_ray(raw, z_vals, rays_d, noise, r, S, white_bkgd, scale * e0, scale * e1, scale * e2, 0.f, 0.f, nullptr,
                        d_raw, lane);
  }
  if (lane == 0) part[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) s += part[w];
    scratch[1 + blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned int*>(scratch), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // last block: the threads add the per-block partials (fixed assignment and tree => deterministic)
  __threadfence();
  double tot = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) tot += *(volatile double*)(scratch + 1 + b);
  tot = warp_sum(tot);
  __syncthreads();
  if (lane == 0) part[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    tot = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) tot += part[w];
    *reinterpret_cast<unsigned int*>(scratch) = 0u;
    *loss = (float)(tot / (double)(3 * (int64_t)R));
  }
}

// ------------------------------------------------------------------------------------------
// Hierarchical sampling (reference renderer.py:165-199) + sorted merge (:86-90).
// One warp per ray; cdf and bins staged in shared memory; binary search = searchsorted(right).
// ------------------------------------------------------------------------------------------
constexpr int kPdfWarps = 4;
constexpr int kMaxBins = NERF_MAX_SAMPLES;          // cdf length <= 512
constexpr int kMaxSort = 2 * NERF_MAX_SAMPLES;      // S_c + N_imp <= 1024

struct PdfArgs {
  const float* bins; int64_t bins_stride;           // explicit bins (generic entry) or
  const float* z_coarse;                            // bins = mids of z_coarse (fused entry)
  const float* weights; int64_t weights_stride; int weights_offset;
  const float* u; int u_shared;
  int R, NB, N_imp, S_c;
  float* samples; int64_t* inds; float* cdf_out; float* z_fine;
  // shared-memory layout per warp, in floats (launch_pdf): cdf[lay_a] | bins[lay_a] | samples[lay_p] | coarse depths[lay_z],
  // sized for THIS launch's NB / N_imp / S_c (the maxima would cap the kernel at 7 blocks per SM)
  int lay_a, lay_p, lay_z;
};

// Bitonic sort of 32 * EPL values held in registers (value i of the sequence = element i % EPL of lane i / EPL):
// compare-exchange partners closer than EPL are in the same lane, the others one __shfl_xor away -- no shared-memory
// round trips and no __syncwarp per stage (the shared-memory network below spent 69 % of its time in the LSU pipe).
// min / max of a pair keeps the multiset, so the result equals torch.sort's values bit for bit.
template <int EPL>
__device__ __forceinline__ void bitonic_regs(float (&e)[EPL], int lane) {
  constexpr int N = 32 * EPL;
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= EPL) {
        const int lm = stride / EPL;
        const bool lower = (lane & lm) == 0;                 // this element is the lower index of its pair
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
          const bool up = ((lane * EPL + j) & size) == 0;
          const float o = __shfl_xor_sync(0xffffffffu, e[j], lm);
          e[j] = (lower == up) ? fminf(e[j], o) : fmaxf(e[j], o);
        }
      } else {
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
          if ((j & stride) == 0) {
            const bool up = ((lane * EPL + j) & size) == 0;
            const float x = e[j], y = e[j | stride];
            const float lo = fminf(x, y), hi = fmaxf(x, y);
            e[j] = up ? lo : hi;
            e[j | stride] = up ? hi : lo;
          }
        }
      }
    }
  }
}
template <int EPL>
__device__ __forceinline__ void sort_in_regs(float* buf, int n, int lane) {      // buf[0 .. 32 * EPL): n values, rest ignored
  float e[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) { const int i = lane * EPL + j; e[j] = i < n ? buf[i] : CUDART_INF_F; }
  __syncwarp();
  bitonic_regs<EPL>(e, lane);
#pragma unroll
  for (int j = 0; j < EPL; ++j) buf[lane * EPL + j] = e[j];
  __syncwarp();
}

template <bool kMerge>
__global__ void __launch_bounds__(kPdfWarps * 32) sample_pdf_kernel(PdfArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kPdfWarps + warp;
  const int per_warp = 2 * a.lay_a + (kMerge ? a.lay_p + a.lay_z : 0);
  float* cdf = smem + (size_t)warp * per_warp;
  float* bins = cdf + a.lay_a;
  float* sortbuf = bins + a.lay_a;
  if (r >= a.R) return;
  const int NB = a.NB, NW = a.NB - 1;
  // bins
  if (a.z_coarse != nullptr) {
    const float* z = a.z_coarse + (int64_t)r * a.S_c;
    for (int k = lane; k < NB; k += 32) bins[k] = __fmul_rn(0.5f, __fadd_rn(z[k + 1], z[k]));   // :86
  } else {
    const float* b = a.bins + (int64_t)r * a.bins_stride;
    for (int k = lane; k < NB; k += 32) bins[k] = b[k];
  }
  // pdf normaliser: fp64 accumulate, round once (oracle.pdf_to_cdf)
  const float* wp = a.weights + (int64_t)r * a.weights_stride + a.weights_offset;
  double tot = 0.0;
  for (int k = lane; k < NW; k += 32) tot += (double)__fadd_rn(wp[k], 1e-5f);                    // :172
  const float total = (float)warp_sum(tot);
  // cdf = [0, cumsum(pdf)]: fp64 prefix, each prefix rounded to fp32                           // :173-175
  double carry = 0.0;
  for (int b0 = 0; b0 < NW; b0 += 32) {
    const int k = b0 + lane;
    const float pdf = k < NW ? __fdiv_rn(__fadd_rn(wp[k], 1e-5f), total) : 0.f;
    double v = warp_scan_add((double)pdf, lane) + carry;
    if (k < NW) cdf[k + 1] = (float)v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  if (a.cdf_out != nullptr)
    for (int k = lane; k < NB; k += 32) a.cdf_out[(int64_t)r * NB + k] = cdf[k];
  // inverse cdf
  for (int j = lane; j < a.N_imp; j += 32) {
    const float u = a.u_shared ? a.u[j] : a.u[(int64_t)r * a.N_imp + j];
    int lo = 0, hi = NB;                         // searchsorted(right=True): #{k : cdf[k] <= u}   :185
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = max(lo - 1, 0), above = min(lo, NB - 1);                                   // :186-187
    const float cb = cdf[below], ca = cdf[above], bb = bins[below], ba = bins[above];
    float denom = __fsub_rn(ca, cb);                                                             // :194
    if (denom < 1e-5f) denom = 1.f;                                                              // :195
    const float t = __fdiv_rn(__fsub_rn(u, cb), denom);                                          // :196
    const float s = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));                              // :197
    if (a.samples != nullptr) a.samples[(int64_t)r * a.N_imp + j] = s;
    if (a.inds != nullptr) a.inds[(int64_t)r * a.N_imp + j] = (int64_t)lo;
    if (kMerge) sortbuf[j] = s;
  }
  if (kMerge) {
    // z_fine = sort(cat[z_coarse, z_samples])                                                   // :90
    // A true multiset sort (bit-identical to torch.sort's values): bitonic-sort the N_imp samples, then
    // merge them with the coarse depths by rank (two binary searches per element) -- 2.5x fewer
    // compare-exchange steps than sorting all S_c + N_imp values.  The coarse depths are sorted by
    // construction (renderer.py:52-61); if a caller passes unsorted ones everything is sorted instead.
    const int n = a.S_c + a.N_imp;
    const float* z = a.z_coarse + (int64_t)r * a.S_c;
    float* zc = sortbuf + a.lay_p;                   // coarse depths  [S_c]
    float* merged = cdf;                             // cdf | bins are dead now: [2 * lay_a] >= pow2(n)
    bool sorted_in = true;
    for (int k = lane; k + 1 < a.S_c; k += 32)
      if (z[k + 1] < z[k]) sorted_in = false;
    sorted_in = __all_sync(0xffffffffu, sorted_in);
    auto bitonic = [&](float* buf, int npad) {
      for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int i = lane; i < (npad >> 1); i += 32) {
            const int lo_i = 2 * i - (i & (stride - 1));       // index with bit `stride` cleared
            const int hi_i = lo_i + stride;
            const bool up = (lo_i & size) == 0;
            const float x = buf[lo_i], y = buf[hi_i];
            if ((x > y) == up) { buf[lo_i] = y; buf[hi_i] = x; }
          }
          __syncwarp();
        }
      }
    };
    if (sorted_in) {
      int npad = 64;
      while (npad < a.N_imp) npad <<= 1;
      for (int k = lane; k < a.S_c; k += 32) zc[k] = z[k];
      __syncwarp();
      // the samples are already in order when u is (the deterministic linspace of render(), renderer.py:179) and no
      // rounding inverted a pair: nothing to sort then
      bool in_order = true;
      for (int k = lane; k + 1 < a.N_imp; k += 32)
        if (sortbuf[k + 1] < sortbuf[k]) in_order = false;
      in_order = __all_sync(0xffffffffu, in_order);
      if (!in_order) {
        if (npad == 64) sort_in_regs<2>(sortbuf, a.N_imp, lane);
        else if (npad == 128) sort_in_regs<4>(sortbuf, a.N_imp, lane);
        else if (npad == 256) sort_in_regs<8>(sortbuf, a.N_imp, lane);
        else {
          for (int k = a.N_imp + lane; k < npad; k += 32) sortbuf[k] = CUDART_INF_F;
          __syncwarp();
          bitonic(sortbuf, npad);
        }
      }
      for (int i = lane; i < a.S_c; i += 32) {       // rank of a coarse depth among the samples: #{j : zs[j] < a}
        const float v = zc[i];
        int lo = 0, hi = a.N_imp;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sortbuf[mid] < v) lo = mid + 1; else hi = mid; }
        merged[i + lo] = v;
      }
      for (int j = lane; j < a.N_imp; j += 32) {     // rank of a sample among the coarse depths: #{i : zc[i] <= b}
        const float v = sortbuf[j];
        int lo = 0, hi = a.S_c;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (zc[mid] <= v) lo = mid + 1; else hi = mid; }
        merged[j + lo] = v;
      }
    } else {
      int npad = 64;
      while (npad < n) npad <<= 1;
      __syncwarp();
      for (int k = lane; k < a.N_imp; k += 32) merged[a.S_c + k] = sortbuf[k];
      for (int k = lane; k < a.S_c; k += 32) merged[k] = z[k];
      for (int k = n + lane; k < npad; k += 32) merged[k] = CUDART_INF_F;
      __syncwarp();
      bitonic(merged, npad);
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) a.z_fine[(int64_t)r * n + k] = merged[k];
  }
}

// ------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam single-tensor order; reference scripts/train.py:258,387)
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float one_minus_b1, float b2,
                            float one_minus_b2, float neg_step_size, float bc2_sqrt, float eps,
                            float grad_scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  float mi = m[i], vi = v[i];
  mi = __fadd_rn(mi, __fmul_rn(__fsub_rn(gi, mi), one_minus_b1));             // lerp_(grad, 1-beta1)
  vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(__fmul_rn(gi, gi), one_minus_b2));  // mul_(b2).addcmul_
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
  p[i] = __fadd_rn(p[i], __fdiv_rn(__fmul_rn(neg_step_size, mi), denom));  // addcdiv_: p + value*m/denom
  m[i] = mi;
  v[i] = vi;
}

// mean((pred-target)^2) and its gradient                          (reference scripts/train.py:376)
__global__ void mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                           float* __restrict__ loss, float* __restrict__ d_pred) {
  __shared__ double part[32];
  double acc = 0.0;
  const float scale = 2.f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = pred[i] - target[i];
    acc += (double)d * (double)d;
    if (d_pred != nullptr) d_pred[i] = scale * d;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) *loss = (float)(v / (double)n);
  }
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_positional_encoding(const float* x, int64_t n, int d, const float* freqs, int L,
                                        int include_input, float* out, void* stream) {
  NERF_CHECK_ARG(n >= 0 && d >= 1 && L >= 0, "nerf_positional_encoding: bad shape n=%lld d=%d L=%d", (long long)n, d, L);
  if (n == 0) return 0;
  posenc_kernel<<<ceil_div(n * d, 256), 256, 0, (cudaStream_t)stream>>>(x, n, d, freqs, L, include_input, out);
  NERF_LAUNCH_CHECK("posenc_kernel");
  return 0;
}

extern "C" int nerf_stratified_z(const float* t_vals, const float* t_rand, int R, int S, float near_,
                                 float far_, float* z_vals, void* stream) {
  NERF_CHECK_ARG(R >= 0 && S >= 1, "nerf_stratified_z: bad shape R=%d S=%d", R, S);
  if (R == 0) return 0;
  const int64_t n = (int64_t)R * S;
  stratified_z_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(t_vals, t_rand, R, S, near_, far_, z_vals);
  NERF_LAUNCH_CHECK("stratified_z_kernel");
  return 0;
}

extern "C" int nerf_composite_fwd(const float* raw, const float* z_vals, const float* rays_d,
                                  const float* noise, int R, int S, int white_bkgd, float* rgb_map,
                                  float* depth_map, float* acc_map, float* weights, void* stream) {
  NERF_CHECK_ARG(R >= 0 && S >= 1, "nerf_composite_fwd: bad shape R=%d S=%d", R, S);
  NERF_CHECK_ARG(((uintptr_t)raw & 15) == 0, "nerf_composite_fwd: raw must be 16-byte aligned");
  if (R == 0) return 0;
  const dim3 grid(ceil_div(R, kWarpsPerBlock)), block(kWarpsPerBlock * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define NERF_FWD_NB(nb) composite_fwd_regs_kernel<nb><<<grid, block, 0, st>>>((const float4*)raw, z_vals, rays_d, noise, R, S, white_bkgd, rgb_map, depth_map, acc_map, weights)
  if (S <= 64) NERF_FWD_NB(2);
  else if (S <= 128) NERF_FWD_NB(4);
  else if (S <= 192) NERF_FWD_NB(6);
  else if (S <= 256) NERF_FWD_NB(8);
  else composite_fwd_kernel<<<grid, block, 0, st>>>((const float4*)raw, z_vals, rays_d, noise, R, S, white_bkgd, rgb_map, depth_map, acc_map, weights);
#undef NERF_FWD_NB
  NERF_LAUNCH_CHECK("composite_fwd_kernel");
  return 0;
}

extern "C" int nerf_composite_bwd(const float* raw, const float* z_vals, const float* rays_d,
                                  const float* noise, int R, int S, int white_bkgd, const float* d_rgb_map,
                                  const float* d_depth, const float* d_acc, const float* d_weights,
                                  float* d_raw, void* stream) {
  NERF_CHECK_ARG(R >= 0 && S >= 1, "nerf_composite_bwd: bad shape R=%d S=%d", R, S);
  NERF_CHECK_ARG((((uintptr_t)raw | (uintptr_t)d_raw) & 15) == 0, "nerf_composite_bwd: raw/d_raw must be 16-byte aligned");
  if (R == 0) return 0;
  const dim3 grid(ceil_div(R, kWarpsPerBlock)), block(kWarpsPerBlock * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define NERF_BWD_NB(nb) composite_bwd_regs_kernel<nb><<<grid, block, 0, st>>>((const float4*)raw, z_vals, rays_d, noise, R, S, white_bkgd, d_rgb_map, d_depth, d_acc, d_weights, (float4*)d_raw)
  if (S <= 64) NERF_BWD_NB(2);
  else if (S <= 128) NERF_BWD_NB(4);
  else if (S <= 192) NERF_BWD_NB(6);
  else if (S <= 256) NERF_BWD_NB(8);
  else composite_bwd_kernel<<<grid, block, 0, st>>>((const float4*)raw, z_vals, rays_d, noise, R, S, white_bkgd, d_rgb_map, d_depth, d_acc, d_weights, (float4*)d_raw);
#undef NERF_BWD_NB
  NERF_LAUNCH_CHECK("composite_bwd_kernel");
  return 0;
}

extern "C" size_t nerf_composite_train_scratch_bytes(int R) {
  return (size_t)(1 + ceil_div(R > 0 ? R : 1, kWarpsPerBlock)) * sizeof(double);
}

extern "C" int nerf_composite_train(const float* raw, const float* z_vals, const float* rays_d, const float* noise, int R,
                                    int S, int white_bkgd, const float* target, float* rgb_map, float* depth_map,
                                    float* acc_map, float* d_raw, float* loss, void* scratch, float* zero_buf,
                                    int64_t zero_n, void* stream) {
  NERF_CHECK_ARG(R >= 1 && S >= 1, "nerf_composite_train: bad shape R=%d S=%d", R, S);
  NERF_CHECK_ARG(raw && z_vals && rays_d && target && rgb_map && depth_map && acc_map && d_raw && loss && scratch,
                 "nerf_composite_train: null pointer");
  NERF_CHECK_ARG((((uintptr_t)raw | (uintptr_t)d_raw | (uintptr_t)zero_buf) & 15) == 0 && ((uintptr_t)scratch & 7) == 0,
                 "nerf_composite_train: raw/d_raw/zero_buf must be 16-byte aligned, scratch 8-byte aligned");
  NERF_CHECK_ARG(zero_n >= 0 && (zero_n & 3) == 0, "nerf_composite_train: zero_n must be a multiple of 4 (got %lld)", (long long)zero_n);
  const dim3 grid(ceil_div(R, kWarpsPerBlock)), block(kWarpsPerBlock * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define NERF_TRAIN_NB(nb) composite_train_kernel<nb><<<grid, block, 0, st>>>((const float4*)raw, z_vals, rays_d, noise, R, S, white_bkgd, target, rgb_map, depth_map, acc_map, (float4*)d_raw, loss, (double*)scratch, (float4*)zero_buf, zero_buf != nullptr ? zero_n / 4 : 0)
  if (S <= 64) NERF_TRAIN_NB(2);
  else if (S <= 128) NERF_TRAIN_NB(4);
  else if (S <= 192) NERF_TRAIN_NB(6);
  else if (S <= 256) NERF_TRAIN_NB(8);
  else NERF_TRAIN_NB(0);
#undef NERF_TRAIN_NB
  NERF_LAUNCH_CHECK("composite_train_kernel");
  return 0;
}

static int launch_pdf(const PdfArgs& a_in, bool merge, cudaStream_t st) {
  PdfArgs a = a_in;
  auto pad32 = [](int x) { return (x + 31) / 32 * 32; };
  auto pow2_64 = [](int x) { int p = 64; while (p < x) p <<= 1; return p; };
  // cdf | bins hold NB floats each and, once dead, the merged depths (padded to a power of two on the fall-back sort);
  // the sample region is sorted in place in blocks of 32 * EPL (or a power of two)
  a.lay_a = pad32(a.NB);
  if (merge && 2 * a.lay_a < pow2_64(a.S_c + a.N_imp)) a.lay_a = pow2_64(a.S_c + a.N_imp) / 2;
  a.lay_p = merge ? pow2_64(a.N_imp) : 0;
  a.lay_z = merge ? pad32(a.S_c) : 0;
  const size_t smem = (size_t)kPdfWarps * (2 * a.lay_a + a.lay_p + a.lay_z) * sizeof(float);
  const size_t smem_max = (size_t)kPdfWarps * (2 * kMaxBins + kMaxSort + kMaxBins) * sizeof(float);
  NERF_CHECK_ARG(smem <= smem_max, "sample_pdf: %d bins / %d samples exceed the shared-memory layout", a.NB, a.N_imp);
  static DeviceOnce attr_set;                       // per device (function attributes are per context)
  if (merge) {
    DeviceProps dp;
    int rc = current_device(&dp);
    if (rc) return rc;
    if (attr_set.needed(dp.ordinal)) {
      NERF_CUDA(cudaFuncSetAttribute(sample_pdf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
      attr_set.mark(dp.ordinal);
    }
    sample_pdf_kernel<true><<<ceil_div(a.R, kPdfWarps), kPdfWarps * 32, smem, st>>>(a);
  } else {
    sample_pdf_kernel<false><<<ceil_div(a.R, kPdfWarps), kPdfWarps * 32, smem, st>>>(a);
  }
  NERF_LAUNCH_CHECK("sample_pdf_kernel");
  return 0;
}

extern "C" int nerf_sample_pdf(const float* bins, int64_t bins_stride, const float* weights,
                               int64_t weights_stride, const float* u, int u_shared, int R, int NB, int N_imp,
                               float* samples, int64_t* inds, float* cdf, void* stream) {
  NERF_CHECK_ARG(R >= 0 && NB >= 2 && NB <= kMaxBins && N_imp >= 1, "nerf_sample_pdf: bad shape R=%d NB=%d N_imp=%d (NB<=%d)", R, NB, N_imp, kMaxBins);
  if (R == 0) return 0;
  PdfArgs a{};
  a.bins = bins; a.bins_stride = bins_stride; a.z_coarse = nullptr;
  a.weights = weights; a.weights_stride = weights_stride; a.weights_offset = 0;
  a.u = u; a.u_shared = u_shared; a.R = R; a.NB = NB; a.N_imp = N_imp; a.S_c = 0;
  a.samples = samples; a.inds = inds; a.cdf_out = cdf; a.z_fine = nullptr;
  return launch_pdf(a, false, (cudaStream_t)stream);
}

extern "C" int nerf_resample_merge(const float* z_coarse, const float* weights, const float* u, int u_shared,
                                   int R, int S_c, int N_imp, float* z_fine, float* z_samples, int64_t* inds,
                                   float* cdf, void* stream) {
  NERF_CHECK_ARG(R >= 0 && S_c >= 3 && S_c <= kMaxBins && N_imp >= 1 && S_c + N_imp <= kMaxSort,
                 "nerf_resample_merge: bad shape R=%d S_c=%d N_imp=%d (S_c<=%d, S_c+N_imp<=%d)", R, S_c, N_imp, kMaxBins, kMaxSort);
  if (R == 0) return 0;
  PdfArgs a{};
  a.bins = nullptr; a.bins_stride = 0; a.z_coarse = z_coarse;
  a.weights = weights; a.weights_stride = S_c; a.weights_offset = 1;     // weights[..., 1:-1]  :87
  a.u = u; a.u_shared = u_shared; a.R = R; a.NB = S_c - 1; a.N_imp = N_imp; a.S_c = S_c;
  a.samples = z_samples; a.inds = inds; a.cdf_out = cdf; a.z_fine = z_fine;
  return launch_pdf(a, true, (cudaStream_t)stream);
}

extern "C" int nerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                              double lr, double beta1, double beta2, double eps, int64_t step, float grad_scale,
                              void* stream) {
  NERF_CHECK_ARG(n >= 0 && step >= 1, "nerf_adam_step: bad n=%lld step=%lld", (long long)n, (long long)step);
  if (n == 0) return 0;
  // python-double scalar arithmetic as in torch/optim/adam.py, rounded to float at the tensor op
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  adam_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
      (float)(-(lr / bc1)), (float)sqrt(bc2), (float)eps, grad_scale);
  NERF_LAUNCH_CHECK("adam_kernel");
  return 0;
}

extern "C" int nerf_mse_loss(const float* pred, const float* target, int64_t n, float* loss, float* d_pred,
                             void* stream) {
  NERF_CHECK_ARG(n >= 1, "nerf_mse_loss: bad n=%lld", (long long)n);
  mse_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, target, n, loss, d_pred);
  NERF_LAUNCH_CHECK("mse_kernel");
  return 0;
}
