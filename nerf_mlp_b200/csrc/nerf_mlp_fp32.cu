// fp32 check mode of the MLP (NERF_PREC_FP32): the reference's op sequence (model.py:57-81 and its
// autograd) as per-layer CUDA-core SGEMMs with fused bias/ReLU/mask epilogues.  This is the 1e-4
// parity gate of north_star and the on-device yardstick the tcgen05 kernels are debugged against;
// it is not the throughput path.
#include "nerf_common.cuh"

namespace nerf {

// ------------------------------------------------------------------------------------------
// Encoding kernels (fp32 mode materialises the encodings, exactly like the reference does)
// ------------------------------------------------------------------------------------------
constexpr int kXLd = 319;   // [x_enc(63) | h4(256)]  = input of layer 5        (model.py:62-63)
constexpr int kVLd = 283;   // [bottleneck(256) | d_enc(27)] = input of view    (model.py:72)

__device__ __forceinline__ void pe_write(float* dst, float x, float y, float z, int L) {
  dst[0] = x; dst[1] = y; dst[2] = z;
  float f = 1.f;
  for (int k = 0; k < L; ++k, f *= 2.f) {                      // model.py:23-25, freq = 2^k exactly
    float sx, cx, sy, cy, sz, cz;
    sincosf(f * x, &sx, &cx); sincosf(f * y, &sy, &cy); sincosf(f * z, &sz, &cz);
    float* o = dst + 3 + 6 * k;
    o[0] = sx; o[1] = sy; o[2] = sz; o[3] = cx; o[4] = cy; o[5] = cz;
  }
}

// rows [row0, row0+mb) of the flattened [R*S] sample list
__global__ void encode_rays_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                   const float* __restrict__ z_vals, int S, float coord_scale, int64_t row0,
                                   int64_t mb, float* __restrict__ X, float* __restrict__ V) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mb) return;
  const int64_t row = row0 + i;
  const int64_t r = row / S;
  const float ox = rays_o[3 * r], oy = rays_o[3 * r + 1], oz = rays_o[3 * r + 2];
  const float dx = rays_d[3 * r], dy = rays_d[3 * r + 1], dz = rays_d[3 * r + 2];
  const float z = z_vals[row];
  float px = __fadd_rn(ox, __fmul_rn(dx, z));                                   // renderer.py:63
  float py = __fadd_rn(oy, __fmul_rn(dy, z));
  float pz = __fadd_rn(oz, __fmul_rn(dz, z));
  if (coord_scale != 1.f) { px = __fmul_rn(px, coord_scale); py = __fmul_rn(py, coord_scale); pz = __fmul_rn(pz, coord_scale); }  // :67-68
  pe_write(X + i * kXLd, px, py, pz, 10);                                        // :70
  const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const float inv = __fadd_rn(n, 1e-8f);                                         // :72
  pe_write(V + i * kVLd + 256, __fdiv_rn(dx, inv), __fdiv_rn(dy, inv), __fdiv_rn(dz, inv), 4);  // :73-74
}

__global__ void copy_encoded_kernel(const float* __restrict__ x_enc, const float* __restrict__ d_enc, int64_t mb,
                                    float* __restrict__ X, float* __restrict__ V) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mb * 90) return;
  const int64_t row = i / 90;
  const int c = (int)(i % 90);
  if (c < 63) X[row * kXLd + c] = x_enc[row * 63 + c];
  else V[row * kVLd + 256 + (c - 63)] = d_enc[row * 27 + (c - 63)];
}

// ------------------------------------------------------------------------------------------
// Generic strided SGEMM: C(m,n) (+)= epi( sum_k A(m,k) B(n,k) )
// ------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A; int64_t a_rs, a_cs;      // A(m,k) = A[m*a_rs + k*a_cs]
  const float* B; int64_t b_rs, b_cs;      // B(n,k) = B[n*b_rs + k*b_cs]
  float* C; int64_t ldc;                   // C(m,n) = C[m*ldc + n]
  int64_t M; int N; int64_t K;
  const float* bias; int relu;             // + bias[n], ReLU
  const float* r1_row; int64_t r1_rs; const float* r1_col;   // + r1_row[m*r1_rs] * r1_col[n]
  const float* mask; int64_t ldmask;       // * (mask(m,n) > 0)
  int atomic; int64_t k_chunk;             // split-K over blockIdx.z, atomicAdd into C
};

constexpr int BM = 128, BN = 128, BK = 16;

__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
  const int64_t kend = min(g.K, kbeg + g.k_chunk);
  const int tx = t & 15, ty = t >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_kc = (g.a_cs == 1);   // K-contiguous A
  const bool b_kc = (g.b_cs == 1);
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- load A tile (BM x BK) ----
    if (a_kc) {
      const int row = t >> 1, kk = (t & 1) * 8;
      const int64_t m = m0 + row;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t k = k0 + kk + i;
        As[kk + i][row] = (m < g.M && k < kend) ? g.A[m * g.a_rs + k] : 0.f;
      }
    } else {
      const int kk = t >> 4, mm = (t & 15) * 8;
      const int64_t k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + mm + i;
        As[kk][mm + i] = (m < g.M && k < kend) ? g.A[m * g.a_rs + k * g.a_cs] : 0.f;
      }
    }
    // ---- load B tile (BN x BK) ----
    if (b_kc) {
      const int row = t >> 1, kk = (t & 1) * 8;
      const int n = n0 + row;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t k = k0 + kk + i;
        Bs[kk + i][row] = (n < g.N && k < kend) ? g.B[(int64_t)n * g.b_rs + k] : 0.f;
      }
    } else {
      const int kk = t >> 4, nn = (t & 15) * 8;
      const int64_t k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = n0 + nn + i;
        Bs[kk][nn + i] = (n < g.N && k < kend) ? g.B[(int64_t)n * g.b_rs + k * g.b_cs] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      *(float4*)&a[0] = *(const float4*)&As[kk][ty * 8];
      *(float4*)&a[4] = *(const float4*)&As[kk][ty * 8 + 4];
      *(float4*)&b[0] = *(const float4*)&Bs[kk][tx * 8];
      *(float4*)&b[4] = *(const float4*)&Bs[kk][tx * 8 + 4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
    const float r1 = g.r1_row != nullptr ? g.r1_row[m * g.r1_rs] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias != nullptr) v += g.bias[n];
      if (g.r1_row != nullptr) v = fmaf(r1, g.r1_col[n], v);
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask != nullptr) v = (g.mask[m * g.ldmask + n] > 0.f) ? v : 0.f;
      float* c = g.C + m * g.ldc + n;
      if (g.atomic) atomicAdd(c, v); else *c = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Forward-layer SGEMM (the hot GEMMs of the check mode: C = relu(A W^T + b), A and W both K-contiguous, N >= 64):
// 128 x 128 x 8 tiles, 8 x 8 outputs per thread, shared memory double-buffered with the next tile's global loads
// held in registers while the current tile is multiplied (one __syncthreads per K-step of 8), 16-byte global loads
// wherever the row pitch allows.  Plain FFMA in K order 0, 1, 2, ...: every output is the same fp32 dot product
// the generic kernel computes (same summation order), so the two kernels are bit-identical.
// (Round 1's generic kernel reached 18.5 TFLOP/s, slower than eager PyTorch on the same GPU.)
// ------------------------------------------------------------------------------------------
constexpr int FBK = 8;
// (Measured and dropped: the same loop on packed FFMA2 -- fma.rn.f32x2, two FMAs per instruction, A stored duplicated in
// shared memory.  Issue-slot use fell from 71 % to 46 % but the FMA pipe stayed at 58-60 % busy and the layer took
// 0.85 ms instead of 0.82: the limit of this kernel is the FMA pipe's operand delivery, not instruction issue.)
template <bool kVecA, bool kVecB>
__global__ void __launch_bounds__(256, 2) sgemm_fwd_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][FBK][BM + 4];
  __shared__ __align__(16) float Bs[2][FBK][BN + 4];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = t & 15, ty = t >> 4;
  const int lrow = t >> 1, lk = (t & 1) * 4;              // this thread loads 4 consecutive k of one row of each tile
  const int64_t am = m0 + lrow;
  const int bn = n0 + lrow;
  const float* ap = g.A + am * g.a_rs + lk;
  const float* bp = g.B + (int64_t)bn * g.b_rs + lk;
  const bool a_ok = am < g.M, b_ok = bn < g.N;
  const int K = (int)g.K;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  auto load_g = [&](int k0) {
    if (kVecA && a_ok && k0 + lk + 3 < K) {
      const float4 v = *reinterpret_cast<const float4*>(ap + k0);
      ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[i] = (a_ok && k0 + lk + i < K) ? ap[k0 + i] : 0.f;
    }
    if (kVecB && b_ok && k0 + lk + 3 < K) {
      const float4 v = *reinterpret_cast<const float4*>(bp + k0);
      rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) rb[i] = (b_ok && k0 + lk + i < K) ? bp[k0 + i] : 0.f;
    }
  };
  auto store_s = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[buf][lk + i][lrow] = ra[i]; Bs[buf][lk + i][lrow] = rb[i]; }
  };
  const int KT = (K + FBK - 1) / FBK;
  load_g(0);
  store_s(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) load_g((kt + 1) * FBK);               // in flight while this tile is multiplied
#pragma unroll
    for (int kk = 0; kk < FBK; ++kk) {
      // a thread's 8 x 8 outputs are rows {4 ty .. +3, 64 + 4 ty .. +3} x columns {4 tx .. +3, 64 + 4 tx .. +3}: the 16 tx
      // lanes of a half-warp then read 256 contiguous bytes of Bs per LDS.128 (conflict-free; with 8 consecutive
      // columns per thread the 32-byte stride made every B read a 2-way bank conflict and the shared-memory pipe as
      // busy as the FMA pipe: 0.97 -> 0.82 ms per 262 144 x 256 x 256 layer)
      float a[8], b[8];
      *(float4*)&a[0] = *(const float4*)&As[buf][kk][ty * 4];
      *(float4*)&a[4] = *(const float4*)&As[buf][kk][64 + ty * 4];
      *(float4*)&b[0] = *(const float4*)&Bs[buf][kk][tx * 4];
      *(float4*)&b[4] = *(const float4*)&Bs[buf][kk][64 + tx * 4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_s(buf ^ 1);
    __syncthreads();
  }
  // ---- epilogue: + bias, ReLU, 16-byte stores where the output pitch allows ----
  const bool vec_c = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && (n0 + BN <= g.N);
  auto col_of = [&](int j) { return n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)); };
  float bias[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int n = col_of(j); bias[j] = (g.bias != nullptr && n < g.N) ? g.bias[n] : 0.f; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = acc[i][j] + bias[j]; if (g.relu) v[j] = fmaxf(v[j], 0.f); }
    float* c = g.C + m * g.ldc;
    if (vec_c) {
      *reinterpret_cast<float4*>(c + col_of(0)) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(c + col_of(4)) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (col_of(j) < g.N) c[col_of(j)] = v[j];
    }
  }
}

// Heads (sigma 256 -> 1, rgb 128 -> 3): N <= 4 outputs per row -- one warp per row, lanes split K, warp-level sum.
// (Through the tiled SGEMM a 1-wide head cost as much as a 128-wide layer.)
__global__ void __launch_bounds__(256) head_gemv_kernel(GemmArgs g) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  const int K = (int)g.K, N = g.N;
  for (int64_t m = warp0; m < g.M; m += nwarps) {
    const float* a = g.A + m * g.a_rs;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < K; k += 32) {
      const float av = a[k];
#pragma unroll
      for (int n = 0; n < 4; ++n)
        if (n < N) acc[n] = fmaf(av, g.B[(int64_t)n * g.b_rs + k], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float v = warp_sum(acc[n]);
      if (lane == 0 && n < N) g.C[m * g.ldc + n] = v + (g.bias != nullptr ? g.bias[n] : 0.f);
    }
  }
}

static int run_gemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M == 0 || g.N == 0) return 0;
  const bool plain = g.a_cs == 1 && g.b_cs == 1 && !g.atomic && g.mask == nullptr && g.r1_row == nullptr && g.k_chunk >= g.K;
  if (plain && g.N <= 4 && !g.relu) {
    int64_t blocks = ceil_div(g.M, 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    head_gemv_kernel<<<(unsigned)blocks, 256, 0, st>>>(g);
    NERF_LAUNCH_CHECK("head_gemv_kernel");
    return 0;
  }
  if (plain && g.N >= 64) {
    const bool va = ((g.a_rs & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    const bool vb = ((g.b_rs & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);
    dim3 grid(ceil_div(g.M, BM), ceil_div(g.N, BN));
    // (the mixed variants -- one operand with 16-byte loads, the other scalar -- measured SLOWER than all-scalar:
    // 1.04 vs ~0.85 ms per 262 144 x 256 x 256 layer, so a misaligned operand sends both down the scalar path)
    if (va && vb) sgemm_fwd_kernel<true, true><<<grid, 256, 0, st>>>(g);
    else sgemm_fwd_kernel<false, false><<<grid, 256, 0, st>>>(g);
    NERF_LAUNCH_CHECK("sgemm_fwd_kernel");
    return 0;
  }
  dim3 grid(ceil_div(g.M, BM), ceil_div(g.N, BN), g.atomic ? ceil_div(g.K, g.k_chunk) : 1);
  sgemm_kernel<<<grid, 256, 0, st>>>(g);
  NERF_LAUNCH_CHECK("sgemm_kernel");
  return 0;
}

// db[n] += sum_m dY(m,n)
__global__ void colsum_kernel(const float* __restrict__ dY, int64_t ld, int64_t M, int N, int64_t rows_per_block,
                              float* __restrict__ db) {
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int64_t mb = (int64_t)blockIdx.x * rows_per_block;
  const int64_t me = min(M, mb + rows_per_block);
  float acc = 0.f;
  for (int64_t m = mb; m < me; ++m) acc += dY[m * ld + n];
  atomicAdd(db + n, acc);
}

static int run_colsum(const float* dY, int64_t ld, int64_t M, int N, float* db, cudaStream_t st) {
  if (M == 0) return 0;
  const int64_t rpb = 512;
  dim3 grid(ceil_div(M, rpb), ceil_div(N, 128));
  colsum_kernel<<<grid, 128, 0, st>>>(dY, ld, M, N, rpb, db);
  NERF_LAUNCH_CHECK("colsum_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------
// Workspace: per-row float counts
// ------------------------------------------------------------------------------------------
// X319 | H0 H1 H2 H3 | H5 H6 H7 | V283 | HV128           (forward; all kept when save=1)
// dHV128 | dA320 | dB320                                  (backward scratch, save=1 only)
constexpr int kFwdFloats = kXLd + 7 * 256 + kVLd + 128;          // 2522
constexpr int kBwdFloats = 128 + 320 + 320;
constexpr int64_t kInferBlockRows = 262144;

int mlp_fp32_workspace_floats_per_row(int save) { return save ? kFwdFloats + kBwdFloats : kFwdFloats; }

struct Fp32Bufs {
  float *X, *H[8], *V, *HV, *dHV, *dA, *dB;
  int64_t ldH[8];
};

static Fp32Bufs carve(float* ws, int64_t rows, bool bwd) {
  Fp32Bufs b{};
  float* p = ws;
  b.X = p; p += rows * kXLd;
  for (int i = 0; i < 8; ++i) {
    if (i == 4) { b.H[4] = b.X + 63; b.ldH[4] = kXLd; continue; }   // layer-4 output lives inside X (columns 63..318)
    b.H[i] = p; b.ldH[i] = 256; p += rows * 256;
  }
  b.V = p; p += rows * kVLd;
  b.HV = p; p += rows * 128;
  if (bwd) { b.dHV = p; p += rows * 128; b.dA = p; p += rows * 320; b.dB = p; p += rows * 320; }
  return b;
}

static GemmArgs fwd_args(const float* A, int64_t lda, const float* params, int layer, int k_in, float* C,
                         int64_t ldc, int64_t M, int relu) {
  GemmArgs g{};
  g.A = A; g.a_rs = lda; g.a_cs = 1;
  g.B = params + w_off(layer); g.b_rs = kIn[layer]; g.b_cs = 1;
  g.C = C; g.ldc = ldc; g.M = M; g.N = kOut[layer]; g.K = k_in;
  g.bias = params + b_off(layer); g.relu = relu; g.k_chunk = k_in;
  return g;
}

int mlp_fp32_forward(const float* rays_o, const float* rays_d, const float* z_vals, int R, int S,
                     float coord_scale, const float* x_enc, const float* d_enc, int64_t M,
                     const float* params, float* out, float* ws, size_t ws_bytes, int save, cudaStream_t st) {
  const int64_t block_rows = save ? M : (M < kInferBlockRows ? M : kInferBlockRows);
  const size_t need = (size_t)block_rows * mlp_fp32_workspace_floats_per_row(save) * sizeof(float);
  NERF_CHECK_ARG(ws_bytes >= need, "mlp fp32 forward: workspace too small (%zu < %zu)", ws_bytes, need);
  for (int64_t row0 = 0; row0 < M; row0 += block_rows) {
    const int64_t mb = (M - row0 < block_rows) ? (M - row0) : block_rows;
    Fp32Bufs b = carve(ws, block_rows, save != 0);
    if (x_enc != nullptr) {
      copy_encoded_kernel<<<ceil_div(mb * 90, 256), 256, 0, st>>>(x_enc + row0 * 63, d_enc + row0 * 27, mb, b.X, b.V);
      NERF_LAUNCH_CHECK("copy_encoded_kernel");
    } else {
      encode_rays_kernel<<<ceil_div(mb, 128), 128, 0, st>>>(rays_o, rays_d, z_vals, S, coord_scale, row0, mb, b.X, b.V);
      NERF_LAUNCH_CHECK("encode_rays_kernel");
    }
    float* o = out + row0 * 4;
    int rc;
    if ((rc = run_gemm(fwd_args(b.X, kXLd, params, 0, 63, b.H[0], 256, mb, 1), st))) return rc;
    for (int l = 1; l < 8; ++l) {
      const float* A = (l == 5) ? b.X : b.H[l - 1];
      const int64_t lda = (l == 5) ? kXLd : b.ldH[l - 1];
      if ((rc = run_gemm(fwd_args(A, lda, params, l, kIn[l], b.H[l], b.ldH[l], mb, 1), st))) return rc;
    }
    if ((rc = run_gemm(fwd_args(b.H[7], 256, params, L_SIGMA, 256, o + 3, 4, mb, 0), st))) return rc;   // model.py:69
    if ((rc = run_gemm(fwd_args(b.H[7], 256, params, L_BOTT, 256, b.V, kVLd, mb, 0), st))) return rc;    // :70
    if ((rc = run_gemm(fwd_args(b.V, kVLd, params, L_VIEW, 283, b.HV, 128, mb, 1), st))) return rc;      // :72-74
    if ((rc = run_gemm(fwd_args(b.HV, 128, params, L_RGB, 128, o, 4, mb, 0), st))) return rc;            // :75,77
  }
  return 0;
}

// dW[layer] += dY^T . X ; db[layer] += colsum(dY)
static int wgrad(const float* dY, int64_t ld_dy, const float* Xin, int64_t ld_x, int layer, int64_t M,
                 float* grads, cudaStream_t st) {
  GemmArgs g{};
  g.A = dY; g.a_rs = 1; g.a_cs = ld_dy;          // A(m'=out, k=row)
  g.B = Xin; g.b_rs = 1; g.b_cs = ld_x;          // B(n'=in,  k=row)
  g.C = grads + w_off(layer); g.ldc = kIn[layer];
  g.M = kOut[layer]; g.N = kIn[layer]; g.K = M;
  g.atomic = 1; g.k_chunk = 4096;
  int rc = run_gemm(g, st);
  if (rc) return rc;
  return run_colsum(dY, ld_dy, M, kOut[layer], grads + b_off(layer), st);
}

// dX(m, 0..n_in) = (dY . W[layer][:, col0:col0+n_in]) [+ rank-1] [* (mask > 0)]
static int dgrad(const float* dY, int64_t ld_dy, const float* params, int layer, int col0, int n_in, float* dX,
                 int64_t ld_dx, const float* mask, int64_t ldmask, const float* r1_row, int64_t r1_rs,
                 const float* r1_col, int64_t M, cudaStream_t st) {
  GemmArgs g{};
  g.A = dY; g.a_rs = ld_dy; g.a_cs = 1;
  g.B = params + w_off(layer) + col0; g.b_rs = 1; g.b_cs = kIn[layer];   // B(n=in, k=out) = W[out][col0+in]
  g.C = dX; g.ldc = ld_dx; g.M = M; g.N = n_in; g.K = kOut[layer]; g.k_chunk = g.K;
  g.mask = mask; g.ldmask = ldmask; g.r1_row = r1_row; g.r1_rs = r1_rs; g.r1_col = r1_col;
  return run_gemm(g, st);
}

int mlp_fp32_backward(const float* d_raw, int64_t M, const float* params, float* grads, float* ws,
                      size_t ws_bytes, cudaStream_t st) {
  const size_t need = (size_t)M * mlp_fp32_workspace_floats_per_row(1) * sizeof(float);
  NERF_CHECK_ARG(ws_bytes >= need, "mlp fp32 backward: workspace too small (%zu < %zu)", ws_bytes, need);
  Fp32Bufs b = carve(ws, M, true);
  int rc;
  // rgb_linear (model.py:75)
  if ((rc = wgrad(d_raw, 4, b.HV, 128, L_RGB, M, grads, st))) return rc;
  if ((rc = dgrad(d_raw, 4, params, L_RGB, 0, 128, b.dHV, 128, b.HV, 128, nullptr, 0, nullptr, M, st))) return rc;
  // view_linear (:73); only the bottleneck part of its input needs a gradient
  if ((rc = wgrad(b.dHV, 128, b.V, kVLd, L_VIEW, M, grads, st))) return rc;
  if ((rc = dgrad(b.dHV, 128, params, L_VIEW, 0, 256, b.dA, 256, nullptr, 0, nullptr, 0, nullptr, M, st))) return rc;
  // bottleneck (:70) and sigma (:69) both read h7
  if ((rc = wgrad(b.dA, 256, b.H[7], 256, L_BOTT, M, grads, st))) return rc;
  if ((rc = wgrad(d_raw + 3, 4, b.H[7], 256, L_SIGMA, M, grads, st))) return rc;
  if ((rc = dgrad(b.dA, 256, params, L_BOTT, 0, 256, b.dB, 256, b.H[7], 256, d_raw + 3, 4,
                  params + w_off(L_SIGMA), M, st))) return rc;
  float* cur = b.dB;   // d(pre-activation) of layer l
  float* nxt = b.dA;
  for (int l = 7; l >= 0; --l) {
    const float* Xin = (l == 0 || l == 5) ? b.X : b.H[l - 1];
    const int64_t ldx = (l == 0 || l == 5) ? kXLd : b.ldH[l - 1];
    if ((rc = wgrad(cur, 256, Xin, ldx, l, M, grads, st))) return rc;
    if (l == 0) break;
    const int col0 = (l == 5) ? 63 : 0;          // skip: only the h part of [x,h] carries gradient
    if ((rc = dgrad(cur, 256, params, l, col0, 256, nxt, 256, b.H[l - 1], b.ldH[l - 1], nullptr, 0, nullptr, M, st))) return rc;
    float* t = cur; cur = nxt; nxt = t;
  }
  return 0;
}

}  // namespace nerf
