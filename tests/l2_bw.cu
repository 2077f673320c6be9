// Development micro-benchmark (not a test): how fast can all SMs pull bytes with bulk-async copies (cp.async.bulk,
// the wgrad kernel's load path) from a buffer that is L2-resident vs one that streams from HBM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/l2_bw tests/l2_bw.cu && ./tests/l2_bw
// Prints GB/s for working sets of 16 MB .. 4 GB and ring depths of 2..6 x 32 KB per CTA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kStage = 32768;

// one elected thread per CTA: ring of `stages` x 32 KB; CTA b reads blocks b, b+grid, ... of the buffer, `reps` passes
__global__ void __launch_bounds__(128, 1) pull_kernel(const uint8_t* buf, size_t nblocks, int reps, int stages, int req = kStage) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + stages * kStage;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  long long issued = 0, waited = 0;
  for (int r = 0; r < reps; ++r) {
    for (size_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
      if (issued - waited == stages) {                 // ring full: wait for the oldest
        const int s = (int)(waited % stages);
        while (!mbar_try_wait(bar0 + 8 * s, (uint32_t)((waited / stages) & 1))) {}
        ++waited;
      }
      const int s = (int)(issued % stages);
      mbar_expect_tx(bar0 + 8 * s, kStage);
      for (int o = 0; o < kStage; o += req) bulk_g2s(sbase + s * kStage + o, buf + b * kStage + o, req, bar0 + 8 * s);
      ++issued;
    }
  }
  while (waited < issued) {
    const int s = (int)(waited % stages);
    while (!mbar_try_wait(bar0 + 8 * s, (uint32_t)((waited / stages) & 1))) {}
    ++waited;
  }
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const size_t cap = (size_t)4 << 30;
  uint8_t* buf;
  if (cudaMalloc(&buf, cap) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 1, cap);
  cudaFuncSetAttribute(pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kStage + 64);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%s, %d SMs, L2 %d MB\n", p.name, sms, p.l2CacheSize >> 20);
  const size_t sets[] = {(size_t)16 << 20, (size_t)32 << 20, (size_t)48 << 20, (size_t)64 << 20, (size_t)96 << 20, (size_t)256 << 20, (size_t)4 << 30};
  for (size_t ws : sets) {
    for (int stages : {2, 3, 4, 6}) {
      const size_t nblocks = ws / kStage;
      const int reps = (int)(((size_t)8 << 30) / ws);     // ~8 GB moved per measurement
      pull_kernel<<<sms, 128, stages * kStage + 64>>>(buf, nblocks, 2, stages);   // warm (fills L2 when it fits)
      cudaEventRecord(e0);
      pull_kernel<<<sms, 128, stages * kStage + 64>>>(buf, nblocks, reps, stages);
      cudaEventRecord(e1);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("working set %5zu MB  ring %d x 32 KB/SM: %7.0f GB/s  (%.1f B/clk/SM at 1.9 GHz)\n", ws >> 20, stages,
             (double)ws * reps / (ms * 1e-3) / 1e9, (double)ws * reps / (ms * 1e-3) / sms / 1.9e9);
    }
  }
  // request size: one stage = 32 KB moved as 32 KB / req bulk copies
  for (size_t ws : {(size_t)48 << 20, (size_t)4 << 30}) {
    for (int req : {32768, 16384, 8192, 4096, 2048}) {
      const int stages = 3;
      const size_t nblocks = ws / kStage;
      const int reps = (int)(((size_t)8 << 30) / ws);
      pull_kernel<<<sms, 128, stages * kStage + 64>>>(buf, nblocks, 2, stages, req);
      cudaEventRecord(e0);
      pull_kernel<<<sms, 128, stages * kStage + 64>>>(buf, nblocks, reps, stages, req);
      cudaEventRecord(e1);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("working set %5zu MB  ring 3 x 32 KB/SM, requests of %5d B: %7.0f GB/s  (%.1f B/clk/SM at 1.9 GHz)\n", ws >> 20, req,
             (double)ws * reps / (ms * 1e-3) / 1e9, (double)ws * reps / (ms * 1e-3) / sms / 1.9e9);
    }
  }
  return 0;
}
