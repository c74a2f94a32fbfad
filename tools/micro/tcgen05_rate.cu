// Micro-benchmark: issue cost of small tcgen05.mma kind::tf32 instructions (M=128, K=8, A in TMEM, B in smem,
// no-swizzle K-major) as a function of N, and of the accumulate-chain shape.  Answers: are the fused RealNVP kernels
// (hundreds of N=16 / N=64 MMAs per layer and row tile) bound by the per-instruction cost of the tensor core?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../probaforms_b200/csrc tcgen05_rate.cu -o tcgen05_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"
using namespace tc05;

// mode 0: all MMAs accumulate into ONE accumulator (dependent chain); mode 1: round-robin over 4 accumulators
// traffic: warps 4..11 hammer TMEM with tcgen05.ld/st x32 (as the epilogue warps of the product kernels do) meanwhile
__global__ void __launch_bounds__(384, 1) k(long long* out, int N, int n_mma, int mode, int a_tmem, int traffic) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  __shared__ volatile int done;
  if (tid == 0) done = 0;
  for (int i = tid; i < 256 * 8 + 128 * 8; i += 384) sm[i] = 1.0f;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_slot;
  if (warp == 0) {
    // warp-uniform control flow, one elected lane issues (as the product kernels do): descriptors stay in uniform registers
    const bool leader = elect_one();
    const uint32_t idesc = idesc_tf32(128, N);
    const uint64_t bdesc = smem_desc_kmajor_nosw(smem_u32(sm), 128u, 256u);
    const uint64_t adesc = smem_desc_kmajor_nosw(smem_u32(sm + 256 * 8), 128u, 256u);
    const long long t0 = clock64();
    const uint32_t dstep = mode ? (uint32_t)(N < 64 ? N : 64) : 0u;
    if (leader) {
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = tb + 256 + (uint32_t)(u & 3) * dstep;
          if (a_tmem) mma_tf32_ts(d, tb + 128 + 8 * u, bdesc, idesc, 1u);
          else mma_tf32_ss(d, adesc, bdesc, idesc, 1u);
        }
      }
    }
    __syncwarp();
    const long long t1 = clock64();
    if (leader) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (leader) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
      done = 1;
    }
  } else if (warp >= 4 && traffic) {
    const uint32_t trow = tb + ((uint32_t)((warp & 3) * 32) << 16) + (warp >= 8 ? 64u : 0u);   // columns 0..127: not the accumulators
    uint32_t r[32];
    while (!done) {
      tmem_ld_x32(trow, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = r[j] * 3u + 1u;
      tmem_st_x32(trow + 32, r);
      tmem_wait_st();
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  const int smem = (256 * 8 + 128 * 8) * 4;
  for (int traffic = 0; traffic < 2; ++traffic)
  for (int a_tmem = 1; a_tmem >= 0; --a_tmem)
    for (int mode = 0; mode < 2; ++mode)
      for (int N : {16, 32, 64, 128, 256}) {
        if (mode == 1 && N > 64) continue;
        if (traffic && (mode == 1 || a_tmem == 0)) continue;
        const int n = 2000;
        k<<<148, 384, smem>>>(d, N, 16, mode, a_tmem, traffic);
        k<<<148, 384, smem>>>(d, N, n, mode, a_tmem, traffic);
        long long h[2];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%sA %s, %s, N=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA (%.0f%% of 1934 MAC/clk)  %s\n",
               traffic ? "[TMEM ld/st traffic] " : "", a_tmem ? "TMEM" : "smem", mode ? "4 accumulators" : "1 accumulator ", N, (double)h[0] / n, (double)h[1] / n,
               100.0 * 128 * N * 8 / ((double)h[1] / n) / 1934.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
