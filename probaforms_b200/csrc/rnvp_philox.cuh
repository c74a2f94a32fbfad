// In-kernel prior draws for NormalizingFlow.sample (reference nflow.py:141: X = prior.sample((n,)), a standard normal
// per row and feature).  The reference draws them from torch's CPU generator and round-trips them through memory; here
// the inverse kernels generate them in registers, keyed on the GLOBAL row index so that the result does not depend on
// how the rows are sharded over GPUs or launches:
//
//   (x0, x1, x2, x3) = Philox4x32-10(counter = (row_lo, row_hi, j / 4, 0), key = (seed_lo, seed_hi))
//   u_k = ((x_k >> 9) + 0.5) * 2^-23                      (exact in fp32, in (0, 1))
//   eps[row][4*(j/4) + 0 | 1] = sqrt(-2 ln u0) * (cos | sin)(2 pi u1)      (Box-Muller)
//   eps[row][4*(j/4) + 2 | 3] = sqrt(-2 ln u2) * (cos | sin)(2 pi u3)
//
// oracle/realnvp_oracle.py:philox_normal restates exactly this in numpy (tests compare the two).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rnvp_rng {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ float unit_open(uint32_t x) { return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-7f; }

// the four standard normals of features 4*jblk .. 4*jblk+3 of global row `row`
__device__ __forceinline__ float4 normal4(unsigned long long seed, long long row, int jblk) {
  const uint4 x = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)((unsigned long long)row >> 32), (uint32_t)jblk, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float r0 = sqrtf(-2.0f * logf(unit_open(x.x))), r1 = sqrtf(-2.0f * logf(unit_open(x.z)));
  float s0, c0, s1, c1;
  sincospif(2.0f * unit_open(x.y), &s0, &c0);
  sincospif(2.0f * unit_open(x.w), &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__device__ __forceinline__ float normal1(unsigned long long seed, long long row, int j) {
  const float4 v = normal4(seed, row, j >> 2);
  const int q = j & 3;
  return q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w));
}

}  // namespace rnvp_rng
