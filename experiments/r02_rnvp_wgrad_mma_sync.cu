// Weight-gradient sweep of the fit step on the tensor-core path.
//
// The backward sweep (rnvp_mma.cu) runs with one thread per row, so the per-layer deltas it produces are row-major;
// the weight gradients are contractions OVER ROWS:
//     dW1[j][k] = sum_r delta1[r][j] * u[r][k]      db1[j] = sum_r delta1[r][j]       (first Linear of nn_t | nn_s)
//     dW2[i][j] = sum_r delta2[r][i] * h[r][j]      db2[i] = sum_r delta2[r][i]       (last Linear)
// i.e. tall-skinny GEMMs with K = rows.  The backward sweep leaves one record per (layer, row),
//     [ h (2H: nn_t | nn_s) | u = [x_K, c, 0..] (K1P) | delta2 (2*TP: t | s) ]   (REC floats)
// (delta1 = (delta2 W2) * act'(h) is recomputed here: one more small MMA instead of 1 KB of HBM traffic per layer-row),
// stored in blocks of 32 rows as [layer][block][column group of 4][32 slots][4] (slot = row ^ (group & 1)): the backward
// sweep's per-row float4 stores coalesce to 512 B per warp, a block is ONE contiguous TMA bulk copy, and every mma
// fragment load from it is bank-conflict free.  This kernel streams the array (written once, read once) through a
// 3-stage ring and contracts with warp-level mma.sync.m16n8k8 TF32 in the error-compensated
// 3-pass split (fp32-grade), main and correction products in separate register accumulators (the tensor core truncates
// on accumulation); operands are split in registers, so the 16 consumer warps never synchronise with each other.  One
// CTA owns one (layer, row-slice): the gradients of its slice stay in registers for the whole slice and are flushed
// once with red.global.add, so the kernel is bound by reading 2.2 KB per layer-row from HBM.
#include <cuda_runtime.h>
#include <stdint.h>
#include "rnvp_wgrad.h"
#include "tc05.cuh"

namespace {
using namespace tc05;

constexpr int WG_WARPS = 16;                      // consumer warps: one 16-unit slab of the 2H hidden units each
constexpr int WG_THREADS = WG_WARPS * 32;         // lane 0 of warp 0 doubles as the TMA producer (register budget: 128)
constexpr int WG_ROWS = 32;                       // rows per block of the record array = rows per pipeline stage
constexpr int WG_STAGES = 3;

// A- and B-operand split for the 3-pass TF32 products.  mma.sync reads only the upper 19 bits of a tf32 operand, so the
// fp32 value itself serves as "hi" (the hardware uses trunc(v)) and lo = v - trunc(v) is exact in fp32 (2 instructions per
// element instead of 3 with a rounded hi; the 16 warps spend more issue slots on splitting than on MMAs).  What is lost is
// the part of lo below ITS upper 19 bits: < 2^-20 |v| per operand, far below the fp32 noise of a 65k-row sum.
__device__ __forceinline__ void split_frag(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v);
  lo = __float_as_uint(v - __uint_as_float(hi & 0xFFFFE000u));
}
__device__ __forceinline__ void mma_1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// float offset of (column c, row r) inside a 32-row block: [column group of 4][32 slots][4], slot = r ^ (group & 1)
__device__ __forceinline__ int blk_off(int c, int r) { return (c >> 2) * 128 + ((r ^ ((c >> 2) & 1)) << 2) + (c & 3); }

// NT1 = 8-wide column tiles of dW1 (ceil8(|K|+Cd)/8), NT2 = column tiles of dW2 per net (|T|/8)
template <int NT1, int NT2>
__global__ void __launch_bounds__(WG_THREADS, 1) rnvp_wgrad_kernel(const __grid_constant__ RnvpWgradArgs a) {
  extern __shared__ __align__(128) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int H = a.H, H2 = 2 * H, K1P = NT1 * 8, TP = NT2 * 8;
  const int REC = a.rec;                                           // floats per record = 2H + K1P + 2*TP
  const int stage_floats = WG_ROWS * REC;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + WG_STAGES * stage_floats);
  uint64_t* empty = full + WG_STAGES;
  const int layer = blockIdx.x / a.n_slices, slice = blockIdx.x - layer * a.n_slices;
  const long long blocks_total = a.Npad / WG_ROWS;
  const long long per = (blocks_total + a.n_slices - 1) / a.n_slices;
  const long long blk0 = slice * per, blk1 = blk0 + per < blocks_total ? blk0 + per : blocks_total;
  const long long nblk = blk1 > blk0 ? blk1 - blk0 : 0;

  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], WG_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  // ---- producer (lane 0 of warp 0): one bulk copy per stage = one 32-row block of this layer's records; issued as far
  // ahead as the ring allows without blocking, and blocking only for the block warp 0 itself needs next
  const float* gR = a.gR + ((size_t)layer * blocks_total + blk0) * stage_floats;
  long long pb = 0;
  auto produce = [&](long long b) {
    while (pb < nblk && pb < b + WG_STAGES) {
      const int st = (int)(pb % WG_STAGES);
      if (pb >= WG_STAGES) {
        const uint32_t par = (uint32_t)((pb / WG_STAGES - 1) & 1);
        if (pb <= b) mbar_wait(&empty[st], par);
        else if (!mbar_try_wait(&empty[st], par)) break;
      }
      const uint32_t bytes = (uint32_t)stage_floats * 4u;
      mbar_expect_tx(&full[st], bytes);
      bulk_g2s(sm + st * stage_floats, gR + (size_t)pb * stage_floats, bytes, &full[st]);
      ++pb;
    }
  };

  // ---- consumers.  c*m: main products of the current stage (short truncating chains), s*: their running fp32 sums
  // (round-to-nearest adds, one per stage); c*c: the tiny split-correction products (their truncation error is negligible)
  float c1m[NT1][4], c1c[NT1][4], c2m[NT2][4], c2c[NT2][4], s1[NT1][4], s2[NT2][4];
#pragma unroll
  for (int j = 0; j < NT1; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) { c1m[j][e] = 0.f; c1c[j][e] = 0.f; s1[j][e] = 0.f; }
#pragma unroll
  for (int j = 0; j < NT2; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) { c2m[j][e] = 0.f; c2c[j][e] = 0.f; s2[j][e] = 0.f; }
  float db1a = 0.f, db1b = 0.f;                              // bias gradients of units j0+g and j0+g+8 (rows t, t+4 of every 8)
  float db2[NT2];                                            // first warp of each net: column 8j+g of delta2
#pragma unroll
  for (int j = 0; j < NT2; ++j) db2[j] = 0.f;

  const int mtiles = H2 / 16;
  const bool active = warp < mtiles;                         // 2H <= 256: one m-tile per warp covers all hidden units
  const int j0 = 16 * warp;
  const int net = active ? j0 / H : 0;                       // m-tile lies entirely in one net (H % 16 == 0)
  const bool sums_b2 = active && (j0 == net * H);
  // Stage-relative float offsets of this thread's fragment elements (see blk_off): every column it touches is
  // (multiple of 8) + g, so the slot swizzle bit is par = (g >> 2) & 1 for all of them and rows r0+2t / r0+2t+1 sit in
  // slots r0 + (2t ^ par) / r0 + ((2t+1) ^ par); r0, the column tile j and the K step are compile-time immediates.
  const int par = (g >> 2) & 1;
  const int oa = (g >> 2) * 128 + ((2 * t) ^ par) * 4 + (g & 3), ob = (g >> 2) * 128 + ((2 * t + 1) ^ par) * 4 + (g & 3);
  const int iH = (j0 >> 2) * 128, iU = (H2 >> 2) * 128, iE = ((H2 + K1P + net * TP) >> 2) * 128;
  // B operand of the dh product: column (8kk + t) of delta2 (group parity 0) and column (8kk + 4 + t) (parity 1), row r0+g
  const int iB0 = iE + g * 4 + t, iB1 = iE + 128 + (g ^ 1) * 4 + t;
  // delta1 is not stored: dh = delta2 W2 is recomputed here (K = TP) and delta1 = dh * act'(h).  A operand of that product:
  // this warp's 16 columns of W2 (rows = transformed features), TF32 hi / lo, loaded once
  uint32_t wh[NT2][4], wl[NT2][4];
  {
    const RnvpWgradLayer& lw0 = a.layers[layer];
    const float* w2 = a.packed + lw0.w2_off[net];
    const int u0 = j0 + g - net * H;
#pragma unroll
    for (int kk = 0; kk < NT2; ++kk) {
      const float v0 = active ? w2[(8 * kk + t) * lw0.Ks2 + u0] : 0.f, v1 = active ? w2[(8 * kk + t) * lw0.Ks2 + u0 + 8] : 0.f;
      const float v2 = active ? w2[(8 * kk + t + 4) * lw0.Ks2 + u0] : 0.f, v3 = active ? w2[(8 * kk + t + 4) * lw0.Ks2 + u0 + 8] : 0.f;
      split_frag(v0, wh[kk][0], wl[kk][0]);
      split_frag(v1, wh[kk][1], wl[kk][1]);
      split_frag(v2, wh[kk][2], wl[kk][2]);
      split_frag(v3, wh[kk][3], wl[kk][3]);
    }
  }
  const bool is_tanh = a.act == 1;

  for (long long b = 0; b < nblk; ++b) {
    const int st = (int)(b % WG_STAGES);
    if (tid == 0) produce(b);
    mbar_wait(&full[st], (uint32_t)((b / WG_STAGES) & 1));
    const float* S = sm + st * stage_floats;
    if (active) {
#pragma unroll
      for (int r0 = 0; r0 < WG_ROWS; r0 += 8) {
        if (tid == 0 && r0) produce(b);
        // rows of this thread's K slots: the dh product below leaves delta1 in accumulator layout (unit g / g+8, rows
        // 2t / 2t+1 of the 8); used as the A fragment of the gradient products that makes K slot t <-> row 2t and K slot
        // t+4 <-> row 2t+1, and the B fragments (u, delta2) and h are read with the same row permutation
        float dm[4] = {0.f, 0.f, 0.f, 0.f}, dc[4] = {0.f, 0.f, 0.f, 0.f};
        {
          uint32_t bh[NT2][2], bl[NT2][2];
#pragma unroll
          for (int kk = 0; kk < NT2; ++kk) {
            split_frag(S[iB0 + 256 * kk + 4 * r0], bh[kk][0], bl[kk][0]);
            split_frag(S[iB1 + 256 * kk + 4 * r0], bh[kk][1], bl[kk][1]);
          }
#pragma unroll
          for (int kk = 0; kk < NT2; ++kk) mma_1688(dc, wl[kk], bh[kk][0], bh[kk][1]);
#pragma unroll
          for (int kk = 0; kk < NT2; ++kk) mma_1688(dm, wh[kk], bh[kk][0], bh[kk][1]);
#pragma unroll
          for (int kk = 0; kk < NT2; ++kk) mma_1688(dc, wh[kk], bl[kk][0], bl[kk][1]);
        }
        // B fragments: u (K = rows, N = weight columns) and delta2 (N = outputs i), split in registers
        uint32_t uh[NT1][2], ul[NT1][2], eh[NT2][2], el[NT2][2];
#pragma unroll
        for (int j = 0; j < NT1; ++j) {
          split_frag(S[iU + oa + 256 * j + 4 * r0], uh[j][0], ul[j][0]);
          split_frag(S[iU + ob + 256 * j + 4 * r0], uh[j][1], ul[j][1]);
        }
#pragma unroll
        for (int j = 0; j < NT2; ++j) {
          const float e0 = S[iE + oa + 256 * j + 4 * r0], e1 = S[iE + ob + 256 * j + 4 * r0];
          if (sums_b2) db2[j] += e0 + e1;
          split_frag(e0, eh[j][0], el[j][0]);
          split_frag(e1, eh[j][1], el[j][1]);
        }
        // A fragments: h^T (dW2^T) and delta1^T = (dh * act'(h))^T (dW1)
        uint32_t ah[4], al[4], hh[4], hl[4];
        const float h0 = S[iH + oa + 4 * r0], h1 = S[iH + oa + 256 + 4 * r0], h2 = S[iH + ob + 4 * r0], h3 = S[iH + ob + 256 + 4 * r0];
        float d0 = dm[0] + dc[0], d1 = dm[2] + dc[2], d2 = dm[1] + dc[1], d3 = dm[3] + dc[3];
        if (is_tanh) {
          d0 *= fmaf(-h0, h0, 1.0f); d1 *= fmaf(-h1, h1, 1.0f); d2 *= fmaf(-h2, h2, 1.0f); d3 *= fmaf(-h3, h3, 1.0f);
        } else {
          d0 = h0 > 0.f ? d0 : 0.f; d1 = h1 > 0.f ? d1 : 0.f; d2 = h2 > 0.f ? d2 : 0.f; d3 = h3 > 0.f ? d3 : 0.f;
        }
        db1a += d0 + d2;
        db1b += d1 + d3;
        split_frag(d0, ah[0], al[0]);
        split_frag(d1, ah[1], al[1]);
        split_frag(d2, ah[2], al[2]);
        split_frag(d3, ah[3], al[3]);
        split_frag(h0, hh[0], hl[0]);
        split_frag(h1, hh[1], hl[1]);
        split_frag(h2, hh[2], hl[2]);
        split_frag(h3, hh[3], hl[3]);
        // independent accumulators are interleaved so that no two consecutive MMAs depend on each other
#pragma unroll
        for (int j = 0; j < NT1; ++j) mma_1688(c1c[j], al, uh[j][0], uh[j][1]);
#pragma unroll
        for (int j = 0; j < NT2; ++j) mma_1688(c2c[j], hl, eh[j][0], eh[j][1]);
#pragma unroll
        for (int j = 0; j < NT1; ++j) mma_1688(c1m[j], ah, uh[j][0], uh[j][1]);
#pragma unroll
        for (int j = 0; j < NT2; ++j) mma_1688(c2m[j], hh, eh[j][0], eh[j][1]);
#pragma unroll
        for (int j = 0; j < NT1; ++j) mma_1688(c1c[j], ah, ul[j][0], ul[j][1]);
#pragma unroll
        for (int j = 0; j < NT2; ++j) mma_1688(c2c[j], hh, el[j][0], el[j][1]);
      }
      // fold this stage's main products into the running sums
#pragma unroll
      for (int j = 0; j < NT1; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { s1[j][e] += c1m[j][e]; c1m[j][e] = 0.f; }
#pragma unroll
      for (int j = 0; j < NT2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { s2[j][e] += c2m[j][e]; c2m[j][e] = 0.f; }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);                   // this warp is done reading the stage
  }
  if (!active) return;

  // ---- flush this slice's gradients (one red.add per value per CTA)
  const RnvpWgradLayer& lw = a.layers[layer];
  const int ja = j0 + g - net * H, jb = ja + 8;              // hidden units within the net
  float* gw1 = a.gpacked + lw.w1_off[net];
  float* gw2 = a.gpacked + lw.w2_off[net];
#pragma unroll
  for (int j = 0; j < NT1; ++j) {
    const int k = 8 * j + 2 * t;
    if (k < lw.Ks1) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gw1 + ja * lw.Ks1 + k), "f"(s1[j][0] + c1c[j][0]), "f"(s1[j][1] + c1c[j][1]) : "memory");
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gw1 + jb * lw.Ks1 + k), "f"(s1[j][2] + c1c[j][2]), "f"(s1[j][3] + c1c[j][3]) : "memory");
    }
  }
#pragma unroll
  for (int j = 0; j < NT2; ++j) {
    const int i0 = 8 * j + 2 * t;                            // dW2 is stored [i][unit]
    atomicAdd(gw2 + i0 * lw.Ks2 + ja, s2[j][0] + c2c[j][0]);
    atomicAdd(gw2 + (i0 + 1) * lw.Ks2 + ja, s2[j][1] + c2c[j][1]);
    atomicAdd(gw2 + i0 * lw.Ks2 + jb, s2[j][2] + c2c[j][2]);
    atomicAdd(gw2 + (i0 + 1) * lw.Ks2 + jb, s2[j][3] + c2c[j][3]);
  }
  // bias gradients: reduce over the 4 lanes (t) that share a unit / the 4 lanes... that hold other rows of the column
  db1a += __shfl_xor_sync(0xffffffffu, db1a, 1); db1a += __shfl_xor_sync(0xffffffffu, db1a, 2);
  db1b += __shfl_xor_sync(0xffffffffu, db1b, 1); db1b += __shfl_xor_sync(0xffffffffu, db1b, 2);
  if (t == 0) {
    atomicAdd(a.gpacked + lw.b1_off[net] + ja, db1a);
    atomicAdd(a.gpacked + lw.b1_off[net] + jb, db1b);
  }
  if (sums_b2) {
#pragma unroll
    for (int j = 0; j < NT2; ++j) {
      float v = db2[j];
      v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t == 0) atomicAdd(a.gpacked + lw.b2_off[net] + 8 * j + g, v);
    }
  }
}

template <int NT1>
cudaError_t launch_nt2(int NT2, const RnvpWgradArgs& a, int grid, size_t smem, cudaStream_t st) {
#define RNVP_WG_LAUNCH(N2)                                                                              \
  {                                                                                                      \
    auto k = rnvp_wgrad_kernel<NT1, N2>;                                                                 \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e != cudaSuccess) return e;                                                                      \
    k<<<grid, WG_THREADS, smem, st>>>(a);                                                                \
    return cudaGetLastError();                                                                           \
  }
  if (NT2 == 2) RNVP_WG_LAUNCH(2)       // D = 32 flows (the tcgen05 backward sweep that feeds this kernel supports only those)
#undef RNVP_WG_LAUNCH
  return cudaErrorInvalidValue;
}

}  // namespace

size_t rnvp_wgrad_smem_bytes(int rec, int bw) { (void)bw; return (size_t)(WG_STAGES * WG_ROWS * rec) * 4 + 16 * WG_STAGES; }

cudaError_t rnvp_launch_wgrad(int NT1, int NT2, const RnvpWgradArgs& a, int grid, size_t smem, cudaStream_t st) {
  switch (NT1) {
    case 2: return launch_nt2<2>(NT2, a, grid, smem, st);
    case 3: return launch_nt2<3>(NT2, a, grid, smem, st);
    case 4: return launch_nt2<4>(NT2, a, grid, smem, st);
    case 5: return launch_nt2<5>(NT2, a, grid, smem, st);
    case 6: return launch_nt2<6>(NT2, a, grid, smem, st);
    default: return cudaErrorInvalidValue;
  }
}
