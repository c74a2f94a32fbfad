// Weight-gradient sweep of the fit step on the tensor-core path.
//
// The backward sweep (rnvp_mma.cu) runs with one thread per row, so the per-layer deltas it produces are row-major;
// the weight gradients are contractions OVER ROWS:
//     dW1[j][k] = sum_r delta1[r][j] * u[r][k]      db1[j] = sum_r delta1[r][j]       (first Linear of nn_t | nn_s)
//     dW2[i][j] = sum_r delta2[r][i] * h[r][j]      db2[i] = sum_r delta2[r][i]       (last Linear)
// i.e. tall-skinny GEMMs with K = rows.  The backward sweep leaves one record per (layer, row),
//     [ delta1 (2H: nn_t | nn_s) | h (2H) | u = [x_K, c, 0..] (K1P) | delta2 (2*TP: t | s) | pad ]   (REC floats, REC == 8 mod 32)
// and this kernel streams the [layer][row][REC] array (written once, read once) through a 2-stage ring of ONE TMA bulk
// copy per 32-row stage (the padded record stride makes every mma fragment load bank-conflict free) and contracts with
// warp-level mma.sync.m16n8k8 TF32 in the error-compensated 3-pass split (fp32-grade), main and correction products in
// separate register accumulators (the tensor core truncates on accumulation).  One CTA owns one (layer, row-slice):
// the gradients of its slice stay in registers for the whole slice and are flushed once with red.global.add, so the
// kernel is bound by reading 2.2 KB per layer-row from HBM.
#include <cuda_runtime.h>
#include <stdint.h>
#include "rnvp_wgrad.h"
#include "tc05.cuh"

namespace {
using namespace tc05;

constexpr int WG_THREADS = 512;             // 16 warps: one 16-unit slab of the 2H hidden units each
constexpr int WG_ROWS = 32;                 // rows per pipeline stage

// A-operand split (3 instructions): hi rounded to TF32, lo = exact fp32 remainder (the tensor core drops its low 13 bits:
// |error| < 2^-21 |v|, far below the fp32 noise of a 65k-row sum)
__device__ __forceinline__ void split_frag(float v, uint32_t& hi, uint32_t& lo) {
  hi = round_tf32(v);
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT1 = 8-wide column tiles of dW1 (ceil8(|K|+Cd)/8 <= 7), NT2 = column tiles of dW2 per net (|T|/8)
template <int NT1, int NT2>
__global__ void __launch_bounds__(WG_THREADS, 1) rnvp_wgrad_kernel(const __grid_constant__ RnvpWgradArgs a) {
  extern __shared__ __align__(128) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int H = a.H, H2 = 2 * H, K1P = NT1 * 8, TP = NT2 * 8;
  const int REC = a.rec;                                           // record stride == 8 (mod 32): conflict-free fragments
  const int stage_floats = WG_ROWS * REC;
  const int BW = K1P + 2 * TP;                                    // B-operand floats per record (u | delta2), contiguous
  float* lo_buf = sm + 2 * stage_floats;                           // [WG_ROWS][BW + 8] TF32 remainders of u | delta2
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * stage_floats + WG_ROWS * (NT1 * 8 + 2 * NT2 * 8 + 8));
  const int layer = blockIdx.x / a.n_slices, slice = blockIdx.x - layer * a.n_slices;
  const long long blocks_total = a.Npad / WG_ROWS;
  const long long per = (blocks_total + a.n_slices - 1) / a.n_slices;
  const long long blk0 = slice * per, blk1 = blk0 + per < blocks_total ? blk0 + per : blocks_total;
  const long long nblk = blk1 > blk0 ? blk1 - blk0 : 0;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const float* gR = a.gR + (size_t)layer * a.Npad * REC;
  // one bulk copy per stage: 32 consecutive records
  auto issue = [&](long long b, int st) {
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)stage_floats * 4u;
      mbar_expect_tx(&bars[st], bytes);
      bulk_g2s(sm + st * stage_floats, gR + (size_t)(blk0 + b) * WG_ROWS * REC, bytes, &bars[st]);
    }
  };
  if (warp == 0) {
    if (nblk > 0) issue(0, 0);
    if (nblk > 1) issue(1, 1);
  }

  // this warp's m-tiles: 16-row slabs of the 2H hidden units (nn_t units first, then nn_s)
  const int mtiles = H2 / 16;
  // c*m: main products of the current stage (short truncating chains), s*: their running fp32 sums (round-to-nearest
  // adds, one per stage); c*c: the tiny split-correction products (their truncation error is negligible)
  float c1m[1][NT1][4], c1c[1][NT1][4], c2m[1][NT2][4], c2c[1][NT2][4], cb[1][4];
  float s1[1][NT1][4], s2[1][NT2][4], sb[1][4];
#pragma unroll
  for (int m = 0; m < 1; ++m) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { cb[m][e] = 0.f; sb[m][e] = 0.f; }
#pragma unroll
    for (int j = 0; j < NT1; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { c1m[m][j][e] = 0.f; c1c[m][j][e] = 0.f; s1[m][j][e] = 0.f; }
#pragma unroll
    for (int j = 0; j < NT2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { c2m[m][j][e] = 0.f; c2c[m][j][e] = 0.f; s2[m][j][e] = 0.f; }
  }
  float db2 = 0.f;                                          // thread tid < 2*TP sums column tid of delta2
  const uint32_t one = (g == 0) ? 0x3f800000u : 0u;

  {
    const int mt0 = warp;                                    // 2H <= 256: one m-tile per warp covers all hidden units
    for (long long b = 0; b < nblk; ++b) {
      const int st = (int)(b & 1);
      mbar_wait(&bars[st], (uint32_t)((b >> 1) & 1));
      const float* D1 = sm + st * stage_floats;                // field offsets inside a record
      const float* Hh = D1 + H2;
      const float* U = Hh + H2;
      const float* E2 = U + K1P;
      const int sH = REC, sU = REC, sE = REC, sL = BW + 8;
      float db2_part = 0.f;
      {   // split the shared B operands once per stage (every warp used to redo this): hi in place, lo to lo_buf
        float* Uw = const_cast<float*>(U);
        for (int e = tid; e < WG_ROWS * BW; e += WG_THREADS) {
          const int r = e / BW, c = e - r * BW;
          const float v = Uw[r * REC + c];
          const uint32_t hi = round_tf32(v);
          Uw[r * REC + c] = __uint_as_float(hi);
          lo_buf[r * sL + c] = __uint_as_float(round_tf32(v - __uint_as_float(hi)));
        }
      }
      __syncthreads();
      if (tid < 2 * TP) {
#pragma unroll 8
        for (int r = 0; r < WG_ROWS; ++r) db2_part += E2[r * sE + tid] + lo_buf[r * sL + K1P + tid];   // hi + lo
        db2 += db2_part;
      }
#pragma unroll
      for (int r0 = 0; r0 < WG_ROWS; r0 += 8) {
        // B fragments: u (K = rows, N = weight columns) and delta2 (N = outputs i)
        uint32_t uh[NT1][2], ul[NT1][2];
#pragma unroll
        for (int j = 0; j < NT1; ++j) {
          uh[j][0] = __float_as_uint(U[(r0 + t) * sU + 8 * j + g]);
          uh[j][1] = __float_as_uint(U[(r0 + t + 4) * sU + 8 * j + g]);
          ul[j][0] = __float_as_uint(lo_buf[(r0 + t) * sL + 8 * j + g]);
          ul[j][1] = __float_as_uint(lo_buf[(r0 + t + 4) * sL + 8 * j + g]);
        }
#pragma unroll
        for (int m = 0; m < 1; ++m) {
          const int mt = mt0 + m;
          if (mt < mtiles) {
            const int j0 = 16 * mt + g;
            const int net = (16 * mt) / H;                  // m-tile lies entirely in one net (H % 16 == 0)
            // A fragments: delta1^T (dW1) and h^T (dW2^T); independent accumulators are interleaved so that no two
            // consecutive MMAs depend on each other
            uint32_t ah[4], al[4], hh[4], hl[4];
            split_frag(D1[(r0 + t) * sH + j0], ah[0], al[0]);
            split_frag(D1[(r0 + t) * sH + j0 + 8], ah[1], al[1]);
            split_frag(D1[(r0 + t + 4) * sH + j0], ah[2], al[2]);
            split_frag(D1[(r0 + t + 4) * sH + j0 + 8], ah[3], al[3]);
            split_frag(Hh[(r0 + t) * sH + j0], hh[0], hl[0]);
            split_frag(Hh[(r0 + t) * sH + j0 + 8], hh[1], hl[1]);
            split_frag(Hh[(r0 + t + 4) * sH + j0], hh[2], hl[2]);
            split_frag(Hh[(r0 + t + 4) * sH + j0 + 8], hh[3], hl[3]);
            uint32_t eh[NT2][2], el[NT2][2];
#pragma unroll
            for (int j = 0; j < NT2; ++j) {
              eh[j][0] = __float_as_uint(E2[(r0 + t) * sE + net * TP + 8 * j + g]);
              eh[j][1] = __float_as_uint(E2[(r0 + t + 4) * sE + net * TP + 8 * j + g]);
              el[j][0] = __float_as_uint(lo_buf[(r0 + t) * sL + K1P + net * TP + 8 * j + g]);
              el[j][1] = __float_as_uint(lo_buf[(r0 + t + 4) * sL + K1P + net * TP + 8 * j + g]);
            }
            mma_1688(cb[m], ah, one, one);
#pragma unroll
            for (int j = 0; j < NT1; ++j) mma_1688(c1c[m][j], al, uh[j][0], uh[j][1]);
#pragma unroll
            for (int j = 0; j < NT2; ++j) mma_1688(c2c[m][j], hl, eh[j][0], eh[j][1]);
#pragma unroll
            for (int j = 0; j < NT1; ++j) mma_1688(c1m[m][j], ah, uh[j][0], uh[j][1]);
#pragma unroll
            for (int j = 0; j < NT2; ++j) mma_1688(c2m[m][j], hh, eh[j][0], eh[j][1]);
            mma_1688(cb[m], al, one, one);
#pragma unroll
            for (int j = 0; j < NT1; ++j) mma_1688(c1c[m][j], ah, ul[j][0], ul[j][1]);
#pragma unroll
            for (int j = 0; j < NT2; ++j) mma_1688(c2c[m][j], hh, el[j][0], el[j][1]);
          }
        }
      }
      // fold this stage's main products into the running sums
#pragma unroll
      for (int m = 0; m < 1; ++m) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { sb[m][e] += cb[m][e]; cb[m][e] = 0.f; }
#pragma unroll
        for (int j = 0; j < NT1; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) { s1[m][j][e] += c1m[m][j][e]; c1m[m][j][e] = 0.f; }
#pragma unroll
        for (int j = 0; j < NT2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) { s2[m][j][e] += c2m[m][j][e]; c2m[m][j][e] = 0.f; }
      }
      __syncthreads();                                       // everyone is done with this stage
      if (warp == 0 && b + 2 < nblk) issue(b + 2, st);
    }
  }

  // ---- flush this slice's gradients (one red.add per value per CTA)
  const RnvpWgradLayer& lw = a.layers[layer];
  for (int m = 0; m < 1; ++m) {
    const int mt = warp + m;
    if (mt >= mtiles) continue;
    const int net = (16 * mt) / H;
    const int ja = 16 * mt + g - net * H, jb = ja + 8;       // hidden unit within the net
    float* gw1 = a.gpacked + lw.w1_off[net];
    float* gw2 = a.gpacked + lw.w2_off[net];
#pragma unroll
    for (int j = 0; j < NT1; ++j) {
      const int k = 8 * j + 2 * t;
      if (k < lw.Ks1) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gw1 + ja * lw.Ks1 + k), "f"(s1[m][j][0] + c1c[m][j][0]), "f"(s1[m][j][1] + c1c[m][j][1]) : "memory");
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gw1 + jb * lw.Ks1 + k), "f"(s1[m][j][2] + c1c[m][j][2]), "f"(s1[m][j][3] + c1c[m][j][3]) : "memory");
      }
    }
#pragma unroll
    for (int j = 0; j < NT2; ++j) {
      const int i0 = 8 * j + 2 * t;                          // dW2 is stored [i][unit]
      atomicAdd(gw2 + i0 * lw.Ks2 + ja, s2[m][j][0] + c2c[m][j][0]);
      atomicAdd(gw2 + (i0 + 1) * lw.Ks2 + ja, s2[m][j][1] + c2c[m][j][1]);
      atomicAdd(gw2 + i0 * lw.Ks2 + jb, s2[m][j][2] + c2c[m][j][2]);
      atomicAdd(gw2 + (i0 + 1) * lw.Ks2 + jb, s2[m][j][3] + c2c[m][j][3]);
    }
    if (t == 0) {
      atomicAdd(a.gpacked + lw.b1_off[net] + ja, sb[m][0]);
      atomicAdd(a.gpacked + lw.b1_off[net] + jb, sb[m][2]);
    }
  }
  if (tid < 2 * TP) {
    const int net = tid / TP, i = tid - net * TP;
    atomicAdd(a.gpacked + lw.b2_off[net] + i, db2);
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
  if (NT2 == 2) RNVP_WG_LAUNCH(2)
  if (NT2 == 4) RNVP_WG_LAUNCH(4)
#undef RNVP_WG_LAUNCH
  return cudaErrorInvalidValue;
}

}  // namespace

size_t rnvp_wgrad_smem_bytes(int rec, int bw) { return (size_t)(2 * WG_ROWS * rec + WG_ROWS * (bw + 8)) * 4 + 64; }

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
