// tcgen05 (5th-gen tensor core) path of the RealNVP hot path: fused forward (log-density) and inverse
// (sampling) passes for flows whose conditioner GEMMs are real contractions (hidden width >= 32,
// D/2 in {16, 32}: BASELINE configs c3, c4).
//
// rnvp_mma_kernel -- one persistent CTA per SM, 11 warps, processes PAIRS of 128-row tiles:
//   warp 0      TMA producer: per coupling layer one bulk copy of the W1 image and one of the W2 image
//               (TF32 hi/lo splits, pre-tiled in the no-swizzle K-major core-matrix layout) into smem
//   warps 1,10  MMA issuers of tile 0 / tile 1 (one elected thread each): tcgen05.mma kind::tf32, A operands in TMEM,
//               accumulators in TMEM, error-compensated split (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo) = fp32-grade accuracy
//   warps 2-5   epilogue group of tile 0, warps 6-9 of tile 1: ONE THREAD PER ROW.  The row (x, c, log-det)
//               lives in registers for the whole flow; per layer the thread writes u=[x_K,c,1] to TMEM,
//               turns each GEMM1 accumulator chunk into tanh(.) hi/lo in place (tcgen05.ld/st), and applies
//               the coupling y_T = x_T*exp(s)+t from the GEMM2 accumulators.
//   While one tile's threads run the tanh epilogue (MUFU-bound), the tensor core runs the other tile's MMAs.
//   Hidden units are processed in chunks of CU per net so that a tile needs <= 256 TMEM columns
//   (u_hi, u_lo, D1/A_hi, A_lo, [D2 | C2] per net).  b1 rides in GEMM1 as an extra K column of ones.
//
//   MODE 2 (fit step, D = 32 flows): the forward sweep additionally stashes (x_T, s) per layer and writes every
//   h = act(.) into the activation record of its (layer, row); the same CTA then sweeps the layers BACKWARDS with the
//   same row ownership: delta2 from the stash -> TMEM, per 16-unit half-chunk dh = delta2 W2 on the tensor core,
//   delta1 = dh * act'(h) with h read back from the record, du += delta1 W1[:, x_K] (transposed K-major weight
//   images W2T / W1T).  u and delta2 complete the record; rnvp_wgrad_kernel contracts the records over rows.
//
// Also here: the primitive self-test: D[128 x N] = A[128 x K] * B[N x K]^T with A
// staged in TMEM (tcgen05.st, one thread per row), B in shared memory in the no-swizzle K-major
// core-matrix layout, kind::tf32 MMAs issued by one thread, completion through tcgen05.commit on an
// mbarrier, and the accumulator read back with tcgen05.ld.  passes = 1: plain TF32;
// passes = 3: error-compensated split (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo), fp32-grade accuracy.
// The fused coupling-layer kernels are built from exactly these pieces.
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc05.cuh"
#include "rnvp_mma.h"
#include "rnvp_philox.cuh"

namespace {
using namespace tc05;

// float offset of element (n, k) of an [N x K] K-major operand in the core-matrix tiled layout
__host__ __device__ inline int tiled_off(int n, int k, int K) { return (n >> 3) * (K >> 2) * 32 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }

__global__ void __launch_bounds__(128, 1) mma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                              float* __restrict__ D, int N, int K, int passes) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* Bhi = sm;
  float* Blo = sm + N * K;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int e = tid; e < N * K; e += 128) {
    const int n = e / K, k = e - n * K;
    uint32_t hi, lo;
    split_tf32(B[e], hi, lo);
    // passes 4 / 5: B stored MN-major (4 consecutive n contiguous, 8 k at 16 B stride, n-groups 128 B apart, k-groups after)
    const int off = passes >= 4 ? (n & 3) + (k & 7) * 4 + (n >> 2) * 32 + (k >> 3) * (N >> 2) * 32 : tiled_off(n, k, K);
    Bhi[off] = __uint_as_float(hi);
    Blo[off] = __uint_as_float(lo);
  }
  // generic-proxy smem writes must be visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t colA_hi = 0, colA_lo = 64, colD = 128;

  // A: thread r owns row r
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_tf32(A[tid * K + k0 + j], hi[j], lo[j]);
    tmem_st_x8(tbase + lane_base + colA_hi + k0, hi);
    tmem_st_x8(tbase + lane_base + colA_lo + k0, lo);
  }
  tmem_wait_st();
  fence_before_sync();
  __syncthreads();

  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = idesc_tf32(128, N);
    const uint32_t sbo = (uint32_t)(K >> 2) * 128u, lbo = 128u;
    uint32_t acc = 0;
    if (passes >= 4) {
      // MN-major B: passes 4 = (LBO = k-group stride, SBO = n-group stride) as in the CUTLASS canonical form, 5 = swapped
      const uint32_t kg = (uint32_t)(N >> 2) * 128u, ng = 128u;
      for (int j = 0; j < K / 8; ++j) {
        const uint64_t bdesc = smem_desc_kmajor_nosw(smem_u32(Bhi) + (uint32_t)j * kg, passes == 4 ? kg : ng, passes == 4 ? ng : kg);
        mma_tf32_ts(tbase + colD, tbase + colA_hi + 8 * j, bdesc, idesc | (1u << 16), acc);
        acc = 1;
      }
    }
    for (int p = 0; p < (passes >= 4 ? 0 : passes); ++p) {
      const uint32_t a_col = (p == 1) ? colA_lo : colA_hi;
      const float* Bp = (p == 2) ? Blo : Bhi;
      for (int j = 0; j < K / 8; ++j) {
        const uint64_t bdesc = smem_desc_kmajor_nosw(smem_u32(Bp) + (uint32_t)j * 256u, lbo, sbo);
        mma_tf32_ts(tbase + colD, tbase + a_col + 8 * j, bdesc, idesc, acc);
        acc = 1;
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    tmem_ld_x16(tbase + lane_base + colD + n0, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[tid * N + n0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// ============================================================ fused forward / inverse kernel
constexpr int MMA_THREADS = 352;        // warp 0 TMA producer, warp 1 / warp 10 MMA issuers of tile 0 / 1, warps 2-9 epilogue

template <int ACT>
__device__ __forceinline__ float act_mma(float v) {
  if (ACT == 1) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v));        // v = 2 log2(e) * pre-activation: the W1 image is pre-scaled (rnvp_planner.h)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
  }
  return fmaxf(v, 0.0f);
}

// trace slot `who` (0 / 1: epilogue group of tile 0 / 1, 2 / 3: their MMA issuers) of CTA 0: (tag, clock64), 2048 events per slot
__device__ __forceinline__ void trace_ev(const RnvpMmaArgs& a, int who, int& n, int tag) {
  if (a.trace && blockIdx.x == 0 && n < 2048) {
    a.trace[(who * 2048 + n) * 2] = tag;
    a.trace[(who * 2048 + n) * 2 + 1] = clock64();
    ++n;
  }
}

// exp(x) = 2^(x*log2e) on the MUFU with the argument's rounding error compensated: t = rn(x*log2e), r = x*log2e - t
// (exact residual through FMA + the low part of log2e), exp = ex2(t)*(1 + r*ln2).  ~2 ulp, 6 instructions; expf() costs
// ~25 and sat on the critical path of every coupling layer (16 per row and layer, forward and backward).
__device__ __forceinline__ float exp_mma(float x) {
  const float t = x * 1.4426950216293335f;                                   // log2e rounded to fp32
  const float r = fmaf(x, 1.9259629911266175e-8f, fmaf(x, 1.4426950216293335f, -t));   // + x * (log2e - fp32(log2e))
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return fmaf(e, r * 0.6931471805599453f, e);
}

enum { B_W1F = 0, B_W1E, B_W2F, B_W2E, B_UF0, B_UF1, B_D1F0, B_D1F1, B_AF0, B_AF1, B_D2F0, B_D2F1, B_COUNT };

template <int DH, int CDMAX, int CU, bool NETSEQ, int ACT, int MODE>
__global__ void __launch_bounds__(MMA_THREADS, 1) rnvp_mma_kernel(const __grid_constant__ RnvpMmaArgs a) {
  constexpr int K1PMAX = (DH + CDMAX + 1 + 7) & ~7;
  constexpr int NTP = (DH + 15) & ~15;
  // NETSEQ: the chunks of nn_t are processed before those of nn_s (t is parked in registers meanwhile), which halves
  // the accumulator columns so that D=64 flows also get the separate correction accumulator C2
  constexpr int D1W = NETSEQ ? CU : 2 * CU;          // columns of one GEMM1 accumulator chunk
  constexpr int D2W = NETSEQ ? NTP : 2 * NTP;        // columns of the GEMM2 accumulator(s) alive at a time
  constexpr int U_HI = 0, U_LO = K1PMAX, D1C = 2 * K1PMAX, A_LO = D1C + D1W, D2C = A_LO + D1W;
  // The tensor core accumulates in fp32 with TRUNCATION (measured: -0.3 ulp bias per accumulation, see
  // tools/mma_rounding_probe.py), so the tiny TF32-split correction products are kept out of the long main chains:
  // in GEMM1 they are issued first (while the accumulator is still tiny), in GEMM2 they get their own accumulator C2
  // (added in the epilogue) whenever the 512 TMEM columns allow it.
  constexpr bool USE_C2 = 2 * (D2C + 2 * D2W) <= 512;
  // GEMM2 accumulators per net: [D2 (NTP) | C2 (NTP)]; concurrent nets: nn_t's pair, then nn_s's pair
  constexpr int T_D2 = D2C, T_C2 = D2C + NTP, S_D2 = D2C + (NETSEQ ? 0 : 2 * NTP), S_C2 = S_D2 + NTP;
  constexpr int TILE_COLS = D2C + (USE_C2 ? 2 : 1) * D2W;
  static_assert(2 * TILE_COLS <= 512, "two row tiles must fit the 512 TMEM columns");
  static_assert((K1PMAX - DH) % 8 == 0 && DH % 8 == 0, "u is written in 8-column pieces");

  extern __shared__ __align__(128) float sm[];
  float* w1buf = sm;
  float* w2buf = sm + a.w1_floats;
  // w2buf also holds, after the W2 blocks, the b2 images: per net [hi | lo] of an [NTP x 8] operand whose only
  // non-zero column multiplies the constant-one column of u, so that GEMM2 starts from the bias
  float* w1tbuf = w2buf + a.w2_floats;                 // backward sweep only: W1T image (the W2T image reuses w2buf)
  uint64_t* bars = reinterpret_cast<uint64_t*>(w1tbuf + a.wt_floats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = a.H, Cd = a.Cd, D = 2 * DH;
  const int NC = H / CU;
  const int NCS = NETSEQ ? 2 * NC : NC;              // chunk steps per layer
  const int K1P = (DH + Cd + 1 + 7) & ~7;
  const int nL = a.l1 - a.l0;

  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bars[B_W1F], 1); mbar_init(&bars[B_W1E], 2); mbar_init(&bars[B_W2F], 1); mbar_init(&bars[B_W2E], 2);   // E: one commit per issuer
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bars[B_UF0 + g], 128); mbar_init(&bars[B_D1F0 + g], 1);
      mbar_init(&bars[B_AF0 + g], 128); mbar_init(&bars[B_D2F0 + g], 1);
    }
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const int my_pairs = (a.n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t ph1 = 0, ph2 = 0;
      long long step = 0;
      const uint32_t w1_bytes = (uint32_t)(4 * H * K1P) * 4u;          // actual image size for this Cd
      const uint32_t w2_bytes = (uint32_t)a.w2_floats * 4u;
      const int n_sweeps = (MODE == 2 && !NETSEQ && a.do_bwd) ? 2 : 1;   // fit step: forward sweep, then backward sweep
      for (int it = 0; it < my_pairs; ++it)
        for (int sw = 0; sw < n_sweeps; ++sw)
          for (int li = 0; li < nL; ++li, ++step) {
            const int i = (MODE == 1 || sw == 1) ? a.l1 - 1 - li : a.l0 + li;
            const float* src = a.wimg + (size_t)i * a.layer_floats;
            const uint32_t wt_bytes = (uint32_t)a.wt_floats * 4u;
            if (step > 0) { mbar_wait(&bars[B_W1E], ph1); ph1 ^= 1; }
            if (sw) {       // backward sweep: only the W1T image (nothing is recomputed, the W1 image is not needed)
              mbar_expect_tx(&bars[B_W1F], wt_bytes);
              bulk_g2s(w1tbuf, src + a.w1_floats + a.w2_floats + a.wt_floats, wt_bytes, &bars[B_W1F]);
            } else {
              mbar_expect_tx(&bars[B_W1F], w1_bytes);
              bulk_g2s(w1buf, src, w1_bytes, &bars[B_W1F]);
            }
            if (step > 0) { mbar_wait(&bars[B_W2E], ph2); ph2 ^= 1; }
            if (sw) {       // backward sweep: the W2T image takes the place of the W2 image
              mbar_expect_tx(&bars[B_W2F], wt_bytes);
              bulk_g2s(w2buf, src + a.w1_floats + a.w2_floats, wt_bytes, &bars[B_W2F]);
            } else {
              mbar_expect_tx(&bars[B_W2F], w2_bytes);
              bulk_g2s(w2buf, src + a.w1_floats, w2_bytes, &bars[B_W2F]);
            }
          }
    }
  } else if (warp == 1 || warp == 10) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 for tile 0, warp 10 for tile 1
    // (issuing costs ~2x the tensor-pipe time of these small MMAs, so one issuer for both tiles was the bottleneck).
    // The whole warp runs this role (warp-uniform control flow and descriptor arithmetic, so the descriptors
    // live in uniform registers); one elected lane issues the tcgen05.mma / tcgen05.commit instructions.
    const int g = warp == 1 ? 0 : 1;
    uint32_t ph_w1 = 0, ph_w2 = 0, ph_u = 0, ph_a = 0;
    const uint32_t idesc1 = idesc_tf32(128, D1W), idesc2 = idesc_tf32(128, NTP);
    const uint32_t lbo = (128u >> 4) << 16;
    const uint32_t w1_lo = ((smem_u32(w1buf) & 0x3FFFFu) >> 4) | lbo, w2_lo = ((smem_u32(w2buf) & 0x3FFFFu) >> 4) | lbo;
    const uint32_t hi1 = ((uint32_t)(K1P >> 2) * 128u >> 4) | (1u << 14);       // SBO, descriptor version 1
    const uint32_t hi2 = ((uint32_t)(CU >> 2) * 128u >> 4) | (1u << 14);
    const uint32_t hib = (256u >> 4) | (1u << 14);                              // b2 images: K = 8
    const uint32_t chunk1 = (uint32_t)(D1W * K1P) * 4u >> 4;                 // one hi (or lo) W1 chunk image, in 16 B units
    const uint32_t blk2 = (uint32_t)(NTP * CU) * 4u >> 4;                       // one hi (or lo) W2 (chunk, net) image
    const uint32_t b2_lo = w2_lo + ((uint32_t)(4 * NTP * H) * 4u >> 4);         // b2 images follow the W2 blocks
    const uint32_t blkb = (uint32_t)(NTP * 8) * 4u >> 4;
    const int nk1 = K1P >> 3;
    const int k_one = (DH + Cd) & ~7;                                           // u slice holding the constant one
    const bool leader = elect_one();
    auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    // GEMM1 chunk c of tile g: D1 = [u_lo*W1_hi + u_hi*W1_lo] + u_hi*W1_hi   (corrections first)
    auto gemm1 = [&](int g, int c) {
      const uint32_t tb = tbase + (uint32_t)(g * TILE_COLS);
      const uint32_t bh = w1_lo + (uint32_t)(2 * c) * chunk1, bl = bh + chunk1;
#pragma unroll
      for (int j = 0; j < K1PMAX / 8; ++j)
        if (j < nk1) mma_tf32_ts(tb + D1C, tb + U_LO + 8 * j, desc(bh + 16u * j, hi1), idesc1, j ? 1u : 0u);
#pragma unroll
      for (int j = 0; j < K1PMAX / 8; ++j)
        if (j < nk1) mma_tf32_ts(tb + D1C, tb + U_HI + 8 * j, desc(bl + 16u * j, hi1), idesc1, 1u);
#pragma unroll
      for (int j = 0; j < K1PMAX / 8; ++j)
        if (j < nk1) mma_tf32_ts(tb + D1C, tb + U_HI + 8 * j, desc(bh + 16u * j, hi1), idesc1, 1u);
    };
    // GEMM2 of chunk step cc of tile g.  Per net the accumulators sit side by side, [D2 | C2] (NTP columns each): main
    // products into D2 (seeded with b2_hi), split corrections into C2 (seeded with b2_lo).  The hi and lo images of a W2
    // block (and of b2) are adjacent in shared memory with the same row-group stride, so A_hi x [B_hi ; B_lo] is ONE MMA
    // of N = 2*NTP that feeds both accumulators: 2 instead of 3 instructions per K step (the issue cost of these small
    // MMAs, 10 + N/2 cycles, is what bounds the MMA phase).
    static_assert(USE_C2, "the merged GEMM2 needs the correction accumulator");
    const uint32_t idesc2m = idesc_tf32(128, 2 * NTP);
    auto gemm2_net = [&](uint32_t tb, int net, int c, uint32_t d2, uint32_t ah, uint32_t al) {
      const uint32_t bh = w2_lo + (uint32_t)((c * 2 + net) * 2) * blk2;
      const uint32_t c2 = d2 + NTP;
      if (c == 0)       // [D2 | C2] = 1 * [b2_hi ; b2_lo]
        mma_tf32_ts(d2, tb + U_HI + k_one, desc(b2_lo + (uint32_t)(net * 2) * blkb, hib), idesc2m, 0u);
#pragma unroll
      for (int j = 0; j < CU / 8; ++j) mma_tf32_ts(c2, al + 8 * j, desc(bh + 16u * j, hi2), idesc2, 1u);
#pragma unroll
      for (int j = 0; j < CU / 8; ++j) mma_tf32_ts(d2, ah + 8 * j, desc(bh + 16u * j, hi2), idesc2m, 1u);
    };
    auto gemm2 = [&](int g, int cc) {
      const uint32_t tb = tbase + (uint32_t)(g * TILE_COLS);
      if (NETSEQ) {
        const int net = cc >= NC ? 1 : 0;
        gemm2_net(tb, net, cc - net * NC, tb + D2C, tb + D1C, tb + A_LO);
      } else {
#pragma unroll
        for (int net = 0; net < 2; ++net)
          gemm2_net(tb, net, cc, tb + D2C + net * 2 * NTP, tb + D1C + net * CU, tb + A_LO + net * CU);
      }
    };
    for (int it = 0; it < my_pairs; ++it) {
      for (int li = 0; li < nL; ++li) {
        mbar_wait(&bars[B_W1F], ph_w1); ph_w1 ^= 1;
        mbar_wait(&bars[B_UF0 + g], ph_u); ph_u ^= 1;
        fence_after_sync();
        if (leader) {
          gemm1(g, 0);
          if (NCS == 1) mma_commit(&bars[B_W1E]);
          mma_commit(&bars[B_D1F0 + g]);
        }
        __syncwarp();
        for (int cc = 0; cc < NCS; ++cc) {
          mbar_wait(&bars[B_AF0 + g], ph_a); ph_a ^= 1;
          if (cc == 0) { mbar_wait(&bars[B_W2F], ph_w2); ph_w2 ^= 1; }
          fence_after_sync();
          if (leader) {
            gemm2(g, cc);
            if (cc + 1 < NCS) {
              gemm1(g, cc + 1);
              if (cc + 2 == NCS) mma_commit(&bars[B_W1E]);             // this tile's last GEMM1 of the layer issued
              mma_commit(&bars[B_D1F0 + g]);
            }
            if (cc + 1 == NCS) mma_commit(&bars[B_W2E]);               // this tile's last GEMM2 of the layer issued
            if (cc + 1 == NCS || (NETSEQ && cc + 1 == NC)) mma_commit(&bars[B_D2F0 + g]);
          }
          __syncwarp();
        }
      }
      if constexpr (MODE == 2 && !NETSEQ) {
        if (a.do_bwd) {
          // ---------------- backward sweep: per half-chunk of 16 units per net
          //   gemmA: DHB = delta2 W2                       (dh; h itself comes back from the record the forward sweep wrote)
          //   gemmB: DU += delta1 W1[:, x_K columns]       (du)
          // The transposed products read K-major images of W2^T (in w2buf during this sweep) and W1^T (w1tbuf): per
          // (half-chunk, net) a [16 x 16] block [hi | lo].  (tf32 operands cannot be read MN-major without swizzle.)
          constexpr int D1B = 2 * K1PMAX, DHB = D1B + 32, E2H = DHB + 32, E2L = E2H + 32, DUM = E2L + 32, DUC = DUM + 16;
          static_assert(DUC + 16 <= TILE_COLS, "backward TMEM map must fit the tile");
          static_assert(DH == 16, "transposed images are [16 x 16] blocks");
          const uint32_t idescB = idesc_tf32(128, 16);
          const uint32_t hit = (512u >> 4) | (1u << 14);                            // [16 x 16] blocks: SBO = 4 core matrices
          const uint32_t w1t_lo = ((smem_u32(w1tbuf) & 0x3FFFFu) >> 4) | lbo;
          const int NCB = H / 16;
          auto gemmA = [&](int g, int hc) {
            const uint32_t tb = tbase + (uint32_t)(g * TILE_COLS);
#pragma unroll
            for (int net = 0; net < 2; ++net) {
              const uint32_t bh = w2_lo + (uint32_t)((hc * 2 + net) * 2) * 64u, bl = bh + 64u;
              const uint32_t dst = tb + DHB + 16 * net;
#pragma unroll
              for (int j = 0; j < 2; ++j)
                mma_tf32_ts(dst, tb + E2L + 16 * net + 8 * j, desc(bh + 16u * j, hit), idescB, j ? 1u : 0u);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                mma_tf32_ts(dst, tb + E2H + 16 * net + 8 * j, desc(bl + 16u * j, hit), idescB, 1u);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                mma_tf32_ts(dst, tb + E2H + 16 * net + 8 * j, desc(bh + 16u * j, hit), idescB, 1u);
            }
          };
          auto gemmB = [&](int g, int hc) {
            const uint32_t tb = tbase + (uint32_t)(g * TILE_COLS);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {                                       // (net, K half) of this half-chunk
              const uint32_t bh = w1t_lo + (uint32_t)((hc * 2 + (ks >> 1)) * 2) * 64u + 16u * (ks & 1), bl = bh + 64u;
              const uint32_t acol = 16 * (ks >> 1) + 8 * (ks & 1);
              mma_tf32_ts(tb + DUC, tb + D1B + acol, desc(bh, hit), idescB, (hc | ks) ? 1u : 0u);   // delta1_lo * W1_hi
              mma_tf32_ts(tb + DUC, tb + DHB + acol, desc(bl, hit), idescB, 1u);                    // delta1_hi * W1_lo
              mma_tf32_ts(tb + DUM, tb + DHB + acol, desc(bh, hit), idescB, (hc | ks) ? 1u : 0u);   // main
            }
          };
          int ntr = 0;
          for (int li = 0; li < nL; ++li) {
            mbar_wait(&bars[B_W1F], ph_w1); ph_w1 ^= 1;
            mbar_wait(&bars[B_W2F], ph_w2); ph_w2 ^= 1;
            if (leader) trace_ev(a, 2 + g, ntr, 200 + li);
            mbar_wait(&bars[B_UF0 + g], ph_u); ph_u ^= 1;
            fence_after_sync();
            if (leader) {
              trace_ev(a, 2 + g, ntr, 300);
              gemmA(g, 0);
              if (NCB == 1) mma_commit(&bars[B_W2E]);
              mma_commit(&bars[B_D1F0 + g]);
            }
            __syncwarp();
            for (int hc = 0; hc < NCB; ++hc) {
              mbar_wait(&bars[B_AF0 + g], ph_a); ph_a ^= 1;
              fence_after_sync();
              if (leader) {
                trace_ev(a, 2 + g, ntr, 400 + hc);
                gemmB(g, hc);
                if (hc + 1 < NCB) {
                  gemmA(g, hc + 1);
                  if (hc + 2 == NCB) mma_commit(&bars[B_W2E]);           // this tile's last dh of the layer issued
                  mma_commit(&bars[B_D1F0 + g]);
                } else {
                  mma_commit(&bars[B_W1E]);                              // this tile's last du of the layer issued
                  mma_commit(&bars[B_D2F0 + g]);
                }
                trace_ev(a, 2 + g, ntr, 500 + hc);
              }
              __syncwarp();
            }
          }
        }
      }
    }   // pairs
  } else {
    // ------------------------------------------------------------------ epilogue: one thread per row
    const int g = (warp - 2) >> 2;                      // tile of the pair
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const uint32_t trow = tbase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * TILE_COLS);
    uint32_t ph_d1 = 0, ph_d2 = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const long long pair = (long long)blockIdx.x + (long long)it * gridDim.x;
      const long long row = pair * 256 + g * 128 + quarter * 32 + lane;
      const bool valid = row < a.N;
      const long long src = valid ? (a.idx ? a.idx[row] : row) : 0;
      float xa[DH], xb[DH], cc[CDMAX], ld = 0.0f;       // even features, odd features, condition
#pragma unroll
      for (int m = 0; m < DH / 2; ++m) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 1 && a.X == nullptr) { if (valid) v = rnvp_rng::normal4(a.seed, a.row_offset + row, m); }   // sampling: in-kernel prior draw
        else if (valid) v = __ldg(reinterpret_cast<const float4*>(a.X + src * D) + m);
        xa[2 * m] = v.x; xb[2 * m] = v.y; xa[2 * m + 1] = v.z; xb[2 * m + 1] = v.w;
      }
#pragma unroll
      for (int k = 0; k < CDMAX; ++k) cc[k] = (valid && k < Cd) ? __ldg(a.C + src * Cd + k) : 0.0f;
      // static part of u: [c | 1 | 0...] at columns DH.. of U_HI / U_LO (written once per tile)
#pragma unroll
      for (int e0 = 0; e0 < K1PMAX - DH; e0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = e0 + j;
          float v = 0.0f;
          if (k < CDMAX && k < Cd) v = cc[k < CDMAX ? k : 0];
          if (k == Cd) v = 1.0f;
          split_tf32(v, hi[j], lo[j]);
        }
        tmem_st_x8(trow + U_HI + DH + e0, hi);
        tmem_st_x8(trow + U_LO + DH + e0, lo);
      }

      // activation records of (layer i, this thread's row): [layer][block of 32 rows][column group of 4][32 slots][4 floats],
      // slot = (row % 32) ^ (group & 1): a warp-level float4 access covers 512 contiguous bytes, and the weight-gradient
      // sweep's mma fragment loads of a block are bank-conflict free (rnvp_wgrad.cu).  Returns the block's base.
      auto rec_base = [&](int i) -> float* {
        return a.records + (((size_t)i * (size_t)(a.Npad >> 5) + (size_t)(row >> 5)) * (size_t)(a.rec >> 2)) * 128;
      };
      auto layer = [&](float (&xT)[DH], float (&xK)[DH], int i) {
        // ---- u (conditioning half) -> TMEM
#pragma unroll
        for (int e0 = 0; e0 < DH; e0 += 8) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_tf32(xK[e0 + j], hi[j], lo[j]);
          tmem_st_x8(trow + U_HI + e0, hi);
          tmem_st_x8(trow + U_LO + e0, lo);
        }
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&bars[B_UF0 + g]);
        // ---- hidden chunks: D1 -> act -> (A_hi in place, A_lo); NETSEQ: nn_t chunks, park t, nn_s chunks
        float tpark[NETSEQ ? DH : 1];
        for (int cc = 0; cc < NCS; ++cc) {
          mbar_wait(&bars[B_D1F0 + g], ph_d1); ph_d1 ^= 1;
          fence_after_sync();
          if (DH >= 32) {                        // register budget: 16-column pieces for the wide rows
#pragma unroll
            for (int q0 = 0; q0 < D1W; q0 += 16) {
              uint32_t r[16], lo[16];
              tmem_ld_x16(trow + D1C + q0, r);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float h = act_mma<ACT>(__uint_as_float(r[j]));
                split_tf32(h, r[j], lo[j]);
              }
              tmem_st_x16(trow + D1C + q0, r);
              tmem_st_x16(trow + A_LO + q0, lo);
            }
          } else {
#pragma unroll
            for (int q0 = 0; q0 < D1W; q0 += 32) {
              uint32_t r[32], lo[32];
              tmem_ld_x32(trow + D1C + q0, r);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(act_mma<ACT>(__uint_as_float(r[j])));
              if (MODE == 2 && !NETSEQ && a.do_bwd) {
                // fit step: h goes to the activation record of (layer, row) now -- the backward sweep and the weight-gradient
                // sweep read it back instead of recomputing u W1^T and the tanh.  Columns q0.. of a chunk are units
                // cc*CU.. of net q0/CU (D1 = [nn_t chunk | nn_s chunk]).
                float* hb = rec_base(i) + (((q0 / CU) * H + cc * CU) >> 2) * 128;
#pragma unroll
                for (int m = 0; m < 8; ++m)
                  *reinterpret_cast<float4*>(hb + m * 128 + ((lane ^ (m & a.rec_swz)) << 2)) =
                      make_float4(__uint_as_float(r[4 * m]), __uint_as_float(r[4 * m + 1]), __uint_as_float(r[4 * m + 2]),
                                  __uint_as_float(r[4 * m + 3]));
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) split_tf32(__uint_as_float(r[j]), r[j], lo[j]);
              tmem_st_x32(trow + D1C + q0, r);
              tmem_st_x32(trow + A_LO + q0, lo);
            }
          }
          tmem_wait_st();
          fence_before_sync();
          mbar_arrive(&bars[B_AF0 + g]);
          if (NETSEQ && cc + 1 == NC) {          // nn_t complete: t = D2 + C2 into registers
            mbar_wait(&bars[B_D2F0 + g], ph_d2); ph_d2 ^= 1;
            fence_after_sync();
#pragma unroll
            for (int e0 = 0; e0 < DH; e0 += 16) {
              uint32_t tv[16], tc[16];
              tmem_ld_x16(trow + T_D2 + e0, tv);
              if (USE_C2) tmem_ld_x16(trow + T_C2 + e0, tc);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                tpark[NETSEQ ? e0 + j : 0] = USE_C2 ? __uint_as_float(tv[j]) + __uint_as_float(tc[j]) : __uint_as_float(tv[j]);
            }
            fence_before_sync();                 // the nn_s GEMM2 will overwrite D2 / C2 after the next a_full arrival
          }
        }
        // ---- t, s -> coupling
        mbar_wait(&bars[B_D2F0 + g], ph_d2); ph_d2 ^= 1;
        fence_after_sync();
#pragma unroll
        for (int e0 = 0; e0 < DH; e0 += 16) {
          uint32_t tv[16], sv[16], tc[16], sc[16];
          if (!NETSEQ) tmem_ld_x16(trow + T_D2 + e0, tv);
          tmem_ld_x16(trow + S_D2 + e0, sv);
          if (USE_C2) {
            if (!NETSEQ) tmem_ld_x16(trow + T_C2 + e0, tc);
            tmem_ld_x16(trow + S_C2 + e0, sc);
          }
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t;
            if (NETSEQ) t = tpark[NETSEQ ? e0 + j : 0];
            else t = USE_C2 ? __uint_as_float(tv[j]) + __uint_as_float(tc[j]) : __uint_as_float(tv[j]);
            const float s = USE_C2 ? __uint_as_float(sv[j]) + __uint_as_float(sc[j]) : __uint_as_float(sv[j]);
            if (MODE == 2) { sv[j] = __float_as_uint(s); tv[j] = __float_as_uint(xT[e0 + j]); }   // stash s and x_T
            if (MODE != 1) { xT[e0 + j] = fmaf(xT[e0 + j], exp_mma(s), t); ld += s; }
            else xT[e0 + j] = (xT[e0 + j] - t) * exp_mma(-s);
          }
          if (MODE == 2 && !NETSEQ && a.do_bwd) {
            // stash for this kernel's own backward sweep: blocks of 32 rows, [block][layer][float4 group][32 rows][4] --
            // a warp stores / loads 512 contiguous bytes per instruction (row-major cost 32 L1 wavefronts each);
            // padding rows are written too (finite values), so the backward sweep reads unconditionally
            float* sb = a.stash + (((size_t)(row >> 5) * a.L_total + i) * (2 * DH / 4)) * 128 + lane * 4;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              *reinterpret_cast<float4*>(sb + (e0 / 4 + m) * 128) =
                  make_float4(__uint_as_float(tv[4 * m]), __uint_as_float(tv[4 * m + 1]), __uint_as_float(tv[4 * m + 2]),
                              __uint_as_float(tv[4 * m + 3]));
              *reinterpret_cast<float4*>(sb + (DH / 4 + e0 / 4 + m) * 128) =
                  make_float4(__uint_as_float(sv[4 * m]), __uint_as_float(sv[4 * m + 1]), __uint_as_float(sv[4 * m + 2]),
                              __uint_as_float(sv[4 * m + 3]));
            }
          } else if (MODE == 2 && valid) {    // row-major stash [N][L][x_T | s] for the FP32 backward sweep (rnvp_tile.cu)
            float4* sp = reinterpret_cast<float4*>(a.stash + ((size_t)row * a.L_total + i) * (2 * DH) + e0);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              sp[m] = make_float4(__uint_as_float(tv[4 * m]), __uint_as_float(tv[4 * m + 1]), __uint_as_float(tv[4 * m + 2]),
                                  __uint_as_float(tv[4 * m + 3]));
              sp[DH / 4 + m] = make_float4(__uint_as_float(sv[4 * m]), __uint_as_float(sv[4 * m + 1]),
                                           __uint_as_float(sv[4 * m + 2]), __uint_as_float(sv[4 * m + 3]));
            }
          }
        }
      };
      for (int li = 0; li < nL; ++li) {
        const int i = MODE != 1 ? a.l0 + li : a.l1 - 1 - li;
        if ((i & 1) == 0) layer(xa, xb, i);              // even layer transforms the even features
        else layer(xb, xa, i);
      }

      if (valid) {
        if (a.out_x) {
#pragma unroll
          for (int m = 0; m < DH / 2; ++m)
            reinterpret_cast<float4*>(a.out_x + row * D)[m] = make_float4(xa[2 * m], xb[2 * m], xa[2 * m + 1], xb[2 * m + 1]);
        }
      }
      if (MODE != 1) {
        float q = 0.0f;
#pragma unroll
        for (int e = 0; e < DH; ++e) { q = fmaf(xa[e], xa[e], q); q = fmaf(xb[e], xb[e], q); }
        float lp = valid ? ld - 0.5f * (D * 1.8378770664093453f + q) : 0.0f;
        if (valid) {
          if (a.out_logdet) a.out_logdet[row] = ld;
          if (a.out_logp) a.out_logp[row] = lp;
        }
        if (MODE == 2 && a.loss_sum) {
#pragma unroll
          for (int m = 16; m >= 1; m >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, m);
          if (lane == 0) atomicAdd(a.loss_sum, lp);
        }
      }

      // ================================================================ backward sweep (fit step)
      if constexpr (MODE == 2 && !NETSEQ) {
        if (a.do_bwd) {
          constexpr int D1B = 2 * K1PMAX, DHB = D1B + 32, E2H = DHB + 32, E2L = E2H + 32, DUM = E2L + 32, DUC = DUM + 16;
          const int NCB = H / 16;
          const int K1P8 = (DH + Cd + 7) & ~7;
          const long long rloc = pair * 256 + g * 128 + quarter * 32 + lane;       // row inside the padded batch
          // gradient of scale*sum_rows logp w.r.t. the current activations: g_z = -scale*z, g_logdet = scale
          const float gld = valid ? a.scale : 0.0f;
          float ga[DH], gb[DH];
#pragma unroll
          for (int e = 0; e < DH; ++e) { ga[e] = -gld * xa[e]; gb[e] = -gld * xb[e]; }

          int ntr = 0;
          const bool tracer = quarter == 0 && lane == 0;

          auto layer_bwd = [&](float (&xT)[DH], float (&xK)[DH], float (&gT)[DH], float (&gK)[DH], int i) {
            if (tracer) trace_ev(a, g, ntr, 100 + i);
            if (i > a.l0) {       // the next layer's stash block (32 rows x 2*DH floats, contiguous) is needed in ~25 k cycles
              const float* nx = a.stash + (((size_t)(row >> 5) * a.L_total + (i - 1)) * (2 * DH / 4)) * 128 + lane * 32;
#pragma unroll
              for (int q = 0; q < 2 * DH / 32; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + q * 1024));
            }
            float* recb = rec_base(i);
            auto rec_st = [&](int col, float4 v) {
              const int cg = col >> 2;
              *reinterpret_cast<float4*>(recb + cg * 128 + ((lane ^ (cg & a.rec_swz)) << 2)) = v;
            };
            // ---- x_T and s of this layer from the forward stash; delta2 and the new g_T
            uint32_t e2h[2 * DH], e2l[2 * DH];
            {
              const float* sb = a.stash + (((size_t)(row >> 5) * a.L_total + i) * (2 * DH / 4)) * 128 + lane * 4;
#pragma unroll
              for (int m = 0; m < DH / 4; ++m) {
                const float4 xv = *reinterpret_cast<const float4*>(sb + m * 128);
                const float4 sv = *reinterpret_cast<const float4*>(sb + (DH / 4 + m) * 128);
                const float xs4[4] = {xv.x, xv.y, xv.z, xv.w}, ss4[4] = {sv.x, sv.y, sv.z, sv.w};
                float d2t[4], d2s[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int e = 4 * m + q;
                  const float es = exp_mma(ss4[q]);
                  xT[e] = xs4[q];                                   // input of this layer (T half)
                  d2t[q] = gT[e];                                   // dL/dt
                  d2s[q] = fmaf(gT[e] * xs4[q], es, gld);           // dL/ds = g_y*x*exp(s) + g_logdet
                  gT[e] *= es;                                      // dL/dx_T
                  split_tf32(d2t[q], e2h[e], e2l[e]);
                  split_tf32(d2s[q], e2h[DH + e], e2l[DH + e]);
                }
                rec_st(2 * H + K1P8 + 4 * m, make_float4(d2t[0], d2t[1], d2t[2], d2t[3]));
                rec_st(2 * H + K1P8 + DH + 4 * m, make_float4(d2s[0], d2s[1], d2s[2], d2s[3]));
              }
            }
            if (tracer) trace_ev(a, g, ntr, 3);
#pragma unroll
            for (int e0 = 0; e0 < 2 * DH; e0 += 8) {
              uint32_t th[8], tl[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { th[j] = e2h[e0 + j]; tl[j] = e2l[e0 + j]; }
              tmem_st_x8(trow + E2H + e0, th);
              tmem_st_x8(trow + E2L + e0, tl);
            }
            // ---- u = [x_K, c] -> record (the weight-gradient sweep needs it; nothing is recomputed here any more)
#pragma unroll
            for (int m = 0; m < DH / 4; ++m)
              rec_st(2 * H + 4 * m, make_float4(xK[4 * m], xK[4 * m + 1], xK[4 * m + 2], xK[4 * m + 3]));
#pragma unroll
            for (int m = 0; m < CDMAX / 4; ++m)
              if (DH + 4 * m < K1P8)
                rec_st(2 * H + DH + 4 * m, make_float4(cc[4 * m], cc[4 * m + 1], cc[4 * m + 2], cc[4 * m + 3]));
            if (tracer) trace_ev(a, g, ntr, 4);
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(&bars[B_UF0 + g]);
            if (tracer) trace_ev(a, g, ntr, 2);
            // ---- half-chunks: delta1 = dh * act'(h) with dh from the MMA warp and h from the record; delta1 hi/lo -> TMEM (A operand of du)
            for (int hc = 0; hc < NCB; ++hc) {
              // h of this half-chunk (units 16hc.. of both nets) from the record the forward sweep wrote: issued before
              // the wait for dh, so the L2 / HBM latency hides behind the MMAs
              uint32_t pa_h[32];
#pragma unroll
              for (int net = 0; net < 2; ++net)
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                  const int cg = (net * H + 16 * hc) / 4 + m;
                  const float4 v = *reinterpret_cast<const float4*>(recb + cg * 128 + ((lane ^ (cg & a.rec_swz)) << 2));
                  pa_h[16 * net + 4 * m] = __float_as_uint(v.x); pa_h[16 * net + 4 * m + 1] = __float_as_uint(v.y);
                  pa_h[16 * net + 4 * m + 2] = __float_as_uint(v.z); pa_h[16 * net + 4 * m + 3] = __float_as_uint(v.w);
                }
              mbar_wait(&bars[B_D1F0 + g], ph_d1); ph_d1 ^= 1;
              fence_after_sync();
              if (tracer) trace_ev(a, g, ntr, 10 + hc);
              uint32_t pa[32], dh[32];
              tmem_ld_x32(trow + DHB, dh);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float h = __uint_as_float(pa_h[j]);
                const float dp = ACT == 1 ? fmaf(-h, h, 1.0f) : (h > 0.0f ? 1.0f : 0.0f);
                dh[j] = __float_as_uint(__uint_as_float(dh[j]) * dp);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) split_tf32(__uint_as_float(dh[j]), dh[j], pa[j]);     // hi -> dh, lo -> pa
              tmem_st_x32(trow + DHB, dh);
              tmem_st_x32(trow + D1B, pa);
              tmem_wait_st();
              fence_before_sync();
              mbar_arrive(&bars[B_AF0 + g]);
              if (tracer) trace_ev(a, g, ntr, 30 + hc);
            }
            // ---- du -> g_x_K
            mbar_wait(&bars[B_D2F0 + g], ph_d2); ph_d2 ^= 1;
            fence_after_sync();
            if (tracer) trace_ev(a, g, ntr, 60);
            {
              uint32_t um[16], uc[16];
              tmem_ld_x16(trow + DUM, um);
              tmem_ld_x16(trow + DUC, uc);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < DH; ++e) gK[e] += __uint_as_float(um[e]) + __uint_as_float(uc[e]);
            }
          };
          for (int li = 0; li < nL; ++li) {
            const int i = a.l1 - 1 - li;
            if ((i & 1) == 0) layer_bwd(xa, xb, ga, gb, i);
            else layer_bwd(xb, xa, gb, ga, i);
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

template <int DH, int CDMAX, int CU, bool NETSEQ>
cudaError_t launch_mma_shape(int act, int mode, const RnvpMmaArgs& a, int grid, size_t smem, cudaStream_t st) {
#define RNVP_MMA_LAUNCH(ACT_, MODE_)                                                                        \
  {                                                                                                          \
    auto k = rnvp_mma_kernel<DH, CDMAX, CU, NETSEQ, ACT_, MODE_>;                                                    \
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    if (e != cudaSuccess) return e;                                                                          \
    k<<<grid, MMA_THREADS, smem, st>>>(a);                                                                   \
    return cudaGetLastError();                                                                               \
  }
  if (act == 1 && mode == 0) RNVP_MMA_LAUNCH(1, 0)
  if (act == 1 && mode == 1) RNVP_MMA_LAUNCH(1, 1)
  if (act == 1 && mode == 2) RNVP_MMA_LAUNCH(1, 2)
  if (act == 2 && mode == 0) RNVP_MMA_LAUNCH(2, 0)
  if (act == 2 && mode == 1) RNVP_MMA_LAUNCH(2, 1)
  if (act == 2 && mode == 2) RNVP_MMA_LAUNCH(2, 2)
#undef RNVP_MMA_LAUNCH
  return cudaErrorInvalidValue;
}

}  // namespace

size_t rnvp_mma_smem_bytes(int w1_floats, int w2_floats, int w1t_floats) {
  return (size_t)(w1_floats + w2_floats + w1t_floats) * 4 + 8 * B_COUNT + 64;
}

cudaError_t rnvp_launch_mma(int DH, int act, int mode, const RnvpMmaArgs& a, int grid, size_t smem, cudaStream_t st) {
  if (DH == 16) return launch_mma_shape<16, 8, 32, false>(act, mode, a, grid, smem, st);
  if (DH == 32) return launch_mma_shape<32, 16, 32, true>(act, mode, a, grid, smem, st);
  return cudaErrorInvalidValue;
}

cudaError_t rnvp_launch_mma_selftest(const float* A, const float* B, float* D, int N, int K, int passes, cudaStream_t st) {
  const size_t smem = (size_t)2 * N * K * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  mma_selftest_kernel<<<1, 128, smem, st>>>(A, B, D, N, K, passes);
  return cudaGetLastError();
}
