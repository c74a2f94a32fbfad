// Row-per-thread RealNVP kernels for SMALL flows (README / moons shapes: D<=8, Cd<=4, one hidden
// layer): forward (log-density), inverse (sampling) and the fit step (fused forward + backward).
//
// For D=2, H=10 a coupling layer is ~120 FMAs per row: tile machinery, barriers and shared-memory
// round trips would dominate, so here one thread owns RPT whole rows in registers, all coupling
// layers are walked in one launch, and the weights of the entire flow (a few KB) sit in shared
// memory as per-hidden-unit records [w1_x | w1_c | b1 | w2] that every lane reads at the same
// address (broadcast LDS.128, no bank conflicts).  The binding unit is the MUFU/FMA pipe of the
// tanh (2 MUFU + 3 FMA each), not HBM: 16 B/row in, 4..12 B/row out.
//
// Reference semantics: RealNVPLayer.f / .g (realnvp.py:73-129) looped as in nflow.py:109-115 /
// 142-143; masks (arange(D)+i)%2, so even layers transform the even-indexed features (xe) and
// condition on the odd ones (xo), odd layers the other way round.
#include <cuda_runtime.h>
#include <stdint.h>
#include "rnvp_small.h"
#include "rnvp_philox.cuh"

namespace {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float act_f(float v) {
  if (ACT == 1) {
    const float e = ex2_approx(v * 2.8853900817779268f);   // tanh(v) = 1 - 2/(exp(2v)+1)
    return fmaf(-2.0f, rcp_approx(e + 1.0f), 1.0f);
  }
  return fmaxf(v, 0.0f);
}

constexpr int RPT = 4;          // rows per thread
constexpr int THREADS = 256;

// t and s of one coupling layer for RPT rows.  xk: conditioning half, c: condition.
template <int NE, int NC, int ACT>
__device__ __forceinline__ void conditioner_pair(const float* __restrict__ wl, int H, int rec,
                                                 const float (&xk)[RPT][NE], const float (&c)[RPT][NC > 0 ? NC : 1],
                                                 float (&t)[RPT][NE], float (&s)[RPT][NE]) {
  const int net_floats = H * rec + ((NE + 3) & ~3);
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    const float* w = wl + net * net_floats;
    float acc[RPT][NE];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[r][e] = w[H * rec + e];          // b2
    for (int j = 0; j < H; ++j) {
      const float* u = w + j * rec;                                      // [w1x NE | w1c NC | b1 | w2 NE]
      float rv[2 * NE + NC + 1];
      constexpr int NV = (2 * NE + NC + 1 + 3) / 4;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(u + 4 * q);
        if (4 * q + 0 < 2 * NE + NC + 1) rv[4 * q + 0] = v.x;
        if (4 * q + 1 < 2 * NE + NC + 1) rv[4 * q + 1] = v.y;
        if (4 * q + 2 < 2 * NE + NC + 1) rv[4 * q + 2] = v.z;
        if (4 * q + 3 < 2 * NE + NC + 1) rv[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        float a = rv[NE + NC];
#pragma unroll
        for (int e = 0; e < NE; ++e) a = fmaf(rv[e], xk[r][e], a);
#pragma unroll
        for (int k = 0; k < NC; ++k) a = fmaf(rv[NE + k], c[r][k], a);
        const float h = act_f<ACT>(a);
#pragma unroll
        for (int e = 0; e < NE; ++e) acc[r][e] = fmaf(rv[NE + NC + 1 + e], h, acc[r][e]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (net == 0) t[r][e] = acc[r][e];
        else s[r][e] = acc[r][e];
      }
  }
}

template <int NE, int NC, int ACT, int MODE>
__global__ void __launch_bounds__(THREADS) rnvp_small_kernel(const RnvpSmallArgs a) {
  extern __shared__ __align__(16) float wsm[];
  const int D = a.D, Cd = a.Cd, H = a.H, rec = a.rec;
  for (int i = threadIdx.x * 4; i < a.small_floats; i += THREADS * 4)
    *reinterpret_cast<float4*>(wsm + i) = *reinterpret_cast<const float4*>(a.packed_small + i);
  __syncthreads();
  const int layer_floats = 2 * (H * rec + ((NE + 3) & ~3));

  const long long rows_per_block = (long long)THREADS * RPT;
  for (long long base = (long long)blockIdx.x * rows_per_block; base < a.N; base += (long long)gridDim.x * rows_per_block) {
    float xe[RPT][NE], xo[RPT][NE], c[RPT][NC > 0 ? NC : 1], ld[RPT];
    long long row[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      row[r] = base + threadIdx.x + (long long)r * THREADS;
      const bool ok = row[r] < a.N;
      const long long src = ok ? (a.idx ? a.idx[row[r]] : row[r]) : 0;
      ld[r] = 0.0f;
      if (MODE == 1 && a.X == nullptr) {      // sampling: prior draws generated here, keyed on the global row index
#pragma unroll
        for (int q = 0; q < (2 * NE + 3) / 4; ++q) {
          const float4 v = rnvp_rng::normal4(a.seed, a.row_offset + row[r], q);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int j = 4 * q + m;
            if (j < 2 * NE) {
              if (j & 1) xo[r][j >> 1] = (ok && j < D) ? vv[m] : 0.0f;
              else xe[r][j >> 1] = (ok && j < D) ? vv[m] : 0.0f;
            }
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          xe[r][e] = (ok && 2 * e < D) ? __ldg(a.X + src * D + 2 * e) : 0.0f;
          xo[r][e] = (ok && 2 * e + 1 < D) ? __ldg(a.X + src * D + 2 * e + 1) : 0.0f;
        }
      }
#pragma unroll
      for (int k = 0; k < (NC > 0 ? NC : 1); ++k) c[r][k] = (NC > 0 && ok && k < Cd) ? __ldg(a.C + src * Cd + k) : 0.0f;
    }

    float t[RPT][NE], s[RPT][NE];
    if (MODE == 0) {
      for (int i = a.l0; i < a.l1; ++i) {
        const float* wl = wsm + i * layer_floats;
        if ((i & 1) == 0) {            // even layer: T = even features, K = odd features
          conditioner_pair<NE, NC, ACT>(wl, H, rec, xo, c, t, s);
#pragma unroll
          for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int e = 0; e < NE; ++e) { xe[r][e] = fmaf(xe[r][e], expf(s[r][e]), t[r][e]); ld[r] += s[r][e]; }
        } else {
          conditioner_pair<NE, NC, ACT>(wl, H, rec, xe, c, t, s);
#pragma unroll
          for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int e = 0; e < NE; ++e) { xo[r][e] = fmaf(xo[r][e], expf(s[r][e]), t[r][e]); ld[r] += s[r][e]; }
        }
      }
    } else {
      for (int i = a.l1 - 1; i >= a.l0; --i) {
        const float* wl = wsm + i * layer_floats;
        if ((i & 1) == 0) {
          conditioner_pair<NE, NC, ACT>(wl, H, rec, xo, c, t, s);
#pragma unroll
          for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int e = 0; e < NE; ++e) xe[r][e] = (xe[r][e] - t[r][e]) * expf(-s[r][e]);
        } else {
          conditioner_pair<NE, NC, ACT>(wl, H, rec, xe, c, t, s);
#pragma unroll
          for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int e = 0; e < NE; ++e) xo[r][e] = (xo[r][e] - t[r][e]) * expf(-s[r][e]);
        }
      }
    }

#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (row[r] >= a.N) continue;
      if (a.out_x) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (2 * e < D) a.out_x[row[r] * D + 2 * e] = xe[r][e];
          if (2 * e + 1 < D) a.out_x[row[r] * D + 2 * e + 1] = xo[r][e];
        }
      }
      if (MODE == 0) {
        float q = 0.0f;
#pragma unroll
        for (int e = 0; e < NE; ++e) {      // padded features are exactly 0 and add nothing
          q = fmaf(xe[r][e], xe[r][e], q);
          q = fmaf(xo[r][e], xo[r][e], q);
        }
        if (a.out_logdet) a.out_logdet[row[r]] = ld[r];
        if (a.out_logp) a.out_logp[row[r]] = ld[r] - 0.5f * (D * 1.8378770664093453f + q);
      }
    }
  }
}

// ============================================================ fit step for small flows (fused forward + backward)
//
// One thread owns one row through the whole flow: forward sweep (stashing x_T and s per layer in local memory), then
// the backward sweep of d(scale * sum logp)/d(theta) with the hidden activations recomputed per unit (H is ~10).
// Fit step of a small flow.  32 rows per CTA, FOUR warps: warp w = (net = w & 1, half = w >> 1) evaluates half of the hidden
// units of ONE conditioner for all 32 rows (lane = row); the partial t / s (forward) and du (backward) of the four warps
// meet through a double-buffered shared-memory exchange and one __syncthreads per layer, after which every warp applies
// the (cheap) coupling arithmetic redundantly, so all four hold the same row state.  A README-sized step (32 rows, 8 layers,
// H = 10) is latency-bound on ONE warp's dependent instruction stream (IPC ~0.3); splitting the units over four warps cuts
// that stream ~3x.  Weight gradients are sums over rows: the per-row contributions of up to 32 gradient entries are reduced
// over the warp with one halving butterfly and added by the owning lane to a shared-memory copy of the gradient (each
// warp touches only its own net's / units' entries), flushed once through the small-layout -> packed-gradient table.
constexpr int FIT_WARPS = 4;
constexpr int FIT_THREADS = 32 * FIT_WARPS;
constexpr int FIT_ROWS = 32;       // rows per CTA
constexpr int FIT_MAXL = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

template <int NE, int NC>
__device__ __forceinline__ void load_record(const float* __restrict__ u, float (&rv)[2 * NE + NC + 1]) {
  constexpr int NV = (2 * NE + NC + 1 + 3) / 4;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(u + 4 * q);
    if (4 * q + 0 < 2 * NE + NC + 1) rv[4 * q + 0] = v.x;
    if (4 * q + 1 < 2 * NE + NC + 1) rv[4 * q + 1] = v.y;
    if (4 * q + 2 < 2 * NE + NC + 1) rv[4 * q + 2] = v.z;
    if (4 * q + 3 < 2 * NE + NC + 1) rv[4 * q + 3] = v.w;
  }
}

// this warp's share of the work of a layer: conditioner `net`, hidden units [j_lo, j_hi)
struct FitShare {
  int net, j_lo, j_hi, wid, lane;
  bool first_half;               // adds b2 (forward) / reduces the b2 gradient (backward)
};

// forward of one coupling layer (same arithmetic as conditioner_pair + the MODE 0 update; the sum over hidden units is split
// into the two halves of the warps, then t = half0 + half1)
template <int NE, int NC, int ACT>
__device__ __forceinline__ void fit_fwd_layer(const float* __restrict__ wl, int H, int rec, const FitShare& sh, float* __restrict__ xch,
                                              float (&xT)[NE], const float (&xK)[NE], const float (&c)[NC > 0 ? NC : 1],
                                              float* __restrict__ st_x, float* __restrict__ st_s, float& ld) {
  const int net_floats = H * rec + ((NE + 3) & ~3);
  const float* w = wl + sh.net * net_floats;
  float tp[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) tp[e] = sh.first_half ? w[H * rec + e] : 0.0f;          // b2
  constexpr int FU = 5;          // units evaluated side by side: their dependent chains (record -> dot -> tanh -> W2) overlap
  for (int j0 = sh.j_lo; j0 < sh.j_hi; j0 += FU) {
    float h[FU], rv[FU][2 * NE + NC + 1];
#pragma unroll
    for (int jj = 0; jj < FU; ++jj) load_record<NE, NC>(w + min(j0 + jj, sh.j_hi - 1) * rec, rv[jj]);
#pragma unroll
    for (int jj = 0; jj < FU; ++jj) {
      float a = rv[jj][NE + NC];
#pragma unroll
      for (int e = 0; e < NE; ++e) a = fmaf(rv[jj][e], xK[e], a);
#pragma unroll
      for (int k = 0; k < NC; ++k) a = fmaf(rv[jj][NE + k], c[k], a);
      h[jj] = (j0 + jj < sh.j_hi) ? act_f<ACT>(a) : 0.0f;
    }
#pragma unroll
    for (int jj = 0; jj < FU; ++jj)
#pragma unroll
      for (int e = 0; e < NE; ++e) tp[e] = fmaf(rv[jj][NE + NC + 1 + e], h[jj], tp[e]);
  }
#pragma unroll
  for (int e = 0; e < NE; ++e) xch[(sh.wid * NE + e) * FIT_ROWS + sh.lane] = tp[e];
  __syncthreads();
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const float t = xch[(0 * NE + e) * FIT_ROWS + sh.lane] + xch[(2 * NE + e) * FIT_ROWS + sh.lane];
    const float sv = xch[(1 * NE + e) * FIT_ROWS + sh.lane] + xch[(3 * NE + e) * FIT_ROWS + sh.lane];
    st_x[e] = xT[e];
    st_s[e] = sv;
    xT[e] = fmaf(xT[e], expf(sv), t);
    ld += sv;
  }
}

// backward of one coupling layer; gl: this layer's slice of the shared-memory gradient accumulator
template <int NE, int NC, int ACT>
__device__ __forceinline__ void fit_bwd_layer(const float* __restrict__ wl, float* __restrict__ gl, int H, int rec, const FitShare& sh,
                                              float* __restrict__ xch, float (&xT)[NE], const float (&xK)[NE], float (&gT)[NE],
                                              float (&gK)[NE], const float (&c)[NC > 0 ? NC : 1], const float* __restrict__ st_x,
                                              const float* __restrict__ st_s, float gld) {
  const int net_floats = H * rec + ((NE + 3) & ~3);
  const int lane = sh.lane;
  float d2[NE], du[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const float es = expf(st_s[e]);
    d2[e] = sh.net == 0 ? gT[e]                           // dL/dt
                        : fmaf(gT[e] * st_x[e], es, gld); // dL/ds = g_y * x * exp(s) + g_logdet
    gT[e] *= es;                                          // dL/dx_T
    xT[e] = st_x[e];                                      // input of this layer
    du[e] = 0.0f;
  }
  // Weight-gradient contributions of GJ hidden units at a time: the warp holds 32 per-row values V[jj * REC + q]; a
  // halving butterfly (16 + 8 + 4 + 2 + 1 = 31 shuffles) leaves the row-sum of value l in lane l -- 5x fewer shuffles than
  // one 5-step reduction per value, and the GJ units' recomputed activations overlap instead of running back to back.
  constexpr int REC = 2 * NE + NC + 1;
  constexpr int GJ = 32 / REC;
  const int my_jj = lane / REC, my_q = lane - my_jj * REC;
  const float* w = wl + sh.net * net_floats;
  float* gw = gl + sh.net * net_floats;
  for (int j0 = sh.j_lo; j0 < sh.j_hi; j0 += GJ) {
    float V[32];
#pragma unroll
    for (int i = GJ * REC; i < 32; ++i) V[i] = 0.0f;
#pragma unroll
    for (int jj = 0; jj < GJ; ++jj) {
      const int j = j0 + jj;
      const bool live = j < sh.j_hi;
      float rv[REC];
      load_record<NE, NC>(w + (live ? j : sh.j_lo) * rec, rv);
      float a = rv[NE + NC];
#pragma unroll
      for (int e = 0; e < NE; ++e) a = fmaf(rv[e], xK[e], a);
#pragma unroll
      for (int k = 0; k < NC; ++k) a = fmaf(rv[NE + k], c[k], a);
      const float h = act_f<ACT>(a);
      float dh = 0.0f;
#pragma unroll
      for (int e = 0; e < NE; ++e) dh = fmaf(d2[e], rv[NE + NC + 1 + e], dh);
      float d1 = dh * (ACT == 1 ? fmaf(-h, h, 1.0f) : (h > 0.0f ? 1.0f : 0.0f));
      d1 = live ? d1 : 0.0f;
      const float hl = live ? h : 0.0f;
#pragma unroll
      for (int e = 0; e < NE; ++e) du[e] = fmaf(d1, rv[e], du[e]);
      // this row's contribution to the gradient of record j: [w1x | w1c | b1 | w2]
#pragma unroll
      for (int e = 0; e < NE; ++e) V[jj * REC + e] = d1 * xK[e];
#pragma unroll
      for (int k = 0; k < NC; ++k) V[jj * REC + NE + k] = d1 * c[k];
      V[jj * REC + NE + NC] = d1;
#pragma unroll
      for (int e = 0; e < NE; ++e) V[jj * REC + NE + NC + 1 + e] = d2[e] * hl;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int i = 0; i < m; ++i) {
        const float keep = up ? V[i + m] : V[i];
        const float send = up ? V[i] : V[i + m];
        V[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
      }
    }
    if (my_jj < GJ && j0 + my_jj < sh.j_hi) gw[(j0 + my_jj) * rec + my_q] += V[0];     // lane l owns value l of the group
  }
  if (sh.first_half) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const float v = warp_sum(d2[e]);
      if (lane == e) gw[H * rec + e] += v;                // b2
    }
  }
#pragma unroll
  for (int e = 0; e < NE; ++e) xch[(sh.wid * NE + e) * FIT_ROWS + lane] = du[e];
  __syncthreads();
#pragma unroll
  for (int e = 0; e < NE; ++e)
    gK[e] += (xch[(0 * NE + e) * FIT_ROWS + lane] + xch[(1 * NE + e) * FIT_ROWS + lane]) +
             (xch[(2 * NE + e) * FIT_ROWS + lane] + xch[(3 * NE + e) * FIT_ROWS + lane]);
}

template <int NE, int NC, int ACT>
__global__ void __launch_bounds__(FIT_THREADS) rnvp_small_fit_kernel(const RnvpSmallArgs a) {
  extern __shared__ __align__(16) float wsm[];
  float* gsm = wsm + a.small_floats;
  int* tsm = reinterpret_cast<int*>(gsm + a.small_floats);     // small layout -> packed-gradient index, staged with the weights:
                                                               // the flush at the end must not pay a global-load latency per entry
  float* xchg = reinterpret_cast<float*>(tsm + a.small_floats);                        // [2][FIT_WARPS][NE][FIT_ROWS]
  const int D = a.D, Cd = a.Cd, H = a.H, rec = a.rec, L = a.l1;
  for (int i = threadIdx.x * 4; i < a.small_floats; i += FIT_THREADS * 4) {
    *reinterpret_cast<float4*>(wsm + i) = *reinterpret_cast<const float4*>(a.packed_small + i);
    *reinterpret_cast<float4*>(gsm + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<int4*>(tsm + i) = *reinterpret_cast<const int4*>(a.s2g + i);
  }
  __syncthreads();
  const int layer_floats = 2 * (H * rec + ((NE + 3) & ~3));
  FitShare sh;
  sh.lane = threadIdx.x & 31;
  sh.wid = threadIdx.x >> 5;
  sh.net = sh.wid & 1;
  sh.first_half = (sh.wid >> 1) == 0;
  const int per = (H + 1) / 2;
  sh.j_lo = (sh.wid >> 1) * per;
  sh.j_hi = min(H, sh.j_lo + per);
  constexpr int XCH = FIT_WARPS * NE * FIT_ROWS;
  float loss_part = 0.0f;

  // every thread of the CTA runs the same number of iterations (shuffles and barriers need everyone); rows >= N carry zeros
  for (long long base = (long long)blockIdx.x * FIT_ROWS; base < a.N; base += (long long)gridDim.x * FIT_ROWS) {
    const long long row = base + sh.lane;
    const bool ok = row < a.N;
    const long long src = ok ? (a.idx ? a.idx[row] : row) : 0;
    float xe[NE], xo[NE], c[NC > 0 ? NC : 1], ld = 0.0f;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      xe[e] = (ok && 2 * e < D) ? __ldg(a.X + src * D + 2 * e) : 0.0f;
      xo[e] = (ok && 2 * e + 1 < D) ? __ldg(a.X + src * D + 2 * e + 1) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < (NC > 0 ? NC : 1); ++k) c[k] = (NC > 0 && ok && k < Cd) ? __ldg(a.C + src * Cd + k) : 0.0f;

    float st_x[FIT_MAXL][NE], st_s[FIT_MAXL][NE];        // per-layer stash (local memory, L1-resident)
    int call = 0;                                         // exchange-buffer parity: consecutive layer calls alternate (2L calls per
                                                          // row block, so the next block starts on the other buffer again)
    for (int i = 0; i < L; ++i, ++call) {
      const float* wl = wsm + i * layer_floats;
      float* xch = xchg + (call & 1) * XCH;
      if ((i & 1) == 0) fit_fwd_layer<NE, NC, ACT>(wl, H, rec, sh, xch, xe, xo, c, st_x[i], st_s[i], ld);
      else fit_fwd_layer<NE, NC, ACT>(wl, H, rec, sh, xch, xo, xe, c, st_x[i], st_s[i], ld);
    }
    float q = 0.0f;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      q = fmaf(xe[e], xe[e], q);
      q = fmaf(xo[e], xo[e], q);
    }
    const float lp = ld - 0.5f * (D * 1.8378770664093453f + q);
    if (ok && sh.wid == 0) {
      if (a.out_logp) a.out_logp[row] = lp;
      loss_part += lp;
    }
    // backward: g_z = -scale * z, g_logdet = scale
    const float gld = ok ? a.scale : 0.0f;
    float ge[NE], go[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) { ge[e] = -gld * xe[e]; go[e] = -gld * xo[e]; }
    for (int i = L - 1; i >= 0; --i, ++call) {
      const float* wl = wsm + i * layer_floats;
      float* gl = gsm + i * layer_floats;
      float* xch = xchg + (call & 1) * XCH;
      if ((i & 1) == 0) fit_bwd_layer<NE, NC, ACT>(wl, gl, H, rec, sh, xch, xe, xo, ge, go, c, st_x[i], st_s[i], gld);
      else fit_bwd_layer<NE, NC, ACT>(wl, gl, H, rec, sh, xch, xo, xe, go, ge, c, st_x[i], st_s[i], gld);
    }
  }
  if (a.fuse_adam) {
    // single CTA: gsm holds the step's whole gradient (exactly what the flush below would have added to a zero accumulator)
    if (sh.wid == 0 && a.ad.loss_dst) {
      const float v = warp_sum(loss_part);
      if (sh.lane == 0) *a.ad.loss_dst = v * a.ad.loss_scale;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.ad.n; i += FIT_THREADS) {
      const int p = a.ad.f2p[i], p2 = a.ad.f2p2[i];
      const float g = p2 >= 0 ? gsm[p2 - a.ad.small_off] : 0.0f;
      float mi = a.ad.m[i], vi = a.ad.v[i];
      const float th = rnvp_adam_update(g, a.ad.theta[i], mi, vi, a.ad.k);
      a.ad.m[i] = mi; a.ad.v[i] = vi; a.ad.theta[i] = th;
      if (p >= 0) a.ad.packed[p] = th;
      if (p2 >= 0) a.ad.packed[p2] = th;
    }
    return;
  }
  if (a.loss_sum && sh.wid == 0) {
    const float v = warp_sum(loss_part);
    if (sh.lane == 0 && v != 0.0f) atomicAdd(a.loss_sum, v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.small_floats; i += FIT_THREADS) {
    const int t = tsm[i];
    const float g = gsm[i];
    if (t >= 0 && g != 0.0f) atomicAdd(a.gpacked + t, g);
  }
}

template <int NE, int NC, int ACT>
cudaError_t launch_mode(int mode, const RnvpSmallArgs& a, int grid, size_t smem, cudaStream_t st) {
  if (mode == 2) {
    auto k = rnvp_small_fit_kernel<NE, NC, ACT>;
    const size_t fit_smem = 3 * smem + 2 * FIT_WARPS * NE * FIT_ROWS * sizeof(float);
    if (fit_smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fit_smem);
    k<<<grid, FIT_THREADS, fit_smem, st>>>(a);
    return cudaGetLastError();
  }
  if (mode == 0) {
    auto k = rnvp_small_kernel<NE, NC, ACT, 0>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, THREADS, smem, st>>>(a);
  } else {
    auto k = rnvp_small_kernel<NE, NC, ACT, 1>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, THREADS, smem, st>>>(a);
  }
  return cudaGetLastError();
}
template <int NE, int NC>
cudaError_t launch_act(int act, int mode, const RnvpSmallArgs& a, int grid, size_t smem, cudaStream_t st) {
  return act == 1 ? launch_mode<NE, NC, 1>(mode, a, grid, smem, st) : launch_mode<NE, NC, 2>(mode, a, grid, smem, st);
}
template <int NE>
cudaError_t launch_nc(int NC, int act, int mode, const RnvpSmallArgs& a, int grid, size_t smem, cudaStream_t st) {
  switch (NC) {
    case 0: return launch_act<NE, 0>(act, mode, a, grid, smem, st);
    case 1: return launch_act<NE, 1>(act, mode, a, grid, smem, st);
    case 2: return launch_act<NE, 2>(act, mode, a, grid, smem, st);
    case 4: return launch_act<NE, 4>(act, mode, a, grid, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

int rnvp_small_rows_per_block() { return THREADS * RPT; }
int rnvp_small_fit_rows_per_block() { return FIT_ROWS; }
int rnvp_small_fit_max_layers() { return FIT_MAXL; }

cudaError_t rnvp_launch_small(int NE, int NC, int act, int mode, const RnvpSmallArgs& a, int grid, size_t smem,
                              cudaStream_t st) {
  switch (NE) {
    case 1: return launch_nc<1>(NC, act, mode, a, grid, smem, st);
    case 2: return launch_nc<2>(NC, act, mode, a, grid, smem, st);
    case 4: return launch_nc<4>(NC, act, mode, a, grid, smem, st);
    default: return cudaErrorInvalidValue;
  }
}
