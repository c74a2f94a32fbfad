// Arguments of the row-per-thread small-flow kernels (rnvp_small.cu).
#pragma once
#include <stdint.h>
#include "rnvp_adam.cuh"

// Optional Adam update fused behind a fit step that runs as ONE CTA (rnvp_fit_epoch with batches of <= 32 rows, the
// reference's default batch_size): the whole step's gradient is in the CTA's shared memory, so the optimiser update, the
// refresh of the packed copies and the loss hand-off need neither the packed-gradient round trip nor a second launch.
struct RnvpFusedAdam {
  float* theta;                // flat parameters [n]
  float* packed;               // packed copies (tile layout at f2p, small layout at f2p2)
  float* m;
  float* v;
  const int* f2p;
  const int* f2p2;             // absolute index of the small-layout copy in `packed`, or -1
  int n, small_off;
  RnvpAdamCoef k;
  float* loss_dst;             // receives loss_scale * sum_rows logp (may be nullptr)
  float loss_scale;
};

struct RnvpSmallArgs {
  const float* packed_small;   // per layer, per net: H records [w1_x NE | w1_c NC | b1 | w2 NE] (padded to rec), then b2
  const float* X;              // rows (forward) or latent noise (inverse), [N][D]
  const float* C;              // [N][Cd] or nullptr
  const long long* idx;        // optional row gather
  long long N;
  float* out_x;                // z / x, may be nullptr in forward mode
  float* out_logdet;
  float* out_logp;
  int D, Cd, H, rec, small_floats;
  int l0, l1;
  // fit step (rnvp_small_fit_kernel): d(scale * sum_rows logp)/d(theta) accumulated into gpacked (tile layout) through
  // s2g[small index] = gpacked index or -1; loss_sum += sum_rows logp
  float* gpacked;
  const int* s2g;
  float* loss_sum;
  float scale;
  // inverse mode with X == nullptr: the latent rows are drawn in-kernel (rnvp_philox.cuh), keyed on row_offset + row
  unsigned long long seed;
  long long row_offset;
  int fuse_adam;               // fit step only, single-CTA launches only
  RnvpFusedAdam ad;
};
