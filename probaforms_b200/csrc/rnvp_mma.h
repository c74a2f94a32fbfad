// Arguments of the tcgen05 forward / inverse kernel (rnvp_mma.cu).
#pragma once
#include <stdint.h>

struct RnvpMmaArgs {
  const float* wimg;           // per layer [W1 image | W2 image | b2], see rnvp_planner.h build_mma_map
  const float* X;              // rows (forward) or latent noise (inverse), [N][D]
  const float* C;              // [N][Cd] or nullptr
  const long long* idx;        // optional row gather
  long long N;
  float* out_x;                // z / x
  float* out_logdet;
  float* out_logp;
  int Cd, H, l0, l1;
  int layer_floats, w1_floats, w2_floats;
  int n_pairs;                 // pairs of 128-row tiles
  // MODE 2 (forward pass of a fit step): per layer the row's x_T (before the coupling) and s go to a global stash,
  // [row][layer][x_T(DH) | s(DH)], read back by the backward-only tile program; sum of logp accumulates atomically
  float* stash;
  float* loss_sum;
  int L_total;                 // layers of the flow (stash row = L_total * 2 * DH floats)
  // MODE 2 with do_bwd: the kernel continues with the backward sweep (everything but the weight gradients): per (layer,
  // row) it leaves a record [h 2H | u K1P8 | delta2 2*DH] in the blocked records array (rnvp_wgrad.cu) for rnvp_wgrad_kernel
  int do_bwd;
  float scale;                 // d(out)/d(logp_row): g_logdet = scale, g_z = -scale*z
  float* records;
  int rec;
  int rec_swz;                 // slot swizzle mask of the record blocks: 7 (tcgen05 weight-gradient sweep) or 1 (mma.sync sweep)
  long long Npad;
  long long* trace;            // development aid (rnvp_debug_set_trace): CTA 0 logs (tag, clock64) pairs of the backward sweep
  int wt_floats;               // floats of one transposed image (W2T, then W1T) per layer; 0 unless do_bwd
  // MODE 1 with X == nullptr: latent rows drawn in-kernel (rnvp_philox.cuh), keyed on row_offset + row
  unsigned long long seed;
  long long row_offset;
  int Dreal;                   // features per row in memory (rnvp_wide.cu: < 2 * DH for padded shapes, the row loads are guarded)
};
