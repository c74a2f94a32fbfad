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
};
