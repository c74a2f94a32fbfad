// torch.optim.Adam single-tensor update (torch/optim/adam.py _single_tensor_adam) for one parameter: shared by adam_kernel
// (rnvp_api.cu) and the fused tail of the small-flow fit step (rnvp_small.cu).  Scalar step math (bias corrections) is done
// on the host in double, as torch does.
#pragma once

struct RnvpAdamCoef {
  float wd, one_minus_b1, b2, one_minus_b2, step_size, bc2_sqrt, eps;
};

__device__ __forceinline__ float rnvp_adam_update(float g, float th, float& mi, float& vi, const RnvpAdamCoef& k) {
  if (k.wd != 0.0f) g = fmaf(k.wd, th, g);                 // grad.add(param, alpha=weight_decay)
  mi = fmaf(k.one_minus_b1, g - mi, mi);                   // exp_avg.lerp_(grad, 1-beta1)
  vi = fmaf(k.one_minus_b2 * g, g, vi * k.b2);             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
  const float denom = __fsqrt_rn(vi) / k.bc2_sqrt + k.eps; // (sqrt(v)/sqrt(bc2)).add_(eps)
  return fmaf(-k.step_size, mi / denom, th);               // param.addcdiv_(exp_avg, denom, value=-step_size)
}
