// Layer-1 tables of the tensor-core prefilter (tc_pass1.cu): the first linear layer of the distance network is
// separable in its input [q, p] (network_macros_mod.py:139-141),
//   A_i[k] = b1[k] + sum_{c < d} W1[k][c] q_c + W1[k][nin + c] sin q_c + W1[k][2 nin + c] cos q_c          (n, 256)
//   B_j[k] =         sum_{c < P} W1[k][d + c] p_c + W1[k][nin + d + c] sin p_c + W1[k][2 nin + d + c] cos p_c
// fp32 FMAs over the fp32 weights in ONE fixed order, rounded once to fp16 / bf16 -- the same function is used by the
// table kernel and by the fused rank + step kernel (rollout_kernels.cu), so a table does not depend on who wrote it.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.cuh"

__device__ __forceinline__ uint16_t l1_to_half_bits(float v, bool bf16) {
  // clamped to half the fp16 range so that A_i + B_j can never overflow to infinity in the packed add (everything
  // downstream converts with .satfinite); the shipped networks stay below 1e2 here
  v = fminf(fmaxf(v, -30000.f), 30000.f);
  if (bf16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v));
}

// feature k of one table row; xs = [x_0, sin x_0, cos x_0, x_1, ...] of the ncomp input components starting at
// input column comp0; Wf0 = W1^T, [3 nin][256]
__device__ __forceinline__ float l1_feature(const float* __restrict__ Wf0, const float* __restrict__ bias, int k,
                                            const float* xs, int ncomp, int comp0, int nin) {
  float acc = bias ? bias[k] : 0.f;
  for (int c = 0; c < ncomp; ++c) {
    acc = fmaf(Wf0[(size_t)(comp0 + c) * HID + k], xs[3 * c], acc);
    acc = fmaf(Wf0[(size_t)(nin + comp0 + c) * HID + k], xs[3 * c + 1], acc);
    acc = fmaf(Wf0[(size_t)(2 * nin + comp0 + c) * HID + k], xs[3 * c + 2], acc);
  }
  return acc;
}
