// pointwise.cuh -- the three per-voxel steps of the multi-view RL iteration.
//
// Arithmetic spec = the reference CPU kernels (ref: inc/cpu_kernels.h):
//   quotient        :19-26   out = in * float(1. / out)
//   final_values    :28-54   v = psi*integral; !(v>0) -> min; nan/inf -> min;
//                            max(v,min); psi = w*(v-psi)+psi
//   regularized_... :59-90   for v>0: v = float(1.f/lambda) * (sqrt(1+2*lambda*v)-1)  [in double]
// These are device functions so that the FFT passes can fuse them into their
// epilogues; the standalone kernels below only serve the legacy entry points
// compute_quotient / compute_final_values (ref: src/multiviewnative.cu:321-393).
#pragma once
#include "fft_types.cuh"
#include "lmvn_common.cuh"

namespace lmvn {

// view / blurred.  __frcp_rn is the correctly rounded float reciprocal, i.e.
// float(1.0 / double(x)) up to (rare) double rounding; no fast-math.
__device__ __forceinline__ float quotient(float view, float blurred) {
  return __fmul_rn(view, __frcp_rn(blurred));
}

// Tikhonov step in the cancellation-free form
//   (1/l)(sqrt(1+2 l v) - 1) == 2 v / (1 + sqrt(1 + 2 l v)),
// which float32 evaluates to ~1e-7 of the reference's double evaluation (the
// textbook form loses up to 6.6e-3 relative for small v in float32).
__device__ __forceinline__ float tikhonov(float v, const UpdateParams& p) {
  float s = __fsqrt_rn(__fmaf_rn(p.two_lambda, v, 1.0f));
  return __fdiv_rn(__fmul_rn(p.coef, v), __fadd_rn(1.0f, s));
}

__device__ __forceinline__ float rl_update(float psi, float integral, float weight,
                                           const UpdateParams& p) {
  const float last = psi;
  float v = __fmul_rn(last, integral);
  if (v > 0.f) {
    if (p.regularized) v = tikhonov(v, p);
  } else {
    v = p.min_value;  // also catches NaN
  }
  float next = (isnan(v) || isinf(v)) ? p.min_value : fmaxf(v, p.min_value);
  // two roundings like the CPU code, no FMA contraction
  return __fadd_rn(__fmul_rn(weight, __fadd_rn(next, -last)), last);
}

// ---- standalone kernels (legacy API only) -------------------------------------
static __global__ void k_divide(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t n4 = n / 4;
  const float4* in4 = reinterpret_cast<const float4*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (size_t k = i; k < n4; k += stride) {
    float4 a = in4[k], b = out4[k];
    b.x = quotient(a.x, b.x); b.y = quotient(a.y, b.y);
    b.z = quotient(a.z, b.z); b.w = quotient(a.w, b.w);
    out4[k] = b;
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) out[k] = quotient(in[k], out[k]);
}

static __global__ void k_final_values(float* __restrict__ psi, const float* __restrict__ integral,
                               const float* __restrict__ weight, size_t n, UpdateParams p) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t n4 = n / 4;
  float4* psi4 = reinterpret_cast<float4*>(psi);
  const float4* int4p = reinterpret_cast<const float4*>(integral);
  const float4* w4 = reinterpret_cast<const float4*>(weight);
  for (size_t k = i; k < n4; k += stride) {
    float4 a = psi4[k], b = int4p[k], w = w4[k];
    a.x = rl_update(a.x, b.x, w.x, p); a.y = rl_update(a.y, b.y, w.y, p);
    a.z = rl_update(a.z, b.z, w.z, p); a.w = rl_update(a.w, b.w, w.w, p);
    psi4[k] = a;
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) psi[k] = rl_update(psi[k], integral[k], weight[k], p);
}

}  // namespace lmvn
