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

// view / blurred: view * (1 / blurred) like the reference (ref: inc/cpu_kernels.h:19-26, which rounds a
// double reciprocal to float).
// Host emulation of the device's MUFU arithmetic (tests/emu): flush-to-zero on inputs and results like the .ftz forms
// below, so that the emulated parity cases exercise the SAME special-value behaviour as the device (a subnormal blurred
// value gives an infinite quotient on both; the reference gives a finite 8.5e37 .. 3.4e38 there, an infinity below 2.9e-39);
// the values themselves differ from MUFU's by <= 1-2 ulp.
#ifdef LMVN_EMU
static inline float emu_ftz(float x) { return (std::fabs(x) < 1.17549435e-38f) ? std::copysign(0.0f, x) : x; }
#endif

__device__ __forceinline__ float quotient(float view, float blurred) {
#ifdef LMVN_EMU
  return __fmul_rn(view, emu_ftz(1.0f / emu_ftz(blurred)));
#else
  // MUFU reciprocal (<= 1 ulp, IEEE special values for 0 / Inf / NaN): two instructions per voxel instead of
  // the range-checked IEEE sequence with its slow-path branch
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(blurred));
  return __fmul_rn(view, r);
#endif
}

// zero-padded stacks (Epilogue::zero_view_guard): the quotient of a zero voxel of the view is zero
__device__ __forceinline__ float quotient(float view, float blurred, int zero_view_guard) {
  const float q = quotient(view, blurred);
  return (zero_view_guard && view == 0.f) ? 0.f : q;
}

// Building blocks of the update.  It runs once per voxel inside the last transform pass; with IEEE-rounded
// division and square root (range checks, slow paths: FCHK, BSSY/BSYNC, ~120 instructions per voxel) that
// pass was bound by instruction issue, not by memory.  The MUFU approximations are within 1-2 ulp, three
// orders of magnitude inside the parity tolerance (measured: tests/test_gpu_parity.py prints ~1e-6).
__device__ __forceinline__ float rcp_fast(float d) {
#ifdef LMVN_EMU
  return emu_ftz(1.0f / emu_ftz(d));
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
#endif
}
__device__ __forceinline__ float sqrt_fast(float x) {
#ifdef LMVN_EMU
  return emu_ftz(std::sqrt(emu_ftz(x)));
#else
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}

// Tikhonov step in the cancellation-free form
//   (1/l)(sqrt(1+2 l v) - 1) == 2 v / (1 + sqrt(1 + 2 l v)),
// which float32 evaluates to ~1e-7 of the reference's double evaluation (the
// textbook form loses up to 6.6e-3 relative for small v in float32).
__device__ __forceinline__ float tikhonov(float v, const UpdateParams& p) {
  const float s = sqrt_fast(__fmaf_rn(p.two_lambda, v, 1.0f));
  return __fmul_rn(__fmul_rn(p.coef, v), rcp_fast(__fadd_rn(1.0f, s)));
}

__device__ __forceinline__ float rl_update(float psi, float integral, float weight,
                                           const UpdateParams& p) {
  const float last = psi;
  float v = __fmul_rn(last, integral);
  // selects instead of branches: !(v > 0) also catches NaN (ref: inc/cpu_kernels.h:36-38, 66-79)
  // plain RL (lambda <= 0) runs the same formula with two_lambda = 0, coef = 2: 2 v / (1 + sqrt(1)) == v
  // exactly (sqrt(1) and 1/2 are exact in the MUFU approximations), so there is no per-voxel mode branch
  const float t = tikhonov(v, p);
  v = (v > 0.f) ? t : p.min_value;
  // NaN or Inf -> minValue, else max(v, minValue) (ref: inc/cpu_kernels.h:40-47)
  const float next = (fabsf(v) <= 3.402823466e+38f) ? fmaxf(v, p.min_value) : p.min_value;
  // two roundings like the CPU code, no FMA contraction
  return __fadd_rn(__fmul_rn(weight, __fadd_rn(next, -last)), last);
}

// Periodic embedding: a volume of logical extents (lz, ly, lx) sits at offset (oz, oy, ox) inside a larger
// power-of-two volume (nz, ny, nx).  Every voxel OUTSIDE the logical box is overwritten with the logical voxel it
// aliases under periodic continuation.  A circular convolution at the large extents then equals, inside the box,
// the circular convolution at the logical extents (ref semantics: inc/cpu_convolve.h:24) as long as the margins are
// at least the kernel half-widths -- which lets arbitrary stack sizes run on the power-of-two fast path.
static __global__ void k_wrap_exterior(float* __restrict__ vol, int nz, int ny, int nx, int lz, int ly, int lx, int oz,
                                       int oy, int ox) {
  // one block per row (z, y): the source row is computed once, interior rows only touch their x margins
  const int y = blockIdx.x, z = blockIdx.y;
  int sz = (z - oz) % lz, sy = (y - oy) % ly;
  if (sz < 0) sz += lz;
  if (sy < 0) sy += ly;
  const bool interior_row = (z >= oz && z < oz + lz && y >= oy && y < oy + ly);
  float* dst = vol + (size_t(z) * ny + y) * nx;
  const float* src = vol + (size_t(sz + oz) * ny + (sy + oy)) * nx + ox;
  for (int x = threadIdx.x; x < nx; x += blockDim.x) {
    if (interior_row && x >= ox && x < ox + lx) continue;
    int sx = (x - ox) % lx;
    if (sx < 0) sx += lx;
    dst[x] = src[sx];
  }
  (void)nz;
}

// Padded plans move host stacks with ONE contiguous copy through a staging buffer; these two kernels place the
// stack into its box inside the (zero filled) volume and gather it back.  One block per row (z, y) of the volume.
static __global__ void k_place_box(float* __restrict__ vol, const float* __restrict__ box, int ny, int nx, int lz, int ly,
                                   int lx, int oz, int oy, int ox) {
  const int y = blockIdx.x, z = blockIdx.y;
  const bool row_in = (z >= oz && z < oz + lz && y >= oy && y < oy + ly);
  float* dst = vol + (size_t(z) * ny + y) * nx;
  const float* src = box + (size_t(row_in ? z - oz : 0) * ly + (row_in ? y - oy : 0)) * lx;
  for (int x = threadIdx.x; x < nx; x += blockDim.x) dst[x] = (row_in && x >= ox && x < ox + lx) ? src[x - ox] : 0.f;
}
static __global__ void k_gather_box(const float* __restrict__ vol, float* __restrict__ box, int ny, int nx, int ly, int lx,
                                    int oz, int oy, int ox) {
  const int y = blockIdx.x, z = blockIdx.y;  // box coordinates
  const float* src = vol + (size_t(z + oz) * ny + (y + oy)) * nx + ox;
  float* dst = box + (size_t(z) * ly + y) * lx;
  for (int x = threadIdx.x; x < lx; x += blockDim.x) dst[x] = src[x];
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
