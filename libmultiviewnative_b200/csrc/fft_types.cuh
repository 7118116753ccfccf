// fft_types.cuh -- plain structs shared by the host engine and the kernels.
#pragma once
#include "lmvn_common.cuh"

namespace lmvn {

struct UpdateParams {
  float min_value;
  float two_lambda;   // float(2*lambda)
  float coef;         // float(2*lambda*double(float(1/lambda)))  (== 2 up to the rounding of 1/lambda)
  int regularized;    // lambda > 0
};

static inline UpdateParams make_update_params(double lambda, float min_value) {
  UpdateParams p;
  p.min_value = min_value;
  p.regularized = lambda > 0.0 ? 1 : 0;
  if (p.regularized) {
    // ref: inc/cpu_kernels.h:71 `TransferT lambda_inv = 1.f / _lambda;` -- the division
    // happens in double (_lambda is double) and is then rounded to float
    const float lambda_inv = float(1.0 / lambda);
    p.two_lambda = float(2.0 * lambda);
    p.coef = float(2.0 * lambda * double(lambda_inv));
  } else {
    // plain RL through the same arithmetic: 2 v / (1 + sqrt(1 + 0 v)) == v (see rl_update)
    p.two_lambda = 0.f;
    p.coef = 2.f;
  }
  return p;
}

// Destination of a scattered final store (slab-decomposed transform over several GPUs): output
// row r of the transform axis goes to device r >> shift, whose buffer may be peer memory reached
// over NVLink.  address = base[r >> shift] + offset + (r & mask) * row_stride + slow * tile_stride + col.
struct Scatter {
  float2* base[8];
  int shift;
  long long row_stride, tile_stride, offset;
};

namespace gen {

struct AxisPlan {
  int n;
  int nf;
  int factors[16];
  const cplx* tw;  // device table, n entries: exp(-2 pi i k / n)
};

struct RealSource {
  const float* data;
  int wrapped;     // 0: data is the [nz][ny][nx] volume; 1: data is a small kernel to be wrapped
  int kz, ky, kx;  // kernel extents when wrapped
};

enum EpilogueMode { EPI_STORE = 0, EPI_QUOTIENT = 1, EPI_UPDATE = 2 };

struct Epilogue {
  int mode;
  float scale;           // applied to the inverse-transform output before the pointwise step
  const float* view;     // EPI_QUOTIENT: out = view / value
  float* psi;            // EPI_UPDATE: psi = update(psi, value, weights) (out is ignored)
  const float* weights;
  UpdateParams up;
  // EPI_QUOTIENT on zero-padded stacks: a zero of the view gives a zero quotient whatever the blurred value is.
  // The padding of the view is exactly zero and the blurred estimate there is round-off noise that may be exactly
  // zero too (extents rounded up to a fast-path size leave padding no kernel tap reaches); 0 * (1 / 0) would be a
  // NaN that the second convolution spreads over the whole stack.
  int zero_view_guard;
};

// Source index of target position t for a kernel of extent k wrapped into n
// (centre k/2 at the origin), or -1 where the padded volume is zero.
__host__ __device__ __forceinline__ int wrap_src_index(int t, int n, int k) {
  const int kc = k / 2;
  if (t < k - kc) return t + kc;
  if (t >= n - kc) return t - (n - kc);
  return -1;
}

}  // namespace gen
}  // namespace lmvn
