/* A C client of the reference's API, as the reference's own benches are written (ref: bench/bench_gpu_deconvolve.cu:40-80,
 * bench/synthetic_data.hpp:58-96): constant views 16 + 4 i, unit weights, delta kernels of value i + 1 / i + 2,
 * psi0 = view 0.  With delta kernels every voxel evolves independently, so the expected result is a scalar recurrence.
 *
 *   gcc -std=c99 -I include examples/deconvolve_c_client.c -L libmultiviewnative_b200/lib -lmultiviewnative -lm \
 *       -Wl,-rpath,$PWD/libmultiviewnative_b200/lib -o deconvolve_c_client && ./deconvolve_c_client
 *
 * Nothing but the library changes with respect to a build against the reference. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "multiviewnative.h"

int main(void) {
  enum { NV = 6, NZ = 64, NY = 32, NX = 64, K = 3 };
  const size_t n = (size_t)NZ * NY * NX;
  int dims[3] = {NZ, NY, NX}, kdims[3] = {K, K, K};
  view_data views[NV];
  for (int v = 0; v < NV; ++v) {
    float* image = malloc(n * sizeof(float));
    float* weights = malloc(n * sizeof(float));
    float* k1 = calloc(K * K * K, sizeof(float));
    float* k2 = calloc(K * K * K, sizeof(float));
    for (size_t i = 0; i < n; ++i) { image[i] = 16.f + 4.f * v; weights[i] = 1.f; }
    k1[K * K * K / 2] = (float)(v + 1);
    k2[K * K * K / 2] = (float)(v + 2);
    views[v].image_ = image; views[v].kernel1_ = k1; views[v].kernel2_ = k2; views[v].weights_ = weights;
    views[v].image_dims_ = dims; views[v].kernel1_dims_ = kdims; views[v].kernel2_dims_ = kdims; views[v].weights_dims_ = dims;
  }
  float* psi = malloc(n * sizeof(float));
  for (size_t i = 0; i < n; ++i) psi[i] = 16.f;
  workspace w;
  w.data_ = views; w.num_views_ = NV; w.lambda_ = 0.006; w.minValue_ = 1e-3f; w.num_iterations_ = 3;

  if (getNumDevicesCUDA() < 1) { printf("no CUDA device: nothing to run (there is no CPU fallback)\n"); return 0; }
  inplace_gpu_deconvolve(psi, w, selectDeviceWithHighestComputeCapability());

  double p = 16.0;
  for (int it = 0; it < w.num_iterations_; ++it)
    for (int v = 0; v < NV; ++v) {
      double val = p * ((16.0 + 4.0 * v) / (p * (v + 1)) * (v + 2));
      val = (sqrt(1.0 + 2.0 * w.lambda_ * val) - 1.0) / w.lambda_;
      p = val > w.minValue_ ? val : w.minValue_;
    }
  double worst = 0.0;
  for (size_t i = 0; i < n; ++i) { double e = fabs(psi[i] - p) / p; if (e > worst) worst = e; }
  printf("expected %.6f, max relative deviation %.3g -> %s\n", p, worst, worst < 2e-5 ? "OK" : "MISMATCH");
  return worst < 2e-5 ? 0 : 1;
}
