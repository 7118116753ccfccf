/* A torch-free walk over every geometry of the C ABI at small sizes: one-shot calls (native fast path with every rows and
 * strided plan, periodic embedding, zero_padd, generic), persistent plans, the slab-decomposed plan with in-process
 * ranks, the legacy entry points.  Delta kernels make every voxel follow its own scalar recurrence (as in
 * examples/deconvolve_c_client.c) while views, weights and the start value differ from voxel to voxel (a hash of the
 * voxel index), so each case checks itself and a voxel that lands in the wrong place shows up.
 *
 * Two uses (tests/test_abi_walk.py):
 *  - on a GPU, against the product library:
 *      gcc -std=c99 -O1 -I include tools/abi_walk.c -L libmultiviewnative_b200/lib -lmultiviewnative -lm \
 *          -Wl,-rpath,$PWD/libmultiviewnative_b200/lib -o abi_walk && ./abi_walk
 *  - on the CPU, as the bounds check of the kernels' index math: against the host-emulation build of the same .cu
 *    sources under AddressSanitizer with red zones between the sub-buffers of the device arenas
 *    (libmultiviewnative_b200._build.build_emu(asan=True)); compute-sanitizer is not available on the GPU pool.
 * argv[1..] (optional): substrings; only the cases whose name contains one of them run. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lmvn_b200.h"
#include "multiviewnative.h"

static char** g_filters = NULL; /* argv[1..]: run the cases whose name contains one of them (none: all) */
static int g_num_filters = 0;
static int g_failed = 0;

static int skipped(const char* name) {
  if (g_num_filters == 0) return strstr(name, "huge") != NULL; /* 12 GiB of host stacks: only on request */
  for (int i = 0; i < g_num_filters; ++i)
    if (strstr(name, g_filters[i])) return 0;
  return 1;
}

/* per-voxel pattern in [0, 1): a hash of the voxel index in the whole volume */
static float pattern(size_t i, unsigned salt) {
  uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull + (uint64_t)salt * 0xD1B54A32D192ED03ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (float)(h & 1023) / 1024.f;
}
static float view_value(int v, size_t i) { return (16.f + 4.f * v) * (1.f + 0.25f * pattern(i, 11u + v)); }
static float weight_value(int v, size_t i) { return 0.25f + 0.5f * pattern(i, 101u + v); }
static float psi0_value(size_t i) { return 16.f * (1.f + 0.25f * pattern(i, 7u)); }

/* what voxel i holds after `iters` sweeps when kernel1 of view v is (v + 1) * delta and kernel2 is (v + 2) * delta;
 * unit_weights: the slab and legacy cases that set weights = 1 */
static double recurrence(size_t i, int nv, int iters, double lambda, double min_value) {
  double p = psi0_value(i);
  for (int it = 0; it < iters; ++it)
    for (int v = 0; v < nv; ++v) {
      double val = p * (view_value(v, i) / (p * (v + 1)) * (v + 2));
      if (lambda > 0) val = (sqrt(1.0 + 2.0 * lambda * val) - 1.0) / lambda;
      val = val > min_value ? val : min_value;
      p = weight_value(v, i) * (val - p) + p;
    }
  return p;
}

static void verdict(const char* name, double worst, double tol) {
  const int ok = worst < tol;
  printf("%-34s geometry %d  max rel dev %.3g  %s\n", name, lmvn_last_geometry(), worst, ok ? "OK" : "MISMATCH");
  if (!ok) { ++g_failed; printf("    last error: %s\n", lmvn_last_error()); }
  fflush(stdout);
}

/* psi against the recurrence; `first` = index of psi[0] in the whole volume (slabs) */
static void report_recurrence(const char* name, const float* psi, size_t n, size_t first, int nv, int iters,
                              double lambda, double min_value) {
  double worst = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double expect = recurrence(first + i, nv, iters, lambda, min_value);
    const double e = fabs(psi[i] - expect) / expect;
    if (!(e <= worst)) worst = e;
  }
  verdict(name, worst, 5e-5);
}

/* got[i] against scale * ref[i] (ref == NULL: against the constant `scale`) */
static void report(const char* name, const float* got, const float* ref, size_t n, double scale, double tol) {
  double worst = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double expect = ref ? scale * ref[i] : scale;
    const double e = fabs(got[i] - expect) / fabs(expect);
    if (!(e <= worst)) worst = e;
  }
  verdict(name, worst, tol);
}

/* zero_padd geometry: a delta kernel leaves exact zeros in the padding, 0 / 0 there is NaN (in the reference as well,
 * ref: src/gpu_deconvolve_methods.cuh:471-476) and the second convolution spreads it.  Those cases use box kernels,
 * which reach all of the padding, and only check that the result is finite and above the floor. */
static void fill_kernel(float* k, int kk, int centre, float value, int box) {
  for (int i = 0; i < kk; ++i) k[i] = box ? value / kk : 0.f;
  if (!box) k[centre] = value;
}

static void report_finite(const char* name, const float* psi, size_t n, float floor_value) {
  size_t bad = 0;
  for (size_t i = 0; i < n; ++i) bad += !(psi[i] > 2.f * floor_value && psi[i] < 1e6f);
  printf("%-34s geometry %d  %zu voxel(s) not finite or at the floor  %s\n", name, lmvn_last_geometry(), bad,
         bad ? "MISMATCH" : "OK");
  if (bad) { ++g_failed; printf("    last error: %s\n", lmvn_last_error()); }
  fflush(stdout);
}

/* one-shot deconvolution with delta kernels of extent k at the kernel centre (box != 0: uniform kernels) */
static void deconv_case(const char* name, int nz, int ny, int nx, int k, int nv, int iters, double lambda, int box) {
  if (skipped(name)) return;
  const size_t n = (size_t)nz * ny * nx;
  const int kk = k * k * k;
  int dims[3] = {nz, ny, nx}, kdims[3] = {k, k, k};
  view_data* views = calloc(nv, sizeof(view_data));
  for (int v = 0; v < nv; ++v) {
    float* image = malloc(n * sizeof(float));
    float* weights = malloc(n * sizeof(float));
    float* k1 = calloc(kk, sizeof(float));
    float* k2 = calloc(kk, sizeof(float));
    for (size_t i = 0; i < n; ++i) { image[i] = view_value(v, i); weights[i] = weight_value(v, i); }
    /* centre element (k/2, k/2, k/2) */
    const int c = ((k / 2) * k + (k / 2)) * k + (k / 2);
    fill_kernel(k1, kk, c, (float)(v + 1), box);
    fill_kernel(k2, kk, c, (float)(v + 2), box);
    views[v].image_ = image; views[v].kernel1_ = k1; views[v].kernel2_ = k2; views[v].weights_ = weights;
    views[v].image_dims_ = dims; views[v].kernel1_dims_ = kdims; views[v].kernel2_dims_ = kdims;
    views[v].weights_dims_ = dims;
  }
  float* psi = malloc(n * sizeof(float));
  for (size_t i = 0; i < n; ++i) psi[i] = psi0_value(i);
  workspace w;
  w.data_ = views; w.num_views_ = (unsigned short)nv; w.lambda_ = lambda; w.minValue_ = 1e-3f; w.num_iterations_ = iters;
  lmvn_clear_error();
  inplace_gpu_deconvolve(psi, w, 0);
  if (box) report_finite(name, psi, n, 1e-3f);
  else report_recurrence(name, psi, n, 0, nv, iters, lambda, 1e-3);
  for (int v = 0; v < nv; ++v) { free(views[v].image_); free(views[v].weights_); free(views[v].kernel1_); free(views[v].kernel2_); }
  free(views); free(psi);
}

/* image (x) delta kernel of value 3 */
static void conv_case(const char* name, int nz, int ny, int nx, int k) {
  if (skipped(name)) return;
  const size_t n = (size_t)nz * ny * nx;
  int dims[3] = {nz, ny, nx}, kdims[3] = {k, k, k};
  float* im = malloc(n * sizeof(float));
  float* ker = calloc((size_t)k * k * k, sizeof(float));
  float* ref = malloc(n * sizeof(float));
  for (size_t i = 0; i < n; ++i) ref[i] = im[i] = psi0_value(i);
  ker[((k / 2) * k + (k / 2)) * k + (k / 2)] = 3.f;
  lmvn_clear_error();
  inplace_gpu_convolution(im, dims, ker, kdims, 0);
  report(name, im, ref, n, 3.0, 5e-5);
  free(im); free(ker); free(ref);
}

/* persistent plan: set views, iterate in two calls, convolve, profile */
static void plan_case(const char* name, int nz, int ny, int nx, int k, int kind) {
  if (skipped(name)) return;
  enum { NV = 2 };
  const size_t n = (size_t)nz * ny * nx;
  int dims[3] = {nz, ny, nx}, kdims[3] = {k, k, k};
  lmvn_plan* p = NULL;
  lmvn_clear_error();
  int rc = kind == 0   ? lmvn_plan_create(&p, dims, NV, 0)
           : kind == 1 ? lmvn_plan_create_embedded(&p, dims, kdims, NV, 0)
                       : lmvn_plan_create_zero_padded(&p, dims, kdims, NV, 0);
  if (rc != 0) { printf("%-34s plan create failed: %s\n", name, lmvn_last_error()); ++g_failed; return; }
  float* image = malloc(n * sizeof(float));
  float* weights = malloc(n * sizeof(float));
  float* psi = malloc(n * sizeof(float));
  float* k1 = calloc((size_t)k * k * k, sizeof(float));
  float* k2 = calloc((size_t)k * k * k, sizeof(float));
  const int c = ((k / 2) * k + (k / 2)) * k + (k / 2);
  for (int v = 0; v < NV; ++v) {
    for (size_t i = 0; i < n; ++i) { image[i] = view_value(v, i); weights[i] = weight_value(v, i); }
    fill_kernel(k1, k * k * k, c, (float)(v + 1), kind == 2);
    fill_kernel(k2, k * k * k, c, (float)(v + 2), kind == 2);
    rc |= lmvn_plan_set_view(p, v, image, weights, k1, kdims, k2, kdims);
  }
  for (size_t i = 0; i < n; ++i) psi[i] = psi0_value(i);
  rc |= lmvn_plan_set_psi(p, psi);
  rc |= lmvn_plan_iterate(p, 2, 0.006, 1e-3f, NULL);
  rc |= lmvn_plan_iterate(p, 1, 0.006, 1e-3f, NULL);
  {
    char names[64 * 48]; float ms[64]; unsigned long long bytes[64]; int count = 0;
    rc |= lmvn_plan_profile(p, 0.006, 1e-3f, 64, names, ms, bytes, &count);
  }
  rc |= lmvn_plan_get_psi(p, psi);
  if (rc != 0) { printf("%-34s failed: %s\n", name, lmvn_last_error()); ++g_failed; }
  if (kind == 2) report_finite(name, psi, n, 1e-3f);
  else report_recurrence(name, psi, n, 0, NV, 3, 0.006, 1e-3);
  /* psi <- psi (x) kernel1 of view 1 (delta of value 2) */
  float* before = malloc(n * sizeof(float));
  memcpy(before, psi, n * sizeof(float));
  rc = lmvn_plan_convolve(p, 1, 1, 1, NULL);
  rc |= lmvn_plan_get_psi(p, psi);
  if (rc != 0) { printf("%-34s convolve failed: %s\n", name, lmvn_last_error()); ++g_failed; }
  else if (kind == 2) {
    char nm[64];
    snprintf(nm, sizeof nm, "%s/convolve", name);
    report_finite(nm, psi, n, 1e-3f);
  } else {
    char nm[64];
    snprintf(nm, sizeof nm, "%s/convolve", name);
    report(nm, psi, before, n, 2.0, 5e-5);
  }
  lmvn_plan_destroy(p);
  free(image); free(weights); free(psi); free(k1); free(k2); free(before);
}

/* ONE volume over `world` in-process ranks (all on device 0): slabs of nz/world planes, phases issued for every rank
 * with a synchronisation in between, as libmultiviewnative_b200/slabs.py LocalSlabGroup does */
static void dist_case(const char* name, int nz, int ny, int nx, int world) {
  if (skipped(name)) return;
  enum { NV = 2, K = 3, MAXW = 8 };
  int dims[3] = {nz, ny, nx}, kdims[3] = {K, K, K};
  const size_t slab = (size_t)(nz / world) * ny * nx;
  lmvn_dist* r[MAXW] = {0};
  int rc = 0;
  lmvn_clear_error();
  for (int i = 0; i < world; ++i) rc |= lmvn_dist_create(&r[i], dims, NV, i, world, 0);
  if (rc != 0) { printf("%-34s create failed: %s\n", name, lmvn_last_error()); ++g_failed; goto done; }
  for (int i = 0; i < world; ++i)
    for (int j = 0; j < world; ++j)
      if (i != j) rc |= lmvn_dist_connect_local(r[i], j, r[j]);
  float* image = malloc(slab * sizeof(float));
  float* weights = malloc(slab * sizeof(float));
  float* psi = malloc(slab * sizeof(float));
  float k[K * K * K];
  for (int v = 0; v < NV; ++v) {
    for (int i = 0; i < world; ++i) {
      for (size_t j = 0; j < slab; ++j) { image[j] = view_value(v, i * slab + j); weights[j] = weight_value(v, i * slab + j); }
      rc |= lmvn_dist_set_view_slab(r[i], v, image, weights);
    }
    for (int which = 1; which <= 2; ++which) {
      fill_kernel(k, K * K * K, K * K * K / 2, (float)(v + which), 0);
      for (int phase = 0; phase < 2; ++phase) {
        for (int i = 0; i < world; ++i) rc |= lmvn_dist_psf_phase(r[i], v, which, phase, phase == 0 ? k : NULL, kdims);
        for (int i = 0; i < world; ++i) rc |= lmvn_dist_synchronize(r[i]);
      }
    }
  }
  for (int i = 0; i < world; ++i) {
    for (size_t j = 0; j < slab; ++j) psi[j] = psi0_value(i * slab + j);
    rc |= lmvn_dist_set_psi_slab(r[i], psi);
  }
  for (int it = 0; it < 2; ++it)
    for (int v = 0; v < NV; ++v)
      for (int which = 1; which <= 2; ++which)
        for (int phase = 0; phase < 3; ++phase) {
          for (int i = 0; i < world; ++i) rc |= lmvn_dist_conv_phase(r[i], v, which, phase, 0.006, 1e-3f);
          for (int i = 0; i < world; ++i) rc |= lmvn_dist_synchronize(r[i]);
        }
  if (rc != 0) { printf("%-34s failed: %s\n", name, lmvn_last_error()); ++g_failed; }
  for (int i = 0; i < world; ++i) {
    char nm[64];
    snprintf(nm, sizeof nm, "%s/rank%d", name, i);
    if (lmvn_dist_get_psi_slab(r[i], psi) != 0) { printf("%-34s get_psi failed: %s\n", nm, lmvn_last_error()); ++g_failed; }
    else report_recurrence(nm, psi, slab, i * slab, NV, 2, 0.006, 1e-3);
  }
  free(image); free(weights); free(psi);
done:
  for (int i = 0; i < world; ++i) if (r[i]) lmvn_dist_destroy(r[i]);
}

static void legacy_case(void) {
  if (skipped("legacy")) return;
  enum { N = 5000 };
  float *a = malloc(N * sizeof(float)), *b = malloc(N * sizeof(float)), *w = malloc(N * sizeof(float));
  for (int i = 0; i < N; ++i) { a[i] = 10.f; b[i] = 5.f; w[i] = 0.1f; }
  compute_quotient(a, b, N, 0);
  report("legacy/compute_quotient", b, NULL, N, 2.0, 1e-6);
  for (int i = 0; i < N; ++i) { a[i] = 5.f; b[i] = 42.f; }
  compute_final_values(a, b, w, N, 1e-4f, 0.0, 0);
  report("legacy/compute_final_values", a, NULL, N, 25.5, 1e-6);
  free(a); free(b); free(w);
  {
    int dims[3] = {16, 16, 32}, kdims[3] = {3, 3, 3};
    const size_t n = 16 * 16 * 32;
    float *in = malloc(n * sizeof(float)), *out = malloc(n * sizeof(float)), k[27] = {0};
    for (size_t i = 0; i < n; ++i) in[i] = 4.f;
    k[13] = 1.f;
    iterate_fft_plain(in, k, out, dims, kdims, 0);
    /* psi = in; integral = in / (in * 1) = 1; (x) kernel2 (0.1 everywhere, 27 taps) = 2.7; psi * 2.7 */
    report("legacy/iterate_fft_plain", out, NULL, n, 4.0 * 2.7, 5e-5);
    free(in); free(out);
  }
}

int main(int argc, char** argv) {
  g_filters = argv + 1;
  g_num_filters = argc - 1;
  if (getNumDevicesCUDA() < 1) { printf("no CUDA device\n"); return 0; }
  printf("%s\n", lmvn_version());
  /* power-of-two fast path: every rows configuration (nx) and every strided plan (ny, nz) */
  deconv_case("fast/16x16x32", 16, 16, 32, 3, 2, 2, 0.006, 0);
  deconv_case("fast/32x64x64", 32, 64, 64, 3, 2, 2, 0.0, 0);
  deconv_case("fast/64x128x128", 64, 128, 128, 5, 2, 2, 0.006, 0);
  deconv_case("fast/128x16x256", 128, 16, 256, 3, 2, 2, 0.006, 0);
  deconv_case("fast/256x16x512", 256, 16, 512, 3, 1, 2, 0.006, 0);
  deconv_case("fast/16x512x1024", 16, 512, 1024, 3, 1, 2, 0.006, 0);
  deconv_case("fast/1024x16x32", 1024, 16, 32, 3, 1, 2, 0.006, 0);
  deconv_case("fast/16x1024x32", 16, 1024, 32, 3, 1, 2, 0.0, 0);
  deconv_case("huge/1024x1024x1024", 1024, 1024, 1024, 3, 1, 1, 0.006, 0); /* 64-bit index math at the largest shape */
  /* not a fast-path shape: the one-shot call embeds it periodically */
  deconv_case("embedded/20x24x28", 20, 24, 28, 5, 2, 2, 0.006, 0);
  deconv_case("embedded/50x17x100", 50, 17, 100, 4, 2, 2, 0.0, 0);
  conv_case("conv/fast/64x64x64", 64, 64, 64, 5);
  conv_case("conv/embedded/20x24x28", 20, 24, 28, 5);
  plan_case("plan/native/32x32x64", 32, 32, 64, 3, 0);
  plan_case("plan/generic/20x24x28", 20, 24, 28, 3, 0);
  plan_case("plan/embedded/20x24x28", 20, 24, 28, 5, 1);
  plan_case("plan/zero_padded/20x24x28", 20, 24, 28, 5, 2);
  plan_case("plan/zero_padded/28x28x28", 28, 28, 28, 5, 2);
  lmvn_set_padding(LMVN_PAD_ZERO);
  deconv_case("zero/20x24x28", 20, 24, 28, 5, 2, 2, 0.006, 1);
  conv_case("conv/zero/20x24x28", 20, 24, 28, 5);
  lmvn_set_padding(LMVN_PAD_NONE);
  dist_case("slabs/32x32x64/world2", 32, 32, 64, 2);
  dist_case("slabs/64x64x64/world4", 64, 64, 64, 4);
  legacy_case();
  lmvn_release_cached_memory();
  printf("%d case(s) failed\n", g_failed);
  return g_failed ? 1 : 0;
}
