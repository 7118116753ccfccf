/*
 * lmvn_b200.h -- additions of the B200-native build that sit NEXT TO the
 * reference C ABI (include/multiviewnative.h) without changing it.
 *
 *  - lmvn_last_error():  the reference reports failures by exit()ing the host
 *    process (ref: inc/cuda_helpers.cuh:17-24); this build never does, the text
 *    of the last failure of the calling thread is available here instead.
 *  - lmvn_plan_*: a persistent handle for the same deconvolution
 *    (SURVEY.md §8f-1).  Fiji drives the reference by calling
 *    inplace_gpu_deconvolve repeatedly with few iterations each
 *    (ref: bench/bench_gpu_deconvolve.cu:48-49); the handle keeps views, weights
 *    and the precomputed PSF spectra resident so that only psi moves.  The
 *    one-shot entry points of multiviewnative.h are thin wrappers over it.
 *  - lmvn_debug_*: forward / inverse transforms in natural layout, for tests.
 *
 * All functions return 0 on success and non-zero on failure (message via
 * lmvn_last_error()).  Plain C, plain pointers, no C++ or torch types.
 */
#ifndef LMVN_B200_EXT_H
#define LMVN_B200_EXT_H

#include "multiviewnative.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lmvn_plan lmvn_plan;

enum lmvn_strategy {
  LMVN_STRATEGY_AUTO = 0,    /* fused power-of-two path when the shape allows, else generic */
  LMVN_STRATEGY_GENERIC = 1, /* any shape: five separable passes per convolution             */
  LMVN_STRATEGY_FUSED = 2    /* three fused passes per convolution (power-of-two extents)    */
};

typedef struct lmvn_plan_info {
  int dims[3];
  int num_views;
  int device;
  int strategy;                        /* resolved strategy, LMVN_STRATEGY_GENERIC or _FUSED */
  int launches_per_view_iteration;     /* kernels of this library per (view, iteration)      */
  unsigned long long arena_bytes;      /* device memory held by the handle                    */
  unsigned long long real_bytes;       /* S = 4*N                                             */
  unsigned long long spectrum_bytes;   /* C = 8*nz*ny*(nx/2+1)                                */
  unsigned long long alg_bytes_per_view_iteration; /* 7S + 10C                */
} lmvn_plan_info;

LMVN_EXPORT const char* lmvn_last_error(void);
LMVN_EXPORT void lmvn_clear_error(void);
LMVN_EXPORT const char* lmvn_version(void);

/* process-wide default used by plans created afterwards (also env LMVN_STRATEGY=generic|fused) */
LMVN_EXPORT int lmvn_set_default_strategy(int strategy);

LMVN_EXPORT int lmvn_plan_create(lmvn_plan** out, const int* dims_zyx, int num_views, int device);
LMVN_EXPORT void lmvn_plan_destroy(lmvn_plan* plan);
LMVN_EXPORT int lmvn_plan_get_info(const lmvn_plan* plan, lmvn_plan_info* info);

/* Host pointers.  Uploads image and weights of one view and precomputes both PSF spectra
 * (kernel wrap-around + forward transform + 1/N, ref: src/multiviewnative.cpp:146-174). */
LMVN_EXPORT int lmvn_plan_set_view(lmvn_plan* plan, int view, const float* image,
                                   const float* weights, const float* kernel1,
                                   const int* kernel1_dims, const float* kernel2,
                                   const int* kernel2_dims);
LMVN_EXPORT int lmvn_plan_set_psi(lmvn_plan* plan, const float* psi);
LMVN_EXPORT int lmvn_plan_get_psi(lmvn_plan* plan, float* psi);

/* `iterations` sweeps over all views on the device (ref: src/multiviewnative.cpp:191-229).
 * device_ms (may be NULL) receives the CUDA-event time of the loop on the plan's stream. */
LMVN_EXPORT int lmvn_plan_iterate(lmvn_plan* plan, int iterations, double lambda, float min_value,
                                  float* device_ms);

/* psi <- psi (*) kernel{1,2} of `view`, `repeats` times (config 2: FFT convolution GB/s). */
LMVN_EXPORT int lmvn_plan_convolve(lmvn_plan* plan, int view, int which_kernel, int repeats,
                                   float* device_ms);
/* Runs ONE (view 0, iteration) with a CUDA event after every kernel launch (same stream) and
 * reports, per launch, its name, device time and algorithmic bytes.  psi is restored.
 * names: count entries of 48 chars; returns the number of launches in *count. */
LMVN_EXPORT int lmvn_plan_profile(lmvn_plan* plan, double lambda, float min_value, int max_entries,
                                  char* names, float* ms, unsigned long long* alg_bytes, int* count);
/* blocks until everything queued on the plan's stream is done */
LMVN_EXPORT int lmvn_plan_synchronize(lmvn_plan* plan);

/* r2c / c2r of a host volume through the generic passes, natural layout:
 * spectrum = nz*ny*(nx/2+1) interleaved (re,im) pairs.  c2r is unnormalised. */
LMVN_EXPORT int lmvn_debug_rfftn(const float* in, const int* dims_zyx, float* spectrum, int device);
LMVN_EXPORT int lmvn_debug_irfftn(const float* spectrum, const int* dims_zyx, float* out, int device);

#ifdef __cplusplus
}
#endif
#endif /* LMVN_B200_EXT_H */
