// lmvn_common.cuh -- shared definitions of the sm_100a build.
//
// Every kernel in csrc/ is written against CUDA directly.  The only indirection is
// the pair of macros LMVN_LAUNCH / LMVN_DYN_SMEM, which exist so that the very
// same sources can also be compiled by g++ against tests/emu/cuda_emu.h (a host
// emulator of blocks/threads/barriers used by the CPU-only tests to check index
// math).  The product library never defines LMVN_EMU.
#pragma once

#ifdef LMVN_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define LMVN_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define LMVN_DYN_SMEM(type, name)                                  \
  extern __shared__ __align__(16) unsigned char name##_raw_[];     \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

// NVTX ranges around the phases of a call (plan, uploads + PSF spectra, loop, download): visible in Nsight Systems and
// `ncu --nvtx`; header-only NVTX 3, nothing to link, a few nanoseconds when no tool is attached.
#if !defined(LMVN_EMU) && !defined(LMVN_NO_NVTX)
#include <nvtx3/nvToolsExt.h>
namespace lmvn {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace lmvn
#else
namespace lmvn {
struct NvtxRange {
  explicit NvtxRange(const char*) {}
};
}  // namespace lmvn
#endif

// Test builds under AddressSanitizer (host emulation, -DLMVN_ARENA_REDZONE): poisoned red zones between the
// sub-buffers carved out of a device arena, so that index math that strays into a neighbouring buffer is reported.
#if defined(LMVN_EMU) && defined(LMVN_ARENA_REDZONE)
#include <sanitizer/asan_interface.h>
namespace lmvn {
static const size_t kArenaRedzone = 4096;
inline void arena_poison(void* p, size_t n) { ASAN_POISON_MEMORY_REGION(p, n); }
inline void arena_unpoison(void* p, size_t n) { ASAN_UNPOISON_MEMORY_REGION(p, n); }
}  // namespace lmvn
#else
namespace lmvn {
static const size_t kArenaRedzone = 0;
inline void arena_poison(void*, size_t) {}
inline void arena_unpoison(void*, size_t) {}
}  // namespace lmvn
#endif

namespace lmvn {

typedef float2 cplx;

__host__ __device__ __forceinline__ cplx cmake(float re, float im) { return make_float2(re, im); }
// complex add / sub: on sm_100 one packed FP32x2 instruction each (FADD2 / FFMA2)
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_float2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, float s) { return make_float2(a.x * s, a.y * s); }
// multiply by -i (forward quarter turn) and +i
__host__ __device__ __forceinline__ cplx cmul_mi(cplx a) { return make_float2(a.y, -a.x); }
__host__ __device__ __forceinline__ cplx cmul_pi(cplx a) { return make_float2(-a.y, a.x); }

// streaming global accesses: L2 only, keep L1 for the twiddle tables
__device__ __forceinline__ cplx ld_stream(const cplx* p) {
#ifdef LMVN_EMU
  return *p;
#elif defined(LMVN_DIAG_DRY_LOADS)  // diagnostic build: no global loads (the SM-side floor of a pass; results are garbage)
  return cmake(__int_as_float(int(reinterpret_cast<size_t>(p)) | 0x3f000000), 1.f);
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) {
#ifdef LMVN_EMU
  *p = v;
#elif defined(LMVN_DIAG_DRY_STORES)  // diagnostic build: no global stores
  if (v.x == 123.456f) __stcg(p, v);
#else
  __stcg(p, v);
#endif
}

// Makes the value of a bumped pointer opaque to the optimiser.  Without it nvcc re-derives every address of
// an unrolled access sequence from the base (five integer instructions per access: zero extension, carry
// chain, scaled 64-bit add); with it the bump stays ONE IMAD.WIDE per access.
#if defined(LMVN_EMU) || !defined(__CUDA_ARCH__)
#define LMVN_KEEP_PTR(p) ((void)0)
#else
#define LMVN_KEEP_PTR(p) asm volatile("" : "+l"(p))
#endif

// L2 prefetch of the 128-byte line holding p (no register, no scoreboard entry)
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef LMVN_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// ---- error plumbing: nothing in this library exits or throws across the ABI ----
void set_last_error(const char* fmt, ...);
const char* last_error();
void clear_last_error();

#define LMVN_CUDA_TRY(expr)                                                               \
  do {                                                                                    \
    cudaError_t lmvn_e_ = (expr);                                                         \
    if (lmvn_e_ != cudaSuccess) {                                                         \
      ::lmvn::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(lmvn_e_), \
                             __FILE__, __LINE__);                                         \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)

#define LMVN_TRY(expr)            \
  do {                            \
    int lmvn_r_ = (expr);         \
    if (lmvn_r_ != 0) return lmvn_r_; \
  } while (0)

static inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

}  // namespace lmvn
