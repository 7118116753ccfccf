// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny single-process emulator of the CUDA execution model, just large enough
// to run this repository's kernels on the host so that their index math can be
// debugged in a container without a GPU.  It is never linked into the product
// library (libmultiviewnative.so is built by nvcc from the same sources with
// LMVN_EMU undefined); the emulated build lands in tests/emu/_lmvn_emu.so and
// is used by `-m "not gpu"` tests only.
//
// Model: one CUDA thread = one ucontext fiber; the threads of a block are
// scheduled round-robin and yield at barriers / warp collectives; blocks run
// sequentially.  __shared__ variables become function-local statics (valid
// because only one block is alive at a time).
#pragma once
#ifndef LMVN_EMU
#error "cuda_emu.h is for the emulated test build only"
#endif

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- vector types ---------------------------------------------------------
struct alignas(8) float2 { float x, y; };  // CUDA's alignment: a misaligned access faults on the device
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

// ---- qualifiers -----------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace emu {

struct State {
  dim3 grid, block;
  unsigned nthreads = 0;
  unsigned alive = 0;
  unsigned cur = 0;
  // block barrier
  unsigned bar_count = 0;
  unsigned bar_gen = 0;
  // warp collectives
  std::vector<unsigned> warp_count, warp_gen;
  std::vector<uint64_t> warp_slots;  // [2][nwarps][32]
  std::vector<unsigned char> dyn_smem;
  ucontext_t main_ctx;
  std::vector<ucontext_t> ctx;
  std::vector<char*> stacks;
  std::vector<char> done;
  const std::function<void()>* body = nullptr;
};
State& st();
void yield();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void syncthreads();
void warp_sync(unsigned warp, unsigned nlanes);
uint64_t shfl_generic(uint64_t v, int src_lane_or_mask, int mode, int width);
unsigned lanes_in_warp(unsigned warp);

}  // namespace emu

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  emu::warp_sync(tid / 32, emu::lanes_in_warp(tid / 32));
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T>
static inline T emu_shfl(T v, int arg, int mode, int width) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  raw = emu::shfl_generic(raw, arg, mode, width);
  T out;
  std::memcpy(&out, &raw, sizeof(T));
  return out;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) { return emu_shfl(v, src, 0, width); }
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) { return emu_shfl(v, m, 1, width); }
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int d, int width = 32) { return emu_shfl(v, d, 2, width); }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d, int width = 32) { return emu_shfl(v, d, 3, width); }

// ---- intrinsics -----------------------------------------------------------
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
using std::fmaxf;
using std::fminf;
using std::isfinite;
using std::isinf;
using std::isnan;
using std::max;
using std::min;

// ---- runtime shim ---------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct emuEvent_* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 11 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
  char name[256];
  size_t totalGlobalMem;
  int major, minor, multiProcessorCount;
  size_t sharedMemPerBlockOptin;
};
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emu error"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  std::memset(p, 0, sizeof(*p));
  std::snprintf(p->name, sizeof(p->name), "lmvn host emulator");
  p->totalGlobalMem = size_t(8) << 30;
  p->major = 10; p->minor = 0; p->multiProcessorCount = 4;
  p->sharedMemPerBlockOptin = 227 * 1024;
  return 0;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  // poison so that reads of uninitialised device memory show up as NaN
  *p = std::malloc(n ? n : 1);
  if (!*p) return cudaErrorMemoryAllocation;
  std::memset(*p, 0xff, n);
  return 0;
}
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return 0; }
static inline cudaError_t cudaHostUnregister(void*) { return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height,
                                            cudaMemcpyKind, cudaStream_t = 0) {
  for (size_t r = 0; r < height; ++r)
    std::memmove(static_cast<char*>(d) + r * dpitch, static_cast<const char*>(s) + r * spitch, width);
  return 0;
}
struct cudaPitchedPtr { void* ptr; size_t pitch, xsize, ysize; };
struct cudaPos { size_t x, y, z; };
struct cudaExtent { size_t width, height, depth; };
struct cudaMemcpy3DParms { cudaPitchedPtr srcPtr, dstPtr; cudaPos srcPos, dstPos; cudaExtent extent; cudaMemcpyKind kind; };
static inline cudaPitchedPtr make_cudaPitchedPtr(void* p, size_t pitch, size_t xs, size_t ys) { return cudaPitchedPtr{p, pitch, xs, ys}; }
static inline cudaPos make_cudaPos(size_t x, size_t y, size_t z) { return cudaPos{x, y, z}; }
static inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { return cudaExtent{w, h, d}; }
static inline cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms* p, cudaStream_t = 0) {
  for (size_t z = 0; z < p->extent.depth; ++z)
    for (size_t y = 0; y < p->extent.height; ++y) {
      const char* s = static_cast<const char*>(p->srcPtr.ptr) + ((p->srcPos.z + z) * p->srcPtr.ysize + p->srcPos.y + y) * p->srcPtr.pitch + p->srcPos.x;
      char* d = static_cast<char*>(p->dstPtr.ptr) + ((p->dstPos.z + z) * p->dstPtr.ysize + p->dstPos.y + y) * p->dstPtr.pitch + p->dstPos.x;
      std::memmove(d, s, p->extent.width);
    }
  return 0;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = size_t(8) << 30; return 0; }
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
enum { cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return 0; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return 0; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return 0; }

namespace emu {
// arguments are evaluated by the caller and captured BY VALUE, like a real launch
template <typename K, typename... A>
static inline void launch_kernel(dim3 grid, dim3 block, size_t smem, K kernel, A... args) {
  std::function<void()> body = [=]() { kernel(args...); };
  launch(grid, block, smem, body);
}
}  // namespace emu
#define LMVN_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch_kernel((grid), (block), (smem), kernel, __VA_ARGS__)
#define LMVN_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::st().dyn_smem.data())
