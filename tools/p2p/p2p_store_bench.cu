// Development micro-benchmark (not part of the library): how fast can SM stores fill PEER memory over NVLink,
// as a function of the store shape?  One process, two GPUs, cudaDeviceEnablePeerAccess.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_store_bench p2p_store_bench.cu && ./p2p_store_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// pattern 0: 8 B per lane; a half-warp writes one 128-byte segment, consecutive segments `pitch` bytes apart (our tiles)
// pattern 1: 8 B per lane, fully contiguous; pattern 2: 16 B per lane, fully contiguous
template <int PAT>
__global__ void k_store(char* dst, size_t bytes, int pitch) {
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = size_t(gridDim.x) * blockDim.x;
  if (PAT == 0) {
    const size_t nseg = bytes / pitch;  // one 128-byte segment per pitch
    for (size_t seg = tid / 16; seg < nseg; seg += nthreads / 16)
      __stcg(reinterpret_cast<float2*>(dst + seg * pitch) + (tid % 16), make_float2(1.f, 2.f));
  } else if (PAT == 1) {
    for (size_t i = tid; i < bytes / 8; i += nthreads) __stcg(reinterpret_cast<float2*>(dst) + i, make_float2(1.f, 2.f));
  } else {
    for (size_t i = tid; i < bytes / 16; i += nthreads) __stcg(reinterpret_cast<float4*>(dst) + i, make_float4(1.f, 2.f, 3.f, 4.f));
  }
}

template <int PAT>
void run(const char* name, char* dst, size_t bytes, int pitch, int ctas, size_t payload) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_store<PAT><<<ctas, 256>>>(dst, bytes, pitch);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) k_store<PAT><<<ctas, 256>>>(dst, bytes, pitch);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-44s ctas %4d: %7.1f GB/s\n", name, ctas, payload * 5 / (ms * 1e-3) / 1e9);
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs two GPUs\n"); return 0; }
  const size_t bytes = size_t(1) << 30;
  char *local, *peer;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, bytes));
  CK(cudaSetDevice(0)); CK(cudaMalloc(&local, bytes));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  for (int pass = 0; pass < 2; ++pass) {
    char* dst = pass ? peer : local;
    printf("== stores into %s memory\n", pass ? "PEER (NVLink)" : "local");
    for (int ctas : {148, 296, 592, 1184}) {
      run<0>("8 B/lane, 128-B segments, pitch 1152", dst, bytes, 1152, ctas, bytes / 1152 * 128);
      run<0>("8 B/lane, 128-B segments, pitch 128 (contig.)", dst, bytes, 128, ctas, bytes);
      run<1>("8 B/lane contiguous", dst, bytes, 0, ctas, bytes);
      run<2>("16 B/lane contiguous", dst, bytes, 0, ctas, bytes);
    }
  }
  return 0;
}
