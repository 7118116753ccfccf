// piece_bench.cu -- micro-benchmark behind the layout decision of the 3-round-trip schedule:
// how fast can a z-pass-shaped kernel (64 KB tiles, 2 CTAs/SM, every thread 32 loads of 8 bytes at a large
// stride, then 32 stores) stream a 270 MB spectrum when the contiguous piece a half-warp touches per row is
// 128 B (today's tiles: 16 kx columns), 64 B (2 sub-tiles x 8 columns) or 32 B (4 sub-tiles x 4 columns)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o piece_bench piece_bench.cu && ./piece_bench
#include <cstdio>
#include <cuda_runtime.h>

// layout [n2 = NS][z = 512][row = 129][kx = 128] float2; a tile = (row, kx chunk of KXC columns) x all n2 x all z
template <int NS>
__global__ void __launch_bounds__(256, 2) k_piece(const float2* __restrict__ in, float2* __restrict__ out, int rows) {
  constexpr int KXC = 16 / NS;
  constexpr int NZ = 512;
  const int c = threadIdx.x % 16, rg = threadIdx.x / 16;
  const int n2 = c / KXC, kxi = c % KXC;
  const int chunks = 128 / KXC;
  const int tile = blockIdx.x;
  const int row = tile / chunks, chunk = tile % chunks;
  const long long zs = (long long)rows * 128;
  const long long base = (long long)n2 * NZ * zs + (long long)row * 128 + chunk * KXC + kxi;
  float2 v[32];
  const float2* p = in + base + rg * zs;
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r] = __ldcg(p + (long long)r * 16 * zs);
  // a little arithmetic so that nothing is optimised away
#pragma unroll
  for (int r = 0; r < 32; ++r) { v[r].x = v[r].x * 1.0001f + v[(r + 1) & 31].y; }
  float2* q = out + base + rg * zs;
#pragma unroll
  for (int r = 0; r < 32; ++r) __stcg(q + (long long)r * 16 * zs, v[r]);
}

int main() {
  const int rows = 129;
  const size_t elems = size_t(4) * 512 * rows * 128;
  float2 *a, *b;
  cudaMalloc(&a, elems * sizeof(float2));
  cudaMalloc(&b, elems * sizeof(float2));
  cudaMemset(a, 0, elems * sizeof(float2));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ns = 1; ns <= 4; ns *= 2) {
    // every variant moves NS/4 of the array (n2 < NS); bytes = 2 * NS * 512 * rows * 128 * 8
    const int kxc = 16 / ns, tiles = rows * (128 / kxc);
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      if (ns == 1) k_piece<1><<<tiles, 256>>>(a, b, rows);
      if (ns == 2) k_piece<2><<<tiles, 256>>>(a, b, rows);
      if (ns == 4) k_piece<4><<<tiles, 256>>>(a, b, rows);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double bytes = 2.0 * ns * 512 * rows * 128 * 8;
    printf("piece %3d B (%d sub-tiles x %2d columns): %.4f ms  %.0f GB/s  [%s]\n", kxc * 8, ns, kxc, best, bytes / best * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
