// stride_bench.cu -- does the plane stride of the z pass matter to HBM?  A z-pass-shaped copy (64 KB tiles of 512 rows x
// 16 columns, 2 CTAs/SM, every thread 32 loads of 8 bytes one plane apart, in place: 32 stores to the same addresses, plus
// 32 more loads from a second array like K^) over a [512 planes][512 rows][128 columns] complex array whose plane pitch
// is 512 * 128 elements (512 KB: what the split layout gives for config 3) plus a padding of 0, 16, 128, ... elements.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stride_bench stride_bench.cu && ./stride_bench
#include <cstdio>
#include <cuda_runtime.h>

template <bool KHAT>
__global__ void __launch_bounds__(256, 2) k_zcopy(float2* __restrict__ data, const float2* __restrict__ khat, long long pitch) {
  const int c = threadIdx.x % 16, rg = threadIdx.x / 16;
  const int tile = blockIdx.x;           // (y, bx): 8 tiles of 16 columns per row
  const long long base = (long long)(tile / 8) * 128 + (tile % 8) * 16 + c;
  float2 v[32];
  float2* p = data + base + rg * pitch;
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r] = __ldcg(p + (long long)r * 16 * pitch);
  if (KHAT) {
    const float2* k = khat + base + rg * pitch;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float2 w = __ldcg(k + (long long)r * 16 * pitch);
      v[r].x = v[r].x * w.x - v[r].y * w.y;
    }
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r].y = v[r].y * 1.0001f + v[(r + 1) & 31].x;
#pragma unroll
  for (int r = 0; r < 32; ++r) __stcg(p + (long long)r * 16 * pitch, v[r]);
}

// variants of the 3C copy: how K^ is loaded (0 ld.cg, 1 ld.nc, 2 ld.cs), how the result is stored (0 st.cg, 1 st.cs, 2 st.wt),
// resident CTAs per SM, in place or to a third array
template <int KLD, int ST, int BLOCKS, bool INPLACE>
__global__ void __launch_bounds__(256, BLOCKS) k_zcopy_v(float2* __restrict__ data, const float2* __restrict__ khat, float2* __restrict__ out,
                                                         long long pitch) {
  const int c = threadIdx.x % 16, rg = threadIdx.x / 16;
  const int tile = blockIdx.x;
  const long long base = (long long)(tile / 8) * 128 + (tile % 8) * 16 + c;
  float2 v[32];
  float2* p = data + base + rg * pitch;
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r] = __ldcg(p + (long long)r * 16 * pitch);
  const float2* k = khat + base + rg * pitch;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const float2* kp = k + (long long)r * 16 * pitch;
    const float2 w = KLD == 0 ? __ldcg(kp) : (KLD == 1 ? __ldg(kp) : __ldcs(kp));
    v[r].x = v[r].x * w.x - v[r].y * w.y;
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r].y = v[r].y * 1.0001f + v[(r + 1) & 31].x;
  float2* q = (INPLACE ? data : out) + base + rg * pitch;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    float2* qp = q + (long long)r * 16 * pitch;
    if (ST == 0) __stcg(qp, v[r]);
    else if (ST == 1) __stcs(qp, v[r]);
    else __stwt(qp, v[r]);
  }
}

// 2C copies: z-shaped (rows one plane apart) or y-shaped (rows 1 KB apart inside a plane, tile = (plane, 16 columns)), in place
// or to a second array; and a rows-shaped one (every 16-lane group streams whole 1 KB rows, like the x passes)
template <bool YSHAPE, bool INPLACE>
__global__ void __launch_bounds__(256, 2) k_copy2(float2* __restrict__ data, float2* __restrict__ out, long long pitch) {
  const int c = threadIdx.x % 16, rg = threadIdx.x / 16;
  const int tile = blockIdx.x;
  const long long rs = YSHAPE ? 128 : pitch;
  const long long base = (YSHAPE ? (long long)(tile / 8) * pitch : (long long)(tile / 8) * 128) + (tile % 8) * 16 + c;
  float2 v[32];
  const float2* p = data + base + rg * rs;
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r] = __ldcg(p + (long long)r * 16 * rs);
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r].y = v[r].y * 1.0001f + v[(r + 1) & 31].x;
  float2* q = (INPLACE ? data : out) + base + rg * rs;
#pragma unroll
  for (int r = 0; r < 32; ++r) __stcg(q + (long long)r * 16 * rs, v[r]);
}
template <bool INPLACE>
__global__ void __launch_bounds__(128, 4) k_rows2(float2* __restrict__ data, float2* __restrict__ out, long long rows) {
  const int lane = threadIdx.x % 16, group = threadIdx.x / 16;
  for (long long row = ((long long)blockIdx.x * 8 + group) * 2; row < rows; row += (long long)gridDim.x * 16) {
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = __ldcg(data + (row + r / 8) * 128 + lane + 16 * (r % 8));
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r].y = v[r].y * 1.0001f + v[(r + 1) & 15].x;
    float2* q = INPLACE ? data : out;
#pragma unroll
    for (int r = 0; r < 16; ++r) __stcg(q + (row + r / 8) * 128 + lane + 16 * (r % 8), v[r]);
  }
}
// the 3C copy as a persistent kernel (2 CTAs per SM loop over the tiles), in place; PHASED = true keeps "all loads, then all
// stores" per tile (what a transform between them forces), false lets the compiler interleave them (stores through a
// pointer it may assume not to alias the loads)
template <bool PHASED>
__global__ void __launch_bounds__(256, 2) k_zcopy_persistent(float2* data, const float2* __restrict__ khat, float2* __restrict__ alias,
                                                             long long pitch, int tiles) {
  const int c = threadIdx.x % 16, rg = threadIdx.x / 16;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long base = (long long)(tile / 8) * 128 + (tile % 8) * 16 + c;
    float2 v[32];
    const float2* p = data + base + rg * pitch;
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = __ldcg(p + (long long)r * 16 * pitch);
    const float2* k = khat + base + rg * pitch;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float2 w = __ldcg(k + (long long)r * 16 * pitch);
      v[r].x = v[r].x * w.x - v[r].y * w.y;
    }
    if (PHASED) {
#pragma unroll
      for (int r = 0; r < 32; ++r) v[r].y = v[r].y * 1.0001f + v[(r + 1) & 31].x;
    }
    float2* q = (PHASED ? data : alias) + base + rg * pitch;
#pragma unroll
    for (int r = 0; r < 32; ++r) __stcg(q + (long long)r * 16 * pitch, v[r]);
  }
}

// copies in the access shape of the chained x kernels ("links") of config 3: a 16-lane group handles 2 rows per iteration --
// spectrum rows (128 complex = 1 KB) in and out in place, real rows (256 floats = 1 KB) of the operands in, psi out.
// STREAMS = 3: quotient link (spectrum in/out, view in); 5: update link (spectrum in/out, psi in/out, weights in).
template <int STREAMS>
__global__ void __launch_bounds__(128, 4) k_link_copy(float2* spec, const float2* __restrict__ opa, const float2* __restrict__ opb,
                                                      float2* psi, long long rows) {
  const int lane = threadIdx.x % 16, group = threadIdx.x / 16;
  for (long long row = ((long long)blockIdx.x * 8 + group) * 2; row < rows; row += (long long)gridDim.x * 16) {
    float2 v[16], a[16], b[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = __ldcg(spec + (row + r / 8) * 128 + lane + 16 * (r % 8));
#pragma unroll
    for (int r = 0; r < 16; ++r) a[r] = __ldcg((STREAMS == 5 ? psi : opa) + (row + r / 8) * 128 + lane + 16 * (r % 8));
    if (STREAMS == 5) {
#pragma unroll
      for (int r = 0; r < 16; ++r) b[r] = __ldcg(opb + (row + r / 8) * 128 + lane + 16 * (r % 8));
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      v[r].x = v[r].x * a[r].x + (STREAMS == 5 ? b[r].y : 1.f);
      v[r].y = v[r].y * 1.0001f + v[(r + 1) & 15].x;  // the whole group of rows before the first store, like a transform
    }
    if (STREAMS == 5) {
#pragma unroll
      for (int r = 0; r < 16; ++r) __stcg(psi + (row + r / 8) * 128 + lane + 16 * (r % 8), v[r]);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) __stcg(spec + (row + r / 8) * 128 + lane + 16 * (r % 8), v[(r + 3) & 15]);
  }
}

// the merged z pass of a 1024-point axis (1024^3: 513 -> 512 columns, plane pitch 1024 x 512): COLS = 16 -> 128 KB tiles, 512
// threads, one CTA per SM (shipped); COLS = 8 -> 64 KB tiles, 256 threads, two CTAs per SM (half lines).  3C, phased, in place.
template <int COLS>
__global__ void __launch_bounds__(COLS * 32, COLS == 16 ? 1 : 2) k_z1024_copy(float2* data, const float2* __restrict__ khat, long long pitch,
                                                                             int tiles_x) {
  extern __shared__ float2 dummy[];  // occupies the tile's shared memory so that the residency matches
  const int c = threadIdx.x % COLS, rg = threadIdx.x / COLS;  // 32 row groups
  const int tile = blockIdx.x;
  const long long base = (long long)(tile / tiles_x) * 512 + (tile % tiles_x) * COLS + c;
  float2 v[32];
  float2* p = data + base + rg * pitch;
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r] = __ldcg(p + (long long)r * 32 * pitch);
  const float2* k = khat + base + rg * pitch;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const float2 w = __ldcg(k + (long long)r * 32 * pitch);
    v[r].x = v[r].x * w.x - v[r].y * w.y;
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r].y = v[r].y * 1.0001f + v[(r + 1) & 31].x;
  if (v[0].x == 123.456f) dummy[threadIdx.x] = v[1];
#pragma unroll
  for (int r = 0; r < 32; ++r) __stcg(p + (long long)r * 32 * pitch, v[r]);
}

template <typename F>
static void time_it(const char* name, double bytes, cudaEvent_t e0, cudaEvent_t e1, F launch) {
  float best = 1e9f;
  for (int rep = 0; rep < 8; ++rep) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 1 && ms < best) best = ms;
  }
  printf("%-68s: %.4f ms  %.0f GB/s  [%s]\n", name, best, bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

template <int KLD, int ST, int BLOCKS, bool INPLACE>
static void run_variant(const char* name, float2* a, float2* b, float2* c, cudaEvent_t e0, cudaEvent_t e1) {
  const long long pitch = 512LL * 128;
  float best = 1e9f;
  for (int rep = 0; rep < 8; ++rep) {
    cudaEventRecord(e0);
    k_zcopy_v<KLD, ST, BLOCKS, INPLACE><<<512 * 8, 256>>>(a, b, c, pitch);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 1 && ms < best) best = ms;
  }
  printf("3C copy, %-58s: %.4f ms  %.0f GB/s  [%s]\n", name, best, 3.0 * 512 * 512 * 128 * 8 / best * 1e-6,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int pads[] = {0, 16, 32, 128, 144, 1024, 2064};
  const long long max_pitch = 512LL * 128 + 4096;
  const size_t elems = size_t(512) * max_pitch;
  float2 *a, *b;
  cudaMalloc(&a, elems * sizeof(float2));
  cudaMalloc(&b, elems * sizeof(float2));
  cudaMemset(a, 0, elems * sizeof(float2));
  cudaMemset(b, 0, elems * sizeof(float2));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int kh = 0; kh < 2; ++kh)
    for (int pad : pads) {
      const long long pitch = 512LL * 128 + pad;
      float best = 1e9f;
      for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(e0);
        if (kh) k_zcopy<true><<<512 * 8, 256>>>(a, b, pitch);
        else k_zcopy<false><<<512 * 8, 256>>>(a, b, pitch);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 1 && ms < best) best = ms;
      }
      const double bytes = (kh ? 3.0 : 2.0) * 512 * 512 * 128 * 8;
      printf("%s plane pitch 512 KB + %5d B: %.4f ms  %.0f GB/s  [%s]\n", kh ? "data + K^ (3C)" : "data only (2C) ", pad * 8, best,
             bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
  float2* c;
  cudaMalloc(&c, 2 * elems * sizeof(float2));
  cudaMemset(c, 0, 2 * elems * sizeof(float2));
  run_variant<0, 0, 2, true>("K^ ld.cg, st.cg, 2 CTAs/SM, in place (= the z pass)", a, b, c, e0, e1);
  run_variant<1, 0, 2, true>("K^ ld.nc", a, b, c, e0, e1);
  run_variant<2, 0, 2, true>("K^ ld.cs", a, b, c, e0, e1);
  run_variant<0, 1, 2, true>("st.cs", a, b, c, e0, e1);
  run_variant<0, 2, 2, true>("st.wt", a, b, c, e0, e1);
  run_variant<0, 0, 3, true>("3 CTAs/SM (85 registers)", a, b, c, e0, e1);
  run_variant<0, 0, 4, true>("4 CTAs/SM (64 registers)", a, b, c, e0, e1);
  run_variant<0, 0, 2, false>("out of place", a, b, c, e0, e1);
  run_variant<2, 1, 2, false>("K^ ld.cs, st.cs, out of place", a, b, c, e0, e1);
  const long long pitch0 = 512LL * 128;
  const double c2 = 2.0 * 512 * 512 * 128 * 8;
  const double c3 = 3.0 * 512 * 512 * 128 * 8;
  time_it("3C copy, in place, persistent (296 CTAs), loads then stores per tile", c3, e0, e1,
          [&] { k_zcopy_persistent<true><<<296, 256>>>(a, b, a, pitch0, 4096); });
  time_it("3C copy, in place, persistent, loads and stores interleaved", c3, e0, e1,
          [&] { k_zcopy_persistent<false><<<296, 256>>>(a, b, a, pitch0, 4096); });
  time_it("3C copy, in place, one tile per CTA, loads and stores interleaved", c3, e0, e1,
          [&] { k_zcopy_persistent<false><<<4096, 256>>>(a, b, a, pitch0, 4096); });
  {
    const long long rows = 512LL * 512;
    const double cs = double(rows) * 128 * 8;  // one stream
    time_it("link-shaped copy, 3 streams (quotient link), persistent 148 x 16 CTAs", 3 * cs, e0, e1,
            [&] { k_link_copy<3><<<148 * 16, 128>>>(a, b, c, c, rows); });
    time_it("link-shaped copy, 5 streams (update link), one iteration per CTA", 5 * cs, e0, e1,
            [&] { k_link_copy<5><<<unsigned(rows / 16), 128>>>(a, b, c, c + rows * 128, rows); });
    time_it("link-shaped copy, 5 streams (update link), persistent 148 x 16 CTAs", 5 * cs, e0, e1,
            [&] { k_link_copy<5><<<148 * 16, 128>>>(a, b, c, c + rows * 128, rows); });
  }
  {
    // 1024 planes x 256 rows x 512 columns (a quarter of the 1024^3 spectrum: 1 GiB per array)
    float2 *d1, *k1;
    const long long pitch1 = 256LL * 512;
    cudaMalloc(&d1, 1024 * pitch1 * sizeof(float2));
    cudaMalloc(&k1, 1024 * pitch1 * sizeof(float2));
    cudaMemset(d1, 0, 1024 * pitch1 * sizeof(float2));
    cudaMemset(k1, 0, 1024 * pitch1 * sizeof(float2));
    const double b3 = 3.0 * 1024 * pitch1 * 8;
    cudaFuncSetAttribute(k_z1024_copy<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(k_z1024_copy<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    time_it("1024-point z-shaped 3C copy, 128 KB tiles, 512 threads, 1 CTA/SM", b3, e0, e1,
            [&] { k_z1024_copy<16><<<256 * 32, 512, 128 * 1024>>>(d1, k1, pitch1, 32); });
    time_it("1024-point z-shaped 3C copy, 64 KB tiles (8 columns), 256 threads, 2 CTAs/SM", b3, e0, e1,
            [&] { k_z1024_copy<8><<<256 * 64, 256, 64 * 1024>>>(d1, k1, pitch1, 64); });
    cudaFree(d1); cudaFree(k1);
  }
  time_it("2C copy, z-shaped, in place", c2, e0, e1, [&] { k_copy2<false, true><<<512 * 8, 256>>>(a, c, pitch0); });
  time_it("2C copy, z-shaped, out of place", c2, e0, e1, [&] { k_copy2<false, false><<<512 * 8, 256>>>(a, c, pitch0); });
  time_it("2C copy, y-shaped, in place", c2, e0, e1, [&] { k_copy2<true, true><<<512 * 8, 256>>>(a, c, pitch0); });
  time_it("2C copy, y-shaped, out of place", c2, e0, e1, [&] { k_copy2<true, false><<<512 * 8, 256>>>(a, c, pitch0); });
  time_it("2C copy, rows-shaped (4 CTAs of 128 threads per SM), in place", c2, e0, e1,
          [&] { k_rows2<true><<<148 * 16, 128>>>(a, c, 512LL * 512); });
  time_it("2C copy, rows-shaped, out of place", c2, e0, e1, [&] { k_rows2<false><<<148 * 16, 128>>>(a, c, 512LL * 512); });
  return 0;
}
