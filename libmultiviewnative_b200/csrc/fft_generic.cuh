// fft_generic.cuh -- shape-agnostic 3-D R2C/C2R FFT convolution passes.
//
// This is the "any dims" path of the in-house transform (the reference accepts
// arbitrary extents: its tests use 10^3, 16x18x14, 13x17x19, 25^3, 27^3 ...,
// ref: tests/test_gpu_convolve_impl.cu:422-530, tests/test_fftw_numerical_stability.cpp:34-37).
// Every axis is transformed by a shared-memory Stockham auto-sort FFT whose
// radices are the prime factors of the length (4s first); a radix-p stage is a
// direct p-point DFT, so prime lengths degrade to O(n^2) per line but stay
// exact in structure.  Twiddles come from a per-length table computed on the
// host in double precision (never __sinf/--use_fast_math).
//
// One convolution = 5 launches:
//   k_rows_fwd   real rows (or the wrapped small kernel, by index math -- the
//                padded kernel volume of ref: inc/padd_utils.h:11-40 is never
//                materialised) -> half spectrum along x
//   k_cols (y)   forward along y
//   k_cols (z)   forward along z, multiply by K^, inverse along z   [merged]
//   k_cols (y)   inverse along y
//   k_rows_inv   inverse along x + fused pointwise epilogue (store / quotient /
//                RL update), ref: inc/cpu_kernels.h:19-90
// The power-of-two fast path (fft_fast.cuh) replaces these with three fused
// passes; this file is also its on-device reference in the tests.
#pragma once
#include "fft_types.cuh"
#include "lmvn_common.cuh"
#include "pointwise.cuh"

namespace lmvn {
namespace gen {

// FFT of L interleaved lines held in shared memory: element j of line l lives at
// a[j*L + l].  `b` is scratch of the same size.  Returns the buffer that holds
// the (natural order) result.  All threads of the block must call it; the
// caller has synchronised after filling `a`.
__device__ __forceinline__ cplx* smem_fft(cplx* a, cplx* b, const AxisPlan& P, int L, bool inverse) {
  const int n = P.n;
  int Ns = 1;
  for (int f = 0; f < P.nf; ++f) {
    const int R = P.factors[f];
    const int m = n / R;
    const int tstep = n / (Ns * R);
    const int total = n * L;
    for (int w = threadIdx.x; w < total; w += blockDim.x) {
      const int l = w % L;
      const int t = w / L;
      const int j = t % m;
      const int q = t / m;
      const int k = j % Ns;
      const int j0 = (j / Ns) * Ns * R + k;
      const int e = (k + q * Ns) * tstep;  // < n
      float ax = 0.f, ay = 0.f;
      int idx = 0;
      for (int r = 0; r < R; ++r) {
        const cplx v = a[(j + r * m) * L + l];
        cplx tw = P.tw[idx];
        if (inverse) tw.y = -tw.y;
        ax += v.x * tw.x - v.y * tw.y;
        ay += v.x * tw.y + v.y * tw.x;
        idx += e;
        if (idx >= n) idx -= n;
      }
      b[(j0 + q * Ns) * L + l] = cmake(ax, ay);
    }
    __syncthreads();
    cplx* t2 = a; a = b; b = t2;
    Ns *= R;
  }
  return a;
}

// ---- pass 1: real rows -> half spectrum along x --------------------------------
static __global__ void k_rows_fwd(RealSource src, cplx* __restrict__ spec, int nz, int ny, int nx,
                           AxisPlan P, int L) {
  LMVN_DYN_SMEM(cplx, sm);
  cplx* a = sm;
  cplx* b = sm + size_t(nx) * L;
  const long long rows = (long long)nz * ny;
  const long long row0 = (long long)blockIdx.x * L;
  const int nxc = nx / 2 + 1;
  for (int w = threadIdx.x; w < nx * L; w += blockDim.x) {
    const int l = w / nx, x = w % nx;
    const long long row = row0 + l;
    float v = 0.f;
    if (row < rows) {
      if (!src.wrapped) {
        v = src.data[row * nx + x];
      } else {
        const int z = int(row / ny), y = int(row % ny);
        const int sz = wrap_src_index(z, nz, src.kz);
        const int sy = wrap_src_index(y, ny, src.ky);
        const int sx = wrap_src_index(x, nx, src.kx);
        if (sz >= 0 && sy >= 0 && sx >= 0) v = src.data[(size_t(sz) * src.ky + sy) * src.kx + sx];
      }
    }
    a[x * L + l] = cmake(v, 0.f);
  }
  __syncthreads();
  cplx* r = smem_fft(a, b, P, L, false);
  for (int w = threadIdx.x; w < nxc * L; w += blockDim.x) {
    const int l = w / nxc, kx = w % nxc;
    const long long row = row0 + l;
    if (row < rows) spec[row * nxc + kx] = r[kx * L + l];
  }
}

// ---- passes 2-4: complex FFT along a strided axis -------------------------------
// element (o, j, i) lives at data[o*ostride + j*jstride + i], i in [0, inner)
enum ColsMode { COLS_FWD = 0, COLS_INV = 1, COLS_FWD_MUL_INV = 2 };

static __global__ void k_cols(cplx* __restrict__ data, const cplx* __restrict__ khat, long long ostride,
                       long long jstride, long long inner, AxisPlan P, int L, int mode, float scale) {
  LMVN_DYN_SMEM(cplx, sm);
  const int n = P.n;
  cplx* a = sm;
  cplx* b = sm + size_t(n) * L;
  const long long i0 = (long long)blockIdx.x * L;
  const long long base = (long long)blockIdx.y * ostride;
  for (int w = threadIdx.x; w < n * L; w += blockDim.x) {
    const int j = w / L, l = w % L;
    a[w] = (i0 + l < inner) ? data[base + j * jstride + i0 + l] : cmake(0.f, 0.f);
  }
  __syncthreads();
  cplx* r = smem_fft(a, b, P, L, mode == COLS_INV);
  if (mode == COLS_FWD_MUL_INV) {
    cplx* other = (r == a) ? b : a;
    for (int w = threadIdx.x; w < n * L; w += blockDim.x) {
      const int j = w / L, l = w % L;
      if (i0 + l < inner) r[w] = cmul(r[w], khat[base + j * jstride + i0 + l]);
    }
    __syncthreads();
    r = smem_fft(r, other, P, L, true);
  }
  for (int w = threadIdx.x; w < n * L; w += blockDim.x) {
    const int j = w / L, l = w % L;
    if (i0 + l < inner) data[base + j * jstride + i0 + l] = cscale(r[w], scale);
  }
}

// ---- pass 5: half spectrum along x -> real rows + pointwise epilogue -------------
static __global__ void k_rows_inv(const cplx* __restrict__ spec, float* __restrict__ out, int nz, int ny,
                           int nx, AxisPlan P, int L, Epilogue ep) {
  LMVN_DYN_SMEM(cplx, sm);
  cplx* a = sm;
  cplx* b = sm + size_t(nx) * L;
  const long long rows = (long long)nz * ny;
  const long long row0 = (long long)blockIdx.x * L;
  const int nxc = nx / 2 + 1;
  for (int w = threadIdx.x; w < nx * L; w += blockDim.x) {
    const int l = w / nx, x = w % nx;
    const long long row = row0 + l;
    cplx v = cmake(0.f, 0.f);
    if (row < rows) {
      // Hermitian extension of the 1-D spectrum of a real row
      v = (x < nxc) ? spec[row * nxc + x] : cconj(spec[row * nxc + (nx - x)]);
    }
    a[x * L + l] = v;
  }
  __syncthreads();
  cplx* r = smem_fft(a, b, P, L, true);
  for (int w = threadIdx.x; w < nx * L; w += blockDim.x) {
    const int l = w / nx, x = w % nx;
    const long long row = row0 + l;
    if (row >= rows) continue;
    const size_t idx = size_t(row) * nx + x;
    const float val = r[x * L + l].x * ep.scale;
    if (ep.mode == EPI_STORE) {
      out[idx] = val;
    } else if (ep.mode == EPI_QUOTIENT) {
      out[idx] = quotient(ep.view[idx], val, ep.zero_view_guard);
    } else {
      ep.psi[idx] = rl_update(ep.psi[idx], val, ep.weights[idx], ep.up);
    }
  }
}

// multiply a whole spectrum by a constant (kernel-spectrum precompute: the 1/N of
// ref: inc/cpu_convolve.h:271-278 is folded into K^, decision q10)
static __global__ void k_scale_spectrum(cplx* __restrict__ data, size_t n, float s) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) data[i] = cscale(data[i], s);
}

}  // namespace gen
}  // namespace lmvn
