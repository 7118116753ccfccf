// fft_fast.cuh -- power-of-two fast path of the in-house 3-D R2C/C2R convolution.
//
// Why it looks the way it does (B200 numbers, all measured): one SM's fair share of HBM is
// ~23 B/clk while its shared-memory/L1 pipe moves 128 B/clk and it holds only 16 warps of these kernels, so a
// transform that bounces every byte through shared memory five or six times, or that spends instructions on
// address arithmetic and range-checked divisions, is limited on chip, not by HBM.  Hence:
//   * butterflies are radix-8 / 16 / 32 held in REGISTERS: a 512-point axis needs ONE shared-memory exchange
//     (32 x 16), a 256-point real row two warp-private ones;
//   * the first stage of every pass loads straight from global memory into registers and the last stage stores
//     straight from registers -- shared memory only carries the inter-stage exchanges; every strided access costs
//     one IMAD.WIDE (opaque pointer bump), not five integer instructions;
//   * strided (y, z) passes work on tiles of 16 adjacent kx columns, i.e. every global access of a half-warp is one
//     full 128-byte line; the spectrum is stored as nx/2 columns per row (whole tiles, whole lines) plus the Nyquist
//     column as a compact plane that the first CTAs of each launch transform;
//   * forward passes are decimation-in-frequency (natural in, digit-reversed out), inverse passes
//     decimation-in-time (digit-reversed in, natural out); the spectrum lives in HBM in digit-reversed (y, z) order
//     and is never unscrambled -- the PSF spectra are produced by the same passes, so the pointwise product is
//     order-agnostic;
//   * the pointwise work of the iteration is fused into the pass edges: kernel wrap-around into the first pass's
//     loads (ref: inc/padd_utils.h:11-40), the 1/N scale into K^ (ref: inc/cpu_convolve.h:271-278), the spectrum
//     product between the last forward and first inverse z stage (ref: inc/cpu_convolve.h:257-266), quotient / RL
//     update into the x-inverse pass (ref: inc/cpu_kernels.h:19-90) -- which, inside the loop, also runs the
//     x-forward pass of the NEXT convolution on the same rows (k_rows_inv_fwd), so the quotient never reaches HBM.
#pragma once
#include "fft_types.cuh"
#include "lmvn_common.cuh"
#include "pointwise.cuh"
#ifndef LMVN_EMU
#include <cuda_fp16.h>
#endif

namespace lmvn {
namespace fast {

// ------------------------------------------------------------------------------
// register butterflies.  INV = false: forward (e^{-i...}); true: inverse.
// Outputs are in natural order: v[q] = sum_r v[r] w_R^{rq}.
// ------------------------------------------------------------------------------
template <bool INV>
__device__ __forceinline__ cplx rot_q(cplx a) {  // * w_4^1  (forward: -i, inverse: +i)
  return INV ? cmake(-a.y, a.x) : cmake(a.y, -a.x);
}
template <bool INV>
__device__ __forceinline__ cplx mul_tw(cplx a, cplx w) {  // a * w (forward) or a * conj(w) (inverse)
  return INV ? cmulc(a, w) : cmul(a, w);
}

template <int R, bool INV>
struct Bfly;

template <bool INV>
struct Bfly<2, INV> {
  static __device__ __forceinline__ void run(cplx* v) {
    const cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <bool INV>
struct Bfly<4, INV> {
  static __device__ __forceinline__ void run(cplx* v) {
    const cplx c0 = cadd(v[0], v[2]), c2 = csub(v[0], v[2]);
    const cplx c1 = cadd(v[1], v[3]), c3 = rot_q<INV>(csub(v[1], v[3]));
    v[0] = cadd(c0, c1);
    v[2] = csub(c0, c1);
    v[1] = cadd(c2, c3);
    v[3] = csub(c2, c3);
  }
};

template <bool INV>
struct Bfly<8, INV> {
  static __device__ __forceinline__ void run(cplx* v) {
    const float h = 0.70710678118654752440f;
    const cplx a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    const cplx a1 = cadd(v[1], v[5]);
    cplx a5 = csub(v[1], v[5]);
    const cplx a2 = cadd(v[2], v[6]), a6 = rot_q<INV>(csub(v[2], v[6]));
    const cplx a3 = cadd(v[3], v[7]);
    cplx a7 = csub(v[3], v[7]);
    // a5 *= w8^1, a7 *= w8^3
    if (!INV) {
      a5 = cmake(h * (a5.x + a5.y), h * (a5.y - a5.x));
      a7 = cmake(h * (a7.y - a7.x), -h * (a7.x + a7.y));
    } else {
      a5 = cmake(h * (a5.x - a5.y), h * (a5.x + a5.y));
      a7 = cmake(-h * (a7.x + a7.y), h * (a7.x - a7.y));
    }
    const cplx b0 = cadd(a0, a2), b2 = csub(a0, a2);
    const cplx b1 = cadd(a1, a3), b3 = rot_q<INV>(csub(a1, a3));
    const cplx b4 = cadd(a4, a6), b6 = csub(a4, a6);
    const cplx b5 = cadd(a5, a7), b7 = rot_q<INV>(csub(a5, a7));
    v[0] = cadd(b0, b1);
    v[4] = csub(b0, b1);
    v[2] = cadd(b2, b3);
    v[6] = csub(b2, b3);
    v[1] = cadd(b4, b5);
    v[5] = csub(b4, b5);
    v[3] = cadd(b6, b7);
    v[7] = csub(b6, b7);
  }
};

template <bool INV>
struct Bfly<16, INV> {
  static __device__ __forceinline__ void run(cplx* v) {
    // 16 = 4 x 4 decimation in frequency: X[q + 4 q2] = sum_i (DFT4_r(v[i+4r])[q] * w16^{iq}) w4^{i q2}
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin(pi/8)
    const float h = 0.70710678118654752440f;
    cplx a[4][4];  // a[q][i]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cplx t[4] = {v[i], v[i + 4], v[i + 8], v[i + 12]};
      Bfly<4, INV>::run(t);
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q][i] = t[q];
    }
    // twiddles w16^{iq}: exponents 1,2,3 (i=1), 2,4,6 (i=2), 3,6,9 (i=3)
    const cplx w1 = cmake(c1, -s1), w2 = cmake(h, -h), w3 = cmake(s1, -c1);
    const cplx w6 = cmake(-h, -h), w9 = cmake(-c1, s1);
    a[1][1] = mul_tw<INV>(a[1][1], w1);
    a[2][1] = mul_tw<INV>(a[2][1], w2);
    a[3][1] = mul_tw<INV>(a[3][1], w3);
    a[1][2] = mul_tw<INV>(a[1][2], w2);
    a[2][2] = rot_q<INV>(a[2][2]);  // w16^4
    a[3][2] = mul_tw<INV>(a[3][2], w6);
    a[1][3] = mul_tw<INV>(a[1][3], w3);
    a[2][3] = mul_tw<INV>(a[2][3], w6);
    a[3][3] = mul_tw<INV>(a[3][3], w9);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      cplx t[4] = {a[q][0], a[q][1], a[q][2], a[q][3]};
      Bfly<4, INV>::run(t);
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) v[q + 4 * q2] = t[q2];
    }
  }
};

// w_32^k (forward sign), k = 0..31, from first-octant constants; k is a compile-time constant
// after unrolling, so this folds to immediates
__device__ __forceinline__ cplx w32(int k) {
  const float c[9] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                      0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                      0.19509032201612826785f, 0.f};
  // angle = -2 pi k / 32; quadrant folding on cos(a) = c[k], sin(a) = c[8-k] for k in 0..8
  k &= 31;
  const int quad = k >> 3, r = k & 7;
  const float cr = c[r], sr = c[8 - r];     // cos, sin of 2 pi r / 32
  // e^{-i (quad pi/2 + t)} = (-i)^quad (cos t - i sin t)
  switch (quad) {
    case 0: return cmake(cr, -sr);
    case 1: return cmake(-sr, -cr);
    case 2: return cmake(-cr, sr);
    default: return cmake(sr, cr);
  }
}

template <bool INV>
struct Bfly<32, INV> {
  static __device__ __forceinline__ void run(cplx* v) {
    // 32 = 8 x 4 decimation in frequency: X[q + 8 q2] = sum_i (DFT8_r(v[i+4r])[q] * w32^{iq}) w4^{i q2}
    cplx a[8][4];  // a[q][i]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cplx t[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) t[r] = v[i + 4 * r];
      Bfly<8, INV>::run(t);
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q][i] = t[q];
    }
#pragma unroll
    for (int q = 1; q < 8; ++q)
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        const int k = i * q;
        if (k == 8) a[q][i] = rot_q<INV>(a[q][i]);
        else a[q][i] = mul_tw<INV>(a[q][i], w32(k));
      }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      cplx t[4] = {a[q][0], a[q][1], a[q][2], a[q][3]};
      Bfly<4, INV>::run(t);
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) v[q + 8 * q2] = t[q2];
    }
  }
};

// ------------------------------------------------------------------------------
// strided passes (y and z axes): tile = N rows x COLS adjacent kx columns.
// ------------------------------------------------------------------------------
enum StridedMode {
  SM_FWD = 0, SM_INV = 1, SM_FWD_MUL_INV = 2, SM_FWD_SCALE = 3,
  SM_FWD_SCATTER = 4,          // SM_FWD whose final stores go to the per-destination buffers of A.sc
  SM_FWD_MUL_INV_SCATTER = 5   // SM_FWD_MUL_INV likewise
};

struct StridedArgs {
  cplx* data;             // spectrum, in place
  const cplx* khat;       // SM_FWD_MUL_INV: same layout as data
  int row_stride;         // complex elements between consecutive transform rows (fits 32 bits)
  long long tile_stride;  // between consecutive tiles of the slow tile index (blockIdx.y)
  int ncols;              // valid kx columns (nx/2+1)
  const cplx* tw1;        // stage-1 twiddles, [j][q] = w_N^{jq},      j < N/R1, q < R1
  const cplx* tw2;        // stage-2 twiddles, [j][q] = w_{N/R1}^{jq}, j < N/(R1 R2), q < R2
  float scale;            // SM_FWD_SCALE
  int prefetch;           // > 0: every CTA first pulls the tile of block (id + prefetch) into L2
  int prefetch_khat;      // SM_FWD_MUL_INV: pull the CTA's own K^ tile into L2 at kernel entry
  // opt-in half-precision PSF spectra (KH = 1 kernels): khat / nyq_khat point at __half2 elements (same indexing),
  // stored as K^ * (1 / *khat_unscale); the product uses K^ = half value * *khat_unscale
  const float* khat_unscale;
  Scatter sc;             // SM_*_SCATTER
  // Split layout (single-device engine): the half spectrum is stored as nx/2 columns per row -- a whole number of
  // tiles, rows of whole 128-byte lines, no pitch padding -- plus the Nyquist column kx = nx/2 as a compact plane
  // nyq[z'][y'].  The launch is then a flat grid: the first nyq_groups CTAs transform the Nyquist plane (a "column" of
  // such a CTA is one slow index; element (row r, slow s) sits at nyq + r * nyq_rs + s * nyq_cs), the others are the
  // tiles_x x slow ordinary tiles.  nyq_groups < 0: 2-D grid (tile, slow), Nyquist column inside the rows.
  cplx* nyq;
  const cplx* nyq_khat;
  int nyq_groups;
  int nyq_rs, nyq_cs;
  int tiles_x;
  unsigned slow;
};

// number of stages and the radix of stage s for N = R1*R2*R3
template <int N>
struct Radix;
#ifndef LMVN_NARROW_RADIX
// two stages, ONE shared-memory exchange: wide register butterflies
template <> struct Radix<512> { static const int S = 2, R1 = 32, R2 = 16, R3 = 1; };
template <> struct Radix<256> { static const int S = 2, R1 = 16, R2 = 16, R3 = 1; };
template <> struct Radix<128> { static const int S = 2, R1 = 16, R2 = 8, R3 = 1; };
#else
template <> struct Radix<512> { static const int S = 3, R1 = 8, R2 = 8, R3 = 8; };
template <> struct Radix<256> { static const int S = 3, R1 = 8, R2 = 8, R3 = 4; };
template <> struct Radix<128> { static const int S = 3, R1 = 8, R2 = 4, R3 = 4; };
#endif
// 1024 keeps three stages: a 32 x 32 plan would need 64 registers of K^ next to 64 of data in the merged z pass
template <> struct Radix<1024> { static const int S = 3, R1 = 16, R2 = 16, R3 = 4; };
template <> struct Radix<64> { static const int S = 2, R1 = 8, R2 = 8, R3 = 1; };
// ALT = 1: the plan of passes that never merge with the spectrum product (the y passes).  1024 = 32 x 32 there: two
// stages, one exchange.  The digit-reversed order of an axis depends on its plan, so an axis uses ONE plan everywhere
// (forward, inverse, PSF spectra): y may take ALT, z (merged pass, K^ in registers) never does.
template <int N, int ALT> struct Plan : Radix<N> {};
template <> struct Plan<1024, 1> { static const int S = 2, R1 = 32, R2 = 32, R3 = 1; };
template <> struct Radix<32> { static const int S = 2, R1 = 8, R2 = 4, R3 = 1; };
template <> struct Radix<16> { static const int S = 2, R1 = 4, R2 = 4, R3 = 1; };

// tile width in kx columns: N x COLS x 8 bytes of shared memory; 1024 rows take half lines (64 KB tiles)
template <int N> struct Cols { static const int V = (N >= 1024) ? 8 : ((N >= 256) ? 16 : (N >= 128 ? 32 : 64)); };
#ifdef LMVN_STRIDED_HALF
template <> struct Cols<512> { static const int V = 8; };
#endif

static const int kStridedThreads = 256;
// threads per CTA of a strided pass; LMVN_STRIDED_HALF (A/B knob): 512-point passes on half-line tiles
// (8 columns, 32 KB) with 128 threads, four CTAs per SM instead of two
template <int N> struct Threads { static const int V = kStridedThreads; };
#ifdef LMVN_STRIDED_HALF
template <> struct Threads<512> { static const int V = 128; };
#endif
// Tile shape per (length, mode).  The merged 1024-point z pass works on FULL-LINE tiles -- 16 columns, 128 KB of shared
// memory, 512 threads, one CTA per SM: every global access of a half-warp is a whole 128-byte line, where the 8-column
// tiles of the other 1024-point passes use half of every L1 wavefront.  Measured (profiles/r02_1024_wide_tiles.log): z pass
// 4.44 -> 3.27 ms at 1024^3 (-26 %), 1.10 -> 0.79 ms at 1024 x 1024 x 256; the y passes get slower on such tiles (y inverse
// +17 %: with two stages they live on the overlap of two co-resident CTAs), so they keep theirs.  LMVN_1024_NARROW_Z: A/B.
template <int N, int MODE> struct TileCols { static const int V = Cols<N>::V; };
template <int N, int MODE> struct TileThreads { static const int V = Threads<N>::V; };
#ifndef LMVN_1024_NARROW_Z
template <> struct TileCols<1024, 2> { static const int V = 16; };     // SM_FWD_MUL_INV
template <> struct TileThreads<1024, 2> { static const int V = 512; };
template <> struct TileCols<1024, 5> { static const int V = 16; };     // SM_FWD_MUL_INV_SCATTER
template <> struct TileThreads<1024, 5> { static const int V = 512; };
// the scattering y pass of the slab-decomposed plans: its stores are 128-byte segments of peer memory instead of 64-byte ones
// (link bound; the forward y pass runs equally fast on either tile shape)
template <> struct TileCols<1024, 4> { static const int V = 16; };     // SM_FWD_SCATTER
template <> struct TileThreads<1024, 4> { static const int V = 512; };
#endif
#ifndef LMVN_ZMUL_BLOCKS
#define LMVN_ZMUL_BLOCKS 2
#endif
#ifndef LMVN_ZMUL_UNROLL
#define LMVN_ZMUL_UNROLL 1
#endif
#ifndef LMVN_Y_UNROLL
#define LMVN_Y_UNROLL 8
#endif

// resident CTAs per SM the register budget is sized for
template <int N, int MODE> struct StridedBlocks {
  static const int V = (Radix<N>::R1 >= 32 || N >= 1024) ? 2
                       : (((MODE == SM_FWD_MUL_INV || MODE == SM_FWD_MUL_INV_SCATTER) && N >= 64) ? LMVN_ZMUL_BLOCKS : 3);
};

enum Where { W_SMEM = 0, W_GLOBAL = 1, W_GLOBAL_SCALED = 2, W_SCATTER = 3 };

// t[1 .. R-1] = the R - 1 twiddles of one butterfly, a table row of R complex entries (16-byte aligned: R is even).
// Two entries per 128-bit load: these passes are bound by the number of load/store instructions the SM can issue,
// and the twiddle fetches were a fifth of them (62 of 318 per thread and tile in the merged z pass).
template <int R>
__device__ __forceinline__ void load_twiddles(cplx* t, const cplx* __restrict__ row) {
  static_assert(R % 2 == 0, "pairs");
  const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int q = 0; q < R / 2; ++q) {
    const float4 f = __ldg(r4 + q);
    t[2 * q] = cmake(f.x, f.y);
    t[2 * q + 1] = cmake(f.z, f.w);
  }
}

// One radix-R stage of span L on the thread's column.  Forward (INV = false) is a
// decimation-in-frequency stage (butterfly, then twiddle w_L^{jq}); inverse is the
// decimation-in-time mirror (conjugate twiddle, then butterfly).  `sm` and `g`
// already point at this thread's column; rows are addressed with 32-bit offsets.
// PITCH: row pitch of the shared-memory tile (the two-pass schedule pads its rows, fft_x3.cuh).
// HV: the tile belongs to a group of THREADS threads of a larger CTA (fft_tma.cuh: two 256-thread halves, each on its own tile)
template <int THREADS, bool HV>
__device__ __forceinline__ int tile_tid() { return HV ? int(threadIdx.x % THREADS) : int(threadIdx.x); }

template <int N, int R, int L, int COLS, bool INV, int SRC, int DST, int UNROLL = 8, int PITCH = COLS,
          int THREADS = Threads<N>::V, bool HV = false>
__device__ __forceinline__ void strided_stage(cplx* __restrict__ sm, cplx* __restrict__ g, int rs,
                                              const cplx* __restrict__ tws, float scale,
                                              const Scatter* sc = nullptr, long long sc_tile = 0) {
  constexpr int M = L / R;
  constexpr int RG = THREADS / COLS;   // row groups per block
  constexpr int PER_THREAD = (N / R) / RG;
  static_assert(PER_THREAD >= 1, "tile too small for the block");
  const int rg = tile_tid<THREADS, HV>() / COLS;
  // UNROLL bounds how many butterflies the compiler may interleave (register pressure)
#pragma unroll(UNROLL)
  for (int i = 0; i < PER_THREAD; ++i) {
    const int bf = rg + i * RG;
    const int j = bf % M;
    const int row0 = (bf / M) * L + j;
    cplx v[R];
    if (SRC == W_SMEM) {
      const cplx* p = sm + row0 * PITCH;
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = p[r * M * PITCH];
    } else {
      // one 64-bit pointer bumped by a uniform 64-bit step: two integer instructions per access
      // (index arithmetic in 32 bits costs five: zero extension, carry chain, scaled address)
      const cplx* gp = g + (long long)row0 * rs;
      const int step = M * rs;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        v[r] = ld_stream(gp);
        gp += step;
        LMVN_KEEP_PTR(gp);
      }
    }
    cplx t[R];  // fetched before the butterfly so that their latency hides behind it
    if (M > 1) load_twiddles<R>(t, tws + j * R);
    if (INV && M > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], t[q]);
    }
    Bfly<R, INV>::run(v);
    if (!INV && M > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = cmul(v[q], t[q]);
    }
    if (DST == W_SMEM) {
      cplx* p = sm + row0 * PITCH;
#pragma unroll
      for (int q = 0; q < R; ++q) p[q * M * PITCH] = v[q];
    } else if (DST == W_SCATTER) {
      const int mask = (1 << sc->shift) - 1;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int row = row0 + q * M;
        st_stream(sc->base[row >> sc->shift] + sc_tile + (long long)(row & mask) * sc->row_stride, v[q]);
      }
    } else {
      cplx* gp = g + (long long)row0 * rs;
      const int step = M * rs;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        st_stream(gp, DST == W_GLOBAL_SCALED ? cscale(v[q], scale) : v[q]);
        gp += step;
        LMVN_KEEP_PTR(gp);
      }
    }
  }
}

// middle of the merged z pass: last forward stage (span R), spectrum product, first inverse stage.
// The K^ operands come straight from HBM; middle_load() is called BEFORE the barrier that precedes
// the middle so that their latency overlaps the barrier wait and the shared-memory reads.
template <int N, int R, int COLS, int PITCH = COLS, int KH = 0, int THREADS = Threads<N>::V, bool HV = false>
struct Middle {
  static const int RG = THREADS / COLS;
  static const int PT = (N / R) / RG;  // butterflies per thread: 1 or 2 in every plan that is used
  struct K { cplx k[PT][R]; };
  // KH = 1: gk points at __half2 elements (4 bytes per complex value): half the K^ bytes of the pass
  static __device__ __forceinline__ void load(K& o, const cplx* __restrict__ gk, int rs, float unscale = 1.f) {
    const int rg = tile_tid<THREADS, HV>() / COLS;
#pragma unroll
    for (int i = 0; i < PT; ++i) {
#if !defined(LMVN_EMU)
      if (KH) {
        const __half2* hp = reinterpret_cast<const __half2*>(gk) + (long long)((rg + i * RG) * R) * rs;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const unsigned raw = __ldcg(reinterpret_cast<const unsigned*>(hp));
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&raw));
          o.k[i][r] = cmake(f.x * unscale, f.y * unscale);
          hp += rs;
          LMVN_KEEP_PTR(hp);
        }
        continue;
      }
#endif
      const cplx* gp = gk + (long long)((rg + i * RG) * R) * rs;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        o.k[i][r] = ld_stream(gp);
        gp += rs;
        LMVN_KEEP_PTR(gp);
      }
    }
  }
  static __device__ __forceinline__ void run(cplx* __restrict__ sm, const K& o) {
    const int rg = tile_tid<THREADS, HV>() / COLS;
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      cplx* p = sm + (rg + i * RG) * R * PITCH;
      cplx v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = p[r * PITCH];
      Bfly<R, false>::run(v);
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = cmul(v[r], o.k[i][r]);
      Bfly<R, true>::run(v);
#pragma unroll
      for (int r = 0; r < R; ++r) p[r * PITCH] = v[r];
    }
  }
};

// barrier of the threads that share a tile: the whole CTA, or (HV) the 256-thread half `bar` - 1 of a two-tile CTA
template <bool HV, int THREADS>
__device__ __forceinline__ void tile_sync(int bar) {
#if !defined(LMVN_EMU)
  if (HV) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(THREADS) : "memory");
    return;
  }
#endif
  (void)bar;
  __syncthreads();
}

// One tile of a strided pass.  `sm`, `g`, `gk` already point at this thread's column; threads of
// the padding columns of a ragged last tile pass live = false: they skip the stages (every stage
// touches the thread's own column only) but still take part in the barriers.
// SRC1 = W_SMEM: the tile has already been brought into `sm` (bulk tensor copy, fft_tma.cuh) and the first stage works in
// place on it; HV / bar: see tile_sync.
template <int N, int MODE_, int U, int ALT = 0, int KH = 0, int SRC1 = W_GLOBAL, bool HV = false>
__device__ __forceinline__ void strided_tile(const StridedArgs& A, cplx* sm, cplx* g, const cplx* gk, bool live,
                                             long long sc_tile, int rs, int bar = 0) {
  typedef Plan<N, ALT> RX;
  constexpr int COLS = TileCols<N, MODE_>::V, TH = TileThreads<N, MODE_>::V;
  constexpr int R1 = RX::R1, R2 = RX::R2;
  constexpr int R3 = (RX::R3 > 1 ? RX::R3 : 2);  // placeholder radix for the dead 3-stage code of 2-stage sizes
  constexpr int L2 = N / R1, L3 = (RX::S == 3 ? N / (R1 * R2) : 2);
  constexpr bool SCAT = (MODE_ == SM_FWD_SCATTER || MODE_ == SM_FWD_MUL_INV_SCATTER);
  constexpr int MODE = (MODE_ == SM_FWD_SCATTER) ? SM_FWD : (MODE_ == SM_FWD_MUL_INV_SCATTER ? SM_FWD_MUL_INV : MODE_);
  constexpr int DSTG = SCAT ? W_SCATTER : ((MODE == SM_FWD_SCALE) ? W_GLOBAL_SCALED : W_GLOBAL);
  const Scatter* sc = &A.sc;
  if (MODE == SM_FWD || MODE == SM_FWD_SCALE) {
    if (live) strided_stage<N, R1, N, COLS, false, SRC1, W_SMEM, U, COLS, TH, HV>(sm, g, rs, A.tw1, 1.f);
    tile_sync<HV, TH>(bar);
    if (RX::S == 3) {
      if (live) strided_stage<N, R2, L2, COLS, false, W_SMEM, W_SMEM, U, COLS, TH, HV>(sm, g, rs, A.tw2, 1.f);
      tile_sync<HV, TH>(bar);
      if (live) strided_stage<N, R3, L3, COLS, false, W_SMEM, DSTG, U, COLS, TH, HV>(sm, g, rs, nullptr, A.scale, sc, sc_tile);
    } else {
      if (live) strided_stage<N, R2, L2, COLS, false, W_SMEM, DSTG, U, COLS, TH, HV>(sm, g, rs, A.tw2, A.scale, sc, sc_tile);
    }
  } else if (MODE == SM_INV) {
    if (RX::S == 3) {
      if (live) strided_stage<N, R3, L3, COLS, true, SRC1, W_SMEM, U, COLS, TH, HV>(sm, g, rs, nullptr, 1.f);
      tile_sync<HV, TH>(bar);
      if (live) strided_stage<N, R2, L2, COLS, true, W_SMEM, W_SMEM, U, COLS, TH, HV>(sm, g, rs, A.tw2, 1.f);
    } else {
      if (live) strided_stage<N, R2, L2, COLS, true, SRC1, W_SMEM, U, COLS, TH, HV>(sm, g, rs, A.tw2, 1.f);
    }
    tile_sync<HV, TH>(bar);
    if (live) strided_stage<N, R1, N, COLS, true, W_SMEM, W_GLOBAL, U, COLS, TH, HV>(sm, g, rs, A.tw1, 1.f);
  } else {  // SM_FWD_MUL_INV
    if (live) strided_stage<N, R1, N, COLS, false, SRC1, W_SMEM, U, COLS, TH, HV>(sm, g, rs, A.tw1, 1.f);
    if (RX::S == 3) {
      typedef Middle<N, R3, COLS, COLS, KH, TH, HV> MID;
      typename MID::K kk;
      tile_sync<HV, TH>(bar);
      if (live) strided_stage<N, R2, L2, COLS, false, W_SMEM, W_SMEM, U, COLS, TH, HV>(sm, g, rs, A.tw2, 1.f);
      if (live) MID::load(kk, gk, rs, KH ? __ldg(A.khat_unscale) : 1.f);
      tile_sync<HV, TH>(bar);
      if (live) MID::run(sm, kk);
      tile_sync<HV, TH>(bar);
      if (live) strided_stage<N, R2, L2, COLS, true, W_SMEM, W_SMEM, U, COLS, TH, HV>(sm, g, rs, A.tw2, 1.f);
    } else {
      typedef Middle<N, R2, COLS, COLS, KH, TH, HV> MID;
      typename MID::K kk;
      if (live) MID::load(kk, gk, rs, KH ? __ldg(A.khat_unscale) : 1.f);
      tile_sync<HV, TH>(bar);
      if (live) MID::run(sm, kk);
    }
    tile_sync<HV, TH>(bar);
    if (live) strided_stage<N, R1, N, COLS, true, W_SMEM, DSTG, U, COLS, TH, HV>(sm, g, rs, A.tw1, 1.f, sc, sc_tile);
  }
}

template <int N, int MODE> struct StridedMinBlocks {
  static const int W = StridedBlocks<N, MODE>::V * kStridedThreads / TileThreads<N, MODE>::V;
  static const int V = W < 1 ? 1 : W;
};
template <int N, int MODE, int ALT = 0, int KH = 0>
static __global__ void __launch_bounds__(TileThreads<N, MODE>::V, StridedMinBlocks<N, MODE>::V)
    k_strided(StridedArgs A) {
  static_assert(KH == 0 || MODE == SM_FWD_MUL_INV, "half-precision K^ is an option of the merged z pass");
  static_assert(ALT == 0 || (MODE != SM_FWD_MUL_INV && MODE != SM_FWD_MUL_INV_SCATTER), "the merged pass keeps the default plan");
  constexpr int COLS = TileCols<N, MODE>::V, TH = TileThreads<N, MODE>::V;
  LMVN_DYN_SMEM(cplx, smem);  // [N][COLS]
  const int c = threadIdx.x % COLS;
  constexpr bool ZMUL = (MODE == SM_FWD_MUL_INV || MODE == SM_FWD_MUL_INV_SCATTER);
  constexpr int U = ZMUL ? LMVN_ZMUL_UNROLL : LMVN_Y_UNROLL;
  constexpr int LINES = (COLS + 15) / 16;  // 128-byte lines per tile row
  unsigned bx = blockIdx.x, by = blockIdx.y;
  long long n_tiles = (long long)gridDim.x * gridDim.y, tile_id = (long long)by * gridDim.x + bx;
  unsigned tiles_x = gridDim.x;
  if (A.nyq_groups >= 0) {
    if (int(blockIdx.x) < A.nyq_groups) {
      // Nyquist plane: this thread's column is slow index s (the plane is small and L2 resident)
      const unsigned s = blockIdx.x * COLS + c;
      const long long nb = (long long)s * A.nyq_cs;
      // (KH: the K^ pointers count __half2 elements -- same element offsets, half the bytes)
      const cplx* nk = KH ? reinterpret_cast<const cplx*>(reinterpret_cast<const unsigned*>(A.nyq_khat) + nb) : A.nyq_khat + nb;
      strided_tile<N, MODE, U, ALT, KH>(A, smem + c, A.nyq + nb, nk, s < A.slow, 0, A.nyq_rs);
      return;
    }
    tiles_x = unsigned(A.tiles_x);
    tile_id = (long long)blockIdx.x - A.nyq_groups;
    n_tiles = (long long)tiles_x * A.slow;
    bx = unsigned(tile_id % tiles_x);
    by = unsigned(tile_id / tiles_x);
  }
  const int col = bx * COLS + c;
  const long long base = (long long)by * A.tile_stride + col;
  if (ZMUL && A.prefetch_khat) {
    // K^ is first needed two stages from now: start its trip from HBM to L2 right away
    const long long tb = (long long)by * A.tile_stride + bx * COLS;
    for (int i = threadIdx.x; i < N * LINES; i += TH)
      prefetch_l2(A.khat + tb + (long long)(i / LINES) * A.row_stride + (i % LINES) * 16);
  }
  if (A.prefetch > 0) {
    // the block that will take this CTA's slot next: its loads then hit L2 instead of waiting for HBM
    const long long id = tile_id + A.prefetch;
    if (id < n_tiles) {
      const long long fb = (id / tiles_x) * A.tile_stride + (id % tiles_x) * COLS;
      for (int i = threadIdx.x; i < N * LINES; i += TH) {
        const long long off = fb + (long long)(i / LINES) * A.row_stride + (i % LINES) * 16;
        prefetch_l2(A.data + off);
      }
    }
  }
  const long long sc_tile = A.sc.offset + (long long)by * A.sc.tile_stride + col;
  const cplx* gk = KH ? reinterpret_cast<const cplx*>(reinterpret_cast<const unsigned*>(A.khat) + base) : A.khat + base;
  strided_tile<N, MODE, U, ALT, KH>(A, smem + c, A.data + base, gk, col < A.ncols, sc_tile, A.row_stride);
}

// ------------------------------------------------------------------------------
// opt-in half-precision PSF spectra: K^ (float2) -> __half2 scaled by 1 / max |component|
// ------------------------------------------------------------------------------
#ifndef LMVN_EMU
static __global__ void k_khat_absmax(const cplx* __restrict__ k, size_t n, unsigned* out_bits) {
  float m = 0.f;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const cplx v = k[i];
    m = fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order like their bits
}
// dst[i] = half2(src[i] / max); *unscale = max (stored behind the half data, read by the z pass)
static __global__ void k_khat_to_half(const cplx* __restrict__ src, size_t n, const unsigned* max_bits, __half2* dst,
                                      float* unscale) {
  const float mx = fmaxf(__uint_as_float(*max_bits), 1e-38f);
  const float inv = 1.f / mx;
  if (blockIdx.x == 0 && threadIdx.x == 0) *unscale = mx;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const cplx v = src[i];
    dst[i] = __floats2half2_rn(v.x * inv, v.y * inv);
  }
}
#endif

// ------------------------------------------------------------------------------
// x passes: one real row of nx = 2M samples <-> M+1 complex bins; every global access
// is a full 128-byte line.  M = R1 * 16.  A group of 16 lanes works on RPG = 16 / R1
// rows at a time:
//   stage 1  lane j: radix-R1 butterflies over elements j + 16 r of each row, loaded
//            straight from global memory (16 lanes x 8 B = one line per row);
//   exchange through a padded slab, element (q, j) at q*17 + j;
//   stage 2  lane t: ONE radix-16 block (block t % R1 of row t / R1), no twiddles;
//   exchange: Z in natural order;
//   split    lane l: pairs (k, M-k), k = l + 16 i -> X[k], X[M-k], stored line by line.
// Stage-1 and split twiddles depend on the lane only and live in registers.  The
// inverse is the exact mirror and ends in the fused pointwise epilogue.
// ------------------------------------------------------------------------------
static const int kRowThreads = 256;

struct RowArgs {
  gen::RealSource src;   // forward: real input (volume or wrapped kernel)
  cplx* spec;            // forward: output; inverse: input
  float* out;            // inverse: real output
  gen::Epilogue ep;      // inverse: pointwise epilogue
  int nz, ny;
  int nxp;               // spectrum row pitch (complex)
  const cplx* tw_m;      // w_M table (M entries)
  const cplx* tw_nx;     // w_nx^k, k = 0..M
  int prefetch;          // pull the next loop iteration's rows into L2 one iteration ahead
  int z0, nz_wrap;       // wrapped source on a slab of planes: global index of plane 0, global nz
  const cplx* tw_h;      // nx = 1024 only: w_{M/2} table (M/2 entries) of the two half-length sub-transforms
  cplx* nyq;             // split layout: X[nx/2] of row r lives at nyq[r] instead of spec[r * nxp + nx/2]
  // chained kernel on a periodically embedded stack (EMBED): the logical box (lz, ly, lx) sits at (oz, oy, ox); a row
  // outside the box recomputes the interior row it aliases, x positions outside the box take the aliased interior
  // sample, and the forward transform of the result goes to spec_out / nyq_out (the sources are read by other CTAs)
  cplx* spec_out;
  cplx* nyq_out;
  int lz, ly, lx, oz, oy, ox;
  // plane kernels (fft_x3.cuh): spectrum row i of the tile is real-space row rr_base + i * (1 + rr_skip) of the volume
  // (both zero everywhere else: spectrum row = real-space row)
  long long rr_base;
  int rr_skip;
  // plane kernels: first pointwise operand (view for the quotient, psi for the update) of tile row i staged in shared
  // memory at op_smem + i * nx floats for i < op_smem_rows (bulk async copies, fft_x3.cuh); nullptr: read from global memory
  const float* op_smem;
  int op_smem_rows;  // tile rows [0, op_smem_rows) are staged
};
// (compile-time switch: the rows kernels that work on global spectra have no register to spare for the mapping)
template <int SP>
__device__ __forceinline__ long long real_row(const RowArgs& A, long long row) {
  return SP ? A.rr_base + row * (1 + A.rr_skip) : row;
}

// where the spectrum rows of an x pass live: SP = 0 global memory (streaming accesses), SP = 1 shared memory (the plane
// kernels transform the rows of a tile in place)
template <int SP> __device__ __forceinline__ cplx ld_spec(const cplx* p) { return SP ? *p : ld_stream(p); }
template <int SP> __device__ __forceinline__ void st_spec(cplx* p, cplx v) {
  if (SP) *p = v;
  else st_stream(p, v);
}

// a + conj(b) and a - conj(b): one packed FFMA2 each on sm_100
__device__ __forceinline__ cplx cadd_conj(cplx a, cplx b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  return __ffma2_rn(b, make_float2(1.f, -1.f), a);
#else
  return make_float2(a.x + b.x, a.y - b.y);
#endif
}
__device__ __forceinline__ cplx csub_conj(cplx a, cplx b) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  return __ffma2_rn(b, make_float2(-1.f, 1.f), a);
#else
  return make_float2(a.x - b.x, a.y + b.y);
#endif
}

// real-transform split for one (k, M-k) pair:  X[k] = E + w^k O,  X[M-k] = conj(E - w^k O)
__device__ __forceinline__ void r2c_pair(cplx zk, cplx zm, cplx w, cplx& xk, cplx& xm) {
  const cplx e = cscale(cadd_conj(zk, zm), 0.5f);   // (Zk + conj Zm)/2
  const cplx d = cscale(csub_conj(zk, zm), 0.5f);   // (Zk - conj Zm)/2
  const cplx o = cmake(d.y, -d.x);                                    // d / i
  const cplx wo = cmul(w, o);
  xk = cadd(e, wo);
  xm = cconj(csub(e, wo));
}
// inverse split (doubled, so that the result is the UNNORMALISED c2r):
//   Z[k] = E' + i O',  Z[M-k] = conj(E') + i conj(O'),  E' = X[k] + conj X[M-k],  O' = (X[k] - conj X[M-k]) conj(w^k)
__device__ __forceinline__ void c2r_pair(cplx xk, cplx xm, cplx w, cplx& zk, cplx& zm) {
  const cplx e = cadd_conj(xk, xm);
  const cplx d = csub_conj(xk, xm);
  const cplx o = cmulc(d, w);
  zk = cmake(e.x - o.y, e.y + o.x);
  zm = cmake(e.x + o.y, -e.y + o.x);
}

template <int M> struct Row2Cfg {
  static const int R1 = M / 16;           // 2, 4, 8
  static const int RPG = 16 / R1;         // rows per 16-lane group: 8, 4, 2
  static const int RS = M + R1;           // slab row pitch (= 17 * R1), conflict free for both layouts
  static const int GROUPS = kRowThreads / 16;
  static const int ROWS = GROUPS * RPG;   // rows per block iteration
  static const int PAIRS = M / 32;        // (k, M-k) pairs per lane and row
  static const int SMEM = GROUPS * RPG * RS * int(sizeof(cplx));
};

// lane-constant twiddles of the row transforms
template <int M>
struct RowTw {  // in registers
  cplx t1[Row2Cfg<M>::R1];
  cplx tp[Row2Cfg<M>::PAIRS];
  __device__ __forceinline__ void load(const RowArgs& A, int lane) {
#pragma unroll
    for (int q = 1; q < Row2Cfg<M>::R1; ++q) t1[q] = __ldg(A.tw_m + lane * q);
#pragma unroll
    for (int i = 0; i < Row2Cfg<M>::PAIRS; ++i) tp[i] = __ldg(A.tw_nx + lane + 16 * i);
  }
  __device__ __forceinline__ cplx tw1(int q) const { return t1[q]; }
  __device__ __forceinline__ cplx twp(int i) const { return tp[i]; }
};
// The same table in shared memory, [entry][lane], read at the point of use: frees 2 (R1 + PAIRS)
// registers where the epilogue operands need them.
template <int M>
struct RowTwShared {
  static const int ENTRIES = Row2Cfg<M>::R1 + Row2Cfg<M>::PAIRS;
  const cplx* base;  // + lane
  // all threads of the block call fill(), then __syncthreads()
  static __device__ __forceinline__ void fill(cplx* table, const RowArgs& A) {
    for (int i = threadIdx.x; i < ENTRIES * 16; i += blockDim.x) {
      const int e = i / 16, l = i % 16;
      table[i] = (e < Row2Cfg<M>::R1) ? __ldg(A.tw_m + l * e) : __ldg(A.tw_nx + l + 16 * (e - Row2Cfg<M>::R1));
    }
  }
  __device__ __forceinline__ cplx tw1(int q) const { return base[q * 16]; }
  __device__ __forceinline__ cplx twp(int i) const { return base[(Row2Cfg<M>::R1 + i) * 16]; }
};

// forward transform of rows row0 .. row0+RPG-1 by one 16-lane group (slab: the group's exchange area)
// forward transform of RPG rows whose samples are already in registers: v[a * R1 + r] = complex sample
// lane + 16 r of row row0 + a (the distribution the loads of rows_fwd_group produce, and exactly what the
// last butterfly of the inverse leaves behind)
template <int M, typename TW, int SP = 0>
__device__ __forceinline__ void rows_fwd_from_regs(const RowArgs& A, cplx* slab, long long row0, int lane,
                                                   const TW& T, cplx* v);

template <int M, bool WRAPPED, typename TW, int SP = 0>
__device__ __forceinline__ void rows_fwd_group(const RowArgs& A, cplx* slab, long long row0, int lane,
                                               const TW& T) {
  typedef Row2Cfg<M> CF;
  constexpr int R1 = CF::R1, RPG = CF::RPG;
  constexpr int nx = 2 * M;
  cplx v[16];
  // ---- stage 1: loads ----
#pragma unroll
  for (int a = 0; a < RPG; ++a) {
    const long long row = row0 + a;
    if (!WRAPPED) {
      const float2* in = reinterpret_cast<const float2*>(A.src.data + real_row<SP>(A, row) * nx);
#pragma unroll
      for (int r = 0; r < R1; ++r) v[a * R1 + r] = ld_stream(in + lane + 16 * r);
    } else {
      const int z = int(row / A.ny) + A.z0, y = int(row % A.ny);
      const int sz = gen::wrap_src_index(z, A.nz_wrap, A.src.kz);
      const int sy = gen::wrap_src_index(y, A.ny, A.src.ky);
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        const int n = lane + 16 * r;
        float re = 0.f, im = 0.f;
        if (sz >= 0 && sy >= 0) {
          const float* kr = A.src.data + (size_t(sz) * A.src.ky + sy) * A.src.kx;
          const int s0 = gen::wrap_src_index(2 * n, nx, A.src.kx);
          const int s1 = gen::wrap_src_index(2 * n + 1, nx, A.src.kx);
          if (s0 >= 0) re = kr[s0];
          if (s1 >= 0) im = kr[s1];
        }
        v[a * R1 + r] = cmake(re, im);
      }
    }
  }
  rows_fwd_from_regs<M, TW, SP>(A, slab, row0, lane, T, v);
}

template <int M, typename TW, int SP>
__device__ __forceinline__ void rows_fwd_from_regs(const RowArgs& A, cplx* slab, long long row0, int lane,
                                                   const TW& T, cplx* v) {
  typedef Row2Cfg<M> CF;
  constexpr int R1 = CF::R1, RPG = CF::RPG, RS = CF::RS, PAIRS = CF::PAIRS;
  const int q_blk = lane % R1, r_blk = lane / R1;
  // ---- stage 1: radix-R1 ----
#pragma unroll
  for (int a = 0; a < RPG; ++a) {
    Bfly<R1, false>::run(v + a * R1);
#pragma unroll
    for (int q = 0; q < R1; ++q) {
      cplx x = v[a * R1 + q];
      if (q > 0) x = cmul(x, T.tw1(q));
      slab[a * RS + q * 17 + lane] = x;
    }
  }
  __syncwarp();
  // ---- stage 2: one radix-16 block per lane ----
  {
    const cplx* p = slab + r_blk * RS + q_blk * 17;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = p[j];
  }
  Bfly<16, false>::run(v);
  __syncwarp();
  {
    cplx* p = slab + r_blk * RS + q_blk;  // Z[q + R1 q2] in natural order
#pragma unroll
    for (int q2 = 0; q2 < 16; ++q2) p[R1 * q2] = v[q2];
  }
  __syncwarp();
  // ---- real-transform split + full-line stores ----
#pragma unroll
  for (int a = 0; a < RPG; ++a) {
    const cplx* zr = slab + a * RS;
    cplx* orow = A.spec + (row0 + a) * A.nxp;
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int k = lane + 16 * i;
      if (k == 0) {
        const cplx z0 = zr[0];
        const cplx zh = zr[M / 2];  // read before the stores: with SP = 1 the spectrum row IS the slab row
        st_spec<SP>(orow, cmake(z0.x + z0.y, 0.f));
        st_spec<SP>(A.nyq ? A.nyq + (row0 + a) : orow + M, cmake(z0.x - z0.y, 0.f));
        st_spec<SP>(orow + M / 2, cconj(zh));
      } else {
        cplx xk, xm;
        r2c_pair(zr[k], zr[M - k], T.twp(i), xk, xm);
        st_spec<SP>(orow + k, xk);
        st_spec<SP>(orow + (M - k), xm);
      }
    }
  }
  __syncwarp();
}

// inverse transform + pointwise epilogue of rows row0 .. row0+RPG-1 by one 16-lane group
// CHAIN: the real samples the epilogue produces are not (only) stored but forward transformed again and
// written back as the row's spectrum -- the x pass of the NEXT convolution, fused: `integral` never goes to
// HBM at all and the new psi is not read back (saves 3S of the 7S + 18C per (view, iteration)).
template <int M, int EPI, typename TW, bool CHAIN = false, bool EMBED = false, int SP = 0>
__device__ __forceinline__ void rows_inv_group(const RowArgs& A, cplx* slab, long long row0, int lane,
                                               const TW& T) {
  typedef Row2Cfg<M> CF;
  constexpr int R1 = CF::R1, RPG = CF::RPG, RS = CF::RS, PAIRS = CF::PAIRS;
  constexpr int nx = 2 * M;
  const int q_blk = lane % R1, r_blk = lane / R1;
  constexpr int mode = EPI;  // compile time: no branches, no dead operand registers
  // rows the data comes from: the row itself, or (EMBED) the interior row it aliases
  long long srow[RPG];
#pragma unroll
  for (int a = 0; a < RPG; ++a) {
    srow[a] = row0 + a;
    if (EMBED) {
      const int z = int(srow[a] / A.ny), y = int(srow[a] % A.ny);
      int sz = (z - A.oz) % A.lz, sy = (y - A.oy) % A.ly;
      if (sz < 0) sz += A.lz;
      if (sy < 0) sy += A.ly;
      srow[a] = (long long)(sz + A.oz) * A.ny + (sy + A.oy);
    }
  }
  // ---- spectrum loads (full lines) ----
  cplx xs[RPG * PAIRS], xm[RPG * PAIRS];
  cplx xh[RPG];
#pragma unroll
  for (int a = 0; a < RPG; ++a) {
    const cplx* irow = A.spec + srow[a] * A.nxp;
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int k = lane + 16 * i;
      xs[a * PAIRS + i] = ld_spec<SP>(irow + k);
      xm[a * PAIRS + i] = ld_spec<SP>((k == 0 && A.nyq) ? A.nyq + srow[a] : irow + (M - k));  // k = 0 reads X[M]
    }
    if (lane == 0) xh[a] = ld_spec<SP>(irow + M / 2);
  }
  if (SP) __syncwarp();  // the spectrum rows are the slab rows: every lane has read before any lane writes
  // epilogue operands: independent of the transform, fetched now so that their latency
  // overlaps both exchanges (element n = lane + 16 r of each row = real samples 2n, 2n+1)
  float2 oa[16], ob[16];
  if (mode != gen::EPI_STORE) {
    const float* pa = (mode == gen::EPI_QUOTIENT) ? A.ep.view : A.ep.psi;
#pragma unroll
    for (int a = 0; a < RPG; ++a) {
      if (SP && A.op_smem && srow[a] < A.op_smem_rows) {
        const float2* p2 = reinterpret_cast<const float2*>(A.op_smem + srow[a] * nx);
#pragma unroll
        for (int r = 0; r < R1; ++r) oa[a * R1 + r] = p2[lane + 16 * r];
      } else {
        const float2* p2 = reinterpret_cast<const float2*>(pa + real_row<SP>(A, srow[a]) * nx);
#pragma unroll
        for (int r = 0; r < R1; ++r) oa[a * R1 + r] = ld_stream(p2 + lane + 16 * r);
      }
    }
  }
  // ---- inverse split into the slab, natural order ----
#pragma unroll
  for (int a = 0; a < RPG; ++a) {
    cplx* zr = slab + a * RS;
#pragma unroll
    for (int i = 0; i < PAIRS; ++i) {
      const int k = lane + 16 * i;
      if (k == 0) {
        const float x0 = xs[a * PAIRS + i].x, xM = xm[a * PAIRS + i].x;  // imaginary parts ignored (c2r)
        zr[0] = cmake(x0 + xM, x0 - xM);
        zr[M / 2] = cmake(2.f * xh[a].x, -2.f * xh[a].y);
      } else {
        cplx zk, zm;
        c2r_pair(xs[a * PAIRS + i], xm[a * PAIRS + i], T.twp(i), zk, zm);
        zr[k] = zk;
        zr[M - k] = zm;
      }
    }
  }
  __syncwarp();
  if (mode == gen::EPI_UPDATE) {
    // the second operand is fetched once the spectrum registers are free (no spills at 128 registers);
    // its latency still overlaps the radix-16 stage and the second exchange
#pragma unroll
    for (int a = 0; a < RPG; ++a) {
      const float2* p2 = reinterpret_cast<const float2*>(A.ep.weights + real_row<SP>(A, srow[a]) * nx);
#pragma unroll
      for (int r = 0; r < R1; ++r) ob[a * R1 + r] = ld_stream(p2 + lane + 16 * r);
    }
  }
  cplx v[16];
  {
    const cplx* p = slab + r_blk * RS + q_blk;
#pragma unroll
    for (int q2 = 0; q2 < 16; ++q2) v[q2] = p[R1 * q2];
  }
  Bfly<16, true>::run(v);
  __syncwarp();
  {
    cplx* p = slab + r_blk * RS + q_blk * 17;
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = v[j];
  }
  __syncwarp();
  // EMBED: rows outside the box read the psi of the interior row they alias while that row's own CTA updates it,
  // so the new psi goes to a second buffer (A.out)
  float* obase = (mode == gen::EPI_UPDATE) ? ((EMBED && A.out) ? A.out : A.ep.psi) : A.out;
#pragma unroll
  for (int a = 0; a < RPG; ++a) {
#pragma unroll
    for (int q = 0; q < R1; ++q) {
      cplx x = slab[a * RS + q * 17 + lane];
      if (q > 0) x = cmulc(x, T.tw1(q));
      v[a * R1 + q] = x;
    }
    Bfly<R1, true>::run(v + a * R1);
    float2* orow = reinterpret_cast<float2*>(obase + real_row<SP>(A, srow[a]) * nx);
    const bool own_row = !EMBED || srow[a] == row0 + a;  // aliases recompute, only the interior row stores psi
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      float2 val = v[a * R1 + r];  // 1/N lives in K^: no scale here (the host rejects ep.scale != 1)
      if (mode == gen::EPI_QUOTIENT) {
        val.x = quotient(oa[a * R1 + r].x, val.x, A.ep.zero_view_guard);
        val.y = quotient(oa[a * R1 + r].y, val.y, A.ep.zero_view_guard);
      } else if (mode == gen::EPI_UPDATE) {
        val.x = rl_update(oa[a * R1 + r].x, val.x, ob[a * R1 + r].x, A.ep.up);
        val.y = rl_update(oa[a * R1 + r].y, val.y, ob[a * R1 + r].y, A.ep.up);
      }
      if ((!CHAIN || mode == gen::EPI_UPDATE) && own_row) st_stream(orow + lane + 16 * r, val);  // psi is always stored
      if (CHAIN) v[a * R1 + r] = val;
    }
  }
  __syncwarp();
  if (CHAIN && EMBED) {
    // periodic continuation along x inside the row: samples outside [ox, ox + lx) take the interior sample they alias
#pragma unroll
    for (int a = 0; a < RPG; ++a)
#pragma unroll
      for (int r = 0; r < R1; ++r) slab[a * RS + lane + 16 * r] = v[a * R1 + r];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < RPG; ++a) {
      const float* rowf = reinterpret_cast<const float*>(slab + a * RS);
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        const int p0 = 2 * (lane + 16 * r);
        if (p0 < A.ox || p0 >= A.ox + A.lx) {
          int sx = (p0 - A.ox) % A.lx;
          if (sx < 0) sx += A.lx;
          v[a * R1 + r].x = rowf[sx + A.ox];
        }
        if (p0 + 1 < A.ox || p0 + 1 >= A.ox + A.lx) {
          int sx = (p0 + 1 - A.ox) % A.lx;
          if (sx < 0) sx += A.lx;
          v[a * R1 + r].y = rowf[sx + A.ox];
        }
      }
    }
    __syncwarp();
    RowArgs B = A;  // the forward transform writes the OTHER spectrum buffer, at the row's own position
    B.spec = A.spec_out;
    B.nyq = A.nyq_out;
    rows_fwd_from_regs<M, TW, 0>(B, slab, row0, lane, T, v);
  } else if (CHAIN) {
    rows_fwd_from_regs<M, TW, SP>(A, slab, row0, lane, T, v);
  }
}

// CTA shape of the rows kernels (chained and plain): four CTAs of 128 threads per SM instead of two of 256 -- the same 16 warps, but
// finer units for the hardware CTA scheduler (measured: update link 0.246 -> 0.234 ms, quotient link 0.147 -> 0.143)
#ifndef LMVN_LINK_THREADS
#define LMVN_LINK_THREADS 128
#endif
#ifndef LMVN_LINK_BLOCKS
#define LMVN_LINK_BLOCKS 4
#endif
static const int kLinkThreads = LMVN_LINK_THREADS;
#ifndef LMVN_ROWS_FWD_BLOCKS
#define LMVN_ROWS_FWD_BLOCKS 2
#endif
#ifndef LMVN_ROWS_INVQ_BLOCKS
#define LMVN_ROWS_INVQ_BLOCKS 2
#endif
template <int M, bool WRAPPED>
static __global__ void __launch_bounds__(kLinkThreads, LMVN_LINK_BLOCKS) k_rows_fwd2(RowArgs A) {
  typedef Row2Cfg<M> CF;
  constexpr int GROUPS = kLinkThreads / 16;
  LMVN_DYN_SMEM(cplx, sm);
  const int lane = threadIdx.x % 16;
  const int group = threadIdx.x / 16;
  cplx* slab = sm + group * (CF::RPG * CF::RS);
  const long long rows = (long long)A.nz * A.ny;
  RowTw<M> T;
  T.load(A, lane);
  const long long stride = (long long)gridDim.x * GROUPS * CF::RPG;
  for (long long row0 = ((long long)blockIdx.x * GROUPS + group) * CF::RPG; row0 < rows; row0 += stride) {
    if (!WRAPPED && A.prefetch && row0 + stride < rows)  // next iteration's RPG rows = 16 lines
      prefetch_l2(reinterpret_cast<const char*>(A.src.data + (row0 + stride) * (2 * M)) + lane * (CF::RPG * M / 2));
    rows_fwd_group<M, WRAPPED>(A, slab, row0, lane, T);
  }
}

template <int M, int EPI>
static __global__ void __launch_bounds__(kLinkThreads, LMVN_LINK_BLOCKS) k_rows_inv2(RowArgs A) {
  typedef Row2Cfg<M> CF;
  constexpr int GROUPS = kLinkThreads / 16;
  LMVN_DYN_SMEM(cplx, sm);
  __shared__ cplx s_tw[RowTwShared<M>::ENTRIES * 16];
  const int lane = threadIdx.x % 16;
  const int group = threadIdx.x / 16;
  cplx* slab = sm + group * (CF::RPG * CF::RS);
  const long long rows = (long long)A.nz * A.ny;
  RowTwShared<M>::fill(s_tw, A);
  __syncthreads();
  RowTwShared<M> T;
  T.base = s_tw + lane;
  const long long stride = (long long)gridDim.x * GROUPS * CF::RPG;
  for (long long row0 = ((long long)blockIdx.x * GROUPS + group) * CF::RPG; row0 < rows; row0 += stride) {
    if (A.prefetch && row0 + stride < rows) {
      // next iteration: RPG spectrum rows (RPG * nxp * 8 bytes) and RPG operand rows (16 lines each)
      const long long nr = row0 + stride;
      const char* sp = reinterpret_cast<const char*>(A.spec + nr * A.nxp);
      const int spec_bytes = CF::RPG * A.nxp * int(sizeof(cplx));
      for (int b = lane * 128; b < spec_bytes; b += 16 * 128) prefetch_l2(sp + b);
      const int ob = lane * (CF::RPG * M / 2);
      if (EPI == gen::EPI_QUOTIENT) prefetch_l2(reinterpret_cast<const char*>(A.ep.view + nr * (2 * M)) + ob);
      if (EPI == gen::EPI_UPDATE) {
        prefetch_l2(reinterpret_cast<const char*>(A.ep.psi + nr * (2 * M)) + ob);
        prefetch_l2(reinterpret_cast<const char*>(A.ep.weights + nr * (2 * M)) + ob);
      }
    }
    rows_inv_group<M, EPI>(A, slab, row0, lane, T);
  }
}

// inverse x + pointwise + forward x of the next convolution, in place on the spectrum rows
template <int M, int EPI, bool EMBED = false>
static __global__ void __launch_bounds__(kLinkThreads, LMVN_LINK_BLOCKS) k_rows_inv_fwd(RowArgs A) {
  typedef Row2Cfg<M> CF;
  constexpr int GROUPS = kLinkThreads / 16;
  LMVN_DYN_SMEM(cplx, sm);
  __shared__ cplx s_tw[RowTwShared<M>::ENTRIES * 16];
  const int lane = threadIdx.x % 16;
  const int group = threadIdx.x / 16;
  cplx* slab = sm + group * (CF::RPG * CF::RS);
  const long long rows = (long long)A.nz * A.ny;
  RowTwShared<M>::fill(s_tw, A);
  __syncthreads();
  RowTwShared<M> T;
  T.base = s_tw + lane;
  const long long stride = (long long)gridDim.x * GROUPS * CF::RPG;
  for (long long row0 = ((long long)blockIdx.x * GROUPS + group) * CF::RPG; row0 < rows; row0 += stride) {
    if (A.prefetch && row0 + stride < rows) {
      const long long nr = row0 + stride;
      const char* sp = reinterpret_cast<const char*>(A.spec + nr * A.nxp);
      const int spec_bytes = CF::RPG * A.nxp * int(sizeof(cplx));
      for (int b = lane * 128; b < spec_bytes; b += 16 * 128) prefetch_l2(sp + b);
      const int ob = lane * (CF::RPG * M / 2);
      if (EPI == gen::EPI_QUOTIENT) prefetch_l2(reinterpret_cast<const char*>(A.ep.view + nr * (2 * M)) + ob);
      if (EPI == gen::EPI_UPDATE) {
        prefetch_l2(reinterpret_cast<const char*>(A.ep.psi + nr * (2 * M)) + ob);
        prefetch_l2(reinterpret_cast<const char*>(A.ep.weights + nr * (2 * M)) + ob);
      }
    }
    rows_inv_group<M, EPI, RowTwShared<M>, true, EMBED>(A, slab, row0, lane, T);
  }
}

// ------------------------------------------------------------------------------
// x passes for nx = 1024 (M = 512 complex samples per row).  One more radix-2 level
// around the M = 256 scheme, decimation in time, and ONE WARP PER ROW: the half-warp
// h = 0 owns the even complex samples (sub-transform E), h = 1 the odd ones (O).  Lane
// (h, j) loads float2 = complex sample 2 (j + 16 r) + h, so a warp-wide access is 256
// contiguous bytes; each half-warp runs its 256-point sub-transform exactly like the
// M = 256 kernel (radix 16, exchange, radix 16, exchange to natural order, slabs se / so)
// and the last butterfly Z[k] = E[k] + w_512^k O[k], Z[k+256] = E[k] - w_512^k O[k] is
// fused into the real-transform split, 8 (k, 512-k) pairs per lane: the pair needs Z[k]
// and Z[256 + (256-k)] = E[256-k] + conj(w_512^k) O[256-k].  The inverse is the mirror;
// E'[k] (h = 0) / O'[k] (h = 1) for k = j + 16 i are exactly the inputs of the lane's
// first inverse radix-16, so the un-combine costs no extra exchange.
// (Until round 2 a 16-lane group held BOTH sub-transforms of a row, 32 complex values per
// lane: the chained kernel needed 170 registers, 12 warps per SM, and ran at 3.4-4.3 TB/s.
// With 16 values per lane the chained kernels fit 128 registers like the M <= 256 ones;
// the arithmetic per element is unchanged -- results are bit-identical.)
// ------------------------------------------------------------------------------
struct RowWide {
  static const int H = 256, M = 512, NX = 1024;
  static const int RS = 272;                 // 17 * 16: pitch of one half-length slab
  static const int SLAB = 2 * RS;            // E and O (>= 512: also holds Z in natural order)
  static const int GROUPS = kRowThreads / 32;
  static const int ROWS = GROUPS;            // one row per warp and iteration
  static const int SMEM = GROUPS * SLAB * int(sizeof(cplx));
};

__device__ __forceinline__ float4 ld_stream4(const float4* p) {
#ifdef LMVN_EMU
  return *p;
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
#ifdef LMVN_EMU
  *p = v;
#else
  __stcg(p, v);
#endif
}

// v[r] = complex sample 2 m + h of the row, m = j + 16 r (lane = 16 h + j)
__device__ __forceinline__ void rows_fwd_wide_from_regs(const RowArgs& A, cplx* slab, long long row, int lane, cplx* v);

template <bool WRAPPED>
__device__ __forceinline__ void rows_fwd_wide_group(const RowArgs& A, cplx* slab, long long row, int lane) {
  constexpr int nx = RowWide::NX;
  const int h = lane >> 4, j = lane & 15;
  cplx v[16];
  if (!WRAPPED) {
    const float2* in = reinterpret_cast<const float2*>(A.src.data + row * nx);
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = ld_stream(in + 2 * (j + 16 * r) + h);
  } else {
    const int z = int(row / A.ny) + A.z0, y = int(row % A.ny);
    const int sz = gen::wrap_src_index(z, A.nz_wrap, A.src.kz);
    const int sy = gen::wrap_src_index(y, A.ny, A.src.ky);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float f[2] = {0.f, 0.f};
      if (sz >= 0 && sy >= 0) {
        const float* kr = A.src.data + (size_t(sz) * A.src.ky + sy) * A.src.kx;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int sx = gen::wrap_src_index(4 * (j + 16 * r) + 2 * h + c, nx, A.src.kx);
          if (sx >= 0) f[c] = kr[sx];
        }
      }
      v[r] = cmake(f[0], f[1]);
    }
  }
  rows_fwd_wide_from_regs(A, slab, row, lane, v);
}

__device__ __forceinline__ void rows_fwd_wide_from_regs(const RowArgs& A, cplx* slab, long long row, int lane, cplx* v) {
  constexpr int H = RowWide::H, M = RowWide::M, RS = RowWide::RS;
  const int h = lane >> 4, j = lane & 15;
  cplx* sh = slab + h * RS;  // this half-warp's sub-transform
  const cplx* se = slab;
  const cplx* so = slab + RS;
  Bfly<16, false>::run(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    cplx a = v[q];
    if (q > 0) a = cmul(a, __ldg(A.tw_h + j * q));
    sh[q * 17 + j] = a;
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < 16; ++t) v[t] = sh[j * 17 + t];
  Bfly<16, false>::run(v);
  __syncwarp();
#pragma unroll
  for (int q2 = 0; q2 < 16; ++q2) sh[j + 16 * q2] = v[q2];  // E[k] / O[k], k = j + 16 q2, natural order
  __syncwarp();
  cplx* orow = A.spec + row * A.nxp;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = lane + 32 * i;
    if (k == 0) {
      const cplx e0 = se[0], o0 = so[0];
      const cplx z0 = cadd(e0, o0), zh = csub(e0, o0);  // Z[0], Z[256]
      st_stream(orow, cmake(z0.x + z0.y, 0.f));
      st_stream(A.nyq ? A.nyq + row : orow + M, cmake(z0.x - z0.y, 0.f));
      st_stream(orow + H, cconj(zh));
    } else {
      const cplx w5 = __ldg(A.tw_m + k);  // w_512^k
      const cplx zk = cadd(se[k], cmul(so[k], w5));
      const cplx zm = cadd(se[H - k], cmulc(so[H - k], w5));  // Z[512-k]
      cplx xk, xm;
      r2c_pair(zk, zm, __ldg(A.tw_nx + k), xk, xm);
      st_stream(orow + k, xk);
      st_stream(orow + (M - k), xm);
    }
  }
  __syncwarp();
}

template <int EPI, bool CHAIN = false>
__device__ __forceinline__ void rows_inv_wide_group(const RowArgs& A, cplx* slab, long long row, int lane) {
  constexpr int H = RowWide::H, M = RowWide::M, nx = RowWide::NX, RS = RowWide::RS;
  const int h = lane >> 4, j = lane & 15;
  const cplx* irow = A.spec + row * A.nxp;
  // ---- loads + inverse split: Z in natural order ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = lane + 32 * i;
    const cplx xk = ld_stream(irow + k);
    const cplx xm = ld_stream((k == 0 && A.nyq) ? A.nyq + row : irow + (M - k));  // k = 0 reads X[512]
    if (k == 0) {
      const cplx xh = ld_stream(irow + H);
      slab[0] = cmake(xk.x + xm.x, xk.x - xm.x);  // imaginary parts ignored (c2r)
      slab[H] = cmake(2.f * xh.x, -2.f * xh.y);
    } else {
      cplx zk, zm;
      c2r_pair(xk, xm, __ldg(A.tw_nx + k), zk, zm);
      slab[k] = zk;
      slab[M - k] = zm;
    }
  }
  // epilogue operands: element n = 2 (j + 16 r) + h of the row = real samples 2n, 2n+1 (this lane's results); fetched
  // now so that their latency overlaps the transform
  float* obase = (EPI == gen::EPI_UPDATE) ? A.ep.psi : A.out;
  float2* orow = reinterpret_cast<float2*>(obase + row * nx);
  float2 oa[16], ob[16];
  if (EPI != gen::EPI_STORE) {
    const float2* pa = reinterpret_cast<const float2*>((EPI == gen::EPI_QUOTIENT ? A.ep.view : A.ep.psi) + row * nx);
#pragma unroll
    for (int r = 0; r < 16; ++r) oa[r] = ld_stream(pa + 2 * (j + 16 * r) + h);
  }
  __syncwarp();
  // ---- un-combine: E'[k] = Z[k] + Z[k+256] (h = 0), O'[k] = (Z[k] - Z[k+256]) conj(w_512^k) (h = 1), k = j + 16 i ----
  cplx v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int k = j + 16 * i;
    const cplx a = slab[k], b = slab[k + H];
    if (h == 0) {
      v[i] = cadd(a, b);
    } else {
      const cplx d = csub(a, b);
      v[i] = (k == 0) ? d : cmulc(d, __ldg(A.tw_m + k));
    }
  }
  // these are the inputs Z[q_blk + 16 q2] of the lane's first inverse radix-16 (q_blk = j)
  Bfly<16, true>::run(v);
  __syncwarp();
  cplx* sh = slab + h * RS;
#pragma unroll
  for (int t = 0; t < 16; ++t) sh[j * 17 + t] = v[t];
  __syncwarp();
  if (EPI == gen::EPI_UPDATE) {
    const float2* pb = reinterpret_cast<const float2*>(A.ep.weights + row * nx);
#pragma unroll
    for (int r = 0; r < 16; ++r) ob[r] = ld_stream(pb + 2 * (j + 16 * r) + h);
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    cplx a = sh[q * 17 + j];
    if (q > 0) a = cmulc(a, __ldg(A.tw_h + j * q));
    v[q] = a;
  }
  Bfly<16, true>::run(v);
  // v[r] = complex sample 2 m + h, m = j + 16 r = real samples 4 m + 2 h, 4 m + 2 h + 1 (1/N lives in K^: no scale here)
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    float2 val = v[r];
    if (EPI == gen::EPI_QUOTIENT) {
      val.x = quotient(oa[r].x, val.x, A.ep.zero_view_guard);
      val.y = quotient(oa[r].y, val.y, A.ep.zero_view_guard);
    } else if (EPI == gen::EPI_UPDATE) {
      val.x = rl_update(oa[r].x, val.x, ob[r].x, A.ep.up);
      val.y = rl_update(oa[r].y, val.y, ob[r].y, A.ep.up);
    }
    if (!CHAIN || EPI == gen::EPI_UPDATE) st_stream(orow + 2 * (j + 16 * r) + h, val);
    if (CHAIN) v[r] = val;
  }
  __syncwarp();
  if (CHAIN) rows_fwd_wide_from_regs(A, slab, row, lane, v);
}

template <bool WRAPPED>
static __global__ void __launch_bounds__(kRowThreads, 2) k_rows_fwd_wide(RowArgs A) {
  LMVN_DYN_SMEM(cplx, sm);
  const int lane = threadIdx.x % 32;
  const int group = threadIdx.x / 32;
  cplx* slab = sm + group * RowWide::SLAB;
  const long long rows = (long long)A.nz * A.ny;
  const long long stride = (long long)gridDim.x * RowWide::ROWS;
  for (long long row = (long long)blockIdx.x * RowWide::ROWS + group; row < rows; row += stride) {
    if (!WRAPPED && A.prefetch && row + stride < rows)  // next iteration's row: 4 KB = 32 lines
      prefetch_l2(reinterpret_cast<const char*>(A.src.data + (row + stride) * RowWide::NX) + lane * 128);
    rows_fwd_wide_group<WRAPPED>(A, slab, row, lane);
  }
}

// next iteration's spectrum row and operand rows into L2
template <int EPI>
__device__ __forceinline__ void rows_wide_prefetch(const RowArgs& A, long long nr, int lane) {
  const char* sp = reinterpret_cast<const char*>(A.spec + nr * A.nxp);
  for (int b = lane * 128; b < A.nxp * int(sizeof(cplx)); b += 32 * 128) prefetch_l2(sp + b);
  if (EPI != gen::EPI_STORE)
    prefetch_l2(reinterpret_cast<const char*>((EPI == gen::EPI_QUOTIENT ? A.ep.view : A.ep.psi) + nr * RowWide::NX) + lane * 128);
  if (EPI == gen::EPI_UPDATE)
    prefetch_l2(reinterpret_cast<const char*>(A.ep.weights + nr * RowWide::NX) + lane * 128);
}

template <int EPI>
static __global__ void __launch_bounds__(kRowThreads, 2) k_rows_inv_wide(RowArgs A) {
  LMVN_DYN_SMEM(cplx, sm);
  const int lane = threadIdx.x % 32;
  const int group = threadIdx.x / 32;
  cplx* slab = sm + group * RowWide::SLAB;
  const long long rows = (long long)A.nz * A.ny;
  const long long stride = (long long)gridDim.x * RowWide::ROWS;
  for (long long row = (long long)blockIdx.x * RowWide::ROWS + group; row < rows; row += stride) {
    if (A.prefetch && row + stride < rows) rows_wide_prefetch<EPI>(A, row + stride, lane);
    rows_inv_wide_group<EPI>(A, slab, row, lane);
  }
}

// the chained link for nx = 1024: four CTAs of 128 threads per SM like the M <= 256 links
static const int kChainWideThreads = LMVN_LINK_THREADS;
template <int EPI>
static __global__ void __launch_bounds__(kChainWideThreads, LMVN_LINK_BLOCKS) k_rows_inv_fwd_wide(RowArgs A) {
  LMVN_DYN_SMEM(cplx, sm);
  constexpr int GROUPS = kChainWideThreads / 32;
  const int lane = threadIdx.x % 32;
  const int group = threadIdx.x / 32;
  cplx* slab = sm + group * RowWide::SLAB;
  const long long rows = (long long)A.nz * A.ny;
  // one row per warp and CTA (grid = rows / GROUPS), left to the hardware CTA scheduler like the update link of the
  // narrower rows: inside a persistent loop the compiler keeps the lane's 31 twiddles in registers across iterations
  // and the kernel spills
  const long long row = (long long)blockIdx.x * GROUPS + group;
  if (row < rows) rows_inv_wide_group<EPI, true>(A, slab, row, lane);
}

}  // namespace fast
}  // namespace lmvn
