// fft_x3.cuh -- the two-pass ("three HBM round trips per convolution") schedule of the power-of-two fast path.
//
// The five-pass schedule of fft_fast.cuh moves the spectrum between HBM and the SMs five times per convolution
// (x | y | z * K^ * z^-1 | y^-1 | x^-1; four launches in the chained loop).  Every one of those passes already streams
// at 5.3 - 5.9 TB/s, so the only way to go faster is to move fewer bytes.  A whole
// x-y plane (512 x 129 complex = 516 KB for BASELINE config 3) does not fit the 227 KB of one SM, and a 4-CTA cluster
// that does fit it leaves 16 of the 148 SMs idle and pays the x <-> y transposition through DSMEM at ~20 B/clk/SM.
// This file splits the y axis by DIGITS instead:   ny = 128 * R2   (R2 = 4 for ny = 512, 2 for ny = 256)
//
//   pass A  "plane tile"  k_plane<MODE>   one CTA = the 128 rows y = R2 * i + n2 of one z plane (a 128 KB spectrum tile in
//           shared memory):  [128-point inverse along i] -> [x inverse] -> quotient / RL update -> [x forward] ->
//           [128-point forward along i].  Everything the chained rows kernel of fft_fast.cuh did, plus seven of the
//           nine (eight) radix-2 levels of the y transform, without touching HBM in between.
//   pass B  "z middle"    k_zmid5 / k_zmid   the merged z pass of fft_fast.cuh (forward z, * K^, inverse z on 64 KB tiles)
//           with the remaining radix-R2 level of y as a register stage behind the first z stage and, mirrored, behind
//           the middle stage: a tile holds the R2 samples (n2 = 0 .. R2-1) of its short y transforms, as pieces of
//           KXC = 16 / R2 adjacent kx columns.
//
// HBM traffic per (view, iteration): (2C + S) + 3C + (2C + 3S) + 3C = 4S + 10C -- below the contract's 7S + 10C
// (SURVEY.md section 8d), against 4S + 18C of the chained five-pass loop.
//
// STATUS: parity-green (emulator, GPU vs oracle, full config-3 size) and OPT-IN (LMVN_X3=1): measured 1.18 ms per
// (view, iteration) on config 3 against 1.03 ms of the five-pass loop.  The plane pass wins what it should (0.28 / 0.35 ms
// against 0.35 / 0.45 ms for y inverse + link + y forward), the z middle loses more (0.29 against 0.16 ms): its rows are
// 32-byte pieces, and the L1/LSU pipe spends a whole wavefront on each.  DESIGN.md section 3.4 has the numbers, the ncu
// evidence and the variants that were tried (shuffles, TMA-staged operands, L2 look-ahead, 128-bit y stages).
//
// Spectrum layout between the passes ("A layout"):  A[n2][z][p][kx], n2 < R2, p < 129, kx < 128 (complex64).
//   p < 128: position p of the 128-point transform along i (16 x 8 decimation in frequency: p = 8 q1 + q2 holds
//            frequency k1 = q1 + 16 q2);  p = 128: the Nyquist column kx = nx/2 of the tile, TRANSPOSED (entry kx
//            of that row = position kx), so that the array is uniform: pass B sees 129 x 128 identical columns per (n2, z)
//            and needs no special case, pass A reads / writes one contiguous 129 KB tile.
//   after the forward radix-R2 level slot n2 holds ky = k1 + 128 * n2; z is in the digit-reversed order
//   of fft_fast.cuh's merged pass.  K^ is precomputed by the five-pass kernels and permuted into this layout once
//   per view (k_khat_to_a), so the spectrum product never needs a natural order.
#pragma once
#include "fft_fast.cuh"

namespace lmvn {
namespace x3 {

using namespace fast;

static const int kRows = 128;      // rows of a plane tile = length of the in-tile y transform
static const int kM = 128;         // complex samples per row (nx = 256)
static const int kPitch = 136;     // shared-memory row pitch = Row2Cfg<128>::RS: the x transforms run in place on the tile rows
static const int kPlaneThreads = 512;
static const int kTileElems = (kRows + 1) * kM;  // 129 x 128 complex per (n2, z) in HBM
static const size_t kPlaneSmem = size_t(kRows) * kPitch * sizeof(cplx);
// staged operand rows (TMA variant of the chained plane pass): the rows of the FIRST of the two row-loop iterations
// (64 rows x nx floats = 64 KB; all 128 would not fit next to the 136 KB tile) behind the tile, + the mbarrier.  The
// rows of the second iteration are pulled into L2 at the same time and read with ordinary loads.
static const int kStagedRows = kRows / 2;
static const size_t kPlaneOpBytes = size_t(kStagedRows) * 2 * kM * sizeof(float);
static const size_t kPlaneSmemStaged = kPlaneSmem + kPlaneOpBytes + 16;

// ---- bulk asynchronous copies (TMA engine, SASS: UBLKCP) completing on an mbarrier --------------------------------
#if !defined(LMVN_EMU)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_global, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LMVN_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LMVN_MBAR_DONE;\n"
      "bra LMVN_MBAR_WAIT;\n"
      "LMVN_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#endif

enum PlaneMode { PM_CHAIN = 0, PM_BEGIN = 1, PM_END = 2 };

struct PlaneArgs {
  RowArgs row;        // spec = nullptr (set per CTA), ep, src (PM_BEGIN), out (PM_END), tw_m, tw_nx
  cplx* a;            // A layout, transformed in place
  int nz, ny, r2;     // r2 = ny / 128
  const cplx* tw_y;   // [j < 8][q < 16] = w_128^{jq}: stage table of the 128-point transform (16 x 8)
  int prefetch;       // > 0: once its own tile is on chip, a CTA pulls the tile and the operand rows of CTA
                      // blockIdx + prefetch (the one that takes an SM next) from HBM into L2
  int stage_ops;      // chained pass: the rows of the first pointwise operand (view / psi) come in by bulk async copies
                      // (TMA engine, mbarrier completion) issued at kernel entry, and are read from shared memory
};

// L2 prefetch of everything CTA `b` will read: its spectrum tile and its rows of the pointwise operands
template <int MODE, int EPI>
__device__ __forceinline__ void plane_prefetch(const PlaneArgs& P, unsigned b) {
  if (b >= unsigned(P.nz * P.r2)) return;
  const int n2 = b % P.r2, z = b / P.r2;
  if (MODE != PM_BEGIN) {
    const char* t = reinterpret_cast<const char*>(P.a + (size_t(n2) * P.nz + z) * kTileElems);
    for (int i = threadIdx.x; i < int(kTileElems * sizeof(cplx) / 128); i += kPlaneThreads) prefetch_l2(t + size_t(i) * 128);
  }
  // real-space rows y = r2 * i + n2 of plane z: 2 * kM floats = 8 lines each
  const size_t row0 = size_t(z) * P.ny + n2;
  const float* ops[2] = {nullptr, nullptr};
  if (MODE == PM_BEGIN) ops[0] = P.row.src.data;
  else if (EPI == gen::EPI_QUOTIENT) ops[0] = P.row.ep.view;
  else if (EPI == gen::EPI_UPDATE) { ops[0] = P.row.ep.psi; ops[1] = P.row.ep.weights; }
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    if (!ops[o]) continue;
    for (int i = threadIdx.x; i < kRows * 8; i += kPlaneThreads) {
      const size_t row = row0 + size_t(i / 8) * P.r2;
      prefetch_l2(reinterpret_cast<const char*>(ops[o] + row * (2 * kM)) + (i % 8) * 128);
    }
  }
}

// one radix-R butterfly of span L of the 128-point transform along the tile rows, for one column.
// sm: the column in shared memory (row pitch kPitch); g: the column in global memory (row stride g_rs).
template <int R, int L, bool INV, int SRC, int DST>
__device__ __forceinline__ void plane_bfly(cplx* sm, cplx* g, int g_rs, const cplx* __restrict__ tws, int bf) {
  constexpr int Mq = L / R;
  const int j = bf % Mq;
  const int row0 = (bf / Mq) * L + j;
  cplx v[R];
  if (SRC == W_SMEM) {
    const cplx* p = sm + row0 * kPitch;
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = p[r * Mq * kPitch];
  } else {
    const cplx* gp = g + row0 * g_rs;
    const int step = Mq * g_rs;
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = ld_stream(gp + r * step);
  }
  cplx t[R];
  if (Mq > 1) load_twiddles<R>(t, tws + j * R);
  if (INV && Mq > 1) {
#pragma unroll
    for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], t[q]);
  }
  Bfly<R, INV>::run(v);
  if (!INV && Mq > 1) {
#pragma unroll
    for (int q = 1; q < R; ++q) v[q] = cmul(v[q], t[q]);
  }
  if (DST == W_SMEM) {
    cplx* p = sm + row0 * kPitch;
#pragma unroll
    for (int q = 0; q < R; ++q) p[q * Mq * kPitch] = v[q];
  } else {
    cplx* gp = g + row0 * g_rs;
    const int step = Mq * g_rs;
#pragma unroll
    for (int q = 0; q < R; ++q) st_stream(gp + q * step, v[q]);
  }
}

// the same butterfly on TWO adjacent columns at once: every access is 128 bits wide (half the load/store instructions
// of the stage, the twiddles are shared).  sm / g point at the even column of the pair (16-byte aligned).
template <int R, int L, bool INV, int SRC, int DST>
__device__ __forceinline__ void plane_bfly2(cplx* sm, cplx* g, int g_rs, const cplx* __restrict__ tws, int bf) {
  constexpr int Mq = L / R;
  const int j = bf % Mq;
  const int row0 = (bf / Mq) * L + j;
  cplx v0[R], v1[R];
  if (SRC == W_SMEM) {
    const float4* p = reinterpret_cast<const float4*>(sm + row0 * kPitch);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 f = p[r * Mq * (kPitch / 2)];
      v0[r] = cmake(f.x, f.y);
      v1[r] = cmake(f.z, f.w);
    }
  } else {
    const float4* gp = reinterpret_cast<const float4*>(g + row0 * g_rs);
    const int step = Mq * g_rs / 2;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 f = ld_stream4(gp + r * step);
      v0[r] = cmake(f.x, f.y);
      v1[r] = cmake(f.z, f.w);
    }
  }
  cplx t[R];
  if (Mq > 1) load_twiddles<R>(t, tws + j * R);
  if (INV && Mq > 1) {
#pragma unroll
    for (int q = 1; q < R; ++q) { v0[q] = cmulc(v0[q], t[q]); v1[q] = cmulc(v1[q], t[q]); }
  }
  Bfly<R, INV>::run(v0);
  Bfly<R, INV>::run(v1);
  if (!INV && Mq > 1) {
#pragma unroll
    for (int q = 1; q < R; ++q) { v0[q] = cmul(v0[q], t[q]); v1[q] = cmul(v1[q], t[q]); }
  }
  if (DST == W_SMEM) {
    float4* p = reinterpret_cast<float4*>(sm + row0 * kPitch);
#pragma unroll
    for (int q = 0; q < R; ++q) p[q * Mq * (kPitch / 2)] = make_float4(v0[q].x, v0[q].y, v1[q].x, v1[q].y);
  } else {
    float4* gp = reinterpret_cast<float4*>(g + row0 * g_rs);
    const int step = Mq * g_rs / 2;
#pragma unroll
    for (int q = 0; q < R; ++q) st_stream4(gp + q * step, make_float4(v0[q].x, v0[q].y, v1[q].x, v1[q].y));
  }
}

// one stage over the whole tile: 64 pairs of ordinary columns (thread = pair tid % 64, eight row groups; 128-bit
// accesses) and the Nyquist column (tile column 128 <-> row 128 of the tile in HBM, stride 1), which the first threads
// take on top.  Default: one column per thread, 64-bit accesses; LMVN_X3_WIDE_Y (build knob): column pairs, 128-bit
// accesses -- half the load/store instructions of the y stages, measured neutral (quotient link 0.285 -> 0.281 ms,
// update link 0.332 -> 0.349 ms, profiles/r02_x3_v6_probe.json): the pass is not bound by instruction issue.
template <int R, int L, bool INV, int SRC, int DST>
__device__ __forceinline__ void plane_stage(cplx* tile, cplx* gt, const cplx* __restrict__ tws) {
  constexpr int NB = kRows / R;                 // butterflies per column
#ifndef LMVN_X3_WIDE_Y
  constexpr int RG = kPlaneThreads / kM;        // 4 row groups
  const int c = threadIdx.x % kM, rg = threadIdx.x / kM;
#pragma unroll
  for (int i = 0; i < NB / RG; ++i) plane_bfly<R, L, INV, SRC, DST>(tile + c, gt + c, kM, tws, rg + i * RG);
#else
  constexpr int RG = kPlaneThreads / (kM / 2);  // 8 row groups
  static_assert(NB % RG == 0, "whole butterflies per thread");
  const int cp = threadIdx.x % (kM / 2), rg = threadIdx.x / (kM / 2);
#pragma unroll
  for (int i = 0; i < NB / RG; ++i) plane_bfly2<R, L, INV, SRC, DST>(tile + 2 * cp, gt + 2 * cp, kM, tws, rg + i * RG);
#endif
  if (threadIdx.x < NB) plane_bfly<R, L, INV, SRC, DST>(tile + kM, gt + kRows * kM, 1, tws, threadIdx.x);
}

// pass A.  grid = nz * r2 CTAs (one tile each), 512 threads, kPlaneSmem bytes of dynamic shared memory.
template <int MODE, int EPI>
static __global__ void __launch_bounds__(kPlaneThreads, 1) k_plane(PlaneArgs P) {
  typedef Row2Cfg<kM> CF;
  static_assert(CF::RS == kPitch, "the x transforms run in place on the tile rows");
  LMVN_DYN_SMEM(cplx, tile);
  __shared__ cplx s_tw[RowTwShared<kM>::ENTRIES * 16];
  const int n2 = blockIdx.x % P.r2;
  const int z = blockIdx.x / P.r2;
  cplx* gt = P.a + (size_t(n2) * P.nz + z) * kTileElems;
  RowArgs A = P.row;
  A.spec = tile;
  A.nxp = kPitch;
  A.nyq = nullptr;  // X[nx/2] of tile row i lives at tile[i][128]
  A.rr_base = (long long)z * P.ny + n2;
  A.rr_skip = P.r2 - 1;
  RowTwShared<kM>::fill(s_tw, A);
  // TMA-staged operand rows: 128 bulk copies of one row (nx floats = 1 KB) each, issued before the tile is even loaded,
  // completing on one mbarrier while the y stages run; the rows phase reads them from shared memory
  float* opbuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(tile) + kPlaneSmem);
  void* opbar = reinterpret_cast<unsigned char*>(opbuf) + kPlaneOpBytes;
  const bool staged = (MODE == PM_CHAIN) && P.stage_ops;
  if (staged) {
    const float* src = (EPI == gen::EPI_QUOTIENT) ? A.ep.view : A.ep.psi;
#ifndef LMVN_EMU
    if (threadIdx.x == 0) mbar_init(opbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(opbar, unsigned(kPlaneOpBytes));
    if (threadIdx.x < kStagedRows)
      bulk_g2s(opbuf + threadIdx.x * (2 * kM), src + (A.rr_base + (long long)threadIdx.x * P.r2) * (2 * kM), 2 * kM * sizeof(float), opbar);
    else if (threadIdx.x < kRows) {
      const char* row = reinterpret_cast<const char*>(src + (A.rr_base + (long long)threadIdx.x * P.r2) * (2 * kM));
#pragma unroll
      for (int l = 0; l < 8; ++l) prefetch_l2(row + l * 128);
    }
#else
    for (int i = threadIdx.x; i < kStagedRows * 2 * kM; i += kPlaneThreads)
      opbuf[i] = src[(A.rr_base + (long long)(i / (2 * kM)) * P.r2) * (2 * kM) + i % (2 * kM)];
#endif
    A.op_smem = opbuf;
    A.op_smem_rows = kStagedRows;
  }
  if (MODE != PM_BEGIN) {
    // inverse along the tile rows: radix 8 (span 8) straight from HBM, radix 16 (span 128)
    plane_stage<8, 8, true, W_GLOBAL, W_SMEM>(tile, gt, nullptr);
    __syncthreads();
    if (P.prefetch > 0) plane_prefetch<MODE, EPI>(P, blockIdx.x + unsigned(P.prefetch));
    plane_stage<16, 128, true, W_SMEM, W_SMEM>(tile, gt, P.tw_y);
  } else if (P.prefetch > 0) {
    plane_prefetch<MODE, EPI>(P, blockIdx.x + unsigned(P.prefetch));
  }
  __syncthreads();
#ifndef LMVN_EMU
  if (staged) mbar_wait(opbar, 0);
#endif
  {
    RowTwShared<kM> T;
    const int lane = threadIdx.x % 16, group = threadIdx.x / 16;
    T.base = s_tw + lane;
    constexpr int GROUPS = kPlaneThreads / 16;
#pragma unroll 1
    for (int row0 = group * CF::RPG; row0 < kRows; row0 += GROUPS * CF::RPG) {
      cplx* slab = tile + row0 * kPitch;
      if (MODE == PM_BEGIN) rows_fwd_group<kM, false, RowTwShared<kM>, 1>(A, slab, row0, lane, T);
      else rows_inv_group<kM, EPI, RowTwShared<kM>, MODE == PM_CHAIN, false, 1>(A, slab, row0, lane, T);
    }
  }
  if (MODE == PM_END) return;
  __syncthreads();
  plane_stage<16, 128, false, W_SMEM, W_SMEM>(tile, gt, P.tw_y);
  __syncthreads();
  plane_stage<8, 8, false, W_SMEM, W_GLOBAL>(tile, gt, nullptr);
}

// ------------------------------------------------------------------------------
// pass B: forward z, * K^, inverse z on tiles of NZ rows x COLS columns; a column = (n2, kx), COLS = R2 * KXC.
// The radix-R2 level of y needs the R2 samples (n2 = 0 .. R2-1) of one (z, kx) in one thread, the z stages need NZ/R
// rows of one column in one thread, so the level is its own register stage, in place on the shared tile (128-bit
// accesses: two adjacent kx of R2 slots), between the first z stage and the spectrum product, and mirrored after it:
//   global -> radix R1 (z) -> tile | twiddle w_ny^{n2 k1}, radix R2 (y) | radix R2z, * K^, radix R2z^-1 (z) |
//   radix R2^-1 (y), conjugate twiddle | radix R1^-1 (z) -> global
// Measured alternatives, config 3, five-pass z pass = 0.165 ms:
//   * the level across lanes with shuffles right after the loads (8 SHFL per value): 0.367 ms
//     (profiles/r02_x3_v1_shuffle_zmid_probe.json);
//   * the level as the FIRST stage (128-bit global loads -> radix R2 -> tile, then a radix-32 z stage fed from shared
//     memory): 0.393 ms -- that z stage keeps 64 data + 62 twiddle registers live and spills 340 bytes per thread
//     (gpurun_out/r02_x3_probe_2.json -> profiles/r02_x3_v2_probe.json).
// These kernels are bound by the load/store/shuffle pipe of the SM, not by HBM.
// ------------------------------------------------------------------------------
struct ZmidArgs {
  cplx* a;
  const cplx* khat;   // A layout
  int nz, r2;
  const cplx* tw1;    // stage tables of the z plan (fft_fast.cuh)
  const cplx* tw2;
  const cplx* tw_ny;  // w_ny^m, m < ny: the twiddles between the 128-point level and the radix-R2 level of y
};

// tile width of pass B: the strided passes' width, at most 32 columns
template <int NZ> struct ZmidCols { static const int V = Cols<NZ>::V < 32 ? Cols<NZ>::V : 32; };
// shared-memory row pitch: COLS + KXC complex.  The 128-bit accesses of the y level are served per quarter-warp (8 lanes =
// 8 / PAIRS consecutive rows x PAIRS column pairs per slot n2); with this pitch consecutive rows start PAIRS * 16 bytes
// apart (mod 128), so the eight 16-byte accesses of a quarter-warp fall into distinct banks.  (COLS + 2 looked fine on
// paper for whole warps and cost 2-way conflicts on every access of the level: 8.7 M of 25.7 M shared wavefronts of the
// pass, profiles/r02_ncu_x3_v3.md.)
template <int NZ, int R2> struct ZmidPitch { static const int V = ZmidCols<NZ>::V + ZmidCols<NZ>::V / R2; };

// the radix-R2 level of y, in place on the shared tile: item = (z, column pair), R2 values of 16 bytes each
template <int NZ, int R2, bool INV>
__device__ __forceinline__ void zmid_y_level(cplx* smem, const cplx* __restrict__ tw_ny) {
  constexpr int COLS = ZmidCols<NZ>::V, PITCH = ZmidPitch<NZ, R2>::V;
  constexpr int KXC = COLS / R2, PAIRS = KXC / 2;
  constexpr int THREADS = Threads<NZ>::V;
  constexpr int ITEMS = NZ * PAIRS;
  static_assert(KXC >= 2 && KXC % 2 == 0, "128-bit accesses");
  static_assert(ITEMS % THREADS == 0, "whole items per thread");
  constexpr int IPT = ITEMS / THREADS;
  const int pair = threadIdx.x % PAIRS;
  // (recomputed here and not passed in: nothing of this stage stays live through the z stages, which have no register to spare)
  const int p = blockIdx.x / (kM / KXC);
  const int kx0 = (blockIdx.x % (kM / KXC)) * KXC;
  // position along the 128-point level: the tile row, or (Nyquist row) the column; 16 x 8 decimation in frequency:
  // position 8 q1 + q2 holds frequency k1 = q1 + 16 q2.  Twiddle between the levels: w_ny^{n2 k1}.
  cplx w[2][R2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int pos = (p < kRows) ? p : (kx0 + 2 * pair + e);
    const int k1 = (pos >> 3) + ((pos & 7) << 4);
#pragma unroll
    for (int n = 1; n < R2; ++n) w[e][n] = __ldg(tw_ny + n * k1);
  }
#pragma unroll(IPT >= 2 ? 2 : 1)
  for (int i = 0; i < IPT; ++i) {
    const int item = threadIdx.x + i * THREADS;
    const int z = item / PAIRS;
    float4* sp = reinterpret_cast<float4*>(smem + z * PITCH + 2 * pair);
    cplx v0[R2], v1[R2];
#pragma unroll
    for (int n = 0; n < R2; ++n) {
      const float4 f = sp[n * (KXC / 2)];
      v0[n] = cmake(f.x, f.y);
      v1[n] = cmake(f.z, f.w);
    }
    if (!INV) {
#pragma unroll
      for (int n = 1; n < R2; ++n) {
        v0[n] = cmul(v0[n], w[0][n]);
        v1[n] = cmul(v1[n], w[1][n]);
      }
    }
    Bfly<R2, INV>::run(v0);
    Bfly<R2, INV>::run(v1);
    if (INV) {
#pragma unroll
      for (int n = 1; n < R2; ++n) {
        v0[n] = cmulc(v0[n], w[0][n]);
        v1[n] = cmulc(v1[n], w[1][n]);
      }
    }
#pragma unroll
    for (int n = 0; n < R2; ++n) sp[n * (KXC / 2)] = make_float4(v0[n].x, v0[n].y, v1[n].x, v1[n].y);
  }
}

template <int NZ, int R2>
static __global__ void __launch_bounds__(Threads<NZ>::V, StridedBlocks<NZ, SM_FWD_MUL_INV>::V) k_zmid(ZmidArgs Z) {
  typedef Radix<NZ> RX;
  static_assert(RX::S == 2, "two-stage z plans only");
  constexpr int COLS = ZmidCols<NZ>::V, PITCH = ZmidPitch<NZ, R2>::V;
  constexpr int KXC = COLS / R2;     // adjacent kx columns per n2: pieces of KXC * 8 bytes
  constexpr int R1 = RX::R1, RR2 = RX::R2;
  constexpr int U = LMVN_ZMUL_UNROLL;
  LMVN_DYN_SMEM(cplx, smem);  // [NZ][PITCH]
  constexpr int CHUNKS = kM / KXC;
  const int c = threadIdx.x % COLS;
  cplx* sm = smem + c;
  const int rs = kTileElems;  // between consecutive z
  // this thread's column (n2, kx) of tile (p, kx chunk) -- recomputed where it is needed instead of kept in registers
  auto column = [&]() -> long long {
    const int p = blockIdx.x / CHUNKS, kx0 = (blockIdx.x % CHUNKS) * KXC;
    return (long long)(c / KXC) * Z.nz * kTileElems + (long long)p * kM + kx0 + (c % KXC);
  };
  typedef Middle<NZ, RR2, COLS, PITCH> MID;
  typename MID::K kk;
  // the y level and the z stages commute (different axes; the twiddle between the y levels does not depend on z), so
  // the level sits between the first z stage and the middle: every register stage keeps the footprint it has in the
  // five-pass kernel (a radix-32 stage fed from shared memory instead of HBM spills at 128 registers)
  strided_stage<NZ, R1, NZ, COLS, false, W_GLOBAL, W_SMEM, U, PITCH>(sm, Z.a + column(), rs, Z.tw1, 1.f);
  __syncthreads();
  zmid_y_level<NZ, R2, false>(smem, Z.tw_ny);
  MID::load(kk, Z.khat + column(), rs);  // K^ straight from HBM, issued before the barrier that precedes its use
  __syncthreads();
  MID::run(sm, kk);
  __syncthreads();
  zmid_y_level<NZ, R2, true>(smem, Z.tw_ny);
  __syncthreads();
  strided_stage<NZ, R1, NZ, COLS, true, W_SMEM, W_GLOBAL, U, PITCH>(sm, Z.a + column(), rs, Z.tw1, 1.f);
}

// ------------------------------------------------------------------------------
// pass B, version 5 (16-column tiles: nz = 256, 512): the same stages as k_zmid, but the y level is WARP LOCAL.
// After the first z stage a half-warp (16 lanes = the 16 columns of one row group rg) has written exactly the rows
// rg + 16 q of all 16 columns; the y level needs, per row, the R2 slots of a column pair -- all inside that half-warp's
// rows.  Likewise after the middle stage (rows (rg + 16 i) 16 + r).  So the level runs right behind each of the two
// stages after a __syncwarp(), and the pass is back to the TWO block-wide barriers of the five-pass z pass (k_zmid has
// four: 27 % + 9 % of its stall samples sat in the two short phases the extra barriers fence off, profiles/r02_ncu_x3_v3.md).
// Measured: 0.289 ms against 0.290 ms of k_zmid (profiles/r02_x3_v6_probe.json) -- the barriers were not the limiter either.
// What is: the 32-byte pieces.  A warp-wide 64-bit access of this pass touches EIGHT 128-byte lines (4 slots x 2 rows) where
// the five-pass z pass touches two, and the L1/LSU pipe handles one line per cycle whatever part of it is used: 768
// instead of 192 wavefronts per warp and tile for the tile load, the K^ load and the tile store, ~1300 against ~450 in
// total -- and 0.29 against 0.16 ms.  Wider pieces need wider tiles (KXC = 16: 256 KB of shared memory) or fewer slots
// (R2 = 2 for ny = 512: 256-row plane tiles, 272 KB).  The plane tile and the z tile cannot both fit an SM.
// The rows rg + 16 q of a half-warp all start at the same bank (16 rows x pitch x 8 bytes is a multiple of 128 for every
// pitch), so the tile is swizzled: the 16-byte unit u of row z is stored at unit u ^ (((z >> 4) & (8/PAIRS - 1)) * PAIRS).
// A quarter-warp of the level (8/PAIRS rows x PAIRS units per slot) then hits eight distinct units; the 64-bit accesses of
// the z stages (one whole row per half-warp) are conflict free under any permutation of a row's units.
// ------------------------------------------------------------------------------
struct FwdTag { static const bool value = false; };
struct InvTag { static const bool value = true; };

template <int NZ, int R2>
static __global__ void __launch_bounds__(256, 2) k_zmid5(ZmidArgs Z) {
  typedef Radix<NZ> RX;
  static_assert(RX::S == 2 && Cols<NZ>::V == 16 && NZ / RX::R1 == 16, "16-column tiles, 16 row groups");
  constexpr int COLS = 16, KXC = COLS / R2, PAIRS = KXC / 2, PITCH = COLS + KXC;
  constexpr int R1 = RX::R1, RM = RX::R2;       // z plan: R1 x RM (32 x 16, 16 x 16)
  constexpr int PT = (NZ / RM) / 16;            // middle butterflies per thread
  constexpr int SWZ_MASK = 8 / PAIRS - 1;
  LMVN_DYN_SMEM(cplx, smem);                    // [NZ][PITCH], swizzled
  constexpr int CHUNKS = kM / KXC;
  const int c = threadIdx.x % COLS, rg = threadIdx.x / COLS;
  const int rs = kTileElems;
  auto column = [&]() -> long long {
    const int p = blockIdx.x / CHUNKS, kx0 = (blockIdx.x % CHUNKS) * KXC;
    return (long long)(c / KXC) * Z.nz * kTileElems + (long long)p * kM + kx0 + (c % KXC);
  };
  // this lane's twiddles of the y level (one per slot and element of its column pair)
  auto y_twiddles = [&](cplx (&w)[2][R2], int pair) {
    const int p = blockIdx.x / CHUNKS, kx0 = (blockIdx.x % CHUNKS) * KXC;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int pos = (p < kRows) ? p : (kx0 + 2 * pair + e);
      const int k1 = (pos >> 3) + ((pos & 7) << 4);
#pragma unroll
      for (int n = 1; n < R2; ++n) w[e][n] = __ldg(Z.tw_ny + n * k1);
    }
  };
  // the y level on one (row, column pair) item
  auto y_item = [&](int row, int pair, int swz, const cplx (&w)[2][R2], auto inv_tag) {
    constexpr bool INV = decltype(inv_tag)::value;
    float4* rowp = reinterpret_cast<float4*>(smem + row * PITCH);
    cplx v0[R2], v1[R2];
#pragma unroll
    for (int n = 0; n < R2; ++n) {
      const float4 f = rowp[(n * PAIRS + pair) ^ swz];
      v0[n] = cmake(f.x, f.y);
      v1[n] = cmake(f.z, f.w);
    }
    if (!INV) {
#pragma unroll
      for (int n = 1; n < R2; ++n) { v0[n] = cmul(v0[n], w[0][n]); v1[n] = cmul(v1[n], w[1][n]); }
    }
    Bfly<R2, INV>::run(v0);
    Bfly<R2, INV>::run(v1);
    if (INV) {
#pragma unroll
      for (int n = 1; n < R2; ++n) { v0[n] = cmulc(v0[n], w[0][n]); v1[n] = cmulc(v1[n], w[1][n]); }
    }
#pragma unroll
    for (int n = 0; n < R2; ++n) rowp[(n * PAIRS + pair) ^ swz] = make_float4(v0[n].x, v0[n].y, v1[n].x, v1[n].y);
  };

  // ---- z stage 1: radix R1 over rows rg + 16 r, straight from HBM ----
  {
    cplx v[R1];
    const cplx* gp = Z.a + column() + (long long)rg * rs;
    const int step = 16 * rs;
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      v[r] = ld_stream(gp);
      gp += step;
      LMVN_KEEP_PTR(gp);
    }
    cplx t[R1];
    load_twiddles<R1>(t, Z.tw1 + rg * R1);
    Bfly<R1, false>::run(v);
#pragma unroll
    for (int q = 1; q < R1; ++q) v[q] = cmul(v[q], t[q]);
#pragma unroll
    for (int q = 0; q < R1; ++q)
      smem[(rg + 16 * q) * PITCH + ((((c >> 1) ^ ((q & SWZ_MASK) * PAIRS)) << 1) | (c & 1))] = v[q];
  }
  __syncwarp();
  // ---- y level forward on the rows this half-warp has just written ----
  {
    constexpr int IPT = R1 * PAIRS / 16;
    const int pair = c % PAIRS;
    cplx w[2][R2];
    y_twiddles(w, pair);
#pragma unroll(IPT >= 2 ? 2 : 1)
    for (int i = 0; i < IPT; ++i) {
      const int q = (c + 16 * i) / PAIRS;
      y_item(rg + 16 * q, pair, (q & SWZ_MASK) * PAIRS, w, FwdTag());
    }
  }
  // ---- middle: radix RM, * K^, radix RM^-1 on RM consecutive rows ----
  cplx kk[PT][RM];
  {
    const cplx* gk = Z.khat + column();
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      const cplx* gp = gk + (long long)((rg + 16 * i) * RM) * rs;
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        kk[i][r] = ld_stream(gp);
        gp += rs;
        LMVN_KEEP_PTR(gp);
      }
    }
  }
  __syncthreads();
  const int swz_m = (rg & SWZ_MASK) * PAIRS;  // rows (rg + 16 i) RM + r: (row >> 4) = rg + 16 i
  {
    const int cphys = (((c >> 1) ^ swz_m) << 1) | (c & 1);
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      cplx* p = smem + (rg + 16 * i) * RM * PITCH + cphys;
      cplx v[RM];
#pragma unroll
      for (int r = 0; r < RM; ++r) v[r] = p[r * PITCH];
      Bfly<RM, false>::run(v);
#pragma unroll
      for (int r = 0; r < RM; ++r) v[r] = cmul(v[r], kk[i][r]);
      Bfly<RM, true>::run(v);
#pragma unroll
      for (int r = 0; r < RM; ++r) p[r * PITCH] = v[r];
    }
  }
  __syncwarp();
  // ---- y level inverse on the rows this half-warp has just written ----
  {
    constexpr int IPT = PT * RM * PAIRS / 16;
    const int pair = c % PAIRS;
    cplx w[2][R2];
    y_twiddles(w, pair);
#pragma unroll(IPT >= 2 ? 2 : 1)
    for (int j = 0; j < IPT; ++j) {
      const int rr = (c + 16 * j) / PAIRS;  // 0 .. PT * RM - 1
      y_item((rg + 16 * (rr / RM)) * RM + (rr % RM), pair, swz_m, w, InvTag());
    }
  }
  __syncthreads();
  // ---- z stage 1 inverse: rows rg + 16 q -> HBM ----
  {
    cplx v[R1];
#pragma unroll
    for (int q = 0; q < R1; ++q)
      v[q] = smem[(rg + 16 * q) * PITCH + ((((c >> 1) ^ ((q & SWZ_MASK) * PAIRS)) << 1) | (c & 1))];
    cplx t[R1];
    load_twiddles<R1>(t, Z.tw1 + rg * R1);
#pragma unroll
    for (int q = 1; q < R1; ++q) v[q] = cmulc(v[q], t[q]);
    Bfly<R1, true>::run(v);
    cplx* gp = Z.a + column() + (long long)rg * rs;
    const int step = 16 * rs;
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      st_stream(gp, v[r]);
      gp += step;
      LMVN_KEEP_PTR(gp);
    }
  }
}

// ------------------------------------------------------------------------------
// K^ of the five-pass layout (main[z'][y'][kx] with nx/2 columns + Nyquist plane nyq[z'][y']) -> A layout.
// y' of the five-pass y plan: two stages R1y x R2y, position R2y * (ky % R1y) + ky / R1y.
// ------------------------------------------------------------------------------
struct KhatPermArgs {
  const cplx* main;
  const cplx* nyq;
  cplx* out;
  int nz, ny, r2;
  int y_r1, y_r2, y_r3;  // radices of the five-pass y plan (r3 = 1 for two-stage plans)
};

__device__ __forceinline__ int old_y_position(int ky, int r1, int r2, int r3) {
  // decimation in frequency: the first-stage output index is the LOW digit of the frequency and the HIGH digit of the position
  const int q1 = ky % r1, rest = ky / r1;
  if (r3 == 1) return q1 * r2 + rest;
  const int q2 = rest % r2, q3 = rest / r2;
  return (q1 * r2 + q2) * r3 + q3;
}

static __global__ void k_khat_to_a(KhatPermArgs K) {
  // one thread per output element
  const size_t n = size_t(K.r2) * K.nz * kTileElems;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const int kx = int(i % kM);
    const int p = int((i / kM) % (kRows + 1));
    const size_t zn = i / kTileElems;
    const int z = int(zn % K.nz);
    const int n2 = int(zn / K.nz);
    const int k2 = n2;  // slot n2 of pass B's y level holds ky = k1 + 128 n2
    const int pos = (p < kRows) ? p : kx;
    const int ky = ((pos >> 3) + ((pos & 7) << 4)) + kRows * k2;
    const int yo = old_y_position(ky, K.y_r1, K.y_r2, K.y_r3);
    const size_t row = size_t(z) * K.ny + yo;
    K.out[i] = (p < kRows) ? K.main[row * kM + kx] : K.nyq[row];
  }
}

}  // namespace x3
}  // namespace lmvn
