// fft_tma.cuh -- strided (y, z) passes as persistent kernels fed by the TMA engine.  OPT-IN (LMVN_TMA=7): bit-identical to
// k_strided (same stage code) and measured slower on B200, see below.
//
// Why it was built: with its loads and stores taken out, a strided pass of fft_fast.cuh needs 0.055 ms (y) / 0.093 ms
// (merged z) of SM time on config 3, HBM needs 0.077 / 0.116 ms for its bytes -- and the pass takes 0.091 / 0.152 ms, because a
// CTA spends the first part of every tile waiting for its own loads and two co-resident CTAs overlap that only partly
// (profiles/r02_dry_pass_floor.log).  Here the loads are not issued by the threads that compute:
//   * one CTA of 512 threads per SM, persistent; its two 256-thread HALVES each work on a tile of their own (the same tile
//     shape and the very same stage code as k_strided: N rows x 16 kx columns, one 128-byte line per row), synchronised
//     by a named barrier per half;
//   * a ring of tile buffers (three of 64 KB for N = 512, six of 32 KB for N = 256): two in work, the others being filled by
//     `cp.async.bulk.tensor` (SASS UTMALDG; a 3-D tensor map [slow][row][16 columns] of the spectrum, boxes of 256 rows) that
//     completes on an mbarrier.  CTA-local item s goes to half s % 2 and buffer s % BUFS; the half that finishes item s (all
//     its shared-memory reads done, its stores on their way) issues the copy of item s + BUFS into the buffer it just
//     released;
//   * the first stage runs in place on the landed tile (strided_tile<..., SRC1 = W_SMEM>): the wavefront count of the
//     L1/shared pipe is the same as with register loads (a shared-memory read replaces the global load), but no warp ever
//     waits for HBM on the spectrum side;
//   * K^ of the merged z pass still goes HBM -> registers (64 registers per thread are the only place 64 KB per tile
//     fit), the results still leave from registers (st.global.cg, full lines);
//   * the Nyquist plane (split layout) is the tail of the item list: the half copies its tile into the buffer with ordinary
//     loads and runs the same code.
// Two things the first version got wrong, both found on the GPU (tools/tma_debug.py): the buffer is written by the in-place
// stages through the generic proxy and then by the copy through the async proxy -- every thread needs a
// fence.proxy.async before the releasing barrier, not just the issuing one; and an mbarrier wait knows only the PARITY of
// its phase, so the half that runs ahead must first see the previous use of the buffer released (`rel`).
// Measured (B200, profiles/r02_tma_strided_probe.log): config 3 y passes 0.112 against 0.093 ms, merged z 0.160 against
// 0.156 ms; 256^3 y 0.033 against 0.025 ms.  Why it loses: 512 threads x 32 values = 128 KB of tiles are in WORK at any time,
// which leaves room for ONE 64 KB tile in flight per SM (at N = 512), where an SM needs ~100-140 KB in flight to cover
// HBM latency at its 47 GB/s share; k_strided keeps its in-flight data in the REGISTER file (up to 128 KB per SM during
// the load phase of its two CTAs).  The register file is the bigger landing buffer -- the async ring would need > 227 KB.
// Built for sm_100a only; the emulated test build has no TMA and keeps k_strided.
#pragma once
#include "fft_fast.cuh"

#if !defined(LMVN_EMU)
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace lmvn {
namespace tma {

using fast::StridedArgs;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LMVN_TMA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LMVN_TMA_DONE;\n"
      "bra LMVN_TMA_WAIT;\n"
      "LMVN_TMA_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// box (c0 .. , c1 .. , c2) of the tensor map -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

static const int kHalf = 256;      // threads per tile
static const int kThreads = 512;   // two halves
template <int N> struct Cfg {
  static const int COLS = 16;
  static const int TILE_ELEMS = N * COLS;
  static const int TILE_BYTES = TILE_ELEMS * int(sizeof(cplx));
  static const int BOX_ROWS = N < 256 ? N : 256;
  // ring of tile buffers in 192 KB: three 64 KB tiles (N = 512), six 32 KB tiles (N = 256).  CTA-local item s lives in buffer
  // s % BUFS and is copied when item s - BUFS releases it
  static const int BUFS = (192 * 1024) / TILE_BYTES > 6 ? 6 : (192 * 1024) / TILE_BYTES;
  static const size_t SMEM = size_t(BUFS) * TILE_BYTES + 128;  // + mbarriers + release counters
};

// MODE: fast::SM_FWD, SM_INV or SM_FWD_MUL_INV on the default plan of the axis; tile_stride etc. as in k_strided.
// A.nyq_groups >= 0 (split layout): items [0, n_tiles) are tiles, [n_tiles, n_tiles + nyq_groups) Nyquist groups.
template <int N, int MODE>
static __global__ void __launch_bounds__(kThreads, 1)
    k_strided_tma(const __grid_constant__ CUtensorMap map, StridedArgs A, int n_tiles) {
  typedef Cfg<N> CF;
  constexpr int kBufs = CF::BUFS;
  static_assert(kBufs >= 3, "ring");
  static_assert(fast::TileCols<N, MODE>::V == CF::COLS && fast::TileThreads<N, MODE>::V == kHalf, "tile shape of k_strided");
  extern __shared__ __align__(128) unsigned char raw[];
  cplx* bufs = reinterpret_cast<cplx*>(raw);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(raw + size_t(kBufs) * CF::TILE_BYTES);
  // rel[b] = how many items have been consumed from buffer b.  An mbarrier wait only knows the PARITY of the phase it waits
  // for: the half that runs ahead (its own copies land early, the other half still waits for a late one) would take the
  // completed phase k - 2 of a buffer for phase k and read a tile that is still landing.  So the consumer of the k-th use of
  // a buffer first waits until the (k - 1)-th use has been released -- from then on the parity is unambiguous.
  volatile unsigned* rel = reinterpret_cast<volatile unsigned*>(full + kBufs);
  const int half = threadIdx.x / kHalf, ht = threadIdx.x % kHalf;
  const int c = ht % CF::COLS;
  constexpr bool ZMUL = (MODE == fast::SM_FWD_MUL_INV);
  constexpr int U = ZMUL ? LMVN_ZMUL_UNROLL : LMVN_Y_UNROLL;
  const int n_items = n_tiles + (A.nyq_groups > 0 ? A.nyq_groups : 0);
  const unsigned tiles_x = unsigned(A.tiles_x);

  // hands buffer s % 3 to CTA-local item s: a tile is copied into it (the copy completes the buffer's mbarrier phase); a
  // Nyquist group loads through registers and only needs the phase to complete
  auto issue = [&](int s) {
    const int g = int(blockIdx.x + s * gridDim.x);
    if (g >= n_items) return;
    const int b = int(s % kBufs);
    if (g >= n_tiles) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[b])) : "memory");
      return;
    }
    unsigned char* dst = raw + size_t(b) * CF::TILE_BYTES;
    const int bx = int(g % tiles_x), by = int(g / tiles_x);
    mbar_expect_tx(&full[b], CF::TILE_BYTES);
#pragma unroll
    for (int r = 0; r < N; r += CF::BOX_ROWS)
      tma_load_3d(dst + size_t(r) * CF::COLS * sizeof(cplx), &map, bx * CF::COLS * 2, r, by, &full[b]);
  };

  if (threadIdx.x == 0) {
    for (int b = 0; b < kBufs; ++b) {
      mbar_init(&full[b], 1);
      rel[b] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int s = 0; s < kBufs; ++s) issue(s);

  for (int s = half;; s += 2) {
    const int g = int(blockIdx.x + s * gridDim.x);
    if (g >= n_items) break;
    const int b = int(s % kBufs);
    cplx* sm = bufs + size_t(b) * CF::TILE_ELEMS;
    const unsigned use = unsigned(s / kBufs);
    while (rel[b] < use) {
    }
    mbar_wait(&full[b], use & 1);
    cplx* gp;
    const cplx* gk;
    int rs;
    bool live;
    if (g < n_tiles) {
      const unsigned bx = unsigned(g % tiles_x), by = unsigned(g / tiles_x);
      const int col = bx * CF::COLS + c;
      const long long base = (long long)by * A.tile_stride + col;
      gp = A.data + base;
      gk = A.khat + base;
      rs = A.row_stride;
      // (always true -- whole tiles only -- but a run-time condition around every stage keeps the compiler from hoisting
      // the twiddle loads of later stages across the barriers: 500 bytes of spills otherwise)
      live = col < A.ncols;
    } else {
      // Nyquist plane (2 MB, L2 resident): a "column" of this group is one slow index (see k_strided).  The half copies
      // its tile into the buffer with ordinary loads and then runs the very same code as for a landed tile.
      const unsigned sidx = unsigned(g - n_tiles) * CF::COLS + c;
      const long long nb = (long long)sidx * A.nyq_cs;
      gp = A.nyq + nb;
      gk = A.nyq_khat + nb;
      rs = A.nyq_rs;
      live = sidx < A.slow;
      if (live) {
        const cplx* src = gp + (long long)(ht / CF::COLS) * rs;
        cplx* dst = sm + ht;
#pragma unroll 8
        for (int i = 0; i < N / (kHalf / CF::COLS); ++i) {
          *dst = ld_stream(src);
          src += (kHalf / CF::COLS) * rs;
          dst += kHalf;
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(kHalf) : "memory");  // (uniform per half: g is)
    }
    fast::strided_tile<N, MODE, U, 0, 0, fast::W_SMEM, true>(A, sm + c, gp, gk, live, 0, rs, 1 + half);
    // every thread of the half has read what it needs from the buffer: hand it to item s + 3 (the tile the OTHER half
    // takes after its current one)
    // (the in-place stages wrote the buffer through the generic proxy, the copy writes it through the async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(kHalf) : "memory");
    if (ht == 0) {
      rel[b] = use + 1;
      issue(s + kBufs);
    }
  }
}

}  // namespace tma
}  // namespace lmvn
#endif  // !LMVN_EMU
