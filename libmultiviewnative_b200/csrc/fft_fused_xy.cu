// fft_fused_xy.cu -- instantiations and launchers of fast::k_xy (x and y passes of a plane in one
// persistent, ticket-ordered launch; the intermediate stays in L2).  See fft_fast.cuh.
#include <algorithm>

#include "fft_fused_xy.cuh"

namespace lmvn {
namespace fast {

namespace {

template <int M, int NY>
struct XY {
  static size_t smem() {
    return std::max(size_t(Row2Cfg<M>::SMEM), size_t(NY) * Cols<NY>::V * sizeof(cplx));
  }
  static int occupancy() {
#ifdef LMVN_EMU
    return 1;
#else
    int occ = 1 << 30;
    void (*ks[4])(XYArgs) = {k_xy<M, NY, false, gen::EPI_STORE>, k_xy<M, NY, true, gen::EPI_STORE>,
                             k_xy<M, NY, true, gen::EPI_QUOTIENT>, k_xy<M, NY, true, gen::EPI_UPDATE>};
    for (auto k : ks) {
      int n = 0;
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem())) != cudaSuccess) return 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, kRowThreads, smem()) != cudaSuccess) return 0;
      occ = std::min(occ, n);
    }
    return occ;
#endif
  }
  static int launch(bool inverse, const XYArgs& a, int grid, cudaStream_t s) {
    auto kf = k_xy<M, NY, false, gen::EPI_STORE>;
    auto k0 = k_xy<M, NY, true, gen::EPI_STORE>;
    auto k1 = k_xy<M, NY, true, gen::EPI_QUOTIENT>;
    auto k2 = k_xy<M, NY, true, gen::EPI_UPDATE>;
    const dim3 g{unsigned(grid)}, b{unsigned(kRowThreads)};
    if (!inverse) {
      LMVN_LAUNCH(kf, g, b, smem(), s, a);
    } else if (a.rows.ep.mode == gen::EPI_QUOTIENT) {
      LMVN_LAUNCH(k1, g, b, smem(), s, a);
    } else if (a.rows.ep.mode == gen::EPI_UPDATE) {
      LMVN_LAUNCH(k2, g, b, smem(), s, a);
    } else {
      LMVN_LAUNCH(k0, g, b, smem(), s, a);
    }
    return 0;
  }
};

}  // namespace

#ifdef LMVN_STRIDED_HALF
#define LMVN_XY_512(MM, EXPR)
#else
#define LMVN_XY_512(MM, EXPR) case 512: { typedef XY<MM, 512> K; EXPR; } break;
#endif
#define LMVN_XY_DISPATCH(EXPR)                                     \
  switch (M) {                                                     \
    case 32:                                                       \
      switch (ny) {                                                \
        case 128: { typedef XY<32, 128> K; EXPR; } break;          \
        case 256: { typedef XY<32, 256> K; EXPR; } break;          \
        LMVN_XY_512(32, EXPR)                                          \
        default: break;                                            \
      }                                                            \
      break;                                                       \
    case 64:                                                       \
      switch (ny) {                                                \
        case 128: { typedef XY<64, 128> K; EXPR; } break;          \
        case 256: { typedef XY<64, 256> K; EXPR; } break;          \
        LMVN_XY_512(64, EXPR)                                          \
        default: break;                                            \
      }                                                            \
      break;                                                       \
    case 128:                                                      \
      switch (ny) {                                                \
        case 128: { typedef XY<128, 128> K; EXPR; } break;         \
        case 256: { typedef XY<128, 256> K; EXPR; } break;         \
        LMVN_XY_512(128, EXPR)                                         \
        default: break;                                            \
      }                                                            \
      break;                                                       \
    default: break;                                                \
  }

bool xy_supported(int M, int ny, int* ctas_per_sm) {
  int occ = -1;
  LMVN_XY_DISPATCH(occ = K::occupancy())
  if (occ < 0) return false;
  *ctas_per_sm = occ;
  return true;
}

int xy_items_per_plane(int M, int ny, int ncols) {
  int rows_per_item = 0, cols = 0;
  switch (M) {
    case 32: rows_per_item = Row2Cfg<32>::ROWS; break;
    case 64: rows_per_item = Row2Cfg<64>::ROWS; break;
    case 128: rows_per_item = Row2Cfg<128>::ROWS; break;
    default: return 1;
  }
  switch (ny) {
    case 128: cols = Cols<128>::V; break;
    case 256: cols = Cols<256>::V; break;
    case 512: cols = Cols<512>::V; break;
    default: return 1;
  }
  return ny / rows_per_item + (ncols + cols - 1) / cols;
}

int launch_xy(int M, int ny, bool inverse, const XYArgs& a, int grid, cudaStream_t s) {
  int rc = -2;
  LMVN_XY_DISPATCH(rc = K::launch(inverse, a, grid, s))
  if (rc == -2) {
    set_last_error("fused x/y pass: unsupported shape (nx = %d, ny = %d)", 2 * M, ny);
    return -1;
  }
  return rc;
}

}  // namespace fast
}  // namespace lmvn
