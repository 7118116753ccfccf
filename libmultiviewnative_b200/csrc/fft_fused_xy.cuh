// fft_fused_xy.cuh -- host-callable launchers of the persistent x/y kernels (fft_fused_xy.cu).
// Kept in a translation unit of their own so that the template instantiations compile in parallel.
#pragma once
#include "fft_fast.cuh"

namespace lmvn {
namespace fast {

// true when k_xy is instantiated for (M = nx/2, ny); *ctas_per_sm = resident CTAs per SM
bool xy_supported(int M, int ny, int* ctas_per_sm);
int xy_items_per_plane(int M, int ny, int ncols);
// returns 0 or -1 (last error set)
int launch_xy(int M, int ny, bool inverse, const XYArgs& a, int grid, cudaStream_t s);

}  // namespace fast
}  // namespace lmvn
