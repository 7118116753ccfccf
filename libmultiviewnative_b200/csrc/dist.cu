// dist.cu -- ONE volume over several GPUs: slab-decomposed 3-D FFT convolution with the
// all-to-all fused into the transform passes (BASELINE config 5).
//
// Real space is cut into slabs of nz/G planes (psi, views, weights, integral: all pointwise
// work is local).  Spectra of the z pass live in "pencils": all z for ny/G rows of y', with
// the PSF spectra K^1_v, K^2_v precomputed once IN that layout and never moved again.  Per
// convolution:
//
//   phase 0  x rows + y forward on the local planes; the LAST y stage does not store in
//            place: row y' goes straight to the pencil buffer of rank y' / (ny/G)  -- local
//            memory or PEER memory over NVLink (P2P stores), i.e. compute and all-to-all
//            are one kernel and the transfer overlaps the butterflies tile by tile;
//   barrier  (device side, flags in peer memory; stream ordered)
//   phase 1  z forward * K^ * z inverse on the pencil; its last stage scatters plane z to
//            the slab buffer of rank z / (nz/G) the same way;
//   barrier
//   phase 2  y inverse + x rows inverse + quotient / RL update on the local planes.
//
// There is no global scalar in the algorithm, hence no reduction: two exchanges per
// convolution, C (G-1)/G bytes in aggregate each (ref for the single-GPU loop:
// src/multiviewnative.cpp:191-229; the reference has no multi-GPU path for one volume).
//
// Ranks are either separate processes (one per GPU; exchange regions are shared through
// CUDA IPC handles that the host layer moves with torch.distributed) or several handles in
// ONE process (tests: the phases are then issued rank by rank in stream order and the
// barrier is a no-op).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <vector>

#include "engine.cuh"
#include "fft_fast.cuh"
#include "lmvn_b200.h"

#ifdef LMVN_EMU
// host emulation (tests): the peer-visible exchange region is POSIX shared memory, so that the multi-process
// path (handle exchange, peer stores, phases) runs under a gloo group on the CPU
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#endif

namespace lmvn {

namespace {

const int kMaxRanks = 8;
const int kMaxGroups = 8;                                  // column groups of the pipelined loop
const int kChannels = 1 + 2 * kMaxGroups;                  // barrier channels: whole phase, per group x {y scatter, z scatter}
const size_t kFlagBytes = sizeof(unsigned) * kChannels * kMaxRanks;

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
size_t align_up_(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Cross-GPU barrier: every rank raises its flag in every peer's exchange region, then waits for
// all of its own flags.  One CTA, one thread per peer.  The wait is bounded by wall time (%globaltimer):
// a lost peer must not hang the device.  On timeout the error flag is raised and STAYS raised; every later
// barrier of the plan then returns at once, the host sees the flag at the end of the call
// (check_device_error) and the plan refuses further work until lmvn_dist_reset_barrier has been called by
// every rank.  timeout_ns comes from LMVN_BARRIER_TIMEOUT_S (default 600 s: ranks may legitimately be
// seconds apart -- first-call graph instantiation, staging a pageable slab; ProcessSlabPlan.iterate
// additionally lines the hosts up with a torch.distributed barrier before the launch).
// channel: independent barrier sequences (own flags, own epoch) -- the pipelined loop runs one per column group and
// exchange direction, on different streams.
__global__ void k_peer_barrier(unsigned* const* peer_flags, unsigned* my_flags, int rank, int world, unsigned* epoch_counter,
                               unsigned* err, unsigned long long timeout_ns, int channel) {
#ifndef LMVN_EMU
  // the epoch lives on the device (every rank launches the same sequence of barriers), so that a captured
  // launch sequence can be replayed as a CUDA graph
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = ++epoch_counter[channel];
  __syncthreads();
  const unsigned epoch = s_epoch;
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    volatile unsigned* dst = peer_flags[t] + channel * kMaxRanks + rank;
    *dst = epoch;
    __threadfence_system();
    volatile unsigned* src = my_flags + channel * kMaxRanks + t;
    volatile unsigned* verr = err;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*src < epoch) {
      if (*verr) break;  // an earlier barrier of this plan timed out: do not wait again
      __nanosleep(200);
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > timeout_ns) {
        *verr = 1u;
        break;
      }
    }
    __threadfence_system();
  }
#else
  (void)peer_flags; (void)my_flags; (void)rank; (void)world; (void)epoch_counter; (void)err; (void)timeout_ns; (void)channel;
#endif
}

}  // namespace

struct DistDeconv {
  int rank = 0, world = 1, device = 0;
  int nz = 0, ny = 0, nx = 0, nz_l = 0, ny_l = 0, nxp = 0, num_views = 0;
  std::unique_ptr<FastOps> ops;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // local arena
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  float *psi = nullptr, *integral = nullptr, *kernel_stage = nullptr;
  size_t kernel_stage_elems = 0;
  std::vector<float*> image, weights;
  std::vector<cplx*> khat1, khat2;
  std::vector<char> view_set;
  bool psi_set = false;
  // peer-visible exchange region: [slab_work][pencil_work][flags]
  unsigned char* xchg = nullptr;
  size_t xchg_bytes = 0, slab_off = 0, pencil_off = 0, flags_off = 0;
  unsigned char* peer[kMaxRanks] = {nullptr};
  bool peer_ipc[kMaxRanks] = {false};
  bool multi_process = false;
  unsigned* d_epoch = nullptr;  // device-side barrier epoch
#ifndef LMVN_EMU
  cudaGraphExec_t sweep_graph = nullptr;  // one sweep over all views of the chained loop
#endif
  bool use_graph = true;
  double graph_lambda = 0.0;
  float graph_min = 0.f;
  // pipelined loop: the kx columns are cut into `groups` windows; the z pass of window g (and the y-inverse pass behind it)
  // runs on its own streams while window g+1 is still being scattered, per-window barriers instead of whole-phase ones
  // OPT-IN (LMVN_DIST_GROUPS > 1): measured on 8 x B200, 1024^3 (profiles/r02_slab_phase_timing_8gpu.log): 6.12 ms per
  // (view, iteration) unpipelined, 6.23 with 2 windows, 6.39 with 4 -- the scatter kernels fill every CTA slot of the
  // device while they wait on their peer stores (y scatter 1.1 ms, z scatter 1.2 ms for 470 MB out of each GPU = ~400 GB/s),
  // so the passes of the other windows queue behind them instead of running beside them, and only 0.65 ms of the 3.2 ms
  // of a convolution is local work that could hide.  The limiter is the rate of the scattered peer stores, not the order.
  int groups = 1;
  cudaStream_t s_z = nullptr, s_yi = nullptr;
  cudaEvent_t ev_y[kMaxGroups] = {nullptr}, ev_z[kMaxGroups] = {nullptr}, ev_yi[kMaxGroups] = {nullptr}, ev_rows = nullptr;
  unsigned** d_peer_flags = nullptr;  // device array of kMaxRanks pointers
  unsigned* d_err = nullptr;
  bool barrier_failed = false;                         // a device-side barrier timed out; reset_barrier() clears it
  unsigned long long barrier_timeout_ns = 600ull * 1000000000ull;
  // comparator: exchanges staged through local buffers and moved by the caller (NCCL all-to-all)
  bool staged = false;
  bool own_stream = true;
  cplx* stage_send = nullptr;
  cplx* stage_recv = nullptr;
#ifdef LMVN_EMU
  char shm_name[64] = {0};
  int alloc_exchange() {
    static int counter = 0;
    std::snprintf(shm_name, sizeof(shm_name), "/lmvn_emu_%d_%d_%d", int(getpid()), rank, counter++);
    const int fd = shm_open(shm_name, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, off_t(xchg_bytes)) != 0) { set_last_error("emu: shm_open failed"); return -1; }
    void* p = mmap(nullptr, xchg_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { set_last_error("emu: mmap failed"); return -1; }
    xchg = static_cast<unsigned char*>(p);
    return 0;
  }
  void free_exchange() {
    for (int r = 0; r < kMaxRanks; ++r)
      if (peer_ipc[r] && peer[r]) munmap(peer[r], xchg_bytes);
    if (xchg) {
      munmap(xchg, xchg_bytes);
      shm_unlink(shm_name);
    }
  }
#endif

  size_t slab_real() const { return size_t(nz_l) * ny * nx; }
  size_t slab_spec() const { return size_t(nz_l) * ny * nxp; }
  size_t pencil_spec() const { return size_t(nz) * ny_l * nxp; }
  cplx* slab_work(int r) const { return reinterpret_cast<cplx*>(peer[r] + slab_off); }
  cplx* pencil_work(int r) const { return reinterpret_cast<cplx*>(peer[r] + pencil_off); }
  unsigned* flags(int r) const { return reinterpret_cast<unsigned*>(peer[r] + flags_off); }

  ~DistDeconv() {
    if (arena || xchg || stream) cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
#ifdef LMVN_EMU
    free_exchange();
    xchg = nullptr;
#else
    for (int r = 0; r < kMaxRanks; ++r)
      if (peer_ipc[r] && peer[r]) cudaIpcCloseMemHandle(peer[r]);
#endif
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (int g = 0; g < kMaxGroups; ++g) {
      if (ev_y[g]) cudaEventDestroy(ev_y[g]);
      if (ev_z[g]) cudaEventDestroy(ev_z[g]);
      if (ev_yi[g]) cudaEventDestroy(ev_yi[g]);
    }
    if (ev_rows) cudaEventDestroy(ev_rows);
    if (s_z) { cudaStreamSynchronize(s_z); cudaStreamDestroy(s_z); }
    if (s_yi) { cudaStreamSynchronize(s_yi); cudaStreamDestroy(s_yi); }
    if (stream && own_stream) cudaStreamDestroy(stream);
    if (stage_send) cudaFree(stage_send);
    if (stage_recv) cudaFree(stage_recv);
    if (d_peer_flags) cudaFree(d_peer_flags);
    if (d_err) cudaFree(d_err);
    if (d_epoch) cudaFree(d_epoch);
#ifndef LMVN_EMU
    if (sweep_graph) cudaGraphExecDestroy(sweep_graph);
#endif
    if (arena) {
      arena_unpoison(arena, arena_bytes);
      cudaFree(arena);
    }
#ifndef LMVN_EMU
    if (xchg) cudaFree(xchg);
#endif
  }

  int init(const int* d, int nviews, int rank_, int world_, int dev) {
    if (!d || world_ < 1 || world_ > kMaxRanks || rank_ < 0 || rank_ >= world_ || (world_ & (world_ - 1))) {
      set_last_error("slab-decomposed plan: world size must be a power of two <= %d and 0 <= rank < world", kMaxRanks);
      return -1;
    }
    if (nviews <= 0) { set_last_error("invalid number of views %d", nviews); return -1; }
    device = resolve_device(dev);
    if (device < 0) return -1;
    LMVN_CUDA_TRY(cudaSetDevice(device));
    rank = rank_; world = world_; num_views = nviews;
    nz = d[0]; ny = d[1]; nx = d[2];
    if (nz % world || ny % world) {
      set_last_error("slab-decomposed plan: nz = %d and ny = %d must be divisible by the world size %d", nz, ny, world);
      return -1;
    }
    nz_l = nz / world; ny_l = ny / world;
    auto fp = get_fft_plan(device, nz, ny, nx);
    if (!fp) return -1;
    ops = make_fast_ops(fp);
    if (!ops) return -1;
    nxp = ops->nxp_pitch();
    if ((size_t(nz_l) * ny) % 128 != 0) {
      set_last_error("slab-decomposed plan: a slab needs a multiple of 128 rows (nz/G * ny = %zu)", size_t(nz_l) * ny);
      return -1;
    }
    const size_t S = align_up_(slab_real() * sizeof(float), 256);
    const size_t K = align_up_(pencil_spec() * sizeof(cplx), 256);
    kernel_stage_elems = std::min<size_t>(size_t(1) << 24, size_t(nz) * ny * nx);
    const size_t KS = align_up_(kernel_stage_elems * sizeof(float), 256);
    arena_bytes = 2 * S + KS + size_t(nviews) * (2 * S + 2 * K) + kArenaRedzone * (3 + 4 * size_t(nviews));
    slab_off = 0;
    pencil_off = align_up_(slab_spec() * sizeof(cplx), 256);
    flags_off = pencil_off + align_up_(pencil_spec() * sizeof(cplx), 256);
    xchg_bytes = flags_off + kFlagBytes;
    size_t free_b = 0, total_b = 0;
    LMVN_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    if (arena_bytes + xchg_bytes > free_b) {
      set_last_error("slab-decomposed plan needs %.2f GiB of device memory per rank, %.2f GiB free",
                     (arena_bytes + xchg_bytes) / 1073741824.0, free_b / 1073741824.0);
      return -1;
    }
    LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&arena), arena_bytes));
#ifdef LMVN_EMU
    LMVN_TRY(alloc_exchange());
#else
    LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&xchg), xchg_bytes));
#endif
    LMVN_CUDA_TRY(cudaMemset(xchg + flags_off, 0, kFlagBytes));
    LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_peer_flags), sizeof(unsigned*) * kMaxRanks));
    LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_err), sizeof(unsigned)));
    LMVN_CUDA_TRY(cudaMemset(d_err, 0, sizeof(unsigned)));
    LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_epoch), sizeof(unsigned) * kChannels));
    LMVN_CUDA_TRY(cudaMemset(d_epoch, 0, sizeof(unsigned) * kChannels));
    if (const char* e = getenv("LMVN_GRAPH")) use_graph = (*e != '0');
    if (const char* e = getenv("LMVN_BARRIER_TIMEOUT_S")) {
      const double sec = atof(e);
      if (sec > 0) barrier_timeout_ns = (unsigned long long)(sec * 1e9);
    }
    unsigned char* p = arena;
    auto take = [&](size_t bytes) {
      unsigned char* r = p;
      p += bytes;
      arena_poison(p, kArenaRedzone);
      p += kArenaRedzone;
      return r;
    };
    psi = reinterpret_cast<float*>(take(S));
    integral = reinterpret_cast<float*>(take(S));
    kernel_stage = reinterpret_cast<float*>(take(KS));
    image.resize(nviews); weights.resize(nviews); khat1.resize(nviews); khat2.resize(nviews);
    view_set.assign(nviews, 0);
    for (int v = 0; v < nviews; ++v) {
      image[v] = reinterpret_cast<float*>(take(S));
      weights[v] = reinterpret_cast<float*>(take(S));
      khat1[v] = reinterpret_cast<cplx*>(take(K));
      khat2[v] = reinterpret_cast<cplx*>(take(K));
    }
    peer[rank] = xchg;
    LMVN_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    LMVN_CUDA_TRY(cudaEventCreate(&ev0));
    LMVN_CUDA_TRY(cudaEventCreate(&ev1));
    LMVN_CUDA_TRY(cudaStreamCreateWithFlags(&s_z, cudaStreamNonBlocking));
    LMVN_CUDA_TRY(cudaStreamCreateWithFlags(&s_yi, cudaStreamNonBlocking));
    for (int g = 0; g < kMaxGroups; ++g) {
      LMVN_CUDA_TRY(cudaEventCreateWithFlags(&ev_y[g], cudaEventDisableTiming));
      LMVN_CUDA_TRY(cudaEventCreateWithFlags(&ev_z[g], cudaEventDisableTiming));
      LMVN_CUDA_TRY(cudaEventCreateWithFlags(&ev_yi[g], cudaEventDisableTiming));
    }
    LMVN_CUDA_TRY(cudaEventCreateWithFlags(&ev_rows, cudaEventDisableTiming));
    if (const char* e = getenv("LMVN_DIST_GROUPS")) groups = std::max(1, std::min(kMaxGroups, atoi(e)));
    return 0;
  }

  // column windows of the pipelined loop: tile aligned for both the y and the z passes, at most `groups` of them
  int window_count() const {
    const int tile = std::max(ops->strided_tile_cols(ny), ops->strided_tile_cols(nz));
    const int tiles = (nx / 2 + 1 + tile - 1) / tile;
    return std::max(1, std::min(groups, tiles));
  }
  void window(int g, int ng, int* col0, int* ncols) const {
    const int tile = std::max(ops->strided_tile_cols(ny), ops->strided_tile_cols(nz));
    const int tiles = (nx / 2 + 1 + tile - 1) / tile;
    const int t0 = int((long long)tiles * g / ng), t1 = int((long long)tiles * (g + 1) / ng);
    *col0 = t0 * tile;
    *ncols = std::min((t1 - t0) * tile, nx / 2 + 1 - *col0);
  }

  int connected() const {
    for (int r = 0; r < world; ++r)
      if (!peer[r]) { set_last_error("rank %d: peer %d is not connected", rank, r); return -1; }
    return 0;
  }
  int publish_peers() {
    unsigned* h[kMaxRanks] = {nullptr};
    for (int r = 0; r < world; ++r) h[r] = flags(r);
    LMVN_CUDA_TRY(cudaMemcpy(d_peer_flags, h, sizeof(h), cudaMemcpyHostToDevice));
    return 0;
  }

  // ---- phases -------------------------------------------------------------------
  StridedGeom y_geom(int mode) const {
    StridedGeom g;
    g.data = slab_work(rank);
    g.n = ny; g.tw_axis = 1; g.row_stride = nxp; g.tile_stride = (long long)ny * nxp; g.slow = unsigned(nz_l);
    g.mode = mode;
    if (mode == fast::SM_FWD_SCATTER) {
      // row y' of local plane z_l -> rank y'/ny_l, pencil [z = rank*nz_l + z_l][y' % ny_l][kx]
      g.sc.shift = ilog2(ny_l);
      g.sc.row_stride = nxp;
      g.sc.tile_stride = (long long)ny_l * nxp;
      if (!staged) {
        for (int r = 0; r < world; ++r) g.sc.base[r] = pencil_work(r);
        g.sc.offset = (long long)rank * nz_l * ny_l * nxp;
      } else {  // send block of destination r: [z_l][y'_l][kx]; the all-to-all puts it at [src = rank] over there
        for (int r = 0; r < world; ++r) g.sc.base[r] = stage_send + size_t(r) * nz_l * ny_l * nxp;
        g.sc.offset = 0;
      }
    }
    return g;
  }
  StridedGeom z_geom(int mode, const cplx* khat, float scale) const {
    StridedGeom g;
    g.data = pencil_work(rank);
    g.khat = khat;
    g.n = nz; g.tw_axis = 0; g.row_stride = ny_l * nxp; g.tile_stride = nxp; g.slow = unsigned(ny_l);
    g.mode = mode;
    g.scale = scale;
    if (mode == fast::SM_FWD_MUL_INV_SCATTER) {
      // plane z of local row y'_l -> rank z/nz_l, slab [z % nz_l][y' = rank*ny_l + y'_l][kx]
      g.sc.shift = ilog2(nz_l);
      g.sc.tile_stride = nxp;
      if (!staged) {
        for (int r = 0; r < world; ++r) g.sc.base[r] = slab_work(r);
        g.sc.row_stride = (long long)ny * nxp;
        g.sc.offset = (long long)rank * ny_l * nxp;
      } else {  // send block of destination r: [z_l][y'_l][kx]; the receiver still has to interleave the sources
        for (int r = 0; r < world; ++r) g.sc.base[r] = stage_send + size_t(r) * nz_l * ny_l * nxp;
        g.sc.row_stride = (long long)ny_l * nxp;
        g.sc.offset = 0;
      }
    }
    return g;
  }

  int forward_to_pencils(const gen::RealSource& src) {
    LMVN_TRY(ops->rows_fwd_planes(src, slab_work(rank), nz_l, rank * nz_l, nz, stream));
    LMVN_TRY(ops->strided_geom(y_geom(fast::SM_FWD_SCATTER), stream));
    return 0;
  }

  // which: 1 = psi (*) kernel1 -> quotient into integral; 2 = integral (*) kernel2 -> RL update of psi
  int conv_phase(int view, int which, int phase, const UpdateParams& up) {
    LMVN_CUDA_TRY(cudaSetDevice(device));
    if (phase == 0) {
      gen::RealSource src{which == 1 ? psi : integral, 0, 0, 0, 0};
      return forward_to_pencils(src);
    }
    if (phase == 1)
      return ops->strided_geom(z_geom(fast::SM_FWD_MUL_INV_SCATTER, which == 1 ? khat1[view] : khat2[view], 1.f), stream);
    LMVN_TRY(ops->strided_geom(y_geom(fast::SM_INV), stream));
    if (which == 1) {
      gen::Epilogue e{gen::EPI_QUOTIENT, 1.f, image[view], nullptr, nullptr, up};
      return ops->rows_inv_planes(slab_work(rank), integral, e, nz_l, stream);
    }
    gen::Epilogue e{gen::EPI_UPDATE, 1.f, nullptr, psi, weights[view], up};
    return ops->rows_inv_planes(slab_work(rank), psi, e, nz_l, stream);
  }

  // PSF spectrum of `which` kernel of `view` in pencil layout (1/N folded in)
  int psf_phase(int view, int which, int phase, const float* kernel_h, const int* kd) {
    LMVN_CUDA_TRY(cudaSetDevice(device));
    if (phase == 0) {
      if (!kernel_h || !kd) { set_last_error("kernel missing"); return -1; }
      const int dims[3] = {nz, ny, nx};
      for (int a = 0; a < 3; ++a)
        if (kd[a] <= 0 || kd[a] > dims[a]) {
          set_last_error("kernel extent %d along axis %d does not fit the image extent %d", kd[a], a, dims[a]);
          return -1;
        }
      const size_t kn = size_t(kd[0]) * kd[1] * kd[2];
      if (kn > kernel_stage_elems) { set_last_error("kernel too large for the staging buffer"); return -1; }
      LMVN_CUDA_TRY(cudaMemcpyAsync(kernel_stage, kernel_h, kn * sizeof(float), cudaMemcpyHostToDevice, stream));
      gen::RealSource src{kernel_stage, 1, kd[0], kd[1], kd[2]};
      return forward_to_pencils(src);
    }
    const float inv_n = float(1.0 / (double(nz) * ny * nx));
    LMVN_TRY(ops->strided_geom(z_geom(fast::SM_FWD_SCALE, nullptr, inv_n), stream));
    cplx* dst = (which == 1) ? khat1[view] : khat2[view];
    LMVN_CUDA_TRY(cudaMemcpyAsync(dst, pencil_work(rank), pencil_spec() * sizeof(cplx), cudaMemcpyDeviceToDevice, stream));
    return 0;
  }

  int barrier(int channel = 0, cudaStream_t s = nullptr) {
    if (!multi_process || world == 1) return 0;  // one process: stream order is the barrier
    LMVN_CUDA_TRY(cudaSetDevice(device));
    LMVN_LAUNCH(k_peer_barrier, dim3(1), dim3(32), 0, s ? s : stream, d_peer_flags, flags(rank), rank, world, d_epoch, d_err,
                barrier_timeout_ns, channel);
    LMVN_CUDA_TRY(cudaGetLastError());
    return 0;
  }

  // Collective recovery after a timed-out barrier: the caller lines every rank up on the host (no kernel of the
  // plan in flight anywhere), every rank calls this, the caller lines them up again.  Epochs restart from zero.
  int reset_barrier() {
    LMVN_CUDA_TRY(cudaSetDevice(device));
    LMVN_CUDA_TRY(cudaStreamSynchronize(stream));
    LMVN_CUDA_TRY(cudaMemset(d_err, 0, sizeof(unsigned)));
    LMVN_CUDA_TRY(cudaMemset(d_epoch, 0, sizeof(unsigned) * kChannels));
    LMVN_CUDA_TRY(cudaMemset(xchg + flags_off, 0, kFlagBytes));
    LMVN_CUDA_TRY(cudaDeviceSynchronize());
    barrier_failed = false;
    return 0;
  }

  int check_device_error() {
    unsigned e = 0;
    LMVN_CUDA_TRY(cudaMemcpyAsync(&e, d_err, sizeof(e), cudaMemcpyDeviceToHost, stream));
    LMVN_CUDA_TRY(cudaStreamSynchronize(stream));
    if (e) {
      barrier_failed = true;
      set_last_error("rank %d: cross-GPU barrier timed out (a peer did not arrive within %.0f s); the result of this "
                     "call is invalid, call lmvn_dist_reset_barrier on every rank before using the plan again",
                     rank, barrier_timeout_ns * 1e-9);
      return -1;
    }
    return 0;
  }

  int upload_slab(float* dst, const float* src_h) {
    if (!src_h) { set_last_error("null slab buffer"); return -1; }
    LMVN_CUDA_TRY(cudaSetDevice(device));
    LMVN_TRY(copy_to_device(device, dst, src_h, slab_real() * sizeof(float), stream));
    return 0;
  }

  // multi-process only: the whole loop with device-side barriers, no host round trip
  int iterate(int iterations, double lambda, float min_value, float* device_ms) {
    LMVN_TRY(connected());
    if (!psi_set) { set_last_error("psi has not been set"); return -1; }
    for (int v = 0; v < num_views; ++v)
      if (!view_set[v]) { set_last_error("view %d has not been set", v); return -1; }
    if (!multi_process && world > 1) {
      set_last_error("lmvn_dist_iterate needs one process per rank; in-process groups drive lmvn_dist_conv_phase");
      return -1;
    }
    if (barrier_failed) {
      set_last_error("rank %d: an earlier cross-GPU barrier timed out; call lmvn_dist_reset_barrier on every rank", rank);
      return -1;
    }
    LMVN_CUDA_TRY(cudaSetDevice(device));
    const UpdateParams up = make_update_params(lambda, min_value);
    LMVN_TRY(barrier());  // nobody starts scattering before every peer is ready
    LMVN_CUDA_TRY(cudaEventRecord(ev0, stream));
    if (ops->can_chain_rows() && !staged && iterations > 0) {
      // chained loop (see Deconv::iterate): the x-inverse pass also runs the x-forward pass of the next
      // convolution, in place on the slab's spectrum rows
      const int ng = window_count();
      // Pipelined exchange (ng > 1).  Per convolution, on three streams:
      //   stream  (link)   y forward + scatter of window 0, 1, .. back to back; joins; chained rows pass
      //   s_z     (link)   per window: wait for its scatter, barrier [all ranks scattered it], z fwd * K^ * z inv + scatter
      //                    back, barrier [all ranks scattered it back]
      //   s_yi    (local)  per window: y inverse
      // so the z pass of window g (HBM bound: data + K^) and the y-inverse pass of window g-1 (local) run while window
      // g+1 is still crossing NVLink, instead of scatter | barrier | z pass | barrier | y inverse with the links idle
      // during every local phase.  Same kernels on the same data: bit-identical to the unpipelined loop.
      auto pipelined_conv = [&](int v, int which) -> int {
        const cplx* kh = (which == 1) ? khat1[v] : khat2[v];
        for (int g = 0; g < ng; ++g) {
          int c0, nc;
          window(g, ng, &c0, &nc);
          StridedGeom yg = y_geom(fast::SM_FWD_SCATTER);
          yg.col0 = c0; yg.ncols = nc;
          LMVN_TRY(ops->strided_geom(yg, stream));
          LMVN_CUDA_TRY(cudaEventRecord(ev_y[g], stream));
          LMVN_CUDA_TRY(cudaStreamWaitEvent(s_z, ev_y[g], 0));
          LMVN_TRY(barrier(1 + g, s_z));
          StridedGeom zg = z_geom(fast::SM_FWD_MUL_INV_SCATTER, kh, 1.f);
          zg.col0 = c0; zg.ncols = nc;
          LMVN_TRY(ops->strided_geom(zg, s_z));
          LMVN_TRY(barrier(1 + kMaxGroups + g, s_z));
          LMVN_CUDA_TRY(cudaEventRecord(ev_z[g], s_z));
          LMVN_CUDA_TRY(cudaStreamWaitEvent(s_yi, ev_z[g], 0));
          StridedGeom ig = y_geom(fast::SM_INV);
          ig.col0 = c0; ig.ncols = nc;
          LMVN_TRY(ops->strided_geom(ig, s_yi));
          LMVN_CUDA_TRY(cudaEventRecord(ev_yi[g], s_yi));
        }
        for (int g = 0; g < ng; ++g) LMVN_CUDA_TRY(cudaStreamWaitEvent(stream, ev_yi[g], 0));
        return 0;
      };
      auto sweep = [&](bool ends_call) -> int {  // one iteration = one sweep over all views
        for (int v = 0; v < num_views; ++v)
          for (int which = 1; which <= 2; ++which) {
            const bool last = (ends_call && v == num_views - 1 && which == 2);
            if (ng > 1 && multi_process) {
              LMVN_TRY(pipelined_conv(v, which));
            } else {
            LMVN_TRY(ops->strided_geom(y_geom(fast::SM_FWD_SCATTER), stream));
            LMVN_TRY(barrier());
            LMVN_TRY(conv_phase(v, which, 1, up));
            LMVN_TRY(barrier());
            LMVN_TRY(ops->strided_geom(y_geom(fast::SM_INV), stream));
            }
            gen::Epilogue e = (which == 1) ? gen::Epilogue{gen::EPI_QUOTIENT, 1.f, image[v], nullptr, nullptr, up}
                                           : gen::Epilogue{gen::EPI_UPDATE, 1.f, nullptr, psi, weights[v], up};
            if (last) LMVN_TRY(ops->rows_inv_planes(slab_work(rank), psi, e, nz_l, stream));
            else LMVN_TRY(ops->rows_inv_fwd_planes(slab_work(rank), e, nz_l, stream));
          }
        return 0;
      };
      gen::RealSource src{psi, 0, 0, 0, 0};
      LMVN_TRY(ops->rows_fwd_planes(src, slab_work(rank), nz_l, rank * nz_l, nz, stream));
      int done = 0;
#ifndef LMVN_EMU
      // every rank captures and replays the same sequence (kernels, device-side barriers included)
      if (use_graph && iterations > 2) {
        if (!sweep_graph || graph_lambda != lambda || graph_min != min_value) {
          if (sweep_graph) { cudaGraphExecDestroy(sweep_graph); sweep_graph = nullptr; }
          cudaGraph_t g = nullptr;
          LMVN_CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
          const int rc = sweep(false);
          const cudaError_t ce = cudaStreamEndCapture(stream, &g);
          if (rc != 0 || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            (void)cudaGetLastError();
            use_graph = false;
          } else {
            if (cudaGraphInstantiate(&sweep_graph, g, 0) != cudaSuccess) { sweep_graph = nullptr; use_graph = false; (void)cudaGetLastError(); }
            cudaGraphDestroy(g);
            graph_lambda = lambda;
            graph_min = min_value;
          }
        }
        if (sweep_graph)
          for (; done < iterations - 1; ++done) LMVN_CUDA_TRY(cudaGraphLaunch(sweep_graph, stream));
      }
#endif
      for (; done < iterations; ++done) LMVN_TRY(sweep(done == iterations - 1));
    } else {
      for (int it = 0; it < iterations; ++it)
        for (int v = 0; v < num_views; ++v)
          for (int which = 1; which <= 2; ++which) {
            LMVN_TRY(conv_phase(v, which, 0, up));
            LMVN_TRY(barrier());
            LMVN_TRY(conv_phase(v, which, 1, up));
            LMVN_TRY(barrier());
            LMVN_TRY(conv_phase(v, which, 2, up));
          }
    }
    LMVN_CUDA_TRY(cudaEventRecord(ev1, stream));
    LMVN_TRY(check_device_error());
    if (device_ms) LMVN_CUDA_TRY(cudaEventElapsedTime(device_ms, ev0, ev1));
    return 0;
  }
};

}  // namespace lmvn

using namespace lmvn;

struct lmvn_dist {
  DistDeconv d;
};

#define LMVN_DIST_GUARD(h)                          \
  clear_last_error();                               \
  if (!(h)) {                                       \
    set_last_error("null slab-decomposed plan");    \
    return -1;                                      \
  }

extern "C" int lmvn_dist_create(lmvn_dist** out, const int* dims_zyx, int num_views, int rank, int world, int device) {
  clear_last_error();
  if (!out) { set_last_error("null output pointer"); return -1; }
  *out = nullptr;
  lmvn_dist* h = new (std::nothrow) lmvn_dist();
  if (!h) { set_last_error("out of host memory"); return -1; }
  if (h->d.init(dims_zyx, num_views, rank, world, device) != 0) {
    delete h;
    return -1;
  }
  *out = h;
  return 0;
}
extern "C" void lmvn_dist_destroy(lmvn_dist* h) { delete h; }

extern "C" int lmvn_dist_get_info(const lmvn_dist* h, lmvn_dist_info* info) {
  LMVN_DIST_GUARD(h);
  if (!info) { set_last_error("null info"); return -1; }
  const DistDeconv& d = h->d;
  std::memset(info, 0, sizeof(*info));
  info->dims[0] = d.nz; info->dims[1] = d.ny; info->dims[2] = d.nx;
  info->num_views = d.num_views; info->rank = d.rank; info->world = d.world; info->device = d.device;
  info->planes_per_rank = d.nz_l; info->rows_per_rank = d.ny_l; info->spectrum_pitch = d.nxp;
  info->arena_bytes = d.arena_bytes + d.xchg_bytes;
  info->exchange_bytes = d.xchg_bytes;
  const unsigned long long S = (unsigned long long)d.nz * d.ny * d.nx * 4ull;
  const unsigned long long C = (unsigned long long)d.nz * d.ny * (d.nx / 2 + 1) * 8ull;
  info->alg_bytes_per_view_iteration = 7 * S + 10 * C;
  // 4 exchanges per (view, iteration), each moves C (G-1)/G in aggregate
  info->exchange_bytes_per_view_iteration = 4 * C / (unsigned long long)d.world * (unsigned long long)(d.world - 1);
  return 0;
}

extern "C" int lmvn_dist_export_handle(lmvn_dist* h, void* handle64) {
  LMVN_DIST_GUARD(h);
#ifdef LMVN_EMU
  std::memset(handle64, 0, 64);
  std::memcpy(handle64, h->d.shm_name, std::strlen(h->d.shm_name));
  return 0;
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  LMVN_CUDA_TRY(cudaSetDevice(h->d.device));
  cudaIpcMemHandle_t m;
  LMVN_CUDA_TRY(cudaIpcGetMemHandle(&m, h->d.xchg));
  std::memcpy(handle64, &m, 64);
  return 0;
#endif
}

extern "C" int lmvn_dist_connect_ipc(lmvn_dist* h, int peer_rank, const void* handle64) {
  LMVN_DIST_GUARD(h);
#ifdef LMVN_EMU
  {
    DistDeconv& d = h->d;
    if (peer_rank < 0 || peer_rank >= d.world || peer_rank == d.rank) { set_last_error("bad peer rank %d", peer_rank); return -1; }
    char name[65] = {0};
    std::memcpy(name, handle64, 64);
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) { set_last_error("emu: cannot open the exchange region of rank %d", peer_rank); return -1; }
    void* p = mmap(nullptr, d.xchg_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { set_last_error("emu: mmap of the peer region failed"); return -1; }
    d.peer[peer_rank] = static_cast<unsigned char*>(p);
    d.peer_ipc[peer_rank] = true;
    d.multi_process = true;
    return 0;
  }
#else
  DistDeconv& d = h->d;
  if (peer_rank < 0 || peer_rank >= d.world || peer_rank == d.rank) { set_last_error("bad peer rank %d", peer_rank); return -1; }
  LMVN_CUDA_TRY(cudaSetDevice(d.device));
  cudaIpcMemHandle_t m;
  std::memcpy(&m, handle64, 64);
  void* p = nullptr;
  LMVN_CUDA_TRY(cudaIpcOpenMemHandle(&p, m, cudaIpcMemLazyEnablePeerAccess));
  d.peer[peer_rank] = static_cast<unsigned char*>(p);
  d.peer_ipc[peer_rank] = true;
  d.multi_process = true;
  bool all = true;
  for (int r = 0; r < d.world; ++r) all = all && d.peer[r];
  if (all) LMVN_TRY(d.publish_peers());
  return 0;
#endif
}

extern "C" int lmvn_dist_connect_local(lmvn_dist* h, int peer_rank, lmvn_dist* other) {
  LMVN_DIST_GUARD(h);
  DistDeconv& d = h->d;
  if (!other || peer_rank < 0 || peer_rank >= d.world || peer_rank == d.rank || other->d.rank != peer_rank ||
      other->d.world != d.world || other->d.xchg_bytes != d.xchg_bytes) {
    set_last_error("connect_local: peer %d does not match this plan", peer_rank);
    return -1;
  }
  if (other->d.device != d.device) {
    // same process, different GPUs: plain peer access
    LMVN_CUDA_TRY(cudaSetDevice(d.device));
    cudaError_t e = cudaDeviceEnablePeerAccess(other->d.device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      set_last_error("cudaDeviceEnablePeerAccess(%d -> %d) failed", d.device, other->d.device);
      return -1;
    }
    (void)cudaGetLastError();
  }
  d.peer[peer_rank] = other->d.xchg;
  return 0;
}

extern "C" int lmvn_dist_set_view_slab(lmvn_dist* h, int view, const float* image_slab, const float* weights_slab) {
  LMVN_DIST_GUARD(h);
  DistDeconv& d = h->d;
  if (view < 0 || view >= d.num_views) { set_last_error("view index %d out of range", view); return -1; }
  LMVN_TRY(d.upload_slab(d.image[view], image_slab));
  LMVN_TRY(d.upload_slab(d.weights[view], weights_slab));
  d.view_set[view] = 1;
  return 0;
}
extern "C" int lmvn_dist_set_psi_slab(lmvn_dist* h, const float* psi_slab) {
  LMVN_DIST_GUARD(h);
  LMVN_TRY(h->d.upload_slab(h->d.psi, psi_slab));
  h->d.psi_set = true;
  return 0;
}
extern "C" int lmvn_dist_get_psi_slab(lmvn_dist* h, float* psi_slab) {
  LMVN_DIST_GUARD(h);
  DistDeconv& d = h->d;
  if (!psi_slab) { set_last_error("psi is null"); return -1; }
  LMVN_CUDA_TRY(cudaSetDevice(d.device));
  LMVN_TRY(copy_to_host(d.device, psi_slab, d.psi, d.slab_real() * sizeof(float), d.stream));
  LMVN_CUDA_TRY(cudaStreamSynchronize(d.stream));
  return 0;
}

extern "C" int lmvn_dist_psf_phase(lmvn_dist* h, int view, int which_kernel, int phase, const float* kernel,
                                   const int* kernel_dims) {
  LMVN_DIST_GUARD(h);
  DistDeconv& d = h->d;
  if (view < 0 || view >= d.num_views || which_kernel < 1 || which_kernel > 2 || phase < 0 || phase > 1) {
    set_last_error("psf_phase: bad view / kernel / phase");
    return -1;
  }
  LMVN_TRY(d.connected());
  return d.psf_phase(view, which_kernel, phase, kernel, kernel_dims);
}

extern "C" int lmvn_dist_conv_phase(lmvn_dist* h, int view, int which_conv, int phase, double lambda, float min_value) {
  LMVN_DIST_GUARD(h);
  DistDeconv& d = h->d;
  if (view < 0 || view >= d.num_views || which_conv < 1 || which_conv > 2 || phase < 0 || phase > 2) {
    set_last_error("conv_phase: bad view / convolution / phase");
    return -1;
  }
  LMVN_TRY(d.connected());
  return d.conv_phase(view, which_conv, phase, make_update_params(lambda, min_value));
}

extern "C" int lmvn_dist_barrier(lmvn_dist* h) {
  LMVN_DIST_GUARD(h);
  LMVN_TRY(h->d.connected());
  return h->d.barrier();
}

extern "C" int lmvn_dist_reset_barrier(lmvn_dist* h) {
  LMVN_DIST_GUARD(h);
  return h->d.reset_barrier();
}

extern "C" int lmvn_dist_iterate(lmvn_dist* h, int iterations, double lambda, float min_value, float* device_ms) {
  LMVN_DIST_GUARD(h);
  return h->d.iterate(iterations, lambda, min_value, device_ms);
}

extern "C" int lmvn_dist_set_stream(lmvn_dist* h, void* cuda_stream) {
  LMVN_DIST_GUARD(h);
  DistDeconv& d = h->d;
  LMVN_CUDA_TRY(cudaSetDevice(d.device));
  LMVN_CUDA_TRY(cudaStreamSynchronize(d.stream));
  if (d.own_stream) LMVN_CUDA_TRY(cudaStreamDestroy(d.stream));
  d.stream = static_cast<cudaStream_t>(cuda_stream);
  d.own_stream = false;
  return 0;
}

extern "C" int lmvn_dist_set_staged(lmvn_dist* h, int on) {
  LMVN_DIST_GUARD(h);
  DistDeconv& d = h->d;
  LMVN_CUDA_TRY(cudaSetDevice(d.device));
  if (on && !d.stage_send) {
    const size_t bytes = std::max(d.slab_spec(), d.pencil_spec()) * sizeof(cplx);
    LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d.stage_send), bytes));
    LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d.stage_recv), bytes));
  }
  d.staged = on != 0;
  return 0;
}

extern "C" int lmvn_dist_buffer(lmvn_dist* h, int which, void** ptr, unsigned long long* bytes) {
  LMVN_DIST_GUARD(h);
  DistDeconv& d = h->d;
  if (!ptr || !bytes) { set_last_error("null output"); return -1; }
  switch (which) {
    case LMVN_DIST_SLAB_WORK: *ptr = d.slab_work(d.rank); *bytes = d.slab_spec() * sizeof(cplx); return 0;
    case LMVN_DIST_PENCIL_WORK: *ptr = d.pencil_work(d.rank); *bytes = d.pencil_spec() * sizeof(cplx); return 0;
    case LMVN_DIST_STAGE_SEND:
    case LMVN_DIST_STAGE_RECV:
      if (!d.stage_send) { set_last_error("staged exchange has not been enabled"); return -1; }
      *ptr = which == LMVN_DIST_STAGE_SEND ? d.stage_send : d.stage_recv;
      *bytes = std::max(d.slab_spec(), d.pencil_spec()) * sizeof(cplx);
      return 0;
    default: set_last_error("unknown buffer %d", which); return -1;
  }
}

extern "C" int lmvn_dist_synchronize(lmvn_dist* h) {
  LMVN_DIST_GUARD(h);
  LMVN_CUDA_TRY(cudaSetDevice(h->d.device));
  LMVN_CUDA_TRY(cudaStreamSynchronize(h->d.stream));
  return h->d.check_device_error();
}
