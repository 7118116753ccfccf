// fft_fused.cu -- host side of the power-of-two fast path (kernels: fft_fast.cuh).
//
// Strategy LMVN_STRATEGY_FUSED.  One convolution = five passes, every pointwise step fused into a pass edge:
//   rows_fwd             S -> C          real rows (or wrapped PSF) -> half spectrum along x
//   strided y forward    C -> C
//   strided z fwd*K^*inv C, K^ -> C      last forward stage, spectrum product and first inverse stage share registers
//   strided y inverse    C -> C
//   rows_inv             C (+S..) -> S   inverse along x + quotient / RL update
// Inside the RL loop (ConvEngine::chain_*) rows_inv of one convolution and rows_fwd of the next are ONE kernel
// (rows_inv_fwd), four launches per convolution.  Runtime knobs (environment, read when an engine is created) are
// A/B switches of measured design decisions, see DESIGN.md section 3.1:
//   LMVN_CHAIN=0           unchained loop
//   LMVN_PREFETCH=<blocks> y-forward look-ahead (148)  LMVN_PREFETCH_ROWS=0 no next-iteration row prefetch
//   LMVN_NXP_ALIGN=<n>     spectrum pitch alignment (8 complex elements)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "engine.cuh"
#include "fft_fast.cuh"
#include "fft_tma.cuh"
#include "fft_x3.cuh"

namespace lmvn {

namespace {

// twiddle tables + device facts, cached on the FftPlan (process-wide store): creating an engine for a
// shape that has been seen before makes no CUDA allocation and no synchronous copy
struct FastTables {
  int device = 0;
  int num_sms = 148;
  cplx* tw_m = nullptr;
  cplx* tw_nx = nullptr;
  cplx* tw_h = nullptr;  // nx = 1024: w_{M/2}
  cplx* tw_y[2] = {nullptr, nullptr};
  cplx* tw_z[2] = {nullptr, nullptr};
  bool y_alt = false;  // ny = 1024: the y passes run the 32 x 32 plan (fast::Plan<1024, 1>), tw_y holds its tables
  cplx* tw_p128[2] = {nullptr, nullptr};  // two-pass schedule: stage tables of the 128-point in-tile y transform
  cplx* tw_ny_lin = nullptr;              // two-pass schedule: w_ny^m, m < ny
  ~FastTables() {
    cudaSetDevice(device);
    for (int i = 0; i < 2; ++i)
      if (tw_p128[i]) cudaFree(tw_p128[i]);
    if (tw_ny_lin) cudaFree(tw_ny_lin);
    if (tw_h) cudaFree(tw_h);
    if (tw_m) cudaFree(tw_m);
    if (tw_nx) cudaFree(tw_nx);
    for (int i = 0; i < 2; ++i) {
      if (tw_y[i]) cudaFree(tw_y[i]);
      if (tw_z[i]) cudaFree(tw_z[i]);
    }
  }
};

// extents the two-pass schedule (fft_x3.cuh) takes: 128-row plane tiles of nx = 256, ny = 128 * R2 with R2 in {2, 4},
// two-stage z plans whose tiles are at most 32 columns wide
static bool x3_shape_ok(int nz, int ny, int nx) {
  return nx == 256 && (ny == 256 || ny == 512) && (nz == 64 || nz == 128 || nz == 256 || nz == 512);
}

struct FastEngine : ConvEngine, FastOps {
  std::shared_ptr<FastTables> tables;
  int M = 0, nxc = 0, nxp = 0;
  int num_sms = 148;
  // split layout: nx/2 columns per spectrum row + the Nyquist column as a compact plane behind them (see
  // fast::StridedArgs); false for the slab-decomposed engine, which scatters whole rows between devices
  bool split = true;
  int rows_ctas_per_sm = 8;
  int update_ctas_per_sm = 64;
  int y_fwd_prefetch = 148;  // blocks of look-ahead of the L2 prefetch in the forward y pass
  int y_inv_prefetch = 0, z_prefetch = 0;  // measured: no gain for these two
  int khat_prefetch = 0;  // measured: hurts the z pass (plane-strided lines), kept as a knob
  int rows_prefetch = 1;
  // OPT-IN (LMVN_TMA = mask of passes: 1 y forward, 2 y inverse, 4 merged z; 7 = all): strided passes of 512- and 256-point
  // axes as persistent kernels fed by the TMA engine (fft_tma.cuh).  Bit-identical to k_strided and measured SLOWER on B200
  // (config 3: y passes 0.112 against 0.093 ms, z pass 0.160 against 0.156 ms; 256^3: y 0.033 against 0.025 ms --
  // profiles/r02_tma_strided_probe.log): 128 KB of the SM's shared memory are tiles in work, what is left holds ONE 64 KB tile
  // in flight, less than the two co-resident CTAs of k_strided keep in flight in their registers.
  int use_tma = 0;
  cplx* d_tw_m = nullptr;
  cplx* d_tw_nx = nullptr;
  cplx* d_tw_h = nullptr;
  bool y_alt = false;
  // opt-in (LMVN_KHAT_FP16=1, single-device five-pass engine only; OUTSIDE the parity gate): the PSF spectra are stored as
  // __half2 scaled by 1 / max|component| -- the merged z pass then reads 2.5 C instead of 3 C.  Layout of a K^ buffer:
  // khat_elems() half2 values (same indexing as the float layout), then one float: the un-scale factor.
  bool khat_half = false;
  cplx* d_tw_y[2] = {nullptr, nullptr};  // per-stage tables of the y / z passes
  cplx* d_tw_z[2] = {nullptr, nullptr};

  ~FastEngine() override {}
  int strategy() const override { return 2; }
  size_t main_elems() const { return size_t(plan->nz) * plan->ny * nxp; }
  size_t khat_elems() const override { return main_elems() + (split ? size_t(plan->nz) * plan->ny : 0); }
  cplx* nyq_of(cplx* spec) const { return split ? spec + main_elems() : nullptr; }
  const cplx* nyq_of(const cplx* spec) const { return split ? spec + main_elems() : nullptr; }
  size_t work_elems() const override { return khat_elems(); }
  int launches_per_conv() const override { return can_chain() ? 4 : 5; }
  // nx = 1024 with LMVN_CHAIN_WIDE_UPDATE=0: the update link is two launches (rows_inv_fwd below)
  int launches_per_view_iteration() const override {
    return 2 * launches_per_conv() + ((can_chain() && M == 512 && !chain_wide_update) ? 1 : 0);
  }
  unsigned long long S() const { return plan->voxels() * sizeof(float); }
  unsigned long long C() const { return plan->spec_elems() * sizeof(cplx); }

  static int upload_table(cplx** dst, int period, int count) {
    std::vector<cplx> h(count);
    for (int k = 0; k < count; ++k) {
      const double a = -2.0 * M_PI * double(k) / double(period);
      h[k] = cmake(float(cos(a)), float(sin(a)));
    }
    LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(dst), sizeof(cplx) * count));
    LMVN_CUDA_TRY(cudaMemcpy(*dst, h.data(), sizeof(cplx) * count, cudaMemcpyHostToDevice));
    return 0;
  }

  int init() {
    M = plan->nx / 2;
    nxc = plan->nxc;
    {
      // complex elements.  8 = rows start on 64-byte boundaries: every 16-column tile row is still four full
      // 32-byte sectors, and the pitch (129 -> 136 instead of 144) moves 5.5 % fewer spectrum bytes (+1.2 % measured)
      int align = 8;
      if (const char* e = getenv("LMVN_NXP_ALIGN")) align = std::max(1, atoi(e));
      if (const char* e = getenv("LMVN_SPLIT_NYQUIST")) split = split && (*e != '0');
      nxp = split ? M : (nxc + align - 1) / align * align;
    }
    if (const char* e = getenv("LMVN_PREFETCH")) y_fwd_prefetch = std::max(0, atoi(e));
    if (const char* e = getenv("LMVN_PREFETCH_YINV")) y_inv_prefetch = std::max(0, atoi(e));
    if (const char* e = getenv("LMVN_PREFETCH_Z")) z_prefetch = std::max(0, atoi(e));
    if (const char* e = getenv("LMVN_PREFETCH_KHAT")) khat_prefetch = atoi(e);
    if (const char* e = getenv("LMVN_PREFETCH_ROWS")) rows_prefetch = atoi(e);
    if (const char* e = getenv("LMVN_TMA")) use_tma = atoi(e);
    if (const char* e = getenv("LMVN_CHAIN")) chain_ok = (*e != '0');
    if (const char* e = getenv("LMVN_CHAIN_WIDE")) chain_wide = (*e != '0');
    if (const char* e = getenv("LMVN_CHAIN_WIDE_UPDATE")) chain_wide_update = (*e != '0');
    if (const char* e = getenv("LMVN_ROWS_CTAS")) rows_ctas_per_sm = update_ctas_per_sm = std::max(1, atoi(e));
    {
      std::lock_guard<std::mutex> lk(plan->fast_mu);
      if (!plan->fast_tables) {
        std::shared_ptr<FastTables> t(new FastTables());
        t->device = plan->device;
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device) == cudaSuccess && sms > 0)
          t->num_sms = sms;
        LMVN_TRY(upload_table(&t->tw_m, M, M));
        LMVN_TRY(upload_table(&t->tw_nx, plan->nx, M + 1));
        if (M == 512) LMVN_TRY(upload_table(&t->tw_h, M / 2, M / 2));
        t->y_alt = (plan->ny == 1024);
        if (const char* e = getenv("LMVN_Y1024_WIDE")) t->y_alt = t->y_alt && (*e != '0');
        LMVN_TRY(upload_stage_tables(t->tw_y, plan->ny, t->y_alt));
        LMVN_TRY(upload_stage_tables(t->tw_z, plan->nz, false));
        if (x3_shape_ok(plan->nz, plan->ny, plan->nx)) {
          LMVN_TRY(upload_stage_tables(t->tw_p128, 128, false));
          LMVN_TRY(upload_table(&t->tw_ny_lin, plan->ny, plan->ny));
        }
        plan->fast_tables = t;
      }
      tables = std::static_pointer_cast<FastTables>(plan->fast_tables);
    }
    num_sms = tables->num_sms;
    y_alt = tables->y_alt;
    d_tw_m = tables->tw_m; d_tw_nx = tables->tw_nx; d_tw_h = tables->tw_h;
    for (int i = 0; i < 2; ++i) { d_tw_y[i] = tables->tw_y[i]; d_tw_z[i] = tables->tw_z[i]; }
    return 0;
  }

  // [j][q] = w_L^{jq} for the first two stages of the radix plan of length n
  static int upload_stage_tables(cplx** dst, int n, bool alt) {
    int r1, r2;
    switch (n) {
      case 1024:
        r1 = alt ? fast::Plan<1024, 1>::R1 : fast::Radix<1024>::R1;
        r2 = alt ? fast::Plan<1024, 1>::R2 : fast::Radix<1024>::R2;
        break;
      case 512: r1 = fast::Radix<512>::R1; r2 = fast::Radix<512>::R2; break;
      case 256: r1 = fast::Radix<256>::R1; r2 = fast::Radix<256>::R2; break;
      case 128: r1 = fast::Radix<128>::R1; r2 = fast::Radix<128>::R2; break;
      case 64: r1 = fast::Radix<64>::R1; r2 = fast::Radix<64>::R2; break;
      case 32: r1 = fast::Radix<32>::R1; r2 = fast::Radix<32>::R2; break;
      case 16: r1 = fast::Radix<16>::R1; r2 = fast::Radix<16>::R2; break;
      default: set_last_error("fused path: unsupported axis length %d", n); return -1;
    }
    const int spans[2] = {n, n / r1};
    const int radix[2] = {r1, r2};
    for (int s = 0; s < 2; ++s) {
      const int L = spans[s], R = radix[s], M = L / R;
      std::vector<cplx> h(size_t(M) * R);
      for (int j = 0; j < M; ++j)
        for (int q = 0; q < R; ++q) {
          const double a = -2.0 * M_PI * double(j) * double(q) / double(L);
          h[size_t(j) * R + q] = cmake(float(cos(a)), float(sin(a)));
        }
      LMVN_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&dst[s]), sizeof(cplx) * h.size()));
      LMVN_CUDA_TRY(cudaMemcpy(dst[s], h.data(), sizeof(cplx) * h.size(), cudaMemcpyHostToDevice));
    }
    return 0;
  }

  // ---- launches -------------------------------------------------------------
  // second-generation rows kernels (full-line accesses, persistent row loop)
  template <int MM>
  int launch_rows_fwd2(const fast::RowArgs& a, bool wrapped, cudaStream_t s) {
    typedef fast::Row2Cfg<MM> CF;
    const size_t rows = size_t(a.nz) * plan->ny;
    const int groups = fast::kLinkThreads / 16;
    const size_t iters = ceil_div(rows, size_t(groups) * CF::RPG);
    const dim3 grid(unsigned(std::min<size_t>(iters, size_t(num_sms) * rows_ctas_per_sm * (fast::kRowThreads / fast::kLinkThreads))));
    const size_t smem = size_t(groups) * CF::RPG * CF::RS * sizeof(cplx);
    auto kw = fast::k_rows_fwd2<MM, true>;
    auto kp = fast::k_rows_fwd2<MM, false>;
    if (wrapped) {
      LMVN_LAUNCH(kw, grid, dim3(fast::kLinkThreads), smem, s, a);
    } else {
      LMVN_LAUNCH(kp, grid, dim3(fast::kLinkThreads), smem, s, a);
    }
    return 0;
  }
  template <int MM>
  int launch_rows_inv_fwd(const fast::RowArgs& a, cudaStream_t s) {
    typedef fast::Row2Cfg<MM> CF;
    const size_t rows = size_t(a.nz) * plan->ny;
    const int groups = fast::kLinkThreads / 16;
    const size_t iters = ceil_div(rows, size_t(groups) * CF::RPG);
    // measured: the update link (four streams per row) runs best with one loop iteration per CTA, i.e. left to the
    // hardware CTA scheduler (0.268 -> 0.242 ms on config 3); the quotient link with a persistent loop of ~7 iterations
    const int per_sm = (a.ep.mode == gen::EPI_UPDATE) ? update_ctas_per_sm : rows_ctas_per_sm;
    const dim3 grid(unsigned(std::min<size_t>(iters, size_t(num_sms) * per_sm * (fast::kRowThreads / fast::kLinkThreads))));
    const size_t smem = size_t(groups) * CF::RPG * CF::RS * sizeof(cplx);
    auto k1 = fast::k_rows_inv_fwd<MM, gen::EPI_QUOTIENT>;
    auto k2 = fast::k_rows_inv_fwd<MM, gen::EPI_UPDATE>;
    auto e1 = fast::k_rows_inv_fwd<MM, gen::EPI_QUOTIENT, true>;
    auto e2 = fast::k_rows_inv_fwd<MM, gen::EPI_UPDATE, true>;
    const bool emb = a.spec_out != nullptr;
    if (a.ep.mode == gen::EPI_QUOTIENT) {
      if (emb) { LMVN_LAUNCH(e1, grid, dim3(fast::kLinkThreads), smem, s, a); }
      else { LMVN_LAUNCH(k1, grid, dim3(fast::kLinkThreads), smem, s, a); }
    } else {
      if (emb) { LMVN_LAUNCH(e2, grid, dim3(fast::kLinkThreads), smem, s, a); }
      else { LMVN_LAUNCH(k2, grid, dim3(fast::kLinkThreads), smem, s, a); }
    }
    return 0;
  }

  // nx = 1024
  int launch_rows_fwd_wide(const fast::RowArgs& a, bool wrapped, cudaStream_t s) {
    const size_t rows = size_t(a.nz) * plan->ny;
    const dim3 grid(unsigned(std::min<size_t>(ceil_div(rows, fast::RowWide::ROWS), size_t(num_sms) * rows_ctas_per_sm)));
    const size_t smem = fast::RowWide::SMEM;
    auto kw = fast::k_rows_fwd_wide<true>;
    auto kp = fast::k_rows_fwd_wide<false>;
    LMVN_CUDA_TRY(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    LMVN_CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    if (wrapped) {
      LMVN_LAUNCH(kw, grid, dim3(fast::kRowThreads), smem, s, a);
    } else {
      LMVN_LAUNCH(kp, grid, dim3(fast::kRowThreads), smem, s, a);
    }
    return 0;
  }
  int launch_rows_inv_wide(const fast::RowArgs& a, cudaStream_t s) {
    const size_t rows = size_t(a.nz) * plan->ny;
    const dim3 grid(unsigned(std::min<size_t>(ceil_div(rows, fast::RowWide::ROWS), size_t(num_sms) * rows_ctas_per_sm)));
    const size_t smem = fast::RowWide::SMEM;
    auto k0 = fast::k_rows_inv_wide<gen::EPI_STORE>;
    auto k1 = fast::k_rows_inv_wide<gen::EPI_QUOTIENT>;
    auto k2 = fast::k_rows_inv_wide<gen::EPI_UPDATE>;
    LMVN_CUDA_TRY(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    LMVN_CUDA_TRY(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    LMVN_CUDA_TRY(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    switch (a.ep.mode) {
      case gen::EPI_QUOTIENT: LMVN_LAUNCH(k1, grid, dim3(fast::kRowThreads), smem, s, a); break;
      case gen::EPI_UPDATE: LMVN_LAUNCH(k2, grid, dim3(fast::kRowThreads), smem, s, a); break;
      default: LMVN_LAUNCH(k0, grid, dim3(fast::kRowThreads), smem, s, a); break;
    }
    return 0;
  }

  template <int MM>
  int launch_rows_inv2(const fast::RowArgs& a, cudaStream_t s) {
    typedef fast::Row2Cfg<MM> CF;
    const size_t rows = size_t(a.nz) * plan->ny;
    const int groups = fast::kLinkThreads / 16;
    const size_t iters = ceil_div(rows, size_t(groups) * CF::RPG);
    const dim3 grid(unsigned(std::min<size_t>(iters, size_t(num_sms) * rows_ctas_per_sm * (fast::kRowThreads / fast::kLinkThreads))));
    const size_t smem = size_t(groups) * CF::RPG * CF::RS * sizeof(cplx);
    auto k0 = fast::k_rows_inv2<MM, gen::EPI_STORE>;
    auto k1 = fast::k_rows_inv2<MM, gen::EPI_QUOTIENT>;
    auto k2 = fast::k_rows_inv2<MM, gen::EPI_UPDATE>;
    switch (a.ep.mode) {
      case gen::EPI_QUOTIENT: LMVN_LAUNCH(k1, grid, dim3(fast::kLinkThreads), smem, s, a); break;
      case gen::EPI_UPDATE: LMVN_LAUNCH(k2, grid, dim3(fast::kLinkThreads), smem, s, a); break;
      default: LMVN_LAUNCH(k0, grid, dim3(fast::kLinkThreads), smem, s, a); break;
    }
    return 0;
  }

  // z0/nzs select a slab of planes (whole volume by default); wrapped sources are whole-volume only
  int rows_fwd(const gen::RealSource& src_in, cplx* spec, cudaStream_t s, int z0 = 0, int nzs = -1, int wrap_z0 = 0,
               int wrap_nz = 0) {
    if (nzs < 0) nzs = plan->nz;
    fast::RowArgs a;
    std::memset(&a, 0, sizeof(a));
    gen::RealSource src = src_in;
    if (!src.wrapped) src.data += size_t(z0) * plan->ny * plan->nx;
    a.src = src;
    a.spec = spec + size_t(z0) * plan->ny * nxp;
    a.nz = nzs; a.ny = plan->ny; a.nxp = nxp;
    a.tw_m = d_tw_m; a.tw_nx = d_tw_nx;
    a.prefetch = rows_prefetch;
    a.z0 = wrap_z0; a.nz_wrap = wrap_nz > 0 ? wrap_nz : plan->nz;
    a.tw_h = d_tw_h;
    a.nyq = split ? nyq_of(spec) + size_t(z0) * plan->ny : nullptr;
    const bool w = src.wrapped != 0;
    switch (M) {
      case 32: LMVN_TRY(launch_rows_fwd2<32>(a, w, s)); break;
      case 64: LMVN_TRY(launch_rows_fwd2<64>(a, w, s)); break;
      case 128: LMVN_TRY(launch_rows_fwd2<128>(a, w, s)); break;
      case 256: LMVN_TRY(launch_rows_fwd2<256>(a, w, s)); break;
      case 512: LMVN_TRY(launch_rows_fwd_wide(a, w, s)); break;
      default: set_last_error("fused path: unsupported nx"); return -1;
    }
    LMVN_CUDA_TRY(cudaGetLastError());
    mark("fast_rows_fwd", S() + C(), s);
    return 0;
  }

  int rows_inv(const cplx* spec, float* out, const gen::Epilogue& ep_in, cudaStream_t s, int z0 = 0, int nzs = -1) {
    if (nzs < 0) nzs = plan->nz;
    fast::RowArgs a;
    std::memset(&a, 0, sizeof(a));
    const size_t voff = size_t(z0) * plan->ny * plan->nx;
    gen::Epilogue ep = ep_in;
    if (ep.scale != 1.f) {
      set_last_error("fast path: the 1/N scale lives in the PSF spectra, epilogue scale must be 1");
      return -1;
    }
    if (ep.view) ep.view += voff;
    if (ep.psi) ep.psi += voff;
    if (ep.weights) ep.weights += voff;
    a.spec = const_cast<cplx*>(spec) + size_t(z0) * plan->ny * nxp;
    a.out = out ? out + voff : out;
    a.ep = ep;
    a.nz = nzs; a.ny = plan->ny; a.nxp = nxp;
    a.tw_m = d_tw_m; a.tw_nx = d_tw_nx;
    a.prefetch = rows_prefetch;
    a.tw_h = d_tw_h;
    a.nyq = split ? const_cast<cplx*>(nyq_of(spec)) + size_t(z0) * plan->ny : nullptr;
    switch (M) {
      case 32: LMVN_TRY(launch_rows_inv2<32>(a, s)); break;
      case 64: LMVN_TRY(launch_rows_inv2<64>(a, s)); break;
      case 128: LMVN_TRY(launch_rows_inv2<128>(a, s)); break;
      case 256: LMVN_TRY(launch_rows_inv2<256>(a, s)); break;
      case 512: LMVN_TRY(launch_rows_inv_wide(a, s)); break;
      default: set_last_error("fused path: unsupported nx"); return -1;
    }
    LMVN_CUDA_TRY(cudaGetLastError());
    mark(ep.mode == gen::EPI_UPDATE ? "fast_rows_inv_update"
                                    : (ep.mode == gen::EPI_QUOTIENT ? "fast_rows_inv_quotient" : "fast_rows_inv"),
         C() + S() * (ep.mode == gen::EPI_UPDATE ? 3 : (ep.mode == gen::EPI_QUOTIENT ? 2 : 1)), s);
    return 0;
  }

  template <int N, int MODE, int ALT = 0>
  int launch_strided(const fast::StridedArgs& a, dim3 grid, cudaStream_t s) {
    constexpr int COLS = fast::TileCols<N, MODE>::V;
    const size_t smem = size_t(N) * COLS * sizeof(cplx);
    auto kfn = fast::k_strided<N, MODE, ALT>;
    if (smem > 48 * 1024) {  // per device, cheap: set every time
      LMVN_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    }
    LMVN_LAUNCH(kfn, grid, dim3(fast::TileThreads<N, MODE>::V), smem, s, a);
    return 0;
  }
  template <int N>
  int launch_strided_khalf(const fast::StridedArgs& a, dim3 grid, cudaStream_t s) {
#ifndef LMVN_EMU
    constexpr int COLS = fast::TileCols<N, fast::SM_FWD_MUL_INV>::V;
    const size_t smem = size_t(N) * COLS * sizeof(cplx);
    auto kfn = fast::k_strided<N, fast::SM_FWD_MUL_INV, 0, 1>;
    if (smem > 48 * 1024) LMVN_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    LMVN_LAUNCH(kfn, grid, dim3(fast::TileThreads<N, fast::SM_FWD_MUL_INV>::V), smem, s, a);
    return 0;
#else
    (void)a; (void)grid; (void)s;
    set_last_error("half-precision PSF spectra are not available in the emulated build");
    return -1;
#endif
  }
  template <int N>
  int launch_strided_mode(const fast::StridedArgs& a, int mode, dim3 grid, cudaStream_t s) {
    if (a.khat_unscale && mode == fast::SM_FWD_MUL_INV) return launch_strided_khalf<N>(a, grid, s);
    switch (mode) {
      case fast::SM_FWD: return launch_strided<N, fast::SM_FWD>(a, grid, s);
      case fast::SM_INV: return launch_strided<N, fast::SM_INV>(a, grid, s);
      case fast::SM_FWD_MUL_INV: return launch_strided<N, fast::SM_FWD_MUL_INV>(a, grid, s);
      case fast::SM_FWD_SCATTER: return launch_strided<N, fast::SM_FWD_SCATTER>(a, grid, s);
      case fast::SM_FWD_MUL_INV_SCATTER: return launch_strided<N, fast::SM_FWD_MUL_INV_SCATTER>(a, grid, s);
      default: return launch_strided<N, fast::SM_FWD_SCALE>(a, grid, s);
    }
  }
  template <int N>
  int launch_strided_axis(const fast::StridedArgs& a, int mode, dim3 grid, cudaStream_t s, bool alt) {
    if constexpr (N == 1024) {  // the only length with an alternate plan: nothing else is instantiated twice
      if (alt) return launch_strided_alt<N>(a, mode, grid, s);
    }
    (void)alt;
    return launch_strided_mode<N>(a, mode, grid, s);
  }
  // passes of an axis on the alternate plan (never the merged pass)
  template <int N>
  int launch_strided_alt(const fast::StridedArgs& a, int mode, dim3 grid, cudaStream_t s) {
    switch (mode) {
      case fast::SM_FWD: return launch_strided<N, fast::SM_FWD, 1>(a, grid, s);
      case fast::SM_INV: return launch_strided<N, fast::SM_INV, 1>(a, grid, s);
      case fast::SM_FWD_SCATTER: return launch_strided<N, fast::SM_FWD_SCATTER, 1>(a, grid, s);
      case fast::SM_FWD_SCALE: return launch_strided<N, fast::SM_FWD_SCALE, 1>(a, grid, s);
      default: set_last_error("strided pass: mode %d has no alternate plan", mode); return -1;
    }
  }

#ifndef LMVN_EMU
  // ---- persistent TMA-fed strided passes (fft_tma.cuh) ----
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
          q != cudaDriverEntryPointSuccess)
        return nullptr;
      return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
  }
  // tensor map [slow][row][2 M floats] of a spectrum in the split layout; box = 16 complex columns x up to 256 rows
  int spectrum_tensor_map(CUtensorMap* map, cplx* data, int n, long long row_stride, long long tile_stride, unsigned slow) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_last_error("cuTensorMapEncodeTiled is not available"); return -1; }
    const cuuint64_t gdim[3] = {cuuint64_t(2 * M), cuuint64_t(n), cuuint64_t(slow)};
    const cuuint64_t gstr[2] = {cuuint64_t(row_stride) * sizeof(cplx), cuuint64_t(tile_stride) * sizeof(cplx)};
    const cuuint32_t box[3] = {32, cuuint32_t(n < 256 ? n : 256), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, data, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed (%d)", int(r)); return -1; }
    return 0;
  }
  bool tma_shape_ok(const fast::StridedArgs& a, const StridedGeom& g) const {
    const int bit = (g.mode == fast::SM_FWD) ? 1 : (g.mode == fast::SM_INV ? 2 : 4);  // LMVN_TMA is a mask of passes
    if (!(use_tma & bit) || !g.nyq || g.ncols > 0 || g.khat_half || (y_alt && g.tw_axis == 1)) return false;
    if (g.mode != fast::SM_FWD && g.mode != fast::SM_INV && g.mode != fast::SM_FWD_MUL_INV) return false;
    if (M % 16 != 0 || a.ncols != M || !encode_tiled()) return false;
    return g.n == 512 || g.n == 256;
  }
  template <int N, int MODE>
  int launch_strided_tma(const fast::StridedArgs& a, const CUtensorMap& map, long long n_tiles, cudaStream_t s) {
    auto kfn = tma::k_strided_tma<N, MODE>;
    const size_t smem = tma::Cfg<N>::SMEM;
    LMVN_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const long long items = n_tiles + a.nyq_groups;
    const dim3 grid(unsigned(std::min<long long>(num_sms, (items + 1) / 2)));
    static const bool dbg = getenv("LMVN_TMA_SYNC") != nullptr;  // debugging aid: localise a failing launch
    if (dbg) {
      const cudaError_t e0 = cudaStreamSynchronize(s);
      if (e0 != cudaSuccess) { set_last_error("before TMA pass N=%d mode=%d: %s", N, MODE, cudaGetErrorString(e0)); return -1; }
    }
    LMVN_LAUNCH(kfn, grid, dim3(tma::kThreads), smem, s, map, a, int(n_tiles));
    if (dbg) {
      const cudaError_t e1 = cudaStreamSynchronize(s);
      fprintf(stderr, "[tma] N=%d mode=%d data=%p khat=%p rs=%d ts=%lld slow=%u tiles=%d grid=%u: %s\n", N, MODE, (void*)a.data,
              (const void*)a.khat, a.row_stride, a.tile_stride, a.slow, int(n_tiles), grid.x, cudaGetErrorString(e1));
      if (e1 != cudaSuccess) { set_last_error("TMA pass N=%d mode=%d: %s", N, MODE, cudaGetErrorString(e1)); return -1; }
    }
    return 0;
  }
  template <int N>
  int strided_tma(fast::StridedArgs a, const StridedGeom& g, cudaStream_t s) {
    a.tiles_x = M / 16;
    a.nyq_groups = int(ceil_div(size_t(g.slow), size_t(16)));
    a.prefetch = 0;
    a.prefetch_khat = 0;
    CUtensorMap map;
    LMVN_TRY(spectrum_tensor_map(&map, a.data, N, a.row_stride, a.tile_stride, g.slow));
    const long long n_tiles = (long long)a.tiles_x * g.slow;
    switch (g.mode) {
      case fast::SM_FWD: return launch_strided_tma<N, fast::SM_FWD>(a, map, n_tiles, s);
      case fast::SM_INV: return launch_strided_tma<N, fast::SM_INV>(a, map, n_tiles, s);
      default: return launch_strided_tma<N, fast::SM_FWD_MUL_INV>(a, map, n_tiles, s);
    }
  }
#endif

  // axis 1 = y, axis 0 = z
  int strided(cplx* data, const cplx* khat, int axis, int mode, float scale, cudaStream_t s, int z0 = 0,
              int nzs = -1) {
    const FftPlan& p = *plan;
    StridedGeom g;
    if (split) {
      g.nyq = nyq_of(data) + ((axis == 1 && nzs >= 0) ? size_t(z0) * p.ny : 0);
      g.nyq_khat = khat ? nyq_of(khat) : nullptr;
      if (khat && khat_half && mode == fast::SM_FWD_MUL_INV)  // half2 elements: same element offset, half the bytes
        g.nyq_khat = reinterpret_cast<const cplx*>(reinterpret_cast<const unsigned*>(khat) + main_elems());
    }
    g.khat_half = khat && khat_half && mode == fast::SM_FWD_MUL_INV;
    if (axis == 1 && nzs >= 0) data += size_t(z0) * p.ny * nxp;  // y pass on a slab of planes
    g.data = data;
    g.khat = khat;
    g.mode = mode;
    g.scale = scale;
    g.tw_axis = axis;
    if (axis == 1) {
      g.n = p.ny; g.row_stride = nxp; g.tile_stride = (long long)p.ny * nxp; g.slow = unsigned(nzs >= 0 ? nzs : p.nz);
    } else {
      g.n = p.nz; g.row_stride = p.ny * nxp; g.tile_stride = nxp; g.slow = unsigned(p.ny);
    }
    return strided_geom(g, s);
  }

  int strided_tile_cols(int n) const override {  // the widest tile any pass of an axis of this length uses
    switch (n) {
      case 1024: return std::max(fast::Cols<1024>::V, fast::TileCols<1024, fast::SM_FWD_MUL_INV>::V);
      case 512: return fast::Cols<512>::V;
      case 256: return fast::Cols<256>::V;
      case 128: return fast::Cols<128>::V;
      default: return fast::Cols<64>::V;
    }
  }

  int strided_geom(const StridedGeom& g, cudaStream_t s) override {
    fast::StridedArgs a;
    std::memset(&a, 0, sizeof(a));
    a.data = g.data;
    a.khat = g.khat;
    a.ncols = g.nyq ? M : nxc;
    int win_cols = nxc;  // columns the grid covers
    if (g.ncols > 0) {
      if (g.nyq || g.col0 % strided_tile_cols(g.n) != 0 || g.col0 < 0 || g.col0 >= nxc) {
        set_last_error("strided pass: invalid column window");
        return -1;
      }
      win_cols = std::min(g.ncols, nxc - g.col0);
      a.data = g.data + g.col0;
      if (g.khat) a.khat = g.khat + g.col0;
      a.ncols = win_cols;
    }
    a.scale = g.scale;
    a.row_stride = g.row_stride;
    a.tile_stride = g.tile_stride;
    a.sc = g.sc;
    if (g.ncols > 0) a.sc.offset += g.col0;
    a.nyq_groups = -1;
    a.slow = g.slow;
    if (g.nyq) {
      // Nyquist plane nyq[z'][y']: the y pass walks it along y' (row stride 1) with z as the slow index, the z
      // pass along z (row stride ny) with y' as the slow index
      a.nyq = g.nyq;
      a.nyq_khat = g.nyq_khat;
      a.nyq_rs = (g.tw_axis == 1) ? 1 : plan->ny;
      a.nyq_cs = (g.tw_axis == 1) ? plan->ny : 1;
    }
    const int mode = g.mode;
    a.prefetch = (g.tw_axis == 1 && mode == fast::SM_FWD) ? y_fwd_prefetch
                 : (g.tw_axis == 1 && mode == fast::SM_INV) ? y_inv_prefetch
                 : (mode == fast::SM_FWD_MUL_INV) ? z_prefetch : 0;
    a.prefetch_khat = khat_prefetch;
    if (g.khat_half) {
      a.khat_unscale = reinterpret_cast<const float*>(reinterpret_cast<const unsigned*>(g.khat) + khat_elems());
      a.prefetch_khat = 0;
    }
    if (g.tw_axis == 1) { a.tw1 = d_tw_y[0]; a.tw2 = d_tw_y[1]; }
    else { a.tw1 = d_tw_z[0]; a.tw2 = d_tw_z[1]; }
    if (g.n != (g.tw_axis == 1 ? plan->ny : plan->nz)) {
      set_last_error("strided pass: length %d does not match the plan's axis", g.n);
      return -1;
    }
    const unsigned slow = g.slow;
    int rc;
#ifndef LMVN_EMU
    if (tma_shape_ok(a, g)) {
      if (g.n == 512) LMVN_TRY(strided_tma<512>(a, g, s));
      else LMVN_TRY(strided_tma<256>(a, g, s));
      LMVN_CUDA_TRY(cudaGetLastError());
      if (mode == fast::SM_FWD_MUL_INV) mark("fast_z_mul", 3 * C(), s);
      else mark(g.tw_axis == 1 ? (mode == fast::SM_INV ? "fast_y_inv" : "fast_y_fwd") : "fast_z", 2 * C(), s);
      return 0;
    }
#endif
#define LMVN_STRIDED_CASE(NN)                                                                   \
  case NN: {                                                                                    \
    const bool wide = (mode == fast::SM_FWD_MUL_INV || mode == fast::SM_FWD_MUL_INV_SCATTER ||  \
                       mode == fast::SM_FWD_SCATTER);                                           \
    const size_t tcols = wide ? size_t(fast::TileCols<NN, fast::SM_FWD_MUL_INV>::V) : size_t(fast::Cols<NN>::V); \
    dim3 grid(unsigned(ceil_div(size_t(win_cols), tcols)), slow);                               \
    if (g.nyq) {                                                                                \
      a.tiles_x = int(ceil_div(size_t(M), tcols));                                              \
      a.nyq_groups = int(ceil_div(size_t(slow), tcols));                                        \
      grid = dim3(unsigned(a.nyq_groups) + unsigned(a.tiles_x) * slow);                         \
    }                                                                                           \
    rc = launch_strided_axis<NN>(a, mode, grid, s, y_alt && g.tw_axis == 1);                    \
  } break;
    switch (g.n) {
      LMVN_STRIDED_CASE(16)
      LMVN_STRIDED_CASE(32)
      LMVN_STRIDED_CASE(64)
      LMVN_STRIDED_CASE(128)
      LMVN_STRIDED_CASE(256)
      LMVN_STRIDED_CASE(512)
      LMVN_STRIDED_CASE(1024)
      default: set_last_error("fused path: unsupported axis length %d", g.n); return -1;
    }
#undef LMVN_STRIDED_CASE
    if (rc != 0) return rc;
    LMVN_CUDA_TRY(cudaGetLastError());
    const bool zmul = (mode == fast::SM_FWD_MUL_INV || mode == fast::SM_FWD_MUL_INV_SCATTER);
    if (zmul) mark("fast_z_mul", 3 * C(), s);
    else mark(g.tw_axis == 1 ? (mode == fast::SM_INV ? "fast_y_inv" : "fast_y_fwd") : "fast_z", 2 * C(), s);
    return 0;
  }

  // ---- FastOps (slab-decomposed engine) ----
  int nxp_pitch() const override { return nxp; }
  int rows_fwd_planes(const gen::RealSource& src, cplx* spec, int nz_local, int z0_global, int nz_global,
                      cudaStream_t s) override {
    // spec / src.data already point at the first local plane
    return rows_fwd(src, spec, s, 0, nz_local, z0_global, nz_global);
  }
  int rows_inv_planes(const cplx* spec, float* out, const gen::Epilogue& ep, int nz_local, cudaStream_t s) override {
    return rows_inv(spec, out, ep, s, 0, nz_local);
  }
  bool can_chain_rows() const override { return chain_ok && (M <= 256 || chain_wide); }
  int rows_inv_fwd_planes(cplx* spec, const gen::Epilogue& ep, int nz_local, cudaStream_t s) override {
    return rows_inv_fwd(spec, ep, s, nz_local);
  }

  // x inverse + pointwise + x forward of the result, in place on the rows of `spec`
  int rows_inv_fwd(cplx* spec, const gen::Epilogue& ep, cudaStream_t s, int nzs = -1, cplx* spec_out = nullptr,
                   const int* logical = nullptr, const int* offset = nullptr, float* psi_out = nullptr) {
    if (ep.scale != 1.f || (ep.mode != gen::EPI_QUOTIENT && ep.mode != gen::EPI_UPDATE)) {
      set_last_error("chained rows pass: quotient or update epilogue with unit scale expected");
      return -1;
    }
    if (M == 512 && ep.mode == gen::EPI_UPDATE && !spec_out && !chain_wide_update) {
      // A/B (LMVN_CHAIN_WIDE_UPDATE=0): the update link of nx = 1024 as the two passes it fuses.  That was the default while a
      // 16-lane group held a whole row (170 registers, 6.38 ms chained against 4.08 + 1.44 ms at 1024^3); with one warp per
      // row the chained kernel fits 128 registers: 3.91 ms against 2.89 + 1.52 ms (profiles/r02_1024_wide_tiles.log).
      LMVN_TRY(rows_inv(spec, ep.psi, ep, s, 0, nzs));
      gen::RealSource src{ep.psi, 0, 0, 0, 0};
      return rows_fwd(src, spec, s, 0, nzs);
    }
    fast::RowArgs a;
    std::memset(&a, 0, sizeof(a));
    a.spec = spec;
    a.ep = ep;
    a.nz = nzs >= 0 ? nzs : plan->nz; a.ny = plan->ny; a.nxp = nxp;
    a.tw_m = d_tw_m; a.tw_nx = d_tw_nx; a.tw_h = d_tw_h;
    a.prefetch = rows_prefetch;
    a.nyq = split ? nyq_of(spec) : nullptr;
    if (spec_out) {
      if (M > 256 || !logical || !offset) { set_last_error("chained embedded rows pass: unsupported"); return -1; }
      a.spec_out = spec_out;
      a.nyq_out = split ? nyq_of(spec_out) : nullptr;
      a.out = psi_out;
      a.lz = logical[0]; a.ly = logical[1]; a.lx = logical[2];
      a.oz = offset[0]; a.oy = offset[1]; a.ox = offset[2];
      a.prefetch = 0;  // the rows read are not the rows of the next loop iteration
    }
    switch (M) {
      case 32: LMVN_TRY(launch_rows_inv_fwd<32>(a, s)); break;
      case 64: LMVN_TRY(launch_rows_inv_fwd<64>(a, s)); break;
      case 128: LMVN_TRY(launch_rows_inv_fwd<128>(a, s)); break;
      case 256: LMVN_TRY(launch_rows_inv_fwd<256>(a, s)); break;
      case 512: {
        const size_t rows = size_t(a.nz) * plan->ny;
        const int groups = fast::kChainWideThreads / 32;
        const dim3 grid(unsigned(ceil_div(rows, size_t(groups))));  // one row per warp, no loop in the kernel
        const size_t smem = size_t(groups) * fast::RowWide::SLAB * sizeof(cplx);
        auto k1 = fast::k_rows_inv_fwd_wide<gen::EPI_QUOTIENT>;
        auto k2 = fast::k_rows_inv_fwd_wide<gen::EPI_UPDATE>;
        LMVN_CUDA_TRY(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        LMVN_CUDA_TRY(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        if (a.ep.mode == gen::EPI_QUOTIENT) {
          LMVN_LAUNCH(k1, grid, dim3(fast::kChainWideThreads), smem, s, a);
        } else {
          LMVN_LAUNCH(k2, grid, dim3(fast::kChainWideThreads), smem, s, a);
        }
      } break;
      default: set_last_error("chained rows pass: unsupported nx"); return -1;
    }
    LMVN_CUDA_TRY(cudaGetLastError());
    // algorithmic bytes: spectrum in and out, operands in, psi out for the update
    mark(ep.mode == gen::EPI_UPDATE ? "fast_rows_inv_update_fwd" : "fast_rows_inv_quotient_fwd",
         2 * C() + S() * (ep.mode == gen::EPI_UPDATE ? 3 : 1), s);
    return 0;
  }

  bool chain_ok = true;
  // nx = 1024: chained links too (one warp per row, 128 registers; LMVN_CHAIN_WIDE=0 / LMVN_CHAIN_WIDE_UPDATE=0 for A/B)
  bool chain_wide = true;
  bool chain_wide_update = true;
  bool can_chain() const override { return chain_ok && (M <= 256 || chain_wide); }
  int chain_begin(const float* in, cplx* work, cudaStream_t s) override {
    gen::RealSource src{in, 0, 0, 0, 0};
    return rows_fwd(src, work, s);
  }
  int chain_middle(cplx* work, const cplx* khat, cudaStream_t s) override {
    LMVN_TRY(strided(work, nullptr, 1, fast::SM_FWD, 1.f, s));
    LMVN_TRY(strided(work, khat, 0, fast::SM_FWD_MUL_INV, 1.f, s));
    return strided(work, nullptr, 1, fast::SM_INV, 1.f, s);
  }
  int chain_link(cplx* work, const gen::Epilogue& ep, cudaStream_t s) override { return rows_inv_fwd(work, ep, s); }
  int chain_end(cplx* work, const gen::Epilogue& ep, float* out, cudaStream_t s) override {
    return rows_inv(work, out, ep, s);
  }
  bool can_chain_embedded() const override { return chain_ok && M <= 256; }
  int chain_link_embedded(cplx* in, cplx* out, const gen::Epilogue& ep, float* psi_out, const int* logical,
                          const int* offset, cudaStream_t s) override {
    return rows_inv_fwd(in, ep, s, -1, out, logical, offset, psi_out);
  }

  int kernel_spectrum(const float* d_kernel, const int kd[3], cplx* khat, cplx* work, cudaStream_t s) override {
    gen::RealSource src{d_kernel, 1, kd[0], kd[1], kd[2]};
    cplx* dst = khat_half ? work : khat;
    if (!dst) { set_last_error("kernel spectrum: work buffer missing"); return -1; }
    LMVN_TRY(rows_fwd(src, dst, s));
    LMVN_TRY(strided(dst, nullptr, 1, fast::SM_FWD, 1.f, s));
    const float inv_n = float(1.0 / double(plan->voxels()));  // decision q10: 1/N folded into K^
    LMVN_TRY(strided(dst, nullptr, 0, fast::SM_FWD_SCALE, inv_n, s));
#ifndef LMVN_EMU
    if (khat_half) {
      // float2 -> half2 scaled by 1 / max|component| (K^ carries the 1/N factor: far below the half range unscaled)
      unsigned* max_bits = reinterpret_cast<unsigned*>(khat) + khat_elems() + 1;  // scratch word behind the un-scale factor
      float* unscale = reinterpret_cast<float*>(reinterpret_cast<unsigned*>(khat) + khat_elems());
      LMVN_CUDA_TRY(cudaMemsetAsync(max_bits, 0, sizeof(unsigned), s));
      LMVN_LAUNCH(fast::k_khat_absmax, dim3(unsigned(num_sms * 8)), dim3(256), 0, s, work, khat_elems(), max_bits);
      LMVN_LAUNCH(fast::k_khat_to_half, dim3(unsigned(num_sms * 8)), dim3(256), 0, s, work, khat_elems(), max_bits,
                  reinterpret_cast<__half2*>(khat), unscale);
      LMVN_CUDA_TRY(cudaGetLastError());
    }
#endif
    return 0;
  }

  int convolve(const float* in, cplx* work, const cplx* khat, const gen::Epilogue& ep, float* out,
               cudaStream_t s) override {
    gen::RealSource src{in, 0, 0, 0, 0};
    LMVN_TRY(rows_fwd(src, work, s));
    LMVN_TRY(strided(work, nullptr, 1, fast::SM_FWD, 1.f, s));
    LMVN_TRY(strided(work, khat, 0, fast::SM_FWD_MUL_INV, 1.f, s));
    LMVN_TRY(strided(work, nullptr, 1, fast::SM_INV, 1.f, s));
    LMVN_TRY(rows_inv(work, out, ep, s));
    return 0;
  }
};

// ------------------------------------------------------------------------------------------------------------
// Two-pass schedule (fft_x3.cuh): plane-tile pass + z-middle pass, spectrum in the A layout.  The PSF spectra are
// produced by the five-pass kernels of the base class into the work buffer and permuted into the A layout.
// ------------------------------------------------------------------------------------------------------------
struct X3Engine : FastEngine {
  int r2 = 4;
  size_t a_elems() const { return size_t(r2) * plan->nz * x3::kTileElems; }
  size_t khat_elems() const override { return std::max(a_elems(), FastEngine::khat_elems()); }
  size_t work_elems() const override { return khat_elems(); }
  int launches_per_conv() const override { return 2; }
  bool can_chain() const override { return true; }
  bool can_chain_embedded() const override { return false; }
  bool can_chain_rows() const override { return false; }

  // operand rows of the chained plane pass by bulk async copies (TMA engine + mbarrier): OPT-IN, LMVN_X3_TMA=1.  Measured
  // (profiles/r02_x3_v5_tma_probe.json): quotient link 0.285 -> 0.354 ms, update link 0.332 -> 0.410 ms -- the 64 KB staging
  // buffer next to the 136 KB tile leaves ~20 KB of the SM's 228 KB for L1 (twiddle tables, operand lines of the second
  // row iteration), and the copy adds a shared-memory write + read per operand byte to kernels that are bound by that pipe.
  int x3_stage_ops = 0;
  int x3_prefetch = 0;  // CTAs of look-ahead of the plane pass' L2 prefetch (A/B knob LMVN_X3_PREFETCH)
  int init_x3() {
    khat_half = false;
    r2 = plan->ny / x3::kRows;
    if (const char* e = getenv("LMVN_X3_TMA")) x3_stage_ops = (*e != '0');
    if (const char* e = getenv("LMVN_X3_ZMID")) zmid_version = atoi(e);
    x3_prefetch = 0;  // measured: -3 % (quotient) / -11 % (update) with a look-ahead of one CTA per SM (profiles/r02_x3_v2_probe*.json)
    if (const char* e = getenv("LMVN_X3_PREFETCH")) x3_prefetch = std::max(0, atoi(e));
    if (!tables->tw_p128[0] || !tables->tw_ny_lin) { set_last_error("two-pass schedule: tables missing"); return -1; }
    return 0;
  }

  template <int MODE, int EPI>
  int launch_plane(const x3::PlaneArgs& a, cudaStream_t s) {
    auto k = x3::k_plane<MODE, EPI>;
    const size_t smem = (MODE == x3::PM_CHAIN && a.stage_ops) ? x3::kPlaneSmemStaged : x3::kPlaneSmem;
    LMVN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(x3::kPlaneSmemStaged)));
    LMVN_LAUNCH(k, dim3(unsigned(plan->nz * r2)), dim3(x3::kPlaneThreads), smem, s, a);
    LMVN_CUDA_TRY(cudaGetLastError());
    return 0;
  }
  int plane(int mode, cplx* a_buf, const float* in, const gen::Epilogue* ep, float* out, cudaStream_t s) {
    x3::PlaneArgs a;
    std::memset(&a, 0, sizeof(a));
    a.a = a_buf;
    a.nz = plan->nz; a.ny = plan->ny; a.r2 = r2;
    a.tw_y = tables->tw_p128[0];
    a.prefetch = x3_prefetch;
    a.stage_ops = x3_stage_ops;
    a.row.tw_m = d_tw_m; a.row.tw_nx = d_tw_nx;
    a.row.nz = plan->nz; a.row.ny = plan->ny;
    int epi = gen::EPI_STORE;
    if (ep) {
      if (ep->scale != 1.f) { set_last_error("two-pass schedule: the 1/N scale lives in the PSF spectra"); return -1; }
      a.row.ep = *ep;
      epi = ep->mode;
    }
    a.row.out = out;
    if (in) a.row.src = gen::RealSource{in, 0, 0, 0, 0};
    const unsigned long long Sb = S(), Cb = C();
    if (mode == x3::PM_BEGIN) {
      LMVN_TRY((launch_plane<x3::PM_BEGIN, gen::EPI_STORE>(a, s)));
      mark("x3_plane_fwd", Sb + Cb, s);
    } else if (mode == x3::PM_CHAIN) {
      if (epi == gen::EPI_QUOTIENT) {
        LMVN_TRY((launch_plane<x3::PM_CHAIN, gen::EPI_QUOTIENT>(a, s)));
        mark("x3_plane_quotient", 2 * Cb + Sb, s);
      } else if (epi == gen::EPI_UPDATE) {
        LMVN_TRY((launch_plane<x3::PM_CHAIN, gen::EPI_UPDATE>(a, s)));
        mark("x3_plane_update", 2 * Cb + 3 * Sb, s);
      } else { set_last_error("two-pass schedule: chained pass needs a quotient or update epilogue"); return -1; }
    } else {
      switch (epi) {
        case gen::EPI_QUOTIENT: LMVN_TRY((launch_plane<x3::PM_END, gen::EPI_QUOTIENT>(a, s))); break;
        case gen::EPI_UPDATE: LMVN_TRY((launch_plane<x3::PM_END, gen::EPI_UPDATE>(a, s))); break;
        default: LMVN_TRY((launch_plane<x3::PM_END, gen::EPI_STORE>(a, s))); break;
      }
      mark("x3_plane_inv", Cb + Sb * (epi == gen::EPI_UPDATE ? 3 : (epi == gen::EPI_QUOTIENT ? 2 : 1)), s);
    }
    return 0;
  }

  template <int NZ, int RR>
  int launch_zmid(const x3::ZmidArgs& a, cudaStream_t s) {
    constexpr int COLS = x3::ZmidCols<NZ>::V;
    const size_t smem = size_t(NZ) * x3::ZmidPitch<NZ, RR>::V * sizeof(cplx);
    auto k = x3::k_zmid<NZ, RR>;
    LMVN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const unsigned tiles = unsigned((x3::kRows + 1) * (x3::kM / (COLS / RR)));
    LMVN_LAUNCH(k, dim3(tiles), dim3(fast::Threads<NZ>::V), smem, s, a);
    LMVN_CUDA_TRY(cudaGetLastError());
    return 0;
  }
  template <int NZ, int RR>
  int launch_zmid5(const x3::ZmidArgs& a, cudaStream_t s) {
    const size_t smem = size_t(NZ) * (16 + 16 / RR) * sizeof(cplx);
    auto k = x3::k_zmid5<NZ, RR>;
    LMVN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const unsigned tiles = unsigned((x3::kRows + 1) * (x3::kM / (16 / RR)));
    LMVN_LAUNCH(k, dim3(tiles), dim3(256), smem, s, a);
    LMVN_CUDA_TRY(cudaGetLastError());
    return 0;
  }
  int zmid_version = 5;  // LMVN_X3_ZMID=4: the four-barrier form (k_zmid) also for 16-column tiles (A/B)
  int zmid(cplx* a_buf, const cplx* khat, cudaStream_t s) {
    x3::ZmidArgs a;
    std::memset(&a, 0, sizeof(a));
    a.a = a_buf; a.khat = khat; a.nz = plan->nz; a.r2 = r2;
    a.tw1 = d_tw_z[0]; a.tw2 = d_tw_z[1]; a.tw_ny = tables->tw_ny_lin;
    int rc = -1;
    if (zmid_version == 5 && (plan->nz == 256 || plan->nz == 512)) {
      if (r2 == 4) rc = (plan->nz == 512) ? launch_zmid5<512, 4>(a, s) : launch_zmid5<256, 4>(a, s);
      else rc = (plan->nz == 512) ? launch_zmid5<512, 2>(a, s) : launch_zmid5<256, 2>(a, s);
    } else if (r2 == 4) {
      switch (plan->nz) {
        case 64: rc = launch_zmid<64, 4>(a, s); break;
        case 128: rc = launch_zmid<128, 4>(a, s); break;
        case 256: rc = launch_zmid<256, 4>(a, s); break;
        case 512: rc = launch_zmid<512, 4>(a, s); break;
      }
    } else if (r2 == 2) {
      switch (plan->nz) {
        case 64: rc = launch_zmid<64, 2>(a, s); break;
        case 128: rc = launch_zmid<128, 2>(a, s); break;
        case 256: rc = launch_zmid<256, 2>(a, s); break;
        case 512: rc = launch_zmid<512, 2>(a, s); break;
      }
    }
    if (rc != 0) { if (rc == -1 && !*last_error()) set_last_error("two-pass schedule: unsupported z extent"); return -1; }
    mark("x3_z_mul", 3 * C(), s);
    return 0;
  }

  int chain_begin(const float* in, cplx* work, cudaStream_t s) override { return plane(x3::PM_BEGIN, work, in, nullptr, nullptr, s); }
  int chain_middle(cplx* work, const cplx* khat, cudaStream_t s) override { return zmid(work, khat, s); }
  int chain_link(cplx* work, const gen::Epilogue& ep, cudaStream_t s) override { return plane(x3::PM_CHAIN, work, nullptr, &ep, nullptr, s); }
  int chain_end(cplx* work, const gen::Epilogue& ep, float* out, cudaStream_t s) override {
    return plane(x3::PM_END, work, nullptr, &ep, out, s);
  }
  int convolve(const float* in, cplx* work, const cplx* khat, const gen::Epilogue& ep, float* out, cudaStream_t s) override {
    LMVN_TRY(plane(x3::PM_BEGIN, work, in, nullptr, nullptr, s));
    LMVN_TRY(zmid(work, khat, s));
    return plane(x3::PM_END, work, nullptr, &ep, out, s);
  }
  int kernel_spectrum(const float* d_kernel, const int kd[3], cplx* khat, cplx* work, cudaStream_t s) override {
    // five-pass kernels into the work buffer (split layout of the base class), then the permutation into the A layout
    LMVN_TRY(FastEngine::kernel_spectrum(d_kernel, kd, work, nullptr, s));
    x3::KhatPermArgs k;
    k.main = work; k.nyq = nyq_of(work); k.out = khat;
    k.nz = plan->nz; k.ny = plan->ny; k.r2 = r2;
    switch (plan->ny) {
      case 512: k.y_r1 = fast::Radix<512>::R1; k.y_r2 = fast::Radix<512>::R2; k.y_r3 = 1; break;
      default: k.y_r1 = fast::Radix<256>::R1; k.y_r2 = fast::Radix<256>::R2; k.y_r3 = 1; break;
    }
    LMVN_LAUNCH(x3::k_khat_to_a, dim3(unsigned(num_sms * 8)), dim3(256), 0, s, k);
    LMVN_CUDA_TRY(cudaGetLastError());
    return 0;
  }
};

bool axis_ok(int n) { return n == 16 || n == 32 || n == 64 || n == 128 || n == 256 || n == 512 || n == 1024; }
bool nx_ok(int nx) { return nx == 64 || nx == 128 || nx == 256 || nx == 512 || nx == 1024; }

}  // namespace

std::unique_ptr<FastOps> make_fast_ops(std::shared_ptr<FftPlan> plan) {
  const int nx = plan->nx;
  if (!nx_ok(nx) || !axis_ok(plan->ny) || !axis_ok(plan->nz)) {
    set_last_error("slab-decomposed engine: dims %dx%dx%d are not supported by the power-of-two fast path",
                   plan->nz, plan->ny, plan->nx);
    return nullptr;
  }
  std::unique_ptr<FastEngine> e(new FastEngine());
  e->plan = plan;
  e->split = false;  // whole spectrum rows travel between the devices
  if (cudaSetDevice(plan->device) != cudaSuccess) return nullptr;
  if (e->init() != 0) return nullptr;
  return std::unique_ptr<FastOps>(e.release());
}

bool fused_shape_ok(int nz, int ny, int nx) {
  return nx_ok(nx) && axis_ok(ny) && axis_ok(nz) && (size_t(nz) * ny) % 128 == 0;
}

std::unique_ptr<ConvEngine> make_fused_engine(std::shared_ptr<FftPlan> plan) {
  const int nx = plan->nx;
  if (!nx_ok(nx)) return nullptr;
  if (!axis_ok(plan->ny) || !axis_ok(plan->nz)) return nullptr;
  const size_t rows = size_t(plan->nz) * plan->ny;
  if (rows % 128 != 0) return nullptr;
  if (cudaSetDevice(plan->device) != cudaSuccess) return nullptr;
  // The two-pass schedule (fft_x3.cuh) is OPT-IN (LMVN_X3=1): it moves 4S + 10C instead of 4S + 18C per (view, iteration)
  // and is still 12 % slower than the chained five-pass loop on BASELINE config 3 (1.17 vs 1.04 ms, DESIGN.md 3.4,
  // profiles/r02_x3_*.json): the passes are bound by the SM's load/store/shared-memory pipe, which the fusion does
  // not relieve (it trades global for shared wavefronts one for one).
  bool x3 = false;
  if (const char* e = getenv("LMVN_X3")) x3 = (*e != '0') && x3_shape_ok(plan->nz, plan->ny, plan->nx);
  if (const char* e = getenv("LMVN_CHAIN")) x3 = x3 && (*e != '0');  // LMVN_CHAIN=0: the unchained five-pass loop
  if (x3) {
    std::unique_ptr<X3Engine> e(new X3Engine());
    e->plan = plan;
    if (e->init() != 0 || e->init_x3() != 0) return nullptr;
    return std::unique_ptr<ConvEngine>(e.release());
  }
  std::unique_ptr<FastEngine> e(new FastEngine());
  e->plan = plan;
  if (e->init() != 0) return nullptr;
#ifndef LMVN_EMU
  if (const char* h = getenv("LMVN_KHAT_FP16")) e->khat_half = (*h == '1');
#endif
  return std::unique_ptr<ConvEngine>(e.release());
}

}  // namespace lmvn
