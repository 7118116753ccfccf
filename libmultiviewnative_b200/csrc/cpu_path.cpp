// cpu_path.cpp -- inplace_cpu_convolution / inplace_cpu_deconvolve.
//
// The reference exports these two CPU entry points from the same shared object
// (ref: src/multiviewnative.cpp:244-293) and its tests and benches link them, so a
// drop-in must have them.  They are NOT the accelerated path and the GPU entry
// points never fall back to them.  FFTW is replaced by a small host mixed-radix
// Stockham transform (OpenMP over lines); semantics follow the reference's serial
// kernels (ref: inc/cpu_kernels.h:19-90, inc/cpu_convolve.h:217-291).
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "multiviewnative.h"

namespace {

typedef std::complex<float> cf;

struct Axis {
  int n = 0;
  std::vector<int> factors;
  std::vector<cf> tw;
  explicit Axis(int n_) : n(n_), tw(n_) {
    int m = n;
    while (m % 4 == 0) { factors.push_back(4); m /= 4; }
    while (m % 2 == 0) { factors.push_back(2); m /= 2; }
    for (int p = 3; p * p <= m; p += 2)
      while (m % p == 0) { factors.push_back(p); m /= p; }
    if (m > 1) factors.push_back(m);
    for (int k = 0; k < n; ++k) {
      const double a = -2.0 * M_PI * double(k) / double(n);
      tw[k] = cf(float(std::cos(a)), float(std::sin(a)));
    }
  }
  // transforms a (length n) using b as scratch; returns the buffer holding the result
  cf* run(cf* a, cf* b, bool inverse) const {
    int Ns = 1;
    for (int R : factors) {
      const int m = n / R, tstep = n / (Ns * R);
      for (int j = 0; j < m; ++j) {
        const int k = j % Ns, j0 = (j / Ns) * Ns * R + k;
        for (int q = 0; q < R; ++q) {
          const int e = (k + q * Ns) * tstep;
          cf acc(0.f, 0.f);
          int idx = 0;
          for (int r = 0; r < R; ++r) {
            cf t = tw[idx];
            if (inverse) t = std::conj(t);
            acc += a[j + r * m] * t;
            idx += e;
            if (idx >= n) idx -= n;
          }
          b[j0 + q * Ns] = acc;
        }
      }
      std::swap(a, b);
      Ns *= R;
    }
    return a;
  }
};

struct Fft3 {
  int nz, ny, nx, nxc;
  Axis az, ay, ax;
  Fft3(int z, int y, int x) : nz(z), ny(y), nx(x), nxc(x / 2 + 1), az(z), ay(y), ax(x) {}
  size_t spec_elems() const { return size_t(nz) * ny * nxc; }

  void strided(std::vector<cf>& s, const Axis& A, size_t outer, size_t ostride, size_t jstride, size_t inner,
               bool inverse, int nthreads) const {
    const long long lines = (long long)(outer * inner);
#pragma omp parallel num_threads(nthreads)
    {
      std::vector<cf> a(A.n), b(A.n);
#pragma omp for schedule(static)
      for (long long t = 0; t < lines; ++t) {
        const size_t o = size_t(t) / inner, i = size_t(t) % inner;
        cf* base = s.data() + o * ostride + i;
        for (int j = 0; j < A.n; ++j) a[j] = base[j * jstride];
        cf* r = A.run(a.data(), b.data(), inverse);
        for (int j = 0; j < A.n; ++j) base[j * jstride] = r[j];
      }
    }
  }
  void forward(const float* in, std::vector<cf>& s, int nthreads) const {
    s.resize(spec_elems());
    const long long rows = (long long)nz * ny;
#pragma omp parallel num_threads(nthreads)
    {
      std::vector<cf> a(nx), b(nx);
#pragma omp for schedule(static)
      for (long long r = 0; r < rows; ++r) {
        for (int x = 0; x < nx; ++x) a[x] = cf(in[r * nx + x], 0.f);
        cf* o = ax.run(a.data(), b.data(), false);
        std::memcpy(&s[r * nxc], o, sizeof(cf) * nxc);
      }
    }
    strided(s, ay, nz, size_t(ny) * nxc, nxc, nxc, false, nthreads);
    strided(s, az, 1, 0, size_t(ny) * nxc, size_t(ny) * nxc, false, nthreads);
  }
  // unnormalised c2r; destroys s
  void backward(std::vector<cf>& s, float* out, int nthreads) const {
    strided(s, az, 1, 0, size_t(ny) * nxc, size_t(ny) * nxc, true, nthreads);
    strided(s, ay, nz, size_t(ny) * nxc, nxc, nxc, true, nthreads);
    const long long rows = (long long)nz * ny;
#pragma omp parallel num_threads(nthreads)
    {
      std::vector<cf> a(nx), b(nx);
#pragma omp for schedule(static)
      for (long long r = 0; r < rows; ++r) {
        for (int x = 0; x < nx; ++x) a[x] = (x < nxc) ? s[r * nxc + x] : std::conj(s[r * nxc + (nx - x)]);
        cf* o = ax.run(a.data(), b.data(), true);
        for (int x = 0; x < nx; ++x) out[r * nx + x] = o[x].real();
      }
    }
  }
};

bool wrap_kernel(const float* k, const int* kd, const int* d, std::vector<float>& out) {
  for (int a = 0; a < 3; ++a)
    if (kd[a] <= 0 || kd[a] > d[a]) return false;
  out.assign(size_t(d[0]) * d[1] * d[2], 0.f);
  for (int z = 0; z < kd[0]; ++z)
    for (int y = 0; y < kd[1]; ++y)
      for (int x = 0; x < kd[2]; ++x) {
        int tz = z - kd[0] / 2, ty = y - kd[1] / 2, tx = x - kd[2] / 2;  // ref: inc/padd_utils.h:25-36
        if (tz < 0) tz += d[0];
        if (ty < 0) ty += d[1];
        if (tx < 0) tx += d[2];
        out[(size_t(tz) * d[1] + ty) * d[2] + tx] = k[(size_t(z) * kd[1] + y) * kd[2] + x];
      }
  return true;
}

int threads_of(int nthreads) {
#ifdef _OPENMP
  return nthreads > 0 ? nthreads : omp_get_num_procs();
#else
  (void)nthreads;
  return 1;
#endif
}

void convolve(const Fft3& f, float* image, const std::vector<cf>& khat, std::vector<cf>& spec, int nt) {
  f.forward(image, spec, nt);
  const size_t n = f.spec_elems();
  // ref: inc/cpu_convolve.h:257-266 runs this loop and the scale loop below serially; element-wise, so the threaded
  // form gives the same bits
#pragma omp parallel for num_threads(nt) schedule(static)
  for (long long i = 0; i < (long long)n; ++i) spec[i] *= khat[i];
  f.backward(spec, image, nt);
  const size_t vox = size_t(f.nz) * f.ny * f.nx;
  const float scale = float(1.0 / double(vox));
#pragma omp parallel for num_threads(nt) schedule(static)
  for (long long i = 0; i < (long long)vox; ++i) image[i] *= scale;
}

}  // namespace

// the error contract of include/multiviewnative.h: message to stderr + lmvn_last_error(), outputs untouched
namespace lmvn { void set_last_error(const char* fmt, ...); }  // csrc/engine.cu
static void fail(const char* what) {
  lmvn::set_last_error("%s", what);
  std::fprintf(stderr, "[libmultiviewnative] error: %s\n", what);
}

extern "C" void inplace_cpu_convolution(imageType* im, int* imDim, imageType* kernel, int* kernelDim, int nthreads) {
  if (!im || !imDim || !kernel || !kernelDim) return fail("inplace_cpu_convolution: null argument");
  for (int a = 0; a < 3; ++a)
    if (imDim[a] <= 0 || kernelDim[a] <= 0) return fail("inplace_cpu_convolution: non-positive extent");
  std::vector<float> padded;
  if (!wrap_kernel(kernel, kernelDim, imDim, padded)) return fail("inplace_cpu_convolution: kernel does not fit the image");
  const int nt = threads_of(nthreads);
  Fft3 f(imDim[0], imDim[1], imDim[2]);
  std::vector<cf> khat, spec;
  f.forward(padded.data(), khat, nt);
  convolve(f, im, khat, spec, nt);
}

extern "C" void inplace_cpu_deconvolve(imageType* psi, workspace input, int nthreads) {
  // same validation and reporting as the GPU entry point (gpu_deconvolve_impl, csrc/api.cu)
  if (!psi || !input.data_ || input.num_views_ == 0) return fail("inplace_cpu_deconvolve: null psi / no views");
  if (!input.data_[0].image_dims_) return fail("inplace_cpu_deconvolve: view 0 has no image_dims_");
  const int* d = input.data_[0].image_dims_;
  for (int v = 0; v < input.num_views_; ++v) {
    const view_data& vd = input.data_[v];
    if (!vd.image_ || !vd.weights_ || !vd.kernel1_ || !vd.kernel2_ || !vd.image_dims_ || !vd.kernel1_dims_ || !vd.kernel2_dims_)
      return fail("inplace_cpu_deconvolve: a view has a null buffer or dims pointer");
    for (int a = 0; a < 3; ++a) {
      if (vd.image_dims_[a] != d[a] || d[a] <= 0)
        return fail("inplace_cpu_deconvolve: all views must share view 0's (positive) image dims");  // decision q6
      if (vd.kernel1_dims_[a] <= 0 || vd.kernel2_dims_[a] <= 0 || vd.kernel1_dims_[a] > d[a] || vd.kernel2_dims_[a] > d[a])
        return fail("inplace_cpu_deconvolve: a kernel does not fit the image");  // decision q11
    }
  }
  const int nt = threads_of(nthreads);
  Fft3 f(d[0], d[1], d[2]);
  const size_t vox = size_t(d[0]) * d[1] * d[2];
  std::vector<std::vector<cf>> k1(input.num_views_), k2(input.num_views_);
  std::vector<float> padded;
  for (int v = 0; v < input.num_views_; ++v) {
    const view_data& vd = input.data_[v];
    if (!wrap_kernel(vd.kernel1_, vd.kernel1_dims_, d, padded)) return fail("inplace_cpu_deconvolve: kernel1 does not fit the image");
    f.forward(padded.data(), k1[v], nt);
    if (!wrap_kernel(vd.kernel2_, vd.kernel2_dims_, d, padded)) return fail("inplace_cpu_deconvolve: kernel2 does not fit the image");
    f.forward(padded.data(), k2[v], nt);
  }
  std::vector<float> integral(vox);
  std::vector<cf> spec;
  const float minv = input.minValue_;
  const double lambda = input.lambda_;
  const float lambda_inv = float(1.0 / lambda);
  for (int it = 0; it < input.num_iterations_; ++it) {
    for (int v = 0; v < input.num_views_; ++v) {
      const view_data& vd = input.data_[v];
      std::memcpy(integral.data(), psi, vox * sizeof(float));
      convolve(f, integral.data(), k1[v], spec, nt);
#pragma omp parallel for num_threads(nt) schedule(static)
      for (long long i = 0; i < (long long)vox; ++i) {
        const float t = float(1.0 / double(integral[i]));
        integral[i] = vd.image_[i] * t;
      }
      convolve(f, integral.data(), k2[v], spec, nt);
#pragma omp parallel for num_threads(nt) schedule(static)
      for (long long i = 0; i < (long long)vox; ++i) {
        const float last = psi[i];
        float value = last * integral[i];
        if (value > 0.f) {
          if (lambda > 0.) value = float(double(lambda_inv) * (std::sqrt(1. + 2. * lambda * double(value)) - 1.));
        } else {
          value = minv;
        }
        float next = (std::isnan(value) || std::isinf(value)) ? minv : std::max(value, minv);
        psi[i] = vd.weights_[i] * (next - last) + last;
      }
    }
  }
}
