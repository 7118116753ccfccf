"""CPU restatement of libmultiviewnative's multi-view deconvolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (``libmultiviewnative_b200/``)
may import this module; it is the checker for ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.

Parity status: **the deconvolution loop is parity-unpinned by data** -- the
reference's golden TIFF stacks (``psi_i.tif`` ...) are not shipped with its
repository (expected under /dev/shm/libmultiview_data, CMakeLists.txt:25) and
the reference's own CPU path cannot be compiled here (needs FFTW 3 float and
Boost 1.55, neither installed; ``inc/fftw_interface.h:19``,
``inc/cpu_convolve.h:11``).  What *is* pinned: every data-independent
known-answer case the reference's tests hold (8^3 convolution fixture, impulse
response, FFT round trip on integer ramps, the pointwise-kernel constants) --
see ``tests/test_oracle_golden.py`` and ``tests/golden/make_golden.py``.

The arithmetic of the FFT itself lives in FFTW 3 ("3.1 or later", README.md:20,
not vendored, no pinned version).  It is replaced here by pocketfft
(``scipy.fft``) in float32 -- the published algorithm is the same unnormalised
DFT with sign -1 forward, so results agree to float32 round-off (measured
5.8e-7 max relative between pocketfft and MKL on config 1).

All ``file:line`` citations are relative to the reference checkout.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft

F32 = np.float32


# --------------------------------------------------------------------------- #
# kernel placement: inc/padd_utils.h:11-40 (wrapped_insert_at_point) as used by
# no_padd::wrapped_insert_at_offsets (inc/padd_utils.h:91-95)
# --------------------------------------------------------------------------- #
def wrap_kernel(kernel: np.ndarray, dims) -> np.ndarray:
    """Place ``kernel`` into a zero volume of shape ``dims`` with its centre
    element (index k//2 per axis) at the origin, negative offsets wrapping to
    the far end.  Element (z,y,x) -> ((z-kz//2) mod nz, ...).
    """
    dims = tuple(int(d) for d in dims)
    kernel = np.asarray(kernel, dtype=F32)
    if any(k > n for k, n in zip(kernel.shape, dims)):
        # undefined in the reference (writes out of bounds); decision q11
        raise ValueError("kernel larger than image")
    out = np.zeros(dims, dtype=F32)
    idx = []
    for ax in range(3):
        k = kernel.shape[ax]
        i = np.arange(k) - k // 2
        i = np.where(i < 0, i + dims[ax], i)
        idx.append(i)
    out[np.ix_(*idx)] = kernel
    return out


# --------------------------------------------------------------------------- #
# transform: inc/fft_utils.h:55-105 + inc/plan_store.h:99-124
# (fftwf_plan_dft_r2c_3d / c2r_3d, in place, unnormalised)
# --------------------------------------------------------------------------- #
def fft_forward(stack: np.ndarray, workers: int = 1) -> np.ndarray:
    """r2c 3-D transform, float32 in -> complex64 half spectrum (nz,ny,nx//2+1)."""
    return sfft.rfftn(np.asarray(stack, dtype=F32), workers=workers)


def fft_backward(spec: np.ndarray, dims, workers: int = 1) -> np.ndarray:
    """c2r 3-D transform, UNNORMALISED like FFTW's (norm='forward' leaves the
    inverse unscaled)."""
    return sfft.irfftn(spec, s=tuple(dims), norm="forward", workers=workers).astype(F32, copy=False)


def forwarded_kernel(kernel: np.ndarray, dims, workers: int = 1) -> np.ndarray:
    """src/multiviewnative.cpp:146-174: wrap, pad for in-place r2c, transform."""
    return fft_forward(wrap_kernel(kernel, dims), workers)


# --------------------------------------------------------------------------- #
# convolution: inc/cpu_convolve.h:217-291 (half_inplace) and :147-202 (inplace)
# --------------------------------------------------------------------------- #
def half_inplace(image: np.ndarray, khat: np.ndarray, workers: int = 1) -> np.ndarray:
    """image <- c2r(r2c(image) * khat) * float(1/N); circular at the image
    extents (PaddingT = no_padd, inc/cpu_convolve.h:24)."""
    dims = image.shape
    if khat.shape != (dims[0], dims[1], dims[2] // 2 + 1):
        # inc/cpu_convolve.h:226-250 throws std::length_error
        raise ValueError("kernel buffer received is ill shaped for convolution")
    spec = fft_forward(image, workers)
    # inc/cpu_convolve.h:257-266: 4 mul + 2 add in float
    spec = (spec * khat).astype(np.complex64, copy=False)
    out = fft_backward(spec, dims, workers)
    n = int(np.prod(dims))
    scale = F32(1.0 / n)  # inc/cpu_convolve.h:274
    out *= scale
    return out


def inplace_cpu_convolution(image: np.ndarray, kernel: np.ndarray, nthreads: int = 1) -> np.ndarray:
    """src/multiviewnative.cpp:273-293 -> cpu_convolve::inplace.  Returns the
    convolved image (the reference overwrites ``im``)."""
    image = np.asarray(image, dtype=F32)
    w = _workers(nthreads)
    return half_inplace(image, forwarded_kernel(kernel, image.shape, w), w)


# --------------------------------------------------------------------------- #
# pointwise steps: inc/cpu_kernels.h
# --------------------------------------------------------------------------- #
def compute_quotient(inp: np.ndarray, out: np.ndarray) -> np.ndarray:
    """inc/cpu_kernels.h:19-26: temp = float(1. / out); out = in * temp."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        temp = (1.0 / out.astype(np.float64)).astype(F32)
        return (inp * temp).astype(F32, copy=False)


def final_values(psi, integral, weight, min_value) -> np.ndarray:
    """inc/cpu_kernels.h:28-54 (serial semantics; the parallel dispatcher has the
    size/nthreads swap bug, :299-309, decision q2)."""
    min_value = F32(min_value)
    with np.errstate(invalid="ignore", over="ignore"):
        last = psi.astype(F32, copy=False)
        value = (last * integral).astype(F32)
        value = np.where(value > 0, value, min_value).astype(F32)
        bad = ~np.isfinite(value)
        nxt = np.where(bad, min_value, np.maximum(value, min_value)).astype(F32)
        nxt = (weight * (nxt - last)).astype(F32) + last
    return nxt.astype(F32, copy=False)


def regularized_final_values(psi, integral, weight, lam, min_value) -> np.ndarray:
    """inc/cpu_kernels.h:59-90: Tikhonov step evaluated in double,
    lambda_inv = float(1.f / lambda)."""
    min_value = F32(min_value)
    lam = float(lam)
    lambda_inv = F32(1.0 / lam)  # TransferT lambda_inv = 1.f / _lambda  (:71)
    with np.errstate(invalid="ignore", over="ignore"):
        last = psi.astype(F32, copy=False)
        value = (last * integral).astype(F32)
        pos = value > 0
        v64 = np.where(pos, value, 0).astype(np.float64)
        reg = (np.float64(lambda_inv) * (np.sqrt(1.0 + 2.0 * lam * v64) - 1.0)).astype(F32)
        value = np.where(pos, reg, min_value).astype(F32)
        bad = ~np.isfinite(value)
        nxt = np.where(bad, min_value, np.maximum(value, min_value)).astype(F32)
        nxt = (weight * (nxt - last)).astype(F32) + last
    return nxt.astype(F32, copy=False)


# --------------------------------------------------------------------------- #
# the loop: src/multiviewnative.cpp:101-240
# --------------------------------------------------------------------------- #
def _workers(nthreads: int) -> int:
    import os

    if nthreads is None or nthreads <= 0:  # tests pass -1 for "all"
        return os.cpu_count() or 1
    return int(nthreads)


def inplace_cpu_deconvolve(psi, views, kernels1, kernels2, weights, num_iterations,
                           lam=0.0, min_value=1e-4, nthreads=1, khats=None, zero_view_guard=False):
    """Multi-view Richardson-Lucy (Tikhonov if lam > 0).

    zero_view_guard (NOT in the reference; the rule of the new build's zero_padd mode): the
    quotient of a voxel whose view is exactly zero is zero, whatever the blurred estimate is.  Explicitly
    zero-padded stacks otherwise hit 0 * (1 / 0) = NaN in padding the kernels barely reach, and the second
    convolution spreads it over the stack (the reference's GPU path does, ref: inc/cuda_kernels.cuh:14-31).

    psi: (nz,ny,nx) float32 start value; views/weights: sequences of the same
    shape; kernels1/kernels2: sequences of small 3-D PSFs.  Returns the new psi.
    One iteration = one sweep over all views, psi updated after each view
    (src/multiviewnative.cpp:191-229).
    """
    w = _workers(nthreads)
    psi = np.array(psi, dtype=F32, copy=True)
    dims = psi.shape
    nviews = len(views)
    for v in range(nviews):
        if tuple(views[v].shape) != tuple(dims):
            raise ValueError("all views must share psi's dims (decision q6)")
    if khats is None:
        k1 = [forwarded_kernel(kernels1[v], dims, w) for v in range(nviews)]
        k2 = [forwarded_kernel(kernels2[v], dims, w) for v in range(nviews)]
    else:
        k1, k2 = khats
    for _ in range(int(num_iterations)):
        for v in range(nviews):
            integral = half_inplace(psi, k1[v], w)  # :195-201
            integral = compute_quotient(np.asarray(views[v], dtype=F32), integral)  # :204
            if zero_view_guard:
                integral = np.where(np.asarray(views[v]) == 0, F32(0), integral).astype(F32, copy=False)
            integral = half_inplace(integral, k2[v], w)  # :209-211
            if lam > 0:  # :216
                psi = regularized_final_values(psi, integral, np.asarray(weights[v], dtype=F32), lam, min_value)
            else:
                psi = final_values(psi, integral, np.asarray(weights[v], dtype=F32), min_value)
    return psi


# --------------------------------------------------------------------------- #
# float64 referee (accuracy arbiter, not a parity target)
# --------------------------------------------------------------------------- #
def deconvolve_f64(psi, views, kernels1, kernels2, weights, num_iterations, lam=0.0,
                   min_value=1e-4, workers=-1):
    w = _workers(workers)
    psi = np.array(psi, dtype=np.float64)
    dims = psi.shape

    def kh(k):
        return sfft.rfftn(wrap_kernel(k, dims).astype(np.float64), workers=w)

    k1 = [kh(k) for k in kernels1]
    k2 = [kh(k) for k in kernels2]

    def conv(a, k):
        return sfft.irfftn(sfft.rfftn(a, workers=w) * k, s=dims, workers=w)

    for _ in range(int(num_iterations)):
        for v in range(len(views)):
            t = conv(psi, k1[v])
            with np.errstate(divide="ignore", invalid="ignore"):
                t = np.asarray(views[v], dtype=np.float64) / t
            t = conv(t, k2[v])
            val = psi * t
            pos = val > 0
            if lam > 0:
                reg = (np.sqrt(1.0 + 2.0 * lam * np.where(pos, val, 0)) - 1.0) / lam
                val = np.where(pos, reg, min_value)
            else:
                val = np.where(pos, val, min_value)
            val = np.where(np.isfinite(val), np.maximum(val, min_value), min_value)
            psi = np.asarray(weights[v], dtype=np.float64) * (val - psi) + psi
    return psi


# --------------------------------------------------------------------------- #
# direct (spatial) convolution used by the reference's fixtures as ground truth:
# tests/test_algorithms.hpp:9-58
# --------------------------------------------------------------------------- #
def direct_convolve(image: np.ndarray, kernel: np.ndarray, offset) -> np.ndarray:
    """result[i] = sum_k kernel[K-1-k] * image[i - K//2 + k] for i inside
    [offset, shape-offset); elements outside stay equal to ``image`` (the fixture
    starts from a copy, tests/test_fixtures.hpp:228-231)."""
    image = np.asarray(image, dtype=F32)
    kernel = np.asarray(kernel, dtype=F32)
    res = image.copy()
    kz, ky, kx = kernel.shape
    hz, hy, hx = kz // 2, ky // 2, kx // 2
    oz, oy, ox = offset
    kflip = kernel[::-1, ::-1, ::-1]
    for z in range(oz, image.shape[0] - oz):
        for y in range(oy, image.shape[1] - oy):
            for x in range(ox, image.shape[2] - ox):
                acc = F32(0)
                patch = image[z - hz:z - hz + kz, y - hy:y - hy + ky, x - hx:x - hx + kx]
                # float accumulation in kernel z,y,x order like the reference loop
                for v in (kflip * patch).ravel():
                    acc = F32(acc + v)
                res[z, y, x] = acc
    return res


# --------------------------------------------------------------------------- #
# torch (MKL) twin: same restatement on a second, faster FFT back-end.  Used for
# cross-checking the oracle and as the multi-threaded timed CPU baseline.
# --------------------------------------------------------------------------- #
def torch_forwarded_kernels(kernels, dims, nthreads=-1):
    """K^ of every kernel on the torch back-end (src/multiviewnative.cpp:146-174), for `khats=` below."""
    import torch

    torch.set_num_threads(_workers(nthreads))
    return [torch.fft.rfftn(torch.from_numpy(wrap_kernel(k, tuple(dims)))) for k in kernels]


def inplace_cpu_deconvolve_torch(psi, views, kernels1, kernels2, weights, num_iterations,
                                 lam=0.0, min_value=1e-4, nthreads=-1, checkpoints=None, max_units=None, khats=None):
    """khats: optional (K^1 list, K^2 list) from torch_forwarded_kernels (kernels1/2 are then ignored).
    checkpoints: optional dict filled with {iteration count: psi copy} for the counts it holds as keys (the
    full-size parity test reads 1 and 10 iterations from ONE run).  max_units: stop after that many (view,
    iteration) units (bench.py times a bounded sample of the loop through this very function)."""
    import torch

    n = _workers(nthreads)
    torch.set_num_threads(n)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F32))
    psi_t = t(psi).clone()
    dims = tuple(psi_t.shape)
    nvox = int(np.prod(dims))
    scale = float(F32(1.0 / nvox))
    if khats is None:
        k1 = [torch.fft.rfftn(t(wrap_kernel(k, dims))) for k in kernels1]
        k2 = [torch.fft.rfftn(t(wrap_kernel(k, dims))) for k in kernels2]
    else:
        k1, k2 = khats
    views_t = [t(v) for v in views]
    weights_t = [t(wt) for wt in weights]
    mv = float(F32(min_value))
    lam_inv = float(F32(1.0 / lam)) if lam > 0 else 0.0

    def conv(a, k):
        s = torch.fft.rfftn(a)
        s = s * k
        return torch.fft.irfftn(s, s=dims, norm="forward") * scale

    units = 0
    for it in range(int(num_iterations)):
        for v in range(len(views_t)):
            if max_units is not None and units >= max_units:
                return psi_t.numpy()
            units += 1
            integ = conv(psi_t, k1[v])
            integ = views_t[v] * (1.0 / integ.double()).float()
            integ = conv(integ, k2[v])
            val = psi_t * integ
            pos = val > 0
            if lam > 0:
                reg = (lam_inv * (torch.sqrt(1.0 + 2.0 * lam * val.double().clamp_min(0)) - 1.0)).float()
                val = torch.where(pos, reg, torch.full_like(val, mv))
            else:
                val = torch.where(pos, val, torch.full_like(val, mv))
            val = torch.where(torch.isfinite(val), val.clamp_min(mv), torch.full_like(val, mv))
            psi_t = weights_t[v] * (val - psi_t) + psi_t
        if checkpoints is not None and (it + 1) in checkpoints:
            checkpoints[it + 1] = psi_t.numpy().copy()
    return psi_t.numpy()
