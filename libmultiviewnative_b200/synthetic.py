"""Seeded synthetic multi-view data sets (numpy only).

Two families:

* ``make_views`` -- the well-conditioned protocol of SURVEY.md §8(d): a DC
  background of 10 under point sources, anisotropic Gaussian PSFs, weights 1/V,
  psi0 = mean(view0).  The background keeps the blurred estimate away from
  zero so that float32 FFT round-off does not dominate per-voxel *relative*
  comparisons.
* ``reference_bench_views`` -- the reference's own synthetic benchmark protocol
  (bench/synthetic_data.hpp:58-96): constant views 16+4i, unit weights, delta
  kernels 21^3 (value i+1) and 25^3 (value i+2).
"""
from __future__ import annotations

import os

import numpy as np

F32 = np.float32


def gaussian_psf(size: int, sigma_zyx, dtype=F32) -> np.ndarray:
    """Normalised anisotropic Gaussian on a size^3 grid centred at size//2."""
    ax = np.arange(size, dtype=np.float64) - size // 2
    g = [np.exp(-0.5 * (ax / s) ** 2) for s in sigma_zyx]
    psf = g[0][:, None, None] * g[1][None, :, None] * g[2][None, None, :]
    psf /= psf.sum()
    return psf.astype(dtype)


_ORIENT = [
    (4.0, 1.5, 1.5), (1.5, 4.0, 1.5), (1.5, 1.5, 4.0),
    (3.0, 2.0, 1.2), (1.2, 3.0, 2.0), (2.0, 1.2, 3.0),
]


def _circ_conv64(a: np.ndarray, k_small: np.ndarray, workers: int) -> np.ndarray:
    import scipy.fft as sfft

    dims = a.shape
    pad = np.zeros(dims, dtype=np.float64)
    idx = []
    for axn in range(3):
        k = k_small.shape[axn]
        i = np.arange(k) - k // 2
        idx.append(np.where(i < 0, i + dims[axn], i))
    pad[np.ix_(*idx)] = k_small
    return sfft.irfftn(sfft.rfftn(a, workers=workers) * sfft.rfftn(pad, workers=workers), s=dims, workers=workers)


def make_views(dims, num_views=3, kernel_size=31, n_sources=200, seed=20240607, workers=None):
    """Returns dict(psi0, views, kernels1, kernels2, weights, truth), float32."""
    dims = tuple(int(d) for d in dims)
    workers = workers or (os.cpu_count() or 1)
    ksz = [min(kernel_size, d if d % 2 == 1 else d - 1) for d in dims]
    rng = np.random.default_rng(seed)
    truth = np.full(dims, 10.0, dtype=np.float64)
    pos = np.stack([rng.integers(0, d, size=n_sources) for d in dims], axis=1)
    amp = rng.uniform(500.0, 5000.0, size=n_sources)
    np.add.at(truth, (pos[:, 0], pos[:, 1], pos[:, 2]), amp)
    views, k1s, k2s, ws = [], [], [], []
    for v in range(num_views):
        vrng = np.random.default_rng(seed + 1 + v)
        sig = _ORIENT[v % len(_ORIENT)]
        size = max(ksz)
        psf = gaussian_psf(size, sig)
        # crop to per-axis kernel extents (only matters for tiny volumes)
        sl = tuple(slice(size // 2 - k // 2, size // 2 - k // 2 + k) for k in ksz)
        psf = np.ascontiguousarray(psf[sl])
        psf = (psf / psf.sum()).astype(F32)
        blurred = _circ_conv64(truth, psf.astype(np.float64), workers)
        view = np.maximum(blurred + vrng.normal(0.0, 0.5, size=dims), 0.1).astype(F32)
        views.append(view)
        k1s.append(psf)
        k2s.append(np.ascontiguousarray(psf[::-1, ::-1, ::-1]))
        ws.append(np.full(dims, 1.0 / num_views, dtype=F32))
    psi0 = np.full(dims, views[0].mean(dtype=np.float64), dtype=F32)
    return dict(psi0=psi0, views=views, kernels1=k1s, kernels2=k2s, weights=ws,
                truth=truth.astype(F32))


def make_views_fast(dims, num_views=6, kernel_size=41, seed=20240607, workers=None, n_sources=None):
    """``make_views`` at BASELINE config-3 scale: the same protocol -- DC background 10, point
    sources U(500, 5000), the six PSF orientations, noise N(0, 0.5^2), floor 0.1, weights 1/V, psi0 = mean(view 0)
    -- with float32 FFTs for the blur so that 6 x 512x512x256 is generated in seconds (20 000 sources there)."""
    import scipy.fft as sfft

    dims = tuple(int(d) for d in dims)
    workers = workers or (os.cpu_count() or 1)
    rng = np.random.default_rng(seed)
    if n_sources is None:
        n_sources = max(50, int(np.prod(dims) // 3355))  # 20 000 at 512x512x256
    truth = np.full(dims, 10.0, dtype=F32)
    pos = [rng.integers(0, d, size=n_sources) for d in dims]
    np.add.at(truth, tuple(pos), rng.uniform(500.0, 5000.0, size=n_sources).astype(F32))
    tf = sfft.rfftn(truth, workers=workers)
    out = dict(views=[], kernels1=[], kernels2=[], weights=[])
    for v in range(num_views):
        psf = gaussian_psf(kernel_size, _ORIENT[v % len(_ORIENT)])
        pad = np.zeros(dims, dtype=F32)
        idx = []
        for ax in range(3):
            i = np.arange(kernel_size) - kernel_size // 2
            idx.append(np.where(i < 0, i + dims[ax], i))
        pad[np.ix_(*idx)] = psf
        blurred = sfft.irfftn(tf * sfft.rfftn(pad, workers=workers), s=dims, workers=workers)
        noise = np.random.default_rng(seed + 1 + v).standard_normal(dims, dtype=F32) * F32(0.5)
        out["views"].append(np.maximum(blurred + noise, F32(0.1)).astype(F32))
        out["kernels1"].append(psf)
        out["kernels2"].append(np.ascontiguousarray(psf[::-1, ::-1, ::-1]))
        out["weights"].append(np.full(dims, 1.0 / num_views, dtype=F32))
    out["psi0"] = np.full(dims, out["views"][0].mean(dtype=np.float64), dtype=F32)
    return out


def reference_bench_views(dims, num_views=6):
    """bench/synthetic_data.hpp:58-96 -- constant views, delta kernels."""
    dims = tuple(int(d) for d in dims)
    views, k1s, k2s, ws = [], [], [], []
    for i in range(num_views):
        views.append(np.full(dims, 16.0 + 4.0 * i, dtype=F32))
        ws.append(np.ones(dims, dtype=F32))
        k1 = np.zeros((21, 21, 21), dtype=F32)
        k1[10, 10, 10] = i + 1
        k2 = np.zeros((25, 25, 25), dtype=F32)
        k2[12, 12, 12] = i + 2
        k1s.append(k1)
        k2s.append(k2)
    return dict(psi0=views[0].copy(), views=views, kernels1=k1s, kernels2=k2s, weights=ws)


def staircase_dims(lo_exp=6, hi_exp=9):
    """python/generate_dims.py:4-48 (produce_size_strings(6, 10)) -- 64^3, 128x64x64,
    128x128x64, 128^3, ... 512^3 ({z,y,x}; the leading axes grow first)."""
    out = []
    for e in range(lo_exp, hi_exp):
        b = 1 << e
        out += [(b, b, b), (2 * b, b, b), (2 * b, 2 * b, b)]
    out.append((1 << hi_exp,) * 3)
    return out
