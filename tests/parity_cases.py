"""Parity cases shared by the GPU tests (real library, `-m gpu`) and the CPU-only
emulator tests (tests/emu build of the same sources).  Every case goes through
the C ABI (libmultiviewnative_b200.capi) and is checked against the oracle or a
committed golden fixture.  Tolerances are the north star's:
per-voxel relative <= 1e-4 after one iteration, relative L2 <= 1e-3 after ten."""
import os

import numpy as np

from libmultiviewnative_b200.synthetic import make_views
from oracle import mvn_oracle as orc

F32 = np.float32
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PER_VOXEL_TOL_1_ITER = 1e-4
REL_L2_TOL_10_ITER = 1e-3


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))


def max_rel(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b) / np.abs(b.astype(np.float64))))


def fixture():
    return np.load(os.path.join(GOLDEN, "conv_fixture_8.npz"))


def pointwise():
    return np.load(os.path.join(GOLDEN, "pointwise_cases.npz"))


# ---- FFT numerics (ref: tests/test_plan_store.cpp:83-111, test_fftw_numerical_stability.cpp) ----
def case_fft_round_trip(L, dims):
    n = int(np.prod(dims))
    a = np.arange(n, dtype=F32).reshape(dims)
    spec = L.rfftn(a)
    ref = orc.fft_forward(a)
    assert np.max(np.abs(spec - ref)) <= 2e-6 * np.max(np.abs(ref)) * max(1.0, np.log2(n))
    back = L.irfftn(spec, dims) * F32(1.0 / n)
    mse = float(np.mean(((back - a) / a.max()) ** 2))
    assert mse < 1e-4
    if dims == (8, 8, 8):
        np.testing.assert_array_equal(np.floor(back + 0.5), a)


# ---- convolution API (ref: tests/test_gpu_convolve.cpp:197-328, test_cpu_asymm_convolve.cpp) ----
def case_conv_fixture(L, name, entry="inplace_gpu_convolution"):
    fx = fixture()
    im = fx["padded_image"].copy()
    getattr(L, entry)(im, fx["kernel_" + name].copy())
    got = im[1:9, 1:9, 1:9]
    if name == "trivial":
        assert abs(float(im.sum())) < 1e-3
        return
    exp = fx["image"] if name == "identity" else fx["image_folded_by_" + name]
    assert float(got.sum(dtype=np.float64)) == np.float64(exp.sum(dtype=np.float64)) or \
        abs(float(got.sum(dtype=np.float64)) / float(exp.sum(dtype=np.float64)) - 1) < 1e-5
    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=2e-3)


def case_conv_impulse(L, name):
    fx = fixture()
    kern = fx["kernel_" + name].copy()
    im = fx["padded_one"].copy()
    L.inplace_gpu_convolution(im, kern)
    one = im[1:9, 1:9, 1:9]
    assert abs(float(one.sum()) / float(kern.sum()) - 1) < 1e-5
    lo = [one.shape[i] // 2 - kern.shape[i] // 2 for i in range(3)]
    seg = one[lo[0]:lo[0] + kern.shape[0], lo[1]:lo[1] + kern.shape[1], lo[2]:lo[2] + kern.shape[2]]
    np.testing.assert_array_equal(np.floor(seg + 0.5), kern)


def case_conv_identity_asymmetric_image(L):
    dims = (16, 18, 14)  # ref: tests/test_gpu_convolve_impl.cu:422-530, abs < 1e-3 on iota
    img = np.arange(np.prod(dims), dtype=F32).reshape(dims)
    k = np.zeros((3, 3, 3), dtype=F32)
    k[1, 1, 1] = 1
    out = img.copy()
    L.inplace_gpu_convolution(out, k)
    assert np.max(np.abs(out - img)) < 1e-3 * max(1.0, float(img.max()) / 1024)


def case_conv_random_vs_oracle(L, dims, kdims, seed=0):
    rng = np.random.default_rng(seed)
    img = (rng.random(dims) + 1).astype(F32)
    k = rng.random(kdims).astype(F32)
    k /= k.sum()
    exp = orc.inplace_cpu_convolution(img, k)
    out = img.copy()
    L.inplace_gpu_convolution(out, k)
    assert rel_l2(out, exp) <= 1e-5
    assert max_rel(out, exp) <= 1e-4


def case_conv_rejects_oversized_kernel(L):
    from libmultiviewnative_b200.capi import LmvnError

    img = np.ones((8, 8, 8), F32)
    before = img.copy()
    try:
        L.inplace_gpu_convolution(img, np.ones((9, 3, 3), F32))
    except LmvnError:
        np.testing.assert_array_equal(img, before)  # output untouched on failure
        return
    raise AssertionError("oversized kernel was not rejected")


# ---- pointwise (ref: tests/test_gpu_kernels_impl.cu) ----
def case_pointwise(L):
    pw = pointwise()
    out = pw["divide_out"].copy()
    L.compute_quotient(pw["divide_in"].copy(), out)
    np.testing.assert_array_equal(out, pw["divide_expected"])
    n = 256 * 255 + 3  # ragged tail on purpose (ref uses 256x255x257)
    psi = np.full(n, 5.0, F32)
    L.compute_final_values(psi, np.full(n, 42.0, F32), np.full(n, 0.1, F32), 1e-4, 0.0)
    np.testing.assert_array_equal(psi, np.full(n, 25.5, F32))
    psi = np.full(n, 5.0, F32)
    L.compute_final_values(psi, np.full(n, 42.0, F32), np.full(n, 0.1, F32), 1e-4, 0.006)
    np.testing.assert_allclose(psi, np.full(n, pw["const_expected_reg"][0], F32), rtol=3e-7)
    psi = pw["rand_psi"].copy()
    L.compute_final_values(psi, pw["rand_integral"].copy(), pw["rand_weight"].copy(), 1e-4, 0.0)
    np.testing.assert_array_equal(psi, pw["rand_expected_plain"])
    psi = pw["rand_psi"].copy()
    L.compute_final_values(psi, pw["rand_integral"].copy(), pw["rand_weight"].copy(), 1e-4, 0.006)
    # float32 cancellation-free Tikhonov vs the reference's double evaluation
    np.testing.assert_allclose(psi, pw["rand_expected_reg"], rtol=1e-6, atol=1e-7)
    # NaN / Inf handling
    psi = np.array([1.0, 1.0, 3e38, 2.0], F32)
    L.compute_final_values(psi, np.array([np.nan, -np.inf, 3e38, 0.0], F32), np.ones(4, F32), 1e-3, 0.006)
    assert np.isfinite(psi).all() and np.allclose(psi[[0, 1, 3]], 1e-3, atol=1e-6)
    # special values of the quotient, identical on the device (MUFU .ftz forms) and in the emulated build, which
    # flushes subnormals the same way: 1/0 = inf like the reference; a SUBNORMAL blurred value is flushed to zero (the
    # reference's double reciprocal still gives a finite 8.5e37 .. 3.4e38 above 2.9e-39: documented deviation)
    out = np.array([0.0, -0.0, 1e-39, 1.17549435e-38, np.inf, np.nan], F32)
    L.compute_quotient(np.ones(6, F32), out)
    with np.errstate(divide="ignore"):
        ref = orc.compute_quotient(np.ones(6, F32), np.array([0.0, -0.0, 1e-39, 1.17549435e-38, np.inf, np.nan], F32))
    assert out[0] == np.inf and out[1] == -np.inf and out[2] == np.inf and out[4] == 0.0 and np.isnan(out[5])
    np.testing.assert_allclose(out[3], ref[3], rtol=3e-7)
    assert ref[0] == np.inf and ref[4] == 0.0 and np.isnan(ref[5])


# ---- deconvolution (config-1 protocol, reduced size for CPU-side runs) ----
def case_deconvolve_vs_oracle(L, dims, nviews, ksize, lam, iters_list=(1, 10), seed=20240607, n_sources=40):
    d = make_views(dims, num_views=nviews, kernel_size=ksize, n_sources=n_sources, seed=seed, workers=4)
    res = {}
    for iters in iters_list:
        exp = orc.inplace_cpu_deconvolve(d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], iters, lam,
                                         1e-4, nthreads=4)
        psi = d["psi0"].copy()
        L.inplace_gpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], iters, lam, 1e-4)
        res[iters] = (max_rel(psi, exp), rel_l2(psi, exp))
        if iters == 1:
            assert res[iters][0] <= PER_VOXEL_TOL_1_ITER, res
        assert res[iters][1] <= REL_L2_TOL_10_ITER, res
    return res


def case_zero_iterations(L):
    d = make_views((8, 10, 12), num_views=2, kernel_size=5, n_sources=5, workers=1)
    psi = d["psi0"].copy()
    L.inplace_gpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], 0, 0.006, 1e-4)
    np.testing.assert_array_equal(psi, d["psi0"])


def case_deterministic(L):
    d = make_views((16, 16, 16), num_views=2, kernel_size=5, n_sources=8, workers=1)
    outs = []
    for _ in range(2):
        psi = d["psi0"].copy()
        L.inplace_gpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4)
        outs.append(psi)
    np.testing.assert_array_equal(outs[0], outs[1])


def case_mismatched_views_rejected(L):
    from libmultiviewnative_b200.capi import LmvnError

    d = make_views((8, 8, 8), num_views=2, kernel_size=3, n_sources=3, workers=1)
    views = [d["views"][0], np.ones((8, 8, 6), F32)]
    weights = [d["weights"][0], np.ones((8, 8, 6), F32)]
    psi = d["psi0"].copy()
    try:
        L.inplace_gpu_deconvolve(psi, views, d["kernels1"], d["kernels2"], weights, 1, 0.0, 1e-4)
    except LmvnError:
        np.testing.assert_array_equal(psi, d["psi0"])
        return
    raise AssertionError("mismatched view dims were not rejected")


def case_plan_resume_equals_one_shot(L):
    """Calling again with the returned psi resumes (ref: bench/bench_gpu_deconvolve.cu:48-49);
    the persistent handle must give the same bits as the one-shot entry point."""
    d = make_views((16, 12, 20), num_views=3, kernel_size=5, n_sources=8, workers=1)
    one = d["psi0"].copy()
    L.inplace_gpu_deconvolve(one, d["views"], d["kernels1"], d["kernels2"], d["weights"], 3, 0.006, 1e-4)
    with L.plan(d["psi0"].shape, 3) as p:
        for v in range(3):
            p.set_view(v, d["views"][v], d["weights"][v], d["kernels1"][v], d["kernels2"][v])
        p.set_psi(d["psi0"])
        p.iterate(1, 0.006, 1e-4)
        p.iterate(2, 0.006, 1e-4)
        got = p.get_psi()
        info = p.info()
    np.testing.assert_array_equal(got, one)
    assert info.num_views == 3 and tuple(info.dims) == (16, 12, 20)
    assert info.alg_bytes_per_view_iteration == 7 * info.real_bytes + 10 * info.spectrum_bytes


def case_legacy_iterate(L):
    d = make_views((8, 8, 16), num_views=1, kernel_size=5, n_sources=4, workers=1)
    inp, k1 = d["views"][0], d["kernels1"][0]
    out = L.iterate_fft_plain(inp, k1)
    k2 = np.full(k1.shape, 0.1, F32)
    exp = orc.inplace_cpu_deconvolve(inp, [inp], [k1], [k2], [np.ones_like(inp)], 1, 0.0, 1e-4)
    assert max_rel(out, exp) <= 1e-4
    # like the reference, iterate_fft_tikhonov ignores its minValue / lambda arguments and uses 1e-4 / 0.2
    # (ref: src/multiviewnative.cu:582-583); with the hard-coded unit weights its older update rule equals the main one
    out = L.iterate_fft_tikhonov(inp, k1, 0.5, 0.006)
    exp = orc.inplace_cpu_deconvolve(inp, [inp], [k1], [k2], [np.ones_like(inp)], 1, 0.2, 1e-4)
    assert max_rel(out, exp) <= 1e-4
    t = inp.astype(np.float64) * orc.inplace_cpu_convolution(orc.compute_quotient(inp, orc.inplace_cpu_convolution(inp, k1)), k2)
    old_rule = np.maximum((np.sqrt(1.0 + 2.0 * 0.2 * t) - 1.0) / 0.2, 1e-4)  # ref: inc/cuda_kernels.cuh:161-194 with w = 1
    assert max_rel(out, old_rule) <= 1e-4
    im = inp.copy()
    L.convolution3DfftCUDAInPlace(im, k1)
    assert rel_l2(im, orc.inplace_cpu_convolution(inp, k1)) <= 1e-5


GEOMETRY_NATIVE, GEOMETRY_EMBEDDED, GEOMETRY_ZERO_PADDED = 1, 2, 3


# ---- zero_padd mode (ref: inc/padd_utils.h:102-249; the reference's GPU geometry) ----------------
def _zero_pad(a, kmax):
    ext = [a.shape[i] + kmax[i] - 1 for i in range(3)]
    off = [(kmax[i] - 1) // 2 for i in range(3)]
    out = np.zeros(ext, dtype=F32)
    out[off[0]:off[0] + a.shape[0], off[1]:off[1] + a.shape[1], off[2]:off[2] + a.shape[2]] = a
    return out, off


def _crop(a, off, shape):
    return np.ascontiguousarray(a[off[0]:off[0] + shape[0], off[1]:off[1] + shape[1], off[2]:off[2] + shape[2]])


def case_zero_padd_convolution(L, dims, kdims):
    """zero mode == circular convolution of the explicitly zero-padded image (image + kernel - 1, offset
    (kernel - 1) / 2), cropped: no wrap-around inside the image."""
    rng = np.random.default_rng(21)
    img = (rng.random(dims, dtype=F32) + 1).astype(F32)
    k = rng.random(kdims, dtype=F32)
    k /= k.sum()
    padded, off = _zero_pad(img, kdims)
    exp = _crop(orc.inplace_cpu_convolution(padded, k), off, dims)
    L.set_padding(1)
    try:
        got = img.copy()
        L.inplace_gpu_convolution(got, k)
    finally:
        L.set_padding(0)
    assert rel_l2(got, exp) < 1e-5
    circ = img.copy()
    L.inplace_gpu_convolution(circ, k)
    assert rel_l2(circ, exp) > 1e-3  # the circular default differs at the borders


def case_zero_padd_deconvolve(L, dims, ksize, lam=0.006, iters=2):
    """zero mode == the deconvolution of the explicitly zero-padded stacks, cropped.  Quotient rule in the padding:
    a zero of the view gives zero (oracle zero_view_guard).  Where the reference's own arithmetic (0 * (1 / blurred),
    NaN for blurred == 0) stays finite the two rules give the same numbers, and that is checked as well; where it does
    not (Gaussian PSFs underflow in the corners of the padding) the reference collapses to one constant per stack."""
    d = make_views(dims, num_views=2, kernel_size=ksize, n_sources=8, workers=1)
    kmax = [max(k.shape[a] for k in d["kernels1"] + d["kernels2"]) for a in range(3)]
    pv, pw = [], []
    for v, w in zip(d["views"], d["weights"]):
        a, off = _zero_pad(v, kmax)
        pv.append(a)
        pw.append(_zero_pad(w, kmax)[0])
    ppsi, off = _zero_pad(d["psi0"], kmax)
    exp = _crop(orc.inplace_cpu_deconvolve(ppsi, pv, d["kernels1"], d["kernels2"], pw, iters, lam, 1e-4,
                                           zero_view_guard=True), off, dims)
    assert np.isfinite(exp).all() and exp.max() > 1.5 * exp.min()  # a deconvolution, not the NaN collapse
    L.set_padding(1)
    try:
        got = d["psi0"].copy()
        L.inplace_gpu_deconvolve(got, d["views"], d["kernels1"], d["kernels2"], d["weights"], iters, lam, 1e-4)
    finally:
        L.set_padding(0)
    assert max_rel(got, exp) < PER_VOXEL_TOL_1_ITER
    ref = _crop(orc.inplace_cpu_deconvolve(ppsi, pv, d["kernels1"], d["kernels2"], pw, iters, lam, 1e-4), off, dims)
    if np.isfinite(ref).all() and ref.max() > 1.5 * ref.min():  # the reference's arithmetic stayed clean: same numbers
        assert max_rel(got, ref) < PER_VOXEL_TOL_1_ITER


def case_zero_padd_unreached_padding(L, dims=(12, 12, 28), k=5, iters=3, lam=0.006):
    """Constant views and box kernels on extents that round up to a fast-path size (x: 32 -> 64): part of the
    padding is out of reach of every kernel tap, the blurred estimate there is round-off noise with exact zeros, and
    0 * (1 / 0) must not poison the stack.  Expected: the same deconvolution at the reference's own extents
    (image + kernel - 1, where every padding voxel is within reach)."""
    nv = 2
    views = [np.full(dims, 16 + 4 * v, dtype=F32) for v in range(nv)]
    weights = [np.ones(dims, dtype=F32) for _ in range(nv)]
    k1 = [np.full((k, k, k), (v + 1) / k ** 3, dtype=F32) for v in range(nv)]
    k2 = [np.full((k, k, k), (v + 2) / k ** 3, dtype=F32) for v in range(nv)]
    psi0 = np.full(dims, 16, dtype=F32)
    kmax = [k, k, k]
    ppsi, off = _zero_pad(psi0, kmax)
    exp = _crop(orc.inplace_cpu_deconvolve(ppsi, [_zero_pad(v, kmax)[0] for v in views], k1, k2,
                                           [_zero_pad(w, kmax)[0] for w in weights], iters, lam, 1e-3), off, dims)
    assert np.isfinite(exp).all() and exp.min() > 0.1
    L.set_padding(1)
    try:
        got = psi0.copy()
        L.inplace_gpu_deconvolve(got, views, k1, k2, weights, iters, lam, 1e-3)
        assert L.last_geometry() == GEOMETRY_ZERO_PADDED
    finally:
        L.set_padding(0)
    assert max_rel(got, exp) < 1e-4


# ---- periodic embedding: circular semantics on the fast path for arbitrary extents ---------------


def case_embedded_convolution(L, dims, kdims, expect_embedded=True):
    rng = np.random.default_rng(31)
    img = (rng.random(dims, dtype=F32) + 1).astype(F32)
    k = rng.random(kdims, dtype=F32)
    k /= k.sum()
    exp = orc.inplace_cpu_convolution(img, k)
    got = img.copy()
    L.inplace_gpu_convolution(got, k)
    assert L.last_geometry() == (GEOMETRY_EMBEDDED if expect_embedded else GEOMETRY_NATIVE)
    assert rel_l2(got, exp) < 1e-5
    assert float(np.max(np.abs(got - exp))) < 1e-4 * float(np.max(np.abs(exp)))


def case_embedded_deconvolve(L, dims, ksize, lam=0.006, iters_list=(1, 3)):
    d = make_views(dims, num_views=2, kernel_size=ksize, n_sources=10, workers=1)
    for iters in iters_list:
        exp = orc.inplace_cpu_deconvolve(d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], iters, lam, 1e-4)
        got = d["psi0"].copy()
        L.inplace_gpu_deconvolve(got, d["views"], d["kernels1"], d["kernels2"], d["weights"], iters, lam, 1e-4)
        assert L.last_geometry() == GEOMETRY_EMBEDDED
        if iters == 1:
            assert max_rel(got, exp) < PER_VOXEL_TOL_1_ITER
        assert rel_l2(got, exp) < REL_L2_TOL_10_ITER


def case_embedded_plan_equals_one_shot(L, dims=(30, 28, 40), ksize=5):
    """persistent handle with periodic embedding == the one-shot call that embeds by itself"""
    d = make_views(dims, num_views=2, kernel_size=ksize, n_sources=10, workers=1)
    one = d["psi0"].copy()
    L.inplace_gpu_deconvolve(one, d["views"], d["kernels1"], d["kernels2"], d["weights"], 4, 0.006, 1e-4)
    assert L.last_geometry() == GEOMETRY_EMBEDDED
    kmax = [max(k.shape[a] for k in d["kernels1"] + d["kernels2"]) for a in range(3)]
    with L.plan(dims, 2, 0, max_kernel_dims=kmax, geometry="embedded") as p:
        info = p.info()
        assert tuple(info.dims) != tuple(dims) and all(n >= m + k - 1 for n, m, k in zip(info.dims, dims, kmax))
        for v in range(2):
            p.set_view(v, d["views"][v], d["weights"][v], d["kernels1"][v], d["kernels2"][v])
        p.set_psi(d["psi0"])
        p.iterate(2, 0.006, 1e-4)
        p.iterate(2, 0.006, 1e-4)
        got = p.get_psi()
    # a call ends with the plain x-inverse kernel and the next one starts with the plain x-forward kernel instead of the
    # fused link: same arithmetic, but another compilation (FMA contraction) -- identical to a few ulp on the GPU
    assert max_rel(got, one) < 5e-6
