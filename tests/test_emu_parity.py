"""CPU-only: the library's sources compiled against the host emulator
(tests/emu/cuda_emu.h) and driven through the same C ABI and the same parity
cases as the GPU tests.  This checks kernel index math and host logic without
a GPU; it says nothing about performance and is never the product path."""
import numpy as np
import pytest

from tests import parity_cases as pc


@pytest.fixture(scope="module")
def L():
    from libmultiviewnative_b200._build import build_emu
    from libmultiviewnative_b200.capi import Library

    return Library(build_emu())


def test_emu_is_labelled(L):
    assert "emulation" in L.version()


@pytest.mark.parametrize("dims", [(8, 8, 8), (13, 17, 19), (16, 16, 16), (14, 14, 14), (9, 5, 25)])
def test_fft_round_trip(L, dims):
    pc.case_fft_round_trip(L, dims)


@pytest.mark.parametrize("name", ["trivial", "identity", "horizontal", "vertical", "depth", "all1"])
def test_conv_fixture(L, name):
    pc.case_conv_fixture(L, name)


@pytest.mark.parametrize("name", ["asymm_cross", "asymm_one", "asymm_identity"])
def test_conv_impulse(L, name):
    pc.case_conv_impulse(L, name)


def test_conv_identity_asymmetric_image(L):
    pc.case_conv_identity_asymmetric_image(L)


@pytest.mark.parametrize("dims,kdims", [((16, 16, 32), (5, 5, 5)), ((12, 10, 14), (4, 3, 2)), ((8, 8, 8), (8, 8, 8))])
def test_conv_random(L, dims, kdims):
    pc.case_conv_random_vs_oracle(L, dims, kdims)


def test_conv_rejects_oversized_kernel(L):
    pc.case_conv_rejects_oversized_kernel(L)


@pytest.mark.parametrize("dims,kdims", [((12, 10, 14), (4, 3, 2)), ((20, 24, 50), (5, 7, 9))])
def test_zero_padd_convolution(L, dims, kdims):
    pc.case_zero_padd_convolution(L, dims, kdims)


def test_zero_padd_deconvolve(L):
    pc.case_zero_padd_deconvolve(L, (10, 12, 14), 5)


def test_zero_padd_padding_out_of_reach_of_the_kernels(L):
    pc.case_zero_padd_unreached_padding(L)


@pytest.mark.parametrize("dims,kdims", [((20, 24, 50), (5, 7, 9)), ((28, 30, 50), (4, 3, 2)), ((27, 27, 27), (3, 3, 3))])
def test_embedded_convolution(L, dims, kdims):
    pc.case_embedded_convolution(L, dims, kdims)


def test_embedding_not_used_when_it_costs_too_much(L, monkeypatch):
    pc.case_embedded_convolution(L, (9, 10, 15), (3, 3, 3), expect_embedded=False)  # 16 x 16 x 64 would be 12x the voxels
    monkeypatch.setenv("LMVN_EMBED", "0")
    pc.case_embedded_convolution(L, (20, 24, 50), (5, 7, 9), expect_embedded=False)


def test_embedded_deconvolve(L):
    pc.case_embedded_deconvolve(L, (30, 28, 40), 5)


def test_embedded_plan_equals_one_shot(L):
    pc.case_embedded_plan_equals_one_shot(L)


def test_pointwise(L):
    pc.case_pointwise(L)


@pytest.mark.parametrize("lam", [0.0, 0.006])
def test_deconvolve_vs_oracle(L, lam):
    pc.case_deconvolve_vs_oracle(L, (16, 16, 32), 3, 7, lam, iters_list=(1, 3), n_sources=10)


def test_deconvolve_odd_dims(L):
    pc.case_deconvolve_vs_oracle(L, (9, 10, 15), 2, 5, 0.006, iters_list=(1,), n_sources=6)


def test_zero_iterations(L):
    pc.case_zero_iterations(L)


def test_deterministic(L):
    pc.case_deterministic(L)


def test_mismatched_views_rejected(L):
    pc.case_mismatched_views_rejected(L)


def test_plan_resume_equals_one_shot(L):
    pc.case_plan_resume_equals_one_shot(L)


def test_legacy_entry_points(L):
    pc.case_legacy_iterate(L)


def test_cpu_entry_points_match_oracle(L):
    """inplace_cpu_* are link-compat symbols (host FFT); they must agree with the oracle."""
    from libmultiviewnative_b200.synthetic import make_views
    from oracle import mvn_oracle as orc

    d = make_views((12, 16, 20), num_views=2, kernel_size=5, n_sources=6, workers=1)
    psi = d["psi0"].copy()
    L.inplace_cpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4, nthreads=2)
    exp = orc.inplace_cpu_deconvolve(d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4)
    assert pc.max_rel(psi, exp) < 1e-4
    im = d["views"][0].copy()
    L.inplace_cpu_convolution(im, d["kernels1"][0], nthreads=1)
    assert pc.rel_l2(im, orc.inplace_cpu_convolution(d["views"][0], d["kernels1"][0])) < 1e-5


# ---- power-of-two fast path (strategy "fused"): every template instantiation ----
@pytest.fixture()
def LF(L):
    from libmultiviewnative_b200 import capi

    L.set_default_strategy(capi.STRATEGY_FUSED)
    yield L
    L.set_default_strategy(capi.STRATEGY_AUTO)


@pytest.mark.parametrize("dims,kdims", [
    ((16, 16, 64), (5, 5, 5)), ((32, 64, 128), (7, 4, 3)), ((64, 16, 256), (3, 3, 9)), ((16, 128, 64), (5, 5, 5)),
    ((256, 16, 64), (5, 5, 5)), ((16, 512, 64), (3, 9, 3)), ((512, 16, 64), (9, 3, 3)), ((16, 16, 64), (16, 16, 64)),
    ((16, 16, 512), (3, 3, 11)), ((16, 1024, 64), (3, 9, 3)), ((1024, 16, 64), (9, 3, 3)), ((16, 16, 1024), (3, 5, 13)),
])
def test_fused_conv_random(LF, dims, kdims):
    pc.case_conv_random_vs_oracle(LF, dims, kdims)


def test_fused_strategy_is_selected_and_rejected(L, LF):
    from libmultiviewnative_b200 import capi

    with LF.plan((16, 16, 64), 1) as p:
        assert p.info().strategy == capi.STRATEGY_FUSED
    with pytest.raises(capi.LmvnError):
        LF.plan((10, 10, 10), 1)  # not a power of two: no fused plan
    L.set_default_strategy(capi.STRATEGY_AUTO)
    with L.plan((10, 10, 10), 1) as p:
        assert p.info().strategy == capi.STRATEGY_GENERIC
    with L.plan((16, 16, 64), 1) as p:
        assert p.info().strategy == capi.STRATEGY_FUSED


@pytest.mark.parametrize("lam", [0.0, 0.006])
def test_fused_deconvolve_vs_oracle(LF, lam):
    pc.case_deconvolve_vs_oracle(LF, (16, 32, 64), 3, 7, lam, iters_list=(1, 3), n_sources=10)


def test_fused_deconvolve_nx1024(LF):
    pc.case_deconvolve_vs_oracle(LF, (16, 16, 1024), 2, 5, 0.006, iters_list=(1,), n_sources=6)


def test_fused_deconvolve_nx128(LF):
    pc.case_deconvolve_vs_oracle(LF, (16, 16, 128), 2, 5, 0.006, iters_list=(1,), n_sources=6)


def test_fused_equals_generic_closely(L):
    from libmultiviewnative_b200 import capi
    from libmultiviewnative_b200.synthetic import make_views

    d = make_views((32, 16, 64), num_views=2, kernel_size=5, n_sources=8, workers=1)
    outs = {}
    for strat in (capi.STRATEGY_GENERIC, capi.STRATEGY_FUSED):
        L.set_default_strategy(strat)
        psi = d["psi0"].copy()
        L.inplace_gpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4)
        outs[strat] = psi
    L.set_default_strategy(capi.STRATEGY_AUTO)
    assert pc.max_rel(outs[capi.STRATEGY_FUSED], outs[capi.STRATEGY_GENERIC]) < 1e-5


def test_embedded_chained_equals_unchained(L, monkeypatch):
    """the chained link kernel continues the stack periodically by itself (aliased rows recompute their interior row,
    out-of-place spectrum and psi): bit-identical to refilling the exterior before every five-pass convolution"""
    from libmultiviewnative_b200.synthetic import make_views

    dims = (30, 28, 40)
    d = make_views(dims, num_views=3, kernel_size=5, n_sources=10, workers=1)
    outs = {}
    for chain in ("1", "0"):
        monkeypatch.setenv("LMVN_CHAIN", chain)
        psi = d["psi0"].copy()
        L.inplace_gpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], 3, 0.006, 1e-4)
        assert L.last_geometry() == pc.GEOMETRY_EMBEDDED
        outs[chain] = psi
    np.testing.assert_array_equal(outs["1"], outs["0"])


# ---- two-pass schedule (csrc/fft_x3.cuh): plane-tile pass + z-middle pass with the cross-lane y level --------------
@pytest.mark.parametrize("dims", [(64, 256, 256)])
def test_two_pass_schedule_deconvolve(L, dims, monkeypatch):
    """smallest shape the two-pass schedule takes (ny = 128 * 2): one view, two iterations = plane pass in its three
    modes (begin, chained quotient / update links, end), the z-middle pass and the K^ permutation, against the
    oracle; then the same call on the default five-pass schedule (LMVN_X3=0) must agree to round-off."""
    from libmultiviewnative_b200.synthetic import make_views
    from oracle import mvn_oracle as orc

    d = make_views(dims, num_views=1, kernel_size=9, n_sources=40, workers=4)
    exp = orc.inplace_cpu_deconvolve(d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4, nthreads=4)
    monkeypatch.setenv("LMVN_X3", "1")  # opt-in schedule, read when a plan is created
    psi = d["psi0"].copy()
    L.inplace_gpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4)
    assert pc.max_rel(psi, exp) <= pc.PER_VOXEL_TOL_1_ITER
    monkeypatch.setenv("LMVN_X3", "0")
    old = d["psi0"].copy()
    L.inplace_gpu_deconvolve(old, d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4)
    assert 0 < pc.max_rel(psi, old) < 5e-6  # different factorisations: close, not identical (identical = the knob did nothing)
