"""CPU-only: the slab-decomposed (multi-GPU) plans driven as in-process groups on the host
emulator.  The decomposed transform runs the same butterflies on the same data as the
single-device fast path, only with scattered final stores, so the results must be IDENTICAL
to the single-device plan bit for bit, and within the north star's tolerance of the oracle."""
import numpy as np
import pytest

from tests import parity_cases as pc


@pytest.fixture(scope="module")
def L():
    from libmultiviewnative_b200._build import build_emu
    from libmultiviewnative_b200.capi import Library

    return Library(build_emu())


def _run_group(L, d, dims, world, iters, lam):
    from libmultiviewnative_b200.slabs import LocalSlabGroup

    nv = len(d["views"])
    with LocalSlabGroup(L, dims, nv, world) as g:
        info = g.ranks[0].info()
        assert info.world == world and info.planes_per_rank == dims[0] // world and info.rows_per_rank == dims[1] // world
        for v in range(nv):
            g.set_view(v, d["views"][v], d["weights"][v], d["kernels1"][v], d["kernels2"][v])
        g.set_psi(d["psi0"])
        g.iterate(iters, lam, 1e-4)
        return g.get_psi()


@pytest.mark.parametrize("dims,world", [((32, 32, 64), 2), ((32, 32, 64), 4), ((16, 64, 128), 2), ((64, 16, 64), 8)])
def test_group_equals_single_plan_and_oracle(L, dims, world):
    from libmultiviewnative_b200 import capi
    from libmultiviewnative_b200.synthetic import make_views
    from oracle import mvn_oracle as orc

    d = make_views(dims, num_views=2, kernel_size=5, n_sources=8, workers=1)
    lam, iters = 0.006, 2
    got = _run_group(L, d, dims, world, iters, lam)
    L.set_default_strategy(capi.STRATEGY_FUSED)
    try:
        single = d["psi0"].copy()
        L.inplace_gpu_deconvolve(single, d["views"], d["kernels1"], d["kernels2"], d["weights"], iters, lam, 1e-4)
    finally:
        L.set_default_strategy(capi.STRATEGY_AUTO)
    np.testing.assert_array_equal(got, single)
    exp = orc.inplace_cpu_deconvolve(d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], iters, lam, 1e-4)
    assert pc.max_rel(got, exp) < pc.PER_VOXEL_TOL_1_ITER


def test_group_asymmetric_kernels(L):
    """even-sized anisotropic PSFs exercise the wrap-around with a plane offset (ref: inc/padd_utils.h:11-40)"""
    from libmultiviewnative_b200.slabs import LocalSlabGroup
    from oracle import mvn_oracle as orc

    dims = (32, 32, 64)
    rng = np.random.default_rng(3)
    img = (rng.random(dims, dtype=np.float32) + 1).astype(np.float32)
    w = np.full(dims, 0.5, np.float32)
    k1 = rng.random((4, 3, 2), dtype=np.float32)
    k2 = rng.random((6, 5, 8), dtype=np.float32)
    k1 /= k1.sum(); k2 /= k2.sum()
    with LocalSlabGroup(L, dims, 1, 4) as g:
        g.set_view(0, img, w, k1, k2)
        g.set_psi(img)
        g.iterate(1, 0.0, 1e-4)
        got = g.get_psi()
    exp = orc.inplace_cpu_deconvolve(img, [img], [k1], [k2], [w], 1, 0.0, 1e-4)
    assert pc.max_rel(got, exp) < pc.PER_VOXEL_TOL_1_ITER


def test_rejects_bad_decompositions(L):
    from libmultiviewnative_b200 import capi
    from libmultiviewnative_b200.slabs import SlabPlan

    with pytest.raises(capi.LmvnError):
        SlabPlan(L, (32, 32, 64), 1, 0, 3)      # world not a power of two
    with pytest.raises(capi.LmvnError):
        SlabPlan(L, (10, 10, 10), 1, 0, 2)      # not a fast-path shape
    with pytest.raises(capi.LmvnError):
        SlabPlan(L, (32, 32, 64), 1, 2, 2)      # rank out of range
    p = SlabPlan(L, (32, 32, 64), 1, 0, 2)
    with pytest.raises(capi.LmvnError):         # peers not connected
        p.conv_phase(0, 1, 0, 0.0, 1e-4)
    p.close()
