"""Block tiler with halo: host logic on the CPU.  The per-block deconvolution is injected; here
it is the oracle, so the test pins the tiling itself: with wrap padding and a halo of 2 * iterations * views * (k // 2)
the stitched result equals the untiled circular deconvolution."""
import numpy as np
import pytest

from libmultiviewnative_b200 import tiler
from libmultiviewnative_b200.synthetic import make_views
from oracle import mvn_oracle as orc
from tests import parity_cases as pc


def test_plan_covers_every_voxel_once():
    for vol, blk, halo in [((40, 33, 50), (16, 16, 32), (2, 3, 4)), ((16, 16, 16), (16, 16, 16), (0, 0, 0)),
                           ((70, 20, 20), (32, 20, 20), (5, 0, 0))]:
        count = np.zeros(vol, dtype=np.int32)
        for b in tiler.plan_blocks(vol, blk, halo):
            assert b.shape == tuple(blk)
            count[b.dst] += 1
            for a in range(3):
                assert b.keep_lo[a] >= halo[a] or b.start[a] + b.keep_lo[a] == 0 or vol[a] <= blk[a] - 2 * halo[a]
        assert (count == 1).all()


def test_halo_matches_reference_fixture_rule():
    # ref: tests/tiff_fixtures.hpp:241: offset = num_kernel_widths * (extent / 2)
    assert tiler.halo_for([(41, 41, 41), (21, 31, 11)], 1) == (20, 20, 20)
    assert tiler.halo_for([(4, 3, 2)], 2) == (4, 2, 2)


def test_extract_modes():
    v = np.arange(5 * 4 * 3, dtype=np.float32).reshape(5, 4, 3)
    b = tiler.Block(0, (-2, -1, 1), (6, 4, 4), (0, 0, 0), (6, 4, 4))
    w = tiler.extract(v, b, "wrap")
    assert w[0, 0, 0] == v[3, 3, 1] and w[2, 1, 2] == v[0, 0, 0]
    r = tiler.extract(v, b, "reflect")
    assert r[0, 1, 0] == v[2, 0, 1] and r[1, 0, 3] == v[1, 1, 0]
    z = tiler.extract(v, b, "zero")
    assert z[0, 0, 0] == 0 and z[2, 1, 0] == v[0, 0, 1] and z[2, 1, 3] == 0


@pytest.mark.parametrize("iterations", [1, 2])
def test_tiled_equals_untiled_with_sufficient_halo(iterations):
    dims, ks = (40, 36, 44), 3
    d = make_views(dims, num_views=2, kernel_size=ks, n_sources=12, workers=1)
    lam = 0.006
    exp = orc.inplace_cpu_deconvolve(d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], iterations, lam, 1e-4)
    halo = tuple(2 * iterations * 2 * (ks // 2) for _ in range(3))  # two views update psi in sequence

    def run(b):
        return orc.inplace_cpu_deconvolve(b["psi0"], b["views"], b["kernels1"], b["kernels2"], b["weights"], iterations, lam, 1e-4)

    got = tiler.deconvolve_tiled(run, d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], (24, 24, 32), halo,
                                 pad_mode="wrap")
    assert pc.max_rel(got, exp) < 1e-5


def test_default_halo_follows_the_fixture_rule():
    """halo=None: num_kernel_widths * (extent // 2) of the largest PSF (ref: tests/tiff_fixtures.hpp:241)"""
    dims = (24, 20, 28)
    d = make_views(dims, num_views=1, kernel_size=5, n_sources=6, workers=1)
    seen = []

    def run(b):
        seen.append(b["psi0"].shape)
        return b["psi0"] + 1

    got = tiler.deconvolve_tiled(run, d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], (16, 16, 16),
                                 num_kernel_widths=2, pad_mode="reflect")
    assert len(seen) == len(tiler.plan_blocks(dims, (16, 16, 16), (4, 4, 4))) and all(s == (16, 16, 16) for s in seen)
    np.testing.assert_array_equal(got, d["psi0"] + 1)


def test_shards_compose():
    """block b on rank b mod G (blocks.shard): the union of the shards' stitched interiors is the full result"""
    from libmultiviewnative_b200.blocks import shard

    dims = (16, 16, 24)
    d = make_views(dims, num_views=1, kernel_size=3, n_sources=6, workers=1)
    run = lambda b: orc.inplace_cpu_deconvolve(b["psi0"], b["views"], b["kernels1"], b["kernels2"], b["weights"], 1, 0.0, 1e-4)
    args = (d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], (12, 12, 12), (2, 2, 2))
    full = tiler.deconvolve_tiled(run, *args, pad_mode="reflect")
    n = len(tiler.plan_blocks(dims, (12, 12, 12), (2, 2, 2)))
    merged = d["psi0"].copy()
    for rank in range(2):
        part = tiler.deconvolve_tiled(run, *args, pad_mode="reflect", blocks=shard(n, rank, 2))
        changed = part != d["psi0"]
        merged[changed] = part[changed]
    np.testing.assert_array_equal(merged, full)
