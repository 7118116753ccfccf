"""Float32 multi-page TIFF stacks in the reference's fixture layout, cross-checked with Pillow."""
import os

import numpy as np
import pytest

from libmultiviewnative_b200 import tiffstack


def test_round_trip_exact(tmp_path):
    rng = np.random.default_rng(1)
    a = rng.standard_normal((5, 7, 9)).astype(np.float32)
    p = str(tmp_path / "s.tif")
    tiffstack.write_stack(p, a)
    np.testing.assert_array_equal(tiffstack.read_stack(p), a)


def test_pillow_reads_what_we_write_and_vice_versa(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(2)
    a = rng.random((4, 6, 5), dtype=np.float32)
    p = str(tmp_path / "ours.tif")
    tiffstack.write_stack(p, a)
    with Image.open(p) as im:
        assert im.n_frames == 4 and im.mode == "F"
        for z in range(4):
            im.seek(z)
            np.testing.assert_array_equal(np.asarray(im), a[z])
    q = str(tmp_path / "pil.tif")
    frames = [Image.fromarray(a[z], mode="F") for z in range(4)]
    frames[0].save(q, save_all=True, append_images=frames[1:])
    np.testing.assert_array_equal(tiffstack.read_stack(q), a)
    u = str(tmp_path / "u16.tif")
    b = (rng.random((3, 4, 4)) * 60000).astype(np.uint16)
    fr = [Image.fromarray(b[z]) for z in range(3)]
    fr[0].save(u, save_all=True, append_images=fr[1:])
    np.testing.assert_array_equal(tiffstack.read_stack(u), b.astype(np.float32))


def test_fixture_layout_runs_through_the_oracle(tmp_path):
    """views written as input_view_i.tif / kernel{1,2}_view_i.tif / weights_view_i.tif and psi_i.tif golden
    results: what a run against the reference's own (unshipped) golden set would look like"""
    from libmultiviewnative_b200.synthetic import make_views
    from oracle import mvn_oracle as orc

    d = make_views((8, 10, 12), num_views=2, kernel_size=3, n_sources=4, workers=1)
    psi2 = orc.inplace_cpu_deconvolve(d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4)
    tiffstack.save_view_set(str(tmp_path), d["views"], d["kernels1"], d["kernels2"], d["weights"], psi={0: d["psi0"], 2: psi2})
    assert os.path.exists(tmp_path / "input_view_1.tif") and os.path.exists(tmp_path / "psi_2.tif")
    s = tiffstack.load_view_set(str(tmp_path), 2)
    again = orc.inplace_cpu_deconvolve(s["psi"][0], s["views"], s["kernels1"], s["kernels2"], s["weights"], 2, 0.006, 1e-4)
    np.testing.assert_array_equal(again, s["psi"][2])


def test_rejects_non_tiff(tmp_path):
    p = tmp_path / "x.tif"
    p.write_bytes(b"not a tiff at all")
    with pytest.raises(ValueError):
        tiffstack.read_stack(str(p))
