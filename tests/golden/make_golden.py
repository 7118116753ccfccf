"""Re-creates the reference's data-independent known-answer cases and stores them
as small fixtures next to this script.  Run from the repo root:

    python tests/golden/make_golden.py

Nothing here needs the reference checkout at run time -- each case is a
restatement of a fixture the reference's tests *construct in code*
(citations relative to the reference):

* ``conv_fixture_8.npz``  -- convolutionFixture3D<3,8>, tests/test_fixtures.hpp:21-305:
  image 0..511 (8^3), zero-padded 10^3 copy, kernels identity / horizontal /
  vertical / depth ramps (1,2,3) / all-ones, asymmetric (4,3,2) cross / one /
  identity kernels, and the expected results by DIRECT convolution
  (tests/test_algorithms.hpp:9-58) cropped back to 8^3.
* ``pointwise_cases.npz`` -- tests/test_gpu_kernels_impl.cu:24-486: divide
  (10/5, 1/5), final values (psi=5, integral=42, w=0.1, min=1e-4 -> 25.5),
  regularised final values (lambda=0.006 -> 19.10277..), random integrals in
  U(-0.1, 1) with the expected values computed in float64 from the formulas of
  inc/cpu_kernels.h:28-90.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.mvn_oracle import direct_convolve  # noqa: E402

F32 = np.float32


def conv_fixture(image_size=8, ksize=3):
    half = ksize // 2
    image = np.arange(image_size ** 3, dtype=F32).reshape((image_size,) * 3)
    one = np.zeros_like(image)
    one[(image_size // 2,) * 3] = 1
    pad_axis = image_size + 2 * half
    padded_image = np.zeros((pad_axis,) * 3, dtype=F32)
    padded_image[half:half + image_size, half:half + image_size, half:half + image_size] = image
    padded_one = np.zeros_like(padded_image)
    padded_one[(pad_axis // 2,) * 3] = 1

    kernels = {}
    kernels["trivial"] = np.zeros((ksize,) * 3, dtype=F32)
    ident = np.zeros((ksize,) * 3, dtype=F32)
    ident.ravel()[ksize ** 3 // 2] = 1
    kernels["identity"] = ident
    hor = np.zeros((ksize,) * 3, dtype=F32)
    ver = np.zeros((ksize,) * 3, dtype=F32)
    dep = np.zeros((ksize,) * 3, dtype=F32)
    for i in range(ksize):
        hor[half, half, i] = i + 1
        ver[half, i, half] = i + 1
        dep[i, half, half] = i + 1
    kernels["horizontal"], kernels["vertical"], kernels["depth"] = hor, ver, dep
    kernels["all1"] = np.ones((ksize,) * 3, dtype=F32)

    ashape = (ksize + 1, ksize, ksize - 1)
    cross = np.zeros(ashape, dtype=F32)
    aone = np.zeros(ashape, dtype=F32)
    aid = np.zeros(ashape, dtype=F32)
    aid[ashape[0] // 2, ashape[1] // 2, ashape[2] // 2] = 1
    for z in range(ashape[0]):
        for y in range(ashape[1]):
            for x in range(ashape[2]):
                if z == ashape[0] // 2 and y == ashape[1] // 2:
                    cross[z, y, x] = x + 1
                    aone[z, y, x] = 1
                if x == ashape[2] // 2 and y == ashape[1] // 2:
                    cross[z, y, x] = z + 101
                    aone[z, y, x] = 1
                if x == ashape[2] // 2 and z == ashape[0] // 2:
                    cross[z, y, x] = y + 11
                    aone[z, y, x] = 1
    kernels["asymm_cross"], kernels["asymm_one"], kernels["asymm_identity"] = cross, aone, aid

    out = dict(image=image, one=one, padded_image=padded_image, padded_one=padded_one)
    for k, v in kernels.items():
        out["kernel_" + k] = v
    sl = slice(half, half + image_size)
    for name in ("horizontal", "vertical", "depth", "all1"):
        res = direct_convolve(padded_image, kernels[name], (half,) * 3)
        out["image_folded_by_" + name] = np.ascontiguousarray(res[sl, sl, sl])

    # asymmetric: image padded by k//2 per axis (tests/test_fixtures.hpp:236-257)
    aoff = [s // 2 for s in ashape]
    adims = [image_size + 2 * o for o in aoff]
    apad_one = np.zeros(adims, dtype=F32)
    apad_one[adims[0] // 2, adims[1] // 2, adims[2] // 2] = 1
    apad_img = np.zeros(adims, dtype=F32)
    asl = tuple(slice(o, o + image_size) for o in aoff)
    apad_img[asl] = image
    out["asymm_padded_one"] = apad_one
    out["asymm_padded_image"] = apad_img
    for name in ("asymm_cross", "asymm_one", "asymm_identity"):
        res = direct_convolve(apad_one, kernels[name], aoff)
        out["one_folded_by_" + name] = np.ascontiguousarray(res[asl])
    return out


def pointwise_cases(seed=7):
    rng = np.random.default_rng(seed)
    n = 4096
    out = {}
    out["divide_in"] = np.array([10.0, 1.0], dtype=F32)
    out["divide_out"] = np.array([5.0, 5.0], dtype=F32)
    out["divide_expected"] = np.array([2.0, F32(1.0) / F32(5.0)], dtype=F32)
    # constants: psi=5, integral=42, w=0.1, min=1e-4
    out["const_expected_plain"] = np.array([F32(0.1) * (F32(210.0) - F32(5.0)) + F32(5.0)], dtype=F32)
    lam = 0.006
    v = (np.sqrt(1.0 + 2.0 * lam * 210.0) - 1.0) * np.float64(F32(1.0 / lam))
    out["const_expected_reg"] = np.array([F32(0.1) * (F32(v) - F32(5.0)) + F32(5.0)], dtype=F32)
    psi = rng.uniform(0.5, 50.0, n).astype(F32)
    integral = rng.uniform(-0.1, 1.0, n).astype(F32)
    weight = rng.uniform(0.0, 1.0, n).astype(F32)
    out["rand_psi"], out["rand_integral"], out["rand_weight"] = psi, integral, weight
    mn = F32(1e-4)
    val = (psi * integral).astype(F32)
    plain = np.where(val > 0, np.maximum(val, mn), mn).astype(F32)
    out["rand_expected_plain"] = ((weight * (plain - psi)).astype(F32) + psi).astype(F32)
    reg = (np.float64(F32(1.0 / lam)) * (np.sqrt(1.0 + 2.0 * lam * np.where(val > 0, val, 0).astype(np.float64)) - 1.0)).astype(F32)
    reg = np.where(val > 0, np.maximum(reg, mn), mn).astype(F32)
    out["rand_expected_reg"] = ((weight * (reg - psi)).astype(F32) + psi).astype(F32)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "conv_fixture_8.npz"), **conv_fixture())
    np.savez_compressed(os.path.join(HERE, "pointwise_cases.npz"), **pointwise_cases())
    print("wrote", os.listdir(HERE))
