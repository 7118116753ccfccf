#!/usr/bin/env python
"""Pageable-buffer call time of a 256^3 block through inplace_gpu_deconvolve for both staging-copy implementations
(run once per LMVN_STAGING_IMPL value: the choice is made when the library first stages a copy)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.synthetic import make_views_fast  # noqa: E402

dims = (256, 256, 256)
lib = load()
d = make_views_fast(dims, 6, 41, 1)
psi = d["psi0"].copy()
ts = []
for i in range(8):
    np.copyto(psi, d["psi0"])
    t0 = time.perf_counter()
    lib.inplace_gpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], 50, 0.006, 1e-4, 0)
    ts.append((time.perf_counter() - t0) * 1e3)
print("LMVN_STAGING_IMPL=%s: calls (ms) %s -> best %.1f median %.1f" % (os.environ.get("LMVN_STAGING_IMPL", "default"),
                                                                        [round(t, 1) for t in ts[1:]], min(ts[1:]), float(np.median(ts[1:]))))
