#!/usr/bin/env python
"""A/B of two library builds at a shape (per-launch profile + loop time + bitwise comparison of psi after 2 iterations).

    python tools/wide_probe.py z,y,x libA.so libB.so
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.synthetic import gaussian_psf  # noqa: E402

dims = tuple(int(x) for x in sys.argv[1].split(","))
rng = np.random.default_rng(3)
img = (rng.random(dims, dtype=np.float32) + 1.0).astype(np.float32)
w = (rng.random(dims, dtype=np.float32) * 0.5 + 0.5).astype(np.float32)
k = gaussian_psf(31, (4.0, 1.5, 1.5))
ref = None
for path in sys.argv[2:]:
    lib = load(None if path == "-" else path)
    lib.release_cached_memory()
    with lib.plan(dims, 1, 0) as p:
        p.set_view(0, img, w, k, np.ascontiguousarray(k[::-1, ::-1, ::-1]))
        p.set_psi(img)
        p.iterate(2, 0.006, 1e-4)
        psi = p.get_psi()
        p.set_psi(img)
        ms = min(p.iterate(4, 0.006, 1e-4) for _ in range(3)) / 4
        prof = p.profile(0.006, 1e-4)
    lib.release_cached_memory()
    if ref is None:
        ref = psi
    print(json.dumps({"lib": path, "dims_zyx": list(dims), "ms_per_view_iteration": ms,
                      "Gvox_view_iter_per_s": float(np.prod(dims)) / (ms * 1e-3) / 1e9,
                      "max_abs_diff_vs_first": float(np.max(np.abs(psi - ref))),
                      "profile": [{"name": n, "ms": round(t, 4), "alg_GBps": round(b / (t * 1e-3) / 1e9)} for n, t, b in prof]}),
          flush=True)
