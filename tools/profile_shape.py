#!/usr/bin/env python
"""Per-launch profile (lmvn_plan_profile) of one (view, iteration) at a given shape: name, ms, algorithmic GB/s."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.synthetic import gaussian_psf  # noqa: E402

dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1024,1024,1024").split(","))
lib = load()
rng = np.random.default_rng(3)
img = (rng.random(dims, dtype=np.float32) + 1.0).astype(np.float32)
w = np.full(dims, 1.0, np.float32)
k = gaussian_psf(31, (4.0, 1.5, 1.5))
with lib.plan(dims, 1, 0) as p:
    p.set_view(0, img, w, k, np.ascontiguousarray(k[::-1, ::-1, ::-1]))
    p.set_psi(img)
    p.iterate(2, 0.006, 1e-4)
    ms = min(p.iterate(4, 0.006, 1e-4) for _ in range(2)) / 4
    prof = p.profile(0.006, 1e-4)
    info = p.info()
nvox = float(np.prod(dims))
print(json.dumps({"dims_zyx": list(dims), "ms_per_view_iteration": ms, "Gvox_view_iter_per_s": nvox / (ms * 1e-3) / 1e9,
                  "roofline_frac_7S10C_of_6450": info.alg_bytes_per_view_iteration / (ms * 1e-3) / 1e9 / 6450.0,
                  "profile": [{"name": n, "ms": round(t, 4), "alg_GBps": round(b / (t * 1e-3) / 1e9)} for n, t, b in prof]}))
