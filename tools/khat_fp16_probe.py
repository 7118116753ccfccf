#!/usr/bin/env python
"""Opt-in half-precision PSF spectra (LMVN_KHAT_FP16=1) on BASELINE config 3: loop time and error against float32 spectra."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.synthetic import make_views_fast  # noqa: E402

dims, nv, iters = (512, 512, 256), 6, 10
lib = load()
d = make_views_fast(dims, nv, 41, 20240607)
out, psis = {}, {}
for half in (0, 1):
    os.environ["LMVN_KHAT_FP16"] = str(half)
    lib.release_cached_memory()
    with lib.plan(dims, nv, 0) as p:
        for v in range(nv):
            p.set_view(v, d["views"][v], d["weights"][v], d["kernels1"][v], d["kernels2"][v])
        p.set_psi(d["psi0"])
        p.iterate(iters, 0.006, 1e-4)
        psis[half] = p.get_psi()
        p.set_psi(d["psi0"])
        ms = min(p.iterate(iters, 0.006, 1e-4) for _ in range(3))
        prof = p.profile(0.006, 1e-4)
    out["fp16" if half else "f32"] = {"ms_per_view_iteration": ms / (iters * nv),
                                      "Gvox_view_iter_per_s": float(np.prod(dims)) * nv * iters / (ms * 1e-3) / 1e9,
                                      "z_pass_ms": [t for n, t, b in prof if n == "fast_z_mul"]}
a, b = psis[1].astype(np.float64), psis[0].astype(np.float64)
out["fp16_vs_f32_after_%d_iterations" % iters] = {"max_rel": float(np.max(np.abs(a - b) / np.abs(b))),
                                                  "rel_l2": float(np.linalg.norm(a - b) / np.linalg.norm(b))}
print(json.dumps(out))
