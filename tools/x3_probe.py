#!/usr/bin/env python
"""A/B of the two-pass schedule (fft_x3.cuh, LMVN_X3=1) against the chained five-pass schedule (LMVN_X3=0) on the
GPU: same inputs, both results against each other, per-launch profile and loop time of each.

    python tools/x3_probe.py [z,y,x] [views] [iterations]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.synthetic import make_views_fast  # noqa: E402


def run(lib, d, dims, nv, iters, x3):
    os.environ["LMVN_X3"] = "1" if x3 else "0"
    lib.release_cached_memory()
    with lib.plan(dims, nv, 0) as p:
        for v in range(nv):
            p.set_view(v, d["views"][v], d["weights"][v], d["kernels1"][v], d["kernels2"][v])
        p.set_psi(d["psi0"])
        p.iterate(2, 0.006, 1e-4)
        first = p.get_psi()
        p.set_psi(d["psi0"])
        p.iterate(iters, 0.006, 1e-4)  # warm-up (graph capture)
        p.set_psi(d["psi0"])
        ms = min(p.iterate(iters, 0.006, 1e-4) for _ in range(3))
        prof = p.profile(0.006, 1e-4)
        info = p.info()
    return first, ms, prof, info


def main():
    dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "512,512,256").split(","))
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    lib = load()
    d = make_views_fast(dims, nv, 41 if min(dims) >= 128 else 15, 20240607)
    out = {}
    res = {}
    for x3 in (0, 1):
        psi, ms, prof, info = run(lib, d, dims, nv, iters, x3)
        res[x3] = psi
        nvox = float(np.prod(dims))
        out["x3" if x3 else "five_pass"] = {
            "ms_per_view_iteration": ms / (iters * nv), "Gvox_view_iter_per_s": nvox * nv * iters / (ms * 1e-3) / 1e9,
            "launches_per_view_iteration": int(info.launches_per_view_iteration),
            "profile": [{"name": n, "ms": t, "alg_GBps": b / (t * 1e-3) / 1e9} for n, t, b in prof]}
    a, b = res[1].astype(np.float64), res[0].astype(np.float64)
    out["x3_vs_five_pass_after_2_iterations"] = {"max_rel": float(np.max(np.abs(a - b) / np.abs(b))),
                                                 "rel_l2": float(np.linalg.norm(a - b) / np.linalg.norm(b))}
    out["dims_zyx"] = list(dims)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
