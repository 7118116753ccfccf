#!/usr/bin/env python
"""A/B of the L2 carry-over between passes (direction of travel of the passes, evict-first operands) on the GPU:
every variant = (library build, environment knobs read when an engine is created) runs the same inputs in ONE
process; results are compared with the first variant (they must be bit-identical: only the order in which rows and
tiles are visited and cache hints change).

    python tools/l2_probe.py [z,y,x] [views] [iterations] [variant ...]

variant = name:lib:ENV=val,ENV=val   (lib = path of a variant build, '-' = the product library)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.synthetic import make_views_fast  # noqa: E402

KNOBS = ("LMVN_LINK_REVERSE", "LMVN_PREFETCH_LINK", "LMVN_PERSIST", "LMVN_OOP_Z", "LMVN_TMA", "LMVN_PREFETCH", "LMVN_X3", "LMVN_GRAPH", "LMVN_PREFETCH_KHAT_AHEAD", "LMVN_PREFETCH_Z", "LMVN_PREFETCH_YINV", "LMVN_PREFETCH_KHAT")


def run(lib, d, dims, nv, iters, env, long_iters):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)
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
        long_ms = None
        if long_iters:
            p.set_psi(d["psi0"])
            long_ms = [p.iterate(long_iters, 0.006, 1e-4) for _ in range(3)]
        prof = p.profile(0.006, 1e-4)
    return first, ms, long_ms, prof


def main():
    args = sys.argv[1:]
    dims = tuple(int(x) for x in (args[0] if len(args) > 0 else "512,512,256").split(","))
    nv = int(args[1]) if len(args) > 1 else 6
    iters = int(args[2]) if len(args) > 2 else 10
    long_iters = int(os.environ.get("L2_PROBE_LONG", "0"))
    variants = args[3:] or ["ascending:-:LMVN_LINK_REVERSE=0", "reversed:-:LMVN_LINK_REVERSE=1"]
    d = make_views_fast(dims, nv, 41 if min(dims) >= 128 else 15, 20240607)
    libs = {}
    ref = None
    nvox = float(np.prod(dims))
    for spec in variants:
        name, path, envs = (spec.split(":") + ["", ""])[:3]
        env = dict(kv.split("=") for kv in envs.split(",") if kv)
        key = path or "-"
        if key not in libs:
            libs[key] = load(None if key == "-" else key)
        psi, ms, long_ms, prof = run(libs[key], d, dims, nv, iters, env, long_iters)
        if ref is None:
            ref = psi
        rec = {"variant": name, "lib": key, "env": env, "dims_zyx": list(dims),
               "ms_per_view_iteration": ms / (iters * nv),
               "Gvox_view_iter_per_s": nvox * nv * iters / (ms * 1e-3) / 1e9,
               "max_abs_diff_vs_first_variant": float(np.max(np.abs(psi - ref))),
               "profile_ms": {}}
        if long_ms:
            rec["sustained_Gvox_view_iter_per_s"] = [nvox * nv * long_iters / (t * 1e-3) / 1e9 for t in long_ms]
        for n, t, b in prof:
            rec["profile_ms"].setdefault(n, []).append(round(t, 4))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
