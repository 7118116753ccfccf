#!/usr/bin/env python
"""Config-4 block pipeline on one GPU: throughput of 256^3 blocks through inplace_gpu_deconvolve for 1, 2 and 3 calls in
flight, pageable and pinned host buffers, plus the device-resident loop alone (what a block costs without any transfer)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.blocks import run_pipelined  # noqa: E402
from libmultiviewnative_b200.synthetic import make_views_fast  # noqa: E402

dims, views, iters, n = (256, 256, 256), 6, 50, 12
lib = load()
blocks = [make_views_fast(dims, views, 41, 20240607 + 17 * i) for i in range(2)]
units = float(np.prod(dims)) * views * iters
out = {}
with lib.plan(dims, views, 0) as p:
    d = blocks[0]
    for v in range(views):
        p.set_view(v, d["views"][v], d["weights"][v], d["kernels1"][v], d["kernels2"][v])
    p.set_psi(d["psi0"])
    p.iterate(iters, 0.006, 1e-4)
    ms = min(p.iterate(iters, 0.006, 1e-4) for _ in range(3))
out["device_resident_loop"] = {"ms_per_block": ms, "Gvox": units / (ms * 1e-3) / 1e9}
pinned = []
keep = []
for b in blocks:
    q = dict(b)
    for key in ("views", "weights"):
        q[key] = []
        for a in b[key]:
            t = torch.from_numpy(a).pin_memory()
            keep.append(t)
            q[key].append(t.numpy())
    pinned.append(q)
for name, src in (("pageable", blocks), ("pinned", pinned)):
    for depth in (1, 2, 3):
        run_pipelined(lib, lambda b: src[b % 2], [0, 1, 2], iters, 0.006, 1e-4, 0, depth=depth, keep=False)  # warm-up
        t0 = time.perf_counter()
        run_pipelined(lib, lambda b: src[b % 2], list(range(n)), iters, 0.006, 1e-4, 0, depth=depth, keep=False)
        dt = time.perf_counter() - t0
        out["%s_depth_%d" % (name, depth)] = {"ms_per_block": dt / n * 1e3, "Gvox": units * n / dt / 1e9}
print(json.dumps(out, indent=1))
