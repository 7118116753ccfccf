#!/usr/bin/env python
"""Per-kernel timing helper (development tool, GPU only): builds a small resident
plan with random data and prints the CUDA-event time of every launch of one
(view, iteration), best of N repeats, plus algorithmic GB/s.  Select a variant
library with LMVN_LIBRARY=path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import capi  # noqa: E402


def main():
    dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "512,512,256").split(","))
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    strategy = int(os.environ.get("KB_STRATEGY", "0"))
    lib = capi.Library()
    lib.set_default_strategy(strategy)
    rng = np.random.default_rng(0)
    img = (rng.random(dims, dtype=np.float32) + 1.0)
    w = np.full(dims, 0.5, np.float32)
    k = rng.random((21, 21, 21), dtype=np.float32)
    k /= k.sum()
    with lib.plan(dims, 1, 0) as p:
        p.set_view(0, img, w, k, np.ascontiguousarray(k[::-1, ::-1, ::-1]))
        p.set_psi(img)
        info = p.info()
        p.iterate(2, 0.006, 1e-4)
        best = {}
        order = []
        for _ in range(reps):
            agg = {}
            for name, ms, nb in p.profile(0.006, 1e-4):
                a = agg.setdefault(name, [0.0, 0, 0])
                a[0] += ms; a[1] += nb; a[2] += 1
                if name not in order:
                    order.append(name)
            for name, a in agg.items():
                if name not in best or a[0] < best[name][0]:
                    best[name] = a
        total = sum(a[0] for a in best.values())
        alg = info.alg_bytes_per_view_iteration
        loop = min(p.iterate(10, 0.006, 1e-4) for _ in range(3)) / 10.0
        print("lib=%s dims=%s  LOOP (no per-launch events) %.4f ms per view-iteration -> %.1f%% of 6450 GB/s (7S+10C), %.1f Gvox/s" % (
            os.path.basename(lib.path), dims, loop, 100 * alg / (loop * 1e-3) / 1e9 / 6450,
            np.prod(dims) / (loop * 1e-3) / 1e9))
        print("lib=%s dims=%s strategy=%d  view-iteration %.4f ms  -> %.1f%% of 6450 GB/s roofline (7S+10C)" % (
            os.path.basename(lib.path), dims, info.strategy, total, 100 * alg / (total * 1e-3) / 1e9 / 6450))
        for name in order:
            a = best[name]
            print("   %-26s x%d  %8.4f ms each  %7.0f GB/s" % (name, a[2], a[0] / a[2], a[1] / a[2] / (a[0] / a[2] * 1e-3) / 1e9))


if __name__ == "__main__":
    main()
