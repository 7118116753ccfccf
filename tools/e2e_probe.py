#!/usr/bin/env python
"""Where the time of a one-shot inplace_gpu_deconvolve call goes (development tool, GPU only)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402


def main():
    dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "256,256,256").split(","))
    nv, iters = 6, 50
    lib = load()
    rng = np.random.default_rng(0)
    # PAGEABLE=1: plain numpy (malloc) buffers, what a JNA caller hands over; the library then stages them through its
    # own pinned ring (LMVN_STAGED_COPY=0: the driver's pageable path)
    pageable = os.environ.get("PAGEABLE", "0") != "0"
    pin = (lambda a: a) if pageable else (lambda a: torch.from_numpy(a).pin_memory().numpy())
    print("host buffers: %s, LMVN_STAGED_COPY=%s" % ("pageable" if pageable else "pinned", os.environ.get("LMVN_STAGED_COPY", "1")))
    img = pin((rng.random(dims, dtype=np.float32) + 1.0))
    w = pin(np.full(dims, 1.0 / nv, np.float32))
    psi = pin(img.copy())
    k = rng.random((41, 41, 41), dtype=np.float32)
    k /= k.sum()
    views, weights, ks = [img] * nv, [w] * nv, [k] * nv
    for rep in range(4):
        t = [time.perf_counter()]
        p = lib.plan(dims, nv, 0); p.synchronize(); t.append(time.perf_counter())
        for v in range(nv):
            p.set_view(v, img, w, k, k)
        p.synchronize(); t.append(time.perf_counter())
        p.set_psi(psi); p.synchronize(); t.append(time.perf_counter())
        dev = p.iterate(iters, 0.006, 1e-4); t.append(time.perf_counter())
        p.get_psi(psi); t.append(time.perf_counter())
        p.close(); t.append(time.perf_counter())
        names = ["create", "set_views", "set_psi", "iterate", "get_psi", "destroy"]
        print("rep %d: " % rep + "  ".join("%s %.1f" % (n, (b - a) * 1e3) for n, a, b in zip(names, t[:-1], t[1:])) +
              "  | device loop %.1f ms, total %.1f ms" % (dev, (t[-1] - t[0]) * 1e3), flush=True)
    np.copyto(psi, img)
    for rep in range(3):
        t0 = time.perf_counter()
        lib.inplace_gpu_deconvolve(psi, views, ks, ks, weights, iters, 0.006, 1e-4, 0)
        print("one-shot call %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)


if __name__ == "__main__":
    main()
