#!/usr/bin/env python
"""Debug / parity of the TMA-fed strided passes: one convolution per pass mask (LMVN_TMA bit 0 = y forward, 1 = y
inverse, 2 = merged z) in a subprocess each, compared bitwise with LMVN_TMA=0."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2 and sys.argv[1] == "child":
    from libmultiviewnative_b200 import load
    from libmultiviewnative_b200.synthetic import gaussian_psf
    dims = tuple(int(x) for x in sys.argv[2].split(","))
    rng = np.random.default_rng(5)
    img = (rng.random(dims, dtype=np.float32) + 0.5).astype(np.float32)
    k = gaussian_psf(21, (3.0, 2.0, 1.5))
    lib = load()
    out = img.copy()
    lib.inplace_gpu_convolution(out, k, 0)
    np.save(sys.argv[3], out)
    sys.exit(0)

dims = sys.argv[1] if len(sys.argv) > 1 else "512,512,256"
ref = None
blocking = os.environ.get("TMA_DEBUG_BLOCKING", "0")
for mask in (0, 1, 1, 2, 2, 4, 4, 7, 7):
    env = dict(os.environ, LMVN_TMA=str(mask), CUDA_LAUNCH_BLOCKING=blocking)
    path = "/tmp/tma_dbg_%d.npy" % mask
    r = subprocess.run([sys.executable, __file__, "child", dims, path], env=env, capture_output=True, text=True, timeout=120)
    if r.returncode != 0:
        print("mask", mask, "FAILED:", (r.stderr.strip().splitlines() or ["?"])[-1])
        continue
    out = np.load(path)
    if ref is None:
        ref = out
    print("mask", mask, "ok, max abs diff vs mask 0:", float(np.max(np.abs(out - ref))), "finite:", bool(np.isfinite(out).all()))
