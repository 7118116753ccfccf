#!/usr/bin/env python
"""Short ncu target: one view of BASELINE config 3 (512 x 512 x 256, 41^3 PSF), two iterations on the default
schedule (set LMVN_X3=0 for the five-pass schedule)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.synthetic import gaussian_psf  # noqa: E402

dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "512,512,256").split(","))
lib = load()
rng = np.random.default_rng(3)
img = (rng.random(dims, dtype=np.float32) + 1.0).astype(np.float32)
w = np.full(dims, 1.0, np.float32)
k = gaussian_psf(41, (4.0, 1.5, 1.5))
os.environ.setdefault("LMVN_GRAPH", "0")
with lib.plan(dims, 1, 0) as p:
    p.set_view(0, img, w, k, np.ascontiguousarray(k[::-1, ::-1, ::-1]))
    p.set_psi(img)
    ms = p.iterate(2, 0.006, 1e-4)
    print("ok", ms)
