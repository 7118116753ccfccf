#!/usr/bin/env python
"""Static SASS instruction mix of selected kernels in the built library (CPU-only development aid):
    python tools/sass_mix.py 'k_stridedILi512ELi2' 'k_rows_inv2ILi128ELi1'
The fast-path kernels are fully unrolled straight-line code, so static counts track the dynamic mix."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("LMVN_LIBRARY", os.path.join(ROOT, "libmultiviewnative_b200", "lib", "libmultiviewnative.so"))
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
mix = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        mix[cur][m.group(2).split(".")[0]] += 1
for pat in sys.argv[1:] or [""]:
    for name, c in mix.items():
        if pat in name:
            tot = sum(c.values())
            print("%s  total %d" % (name[:90], tot))
            print("   " + "  ".join("%s %d" % (k, v) for k, v in c.most_common(18)))
