"""Runs the emulator parity suites against the AddressSanitizer build of the host-emulation library (red zones between
the arena sub-buffers, exactly sized shared memory): every parity case doubles as a bounds check of the kernels.

    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/run_emu_asan.py [pytest args]

About 5 minutes; the CPU suite itself runs the shorter tests/test_abi_walk.py.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)

from libmultiviewnative_b200 import _build  # noqa: E402

if "libasan" not in os.environ.get("LD_PRELOAD", ""):
    sys.exit("preload libasan (see the docstring): the interpreter itself is not an AddressSanitizer build")
_asan_path = _build.build_emu(asan=True)
_build.build_emu = lambda force=False, asan=False: _asan_path  # the test fixtures load whatever build_emu() returns

import pytest  # noqa: E402

sys.exit(pytest.main(["-x", "-q", "tests/test_emu_parity.py", "tests/test_slabs_emu.py"] + sys.argv[1:]))
