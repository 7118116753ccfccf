"""tools/abi_walk.c: a C program that walks every geometry of the C ABI with self-checking inputs.

CPU: built against the host-emulation library compiled under AddressSanitizer, with poisoned red zones between
the sub-buffers of the device arenas -- the bounds check of the kernels' index math (global and shared memory).
GPU: built against the product library and run as it is (all cases, every rows / strided plan of the fast path).
"""
import os
import shutil
import subprocess

import pytest

from libmultiviewnative_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "abi_walk.c")

# a subset that keeps the CPU suite short; `python tests/test_abi_walk.py` runs all of them (minutes)
ASAN_CASES = ["fast/16x16x32", "conv/", "embedded/20x24x28", "zero/20x24x28", "plan/zero_padded/28x28x28",
              "slabs/32x32x64", "legacy"]


def _compile(out, lib_path, extra=()):
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("no gcc")
    lib_dir = os.path.dirname(lib_path)
    cmd = [cc, "-std=c99", "-O1", "-g", *extra, "-I", os.path.join(ROOT, "include"), SRC, lib_path, "-lm",
           "-Wl,-rpath," + lib_dir, "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return out


def _run_asan(tmp_dir, cases, timeout):
    lib = _build.build_emu(asan=True)
    exe = _compile(os.path.join(str(tmp_dir), "abi_walk_asan"), lib, extra=("-fsanitize=address",))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    res = subprocess.run([exe, *cases], capture_output=True, text=True, timeout=timeout, env=env)
    out = res.stdout + res.stderr
    assert "ERROR: AddressSanitizer" not in out, out[-6000:]
    assert res.returncode == 0, out[-6000:]
    assert "0 case(s) failed" in out and "MISMATCH" not in out, out[-6000:]
    return out


def test_abi_walk_under_address_sanitizer(tmp_path):
    out = _run_asan(tmp_path, ASAN_CASES, timeout=900)
    assert out.count(" OK") >= 12, out


@pytest.mark.parametrize("order", ["reverse", "shuffle"])
def test_abi_walk_with_other_thread_orders(tmp_path, order):
    """The emulated threads of a block run one after the other up to their next barrier; under another order a missing
    barrier in a shared-memory exchange reads stale or poisoned data (tools/run_emu_sanitized.py --order runs all
    emulator parity cases this way)."""
    exe = _compile(os.path.join(str(tmp_path), "abi_walk_emu"), _build.build_emu())
    env = dict(os.environ, LMVN_EMU_ORDER=order)
    res = subprocess.run([exe, "fast/32x64x64", "conv/fast", "embedded/20x24x28", "slabs/32x32x64"], capture_output=True,
                         text=True, timeout=900, env=env)
    out = res.stdout + res.stderr
    assert res.returncode == 0 and "0 case(s) failed" in out and "MISMATCH" not in out, out[-6000:]
    assert out.count(" OK") >= 5, out


@pytest.mark.gpu
def test_abi_walk_on_the_gpu(tmp_path):
    lib = _build.build_cuda()
    exe = _compile(os.path.join(str(tmp_path), "abi_walk"), lib)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-6000:]
    assert "0 case(s) failed" in out and "MISMATCH" not in out, out[-6000:]
    assert out.count(" OK") >= 30, out


if __name__ == "__main__":
    import tempfile
    print(_run_asan(tempfile.mkdtemp(), [], timeout=6 * 3600))
