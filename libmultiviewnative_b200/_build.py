"""In-tree builds.

* ``build_cuda()``  -> libmultiviewnative_b200/lib/libmultiviewnative.so, nvcc, sm_100a
  only.  This is the product: the C-ABI drop-in for the reference's
  libmultiviewnative.so.
* ``build_emu()``   -> tests/emu/_lmvn_emu.so, g++ with -DLMVN_EMU against
  tests/emu/cuda_emu.h.  Test infrastructure only (kernel index math without a
  GPU); never loaded by the package.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmultiviewnative.so")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_PATH = os.path.join(EMU_DIR, "_lmvn_emu.so")

CUDA_SOURCES = ["api.cu", "engine.cu", "fft_fused.cu", "dist.cu"]
HOST_SOURCES = ["cpu_path.cpp"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    out = []
    for d, _, files in os.walk(CSRC):
        out += [os.path.join(d, f) for f in files if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    out += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return out


def _stale(target: str, extra=()) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in list(_sources()) + list(extra))


def build_cuda(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """`out` / `defines` build an experimental variant next to the product library
    (A/B timing of compile-time knobs); the default call builds the product."""
    target = out or LIB_PATH
    if not force and not _stale(target):
        return target
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_ARCH, "--threads", "0", "-lineinfo", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC,-fopenmp,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    srcs = list(CUDA_SOURCES)
    cmd += ["-DLMVN_HAVE_FUSED"]
    cmd += ["-D" + d for d in defines]
    cmd += [os.path.join(CSRC, s) for s in srcs + HOST_SOURCES]
    cmd += ["-o", target, "-lgomp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc build of libmultiviewnative.so failed")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return target


EMU_ASAN_PATH = os.path.join(EMU_DIR, "_lmvn_emu_asan.so")
EMU_UBSAN_PATH = os.path.join(EMU_DIR, "_lmvn_emu_ubsan.so")


def build_emu(force: bool = False, asan: bool = False, ubsan: bool = False) -> str:
    """asan=True: the same library under AddressSanitizer with red zones between the sub-buffers of the device
    arenas (-DLMVN_ARENA_REDZONE) -- the bounds check of the kernels' index math (tests/test_abi_walk.py).
    ubsan=True: under UndefinedBehaviorSanitizer, trapping -- misaligned float2 / float4 accesses (they fault on the
    device), signed overflow in index arithmetic, out-of-range shifts."""
    emu_srcs = [os.path.join(EMU_DIR, "cuda_emu.cpp"), os.path.join(EMU_DIR, "cuda_emu.h")]
    target = EMU_ASAN_PATH if asan else (EMU_UBSAN_PATH if ubsan else EMU_PATH)
    if not force and not _stale(target, emu_srcs):
        return target
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-fopenmp", "-DLMVN_EMU",
           "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function",
           "-I", EMU_DIR, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if asan:
        cmd += ["-fsanitize=address", "-fno-omit-frame-pointer", "-DLMVN_ARENA_REDZONE"]
    elif ubsan:
        cmd += ["-fsanitize=undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer"]
    srcs = list(CUDA_SOURCES)
    cmd += ["-DLMVN_HAVE_FUSED"]
    for s in srcs:
        cmd += ["-x", "c++", os.path.join(CSRC, s)]
    cmd += ["-x", "c++", os.path.join(CSRC, "cpu_path.cpp"), os.path.join(EMU_DIR, "cuda_emu.cpp")]
    cmd += ["-o", target]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ build of the emulated test library failed")
    return target


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "cuda"
    if which in ("cuda", "all"):
        print(build_cuda(force=True, verbose="-v" in sys.argv))
    if which in ("emu", "all"):
        print(build_emu(force=True))
