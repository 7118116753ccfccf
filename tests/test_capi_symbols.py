"""CPU-only checks of the drop-in boundary: the nvcc-built library loads, exports
every symbol include/*.h declares, keeps the reference's struct layouts, and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import shutil

import numpy as np
import pytest

from libmultiviewnative_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs_nvcc = pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                                reason="nvcc not available")


@pytest.fixture(scope="module")
def lib():
    from libmultiviewnative_b200._build import build_cuda

    return capi.Library(build_cuda())


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    names = re.findall(r"(?:FUNCTION_PREFIX|LMVN_EXPORT)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
    return sorted(set(names))


@needs_nvcc
def test_every_declared_symbol_is_exported(lib):
    ref = _declared("multiviewnative.h")
    ext = _declared("lmvn_b200.h")
    assert sorted(ref) == sorted(capi.REFERENCE_SYMBOLS)  # the reference's 16 entry points
    assert sorted(ext) == sorted(capi.EXTENSION_SYMBOLS)
    for name in ref + ext:
        assert hasattr(lib.lib, name), name


def test_struct_layouts_match_reference():
    # ref: inc/multiviewnative.h:15-35 -- 64-byte view_data, 32-byte workspace {ptr@0,u16@8,f64@16,f32@24,i32@28}
    assert C.sizeof(capi.ViewData) == 64
    assert C.sizeof(capi.Workspace) == 32
    assert capi.Workspace.num_views_.offset == 8
    assert capi.Workspace.lambda_.offset == 16
    assert capi.Workspace.minValue_.offset == 24
    assert capi.Workspace.num_iterations_.offset == 28


@needs_nvcc
def test_header_compiles_as_c_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "lmvn_b200.h"\nint main(void){ struct workspace w; (void)w; return sizeof(struct view_data)==64 ? 0 : 1; }\n')
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        exe = tmp_path / ("t_" + cc)
        import subprocess

        lang = ["-x", "c++"] if cc == "g++" else []
        subprocess.check_call([cc, std, "-I", os.path.join(ROOT, "include"), *lang, str(src), "-o", str(exe)])
        assert subprocess.call([str(exe)]) == 0


@needs_nvcc
def test_no_cpu_fallback_without_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert "sm_100a" in lib.version()
    psi = np.ones((8, 8, 8), np.float32)
    before = psi.copy()
    with pytest.raises(capi.LmvnError):
        lib.inplace_gpu_convolution(psi, np.ones((3, 3, 3), np.float32))
    np.testing.assert_array_equal(psi, before)
    with pytest.raises(capi.LmvnError):
        lib.plan((8, 8, 8), 1)


def test_missing_library_raises(tmp_path):
    with pytest.raises(capi.LmvnError):
        capi.Library(str(tmp_path / "nope.so"))


@needs_nvcc
def test_kernels_are_sm100a(lib):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    import subprocess

    out = subprocess.run([cuobjdump, "-lelf", lib.path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


@needs_nvcc
def test_sass_carries_the_blackwell_paths(lib):
    """what the shipped library is made of, checked on its SASS (no GPU needed): packed FP32x2 arithmetic in the transforms,
    the tensor-map TMA copies + mbarrier waits of the TMA-fed strided passes (csrc/fft_tma.cuh), the bulk copies of the
    two-pass schedule (csrc/fft_x3.cuh), and no library FFT / BLAS kernels at all."""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    import subprocess

    sass = subprocess.run([cuobjdump, "-sass", lib.path], capture_output=True, text=True).stdout
    for mnemonic in ("FFMA2", "FADD2", "UTMALDG.3D", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "UBLKCP", "LDG.E.64.STRONG.GPU"):
        assert mnemonic in sass, mnemonic
    funcs = re.findall(r"Function : (\S+)", sass)
    assert any("k_strided_tma" in f for f in funcs) and any("k_rows_inv_fwd" in f for f in funcs)
    assert not any(("cufft" in f.lower()) or ("cublas" in f.lower()) for f in funcs)


def _build_c_client(tmp_path):
    import shutil
    import subprocess

    from libmultiviewnative_b200 import capi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "deconvolve_c_client")
    libdir = os.path.dirname(capi.DEFAULT_LIBRARY)
    subprocess.run([gcc, "-std=c99", "-Wall", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "deconvolve_c_client.c"), "-L", libdir, "-lmultiviewnative", "-lm",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_c_client_links_against_the_drop_in(tmp_path):
    """a plain C client written against the reference's header links with nothing but the library swapped; without a
    device it reports that and exits 0 (no CPU fallback)"""
    import subprocess

    res = subprocess.run([_build_c_client(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr


@pytest.mark.gpu
def test_c_client_runs_on_the_gpu(tmp_path):
    import subprocess

    res = subprocess.run([_build_c_client(tmp_path)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr
