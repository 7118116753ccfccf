"""Float32 multi-page TIFF stacks and the reference's fixture layout.

The reference's tests read their inputs and golden results from multi-page TIFF files -- one directory
(page) per z plane, 32-bit IEEE float, one sample per pixel, uncompressed, written scanline by scanline
(ref: tests/tiff_utils.h:90-160) -- named ``input_view_i.tif``, ``kernel1_view_i.tif``,
``kernel2_view_i.tif``, ``weights_view_i.tif`` and ``psi_i.tif`` (ref: tests/tiff_fixtures.hpp:18-27,
260-286).  Those files are not in the reference repository (ref: CMakeLists.txt:25); this module reads and
writes the same format with numpy only, so that real SPIM data or the golden set, once obtained, run
through the C ABI unchanged.  Classic TIFF (not BigTIFF), both byte orders, stripped, uncompressed;
8/16/32-bit integer and 32/64-bit float samples are converted to float32 on load.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List

import numpy as np

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 8: ("h", 2), 9: ("i", 4),
          11: ("f", 4), 12: ("d", 8), 16: ("Q", 8)}


def write_stack(path: str, stack: np.ndarray) -> None:
    """{z, y, x} float32 -> multi-page TIFF with the tags the reference writes (ref: tests/tiff_utils.h:127-141)."""
    a = np.ascontiguousarray(stack, dtype="<f4")
    if a.ndim != 3:
        raise ValueError("stack must be 3-D {z, y, x}")
    nz, h, w = a.shape
    page_bytes = h * w * 4
    tags_per_page = 13
    ifd_bytes = 2 + 12 * tags_per_page + 4
    with open(path, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, 8))
        pos = 8
        for z in range(nz):
            data_off = pos + ifd_bytes
            nxt = data_off + page_bytes if z + 1 < nz else 0
            entries = [
                (254, 4, 1, 2),             # NewSubfileType: page of a multi-page image
                (256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 32),
                (259, 3, 1, 1),             # no compression
                (262, 3, 1, 1),             # min is black
                (273, 4, 1, data_off), (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, page_bytes),
                (284, 3, 1, 1),             # contiguous
                (297, 3, 2, z | (nz << 16) if nz < 65536 else 0),  # page number z of nz (two shorts)
                (339, 3, 1, 3),             # IEEE float
            ]
            f.write(struct.pack("<H", len(entries)))
            for tag, typ, cnt, val in entries:
                if typ == 3 and cnt == 1:
                    f.write(struct.pack("<HHIHH", tag, typ, cnt, val, 0))
                else:
                    f.write(struct.pack("<HHII", tag, typ, cnt, val))
            f.write(struct.pack("<I", nxt))
            f.write(a[z].tobytes())
            pos = data_off + page_bytes


def _read_values(buf: bytes, bo: str, typ: int, cnt: int, raw: bytes) -> List[int]:
    fmt, size = _TYPES[typ]
    total = size * cnt
    if total <= 4:
        data = raw[:total]
    else:
        off = struct.unpack(bo + "I", raw)[0]
        data = buf[off:off + total]
    if typ == 5:
        vals = struct.unpack(bo + "I" * (2 * cnt), data)
        return [vals[2 * i] // max(1, vals[2 * i + 1]) for i in range(cnt)]
    return list(struct.unpack(bo + fmt * cnt, data))


def read_stack(path: str) -> np.ndarray:
    """Multi-page TIFF -> {z, y, x} float32 (pages in file order, like ref: tests/tiff_utils.h:90-118)."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:2] == b"II":
        bo = "<"
    elif buf[:2] == b"MM":
        bo = ">"
    else:
        raise ValueError(f"{path}: not a TIFF file")
    magic, off = struct.unpack(bo + "HI", buf[2:8])
    if magic != 42:
        raise ValueError(f"{path}: only classic TIFF is supported (magic {magic})")
    pages = []
    while off:
        n = struct.unpack(bo + "H", buf[off:off + 2])[0]
        tags: Dict[int, List[int]] = {}
        for i in range(n):
            e = buf[off + 2 + 12 * i: off + 14 + 12 * i]
            tag, typ, cnt = struct.unpack(bo + "HHI", e[:8])
            if typ in _TYPES:
                tags[tag] = _read_values(buf, bo, typ, cnt, e[8:12])
        off = struct.unpack(bo + "I", buf[off + 2 + 12 * n: off + 6 + 12 * n])[0]
        w, h = tags[256][0], tags[257][0]
        bits = tags.get(258, [1])[0]
        if tags.get(259, [1])[0] != 1:
            raise ValueError(f"{path}: compressed TIFF pages are not supported")
        if tags.get(277, [1])[0] != 1:
            raise ValueError(f"{path}: one sample per pixel expected")
        fmt = tags.get(339, [1])[0]
        kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
        if kind is None or bits not in (8, 16, 32, 64):
            raise ValueError(f"{path}: unsupported sample format {fmt} / {bits} bits")
        dt = np.dtype(f"{bo}{kind}{bits // 8}")
        offs, counts = tags[273], tags.get(279)
        rows_per_strip = tags.get(278, [h])[0]
        page = np.empty((h, w), dtype=np.float32)
        row = 0
        for i, so in enumerate(offs):
            rows = min(rows_per_strip, h - row)
            nbytes = rows * w * dt.itemsize if counts is None else min(counts[i], rows * w * dt.itemsize)
            page[row:row + rows] = np.frombuffer(buf, dtype=dt, count=nbytes // dt.itemsize, offset=so).reshape(rows, w)
            row += rows
        pages.append(page)
    if not pages:
        raise ValueError(f"{path}: no pages")
    shape = pages[0].shape
    if any(p.shape != shape for p in pages):
        raise ValueError(f"{path}: pages of different size")
    return np.stack(pages, axis=0)


# ---- the reference's fixture layout ---------------------------------------------------------------
def view_paths(directory: str, view: int) -> Dict[str, str]:
    return {k: os.path.join(directory, f"{k}_view_{view}.tif") for k in ("input", "kernel1", "kernel2", "weights")}


def save_view_set(directory: str, views, kernels1, kernels2, weights, psi: Dict[int, np.ndarray] = None) -> None:
    os.makedirs(directory, exist_ok=True)
    for v, (im, k1, k2, w) in enumerate(zip(views, kernels1, kernels2, weights)):
        p = view_paths(directory, v)
        write_stack(p["input"], im)
        write_stack(p["kernel1"], k1)
        write_stack(p["kernel2"], k2)
        write_stack(p["weights"], w)
    for it, a in (psi or {}).items():
        write_stack(os.path.join(directory, f"psi_{it}.tif"), a)


def load_view_set(directory: str, num_views: int) -> dict:
    """dict(views, kernels1, kernels2, weights[, psi: {iteration: stack}]) as the rest of this package uses it."""
    out = dict(views=[], kernels1=[], kernels2=[], weights=[], psi={})
    for v in range(num_views):
        p = view_paths(directory, v)
        out["views"].append(read_stack(p["input"]))
        out["kernels1"].append(read_stack(p["kernel1"]))
        out["kernels2"].append(read_stack(p["kernel2"]))
        out["weights"].append(read_stack(p["weights"]))
    for name in sorted(os.listdir(directory)):
        if name.startswith("psi_") and name.endswith(".tif"):
            out["psi"][int(name[4:-4])] = read_stack(os.path.join(directory, name))
    return out
