"""Block tiler with halo: the step BEFORE the hot path in the Fiji pipeline.

An acquisition that is too large for one call (or that should be spread over several GPUs) is cut
into overlapping blocks; every block is an independent ``inplace_gpu_deconvolve`` call (BASELINE
config 4, sharded over GPUs without any collective) and only its interior is kept.  The halo is
what the reference's fixtures add around a stack before deconvolving it -- ``num_kernel_widths *
(kernel_extent / 2)`` voxels per side (ref: tests/tiff_fixtures.hpp:225-258) -- so that the circular
wrap-around of the FFT convolution (ref: inc/cpu_convolve.h:24, ``no_padd``) lands in voxels that are
thrown away.  One (view, iteration) step looks ``2 * (k // 2)`` voxels far (two convolutions) and the
views update psi one after the other, hence ``halo >= 2 * iterations * num_views * (k // 2)`` reproduces
the untiled result exactly; in practice one or two kernel widths are used and the error decays with
the PSF tails.

Padding at the volume border: ``reflect`` for psi and the views, zeros for the weights (a zero
weight freezes psi there, like the reference's zero-padded weight stacks), or ``wrap`` for everything
(makes the tiled result equal to the untiled circular one; used by the tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class Block:
    index: int
    start: Tuple[int, int, int]      # first voxel of the block (halo included) in volume coordinates, may be < 0
    shape: Tuple[int, int, int]      # block extents, halo included
    keep_lo: Tuple[int, int, int]    # interior [keep_lo, keep_hi) in block coordinates
    keep_hi: Tuple[int, int, int]

    @property
    def dst(self):
        return tuple(slice(s + lo, s + hi) for s, lo, hi in zip(self.start, self.keep_lo, self.keep_hi))

    @property
    def src(self):
        return tuple(slice(lo, hi) for lo, hi in zip(self.keep_lo, self.keep_hi))


def halo_for(kernel_shapes: Sequence[Sequence[int]], num_kernel_widths: int = 1) -> Tuple[int, int, int]:
    """ref: tests/tiff_fixtures.hpp:241 -- offset = num_kernel_widths * (extent / 2), per axis, largest PSF."""
    return tuple(int(num_kernel_widths) * max(int(k[a]) // 2 for k in kernel_shapes) for a in range(3))


def plan_blocks(volume_shape: Sequence[int], block_shape: Sequence[int], halo: Sequence[int]) -> List[Block]:
    """Cover the volume with blocks of exactly ``block_shape`` (so that one plan / one FFT shape serves all of
    them); consecutive interiors abut, the last block of an axis is shifted inwards instead of shrunk."""
    vol = tuple(int(v) for v in volume_shape)
    blk = tuple(int(b) for b in block_shape)
    hal = tuple(int(h) for h in halo)
    per_axis = []
    for a in range(3):
        inner = blk[a] - 2 * hal[a]
        if inner <= 0:
            raise ValueError(f"axis {a}: block extent {blk[a]} leaves no interior with a halo of {hal[a]}")
        spans = []
        pos = 0
        while pos < vol[a]:
            lo = pos                       # first kept voxel
            hi = min(pos + inner, vol[a])  # one past the last kept voxel
            start = lo - hal[a]
            if hi - lo < inner and vol[a] >= inner:
                start = vol[a] - inner - hal[a]  # shift the last block inwards, keep only the new voxels
            spans.append((start, lo - start, hi - start))
            pos = hi
        per_axis.append(spans)
    out = []
    for sz in per_axis[0]:
        for sy in per_axis[1]:
            for sx in per_axis[2]:
                out.append(Block(len(out), (sz[0], sy[0], sx[0]), blk, (sz[1], sy[1], sx[1]), (sz[2], sy[2], sx[2])))
    return out


def extract(volume: np.ndarray, block: Block, mode: str = "reflect") -> np.ndarray:
    """The block's voxels, padded where it sticks out of the volume."""
    idx = []
    for a in range(3):
        i = np.arange(block.start[a], block.start[a] + block.shape[a])
        n = volume.shape[a]
        if mode == "wrap":
            i = np.mod(i, n)
        elif mode == "reflect":
            period = 2 * n - 2 if n > 1 else 1
            i = np.mod(i, period)
            i = np.where(i >= n, period - i, i)
        elif mode == "zero":
            pass
        else:
            raise ValueError(mode)
        idx.append(i)
    if mode == "zero":
        out = np.zeros(block.shape, dtype=volume.dtype)
        src, dst = [], []
        for a in range(3):
            lo, hi = max(block.start[a], 0), min(block.start[a] + block.shape[a], volume.shape[a])
            src.append(slice(lo, hi))
            dst.append(slice(lo - block.start[a], hi - block.start[a]))
        out[tuple(dst)] = volume[tuple(src)]
        return out
    return np.ascontiguousarray(volume[np.ix_(*idx)])


def make_block_inputs(block: Block, psi, views, kernels1, kernels2, weights, pad_mode: str = "reflect") -> dict:
    wmode = "wrap" if pad_mode == "wrap" else "zero"
    return {
        "psi0": extract(psi, block, pad_mode),
        "views": [extract(v, block, pad_mode) for v in views],
        "weights": [extract(w, block, wmode) for w in weights],
        "kernels1": list(kernels1), "kernels2": list(kernels2),
    }


def deconvolve_tiled(deconvolve_block: Callable[[dict], np.ndarray], psi: np.ndarray, views, kernels1, kernels2, weights,
                     block_shape: Sequence[int], halo: Optional[Sequence[int]] = None, num_kernel_widths: int = 1,
                     pad_mode: str = "reflect", blocks: Optional[Sequence[int]] = None) -> np.ndarray:
    """Tiles, runs ``deconvolve_block(inputs) -> new psi`` on every block (or on the indices in ``blocks``:
    a rank's shard) and stitches the interiors into a copy of psi.

    ``deconvolve_block`` is e.g. ``lambda b: blocks.deconvolve_block(lib, b, iterations, lam, min_value, device)``.
    """
    if halo is None:
        halo = halo_for([np.shape(k) for k in list(kernels1) + list(kernels2)], num_kernel_widths)
    plan = plan_blocks(psi.shape, block_shape, halo)
    out = np.array(psi, dtype=np.float32, copy=True)
    for b in plan:
        if blocks is not None and b.index not in blocks:
            continue
        res = deconvolve_block(make_block_inputs(b, psi, views, kernels1, kernels2, weights, pad_mode))
        out[b.dst] = res[b.src]
    return out
