"""Sharding of independent deconvolution blocks over the GPUs of one box.

The reference has no multi-GPU path; Fiji tiles a large acquisition into blocks
and calls ``inplace_gpu_deconvolve`` once per block.
Blocks are independent units, so the multi-GPU path is pure sharding -- block b
runs on GPU ``b mod G`` -- with NO collective on the data path.  Two drivers:

* ``run_sharded``: one process per GPU (torchrun / torch.distributed), each rank
  deconvolves its shard; results are optionally gathered on rank 0 for checking.
* ``run_threads``: one process, host threads per GPU (the C library is
  re-entrant per device and ctypes releases the GIL during the call).
* ``run_pipelined``: the block pipeline both are built on -- two calls in flight per
  device, so that block b+1 uploads while block b iterates.
"""
from __future__ import annotations

import threading
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def shard(n_blocks: int, rank: int, world: int) -> List[int]:
    """Indices of the blocks rank ``rank`` of ``world`` owns (round robin)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    return list(range(rank, n_blocks, world))


def deconvolve_block(lib, block: dict, num_iterations: int, lam: float, min_value: float, device: int) -> np.ndarray:
    """One block through the reference-facing entry point; returns the new psi."""
    psi = np.array(block["psi0"], dtype=np.float32, copy=True)
    lib.inplace_gpu_deconvolve(psi, block["views"], block["kernels1"], block["kernels2"], block["weights"],
                               num_iterations, lam, min_value, device)
    return psi


def run_pipelined(lib, make_block: Callable[[int], dict], indices: Sequence[int], num_iterations: int, lam: float,
                  min_value: float, device: int, depth: int = 2, keep: bool = True) -> Dict[int, Optional[np.ndarray]]:
    """Block pipeline on ONE device: ``depth`` calls of ``inplace_gpu_deconvolve`` in flight at a time, each from its
    own host thread (the library is re-entrant per device; every call has its own stream and arena, and the
    library parks up to two arenas per device between calls).  With depth = 2 the uploads, PSF spectra and the
    download of block b+1 overlap the iteration loop of block b -- the copy engines and the SMs work at the same
    time -- which is what the reference's interleaved strategy did with two streams inside one call
    (ref: src/gpu_deconvolve_methods.cuh:153-158).  Results are identical to the sequential order: blocks are
    independent.  ``make_block(b)`` is called on the worker thread, so block generation / disk reads overlap too."""
    from concurrent.futures import ThreadPoolExecutor

    out: Dict[int, Optional[np.ndarray]] = {}
    if not getattr(lib, "reentrant", True):
        depth = 1
    local = threading.local()  # keep=False: one psi buffer per worker thread, reused (a fresh 64 MiB array per block costs
                               # ~16 ms of copying plus the page faults of first-touched memory inside the call)

    def work(b: int):
        block = make_block(b)
        if keep:
            return b, deconvolve_block(lib, block, num_iterations, lam, min_value, device)
        psi0 = np.asarray(block["psi0"], dtype=np.float32)
        buf = getattr(local, "psi", None)
        if buf is None or buf.shape != psi0.shape:
            buf = local.psi = np.empty_like(psi0)
        np.copyto(buf, psi0)
        lib.inplace_gpu_deconvolve(buf, block["views"], block["kernels1"], block["kernels2"], block["weights"],
                                   num_iterations, lam, min_value, device)
        return b, None

    with ThreadPoolExecutor(max_workers=max(1, int(depth))) as pool:
        for b, res in pool.map(work, list(indices)):
            out[b] = res
    return out


def run_shard(lib, make_block: Callable[[int], dict], n_blocks: int, rank: int, world: int, num_iterations: int,
              lam: float, min_value: float, device: int, keep: bool = True, depth: int = 2) -> Dict[int, Optional[np.ndarray]]:
    return run_pipelined(lib, make_block, shard(n_blocks, rank, world), num_iterations, lam, min_value, device,
                         depth=depth, keep=keep)


def run_sharded(lib, make_block: Callable[[int], dict], n_blocks: int, num_iterations: int, lam: float,
                min_value: float, device: Optional[int] = None, gather: bool = False, depth: int = 2):
    """Every rank of the initialised torch.distributed group processes its shard.
    No data-path collective; ``gather=True`` collects the results on rank 0 (testing)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    dev = rank if device is None else device
    mine = run_shard(lib, make_block, n_blocks, rank, world, num_iterations, lam, min_value, dev, keep=gather, depth=depth)
    if not gather or world == 1:
        return mine
    gathered: List[Optional[dict]] = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    merged: Dict[int, np.ndarray] = {}
    for part in gathered:
        merged.update(part)
    return merged


def run_threads(lib, blocks: Sequence[dict], devices: Sequence[int], num_iterations: int, lam: float,
                min_value: float, depth: int = 2) -> List[np.ndarray]:
    """Single process: block b on devices[b mod G], ``depth`` host threads (calls in flight) per device."""
    results: List[Optional[np.ndarray]] = [None] * len(blocks)
    errors: List[BaseException] = []

    def worker(slot: int):
        try:
            mine = run_pipelined(lib, lambda b: blocks[b], shard(len(blocks), slot, len(devices)), num_iterations, lam,
                                 min_value, devices[slot], depth=depth)
            for b, res in mine.items():
                results[b] = res
        except BaseException as exc:  # surfaced to the caller below
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results  # type: ignore[return-value]
