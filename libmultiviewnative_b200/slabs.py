"""One volume over several GPUs: host side of the slab-decomposed plans
(include/lmvn_b200.h, lmvn_dist_*; csrc/dist.cu).

Real space is cut into slabs of nz/G planes, the z pass works on pencils of ny/G
rows, and both all-to-all exchanges of a convolution are stores into peer memory
issued by the transform kernels themselves.  Two ways to form the group:

* ``LocalSlabGroup``  -- every rank is a handle of THIS process (all on one GPU for
  tests, or one GPU each when a single process drives the box, like Fiji's Java
  threads drive the reference).  Phases are issued rank by rank and the host
  synchronises between phases.
* ``ProcessSlabPlan`` -- one process per GPU (``torchrun``): the 64-byte CUDA IPC
  handles of the exchange regions travel through ``torch.distributed``; afterwards
  the whole loop runs on the device with flag barriers in peer memory and no host
  round trip.  An NCCL ``all_to_all_single`` variant of the same exchange is kept
  as the comparator (``iterate_nccl``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from .capi import DistInfo, Library, _dims, _f32, _fp, c_int_p


def slab_of(volume: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The rank's planes of a {z, y, x} stack (a view when the stack is contiguous)."""
    nz = volume.shape[0]
    if nz % world:
        raise ValueError("nz must be divisible by the world size")
    n = nz // world
    return np.ascontiguousarray(volume[rank * n:(rank + 1) * n])


class SlabPlan:
    """One rank's handle."""

    def __init__(self, library: Library, dims, num_views: int, rank: int, world: int, device: int = -1):
        self.L = library
        self.dims = tuple(int(d) for d in dims)
        self.rank, self.world, self.num_views = int(rank), int(world), int(num_views)
        self.handle = C.c_void_p()
        library._check(library.lib.lmvn_dist_create(C.byref(self.handle), C.cast(_dims(self.dims), c_int_p),
                                                    int(num_views), int(rank), int(world), int(device)),
                       "lmvn_dist_create")
        self.slab_shape = (self.dims[0] // self.world, self.dims[1], self.dims[2])

    def close(self):
        if self.handle:
            self.L.lib.lmvn_dist_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> DistInfo:
        info = DistInfo()
        self.L._check(self.L.lib.lmvn_dist_get_info(self.handle, C.byref(info)), "lmvn_dist_get_info")
        return info

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self.L._check(self.L.lib.lmvn_dist_export_handle(self.handle, buf), "lmvn_dist_export_handle")
        return buf.raw

    def connect_ipc(self, peer: int, handle: bytes):
        self.L._check(self.L.lib.lmvn_dist_connect_ipc(self.handle, int(peer), handle), "lmvn_dist_connect_ipc")

    def connect_local(self, peer: "SlabPlan"):
        self.L._check(self.L.lib.lmvn_dist_connect_local(self.handle, peer.rank, peer.handle), "lmvn_dist_connect_local")

    def _slab(self, a, name):
        a = _f32(a, name)
        if a.shape != self.slab_shape:
            raise ValueError(f"{name}: expected the rank's slab {self.slab_shape}, got {a.shape}")
        return a

    def set_view_slab(self, v: int, image_slab, weights_slab):
        image_slab, weights_slab = self._slab(image_slab, "image"), self._slab(weights_slab, "weights")
        self.L._check(self.L.lib.lmvn_dist_set_view_slab(self.handle, int(v), _fp(image_slab), _fp(weights_slab)),
                      "lmvn_dist_set_view_slab")

    def set_psi_slab(self, psi_slab):
        self.L._check(self.L.lib.lmvn_dist_set_psi_slab(self.handle, _fp(self._slab(psi_slab, "psi"))),
                      "lmvn_dist_set_psi_slab")

    def get_psi_slab(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.slab_shape, dtype=np.float32)
        self.L._check(self.L.lib.lmvn_dist_get_psi_slab(self.handle, _fp(self._slab(out, "psi"))),
                      "lmvn_dist_get_psi_slab")
        return out

    def psf_phase(self, v: int, which: int, phase: int, kernel=None):
        if kernel is not None:
            kernel = _f32(kernel, "kernel")
            rc = self.L.lib.lmvn_dist_psf_phase(self.handle, int(v), int(which), int(phase), _fp(kernel),
                                                C.cast(_dims(kernel.shape), c_int_p))
        else:
            rc = self.L.lib.lmvn_dist_psf_phase(self.handle, int(v), int(which), int(phase), None, None)
        self.L._check(rc, "lmvn_dist_psf_phase")

    def conv_phase(self, v: int, which: int, phase: int, lam: float, min_value: float):
        self.L._check(self.L.lib.lmvn_dist_conv_phase(self.handle, int(v), int(which), int(phase), float(lam),
                                                      float(min_value)), "lmvn_dist_conv_phase")

    def barrier(self):
        self.L._check(self.L.lib.lmvn_dist_barrier(self.handle), "lmvn_dist_barrier")

    def iterate(self, iterations: int, lam: float = 0.0, min_value: float = 1e-4) -> float:
        ms = C.c_float(0.0)
        self.L._check(self.L.lib.lmvn_dist_iterate(self.handle, int(iterations), float(lam), float(min_value),
                                                   C.byref(ms)), "lmvn_dist_iterate")
        return float(ms.value)

    def synchronize(self):
        self.L._check(self.L.lib.lmvn_dist_synchronize(self.handle), "lmvn_dist_synchronize")


class LocalSlabGroup:
    """All ranks in this process.  ``devices[r]`` is rank r's GPU (default: all on one device)."""

    def __init__(self, library: Library, dims, num_views: int, world: int, devices: Optional[Sequence[int]] = None):
        self.dims = tuple(int(d) for d in dims)
        self.world, self.num_views = int(world), int(num_views)
        devices = list(devices) if devices is not None else [0] * self.world
        self.ranks: List[SlabPlan] = [SlabPlan(library, dims, num_views, r, world, devices[r]) for r in range(world)]
        for a in self.ranks:
            for b in self.ranks:
                if a is not b:
                    a.connect_local(b)

    def close(self):
        for r in self.ranks:
            r.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _sync(self):
        for r in self.ranks:
            r.synchronize()

    def set_view(self, v: int, image, weights, kernel1, kernel2):
        for r in self.ranks:
            r.set_view_slab(v, slab_of(image, r.rank, self.world), slab_of(weights, r.rank, self.world))
        for which, k in ((1, kernel1), (2, kernel2)):
            k = np.ascontiguousarray(k, dtype=np.float32)
            for r in self.ranks:
                r.psf_phase(v, which, 0, k)
            self._sync()
            for r in self.ranks:
                r.psf_phase(v, which, 1)
            self._sync()

    def set_psi(self, psi):
        for r in self.ranks:
            r.set_psi_slab(slab_of(psi, r.rank, self.world))

    def get_psi(self) -> np.ndarray:
        return np.concatenate([r.get_psi_slab() for r in self.ranks], axis=0)

    def iterate(self, iterations: int, lam: float = 0.0, min_value: float = 1e-4):
        for _ in range(int(iterations)):
            for v in range(self.num_views):
                for which in (1, 2):
                    for phase in (0, 1, 2):
                        for r in self.ranks:
                            r.conv_phase(v, which, phase, lam, min_value)
                        self._sync()


class ProcessSlabPlan(SlabPlan):
    """One process per GPU.  ``dist`` is an initialised ``torch.distributed`` (any backend that can
    all_gather_object; the data path itself never goes through it)."""

    def __init__(self, library: Library, dims, num_views: int, dist, device: int):
        super().__init__(library, dims, num_views, dist.get_rank(), dist.get_world_size(), device)
        self.dist = dist
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, self.export_handle())
        for peer, h in enumerate(handles):
            if peer != self.rank:
                self.connect_ipc(peer, h)
        dist.barrier()

    def set_view(self, v: int, image_slab, weights_slab, kernel1, kernel2):
        """Collective: every rank passes its slab of the view and the whole (small) kernels."""
        self.set_view_slab(v, image_slab, weights_slab)
        for which, k in ((1, kernel1), (2, kernel2)):
            self.barrier()  # the previous reader of the exchange buffers is done everywhere
            self.psf_phase(v, which, 0, np.ascontiguousarray(k, dtype=np.float32))
            self.barrier()
            self.psf_phase(v, which, 1)
        self.synchronize()


    def _host_barrier(self):
        self.synchronize()
        self.dist.barrier()

    def set_view_host_barriers(self, v: int, image_slab, weights_slab, kernel1, kernel2):
        """set_view with host-side barriers between the phases (any backend; used by the CPU tests)."""
        self.set_view_slab(v, image_slab, weights_slab)
        for which, k in ((1, kernel1), (2, kernel2)):
            self._host_barrier()
            self.psf_phase(v, which, 0, np.ascontiguousarray(k, dtype=np.float32))
            self._host_barrier()
            self.psf_phase(v, which, 1)
        self._host_barrier()

    def iterate(self, iterations: int, lam: float = 0.0, min_value: float = 1e-4) -> float:
        """lmvn_dist_iterate with the hosts lined up first: the device-side barriers then only absorb launch
        skew, not seconds of host skew (first-call graph instantiation, staging of pageable slabs)."""
        self._host_barrier()
        return super().iterate(iterations, lam, min_value)

    def reset_barrier(self):
        """Collective recovery after a timed-out device barrier (lmvn_dist_reset_barrier)."""
        self._host_barrier()
        self.L._check(self.L.lib.lmvn_dist_reset_barrier(self.handle), "lmvn_dist_reset_barrier")
        self._host_barrier()

    def iterate_host_barriers(self, iterations: int, lam: float = 0.0, min_value: float = 1e-4):
        """The phase sequence of lmvn_dist_iterate with torch.distributed barriers instead of the device-side
        flag barrier: slower, but independent of peer-visible flags (and what the gloo tests on the CPU drive)."""
        self._host_barrier()
        for _ in range(int(iterations)):
            for v in range(self.num_views):
                for which in (1, 2):
                    self.conv_phase(v, which, 0, lam, min_value)
                    self._host_barrier()
                    self.conv_phase(v, which, 1, lam, min_value)
                    self._host_barrier()
                    self.conv_phase(v, which, 2, lam, min_value)

    # ---- comparator: the same exchanges through NCCL all_to_all_single ------------------------------
    def _tensor(self, which: int, torch):
        ptr, nbytes = C.c_void_p(), C.c_ulonglong()
        self.L._check(self.L.lib.lmvn_dist_buffer(self.handle, which, C.byref(ptr), C.byref(nbytes)), "lmvn_dist_buffer")

        class _Raw:  # device memory owned by the plan, exposed through the CUDA array interface
            __cuda_array_interface__ = {"shape": (int(nbytes.value) // 4,), "typestr": "<f4", "data": (int(ptr.value), False),
                                        "version": 3}
        return torch.as_tensor(_Raw(), device="cuda")

    def enable_nccl_comparator(self):
        import torch

        torch.cuda.synchronize()
        self.synchronize()
        self.L._check(self.L.lib.lmvn_dist_set_stream(self.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                      "lmvn_dist_set_stream")
        self.L._check(self.L.lib.lmvn_dist_set_staged(self.handle, 1), "lmvn_dist_set_staged")
        info = self.info()
        g, nzl, nyl, p2 = self.world, info.planes_per_rank, info.rows_per_rank, 2 * info.spectrum_pitch
        n = g * nzl * nyl * p2  # floats of one local spectrum
        self._t = {
            "slab": self._tensor(0, torch)[:n].view(nzl, g, nyl, p2),
            "pencil": self._tensor(1, torch)[:n],
            "send": self._tensor(2, torch)[:n],
            "recv": self._tensor(3, torch)[:n],
        }

    def iterate_nccl(self, iterations: int, lam: float = 0.0, min_value: float = 1e-4) -> float:
        """Same loop, exchanges = local scatter + torch.distributed.all_to_all_single (NCCL) + one interleaving
        copy on the way back.  Returns the CUDA-event time in ms."""
        import torch

        t, g = self._t, self.world
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.dist.barrier()
        e0.record()
        for _ in range(int(iterations)):
            for v in range(self.num_views):
                for which in (1, 2):
                    self.conv_phase(v, which, 0, lam, min_value)
                    self.dist.all_to_all_single(t["pencil"], t["send"])
                    self.conv_phase(v, which, 1, lam, min_value)
                    self.dist.all_to_all_single(t["recv"], t["send"])
                    t["slab"].copy_(t["recv"].view(g, t["slab"].shape[0], t["slab"].shape[2], t["slab"].shape[3]).permute(1, 0, 2, 3))
                    self.conv_phase(v, which, 2, lam, min_value)
        e1.record()
        torch.cuda.synchronize()
        return float(e0.elapsed_time(e1))
