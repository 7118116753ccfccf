"""Multi-process host logic of the slab-decomposed plans: world_size-2 `gloo` group on
the CPU.  The "device" is the host-emulated build of the library, whose exchange regions are POSIX shared memory,
so the real N > 1 path runs without a GPU: rendezvous, exchange of the 64-byte handles, stores into the peer's
memory from the transform passes, phase barriers, gather for checking."""
import os
import socket

import numpy as np

DIMS, VIEWS, ITERS, LAM = (32, 32, 64), 2, 2, 0.006


def _data():
    from libmultiviewnative_b200.synthetic import make_views

    return make_views(DIMS, num_views=VIEWS, kernel_size=5, n_sources=8, workers=1, seed=11)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, emu_path, out_dir):
    import torch.distributed as dist

    from libmultiviewnative_b200.capi import Library
    from libmultiviewnative_b200.slabs import ProcessSlabPlan, slab_of

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = Library(emu_path)
    d = _data()
    plan = ProcessSlabPlan(lib, DIMS, VIEWS, dist, 0)
    info = plan.info()
    assert (info.rank, info.world, info.planes_per_rank) == (rank, world, DIMS[0] // world)
    for v in range(VIEWS):
        plan.set_view_host_barriers(v, slab_of(d["views"][v], rank, world), slab_of(d["weights"][v], rank, world),
                                    d["kernels1"][v], d["kernels2"][v])
    plan.set_psi_slab(slab_of(d["psi0"], rank, world))
    plan.iterate_host_barriers(ITERS, LAM, 1e-4)
    mine = plan.get_psi_slab()
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    dist.barrier()
    plan.close()
    if rank == 0:
        np.save(os.path.join(out_dir, "psi.npy"), np.concatenate(parts, axis=0))
    dist.destroy_process_group()


def test_two_rank_slab_run_matches_single_plan_and_oracle(tmp_path):
    import torch.multiprocessing as mp

    from libmultiviewnative_b200 import capi
    from libmultiviewnative_b200._build import build_emu
    from oracle import mvn_oracle as orc
    from tests import parity_cases as pc

    emu = build_emu()
    mp.spawn(_worker, args=(2, _free_port(), emu, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "psi.npy")
    d = _data()
    L = capi.Library(emu)
    L.set_default_strategy(capi.STRATEGY_FUSED)
    try:
        single = d["psi0"].copy()
        L.inplace_gpu_deconvolve(single, d["views"], d["kernels1"], d["kernels2"], d["weights"], ITERS, LAM, 1e-4)
    finally:
        L.set_default_strategy(capi.STRATEGY_AUTO)
    np.testing.assert_array_equal(got, single)
    exp = orc.inplace_cpu_deconvolve(d["psi0"], d["views"], d["kernels1"], d["kernels2"], d["weights"], ITERS, LAM, 1e-4)
    assert pc.max_rel(got, exp) < pc.PER_VOXEL_TOL_1_ITER
