#!/usr/bin/env python
"""Multi-process check of the slab-decomposed plans (run under torchrun, one rank per GPU):
every rank deconvolves its slab with P2P-fused exchanges; rank 0 gathers psi and compares it
with the single-GPU plan (identical up to the FMA contraction of the chained x passes: a few ulp),
then prints timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/slab_mp_check.py 256,256,256 2 5
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.slabs import ProcessSlabPlan, slab_of  # noqa: E402
from libmultiviewnative_b200.synthetic import make_views  # noqa: E402


def main():
    dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "128,128,128").split(","))
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load()
    d = make_views(dims, num_views=nv, kernel_size=15, n_sources=100, workers=4, seed=7)  # same on every rank
    plan = ProcessSlabPlan(lib, dims, nv, dist, local)
    for v in range(nv):
        plan.set_view(v, slab_of(d["views"][v], rank, world), slab_of(d["weights"][v], rank, world),
                      d["kernels1"][v], d["kernels2"][v])
    plan.set_psi_slab(slab_of(d["psi0"], rank, world))
    ms = plan.iterate(iters, 0.006, 1e-4)
    mine = torch.from_numpy(plan.get_psi_slab()).cuda()
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    # timing: restart from psi0, more iterations
    plan.set_psi_slab(slab_of(d["psi0"], rank, world))
    plan.iterate(3, 0.006, 1e-4)  # warm-up; also captures the CUDA graph of one sweep
    t = torch.tensor([plan.iterate(10, 0.006, 1e-4)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # comparator: the same exchanges through NCCL all_to_all_single
    plan.enable_nccl_comparator()
    plan.set_psi_slab(slab_of(d["psi0"], rank, world))
    plan.iterate_nccl(iters, 0.006, 1e-4)
    nccl_psi, fused_psi = plan.get_psi_slab(), mine.cpu().numpy()
    nccl_same = bool(np.array_equal(nccl_psi, fused_psi))
    # (the comparator drives the UNCHAINED phases; where the chained kernels of this shape contract their FMAs differently
    # the two agree to a few ulp instead of bit for bit)
    nccl_rel = torch.tensor([float(np.max(np.abs(nccl_psi - fused_psi) / np.abs(fused_psi)))], device="cuda")
    dist.all_reduce(nccl_rel, op=dist.ReduceOp.MAX)
    plan.iterate_nccl(2, 0.006, 1e-4)
    tn = torch.tensor([plan.iterate_nccl(10, 0.006, 1e-4)], device="cuda")
    dist.all_reduce(tn, op=dist.ReduceOp.MAX)
    flags = torch.tensor([1.0 if nccl_same else 0.0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("slab_mp_check NCCL comparator: identical to the P2P-fused path on every rank = %s (max rel %.2g); %.3f ms/(view,iter) vs "
              "%.3f ms fused (fused is %.2fx faster)" % (bool(flags.item() > 0), nccl_rel.item(), tn.item() / (10 * nv),
                                                        t.item() / (10 * nv), tn.item() / t.item()), flush=True)
        assert nccl_rel.item() < 5e-6
        got = torch.cat(parts, 0).cpu().numpy()
        single = d["psi0"].copy()
        lib.inplace_gpu_deconvolve(single, d["views"], d["kernels1"], d["kernels2"], d["weights"], iters, 0.006, 1e-4, local)
        same = bool(np.array_equal(got, single))
        rel = float(np.max(np.abs(got - single) / np.abs(single)))
        with lib.plan(dims, nv, local) as p:
            for v in range(nv):
                p.set_view(v, d["views"][v], d["weights"][v], d["kernels1"][v], d["kernels2"][v])
            p.set_psi(d["psi0"])
            p.iterate(3, 0.006, 1e-4)
            t1 = p.iterate(10, 0.006, 1e-4)
        nvox = float(np.prod(dims))
        print("slab_mp_check dims=%s world=%d views=%d: identical=%s max_rel=%.3g first_call_ms=%.3f | %d ranks %.3f ms/(view,iter) "
              "= %.1f Gvox/s | single GPU %.3f ms = %.1f Gvox/s | speed-up %.2fx" % (
                  dims, world, nv, same, rel, ms, world, t.item() / (10 * nv), nvox * 10 * nv / (t.item() * 1e-3) / 1e9,
                  t1 / (10 * nv), nvox * 10 * nv / (t1 * 1e-3) / 1e9, t1 / t.item()), flush=True)
        assert same or rel < 5e-6  # the single-GPU loop chains its x passes: same source, other FMA contraction
    dist.barrier()
    plan.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
