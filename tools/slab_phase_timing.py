#!/usr/bin/env python
"""Where does a convolution of the slab-decomposed plan spend its time?  One rank per GPU (torchrun); every phase of
`lmvn_dist_conv_phase` is run on all ranks between host barriers and timed on the host (max over ranks), then the
fused loop (`lmvn_dist_iterate`) is timed with the column-window pipeline off and on (LMVN_DIST_GROUPS).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        tools/slab_phase_timing.py 1024,1024,1024
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultiviewnative_b200 import load  # noqa: E402
from libmultiviewnative_b200.slabs import ProcessSlabPlan  # noqa: E402
from libmultiviewnative_b200.synthetic import gaussian_psf  # noqa: E402


def main():
    dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1024,1024,1024").split(","))
    views = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load()
    slab = (dims[0] // world, dims[1], dims[2])
    rng = np.random.default_rng(5 + rank)
    img = (rng.random(slab, dtype=np.float32) + 1.0).astype(np.float32)
    wts = np.full(slab, 1.0 / views, dtype=np.float32)
    k = gaussian_psf(31, (4.0, 1.5, 1.5))
    k2 = np.ascontiguousarray(k[::-1, ::-1, ::-1])
    out = {"dims_zyx": list(dims), "world": world, "views": views}

    def mx(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for groups in (1, 2, 4):
        os.environ["LMVN_DIST_GROUPS"] = str(groups)
        plan = ProcessSlabPlan(lib, dims, views, dist, local)
        for v in range(views):
            plan.set_view(v, img, wts, k, k2)
        plan.set_psi_slab(img)
        if groups == 1:
            # phases of one convolution, host timed (unchained form: x forward | y forward + scatter || z pass + scatter || y inverse | x inverse)
            plan.iterate_host_barriers(1, 0.006, 1e-4)  # warm-up
            ph = [0.0, 0.0, 0.0]
            reps = 3
            for _ in range(reps):
                for which in (1, 2):
                    for phase in (0, 1, 2):
                        plan._host_barrier()
                        t0 = time.perf_counter()
                        plan.conv_phase(0, which, phase, 0.006, 1e-4)
                        plan.synchronize()
                        ph[phase] += time.perf_counter() - t0
            info = plan.info()
            exch = int(info.exchange_bytes_per_view_iteration)
            out["phase_ms_per_conv"] = {"x_fwd+y_fwd_scatter": mx(ph[0]) / (2 * reps) * 1e3, "z_pass_scatter": mx(ph[1]) / (2 * reps) * 1e3,
                                        "y_inv+x_inv": mx(ph[2]) / (2 * reps) * 1e3}
            out["scatter_bytes_out_per_gpu_per_exchange"] = exch / 4 / world
            plan.set_psi_slab(img)
        plan.iterate(3, 0.006, 1e-4)  # warm-up + graph capture
        ms = mx(min(plan.iterate(3, 0.006, 1e-4) for _ in range(2)))
        out["iterate_ms_per_view_iteration_groups_%d" % groups] = ms / (3 * views)
        plan.close()
        lib.release_cached_memory()
        dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
