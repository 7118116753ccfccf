"""Secondary workloads of bench.py (BASELINE.json configs 2, 4 and 5).  The default bench.py run is
config 3; these print the same kind of JSON line for the other measurement rows of SURVEY.md §8d."""
from __future__ import annotations

import json
import os
import time

import numpy as np

SWEEP_SHAPES = [(64, 64, 64), (128, 64, 64), (128, 128, 64), (128, 128, 128), (256, 128, 128), (256, 256, 128),
                (256, 256, 256), (512, 256, 256), (512, 512, 256), (512, 512, 512)]  # ref: python/generate_dims.py:4-48
SWEEP_KERNELS = [15, 21, 31, 41, 63]                                                  # 21 = reference default


def conv_bytes(dims):
    """Device-resident convolution with a precomputed K^: image S->C, (C + K^ C)->C, C->S = 2S + 5C.
    (SURVEY.md §8d's B_conv = 2S + 6C additionally counts building K^ once per call; that part is in e2e_ms.)"""
    nz, ny, nx = dims
    S = 4 * nz * ny * nx
    C = 8 * nz * ny * (nx // 2 + 1)
    return 2 * S + 5 * C


def conv_sweep(args, lib, torch, peak, peak_src, device=0):
    """Config 2: single-view FFT convolution, device-resident (CUDA events) and through
    inplace_gpu_convolution with host pointers; cuFFT (torch.fft) pipeline timed beside it."""
    from libmultiviewnative_b200.synthetic import gaussian_psf
    from oracle import mvn_oracle as orc  # checker only: correctness of every shape, outside the timed regions

    rows = []
    rng = np.random.default_rng(11)
    shapes = SWEEP_SHAPES if not args.dims else [tuple(int(x) for x in args.dims.split(","))]
    for dims in shapes:
        img = (rng.random(dims, dtype=np.float32) + 1.0).astype(np.float32)
        timg = torch.from_numpy(img).pin_memory()
        for ks in SWEEP_KERNELS:
            if ks > min(dims):
                continue
            k = gaussian_psf(ks, (ks / 8.0, ks / 10.0, ks / 12.0))
            with lib.plan(dims, 1, device) as p:
                p.set_view(0, img, img, k, k)
                p.set_psi(img)
                info = p.info()
                p.convolve(0, 1, 1)                      # warm-up (reference: 1 warm-up, 10 timed repeats)
                ms = p.convolve(0, 1, 10) / 10.0
                strategy = info.strategy
            # end to end through the reference entry point, host pointers (pinned)
            work = timg.numpy()
            out = work.copy()
            lib.inplace_gpu_convolution(out, k, device)   # warm-up + result for the check
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                lib.inplace_gpu_convolution(work, k, device)
            e2e_ms = (time.perf_counter() - t0) / reps * 1e3
            np.copyto(work, img)
            # cuFFT comparator: rfftn -> multiply -> irfftn with a precomputed kernel spectrum
            x = torch.from_numpy(img).to("cuda:%d" % device)
            kh = torch.fft.rfftn(torch.from_numpy(orc.wrap_kernel(k, dims)).to(x.device))
            torch.fft.irfftn(torch.fft.rfftn(x) * kh, s=dims)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                y = torch.fft.irfftn(torch.fft.rfftn(x) * kh, s=dims)
            e1.record()
            torch.cuda.synchronize()
            cufft_ms = e0.elapsed_time(e1) / 10.0
            # every shape and every kernel size of the sweep is checked against the oracle (rel L2 <= 1e-5)
            exp = orc.inplace_cpu_convolution(img, k, nthreads=-1)
            err = float(np.linalg.norm(out.astype(np.float64) - exp) / np.linalg.norm(exp.astype(np.float64)))
            assert err < 1e-5, (dims, ks, err)
            del exp
            del x, kh, y
            nb = conv_bytes(dims)
            rows.append({"dims_zyx": list(dims), "kernel": ks, "strategy": int(strategy), "ms": ms, "GBps": nb / (ms * 1e-3) / 1e9,
                         "frac_of_peak": nb / (ms * 1e-3) / 1e9 / peak, "e2e_ms": e2e_ms, "cufft_pipeline_ms": cufft_ms,
                         "speedup_vs_cufft": cufft_ms / ms, "rel_l2_vs_oracle": err})
    big = max(rows, key=lambda r: (np.prod(r["dims_zyx"]), -abs(r["kernel"] - 21)))
    return {
        "metric": "fft_conv_GBps", "value": big["GBps"], "unit": "GB/s", "n_gpus": 1, "steps": 10, "warmup": 1,
        "ms_per_step": big["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "config 2: single-view 3-D FFT convolution sweep, bytes = 2S + 5C per device-resident convolution",
                   "headline_row": {"dims_zyx": big["dims_zyx"], "kernel": big["kernel"]},
                   "peak": peak, "peak_source": peak_src},
        "sweep": rows,
    }


def blocks(args, lib, torch, dist, rank, world, device, fast_views, barrier, max_over_ranks):
    """Config 4: independent 256^3 blocks (6 views, 41^3 PSFs, 50 iterations) sharded b -> GPU b mod G,
    every block through inplace_gpu_deconvolve with host buffers (H2D/D2H inside the timed region)."""
    from libmultiviewnative_b200.blocks import shard

    dims = tuple(int(x) for x in args.dims.split(",")) if args.dims else (256, 256, 256)
    n_blocks = args.blocks
    workers = max(1, (os.cpu_count() or 1) // max(1, world))
    distinct = [fast_views(dims, args.views, args.kernel, 20240607 + 17 * i, workers) for i in range(4)]

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = []
    for d in distinct:
        for key in ("views", "weights"):
            for i, a in enumerate(d[key]):
                tt, d[key][i] = pin(a)
                keep.append(tt)
    tt, psi = pin(distinct[0]["psi0"].copy())
    keep.append(tt)
    mine = shard(n_blocks, rank, world)

    def run_block(b):
        d = distinct[b % len(distinct)]
        np.copyto(psi, d["psi0"])
        t0 = time.perf_counter()
        lib.inplace_gpu_deconvolve(psi, d["views"], d["kernels1"], d["kernels2"], d["weights"], args.iterations, 0.006,
                                   1e-4, device)
        return time.perf_counter() - t0

    run_block(mine[0] if mine else 0)  # warm-up
    barrier()
    t_wall = time.perf_counter()
    busy = sum(run_block(b) for b in mine)
    barrier()
    wall = max_over_ranks(time.perf_counter() - t_wall)
    busy = max_over_ranks(busy)
    nvox = float(np.prod(dims))
    units = nvox * args.views * args.iterations * n_blocks
    return {
        "metric": "rl_deconv_gvoxel_view_iter_per_s", "value": units / busy / 1e9, "unit": "Gvoxel*view*iter/s",
        "n_gpus": world, "steps": n_blocks, "warmup": 1, "ms_per_step": busy / max(1, len(mine)) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config 4: %d independent %dx%dx%d blocks, %d views, %d^3 PSFs, %d iterations, block b on GPU b mod G, "
                               "no collective; every block through inplace_gpu_deconvolve with pinned host buffers" % (
                                   n_blocks, dims[0], dims[1], dims[2], args.views, args.kernel, args.iterations),
                   "blocks_per_rank": len(mine), "wall_s_incl_host_psi_reset": wall},
        "e2e": {"value": units / busy / 1e9, "unit": "Gvoxel*view*iter/s",
                "h2d_bytes_per_step": int((2 * args.views + 1) * nvox * 4), "d2h_bytes_per_step": int(nvox * 4)},
    }


def volume(args, lib, torch, dist, rank, world, device, barrier, max_over_ranks, peak, peak_src):
    """Config 5: ONE volume over G GPUs: slab-decomposed FFT, exchanges fused into the transform kernels
    as stores into peer memory over NVLink.  Strong scaling; G = 1 runs the ordinary plan."""
    from libmultiviewnative_b200.slabs import ProcessSlabPlan
    from libmultiviewnative_b200.synthetic import gaussian_psf

    dims = tuple(int(x) for x in args.dims.split(",")) if args.dims else (512, 512, 256)
    nvox = float(np.prod(dims))
    iters = args.iterations
    rng = np.random.default_rng(5 + rank)
    nz_l = dims[0] // world
    slab = (nz_l, dims[1], dims[2])
    # timing is data independent: one random slab serves as every view, constant weights
    img = (rng.random(slab, dtype=np.float32) + 1.0).astype(np.float32)
    wts = np.full(slab, 1.0 / args.views, dtype=np.float32)
    k = gaussian_psf(args.kernel, (4.0, 1.5, 1.5))
    k2 = np.ascontiguousarray(k[::-1, ::-1, ::-1])
    exch = 0
    if world == 1:
        plan = lib.plan(dims, args.views, device)
        for v in range(args.views):
            plan.set_view(v, img, wts, k, k2)
        plan.set_psi(img)
        run = lambda n: plan.iterate(n, 0.006, 1e-4)
    else:
        plan = ProcessSlabPlan(lib, dims, args.views, dist, device)
        for v in range(args.views):
            plan.set_view(v, img, wts, k, k2)
        plan.set_psi_slab(img)
        exch = int(plan.info().exchange_bytes_per_view_iteration)
        run = lambda n: plan.iterate(n, 0.006, 1e-4)
    for _ in range(args.warmup):
        run(iters)
    barrier()
    ms = 0.0
    for _ in range(args.steps):
        ms += run(iters)
    barrier()
    ms = max_over_ranks(ms)
    units = nvox * args.views * iters * args.steps
    S = 4 * nvox
    C = 8 * dims[0] * dims[1] * (dims[2] // 2 + 1)
    alg = 7 * S + 10 * C
    gbps = alg * args.views * iters * args.steps / (ms * 1e-3) / 1e9
    plan.close()
    return {
        "metric": "rl_deconv_gvoxel_view_iter_per_s", "value": units / (ms * 1e-3) / 1e9, "unit": "Gvoxel*view*iter/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config 5: ONE %dx%dx%d volume, %d views, %d^3 PSFs, %d iterations, slabs of nz/G planes, "
                               "pencils of ny/G rows, exchanges = P2P stores from the y / z transform kernels" % (
                                   dims[0], dims[1], dims[2], args.views, args.kernel, iters),
                   "exchange_bytes_per_view_iteration_all_ranks": exch},
        "roofline": {"bound": "hbm", "achieved": gbps, "peak": peak * world, "unit": "GB/s", "frac": gbps / (peak * world),
                     "peak_source": peak_src + " x n_gpus", "alg_bytes_per_view_iteration": alg,
                     "nvlink_GBps_per_gpu_out": (exch / world) * args.views * iters * args.steps / (ms * 1e-3) / 1e9 if world > 1 else 0.0},
    }


# ------------------------------------------------------------------------------------------------------------
# Sub-records of the default bench.py run: BASELINE configs 4 and 5 in front of the driver (every --gpus N run
# of the headline also measures them; N = 1 emits the single-GPU baselines).
# ------------------------------------------------------------------------------------------------------------
def _pin(torch, a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def blocks_record(lib, torch, dist, rank, world, device, barrier, max_over_ranks, n_blocks=64, n1_blocks=8, depth=2,
                  dims=(256, 256, 256), views=6, kernel=41, iterations=50):
    """Config 4: n_blocks independent 256^3 blocks (6 views, 41^3 PSFs, 50 iterations) through inplace_gpu_deconvolve
    with PAGEABLE host buffers (what JNA hands over), block b on GPU b mod G, no collective, two calls in flight per
    GPU (blocks.run_pipelined).  Strong scaling.  The same-build one-GPU number it is compared with is measured in
    the same invocation: rank 0 alone runs n1_blocks blocks before the sharded run."""
    from libmultiviewnative_b200.blocks import run_pipelined, shard
    from libmultiviewnative_b200.synthetic import make_views_fast

    workers = max(1, (os.cpu_count() or 1) // max(1, world))
    distinct = [make_views_fast(dims, views, kernel, 20240607 + 17 * i, workers) for i in range(2)]  # cycled (host RAM)
    nvox = float(np.prod(dims))

    def timed(indices, calls_in_flight=depth):
        t0 = time.perf_counter()
        run_pipelined(lib, lambda b: distinct[b % len(distinct)], indices, iterations, 0.006, 1e-4, device,
                      depth=calls_in_flight, keep=False)
        return time.perf_counter() - t0

    # the same blocks in page-locked buffers (a caller that can allocate its stacks with cudaHostAlloc)
    keep, pinned = [], []
    for blk in distinct:
        q = dict(blk)
        for key in ("views", "weights"):
            q[key] = []
            for a in blk[key]:
                t, arr = _pin(torch, a)
                keep.append(t)
                q[key].append(arr)
        pinned.append(q)

    def timed_pinned(indices):
        t0 = time.perf_counter()
        run_pipelined(lib, lambda b: pinned[b % len(pinned)], indices, iterations, 0.006, 1e-4, device, depth=depth, keep=False)
        return time.perf_counter() - t0

    timed([0, 1])  # warm-up: plan store, arenas, staging ring
    barrier()
    one_gpu = one_gpu_seq = one_gpu_pinned = None
    if rank == 0:
        per = nvox * views * iterations * n1_blocks / 1e9
        one_gpu = per / timed(list(range(n1_blocks)))
        one_gpu_seq = per / timed(list(range(n1_blocks)), 1)  # the same blocks, one call at a time (no overlap)
        one_gpu_pinned = per / timed_pinned(list(range(n1_blocks)))
    barrier()
    mine = shard(n_blocks, rank, world)
    t0 = time.perf_counter()
    timed(mine)
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    timed_pinned(mine)
    barrier()
    wall_pinned = max_over_ranks(time.perf_counter() - t0)
    del keep, pinned
    if rank != 0:
        return None
    value = nvox * views * iterations * n_blocks / wall / 1e9
    value_pinned = nvox * views * iterations * n_blocks / wall_pinned / 1e9
    return {"workload": "config 4: %d independent %dx%dx%d blocks, %d views, %d^3 PSFs, %d iterations, through "
                        "inplace_gpu_deconvolve with pageable host buffers, block b on GPU b mod G, %d calls in flight per GPU, "
                        "no collective" % (n_blocks, dims[0], dims[1], dims[2], views, kernel, iterations, depth),
            "value": value, "unit": "Gvoxel*view*iter/s", "scaling": "strong", "n_gpus": world, "wall_s": wall,
            "one_gpu_same_build": {"value": one_gpu, "blocks": n1_blocks, "one_call_at_a_time": one_gpu_seq,
                                   "pinned_host_buffers": one_gpu_pinned},
            "speedup_vs_one_gpu": value / one_gpu,
            # the same batch from page-locked host buffers: no host-side staging copy (8 ranks share the host's cores
            # and memory bandwidth for that copy; on the 32-core 8 x B200 box it is the limit of the pageable batch)
            "pinned_host_buffers": {"value": value_pinned, "wall_s": wall_pinned,
                                    "speedup_vs_one_gpu_pinned": value_pinned / one_gpu_pinned},
            "h2d_bytes_per_block": int((2 * views + 1) * nvox * 4), "d2h_bytes_per_block": int(nvox * 4)}


def volume_record(lib, torch, dist, rank, world, device, barrier, max_over_ranks, peak, dims=(1024, 1024, 1024), views=6,
                  kernel=41, iterations=2, warmup=1, link_gbps=770.0):
    """Config 5: ONE 1024^3 6-view volume, slab-decomposed over the G ranks (P2P-fused exchanges); G = 1: the ordinary
    plan.  Also, for G > 1: the same-build single-GPU time of the same volume (rank 0 alone, before the slab plan is
    created) and a 256^3 parity run of the multi-process path against the single-GPU plan."""
    from libmultiviewnative_b200.slabs import ProcessSlabPlan, slab_of
    from libmultiviewnative_b200.synthetic import gaussian_psf, make_views

    nvox = float(np.prod(dims))
    k = gaussian_psf(kernel, (4.0, 1.5, 1.5))
    k2 = np.ascontiguousarray(k[::-1, ::-1, ::-1])
    S = 4 * nvox
    C = 8 * dims[0] * dims[1] * (dims[2] // 2 + 1)
    alg = 7 * S + 10 * C

    def single_gpu_ms():
        rng = np.random.default_rng(5)
        img = (rng.random(dims, dtype=np.float32) + 1.0).astype(np.float32)  # timing is data independent: one stack for all views
        wts = np.full(dims, 1.0 / views, dtype=np.float32)
        with lib.plan(dims, views, device) as plan:
            for v in range(views):
                plan.set_view(v, img, wts, k, k2)
            plan.set_psi(img)
            for _ in range(warmup):
                plan.iterate(iterations, 0.006, 1e-4)
            return plan.iterate(iterations, 0.006, 1e-4)

    rec = {"workload": "config 5: ONE %dx%dx%d volume, %d views, %d^3 PSFs, %d timed iterations (%d warm-up)" % (
        dims[0], dims[1], dims[2], views, kernel, iterations, warmup), "unit": "Gvoxel*view*iter/s", "scaling": "strong",
        "n_gpus": world}
    one_ms = None
    if rank == 0:
        one_ms = single_gpu_ms()
        lib.release_cached_memory()
        rec["one_gpu_same_build"] = {"value": nvox * views * iterations / (one_ms * 1e-3) / 1e9, "ms": one_ms,
                                     "roofline_frac": alg * views * iterations / (one_ms * 1e-3) / 1e9 / peak}
    if world == 1:
        rec["value"] = rec["one_gpu_same_build"]["value"]
        return rec
    barrier()
    # ---- parity of the multi-process path: 256^3 through the same code vs the single-GPU plan ----
    pd = (256, 256, 256)
    d = make_views(pd, num_views=2, kernel_size=15, n_sources=100, workers=4, seed=7)  # same on every rank
    plan = ProcessSlabPlan(lib, pd, 2, dist, device)
    for v in range(2):
        plan.set_view(v, slab_of(d["views"][v], rank, world), slab_of(d["weights"][v], rank, world), d["kernels1"][v], d["kernels2"][v])
    plan.set_psi_slab(slab_of(d["psi0"], rank, world))
    plan.iterate(2, 0.006, 1e-4)
    mine = torch.from_numpy(plan.get_psi_slab()).to("cuda:%d" % device)
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    plan.close()
    if rank == 0:
        got = torch.cat(parts, 0).cpu().numpy()
        single = d["psi0"].copy()
        lib.inplace_gpu_deconvolve(single, d["views"], d["kernels1"], d["kernels2"], d["weights"], 2, 0.006, 1e-4, device)
        rel = float(np.max(np.abs(got - single) / np.abs(single)))
        rec["parity_256cubed_vs_single_gpu"] = {"max_rel": rel, "bit_identical": bool(np.array_equal(got, single)),
                                                "ok": bool(rel < 5e-6)}
    lib.release_cached_memory()
    barrier()
    # ---- the timed volume ----
    nz_l = dims[0] // world
    slab = (nz_l, dims[1], dims[2])
    rng = np.random.default_rng(5 + rank)
    img = (rng.random(slab, dtype=np.float32) + 1.0).astype(np.float32)
    wts = np.full(slab, 1.0 / views, dtype=np.float32)
    plan = ProcessSlabPlan(lib, dims, views, dist, device)
    for v in range(views):
        plan.set_view(v, img, wts, k, k2)
    plan.set_psi_slab(img)
    exch = int(plan.info().exchange_bytes_per_view_iteration)
    for _ in range(warmup):
        plan.iterate(iterations, 0.006, 1e-4)
    barrier()
    ms = max_over_ranks(plan.iterate(iterations, 0.006, 1e-4))
    plan.close()
    lib.release_cached_memory()
    if rank != 0:
        return None
    units = nvox * views * iterations
    link_ms = (exch / world) * views * iterations / (link_gbps * 1e9) * 1e3   # bytes out of one GPU / measured peer rate
    comp_ms = one_ms / world                                                  # the single-GPU time split perfectly
    rec.update({
        "value": units / (ms * 1e-3) / 1e9, "ms": ms, "speedup_vs_one_gpu": one_ms / ms,
        "exchange_bytes_per_view_iteration_all_ranks": exch,
        "nvlink_GBps_per_gpu_out": (exch / world) * views * iterations / (ms * 1e-3) / 1e9,
        "link_ms_model": link_ms, "compute_ms_model": comp_ms,
        # 1 = the shorter of (link time at %g GB/s, compute time = one-GPU time / G) is completely hidden behind the
        # other; 0 = they add up
        "overlap_fraction": max(0.0, min(1.0, (link_ms + comp_ms - ms) / max(1e-9, min(link_ms, comp_ms)))),
        "overlap_definition": "(t_link + t_compute - t_measured) / min(t_link, t_compute), t_link = bytes out of one GPU / "
                              "%g GB/s (measured peer copy rate, B200_PROFILING.md), t_compute = same-build one-GPU time / G" % link_gbps,
        "roofline_frac_of_G_x_hbm": alg * views * iterations / (ms * 1e-3) / 1e9 / (peak * world),
    })
    return rec
