"""Multi-process host logic of the block-sharded path:
world_size-2 `gloo` group on CPU.  The "device" is the host-emulated build of the
library (tests/emu), so the whole N > 1 path -- rendezvous, sharding, per-rank
C-ABI calls, gather for checking -- runs without a GPU."""
import os
import socket

import numpy as np
import pytest

from libmultiviewnative_b200 import blocks


def test_shard_round_robin_covers_every_block_once():
    for n in (0, 1, 7, 64):
        for world in (1, 2, 4, 8):
            seen = sorted(b for r in range(world) for b in blocks.shard(n, r, world))
            assert seen == list(range(n))
            sizes = [len(blocks.shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        blocks.shard(4, 2, 2)


def _make_block(b):
    from libmultiviewnative_b200.synthetic import make_views

    return make_views((16, 16, 64), num_views=2, kernel_size=5, n_sources=6, seed=100 + b, workers=1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, emu_path, n_blocks, out_dir):
    import torch.distributed as dist

    from libmultiviewnative_b200.capi import Library

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = Library(emu_path)
    res = blocks.run_sharded(lib, _make_block, n_blocks, 2, 0.006, 1e-4, device=0, gather=True)
    dist.barrier()
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), **{str(k): v for k, v in res.items()})
    else:
        assert res is None
    dist.destroy_process_group()


def test_two_rank_sharded_run_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    from libmultiviewnative_b200._build import build_emu
    from libmultiviewnative_b200.capi import Library

    emu = build_emu()
    n_blocks = 5  # ragged on purpose: rank 0 gets 3 blocks, rank 1 gets 2
    mp.spawn(_worker, args=(2, _free_port(), emu, n_blocks, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    assert sorted(int(k) for k in got.files) == list(range(n_blocks))
    lib = Library(emu)
    for b in range(n_blocks):
        exp = blocks.deconvolve_block(lib, _make_block(b), 2, 0.006, 1e-4, 0)
        np.testing.assert_array_equal(got[str(b)], exp)  # same code, same bits, whichever rank ran it


def test_run_threads_matches_serial():
    from libmultiviewnative_b200._build import build_emu
    from libmultiviewnative_b200.capi import Library

    lib = Library(build_emu())
    blks = [_make_block(b) for b in range(3)]
    # the emulator is single threaded by design: one "device" slot
    res = blocks.run_threads(lib, blks, [0], 1, 0.0, 1e-4)
    for b in range(3):
        np.testing.assert_array_equal(res[b], blocks.deconvolve_block(lib, blks[b], 1, 0.0, 1e-4, 0))
