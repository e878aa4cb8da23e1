"""Multi-rank host logic on CPU: gloo backend, world_size 2 (and 3), checker = the C port oracle with global-index slabs."""
import os
import socket

import numpy as np
import pytest

from fastnoise2_b200.sharding import slab_range


def test_slab_ranges_partition_exactly():
    for total in (1, 2, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [slab_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tree, shape, result_dir):
    import torch
    import torch.distributed as dist
    from fastnoise2_b200 import sharding
    from tests import helpers
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny, nz = shape
        z0, z1 = sharding.slab_range(nz, rank, world)
        node = helpers.PortNode(helpers.load_port(), tree)
        slab, mm = node.grid3d(0.0, 0.0, 0.0, nx, ny, nz, 2.7, 2.7, 2.7, 1337, rows=(z0 * ny, (z1 - z0) * ny))
        lo, hi = sharding.all_reduce_minmax(float(mm[0]), float(mm[1]))
        counts = [(sharding.slab_range(nz, r, world)[1] - sharding.slab_range(nz, r, world)[0]) * nx * ny for r in range(world)]
        whole = sharding.all_gather_slabs(torch.from_numpy(slab), counts)
        if rank == 0:
            np.save(os.path.join(result_dir, "whole.npy"), whole.numpy())
            np.save(os.path.join(result_dir, "mm.npy"), np.array([lo, hi], np.float32))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_slabs_over_gloo_reassemble_the_grid(tmp_path, world):
    import torch.multiprocessing as mp
    from tests import helpers
    from fastnoise2_b200.api import TREES
    if helpers.load_port() is None:
        pytest.skip("oracle/libfn2oracle.so not built")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    shape = (10, 6, 7)
    mp.spawn(_worker, args=(world, port, TREES["c4_terrain"], shape, str(tmp_path)), nprocs=world, join=True)
    whole = np.load(tmp_path / "whole.npy")
    mm = np.load(tmp_path / "mm.npy")
    want, wmm = helpers.PortNode(helpers.load_port(), TREES["c4_terrain"]).grid3d(0.0, 0.0, 0.0, *shape, 2.7, 2.7, 2.7, 1337)
    assert np.array_equal(whole.view(np.uint32), want.view(np.uint32))
    assert mm[0] == wmm[0] and mm[1] == wmm[1]
