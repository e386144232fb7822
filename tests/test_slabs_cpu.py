"""Host-side logic of the x-slab decomposition, world_size 2 over gloo on CPU: geometry helpers agree across
ranks, every atom has exactly one owner, the 128-byte bootstrap blob of rank 0 reaches the other rank."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch

    from mrmd_b200 import slabs

    gmin, gmax = np.zeros(3), np.array([50.0, 20.0, 20.0])
    rng = np.random.default_rng(5)  # same global configuration on both ranks
    pos = rng.random((5000, 3)) * gmax
    mine = slabs.select_slab(pos, gmin, gmax, rank, world)
    lo, hi = slabs.slab_bounds(gmin, gmax, rank, world)
    assert np.all((pos[mine, 0] >= lo) & (pos[mine, 0] < hi))
    counts = torch.tensor([len(mine)])
    dist.all_reduce(counts)
    assert int(counts) == len(pos)  # a partition: nobody lost, nobody owned twice
    assert np.array_equal(slabs.owner_of(pos[mine, 0], gmin, gmax, world), np.full(len(mine), rank))
    left, right = slabs.neighbours(rank, world)
    assert left == right == 1 - rank
    to_left, to_right = slabs.boundary_shifts(gmin, gmax, rank, world)
    assert (to_left, to_right) == ((50.0, 0.0) if rank == 0 else (0.0, -50.0))
    # an atom leaving rank 0 through the low end of the box lands inside the last rank's slab
    if rank == 0:
        x = -0.3 + to_left
        l1, h1 = slabs.slab_bounds(gmin, gmax, world - 1, world)
        assert l1 <= x < h1
    # bootstrap blob: rank 0's bytes arrive on rank 1 (the real id comes from mrmd_b200_nccl_unique_id on a GPU box)
    blob = np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(blob)
    dist.broadcast(t, src=0)
    assert np.array_equal(t.numpy(), np.arange(128, dtype=np.uint8))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_slab_host_logic_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_balanced_cuts():
    """cost-balanced slab boundaries: equal estimated cost, on lattice planes, minimum width respected, a partition"""
    from mrmd_b200 import slabs

    gmin, gmax = np.array([0.0, 0, 0]), np.array([500.0, 500, 500])
    ratio = 1.67
    cuts = slabs.balanced_cuts(gmin, gmax, 8, 150.0, 350.0, ratio, quantum=1.25, min_width=5.2)
    assert cuts[0] == 0.0 and cuts[-1] == 500.0 and np.all(np.diff(cuts) >= 5.2)
    assert np.allclose(cuts / 1.25, np.round(cuts / 1.25))

    def cost(a, b):
        inside = max(0.0, min(b, 350.0) - max(a, 150.0))
        return (b - a) + ratio * inside

    costs = [cost(cuts[r], cuts[r + 1]) for r in range(8)]
    assert max(costs) - min(costs) <= 2 * 1.25 * (1 + ratio)  # within the rounding to lattice planes
    widths = np.diff(cuts)
    assert widths[0] > widths[3] and widths[7] > widths[4]  # wide over CG, narrow over AT / HY
    x = np.random.default_rng(1).random(1000) * 500.0
    owners = slabs.owner_of(x, gmin, gmax, 8, cuts)
    assert all(cuts[o] <= xi < cuts[o + 1] for xi, o in zip(x, owners))
    # nothing to balance: equal widths
    assert np.allclose(slabs.balanced_cuts(gmin, gmax, 4, 0.0, 0.0, ratio), [0, 125, 250, 375, 500])


def test_slab_bounds_cover_the_box():
    from mrmd_b200 import slabs

    gmin, gmax = np.array([-3.0, 0, 0]), np.array([997.0, 10, 10])
    for n in (2, 3, 4, 8):
        edges = [slabs.slab_bounds(gmin, gmax, r, n) for r in range(n)]
        assert edges[0][0] == gmin[0] and edges[-1][1] == gmax[0]
        for a, b in zip(edges[:-1], edges[1:]):
            assert a[1] == b[0]


def test_bind_host_to_gpu_is_a_noop_without_topology():
    """no CUDA device / no sysfs entry: the helper leaves the affinity alone and says so"""
    from mrmd_b200 import slabs

    before = os.sched_getaffinity(0)
    assert slabs.bind_host_to_gpu(0) is None
    assert os.sched_getaffinity(0) == before
