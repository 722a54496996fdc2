"""Host logic of the tile-row sharding on CPU: partitioning, and world-size-2 gloo runs of the padded band
gather and the flat gradient reduction (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfgs import multigpu as MG


def test_partition_rows_balanced_and_total():
    w = [0, 0, 5, 50, 80, 80, 60, 10, 1, 0, 0, 0]
    for world in (1, 2, 3, 4, 8):
        cuts = MG.partition_rows(w, world)
        assert cuts[0] == 0 and cuts[-1] == len(w) and len(cuts) == world + 1
        assert all(b > a for a, b in zip(cuts, cuts[1:]))          # every band non-empty (12 rows >= world)
    cuts = MG.partition_rows(w, 2)
    left, right = sum(w[:cuts[1]]), sum(w[cuts[1]:])
    assert abs(left - right) <= max(w)
    # fewer rows than ranks: still monotone, covers everything
    cuts = MG.partition_rows([3, 1], 4)
    assert cuts[0] == 0 and cuts[-1] == 2 and all(b >= a for a, b in zip(cuts, cuts[1:]))
    # uniform weights -> equal split
    assert MG.partition_rows([1] * 68, 4) == [0, 17, 34, 51, 68]
    assert MG.tile_rows(1080) == 68 and MG.band_pixel_rows([0, 34, 68], 1, 1080) == (544, 1080)
    assert MG.gaussian_slices(10, 4) == [0, 3, 6, 9, 10] and MG.gaussian_slices(2, 4) == [0, 1, 2, 2, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W, C = 100, 37, 8                          # 7 tile rows, last one partial
        cuts = [0, 5, 7]                              # unequal bands: 80 and 20 pixel rows
        truth = torch.arange(C * H * W, dtype=torch.float32).view(C, H, W)
        local = torch.full((C, H, W), float("nan"))
        y0, y1 = MG.band_pixel_rows(cuts, rank, H)
        local[:, y0:y1] = truth[:, y0:y1]
        full = MG.gather_bands(local, cuts, H)
        ok_gather = bool(torch.equal(full, truth))
        g1 = torch.full((5, 3), float(rank + 1))
        g2 = torch.full((5,), 10.0 * (rank + 1))
        MG.reduce_gradients([g1, g2])
        ok_reduce = bool(torch.all(g1 == 3.0)) and bool(torch.all(g2 == 30.0))
        # blend-adjoint sums: P = 7 Gaussians over 2 ranks -> slices [0,4) and [4,7); every rank ends up with the
        # complete sums of its own slice only (zero rows past P)
        P = 7
        acc = torch.arange(P * 16, dtype=torch.float32).view(P, 16) * (rank + 1)
        mine = MG.reduce_scatter_sums(acc)
        sl = MG.gaussian_slices(P, world)
        want = torch.zeros((4, 16))
        n = sl[rank + 1] - sl[rank]
        want[:n] = torch.arange(P * 16, dtype=torch.float32).view(P, 16)[sl[rank]:sl[rank + 1]] * 3.0
        ok_rs = sl == [0, 4, 7] and mine.shape == (4, 16) and bool(torch.equal(mine, want))
        out.put((rank, ok_gather, ok_reduce and ok_rs))
    finally:
        dist.destroy_process_group()


def test_gather_and_reduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "padded band all_gather did not reassemble the frame"
    assert all(r[2] for r in res), "flat gradient all_reduce mismatch"


def _renderset_worker(rank, world, port, q):
    import os
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist
    from sfgs import renderset as RS
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_views = 7
    cuts = RS.partition_views(n_views, world)
    H, W = 4, 6
    mine = range(cuts[rank], cuts[rank + 1])
    color = torch.stack([torch.full((3, H, W), v, dtype=torch.uint8) for v in mine]) if len(mine) else torch.empty((0, 3, H, W), dtype=torch.uint8)
    depth = torch.stack([torch.full((1, H, W), float(v) + 0.5) for v in mine]) if len(mine) else torch.empty((0, 1, H, W))
    c, z = RS.gather_views(color, depth, cuts)
    ok = c.shape == (n_views, 3, H, W) and all(int(c[v, 0, 0, 0]) == v and abs(float(z[v, 0, 0, 0]) - v - 0.5) < 1e-6 for v in range(n_views))
    q.put((rank, ok, cuts))
    dist.destroy_process_group()


def test_renderset_view_sharding_gathers_every_view_in_order():
    """the IDU render set shards by view; host logic (partition, padded gather, reassembly) on gloo."""
    import torch.multiprocessing as mp
    from sfgs import renderset as RS
    assert RS.partition_views(108, 8) == [0, 14, 28, 42, 56, 69, 82, 95, 108]
    assert RS.partition_views(3, 4) == [0, 1, 2, 3, 3]
    assert len(RS.idu_grid_targets()) == 9
    cams = RS.idu_orbit_cameras(RS.idu_grid_targets(), 85.0, 300.0, num_cams=6, num_samples=2, size=64)
    assert len(cams) == 108 and cams[0].width == 64 and cams[0] is cams[1] and cams[0] is not cams[2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_renderset_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
