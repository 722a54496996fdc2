"""Screen-space (tile-row) sharding of ONE frame across ranks — SURVEY.md §8e, BASELINE.json configs[3].

Nothing like this exists in the reference (its only multi-GPU mode is one scene per process,
scripts/run_jax.py:52-87).  Design:

* every rank holds the whole scene (or at least every Gaussian that can reach its band);
* rank g owns a contiguous band of tile rows [cuts[g], cuts[g+1]) chosen so that the number of tile
  instances per band is balanced (`partition_rows` on the per-row histogram of a previous frame);
* forward: each rank runs the normal pipeline restricted to its band (`tile_rows=` of
  sfgs.rasterizer.rasterize_gaussians), then ONE all_gather of the 8 image planes of the bands
  (padded to the tallest band) assembles the frame on every rank;
* backward: each rank runs the blend adjoint of its band only (`phase=1` of
  sfgs.rasterizer.rasterize_gaussians_backward).  The per-Gaussian adjoint is linear in the 14 blend-adjoint
  sums, so the ranks exchange those ([P,16] floats, 64 B/Gaussian) instead of the final gradients
  ((14+3M)·4 = 248 B/Gaussian at M=16): ONE reduce_scatter(sum) hands every rank the complete sums of its
  slice of Gaussians (`gaussian_slices`), and it finishes only that slice (`phase=2`) — the sharded-optimizer
  layout.  `reduce_gradients` (all_reduce of the final gradients, replicated parameters) remains as the simple
  alternative.

The host logic (partitioning, padded gather/assembly, reduction) is backend-agnostic and covered by
world-size-2 gloo tests on CPU tensors; on B200s the backend is NCCL over NVLink/NVSwitch.
"""
from __future__ import annotations

import os
import sys
from typing import List, Sequence

import torch
import torch.distributed as dist

TILE = 16


def tile_rows(height: int) -> int:
    return (height + TILE - 1) // TILE


def partition_rows(row_weights: Sequence[float], world: int) -> List[int]:
    """Cut `len(row_weights)` tile rows into `world` contiguous bands of (nearly) equal total weight.
    Returns world+1 monotone cut positions, cuts[0] = 0, cuts[-1] = n_rows; every band gets >= 1 row when
    n_rows >= world.  Deterministic: every rank computes the same cuts from the same histogram."""
    n = len(row_weights)
    if world <= 0:
        raise ValueError("world must be positive")
    w = [max(float(x), 0.0) + 1e-9 for x in row_weights]
    total = sum(w)
    cuts = [0]
    acc, r = 0.0, 0
    for g in range(1, world):
        target = total * g / world
        while r < n and acc + w[r] * 0.5 < target:
            acc += w[r]
            r += 1
        lo = cuts[-1] + 1 if n >= world else cuts[-1]          # at least one row per band when possible
        hi = n - (world - g) if n >= world else n
        cuts.append(min(max(r, lo), max(hi, lo)) if n >= world else min(r, n))
        # keep the running sum consistent with the cut actually taken
        acc = sum(w[:cuts[-1]])
        r = cuts[-1]
    cuts.append(n)
    return cuts


def band_pixel_rows(cuts: Sequence[int], rank: int, height: int):
    y0 = min(cuts[rank] * TILE, height)
    y1 = min(cuts[rank + 1] * TILE, height)
    return y0, y1


def gather_bands(local_planes: torch.Tensor, cuts: Sequence[int], height: int, group=None) -> torch.Tensor:
    """local_planes: [C, H, W] with only this rank's band rows valid. Returns the assembled [C, H, W].
    One all_gather of bands padded to the tallest band (bands differ in height after balancing)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    C, H, W = local_planes.shape
    assert H == height
    heights = [band_pixel_rows(cuts, g, height)[1] - band_pixel_rows(cuts, g, height)[0] for g in range(world)]
    hmax = max(max(heights), 1)
    y0, y1 = band_pixel_rows(cuts, rank, height)
    send = torch.zeros((C, hmax, W), dtype=local_planes.dtype, device=local_planes.device)
    send[:, : y1 - y0] = local_planes[:, y0:y1]
    recv = torch.empty((world, C, hmax, W), dtype=local_planes.dtype, device=local_planes.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    out = torch.empty_like(local_planes)
    for g in range(world):
        a, b = band_pixel_rows(cuts, g, height)
        out[:, a:b] = recv[g, :, : b - a]
    return out


def reduce_gradients(grads: Sequence[torch.Tensor], group=None) -> None:
    """Sum the per-band partial per-Gaussian gradients over ranks, in place, with one flat all_reduce."""
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def gaussian_slices(P: int, world: int) -> List[int]:
    """Equal slices of the Gaussian index range for the per-Gaussian adjoint: world+1 boundaries, slice length
    ceil(P/world) (the last one may be shorter or empty)."""
    per = (P + world - 1) // world
    return [min(g * per, P) for g in range(world + 1)]


def reduce_scatter_sums(acc: torch.Tensor, group=None) -> torch.Tensor:
    """acc: this rank's partial [P,16] blend-adjoint sums.  Returns the complete sums of this rank's slice
    ([ceil(P/world),16]; rows past P are zero).  NCCL: one reduce_scatter; backends without it (gloo): all_reduce
    + slice."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    P = acc.shape[0]
    per = (P + world - 1) // world
    if per * world != P:
        acc = torch.cat([acc, acc.new_zeros((per * world - P, acc.shape[1]))], 0)
    if dist.get_backend(group) == "nccl":
        out = torch.empty((per, acc.shape[1]), dtype=acc.dtype, device=acc.device)
        dist.reduce_scatter_tensor(out, acc.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out
    full = acc.clone()
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full[rank * per:(rank + 1) * per].contiguous()


def row_histogram(ranges: torch.Tensor, tiles_x: int) -> List[int]:
    """Instances per tile row from the image buffer's `ranges` [tiles,2]."""
    cnt = (ranges[:, 1] - ranges[:, 0]).long().view(-1, tiles_x).sum(1)
    return [int(v) for v in cnt.tolist()]


# ----------------------------------------------------------------------------- bench leg (NCCL, one frame sharded)
def bench_tilerows(args, rank, world, dev, steps, warmup, metric):
    """`bench.py --shard tilerows`: one JSON line for the tile-row sharded frame (strong scaling)."""
    import json
    rec = run_tilerows(args.P, 256.0, rank, world, dev, steps, warmup)
    if rank == 0:
        print(json.dumps({"metric": metric, "value": rec["value"], "unit": "Mpix/s", "n_gpus": world, "steps": steps,
                          "warmup": warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": rec["workload"], "P": rec["P"], "parallelism": rec["parallelism"],
                                     "cuts": rec["cuts"], "l2": "not flushed (collectives in the loop)"}}))
    dist.destroy_process_group()


def run_tilerows(P, extent, rank, world, dev, steps, warmup):
    """Time `steps` forward+backward passes of ONE 1920x1080 frame of the `P`-Gaussian scene split by tile rows over the
    ranks of the default process group (every rank must call this).  Returns a dict (identical on every rank):
    value [Mpix/s], ms_per_step (device time, max over ranks), the exchange mechanisms used and the band cuts."""
    import numpy as np

    from . import rasterizer as R
    from . import synthetic as S
    scene = S.city_scene(P, seed=0, sh_degree=3, extent=extent)
    cam = S.jax004_camera(1920, 1080)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d = dict(means3D=t(scene.means3D), scales=t(scene.scales), rotations=t(scene.rotations),
             opacities=t(scene.opacities), shs=t(scene.shs), view=t(cam.viewmatrix), proj=t(cam.projmatrix),
             campos=t(cam.campos), bg=torch.zeros(3, device=dev))
    cot = [t(c) for c in S.cotangents(cam.width, cam.height, seed=1)]
    e = torch.empty(0, device=dev)
    H, W = cam.height, cam.width
    rows, tiles_x = tile_rows(H), (W + TILE - 1) // TILE

    last = {"R": 0}

    def fwd(band):
        return R.rasterize_gaussians(d["bg"], d["means3D"], e, d["opacities"], d["scales"], d["rotations"], 1.0, e, e,
                                     e, 0, d["view"], d["proj"], cam.tanfovx, cam.tanfovy, 0.1, H, W, d["shs"], 3,
                                     d["campos"], False, False, tile_rows=band)

    # setup (untimed): one full-frame pass for the per-row histogram -> balanced cuts, identical on every rank
    f0 = fwd(None)
    from . import native as N
    import ctypes as C
    iv = N.ImageView()
    N.check(N.lib().sfgs_image_layout(f0[9].data_ptr(), W, H, C.byref(iv)), "image_layout")
    off = iv.ranges - f0[9].data_ptr()
    ranges = f0[9][off:off + rows * tiles_x * 8].view(torch.int32).view(-1, 2)
    cuts = partition_rows(row_histogram(ranges, tiles_x), world)
    band = (cuts[rank], cuts[rank + 1])

    sl = gaussian_slices(scene.P, world)
    per = sl[1] - sl[0]
    acc_buf = torch.empty((per * world, 16), dtype=torch.float32, device=dev)   # padded to equal slices
    acc_buf[scene.P:].zero_()

    def bwd(f, **kw):
        return R.rasterize_gaussians_backward(d["bg"], d["means3D"], f[5], e, d["scales"], d["rotations"], e, 1.0, e, e,
                                              d["view"], d["proj"], cam.tanfovx, cam.tanfovy, 0.1, cot[0], cot[1],
                                              cot[2], cot[3], e, d["shs"], 3, d["campos"], f[7], f[0], f[8], f[9], f[4],
                                              False, tile_rows=band, **kw)

    def step_nccl():
        f = fwd(band)
        last["R"] = f[0]
        planes = torch.cat([f[1], f[2], f[4], f[3]], 0)            # colour, depth, alpha, normal: 8 planes
        full = gather_bands(planes, cuts, H)
        bwd(f, phase=1, acc=acc_buf[:scene.P])                      # blend adjoint of this band -> partial sums
        mine = torch.empty((per, 16), dtype=torch.float32, device=dev)
        dist.reduce_scatter_tensor(mine, acc_buf, op=dist.ReduceOp.SUM)
        g = bwd(f, phase=2, acc=mine, gauss_range=(sl[rank], sl[rank + 1]))   # finish this rank's slice only
        return full, g

    # Fused alternative: every rank's accumulator slice is symmetric (peer-mapped) memory and the blend adjoint's
    # vector reductions go straight to the owning rank over NVLink — the reduce-scatter happens inside the kernel,
    # tile by tile, and no collective follows it.  Falls back to the NCCL path when peer mapping is unavailable.
    collective, step = "nccl reduce_scatter of the [P,16] blend-adjoint sums", step_nccl
    gather = "image all_gather (NCCL)"

    def step_single():          # world == 1: the same frame, the same loop, no exchange — the strong-scaling baseline
        f = fwd(None)
        last["R"] = f[0]
        return f, bwd_full(f)

    def bwd_full(f):
        return R.rasterize_gaussians_backward(d["bg"], d["means3D"], f[5], e, d["scales"], d["rotations"], e, 1.0, e, e,
                                              d["view"], d["proj"], cam.tanfovx, cam.tanfovy, 0.1, cot[0], cot[1],
                                              cot[2], cot[3], e, d["shs"], 3, d["campos"], f[7], f[0], f[8], f[9], f[4],
                                              False)
    if world == 1:
        collective, gather, step = "none (one GPU)", "none (one GPU)", step_single
    elif os.environ.get("SFGS_PEER_REDUCE", "1") != "0":
        try:
            import torch.distributed._symmetric_memory as symm
            acc_sym = symm.empty((per * 16,), dtype=torch.float32, device=dev)
            hdl = symm.rendezvous(acc_sym, dist.group.WORLD)
            peers = [int(p) for p in hdl.buffer_ptrs]

            # forward: the frame block of every rank is symmetric memory too; the blend kernel stores its band into
            # all of them (the image all-gather, fused), and one barrier publishes the frame
            frame_sym = symm.empty((8 * H * W,), dtype=torch.float32, device=dev)
            fh = symm.rendezvous(frame_sym, dist.group.WORLD)
            frame_peers = [int(p) for p in fh.buffer_ptrs]

            def step_peer():
                # Two device-side barriers per step.  (B) after the forward: every band has landed in every rank's frame
                # block AND every rank's accumulator slice is clear (it is zeroed before the forward, in stream order
                # after this rank's previous phase 2).  (C) after the blend adjoint: every rank's reductions have
                # landed.  No barrier is needed before the forward: a rank passes (C) of the previous step only after
                # its loss and blend adjoint of that step — the last readers of the previous frame — have completed.
                acc_sym.zero_()
                f = R.rasterize_gaussians(d["bg"], d["means3D"], e, d["opacities"], d["scales"], d["rotations"], 1.0,
                                          e, e, e, 0, d["view"], d["proj"], cam.tanfovx, cam.tanfovy, 0.1, H, W,
                                          d["shs"], 3, d["campos"], False, False, tile_rows=band,
                                          out_planes=frame_sym, out_peers=frame_peers)
                last["R"] = f[0]
                hdl.barrier(channel=0)                              # (B)
                full = frame_sym.view(8, H, W)
                bwd(f, phase=1, acc_peers=peers, peer_slice=per)    # reductions land on the owners' slices
                hdl.barrier(channel=1)                              # (C)
                g = bwd(f, phase=2, acc=acc_sym.view(per, 16), gauss_range=(sl[rank], sl[rank + 1]))
                return full, g

            # agreement with the NCCL path on this rank's slice, before it is trusted
            full_ref, g_ref = step_nccl()
            full_new, g_new = step_peer()
            torch.cuda.synchronize(dev)
            worst = 0.0 if torch.equal(full_ref, full_new) else float("inf")   # the frames must agree bit for bit
            for a_, b_ in zip(g_ref, g_new):
                if a_.numel():
                    x, y = a_[sl[rank]:sl[rank + 1]], b_[sl[rank]:sl[rank + 1]]
                    worst = max(worst, float(((x - y).abs().max() / (y.abs().max() + 1e-12)).item()))
            ok = torch.tensor([1.0 if worst <= 1e-3 else 0.0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 1.0:
                gather = "image all-gather fused into the blend kernel (peer stores)"
                collective, step = ("reduce-scatter fused into the blend adjoint: system-scope vector reductions to "
                                    f"peer-mapped slices over NVLink (max rel. deviation from the NCCL path {worst:.1e})"), step_peer
            elif rank == 0:
                print(f"[tilerows] peer path disagrees with the NCCL path (rel {worst:.2e}); using NCCL", file=sys.stderr)
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                print(f"[tilerows] symmetric memory unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)

    multi = world > 1
    import time

    def timed_pass(nwarm):
        """One pass of `steps` steps: CUDA events around the WHOLE loop (what is reported) and around every step (only to
        tell whether the pass was disturbed from outside).  Returns (loop ms max over ranks, clean?, host issue ms/step)."""
        for _ in range(nwarm):
            step()
        if multi:
            dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per = []
        a.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            per.append((e0, e1))
        t_issue = (time.perf_counter() - t0) * 1e3 / steps    # host time to ISSUE a step (incl. the forward's one wait)
        b.record()
        if multi:
            dist.barrier()
        torch.cuda.synchronize(dev)
        loop = a.elapsed_time(b)
        fastest = min(x.elapsed_time(y) for x, y in per)
        # the same rule as bench.py's headline: a pass whose mean step exceeds 1.05x its fastest step was disturbed (the
        # boxes are shared: a descheduled host thread leaves the GPU idle inside the loop) and is measured again
        v = torch.tensor([loop, 1.0 if loop / steps <= 1.05 * fastest else 0.0], dtype=torch.float64, device=dev)
        if multi:
            lo = v.clone()
            dist.all_reduce(v, op=dist.ReduceOp.MAX)           # loop time: max over ranks
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)          # clean only if clean on every rank
            v[1] = lo[1]
        return float(v[0].item()), bool(v[1].item() == 1.0), t_issue

    passes = []
    for k in range(6):
        # untimed warm-up: a fixed count (the ranks must stay in lock step: the fused step contains device-side barriers),
        # long enough to outlast the slow first 0.1-0.3 s of a freshly set-up loop on these boxes (see bench.py settle())
        passes.append(timed_pass(max(warmup, 200) if k == 0 else 20))
        if passes[-1][1]:
            break
    clean = [p_ for p_ in passes if p_[1]]
    total, _, t_host = clean[0] if clean else min(passes, key=lambda p_: p_[0])
    R_band = torch.tensor([float(last["R"])], dtype=torch.float64, device=dev)
    R_all = [torch.zeros_like(R_band) for _ in range(world)]
    if multi:
        dist.all_gather(R_all, R_band)
    else:
        R_all = [R_band]
    # this rank's kernel times per stage (untimed extra steps; the library's own CUDA events), to show what shrinks with N
    per_stage = {}
    for _ in range(3):
        N.profile_enable(True)
        step()
        torch.cuda.synchronize(dev)
        for k, v in N.profile_read().items():
            per_stage.setdefault(k, []).append(v[0] / v[1] if v[1] else 0.0)
    N.profile_enable(False)
    stages = {k: round(sorted(v)[len(v) // 2], 4) for k, v in per_stage.items()}
    if multi:
        dist.barrier()
    return {"stages_ms_rank0": stages, "kernel_ms_rank0": round(sum(stages.values()), 4), "host_issue_ms_rank0": round(t_host, 4),"workload": f"one 1920x1080 frame of the {scene.P}-Gaussian scene (extent {extent:g}) sharded by tile rows",
            "P": scene.P, "value": round(steps * H * W / (total / 1e3) / 1e6, 2), "ms_per_step": round(total / steps, 4),
            "steps": steps, "warmup": warmup,
            "parallelism": f"tilerows x{world}: {gather} + {collective}; each rank finishes the gradients of P/N Gaussians",
            "gather": gather, "collective": collective, "cuts": cuts,
            "instances_per_band": [int(x.item()) for x in R_all],
            "timing": "CUDA events around the whole loop of `steps` steps, max over ranks; a pass whose mean step exceeds 1.05x its "
                      "fastest step (per-step events) is measured again, <= 6 passes",
            "passes_ms_per_step": [round(p_[0] / steps, 4) for p_ in passes], "clean_pass_found": bool(clean)}
