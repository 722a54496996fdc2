"""Per-step device / host time of the C-ABI fwd+bwd loop from a cold start, with the caching allocator's cudaMalloc count:
shows what the first passes of a fresh configuration cost (python tools/step_trace.py [P], needs a GPU)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "skyfall-gs_b200"))
import numpy as np
import torch

import bench as B
from sfgs import native
from sfgs import synthetic as S

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
flush_on = (sys.argv[2] != "noflush") if len(sys.argv) > 2 else True
dev = torch.device("cuda:0")
scene = S.city_scene(P, seed=0, sh_degree=3, extent=256.0 * (P / 1e6) ** 0.5 if P > 1_000_000 else 256.0)
cam = S.jax004_camera(1920, 1080)
d = B.make_inputs(scene, cam, dev)
flush = B.L2Flusher(dev)
native.lib()
evs, host = [], []
stats0 = torch.cuda.memory_stats(dev)
marks = []
for i in range(120):
    if flush_on:
        flush()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    B.step_ours(d, cam)
    b.record()
    host.append((time.perf_counter() - t0) * 1e3)
    evs.append((a, b))
    if i % 20 == 19:
        st = torch.cuda.memory_stats(dev)
        marks.append((st["num_device_alloc"], st["num_device_free"], native.lib().sfgs_overflow_reruns()))
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in evs]
for k in range(6):
    seg, hs = ms[20 * k:20 * k + 20], host[20 * k:20 * k + 20]
    print(f"P={P} flush={flush_on} steps {20*k:3d}-{20*k+19:3d}: device median {np.median(seg):.3f} max {max(seg):.3f} ms | host issue median "
          f"{np.median(hs):.3f} max {max(hs):.3f} ms | cudaMalloc so far {marks[k][0]} cudaFree {marks[k][1]} overflow reruns {marks[k][2]}")
