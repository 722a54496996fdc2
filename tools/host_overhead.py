"""Where does the HOST time of one public-API training step go?  (python tools/host_overhead.py, needs a GPU)

Prints, for the bench's e2e step, the wall time the host needs to ENQUEUE each part (no device sync inside the
loop except the forward's own wait for num_rendered) next to the device time of the whole step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "skyfall-gs_b200"))
import torch

import bench as B
from sfgs import synthetic as S

dev = torch.device("cuda:0")
scene = S.city_scene(1_000_000, seed=0, sh_degree=3)
cam = B.camera_for_rank(0, 1)
import diff_gauss
e = B.E2E(scene, cam, dev, diff_gauss)
for _ in range(5):
    e.step()
torch.cuda.synchronize()
n = 30
t_host, t_dev = [], []
for _ in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record(); e.step(); b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t_host.append((t1 - t0) * 1e3); t_dev.append(a.elapsed_time(b))
med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
print(f"e2e step (device idle at its start): host enqueue median {med(t_host):.3f} ms, device median {med(t_dev):.3f} ms")
# back-to-back steps, the way the bench times them: device time per step when the host runs ahead
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
    e.step()
b.record()
torch.cuda.synchronize()
print(f"e2e back to back: {a.elapsed_time(b) / n:.3f} ms per step")

import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    e.step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative")
st.print_stats(28)
