"""Golden vectors from the reference's own PYTHON code (importable in the CPU container; /root/reference is
absent on the GPU box, so the vectors are committed):
  ssim_ref.npz  utils/loss_utils.py::ssim (conv2d SSIM, :33-63) value map mean + autograd gradient w.r.t. img1
  sh_ref.npz    utils/sh_utils.py::eval_sh (:57-112) for degrees 0..3
Run:  python tests/golden/make_python_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SFGS_REFERENCE", "/root/reference")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    lu = load(os.path.join(REF, "utils", "loss_utils.py"), "ref_loss_utils")
    sh = load(os.path.join(REF, "utils", "sh_utils.py"), "ref_sh_utils")
    rng = np.random.default_rng(2024)
    # ---- SSIM: B=2, C=3, 70x53 (not multiples of the tile sizes), correlated images
    a = rng.uniform(0, 1, size=(2, 3, 70, 53)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.2, size=a.shape), 0, 1).astype(np.float32)
    ta = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    tb = torch.tensor(b, dtype=torch.float64)
    win = lu.create_window(11, 3).double()
    val = lu._ssim(ta, tb, win, 11, 3, True)
    val.backward()
    np.savez_compressed(os.path.join(HERE, "ssim_ref.npz"), img1=a, img2=b, mean=np.float64(val.item()),
                        grad=ta.grad.numpy())
    # ---- SH evaluation
    dirs = rng.normal(size=(512, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    coef = rng.normal(0, 0.5, size=(512, 3, 16))
    out = {"dirs": dirs.astype(np.float32), "coef": coef.astype(np.float32)}
    for deg in range(4):
        out[f"deg{deg}"] = sh.eval_sh(deg, torch.tensor(out["coef"], dtype=torch.float64),
                                      torch.tensor(out["dirs"], dtype=torch.float64)).numpy()
    np.savez_compressed(os.path.join(HERE, "sh_ref.npz"), **out)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
