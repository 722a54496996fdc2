#!/bin/bash
mkdir -p gpurun_out
echo "== no TMA"; SFGS_SSIM_TMA=0 CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_siblings.py -x -q -k "ssim" 2>&1 | tail -8
echo "== TMA"; CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_siblings.py -x -q -k "ssim" 2>&1 | tail -8
echo "== sanitizer TMA small"; timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_siblings.py -x -q -k "reference_cuda_kernel and 33" 2>&1 | grep -v "^$" | tail -25
