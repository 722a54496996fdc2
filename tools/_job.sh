mkdir -p gpurun_out
t0=$(date +%s)
timeout 1700 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/k6_alltests.log 2>&1; echo "alltests rc=$?"; tail -16 gpurun_out/k6_alltests.log
echo "alltests took $(( $(date +%s) - t0 )) s"
for cam in jax orbit; do timeout 300 python tests/gpu_profile_case.py --iters 40 --stages --time --cam $cam > gpurun_out/k6_st_$cam.log 2>&1; echo "$cam: $(tail -2 gpurun_out/k6_st_$cam.log)"; done
timeout 300 python tests/gpu_profile_case.py --iters 12 --stages --time --P 5000000 > gpurun_out/k6_st_dense.log 2>&1; echo "dense: $(tail -1 gpurun_out/k6_st_dense.log)"
