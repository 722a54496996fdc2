#!/bin/bash
# Usage (on the GPU box, from the repo root): bash tools/gpu_job.sh <tag> <step> [<step> ...]
# Every step writes its log under gpurun_out/<tag>_<step>.log; nothing here is a benchmark result by itself.
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    newtests)  timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "benched or needles or adjudicated or randomised or golden" > gpurun_out/${tag}_newtests.log 2>&1; echo "newtests rc=$?" ;;
    alltests)  timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${tag}_alltests.log 2>&1; echo "alltests rc=$?"; tail -5 gpurun_out/${tag}_alltests.log ;;
    bench)     timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json | head -c 3000 ;;
    benchref)  timeout 600 python bench.py --impl reference > gpurun_out/${tag}_benchref.json 2> gpurun_out/${tag}_benchref.err; echo "benchref rc=$?"; cat gpurun_out/${tag}_benchref.json | head -c 1500 ;;
    host)      timeout 300 python tools/host_overhead.py > gpurun_out/${tag}_host.log 2>&1; echo "host rc=$?"; head -40 gpurun_out/${tag}_host.log ;;
    launches)  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_launches.log 2>&1; echo "launches rc=$?" ;;
    ncufull)   timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess|tile_sort|gauss_bwd|scatter|tile_scan|acc_clear|ssim_|appearance_|filter3d" -s 14 -c 16 -o gpurun_out/${tag}_full -f python tests/gpu_profile_case.py --iters 2 --siblings > gpurun_out/${tag}_ncufull.log 2>&1; echo "ncufull rc=$?" ;;
    benchN)    n=${NGPU:-2}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err; echo "benchN($n) rc=$?"; head -c 2500 gpurun_out/${tag}_bench_n$n.json ;;
    benchrefN) n=${NGPU:-2}; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_benchref_n$n.json 2> gpurun_out/${tag}_benchref_n$n.err; echo "benchrefN($n) rc=$?"; head -c 600 gpurun_out/${tag}_benchref_n$n.json ;;
    sanitize)  for tool in memcheck racecheck synccheck; do timeout 600 compute-sanitizer --tool $tool python tests/gpu_sanitize_case.py > gpurun_out/${tag}_san_$tool.log 2>&1; echo "sanitize $tool rc=$? $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${tag}_san_$tool.log | tail -1)"; done ;;
    *)         echo "unknown step $step" ;;
  esac
done
