#!/bin/bash
# One GPU-box visit: GPU tests, smoke, bench, ncu launch list + one full capture of the top kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_ncu]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ -z "$2" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 64 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdnn_gemm -s 12 -c 3 -o gpurun_out/${TAG}_gemm python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:stats_pool -s 2 -c 1 -o gpurun_out/${TAG}_pool python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_pool.log 2>&1; echo "ncu pool rc=$?"
fi
ls -la gpurun_out | tail -20
