#!/bin/bash
# Four-GPU visit (gpurun --gpus 4): scaling bench + sharded extraction / all-pairs EER demo.
TAG=${1:-r02i}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_4gpu.json 2> gpurun_out/${TAG}_bench_4gpu.err; echo "bench 4gpu rc=$?"; tail -1 gpurun_out/${TAG}_bench_4gpu.json | cut -c1-400; tail -2 gpurun_out/${TAG}_bench_4gpu.err
timeout 300 $TR --master-port 29523 tools/multi_gpu_demo.py > gpurun_out/${TAG}_demo.log 2>&1; echo "demo rc=$?"; grep multi_gpu_demo gpurun_out/${TAG}_demo.log
