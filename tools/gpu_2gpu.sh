#!/bin/bash
# Two-GPU visit (gpurun --gpus 2): scaling bench, reference arm under torchrun, sharded extraction + all-pairs EER demo.
TAG=${1:-r01y}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -1 gpurun_out/${TAG}_bench_2gpu.json | cut -c1-700; tail -3 gpurun_out/${TAG}_bench_2gpu.err
timeout 300 $TR --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/${TAG}_ref_2gpu.json 2> gpurun_out/${TAG}_ref_2gpu.err; echo "reference arm 2gpu rc=$?"; tail -1 gpurun_out/${TAG}_ref_2gpu.json | cut -c1-400
timeout 300 $TR --master-port 29513 tools/multi_gpu_demo.py > gpurun_out/${TAG}_demo.log 2>&1; echo "demo rc=$?"; grep multi_gpu_demo gpurun_out/${TAG}_demo.log
timeout 600 python -m pytest tests/test_gpu_plda_train.py tests/test_gpu_kernels.py -m gpu -q -k "plda or snowdar or im2col" > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -4 gpurun_out/${TAG}_pytest_new.log
