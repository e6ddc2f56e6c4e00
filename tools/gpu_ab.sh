#!/bin/bash
# Scratch GPU visit: F-TDNN blueprint tests; memcheck over the tcgen05 layer / pooling / Res2Net / extractor tests.
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "factored or snowdar" > gpurun_out/${TAG}_pytest_ftdnn.log 2>&1; echo "pytest ftdnn rc=$?"; tail -12 gpurun_out/${TAG}_pytest_ftdnn.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tdnn_gemm_vs_oracle or fused_pooling or zero_padding or stats_pool or im2col or edge_lengths" > gpurun_out/${TAG}_memcheck_gemm.log 2>&1; echo "memcheck gemm rc=$?"; grep -h "passed\|failed\|ERROR SUMMARY" gpurun_out/${TAG}_memcheck_gemm.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ecapa.py -m gpu -q -x -k "res2net or se_apply or attn or native" > gpurun_out/${TAG}_memcheck_ecapa.log 2>&1; echo "memcheck ecapa rc=$?"; grep -h "passed\|failed\|ERROR SUMMARY" gpurun_out/${TAG}_memcheck_ecapa.log | tail -3
