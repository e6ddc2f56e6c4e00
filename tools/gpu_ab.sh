#!/bin/bash
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tdnn_gemm_vs_oracle or fused_pooling or zero_padding or stats_pool or im2col or edge_lengths or split_frames" > gpurun_out/${TAG}_memcheck_gemm.log 2>&1; echo "memcheck gemm rc=$?"; grep -h "passed\|failed\|ERROR SUMMARY" gpurun_out/${TAG}_memcheck_gemm.log | tail -3
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "stats_pool or full_size" > gpurun_out/${TAG}_pytest_pool.log 2>&1; echo "pytest pool rc=$?"; tail -3 gpurun_out/${TAG}_pytest_pool.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e pool frac %.3f'%(d['value'], d['e2e']['value'], d['roofline_stats_pool']['frac']))"
