#!/bin/bash
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ecapa.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/bench_ecapa.py 10 --profile > gpurun_out/${TAG}_ecapa.txt 2>&1; tail -22 gpurun_out/${TAG}_ecapa.txt
XVB_ECAPA_RES2NET=gemm timeout 300 python tools/bench_ecapa.py 10 | tail -1
