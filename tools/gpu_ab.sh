#!/bin/bash
TAG=${1:-ab}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/layer_times.py > gpurun_out/${TAG}_$name.txt 2>&1; echo "$name: $(tail -1 gpurun_out/${TAG}_$name.txt)"; }
run narrow XVB_GEMM_WIDE=0
run wide XVB_GEMM_WIDE=1
run narrow_skipmma XVB_GEMM_WIDE=0 XVB_GEMM_DEBUG=2
XVB_GEMM_WIDE=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
