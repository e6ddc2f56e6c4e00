#!/bin/bash
TAG=${1:-ab}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" XVB_GEMM_WIDE=0 timeout 300 python tools/layer_times.py > gpurun_out/${TAG}_$name.txt 2>&1; echo "$name: $(tail -1 gpurun_out/${TAG}_$name.txt)"; }
run base XVB_GEMM_DEBUG=0
run no_store_issue XVB_GEMM_DEBUG=4
run no_ldtm XVB_GEMM_DEBUG=8
run no_store_no_ldtm XVB_GEMM_DEBUG=12
run no_store_no_ldtm_no_mma XVB_GEMM_DEBUG=14
run direct_nomma XVB_GEMM_DEBUG=2 XVB_GEMM_STORE=direct
