#!/bin/bash
# Scratch GPU visit: native ECAPA extractor (tests + bench native vs Python twin) and the newest tests.
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ecapa.py tests/test_gpu_deploy.py tests/test_gpu_scoring.py tests/test_gpu_plda_train.py -m gpu -q > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest_new.log
timeout 300 python tools/bench_ecapa.py 10 > gpurun_out/${TAG}_ecapa_native.json 2> gpurun_out/${TAG}_ecapa_native.err; echo "ecapa native rc=$?"; cat gpurun_out/${TAG}_ecapa_native.json; tail -3 gpurun_out/${TAG}_ecapa_native.err
XVB_ECAPA_NATIVE=0 timeout 300 python tools/bench_ecapa.py 10 > gpurun_out/${TAG}_ecapa_python.json 2> gpurun_out/${TAG}_ecapa_python.err; echo "ecapa python rc=$?"; cat gpurun_out/${TAG}_ecapa_python.json
