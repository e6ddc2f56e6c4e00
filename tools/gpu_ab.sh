#!/bin/bash
# Scratch GPU visit: new tests (histogram, deploy, fbank) + scoring bench.
TAG=${1:-r01u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trial_histogram.py tests/test_gpu_deploy.py tests/test_gpu_fbank.py -m gpu -q > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -40 gpurun_out/${TAG}_pytest_new.log
timeout 900 python tools/bench_scoring.py > gpurun_out/${TAG}_scoring_bench.json 2> gpurun_out/${TAG}_scoring_bench.err; echo "scoring bench rc=$?"; cat gpurun_out/${TAG}_scoring_bench.json; tail -5 gpurun_out/${TAG}_scoring_bench.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
