#!/bin/bash
# Scratch GPU visit: trial-histogram tests + scoring bench + bench sanity.
TAG=${1:-r01s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trial_histogram.py -m gpu -x -q > gpurun_out/${TAG}_pytest_hist.log 2>&1; echo "pytest hist rc=$?"; tail -25 gpurun_out/${TAG}_pytest_hist.log
timeout 600 python tools/bench_scoring.py > gpurun_out/${TAG}_scoring_bench.json 2> gpurun_out/${TAG}_scoring_bench.err; echo "scoring bench rc=$?"; cat gpurun_out/${TAG}_scoring_bench.json; tail -5 gpurun_out/${TAG}_scoring_bench.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
