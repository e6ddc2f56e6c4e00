#!/bin/bash
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/bench_scoring.py > gpurun_out/${TAG}_scoring.json 2> gpurun_out/${TAG}_scoring.err; echo "scoring rc=$?"; cat gpurun_out/${TAG}_scoring.json; tail -5 gpurun_out/${TAG}_scoring.err
timeout 300 python tools/bench_ecapa.py 10 > gpurun_out/${TAG}_ecapa.json 2>&1; tail -1 gpurun_out/${TAG}_ecapa.json
