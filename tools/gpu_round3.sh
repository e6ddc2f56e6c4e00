#!/bin/bash
# GPU visit: full GPU test suite with durations, then the bench with lanes on and off.
TAG=${1:-r04}; shift
mkdir -p gpurun_out
for WHAT in "$@"; do
case $WHAT in
tests)
  timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${TAG}_pytest_gpu.log ;;
newtests)
  timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ecapa.py -m gpu -q -x --durations=8 -k "snowdar or attn_head or shard or lanes" > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -20 gpurun_out/${TAG}_pytest_new.log ;;
ab)
  for L in 1 0; do
    XVB_LANES=$L XVB_BENCH_UTTS=32768 XVB_BENCH_ECAPA_UTTS=8192 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_lanes$L.json 2> gpurun_out/${TAG}_bench_lanes$L.err; echo "lanes=$L rc=$?"; tail -3 gpurun_out/${TAG}_bench_lanes$L.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_lanes$L.json").read().strip().splitlines()[-1])
    print("lanes=$L value %.4e e2e %.4e ms/step %.2f | ecapa %.4e e2e %.4e | clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config3_ecapa"]["value"], d["config3_ecapa"]["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("parse failed", e)
PY
  done ;;
esac
done
