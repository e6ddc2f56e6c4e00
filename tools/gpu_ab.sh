#!/bin/bash
# Scratch GPU visit: all GPU tests, bench A/B of the im2col first layer + split-K, scoring bench.
TAG=${1:-r01v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -25 gpurun_out/${TAG}_pytest_gpu.log
for CFG in "1 1" "0 0" "1 0" "0 1"; do
  set -- $CFG
  XVB_IM2COL=$1 XVB_SPLITK=$2 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_i$1_s$2.json 2> gpurun_out/${TAG}_bench_i$1_s$2.err; echo "bench(im2col=$1 splitk=$2) rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_i$1_s$2.json").read().strip().splitlines()[-1])
    print("im2col=$1 splitk=$2 value %.4e e2e %.4e ms/step %.4f link %.1f GB/s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"].get("h2d_link_gbs_measured", 0)))
    print({k: round(v*1e3,1) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench_i$1_s$2.err").read()[-1500:])
PY
done
cp gpurun_out/${TAG}_bench_i1_s1.json gpurun_out/${TAG}_bench.json
timeout 900 python tools/bench_scoring.py > gpurun_out/${TAG}_scoring_bench.json 2> gpurun_out/${TAG}_scoring_bench.err; echo "scoring bench rc=$?"; cat gpurun_out/${TAG}_scoring_bench.json; tail -5 gpurun_out/${TAG}_scoring_bench.err
timeout 300 python tools/bench_ecapa.py 10 > gpurun_out/${TAG}_ecapa_bench.json 2> gpurun_out/${TAG}_ecapa_bench.err; echo "ecapa bench rc=$?"; cat gpurun_out/${TAG}_ecapa_bench.json; tail -3 gpurun_out/${TAG}_ecapa_bench.err
