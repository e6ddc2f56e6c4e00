#!/bin/bash
# Scratch GPU visit: 64-column plane-at-a-time epilogue (128-byte store rows): tests + A/B.
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
for K in 1 0; do
  XVB_GEMM_BOX64=$K timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_box$K.json 2> gpurun_out/${TAG}_bench_box$K.err; echo "bench(box64=$K) rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_box$K.json").read().strip().splitlines()[-1])
    print("box64=$K value %.4e e2e %.4e ms/step %.4f exec_frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["executed_frac"]), {k: round(v*1e3,1) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench_box$K.err").read()[-1500:])
PY
  XVB_GEMM_BOX64=$K timeout 300 python tools/bench_ecapa.py 10 2>/dev/null | tail -1 | cut -c1-200
done
