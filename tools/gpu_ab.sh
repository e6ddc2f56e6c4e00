#!/bin/bash
# A/B bench of GEMM knobs in one GPU visit.  Usage: bash tools/gpu_ab.sh <tag>
TAG=${1:-ab}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().splitlines()[-1])
    print("$name: value %.3e e2e %.3e ms/step %.3f exec_frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["executed_frac"]), {k: round(v*1e3) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/${TAG}_$name.err").read()[-1500:])
PY
}
run direct XVB_GEMM_STORE=direct
run tma XVB_GEMM_STORE=tma
run direct_bn128 XVB_GEMM_STORE=direct XVB_GEMM_BN=128
run direct_cta1 XVB_GEMM_STORE=direct XVB_GEMM_CTA=1
XVB_GEMM_STORE=direct timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/bench_ecapa.py 10
