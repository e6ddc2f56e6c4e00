#!/bin/bash
# Scratch GPU visit: register-direct epilogue stores for few-K-block layers: tests + A/B.
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
for MODE in auto tma reg; do
  XVB_GEMM_STORE=$MODE timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_$MODE.json 2> gpurun_out/${TAG}_bench_$MODE.err; echo "bench($MODE) rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$MODE.json").read().strip().splitlines()[-1])
    print("$MODE value %.4e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v*1e3,1) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench_$MODE.err").read()[-1500:])
PY
  XVB_GEMM_STORE=$MODE timeout 300 python tools/bench_ecapa.py 10 2>/dev/null | tail -1 | cut -c1-200
done
XVB_GEMM_STORE=tma timeout 300 python tools/bench_scoring.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tma', {k:(round(v.get('ms',v.get('narrow_window_pass_ms',0)),2)) for k,v in d.items()})"
timeout 300 python tools/bench_scoring.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('auto', {k:(round(v.get('ms',v.get('narrow_window_pass_ms',0)),2)) for k,v in d.items()})"
