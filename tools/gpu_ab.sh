#!/bin/bash
# Scratch GPU visit: double-buffered sub-slab epilogue: all GPU tests + bench + ECAPA + scoring.
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.4e e2e %.4e ms/step %.4f exec_frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["executed_frac"]), {k: round(v*1e3,1) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
timeout 300 python tools/bench_ecapa.py 10 2>/dev/null | tail -1 | cut -c1-220
timeout 300 python tools/bench_scoring.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:(round(v.get('ms',v.get('narrow_window_pass_ms',0)),2)) for k,v in d.items()})"
