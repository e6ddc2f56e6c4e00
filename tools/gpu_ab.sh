#!/bin/bash
TAG=${1:-ab}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/layer_times.py > gpurun_out/${TAG}_$name.txt 2>&1; echo "$name: $(tail -1 gpurun_out/${TAG}_$name.txt)"; }
run wide XVB_GEMM_WIDE=1
run narrow XVB_GEMM_WIDE=0
run wide_skip_both XVB_GEMM_WIDE=1 XVB_GEMM_DEBUG=3
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/bench_ecapa.py 10
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e exec_frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['executed_frac']))"
