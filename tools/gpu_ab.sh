#!/bin/bash
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e exec_frac %.3f pool_frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['executed_frac'], d['roofline_stats_pool']['frac'])); print({k: round(v*1e3) for k,v in d['kernel_ms'].items()})" || tail -20 gpurun_out/${TAG}_bench.err
