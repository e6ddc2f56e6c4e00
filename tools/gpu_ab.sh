#!/bin/bash
TAG=${1:-ab}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/layer_times.py > gpurun_out/${TAG}_$name.txt 2>&1; echo "$name: $(tail -1 gpurun_out/${TAG}_$name.txt)"; }
run pdl1 XVB_PDL=1
run pdl0 XVB_PDL=0
for P in 1 0; do XVB_PDL=$P timeout 300 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl=$P value %.4e e2e %.3e ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"; done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
