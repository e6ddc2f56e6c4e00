#!/bin/bash
# N-GPU visit (gpurun --gpus N): the scaling bench line with the all-gather and the sharded all-pairs / PLDA back end.
# Usage: bash tools/gpu_ngpu.sh <tag> <N> [steps]
TAG=${1:-r03n}; N=${2:-2}; STEPS=${3:-10}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench ${N}gpu rc=$?"; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("value %.4e e2e %.4e ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    print("phases", d["phases_ms"]); print("clocks", d["clocks"]); print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("api",)})
    print("c4", d["config4"]); print("c3", {k: v for k, v in d["config3_ecapa"].items() if k != "roofline"}); print("c5", d["config5"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_ref_${N}gpu.json 2> gpurun_out/${TAG}_ref_${N}gpu.err; echo "reference arm rc=$?"; tail -1 gpurun_out/${TAG}_ref_${N}gpu.json | cut -c1-300
