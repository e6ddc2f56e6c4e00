#!/bin/bash
# One GPU-box visit: GPU tests, smoke, bench, ncu launch list + one full capture of the top kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [noncu]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 600 python tools/bench_scoring.py > gpurun_out/${TAG}_scoring_bench.json 2> gpurun_out/${TAG}_scoring_bench.err; echo "scoring bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_scoring_bench.json
for MODE in ${MODES:-2}; do
  XVB_GEMM_CTA=$MODE timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu_cta$MODE.log 2>&1; echo "pytest(cta=$MODE) rc=$?"
  tail -12 gpurun_out/${TAG}_pytest_gpu_cta$MODE.log
  XVB_GEMM_CTA=$MODE timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cta$MODE.json 2> gpurun_out/${TAG}_bench_cta$MODE.err; echo "bench(cta=$MODE) rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_cta$MODE.json").read().strip().splitlines()[-1])
    print("cta=$MODE value %.3e e2e %.3e ms/step %.3f gemm_frac_exec %.3f pool_frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["executed_frac"], d["roofline_stats_pool"]["frac"]))
    print(d["kernel_ms"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench_cta$MODE.err").read()[-2000:])
PY
done
timeout 300 python tools/bench_ecapa.py 10 > gpurun_out/${TAG}_ecapa_bench.json 2> gpurun_out/${TAG}_ecapa_bench.err; echo "ecapa bench rc=$?"; cat gpurun_out/${TAG}_ecapa_bench.json; tail -3 gpurun_out/${TAG}_ecapa_bench.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ -z "$2" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 64 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tdnn_gemm -s 12 -c 6 -o gpurun_out/${TAG}_gemm python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:stats_pool -s 2 -c 1 -o gpurun_out/${TAG}_pool python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_pool.log 2>&1; echo "ncu pool rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tdnn_gemm -c 2 -o gpurun_out/${TAG}_hist python tools/hist_once.py > gpurun_out/${TAG}_ncu_hist.log 2>&1; echo "ncu hist rc=$?"
fi
ls -la gpurun_out | tail -24
